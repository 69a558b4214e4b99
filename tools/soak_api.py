"""Randomised differential soak of the entry points AROUND the fused step, against the CPU oracle, for MSJ, MSJ-shaped
robots with other limits and other robots:
  * stand-alone compute_reward / _did_reach_goal (`roboy_compute_reward`) on states and goals salted with NaN, +-inf, 0,
    the bounds, 3e38 and denormals, with and without goal velocities and infeasible flags;
  * the external-simulator feed (`roboy_step_external` / `roboy_reset_external`) on such wire values;
  * state / goal / step-number injection (`roboy_set_state`, `roboy_set_goal`, `roboy_set_step_num`) followed by fused steps
    whose actions hold (so the injected state is what the env sees).
Observations, done masks, goals, step words bit-exact (NaN payloads included); rewards within 1e-6, non-finite ones equal.
usage: python tools/soak_api.py [seconds] [out.json]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from cuda_adaptor import robot_from_bounds
from gym_roboy_b200.envs import RoboyEnv
from gym_roboy_b200.envs.simulations import CudaSimulationClient
from oracle import oracle as orc
from soak_generic import random_robot

SPECIALS = np.array([np.nan, np.inf, -np.inf, 0.0, -0.0, 3e38, -3e38, 1e-42, -1e-42, 1e-30], np.float32)


def salted(rng, lo, hi, n, widen=1.3, p=0.02):
    """[n, J] float32 uniform in the widened box, a fraction p of the entries replaced by special values."""
    lo, hi = np.asarray(lo, np.float64), np.asarray(hi, np.float64)
    c, h = (lo + hi) / 2, (hi - lo) / 2 * widen
    x = rng.uniform(c - h, c + h, (n, lo.size)).astype(np.float32)
    m = rng.random(x.shape) < p
    x[m] = rng.choice(SPECIALS, int(m.sum()))
    return x


def same_bits(a, b):
    return np.array_equal(np.ascontiguousarray(a, np.float32).view(np.uint32), np.ascontiguousarray(b, np.float32).view(np.uint32))


def reward_mismatch(got, want):
    got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
    fin = np.isfinite(want)
    if not np.array_equal(np.isnan(got), np.isnan(want)): return "NaN pattern"
    if not np.array_equal(got[~fin & ~np.isnan(want)], want[~fin & ~np.isnan(want)]): return "infinities"
    if fin.any():
        rel = np.abs(got[fin] - want[fin]) / np.maximum(np.abs(want[fin]), 1e-30)
        if not rel.max() <= 1e-6: return "rel %g" % rel.max()
    return None


def soak(budget=60.0, master_seed=20261018):
    master = np.random.default_rng(master_seed)
    t0 = time.time()
    summary = {"configs": 0, "rewards_checked": 0, "external_steps": 0, "injected_steps": 0, "robots": {"msj": 0, "msj_shaped": 0, "other": 0},
               "mismatches": []}
    while time.time() - t0 < budget:
        kind = int(master.integers(0, 3))
        if kind == 0:
            b, name = {}, "msj"
        elif kind == 1:
            b = dict(angle_low=-master.uniform(0.0, 3.0) * master.integers(0, 2), angle_high=master.uniform(0.3, 3.1),
                     vel_low=-master.uniform(0.0, 1.0) * master.integers(0, 2), vel_high=master.uniform(0.1, 1.0),
                     act_low=-master.uniform(0.05, 0.9) * master.integers(0, 2), act_high=master.uniform(0.05, 0.9))
            name = "msj_shaped"
        else:
            b, name = random_robot(master, int(master.integers(1, 16))), "other"
        J, A, _, bb = orc.robot_bounds(b)
        if name == "other" and J == 3 and A == 8:
            continue
        a_lo, a_hi = np.broadcast_to(bb["angle_low"], (J,)).astype(np.float32), np.broadcast_to(bb["angle_high"], (J,)).astype(np.float32)
        v_lo, v_hi = np.broadcast_to(bb["vel_low"], (J,)).astype(np.float32), np.broadcast_to(bb["vel_high"], (J,)).astype(np.float32)
        n = int(master.choice([33, 1000, 5000, 20_011]))
        seed = int(master.integers(0, 2 ** 63))
        flags = dict(penalty=bool(master.integers(0, 2)), bonus=bool(master.integers(0, 2)))
        client = CudaSimulationClient(robot=robot_from_bounds(b) if b else None, num_envs=n, seed=seed, device="cuda:0")
        env = RoboyEnv(client, joint_vel_penalty=flags["penalty"], is_agent_getting_bonus_for_reaching_goal=flags["bonus"],
                       auto_reset=False, strict=False)
        env._single = False
        ora = orc.OracleEnv(n, seed=seed, joint_vel_penalty=flags["penalty"], bonus=flags["bonus"], auto_reset=False, threads=16, **b)
        rng = np.random.default_rng(seed & 0xffffffff)
        tag = dict(robot=name, J=J, A=A, n=n, seed=seed, bounds={k: np.asarray(v, np.float64).tolist() for k, v in b.items()}, **flags)
        try:
            if not same_bits(env.reset().cpu().numpy(), ora.reset()): raise AssertionError("reset obs")
            # ---- stand-alone compute_reward ----
            for rep in range(3):
                k = 20_000
                q, qd = salted(rng, a_lo, a_hi, k), salted(rng, v_lo, v_hi, k, widen=3.0)
                gq = salted(rng, a_lo, a_hi, k, widen=1.0, p=0.01)
                near = rng.random(k) < 0.3       # goals next to the state: both sides of the reached threshold
                gq[near] = (q[near].astype(np.float64) + rng.normal(size=(int(near.sum()), J)) * float(orc.thresholds(ora.cfg)[0]) / np.sqrt(J)).astype(np.float32)
                gqd = None if rep == 0 else salted(rng, v_lo, v_hi, k, widen=1.0, p=0.01)
                feas = (rng.random(k) < 0.8).astype(np.uint8)
                r_c, reached_c = client.compute_reward(q, qd, feas, gq, gqd)
                r_o, reached_o, _ = orc.compute_reward(ora.cfg, q, qd, feas, gq, gqd)
                if not np.array_equal(reached_c.cpu().numpy(), reached_o): raise AssertionError("compute_reward reached, rep %d" % rep)
                why = reward_mismatch(r_c.cpu().numpy(), r_o)
                if why: raise AssertionError("compute_reward reward (%s), rep %d" % (why, rep))
                summary["rewards_checked"] += k
            # ---- external feed ----
            for t in range(6):
                q, qd = salted(rng, a_lo, a_hi, n, widen=1.1), salted(rng, v_lo, v_hi, n, widen=2.0)
                feas = (rng.random(n) < 0.85).astype(np.uint8)
                if t % 2:
                    g = np.clip(q + np.float32(0.001), a_lo, a_hi).astype(np.float32)
                    g[~np.isfinite(g)] = a_lo[0]
                    idx = np.arange(0, n, 3)
                    client.set_goal(g[idx], idx=idx); ora.goal[:, idx] = g[idx].T
                client.step_external(q, qd, feas)
                o_obs, o_rew, o_done = ora.step_external(q, qd, feas)
                if not same_bits(client.obs.cpu().numpy(), o_obs): raise AssertionError("external obs, step %d" % t)
                if not np.array_equal(client.done.cpu().numpy().astype(bool), o_done.astype(bool)): raise AssertionError("external done, step %d" % t)
                why = reward_mismatch(client.reward.cpu().numpy(), o_rew)
                if why: raise AssertionError("external reward (%s), step %d" % (why, t))
                m = o_done.astype(np.uint8)
                if m.any():
                    client.reset_external(q, qd, mask=m)
                    want = ora.reset_external(q, qd, m)
                    if not same_bits(client.obs.cpu().numpy()[m.astype(bool)], want[m.astype(bool)]): raise AssertionError("external reset obs, step %d" % t)
                summary["external_steps"] += n
            if not same_bits(client.goal.cpu().numpy(), ora.goal): raise AssertionError("goals after the external feed")
            # ---- injection + holding steps ----
            zero_action, can_hold = orc.hold_action(b) if b else (np.zeros(8, np.float32), True)
            for t in range(6):
                idx = np.sort(rng.choice(n, size=max(1, n // 3), replace=False))
                q, qd = salted(rng, a_lo, a_hi, idx.size, widen=1.05), salted(rng, v_lo, v_hi, idx.size, widen=1.5)
                feas = (rng.random(idx.size) < 0.7).astype(np.uint8)
                client.set_state(q, qd, feas, idx=idx)
                ora.held[0:J, idx] = q.T; ora.held[J:2 * J, idx] = qd.T
                ora.step_flags[idx] = (ora.step_flags[idx] & np.uint32(orc.STEP_MASK)) | np.where(feas, 0, orc.F_HELD_INFEASIBLE).astype(np.uint32)
                steps = rng.integers(1, 402, idx.size).astype(np.int32)
                client.set_step_num(steps, idx=idx)
                ora.step_flags[idx] = (ora.step_flags[idx] & ~np.uint32(orc.STEP_MASK)) | steps.astype(np.uint32)
                a = rng.uniform(-1, 1, (n, A)).astype(np.float32)
                a[rng.random(n) < 0.6] = zero_action
                ora_held_before, goal_before, flags_before = ora.held.copy(), ora.goal.copy(), ora.step_flags.copy()
                twin = None
                if t == 3:   # checkpoint -> a fresh handle -> the same step: every output and the whole state, bit for bit
                    sd = dict(client.state_dict(), err_flags=0)
                    twin = CudaSimulationClient(robot=robot_from_bounds(b) if b else None, num_envs=n, seed=0, device="cuda:0")
                    twin_env = RoboyEnv(twin, joint_vel_penalty=flags["penalty"], is_agent_getting_bonus_for_reaching_goal=flags["bonus"],
                                        auto_reset=False, strict=False)
                    twin_env._single = False
                    twin.load_state_dict(sd)
                obs, rew, done, _ = env.step(torch.as_tensor(a, device="cuda:0"))
                if twin is not None:
                    o2, r2, d2, _ = twin_env.step(torch.as_tensor(a, device="cuda:0"))
                    same = (torch.equal(obs.view(torch.int32), o2.view(torch.int32)) and torch.equal(rew.view(torch.int32), r2.view(torch.int32))
                            and torch.equal(done, d2) and torch.equal(client.goal.view(torch.int32), twin.goal.view(torch.int32))
                            and torch.equal(client.step_flags, twin.step_flags) and twin.counter == client.counter)
                    twin.close()
                    summary["checkpoints"] = summary.get("checkpoints", 0) + 1
                    if not same: raise AssertionError("resumed handle diverges from the original")
                o_obs, o_rew, o_done = ora.step(a)
                if not same_bits(obs.cpu().numpy(), o_obs): raise AssertionError("injected obs, step %d" % t)
                if not np.array_equal(done.cpu().numpy().astype(bool), o_done.astype(bool)): raise AssertionError("injected done, step %d" % t)
                why = reward_mismatch(rew.cpu().numpy(), o_rew)
                if why:
                    got = rew.cpu().numpy().astype(np.float64)
                    bad = np.flatnonzero((np.isnan(got) != np.isnan(o_rew)) | (~np.isnan(o_rew) & ~(np.abs(got - o_rew) <= 1e-6 * np.abs(o_rew)) & (got != o_rew)))
                    i = int(bad[0])
                    tag["detail"] = dict(env=i, n_bad=int(bad.size), got=float(got[i]), want=float(o_rew[i]), held=ora_held_before[:, i].tolist(),
                                         goal=goal_before[:, i].tolist(), flags=int(flags_before[i]), action_is_hold=bool(np.array_equal(a[i], zero_action)),
                                         obs=o_obs[i].tolist())
                    raise AssertionError("injected reward (%s), step %d" % (why, t))
                m = o_done.astype(np.uint8)
                if m.any():
                    env.reset(mask=torch.as_tensor(m)); ora.reset(m)
                summary["injected_steps"] += n
            # ---- the un-fused plug-in calls, composed the way the reference's RoboyEnv composes them (level 1 of
            # INTEGRATION.md): forward_step_command / compute_reward / get_new_goal_joint_angles / forward_reset_command on a
            # second handle reproduce the fused env's states, goals, done flags and rewards ----
            plug = CudaSimulationClient(robot=robot_from_bounds(b) if b else None, num_envs=n, seed=seed, device="cuda:0")
            fused = CudaSimulationClient(robot=robot_from_bounds(b) if b else None, num_envs=n, seed=seed, device="cuda:0")
            fenv = RoboyEnv(fused, joint_vel_penalty=flags["penalty"], is_agent_getting_bonus_for_reaching_goal=flags["bonus"],
                            auto_reset=False, strict=False)
            fenv._single = False
            plug.configure_env(flags["penalty"], flags["bonus"], False)
            plug.set_reward_range(*fenv.reward_range)
            t_hi, t_lo = np.broadcast_to(bb["act_high"], (A,)).astype(np.float32), np.broadcast_to(bb["act_low"], (A,)).astype(np.float32)
            slope = ((t_hi - t_lo) / np.float32(2.0)).astype(np.float32)

            def plug_reset(mask=None):
                plug.forward_reset_command(mask)
                goal = plug.get_new_goal_joint_angles()
                idx = None if mask is None else torch.nonzero(torch.as_tensor(mask)).flatten()
                plug.set_goal(goal if idx is None else goal[idx.to(goal.device)], idx=idx)

            fenv.reset(); plug_reset()
            stp = np.full(n, 396, np.int32)
            fused.set_step_num(stp)
            plug_steps = stp.copy()
            for t in range(8):
                a = rng.uniform(-1, 1, (n, A)).astype(np.float32)
                a[rng.random(n) < 0.1] = zero_action
                obs, rew, done, _ = fenv.step(torch.as_tensor(a, device="cuda:0"))
                rescaled = (slope * (a - np.float32(1.0))).astype(np.float32) + t_hi               # roboy_env.py:54-57,157-158
                goal_before_step = plug.goal.t().contiguous()
                state = plug.forward_step_command(torch.as_tensor(rescaled))
                if not (same_bits(state.joint_angles.cpu().numpy(), obs[:, 0:J].cpu().numpy()) and same_bits(state.joint_vels.cpu().numpy(), obs[:, J:2 * J].cpu().numpy())):
                    raise AssertionError("un-fused state, step %d" % t)
                if not same_bits(obs[:, 2 * J:].cpu().numpy(), goal_before_step.cpu().numpy()): raise AssertionError("un-fused goal, step %d" % t)
                r_p, reached_p = plug.compute_reward(state.joint_angles.to(torch.float32), state.joint_vels.to(torch.float32),
                                                      state.is_feasible.to(torch.uint8), goal_before_step, None)
                plug_steps += 1
                d_p = reached_p.cpu().numpy() | (plug_steps > 400)
                if not np.array_equal(d_p, done.cpu().numpy().astype(bool)): raise AssertionError("un-fused done, step %d" % t)
                why = reward_mismatch(r_p.cpu().numpy(), rew.cpu().numpy().astype(np.float64))
                if why: raise AssertionError("un-fused reward (%s), step %d" % (why, t))
                if d_p.any():
                    m = d_p.astype(np.uint8)
                    idx = torch.nonzero(torch.as_tensor(d_p)).flatten().to(plug.goal.device)
                    plug.set_goal(plug.get_new_goal_joint_angles()[idx], idx=idx)                   # roboy_env.py:67-68
                    fenv.reset(mask=torch.as_tensor(m)); plug_reset(m)
                    plug_steps[d_p] = 1
                    if not same_bits(fused.goal.cpu().numpy(), plug.goal.cpu().numpy()): raise AssertionError("un-fused goals after reset, step %d" % t)
                summary["unfused_steps"] = summary.get("unfused_steps", 0) + n
            if fused.counter != plug.counter: raise AssertionError("un-fused call counter %r vs %r" % (plug.counter, fused.counter))
            plug.close(); fused.close()
            s, so = client.stats(), ora.stats()
            for key in ("steps", "episodes", "successes", "timeouts", "sum_episode_len", "holds", "violations"):
                if s[key] != so[key]: raise AssertionError("stat %s: %r vs %r" % (key, s[key], so[key]))
            if not same_bits(client.goal.cpu().numpy(), ora.goal): raise AssertionError("final goals")
            if not np.array_equal(client.step_flags.cpu().numpy().astype(np.uint32), ora.step_flags): raise AssertionError("final step words")
        except AssertionError as err:
            summary["mismatches"].append(dict(tag, error=str(err)))
        summary["configs"] += 1
        summary["robots"][name] += 1
        client.close()
    summary["seconds"] = time.time() - t0
    return summary


if __name__ == "__main__":
    budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
    out = sys.argv[2] if len(sys.argv) > 2 else None
    summary = soak(budget)
    print(json.dumps(summary))
    if out:
        json.dump(summary, open(out, "w"), indent=1)
    sys.exit(1 if summary["mismatches"] else 0)
