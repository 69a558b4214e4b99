"""Randomised differential soak of the GENERIC fused step (robots other than MSJ): random joint / tendon counts, one
asymmetric bound per component, random flags and env-id bases, CUDA vs the CPU oracle -- observations, done masks, goals and
step words bit-exact, rewards within 1e-6.  Goals are planted around the reached threshold (both sides), at the midpoint
of the angle space (a zero numerator: the step's proved division hands over to the IEEE path) and next to the bounds;
actions include holds, near-holds, out-of-range values and NaN.
usage: python tools/soak_generic.py [seconds] [out.json]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from cuda_adaptor import robot_from_bounds
from gym_roboy_b200.envs import RoboyEnv
from gym_roboy_b200.envs.simulations import CudaSimulationClient
from oracle import oracle as orc


def random_robot(rng, J):
    A = int(rng.integers(1, 65))
    a_hi = rng.uniform(0.3, 3.1, J); a_lo = -rng.uniform(0.3, 3.1, J)
    v_hi = rng.uniform(0.1, 1.0, J); v_lo = -rng.uniform(0.1, 1.0, J)
    t_hi = rng.uniform(0.05, 0.9, A); t_lo = -rng.uniform(0.05, 0.9, A)
    kind = rng.integers(0, 4)
    if kind == 0:     # symmetric dyadic tendon ranges: a robot that CAN hold
        t_hi = np.full(A, rng.choice([0.25, 0.5, 0.125])); t_lo = -t_hi
    elif kind == 1:   # symmetric angle space: 2*v - max - min is exactly 2*v, the midpoint exactly 0
        a_lo = -a_hi
        if rng.integers(0, 2):   # ... and symmetric velocity space: the float32 penalty path (PenaltyF32) up to 8 joints
            v_lo = -v_hi
    elif kind == 2:   # one-sided spaces (a one-sided tendon range holds at a = -1, the edge of the action space)
        a_lo = np.zeros(J); v_lo = np.zeros(J)
        if rng.integers(0, 2):
            t_lo = np.zeros(A)
    return dict(angle_low=a_lo, angle_high=a_hi, vel_low=v_lo, vel_high=v_hi, act_low=t_lo, act_high=t_hi)


def soak(budget=120.0, master_seed=20261018):
    master = np.random.default_rng(master_seed)
    t0 = time.time()
    summary = {"configs": 0, "env_steps": 0, "joint_counts": {}, "fast_division": 0, "successes": 0, "timeouts": 0, "holds": 0,
               "violations": 0, "worst_reward_rel": 0.0, "mismatches": []}
    while time.time() - t0 < budget:
        J = int(master.integers(1, 16))
        b = random_robot(master, J)
        _, A, _, bb = orc.robot_bounds(b)
        if J == 3 and A == 8:
            continue
        n = int(master.choice([33, 1000, 4096, 20_011, 65_536]))
        seed = int(master.integers(0, 2 ** 63))
        flags = dict(penalty=bool(master.integers(0, 2)), bonus=bool(master.integers(0, 2)), auto_reset=bool(master.integers(0, 2)))
        base = int(master.choice([0, 2 ** 31 - 7, 2 ** 40 + 12345]))
        T = int(master.integers(20, 120))
        client = CudaSimulationClient(robot=robot_from_bounds(b), num_envs=n, seed=seed, env_id_base=base, device="cuda:0")
        env = RoboyEnv(client, joint_vel_penalty=flags["penalty"], is_agent_getting_bonus_for_reaching_goal=flags["bonus"],
                       auto_reset=flags["auto_reset"], strict=False)
        env._single = False
        ora = orc.OracleEnv(n, seed=seed, env_id_base=base, joint_vel_penalty=flags["penalty"], bonus=flags["bonus"],
                            auto_reset=flags["auto_reset"], threads=16, **b)
        rng = np.random.default_rng(seed & 0xffffffff)
        env.reset(); ora.reset()
        steps = rng.integers(1, 400, n).astype(np.int32)
        client.set_step_num(steps)
        ora.step_flags[:] = (ora.step_flags & ~np.uint32(orc.STEP_MASK)) | steps.astype(np.uint32)
        zero_action, _ = orc.hold_action(b)
        thr = float(orc.thresholds(ora.cfg)[0])
        lo, hi = np.broadcast_to(bb["angle_low"], (J,)).astype(np.float32), np.broadcast_to(bb["angle_high"], (J,)).astype(np.float32)
        mid = ((lo.astype(np.float64) + hi.astype(np.float64)) / 2).astype(np.float32)
        tag = dict(J=J, A=A, n=n, seed=seed, base=base, T=T, **flags)
        tag["reward_range"] = [list(map(float, env.reward_range)), list(map(float, ora.reward_range))]
        tag["bounds"] = {k: np.asarray(v, np.float64).tolist() for k, v in b.items()}
        viol = [0, 0]
        summary["fast_division"] += int(client.fast_division)
        summary["penalty_float32"] = summary.get("penalty_float32", 0) + int(client.penalty_float32 and flags["penalty"])
        try:
            for t in range(T):
                if t % 7 == 3:
                    q, _ = orc.draw_state(seed, np.arange(base, base + n, dtype=np.uint64), ora.counter + 1, bb["angle_low"], bb["angle_high"], J=J)
                    d = rng.normal(size=(n, J)); d /= np.linalg.norm(d, axis=1, keepdims=True)
                    r = thr * (1.0 + rng.choice([1e-7, 1e-6, 1e-5, 1e-3, 0.1], n) * rng.choice([-1, 1], n))
                    g = np.clip(q.astype(np.float64) + d * r[:, None], lo, hi).astype(np.float32)
                    pick = rng.random(n)
                    g[pick < 0.05] = mid                                         # zero numerators
                    g[(pick >= 0.05) & (pick < 0.08)] = lo                       # on the bounds
                    g[(pick >= 0.08) & (pick < 0.10)] = np.nextafter(hi, np.float32(-10))
                    client.set_goal(g); ora.goal[:] = g.T
                a = rng.uniform(-1, 1, (n, A)).astype(np.float32)
                a[rng.random(n) < 0.01] = zero_action
                a[rng.random(n) < 0.003] = np.nextafter(zero_action, np.float32(1))
                if t % 5 == 1:
                    a[rng.integers(0, n), rng.integers(0, A)] = [np.nan, 1.5, -1.0000001][t % 3]
                obs, rew, done, _ = env.step(torch.as_tensor(a, device="cuda:0"))
                o_obs, o_rew, o_done = ora.step(a)
                if not np.array_equal(done.cpu().numpy(), o_done): raise AssertionError("done mask, step %d" % t)
                if not np.array_equal(obs.cpu().numpy(), o_obs): raise AssertionError("obs, step %d" % t)
                rel = float((np.abs(rew.cpu().numpy().astype(np.float64) - o_rew) / np.maximum(np.abs(o_rew), 1e-30)).max())
                summary["worst_reward_rel"] = max(summary["worst_reward_rel"], rel)
                if rel > 1e-6: raise AssertionError("reward rel %g, step %d" % (rel, t))
                for i, (r_, rr) in enumerate(((rew.cpu().numpy().astype(np.float64), env.reward_range), (o_rew.astype(np.float64), ora.reward_range))):
                    viol[i] += int((~((rr[0] <= r_) & (r_ <= rr[1]))).sum())
                tag["violations_recomputed_from_rewards"] = list(viol)
                if not flags["auto_reset"]:
                    dd = o_done.astype(np.uint8)
                    if dd.any():
                        env.reset(mask=torch.as_tensor(dd)); ora.reset(dd)
            s, so = client.stats(), ora.stats()
            for k in ("steps", "episodes", "successes", "timeouts", "sum_episode_len", "holds", "violations"):
                if s[k] != so[k]: raise AssertionError("stat %s: %r vs %r" % (k, s[k], so[k]))
            if client.errors() != ora.errors(): raise AssertionError("error word / first offending env")
            if not np.array_equal(client.goal.cpu().numpy(), ora.goal): raise AssertionError("final goals")
            if not np.array_equal(client.step_flags.cpu().numpy().astype(np.uint32), ora.step_flags): raise AssertionError("final step words")
            for k in ("successes", "timeouts", "holds", "violations"):
                summary[k] += int(s[k])
        except AssertionError as err:
            summary["mismatches"].append(dict(tag, error=str(err)))
        summary["configs"] += 1
        summary["env_steps"] += n * T
        summary["joint_counts"][str(J)] = summary["joint_counts"].get(str(J), 0) + 1
        client.close()
    summary["seconds"] = time.time() - t0
    return summary


if __name__ == "__main__":
    budget = float(sys.argv[1]) if len(sys.argv) > 1 else 120.0
    out = sys.argv[2] if len(sys.argv) > 2 else None
    summary = soak(budget)
    print(json.dumps(summary))
    if out:
        json.dump(summary, open(out, "w"), indent=1)
    sys.exit(1 if summary["mismatches"] else 0)
