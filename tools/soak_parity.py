"""Randomised differential soak: CUDA step kernel vs the CPU oracle over many seeds, sizes and flag combinations, with
goals planted around the goal-reached threshold so that the kernel's cheap pre-test / exact test boundary is crossed in
both directions.  usage: python tools/soak_parity.py [seconds] [out.json]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from cuda_adaptor import robot_from_bounds
from gym_roboy_b200.envs import RoboyEnv
from gym_roboy_b200.envs.simulations import CudaSimulationClient
from oracle import oracle as orc


def soak(budget=120.0, master_seed=20261018):
    """Run random configurations for `budget` seconds; returns the summary dict (mismatches listed, never raised)."""
    master = np.random.default_rng(master_seed)
    t0 = time.time()
    summary = {"configs": 0, "env_steps": 0, "successes": 0, "timeouts": 0, "holds": 0, "violations": 0, "worst_reward_rel": 0.0,
               "mismatches": []}
    summary["other_limits"] = 0
    while time.time() - t0 < budget:
        n = int(master.choice([1, 31, 33, 257, 1000, 4096, 5000, 12345]))
        seed = int(master.integers(0, 2 ** 63))
        flags = dict(penalty=bool(master.integers(0, 2)), bonus=bool(master.integers(0, 2)), auto_reset=bool(master.integers(0, 2)))
        base = int(master.choice([0, 1, 2 ** 31 - 7, 2 ** 40 + 12345]))
        T = int(master.integers(30, 450))
        b = {}
        kind = int(master.integers(0, 4))   # MSJ itself half of the time, else an MSJ-shaped robot with other limits
        if kind == 2:
            ah, vh, th = master.uniform(0.3, 3.1), master.uniform(0.1, 1.0), master.choice([0.25, 0.5, 0.1, 0.37])
            b = dict(angle_low=-ah, angle_high=ah, vel_low=-vh, vel_high=vh, act_low=-th, act_high=th)
        elif kind == 3:
            b = dict(angle_low=-master.uniform(0.0, 3.0) * master.integers(0, 2), angle_high=master.uniform(0.3, 3.1),
                     vel_low=-master.uniform(0.0, 1.0) * master.integers(0, 2), vel_high=master.uniform(0.1, 1.0),
                     act_low=-master.uniform(0.05, 0.9) * master.integers(0, 2), act_high=master.uniform(0.05, 0.9))
        summary["other_limits"] += bool(b)
        a_lo, a_hi = (np.float32(b["angle_low"]), np.float32(b["angle_high"])) if b else (-orc.PI32, orc.PI32)
        client = CudaSimulationClient(robot=robot_from_bounds(b) if b else None, num_envs=max(n, 2) if n == 1 else n, seed=seed,
                                      env_id_base=base, device="cuda:0")
        n = client.num_envs
        summary["penalty_float32"] = summary.get("penalty_float32", 0) + int(client.penalty_float32 and flags["penalty"])
        env = RoboyEnv(client, joint_vel_penalty=flags["penalty"], is_agent_getting_bonus_for_reaching_goal=flags["bonus"],
                       auto_reset=flags["auto_reset"], strict=False)
        ora = orc.OracleEnv(n, seed=seed, env_id_base=base, joint_vel_penalty=flags["penalty"], bonus=flags["bonus"],
                            auto_reset=flags["auto_reset"], threads=8, **b)
        thr = float(orc.thresholds(ora.cfg)[0])
        zero_action = orc.hold_action(b)[0] if b else np.zeros(8, np.float32)
        mid = np.float32((np.float64(a_lo) + np.float64(a_hi)) / 2)
        rng = np.random.default_rng(seed & 0xffffffff)
        env.reset(); ora.reset()
        steps = rng.integers(1, 400, n).astype(np.int32)
        client.set_step_num(steps)
        ora.step_flags[:] = (ora.step_flags & ~np.uint32(orc.STEP_MASK)) | steps.astype(np.uint32)
        tag = dict(n=n, seed=seed, base=base, T=T, bounds={k: float(v) for k, v in b.items()}, **flags)
        try:
            for t in range(T):
                if t % 7 == 3:   # plant goals at a distance spread tightly around the reached threshold (both sides)
                    q, _ = orc.draw_state(seed, np.arange(base, base + n, dtype=np.uint64), ora.counter + 1, a_lo, a_hi, J=3)
                    d = rng.normal(size=(n, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
                    r = thr * (1.0 + rng.choice([1e-7, 1e-6, 1e-5, 1e-3, 0.1], n) * rng.choice([-1, 1], n))
                    g = np.clip((q.astype(np.float64) + d * r[:, None]), a_lo, a_hi).astype(np.float32)
                    pick = rng.random(n)
                    g[pick < 0.05] = np.clip(np.float32(0.0), a_lo, a_hi)         # the zero state itself (held after a reset)
                    g[(pick >= 0.05) & (pick < 0.07)] = a_lo                      # on the bounds
                    g[(pick >= 0.07) & (pick < 0.09)] = np.nextafter(a_hi, np.float32(-10))
                    g[(pick >= 0.09) & (pick < 0.12)] = mid                       # a zero numerator of the normalisation
                    client.set_goal(g); ora.goal[:] = g.T
                a = rng.uniform(-1, 1, (n, 8)).astype(np.float32)
                a[rng.random(n) < 0.02] = zero_action
                a[rng.random(n) < 0.003] = np.nextafter(zero_action, np.float32(1)) if b else np.float32(1e-9)   # next to the hold interval
                if t % 5 == 1:
                    a[rng.integers(0, n), rng.integers(0, 8)] = [np.nan, 1.5, -1.0000001][t % 3]
                obs, rew, done, _ = env.step(torch.as_tensor(a, device="cuda:0"))
                o_obs, o_rew, o_done = ora.step(a)
                if not np.array_equal(done.cpu().numpy(), o_done): raise AssertionError("done mask, step %d" % t)
                if not np.array_equal(obs.cpu().numpy(), o_obs): raise AssertionError("obs, step %d" % t)
                rel = float((np.abs(rew.cpu().numpy().astype(np.float64) - o_rew) / np.maximum(np.abs(o_rew), 1e-30)).max())
                summary["worst_reward_rel"] = max(summary["worst_reward_rel"], rel)
                if rel > 1e-6: raise AssertionError("reward rel %g, step %d" % (rel, t))
                if not flags["auto_reset"]:
                    dd = o_done.astype(np.uint8)
                    if dd.any():
                        env.reset(mask=torch.as_tensor(dd)); ora.reset(dd)
            s, so = client.stats(), ora.stats()
            for k in ("steps", "episodes", "successes", "timeouts", "sum_episode_len", "holds", "violations"):
                if s[k] != so[k]: raise AssertionError("stat %s: %r vs %r" % (k, s[k], so[k]))
            if client.errors()[0] != ora.errors()[0] or (client.errors()[0] and client.errors()[1] != ora.errors()[1]):
                raise AssertionError("error word / first offending env: %r vs %r" % (client.errors(), ora.errors()))
            if not np.array_equal(client.goal.cpu().numpy(), ora.goal): raise AssertionError("final goals")
            if not np.array_equal(client.step_flags.cpu().numpy().astype(np.uint32), ora.step_flags): raise AssertionError("final step words")
            for k in ("successes", "timeouts", "holds", "violations"):
                summary[k] += int(s[k])
        except AssertionError as err:
            summary["mismatches"].append(dict(tag, error=str(err)))
        summary["configs"] += 1
        summary["env_steps"] += n * T
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
