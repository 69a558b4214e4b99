"""Randomised differential soak of the HOST-BUFFER step (`roboy_step_host`, the end-to-end path of bench.py): random
population sizes (ragged against the stage size), stage sizes, stream counts, ramp on/off, ring / split pattern, staged /
mapped modes, autotune, done-index list and terminal observations on/off, MSJ and other robots, device steps interleaved
between host steps -- against the CPU oracle: observations, done masks, done-index lists bit-exact, rewards within 1e-6.
usage: python tools/soak_host.py [seconds] [out.json]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from cuda_adaptor import robot_from_bounds
from gym_roboy_b200 import _native
from gym_roboy_b200.envs import RoboyEnv
from gym_roboy_b200.envs.simulations import CudaSimulationClient
from oracle import oracle as orc
from soak_generic import random_robot


def soak(budget=60.0, master_seed=20261018):
    master = np.random.default_rng(master_seed)
    t0 = time.time()
    summary = {"configs": 0, "env_steps": 0, "modes": {}, "robots": {"msj": 0, "other": 0}, "worst_reward_rel": 0.0, "mismatches": []}
    while time.time() - t0 < budget:
        generic = bool(master.integers(0, 3) == 0)
        b = random_robot(master, int(master.integers(1, 16))) if generic else {}
        J, A = (orc.robot_bounds(b)[0], orc.robot_bounds(b)[1]) if generic else (3, 8)
        if generic and J == 3 and A == 8:
            continue
        n = int(master.choice([1, 31, 1000, 4097, 65_537, 300_001, 1_048_576 + 5]))
        seed = int(master.integers(0, 2 ** 63))
        flags = dict(penalty=bool(master.integers(0, 2)), bonus=bool(master.integers(0, 2)), auto_reset=bool(master.integers(0, 2)))
        pipe = dict(stage_envs=int(master.choice([32, 4096, 1 << 14, 1 << 16, 1 << 19])), n_streams=int(master.integers(1, 5)),
                    ramp=bool(master.integers(0, 2)), ring=bool(master.integers(0, 2)))
        pipe["stage_envs"] = max(pipe["stage_envs"], 32 * ((n + 32 * 200 - 1) // (32 * 200)))   # at most ~200 stages per step
        mode = int(master.choice([_native.HOST_STAGED, _native.HOST_STAGED, _native.HOST_MAPPED_OUT, _native.HOST_MAPPED_ALL]))
        autotune = bool(master.integers(0, 4) == 0)
        done_index = bool(master.integers(0, 2))
        T = int(master.integers(4, 16)) if n > 100_000 else int(master.integers(10, 60))
        n_eff = max(n, 2) if n == 1 else n
        client = CudaSimulationClient(robot=robot_from_bounds(b) if generic else None, num_envs=n_eff, seed=seed, device="cuda:0")
        env = RoboyEnv(client, joint_vel_penalty=flags["penalty"], is_agent_getting_bonus_for_reaching_goal=flags["bonus"],
                       auto_reset=flags["auto_reset"], strict=False)
        env._single = False
        ora = orc.OracleEnv(n_eff, seed=seed, joint_vel_penalty=flags["penalty"], bonus=flags["bonus"],
                            auto_reset=flags["auto_reset"], threads=16, **b)
        n = n_eff
        rng = np.random.default_rng(seed & 0xffffffff)
        env.reset(); ora.reset()
        steps = rng.integers(300, 401, n).astype(np.int32)
        client.set_step_num(steps)
        ora.step_flags[:] = (ora.step_flags & ~np.uint32(orc.STEP_MASK)) | steps.astype(np.uint32)
        if autotune:
            client.set_host_autotune(True)
        else:
            client.set_host_pipeline(**pipe)
        client.set_host_mode(mode)
        if done_index:
            client.enable_terminal_obs(True)
            client.enable_done_index(True)
        a_h, obs_h, rew_h, done_h = client.host_buffers(write_combined_actions=bool(master.integers(0, 2)))
        zero_action, _ = orc.hold_action(b) if generic else (np.zeros(8, np.float32), True)
        tag = dict(generic=generic, J=J, A=A, n=n, seed=seed, T=T, mode=mode, autotune=autotune, done_index=done_index, **pipe, **flags)
        try:
            for t in range(T):
                a = rng.uniform(-1, 1, (n, A)).astype(np.float32)
                a[rng.random(n) < 0.02] = zero_action
                if t % 4 == 2:   # a device-buffer step in between: the two entry points share the call counter and the state
                    obs, rew, done, _ = env.step(torch.as_tensor(a, device="cuda:0"))
                    got = (obs.cpu().numpy(), rew.cpu().numpy(), done.cpu().numpy().astype(bool))
                else:
                    a_h[...] = a
                    obs_h[...] = np.nan; rew_h[...] = np.nan; done_h[...] = 7
                    client.step_host(a_h, obs_h, rew_h, done_h)
                    got = (obs_h.copy(), rew_h.copy(), done_h.astype(bool))
                    if done_h.max() > 1: raise AssertionError("done bytes not all written, step %d" % t)
                if done_index:
                    o_obs, o_rew, o_done, o_term = ora.step(a, want_terminal_obs=True)
                else:
                    o_obs, o_rew, o_done = ora.step(a)
                if not np.array_equal(got[2], o_done.astype(bool)): raise AssertionError("done mask, step %d" % t)
                if not np.array_equal(got[0], o_obs): raise AssertionError("obs, step %d" % t)
                rel = float((np.abs(got[1].astype(np.float64) - o_rew) / np.maximum(np.abs(o_rew), 1e-30)).max())
                summary["worst_reward_rel"] = max(summary["worst_reward_rel"], rel)
                if not rel <= 1e-6: raise AssertionError("reward rel %g, step %d" % (rel, t))
                if done_index:
                    idx, rows = client.done_indices(with_terminal_obs=flags["auto_reset"])
                    want = np.flatnonzero(o_done).astype(np.int32)
                    if not np.array_equal(idx.cpu().numpy(), want): raise AssertionError("done index list, step %d" % t)
                    if flags["auto_reset"] and not np.array_equal(rows.cpu().numpy(), o_term[want]):
                        raise AssertionError("terminal rows, step %d" % t)
                if not flags["auto_reset"]:
                    dd = o_done.astype(np.uint8)
                    if dd.any():
                        env.reset(mask=torch.as_tensor(dd)); ora.reset(dd)
            s, so = client.stats(), ora.stats()
            for k in ("steps", "episodes", "successes", "timeouts", "sum_episode_len", "holds", "violations"):
                if s[k] != so[k]: raise AssertionError("stat %s: %r vs %r" % (k, s[k], so[k]))
            if client.counter != ora.counter: raise AssertionError("call counter %r vs %r" % (client.counter, ora.counter))
        except AssertionError as err:
            summary["mismatches"].append(dict(tag, error=str(err)))
        summary["configs"] += 1
        summary["env_steps"] += n * T
        summary["modes"][str(mode)] = summary["modes"].get(str(mode), 0) + 1
        summary["robots"]["other" if generic else "msj"] += 1
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
