"""Randomised soak of the fused rollout kernels: random seeds, ragged sizes, env-id bases, T, kernel (float32 FFMA2 with 1 / 2
envs per thread; tensor cores: one tile per group, ping-pong, merged, exact), env flags.  The env outputs of every rollout
are replayed through the CPU oracle on the stored actions; the values are checked against the torch policy.
usage: python tools/soak_rollout.py [seconds] [out.json]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gym_roboy_b200.envs import RoboyEnv
from gym_roboy_b200.envs.simulations import CudaSimulationClient
from gym_roboy_b200.rollout import MlpPolicy, RolloutCollector
from oracle import oracle as orc

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 120.0
out = sys.argv[2] if len(sys.argv) > 2 else None
master = np.random.default_rng(20261019)
KERNELS = [("fp32", 1), ("fp32", 2), ("tc", 1), ("tc", 2), ("tc", 3), ("tc_exact", 0)]
t0 = time.time()
summary = {"rollouts": 0, "env_steps": 0, "episodes": 0, "holds": 0, "worst_reward_rel": 0.0, "worst_value_abs": {}, "mismatches": []}
while time.time() - t0 < budget:
    n = int(master.choice([2, 33, 129, 1000, 4096, 5001, 40000]))
    seed = int(master.integers(0, 2 ** 63))
    base = int(master.choice([0, 3, 2 ** 33 + 5]))
    T = int(master.integers(3, 48))
    mode, ept = KERNELS[int(master.integers(0, len(KERNELS)))]
    penalty, bonus = bool(master.integers(0, 2)), bool(master.integers(0, 2))
    tag = dict(n=n, seed=seed, base=base, T=T, mode=mode, ept=ept, penalty=penalty, bonus=bonus)
    torch.manual_seed(seed & 0xffff)
    policy = MlpPolicy().to("cuda:0")
    with torch.no_grad():
        policy.log_std.fill_(float(master.choice([-30.0, -1.0, 0.0])))     # -30: actions ~ mean; some policies hold
        if master.random() < 0.2:
            for q in policy.pi.parameters(): q.zero_()                      # mean 0 (+ tiny std): the Stub's hold branch
    client = CudaSimulationClient(num_envs=n, seed=seed, env_id_base=base, device="cuda:0")
    env = RoboyEnv(client, joint_vel_penalty=penalty, is_agent_getting_bonus_for_reaching_goal=bonus, strict=False)
    col = RolloutCollector(env, policy, n_steps=T, fused=mode, envs_per_thread=ept, noise_seed=seed ^ 0x5555)
    ora = orc.OracleEnv(n, seed=seed, env_id_base=base, joint_vel_penalty=penalty, bonus=bonus, threads=8)
    try:
        if not np.array_equal(ora.reset(), col.obs[0].cpu().numpy()): raise AssertionError("reset obs")
        steps = np.random.default_rng(seed & 0xffffffff).integers(1, 400, n).astype(np.int32)
        client.set_step_num(steps)
        ora.step_flags[:] = (ora.step_flags & ~np.uint32(orc.STEP_MASK)) | steps.astype(np.uint32)
        obs0 = col.obs[0].clone()
        col.collect(); torch.cuda.synchronize()
        acts = np.clip(col.actions.cpu().numpy(), -1.0, 1.0)
        obs, rew, done = col.obs.cpu().numpy(), col.rewards.cpu().numpy(), col.dones.cpu().numpy().astype(bool)
        for t in range(T):
            o, r, d = ora.step(acts[t])
            if not np.array_equal(obs[t + 1], o): raise AssertionError("obs, step %d" % t)
            if not np.array_equal(done[t], d): raise AssertionError("done, step %d" % t)
            rel = float((np.abs(rew[t].astype(np.float64) - r) / np.maximum(np.abs(r), 1e-30)).max())
            summary["worst_reward_rel"] = max(summary["worst_reward_rel"], rel)
            if rel > 1e-6: raise AssertionError("reward rel %g, step %d" % (rel, t))
        s, so = client.stats(), ora.stats()
        for k in ("steps", "episodes", "successes", "timeouts", "holds", "violations"):
            if s[k] != so[k]: raise AssertionError("stat %s: %r vs %r" % (k, s[k], so[k]))
        if not np.array_equal(client.step_flags.cpu().numpy().astype(np.uint32), ora.step_flags): raise AssertionError("step words")
        with torch.no_grad():
            _, value = policy(obs0)
        verr = float((col.values[0] - value).abs().max())
        summary["worst_value_abs"][mode] = max(summary["worst_value_abs"].get(mode, 0.0), verr)
        if verr > (4e-3 if mode == "tc" else 2e-5): raise AssertionError("value err %g" % verr)
        summary["episodes"] += int(s["episodes"]); summary["holds"] += int(s["holds"])
    except AssertionError as err:
        summary["mismatches"].append(dict(tag, error=str(err)))
    summary["rollouts"] += 1
    summary["env_steps"] += n * T
    client.close()
summary["seconds"] = time.time() - t0
print(json.dumps(summary))
if out:
    json.dump(summary, open(out, "w"), indent=1)
