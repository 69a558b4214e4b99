"""SURVEY.md 8f row 1 / BASELINE.json configs[4], fused form: `roboy_policy_rollout` runs the MlpPolicy, the
Gaussian sample, the clip and the env step for all T steps inside one kernel.  Checked here:
  * env outputs (obs / done bit-exact, reward <= 1e-6 rel) against the CPU oracle replaying clip(actions);
  * policy outputs (action mean via the recorded noise, value, log-density) against the torch float32 MlpPolicy;
  * the Gaussian noise against a numpy restatement of Philox + Box-Muller, and its moments;
  * one env per thread vs two envs per thread, and one shard vs two half shards: bit-identical;
  * CUDA-graph replay and the hand-over to the next rollout.
"""
import math

import numpy as np
import pytest
import torch

from oracle import oracle as orc

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


def make(n, seed, T, fused=True, env_id_base=0, noise_seed=77, envs_per_thread=0, policy_seed=0):
    from gym_roboy_b200.envs import RoboyEnv
    from gym_roboy_b200.envs.simulations import CudaSimulationClient
    from gym_roboy_b200.rollout import MlpPolicy, RolloutCollector
    torch.manual_seed(policy_seed)
    policy = MlpPolicy().to(DEV)
    with torch.no_grad():
        policy.log_std.copy_(torch.linspace(-1.0, 0.2, 8))     # distinct per-dimension std
    client = CudaSimulationClient(num_envs=n, seed=seed, device=DEV, env_id_base=env_id_base)
    env = RoboyEnv(client)
    col = RolloutCollector(env, policy, n_steps=T, fused=fused, noise_seed=noise_seed, envs_per_thread=envs_per_thread)
    return policy, client, col


def numpy_noise(noise_seed, gids, t):
    """Restatement of the kernel's noise: two Philox blocks (stream 4, sub 0 / 1) -> four Box-Muller pairs."""
    z = []
    for sub in (0, 1):
        x = orc._block(noise_seed, gids, t, 4, sub)
        for a, b in ((x[0], x[1]), (x[2], x[3])):
            u1 = (a >> np.uint32(8)).astype(np.float64) * 2.0 ** -24 + 2.0 ** -25
            u2 = (b >> np.uint32(8)).astype(np.float64) * 2.0 ** -24
            r = np.sqrt(-2.0 * np.log(u1))
            z += [r * np.cos(2 * np.pi * u2), r * np.sin(2 * np.pi * u2)]
    return np.stack(z, axis=-1)


# the float32 FFMA2 kernel agrees with the torch float32 policy to ~1e-6; the tensor-core kernel computes with
# float16 operands (10-bit mantissa), float32 accumulation and tanh.approx: ~1e-3
# ("tc_exact": split float16 operands, three MMAs per product, accurate tanh: float32-level again)
POLICY_TOL = {"fp32": dict(rtol=1e-5, atol=2e-5), "tc": dict(rtol=0, atol=4e-3), "tc_exact": dict(rtol=1e-5, atol=2e-5)}


@pytest.mark.parametrize("n,ept,mode", [(4096, 0, "fp32"), (2531, 2, "fp32"), (40000, 0, "fp32"),
                                        (4096, 0, "tc"), (2531, 2, "tc"), (80000, 0, "tc"), (4096, 1, "tc"), (80000, 3, "tc"),
                                        (4096, 0, "tc_exact"), (2531, 0, "tc_exact"), (80000, 0, "tc_exact")])
def test_fused_rollout_against_torch_policy_and_oracle_env(n, ept, mode):
    T, seed = 10, 21
    tol = POLICY_TOL[mode]
    policy, client, col = make(n, seed, T, fused=mode, envs_per_thread=ept)
    col.noise = torch.zeros((T, n, 8), dtype=torch.float32, device=DEV)
    client.set_step_num(torch.full((n,), 395, dtype=torch.int32))   # every env times out inside the rollout
    ora = orc.OracleEnv(n, seed=seed)
    ora.reset()
    ora.step_flags[:] = (ora.step_flags & ~np.uint32(orc.STEP_MASK)) | np.uint32(395)
    t0 = client.counter
    obs0 = col.obs[0].clone()
    col.collect()
    torch.cuda.synchronize()
    assert client.errors() == (0, None)
    assert torch.equal(col.obs[0], col.obs[T])                      # handed over to the next rollout
    std = policy.log_std.detach().exp()
    lognorm = -0.5 * math.log(2 * math.pi) * 8 - float(policy.log_std.sum())
    obs_t = obs0
    with torch.no_grad():
        for t in range(T):
            mean, value = policy(obs_t)
            assert torch.allclose(col.values[t], value, **tol)
            z = col.noise[t]
            assert torch.allclose(col.actions[t], mean + std * z, **tol)
            assert torch.allclose(col.logp[t], lognorm - 0.5 * (z * z).sum(-1), rtol=1e-5, atol=1e-4)
            want_z = numpy_noise(77, np.arange(n, dtype=np.uint64), t0 + 1 + t)
            assert np.allclose(z.cpu().numpy(), want_z, rtol=0, atol=2e-4)
            # the env saw the clipped action: replay it through the oracle
            a = np.clip(col.actions[t].cpu().numpy(), -1.0, 1.0)
            o, r, d = ora.step(a)
            assert np.array_equal(col.obs[t + 1].cpu().numpy() if t + 1 < T else col.obs[T].cpu().numpy(), o)
            assert np.array_equal(col.dones[t].cpu().numpy().astype(bool), d)
            assert np.allclose(col.rewards[t].cpu().numpy(), r, rtol=1e-6, atol=0)
            obs_t = col.obs[t + 1]
        assert torch.allclose(col.values[T], policy(col.obs[T])[1], **tol)
    assert int(col.dones.sum()) >= n
    s, so = client.stats(), ora.stats()
    assert all(s[k] == so[k] for k in ("steps", "episodes", "successes", "timeouts", "holds", "violations")), (s, so)
    assert np.array_equal(client.goal.cpu().numpy(), ora.goal)
    assert np.array_equal(client.step_flags.cpu().numpy().astype(np.uint32), ora.step_flags)
    assert client.counter == t0 + T == ora.counter
    zz = col.noise.flatten().double()
    assert abs(float(zz.mean())) < 5e-3 and abs(float(zz.var()) - 1.0) < 1e-2 and abs(float((zz ** 4).mean()) - 3.0) < 0.1


@pytest.mark.parametrize("mode", ["fp32", "tc"])
def test_one_and_two_envs_per_thread_and_sharding_are_bit_identical(mode):
    """fp32: one or two envs per thread; tc: one or two 128-env tiles per thread group (ping-pong), or the merged form."""
    n, T = 6000, 6
    ref = None
    for ept in ((1, 2) if mode == "fp32" else (1, 2, 3)):
        _, client, col = make(n, 5, T, fused=mode, envs_per_thread=ept)
        col.collect()
        torch.cuda.synchronize()
        got = {k: getattr(col, k).clone() for k in ("obs", "actions", "logp", "values", "rewards", "dones", "adv", "ret")}
        if ref is None:
            ref = got
        else:
            for k in ref:
                assert torch.equal(ref[k], got[k]), k
    shards_of_one_population(n, T, ref, mode)


@pytest.mark.parametrize("mode", ["fp32", "tc", "tc_exact"])
def test_sharding_invariance(mode):
    n, T = 6000, 6
    _, client, col = make(n, 5, T, fused=mode)
    col.collect()
    torch.cuda.synchronize()
    ref = {k: getattr(col, k).clone() for k in ("obs", "actions", "logp", "values", "rewards", "dones")}
    shards_of_one_population(n, T, ref, mode)


def shards_of_one_population(n, T, ref, mode):
    # two shards of the same population (global env ids 0..2999 and 3000..5999)
    half = n // 2
    for base in (0, half):
        _, client, col = make(half, 5, T, env_id_base=base, fused=mode)
        col.collect()
        torch.cuda.synchronize()
        for k in ("actions", "logp", "values", "rewards", "dones"):
            assert torch.equal(getattr(col, k), ref[k][:, base:base + half]), k
        assert torch.equal(col.obs[1:T], ref["obs"][1:T, base:base + half])


@pytest.mark.parametrize("mode", ["fp32", "tc", "tc_exact"])
def test_holds_and_nan_actions_inside_the_fused_kernel(mode):
    """A zero policy (mean 0, std tiny) makes every env take the Stub's hold branch; a NaN weight trips the
    action assert (roboy_env.py:52) exactly as T un-fused steps would."""
    n, T = 1024, 4
    policy, client, col = make(n, 9, T, fused=mode)
    with torch.no_grad():
        for p_ in policy.pi.parameters():
            p_.zero_()
        policy.log_std.fill_(-40.0)                                  # std = 4e-18: actions ~ 0 -> hold
    ora = orc.OracleEnv(n, seed=9)
    ora.reset()
    col.collect()
    torch.cuda.synchronize()
    for t in range(T):
        a = np.clip(col.actions[t].cpu().numpy(), -1, 1)
        assert np.abs(a).max() < 1e-12
        o, r, d = ora.step(a)
        assert np.array_equal(col.obs[t + 1].cpu().numpy(), o) and np.array_equal(col.dones[t].cpu().numpy().astype(bool), d)
        assert np.allclose(col.rewards[t].cpu().numpy(), r, rtol=1e-6, atol=0)
    assert client.stats()["holds"] == n * T == ora.stats()["holds"]
    with torch.no_grad():
        policy.pi[4].bias[3] = float("nan")
    col.collect()
    torch.cuda.synchronize()
    flags, first = client.errors()
    from gym_roboy_b200 import _native
    assert flags & _native.ERR_ACTION and first == 0
    assert torch.isnan(col.actions[:, :, 3]).all()


@pytest.mark.parametrize("mode", ["fp32", "tc", "tc_exact"])
def test_fused_rollout_in_a_cuda_graph_and_against_the_unfused_collector(mode):
    n, T = 4096, 8
    policy, client, col = make(n, 3, T, fused=mode)
    sd = client.state_dict()
    obs0 = col.obs[0].clone()
    col.collect()
    torch.cuda.synchronize()
    eager = {k: getattr(col, k).clone() for k in ("obs", "actions", "logp", "values", "rewards", "dones", "adv", "ret")}
    col.capture()
    client.load_state_dict(sd)
    col.obs[0].copy_(obs0)
    col.collect()                                                   # graph replay
    torch.cuda.synchronize()
    for k, v in eager.items():
        if k == "obs":
            assert torch.equal(col.obs[1:], v[1:]), k
        else:
            assert torch.equal(getattr(col, k), v), k
    # the un-fused collector on the same policy, fed the fused kernel's actions, yields the same env trajectory
    client.load_state_dict(sd)
    for t in range(T):
        client.step_fused(torch.clamp(eager["actions"][t], -1, 1).contiguous())
        assert torch.equal(client.obs, eager["obs"][t + 1]) and torch.equal(client.reward, eager["rewards"][t])
        assert torch.equal(client.done_u8, eager["dones"][t])


@pytest.mark.parametrize("mode,ept", [("fp32", 1), ("fp32", 2), ("tc", 1), ("tc", 2), ("tc", 3), ("tc_exact", 0)])
@pytest.mark.parametrize("n", [2, 33, 257, 1001, 2531])
def test_ragged_sizes_write_nothing_outside_their_buffers(n, mode, ept):
    """(compute-sanitizer is not available on the GPU pool.)  Every rollout buffer is carved out of one arena with
    sentinel-filled guard bands on both sides; after two rollouts at ragged sizes the guards must be untouched and
    every slot inside must have been written."""
    T, guard = 3, 256
    policy, client, col = make(n, 31, T, fused=mode, envs_per_thread=ept)
    shapes = {"obs": (T + 1, n, 9), "actions": (T, n, 8), "logp": (T, n), "values": (T + 1, n), "rewards": (T, n),
              "noise": (T, n, 8), "adv": (T, n), "ret": (T, n)}
    sentinel = -12345.0
    total = sum(guard + ((int(np.prod(s)) + 3) // 4) * 4 for s in shapes.values()) + guard
    arena = torch.full((total,), sentinel, dtype=torch.float32, device=DEV)
    off, spans = guard, {}
    for k, shp in shapes.items():
        cnt = int(np.prod(shp))
        view = arena[off:off + cnt].view(shp)
        if k == "obs":
            view[0].copy_(col.obs[0])
        setattr(col, k, view)
        spans[k] = (off, cnt)
        off += ((cnt + 3) // 4) * 4 + guard
    dones_arena = torch.full((T * n + 2 * guard,), 0x5A, dtype=torch.uint8, device=DEV)
    col.dones = dones_arena[guard:guard + T * n].view(T, n)
    col.collect()
    col.collect()
    torch.cuda.synchronize()
    inside = torch.zeros(total, dtype=torch.bool, device=DEV)
    for k, (o, cnt) in spans.items():
        inside[o:o + cnt] = True
        assert not bool((arena[o:o + cnt] == sentinel).any()), k + " has unwritten slots"
    assert bool((arena[~inside] == sentinel).all()), "a float buffer was written outside its bounds"
    assert bool((dones_arena[:guard] == 0x5A).all()) and bool((dones_arena[guard + T * n:] == 0x5A).all())
    assert bool((col.dones <= 1).all())
    assert client.errors() == (0, None)


@pytest.mark.parametrize("mode", ["fp32", "tc", "tc_exact"])
def test_long_fused_rollout_crosses_several_episode_ends(mode):
    """T = 900 steps in one launch: every env times out at least twice inside the kernel (goal re-draw, reset
    observation, step counter restart), all of it replayed through the oracle on the stored actions."""
    n, T, seed = 384, 900, 17
    policy, client, col = make(n, seed, T, fused=mode)
    ora = orc.OracleEnv(n, seed=seed)
    assert np.array_equal(ora.reset(), col.obs[0].cpu().numpy())
    col.collect()
    torch.cuda.synchronize()
    acts = np.clip(col.actions.cpu().numpy(), -1.0, 1.0)
    obs, rew, done = col.obs.cpu().numpy(), col.rewards.cpu().numpy(), col.dones.cpu().numpy().astype(bool)
    for t in range(T):
        o, r, d = ora.step(acts[t])
        assert np.array_equal(obs[t + 1] if t + 1 < T else col.obs[T].cpu().numpy(), o), t
        assert np.array_equal(done[t], d), t
        assert np.allclose(rew[t], r, rtol=1e-6, atol=0), t
    assert done.sum(axis=0).min() >= 2
    s, so = client.stats(), ora.stats()
    assert all(s[k] == so[k] for k in ("steps", "episodes", "successes", "timeouts", "holds", "violations")), (s, so)
    assert np.array_equal(client.goal.cpu().numpy(), ora.goal)
    assert np.array_equal(client.step_flags.cpu().numpy().astype(np.uint32), ora.step_flags)
    assert client.errors() == (0, None)


def test_env_trajectory_is_identical_across_the_three_fused_kernels():
    """The Stub's next state does not depend on the action unless it is (numerically) zero, so the ENV outputs of a
    rollout are a function of the env seed alone: the float32, tensor-core and exact tensor-core kernels -- whose
    policy outputs differ by up to 1e-3 -- must produce bit-identical observations, rewards and done masks."""
    n, T = 8192, 64
    outs = {}
    for mode in ("fp32", "tc", "tc_exact"):
        _, client, col = make(n, 123, T, fused=mode)
        client.set_step_num(torch.randint(1, 400, (n,), dtype=torch.int32, device=DEV, generator=torch.Generator(DEV).manual_seed(1)))
        col.collect()
        torch.cuda.synchronize()
        assert float(col.actions.abs().amax(-1).min()) > 1e-3         # no action vector anywhere near the hold interval
        outs[mode] = (col.obs[1:].clone(), col.rewards.clone(), col.dones.clone(), client.goal.clone(), client.step_flags.clone())
        assert int(col.dones.sum()) > 0
    for mode in ("tc", "tc_exact"):
        for a, b in zip(outs["fp32"], outs[mode]):
            assert torch.equal(a, b), mode
    # ... while the policy outputs agree only to the kernels' tolerances (checked elsewhere)


@pytest.mark.parametrize("mode", ["fp32", "tc", "tc_exact"])
def test_saturated_policy(mode):
    """Weights scaled up until most tanh units saturate (|pre-activation| up to ~60): no NaN / inf anywhere, values and
    means still track the torch float32 policy (relative to their now larger magnitude)."""
    n, T = 2048, 3
    policy, client, col = make(n, 41, T, fused=mode)
    with torch.no_grad():
        for net in (policy.pi, policy.vf):
            net[0].weight.mul_(8.0); net[2].weight.mul_(8.0); net[4].weight.mul_(4.0)
    col.noise = torch.zeros((T, n, 8), dtype=torch.float32, device=DEV)
    obs0 = col.obs[0].clone()
    col.collect()
    torch.cuda.synchronize()
    for k in ("actions", "logp", "values", "rewards", "adv", "ret"):
        assert bool(torch.isfinite(getattr(col, k)).all()), k
    with torch.no_grad():
        mean, value = policy(obs0)
    scale = float(mean.abs().max())
    tol = 5e-3 if mode == "tc" else 5e-5
    got_mean = col.actions[0] - policy.log_std.detach().exp() * col.noise[0]
    assert float((got_mean - mean).abs().max()) <= tol * max(scale, 1.0)
    assert float((col.values[0] - value).abs().max()) <= tol * max(float(value.abs().max()), 1.0)
    assert client.errors() == (0, None)
