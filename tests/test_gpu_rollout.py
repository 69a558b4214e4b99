"""the rollout consumer (policy on the GPU, buffers written in place by the
step kernel, GAE kernel, CUDA-graph capture) and the VecEnv adapter."""
import numpy as np
import pytest
import torch

from oracle import oracle as orc

pytestmark = pytest.mark.gpu


def gae_numpy(rew, val, done, last_val, gamma, lam):
    """stable-baselines PPO2 runner recurrence restated (test-only reference)."""
    T, n = rew.shape
    adv = np.zeros((T, n), np.float64)
    nxt, g = last_val.astype(np.float64), np.zeros(n)
    for t in reversed(range(T)):
        nt = 1.0 - done[t]
        delta = rew[t] + gamma * nxt * nt - val[t]
        g = delta + gamma * lam * nt * g
        adv[t] = g
        nxt = val[t].astype(np.float64)
    return adv, adv + val


def test_gae_kernel_matches_the_runner_recurrence():
    from gym_roboy_b200.rollout import gae
    rng = np.random.default_rng(0)
    T, n = 128, 5000
    rew = rng.normal(-3, 2, (T, n)).astype(np.float32)
    val = rng.normal(-50, 10, (T, n)).astype(np.float32)
    done = (rng.random((T, n)) < 0.02)
    last = rng.normal(-50, 10, n).astype(np.float32)
    to = lambda a: torch.as_tensor(a, device="cuda:0")  # noqa: E731
    adv, ret = gae(to(rew), to(val), to(done.astype(np.uint8)), to(last), 0.99, 0.95)
    want_adv, want_ret = gae_numpy(rew, val, done, last, 0.99, 0.95)
    assert np.allclose(adv.cpu().numpy(), want_adv, rtol=2e-4, atol=2e-3)
    assert np.allclose(ret.cpu().numpy(), want_ret, rtol=2e-4, atol=2e-3)


def test_collector_buffers_are_what_the_env_produced_and_graph_replay_matches_eager():
    from gym_roboy_b200.envs import RoboyEnv
    from gym_roboy_b200.envs.simulations import CudaSimulationClient
    from gym_roboy_b200.rollout import MlpPolicy, RolloutCollector
    n, T = 4096, 16
    outs = []
    for use_graph in (False, True):
        torch.manual_seed(0)
        torch.cuda.manual_seed(0)
        policy = MlpPolicy().to("cuda:0")
        client = CudaSimulationClient(num_envs=n, seed=11, device="cuda:0")
        env = RoboyEnv(client)
        col = RolloutCollector(env, policy, n_steps=T)
        client.set_step_num(torch.full((n,), 396, dtype=torch.int32))   # episodes end inside the rollout
        if use_graph:
            sd = client.state_dict()
            col.capture()
            client.load_state_dict(sd)                                   # undo the warm-up rollout
            col.obs[0].copy_(outs[0]["obs"][0])                          # ... and its last observation
        torch.cuda.manual_seed(123)
        col.collect()
        torch.cuda.synchronize()
        outs.append({k: getattr(col, k).clone() for k in ("obs", "actions", "rewards", "dones", "values", "adv", "ret")})
        # the buffers were written by the step kernel itself: replay them through the oracle
        if not use_graph:
            ora = orc.OracleEnv(n, seed=11)
            ora.reset()
            ora.step_flags[:] = (ora.step_flags & ~np.uint32(orc.STEP_MASK)) | np.uint32(396)
            for t in range(T):
                a = np.clip(col.actions[t].cpu().numpy(), -1.0, 1.0)         # stored un-clipped, clipped on the device for env.step
                o, r, d = ora.step(a)
                assert np.array_equal(col.obs[t + 1].cpu().numpy(), o)
                assert np.array_equal(col.dones[t].cpu().numpy().astype(bool), d)
                assert np.allclose(col.rewards[t].cpu().numpy(), r, rtol=1e-6, atol=0)
            assert col.dones.sum() >= n                                      # everyone timed out once
            want_adv, want_ret = gae_numpy(col.rewards.cpu().numpy(), col.values[:T].cpu().numpy(),
                                           col.dones.cpu().numpy().astype(bool), col.values[T].cpu().numpy(), 0.99, 0.95)
            assert np.allclose(col.adv.cpu().numpy(), want_adv, rtol=2e-4, atol=2e-3)
    # env trajectory under the graph is a valid rollout too (same shapes, episodes end, rewards in range)
    g = outs[1]
    assert g["dones"].sum() >= n and float(g["rewards"].max()) <= 999.0 and float(g["rewards"].min()) >= -33.0
    assert torch.isfinite(g["adv"]).all()


def test_vec_env_adapter():
    from gym_roboy_b200.vec_env import RoboyVecEnv
    n = 1000
    venv = RoboyVecEnv(n, seed=4, clip_actions=True)
    ora = orc.OracleEnv(n, seed=4)
    obs = venv.reset()
    assert obs.shape == (n, 9) and np.array_equal(obs, ora.reset())
    venv.client.set_step_num(np.full(n, 399, np.int32))
    ora.step_flags[:] = (ora.step_flags & ~np.uint32(orc.STEP_MASK)) | np.uint32(399)
    rng = np.random.default_rng(0)
    for t in range(3):
        a = rng.uniform(-1.5, 1.5, (n, 8)).astype(np.float32)       # the adapter clips like the SB runner
        obs, rew, done, infos = venv.step(a)
        o, r, d, term = ora.step(np.clip(a, -1, 1), want_terminal_obs=True)
        assert np.array_equal(obs, o) and np.array_equal(done, d) and np.allclose(rew, r, rtol=1e-6, atol=0)
        assert isinstance(infos, list) and len(infos) == n
        for i in np.flatnonzero(d):
            assert np.array_equal(infos[i]["terminal_observation"], term[i])
        assert all("terminal_observation" not in infos[i] for i in np.flatnonzero(~d))
    assert venv.client.errors() == (0, None)
    # per-env accessors, stable-baselines style
    assert venv.get_attr("step_num", indices=[0, 5]) == [int(venv.client.step_num[0]), int(venv.client.step_num[5])]
    assert len(venv.get_attr("reward_range")) == n and len(venv.env_method("render", indices=[1, 2, 3])) == 3
    venv.set_attr("step_num", 7, indices=[3])
    assert venv.get_attr("step_num", indices=3) == [7]
    venv.close()


def test_vec_env_adapter_asserts_like_the_reference_env_when_not_clipping():
    """roboy_env.py:52: the env asserts the action range (the SB runner clips BEFORE env.step); a NaN fails too."""
    from gym_roboy_b200.vec_env import RoboyVecEnv
    venv = RoboyVecEnv(64, seed=1)
    venv.reset()
    a = np.zeros((64, 8), np.float32)
    venv.step(a)
    a[17, 3] = 1.0000001
    with pytest.raises(AssertionError, match="17"):
        venv.step(a)
    a[17, 3] = np.nan
    with pytest.raises(AssertionError):
        venv.step(a)
    a[17, 3] = 1.0
    venv.step(a)
    venv.close()
