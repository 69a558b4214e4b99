"""Full-population parity on the BENCHMARKED configurations (BASELINE.json configs[2] = 262,144 envs and the
bench.py headline = 16,777,216 envs): every env of the CUDA path against the CPU oracle on the same seeded
inputs -- obs / done / goal / step word bit-exact, rewards within 1e-6 relative (north_star) -- plus the
done-index list against `nonzero(done)`, and a time-boxed slice of the randomised differential soak."""
import numpy as np
import pytest
import torch

from oracle import oracle as orc

pytestmark = pytest.mark.gpu
RTOL = 1e-6  # north_star tolerance for fp32 rewards

FLAG_COMBOS = [dict(), dict(penalty=True), dict(bonus=False), dict(auto_reset=False),
               dict(penalty=True, bonus=False, auto_reset=False)]
FLAG_IDS = ["default", "penalty", "nobonus", "manual_reset", "penalty_nobonus_manual"]


def make_pair(n, seed, **flags):
    from gym_roboy_b200.envs import RoboyEnv
    from gym_roboy_b200.envs.simulations import CudaSimulationClient
    client = CudaSimulationClient(num_envs=n, seed=seed, device="cuda:0")
    env = RoboyEnv(client, joint_vel_penalty=flags.get("penalty", False),
                   is_agent_getting_bonus_for_reaching_goal=flags.get("bonus", True),
                   auto_reset=flags.get("auto_reset", True), strict=False)
    ora = orc.OracleEnv(n, seed=seed, joint_vel_penalty=flags.get("penalty", False), bonus=flags.get("bonus", True),
                        auto_reset=flags.get("auto_reset", True), threads=16)
    return env, client, ora


def set_phases(client, ora, steps):
    client.set_step_num(steps)
    ora.step_flags[:] = (ora.step_flags & ~np.uint32(orc.STEP_MASK)) | steps.astype(np.uint32)


def compare(env, client, ora, a, t, check_done_index=True):
    obs, rew, done, _ = env.step(torch.as_tensor(a, device="cuda:0"))
    o_obs, o_rew, o_done = ora.step(a)
    d = done.cpu().numpy()
    assert np.array_equal(d, o_done), "done mask differs at step %d" % t
    assert np.array_equal(obs.cpu().numpy(), o_obs), "obs differ at step %d" % t
    rel = np.abs(rew.cpu().numpy().astype(np.float64) - o_rew) / np.maximum(np.abs(o_rew.astype(np.float64)), 1e-30)
    assert rel.max() <= RTOL, "reward rel err %g at step %d" % (rel.max(), t)
    if check_done_index:   # north_star: "bit-exact for done/reset masks AND indices"
        idx, _ = client.done_indices()
        assert np.array_equal(idx.cpu().numpy(), np.flatnonzero(o_done).astype(np.int32)), "done index list, step %d" % t
    return d


@pytest.mark.parametrize("flags", FLAG_COMBOS, ids=FLAG_IDS)
def test_262144_envs_50_steps_every_env_matches_oracle(flags):
    """BASELINE.json configs[2] as a population: 262,144 envs x 50 steps, all five flag combinations."""
    n, T, seed = 262_144, 50, 4321
    env, client, ora = make_pair(n, seed, **flags)
    client.enable_done_index(True)
    rng = np.random.default_rng(11)
    assert np.array_equal(env.reset().cpu().numpy(), ora.reset())
    set_phases(client, ora, rng.integers(1, 401, n).astype(np.int32))      # timeouts from the first step on
    pi = orc.PI32
    for t in range(T):
        if t % 10 == 3:   # goals next to the state the sampled branch will produce -> successes, both sides of the threshold
            q, _ = orc.draw_state(seed, np.arange(n), ora.counter + 1)
            off = rng.choice([0.003, 0.0314, 0.0315, 0.05], (n, 1)).astype(np.float32)
            g = np.clip(q + off, -pi, pi).astype(np.float32)
            client.set_goal(g); ora.goal[:] = g.T
        a = rng.uniform(-1, 1, (n, 8)).astype(np.float32)
        a[rng.random(n) < 0.01] = 0.0                                       # hold branch
        a[rng.random(n) < 0.001] = np.float32(2.0 ** -25)                   # upper edge of the hold interval
        d = compare(env, client, ora, a, t)
        if not flags.get("auto_reset", True) and d.any():
            m = d.astype(np.uint8)
            assert np.array_equal(env.reset(mask=torch.as_tensor(m)).cpu().numpy()[d], ora.reset(m)[d])
    assert np.array_equal(client.goal.cpu().numpy(), ora.goal)
    assert np.array_equal(client.step_flags.cpu().numpy().astype(np.uint32), ora.step_flags)
    s, so = client.stats(), ora.stats()
    for k in ("steps", "episodes", "successes", "timeouts", "sum_episode_len", "holds", "violations"):
        assert s[k] == so[k], (k, s[k], so[k])
    assert abs(s["sum_reward"] - so["sum_reward"]) <= 1e-6 * abs(so["sum_reward"])
    assert s["successes"] > 100 and s["timeouts"] > 1000 and s["holds"] > 1000
    assert client.errors()[0] == ora.errors()[0]


def test_headline_16777216_envs_every_env_matches_oracle():
    """The bench.py headline configuration (16,777,216 envs on one GPU) as a population, 2 steps."""
    n, seed = 1 << 24, 1234
    free, _ = torch.cuda.mem_get_info()
    if free < 8 << 30:
        pytest.skip("needs 8 GB of free HBM")
    env, client, ora = make_pair(n, seed)
    client.enable_done_index(True)
    assert np.array_equal(env.reset().cpu().numpy(), ora.reset())
    rng = np.random.default_rng(5)
    set_phases(client, ora, (np.arange(n, dtype=np.int64) % 400 + 1).astype(np.int32))
    total_done = 0
    for t in range(2):
        a = rng.random((n, 8), dtype=np.float32) * np.float32(2) - np.float32(1)
        a[:: 97] = 0.0
        d = compare(env, client, ora, a, t)
        total_done += int(d.sum())
    assert np.array_equal(client.goal.cpu().numpy(), ora.goal)
    assert np.array_equal(client.step_flags.cpu().numpy().astype(np.uint32), ora.step_flags)
    s, so = client.stats(), ora.stats()
    for k in ("steps", "episodes", "successes", "timeouts", "sum_episode_len", "holds", "violations"):
        assert s[k] == so[k], (k, s[k], so[k])
    assert total_done > 2 * n // 400 - 10 and s["holds"] >= 2 * (n // 97)
    client.close()


@pytest.mark.parametrize("n", [1, 31, 32, 33, 1000, 32768, 32769, 100_003])
def test_done_index_list_and_terminal_rows(n):
    """roboy_done_indices: ascending ids == nonzero(done), terminal rows == terminal_obs[ids], at ragged sizes and
    across several 32,768-env tiles; capacity smaller than the count drops the tail but reports the full count."""
    from gym_roboy_b200.envs import RoboyEnv
    from gym_roboy_b200.envs.simulations import CudaSimulationClient
    client = CudaSimulationClient(num_envs=n, seed=9, device="cuda:0")
    env = RoboyEnv(client, strict=False, auto_reset=True)
    env._single = False
    client.enable_terminal_obs(True)
    client.enable_done_index(True)
    ora = orc.OracleEnv(n, seed=9)
    env.reset(); ora.reset()
    rng = np.random.default_rng(n)
    steps = np.where(rng.random(n) < 0.3, 400, rng.integers(1, 300, n)).astype(np.int32)   # ~30 % time out at once
    set_phases(client, ora, steps)
    for t in range(3):
        a = rng.uniform(-1, 1, (n, 8)).astype(np.float32)
        obs, rew, done, _ = env.step(torch.as_tensor(a, device="cuda:0"))
        _, _, o_done, term = ora.step(a, want_terminal_obs=True)
        want = np.flatnonzero(o_done).astype(np.int32)
        idx, rows = client.done_indices(with_terminal_obs=True)
        assert np.array_equal(idx.cpu().numpy(), want), t
        assert np.array_equal(rows.cpu().numpy(), term[want]), t
        if want.size > 2:
            idx2, rows2 = client.done_indices(with_terminal_obs=True, capacity=want.size - 2)
            assert np.array_equal(idx2.cpu().numpy(), want[:-2]) and int(client._done_count.item()) == want.size
        if t == 1:   # nobody done: empty list
            set_phases(client, ora, np.full(n, 5, np.int32))
    assert client.done_indices()[0].numel() == 0


@pytest.mark.parametrize("mode", ["staged", "staged_noramp", "staged_split", "staged_one_stream", "mapped_out", "mapped_all"])
def test_host_buffer_modes_equal_the_device_step(mode):
    """roboy_step_host with copy-engine staging (with and without the short first stages), with the kernel storing
    straight into page-locked host memory, and with it reading the actions from there too: identical results."""
    from gym_roboy_b200 import _native
    from test_gpu_parity import actions_for
    n = 300_003   # several pipeline stages, ragged tail
    env_a, client_a, _ = make_pair(n, seed=3)
    env_b, client_b, _ = make_pair(n, seed=3)
    client_a.set_host_pipeline(stage_envs=1 << 15, n_streams=1 if mode == "staged_one_stream" else 3, ramp=mode != "staged_noramp",
                               ring=mode != "staged_split")
    client_a.set_host_mode({"mapped_out": _native.HOST_MAPPED_OUT, "mapped_all": _native.HOST_MAPPED_ALL}.get(mode, _native.HOST_STAGED))
    client_a.enable_done_index(True)
    env_a.reset(); env_b.reset()
    set_a = (np.arange(n) % 400 + 1).astype(np.int32)
    client_a.set_step_num(set_a); client_b.set_step_num(set_a)
    a_h, obs, rew, done = client_a.host_buffers(write_combined_actions=mode == "mapped_all")
    rng = np.random.default_rng(1)
    for t in range(3):
        a_h[...] = actions_for(rng, n)
        client_a.step_host(a_h, obs, rew, done)
        o2, r2, d2, _ = env_b.step(torch.as_tensor(np.array(a_h), device="cuda:0"))
        assert np.array_equal(obs, o2.cpu().numpy()) and np.array_equal(rew, r2.cpu().numpy())
        assert np.array_equal(done.astype(bool), d2.cpu().numpy())
        assert np.array_equal(client_a.done_indices()[0].cpu().numpy(), np.flatnonzero(done).astype(np.int32))
    assert client_a.counter == client_b.counter == 1 + 3
    assert torch.equal(client_a.goal, client_b.goal) and torch.equal(client_a.step_flags, client_b.step_flags)
    sa, sb = client_a.stats(), client_b.stats()
    for k in sa:
        assert sa[k] == sb[k] if k != "sum_reward" else abs(sa[k] - sb[k]) <= 1e-6 * abs(sb[k]), k
    # the copy probe moves the same bytes and leaves the env alone
    ms = client_a.copy_probe(a_h, obs, rew, done, directions=3, iters=2)
    assert ms > 0 and client_a.counter == 4
    assert client_a.copy_probe(a_h, obs, rew, done, directions=2, monolithic=True, iters=1) > 0
    client_a.close()


def test_mapped_host_mode_rejects_pageable_buffers():
    from gym_roboy_b200 import _native
    n = 4096
    env, client, _ = make_pair(n, seed=1)
    client.set_host_mode(_native.HOST_MAPPED_OUT)
    a = np.zeros((n, 8), np.float32)
    with pytest.raises(_native.RoboyNativeError, match="page-locked"):
        client.step_host(a, np.empty((n, 9), np.float32), np.empty(n, np.float32), np.empty(n, np.uint8))
    client.set_host_mode(_native.HOST_STAGED)
    client.step_host(a, np.empty((n, 9), np.float32), np.empty(n, np.float32), np.empty(n, np.uint8))   # pageable is fine here


def test_step_host_is_ordered_after_work_on_other_streams():
    """a reset / injection issued on a non-blocking side stream right before roboy_step_host must be seen by
    it (the call is device-synchronous)."""
    n = 200_000
    env_a, client_a, _ = make_pair(n, seed=8)
    env_b, client_b, _ = make_pair(n, seed=8)
    env_a.reset(); env_b.reset()
    goal = torch.zeros((n, 3), device="cuda:0") + 0.25
    steps = torch.full((n,), 400, dtype=torch.int32, device="cuda:0")
    side = torch.cuda.Stream()
    torch.cuda.synchronize()
    with torch.cuda.stream(side):
        torch.cuda._sleep(20_000_000)        # ~10 ms of busy GPU before the injection lands
        client_a.set_goal(goal)
        client_a.set_step_num(steps)
    client_b.set_goal(goal); client_b.set_step_num(steps)
    a = np.random.default_rng(0).uniform(-1, 1, (n, 8)).astype(np.float32)
    obs = np.empty((n, 9), np.float32); rew = np.empty(n, np.float32); done = np.empty(n, np.uint8)
    client_a.step_host(a, obs, rew, done)
    o2, r2, d2, _ = env_b.step(torch.as_tensor(a, device="cuda:0"))
    assert done.all() and np.array_equal(obs, o2.cpu().numpy()) and np.array_equal(rew, r2.cpu().numpy())


def test_checkpoint_carries_the_observation_and_checks_the_shard():
    """the last observation exists only in the obs buffer; a resumed closed-loop collector must continue
    the uninterrupted trajectory, and a checkpoint of another shard must be refused."""
    from gym_roboy_b200.envs import RoboyEnv
    from gym_roboy_b200.envs.simulations import CudaSimulationClient
    from gym_roboy_b200.rollout import MlpPolicy, RolloutCollector
    n, T = 4096, 6

    def collector(seed):
        torch.manual_seed(0)
        client = CudaSimulationClient(num_envs=n, seed=seed, device="cuda:0")
        return RolloutCollector(RoboyEnv(client), MlpPolicy().to("cuda:0"), n_steps=T, fused="fp32", noise_seed=5)

    a = collector(77)
    a.collect()
    sd = a.state_dict()
    assert "obs" in sd["env"] and sd["env"]["num_envs"] == n
    a.collect()
    want = (a.obs.clone(), a.actions.clone(), a.rewards.clone(), a.dones.clone())
    b = collector(123)                       # different seed, fresh obs: everything must come from the checkpoint
    b.load_state_dict(sd)
    assert torch.equal(b.client.obs, sd["env"]["obs"])
    b.collect()
    for x, y in zip(want, (b.obs, b.actions, b.rewards, b.dones)):
        assert torch.equal(x, y)
    other = CudaSimulationClient(num_envs=n, seed=1, env_id_base=n, device="cuda:0")
    with pytest.raises(ValueError, match="env_id_base"):
        other.load_state_dict(sd["env"])


def test_env_seed_redraws_the_first_goal():
    """RoboyEnv(client, seed=s) must start from the goal / held state a client built with seed s starts from
    (the reference seeds before it draws its first goal, roboy_env.py:15 then :37)."""
    from gym_roboy_b200.envs import RoboyEnv
    from gym_roboy_b200.envs.simulations import CudaSimulationClient
    n = 1000
    c1 = CudaSimulationClient(num_envs=n, device="cuda:0")            # os.urandom seed
    e1 = RoboyEnv(c1, seed=99)
    c2 = CudaSimulationClient(num_envs=n, seed=99, device="cuda:0")
    e2 = RoboyEnv(c2)
    assert torch.equal(c1.goal, c2.goal) and torch.equal(c1.held, c2.held) and c1.counter == c2.counter == 0
    a = torch.rand((n, 8), device="cuda:0") * 2 - 1
    o1, r1, d1, _ = e1.step(a)
    o2, r2, d2, _ = e2.step(a)
    assert torch.equal(o1, o2) and torch.equal(r1, r2) and torch.equal(d1, d2)


def test_episode_length_statistic_without_auto_reset():
    """sum_episode_len is accumulated with auto_reset off and on the external-feed path too."""
    n = 2048
    env, client, ora = make_pair(n, seed=2, auto_reset=False)
    env.reset(); ora.reset()
    set_phases(client, ora, np.full(n, 400, np.int32))
    a = np.random.default_rng(0).uniform(-1, 1, (n, 8)).astype(np.float32)
    compare(env, client, ora, a, 0, check_done_index=False)
    s = client.stats()
    assert s["episodes"] == n and s["sum_episode_len"] == 400 * n == ora.stats()["sum_episode_len"]
    from gym_roboy_b200.sharding import summarize
    summ = summarize(client.stats_tensor)
    assert summ["mean_episode_len"] == 400.0 and summ["mean_episode_return"] is not None
    q = np.zeros((n, 3), np.float32); qd = np.zeros((n, 3), np.float32)
    client.clear_stats()
    client.step_external(q, qd)
    assert client.stats()["sum_episode_len"] == 401 * n      # still done (no reset in between): one step longer


def test_randomised_soak_slice():
    """20 s of tools/soak_parity.py: random seeds, sizes, env-id bases and flags, goals planted around the reached
    threshold; zero mismatches allowed."""
    import importlib.util
    import os
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools", "soak_parity.py")
    spec = importlib.util.spec_from_file_location("soak_parity", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    master_seed = int(np.random.SeedSequence().entropy % (2 ** 32))
    summary = mod.soak(20.0, master_seed=master_seed)
    assert summary["mismatches"] == [], (master_seed, summary["mismatches"][:3])
    assert summary["configs"] >= 5 and summary["successes"] > 0 and summary["timeouts"] > 0 and summary["holds"] > 0
    assert summary["worst_reward_rel"] <= RTOL


def test_host_path_soak_slice():
    """12 seconds of tools/soak_host.py: `roboy_step_host` under random stage sizes, stream counts, ramp / pattern / mode,
    autotune, done-index lists and terminal rows, MSJ and other robots, device steps in between -- against the oracle."""
    import os, sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    from soak_host import soak
    summary = soak(12.0, master_seed=11)
    assert not summary["mismatches"], summary["mismatches"][:3]
    assert summary["configs"] >= 3


def test_api_soak_slice():
    """12 seconds of tools/soak_api.py: stand-alone compute_reward, the external-simulator feed and state / goal / step
    injection on values salted with NaN, infinities, the bounds, 3e38 and denormals, for MSJ, MSJ-shaped robots with other
    limits and other robots -- against the oracle (NaN payloads included)."""
    import os, sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    from soak_api import soak
    summary = soak(12.0, master_seed=5)
    assert not summary["mismatches"], [{k: v for k, v in m.items() if k != "bounds"} for m in summary["mismatches"][:3]]
    assert summary["configs"] >= 3 and summary["rewards_checked"] > 0 and summary["injected_steps"] > 0


def test_nan_velocity_in_a_held_state_makes_the_penalty_reward_nan():
    """roboy_env.py:99 takes np.linalg.norm of the velocity difference directly -- no NaN -> 0 as in _l2_distance (:139): an
    env holding a state whose velocity has a NaN component gets a NaN reward with joint_vel_penalty on (and a finite one
    without).  The tuned kernels' hold path used the guarded norm there; found by tools/soak_api.py."""
    import numpy as np, torch
    from gym_roboy_b200.envs import RoboyEnv
    from gym_roboy_b200.envs.simulations import CudaSimulationClient
    from oracle import oracle as orc
    for penalty in (True, False):
        n = 64
        client = CudaSimulationClient(num_envs=n, seed=3, device="cuda:0")
        env = RoboyEnv(client, joint_vel_penalty=penalty, auto_reset=False, strict=False)
        ora = orc.OracleEnv(n, seed=3, joint_vel_penalty=penalty, auto_reset=False)
        env.reset(); ora.reset()
        q = np.full((n, 3), 0.25, np.float32); qd = np.full((n, 3), 0.1, np.float32)
        qd[::2, 1] = np.nan
        client.set_state(q, qd, np.ones(n, np.uint8))
        ora.held[0:3] = q.T; ora.held[3:6] = qd.T
        ora.step_flags[:] = ora.step_flags & np.uint32(orc.STEP_MASK)
        a = np.zeros((n, 8), np.float32)                       # hold: the injected state is the env's state
        obs, rew, done, _ = env.step(torch.as_tensor(a, device="cuda:0"))
        o_obs, o_rew, o_done = ora.step(a)
        r = rew.cpu().numpy()
        assert np.array_equal(np.isnan(r), np.isnan(o_rew)) and np.isnan(r[::2]).all() == penalty and not np.isnan(r[1::2]).any()
        assert np.allclose(r[1::2], o_rew[1::2], rtol=1e-6, atol=0)
        assert np.array_equal(obs.cpu().numpy().view(np.uint32), o_obs.view(np.uint32)) and np.array_equal(done.cpu().numpy(), o_done)
        assert client.stats()["violations"] == ora.stats()["violations"] == (n // 2 if penalty else 0)
