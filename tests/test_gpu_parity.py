"""CUDA path vs the CPU oracle on identical seeded inputs (bit-exact done/obs/state, rewards to
1e-6 relative), at sizes the oracle finishes in seconds, plus size-independent properties at
BASELINE.json's full sizes."""
import numpy as np
import pytest
import torch

from oracle import oracle as orc

pytestmark = pytest.mark.gpu
RTOL = 1e-6  # north_star tolerance for fp32 rewards


def to_np(x):
    return x.cpu().numpy() if torch.is_tensor(x) else np.asarray(x)


def make_pair(n, seed, **flags):
    from gym_roboy_b200.envs import RoboyEnv
    from gym_roboy_b200.envs.simulations import CudaSimulationClient
    base = flags.pop("env_id_base", 0)
    client = CudaSimulationClient(num_envs=n, seed=seed, env_id_base=base, device="cuda:0")
    env = RoboyEnv(client, joint_vel_penalty=flags.get("penalty", False),
                   is_agent_getting_bonus_for_reaching_goal=flags.get("bonus", True),
                   auto_reset=flags.get("auto_reset", True), strict=False)
    ora = orc.OracleEnv(n, seed=seed, env_id_base=base, joint_vel_penalty=flags.get("penalty", False),
                        bonus=flags.get("bonus", True), auto_reset=flags.get("auto_reset", True), threads=8)
    return env, client, ora


def actions_for(rng, n, hold_frac=0.01):
    a = rng.uniform(-1, 1, (n, 8)).astype(np.float32)
    a[rng.random(n) < hold_frac] = 0.0
    up, dn = np.float32(2.0 ** -25), np.float32(-(2.0 ** -24))
    k = min(n, 8)
    rows = [np.full(8, up), np.full(8, np.nextafter(up, np.float32(1))), np.full(8, dn),
            np.full(8, np.nextafter(dn, np.float32(-1))), np.array([0, 0, 0, 0, 0, 0, 0, 1e-7]),
            np.full(8, 1.0), np.full(8, -1.0), np.array([up, dn, 0, 0, dn, up, 0, 0])]
    idx = rng.choice(n, size=k, replace=False)
    for i, r in zip(idx, rows):
        a[i] = r.astype(np.float32)
    return a


def compare_step(env, client, ora, a, t):
    obs, rew, done, _ = env.step(torch.as_tensor(a, device="cuda:0"))
    o_obs, o_rew, o_done = ora.step(a)
    assert np.array_equal(done.cpu().numpy(), o_done), "done mask differs at step %d" % t
    assert np.array_equal(obs.cpu().numpy(), o_obs), "obs differ at step %d" % t
    r = rew.cpu().numpy().astype(np.float64)
    rel = np.abs(r - o_rew) / np.maximum(np.abs(o_rew.astype(np.float64)), 1e-30)
    assert rel.max() <= RTOL, "reward rel err %g at step %d" % (rel.max(), t)
    return float(rel.max())


@pytest.mark.parametrize("n", [1, 31, 32, 33, 1000, 4096])
def test_construction_and_reset_match_oracle(n):
    env, client, ora = make_pair(n, seed=99)
    assert np.array_equal(client.goal.cpu().numpy(), ora.goal)
    assert np.array_equal(client.held.cpu().numpy(), ora.held)
    assert np.array_equal(client.step_flags.cpu().numpy().astype(np.uint32), ora.step_flags)
    obs = env.reset()
    assert np.array_equal(to_np(obs).reshape(n, 9), ora.reset())
    assert np.array_equal(client.step_flags.cpu().numpy().astype(np.uint32), ora.step_flags)
    mask = (np.arange(n) % 3 == 0).astype(np.uint8)
    obs = to_np(env.reset(mask=torch.as_tensor(mask))).reshape(n, 9)
    o = ora.reset(mask)
    assert np.array_equal(obs[mask.astype(bool)], o[mask.astype(bool)])
    assert np.array_equal(client.goal.cpu().numpy(), ora.goal)


@pytest.mark.parametrize("flags", [dict(), dict(penalty=True), dict(bonus=False), dict(auto_reset=False),
                                   dict(penalty=True, bonus=False, auto_reset=False)],
                         ids=["default", "penalty", "nobonus", "manual_reset", "penalty_nobonus_manual"])
def test_rollout_4096_envs_matches_oracle(flags):
    """configs[1]: 4,096 envs, T > 2 x 400 so the timeout reset fires twice per env."""
    n, T = 4096, 1001
    env, client, ora = make_pair(n, seed=1234, **flags)
    rng = np.random.default_rng(7)
    env.reset(); ora.reset()
    # de-synchronise episode phases and plant goals the sampled state will hit
    steps = rng.integers(1, 400, n).astype(np.int32)
    client.set_step_num(steps)
    ora.step_flags[:] = (ora.step_flags & ~np.uint32(orc.STEP_MASK)) | steps.astype(np.uint32)
    worst = 0.0
    for t in range(T):
        if t % 50 == 10:   # goals next to the state the Philox draw will produce -> goal reached on the sampled branch
            c = ora.counter + 1
            q, _ = orc.draw_state(1234, np.arange(n), c)
            g = np.clip(q + np.float32(0.005), -orc.PI32, orc.PI32).astype(np.float32)
            client.set_goal(g); ora.goal[:] = g.T
        worst = max(worst, compare_step(env, client, ora, actions_for(rng, n), t))
        if not flags.get("auto_reset", True):
            d = client.done.cpu().numpy()
            if d.any():
                env.reset(mask=torch.as_tensor(d.astype(np.uint8))); ora.reset(d.astype(np.uint8))
        if t % 100 == 0:
            assert np.array_equal(client.goal.cpu().numpy(), ora.goal)
            assert np.array_equal(client.step_flags.cpu().numpy().astype(np.uint32), ora.step_flags)
    assert np.array_equal(client.goal.cpu().numpy(), ora.goal)
    assert np.array_equal(client.step_flags.cpu().numpy().astype(np.uint32), ora.step_flags)
    s, so = client.stats(), ora.stats()
    for k in ("steps", "episodes", "successes", "timeouts", "sum_episode_len", "holds", "violations"):
        assert s[k] == so[k], (k, s[k], so[k])
    assert abs(s["sum_reward"] - so["sum_reward"]) <= 1e-6 * abs(so["sum_reward"])
    assert s["successes"] > 0 and s["timeouts"] > 2 * n * 0.9 and s["holds"] > 0
    assert client.errors()[0] == ora.errors()[0]
    if client.errors()[0]:
        assert client.errors()[1] == ora.errors()[1]


def test_held_state_injection_and_thresholds():
    """Near-threshold goal distances on the hold branch, float32 and float64 held states,
    infeasible penalty -- compared env by env with the oracle."""
    n = 2048
    env, client, ora = make_pair(n, seed=5)
    env.reset(); ora.reset()
    rng = np.random.default_rng(3)
    thr_a, thr_v = orc.thresholds(ora.cfg)
    q = rng.uniform(-3, 3, (n, 3)).astype(np.float32)
    scale = (1 + rng.uniform(-4e-7, 4e-7, n)).astype(np.float32)
    direction = rng.normal(size=(n, 3)); direction /= np.linalg.norm(direction, axis=1, keepdims=True)
    goal = np.clip(q + (direction * thr_a * scale[:, None]).astype(np.float32), -orc.PI32, orc.PI32).astype(np.float32)
    vdir = rng.normal(size=(n, 3)); vdir /= np.linalg.norm(vdir, axis=1, keepdims=True)
    vscale = np.where(rng.random(n) < 0.5, 1 + rng.uniform(-4e-7, 4e-7, n), rng.uniform(0, 0.9, n))
    qd = (vdir * thr_v * vscale[:, None]).astype(np.float32)
    feasible = (rng.random(n) > 0.2)
    half = n // 2   # first half: float32 held states; second half keeps the float64 zero state
    idx = np.arange(half)
    client.set_state(q[:half], qd[:half], feasible[:half].astype(np.uint8), idx=idx)
    ora.held[0:3, :half] = q[:half].T; ora.held[3:6, :half] = qd[:half].T
    ora.step_flags[:half] = (ora.step_flags[:half] & np.uint32(orc.STEP_MASK)) | \
        np.where(feasible[:half], 0, orc.F_HELD_INFEASIBLE).astype(np.uint32)
    goal[half:] = (direction[half:] * thr_a * scale[half:, None]).astype(np.float32)  # around the zero state
    client.set_goal(goal); ora.goal[:] = goal.T
    a = np.zeros((n, 8), np.float32)
    compare_step(env, client, ora, a, 0)
    d = client.done.cpu().numpy()
    assert 0.2 < d[:half].mean() < 0.8 and 0.2 < d[half:].mean() < 0.8   # the threshold really was straddled
    assert client.stats()["holds"] == n


def test_compute_reward_kats():
    """SURVEY.md 8a KAT table (values produced by the reference itself) + oracle on random states."""
    from gym_roboy_b200.envs import RoboyEnv
    from gym_roboy_b200.envs.robots import RobotState
    from gym_roboy_b200.envs.simulations import CudaSimulationClient
    kats = [  # q, qd, goal, feasible, reward(pen=False), reward(pen=True), reached
        ((0.5, -1.0, 2.0), (0.1, -0.2, 0.3), (0.25, -0.75, 1.5), True, -1.2152187824249268, -2.5922478480921027, False),
        ((0.26, -0.74, 1.51), (0.05, 0, -0.05), (0.25, -0.75, 1.5), True, 998.9944458007812, 998.4434189188644, True),
        ((0.26, -0.74, 1.51), (0.3, 0.3, 0.3), (0.25, -0.75, 1.5), True, -1.0055285692214966, -2.7323260694901483, False),
        ((3, -3, 3), (0.5, -0.5, 0.5), (-3, 3, -3), False, -28.329681396484375, -73.53261015329933, False),
        ((1, 1, 1), (0, 0, 0), (1, 1, 1), True, 999.0, 998.6321206092834, True),
    ]
    for pen in (False, True):
        env = RoboyEnv(CudaSimulationClient(num_envs=1, seed=0, device="cuda:0"), joint_vel_penalty=pen)
        for q, qd, g, feas, r0, r1, reached in kats:
            cur = RobotState(np.array(q, np.float32), np.array(qd, np.float32), feas)
            goal = RobotState(np.array(g, np.float32), np.zeros(3), True)   # float64 zero goal velocities
            want = r1 if pen else r0
            got = env.compute_reward(cur, goal)
            assert isinstance(got, float)
            assert abs(got - want) <= RTOL * abs(want), (q, pen, got, want)
            assert env._did_reach_goal(cur, goal) is reached
    # reward_range pins (SURVEY.md 8a a9 / test_roboy_env.py:82-89)
    env = RoboyEnv(CudaSimulationClient(num_envs=1, seed=0, device="cuda:0"))
    # (numpy's float32 exp and CUDA's expf may differ by an ulp, hence rtol and not ==)
    assert np.allclose(env.reward_range, (-32.947744369506836, 999.0), rtol=RTOL, atol=0)
    envp = RoboyEnv(CudaSimulationClient(num_envs=1, seed=0, device="cuda:0"), joint_vel_penalty=True)
    assert np.allclose(envp.reward_range, (-143.61798095703125, 998.6321411132812), rtol=RTOL, atol=0)
    envn = RoboyEnv(CudaSimulationClient(num_envs=1, seed=0, device="cuda:0"),
                    is_agent_getting_bonus_for_reaching_goal=False)
    assert np.allclose(envn.reward_range, (-32.947744369506836, -1.0), rtol=RTOL, atol=0)
    # random batch vs the oracle, float32 goal velocities (all-float32 numpy path) and float64-zero ones
    rng = np.random.default_rng(0)
    k = 20000
    q = rng.uniform(-np.pi, np.pi, (k, 3)).astype(np.float32)
    g = np.where(rng.random((k, 1)) < 0.3, q + rng.uniform(-0.04, 0.04, (k, 3)), rng.uniform(-3.1, 3.1, (k, 3))).astype(np.float32)
    qd = (rng.uniform(-0.5, 0.5, (k, 3)) * rng.random((k, 1))).astype(np.float32)
    gqd = rng.uniform(-0.1, 0.1, (k, 3)).astype(np.float32)
    feas = rng.random(k) > 0.1
    for pen in (False, True):
        client = CudaSimulationClient(num_envs=1, seed=0, device="cuda:0")
        client.configure_env(pen, True, True)
        cfg = orc.make_cfg(1, joint_vel_penalty=pen)
        for goal_qd in (None, gqd):
            r, reached = client.compute_reward(q, qd, feas.astype(np.uint8), g, goal_qd)
            ro, reached_o, _ = orc.compute_reward(cfg, q, qd, feas, g, goal_qd)
            assert np.array_equal(reached.cpu().numpy(), reached_o)
            assert reached_o.sum() > 100
            rel = np.abs(r.cpu().numpy() - ro) / np.abs(ro)
            assert rel.max() <= RTOL


def test_sharding_invariance_and_determinism():
    """Philox counters use the global env id: two half-shards == one full shard, bit for bit;
    and the same seed replays the same rollout."""
    n, T = 8192, 60
    rng = np.random.default_rng(11)
    acts = [actions_for(rng, n) for _ in range(T)]
    outs = {}
    for label, shards in (("full", [(0, n)]), ("halves", [(0, n // 2), (n // 2, n)]), ("again", [(0, n)])):
        pieces = []
        for b, e in shards:
            env, client, _ = make_pair(e - b, seed=77, env_id_base=b)
            env.reset()
            client.set_step_num(np.full(e - b, 380, np.int32))
            rec = []
            for t in range(T):
                obs, rew, done, _ = env.step(torch.as_tensor(acts[t][b:e], device="cuda:0"))
                rec.append((obs.cpu().numpy().copy(), rew.cpu().numpy().copy(), done.cpu().numpy().copy()))
            pieces.append((rec, client.stats()))
        outs[label] = pieces
    for t in range(T):
        for k in range(3):
            full = outs["full"][0][0][t][k]
            assert np.array_equal(full, np.concatenate([p[0][t][k] for p in outs["halves"]]))
            assert np.array_equal(full, outs["again"][0][0][t][k])
    for key in ("steps", "episodes", "successes", "timeouts", "holds"):
        assert outs["full"][0][1][key] == sum(p[1][key] for p in outs["halves"])
    assert outs["full"][0][1]["episodes"] >= n


def test_host_buffer_step_equals_device_step():
    n = 300_000   # more than one pipeline stage, ragged tail
    env_a, client_a, _ = make_pair(n, seed=3)
    env_b, client_b, _ = make_pair(n, seed=3)
    client_a.set_host_pipeline(stage_envs=1 << 16, n_streams=3)
    env_a.reset(); env_b.reset()
    rng = np.random.default_rng(1)
    for t in range(3):
        a = actions_for(rng, n)
        obs = np.empty((n, 9), np.float32); rew = np.empty(n, np.float32); done = np.empty(n, np.uint8)
        client_a.step_host(a, obs, rew, done)
        o2, r2, d2, _ = env_b.step(torch.as_tensor(a, device="cuda:0"))
        assert np.array_equal(obs, o2.cpu().numpy()) and np.array_equal(rew, r2.cpu().numpy())
        assert np.array_equal(done.astype(bool), d2.cpu().numpy())
    sa, sb = client_a.stats(), client_b.stats()
    for k in sa:   # counters are exact; the reward sum is reduced in a different order per launch shape
        assert sa[k] == sb[k] if k != "sum_reward" else abs(sa[k] - sb[k]) <= 1e-6 * abs(sb[k]), k


def test_outputs_are_zero_copy_views_of_the_library_buffers():
    import ctypes
    from gym_roboy_b200 import _native
    env, client, _ = make_pair(1024, seed=1)
    for which, t in ((_native.BUF_OBS, client.obs), (_native.BUF_REWARD, client.reward), (_native.BUF_DONE, client.done_u8),
                     (_native.BUF_GOAL, client.goal), (_native.BUF_STEP_FLAGS, client.step_flags)):
        ptr, nbytes = ctypes.c_void_p(), ctypes.c_uint64()
        _native.check(client._lib.roboy_buffer(client._h, which, ctypes.byref(ptr), ctypes.byref(nbytes)))
        assert t.data_ptr() == ptr.value and t.numel() * t.element_size() == nbytes.value and t.is_cuda
    obs, rew, done, _ = env.step(torch.zeros((1024, 8), device="cuda:0"))
    assert obs.data_ptr() == client.obs.data_ptr() and done.dtype == torch.bool


def test_checkpoint_resume_is_exact():
    n = 5000
    env, client, _ = make_pair(n, seed=21)
    env.reset()
    rng = np.random.default_rng(2)
    acts = [torch.as_tensor(actions_for(rng, n), device="cuda:0") for _ in range(20)]
    for a in acts[:10]:
        env.step(a)
    sd = client.state_dict()
    tail = []
    for a in acts[10:]:
        o, r, d, _ = env.step(a)
        tail.append((o.clone(), r.clone(), d.clone()))
    env2, client2, _ = make_pair(n, seed=999)   # different seed: everything must come from the checkpoint
    client2.load_state_dict(sd)
    for a, (o, r, d) in zip(acts[10:], tail):
        o2, r2, d2, _ = env2.step(a)
        assert torch.equal(o, o2) and torch.equal(r, r2) and torch.equal(d, d2)


@pytest.mark.parametrize("n", [262_144, 4_194_304])
def test_full_size_properties(n):
    """configs[2]/[3] sizes: properties that need no oracle."""
    env, client, _ = make_pair(4, seed=0)  # noqa: F841 (warm the library)
    from gym_roboy_b200.envs import RoboyEnv
    from gym_roboy_b200.envs.simulations import CudaSimulationClient
    client = CudaSimulationClient(num_envs=n, seed=42, device="cuda:0")
    env = RoboyEnv(client)
    env.reset()
    client.set_step_num(torch.randint(1, 400, (n,), dtype=torch.int32, device="cuda:0"))
    gen = torch.Generator(device="cuda:0"); gen.manual_seed(0)
    pi = float(orc.PI32)
    total_done = 0
    for t in range(5):
        goal_before = client.goal.clone()
        steps_before = client.step_num.clone()
        a = torch.rand((n, 8), device="cuda:0", generator=gen) * 2 - 1
        obs, rew, done, _ = env.step(a)
        assert obs.shape == (n, 9) and rew.shape == (n,) and done.shape == (n,)
        nd = ~done
        assert torch.equal(obs[nd][:, 6:9], goal_before.t()[nd])          # obs carries the goal in force
        assert torch.equal(obs[done][:, 6:9], client.goal.t()[done])      # reset obs carries the NEW goal
        assert (obs[done][:, 0:6] == 0).all()                             # reset obs: zero state
        assert (obs[:, 0:6].abs() <= pi).all() and (obs[:, 6:9].abs() <= pi).all()
        assert torch.equal(client.step_num[nd], steps_before[nd] + 1)
        assert (client.step_num[done] == 1).all()
        lo, hi = env.reward_range
        assert (rew >= lo).all() and (rew <= hi).all()
        assert (rew[nd] <= -1.0).all()                                    # -exp(d) <= -1 without the bonus
        total_done += int(done.sum())
    s = client.stats()
    assert s["steps"] == 5 * n and s["episodes"] == total_done and s["violations"] == 0
    assert s["episodes"] == s["successes"] + s["timeouts"] and total_done > 0
    assert client.errors() == (0, None)


def test_steps_captured_in_a_cuda_graph_replay_like_eager_steps():
    """The call counter is device state, so a captured sequence of steps can be replayed and each
    replay continues the trajectory exactly as eager stepping would."""
    n, k, rounds = 20000, 8, 3
    rng = np.random.default_rng(5)
    acts = [torch.as_tensor(actions_for(rng, n), device="cuda:0") for _ in range(k)]
    env_a, client_a, _ = make_pair(n, seed=8)
    env_b, client_b, _ = make_pair(n, seed=8)
    for env, client in ((env_a, client_a), (env_b, client_b)):
        env.reset()
        client.set_step_num(np.full(n, 390, np.int32))     # timeouts (goal draws) inside the captured region
    obs_buf = torch.zeros((k, n, 9), device="cuda:0")
    rew_buf = torch.zeros((k, n), device="cuda:0")
    done_buf = torch.zeros((k, n), dtype=torch.uint8, device="cuda:0")
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        client_b.step_fused(acts[0])                        # warm-up outside capture (occupancy query etc.)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    client_b.load_state_dict(client_a.state_dict())         # undo the warm-up step exactly
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        for t in range(k):
            client_b.step_fused(acts[t], obs=obs_buf[t], reward=rew_buf[t], done=done_buf[t])
    client_b.load_state_dict(client_a.state_dict())         # capture does not execute; be explicit anyway
    for r in range(rounds):
        graph.replay()
        torch.cuda.synchronize()
        for t in range(k):
            o, rw, d, _ = env_a.step(acts[t])
            assert torch.equal(o, obs_buf[t]) and torch.equal(rw, rew_buf[t]) and torch.equal(d, done_buf[t].bool()), (r, t)
    assert client_a.counter == client_b.counter == 1 + rounds * k
    assert client_a.stats()["episodes"] == client_b.stats()["episodes"] > 0


OTHER_ROBOTS = {
    "symmetric": dict(angle_low=-2.0, angle_high=2.0, vel_low=-0.7, vel_high=0.7, act_low=-0.5, act_high=0.5),
    "asymmetric": dict(angle_low=-1.0, angle_high=2.5, vel_low=-0.25, vel_high=0.75, act_low=-0.1, act_high=0.4),
}


@pytest.mark.parametrize("name", sorted(OTHER_ROBOTS))
@pytest.mark.parametrize("penalty", [False, True])
def test_other_robot_bounds_match_oracle(name, penalty):
    """a RoboyRobot plug-in with other spaces.  The kernel takes the generic
    IEEE-division path (the 3-instruction division is only proved for the MSJ spans) and a different
    hold interval; the oracle for these robots is pinned against the reference in
    tests/test_oracle_vs_reference.py::test_other_robot_bounds_match_reference."""
    from gym_roboy_b200.envs import RoboyEnv
    from gym_roboy_b200.envs.robots import RoboyRobot
    from gym_roboy_b200.envs.simulations import CudaSimulationClient
    from gym_roboy_b200.spaces import Box
    b = OTHER_ROBOTS[name]

    class OtherRobot(RoboyRobot):
        _A = Box(b["angle_low"], b["angle_high"], (3,), "float32")
        _V = Box(b["vel_low"], b["vel_high"], (3,), "float32")
        _T = Box(b["act_low"], b["act_high"], (8,), "float32")
        get_action_space = classmethod(lambda cls: cls._T)
        get_joint_angles_space = classmethod(lambda cls: cls._A)
        get_joint_vels_space = classmethod(lambda cls: cls._V)

    n, T = 4096, 450
    client = CudaSimulationClient(robot=OtherRobot(), num_envs=n, seed=6, device="cuda:0")
    env = RoboyEnv(client, joint_vel_penalty=penalty, strict=False)
    ora = orc.OracleEnv(n, seed=6, joint_vel_penalty=penalty, threads=8, **b)
    assert np.allclose(env.reward_range, ora.reward_range, rtol=RTOL, atol=0)
    assert np.array_equal(client.goal.cpu().numpy(), ora.goal) and np.array_equal(client.held.cpu().numpy(), ora.held)
    env.reset(); ora.reset()
    rng = np.random.default_rng(2)
    zero_action = np.float32(1 - 2 * b["act_high"] / (b["act_high"] - b["act_low"]))
    steps = rng.integers(1, 400, n).astype(np.int32)
    client.set_step_num(steps)
    ora.step_flags[:] = (ora.step_flags & ~np.uint32(orc.STEP_MASK)) | steps.astype(np.uint32)
    for t in range(T):
        a = rng.uniform(-1, 1, (n, 8)).astype(np.float32)
        hold_rows = rng.random(n) < 0.02
        a[hold_rows] = zero_action
        a[rng.random(n) < 0.01] = np.nextafter(zero_action, np.float32(1))   # right next to the hold interval
        if t % 40 == 5:
            q, _ = orc.draw_state(6, np.arange(n), ora.counter + 1, b["angle_low"], b["angle_high"])
            g = np.clip(q + np.float32(0.004), b["angle_low"], b["angle_high"]).astype(np.float32)
            client.set_goal(g); ora.goal[:] = g.T
        compare_step(env, client, ora, a, t)
    assert np.array_equal(client.goal.cpu().numpy(), ora.goal)
    assert np.array_equal(client.step_flags.cpu().numpy().astype(np.uint32), ora.step_flags)
    s, so = client.stats(), ora.stats()
    for k in ("steps", "episodes", "successes", "timeouts", "holds", "violations"):
        assert s[k] == so[k], k
    assert s["holds"] > 0 and s["timeouts"] > 0


@pytest.mark.parametrize("n", [1, 2, 31, 32, 33, 63, 95, 255, 257, 1000, 4097])
def test_ragged_sizes_with_guarded_output_buffers(n):
    """compute-sanitizer is not available on the GPU pool, so out-of-bounds stores are hunted with
    canaries: the step writes into caller tensors that sit inside larger, pre-filled buffers."""
    from gym_roboy_b200.envs import RoboyEnv
    from gym_roboy_b200.envs.simulations import CudaSimulationClient
    client = CudaSimulationClient(num_envs=n, seed=13, device="cuda:0")
    env = RoboyEnv(client, strict=False, auto_reset=True)
    env._single = False
    ora = orc.OracleEnv(n, seed=13)
    env.reset(); ora.reset()
    steps = (np.arange(n) % 5 + 397).astype(np.int32)
    client.set_step_num(steps)
    ora.step_flags[:] = (ora.step_flags & ~np.uint32(orc.STEP_MASK)) | steps.astype(np.uint32)
    pad, canary = 64, 7777.0
    obs_big = torch.full((n + 2 * pad, 9), canary, device="cuda:0")
    rew_big = torch.full((n + 2 * pad,), canary, device="cuda:0")
    done_big = torch.full((n + 2 * pad,), 77, dtype=torch.uint8, device="cuda:0")
    # the obs view is 16-byte aligned for even n and only 4-byte aligned for odd n (offset (pad+n%2)*36 B):
    # both the bulk-store path and the scalar fallback get exercised
    off = pad + (n % 2)
    obs, rew, done = obs_big[off:off + n], rew_big[pad:pad + n], done_big[pad:pad + n]
    assert (obs.data_ptr() % 16 == 0) == (n % 2 == 0)
    rng = np.random.default_rng(n)
    for t in range(6):
        a = rng.uniform(-1, 1, (n, 8)).astype(np.float32)
        a[rng.random(n) < 0.2] = 0
        client.step_fused(torch.as_tensor(a, device="cuda:0"), obs=obs, reward=rew, done=done)
        o, r, d = ora.step(a)
        assert np.array_equal(obs.cpu().numpy(), o) and np.array_equal(done.cpu().numpy().astype(bool), d)
        assert np.allclose(rew.cpu().numpy(), r, rtol=RTOL, atol=0)
        assert (obs_big[:off] == canary).all() and (obs_big[off + n:] == canary).all(), "obs store outside its range"
        for big, val in ((rew_big, canary), (done_big, 77)):
            assert (big[:pad] == val).all() and (big[pad + n:] == val).all(), "store outside the output range"
    assert np.array_equal(client.goal.cpu().numpy(), ora.goal)
    assert np.array_equal(client.step_flags.cpu().numpy().astype(np.uint32), ora.step_flags)
    assert client.stats()["steps"] == 6 * n and client.stats()["episodes"] == ora.stats()["episodes"] > 0


@pytest.mark.parametrize("n,T", [(4096, 37), (1000, 12), (33, 420), (262144, 8)])
def test_open_loop_step_many_equals_sequential_steps(n, T):
    """roboy_step_many (state in registers across T steps, 73 B/env-step) must be bit-identical to T
    roboy_step launches on the same pre-recorded actions -- including episode ends inside the window."""
    rng = np.random.default_rng(n + T)
    acts = np.stack([actions_for(rng, n, hold_frac=0.05) for _ in range(T)])
    env_a, client_a, _ = make_pair(n, seed=31)
    env_b, client_b, _ = make_pair(n, seed=31)
    for env, client in ((env_a, client_a), (env_b, client_b)):
        env._single = False
        env.reset()
        client.set_step_num((np.arange(n) % 400 + 1).astype(np.int32))
    a_dev = torch.as_tensor(acts, device="cuda:0")
    obs, rew, done = client_a.step_many(a_dev)
    for t in range(T):
        client_b.step_fused(a_dev[t])
        assert torch.equal(obs[t], client_b.obs) and torch.equal(rew[t], client_b.reward), t
        assert torch.equal(done[t], client_b.done_u8), t
    assert torch.equal(client_a.goal, client_b.goal) and torch.equal(client_a.step_flags, client_b.step_flags)
    assert client_a.counter == client_b.counter == 1 + T
    sa, sb = client_a.stats(), client_b.stats()
    for k in sa:
        assert sa[k] == sb[k] if k != "sum_reward" else abs(sa[k] - sb[k]) <= 1e-6 * abs(sb[k]), k
    assert sa["episodes"] > 0 and sa["holds"] > 0
    # and a second window continues the same trajectory
    obs2, _, _ = client_a.step_many(a_dev[:3])
    for t in range(3):
        client_b.step_fused(a_dev[t])
        assert torch.equal(obs2[t], client_b.obs)


def test_maximum_size_shard_matches_the_oracle_at_both_ends():
    """134,217,728 envs on one GPU (the largest total of SURVEY.md 8d config C4; ~15 GB of HBM, observation offsets
    beyond 4 GiB): because draws depend only on the global env id, the first and the last 4,096 envs must equal a
    4,096-env oracle shard with the matching env_id_base, bit for bit."""
    from gym_roboy_b200.envs import RoboyEnv
    from gym_roboy_b200.envs.simulations import CudaSimulationClient
    n, k, seed = 1 << 27, 4096, 77
    free, _ = torch.cuda.mem_get_info()
    if free < 24 << 30:
        pytest.skip("needs 24 GB of free HBM")
    client = CudaSimulationClient(num_envs=n, seed=seed, device="cuda:0")
    env = RoboyEnv(client)
    obs0 = env.reset()
    ends = {0: orc.OracleEnv(k, seed=seed, env_id_base=0), n - k: orc.OracleEnv(k, seed=seed, env_id_base=n - k)}
    for base, ora in ends.items():
        assert np.array_equal(ora.reset(), obs0[base:base + k].cpu().numpy())
        ora.step_flags[:] = (ora.step_flags & ~np.uint32(orc.STEP_MASK)) | np.uint32(399)
    client.set_step_num(torch.full((k,), 399, dtype=torch.int32), idx=torch.arange(0, k))
    client.set_step_num(torch.full((k,), 399, dtype=torch.int32), idx=torch.arange(n - k, n))
    gen = torch.Generator(device="cuda:0"); gen.manual_seed(5)
    a = torch.empty((n, 8), device="cuda:0")
    for t in range(3):
        a.uniform_(-1, 1, generator=gen)
        a[n - 7:] = 0.0                                                   # hold branch at the very end of the shard
        obs, rew, done, _ = env.step(a)
        for base, ora in ends.items():
            o, r, d = ora.step(a[base:base + k].cpu().numpy())
            assert np.array_equal(obs[base:base + k].cpu().numpy(), o), (t, base)
            assert np.array_equal(done[base:base + k].cpu().numpy(), d), (t, base)
            assert np.allclose(rew[base:base + k].cpu().numpy(), r, rtol=1e-6, atol=0), (t, base)
    s = client.stats()
    assert s["steps"] == 3 * n and s["holds"] >= 3 * 7 and s["violations"] == 0
    assert client.errors() == (0, None)
    client.close()
