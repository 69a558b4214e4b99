"""The four SimulationClient calls of CudaSimulationClient (un-fused kernels) against the oracle's
draws, and the reference-style single-env protocol driven through them."""
import numpy as np
import pytest
import torch

from oracle import oracle as orc

pytestmark = pytest.mark.gpu


def test_plugin_calls_match_the_stub_semantics():
    from gym_roboy_b200.envs.simulations import CudaSimulationClient
    n, seed = 513, 31
    c = CudaSimulationClient(num_envs=n, seed=seed, env_id_base=7, device="cuda:0")
    gids = np.arange(7, 7 + n)
    # __init__: random held state (simulation_client.py:31)
    s = c.read_state()
    q0, qd0 = orc.draw_state(seed, gids, 0)
    assert np.array_equal(s.joint_angles.cpu().numpy(), q0)
    assert np.array_equal(s.joint_vels.cpu().numpy(), qd0)
    assert s.is_feasible.all()
    # zero action -> the held state, unchanged and not advanced (:38-39)
    h = c.forward_step_command(torch.zeros((n, 8)))
    assert torch.equal(h.joint_angles, s.joint_angles) and torch.equal(h.joint_vels, s.joint_vels)
    # non-zero action -> fresh sample that is NOT stored (:40)
    a = torch.full((n, 8), 0.1)
    a[::2] = 0.0
    a[::2, 3] = 2e-8   # just above numpy's atol
    a[1::4] = 0.0      # these hold
    r = c.forward_step_command(a)
    t = c.counter
    fresh_q, _ = orc.draw_state(seed, gids, t)
    hold = np.zeros(n, bool); hold[1::4] = True
    got = r.joint_angles.cpu().numpy()
    assert np.array_equal(got[~hold], fresh_q[~hold]) and np.array_equal(got[hold], s.joint_angles.cpu().numpy()[hold])
    again = c.read_state()
    assert torch.equal(again.joint_angles, s.joint_angles)
    # reset -> zero state (:42-44); goals differ call to call and stay inside the angle space (:46-47)
    z = c.forward_reset_command()
    assert not z.joint_angles.any() and not z.joint_vels.any()
    g1, g2 = c.get_new_goal_joint_angles(), c.get_new_goal_joint_angles()
    assert not torch.equal(g1, g2)
    assert (g1.abs() <= float(orc.PI32)).all() and (g2.abs() <= float(orc.PI32)).all()
    assert np.array_equal(g1.cpu().numpy(), orc.draw_goal(seed, gids, c.counter, sub=0))
    assert np.array_equal(g2.cpu().numpy(), orc.draw_goal(seed, gids, c.counter, sub=1))


def test_single_env_view_returns_reference_types():
    from gym_roboy_b200.envs.robots import RobotState
    from gym_roboy_b200.envs.simulations import CudaSimulationClient
    c = CudaSimulationClient(num_envs=1, seed=5, device="cuda:0")
    s = c.forward_step_command([0.1] * 8)
    assert isinstance(s, RobotState) and isinstance(s.joint_angles, np.ndarray) and s.joint_angles.shape == (3,)
    assert isinstance(s.is_feasible, bool)
    g = c.get_new_goal_joint_angles()
    assert isinstance(g, np.ndarray) and g.shape == (3,) and g.dtype == np.float32
    with pytest.raises(AssertionError):
        c.forward_step_command([0.0] * 7)            # simulation_client.py:37


def test_external_simulator_feed_matches_oracle():
    """roboy_step_external / roboy_reset_external vs the oracle (itself pinned against the reference
    in tests/test_oracle_vs_reference.py::test_external_simulator_feed_matches_reference)."""
    from gym_roboy_b200.envs import RoboyEnv
    from gym_roboy_b200.envs.simulations import CudaSimulationClient
    n, T = 3000, 40
    rng = np.random.default_rng(4)
    for penalty in (False, True):
        client = CudaSimulationClient(num_envs=n, seed=9, device="cuda:0")
        env = RoboyEnv(client, joint_vel_penalty=penalty, auto_reset=False, strict=False)
        o = orc.OracleEnv(n, seed=9, joint_vel_penalty=penalty, auto_reset=False)
        q0 = rng.uniform(-3, 3, (n, 3)).astype(np.float32)
        qd0 = rng.uniform(-0.5, 0.5, (n, 3)).astype(np.float32)
        assert np.array_equal(env.reset_from_states(q0, qd0).cpu().numpy(), o.reset_external(q0, qd0))
        steps = rng.integers(370, 400, n).astype(np.int32)
        client.set_step_num(steps)
        o.step_flags[:] = (o.step_flags & ~np.uint32(orc.STEP_MASK)) | steps.astype(np.uint32)
        reached = 0
        for t in range(T):
            q = rng.uniform(-3, 3, (n, 3)).astype(np.float32)
            qd = rng.uniform(-0.5, 0.5, (n, 3)).astype(np.float32)
            near = rng.random(n) < 0.3
            q[near] = np.clip(o.goal.T[near] + rng.uniform(-0.04, 0.04, (near.sum(), 3)), -3.1, 3.1).astype(np.float32)
            qd[near] = rng.uniform(-0.25, 0.25, (near.sum(), 3)).astype(np.float32)
            feas = (rng.random(n) > 0.2).astype(np.uint8)
            obs, rew, done, _ = env.step_from_states(q, qd, feas)
            oo, orw, od = o.step_external(q, qd, feas)
            assert np.array_equal(done.cpu().numpy(), od) and np.array_equal(obs.cpu().numpy(), oo)
            rel = np.abs(rew.cpu().numpy() - orw) / np.abs(orw)
            assert rel.max() <= 1e-6
            reached += int((od & (orw > 500)).sum())
            if od.any():
                ro = env.reset_from_states(q, qd, mask=od.astype(np.uint8)).cpu().numpy()
                assert np.array_equal(ro[od], o.reset_external(q, qd, od)[od])
            assert np.array_equal(client.goal.cpu().numpy(), o.goal)
            assert np.array_equal(client.step_num.cpu().numpy(), o.step_num)
        assert reached > 100
        s, so = client.stats(), o.stats()
        for k in ("steps", "episodes", "successes", "timeouts", "violations"):
            assert s[k] == so[k], k


def test_unfused_plugin_protocol_reproduces_the_fused_env():
    """INTEGRATION.md level 1: the four SimulationClient calls, issued in the order the reference's
    RoboyEnv issues them (step: forward_step_command, then on done _set_new_goal ->
    get_new_goal_joint_angles; reset: forward_reset_command, read_state, get_new_goal_joint_angles),
    walk exactly the trajectory the fused step kernel walks from the same seed."""
    from gym_roboy_b200.envs import RoboyEnv
    from gym_roboy_b200.envs.simulations import CudaSimulationClient
    n, T, seed = 2000, 30, 77
    fused_client = CudaSimulationClient(num_envs=n, seed=seed, device="cuda:0")
    fused = RoboyEnv(fused_client, auto_reset=False, strict=False)
    plug = CudaSimulationClient(num_envs=n, seed=seed, device="cuda:0")     # driven call by call
    act_space = plug.robot.get_action_space()
    lo, hi = np.float32(act_space.low[0]), np.float32(act_space.high[0])
    slope = np.float32((hi - lo) / np.float32(2.0))

    def plug_reset(mask=None):
        state = plug.forward_reset_command(mask)                              # roboy_env.py:83-84
        goal = plug.get_new_goal_joint_angles()                               # :86 -> _set_new_goal
        idx = None if mask is None else torch.nonzero(torch.as_tensor(mask)).flatten()
        plug.set_goal(goal if idx is None else goal[idx.to(goal.device)], idx=idx)
        return state

    assert torch.equal(fused_client.goal, plug.goal) and torch.equal(fused_client.held, plug.held)
    fused.reset(); plug_reset()
    assert torch.equal(fused_client.goal, plug.goal)
    fused_client.set_step_num(np.full(n, 395, np.int32))
    rng = np.random.default_rng(0)
    for t in range(T):
        a = rng.uniform(-1, 1, (n, 8)).astype(np.float32)
        a[rng.random(n) < 0.1] = 0.0
        obs, _, done, _ = fused.step(torch.as_tensor(a, device="cuda:0"))
        rescaled = slope * (a - np.float32(1.0)) + hi                          # roboy_env.py:54-57,157-158
        state = plug.forward_step_command(torch.as_tensor(rescaled))          # :59
        assert torch.equal(state.joint_angles, obs[:, 0:3]) and torch.equal(state.joint_vels, obs[:, 3:6])
        assert torch.equal(obs[:, 6:9], plug.goal.t())                        # obs carries the goal in force
        d = done.cpu().numpy()
        if d.any():
            idx = torch.nonzero(done).flatten()
            plug.set_goal(plug.get_new_goal_joint_angles()[idx], idx=idx)     # :67-68 _set_new_goal()
            assert torch.equal(fused_client.goal, plug.goal)
            fused.reset(mask=done.to(torch.uint8))                            # the caller's reset()
            plug_reset(d.astype(np.uint8))
            assert torch.equal(fused_client.goal, plug.goal)
    assert fused_client.counter == plug.counter
    assert fused_client.stats()["holds"] == plug.stats()["holds"] > 0


@pytest.mark.parametrize("robot", ["msj", "five_joints_11_tendons_per_component"])
def test_external_state_outside_the_angle_space_sets_its_error_bit(robot):
    """The reference's client builds every state it receives with robot.new_state() (ros_simulation_client.py:40-46),
    whose assert (roboy_robot.py:76) rejects joint angles outside the angle space and NaN; velocities are not checked
    (:77).  roboy_step_external / roboy_reset_external report that as ROBOY_ERR_STATE_BOUNDS + the first offending env and
    carry on with the batch -- same word, same env, same violation count as the oracle (pinned against the reference's
    new_state in tests/test_oracle_vs_reference.py)."""
    from cuda_adaptor import robot_from_bounds
    from gym_roboy_b200 import _native
    from gym_roboy_b200.envs import RoboyEnv
    from gym_roboy_b200.envs.simulations import CudaSimulationClient
    from test_oracle_vs_reference import GENERIC_ROBOTS
    b = {} if robot == "msj" else GENERIC_ROBOTS[robot]
    J, _, _, bb = orc.robot_bounds(b)
    n = 5000
    rng = np.random.default_rng(12)
    lo, hi = bb["angle_low"], bb["angle_high"]

    def states(first_bad):
        q = rng.uniform(lo, hi, (n, J)).astype(np.float32)
        qd = (rng.uniform(bb["vel_low"], bb["vel_high"], (n, J)) * 3).astype(np.float32)     # velocities may leave their space
        q[first_bad + 3, 0] = np.nextafter(hi[0], np.float32(10))
        q[first_bad, J - 1] = np.nextafter(lo[J - 1], np.float32(-10))
        q[first_bad + 40, J // 2] = np.nan
        q[first_bad + 41, 0] = np.inf
        q[first_bad + 50] = hi                                                               # on the bounds: inside
        q[first_bad + 51] = lo
        return q, qd

    client = CudaSimulationClient(robot=robot_from_bounds(b) if b else None, num_envs=n, seed=3, env_id_base=1000, device="cuda:0")
    env = RoboyEnv(client, auto_reset=False, strict=False)
    o = orc.OracleEnv(n, seed=3, env_id_base=1000, auto_reset=False, **b)
    q, qd = states(700)
    obs = env.reset_from_states(q, qd).cpu().numpy()
    assert np.array_equal(obs.view(np.uint32), o.reset_external(q, qd).view(np.uint32))
    assert client.errors() == (_native.ERR_STATE_BOUNDS, 1700) and o.errors() == (orc.ERR_STATE_BOUNDS, 1700)
    client.clear_errors()
    q, qd = states(200)
    obs, rew, done, _ = env.step_from_states(q, qd)
    oo, orw, od = o.step_external(q, qd)
    assert np.array_equal(obs.cpu().numpy().view(np.uint32), oo.view(np.uint32)) and np.array_equal(done.cpu().numpy(), od)
    # (the env handed an infinite angle also gets an infinite reward: both bits, in the kernel and in the oracle)
    both = _native.ERR_STATE_BOUNDS | _native.ERR_REWARD_RANGE
    assert client.errors() == (both, 1200) and o.errors() == (both, 1200), (client.errors(), o.errors())
    assert client.stats()["violations"] == o.stats()["violations"] == 4
    with pytest.raises(AssertionError, match="roboy_robot.py:76"):
        env.check_errors()
