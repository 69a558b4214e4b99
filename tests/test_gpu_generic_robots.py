"""robots with other joint / tendon counts and one bound per component (roboy_robot.py:21-33,
README.md:6-7 "MSJ platform, Upper Body, etc.") on the generic kernels -- CUDA against the CPU oracle, which
tests/test_oracle_vs_reference.py pins against the unmodified reference for the same robots."""
import contextlib
import os

import numpy as np
import pytest
import torch

from cuda_adaptor import robot_from_bounds
from oracle import oracle as orc
from test_oracle_vs_reference import GENERIC_ROBOTS

pytestmark = pytest.mark.gpu
RTOL = 1e-6


def make_pair(b, n, seed, penalty=False, auto_reset=True, env_id_base=0):
    from gym_roboy_b200.envs import RoboyEnv
    from gym_roboy_b200.envs.simulations import CudaSimulationClient
    client = CudaSimulationClient(robot=robot_from_bounds(b), num_envs=n, seed=seed, env_id_base=env_id_base, device="cuda:0")
    env = RoboyEnv(client, joint_vel_penalty=penalty, auto_reset=auto_reset, strict=False)
    env._single = False
    ora = orc.OracleEnv(n, seed=seed, env_id_base=env_id_base, joint_vel_penalty=penalty, auto_reset=auto_reset, threads=8, **b)
    return env, client, ora


def set_phases(client, ora, steps):
    client.set_step_num(steps)
    ora.step_flags[:] = (ora.step_flags & ~np.uint32(orc.STEP_MASK)) | steps.astype(np.uint32)


def compare_step(env, client, ora, a, t):
    obs, rew, done, _ = env.step(torch.as_tensor(a, device="cuda:0"))
    o_obs, o_rew, o_done = ora.step(a)
    assert np.array_equal(done.cpu().numpy(), o_done), "done mask differs at step %d" % t
    assert np.array_equal(obs.cpu().numpy(), o_obs), "obs differ at step %d" % t
    rel = np.abs(rew.cpu().numpy().astype(np.float64) - o_rew) / np.maximum(np.abs(o_rew.astype(np.float64)), 1e-30)
    assert rel.max() <= RTOL, "reward rel err %g at step %d" % (rel.max(), t)
    return o_done


@pytest.mark.parametrize("penalty", [False, True])
@pytest.mark.parametrize("name", sorted(GENERIC_ROBOTS))
def test_generic_robot_rollout_matches_oracle(name, penalty):
    b = GENERIC_ROBOTS[name]
    J, A, _, bb = orc.robot_bounds(b)
    n, T, seed = 4096, 430, 6
    env, client, ora = make_pair(b, n, seed, penalty)
    assert not client.msj_kernels and (client.dim_joint, client.dim_action, client.dim_obs) == (J, A, 3 * J)
    assert client.obs.shape == (n, 3 * J) and client.goal.shape == (J, n) and client.held.shape == (2 * J, n)
    assert np.allclose(env.reward_range, ora.reward_range, rtol=RTOL, atol=0)
    assert np.array_equal(client.goal.cpu().numpy(), ora.goal) and np.array_equal(client.held.cpu().numpy(), ora.held)
    assert np.array_equal(env.reset().cpu().numpy(), ora.reset())
    rng = np.random.default_rng(2)
    zero_action, can_hold = orc.hold_action(b)
    set_phases(client, ora, rng.integers(1, 400, n).astype(np.int32))
    client.enable_done_index(True)
    thr_a, _ = orc.thresholds(ora.cfg)
    for t in range(T):
        a = rng.uniform(-1, 1, (n, A)).astype(np.float32)
        hold = rng.random(n) < 0.03
        a[hold] = zero_action
        a[rng.random(n) < 0.01] = np.nextafter(zero_action, np.float32(1))   # right next to the hold interval
        if t % 40 == 5:     # goals around the next state (sampled rows) / the held zero state (hold rows), both sides of the threshold
            q, _ = orc.draw_state(seed, np.arange(n), ora.counter + 1, bb["angle_low"], bb["angle_high"], J=J)
            d = rng.normal(size=(n, J)); d /= np.linalg.norm(d, axis=1, keepdims=True)
            r = float(thr_a) * (1 + rng.choice([-1e-7, 1e-7, -0.3, 0.2], n))
            base = np.where(hold[:, None], 0.0, q.astype(np.float64))
            g = np.clip(base + d * r[:, None], bb["angle_low"], bb["angle_high"]).astype(np.float32)
            client.set_goal(g); ora.goal[:] = g.T
        if t % 60 == 20:    # injected float32 held states, some infeasible
            q = rng.uniform(bb["angle_low"], bb["angle_high"], (n, J)).astype(np.float32)
            qd = (rng.uniform(bb["vel_low"], bb["vel_high"], (n, J)) * 0.2).astype(np.float32)
            feas = rng.random(n) < 0.6
            client.set_state(q, qd, feas.astype(np.uint8))
            ora.held[0:J] = q.T; ora.held[J:2 * J] = qd.T
            ora.step_flags[:] = (ora.step_flags & np.uint32(orc.STEP_MASK)) | np.where(feas, 0, orc.F_HELD_INFEASIBLE).astype(np.uint32)
        od = compare_step(env, client, ora, a, t)
        if t % 50 == 0:
            assert np.array_equal(client.done_indices()[0].cpu().numpy(), np.flatnonzero(od).astype(np.int32))
    assert np.array_equal(client.goal.cpu().numpy(), ora.goal)
    assert np.array_equal(client.step_flags.cpu().numpy().astype(np.uint32), ora.step_flags)
    s, so = client.stats(), ora.stats()
    for k in ("steps", "episodes", "successes", "timeouts", "sum_episode_len", "holds", "violations"):
        assert s[k] == so[k], (k, s[k], so[k])
    assert abs(s["sum_reward"] - so["sum_reward"]) <= 1e-6 * abs(so["sum_reward"])
    assert s["timeouts"] > n * 0.9 and (s["holds"] > 0) == can_hold
    assert client.errors()[0] == ora.errors()[0]
    if client.errors()[0]:
        assert client.errors()[1] == ora.errors()[1]


def test_which_robots_run_the_tuned_kernels():
    from gym_roboy_b200.envs.simulations import CudaSimulationClient
    msj_like = dict(angle_low=-2.0, angle_high=2.0, vel_low=-0.7, vel_high=0.7, act_low=-0.5, act_high=0.5)
    assert CudaSimulationClient(num_envs=64, seed=1, device="cuda:0").msj_kernels
    assert CudaSimulationClient(robot=robot_from_bounds(msj_like), num_envs=64, seed=1, device="cuda:0").msj_kernels
    for name in GENERIC_ROBOTS:
        assert not CudaSimulationClient(robot=robot_from_bounds(GENERIC_ROBOTS[name]), num_envs=64, seed=1, device="cuda:0").msj_kernels
    with pytest.raises(ValueError, match="joints"):
        CudaSimulationClient(robot=robot_from_bounds(dict(dim_joint=16, dim_action=8)), num_envs=4, device="cuda:0")


@contextlib.contextmanager
def environ(**kv):
    """os.environ entries for the duration of the block (the library reads its switches when a handle is created and when
    RoboyEnv hands it the reward range)"""
    old = {k: os.environ.get(k) for k in kv}
    os.environ.update(kv)
    try:
        yield
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


# The velocity penalty of a sampled state is evaluated in float32 away from the bounds of reward_range (PenaltyF32,
# msj_math.cuh; tests/test_gpu_penalty_f32.py) while the general path -- which the generic step also takes for a division
# numerator outside the proved range -- keeps the float64 expression: rewards of two handles that route an env differently
# then agree to ~2e-7, not bit for bit.  The bit-equality tests below therefore pin both handles to the float64
# expression when the penalty is on ("float64"); "float32" runs the product default and compares rewards to PENALTY_REL.
PENALTY_MODES = [(False, "float32"), (True, "float64"), (True, "float32")]
PENALTY_REL = 5e-7


def rewards_agree(r1, r2, penalty, mode):
    if not penalty or mode == "float64":
        return torch.equal(r1.view(torch.int32), r2.view(torch.int32))   # bit pattern: NaN-safe
    a, b = r1.double(), r2.double()
    same = r1.view(torch.int32) == r2.view(torch.int32)
    return bool((same | ((a - b).abs() <= PENALTY_REL * b.abs())).all().item())


@pytest.mark.parametrize("penalty,mode", PENALTY_MODES)
def test_msj_through_the_generic_kernels_equals_the_tuned_kernels(penalty, mode):
    """The tuned MSJ kernels are an instantiation of the generic path: same seed, same actions -> the same bits
    (rewards included: the 3-instruction division is proved identical to IEEE division)."""
    from gym_roboy_b200.envs import RoboyEnv
    from gym_roboy_b200.envs.simulations import CudaSimulationClient
    n, T = 40_001, 30
    with environ(ROBOY_B200_PENALTY_F64="1" if mode == "float64" else "0"):
        tuned_c = CudaSimulationClient(num_envs=n, seed=3, device="cuda:0")
        with environ(ROBOY_B200_FORCE_GENERIC="1"):
            gen_c = CudaSimulationClient(num_envs=n, seed=3, device="cuda:0")
        assert tuned_c.msj_kernels and not gen_c.msj_kernels
        tuned, gen = RoboyEnv(tuned_c, joint_vel_penalty=penalty, strict=False), RoboyEnv(gen_c, joint_vel_penalty=penalty, strict=False)
    assert tuned_c.penalty_float32 == gen_c.penalty_float32 == (mode == "float32")
    assert torch.equal(tuned.reset(), gen.reset())
    steps = (np.arange(n) % 400 + 1).astype(np.int32)
    tuned_c.set_step_num(steps); gen_c.set_step_num(steps)
    rng = np.random.default_rng(0)
    for t in range(T):
        a = rng.uniform(-1, 1, (n, 8)).astype(np.float32)
        a[rng.random(n) < 0.05] = 0.0
        a_dev = torch.as_tensor(a, device="cuda:0")
        o1, r1, d1, _ = tuned.step(a_dev)
        o2, r2, d2, _ = gen.step(a_dev)
        assert torch.equal(o1, o2) and torch.equal(d1, d2) and rewards_agree(r1, r2, penalty, mode), t
    assert torch.equal(tuned_c.goal, gen_c.goal) and torch.equal(tuned_c.step_flags, gen_c.step_flags)
    s1, s2 = tuned_c.stats(), gen_c.stats()
    for k in s1:
        assert s1[k] == s2[k] if k != "sum_reward" else abs(s1[k] - s2[k]) <= 1e-9 * abs(s2[k]), k


def test_generic_robot_host_path_step_many_unfused_calls_and_external_feed():
    b = GENERIC_ROBOTS["five_joints_11_tendons_per_component"]
    J, A, _, bb = orc.robot_bounds(b)
    n, seed = 70_003, 11
    env, client, ora = make_pair(b, n, seed)
    env2, client2, _ = make_pair(b, n, seed)
    env.reset(); env2.reset(); ora.reset()
    steps = (np.arange(n) % 400 + 1).astype(np.int32)
    set_phases(client, ora, steps); client2.set_step_num(steps)
    rng = np.random.default_rng(4)
    zero_action, _ = orc.hold_action(b)
    # ---- host buffers (several stages, ragged) ----
    client.set_host_pipeline(stage_envs=1 << 14, n_streams=2)
    a_h, obs_h, rew_h, done_h = client.host_buffers()
    for t in range(2):
        a = rng.uniform(-1, 1, (n, A)).astype(np.float32); a[rng.random(n) < 0.05] = zero_action
        a_h[...] = a
        client.step_host(a_h, obs_h, rew_h, done_h)
        o, r, d = ora.step(a)
        assert np.array_equal(obs_h, o) and np.array_equal(done_h.astype(bool), d) and np.allclose(rew_h, r, rtol=RTOL, atol=0)
        client2.step_fused(torch.as_tensor(a, device="cuda:0"))
    # ---- T steps in one call (generic robots: T launches) ----
    acts = rng.uniform(-1, 1, (3, n, A)).astype(np.float32)
    obs, rew, done = client.step_many(torch.as_tensor(acts, device="cuda:0"))
    for t in range(3):
        o, r, d = ora.step(acts[t])
        assert np.array_equal(obs[t].cpu().numpy(), o) and np.array_equal(done[t].cpu().numpy().astype(bool), d)
    assert client.counter == ora.counter == 1 + 2 + 3
    # ---- external feed ----
    q = rng.uniform(bb["angle_low"], bb["angle_high"], (n, J)).astype(np.float32)
    qd = (rng.uniform(bb["vel_low"], bb["vel_high"], (n, J)) * 0.3).astype(np.float32)
    feas = (rng.random(n) < 0.8).astype(np.uint8)
    near = np.clip(q + np.float32(0.002), bb["angle_low"], bb["angle_high"]).astype(np.float32)
    client.set_goal(near[::3], idx=np.arange(0, n, 3))
    ora.goal[:, ::3] = near[::3].T
    client.step_external(q, qd, feas)
    o, r, d = ora.step_external(q, qd, feas)
    assert np.array_equal(client.obs.cpu().numpy(), o) and np.array_equal(client.done.cpu().numpy(), d)
    assert np.allclose(client.reward.cpu().numpy(), r, rtol=RTOL, atol=0) and d.sum() > n // 10
    assert np.array_equal(client.goal.cpu().numpy(), ora.goal)
    client.reset_external(q, qd, mask=d.astype(np.uint8))
    assert np.array_equal(client.obs.cpu().numpy()[d], ora.reset_external(q, qd, d.astype(np.uint8))[d])


def test_generic_robot_unfused_plugin_calls_reproduce_the_fused_env():
    b = GENERIC_ROBOTS["six_joints_14_tendons"]
    J, A, _, bb = orc.robot_bounds(b)
    n, T, seed = 3000, 25, 77
    fused_env, fused_client, _ = make_pair(b, n, seed, auto_reset=False)
    from gym_roboy_b200.envs.simulations import CudaSimulationClient
    plug = CudaSimulationClient(robot=fused_client.robot, num_envs=n, seed=seed, device="cuda:0")
    hi, lo = bb["act_high"], bb["act_low"]
    slope = ((hi - lo) / np.float32(2.0)).astype(np.float32)

    def plug_reset(mask=None):
        state = plug.forward_reset_command(mask)
        goal = plug.get_new_goal_joint_angles()
        idx = None if mask is None else torch.nonzero(torch.as_tensor(mask)).flatten()
        plug.set_goal(goal if idx is None else goal[idx.to(goal.device)], idx=idx)
        return state

    assert torch.equal(fused_client.goal, plug.goal) and torch.equal(fused_client.held, plug.held)
    assert torch.equal(plug.get_new_goal_joint_angles().t(), fused_client.goal)      # first goal asked = construction goal
    fused_env.reset(); plug_reset()
    assert torch.equal(fused_client.goal, plug.goal)
    fused_client.set_step_num(np.full(n, 395, np.int32))
    rng = np.random.default_rng(0)
    zero_action, _ = orc.hold_action(b)
    for t in range(T):
        a = rng.uniform(-1, 1, (n, A)).astype(np.float32)
        a[rng.random(n) < 0.1] = zero_action
        obs, _, done, _ = fused_env.step(torch.as_tensor(a, device="cuda:0"))
        rescaled = (slope * (a - np.float32(1.0))).astype(np.float32) + hi            # roboy_env.py:54-57,157-158
        state = plug.forward_step_command(torch.as_tensor(rescaled))
        assert torch.equal(state.joint_angles, obs[:, 0:J]) and torch.equal(state.joint_vels, obs[:, J:2 * J])
        assert torch.equal(obs[:, 2 * J:], plug.goal.t())
        d = done.cpu().numpy()
        if d.any():
            idx = torch.nonzero(done).flatten()
            plug.set_goal(plug.get_new_goal_joint_angles()[idx], idx=idx)
            assert torch.equal(fused_client.goal, plug.goal)
            fused_env.reset(mask=done.to(torch.uint8)); plug_reset(d.astype(np.uint8))
            assert torch.equal(fused_client.goal, plug.goal)
    assert fused_client.counter == plug.counter and fused_client.stats()["holds"] == plug.stats()["holds"] > 0


def test_vec_env_adapter_with_another_robot():
    from gym_roboy_b200.vec_env import RoboyVecEnv
    b = GENERIC_ROBOTS["six_joints_14_tendons"]
    n = 500
    venv = RoboyVecEnv(n, seed=4, robot=robot_from_bounds(b))
    ora = orc.OracleEnv(n, seed=4, **b)
    obs = venv.reset()
    assert obs.shape == (n, 18) and np.array_equal(obs, ora.reset())
    assert venv.observation_space.shape == (18,) and venv.action_space.shape == (14,)
    venv.client.set_step_num(np.full(n, 399, np.int32))
    ora.step_flags[:] = (ora.step_flags & ~np.uint32(orc.STEP_MASK)) | np.uint32(399)
    rng = np.random.default_rng(0)
    for t in range(3):
        a = rng.uniform(-1, 1, (n, 14)).astype(np.float32)
        obs, rew, done, infos = venv.step(a)
        o, r, d, term = ora.step(a, want_terminal_obs=True)
        assert np.array_equal(obs, o) and np.array_equal(done, d) and np.allclose(rew, r, rtol=1e-6, atol=0)
        for i in np.flatnonzero(d):
            assert np.array_equal(infos[i]["terminal_observation"], term[i])
    venv.close()


@pytest.mark.parametrize("J", list(range(1, 16)))
def test_every_joint_count_instantiation_matches_oracle(J):
    """generic_step_kernel<J> for J = 1..15, each on a random robot (per-component asymmetric bounds, random tendon
    count), both penalty modes alternating, ragged size; the same robots are pinned against the reference on the CPU
    (tests/test_oracle_vs_reference.py::test_every_joint_count_matches_reference)."""
    from test_oracle_vs_reference import random_robot
    rng = np.random.default_rng(1000 + J)
    b = random_robot(rng, J)
    _, A, _, bb = orc.robot_bounds(b)
    n, T = 3001, 70
    env, client, ora = make_pair(b, n, seed=J, penalty=bool(J % 2))
    assert not client.msj_kernels or (J == 3 and A == 8)
    assert np.array_equal(env.reset().cpu().numpy(), ora.reset())
    set_phases(client, ora, rng.integers(340, 401, n).astype(np.int32))
    zero_action, can_hold = orc.hold_action(b)
    thr_a, _ = orc.thresholds(ora.cfg)
    client.enable_terminal_obs(True)
    for t in range(T):
        a = rng.uniform(-1, 1, (n, A)).astype(np.float32)
        hold = rng.random(n) < 0.05
        a[hold] = zero_action
        a[rng.random(n) < 0.002] = np.float32(1.5)          # outside the action space: error word, first offending env
        if t % 7 == 3:
            a[rng.integers(0, n), rng.integers(0, A)] = np.nan   # one NaN component somewhere: fails the same assert
        if t % 10 == 4:
            q, _ = orc.draw_state(J, np.arange(n), ora.counter + 1, bb["angle_low"], bb["angle_high"], J=J)
            d = rng.normal(size=(n, J)); d /= np.linalg.norm(d, axis=1, keepdims=True)
            base = np.where(hold[:, None], 0.0, q.astype(np.float64))
            g = np.clip(base + d * float(thr_a) * (1 + rng.choice([-1e-7, 1e-7, -0.4, 0.3], n))[:, None],
                        bb["angle_low"], bb["angle_high"]).astype(np.float32)
            client.set_goal(g); ora.goal[:] = g.T
        obs, rew, done, info = env.step(torch.as_tensor(a, device="cuda:0"))
        o_obs, o_rew, o_done, o_term = ora.step(a, want_terminal_obs=True)
        assert np.array_equal(done.cpu().numpy(), o_done) and np.array_equal(obs.cpu().numpy(), o_obs), t
        rel = np.abs(rew.cpu().numpy().astype(np.float64) - o_rew) / np.maximum(np.abs(o_rew.astype(np.float64)), 1e-30)
        assert rel.max() <= RTOL, (t, rel.max())
        assert np.array_equal(client.terminal_obs.cpu().numpy()[o_done], o_term[o_done]), t
    assert np.array_equal(client.goal.cpu().numpy(), ora.goal)
    assert np.array_equal(client.step_flags.cpu().numpy().astype(np.uint32), ora.step_flags)
    s, so = client.stats(), ora.stats()
    for k in ("steps", "episodes", "successes", "timeouts", "sum_episode_len", "holds", "violations"):
        assert s[k] == so[k], (k, s[k], so[k])
    assert client.errors() == ora.errors() and client.errors()[0] & 1 and s["episodes"] > n // 2


@pytest.mark.parametrize("penalty,mode", PENALTY_MODES)
@pytest.mark.parametrize("name", sorted(GENERIC_ROBOTS))
def test_proved_fast_division_equals_ieee_division(name, penalty, mode):
    """The generic step divides by the robot's spans with a three-instruction core after checking it, at construction,
    against IEEE division over all 2^32 numerators.  Same seed, same actions, same goals (some far outside the proved
    numerator range, some making a numerator exactly zero) -> the same bits as a handle kept on IEEE division."""
    from gym_roboy_b200.envs import RoboyEnv
    from gym_roboy_b200.envs.simulations import CudaSimulationClient
    b = GENERIC_ROBOTS[name]
    J, A, _, bb = orc.robot_bounds(b)
    n, T = 20_011, 24
    with environ(ROBOY_B200_PENALTY_F64="1" if mode == "float64" else "0"):
        fast_c = CudaSimulationClient(robot=robot_from_bounds(b), num_envs=n, seed=9, device="cuda:0")
        with environ(ROBOY_B200_GENERIC_FASTDIV="0"):
            ieee_c = CudaSimulationClient(robot=robot_from_bounds(b), num_envs=n, seed=9, device="cuda:0")
        assert fast_c.fast_division and not ieee_c.fast_division
        fast, ieee = [RoboyEnv(c, joint_vel_penalty=penalty, strict=False) for c in (fast_c, ieee_c)]
    assert fast_c.penalty_float32 == ieee_c.penalty_float32 and not (mode == "float64" and fast_c.penalty_float32)
    assert torch.equal(fast.reset(), ieee.reset())
    rng = np.random.default_rng(4)
    lo, hi = np.broadcast_to(bb["angle_low"], (J,)).astype(np.float32), np.broadcast_to(bb["angle_high"], (J,)).astype(np.float32)
    for t in range(T):
        if t % 6 == 2:   # goals written straight into the state rows: the midpoint (numerator 0), denormals, huge values, inf
            g = fast_c.goal.clone()
            mid = torch.as_tensor((lo.astype(np.float64) + hi.astype(np.float64)) / 2, dtype=torch.float32, device="cuda:0")
            g[:, 0::7] = mid[:, None]
            g[:, 1::7] = 1e-42
            g[:, 2::7] = 3e38
            g[:, 3::7] = float("inf")
            g[:, 4::7] = -1e25
            fast_c.goal.copy_(g); ieee_c.goal.copy_(g)
        a = torch.as_tensor(rng.uniform(-1, 1, (n, A)).astype(np.float32), device="cuda:0")
        o1, r1, d1, _ = fast.step(a)
        o2, r2, d2, _ = ieee.step(a)
        assert torch.equal(o1, o2) and torch.equal(d1, d2), t
        assert rewards_agree(r1, r2, penalty, mode), t
    assert torch.equal(fast_c.goal, ieee_c.goal) and torch.equal(fast_c.step_flags, ieee_c.step_flags)


@pytest.mark.parametrize("name", ["six_joints_14_tendons", "fifteen_joints_64_tendons"])
def test_quiet_chunks_and_unaligned_action_slices(name):
    """The whole-chunk action test reads 32 rows as float4 when the slice is 16-byte aligned and falls back to one row per
    lane otherwise.  A batch where most chunks contain no hold and no bad action (the quick path) and a few contain exactly
    one of either, stepped one call at a time (aligned) and two steps per call (`step_many`: the second slice of an odd
    population is only 4-byte aligned) -- both against the oracle."""
    b = GENERIC_ROBOTS[name]
    J, A, _, bb = orc.robot_bounds(b)
    n, T = 50_003, 12
    e1, c1, ora = make_pair(b, n, 11)
    e2, c2, _ = make_pair(b, n, 11)
    first = e1.reset()
    assert torch.equal(first, e2.reset()) and np.array_equal(first.cpu().numpy(), ora.reset())
    zero_action, can_hold = orc.hold_action(b)
    rng = np.random.default_rng(5)
    assert (n * A * 4) % 16 != 0 or A % 4 == 0   # 14 tendons: the second slice is only 8-byte aligned
    for t in range(0, T, 2):
        pair = []
        for u in range(2):
            a = rng.uniform(-1, 1, (n, A)).astype(np.float32)
            for i in rng.integers(0, n, 40):
                a[i] = zero_action                                   # 40 chunks of 1563 hold one env
            for i in rng.integers(0, n, 20):
                a[i, rng.integers(0, A)] = [np.nan, 1.0000001, -3.0][(t + u) % 3]
            a[rng.integers(0, n, 20)] = np.nextafter(zero_action, np.float32(-1))   # next to the hold interval, not inside
            pair.append(a)
        both = torch.as_tensor(np.stack(pair), device="cuda:0")
        obs_m, rew_m, done_m = c2.step_many(both)
        for u in range(2):
            o1, r1, d1, _ = e1.step(both[u].clone())
            o_obs, o_rew, o_done = ora.step(pair[u])
            assert np.array_equal(o1.cpu().numpy(), o_obs) and np.array_equal(d1.cpu().numpy(), o_done), (t, u)
            assert torch.equal(obs_m[u], o1) and torch.equal(done_m[u].bool(), d1.bool()), (t, u)
            assert torch.equal(rew_m[u].view(torch.int32), r1.view(torch.int32)), (t, u)
    assert c1.errors() == c2.errors() == ora.errors() and c1.errors()[0] & 1
    s1, s2, so = c1.stats(), c2.stats(), ora.stats()
    for k in ("steps", "episodes", "successes", "timeouts", "holds", "violations"):
        assert s1[k] == s2[k] == so[k], k
    assert (s1["holds"] > 0) == can_hold


def test_generic_soak_slice():
    """15 seconds of tools/soak_generic.py: random robots (1..15 joints, 1..64 tendons, asymmetric / one-sided / symmetric
    spaces), random flags, goals around the reached threshold and at the midpoint of the angle space, bad actions."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    from soak_generic import soak
    summary = soak(15.0, master_seed=7)
    assert not summary["mismatches"], summary["mismatches"][:3]
    assert summary["configs"] >= 3 and summary["fast_division"] == summary["configs"]
    assert summary["successes"] > 0 and summary["timeouts"] > 0 and summary["violations"] > 0


@pytest.mark.parametrize("robot", ["msj", "six_joints_14_tendons"])
def test_env_sitting_on_its_goal_is_inside_the_reward_range(robot):
    """joint_vel_penalty on, bonus off: max_reward = fl32(-1 - e^-1) is an exact tie in float32 and must not come out below
    the float64 reward of an env that holds the zero state with its goal planted exactly there (distance 0, velocity 0) --
    the reference raises nothing for that env (roboy_env.py:109); found by tools/soak_generic.py."""
    from gym_roboy_b200.envs import RoboyEnv
    from gym_roboy_b200.envs.simulations import CudaSimulationClient
    b = {} if robot == "msj" else GENERIC_ROBOTS[robot]
    J, A, _, bb = orc.robot_bounds(b) if b else (3, 8, None, None)
    n = 4099
    client = CudaSimulationClient(robot=robot_from_bounds(b) if b else None, num_envs=n, seed=2, device="cuda:0")
    env = RoboyEnv(client, joint_vel_penalty=True, is_agent_getting_bonus_for_reaching_goal=False, auto_reset=False, strict=False)
    ora = orc.OracleEnv(n, seed=2, joint_vel_penalty=True, bonus=False, auto_reset=False, threads=4, **b)
    assert env.reward_range == tuple(ora.reward_range), (env.reward_range, ora.reward_range)   # bit-equal: both exps correctly rounded
    env.reset(); ora.reset()
    zero_action, can_hold = orc.hold_action(b) if b else (np.zeros(8, np.float32), True)
    assert can_hold
    g = np.zeros((n, J), np.float32)
    client.set_goal(g); ora.goal[:] = g.T
    a = np.tile(zero_action, (n, 1)).astype(np.float32)
    obs, rew, done, _ = env.step(torch.as_tensor(a, device="cuda:0"))
    o_obs, o_rew, o_done = ora.step(a)
    assert np.array_equal(obs.cpu().numpy(), o_obs) and np.array_equal(done.cpu().numpy(), o_done) and o_done.all()
    assert np.allclose(rew.cpu().numpy(), o_rew, rtol=RTOL, atol=0)
    assert np.all(rew.cpu().numpy() == np.float32(env.reward_range[1]))
    assert client.errors()[0] == ora.errors()[0] == 0
    assert client.stats()["violations"] == ora.stats()["violations"] == 0


def test_large_population_of_a_64_tendon_robot_addresses_beyond_4_gib():
    """17,000,005 envs of the 15-joint / 64-tendon robot: the action array is 4.35 GB and the observations 3.06 GB, so
    every row offset of the second half needs more than 32 bits.  Two steps against the oracle, bit-exact."""
    b = GENERIC_ROBOTS["fifteen_joints_64_tendons"]
    J, A, _, bb = orc.robot_bounds(b)
    n = 17_000_005
    env, client, ora = make_pair(b, n, seed=21)
    assert np.array_equal(env.reset().cpu().numpy(), ora.reset())
    rng = np.random.default_rng(8)
    steps = rng.integers(399, 401, n).astype(np.int32)     # half the population times out on the first step
    set_phases(client, ora, steps)
    zero_action, _ = orc.hold_action(b)
    for t in range(2):
        a = rng.random((n, A), dtype=np.float32) * 2 - 1
        a[::1001] = zero_action
        a[n - 3, A - 1] = 1.5                                  # the last rows matter: out of range at the far end
        obs, rew, done, _ = env.step(torch.as_tensor(a, device="cuda:0"))
        o_obs, o_rew, o_done = ora.step(a)
        assert np.array_equal(done.cpu().numpy(), o_done), t
        assert np.array_equal(obs.cpu().numpy(), o_obs), t
        r = rew.cpu().numpy().astype(np.float64)
        assert (np.abs(r - o_rew) <= RTOL * np.abs(o_rew)).all(), t
        del a, obs, rew, done, o_obs, o_rew, r
    assert np.array_equal(client.goal.cpu().numpy(), ora.goal)
    assert np.array_equal(client.step_flags.cpu().numpy().astype(np.uint32), ora.step_flags)
    assert client.errors() == ora.errors() and client.errors()[1] == n - 3
    s, so = client.stats(), ora.stats()
    for k in ("steps", "episodes", "successes", "timeouts", "holds", "violations"):
        assert s[k] == so[k], k
    assert s["episodes"] > n // 3


MSJ_SHAPED_OTHER_LIMITS = {
    "symmetric": dict(angle_low=-2.0, angle_high=2.0, vel_low=-0.7, vel_high=0.7, act_low=-0.5, act_high=0.5),
    "one_sided": dict(angle_low=0.0, angle_high=2.5, vel_low=-0.3, vel_high=0.9, act_low=0.0, act_high=0.4),
    "asymmetric": dict(angle_low=-1.1, angle_high=2.9, vel_low=-0.45, vel_high=0.2, act_low=-0.3, act_high=0.1),
}


@pytest.mark.parametrize("penalty", [False, True])
@pytest.mark.parametrize("name", sorted(MSJ_SHAPED_OTHER_LIMITS))
def test_msj_shaped_robot_with_other_limits_runs_the_tuned_kernels_with_the_checked_division(name, penalty):
    """3 joints / 8 tendons / uniform bounds that are not MSJ's: the tuned kernels in their range-checked instantiation (the
    division proved on this device at construction, numerators outside the proved range -- exactly zero here -- through
    IEEE division; hold pre-filter around the centre of the hold interval, which a one-sided tendon range puts at -1).
    Against the oracle, and bit for bit (rewards included) against a handle kept on IEEE division; closed loop, open loop
    (`step_many`) and the host-buffer path."""
    from gym_roboy_b200.envs import RoboyEnv
    from gym_roboy_b200.envs.simulations import CudaSimulationClient
    b = MSJ_SHAPED_OTHER_LIMITS[name]
    J, A, _, bb = orc.robot_bounds(b)
    n, T, seed = 20_011, 120, 17
    env, client, ora = make_pair(b, n, seed, penalty)
    os.environ["ROBOY_B200_GENERIC_FASTDIV"] = "0"
    try:
        ieee_c = CudaSimulationClient(robot=robot_from_bounds(b), num_envs=n, seed=seed, device="cuda:0")
    finally:
        del os.environ["ROBOY_B200_GENERIC_FASTDIV"]
    ieee = RoboyEnv(ieee_c, joint_vel_penalty=penalty, auto_reset=True, strict=False)
    ieee._single = False
    assert client.msj_kernels and ieee_c.msj_kernels and client.fast_division and not ieee_c.fast_division
    first = env.reset()
    assert torch.equal(first, ieee.reset()) and np.array_equal(first.cpu().numpy(), ora.reset())
    rng = np.random.default_rng(3)
    steps = rng.integers(1, 400, n).astype(np.int32)
    set_phases(client, ora, steps); ieee_c.set_step_num(steps)
    zero_action, can_hold = orc.hold_action(b)
    thr_a, _ = orc.thresholds(ora.cfg)
    lo, hi = np.float32(bb["angle_low"]), np.float32(bb["angle_high"])
    mid = np.float32((np.float64(lo) + np.float64(hi)) / 2)
    for t in range(T):
        a = rng.uniform(-1, 1, (n, A)).astype(np.float32)
        a[rng.random(n) < 0.03] = zero_action
        a[rng.random(n) < 0.01] = np.nextafter(zero_action, np.float32(1))
        if t % 9 == 4:
            q, _ = orc.draw_state(seed, np.arange(n), ora.counter + 1, bb["angle_low"], bb["angle_high"], J=J)
            d = rng.normal(size=(n, J)); d /= np.linalg.norm(d, axis=1, keepdims=True)
            g = np.clip(q.astype(np.float64) + d * float(thr_a) * (1 + rng.choice([-1e-7, 1e-7, -0.3, 0.2], n))[:, None], lo, hi).astype(np.float32)
            pick = rng.random(n)
            g[pick < 0.1] = mid          # zero numerators: outside the proved range
            g[(pick >= 0.1) & (pick < 0.15)] = lo
            client.set_goal(g); ieee_c.set_goal(g); ora.goal[:] = g.T
        a_dev = torch.as_tensor(a, device="cuda:0")
        o1, r1, d1, _ = env.step(a_dev)
        o2, r2, d2, _ = ieee.step(a_dev)
        assert torch.equal(o1, o2) and torch.equal(d1, d2) and torch.equal(r1.view(torch.int32), r2.view(torch.int32)), t
        o_obs, o_rew, o_done = ora.step(a)
        assert np.array_equal(d1.cpu().numpy(), o_done) and np.array_equal(o1.cpu().numpy(), o_obs), t
        rel = np.abs(r1.cpu().numpy().astype(np.float64) - o_rew) / np.maximum(np.abs(o_rew.astype(np.float64)), 1e-30)
        assert rel.max() <= RTOL, (t, rel.max())
    # open loop and host buffers on the same handles
    acts = rng.uniform(-1, 1, (3, n, A)).astype(np.float32); acts[:, ::17] = zero_action
    obs_m, rew_m, done_m = client.step_many(torch.as_tensor(acts, device="cuda:0"))
    obs_i, rew_i, done_i = ieee_c.step_many(torch.as_tensor(acts, device="cuda:0"))
    assert torch.equal(obs_m, obs_i) and torch.equal(done_m, done_i) and torch.equal(rew_m.view(torch.int32), rew_i.view(torch.int32))
    for t in range(3):
        o_obs, o_rew, o_done = ora.step(acts[t])
        assert np.array_equal(obs_m[t].cpu().numpy(), o_obs) and np.array_equal(done_m[t].cpu().numpy().astype(bool), o_done), t
    a_h, obs_h, rew_h, done_h = client.host_buffers()
    a_h[...] = rng.uniform(-1, 1, (n, A)).astype(np.float32); a_h[::13] = zero_action
    client.set_host_pipeline(stage_envs=4096, n_streams=2)
    client.step_host(a_h, obs_h, rew_h, done_h)
    o_obs, o_rew, o_done = ora.step(np.array(a_h))
    assert np.array_equal(obs_h, o_obs) and np.array_equal(done_h.astype(bool), o_done) and np.allclose(rew_h, o_rew, rtol=RTOL, atol=0)
    s, so = client.stats(), ora.stats()
    for k in ("steps", "episodes", "successes", "timeouts", "sum_episode_len", "holds", "violations"):
        assert s[k] == so[k], (k, s[k], so[k])
    assert (s["holds"] > 0) == can_hold and s["successes"] > 0
    assert np.array_equal(client.goal.cpu().numpy(), ora.goal)


@pytest.mark.parametrize("shape", ["msj_shaped", "generic"])
def test_one_sided_tendon_range_holds_at_the_edge_of_the_action_space_and_just_outside(shape):
    """Tendon range [0, hi]: the rescaled action is 0 at a = -1, and a = -1.0000001 -- outside the action space, so the
    error word is set -- still rescales to within 1e-8 of zero: numpy's allclose (simulation_client.py:38) holds there, and
    so must the kernels (the hold interval is searched over every float32, not only [-1, 1]).  Found by tools/soak_parity.py."""
    b = dict(angle_low=-1.5, angle_high=2.0, vel_low=-0.5, vel_high=0.5, act_low=0.0, act_high=0.3)
    if shape == "generic":
        b = dict(b, dim_joint=4, dim_action=5)
    J, A, _, bb = orc.robot_bounds(b)
    n = 1000
    env, client, ora = make_pair(b, n, seed=5, auto_reset=False)
    assert client.msj_kernels == (shape == "msj_shaped")
    assert np.array_equal(env.reset().cpu().numpy(), ora.reset())
    rng = np.random.default_rng(0)
    below = np.nextafter(np.float32(-1), np.float32(-2))
    for t in range(4):
        a = rng.uniform(-1, 1, (n, A)).astype(np.float32)
        a[0::5] = -1.0                      # holds
        a[1::5] = below                     # holds too, and is an action error
        a[2::5, 0] = below                  # one component outside, the others random: no hold, action error
        obs, rew, done, _ = env.step(torch.as_tensor(a, device="cuda:0"))
        o_obs, o_rew, o_done = ora.step(a)
        assert np.array_equal(obs.cpu().numpy(), o_obs) and np.array_equal(done.cpu().numpy(), o_done), t
    s, so = client.stats(), ora.stats()
    assert s["holds"] == so["holds"] == 4 * (len(range(0, n, 5)) + len(range(1, n, 5)))
    assert s["violations"] == so["violations"] and client.errors() == ora.errors() and client.errors() == (1, 1)


CANARY_ROBOTS = {
    "six_joints_14_tendons": GENERIC_ROBOTS["six_joints_14_tendons"],
    "fifteen_joints_64_tendons": GENERIC_ROBOTS["fifteen_joints_64_tendons"],
    "one_joint_3_tendons_one_sided": dict(dim_joint=1, dim_action=3, angle_low=0.0, angle_high=2.0, vel_low=0.0, vel_high=0.5, act_low=0.0, act_high=0.4),
    "eleven_joints_17_tendons": dict(dim_joint=11, dim_action=17, angle_low=-1.0, angle_high=2.0, vel_low=-0.5, vel_high=0.25, act_low=-0.125, act_high=0.125),
    "msj_shaped_other_limits": MSJ_SHAPED_OTHER_LIMITS["one_sided"],
}


@pytest.mark.parametrize("name", sorted(CANARY_ROBOTS))
def test_caller_buffers_are_written_inside_their_bounds_only(name):
    """(compute-sanitizer is not available on the GPU pool.)  Every output the fused step and the open-loop step write
    into CALLER buffers lands between two guard bands that stay untouched, for ragged populations
    (1 short of / 1 past a warp, a CTA, ...), and equals what a twin handle leaves in its own buffers."""
    b = CANARY_ROBOTS[name]
    J, A, _, bb = orc.robot_bounds(b)
    D, G = 3 * J, 2048
    zero_action, _ = orc.hold_action(b)
    rng = np.random.default_rng(1)
    for n in (2, 31, 33, 255, 257, 1000, 4097):
        for penalty in (False, True):
            e1, c1, _ = make_pair(b, n, seed=n, penalty=penalty)
            e2, c2, _ = make_pair(b, n, seed=n, penalty=penalty)
            e1.reset(); e2.reset()
            steps = rng.integers(396, 401, n).astype(np.int32)
            c1.set_step_num(steps); c2.set_step_num(steps)
            T = 3
            fbuf = torch.full((G + T * n * D + G,), float("nan"), device="cuda:0"); fbuf.view(torch.int32).fill_(0x7FC0DEAD)
            rbuf = torch.empty((G + T * n + G,), device="cuda:0"); rbuf.view(torch.int32).fill_(0x7FC0DEAD)
            dbuf = torch.full((G + T * n + G,), 0xAB, dtype=torch.uint8, device="cuda:0")
            obs, rew, done = fbuf[G:G + n * D].view(n, D), rbuf[G:G + n], dbuf[G:G + n]

            def guards_intact():
                ok = bool((fbuf[:G].view(torch.int32) == 0x7FC0DEAD).all()) and bool((fbuf[G + T * n * D:].view(torch.int32) == 0x7FC0DEAD).all())
                ok = ok and bool((rbuf[:G].view(torch.int32) == 0x7FC0DEAD).all()) and bool((rbuf[G + T * n:].view(torch.int32) == 0x7FC0DEAD).all())
                return ok and bool((dbuf[:G] == 0xAB).all()) and bool((dbuf[G + T * n:] == 0xAB).all())

            for t in range(2):
                a = rng.uniform(-1, 1, (n, A)).astype(np.float32); a[::3] = zero_action; a[1::11, A - 1] = np.nan
                a_dev = torch.as_tensor(a, device="cuda:0")
                c1.step_fused(a_dev, obs, rew, done)
                c2.step_fused(a_dev)
                torch.cuda.synchronize()
                assert guards_intact(), (n, penalty, t)
                assert torch.equal(obs.view(torch.int32), c2.obs.view(torch.int32)) and torch.equal(rew.view(torch.int32), c2.reward.view(torch.int32))
                assert torch.equal(done, c2.done)
                # unused tail of the three-step buffers stays untouched too
                assert bool((fbuf[G + n * D:G + T * n * D].view(torch.int32) == 0x7FC0DEAD).all()) and bool((dbuf[G + n:G + T * n] == 0xAB).all())
            acts = torch.as_tensor(rng.uniform(-1, 1, (T, n, A)).astype(np.float32), device="cuda:0")
            o3, r3, d3 = c1.step_many(acts, fbuf[G:G + T * n * D].view(T, n, D), rbuf[G:G + T * n].view(T, n), dbuf[G:G + T * n].view(T, n))
            o3b, r3b, d3b = c2.step_many(acts)
            torch.cuda.synchronize()
            assert guards_intact(), (n, penalty, "step_many")
            assert torch.equal(o3.view(torch.int32), o3b.view(torch.int32)) and torch.equal(d3, d3b) and torch.equal(r3.view(torch.int32), r3b.view(torch.int32))
            c1.close(); c2.close()
