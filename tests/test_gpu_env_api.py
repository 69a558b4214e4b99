"""The reference's own unit tests (gym_roboy/envs/tests/test_roboy_env.py:28-207), restated
against this package's single-env surface: RoboyEnv(CudaSimulationClient(num_envs=1)).
Each test names the reference test it mirrors."""
from itertools import combinations

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture()
def parts():
    from gym_roboy_b200.envs import RoboyEnv
    from gym_roboy_b200.envs.robots import MsjRobot
    from gym_roboy_b200.envs.simulations import CudaSimulationClient
    robot = MsjRobot()
    client = CudaSimulationClient(robot=robot, num_envs=1, seed=17, device="cuda:0")
    return robot, client, RoboyEnv(simulation_client=client)


def new_env(**kw):
    from gym_roboy_b200.envs import RoboyEnv
    from gym_roboy_b200.envs.simulations import CudaSimulationClient
    return RoboyEnv(simulation_client=CudaSimulationClient(num_envs=1, seed=23, device="cuda:0"), **kw)


def test_roboy_env_step(parts):                                   # test_roboy_env.py:28-33
    _, _, env = parts
    env.reset()
    obs, reward, done, _ = env.step(env.action_space.sample())
    assert isinstance(obs, np.ndarray) and obs.shape == (9,)
    assert isinstance(reward, float)
    assert isinstance(done, bool), str(type(done))


def test_roboy_env_reset(parts):                                  # :36-46
    robot, _, env = parts
    num_zero = 2 * robot.get_joint_angles_space().shape[0]
    all_obs = [env.reset() for _ in range(5)]
    for obs in all_obs:
        assert np.allclose(obs[:num_zero], 0) and isinstance(obs, np.ndarray)
    for o1, o2 in combinations(all_obs, 2):
        assert np.allclose(o1[:num_zero], o2[:num_zero])
        assert not np.allclose(o1[num_zero:], o2[num_zero:])      # every reset draws a new goal


def test_new_goal_is_different_and_feasible(parts):               # :49-57
    robot, _, env = parts
    for _ in range(3):
        env._set_new_goal()
        old_goal = env._goal_state
        env._set_new_goal()
        new_goal = env._goal_state
        assert not np.allclose(old_goal.joint_angles, new_goal.joint_angles)
        assert np.all(robot.get_joint_angles_space().low <= new_goal.joint_angles)
        assert np.all(new_goal.joint_angles <= robot.get_joint_angles_space().high)


def test_reaching_goal_angle_delivers_maximum_reward(parts):      # :60-68
    _, _, env = parts
    env.reset()
    env._set_new_goal(goal_joint_angle=env._last_state.joint_angles)
    _, reward, done, _ = env.step(np.zeros(len(env.action_space.low)))
    assert np.isclose(reward, env.reward_range[1]) and done


def test_reaching_goal_but_moving_is_not_done(parts):             # :71-79
    robot, _, env = parts
    env.reset()
    state = env._last_state
    env._set_new_goal(goal_joint_angle=state.joint_angles)
    state.joint_vels = robot.get_joint_vels_space().high
    assert not env._did_reach_goal(current_state=state, goal_state=env._goal_state)


def test_joint_vel_penalty_affects_worst_possible_reward(parts):  # :82-89
    robot, _, _ = parts
    env = new_env(joint_vel_penalty=False)
    largest = np.linalg.norm(2 * np.ones(robot.get_joint_angles_space().shape))
    worst = -np.exp(largest) - abs(env._PENALTY_FOR_TOUCHING_BOUNDARY)
    assert np.isclose(env.reward_range[0], worst)
    assert new_env(joint_vel_penalty=True).reward_range[0] < worst


def test_reward_is_lower_with_joint_vel_penalty(parts):           # :92-108
    from gym_roboy_b200.envs.robots import MsjRobot, RobotState
    s = MsjRobot.new_random_state()
    moving = RobotState(s.joint_angles, MsjRobot.get_joint_vels_space().sample(), True)
    goal = RobotState(s.joint_angles, np.zeros(3), True)
    r_plain = new_env(joint_vel_penalty=False).compute_reward(moving, goal)
    r_pen = new_env(joint_vel_penalty=True).compute_reward(moving, goal)
    assert r_plain > r_pen


def test_agent_gets_bonus_when_reaching_the_goal():               # :111-122
    e0 = new_env(is_agent_getting_bonus_for_reaching_goal=False)
    e0.reset()
    r0 = e0.compute_reward(current_state=e0._goal_state, goal_state=e0._goal_state)
    e1 = new_env(is_agent_getting_bonus_for_reaching_goal=True)
    e1.reset()
    r1 = e1.compute_reward(current_state=e1._goal_state, goal_state=e1._goal_state)
    assert np.allclose(r1 - r0, e1._BONUS_FOR_REACHING_GOAL)


def test_render_does_nothing(parts):                              # :125-126
    parts[2].render()


@pytest.mark.parametrize("penalty", [True, False], ids=["with joint_vel penalty", "no joint_vel penalty"])
def test_reward_monotonously_improves_during_approach(penalty):   # :129-167
    from gym_roboy_b200.envs.robots import MsjRobot, RobotState
    env = new_env(joint_vel_penalty=penalty, strict=False)
    robot = MsjRobot()
    for space in (robot.get_joint_angles_space(), robot.get_joint_vels_space()):
        space.seed(0)
    goal = robot.new_random_zero_vels_state()
    starts = [robot.new_random_state() for _ in range(40)] + [robot.new_random_zero_vels_state()]
    if penalty:
        starts.append(robot.new_random_zero_angles_state())
    for cur in starts:
        rewards = []
        for _ in range(7):
            rewards.append(env.compute_reward(current_state=cur, goal_state=goal))
            cur = RobotState.interpolate(cur, goal)
        assert all(x < y for x, y in zip(rewards, rewards[1:])), rewards


def test_maximum_episode_length():                                # :170-180
    env = new_env()
    env.reset()
    env.step_num = env._MAX_EPISODE_LENGTH - 1
    _, _, done, _ = env.step(np.zeros(env.action_space.shape))
    assert not done
    assert not env._did_reach_goal(env._last_state, env._goal_state)
    _, _, done, _ = env.step(env.action_space.sample())
    assert done
    # reference quirk: done does not reset step_num; the next step is done again
    _, _, done, _ = env.step(env.action_space.sample())
    assert done and env.step_num == env._MAX_EPISODE_LENGTH + 2


def test_reset_sets_step_number_to_one(parts):                    # :183-188
    _, _, env = parts
    env.step(env.action_space.sample())
    assert env.step_num != 1
    env.reset()
    assert env.step_num == 1


def test_action_outside_action_space_raises(parts):               # roboy_env.py:52 ([probe] in SURVEY 8a a2)
    _, _, env = parts
    env.reset()
    for bad in (np.full(8, 1.0000001, np.float32), np.full(8, np.nan, np.float32), np.zeros(7, np.float32)):
        with pytest.raises(AssertionError):
            env.step(bad)


def test_goal_outside_angle_space_raises(parts):                  # roboy_robot.py:76
    _, _, env = parts
    with pytest.raises(AssertionError):
        env._set_new_goal(goal_joint_angle=np.array([4.0, 0.0, 0.0], np.float32))


def test_batched_env_flags_bad_actions_on_device():
    import torch
    from gym_roboy_b200.envs import RoboyEnv
    from gym_roboy_b200.envs.simulations import CudaSimulationClient
    env = RoboyEnv(CudaSimulationClient(num_envs=100, seed=1, env_id_base=1000, device="cuda:0"))
    env.reset()
    a = torch.zeros((100, 8), device="cuda:0")
    a[37, 5] = 1.5
    a[80, 0] = float("nan")
    env.step(a)
    word, first = env._simulation_client.errors()
    assert word == 1 and first == 1037
    assert env.episode_stats()["violations"] == 2
    with pytest.raises(AssertionError):
        env.check_errors()
    env.check_errors()   # cleared by the raise


def test_make_msj_control_v1():
    import gym_roboy_b200
    env = gym_roboy_b200.make("msj-control-v1")                   # README.md:24 of the reference
    obs = env.reset()
    assert obs.shape == (9,) and env.observation_space.shape == (9,) and env.action_space.shape == (8,)
    venv = gym_roboy_b200.make("msj-control-v1", num_envs=64, seed=3)
    assert venv.reset().shape == (64, 9)
