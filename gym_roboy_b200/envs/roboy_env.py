"""`RoboyEnv`: the goal-reaching environment, for N environments per launch on a B200.

Same constructor, methods and attributes as the reference env (envs/roboy_env.py:10-158), so
code written against `RoboyEnv(simulation_client, seed, joint_vel_penalty,
is_agent_getting_bonus_for_reaching_goal)` keeps working:

* `CudaSimulationClient(num_envs=1)`  ->  the classic single-env gym surface: `step(action)`
  takes 8 floats and returns `(np.ndarray(9,), float, bool, {})`; contract violations raise
  `AssertionError` exactly where the reference does (roboy_env.py:52,109); no auto-reset.
* `CudaSimulationClient(num_envs=N)`  ->  the vectorised surface: `step(actions[N,8])` returns
  zero-copy CUDA tensors `(obs[N,9], reward[N], done[N], info)` and finished envs are reset
  inside the same kernel (what the vec-env worker does around the reference env).  Violations
  are recorded on the device (`check_errors()` raises) instead of raising per env.

Everything `step` / `reset` / `compute_reward` compute happens in the CUDA kernels behind
include/roboy_b200.h; this file is host-side plumbing only.
"""
import numpy as np
import torch

from .. import _native
from ..spaces import Box
from .robots import RobotState
from .simulations import CudaSimulationClient

try:  # gym is optional: subclass its GoalEnv when present so isinstance checks keep working
    import gym as _gym
    _GoalEnvBase = _gym.GoalEnv
except Exception:  # pragma: no cover - gym is not in the image
    class _GoalEnvBase:
        metadata = {"render.modes": []}
        reward_range = (-float("inf"), float("inf"))

        def close(self):
            pass

        @property
        def unwrapped(self):
            return self


class RoboyEnv(_GoalEnvBase):

    def __init__(self, simulation_client: CudaSimulationClient = None, seed: int = None,
                 joint_vel_penalty: bool = False, is_agent_getting_bonus_for_reaching_goal: bool = True,
                 auto_reset: bool = None, strict: bool = None, num_envs: int = None, device=None):
        if simulation_client is None:
            simulation_client = CudaSimulationClient(num_envs=num_envs or 1, device=device, seed=seed)
        if not isinstance(simulation_client, CudaSimulationClient):
            raise TypeError("this RoboyEnv runs fused on the GPU and needs a CudaSimulationClient; other "
                            "SimulationClient implementations plug into the reference's RoboyEnv instead")
        self._simulation_client = client = simulation_client
        self.num_envs = client.num_envs
        self._single = self.num_envs == 1
        self._joint_vel_penalty = bool(joint_vel_penalty)
        self._is_agent_getting_bonus_for_reaching_goal = bool(is_agent_getting_bonus_for_reaching_goal)
        self._auto_reset = (not self._single) if auto_reset is None else bool(auto_reset)
        self._strict = self._single if strict is None else bool(strict)
        self._robot = robot = client.robot
        if seed is not None:
            # the reference seeds before it draws its first goal (roboy_env.py:15 then :37): redo the construction draws
            client.reseed(seed)
        client.configure_env(self._joint_vel_penalty, self._is_agent_getting_bonus_for_reaching_goal, self._auto_reset)

        angles, vels = robot.get_joint_angles_space(), robot.get_joint_vels_space()
        self._GOAL_JOINT_VEL = robot.new_zero_state().joint_angles       # roboy_env.py:23 (float64 zeros)
        self._PENALTY_FOR_TOUCHING_BOUNDARY = 1                          # roboy_env.py:26
        self._BONUS_FOR_REACHING_GOAL = 1000                             # roboy_env.py:27
        self._MAX_EPISODE_LENGTH = 400                                   # roboy_env.py:28
        self.action_space = Box(-1, 1, robot.get_action_space().shape, "float32")             # roboy_env.py:31
        self.observation_space = Box(np.concatenate((angles.low, vels.low, angles.low)),      # roboy_env.py:32-36
                                     np.concatenate((angles.high, vels.high, angles.high)), dtype="float32")
        self.reward_range = self._create_reward_range()
        client.set_reward_range(*self.reward_range)
        self._last_obs = None
        self._info = {}                      # batched: {'terminal_observation': tensor} once enabled on the client
        client.info_sink = self._info
        if client.terminal_obs is not None:
            self._info["terminal_observation"] = client.terminal_obs

    # ------------------------------------------------------------------ reference surface
    def _create_reward_range(self):
        """roboy_env.py:40-49: best and worst reward, from two compute_reward calls (on the GPU)."""
        angles, vels = self._robot.get_joint_angles_space(), self._robot.get_joint_vels_space()
        q = np.stack([angles.high, angles.low])
        qd = np.stack([vels.high, vels.low])
        goal_q, goal_qd = np.stack([angles.high] * 2), np.stack([vels.high] * 2)
        reward, _ = self._simulation_client.compute_reward(q, qd, [1, 0], goal_q, goal_qd, check_range=_native.REWARD_RANGE_PROBE)
        max_reward, min_reward = reward.tolist()
        return min_reward, max_reward

    def step(self, action):
        client = self._simulation_client
        if not self._single and not self._strict and torch.is_tensor(action) and action.is_cuda \
                and action.dtype == torch.float32 and action.is_contiguous():
            client.step_fused(action)   # hot path of a batched trainer: no conversions, no sync
            self._last_obs = client.obs
            return client.obs, client.reward, client.done, self._info
        if self._single and not torch.is_tensor(action):
            a = np.asarray(action)
            assert self.action_space.contains(a)                                              # roboy_env.py:52
            actions = torch.as_tensor(a.astype(np.float32), device=client.device).reshape(1, -1)
        else:
            actions = action if torch.is_tensor(action) else torch.as_tensor(np.asarray(action, np.float32))
            actions = actions.to(device=client.device, dtype=torch.float32).reshape(self.num_envs, -1).contiguous()
        client.step_fused(actions)
        self._last_obs = client.obs
        if self._strict:
            self.check_errors()
        if self._single:
            return (client.obs[0].cpu().numpy(), float(client.reward[0].item()), bool(client.done[0].item()), {})
        return client.obs, client.reward, client.done, self._info

    def step_from_states(self, joint_angles, joint_vels, is_feasible=None):
        """`step` for envs whose simulator runs elsewhere: the caller has already
        forwarded its actions; this computes obs / reward / done / goal resample for the returned
        states (`[N,3]`, `[N,3]`, `[N]`).  No auto-reset -- see `reset_from_states`."""
        client = self._simulation_client
        client.step_external(joint_angles, joint_vels, is_feasible)
        self._last_obs = client.obs
        if self._strict:
            self.check_errors()
        return client.obs, client.reward, client.done, {}

    def reset_from_states(self, joint_angles, joint_vels, mask=None):
        client = self._simulation_client
        client.reset_external(joint_angles, joint_vels, mask)
        self._last_obs = client.obs
        return client.obs

    def reset(self, mask=None):
        """roboy_env.py:82-87.  Batched envs may reset only the envs selected by a `[N]` mask."""
        client = self._simulation_client
        client.reset_fused(mask)
        self._last_obs = client.obs
        return client.obs[0].cpu().numpy() if self._single else client.obs

    def render(self, mode="human"):
        pass

    def seed(self, seed=None):
        """roboy_env.py:114-115.  Re-keys the device generator (the reference seeds numpy's global RNG)."""
        if seed is not None:
            self._simulation_client.set_seed(seed)

    def compute_reward(self, current_state: RobotState, goal_state: RobotState, info=None):
        """roboy_env.py:92-112 on the GPU: python float for one state, float64 tensor `[k]` for a batch."""
        reward, _ = self._reward_and_reached(current_state, goal_state, check_range=True)
        if self._strict:
            self.check_errors()
        return float(reward[0].item()) if reward.numel() == 1 and not torch.is_tensor(current_state.joint_angles) \
            else reward

    def _did_reach_goal(self, current_state: RobotState, goal_state: RobotState):
        """roboy_env.py:125-134"""
        _, reached = self._reward_and_reached(current_state, goal_state, check_range=False)
        return bool(reached[0].item()) if reached.numel() == 1 and not torch.is_tensor(current_state.joint_angles) \
            else reached

    def _reward_and_reached(self, current_state, goal_state, check_range):
        gv = goal_state.joint_vels
        # the env's own goal carries float64 zero velocities (roboy_env.py:23); the kernel has that
        # case built in (goal_qd = NULL).  Anything else is taken as float32 values.
        zero64 = (not torch.is_tensor(gv)) and np.asarray(gv).dtype == np.float64 and not np.any(gv)
        feasible = current_state.is_feasible
        if isinstance(feasible, bool):
            feasible = [feasible]
        return self._simulation_client.compute_reward(current_state.joint_angles, current_state.joint_vels, feasible,
                                                      goal_state.joint_angles, None if zero64 else gv,
                                                      check_range=check_range)

    def _set_new_goal(self, goal_joint_angle=None):
        """roboy_env.py:117-123: a given goal (bounds-checked) or a fresh draw from the client."""
        client = self._simulation_client
        if goal_joint_angle is None:
            goal_joint_angle = client.get_new_goal_joint_angles()
        if self._single:
            self._robot.new_state(joint_angle=np.asarray(goal_joint_angle), joint_vel=self._GOAL_JOINT_VEL,
                                  is_feasible=True)                                           # roboy_robot.py:76
        client.set_goal(goal_joint_angle)
        if self._strict:
            self.check_errors()

    def _reached_max_steps(self):
        return self.step_num > self._MAX_EPISODE_LENGTH

    # ------------------------------------------------------------------ attributes the reference's tests poke
    @property
    def step_num(self):
        s = self._simulation_client.step_num
        return int(s[0].item()) if self._single else s

    @step_num.setter
    def step_num(self, value):
        self._simulation_client.set_step_num(value)

    @property
    def _goal_state(self) -> RobotState:
        g = self._simulation_client.goal
        if self._single:
            return RobotState(g[:, 0].cpu().numpy(), self._GOAL_JOINT_VEL, True)
        return RobotState(g.t(), torch.zeros_like(g.t()), torch.ones(self.num_envs, dtype=torch.bool, device=g.device))

    @property
    def _last_state(self) -> RobotState:
        if self._last_obs is None:
            return None
        o = self._last_obs
        J = self._simulation_client.dim_joint
        if self._single:
            o = o[0].cpu().numpy()
            return RobotState(o[0:J], o[J:2 * J], True)
        return RobotState(o[:, 0:J], o[:, J:2 * J], torch.ones(self.num_envs, dtype=torch.bool, device=o.device))

    # ------------------------------------------------------------------ batched extras
    def check_errors(self):
        """Raise the AssertionError the reference would have raised (roboy_env.py:52,109; roboy_robot.py:76)."""
        flags, first = self._simulation_client.errors()
        if flags:
            what = [name for bit, name in ((_native.ERR_ACTION, "action outside action_space (roboy_env.py:52)"),
                                           (_native.ERR_REWARD_RANGE, "reward outside reward_range (roboy_env.py:109)"),
                                           (_native.ERR_GOAL_BOUNDS, "goal outside the joint angle space (roboy_robot.py:76)"),
                                           (_native.ERR_STATE_BOUNDS, "external state outside the joint angle space (roboy_robot.py:76)"))
                    if flags & bit]
            self._simulation_client.clear_errors()
            raise AssertionError("{}; first offending env id {}".format("; ".join(what), first))

    def episode_stats(self):
        """Counters accumulated by the step kernel (this shard only; see sharding.all_reduce_stats)."""
        return self._simulation_client.stats()

    def close(self):
        self._simulation_client.close()


def _l2_distance(joint_angle1, joint_angle2):
    """Host helper with the reference's semantics (roboy_env.py:137-140): NaN differences count as 0."""
    diff = np.subtract(joint_angle1, joint_angle2)
    diff[np.isnan(diff)] = 0
    return np.linalg.norm(diff, ord=2)


def _rescale_from_one_space_to_other(input_val: np.ndarray, input_space: Box, output_space: Box) -> np.ndarray:
    """Host helper with the reference's semantics (roboy_env.py:143-158): the affine map that sends
    `input_space` onto `output_space`, evaluated as `slope * (x - in.high) + out.high`.  The step kernel
    applies the same float32 map to every action (as the pre-image of its hold test); this version
    exists for callers and tests that use the helper directly."""
    if not isinstance(input_val, np.ndarray):
        raise TypeError("input_val must be a numpy array")
    assert input_space.shape == output_space.shape
    assert input_space.contains(input_val)
    slope = (output_space.high - output_space.low) / (input_space.high - input_space.low)
    return slope * (input_val - input_space.high) + output_space.high

