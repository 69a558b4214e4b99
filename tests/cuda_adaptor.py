"""Adaptor driving the CUDA path (through CudaSimulationClient -> ctypes -> C-ABI) for golden_replay."""
import numpy as np
import torch

from golden_replay import Adaptor
from gym_roboy_b200.envs import RoboyEnv
from gym_roboy_b200.envs.simulations import CudaSimulationClient


def robot_from_bounds(bounds):
    """A RoboyRobot plug-in (this package's base class) with the given spaces; None -> MsjRobot."""
    if bounds is None:
        return None
    from gym_roboy_b200.envs.robots import RoboyRobot
    from gym_roboy_b200.spaces import Box
    from oracle import oracle as orc
    _, _, _, b = orc.robot_bounds(bounds)

    class FixtureRobot(RoboyRobot):
        _A = Box(b["angle_low"], b["angle_high"], dtype="float32")
        _V = Box(b["vel_low"], b["vel_high"], dtype="float32")
        _T = Box(b["act_low"], b["act_high"], dtype="float32")
        get_action_space = classmethod(lambda cls: cls._T)
        get_joint_angles_space = classmethod(lambda cls: cls._A)
        get_joint_vels_space = classmethod(lambda cls: cls._V)

    return FixtureRobot()


class CudaAdaptor(Adaptor):
    def __init__(self, fx, env_id_base=0):
        self.client = CudaSimulationClient(robot=robot_from_bounds(fx.get("bounds")), num_envs=fx["N"], seed=fx["seed"],
                                           env_id_base=env_id_base, device="cuda:0")
        self.env = RoboyEnv(self.client, joint_vel_penalty=fx["joint_vel_penalty"],
                            is_agent_getting_bonus_for_reaching_goal=fx["bonus"], auto_reset=fx["auto_reset"],
                            strict=False)
        self.client.enable_terminal_obs(True)

    def goals(self):
        return self.client.goal.t().cpu().numpy()

    def step_nums(self):
        return self.client.step_num.cpu().numpy()

    def set_goal(self, e, g):
        self.client.set_goal(np.asarray(g, np.float32), idx=[e])

    def set_state(self, e, q, qd, feasible):
        self.client.set_state(q, qd, feasible=[feasible], idx=[e])

    def set_step_num(self, e, k):
        self.client.set_step_num([k], idx=[e])

    def reset(self, mask=None):
        n = self.client.num_envs
        if mask is None:
            return self.env.reset().cpu().numpy().reshape(n, -1).copy()
        out = self.env.reset(mask=torch.as_tensor(np.asarray(mask, np.uint8)))
        return out.cpu().numpy().reshape(n, -1).copy()

    def step(self, actions):
        a = torch.as_tensor(np.ascontiguousarray(actions, np.float32), device="cuda:0")
        obs, reward, done, info = self.env.step(a)
        n = self.client.num_envs
        return (obs.cpu().numpy().reshape(n, -1), reward.cpu().numpy().reshape(n), done.cpu().numpy().reshape(n),
                info["terminal_observation"].cpu().numpy())

    def violations(self):
        word, first = self.client.errors()
        return word, first
