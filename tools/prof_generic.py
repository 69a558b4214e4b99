# profiling driver: the generic fused step for a 6-joint / 14-tendon robot, steady state (used under ncu)
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from gym_roboy_b200.envs import RoboyEnv
from gym_roboy_b200.envs.robots import RoboyRobot
from gym_roboy_b200.envs.simulations import CudaSimulationClient
from gym_roboy_b200.spaces import Box
J, A = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (6, 14)


class R(RoboyRobot):
    _A = Box(-2.5, 2.5, (J,), "float32"); _V = Box(-0.6, 0.6, (J,), "float32"); _T = Box(-0.2, 0.2, (A,), "float32")
    get_action_space = classmethod(lambda cls: cls._T)
    get_joint_angles_space = classmethod(lambda cls: cls._A)
    get_joint_vels_space = classmethod(lambda cls: cls._V)


n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 22
c = CudaSimulationClient(robot=R(), num_envs=n, seed=1, device="cuda:0"); e = RoboyEnv(c, strict=False); e.reset()
c.set_step_num(((torch.arange(n, device="cuda:0") % 400) + 1).to(torch.int32))
g = torch.Generator(device="cuda:0"); g.manual_seed(0)
a = [torch.rand((n, A), device="cuda:0", generator=g) * 2 - 1 for _ in range(2)]
for i in range(4): c.step_fused(a[i & 1])
torch.cuda.synchronize(); print("ok", c.stats()["steps"])
