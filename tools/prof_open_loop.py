# profiling driver: the open-loop T-step kernel (roboy_step_many), steady-state episode phases (used under ncu)
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gym_roboy_b200.envs import RoboyEnv
from gym_roboy_b200.envs.simulations import CudaSimulationClient
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 22
T = int(sys.argv[2]) if len(sys.argv) > 2 else 16
c = CudaSimulationClient(num_envs=n, seed=1234, device="cuda:0"); e = RoboyEnv(c); e.reset()
c.set_step_num(((torch.arange(n, device="cuda:0") % 400) + 1).to(torch.int32))
g = torch.Generator(device="cuda:0"); g.manual_seed(0)
a = torch.rand((T, n, 8), device="cuda:0", generator=g) * 2 - 1
obs = torch.empty((T, n, 9), device="cuda:0"); rew = torch.empty((T, n), device="cuda:0")
dn = torch.empty((T, n), dtype=torch.uint8, device="cuda:0")
for _ in range(3):
    c.step_many(a, obs, rew, dn)
torch.cuda.synchronize(); print("ok", c.stats()["steps"])
