# profiling driver: a few fused steps at a given size (used under ncu)
import sys, torch
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gym_roboy_b200.envs import RoboyEnv
from gym_roboy_b200.envs.simulations import CudaSimulationClient
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 24
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
c = CudaSimulationClient(num_envs=n, seed=1234, device="cuda:0"); e = RoboyEnv(c); e.reset()
c.set_step_num(((torch.arange(n, device="cuda:0") % 400) + 1).to(torch.int32))   # steady state: 1/400 of the envs finish per step
g = torch.Generator(device="cuda:0"); g.manual_seed(0)
a = [torch.rand((n, 8), device="cuda:0", generator=g) * 2 - 1 for _ in range(2)]
for i in range(steps): e.step(a[i & 1])
torch.cuda.synchronize(); print("ok", c.stats()["steps"])
