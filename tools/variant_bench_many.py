"""Kernel-experiment helper: time roboy_step_many (open loop) for one build of the library."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gym_roboy_b200.envs import RoboyEnv
from gym_roboy_b200.envs.simulations import CudaSimulationClient
n, T = 1 << 22, 16
c = CudaSimulationClient(num_envs=n, seed=1234, device="cuda:0"); e = RoboyEnv(c); e.reset()
g = torch.Generator(device="cuda:0"); g.manual_seed(1)
a = torch.rand((T, n, 8), device="cuda:0", generator=g) * 2 - 1
obs = torch.empty((T, n, 9), device="cuda:0"); rew = torch.empty((T, n), device="cuda:0"); dn = torch.empty((T, n), dtype=torch.uint8, device="cuda:0")
for _ in range(3): c.step_many(a, obs, rew, dn)
torch.cuda.synchronize()
s, f = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(5): c.step_many(a, obs, rew, dn)
f.record(); torch.cuda.synchronize()
ms = s.elapsed_time(f) / 5
print("%-20s open-loop %.4f ms/launch  %.3e env-steps/s  %.0f GB/s (73 B)" % (os.path.basename(os.environ.get("ROBOY_B200_LIB", "default")), ms, n * T / ms * 1e3, 73 * n * T / ms / 1e6))
