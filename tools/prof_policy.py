# profiling driver: a few fused policy rollouts at a given size (used under ncu)
import sys, torch
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gym_roboy_b200.envs import RoboyEnv
from gym_roboy_b200.envs.simulations import CudaSimulationClient
from gym_roboy_b200.rollout import MlpPolicy, RolloutCollector
n = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
T = int(sys.argv[2]) if len(sys.argv) > 2 else 16
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
mode = sys.argv[4] if len(sys.argv) > 4 else "fp32"
torch.manual_seed(0)
c = CudaSimulationClient(num_envs=n, seed=1234, device="cuda:0")
col = RolloutCollector(RoboyEnv(c), MlpPolicy().to("cuda:0"), n_steps=T, fused=mode)
for _ in range(reps): col.collect()
torch.cuda.synchronize(); print("ok", c.stats()["steps"])
