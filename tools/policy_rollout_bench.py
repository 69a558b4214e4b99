import sys, time, torch
sys.path.insert(0, '.')
from gym_roboy_b200.envs import RoboyEnv
from gym_roboy_b200.envs.simulations import CudaSimulationClient
from gym_roboy_b200.rollout import MlpPolicy, RolloutCollector
for n, ept in ((4096, 0), (32768, 0), (262144, 0), (262144, 1), (1048576, 0)):
    torch.manual_seed(0)
    c = CudaSimulationClient(num_envs=n, seed=1, device='cuda:0')
    col = RolloutCollector(RoboyEnv(c), MlpPolicy().to('cuda:0'), n_steps=128, fused=True, envs_per_thread=ept)
    for _ in range(2): col.collect()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5
    s.record()
    for _ in range(reps): col.collect()
    e.record(); torch.cuda.synchronize()
    ms = s.elapsed_time(e) / reps
    print(n, ept, 'ms/rollout', ms, 'env-steps/s %.3e' % (n * 128 / ms * 1e3), 'GFLOP/s %.1f' % (n*128*20736/ms*1e3/1e9), flush=True)
    del col, c
