# closed-loop rollout micro-bench: fused float32 and tensor-core (TF32) kernels at several sizes
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gym_roboy_b200.envs import RoboyEnv
from gym_roboy_b200.envs.simulations import CudaSimulationClient
from gym_roboy_b200.rollout import MlpPolicy, RolloutCollector
cases = [(4096, "fp32"), (4096, "tc"), (32768, "tc"), (262144, "fp32"), (262144, "tc"), (1048576, "fp32"), (1048576, "tc")]
if len(sys.argv) > 1:
    cases = [tuple(a.split(":")) for a in sys.argv[1:]]
for case in cases:
    n, mode, ept = int(case[0]), case[1], int(case[2]) if len(case) > 2 else 0
    torch.manual_seed(0)
    c = CudaSimulationClient(num_envs=n, seed=1, device='cuda:0')
    col = RolloutCollector(RoboyEnv(c), MlpPolicy().to('cuda:0'), n_steps=128, fused=mode, envs_per_thread=ept)
    for _ in range(2): col.collect()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5
    s.record()
    for _ in range(reps): col.collect()
    e.record(); torch.cuda.synchronize()
    ms = s.elapsed_time(e) / reps
    print(n, mode, ept, 'ms/rollout %.3f' % ms, 'env-steps/s %.3e' % (n * 128 / ms * 1e3), 'TFLOP/s %.1f' % (n*128*20736/ms*1e3/1e12), flush=True)
    del col, c
