import os, sys, torch
sys.path.insert(0, '.')
from gym_roboy_b200.envs import RoboyEnv
from gym_roboy_b200.envs.simulations import CudaSimulationClient
for n in (65536, 262144, 1048576, 4194304):
    c = CudaSimulationClient(num_envs=n, seed=1234, device="cuda:0"); e = RoboyEnv(c); e.reset()
    g = torch.Generator(device="cuda:0"); g.manual_seed(0)
    acts = [torch.rand((n, 8), device="cuda:0", generator=g) * 2 - 1 for _ in range(2)]
    gr = torch.cuda.CUDAGraph()
    for i in range(3): c.step_fused(acts[i & 1])
    torch.cuda.synchronize()
    with torch.cuda.graph(gr):
        for i in range(100): c.step_fused(acts[i & 1])
    for _ in range(2): gr.replay()
    torch.cuda.synchronize()
    s, f = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(5): gr.replay()
    f.record(); torch.cuda.synchronize()
    us = s.elapsed_time(f) / 500 * 1e3
    print("%8d envs  %.2f us/step  %.3e env-steps/s  %.0f GB/s  grid=%s" % (n, us, n / us * 1e6, 93 * n / us / 1e3, c.step_geometry()), flush=True)
