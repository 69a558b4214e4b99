"""Experiment: roboy_step_host throughput vs pipeline stage size / stream count."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gym_roboy_b200.envs import RoboyEnv
from gym_roboy_b200.envs.simulations import CudaSimulationClient
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 24
c = CudaSimulationClient(num_envs=n, seed=1, device="cuda:0"); e = RoboyEnv(c); e.reset()
pin = lambda shape, dt: torch.empty(shape, dtype=dt).pin_memory()
a = pin((n, 8), torch.float32); a.uniform_(-1, 1)
obs, rew, done = pin((n, 9), torch.float32), pin((n,), torch.float32), pin((n,), torch.uint8)
bufs = (a.numpy(), obs.numpy(), rew.numpy(), done.numpy())
for stage in (1 << 16, 1 << 17, 1 << 18, 1 << 19, 1 << 20, 1 << 21, 1 << 22):
    for ns in (2, 3, 4, 8):
        c.set_host_pipeline(stage, ns)
        for _ in range(2): c.step_host(*bufs)
        t0 = time.perf_counter(); k = 6
        for _ in range(k): c.step_host(*bufs)
        dt = (time.perf_counter() - t0) / k
        print("stage %8d streams %d: %7.3f ms/step  %.3e env-steps/s  D2H %.1f GB/s H2D %.1f GB/s" % (
            stage, ns, dt * 1e3, n / dt, 41 * n / dt / 1e9, 32 * n / dt / 1e9), flush=True)
