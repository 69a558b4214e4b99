"""roboy_step_many (open loop, 73 B per env-step) over population sizes and window lengths: does the footprint (T planes of
n envs touched at once) bound it?  usage: python tools/open_loop_sweep.py"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gym_roboy_b200.envs import RoboyEnv
from gym_roboy_b200.envs.simulations import CudaSimulationClient
out = []
for n, T in ((1 << 20, 16), (1 << 22, 4), (1 << 22, 16), (1 << 22, 64), (1 << 24, 4), (1 << 24, 16)):
    c = CudaSimulationClient(num_envs=n, seed=1234, device="cuda:0")
    e = RoboyEnv(c); e.reset()
    c.set_step_num(((torch.arange(n, device="cuda:0") % 400) + 1).to(torch.int32))
    g = torch.Generator(device="cuda:0"); g.manual_seed(0)
    a = torch.rand((T, n, 8), device="cuda:0", generator=g) * 2 - 1
    obs = torch.empty((T, n, 9), device="cuda:0"); rew = torch.empty((T, n), device="cuda:0")
    dn = torch.empty((T, n), dtype=torch.uint8, device="cuda:0")
    for _ in range(3):
        c.step_many(a, obs, rew, dn)
    torch.cuda.synchronize()
    s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(5):
        c.step_many(a, obs, rew, dn)
    t.record(); torch.cuda.synchronize()
    ms = s.elapsed_time(t) / 5
    out.append({"envs": n, "T": T, "footprint_GB": 73 * n * T / 1e9, "ms": ms, "env_steps_per_s": n * T / ms * 1e3,
                "frac_of_6544": 73 * n * T / ms / 1e6 / 6544})
    c.close(); del c, e, a, obs, rew, dn
    torch.cuda.empty_cache()
print(json.dumps(out, indent=1))
