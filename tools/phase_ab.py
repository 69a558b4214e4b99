"""A/B of the step kernel at 16,777,216 envs: all envs in lock-step right after reset (no episode ends for 400 steps -- what
round 1's bench timed) vs episode phases spread over the 400-step episode (1/400 of the envs finish in every step -- the
steady state bench.py times now).  usage: python tools/phase_ab.py [envs] [steps]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gym_roboy_b200.envs import RoboyEnv
from gym_roboy_b200.envs.simulations import CudaSimulationClient

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 24
K = int(sys.argv[2]) if len(sys.argv) > 2 else 100
out = {}
gen = torch.Generator(device="cuda:0"); gen.manual_seed(0)
acts = [torch.rand((n, 8), device="cuda:0", generator=gen) * 2 - 1 for _ in range(2)]
for label in ("lockstep", "phases", "phases_done_index"):
    c = CudaSimulationClient(num_envs=n, seed=1234, device="cuda:0")
    e = RoboyEnv(c); e.reset()
    if label.startswith("phases"):
        c.set_step_num(((torch.arange(n, device="cuda:0") % 400) + 1).to(torch.int32))
    if label == "phases_done_index":
        c.enable_done_index(True)
    for i in range(10):
        c.step_fused(acts[i & 1])
    torch.cuda.synchronize()
    s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for i in range(K):
        c.step_fused(acts[i & 1])
    t.record(); torch.cuda.synchronize()
    ms = s.elapsed_time(t) / K
    out[label] = {"ms_per_step": ms, "GBps": 93 * n / ms / 1e6, "frac_of_6544": 93 * n / ms / 1e6 / 6544, "episodes": c.stats()["episodes"]}
    c.close(); del c, e
# joint_vel_penalty=True (float64 promotion on the reward path), steady state
c = CudaSimulationClient(num_envs=n, seed=1234, device="cuda:0")
e = RoboyEnv(c, joint_vel_penalty=True, strict=False); e.reset()
c.set_step_num(((torch.arange(n, device="cuda:0") % 400) + 1).to(torch.int32))
for i in range(10):
    c.step_fused(acts[i & 1])
torch.cuda.synchronize()
s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for i in range(K):
    c.step_fused(acts[i & 1])
t.record(); torch.cuda.synchronize()
ms = s.elapsed_time(t) / K
out["penalty_phases"] = {"ms_per_step": ms, "GBps": 93 * n / ms / 1e6, "frac_of_6544": 93 * n / ms / 1e6 / 6544}
c.close(); del c, e
# open loop (roboy_step_many, 73 B per env-step), 4,194,304 envs x 16 steps
n2, T = 1 << 22, 16
c = CudaSimulationClient(num_envs=n2, seed=1234, device="cuda:0")
e = RoboyEnv(c); e.reset()
c.set_step_num(((torch.arange(n2, device="cuda:0") % 400) + 1).to(torch.int32))
a = torch.rand((T, n2, 8), device="cuda:0", generator=gen) * 2 - 1
obs = torch.empty((T, n2, 9), device="cuda:0"); rew = torch.empty((T, n2), device="cuda:0")
dn = torch.empty((T, n2), dtype=torch.uint8, device="cuda:0")
for _ in range(3):
    c.step_many(a, obs, rew, dn)
torch.cuda.synchronize()
s.record()
for _ in range(5):
    c.step_many(a, obs, rew, dn)
t.record(); torch.cuda.synchronize()
ms = s.elapsed_time(t) / 5
out["open_loop"] = {"ms_per_launch": ms, "env_steps_per_s": n2 * T / ms * 1e3, "frac_of_6544": 73 * n2 * T / ms / 1e6 / 6544}
print(json.dumps(out, indent=1))
