"""Kernel-experiment helper: roboy_step_many (open loop, 73 B per env-step), steady state, for ONE build of the library
(ROBOY_B200_LIB=...), with a digest of the outputs so that builds can be compared.  usage: python tools/open_loop_variant.py"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gym_roboy_b200.envs import RoboyEnv
from gym_roboy_b200.envs.simulations import CudaSimulationClient
res = {"lib": os.environ.get("ROBOY_B200_LIB", "default")}
for n, T in ((1 << 22, 4), (1 << 22, 16), (1 << 24, 16), (1 << 22, 64)):
    c = CudaSimulationClient(num_envs=n, seed=1234, device="cuda:0")
    e = RoboyEnv(c); e.reset()
    c.set_step_num(((torch.arange(n, device="cuda:0") % 400) + 1).to(torch.int32))
    g = torch.Generator(device="cuda:0"); g.manual_seed(0)
    a = torch.rand((T, n, 8), device="cuda:0", generator=g) * 2 - 1
    obs = torch.empty((T, n, 9), device="cuda:0"); rew = torch.empty((T, n), device="cuda:0")
    dn = torch.empty((T, n), dtype=torch.uint8, device="cuda:0")
    c.step_many(a, obs, rew, dn)
    digest = [int(obs.view(torch.int32).sum(dtype=torch.int64).item()), int(rew.view(torch.int32).sum(dtype=torch.int64).item()),
              int(dn.sum(dtype=torch.int64).item()), int(c.goal.view(torch.int32).sum(dtype=torch.int64).item())]
    for _ in range(2):
        c.step_many(a, obs, rew, dn)
    torch.cuda.synchronize()
    s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    K = 6
    s.record()
    for _ in range(K):
        c.step_many(a, obs, rew, dn)
    t.record(); torch.cuda.synchronize()
    ms = s.elapsed_time(t) / K
    res["%dx%d" % (n, T)] = {"frac_of_6544": round(73 * n * T / ms / 1e6 / 6544, 4), "env_steps_per_s": n * T / ms * 1e3, "digest": digest}
    c.close(); del c, e, a, obs, rew, dn
    torch.cuda.empty_cache()
print(json.dumps(res))
