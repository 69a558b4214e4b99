# time the fused env step for the four (joint_vel_penalty, bonus) flag combinations at an HBM-bound size
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gym_roboy_b200.envs import RoboyEnv
from gym_roboy_b200.envs.simulations import CudaSimulationClient
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 24
g = torch.Generator(device="cuda:0"); g.manual_seed(0)
acts = [torch.rand((n, 8), device="cuda:0", generator=g) * 2 - 1 for _ in range(2)]
for penalty in (False, True):
    for bonus in (True, False):
        c = CudaSimulationClient(num_envs=n, seed=1234, device="cuda:0")
        e = RoboyEnv(c, joint_vel_penalty=penalty, is_agent_getting_bonus_for_reaching_goal=bonus, strict=False)
        e.reset()
        for i in range(5): e.step(acts[i & 1])
        torch.cuda.synchronize()
        s, f = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for i in range(40): e.step(acts[i & 1])
        f.record(); torch.cuda.synchronize()
        ms = s.elapsed_time(f) / 40
        print("penalty=%-5s bonus=%-5s %.4f ms/step  %.3e env-steps/s  %.0f GB/s (%.1f%% of 6544)  violations=%d" % (
            penalty, bonus, ms, n / ms * 1e3, 93 * n / ms / 1e6, 93 * n / ms / 1e6 / 65.44, c.stats()["violations"]), flush=True)
        c.close()
