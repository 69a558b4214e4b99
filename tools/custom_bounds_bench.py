"""Throughput of the tuned MSJ-shaped step (3 joints, 8 tendons, uniform bounds) for MSJ's own limits and for other limits
(which run the same kernel on IEEE division unless their spans are proved).  usage: python tools/custom_bounds_bench.py [envs]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from cuda_adaptor import robot_from_bounds
from gym_roboy_b200.envs import RoboyEnv
from gym_roboy_b200.envs.simulations import CudaSimulationClient

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 24
out = {}
for name, b in (("msj", None), ("msj_shaped_other_limits", dict(angle_low=-2.0, angle_high=2.0, vel_low=-0.7, vel_high=0.7, act_low=-0.5, act_high=0.5)),
                ("msj_shaped_one_sided", dict(angle_low=0.0, angle_high=2.5, vel_low=-0.3, vel_high=0.9, act_low=0.0, act_high=0.4))):
    for penalty in (False, True):
        c = CudaSimulationClient(robot=robot_from_bounds(b), num_envs=n, seed=1, device="cuda:0")
        e = RoboyEnv(c, joint_vel_penalty=penalty, strict=False); e.reset()
        c.set_step_num(((torch.arange(n, device="cuda:0") % 400) + 1).to(torch.int32))
        g = torch.Generator(device="cuda:0"); g.manual_seed(0)
        acts = [torch.rand((n, 8), device="cuda:0", generator=g) * 2 - 1 for _ in range(2)]
        for i in range(5):
            c.step_fused(acts[i & 1])
        torch.cuda.synchronize()
        s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        K = 40
        s.record()
        for i in range(K):
            c.step_fused(acts[i & 1])
        t.record(); torch.cuda.synchronize()
        ms = s.elapsed_time(t) / K
        out["%s%s" % (name, "+penalty" if penalty else "")] = {"ms_per_step": ms, "frac_of_6544": 93 * n / ms / 1e6 / 6544,
                                                              "msj_kernels": c.msj_kernels, "fast_division": c.fast_division}
        c.close(); del c, e, acts
print(json.dumps(out, indent=1))
