"""Throughput of the generic (non-MSJ) fused step: a 6-joint / 14-tendon robot and a 15-joint / 64-tendon one, steady state.
Algorithmic bytes per env-step = 4A + 16J + 13 (read: action 4A, goal 4J, step word 4; written: obs 12J, reward 4, done 1,
step word 4).  usage: python tools/generic_bench.py [envs]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gym_roboy_b200.envs import RoboyEnv
from gym_roboy_b200.envs.robots import RoboyRobot
from gym_roboy_b200.envs.simulations import CudaSimulationClient
from gym_roboy_b200.spaces import Box


def robot(J, A, per_component=False):
    lo = -np.linspace(1.0, 3.0, J) if per_component else -2.5
    hi = np.linspace(0.5, 3.1, J) if per_component else 2.5

    class R(RoboyRobot):
        _A = Box(lo, hi, (J,), "float32") if not per_component else Box(lo, hi, dtype="float32")
        _V = Box(-0.6, 0.6, (J,), "float32")
        _T = Box(-0.2, 0.2, (A,), "float32")
        get_action_space = classmethod(lambda cls: cls._T)
        get_joint_angles_space = classmethod(lambda cls: cls._A)
        get_joint_vels_space = classmethod(lambda cls: cls._V)
    return R()


n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 22
out = {}
for name, J, A, pc in (("6_joints_14_tendons", 6, 14, False), ("15_joints_64_tendons_per_component", 15, 64, True), ("msj_forced_generic", 3, 8, False)):
    if name == "msj_forced_generic":
        os.environ["ROBOY_B200_FORCE_GENERIC"] = "1"
        c = CudaSimulationClient(num_envs=n, seed=1, device="cuda:0")
        del os.environ["ROBOY_B200_FORCE_GENERIC"]
    else:
        c = CudaSimulationClient(robot=robot(J, A, pc), num_envs=n, seed=1, device="cuda:0")
    e = RoboyEnv(c, strict=False); e.reset()
    c.set_step_num(((torch.arange(n, device="cuda:0") % 400) + 1).to(torch.int32))
    g = torch.Generator(device="cuda:0"); g.manual_seed(0)
    acts = [torch.rand((n, A), device="cuda:0", generator=g) * 2 - 1 for _ in range(2)]
    for i in range(5):
        c.step_fused(acts[i & 1])
    torch.cuda.synchronize()
    s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    K = 30
    s.record()
    for i in range(K):
        c.step_fused(acts[i & 1])
    t.record(); torch.cuda.synchronize()
    ms = s.elapsed_time(t) / K
    b = 4 * A + 16 * J + 13
    out[name] = {"envs": n, "ms_per_step": ms, "env_steps_per_s": n / ms * 1e3, "algorithmic_bytes_per_env_step": b,
                 "GBps": b * n / ms / 1e6, "frac_of_6544": b * n / ms / 1e6 / 6544, "msj_kernels": c.msj_kernels}
    c.close(); del c, e, acts
print(json.dumps(out, indent=1))
