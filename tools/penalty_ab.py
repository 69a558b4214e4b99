"""A/B of the velocity-penalty evaluation on one box: float32 with the float64 re-run near reward_range's bounds
(PenaltyF32, msj_math.cuh; the default) against the float64 expression everywhere (ROBOY_B200_PENALTY_F64=1), steady
state, joint_vel_penalty=True.  Rounds alternate so clock drift hits both alike.  usage: python tools/penalty_ab.py [rounds]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gym_roboy_b200.envs import RoboyEnv
from gym_roboy_b200.envs.robots import RoboyRobot
from gym_roboy_b200.envs.simulations import CudaSimulationClient
from gym_roboy_b200.spaces import Box


def robot(J, A, a, v, t):
    class R(RoboyRobot):
        _A = Box(-a, a, (J,), "float32")
        _V = Box(-v, v, (J,), "float32")
        _T = Box(-t, t, (A,), "float32")
        get_action_space = classmethod(lambda cls: cls._T)
        get_joint_angles_space = classmethod(lambda cls: cls._A)
        get_joint_vels_space = classmethod(lambda cls: cls._V)
    return R()


CASES = (("msj", None, 3, 8, 1 << 24), ("msj_shaped_other_limits", robot(3, 8, 2.0, 0.7, 0.4), 3, 8, 1 << 24),
         ("6_joints_14_tendons", robot(6, 14, 2.5, 0.6, 0.2), 6, 14, 1 << 22))
rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 2
out = {}
for name, rob, J, A, n in CASES:
    g = torch.Generator(device="cuda:0"); g.manual_seed(0)
    acts = [torch.rand((n, A), device="cuda:0", generator=g) * 2 - 1 for _ in range(2)]
    b = 4 * A + 16 * J + 13
    res = {"float32": [], "float64": []}
    for r in range(rounds):
        for mode in ("float32", "float64"):
            if mode == "float64":
                os.environ["ROBOY_B200_PENALTY_F64"] = "1"
            try:
                kw = {"robot": rob} if rob is not None else {}
                c = CudaSimulationClient(num_envs=n, seed=1234, device="cuda:0", **kw)
                e = RoboyEnv(c, joint_vel_penalty=True, strict=False)
            finally:
                os.environ.pop("ROBOY_B200_PENALTY_F64", None)
            assert c.penalty_float32 == (mode == "float32")
            e.reset()
            c.set_step_num(((torch.arange(n, device="cuda:0") % 400) + 1).to(torch.int32))
            for i in range(6):
                c.step_fused(acts[i & 1])
            torch.cuda.synchronize()
            s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            K = 40
            s.record()
            for i in range(K):
                c.step_fused(acts[i & 1])
            t.record(); torch.cuda.synchronize()
            ms = s.elapsed_time(t) / K
            res[mode].append(round(b * n / ms / 1e6 / 6544, 4))
            viol = c.stats()["violations"]
            c.close(); del c, e
    out[name] = {"envs": n, "algorithmic_bytes_per_env_step": b, "frac_of_6544_GBps": res, "violations_last_run": viol}
    del acts
print(json.dumps(out, indent=1))
