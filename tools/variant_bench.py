"""Kernel-experiment helper: time the fused step for one build of the library.
usage: ROBOY_B200_LIB=path python tools/variant_bench.py [envs] [steps]   -> prints one line"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gym_roboy_b200.envs import RoboyEnv
from gym_roboy_b200.envs.simulations import CudaSimulationClient
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 24
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 40
# parity spot check against the oracle first (4096 envs, 3 steps)
from oracle import oracle as orc
c = CudaSimulationClient(num_envs=4099, seed=7, device="cuda:0"); e = RoboyEnv(c); o = orc.OracleEnv(4099, seed=7)
e.reset(); o.reset(); rng = np.random.default_rng(0); ok = True
st = (np.arange(4099) % 400 + 1).astype(np.int32); c.set_step_num(st); o.step_flags[:] = (o.step_flags & ~np.uint32(orc.STEP_MASK)) | st.astype(np.uint32)
for t in range(3):
    a = rng.uniform(-1, 1, (4099, 8)).astype(np.float32); a[rng.random(4099) < 0.02] = 0
    ob, rw, dn, _ = e.step(torch.as_tensor(a, device="cuda:0")); oo, orw, od = o.step(a)
    ok &= np.array_equal(ob.cpu().numpy(), oo) and np.array_equal(dn.cpu().numpy(), od) and \
        (np.abs(rw.cpu().numpy() - orw) / np.abs(orw)).max() < 1e-6
c = CudaSimulationClient(num_envs=n, seed=1234, device="cuda:0"); e = RoboyEnv(c); e.reset()
g = torch.Generator(device="cuda:0"); g.manual_seed(0)
acts = [torch.rand((n, 8), device="cuda:0", generator=g) * 2 - 1 for _ in range(2)]
for i in range(5): e.step(acts[i & 1])
torch.cuda.synchronize()
s, f = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for i in range(steps): e.step(acts[i & 1])
f.record(); torch.cuda.synchronize()
ms = s.elapsed_time(f) / steps
print("%-28s parity=%s  %.4f ms/step  %.3e env-steps/s  %.0f GB/s (%.1f%% of 6544)  grid=%s" % (
    os.path.basename(os.environ.get("ROBOY_B200_LIB", "default")), ok, ms, n / ms * 1e3, 93 * n / ms / 1e6,
    93 * n / ms / 1e6 / 65.44, c.step_geometry()["grid"]))
