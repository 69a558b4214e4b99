import sys, time, ctypes, torch
sys.path.insert(0, '.')
from gym_roboy_b200 import _native
from gym_roboy_b200.envs import RoboyEnv
from gym_roboy_b200.envs.simulations import CudaSimulationClient
n = 4096
c = CudaSimulationClient(num_envs=n, seed=1, device='cuda:0'); e = RoboyEnv(c); e.reset()
a = torch.rand((n, 8), device='cuda:0') * 2 - 1
def timeit(f, k=20000):
    for _ in range(200): f()
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(k): f()
    torch.cuda.synchronize(); return (time.perf_counter() - t) / k * 1e6
print('env.step            %.2f us' % timeit(lambda: e.step(a)))
print('client.step_fused   %.2f us' % timeit(lambda: c.step_fused(a)))
L = _native.load(); h = c._h
pa = ctypes.c_void_p(a.data_ptr()); st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
print('raw ctypes call     %.2f us' % timeit(lambda: L.roboy_step(h, pa, None, None, None, st)))
print('stream lookup       %.2f us' % timeit(lambda: torch.cuda.current_stream(c.device).cuda_stream, 100000))
print('data_ptr + c_void_p %.2f us' % timeit(lambda: ctypes.c_void_p(a.data_ptr()), 100000))
print('a.to(...).reshape.contiguous %.2f us' % timeit(lambda: a.to(device=c.device, dtype=torch.float32).reshape(n, -1).contiguous(), 100000))
raw = getattr(torch._C, "_cuda_getCurrentRawStream", None)
if raw is not None:
    print('raw stream lookup   %.2f us' % timeit(lambda: raw(0), 100000))
print('null_step           %.2f us' % timeit(lambda: c.null_step()))
