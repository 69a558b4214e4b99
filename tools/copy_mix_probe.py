"""What does the memory system give a streaming kernel with the step kernels' read : write mixes?  torch elementwise
kernels as neutral probes (they are not the product): copy (1:1, what MEASURED_PEAKS.json's hbm_gbs is), a write-only fill,
a read-heavy reduction-free op (2 reads : 1 write) and a write-heavy one (1 read : 2 writes via two outputs).
usage: python tools/copy_mix_probe.py"""
import json, torch
dev = "cuda:0"
n = 1 << 28   # 1 GiB per float32 tensor
a = torch.empty(n, device=dev); b = torch.empty(n, device=dev); c = torch.empty(n, device=dev)
a.normal_(); b.normal_()


def timed(fn, nbytes, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); t.record(); torch.cuda.synchronize()
        best = min(best, s.elapsed_time(t))
    return nbytes / best / 1e6


out = {
    "copy_1r_1w_GBps": timed(lambda: c.copy_(a), 8 * n),
    "fill_0r_1w_GBps": timed(lambda: c.fill_(1.0), 4 * n),
    "add_2r_1w_GBps": timed(lambda: torch.add(a, b, out=c), 12 * n),
    "sum_1r_0w_GBps": timed(lambda: a.sum(), 4 * n),
}
print(json.dumps(out))
