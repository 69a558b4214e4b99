"""What bounds roboy_step_host (bench.py's e2e) at 1/2/4/8 ranks?  Run under torch.distributed.run (or alone):

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/e2e_probe.py [envs_per_gpu]

Per configuration every rank runs the same thing at the same time (barrier first); the slowest rank's wall clock counts.
  copy probes   the H2D (32 B/env) and D2H (41 B/env) copies of roboy_step_host without the kernel: each direction
                alone, both at once, staged pattern and monolithic
  e2e           roboy_step_host itself: staged / mapped_out / mapped_all, stage sizes, stream counts, with and without
                binding the process (and so its page-locked buffers) to the GPU's NUMA node
Rank 0 prints one JSON document (and writes it to the path in $ROBOY_PROBE_OUT if set)."""
import json
import os
import subprocess
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from gym_roboy_b200 import _native
from gym_roboy_b200.envs import RoboyEnv
from gym_roboy_b200.envs.simulations import CudaSimulationClient

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
envs = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 24
torch.cuda.set_device(local)
numa = None
if os.environ.get("ROBOY_PROBE_NUMA", "1") == "1" and world > 1:
    from gym_roboy_b200.sharding import bind_to_gpu_numa_node
    numa = bind_to_gpu_numa_node(local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dev = torch.device("cuda", local)


def max_over_ranks(x):
    if world == 1:
        return x
    t = torch.tensor([x], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()


client = CudaSimulationClient(num_envs=envs, seed=1, env_id_base=rank * envs, device=dev)
env = RoboyEnv(client)
env.reset()
out = {"world": world, "envs_per_gpu": envs, "numa_bound_cpus": len(numa) if numa else None, "probes": {}, "e2e": []}
MODES = {"staged": _native.HOST_STAGED, "mapped_out": _native.HOST_MAPPED_OUT, "mapped_all": _native.HOST_MAPPED_ALL}
GB = 1e9


def run(write_combined):
    a, obs, rew, done = client.host_buffers(write_combined_actions=write_combined)
    a[...] = np.random.default_rng(rank).random((envs, 8), dtype=np.float32) * 2 - 1
    tag = "wc" if write_combined else "plain"
    for name, directions, mono, nbytes in (("h2d_alone", 1, True, 32), ("d2h_alone", 2, True, 41), ("both_monolithic", 3, True, 73),
                                           ("both_staged", 3, False, 73)):
        barrier()
        ms = max_over_ranks(client.copy_probe(a, obs, rew, done, directions=directions, monolithic=mono, iters=4))
        out["probes"]["%s_%s" % (name, tag)] = {"ms": ms, "GBps_per_gpu": nbytes * envs / (ms * 1e-3) / GB,
                                                "GBps_all": nbytes * envs * world / (ms * 1e-3) / GB,
                                                "env_steps_per_s_all": envs * world / (ms * 1e-3)}
    configs = [("staged", 1 << 19, 2, True, True), ("staged", 1 << 19, 2, False, True), ("staged", 1 << 19, 3, True, True),
               ("staged", 1 << 20, 2, True, True), ("staged", 1 << 21, 2, True, True), ("staged", 1 << 18, 2, True, True),
               ("staged", 1 << 22, 2, True, True), ("staged", 1 << 19, 3, True, False), ("staged", 1 << 19, 1, True, True),
               ("mapped_out", 1 << 19, 2, True, True), ("mapped_all", 0, 1, True, True)]
    for mode, stage, streams, ramp, ring in configs:
        client.set_host_mode(MODES[mode])
        if stage:
            client.set_host_pipeline(stage_envs=stage, n_streams=streams, ramp=ramp, ring=ring)
        try:
            for _ in range(2):
                client.step_host(a, obs, rew, done)
            barrier()
            t0 = time.perf_counter()
            k = 6
            for _ in range(k):
                client.step_host(a, obs, rew, done)
            dt = max_over_ranks(time.perf_counter() - t0)
            out["e2e"].append({"buffers": tag, "mode": mode, "stage_envs": stage, "streams": streams, "ramp": ramp, "ring": ring,
                               "ms_per_step": 1e3 * dt / k, "env_steps_per_s_all": envs * world * k / dt,
                               "GBps_all": 73 * envs * world * k / dt / GB})
        except Exception as exc:
            out["e2e"].append({"buffers": tag, "mode": mode, "error": str(exc)[:200]})
    client.set_host_mode(_native.HOST_STAGED)
    client.set_host_pipeline()


run(False)
if os.environ.get("ROBOY_PROBE_WC"):
    run(True)
if rank == 0:
    def sh(cmd):
        try:
            return subprocess.run(cmd, shell=True, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=30).stdout[-4000:]
        except Exception as exc:
            return str(exc)
    out["system"] = {"topo": sh("nvidia-smi topo -m"), "lscpu": sh("lscpu | head -25"),
                     "pcie": sh("nvidia-smi --query-gpu=index,pci.bus_id,pcie.link.gen.current,pcie.link.width.current --format=csv"),
                     "numa_nodes": sh("cat /sys/devices/system/node/online; for d in /sys/bus/pci/devices/*; do "
                                      "c=$(cat $d/class); if [ \"$c\" = 0x030200 ]; then echo $d $(cat $d/numa_node); fi; done"),
                     "meminfo": sh("head -3 /proc/meminfo")}
    best = max((r for r in out["e2e"] if "env_steps_per_s_all" in r), key=lambda r: r["env_steps_per_s_all"])
    out["best_e2e"] = best
    txt = json.dumps(out, indent=1)
    print(txt)
    if os.environ.get("ROBOY_PROBE_OUT"):
        open(os.environ["ROBOY_PROBE_OUT"], "w").write(txt)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
