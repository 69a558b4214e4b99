#!/usr/bin/env python
"""bench.py -- MSJ env-steps/sec of the fused batched env step on B200 (contract: see DESIGN.md §Measurement).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--envs E] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the hot path (one launch of the fused step kernel: RoboyEnv.step +
StubSimulationClient update + reward + done + auto-reset) over E envs PER GPU (weak scaling),
with a fresh synthetic action batch each step.  Rank 0 prints ONE JSON line.

  value      env-steps/s, device-timed (CUDA events on the launching stream, max over ranks),
             actions already resident in HBM.
  roofline   algorithmic bytes (93 B per env-step, closed loop -- SURVEY.md 8d) / mean kernel
             duration measured per launch with CUDA events, against MEASURED_PEAKS.json hbm_gbs.
  e2e        the same step through the C-ABI host-buffer call (roboy_step_host): pinned host
             actions -> H2D -> kernel -> D2H of obs, reward, done, all inside the timed region.
  cpu_baseline  the CPU oracle (a C port of the reference's algorithm; oracle/) on all host
             cores, on a bounded sample -- a reported baseline, not the product.
  --impl reference   times that CPU port alone, same metric and config.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

BYTES_PER_ENV_STEP = 93            # reads 48 (action 32 + goal 12 + step 4) + writes 45 (obs 36 + reward 4 + done 1 + step 4)
H2D_BYTES, D2H_BYTES = 32, 41      # per env-step through host buffers
DEFAULT_ENVS = 1 << 24             # 16,777,216 envs per GPU: 1.56 GB per step, far beyond the 126 MB L2
SWEEP = (4096, 32768, 262144, 1048576, 4194304, 16777216)
METRIC, UNIT = "MSJ env-steps/sec (device-timed)", "env-steps/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--envs", type=int, default=DEFAULT_ENVS, help="envs per GPU")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-sweep", action="store_true")
    ap.add_argument("--stats-every", type=int, default=100, help="steps between the NCCL episode-stat reductions (N > 1)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=0, help="0: min(steps, 12)")
    return ap.parse_args()


def workload_config(envs, n_gpus):
    return {
        "workload": "MsjRobot batched env, fused step+reward+done+auto-reset, {:,} envs per B200 "
                    "(BASELINE.json configs[2] HBM-roofline sweep, HBM-bound point; x{} GPUs = configs[3]); "
                    "configs[1] (4,096 envs) and the other sweep sizes are in `sweep`".format(envs, n_gpus),
        "envs_per_gpu": envs, "total_envs": envs * n_gpus, "robot": "MsjRobot", "client": "in-process stub semantics",
        "loop": "closed-loop, one launch per step, 93 B/env-step algorithmic",
        "actions": "U(-1,1) float32 [N,8], torch.Generator(seed=0), 2 rotating batches",
        "l2": "inputs+outputs per step ({:.2f} GB) exceed the 126 MB L2; no explicit flush".format(envs * BYTES_PER_ENV_STEP / 1e9),
        "parallelism": "env shards, dp{}".format(n_gpus), "stats_allreduce_every": 100,
    }


class stdout_to_stderr:
    """Temporarily point file descriptor 1 at stderr (for native libraries that print to stdout)."""

    def __enter__(self):
        sys.stdout.flush()
        self._saved = os.dup(1)
        os.dup2(2, 1)

    def __exit__(self, *a):
        sys.stdout.flush()
        os.dup2(self._saved, 1)
        os.close(self._saved)


# --------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """Samples SM clock and throttle reasons via NVML while the timed region runs."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
                 "hw_power_brake": 0x80}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.002)

    def __enter__(self):
        if self.nv:
            self._thread = threading.Thread(target=self._loop, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thread:
            self._thread.join()

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# --------------------------------------------------------------------------------------------- CPU port
def cpu_port_rate(envs, steps, warmup, threads):
    """env-steps/s of the oracle (C port of the reference algorithm) on `threads` host threads."""
    import numpy as np
    from oracle import oracle as orc
    env = orc.OracleEnv(envs, seed=0, threads=threads)
    env.reset()
    rng = np.random.default_rng(0)
    acts = [rng.uniform(-1, 1, (envs, 8)).astype(np.float32) for _ in range(2)]
    for i in range(warmup):
        env.step(acts[i & 1])
    t0 = time.perf_counter()
    for i in range(steps):
        env.step(acts[i & 1])
    dt = time.perf_counter() - t0
    return envs * steps / dt, dt


def reference_python_context():
    """The unmodified Python reference timed in the build container (oracle/time_reference.py); context
    only -- it is not measured on this box and is not the `value` of any baseline object."""
    path = os.path.join(ROOT, "profiles", "r1_reference_python_cpu.json")
    if not os.path.exists(path):
        return None
    j = json.load(open(path))
    return {"where": j["where"], "cores": j["cores"],
            "single_env_steps_per_s": j["single_env_steps_per_s"]["median"],
            "vec_env_restated_steps_per_s": j["vec_env_restated_steps_per_s"]["value"],
            "independent_processes_steps_per_s": j["independent_processes_steps_per_s"]["value"]}


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = host_threads()
    sample = min(args.envs, 1 << 21)     # bounded sample of the workload: 2,097,152 envs per step
    # keep the whole run within a couple of minutes whatever the core count
    rate, dt = cpu_port_rate(sample, max(1, args.steps), max(0, args.warmup), threads)
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / max(1, args.steps), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.envs, args.gpus),
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": "{:,} envs x {} steps per timed run, oracle/roboy_oracle.c (C port of the "
                                   "reference algorithm; the Python reference itself cannot travel to this box), "
                                   "pthreads over all host cores".format(sample, args.steps),
                         "reference_python": reference_python_context()},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------- GPU arm
def device_timed(env, client, actions, steps, warmup, torch, dist, world, stats_every=100, per_step_events=True):
    """K steps, barrier + synchronize on both sides, CUDA events on the launching stream.
    Returns (elapsed_ms_total_max_over_ranks, mean_kernel_ms, median_kernel_ms, launches).  per_step_events=False
    records only the first and the last event (host-bound sizes: an event record per step costs more than the step)."""
    from gym_roboy_b200.sharding import all_reduce_stats
    for i in range(warmup):
        env.step(actions[i % len(actions)])
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    launches0 = client.launch_count()
    ev[0].record()
    for i in range(steps):
        env.step(actions[i % len(actions)])
        if per_step_events or i == steps - 1:
            ev[i + 1].record()
        if world > 1 and (i + 1) % stats_every == 0:
            # the tiny episode-stat reduction (64 bytes), stream-ordered between two steps: ~20 us per 100 steps.  On a side
            # stream the NCCL kernel competes for SM slots with step grids that programmatic dependent launch has already
            # queued, which measured 10 % slower at 2 and 8 GPUs.
            all_reduce_stats(client.stats_tensor)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
    total_ms = ev[0].elapsed_time(ev[steps])
    per = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(steps)) if per_step_events else [total_ms / steps]
    mean_kernel_ms = sum(per) / len(per)
    if world > 1:
        t = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    return total_ms, mean_kernel_ms, per[len(per) // 2], client.launch_count() - launches0


def run_b200_arm(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback "
                         "(use --impl reference for the CPU port)")
    torch.cuda.set_device(local)
    numa_cpus = None
    if world > 1 and not os.environ.get("ROBOY_BENCH_NO_NUMA_BIND"):
        from gym_roboy_b200.sharding import bind_to_gpu_numa_node
        numa_cpus = bind_to_gpu_numa_node(local)   # pinned host buffers on the GPU's NUMA node
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        with stdout_to_stderr():   # NCCL prints its version banner on stdout; rank 0 must print ONE JSON line
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            dist.barrier()
            torch.cuda.synchronize()
    from gym_roboy_b200.envs import RoboyEnv
    from gym_roboy_b200.envs.simulations import CudaSimulationClient
    from gym_roboy_b200.sharding import all_reduce_stats, shard_range, summarize

    dev = torch.device("cuda", local)
    envs = args.envs
    begin, end = shard_range(envs * world, world, rank)

    def make(n, base=0):
        client = CudaSimulationClient(num_envs=n, seed=1234, env_id_base=base, device=dev)
        env = RoboyEnv(client)
        env.reset()
        return env, client

    def make_actions(n, k=2):
        gen = torch.Generator(device=dev)
        gen.manual_seed(rank)            # seed 0 on rank 0
        return [torch.rand((n, 8), device=dev, generator=gen) * 2 - 1 for _ in range(k)]

    env, client = make(end - begin, begin)
    actions = make_actions(end - begin)
    # The timed region is ONE CUDA-event pair around the K back-to-back launches: an event recorded between two
    # launches would keep the next step kernel from starting under the tail of the previous one (programmatic dependent
    # launch).  The average launch duration of the roofline is that region divided by its K launches; a second, untimed
    # pass with an event per launch gives the per-launch median as a diagnostic.
    with ClockSampler(local) as clocks:
        total_ms, mean_kernel_ms, _, launches = device_timed(
            env, client, actions, args.steps, args.warmup, torch, dist, world, stats_every=args.stats_every,
            per_step_events=False)
    value = envs * world * args.steps / (total_ms * 1e-3)
    _, per_launch_mean_ms, median_kernel_ms, _ = device_timed(env, client, actions, min(args.steps, 50), 3, torch, dist, world)

    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (of measured)"
    else:
        peak, peak_src = 6650.0, "B200_PROFILING.md fallback (of fallback)"
    achieved = BYTES_PER_ENV_STEP * envs / (mean_kernel_ms * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")   # dram bytes per launch from the committed ncu capture
    if os.path.exists(tpath):
        tj = json.load(open(tpath))
        if int(tj.get("envs", 0)) == envs:
            traffic = tj.get("dram_bytes_per_launch")
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_src, "kernel": "roboy::step_kernel<false,true,true>",
                "algorithmic_bytes_per_launch": BYTES_PER_ENV_STEP * envs, "kernel_ms_mean": mean_kernel_ms,
                "timing": "one CUDA-event pair around the {} timed launches on the launching stream; mean = region / launches".format(args.steps),
                "per_launch_events_pass": {"kernel_ms_mean": per_launch_mean_ms, "kernel_ms_median": median_kernel_ms,
                                           "launches": min(args.steps, 50)}}
    stats = summarize(all_reduce_stats(client.stats_tensor))

    # ---- end to end through host buffers (C-ABI roboy_step_host) ----
    e2e = None
    if not args.no_e2e:
        n = end - begin
        k2 = args.e2e_steps or min(args.steps, 12)
        a_host = [torch.empty((n, 8), dtype=torch.float32).pin_memory() for _ in range(2)]
        for a, src in zip(a_host, actions):
            a.copy_(src)
        obs_h = torch.empty((n, 9), dtype=torch.float32).pin_memory()
        rew_h = torch.empty(n, dtype=torch.float32).pin_memory()
        done_h = torch.empty(n, dtype=torch.uint8).pin_memory()
        bufs = ([a.numpy() for a in a_host], obs_h.numpy(), rew_h.numpy(), done_h.numpy())
        for i in range(3):
            client.step_host(bufs[0][i & 1], bufs[1], bufs[2], bufs[3])
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        l0 = client.launch_count()
        t0 = time.perf_counter()
        for i in range(k2):
            client.step_host(bufs[0][i & 1], bufs[1], bufs[2], bufs[3])   # returns with outputs in host memory
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        e2e = {"value": envs * world * k2 / dt, "unit": UNIT, "h2d_bytes_per_step": H2D_BYTES * n,
               "d2h_bytes_per_step": D2H_BYTES * n, "steps": k2, "ms_per_step": 1e3 * dt / k2,
               "api": "roboy_step_host (C-ABI): pinned host actions -> H2D -> step kernel -> D2H obs+reward+done, "
                      "pipelined over 2 streams in 524,288-env stages", "gpu_launches": client.launch_count() - l0,
               "checksum": float(rew_h[:1024].double().sum()),
               "numa_bound_cpus": len(numa_cpus) if numa_cpus else None}
        del a_host, obs_h, rew_h, done_h, bufs
    del env, client, actions
    torch.cuda.empty_cache()

    # ---- size sweep (rank 0, N = 1 only): SURVEY.md 8d config C3 ----
    sweep = None
    if world == 1 and not args.no_sweep:
        sweep = []
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # 256 MiB > L2
        for n in SWEEP:
            e, c = make(n)
            acts = make_actions(n)
            k = max(20, min(args.steps, 200))
            # (below ~1M envs the eager loop is bound by the host, where a CUDA event per step would dominate)
            tot, mean_ms, med_ms, _ = device_timed(e, c, acts, k, max(3, args.warmup), torch, dist, 1,
                                                   per_step_events=n > 1048576)
            row = {"envs": n, "env_steps_per_s": n * k / (tot * 1e-3), "ms_per_step": tot / k,
                   "GBps_algorithmic": BYTES_PER_ENV_STEP * n / (mean_ms * 1e-3) / 1e9,
                   "frac_of_peak": BYTES_PER_ENV_STEP * n / (mean_ms * 1e-3) / 1e9 / peak,
                   "regime": "launch-bound" if n <= 32768 else ("L2-resident" if BYTES_PER_ENV_STEP * n < 126e6 else "HBM-bound")}
            if n <= 4194304:   # launch/host-bound sizes: the same steps captured in one CUDA graph and replayed
                kg = 100
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    for i in range(kg):
                        c.step_fused(acts[i & 1])
                for _ in range(2):
                    g.replay()
                torch.cuda.synchronize()
                s_, e_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s_.record()
                for _ in range(3):
                    g.replay()
                e_.record()
                torch.cuda.synchronize()
                gms = s_.elapsed_time(e_) / (3 * kg)
                row["cuda_graph_ms_per_step"] = gms
                row["cuda_graph_env_steps_per_s"] = n / (gms * 1e-3)
                row["cuda_graph_GBps_algorithmic"] = BYTES_PER_ENV_STEP * n / (gms * 1e-3) / 1e9
                del g
            if BYTES_PER_ENV_STEP * n < 2 * 126e6:     # fits (mostly) in L2: also time with an explicit L2 flush per step
                times = []
                for i in range(12):
                    flush.fill_(i)
                    s_, e_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    s_.record(); e.step(acts[i & 1]); e_.record()
                    torch.cuda.synchronize()
                    times.append(s_.elapsed_time(e_))
                times = sorted(times[2:])
                row["l2_flushed_ms"] = times[len(times) // 2]
                row["l2_flushed_GBps_algorithmic"] = BYTES_PER_ENV_STEP * n / (row["l2_flushed_ms"] * 1e-3) / 1e9
            sweep.append(row)
            del e, c, acts
        del flush

    # ---- open-loop fused T-step variant (73 B per env-step, state stays in registers) ----
    open_loop = None
    if world == 1 and not args.no_sweep:
        n, T = 1 << 22, 16     # 4,194,304 envs x 16 pre-recorded steps: 4.9 GB per launch
        e, c = make(n)
        gen = torch.Generator(device=dev); gen.manual_seed(1)
        a = torch.rand((T, n, 8), device=dev, generator=gen) * 2 - 1
        obs = torch.empty((T, n, 9), device=dev); rew = torch.empty((T, n), device=dev)
        dn = torch.empty((T, n), dtype=torch.uint8, device=dev)
        for _ in range(3):
            c.step_many(a, obs, rew, dn)
        torch.cuda.synchronize()
        s_, e_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 5
        s_.record()
        for _ in range(reps):
            c.step_many(a, obs, rew, dn)
        e_.record()
        torch.cuda.synchronize()
        ms = s_.elapsed_time(e_) / reps
        open_loop = {"envs": n, "T": T, "ms_per_launch": ms, "env_steps_per_s": n * T / (ms * 1e-3),
                     "algorithmic_bytes_per_env_step": 73, "GBps_algorithmic": 73 * n * T / (ms * 1e-3) / 1e9,
                     "frac_of_peak": 73 * n * T / (ms * 1e-3) / 1e9 / peak,
                     "api": "roboy_step_many: pre-recorded actions [T,N,8], one launch, bit-identical to T roboy_step calls"}
        del e, c, a, obs, rew, dn
        torch.cuda.empty_cache()

    # ---- closed-loop policy rollouts (BASELINE configs[1] size and configs[4] shape), every N ----
    rollout = None
    if not args.no_sweep:
        from gym_roboy_b200.rollout import MlpPolicy, RolloutCollector
        rollout = []
        fp32_peak = 148 * 128 * 2 * 1.965e9 / 1e12   # nominal float32 FMA peak of one B200 at max clock, TFLOP/s
        mufu_peak = 148 * 16 * 1.965e9               # MUFU results per second of one B200 at max clock (16 lanes per SM)
        for n, mode in ((4096, "torch"), (4096, "torch+graph"), (4096, "fused"), (4096, "fused_tc_exact"), (4096, "fused_tc"),
                        (262144, "torch"), (262144, "torch+graph"), (262144, "fused"), (262144, "fused_tc_exact"),
                        (262144, "fused_tc"), (1048576, "fused"), (1048576, "fused_tc_exact"), (1048576, "fused_tc")):
            torch.manual_seed(0)
            b0, _ = shard_range(n * world, world, rank)
            e, c = make(n, b0)
            fused = mode == "fused"
            col = RolloutCollector(e, MlpPolicy().to(dev), n_steps=128, fused={"fused": "fp32", "fused_tc": "tc", "fused_tc_exact": "tc_exact"}.get(mode, False))
            if mode == "torch+graph":
                col.capture()
            for _ in range(2):
                col.collect()
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
                torch.cuda.synchronize()
            s_, e_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 5
            s_.record()
            for _ in range(reps):
                col.collect()
            e_.record()
            torch.cuda.synchronize()
            ms = s_.elapsed_time(e_) / reps
            if world > 1:
                tms = torch.tensor([ms], dtype=torch.float64, device="cuda")
                dist.all_reduce(tms, op=dist.ReduceOp.MAX)
                ms = float(tms.item())
            row = {"envs_per_gpu": n, "n_gpus": world, "n_steps": 128, "mode": mode, "cuda_graph": mode == "torch+graph",
                   "ms_per_rollout": ms, "env_steps_per_s": n * world * 128 / (ms * 1e-3)}
            if fused:
                # 2 networks x (9*64 + 64*64 + 64*8) FMA per env-step (the value head is padded to 8 columns)
                tflops = 2 * 2 * (9 * 64 + 64 * 64 + 64 * 8) * n * 128 / (ms * 1e-3) / 1e12
                row.update({"what": "roboy_policy_rollout: both MlpPolicy networks (float32 FFMA2), Philox Gaussian sample, clip "
                                    "and env step for all 128 steps in ONE launch, state in registers; + GAE kernel",
                            "bound": "fp32 FMA", "tflops_fp32": tflops, "frac_of_nominal_fp32_peak": tflops / fp32_peak,
                            "nominal_fp32_peak_tflops": fp32_peak})
            elif mode == "fused_tc_exact":
                row.update({"what": "roboy_policy_rollout_tc(exact=1): tensor cores with every operand split into two float16 "
                                    "halves (three MMAs per product) and the accurate tanh; agrees with the float32 policy to "
                                    "~1e-6 like `fused`; + GAE kernel",
                            "bound": "MUFU (tanh = ex2 + rcp)", "mufu_per_env_step": 539,
                            "frac_of_nominal_mufu_peak": 539 * n * 128 / (ms * 1e-3) / mufu_peak})
            elif mode == "fused_tc":
                # the MUFU unit bounds this kernel: 256 tanh per env-step (+ ~27 for the Gaussian noise and the env's exp / sqrt)
                row.update({"what": "roboy_policy_rollout_tc: as `fused`, with the matrix products on the tensor cores (tcgen05.mma "
                                    "kind::f16, float32 accumulators in TMEM, four 128-env tiles per SM); agrees with the float32 "
                                    "policy to ~1e-3; + GAE kernel",
                            "bound": "MUFU (tanh)", "mufu_per_env_step": 283,
                            "frac_of_nominal_mufu_peak": 283 * n * 128 / (ms * 1e-3) / mufu_peak,
                            "hbm_GBps_written": 81 * n * 128 / (ms * 1e-3) / 1e9})
            else:
                row["what"] = ("torch MlpPolicy 9-64-64-8 (+value net) forward, Gaussian sample, device clip, fused env "
                               "step writing into [T,N] buffers, GAE kernel; policy replicated per GPU")
            rollout.append(row)
            del col, e, c
            torch.cuda.empty_cache()

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = host_threads()
        sample = 1 << 21
        rate1, dt1 = cpu_port_rate(sample, 2, 1, threads)              # calibrate
        k = max(2, min(200, int(8.0 / max(dt1 / 2, 1e-4))))            # ~8 s of wall clock on all cores
        rate, dt = cpu_port_rate(sample, k, 1, threads)
        cpu_baseline = {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
                        "sample": "{:,} envs x {} steps ({:.1f} s wall on {} threads), oracle/roboy_oracle.c".format(
                            sample, k, dt, threads),
                        "reference_python": reference_python_context()}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": workload_config(envs, world), "roofline": roofline,
            "cpu_baseline": cpu_baseline, "e2e": e2e, "gpu_launches": launches, "clocks": clocks.summary(),
            "episode_stats": stats, "sweep": sweep, "open_loop": open_loop, "rollout": rollout, "impl": "b200",
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_b200_arm(args)


if __name__ == "__main__":
    main()
