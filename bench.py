#!/usr/bin/env python
"""bench.py -- MSJ env-steps/sec of the fused batched env step on B200 (contract: see DESIGN.md §Measurement).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--envs E] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the hot path (one launch of the fused step kernel: RoboyEnv.step +
StubSimulationClient update + reward + done + auto-reset) over E envs PER GPU (weak scaling),
with a fresh synthetic action batch each step.  Rank 0 prints ONE JSON line.

  value      env-steps/s, device-timed (CUDA events on the launching stream, max over ranks),
             actions already resident in HBM.
  roofline   algorithmic bytes (93 B per env-step, closed loop -- SURVEY.md 8d) / average launch duration =
             the CUDA-event-timed region / its K launches (one event pair: an event between two launches would
             defeat the programmatic dependent launch), against MEASURED_PEAKS.json hbm_gbs.
  parity_gate  before anything is timed, the benchmarked env shard itself is stepped GATE_STEPS times and the first
             GATE_ENVS envs are compared with the CPU oracle on the same actions (obs / done / goal / step word
             bit-exact, rewards 1e-6 relative).  A failed gate prints no `value`.  Both arms draw the same actions for
             that prefix and print `parity_checksum` over it, so the two records can be compared with each other.
  e2e        the same step through the C-ABI host-buffer call (roboy_step_host): page-locked host
             actions -> H2D -> kernel -> D2H of obs, reward, done, all inside the timed region; `copy_ceiling` is the
             same copies without the kernel (roboy_host_copy_probe), `frac_of_ceiling` = e2e / that.
  cpu_baseline  the CPU oracle (a C port of the reference's algorithm; oracle/) on all host
             cores, on a bounded sample -- a reported baseline, not the product; `reference_python` is the unmodified
             Python reference (baseline/_ref) timed on THIS box (oracle/time_reference.py --quick).
  --impl reference   times that CPU port alone, same metric and config, same actions for the gate prefix.
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
SEED = 1234                        # Philox key of the benchmarked population (both arms)
GATE_ENVS, GATE_STEPS = 1 << 21, 3  # the parity gate: the first 2,097,152 envs of the benchmarked shard, 3 steps
EPISODE_LEN = 400


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
        "actions": "U(-1,1) float32 [N,8], 2 rotating batches: the first {:,} envs from numpy default_rng(20261018) "
                   "(identical in the reference arm), the rest from torch.Generator(seed=rank)".format(min(envs, GATE_ENVS)),
        "episode_phases": "step_num = global_env_id % 400 + 1 after reset: 1/400 of the envs end an episode in every step",
        "l2": "inputs+outputs per step ({:.2f} GB) exceed the 126 MB L2; no explicit flush".format(envs * BYTES_PER_ENV_STEP / 1e9),
        "parallelism": "env shards, dp{}".format(n_gpus),
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


# --------------------------------------------------------------------------------------------- shared inputs, checksums
def gate_action_batches(n):
    """The two rotating action batches of the first `n` (<= GATE_ENVS) envs -- ONE generator for both arms."""
    import numpy as np
    rng = np.random.default_rng(20261018)
    out = []
    for _ in range(2):
        a = rng.random((GATE_ENVS, 8), dtype=np.float32) * np.float32(2) - np.float32(1)
        out.append(np.ascontiguousarray(a[:n]))
    return out


def episode_phases(begin, n):
    """step_num after the initial reset: spread over the episode so that every step ends 1/400 of the episodes."""
    import numpy as np
    return ((np.arange(begin, begin + n, dtype=np.int64) % EPISODE_LEN) + 1).astype(np.int32)


class Checksum:
    """Order-independent digests of the gate prefix's outputs; printed by BOTH arms (`parity_checksum`)."""

    def __init__(self):
        self.done_count, self.obs_xor, self.obs_sum, self.reward_sum, self.steps = 0, 0, 0, 0.0, 0

    def add(self, obs, reward, done):
        import numpy as np
        w = np.ascontiguousarray(obs, np.float32).view(np.uint32)
        self.obs_xor ^= int(np.bitwise_xor.reduce(w, axis=None))
        self.obs_sum = (self.obs_sum + int(w.sum(dtype=np.uint64))) & (2 ** 64 - 1)
        self.done_count += int(np.count_nonzero(done))
        self.reward_sum += float(np.asarray(reward, np.float64).sum())
        self.steps += 1

    def result(self, envs, goal, step_words):
        import numpy as np
        g = np.ascontiguousarray(goal, np.float32).view(np.uint32)
        sw = np.ascontiguousarray(step_words).astype(np.uint32)
        return {"envs": envs, "steps": self.steps, "done_count": self.done_count, "obs_xor32": "%08x" % self.obs_xor,
                "obs_sum64": "%016x" % self.obs_sum, "goal_xor32": "%08x" % int(np.bitwise_xor.reduce(g, axis=None)),
                "step_word_xor32": "%08x" % int(np.bitwise_xor.reduce(sw, axis=None)),
                "reward_sum": self.reward_sum,
                "note": "bit patterns are exact digests (equal across arms); reward_sum differs in the last digits only "
                        "(expf is not the same function on every platform -- rewards are gated at 1e-6 relative)"}


# --------------------------------------------------------------------------------------------- CPU port
def make_oracle_env(envs, threads):
    """The CPU port on the benchmark's population prefix: same seed, env ids, phases as the GPU arm's rank 0."""
    import numpy as np
    from oracle import oracle as orc
    env = orc.OracleEnv(envs, seed=SEED, env_id_base=0, threads=threads)
    env.reset()
    env.step_flags[:] = (env.step_flags & ~np.uint32(orc.STEP_MASK)) | episode_phases(0, envs).astype(np.uint32)
    return env


def cpu_port_rate(envs, steps, warmup, threads, want_checksum=False):
    """env-steps/s of the oracle (C port of the reference algorithm) on `threads` host threads; optionally the gate
    prefix's checksum over its first GATE_STEPS steps."""
    env = make_oracle_env(envs, threads)
    n_gate = min(envs, GATE_ENVS)
    acts = gate_action_batches(n_gate)
    if envs > n_gate:
        import numpy as np
        rng = np.random.default_rng(1)
        acts = [np.concatenate([a, rng.uniform(-1, 1, (envs - n_gate, 8)).astype(np.float32)]) for a in acts]
    checksum = None
    done_steps = 0
    if want_checksum:
        cs = Checksum()
        for i in range(GATE_STEPS):
            obs, rew, done = env.step(acts[i & 1])
            cs.add(obs[:n_gate], rew[:n_gate], done[:n_gate])
        checksum = cs.result(n_gate, env.goal[:, :n_gate], env.step_flags[:n_gate])
        done_steps = GATE_STEPS
    for i in range(warmup):
        env.step(acts[(done_steps + i) & 1])
    t0 = time.perf_counter()
    for i in range(steps):
        env.step(acts[(done_steps + warmup + i) & 1])
    dt = time.perf_counter() - t0
    return envs * steps / dt, dt, checksum


def reference_python_on_this_box():
    """The UNMODIFIED Python reference (baseline/_ref, or /root/reference in the build container) timed on THIS box's
    host cores by oracle/time_reference.py --quick (a subprocess: it forks one worker per core)."""
    import subprocess
    try:
        from oracle import reference_harness as rh
        if not rh.available():
            raise RuntimeError("reference not present (baseline/_ref is installed by __graft_entry__.build())")
        out = subprocess.run([sys.executable, "-m", "oracle.time_reference", "--quick"], cwd=ROOT, stdout=subprocess.PIPE,
                             stderr=subprocess.PIPE, text=True, timeout=180)
        if out.returncode != 0:
            raise RuntimeError(out.stderr.strip().splitlines()[-1] if out.stderr.strip() else "exit %d" % out.returncode)
        return json.loads(out.stdout.strip().splitlines()[-1])
    except Exception as exc:   # context only: never fail the bench for it
        path = os.path.join(ROOT, "profiles", "r1_reference_python_cpu.json")
        ctx = {"unavailable_on_this_box": str(exc)[:200]}
        if os.path.exists(path):
            j = json.load(open(path))
            ctx.update({"where": j["where"], "cores": j["cores"],
                        "single_env_steps_per_s": j["single_env_steps_per_s"]["median"],
                        "vec_env_restated_steps_per_s": j["vec_env_restated_steps_per_s"]["value"],
                        "independent_processes_steps_per_s": j["independent_processes_steps_per_s"]["value"]})
        return ctx


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
    sample = min(args.envs, GATE_ENVS)     # bounded sample of the workload: the gate prefix, 2,097,152 envs per step
    rate, dt, checksum = cpu_port_rate(sample, max(1, args.steps), max(0, args.warmup), threads, want_checksum=True)
    cfg = workload_config(args.envs, args.gpus)
    cfg["reference_arm_sample"] = ("each timed step covers the first {:,} of the {:,} envs (same seed, env ids, episode phases "
                                   "and actions as the b200 arm's first {:,} envs)".format(sample, args.envs, sample))
    cfg["envs_per_step_in_this_arm"] = sample
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / max(1, args.steps), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": "{:,} envs x {} steps per timed run, oracle/roboy_oracle.c (C port of the "
                                   "reference algorithm), pthreads over all host cores".format(sample, args.steps),
                         "reference_python": reference_python_on_this_box()},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "parity_checksum": checksum, "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------- GPU arm
def device_timed(env, client, actions, steps, warmup, torch, dist, world, stats_every=100, per_step_events=True):
    """K steps, barrier + synchronize on both sides, CUDA events on the launching stream.
    Returns (elapsed_ms_total_max_over_ranks, mean_kernel_ms, median_kernel_ms, launches, collectives).
    per_step_events=False records only the first and the last event (host-bound sizes: an event record per step costs
    more than the step; and an event between two launches defeats the programmatic dependent launch)."""
    from gym_roboy_b200.sharding import all_reduce_stats
    for i in range(warmup):
        env.step(actions[i % len(actions)])
    torch.cuda.synchronize()
    if world > 1:
        all_reduce_stats(client.stats_tensor.clone())   # NCCL communicator warm-up outside the timed region
        dist.barrier()
        torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    launches0 = client.launch_count()
    collectives = 0
    ev[0].record()
    for i in range(steps):
        env.step(actions[i % len(actions)])
        if per_step_events or i == steps - 1:
            ev[i + 1].record()
        if world > 1 and (i + 1) % stats_every == 0 and i + 1 < steps:
            # the tiny episode-stat reduction (64 bytes), stream-ordered between two steps.  On a side
            # stream the NCCL kernel competes for SM slots with step grids that programmatic dependent launch has already
            # queued, which measured 10 % slower at 2 and 8 GPUs.
            all_reduce_stats(client.stats_tensor)
            collectives += 1
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
    return total_ms, mean_kernel_ms, per[len(per) // 2], client.launch_count() - launches0, collectives


def parity_gate(env, client, actions, begin, torch):
    """Step the BENCHMARKED shard GATE_STEPS times and compare its first GATE_ENVS envs with the CPU oracle (the
    checker) on the same actions.  Returns (gate dict, checksum dict)."""
    import numpy as np
    n_gate = min(client.num_envs, GATE_ENVS)
    ora = make_oracle_env(n_gate, host_threads()) if begin == 0 else None
    cs = Checksum()
    gate = {"envs": n_gate, "steps": GATE_STEPS, "done_xor": 0, "obs_xor": 0, "goal_xor": 0, "step_word_xor": 0,
            "reward_max_rel": 0.0, "done_count": 0,
            "what": "the benchmarked shard's first {:,} envs vs oracle/roboy_oracle.c on the same actions; *_xor = number "
                    "of differing elements".format(n_gate)}
    for i in range(GATE_STEPS):
        obs, rew, done, _ = env.step(actions[i & 1])
        o, r, d = obs[:n_gate].cpu().numpy(), rew[:n_gate].cpu().numpy(), done[:n_gate].cpu().numpy()
        cs.add(o, r, d)
        if ora is not None:
            oo, orr, od = ora.step(actions[i & 1][:n_gate].cpu().numpy())
            gate["done_xor"] += int(np.count_nonzero(d != od))
            gate["obs_xor"] += int(np.count_nonzero(o.view(np.uint32) != oo.view(np.uint32)))
            rel = np.abs(r.astype(np.float64) - orr) / np.maximum(np.abs(orr.astype(np.float64)), 1e-30)
            gate["reward_max_rel"] = max(gate["reward_max_rel"], float(rel.max()))
            gate["done_count"] += int(np.count_nonzero(od))
    goal = client.goal[:, :n_gate].cpu().numpy()
    words = client.step_flags[:n_gate].cpu().numpy().astype(np.uint32)
    if ora is not None:
        gate["goal_xor"] = int(np.count_nonzero(goal.view(np.uint32) != ora.goal.view(np.uint32)))
        gate["step_word_xor"] = int(np.count_nonzero(words != ora.step_flags))
    gate["passed"] = bool(ora is not None and gate["done_xor"] == 0 and gate["obs_xor"] == 0 and gate["goal_xor"] == 0 and
                          gate["step_word_xor"] == 0 and gate["reward_max_rel"] <= 1e-6 and gate["done_count"] > 0)
    return gate, cs.result(n_gate, goal, words)


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
    from gym_roboy_b200 import _native
    from gym_roboy_b200.envs import RoboyEnv
    from gym_roboy_b200.envs.simulations import CudaSimulationClient
    from gym_roboy_b200.sharding import all_reduce_stats, shard_range, summarize

    dev = torch.device("cuda", local)
    envs = args.envs
    begin, end = shard_range(envs * world, world, rank)

    def make(n, base=0):
        client = CudaSimulationClient(num_envs=n, seed=SEED, env_id_base=base, device=dev)
        env = RoboyEnv(client)
        env.reset()
        client.set_step_num(torch.as_tensor(episode_phases(base, n), device=dev))
        return env, client

    gate_batches = gate_action_batches(min(envs, GATE_ENVS)) if rank == 0 else None

    def make_actions(n, k=2, base=0):
        gen = torch.Generator(device=dev)
        gen.manual_seed(rank)            # seed 0 on rank 0
        acts = [torch.rand((n, 8), device=dev, generator=gen) * 2 - 1 for _ in range(k)]
        if base == 0 and gate_batches is not None:    # the prefix both arms share (one numpy generator)
            m = min(n, gate_batches[0].shape[0])
            for a, g in zip(acts, gate_batches):
                a[:m].copy_(torch.from_numpy(g[:m]).pin_memory(), non_blocking=False)
        return acts

    env, client = make(end - begin, begin)
    actions = make_actions(end - begin, base=begin)

    # ---- parity gate (before any timing counts) on the benchmarked shard itself, rank 0 ----
    gate, checksum = (None, None)
    if rank == 0:
        gate, checksum = parity_gate(env, client, actions, begin, torch)
    else:
        for i in range(GATE_STEPS):
            env.step(actions[i & 1])
    gate_ok = torch.tensor([1 if (gate is None or gate["passed"]) else 0], device=dev)
    if world > 1:
        dist.all_reduce(gate_ok, op=dist.ReduceOp.MIN)
    if int(gate_ok.item()) == 0:
        if rank == 0:
            print(json.dumps({"metric": METRIC, "value": None, "unit": UNIT, "n_gpus": world, "impl": "b200",
                              "error": "parity gate FAILED: no throughput is reported for a kernel whose results differ "
                                       "from the oracle's", "parity_gate": gate, "parity_checksum": checksum}), flush=True)
        if world > 1:
            dist.destroy_process_group()
        raise SystemExit(3)

    # The timed region is ONE CUDA-event pair around the K back-to-back launches: an event recorded between two
    # launches would keep the next step kernel from starting under the tail of the previous one (programmatic dependent
    # launch).  The average launch duration of the roofline is that region divided by its K launches; a second, untimed
    # pass with an event per launch gives the per-launch median as a diagnostic.
    # Episode statistics are all-reduced over NCCL every `stats_every` steps; clamped so that even the driver's short
    # runs time at least one collective.
    stats_every = max(1, min(args.stats_every, max(1, args.steps // 2)))
    with ClockSampler(local) as clocks:
        total_ms, mean_kernel_ms, _, launches, collectives = device_timed(
            env, client, actions, args.steps, args.warmup, torch, dist, world, stats_every=stats_every,
            per_step_events=False)
    value = envs * world * args.steps / (total_ms * 1e-3)
    _, per_launch_mean_ms, median_kernel_ms, _, _ = device_timed(env, client, actions, min(args.steps, 50), 3, torch, dist,
                                                                 world, stats_every=10 ** 9)

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
        a_h, obs_h, rew_h, done_h = client.host_buffers()
        a_h2 = client.host_buffers()[0]
        a_host = [a_h, a_h2]
        for a, src in zip(a_host, actions):
            torch.from_numpy(a).copy_(src)
        modes = {"staged": _native.HOST_STAGED, "mapped_out": _native.HOST_MAPPED_OUT, "mapped_all": _native.HOST_MAPPED_ALL}

        def time_e2e(mode, ramp=True, steps=k2, streams=None):
            client.set_host_mode(modes[mode])
            if streams is None:
                client.set_host_autotune(True)     # the library's default: ring, stream count tuned on calls 2 and 3
            else:
                client.set_host_pipeline(n_streams=streams, ramp=ramp)
            for i in range(3):
                client.step_host(a_host[i & 1], obs_h, rew_h, done_h)
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            l0 = client.launch_count()
            t0 = time.perf_counter()
            for i in range(steps):
                client.step_host(a_host[i & 1], obs_h, rew_h, done_h)   # returns with outputs in host memory
            dt = time.perf_counter() - t0
            if world > 1:
                t = torch.tensor([dt], dtype=torch.float64, device="cuda")
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                dt = float(t.item())
            return envs * world * steps / dt, dt, client.launch_count() - l0

        def probe(directions, monolithic):
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            ms = client.copy_probe(a_host[0], obs_h, rew_h, done_h, directions=directions, monolithic=monolithic, iters=4)
            if world > 1:
                t = torch.tensor([ms], dtype=torch.float64, device="cuda")
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms = float(t.item())
            return ms

        variants = {}
        for name, mode, ramp, streams in (("ring2_noramp", "staged", False, 2), ("ring2", "staged", True, 2), ("one_stream", "staged", True, 1),
                                          ("mapped_out", "mapped_out", True, 2), ("mapped_all", "mapped_all", True, 1)):
            try:
                variants[name] = time_e2e(mode, ramp, steps=max(3, k2 // 2), streams=streams)[0]
            except Exception as exc:
                variants[name] = "failed: %s" % str(exc)[:120]
        rate, dt, e2e_launches = time_e2e("staged")          # the library's default configuration: THE e2e number
        pipeline = client.host_pipeline()
        # the copy ceiling: the same bytes over the same buffers without the kernel, all ranks at once
        both_staged, both_mono = probe(3, False), probe(3, True)
        h2d_ms, d2h_ms = probe(1, True), probe(2, True)
        best_ms = min(both_staged, both_mono)
        ceiling = envs * world / (best_ms * 1e-3)
        e2e = {"value": rate, "unit": UNIT, "h2d_bytes_per_step": H2D_BYTES * n,
               "d2h_bytes_per_step": D2H_BYTES * n, "steps": k2, "ms_per_step": 1e3 * dt / k2,
               "api": "roboy_step_host (C-ABI), library defaults: page-locked host actions -> H2D -> step kernel -> D2H obs+reward+done, "
                      "{:,}-env stages (first stages shorter) on a ring of {} stream(s) -- the stream count is tuned by the "
                      "library on its 2nd and 3rd call".format(pipeline["stage_envs"], pipeline["n_streams"]),
               "gpu_launches": e2e_launches,
               "copy_ceiling": {"value": ceiling, "unit": UNIT, "ms_per_pass": best_ms,
                                "what": "roboy_host_copy_probe: the same H2D (32 B/env) and D2H (41 B/env) copies over the same "
                                        "buffers WITHOUT the kernel, both directions concurrently, all {} ranks at once, "
                                        "max over ranks; best of the staged pattern and one monolithic copy per array".format(world),
                                "staged_pattern_ms": both_staged, "monolithic_ms": both_mono,
                                "h2d_alone_GBps_per_gpu": H2D_BYTES * n / (h2d_ms * 1e-3) / 1e9,
                                "d2h_alone_GBps_per_gpu": D2H_BYTES * n / (d2h_ms * 1e-3) / 1e9,
                                "both_GBps_per_gpu": (H2D_BYTES + D2H_BYTES) * n / (best_ms * 1e-3) / 1e9,
                                "both_GBps_all_gpus": (H2D_BYTES + D2H_BYTES) * n * world / (best_ms * 1e-3) / 1e9},
               "frac_of_ceiling": rate / ceiling, "variants_env_steps_per_s": variants,
               "checksum": float(np.asarray(rew_h[:1024], np.float64).sum()),
               "numa_bound_cpus": len(numa_cpus) if numa_cpus else None}
        client.set_host_mode(_native.HOST_STAGED)
        del a_host, a_h, a_h2, obs_h, rew_h, done_h
    del env, actions
    client.close()
    del client
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
            tot, mean_ms, med_ms, _, _ = device_timed(e, c, acts, k, max(3, args.warmup), torch, dist, 1,
                                                      per_step_events=n > 1048576)
            row = {"envs": n, "env_steps_per_s": n * k / (tot * 1e-3), "ms_per_step": tot / k,
                   "GBps_algorithmic": BYTES_PER_ENV_STEP * n / (mean_ms * 1e-3) / 1e9,
                   "frac_of_peak": BYTES_PER_ENV_STEP * n / (mean_ms * 1e-3) / 1e9 / peak,
                   "regime": "launch-bound" if n <= 32768 else ("L2-resident" if BYTES_PER_ENV_STEP * n < 126e6 else "HBM-bound")}
            if n <= 4194304:
                # launch floor: an EMPTY kernel launched exactly like the step kernel (roboy_null_step), through the
                # same Python -> ctypes -> C-ABI path, eager and replayed from a CUDA graph
                for _ in range(20):
                    c.null_step()
                torch.cuda.synchronize()
                s_, e_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s_.record()
                for _ in range(200):
                    c.null_step()
                e_.record()
                torch.cuda.synchronize()
                row["launch_floor_eager_ms"] = s_.elapsed_time(e_) / 200
                g0 = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g0):
                    for _ in range(100):
                        c.null_step()
                g0.replay()
                torch.cuda.synchronize()
                s_.record()
                for _ in range(3):
                    g0.replay()
                e_.record()
                torch.cuda.synchronize()
                row["launch_floor_graph_ms"] = s_.elapsed_time(e_) / 300
                row["eager_minus_floor_ms"] = row["ms_per_step"] - row["launch_floor_eager_ms"]
                del g0
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
                row["cuda_graph_minus_floor_ms"] = gms - row["launch_floor_graph_ms"]
                row["reading"] = "{:.1f} us per step in a CUDA graph, of which {:.1f} us is the launch floor of an empty kernel".format(
                    gms * 1e3, row["launch_floor_graph_ms"] * 1e3)
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

    # ---- the flag variant of the same step kernel: joint_vel_penalty=True (roboy_env.py:98-100), steady state ----
    penalty_variant = None
    if world == 1 and not args.no_sweep:
        n = 1 << 24
        c = CudaSimulationClient(num_envs=n, seed=SEED, device=dev)
        e = RoboyEnv(c, joint_vel_penalty=True, strict=False)
        e.reset()
        c.set_step_num(torch.as_tensor(episode_phases(0, n), device=dev))
        gen = torch.Generator(device=dev); gen.manual_seed(2)
        a = [torch.rand((n, 8), device=dev, generator=gen) * 2 - 1 for _ in range(2)]
        for i in range(5):
            c.step_fused(a[i & 1])
        torch.cuda.synchronize()
        s_, e_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 30
        s_.record()
        for i in range(reps):
            c.step_fused(a[i & 1])
        e_.record()
        torch.cuda.synchronize()
        ms = s_.elapsed_time(e_) / reps
        penalty_variant = {"envs": n, "ms_per_step": ms, "env_steps_per_s": n / (ms * 1e-3),
                           "GBps_algorithmic": BYTES_PER_ENV_STEP * n / (ms * 1e-3) / 1e9,
                           "frac_of_peak": BYTES_PER_ENV_STEP * n / (ms * 1e-3) / 1e9 / peak,
                           "penalty_float32": c.penalty_float32,
                           "note": "velocity penalty of sampled states in float32, the reference's float64 expression re-run "
                                   "within 1e-5 of reward_range's bounds (DESIGN.md 5); reward_range violations counted: %d"
                                   % int(c.stats()["violations"])}
        del e, c, a
        torch.cuda.empty_cache()

    # ---- a robot that is not MSJ: 6 joints / 14 tendons on the generic step kernel ----
    generic_robot = generic_robot_15 = None
    if world == 1 and not args.no_sweep:
        from gym_roboy_b200.envs.robots import RoboyRobot
        from gym_roboy_b200.spaces import Box

        def time_generic(J, A, n, a_lo, a_hi, what):
            class OtherRobot(RoboyRobot):
                _A, _V, _T = Box(a_lo, a_hi, (J,), "float32"), Box(-0.6, 0.6, (J,), "float32"), Box(-0.2, 0.2, (A,), "float32")
                get_action_space = classmethod(lambda cls: cls._T)
                get_joint_angles_space = classmethod(lambda cls: cls._A)
                get_joint_vels_space = classmethod(lambda cls: cls._V)

            c = CudaSimulationClient(robot=OtherRobot(), num_envs=n, seed=SEED, device=dev)
            e = RoboyEnv(c, strict=False)
            e.reset()
            c.set_step_num(torch.as_tensor(episode_phases(0, n), device=dev))
            gen = torch.Generator(device=dev); gen.manual_seed(2)
            acts = [torch.rand((n, A), device=dev, generator=gen) * 2 - 1 for _ in range(2)]
            for i in range(5):
                c.step_fused(acts[i & 1])
            torch.cuda.synchronize()
            s_, e_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            k = 30
            s_.record()
            for i in range(k):
                c.step_fused(acts[i & 1])
            e_.record()
            torch.cuda.synchronize()
            ms = s_.elapsed_time(e_) / k
            nbytes = 4 * A + 16 * J + 13    # read: action 4A, goal 4J, step word 4; written: obs 12J, reward 4, done 1, step word 4
            row = {"robot": what, "envs": n, "ms_per_step": ms,
                   "env_steps_per_s": n / (ms * 1e-3), "algorithmic_bytes_per_env_step": nbytes,
                   "GBps_algorithmic": nbytes * n / (ms * 1e-3) / 1e9, "frac_of_peak": nbytes * n / (ms * 1e-3) / 1e9 / peak,
                   "kernel": "roboy::generic_step_kernel<%d, %s>" % (J, "true" if c.fast_division else "false"),
                   "msj_kernels": c.msj_kernels, "fast_division_proved_at_create": c.fast_division}
            c.close()
            del e, c, acts
            torch.cuda.empty_cache()
            return row

        generic_robot = time_generic(6, 14, 1 << 22, -2.5, 2.5, "6 joints, 14 tendons (a RoboyRobot plug-in other than MsjRobot)")
        generic_robot_15 = time_generic(15, 64, 1 << 22, -np.linspace(1.0, 3.0, 15), np.linspace(0.5, 3.1, 15),
                                        "15 joints, 64 tendons, one bound per joint (the largest robot the kernels take)")

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
        sample = GATE_ENVS
        rate1, dt1, _ = cpu_port_rate(sample, 2, 1, threads)           # calibrate
        k = max(2, min(200, int(8.0 / max(dt1 / 2, 1e-4))))            # ~8 s of wall clock on all cores
        rate, dt, _ = cpu_port_rate(sample, k, 1, threads)
        cpu_baseline = {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
                        "sample": "{:,} envs x {} steps ({:.1f} s wall on {} threads), oracle/roboy_oracle.c".format(
                            sample, k, dt, threads),
                        "reference_python": reference_python_on_this_box()}

    if rank == 0:
        cfg = workload_config(envs, world)
        cfg["stats_allreduce_every"] = stats_every if world > 1 else None
        cfg["collectives_in_timed_region"] = collectives
        if sweep:   # BASELINE.json's named single-GPU configs, kept at top level (the driver record drops `sweep`)
            named = {"configs[1] 4,096 envs": 4096, "configs[2] 262,144 envs": 262144}
            keep = ("env_steps_per_s", "ms_per_step", "frac_of_peak", "cuda_graph_ms_per_step", "cuda_graph_env_steps_per_s",
                    "launch_floor_eager_ms", "launch_floor_graph_ms", "l2_flushed_ms", "l2_flushed_GBps_algorithmic", "regime", "reading")
            by_n = {r["envs"]: r for r in sweep}
            cfg["baseline_configs"] = {k: {f: by_n[n].get(f) for f in keep} for k, n in named.items() if n in by_n}
            roofline["at_baseline_configs"] = {
                k: {"cuda_graph_GBps_algorithmic": by_n[n].get("cuda_graph_GBps_algorithmic"),
                    "cuda_graph_frac_of_peak": (by_n[n].get("cuda_graph_GBps_algorithmic") or 0) / peak,
                    "l2_flushed_frac_of_peak": (by_n[n].get("l2_flushed_GBps_algorithmic") or 0) / peak,
                    "launch_floor_graph_ms": by_n[n].get("launch_floor_graph_ms"), "regime": by_n[n]["regime"]}
                for k, n in named.items() if n in by_n}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": cfg, "roofline": roofline,
            "cpu_baseline": cpu_baseline, "e2e": e2e, "gpu_launches": launches, "clocks": clocks.summary(),
            "parity_gate": gate, "parity_checksum": checksum, "collectives_in_timed_region": collectives,
            "episode_stats": stats, "sweep": sweep, "open_loop": open_loop, "joint_vel_penalty": penalty_variant, "generic_robot": generic_robot, "generic_robot_15_joints": generic_robot_15, "rollout": rollout,
            "impl": "b200",
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
