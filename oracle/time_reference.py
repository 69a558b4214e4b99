"""Time the UNMODIFIED Python reference on this machine's CPU -- TEST/BENCH INFRASTRUCTURE.

    python -m oracle.time_reference            full run, writes profiles/r1_reference_python_cpu.json
    python -m oracle.time_reference --quick    ~5 s, prints ONE JSON line (bench.py runs this on the bench box)

Needs the reference: /root/reference in the build container, or the unmodified copy under baseline/_ref/ that
__graft_entry__.build() installs and that travels to the GPU box.

SURVEY.md 8d CPU baseline: (i) one reference RoboyEnv(StubSimulationClient(MsjRobot())) on one core,
U(-1,1) float32 actions, reset() on done; (ii) the "vectorised-env path": stable-baselines is not
installable, so SubprocVecEnv is restated as P worker processes each owning one reference env,
lock-stepped through pipes with reset-on-done in the worker; (iii) the no-pipe upper bound
(P independent loops).  bench.py attaches the --quick numbers, measured on the box it runs on, to its
cpu_baseline object (`reference_python`).
"""
import contextlib
import io
import json
import multiprocessing as mp
import os
import sys
import time

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)
from oracle.reference_harness import _import_reference  # noqa: E402


def _make_env():
    RoboyEnv, MsjRobot, _, Stub = _import_reference()
    with contextlib.redirect_stdout(io.StringIO()):
        return RoboyEnv(Stub(robot=MsjRobot()))


def single_env(steps, warmup=1000, seed=0):
    env = _make_env()
    rng = np.random.default_rng(seed)
    acts = rng.uniform(-1, 1, (1024, 8)).astype(np.float32)
    with contextlib.redirect_stdout(io.StringIO()):
        env.reset()
        t0 = None
        for i in range(warmup + steps):
            if i == warmup:
                t0 = time.perf_counter()
            _, _, done, _ = env.step(acts[i & 1023])
            if done:
                env.reset()
    return steps / (time.perf_counter() - t0)


def _worker(conn, seed):
    env = _make_env()
    with contextlib.redirect_stdout(io.StringIO()):
        env.reset()
        while True:
            a = conn.recv()
            if a is None:
                break
            obs, r, d, info = env.step(a)
            if d:
                obs = env.reset()          # SubprocVecEnv worker semantics
            conn.send((obs, r, d, info))


def vec_env(P, steps):
    ctx = mp.get_context("fork")
    pipes, procs = [], []
    for k in range(P):
        a, b = ctx.Pipe()
        p = ctx.Process(target=_worker, args=(b, k), daemon=True)
        p.start()
        pipes.append(a)
        procs.append(p)
    rng = np.random.default_rng(0)
    acts = rng.uniform(-1, 1, (256, P, 8)).astype(np.float32)
    for i in range(200):
        for k, c in enumerate(pipes):
            c.send(acts[i & 255, k])
        [c.recv() for c in pipes]
    t0 = time.perf_counter()
    for i in range(steps):
        for k, c in enumerate(pipes):
            c.send(acts[i & 255, k])
        [c.recv() for c in pipes]
    dt = time.perf_counter() - t0
    for c in pipes:
        c.send(None)
    for p in procs:
        p.join(timeout=5)
    return P * steps / dt


def _indep(q, steps):
    q.put(single_env(steps, warmup=500))


def independent(P, steps):
    ctx = mp.get_context("fork")
    q = ctx.Queue()
    procs = [ctx.Process(target=_indep, args=(q, steps)) for _ in range(P)]
    t0 = time.perf_counter()
    for p in procs:
        p.start()
    rates = [q.get() for _ in procs]
    for p in procs:
        p.join()
    return sum(rates), time.perf_counter() - t0


def quick():
    """Bounded (~5 s) version for bench.py: measured on whatever box runs it, core count stated."""
    from oracle import reference_harness as rh
    cores = len(os.sched_getaffinity(0))
    single = sorted(single_env(3000, warmup=300) for _ in range(2))
    vec = vec_env(cores, 400)
    ind, _ = independent(cores, 3000)
    return {"what": "unmodified Python reference (gym-roboy RoboyEnv + StubSimulationClient) through tests/_shim",
            "where": "this box", "reference_root": rh.REFERENCE_ROOT, "cores": cores,
            "python": sys.version.split()[0], "numpy": np.__version__,
            "single_env_steps_per_s": single[-1], "vec_env_restated_steps_per_s": vec,
            "independent_processes_steps_per_s": ind,
            "note": "restated SubprocVecEnv (stable-baselines is not installable): one process per core, pipes, lock-step, "
                    "reset in the worker; `independent` = the same processes without pipes (upper bound)"}


if __name__ == "__main__" and "--quick" in sys.argv:
    print(json.dumps(quick()))
    sys.exit(0)

if __name__ == "__main__":
    cores = len(os.sched_getaffinity(0))
    trials = sorted(single_env(20000) for _ in range(3))
    vec = vec_env(cores, 3000)
    ind, _ = independent(cores, 10000)
    out = {
        "what": "unmodified Python reference (gym-roboy RoboyEnv + StubSimulationClient) through tests/_shim",
        "where": "build container (no GPU); the reference cannot travel to the GPU box",
        "cores": cores, "python": sys.version.split()[0], "numpy": np.__version__,
        "single_env_steps_per_s": {"median": trials[1], "best": trials[2], "trials": trials, "steps": 20000},
        "vec_env_restated_steps_per_s": {"value": vec, "processes": cores, "note": "restated SubprocVecEnv: pipes, lock-step, reset in worker"},
        "independent_processes_steps_per_s": {"value": ind, "processes": cores, "note": "no pipes: upper bound"},
    }
    path = os.path.join(_ROOT, "profiles", "r1_reference_python_cpu.json")
    json.dump(out, open(path, "w"), indent=1)
    print(json.dumps(out, indent=1))
