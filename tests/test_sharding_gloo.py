"""Multi-rank host logic on CPU: two gloo ranks each own a shard of the env population (the
oracle stands in for the GPU shard -- test only), all-reduce the episode statistics, and the
union of the shards equals the single-rank population bit for bit."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gym_roboy_b200.sharding import all_reduce_stats, shard_range, summarize
from oracle import oracle as orc

TOTAL, T, SEED = 3001, 40, 9


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _actions(t):
    rng = np.random.default_rng(1000 + t)
    a = rng.uniform(-1, 1, (TOTAL, 8)).astype(np.float32)
    a[rng.random(TOTAL) < 0.05] = 0
    return a


def _run_shard(begin, end):
    env = orc.OracleEnv(end - begin, seed=SEED, env_id_base=begin)
    env.reset()
    env.step_flags[:] = (env.step_flags & ~np.uint32(orc.STEP_MASK)) | np.uint32(390)
    outs = []
    for t in range(T):
        obs, rew, done = env.step(_actions(t)[begin:end])
        outs.append((obs, rew, done))
    stats = torch.tensor([env.stats()[k] for k in orc.STAT_NAMES], dtype=torch.float64)
    return outs, stats


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    begin, end = shard_range(TOTAL, world, rank)
    outs, stats = _run_shard(begin, end)
    total = all_reduce_stats(stats)
    q.put((rank, begin, end, [o[2].copy() for o in outs], [o[0].copy() for o in outs], total.tolist(), stats.tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_shards_equal_single_population():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    full_outs, full_stats = _run_shard(0, TOTAL)
    for t in range(T):
        assert np.array_equal(np.concatenate([r[3][t] for r in res]), full_outs[t][2])
        assert np.array_equal(np.concatenate([r[4][t] for r in res]), full_outs[t][0])
    reduced = res[0][5]
    assert reduced == res[1][5]                                    # every rank holds the same sums
    for k, name in enumerate(orc.STAT_NAMES):
        if name == "sum_reward":
            assert abs(reduced[k] - full_stats[k].item()) <= 1e-9 * abs(full_stats[k].item())
        else:
            assert reduced[k] == full_stats[k].item(), name
    assert reduced[1] >= TOTAL                                     # everyone timed out at least once
    assert summarize(torch.tensor(reduced))["steps"] == TOTAL * T


def test_all_reduce_is_identity_without_a_process_group():
    s = torch.arange(8, dtype=torch.float64)
    out = all_reduce_stats(s)
    assert torch.equal(out, s) and out.data_ptr() != s.data_ptr()
