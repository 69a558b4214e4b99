"""Sharding of the env population over the GPUs of one box, one process per GPU.

The reference's only parallelism is one env per OS process under a vec-env wrapper
(gym_roboy/train_parallel.py:28-29); envs never interact.  So a population of `total_envs` is cut
into contiguous blocks of global env ids, one block per rank, and the step path needs NO
collective.  Philox counters use the GLOBAL env id, which makes every env's trajectory independent
of the number of ranks.  The only communication is the all-reduce (sum) of the 8 episode
statistics the step kernel accumulates -- 64 bytes, NCCL over NVLink on GPUs, gloo in CPU tests.
"""
import os

import torch
import torch.distributed as dist


def shard_range(total_envs: int, world_size: int, rank: int):
    """Contiguous block [begin, end) of global env ids owned by `rank`; blocks differ by at most
    one env and every block start is what `env_id_base` is set to."""
    if not (0 <= rank < world_size):
        raise ValueError("rank {} outside world of {}".format(rank, world_size))
    base, extra = divmod(int(total_envs), int(world_size))
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def bind_to_gpu_numa_node(device_index: int):
    """Restrict this process to the CPUs NVML reports as local to GPU `device_index`, so that pinned
    host buffers (first touch) and the copy-issuing thread sit on the GPU's NUMA node.  Matters for the
    host-buffer path with several ranks per box; a no-op when NVML or the affinity call is unavailable.
    Returns the CPU set chosen (or None)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        handle = pynvml.nvmlDeviceGetHandleByIndex(device_index)
        words = pynvml.nvmlDeviceGetCpuAffinity(handle, (os.cpu_count() + 63) // 64)
        local = {w * 64 + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
        allowed = local & set(os.sched_getaffinity(0))
        if allowed:
            os.sched_setaffinity(0, allowed)
            return sorted(allowed)
    except Exception:
        pass
    return None


def all_reduce_stats(stats: torch.Tensor, group=None) -> torch.Tensor:
    """Sum the per-shard statistics vector (float64 [8]) over all ranks; returns a new tensor.
    With no process group initialised (single GPU) this is the identity."""
    out = stats.detach().clone()
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(out, op=dist.ReduceOp.SUM, group=group)
    return out


def summarize(stats: torch.Tensor) -> dict:
    """Population-level episode metrics from a (reduced) statistics vector."""
    s = stats.tolist()
    steps, episodes, successes, timeouts, sum_reward, sum_eplen, holds, violations = s
    return dict(
        steps=steps, episodes=episodes, successes=successes, timeouts=timeouts, holds=holds, violations=violations,
        success_rate=successes / episodes if episodes else 0.0,
        mean_step_reward=sum_reward / steps if steps else 0.0,
        mean_episode_len=sum_eplen / episodes if episodes else None,
        # total reward per finished episode: exact once many episodes have finished (the steps of episodes still in
        # flight are in the numerator only)
        mean_episode_return=sum_reward / episodes if episodes else None,
    )


def make_sharded_env(total_envs: int, seed: int = 0, device=None, **env_kwargs):
    """Build this rank's shard of a `total_envs` population (RANK / WORLD_SIZE from torch.distributed)."""
    from .envs import RoboyEnv
    from .envs.simulations import CudaSimulationClient

    world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank() if world > 1 else 0
    begin, end = shard_range(total_envs, world, rank)
    client = CudaSimulationClient(num_envs=end - begin, device=device, seed=seed, env_id_base=begin)
    return RoboyEnv(client, **env_kwargs)
