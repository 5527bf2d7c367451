"""Multi-GPU plumbing: one process per GPU, environments sharded in contiguous blocks, no collective
on the step path.  The only exchange is an all-gather of the finished-episode statistics vector
(the keys of the reference's per-episode logging dict, src/env/env/env.py:115-125)."""
from __future__ import annotations

import torch
import torch.distributed as dist

from ._native import EPISODE_STAT_KEYS, NUM_EPISODE_STATS


def bind_host_to_device(device_index: int) -> dict:
    """Pin the calling process to the CPU cores NVML reports as local to GPU `device_index` (its NUMA node / socket).

    Opt-in, for the host face (`evac_step_host`: NumPy in / NumPy out) under one process per GPU: page-locked buffers
    allocated afterwards are first-touched on the GPU's own socket, so the per-step D2H copy of the observation block
    (1.5 KB per environment) does not cross the inter-socket link; with 8 ranks copying at once an unbound rank whose
    buffers sit on the other socket is bound by that link, not by its PCIe lanes.  Returns what was done (never raises:
    without NVML or on a single-node host it is a no-op)."""
    import os

    info = {"bound": False}
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(int(device_index))
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
        allowed = os.sched_getaffinity(0)
        cpus = (cpus & allowed) or allowed
        os.sched_setaffinity(0, cpus)
        info.update(bound=True, cpus=len(cpus), first_cpu=min(cpus), last_cpu=max(cpus))
    except Exception as exc:  # NVML missing, cpuset restrictions, non-Linux: leave the affinity alone
        info["reason"] = repr(exc)
    return info


def shard_offset(rank: int, envs_per_rank: int) -> int:
    """Global index of this rank's first environment; passed as `env_index_offset` so every
    environment draws the same Philox streams regardless of the number of ranks."""
    return int(rank) * int(envs_per_rank)


def shard_counts(total_envs: int, world_size: int):
    """Contiguous block partition of `total_envs` environments over `world_size` ranks ->
    list of (offset, count); the first `total_envs % world_size` ranks get one extra."""
    base, rem = divmod(int(total_envs), int(world_size))
    out, off = [], 0
    for r in range(world_size):
        c = base + (1 if r < rem else 0)
        out.append((off, c))
        off += c
    return out


def allgather_totals_tensor(totals: torch.Tensor) -> torch.Tensor:
    """totals: [1 + NUM_EPISODE_STATS] float64 on this rank's device (count of finished episodes and the
    sums of their statistics).  Returns [world, 1 + NUM_EPISODE_STATS] (same on every rank).  Uses the
    default process group (NCCL on GPUs, gloo in the CPU tests)."""
    assert totals.numel() == 1 + NUM_EPISODE_STATS
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return totals.reshape(1, -1).clone()
    world = dist.get_world_size()
    out = torch.empty((world, totals.numel()), dtype=totals.dtype, device=totals.device)
    dist.all_gather_into_tensor(out, totals.contiguous()) if totals.is_cuda else dist.all_gather(
        list(out.unbind(0)), totals.contiguous())
    return out


def allgather_episode_totals(env) -> torch.Tensor:
    """All-gather the finished-episode totals of `env` (an EvacuationEnv) across ranks."""
    _, _, totals = env.unwrapped.episode_statistics()
    return allgather_totals_tensor(totals)


def summarize_totals(gathered: torch.Tensor) -> dict:
    """Mean per-episode statistics over all ranks, keyed like the reference's logging dict."""
    g = gathered.double().sum(dim=0)
    n = max(float(g[0]), 1.0)
    out = {"episodes": float(g[0])}
    for i, k in enumerate(EPISODE_STAT_KEYS):
        out[k] = float(g[1 + i]) / n
    return out
