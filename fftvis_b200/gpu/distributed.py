"""Multi-GPU layer: one process per GPU (``torchrun``), frequency-major sharding, NCCL gather.

The reference's only parallelism is a Ray task farm over (frequency-chunk, time-chunk) blocks whose
results are disjoint slices ``vis[tc][..., fc] = future`` (/root/reference/src/fftvis/cpu/
cpu_simulate.py:711-835, 843-847; chunking rule core/utils.py:122-187).  Here every rank simulates a
contiguous block of frequencies for ALL times (rotate + horizon cut is recomputed per rank: O(Nsrc)
per time) and the finished ``(nf_local, nt, P, nbls)`` slabs are collected with ONE collective over
NVLink (``all_gather_into_tensor`` on equal, padded blocks, or a gather to rank 0).  There is no
exchange step inside the path, so no other collective exists.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def shard_frequencies(nfreqs: int, world_size: int) -> list[tuple[int, int]]:
    """Contiguous, balanced [lo, hi) frequency blocks, one per rank (first ranks get the remainder);
    ranks beyond ``nfreqs`` get empty blocks."""
    base, rem = divmod(int(nfreqs), int(world_size))
    out, lo = [], 0
    for r in range(world_size):
        n = base + (1 if r < rem else 0)
        out.append((lo, lo + n))
        lo += n
    return out


def gather_slabs(local: torch.Tensor, shards: list[tuple[int, int]], group=None, dst: int | None = 0):
    """Collect per-rank ``(nf_local, ...)`` slabs into the full ``(nf, ...)`` tensor.

    ``dst=None``: every rank gets the result (all-gather); otherwise only ``dst`` does (others get
    ``None``).  Blocks are padded to the largest shard so that one fixed-size collective suffices."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    nmax = max(hi - lo for lo, hi in shards)
    tail = tuple(local.shape[1:])
    pad = torch.zeros((nmax,) + tail, dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    # complex tensors travel as real pairs (NCCL has no complex dtype)
    send = torch.view_as_real(pad) if pad.is_complex() else pad
    if dst is None:
        recv = torch.empty((world * send.shape[0],) + tuple(send.shape[1:]), dtype=send.dtype, device=send.device)
        dist.all_gather_into_tensor(recv, send.contiguous(), group=group)
        recv = recv.view((world,) + tuple(send.shape))
    else:
        recv_list = [torch.empty_like(send) for _ in range(world)] if rank == dst else None
        dist.gather(send.contiguous(), recv_list, dst=dst, group=group)
        if rank != dst:
            return None
        recv = torch.stack(recv_list)
    if local.is_complex():
        recv = torch.view_as_complex(recv)
    parts = [recv[r, : hi - lo] for r, (lo, hi) in enumerate(shards)]
    return torch.cat(parts, dim=0)


def simulate_vis_sharded(engine, *, group=None, dst: int | None = 0, **simulate_kwargs):
    """Frequency-sharded ``simulate``: every rank passes the SAME full inputs; rank ``dst`` (or every
    rank when ``dst`` is None) returns the full host array, the others ``None``.

    ``engine`` is a ``GPUSimulationEngine`` bound to this rank's device."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    freqs = np.atleast_1d(np.asarray(simulate_kwargs["freqs"]))
    shards = shard_frequencies(freqs.size, world)
    plan = engine.prepare(freq_range=shards[rank], **simulate_kwargs)
    out = engine.run_plan(plan)
    engine.check_source_buffer(plan)
    full = gather_slabs(out, shards, group=group, dst=dst)
    if full is None:
        return None
    return engine.finish(plan, full)
