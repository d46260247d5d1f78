"""Host-side multi-rank logic on CPU: frequency sharding and the slab gather over ``gloo`` with
world_size 2 (the NCCL path runs the same code on the GPU box)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from fftvis_b200.gpu.distributed import SlabGather, gather_slabs, shard_frequencies, shard_inputs


def test_shard_frequencies_balanced_and_contiguous():
    assert shard_frequencies(1024, 8) == [(i * 128, (i + 1) * 128) for i in range(8)]
    s = shard_frequencies(10, 4)
    assert s == [(0, 3), (3, 6), (6, 8), (8, 10)]
    assert shard_frequencies(2, 4) == [(0, 1), (1, 2), (2, 2), (2, 2)]
    for nf, w in [(1, 1), (7, 3), (1024, 5)]:
        s = shard_frequencies(nf, w)
        assert s[0][0] == 0 and s[-1][1] == nf and all(a[1] == b[0] for a, b in zip(s, s[1:]))
        sizes = [hi - lo for lo, hi in s]
        assert max(sizes) - min(sizes) <= 1


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, nf, dst, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        shards = shard_frequencies(nf, world)
        lo, hi = shards[rank]
        rng = np.random.default_rng(0)
        full = rng.normal(size=(nf, 3, 4, 5)) + 1j * rng.normal(size=(nf, 3, 4, 5))
        local = torch.as_tensor(full[lo:hi].copy())
        got = gather_slabs(local, shards, dst=dst)
        if dst is None or rank == dst:
            q.put((rank, bool(np.array_equal(got.numpy(), full))))
        else:
            q.put((rank, got is None))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("nf,dst", [(7, 0), (8, None), (1, 0)])
def test_gather_slabs_gloo_world2(nf, dst):
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, nf, dst, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok in res), res


def _slab_worker(rank, world, port, nf, nt, dst, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        shards = shard_frequencies(nf, world)
        lo, hi = shards[rank]
        rng = np.random.default_rng(1)
        want = rng.normal(size=(nt, nf, 4, 5)) + 1j * rng.normal(size=(nt, nf, 4, 5))
        if rank == dst:
            full = torch.zeros((nt, nf, 4, 5), dtype=torch.complex128)
            sg = SlabGather(shards, full, None, dst=dst)
        else:
            local = torch.zeros((nt, hi - lo, 4, 5), dtype=torch.complex128)
            sg = SlabGather(shards, None, local, dst=dst)
        for t in range(nt):                      # "compute" slab t, then post its transfer
            block = torch.as_tensor(want[t, lo:hi])
            if rank == dst:
                full[t, lo:hi] = block
            else:
                local[t] = block
            sg.post(t)
        sg.finish()
        q.put((rank, bool(np.array_equal(full.numpy(), want)) if rank == dst else True))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("nf,nt,dst", [(7, 3, 0), (8, 2, 1), (1, 2, 0)])
def test_per_time_slab_gather_gloo_world2(nf, nt, dst):
    """The per-slab point-to-point gather of run_sharded: ragged shards, either destination, a rank
    with an empty shard."""
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_slab_worker, args=(r, world, port, nf, nt, dst, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok in res), res


def test_shard_inputs_cuts_the_frequency_axis():
    kw = dict(freqs=np.arange(10.0), fluxes=np.arange(30.0).reshape(3, 10), beam_coefs=np.ones((4, 2, 10)), eps=1e-9)
    s = shard_inputs(kw, 3, 6)
    assert s["freqs"].tolist() == [3.0, 4.0, 5.0] and s["fluxes"].shape == (3, 3) and s["beam_coefs"].shape == (4, 2, 3)
    assert s["eps"] == 1e-9 and kw["fluxes"].shape == (3, 10)
    assert shard_inputs(dict(freqs=np.arange(4.0), fluxes=np.ones((2, 4, 2, 2))), 0, 2)["fluxes"].shape == (2, 2, 2, 2)


def _shared_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from fftvis_b200.gpu.distributed import SharedHostResult
        shards = shard_frequencies(5, world)
        shape, ok = (5, 2, 1, 3), True
        for slot in (0, 1, 0):                                 # two alternating segments, then the cached first one
            seg = SharedHostResult.get(int(np.prod(shape)) * 16, None, 0, slot)
            full = seg.array(shape, np.complex128)
            lo, hi = shards[rank]
            full[lo:hi] = (rank + 1) * (slot + 1) + 1j * np.arange(lo, hi)[:, None, None, None]
            dist.barrier()
            if rank == 0:
                for r, (a, b) in enumerate(shards):
                    ok &= bool(np.all(full[a:b].real == (r + 1) * (slot + 1)))
                    ok &= bool(np.all(full[a:b].imag == np.arange(a, b)[:, None, None, None]))
            dist.barrier()
        # a segment that cannot be created (here: every rank's constructor raises) is reported as not ok on EVERY
        # rank, so that all of them fall back to the gather together
        from multiprocessing import shared_memory

        class Boom:
            def __init__(self, *a, **k):
                raise OSError("no space left on /dev/shm")

        real = shared_memory.SharedMemory
        shared_memory.SharedMemory = Boom
        try:
            seg = SharedHostResult.get(12345, None, 0, 7)
        finally:
            shared_memory.SharedMemory = real
        ok &= seg.ok is False
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


def test_shared_host_result_gloo_world2():
    """Every rank writes its frequency block into the shared host array; rank 0 reads all of them."""
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_shared_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok in res), res
