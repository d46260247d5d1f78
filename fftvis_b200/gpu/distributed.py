"""Multi-GPU layer: one process per GPU (``torchrun``), frequency-major sharding, NCCL gather.

The reference's only parallelism is a Ray task farm over (frequency-chunk, time-chunk) blocks whose
results are disjoint slices ``vis[tc][..., fc] = future`` (/root/reference/src/fftvis/cpu/
cpu_simulate.py:711-835, 843-847; chunking rule core/utils.py:122-187).  Here every rank simulates a
contiguous block of frequencies for ALL times (rotate + horizon cut is recomputed per rank: O(Nsrc)
per time) and the result is collected on one rank over NVLink.  There is no exchange step inside the
path, so the gather is the only collective.

Data movement, designed so that nothing is copied twice:

* every rank computes into a TIME-major buffer ``(nt, nf_local, P, nbls)`` (``run_plan`` writes through
  a permuted view), so each finished time slab is one contiguous block;
* as soon as slab ``t`` is enqueued, its transfer is posted on a communication stream: the senders
  ``isend`` the slab, the destination ``irecv``s every peer's slab straight into its place
  ``full[t, lo_r:hi_r]`` of the gathered ``(nt, nf, P, nbls)`` array (grouped ``ncclSend`` /
  ``ncclRecv``: the point-to-point form of ``ncclGather``, which also serves unequal shards) -- no
  padding, no stack / cat pass; the destination computes its own block in place inside ``full``;
* the transfers of slab ``t`` overlap the computation of slabs ``t+1...``; on the destination the
  gathered slab is then copied to the caller's page-locked host array (``(nf, nt, ...)``, the
  reference's layout) by a strided D2H on a third stream, so the PCIe copy hides under the compute too.
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


def _as_real(t: torch.Tensor) -> torch.Tensor:
    """Complex tensors travel as real pairs (NCCL has no complex dtype)."""
    return torch.view_as_real(t) if t.is_complex() else t


class SlabGather:
    """Gather of per-rank time slabs into ``full (nt, nf, ...)`` on rank ``dst``, one grouped
    point-to-point exchange per time slab.

    ``local_t``: this rank's ``(nt, nf_local, ...)`` buffer (on ``dst`` it is ``None``: the destination
    computes in place in ``full[:, lo:hi]``).  ``post(t)`` enqueues the transfer of slab ``t`` behind
    the work already enqueued on the current stream; ``finish()`` makes the current stream wait for
    every transfer.  Works on CUDA tensors over NCCL (asynchronously, on ``comm_stream``) and on CPU
    tensors over gloo (the CPU tests)."""

    def __init__(self, shards, full, local_t, group=None, dst: int = 0, comm_stream=None):
        self.shards = list(shards)
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.dst = int(dst)
        self.dst_global = dist.get_global_rank(group, self.dst) if group is not None else self.dst
        self.full, self.local_t = full, local_t
        self.comm_stream = comm_stream
        self.pending = []
        self.posted = 0
        if self.rank == self.dst:
            if full is None:
                raise ValueError("the destination rank needs the gathered array")
        elif local_t is None:
            raise ValueError("a sending rank needs its local buffer")

    def _peers(self):
        return [r for r in range(self.world) if r != self.dst and self.shards[r][1] > self.shards[r][0]]

    def _ops(self, t: int):
        ops = []
        if self.rank == self.dst:
            for r in self._peers():
                lo, hi = self.shards[r]
                src = dist.get_global_rank(self.group, r) if self.group is not None else r
                ops.append(dist.P2POp(dist.irecv, _as_real(self.full[t, lo:hi]), src, self.group))
        elif self.shards[self.rank][1] > self.shards[self.rank][0]:
            ops.append(dist.P2POp(dist.isend, _as_real(self.local_t[t]), self.dst_global, self.group))
        return ops

    def post(self, t: int):
        ops = self._ops(t)
        self.posted += 1
        if not ops:
            return
        if self.comm_stream is not None:
            cur = torch.cuda.current_stream()
            self.comm_stream.wait_stream(cur)
            with torch.cuda.stream(self.comm_stream):
                reqs = dist.batch_isend_irecv(ops)
                for q in reqs:
                    q.wait()          # NCCL: orders comm_stream behind the transfer, does not block the host
        else:
            self.pending.extend(dist.batch_isend_irecv(ops))

    def finish(self):
        for q in self.pending:
            q.wait()
        self.pending = []
        if self.comm_stream is not None:
            torch.cuda.current_stream().wait_stream(self.comm_stream)


def gather_slabs(local: torch.Tensor, shards: list[tuple[int, int]], group=None, dst: int | None = 0):
    """Collect per-rank ``(nf_local, ...)`` blocks into the full ``(nf, ...)`` tensor in one exchange,
    receiving straight into the result (no padding, stacking or concatenation).

    ``dst=None``: every rank gets the result (one ``all_gather_into_tensor`` when the shards are
    equal, else a broadcast of the gathered array); otherwise only ``dst`` does (others get ``None``)."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sizes = [hi - lo for lo, hi in shards]
    nf = shards[-1][1]
    tail = tuple(local.shape[1:])
    local = local.contiguous()
    if dst is None and len(set(sizes)) == 1:
        full = torch.empty((nf,) + tail, dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(_as_real(full), _as_real(local), group=group)
        return full
    root = 0 if dst is None else int(dst)
    full = torch.empty((nf,) + tail, dtype=local.dtype, device=local.device) if (rank == root or dst is None) else None
    lo, hi = shards[rank]
    if rank == root:
        full[lo:hi] = local
    sg = SlabGather(shards, full[None] if rank == root else None, local[None] if rank != root else None,
                    group=group, dst=root)
    sg.post(0)
    sg.finish()
    if dst is None:
        dist.broadcast(_as_real(full), dist.get_global_rank(group, root) if group is not None else root, group=group)
    return full if (rank == root or dst is None) else None


def shard_inputs(kw: dict, lo: int, hi: int) -> dict:
    """This rank's view of the ``simulate`` arguments: the frequency axis of ``freqs``, ``fluxes`` and
    ``beam_coefs`` cut to ``[lo, hi)`` (tabulated beams are cut by the engine when it interpolates them
    to the shard's frequencies), so that a rank uploads and holds only its own shard."""
    out = dict(kw)
    freqs = np.atleast_1d(np.asarray(kw["freqs"]))
    out["freqs"] = freqs[lo:hi]
    fl = np.asarray(kw["fluxes"])
    out["fluxes"] = fl[:, lo:hi]
    bc = kw.get("beam_coefs")
    if bc is not None:
        bc = np.asarray(bc)
        out["beam_coefs"] = bc[:, :, lo:hi] if bc.ndim == 3 else bc
    return out


_CPU_ONLY = ("nprocesses", "nthreads", "force_use_ray", "trace_mem", "enable_memory_monitor")


SHARED_HOST_MAX_BYTES = int(__import__("os").environ.get("FV_SHARED_HOST_MAX_BYTES", 16 << 30))
_NO_SHARED = object()


def _simulate_shared(engine, plan, shards, nfreqs: int, group, root: int):
    """``host_result="shared"``: this rank's block of the result streamed into the shared host array."""
    from .gpu_simulate import _CDT, _NP_C
    rank = dist.get_rank(group)
    P = 4 if plan.polarized else 1
    lo, hi = shards[rank]
    shape = (nfreqs, plan.ntimes, P, plan.nbls)
    cdt = _NP_C[plan.precision]
    nbytes = int(np.prod(shape)) * np.dtype(cdt).itemsize
    st = engine.__dict__.setdefault("_shared_state", {"slot": 0})
    st["slot"] ^= 1
    seg = SharedHostResult.get(nbytes, group, root, st["slot"])
    if not seg.ok:
        return _NO_SHARED
    full = seg.array(shape, cdt)
    mine = torch.from_numpy(full[lo:hi])                       # contiguous: the frequency axis is the outermost
    if hi > lo:
        if seg.registered:
            engine.run_plan(plan, host_out=mine)
        else:                                                  # page-locking refused: one plain copy at the end
            out = engine.run_plan(plan)
            torch.cuda.current_stream(plan.device).synchronize()
            mine.copy_(out.cpu())
    bad = torch.zeros(1, dtype=torch.int32, device=plan.device)
    if plan.work:
        bad += (plan.work["counts"] < 0).any().to(torch.int32)
    dist.all_reduce(bad, op=dist.ReduceOp.MAX, group=group)    # also orders every rank's copies before the read
    torch.cuda.current_stream(plan.device).synchronize()
    if int(bad.item()):
        engine.check_source_buffer(plan)
        raise ValueError("source_buffer too small on another rank's frequency shard. Increase source_buffer.")
    dist.barrier(group)
    return engine._shape_result(plan, full) if rank == root else None


def run_sharded(engine, plan, shards, group=None, dst: int = 0, full: torch.Tensor | None = None,
                host_out: torch.Tensor | None = None, work: dict | None = None):
    """One frequency-sharded pass with the per-slab gather (see the module docstring).  ``plan`` is this
    rank's plan of ITS shard.  On ``dst`` returns ``full`` -- the gathered ``(nt, nf, P, nbls)`` device
    array (time-major) -- and, when ``host_out`` ``(nf, nt, P, nbls)`` is given, streams every gathered
    slab into it; on the other ranks returns ``None``.  ``work`` caches the buffers between calls."""
    from .gpu_simulate import _CDT, _copy_stream
    rank = dist.get_rank(group)
    dev = plan.device
    P = 4 if plan.polarized else 1
    nt, nbls = plan.ntimes, plan.nbls
    nf = shards[-1][1]
    lo, hi = shards[rank]
    cdt = _CDT[plan.precision]
    work = work if work is not None else {}
    with torch.cuda.device(dev):
        if "comm" not in work:
            work["comm"] = torch.cuda.Stream(device=dev)
        if rank == dst:
            if full is None:
                full = work.get("full")
                if full is None or tuple(full.shape) != (nt, nf, P, nbls) or full.dtype != cdt:
                    full = work["full"] = torch.empty((nt, nf, P, nbls), dtype=cdt, device=dev)
            mine = full[:, lo:hi].permute(1, 0, 2, 3)
            local_t = None
        else:
            local_t = work.get("local")
            if local_t is None or tuple(local_t.shape) != (nt, hi - lo, P, nbls) or local_t.dtype != cdt:
                local_t = work["local"] = torch.empty((nt, hi - lo, P, nbls), dtype=cdt, device=dev)
            mine = local_t.permute(1, 0, 2, 3)
        sg = SlabGather(shards, full, local_t, group=group, dst=dst, comm_stream=work["comm"])
        st = torch.cuda.current_stream()
        copy_st = _copy_stream(dev) if (rank == dst and host_out is not None) else None
        if copy_st is not None:
            if tuple(host_out.shape) != (nf, nt, P, nbls) or host_out.dtype != cdt or not host_out.is_contiguous():
                raise ValueError("host_out must be a contiguous (nf, nt, P, nbls) host tensor of the result dtype")
        view_ft = full.permute(1, 0, 2, 3) if rank == dst else None

        def hook(to):
            sg.post(to)
            if copy_st is not None:
                # gathered slab `to` -> host, behind this rank's own block (current stream) and the receives
                copy_st.wait_stream(work["comm"])
                engine._stream_slab(view_ft, host_out, to, st, copy_st)

        engine.run_plan(plan, out=mine, slab_hook=hook)
        sg.finish()
        if copy_st is not None:
            st.wait_stream(copy_st)
    return full if rank == dst else None


class SharedHostResult:
    """The result array in POSIX shared memory, mapped and page-locked (``cudaHostRegister``) by every rank of
    one node: each rank streams the finished time slabs of ITS frequency block over its own PCIe link straight
    into its rows of the ``(nf, nt, P, nbls)`` array, instead of sending them to one rank that owns the only
    host copy (whose single link then bounds the call).  Two segments alternate between calls, so the array a
    call returned stays intact during the next call; segments are cached per size."""

    _cache: dict = {}

    def __init__(self, nbytes: int, group, root: int, tag: int):
        from multiprocessing import shared_memory
        rank = dist.get_rank(group)
        name = [None]
        self.shm = None
        if rank == root:
            try:
                # tmpfs lets a segment be created larger than the space it has and faults (SIGBUS) on the first write
                # beyond it: ask for the room first
                import os
                vfs = os.statvfs("/dev/shm")
                if vfs.f_bavail * vfs.f_frsize < int(nbytes) + (64 << 20):
                    raise OSError("not enough room in /dev/shm")
                self.shm = shared_memory.SharedMemory(create=True, size=max(int(nbytes), 1))
                name[0] = self.shm.name
            except Exception:                 # /dev/shm too small, no permission ...: every rank learns it below
                self.shm = None
        dist.broadcast_object_list(name, src=dist.get_global_rank(group, root) if group is not None else root,
                                   group=group)
        self.ok = name[0] is not None
        if not self.ok:
            return
        if rank != root:
            self.shm = shared_memory.SharedMemory(name=name[0])
            try:                              # attaching registered the segment with this process' resource tracker,
                from multiprocessing import resource_tracker      # which would unlink it again at exit
                resource_tracker.unregister(self.shm._name, "shared_memory")
            except Exception:
                pass
        self.owner = rank == root
        self.nbytes = int(nbytes)
        self.bytes_view = np.ndarray((self.nbytes,), dtype=np.uint8, buffer=self.shm.buf)
        self.registered = False
        if torch.cuda.is_available() and self.nbytes:
            rt = torch.cuda.cudart()
            rc = rt.cudaHostRegister(self.bytes_view.ctypes.data, self.nbytes, 0)
            self.registered = int(rc) == 0
            if not self.registered:
                # locked-memory limit, unsupported mapping ...: clear the error so that it does not surface at the next
                # CUDA call; the caller then copies its block once at the end instead of streaming slabs
                try:
                    rt.cudaGetLastError()
                except Exception:
                    pass
        dist.barrier(group)
        if rank == root:
            self.shm.unlink()                 # the mappings keep it alive; nothing is left behind in /dev/shm

    @classmethod
    def get(cls, nbytes: int, group, root: int, slot: int):
        key = (int(nbytes), id(group), int(root), int(slot))
        if key not in cls._cache:
            cls._cache[key] = cls(nbytes, group, root, slot)
        return cls._cache[key]

    def array(self, shape, dtype) -> np.ndarray:
        return np.ndarray(tuple(shape), dtype=dtype, buffer=self.shm.buf)


def simulate_vis_sharded(engine, *, group=None, dst: int | None = 0, shards=None, host_result: str = "gather",
                         **simulate_kwargs):
    """Frequency-sharded ``simulate``: every rank passes the SAME full inputs (the arguments of
    ``GPUSimulationEngine.simulate``); rank ``dst`` (or every rank when ``dst`` is None) returns the full
    host array ``(nf, nt, [2, 2,] nbls)``, the others ``None``.

    ``shards``: when given (``shard_frequencies(nfreqs, world)``), the per-frequency inputs (``freqs``,
    ``fluxes``, ``beam_coefs``, tabulated beams) are ALREADY this rank's block -- for workloads whose
    full per-frequency inputs should never exist on one host (cfg5: 21 GB of basis-beam tables).

    ``host_result``: ``"gather"`` -- the finished time slabs are collected on ``dst``'s GPU by NCCL and copied to
    its page-locked result from there (the gathered array also stays device-resident on ``dst``);
    ``"shared"`` (ranks of one node, ``dst`` not None) -- every rank streams its own block straight into a
    shared page-locked host array (``SharedHostResult``): no device-side gather, all PCIe links in use.

    ``engine`` is a ``GPUSimulationEngine`` bound to this rank's device."""
    if host_result not in ("gather", "shared"):
        raise ValueError("host_result must be 'gather' or 'shared'")
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    kw = {k: v for k, v in simulate_kwargs.items() if k not in _CPU_ONLY}
    root = 0 if dst is None else int(dst)
    if shards is None:
        nfreqs = int(np.size(kw["freqs"]))
        shards = shard_frequencies(nfreqs, world)
        kw = shard_inputs(kw, *shards[rank])
    else:
        shards = [tuple(x) for x in shards]
        nfreqs = shards[-1][1]
        if int(np.size(kw["freqs"])) != shards[rank][1] - shards[rank][0]:
            raise ValueError("with shards=..., freqs must be this rank's block of the frequency axis")
    plan = engine.prepare(**kw)
    P = 4 if plan.polarized else 1
    if host_result == "shared" and dst is not None:
        # beyond SHARED_HOST_MAX_BYTES (page-locking tens of GB of shared memory is refused on common settings), or
        # where the segment cannot be created, the call goes through the NCCL gather instead -- the same decision on
        # every rank
        nbytes = nfreqs * plan.ntimes * P * plan.nbls * (8 * plan.precision)
        if nbytes <= SHARED_HOST_MAX_BYTES:
            res = _simulate_shared(engine, plan, shards, nfreqs, group, root)
            if res is not _NO_SHARED:
                return res
    host = None
    if rank == root:
        from .gpu_simulate import _CDT
        try:
            host = torch.empty((nfreqs, plan.ntimes, P, plan.nbls), dtype=_CDT[plan.precision], pin_memory=True)
        except RuntimeError:
            host = None
    full = run_sharded(engine, plan, shards, group=group, dst=root, host_out=host,
                       work=engine.__dict__.setdefault("_shard_work", {}))
    # a chunk overflow on ANY rank must raise on EVERY rank (no rank may be left waiting in a collective)
    bad = torch.zeros(1, dtype=torch.int32, device=plan.device)
    if plan.work:
        bad += (plan.work["counts"] < 0).any().to(torch.int32)
    dist.all_reduce(bad, op=dist.ReduceOp.MAX, group=group)
    if int(bad.item()):
        engine.check_source_buffer(plan)
        raise ValueError("source_buffer too small on another rank's frequency shard. Increase source_buffer.")
    torch.cuda.current_stream(plan.device).synchronize()
    res = None
    if rank == root:
        if host is None:
            host = full.permute(1, 0, 2, 3).contiguous().cpu()
        res = engine._shape_result(plan, host.numpy())
    if dst is None:
        box = [res]
        dist.broadcast_object_list(box, src=dist.get_global_rank(group, root) if group is not None else root,
                                   group=group)
        res = box[0]
    return res
