"""GPU simulation engine: fills ``GPUSimulationEngine`` (stub at
/root/reference/src/fftvis/gpu/gpu_simulate.py:20-91) with the CPU engine's real interface
(/root/reference/src/fftvis/cpu/cpu_simulate.py:537-569 ``simulate``, :856-884
``_evaluate_vis_chunk``).

Structure (B200-first, not the reference's per-(time, frequency) Python loop):

* ``simulate``  = host front half (dtype casts, default baselines, gridding / plane rotation;
  cpu_simulate.py:583-681)  ->  ``SimulationPlan`` (everything uploaded once: catalogue unit
  vectors, frequency-major fluxes, baseline tables per beam pair, beam tables)  ->
  ``run_plan`` (device-only loop)  ->  one D2H of the finished ``(nf, nt, [2, 2,] nbls)`` array.
* ``run_plan`` per time step: ONE fused rotate + horizon-cut + compaction + az/za + array-plane
  rotation pass (``fv_rotate_cut``); then, per *batch of frequencies* (they share the
  above-horizon source set, the NU points differ only by the scalar frequency): beam evaluation
  + apparent coherency (``fv_weights``) and a frequency-batched NUFFT whose epilogue applies the
  flipped-baseline conjugation, the feed-axis swap and the scatter into the output array in its
  final layout (``fv_nufft2d1`` for gridded arrays, ``fv_nufft3`` otherwise).  The live source
  count never leaves the device, so a whole time step is enqueued without host synchronisation
  on the type-1 path (type 3 reads the NU-point extents back once per time step to size its
  grids, as finufft does inside every call).

No CPU fallback: every entry point raises ``FVError`` without the CUDA library or a GPU.
"""
from __future__ import annotations

import logging
import math
from dataclasses import dataclass, field

import numpy as np
import torch

from ..beam_models import as_beam_model
from ..core import antenna_gridding, catalog, coords
from ..core import utils as core_utils
from ..core.simulate import SimulationEngine, default_accuracy_dict
from . import _lib
from .beams import BeamTiles, DeviceBeam, GPUBeamEvaluator, launch_weights, launch_weights_basis, resolve_interpolation
from .nufft import ModeSet, NufftPlan, default_plan

logger = logging.getLogger(__name__)

_RDT = {1: torch.float32, 2: torch.float64}
_CDT = {1: torch.complex64, 2: torch.complex128}
_NP_R = {1: np.float32, 2: np.float64}
_NP_C = {1: np.complex64, 2: np.complex128}
_FEED_SWAP = (0, 2, 1, 3)     # transform [a*2+p] -> output slot [p*2+a]  (cpu_simulate.py:300)


@dataclass
class _PairTable:
    """Device tables of one unique beam pair (cpu/beams.py:91-127 routing)."""
    bi: int
    bj: int
    nk: int
    kmap: torch.Tensor | None          # int32 baseline indices (None: all baselines in order)
    conj: torch.Tensor | None          # uint8 flipped flags (None: none flipped)
    m1: torch.Tensor | None = None     # type 1: signed integer modes (flip already applied)
    m2: torch.Tensor | None = None
    modes: ModeSet | None = None       # type 1: the same modes bucketed by m1 (fused path)
    ncols: int = 0                     # type 1: number of distinct m1 (columns the fused path keeps)
    uvw: list | None = None            # type 3: per-unit-frequency targets (flip already applied)
    ulim: list | None = None           # type 3: {min, max} of each target coordinate


@dataclass
class SimulationPlan:
    """Device-resident state of one ``simulate`` call (or one rank's frequency shard of it)."""
    precision: int
    polarized: bool
    polarized_sky: bool
    nfeeds: int
    eps: float
    upsample_factor: float
    use_type1: bool
    is_coplanar: bool
    n_modes: int | None
    nbls: int
    nsrc: int
    nchunks: int
    n_cap: int
    freqs_host: np.ndarray             # working-precision frequencies (ALL of them)
    f_lo: int
    f_hi: int
    enu_mats: np.ndarray               # (nt, 3, 3) fp64
    astrom: np.ndarray | None          # (nt, 10) fp64 deflection / aberration block (None: rotation only)
    plane_mat: np.ndarray              # (3, 3) working precision, as fp64
    eq_xyz: torch.Tensor               # (3, nsrc) fp64
    flux: torch.Tensor                 # (nf_total, nsrc) or (nf_total, 4, nsrc) cplx
    freqs_dev: torch.Tensor            # (nf_total,) fp64 of the working-precision values
    beams: list                        # DeviceBeam per beam_list entry
    pairs: list                        # _PairTable per unique beam pair (standard path)
    basis: dict | None                 # basis path: coefs (nant, K, nf), ant1, ant2
    freq_batch: int
    device: torch.device
    work: dict = field(default_factory=dict)

    @property
    def ntimes(self) -> int:
        return self.enu_mats.shape[0]

    @property
    def nf_local(self) -> int:
        return self.f_hi - self.f_lo


_COPY_STREAMS: dict = {}


def _copy_stream(dev) -> torch.cuda.Stream:
    """One side stream per device for the result slabs' D2H copies."""
    key = torch.device(dev).index
    if key not in _COPY_STREAMS:
        _COPY_STREAMS[key] = torch.cuda.Stream(device=dev)
    return _COPY_STREAMS[key]


_SIDE_STREAMS: dict = {}


def _side_stream(dev, i: int = 0) -> torch.cuda.Stream:
    """Side compute streams of a device: alternate frequency batches of the fused type-1 path run on them, so
    that the tail of one batch's kernels overlaps the head of the next batch's."""
    key = (torch.device(dev).index, i)
    if key not in _SIDE_STREAMS:
        _SIDE_STREAMS[key] = torch.cuda.Stream(device=dev)
    return _SIDE_STREAMS[key]


def _next235even(n: int) -> int:
    return int(_lib.lib().fv_next235even(int(n)))


class GPUSimulationEngine(SimulationEngine):
    """GPU implementation of the simulation engine."""

    def __init__(self, device=None, freq_batch: int | None = None, grid_budget_bytes: int = 48 << 20,
                 type1_method: str = "fused"):
        if type1_method not in ("fused", "cufft"):
            raise ValueError("type1_method must be 'fused' or 'cufft'")
        self.device = device
        self.type1_method = type1_method
        self.freq_batch = freq_batch
        self.grid_budget_bytes = int(grid_budget_bytes)
        self.last_fft_ms = None
        self._nufft = None
        # table beams of order 0 / 1: "sort" = sort the live set by beam-grid tile, taps gathered from global
        # memory by neighbouring threads (measured fastest: the weights kernels are bound by their fp64
        # arithmetic at low occupancy, not by table traffic); True = patches staged in shared memory by bulk
        # copies (csrc/weights_tiled.cuh); False = catalogue order
        self.beam_tiles = "sort"
        # per-stage device time of the stages outside the NUFFT plan (rotate + cut, tile sort, weights): CUDA
        # events around every launch while ``time_stages`` is set (bench.py); read with ``stage_times()``
        self.time_stages = False
        self._stage_events = []
        # fused type-1 path: odd frequency batches on a side stream with their own strengths / work buffers
        # (the kernels of consecutive batches then overlap: pass 2's latency-bound waves and the wave tails of
        # pass 1 fill with the other batch's CTAs); False = one stream
        self.two_streams = True
        self.side_streams = 2                            # measured on cfg2: 178.7 / 174.7 / 172.1 ms per step with 1 / 2 / 3

    # ------------------------------------------------------------------------------------------
    def _timed(self, name, st, fn, *a, **k):
        if not self.time_stages:
            return fn(*a, **k)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        r = fn(*a, **k)
        e1.record(st)
        self._stage_events.append((name, e0, e1))
        return r

    def stage_times(self, reset=True) -> dict:
        """``{stage: (milliseconds, launches)}`` of the engine-level stages recorded since the last reset."""
        torch.cuda.synchronize()
        out = {}
        for name, e0, e1 in self._stage_events:
            ms, n = out.get(name, (0.0, 0))
            out[name] = (ms + e0.elapsed_time(e1), n + 1)
        if reset:
            self._stage_events = []
        return out

    def _device(self) -> torch.device:
        _lib.require_gpu()
        if self.device is not None:
            return torch.device(self.device)
        return torch.device("cuda", torch.cuda.current_device())

    def nufft_plans(self, dev) -> list:
        """Every NUFFT plan ``run_plan`` launches on: the current stream's and, with ``two_streams``, the side
        stream's (for callers that switch the per-stage timers on and read them: bench.py)."""
        plans = [self._nufft_plan(dev)]
        if self.two_streams:
            for i in range(max(1, int(self.side_streams))):
                with torch.cuda.device(dev), torch.cuda.stream(_side_stream(dev, i)):
                    plans.append(default_plan())
        return plans

    def _nufft_plan(self, dev) -> NufftPlan:
        """The process-wide plan of (device, current stream): its work areas and twiddle tables outlive
        the engine, so that back-to-back ``simulate_vis`` calls do not pay a plan teardown + rebuild."""
        with torch.cuda.device(dev):
            self._nufft = default_plan()
        return self._nufft

    # ------------------------------------------------------------------------------------------
    def prepare(self, ants, freqs, fluxes, beam_list, ra, dec, times, telescope_loc,
                baselines=None, beam_idx=None, precision=2, polarized=False, eps=None,
                upsample_factor=2, beam_spline_opts=None, flat_array_tol=1e-6,
                coord_method_params=None, force_use_type3=False, nchunks=1, source_buffer=1.0,
                beam_coefs=None, freq_range=None, coord_method="CoordinateRotationERFA",
                interpolation_function="az_za_map_coordinates") -> SimulationPlan:
        """Host front half of ``simulate`` (cpu_simulate.py:583-709) + upload of every input.
        ``freq_range=(lo, hi)`` restricts the plan to that slice of ``freqs`` (one rank's shard)."""
        dev = self._device()
        if precision not in (1, 2):
            raise ValueError("precision must be 1 or 2")
        rd, cd = _NP_R[precision], _NP_C[precision]
        nfreqs = int(np.size(freqs))
        nbeam, nant = len(beam_list), len(ants)
        if eps is None:
            eps = default_accuracy_dict[precision]
        ra = np.asarray(ra).astype(rd, copy=False)
        dec = np.asarray(dec).astype(rd, copy=False)
        freqs = np.atleast_1d(np.asarray(freqs)).astype(rd, copy=False)
        beam_idx = core_utils.validate_beam_idx(beam_idx, beam_coefs, nbeam, nant)
        if beam_coefs is not None and not polarized:
            raise ValueError(
                "Basis decomposition is not compatible with unpolarized simulations. "
                "Set polarized=True to use beam_coefs.")
        ants = {k: np.asarray(v, dtype=float) for k, v in ants.items()}
        if baselines is None:
            baselines = core_utils.get_pos_reds(ants, include_autos=True, representatives_only=True)
        baselines = [tuple(b) for b in baselines]
        nbls = len(baselines)
        fluxes = np.asarray(fluxes)
        if fluxes.ndim == 2:
            # Stokes I: the 0.5 I scaling and the cast (cpu/utils.py:52-53, cpu_simulate.py:622-626)
            # are applied on the device after the upload of the caller's array as it is
            coherency, pol_sky = None, False
        else:
            coherency, pol_sky = catalog.prepare_source_catalog(fluxes, polarized_beam=polarized)
        nsrc = int(np.size(dec))

        # ---- array geometry: gridded -> type 1; else plane rotation -> type 3 (:628-681)
        antnums = list(ants.keys())
        a_index = {a: i for i, a in enumerate(antnums)}
        antvecs = np.array([ants[a] for a in antnums], dtype=rd)
        basis_matrix, n_modes = None, None
        if np.abs(antvecs[:, -1]).max() > flat_array_tol or force_use_type3:
            is_gridded = False
        else:
            is_gridded, gridded, basis_matrix = antenna_gridding.check_antpos_griddability(ants)
        i0 = np.array([a_index[b[0]] for b in baselines], dtype=np.int64)
        i1 = np.array([a_index[b[1]] for b in baselines], dtype=np.int64)
        if not is_gridded:
            n_modes = None
            rot = np.ascontiguousarray(core_utils.get_plane_to_xy_rotation_matrix(antvecs).T)
            rants = np.dot(rot, antvecs.T)
            bls = (rants[:, i1] - rants[:, i0]) if nbls else np.zeros((3, 0))
            is_coplanar = bool(np.all(np.abs(bls[2]) <= flat_array_tol))
            bls = (bls / core_utils.speed_of_light).astype(rd)
            plane = rot.astype(rd)
        else:
            logger.info("Using gridded coordinates for the array. Type 1 transform will be used.")
            gpos = np.array([gridded[a] for a in antnums])
            bls = np.round(gpos[i1] - gpos[i0]).astype(int).T
            n_modes = 2 * int(np.round(np.max(np.abs(bls)))) + 1 if nbls else 1
            plane = (basis_matrix / core_utils.speed_of_light).astype(rd).T
            is_coplanar = True

        # ---- per-time rotation matrices (stage a1, host part)
        params = dict(coord_method_params or {})
        enu_mats, astrom = coords.coordinate_blocks(times, telescope_loc, coord_method, params)
        eq = params.get("eq_xyz")
        if eq is None:
            eq = coords.equatorial_unit_vectors(ra, dec)
        nchunks = max(1, min(int(nchunks), max(nsrc, 1)))
        chunk_size = int(math.ceil(nsrc / nchunks)) if nsrc else 0
        return self._plan_from_parts(
            dev, precision=precision, polarized=polarized, eps=eps, upsample_factor=upsample_factor,
            beam_spline_opts=beam_spline_opts, interpolation_function=interpolation_function, freqs=freqs,
            fluxes=fluxes if coherency is None else coherency, flux_is_coherency=coherency is not None,
            pol_sky=pol_sky, eq=eq, enu_mats=enu_mats, astrom=astrom, bls=bls, plane=plane,
            is_gridded=is_gridded, is_coplanar=is_coplanar, n_modes=n_modes, antnums=antnums,
            baselines=baselines, beam_list=beam_list, beam_idx=beam_idx, beam_coefs=beam_coefs,
            nchunks=nchunks, n_cap=max(1, int(chunk_size * source_buffer)), freq_range=freq_range)

    def _plan_from_parts(self, dev, *, precision, polarized, eps, upsample_factor, beam_spline_opts,
                         interpolation_function, freqs, fluxes, flux_is_coherency, pol_sky, eq, enu_mats, astrom,
                         bls, plane, is_gridded, is_coplanar, n_modes, antnums, baselines, beam_list, beam_idx,
                         beam_coefs, nchunks, n_cap, freq_range=None) -> SimulationPlan:
        """Uploads and device tables of a plan whose host geometry is already known: the baselines ``bls``
        (integer grid offsets of a gridded array, else seconds in the array plane), the matrix ``plane``
        that takes ENU unit vectors to those axes, per-time ENU matrices and the catalogue.  ``prepare``
        derives them from antenna positions (cpu_simulate.py:583-709); ``_evaluate_vis_chunk`` receives them
        in the reference's own argument form (cpu_simulate.py:856-884).  ``flux_is_coherency``: ``fluxes`` is
        already the coherency the reference's coordinate manager holds (0.5 I applied, cpu/utils.py:52-53)."""
        rd, cd = _NP_R[precision], _NP_C[precision]
        nfreqs, nbls, nsrc = int(freqs.size), len(baselines), int(np.shape(eq)[1])
        nbeam, nant = len(beam_list), len(antnums)
        a_index = {a: i for i, a in enumerate(antnums)}
        i0 = np.array([a_index[b[0]] for b in baselines], dtype=np.int64)
        i1 = np.array([a_index[b[1]] for b in baselines], dtype=np.int64)
        eq = np.ascontiguousarray(eq, dtype=np.float64)
        f_lo, f_hi = (0, nfreqs) if freq_range is None else (int(freq_range[0]), int(freq_range[1]))
        with torch.cuda.device(dev):
            rdt, cdt = _RDT[precision], _CDT[precision]
            eq_d = torch.as_tensor(eq).to(dev)
            # catalogue, frequency-major so that the gather through ascending src_idx coalesces
            # catalogue: uploaded as given, cast and transposed to frequency-major ON the device
            if pol_sky:
                coh_d = torch.as_tensor(np.ascontiguousarray(fluxes)).to(dev).to(cdt)
                flux_d = coh_d.permute(1, 2, 3, 0).reshape(nfreqs, 4, nsrc).contiguous()
            elif flux_is_coherency:
                flux_d = torch.as_tensor(np.ascontiguousarray(fluxes)).to(dev).to(cdt).t().contiguous()
            else:
                # Stokes I: the 0.5 I scaling and the cast (cpu/utils.py:52-53, cpu_simulate.py:622-626)
                raw = torch.as_tensor(np.ascontiguousarray(fluxes)).to(dev)
                half = raw * 0.5 if raw.is_complex() else raw.to(torch.float64) * 0.5
                del raw
                flux_d = half.to(cdt).t().contiguous()   # product in the caller's precision, then the cast
                del half
            freqs_d = torch.as_tensor(freqs.astype(np.float64)).to(dev)

            order = resolve_interpolation(interpolation_function, beam_spline_opts)
            models = []
            for b in beam_list:
                m = as_beam_model(b)
                if not polarized and m.beam_type != "power":
                    m = m.to_power()
                if polarized and m.beam_type == "power":
                    raise ValueError("polarized=True needs E-field beams")
                if hasattr(m, "interp_freq") and m.Nfreqs > 1:
                    m = m.interp_freq(freqs.astype(np.float64))
                models.append(m)
            dbeams = [DeviceBeam(m, precision, order, dev) for m in models]

            pairs, basis = [], None
            if beam_coefs is not None:
                coefs = np.asarray(beam_coefs)
                if coefs.ndim == 2:
                    coefs = np.repeat(coefs[:, :, None], nfreqs, axis=2)
                if coefs.shape != (nant, nbeam, nfreqs):
                    raise ValueError("beam_coefs must have shape (nant, K, nfreqs)")
                basis = dict(
                    coefs=torch.as_tensor(np.ascontiguousarray(coefs)).to(dev).to(cdt),
                    ant1=torch.as_tensor(i0.astype(np.int32)).to(dev),
                    ant2=torch.as_tensor(i1.astype(np.int32)).to(dev), K=nbeam, nant=nant)
                pairs.append(self._pair_table(0, 0, np.arange(nbls), np.zeros(nbls, bool), bls,
                                              is_gridded, is_coplanar, rd, dev, nbls, n_modes))
            else:
                upairs, to_bls, to_flip = GPUBeamEvaluator.prepare_beam_evaluation(antnums, baselines, beam_idx)
                for (bi, bj) in upairs:
                    idx = np.asarray(to_bls[(bi, bj)], dtype=np.int64)
                    if idx.size == 0:
                        continue
                    fl = np.asarray(to_flip[(bi, bj)], dtype=bool)
                    pairs.append(self._pair_table(int(bi), int(bj), idx, fl, bls, is_gridded,
                                                  is_coplanar, rd, dev, nbls, n_modes))

        P = 4 if polarized else 1
        fb = self.freq_batch
        if fb is None:
            ncols = max((pt.ncols for pt in pairs), default=None) if is_gridded else None
            npairs = nbeam * (nbeam + 1) // 2 if (beam_coefs is not None and nbeam <= 5) else 1
            fb = self._auto_batch(is_gridded, n_modes, P * npairs, precision, eps, float(upsample_factor), n_cap, ncols)
            # equal batches: a short last batch (a rank's 128 frequencies as 54 + 54 + 20) leaves the streams that
            # run consecutive batches side by side unevenly loaded
            nfl = max(1, f_hi - f_lo)
            if nfl > fb:
                fb = -(-nfl // -(-nfl // fb))
        return SimulationPlan(
            precision=precision, polarized=polarized, polarized_sky=pol_sky, nfeeds=2 if polarized else 1,
            eps=float(eps), upsample_factor=float(upsample_factor), use_type1=is_gridded,
            is_coplanar=is_coplanar, n_modes=n_modes, nbls=nbls, nsrc=nsrc, nchunks=nchunks, n_cap=n_cap,
            freqs_host=freqs, f_lo=f_lo, f_hi=f_hi, enu_mats=enu_mats, astrom=astrom,
            plane_mat=np.ascontiguousarray(plane, dtype=np.float64), eq_xyz=eq_d, flux=flux_d,
            freqs_dev=freqs_d, beams=dbeams, pairs=pairs, basis=basis, freq_batch=int(fb), device=dev)

    @staticmethod
    def _pair_table(bi, bj, idx, fl, bls, is_gridded, is_coplanar, rd, dev, nbls, n_modes=None) -> _PairTable:
        all_in_order = idx.size == nbls and np.array_equal(idx, np.arange(nbls))
        kmap = None if all_in_order else torch.as_tensor(idx.astype(np.int32)).to(dev)
        conj = torch.as_tensor(fl.astype(np.uint8)).to(dev) if fl.any() else None
        pt = _PairTable(bi=bi, bj=bj, nk=int(idx.size), kmap=kmap, conj=conj)
        if is_gridded:
            m = np.where(fl, -bls[:, idx], bls[:, idx])            # cpu_simulate.py:259
            pt.m1 = torch.as_tensor(np.ascontiguousarray(m[0]).astype(np.int32)).to(dev)
            pt.m2 = torch.as_tensor(np.ascontiguousarray(m[1]).astype(np.int32)).to(dev)
            pt.modes = ModeSet(m[0], m[1], n_modes)
            pt.ncols = int(np.unique(m[0]).size)
        else:
            q = np.where(fl, -bls[:, idx], bls[:, idx]).astype(rd)  # cpu_simulate.py:271
            dim = 2 if is_coplanar else 3
            pt.uvw = [torch.as_tensor(np.ascontiguousarray(q[d])).to(dev) for d in range(dim)]
            pt.ulim = [float(v) for d in range(dim) for v in (q[d].min(), q[d].max())]
        return pt

    def _auto_batch(self, type1, n_modes, P, precision, eps, upsampfac, n_cap, ncols=None) -> int:
        """Frequencies per batch.  cuFFT type-1 path: keep the batch's fine grids near the L2 budget.
        Fused type-1 path: keep the half-transformed array T (ncols x nf per frequency and product)
        inside L2 and make the strip CTAs of pass 1 fill whole waves of the 148 SMs."""
        import ctypes
        csize = 8 * precision
        by_w = max(1, ((1 << 30) if P <= 4 else (4 << 30)) // max(1, P * n_cap * csize))   # strengths buffer <= 1 GiB (4 GiB: basis pairs)
        if not type1:
            return int(max(1, min(256, by_w, 4)))                      # type-3 grids are large
        w = ctypes.c_int(0); beta = ctypes.c_double(0)
        _lib.check(_lib.lib().fv_kernel_params(eps, upsampfac, precision, ctypes.byref(w), ctypes.byref(beta)))
        nf = _next235even(max(int(upsampfac * n_modes), 2 * w.value))
        if self.type1_method == "cufft":
            by_grid = max(1, self.grid_budget_bytes // max(P * nf * nf * csize, 1))
            return int(max(1, min(256, by_grid, by_w)))
        per_f = P * (ncols or n_modes) * nf * csize
        nb_max = int(max(1, min(128, by_w, (96 << 20) // max(per_f, 1))))
        row_bytes = (nf + 1) * csize
        if row_bytes * nf <= 160 * 1024:
            strips = 1
        else:
            rows = max(1, min(32, (190 * 1024) // row_bytes))
            rows -= rows % 8 if rows > 8 else 0
            strips = -(-nf // rows)
        nc = ncols or n_modes
        xdirect = precision == 1 and strips > 1 and nf >= 2 * (24 + w.value)
        if xdirect:
            # x-direct pass 1 (csrc/type1_xdirect.cuh): strips of 24 rows, groups of 256 columns, three CTAs per SM
            strips, slots = -(-nf // 24) * -(-nc // 256), 3 * 148
        else:
            slots = 148
        # pass 2: groups of <= 8 / 16 columns, three (single) / one (double precision) CTAs per SM
        cpc = max(1, min(16, ((74 if precision == 1 else 100) * 1024 - nf * csize) // row_bytes))
        cpc -= cpc % 8 if cpc >= 8 else 0
        groups, slots2 = -(-nc // cpc), (3 if precision == 1 else 1) * 148
        best, best_cost = nb_max, float("inf")
        for nb in range(max(1, nb_max // 2), nb_max + 1):
            # whole waves of both passes per frequency; a pass-1 wave weighs ~1.5 pass-2 waves (measured on cfg2)
            cost = (1.5 * -(-strips * nb * P // slots) + (-(-groups * nb * P // slots2) if xdirect else 0)) / nb
            if cost <= best_cost + 1e-12:
                best, best_cost = nb, cost
        return int(best)

    # ------------------------------------------------------------------------------------------
    def _workspace(self, plan: SimulationPlan):
        if plan.work:
            return plan.work
        dev, prec = plan.device, plan.precision
        rdt, cdt = _RDT[prec], _CDT[prec]
        P = 4 if plan.polarized else 1
        n_cap, nb = plan.n_cap, plan.freq_batch
        w = plan.work
        w["xyz"] = torch.empty((3, n_cap), dtype=rdt, device=dev)
        w["az"] = torch.empty(n_cap, dtype=rdt, device=dev)
        w["za"] = torch.empty(n_cap, dtype=rdt, device=dev)
        w["src_idx"] = torch.empty(n_cap, dtype=torch.int32, device=dev)
        w["n_dev"] = torch.zeros(1, dtype=torch.int32, device=dev)
        w["counts"] = torch.zeros((plan.ntimes, plan.nchunks), dtype=torch.int32, device=dev)
        sb = int(_lib.lib().fv_rotate_cut_scratch_bytes(max(plan.nsrc, 1)))
        w["scratch"] = torch.empty(sb, dtype=torch.uint8, device=dev)
        w["W"] = torch.empty((nb, P, n_cap), dtype=cdt, device=dev)
        # beam tables staged in shared memory: every beam of the plan a table of order 0 / 1 on one grid
        # (the basis path stages its K <= 5 beams together; more than 5 go pair by pair through fv_weights)
        nbm = len(plan.beams)
        w["tiles_ok"] = n_cap >= 2048 and BeamTiles.supported(plan.beams[:min(nbm, 6)]) and \
            (nbm <= 6 or all(BeamTiles.supported([plan.beams[0], b]) for b in plan.beams[6:])) and \
            (plan.basis is None or 4 * (plan.basis["K"] * (plan.basis["K"] + 1) // 2) <= 64)
        if w["tiles_ok"]:
            with torch.cuda.device(dev):
                w["tiles"] = BeamTiles()
        if plan.basis is not None:
            npairs = plan.basis["K"] * (plan.basis["K"] + 1) // 2
            if 4 * npairs <= 64:                        # K <= 5: all pairs as one batched transform
                w["Wb"] = torch.empty((nb, 4 * npairs, n_cap), dtype=cdt, device=dev)
                w["vkl"] = torch.empty((nb, 4 * npairs, plan.nbls), dtype=cdt, device=dev)
            else:
                w["vkl"] = torch.empty((nb, 4, plan.nbls), dtype=cdt, device=dev)
        return w

    def run_plan(self, plan: SimulationPlan, out: torch.Tensor | None = None,
                 time_range=None, host_out: torch.Tensor | None = None, slab_hook=None) -> torch.Tensor:
        """Device-only hot loop (the GPU form of ``_evaluate_vis_chunk``, cpu_simulate.py:936-1069).
        Returns the device tensor ``(nf_local, nt, P, nbls)`` in the final output layout.

        ``out``: where to write.  Any view shaped ``(nf_local, nt, P, nbls)`` whose last two axes are
        contiguous; the frequency and time strides are free, so the same loop fills the reference's
        frequency-major array, a time-major ``(nt, nf_local, P, nbls)`` buffer (``buf.permute(1, 0, 2,
        3)``: every time slab contiguous, ready for a send) or this rank's frequency block inside the
        gathered array of all ranks (gpu/distributed.py).

        ``host_out``: a page-locked host tensor ``(nf_local, nt, P, nbls)``, contiguous.  When given,
        every finished time slab ``out[:, t]`` is copied into it on a second stream
        (``fv_memcpy2d_async``) while the next time steps are computed -- the device form of the
        reference's ``vis[tc][..., fc] = future`` scatter (cpu_simulate.py:846-847); the caller
        synchronises the device before reading it.

        ``slab_hook(to)``: called after all work of time slab ``to`` has been enqueued on the current
        stream (the sharded driver posts that slab's NCCL transfer there)."""
        dev, prec = plan.device, plan.precision
        L = _lib.lib()
        P = 4 if plan.polarized else 1
        nt_all = plan.ntimes
        t_lo, t_hi = (0, nt_all) if time_range is None else time_range
        nt = t_hi - t_lo
        nfl = plan.nf_local
        cdt = _CDT[prec]
        with torch.cuda.device(dev):
            nufft = self._nufft_plan(dev)
            st = torch.cuda.current_stream()
            if nufft.stream.cuda_stream != st.cuda_stream:
                raise RuntimeError("run_plan must run on the stream its NUFFT plan was created on")
            if out is None:
                out = torch.zeros((nfl, nt, P, plan.nbls), dtype=cdt, device=dev)
            else:
                if tuple(out.shape) != (nfl, nt, P, plan.nbls) or out.dtype != cdt:
                    raise ValueError(f"out must be {(nfl, nt, P, plan.nbls)} {cdt}")
                if out.numel() and (out.stride(3) != 1 or out.stride(2) != plan.nbls):
                    raise ValueError("the last two axes (P, nbls) of out must be contiguous")
                out.zero_()
            if nfl == 0 or nt == 0 or plan.nbls == 0 or plan.nsrc == 0:
                if host_out is not None:
                    host_out.zero_()
                if slab_hook is not None:
                    for to in range(nt):
                        slab_hook(to)
                return out
            w = self._workspace(plan)
            esz = out.element_size()
            mode = 0 if not plan.polarized else (2 if plan.polarized_sky else 1)
            pmap = _FEED_SWAP if plan.polarized else (0, 1, 2, 3)
            dim = 2 if (plan.use_type1 or plan.is_coplanar) else 3
            chunk = int(math.ceil(plan.nsrc / plan.nchunks))
            freqs64 = plan.freqs_host.astype(np.float64)
            copy_st = None
            if host_out is not None:
                if tuple(host_out.shape) != tuple(out.shape) or host_out.dtype != out.dtype \
                        or not host_out.is_contiguous():
                    raise ValueError("host_out must be a contiguous host tensor shaped and typed like the result")
                copy_st = _copy_stream(dev)
                copy_st.wait_stream(st)                      # the zero fill precedes every slab copy
            s_f, s_t = out.stride(0), out.stride(1)          # in elements
            two = bool(self.two_streams and plan.basis is None and plan.nf_local > plan.freq_batch and
                       ((plan.use_type1 and self.type1_method == "fused") or
                        (not plan.use_type1 and self.two_streams == "all")))
            if two:
                nside = max(1, int(self.side_streams))
                sides = [_side_stream(dev, i) for i in range(nside)]
                nufft2s = []
                for sd in sides:
                    with torch.cuda.stream(sd):
                        nufft2s.append(default_plan())       # the side stream's own plan (work buffers, T)
                if len(w.get("W2", [])) < nside:
                    w["W2"] = [torch.empty_like(w["W"]) for _ in range(nside)]
                ev_rc = torch.cuda.Event()
            for to, ti in enumerate(range(t_lo, t_hi)):
                for ch in range(plan.nchunks):
                    lo, hi = ch * chunk, min(plan.nsrc, (ch + 1) * chunk)
                    if lo >= hi:
                        continue
                    _lib.check(self._timed("rotate_cut", st, L.fv_rotate_cut,
                        prec, plan.eq_xyz.data_ptr(), plan.nsrc, lo, hi,
                        _lib.doubles(plan.enu_mats[ti].ravel()),
                        _lib.doubles(plan.astrom[ti]) if plan.astrom is not None else None,
                        _lib.doubles(plan.plane_mat.ravel()),
                        w["xyz"].data_ptr(), w["az"].data_ptr(), w["za"].data_ptr(),
                        w["src_idx"].data_ptr(), plan.n_cap, w["n_dev"].data_ptr(),
                        w["scratch"].data_ptr(), st.cuda_stream), "fv_rotate_cut")
                    w["counts"][ti, ch:ch + 1].copy_(w["n_dev"])
                    tiles = None
                    if self.beam_tiles and w.get("tiles_ok"):
                        tiles = w["tiles"]
                        self._timed("tile_sort", st, tiles.sort, prec, plan.beams[0], w["xyz"], w["az"], w["za"],
                                    w["src_idx"], w["n_dev"], plan.n_cap)
                        if self.beam_tiles == "sort":        # sorted live set, taps gathered from global memory
                            tiles = None
                    xlim = None
                    if not plan.use_type1:
                        import ctypes
                        lim = (ctypes.c_double * 6)()
                        _lib.check(L.fv_minmax(nufft.handle, prec, dim, w["xyz"][0].data_ptr(),
                                               w["xyz"][1].data_ptr(), w["xyz"][2].data_ptr(),
                                               w["n_dev"].data_ptr(), 0, lim), "fv_minmax")
                        xlim = [lim[i] for i in range(2 * dim)]
                        if not (xlim[0] <= xlim[1]):
                            continue                         # nothing above the horizon
                    if two:
                        ev_rc.record(st)
                        for sd in sides:
                            sd.wait_event(ev_rc)             # the live set of this (time, chunk) is ready
                    for jb, f0 in enumerate(range(plan.f_lo, plan.f_hi, plan.freq_batch)):
                        nb = min(plan.freq_batch, plan.f_hi - f0)
                        scale = freqs64[f0:f0 + nb]
                        obase = out.data_ptr() + ((f0 - plan.f_lo) * s_f + to * s_t) * esz
                        sb_, sp_ = s_f, plan.nbls
                        if two and (jb % (nside + 1)) and tiles is None:
                            si = jb % (nside + 1) - 1
                            side, nufft2, W2 = sides[si], nufft2s[si], w["W2"][si]
                            for pt in plan.pairs:
                                self._timed("weights", side, launch_weights, prec, mode, plan.beams[pt.bi],
                                            plan.beams[pt.bj], w["az"], w["za"], w["src_idx"], w["n_dev"], plan.n_cap,
                                            plan.freqs_dev, f0, nb, plan.flux, plan.nsrc, W2, None, side)
                                epi = _lib.make_epilogue(
                                    obase, sb_, sp_, pmap, pt.kmap.data_ptr() if pt.kmap is not None else 0,
                                    pt.conj.data_ptr() if pt.conj is not None else 0, accumulate=ch > 0)
                                self._nufft_batch(plan, w, nufft2, pt, dim, xlim, scale, nb, epi, W=W2[:nb])
                            continue
                        if plan.basis is not None:
                            self._basis_batch(plan, w, nufft, mode, dim, xlim, f0, nb, scale, obase, sb_, sp_, st,
                                              tiles=tiles)
                            continue
                        for pt in plan.pairs:
                            if tiles is not None:
                                bb = [plan.beams[pt.bi]] if pt.bi == pt.bj else [plan.beams[pt.bi], plan.beams[pt.bj]]
                                self._timed("weights", st, tiles.weights, prec, mode, bb, False, w["az"], w["za"],
                                            w["src_idx"], plan.n_cap, plan.freqs_dev, f0, nb, plan.flux, plan.nsrc,
                                            w["W"])
                            else:
                                self._timed("weights", st, launch_weights, prec, mode, plan.beams[pt.bi],
                                            plan.beams[pt.bj], w["az"], w["za"], w["src_idx"], w["n_dev"], plan.n_cap,
                                            plan.freqs_dev, f0, nb, plan.flux, plan.nsrc, w["W"], None, st)
                            epi = _lib.make_epilogue(
                                obase, sb_, sp_, pmap, pt.kmap.data_ptr() if pt.kmap is not None else 0,
                                pt.conj.data_ptr() if pt.conj is not None else 0,
                                # every baseline belongs to exactly one beam pair and `out` starts at zero: only the
                                # later source chunks add to what is there (cpu_simulate.py:1069 `vis[...] +=`)
                                accumulate=ch > 0)
                            self._nufft_batch(plan, w, nufft, pt, dim, xlim, scale, nb, epi)
                    if two:
                        for sd in sides:
                            st.wait_stream(sd)               # chunk complete; the live set may be rewritten
                if copy_st is not None:
                    self._stream_slab(out, host_out, to, st, copy_st)
                if slab_hook is not None:
                    slab_hook(to)
            if copy_st is not None:
                st.wait_stream(copy_st)                      # `out` may be reused once the copies are done
            return out

    @staticmethod
    def _stream_slab(out, host_out, to, st, copy_st):
        """Enqueue the D2H of time slab ``to`` (nf rows of P * nbls elements; rows ``out.stride(0)``
        apart on the device, ``nt`` slabs apart on the host) behind the work enqueued so far."""
        esz = out.element_size()
        nfl, nt = out.shape[0], out.shape[1]
        slab = out.shape[2] * out.shape[3] * esz
        copy_st.wait_stream(st)
        _lib.check(_lib.lib().fv_memcpy2d_async(
            host_out.data_ptr() + to * slab, nt * slab, out.data_ptr() + to * out.stride(1) * esz,
            out.stride(0) * esz, slab, nfl, 0, copy_st.cuda_stream), "fv_memcpy2d_async")

    def _nufft_batch(self, plan, w, nufft, pt, dim, xlim, scale, nb, epi, W=None):
        W = w["W"][:nb] if W is None else W
        if plan.use_type1 and self.type1_method == "fused":
            nufft.type1_fused(plan.precision, w["xyz"][0], w["xyz"][1], w["n_dev"], scale, W, pt.modes,
                              plan.eps, plan.upsample_factor, epi)
        elif plan.use_type1:
            nufft.type1(plan.precision, w["xyz"][0], w["xyz"][1], w["n_dev"], scale, W, plan.n_modes,
                        pt.m1, pt.m2, plan.eps, plan.upsample_factor, epi)
        else:
            nufft.type3(plan.precision, dim, w["xyz"], w["n_dev"], xlim, pt.uvw, pt.ulim, scale, W,
                        plan.eps, plan.upsample_factor, epi)

    def _basis_batch(self, plan, w, nufft, mode, dim, xlim, f0, nb, scale, obase, sb_, sp_, st, tiles=None):
        """K (K + 1) / 2 transforms over all baselines + contraction (cpu_simulate.py:416-468): every basis
        beam is evaluated once per (source, frequency) (``fv_weights_basis``), the pair products are the
        strengths of ONE batched transform with 4 K (K + 1) / 2 components on the fused type-1 path (the
        other paths take them four at a time), and one kernel contracts all pairs."""
        import ctypes
        L = _lib.lib()
        b = plan.basis
        pt = plan.pairs[0]
        K = b["K"]
        npairs = K * (K + 1) // 2
        if "Wb" not in w:
            return self._basis_batch_pairs(plan, w, nufft, mode, dim, xlim, f0, nb, scale, obase, sb_, sp_, st)
        Wb, vkl = w["Wb"][:nb], w["vkl"]
        if tiles is not None:
            self._timed("weights", st, tiles.weights, plan.precision, mode, plan.beams[:K], True, w["az"], w["za"],
                        w["src_idx"], plan.n_cap, plan.freqs_dev, f0, nb, plan.flux, plan.nsrc, Wb)
        else:
            self._timed("weights", st, launch_weights_basis, plan.precision, mode, plan.beams[:K], w["az"], w["za"],
                        w["src_idx"], w["n_dev"], plan.n_cap, plan.freqs_dev, f0, nb, plan.flux, plan.nsrc, Wb, st)
        if plan.use_type1 and self.type1_method == "fused":
            epi = _lib.make_epilogue(vkl.data_ptr(), vkl.stride(0), vkl.stride(1), _FEED_SWAP)
            self._nufft_batch(plan, w, nufft, pt, dim, xlim, scale, nb, epi, W=Wb)
        else:
            for q in range(npairs):
                w["W"][:nb].copy_(Wb[:, 4 * q:4 * q + 4])
                epi = _lib.make_epilogue(vkl.data_ptr() + 4 * q * vkl.stride(1) * vkl.element_size(),
                                         vkl.stride(0), vkl.stride(1), _FEED_SWAP)
                self._nufft_batch(plan, w, nufft, pt, dim, xlim, scale, nb, epi)
        epo = _lib.make_epilogue(obase, sb_, sp_, (0, 1, 2, 3), accumulate=True)
        _lib.check(L.fv_basis_contract_all(
            plan.precision, vkl.data_ptr(), nb, plan.nbls, b["coefs"].data_ptr(), b["nant"], K,
            plan.freqs_host.size, f0, b["ant1"].data_ptr(), b["ant2"].data_ptr(),
            ctypes.byref(epo), st.cuda_stream), "fv_basis_contract_all")

    def _basis_batch_pairs(self, plan, w, nufft, mode, dim, xlim, f0, nb, scale, obase, sb_, sp_, st):
        """More than five basis beams (over 64 batched components): one transform per pair."""
        import ctypes
        L = _lib.lib()
        b = plan.basis
        pt = plan.pairs[0]
        vkl = w["vkl"]
        for k in range(b["K"]):
            for l in range(k, b["K"]):
                launch_weights(plan.precision, mode, plan.beams[k], plan.beams[l], w["az"], w["za"],
                               w["src_idx"], w["n_dev"], plan.n_cap, plan.freqs_dev, f0, nb, plan.flux,
                               plan.nsrc, w["W"], None, st)
                epi = _lib.make_epilogue(vkl.data_ptr(), vkl.stride(0), vkl.stride(1), _FEED_SWAP)
                self._nufft_batch(plan, w, nufft, pt, dim, xlim, scale, nb, epi)
                epo = _lib.make_epilogue(obase, sb_, sp_, (0, 1, 2, 3), accumulate=True)
                _lib.check(L.fv_basis_contract(
                    plan.precision, vkl.data_ptr(), nb, plan.nbls, b["coefs"].data_ptr(), b["nant"], b["K"],
                    plan.freqs_host.size, f0, k, l, b["ant1"].data_ptr(), b["ant2"].data_ptr(),
                    ctypes.byref(epo), st.cuda_stream), "fv_basis_contract")

    def check_source_buffer(self, plan: SimulationPlan):
        """Raise like matvis' ``select_chunk`` when a chunk overflowed its buffer
        (reference call site cpu_simulate.py:940)."""
        if not plan.work:
            return
        counts = plan.work["counts"].cpu().numpy()
        if (counts < 0).any():
            raise ValueError(
                f"source_buffer too small: {int(-counts.min())} sources above the horizon in one chunk "
                f"but the buffer holds {plan.n_cap}. Increase source_buffer.")

    # ------------------------------------------------------------------------------------------
    def simulate(self, ants, freqs, fluxes, beam_list, ra, dec, times, telescope_loc,
                 baselines=None, beam_idx=None, precision=2, polarized=False, eps=None,
                 upsample_factor=2, beam_spline_opts=None, flat_array_tol=1e-6,
                 interpolation_function="az_za_map_coordinates", nprocesses=1, nthreads=None,
                 coord_method="CoordinateRotationERFA", coord_method_params=None,
                 force_use_ray=False, force_use_type3=False, trace_mem=False,
                 enable_memory_monitor=False, nchunks=1, source_buffer=1.0, beam_coefs=None):
        """Simulate visibilities on the GPU; same parameters and return value as the CPU engine
        (cpu_simulate.py:537-569, return at :850-854).  ``nprocesses``, ``nthreads``,
        ``force_use_ray``, ``trace_mem`` and ``enable_memory_monitor`` are CPU-only knobs and are
        accepted and ignored.  ``coord_method`` / ``coord_method_params`` select the coordinate model
        (core/coords.py ``coordinate_blocks``; unknown names raise ``KeyError`` as in the reference);
        ``interpolation_function`` is validated by ``gpu/beams.py`` (never silently ignored)."""
        plan = self.prepare(ants, freqs, fluxes, beam_list, ra, dec, times, telescope_loc,
                            baselines=baselines, beam_idx=beam_idx, precision=precision,
                            polarized=polarized, eps=eps, upsample_factor=upsample_factor,
                            beam_spline_opts=beam_spline_opts, flat_array_tol=flat_array_tol,
                            coord_method_params=coord_method_params, force_use_type3=force_use_type3,
                            nchunks=nchunks, source_buffer=source_buffer, beam_coefs=beam_coefs,
                            coord_method=coord_method, interpolation_function=interpolation_function)
        host = self._pinned_result(plan)
        out = self.run_plan(plan, host_out=host)
        if host is None:
            self.check_source_buffer(plan)
            return self.finish(plan, out)
        self.check_source_buffer(plan)                       # reads the live counts back: synchronises
        torch.cuda.current_stream(plan.device).synchronize()
        return self._shape_result(plan, host.numpy())

    @staticmethod
    def _pinned_result(plan: SimulationPlan):
        """Page-locked host block for the result (torch's caching host allocator reuses it across
        calls); None when page-locking is refused (ulimit / memory pressure)."""
        P = 4 if plan.polarized else 1
        try:
            return torch.empty((plan.nf_local, plan.ntimes, P, plan.nbls), dtype=_CDT[plan.precision],
                               pin_memory=True)
        except RuntimeError:
            return None

    @staticmethod
    def _shape_result(plan: SimulationPlan, res: np.ndarray) -> np.ndarray:
        nf, nt = res.shape[0], res.shape[1]
        if plan.polarized:
            return res.reshape(nf, nt, 2, 2, plan.nbls)
        return res.reshape(nf, nt, plan.nbls)

    @staticmethod
    def finish(plan: SimulationPlan, out: torch.Tensor) -> np.ndarray:
        """D2H + final shape ``(nf, nt, 2, 2, nbls)`` / ``(nf, nt, nbls)`` (cpu_simulate.py:850-854).
        The copy lands in page-locked host memory (torch's caching host allocator reuses the block
        across calls), which is ~20x faster than a pageable D2H for the multi-GB result; the returned
        array is a view of that block and keeps it alive."""
        try:
            host = torch.empty(out.shape, dtype=out.dtype, pin_memory=True)
            host.copy_(out, non_blocking=True)
            torch.cuda.current_stream(out.device).synchronize()
        except RuntimeError:          # page-locking refused (ulimit / memory pressure): plain copy
            host = out.cpu()
        res = host.numpy()
        nf, nt = res.shape[0], res.shape[1]
        if plan.polarized:
            return res.reshape(nf, nt, 2, 2, plan.nbls)
        return res.reshape(nf, nt, plan.nbls)

    # ------------------------------------------------------------------------------------------
    def _plan_from_chunk_args(self, beam_list, coord_mgr, rotation_matrix, antnums, baselines, bls, freqs,
                              complex_dtype, nfeeds, beam_idx, polarized, polarized_sky_model, eps, upsample_factor,
                              beam_spline_opts, interpolation_function, is_coplanar, use_type1, basis_matrix,
                              type1_n_modes, nchunks, beam_coefs) -> SimulationPlan:
        """Device plan from the reference's chunk-evaluator arguments (cpu_simulate.py:856-884)."""
        dev = self._device()
        m = coords.manager_inputs(coord_mgr)
        precision = 1 if (complex_dtype is not None and np.dtype(complex_dtype) == np.complex64) else \
            (2 if complex_dtype is not None else int(getattr(coord_mgr, "precision", 2)))
        rd = _NP_R[precision]
        if nfeeds is not None and int(nfeeds) != (2 if polarized else 1):
            raise ValueError("nfeeds must be 2 for polarized and 1 for unpolarized simulations")
        if eps is None:
            eps = default_accuracy_dict[precision]
        freqs = np.atleast_1d(np.asarray(freqs)).astype(rd, copy=False)
        baselines = [tuple(b) for b in baselines]
        if antnums is None:
            antnums = sorted({a for b in baselines for a in b})
        antnums = list(antnums)
        beam_idx = core_utils.validate_beam_idx(beam_idx, beam_coefs, len(beam_list), len(antnums))
        bls = np.asarray(bls)
        # ENU -> the axes the baselines are expressed in (cpu_simulate.py:961-965): array-plane rotation,
        # then the lattice basis of a gridded array (already divided by c by the caller, :676-677)
        plane = np.eye(3) if rotation_matrix is None else np.asarray(rotation_matrix, dtype=np.float64)
        if basis_matrix is not None:
            plane = np.asarray(basis_matrix, dtype=np.float64).T @ plane
        if use_type1:
            bls = np.round(bls).astype(int)
            n_modes = int(type1_n_modes) if type1_n_modes is not None else 2 * int(np.abs(bls).max()) + 1
        else:
            bls, n_modes = bls.astype(rd), None
        flux = m["flux"]
        pol_sky = bool(polarized_sky_model) or flux.ndim == 4
        enu_mats, astrom = coords.coordinate_blocks(m["times"], m["telescope_loc"], m["method"], m["params"])
        eq = m["params"].get("eq_xyz")
        if eq is None:
            eq = coords.equatorial_unit_vectors(m["ra"].astype(rd), m["dec"].astype(rd))
        nsrc = int(m["ra"].size)
        # the manager's chunk size and nchunks describe the same split (cpu_simulate.py:691, :939)
        nchunks = max(int(nchunks), -(-nsrc // max(m["chunk_size"], 1))) if nsrc else 1
        nchunks = max(1, min(nchunks, max(nsrc, 1)))
        chunk_size = int(math.ceil(nsrc / nchunks)) if nsrc else 0
        return self._plan_from_parts(
            dev, precision=precision, polarized=polarized, eps=eps, upsample_factor=upsample_factor,
            beam_spline_opts=beam_spline_opts, interpolation_function=interpolation_function, freqs=freqs,
            fluxes=flux, flux_is_coherency=True, pol_sky=pol_sky, eq=eq, enu_mats=enu_mats, astrom=astrom,
            bls=bls, plane=plane, is_gridded=bool(use_type1), is_coplanar=bool(is_coplanar) or bool(use_type1),
            n_modes=n_modes, antnums=antnums, baselines=baselines, beam_list=beam_list, beam_idx=beam_idx,
            beam_coefs=beam_coefs, nchunks=nchunks, n_cap=max(1, int(chunk_size * m["source_buffer"])))

    def _evaluate_vis_chunk(self, time_idx, freq_idx, beam_list=None, coord_mgr=None, rotation_matrix=None,
                            antnums=None, baselines=None, bls=None, freqs=None, complex_dtype=None, nfeeds=None,
                            beam_idx=None, polarized=False, polarized_sky_model=False, eps=None,
                            upsample_factor=2, beam_spline_opts=None,
                            interpolation_function="az_za_map_coordinates", n_threads=1, is_coplanar=False,
                            use_type1=False, basis_matrix=None, type1_n_modes=None, trace_mem=False, nchunks=1,
                            beam_coefs=None, plan: SimulationPlan = None):
        """One (time-slice, frequency-slice) block as ``(nt_here, nbls, nfeed, nfeed, nf_here)``; the CPU
        engine's signature (cpu_simulate.py:856-884, called directly by tests/test_cpu_simulate.py:1068-1087).

        Two call forms.  The reference's: ``coord_mgr`` (``core.coords.CoordinateRotation`` or a matvis
        manager: catalogue, times, site, source positions, chunking), the baselines ``bls`` in the array
        plane with ``rotation_matrix`` (type 3) or as integer grid offsets with ``basis_matrix`` /
        ``type1_n_modes`` (type 1), exactly as ``simulate`` hands them to its workers; a device plan is
        built from them.  Or ``plan=<SimulationPlan from prepare()>``: the GPU engine's own unit of state,
        which skips the uploads.  ``n_threads`` and ``trace_mem`` are CPU knobs and are ignored."""
        if plan is None:
            if coord_mgr is None or beam_list is None or bls is None or freqs is None or baselines is None:
                raise TypeError("_evaluate_vis_chunk needs either plan=<SimulationPlan> or the reference's "
                                "arguments (beam_list, coord_mgr, rotation_matrix, antnums, baselines, bls, freqs, ...)")
            plan = self._plan_from_chunk_args(
                beam_list, coord_mgr, rotation_matrix, antnums, baselines, bls, freqs, complex_dtype, nfeeds,
                beam_idx, polarized, polarized_sky_model, eps, upsample_factor, beam_spline_opts,
                interpolation_function, is_coplanar, use_type1, basis_matrix, type1_n_modes, nchunks, beam_coefs)
        nt_all, nf_all = plan.ntimes, plan.freqs_host.size
        ts = range(nt_all)[time_idx]
        fs = range(nf_all)[freq_idx]
        if len(ts) == 0 or len(fs) == 0:
            return np.zeros((len(ts), plan.nbls, plan.nfeeds, plan.nfeeds, len(fs)), _NP_C[plan.precision])
        if ts.step != 1 or fs.step != 1:
            raise ValueError("time_idx / freq_idx must be contiguous slices")
        sub = SimulationPlan(**{**plan.__dict__, "f_lo": fs.start, "f_hi": fs.stop, "work": plan.work})
        out = self.run_plan(sub, time_range=(ts.start, ts.stop))
        self.check_source_buffer(sub)
        res = out.cpu().numpy().reshape(len(fs), len(ts), plan.nfeeds, plan.nfeeds, plan.nbls)
        return np.ascontiguousarray(np.transpose(res, (1, 4, 2, 3, 0)))
