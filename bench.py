#!/usr/bin/env python
"""Headline benchmark: visibility terms / second (Nsrc * Nbl * Nfreq * Ntime / s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload cfg2]

A "step" is one pass of the hot path (``GPUSimulationEngine.run_plan``: rotate + horizon cut, beam +
coherency weights, batched NUFFT, epilogue) over the whole workload with every input already resident
in HBM.  Default workload = BASELINE.json configs[1]: HERA-350-like gridded hex array (type-1 path),
10 000 point sources, 1024 frequencies 100-200 MHz, 60 times, unpolarised Airy beam, single precision.
At N > 1 (torchrun, one rank per GPU) the SAME workload is sharded over frequency (contiguous blocks, all
times per rank) and the finished time slabs are gathered on rank 0 over NCCL while later slabs compute
(fftvis_b200/gpu/distributed.py): strong scaling, the gather inside the timed region.  ``--scaling weak``
keeps round 1's replica mode (every rank the whole workload on its own block of times, no collective).

One JSON line on stdout (rank 0).  ``e2e`` = the same metric through ``fftvis_b200.simulate_vis`` with
host numpy buffers (planning, H2D, compute, D2H inside the timed region).  ``roofline`` = the dominant
kernel of this library (per-launch CUDA events recorded live in the timed region).  ``cpu_baseline`` /
``--impl reference`` = the CPU restatement of the reference pipeline (oracle/; finufft, matvis and
pyuvdata cannot be installed offline, so the reference itself cannot run) on a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

START_JD = 2459845.0
CADENCE_S = 10.0


# --------------------------------------------------------------------------------------------
# workloads (BASELINE.json configs)
# --------------------------------------------------------------------------------------------
def make_workload(name: str, nfreq=None, ntimes=None, nsrc=None, time_block: int = 0, freq_range=None):
    """``freq_range=(lo, hi)``: generate only that block of the workload's frequencies (one rank's shard
    of the frequency-sharded run; every per-frequency input -- fluxes, beam tables, basis coefficients --
    is a function of the frequency alone, so a shard equals the slice of the full workload)."""
    from fftvis_b200 import AiryBeam, GaussianBeam, HERA_LOCATION, synth
    w = dict(name=name, telescope_loc=HERA_LOCATION, kwargs={})
    fsl = slice(None) if freq_range is None else slice(int(freq_range[0]), int(freq_range[1]))
    if name == "cfg1":
        nfreq, ntimes, nsrc = nfreq or 2, ntimes or 1, nsrc or 100
        w.update(ants=synth.hex_rows((3, 4, 3)), beam=GaussianBeam(diameter=14.0), precision=2, polarized=False,
                 desc="tests-scale: 10-antenna hex, 100 sources, 2 freqs, 1 time, Gaussian beam, f64")
        freqs = np.linspace(100e6, 110e6, nfreq)[fsl]
        sky = synth.random_sky(nsrc, freqs, seed=42)
    elif name == "cfg2":
        nfreq, ntimes, nsrc = nfreq or 1024, ntimes or 60, nsrc or 10000
        w.update(ants=synth.hera350_like(), beam=AiryBeam(diameter=14.0), precision=1, polarized=False,
                 desc="HERA-350-like gridded hex (type-1 path), 10k GLEAM-like point sources, "
                      "1024 freqs 100-200 MHz, 60 times, unpolarized Airy beam, single precision")
        freqs = np.linspace(100e6, 200e6, nfreq)[fsl]
        sky = synth.random_sky(nsrc, freqs, seed=42, kind="gleam")
    elif name == "cfg3":
        nfreq, ntimes, nsrc = nfreq or 1024, ntimes or 60, nsrc or 100000
        freqs = np.linspace(100e6, 200e6, nfreq)[fsl]
        w.update(ants=synth.hex_array(11), beam=synth.synthetic_uvbeam(freqs, naz=360, nza=181), precision=2,
                 polarized=True, kwargs=dict(beam_spline_opts={"order": 1}),
                 desc="HERA-331 polarized: synthetic UVBeam E-field on az/za grid, 4 pol products, "
                      "100k sources, 1024 freqs, 60 times, f64")
        sky = synth.random_sky(nsrc, freqs, seed=42, kind="gleam")
    elif name == "cfg4":
        nfreq, ntimes, nsrc = nfreq or 512, ntimes or 1, nsrc or 3145728
        ants = synth.random_array(256, radius=150.0, zspan=2.0, seed=42)
        w.update(ants=ants, beam=GaussianBeam(diameter=14.0), precision=2, polarized=False,
                 kwargs=dict(baselines=synth.all_baselines(ants)),
                 desc="non-gridded random 256-antenna layout (32640 baselines), 3-D type 3, diffuse sky "
                      "(3.1M pixels, ~1.5M above horizon), 512 freqs, f64")
        freqs = np.linspace(100e6, 200e6, nfreq)[fsl]
        sky = synth.random_sky(nsrc, freqs, seed=42, kind="diffuse")
    elif name == "cfg5":
        nfreq, ntimes, nsrc = nfreq or 1024, ntimes or 120, nsrc or 100000
        freqs = np.linspace(100e6, 200e6, nfreq)[fsl]
        ants = synth.hex_array(7)
        ants[len(ants)] = np.array([7 * synth.HEX_SPACING, 0.0, 0.0])          # 127 + 1 antennas, on the lattice
        K = 5
        basis = [synth.synthetic_uvbeam(freqs, naz=360, nza=181, seed=s, perturb=0.3) for s in range(K)]
        for b in basis:
            b.data_array = b.data_array.real.astype(complex)      # real basis beams (reference's upper-triangle trick)
        rng = np.random.default_rng(42)
        coefs = (rng.normal(size=(len(ants), K, nfreq)) + 1j * rng.normal(size=(len(ants), K, nfreq)))[:, :, fsl]
        w.update(ants=ants, beam=basis, precision=2, polarized=True,
                 kwargs=dict(baselines=synth.all_baselines(ants, autos=True), beam_coefs=coefs,
                             beam_spline_opts={"order": 1}),
                 desc="per-antenna beams via 5-term beam-basis decomposition, 128 antennas, polarized, 100k "
                      "sources, 1024 freqs x 120 times")
        sky = synth.random_sky(nsrc, freqs, seed=42, kind="gleam")
    else:
        raise SystemExit(f"unknown workload {name}")
    t0 = START_JD + time_block * ntimes * CADENCE_S / 86400.0
    w.update(freqs=freqs, times=t0 + np.arange(ntimes) * CADENCE_S / 86400.0, ra=sky[0], dec=sky[1],
             fluxes=sky[2], nfreq=nfreq, ntimes=ntimes, nsrc=nsrc)
    return w


def n_baselines(w) -> int:
    if "baselines" in w["kwargs"]:
        return len(w["kwargs"]["baselines"])
    from fftvis_b200.core import utils
    return len(utils.get_pos_reds(w["ants"], include_autos=True))


def terms(w, nbls, nfreq=None, ntimes=None) -> float:
    return float(w["nsrc"]) * nbls * (nfreq or w["nfreq"]) * (ntimes or w["ntimes"])


# --------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi during the timed region)
# --------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.gpu)], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in Path(self.path).read_text().splitlines():
            p = [x.strip() for x in line.split(",")]
            if len(p) < 8:
                continue
            try:
                sm.append(float(p[1])); mx.append(float(p[2])); power.append(float(p[3]))
            except ValueError:
                continue
            for nm, v in zip(names, p[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        os.unlink(self.path)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "power_w_max": float(max(power)),
                "samples": len(sm), "reasons": sorted(reasons)}


# --------------------------------------------------------------------------------------------
# CPU arm (oracle = restatement of the reference pipeline)
# --------------------------------------------------------------------------------------------
def cpu_sample(w, nbls, budget_s: float = 15.0):
    """Time oracle.pipeline.simulate_cpu on a bounded (time, frequency) sample of the workload with
    every host core; returns (terms/s, description, cores, seconds)."""
    from oracle import nufft_cpu, pipeline
    # every host core, whatever OMP_NUM_THREADS says (torchrun sets it to 1): the oracle's OpenMP
    # regions take an explicit thread count
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    beam = w["beam"] if w["polarized"] else w["beam"].to_power()
    beam_list = beam if isinstance(beam, list) else [beam]
    base = dict(precision=w["precision"], polarized=w["polarized"], nthreads=cores, **w["kwargs"])

    def run(nf, nt):
        fs = slice(0, nf)
        t0 = time.perf_counter()
        pipeline.simulate_cpu(w["ants"], w["fluxes"], w["ra"], w["dec"], w["freqs"], w["times"], beam_list,
                              w["telescope_loc"], freq_slice=fs, time_slice=slice(0, nt), **base)
        return time.perf_counter() - t0

    # grow the sample geometrically until one run costs between budget/4 and budget seconds; that
    # last run is the measurement (its one-off planning share is then small)
    total = w["nfreq"] * w["ntimes"]
    units = 2
    while True:
        nf = int(min(w["nfreq"], units))
        nt = int(max(1, min(w["ntimes"], units // nf)))
        dt = run(nf, nt)
        if dt >= budget_s / 4 or nf * nt >= total:
            break
        units = int(min(total, max(units * 2, units * min(8.0, 0.6 * budget_s / max(dt, 1e-3)))))
    val = terms(w, nbls, nf, nt) / dt
    return val, f"{nf} of {w['nfreq']} freqs x {nt} of {w['ntimes']} times of the same workload, {dt:.1f} s", cores, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w = apply_overrides(make_workload(args.workload, args.nfreq, args.ntimes, args.nsrc), args)
    nbls = n_baselines(w)
    vals, secs, sample, cores = [], [], "", 1
    for i in range(args.warmup + args.steps):
        v, sample, cores, dt = cpu_sample(w, nbls, budget_s=args.cpu_budget)
        if i >= args.warmup:
            vals.append(v); secs.append(dt)
    val = float(np.mean(vals))
    line = {
        "impl": "reference", "metric": "vis terms/sec (Nsrc*Nbl*Nfreq*Ntime/s)", "value": val, "unit": "terms/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean(secs)),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32" if w["precision"] == 1 else "f64", "data": "synthetic",
        "config": {"workload": f"{w['name']}: {w['desc']}", "nsrc": w["nsrc"], "nbls": nbls, "nfreq": w["nfreq"],
                   "ntimes": w["ntimes"]},
        "cpu_baseline": {"value": val, "unit": "terms/s", "cores": cores, "kind": "port", "sample": sample,
                         "note": "CPU restatement of the reference pipeline (finufft/matvis/pyuvdata unavailable offline)"},
        "e2e": {"value": val, "unit": "terms/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------
def kernel_geometry(plan):
    """(w, nf) of the type-1 fine grid of ``plan``."""
    import ctypes
    from fftvis_b200.gpu import _lib
    w_, beta = ctypes.c_int(0), ctypes.c_double(0)
    _lib.lib().fv_kernel_params(plan.eps, plan.upsample_factor, plan.precision, ctypes.byref(w_), ctypes.byref(beta))
    nf = _lib.lib().fv_next235even(max(int(plan.upsample_factor * plan.n_modes), 2 * w_.value)) if plan.n_modes else 0
    return w_.value, int(nf)


def spread_alg_bytes(plan, n_live: float, nb: float) -> float:
    """Algorithmic bytes of ONE launch of the type-1 spreader (the per-unit
    figure of SURVEY.md section 8(d) restricted to this kernel, times the ``nb`` frequencies of a launch):
    NU coordinates and strengths read once, the fine grid written once.  The fused kernel keeps the
    grid in shared memory, so its real DRAM traffic (``traffic``) is far below this figure."""
    r = 4 * plan.precision
    c = 2 * r
    P = 4 if plan.polarized else 1
    _, nf = kernel_geometry(plan)
    return nb * (n_live * (2 * r + P * c) + P * c * nf * nf)


def run_gpu(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (fftvis_b200 has no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        # keep stdout to the one JSON line: NCCL prints its version banner to stdout at NCCL_DEBUG=VERSION
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    import fftvis_b200
    from fftvis_b200.gpu import GPUSimulationEngine, _lib
    from fftvis_b200.gpu import distributed as fvdist

    strong = world > 1 and args.scaling == "strong"
    full_nf = make_workload_nfreq(args)
    shards = fvdist.shard_frequencies(full_nf, world) if strong else None
    if strong:
        w = make_workload(args.workload, args.nfreq, args.ntimes, args.nsrc, freq_range=shards[rank])
    else:
        w = make_workload(args.workload, args.nfreq, args.ntimes, args.nsrc, time_block=rank)
    w = apply_overrides(w, args)
    nbls = n_baselines(w)
    beam = w["beam"] if w["polarized"] else w["beam"].to_power()
    beam_list = beam if isinstance(beam, list) else [beam]
    eng = GPUSimulationEngine(freq_batch=args.freq_batch, type1_method=args.type1_method)
    eng.two_streams = False if args.one_stream else (os.environ.get("FV_TWO_STREAMS") or True)
    eng.side_streams = int(os.environ.get("FV_SIDE_STREAMS", eng.side_streams))
    if args.no_beam_tiles:
        eng.beam_tiles = False
    elif os.environ.get("FV_BEAM_TILES"):                    # "sort" (default) or "smem"
        eng.beam_tiles = True if os.environ["FV_BEAM_TILES"] == "smem" else os.environ["FV_BEAM_TILES"]
    plan = eng.prepare(w["ants"], w["freqs"], w["fluxes"], beam_list, w["ra"], w["dec"], w["times"],
                       w["telescope_loc"], precision=w["precision"], polarized=w["polarized"], **w["kwargs"])
    P = 4 if plan.polarized else 1
    cdt = torch.complex64 if plan.precision == 1 else torch.complex128
    nufft = eng._nufft_plan(plan.device)
    swork = {}
    out = None
    if not strong:
        out = torch.zeros((plan.nf_local, plan.ntimes, P, plan.nbls), dtype=cdt, device=plan.device)

    def step():
        if strong:
            # this rank's frequency block; every finished time slab is sent to rank 0 (NCCL send / recv
            # into its place in the gathered array) while the next slabs compute
            fvdist.run_sharded(eng, plan, shards, dst=0, work=swork)
        else:
            eng.run_plan(plan, out=out)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    eng.check_source_buffer(plan)
    barrier()
    # Live in the timed region: CUDA events around every launch of the dominant stage (the spreader / pass 1 of
    # type 1; spreader + inner FFT of type 3), which feed ``roofline``.  The events of the minor stages (prep,
    # gather, weights, rotate ...) would add ~9 k event records per cfg2 step to the timed region, so their
    # breakdown comes from ONE extra step after it (``stages_from``).
    dominant = (1 << 1) if plan.use_type1 else ((1 << 1) | (1 << 2) | (1 << 4))
    plans = eng.nufft_plans(plan.device)                 # main stream + the side stream of alternate batches

    def merged_stage_times():
        tot = {}
        for pl in plans:
            for k, (ms_, n_) in pl.stage_times().items():
                a_, b_ = tot.get(k, (0.0, 0))
                tot[k] = (a_ + ms_, b_ + n_)
        return tot

    for pl in plans:
        pl.set_option("timing_mask", dominant)
        pl.set_timing(True)
        pl.reset_timing()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = _lib.lib().fv_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = _lib.lib().fv_launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    stages = merged_stage_times()
    # one extra step with every stage timed
    for pl in plans:
        pl.set_option("timing_mask", (1 << 6) - 1)
        pl.reset_timing()
    eng.time_stages = True
    eng.stage_times()
    overlapped = bool(eng.two_streams)
    eng.two_streams = False                              # kernels one after the other: clean per-stage times
    xe0, xe1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    xe0.record()
    step()
    xe1.record()
    barrier()
    eng.two_streams = overlapped
    extra_ms = xe0.elapsed_time(xe1)
    stages_all = merged_stage_times()
    stages_all.update(eng.stage_times())                 # rotate + cut, beam-tile sort, weights
    eng.time_stages = False
    for pl in plans:
        pl.set_timing(False)
    tms = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms_max = float(tms.item())
    work_units = 1 if strong else world                  # strong: ONE workload over all ranks
    total_terms = terms(w, nbls) * work_units * args.steps
    value = total_terms / (ms_max * 1e-3)
    # live (above-horizon) sources, read back from the run: the horizon cut's own counts
    counts = plan.work["counts"].cpu().numpy() if plan.work else np.zeros((1, 1))
    n_live = float(counts.sum(axis=1).mean())
    gathered_bytes = 0
    if strong:
        esz = 8 * w["precision"]
        gathered_bytes = int(sum(hi - lo for r, (lo, hi) in enumerate(shards) if r != 0) * plan.ntimes * P * plan.nbls * esz)

    # ---- end to end through the public API with host buffers ------------------------------------
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    call = dict(ants=w["ants"], fluxes=w["fluxes"], ra=w["ra"], dec=w["dec"], freqs=w["freqs"], times=w["times"],
                telescope_loc=w["telescope_loc"], precision=w["precision"], polarized=w["polarized"], **w["kwargs"])
    fb, plan_nf_local = plan.freq_batch, plan.nf_local
    del out, plan
    swork.clear()
    torch.cuda.empty_cache()
    if strong:
        eng_e = GPUSimulationEngine(freq_batch=args.freq_batch, type1_method=args.type1_method)
        # every rank streams its block into a shared page-locked host array (all PCIe links); the variant that
        # gathers on rank 0's GPU by NCCL and copies from there is timed once as ``e2e.gathered_value``
        e2e_call = lambda: fvdist.simulate_vis_sharded(eng_e, dst=0, shards=shards, beam_list=beam_list,
                                                       host_result="shared", **call)
        e2e_gather_call = lambda: fvdist.simulate_vis_sharded(eng_e, dst=0, shards=shards, beam_list=beam_list,
                                                              host_result="gather", **call)
    else:
        e2e_call = lambda: fftvis_b200.simulate_vis(beam=w["beam"], **call)
    e2e_val, h2d, d2h, first_s = None, 0, 0, None
    if not args.no_e2e:
        # warm-up: two results alive at once, so that torch's caching host allocator owns the two
        # page-locked result blocks a steady stream of calls alternates between (W >= 3 calls in all);
        # the cold first call (page-locking the result block, plan tables) is reported separately
        t0 = time.perf_counter()
        wa = e2e_call()
        first_s = time.perf_counter() - t0
        wb = e2e_call()
        del wa, wb
        res = e2e_call()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            res = e2e_call()
        barrier()
        dt = time.perf_counter() - t0
        tdt = torch.tensor([dt], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tdt, op=dist.ReduceOp.MAX)
        e2e_val = terms(w, nbls) * work_units * e2e_steps / float(tdt.item())
        e2e_gathered = None
        if strong:
            del res
            wg = e2e_gather_call()                             # warm (page-locks rank 0's result block)
            del wg
            barrier()
            t0 = time.perf_counter()
            res = e2e_gather_call()
            barrier()
            tg = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
            dist.all_reduce(tg, op=dist.ReduceOp.MAX)
            e2e_gathered = terms(w, nbls) / float(tg.item())
        csz = 8 * w["precision"]
        h2d = int(np.asarray(w["fluxes"]).size * csz + 3 * 8 * w["nsrc"] + 8 * np.size(w["freqs"]))
        d2h = int(res.size * res.itemsize) if res is not None else 0
        hb = torch.tensor([h2d, d2h], dtype=torch.int64, device="cuda")
        if world > 1:
            dist.all_reduce(hb)                      # bytes over all ranks
        h2d, d2h = int(hb[0].item()), int(hb[1].item())

    if rank == 0:
        peaks = {}
        pk = ROOT / "MEASURED_PEAKS.json"
        if pk.exists():
            peaks = json.loads(pk.read_text())
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
        eng2 = GPUSimulationEngine(freq_batch=args.freq_batch, type1_method=args.type1_method)
        plan2 = eng2.prepare(w["ants"], w["freqs"], w["fluxes"], beam_list, w["ra"], w["dec"], w["times"][:1],
                             w["telescope_loc"], precision=w["precision"], polarized=w["polarized"], **w["kwargs"])
        roof = build_roofline(args, eng, plan2, stages, n_live, hbm_peak, peak_src, w, plan_nf_local, ms / args.steps)
        stage_share = {k: {"ms": v[0], "launches": v[1], "share_of_step": v[0] / extra_ms if extra_ms else None}
                       for k, v in stages_all.items()}
        stage_share["_from"] = {"ms": extra_ms, "launches": 1, "share_of_step": 1.0,
                                "note": "one extra step after the timed region with every stage timed and all kernels "
                                        "on one stream (the timed region overlaps alternate frequency batches on two "
                                        "streams and times only the dominant stage, for the roofline)"}
        if roof is not None and plan2.use_type1 and stages_all.get("spread", (0, 0))[1]:
            sp_ms1, sp_n1 = stages_all["spread"]
            roof["step_level"] = {
                "achieved": roof["alg_bytes_per_launch"] * roof["launches"] / (ms * 1e-3) / 1e9,
                "frac": roof["alg_bytes_per_launch"] * roof["launches"] / (ms * 1e-3) / 1e9 / roof["peak"],
                "note": "all of this kernel's algorithmic bytes over the WHOLE timed region's wall time (every other "
                        "kernel's time included): the sustained figure when launches of several streams overlap"}
            roof["alone"] = {"avg_launch_ms": sp_ms1 / sp_n1,
                             "frac": roof["alg_bytes_per_launch"] / (sp_ms1 / sp_n1 * 1e-3) / 1e9 / roof["peak"],
                             "note": "the same kernel timed in the single-stream extra step (no co-running kernels); "
                                     "`achieved` / `frac` above are from the timed region, where the launches of several "
                                     "streams overlap and each launch therefore lasts longer"}
        cpu = None
        if world == 1 and not args.no_cpu:
            cpu_val, cpu_desc, cores, _ = cpu_sample(w, nbls, budget_s=args.cpu_budget)
            cpu = {"value": cpu_val, "unit": "terms/s", "cores": cores, "kind": "port", "sample": cpu_desc}
        sharding = ("frequency-sharded over the ranks (contiguous blocks, all times each); every finished time slab "
                    "is gathered on rank 0 by NCCL send/recv into its place in the result while later slabs compute"
                    if strong else ("single GPU" if world == 1 else
                                    "each rank simulates its own block of times; no data-path collective"))
        line = {
            "metric": "vis terms/sec (Nsrc*Nbl*Nfreq*Ntime/s)", "value": value, "unit": "terms/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_max / args.steps,
            "higher_is_better": True, "scaling": "strong" if (strong or world == 1 and args.scaling == "strong") else "weak",
            "vs_baseline": None,
            "dtype": "f32" if w["precision"] == 1 else "f64", "data": "synthetic",
            "config": {"workload": f"{w['name']}: {w['desc']}", "nsrc": w["nsrc"], "nbls": nbls, "nfreq": w["nfreq"],
                       "ntimes": w["ntimes"], "n_modes": plan2.n_modes, "freq_batch": fb,
                       "eps": plan2.eps, "type": 1 if plan2.use_type1 else 3, "type1_method": args.type1_method,
                       "n_live_mean": n_live,
                       "l2": "inputs + outputs streamed per step exceed the 126 MB L2 (no explicit flush)",
                       "sharding": sharding, "gathered_bytes_per_step": gathered_bytes},
            "e2e": None if e2e_val is None else
                   {"value": e2e_val, "unit": "terms/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps, "first_call_s": first_s,
                    "api": ("fftvis_b200.gpu.distributed.simulate_vis_sharded(host numpy in, host numpy out on rank 0; "
                            "host_result='shared': every rank copies its frequency block into one shared page-locked array)"
                            if strong else "fftvis_b200.simulate_vis(host numpy in, host numpy out)"),
                    **({"gathered_value": e2e_gathered,
                        "gathered_note": "same call with host_result='gather': slabs collected on rank 0's GPU by NCCL, "
                                         "one D2H link (1 call)"} if strong and e2e_gathered is not None else {})},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roof,
            "stages": stage_share,
            "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def apply_overrides(w, args):
    """Command-line overrides of a workload's simulate() keywords."""
    if getattr(args, "force_type3", False):
        w["kwargs"] = dict(w["kwargs"], force_use_type3=True)
        w["desc"] += " [force_use_type3]"
    return w


def make_workload_nfreq(args) -> int:
    return int(args.nfreq or {"cfg1": 2, "cfg2": 1024, "cfg3": 1024, "cfg4": 512, "cfg5": 1024}[args.workload])


def build_roofline(args, eng, plan2, stages, n_live, hbm_peak, peak_src, w, nf_local, step_ms):
    """``roofline`` of the kernel with the largest share of the step (stage timers: CUDA events around
    every launch, live in the timed region)."""
    P = 4 if plan2.polarized else 1
    r = 4 * plan2.precision
    c = 2 * r
    fb = plan2.freq_batch
    units = nf_local * len(w["times"])                      # (time, frequency) units per step on this rank
    if plan2.use_type1:
        sp_ms, sp_n = stages["spread"]
        if not sp_n:
            return None
        K = plan2.basis["K"] if plan2.basis is not None else 0
        ntrans = K * (K + 1) // 2 if K else max(1, len(plan2.pairs))      # transforms per (time, frequency) unit
        nb_mean = units * args.steps * ntrans / sp_n                       # transforms per launch (ragged last batch)
        alg = spread_alg_bytes(plan2, n_live, nb_mean)
        ach = alg / (sp_ms / sp_n * 1e-3) / 1e9
        wk, nf = kernel_geometry(plan2)
        small = nf * (nf + 1) * c <= 160 * 1024
        xdirect = eng.type1_method == "fused" and plan2.precision == 1 and not small and nf >= 2 * (24 + wk)
        if eng.type1_method != "fused":
            kname = "spread_kernel (type 1, global grid)"
        elif xdirect:
            kname = "t1_xdirect_kernel (pass 1 without an x grid: Fourier sum over the needed columns, gridded in y, accumulators in registers)"
        elif small:
            kname = "t1s_spread_kernel (whole grid per CTA, bin-sorted sources, register windows)"
        else:
            kname = "t1_spread_fftx_kernel (fused spread + FFT-x, strip of the grid in shared memory)"
        real = ("on-chip (FP32 multiply-add issue + sin/cos unit): no fine grid exists, DRAM traffic is the write of T only, "
                "so the HBM fraction is an algorithmic yardstick, not the limiter") if xdirect else \
               ("on-chip (shared-memory pipe / instruction issue): the fine grid never leaves the SM, "
                "so the HBM fraction is an algorithmic yardstick, not the limiter")
        traffic, on_chip = None, None
        tf = ROOT / "profiles" / "r02_traffic.json"
        if tf.exists() and eng.type1_method == "fused":
            prof = json.loads(tf.read_text())
            per_f = prof.get(w["name"], {}).get("pass1_bytes_per_transform")     # ncu dram read + write / transforms
            traffic = None if per_f is None else per_f * nb_mean
            on_chip = prof.get(w["name"] + "_ncu")        # ncu: issue-active and shared-memory pipe utilisation
        return {"kernel": kname, "bound": "hbm", "achieved": ach, "peak": hbm_peak,
                "unit": "GB/s", "frac": ach / hbm_peak, "traffic": traffic, "peak_source": peak_src,
                "alg_bytes_per_launch": alg, "avg_launch_ms": sp_ms / sp_n, "launches": sp_n,
                "real_bound": real, "on_chip_ncu": on_chip}
    # type 3: spreader and the pruned FFT passes; report the larger one
    sp_ms, sp_n = stages["spread"]
    ff_ms, ff_n = stages["fft"]
    dc_ms, dc_n = stages["deconv"]
    geo = eng._nufft.last_type3_geometry() if hasattr(eng._nufft, "last_type3_geometry") else None
    if not geo or not sp_n:
        return None
    G1, G2, dim = geo["G1"], geo["G2"], geo["dim"]
    fft_total = ff_ms + dc_ms
    if fft_total >= sp_ms:
        alg_unit = P * c * (G1 + G2)
        kname = "type-3 inner FFT (pruned shared-memory passes, deconvolution fused)" if geo["own_fft"] else \
                "type-3 deconvolve + pad + cuFFT"
        t = fft_total
    else:
        alg_unit = n_live * (dim * r + P * c) + P * c * G1
        kname = "t3_col_spread_kernel (bin-sorted column tiles)" if geo["tiles"] else "spread_kernel (global atomics)"
        t = sp_ms
    per_step_units = units
    steps = args.steps
    ach = alg_unit * per_step_units * steps / (t * 1e-3) / 1e9
    return {"kernel": kname, "bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s",
            "frac": ach / hbm_peak, "traffic": None, "peak_source": peak_src,
            "alg_bytes_per_unit": alg_unit, "units_per_step": per_step_units, "kernel_ms_per_step": t / steps,
            "share_of_step": t / steps / step_ms if step_ms else None,
            "geometry": geo}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="gpu", choices=["gpu", "reference"])
    ap.add_argument("--workload", default="cfg2")
    ap.add_argument("--nfreq", type=int, default=None)
    ap.add_argument("--ntimes", type=int, default=None)
    ap.add_argument("--nsrc", type=int, default=None)
    ap.add_argument("--freq-batch", type=int, default=None)
    ap.add_argument("--type1-method", default="fused", choices=["fused", "cufft"])
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--cpu-budget", type=float, default=15.0, help="seconds of CPU work per cpu sample")
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"],
                    help="N > 1: strong = one workload frequency-sharded with the NCCL gather (default); "
                         "weak = every rank the whole workload on its own block of times")
    ap.add_argument("--force-type3", action="store_true",
                    help="type-3 transforms even for a gridded array (the reference's force_use_type3)")
    ap.add_argument("--no-beam-tiles", action="store_true",
                    help="table beams gathered from global memory instead of staged in shared memory")
    ap.add_argument("--one-stream", action="store_true", help="no side stream for alternate frequency batches")
    ap.add_argument("--no-e2e", action="store_true", help="skip the end-to-end (host buffers) leg")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
