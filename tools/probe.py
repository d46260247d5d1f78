"""Development aid: time prepare / run_plan / D2H of a (possibly reduced) BASELINE workload."""
import argparse, json, sys, time
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import bench
from fftvis_b200.gpu import GPUSimulationEngine

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="cfg2"); ap.add_argument("--nfreq", type=int); ap.add_argument("--ntimes", type=int)
ap.add_argument("--nsrc", type=int); ap.add_argument("--reps", type=int, default=2); ap.add_argument("--freq-batch", type=int)
ap.add_argument("--force3", action="store_true"); ap.add_argument("--upsamp", type=float); ap.add_argument("--flo", type=int); ap.add_argument("--fhi", type=int); ap.add_argument("--precision", type=int); ap.add_argument("--eps", type=float)
a = ap.parse_args()
w = bench.make_workload(a.workload, a.nfreq, a.ntimes, a.nsrc)
nbls = bench.n_baselines(w)
beam = w["beam"] if w["polarized"] else w["beam"].to_power()
beam_list = beam if isinstance(beam, list) else [beam]
prec = a.precision or w["precision"]
eng = GPUSimulationEngine(freq_batch=a.freq_batch)
kw = dict(w["kwargs"]); kw["force_use_type3"] = a.force3
if a.eps: kw["eps"] = a.eps
if a.upsamp: kw["upsample_factor"] = a.upsamp
res = {}
for rep in range(a.reps):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    plan = eng.prepare(w["ants"], w["freqs"], w["fluxes"], beam_list, w["ra"], w["dec"], w["times"], w["telescope_loc"],
                       precision=prec, polarized=w["polarized"],
                       freq_range=(a.flo, a.fhi) if a.fhi else None, **kw)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    nufft = eng._nufft_plan(plan.device); nufft.set_timing(True); nufft.reset_timing()
    out = eng.run_plan(plan)
    torch.cuda.synchronize(); t2 = time.perf_counter()
    st = nufft.stage_times(); nufft.set_timing(False)
    host = eng.finish(plan, out)
    t3 = time.perf_counter()
    res = dict(prepare_s=t1 - t0, run_s=t2 - t1, d2h_s=t3 - t2, terms_per_s=bench.terms(w, nbls, nfreq=plan.nf_local) / (t2 - t1), nf_run=plan.nf_local,
               nbls=nbls, type1=plan.use_type1, coplanar=plan.is_coplanar, n_modes=plan.n_modes, freq_batch=plan.freq_batch,
               stages={k: round(v[0], 2) for k, v in st.items() if v[1]}, out_gb=host.nbytes / 1e9,
               plan_bytes_gb=nufft.bytes() / 1e9, finite=bool(np.isfinite(host).all()))
    del plan, out, host
print(json.dumps(res))
