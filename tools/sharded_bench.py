"""torchrun --nproc-per-node N tools/sharded_bench.py [--workload cfg2] : STRONG scaling of one full
workload, frequency-sharded over the ranks with one NCCL gather of the result slabs to rank 0
(fftvis_b200.gpu.distributed.simulate_vis_sharded), checked against the single-GPU result.
Prints one JSON line on rank 0 (development aid; bench.py's N > 1 line is the weak-scaling contract)."""
import argparse, json, os, sys, time
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch, torch.distributed as dist
import bench
from fftvis_b200.gpu import GPUSimulationEngine
from fftvis_b200.gpu.distributed import simulate_vis_sharded

ap = argparse.ArgumentParser(); ap.add_argument("--workload", default="cfg2"); ap.add_argument("--reps", type=int, default=3)
a = ap.parse_args()
local = int(os.environ.get("LOCAL_RANK", 0)); torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
w = bench.make_workload(a.workload)
nbls = bench.n_baselines(w)
beam = w["beam"] if w["polarized"] else w["beam"].to_power()
kw = dict(ants=w["ants"], freqs=w["freqs"], fluxes=w["fluxes"], beam_list=beam if isinstance(beam, list) else [beam],
          ra=w["ra"], dec=w["dec"], times=w["times"], telescope_loc=w["telescope_loc"], precision=w["precision"],
          polarized=w["polarized"], **w["kwargs"])
eng = GPUSimulationEngine()
times = []
for rep in range(a.reps + 2):                       # two warm-up repetitions
    dist.barrier(); torch.cuda.synchronize(); t0 = time.perf_counter()
    got = simulate_vis_sharded(eng, dst=0, **kw)
    torch.cuda.synchronize(); dist.barrier(); dt = time.perf_counter() - t0
    if rep >= 2:
        times.append(dt)
if rank == 0:
    ref = eng.simulate(**kw)
    err = float(np.linalg.norm(got - ref) / np.linalg.norm(ref))
    dt = float(np.median(times))
    print(json.dumps({"workload": w["name"], "n_gpus": world, "scaling": "strong", "seconds_per_call": dt,
                      "terms_per_s": bench.terms(w, nbls) / dt, "rel_err_vs_single_gpu": err,
                      "includes": "host planning + H2D on every rank, frequency-sharded run, NCCL gather to rank 0, D2H"}), flush=True)
dist.barrier(); dist.destroy_process_group()
