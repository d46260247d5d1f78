"""Development aid: cProfile of the second simulate_vis call on the cfg2 workload."""
import cProfile, pstats, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import bench, fftvis_b200
w = bench.make_workload("cfg2")
call = dict(ants=w["ants"], fluxes=w["fluxes"], ra=w["ra"], dec=w["dec"], freqs=w["freqs"], times=w["times"],
            beam=w["beam"], telescope_loc=w["telescope_loc"], precision=w["precision"], polarized=w["polarized"])
r1 = fftvis_b200.simulate_vis(**call)
r2 = fftvis_b200.simulate_vis(**call)
del r1, r2
for i in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    r = fftvis_b200.simulate_vis(**call)
    torch.cuda.synchronize(); print("call", i, time.perf_counter() - t0, flush=True)
pr = cProfile.Profile(); pr.enable()
r = fftvis_b200.simulate_vis(**call)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
