"""torchrun --nproc-per-node N tools/sharded_check.py : frequency-sharded simulate + NCCL gather must
equal the single-GPU result (development aid / multi-GPU smoke)."""
import os, sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch, torch.distributed as dist
from fftvis_b200 import AiryBeam, HERA_LOCATION, synth
from fftvis_b200.gpu import GPUSimulationEngine
from fftvis_b200.gpu.distributed import simulate_vis_sharded

local = int(os.environ.get("LOCAL_RANK", 0)); torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank = dist.get_rank()
ants = synth.hex_array(4); freqs = np.linspace(100e6, 200e6, 11)        # ragged shards at N = 2, 4, 8
ra, dec, flux = synth.random_sky(3000, freqs, seed=1)
times = 2459845.0 + np.arange(3) * 10 / 86400
kw = dict(ants=ants, freqs=freqs, fluxes=flux, beam_list=[AiryBeam(diameter=14.0).to_power()], ra=ra, dec=dec, times=times,
          telescope_loc=HERA_LOCATION, precision=2, eps=1e-12)
eng = GPUSimulationEngine()
got = simulate_vis_sharded(eng, dst=0, **kw)
shared = simulate_vis_sharded(eng, dst=0, host_result="shared", **kw)
shared2 = simulate_vis_sharded(eng, dst=0, host_result="shared", **kw)       # the other segment
if rank == 0:
    print("shared host block equals the gathered result:", bool(np.array_equal(shared, got)),
          bool(np.array_equal(shared2, got)), flush=True)
    assert np.array_equal(shared, got) and np.array_equal(shared2, got)
    ref = eng.simulate(**kw)
    err = np.linalg.norm(got - ref) / np.linalg.norm(ref)
    print(f"sharded over {dist.get_world_size()} ranks vs single GPU: rel err {err:.2e}", flush=True)
    print("bit-identical:", bool(np.array_equal(got, ref)), flush=True)
    assert err < 1e-12
dist.barrier(); dist.destroy_process_group()
