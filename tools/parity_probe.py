"""Development aid: where does the fp64 parity error at full size come from? (coordinates vs NUFFT)"""
import sys, time
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from fftvis_b200 import GaussianBeam, HERA_LOCATION, simulate_vis, synth
from oracle import pipeline
def relerr(a, b): return float(np.linalg.norm(np.ravel(a - b)) / np.linalg.norm(np.ravel(b)))
T0 = np.array([2459845.0])
# type 3 fp32 sigma 1.25
from fftvis_b200.gpu import gpu_nufft3d, gpu_nufft2d
from oracle import nufft_cpu as nc
from fftvis_b200.gpu.nufft import default_plan
for seed in range(3):
    rng = np.random.default_rng(seed)
    n, nk = 900, 130
    lm = rng.uniform(-0.7, 0.7, (2, n)); nn = np.sqrt(1 - (lm**2).sum(0))
    xs = [(2 * np.pi * a).astype(np.float32) for a in (lm[0], lm[1], nn)]
    ss = [rng.uniform(-12, 12, nk).astype(np.float32), rng.uniform(-12, 12, nk).astype(np.float32), rng.uniform(-0.5, 0.5, nk).astype(np.float32)]
    c = (rng.normal(size=(2, n)) + 1j * rng.normal(size=(2, n))).astype(np.complex64)
    for up in (2.0, 1.25):
        for dim in (2, 3):
          for opt in ((1, 1), (0, 1), (1, 0), (0, 0)):
            default_plan().set_option("t3_fft", 2 * opt[0]); default_plan().set_option("t3_tiles", opt[1])
            got = gpu_nufft3d(*xs, c, *ss, 6e-8, upsample_factor=up) if dim == 3 else gpu_nufft2d(xs[0], xs[1], c, ss[0], ss[1], 6e-8, upsample_factor=up)
            want = nc.direct_sum(xs[0], xs[1], xs[2] if dim == 3 else None, c, ss[0], ss[1], ss[2] if dim == 3 else None)
            cpu = nc.nufft_type3(xs[:dim], c, ss[:dim], 6e-8, upsampfac=up)
            print(f"t3 f32 seed={seed} up={up} dim={dim} ownfft,tiles={opt}: gpu {relerr(got, want):.2e} cpu {relerr(cpu, want):.2e}", flush=True)
