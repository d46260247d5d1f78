"""Small end-to-end cases for compute-sanitizer (memcheck / racecheck): type 1 fused (segment and
multi-row spreaders), type 1 cuFFT, type 3 2-D and tiled 3-D, polarised weights."""
import sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from fftvis_b200 import AiryBeam, HERA_LOCATION, simulate_vis, synth
from fftvis_b200.gpu import GPUSimulationEngine, gpu_nufft2d_type1

rng = np.random.default_rng(0)
freqs = np.linspace(100e6, 120e6, 3)
times = np.array([2459845.0])
ra, dec, flux = synth.random_sky(600, freqs, seed=1)
# big-grid fused path (segment spreader): n_modes 465 via direct call
x = rng.uniform(-60, 60, 800).astype(np.float32); y = rng.uniform(-60, 60, 800).astype(np.float32)
c = (rng.normal(size=(1, 800)) + 1j * rng.normal(size=(1, 800))).astype(np.complex64)
idx = rng.integers(-232, 233, size=(2, 50))
print("fused big", np.abs(gpu_nufft2d_type1(x, y, c, 465, idx, 6e-8)).sum())
print("cufft", np.abs(gpu_nufft2d_type1(x, y, c, 41, np.clip(idx, -20, 20), 6e-8, method="cufft")).sum())
ants = synth.hex_array(3)
print("t1 small pol", np.abs(simulate_vis(ants, flux, ra, dec, freqs, times, synth.synthetic_uvbeam(freqs, naz=36, nza=19), HERA_LOCATION,
                                         polarized=True, precision=2, eps=1e-10, beam_spline_opts={"order": 3})).sum())
print("t3 2d", np.abs(simulate_vis(ants, flux, ra, dec, freqs, times, AiryBeam(diameter=14.0), HERA_LOCATION, precision=1,
                                  force_use_type3=True)).sum())
ants3 = synth.random_array(8, radius=40.0, zspan=2.0, seed=3)
print("t3 3d", np.abs(simulate_vis(ants3, flux, ra, dec, freqs, times, AiryBeam(diameter=14.0), HERA_LOCATION, precision=2, eps=1e-10,
                                  baselines=synth.all_baselines(ants3))).sum())
