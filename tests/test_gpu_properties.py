"""Size-independent properties at (or near) the BASELINE sizes, where the oracle is too slow to be the
checker: linearity in the sky, conjugate symmetry of flipped baselines, type-1 == type-3 on a gridded
array, independence of the frequency batching, and a sub-sampled direct-sum check at the full cfg2
array size.  (The reference's own tests use the same kind of self-consistency: type 1 == type 3,
tests/test_cpu_simulate.py:199-271; basis == per-antenna, tests/test_beam_basis.py:310-431.)"""
import numpy as np
import pytest

from gpu_helpers import f32_bar, relerr

pytestmark = pytest.mark.gpu

TIMES = 2459845.0 + np.arange(3) * 10.0 / 86400.0


def _cfg2(nsrc=10000, nfreq=48, seed=42):
    from fftvis_b200 import AiryBeam, HERA_LOCATION, synth
    ants = synth.hera350_like()
    freqs = np.linspace(100e6, 200e6, nfreq)
    ra, dec, flux = synth.random_sky(nsrc, freqs, seed=seed, kind="gleam")
    return ants, flux, ra, dec, freqs, AiryBeam(diameter=14.0), HERA_LOCATION


def test_cfg2_array_subsampled_direct_sum_f32_and_f64():
    """Full HERA-350-like array (5861 baselines, n_modes 465, nf 960), 10k sources: a random subset of
    baselines against the term-by-term fp64 measurement equation."""
    from fftvis_b200 import simulate_vis
    from fftvis_b200.core import utils
    from oracle import pipeline
    ants, flux, ra, dec, freqs, beam, loc = _cfg2(nfreq=3)
    reds = [r[0] for r in utils.get_pos_reds(ants, include_autos=True)]
    rng = np.random.default_rng(0)
    sub = [reds[i] for i in sorted(rng.choice(len(reds), 64, replace=False))]
    direct = pipeline.simulate_direct(ants, flux, ra, dec, freqs, TIMES[:1], [beam.to_power()], loc, baselines=sub,
                                      precision=2)
    v64 = simulate_vis(ants, flux, ra, dec, freqs, TIMES[:1], beam, loc, baselines=sub, precision=2, eps=1e-12)
    assert relerr(v64, direct) < 1e-11
    v32 = simulate_vis(ants, flux, ra, dec, freqs, TIMES[:1], beam, loc, baselines=sub, precision=1)
    d32 = pipeline.simulate_direct(ants, flux, ra, dec, freqs, TIMES[:1], [beam.to_power()], loc, baselines=sub,
                                   precision=1)
    # fp32: the bar is the CPU path's own fp32 error on the same inputs (input rounding of phases up to
    # ~60 rad), and the two fp32 results are compared with each other directly
    c32 = pipeline.simulate_cpu(ants, flux, ra, dec, freqs, TIMES[:1], [beam.to_power()], loc, baselines=sub,
                                precision=1)
    assert relerr(v32, d32) < f32_bar(c32, d32)
    assert relerr(v32, c32) < 2 * f32_bar(c32, d32)
    # the same baselines inside the full default baseline set give the same numbers
    full = simulate_vis(ants, flux, ra, dec, freqs, TIMES[:1], beam, loc, precision=2, eps=1e-12)
    idx = [reds.index(b) for b in sub]
    assert relerr(full[..., idx], v64) < 1e-12


def test_cfg2_linearity_and_batch_independence():
    from fftvis_b200.gpu import GPUSimulationEngine
    ants, flux, ra, dec, freqs, beam, loc = _cfg2(nfreq=41)
    rng = np.random.default_rng(1)
    f2 = flux * rng.uniform(0.0, 2.0, size=(flux.shape[0], 1))
    eng = GPUSimulationEngine()
    args = lambda f: (ants, freqs, f, [beam.to_power()], ra, dec, TIMES, loc)
    a = eng.simulate(*args(flux), precision=1)
    b = eng.simulate(*args(f2), precision=1)
    ab = eng.simulate(*args(flux + 2.0 * f2), precision=1)
    assert relerr(ab, a + 2.0 * b) < 5e-6
    other = GPUSimulationEngine(freq_batch=7).simulate(*args(flux), precision=1)
    assert relerr(other, a) < 1e-6           # batching changes nothing but fp32 summation order of nothing
    cufft = GPUSimulationEngine(type1_method="cufft", freq_batch=5).simulate(*args(flux), precision=1)
    assert relerr(cufft, a) < 1e-5           # two fp32 transforms of the same inputs (each ~3e-6 from the truth)


def test_flipped_baselines_are_conjugates_and_type1_equals_type3():
    from fftvis_b200 import AiryBeam, HERA_LOCATION, simulate_vis, synth
    ants = synth.hex_array(6)
    freqs = np.linspace(100e6, 200e6, 5)
    ra, dec, flux = synth.random_sky(20000, freqs, seed=3, kind="gleam")
    beam = AiryBeam(diameter=14.0)
    keys = list(ants.keys())
    rng = np.random.default_rng(4)
    pairs = [(int(keys[i]), int(keys[j])) for i, j in rng.integers(0, len(keys), size=(200, 2))]
    flipped = [(b, a) for a, b in pairs]
    kw = dict(precision=2, eps=1e-12)
    v = simulate_vis(ants, flux, ra, dec, freqs, TIMES, beam, HERA_LOCATION, baselines=pairs, **kw)
    vf = simulate_vis(ants, flux, ra, dec, freqs, TIMES, beam, HERA_LOCATION, baselines=flipped, **kw)
    assert relerr(vf, np.conj(v)) < 1e-11
    v3 = simulate_vis(ants, flux, ra, dec, freqs, TIMES, beam, HERA_LOCATION, baselines=pairs, force_use_type3=True, **kw)
    assert relerr(v3, v) < 1e-10


def test_cfg4_like_3d_type3_subsampled_direct_sum():
    """Random 256-antenna non-flat layout (all 32 640 baselines), 200k diffuse pixels, fp64: the tiled
    3-D spreader at scale against a direct sum over a random subset of baselines."""
    from fftvis_b200 import GaussianBeam, HERA_LOCATION, simulate_vis, synth
    from oracle import pipeline
    ants = synth.random_array(256, radius=150.0, zspan=2.0, seed=42)
    freqs = np.array([150e6, 200e6])
    ra, dec, flux = synth.random_sky(200000, freqs, seed=42, kind="diffuse")
    beam = GaussianBeam(diameter=14.0)
    bls = synth.all_baselines(ants)
    got = simulate_vis(ants, flux, ra, dec, freqs, TIMES[:1], beam, HERA_LOCATION, baselines=bls, precision=2, eps=1e-12)
    assert got.shape == (2, 1, 32640) and np.isfinite(got).all()
    rng = np.random.default_rng(5)
    idx = sorted(rng.choice(len(bls), 48, replace=False))
    direct = pipeline.simulate_direct(ants, flux, ra, dec, freqs, TIMES[:1], [beam.to_power()], HERA_LOCATION,
                                      baselines=[bls[i] for i in idx], precision=2)
    assert relerr(got[..., idx], direct) < 1e-10


def test_fused_type1_is_bitwise_reproducible():
    """The fused type-1 path uses ownership (warps own column segments / row blocks) instead of
    atomics, with rank-ordered hit lists: repeated runs must be bit-identical.  (compute-sanitizer is
    closed on the GPU pool, so this doubles as the race check of the shared-memory spreader.)"""
    from fftvis_b200.gpu import GPUSimulationEngine
    ants, flux, ra, dec, freqs, beam, loc = _cfg2(nfreq=37)
    eng = GPUSimulationEngine()
    args = (ants, freqs, flux, [beam.to_power()], ra, dec, TIMES[:2], loc)
    runs = [eng.simulate(*args, precision=1) for _ in range(3)]
    assert np.array_equal(runs[0], runs[1]) and np.array_equal(runs[0], runs[2])
    # small-grid (row-block ownership) spreader, fp64, four products
    from fftvis_b200 import HERA_LOCATION, synth
    f2 = np.linspace(100e6, 200e6, 5)
    ra2, dec2, fl2 = synth.random_sky(20000, f2, seed=9)
    tb = synth.synthetic_uvbeam(f2, naz=72, nza=37)
    a2 = (synth.hex_array(11), f2, fl2, [tb], ra2, dec2, TIMES[:1], HERA_LOCATION)
    r = [eng.simulate(*a2, precision=2, polarized=True, beam_spline_opts={"order": 1}) for _ in range(3)]
    assert np.array_equal(r[0], r[1]) and np.array_equal(r[0], r[2])
