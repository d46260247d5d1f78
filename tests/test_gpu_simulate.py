"""End-to-end parity of ``simulate_vis(backend="gpu")`` against the oracle: the reference-structured
CPU pipeline (oracle.pipeline.simulate_cpu) and the term-by-term fp64 measurement equation
(oracle.pipeline.simulate_direct).  Mirrors the reference's own integration tests
(tests/test_cpu_simulate.py:75-271 type-1 == type-3 and vs a direct simulator; tests/test_beam_basis.py
:310-431 basis == per-antenna; tests/test_wrapper.py:318 f32 ~ f64).
Bar: relative L2 error <= 10 x eps (fp64); fp32 is limited by input rounding (see test_gpu_nufft)."""
import numpy as np
import pytest

from gpu_helpers import TIMES, f32_bar, hex_ants, relerr, small_sky

pytestmark = pytest.mark.gpu

FREQS = np.array([100e6, 110e6])


def _cfg1(nsrc=100, polarized_sky=False):
    from fftvis_b200 import HERA_LOCATION
    from fftvis_b200 import synth
    ants = synth.hex_rows((3, 4, 3))        # the 10-antenna hex of BASELINE configs[0]
    ra, dec, flux = small_sky(nsrc, FREQS, polarized=polarized_sky)
    return ants, flux, ra, dec, HERA_LOCATION


@pytest.mark.parametrize("force3", [False, True])
@pytest.mark.parametrize("precision,eps", [(2, 1e-13), (2, 1e-10), (1, 6e-8)])
def test_cfg1_unpolarized_gaussian(precision, eps, force3):
    """BASELINE configs[0]: 10-antenna hex, 100 sources, 2 freqs, unpolarized Gaussian beam."""
    from fftvis_b200 import GaussianBeam, simulate_vis
    from oracle import pipeline
    ants, flux, ra, dec, loc = _cfg1()
    beam = GaussianBeam(diameter=14.0)
    got = simulate_vis(ants, flux, ra, dec, FREQS, TIMES[:1], beam, loc, precision=precision, eps=eps,
                       force_use_type3=force3)
    cpu = pipeline.simulate_cpu(ants, flux, ra, dec, FREQS, TIMES[:1], [beam.to_power()], loc,
                                precision=precision, eps=eps, force_use_type3=force3)
    direct = pipeline.simulate_direct(ants, flux, ra, dec, FREQS, TIMES[:1], [beam.to_power()], loc,
                                      precision=precision)
    assert got.shape == cpu.shape == (2, 1, got.shape[-1])
    assert got.dtype == (np.complex64 if precision == 1 else np.complex128)
    if precision == 2:
        assert relerr(got, direct) < 10 * eps
        assert relerr(got, cpu) < 10 * eps
    else:
        assert relerr(got, direct) < f32_bar(cpu, direct, eps)
        assert relerr(got, cpu) < 2 * f32_bar(cpu, direct, eps)


@pytest.mark.parametrize("force3", [False, True])
@pytest.mark.parametrize("pol_sky", [False, True])
def test_polarized_table_beam(pol_sky, force3):
    from fftvis_b200 import simulate_vis, synth
    from oracle import pipeline
    ants, flux, ra, dec, loc = _cfg1(300, polarized_sky=pol_sky)
    beam = synth.synthetic_uvbeam(FREQS, naz=72, nza=37)
    kw = dict(precision=2, eps=1e-12, polarized=True, beam_spline_opts={"order": 1}, force_use_type3=force3)
    got = simulate_vis(ants, flux, ra, dec, FREQS, TIMES, beam, loc, **kw)
    cpu = pipeline.simulate_cpu(ants, flux, ra, dec, FREQS, TIMES, [beam], loc, **kw)
    assert got.shape == cpu.shape and got.shape[:4] == (2, 2, 2, 2)
    assert relerr(got, cpu) < 1e-11
    direct = pipeline.simulate_direct(ants, flux, ra, dec, FREQS, TIMES, [beam], loc, precision=2,
                                      polarized=True, beam_spline_opts={"order": 1})
    assert relerr(got, direct) < 1e-11


@pytest.mark.parametrize("polarized", [False, True])
@pytest.mark.parametrize("force3", [False, True])
def test_per_antenna_beams_with_flipped_pairs(polarized, force3):
    """Two beam types alternating over antennas, explicit baselines in both orders
    (reference tests/test_cpu_simulate.py:273-382 uses one un-flipped baseline; cpu/beams.py:115-125)."""
    from fftvis_b200 import AiryBeam, GaussianBeam, simulate_vis, synth
    from oracle import pipeline
    ants, flux, ra, dec, loc = _cfg1(200)
    nant = len(ants)
    beam_idx = np.arange(nant) % 2
    if polarized:
        beams = [synth.synthetic_uvbeam(FREQS, naz=72, nza=37, seed=s, perturb=0.3) for s in (1, 2)]
    else:
        beams = [GaussianBeam(diameter=14.0), AiryBeam(diameter=12.0)]
    baselines = [(0, 1), (1, 0), (1, 2), (2, 4), (3, 3), (5, 2), (4, 4), (6, 9)]
    kw = dict(precision=2, eps=1e-12, polarized=polarized, beam_spline_opts={"order": 1},
              force_use_type3=force3, beam_idx=beam_idx, baselines=baselines)
    got = simulate_vis(ants, flux, ra, dec, FREQS, TIMES, beams, loc, **kw)
    cpu_beams = beams if polarized else [b.to_power() for b in beams]
    cpu = pipeline.simulate_cpu(ants, flux, ra, dec, FREQS, TIMES, cpu_beams, loc, **kw)
    assert relerr(got, cpu) < 1e-11


def test_basis_path_matches_per_antenna_path():
    """reference tests/test_beam_basis.py:310-431 (atol 1e-5 there)."""
    from fftvis_b200 import simulate_vis, synth
    from oracle import pipeline
    ants, flux, ra, dec, loc = _cfg1(150)
    nant, K = len(ants), 3
    rng = np.random.default_rng(42)
    basis = [synth.synthetic_uvbeam(FREQS, naz=72, nza=37, seed=s, perturb=0.3) for s in range(K)]
    for b in basis:                      # real-valued basis beams: the regime where the reference's
        b.data_array = b.data_array.real.astype(complex)   # upper-triangle trick is exact
    coefs = rng.normal(size=(nant, K, FREQS.size)) + 1j * rng.normal(size=(nant, K, FREQS.size))
    baselines = [(0, 1), (2, 5), (3, 3), (9, 4), (1, 0)]
    kw = dict(precision=2, eps=1e-12, polarized=True, beam_spline_opts={"order": 1}, baselines=baselines)
    got = simulate_vis(ants, flux, ra, dec, FREQS, TIMES, basis, loc, beam_coefs=coefs, **kw)
    cpu = pipeline.simulate_cpu(ants, flux, ra, dec, FREQS, TIMES, basis, loc, beam_coefs=coefs, **kw)
    assert relerr(got, cpu) < 1e-11
    direct = pipeline.simulate_direct(ants, flux, ra, dec, FREQS, TIMES, basis, loc, precision=2, polarized=True,
                                      beam_spline_opts={"order": 1}, beam_coefs=coefs, baselines=baselines)
    assert relerr(got, direct) < 1e-10
    got3 = simulate_vis(ants, flux, ra, dec, FREQS, TIMES, basis, loc, beam_coefs=coefs, force_use_type3=True, **kw)
    assert relerr(got3, got) < 1e-11


def test_tilted_nonflat_array_uses_3d_and_matches_direct():
    from fftvis_b200 import AiryBeam, HERA_LOCATION, simulate_vis, synth
    from oracle import pipeline
    ants = synth.random_array(12, radius=60.0, zspan=2.0, seed=42)
    ra, dec, flux = small_sky(400, FREQS)
    beam = AiryBeam(diameter=14.0)
    bls = synth.all_baselines(ants, autos=True)
    got = simulate_vis(ants, flux, ra, dec, FREQS, TIMES, beam, HERA_LOCATION, precision=2, eps=1e-12, baselines=bls)
    direct = pipeline.simulate_direct(ants, flux, ra, dec, FREQS, TIMES, [beam.to_power()], HERA_LOCATION,
                                      precision=2, baselines=bls)
    assert got.shape == (2, 2, len(bls))
    assert relerr(got, direct) < 1e-11
    # tilted but flat plane: rotation to the xy plane then the 2-D transform
    tilt = {k: np.array([v[0], v[1], 0.02 * v[0] - 0.01 * v[1]]) for k, v in ants.items()}
    got = simulate_vis(tilt, flux, ra, dec, FREQS, TIMES, beam, HERA_LOCATION, precision=2, eps=1e-12, baselines=bls)
    direct = pipeline.simulate_direct(tilt, flux, ra, dec, FREQS, TIMES, [beam.to_power()], HERA_LOCATION,
                                      precision=2, baselines=bls)
    assert relerr(got, direct) < 1e-11


def test_source_chunks_and_frequency_batches_do_not_change_the_answer():
    from fftvis_b200 import GaussianBeam, HERA_LOCATION
    from fftvis_b200.gpu import GPUSimulationEngine
    ants = hex_ants(3)
    freqs = np.linspace(100e6, 120e6, 7)
    ra, dec, flux = small_sky(1000, freqs)
    beam = GaussianBeam(diameter=14.0).to_power()
    args = (ants, freqs, flux, [beam], ra, dec, TIMES, HERA_LOCATION)
    base = GPUSimulationEngine().simulate(*args, precision=2, eps=1e-12)
    chunked = GPUSimulationEngine(freq_batch=3).simulate(*args, precision=2, eps=1e-12, nchunks=3, source_buffer=1.0)
    assert relerr(chunked, base) < 1e-12
    with pytest.raises(ValueError, match="source_buffer"):
        GPUSimulationEngine().simulate(*args, precision=2, eps=1e-12, nchunks=1, source_buffer=0.3)


def test_evaluate_vis_chunk_slices_match_simulate():
    from fftvis_b200 import GaussianBeam, HERA_LOCATION
    from fftvis_b200.gpu import GPUSimulationEngine
    ants = hex_ants(2)
    freqs = np.linspace(100e6, 120e6, 5)
    ra, dec, flux = small_sky(200, freqs)
    eng = GPUSimulationEngine()
    beam = GaussianBeam(diameter=14.0).to_power()
    full = eng.simulate(ants, freqs, flux, [beam], ra, dec, TIMES, HERA_LOCATION, precision=2, eps=1e-12)
    plan = eng.prepare(ants, freqs, flux, [beam], ra, dec, TIMES, HERA_LOCATION, precision=2, eps=1e-12)
    blk = eng._evaluate_vis_chunk(slice(1, 2), slice(2, 5), plan=plan)
    assert blk.shape == (1, full.shape[-1], 1, 1, 3)
    np.testing.assert_allclose(blk[0, :, 0, 0, :].T, full[2:5, 1], rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize("gridded", [False, True])
@pytest.mark.parametrize("polarized", [False, True])
def test_evaluate_vis_chunk_reference_call_form(gridded, polarized):
    """The chunk evaluator called the way the reference's own test calls the CPU engine
    (/root/reference/tests/test_cpu_simulate.py:1036-1087): coordinate manager + plane-rotated baselines in
    seconds (type 3) or integer lattice offsets with the basis matrix (type 1), against ``simulate``."""
    from fftvis_b200 import GaussianBeam, HERA_LOCATION
    from fftvis_b200.core import antenna_gridding, utils
    from fftvis_b200.core.coords import CoordinateRotation
    from fftvis_b200.gpu import GPUSimulationEngine
    ants = hex_ants(2)
    if not gridded:                       # a perturbed hex is not a lattice: type 3
        rng = np.random.default_rng(5)
        ants = {k: v + np.r_[rng.uniform(-0.4, 0.4, 2), 0.0] for k, v in ants.items()}
    freqs = np.linspace(100e6, 120e6, 4)
    ra, dec, flux = small_sky(150, freqs)
    beam = GaussianBeam(diameter=14.0)
    beam = beam if polarized else beam.to_power()
    eng = GPUSimulationEngine()
    full = eng.simulate(ants, freqs, flux, [beam], ra, dec, TIMES, HERA_LOCATION, precision=2, eps=1e-12,
                        polarized=polarized)
    # ---- the front half of simulate, restated as the reference's test does
    baselines = [r[0] for r in utils.get_pos_reds(ants, include_autos=True)]
    antnums = list(ants.keys())
    antvecs = np.array([ants[a] for a in antnums], dtype=np.float64)
    coord_mgr = CoordinateRotation._methods["CoordinateRotationERFA"]
    coord_mgr = CoordinateRotation(flux=0.5 * flux, times=TIMES, telescope_loc=HERA_LOCATION, skycoords=(ra, dec),
                                   precision=2, source_buffer=1.0, chunk_size=ra.size, method=coord_mgr)
    if gridded:
        ok, gpos, basis = antenna_gridding.check_antpos_griddability(ants)
        assert ok
        bls = np.round(np.array([gpos[b[1]] - gpos[b[0]] for b in baselines]).T).astype(int)
        kw = dict(rotation_matrix=np.eye(3), bls=bls, use_type1=True, basis_matrix=basis / utils.speed_of_light,
                  type1_n_modes=2 * int(np.abs(bls).max()) + 1, is_coplanar=True)
    else:
        rot = np.ascontiguousarray(utils.get_plane_to_xy_rotation_matrix(antvecs).T)
        rants = rot @ antvecs.T
        idx = {a: i for i, a in enumerate(antnums)}
        bls = np.array([rants[:, idx[b[1]]] - rants[:, idx[b[0]]] for b in baselines]).T / utils.speed_of_light
        kw = dict(rotation_matrix=rot, bls=bls, is_coplanar=True)
    nf = 2 if polarized else 1
    blk = eng._evaluate_vis_chunk(
        time_idx=slice(None), freq_idx=slice(1, 4), beam_list=[beam], coord_mgr=coord_mgr, antnums=antnums,
        baselines=baselines, freqs=freqs, complex_dtype=np.complex128, nfeeds=nf, polarized=polarized, eps=1e-12,
        beam_spline_opts=None, interpolation_function="az_za_map_coordinates", n_threads=1, trace_mem=False, **kw)
    assert blk.shape == (len(TIMES), len(baselines), nf, nf, 3)
    want = full[1:4].reshape(3, len(TIMES), nf, nf, len(baselines))
    np.testing.assert_allclose(np.transpose(blk, (4, 0, 2, 3, 1)), want, rtol=1e-10, atol=1e-10 * np.abs(want).max())
    with pytest.raises(TypeError, match="plan="):
        eng._evaluate_vis_chunk(slice(None), slice(None))


def test_cubic_spline_beam_matches_cpu_pipeline():
    """beam_spline_opts={"order": 3}: cubic B-spline interpolation of the UVBeam table (host prefilter,
    device 4 x 4 taps) against scipy's map_coordinates inside the CPU pipeline."""
    from fftvis_b200 import simulate_vis, synth
    from oracle import pipeline
    ants, flux, ra, dec, loc = _cfg1(300)
    beam = synth.synthetic_uvbeam(FREQS, naz=72, nza=37)
    kw = dict(precision=2, eps=1e-12, polarized=True, beam_spline_opts={"order": 3})
    got = simulate_vis(ants, flux, ra, dec, FREQS, TIMES, beam, loc, **kw)
    cpu = pipeline.simulate_cpu(ants, flux, ra, dec, FREQS, TIMES, [beam], loc, **kw)
    assert relerr(got, cpu) < 1e-11
    lin = simulate_vis(ants, flux, ra, dec, FREQS, TIMES, beam, loc, precision=2, eps=1e-12, polarized=True,
                       beam_spline_opts={"order": 1})
    assert 1e-6 < relerr(lin, got) < 1e-1          # the two interpolation orders genuinely differ
    with pytest.raises(NotImplementedError):
        simulate_vis(ants, flux, ra, dec, FREQS, TIMES, beam, loc, polarized=True, beam_spline_opts={"order": 2})


def test_type1_fused_and_cufft_paths_agree():
    from fftvis_b200 import AiryBeam, HERA_LOCATION, synth
    from fftvis_b200.gpu import GPUSimulationEngine
    ants = hex_ants(4)
    freqs = np.linspace(100e6, 200e6, 9)
    ra, dec, flux = small_sky(2000, freqs)
    beam = synth.synthetic_uvbeam(freqs, naz=72, nza=37)
    args = (ants, freqs, flux, [beam], ra, dec, TIMES, HERA_LOCATION)
    kw = dict(precision=2, eps=1e-12, polarized=True, beam_spline_opts={"order": 1})
    a = GPUSimulationEngine(type1_method="fused", freq_batch=4).simulate(*args, **kw)
    b = GPUSimulationEngine(type1_method="cufft", freq_batch=4).simulate(*args, **kw)
    assert relerr(a, b) < 1e-11


def test_f32_close_to_f64_and_no_sources_above_horizon():
    """reference tests/test_wrapper.py:318 (rtol = atol = 1e-5 there, on O(1) visibilities)."""
    from fftvis_b200 import AiryBeam, HERA_LOCATION, simulate_vis
    ants = hex_ants(3)
    ra, dec, flux = small_sky(500, FREQS)
    beam = AiryBeam(diameter=14.0)
    v64 = simulate_vis(ants, flux, ra, dec, FREQS, TIMES, beam, HERA_LOCATION, precision=2)
    v32 = simulate_vis(ants, flux, ra, dec, FREQS, TIMES, beam, HERA_LOCATION, precision=1)
    from oracle import pipeline
    c32 = pipeline.simulate_cpu(ants, flux, ra, dec, FREQS, TIMES, [beam.to_power()], HERA_LOCATION, precision=1)
    assert v32.dtype == np.complex64 and relerr(v32, v64) < f32_bar(c32, v64)
    assert relerr(v32, c32) < 2 * f32_bar(c32, v64)
    # every source below the horizon -> exact zeros
    below_dec = np.full(5, np.deg2rad(80.0))
    out = simulate_vis(ants, flux[:5], ra[:5], below_dec, FREQS, TIMES, beam, HERA_LOCATION, precision=2)
    assert np.all(out == 0)
    out = simulate_vis(ants, flux[:5], ra[:5], below_dec, FREQS, TIMES, beam, HERA_LOCATION, precision=2,
                       force_use_type3=True)
    assert np.all(out == 0)


@pytest.mark.parametrize("force3", [False, True])
def test_polarized_analytic_and_uniform_beams(force3):
    from fftvis_b200 import AiryBeam, UniformBeam, simulate_vis
    from oracle import pipeline
    ants, flux, ra, dec, loc = _cfg1(150)
    for beam in (AiryBeam(diameter=14.0), UniformBeam()):
        kw = dict(precision=2, eps=1e-12, polarized=True, force_use_type3=force3)
        got = simulate_vis(ants, flux, ra, dec, FREQS, TIMES, beam, loc, **kw)
        cpu = pipeline.simulate_cpu(ants, flux, ra, dec, FREQS, TIMES, [beam], loc, **kw)
        assert relerr(got, cpu) < 1e-11
        # an unpolarised analytic beam gives identical xx / yy and identical cross products
        np.testing.assert_allclose(got[:, :, 0, 0], got[:, :, 1, 1], rtol=1e-12, atol=1e-12)


def test_degenerate_inputs():
    """Autos only (zero-length baselines), a single source, one antenna, and an empty sky."""
    from fftvis_b200 import GaussianBeam, HERA_LOCATION, simulate_vis
    from oracle import pipeline
    beam = GaussianBeam(diameter=14.0)
    ants = {0: np.array([0.0, 0.0, 0.0]), 5: np.array([14.6, 0.0, 0.0])}
    ra, dec, flux = small_sky(50, FREQS)
    autos = [(0, 0), (5, 5)]
    got = simulate_vis(ants, flux, ra, dec, FREQS, TIMES, beam, HERA_LOCATION, baselines=autos, precision=2)
    ref = pipeline.simulate_direct(ants, flux, ra, dec, FREQS, TIMES, [beam.to_power()], HERA_LOCATION,
                                   baselines=autos, precision=2)
    assert relerr(got, ref) < 1e-11 and np.abs(got.imag).max() < 1e-9 * np.abs(got.real).max()
    one = {7: np.array([1.0, 2.0, 0.0])}
    got = simulate_vis(one, flux, ra, dec, FREQS, TIMES, beam, HERA_LOCATION, precision=2)
    ref = pipeline.simulate_direct(one, flux, ra, dec, FREQS, TIMES, [beam.to_power()], HERA_LOCATION, precision=2)
    assert got.shape == ref.shape == (2, 2, 1) and relerr(got, ref) < 1e-11
    got = simulate_vis(ants, flux[:1], ra[:1], dec[:1], FREQS, TIMES, beam, HERA_LOCATION, precision=2)
    ref = pipeline.simulate_direct(ants, flux[:1], ra[:1], dec[:1], FREQS, TIMES, [beam.to_power()], HERA_LOCATION,
                                   precision=2)
    assert relerr(got, ref) < 1e-11 or np.all(ref == 0)
    empty = simulate_vis(ants, flux[:0], ra[:0], dec[:0], FREQS, TIMES, beam, HERA_LOCATION, precision=2)
    assert empty.shape[:2] == (2, 2) and np.all(empty == 0)


def test_wrapper_errors_match_reference_strings():
    """reference tests/test_wrapper.py:123-141, tests/test_beam_basis.py:459,476."""
    from fftvis_b200 import GaussianBeam, HERA_LOCATION, simulate_vis
    ants = hex_ants(2)
    ra, dec, flux = small_sky(10, FREQS)
    b = GaussianBeam(diameter=14.0)
    with pytest.raises(ValueError, match="beam_idx must be provided"):
        simulate_vis(ants, flux, ra, dec, FREQS, TIMES, [b, b], HERA_LOCATION)
    with pytest.raises(ValueError, match="beam_idx must be length nant"):
        simulate_vis(ants, flux, ra, dec, FREQS, TIMES, [b, b], HERA_LOCATION, beam_idx=np.zeros(3, int))
    with pytest.raises(ValueError, match="not compatible with unpolarized"):
        simulate_vis(ants, flux, ra, dec, FREQS, TIMES, [b], HERA_LOCATION, beam_coefs=np.ones((7, 1, 2)))
    with pytest.raises(ValueError, match="Unsupported backend"):
        simulate_vis(ants, flux, ra, dec, FREQS, TIMES, b, HERA_LOCATION, backend="tpu")


@pytest.mark.parametrize("nchunks", [1, 3])
def test_streamed_result_equals_plain_copy(nchunks):
    """``simulate`` streams every finished time slab to the pinned result on a copy stream
    (fv_memcpy2d_async) while later time steps compute; the array it returns must be bit-identical to
    one plain device-to-host copy of ``run_plan``'s output, for one and for several source chunks."""
    import torch
    from fftvis_b200 import AiryBeam, HERA_LOCATION
    from fftvis_b200.gpu import GPUSimulationEngine
    ants = hex_ants(19)
    freqs = np.linspace(100e6, 120e6, 5)
    times = 2459845.0 + np.arange(4) * 600 / 86400.0
    ra, dec, flux = small_sky(700, freqs)
    kw = dict(precision=1, polarized=False, nchunks=nchunks, source_buffer=1.0)
    eng = GPUSimulationEngine()
    beam = AiryBeam(diameter=14.0).to_power()
    streamed = eng.simulate(ants, freqs, flux, [beam], ra, dec, times, HERA_LOCATION, **kw)
    plan = eng.prepare(ants, freqs, flux, [beam], ra, dec, times, HERA_LOCATION, **kw)
    out = eng.run_plan(plan)
    torch.cuda.synchronize()
    plain = out.cpu().numpy().reshape(streamed.shape)
    assert streamed.shape == (freqs.size, times.size, plain.shape[-1])
    assert np.array_equal(streamed, plain)
    assert np.abs(streamed).max() > 0


def test_time_major_output_and_sharded_driver_world1():
    """``run_plan`` through a permuted (time-major) view, and the frequency-sharded driver
    (gpu/distributed.py ``simulate_vis_sharded``: per-slab gather + host streaming) on a one-rank NCCL
    group, must both reproduce ``simulate`` bit for bit.  (world_size 2 runs on CPU over gloo in
    tests/test_distributed_cpu.py and on 2..8 GPUs in tools/sharded_check.py.)"""
    import os
    import socket
    import torch
    import torch.distributed as dist
    from fftvis_b200 import HERA_LOCATION, synth
    from fftvis_b200.gpu import GPUSimulationEngine
    from fftvis_b200.gpu.distributed import simulate_vis_sharded
    ants = hex_ants(4)
    freqs = np.linspace(100e6, 200e6, 7)
    times = 2459845.0 + np.arange(3) * 600 / 86400.0
    ra, dec, flux = small_sky(900, freqs)
    beam = synth.synthetic_uvbeam(freqs, naz=72, nza=37)
    kw = dict(ants=ants, freqs=freqs, fluxes=flux, beam_list=[beam], ra=ra, dec=dec, times=times,
              telescope_loc=HERA_LOCATION, precision=2, eps=1e-12, polarized=True, beam_spline_opts={"order": 1})
    eng = GPUSimulationEngine(freq_batch=3)
    ref = eng.simulate(**kw)
    plan = eng.prepare(**kw)
    buf = torch.empty((times.size, freqs.size, 4, plan.nbls), dtype=torch.complex128, device="cuda")
    eng.run_plan(plan, out=buf.permute(1, 0, 2, 3))
    tm = buf.cpu().numpy().transpose(1, 0, 2, 3).reshape(ref.shape)
    assert np.array_equal(tm, ref)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", torch.cuda.current_device()))
    try:
        got = simulate_vis_sharded(eng, dst=0, nprocesses=4, **kw)      # CPU-only knobs are accepted
        assert got.shape == ref.shape and np.array_equal(got, ref)
        got_all = simulate_vis_sharded(eng, dst=None, **kw)
        assert np.array_equal(got_all, ref)
    finally:
        dist.destroy_process_group()
