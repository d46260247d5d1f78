"""Beam tables staged in shared memory (csrc/weights_tiled.cuh: live set sorted by beam-grid tile, patches copied
by cp.async.bulk / plain loads, sources interpolated from shared memory) against the gather-from-global kernels
(``engine.beam_tiles = False``) and, through them, the oracle: the two evaluate the same taps with the same
weights, only the order of the sources -- hence of the NUFFT's sums -- differs."""
import numpy as np
import pytest

from gpu_helpers import TIMES, hex_ants, relerr, small_sky

pytestmark = pytest.mark.gpu

FREQS = np.linspace(100e6, 130e6, 7)


def _run(tiles, beam, polarized, precision, order, nsrc=6000, **kw):
    from fftvis_b200 import HERA_LOCATION
    from fftvis_b200.gpu import GPUSimulationEngine
    ants = hex_ants(3)
    ra, dec, flux = small_sky(nsrc, FREQS, seed=3)
    eng = GPUSimulationEngine(freq_batch=3)
    eng.beam_tiles = tiles
    beam_list = beam if isinstance(beam, list) else [beam]
    plan = eng.prepare(ants, FREQS, flux, beam_list, ra, dec, TIMES, HERA_LOCATION, precision=precision,
                       polarized=polarized, eps=1e-12 if precision == 2 else 6e-8,
                       beam_spline_opts={"order": order}, **kw)
    out = eng.run_plan(plan)
    eng.check_source_buffer(plan)
    assert bool(plan.work["tiles_ok"]) and (("tiles" in plan.work) is True)
    return eng.finish(plan, out)


@pytest.mark.parametrize("precision", [2, 1])
@pytest.mark.parametrize("order", [1, 0])
@pytest.mark.parametrize("polarized", [True, False])
def test_tiled_tables_match_gathered_tables(precision, order, polarized):
    from fftvis_b200 import synth
    beam = synth.synthetic_uvbeam(FREQS, naz=90, nza=46)        # 4-degree grid: 3 x 6 tiles, partial edge tiles
    if not polarized:
        beam = beam.to_power()
    a = _run(True, beam, polarized, precision, order)
    b = _run(False, beam, polarized, precision, order)
    assert np.isfinite(a).all()
    assert relerr(a, b) < (1e-12 if precision == 2 else 3e-6)
    a2 = _run(True, beam, polarized, precision, order)
    assert np.array_equal(a, a2)                                 # stable sort: bitwise reproducible


def test_tiled_tables_with_source_chunks_and_two_beams():
    """Two different table beams on one grid (pair form with K = 2) and the catalogue in three chunks."""
    from fftvis_b200 import synth
    b0 = synth.synthetic_uvbeam(FREQS, naz=72, nza=37, seed=0)
    b1 = synth.synthetic_uvbeam(FREQS, naz=72, nza=37, seed=1, perturb=0.05)
    idx = np.arange(len(hex_ants(3))) % 2
    kw = dict(beam_idx=idx, nchunks=3, source_buffer=1.0, nsrc=9000)
    a = _run(True, [b0, b1], True, 2, 1, **kw)
    b = _run(False, [b0, b1], True, 2, 1, **kw)
    assert relerr(a, b) < 1e-12


def test_tiled_tables_basis_path():
    """K = 3 basis beams staged together, all six pair products from shared memory."""
    from fftvis_b200 import synth
    beams = [synth.synthetic_uvbeam(FREQS, naz=72, nza=37, seed=s, perturb=0.05 * s) for s in range(3)]
    nant = len(hex_ants(3))
    rng = np.random.default_rng(0)
    coefs = rng.normal(size=(nant, 3)) + 1j * rng.normal(size=(nant, 3))
    a = _run(True, beams, True, 2, 1, beam_coefs=coefs)
    b = _run(False, beams, True, 2, 1, beam_coefs=coefs)
    assert relerr(a, b) < 1e-12
