"""Parity at the FULL sizes of BASELINE.json configs[2], [3] and [4] in every dimension that changes the
numerics of one (time, frequency) unit -- array, baseline set, catalogue size, products, beam tables,
precision, eps -- on a few of the independent (time, frequency) units, against the term-by-term fp64
measurement equation (oracle.pipeline.simulate_direct) evaluated on a random subset of baselines.
(The time and frequency axes only repeat independent units; bench.py runs all of them.)

Two bars, both stated from north_star's 10 x eps:
* transform parity: the engine and the direct sum are given the SAME per-time equatorial -> ENU rotation
  (``coord_method_params={"rotation_matrices": ...}`` from the oracle's own chain), so the comparison is of
  the horizon cut, beams, coherency and NUFFT: relative L2 error <= 10 x eps;
* whole-chain parity with the two independent coordinate chains (product: core/astrometry.py, oracle:
  oracle/coords.py): the Earth rotation angle at JD 2459845 is only defined to ~1e-14 rad in fp64 (the
  product of 8300 days and the rotation rate rounds at that level in ANY implementation, erfa's included),
  and a rotation error d shows up in a visibility phase as 2 pi f |b| d / c, ~1.2e3 d for a 292 m baseline
  at 200 MHz; the bar there is 10 x eps + kappa x 4e-14 with kappa = 2 pi f_max |b|_max / c.

cfg3: HERA-331 hex, 631 unique baselines, 100k sources, P = 4, az/za E-field table, f64 eps 1e-13 (w = 14,
      fine grid 90 x 90 held whole in one CTA: the small-grid spreader) and f32.
cfg4: 256 random antennas, all 32 640 baselines, 3 145 728 diffuse pixels (1.57 M above the horizon),
      3-D type 3, f64 eps 1e-13 (tiled spreader + pruned FFT at 1620 x 1620 x 30) at the top frequency.
cfg5: 128 antennas, all 8 256 baselines (autos included), K = 5 basis beams with complex coefficients,
      100k sources, P = 4: the basis path (15 transforms per unit + contraction)."""
import numpy as np
import pytest

from gpu_helpers import f32_bar, relerr

pytestmark = pytest.mark.gpu

T0 = np.array([2459845.0])


def _shared_rotation():
    from fftvis_b200 import HERA_LOCATION
    from oracle import coords as ocoords
    return dict(coord_method="CoordinateRotationERA",
                coord_method_params={"rotation_matrices": ocoords.rotation_matrices(T0, HERA_LOCATION)})


def _kappa(ants, fmax):
    pos = np.array(list(ants.values()))
    bmax = np.sqrt(((pos[:, None, :] - pos[None, :, :]) ** 2).sum(-1)).max()
    return 2 * np.pi * fmax * bmax / 299792458.0


@pytest.mark.parametrize("precision", [2, 1])
def test_cfg3_full_size_subsampled_direct_sum(precision):
    from fftvis_b200 import HERA_LOCATION, simulate_vis, synth
    from oracle import pipeline
    ants = synth.hex_array(11)
    assert len(ants) == 331
    freqs = np.array([100e6, 200e6])                       # both ends of the band (largest phases at 200 MHz)
    ra, dec, flux = synth.random_sky(100000, freqs, seed=42, kind="gleam")
    beam = synth.synthetic_uvbeam(freqs, naz=360, nza=181)
    kw = dict(precision=precision, polarized=True, beam_spline_opts={"order": 1}, **_shared_rotation())
    got = simulate_vis(ants, flux, ra, dec, freqs, T0, beam, HERA_LOCATION, **kw)
    assert got.shape == (2, 1, 2, 2, 631)
    direct = pipeline.simulate_direct(ants, flux, ra, dec, freqs, T0, [beam], HERA_LOCATION, **kw)
    if precision == 2:
        assert relerr(got, direct) < 10 * 1e-13
        # whole chain, independent coordinate managers (default ERFA-structured model on both sides)
        kw2 = dict(precision=2, polarized=True, beam_spline_opts={"order": 1})
        got2 = simulate_vis(ants, flux, ra, dec, freqs, T0, beam, HERA_LOCATION, **kw2)
        direct2 = pipeline.simulate_direct(ants, flux, ra, dec, freqs, T0, [beam], HERA_LOCATION, **kw2)
        assert relerr(got2, direct2) < 10 * 1e-13 + _kappa(ants, freqs.max()) * 4e-14
    else:
        cpu = pipeline.simulate_cpu(ants, flux, ra, dec, freqs, T0, [beam], HERA_LOCATION, **kw)
        assert relerr(got, direct) < f32_bar(cpu, direct)
        assert relerr(got, cpu) < 2 * f32_bar(cpu, direct)


def test_cfg4_full_size_subsampled_direct_sum():
    from fftvis_b200 import GaussianBeam, HERA_LOCATION, simulate_vis, synth
    from oracle import pipeline
    ants = synth.random_array(256, radius=150.0, zspan=2.0, seed=42)
    freqs = np.array([200e6])
    ra, dec, flux = synth.random_sky(3145728, freqs, seed=42, kind="diffuse")
    beam = GaussianBeam(diameter=14.0)
    bls = synth.all_baselines(ants)
    rot = _shared_rotation()
    got = simulate_vis(ants, flux, ra, dec, freqs, T0, beam, HERA_LOCATION, baselines=bls, precision=2, **rot)
    assert got.shape == (1, 1, 32640) and np.isfinite(got).all()
    rng = np.random.default_rng(5)
    idx = sorted(rng.choice(len(bls), 96, replace=False))
    direct = pipeline.simulate_direct(ants, flux, ra, dec, freqs, T0, [beam.to_power()], HERA_LOCATION,
                                      baselines=[bls[i] for i in idx], precision=2, **rot)
    assert relerr(got[..., idx], direct) < 10 * 1e-13


def test_cfg5_full_array_basis_path_subsampled_direct_sum():
    from fftvis_b200 import HERA_LOCATION, simulate_vis, synth
    from oracle import pipeline
    ants = synth.hex_array(7)
    ants[len(ants)] = np.array([7 * synth.HEX_SPACING, 0.0, 0.0])
    assert len(ants) == 128
    freqs = np.array([100e6, 200e6])
    K = 5
    basis = [synth.synthetic_uvbeam(freqs, naz=360, nza=181, seed=s, perturb=0.3) for s in range(K)]
    for b in basis:
        b.data_array = b.data_array.real.astype(complex)   # real basis beams
    rng = np.random.default_rng(42)
    coefs = rng.normal(size=(128, K, 2)) + 1j * rng.normal(size=(128, K, 2))
    ra, dec, flux = synth.random_sky(100000, freqs, seed=42, kind="gleam")
    bls = synth.all_baselines(ants, autos=True)
    assert len(bls) == 8256
    kw = dict(precision=2, polarized=True, beam_spline_opts={"order": 1}, beam_coefs=coefs, **_shared_rotation())
    got = simulate_vis(ants, flux, ra, dec, freqs, T0, basis, HERA_LOCATION, baselines=bls, **kw)
    assert got.shape == (2, 1, 2, 2, 8256)
    idx = sorted(rng.choice(len(bls), 64, replace=False))
    direct = pipeline.simulate_direct(ants, flux, ra, dec, freqs, T0, basis, HERA_LOCATION,
                                      baselines=[bls[i] for i in idx], **kw)
    assert relerr(got[..., idx], direct) < 10 * 1e-13
