"""The oracle pinned against what can be pinned offline:
  * the reference's coherency kernels (golden vectors compiled from cpu/beams.py:129-246) and
    its einsum identities (reference tests/test_cpu_beams.py:102,606,870,953);
  * the CPU NUFFT restatement against the fp64 direct sum, to the requested eps."""
import numpy as np
import pytest

from oracle import beams as ob
from oracle import nufft_cpu as nc


def relerr(a, b):
    return np.linalg.norm(np.ravel(a) - np.ravel(b)) / np.linalg.norm(np.ravel(b))


def test_coherency_kernels_match_reference_golden(golden):
    g, _ = golden
    bi, bj, flux, coh = g["coh/beam_i"], g["coh/beam_j"], g["coh/flux"], g["coh/coherency"]
    np.testing.assert_allclose(ob.apparent_polarized_beam(bi, flux), g["coh/out_polarized_beam"], rtol=1e-12)
    np.testing.assert_allclose(ob.apparent_polarized(bi, coh), g["coh/out_polarized"], rtol=1e-12)
    np.testing.assert_allclose(ob.apparent_polarized_beam_pair(bi, bj, flux), g["coh/out_beam_pair"], rtol=1e-12)
    np.testing.assert_allclose(ob.apparent_polarized_pair(bi, bj, coh), g["coh/out_pair"], rtol=1e-12)


def test_coherency_known_answer_from_reference_tests():
    # reference tests/test_cpu_beams.py:99-109
    beam = np.arange(12).reshape((2, 2, 3)).astype(complex)
    flux = np.arange(3).astype(float)
    want = np.einsum("bas,s,bcs->acs", beam.conj(), flux, beam)
    np.testing.assert_allclose(ob.apparent_polarized_beam(beam, flux), want)


def test_kernel_params():
    assert nc.kernel_params(6e-8, 2.0, 1)[0] == 9
    assert nc.kernel_params(1e-10, 2.0, 2)[0] == 11
    assert nc.kernel_params(1e-13, 2.0, 2)[0] == 14
    assert nc.kernel_params(1e-12, 2.0, 2)[0] == 13
    assert nc.kernel_params(6e-8, 1.25, 1)[0] == 12
    assert nc.next235even(7 * 2) == 16 and nc.next235even(82) == 90 and nc.next235even(28) == 30
    assert nc.type1_grid_size(7, 14) == 30 and nc.type1_grid_size(41, 9) == 90


@pytest.mark.parametrize("prec,eps", [(2, 1e-13), (2, 1e-10), (2, 1e-6), (1, 6e-8), (1, 1e-4)])
@pytest.mark.parametrize("upsamp", [2.0, 1.25])
def test_type1_vs_direct(prec, eps, upsamp):
    rng = np.random.default_rng(0)
    n, N = 500, 21
    rd, cd = (np.float32, np.complex64) if prec == 1 else (np.float64, np.complex128)
    x = rng.uniform(-40, 40, n).astype(rd)
    y = rng.uniform(-40, 40, n).astype(rd)
    c = (rng.normal(size=(3, n)) + 1j * rng.normal(size=(3, n))).astype(cd)
    F = nc.nufft2d1(x, y, c, N, eps, upsampfac=upsamp)
    k = np.fft.ifftshift(np.arange(-(N // 2), N // 2 + 1))
    k1, k2 = np.meshgrid(k, k, indexing="ij")
    want = nc.direct_sum(x, y, None, c, k1.ravel(), k2.ravel(), None).reshape(3, N, N)
    tol = 10 * eps if prec == 2 else max(10 * eps, 3e-5)   # f32 folding of |x|~40 rad
    if upsamp == 1.25:
        tol = max(tol, 1e-9)    # w is clamped at 16: sigma=1.25 cannot reach below ~1e-10
    assert relerr(F, want) < tol


@pytest.mark.parametrize("prec,eps", [(2, 1e-13), (2, 1e-10), (1, 6e-8)])
@pytest.mark.parametrize("dim", [2, 3])
def test_type3_vs_direct(prec, eps, dim):
    rng = np.random.default_rng(1)
    n, nk = 400, 60
    rd, cd = (np.float32, np.complex64) if prec == 1 else (np.float64, np.complex128)
    lm = rng.uniform(-0.7, 0.7, (2, n))
    nn = np.sqrt(1 - (lm**2).sum(0))
    xs = [(2 * np.pi * a).astype(rd) for a in (lm[0], lm[1], nn)][:dim]
    ss = [rng.uniform(-12, 12, nk).astype(rd), rng.uniform(-12, 12, nk).astype(rd),
          rng.uniform(-0.5, 0.5, nk).astype(rd)][:dim]
    c = (rng.normal(size=(2, n)) + 1j * rng.normal(size=(2, n))).astype(cd)
    got = nc.nufft_type3(xs, c, ss, eps)
    want = nc.direct_sum(xs[0], xs[1], xs[2] if dim == 3 else None, c, ss[0], ss[1],
                         ss[2] if dim == 3 else None)
    tol = 10 * eps if prec == 2 else 2e-5
    assert relerr(got, want) < tol


def test_spread_thread_invariance():
    rng = np.random.default_rng(3)
    n = 300
    pts = [rng.uniform(-np.pi, np.pi, n), rng.uniform(-np.pi, np.pi, n)]
    c = (rng.normal(size=(1, n)) + 1j * rng.normal(size=(1, n)))
    a = nc.spread(pts, c, (40, 64), 7, 2.3 * 7, nthreads=1)
    b = nc.spread(pts, c, (40, 64), 7, 2.3 * 7, nthreads=4)
    np.testing.assert_allclose(a, b, rtol=1e-12, atol=1e-14)
    assert a.shape == (1, 64, 40)


@pytest.mark.parametrize("eps", [1e-12, 1e-6, 1e-3])
def test_xdirect_hybrid_restatement_matches_direct_sum(eps):
    """The algebra of the product's x-direct pass 1 (exact Fourier sum along x, kernel + FFT + deconvolution along
    y only) restated in numpy: never worse than the two-dimensional transform's tolerance."""
    from oracle import nufft_cpu as nc
    rng = np.random.default_rng(11)
    n, nk, N = 400, 60, 41
    x, y = rng.uniform(-30, 30, n), rng.uniform(-30, 30, n)
    x[:4] = [np.pi, -np.pi, 3 * np.pi - 1e-3, 1e-3 - np.pi]                 # footprints across the periodic edge in y too
    y[2:6] = [np.pi - 1e-3, -np.pi, 5 * np.pi, -3 * np.pi + 2e-3]
    c = rng.normal(size=(2, n)) + 1j * rng.normal(size=(2, n))
    idx = rng.integers(-(N // 2), N // 2 + 1, size=(2, nk))
    idx[:, :3] = [[-(N // 2), N // 2, 0], [N // 2, -(N // 2), 0]]
    got = nc.nufft2d1_xdirect(x, y, c, idx[0], idx[1], eps)
    want = nc.direct_sum(x, y, None, c, idx[0], idx[1], None)
    full = nc.cpu_nufft2d_type1(x, y, c, N, idx, eps)
    err = np.linalg.norm(got - want) / np.linalg.norm(want)
    assert err < 10 * eps
    assert err <= 1.5 * np.linalg.norm(full - want) / np.linalg.norm(want) + 1e-14


def test_half_length_form_of_the_padded_fft():
    """X[2k] = (-1)^k FFT_nin(y)[k], X[2k+1] = (-1)^k (-i) FFT_nin(y_j e^{2 pi i j / n})[k] (csrc/type3_fft.cuh) against
    numpy's transform of the explicitly padded vector."""
    from oracle import nufft_cpu as nc
    rng = np.random.default_rng(3)
    for nin in (30, 36, 1620):
        data = rng.normal(size=(3, nin)) + 1j * rng.normal(size=(3, nin))
        n = 2 * nin
        padded = np.zeros((3, n), complex)
        m = np.arange(nin) - nin // 2
        padded[:, m % n] = data
        want = np.fft.ifft(padded, axis=-1) * n                              # sum_j x_j e^{+2 pi i j k / n}
        got = nc.padded_fft_half_length(data)
        assert np.linalg.norm(got - want) / np.linalg.norm(want) < 1e-13
