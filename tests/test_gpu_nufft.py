"""GPU NUFFTs through the C ABI vs the oracle (fp64 direct sum + CPU NUFFT restatement).
Tolerance: relative L2 error <= 10 x eps in fp64 (north_star); in fp32 the floor is the rounding
of the *inputs* (phases up to ~60 rad carry ~4e-6 rad of fp32 rounding), so the bar there is the
oracle's own fp32 error x 3 (SURVEY.md section 7 "fp32 parity floor")."""
import numpy as np
import pytest

from gpu_helpers import f32_bar, relerr

pytestmark = pytest.mark.gpu


def _types(prec):
    return (np.float32, np.complex64) if prec == 1 else (np.float64, np.complex128)


@pytest.mark.parametrize("prec,eps", [(2, 1e-13), (2, 1e-10), (2, 1e-6), (1, 6e-8), (1, 1e-4)])
@pytest.mark.parametrize("upsamp", [2.0, 1.25])
@pytest.mark.parametrize("ntr", [1, 4])
@pytest.mark.parametrize("method", ["fused", "cufft"])
def test_type1_vs_oracle(prec, eps, upsamp, ntr, method):
    from fftvis_b200.gpu import gpu_nufft2d_type1
    from oracle import nufft_cpu as nc
    rng = np.random.default_rng(0)
    n, N = 700, 21
    rd, cd = _types(prec)
    x = rng.uniform(-40, 40, n).astype(rd)
    y = rng.uniform(-40, 40, n).astype(rd)
    c = (rng.normal(size=(ntr, n)) + 1j * rng.normal(size=(ntr, n))).astype(cd)
    idx = rng.integers(-(N // 2), N // 2 + 1, size=(2, 97))
    got = gpu_nufft2d_type1(x, y, c, N, idx, eps, upsample_factor=upsamp, method=method)
    assert got.shape == (ntr, 97) and got.dtype == cd
    want = nc.direct_sum(x, y, None, c, idx[0], idx[1], None)
    cpu = nc.cpu_nufft2d_type1(x, y, c, N, idx, eps, upsamp)
    floor = 1e-9 if upsamp == 1.25 else 0.0
    if prec == 2:
        assert relerr(got, want) < max(10 * eps, floor)
    else:
        assert relerr(got, want) < f32_bar(cpu, want, eps)
        # directly against the CPU path in the same precision: the two differ by at most the sum of
        # their own fp32 rounding errors
        assert relerr(got, cpu) < 2 * f32_bar(cpu, want, eps)
        return
    assert relerr(got, cpu) < max(10 * eps, floor)


@pytest.mark.parametrize("prec,eps", [(2, 1e-13), (2, 1e-10), (1, 6e-8)])
@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("upsamp", [2.0, 1.25])
def test_type3_vs_oracle(prec, eps, dim, upsamp):
    from fftvis_b200.gpu import gpu_nufft2d, gpu_nufft3d
    from oracle import nufft_cpu as nc
    rng = np.random.default_rng(1)
    n, nk = 900, 130
    rd, cd = _types(prec)
    lm = rng.uniform(-0.7, 0.7, (2, n))
    nn = np.sqrt(1 - (lm**2).sum(0))
    xs = [(2 * np.pi * a).astype(rd) for a in (lm[0], lm[1], nn)][:dim]
    ss = [rng.uniform(-12, 12, nk).astype(rd), rng.uniform(-12, 12, nk).astype(rd),
          rng.uniform(-0.5, 0.5, nk).astype(rd)][:dim]
    c = (rng.normal(size=(2, n)) + 1j * rng.normal(size=(2, n))).astype(cd)
    if dim == 2:
        got = gpu_nufft2d(xs[0], xs[1], c, ss[0], ss[1], eps, upsample_factor=upsamp)
    else:
        got = gpu_nufft3d(xs[0], xs[1], xs[2], c, ss[0], ss[1], ss[2], eps, upsample_factor=upsamp)
    want = nc.direct_sum(xs[0], xs[1], xs[2] if dim == 3 else None, c, ss[0], ss[1],
                         ss[2] if dim == 3 else None)
    floor = 1e-9 if upsamp == 1.25 else 0.0
    if prec == 2:
        tol = max(10 * eps, floor)
    else:
        # fp32: the yardstick is the CPU restatement's own fp32 error on the same inputs (sigma = 1.25
        # amplifies fp32 rounding through the larger 1/phihat deconvolution factors)
        cpu = nc.nufft_type3(xs, c, ss, eps, upsampfac=upsamp)
        tol = f32_bar(cpu, want, eps)
        assert relerr(got, cpu) < 2 * tol
    assert got.shape == want.shape
    assert relerr(got, want) < tol


@pytest.mark.parametrize("n_modes,rows", [(465, 0), (465, 7), (121, 0), (41, 0), (41, 16), (7, 0), (251, 5)])
@pytest.mark.parametrize("prec", [1, 2])
def test_type1_fused_grid_sizes_and_strip_heights(n_modes, rows, prec):
    """Fused path on the grid sizes of the BASELINE configs (nf = 960, 250, 90, 30, 512) with
    automatic and forced strip heights (strips that do not divide nf, single-strip grids), against
    the direct sum and the cuFFT path."""
    from fftvis_b200.gpu import gpu_nufft2d_type1
    from fftvis_b200.gpu.nufft import default_plan
    from oracle import nufft_cpu as nc
    rng = np.random.default_rng(n_modes + rows)
    n, nk = 3000, 400
    rd, cd = _types(prec)
    eps = 6e-8 if prec == 1 else 1e-12
    x = rng.uniform(-60, 60, n).astype(rd)
    y = rng.uniform(-60, 60, n).astype(rd)
    c = (rng.normal(size=(2, n)) + 1j * rng.normal(size=(2, n))).astype(cd)
    h = n_modes // 2
    idx = rng.integers(-h, h + 1, size=(2, nk))
    idx[:, :4] = [[-h, h, 0, h], [h, -h, 0, h]]
    plan = default_plan()
    plan.set_option("t1_rows", rows)
    try:
        got = gpu_nufft2d_type1(x, y, c, n_modes, idx, eps, method="fused")
    finally:
        plan.set_option("t1_rows", 0)
    ref = gpu_nufft2d_type1(x, y, c, n_modes, idx, eps, method="cufft")
    want = nc.direct_sum(x, y, None, c, idx[0], idx[1], None)
    if prec == 2:
        assert relerr(got, want) < 10 * eps
        assert relerr(got, ref) < 10 * eps
    else:
        cpu = nc.cpu_nufft2d_type1(x, y, c, n_modes, idx, eps, 2.0)
        assert relerr(got, want) < f32_bar(cpu, want, eps)
        assert relerr(got, cpu) < 2 * f32_bar(cpu, want, eps)


def test_type3_offcentre_points_and_targets():
    """Non-zero centres exercise the pre- and post-phases."""
    from fftvis_b200.gpu import gpu_nufft3d
    from oracle import nufft_cpu as nc
    rng = np.random.default_rng(5)
    n, nk = 500, 77
    x, y = rng.uniform(1.0, 3.0, n), rng.uniform(-2.0, -0.5, n)
    z = rng.uniform(4.0, 6.0, n)
    u, v, w = rng.uniform(20, 30, nk), rng.uniform(-40, -25, nk), rng.uniform(3, 5, nk)
    c = rng.normal(size=n) + 1j * rng.normal(size=n)
    got = gpu_nufft3d(x, y, z, c, u, v, w, 1e-12)
    want = nc.direct_sum(x, y, z, c, u, v, w)
    assert got.shape == (nk,)
    assert relerr(got, want) < 1e-11


def test_single_point_and_single_target():
    from fftvis_b200.gpu import gpu_nufft2d
    got = gpu_nufft2d(np.array([0.3]), np.array([-0.2]), np.array([2.0 + 1j]), np.array([5.0]),
                      np.array([-3.0]), 1e-12)
    want = (2.0 + 1j) * np.exp(1j * (5.0 * 0.3 + 3.0 * 0.2))
    assert abs(got[0] - want) < 1e-10


def test_empty_inputs():
    from fftvis_b200.gpu import gpu_nufft2d, gpu_nufft2d_type1
    e = np.zeros(0)
    out = gpu_nufft2d(e, e, np.zeros(0, complex), np.array([1.0, 2.0]), np.array([0.5, 0.1]), 1e-10)
    assert out.shape == (2,) and np.all(out == 0)
    out = gpu_nufft2d_type1(e, e, np.zeros((4, 0), complex), 7, np.array([[0, 1], [2, -3]]), 1e-10)
    assert out.shape == (4, 2) and np.all(out == 0)


def test_type1_mode_out_of_range_raises():
    from fftvis_b200.gpu import gpu_nufft2d_type1
    x = np.array([0.1, 0.2])
    with pytest.raises(IndexError):
        gpu_nufft2d_type1(x, x, np.ones(2, complex), 7, np.array([[9], [0]]), 1e-10)


def test_frequency_batched_type1_and_type3_match_one_by_one():
    """The batch dimension (frequencies sharing one source set) against per-frequency calls."""
    import torch
    from fftvis_b200.gpu import _lib
    from fftvis_b200.gpu.nufft import default_plan
    from oracle import nufft_cpu as nc
    rng = np.random.default_rng(7)
    n, nk, nb, ntr = 1500, 64, 5, 4
    bx, by = rng.uniform(-0.3, 0.3, n), rng.uniform(-0.3, 0.3, n)
    scale = np.linspace(100.0, 200.0, nb)
    W = rng.normal(size=(nb, ntr, n)) + 1j * rng.normal(size=(nb, ntr, n))
    m = rng.integers(-10, 11, size=(2, nk))
    dev = "cuda"
    t = lambda a, dt: torch.as_tensor(np.ascontiguousarray(a)).to(dev, dt)
    n_dev = torch.tensor([n], dtype=torch.int32, device=dev)
    out = torch.zeros((nb, ntr, nk), dtype=torch.complex128, device=dev)
    epi = _lib.make_epilogue(out.data_ptr(), out.stride(0), out.stride(1))
    plan = default_plan()
    plan.type1(2, t(bx, torch.float64), t(by, torch.float64), n_dev, scale, t(W, torch.complex128), 21,
               t(m[0], torch.int32), t(m[1], torch.int32), 1e-12, 2.0, epi)
    got = out.cpu().numpy()
    from fftvis_b200.gpu.nufft import ModeSet
    out2 = torch.zeros_like(out)
    epi2 = _lib.make_epilogue(out2.data_ptr(), out2.stride(0), out2.stride(1))
    modes = ModeSet(m[0], m[1], 21)
    plan.type1_fused(2, t(bx, torch.float64), t(by, torch.float64), n_dev, scale, t(W, torch.complex128), modes,
                     1e-12, 2.0, epi2)
    got2 = out2.cpu().numpy()
    for b in range(nb):
        want = nc.direct_sum(bx * scale[b], by * scale[b], None, W[b], m[0], m[1], None)
        assert relerr(got[b], want) < 1e-11
        assert relerr(got2[b], want) < 1e-11
    # type 3: targets scale with the frequency
    x = [2 * np.pi * rng.uniform(-0.6, 0.6, n) for _ in range(2)]
    u = [rng.uniform(-0.08, 0.08, nk) for _ in range(2)]
    out.zero_()
    plan.type3(2, 2, [t(a, torch.float64) for a in x], n_dev, None, [t(a, torch.float64) for a in u], None,
               scale, t(W, torch.complex128), 1e-12, 2.0, epi)
    got = out.cpu().numpy()
    for b in range(nb):
        want = nc.direct_sum(x[0], x[1], None, W[b], u[0] * scale[b], u[1] * scale[b], None)
        assert relerr(got[b], want) < 1e-11


@pytest.mark.parametrize("prec,eps", [(2, 1e-12), (2, 1e-13), (1, 6e-8)])
def test_type3_3d_tiled_spreader_matches_atomic_spreader_and_direct_sum(prec, eps):
    """Thin-z 3-D grids use the bin-sorted column-tile spreader (no atomics); it must agree with the
    global-atomics spreader and the direct sum, over a frequency batch with off-centre targets."""
    import torch
    from fftvis_b200.gpu import _lib
    from fftvis_b200.gpu.nufft import default_plan
    from oracle import nufft_cpu as nc
    rng = np.random.default_rng(12)
    n, nk, nb, ntr = 20000, 300, 3, 2
    rd, cd = _types(prec)
    rdt, cdt = (torch.float32, torch.complex64) if prec == 1 else (torch.float64, torch.complex128)
    lm = rng.uniform(-0.7, 0.7, (2, n))
    x = [(2 * np.pi * v).astype(rd) for v in (lm[0], lm[1], np.sqrt(1 - (lm**2).sum(0)))]
    u = [rng.uniform(-3e-7, 5e-7, nk).astype(rd), rng.uniform(-4e-7, 4e-7, nk).astype(rd),
         rng.uniform(-6e-9, 6e-9, nk).astype(rd)]
    scale = np.array([1.0e8, 1.01e8, 1.02e8])
    W = (rng.normal(size=(nb, ntr, n)) + 1j * rng.normal(size=(nb, ntr, n))).astype(cd)
    t = lambda a, dt: torch.as_tensor(np.ascontiguousarray(a)).to("cuda", dt)
    xs, us, Wd = [t(a, rdt) for a in x], [t(a, rdt) for a in u], t(W, cdt)
    n_dev = torch.tensor([n], dtype=torch.int32, device="cuda")
    plan = default_plan()
    outs = []
    for tiles in (1, 0, 1, 1):
        plan.set_option("t3_tiles", tiles)
        out = torch.zeros((nb, ntr, nk), dtype=cdt, device="cuda")
        epi = _lib.make_epilogue(out.data_ptr(), out.stride(0), out.stride(1))
        plan.type3(prec, 3, xs, n_dev, None, us, None, scale, Wd, eps, 2.0, epi)
        outs.append(out.cpu().numpy())
    plan.set_option("t3_tiles", 1)
    # tile lists are sorted by source index: the tiled spreader's sums are bitwise reproducible
    assert np.array_equal(outs[0], outs[2]) and np.array_equal(outs[0], outs[3])
    for b in range(nb):
        uu = [(a * rd(scale[b])).astype(rd) for a in u]
        want = nc.direct_sum(x[0], x[1], x[2], W[b], uu[0], uu[1], uu[2])
        tol = 10 * eps if prec == 2 else f32_bar(nc.nufft_type3(x, W[b], uu, eps), want, eps)
        assert relerr(outs[0][b], want) < tol
        assert relerr(outs[1][b], want) < tol
        assert relerr(outs[0][b], outs[1][b]) < (1e-12 if prec == 2 else 2 * tol)


@pytest.mark.parametrize("prec,dim", [(2, 2), (2, 3), (1, 2), (1, 3)])
def test_type3_pruned_fft_matches_cufft_path(prec, dim):
    """Inner FFT of type 3: own pruned shared-memory passes (deconvolution fused) vs cuFFT on the
    padded grid, and both against the direct sum, on a frequency batch."""
    import torch
    from fftvis_b200.gpu import _lib
    from fftvis_b200.gpu.nufft import default_plan
    from oracle import nufft_cpu as nc
    rng = np.random.default_rng(40 + dim)
    n, nk, nb, ntr = 4000, 257, 2, 2
    rd, cd = _types(prec)
    rdt, cdt = (torch.float32, torch.complex64) if prec == 1 else (torch.float64, torch.complex128)
    eps = 6e-8 if prec == 1 else 1e-12
    lm = rng.uniform(-0.7, 0.7, (2, n))
    x = [(2 * np.pi * v).astype(rd) for v in (lm[0], lm[1], np.sqrt(1 - (lm**2).sum(0)))][:dim]
    u = [rng.uniform(-2e-7, 3e-7, nk).astype(rd), rng.uniform(-3e-7, 3e-7, nk).astype(rd),
         rng.uniform(-5e-9, 5e-9, nk).astype(rd)][:dim]
    scale = np.array([1.0e8, 1.3e8])
    W = (rng.normal(size=(nb, ntr, n)) + 1j * rng.normal(size=(nb, ntr, n))).astype(cd)
    t = lambda a, dt: torch.as_tensor(np.ascontiguousarray(a)).to("cuda", dt)
    xs, us, Wd = [t(a, rdt) for a in x], [t(a, rdt) for a in u], t(W, cdt)
    if dim == 2:
        xs, us = xs + [xs[0]], us + [us[0]]          # placeholders (ignored for dim = 2)
    n_dev = torch.tensor([n], dtype=torch.int32, device="cuda")
    plan = default_plan()
    outs = []
    # own passes in the half-length form (where ng = 2 nf: two nf-point transforms per padded vector), own passes
    # at full length, cuFFT on the padded grid
    for own, half in ((2, 1), (2, 0), (0, 0)):
        plan.set_option("t3_fft", own)
        plan.set_option("t3_half", half)
        out = torch.zeros((nb, ntr, nk), dtype=cdt, device="cuda")
        epi = _lib.make_epilogue(out.data_ptr(), out.stride(0), out.stride(1))
        plan.type3(prec, dim, xs, n_dev, None, us, None, scale, Wd, eps, 2.0, epi)
        outs.append(out.cpu().numpy())
    geo = plan.last_type3_geometry()
    plan.set_option("t3_fft", 1)
    plan.set_option("t3_half", 1)
    # the half-length form really ran in at least one dimension of this case
    assert any(g == 2 * f and f % 2 == 0 for f, g in zip(geo["nf"], geo["ng"])), geo
    for b in range(nb):
        uu = [(a * rd(scale[b])).astype(rd) for a in u]
        want = nc.direct_sum(x[0], x[1], x[2] if dim == 3 else None, W[b], uu[0], uu[1], uu[2] if dim == 3 else None)
        tol = 10 * eps if prec == 2 else f32_bar(nc.nufft_type3(x, W[b], uu, eps), want, eps)
        for o in outs:
            assert relerr(o[b], want) < tol
        assert relerr(outs[0][b], outs[2][b]) < (1e-12 if prec == 2 else 2 * tol)
        assert relerr(outs[0][b], outs[1][b]) < (1e-12 if prec == 2 else 2 * tol)


def test_epilogue_conj_kmap_pmap_accumulate():
    import torch
    from fftvis_b200.gpu import _lib
    from fftvis_b200.gpu.nufft import default_plan
    from oracle import nufft_cpu as nc
    rng = np.random.default_rng(8)
    n, nk = 300, 10
    x = [2 * np.pi * rng.uniform(-0.6, 0.6, n) for _ in range(2)]
    u = [rng.uniform(-8, 8, nk) for _ in range(2)]
    W = rng.normal(size=(1, 4, n)) + 1j * rng.normal(size=(1, 4, n))
    kmap = rng.permutation(16)[:nk].astype(np.int32)
    conj = (rng.uniform(size=nk) < 0.5).astype(np.uint8)
    dev = "cuda"
    t = lambda a, dt: torch.as_tensor(np.ascontiguousarray(a)).to(dev, dt)
    out = torch.ones((1, 4, 16), dtype=torch.complex128, device=dev)
    km, cj = t(kmap, torch.int32), t(conj, torch.uint8)
    epi = _lib.make_epilogue(out.data_ptr(), out.stride(0), out.stride(1), (0, 2, 1, 3), km.data_ptr(),
                             cj.data_ptr(), accumulate=True)
    n_dev = torch.tensor([n], dtype=torch.int32, device=dev)
    default_plan().type3(2, 2, [t(a, torch.float64) for a in x], n_dev, None,
                         [t(a, torch.float64) for a in u], None, [1.0], t(W, torch.complex128), 1e-12, 2.0, epi)
    got = out.cpu().numpy()[0]
    ref = nc.direct_sum(x[0], x[1], None, W[0], u[0], u[1], None)
    ref = np.where(conj[None, :].astype(bool), ref.conj(), ref)
    want = np.ones((4, 16), complex)
    for p, slot in enumerate((0, 2, 1, 3)):
        want[slot, kmap] += ref[p]
    np.testing.assert_allclose(got, want, rtol=0, atol=1e-9)


def test_gpu_direct_sum_matches_oracle():
    import ctypes
    import torch
    from fftvis_b200.gpu import _lib
    from oracle import nufft_cpu as nc
    rng = np.random.default_rng(9)
    n, nk = 257, 33
    x = [rng.uniform(-3, 3, n) for _ in range(3)]
    u = [rng.uniform(-20, 20, nk) for _ in range(3)]
    W = rng.normal(size=(2, 2, n)) + 1j * rng.normal(size=(2, 2, n))
    t = lambda a, dt: torch.as_tensor(np.ascontiguousarray(a)).to("cuda", dt)
    xs, us = [t(a, torch.float64) for a in x], [t(a, torch.float64) for a in u]
    Wd = t(W, torch.complex128)
    out = torch.zeros((2, 2, nk), dtype=torch.complex128, device="cuda")
    epi = _lib.make_epilogue(out.data_ptr(), out.stride(0), out.stride(1))
    n_dev = torch.tensor([n], dtype=torch.int32, device="cuda")
    _lib.check(_lib.lib().fv_direct_sum(2, 3, xs[0].data_ptr(), xs[1].data_ptr(), xs[2].data_ptr(),
                                        n_dev.data_ptr(), n, us[0].data_ptr(), us[1].data_ptr(),
                                        us[2].data_ptr(), nk, _lib.doubles([1.0, 0.5]), 2, 2, Wd.data_ptr(),
                                        ctypes.byref(epi), torch.cuda.current_stream().cuda_stream))
    got = out.cpu().numpy()
    for b, s in enumerate((1.0, 0.5)):
        want = nc.direct_sum(x[0], x[1], x[2], W[b], u[0] * s, u[1] * s, u[2] * s)
        assert relerr(got[b], want) < 1e-13


@pytest.mark.parametrize("prec,eps", [(2, 1e-13), (2, 1e-10), (2, 1e-6), (1, 6e-8), (1, 1e-6)])
@pytest.mark.parametrize("n_modes,ntr,upsamp", [(41, 4, 2.0), (41, 9, 2.0), (33, 1, 2.0), (61, 2, 1.25)])
def test_type1_small_grid_register_window_path(prec, eps, n_modes, ntr, upsamp):
    """Small-grid path (type1_small.cuh: sources bin-sorted into 2 x 2 origin bins, the bin's window in
    registers, phases of disjoint windows): forced on, against the direct sum, the strip kernel and itself
    (bit-reproducible), with more than four transforms per frequency (batched basis pairs) and grid sizes
    whose bins do not divide evenly into classes."""
    from fftvis_b200.gpu import gpu_nufft2d_type1
    from fftvis_b200.gpu.nufft import default_plan
    from oracle import nufft_cpu as nc
    rng = np.random.default_rng(n_modes + ntr)
    n, nk = 20000, 300
    rd, cd = _types(prec)
    x = rng.uniform(-60, 60, n).astype(rd)
    y = rng.uniform(-60, 60, n).astype(rd)
    c = (rng.normal(size=(ntr, n)) + 1j * rng.normal(size=(ntr, n))).astype(cd)
    h = n_modes // 2
    idx = rng.integers(-h, h + 1, size=(2, nk))
    idx[:, :4] = [[-h, h, 0, h], [h, -h, 0, h]]
    plan = default_plan()
    outs = []
    try:
        for small in (2, 2, 0):
            plan.set_option("t1_small", small)
            outs.append(gpu_nufft2d_type1(x, y, c, n_modes, idx, eps, upsample_factor=upsamp, method="fused"))
    finally:
        plan.set_option("t1_small", 1)
    assert np.array_equal(outs[0], outs[1])
    want = nc.direct_sum(x, y, None, c, idx[0], idx[1], None)
    floor = 1e-9 if upsamp == 1.25 else 0.0
    if prec == 2:
        assert relerr(outs[0], want) < max(10 * eps, floor)
        assert relerr(outs[0], outs[2]) < max(10 * eps, floor)
    else:
        cpu = nc.cpu_nufft2d_type1(x, y, c, n_modes, idx, eps, upsamp)
        assert relerr(outs[0], want) < f32_bar(cpu, want, eps)
        assert relerr(outs[0], cpu) < 2 * f32_bar(cpu, want, eps)


@pytest.mark.parametrize("eps", [6e-8, 1e-5, 1e-3, 1e-1])
@pytest.mark.parametrize("n_modes,ntr,nk", [(465, 1, 700), (251, 4, 300), (999, 1, 900), (131, 2, 200)])
def test_type1_xdirect_pass1(eps, n_modes, ntr, nk):
    """Single-precision pass 1 without an x grid (type1_xdirect.cuh) against the direct sum, the CPU
    restatement and the strip kernel it replaces: kernel widths from 2 to 9, more than 256 needed columns
    (two column groups), more than 64 strips (nf = 2000: row compare instead of strip masks), strips that do
    not divide nf (nf = 270), sources whose footprint wraps around both grid edges."""
    from fftvis_b200.gpu import gpu_nufft2d_type1
    from fftvis_b200.gpu.nufft import default_plan
    from oracle import nufft_cpu as nc
    rng = np.random.default_rng(n_modes + ntr)
    n = 4000
    x = rng.uniform(-60, 60, n).astype(np.float32)
    y = rng.uniform(-60, 60, n).astype(np.float32)
    # footprints across the periodic edge of the fine grid in x and y
    x[:8] = np.float32(np.pi) * np.array([1, -1, 1, -1, 3, -3, 1, 1], np.float32) + np.float32(1e-3) * np.arange(-4, 4, dtype=np.float32)
    y[4:12] = np.float32(np.pi) * np.array([1, -1, 1, -1, 3, -3, 5, -5], np.float32) + np.float32(2e-3) * np.arange(-4, 4, dtype=np.float32)
    c = (rng.normal(size=(ntr, n)) + 1j * rng.normal(size=(ntr, n))).astype(np.complex64)
    h = n_modes // 2
    idx = rng.integers(-h, h + 1, size=(2, nk))
    idx[:, :4] = [[-h, h, 0, h], [h, -h, 0, h]]
    plan = default_plan()
    got = gpu_nufft2d_type1(x, y, c, n_modes, idx, eps, method="fused")
    plan.set_option("t1_xdirect", 0)
    try:
        old = gpu_nufft2d_type1(x, y, c, n_modes, idx, eps, method="fused")
    finally:
        plan.set_option("t1_xdirect", 1)
    want = nc.direct_sum(x, y, None, c, idx[0], idx[1], None)
    cpu = nc.cpu_nufft2d_type1(x, y, c, n_modes, idx, eps, 2.0)
    bar = f32_bar(cpu, want, eps)
    assert relerr(got, want) < bar
    assert relerr(got, cpu) < 2 * bar
    assert relerr(got, old) < 2 * bar
    assert not np.array_equal(got, old)          # the two pass-1 kernels really are different code paths
    again = gpu_nufft2d_type1(x, y, c, n_modes, idx, eps, method="fused")
    assert np.array_equal(got, again)            # fixed summation order: bitwise reproducible


@pytest.mark.parametrize("extent", [0.4, 1.2, 3.0])
def test_type3_3d_tiny_grids_wrapping_footprints(extent):
    """Very short baselines give 3-D type-3 grids of 16..32 cells per side, where a kernel footprint wraps
    around the grid and re-enters its first tile: every source must still be spread exactly once."""
    from fftvis_b200.gpu import gpu_nufft3d
    from oracle import nufft_cpu as nc
    rng = np.random.default_rng(int(extent * 10))
    n, nk, eps = 3000, 40, 1e-9
    lm = rng.uniform(-0.7, 0.7, (2, n))
    x = [2 * np.pi * lm[0], 2 * np.pi * lm[1], 2 * np.pi * np.sqrt(1 - (lm**2).sum(0))]
    u = [rng.uniform(-extent, extent, nk), rng.uniform(-extent, extent, nk), rng.uniform(-0.05, 0.05, nk)]
    c = rng.normal(size=(1, n)) + 1j * rng.normal(size=(1, n))
    got = gpu_nufft3d(x[0], x[1], x[2], c, u[0], u[1], u[2], eps)
    want = nc.direct_sum(x[0], x[1], x[2], c, u[0], u[1], u[2])
    assert relerr(got, want) < 10 * eps
