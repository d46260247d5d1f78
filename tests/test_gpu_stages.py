"""Stages a1-a5 and a10 through the C ABI vs numpy restatements / the reference's own einsum
identities (reference tests/test_cpu_beams.py:102,606,870,953)."""
import ctypes

import numpy as np
import pytest

from gpu_helpers import relerr

pytestmark = pytest.mark.gpu


def _rotate_cut(prec, eq, enu, plane, lo=0, hi=None, n_cap=None, astrom=None):
    import torch
    from fftvis_b200.gpu import _lib
    nsrc = eq.shape[1]
    hi = nsrc if hi is None else hi
    n_cap = nsrc if n_cap is None else n_cap
    rdt = torch.float32 if prec == 1 else torch.float64
    eq_d = torch.as_tensor(np.ascontiguousarray(eq)).cuda()
    xyz = torch.full((3, n_cap), -77.0, dtype=rdt, device="cuda")
    az = torch.zeros(n_cap, dtype=rdt, device="cuda")
    za = torch.zeros(n_cap, dtype=rdt, device="cuda")
    idx = torch.full((n_cap,), -1, dtype=torch.int32, device="cuda")
    n_dev = torch.zeros(1, dtype=torch.int32, device="cuda")
    scratch = torch.empty(int(_lib.lib().fv_rotate_cut_scratch_bytes(nsrc)), dtype=torch.uint8, device="cuda")
    _lib.check(_lib.lib().fv_rotate_cut(prec, eq_d.data_ptr(), nsrc, lo, hi, _lib.doubles(enu.ravel()),
                                        _lib.doubles(astrom) if astrom is not None else None,
                                        _lib.doubles(plane.ravel()), xyz.data_ptr(), az.data_ptr(),
                                        za.data_ptr(), idx.data_ptr(), n_cap, n_dev.data_ptr(),
                                        scratch.data_ptr(), torch.cuda.current_stream().cuda_stream))
    n = int(n_dev.item())
    return n, xyz.cpu().numpy(), az.cpu().numpy(), za.cpu().numpy(), idx.cpu().numpy()


@pytest.mark.parametrize("prec", [1, 2])
@pytest.mark.parametrize("nsrc", [1, 31, 1024, 1025, 40001])
@pytest.mark.parametrize("method,params", [("CoordinateRotationERFA", {}), ("CoordinateRotationERA", {}),
                                           ("CoordinateRotationAstropy", {"dut1": -0.03, "xp": 1e-6, "yp": 2e-6})])
def test_rotate_cut_vs_oracle_chain(prec, nsrc, method, params):
    """The fused rotate / cut / az-za kernel fed by the product's per-time blocks against the ORACLE's
    own coordinate chain (oracle/coords.py shares no code with fftvis_b200.core)."""
    from fftvis_b200.core import coords
    from oracle import coords as ocoords
    rng = np.random.default_rng(nsrc)
    rd = np.float32 if prec == 1 else np.float64
    ra = rng.uniform(0, 2 * np.pi, nsrc).astype(rd)
    dec = np.arcsin(rng.uniform(-1, 1, nsrc)).astype(rd)
    eq = coords.equatorial_unit_vectors(ra, dec)
    tjd = np.array([2459845.3])
    mats, ast = coords.coordinate_blocks(tjd, coords.HERA_LOCATION, method, params)
    enu = mats[0]
    th = 0.01
    plane = np.array([[np.cos(th), 0, np.sin(th)], [0, 1, 0], [-np.sin(th), 0, np.cos(th)]]).astype(rd).astype(float)
    n, xyz, az, za, idx = _rotate_cut(prec, eq, enu, plane, astrom=None if ast is None else ast[0])
    topo64 = ocoords.topocentric_enu(ra, dec, tjd, coords.HERA_LOCATION, method, params)[0]
    topo = topo64.astype(rd)
    # sources within rounding of the horizon may fall on either side: exclude them from the set check
    sure = np.abs(topo64[2]) > 1e-12
    up = np.nonzero(topo[2] > 0)[0]
    got_set = np.zeros(nsrc, bool)
    got_set[idx[:n]] = True
    assert np.array_equal(got_set[sure], (topo[2] > 0)[sure])
    assert np.all(np.diff(idx[:n]) > 0)                  # order-preserving compaction
    if n != up.size or not np.array_equal(idx[:n], up):
        return                                           # a horizon-grazing source: the value checks need equal sets
    if n == 0:
        return
    tp = topo[:, up]
    waz, wza = ocoords.enu_to_az_za(tp[0], tp[1], "uvbeam")
    tol = 2e-6 if prec == 1 else 1e-12
    # az near 0 / 2 pi may wrap differently by one ulp: compare on the circle
    d = np.abs(np.angle(np.exp(1j * (az[:n].astype(float) - waz.astype(float)))))
    assert d.max() < (3e-4 if prec == 1 else 1e-7)       # atan2 near the zenith is ill-conditioned
    np.testing.assert_allclose(za[:n], wza, atol=5e-4 if prec == 1 else 2e-8)
    want = (plane.astype(rd) @ tp).astype(rd) * rd(2 * np.pi)
    np.testing.assert_allclose(xyz[:, :n], want, rtol=0, atol=tol * 10)
    assert np.all(xyz[:, n:] == -77.0)


def test_rotate_cut_slice_and_overflow():
    from fftvis_b200.core import coords
    rng = np.random.default_rng(3)
    nsrc = 5000
    eq = coords.equatorial_unit_vectors(rng.uniform(0, 2 * np.pi, nsrc), np.arcsin(rng.uniform(-1, 1, nsrc)))
    enu = coords.eq_to_enu_matrices(np.array([2459845.0]), coords.HERA_LOCATION)[0]
    topo = enu @ eq
    n, _, _, _, idx = _rotate_cut(2, eq, enu, np.eye(3), lo=1000, hi=3100)
    up = 1000 + np.nonzero(topo[2, 1000:3100] > 0)[0]
    assert n == up.size
    np.testing.assert_array_equal(idx[:n], up)
    total = int((topo[2] > 0).sum())
    n, *_ = _rotate_cut(2, eq, enu, np.eye(3), n_cap=total - 1)
    assert n == -total                                    # overflow is reported, never truncated
    n, *_ = _rotate_cut(2, eq, enu, np.eye(3), lo=10, hi=10)
    assert n == 0


@pytest.mark.parametrize("prec", [1, 2])
def test_inplace_rot(prec):
    from fftvis_b200.gpu import inplace_rot
    rng = np.random.default_rng(0)
    rd = np.float32 if prec == 1 else np.float64
    b = rng.normal(size=(3, 1001)).astype(rd)
    rot = np.array([[0.0, -1.0, 0.0], [1.0, 0.0, 0.0], [0.0, 0.0, 1.0]])   # reference test_core_utils.py:138-170
    want = rot.astype(rd) @ b
    inplace_rot(rot, b)
    np.testing.assert_allclose(b, want, rtol=1e-6 if prec == 1 else 1e-14)


@pytest.mark.parametrize("prec", [1, 2])
def test_four_coherency_products_vs_einsum(prec):
    from fftvis_b200.gpu import GPUBeamEvaluator
    rng = np.random.default_rng(11)
    cd = np.complex64 if prec == 1 else np.complex128
    n = 513
    c = lambda *s: (rng.normal(size=s) + 1j * rng.normal(size=s)).astype(cd)
    bi, bj, flux, coh = c(2, 2, n), c(2, 2, n), c(n), c(2, 2, n)
    ev = GPUBeamEvaluator()
    rtol = 2e-5 if prec == 1 else 1e-12
    b = bi.copy(); ev.get_apparent_flux_polarized_beam(b, flux)
    np.testing.assert_allclose(b, np.einsum("bas,s,bcs->acs", bi.conj(), flux, bi), rtol=rtol, atol=rtol)
    b = bi.copy(); ev.get_apparent_flux_polarized(b, coh)
    np.testing.assert_allclose(b, np.einsum("kin,kmn,mjn->ijn", bi.conj(), coh, bi), rtol=rtol, atol=rtol)
    out = np.zeros_like(bi); ev.get_apparent_flux_polarized_beam_pair(bi, bj, flux, out)
    np.testing.assert_allclose(out, np.einsum("bas,s,bps->aps", bi.conj(), flux, bj), rtol=rtol, atol=rtol)
    out = np.zeros_like(bi); ev.get_apparent_flux_polarized_pair(bi, bj, coh, out)
    np.testing.assert_allclose(out, np.einsum("bas,bks,kps->aps", bi.conj(), coh, bj), rtol=rtol, atol=rtol)


def test_coherency_golden_from_reference(golden):
    from fftvis_b200.gpu import GPUBeamEvaluator
    g, _ = golden
    bi, bj, flux, coh = g["coh/beam_i"], g["coh/beam_j"], g["coh/flux"], g["coh/coherency"]
    ev = GPUBeamEvaluator()
    out = np.zeros_like(bi); ev.get_apparent_flux_polarized_pair(bi, bj, coh, out)
    np.testing.assert_allclose(out, g["coh/out_pair"], rtol=1e-12)
    out = np.zeros_like(bi); ev.get_apparent_flux_polarized_beam_pair(bi, bj, flux, out)
    np.testing.assert_allclose(out, g["coh/out_beam_pair"], rtol=1e-12)


@pytest.mark.parametrize("kind", ["gaussian", "airy", "table0", "table1", "table_endpoint", "table3", "table3_endpoint"])
@pytest.mark.parametrize("polarized", [False, True])
def test_evaluate_beam_vs_oracle(kind, polarized):
    from fftvis_b200 import AiryBeam, GaussianBeam, synth
    from fftvis_b200.gpu import GPUBeamEvaluator
    from oracle import beams as ob
    rng = np.random.default_rng(4)
    n, freq = 2000, 150e6
    az = rng.uniform(0, 2 * np.pi, n)
    za = rng.uniform(0, np.pi / 2, n)
    az[:3] = [0.0, 2 * np.pi - 1e-9, 1e-9]
    za[:3] = [0.0, np.pi / 2, 1e-7]
    opts = {"order": 1}
    if kind == "gaussian":
        beam = GaussianBeam(diameter=14.0)
    elif kind == "airy":
        beam = AiryBeam(diameter=14.0)
    else:
        beam = synth.synthetic_uvbeam([freq], naz=90, nza=46, include_endpoint=kind.endswith("endpoint"))
        if kind == "table0":
            opts = {"order": 0}
        if kind.startswith("table3"):
            opts = {"order": 3}
    model = beam if polarized else (beam.to_power())
    got = GPUBeamEvaluator().evaluate_beam(model, az, za, polarized, freq, spline_opts=opts)
    want = ob.evaluate_beam(model, az, za, polarized, freq, 0, opts)
    assert got.shape == ((2, 2, n) if polarized else (n,))
    if kind == "table0":
        # nearest-neighbour: ties at half-integers may round differently; compare away from them
        zi, ai, _ = ob.table_fractional_index(model, az, za)
        ok = (np.abs(zi - np.floor(zi) - 0.5) > 1e-6) & (np.abs(ai - np.floor(ai) - 0.5) > 1e-6)
        np.testing.assert_allclose(got[..., ok], want[..., ok], rtol=1e-12, atol=1e-14)
    else:
        np.testing.assert_allclose(got, want, rtol=1e-10, atol=1e-13)


def test_evaluate_beam_check_raises_on_nan():
    from fftvis_b200 import synth
    from fftvis_b200.gpu import GPUBeamEvaluator
    beam = synth.synthetic_uvbeam([150e6], naz=36, nza=19)
    beam.data_array[0, 0, 0, 3, 4] = np.nan
    az, za = np.array([beam.axis1_array[4]]), np.array([beam.axis2_array[3]])
    with pytest.raises(ValueError, match="Beam interpolation resulted in an invalid value"):
        GPUBeamEvaluator().evaluate_beam(beam, az, za, True, 150e6, check=True, spline_opts={"order": 1})


def test_evaluator_defaults_and_prepare_beam_evaluation():
    """reference tests/test_gpu_beams.py:7-19 (constructor defaults) and the routing table of
    cpu/beams.py:91-127 (tests/test_cpu_beams.py:708-854)."""
    from fftvis_b200.gpu import GPUBeamEvaluator
    ev = GPUBeamEvaluator()
    assert ev.beam_list == [] and ev.beam_idx is None and ev.polarized is False
    assert ev.nant == 0 and ev.freq == 0.0 and ev.nsrc == 0 and ev.spline_opts == {} and ev.precision == 2
    pairs, to_bls, to_flip = ev.prepare_beam_evaluation([0, 1, 2], [(0, 1), (1, 0), (2, 2), (0, 2)],
                                                        np.array([0, 1, 0]))
    assert [tuple(int(v) for v in p) for p in pairs] == [(0, 0), (0, 1), (1, 1)]
    assert to_bls[(0, 0)] == [2, 3] and to_bls[(0, 1)] == [0, 1] and to_bls[(1, 1)] == []
    assert to_flip[(0, 1)] == [False, True]


@pytest.mark.parametrize("mode", [0, 1, 2])
@pytest.mark.parametrize("prec", [1, 2])
def test_weights_batch_vs_oracle(mode, prec):
    """fv_weights (beam evaluation + apparent coherency, frequency batch, gather through src_idx)
    vs oracle.beams.compute_apparent_coherency."""
    import torch
    from fftvis_b200 import synth
    from fftvis_b200.gpu.beams import DeviceBeam, launch_weights
    from oracle import beams as ob
    rng = np.random.default_rng(21)
    nsrc, n, nf_tot, f0, nb = 3000, 1700, 6, 2, 3
    freqs = np.linspace(100e6, 200e6, nf_tot)
    rd, cd = (np.float32, np.complex64) if prec == 1 else (np.float64, np.complex128)
    rdt, cdt = (torch.float32, torch.complex64) if prec == 1 else (torch.float64, torch.complex128)
    az = rng.uniform(0, 2 * np.pi, n).astype(rd)
    za = rng.uniform(0, np.pi / 2, n).astype(rd)
    src = np.sort(rng.permutation(nsrc)[:n]).astype(np.int32)
    polarized = mode != 0
    if polarized:
        bi = synth.synthetic_uvbeam(freqs, naz=72, nza=37, seed=1, perturb=0.3)
        bj = synth.synthetic_uvbeam(freqs, naz=72, nza=37, seed=2, perturb=0.3)
    else:
        bi = synth.synthetic_uvbeam(freqs, naz=72, nza=37, seed=1, perturb=0.3).to_power()
        bj = synth.synthetic_uvbeam(freqs, naz=72, nza=37, seed=2, perturb=0.3).to_power()
    if mode == 2:
        flux = (rng.normal(size=(nsrc, nf_tot, 2, 2)) + 1j * rng.normal(size=(nsrc, nf_tot, 2, 2))).astype(cd)
        flux_dev = np.ascontiguousarray(np.transpose(flux, (1, 2, 3, 0)).reshape(nf_tot, 4, nsrc))
    else:
        flux = rng.uniform(0.5, 2, size=(nsrc, nf_tot)).astype(cd)
        flux_dev = np.ascontiguousarray(flux.T)
    t = lambda a, dt: torch.as_tensor(np.ascontiguousarray(a)).to("cuda", dt)
    P = 4 if polarized else 1
    n_cap = n + 13
    pad = lambda a: np.concatenate([a, np.zeros(n_cap - n, a.dtype)])
    out = torch.zeros((nb, P, n_cap), dtype=cdt, device="cuda")
    launch_weights(prec, mode, DeviceBeam(bi, prec, 1), DeviceBeam(bj, prec, 1), t(pad(az), rdt), t(pad(za), rdt),
                   t(pad(src), torch.int32), torch.tensor([n], dtype=torch.int32, device="cuda"), n_cap,
                   t(freqs, torch.float64), f0, nb, t(flux_dev, cdt), nsrc, out)
    got = out.cpu().numpy()
    for b in range(nb):
        fi = f0 + b
        bev = [ob.evaluate_beam(m, az, za, polarized, freqs[fi], fi, {"order": 1}).astype(cd) for m in (bi, bj)]
        want = ob.compute_apparent_coherency(bev, 0, 1, flux[src, fi], polarized, mode == 2)
        assert relerr(got[b, :, :n], want) < (3e-6 if prec == 1 else 1e-13)
        assert np.all(got[b, :, n:] == 0)


@pytest.mark.parametrize("prec", [1, 2])
def test_basis_contract_vs_numpy(prec):
    import torch
    from fftvis_b200.gpu import _lib
    rng = np.random.default_rng(31)
    nb, nk, nant, K, nf_tot, f0 = 3, 77, 9, 4, 5, 1
    cd = np.complex64 if prec == 1 else np.complex128
    cdt = torch.complex64 if prec == 1 else torch.complex128
    c = lambda *s: (rng.normal(size=s) + 1j * rng.normal(size=s)).astype(cd)
    vkl, coefs = c(nb, 4, nk), c(nant, K, nf_tot)
    a1 = rng.integers(0, nant, nk).astype(np.int32)
    a2 = rng.integers(0, nant, nk).astype(np.int32)
    t = lambda a, dt: torch.as_tensor(np.ascontiguousarray(a)).to("cuda", dt)
    vkl_d, coefs_d, a1_d, a2_d = t(vkl, cdt), t(coefs, cdt), t(a1, torch.int32), t(a2, torch.int32)
    for kk, ll in [(1, 1), (0, 3)]:
        out = torch.zeros((nb, 4, nk), dtype=cdt, device="cuda")
        epi = _lib.make_epilogue(out.data_ptr(), out.stride(0), out.stride(1), accumulate=True)
        _lib.check(_lib.lib().fv_basis_contract(prec, vkl_d.data_ptr(), nb, nk, coefs_d.data_ptr(),
                                                nant, K, nf_tot, f0, kk, ll, a1_d.data_ptr(),
                                                a2_d.data_ptr(), ctypes.byref(epi),
                                                torch.cuda.current_stream().cuda_stream))
        got = out.cpu().numpy().reshape(nb, 2, 2, nk)
        v = vkl.reshape(nb, 2, 2, nk)
        want = np.zeros_like(v)
        for b in range(nb):
            f = f0 + b
            w1 = coefs[a1, kk, f].conj() * coefs[a2, ll, f]
            want[b] = w1 * v[b]
            if kk != ll:
                want[b] += (coefs[a1, ll, f].conj() * coefs[a2, kk, f]) * np.swapaxes(v[b], 0, 1)
        assert relerr(got, want) < (1e-5 if prec == 1 else 1e-13)


def test_library_fails_loudly_on_bad_arguments():
    from fftvis_b200.gpu import _lib
    with pytest.raises(_lib.FVError, match="prec must be 1 or 2"):
        _lib.check(_lib.lib().fv_inplace_rot(3, _lib.doubles(np.eye(3).ravel()), None, 0, None))
