"""compute_beam_basis (SVD eigenbeams on the GPU) -- mirrors the reference's tests/test_beam_basis.py
(:75-260 shapes / rank / reconstruction, :262-303 argument errors, :310-431 basis simulation ==
per-antenna simulation)."""
import numpy as np
import pytest

from gpu_helpers import relerr

_FREQ = 150e6


def _beams():
    from fftvis_b200 import AiryBeam
    return AiryBeam(diameter=14.0), AiryBeam(diameter=7.0)


def _flat(eigenbeams):
    return np.stack([eb.data_array[:, :, 0].ravel() for eb in eigenbeams], axis=0)


# ---- argument errors: raised before any device work (reference :262-303) -----------------------
def test_argument_errors():
    from fftvis_b200 import compute_beam_basis
    a, _ = _beams()
    with pytest.raises(ValueError, match="at least one beam"):
        compute_beam_basis([], freq=_FREQ, polarized=True)
    with pytest.raises(ValueError, match="threshold"):
        compute_beam_basis([a], freq=_FREQ, polarized=True, threshold=0.0)
    with pytest.raises(ValueError, match="threshold"):
        compute_beam_basis([a], freq=_FREQ, polarized=True, threshold=1.5)
    with pytest.raises(ValueError, match="scalar freq"):
        compute_beam_basis([a], freq=np.array([100e6, 150e6]), polarized=True)
    with pytest.raises(ValueError, match="supplied together"):
        compute_beam_basis([a], freq=_FREQ, polarized=True, axis1_array=np.linspace(0, 2 * np.pi, 36))
    with pytest.raises(ValueError, match="supplied together"):
        compute_beam_basis([a], freq=_FREQ, polarized=True, axis2_array=np.linspace(0, np.pi / 2, 19))
    with pytest.raises(ValueError, match="requires efield beams"):
        compute_beam_basis([a.to_power()], freq=_FREQ, polarized=True)


# ---- shapes and rank (reference :75-170) ---------------------------------------------------------
@pytest.mark.gpu
def test_shapes_and_custom_axes():
    from fftvis_b200 import compute_beam_basis
    a, b = _beams()
    eb, coefs = compute_beam_basis([a], freq=_FREQ, polarized=True)
    assert len(eb) == 1 and coefs.shape == (1, 1)
    assert eb[0].data_array.shape == (2, 2, 1, 181, 361) and eb[0].beam_type == "efield"
    eb, coefs = compute_beam_basis([a, b], freq=_FREQ, polarized=True)
    assert coefs.shape[0] == 2 and coefs.shape[1] == len(eb)
    axis1, axis2 = np.linspace(0, 2 * np.pi, 36), np.linspace(0, np.pi / 2, 19)
    eb, _ = compute_beam_basis([a], freq=_FREQ, polarized=True, axis1_array=axis1, axis2_array=axis2)
    assert eb[0].axis1_array.shape == axis1.shape and eb[0].axis2_array.shape == axis2.shape
    assert eb[0].data_array.shape == (2, 2, 1, 19, 36)
    eb, coefs = compute_beam_basis([a, b], freq=_FREQ, polarized=False, n_axis1=73, n_axis2=37)
    assert eb[0].data_array.shape == (1, 1, 1, 37, 73) and eb[0].beam_type == "power" and coefs.shape[0] == 2


@pytest.mark.gpu
def test_rank_tracks_beam_diversity():
    from fftvis_b200 import compute_beam_basis
    a, b = _beams()
    eb, coefs = compute_beam_basis([a, a, a], freq=_FREQ, polarized=True, threshold=1e-3)
    assert len(eb) == 1 and coefs.shape == (3, 1)
    eb, coefs = compute_beam_basis([a, b], freq=_FREQ, polarized=True, threshold=1e-12)
    assert len(eb) == 2
    k_tight = len(compute_beam_basis([a, b], freq=_FREQ, polarized=True, threshold=1e-12)[0])
    k_loose = len(compute_beam_basis([a, b], freq=_FREQ, polarized=True, threshold=0.99)[0])
    assert k_loose <= k_tight


# ---- reconstruction (reference :172-260) -----------------------------------------------------------
@pytest.mark.gpu
def test_reconstruction_and_orthonormality():
    from fftvis_b200 import compute_beam_basis
    from fftvis_b200.gpu import GPUBeamEvaluator
    a, b = _beams()
    n1, n2 = 73, 37
    eb, coefs = compute_beam_basis([a, b, a], freq=_FREQ, polarized=True, n_axis1=n1, n_axis2=n2)
    V = _flat(eb)
    assert np.allclose(V @ V.conj().T, np.eye(len(eb)), atol=1e-12)      # right-singular vectors
    recon = coefs @ V
    assert np.allclose(recon[0], recon[2], atol=1e-13)                   # identical beams, identical rows
    # the rows reproduce the beams evaluated on the same grid
    az = np.tile(np.linspace(0.0, 2 * np.pi, n1), n2)
    za = np.repeat(np.linspace(0.0, np.pi, n2), n1)
    ev = GPUBeamEvaluator()
    for row, beam in zip(recon, (a, b, a)):
        want = ev.evaluate_beam(beam, az, za, True, _FREQ).ravel()
        assert relerr(row, want) < 1e-12
    # truncation loses norm (reference :230-260)
    eb1, c1 = compute_beam_basis([a, b], freq=_FREQ, polarized=True, threshold=0.99, n_axis1=n1, n_axis2=n2)
    eb2, c2 = compute_beam_basis([a, b], freq=_FREQ, polarized=True, threshold=1e-12, n_axis1=n1, n_axis2=n2)
    assert np.linalg.norm(c1 @ _flat(eb1)) <= np.linalg.norm(c2 @ _flat(eb2)) + 1e-12


# ---- basis simulation == per-antenna simulation (reference :310-431) -------------------------------
def _sim_params():
    from fftvis_b200 import HERA_LOCATION
    ants = {0: np.array([0.0, 0.0, 0.0]), 1: np.array([14.6, 0.0, 0.0]), 2: np.array([0.0, 14.6, 0.0])}
    rng = np.random.default_rng(42)
    nsrc = 30
    return dict(ants=ants, freqs=np.array([_FREQ]), ra=rng.uniform(0, 2 * np.pi, nsrc),
                dec=rng.uniform(-np.pi / 4, np.pi / 4, nsrc), fluxes=rng.uniform(0.5, 2.0, (nsrc, 1)),
                times=np.array([2458119.5]), telescope_loc=HERA_LOCATION,
                baselines=[(0, 0), (0, 1), (0, 2), (1, 2)])


@pytest.mark.gpu
@pytest.mark.parametrize("different", [False, True])
def test_basis_simulation_matches_per_antenna_simulation(different):
    from fftvis_b200 import UVBeamTable, compute_beam_basis, simulate_vis
    a, b = _beams()
    p = _sim_params()
    nant = len(p["ants"])
    if different:
        beam_list, beam_idx = [a, b], np.array([i % 2 for i in range(nant)])
        eb, coefs = compute_beam_basis(beam_list, freq=_FREQ, polarized=True, threshold=1e-12)
        coefs_per_ant = coefs[beam_idx, :, np.newaxis]
    else:
        beam_list, beam_idx = [a] * nant, np.arange(nant)
        eb, coefs = compute_beam_basis([a], freq=_FREQ, polarized=True)
        coefs_per_ant = np.tile(coefs[np.newaxis], (nant, 1, 1))
    kw = dict(polarized=True, eps=1e-10, beam_spline_opts={"order": 3})
    vis_basis = simulate_vis(beam=eb, beam_coefs=coefs_per_ant, **kw, **p)
    assert vis_basis.shape == (1, 1, 2, 2, 4) and np.isfinite(vis_basis).all()
    # (1) against the same beams as az/za tables rebuilt from the basis: interpolation is linear in
    #     the table, so the two paths agree to NUFFT accuracy
    V = np.stack([e.data_array for e in eb], axis=0)
    uniq = coefs if different else coefs[:1]
    tables = [UVBeamTable(np.tensordot(c, V, axes=(0, 0)), eb[0].axis1_array, eb[0].axis2_array,
                          eb[0].freq_array, "efield") for c in uniq]
    idx = beam_idx if different else np.zeros(nant, dtype=int)
    vis_tab = simulate_vis(beam=tables, beam_idx=idx, **kw, **p)
    assert relerr(vis_basis, vis_tab) < 1e-8
    # (2) against the analytic beams themselves (the reference's comparison, atol 1e-5 there): limited
    #     by the cubic-spline interpolation of the 1-degree eigenbeam grid
    vis_ref = simulate_vis(beam=beam_list, beam_idx=beam_idx, **kw, **p)
    assert vis_ref.shape == vis_basis.shape
    err = relerr(vis_basis, vis_ref)
    print("basis vs analytic relerr", err)
    assert err < 1e-4


# ---- against numpy.linalg.svd of ORACLE-evaluated beams (reference core/beam_basis.py:128-151) -----
@pytest.mark.gpu
@pytest.mark.parametrize("polarized", [True, False])
def test_singular_values_and_subspace_match_numpy_svd_of_oracle_beams(polarized):
    """The device QR + SVD against the reference's own recipe done on the CPU: beams evaluated by the
    oracle's numpy beam models on the common grid, flattened, ``numpy.linalg.svd(full_matrices=False)``,
    ``s / s[0] >= threshold`` (core/beam_basis.py:128-151).  Compared: the number of retained
    eigenbeams, the singular values (= column norms of ``beam_coefs``), the spanned subspace (the
    singular vectors themselves are defined up to a phase) and the reconstructed beams."""
    from fftvis_b200 import AiryBeam, GaussianBeam, compute_beam_basis, synth
    from fftvis_b200.beam_models import as_beam_model
    from oracle import beams as obeams
    n1, n2 = 91, 46
    axis1, axis2 = np.linspace(0.0, 2 * np.pi, n1), np.linspace(0.0, np.pi, n2)
    analytic = [AiryBeam(diameter=14.0), AiryBeam(diameter=13.0), GaussianBeam(diameter=14.0),
                GaussianBeam(diameter=12.0), AiryBeam(diameter=14.0)]            # rank 4 of 5
    tables = [synth.synthetic_uvbeam([_FREQ], naz=90, nza=46, seed=s, perturb=0.3) for s in range(3)]
    beam_list = analytic + (tables if polarized else [])
    eb, coefs = compute_beam_basis(beam_list, freq=_FREQ, polarized=polarized, threshold=1e-10,
                                   axis1_array=axis1, axis2_array=axis2)
    az, za = np.tile(axis1, n2), np.repeat(axis2, n1)
    rows = []
    for b in beam_list:
        m = as_beam_model(b)
        if not polarized:
            m = m.to_power()
        r = obeams.evaluate_beam(m, az, za, polarized, _FREQ, 0, {"order": 1})
        rows.append(np.asarray(r).reshape(-1))
    flat = np.stack(rows)
    u, s, vh = np.linalg.svd(flat, full_matrices=False)
    K = int((s / s[0] >= 1e-10).sum())
    assert len(eb) == K and coefs.shape == (len(beam_list), K)
    assert K == (7 if polarized else 4)
    np.testing.assert_allclose(np.linalg.norm(coefs, axis=0), s[:K], rtol=1e-10)
    V = _flat(eb)
    Vn = vh[:K]
    # same subspace: projecting numpy's right-singular vectors onto ours loses nothing
    assert np.linalg.norm(Vn @ V.conj().T @ V - Vn) < 1e-9
    # well-separated singular values: vectors agree up to a unit phase
    for k in range(K):
        if min(abs(s[k] - s[j]) for j in range(len(s)) if j != k) > 1e-6 * s[0]:
            assert abs(abs(np.vdot(Vn[k], V[k])) - 1.0) < 1e-9
    assert relerr(coefs @ V, flat) < 1e-10


@pytest.mark.gpu
def test_basis_simulation_matches_oracle_per_antenna_cpu_pipeline():
    """compute_beam_basis -> simulate_vis(beam_coefs=...) on the GPU against the ORACLE's CPU pipeline run
    with the original per-antenna beams (reference tests/test_beam_basis.py:310-431 compares the two
    CPU paths at atol 1e-5; here the basis grid is the table beams' own grid, so the decomposition is
    exact and the bar is the NUFFT's)."""
    from fftvis_b200 import compute_beam_basis, simulate_vis, synth
    from oracle import pipeline
    p = _sim_params()
    nant = len(p["ants"])
    beams = [synth.synthetic_uvbeam(p["freqs"], naz=72, nza=37, seed=s, perturb=0.3) for s in range(3)]
    # real-valued beams: the reference's lower-triangle shortcut (swapaxes without conjugation, SURVEY
    # App. D.5) is exact only for a real basis
    for b in beams:
        b.data_array = b.data_array.real.astype(complex)
    eb, coefs = compute_beam_basis(beams, freq=_FREQ, polarized=True, threshold=1e-12)
    assert len(eb) == 3
    beam_idx = np.arange(nant) % 3
    kw = dict(polarized=True, eps=1e-12, beam_spline_opts={"order": 1})
    vis_basis = simulate_vis(beam=eb, beam_coefs=coefs[beam_idx, :, np.newaxis], **kw, **p)
    q = {k: v for k, v in p.items() if k != "fluxes"}
    cpu = pipeline.simulate_cpu(fluxes=p["fluxes"], beam_list=beams, beam_idx=beam_idx, **kw, **q)
    direct = pipeline.simulate_direct(fluxes=p["fluxes"], beam_list=beams, beam_idx=beam_idx, polarized=True,
                                      beam_spline_opts={"order": 1}, **q)
    assert relerr(cpu, direct) < 1e-11
    assert relerr(vis_basis, cpu) < 1e-10
    assert relerr(vis_basis, direct) < 1e-10
