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
