"""Host planners vs golden vectors produced by the reference's own modules
(tests/golden/make_golden.py; reference files core/utils.py, core/antenna_gridding.py,
cpu/utils.py, cpu/beams.py:129-246)."""
import numpy as np
import pytest

from fftvis_b200.core import antenna_gridding as ag
from fftvis_b200.core import catalog, utils

ARRAYS = ["hex2", "hex3", "hex4", "line", "square", "random", "tilted", "holey_hex4", "sheared_square"]


def _ants(g, name):
    return {int(k): v for k, v in zip(g[f"{name}/antkeys"], g[f"{name}/antpos"])}


@pytest.mark.parametrize("name", ARRAYS)
def test_get_pos_reds_matches_reference(golden, name):
    g, _ = golden
    ants = _ants(g, name)
    reds = utils.get_pos_reds(ants, include_autos=True)
    assert [tuple(r[0]) for r in reds] == [tuple(x) for x in g[f"{name}/red_first"]]
    assert [len(r) for r in reds] == list(g[f"{name}/red_sizes"])
    assert [tuple(b) for r in reds for b in r] == [tuple(x) for x in g[f"{name}/red_flat"]]
    reds_na = utils.get_pos_reds(ants, include_autos=False)
    assert [tuple(r[0]) for r in reds_na] == [tuple(x) for x in g[f"{name}/red_first_noautos"]]


@pytest.mark.parametrize("name", ARRAYS)
def test_plane_rotation_matches_reference(golden, name):
    g, _ = golden
    vec = g[f"{name}/antpos"]
    np.testing.assert_allclose(utils.get_plane_to_xy_rotation_matrix(vec), g[f"{name}/plane_rot"],
                               rtol=0, atol=1e-14)


@pytest.mark.parametrize("name", ARRAYS)
def test_griddability_matches_reference(golden, name):
    g, meta = golden
    ants = _ants(g, name)
    ok, gridded, basis = ag.check_antpos_griddability(ants)
    assert ok == meta[f"{name}/griddable"]
    np.testing.assert_array_equal(np.asarray(basis, float), g[f"{name}/basis"])
    if ok:
        np.testing.assert_array_equal(np.array([gridded[k] for k in ants]), g[f"{name}/gridded"])


def test_reference_gridding_cases():
    # reference tests/test_antenna_gridding.py:59-82
    assert ag.check_antpos_griddability({i: np.array([i * 3.0, 0, 0]) for i in range(4)})[0]
    sq = {i * 5 + j: np.array([i * 2.0, j * 2.0, 0]) for i in range(5) for j in range(5)}
    assert ag.check_antpos_griddability(sq)[0]
    rng = np.random.default_rng(42)
    rnd = {i: np.append(rng.uniform(0, 100, 2), 0) for i in range(10)}
    assert not ag.check_antpos_griddability(rnd)[0]
    assert not ag.check_antpos_griddability({0: np.zeros(3)})[0]


def test_task_chunks_match_reference(golden):
    _, meta = golden
    for key, want in meta["task_chunks"].items():
        args = eval(key)
        npz, fc, tc, nf, nt = utils.get_task_chunks(*args)
        assert (npz, nf, nt) == (want["nproc"], want["nf"], want["nt"])
        assert [[s.start, s.stop] for s in fc] == want["fc"]
        assert [[s.start, s.stop] for s in tc] == want["tc"]
    # reference tests/test_core_utils.py:26-45
    assert utils.get_task_chunks(3, 30, 1)[3] == 10
    assert utils.get_task_chunks(10, 5, 1)[0] == 1


def test_source_catalog_matches_reference(golden):
    g, _ = golden
    c, pol = catalog.prepare_source_catalog(g["catalog/sky_i"], False)
    assert not pol
    np.testing.assert_array_equal(c, g["catalog/coh_i"])
    c, pol = catalog.prepare_source_catalog(g["catalog/sky_iquv"], True)
    assert pol
    np.testing.assert_allclose(c, g["catalog/coh_iquv"], rtol=1e-15)
    with pytest.raises(ValueError, match="polarized_beam=False requires sky_model to be 2D"):
        catalog.prepare_source_catalog(g["catalog/sky_iquv"], False)
    with pytest.raises(ValueError, match="polarized_beam=True requires sky_model"):
        catalog.prepare_source_catalog(np.zeros((3, 2, 5)), True)


def test_inplace_rot_base(golden):
    g, _ = golden
    b = g["rot/b_in"].copy()
    utils.inplace_rot_base(g["rot/rot"], b)
    np.testing.assert_allclose(b, g["rot/b_out"], rtol=1e-14, atol=1e-15)


def test_validate_beam_idx_messages():
    # strings matched by reference tests/test_wrapper.py:123-141, tests/test_beam_basis.py:459,476
    with pytest.raises(ValueError, match="beam_idx must be provided"):
        utils.validate_beam_idx(None, None, 2, 3)
    with pytest.raises(ValueError, match="beam_idx must be length nant"):
        utils.validate_beam_idx(np.zeros(2, int), None, 1, 3)
    with pytest.raises(ValueError, match="beam_idx contains indices greater"):
        utils.validate_beam_idx(np.array([0, 1, 5]), None, 2, 3)
    with pytest.raises(ValueError, match="beam_idx should not be provided when beam_coefs"):
        utils.validate_beam_idx(np.zeros(3, int), np.zeros((3, 2, 1)), 2, 3)
    np.testing.assert_array_equal(utils.validate_beam_idx(None, None, 3, 3), np.arange(3))
    assert utils.validate_beam_idx(None, None, 1, 3) is None
