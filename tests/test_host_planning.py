"""Host-side planning that needs no GPU: value-keyed memoisation of the planners, the coordinate-manager
carrier of the reference's chunk-evaluator call form, and the frequency-batch choice of the fused type-1 path."""
import types

import numpy as np
import pytest


def test_memo_by_value_hits_on_equal_values_and_returns_independent_copies():
    from fftvis_b200.core.utils import memo_by_value
    calls = []

    @memo_by_value(maxsize=2)
    def planner(antpos, tol=1e-9, flag=False):
        calls.append(1)
        return {"basis": np.array([[1.0, 2.0], [3.0, 4.0]]) * len(antpos)}, [(0, 1), (1, 2)]

    a = {0: np.zeros(3), 1: np.array([14.6, 0.0, 0.0])}
    r1 = planner(a, tol=1e-9)
    r2 = planner({0: np.zeros(3), 1: np.array([14.6, 0.0, 0.0])}, tol=1e-9)      # equal values, new objects
    assert len(calls) == 1 and np.array_equal(r1[0]["basis"], r2[0]["basis"])
    r1[0]["basis"][0, 0] = -1.0                                                   # callers may modify what they get
    assert planner(a, tol=1e-9)[0]["basis"][0, 0] == 2.0 and len(calls) == 1
    planner({0: np.zeros(3), 1: np.array([14.6, 0.0, 1e-12])})                    # any change of a value recomputes
    planner(a, tol=1e-8)
    assert len(calls) == 3
    planner(a, tol=1e-9)                                                          # evicted (maxsize 2): recomputed
    assert len(calls) == 4
    planner(a, tol=lambda: None)                                                  # unhashable argument: no caching
    planner(a, tol=lambda: None)
    assert len(calls) == 6


def test_memoised_planners_match_uncached_results():
    from fftvis_b200 import synth
    from fftvis_b200.core import antenna_gridding, utils
    ants = synth.hex_array(4)
    first = utils.get_pos_reds(ants, include_autos=True)
    again = utils.get_pos_reds({k: v.copy() for k, v in ants.items()}, include_autos=True)
    assert first == again and first is not again and first[0] is not again[0]
    g1 = antenna_gridding.check_antpos_griddability(ants)
    g2 = antenna_gridding.check_antpos_griddability(ants)
    assert g1[0] == g2[0] and np.array_equal(g1[2], g2[2]) and g1[2] is not g2[2]
    moved = {k: v + (0.37 if k == 5 else 0.0) for k, v in ants.items()}           # one antenna off the lattice
    assert antenna_gridding.check_antpos_griddability(moved)[0] is False


def test_coordinate_rotation_carrier_and_manager_inputs():
    """What _evaluate_vis_chunk reads from a coord_mgr (cpu_simulate.py:693-704): ours, and a matvis-like object
    with astropy-like attributes (duck-typed)."""
    from fftvis_b200 import HERA_LOCATION
    from fftvis_b200.core.coords import CoordinateRotation, manager_inputs
    ra, dec = np.array([0.1, 0.2, 0.3]), np.array([-0.5, -0.4, -0.6])
    flux = np.ones((3, 4))
    m = CoordinateRotation(flux=flux, times=np.array([2459845.0]), telescope_loc=HERA_LOCATION, skycoords=(ra, dec),
                           precision=2, source_buffer=0.75, chunk_size=2, update_bcrs_every=30.0)
    got = manager_inputs(m)
    assert got["method"] == "CoordinateRotationERFA" and got["chunk_size"] == 2 and got["source_buffer"] == 0.75
    assert got["params"] == {"update_bcrs_every": 30.0} and np.array_equal(got["ra"], ra)
    assert CoordinateRotation._methods["CoordinateRotationAstropy"] == "CoordinateRotationAstropy"
    with pytest.raises(KeyError):
        CoordinateRotation(flux=flux, times=[0.0], telescope_loc=HERA_LOCATION, skycoords=(ra, dec), method="nope")
    angle = lambda v: types.SimpleNamespace(rad=np.asarray(v))
    second = types.SimpleNamespace(to_value=lambda unit: 120.0)
    CoordinateRotationAstropy = type("CoordinateRotationAstropy", (), {})    # matvis' class name selects the method
    matvis_like = CoordinateRotationAstropy()
    matvis_like.__dict__.update(
        flux=flux, times=types.SimpleNamespace(jd=np.array([2459845.0])), telescope_loc=HERA_LOCATION,
        skycoords=types.SimpleNamespace(ra=angle(ra), dec=angle(dec)), chunk_size=3, source_buffer=1.0,
        update_bcrs_every=second)
    got = manager_inputs(matvis_like)
    assert got["method"] == "CoordinateRotationAstropy" and got["params"]["update_bcrs_every"] == 120.0
    assert np.array_equal(got["dec"], dec)


def test_auto_batch_fills_whole_waves_of_both_passes():
    """cfg2's geometry (n_modes 465 -> nf 960, 241 needed columns, single precision): the batch fills whole waves
    of the x-direct pass 1 (40 strips, three CTAs per SM) and of pass 2 (31 column blocks) and keeps T <= 96 MB."""
    from fftvis_b200.gpu.gpu_simulate import GPUSimulationEngine
    eng = GPUSimulationEngine.__new__(GPUSimulationEngine)
    eng.type1_method, eng.grid_budget_bytes = "fused", 48 << 20
    nb = eng._auto_batch(True, 465, 1, 1, 6e-8, 2.0, 10000, ncols=241)
    assert 27 <= nb <= 54
    slots = 3 * 148
    for ctas in (40 * nb, 31 * nb):
        waves = -(-ctas // slots)
        assert ctas / (waves * slots) > 0.9
    assert nb * 241 * 960 * 8 <= 100 << 20
    # double precision, small grid held whole in one CTA: one strip per transform
    assert eng._auto_batch(True, 41, 4, 2, 1e-13, 2.0, 100000, ncols=41) >= 1
    # type 3: small batches (the grids are large)
    assert eng._auto_batch(False, None, 1, 2, 1e-13, 2.0, 3000000) <= 4
