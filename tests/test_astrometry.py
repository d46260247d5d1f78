"""The coordinate model (stage a1, host part): product (fftvis_b200/core/astrometry.py, coords.py)
against the oracle's separately written chain (oracle/coords.py), both against SOFA/ERFA's published
test values (t_erfa_c.c: era00, pfw06, nut06a, pnm06a, s06, epv00), and -- wherever ``erfa`` is
installed -- against erfa itself.  Reference call sites: cpu_simulate.py:693-709, 937-940."""
import json
from pathlib import Path

import numpy as np
import pytest

from fftvis_b200.core import astrometry as A
from fftvis_b200.core import coords
from oracle import coords as oc

MAS = A.AS2R * 1e-3
MJD0 = 2400000.5


def _t(mjd):
    return ((MJD0 - A.DJ00) + mjd) / A.DJC


def test_series_tables_shared_with_the_oracle_are_identical():
    d = json.loads((Path(oc.__file__).parent / "data" / "iau_series.json").read_text())
    np.testing.assert_array_equal(np.array(d["nut"]), A._NUT)
    np.testing.assert_array_equal(np.array(d["s06_t0"]), A._S06_T0)
    np.testing.assert_array_equal(np.array(d["s06_t1"]), A._S06_T1)
    np.testing.assert_array_equal(np.array(d["s06_t2"]), A._S06_T2)
    assert [tuple(x) for x in d["leap"]] == [tuple(x) for x in A._LEAP]
    assert {k: tuple(v) for k, v in d["elements"].items()} == A._ELEMENTS


def test_published_erfa_values_exact_parts():
    # t_erfa_c.c: era00(2400000.5, 54388.0), pfw06(2400000.5, 50123.9999), obl06 via pfw06's epsa
    for f in (A.earth_rotation_angle, lambda j: np.array([oc.era00(float(j))])):
        assert abs(f(MJD0 + 54388.0)[0] - 0.4022837240028158102) < 1e-12
    want = (-0.2243387670997995690e-5, 0.4091014602391312808, -0.9501954178013031895e-3, 0.4091014316587367491)
    for f in (A.fukushima_williams, oc.pfw06):
        np.testing.assert_allclose(f(_t(50123.9999)), want, rtol=0, atol=1e-16)
    assert A.tai_minus_utc(2459845.0) == 37.0 and oc.dat(2453736.5) == 33.0 and oc.dat(2453736.4) == 32.0


def test_published_erfa_values_truncated_parts():
    """Truncation budget: nutation < 2 mas, NPB matrix < 2 mas, CIO locator < 1e-11 rad, Earth velocity
    2e-5 relative (0.4 mas of aberration)."""
    t = _t(53736.0)
    for f in (A.nutation, oc.nut06a_truncated):
        dpsi, deps = f(t)
        assert abs(dpsi + 0.9630912025820308797e-5) < 2.0 * MAS
        assert abs(deps - 0.4063238496887249798e-4) < 0.2 * MAS
    ref = np.array([[0.9999995832794205484, 0.8372382772630962111e-3, 0.3639684771140623099e-3],
                    [-0.8372533744743683605e-3, 0.9999996486492861646, 0.4132905944611019498e-4],
                    [-0.3639337469629464969e-3, -0.4163377605910663999e-4, 0.9999999329094260057]])
    for f in (A.npb_matrix, lambda tt: np.array(oc.pnm06a(tt))):
        assert np.abs(f(_t(50123.9999)) - ref).max() < 2.0 * MAS
    x, y = 0.5791308486706011000e-3, 0.4020579816732961219e-4
    for f in (A.cio_locator, oc.s06):
        assert abs(f(t, x, y) + 0.1220032213076463117e-7) < 1e-11
    eh = np.array([-0.7757238809297706813, 0.5598052241363340596, 0.2426998466481686993])
    vb = np.array([-0.1091874268116823295e-1, -0.1246525461732861538e-1, -0.5404773180966231279e-2])
    h, _, v = A.earth_posvel(_t(53411.52501161))
    assert np.linalg.norm(h - eh) < 1e-4 and np.linalg.norm(v - vb) / np.linalg.norm(vb) < 5e-5
    h2, v2 = oc.epv_approx(_t(53411.52501161))
    assert np.linalg.norm(np.array(h2) - eh) < 1e-4 and np.linalg.norm(np.array(v2) - vb) / np.linalg.norm(vb) < 5e-5


def test_published_erfa_ab_and_ld_values():
    # t_erfa_c.c t_ab / t_ldsun
    pnat = np.array([[-0.76321968546737951], [-0.60869453983060384], [-0.21676408580639883]])
    v = [2.1044018893653786e-5, -8.9108923304429319e-5, -3.8633714797716569e-5]
    got = oc.ab(pnat, v, 0.99980921395708788, 0.99999999506209258)[:, 0]
    np.testing.assert_allclose(got, [-0.7631631094219556269, -0.6087553082505590832, -0.2167926269368471279], atol=1e-12)
    row = np.array([0, 0, 1, 1e9, *v, 0.99999999506209258, 1e-6, 1.0])        # Sun far away: deflection ~ 0
    got2 = A.apply_astrom(pnat, row)[:, 0]
    np.testing.assert_allclose(got2, got, atol=1e-11)
    p = np.array([[-0.763276255], [-0.608633767], [-0.216735543]])
    e = [-0.973644023, -0.20925523, -0.0907169552]
    got = oc.ldsun(p, e, 0.999809214)[:, 0]
    np.testing.assert_allclose(got, [-0.7632762580731413169, -0.6086337635262647900, -0.2167355419322321302], atol=1e-12)


@pytest.mark.parametrize("method,params", [
    ("CoordinateRotationERFA", {}), ("CoordinateRotationAstropy", {"dut1": 0.05, "xp": 1e-6, "yp": 1.5e-6}),
    ("CoordinateRotationERA", {}), ("CoordinateRotationERFA", {"update_bcrs_every": 1000.0})])
def test_product_blocks_equal_oracle_chain(method, params):
    rng = np.random.default_rng(1)
    ra, dec = rng.uniform(0, 2 * np.pi, 3000), np.arcsin(rng.uniform(-1, 1, 3000))
    times = 2459845.0 + np.arange(5) * 0.01
    mats, ast = coords.coordinate_blocks(times, coords.HERA_LOCATION, method, params)
    eq = coords.equatorial_unit_vectors(ra, dec)
    got = np.stack([mats[i] @ (A.apply_astrom(eq, ast[i]) if ast is not None else eq) for i in range(times.size)])
    want = oc.topocentric_enu(ra, dec, times, coords.HERA_LOCATION, method, params)
    assert np.abs(got - want).max() < 1e-13


def test_model_magnitudes():
    """Sanity of the pieces: precession-nutation moves the pole ~7.6 arcmin by 2022, annual aberration is
    <= 20.5 arcsec, deflection is sub-mas away from the Sun, update_bcrs_every = inf changes < 1 arcsec over 1 h."""
    times = np.array([2459845.0])
    full, ast = coords.coordinate_blocks(times, coords.HERA_LOCATION)
    era, none = coords.coordinate_blocks(times, coords.HERA_LOCATION, "CoordinateRotationERA")
    assert none is None
    ang = np.degrees(np.arccos((np.trace(full[0] @ era[0].T) - 1) / 2)) * 60
    assert 7.0 < ang < 8.2
    rng = np.random.default_rng(0)
    eq = coords.equatorial_unit_vectors(rng.uniform(0, 2 * np.pi, 5000), np.arcsin(rng.uniform(-1, 1, 5000)))
    moved = A.apply_astrom(eq, ast[0])
    sep = np.degrees(np.arccos(np.clip(np.sum(eq * moved, 0), -1, 1))) * 3600
    assert 15.0 < sep.max() < 20.6
    with pytest.raises(KeyError):
        coords.coordinate_blocks(times, coords.HERA_LOCATION, "NoSuchMethod")
    with pytest.raises(TypeError):
        coords.coordinate_blocks(times, coords.HERA_LOCATION, coord_method_params={"bogus": 1})
    t2 = 2459845.0 + np.array([0.0, 1.0 / 24])
    a = oc.topocentric_enu([1.0], [-0.5], t2, coords.HERA_LOCATION, "CoordinateRotationERFA", {"update_bcrs_every": 1e9})
    b = oc.topocentric_enu([1.0], [-0.5], t2, coords.HERA_LOCATION, "CoordinateRotationERFA", {})
    assert np.abs(a - b).max() < 5e-6 and np.abs(a[0] - b[0]).max() == 0.0


def test_against_erfa_when_installed():
    """Runs wherever ``erfa`` exists (not in this image): ICRS -> observed az/zd through erfa.atco13
    with zero refraction against the oracle chain, to the model's truncation budget (5 mas)."""
    erfa = pytest.importorskip("erfa")
    rng = np.random.default_rng(2)
    ra, dec = rng.uniform(0, 2 * np.pi, 200), np.arcsin(rng.uniform(-0.95, 0.95, 200))
    jd = 2459845.25
    lat, lon, h = oc.site(coords.HERA_LOCATION)
    aob, zob, *_ = erfa.atco13(ra, dec, 0, 0, 0, 0, jd, 0.0, 0.0, lon, lat, h, 0, 0, 0, 0, 0, 0.55)
    want = np.stack([np.sin(aob) * np.sin(zob), np.cos(aob) * np.sin(zob), np.cos(zob)])
    got = oc.topocentric_enu(ra, dec, np.array([jd]), coords.HERA_LOCATION)[0]
    assert np.abs(got - want).max() < 5 * MAS
