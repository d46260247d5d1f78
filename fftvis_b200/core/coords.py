"""Host-side coordinate manager of stage a1: everything that is *per time*,
not per source.  The per-source work (light deflection, aberration, 3x3 rotation of Ns unit
vectors, horizon cut, compaction, az/za) is the CUDA kernel ``fv_rotate_cut`` (csrc/rotate_cut.cu).

The reference obtains topocentric vectors from matvis' ``CoordinateRotationERFA`` /
``CoordinateRotationAstropy`` (call sites /root/reference/src/fftvis/cpu/cpu_simulate.py:693-709,
937-940).  ``coordinate_blocks`` is this repo's form of that manager: ``coord_method`` selects the
ERFA-structured model of ``core/astrometry.py`` (both matvis names map to it: they are the same
transformation through two libraries) or the explicit Earth-rotation-only model
(``"CoordinateRotationERA"``); ``coord_method_params`` carries matvis' ``update_bcrs_every`` plus
the IERS quantities that cannot be looked up offline (``dut1`` seconds, ``xp`` / ``yp`` radians) and
the escape hatches ``rotation_matrices`` / ``astrom`` / ``eq_xyz`` for callers who run erfa themselves.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from . import astrometry
from .utils import memo_by_value

TWO_PI = 2.0 * np.pi


@dataclass(frozen=True)
class TelescopeLocation:
    """Geodetic site: the stand-in for ``astropy.coordinates.EarthLocation``."""

    lat_deg: float
    lon_deg: float
    height_m: float = 0.0

    @property
    def lat_rad(self) -> float:
        return float(np.deg2rad(self.lat_deg))

    @property
    def lon_rad(self) -> float:
        return float(np.deg2rad(self.lon_deg))


HERA_LOCATION = TelescopeLocation(-30.72152612068957, 21.428303826863015, 1051.69)


def _angle_rad(q) -> float:
    for attr in ("rad", "radian"):
        if hasattr(q, attr):
            return float(getattr(q, attr))
    if hasattr(q, "to_value"):
        return float(q.to_value("rad"))
    return float(q)


def site_height(telescope_loc) -> float:
    """Height above the WGS84 ellipsoid in metres (0 when the location does not carry one)."""
    if isinstance(telescope_loc, TelescopeLocation):
        return float(telescope_loc.height_m)
    h = getattr(telescope_loc, "height", None)
    if h is not None:
        return float(h.to_value("m")) if hasattr(h, "to_value") else float(h)
    if hasattr(telescope_loc, "lat"):
        return 0.0
    t = tuple(telescope_loc)
    return float(t[2]) if len(t) > 2 else 0.0


def site_lat_lon(telescope_loc) -> tuple[float, float]:
    """(lat, lon) in radians from a TelescopeLocation, an astropy EarthLocation (duck-typed:
    ``.lat``/``.lon`` angles) or a ``(lat_deg, lon_deg[, height])`` tuple."""
    if isinstance(telescope_loc, TelescopeLocation):
        return telescope_loc.lat_rad, telescope_loc.lon_rad
    if hasattr(telescope_loc, "lat") and hasattr(telescope_loc, "lon"):
        return _angle_rad(telescope_loc.lat), _angle_rad(telescope_loc.lon)
    t = tuple(telescope_loc)
    return float(np.deg2rad(t[0])), float(np.deg2rad(t[1]))


def times_to_jd(times) -> np.ndarray:
    """UTC Julian dates from an ndarray or an astropy ``Time`` (duck-typed ``.utc.jd`` / ``.jd``);
    the reference wraps bare arrays as ``Time(times, format="jd")``, i.e. UTC (cpu_simulate.py:688-689)."""
    if hasattr(times, "utc"):
        try:
            return np.atleast_1d(np.asarray(times.utc.jd, dtype=np.float64))
        except Exception:
            pass
    if hasattr(times, "jd"):
        return np.atleast_1d(np.asarray(times.jd, dtype=np.float64))
    return np.atleast_1d(np.asarray(times, dtype=np.float64))


def earth_rotation_angle(jd_ut1: np.ndarray) -> np.ndarray:
    """IAU 2000 Earth rotation angle (radians, [0, 2pi))."""
    return astrometry.earth_rotation_angle(jd_ut1)


# the per-time blocks depend on (times, site, IERS parameters) only: memoised on those values
_astrom_blocks = memo_by_value()(astrometry.astrom_blocks)


def coordinate_blocks(times, telescope_loc, coord_method: str = "CoordinateRotationERFA",
                      coord_method_params: dict | None = None) -> tuple[np.ndarray, np.ndarray | None]:
    """Per-time blocks ``(enu_mats (nt, 3, 3), astrom (nt, 10) or None)`` handed to ``fv_rotate_cut``.

    ``coord_method``: ``"CoordinateRotationERFA"`` / ``"CoordinateRotationAstropy"`` (the reference's
    two managers, cpu_simulate.py:693; here one model: IAU 2006/2000 bias-precession-nutation, CIO
    locator, annual + diurnal aberration, solar light deflection, polar motion) or
    ``"CoordinateRotationERA"`` (Earth rotation angle + latitude only).  Unknown names raise
    ``KeyError`` like the reference's ``CoordinateRotation._methods[coord_method]`` lookup."""
    params = dict(coord_method_params or {})
    if coord_method not in astrometry.COORD_METHODS:
        raise KeyError(coord_method)
    known = {"update_bcrs_every", "dut1", "xp", "yp", "rotation_matrices", "astrom", "eq_xyz"}
    extra = set(params) - known
    if extra:
        raise TypeError(f"unexpected coord_method_params for {coord_method}: {sorted(extra)}")
    if "rotation_matrices" in params:
        mats = np.ascontiguousarray(params["rotation_matrices"], dtype=np.float64)
        ast = params.get("astrom")
        return mats, (None if ast is None else np.ascontiguousarray(ast, dtype=np.float64).reshape(len(mats), 10))
    lat, lon = site_lat_lon(telescope_loc)
    blk = _astrom_blocks(
        times_to_jd(times), lat, lon, site_height(telescope_loc), dut1=float(params.get("dut1", 0.0)),
        xp=float(params.get("xp", 0.0)), yp=float(params.get("yp", 0.0)),
        update_bcrs_every=float(params.get("update_bcrs_every", 0.0)),
        era_only=coord_method == "CoordinateRotationERA")
    return blk["enu"], (None if coord_method == "CoordinateRotationERA" else blk["astrom"])


def eq_to_enu_matrices(times, telescope_loc) -> np.ndarray:
    """(ntimes, 3, 3) fp64 matrices of the Earth-rotation-only model: enu = M @ (cos d cos a, cos d sin a, sin d)."""
    return coordinate_blocks(times, telescope_loc, "CoordinateRotationERA")[0]


def equatorial_unit_vectors(ra, dec) -> np.ndarray:
    """(3, Ns) fp64 unit vectors from ra/dec (radians), evaluated in the dtype-cast inputs
    (the reference casts ra/dec to the working precision first, cpu_simulate.py:601-604)."""
    ra = np.asarray(ra).astype(np.float64)
    dec = np.asarray(dec).astype(np.float64)
    cd = np.cos(dec)
    return np.stack([cd * np.cos(ra), cd * np.sin(ra), np.sin(dec)])


def enu_to_az_za(enu_e, enu_n, orientation: str = "uvbeam"):
    """East/north direction cosines -> (az, za).  ``uvbeam``: azimuth from East toward North in
    [0, 2pi) (matvis coordinates.enu_to_az_za as recalled in SURVEY.md Appendix B.2; call site
    reference cpu_simulate.py:957-959)."""
    e = np.asarray(enu_e)
    n = np.asarray(enu_n)
    r2 = e * e + n * n
    zeta = np.sqrt(np.clip(1.0 - r2, 0.0, None))
    za = np.pi / 2 - np.arcsin(zeta)
    az = np.arctan2(e, n)
    if orientation == "uvbeam":
        az = np.pi / 2 - az
    elif orientation != "astropy":
        raise ValueError("orientation must be 'astropy' or 'uvbeam'")
    az = np.mod(az, TWO_PI)
    return az.astype(e.dtype, copy=False), za.astype(e.dtype, copy=False)


class CoordinateRotation:
    """The argument the reference's chunk evaluator calls ``coord_mgr`` (matvis ``CoordinateRotation``,
    constructed at cpu_simulate.py:693-704 and consumed at :856-940): the catalogue as coherency, the
    observation times, the site, the source positions and the chunking.  Here it only *carries* those
    inputs -- the rotation itself is the ``fv_rotate_cut`` kernel -- so that
    ``GPUSimulationEngine._evaluate_vis_chunk`` can be called with the reference's own keyword set.
    ``skycoords`` is an astropy ``SkyCoord`` (duck-typed ``.ra.rad`` / ``.dec.rad``) or an ``(ra, dec)``
    pair in radians.  A matvis manager passed instead is read through the same attribute names."""

    _methods = {name: name for name in astrometry.COORD_METHODS}
    method = "CoordinateRotationERFA"

    def __init__(self, flux, times, telescope_loc, skycoords, chunk_size=None, source_buffer=1.0, precision=2,
                 method=None, **coord_method_params):
        self.flux = np.asarray(flux)
        self.times = times
        self.telescope_loc = telescope_loc
        self.skycoords = skycoords
        self.nsrc = int(np.shape(self.flux)[0])
        self.chunk_size = int(chunk_size) if chunk_size else self.nsrc
        self.source_buffer = float(source_buffer)
        self.precision = int(precision)
        if method is not None:
            if method not in astrometry.COORD_METHODS:
                raise KeyError(method)
            self.method = method
        self.coord_method_params = dict(coord_method_params)

    def setup(self):          # matvis allocates its device/host buffers here; nothing to do
        return None


def manager_inputs(coord_mgr) -> dict:
    """What ``_evaluate_vis_chunk`` needs from a ``coord_mgr`` (this module's ``CoordinateRotation`` or a
    matvis one, duck-typed): ra/dec in radians, flux, times, site, chunk size, buffer, method + parameters."""
    sky = coord_mgr.skycoords
    if hasattr(sky, "ra") and hasattr(sky, "dec"):
        ra = np.asarray(getattr(sky.ra, "rad", sky.ra), dtype=np.float64)
        dec = np.asarray(getattr(sky.dec, "rad", sky.dec), dtype=np.float64)
    else:
        ra, dec = (np.asarray(v, dtype=np.float64) for v in sky)
    method = getattr(coord_mgr, "method", None) or type(coord_mgr).__name__
    if method not in astrometry.COORD_METHODS:
        method = "CoordinateRotationERFA"
    params = dict(getattr(coord_mgr, "coord_method_params", {}) or {})
    if "update_bcrs_every" not in params and hasattr(coord_mgr, "update_bcrs_every"):
        ub = coord_mgr.update_bcrs_every
        params["update_bcrs_every"] = float(ub.to_value("s")) if hasattr(ub, "to_value") else float(ub)
    return dict(ra=ra, dec=dec, flux=np.asarray(coord_mgr.flux), times=coord_mgr.times,
                telescope_loc=coord_mgr.telescope_loc, chunk_size=int(coord_mgr.chunk_size),
                source_buffer=float(getattr(coord_mgr, "source_buffer", 1.0)), method=method, params=params)
