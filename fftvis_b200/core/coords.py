"""Host-side astrometry for stage a1: everything that is *per time*, not
per source.  The per-source work (3x3 rotation of Ns unit vectors, horizon cut, compaction,
az/za) is the CUDA kernel ``fv_rotate_cut`` (csrc/rotate_cut.cuh).

The reference obtains topocentric vectors from matvis' ``CoordinateRotationERFA`` /
``CoordinateRotationAstropy`` (call sites /root/reference/src/fftvis/cpu/cpu_simulate.py:693-709,
937-940); matvis, astropy and erfa are absent from this image, so the rotation is stated here in
closed form: Earth-rotation-angle sidereal rotation about the celestial pole followed by the
latitude tilt.  Aberration, light deflection, precession-nutation and polar motion are NOT
applied (SURVEY.md section 8(f) rank 2 "coordinate manager on device" is the next row that adds
them); callers who have erfa can pass their own per-time matrices / apparent unit vectors via
``coord_method_params={"rotation_matrices": ..., "eq_xyz": ...}``.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

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
    """Julian dates (treated as UT1) from an ndarray or an astropy ``Time`` (duck-typed ``.jd``)."""
    if hasattr(times, "ut1"):
        try:
            return np.atleast_1d(np.asarray(times.ut1.jd, dtype=np.float64))
        except Exception:  # no IERS data offline: fall through to .jd
            pass
    if hasattr(times, "jd"):
        return np.atleast_1d(np.asarray(times.jd, dtype=np.float64))
    return np.atleast_1d(np.asarray(times, dtype=np.float64))


def earth_rotation_angle(jd_ut1: np.ndarray) -> np.ndarray:
    """IAU 2000 Earth rotation angle (radians, [0, 2pi)); JD split to keep fp64 precision."""
    jd = np.asarray(jd_ut1, dtype=np.float64)
    d = jd - 2451545.0
    frac = np.mod(jd, 1.0)  # (JD - 2451545.0) mod 1: the whole-turn part of Tu drops out
    theta = TWO_PI * np.mod(frac + 0.7790572732640 + 0.00273781191135448 * d, 1.0)
    return np.mod(theta, TWO_PI)


def eq_to_enu_matrices(times, telescope_loc) -> np.ndarray:
    """(ntimes, 3, 3) fp64 matrices M with  enu = M @ (cos d cos a, cos d sin a, sin d)."""
    lat, lon = site_lat_lon(telescope_loc)
    th = earth_rotation_angle(times_to_jd(times)) + lon
    c, s = np.cos(th), np.sin(th)
    sl, cl = np.sin(lat), np.cos(lat)
    m = np.zeros((th.size, 3, 3))
    # east
    m[:, 0, 0], m[:, 0, 1] = -s, c
    # north
    m[:, 1, 0], m[:, 1, 1], m[:, 1, 2] = -sl * c, -sl * s, cl
    # up
    m[:, 2, 0], m[:, 2, 1], m[:, 2, 2] = cl * c, cl * s, sl
    return m


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
