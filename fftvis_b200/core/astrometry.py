"""Per-time astrometry block of stage a1: ICRS -> observed (topocentric ENU), host part.

The reference gets topocentric unit vectors from matvis' ``CoordinateRotationERFA`` /
``CoordinateRotationAstropy`` (call sites /root/reference/src/fftvis/cpu/cpu_simulate.py:693-709,
913, 937-940), i.e. the ERFA chain  ``apco`` (star-independent parameters per time) ->
``atciqz`` (light deflection by the Sun, annual + diurnal aberration, bias-precession-nutation)
-> ``atioq`` without refraction (Earth rotation, polar motion, latitude).  matvis, astropy and erfa
are not installable offline, so that chain is built here from the published IAU/SOFA algorithms:

* time scales: UTC -> TT through the leap-second table, UTC -> UT1 through a caller-supplied
  ``dut1`` (IERS data are not available offline; default 0 s);
* bias-precession-nutation: IAU 2006 precession in Fukushima-Williams angles (the ``pfw06``
  polynomials), nutation = the leading luni-solar terms of the IAU 2000 series with the 2000B
  planetary offset and the IAU 2006 J2 adjustment (``nut06a``'s correction), composed as
  ``fw2m(gamb, phib, psib + dpsi, epsa + deps)``; CIP X, Y from that matrix, CIO locator ``s``
  from the ``s06`` polynomial + leading periodic terms, ``c2ixys`` for the CIO-based matrix;
* Earth ephemeris (``epv00`` stand-in): Keplerian mean elements of the Earth-Moon barycentre,
  Jupiter and Saturn (Standish's approximate elements), a three-term lunar orbit for the
  Earth-EMB offset; velocities by central differences;
* observer: WGS84 geodetic -> geocentric, ``pvtob`` with polar motion and ``s'``; ``apcs``.

Truncation budget (worst case, documented in DESIGN.md): nutation <= ~1 mas, aberration ~1-2 mas
(ephemeris velocity ~1e-4 relative), everything else < 0.1 mas -- against ~18 arcmin for the
Earth-rotation-only model this replaces.  The per-SOURCE arithmetic (deflection, aberration, the
3x3 rotation, horizon cut) runs on the GPU in ``fv_rotate_cut`` from the ``fv_astrom`` block made
here; ``update_bcrs_every`` freezes the ICRS -> CIRS part (deflection, aberration, NPB) for that
many seconds, like matvis.
"""
from __future__ import annotations

import math

import numpy as np

AS2R = math.pi / (180.0 * 3600.0)
TWO_PI = 2.0 * math.pi
DJ00 = 2451545.0
DJC = 36525.0
DAU = 149597870.7e3          # m
C_LIGHT = 299792458.0        # m / s
DAYSEC = 86400.0
AULT = DAU / C_LIGHT         # light time for 1 au (s)
SRS = 1.97412574336e-8       # Schwarzschild radius of the Sun (au)
WGS84_A = 6378137.0
WGS84_F = 1.0 / 298.257223563
EARTH_OM = 1.00273781191135448 * TWO_PI / DAYSEC      # rad per UT1 second

COORD_METHODS = ("CoordinateRotationERFA", "CoordinateRotationAstropy", "CoordinateRotationERA")

# (JD of the UTC day the step takes effect, TAI - UTC in seconds)
_LEAP = (
    (2441317.5, 10), (2441499.5, 11), (2441683.5, 12), (2442048.5, 13), (2442413.5, 14), (2442778.5, 15),
    (2443144.5, 16), (2443509.5, 17), (2443874.5, 18), (2444239.5, 19), (2444786.5, 20), (2445151.5, 21),
    (2445516.5, 22), (2446247.5, 23), (2447161.5, 24), (2447892.5, 25), (2448257.5, 26), (2448804.5, 27),
    (2449169.5, 28), (2449534.5, 29), (2450083.5, 30), (2450630.5, 31), (2451179.5, 32), (2453736.5, 33),
    (2454832.5, 34), (2456109.5, 35), (2457204.5, 36), (2457754.5, 37),
)

# Leading luni-solar nutation terms (IAU 2000B ordering; units 0.1 microarcsecond):
#   l  l'  F  D  Om |  psi_sin  t*psi_sin  psi_cos |  eps_cos  t*eps_cos  eps_sin
_NUT = np.array([
    [0, 0, 0, 0, 1, -172064161.0, -174666.0, 33386.0, 92052331.0, 9086.0, 15377.0],
    [0, 0, 2, -2, 2, -13170906.0, -1675.0, -13696.0, 5730336.0, -3015.0, -4587.0],
    [0, 0, 2, 0, 2, -2276413.0, -234.0, 2796.0, 978459.0, -485.0, 1374.0],
    [0, 0, 0, 0, 2, 2074554.0, 207.0, -698.0, -897492.0, 470.0, -291.0],
    [0, 1, 0, 0, 0, 1475877.0, -3633.0, 11817.0, 73871.0, -184.0, -1924.0],
    [0, 1, 2, -2, 2, -516821.0, 1226.0, -524.0, 224386.0, -677.0, -174.0],
    [1, 0, 0, 0, 0, 711159.0, 73.0, -872.0, -6750.0, 0.0, 358.0],
    [0, 0, 2, 0, 1, -387298.0, -367.0, 380.0, 200728.0, 18.0, 318.0],
    [1, 0, 2, 0, 2, -301461.0, -36.0, 816.0, 129025.0, -63.0, 367.0],
    [0, -1, 2, -2, 2, 215829.0, -494.0, 111.0, -95929.0, 299.0, 132.0],
    [0, 0, 2, -2, 1, 128227.0, 137.0, 181.0, -68982.0, -9.0, 39.0],
    [-1, 0, 2, 0, 2, 123457.0, 11.0, 19.0, -53311.0, 32.0, -4.0],
    [-1, 0, 0, 2, 0, 156994.0, 10.0, -168.0, -1235.0, 0.0, 82.0],
    [1, 0, 0, 0, 1, 63110.0, 63.0, 27.0, -33228.0, 0.0, -9.0],
    [-1, 0, 0, 0, 1, -57976.0, -63.0, -189.0, 31429.0, 0.0, -75.0],
    [-1, 0, 2, 2, 2, -59641.0, -11.0, 149.0, 25543.0, -11.0, 66.0],
    [1, 0, 2, 0, 1, -51613.0, -42.0, 129.0, 26366.0, 0.0, 78.0],
    [-2, 0, 2, 0, 1, 45893.0, 50.0, 31.0, -24236.0, -10.0, 20.0],
    [0, 0, 0, 2, 0, 63384.0, 11.0, -150.0, -1220.0, 0.0, 29.0],
    [0, 0, 2, 2, 2, -38571.0, -1.0, 158.0, 16452.0, -11.0, 68.0],
    [0, -2, 2, -2, 2, 32481.0, 0.0, 0.0, -13870.0, 0.0, 0.0],
    [-2, 0, 0, 2, 0, -47722.0, 0.0, -18.0, 477.0, 0.0, -25.0],
    [2, 0, 2, 0, 2, -31046.0, -1.0, 131.0, 13238.0, -11.0, 59.0],
    [1, 0, 2, -2, 2, 28593.0, 0.0, -1.0, -12338.0, 10.0, -3.0],
    [-1, 0, 2, 0, 1, 20441.0, 21.0, 10.0, -10758.0, 0.0, -3.0],
    [2, 0, 0, 0, 0, 29243.0, 0.0, -74.0, -609.0, 0.0, 13.0],
    [0, 0, 2, 0, 0, 25887.0, 0.0, -66.0, -550.0, 0.0, 11.0],
    [0, 1, 0, 0, 1, -14053.0, -25.0, 79.0, 8551.0, -2.0, -45.0],
    [-1, 0, 0, 2, 1, 15164.0, 10.0, 11.0, -8001.0, 0.0, -1.0],
    [0, 2, 2, -2, 2, -15794.0, 72.0, -16.0, 6850.0, -42.0, -5.0],
    [0, 0, -2, 2, 0, 21783.0, 0.0, 13.0, -167.0, 0.0, 13.0],
    [1, 0, 0, -2, 1, -12873.0, -10.0, -37.0, 6953.0, 0.0, -14.0],
    [0, -1, 0, 0, 1, -12654.0, 11.0, 63.0, 6415.0, 0.0, 26.0],
    [-1, 0, 2, 2, 1, -10204.0, 0.0, 25.0, 5222.0, 0.0, 15.0],
    [0, 2, 0, 0, 0, 16707.0, -85.0, -10.0, 168.0, -1.0, 10.0],
    [1, 0, 2, 2, 2, -7691.0, 0.0, 44.0, 3268.0, 0.0, 19.0],
    [-2, 0, 2, 0, 0, -11024.0, 0.0, -14.0, 104.0, 0.0, 2.0],
    [0, 1, 2, 0, 2, 7566.0, -21.0, -11.0, -3250.0, 0.0, -5.0],
    [0, 0, 2, 2, 1, -6637.0, -11.0, 25.0, 3353.0, 0.0, 14.0],
    [0, -1, 2, 0, 2, -7141.0, 21.0, 8.0, 3070.0, 0.0, 4.0],
    [0, 0, 0, 2, 1, -6302.0, -11.0, 2.0, 3272.0, 0.0, 4.0],
    [1, 0, 2, -2, 1, 5800.0, 10.0, 2.0, -3045.0, 0.0, -1.0],
    [2, 0, 2, -2, 2, 6443.0, 0.0, -7.0, -2768.0, 0.0, -4.0],
    [-2, 0, 0, 2, 1, -5774.0, -11.0, -15.0, 3041.0, 0.0, -5.0],
    [2, 0, 2, 0, 1, -5350.0, 0.0, 21.0, 2695.0, 0.0, 12.0],
    [0, -1, 2, -2, 1, -4752.0, -11.0, -3.0, 2719.0, 0.0, -3.0],
    [0, 0, 0, -2, 1, -4940.0, -11.0, -21.0, 2720.0, 0.0, -9.0],
    [-1, -1, 0, 2, 0, 7350.0, 0.0, -8.0, -51.0, 0.0, 4.0],
    [2, 0, 0, -2, 1, 4065.0, 0.0, 6.0, -2206.0, 0.0, 1.0],
    [1, 0, 0, 2, 0, 6579.0, 0.0, -24.0, -199.0, 0.0, 2.0],
])
_NUT_PLANETARY = (-0.135e-3 * AS2R, 0.388e-3 * AS2R)     # fixed offset standing in for the planetary terms

# CIO locator s + XY/2: polynomial (arcsec) and leading periodic terms (l, l', F, D, Om | sin, cos) in arcsec
_S06_POLY = (94.00e-6, 3808.65e-6, -122.68e-6, -72574.11e-6, 27.98e-6, 15.62e-6)
_S06_T0 = np.array([
    [0, 0, 0, 0, 1, -2640.73e-6, 0.39e-6],
    [0, 0, 0, 0, 2, -63.53e-6, 0.02e-6],
    [0, 0, 2, -2, 3, -11.75e-6, -0.01e-6],
    [0, 0, 2, -2, 1, -11.21e-6, -0.01e-6],
    [0, 0, 2, -2, 2, 4.57e-6, 0.00e-6],
    [0, 0, 2, 0, 3, -2.02e-6, 0.00e-6],
    [0, 0, 2, 0, 1, -1.98e-6, 0.00e-6],
    [0, 0, 0, 0, 3, 1.72e-6, 0.00e-6],
    [0, 1, 0, 0, 1, 1.41e-6, 0.01e-6],
    [0, 1, 0, 0, -1, 1.26e-6, 0.01e-6],
    [1, 0, 0, 0, -1, 0.63e-6, 0.00e-6],
    [1, 0, 0, 0, 1, 0.63e-6, 0.00e-6],
])
_S06_T1 = np.array([[0, 0, 0, 0, 2, -0.07e-6, 3.57e-6], [0, 0, 0, 0, 1, 1.73e-6, -0.03e-6]])
_S06_T2 = np.array([[0, 0, 0, 0, 1, 743.52e-6, -0.17e-6], [0, 0, 2, -2, 2, 56.91e-6, 0.06e-6],
                    [0, 0, 2, 0, 2, 9.84e-6, -0.01e-6], [0, 0, 0, 0, 2, -8.85e-6, 0.01e-6]])

# Approximate Keplerian elements (J2000 ecliptic; au, deg, deg / century): a, da, e, de, I, dI, L, dL, peri, dperi, node, dnode
_ELEMENTS = {
    "emb": (1.00000261, 0.00000562, 0.01671123, -0.00004392, -0.00001531, -0.01294668,
            100.46457166, 35999.37244981, 102.93768193, 0.32327364, 0.0, 0.0),
    "jupiter": (5.20288700, -0.00011607, 0.04838624, -0.00013253, 1.30439695, -0.00183714,
                34.39644051, 3034.74612775, 14.72847983, 0.21252668, 100.47390909, 0.20469106),
    "saturn": (9.53667594, -0.00125060, 0.05386179, -0.00050991, 2.48599187, 0.00193609,
               49.95424423, 1222.49362201, 92.59887831, -0.41897216, 113.66242448, -0.28867794),
}
_INV_MASS = {"jupiter": 1047.348644, "saturn": 3497.9018}     # Sun / planet-system mass
_EARTH_MOON_MASS = 81.30056
_EPS0 = 84381.406 * AS2R                                      # J2000 mean obliquity (IAU 2006)


# ---------------------------------------------------------------------------------------------
# time scales
# ---------------------------------------------------------------------------------------------
def tai_minus_utc(jd_utc: float) -> float:
    out = 10.0
    for jd0, dat in _LEAP:
        if jd_utc >= jd0:
            out = float(dat)
    return out


def julian_centuries_tt(jd_utc) -> np.ndarray:
    jd = np.atleast_1d(np.asarray(jd_utc, dtype=np.float64))
    dtt = np.array([tai_minus_utc(float(j)) + 32.184 for j in jd]) / DAYSEC
    return ((jd - DJ00) + dtt) / DJC


def earth_rotation_angle(jd_ut1) -> np.ndarray:
    """IAU 2000 Earth rotation angle, radians in [0, 2 pi) (``era00``)."""
    jd = np.atleast_1d(np.asarray(jd_ut1, dtype=np.float64))
    d = jd - DJ00
    frac = np.mod(jd, 1.0)       # (jd - DJ00) differs from jd by a half-integer + integer: handled by the 0.779.. constant
    return TWO_PI * np.mod(frac + 0.7790572732640 + 0.00273781191135448 * d, 1.0)


# ---------------------------------------------------------------------------------------------
# rotation helpers (frame rotations, SOFA convention: positive angle rotates the FRAME anticlockwise)
# ---------------------------------------------------------------------------------------------
def _rx(a):
    c, s = math.cos(a), math.sin(a)
    return np.array([[1.0, 0.0, 0.0], [0.0, c, s], [0.0, -s, c]])


def _ry(a):
    c, s = math.cos(a), math.sin(a)
    return np.array([[c, 0.0, -s], [0.0, 1.0, 0.0], [s, 0.0, c]])


def _rz(a):
    c, s = math.cos(a), math.sin(a)
    return np.array([[c, s, 0.0], [-s, c, 0.0], [0.0, 0.0, 1.0]])


# ---------------------------------------------------------------------------------------------
# precession-nutation
# ---------------------------------------------------------------------------------------------
def fundamental_arguments(t: float):
    """Delaunay arguments l, l', F, D, Omega (radians; IERS 2003 polynomials), t = TT centuries."""
    def poly(c):
        return math.fmod(c[0] + t * (c[1] + t * (c[2] + t * (c[3] + t * c[4]))), 1296000.0) * AS2R
    return (poly((485868.249036, 1717915923.2178, 31.8792, 0.051635, -0.00024470)),
            poly((1287104.793048, 129596581.0481, -0.5532, 0.000136, -0.00001149)),
            poly((335779.526232, 1739527262.8478, -12.7512, -0.001037, 0.00000417)),
            poly((1072260.703692, 1602961601.2090, -6.3706, 0.006593, -0.00003169)),
            poly((450160.398036, -6962890.5431, 7.4722, 0.007702, -0.00005939)))


def nutation(t: float):
    """(dpsi, deps) in radians: truncated IAU 2000 luni-solar series + planetary offset, with the
    IAU 2006 adjustments of ``nut06a``."""
    fa = np.array(fundamental_arguments(t))
    arg = _NUT[:, :5] @ fa
    sa, ca = np.sin(arg), np.cos(arg)
    dpsi = float(np.sum((_NUT[:, 5] + _NUT[:, 6] * t) * sa + _NUT[:, 7] * ca)) * 1e-7 * AS2R + _NUT_PLANETARY[0]
    deps = float(np.sum((_NUT[:, 8] + _NUT[:, 9] * t) * ca + _NUT[:, 10] * sa)) * 1e-7 * AS2R + _NUT_PLANETARY[1]
    fj2 = -2.7774e-6 * t
    return dpsi + dpsi * (0.4697e-6 + fj2), deps + deps * fj2


def fukushima_williams(t: float):
    """IAU 2006 bias-precession angles gamma_bar, phi_bar, psi_bar and the mean obliquity (``pfw06``)."""
    gamb = (-0.052928 + (10.556378 + (0.4932044 + (-0.00031238 + (-0.000002788 + 0.0000000260 * t) * t) * t) * t) * t) * AS2R
    phib = (84381.412819 + (-46.811016 + (0.0511268 + (0.00053289 + (-0.000000440 - 0.0000000176 * t) * t) * t) * t) * t) * AS2R
    psib = (-0.041775 + (5038.481484 + (1.5584175 + (-0.00018522 + (-0.000026452 - 0.0000000148 * t) * t) * t) * t) * t) * AS2R
    epsa = (84381.406 + (-46.836769 + (-0.0001831 + (0.00200340 + (-0.000000576 - 0.0000000434 * t) * t) * t) * t) * t) * AS2R
    return gamb, phib, psib, epsa


def npb_matrix(t: float) -> np.ndarray:
    """Bias-precession-nutation matrix GCRS -> true equator and equinox of date (``pnm06a``)."""
    gamb, phib, psib, epsa = fukushima_williams(t)
    dpsi, deps = nutation(t)
    return _rx(-(epsa + deps)) @ _rz(-(psib + dpsi)) @ _rx(phib) @ _rz(gamb)


def cio_locator(t: float, x: float, y: float) -> float:
    """CIO locator s (radians) given the CIP coordinates (``s06``, leading terms)."""
    fa = np.array(fundamental_arguments(t))

    def series(tab):
        a = tab[:, :5] @ fa
        return float(np.sum(tab[:, 5] * np.sin(a) + tab[:, 6] * np.cos(a)))
    w = list(_S06_POLY)
    w[0] += series(_S06_T0)
    w[1] += series(_S06_T1)
    w[2] += series(_S06_T2)
    s = w[0] + (w[1] + (w[2] + (w[3] + (w[4] + w[5] * t) * t) * t) * t) * t
    return s * AS2R - x * y / 2.0


def c2i_matrix(x: float, y: float, s: float) -> np.ndarray:
    """Celestial-to-intermediate matrix from CIP X, Y and the CIO locator (``c2ixys``)."""
    r2 = x * x + y * y
    e = math.atan2(y, x) if r2 > 0.0 else 0.0
    d = math.atan(math.sqrt(r2 / (1.0 - r2)))
    return _rz(-(e + s)) @ _ry(d) @ _rz(e)


# ---------------------------------------------------------------------------------------------
# ephemeris (epv00 stand-in)
# ---------------------------------------------------------------------------------------------
def _kepler_xyz(name: str, T: float) -> np.ndarray:
    """Heliocentric J2000-ecliptic position (au) from the approximate mean elements."""
    a0, da, e0, de, i0, di, l0, dl, p0, dp, n0, dn = _ELEMENTS[name]
    a, e = a0 + da * T, e0 + de * T
    inc, L, peri, node = (math.radians(v) for v in (i0 + di * T, l0 + dl * T, p0 + dp * T, n0 + dn * T))
    M = math.fmod(L - peri, TWO_PI)
    E = M
    for _ in range(12):
        E = E - (E - e * math.sin(E) - M) / (1.0 - e * math.cos(E))
    xp, yp = a * (math.cos(E) - e), a * math.sqrt(1.0 - e * e) * math.sin(E)
    w = peri - node
    cw, sw, cn, sn, ci, si = math.cos(w), math.sin(w), math.cos(node), math.sin(node), math.cos(inc), math.sin(inc)
    return np.array([(cw * cn - sw * sn * ci) * xp + (-sw * cn - cw * sn * ci) * yp,
                     (cw * sn + sw * cn * ci) * xp + (-sw * sn + cw * cn * ci) * yp,
                     (sw * si) * xp + (cw * si) * yp])


def _moon_geocentric(T: float) -> np.ndarray:
    """Geocentric ecliptic-of-date position of the Moon (au), three-term orbit (good to ~1 %)."""
    d = T * DJC
    Lm = math.radians(218.316 + 13.176396 * d)
    Mm = math.radians(134.963 + 13.064993 * d)
    F = math.radians(93.272 + 13.229350 * d)
    lam = Lm + math.radians(6.289) * math.sin(Mm)
    beta = math.radians(5.128) * math.sin(F)
    r = (385001.0 - 20905.0 * math.cos(Mm)) * 1e3 / DAU
    return r * np.array([math.cos(beta) * math.cos(lam), math.cos(beta) * math.sin(lam), math.sin(beta)])


def _earth_positions(T: float):
    """(heliocentric, barycentric) Earth position, J2000 ecliptic, au."""
    emb = _kepler_xyz("emb", T)
    earth_h = emb - _moon_geocentric(T) / (1.0 + _EARTH_MOON_MASS)
    msum, sun_b = 1.0, np.zeros(3)
    for name, inv in _INV_MASS.items():
        sun_b = sun_b - _kepler_xyz(name, T) / inv
        msum += 1.0 / inv
    sun_b /= msum
    return earth_h, earth_h + sun_b


def earth_posvel(t: float):
    """Earth heliocentric position (au) and barycentric position / velocity (au, au / day) in the
    ICRS-aligned equatorial frame at TT century ``t`` (the quantities ``apco`` takes from ``epv00``)."""
    h = 0.05 / DJC                                 # +- 0.05 day central difference
    eh, eb = _earth_positions(t)
    _, eb_p = _earth_positions(t + h)
    _, eb_m = _earth_positions(t - h)
    vel = (eb_p - eb_m) / 0.1
    rot = _rx(-_EPS0)                              # ecliptic -> equatorial
    return rot @ eh, rot @ eb, rot @ vel


# ---------------------------------------------------------------------------------------------
# observer + the astrom block
# ---------------------------------------------------------------------------------------------
def geodetic_to_geocentric(lon: float, lat: float, height: float) -> np.ndarray:
    """WGS84 geodetic -> geocentric xyz in metres (``gd2gc``)."""
    sp, cp = math.sin(lat), math.cos(lat)
    w = (1.0 - WGS84_F) ** 2
    ac = WGS84_A / math.sqrt(cp * cp + w * sp * sp)
    r = (ac + height) * cp
    return np.array([r * math.cos(lon), r * math.sin(lon), (w * ac + height) * sp])


def astrom_blocks(jd_utc, lat: float, lon: float, height: float, dut1: float = 0.0, xp: float = 0.0,
                  yp: float = 0.0, update_bcrs_every: float = 0.0, era_only: bool = False) -> dict:
    """Star-independent parameters of every time step.

    Returns ``enu`` (nt, 3, 3): the matrix taking the (deflected, aberrated) ICRS direction to local
    East-North-Up = latitude tilt . linearised polar motion . R3(local ERA) . C2I; and ``astrom``
    (nt, 10): Sun -> observer unit vector (3), its length in au, observer barycentric velocity / c
    (3), sqrt(1 - v^2), the deflection limiter, and a flag (0: skip deflection + aberration).
    ``era_only`` gives the Earth-rotation-angle + latitude model (no NPB, aberration, deflection)."""
    jd = np.atleast_1d(np.asarray(jd_utc, dtype=np.float64))
    nt = jd.size
    enu = np.zeros((nt, 3, 3))
    astrom = np.zeros((nt, 10))
    sphi, cphi = math.sin(lat), math.cos(lat)
    tilt = np.array([[0.0, 1.0, 0.0], [-sphi, 0.0, cphi], [cphi, 0.0, sphi]])      # (-HA, Dec) frame -> ENU
    theta = earth_rotation_angle(jd + dut1 / DAYSEC)
    if era_only:
        for i in range(nt):
            enu[i] = tilt @ _rz(float(theta[i]) + lon)
        astrom[:, 3], astrom[:, 7] = 1.0, 1.0
        return dict(enu=enu, astrom=astrom)
    tt = julian_centuries_tt(jd)
    frozen, last = None, None
    for i in range(nt):
        t = float(tt[i])
        th = float(theta[i])
        sp = -47e-6 * t * AS2R
        # CIRS -> apparent (-HA, Dec): Earth rotation, polar motion, longitude (apco's matrix r)
        r = _rz(lon) @ _rx(-yp) @ _ry(-xp) @ _rz(th + sp)
        eral = math.atan2(r[0, 1], r[0, 0])
        xpl = math.atan2(r[0, 2], math.hypot(r[0, 0], r[0, 1]))
        ypl = -math.atan2(r[1, 2], r[2, 2])
        pm = np.array([[1.0, 0.0, xpl], [0.0, 1.0, -ypl], [-xpl, ypl, 1.0]])     # atioq's linearised polar motion
        if frozen is None or update_bcrs_every <= 0.0 or abs(float(jd[i]) - last) * DAYSEC >= update_bcrs_every:
            npb = npb_matrix(t)
            x, y = float(npb[2, 0]), float(npb[2, 1])
            c2i = c2i_matrix(x, y, cio_locator(t, x, y))
            # observer's geocentric position / velocity, CIRS then GCRS (pvtob + trxpv)
            xyzm = geodetic_to_geocentric(lon, lat, height)
            pom = _rx(-yp) @ _ry(-xp) @ _rz(sp)
            ox, oy, oz = pom.T @ xyzm
            s_, c_ = math.sin(th), math.cos(th)
            pos = c2i.T @ np.array([c_ * ox - s_ * oy, s_ * ox + c_ * oy, oz])
            vel = c2i.T @ np.array([EARTH_OM * (-s_ * ox - c_ * oy), EARTH_OM * (c_ * ox - s_ * oy), 0.0])
            e_h, _, e_v = earth_posvel(t)
            ph = e_h + pos / DAU
            em = float(np.linalg.norm(ph))
            v = (e_v + vel / (DAU / DAYSEC)) * (AULT / DAYSEC)
            bm1 = math.sqrt(1.0 - float(v @ v))
            em2 = max(em * em, 1.0)
            frozen = (c2i, np.concatenate([ph / em, [em], v, [bm1, 1e-6 / em2, 1.0]]))
            last = float(jd[i])
        enu[i] = tilt @ pm @ _rz(eral) @ frozen[0]
        astrom[i] = frozen[1]
    return dict(enu=enu, astrom=astrom)


def apply_astrom(eq_xyz: np.ndarray, astrom_row: np.ndarray) -> np.ndarray:
    """Host form of the per-source part (``ldsun`` + ``ab``) for ONE time step: (3, n) ICRS unit
    vectors -> proper directions.  The product runs this on the GPU (csrc/rotate_cut.cu); this numpy
    form serves ``evaluate``-style host callers and the CPU tests of the block itself."""
    p = np.asarray(eq_xyz, dtype=np.float64)
    if astrom_row[9] == 0.0:
        return p
    e, em, v, bm1, dlim = astrom_row[0:3], astrom_row[3], astrom_row[4:7], astrom_row[7], astrom_row[8]
    qpe = p + e[:, None]
    w = SRS / em / np.maximum(np.sum(p * qpe, axis=0), dlim)
    eq = np.cross(e[:, None], p, axis=0)
    p1 = p + w * np.cross(p, eq, axis=0)
    pdv = v @ p1
    w1 = 1.0 + pdv / (1.0 + bm1)
    w2 = SRS / em
    q = p1 * bm1 + w1 * v[:, None] + w2 * (v[:, None] - pdv * p1)
    return q / np.linalg.norm(q, axis=0)
