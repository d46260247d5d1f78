"""TEST INFRASTRUCTURE ONLY -- the oracle's own coordinate chain (nothing here imports
``fftvis_b200``).  Only tests/, __graft_entry__.smoke() and bench.py's CPU legs use it.

Restates what the reference obtains from matvis' coordinate managers (call sites
/root/reference/src/fftvis/cpu/cpu_simulate.py:693-709, 913, 937-940, 957-959):

    CoordinateRotationERFA / CoordinateRotationAstropy:
        ICRS --(erfa apco: star-independent parameters per time)
             --(atciqz: solar light deflection, aberration, bias-precession-nutation)--> CIRS
             --(atioq without refraction: Earth rotation angle, polar motion, latitude)--> az / zenith distance
             --> East / North / Up direction cosines
    enu_to_az_za(e, n, "uvbeam")

matvis / erfa / astropy are third-party and absent (matvis >= 1.3.2 pinned by the reference's
pyproject.toml:34; erfa transitively), so the published SOFA algorithms are restated function by
function with SOFA's own names (``era00``, ``pfw06``, ``fw2m``, ``bpn2xy``, ``s06``, ``c2ixys``,
``pom00``, ``gd2gc``, ``pvtob``, ``apcs``, ``ldsun``, ``ab``, ``atioq``).  PARITY UNPINNED against
erfa itself in this container; what pins it: the SOFA/ERFA published test values checked in
tests/test_astrometry.py (era00, pfw06 exact; nutation, pnm06a, s06, epv00 to the documented
truncation: < 2 mas), and an import-guarded comparison with ``erfa`` that runs wherever it exists.
The series coefficients (truncated luni-solar nutation, s06 terms, approximate planetary elements)
are DATA shared with the product through oracle/data/iau_series.json (a test asserts both copies are
equal); all CODE is separate and follows a different route from the product's: the product folds
everything after the aberration into one 3x3 per time, the oracle goes through spherical RA/Dec,
hour angle and az / zenith distance like ``atioq``.
"""
from __future__ import annotations

import json
import math
from pathlib import Path

import numpy as np

_D = json.loads((Path(__file__).resolve().parent / "data" / "iau_series.json").read_text())

DAS2R = 4.848136811095359935899141e-6
D2PI = 6.283185307179586476925287
DJ00, DJC, DAYSEC = 2451545.0, 36525.0, 86400.0
DAU = 149597870.7e3
CMPS = 299792458.0
AULT = DAU / CMPS
SRS = 1.97412574336e-8
METHODS = ("CoordinateRotationERFA", "CoordinateRotationAstropy", "CoordinateRotationERA")


# ---- small SOFA-style vector / matrix tools -------------------------------------------------------
def ir():
    return [[1.0, 0.0, 0.0], [0.0, 1.0, 0.0], [0.0, 0.0, 1.0]]


def rx(phi, r):
    s, c = math.sin(phi), math.cos(phi)
    a = [c * r[1][j] + s * r[2][j] for j in range(3)]
    b = [-s * r[1][j] + c * r[2][j] for j in range(3)]
    r[1], r[2] = a, b
    return r


def ry(theta, r):
    s, c = math.sin(theta), math.cos(theta)
    a = [c * r[0][j] - s * r[2][j] for j in range(3)]
    b = [s * r[0][j] + c * r[2][j] for j in range(3)]
    r[0], r[2] = a, b
    return r


def rz(psi, r):
    s, c = math.sin(psi), math.cos(psi)
    a = [c * r[0][j] + s * r[1][j] for j in range(3)]
    b = [-s * r[0][j] + c * r[1][j] for j in range(3)]
    r[0], r[1] = a, b
    return r


def rxp(r, p):
    return [sum(r[i][j] * p[j] for j in range(3)) for i in range(3)]


def trxp(r, p):
    return [sum(r[j][i] * p[j] for j in range(3)) for i in range(3)]


# ---- time ------------------------------------------------------------------------------------------
def dat(jd_utc):
    out = 10.0
    for jd0, d in _D["leap"]:
        if jd_utc >= jd0:
            out = float(d)
    return out


def tt_centuries(jd_utc):
    return ((jd_utc - DJ00) + (dat(jd_utc) + 32.184) / DAYSEC) / DJC


def era00(jd_ut1):
    t = jd_ut1 - DJ00
    f = math.fmod(jd_ut1, 1.0)
    return math.fmod(D2PI * (f + 0.7790572732640 + 0.00273781191135448 * t), D2PI) % D2PI


# ---- precession-nutation ----------------------------------------------------------------------------
def delaunay(t):
    c = ((485868.249036, 1717915923.2178, 31.8792, 0.051635, -0.00024470),
         (1287104.793048, 129596581.0481, -0.5532, 0.000136, -0.00001149),
         (335779.526232, 1739527262.8478, -12.7512, -0.001037, 0.00000417),
         (1072260.703692, 1602961601.2090, -6.3706, 0.006593, -0.00003169),
         (450160.398036, -6962890.5431, 7.4722, 0.007702, -0.00005939))
    return [math.fmod(k[0] + k[1] * t + k[2] * t**2 + k[3] * t**3 + k[4] * t**4, 1296000.0) * DAS2R for k in c]


def nut06a_truncated(t):
    fa = delaunay(t)
    dp = de = 0.0
    for row in reversed(_D["nut"]):                      # small terms first, like SOFA
        arg = math.fmod(sum(row[i] * fa[i] for i in range(5)), D2PI)
        sa, ca = math.sin(arg), math.cos(arg)
        dp += (row[5] + row[6] * t) * sa + row[7] * ca
        de += (row[8] + row[9] * t) * ca + row[10] * sa
    dpsi = dp * 1e-7 * DAS2R + _D["nut_planetary_mas"][0] * 1e-3 * DAS2R
    deps = de * 1e-7 * DAS2R + _D["nut_planetary_mas"][1] * 1e-3 * DAS2R
    fj2 = -2.7774e-6 * t
    return dpsi + dpsi * (0.4697e-6 + fj2), deps + deps * fj2


def pfw06(t):
    gamb = (-0.052928 + (10.556378 + (0.4932044 + (-0.00031238 + (-0.000002788 + (0.0000000260) * t) * t) * t) * t) * t) * DAS2R
    phib = (84381.412819 + (-46.811016 + (0.0511268 + (0.00053289 + (-0.000000440 + (-0.0000000176) * t) * t) * t) * t) * t) * DAS2R
    psib = (-0.041775 + (5038.481484 + (1.5584175 + (-0.00018522 + (-0.000026452 + (-0.0000000148) * t) * t) * t) * t) * t) * DAS2R
    epsa = (84381.406 + (-46.836769 + (-0.0001831 + (0.00200340 + (-0.000000576 + (-0.0000000434) * t) * t) * t) * t) * t) * DAS2R
    return gamb, phib, psib, epsa


def fw2m(gamb, phib, psi, eps):
    r = ir()
    rz(gamb, r)
    rx(phib, r)
    rz(-psi, r)
    rx(-eps, r)
    return r


def pnm06a(t):
    gamb, phib, psib, epsa = pfw06(t)
    dp, de = nut06a_truncated(t)
    return fw2m(gamb, phib, psib + dp, epsa + de)


def s06(t, x, y):
    fa = delaunay(t)

    def series(tab):
        tot = 0.0
        for row in reversed(tab):
            a = sum(row[i] * fa[i] for i in range(5))
            tot += row[5] * math.sin(a) + row[6] * math.cos(a)
        return tot
    sp = list(_D["s06_poly"])
    w0, w1, w2 = sp[0] + series(_D["s06_t0"]), sp[1] + series(_D["s06_t1"]), sp[2] + series(_D["s06_t2"])
    return (w0 + (w1 + (w2 + (sp[3] + (sp[4] + sp[5] * t) * t) * t) * t) * t) * DAS2R - x * y / 2.0


def c2ixys(x, y, s):
    r2 = x * x + y * y
    e = math.atan2(y, x) if r2 > 0.0 else 0.0
    d = math.atan(math.sqrt(r2 / (1.0 - r2)))
    r = ir()
    rz(e, r)
    ry(d, r)
    rz(-(e + s), r)
    return r


# ---- ephemeris stand-in for epv00 -----------------------------------------------------------------
def _kepler(name, T):
    a0, da, e0, de, i0, di, l0, dl, p0, dp, n0, dn = _D["elements"][name]
    a, e = a0 + da * T, e0 + de * T
    inc, L, peri, node = (math.radians(v) for v in (i0 + di * T, l0 + dl * T, p0 + dp * T, n0 + dn * T))
    M = math.fmod(L - peri, D2PI)
    E = M + e * math.sin(M)
    for _ in range(20):                                   # fixed-point form (the product uses Newton steps)
        E = M + e * math.sin(E)
    xo, yo = a * (math.cos(E) - e), a * math.sqrt(1 - e * e) * math.sin(E)
    # perifocal -> ecliptic: Rz(-node) Rx(-inc) Rz(-argp)
    argp = peri - node
    r = ir()
    rz(-argp, r)
    rx(-inc, r)
    rz(-node, r)
    return rxp(r, [xo, yo, 0.0])


def _moon(T):
    d = T * DJC
    Lm, Mm, F = (math.radians(v) for v in (218.316 + 13.176396 * d, 134.963 + 13.064993 * d, 93.272 + 13.229350 * d))
    lam, beta = Lm + math.radians(6.289) * math.sin(Mm), math.radians(5.128) * math.sin(F)
    rr = (385001.0 - 20905.0 * math.cos(Mm)) * 1e3 / DAU
    return [rr * math.cos(beta) * math.cos(lam), rr * math.cos(beta) * math.sin(lam), rr * math.sin(beta)]


def _earth_ecl(T):
    emb, moon = _kepler("emb", T), _moon(T)
    eh = [emb[i] - moon[i] / (1.0 + _D["earth_moon_mass"]) for i in range(3)]
    msum, sb = 1.0, [0.0, 0.0, 0.0]
    for name, inv in _D["inv_mass"].items():
        p = _kepler(name, T)
        sb = [sb[i] - p[i] / inv for i in range(3)]
        msum += 1.0 / inv
    return eh, [eh[i] + sb[i] / msum for i in range(3)]


def epv_approx(t):
    """(heliocentric position au, barycentric velocity au/day), equatorial ICRS-aligned."""
    h = 0.05 / DJC
    eh, _ = _earth_ecl(t)
    _, bp = _earth_ecl(t + h)
    _, bm = _earth_ecl(t - h)
    vel = [(bp[i] - bm[i]) / 0.1 for i in range(3)]
    r = rx(-84381.406 * DAS2R, ir())
    return rxp(r, eh), rxp(r, vel)


# ---- observer ------------------------------------------------------------------------------------
def gd2gc_wgs84(elong, phi, height):
    a, f = 6378137.0, 1.0 / 298.257223563
    sp, cp = math.sin(phi), math.cos(phi)
    w = (1.0 - f) ** 2
    d = cp * cp + w * sp * sp
    ac = a / math.sqrt(d)
    as_ = w * ac
    r = (ac + height) * cp
    return [r * math.cos(elong), r * math.sin(elong), (as_ + height) * sp]


def pom00(xp, yp, sp):
    r = ir()
    rz(sp, r)
    ry(-xp, r)
    rx(-yp, r)
    return r


def pvtob(elong, phi, hm, xp, yp, sp, theta):
    om = 1.00273781191135448 * D2PI / DAYSEC
    xyz = trxp(pom00(xp, yp, sp), gd2gc_wgs84(elong, phi, hm))
    s, c = math.sin(theta), math.cos(theta)
    x, y, z = xyz
    return [c * x - s * y, s * x + c * y, z], [om * (-s * x - c * y), om * (c * x - s * y), 0.0]


def apco(jd_utc, elong, phi, hm, dut1=0.0, xp=0.0, yp=0.0):
    """Star-independent parameters of one epoch (the erfa ``apco`` fields this path needs)."""
    t = tt_centuries(jd_utc)
    theta = era00(jd_utc + dut1 / DAYSEC)
    sp = -47e-6 * t * DAS2R
    r = ir()
    rz(theta + sp, r)
    ry(-xp, r)
    rx(-yp, r)
    rz(elong, r)
    a, b = r[0][0], r[0][1]
    eral = math.atan2(b, a) if (a != 0.0 or b != 0.0) else 0.0
    xpl = math.atan2(r[0][2], math.sqrt(a * a + b * b))
    a, b = r[1][2], r[2][2]
    ypl = -math.atan2(a, b) if (a != 0.0 or b != 0.0) else 0.0
    rbpn = pnm06a(t)
    x, y = rbpn[2][0], rbpn[2][1]
    bpn = c2ixys(x, y, s06(t, x, y))
    pc, vc = pvtob(elong, phi, hm, xp, yp, sp, theta)
    pos, vel = trxp(bpn, pc), trxp(bpn, vc)
    ehp, ebv = epv_approx(t)
    ph = [ehp[i] + pos[i] / DAU for i in range(3)]
    em = math.sqrt(sum(q * q for q in ph))
    v = [(ebv[i] + vel[i] / (DAU / DAYSEC)) * (AULT / DAYSEC) for i in range(3)]
    return dict(eh=[q / em for q in ph], em=em, v=v, bm1=math.sqrt(1.0 - sum(q * q for q in v)), bpn=bpn,
                eral=eral, xpl=xpl, ypl=ypl, sphi=math.sin(phi), cphi=math.cos(phi), theta=theta)


# ---- per-source steps (numpy over the source axis) -----------------------------------------------------
def ldsun(p, e, em):
    em2 = max(em * em, 1.0)
    dlim = 1e-6 / em2
    e = np.asarray(e)[:, None]
    qdqpe = np.maximum(np.einsum("is,is->s", p, p + e), dlim)
    w = SRS / em / qdqpe
    eq = np.cross(np.broadcast_to(e, p.shape), p, axis=0)
    return p + w * np.cross(p, eq, axis=0)


def ab(pnat, v, s, bm1):
    v = np.asarray(v)[:, None]
    pdv = np.einsum("is,is->s", pnat, np.broadcast_to(v, pnat.shape))
    w1 = 1.0 + pdv / (1.0 + bm1)
    w2 = SRS / s
    p = pnat * bm1 + w1 * v + w2 * (v - pdv * pnat)
    return p / np.sqrt(np.einsum("is,is->s", p, p))


def atioq_enu(ri, di, astrom):
    """CIRS RA/Dec -> East/North/Up direction cosines: ``atioq`` with zero refraction, returned as
    cartesian (e, n, u) = (sin az sin zd, cos az sin zd, cos zd)."""
    ha = ri - astrom["eral"]
    x, y, z = np.cos(ha) * np.cos(di), np.sin(ha) * np.cos(di), np.sin(di)      # s2c(ri - eral, di): -HA, Dec
    xpl, ypl = astrom["xpl"], astrom["ypl"]
    xhd, yhd, zhd = x + xpl * z, y - ypl * z, z - xpl * x + ypl * y
    sphi, cphi = astrom["sphi"], astrom["cphi"]
    xaet, yaet, zaet = sphi * xhd - cphi * zhd, yhd, cphi * xhd + sphi * zhd
    az = np.arctan2(yaet, -xaet)
    r = np.hypot(xaet, yaet)
    zd = np.arctan2(r, zaet)
    # the vector is unit to ~1e-12 (linearised polar motion): rebuild it from the two angles times its length
    nrm = np.sqrt(xaet**2 + yaet**2 + zaet**2)
    return nrm * np.sin(az) * np.sin(zd), nrm * np.cos(az) * np.sin(zd), nrm * np.cos(zd)


def site(telescope_loc):
    """(lat, lon, height) in radians / metres from a TelescopeLocation-like object, an astropy
    EarthLocation (duck-typed) or a (lat_deg, lon_deg[, height_m]) tuple."""
    if hasattr(telescope_loc, "lat_deg"):
        return (math.radians(telescope_loc.lat_deg), math.radians(telescope_loc.lon_deg),
                float(getattr(telescope_loc, "height_m", 0.0)))
    if hasattr(telescope_loc, "lat") and hasattr(telescope_loc, "lon"):
        def rad(q):
            return float(q.rad) if hasattr(q, "rad") else float(q.to_value("rad"))
        h = getattr(telescope_loc, "height", 0.0)
        return rad(telescope_loc.lat), rad(telescope_loc.lon), float(h.to_value("m")) if hasattr(h, "to_value") else float(h)
    t = tuple(telescope_loc)
    return math.radians(t[0]), math.radians(t[1]), float(t[2]) if len(t) > 2 else 0.0


def jd_of(times):
    if hasattr(times, "utc"):
        try:
            return np.atleast_1d(np.asarray(times.utc.jd, dtype=np.float64))
        except Exception:
            pass
    if hasattr(times, "jd"):
        return np.atleast_1d(np.asarray(times.jd, dtype=np.float64))
    return np.atleast_1d(np.asarray(times, dtype=np.float64))


def topocentric_enu(ra, dec, times, telescope_loc, coord_method="CoordinateRotationERFA",
                    coord_method_params=None):
    """(nt, 3, Ns) fp64 East/North/Up direction cosines of ICRS (ra, dec) at every time."""
    if coord_method not in METHODS:
        raise KeyError(coord_method)
    prm = dict(coord_method_params or {})
    ra = np.asarray(ra, dtype=np.float64)
    dec = np.asarray(dec, dtype=np.float64)
    lat, lon, hm = site(telescope_loc)
    jds = jd_of(times)
    out = np.empty((jds.size, 3, ra.size))
    p0 = np.stack([np.cos(dec) * np.cos(ra), np.cos(dec) * np.sin(ra), np.sin(dec)])
    if "rotation_matrices" in prm:
        # the caller's own per-time equatorial -> ENU matrices (the engine's escape hatch for users who run
        # erfa themselves): a plain rotation of the catalogue vectors
        mats = np.asarray(prm["rotation_matrices"], dtype=np.float64).reshape(jds.size, 3, 3)
        return np.einsum("tij,js->tis", mats, p0)
    upd = float(prm.get("update_bcrs_every", 0.0))
    held, held_jd = None, None
    for i, jd in enumerate(jds):
        jd = float(jd)
        if coord_method == "CoordinateRotationERA":
            ast = dict(eral=era00(jd + float(prm.get("dut1", 0.0)) / DAYSEC) + lon, xpl=0.0, ypl=0.0,
                       sphi=math.sin(lat), cphi=math.cos(lat))
            ri, di = ra, dec
        else:
            ast = apco(jd, lon, lat, hm, float(prm.get("dut1", 0.0)), float(prm.get("xp", 0.0)), float(prm.get("yp", 0.0)))
            if held is None or upd <= 0.0 or abs(jd - held_jd) * DAYSEC >= upd:
                ppr = ab(ldsun(p0, ast["eh"], ast["em"]), ast["v"], ast["em"], ast["bm1"])
                pi_ = np.asarray(ast["bpn"]) @ ppr
                held = (np.mod(np.arctan2(pi_[1], pi_[0]), D2PI), np.arctan2(pi_[2], np.hypot(pi_[0], pi_[1])))
                held_jd = jd
            ri, di = held
        out[i] = atioq_enu(ri, di, ast)
    return out


def rotation_matrices(times, telescope_loc, coord_method="CoordinateRotationERA", coord_method_params=None):
    """(nt, 3, 3) equatorial -> ENU matrices of this module's chain for the rotation-only part of
    ``coord_method`` (its action on the three coordinate axes, with no aberration / deflection): lets a
    test hand the SAME per-time rotation to the engine and to the direct sum, so that what is compared is
    the transform and not two roundings of the Earth rotation angle (a 1e-15 rad difference in the
    rotation is 2 pi f |b| / c ~ 1e3 times larger in the phase of a 300 m baseline at 200 MHz)."""
    ra = np.array([0.0, np.pi / 2, 0.0])
    dec = np.array([0.0, 0.0, np.pi / 2])
    enu = topocentric_enu(ra, dec, times, telescope_loc, coord_method, coord_method_params)   # (nt, 3, 3): columns
    return np.ascontiguousarray(enu)


def enu_to_az_za(enu_e, enu_n, orientation="uvbeam"):
    """matvis ``coordinates.enu_to_az_za`` as recalled in SURVEY.md Appendix B.2 (call site
    cpu_simulate.py:957-959): zeta = sqrt(1 - e^2 - n^2) (0 outside the unit disc), az = atan2(e, n),
    za = pi/2 - asin(zeta); "uvbeam": az <- pi/2 - az; wrapped to [0, 2 pi)."""
    e, n = np.asarray(enu_e), np.asarray(enu_n)
    lsqr = e**2 + n**2
    zeta = np.where(lsqr < 1.0, np.sqrt(np.abs(1.0 - lsqr)), 0.0)
    az = np.arctan2(e, n)
    za = 0.5 * np.pi - np.arcsin(zeta)
    if orientation == "uvbeam":
        az = 0.5 * np.pi - az
    elif orientation != "astropy":
        raise ValueError("orientation must be 'astropy' or 'uvbeam'")
    az = np.where(az < 0, az + D2PI, az)
    az = np.where(az >= D2PI, az - D2PI, az)
    return az.astype(e.dtype, copy=False), za.astype(e.dtype, copy=False)
