"""Synthetic inputs of the convection tests: L137-like hybrid coefficients and moist soundings that
make the Emanuel scheme fire in part of the columns (test infrastructure)."""
import numpy as np


def hybrid_levels(nuvz=138):
    """akm, bkm (half levels, index 1 = surface) and akz, bkz (layer centres; index 1 = surface) as
    FLEXPART holds them (src/gridcheck_ecmwf.f90:470-530), 1-based arrays of length nuvz + 1; nconvlev
    as derived at src/gridcheck_ecmwf.f90:553-566."""
    k = np.arange(1, nuvz + 1)
    eta = ((nuvz - k) / (nuvz - 1.0)) ** 1.35
    b = eta ** 2.2
    a = 101325.0 * (eta - b) + 1.0 * (1 - eta)       # ~1 Pa at the top: no log(0)
    akm, bkm = np.zeros(nuvz + 1, np.float32), np.zeros(nuvz + 1, np.float32)
    akm[1:], bkm[1:] = a, b
    akz, bkz = np.zeros(nuvz + 1, np.float32), np.zeros(nuvz + 1, np.float32)
    akz[1], bkz[1] = 0.0, 1.0
    akz[2:] = 0.5 * (akm[1:-1] + akm[2:])
    bkz[2:] = 0.5 * (bkm[1:-1] + bkm[2:])
    nconvlev = nuvz - 2
    for i in range(1, nuvz - 1):
        if akz[i] + bkz[i] * np.float32(101325.0) < 5000.0:
            nconvlev = i
            break
    nconvlev = min(nconvlev, nuvz - 1 - 1)           # nconvlevmax - 1 with nuvzmax = nuvz
    return akm, bkm, akz, bkz, nconvlev


def sounding(rs, akz, bkz, nuvz):
    """tconv, qconv (1-based, levels 1..nuvz-1 = FLEXPART's tth/qvh(kz+1)), psconv, tt2conv, td2conv"""
    ps = np.float32(rs.uniform(96000.0, 103000.0))
    t0 = rs.uniform(285.0, 305.0)
    rh0 = rs.uniform(0.55, 0.98)
    p = (akz[2:nuvz + 1] + bkz[2:nuvz + 1] * ps).astype(np.float64)       # pconv(1..nuvz-1)
    t = np.maximum(t0 * (p / ps) ** 0.19, rs.uniform(200.0, 215.0))
    t = t + rs.normal(0.0, 0.3, t.shape)
    es = 611.2 * np.exp(17.67 * (t - 273.15) / (t - 29.65))
    qs = 0.622 * es / np.maximum(p - 0.378 * es, 1.0)
    rh = np.clip(rh0 * (p / ps) ** rs.uniform(0.5, 2.0), 0.02, 1.0)
    q = np.clip(rh * qs, 1e-7, 0.03)
    tconv, qconv = np.zeros(nuvz + 2, np.float32), np.zeros(nuvz + 2, np.float32)
    tconv[1:nuvz], qconv[1:nuvz] = t, q
    tt2 = np.float32(t0 + rs.uniform(-1.0, 2.0))
    td2 = np.float32(tt2 - rs.uniform(0.5, 8.0))
    return tconv, qconv, ps, tt2, td2


def conv_fields(cb, akz, bkz, nuvz, seed, tshift=0.0, nest=0):
    """ps, tt2, td2 (nxmax,nymax) and tth, qvh (nxmax,nymax,nuvzmax = nzmax) of one time level: a smooth
    field of soundings (warm and moist in the tropics so that part of the columns convects); nest >= 1:
    the same on nested input grid `nest` ((nxmaxn,nymaxn) extents, its own coordinates)"""
    c = cb.cfg
    rs = np.random.RandomState(seed)
    nzm = c.nzmax
    if nest:
        l = nest - 1
        nxm, nym = c.nxmaxn, c.nymaxn
        lon = (c.xln[l] + np.arange(nxm)[:, None] / c.xresoln[l]) * c.dx + c.xlon0
        lat = (c.yln[l] + np.arange(nym)[None, :] / c.yresoln[l]) * c.dy + c.ylat0
        tshift = tshift + 0.7 * nest          # (a nest that differs from the mother grid shows a wrong pick)
    else:
        nxm, nym = c.nxmax, c.nymax
        lon = np.arange(nxm)[:, None] * c.dx + c.xlon0
        lat = np.arange(nym)[None, :] * c.dy + c.ylat0
    ps = (100000.0 + 1500.0 * np.sin(np.deg2rad(2 * lon)) * np.cos(np.deg2rad(lat))).astype(np.float32)
    t0 = 272.0 + 32.0 * np.cos(np.deg2rad(lat)) ** 2 + 2.0 * np.sin(np.deg2rad(3 * lon)) + tshift
    rh0 = np.clip(0.55 + 0.42 * np.cos(np.deg2rad(lat)) ** 2 + 0.1 * np.sin(np.deg2rad(5 * lon)), 0.2, 0.98)
    tth = np.zeros((nxm, nym, nzm), np.float32, order="F")
    qvh = np.zeros((nxm, nym, nzm), np.float32, order="F")
    ttrop = 205.0 + 8.0 * np.cos(np.deg2rad(lat)) ** 2
    for k in range(2, nuvz + 1):                                  # Fortran level k (tconv(k-1))
        p = akz[k] + bkz[k] * ps.astype(np.float64)
        t = np.maximum(t0 * (p / ps) ** 0.19, ttrop) + 0.2 * np.sin(0.37 * k + np.deg2rad(lon))
        es = 611.2 * np.exp(17.67 * (t - 273.15) / (t - 29.65))
        qs = 0.622 * es / np.maximum(p - 0.378 * es, 1.0)
        rh = np.clip(rh0 * (p / ps) ** 1.2, 0.02, 1.0)
        tth[:, :, k - 1] = t
        qvh[:, :, k - 1] = np.clip(rh * qs, 1e-7, 0.03)
    tth[:, :, 0] = tth[:, :, 1]; qvh[:, :, 0] = qvh[:, :, 1]
    tt2 = (t0 + 0.5 + 0.0 * lon).astype(np.float32)
    td2 = (tt2 - 2.0 - 4.0 * (1.0 - rh0)).astype(np.float32)
    del rs
    return ps, np.asfortranarray(tt2), np.asfortranarray(td2), tth, qvh
