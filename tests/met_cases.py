"""Synthetic model-level wind fields as readwind_ecmwf leaves them (uuh, vvh, wwh, tth, qvh on eta
levels; ps, 2 m temperature / dew point, surface fluxes, precipitation): the input of calcpar +
verttransform_ecmwf.  Test infrastructure (also used by bench.py's next_rows leg)."""
import ctypes as C

import numpy as np

import conv_cases


def raw_fields(cb, akz, bkz, nuvz, seed=1, tshift=0.0, mountain=True):
    """dict of Fortran-ordered float32 arrays with the reference's padded extents:
    uuh, vvh, tth, qvh (nxmax,nymax,nuvzmax=nzmax), wwh (nxmax,nymax,nwzmax=nzmax),
    ps, tt2, td2, sshf, surfstr, lsprec, convprec, tcc (nxmax,nymax)"""
    c = cb.cfg
    ps, tt2, td2, tth, qvh = conv_cases.conv_fields(cb, akz, bkz, nuvz, seed, tshift)
    nxm, nym, nzm = c.nxmax, c.nymax, c.nzmax
    lon = (np.arange(nxm)[:, None] * c.dx + c.xlon0).astype(np.float64)
    lat = (np.arange(nym)[None, :] * c.dy + c.ylat0).astype(np.float64)
    if mountain:   # two massifs: level heights differ between neighbours, column tops fall below the reference's
        bump = 28000.0 * np.exp(-(((lon - 85.0) / 25.0) ** 2 + ((lat - 33.0) / 12.0) ** 2)) \
            + 20000.0 * np.exp(-(((lon + 70.0) / 12.0) ** 2 + ((lat + 20.0) / 25.0) ** 2))
        ps = (ps - bump).astype(np.float32)
        # (tth/qvh stay as they are: a synthetic atmosphere need not be hydrostatically consistent)
    ps = np.asfortranarray(ps.astype(np.float32))
    # dry air under the clouds over a third of the globe (below-cloud scavenging classes 4 and 5)
    qvh = qvh.copy(order="F")
    dry = (np.sin(np.deg2rad(1.5 * lon + 0.0 * lat)) > 0.5)
    for k in range(1, nuvz + 1):
        eta = (akz[k] + bkz[k] * 101325.0) / 101325.0 if k > 1 else 1.0
        if eta > 0.94:
            qvh[:, :, k - 1] = np.where(dry, 0.45 * qvh[:, :, k - 1], qvh[:, :, k - 1])
    # a very cold dome: its columns end below the upper height levels of the reference column
    # (the `height(iz) > uvzlev(nuvz)` branch of src/verttransform_ecmwf.f90:293)
    tth = tth.copy(order="F")
    dome = 80.0 * np.exp(-(((lon + 120.0) / 14.0) ** 2 + ((lat - 50.0) / 10.0) ** 2))
    tth -= dome[:, :, None].astype(np.float32)
    rs = np.random.RandomState(seed + 100)
    uuh = np.zeros((nxm, nym, nzm), np.float32, order="F")
    vvh = np.zeros((nxm, nym, nzm), np.float32, order="F")
    wwh = np.zeros((nxm, nym, nzm), np.float32, order="F")
    for k in range(1, nuvz + 1):
        eta = (akz[k] + bkz[k] * 101325.0) / 101325.0 if k > 1 else 1.0
        jet = 35.0 * (1.0 - eta) ** 0.7 * np.cos(np.deg2rad(lat)) ** 2
        uuh[:, :, k - 1] = 4.0 + jet * (1.0 + 0.3 * np.sin(np.deg2rad(3 * lon))) + 2.0 * np.sin(0.21 * k + np.deg2rad(lon))
        vvh[:, :, k - 1] = 6.0 * np.sin(np.deg2rad(2 * lon + 40.0 * eta)) * np.cos(np.deg2rad(lat)) + 1.5 * np.cos(0.17 * k)
        wwh[:, :, k - 1] = 0.4 * eta * (1.0 - eta) * 4.0 * np.sin(np.deg2rad(4 * lon)) * np.sin(np.deg2rad(3 * lat)) \
            + 0.02 * np.sin(0.5 * k + np.deg2rad(lat))
    uuh[:, :, 0] = 0.6 * uuh[:, :, 1]      # 10 m wind at level 1
    vvh[:, :, 0] = 0.6 * vvh[:, :, 1]
    sshf = (-160.0 * np.cos(np.deg2rad(lat)) * np.sin(np.deg2rad(lon + 30.0)) + 15.0 + 0.0 * lon).astype(np.float32)
    surfstr = (0.02 + 0.25 * np.abs(np.sin(np.deg2rad(2 * lon)) * np.cos(np.deg2rad(lat)))).astype(np.float32)
    lsprec = np.clip(2.5 * np.sin(np.deg2rad(3 * lon + 20.0)) * np.sin(np.deg2rad(2 * lat + 10.0)) - 0.8, 0.0, None)
    convprec = np.clip(3.0 * np.cos(np.deg2rad(lat)) ** 6 * np.sin(np.deg2rad(5 * lon)) - 0.6, 0.0, None)
    tcc = np.clip(0.5 + 0.5 * np.sin(np.deg2rad(2 * lon)) * np.cos(np.deg2rad(3 * lat)), 0.0, 1.0)
    # cloud liquid / ice water (readclouds): decks between 800 and 400 hPa over two thirds of the globe, ice above
    # 500 hPa; every precipitating column holds cloud water (the reference's below-cloud test reads a scalar
    # that only a column with cloud water sets, src/verttransform_ecmwf.f90:634-665)
    clwch = np.zeros((nxm, nym, nzm), np.float32, order="F")
    ciwch = np.zeros((nxm, nym, nzm), np.float32, order="F")
    cloudy = (np.sin(np.deg2rad(2.5 * lon + 20.0)) * np.cos(np.deg2rad(2.0 * lat)) > -0.4) | (lsprec + convprec > 0.0)
    for k in range(2, nuvz + 1):
        eta = (akz[k] + bkz[k] * 101325.0) / 101325.0
        if 0.4 < eta < 0.8:
            clwch[:, :, k - 1] = np.where(cloudy, 2.0e-4 * np.sin(np.pi * (eta - 0.4) / 0.4) * (1.0 + 0.5 * np.sin(np.deg2rad(4 * lon))), 0.0)
        if 0.25 < eta < 0.5:
            ciwch[:, :, k - 1] = np.where(cloudy, 5.0e-5 * np.sin(np.pi * (eta - 0.25) / 0.25), 0.0)
    del rs
    F = lambda a: np.asfortranarray(np.broadcast_to(a, (nxm, nym)).astype(np.float32))
    return dict(clwch=clwch, ciwch=ciwch, uuh=uuh, vvh=vvh, wwh=wwh, tth=tth, qvh=qvh, ps=ps, tt2=tt2, td2=td2, sshf=F(sshf),
                surfstr=F(surfstr), lsprec=F(lsprec), convprec=F(convprec), tcc=F(tcc))


def reference_heights(cb, raw, akz, bkz, nuvz):
    """height(1:nuvz) as the first call of verttransform_ecmwf builds it (src/verttransform_ecmwf.f90:
    131-163) -- numpy restatement used to SET UP cases (the tests read the reference's own array)"""
    c = cb.cfg
    ps = raw["ps"]
    for jy in range(c.ny):
        hit = np.nonzero(ps[:c.nx, jy] > 100000.0)[0]
        if len(hit):
            return int(hit[0]), jy
    raise ValueError("no grid point with ps > 1000 hPa")
