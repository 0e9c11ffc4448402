"""Shared pieces of tests/test_metproc.py (CPU) and tests/test_gpu_metproc.py: running the reference's
calcpar + verttransform_ecmwf (oracle/_ref) on a synthetic wind field and comparing a set of
transformed fields with its arrays.  Test infrastructure."""
import ctypes as C

import numpy as np

import ref_api

def reference_run(cb, raw, akm, bkm, akz, bkz, nuvz, lsubgrid=0, excessoro=None, timing=None, readclouds=0, sumclouds=0):
    """calcpar + verttransform_ecmwf of the reference on time slot 1; returns (ref, height)"""
    c = cb.cfg
    ref = ref_api.Ref(cb, maxrand=1000)
    for k, v in dict(nuvz=nuvz, nwz=nuvz, nz=nuvz, nmixz=0, lsubgrid=lsubgrid, readclouds=readclouds, sumclouds=sumclouds).items():
        ref.set(k, v)
    for nm, a in (("akz", akz), ("bkz", bkz), ("akm", akm), ("bkm", bkm), ("aknew", akz), ("bknew", bkz)):
        ref.arr(nm)[:nuvz] = a[1:nuvz + 1]
    for nm in ("ps", "tt2", "td2", "sshf", "surfstr", "lsprec", "convprec", "tcc"):
        ref.arr(nm)[:, :, 0, 0] = raw[nm]
    ref.arr("tth")[:, :, :, 0] = raw["tth"]
    ref.arr("qvh")[:, :, :, 0] = raw["qvh"]
    if readclouds:
        ref.arr("clwch")[:, :, :, 0] = raw["clwch"] + (raw["ciwch"] if sumclouds else 0.0)
        if not sumclouds:
            ref.arr("ciwch")[:, :, :, 0] = raw["ciwch"]
    if excessoro is not None:
        ref.arr("excessoro")[:, :] = excessoro
    n, fmt = C.c_int(1), C.c_int(2)   # GRIBFILE_CENTRE_ECMWF
    pvh = np.zeros_like(raw["uuh"])
    P = lambda a: a.ctypes.data_as(C.POINTER(C.c_float))
    import time
    t0 = time.perf_counter()
    ref.L.f_calcpar(C.byref(n), P(raw["uuh"]), P(raw["vvh"]), P(pvh), C.byref(fmt))
    t1 = time.perf_counter()
    ref.L.f_verttransform_ecmwf(C.byref(n), P(raw["uuh"]), P(raw["vvh"]), P(raw["wwh"]), P(pvh))
    if timing is not None:
        timing.update(calcpar_s=t1 - t0, verttransform_s=time.perf_counter() - t1)
    return ref, ref.arr("height")[:nuvz].copy(), pvh



NEST = dict(xlon0n=-20.0, ylat0n=25.0, nxn=61, nyn=41, dxn=1.0, dyn=1.0)   # a 1 deg nest inside the 5 deg mother grid


def nest_configs(**kw):
    """(cb, cbn): the run configuration with one nested input grid, and a configuration whose MOTHER grid is
    that nest's grid (for the field synthesizer only)"""
    import cases
    g = NEST
    cb = cases.config_small(met_nests=((g["xlon0n"], g["ylat0n"], g["nxn"], g["nyn"], g["dxn"], g["dyn"]),), **kw)
    kwn = {k: v for k, v in kw.items() if k not in ("nx", "ny")}
    kwn["height"] = cb.height      # (config_small's default ladder otherwise)
    cbn = cases.config_small(nx=g["nxn"], ny=g["nyn"], dx=g["dxn"], dy=g["dyn"], xlon0=g["xlon0n"], ylat0=g["ylat0n"],
                             outlon0=g["xlon0n"], outlat0=g["ylat0n"], numxgrid=10, numygrid=10, dxout=1.0, dyout=1.0, **kwn)
    assert cb.cfg.nxmaxn == g["nxn"] and cb.cfg.nymaxn == g["nyn"]
    return cb, cbn


def reference_run_nest(ref, cb, rawn, nuvz, lsubgrid=0, excessoron=None):
    """calcpar_nests + verttransform_nests (+ calcpv_nests) of the reference for nest 1 on time slot 1, on a Ref
    that has already run the mother grid (height, akz .. are set); returns pvhn"""
    g = NEST
    for k, v in dict(numbnests=1, lsubgrid=lsubgrid).items():
        ref.set(k, v)
    ref.arr("nxn")[0] = g["nxn"]; ref.arr("nyn")[0] = g["nyn"]
    ref.arr("dxn")[0] = g["dxn"]; ref.arr("dyn")[0] = g["dyn"]
    for nm in ("xlon0n", "ylat0n"):
        if ref.has(nm):
            ref.arr(nm)[0] = g[nm]
    ref.arr("xresoln")[1] = cb.cfg.xresoln[0]; ref.arr("yresoln")[1] = cb.cfg.yresoln[0]   # xresoln(0:maxnests)
    ref.arr("readclouds_nest")[0] = 0
    for nm in ("ps", "tt2", "td2", "sshf", "surfstr", "lsprec", "convprec"):
        ref.arr(nm + "n")[:, :, 0, 0, 0] = rawn[nm]
    ref.arr("tthn")[:, :, :, 0, 0] = rawn["tth"]
    ref.arr("qvhn")[:, :, :, 0, 0] = rawn["qvh"]
    if excessoron is not None:
        ref.arr("excessoron")[:, :, 0] = excessoron
    n, fmt = C.c_int(1), C.c_int(2)
    pvhn = np.zeros_like(rawn["uuh"])
    P = lambda a: a.ctypes.data_as(C.POINTER(C.c_float))
    ref.L.f_calcpar_nests(C.byref(n), P(rawn["uuh"]), P(rawn["vvh"]), P(pvhn), C.byref(fmt))
    ref.L.f_verttransform_nests(C.byref(n), P(rawn["uuh"]), P(rawn["vvh"]), P(rawn["wwh"]), P(pvhn))
    return pvhn


def compare_fields_nest(ref, got, nuvz):
    """got: dict name -> [k][jy][ix] (or [jy][ix]) over the nest; against uun .. of nest 1, slot 1 (bitwise)"""
    g = NEST
    nx, ny = g["nxn"], g["nyn"]
    bad = {}
    for nm, a in got.items():
        if a.ndim == 3:
            b = np.transpose(ref.arr(nm + "n")[:nx, :ny, :nuvz, 0, 0], (2, 1, 0))
        else:
            b = ref.arr(nm + "n")[:nx, :ny, 0, 0, 0].T
        if nm == "clouds":
            if not np.array_equal(a, b):
                bad[nm] = int((a != b).sum())
            continue
        assert np.isfinite(b).all(), nm
        if not np.array_equal(np.ascontiguousarray(a).view(np.uint32), np.ascontiguousarray(b).view(np.uint32)):
            bad[nm] = (int((a != b).sum()), float(np.abs(a - b).max()))
    return bad


def compare_fields(cb, ref, got, nuvz, exact=True, tol=0.0):
    """got: dict name -> array [k][jy][ix] (or [jy][ix]) over the used grid; against the reference's
    slot-1 arrays.  Returns the fields that differ."""
    c = cb.cfg
    nx, ny = c.nx, c.ny
    r3 = lambda nm: np.transpose(ref.arr(nm)[:nx, :ny, :nuvz, 0], (2, 1, 0))
    r2 = lambda nm: ref.arr(nm)[:nx, :ny, 0, 0].T
    bad = {}
    for nm, a in got.items():
        if nm in ("uupol", "vvpol"):
            continue
        if nm == "ctwc":
            b = ref.arr("ctwc")[:nx, :ny, 0].T
            if not np.array_equal(np.ascontiguousarray(a).view(np.uint32), np.ascontiguousarray(b).view(np.uint32)):
                bad[nm] = (int((a != b).sum()), float(np.abs(a - b).max()))
            continue
        b = r3(nm) if a.ndim == 3 else r2(nm)
        if nm == "clouds":
            if not np.array_equal(a, b):
                bad[nm] = int((a != b).sum())
            continue
        assert np.isfinite(b).all(), nm
        if exact:
            if not np.array_equal(np.ascontiguousarray(a).view(np.uint32), np.ascontiguousarray(b).view(np.uint32)):
                bad[nm] = (int((a != b).sum()), float(np.abs(a - b).max()))
        else:
            d = np.abs(a - b) / np.maximum(np.abs(b), 1e-20)
            if d.max() > tol:
                bad[nm] = float(d.max())
    # polar stereographic winds: only defined poleward of the switch latitudes; cc2gll works in double
    # (sin / cos of the device's and glibc's libm agree to an ulp of double: a float ulp at most)
    jn, js = int(c.switchnorthg) - 2, int(c.switchsouthg) + 3
    for nm in ("uupol", "vvpol"):
        if nm not in got:
            continue
        b = r3(nm)
        a = got[nm]
        for sl in (slice(jn, ny), slice(0, js + 1)):
            if not np.allclose(a[:, sl], b[:, sl], rtol=2e-6, atol=1e-6):
                bad[nm] = float(np.abs(a[:, sl] - b[:, sl]).max())
    return bad
