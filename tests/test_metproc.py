"""calcpar + verttransform_ecmwf: the device code of flexpart_b200/csrc/
fpb_metproc.cuh, compiled for the host by tests/met_host_check.cpp, against the reference's own
routines (src/calcpar.f90, scalev.f90, obukhov.f90, richardson.f90, verttransform_ecmwf.f90 run from
their sources through oracle/f2c).  The GPU twin of this test (tests/test_gpu_metproc.py) runs the
same comparison through the C ABI."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import flexpart_b200 as fb
import cases
import conv_cases
import met_cases
import ref_api
from metproc_common import reference_run, compare_fields

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.skipif(not ref_api.available(), reason="oracle/_ref/libflexref.so not built")

_fp = lambda a: a.ctypes.data_as(C.c_void_p)


class MetGrid(C.Structure):
    _fields_ = [("nx", C.c_int), ("ny", C.c_int), ("nz", C.c_int), ("nuvz", C.c_int), ("nwz", C.c_int),
                ("nxd", C.c_int), ("nyd", C.c_int),
                ("dx", C.c_float), ("dy", C.c_float), ("xlon0", C.c_float), ("ylat0", C.c_float),
                ("dxconst", C.c_float), ("dyconst", C.c_float),
                ("nglobal", C.c_int), ("sglobal", C.c_int), ("xglobal", C.c_int),
                ("switchnorthg", C.c_float), ("switchsouthg", C.c_float),
                ("northpolemap", C.c_float * 9), ("southpolemap", C.c_float * 9),
                ("lsubgrid", C.c_int), ("readclouds", C.c_int), ("nest", C.c_int), ("xresol", C.c_float), ("yresol", C.c_float)] + \
               [(n, C.c_void_p) for n in ("akz", "bkz", "akm", "bkm", "height", "cosf", "UV", "W", "TQ", "PV", "theta", "SF1",
                                          "SF2", "excessoro", "CLW", "CIW", "clw", "uvzlev", "A", "G", "T", "P", "S", "trop", "R", "Cl", "Q")]


def _host_lib():
    out = os.path.join(ROOT, "tests", "_build")
    os.makedirs(out, exist_ok=True)
    so = os.path.join(out, "libmetcheck.so")
    src = os.path.join(ROOT, "tests", "met_host_check.cpp")
    hdr = os.path.join(ROOT, "flexpart_b200", "csrc", "fpb_metproc.cuh")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        inc = "/usr/local/cuda/include"
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-ffp-contract=off", "-shared", "-I" + inc,
                               "-o", so, src])
    L = C.CDLL(so)
    L.met_check_sizeof_grid.restype = C.c_ulong
    assert L.met_check_sizeof_grid() == C.sizeof(MetGrid)
    return L


def device_layout_inputs(cb, raw, pvh, akm, bkm, akz, bkz, nuvz, height):
    """arrays of MetGrid in [k][jy][ix] order (what fpb_calcpar_verttransform builds on the device)"""
    c = cb.cfg
    nx, ny = c.nx, c.ny
    lev = lambda a, n: np.ascontiguousarray(np.transpose(a[:nx, :ny, :n], (2, 1, 0)))
    k = {}
    k["UV"] = np.ascontiguousarray(np.stack([lev(raw["uuh"], nuvz), lev(raw["vvh"], nuvz)], axis=-1))
    k["W"] = lev(raw["wwh"], nuvz)
    k["TQ"] = np.ascontiguousarray(np.stack([lev(raw["tth"], nuvz), lev(raw["qvh"], nuvz)], axis=-1))
    k["PV"] = np.zeros_like(k["W"])                      # computed by the calcpv pass
    k["theta"] = np.zeros_like(k["W"])
    s2 = lambda a: np.ascontiguousarray(a[:nx, :ny].T)
    k["SF1"] = np.ascontiguousarray(np.stack([s2(raw[n]) for n in ("ps", "tt2", "td2", "sshf")], axis=-1))
    k["SF2"] = np.ascontiguousarray(np.stack([s2(raw[n]) for n in ("surfstr", "lsprec", "convprec", "tcc")], axis=-1))
    for nm, a in (("akz", akz), ("bkz", bkz), ("akm", akm), ("bkm", bkm)):
        k[nm] = np.ascontiguousarray(a[:nuvz + 1], np.float32)
    k["height"] = np.ascontiguousarray(height, np.float32)
    lat = (np.arange(ny, dtype=np.float32) * np.float32(c.dy) + np.float32(c.ylat0)) * np.float32(np.float32(3.14159265) / np.float32(180.0))
    k["cosf"] = (np.float32(1.0) / np.cos(lat.astype(np.float64)).astype(np.float32)).astype(np.float32)
    return k


def make_grid(cb, k, nuvz, lsubgrid=0, readclouds=0):
    c = cb.cfg
    nx, ny = c.nx, c.ny
    g = MetGrid()
    g.nx, g.ny, g.nz, g.nuvz, g.nwz, g.nxd, g.nyd = nx, ny, nuvz, nuvz, nuvz, nx, ny
    for f in ("dx", "dy", "xlon0", "ylat0", "dxconst", "dyconst", "nglobal", "sglobal", "xglobal", "switchnorthg",
              "switchsouthg"):
        setattr(g, f, getattr(c, f))
    g.northpolemap[:] = list(c.northpolemap); g.southpolemap[:] = list(c.southpolemap)
    g.lsubgrid, g.readclouds = lsubgrid, readclouds
    out = dict(uvzlev=np.zeros((nuvz, ny, nx), np.float32), A=np.zeros((nuvz, ny, nx, 4), np.float32),
               G=np.zeros((nuvz, ny, nx), np.float32), T=np.zeros((nuvz, ny, nx), np.float32),
               P=np.zeros((nuvz, ny, nx, 2), np.float32), S=np.zeros((ny, nx, 4), np.float32),
               trop=np.zeros((ny, nx), np.float32), R=np.zeros((ny, nx, 4), np.float32),
               Cl=np.zeros((nuvz, ny, nx), np.int8), Q=np.zeros((nuvz, ny, nx, 2), np.float32))
    for nm, a in list(k.items()) + list(out.items()):
        setattr(g, nm, _fp(a))
    return g, out


def compare(cb, ref, out, nuvz):
    got = {"uu": out["A"][..., 0], "vv": out["A"][..., 1], "ww": out["A"][..., 2], "rho": out["A"][..., 3],
           "drhodz": out["G"], "tt": out["T"], "pv": out["Q"][..., 0], "qv": out["Q"][..., 1],
           "hmix": out["S"][..., 0], "ustar": out["S"][..., 1], "wstar": out["S"][..., 2], "oli": out["S"][..., 3],
           "tropopause": out["trop"], "clouds": out["Cl"], "uupol": out["P"][..., 0], "vvpol": out["P"][..., 1]}
    return compare_fields(cb, ref, got, nuvz)


def test_nested_grid_host_build_matches_reference():
    """calcpar_nests + verttransform_nests + calcpv_nests (src/getfields.f90:131-134): the column code with the
    nest's geometry, no poles, no wrap, xresoln / yresoln in the slope term"""
    from metproc_common import nest_configs, reference_run_nest, compare_fields_nest, NEST
    nuvz = 40
    kw = dict(nrel=1, npart_each=8, nz=nuvz, wetdepspec=(1,), weta_gas=(2.0e-5,), wetb_gas=(0.62,))
    cb0, _ = nest_configs(**kw, height=fb.synth_heights(nuvz))
    akm, bkm, akz, bkz, _ = conv_cases.hybrid_levels(nuvz)
    raw = met_cases.raw_fields(cb0, akz, bkz, nuvz, seed=1)
    ref, height, _ = reference_run(cb0, raw, akm, bkm, akz, bkz, nuvz)
    cb, cbn = nest_configs(**kw, height=height)
    rawn = met_cases.raw_fields(cbn, akz, bkz, nuvz, seed=4)
    pvhn = reference_run_nest(ref, cb, rawn, nuvz)
    k = device_layout_inputs(cbn, rawn, pvhn, akm, bkm, akz, bkz, nuvz, height)
    g, out = make_grid(cbn, k, nuvz)
    g.nglobal = g.sglobal = g.xglobal = 0
    g.nest, g.xresol, g.yresol = 1, cb.cfg.xresoln[0], cb.cfg.yresoln[0]
    g.dxconst, g.dyconst = cb.cfg.dxconst, cb.cfg.dyconst       # the mother grid's (src/verttransform_nests.f90:322)
    _host_lib().met_check_run(C.byref(g))
    got = {"uu": out["A"][..., 0], "vv": out["A"][..., 1], "ww": out["A"][..., 2], "rho": out["A"][..., 3],
           "drhodz": out["G"], "tt": out["T"], "hmix": out["S"][..., 0], "ustar": out["S"][..., 1],
           "wstar": out["S"][..., 2], "oli": out["S"][..., 3], "tropopause": out["trop"], "clouds": out["Cl"],
           "pv": out["Q"][..., 0], "qv": out["Q"][..., 1]}
    bad = compare_fields_nest(ref, got, nuvz)
    assert not bad, bad
    assert len(set(np.unique(out["Cl"]))) >= 4 and np.abs(out["A"][..., 2]).max() > 0.01


@pytest.mark.parametrize("sumclouds", [0, 1])
def test_readclouds_branch_matches_reference(sumclouds):
    """cloud water from the input (src/verttransform_ecmwf.f90:610-681): clwc (+ ciwc) on the height levels,
    the column total ctwc and the in-cloud / below-cloud classes"""
    nuvz = 40
    cb = cases.config_small(nrel=1, npart_each=8, nz=nuvz, wetdepspec=(1,), weta_gas=(2.0e-5,), wetb_gas=(0.62,),
                            readclouds=1)
    c = cb.cfg
    akm, bkm, akz, bkz, _ = conv_cases.hybrid_levels(nuvz)
    raw = met_cases.raw_fields(cb, akz, bkz, nuvz, seed=3)
    ref, height, pvh = reference_run(cb, raw, akm, bkm, akz, bkz, nuvz, readclouds=1, sumclouds=sumclouds)
    k = device_layout_inputs(cb, raw, pvh, akm, bkm, akz, bkz, nuvz, height)
    lev = lambda a: np.ascontiguousarray(np.transpose(a[:c.nx, :c.ny, :nuvz], (2, 1, 0)))
    k["CLW"] = lev(raw["clwch"] + (raw["ciwch"] if sumclouds else 0.0))
    if not sumclouds:
        k["CIW"] = lev(raw["ciwch"])
    k["clw"] = np.zeros_like(k["W"])
    g, out = make_grid(cb, k, nuvz, readclouds=1)
    _host_lib().met_check_run(C.byref(g))
    got = {"clouds": out["Cl"], "ctwc": out["R"][..., 3], "rho": out["A"][..., 3], "qv": out["Q"][..., 1]}
    bad = compare_fields(cb, ref, got, nuvz)
    assert not bad, bad
    assert set(np.unique(out["Cl"])) == {0, 2, 3, 4, 5} and (out["R"][..., 3] > 0).mean() > 0.5


@pytest.mark.parametrize("seed", [1, 2])
def test_calcpar_verttransform_host_build_matches_reference(seed):
    nuvz = 40
    cb = cases.config_small(nrel=1, npart_each=8, nz=nuvz, wetdepspec=(1,), weta_gas=(2.0e-5,), wetb_gas=(0.62,))
    c = cb.cfg
    akm, bkm, akz, bkz, _ = conv_cases.hybrid_levels(nuvz)
    raw = met_cases.raw_fields(cb, akz, bkz, nuvz, seed=seed)
    ref, height, pvh = reference_run(cb, raw, akm, bkm, akz, bkz, nuvz)
    assert height[0] == 0.0 and (np.diff(height) > 0).all() and height[-1] > 20000.0
    # the host helper a device-side caller derives fpb_config::height with
    hh, (ixm, jym) = fb.verttransform_heights(cb, nuvz, akz[1:], bkz[1:], raw)
    assert np.array_equal(hh.view(np.uint32), height.view(np.uint32)) and raw["ps"][ixm, jym] > 100000.0
    k = device_layout_inputs(cb, raw, pvh, akm, bkm, akz, bkz, nuvz, height)
    g, out = make_grid(cb, k, nuvz)
    _host_lib().met_check_run(C.byref(g))
    bad = compare(cb, ref, out, nuvz)
    assert not bad, bad
    # calcpv: the potential vorticity on the eta levels (the reference's pvh after calcpar)
    pv_ref = np.transpose(pvh[:c.nx, :c.ny, :nuvz], (2, 1, 0))
    assert np.abs(pv_ref).max() > 0.1
    assert np.array_equal(k["PV"].view(np.uint32), np.ascontiguousarray(pv_ref).view(np.uint32)), \
        (int((k["PV"] != pv_ref).sum()), float(np.abs(k["PV"] - pv_ref).max()))
    # the case exercises what it should: all cloud classes, both stabilities, columns whose top lies
    # below the reference column's, a tropopause everywhere
    cl = out["Cl"]
    assert set(np.unique(cl)) == {0, 1, 2, 3, 4, 5}
    assert (out["S"][..., 2] > 0).any() and (out["S"][..., 2] == 0).any()
    assert (out["uvzlev"][-1] < height[-2]).any()
    assert (out["trop"][1:-1] > 2500.0).all()
