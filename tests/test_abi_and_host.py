"""CPU tests of the drop-in boundary and the host logic: the C-ABI libraries
load and export every symbol include/*.h declares, struct layouts match, the
engine refuses to run without a CUDA device (no CPU fallback), and the host
time loop (timemanager replay) keeps the reference's schedule."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

import flexpart_b200 as fb
from flexpart_b200 import abi
import cases
from oracle_api import Oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(header, prefix):
    txt = open(os.path.join(ROOT, "include", header)).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(%s[a-z0-9_]+)\s*\(" % prefix, txt)))


def test_engine_library_exports_every_declared_symbol():
    names = _declared("fpb.h", "fpb_")
    assert len(names) >= 20
    lib = C.CDLL(abi.ENGINE_LIB)  # loading needs no GPU
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_host_library_exports_every_declared_symbol():
    names = _declared("fpb_host.h", "fpbh_")
    assert len(names) >= 12
    lib = C.CDLL(abi.HOST_LIB)
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_struct_layouts_match_the_headers():
    L = fb.load_engine_lib()  # raises on a fpb_config size mismatch
    assert L.fpb_abi_version() == abi.ABI_VERSION
    # compile a probe against the real headers and compare every struct size
    src = r'''
#include <stdio.h>
#include "fpb_host.h"
int main(void){printf("%zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(fpb_config), sizeof(fpb_met_ptrs),
 sizeof(fpb_particle_ptrs), sizeof(fpb_step_stats), sizeof(fpbh_releases), sizeof(fpbh_run),
 sizeof(fpbh_run_result), sizeof(fpbh_engine)); return 0;}
'''
    exe = os.path.join(ROOT, "tests", "_abi_probe")
    subprocess.run(["gcc", "-x", "c", "-", "-I", os.path.join(ROOT, "include"), "-o", exe],
                   input=src.encode(), check=True)
    out = subprocess.check_output([exe]).decode().split()
    os.remove(exe)
    expect = [C.sizeof(t) for t in (abi.FpbConfig, abi.FpbMetPtrs, abi.FpbParticlePtrs, abi.FpbStepStats,
                                    abi.FpbhReleases, abi.FpbhRun, abi.FpbhRunResult, abi.FpbhEngine)]
    assert [int(x) for x in out] == expect


def test_engine_fails_loudly_without_a_gpu():
    """No CPU fallback: on a machine without a CUDA device fpb_init must fail
    with a message, never silently compute elsewhere."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    cb = cases.config_small()
    with pytest.raises(fb.FpbError, match="no CUDA device"):
        fb.Engine(cb)


def test_bad_arguments_are_rejected():
    L = fb.load_engine_lib()
    h = C.c_void_p()
    cb = cases.config_small()
    bad = cb.clone(abi_version=99)
    assert L.fpb_init(C.byref(bad.cfg), C.byref(h)) != 0
    assert b"abi_version" in L.fpb_last_error()
    bad = cb.clone(nz=1000)
    assert L.fpb_init(C.byref(bad.cfg), C.byref(h)) != 0
    assert b"nz" in L.fpb_last_error()
    assert L.fpb_step(None, 0, 0, None) != 0
    assert L.fpb_conccalc(None, 0, 1.0) != 0
    assert L.fpb_upload_met(None, 1, None) != 0


def test_readcommand_semantics():
    """src/readcommand.f90:244-272,377-383,627-634."""
    c = fb.make_config(nx=73, ny=37, nz=10, dx=5.0, dy=5.0, height=fb.synth_heights(10), ctl=-5.0, ifine=4).cfg
    assert (c.method, c.mintime, c.turbswitch, c.ifine) == (0, 900, 0, 1)
    assert abs(c.ctl + 0.2) < 1e-7
    c = fb.make_config(nx=73, ny=37, nz=10, dx=5.0, dy=5.0, height=fb.synth_heights(10), ctl=5.0, ifine=4).cfg
    assert (c.method, c.mintime, c.turbswitch, c.ifine) == (1, 1, 1, 4)
    assert abs(c.ctl - 0.2) < 1e-7 and abs(c.fine - 0.25) < 1e-7
    c = fb.make_config(nx=73, ny=37, nz=10, dx=5.0, dy=5.0, height=fb.synth_heights(10), ctl=2.0, ifine=4,
                       cblflag=1, lsynctime=1800).cfg
    assert c.turbswitch == 1 and c.lsynctime == 1200 and abs(c.ctl - 0.2) < 1e-7 and c.ifine == 11
    c = fb.make_config(nx=73, ny=37, nz=10, dx=5.0, dy=5.0, height=fb.synth_heights(10), ldirect=-1).cfg
    assert c.lsynctime == -900 and c.mintime == 900   # mintime is set before the sign flip, readcommand.f90:384 vs :631
    with pytest.raises(fb.FpbError, match="EITHER -1 OR 1"):
        fb.make_config(nx=73, ny=37, nz=10, dx=5.0, dy=5.0, height=fb.synth_heights(10), ldirect=0)


def test_gridcheck_semantics():
    """src/gridcheck_ecmwf.f90:300-366."""
    c = fb.make_config().cfg
    assert (c.xglobal, c.nglobal, c.sglobal) == (1, 1, 1)
    assert c.switchnorthg == 165.0 and c.switchsouthg == 15.0
    assert c.nxmin1 == 360 and c.nymin1 == 180
    assert abs(c.dxconst - 180.0 / (1.0 * 6.371e6 * 3.14159265)) < 1e-12
    c = fb.make_config(nx=101, ny=81, nz=10, dx=0.25, dy=0.25, xlon0=0.0, ylat0=40.0,
                       height=fb.synth_heights(10), outlon0=0.0, outlat0=40.0, numxgrid=10, numygrid=10).cfg
    assert (c.xglobal, c.nglobal, c.sglobal) == (0, 0, 0)
    assert c.switchnorthg == 999999.0 and c.switchsouthg == 999999.0


def test_synthetic_met_invariants():
    """What the hot path divides by or searches on."""
    cb = cases.config_small()
    c = cb.cfg
    m = fb.MetFields(cb).synth(3600)
    nx, ny, nz = c.nx, c.ny, c.nz
    h = cb.height
    assert h[0] == 0.0 and np.all(np.diff(h) > 0)
    assert m.hmix[:nx, :ny].min() >= 100.0 and m.hmix[:nx, :ny].max() <= 4500.0
    assert m.ustar[:nx, :ny].min() >= 1e-8 and m.rho[:nx, :ny, :nz].min() > 0
    assert np.all(np.isfinite(m.oli)) and np.abs(m.oli[:nx, :ny]).min() >= 1e-3 - 1e-9
    # cyclic column repeats the first one
    for f in (m.uu, m.vv, m.rho, m.hmix):
        np.testing.assert_allclose(f[0], f[nx - 1], rtol=0, atol=2e-4)
    # drhodz by the reference's centred differences (verttransform_ecmwf.f90:392-398)
    k = 5
    ref = (m.rho[:nx, :ny, k + 1] - m.rho[:nx, :ny, k - 1]) / (h[k + 1] - h[k - 1])
    np.testing.assert_allclose(m.drhodz[:nx, :ny, k], ref, rtol=1e-6)
    np.testing.assert_array_equal(m.drhodz[:nx, :ny, nz - 1], m.drhodz[:nx, :ny, nz - 2])
    # polar winds exist poleward of the switch rows and are constant on the pole rows
    j0 = int(c.switchnorthg) - 2
    assert np.abs(m.uupol[:nx, j0:ny, 0]).max() > 0
    assert np.ptp(m.uupol[:nx, ny - 1, 3]) == 0 and np.ptp(m.vvpol[:nx, 0, 3]) == 0
    assert np.ptp(m.ww[:nx, ny - 1, 7]) == 0


def test_timemanager_schedule_matches_reference_loop():
    """Sampling weights 0.5 at the window ends, output every LOUTSTEP with a
    second half-weight sample when the next window starts at the same time
    (src/timemanager.f90:350-365, 376-464), exit at ideltas (:509)."""
    cb = cases.config_small(nrel=1, npart_each=50)
    rel = cases.releases_boxes(cb)
    calls = []

    class Rec(Oracle):
        pass
    o = Rec(cb)
    o.fill_rannumb(50000, -320)
    v = o.vtable()
    # (a struct field access returns a view of the slot: copy the raw addresses)
    orig_conc = abi.CONC_FN(C.cast(v.conccalc, C.c_void_p).value)
    orig_step = abi.STEP_FN(C.cast(v.step, C.c_void_p).value)

    def conc(self_, itime, weight):
        calls.append(("conc", itime, weight))
        return orig_conc(self_, itime, weight)

    def step(self_, itime, ldeltat, st):
        calls.append(("step", itime, ldeltat))
        return orig_step(self_, itime, ldeltat, st)
    v.conccalc = abi.CONC_FN(conc)
    v.step = abi.STEP_FN(step)
    run = fb.RunSpec(ideltas=3 * 3600, loutstep=3600, loutaver=3600, loutsample=900)
    res, outs = fb.timemanager(cb, rel, run, v)
    assert res.syncs == 12 and res.outputs == 3
    assert [o_["itime"] for o_ in outs] == [3600, 7200, 10800]
    assert all(o_["outnum"] == 4.0 for o_ in outs)
    conc_calls = [c_ for c_ in calls if c_[0] == "conc"]
    w = {}
    for _, t, wt in conc_calls:
        w.setdefault(t, []).append(wt)
    assert w[0] == [0.5] and w[900] == [1.0] and w[3600] == [0.5, 0.5] and w[10800] == [0.5, 0.5]
    steps = [c_ for c_ in calls if c_[0] == "step"]
    assert [t for _, t, _ in steps] == list(range(0, 10800, 900))
    # ldeltat: time since the last deposition-decay reference (timemanager.f90:514-518)
    ld = {t: l for _, t, l in steps}
    assert ld[0] == 0 - (1800 - 3600) and ld[1800] == 0 and ld[2700] == 900 and ld[5400] == 0
    # mass: every output holds outnum * total released mass
    for o_ in outs:
        assert abs(o_["gridunc"].sum() - 4.0 * 50 * (1.0 / 50)) < 1e-4


def test_timemanager_backward_run_counts_down():
    """LDIRECT=-1: lsynctime < 0, the loop counts down to a negative ideltas and
    the met bracket is ordered in run direction (src/readcommand.f90:627-634)."""
    cb = cases.config_small(nrel=1, npart_each=40, ldirect=-1, ctl=-5.0)
    rel = cases.releases_boxes(cb, zmax=1000.0)
    o = Oracle(cb)
    o.fill_rannumb(50000, -320)
    run = fb.RunSpec(ideltas=-4 * 900, ldirect=-1)
    res, outs = fb.timemanager(cb, rel, run, o.vtable())
    assert res.syncs == 4 and res.particle_steps == 160 and res.outputs == 1
    assert outs[0]["itime"] == -3600
    p = fb.Particles(cb.cfg.maxpart, 1)
    p.numpart = 40
    o.pull_particles(p)
    assert np.all(p.itra1[:40] == -3600)


def test_two_rank_partition_and_grid_reduce_gloo():
    """Multi-GPU semantics on CPU: particles split round-robin over 2 ranks
    (releaseparticles_mpi.f90:141-152), each with a full met replica; the only
    collective is a sum-reduce of gridunc to rank 0 (mpi_mod.f90:2395-2579).
    The reduced grid equals the single-rank grid."""
    script = os.path.join(ROOT, "tests", "_gloo_worker.py")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29531", script],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "GLOO_OK" in out.stdout


def test_timemanager_release_into_reused_slots_keeps_live_particles():
    """Continuous releases with terminations by age: new particles take the dead slots between
    live ones (src/releaseparticles.f90:139-160).  The time loop must hand the engine the new rows
    only -- its host copies of the live rows in between are stale.  Checked against a loop that
    works on the oracle's own state (fpo_releaseparticles + fpo_step, no host mirror)."""
    cb = cases.config_small(nrel=3, npart_each=400, maxpart=1300, lage=(2700,))
    c = cb.cfg
    rel = cases.releases_boxes(cb, seed=11, start=0, end=3600)
    nsteps = 4
    ora = Oracle(cb); ora.fill_rannumb()
    res, _ = fb.timemanager(cb, rel, fb.RunSpec(ideltas=nsteps * 900), ora.vtable())

    ref = Oracle(cb); ref.fill_rannumb()
    _pf, _pi = C.POINTER(C.c_float), C.POINTER(C.c_int32)
    fp = lambda a: a.ctypes.data_as(_pf)
    xmasssave = np.zeros(c.numpoint, np.float32)
    for k in range(nsteps + 1):
        itime = k * 900
        if k % 12 == 0:   # the loop's 3-hourly synthetic fields (fpbh_timemanager)
            ref.upload_met(1, fb.MetFields(cb).synth(itime)); ref.upload_met(2, fb.MetFields(cb).synth(itime + 10800))
            ref.set_met_bracket((1, 2), (itime, itime + 10800))
        assert ref.L.fpo_releaseparticles(ref.S, itime, c.numpoint, rel.start.ctypes.data_as(_pi),
                                          rel.end.ctypes.data_as(_pi), fp(rel.xpoint1), fp(rel.ypoint1), fp(rel.xpoint2),
                                          fp(rel.ypoint2), fp(rel.zpoint1), fp(rel.zpoint2), fp(xmasssave), 99999999) == 0
        if k == nsteps:
            break
        ref.step(itime, 0)
    n = res.numpart_final
    a, b = fb.Particles(c.maxpart, 1), fb.Particles(c.maxpart, 1)
    a.numpart = b.numpart = n
    ora.pull_particles(a); ref.pull_particles(b)
    assert np.array_equal(a.itra1[:n], b.itra1[:n])
    live = b.itra1[:n] != fb.ITRA_DEAD
    assert (~live).any() and (b.itramem[:n][live] < nsteps * 900).any()   # dead slots, and older particles still around
    for f in ("xtra1", "ytra1", "ztra1", "idt", "npoint", "itramem"):
        assert np.array_equal(getattr(a, f)[:n][live], getattr(b, f)[:n][live]), f
    stepped = live & (b.itramem[:n] < nsteps * 900)   # (initialize sets the velocities of the newest ones)
    for f in ("uap", "ucp", "uzp"):
        assert np.array_equal(getattr(a, f)[:n][stepped], getattr(b, f)[:n][stepped]), f


# ---------------------------------------------------------------------------------------------
# include/fpb_mod.f90 (the ISO_C_BINDING module of INTEGRATION.md) against include/fpb.h
# ---------------------------------------------------------------------------------------------
_F_KIND = {"integer(c_int32_t)": 4, "integer(c_int64_t)": 8, "integer(c_int16_t)": 2, "integer(c_int8_t)": 1,
           "integer(c_int)": 4, "integer(c_size_t)": 8, "real(c_float)": 4, "real(c_double)": 8, "type(c_ptr)": 8}


def _fortran_module():
    """parse include/fpb_mod.f90 on its own: parameters, derived types (field, kind, extent), interfaces"""
    txt = open(os.path.join(ROOT, "include", "fpb_mod.f90")).read()
    txt = re.sub(r"&\s*\n\s*", " ", txt)
    params = {m.group(1): int(m.group(2)) for m in re.finditer(r"parameter\s*::\s*(\w+)\s*=\s*(-?\d+)", txt)}
    types = {}
    for m in re.finditer(r"type, bind\(C\) :: (\w+)\n(.*?)end type", txt, flags=re.S):
        fields = []
        for line in m.group(2).strip().splitlines():
            mm = re.match(r"\s*(\S+(?:\(\w+\))?)\s*::\s*(\w+)(?:\((\w+)\))?\s*$", line)
            assert mm, line
            ext = mm.group(3)
            fields.append((mm.group(2), mm.group(1), (params[ext] if ext in params else int(ext)) if ext else 1))
        types[m.group(1)] = fields
    funcs = set(re.findall(r"function (fpb_\w+)\(", txt))
    return params, types, funcs


def test_timemanager_domainfill_boundary_and_splitting_on_the_oracle():
    """mdomainfill = 1 in the host loop (src/timemanager.f90:230-241,472-503): init_domainfill at itime 0,
    boundcond_domainfill afterwards, splitting at loutend once itsplit is reached -- here with the oracle
    behind the engine table (tests/test_gpu_timeloop.py runs the CUDA engine through the same loop)."""
    from oracle_api import Oracle
    cb = cases.config_small(nrel=1, npart_each=20000, maxpart=60000, mdomainfill=1, nclassunc=2)
    rel = fb.Releases(cb, lon1=[-60.0], lon2=[70.0], lat1=[-30.0], lat2=[45.0], z1=[0.0], z2=[100.0], start=[0],
                      end=[0], itsplit=5400)
    ora = Oracle(cb)
    ora.fill_rannumb()
    r, outs = fb.timemanager(cb, rel, fb.RunSpec(ideltas=10 * 900), ora.vtable())
    assert r.syncs == 10 and len(outs) == 2 and r.split_calls == 1 and r.boundary_particles > 0
    n0 = 19984                      # init_domainfill: nint(0.999 * npart * share) summed over the columns
    assert abs(r.numpart_final - 2 * n0) < 200
    p = fb.Particles(cb.cfg.maxpart, 1); p.numpart = r.numpart_final
    ora.pull_particles(p)
    live = p.itra1[:p.numpart] != fb.ITRA_DEAD
    # the halves of a split particle carry half the mass each: the air mass of the box stays what
    # init_domainfill distributed (up to what left and entered through the boundaries in 10 steps)
    o2 = Oracle(cb)
    m0, m1 = cases.met_pair(cb)
    o2.upload_met(1, m0); o2.upload_met(2, m1)
    c = cb.cfg
    pts = [float(np.float32(v)) for v in ((-60.0 - c.xlon0) / c.dx, (-30.0 - c.ylat0) / c.dy, (70.0 - c.xlon0) / c.dx,
                                          (45.0 - c.ylat0) / c.dy)]
    _, info = o2.init_domainfill(pts)
    m = p.xmass1[:p.numpart, 0][live].astype(np.float64).sum()
    assert live.sum() > 30000 and abs(m / info["colmasstotal"] - 1.0) < 0.1, m / info["colmasstotal"]
    # a run without the domain-filling entry points in the table is refused, not silently different
    v = ora.vtable()
    v.init_domainfill = abi.INIT_DF_FN()
    with pytest.raises(fb.FpbError, match="init_domainfill"):
        fb.timemanager(cb, rel, fb.RunSpec(ideltas=900), v)


def test_fortran_module_is_generated_from_the_header():
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import gen_fortran_module as g
    assert open(g.OUT).read() == g.generate(), "include/fpb_mod.f90 is stale: python tools/gen_fortran_module.py"


def test_fortran_module_layout_matches_the_c_compiler():
    """Field order, kinds and array extents of every derived type, laid out by the C rules a bind(C)
    type follows, against gcc's own offsetof/sizeof of the header's structs; the parameters against
    the header's constants; an interface for every exported symbol."""
    params, types, funcs = _fortran_module()
    assert set(types) == {"fpb_config", "fpb_met_ptrs", "fpb_particle_ptrs", "fpb_step_stats", "fpb_partout_ptrs",
                          "fpb_release_points", "fpb_domainfill_info", "fpb_conv_ptrs", "fpb_rawmet_ptrs", "fpb_partav_ptrs",
                          "fpb_met_out_ptrs"}
    lines = []
    for t, fields in types.items():
        for f, _, _ in fields:
            lines.append(f'printf("{t}.{f} %zu\\n", offsetof({t}, {f}));')
        lines.append(f'printf("{t} %zu\\n", sizeof({t}));')
    for k in params:
        lines.append(f'printf("{k} %d\\n", (int){k});')
    src = "#include <stdio.h>\n#include <stddef.h>\n#include \"fpb.h\"\nint main(void){\n" + "\n".join(lines) + "\nreturn 0;}\n"
    exe = os.path.join(ROOT, "tests", "_fmod_probe")
    subprocess.run(["gcc", "-x", "c", "-", "-I", os.path.join(ROOT, "include"), "-o", exe], input=src.encode(), check=True)
    got = dict(l.split() for l in subprocess.check_output([exe]).decode().splitlines())
    os.remove(exe)
    for k, v in params.items():
        assert int(got[k]) == v, k
    for t, fields in types.items():
        off, align = 0, 1
        for f, kind, n in fields:
            sz = _F_KIND[kind]
            off = (off + sz - 1) // sz * sz
            assert int(got[f"{t}.{f}"]) == off, (t, f, got[f"{t}.{f}"], off)
            off += sz * n
            align = max(align, sz)
        assert int(got[t]) == (off + align - 1) // align * align, t
    # every field of the C structs is there (the probe above would not notice a field missing at the end)
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import gen_fortran_module as g
    _, structs, _ = g.parse()
    for name, cf in structs:
        assert [f[2] for f in cf] == [f[0] for f in types[name]], name
    assert funcs == set(_declared("fpb.h", "fpb_")), funcs ^ set(_declared("fpb.h", "fpb_"))
