"""Pins the hand-written oracle against the REFERENCE'S OWN SOURCES.

The reference is Fortran and this image has no Fortran compiler.  oracle/f2c/
f90toc.py transpiles the reference's hot-path files (advance.f90, initialize.f90,
interpol_*.f90 and their _nests twins, hanna*.f90, cbl.f90, windalign.f90,
random_mod.f90, get_settling.f90, conccalc.f90, ...) statement by statement to C,
from where they lie under /root/reference/src, and the result is compiled into
oracle/_ref/libflexref.so (git-ignored, travels to the GPU box).  These tests run
that code and the oracle on the same inputs, call by call in the reference's
particle order, and require bit-identical results (oracle in strict_reference
mode: the reference's module-state leaks between calls included).

What the transpiled code does not fix is the transcendental library: both sides
call oracle/fpo_math.h (correctly rounded exp/log/pow/sin/cos/erf), see DESIGN.md."""
import ctypes as C

import numpy as np
import pytest

import flexpart_b200 as fb
import cases
import ref_api
from oracle_api import Oracle, load

pytestmark = pytest.mark.skipif(not ref_api.available(), reason="oracle/_ref/libflexref.so not built (no /root/reference)")

MAXRAND = 20000


def _pair(cb, mets, nest_mets=(), bracket=(0, 10800)):
    ref = ref_api.Ref(cb, maxrand=MAXRAND)
    ora = Oracle(cb, strict_reference=True)
    ora.fill_rannumb(MAXRAND, -320)
    tab = ref.fill_rannumb(-320)
    # FLEXPART.f90:56-59 through the reference's gasdev1/ran3 == the oracle's table, bit for bit
    assert np.array_equal(tab, ora.rannumb(MAXRAND))
    for e in (ref, ora):
        e.upload_met(1, mets[0]); e.upload_met(2, mets[1])
        for nest, (n0, n1) in enumerate(nest_mets, start=1):
            e.upload_met_nest(1, nest, n0); e.upload_met_nest(2, nest, n1)
        e.set_met_bracket((1, 2), bracket)
    return ref, ora


def _oracle_calls(ora):
    L = load()
    S = ora.S
    cf, ci = C.c_float, C.c_int32

    def initialize(itime, ldt, xt, yt, zt):
        ldt_c, up, vp, wp, us, vs, ws, icbt = ci(ldt), cf(), cf(), cf(), cf(), cf(), cf(), C.c_int16(0)
        L.fpo_initialize(S, itime, C.byref(ldt_c), C.byref(up), C.byref(vp), C.byref(wp), C.byref(us), C.byref(vs),
                         C.byref(ws), C.c_double(xt), C.c_double(yt), cf(zt), C.byref(icbt))
        return ldt_c.value, up.value, vp.value, wp.value, us.value, vs.value, ws.value, icbt.value

    def advance(itime, nrelpoint, ldt, up, vp, wp, us, vs, ws, xt, yt, zt, icbt, maxspec):
        a = dict(ldt=ci(ldt), up=cf(up), vp=cf(vp), wp=cf(wp), us=cf(us), vs=cf(vs), ws=cf(ws), nstop=C.c_int(0),
                 xt=C.c_double(xt), yt=C.c_double(yt), zt=cf(zt), icbt=C.c_int16(icbt))
        prob = (cf * 8)()
        L.fpo_advance(S, itime, nrelpoint, C.byref(a["ldt"]), C.byref(a["up"]), C.byref(a["vp"]), C.byref(a["wp"]),
                      C.byref(a["us"]), C.byref(a["vs"]), C.byref(a["ws"]), C.byref(a["nstop"]), C.byref(a["xt"]),
                      C.byref(a["yt"]), C.byref(a["zt"]), prob, C.byref(a["icbt"]))
        out = {k: v.value for k, v in a.items()}
        out["prob"] = np.array(prob[:maxspec], np.float32)
        return out

    L.fpo_initialize.argtypes = [C.c_void_p, C.c_int, C.POINTER(ci), C.POINTER(cf), C.POINTER(cf), C.POINTER(cf),
                                 C.POINTER(cf), C.POINTER(cf), C.POINTER(cf), C.c_double, C.c_double, cf,
                                 C.POINTER(C.c_int16)]
    L.fpo_advance.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(ci)] + [C.POINTER(cf)] * 6 + \
                             [C.POINTER(C.c_int), C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(cf),
                              C.POINTER(cf), C.POINTER(C.c_int16)]
    return initialize, advance


def _bits(x):
    return np.asarray(x, np.float32).view(np.uint32) if not isinstance(x, float) or True else x


def _same(a, b):
    """bitwise equality of python floats that carry float32 / float64 values, ints, arrays"""
    if isinstance(a, np.ndarray):
        return np.array_equal(a.view(np.uint32), b.view(np.uint32))
    if isinstance(a, float):
        return np.float64(a).tobytes() == np.float64(b).tobytes() or (np.isnan(a) and np.isnan(b))
    return a == b


def _run_particle_loop(cb, ref, ora, p, nsteps, dt=None):
    """The reference's particle loop order (src/timemanager.f90:531-611): initialize for new
    particles, then advance, particle after particle, both implementations side by side."""
    c = cb.cfg
    o_init, o_adv = _oracle_calls(ora)
    n = p.numpart
    dt = dt or c.lsynctime
    st = dict(xt=p.xtra1[:n].copy(), yt=p.ytra1[:n].copy(), zt=p.ztra1[:n].copy(), ldt=p.idt[:n].copy(),
              up=np.zeros(n, np.float32), vp=np.zeros(n, np.float32), wp=np.zeros(n, np.float32),
              us=np.zeros(n, np.float32), vs=np.zeros(n, np.float32), ws=np.zeros(n, np.float32),
              icbt=np.ones(n, np.int16), alive=np.ones(n, bool))
    counts = dict(calls=0, pbl_like=0, nstop=0)
    for k in range(nsteps):
        itime = k * dt
        for j in range(n):
            if not st["alive"][j]:
                continue
            if k == 0:
                ri = ref.initialize(itime, int(st["ldt"][j]), st["xt"][j], st["yt"][j], float(st["zt"][j]))
                oi = o_init(itime, int(st["ldt"][j]), st["xt"][j], st["yt"][j], float(st["zt"][j]))
                assert all(_same(a, b) for a, b in zip(ri, oi)), ("initialize", j, ri, oi)
                st["ldt"][j], st["up"][j], st["vp"][j], st["wp"][j], st["us"][j], st["vs"][j], st["ws"][j], st["icbt"][j] = ri
            args = (itime, int(p.npoint[j]), int(st["ldt"][j]), float(st["up"][j]), float(st["vp"][j]), float(st["wp"][j]),
                    float(st["us"][j]), float(st["vs"][j]), float(st["ws"][j]), float(st["xt"][j]), float(st["yt"][j]),
                    float(st["zt"][j]), int(st["icbt"][j]))
            ra = ref.advance(*args)
            oa = o_adv(*args, c.maxspec)
            for key in ra:
                assert _same(ra[key], oa[key] if key != "prob" else oa[key][:len(ra[key])]), ("advance", k, j, key, ra[key], oa[key])
            counts["calls"] += 1
            if ra["nstop"] > 1:
                st["alive"][j] = False
                counts["nstop"] += 1
                continue
            for key, dst in (("ldt", "ldt"), ("up", "up"), ("vp", "vp"), ("wp", "wp"), ("us", "us"), ("vs", "vs"),
                             ("ws", "ws"), ("xt", "xt"), ("yt", "yt"), ("zt", "zt"), ("icbt", "icbt")):
                st[dst][j] = ra[key]
    return counts, st


def test_reference_random_mod_matches_oracle():
    """ran1, ran3, gasdev (src/random_mod.f90) against the oracle's restatement, call by call."""
    cb = cases.config_small(nrel=1, npart_each=8)
    ref = ref_api.Ref(cb, maxrand=MAXRAND)
    ora = Oracle(cb)
    L, R = load(), ref.L
    R.f_ran1.restype = R.f_ran3.restype = R.f_gasdev.restype = C.c_float
    L.fpo_gasdev.restype = C.c_float
    L.fpo_gasdev.argtypes = [C.c_void_p, C.POINTER(C.c_int32)]
    for fn_r, fn_o, seed in ((R.f_ran1, L.fpo_ran1, -7), (R.f_ran3, L.fpo_ran3, -320), (R.f_ran3, L.fpo_ran3, -7),
                             (R.f_gasdev, L.fpo_gasdev, -11)):
        ir, io = C.c_int(seed), C.c_int32(seed)
        for _ in range(3000):
            a, b = fn_r(C.byref(ir)), fn_o(ora.S, C.byref(io))
            assert np.float32(a).tobytes() == np.float32(b).tobytes() and ir.value == io.value


@pytest.mark.parametrize("ctl,label", [(5.0, "hanna method 1"), (-5.0, "hanna1 method 0 + Petterssen")])
def test_reference_advance_initialize_bit_identical(ctl, label):
    """advance + initialize (with interpol_all/misslev/wind/wind_short, hanna/hanna1/hanna_short,
    windalign, polar cmapf branch, reflection, Petterssen) on the synthetic global met:
    every output of every call bit-identical, cross-particle module-state leaks included."""
    cb = cases.config_small(nrel=4, npart_each=150, ctl=ctl)
    n = 600
    p = cases.seeded_particles(cb, n, zmax=7000.0, lat_range=(-89.0, 89.0))
    p.ztra1[:200] = np.random.RandomState(1).uniform(1.0, 300.0, 200).astype(np.float32)
    ref, ora = _pair(cb, cases.met_pair(cb))
    counts, st = _run_particle_loop(cb, ref, ora, p, 4)
    assert counts["calls"] >= 4 * n - 50


def test_reference_nested_grids_and_drydep_bit_identical():
    """nested met input (interpol_*_nests, nest choice, hmixn/tropopausen), dry-deposition
    probability (interpol_vdep(_nests)) and settling (get_settling + viscosity)."""
    nests = [(-40.0, 10.0, 81, 41, 1.0, 1.0)]
    cb = cases.config_small(nrel=4, npart_each=128, met_nests=nests, nspec=2, drydepspec=(1, 0), xmass=np.ones((4, 2)),
                            lsettling=1, density=(0.0, 2000.0), dquer=(0.0, 5.0), vsetaver=(0.0, -0.01),
                            cunningham=(1.0, 1.1))
    n = 512
    p = cases.seeded_particles(cb, n, zmax=5000.0, lat_range=(0.0, 60.0), nspec=2)
    r = np.random.RandomState(3)
    p.xtra1[:384] = (r.uniform(-50.0, 50.0, 384) - cb.cfg.xlon0) / cb.cfg.dx
    p.ztra1[:128] = r.uniform(1.0, 40.0, 128).astype(np.float32)
    nm = (fb.MetFields(cb, nest=1).synth(0), fb.MetFields(cb, nest=1).synth(10800))
    for m in nm:
        m.uu += 3.0; m.hmix *= 1.2; m.vdep *= 2.0
    ref, ora = _pair(cb, cases.met_pair(cb), nest_mets=[nm])
    counts, st = _run_particle_loop(cb, ref, ora, p, 3)
    assert counts["calls"] >= 3 * n - 50


def test_reference_cbl_bit_identical():
    """CBLFLAG=1: cbl, re_initialize_particle, initialize_cbl_vel (ran3 + gasdev from the global stream)."""
    cb = cases.config_small(nrel=2, npart_each=150, cblflag=1, ctl=10.0, ifine=5)
    n = 300
    p = cases.seeded_particles(cb, n, zmax=1500.0, lat_range=(-50.0, 50.0))
    ref, ora = _pair(cb, cases.met_pair(cb))
    counts, st = _run_particle_loop(cb, ref, ora, p, 3)
    assert counts["calls"] >= 3 * n - 20


@pytest.mark.parametrize("ctl", [5.0, -5.0])
def test_reference_backward_run_bit_identical(ctl):
    """LDIRECT=-1, method 1 and method 0 (CTL<0: mintime must stay +|lsynctime|, src/readcommand.f90:384
    vs :631, or ldt=max(ldt,mintime) (src/advance.f90:510) goes negative and the Petterssen corrector
    (src/advance.f90:829) is skipped)."""
    cb = cases.config_small(nrel=2, npart_each=128, ldirect=-1, ctl=ctl)
    assert cb.cfg.mintime == (1 if ctl > 0 else 900) and cb.cfg.lsynctime == -900
    n = 256
    p = cases.seeded_particles(cb, n, zmax=4000.0)
    mets = (fb.MetFields(cb).synth(0), fb.MetFields(cb).synth(-10800))
    ref, ora = _pair(cb, mets, bracket=(0, -10800))
    counts, st = _run_particle_loop(cb, ref, ora, p, 3, dt=-900)
    assert counts["calls"] >= 3 * n - 20


@pytest.mark.parametrize("ind_samp", [0, -1])
def test_reference_conccalc_bit_identical(ind_samp):
    """conccalc (mother + nested output grid, age classes, kernel / no-kernel branches,
    density-weighted sampling, receptors): gridunc, griduncn, creceptor bit-identical
    (same serial accumulation order)."""
    cb = cases.config_small(nrel=3, npart_each=400, nspec=2, ind_samp=ind_samp, lage=(7200, 86400 * 10),
                            nest=(-60.0, -30.0, 48, 24, 2.5, 2.5), ioutputforeachrelease=1, xmass=np.ones((3, 2)),
                            receptors=((30.0, 20.0, 1.0e10), (40.5, 18.2, 2.0e10)))
    c = cb.cfg
    n = 1200
    p = cases.seeded_particles(cb, n, zmax=6000.0, lat_range=(-70.0, 70.0), nspec=2)
    p.itramem[:400] = -30000          # old enough for the kernel branch and the 2nd age class
    p.xmass1[:n, 1] = 0.5
    p.itra1[:n] = 900
    r = np.random.RandomState(5)        # a cloud of low particles around the receptors
    p.xtra1[400:500] = r.uniform(28.0, 42.0, 100); p.ytra1[400:500] = r.uniform(17.0, 21.0, 100)
    p.ztra1[400:500] = r.uniform(1.0, 140.0, 100).astype(np.float32)
    ref, ora = _pair(cb, cases.met_pair(cb))
    ref.push_particles(p)
    ora.push_particles(p)
    for e in (ref, ora):
        e.conccalc(900, 0.5)
        e.conccalc(900, 1.0)
    go = ora.fetch_grids(zero_conc=False)
    assert go["gridunc"].sum() > 0 and go["griduncn"].sum() > 0 and go["creceptor"].sum() > 0
    assert np.array_equal(ref.arr("gridunc").view(np.uint32), go["gridunc"].view(np.uint32))
    assert np.array_equal(ref.arr("griduncn").view(np.uint32), go["griduncn"].view(np.uint32))
    assert np.array_equal(ref.arr("creceptor")[:, :c.maxspec].view(np.uint32), go["creceptor"][:, :c.maxspec].view(np.uint32))


@pytest.mark.parametrize("readclouds", [0, 1])
def test_reference_wetdepo_bit_identical(readclouds):
    """wetdepo + get_wetscav + interpol_rain(_nests) + wetdepokernel(_nest):
    particle masses and wetgridunc / wetgriduncn bit-identical (serial accumulation order)."""
    cb = cases.config_small(nrel=3, npart_each=512, nspec=2, wetdepspec=(1, 1), weta_gas=(2.0e-5, -1.0),
                            wetb_gas=(0.62, -1.0), henry=(1.0e-2, 0.0), crain_aero=(-1.0, 1.0),
                            csnow_aero=(-1.0, 1.0), ccn_aero=(-1.0, 0.9), in_aero=(-1.0, 0.1),
                            dquer=(0.0, 0.6), decay=(0.0, 2.0e-6), readclouds=readclouds,
                            met_nests=[(-60.0, -20.0, 121, 81, 1.0, 1.0)], nest=(-60.0, -30.0, 48, 24, 2.5, 2.5),
                            ioutputforeachrelease=1, lage=(7200, 86400 * 10), xmass=np.ones((3, 2)))
    c = cb.cfg
    n = 1536
    p = cases.seeded_particles(cb, n, zmax=9000.0, lat_range=(-70.0, 70.0), nspec=2)
    p.itramem[:512] = -30000
    p.xmass1[:n, 1] = 0.5
    p.itra1[:n] = 900
    p.itra1[100:120] = 1800           # not yet due
    p.itra1[120:130] = fb.ITRA_DEAD
    mets = cases.met_pair(cb)
    nm = (fb.MetFields(cb, nest=1).synth(0), fb.MetFields(cb, nest=1).synth(10800))
    for m in nm:
        m.lsprec *= 1.5; m.tt -= 5.0; m.ctwc *= 2.0
    ref, ora = _pair(cb, mets, nest_mets=[nm])
    for slot in (1, 2):
        ref.upload_rain(slot, mets[slot - 1])
        ref.upload_rain(slot, nm[slot - 1], nest=1)
    ref.push_particles(p)
    ora.push_particles(p)
    # loutnext = 1800, loutstep = 3600: itime <= loutnext -> ldeltat = itime - (loutnext - loutstep)
    ref.wetdepo(900, 900, 1800)
    ora.wetdepo(900, 900, 900 - (1800 - 3600))
    q = fb.Particles(c.maxpart, c.nspec); q.numpart = n
    ora.pull_particles(q)
    assert np.array_equal(ref.arr("xmass1")[:n, :2].view(np.uint32), q.xmass1[:n].view(np.uint32))
    wo = ora.fetch_wetgrids()
    assert wo["wetgridunc"][:, :, 0].sum() > 0 and wo["wetgridunc"][:, :, 1].sum() > 0 and wo["wetgriduncn"].sum() > 0
    assert np.array_equal(ref.arr("wetgridunc").view(np.uint32), wo["wetgridunc"].view(np.uint32))
    assert np.array_equal(ref.arr("wetgriduncn").view(np.uint32), wo["wetgriduncn"].view(np.uint32))


def test_reference_drydepokernel_and_pieces_bit_identical():
    """drydepokernel(_nest), windalign, hanna / hanna1 / hanna_short and the cmapf_mod
    transformations called directly with random arguments."""
    cb = cases.config_small(nrel=2, npart_each=8, nspec=2, drydepspec=(1, 1), xmass=np.ones((2, 2)),
                            nest=(-60.0, -30.0, 48, 24, 2.5, 2.5), ioutputforeachrelease=1, lage=(7200, 86400 * 10))
    c = cb.cfg
    ref, ora = _pair(cb, cases.met_pair(cb))
    L, R = load(), ref.L
    r = np.random.RandomState(11)
    cf, ci = C.c_float, C.c_int
    # drydepokernel / _nest
    L.fpo_drydepokernel.argtypes = L.fpo_drydepokernel_nest.argtypes = [C.c_void_p, ci, C.POINTER(cf), cf, cf, ci, ci]
    for _ in range(400):
        dep = (cf * 5)(*(r.uniform(0, 1, 5) * (r.uniform(0, 1, 5) > 0.3)))
        x, y = float(np.float32(r.uniform(0, c.nxmin1))), float(np.float32(r.uniform(0, c.nymin1)))
        nage, kp = int(r.randint(1, 3)), int(r.randint(1, 3))
        for fr, fo in ((R.f_drydepokernel, L.fpo_drydepokernel), (R.f_drydepokernel_nest, L.fpo_drydepokernel_nest)):
            fr(C.byref(ci(1)), dep, C.byref(cf(x)), C.byref(cf(y)), C.byref(ci(nage)), C.byref(ci(kp)))
            fo(ora.S, 1, dep, x, y, nage, kp)
    go = ora.fetch_grids(zero_conc=False)
    assert go["drygridunc"].sum() > 0 and go["drygriduncn"].sum() > 0
    assert np.array_equal(ref.arr("drygridunc").view(np.uint32), go["drygridunc"].view(np.uint32))
    assert np.array_equal(ref.arr("drygriduncn").view(np.uint32), go["drygriduncn"].view(np.uint32))
    # windalign
    L.fpo_windalign.argtypes = [cf] * 4 + [C.POINTER(cf)] * 2
    for _ in range(500):
        a = [float(np.float32(v)) for v in r.normal(0, 5, 4)]
        ux, vy, ox, oy = cf(), cf(), cf(), cf()
        R.f_windalign(*[C.byref(cf(v)) for v in a], C.byref(ux), C.byref(vy))
        L.fpo_windalign(*a, C.byref(ox), C.byref(oy))
        assert (ux.value, vy.value) == (ox.value, oy.value)
    # cmapf_mod: cll2xy, cxy2ll, cgszll on the run's polar maps
    R.f_cgszll.restype = cf
    L.fpo_cgszll.restype = cf
    for mp in ("northpolemap", "southpolemap"):
        mr = ref.arr(mp).ctypes.data_as(C.POINTER(cf))
        mo = (cf * 9)(*getattr(c, mp))
        for _ in range(500):
            lat = float(np.float32(r.uniform(60, 90) * (1 if mp[0] == "n" else -1)))
            lon = float(np.float32(r.uniform(-180, 180)))
            x1, y1, x2, y2 = cf(), cf(), cf(), cf()
            R.f_cll2xy(mr, C.byref(cf(lat)), C.byref(cf(lon)), C.byref(x1), C.byref(y1))
            L.fpo_cll2xy(mo, lat, lon, C.byref(x2), C.byref(y2))
            assert (x1.value, y1.value) == (x2.value, y2.value)
            la1, lo1, la2, lo2 = cf(), cf(), cf(), cf()
            R.f_cxy2ll(mr, C.byref(x1), C.byref(y1), C.byref(la1), C.byref(lo1))
            L.fpo_cxy2ll(mo, x1.value, y1.value, C.byref(la2), C.byref(lo2))
            assert (la1.value, lo1.value) == (la2.value, lo2.value)
            assert R.f_cgszll(mr, C.byref(cf(lat)), C.byref(cf(lon))) == L.fpo_cgszll(mo, lat, lon)


@pytest.mark.parametrize("ctl", [5.0, -5.0])
def test_reference_timemanager_particle_loop_bit_identical(ctl):
    """The particle loop of timemanager itself (src/timemanager.f90:531-712: age class,
    initialize for new particles, advance, nstop, radioactive decay, dry-deposition split with
    the ldeltat back-correction, xmassfract / age terminations, drydepokernel(_nest)) against
    the oracle's fpo_step: every particle array and the deposition grids, 4 intervals."""
    cb = cases.config_small(nrel=3, npart_each=300, nspec=2, decay=[0.0, 1.0e-5], drydepspec=[1, 1], ctl=ctl,
                            lage=(3600, 86400 * 10), nest=(-60.0, -30.0, 48, 24, 2.5, 2.5), ioutputforeachrelease=1)
    c = cb.cfg
    n = 900
    p = cases.seeded_particles(cb, n, zmax=400.0, lat_range=(-60.0, 60.0), nspec=2)
    p.ztra1[300:600] = np.random.RandomState(2).uniform(500.0, 9000.0, 300).astype(np.float32)
    p.itramem[:100] = -86400 * 10 + 1800      # reach lage(nageclass) during the run -> age termination
    p.xmass1[:n, 1] = 0.5
    p.xmass1[200:220, :] = 1.0e-9             # xmassfract < minmass -> mass termination
    ref, ora = _pair(cb, cases.met_pair(cb))
    ref.push_state(p)
    ora.push_particles(p)
    tot_term = 0
    for k in range(4):
        itime, ldeltat = k * 900, 450 + 900 * (k % 2)
        ref.particle_loop(itime, ldeltat)
        st = ora.step(itime, ldeltat)
        tot_term += st["n_terminated"]
        pr = fb.Particles(c.maxpart, c.nspec); pr.numpart = n
        po = fb.Particles(c.maxpart, c.nspec); po.numpart = n
        ref.pull_state(pr); ora.pull_particles(po)
        for f in ref.STATE:
            a, b = getattr(pr, f)[:n], getattr(po, f)[:n]
            assert np.array_equal(a.view(np.uint8), b.view(np.uint8)), (k, f)
        assert np.array_equal(pr.xmass1[:n].view(np.uint32), po.xmass1[:n].view(np.uint32)), k
    assert tot_term >= 120
    go = ora.fetch_grids(zero_conc=False)
    assert go["drygridunc"].sum() > 0 and go["drygriduncn"].sum() > 0
    assert np.array_equal(ref.arr("drygridunc").view(np.uint32), go["drygridunc"].view(np.uint32))
    assert np.array_equal(ref.arr("drygriduncn").view(np.uint32), go["drygriduncn"].view(np.uint32))


@pytest.mark.parametrize("kind", ["dry", "wet"])
def test_reference_backward_receptor_scavenging_bit_identical(kind):
    """The RECEPTOR block of the particle loop (src/timemanager.f90:563-598) in a backward deposition
    run: xscav_frac1 is set once after the release from get_vdep_prob (IND_RECEPTOR = 4) or
    get_wetscav * release depth * grfraction (IND_RECEPTOR = 3); the species that is not deposited
    loses its mass.  Particle loop of the reference against fpo_step, two intervals."""
    kw = dict(nrel=3, npart_each=300, ldirect=-1, nspec=2, lage=(86400 * 10,), ioutputforeachrelease=1,
              xmass=np.ones((3, 2)))
    if kind == "dry":
        kw.update(ind_receptor=4, drydepspec=(1, 0))
    else:
        kw.update(ind_receptor=3, wetdepspec=(1, 0), weta_gas=(2.0e-5, -1.0), wetb_gas=(0.62, -1.0), henry=(1.0e-2, 0.0))
    cb = cases.config_small(**kw)
    c = cb.cfg
    assert (c.drybkdep, c.wetbkdep) == ((1, 0) if kind == "dry" else (0, 1))
    n = 900
    p = cases.seeded_particles(cb, n, zmax=60.0 if kind == "dry" else 9000.0, lat_range=(-70.0, 70.0), nspec=2)
    p.xscav_frac1[:n] = -1.0                       # releaseparticles.f90:171
    p.xmass1[:n, 1] = 0.5
    rel = cases.releases_boxes(cb, seed=5, zmax=1500.0)
    mets = (fb.MetFields(cb).synth(0), fb.MetFields(cb).synth(-10800))
    ref, ora = _pair(cb, mets, bracket=(0, -10800))
    if kind == "wet":
        for slot in (1, 2):
            ref.upload_rain(slot, mets[slot - 1])
    ref.arr("zpoint1")[:] = rel.zpoint1; ref.arr("zpoint2")[:] = rel.zpoint2
    ora.set_releases(rel)
    ref.push_state(p)
    ora.push_particles(p)
    for k in range(2):
        itime = -k * 900
        ref.particle_loop(itime, 450)
        ora.step(itime, 450)
        pr = fb.Particles(c.maxpart, c.nspec); pr.numpart = n
        po = fb.Particles(c.maxpart, c.nspec); po.numpart = n
        ref.pull_state(pr); ora.pull_particles(po)
        for f in ref.STATE:
            a, b = getattr(pr, f)[:n], getattr(po, f)[:n]
            assert np.array_equal(a.view(np.uint8), b.view(np.uint8)), (k, f)
        assert np.array_equal(pr.xmass1[:n].view(np.uint32), po.xmass1[:n].view(np.uint32)), k
        xs = ref.arr("xscav_frac1")[:n, :2]
        assert np.array_equal(xs.view(np.uint32), po.xscav_frac1[:n].view(np.uint32)), k
        assert (xs >= 0).all() and (xs[:, 0] > 0).sum() > 50 and (xs[:, 1] == 0).all()
        assert (po.xmass1[:n, 1] == 0).all()      # the species without deposition contributes nothing
        if kind == "wet":
            assert ((xs[:, 0] == 0) & (po.xmass1[:n, 0] == 0)).sum() > 50   # particles outside the rain


def test_reference_releaseparticles_bit_identical():
    """releaseparticles (src/releaseparticles.f90:69-378): release counts per interval, the
    free-slot search, the ran1 position stream and the particle masses, against the oracle's
    restatement (which tests/test_oracle_pins.py in turn equates with the host library's)."""
    cb = cases.config_small(nrel=3, npart_each=700, maxpart=2600)
    c = cb.cfg
    rel = cases.releases_boxes(cb, seed=3, start=0, end=3600)
    ref = ref_api.Ref(cb, maxrand=MAXRAND)
    o = Oracle(cb)
    L = o.L
    for nm in ("xpoint1", "xpoint2", "ypoint1", "ypoint2", "zpoint1", "zpoint2"):
        ref.arr(nm)[:] = getattr(rel, nm)
    ref.arr("ireleasestart")[:] = rel.start
    ref.arr("ireleaseend")[:] = rel.end
    ref.arr("kindz")[:] = 1
    ref.arr("xmasssave")[:] = 0.0
    for nm in ("area_hour", "point_hour", "area_dow", "point_dow"):
        ref.arr(nm)[:] = 1.0              # no emission variation (readreleases default)
    ref.set("ind_rel", 0); ref.set("itsplit", 99999999); ref.set("bdate", 2455197.5)
    ref.set("numpart", 0); ref.set("numparticlecount", 0)
    ref.arr("itra1")[:] = fb.ITRA_DEAD   # FLEXPART.f90:315-317
    xmasssave = np.zeros(c.numpoint, np.float32)
    _pf, _pi = C.POINTER(C.c_float), C.POINTER(C.c_int32)
    fp = lambda a: a.ctypes.data_as(_pf)
    for itime in range(0, 4500, 900):
        rc = L.fpo_releaseparticles(o.S, itime, c.numpoint, rel.start.ctypes.data_as(_pi), rel.end.ctypes.data_as(_pi),
                                    fp(rel.xpoint1), fp(rel.ypoint1), fp(rel.xpoint2), fp(rel.ypoint2),
                                    fp(rel.zpoint1), fp(rel.zpoint2), fp(xmasssave), 99999999)
        assert rc == 0
        ref.L.f_releaseparticles(C.byref(C.c_int(itime)))
        k = ref.get("numpart")
        q = fb.Particles(c.maxpart, 1); q.numpart = k
        o.pull_particles(q)
        for f in ("xtra1", "ytra1", "ztra1", "itra1", "itramem", "npoint", "nclass", "idt"):
            a, b = ref.arr(f)[:k], getattr(q, f)[:k]
            assert np.array_equal(a.view(np.uint8), np.ascontiguousarray(b).view(np.uint8)), (itime, f)
        assert np.array_equal(ref.arr("xmass1")[:k, 0].view(np.uint32), q.xmass1[:k, 0].view(np.uint32))
        # the released particles "advance": their slots stay occupied
        ref.arr("itra1")[:k] = itime + 900
        q.itra1[:k] = itime + 900
        o.push_particles(q, 0, k)
    assert k == 3 * (87 + 175 * 3 + 88)


@pytest.mark.parametrize("box,npart1", [((-180.0, -90.0, 180.0, 90.0), 30000),    # global: gdomainfill
                                        ((-180.0, -90.0, 180.0, 90.0), 2500),     # sparse: ncolumn <= 20 -> random heights
                                        ((-40.0, 10.0, 65.0, 72.5), 12000)])      # limited domain
def test_reference_init_domainfill_bit_identical(box, npart1):
    """init_domainfill (src/init_domainfill.f90:55-283, MDOMAINFILL = 1): the domain box, column air
    masses, particles per column, the pressure-equidistant (or, for thin columns, random) heights,
    the ran1 positions, masses and the out-of-domain termination against the oracle's restatement."""
    cb = cases.config_small(nrel=1, npart_each=npart1, maxpart=npart1 + 2000, mdomainfill=1, nclassunc=3)
    c = cb.cfg
    m0, m1 = cases.met_pair(cb)
    ref = ref_api.Ref(cb, maxrand=MAXRAND)
    o = Oracle(cb)
    for e in (ref, o):
        e.upload_met(1, m0); e.upload_met(2, m1)
    pts = [(box[0] - c.xlon0) / c.dx, (box[1] - c.ylat0) / c.dy, (box[2] - c.xlon0) / c.dx, (box[3] - c.ylat0) / c.dy]
    pts = [float(np.float32(v)) for v in pts]
    for nm, v in zip(("xpoint1", "ypoint1", "xpoint2", "ypoint2"), pts):
        ref.arr(nm)[0] = v
    ref.set("ipin", 0); ref.set("itsplit", 99999999); ref.set("numpart", 0); ref.set("numparticlecount", 0)
    ref.set("gdomainfill", 0)
    ref.arr("itra1")[:] = fb.ITRA_DEAD
    ref.L.f_init_domainfill()
    n, info = o.init_domainfill(pts)
    assert n == ref.get("numpart") and n > 0.9 * npart1
    assert info["nx_we"] == tuple(ref.arr("nx_we")) and info["ny_sn"] == tuple(ref.arr("ny_sn"))
    assert info["gdomainfill"] == ref.get("gdomainfill") == (1 if box[0] == -180.0 else 0)
    assert info["numcolumn"] == ref.get("numcolumn")
    assert np.float32(info["xmassperparticle"]).tobytes() == np.float32(ref.get("xmassperparticle")).tobytes()
    q = fb.Particles(c.maxpart, 1); q.numpart = n
    o.pull_particles(q)
    for f in ("xtra1", "ytra1", "ztra1", "itra1", "itramem", "npoint", "nclass", "idt", "itrasplit"):
        a, b = ref.arr(f)[:n], getattr(q, f)[:n]
        assert np.array_equal(a.view(np.uint8), np.ascontiguousarray(b).view(np.uint8)), f
    assert np.array_equal(ref.arr("xmass1")[:n, 0].view(np.uint32), q.xmass1[:n, 0].view(np.uint32))
    assert len(np.unique(q.nclass[:n])) == 3
    # every particle carries its column's share: the masses add up to the air mass of the box
    # (thin columns round to 0 or 1 particle: only the dense cases close to better than a per cent)
    assert abs(q.xmass1[:n, 0].astype(np.float64).sum() / info["colmasstotal"] - 1.0) < (2e-3 if npart1 >= 10000 else 0.1)


def test_reference_boundcond_domainfill_bit_identical():
    """boundcond_domainfill (src/boundcond_domainfill.f90:54-560) on a limited domain-filling box:
    the boundary release heights memorised by init_domainfill (:287-389), then per call the
    termination of particles outside the box, the mass flux through every boundary location, the
    accumulated masses and the particles created from them (slots, positions, ran1 stream)."""
    npart1 = 60000
    cb = cases.config_small(nrel=1, npart_each=npart1, maxpart=npart1 + 20000, mdomainfill=1, nclassunc=3)
    c = cb.cfg
    m0, m1 = cases.met_pair(cb)
    ref = ref_api.Ref(cb, maxrand=MAXRAND)
    o = Oracle(cb)
    for e in (ref, o):
        e.upload_met(1, m0); e.upload_met(2, m1)
        e.set_met_bracket((1, 2), (0, 10800))
    box = (-60.0, -30.0, 70.0, 45.0)
    pts = [(box[0] - c.xlon0) / c.dx, (box[1] - c.ylat0) / c.dy, (box[2] - c.xlon0) / c.dx, (box[3] - c.ylat0) / c.dy]
    pts = [float(np.float32(v)) for v in pts]
    for nm, v in zip(("xpoint1", "ypoint1", "xpoint2", "ypoint2"), pts):
        ref.arr(nm)[0] = v
    ref.set("ipin", 0); ref.set("itsplit", 99999999); ref.set("numpart", 0); ref.set("numparticlecount", 0)
    ref.set("gdomainfill", 0); ref.set("ipout", 0)
    ref.arr("itra1")[:] = fb.ITRA_DEAD
    ref.L.f_init_domainfill()
    n, info = o.init_domainfill(pts)
    assert info["gdomainfill"] == 0 and n == ref.get("numpart")
    nloc, _ = o.boundcond_locations()
    assert nloc == int(ref.arr("numcolumn_we").sum() + ref.arr("numcolumn_sn").sum()) and nloc > 1000
    assert ref.arr("numcolumn_we").max() >= 3
    rs = np.random.RandomState(4)
    created_total = 0
    for k in range(10):
        itime = k * c.lsynctime
        q = fb.Particles(c.maxpart, 1); q.numpart = o.numpart()
        o.pull_particles(q)
        npt = q.numpart
        if k:   # the particle loop moved the particles on: here a random walk, some of them out of the box
            mv = q.itra1[:npt] == itime
            q.xtra1[:npt][mv] += rs.normal(0.0, 0.8, mv.sum())
            q.ytra1[:npt][mv] += rs.normal(0.0, 0.8, mv.sum())
            o.push_particles(q)
            ref.arr("xtra1")[:npt] = q.xtra1[:npt]; ref.arr("ytra1")[:npt] = q.ytra1[:npt]
        itime_c = C.c_int(itime); lout = C.c_int(10 ** 9)
        ref.L.f_boundcond_domainfill(C.byref(itime_c), C.byref(lout))
        created = o.boundcond_domainfill(itime)
        created_total += created
        n = o.numpart()
        assert n == ref.get("numpart"), k
        q = fb.Particles(c.maxpart, 1); q.numpart = n
        o.pull_particles(q)
        for f in ("xtra1", "ytra1", "ztra1", "itra1", "itramem", "npoint", "nclass", "idt", "itrasplit"):
            a, b = ref.arr(f)[:n], getattr(q, f)[:n]
            assert np.array_equal(a.view(np.uint8), np.ascontiguousarray(b).view(np.uint8)), (k, f)
        assert np.array_equal(ref.arr("xmass1")[:n, 0].view(np.uint32), q.xmass1[:n, 0].view(np.uint32)), k
        assert o.numparticlecount() == ref.get("numparticlecount")
        # advance the clock of the live particles as the particle loop would
        live = q.itra1[:n] == itime
        q.itra1[:n][live] = itime + c.lsynctime
        o.push_particles(q)
        ref.arr("itra1")[:n] = q.itra1[:n]
    assert created_total > 80
    acc_ref = float(ref.arr("acc_mass_we").astype(np.float64).sum() + ref.arr("acc_mass_sn").astype(np.float64).sum())
    assert abs(o.boundcond_locations()[1] - acc_ref) <= 1e-9 * abs(acc_ref)


def test_reference_readcommand_derivations_match_host():
    """fpbh_readcommand (turbulence switches, ifine, fine, ctl := 1/ctl, method, mintime) against
    src/readcommand.f90:244-272,379-385 run from the reference's source."""
    cb = cases.config_small(nrel=1, npart_each=8)
    ref = ref_api.Ref(cb, maxrand=MAXRAND)
    from flexpart_b200.abi import FpbConfig, load_host_lib
    H = load_host_lib()
    for ctl in (-5.0, -1.0, 0.05, 0.1, 1.0, 3.0, 5.0, 7.5, 40.0):
        for ifine in (0, 1, 4, 5, 20):
            for cbl in (0, 1):
                for lsync in (300, 900, 3600):
                    ref.set("ctl", ctl); ref.set("ifine", ifine); ref.set("cblflag", cbl); ref.set("lsynctime", lsync)
                    ref.set("turbswitch", 0)
                    ref.L.f_rc_turbulence_switches()
                    ref.L.f_rc_method()
                    c = FpbConfig()
                    c.ldirect, c.lsynctime, c.ctl, c.ifine, c.cblflag = 1, lsync, ctl, ifine, cbl
                    assert H.fpbh_readcommand(C.byref(c)) == 0
                    got = (c.ifine, c.turbswitch, c.fine, c.ctl, c.method, c.mintime, c.lsynctime)
                    exp = (ref.get("ifine"), ref.get("turbswitch"), ref.get("fine"), ref.get("ctl"), ref.get("method"),
                           ref.get("mintime"), ref.get("lsynctime"))
                    assert got == exp, (ctl, ifine, cbl, lsync, got, exp)
                    # backward run: only lsynctime changes sign, after method/mintime were derived
                    # (src/readcommand.f90:627-634)
                    b = FpbConfig()
                    b.ldirect, b.lsynctime, b.ctl, b.ifine, b.cblflag = -1, lsync, ctl, ifine, cbl
                    assert H.fpbh_readcommand(C.byref(b)) == 0
                    assert (b.ifine, b.turbswitch, b.fine, b.ctl, b.method, b.mintime, b.lsynctime) == \
                        exp[:6] + (-exp[6],), (ctl, ifine, cbl, lsync)


def test_reference_outgrid_geometry_and_sparse_dump_bit_identical():
    """outgrid_init's areas and volumes (src/outgrid_init.f90:51-99) and concoutput's work on one
    (ks, kp, nage) grid -- factor3d (:214-225), the class mean (:275-335, mean_mod), and the three
    sparse dumps (wet :353-381, dry :390-418, concentration :429-467) -- against the oracle's
    restatement (oracle/fpo_output.c), which the GPU tests in turn equate with the device's."""
    cb = cases.config_small(nrel=2, npart_each=10, nclassunc=3, ioutputforeachrelease=1, lage=(1800, 86400),
                            nspec=2, drydepspec=(1, 1), wetdepspec=(1, 0), weta_gas=(2.0e-5, -1.0),
                            wetb_gas=(0.62, -1.0), henry=(1.0e-2, 0.0))
    c = cb.cfg
    ref = ref_api.Ref(cb, maxrand=MAXRAND)
    o = Oracle(cb)
    L = o.L
    _pf, _pi = C.POINTER(C.c_float), C.POINTER(C.c_int32)
    outlat0 = np.float32(c.ylat0) - np.float32(c.youtshift)
    ref.set("outlat0", float(outlat0))
    ref.L.f_og_geometry()
    area = np.zeros((c.numxgrid, c.numygrid), np.float32, order="F")
    vol = np.zeros((c.numxgrid, c.numygrid, c.numzgrid), np.float32, order="F")
    L.fpo_outgrid_geometry(C.byref(c), 0, float(outlat0), area.ctypes.data_as(_pf), vol.ctypes.data_as(_pf))
    assert np.array_equal(ref.arr("area").view(np.uint32), area.view(np.uint32))
    assert np.array_equal(ref.arr("volume").view(np.uint32), vol.view(np.uint32))

    # random sparse grids with runs, gaps, denormals and exact zeros; classes differ
    r = np.random.RandomState(4)
    n2, n3 = c.numxgrid * c.numygrid, c.numxgrid * c.numygrid * c.numzgrid
    for name in ("gridunc", "drygridunc", "wetgridunc"):
        a = ref.arr(name)
        v = r.uniform(0.0, 1.0, a.shape).astype(np.float32) * (r.uniform(size=a.shape) < 0.35)
        v[r.uniform(size=a.shape) < 0.01] = 1e-39
        blk = r.uniform(size=a.shape[:2]) < 0.5          # whole columns empty: longer gaps
        v[blk] = 0.0
        a[...] = v
    ref.set("wetdep", 1)
    outnum = np.float32(7.0)
    ref.L.f_co_factor3d(C.byref(C.c_float(outnum)))
    tot_mu = np.asfortranarray(r.uniform(0.5, 2.0, (c.maxspec, c.maxpointspec_act)).astype(np.float32))
    # densityoutgrid (:164-190) from a synthetic rho at memind(2)
    m2 = fb.MetFields(cb).synth(10800)
    ref.upload_met(2, m2); ref.set_met_bracket((1, 2), (0, 10800))
    ref.set("outlon0", float(np.float32(c.xlon0) - np.float32(c.xoutshift)))
    ref.arr("weightmolar")[:2] = (350.5, 28.0)
    ref.L.f_co_density()
    dens = np.zeros(n3, np.float32)
    rho2 = np.ascontiguousarray(m2.rho.reshape(-1, order="F"))
    L.fpo_density_outgrid(C.byref(c), cb.height.ctypes.data_as(_pf), 0, float(ref.get("outlon0")), float(outlat0),
                          rho2.ctypes.data_as(_pf), dens.ctypes.data_as(_pf))
    assert np.array_equal(ref.arr("densityoutgrid").reshape(-1, order="F").view(np.uint32), dens.view(np.uint32))
    volf = np.ascontiguousarray(vol.reshape(-1, order="F")); areaf = np.ascontiguousarray(area.reshape(-1, order="F"))
    total = 0
    for ks in (1, 2):
        for kp in (1, 2):
            for nage in (1, 2):
                ref.L.f_co_mean(C.byref(C.c_int(ks)), C.byref(C.c_int(kp)), C.byref(C.c_int(nage)))
                for which, fn, gname, geom, n in ((2, "f_co_wet", "wetgridunc", areaf, n2), (1, "f_co_dry", "drygridunc", areaf, n2),
                                                  (0, "f_co_conc", "gridunc", volf, n3), (3, "f_co_pptv", "gridunc", volf, n3)):
                    ci, cr = C.c_int(0), C.c_int(0)
                    ref.arr("sparse_dump_i")[:] = -7; ref.arr("sparse_dump_r")[:] = np.nan
                    if which == 0:
                        getattr(ref.L, fn)(C.byref(C.c_int(ks)), C.byref(C.c_int(kp)), tot_mu.ctypes.data_as(_pf),
                                           C.byref(ci), C.byref(cr))
                    elif which == 3:
                        getattr(ref.L, fn)(C.byref(C.c_int(ks)), C.byref(C.c_float(outnum)), C.byref(ci), C.byref(cr))
                    else:
                        getattr(ref.L, fn)(C.byref(ci), C.byref(cr))
                    flat = np.ascontiguousarray(ref.arr(gname).reshape(-1, order="F"))
                    di, dr = np.zeros(n, np.int32), np.zeros(n, np.float32)
                    oi, orr = C.c_int32(), C.c_int32()
                    L.fpo_concoutput_sparse(C.byref(c), 0, which, flat.ctypes.data_as(_pf), geom.ctypes.data_as(_pf),
                                            dens.ctypes.data_as(_pf) if which == 3 else None, ks, kp, nage,
                                            float(outnum), float(tot_mu[ks - 1, kp - 1]) if which != 3 else float(ref.arr("weightmolar")[ks - 1]),
                                            3600, C.byref(oi),
                                            di.ctypes.data_as(_pi), C.byref(orr), dr.ctypes.data_as(_pf))
                    assert (ci.value, cr.value) == (oi.value, orr.value), (ks, kp, nage, which)
                    assert np.array_equal(ref.arr("sparse_dump_i")[:ci.value], di[:ci.value])
                    assert np.array_equal(ref.arr("sparse_dump_r")[:cr.value].view(np.uint32), dr[:cr.value].view(np.uint32)), (ks, kp, nage, which)
                    total += cr.value
    assert total > 1000


def test_reference_partoutput_record_bit_identical():
    """The per-particle interpolation of partoutput (src/partoutput.f90:48-50,73-179) against the
    oracle's restatement (oracle/fpo_output.c), which the GPU tests equate with the device's."""
    cb = cases.config_small(nrel=1, npart_each=400)
    c = cb.cfg
    ref = ref_api.Ref(cb, maxrand=MAXRAND)
    L = Oracle(cb).L
    m = [fb.MetFields(cb).synth(0), fb.MetFields(cb).synth(10800)]
    ref.upload_met(1, m[0]); ref.upload_met(2, m[1]); ref.set_met_bracket((1, 2), (0, 10800))
    r = np.random.RandomState(3)
    shp3, shp2 = m[0].uu.shape, m[0].hmix.shape
    pv = [np.asfortranarray(r.normal(0, 2e-6, shp3).astype(np.float32)) for _ in range(2)]
    qv = [np.asfortranarray(r.uniform(0, 0.02, shp3).astype(np.float32)) for _ in range(2)]
    oro = np.asfortranarray(r.uniform(0, 3000.0, shp2).astype(np.float32))
    ref.arr("oro")[...] = oro
    for s in (0, 1):
        ref.arr("pv")[:, :, :, s] = pv[s]; ref.arr("qv")[:, :, :, s] = qv[s]
    p = cases.seeded_particles(cb, 400, zmax=12000.0, lat_range=(-89.0, 89.0))
    ref.push_particles(p)
    itime = 2700
    f = lambda: C.c_float(0.0)
    dt1, dt2, dtt = f(), f(), f()
    ref.L.f_pp_dt(C.byref(C.c_int(itime)), C.byref(dt1), C.byref(dt2), C.byref(dtt))
    _pf = C.POINTER(C.c_float)
    pair = lambda a, b: (_pf * 2)(a.ctypes.data_as(_pf), b.ctypes.data_as(_pf))
    args = [pair(pv[0], pv[1]), pair(qv[0], qv[1]), pair(m[0].tt, m[1].tt), pair(m[0].rho, m[1].rho),
            pair(m[0].hmix, m[1].hmix), pair(m[0].tropopause, m[1].tropopause)]
    memtime = (C.c_int32 * 2)(0, 10800)
    out = np.zeros(9, np.float32)
    for i in range(400):
        v = [f() for _ in range(9)]
        ref.L.f_pp_record(C.byref(C.c_int(i + 1)), C.byref(dt1), C.byref(dt2), C.byref(dtt), *[C.byref(x) for x in v])
        L.fpo_partoutput_record(C.byref(c), cb.height.ctypes.data_as(_pf), itime, memtime, float(p.xtra1[i]), float(p.ytra1[i]),
                                float(p.ztra1[i]), oro.ctypes.data_as(_pf), *args, out.ctypes.data_as(_pf))
        got = np.array([x.value for x in v], np.float32)
        assert np.array_equal(got.view(np.uint32), out.view(np.uint32)), (i, got, out)


def test_reference_particle_splitting_bit_identical():
    """The splitting block of timemanager (src/timemanager.f90:473-504) against the oracle's loop:
    same candidates, same slots, halved masses, doubled itrasplit, saturation at maxpart."""
    cb = cases.config_small(nrel=2, npart_each=300, maxpart=800, nspec=2)
    c = cb.cfg
    ref = ref_api.Ref(cb, maxrand=MAXRAND)
    o = Oracle(cb)
    n = 600
    p = cases.seeded_particles(cb, n, zmax=3000.0, nspec=2)
    r = np.random.RandomState(8)
    p.itrasplit[:n] = r.choice([900, 1800, 99999999], n)
    p.itramem[:n] = r.choice([0, -900], n)
    p.itra1[:n:7] = fb.ITRA_DEAD
    for nm in ("uap", "ucp", "uzp", "us", "vs", "ws"):
        getattr(p, nm)[:n] = r.normal(size=n).astype(np.float32)
    p.idt[:n] = r.randint(1, 900, n)
    o.push_particles(p)
    ref.push_particles(p); ref.push_state(p)
    ref.arr("itrasplit")[:n] = p.itrasplit[:n]
    ref.set("itsplit", 900)
    for itime in (900, 1800):
        ref.L.f_tm_split(C.byref(C.c_int(itime)))
        o.L.fpo_split_particles(o.S, itime)
        k = ref.get("numpart")
        assert k == o.L.fpo_numpart(o.S)
        q = fb.Particles(c.maxpart, 2); q.numpart = k
        o.pull_particles(q)
        for f in ("xtra1", "ytra1", "ztra1", "itra1", "itramem", "itrasplit", "npoint", "nclass", "idt", "uap", "ucp",
                  "uzp", "us", "vs", "ws", "cbt"):
            a, b = ref.arr(f)[:k], getattr(q, f)[:k]
            assert np.array_equal(np.ascontiguousarray(a).view(np.uint8), np.ascontiguousarray(b).view(np.uint8)), (itime, f)
        assert np.array_equal(ref.arr("xmass1")[:k, :2].view(np.uint32), q.xmass1[:k, :2].view(np.uint32))
    assert k == c.maxpart          # the second round ran out of slots
