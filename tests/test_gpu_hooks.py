"""The two optional hooks of the particle loop (src/timemanager.f90:614-623) on the device -- calcfluxes
(iflux = 1) and partpos_average (ipout = 3), run by fpb_step / fpb_step_host around the step kernels --
against the reference's own routines (oracle/_ref: src/calcfluxes.f90, src/partpos_average.f90 from their
sources) called particle by particle on the positions the engine itself produced: flux and every part_av_*
sum must be bit-identical (the masses are powers of two, so the flux sums do not depend on the order of the
atomics)."""
import ctypes as C

import numpy as np
import pytest

import flexpart_b200 as fb
import cases
import ref_api

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not ref_api.available(), reason="oracle/_ref/libflexref.so not built")]

AV = ("cartx", "carty", "cartz", "z", "topo", "pv", "qv", "tt", "uu", "vv", "rho", "tro", "hmix", "energy")


def _setup(host_mode, sort_interval, xglobal_jump=False):
    kw = dict(nrel=4, npart_each=512, nspec=2, lage=(7200, 86400 * 10), ioutputforeachrelease=1, iflux=1, ipout=3,
              math_mode=fb.MATH_FAST, rng_mode=fb.RNG_PHILOX_INDEX, sort_interval=sort_interval, decay=[0.0, 1.0e-5],
              outlon0=-180.0, outlat0=-60.0, numxgrid=72, numygrid=24, dxout=5.0, dyout=5.0)
    cb = cases.config_small(**kw)
    c = cb.cfg
    n = 2048
    p = cases.seeded_particles(cb, n, zmax=6000.0, lat_range=(-70.0, 70.0), nspec=2)
    p.ztra1[:700] = np.random.RandomState(4).uniform(1.0, 400.0, 700).astype(np.float32)
    p.xmass1[:n, 0] = 1.0
    p.xmass1[:n, 1] = 0.5
    p.itramem[:n:5] = -6300            # a fifth sits in the second age class after the first step
    if xglobal_jump:                   # some next to the date line of the global domain
        p.xtra1[:200] = np.random.RandomState(5).uniform(c.nx - 1.2, c.nx - 1.001, 200)
        p.xtra1[200:400] = np.random.RandomState(6).uniform(0.001, 0.2, 200)
    m0, m1 = cases.met_pair(cb)
    r = np.random.RandomState(9)
    shp3, shp2 = m0.uu.shape, m0.hmix.shape
    pv = [np.asfortranarray(r.normal(0, 2e-6, shp3).astype(np.float32)) for _ in range(2)]
    qv = [np.asfortranarray(r.uniform(0, 0.02, shp3).astype(np.float32)) for _ in range(2)]
    oro = np.asfortranarray(r.uniform(0, 3000.0, shp2).astype(np.float32))
    eng = fb.Engine(cb)
    eng.fill_rannumb()
    eng.upload_met(1, m0); eng.upload_met(2, m1); eng.set_met_bracket((1, 2), (0, 10800))
    eng.set_orography(oro); eng.upload_pvqv(1, pv[0], qv[0]); eng.upload_pvqv(2, pv[1], qv[1])
    ref = ref_api.Ref(cb, maxrand=1000)
    ref.upload_met(1, m0); ref.upload_met(2, m1); ref.set_met_bracket((1, 2), (0, 10800))
    ref.arr("oro")[...] = oro
    for s in range(2):
        ref.arr("pv")[:, :, :, s] = pv[s]; ref.arr("qv")[:, :, :, s] = qv[s]
    oh = np.array([c.outheight[k] for k in range(c.numzgrid)], np.float32)
    half = ref.arr("outheighthalf")
    half[0] = oh[0] / np.float32(2.0)                       # src/readoutgrid.f90:194-197
    for k in range(1, c.numzgrid):
        half[k] = (oh[k - 1] + oh[k]) / np.float32(2.0)
    return cb, eng, ref, p, n


def _reference_hooks(ref, cb, pre, post, itime, n):
    """calcfluxes + partpos_average of the reference for every particle the step advanced"""
    c = cb.cfg
    ref.push_state(post)
    ref.arr("xmass1")[:n, :c.nspec] = pre.xmass1[:n]          # the hook runs before decay / deposition
    lage = [c.lage[k] for k in range(c.nageclass)]
    for j in np.nonzero(pre.itra1[:n] == itime)[0]:
        itage = abs(int(pre.itra1[j]) - int(pre.itramem[j]))
        nage = next((k + 1 for k, v in enumerate(lage) if itage < v), c.nageclass + 1)
        if post.itra1[j] != fb.ITRA_DEAD:
            ref.L.f_partpos_average(C.byref(C.c_int(itime)), C.byref(C.c_int(int(j) + 1)))
        if nage <= c.nageclass:
            ref.L.f_calcfluxes(C.byref(C.c_int(nage)), C.byref(C.c_int(int(j) + 1)),
                               C.byref(C.c_float(np.float32(pre.xtra1[j]))), C.byref(C.c_float(np.float32(pre.ytra1[j]))),
                               C.byref(C.c_float(pre.ztra1[j])))


def _same_flux(f, fr):
    """species 1 carries unit masses: sums exact whatever the order of the atomics; species 2 decays"""
    assert f.shape == fr.shape
    assert np.array_equal(f[:, :, :, :, 0], fr[:, :, :, :, 0])
    assert np.array_equal(f[:, :, :, :, 1] > 0, fr[:, :, :, :, 1] > 0)
    assert np.allclose(f[:, :, :, :, 1], fr[:, :, :, :, 1], rtol=1e-5, atol=0.0)


@pytest.mark.parametrize("host_mode,sort_interval,jump", [(False, 0, False), (False, 1, True), (True, 0, True)])
def test_calcfluxes_and_partpos_average_match_the_reference_routines(host_mode, sort_interval, jump, monkeypatch):
    if host_mode:
        monkeypatch.setenv("FPB_HOST_CHUNKS", "3")
    cb, eng, ref, p, n = _setup(host_mode, sort_interval, jump)
    c = cb.cfg
    if not host_mode:
        eng.push_particles(p)
    cur = p
    for k in range(4):
        itime = k * 900
        pre = fb.Particles(c.maxpart, c.nspec); pre.numpart = n
        if host_mode:
            for f in ref.STATE:
                getattr(pre, f)[:n] = getattr(cur, f)[:n]
            pre.xmass1[:n] = cur.xmass1[:n]
            st = eng.step_host(cur, itime, 450, conc_weight=0.0)
            post = cur
        else:
            eng.pull_particles(pre)
            st = eng.step(itime, 450)
            post = fb.Particles(c.maxpart, c.nspec); post.numpart = n
            eng.pull_particles(post)
        assert st["n_active"] > 0
        _reference_hooks(ref, cb, pre, post, itime, n)
        if k == 1:   # an output interval in between: fetch + reset on both sides
            f = eng.fetch_fluxes(zero=True)
            assert f.sum() > 0
            _same_flux(f, ref.arr("flux"))
            ref.arr("flux")[...] = 0.0
    f, fr = eng.fetch_fluxes(), ref.arr("flux")
    _same_flux(f, fr)
    # both horizontal axes and both vertical directions were exercised, in both age classes
    assert sum(bool(f[i].sum() > 0) for i in range(6)) >= 4 and f[4].sum() > 0 and f[5].sum() > 0
    assert f[..., 0].sum() > 0 and f[..., 1].sum() > 0
    av = eng.fetch_partpos_average(n, zero=True)
    assert np.array_equal(av["npart_av"], ref.arr("npart_av")[:n]) and av["npart_av"].max() == 4
    for nm in AV:
        a, b = av[nm], ref.arr("part_av_" + nm)[:n]
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), (nm, int((a != b).sum()), float(np.abs(a - b).max()))
    again = eng.fetch_partpos_average(n)
    assert again["npart_av"].max() == 0 and all(not again[nm].any() for nm in AV)
    eng.close()


def test_hooks_need_their_fields():
    cb = cases.config_small(nrel=1, npart_each=64, ipout=3)
    eng = fb.Engine(cb)
    m0, m1 = cases.met_pair(cb)
    eng.fill_rannumb()
    eng.upload_met(1, m0); eng.upload_met(2, m1); eng.set_met_bracket((1, 2), (0, 10800))
    eng.push_particles(cases.seeded_particles(cb, 64))
    with pytest.raises(fb.FpbError, match="partpos_average"):
        eng.step(0, 450)
    with pytest.raises(fb.FpbError, match="iflux"):
        eng.fetch_fluxes()
    eng.close()


@pytest.mark.parametrize("linit_cond,regional", [(2, True), (1, False)])
def test_initial_cond_calc_matches_the_reference_routine(linit_cond, regional):
    """LINIT_COND (backward runs): initial_cond_calc for the particles the loop terminates -- leaving a regional
    domain (src/timemanager.f90:631, masses from before the step) or reaching the maximum age (:702) -- and, at the
    end, for every particle still active (:733-737), against src/initial_cond_calc.f90 called for the same
    particles on the engine's own positions.  linit_cond = 1 (division by rho at the particle) runs on the global
    grid: outside the domain the reference reads past its arrays."""
    kw = dict(nrel=4, npart_each=512, nspec=2, lage=(2700,), ioutputforeachrelease=1, ldirect=-1, linit_cond=linit_cond,
              math_mode=fb.MATH_FAST, rng_mode=fb.RNG_PHILOX_INDEX, sort_interval=1)
    if regional:   # 144 x 72 degrees: particles near the edges leave with the wind (nstop = 3)
        kw.update(dx=2.0, dy=2.0, xlon0=-60.0, ylat0=5.0, outlon0=-60.0, outlat0=5.0, numxgrid=70, numygrid=35,
                  dxout=2.0, dyout=2.0)
    cb = cases.config_small(**kw)
    c = cb.cfg
    assert c.lsynctime == -900 and bool(c.xglobal) == (not regional)
    n = 2048
    lat = (c.ylat0 + 2.0, c.ylat0 + (c.ny - 1) * c.dy - 2.0) if regional else (-70.0, 70.0)
    p = cases.seeded_particles(cb, n, zmax=6000.0, lat_range=lat, nspec=2)
    r = np.random.RandomState(12)
    if regional:
        p.xtra1[:n] = r.uniform(0.05, c.nx - 1.05, n)
        p.xtra1[:300] = r.uniform(0.001, 0.02, 300)                 # right at the western / eastern edge
        p.xtra1[300:600] = r.uniform(c.nx - 1.02, c.nx - 1.001, 300)
    p.xmass1[:n, 0] = 1.0
    p.xmass1[:n, 1] = 0.5
    p.itramem[:n:4] = 900              # a quarter reaches the maximum age two steps earlier
    p.itramem[1:n:4] = -1800           # a quarter is young enough to be active at the end
    mets = (fb.MetFields(cb).synth(0), fb.MetFields(cb).synth(-10800))
    eng = fb.Engine(cb)
    eng.fill_rannumb()
    eng.upload_met(1, mets[0]); eng.upload_met(2, mets[1]); eng.set_met_bracket((1, 2), (0, -10800))
    ref = ref_api.Ref(cb, maxrand=1000)
    ref.upload_met(1, mets[0]); ref.upload_met(2, mets[1]); ref.set_met_bracket((1, 2), (0, -10800))
    ref.set("linit_cond", linit_cond)
    ref.arr("init_cond")[...] = 0.0
    eng.push_particles(p)
    n_stop = n_age = 0

    def ref_calc(post, mass, rows, it):
        q = fb.Particles(c.maxpart, c.nspec); q.numpart = n
        for f in ref.STATE:
            getattr(q, f)[:n] = getattr(post, f)[:n]
        q.xmass1[:n] = mass[:n]
        q.itra1[:n] = fb.ITRA_DEAD
        q.itra1[rows] = it
        ref.push_state(q)
        for j in rows:
            ref.L.f_initial_cond_calc(C.byref(C.c_int(it)), C.byref(C.c_int(int(j) + 1)))

    for k in range(4):
        itime = -900 * k
        pre = fb.Particles(c.maxpart, c.nspec); pre.numpart = n
        eng.pull_particles(pre)
        eng.step(itime, 450)
        post = fb.Particles(c.maxpart, c.nspec); post.numpart = n
        eng.pull_particles(post)
        died = (pre.itra1[:n] == itime) & (post.itra1[:n] == fb.ITRA_DEAD)
        x, y = post.xtra1[:n], post.ytra1[:n]
        outside = (x < 0.0) | (x >= np.float32(c.nxmin1)) | (y < 0.0) | (y > np.float32(c.nymin1))
        stop = np.nonzero(died & outside)[0]
        age = np.nonzero(died & ~outside)[0]
        assert (np.abs(itime - 900 - post.itramem[age]) >= 2700).all()
        n_stop += stop.size; n_age += age.size
        ref_calc(post, pre.xmass1, stop, itime)
        ref_calc(post, post.xmass1, age, itime - 900)
    got, want = eng.fetch_init_cond(), ref.arr("init_cond")
    assert got.shape == want.shape and want.sum() > 0
    assert np.array_equal(got > 0, want > 0) and np.allclose(got, want, rtol=1e-5, atol=0.0)
    assert n_age > 400 and (n_stop > 50) == regional, (n_stop, n_age)
    # the end of the run: everything still active
    itime = -3600
    post = fb.Particles(c.maxpart, c.nspec); post.numpart = n
    eng.pull_particles(post)
    alive = np.nonzero(post.itra1[:n] == itime)[0]
    assert alive.size > 100
    eng.initial_cond_final(itime)
    ref_calc(post, post.xmass1, alive, itime)
    got, want = eng.fetch_init_cond(zero=True), ref.arr("init_cond")
    assert np.array_equal(got > 0, want > 0) and np.allclose(got, want, rtol=1e-5, atol=0.0)
    assert not eng.fetch_init_cond().any()
    eng.close()
