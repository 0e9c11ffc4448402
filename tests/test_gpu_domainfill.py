"""Device-side init_domainfill (src/init_domainfill.f90:55-283; SURVEY.md 8f rank 2, BASELINE
configs[4]) against the oracle's restatement, which tests/test_ref_transpiled.py pins bit for bit
against the reference's own routine."""
import numpy as np
import pytest

import flexpart_b200 as fb
import cases
from oracle_api import Oracle

pytestmark = pytest.mark.gpu

FIELDS = ("xtra1", "ytra1", "ztra1", "itra1", "itramem", "npoint", "nclass", "idt", "itrasplit")


def _box(c, lon0, lat0, lon1, lat1):
    return [float(np.float32(v)) for v in ((lon0 - c.xlon0) / c.dx, (lat0 - c.ylat0) / c.dy,
                                           (lon1 - c.xlon0) / c.dx, (lat1 - c.ylat0) / c.dy)]


@pytest.mark.parametrize("math_mode", [fb.MATH_STRICT, fb.MATH_FAST])
@pytest.mark.parametrize("box,npart1", [((-180.0, -90.0, 180.0, 90.0), 60000),
                                        ((-180.0, -90.0, 180.0, 90.0), 2500),    # thin columns: random heights
                                        ((-40.0, 10.0, 65.0, 72.5), 12000)])     # limited domain
def test_reference_stream_is_bit_identical_to_the_oracle(box, npart1, math_mode):
    cb = cases.config_small(nrel=1, npart_each=npart1, maxpart=npart1 + 2000, mdomainfill=1, nclassunc=3,
                            math_mode=math_mode, sort_interval=1)
    c = cb.cfg
    m0, m1 = cases.met_pair(cb)
    eng, ora = fb.Engine(cb), Oracle(cb)
    for e in (eng, ora):
        e.upload_met(1, m0); e.upload_met(2, m1); e.set_met_bracket((1, 2), (0, 10800))
    pts = _box(c, *box)
    no, io = ora.init_domainfill(pts)
    ng, ig = eng.init_domainfill(pts)
    assert ng == no and no > 0.9 * npart1
    for k in ("nx_we", "ny_sn", "gdomainfill", "numcolumn", "numparttot"):
        assert ig[k] == io[k], k
    assert np.float32(ig["colmasstotal"]).tobytes() == np.float32(io["colmasstotal"]).tobytes()
    assert np.float32(ig["xmassperparticle"]).tobytes() == np.float32(io["xmassperparticle"]).tobytes()
    pg, po = fb.Particles(c.maxpart, 1), fb.Particles(c.maxpart, 1)
    pg.numpart = po.numpart = no
    eng.pull_particles(pg); ora.pull_particles(po)
    for f in FIELDS:
        assert np.array_equal(getattr(pg, f)[:no], getattr(po, f)[:no]), f
    assert np.array_equal(pg.xmass1[:no], po.xmass1[:no])
    # global domain: boundcond_domainfill returns at once (src/boundcond_domainfill.f90:54)
    if io["gdomainfill"]:
        assert eng.boundcond_domainfill(900) == (no, 0)
    # the created particles step like any others (initialize runs for them: itramem == itime == 0)
    fill = lambda e: e.fill_rannumb()
    fill(eng); fill(ora)
    if math_mode == fb.MATH_STRICT:
        for e in (eng, ora):
            e.conccalc(0, 1.0)
        sg, so = eng.step(0), ora.step(0)
        assert sg == so and sg["n_init"] == sg["n_active"] == int((po.itra1[:no] == 0).sum())
        eng.pull_particles(pg); ora.pull_particles(po)
        for f in FIELDS + ("uap", "ucp", "uzp", "us", "vs", "ws"):
            assert np.array_equal(getattr(pg, f)[:no], getattr(po, f)[:no]), f
    eng.close()


@pytest.mark.parametrize("math_mode", [fb.MATH_STRICT, fb.MATH_FAST])
def test_boundcond_domainfill_limited_box_bit_identical(math_mode):
    """boundcond_domainfill on the device (src/boundcond_domainfill.f90:54-560) interleaved with the
    particle loop and the cell sort: terminations, accumulated boundary masses, the slots the new
    particles take and their ran1 positions are those of the oracle (pinned against the reference's
    routine in tests/test_ref_transpiled.py), call after call."""
    npart1 = 150000
    cb = cases.config_small(nrel=1, npart_each=npart1, maxpart=npart1 + 30000, mdomainfill=1, nclassunc=3,
                            math_mode=math_mode, sort_interval=1)
    c = cb.cfg
    m0, m1 = cases.met_pair(cb)
    eng, ora = fb.Engine(cb), Oracle(cb)
    for e in (eng, ora):
        e.fill_rannumb()
        e.upload_met(1, m0); e.upload_met(2, m1); e.set_met_bracket((1, 2), (0, 10800))
    pts = _box(c, -60.0, -30.0, 70.0, 45.0)
    no, io = ora.init_domainfill(pts)
    ng, ig = eng.init_domainfill(pts)
    assert ng == no and ig["gdomainfill"] == 0
    created = terminated = 0
    exact = math_mode == fb.MATH_STRICT
    for k in range(6):
        itime = k * c.lsynctime
        if not exact and k:     # fast math: re-inject the oracle's particle state before every call
            q = fb.Particles(c.maxpart, 1); q.numpart = ora.numpart()
            ora.pull_particles(q)
            eng.push_particles(q)
        po0 = fb.Particles(c.maxpart, 1); po0.numpart = ora.numpart()
        ora.pull_particles(po0)
        mo = ora.boundcond_domainfill(itime)
        n, mg = eng.boundcond_domainfill(itime)
        assert mg == mo and n == ora.numpart(), (k, mg, mo, n, ora.numpart())
        created += mo
        pg, po = fb.Particles(c.maxpart, 1), fb.Particles(c.maxpart, 1)
        pg.numpart = po.numpart = n
        eng.pull_particles(pg); ora.pull_particles(po)
        terminated += int(((po0.itra1[:po0.numpart] == itime) & (po.itra1[:po0.numpart] == fb.ITRA_DEAD)).sum())
        for f in FIELDS:
            assert np.array_equal(getattr(pg, f)[:n], getattr(po, f)[:n]), (k, f)
        assert np.array_equal(pg.xmass1[:n], po.xmass1[:n]), k
        sg, so = eng.step(itime), ora.step(itime)
        assert sg["n_active"] == so["n_active"] and sg["n_init"] == so["n_init"], k
        if exact:
            eng.pull_particles(pg); ora.pull_particles(po)
            for f in FIELDS + ("uap", "ucp", "uzp", "us", "vs", "ws"):
                assert np.array_equal(getattr(pg, f)[:n], getattr(po, f)[:n]), (k, f)
    assert created > 300 and terminated > 100, (created, terminated)
    eng.close()


def test_boundcond_domainfill_philox_rank_partition():
    """Production RNG: every rank accumulates the same boundary fluxes and keeps every N-th new
    particle; the union over the ranks is the one-rank result (particles keyed by their global
    count), and the new particles sit on the boundary they entered through."""
    npart1 = 300000
    base = dict(nrel=1, npart_each=npart1, maxpart=npart1 + 30000, mdomainfill=1, rng_mode=fb.RNG_PHILOX_INDEX)
    m0 = m1 = None

    def run(cb):
        nonlocal m0, m1
        if m0 is None:
            m0, m1 = cases.met_pair(cb)
        c = cb.cfg
        eng = fb.Engine(cb)
        eng.upload_met(1, m0); eng.upload_met(2, m1); eng.set_met_bracket((1, 2), (0, 10800))
        n, info = eng.init_domainfill(_box(c, -60.0, -30.0, 70.0, 45.0))
        new = {}
        for k in range(4):
            itime = k * c.lsynctime
            n, m = eng.boundcond_domainfill(itime)
            p = fb.Particles(c.maxpart, 1); p.numpart = n
            eng.pull_particles(p)
            sel = np.nonzero((p.itramem[:n] == itime) & (p.itra1[:n] == itime) & (p.npoint[:n] > info["numparttot"]))[0] \
                if k else np.nonzero((p.itra1[:n] == 0) & (p.npoint[:n] > info["numparttot"]))[0]
            assert len(sel) == m, (k, len(sel), m)
            for s in sel:
                new[int(p.npoint[s])] = (float(p.xtra1[s]), float(p.ytra1[s]), float(p.ztra1[s]), int(p.nclass[s]),
                                         float(p.xmass1[s, 0]), itime)
            live = p.itra1[:n] == itime          # the particle loop would move the clock on
            p.itra1[:n][live] = itime + c.lsynctime
            eng.push_particles(p)
        eng.close()
        return info, new

    info, one = run(cases.config_small(**base))
    assert len(one) > 500
    nx_we, ny_sn = info["nx_we"], info["ny_sn"]
    for x, y, z, nc, xm, _ in one.values():
        assert x in (float(nx_we[0]), float(nx_we[1])) or y in (float(ny_sn[0]), float(ny_sn[1]))
        assert nx_we[0] - 0.5 <= x <= nx_we[1] + 0.5 and ny_sn[0] - 0.5 <= y <= ny_sn[1] + 0.5 and z > 0.0
        assert np.float32(xm) == np.float32(info["xmassperparticle"])
    union = {}
    for r in range(3):
        _, part = run(cases.config_small(**dict(base, part_id_stride=3, part_id_offset=r)))
        assert not (set(part) & set(union))
        assert all((g - 1) % 3 == r for g in part)
        union.update(part)
    assert union == one


def test_philox_fill_properties_and_rank_partition():
    """Production RNG: same columns, counts, heights and masses as the reference stream (only the
    uniforms differ), positions inside their cells; N ranks with part_id_stride = N each hold
    every N-th particle of the one-rank result, bit for bit."""
    npart1 = 200000
    base = dict(nrel=1, npart_each=npart1, maxpart=npart1 + 2000, mdomainfill=1, rng_mode=fb.RNG_PHILOX_INDEX)
    cb = cases.config_small(**base)
    c = cb.cfg
    m0, m1 = cases.met_pair(cb)
    pts = _box(c, -180.0, -90.0, 180.0, 90.0)

    def fill(cb):
        eng = fb.Engine(cb)
        eng.upload_met(1, m0); eng.upload_met(2, m1); eng.set_met_bracket((1, 2), (0, 10800))
        n, info = eng.init_domainfill(pts)
        p = fb.Particles(cb.cfg.maxpart, 1); p.numpart = n
        eng.pull_particles(p)
        eng.close()
        return n, info, p

    n1, info, p1 = fill(cb)
    ora = Oracle(cases.config_small(**dict(base, rng_mode=fb.RNG_REFERENCE)))
    ora.upload_met(1, m0); ora.upload_met(2, m1)
    no, io = ora.init_domainfill(pts)
    po = fb.Particles(c.maxpart, 1); po.numpart = no
    ora.pull_particles(po)
    # (numpart drops the dead particles at the end of the arrays, and which particles of the last,
    # polar row fall outside the domain depends on the uniforms)
    assert abs(n1 - no) <= 64 and info["numparttot"] == io["numparttot"] and info["numcolumn"] == io["numcolumn"]
    full_n1 = n1
    n1 = min(n1, no)
    live = p1.itra1[:n1] == 0
    assert live.mean() > 0.999
    # particle g of both runs sits in the same column with the same mass; dense columns (> 20
    # particles) also share the pressure-equidistant height
    assert np.array_equal(p1.xmass1[:n1], po.xmass1[:n1])
    assert np.array_equal(p1.npoint[:n1], po.npoint[:n1])
    assert (np.abs(p1.xtra1[:n1] - po.xtra1[:n1]) <= 1.0).all() and (np.abs(p1.ytra1[:n1] - po.ytra1[:n1]) <= 1.0).all()
    assert (p1.ztra1[:n1] == po.ztra1[:n1]).mean() > 0.95
    assert not np.array_equal(p1.xtra1[:n1], po.xtra1[:n1])
    # uniform inside the cell: mean offset from the cell centre ~ 0, spread ~ 1/sqrt(12)
    fx = p1.xtra1[:n1][live] - np.round(p1.xtra1[:n1][live])
    assert abs(fx.mean()) < 0.01 and abs(fx.std() - 12 ** -0.5) < 0.01
    total = float(p1.xmass1[:n1, 0].astype(np.float64).sum())
    assert abs(total / info["colmasstotal"] - 1.0) < 2e-3
    # 3 "ranks"
    got = 0
    for r in range(3):
        nr, ir, pr = fill(cases.config_small(**dict(base, part_id_stride=3, part_id_offset=r)))
        assert ir["numparttot"] == info["numparttot"]
        idx = r + 3 * np.arange(nr)          # local slot s holds global particle r + 3 s
        idx = idx[idx < full_n1]
        for f in FIELDS:
            assert np.array_equal(getattr(pr, f)[:len(idx)], getattr(p1, f)[idx]), (r, f)
        assert np.array_equal(pr.xmass1[:len(idx)], p1.xmass1[idx])
        got += len(idx)
    assert full_n1 - 6 <= got <= full_n1     # (each rank drops its own dead tail)


def test_full_size_fill_conserves_the_air_mass():
    """configs[4] geometry: 721 x 361 x 138, 4 M particles; total particle mass = air mass of the
    atmosphere (5.1e18 kg on the synthetic fields), every live particle inside the domain and below
    the model top, mass per particle uniform within a column's rounding."""
    n = 4_000_000
    cb = fb.make_config(nx=721, ny=361, nz=138, dx=0.5, dy=0.5, xlon0=-180.0, ylat0=-90.0, lsynctime=900, ctl=-5.0,
                        ifine=4, outlon0=-180.0, outlat0=-90.0, numxgrid=720, numygrid=360, dxout=0.5, dyout=0.5,
                        outheights=(100.0, 1000.0, 5000.0, 100000.0), lage=(86400 * 20,), ioutputforeachrelease=0,
                        npart=(n,), nspec=1, maxpart=n + 200000, mdomainfill=1, rng_mode=fb.RNG_PHILOX_INDEX,
                        sort_interval=1)
    c = cb.cfg
    m0 = fb.MetFields(cb).synth(0)
    eng = fb.Engine(cb)
    eng.upload_met(1, m0); eng.upload_met(2, fb.MetFields(cb).synth(10800)); eng.set_met_bracket((1, 2), (0, 10800))
    num, info = eng.init_domainfill(_box(c, -180.0, -90.0, 180.0, 90.0))
    assert info["gdomainfill"] == 1 and abs(num - 0.999 * n) < 0.01 * n and info["numparttot"] >= num
    p = fb.Particles(c.maxpart, 1); p.numpart = num
    eng.pull_particles(p)
    live = p.itra1[:num] == 0
    assert live.mean() > 0.9999
    x, y, z = p.xtra1[:num][live], p.ytra1[:num][live], p.ztra1[:num][live]
    assert x.min() >= 0 and x.max() < c.nxmin1 and y.min() >= 0 and y.max() < c.nymin1
    assert z.min() >= 0 and z.max() <= cb.height[c.nz - 1]
    mass = p.xmass1[:num, 0].astype(np.float64)
    assert abs(mass.sum() / info["colmasstotal"] - 1.0) < 1e-3 and 4.5e18 < info["colmasstotal"] < 6e18
    assert abs(np.median(mass) / info["xmassperparticle"] - 1.0) < 0.05
    eng.fill_rannumb()
    eng.conccalc(0, 1.0)
    st = eng.step(0)
    assert st["n_active"] == int(live.sum()) == st["n_init"] and st["n_terminated"] == 0
    g = eng.fetch_grids()["gridunc"]
    assert abs(float(g.astype(np.float64).sum()) / mass[live].sum() - 1.0) < 1e-4
    eng.close()
