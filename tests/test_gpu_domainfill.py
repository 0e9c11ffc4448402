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
    # global domain: boundcond_domainfill returns at once (src/boundcond_domainfill.f90:54);
    # the inflow boundary of a limited domain is refused, not silently skipped
    if io["gdomainfill"]:
        eng.boundcond_domainfill(900)
    else:
        with pytest.raises(fb.FpbError, match="limited domain"):
            eng.boundcond_domainfill(900)
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
