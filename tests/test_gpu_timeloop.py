"""The host time loop (fpbh_timemanager, src/timemanager.f90:152-729) with the callers either side of
the particle loop on the device: domain filling with the inflow boundary and particle splitting
against the oracle run through the same loop, and a run whose wind fields are model-level fields
transformed by the engine (calcpar + verttransform) with convective mixing every step."""
import numpy as np
import pytest

import flexpart_b200 as fb
import cases
from oracle_api import Oracle

pytestmark = pytest.mark.gpu

INT_FIELDS = ("itra1", "npoint", "nclass", "idt", "itramem", "itrasplit", "cbt")
FLOAT_FIELDS = ("xtra1", "ytra1", "ztra1", "uap", "ucp", "uzp", "us", "vs", "ws")


def test_domainfill_box_with_boundary_and_splitting_matches_the_oracle_run():
    """init_domainfill at itime 0, boundcond_domainfill every later step (a limited box: particles leave,
    new ones enter through the four boundaries), splitting at loutend once itsplit is reached
    (src/timemanager.f90:230-241,472-503): strict math + reference RNG streams, the engine and the
    oracle driven by the same loop end with bit-identical particles and grids."""
    cb = cases.config_small(nrel=1, npart_each=40000, maxpart=120000, mdomainfill=1, nclassunc=2,
                            math_mode=fb.MATH_STRICT, scatter_mode=fb.SCATTER_DETERMINISTIC, sort_interval=2)
    c = cb.cfg
    rel = fb.Releases(cb, lon1=[-60.0], lon2=[70.0], lat1=[-30.0], lat2=[45.0], z1=[0.0], z2=[100.0], start=[0],
                      end=[0], itsplit=5400)
    run = fb.RunSpec(ideltas=12 * 900)
    eng, ora = fb.Engine(cb), Oracle(cb)
    eng.fill_rannumb(); ora.fill_rannumb()
    rg, og = fb.timemanager(cb, rel, run, eng.vtable(device_release=True))
    ro, oo = fb.timemanager(cb, rel, run, ora.vtable())
    assert (rg.syncs, rg.numpart_final, rg.boundary_particles, rg.split_calls, rg.particle_steps) == \
           (ro.syncs, ro.numpart_final, ro.boundary_particles, ro.split_calls, ro.particle_steps)
    assert ro.boundary_particles > 10 and ro.split_calls >= 1 and ro.numpart_final > 70000
    n = ro.numpart_final
    pg, po = fb.Particles(c.maxpart, 1), fb.Particles(c.maxpart, 1)
    pg.numpart = po.numpart = n
    eng.pull_particles(pg); ora.pull_particles(po)
    for f in INT_FIELDS + FLOAT_FIELDS:
        assert np.array_equal(getattr(pg, f)[:n], getattr(po, f)[:n]), f
    assert np.array_equal(pg.xmass1[:n], po.xmass1[:n])
    assert len(og) == len(oo) >= 2
    for a, b in zip(og, oo):
        assert a["itime"] == b["itime"] and a["outnum"] == b["outnum"]
        assert np.array_equal(a["gridunc"], b["gridunc"]) and b["gridunc"].sum() > 0
    eng.close()


@pytest.mark.parametrize("ldirect", [1, -1])
def test_raw_met_and_convection_in_the_loop_equal_the_same_calls_made_by_hand(ldirect):
    """fpbh_run::met_raw + lconvection: every new wind field is a model-level field that the engine
    turns into a met slot (fpb_calcpar_verttransform), convmix runs after the release (forward) or
    before the new fields (backward, src/timemanager.f90:183-193,258-263).  The loop's result equals
    the same sequence of C-ABI calls issued from here, and convection has moved particles."""
    nuvz = 40
    akm, bkm, akz, bkz, nconvlev = fb.synth_hybrid_levels(nuvz)
    kw = dict(nrel=4, npart_each=3000, nz=nuvz, ldirect=ldirect, math_mode=fb.MATH_STRICT)
    cb0 = cases.config_small(**kw, height=fb.synth_heights(nuvz))
    hh, _ = fb.verttransform_heights(cb0, nuvz, akz, bkz, fb.synth_rawmet(cb0, nuvz, akz, bkz, 0))
    cb = cases.config_small(**kw, height=hh)
    c = cb.cfg
    rel = cases.releases_boxes(cb, seed=4, zmax=3000.0, lat_range=(-35.0, 35.0))
    nsteps = 8
    # output window far away: no sampling inside the run (the loop's schedule is tested elsewhere)
    run = fb.RunSpec(ideltas=ldirect * nsteps * 900, loutstep=360000, loutaver=900, loutsample=900, met_interval=3600,
                     ldirect=ldirect, met_raw=True, lconvection=True)
    eng = fb.Engine(cb)
    eng.fill_rannumb()
    r, _ = fb.timemanager(cb, rel, run, eng.vtable(device_release=True))
    assert r.syncs == nsteps and r.convmix_calls in (nsteps, nsteps + 1, nsteps - 1) and r.convecting_columns > 0
    n = r.numpart_final
    p1 = fb.Particles(c.maxpart, 1); p1.numpart = n
    eng.pull_particles(p1)
    eng.close()

    # the same calls by hand
    e2 = fb.Engine(cb)
    e2.fill_rannumb()
    e2.set_vertical(nuvz, akm, bkm, akz, bkz)
    e2.set_convection(nuvz, c.nzmax, nconvlev, akz, bkz, akm, bkm)
    e2.set_releases(rel)
    memind, memtime = [1, 2], [0, ldirect * 3600]
    e2.calcpar_verttransform(1, fb.synth_rawmet(cb, nuvz, akz, bkz, memtime[0]))
    e2.calcpar_verttransform(2, fb.synth_rawmet(cb, nuvz, akz, bkz, memtime[1]))
    loutnext = ldirect * 360000 // 2
    calls = 0
    for k in range(nsteps + 1):
        itime = ldirect * k * 900
        if ldirect == -1 and itime < 0:
            e2.convmix(itime); calls += 1
        if ldirect * memtime[1] <= ldirect * itime:          # getfields: next field into the older slot
            memind = [memind[1], memind[0]]
            memtime = [memtime[1], memtime[1] + ldirect * 3600]
            e2.calcpar_verttransform(memind[1], fb.synth_rawmet(cb, nuvz, akz, bkz, memtime[1]))
        e2.set_met_bracket(tuple(memind), tuple(memtime))
        e2.release_particles(itime)
        if ldirect == 1:
            e2.convmix(itime); calls += 1
        if k == nsteps:
            break
        e2.step(itime, itime - (loutnext - ldirect * 360000) if itime < loutnext else itime - loutnext)
    assert calls == r.convmix_calls
    p2 = fb.Particles(c.maxpart, 1); p2.numpart = n
    e2.pull_particles(p2)
    for f in INT_FIELDS + FLOAT_FIELDS:
        assert np.array_equal(getattr(p1, f)[:n], getattr(p2, f)[:n]), f
    # convection is what it claims to be: the same run without it ends elsewhere
    e3 = fb.Engine(cb)
    e3.fill_rannumb()
    run3 = fb.RunSpec(ideltas=ldirect * nsteps * 900, loutstep=360000, loutaver=900, loutsample=900, met_interval=3600,
                      ldirect=ldirect, met_raw=True, lconvection=False)
    r3, _ = fb.timemanager(cb, rel, run3, e3.vtable(device_release=True))
    p3 = fb.Particles(c.maxpart, 1); p3.numpart = n
    e3.pull_particles(p3)
    # (redist draws from the same ran3 stream as the particle loop: every later draw shifts)
    moved = (p3.ztra1[:n] != p1.ztra1[:n]).mean()
    assert r3.convmix_calls == 0 and moved > 0.001, moved
    e2.close(); e3.close()
