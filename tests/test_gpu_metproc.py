"""calcpar + verttransform_ecmwf on the device (fpb_calcpar_verttransform)
through the C ABI against the reference's own routines (oracle/_ref, src/calcpar.f90 +
src/verttransform_ecmwf.f90 run from their sources), and the slot it builds against the same fields
uploaded from the host."""
import numpy as np
import pytest

import flexpart_b200 as fb
import cases
import conv_cases
import met_cases
import ref_api
from metproc_common import reference_run, compare_fields
from oracle_api import Oracle

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not ref_api.available(), reason="oracle/_ref/libflexref.so not built")]

FIELDS3 = ("uu", "vv", "ww", "rho", "drhodz", "tt", "qv", "pv", "uupol", "vvpol", "clouds")
FIELDS2 = ("hmix", "ustar", "wstar", "oli", "tropopause")


def _device_fields(cb, eng, nuvz, slot=1):
    c = cb.cfg
    f = eng.fetch_met(slot)
    got = {}
    for nm in FIELDS3:
        if nm in f:
            got[nm] = np.ascontiguousarray(np.transpose(f[nm][:c.nx, :c.ny, :nuvz], (2, 1, 0)))
    for nm in FIELDS2:
        got[nm] = np.ascontiguousarray(f[nm][:c.nx, :c.ny].T)
    return got, f


@pytest.mark.parametrize("nx,ny,nuvz,seed", [(73, 37, 40, 1), (181, 91, 138, 3)])
def test_device_slot_is_bit_identical_to_the_reference_routines(nx, ny, nuvz, seed):
    kw = dict(nrel=1, npart_each=8, nx=nx, ny=ny, nz=nuvz, wetdepspec=(1,), weta_gas=(2.0e-5,), wetb_gas=(0.62,))
    cb0 = cases.config_small(**kw, height=fb.synth_heights(nuvz))
    akm, bkm, akz, bkz, _ = conv_cases.hybrid_levels(nuvz)
    raw = met_cases.raw_fields(cb0, akz, bkz, nuvz, seed=seed)
    ref, height, pvh = reference_run(cb0, raw, akm, bkm, akz, bkz, nuvz)
    hh, _ = fb.verttransform_heights(cb0, nuvz, akz[1:], bkz[1:], raw)
    assert np.array_equal(hh, height)
    cb = cases.config_small(**kw, height=hh)
    eng = fb.Engine(cb)
    eng.set_vertical(nuvz, akm[1:], bkm[1:], akz[1:], bkz[1:])
    ms = eng.calcpar_verttransform(1, raw)       # (pvh not handed in: calcpv runs on the device)
    assert ms > 0.0
    got, _ = _device_fields(cb, eng, nuvz)
    bad = compare_fields(cb, ref, got, nuvz)
    assert not bad, bad
    assert set(np.unique(got["clouds"])) == {0, 1, 2, 3, 4, 5}
    eng.close()


@pytest.mark.parametrize("sumclouds", [0, 1])
def test_readclouds_slot_is_bit_identical_to_the_reference_routines(sumclouds):
    """cloud water read from the input (src/verttransform_ecmwf.f90:610-681): classes and column total ctwc"""
    nuvz = 40
    kw = dict(nrel=1, npart_each=8, nz=nuvz, wetdepspec=(1,), weta_gas=(2.0e-5,), wetb_gas=(0.62,), readclouds=1)
    cb0 = cases.config_small(**kw, height=fb.synth_heights(nuvz))
    akm, bkm, akz, bkz, _ = conv_cases.hybrid_levels(nuvz)
    raw = met_cases.raw_fields(cb0, akz, bkz, nuvz, seed=3)
    ref, height, pvh = reference_run(cb0, raw, akm, bkm, akz, bkz, nuvz, readclouds=1, sumclouds=sumclouds)
    cb = cases.config_small(**kw, height=height)
    c = cb.cfg
    eng = fb.Engine(cb)
    eng.set_vertical(nuvz, akm[1:], bkm[1:], akz[1:], bkz[1:])
    dev_raw = dict(raw)
    if sumclouds:
        dev_raw["clwch"] = raw["clwch"] + raw["ciwch"]
        dev_raw["ciwch"] = None
    eng.calcpar_verttransform(1, dev_raw)
    got, f = _device_fields(cb, eng, nuvz)
    got["ctwc"] = np.ascontiguousarray(f["ctwc"][:c.nx, :c.ny].T)
    bad = compare_fields(cb, ref, got, nuvz)
    assert not bad, bad
    assert set(np.unique(got["clouds"])) == {0, 2, 3, 4, 5} and (got["ctwc"] > 0).mean() > 0.5
    eng.close()


@pytest.mark.parametrize("wet", [True, False])
def test_nested_grid_slot_is_bit_identical_to_the_reference_routines(wet):
    """fpb_calcpar_verttransform_nest against calcpar_nests + verttransform_nests + calcpv_nests run from their
    sources (src/getfields.f90:131-134), and a particle step through the nest on the slots built that way"""
    from metproc_common import nest_configs, reference_run_nest, compare_fields_nest, NEST
    nuvz = 40
    kw = dict(nrel=1, npart_each=512, nz=nuvz, math_mode=fb.MATH_STRICT)
    if wet:
        kw.update(wetdepspec=(1,), weta_gas=(2.0e-5,), wetb_gas=(0.62,))
    cb0, _ = nest_configs(**kw, height=fb.synth_heights(nuvz))
    akm, bkm, akz, bkz, _ = conv_cases.hybrid_levels(nuvz)
    raw = met_cases.raw_fields(cb0, akz, bkz, nuvz, seed=1)
    ref, height, _ = reference_run(cb0, raw, akm, bkm, akz, bkz, nuvz)
    cb, cbn = nest_configs(**kw, height=height)
    rawn = met_cases.raw_fields(cbn, akz, bkz, nuvz, seed=4)
    reference_run_nest(ref, cb, rawn, nuvz)
    g = NEST
    eng = fb.Engine(cb)
    eng.set_vertical(nuvz, akm[1:], bkm[1:], akz[1:], bkz[1:])
    for slot in (1, 2):
        eng.calcpar_verttransform(slot, raw)
        ms = eng.calcpar_verttransform_nest(slot, 1, rawn, g["dxn"], g["dyn"], g["xlon0n"], g["ylat0n"])
        assert ms > 0.0
    f = eng.fetch_met(1, nest=1)
    names3 = ("uu", "vv", "ww", "rho", "drhodz") + (("tt", "clouds") if wet else ())
    got = {nm: np.ascontiguousarray(np.transpose(f[nm][:g["nxn"], :g["nyn"], :nuvz], (2, 1, 0))) for nm in names3}
    for nm in FIELDS2:
        got[nm] = np.ascontiguousarray(f[nm][:g["nxn"], :g["nyn"]].T)
    bad = compare_fields_nest(ref, got, nuvz)
    assert not bad, bad
    # particles inside the nest step on the device-built nest slots without leaving the fields
    eng.fill_rannumb()
    eng.set_met_bracket((1, 2), (0, 10800))
    p = cases.seeded_particles(cb, 512, zmax=3000.0, lat_range=(g["ylat0n"] + 3.0, g["ylat0n"] + g["nyn"] - 4.0))
    p.xtra1[:512] = (np.random.RandomState(3).uniform(g["xlon0n"] + 3.0, g["xlon0n"] + g["nxn"] - 4.0, 512) - cb.cfg.xlon0) / cb.cfg.dx
    eng.push_particles(p)
    st = eng.step(0, 450)
    assert st["n_active"] == 512 and st["n_nonfinite"] == 0
    q = fb.Particles(cb.cfg.maxpart, 1); q.numpart = 512
    eng.pull_particles(q)
    assert np.isfinite(q.xtra1[:512]).all() and np.isfinite(q.ztra1[:512]).all() and (q.itra1[:512] == 900).all()
    eng.close()


def test_particles_step_alike_on_device_built_and_uploaded_slots():
    """The slot fpb_calcpar_verttransform builds is what fpb_upload_met makes of the same fields: the
    particle loop, wet deposition and conccalc give bit-identical results on both, and the oracle on
    the fetched fields agrees with them (strict math)."""
    nuvz = 40
    kw = dict(nrel=2, npart_each=2000, nz=nuvz, wetdepspec=(1,), weta_gas=(2.0e-5,), wetb_gas=(0.62,),
              henry=(1.0e-2,), math_mode=fb.MATH_STRICT, scatter_mode=fb.SCATTER_DETERMINISTIC)
    cb0 = cases.config_small(**kw, height=fb.synth_heights(nuvz))
    akm, bkm, akz, bkz, _ = conv_cases.hybrid_levels(nuvz)
    raws = [met_cases.raw_fields(cb0, akz, bkz, nuvz, seed=1, tshift=t) for t in (0.0, 1.5)]
    hh, _ = fb.verttransform_heights(cb0, nuvz, akz[1:], bkz[1:], raws[0])
    cb = cases.config_small(**kw, height=hh)
    c = cb.cfg
    dev, up, ora = fb.Engine(cb), fb.Engine(cb), Oracle(cb)
    dev.set_vertical(nuvz, akm[1:], bkm[1:], akz[1:], bkz[1:])
    for slot, raw in enumerate(raws, start=1):
        dev.calcpar_verttransform(slot, raw)
        f = dev.fetch_met(slot)
        m = fb.MetFields(cb)
        for nm in ("uu", "vv", "ww", "rho", "drhodz", "tt", "uupol", "vvpol", "hmix", "ustar", "wstar", "oli",
                   "tropopause", "clouds"):
            getattr(m, nm)[...] = f[nm]
        for nm in ("lsprec", "convprec", "tcc"):
            getattr(m, nm)[...] = raw[nm]
        up.upload_met(slot, m)
        ora.upload_met(slot, m)
    p = cases.seeded_particles(cb, 4000, zmax=6000.0, lat_range=(-85.0, 85.0))
    for e in (dev, up, ora):
        e.fill_rannumb()
        e.set_met_bracket((1, 2), (0, 10800))
        e.push_particles(p)
    for k in range(4):
        itime = k * 900
        for e in (dev, up, ora):
            if k:
                e.wetdepo(itime, 900, 450)
            e.conccalc(itime, 1.0)
        sd, su, so = dev.step(itime, 450), up.step(itime, 450), ora.step(itime, 450)
        assert sd == su == so, (k, sd, su, so)
    n = 4000
    pd, pu, po = (fb.Particles(c.maxpart, 1) for _ in range(3))
    for q in (pd, pu, po):
        q.numpart = n
    dev.pull_particles(pd); up.pull_particles(pu); ora.pull_particles(po)
    for f in ("xtra1", "ytra1", "ztra1", "itra1", "idt", "uap", "ucp", "uzp", "us", "vs", "ws", "cbt"):
        assert np.array_equal(getattr(pd, f)[:n], getattr(pu, f)[:n]), f
        assert np.array_equal(getattr(pd, f)[:n], getattr(po, f)[:n]), f
    assert np.array_equal(pd.xmass1[:n], pu.xmass1[:n]) and np.array_equal(pd.xmass1[:n], po.xmass1[:n])
    gd, gu, go = dev.fetch_grids(), up.fetch_grids(), ora.fetch_grids()
    assert gd["gridunc"].sum() > 0 and np.array_equal(gd["gridunc"], gu["gridunc"]) and np.array_equal(gd["gridunc"], go["gridunc"])
    wd, wu = dev.fetch_wetgrids(), up.fetch_wetgrids()
    assert wd["wetgridunc"].sum() > 0 and np.array_equal(wd["wetgridunc"], wu["wetgridunc"])
    for e in (dev, up):
        e.close()


def test_full_size_field_transforms_in_milliseconds():
    """0.5 deg x 138 levels (BASELINE configs[1]/[4] geometry): the transformation is upload-bound;
    sanity of the fields (hydrostatic consistency of rho and the level heights, limits of calcpar)."""
    nuvz = 138
    cb0 = fb.make_config(nx=721, ny=361, nz=nuvz, dx=0.5, dy=0.5, xlon0=-180.0, ylat0=-90.0, numxgrid=720, numygrid=360,
                         dxout=0.5, dyout=0.5, outlon0=-180.0, outlat0=-90.0, npart=(8,), maxpart=64,
                         height=fb.synth_heights(nuvz))
    akm, bkm, akz, bkz, _ = conv_cases.hybrid_levels(nuvz)
    raw = met_cases.raw_fields(cb0, akz, bkz, nuvz, seed=5)
    hh, _ = fb.verttransform_heights(cb0, nuvz, akz[1:], bkz[1:], raw)
    cb = fb.make_config(nx=721, ny=361, nz=nuvz, dx=0.5, dy=0.5, xlon0=-180.0, ylat0=-90.0, numxgrid=720, numygrid=360,
                        dxout=0.5, dyout=0.5, outlon0=-180.0, outlat0=-90.0, npart=(8,), maxpart=64, height=hh)
    c = cb.cfg
    eng = fb.Engine(cb)
    eng.set_vertical(nuvz, akm[1:], bkm[1:], akz[1:], bkz[1:])
    eng.calcpar_verttransform(1, raw)
    ms = min(eng.calcpar_verttransform(1, raw) for _ in range(2))
    print(f"fpb_calcpar_verttransform 721x361x138: {ms:.1f} ms on the device (upload + kernels), "
          f"kernels {eng.metproc_kernel_ms:.2f} ms")
    assert ms < 1000.0
    f = eng.fetch_met(1, fields=("rho", "hmix", "ustar", "oli", "tropopause", "ww", "uu"))
    assert np.isfinite(f["rho"][:c.nx, :c.ny]).all() and (f["rho"][:c.nx, :c.ny] > 0).all()
    assert f["hmix"][:c.nx, :c.ny].min() >= 100.0 and f["hmix"][:c.nx, :c.ny].max() <= 4500.0
    assert f["ustar"][:c.nx, :c.ny].min() >= 1e-8 and np.isfinite(f["oli"][:c.nx, :c.ny]).all()
    assert np.isfinite(f["ww"][:c.nx, :c.ny]).all() and np.abs(f["ww"][:c.nx, :c.ny]).max() < 50.0
    eng.close()
