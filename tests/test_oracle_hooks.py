"""The oracle's restatement of the loop's optional hooks (oracle/fpo_hooks.c: calcfluxes, partpos_average,
initial_cond_calc) pinned against the reference's own routines run from their sources (oracle/_ref).  CPU only."""
import ctypes as C

import numpy as np
import pytest

import flexpart_b200 as fb
import cases
import ref_api
from oracle_api import load

pytestmark = pytest.mark.skipif(not ref_api.available(), reason="oracle/_ref/libflexref.so not built")
_pf = C.POINTER(C.c_float)
P = lambda a: a.ctypes.data_as(_pf)


def _pair(a, b):
    return (_pf * 2)(P(a), P(b))


def _moves(cb, n, seed):
    """old and new positions: short moves, long ones, some across the date line / out of the output grid"""
    c = cb.cfg
    r = np.random.RandomState(seed)
    x0 = r.uniform(0.01, c.nx - 1.01, n); y0 = r.uniform(2.0, c.ny - 3.0, n)
    z0 = r.uniform(1.0, 12000.0, n).astype(np.float32)
    x1 = x0 + r.normal(0.0, 0.6, n); y1 = y0 + r.normal(0.0, 0.6, n)
    z1 = np.abs(z0 + r.normal(0.0, 900.0, n)).astype(np.float32)
    far = r.rand(n) < 0.1
    x1[far] = np.mod(x0[far] + r.uniform(-45.0, 45.0, far.sum()), c.nx - 1)
    x1 = np.clip(x1, 0.0, c.nx - 1.0)
    y1 = np.clip(y1, 0.0, c.ny - 1.0)
    return x0, y0, z0, x1, y1, z1


@pytest.mark.parametrize("foreach", [0, 1])
def test_calcfluxes_restatement_is_bit_identical(foreach):
    cb = cases.config_small(nrel=3, npart_each=4, nspec=2, lage=(7200, 86400), ioutputforeachrelease=foreach,
                            outlon0=-150.0, outlat0=-60.0, numxgrid=60, numygrid=24, dxout=5.0, dyout=5.0)
    c = cb.cfg
    L = load()
    ref = ref_api.Ref(cb, maxrand=1000)
    oh = np.array([c.outheight[k] for k in range(c.numzgrid)], np.float32)
    half = ref.arr("outheighthalf")
    half[0] = oh[0] / np.float32(2.0); half[1:] = (oh[:-1] + oh[1:]) / np.float32(2.0)
    n = 3000
    x0, y0, z0, x1, y1, z1 = _moves(cb, n, 3)
    r = np.random.RandomState(4)
    q = fb.Particles(c.maxpart if c.maxpart >= n else n, 2)
    flux = np.zeros((6, c.numxgrid, c.numygrid, c.numzgrid, c.nspec, c.maxpointspec_act, c.nageclass), np.float32, order="F")
    ref.set("maxpart", c.maxpart)
    for j in range(n):
        slot = j % c.maxpart
        nage, npoint = int(r.randint(1, 3)), int(r.randint(1, 4))
        mass = np.array([r.uniform(0.1, 2.0), r.uniform(0.1, 2.0)], np.float32)
        ref.arr("xtra1")[slot] = x1[j]; ref.arr("ytra1")[slot] = y1[j]; ref.arr("ztra1")[slot] = z1[j]
        ref.arr("npoint")[slot] = npoint
        ref.arr("xmass1")[slot, :2] = mass
        ref.L.f_calcfluxes(C.byref(C.c_int(nage)), C.byref(C.c_int(slot + 1)), C.byref(C.c_float(np.float32(x0[j]))),
                           C.byref(C.c_float(np.float32(y0[j]))), C.byref(C.c_float(z0[j])))
        L.fpo_calcfluxes(C.byref(c), P(flux), nage, npoint, C.c_float(np.float32(x0[j])), C.c_float(np.float32(y0[j])),
                         C.c_float(z0[j]), C.c_double(x1[j]), C.c_double(y1[j]), C.c_float(z1[j]), P(mass))
    fr = ref.arr("flux")
    assert fr.shape == flux.shape and all(fr[i].sum() > 0 for i in range(6))
    assert np.array_equal(flux.view(np.uint32), np.asfortranarray(fr).view(np.uint32))


def test_partpos_average_restatement_is_bit_identical():
    cb = cases.config_small(nrel=1, npart_each=8)
    c = cb.cfg
    L = load()
    m0, m1 = cases.met_pair(cb)
    r = np.random.RandomState(9)
    shp3, shp2 = m0.uu.shape, m0.hmix.shape
    pv = [np.asfortranarray(r.normal(0, 2e-6, shp3).astype(np.float32)) for _ in range(2)]
    qv = [np.asfortranarray(r.uniform(0, 0.02, shp3).astype(np.float32)) for _ in range(2)]
    oro = np.asfortranarray(r.uniform(0, 3000.0, shp2).astype(np.float32))
    ref = ref_api.Ref(cb, maxrand=1000)
    ref.upload_met(1, m0); ref.upload_met(2, m1); ref.set_met_bracket((1, 2), (0, 10800))
    ref.arr("oro")[...] = oro
    for s in range(2):
        ref.arr("pv")[:, :, :, s] = pv[s]; ref.arr("qv")[:, :, :, s] = qv[s]
    memtime = (C.c_int32 * 2)(0, 10800)
    args = [_pair(pv[0], pv[1]), _pair(qv[0], qv[1]), _pair(m0.tt, m1.tt), _pair(m0.uu, m1.uu), _pair(m0.vv, m1.vv),
            _pair(m0.rho, m1.rho), _pair(m0.hmix, m1.hmix), _pair(m0.tropopause, m1.tropopause)]
    names = ("cartx", "carty", "cartz", "z", "topo", "pv", "qv", "tt", "uu", "vv", "rho", "tro", "hmix", "energy")
    out = np.zeros(14, np.float32)
    itime = 2700
    for j in range(400):
        x, y, z = r.uniform(0.0, c.nx - 1.001), r.uniform(0.0, c.ny - 1.001), np.float32(r.uniform(0.0, 15000.0))
        ref.arr("xtra1")[0] = x; ref.arr("ytra1")[0] = y; ref.arr("ztra1")[0] = z
        for nm in names:
            ref.arr("part_av_" + nm)[0] = 0.0
        ref.arr("npart_av")[0] = 0
        ref.L.f_partpos_average(C.byref(C.c_int(itime)), C.byref(C.c_int(1)))
        L.fpo_partpos_average(C.byref(c), cb.height.ctypes.data_as(_pf), itime, memtime, C.c_double(x), C.c_double(y),
                              C.c_float(z), P(oro), *args, P(out))
        want = np.array([ref.arr("part_av_" + nm)[0] for nm in names], np.float32)
        assert ref.arr("npart_av")[0] == 1
        assert np.array_equal(out.view(np.uint32), want.view(np.uint32)), (j, dict(zip(names, zip(out, want))))


@pytest.mark.parametrize("linit_cond", [1, 2])
def test_initial_cond_calc_restatement_is_bit_identical(linit_cond):
    cb = cases.config_small(nrel=3, npart_each=4, nspec=2, ioutputforeachrelease=1, ldirect=-1,
                            outlon0=-150.0, outlat0=-60.0, numxgrid=60, numygrid=24, dxout=5.0, dyout=5.0)
    c = cb.cfg
    L = load()
    mets = (fb.MetFields(cb).synth(0), fb.MetFields(cb).synth(-10800))
    ref = ref_api.Ref(cb, maxrand=1000)
    ref.upload_met(1, mets[0]); ref.upload_met(2, mets[1]); ref.set_met_bracket((1, 2), (0, -10800))
    ref.set("linit_cond", linit_cond)
    ref.arr("init_cond")[...] = 0.0
    ic = np.zeros((c.numxgrid, c.numygrid, c.numzgrid, c.maxspec, c.maxpointspec_act), np.float32, order="F")
    r = np.random.RandomState(5)
    itime = -1800
    for j in range(3000):
        # (below the model top: above it the reference's level search leaves indz undefined)
        x, y, z = r.uniform(0.0, c.nx - 1.001), r.uniform(0.0, c.ny - 1.001), np.float32(r.uniform(0.0, cb.height[c.nz - 1] - 1.0))
        npoint = int(r.randint(1, 4))
        mass = np.array([r.uniform(0.1, 2.0), r.uniform(0.1, 2.0)], np.float32)
        ref.arr("xtra1")[0] = x; ref.arr("ytra1")[0] = y; ref.arr("ztra1")[0] = z
        ref.arr("itra1")[0] = itime; ref.arr("npoint")[0] = npoint
        ref.arr("xmass1")[0, :2] = mass
        ref.L.f_initial_cond_calc(C.byref(C.c_int(itime)), C.byref(C.c_int(1)))
        L.fpo_initial_cond_calc(C.byref(c), cb.height.ctypes.data_as(_pf), P(ic), linit_cond, C.c_double(x), C.c_double(y),
                                C.c_float(z), npoint, P(mets[1].rho), P(mass))
    want = ref.arr("init_cond")
    assert want.shape == ic.shape and want.sum() > 0 and (want > 0).sum() > 500
    assert np.array_equal(ic.view(np.uint32), np.asfortranarray(want).view(np.uint32))
