"""Convective mixing (convmix / calcmatrix / convect43c / redist).

CPU part (this file, no GPU): the column code of the device (flexpart_b200/csrc/fpb_convect.cuh,
compiled for the host by tests/conv_host_check.cpp) against the reference's own routines
(oracle/_ref/libflexref.so: calcmatrix, convect, tlift, f_qvsat, ew, redist transpiled from the
Fortran): lconv, cbmf, nconvtop, the redistribution matrix, the subsidence, the half-level heights
and the new particle heights, bit for bit, over many random soundings, forward and backward."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import flexpart_b200 as fb
import cases
import conv_cases
import ref_api
from oracle_api import Oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_pf, _pi = C.POINTER(C.c_float), C.POINTER(C.c_int)
fp = lambda a: a.ctypes.data_as(_pf)


@pytest.fixture(scope="module")
def hostlib(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("conv") / "libconvcheck.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-ffp-contract=off", "-shared", "-o", so,
                           os.path.join(ROOT, "tests", "conv_host_check.cpp")])
    L = C.CDLL(so)
    L.conv_check_column.argtypes = [C.c_int, C.c_int, _pf, _pf, _pf, _pf, _pf, _pf, C.c_float, C.c_float, C.c_float,
                                    C.c_float, _pf, C.c_int, C.c_int, C.c_int, _pf, _pf, _pi, _pi, _pf, _pf, _pf, _pi]
    L.conv_check_column_rows.argtypes = [C.c_int, C.c_int, _pf, _pf, _pf, _pf, _pf, _pf, C.c_float, C.c_float, _pf, C.c_int,
                                         _pi, _pf, _pf]
    return L


@pytest.mark.parametrize("rows", [1, 8, 5])
def test_rows_of_the_level_pair_loops_are_independent(hostlib, rows):
    """What conv_mix_kernel / conv_assembly_kernel rely on: the rows of MENT (mixing fractions + normalisation) and of
    the redistribution matrix can be worked in any order, on a pool that is not zeroed, with the bits of the
    sequential routine (which test_column_code_is_bit_identical_to_the_reference_routines pins to the reference)."""
    nuvz = 138
    akm, bkm, akz, bkz, nconvlev = conv_cases.hybrid_levels(nuvz)
    rs = np.random.RandomState(23)
    L = nconvlev + 3
    n_conv = 0
    for col in range(120):
        tconv, qconv, ps, tt2, td2 = conv_cases.sounding(rs, akz, bkz, nuvz)
        cbmf0 = np.float32(rs.choice([0.0, 0.004, 0.02]))
        out = []
        for which in (0, 1):
            cb_mf = np.array([cbmf0], np.float32)
            ntop, ld, used = C.c_int(0), C.c_int(0), C.c_int(0)
            fm = np.zeros(L * L, np.float32); sub = np.zeros(nuvz + 2, np.float32); uvz = np.zeros(nuvz + 2, np.float32)
            z = np.zeros(1, np.float32); rn = np.zeros(1, np.float32)
            if which == 0:
                lc = hostlib.conv_check_column(nuvz, nconvlev, fp(akz), fp(bkz), fp(akm), fp(bkm), fp(tconv), fp(qconv), ps,
                                               tt2, td2, 900.0, fp(cb_mf), 1, 900, 0, fp(z), fp(rn), C.byref(used),
                                               C.byref(ntop), fp(fm), fp(sub), fp(uvz), C.byref(ld))
            else:
                lc = hostlib.conv_check_column_rows(nuvz, nconvlev, fp(akz), fp(bkz), fp(akm), fp(bkm), fp(tconv), fp(qconv),
                                                    ps, 900.0, fp(cb_mf), rows, C.byref(ntop), fp(fm), fp(sub))
            nt = ntop.value
            out.append((lc, cb_mf[0].tobytes(), nt, fm.reshape(L, L)[1:nt + 1, 1:nt + 1].tobytes() if lc else b"",
                        sub[1:nt + 1].tobytes() if lc else b""))
        assert out[0] == out[1], col
        n_conv += out[0][0]
    assert n_conv >= 20, n_conv


@pytest.mark.parametrize("ldirect", [1, -1])
def test_column_code_is_bit_identical_to_the_reference_routines(hostlib, ldirect):
    if not ref_api.available():
        pytest.skip("oracle/_ref/libflexref.so not built")
    nuvz = 138
    akm, bkm, akz, bkz, nconvlev = conv_cases.hybrid_levels(nuvz)
    cb = cases.config_small(nrel=1, npart_each=64, nz=nuvz, height=fb.synth_heights(nuvz), ldirect=ldirect)
    c = cb.cfg
    ref = ref_api.Ref(cb, maxrand=2000)
    ref.set("nuvz", nuvz); ref.set("nconvlev", nconvlev)
    for nm, a in (("akm", akm), ("bkm", bkm), ("akz", akz), ("bkz", bkz)):
        ref.arr(nm)[:nuvz] = a[1:nuvz + 1]
    ora = Oracle(cb)                      # for its ran3: the stream redist draws from (iseed = -88)
    seed = C.c_int32(-88)
    R = ref.L
    rs = np.random.RandomState(17)
    n_conv = n_moved = 0
    np_col = 40
    for col in range(300):
        tconv, qconv, ps, tt2, td2 = conv_cases.sounding(rs, akz, bkz, nuvz)
        cbmf0 = np.float32(rs.choice([0.0, 0.0, 0.004, 0.02]))
        z0 = rs.uniform(5.0, 16000.0, np_col).astype(np.float32)
        # --- the reference: calcmatrix, then redist for every particle of the column
        ref.arr("tconv")[:nuvz - 1] = tconv[1:nuvz]; ref.arr("qconv")[:nuvz - 1] = qconv[1:nuvz]
        ref.set("psconv", float(ps)); ref.set("tt2conv", float(tt2)); ref.set("td2conv", float(td2))
        lconv, cbmf = C.c_int(0), C.c_float(float(cbmf0))
        R.f_calcmatrix(C.byref(lconv), C.byref(C.c_float(900.0)), C.byref(cbmf), C.byref(C.c_int(2)))
        zr = z0.copy()
        if lconv.value:
            ref.arr("ztra1")[:np_col] = z0
            ktop = C.c_int(0)
            for i in range(np_col):
                R.f_redist(C.byref(C.c_int(i + 1)), C.byref(ktop), C.byref(C.c_int(0)))
            zr = ref.arr("ztra1")[:np_col].copy()
        # --- the device's column code on the host, same uniforms
        draws = int(ref.L.f_ran3.restype is None)   # (unused)
        rn = np.zeros(np_col, np.float32)
        # the reference consumed its ran3 stream; replay the same stream from the oracle
        state_before = None
        zc = z0.copy()
        cb_mf = np.array([cbmf0], np.float32)
        used, ntop, ld = C.c_int(0), C.c_int(0), C.c_int(0)
        L = nconvlev + 3
        fm = np.zeros(L * L, np.float32); sub = np.zeros(nuvz + 2, np.float32); uvz = np.zeros(nuvz + 2, np.float32)
        # first pass with zero uniforms only to learn how many are consumed, then draw them and redo
        lc = hostlib.conv_check_column(nuvz, nconvlev, fp(akz), fp(bkz), fp(akm), fp(bkm), fp(tconv), fp(qconv), ps, tt2,
                                       td2, 900.0, fp(cb_mf), ldirect, c.lsynctime, np_col, fp(zc), fp(rn),
                                       C.byref(used), C.byref(ntop), fp(fm), fp(sub), fp(uvz), C.byref(ld))
        assert lc == lconv.value, col
        assert cb_mf[0].tobytes() == np.float32(cbmf.value).tobytes(), (col, cb_mf[0], cbmf.value)
        if not lc:
            continue
        n_conv += 1
        rn = np.array([ora.L.fpo_ran3(ora.S, C.byref(seed)) for _ in range(used.value)] + [0.0] * (np_col - used.value),
                      np.float32)
        zc = z0.copy(); cb_mf[0] = cbmf0
        hostlib.conv_check_column(nuvz, nconvlev, fp(akz), fp(bkz), fp(akm), fp(bkm), fp(tconv), fp(qconv), ps, tt2, td2,
                                  900.0, fp(cb_mf), ldirect, c.lsynctime, np_col, fp(zc), fp(rn), C.byref(used),
                                  C.byref(ntop), fp(fm), fp(sub), fp(uvz), C.byref(ld))
        assert ntop.value == ref.get("nconvtop"), col
        nt = ntop.value
        fr = ref.arr("fmassfrac")[:nt, :nt]
        fg = fm.reshape(ld.value, ld.value).T[1:nt + 1, 1:nt + 1]         # (k,kk) at [k + ld*kk]
        assert np.array_equal(fr.view(np.uint32), np.ascontiguousarray(fg).view(np.uint32)), col
        assert np.array_equal(ref.arr("sub")[:nt].view(np.uint32), sub[1:nt + 1].view(np.uint32)), col
        assert np.array_equal(zr.view(np.uint32), zc.view(np.uint32)), (col, np.abs(zr - zc).max())
        n_moved += int((zc != z0).sum())
    assert n_conv >= 40 and n_moved > 200, (n_conv, n_moved)


# ------------------------------------------------------------------------------------------
# GPU: fpb_convmix against the reference's convmix
# ------------------------------------------------------------------------------------------
def _conv_setup(rng_mode, ldirect=1, n=6000, sort_interval=0, met_nests=(), iflux=0):
    nuvz = 138
    akm, bkm, akz, bkz, nconvlev = conv_cases.hybrid_levels(nuvz)
    cb = cases.config_small(nrel=4, npart_each=n // 4, nz=nuvz, height=fb.synth_heights(nuvz), ldirect=ldirect,
                            rng_mode=rng_mode, math_mode=fb.MATH_STRICT, sort_interval=sort_interval,
                            met_nests=met_nests, iflux=iflux)
    sign = 1 if ldirect == 1 else -1
    f0 = conv_cases.conv_fields(cb, akz, bkz, nuvz, 1)
    f1 = conv_cases.conv_fields(cb, akz, bkz, nuvz, 2, tshift=1.5)
    p = cases.seeded_particles(cb, n, zmax=15000.0, lat_range=(-60.0, 60.0))
    p.itra1[n - 200:n] = 900 * sign          # not due: must not be touched
    return cb, (akm, bkm, akz, bkz, nconvlev, nuvz), (f0, f1), p, sign


@pytest.mark.gpu
@pytest.mark.parametrize("ldirect", [1, -1])
@pytest.mark.parametrize("sort_interval,iflux", [(0, 0), (1, 1)])
def test_convmix_reference_stream_is_bit_identical(ldirect, sort_interval, iflux):
    """convmix on the device, the reference's ran3 stream replayed in its sort2 visiting order, against
    the reference's own convmix (oracle/_ref): every particle height bit-identical over three calls
    (cbaseflux carried from call to call), forward and backward, rows cell-sorted or not.  iflux = 1: the
    gross fluxes of the convective displacements (calcfluxes, src/convmix.f90:205-218) as well."""
    if not ref_api.available():
        pytest.skip("oracle/_ref/libflexref.so not built")
    cb, (akm, bkm, akz, bkz, nconvlev, nuvz), (f0, f1), p, sign = _conv_setup(fb.RNG_REFERENCE, ldirect, 6000,
                                                                                  sort_interval, iflux=iflux)
    c, n = cb.cfg, p.numpart
    mets = (fb.MetFields(cb).synth(0), fb.MetFields(cb).synth(10800 * sign))
    ref = ref_api.Ref(cb, maxrand=2000)
    ref.set("nuvz", nuvz); ref.set("nconvlev", nconvlev)
    for nm, a in (("akm", akm), ("bkm", bkm), ("akz", akz), ("bkz", bkz)):
        ref.arr(nm)[:nuvz] = a[1:nuvz + 1]
    for slot, f in ((1, f0), (2, f1)):
        for nm, a in zip(("ps", "tt2", "td2"), f[:3]):
            ref.arr(nm)[:, :, 0, slot - 1] = a
        ref.arr("tth")[:, :, :, slot - 1] = f[3]; ref.arr("qvh")[:, :, :, slot - 1] = f[4]
    ref.set_met_bracket((1, 2), (0, 10800 * sign))
    ref.arr("cbaseflux")[:] = 0.0
    ref.push_state(p)
    if iflux:
        ref.set("iflux", 1)
        oh = np.array([cb.cfg.outheight[k] for k in range(cb.cfg.numzgrid)], np.float32)
        half = ref.arr("outheighthalf")          # src/readoutgrid.f90:194-197
        half[0] = oh[0] / np.float32(2.0)
        half[1:] = (oh[:-1] + oh[1:]) / np.float32(2.0)
    eng = fb.Engine(cb)
    eng.upload_met(1, mets[0]); eng.upload_met(2, mets[1]); eng.set_met_bracket((1, 2), (0, 10800 * sign))
    eng.set_convection(nuvz, c.nzmax, nconvlev, akz[1:], bkz[1:], akm[1:], bkm[1:])
    eng.upload_convmet(1, *f0); eng.upload_convmet(2, *f1)
    eng.push_particles(p)
    if sort_interval:
        eng.sort_particles()
    moved = 0
    for k in range(3):
        itime = 0
        z_before = ref.arr("ztra1")[:n].copy()
        ref.L.f_convmix(C.byref(C.c_int(itime)), C.byref(C.c_int(2)))
        ncol, nconv = eng.convmix(itime)
        q = fb.Particles(c.maxpart, 1); q.numpart = n
        eng.pull_particles(q)
        zr = ref.arr("ztra1")[:n]
        assert nconv > 20 and ncol > nconv
        assert np.array_equal(zr.view(np.uint32), q.ztra1[:n].view(np.uint32)), (k, np.abs(zr - q.ztra1[:n]).max())
        assert np.array_equal(q.ztra1[n - 200:n], p.ztra1[n - 200:n])
        moved += int((zr != z_before).sum())
    assert moved > 500
    if iflux:   # unit masses: the sums are exact whatever the order of the atomics
        fg, fr = eng.fetch_fluxes(), ref.arr("flux")
        assert fg.shape == fr.shape and np.array_equal(fg, fr), (fg.sum(axis=(1, 2, 3, 4, 5, 6)), fr.sum(axis=(1, 2, 3, 4, 5, 6)))
        assert fg[4].sum() > 0 and fg[5].sum() > 0, (fg[4].sum(), fg[5].sum())
        assert not fg[:4].any()     # a vertical displacement crosses no lateral face
    eng.close()


@pytest.mark.gpu
def test_convmix_philox_moves_only_convecting_columns():
    """production RNG: same columns convect (the column work does not depend on the RNG), particles
    outside them keep their height, moved ones stay between the ground and the model top."""
    cb, (akm, bkm, akz, bkz, nconvlev, nuvz), (f0, f1), p, sign = _conv_setup(fb.RNG_PHILOX_INDEX, 1, 40000, 1)
    c, n = cb.cfg, p.numpart
    eng = fb.Engine(cb)
    eng.fill_rannumb()
    eng.upload_met(1, fb.MetFields(cb).synth(0)); eng.upload_met(2, fb.MetFields(cb).synth(10800))
    eng.set_met_bracket((1, 2), (0, 10800))
    eng.set_convection(nuvz, c.nzmax, nconvlev, akz[1:], bkz[1:], akm[1:], bkm[1:])
    eng.upload_convmet(1, *f0); eng.upload_convmet(2, *f1)
    eng.push_particles(p)
    ncol, nconv = eng.convmix(0)
    q = fb.Particles(c.maxpart, 1); q.numpart = n
    eng.pull_particles(q)
    moved = q.ztra1[:n] != p.ztra1[:n]
    assert 0 < nconv < ncol and 0.02 < moved.mean() < 0.9
    assert (q.ztra1[:n] >= 0).all() and (q.ztra1[:n] <= cb.height[c.nz - 1]).all()
    assert not moved[n - 200:].any()
    # a second engine gives the same result (counter-based stream), also through a step in between
    eng.conccalc(0, 1.0); eng.step(0)
    eng.close()


@pytest.mark.gpu
def test_convmix_nested_input_grids_bit_identical():
    """Two nested input grids (src/convmix.f90:100-134,198-281): a particle takes part in the columns of the
    innermost nest it is in, with the nest's own soundings and cbasefluxn; the reference visits the mother
    grid first, then nest by nest, each with its own sort2 -- the replayed ran3 stream follows that."""
    if not ref_api.available():
        pytest.skip("oracle/_ref/libflexref.so not built")
    nests = [(-40.0, -10.0, 81, 41, 1.0, 1.0), (-10.0, 0.0, 81, 61, 0.25, 0.25)]
    cb, (akm, bkm, akz, bkz, nconvlev, nuvz), (f0, f1), p, sign = _conv_setup(fb.RNG_REFERENCE, 1, 8000, 1, nests)
    c, n = cb.cfg, p.numpart
    assert c.numbnests == 2
    r = np.random.RandomState(3)
    p.xtra1[:5000] = (r.uniform(-50.0, 50.0, 5000) - c.xlon0) / c.dx      # most particles in and around the nests
    p.ytra1[:5000] = (r.uniform(-20.0, 40.0, 5000) - c.ylat0) / c.dy
    fn = {l: (conv_cases.conv_fields(cb, akz, bkz, nuvz, 1, nest=l), conv_cases.conv_fields(cb, akz, bkz, nuvz, 2, 1.5, nest=l))
          for l in (1, 2)}
    ref = ref_api.Ref(cb, maxrand=2000)
    ref.set("nuvz", nuvz); ref.set("nconvlev", nconvlev)
    for nm, a in (("akm", akm), ("bkm", bkm), ("akz", akz), ("bkz", bkz)):
        ref.arr(nm)[:nuvz] = a[1:nuvz + 1]
    for slot, f in ((1, f0), (2, f1)):
        for nm, a in zip(("ps", "tt2", "td2"), f[:3]):
            ref.arr(nm)[:, :, 0, slot - 1] = a
        ref.arr("tth")[:, :, :, slot - 1] = f[3]; ref.arr("qvh")[:, :, :, slot - 1] = f[4]
        for l in (1, 2):
            g = fn[l][slot - 1]
            for nm, a in zip(("psn", "tt2n", "td2n"), g[:3]):
                ref.arr(nm)[:, :, 0, slot - 1, l - 1] = a
            ref.arr("tthn")[:, :, :, slot - 1, l - 1] = g[3]; ref.arr("qvhn")[:, :, :, slot - 1, l - 1] = g[4]
    ref.set_met_bracket((1, 2), (0, 10800))
    ref.arr("cbaseflux")[:] = 0.0; ref.arr("cbasefluxn")[:] = 0.0
    ref.push_state(p)
    eng = fb.Engine(cb)
    m0, m1 = fb.MetFields(cb).synth(0), fb.MetFields(cb).synth(10800)
    eng.upload_met(1, m0); eng.upload_met(2, m1)
    for l in (1, 2):
        eng.upload_met_nest(1, l, fb.MetFields(cb, nest=l).synth(0)); eng.upload_met_nest(2, l, fb.MetFields(cb, nest=l).synth(10800))
    eng.set_met_bracket((1, 2), (0, 10800))
    eng.set_convection(nuvz, c.nzmax, nconvlev, akz[1:], bkz[1:], akm[1:], bkm[1:])
    eng.upload_convmet(1, *f0); eng.upload_convmet(2, *f1)
    for l in (1, 2):
        eng.upload_convmet(1, *fn[l][0], nest=l); eng.upload_convmet(2, *fn[l][1], nest=l)
    eng.push_particles(p)
    eng.sort_particles()
    for k in range(2):
        ref.L.f_convmix(C.byref(C.c_int(0)), C.byref(C.c_int(2)))
        ncol, nconv = eng.convmix(0)
        q = fb.Particles(c.maxpart, 1); q.numpart = n
        eng.pull_particles(q)
        zr = ref.arr("ztra1")[:n]
        assert nconv > 20
        assert np.array_equal(zr.view(np.uint32), q.ztra1[:n].view(np.uint32)), (k, np.abs(zr - q.ztra1[:n]).max())
    assert (ref.arr("cbasefluxn")[:, :, 0] > 0).sum() > 5 and (ref.arr("cbasefluxn")[:, :, 1] > 0).sum() > 5
    eng.close()
