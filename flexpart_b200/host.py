"""Host-side helpers: run constants, met arrays, particle arrays, releases and
the timemanager driver (bindings of libfpb_host.so, include/fpb_host.h)."""
import ctypes as C

import numpy as np

from . import abi
from .abi import (FpbConfig, FpbMetPtrs, FpbParticlePtrs, FpbStepStats, FpbhReleases, FpbhRun,
                  FpbhRunResult, FpbhEngine, FpbError, load_host_lib)

_pf = C.POINTER(C.c_float)


def _fp(a):
    return a.ctypes.data_as(_pf) if a is not None else None


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32)) if a is not None else None


def _hcheck(rc):
    if rc != 0:
        raise FpbError(load_host_lib().fpbh_last_error().decode())


def synth_heights(nz):
    h = np.zeros(nz, np.float32)
    _hcheck(load_host_lib().fpbh_synth_heights(nz, _fp(h)))
    return h


class ConfigBundle:
    """fpb_config plus the numpy arrays its pointers refer to (kept alive)."""

    def __init__(self, cfg, height, npart, xmass):
        self.cfg, self.height, self.npart, self.xmass = cfg, height, npart, xmass
        cfg.height = _fp(height)
        cfg.npart = _ip(npart)
        cfg.xmass = _fp(xmass)

    def clone(self, **over):
        c = FpbConfig()
        C.memmove(C.byref(c), C.byref(self.cfg), C.sizeof(FpbConfig))
        for k, v in over.items():
            setattr(c, k, v)
        return ConfigBundle(c, self.height, self.npart, self.xmass)


def make_config(nx=361, ny=181, nz=138, dx=1.0, dy=1.0, xlon0=-180.0, ylat0=-90.0,
                nxmax=None, nymax=None, nzmax=None,
                ldirect=1, lsynctime=900, ctl=-5.0, ifine=4, cblflag=0, turboff=0,
                mdomainfill=0, lsettling=0, ind_samp=0,
                nspec=1, decay=None, drydepspec=None, density=None, dquer=None, vsetaver=None,
                cunningham=None, lage=(1728000,),
                outlon0=-25.0, outlat0=10.0, numxgrid=85, numygrid=65, dxout=1.0, dyout=1.0,
                outheights=(100.0, 500.0, 1000.0, 50000.0), nest=None,
                ioutputforeachrelease=1, lusekerneloutput=1, lparticlecountoutput=0,
                npart=(10000,), xmass=None, maxspec=5, nclassunc=1, receptors=(),
                maxpart=None, device=0, rng_mode=abi.RNG_REFERENCE, math_mode=abi.MATH_FAST,
                scatter_mode=abi.SCATTER_ATOMIC, seed=0x5EEDF1E0, height=None,
                part_id_stride=1, part_id_offset=0, sort_interval=0, met_nests=(),
                wetdepspec=None, weta_gas=None, wetb_gas=None, crain_aero=None, csnow_aero=None,
                ccn_aero=None, in_aero=None, henry=None, readclouds=0, ind_receptor=1, iflux=0, ipout=0, linit_cond=0):
    """Run constants for the engine, derived the way the reference's
    gridcheck_ecmwf / readcommand / readoutgrid / readreleases derive them."""
    L = load_host_lib()
    c = FpbConfig()
    c.abi_version = abi.ABI_VERSION
    c.nx, c.ny, c.nz = nx, ny, nz
    c.nxmax, c.nymax, c.nzmax = nxmax or nx, nymax or ny, nzmax or nz
    c.dx, c.dy, c.xlon0, c.ylat0 = dx, dy, xlon0, ylat0
    _hcheck(L.fpbh_gridcheck(C.byref(c)))
    for (xlon0n, ylat0n, nxn, nyn, dxn, dyn) in met_nests:  # nested input grids, gridcheck_nests
        _hcheck(L.fpbh_gridcheck_nest(C.byref(c), xlon0n, ylat0n, nxn, nyn, dxn, dyn))
    c.ldirect, c.lsynctime, c.ctl, c.ifine, c.cblflag = ldirect, lsynctime, ctl, ifine, cblflag
    _hcheck(L.fpbh_readcommand(C.byref(c)))
    c.turboff, c.mdomainfill, c.mquasilag, c.lsettling = turboff, mdomainfill, 0, lsettling
    c.ind_samp = ind_samp
    c.ioutputforeachrelease = ioutputforeachrelease
    c.lusekerneloutput, c.lparticlecountoutput = lusekerneloutput, lparticlecountoutput
    c.nspec, c.maxspec = nspec, max(maxspec, nspec)

    def setarr(name, vals, default):
        vals = list(vals) if vals is not None else [default] * nspec
        for k in range(nspec):
            getattr(c, name)[k] = vals[k]
    setarr("decay", decay, 0.0)
    setarr("drydepspec", drydepspec, 0)
    setarr("density", density, 0.0)
    setarr("dquer", dquer, 0.0)
    setarr("vsetaver", vsetaver, 0.0)
    setarr("cunningham", cunningham, 1.0)
    c.drydep = 1 if any(c.drydepspec[k] for k in range(nspec)) else 0
    # wet deposition switches, readspecies / readreleases.f90:349-370 (negative = off)
    setarr("wetdepspec", wetdepspec, 0)
    setarr("weta_gas", weta_gas, -1.0)
    setarr("wetb_gas", wetb_gas, -1.0)
    setarr("crain_aero", crain_aero, -1.0)
    setarr("csnow_aero", csnow_aero, -1.0)
    setarr("ccn_aero", ccn_aero, -1.0)
    setarr("in_aero", in_aero, -1.0)
    setarr("henry", henry, 0.0)
    c.wetdep = 1 if any(c.wetdepspec[k] for k in range(nspec)) else 0
    # backward runs, src/readcommand.f90:320-339: IND_RECEPTOR 3 = wet, 4 = dry deposition at the receptor
    c.wetbkdep = 1 if (ldirect == -1 and ind_receptor == 3) else 0
    c.drybkdep = 1 if (ldirect == -1 and ind_receptor == 4) else 0
    c.readclouds = readclouds
    for l in range(c.numbnests):
        c.readclouds_nest[l] = readclouds
    c.nageclass = len(lage)
    for k, v in enumerate(lage):
        c.lage[k] = v
    c.maxageclass = max(1, len(lage))
    oh = np.asarray(outheights, np.float32)
    _hcheck(L.fpbh_readoutgrid(C.byref(c), outlon0, outlat0, numxgrid, numygrid, dxout, dyout,
                               _fp(oh), len(oh)))
    if nest is not None:
        _hcheck(L.fpbh_readoutgrid_nest(C.byref(c), *nest))
    npart_a = np.asarray(npart, np.int32)
    numpoint = len(npart_a)
    c.numpoint = numpoint
    c.maxpointspec_act = numpoint if ioutputforeachrelease == 1 else 1
    c.nclassunc = nclassunc
    if xmass is None:
        xmass_a = np.ones((numpoint, c.maxspec), np.float32, order="F")
    else:
        xmass_a = np.asfortranarray(np.asarray(xmass, np.float32).reshape(numpoint, -1))
        if xmass_a.shape[1] < c.maxspec:
            pad = np.zeros((numpoint, c.maxspec), np.float32, order="F")
            pad[:, :xmass_a.shape[1]] = xmass_a
            xmass_a = pad
    c.numreceptor = len(receptors)
    for k, (xr, yr, area) in enumerate(receptors):
        c.xreceptor[k], c.yreceptor[k], c.receptorarea[k] = xr, yr, area
    c.maxpart = maxpart or int(npart_a.sum())
    c.device, c.rng_mode, c.math_mode, c.scatter_mode, c.seed = device, rng_mode, math_mode, scatter_mode, seed
    c.part_id_stride, c.part_id_offset = part_id_stride, part_id_offset
    c.sort_interval = sort_interval
    c.iflux, c.ipout, c.linit_cond = iflux, ipout, linit_cond
    h = np.ascontiguousarray(height, np.float32) if height is not None else synth_heights(nz)
    return ConfigBundle(c, h, npart_a, xmass_a)


class MetFields:
    """One time level of the com_mod met arrays (padded, Fortran order)."""
    NAMES3 = ("uu", "vv", "ww", "rho", "drhodz", "tt", "uupol", "vvpol")
    NAMES2 = ("hmix", "ustar", "wstar", "oli", "tropopause", "lsprec", "convprec", "tcc", "ctwc")

    def __init__(self, cb, nest=0):
        """nest = 0: mother grid; nest = l >= 1: nested input grid l (uun.. of com_mod)."""
        c = cb.cfg
        self.cb, self.nest = cb, nest
        nxm, nym = (c.nxmax, c.nymax) if nest == 0 else (c.nxmaxn, c.nymaxn)
        for n in self.NAMES3:
            setattr(self, n, np.zeros((nxm, nym, c.nzmax), np.float32, order="F"))
        for n in self.NAMES2:
            setattr(self, n, np.zeros((nxm, nym), np.float32, order="F"))
        self.vdep = np.zeros((nxm, nym, c.maxspec), np.float32, order="F")
        self.clouds = np.zeros((nxm, nym, c.nzmax), np.int8, order="F")
        self.ptrs = FpbMetPtrs()
        for n in self.NAMES3 + self.NAMES2 + ("vdep",):
            setattr(self.ptrs, n, _fp(getattr(self, n)))
        self.ptrs.clouds = self.clouds.ctypes.data_as(C.POINTER(C.c_int8))

    def synth(self, time_s):
        L = load_host_lib()
        if self.nest == 0:
            _hcheck(L.fpbh_synth_met(C.byref(self.cb.cfg), _fp(self.cb.height), int(time_s), C.byref(self.ptrs)))
        else:
            _hcheck(L.fpbh_synth_met_nest(C.byref(self.cb.cfg), _fp(self.cb.height), int(time_s), self.nest,
                                          C.byref(self.ptrs)))
        return self

    def homogeneous(self, u=10.0, v=0.0, w=0.0):
        _hcheck(load_host_lib().fpbh_homogeneous_met(C.byref(self.cb.cfg), u, v, w, C.byref(self.ptrs)))
        return self


class Particles:
    """Host mirror of the particle arrays (src/com_mod.f90:675-695)."""
    F32 = ("ztra1", "uap", "ucp", "uzp", "us", "vs", "ws")
    I32 = ("itra1", "npoint", "nclass", "idt", "itramem", "itrasplit")

    def __init__(self, maxpart, nspec, pinned=False):
        self.maxpart, self.nspec = maxpart, nspec
        self._pin = []
        if pinned:  # page-locked host arrays (torch is the allocator, nothing more)
            import torch

            def zeros(shape, dt):
                n = int(np.prod(shape))
                t = torch.zeros(n, dtype=getattr(torch, np.dtype(dt).name), pin_memory=True)
                self._pin.append(t)
                return t.numpy().reshape(shape, order="F")
        else:
            def zeros(shape, dt):
                return np.zeros(shape, dt, order="F")
        self.xtra1 = zeros(maxpart, np.float64)
        self.ytra1 = zeros(maxpart, np.float64)
        for n in self.F32:
            setattr(self, n, zeros(maxpart, np.float32))
        for n in self.I32:
            setattr(self, n, zeros(maxpart, np.int32))
        self.itra1[:] = abi.ITRA_DEAD
        self.cbt = zeros(maxpart, np.int16)
        self.cbt[:] = 1
        self.xmass1 = zeros((maxpart, nspec), np.float32)
        self.xscav_frac1 = zeros((maxpart, nspec), np.float32)
        self.numpart = 0
        p = FpbParticlePtrs()
        p.xtra1 = self.xtra1.ctypes.data_as(C.POINTER(C.c_double))
        p.ytra1 = self.ytra1.ctypes.data_as(C.POINTER(C.c_double))
        for n in self.F32:
            setattr(p, n, _fp(getattr(self, n)))
        for n in self.I32:
            setattr(p, n, _ip(getattr(self, n)))
        p.cbt = self.cbt.ctypes.data_as(C.POINTER(C.c_int16))
        p.xmass1 = _fp(self.xmass1)
        p.xscav_frac1 = _fp(self.xscav_frac1)
        p.ld = maxpart
        self.ptrs = p

    def copy(self):
        q = Particles(self.maxpart, self.nspec)
        for n in ("xtra1", "ytra1", "cbt", "xmass1", "xscav_frac1") + self.F32 + self.I32:
            getattr(q, n)[...] = getattr(self, n)
        q.numpart = self.numpart
        return q


class Releases:
    """RELEASES after coordtrafo (grid units) and time snapping."""

    def __init__(self, cb, lon1, lon2, lat1, lat2, z1, z2, start, end, itsplit=99999999):
        c = cb.cfg
        a = lambda v, t: np.atleast_1d(np.asarray(v, t)).copy()
        # coordtrafo.f90:37-42
        self.xpoint1 = ((a(lon1, np.float32) - np.float32(c.xlon0)) / np.float32(c.dx)).astype(np.float32)
        self.xpoint2 = ((a(lon2, np.float32) - np.float32(c.xlon0)) / np.float32(c.dx)).astype(np.float32)
        self.ypoint1 = ((a(lat1, np.float32) - np.float32(c.ylat0)) / np.float32(c.dy)).astype(np.float32)
        self.ypoint2 = ((a(lat2, np.float32) - np.float32(c.ylat0)) / np.float32(c.dy)).astype(np.float32)
        self.zpoint1, self.zpoint2 = a(z1, np.float32), a(z2, np.float32)
        self.start, self.end = a(start, np.int32), a(end, np.int32)
        n = len(self.xpoint1)
        assert n == c.numpoint, "release count must match config numpoint"
        for arr in (self.xpoint2, self.ypoint1, self.ypoint2, self.zpoint1, self.zpoint2, self.start, self.end):
            assert len(arr) == n
        r = FpbhReleases()
        r.numpoint = n
        r.ireleasestart, r.ireleaseend = _ip(self.start), _ip(self.end)
        r.xpoint1, r.ypoint1, r.xpoint2, r.ypoint2 = map(_fp, (self.xpoint1, self.ypoint1, self.xpoint2, self.ypoint2))
        r.zpoint1, r.zpoint2 = _fp(self.zpoint1), _fp(self.zpoint2)
        r.itsplit = itsplit
        self.c_struct = r


def synth_hybrid_levels(nuvz):
    """(akm, bkm, akz, bkz, nconvlev): the synthetic vertical structure of fpbh_timemanager's met_raw runs
    (arrays (1:nuvz), 0-based)"""
    arrs = [np.zeros(nuvz, np.float32) for _ in range(4)]
    n = C.c_int32(0)
    _hcheck(load_host_lib().fpbh_synth_hybrid_levels(nuvz, *[_fp(a) for a in arrs], C.byref(n)))
    return (*arrs, n.value)


def synth_rawmet(cb, nuvz, akz, bkz, time_s):
    """one time level of fpbh_synth_rawmet as a dict of Fortran-ordered arrays (the `raw` argument of
    Engine.calcpar_verttransform)"""
    from .abi import FpbRawmetPtrs
    c = cb.cfg
    raw, m = {}, FpbRawmetPtrs()
    for n in ("uuh", "vvh", "tth", "qvh", "wwh"):
        raw[n] = np.zeros((c.nxmax, c.nymax, c.nzmax), np.float32, order="F")
    for n in ("ps", "tt2", "td2", "sshf", "surfstr", "lsprec", "convprec", "tcc"):
        raw[n] = np.zeros((c.nxmax, c.nymax), np.float32, order="F")
    for n, a in raw.items():
        setattr(m, n, _fp(a))
    keep = [np.ascontiguousarray(akz[:nuvz], np.float32), np.ascontiguousarray(bkz[:nuvz], np.float32)]
    _hcheck(load_host_lib().fpbh_synth_rawmet(C.byref(c), nuvz, _fp(keep[0]), _fp(keep[1]), int(time_s), C.byref(m)))
    return raw


def verttransform_heights(cb, nuvz, akz, bkz, raw):
    """height(1:nuvz) of verttransform_ecmwf's first call (src/verttransform_ecmwf.f90:131-163) from a raw
    wind field (dict with ps, tt2, td2, tth, qvh, Fortran order); akz, bkz 0-based (1:nuvz).
    Returns (height, (ixm, jym))."""
    h = np.zeros(nuvz, np.float32)
    keep = [np.ascontiguousarray(akz[:nuvz], np.float32), np.ascontiguousarray(bkz[:nuvz], np.float32)] + \
           [np.asfortranarray(raw[n], np.float32) for n in ("ps", "tt2", "td2", "tth", "qvh")]
    ix, jy = C.c_int32(0), C.c_int32(0)
    _hcheck(load_host_lib().fpbh_verttransform_heights(C.byref(cb.cfg), nuvz, *[_fp(a) for a in keep], _fp(h),
                                                       C.byref(ix), C.byref(jy)))
    return h, (ix.value, jy.value)


def outgrid_geometry(cb, outlat0, nest=0):
    """area(numxgrid,numygrid), volume(numxgrid,numygrid,numzgrid) of outgrid_init (Fortran order)."""
    c = cb.cfg
    nx, ny = (c.numxgridn, c.numygridn) if nest else (c.numxgrid, c.numygrid)
    area = np.zeros((nx, ny), np.float32, order="F")
    volume = np.zeros((nx, ny, c.numzgrid), np.float32, order="F")
    _hcheck(load_host_lib().fpbh_outgrid_geometry(C.byref(c), nest, outlat0, _fp(area), _fp(volume)))
    return area, volume


class ReleaseState:
    def __init__(self, numpoint, mp_pid=0):
        """mp_pid > 0: the rank's ran1 seed offset of the MPI build (src/mpi_mod.f90:331-335)."""
        self._L = load_host_lib()
        self.h = self._L.fpbh_release_state_new(numpoint)
        self._L.fpbh_release_state_set_rank(self.h, mp_pid)

    def __del__(self):
        if getattr(self, "h", None):
            self._L.fpbh_release_state_free(self.h)
            self.h = None


def release_particles(cb, rel, state, itime, parts):
    """releaseparticles(itime) on the host mirror; returns (first, count) of changed rows."""
    numpart = C.c_int32(parts.numpart)
    first, n = C.c_int32(0), C.c_int32(0)
    _hcheck(load_host_lib().fpbh_releaseparticles(C.byref(cb.cfg), _fp(cb.height), C.byref(rel.c_struct),
                                                  state.h, itime, C.byref(parts.ptrs), C.byref(numpart),
                                                  C.byref(first), C.byref(n)))
    parts.numpart = numpart.value
    return first.value, n.value


class RunSpec:
    def __init__(self, ideltas, loutstep=3600, loutaver=3600, loutsample=900, met_interval=10800,
                 homogeneous=None, max_steps=0, ldirect=1, met_raw=False, lconvection=False):
        r = FpbhRun()
        s = -1 if ldirect < 0 else 1
        r.ideltas = ideltas
        r.loutstep, r.loutaver, r.loutsample = s * loutstep, s * loutaver, s * loutsample
        r.met_interval = met_interval
        if homogeneous is not None:
            r.met_homogeneous = 1
            r.met_u, r.met_v, r.met_w = homogeneous
        r.max_steps = max_steps
        r.met_raw, r.lconvection = int(met_raw), int(lconvection)
        self.c_struct = r


def timemanager(cb, rel, run, engine_vtable, on_output=None):
    """Run the reference's time loop on an engine (see fpbh_timemanager).
    `engine_vtable` is an abi.FpbhEngine; returns (FpbhRunResult, outputs)."""
    c = cb.cfg
    outs = []
    outer = c.maxspec * c.maxpointspec_act * c.nclassunc * c.maxageclass
    ng = c.numxgrid * c.numygrid * c.numzgrid * outer
    nd = c.numxgrid * c.numygrid * outer
    ngn = c.numxgridn * c.numygridn * c.numzgrid * outer
    ndn = c.numxgridn * c.numygridn * outer

    def cb_out(user, itime, outnum, g, gn, d, dn, cr):
        rec = {"itime": itime, "outnum": outnum,
               "gridunc": np.ctypeslib.as_array(g, shape=(ng,)).copy(),
               "drygridunc": np.ctypeslib.as_array(d, shape=(nd,)).copy(),
               "creceptor": np.ctypeslib.as_array(cr, shape=(abi.MAXRECEPTOR * c.maxspec,)).copy()}
        if c.nested_output == 1 and gn:
            rec["griduncn"] = np.ctypeslib.as_array(gn, shape=(ngn,)).copy()
            rec["drygriduncn"] = np.ctypeslib.as_array(dn, shape=(ndn,)).copy()
        outs.append(rec)
        if on_output:
            on_output(rec)
        return 0

    fn = abi.OUTPUT_FN(cb_out)
    res = FpbhRunResult()
    _hcheck(load_host_lib().fpbh_timemanager(C.byref(c), _fp(cb.height), C.byref(rel.c_struct),
                                             C.byref(run.c_struct), C.byref(engine_vtable), fn, None,
                                             C.byref(res)))
    return res, outs
