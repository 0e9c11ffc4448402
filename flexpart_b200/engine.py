"""Engine: Python handle on libfpb.so (include/fpb.h)."""
import ctypes as C

import numpy as np

from . import abi
from .abi import FpbStepStats, FpbError, FpbhEngine, FpbReleasePoints, FpbPartoutPtrs, load_engine_lib

_pf = C.POINTER(C.c_float)


def _fp(a):
    return a.ctypes.data_as(_pf) if a is not None else None


class Engine:
    """One engine instance = one GPU's share of the particles + a met replica."""

    def __init__(self, cb):
        self.L = load_engine_lib()
        self.cb = cb
        self.h = C.c_void_p()
        self._check(self.L.fpb_init(C.byref(cb.cfg), C.byref(self.h)))
        c = cb.cfg
        outer = c.maxspec * c.maxpointspec_act * c.nclassunc * c.maxageclass
        self.shape_grid = (c.numxgrid, c.numygrid, c.numzgrid, c.maxspec, c.maxpointspec_act,
                           c.nclassunc, c.maxageclass)
        self.shape_dry = (c.numxgrid, c.numygrid, c.maxspec, c.maxpointspec_act, c.nclassunc, c.maxageclass)
        self.shape_gridn = (c.numxgridn, c.numygridn, c.numzgrid, c.maxspec, c.maxpointspec_act,
                            c.nclassunc, c.maxageclass)
        self.shape_dryn = (c.numxgridn, c.numygridn, c.maxspec, c.maxpointspec_act, c.nclassunc, c.maxageclass)
        self._outer = outer

    def _check(self, rc):
        if rc != 0:
            raise FpbError(self.L.fpb_last_error().decode())

    def close(self):
        if self.h:
            self.L.fpb_finalize(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # --- RNG
    def set_rannumb(self, table):
        t = np.ascontiguousarray(table, np.float32)
        self._check(self.L.fpb_set_rannumb(self.h, _fp(t), len(t)))

    def fill_rannumb(self, maxrand=1000000, idummy=-320):
        self._check(self.L.fpb_fill_rannumb(self.h, maxrand, idummy))

    # --- met
    def upload_met(self, slot, met):
        self._check(self.L.fpb_upload_met(self.h, slot, C.byref(met.ptrs)))

    def upload_met_begin(self, slot, met):
        """read-ahead upload (slot 1..3): returns once the copies are enqueued; keep `met` alive and
        untouched until upload_met_end()"""
        self._met_keep = met
        self._check(self.L.fpb_upload_met_begin(self.h, slot, C.byref(met.ptrs)))

    def upload_met_end(self):
        ms = C.c_float(0.0)
        self._check(self.L.fpb_upload_met_end(self.h, C.byref(ms)))
        return ms.value

    def host_register(self, *arrays):
        """page-lock numpy arrays (cudaHostRegister) for asynchronous copies"""
        for a in arrays:
            self._check(self.L.fpb_host_register(a.ctypes.data, a.nbytes))

    def host_unregister(self, *arrays):
        for a in arrays:
            self._check(self.L.fpb_host_unregister(a.ctypes.data))

    def upload_met_nest(self, slot, nest, met):
        self._check(self.L.fpb_upload_met_nest(self.h, slot, nest, C.byref(met.ptrs)))

    def set_met_bracket(self, memind, memtime, lwindinterv=None):
        mi = (C.c_int32 * 2)(*memind)
        mt = (C.c_int32 * 2)(*memtime)
        if lwindinterv is None:
            lwindinterv = abs(memtime[1] - memtime[0])
        self._check(self.L.fpb_set_met_bracket(self.h, mi, mt, lwindinterv))

    # --- particles
    def push_particles(self, parts, first=0, count=None):
        count = parts.numpart - first if count is None else count
        self._check(self.L.fpb_push_particles(self.h, first, count, C.byref(parts.ptrs)))

    def pull_particles(self, parts, first=0, count=None):
        count = parts.numpart - first if count is None else count
        self._check(self.L.fpb_pull_particles(self.h, first, count, C.byref(parts.ptrs)))

    def set_numpart(self, n):
        self._check(self.L.fpb_set_numpart(self.h, n))

    # --- hot path
    def step(self, itime, ldeltat=0, stats=True):
        st = FpbStepStats()
        self._check(self.L.fpb_step(self.h, itime, ldeltat, C.byref(st) if stats else None))
        return st.as_dict() if stats else None

    def step_host(self, parts, itime, ldeltat=0, conc_weight=0.0, stats=True):
        """fpb_step_host: one interval on host-owned particle arrays (chunked, copies overlapped)."""
        st = FpbStepStats()
        self._check(self.L.fpb_step_host(self.h, itime, ldeltat, parts.numpart, C.byref(parts.ptrs),
                                         conc_weight, C.byref(st) if stats else None))
        return st.as_dict() if stats else None

    def set_outgrid_geometry(self, area, volume, arean=None, volumen=None):
        """area/volume of outgrid_init (host.outgrid_geometry), Fortran order."""
        keep = [np.asfortranarray(a, np.float32) if a is not None else None for a in (area, volume, arean, volumen)]
        self._check(self.L.fpb_set_outgrid_geometry(self.h, *[_fp(a) if a is not None else None for a in keep]))

    def set_orography(self, oro):
        self._check(self.L.fpb_set_orography(self.h, _fp(np.asfortranarray(oro, np.float32))))

    def upload_pvqv(self, slot, pv, qv):
        pv, qv = np.asfortranarray(pv, np.float32), np.asfortranarray(qv, np.float32)
        self._check(self.L.fpb_upload_pvqv(self.h, slot, _fp(pv), _fp(qv)))

    def partoutput(self, itime):
        """records of partoutput(itime): dict of arrays, one entry per particle with itra1 == itime."""
        c = self.cb.cfg
        mp = c.maxpart
        out = {k: np.zeros(mp, np.int32) for k in ("npoint", "itramem")}
        out.update({k: np.zeros(mp, np.float32) for k in ("xlon", "ylat", "ztra1", "topo", "pvi", "qvi", "rhoi", "hmixi", "tri", "tti")})
        out["xmass1"] = np.zeros((mp, c.nspec), np.float32, order="F")
        r = FpbPartoutPtrs()
        for k, a in out.items():
            setattr(r, k, a.ctypes.data_as(C.POINTER(C.c_int32 if a.dtype == np.int32 else C.c_float)))
        r.ld = mp
        n = C.c_int32(0)
        self._check(self.L.fpb_partoutput(self.h, itime, C.byref(n), C.byref(r)))
        return {k: a[:n.value].copy() for k, a in out.items()}

    def fetch_fluxes(self, zero=False):
        """flux(6, numxgrid, numygrid, numzgrid, nspec, maxpointspec_act, nageclass) of calcfluxes (iflux = 1)"""
        c = self.cb.cfg
        f = np.zeros((6, c.numxgrid, c.numygrid, c.numzgrid, c.nspec, c.maxpointspec_act, c.nageclass), np.float32, order="F")
        self._check(self.L.fpb_fetch_fluxes(self.h, _fp(f), 1 if zero else 0))
        return f

    def fetch_init_cond(self, zero=False):
        """init_cond(numxgrid, numygrid, numzgrid, maxspec, maxpointspec_act) of initial_cond_calc (linit_cond > 0)"""
        c = self.cb.cfg
        f = np.zeros((c.numxgrid, c.numygrid, c.numzgrid, c.maxspec, c.maxpointspec_act), np.float32, order="F")
        self._check(self.L.fpb_fetch_init_cond(self.h, _fp(f), 1 if zero else 0))
        return f

    def initial_cond_final(self, itime):
        self._check(self.L.fpb_initial_cond_final(self.h, itime))

    def fetch_partpos_average(self, numpart, zero=False):
        """npart_av and the part_av_* sums of partpos_average (ipout = 3) for the first numpart slots"""
        from .abi import FpbPartavPtrs
        r = FpbPartavPtrs()
        out = {"npart_av": np.zeros(numpart, np.int32)}
        r.npart_av = out["npart_av"].ctypes.data_as(C.POINTER(C.c_int32))
        for name, _ in FpbPartavPtrs._fields_[1:]:
            out[name] = np.zeros(numpart, np.float32)
            setattr(r, name, _fp(out[name]))
        self._check(self.L.fpb_fetch_partpos_average(self.h, numpart, C.byref(r), 1 if zero else 0))
        return out

    def set_outgrid_origin(self, outlon0, outlat0, outlon0n=0.0, outlat0n=0.0):
        self._check(self.L.fpb_set_outgrid_origin(self.h, outlon0, outlat0, outlon0n, outlat0n))

    def concoutput_sparse(self, which, ks, kp, nage, outnum, tot_mu=1.0, loutaver=3600, nest=0):
        """sparse dump of one (ks, kp, nage) grid: (sparse_dump_i, sparse_dump_r) of
        src/concoutput.f90:352-475; which = 0 concentration, 1 dry, 2 wet deposition."""
        c = self.cb.cfg
        n = (c.numxgridn * c.numygridn if nest else c.numxgrid * c.numygrid) * (c.numzgrid if which in (0, 3) else 1)
        di, dr = np.zeros(n, np.int32), np.zeros(n, np.float32)
        ci, cr = C.c_int32(0), C.c_int32(0)
        self._check(self.L.fpb_concoutput_sparse(self.h, nest, which, ks, kp, nage, outnum, tot_mu, loutaver,
                                                 C.byref(ci), di.ctypes.data_as(C.POINTER(C.c_int32)), C.byref(cr), _fp(dr)))
        return di[:ci.value].copy(), dr[:cr.value].copy()

    def set_releases(self, rel, mp_pid=0):
        """Release points for the device-side releaseparticles (host.Releases)."""
        r = FpbReleasePoints()
        for name, _ in FpbReleasePoints._fields_[:-1]:
            setattr(r, name, getattr(rel.c_struct, name))
        r.mp_pid = mp_pid
        self._rel_keep = rel
        self._check(self.L.fpb_set_releases(self.h, C.byref(r)))

    def release_particles(self, itime):
        """releaseparticles(itime) on the device; returns (numpart, particles created)."""
        n, m = C.c_int32(0), C.c_int32(0)
        self._check(self.L.fpb_releaseparticles(self.h, itime, C.byref(n), C.byref(m)))
        return n.value, m.value

    # ---- the grid exchange of a multi-GPU run (mpif_tm_reduce_grid slot), NCCL behind the C ABI
    @staticmethod
    def comm_unique_id():
        """128-byte NCCL id, created on rank 0 and broadcast by the host"""
        buf = C.create_string_buffer(128)
        L = load_engine_lib()
        if L.fpb_comm_unique_id(buf):
            raise FpbError(L.fpb_last_error().decode())
        return buf.raw

    def comm_init(self, unique_id, rank, nranks):
        buf = C.create_string_buffer(bytes(unique_id), 128)
        self._check(self.L.fpb_comm_init(self.h, buf, rank, nranks))

    def reduce_grids_begin(self):
        self._check(self.L.fpb_reduce_grids_begin(self.h))

    def reduce_grids_end(self, fetch=True):
        """waits for the exchange; with fetch=True rank 0 gets the summed grids (reference layout)"""
        c = self.cb.cfg
        out = {}
        if fetch:
            out["gridunc"] = np.zeros(self.shape_grid, np.float32, order="F")
            out["drygridunc"] = np.zeros(self.shape_dry, np.float32, order="F")
            if c.nested_output == 1:
                out["griduncn"] = np.zeros(self.shape_gridn, np.float32, order="F")
                out["drygriduncn"] = np.zeros(self.shape_dryn, np.float32, order="F")
            if c.wetdep:
                out["wetgridunc"] = np.zeros(self.shape_dry, np.float32, order="F")
                if c.nested_output == 1:
                    out["wetgriduncn"] = np.zeros(self.shape_dryn, np.float32, order="F")
            if c.numreceptor > 0:
                out["creceptor"] = np.zeros((abi.MAXRECEPTOR, c.maxspec), np.float32, order="F")
        g = lambda k: _fp(out[k]) if k in out else None
        self._check(self.L.fpb_reduce_grids_end(self.h, g("gridunc"), g("griduncn"), g("drygridunc"),
                                                g("drygriduncn"), g("wetgridunc"), g("wetgriduncn"), g("creceptor")))
        return out

    def reduce_grids_device(self, which=0):
        """(device pointer, floats, ms of the last reduce) of a summed staging buffer; waits for the exchange"""
        p, n, ms = C.c_void_p(), C.c_size_t(), C.c_float()
        self._check(self.L.fpb_reduce_grids_device(self.h, which, C.byref(p), C.byref(n), C.byref(ms)))
        return p.value, n.value, ms.value

    # ---- convective mixing (convmix / calcmatrix / convect / redist on the device)
    def set_convection(self, nuvz, nuvzmax, nconvlev, akz, bkz, akm, bkm):
        """akz.. are the Fortran arrays (1:nuvz) as numpy arrays of length >= nuvz (0-based)"""
        arrs = [np.ascontiguousarray(a[:nuvz], np.float32) for a in (akz, bkz, akm, bkm)]
        self._check(self.L.fpb_set_convection(self.h, nuvz, nuvzmax, nconvlev, *[_fp(a) for a in arrs]))

    def upload_convmet(self, slot, ps, tt2, td2, tth, qvh, nest=0):
        """ps, tt2, td2 (nxmax,nymax) and tth, qvh (nxmax,nymax,nuvzmax), Fortran order, float32;
        nest >= 1: the fields of that nested input grid (nxmaxn, nymaxn extents)"""
        from .abi import FpbConvPtrs
        m = FpbConvPtrs()
        keep = [np.asfortranarray(a, np.float32) for a in (ps, tt2, td2, tth, qvh)]
        m.ps, m.tt2, m.td2, m.tth, m.qvh = [_fp(a) for a in keep]
        if nest:
            self._check(self.L.fpb_upload_convmet_nest(self.h, slot, nest, C.byref(m)))
        else:
            self._check(self.L.fpb_upload_convmet(self.h, slot, C.byref(m)))

    def convmix(self, itime):
        """convmix(itime); returns (occupied columns, convecting columns)"""
        nc, nv = C.c_int32(0), C.c_int32(0)
        self._check(self.L.fpb_convmix(self.h, itime, C.byref(nc), C.byref(nv)))
        return nc.value, nv.value

    # ---- calcpar + verttransform_ecmwf on the device
    def set_vertical(self, nuvz, akm, bkm, akz, bkz, nwz=None, nuvzmax=None, nwzmax=None):
        """akm.. are the Fortran arrays (1:nuvz) as numpy arrays of length >= nuvz (0-based)"""
        c = self.cb.cfg
        arrs = [np.ascontiguousarray(a[:nuvz], np.float32) for a in (akm, bkm, akz, bkz)]
        self._check(self.L.fpb_set_vertical(self.h, nuvz, nwz or nuvz, nuvzmax or c.nzmax, nwzmax or c.nzmax,
                                            *[_fp(a) for a in arrs]))

    def calcpar_verttransform(self, slot, raw, lsubgrid=0):
        """raw: dict of Fortran-ordered float32 arrays (uuh, vvh, wwh, tth, qvh, [pvh,] ps, tt2, td2, sshf,
        surfstr, [lsprec, convprec, tcc, excessoro]) with the reference's padded extents; the met slot is
        built on the device.  Returns the device time (upload + kernels) in ms; the kernels' share is
        left in self.metproc_kernel_ms."""
        from .abi import FpbRawmetPtrs
        m = FpbRawmetPtrs()
        keep = {}
        for name, _ in FpbRawmetPtrs._fields_:
            a = raw.get(name)
            if a is not None:
                keep[name] = np.asfortranarray(a, np.float32)
                setattr(m, name, _fp(keep[name]))
        ms = (C.c_float * 2)()
        self._check(self.L.fpb_calcpar_verttransform(self.h, slot, C.byref(m), lsubgrid, ms))
        self.metproc_kernel_ms = ms[1]
        return ms[0]

    def calcpar_verttransform_nest(self, slot, nest, raw, dxn, dyn, xlon0n, ylat0n, lsubgrid=0):
        """the same for nested input grid `nest` (arrays with the nest's padded extents)"""
        from .abi import FpbRawmetPtrs
        m = FpbRawmetPtrs()
        keep = {}
        for name, _ in FpbRawmetPtrs._fields_:
            a = raw.get(name)
            if a is not None:
                keep[name] = np.asfortranarray(a, np.float32)
                setattr(m, name, _fp(keep[name]))
        ms = (C.c_float * 2)()
        self._check(self.L.fpb_calcpar_verttransform_nest(self.h, slot, nest, C.byref(m), lsubgrid, dxn, dyn, xlon0n, ylat0n, ms))
        return ms[0]

    def upload_vdep(self, slot, vdep):
        a = np.asfortranarray(vdep, np.float32)
        self._check(self.L.fpb_upload_vdep(self.h, slot, _fp(a)))

    def fetch_met(self, slot, fields=None, nest=0):
        """the transformed fields of a slot (of nested input grid `nest`) in the reference's padded layout"""
        from .abi import FpbMetOutPtrs
        c = self.cb.cfg
        if nest:
            class _N:   # the nest's padded extents
                nxmax, nymax, nzmax, wetdep = c.nxmaxn, c.nymaxn, c.nzmax, c.wetdep
            c = _N
            skip = ("uupol", "vvpol", "pv", "qv") + (() if c.wetdep else ("tt",))
            fields = [n for n, _ in FpbMetOutPtrs._fields_ if n not in skip and (fields is None or n in fields)]
        o, out = FpbMetOutPtrs(), {}
        for name, _ in FpbMetOutPtrs._fields_:
            if fields is not None and name not in fields:
                continue
            if name == "clouds":
                if not c.wetdep:
                    continue
                out[name] = np.zeros((c.nxmax, c.nymax, c.nzmax), np.int8, order="F")
                o.clouds = out[name].ctypes.data_as(C.POINTER(C.c_int8))
                continue
            if name == "ctwc" and not c.wetdep:
                continue
            shape = (c.nxmax, c.nymax) if name in ("hmix", "ustar", "wstar", "oli", "tropopause", "ctwc") else (c.nxmax, c.nymax, c.nzmax)
            out[name] = np.zeros(shape, np.float32, order="F")
            setattr(o, name, _fp(out[name]))
        if nest:
            self._check(self.L.fpb_fetch_met_nest(self.h, slot, nest, C.byref(o)))
        else:
            self._check(self.L.fpb_fetch_met(self.h, slot, C.byref(o)))
        return out

    def init_domainfill(self, box, itsplit=99999999):
        """init_domainfill (src/init_domainfill.f90:55-283) over the box (xpoint1, ypoint1, xpoint2,
        ypoint2) in grid units, on the device; returns (numpart, info dict)."""
        from .abi import FpbDomainfillInfo
        n, info = C.c_int32(0), FpbDomainfillInfo()
        self._check(self.L.fpb_init_domainfill(self.h, box[0], box[1], box[2], box[3], itsplit, C.byref(n),
                                               C.byref(info)))
        return n.value, dict(nx_we=tuple(info.nx_we), ny_sn=tuple(info.ny_sn), gdomainfill=info.gdomainfill,
                             numcolumn=info.numcolumn, numparttot=info.numparttot,
                             colmasstotal=info.colmasstotal, xmassperparticle=info.xmassperparticle)

    def boundcond_domainfill(self, itime, loutend=0):
        """boundcond_domainfill(itime, loutend) (src/boundcond_domainfill.f90:54-560); returns
        (numpart, particles created by this rank)."""
        n, m = C.c_int32(0), C.c_int32(0)
        self._check(self.L.fpb_boundcond_domainfill(self.h, itime, loutend, C.byref(n), C.byref(m)))
        return n.value, m.value

    def split_particles(self, itime):
        """particle splitting of timemanager (src/timemanager.f90:472-503); returns the new numpart."""
        n = C.c_int32(0)
        self._check(self.L.fpb_split_particles(self.h, itime, C.byref(n)))
        return n.value

    def wetdepo(self, itime, ltsample, ldeltat=0):
        """wetdepo(itime, ltsample, loutnext) with ldeltat precomputed (src/wetdepo.f90:55-63)."""
        self._check(self.L.fpb_wetdepo(self.h, itime, ltsample, ldeltat))

    def fetch_wetgrids(self):
        c = self.cb.cfg
        out = {"wetgridunc": np.zeros(self.shape_dry, np.float32, order="F")}
        wn = None
        if c.nested_output == 1:
            out["wetgriduncn"] = np.zeros(self.shape_dryn, np.float32, order="F")
            wn = out["wetgriduncn"]
        self._check(self.L.fpb_fetch_wetgrids(self.h, _fp(out["wetgridunc"]), _fp(wn)))
        return out

    def conccalc(self, itime, weight):
        self._check(self.L.fpb_conccalc(self.h, itime, weight))

    def fetch_grids(self, zero_conc=True):
        c = self.cb.cfg
        out = {"gridunc": np.zeros(self.shape_grid, np.float32, order="F"),
               "drygridunc": np.zeros(self.shape_dry, np.float32, order="F"),
               "creceptor": np.zeros((abi.MAXRECEPTOR, c.maxspec), np.float32, order="F")}
        gn = dn = None
        if c.nested_output == 1:
            out["griduncn"] = np.zeros(self.shape_gridn, np.float32, order="F")
            out["drygriduncn"] = np.zeros(self.shape_dryn, np.float32, order="F")
            gn, dn = out["griduncn"], out["drygriduncn"]
        self._check(self.L.fpb_fetch_grids(self.h, _fp(out["gridunc"]), _fp(gn), _fp(out["drygridunc"]),
                                           _fp(dn), _fp(out["creceptor"]), 1 if zero_conc else 0))
        return out

    def scale_depgrids(self, factors):
        f = np.ascontiguousarray(factors, np.float32)
        self._check(self.L.fpb_scale_depgrids(self.h, _fp(f)))

    def zero_conc_grids(self):
        self._check(self.L.fpb_zero_conc_grids(self.h))

    def sort_particles(self):
        self._check(self.L.fpb_sort_particles(self.h))

    def grid_device_ptr(self, which):
        p, n = C.c_void_p(), C.c_size_t()
        self._check(self.L.fpb_grid_device_ptr(self.h, which, C.byref(p), C.byref(n)))
        return p.value, n.value

    def kernel_times(self):
        """(step_ms, conccalc_ms) of the most recent calls, device time."""
        a, b = C.c_float(), C.c_float()
        self._check(self.L.fpb_kernel_times(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def get_rannumb(self, n):
        out = np.zeros(n, np.float32)
        self._check(self.L.fpb_get_rannumb(self.h, _fp(out), n))
        return out

    @property
    def stream(self):
        return self.L.fpb_stream(self.h)

    @property
    def launch_count(self):
        return int(self.L.fpb_launch_count(self.h))

    def vtable(self, device_release=False):
        """fpbh_engine table pointing at the library's own entry points.
        device_release: releaseparticles runs on the device (fpb_releaseparticles)."""
        L, a = self.L, abi
        v = FpbhEngine()
        v.self = self.h
        cast = lambda fn, T: C.cast(fn, T)
        v.upload_met = cast(L.fpb_upload_met, a.UPLOAD_MET_FN)
        v.set_met_bracket = cast(L.fpb_set_met_bracket, a.SET_BRACKET_FN)
        v.push_particles = cast(L.fpb_push_particles, a.PUSH_FN)
        v.pull_particles = cast(L.fpb_pull_particles, a.PUSH_FN)
        v.set_numpart = cast(L.fpb_set_numpart, a.SET_NUMPART_FN)
        v.step = cast(L.fpb_step, a.STEP_FN)
        v.conccalc = cast(L.fpb_conccalc, a.CONC_FN)
        v.fetch_grids = cast(L.fpb_fetch_grids, a.FETCH_FN)
        v.scale_depgrids = cast(L.fpb_scale_depgrids, a.SCALE_FN)
        v.wetdepo = cast(L.fpb_wetdepo, a.WETDEPO_FN)
        if device_release:
            v.set_releases = cast(L.fpb_set_releases, a.SET_RELEASES_FN)
            v.releaseparticles = cast(L.fpb_releaseparticles, a.RELEASE_FN)
        # domain filling, splitting, calcpar + verttransform, convective mixing
        v.init_domainfill = cast(L.fpb_init_domainfill, a.INIT_DF_FN)
        v.boundcond_domainfill = cast(L.fpb_boundcond_domainfill, a.BOUNDCOND_FN)
        if device_release:
            v.split_particles = cast(L.fpb_split_particles, a.SPLIT_FN)
        v.set_vertical = cast(L.fpb_set_vertical, a.SET_VERTICAL_FN)
        v.calcpar_verttransform = cast(L.fpb_calcpar_verttransform, a.CALCPAR_VT_FN)
        v.set_convection = cast(L.fpb_set_convection, a.SET_CONV_FN)
        v.convmix = cast(L.fpb_convmix, a.CONVMIX_FN)
        return v
