"""ctypes binding of the CPU oracle (oracle/liboracle.so).  Test infrastructure:
imported only by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs."""
import ctypes as C
import os
import subprocess

import numpy as np

from flexpart_b200 import abi
from flexpart_b200.abi import FpbConfig, FpbMetPtrs, FpbParticlePtrs, FpbStepStats, FpbhEngine

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
_pf, _pi = C.POINTER(C.c_float), C.POINTER(C.c_int32)
_libs = {}


def _fp(a):
    return a.ctypes.data_as(_pf) if a is not None else None


def load(libm_float=False):
    name = "liboracle_libmf.so" if libm_float else "liboracle.so"
    if name in _libs:
        return _libs[name]
    path = os.path.join(ORACLE_DIR, name)
    if not os.path.exists(path):
        subprocess.check_call(["make", "-C", ORACLE_DIR], stdout=subprocess.DEVNULL)
    L = C.CDLL(path)
    S = C.c_void_p
    L.fpo_create.argtypes = [C.POINTER(FpbConfig), C.c_int]
    L.fpo_create.restype = S
    L.fpo_destroy.argtypes = [S]
    L.fpo_ran1.argtypes = [S, _pi]; L.fpo_ran1.restype = C.c_float
    L.fpo_ran3.argtypes = [S, _pi]; L.fpo_ran3.restype = C.c_float
    L.fpo_fill_rannumb.argtypes = [S, C.c_int, C.c_int]
    L.fpo_set_rannumb.argtypes = [S, _pf, C.c_int]
    L.fpo_rannumb.argtypes = [S]; L.fpo_rannumb.restype = _pf
    L.fpo_set_index_uniforms.argtypes = [S, _pf, C.c_long]
    L.fpo_set_met.argtypes = [S, C.c_int, C.POINTER(FpbMetPtrs)]
    L.fpo_set_met_nest.argtypes = [S, C.c_int, C.c_int, C.POINTER(FpbMetPtrs)]
    L.fpo_set_met_bracket.argtypes = [S, _pi, _pi, C.c_int]
    L.fpo_push_particles.argtypes = [S, C.c_int, C.c_int, C.POINTER(FpbParticlePtrs)]
    L.fpo_pull_particles.argtypes = [S, C.c_int, C.c_int, C.POINTER(FpbParticlePtrs)]
    L.fpo_set_numpart.argtypes = [S, C.c_int]
    L.fpo_step.argtypes = [S, C.c_int, C.c_int, C.POINTER(FpbStepStats)]
    L.fpo_conccalc.argtypes = [S, C.c_int, C.c_float]
    L.fpo_wetdepo.argtypes = [S, C.c_int, C.c_int, C.c_int]
    L.fpo_fetch_wetgrids.argtypes = [S, _pf, _pf]
    L.fpo_fetch_grids.argtypes = [S, _pf, _pf, _pf, _pf, _pf, C.c_int]
    L.fpo_scale_depgrids.argtypes = [S, _pf]
    L.fpo_windalign.argtypes = [C.c_float] * 4 + [_pf, _pf]
    L.fpo_stlmbr.argtypes = [_pf, C.c_float, C.c_float]
    L.fpo_stcm2p.argtypes = [_pf] + [C.c_float] * 8
    L.fpo_cll2xy.argtypes = [_pf, C.c_float, C.c_float, _pf, _pf]
    L.fpo_cxy2ll.argtypes = [_pf, C.c_float, C.c_float, _pf, _pf]
    L.fpo_cgszll.argtypes = [_pf, C.c_float, C.c_float]; L.fpo_cgszll.restype = C.c_float
    L.fpo_cc2gll.argtypes = [_pf, C.c_float, C.c_float, C.c_float, C.c_float, _pf, _pf]
    L.fpo_releaseparticles.argtypes = [S, C.c_int, C.c_int, _pi, _pi] + [_pf] * 6 + [_pf, C.c_int]
    L.fpo_releaseparticles.restype = C.c_int
    L.fpo_split_particles.argtypes = [S, C.c_int]
    L.fpo_set_release_heights.argtypes = [S, _pf, _pf, C.c_int]
    L.fpo_init_domainfill.argtypes = [S, C.c_float, C.c_float, C.c_float, C.c_float, C.c_int, _pi, _pf]
    L.fpo_init_domainfill.restype = C.c_int
    L.fpo_boundcond_domainfill.argtypes = [S, C.c_int, C.c_int]
    L.fpo_boundcond_domainfill.restype = C.c_int
    L.fpo_boundcond_locations.argtypes = [S, C.POINTER(C.c_double)]
    L.fpo_boundcond_locations.restype = C.c_int
    L.fpo_numpart.argtypes = [S]; L.fpo_numpart.restype = C.c_int
    L.fpo_numparticlecount.argtypes = [S]; L.fpo_numparticlecount.restype = C.c_int
    L.fpo_outgrid_geometry.argtypes = [C.POINTER(FpbConfig), C.c_int, C.c_float, _pf, _pf]
    L.fpo_partoutput_record.argtypes = [C.POINTER(FpbConfig), _pf, C.c_int, _pi, C.c_double, C.c_double, C.c_float, _pf] + \
        [C.POINTER(_pf)] * 6 + [_pf]
    L.fpo_density_outgrid.argtypes = [C.POINTER(FpbConfig), _pf, C.c_int, C.c_float, C.c_float, _pf, _pf]
    L.fpo_concoutput_sparse.argtypes = [C.POINTER(FpbConfig), C.c_int, C.c_int, _pf, _pf, _pf, C.c_int, C.c_int, C.c_int,
                                        C.c_float, C.c_float, C.c_int, _pi, _pi, _pi, _pf]
    L.fpo_mp_step.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, C.c_float]
    L.fpo_mp_step.restype = C.c_double
    _libs[name] = L
    return L


class Oracle:
    """Sequential CPU restatement of the hot path with the Engine's interface."""

    def __init__(self, cb, strict_reference=False, libm_float=False):
        self.L = load(libm_float)
        self.cb = cb
        self.S = C.c_void_p(self.L.fpo_create(C.byref(cb.cfg), 1 if strict_reference else 0))
        self._keep = []

    def close(self):
        if self.S:
            self.L.fpo_destroy(self.S)
            self.S = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def fill_rannumb(self, maxrand=1000000, idummy=-320):
        self.L.fpo_fill_rannumb(self.S, maxrand, idummy)

    def set_rannumb(self, table):
        t = np.ascontiguousarray(table, np.float32)
        self.L.fpo_set_rannumb(self.S, _fp(t), len(t))

    def set_releases(self, rel, mp_pid=0):
        """the release heights the backward wet-scavenging block needs (src/timemanager.f90:590-591)"""
        self.L.fpo_set_release_heights(self.S, _fp(rel.zpoint1), _fp(rel.zpoint2), len(rel.zpoint1))

    def init_domainfill(self, box, itsplit=99999999):
        """init_domainfill over the box (xpoint1, ypoint1, xpoint2, ypoint2) in grid units;
        returns (numpart, info) like Engine.init_domainfill"""
        out, fout = np.zeros(8, np.int32), np.zeros(2, np.float32)
        rc = self.L.fpo_init_domainfill(self.S, box[0], box[1], box[2], box[3], itsplit,
                                        out.ctypes.data_as(_pi), _fp(fout))
        if rc:
            raise RuntimeError("init_domainfill: numpart exceeds maxpart")
        return self.L.fpo_numpart(self.S), dict(nx_we=(int(out[0]), int(out[1])), ny_sn=(int(out[2]), int(out[3])),
                                                gdomainfill=int(out[4]), numcolumn=int(out[5]), numparttot=int(out[6]),
                                                colmasstotal=float(fout[0]), xmassperparticle=float(fout[1]))

    def boundcond_domainfill(self, itime, itsplit=99999999):
        """boundcond_domainfill(itime); returns the number of particles created"""
        n = self.L.fpo_boundcond_domainfill(self.S, itime, itsplit)
        if n < 0:
            raise RuntimeError("boundcond_domainfill: too many particles required")
        return n

    def numpart(self):
        return self.L.fpo_numpart(self.S)

    def numparticlecount(self):
        return self.L.fpo_numparticlecount(self.S)

    def boundcond_locations(self):
        acc = C.c_double(0.0)
        n = self.L.fpo_boundcond_locations(self.S, C.byref(acc))
        return n, acc.value

    def set_index_uniforms(self, u):
        """validation hook: the uniforms behind the next nrand draws (tests/philox_ref.py)"""
        if u is None:
            self.L.fpo_set_index_uniforms(self.S, None, 0)
        else:
            u = np.ascontiguousarray(u, np.float32)
            self.L.fpo_set_index_uniforms(self.S, _fp(u), len(u))

    def rannumb(self, n):
        return np.ctypeslib.as_array(self.L.fpo_rannumb(self.S), shape=(n,)).copy()

    def upload_met(self, slot, met):
        self._keep.append(met)  # the oracle reads the host arrays in place
        self._keep = self._keep[-4:]
        self.L.fpo_set_met(self.S, slot, C.byref(met.ptrs))

    def upload_met_nest(self, slot, nest, met):
        self._keep_n = getattr(self, "_keep_n", {})
        self._keep_n[(slot, nest)] = met
        self.L.fpo_set_met_nest(self.S, slot, nest, C.byref(met.ptrs))

    def set_met_bracket(self, memind, memtime, lwindinterv=None):
        if lwindinterv is None:
            lwindinterv = abs(memtime[1] - memtime[0])
        self.L.fpo_set_met_bracket(self.S, (C.c_int32 * 2)(*memind), (C.c_int32 * 2)(*memtime), lwindinterv)

    def push_particles(self, parts, first=0, count=None):
        count = parts.numpart - first if count is None else count
        self.L.fpo_push_particles(self.S, first, count, C.byref(parts.ptrs))

    def pull_particles(self, parts, first=0, count=None):
        count = parts.numpart - first if count is None else count
        self.L.fpo_pull_particles(self.S, first, count, C.byref(parts.ptrs))

    def set_numpart(self, n):
        self.L.fpo_set_numpart(self.S, n)

    def step(self, itime, ldeltat=0, stats=True):
        st = FpbStepStats()
        self.L.fpo_step(self.S, itime, ldeltat, C.byref(st))
        return st.as_dict()

    def conccalc(self, itime, weight):
        self.L.fpo_conccalc(self.S, itime, weight)

    def wetdepo(self, itime, ltsample, ldeltat=0):
        self.L.fpo_wetdepo(self.S, itime, ltsample, ldeltat)

    def fetch_wetgrids(self):
        c = self.cb.cfg
        sd = (c.numxgrid, c.numygrid, c.maxspec, c.maxpointspec_act, c.nclassunc, c.maxageclass)
        out = {"wetgridunc": np.zeros(sd, np.float32, order="F")}
        wn = None
        if c.nested_output == 1:
            out["wetgriduncn"] = np.zeros((c.numxgridn, c.numygridn) + sd[2:], np.float32, order="F")
            wn = out["wetgriduncn"]
        self.L.fpo_fetch_wetgrids(self.S, _fp(out["wetgridunc"]), _fp(wn))
        return out

    def fetch_grids(self, zero_conc=True):
        c = self.cb.cfg
        sg = (c.numxgrid, c.numygrid, c.numzgrid, c.maxspec, c.maxpointspec_act, c.nclassunc, c.maxageclass)
        sd = (c.numxgrid, c.numygrid, c.maxspec, c.maxpointspec_act, c.nclassunc, c.maxageclass)
        out = {"gridunc": np.zeros(sg, np.float32, order="F"), "drygridunc": np.zeros(sd, np.float32, order="F"),
               "creceptor": np.zeros((abi.MAXRECEPTOR, c.maxspec), np.float32, order="F")}
        gn = dn = None
        if c.nested_output == 1:
            sgn = (c.numxgridn, c.numygridn) + sg[2:]
            sdn = (c.numxgridn, c.numygridn) + sd[2:]
            out["griduncn"] = np.zeros(sgn, np.float32, order="F")
            out["drygriduncn"] = np.zeros(sdn, np.float32, order="F")
            gn, dn = out["griduncn"], out["drygriduncn"]
        self.L.fpo_fetch_grids(self.S, _fp(out["gridunc"]), _fp(gn), _fp(out["drygridunc"]), _fp(dn),
                               _fp(out["creceptor"]), 1 if zero_conc else 0)
        return out

    def scale_depgrids(self, factors):
        f = np.ascontiguousarray(factors, np.float32)
        self.L.fpo_scale_depgrids(self.S, _fp(f))

    def vtable(self):
        L, a = self.L, abi
        v = FpbhEngine()
        v.self = self.S
        cast = lambda fn, T: C.cast(fn, T)
        v.upload_met = cast(L.fpo_vt_upload_met, a.UPLOAD_MET_FN)
        v.set_met_bracket = cast(L.fpo_vt_set_met_bracket, a.SET_BRACKET_FN)
        v.push_particles = cast(L.fpo_vt_push_particles, a.PUSH_FN)
        v.pull_particles = cast(L.fpo_vt_pull_particles, a.PUSH_FN)
        v.set_numpart = cast(L.fpo_vt_set_numpart, a.SET_NUMPART_FN)
        v.step = cast(L.fpo_vt_step, a.STEP_FN)
        v.conccalc = cast(L.fpo_vt_conccalc, a.CONC_FN)
        v.fetch_grids = cast(L.fpo_vt_fetch_grids, a.FETCH_FN)
        v.scale_depgrids = cast(L.fpo_vt_scale_depgrids, a.SCALE_FN)
        v.wetdepo = cast(L.fpo_vt_wetdepo, a.WETDEPO_FN)
        v.init_domainfill = cast(L.fpo_vt_init_domainfill, a.INIT_DF_FN)
        v.boundcond_domainfill = cast(L.fpo_vt_boundcond_domainfill, a.BOUNDCOND_FN)
        v.split_particles = cast(L.fpo_vt_split_particles, a.SPLIT_FN)
        return v
