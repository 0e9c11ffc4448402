"""ctypes harness for oracle/_ref/libflexref.so: the reference's own Fortran
sources for the hot path, transpiled to C by oracle/f2c/f90toc.py and compiled
here (test infrastructure; see oracle/f2c/Makefile).  The module variables of
com_mod / par_mod / interpol_mod / hanna_mod / unc_mod / outg_mod / point_mod
are C globals reached by name through ref_ptr()."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_LIB = os.path.join(ROOT, "oracle", "_ref", "libflexref.so")
_CT = {"int": C.c_int, "float": C.c_float, "double": C.c_double, "short": C.c_short,
       "signed char": C.c_int8, "long long": C.c_longlong}
_NP = {"int": np.int32, "float": np.float32, "double": np.float64, "short": np.int16,
       "signed char": np.int8, "long long": np.int64}


def available():
    if os.path.isdir("/root/reference/src"):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle", "f2c")], stdout=subprocess.DEVNULL)
    return os.path.exists(REF_LIB)


class Ref:
    """One instance of the transpiled reference (process-global state, like the
    Fortran modules it comes from: use one configuration per process run)."""

    def __init__(self, cb, maxrand=20000):
        # the library carries the Fortran modules' global state and the SAVEd locals
        # (ran3's table, the idummy of advance/initialize, ...): load a private copy per instance
        import shutil
        import tempfile
        self._tmp = tempfile.NamedTemporaryFile(suffix=".so", delete=False)
        self._tmp.close()
        shutil.copyfile(REF_LIB, self._tmp.name)
        L = self.L = C.CDLL(self._tmp.name)
        os.unlink(self._tmp.name)
        L.ref_ptr.restype = C.c_void_p
        L.ref_ptr.argtypes = [C.c_char_p]
        L.ref_extent.restype = C.c_long
        L.ref_extent.argtypes = [C.c_char_p, C.c_int]
        L.ref_type.restype = C.c_char_p
        L.ref_type.argtypes = [C.c_char_p]
        self.cb = cb
        c = cb.cfg
        L.ref_defaults()
        # par_mod extents (compile-time parameters in the reference, run-time here)
        for k, v in dict(nxmax=c.nxmax, nymax=c.nymax, nzmax=c.nzmax, maxpart=c.maxpart, maxspec=c.maxspec,
                         maxnests=max(c.numbnests, 1), nxmaxn=max(c.nxmaxn, 1), nymaxn=max(c.nymaxn, 1),
                         maxageclass=c.maxageclass, nclassunc=c.nclassunc, maxrand=maxrand, numwfmem=2).items():
            self.set(k, v)
        for k, v in dict(nuvzmax=c.nzmax, nwzmax=c.nzmax, nconvlevmax=c.nzmax - 1, na=c.nzmax).items():
            if self.has(k):     # convection (conv_mod extents)
                self.set(k, v)
        # extents of the allocatable arrays
        for k in ("numxgrid", "numygrid", "numzgrid", "maxpointspec_act", "numpoint", "numreceptor"):
            self.set(k, getattr(c, k))
        for k in ("nspec", "nageclass"):          # extents of flux_mod's flux
            self.set(k, getattr(c, k))
        self.set("numxgridn", max(c.numxgridn, 1)); self.set("numygridn", max(c.numygridn, 1))
        L.ref_alloc()
        self.set("numxgridn", c.numxgridn); self.set("numygridn", c.numygridn)
        # com_mod scalars the path reads
        for k in ("nxmin1", "nymin1", "dx", "dy", "xlon0", "ylat0", "dxconst", "dyconst", "xglobal", "nglobal",
                  "sglobal", "switchnorthg", "switchsouthg", "ldirect", "lsynctime", "method", "mintime", "ifine",
                  "turbswitch", "cblflag", "mdomainfill", "lsettling", "ctl", "fine", "d_trop", "d_strat",
                  "turbmesoscale", "ind_samp", "ioutputforeachrelease", "drydep", "drybkdep", "wetbkdep",
                  "nested_output", "nspec", "nageclass", "dxout", "dyout", "xoutshift", "youtshift", "dxoutn",
                  "dyoutn", "xoutshiftn", "youtshiftn", "numbnests"):
            self.set(k, getattr(c, k))
        self.set("nz", c.nz)
        for k in ("nx", "ny"):
            if self.has(k):
                self.set(k, getattr(c, k))
        for l in range(c.numbnests):
            for k in ("nxn", "nyn"):
                if self.has(k):
                    self.arr(k)[l] = getattr(c, k)[l]
        self.set("nmixz", c.nz)
        assert c.turboff == 0, "turboff is a compile-time .false. in the reference (src/com_mod.f90:778)"
        assert c.lusekerneloutput == 1 and c.lparticlecountoutput == 0, "par_mod compile-time switches"
        self.arr("northpolemap")[:] = list(c.northpolemap)
        self.arr("southpolemap")[:] = list(c.southpolemap)
        self.arr("height")[:c.nz] = cb.height[:c.nz]
        self.arr("lage")[:c.nageclass] = [c.lage[k] for k in range(c.nageclass)]
        self.arr("outheight")[:] = [c.outheight[k] for k in range(c.numzgrid)]
        for nm in ("density", "dquer", "vsetaver", "cunningham"):
            self.arr(nm)[:c.nspec] = [getattr(c, nm)[k] for k in range(c.nspec)]
        self.arr("drydepspec")[:c.nspec] = [c.drydepspec[k] for k in range(c.nspec)]
        if self.has("npart"):
            self.arr("npart")[:] = cb.npart
        for k, v in dict(mquasilag=0, ipout=0, iflux=0, linit_cond=0, verbosity=0).items():
            if self.has(k):
                self.set(k, v)
        self.arr("xmass")[:, :] = cb.xmass[:, :c.maxspec]
        for nm in ("xreceptor", "yreceptor", "receptorarea"):
            self.arr(nm)[:c.numreceptor] = [getattr(c, nm)[k] for k in range(c.numreceptor)]
        for l in range(c.numbnests):
            for nm in ("xln", "yln", "xrn", "yrn", "xresoln", "yresoln"):
                a = self.arr(nm)
                a[l + (1 if nm.endswith("resoln") else 0)] = getattr(c, nm)[l]  # xresoln(0:maxnests)
        # wet deposition (readspecies / readreleases)
        if self.has("wetdepspec"):
            for nm in ("wetdepspec", "weta_gas", "wetb_gas", "crain_aero", "csnow_aero", "ccn_aero", "in_aero", "henry",
                       "decay"):
                self.arr(nm)[:c.nspec] = [getattr(c, nm)[k] for k in range(c.nspec)]
            self.set("readclouds", c.readclouds)
            for l in range(c.numbnests):
                self.arr("readclouds_nest")[l] = c.readclouds_nest[l]
            self.set("loutstep", 3600)
        self.numpart = 0

    # ---- access by name
    def has(self, name):
        return bool(self.L.ref_ptr(name.encode()))

    def _ptr(self, name):
        p = self.L.ref_ptr(name.encode())
        if not p:
            raise KeyError(name)
        return p, self.L.ref_type(name.encode()).decode()

    def set(self, name, value):
        p, t = self._ptr(name)
        C.cast(p, C.POINTER(_CT[t]))[0] = value

    def get(self, name):
        p, t = self._ptr(name)
        return C.cast(p, C.POINTER(_CT[t]))[0]

    def arr(self, name):
        """numpy view (Fortran order) of a module array"""
        p, t = self._ptr(name)
        shape, d = [], 0
        while True:
            e = self.L.ref_extent(name.encode(), d)
            if e < 0:
                break
            shape.append(int(e))
            d += 1
        n = int(np.prod(shape))
        flat = np.ctypeslib.as_array(C.cast(p, C.POINTER(_CT[t])), shape=(n,))
        return flat.reshape(shape, order="F")

    # ---- meteorology: the slices (:,:,:,slot[,nest]) of the com_mod arrays
    def upload_met(self, slot, m):
        for nm in ("uu", "vv", "ww", "rho", "drhodz", "tt", "uupol", "vvpol"):
            self.arr(nm)[:, :, :, slot - 1] = getattr(m, nm)
        for nm in ("hmix", "ustar", "wstar", "oli", "tropopause"):
            self.arr(nm)[:, :, 0, slot - 1] = getattr(m, nm)
        self.arr("vdep")[:, :, :, slot - 1] = m.vdep

    def upload_rain(self, slot, m, nest=0):
        sfx, idx = ("", (slot - 1,)) if nest == 0 else ("n", (slot - 1, nest - 1))
        for nm in ("lsprec", "convprec", "tcc"):
            self.arr(nm + sfx)[(slice(None), slice(None), 0) + idx] = getattr(m, nm)
        self.arr("ctwc" + sfx)[(slice(None), slice(None)) + idx] = m.ctwc
        self.arr("clouds" + sfx)[(slice(None), slice(None), slice(None)) + idx] = m.clouds
        if nest:
            self.arr("ttn")[(slice(None), slice(None), slice(None)) + idx] = m.tt

    def upload_met_nest(self, slot, nest, m):
        for nm in ("uu", "vv", "ww", "rho", "drhodz"):
            self.arr(nm + "n")[:, :, :, slot - 1, nest - 1] = getattr(m, nm)
        for nm in ("hmix", "ustar", "wstar", "oli", "tropopause"):
            self.arr(nm + "n")[:, :, 0, slot - 1, nest - 1] = getattr(m, nm)
        self.arr("vdepn")[:, :, :, slot - 1, nest - 1] = m.vdep

    def set_met_bracket(self, memind, memtime, lwindinterv=None):
        self.arr("memind")[:2] = memind
        self.arr("memtime")[:2] = memtime
        self.set("lwindinterv", abs(memtime[1] - memtime[0]) if lwindinterv is None else lwindinterv)

    # ---- FLEXPART.f90:47,56-59: the rannumb table
    def fill_rannumb(self, idummy=-320):
        n = self.get("maxrand")
        tab = self.arr("rannumb")
        L = self.L
        idum = C.c_int(idummy)
        a, b = C.c_float(), C.c_float()
        for j in range(1, n, 2):          # do j=1,maxrand-1,2: gasdev1(idummy,rannumb(j),rannumb(j+1))
            L.f_gasdev1(C.byref(idum), C.byref(a), C.byref(b))
            tab[j - 1], tab[j] = a.value, b.value
        L.f_gasdev1(C.byref(idum), C.byref(a), C.byref(b))   # gasdev1(idummy,rannumb(maxrand),rannumb(maxrand-1))
        tab[n - 1], tab[n - 2] = a.value, b.value
        return tab

    # ---- the two calls of the particle loop, src/timemanager.f90:553-555,609-611
    def initialize(self, itime, ldt, xt, yt, zt):
        i = lambda v: C.c_int(v)
        f = lambda: C.c_float(0.0)
        ldt_c, up, vp, wp, us, vs, ws = i(ldt), f(), f(), f(), f(), f(), f()
        icbt = C.c_short(0)
        self.L.f_initialize(C.byref(i(itime)), C.byref(ldt_c), C.byref(up), C.byref(vp), C.byref(wp), C.byref(us),
                            C.byref(vs), C.byref(ws), C.byref(C.c_double(xt)), C.byref(C.c_double(yt)),
                            C.byref(C.c_float(zt)), C.byref(icbt))
        return ldt_c.value, up.value, vp.value, wp.value, us.value, vs.value, ws.value, icbt.value

    def advance(self, itime, nrelpoint, ldt, up, vp, wp, us, vs, ws, xt, yt, zt, icbt):
        cf = C.c_float
        a = dict(ldt=C.c_int(ldt), up=cf(up), vp=cf(vp), wp=cf(wp), us=cf(us), vs=cf(vs), ws=cf(ws), nstop=C.c_int(0),
                 xt=C.c_double(xt), yt=C.c_double(yt), zt=cf(zt), icbt=C.c_short(icbt))
        prob = (cf * max(self.cb.cfg.maxspec, 1))()
        self.L.f_advance(C.byref(C.c_int(itime)), C.byref(C.c_int(nrelpoint)), C.byref(a["ldt"]), C.byref(a["up"]),
                         C.byref(a["vp"]), C.byref(a["wp"]), C.byref(a["us"]), C.byref(a["vs"]), C.byref(a["ws"]),
                         C.byref(a["nstop"]), C.byref(a["xt"]), C.byref(a["yt"]), C.byref(a["zt"]), prob,
                         C.byref(a["icbt"]))
        out = {k: v.value for k, v in a.items()}
        out["prob"] = np.array(prob[:], np.float32)
        return out

    # ---- particles for conccalc
    def push_particles(self, p):
        n = p.numpart
        self.numpart = n
        self.set("numpart", n)
        for nm in ("xtra1", "ytra1", "ztra1", "itra1", "itramem", "npoint", "nclass"):
            self.arr(nm)[:n] = getattr(p, nm)[:n]
        self.arr("xmass1")[:n, :p.nspec] = p.xmass1[:n]
        self.arr("xscav_frac1")[:n, :p.nspec] = p.xscav_frac1[:n]

    STATE = ("xtra1", "ytra1", "ztra1", "itra1", "itramem", "npoint", "nclass", "idt", "uap", "ucp", "uzp",
             "us", "vs", "ws", "cbt")

    def push_state(self, p):
        """every particle array of com_mod the loop reads (src/com_mod.f90:675-695)"""
        n = p.numpart
        self.set("numpart", n)
        for nm in self.STATE:
            self.arr(nm)[:n] = getattr(p, nm)[:n]
        self.arr("xmass1")[:n, :p.nspec] = p.xmass1[:n]
        self.arr("xscav_frac1")[:n, :p.nspec] = p.xscav_frac1[:n]

    def pull_state(self, p):
        n = p.numpart
        for nm in self.STATE:
            getattr(p, nm)[:n] = self.arr(nm)[:n]
        p.xmass1[:n] = self.arr("xmass1")[:n, :p.nspec]

    def particle_loop(self, itime, ldeltat):
        """src/timemanager.f90:531-712"""
        self.L.f_tm_particle_loop(C.byref(C.c_int(itime)), C.byref(C.c_int(ldeltat)))

    def wetdepo(self, itime, ltsample, loutnext):
        self.L.f_wetdepo(C.byref(C.c_int(itime)), C.byref(C.c_int(ltsample)), C.byref(C.c_int(loutnext)))

    def conccalc(self, itime, weight):
        self.L.f_conccalc(C.byref(C.c_int(itime)), C.byref(C.c_float(weight)))
