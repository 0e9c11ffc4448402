"""ctypes mirror of include/fpb.h and include/fpb_host.h."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))

MAXSPEC, MAXAGECLASS, MAXZGRID, MAXRECEPTOR, MAXNESTS = 8, 8, 64, 20, 3
ABI_VERSION = 3
ITRA_DEAD = -999999999
RNG_REFERENCE, RNG_PHILOX_INDEX, RNG_PHILOX = 0, 1, 2
MATH_FAST, MATH_STRICT = 0, 1
SCATTER_ATOMIC, SCATTER_DETERMINISTIC = 0, 1

_f, _i = C.c_float, C.c_int32
_pf, _pi = C.POINTER(C.c_float), C.POINTER(C.c_int32)


class FpbError(RuntimeError):
    pass


class FpbConfig(C.Structure):
    _fields_ = [
        ("abi_version", _i),
        ("nx", _i), ("ny", _i), ("nz", _i), ("nxmax", _i), ("nymax", _i), ("nzmax", _i),
        ("nxmin1", _i), ("nymin1", _i),
        ("dx", _f), ("dy", _f), ("xlon0", _f), ("ylat0", _f), ("dxconst", _f), ("dyconst", _f),
        ("xglobal", _i), ("nglobal", _i), ("sglobal", _i),
        ("switchnorthg", _f), ("switchsouthg", _f),
        ("northpolemap", _f * 9), ("southpolemap", _f * 9), ("eps", _f),
        ("numbnests", _i), ("nxn", _i * MAXNESTS), ("nyn", _i * MAXNESTS), ("nxmaxn", _i), ("nymaxn", _i),
        ("xln", _f * MAXNESTS), ("yln", _f * MAXNESTS), ("xrn", _f * MAXNESTS), ("yrn", _f * MAXNESTS),
        ("xresoln", _f * MAXNESTS), ("yresoln", _f * MAXNESTS),
        ("ldirect", _i), ("lsynctime", _i), ("method", _i), ("mintime", _i), ("ifine", _i),
        ("turbswitch", _i), ("cblflag", _i), ("mdomainfill", _i), ("mquasilag", _i), ("lsettling", _i),
        ("ctl", _f), ("fine", _f), ("d_trop", _f), ("d_strat", _f), ("turbmesoscale", _f),
        ("turboff", _i), ("ind_samp", _i),
        ("ioutputforeachrelease", _i), ("lusekerneloutput", _i), ("lparticlecountoutput", _i),
        ("drydep", _i), ("drybkdep", _i), ("wetbkdep", _i), ("nested_output", _i),
        ("nspec", _i), ("decay", _f * MAXSPEC), ("drydepspec", _i * MAXSPEC),
        ("density", _f * MAXSPEC), ("dquer", _f * MAXSPEC), ("vsetaver", _f * MAXSPEC),
        ("cunningham", _f * MAXSPEC),
        ("wetdep", _i), ("wetdepspec", _i * MAXSPEC), ("weta_gas", _f * MAXSPEC), ("wetb_gas", _f * MAXSPEC),
        ("crain_aero", _f * MAXSPEC), ("csnow_aero", _f * MAXSPEC), ("ccn_aero", _f * MAXSPEC),
        ("in_aero", _f * MAXSPEC), ("henry", _f * MAXSPEC), ("readclouds", _i),
        ("readclouds_nest", _i * MAXNESTS),
        ("nageclass", _i), ("lage", _i * MAXAGECLASS),
        ("numxgrid", _i), ("numygrid", _i), ("numzgrid", _i),
        ("dxout", _f), ("dyout", _f), ("xoutshift", _f), ("youtshift", _f),
        ("outheight", _f * MAXZGRID),
        ("numxgridn", _i), ("numygridn", _i),
        ("dxoutn", _f), ("dyoutn", _f), ("xoutshiftn", _f), ("youtshiftn", _f),
        ("maxpointspec_act", _i), ("nclassunc", _i), ("maxageclass", _i), ("maxspec", _i),
        ("numreceptor", _i), ("xreceptor", _f * MAXRECEPTOR), ("yreceptor", _f * MAXRECEPTOR),
        ("receptorarea", _f * MAXRECEPTOR),
        ("numpoint", _i), ("npart", _pi), ("xmass", _pf),
        ("height", _pf),
        ("maxpart", _i), ("device", _i), ("rng_mode", _i), ("math_mode", _i), ("scatter_mode", _i),
        ("seed", C.c_uint64), ("part_id_stride", _i), ("part_id_offset", _i),
        ("sort_interval", _i), ("iflux", _i), ("ipout", _i), ("linit_cond", _i), ("reserved", _i * 4),
    ]


class FpbMetPtrs(C.Structure):
    _fields_ = [(n, _pf) for n in ("uu", "vv", "ww", "rho", "drhodz", "tt", "uupol", "vvpol",
                                   "hmix", "ustar", "wstar", "oli", "tropopause", "vdep",
                                   "lsprec", "convprec", "tcc", "ctwc")] + [("clouds", C.POINTER(C.c_int8))]


class FpbParticlePtrs(C.Structure):
    _fields_ = [
        ("xtra1", C.POINTER(C.c_double)), ("ytra1", C.POINTER(C.c_double)), ("ztra1", _pf),
        ("itra1", _pi), ("npoint", _pi), ("nclass", _pi), ("idt", _pi), ("itramem", _pi),
        ("itrasplit", _pi),
        ("uap", _pf), ("ucp", _pf), ("uzp", _pf), ("us", _pf), ("vs", _pf), ("ws", _pf),
        ("cbt", C.POINTER(C.c_int16)), ("xmass1", _pf), ("xscav_frac1", _pf), ("ld", _i),
    ]


class FpbStepStats(C.Structure):
    _fields_ = [(n, C.c_int64) for n in ("n_active", "n_init", "n_terminated", "n_pbl",
                                         "n_substeps", "n_petterssen", "n_nan_cbl", "n_nonfinite")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


class FpbhReleases(C.Structure):
    _fields_ = [("numpoint", _i), ("ireleasestart", _pi), ("ireleaseend", _pi),
                ("xpoint1", _pf), ("ypoint1", _pf), ("xpoint2", _pf), ("ypoint2", _pf),
                ("zpoint1", _pf), ("zpoint2", _pf), ("itsplit", _i)]


class FpbReleasePoints(C.Structure):
    _fields_ = FpbhReleases._fields_ + [("mp_pid", _i)]


class FpbPartoutPtrs(C.Structure):
    _fields_ = [("npoint", _pi), ("xlon", _pf), ("ylat", _pf), ("ztra1", _pf), ("itramem", _pi), ("topo", _pf),
                ("pvi", _pf), ("qvi", _pf), ("rhoi", _pf), ("hmixi", _pf), ("tri", _pf), ("tti", _pf),
                ("xmass1", _pf), ("ld", _i)]


class FpbPartavPtrs(C.Structure):
    _fields_ = [("npart_av", _pi)] + [(n, _pf) for n in ("cartx", "carty", "cartz", "z", "topo", "pv", "qv", "tt", "uu", "vv",
                                                          "rho", "tro", "hmix", "energy")]


class FpbDomainfillInfo(C.Structure):
    _fields_ = [("nx_we", _i * 2), ("ny_sn", _i * 2), ("gdomainfill", _i), ("numcolumn", _i),
                ("numparttot", _i), ("colmasstotal", _f), ("xmassperparticle", _f)]


class FpbConvPtrs(C.Structure):
    _fields_ = [(n, _pf) for n in ("ps", "tt2", "td2", "tth", "qvh")]


class FpbRawmetPtrs(C.Structure):
    _fields_ = [(n, _pf) for n in ("uuh", "vvh", "tth", "qvh", "pvh", "wwh", "ps", "tt2", "td2", "sshf", "surfstr",
                                   "lsprec", "convprec", "tcc", "excessoro", "clwch", "ciwch")]


class FpbMetOutPtrs(C.Structure):
    _fields_ = [(n, _pf) for n in ("uu", "vv", "ww", "rho", "drhodz", "tt", "qv", "pv", "uupol", "vvpol", "hmix",
                                   "ustar", "wstar", "oli", "tropopause")] + [("clouds", C.POINTER(C.c_int8)), ("ctwc", _pf)]


class FpbhRun(C.Structure):
    _fields_ = [("ideltas", _i), ("loutstep", _i), ("loutaver", _i), ("loutsample", _i),
                ("met_interval", _i), ("met_homogeneous", _i),
                ("met_u", _f), ("met_v", _f), ("met_w", _f), ("max_steps", _i), ("met_raw", _i), ("lconvection", _i)]


class FpbhRunResult(C.Structure):
    _fields_ = [("particle_steps", C.c_int64), ("substeps", C.c_int64), ("syncs", _i),
                ("outputs", _i), ("numpart_final", _i), ("t_step_s", C.c_double),
                ("t_conc_s", C.c_double), ("convmix_calls", _i), ("convecting_columns", _i),
                ("boundary_particles", _i), ("split_calls", _i)]


_pmet, _ppart = C.POINTER(FpbMetPtrs), C.POINTER(FpbParticlePtrs)
UPLOAD_MET_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, _i, _pmet)
SET_BRACKET_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, _pi, _pi, _i)
PUSH_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, _i, _i, _ppart)
SET_NUMPART_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, _i)
STEP_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, _i, _i, C.POINTER(FpbStepStats))
CONC_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, _i, _f)
FETCH_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, _pf, _pf, _pf, _pf, _pf, _i)
SCALE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, _pf)
WETDEPO_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, _i, _i, _i)
SET_RELEASES_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p)
RELEASE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, _i, _pi, _pi)
INIT_DF_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, _f, _f, _f, _f, _i, _pi, C.c_void_p)
BOUNDCOND_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, _i, _i, _pi, _pi)
SPLIT_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, _i, _pi)
SET_VERTICAL_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, _i, _i, _i, _i, _pf, _pf, _pf, _pf)
CALCPAR_VT_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, _i, C.c_void_p, _i, _pf)
SET_CONV_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, _i, _i, _i, _pf, _pf, _pf, _pf)
CONVMIX_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, _i, _pi, _pi)
OUTPUT_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, _i, _f, _pf, _pf, _pf, _pf, _pf)


class FpbhEngine(C.Structure):
    _fields_ = [("self", C.c_void_p), ("upload_met", UPLOAD_MET_FN),
                ("set_met_bracket", SET_BRACKET_FN), ("push_particles", PUSH_FN),
                ("pull_particles", PUSH_FN), ("set_numpart", SET_NUMPART_FN), ("step", STEP_FN),
                ("conccalc", CONC_FN), ("fetch_grids", FETCH_FN), ("scale_depgrids", SCALE_FN),
                ("wetdepo", WETDEPO_FN), ("set_releases", SET_RELEASES_FN), ("releaseparticles", RELEASE_FN),
                ("init_domainfill", INIT_DF_FN), ("boundcond_domainfill", BOUNDCOND_FN),
                ("split_particles", SPLIT_FN), ("set_vertical", SET_VERTICAL_FN),
                ("calcpar_verttransform", CALCPAR_VT_FN), ("set_convection", SET_CONV_FN), ("convmix", CONVMIX_FN)]


# FPB_ENGINE_LIB: load another build of the same library (kernel A/B experiments)
ENGINE_LIB = os.environ.get("FPB_ENGINE_LIB") or os.path.join(_HERE, "libfpb.so")
HOST_LIB = os.path.join(_HERE, "libfpb_host.so")
_engine = None
_host = None


def load_engine_lib():
    """Load libfpb.so.  Raises if it has not been built: there is no fallback."""
    global _engine
    if _engine is not None:
        return _engine
    if not os.path.exists(ENGINE_LIB):
        raise FpbError(f"{ENGINE_LIB} is missing: build it with __graft_entry__.build() "
                       "(nvcc, sm_100a); flexpart_b200 has no CPU fallback")
    L = C.CDLL(ENGINE_LIB)
    H = C.c_void_p
    L.fpb_last_error.restype = C.c_char_p
    L.fpb_abi_version.restype = C.c_int
    L.fpb_config_sizeof.restype = C.c_size_t
    L.fpb_init.argtypes = [C.POINTER(FpbConfig), C.POINTER(H)]
    L.fpb_finalize.argtypes = [H]
    L.fpb_set_rannumb.argtypes = [H, _pf, _i]
    L.fpb_fill_rannumb.argtypes = [H, _i, _i]
    L.fpb_upload_met.argtypes = [H, _i, _pmet]
    L.fpb_upload_met_nest.argtypes = [H, _i, _i, _pmet]
    L.fpb_upload_met_begin.argtypes = [H, _i, _pmet]
    L.fpb_upload_met_end.argtypes = [H, _pf]
    L.fpb_host_register.argtypes = [C.c_void_p, C.c_size_t]
    L.fpb_host_unregister.argtypes = [C.c_void_p]
    L.fpb_set_met_bracket.argtypes = [H, _pi, _pi, _i]
    L.fpb_push_particles.argtypes = [H, _i, _i, _ppart]
    L.fpb_pull_particles.argtypes = [H, _i, _i, _ppart]
    L.fpb_set_numpart.argtypes = [H, _i]
    L.fpb_step.argtypes = [H, _i, _i, C.POINTER(FpbStepStats)]
    L.fpb_wetdepo.argtypes = [H, _i, _i, _i]
    L.fpb_set_releases.argtypes = [H, C.POINTER(FpbReleasePoints)]
    L.fpb_set_outgrid_geometry.argtypes = [H, _pf, _pf, _pf, _pf]
    L.fpb_set_outgrid_origin.argtypes = [H, _f, _f, _f, _f]
    L.fpb_set_orography.argtypes = [H, _pf]
    L.fpb_upload_pvqv.argtypes = [H, _i, _pf, _pf]
    L.fpb_partoutput.argtypes = [H, _i, _pi, C.POINTER(FpbPartoutPtrs)]
    L.fpb_fetch_fluxes.argtypes = [H, _pf, _i]
    L.fpb_fetch_init_cond.argtypes = [H, _pf, _i]
    L.fpb_initial_cond_final.argtypes = [H, _i]
    L.fpb_fetch_partpos_average.argtypes = [H, _i, C.POINTER(FpbPartavPtrs), _i]
    L.fpb_concoutput_sparse.argtypes = [H, _i, _i, _i, _i, _i, _f, _f, _i, _pi, _pi, _pi, _pf]
    L.fpb_releaseparticles.argtypes = [H, _i, _pi, _pi]
    L.fpb_split_particles.argtypes = [H, _i, _pi]
    L.fpb_fetch_wetgrids.argtypes = [H, _pf, _pf]
    L.fpb_comm_unique_id.argtypes = [C.c_void_p]
    L.fpb_comm_init.argtypes = [H, C.c_void_p, _i, _i]
    L.fpb_reduce_grids_begin.argtypes = [H]
    L.fpb_reduce_grids_end.argtypes = [H, _pf, _pf, _pf, _pf, _pf, _pf, _pf]
    L.fpb_reduce_grids_device.argtypes = [H, _i, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t), _pf]
    L.fpb_comm_finalize.argtypes = [H]
    L.fpb_set_convection.argtypes = [H, _i, _i, _i, _pf, _pf, _pf, _pf]
    L.fpb_upload_convmet.argtypes = [H, _i, C.POINTER(FpbConvPtrs)]
    L.fpb_upload_convmet_nest.argtypes = [H, _i, _i, C.POINTER(FpbConvPtrs)]
    L.fpb_convmix.argtypes = [H, _i, _pi, _pi]
    L.fpb_set_vertical.argtypes = [H, _i, _i, _i, _i, _pf, _pf, _pf, _pf]
    L.fpb_calcpar_verttransform.argtypes = [H, _i, C.POINTER(FpbRawmetPtrs), _i, _pf]
    L.fpb_upload_vdep.argtypes = [H, _i, _pf]
    L.fpb_fetch_met.argtypes = [H, _i, C.POINTER(FpbMetOutPtrs)]
    L.fpb_calcpar_verttransform_nest.argtypes = [H, _i, _i, C.POINTER(FpbRawmetPtrs), _i, _f, _f, _f, _f, _pf]
    L.fpb_fetch_met_nest.argtypes = [H, _i, _i, C.POINTER(FpbMetOutPtrs)]
    L.fpb_init_domainfill.argtypes = [H, _f, _f, _f, _f, _i, _pi, C.POINTER(FpbDomainfillInfo)]
    L.fpb_boundcond_domainfill.argtypes = [H, _i, _i, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
    L.fpb_step_host.argtypes = [H, _i, _i, _i, C.POINTER(FpbParticlePtrs), C.c_float,
                                C.POINTER(FpbStepStats)]
    L.fpb_conccalc.argtypes = [H, _i, _f]
    L.fpb_fetch_grids.argtypes = [H, _pf, _pf, _pf, _pf, _pf, _i]
    L.fpb_scale_depgrids.argtypes = [H, _pf]
    L.fpb_grid_device_ptr.argtypes = [H, _i, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
    L.fpb_zero_conc_grids.argtypes = [H]
    L.fpb_sort_particles.argtypes = [H]
    L.fpb_stream.argtypes = [H]
    L.fpb_stream.restype = C.c_void_p
    L.fpb_launch_count.argtypes = [H]
    L.fpb_launch_count.restype = C.c_int64
    L.fpb_kernel_times.argtypes = [H, _pf, _pf]
    L.fpb_get_rannumb.argtypes = [H, _pf, _i]
    if L.fpb_config_sizeof() != C.sizeof(FpbConfig):
        raise FpbError(f"fpb_config layout mismatch: C {L.fpb_config_sizeof()} vs ctypes {C.sizeof(FpbConfig)}")
    _engine = L
    return L


def load_host_lib():
    global _host
    if _host is not None:
        return _host
    if not os.path.exists(HOST_LIB):
        raise FpbError(f"{HOST_LIB} is missing: build it with __graft_entry__.build()")
    L = C.CDLL(HOST_LIB)
    L.fpbh_last_error.restype = C.c_char_p
    L.fpbh_gridcheck.argtypes = [C.POINTER(FpbConfig)]
    L.fpbh_gridcheck_nest.argtypes = [C.POINTER(FpbConfig), _f, _f, _i, _i, _f, _f]
    L.fpbh_readcommand.argtypes = [C.POINTER(FpbConfig)]
    L.fpbh_readoutgrid.argtypes = [C.POINTER(FpbConfig), _f, _f, _i, _i, _f, _f, _pf, _i]
    L.fpbh_readoutgrid_nest.argtypes = [C.POINTER(FpbConfig), _f, _f, _i, _i, _f, _f]
    L.fpbh_synth_heights.argtypes = [_i, _pf]
    L.fpbh_synth_hybrid_levels.argtypes = [_i, _pf, _pf, _pf, _pf, _pi]
    L.fpbh_synth_rawmet.argtypes = [C.POINTER(FpbConfig), _i, _pf, _pf, _i, C.POINTER(FpbRawmetPtrs)]
    L.fpbh_verttransform_heights.argtypes = [C.POINTER(FpbConfig), _i, _pf, _pf, _pf, _pf, _pf, _pf, _pf, _pf, _pi, _pi]
    L.fpbh_synth_met.argtypes = [C.POINTER(FpbConfig), _pf, _i, _pmet]
    L.fpbh_synth_met_nest.argtypes = [C.POINTER(FpbConfig), _pf, _i, _i, _pmet]
    L.fpbh_homogeneous_met.argtypes = [C.POINTER(FpbConfig), _f, _f, _f, _pmet]
    L.fpbh_stlmbr.argtypes = [_pf, _f, _f]
    L.fpbh_stcm2p.argtypes = [_pf] + [_f] * 8
    L.fpbh_cc2gll.argtypes = [_pf, _f, _f, _f, _f, _pf, _pf]
    L.fpbh_cll2xy.argtypes = [_pf, _f, _f, _pf, _pf]
    L.fpbh_cxy2ll.argtypes = [_pf, _f, _f, _pf, _pf]
    L.fpbh_release_state_new.argtypes = [_i]
    L.fpbh_release_state_new.restype = C.c_void_p
    L.fpbh_release_state_free.argtypes = [C.c_void_p]
    L.fpbh_release_state_set_rank.argtypes = [C.c_void_p, _i]
    L.fpbh_release_state_set_rank.restype = None
    L.fpbh_outgrid_geometry.argtypes = [C.POINTER(FpbConfig), _i, _f, _pf, _pf]
    L.fpbh_releaseparticles.argtypes = [C.POINTER(FpbConfig), _pf, C.POINTER(FpbhReleases),
                                        C.c_void_p, _i, _ppart, _pi, _pi, _pi]
    L.fpbh_timemanager.argtypes = [C.POINTER(FpbConfig), _pf, C.POINTER(FpbhReleases),
                                   C.POINTER(FpbhRun), C.POINTER(FpbhEngine), OUTPUT_FN,
                                   C.c_void_p, C.POINTER(FpbhRunResult)]
    _host = L
    return L
