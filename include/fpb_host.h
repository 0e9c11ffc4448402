/*
 * fpb_host.h -- C ABI of the host-side companion library (libfpb_host.so).
 *
 * The reference's host is Fortran (no Fortran compiler exists in the build
 * image), so the host pieces a run needs on top of the engine
 * are provided in C++ with the reference's own names and semantics:
 *   - run-constant derivation: gridcheck_ecmwf (global/pole flags, pole
 *     maps), readcommand (CTL/IFINE/method switches), readoutgrid (shifts);
 *   - a synthetic "ECMWF-shaped" met generator that fills the arrays
 *     verttransform_ecmwf/calcpar would produce (incl. drhodz and uupol/vvpol);
 *   - releaseparticles (release counts, slot search, ran1 position stream);
 *   - timemanager: the reference's time loop driving an engine through a
 *     table of function pointers with the fpb_* signatures.
 * CPU only; needs no CUDA device.
 */
#ifndef FPB_HOST_H
#define FPB_HOST_H

#include "fpb.h"

#ifdef __cplusplus
extern "C" {
#endif

const char *fpbh_last_error(void);

/* gridcheck_ecmwf: from nx,ny,nz,nxmax,nymax,nzmax,dx,dy,xlon0,ylat0 derive
 * nxmin1,nymin1,xglobal,nglobal,sglobal,switchnorthg/southg,north/southpolemap,
 * dxconst,dyconst,eps (src/gridcheck_ecmwf.f90:300-366, src/par_mod.f90:123,
 * src/advance.f90:107). */
int fpbh_gridcheck(fpb_config *cfg);
/* gridcheck_nests (src/gridcheck_nests.f90:359-389): appends one nested input
 * grid (after fpbh_gridcheck): numbnests, nxn, nyn, nxmaxn, nymaxn, xresoln,
 * yresoln, xln, xrn, yln, yrn; "Nested domain does not fit into mother domain"
 * is an error as in the reference. */
int fpbh_gridcheck_nest(fpb_config *cfg, float xlon0n, float ylat0n, int32_t nxn, int32_t nyn,
                        float dxn, float dyn);

/* readcommand derivations (src/readcommand.f90:244-272,377-383,622-634):
 * in: ldirect, lsynctime (>0 as in COMMAND), ctl (as in COMMAND), ifine,
 * cblflag; out: ifine, turbswitch, fine, ctl:=1/ctl, method, mintime,
 * lsynctime sign, d_trop, d_strat, turbmesoscale defaults. */
int fpbh_readcommand(fpb_config *cfg);

/* readoutgrid (src/readoutgrid.f90:199-200): xoutshift=xlon0-outlon0 etc. */
int fpbh_readoutgrid(fpb_config *cfg, float outlon0, float outlat0, int32_t numxgrid,
                     int32_t numygrid, float dxout, float dyout, const float *outheights,
                     int32_t numzgrid);
int fpbh_readoutgrid_nest(fpb_config *cfg, float outlon0n, float outlat0n, int32_t numxgridn,
                          int32_t numygridn, float dxoutn, float dyoutn);

/* 137-layer-like level heights: height[0]=0, strictly increasing, ~80 km top
 * (shape of src/verttransform_ecmwf.f90:153-168 output). */
int fpbh_synth_heights(int32_t nz, float *height);

/* height(1:nuvz) as the first call of verttransform_ecmwf derives it (src/verttransform_ecmwf.f90:
 * 131-163): heights of the eta levels of the first grid point (south to north, west to east) whose
 * surface pressure exceeds 1000 hPa.  For callers that leave calcpar + verttransform to the device
 * (fpb_calcpar_verttransform): fpb_config::height must be this array.  akz, bkz: the Fortran arrays
 * (1:nuvz); ps, tt2, td2 (nxmax,nymax); tth, qvh (nxmax,nymax,nuvzmax); *ixm, *jym may be NULL. */
int fpbh_verttransform_heights(const fpb_config *cfg, int32_t nuvz, const float *akz, const float *bkz,
                               const float *ps, const float *tt2, const float *td2, const float *tth,
                               const float *qvh, float *height, int32_t *ixm, int32_t *jym);

/* Synthetic hybrid coefficients (L137-like; akm, bkm half levels, akz, bkz layer centres with level 1
 * the surface, src/gridcheck_ecmwf.f90:470-566) and one time level of synthetic model-level fields
 * (uuh, vvh, wwh, tth, qvh, ps, tt2, td2, sshf, surfstr, [lsprec, convprec, tcc]) in the reference's
 * padded layout: the input of fpb_calcpar_verttransform when the run has no GRIB files.  Arrays of
 * the coefficients are the Fortran (1:nuvz), 0-based. */
int fpbh_synth_hybrid_levels(int32_t nuvz, float *akm, float *bkm, float *akz, float *bkz, int32_t *nconvlev);
int fpbh_synth_rawmet(const fpb_config *cfg, int32_t nuvz, const float *akz, const float *bkz, int32_t time_s,
                      const fpb_rawmet_ptrs *out);

/* Fill one time level of synthetic met into caller arrays with
 * the padded Fortran layout; `time_s` moves the phase of the fields.
 * All non-NULL pointers of `out` are written. */
int fpbh_synth_met(const fpb_config *cfg, const float *height, int32_t time_s,
                   const fpb_met_ptrs *out);
/* the same fields on nested input grid `nest` (1..numbnests), padded to
 * (nxmaxn, nymaxn, nzmax) */
int fpbh_synth_met_nest(const fpb_config *cfg, const float *height, int32_t time_s, int32_t nest,
                        const fpb_met_ptrs *out);
/* homogeneous fields of mpi_mod's set_fields_synthetic (src/mpi_mod.f90:2940-2973) */
int fpbh_homogeneous_met(const fpb_config *cfg, float u, float v, float w,
                         const fpb_met_ptrs *out);

/* conformal-map helpers (src/cmapf_mod.f90), exported for tests */
void fpbh_stlmbr(float *strcmp, float tnglat, float xlong);
void fpbh_stcm2p(float *strcmp, float x1, float y1, float xlat1, float xlong1, float x2,
                 float y2, float xlat2, float xlong2);
void fpbh_cc2gll(const float *strcmp, float xlat, float xlong, float ue, float vn, float *ug,
                 float *vg);
void fpbh_cll2xy(const float *strcmp, float xlat, float xlong, float *x, float *y);
void fpbh_cxy2ll(const float *strcmp, float x, float y, float *xlat, float *xlong);

/* RELEASES after coordtrafo: boxes in grid units (src/coordtrafo.f90:37-42),
 * times relative to the simulation start and snapped to lsynctime
 * (src/FLEXPART.f90:401-404). */
typedef struct fpbh_releases {
  int32_t numpoint;
  const int32_t *ireleasestart, *ireleaseend;
  const float *xpoint1, *ypoint1, *xpoint2, *ypoint2, *zpoint1, *zpoint2;
  int32_t itsplit;
} fpbh_releases;

/* outgrid_init's cell areas and volumes (src/outgrid_init.f90:48-100; nest = 1:
 * src/outgrid_init_nest.f90:84-117), the inputs of fpb_set_outgrid_geometry.
 * outlat0: southern edge of the (nested) output grid in degrees. */
int fpbh_outgrid_geometry(const fpb_config *cfg, int32_t nest, float outlat0, float *area, float *volume);

typedef struct fpbh_release_state fpbh_release_state;
fpbh_release_state *fpbh_release_state_new(int32_t numpoint);
void fpbh_release_state_free(fpbh_release_state *s);
/* MPI build: rank mp_pid > 0 offsets the SAVEd idummy of releaseparticles by mp_seed
 * (src/mpi_mod.f90:331-335, src/releaseparticles_mpi.f90:56-65).  Call before the first release. */
void fpbh_release_state_set_rank(fpbh_release_state *s, int32_t mp_pid);

/* releaseparticles(itime), src/releaseparticles.f90:69-378 (zkind 1,
 * EMISVAR factors 1).  Works on the host mirror `p` (itra1 must be current).
 * On return first_changed and n_changed bound the rows that were written and
 * *numpart is updated.  Returns 1 when maxpart is exceeded (label 996). */
int fpbh_releaseparticles(const fpb_config *cfg, const float *height, const fpbh_releases *rel,
                          fpbh_release_state *st, int32_t itime, const fpb_particle_ptrs *p,
                          int32_t *numpart, int32_t *first_changed, int32_t *n_changed);

/* engine seen by the time loop: the fpb_* entry points (or any
 * implementation with the same contract), `self` passed as first argument */
typedef struct fpbh_engine {
  void *self;
  int (*upload_met)(void *self, int32_t slot, const fpb_met_ptrs *met);
  int (*set_met_bracket)(void *self, const int32_t memind[2], const int32_t memtime[2],
                         int32_t lwindinterv);
  int (*push_particles)(void *self, int32_t first, int32_t count, const fpb_particle_ptrs *p);
  int (*pull_particles)(void *self, int32_t first, int32_t count, const fpb_particle_ptrs *p);
  int (*set_numpart)(void *self, int32_t numpart);
  int (*step)(void *self, int32_t itime, int32_t ldeltat, fpb_step_stats *stats);
  int (*conccalc)(void *self, int32_t itime, float weight);
  int (*fetch_grids)(void *self, float *gridunc, float *griduncn, float *drygridunc,
                     float *drygriduncn, float *creceptor, int32_t zero_conc);
  int (*scale_depgrids)(void *self, const float *factor);
  /* may be NULL: wet deposition is then left to the caller (src/timemanager.f90:164-169) */
  int (*wetdepo)(void *self, int32_t itime, int32_t ltsample, int32_t ldeltat);
  /* may both be NULL: releaseparticles then runs on the host mirror (fpbh_releaseparticles) and the
   * new rows are pushed; otherwise the engine creates the particles itself (fpb_releaseparticles) */
  int (*set_releases)(void *self, const fpb_release_points *rel);
  int (*releaseparticles)(void *self, int32_t itime, int32_t *numpart, int32_t *n_released);
  /* domain-filling runs (cfg->mdomainfill >= 1, src/timemanager.f90:230-241): both required then */
  int (*init_domainfill)(void *self, float xpoint1, float ypoint1, float xpoint2, float ypoint2, int32_t itsplit,
                         int32_t *numpart, fpb_domainfill_info *info);
  int (*boundcond_domainfill)(void *self, int32_t itime, int32_t loutend, int32_t *numpart, int32_t *n_created);
  /* may be NULL: no particle splitting (src/timemanager.f90:472-503) */
  int (*split_particles)(void *self, int32_t itime, int32_t *numpart);
  /* fpbh_run::met_raw: the wind fields are model-level fields and the engine does calcpar +
   * verttransform (src/getfields.f90:126-129) */
  int (*set_vertical)(void *self, int32_t nuvz, int32_t nwz, int32_t nuvzmax, int32_t nwzmax, const float *akm,
                      const float *bkm, const float *akz, const float *bkz);
  int (*calcpar_verttransform)(void *self, int32_t slot, const fpb_rawmet_ptrs *raw, int32_t lsubgrid, float *device_ms);
  /* fpbh_run::lconvection (needs met_raw: the convection scheme reads the model-level fields) */
  int (*set_convection)(void *self, int32_t nuvz, int32_t nuvzmax, int32_t nconvlev, const float *akz, const float *bkz,
                        const float *akm, const float *bkm);
  int (*convmix)(void *self, int32_t itime, int32_t *ncolumns, int32_t *nconvecting);
} fpbh_engine;

/* one output interval handed to the caller (the concoutput slot,
 * src/timemanager.f90:376-436); arrays are in the reference layout and are
 * only valid during the call.  The fluxoutput / partoutput_average slots of the same place (:439,:455)
 * are the callback's too: with cfg->iflux = 1 / cfg->ipout = 3 the engine's step has accumulated the
 * fluxes / averages (the loop's two hooks, :614-623), and the callback fetches and clears them with
 * fpb_fetch_fluxes(h, flux, 1) / fpb_fetch_partpos_average(h, numpart, av, 1). */
typedef int (*fpbh_output_fn)(void *user, int32_t itime, float outnum, const float *gridunc,
                              const float *griduncn, const float *drygridunc,
                              const float *drygriduncn, const float *creceptor);

typedef struct fpbh_run {
  int32_t ideltas;                        /* signed run length (s) */
  int32_t loutstep, loutaver, loutsample; /* signed like readcommand leaves them */
  int32_t met_interval;                   /* s between synthetic wind fields */
  int32_t met_homogeneous;                /* 1: set_fields_synthetic-style met */
  float met_u, met_v, met_w;
  int32_t max_steps;                      /* >0: stop after that many syncs */
  int32_t met_raw;                        /* 1: synthetic model-level fields (fpbh_synth_rawmet) through the
                                           * engine's calcpar_verttransform; fpb_config::height must then be
                                           * fpbh_verttransform_heights of the field at time 0 */
  int32_t lconvection;                    /* 1: convmix every step (COMMAND LCONVECTION), needs met_raw */
} fpbh_run;

typedef struct fpbh_run_result {
  int64_t particle_steps; /* sum over syncs of particles with itra1 == itime */
  int64_t substeps;
  int32_t syncs, outputs, numpart_final;
  double t_step_s, t_conc_s; /* host wall time inside engine->step / conccalc */
  int32_t convmix_calls, convecting_columns; /* lconvection: calls made, columns that convected (summed) */
  int32_t boundary_particles, split_calls;   /* domain filling: particles created at the boundaries */
} fpbh_run_result;

/* timemanager(metdata_format): src/timemanager.f90:152-729 with the engine
 * in place of the particle loop and conccalc; met comes from fpbh_synth_met, or -- fpbh_run::met_raw --
 * as model-level fields from fpbh_synth_rawmet that the engine transforms itself.  Domain-filling
 * runs (init_domainfill / boundcond_domainfill), convective mixing (forward: after the release,
 * backward: before the new fields, :183-193,258-263) and particle splitting (:472-503) are part of
 * the loop when the configuration asks for them and the engine table has the entry points. */
int fpbh_timemanager(const fpb_config *cfg, const float *height, const fpbh_releases *rel,
                     const fpbh_run *run, const fpbh_engine *eng, fpbh_output_fn out, void *user,
                     fpbh_run_result *result);

#ifdef __cplusplus
}
#endif
#endif /* FPB_HOST_H */
