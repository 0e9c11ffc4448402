/*
 * fpo.h -- CPU ORACLE for the FLEXPART per-particle timestep hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may build, load or call it.
 *
 * It is a sequential, module-global-style restatement in plain C of the
 * reference's Fortran algorithm (MeteoSwiss/flexpart, src/ *.f90); every
 * function cites the file:line it follows.  It shares only the public type
 * definitions of include/fpb.h (fpb_config, fpb_met_ptrs, fpb_particle_ptrs)
 * so tests can hand one problem description to both sides.
 *
 * PINNED AGAINST THE REFERENCE'S OWN SOURCES.  The reference ships no golden
 * vectors for this path and is Fortran, with no Fortran compiler in the build
 * image; oracle/f2c/f90toc.py therefore transpiles the reference's hot-path
 * files (read where they lie under /root/reference/src, nothing copied) to C,
 * oracle/f2c/Makefile compiles them into oracle/_ref/libflexref.so, and
 * tests/test_ref_transpiled.py requires this oracle (strict_reference mode) to be
 * bit-identical to that code call by call: random_mod, initialize, advance with
 * every interpol / hanna / cbl / cmapf / settling branch, nested input grids,
 * conccalc, drydepokernel(_nest), wetdepo.  What that cannot pin is the
 * transcendental library (both sides call fpo_math.h) and gfortran's own code
 * generation; see DESIGN.md section 2.  Analytic known-answer cases, invariants
 * and the Numerical-Recipes generator recurrences pin it independently (tests/).
 *
 * Arithmetic: default `real` = float, `real(dp)` = double, no contraction
 * (build with -O2 -ffp-contract=off, mirroring src/makefile_meteoswiss:103-111).
 * Transcendentals go through fpo_math.h: by default correctly rounded
 * (double evaluation, rounded once to float); -DFPO_LIBM_FLOAT selects
 * glibc's float routines, which is what gfortran would link.
 */
#ifndef FPO_H
#define FPO_H

#include <stddef.h>
#include <stdint.h>

#include "../include/fpb.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct fpo_state {
  fpb_config c;   /* by-value copy; c.height/npart/xmass repointed to own copies */
  float *height;  /* 1-based: height[1..nz] */
  int32_t *npart; /* 1-based */
  float *xmass;   /* xmass[(i-1) + numpoint*(k-1)] */

  /* meteorology: Fortran slot 1..2 */
  fpb_met_ptrs met[3];
  /* nested input grids: metn[l][slot], l = 1..numbnests (uun.. of src/com_mod.f90:501-529) */
  fpb_met_ptrs metn[FPB_MAXNESTS + 1][3];
  int memind[3];
  int memtime[3];
  int lwindinterv;

  /* rannumb(maxrand), 1-based */
  float *rannumb;
  int maxrand;

  /* random_mod SAVEd state, src/random_mod.f90 */
  int r3_inext, r3_inextp, r3_ma[56], r3_iff;
  int r1_iv[33], r1_iy;
  int gd_iset;
  float gd_gset;
  /* SAVEd `idummy` locals: src/advance.f90:120, src/initialize.f90:64,
   * src/releaseparticles.f90 (idummy=-7) */
  int idummy_advance, idummy_initialize, idummy_release;
  int idummy_domainfill;  /* src/init_domainfill.f90:47 (idummy = -11) */
  int numparticlecount;   /* src/com_mod.f90:676 */
  /* inflow boundary of a limited domain-filling box (src/com_mod.f90:244-253), set by fpo_init_domainfill */
  struct {
    int nx_we[2], ny_sn[2], gdomainfill;
    float xmassperparticle;
    int32_t *numcolumn_we, *numcolumn_sn;        /* (2, 0:nymax-1) / (2, 0:nxmax-1) */
    float *zcolumn_we, *zcolumn_sn;              /* (2, 0:n-1, 0:maxcolumn+1): third index 0 holds 0 */
    float *acc_mass_we, *acc_mass_sn;
    int idummy;                                  /* src/boundcond_domainfill.f90:49 */
  } bc;
  float *zpoint1, *zpoint2; /* 1-based release heights (point_mod), backward wet scavenging only */
  float settling_saved; /* src/advance.f90:121 */

  /* particles, 1-based arrays of maxpart+1 */
  int maxpart, numpart;
  double *xtra1, *ytra1;
  float *ztra1;
  int32_t *itra1, *npoint, *nclass, *idt, *itramem, *itrasplit;
  float *uap, *ucp, *uzp, *us, *vs, *ws;
  int16_t *cbt;
  float *xmass1;      /* xmass1[j + (maxpart+1)*(k-1)] */
  float *xscav_frac1; /* same layout */

  /* interpol_mod, src/interpol_mod.f90 */
  float *uprof, *vprof, *wprof, *usigprof, *vsigprof, *wsigprof, *rhoprof,
      *rhogradprof; /* 1-based [nzmax+1] */
  float u, v, w, usig, vsig, wsig;
  float p1, p2, p3, p4, ddx, ddy, rddx, rddy, dtt, dt1, dt2;
  int ix, jy, ixp, jyp, ngrid, indz, indzp;
  int depoindicator[FPB_MAXSPEC + 1];
  unsigned char *indzindicator; /* 1-based */

  /* hanna_mod, src/hanna_mod.f90 */
  float ust, wst, ol, h, zeta, sigu, sigv, tlu, tlv, tlw, sigw, dsigwdz,
      dsigw2dz;

  /* grids (reference layout, maxspec species extent) */
  float *gridunc, *griduncn, *drygridunc, *drygriduncn, *creceptor;
  float *wetgridunc, *wetgriduncn; /* same layout as drygridunc(n) */

  /* behaviour switch ("cross-particle stale state"):
   * 1 = reproduce the reference's leaks between calls,
   * 0 = "defined" behaviour the device implements. */
  int strict_reference;

  /* validation hooks */
  long n_ran3_draws;
  /* when set, the uniform behind `nrand=int(ran3(idummy)*real(maxrand-1))+1` (src/advance.f90:153,
   * src/initialize.f90:68) is popped from this queue instead of ran3: lets a test hand the oracle the
   * engine's counter-based (Philox) index stream, so that the production RNG mode can be compared */
  float *index_queue;
  long index_queue_n, index_queue_pos;

  /* counters */
  long nan_count, nan_count2;
  fpb_step_stats last;
  /* per-particle trace of the last step (tests): branch + substeps */
  int32_t *trace_nsub; /* 1-based, may be NULL */
} fpo_state;

/* lifecycle */
fpo_state *fpo_create(const fpb_config *cfg, int strict_reference);
void fpo_destroy(fpo_state *S);

/* random_mod */
float fpo_ran1(fpo_state *S, int *idum);
float fpo_ran3(fpo_state *S, int *idum);
float fpo_gasdev(fpo_state *S, int *idum);
void fpo_gasdev1(fpo_state *S, int *idum, float *r1, float *r2);
void fpo_fill_rannumb(fpo_state *S, int maxrand, int idummy);
void fpo_set_rannumb(fpo_state *S, const float *tab, int n);
const float *fpo_rannumb(fpo_state *S); /* 0-based view of the table */
float fpo_index_uniform(fpo_state *S, int *idum);
void fpo_set_index_uniforms(fpo_state *S, const float *u, long n);

/* meteorology */
void fpo_set_met(fpo_state *S, int slot, const fpb_met_ptrs *m);
void fpo_set_met_nest(fpo_state *S, int slot, int nest, const fpb_met_ptrs *m);
void fpo_set_met_bracket(fpo_state *S, const int memind[2],
                         const int memtime[2], int lwindinterv);

/* particles */
void fpo_push_particles(fpo_state *S, int first, int count,
                        const fpb_particle_ptrs *p);
void fpo_pull_particles(fpo_state *S, int first, int count,
                        const fpb_particle_ptrs *p);
void fpo_set_numpart(fpo_state *S, int numpart);
int fpo_numpart(const fpo_state *S);
const int32_t *fpo_trace_nsub(const fpo_state *S);

/* hot path */
void fpo_initialize(fpo_state *S, int itime, int32_t *ldt, float *up, float *vp,
                    float *wp, float *usigold, float *vsigold, float *wsigold,
                    double xt, double yt, float zt, int16_t *icbt);
void fpo_advance(fpo_state *S, int itime, int nrelpoint, int32_t *ldt,
                 float *up, float *vp, float *wp, float *usigold,
                 float *vsigold, float *wsigold, int *nstop, double *xt,
                 double *yt, float *zt, float *prob, int16_t *icbt);
void fpo_step(fpo_state *S, int itime, int ldeltat, fpb_step_stats *stats);
void fpo_conccalc(fpo_state *S, int itime, float weight);
void fpo_drydepokernel(fpo_state *S, int nunc, const float *deposit, float x,
                       float y, int nage, int kp);
void fpo_drydepokernel_nest(fpo_state *S, int nunc, const float *deposit,
                            float x, float y, int nage, int kp);
void fpo_fetch_grids(fpo_state *S, float *gridunc, float *griduncn,
                     float *drygridunc, float *drygriduncn, float *creceptor,
                     int zero_conc);
void fpo_scale_depgrids(fpo_state *S, const float *factor);
/* wetdepo(itime,ltsample,loutnext) with ldeltat precomputed (src/wetdepo.f90:55-63) */
void fpo_wetdepo(fpo_state *S, int itime, int ltsample, int ldeltat);
void fpo_fetch_wetgrids(fpo_state *S, float *wetgridunc, float *wetgriduncn);

/* fpo_output.c: outgrid_init's cell geometry and concoutput's sparse dump of one (ks, kp, nage) */
void fpo_outgrid_geometry(const fpb_config *c, int nest, float outlat0, float *area, float *volume);
void fpo_partoutput_record(const fpb_config *c, const float *height, int itime, const int32_t memtime[2],
                           double xtra1, double ytra1, float ztra1, const float *oro,
                           const float *pv[2], const float *qv[2], const float *tt[2], const float *rho[2],
                           const float *hmix[2], const float *tropopause[2], float out[9]);

/* the optional hooks of the particle loop (fpo_hooks.c): src/calcfluxes.f90, src/partpos_average.f90,
 * src/initial_cond_calc.f90 */
void fpo_calcfluxes(const fpb_config *c, float *flux, int nage, int npoint, float xold, float yold, float zold,
                    double xtra1, double ytra1, float ztra1, const float *mass);
void fpo_partpos_average(const fpb_config *c, const float *height, int itime, const int32_t memtime[2], double xtra1,
                         double ytra1, float ztra1, const float *oro, const float *pv[2], const float *qv[2],
                         const float *tt[2], const float *uu[2], const float *vv[2], const float *rho[2],
                         const float *hmix[2], const float *tropopause[2], float out[14]);
void fpo_initial_cond_calc(const fpb_config *c, const float *height, float *init_cond, int linit_cond, double xtra1,
                           double ytra1, float ztra1, int npoint, const float *rho2, const float *mass);
void fpo_density_outgrid(const fpb_config *c, const float *height, int nest, float outlon0, float outlat0,
                         const float *rho, float *densityoutgrid);
void fpo_concoutput_sparse(const fpb_config *c, int nest, int which, const float *grid_ref,
                           const float *geom, const float *density, int ks, int kp, int nage, float outnum, float tot_mu,
                           int loutaver, int32_t *sp_count_i, int32_t *sparse_dump_i,
                           int32_t *sp_count_r, float *sparse_dump_r);

/* pieces exported for unit tests */
void fpo_hanna(fpo_state *S, float z);
void fpo_hanna1(fpo_state *S, float z);
void fpo_hanna_short(fpo_state *S, float z);
void fpo_windalign(float u, float v, float ffap, float ffcp, float *ux,
                   float *vy);
/* the *_nests twins (src/interpol_all_nests.f90 etc.) are the same functions:
 * with S->ngrid > 0 they read the nest's arrays and xt,yt are nest coordinates */
void fpo_interpol_all(fpo_state *S, int itime, float xt, float yt, float zt);
void fpo_interpol_misslev(fpo_state *S, int n);
void fpo_interpol_wind(fpo_state *S, int itime, float xt, float yt, float zt);
void fpo_interpol_wind_short(fpo_state *S, int itime, float xt, float yt,
                             float zt);
void fpo_interpol_vdep(fpo_state *S, int level, float *vdepo);
void fpo_get_settling(fpo_state *S, int itime, float xt, float yt, float zt,
                      int nsp, float *settling);
void fpo_cbl(fpo_state *S, float wp, float zp, float ust, float wst, float h,
             float rhoa, float rhograd, float sigmaw, float dsigmawdz,
             float tlw, float *ptot, float *Q, float *phi, float *ath,
             float *bth, float ol, int *flagrein);
void fpo_re_initialize_particle(fpo_state *S, float zp, float ust, float wst,
                                float h, float sigmaw, float *wp, int *nrand,
                                float ol);
void fpo_initialize_cbl_vel(fpo_state *S, int *idum, float zp, float ust,
                            float wst, float h, float sigmaw, float *wp,
                            float ol);

void fpo_initialize_cbl_vel_defined(fpo_state *S, float dcas, float dcas1,
                                    float zp, float wst, float h, float sigmaw,
                                    float *wp, float ol);

/* cmapf_mod */
void fpo_stlmbr(float *strcmp, float tnglat, float xlong);
void fpo_stcm2p(float *strcmp, float x1, float y1, float xlat1, float xlong1,
                float x2, float y2, float xlat2, float xlong2);
void fpo_cll2xy(const float *strcmp, float xlat, float xlong, float *x,
                float *y);
void fpo_cxy2ll(const float *strcmp, float x, float y, float *xlat,
                float *xlong);
float fpo_cgszll(const float *strcmp, float xlat, float xlong);
void fpo_cc2gll(const float *strcmp, float xlat, float xlong, float ue,
                float vn, float *ug, float *vg);

/* backward-run receptor scavenging of the particle loop (src/timemanager.f90:571-598) */
void fpo_set_release_heights(fpo_state *S, const float *zpoint1, const float *zpoint2, int numpoint);
void fpo_get_vdep_prob(fpo_state *S, int itime, double xt, double yt, float zt, float *prob);
float fpo_get_wetscav(fpo_state *S, int itime, int ltsample, int jpart, int ks, float *grfraction1);
void fpo_interpol_weights(fpo_state *S, int itime, float xt, float yt);

/* init_domainfill (fpo_domainfill.c) */
void fpo_domainfill_gridarea(const fpb_config *c, const int ny_sn[2], float *gridarea);
int fpo_init_domainfill(fpo_state *S, float xpoint1, float ypoint1, float xpoint2, float ypoint2,
                        int itsplit, int32_t *out, float *fout);
/* boundcond_domainfill(itime): returns the number of particles created, -1 when maxpart is exceeded */
int fpo_boundcond_domainfill(fpo_state *S, int itime, int itsplit);
/* total number of boundary release locations and the sum of the accumulated masses (diagnostics) */
int fpo_boundcond_locations(fpo_state *S, double *accmass_sum);
int fpo_numparticlecount(const fpo_state *S);

/* releaseparticles (integer semantics + ran1 stream) */
void fpo_split_particles(fpo_state *S, int itime);
int fpo_releaseparticles(fpo_state *S, int itime, int numpoint,
                         const int32_t *ireleasestart,
                         const int32_t *ireleaseend, const float *xpoint1,
                         const float *ypoint1, const float *xpoint2,
                         const float *ypoint2, const float *zpoint1,
                         const float *zpoint2, float *xmasssave, int itsplit);

#ifdef __cplusplus
}
#endif
#endif
