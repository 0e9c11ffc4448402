/*
 * fpb.h -- C ABI of the B200 particle-timestep engine (libfpb.so).
 *
 * This is the drop-in boundary for FLEXPART's per-particle hot path.  The
 * reference (MeteoSwiss/flexpart, Fortran 90) has no FFI of its own; the seam
 * is timemanager's particle loop and its conccalc calls.  Every entry point
 * below names the reference lines it replaces (paths relative to the
 * reference tree).  A Fortran ISO_C_BINDING interface module for these
 * symbols is given as text in INTEGRATION.md.
 *
 * Conventions
 *   - plain C types only; every call returns 0 on success, non-zero on error,
 *     and fpb_last_error() then holds a message (the reference convention is
 *     banner + `stop 1`, e.g. src/timemanager.f90:205-208 -- the Fortran shim
 *     does `if (ierr /= 0) stop 1`).
 *   - all calls are synchronous on return.
 *   - host arrays are Fortran column-major with the reference's padded extents
 *     and stay owned by the caller; the library owns all device memory.
 *   - particle slot indices and array offsets are 0-based here (Fortran slot
 *     j is offset j-1); model times are FLEXPART's integer seconds.
 *   - there is no CPU fallback: every compute entry point needs a CUDA device.
 */
#ifndef FPB_H
#define FPB_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FPB_ABI_VERSION 3
#define FPB_MAXSPEC 8      /* >= par_mod maxspec (5), src/par_mod.f90:210 */
#define FPB_MAXAGECLASS 8  /* >= par_mod maxageclass */
#define FPB_MAXZGRID 64    /* output-grid levels */
#define FPB_MAXRECEPTOR 20 /* par_mod maxreceptor, src/par_mod.f90:204 */
#define FPB_NSLOTS 3       /* met time levels on the device: memind slots 1, 2 + a read-ahead slot
                              (numwfmem = 3 of the reference's MPI build, src/par_mod.f90:226-227) */
#define FPB_MAXNESTS 3     /* >= par_mod maxnests (0 shipped, 1 MeteoSwiss:
                              src/par_mod.f90:152, src/par_mod_meteoswiss.f90:154) */

/* Values of itra1 for a terminated particle, src/timemanager.f90:632 */
#define FPB_ITRA_DEAD (-999999999)

/* fpb_config.rng_mode */
enum {
  FPB_RNG_REFERENCE = 0, /* ran3-indexed rannumb table, reference draw order
                            (validation; src/advance.f90:153) */
  FPB_RNG_PHILOX_INDEX = 1, /* Philox4x32-10 replaces ran3 as the table index */
  FPB_RNG_PHILOX = 2     /* Philox4x32-10 + Box-Muller clipped to +-3, no table */
};

/* fpb_config.math_mode */
enum {
  FPB_MATH_FAST = 0,  /* FMA contraction on, float intrinsics */
  FPB_MATH_STRICT = 1 /* no contraction, correctly rounded transcendentals:
                         bit-comparable with the CPU oracle */
};

/* fpb_config.scatter_mode (conccalc / drydepokernel accumulation) */
enum {
  FPB_SCATTER_ATOMIC = 0,       /* red.global.add.f32 */
  FPB_SCATTER_DETERMINISTIC = 1 /* sort by cell + ordered segmented sum: the concentration grids, the dry and wet
                                 * deposition grids and the receptor sums are bit-identical to the serial loops */
};

/*
 * Run constants: the subset of par_mod / com_mod / outg_mod / unc_mod /
 * point_mod the hot path reads.  Copied by value in fpb_init.
 */
typedef struct fpb_config {
  int32_t abi_version; /* FPB_ABI_VERSION */

  /* --- meteorological grid, src/gridcheck_ecmwf.f90:300-366 ------------- */
  int32_t nx, ny, nz;          /* used extents; nz = nuvz = levels+1 */
  int32_t nxmax, nymax, nzmax; /* padded host extents, src/par_mod.f90:142-154 */
  int32_t nxmin1, nymin1;
  float dx, dy, xlon0, ylat0;
  float dxconst, dyconst;      /* 180/(dx*r_earth*pi), src/gridcheck_ecmwf.f90 */
  int32_t xglobal, nglobal, sglobal;
  float switchnorthg, switchsouthg;
  float northpolemap[9], southpolemap[9]; /* cmapf_mod stlmbr/stcm2p result */
  float eps;                   /* nxmax/3.e5, src/advance.f90:107 */

  /* --- nested meteorological input grids, src/gridcheck_nests.f90:340-385 --
   * Nest l (1-based in the reference, index l-1 here) covers
   * xln < x < xrn, yln < y < yrn in mother-grid units; a particle inside it
   * (highest l first, src/advance.f90:166-173) is interpolated from the nest's
   * arrays at xtn = (xt-xln)*xresoln, ytn = (yt-yln)*yresoln. */
  int32_t numbnests;
  int32_t nxn[FPB_MAXNESTS], nyn[FPB_MAXNESTS]; /* used extents */
  int32_t nxmaxn, nymaxn;                       /* padded host extents, src/par_mod.f90:152 */
  float xln[FPB_MAXNESTS], yln[FPB_MAXNESTS], xrn[FPB_MAXNESTS], yrn[FPB_MAXNESTS];
  float xresoln[FPB_MAXNESTS], yresoln[FPB_MAXNESTS]; /* dx/dxn, dy/dyn */

  /* --- COMMAND, src/readcommand.f90:244-272,377-383,622-634 ------------- */
  int32_t ldirect;    /* +1 / -1 */
  int32_t lsynctime;  /* negative for backward runs */
  int32_t method, mintime, ifine;
  int32_t turbswitch, cblflag, mdomainfill, mquasilag, lsettling;
  float ctl;          /* already inverted: 1/CTL */
  float fine;         /* 1/ifine */
  float d_trop, d_strat, turbmesoscale; /* src/par_mod.f90:79 */
  int32_t turboff;    /* com_mod turboff, src/com_mod.f90:778 */
  int32_t ind_samp;   /* 0 or -1 */
  int32_t ioutputforeachrelease, lusekerneloutput, lparticlecountoutput;
  int32_t drydep, drybkdep, wetbkdep, nested_output;

  /* --- species, src/readspecies.f90 / readreleases.f90:349-386 ---------- */
  int32_t nspec;
  float decay[FPB_MAXSPEC];
  int32_t drydepspec[FPB_MAXSPEC];
  float density[FPB_MAXSPEC], dquer[FPB_MAXSPEC], vsetaver[FPB_MAXSPEC],
      cunningham[FPB_MAXSPEC];

  /* --- wet deposition, src/readspecies.f90 + src/readreleases.f90:349-370 -- */
  int32_t wetdep;                     /* WETDEP */
  int32_t wetdepspec[FPB_MAXSPEC];    /* WETDEPSPEC */
  float weta_gas[FPB_MAXSPEC], wetb_gas[FPB_MAXSPEC];     /* below-cloud, gases */
  float crain_aero[FPB_MAXSPEC], csnow_aero[FPB_MAXSPEC]; /* below-cloud, aerosols */
  float ccn_aero[FPB_MAXSPEC], in_aero[FPB_MAXSPEC];      /* in-cloud, aerosols */
  float henry[FPB_MAXSPEC];                               /* in-cloud, gases */
  int32_t readclouds;                 /* ctwc holds cloud water read from the met input */
  int32_t readclouds_nest[FPB_MAXNESTS];

  /* --- age classes, src/readageclasses.f90 ------------------------------ */
  int32_t nageclass;
  int32_t lage[FPB_MAXAGECLASS];

  /* --- output grids, src/readoutgrid.f90:199-200, outgrid_init.f90:192 -- */
  int32_t numxgrid, numygrid, numzgrid;
  float dxout, dyout, xoutshift, youtshift;
  float outheight[FPB_MAXZGRID];
  int32_t numxgridn, numygridn;
  float dxoutn, dyoutn, xoutshiftn, youtshiftn;
  int32_t maxpointspec_act, nclassunc, maxageclass;
  int32_t maxspec; /* species extent of the host grid arrays (reference: 5) */

  /* --- receptors, src/readreceptors.f90 --------------------------------- */
  int32_t numreceptor;
  float xreceptor[FPB_MAXRECEPTOR], yreceptor[FPB_MAXRECEPTOR],
      receptorarea[FPB_MAXRECEPTOR];

  /* --- releases, src/readreleases.f90 ----------------------------------- */
  int32_t numpoint;
  const int32_t *npart; /* [numpoint] */
  const float *xmass;   /* (numpoint, maxspec) column-major: xmass[i + numpoint*k] */

  /* --- vertical levels, src/verttransform_ecmwf.f90:153-168 ------------- */
  const float *height; /* [nz], height[0] = 0 */

  /* --- engine ----------------------------------------------------------- */
  int32_t maxpart;      /* particle capacity on this device */
  int32_t device;       /* CUDA device ordinal */
  int32_t rng_mode, math_mode, scatter_mode;
  uint64_t seed;        /* Philox key */
  int32_t part_id_stride, part_id_offset; /* global id = offset + stride*slot:
                           multi-GPU partition, src/releaseparticles_mpi.f90:141 */
  int32_t sort_interval; /* >0: fpb_step re-orders the device rows by met cell
                            every that many steps (results do not depend on it) */
  /* --- optional hooks of the particle loop, src/timemanager.f90:614-623 -- */
  int32_t iflux;         /* 1: calcfluxes after every advance (src/calcfluxes.f90); fpb_fetch_fluxes */
  int32_t ipout;         /* 3: partpos_average after every advance (src/partpos_average.f90);
                            fpb_fetch_partpos_average.  Other values: nothing happens in the loop */
  int32_t linit_cond;    /* 1 (mass units) / 2 (mass mixing ratio): initial_cond_calc for every particle the loop
                            terminates (src/initial_cond_calc.f90; backward runs); fpb_fetch_init_cond */
  int32_t reserved[4];
} fpb_config;

/* One time level of the meteorological arrays the hot path gathers from
 * (src/com_mod.f90:355-371,410-427,451).  3-D: (nxmax,nymax,nzmax); 2-D:
 * (nxmax,nymax); vdep: (nxmax,nymax,maxspec).  uupol/vvpol may be NULL when
 * neither pole is in the domain; tt may be NULL when lsettling == 0; vdep may
 * be NULL when drydep == 0.  For a nested grid (fpb_upload_met_nest) the
 * extents are (nxmaxn,nymaxn,nzmax) / (nxmaxn,nymaxn) / (nxmaxn,nymaxn,maxspec)
 * (src/com_mod.f90:501-529) and uupol/vvpol/tt are not read. */
typedef struct fpb_met_ptrs {
  const float *uu, *vv, *ww, *rho, *drhodz, *tt, *uupol, *vvpol;
  const float *hmix, *ustar, *wstar, *oli, *tropopause;
  const float *vdep;
  /* wet deposition only (wetdep != 0; may be NULL otherwise): lsprec, convprec,
   * tcc (nxmax,nymax) [mm/h, mm/h, 0..1], ctwc (nxmax,nymax) total cloud water
   * (read when readclouds), clouds (nxmax,nymax,nzmax) integer(kind=1) cloud /
   * precipitation class of verttransform_ecmwf (src/com_mod.f90:376-395);
   * tt is then required as well */
  const float *lsprec, *convprec, *tcc, *ctwc;
  const int8_t *clouds;
} fpb_met_ptrs;

/* Particle arrays (src/com_mod.f90:675-695).  xmass1 / xscav_frac1 are
 * (ld, nspec) column-major.  Any pointer may be NULL in fpb_pull_particles to
 * skip that array; xscav_frac1 may be NULL unless drybkdep/wetbkdep. */
typedef struct fpb_particle_ptrs {
  double *xtra1, *ytra1;
  float *ztra1;
  int32_t *itra1, *npoint, *nclass, *idt, *itramem, *itrasplit;
  float *uap, *ucp, *uzp, *us, *vs, *ws;
  int16_t *cbt;
  float *xmass1;
  float *xscav_frac1;
  int32_t ld;
} fpb_particle_ptrs;

typedef struct fpb_step_stats {
  int64_t n_active;     /* particles with itra1 == itime on entry */
  int64_t n_init;       /* of those, initialize() calls */
  int64_t n_terminated; /* set to FPB_ITRA_DEAD in this step */
  int64_t n_pbl;        /* took the PBL branch (zeta <= 1) on entry */
  int64_t n_substeps;   /* label-100 loop iterations, summed */
  int64_t n_petterssen; /* Petterssen corrector applied */
  int64_t n_nan_cbl;    /* nan_count + nan_count2, src/advance.f90:421,439 */
  int64_t n_nonfinite;  /* of n_terminated: position not finite after advance (the reference
                           carries such a particle on and indexes out of bounds with it) */
} fpb_step_stats;

typedef struct fpb_handle fpb_handle;

const char *fpb_last_error(void);
int fpb_abi_version(void);
size_t fpb_config_sizeof(void);

/* after outgrid_init, before timemanager: src/FLEXPART.f90:342-453 */
int fpb_init(const fpb_config *cfg, fpb_handle **out);
/* src/timemanager.f90:760-774 */
int fpb_finalize(fpb_handle *h);

/* rannumb(maxrand) filled at src/FLEXPART.f90:56-59 */
int fpb_set_rannumb(fpb_handle *h, const float *rannumb, int32_t n);
/* Fill the table inside the library with the same ran3/gasdev1 recipe
 * (src/random_mod.f90:70-139, src/FLEXPART.f90:47,56-59). */
int fpb_fill_rannumb(fpb_handle *h, int32_t maxrand, int32_t idummy);

/* after getfields returned a new field: src/timemanager.f90:200,
 * src/getfields.f90:109-139.  slot is the Fortran slot index 1 or 2 (3: read-ahead, below). */
int fpb_upload_met(fpb_handle *h, int32_t slot, const fpb_met_ptrs *met);
/* The same in two halves, for reading ahead: _begin enqueues the copies and the re-packing of one
 * time level on the engine's upload stream and returns (with page-locked source arrays -- see
 * fpb_host_register -- at once; the arrays must stay untouched until _end), the steps of the
 * current bracket go on meanwhile; _end waits and reports the device time of the upload [ms].
 * A third slot exists for this (FPB_NSLOTS = 3; the numwfmem = 3 of the reference's MPI build with
 * a dedicated reader process, src/par_mod.f90:226-227, src/getfields_mpi.f90): the field after next
 * goes to the slot that is not in the bracket, and fpb_set_met_bracket(memind) then names any two
 * of the three.  fpb_set_met_bracket waits for an upload in flight. */
int fpb_upload_met_begin(fpb_handle *h, int32_t slot, const fpb_met_ptrs *met);
int fpb_upload_met_end(fpb_handle *h, float *upload_ms /* may be NULL */);
/* cudaHostRegister / cudaHostUnregister for a host that does not link the CUDA runtime: page-locks
 * the met arrays (asynchronous uploads) or the particle arrays (fpb_step_host) */
int fpb_host_register(void *p, size_t bytes);
int fpb_host_unregister(void *p);
/* the same for nested input grid `nest` (1..numbnests): the slices
 * uun(:,:,:,slot,nest) .. of src/com_mod.f90:501-529, filled by
 * readwind_nests / verttransform_nests / calcpar_nests (src/getfields.f90:141-170) */
int fpb_upload_met_nest(fpb_handle *h, int32_t slot, int32_t nest, const fpb_met_ptrs *met);
/* memind(1:2), memtime(1:2), lwindinterv: src/getfields.f90:96-176 */
int fpb_set_met_bracket(fpb_handle *h, const int32_t memind[2],
                        const int32_t memtime[2], int32_t lwindinterv);

/* after releaseparticles / init_domainfill / boundcond_domainfill / split:
 * src/timemanager.f90:230-251,473-504.  Copies rows [first, first+count). */
int fpb_push_particles(fpb_handle *h, int32_t first, int32_t count,
                       const fpb_particle_ptrs *p);
int fpb_set_numpart(fpb_handle *h, int32_t numpart);
/* before partoutput / plumetraj / wetdepo / convmix: src/timemanager.f90:453 */
int fpb_pull_particles(fpb_handle *h, int32_t first, int32_t count,
                       const fpb_particle_ptrs *p);

/* replaces the particle loop src/timemanager.f90:531-712 (initialize,
 * advance, decay, dry-deposition split + drydepokernel, termination). */
int fpb_step(fpb_handle *h, int32_t itime, int32_t ldeltat,
             fpb_step_stats *stats /* may be NULL */);

/* replaces conccalc(itime,weight): src/timemanager.f90:364,463 */
int fpb_conccalc(fpb_handle *h, int32_t itime, float weight);

/* One whole synchronisation interval for a host that keeps the particle
 * arrays (wetdepo / convmix / partoutput still on the CPU between steps):
 * rows [0,numpart) of the caller's arrays go in, conccalc(itime,conc_weight)
 * (skipped when conc_weight <= 0: outside the sampling window,
 * src/timemanager.f90:340-364) and the particle loop src/timemanager.f90:531-712
 * run, and the arrays the loop writes come back (xtra1, ytra1, ztra1, itra1,
 * idt, uap, ucp, uzp, us, vs, ws, cbt, xmass1).  Equivalent to
 * fpb_push_particles + fpb_conccalc + fpb_step + fpb_pull_particles, but cut
 * into row chunks that run on separate streams so the host<->device copies of
 * one chunk overlap the kernels of the others; give it page-locked arrays.
 * Synchronous on return.  With FPB_SCATTER_DETERMINISTIC the chunks add to the grids in slot order
 * (bit-identical to the resident path and to the serial loop).  Only the arrays the loop
 * reads are uploaded (not itrasplit; xscav_frac1 only in backward deposition runs): the device copies
 * of those two are undefined afterwards, push before pulling them. */
int fpb_step_host(fpb_handle *h, int32_t itime, int32_t ldeltat, int32_t numpart,
                  const fpb_particle_ptrs *p, float conc_weight,
                  fpb_step_stats *stats /* may be NULL */);

/* replaces wetdepo(itime,lsynctime,loutnext): src/timemanager.f90:164-169,
 * src/wetdepo.f90:70-147 with get_wetscav, interpol_rain(_nests) and
 * wetdepokernel(_nest).  ltsample = lsynctime; ldeltat as computed at
 * src/wetdepo.f90:55-63.  Particles with itra1 <= itime (>= for backward
 * runs) lose mass to wetgridunc/wetgriduncn.   */
int fpb_wetdepo(fpb_handle *h, int32_t itime, int32_t ltsample, int32_t ldeltat);

/* Output-grid geometry of outgrid_init (src/outgrid_init.f90:48-100): area(numxgrid,numygrid) and
 * volume(numxgrid,numygrid,numzgrid); the nested grid's (src/outgrid_init_nest.f90) may be NULL. */
int fpb_set_outgrid_geometry(fpb_handle *h, const float *area, const float *volume,
                             const float *arean, const float *volumen);
/* lower left corners outlon0, outlat0 (and outlon0n, outlat0n) as read by readoutgrid(_nest): only
 * needed for the mixing-ratio record (which = 3), whose densityoutgrid is looked up from them */
int fpb_set_outgrid_origin(fpb_handle *h, float outlon0, float outlat0, float outlon0n, float outlat0n);

/* The body of concoutput's loop over (ks, kp, nage) for the sparse binary output (iout = 1;
 * src/concoutput.f90:287-475, concoutput_nest.f90 alike): mean over the uncertainty classes x
 * nclassunc, conversion (1.e12/volume/outnum/tot_mu forward, abs(loutaver)/outnum/tot_mu backward;
 * 1.e12/area for deposition) and the run-length dump -- the index of the first cell of every run of
 * cells above tiny(0.0) and the values with the sign alternating from run to run -- built on the
 * device, so only the compacted lists cross the bus.  which: 0 concentration, 1 dry deposition,
 * 2 wet deposition, 3 mixing ratio (iout = 2, 3; :563-593: 1.e12/volume/outnum*weightair/
 * weightmolar(ks)/densityoutgrid with the air density of time level memind(2), :164-190 -- pass
 * weightmolar(ks) in tot_mu); nest: 0 mother, 1 nested output grid; ks, kp, nage 1-based.  The buffers hold
 * numxgrid*numygrid*numzgrid entries (numxgrid*numygrid for deposition).  The grid totals and
 * their uncertainty (gridtotal, gridsigmatotal: diagnostics on stdout) are not computed.
 *  */
int fpb_concoutput_sparse(fpb_handle *h, int32_t nest, int32_t which, int32_t ks, int32_t kp,
                          int32_t nage, float outnum, float tot_mu, int32_t loutaver,
                          int32_t *sp_count_i, int32_t *sparse_dump_i, int32_t *sp_count_r,
                          float *sparse_dump_r);

/* replaces the particle-splitting block of timemanager (src/timemanager.f90:472-503): every
 * particle (dead ones included, as in the reference) whose itrasplit has been reached is duplicated
 * into the slot numpart + its rank among such particles, both halves with half the mass and
 * itrasplit doubled; candidates that would exceed maxpart stay untouched.  The caller keeps the
 * outer test `ldirect*itime >= ldirect*itsplit`.  Needs fpb_set_releases (for the block-scan scratch)
 * or any earlier fpb_releaseparticles.  *numpart returns the new numpart. */
int fpb_split_particles(fpb_handle *h, int32_t itime, int32_t *numpart);

/* partoutput (src/partoutput.f90:66-192): the particle dump's
 * records -- position in degrees, height, release point and time, and topography, potential
 * vorticity, humidity, density, mixing height, tropopause and temperature interpolated to the
 * particle -- built on the device for the particles with itra1 == itime, in slot order.  Needs
 * the orography (once) and pv, qv of both time levels (after each fpb_upload_met of a slot). */
int fpb_set_orography(fpb_handle *h, const float *oro /* (nxmax,nymax) */);
int fpb_upload_pvqv(fpb_handle *h, int32_t slot, const float *pv, const float *qv /* (nxmax,nymax,nzmax) */);
typedef struct fpb_partout_ptrs { /* each [>= number of active particles]; xmass1 (ld, nspec) column-major */
  int32_t *npoint;
  float *xlon, *ylat, *ztra1;
  int32_t *itramem;
  float *topo, *pvi, *qvi, *rhoi, *hmixi, *tri, *tti;
  float *xmass1;
  int32_t ld;
} fpb_partout_ptrs;
int fpb_partoutput(fpb_handle *h, int32_t itime, int32_t *nrecords, const fpb_partout_ptrs *out);

/* The two optional hooks of the particle loop (src/timemanager.f90:614-623) run inside fpb_step / fpb_step_host
 * when the configuration asks for them:
 *   iflux = 1  calcfluxes (src/calcfluxes.f90): every advanced particle adds its masses (from before the step's
 *              decay / deposition, as the reference's call order has it) to the faces of the output-grid cells
 *              it crossed.  fpb_fetch_fluxes copies flux(6, 0:numxgrid-1, 0:numygrid-1, numzgrid, nspec,
 *              maxpointspec_act, nageclass) out (NULL: no copy) and, with zero != 0, clears it the way
 *              fluxoutput does after writing (src/fluxoutput.f90:288-303).  fpb_convmix adds the fluxes of its
 *              vertical displacements the same way (src/convmix.f90:205-218).
 *   ipout = 3  partpos_average (src/partpos_average.f90): per-particle running sums for partoutput_average;
 *              needs fpb_set_orography and fpb_upload_pvqv like fpb_partoutput.  fpb_fetch_partpos_average
 *              copies the first numpart slots of npart_av / part_av_* (any pointer may be NULL) and, with
 *              zero != 0, clears them (src/partoutput_average.f90:171-187).  A particle the step has just
 *              terminated is left out (the reference averages it at a position that may lie outside the
 *              fields; it is never written).
 * flux and init_cond are accumulated with float atomics in every scatter_mode (the order of the additions is not
 * the particle order; sums of exactly representable masses are exact, others agree to 1e-6 relative). */
typedef struct fpb_partav_ptrs {
  int32_t *npart_av;
  float *cartx, *carty, *cartz, *z, *topo, *pv, *qv, *tt, *uu, *vv, *rho, *tro, *hmix, *energy;
} fpb_partav_ptrs;
int fpb_fetch_fluxes(fpb_handle *h, float *flux, int32_t zero);
/*   linit_cond = 1, 2  initial_cond_calc (src/initial_cond_calc.f90, backward runs): a particle that leaves the
 *              domain (src/timemanager.f90:631; its masses from before the step) or reaches the maximum age
 *              (:702; its masses after decay and deposition) adds xmass1 / rho (1) or xmass1 (2) to the
 *              sensitivity-to-initial-conditions grid through the output kernel.  fpb_initial_cond_final does the
 *              same for every particle still active at the end of the run (:733-737; call it with the final itime
 *              in place of that loop); fpb_fetch_init_cond copies init_cond(0:numxgrid-1, 0:numygrid-1, numzgrid,
 *              maxspec, maxpointspec_act) out and, with zero != 0, clears it.  linit_cond = 1 reads rho of memind(2)
 *              at the particle: for a particle outside the domain the reference reads past its arrays there, the
 *              engine takes the nearest grid point. */
int fpb_initial_cond_final(fpb_handle *h, int32_t itime);
int fpb_fetch_init_cond(fpb_handle *h, float *init_cond, int32_t zero);
int fpb_fetch_partpos_average(fpb_handle *h, int32_t numpart, const fpb_partav_ptrs *out, int32_t zero);

/* Release points (src/point_mod.f90:15-26 after the conversions of src/readreleases.f90 and
 * src/FLEXPART.f90:401-404): coordinates in grid units, heights in metres above ground
 * (zkind 1), times in seconds relative to the start and snapped to lsynctime. */
typedef struct fpb_release_points {
  int32_t numpoint;
  const int32_t *ireleasestart, *ireleaseend;
  const float *xpoint1, *ypoint1, *xpoint2, *ypoint2, *zpoint1, *zpoint2;
  int32_t itsplit;
  int32_t mp_pid; /* MPI rank of the reference's seed offset (src/mpi_mod.f90:331-335); 0 = serial */
} fpb_release_points;

/* copies the release points to the device; resets the release state (xmasssave, ran1 seed) */
int fpb_set_releases(fpb_handle *h, const fpb_release_points *rel);

/* replaces `call releaseparticles(itime)` (src/timemanager.f90:230-233, src/releaseparticles.f90:69-378;
 * EMISVAR factors 1, zkind 1, mass units): the new particles are created on the device, in the slots
 * the reference's search finds (first slots whose itra1 differs from itime), so no particle row
 * crosses the bus.  FPB_RNG_REFERENCE: positions from the reference's ran1 stream, bit-identical;
 * Philox modes: four uniforms per particle from its counter stream.  *numpart returns the new
 * numpart, *n_released the number of particles created.  Fails (label 996 of the reference) when
 * fewer free slots than new particles are left.   */
int fpb_releaseparticles(fpb_handle *h, int32_t itime, int32_t *numpart /* may be NULL */,
                         int32_t *n_released /* may be NULL */);
/* Domain-filling runs (MDOMAINFILL = 1; BASELINE configs[4]).  fpb_init_domainfill replaces `call
 * init_domainfill` (src/timemanager.f90:230-235, src/init_domainfill.f90:55-283): the box of release
 * point 1 (grid units) is filled with npart(1) particles, each column holding a number proportional
 * to its air mass ((p(1)-p(nz))/g*area from rho*r_air*tt of time slot 1, so fpb_upload_met(1, ..)
 * with tt must have happened), at pressure-equidistant heights (random ones in columns of <= 20
 * particles), every particle carrying its share of the column mass in species 1.  The particles are
 * created on the device, in the reference's order (columns south to north, west to east):
 * FPB_RNG_REFERENCE replays the reference's ran1 stream bit for bit, the Philox modes draw from the
 * particle's counter stream.  With part_id_stride = N, part_id_offset = r the call keeps every N-th
 * particle (global index g with g mod N = r, in slot g / N): the round-robin distribution of
 * src/init_domainfill_mpi.f90:86-104 without the root process or any exchange.  Needs an empty
 * particle set (ipin = 0).  For a box that is not the whole globe the second half of the routine
 * (:287-389) memorises the release heights of the four inflow boundaries.  Not built:
 * MDOMAINFILL = 2 (stratospheric ozone, :236-251) and restarts (ipin = 1, boundcond.bin).
 * fpb_boundcond_domainfill replaces `call boundcond_domainfill(itime,loutend)`
 * (src/timemanager.f90:240, src/boundcond_domainfill.f90:54-560): for a global domain (gdomainfill)
 * the reference returns at once and so does this.  For a limited box the particles of the current
 * step that left the box are terminated, the air-mass flux through every boundary release location
 * (wind and density of the met bracket at itime, so fpb_set_met_bracket must have been called) is
 * accumulated, and wherever half a particle mass has accumulated new particles are created in the
 * slots the reference's search finds -- all on the device; per call one int per release location
 * comes back to the host.  FPB_RNG_REFERENCE replays the routine's own ran1 stream (bit-identical),
 * the Philox modes draw from the new particle's counter stream.  With part_id_stride = N every rank
 * accumulates the same fluxes and keeps every N-th new particle.  *numpart returns the new numpart,
 * *n_created the particles this rank created.  A boundary column with exactly two release heights
 * reads element 0 of a 1-based array in the reference (:115); that element is 0 here. */
typedef struct fpb_domainfill_info {
  int32_t nx_we[2], ny_sn[2]; /* domain box in met-grid indices, src/com_mod.f90:245 */
  int32_t gdomainfill;        /* 1: global domain filling, no boundary conditions */
  int32_t numcolumn;          /* largest number of particles in one column */
  int32_t numparttot;         /* particles created over all ranks */
  float colmasstotal;         /* air mass of the box [kg] */
  float xmassperparticle;     /* colmasstotal / numparttot */
} fpb_domainfill_info;
int fpb_init_domainfill(fpb_handle *h, float xpoint1, float ypoint1, float xpoint2, float ypoint2,
                        int32_t itsplit, int32_t *numpart /* may be NULL */,
                        fpb_domainfill_info *info /* may be NULL */);
int fpb_boundcond_domainfill(fpb_handle *h, int32_t itime, int32_t loutend, int32_t *numpart /* may be NULL */,
                             int32_t *n_created /* may be NULL */);

/* calcpar + verttransform_ecmwf on the device.
 * fpb_calcpar_verttransform replaces the pair `call calcpar(n,uuh,vvh,pvh,metdata_format)` /
 * `call verttransform_ecmwf(n,uuh,vvh,wwh,pvh)` of getfields (src/getfields.f90:126-129,161-164,
 * 177-180): the wind field as readwind_ecmwf leaves it is copied to the device once and becomes
 * met slot `slot` there -- friction velocity, Obukhov length, mixing height, convective velocity
 * scale and thermal tropopause (src/calcpar.f90:78-258, scalev.f90, obukhov.f90, richardson.f90),
 * potential vorticity on the eta levels (src/calcpv.f90),
 * u, v, T, q, PV, density and density gradient on the height levels, the vertical wind in m/s with
 * the slope term of the eta surfaces, polar-stereographic winds and pole rows, the cloud /
 * precipitation classes -- parameterised from the humidity or, with readclouds, from the input's cloud
 * water, with the column total ctwc (src/verttransform_ecmwf.f90:198-724; readclouds: the scalar
 * cloudh_min the reference carries from column to column starts at 0 in every column here) -- so the transformed
 * fields never exist on the host.  The same arithmetic in the same order as the reference (no
 * contraction, transcendentals evaluated in double): bit-identical fields.  The slot is then what
 * fpb_upload_met would have produced; when fpb_set_convection has been called with the same nuvz
 * the convection fields of the slot (fpb_upload_convmet) are set as well.
 * Needs fpb_set_vertical first.  ECMWF layout only: nz = nuvz = nwz.  `height` of fpb_config must be
 * what verttransform_ecmwf's first call derives (fpbh_verttransform_heights in fpb_host.h).
 * Not built: dry-deposition velocities (getvdep: land-use inventory; upload vdep with
 * fpb_upload_vdep), the NCEP/GFS variant.
 * fpb_fetch_met copies a slot back in the reference's padded layout (any pointer may be NULL): for
 * the parts of a host model that still read the transformed fields, and for the tests. */
typedef struct fpb_rawmet_ptrs {
  const float *uuh, *vvh, *tth, *qvh; /* (nxmax, nymax, nuvzmax) */
  const float *pvh;                   /* (nxmax, nymax, nuvzmax); NULL: calcpv runs on the device */
  const float *wwh;                   /* (nxmax, nymax, nwzmax) */
  const float *ps, *tt2, *td2, *sshf, *surfstr, *lsprec, *convprec, *tcc; /* (nxmax, nymax) */
  const float *excessoro;             /* (nxmax, nymax), lsubgrid = 1 only */
  const float *clwch, *ciwch;         /* (nxmax, nymax, nuvzmax), readclouds only: cloud liquid (+ ice: ciwch NULL when
                                       * the input holds their sum, `sumclouds`) water content */
} fpb_rawmet_ptrs;
typedef struct fpb_met_out_ptrs {
  float *uu, *vv, *ww, *rho, *drhodz, *tt, *qv, *pv, *uupol, *vvpol; /* (nxmax, nymax, nzmax) */
  float *hmix, *ustar, *wstar, *oli, *tropopause;                    /* (nxmax, nymax) */
  int8_t *clouds;                                                     /* (nxmax, nymax, nzmax) */
  float *ctwc;                                                        /* (nxmax, nymax), wet deposition only */
} fpb_met_out_ptrs;
int fpb_set_vertical(fpb_handle *h, int32_t nuvz, int32_t nwz, int32_t nuvzmax, int32_t nwzmax, const float *akm,
                     const float *bkm, const float *akz, const float *bkz);
int fpb_calcpar_verttransform(fpb_handle *h, int32_t slot, const fpb_rawmet_ptrs *raw, int32_t lsubgrid,
                              float *device_ms /* may be NULL; [0] upload + kernels, [1] kernels only (ms) */);
int fpb_upload_vdep(fpb_handle *h, int32_t slot, const float *vdep /* (nxmax, nymax, maxspec) */);
int fpb_fetch_met(fpb_handle *h, int32_t slot, const fpb_met_out_ptrs *out);
/* The same for nested input grid `nest` (1-based), replacing the calcpar_nests / verttransform_nests pair of
 * getfields (src/getfields.f90:131-134; src/calcpar_nests.f90, src/verttransform_nests.f90, src/calcpv_nests.f90):
 * the raw arrays have the nest's padded extents (nxmaxn, nymaxn, ..); dxn, dyn, xlon0n, ylat0n are the nest's
 * own grid constants as gridcheck_nests read them (src/gridcheck_nests.f90: they enter cosf, the latitude
 * dependent tropopause search and calcpv); bit-identical fields, as for the mother grid.  The nest's slot is
 * then what fpb_upload_met_nest would have produced (and fpb_upload_convmet_nest, when fpb_set_convection was
 * called with the same nuvz).  Not built: readclouds_nest, getvdep_nests (upload vdepn: fpb_upload_met_nest
 * is still the way to hand over deposition velocities).  fpb_fetch_met_nest: uupol / vvpol / pv / qv do not
 * exist for nests (left untouched); tt only when the run has wet deposition. */
int fpb_calcpar_verttransform_nest(fpb_handle *h, int32_t slot, int32_t nest, const fpb_rawmet_ptrs *raw,
                                   int32_t lsubgrid, float dxn, float dyn, float xlon0n, float ylat0n,
                                   float *device_ms /* may be NULL */);
int fpb_fetch_met_nest(fpb_handle *h, int32_t slot, int32_t nest, const fpb_met_out_ptrs *out);

/* Convective mixing (LCONVECTION = 1, the shipped default).
 * fpb_convmix replaces `call convmix(itime,metdata_format)` (src/timemanager.f90:183-193,258-263;
 * src/convmix.f90:60-196): the active particles are grouped by grid column, every occupied column's
 * sounding is interpolated in time and run through calcmatrix / the Emanuel scheme
 * (src/calcmatrix.f90, src/convect43c.f90; one thread per column), and the particles of the
 * convecting columns get their new height from redist (src/redist.f90) -- all on the device, the
 * particles stay resident.  The cloud base mass flux of every column (cbaseflux) is kept on the
 * device between calls.  The column arithmetic is the reference's statement by statement and is
 * evaluated without contraction in every math mode.  FPB_RNG_REFERENCE replays the reference's
 * ran3 stream in its sort2 visiting order (bit-identical heights); the Philox modes draw the
 * uniform from the particle's counter stream.
 *   fpb_set_convection  once: nuvz, the padded level extent of tth/qvh (nuvzmax), nconvlev
 *                       (src/gridcheck_ecmwf.f90:553-566) and akz, bkz, akm, bkm (1:nuvz)
 *   fpb_upload_convmet  next to every fpb_upload_met: ps, tt2, td2 (nxmax,nymax) and tth, qvh
 *                       (nxmax,nymax,nuvzmax) of the time level (src/com_mod.f90:372-417)
 *   fpb_upload_convmet_nest  the same fields of nested input grid `nest` (psn, tt2n, td2n (nxmaxn,
 *                       nymaxn); tthn, qvhn (nxmaxn,nymaxn,nuvzmax)): a particle inside a nest takes
 *                       part in the nest's columns only (the innermost one, src/convmix.f90:100-134,
 *                       198-281), with the nest's own cbasefluxn
 * ECMWF fields (metdata_format = GRIBFILE_CENTRE_ECMWF).  With iflux = 1 every displaced particle's
 * calcfluxes call (src/convmix.f90:205-218) adds to the flux array fpb_fetch_fluxes returns. */
typedef struct fpb_conv_ptrs {
  const float *ps, *tt2, *td2;
  const float *tth, *qvh;
} fpb_conv_ptrs;
int fpb_set_convection(fpb_handle *h, int32_t nuvz, int32_t nuvzmax, int32_t nconvlev, const float *akz,
                       const float *bkz, const float *akm, const float *bkm);
int fpb_upload_convmet(fpb_handle *h, int32_t slot, const fpb_conv_ptrs *met);
int fpb_upload_convmet_nest(fpb_handle *h, int32_t slot, int32_t nest, const fpb_conv_ptrs *met);
int fpb_convmix(fpb_handle *h, int32_t itime, int32_t *ncolumns /* occupied columns, may be NULL */,
                int32_t *nconvecting /* of those, columns with convection, may be NULL */);

/* wetgridunc(0:numxgrid-1,0:numygrid-1,maxspec,maxpointspec_act,nclassunc,
 * maxageclass) and the nested twin (may be NULL): cumulative, never zeroed,
 * decayed by fpb_scale_depgrids like drygridunc */
int fpb_fetch_wetgrids(fpb_handle *h, float *wetgridunc, float *wetgriduncn);

/* before concoutput*: src/timemanager.f90:376 (MPI build:
 * mpif_tm_reduce_grid, src/timemanager_mpi.f90:468).  Copies the device
 * grids into the caller's arrays in the reference layout
 *   gridunc (numxgrid,numygrid,numzgrid,maxspec,maxpointspec_act,nclassunc,maxageclass)
 *   drygridunc (numxgrid,numygrid,maxspec,maxpointspec_act,nclassunc,maxageclass)
 *   creceptor (maxreceptor, maxspec)
 * (src/outgrid_init.f90:192-201); NULL pointers are skipped.  With
 * zero_conc != 0 gridunc/griduncn/creceptor are then zeroed on the device
 * (src/concoutput.f90:719-720); deposition grids stay cumulative. */
int fpb_fetch_grids(fpb_handle *h, float *gridunc, float *griduncn,
                    float *drygridunc, float *drygriduncn, float *creceptor,
                    int32_t zero_conc);
/* decay of deposited mass: src/timemanager.f90:269-304;
 * factor[ks] multiplies drygridunc(:,:,ks,...) (and the nested twin). */
int fpb_scale_depgrids(fpb_handle *h, const float *factor_per_species);

/* Device views for the host's collective (one NCCL reduce per output
 * interval, the mpif_tm_reduce_grid slot src/mpi_mod.f90:2395-2579).
 * which: 0 gridunc, 1 griduncn, 2 drygridunc, 3 drygriduncn, 4 creceptor.
 * The device layout is packed to nspec species (not maxspec). */
int fpb_grid_device_ptr(fpb_handle *h, int32_t which, void **dptr,
                        size_t *nfloats);
int fpb_zero_conc_grids(fpb_handle *h);

/* The grid exchange of a multi-GPU run: one process per GPU, particles partitioned, full met replica
 * and private grids per GPU; at each output interval the grids are summed to rank 0 -- the slot of
 * mpif_tm_reduce_grid(_nest) (src/mpi_mod.f90:2395-2579, called at src/timemanager_mpi.f90:468-485)
 * -- as ONE NCCL reduce group over NVLink, on a high-priority side stream so that the next interval's
 * steps proceed meanwhile.  NCCL is bound at run time (dlopen), libfpb.so does not link against it.
 *   fpb_comm_unique_id   rank 0 creates the 128-byte NCCL id; the host broadcasts it (MPI_Bcast in
 *                        the Fortran host, src/mpi_mod.f90:162 mpif_init is where the ranks meet)
 *   fpb_comm_init        every rank, after fpb_init: joins the communicator (nranks = 1: no NCCL)
 *   fpb_reduce_grids_begin  gridunc, griduncn, drygridunc(n), wetgridunc(n) are copied to staging
 *                        buffers (creceptor too) and gridunc/griduncn/creceptor zeroed (src/concoutput.f90:719-720) on the
 *                        engine's stream; the staging buffers are summed to rank 0.  Returns at once.
 *   fpb_reduce_grids_end waits; on rank 0 the sums are written to the caller's arrays in the
 *                        reference layout (see fpb_fetch_grids; NULL = skip).  Deposition grids stay
 *                        cumulative per rank, rank 0 receives their sum (the drygridunc0 of the reference).
 *   fpb_reduce_grids_device  device pointer of a summed staging buffer (which: 0 gridunc, 1 griduncn,
 *                        2 drygridunc, 3 drygriduncn, 4 wetgridunc, 5 wetgriduncn, 6 creceptor; nspec-packed) and
 *                        the device time of the last reduce [ms], for a host that post-processes on
 *                        the GPU (fpb_concoutput_sparse) or reports the exchange cost. */
int fpb_comm_unique_id(void *id128);
int fpb_comm_init(fpb_handle *h, const void *id128, int32_t rank, int32_t nranks);
int fpb_reduce_grids_begin(fpb_handle *h);
int fpb_reduce_grids_end(fpb_handle *h, float *gridunc, float *griduncn, float *drygridunc,
                         float *drygriduncn, float *wetgridunc, float *wetgriduncn, float *creceptor);
int fpb_reduce_grids_device(fpb_handle *h, int32_t which, void **dptr, size_t *nfloats, float *reduce_ms);
int fpb_comm_finalize(fpb_handle *h);

/* Re-order the device-resident particles by met-grid cell for gather
 * locality.  Slot identity seen through push/pull is preserved. */
int fpb_sort_particles(fpb_handle *h);

/* CUDA stream the engine launches on (for the caller's events), and the
 * number of kernels the engine has launched so far. */
void *fpb_stream(fpb_handle *h);
int64_t fpb_launch_count(fpb_handle *h);

/* Device time (ms, CUDA events on the engine's stream) of the kernels of the
 * most recent fpb_step / fpb_conccalc call; for the roofline report. */
int fpb_kernel_times(fpb_handle *h, float *step_ms, float *conccalc_ms);

/* Export the rannumb table the engine holds (validation: hand the same
 * Gaussian stream to another implementation). */
int fpb_get_rannumb(fpb_handle *h, float *out, int32_t n);

#ifdef __cplusplus
}
#endif
#endif /* FPB_H */
