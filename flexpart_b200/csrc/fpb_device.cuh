// fpb_device.cuh -- device-side types shared by the engine's kernels.
//
// Data layout in HBM (DESIGN.md "Data layout"):
//   met, per time slot:  A[k][jy][ix] = float4{uu,vv,ww,rho}
//                        G[k][jy][ix] = drhodz, T[k][jy][ix] = tt,
//                        P[k][jy][ix] = float2{uupol,vvpol}
//                        S[jy][ix]    = float4{hmix,ustar,wstar,oli}
//                        trop[jy][ix], vdep[ks][jy][ix]
//     -> the 4 values a bilinear corner needs sit in one 16-B word and the
//        x-neighbour is the adjacent word: a corner PAIR is one 32-B sector.
//   particles: structure of arrays, one array per com_mod variable
//     (src/com_mod.f90:675-695), particle index fastest.
//   grids: reference index order (src/outgrid_init.f90:192-201) packed to
//     nspec species.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/fpb.h"

#define FPB_MAXNZ 160

// two x-neighbours of the wind/density word in one aligned 32-byte sector: AP[i] = {A[i], A[i+1]}
// (the same values a second time; one 256-bit load per corner pair in interp_wind)
struct __align__(32) MetPair { float4 a, b; };

struct DevMetSlot {
  const float4 *A;
  const MetPair *AP; // or null
  const float *G;  // drhodz
  const float *T;  // tt (settling only)
  const float2 *P; // uupol, vvpol (poleward of the switch latitudes only)
  const float4 *S;
  const float *trop;
  const float *vdep;
  const float4 *R;   // {lsprec, convprec, tcc, ctwc}[jy][ix] (wet deposition only)
  const int8_t *C;   // clouds[k][jy][ix] (wet deposition only)
};

struct DevParticles {
  double *xtra1, *ytra1;
  float *ztra1;
  int32_t *itra1, *npoint, *nclass, *idt, *itramem, *itrasplit;
  float *uap, *ucp, *uzp, *us, *vs, *ws;
  int16_t *cbt;
  float *xmass1;      // [nspec][maxpart]
  float *xscav_frac1; // [nspec][maxpart] or null
  int32_t *slot;      // slot[row] = caller-visible slot index of device row `row`
  int32_t maxpart;
};

// Scalars the kernels read; passed by value as a __grid_constant__ kernel
// parameter (constant bank), one copy per launch so several engine handles
// can coexist in a process.
struct DevCfg {
  // met grid
  int nx, ny, nz, nxd, nyd; // nxd/nyd: device extents of the packed met arrays
  int nymax;
  int nxmin1, nymin1;
  float dx, dy, xlon0, ylat0, dxconst, dyconst;
  int xglobal, nglobal, sglobal;
  float switchnorthg, switchsouthg;
  float northpolemap[9], southpolemap[9];
  float eps;
  // nested input grids (src/gridcheck_nests.f90:359-389); nxdn/nydn: device extents
  int numbnests;
  int nxdn[FPB_MAXNESTS], nydn[FPB_MAXNESTS];
  float xln[FPB_MAXNESTS], yln[FPB_MAXNESTS], xrn[FPB_MAXNESTS], yrn[FPB_MAXNESTS];
  float xresoln[FPB_MAXNESTS], yresoln[FPB_MAXNESTS];
  // command
  int ldirect, lsynctime, method, mintime, ifine;
  int turbswitch, cblflag, mdomainfill, mquasilag, lsettling, turboff;
  float ctl, fine, d_trop, d_strat, turbmesoscale;
  int ind_samp, ioutputforeachrelease, lusekerneloutput, lparticlecountoutput;
  int drydep, drybkdep, wetbkdep, nested_output;
  int linit_cond;
  // species
  int nspec;
  float decay[FPB_MAXSPEC];
  int drydepspec[FPB_MAXSPEC];
  float density[FPB_MAXSPEC], dquer[FPB_MAXSPEC], vsetaver[FPB_MAXSPEC],
      cunningham[FPB_MAXSPEC];
  int wetdepspec[FPB_MAXSPEC];
  float weta_gas[FPB_MAXSPEC], wetb_gas[FPB_MAXSPEC], crain_aero[FPB_MAXSPEC], csnow_aero[FPB_MAXSPEC],
      ccn_aero[FPB_MAXSPEC], in_aero[FPB_MAXSPEC], henry[FPB_MAXSPEC];
  int readclouds, readclouds_nest[FPB_MAXNESTS];
  int nageclass;
  int lage[FPB_MAXAGECLASS];
  // out grids
  int numxgrid, numygrid, numzgrid;
  float dxout, dyout, xoutshift, youtshift;
  float outheight[FPB_MAXZGRID];
  int numxgridn, numygridn;
  float dxoutn, dyoutn, xoutshiftn, youtshiftn;
  int maxpointspec_act, nclassunc, maxageclass;
  int numreceptor;
  float xreceptor[FPB_MAXRECEPTOR], yreceptor[FPB_MAXRECEPTOR],
      receptorarea[FPB_MAXRECEPTOR];
  int numpoint;
  // per-step
  int itime, ldeltat;
  int memind[2];  // 0-based device slot of the older / newer field
  int memtime[2];
  int lwindinterv;
  int maxrand;
  int rng_mode;
  unsigned long long seed;
  int part_id_stride, part_id_offset;
  int numpart;
  float weight; // conccalc
};

// Hand-over between fpb_pbl_kernel and fpb_finish_kernel (fpb_step.cuh): what
// the rest of advance() needs from the sub-step loop.  Rows with
// itra1 == itime get `flags`; the float4 rows only when SC_PBL is set.
struct DevScratch {
  int32_t *flags; // SC_* bits (fpb_step.cuh)
  float4 *s0; // dxsave, dysave, dawsave, dcwsave
  float4 *s1; // u, v, w, indz of the last sub-step (bits)
  int2 *s2;   // nrand, itimec
  float *prob; // [nspec][maxpart], dry-deposition probability (drydep runs only)
};

// Deterministic deposition (FPB_SCATTER_DETERMINISTIC): instead of atomics the kernels write one
// (cell key, values) record per touched cell -- record id = 4*(slot - slot_base) + corner -- and the
// records are then sorted by key and added run by run in record (= particle) order, like conccalc's
// (fpb_scatter.cu).  grid 0 = mother output grid, 1 = nested.  keys == null: atomics.
struct DevDepRecords {
  unsigned *keys[2]; // [nrec], pre-filled with 0xffffffff
  float *vals[2];    // [nspec][nrec]
  size_t nrec;
  int slot_base;
};

struct DevStepArgs {
  DevCfg cfg;
  DevMetSlot met[2];   // [0] = memind(1) (older field), [1] = memind(2)
  DevMetSlot met_lit1; // Fortran slot 1, for the reference's literal-slot reads
  DevMetSlot metn[FPB_MAXNESTS][2];         // nested input grids, ordered like met[]
  const float *tropn_lit1[FPB_MAXNESTS];    // tropopausen(:,:,1,1,l), src/advance.f90:263
  DevParticles p;
  const float *height;   // [nz] 0-based (height[0] = level 1)
  const float *rannumb;  // 0-based table, rannumb[i-1] = Fortran rannumb(i)
  const int32_t *npart;  // [numpoint]
  const float *xmass;    // [nspec][numpoint] (device-packed)
  const int32_t *nrand_init; // reference RNG mode: per-slot nrand for initialize
  const int32_t *nrand_adv;  //                      and for advance
  float *drygridunc, *drygriduncn;
  unsigned long long *stats; // 8 counters, fpb_step_stats order
  int *work_counter;         // next unclaimed particle row (persistent sub-step kernel)
  DevScratch sc;
  DevDepRecords dep;         // dry deposition records (deterministic mode), else keys = null
  float grid_frac;           // persistent sub-step grid as a fraction of one resident wave (0 = 1: all of it)
};

struct DevConcArgs {
  DevCfg cfg;
  DevMetSlot met[2];   // ordered like DevStepArgs::met
  DevParticles p;
  const float *height;
  float *gridunc, *griduncn;
  float *crec_acc; // [numreceptor][nspec] accumulators of c(ks)
  int slot_base;   // deterministic path on a row chunk (fpb_step_host): record id = 4*(slot - slot_base) + corner
  float *rec_vals; // deterministic receptors: [numreceptor*nspec][nslots] xmass1*kernel of every particle, or null
  int rec_nslots;
};

// wetdepo (src/wetdepo.f90): one time level per grid, chosen on the host the way
// get_wetscav does (src/get_wetscav.f90:113-117)
struct DevWetArgs {
  DevCfg cfg;
  DevMetSlot met;                 // mother grid, time level n
  DevMetSlot metn[FPB_MAXNESTS];  // nested grids, time level n
  DevParticles p;
  const float *height;
  float *wetgridunc, *wetgriduncn;
  int ltsample;
  DevDepRecords dep; // wet deposition records (deterministic mode), else keys = null
};

// backward-run receptor scavenging of the particle loop (src/timemanager.f90:563-598)
struct DevBkdepArgs {
  DevWetArgs w;                        // cfg, particles, height; met / metn = the time level get_wetscav picks
  DevMetSlot vmet[2];                  // memind(1), memind(2): vdep for get_vdep_prob
  DevMetSlot vmetn[FPB_MAXNESTS][2];
  const float *zpoint1, *zpoint2;      // [numpoint] release heights (wet scavenging)
};

// releaseparticles on the device (fpb_release.cuh)
struct DevReleaseArgs {
  DevCfg cfg;                 // cfg.itime = the release time
  DevParticles p;
  const int32_t *row_of_slot;
  int permuted;               // rows != slots
  int numpart_old;            // slots >= numpart_old have never been used
  int numpoint, n_new, itsplit;
  float ztop;                 // height(nz)
  const float *xpoint1, *ypoint1, *xpoint2, *ypoint2, *zpoint1, *zpoint2;
  const int32_t *offsets;     // [numpoint + 1] first new particle of each point
  const float *uniforms;      // [4 * n_new] ran1 stream of the call, or null (Philox)
  const float *xmass;         // [nspec][numpoint]
  const int32_t *npart;
  unsigned *block_counts;     // [ceil(maxpart / 1024)]
  int *out;                   // [0] max(new slot) + 1 (atomicMax), [1] free slots
};

// particle splitting (fpb_release.cuh)
struct DevSplitArgs {
  DevCfg cfg;                 // cfg.itime
  DevParticles p;
  const int32_t *row_of_slot;
  int permuted, numpart_old;
  unsigned *block_counts;
  int *total;                 // candidates found
};

#ifdef __CUDACC__
// Philox4x32-10 (Salmon et al. 2011), counter = (particle id, time, stream, block)
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; r++) {
    uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
    uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += W0;
    key.y += W1;
  }
  return ctr;
}
__device__ __forceinline__ float u01(uint32_t x) { return (float)(x >> 8) * (1.0f / 16777216.0f); }
#endif

// launchers (one set per math mode; defined in fpb_kernels.cu compiled twice)
#define FPB_DECL_LAUNCHERS(SUF)                                               \
  void fpbk_init_##SUF(const DevStepArgs &a, cudaStream_t st);                \
  void fpbk_step_##SUF(const DevStepArgs &a, cudaStream_t st);                \
  void fpbk_conccalc_##SUF(const DevConcArgs &a, cudaStream_t st);            \
  void fpbk_receptor_##SUF(const DevConcArgs &a, cudaStream_t st);            \
  void fpbk_wetdepo_##SUF(const DevWetArgs &a, cudaStream_t st);              \
  void fpbk_bkdep_##SUF(const DevBkdepArgs &a, cudaStream_t st);              \
  void fpbk_release_##SUF(const DevReleaseArgs &a, cudaStream_t st);          \
  void fpbk_split_##SUF(const DevSplitArgs &a, cudaStream_t st);              \
  void fpbk_conc_emit_##SUF(const DevConcArgs &a, int nest_sel, unsigned *keys, \
                            float *vals, size_t nrec, cudaStream_t st);
FPB_DECL_LAUNCHERS(fast)
FPB_DECL_LAUNCHERS(strict)
