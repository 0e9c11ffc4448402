// fpb_domainfill.cuh -- init_domainfill on the device (BASELINE
// configs[4]): the domain-filling particles are created where they live, no particle row crosses
// the bus.  Kernels in fpb_domainfill.cu.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

#include "fpb_device.cuh"

struct DomainfillArgs {
  DevCfg cfg;
  DevParticles p;          // rows == slots (the call starts from an empty, unpermuted state)
  const float4 *A1;        // {uu,vv,ww,rho} of Fortran slot 1 (literal in the reference), [k][jy][ix]
  const float *T1;         // tt of slot 1
  const float *height;     // [nz]
  int nx0, nx1, ny0, ny1;  // nx_we(1:2), ny_sn(1:2)
  int ncolx, ncols;        // columns per row / in the box, column = (jy - ny0) * ncolx + (ix - nx0)
  const float *gridarea;   // [ny], indexed by jy
  float *colmass;          // [ncols]
  float *total;            // [1] colmasstotal (sequential float sum in the reference's loop order)
  int32_t *ncolumn;        // [ncols] particles per column
  unsigned *colstart;      // [ncols] first global particle index of the column (exclusive scan)
  unsigned *block_sums;    // [ceil(ncols / 1024)]
  int *out;                // [0] numparttot, [1] numcolumn, [2] local numpart (max live slot + 1),
                           // [3] columns whose pressure profile matched 0 or > 1 layers (reference RNG)
  float npart1;            // real(npart(1))
  int itsplit;
  int id_stride, id_offset; // this rank keeps global particle g when g % id_stride == id_offset,
                            // in local slot g / id_stride (src/init_domainfill_mpi.f90:86-104)
  const float *uniforms;    // reference RNG: the call's ran1 stream, in the reference's draw order
  const unsigned long long *u_off; // [ncols] first stream entry of each column
};

void fpb_domainfill_launch(const DomainfillArgs &a, cudaStream_t st, int64_t *launches, int phase);
// pressure profiles pp(kz) = rho*r_air*tt of the listed columns (init_domainfill.f90:322-324), out[n][nz]
void fpb_domainfill_profiles(const DomainfillArgs &a, const int2 *cols, int n, float *out, cudaStream_t st);

// ---- boundcond_domainfill (src/boundcond_domainfill.f90:54-560): inflow boundary of a limited box
// One release location of the boundary: what does not change during the run is worked out once by
// the host from the heights memorised by init_domainfill (:104-143, :343-381).
struct BcLoc {
  int gx, gy;        // grid point whose wind and density give the mass flux
  int indz;          // 1-based model level below the release height
  float dz1, dz2, dz;
  float boundarea;
  float za, zb;      // height of a new particle: za (+ u * zb at an interior height)
  int idx;           // jy (west/east boundary) or ix (south/north)
  int flags;         // see BC_* below
};
enum { BC_WE = 1, BC_K2 = 2, BC_EDGE_LOW = 4, BC_EDGE_HIGH = 8, BC_ZDRAW = 16 };

struct BoundcondArgs {
  DevCfg cfg;                  // cfg.itime = the call's itime
  DevParticles p;
  const int32_t *row_of_slot;
  int permuted, numpart_old;
  const float4 *A[2];          // {uu,vv,ww,rho} of memind(1), memind(2)
  float dt1, dt2, dtt;
  int nx0, nx1, ny0, ny1, check_x;
  const BcLoc *loc;
  int nloc;
  float *acc_mass;             // [nloc]
  int32_t *mmass;              // [nloc] particles each location releases in this call
  const int32_t *first;        // [nloc + 1] exclusive prefix of mmass
  const float *uniforms;       // reference RNG: the call's ran1 stream in the reference's draw order
  const int32_t *u_off;        // [nloc] first stream entry of each location
  int n_new;                   // particles of the call over all ranks
  int n_mine, r0;              // this rank creates r = r0, r0 + stride, ... (n_mine of them)
  int numparticlecount;        // before the call
  float xmassperparticle;
  int itsplit, id_stride;
  unsigned *block_counts;      // [ceil(maxpart / 1024) + 1]
  int *out;                    // [0] max(new slot) + 1, [1] free slots
};
// phase 0: terminate + fluxes -> mmass; phase 1: create the particles
void fpb_boundcond_launch(const BoundcondArgs &a, cudaStream_t st, int64_t *launches, int phase);
