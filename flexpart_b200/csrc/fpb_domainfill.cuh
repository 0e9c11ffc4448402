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
