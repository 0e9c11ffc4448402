// fpb_convmix.cuh -- arguments of the convective-mixing kernels (fpb_convect.cu)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "fpb_device.cuh"

struct ConvmixArgs {
  DevCfg cfg;                 // cfg.itime, cfg.memtime set
  DevParticles p;
  int nrows;                  // rows to look at (live rows lead the arrays after a cell sort)
  int nuvz, nconvlev;
  const float *akz, *bkz, *akm, *bkm; // device, 1-based (element k at [k])
  // grid g = 0: mother grid, g = l: nested input grid l (src/convmix.f90:198-281)
  const float2 *CT[FPB_MAXNESTS + 1][2]; // {tth, qvh}[k][jy][ix] of memind(1), memind(2), k = 0..nuvz-1 (Fortran level k+1)
  const float4 *CS[FPB_MAXNESTS + 1][2]; // {ps, tt2, td2, -}[jy][ix]
  float *cbaseflux[FPB_MAXNESTS + 1];    // cloud base mass flux of every column, kept between calls
  int gnx[FPB_MAXNESTS + 1], gnxd[FPB_MAXNESTS + 1], gnyd[FPB_MAXNESTS + 1]; // nx of igrid = 1 + jy*nx + ix; device extents
  int col_bits;               // key = (grid << col_bits) | (jy*nx + ix)
  int ecmwf_eps;              // nest test with the eps margin (metdata_format = ECMWF, src/convmix.f90:104-110)
  float ztop;                 // height(nz)
  unsigned *keys, *ids;       // column key / row of every row (sort input)
  const unsigned *sorted_ids; // rows in column order
  int32_t *key_by_slot;       // reference RNG: igrid(ipart) in slot order, or null
  unsigned *block_counts;
  int32_t *colidx;            // [nrows] column index of every sorted position
  unsigned *col_key;          // [ncols]
  int32_t *col_start;         // [ncols + 1] first sorted position of every column
  int32_t *col_lconv;         // [ncols] nconvtop when the column convects, else 0
  float *pool;                // work pool: CONV_BATCH columns x conv_pool_floats()
  float *pool2;               // the columns' final MENT and FMASS, contiguous per column: CONV_BATCH x fpb_convmix_pool2_floats()
  void *col_state;            // [ncols] fpbconv::ConvState: what the halves of the column code hand over
  uint8_t *draws;             // reference RNG: [slot] the particle draws a uniform
  const float *rn_by_slot;    // reference RNG: the uniforms, by slot; null: Philox
  float *flux;                // iflux = 1: calcfluxes after redist (src/convmix.f90:205-218), else null
};

void fpb_convmix_keys(const ConvmixArgs &a, cudaStream_t st);
void fpb_convmix_heads(const ConvmixArgs &a, const unsigned *sorted_keys, int *total, cudaStream_t st);
void fpb_convmix_columns(const ConvmixArgs &a, int c0, int c1, cudaStream_t st);
void fpb_convmix_redist(const ConvmixArgs &a, int c0, int i0, int i1, int mode, cudaStream_t st);
size_t fpb_convmix_pool_floats(int nuvz, int nconvlev);
size_t fpb_convmix_pool2_floats(int nconvlev); // per column: contiguous MENT (assembly layout) + FMASS
size_t fpb_convmix_state_bytes();
