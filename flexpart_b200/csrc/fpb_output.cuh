// fpb_output.cuh -- concoutput's per-(species, release, age class) work on the device
// (SURVEY.md section 8f, rank 4): class mean, unit conversion and the sparse run-length dump.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

struct SparseDumpArgs {
  const float *grid;        // first class slice of (ks, kp, nage): [ncells]
  size_t class_stride;      // floats between the slices of consecutive uncertainty classes
  int nclassunc, ncells;
  const float *geom;        // which = 0: volume[ncells]; 1, 2: area[ncells]
  int which;                // 0 concentration, 1 dry deposition, 2 wet deposition
  int ldirect;
  float outnum, tot_mu, loutaver_abs;
  int index_offset;         // added to the linear cell index (numxgrid*numygrid for concentrations)
  unsigned *block_counts;   // [2 * nblocks] scratch: non-zero cells, run starts
  int32_t *out_i;           // sparse_dump_i
  float *out_r;             // sparse_dump_r
  int *counts;              // [0] sp_count_i, [1] sp_count_r
};

void fpb_sparse_dump(const SparseDumpArgs &a, cudaStream_t st);
