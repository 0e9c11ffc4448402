// fpb_scatter.cuh -- deterministic accumulation: sort (cell key, record id)
// pairs with a stable LSD radix sort, then add every run of equal keys to its
// grid cell in record order (one writer per cell, no atomics).  Record order
// is particle order, so a cell receives its contributions in exactly the
// sequence the reference's serial loop adds them (src/conccalc.f90:50-444):
// the result is bit-identical to the sequential float accumulation.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "fpb_device.cuh"

struct ScatterWork {
  unsigned *keys[2] = {nullptr, nullptr};
  unsigned *ids[2] = {nullptr, nullptr};
  float *vals = nullptr;
  unsigned *hist = nullptr;
  size_t cap_rec = 0, cap_vals = 0, cap_hist = 0;
};

int scatter_conccalc_deterministic(ScatterWork &w, const DevConcArgs &a, bool strict,
                                   cudaStream_t st, int64_t *launches);
// deposition records written by fpb_finish_kernel / fpb_wetdepo_kernel (DevDepRecords) -> grid
int scatter_records_deterministic(ScatterWork &w, const unsigned *keys, const float *vals, size_t nrec, int nspec,
                                  float *grid, int nxyz, unsigned long long ncell, cudaStream_t st, int64_t *launches);
// receptor sums in particle order: vals[nsums][nslots] -> acc[nsums] += sequential sum
int scatter_receptor_ordered(const float *vals, int nsums, int nslots, float *acc, cudaStream_t st);
// stable sort of n (key, id) pairs on keys' low `bits` bits; result in
// w.keys[*out] / w.ids[*out]
int scatter_sort_pairs(ScatterWork &w, size_t n, int bits, cudaStream_t st, int64_t *launches,
                       int *out);
int scatter_reserve(ScatterWork &w, size_t nrec, int nspec);
void scatter_free(ScatterWork &w);
const char *scatter_error();
