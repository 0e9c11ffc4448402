// fpb_scatter.cu -- stable LSD radix sort + ordered segmented sum (see
// fpb_scatter.cuh).  Integer/ordering work only: no math-mode variants.
#include <stdio.h>

#include <string>

#include "fpb_scatter.cuh"

static std::string s_err;
const char *scatter_error() { return s_err.c_str(); }
#define SCK(call)                                                             \
  do {                                                                        \
    cudaError_t e_ = (call);                                                  \
    if (e_ != cudaSuccess) {                                                  \
      s_err = std::string(#call) + ": " + cudaGetErrorString(e_);             \
      return 1;                                                               \
    }                                                                         \
  } while (0)

namespace {

constexpr int RADIX_BITS = 8, RADIX = 1 << RADIX_BITS;
constexpr int SORT_THREADS = 256, SORT_WARPS = SORT_THREADS / 32;
#ifndef FPB_SORT_ROUNDS
#define FPB_SORT_ROUNDS 16
#endif
constexpr int SORT_ROUNDS = FPB_SORT_ROUNDS;          // elements per thread
constexpr int SORT_TILE = SORT_THREADS * SORT_ROUNDS; // elements per block

__global__ void iota_fill_kernel(unsigned *keys, unsigned *ids, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    keys[i] = 0xffffffffu;
    ids[i] = (unsigned)i;
  }
}

// pass 1: per-block digit histogram -> hist[digit * nblocks + block]
__global__ void __launch_bounds__(SORT_THREADS)
radix_hist_kernel(const unsigned *keys, size_t n, int shift, unsigned *hist, int nblocks) {
  __shared__ unsigned sh[RADIX];
  for (int d = threadIdx.x; d < RADIX; d += SORT_THREADS) sh[d] = 0;
  __syncthreads();
  const size_t base = (size_t)blockIdx.x * SORT_TILE;
  for (int r = 0; r < SORT_ROUNDS; r++) {
    size_t i = base + (size_t)r * SORT_THREADS + threadIdx.x;
    if (i < n) atomicAdd(&sh[(keys[i] >> shift) & (RADIX - 1)], 1u);
  }
  __syncthreads();
  for (int d = threadIdx.x; d < RADIX; d += SORT_THREADS) hist[(size_t)d * nblocks + blockIdx.x] = sh[d];
}

// pass 2: one block per digit scans that digit's row of per-block counts in
// place (exclusive) and leaves the digit total in totals[digit]; the scatter
// kernel turns the 256 totals into digit base offsets itself.
__global__ void __launch_bounds__(256) digit_scan_kernel(unsigned *hist, int nblocks, unsigned *totals) {
  __shared__ unsigned warp_tot[8];
  __shared__ unsigned carry;
  unsigned *row = hist + (size_t)blockIdx.x * nblocks;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int base = 0; base < nblocks; base += 256) {
    const int i = base + threadIdx.x;
    const unsigned v = (i < nblocks) ? row[i] : 0u;
    unsigned x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      unsigned y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) warp_tot[wid] = x;
    __syncthreads();
    unsigned before = carry;
    for (int w = 0; w < wid; w++) before += warp_tot[w];
    if (i < nblocks) row[i] = before + (x - v);
    __syncthreads();
    if (threadIdx.x == 255) carry = before + x;
    __syncthreads();
  }
  if (threadIdx.x == 0) totals[blockIdx.x] = carry;
}

// pass 3: stable scatter.  Element order inside a block is (round, thread);
// the rank of an element among equal digits = earlier rounds + earlier warps
// of this round + earlier lanes of this warp.
__global__ void __launch_bounds__(SORT_THREADS)
radix_scatter_kernel(const unsigned *keys, const unsigned *ids, unsigned *keys_out,
                     unsigned *ids_out, size_t n, int shift, const unsigned *hist, int nblocks,
                     const unsigned *totals) {
  __shared__ unsigned run[RADIX];                // global offset + count so far
  __shared__ unsigned wcnt[SORT_WARPS][RADIX];   // this round's per-warp counts
  __shared__ unsigned wsum[SORT_WARPS];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  static_assert(RADIX == SORT_THREADS, "one thread per digit in the base-offset scan");
  { // digit base = exclusive scan of the digit totals
    const unsigned v = totals[threadIdx.x];
    unsigned x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      unsigned y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) wsum[wid] = x;
    __syncthreads();
    unsigned before = 0;
    for (int w = 0; w < wid; w++) before += wsum[w];
    run[threadIdx.x] = before + (x - v) + hist[(size_t)threadIdx.x * nblocks + blockIdx.x];
  }
  const size_t base = (size_t)blockIdx.x * SORT_TILE;
  for (int r = 0; r < SORT_ROUNDS; r++) {
    for (int d = threadIdx.x; d < RADIX * SORT_WARPS; d += SORT_THREADS) (&wcnt[0][0])[d] = 0;
    __syncthreads();
    const size_t i = base + (size_t)r * SORT_THREADS + threadIdx.x;
    const bool valid = i < n;
    unsigned key = valid ? keys[i] : 0u;
    const unsigned dg = (key >> shift) & (RADIX - 1);
    // lanes of this warp with the same digit (invalid lanes match nobody valid)
    unsigned peers = __match_any_sync(0xffffffffu, valid ? dg : (RADIX + (unsigned)lane));
    const unsigned rank_in_warp = __popc(peers & ((1u << lane) - 1u));
    if (valid && rank_in_warp == 0) wcnt[wid][dg] = __popc(peers);
    __syncthreads();
    if (valid) {
      unsigned off = run[dg];
      for (int w = 0; w < wid; w++) off += wcnt[w][dg];
      off += rank_in_warp;
      keys_out[off] = key;
      ids_out[off] = ids[i];
    }
    __syncthreads();
    for (int d = threadIdx.x; d < RADIX; d += SORT_THREADS) {
      unsigned t = 0;
#pragma unroll
      for (int w = 0; w < SORT_WARPS; w++) t += wcnt[w][d];
      run[d] += t;
    }
    __syncthreads();
  }
}

// ordered segmented sum: the thread at the head of a run of equal keys walks
// the run and adds the values to the cell in record order.
__global__ void __launch_bounds__(256)
segsum_kernel(const unsigned *keys, const unsigned *ids, const float *vals, size_t nrec,
              float *grid, int nxyz, int nspec) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nrec) return;
  const unsigned key = keys[i];
  if (key == 0xffffffffu) return;
  if (i > 0 && keys[i - 1] == key) return; // not a run head
  const size_t inner = key % (unsigned)nxyz, rest = key / (unsigned)nxyz;
  for (int ks = 0; ks < nspec; ks++) {
    float *cell = grid + inner + (size_t)nxyz * (ks + (size_t)nspec * rest);
    float acc = *cell;
    const float *v = vals + (size_t)ks * nrec;
    for (size_t j = i; j < nrec && keys[j] == key; j++) acc = acc + v[ids[j]];
    *cell = acc;
  }
}

} // namespace

void scatter_free(ScatterWork &w) {
  for (int k = 0; k < 2; k++) {
    cudaFree(w.keys[k]);
    cudaFree(w.ids[k]);
    w.keys[k] = w.ids[k] = nullptr;
  }
  cudaFree(w.vals);
  cudaFree(w.hist);
  w.vals = nullptr;
  w.hist = nullptr;
  w.cap_rec = w.cap_vals = w.cap_hist = 0;
}

int scatter_reserve(ScatterWork &w, size_t nrec, int nspec) {
  if (nrec > w.cap_rec) {
    for (int k = 0; k < 2; k++) {
      cudaFree(w.keys[k]);
      cudaFree(w.ids[k]);
      SCK(cudaMalloc((void **)&w.keys[k], nrec * sizeof(unsigned)));
      SCK(cudaMalloc((void **)&w.ids[k], nrec * sizeof(unsigned)));
    }
    w.cap_rec = nrec;
  }
  if (nrec * nspec > w.cap_vals) {
    cudaFree(w.vals);
    SCK(cudaMalloc((void **)&w.vals, nrec * nspec * sizeof(float)));
    w.cap_vals = nrec * nspec;
  }
  size_t nblocks = (nrec + SORT_TILE - 1) / SORT_TILE;
  if ((nblocks + 1) * RADIX > w.cap_hist) { // per-block counts + one row of digit totals
    cudaFree(w.hist);
    SCK(cudaMalloc((void **)&w.hist, (nblocks + 1) * RADIX * sizeof(unsigned)));
    w.cap_hist = (nblocks + 1) * RADIX;
  }
  return 0;
}

int scatter_sort_pairs(ScatterWork &w, size_t n, int bits, cudaStream_t st, int64_t *launches,
                       int *out) {
  const int nblocks = (int)((n + SORT_TILE - 1) / SORT_TILE);
  int cur = 0;
  for (int shift = 0; shift < bits; shift += RADIX_BITS) {
    radix_hist_kernel<<<nblocks, SORT_THREADS, 0, st>>>(w.keys[cur], n, shift, w.hist, nblocks);
    digit_scan_kernel<<<RADIX, 256, 0, st>>>(w.hist, nblocks, w.hist + (size_t)nblocks * RADIX);
    radix_scatter_kernel<<<nblocks, SORT_THREADS, 0, st>>>(w.keys[cur], w.ids[cur], w.keys[cur ^ 1],
                                                           w.ids[cur ^ 1], n, shift, w.hist, nblocks,
                                                           w.hist + (size_t)nblocks * RADIX);
    if (launches) *launches += 3;
    cur ^= 1;
  }
  SCK(cudaGetLastError());
  *out = cur;
  return 0;
}

namespace {
__global__ void iota_ids_kernel(unsigned *ids, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) ids[i] = (unsigned)i;
}
// creceptor's c(ks) of one (receptor, species): the particles' contributions added in slot order
// (src/conccalc.f90:476-492) by one warp -- every lane ends with the same sequential sum
__global__ void __launch_bounds__(32) receptor_seqsum_kernel(const float *vals, int nslots, float *acc) {
  const float *v = vals + (size_t)blockIdx.x * nslots;
  const int lane = threadIdx.x;
  float sum = acc[blockIdx.x];
  for (int base = 0; base < nslots; base += 32) {
    const float x = (base + lane < nslots) ? v[base + lane] : 0.f;
    const int n = min(32, nslots - base);
    for (int k = 0; k < n; k++) sum = sum + __shfl_sync(0xffffffffu, x, k);
  }
  if (lane == 0) acc[blockIdx.x] = sum;
}
} // namespace

// records (keys[nrec], vals[nspec][nrec]) written by a kernel -> grid, every cell in record order
int scatter_records_deterministic(ScatterWork &w, const unsigned *keys, const float *vals, size_t nrec, int nspec,
                                  float *grid, int nxyz, unsigned long long ncell, cudaStream_t st, int64_t *launches) {
  if (nrec == 0) return 0;
  if (ncell >= 0xffffffffull || nrec >= 0xffffffffull) {
    s_err = "deterministic scatter: more than 2^32 cells or records";
    return 1;
  }
  if (scatter_reserve(w, nrec, 1)) return 1;
  int bits = 1;
  while ((1ull << bits) < ncell + 1) bits++;
  bits = ((bits + RADIX_BITS - 1) / RADIX_BITS) * RADIX_BITS;
  if (bits < 32) bits = (bits + RADIX_BITS <= 32) ? bits + RADIX_BITS : 32; // the all-ones keys sort last
  SCK(cudaMemcpyAsync(w.keys[0], keys, nrec * sizeof(unsigned), cudaMemcpyDeviceToDevice, st));
  iota_ids_kernel<<<(unsigned)((nrec + 255) / 256), 256, 0, st>>>(w.ids[0], nrec);
  int cur = 0;
  if (scatter_sort_pairs(w, nrec, bits, st, launches, &cur)) return 1;
  segsum_kernel<<<(unsigned)((nrec + 255) / 256), 256, 0, st>>>(w.keys[cur], w.ids[cur], vals, nrec, grid, nxyz, nspec);
  if (launches) *launches += 2;
  SCK(cudaGetLastError());
  return 0;
}

int scatter_receptor_ordered(const float *vals, int nsums, int nslots, float *acc, cudaStream_t st) {
  if (nsums <= 0 || nslots <= 0) return 0;
  receptor_seqsum_kernel<<<nsums, 32, 0, st>>>(vals, nslots, acc);
  SCK(cudaGetLastError());
  return 0;
}

int scatter_conccalc_deterministic(ScatterWork &w, const DevConcArgs &a, bool strict,
                                   cudaStream_t st, int64_t *launches) {
  const DevCfg &c = a.cfg;
  const size_t nrec = 4 * (size_t)c.numpart; // record id = 4*slot + corner; slots < numpart
  if (nrec == 0) return 0;
  if (nrec >= 0xffffffffull) {
    s_err = "deterministic scatter: more than 2^32 records";
    return 1;
  }
  if (scatter_reserve(w, nrec, c.nspec)) return 1;
  for (int nest = 0; nest <= (c.nested_output == 1 ? 1 : 0); nest++) {
    const int nxg = nest ? c.numxgridn : c.numxgrid, nyg = nest ? c.numygridn : c.numygrid;
    const int nxyz = nxg * nyg * c.numzgrid;
    const unsigned long long ncell =
        (unsigned long long)nxyz * c.maxpointspec_act * c.nclassunc * c.maxageclass;
    if (ncell >= 0xffffffffull) {
      s_err = "deterministic scatter: grid has more than 2^32 cells";
      return 1;
    }
    int bits = 1;
    while ((1ull << bits) < ncell + 1) bits++;
    // all-ones keys (= no record) must sort last: widen to cover them
    bits = ((bits + RADIX_BITS - 1) / RADIX_BITS) * RADIX_BITS;
    if (bits < 32) bits = (bits + RADIX_BITS <= 32) ? bits + RADIX_BITS : 32;
    iota_fill_kernel<<<(unsigned)((nrec + 255) / 256), 256, 0, st>>>(w.keys[0], w.ids[0], nrec);
    if (strict) fpbk_conc_emit_strict(a, nest, w.keys[0], w.vals, nrec, st);
    else fpbk_conc_emit_fast(a, nest, w.keys[0], w.vals, nrec, st);
    if (launches) *launches += 2;
    int cur = 0;
    if (scatter_sort_pairs(w, nrec, bits, st, launches, &cur)) return 1;
    segsum_kernel<<<(unsigned)((nrec + 255) / 256), 256, 0, st>>>(
        w.keys[cur], w.ids[cur], w.vals, nrec, nest ? a.griduncn : a.gridunc, nxyz, c.nspec);
    if (launches) *launches += 1;
  }
  SCK(cudaGetLastError());
  return 0;
}
