// fpb_sort.cu -- see fpb_sort.cuh.  Integer / data-movement kernels only.
#include "fpb_sort.cuh"

namespace {

__global__ void __launch_bounds__(256)
build_keys_kernel(DevCfg c, DevParticles p, const float *height, int nrows, unsigned *keys,
                  unsigned *ids, unsigned *nlive) {
  __shared__ float sh[FPB_MAXNZ];
  for (int i = threadIdx.x; i < c.nz; i += blockDim.x) sh[i] = height[i];
  __syncthreads();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned live = 0;
  if (i < nrows) {
    unsigned key = 0xffffffffu;
    if (p.itra1[i] != FPB_ITRA_DEAD) {
      live = 1;
      int ix = (int)p.xtra1[i], jy = (int)p.ytra1[i];
      ix = min(max(ix, 0), c.nxd - 1);
      jy = min(max(jy, 0), c.nyd - 1);
      const float zt = p.ztra1[i];
      int lo = 2, hi = c.nz; // level search of src/interpol_all.f90:118-125
      while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (sh[mid - 1] > zt) hi = mid; else lo = mid + 1;
      }
      key = (unsigned)(((lo - 2) * c.nyd + jy) * c.nxd + ix);
    }
    keys[i] = key;
    ids[i] = (unsigned)i;
  }
  unsigned tot = __reduce_add_sync(0xffffffffu, live);
  if ((threadIdx.x & 31) == 0 && tot) atomicAdd(nlive, tot);
}

__global__ void __launch_bounds__(256)
permute_kernel(DevParticles s, DevParticles d, const unsigned *ids, int nrows, int nspec) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nrows) return;
  const unsigned j = ids[i];
  d.xtra1[i] = s.xtra1[j]; d.ytra1[i] = s.ytra1[j]; d.ztra1[i] = s.ztra1[j];
  d.itra1[i] = s.itra1[j]; d.npoint[i] = s.npoint[j]; d.nclass[i] = s.nclass[j];
  d.idt[i] = s.idt[j]; d.itramem[i] = s.itramem[j]; d.itrasplit[i] = s.itrasplit[j];
  d.uap[i] = s.uap[j]; d.ucp[i] = s.ucp[j]; d.uzp[i] = s.uzp[j];
  d.us[i] = s.us[j]; d.vs[i] = s.vs[j]; d.ws[i] = s.ws[j];
  d.cbt[i] = s.cbt[j];
  d.slot[i] = s.slot[j];
  for (int k = 0; k < nspec; k++) {
    d.xmass1[(size_t)k * d.maxpart + i] = s.xmass1[(size_t)k * s.maxpart + j];
    d.xscav_frac1[(size_t)k * d.maxpart + i] = s.xscav_frac1[(size_t)k * s.maxpart + j];
  }
}

__global__ void invert_kernel(const int32_t *slot, int32_t *row_of_slot, int nrows) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nrows) row_of_slot[slot[i]] = i;
}

__global__ void iota_kernel(int32_t *a, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) a[i] = i;
}

// staging row (first + t) <- device row row_of_slot[first + t]
__global__ void __launch_bounds__(256)
gather_kernel(DevParticles s, DevParticles d, const int32_t *row_of_slot, int first, int count,
              int nspec) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= count) return;
  const int i = first + t, j = row_of_slot[i];
  d.xtra1[i] = s.xtra1[j]; d.ytra1[i] = s.ytra1[j]; d.ztra1[i] = s.ztra1[j];
  d.itra1[i] = s.itra1[j]; d.npoint[i] = s.npoint[j]; d.nclass[i] = s.nclass[j];
  d.idt[i] = s.idt[j]; d.itramem[i] = s.itramem[j]; d.itrasplit[i] = s.itrasplit[j];
  d.uap[i] = s.uap[j]; d.ucp[i] = s.ucp[j]; d.uzp[i] = s.uzp[j];
  d.us[i] = s.us[j]; d.vs[i] = s.vs[j]; d.ws[i] = s.ws[j];
  d.cbt[i] = s.cbt[j];
  for (int k = 0; k < nspec; k++) {
    d.xmass1[(size_t)k * d.maxpart + i] = s.xmass1[(size_t)k * s.maxpart + j];
    d.xscav_frac1[(size_t)k * d.maxpart + i] = s.xscav_frac1[(size_t)k * s.maxpart + j];
  }
}

__global__ void __launch_bounds__(256)
scatter_kernel(DevParticles s, DevParticles d, const int32_t *row_of_slot, int first, int count,
               int nspec, bool have_split, bool have_scav) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= count) return;
  const int i = first + t, j = row_of_slot[i];
  d.xtra1[j] = s.xtra1[i]; d.ytra1[j] = s.ytra1[i]; d.ztra1[j] = s.ztra1[i];
  d.itra1[j] = s.itra1[i]; d.npoint[j] = s.npoint[i]; d.nclass[j] = s.nclass[i];
  d.idt[j] = s.idt[i]; d.itramem[j] = s.itramem[i];
  if (have_split) d.itrasplit[j] = s.itrasplit[i];
  d.uap[j] = s.uap[i]; d.ucp[j] = s.ucp[i]; d.uzp[j] = s.uzp[i];
  d.us[j] = s.us[i]; d.vs[j] = s.vs[i]; d.ws[j] = s.ws[i];
  d.cbt[j] = s.cbt[i];
  for (int k = 0; k < nspec; k++) {
    d.xmass1[(size_t)k * d.maxpart + j] = s.xmass1[(size_t)k * s.maxpart + i];
    if (have_scav) d.xscav_frac1[(size_t)k * d.maxpart + j] = s.xscav_frac1[(size_t)k * s.maxpart + i];
  }
}

inline unsigned nb(int n, int t) { return (unsigned)((n + t - 1) / t); }
} // namespace

void sortk_build_keys(const DevCfg &c, const DevParticles &p, const float *height, int nrows,
                      unsigned *keys, unsigned *ids, unsigned *d_nlive, cudaStream_t st) {
  cudaMemsetAsync(d_nlive, 0, sizeof(unsigned), st);
  build_keys_kernel<<<nb(nrows, 256), 256, 0, st>>>(c, p, height, nrows, keys, ids, d_nlive);
}
void sortk_permute(const DevParticles &src, const DevParticles &dst, const unsigned *ids,
                   int nrows, int nspec, cudaStream_t st) {
  permute_kernel<<<nb(nrows, 256), 256, 0, st>>>(src, dst, ids, nrows, nspec);
}
void sortk_invert(const int32_t *slot, int32_t *row_of_slot, int nrows, cudaStream_t st) {
  invert_kernel<<<nb(nrows, 256), 256, 0, st>>>(slot, row_of_slot, nrows);
}
void sortk_iota(int32_t *a, int n, cudaStream_t st) {
  iota_kernel<<<nb(n, 256), 256, 0, st>>>(a, n);
}
void sortk_gather_to_staging(const DevParticles &rows, const DevParticles &stg,
                             const int32_t *row_of_slot, int first, int count, int nspec,
                             cudaStream_t st) {
  gather_kernel<<<nb(count, 256), 256, 0, st>>>(rows, stg, row_of_slot, first, count, nspec);
}
void sortk_scatter_from_staging(const DevParticles &stg, const DevParticles &rows,
                                const int32_t *row_of_slot, int first, int count, int nspec,
                                bool have_split, bool have_scav, cudaStream_t st) {
  scatter_kernel<<<nb(count, 256), 256, 0, st>>>(stg, rows, row_of_slot, first, count, nspec,
                                                  have_split, have_scav);
}
