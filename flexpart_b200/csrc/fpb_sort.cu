// fpb_sort.cu -- see fpb_sort.cuh.  Integer / data-movement kernels only.
#include "fpb_sort.cuh"

namespace {

// Turbulence regime of the coming advance() call, as the top key bits: rows of
// one regime become contiguous, so the lanes of a warp of fpb_pbl_kernel take
// the same hanna*() branch.  0: h/|ol| < 1, 1: ol < 0, 2: stable, 3: above the
// PBL (no sub-steps).  A grouping heuristic only -- results do not depend on
// the row order -- so the interpolation here is the plain formula.
struct RegimeMet {
  const float4 *S0, *S1; // {hmix, ustar, wstar, oli} of memind(1), memind(2); null: no regime bits
  float dt1, dt2;
};

// key = (level, jy >> sy, ix >> sx) packed level-major into lb + yb + xb bits: the cell is
// coarsened to a tile when the exact cell index would push the sort to a 4th radix pass
struct KeyLayout { int sx, sy, xb, yb, cell_bits; };

__global__ void __launch_bounds__(256)
build_keys_kernel(DevCfg c, DevParticles p, const float *height, int nrows, unsigned *keys,
                  unsigned *ids, unsigned *nlive, RegimeMet rm, KeyLayout kl) {
  __shared__ float sh[FPB_MAXNZ];
  for (int i = threadIdx.x; i < c.nz; i += blockDim.x) sh[i] = height[i];
  __syncthreads();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned live = 0;
  if (i < nrows) {
    unsigned key = 0xffffffffu;
    if (p.itra1[i] != FPB_ITRA_DEAD) {
      live = 1;
      const double xd = p.xtra1[i], yd = p.ytra1[i];
      int ix = (int)xd, jy = (int)yd;
      ix = min(max(ix, 0), c.nxd - 1);
      jy = min(max(jy, 0), c.nyd - 1);
      const float zt = p.ztra1[i];
      int lo = 2, hi = c.nz; // level search of src/interpol_all.f90:118-125
      while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (sh[mid - 1] > zt) hi = mid; else lo = mid + 1;
      }
      key = ((((unsigned)(lo - 2) << kl.yb) | (unsigned)(jy >> kl.sy)) << kl.xb) | (unsigned)(ix >> kl.sx);
      if (rm.S0) {
        const int ixp = min(ix + 1, c.nxd - 1), jyp = min(jy + 1, c.nyd - 1);
        const float ddx = (float)xd - (float)ix, ddy = (float)yd - (float)jy;
        const float p1 = (1.f - ddx) * (1.f - ddy), p2 = ddx * (1.f - ddy), p3 = (1.f - ddx) * ddy, p4 = ddx * ddy;
        const int o00 = ix + c.nxd * jy, o10 = ixp + c.nxd * jy, o01 = ix + c.nxd * jyp, o11 = ixp + c.nxd * jyp;
        const float4 a0 = __ldg(rm.S0 + o00), b0 = __ldg(rm.S0 + o10), c0 = __ldg(rm.S0 + o01), d0 = __ldg(rm.S0 + o11);
        const float4 a1 = __ldg(rm.S1 + o00), b1 = __ldg(rm.S1 + o10), c1 = __ldg(rm.S1 + o01), d1 = __ldg(rm.S1 + o11);
        const float h = fmaxf(fmaxf(fmaxf(a0.x, b0.x), fmaxf(c0.x, d0.x)), fmaxf(fmaxf(a1.x, b1.x), fmaxf(c1.x, d1.x)));
        const float oli = ((p1 * a0.w + p2 * b0.w + p3 * c0.w + p4 * d0.w) * rm.dt2 +
                           (p1 * a1.w + p2 * b1.w + p3 * c1.w + p4 * d1.w) * rm.dt1);
        unsigned regime;
        if (!(zt <= h)) regime = 3u;
        else if (h * fabsf(oli) < fabsf(rm.dt1 + rm.dt2)) regime = 0u;
        else regime = ((oli < 0.f) != ((rm.dt1 + rm.dt2) < 0.f)) ? 1u : 2u;
        key |= regime << kl.cell_bits;
      }
    }
    keys[i] = key;
    ids[i] = (unsigned)i;
  }
  unsigned tot = __reduce_add_sync(0xffffffffu, live);
  if ((threadIdx.x & 31) == 0 && tot) atomicAdd(nlive, tot);
}

__global__ void __launch_bounds__(256)
permute_kernel(DevParticles s, DevParticles d, const unsigned *ids, int nrows, int nspec,
               int32_t *row_of_slot, int base) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nrows) return;
  const unsigned j = ids[i];
  d.xtra1[i] = s.xtra1[j]; d.ytra1[i] = s.ytra1[j]; d.ztra1[i] = s.ztra1[j];
  d.itra1[i] = s.itra1[j]; d.npoint[i] = s.npoint[j]; d.nclass[i] = s.nclass[j];
  d.idt[i] = s.idt[j]; d.itramem[i] = s.itramem[j]; d.itrasplit[i] = s.itrasplit[j];
  d.uap[i] = s.uap[j]; d.ucp[i] = s.ucp[j]; d.uzp[i] = s.uzp[j];
  d.us[i] = s.us[j]; d.vs[i] = s.vs[j]; d.ws[i] = s.ws[j];
  d.cbt[i] = s.cbt[j];
  const int32_t sl = s.slot[j];
  d.slot[i] = sl;
  row_of_slot[sl] = base + i; // the inverse map, kept in the same pass
  for (int k = 0; k < nspec; k++) {
    d.xmass1[(size_t)k * d.maxpart + i] = s.xmass1[(size_t)k * s.maxpart + j];
    d.xscav_frac1[(size_t)k * d.maxpart + i] = s.xscav_frac1[(size_t)k * s.maxpart + j];
  }
}

// ---- packed permute: a random-row gather of 18 separate 2/4/8-byte arrays fetches a 32-byte sector per
// value (8x the useful bytes for the 4-byte fields).  For large row counts the rows are first packed
// into array-of-structures records (coalesced reads, records written back to back), then every
// destination row reads ITS source record -- 3 to 4 whole sectors -- and writes the arrays coalesced.
// Record: [xtra1 ytra1 | ztra1 itra1 npoint nclass] [idt itramem itrasplit uap ucp uzp us vs]
//         [ws slot cbt+pad xmass1(1:nspec) xscav_frac1(1:nspec)], padded to 16 bytes.
template <int NSPEC> struct PackedRow {
  static constexpr int WORDS = ((76 + 8 * NSPEC + 15) / 16) * 4;
};

template <int NSPEC>
__global__ void __launch_bounds__(256) pack_rows_kernel(DevParticles s, uint4 *rec, int nrows) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nrows) return;
  constexpr int W = PackedRow<NSPEC>::WORDS;
  unsigned w[W];
#pragma unroll
  for (int k = 0; k < W; k++) w[k] = 0u;
  const unsigned long long x = (unsigned long long)__double_as_longlong(s.xtra1[i]);
  const unsigned long long y = (unsigned long long)__double_as_longlong(s.ytra1[i]);
  w[0] = (unsigned)x; w[1] = (unsigned)(x >> 32); w[2] = (unsigned)y; w[3] = (unsigned)(y >> 32);
  w[4] = __float_as_uint(s.ztra1[i]); w[5] = (unsigned)s.itra1[i]; w[6] = (unsigned)s.npoint[i]; w[7] = (unsigned)s.nclass[i];
  w[8] = (unsigned)s.idt[i]; w[9] = (unsigned)s.itramem[i]; w[10] = (unsigned)s.itrasplit[i];
  w[11] = __float_as_uint(s.uap[i]); w[12] = __float_as_uint(s.ucp[i]); w[13] = __float_as_uint(s.uzp[i]);
  w[14] = __float_as_uint(s.us[i]); w[15] = __float_as_uint(s.vs[i]);
  w[16] = __float_as_uint(s.ws[i]); w[17] = (unsigned)s.slot[i]; w[18] = (unsigned)(unsigned short)s.cbt[i];
#pragma unroll
  for (int k = 0; k < NSPEC; k++) {
    w[19 + k] = __float_as_uint(s.xmass1[(size_t)k * s.maxpart + i]);
    w[19 + NSPEC + k] = __float_as_uint(s.xscav_frac1[(size_t)k * s.maxpart + i]);
  }
  uint4 *r = rec + (size_t)i * (W / 4);
#pragma unroll
  for (int k = 0; k < W / 4; k++) r[k] = make_uint4(w[4 * k], w[4 * k + 1], w[4 * k + 2], w[4 * k + 3]);
}

template <int NSPEC>
__global__ void __launch_bounds__(256) unpack_rows_kernel(const uint4 *rec, DevParticles d, const unsigned *ids, int nrows,
                                                          int32_t *row_of_slot, int base) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nrows) return;
  constexpr int W = PackedRow<NSPEC>::WORDS;
  const uint4 *r = rec + (size_t)ids[i] * (W / 4);
  unsigned w[W];
#pragma unroll
  for (int k = 0; k < W / 4; k++) {
    const uint4 v = __ldg(r + k);
    w[4 * k] = v.x; w[4 * k + 1] = v.y; w[4 * k + 2] = v.z; w[4 * k + 3] = v.w;
  }
  d.xtra1[i] = __longlong_as_double((long long)((unsigned long long)w[0] | ((unsigned long long)w[1] << 32)));
  d.ytra1[i] = __longlong_as_double((long long)((unsigned long long)w[2] | ((unsigned long long)w[3] << 32)));
  d.ztra1[i] = __uint_as_float(w[4]); d.itra1[i] = (int)w[5]; d.npoint[i] = (int)w[6]; d.nclass[i] = (int)w[7];
  d.idt[i] = (int)w[8]; d.itramem[i] = (int)w[9]; d.itrasplit[i] = (int)w[10];
  d.uap[i] = __uint_as_float(w[11]); d.ucp[i] = __uint_as_float(w[12]); d.uzp[i] = __uint_as_float(w[13]);
  d.us[i] = __uint_as_float(w[14]); d.vs[i] = __uint_as_float(w[15]); d.ws[i] = __uint_as_float(w[16]);
  const int32_t sl = (int32_t)w[17];
  d.slot[i] = sl;
  d.cbt[i] = (int16_t)(unsigned short)w[18];
  row_of_slot[sl] = base + i;
#pragma unroll
  for (int k = 0; k < NSPEC; k++) {
    d.xmass1[(size_t)k * d.maxpart + i] = __uint_as_float(w[19 + k]);
    d.xscav_frac1[(size_t)k * d.maxpart + i] = __uint_as_float(w[19 + NSPEC + k]);
  }
}

__global__ void invert_kernel(const int32_t *slot, int32_t *row_of_slot, int nrows, int base) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nrows) row_of_slot[slot[i]] = base + i;
}

// staging row slot[i] <- device row i, only the arrays the particle loop writes
// (src/timemanager.f90:531-712: position, itra1, idt, turbulent velocities, cbt, masses)
__global__ void __launch_bounds__(256)
scatter_back_kernel(DevParticles rows, DevParticles stg, int count, int nspec, bool scav) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const int s = rows.slot[i];
  stg.xtra1[s] = rows.xtra1[i]; stg.ytra1[s] = rows.ytra1[i]; stg.ztra1[s] = rows.ztra1[i];
  stg.itra1[s] = rows.itra1[i]; stg.idt[s] = rows.idt[i];
  stg.uap[s] = rows.uap[i]; stg.ucp[s] = rows.ucp[i]; stg.uzp[s] = rows.uzp[i];
  stg.us[s] = rows.us[i]; stg.vs[s] = rows.vs[i]; stg.ws[s] = rows.ws[i];
  stg.cbt[s] = rows.cbt[i];
  for (int k = 0; k < nspec; k++) {
    stg.xmass1[(size_t)k * stg.maxpart + s] = rows.xmass1[(size_t)k * rows.maxpart + i];
    if (scav) stg.xscav_frac1[(size_t)k * stg.maxpart + s] = rows.xscav_frac1[(size_t)k * rows.maxpart + i];
  }
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

static int bits_for(int n) { // bits needed for values 0..n-1
  int b = 1;
  while ((1 << b) < n) b++;
  return b;
}

int sortk_key_bits(const DevCfg &c, bool regime) {
  // level bits + y bits + x bits (+2 regime) + 1 (dead rows = all ones, must sort last) <= 24 -> 3 passes
  return sortk_key_layout_bits(c) + (regime ? 2 : 0) + 1;
}

static KeyLayout key_layout(const DevCfg &c) {
  KeyLayout k;
  const int lb = bits_for(c.nz > 1 ? c.nz - 1 : 1);
  k.xb = bits_for(c.nxd); k.yb = bits_for(c.nyd); k.sx = k.sy = 0;
  // method 1 (sub-stepping, sorted every step): tiles, 3 passes; method 0 (gather-bound, sorted
  // rarely): exact cells, the locality is worth the 4th pass
  while (c.method == 1 && lb + k.xb + k.yb > 21 && (k.xb > 1 || k.yb > 1)) { // drop low bits, x and y in turn
    if (k.xb >= k.yb && k.xb > 1) { k.xb--; k.sx++; } else { k.yb--; k.sy++; }
  }
  k.cell_bits = lb + k.xb + k.yb;
  return k;
}

int sortk_key_layout_bits(const DevCfg &c) { return key_layout(c).cell_bits; }

void sortk_build_keys(const DevCfg &c, const DevParticles &p, const float *height, int nrows,
                      unsigned *keys, unsigned *ids, unsigned *d_nlive, cudaStream_t st,
                      const DevMetSlot *met) {
  cudaMemsetAsync(d_nlive, 0, sizeof(unsigned), st);
  RegimeMet rm;
  rm.S0 = met ? met[0].S : nullptr;
  rm.S1 = met ? met[1].S : nullptr;
  rm.dt1 = (float)(c.itime - c.memtime[0]);
  rm.dt2 = (float)(c.memtime[1] - c.itime);
  build_keys_kernel<<<nb(nrows, 256), 256, 0, st>>>(c, p, height, nrows, keys, ids, d_nlive, rm, key_layout(c));
}
void sortk_permute(const DevParticles &src, const DevParticles &dst, const unsigned *ids,
                   int nrows, int nspec, cudaStream_t st, int32_t *row_of_slot, int base) {
  permute_kernel<<<nb(nrows, 256), 256, 0, st>>>(src, dst, ids, nrows, nspec, row_of_slot, base);
}
size_t sortk_packed_bytes(int nrows, int nspec) {
  if (nspec < 1 || nspec > 4) return 0; // (more species: the plain gather)
  return (size_t)nrows * (((76 + 8 * nspec + 15) / 16) * 16);
}
void sortk_permute_packed(const DevParticles &src, const DevParticles &dst, const unsigned *ids, int nrows, int nspec,
                          cudaStream_t st, int32_t *row_of_slot, void *records, int base) {
  uint4 *rec = reinterpret_cast<uint4 *>(records);
  const unsigned g = nb(nrows, 256);
  switch (nspec) {
    case 1: pack_rows_kernel<1><<<g, 256, 0, st>>>(src, rec, nrows); unpack_rows_kernel<1><<<g, 256, 0, st>>>(rec, dst, ids, nrows, row_of_slot, base); break;
    case 2: pack_rows_kernel<2><<<g, 256, 0, st>>>(src, rec, nrows); unpack_rows_kernel<2><<<g, 256, 0, st>>>(rec, dst, ids, nrows, row_of_slot, base); break;
    case 3: pack_rows_kernel<3><<<g, 256, 0, st>>>(src, rec, nrows); unpack_rows_kernel<3><<<g, 256, 0, st>>>(rec, dst, ids, nrows, row_of_slot, base); break;
    default: pack_rows_kernel<4><<<g, 256, 0, st>>>(src, rec, nrows); unpack_rows_kernel<4><<<g, 256, 0, st>>>(rec, dst, ids, nrows, row_of_slot, base); break;
  }
}
void sortk_invert(const int32_t *slot, int32_t *row_of_slot, int nrows, cudaStream_t st, int base) {
  invert_kernel<<<nb(nrows, 256), 256, 0, st>>>(slot, row_of_slot, nrows, base);
}
void sortk_scatter_back(const DevParticles &rows, const DevParticles &stg, int count, int nspec,
                        cudaStream_t st, bool scav) {
  scatter_back_kernel<<<nb(count, 256), 256, 0, st>>>(rows, stg, count, nspec, scav);
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
