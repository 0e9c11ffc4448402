// fpb_domainfill.cu -- init_domainfill (src/init_domainfill.f90:55-283, MDOMAINFILL = 1) as kernels.
//
// The reference walks the columns of the domain box (jy outer, ix inner) and appends each column's
// particles to the particle arrays; the number per column follows the column's share of the air
// mass.  Here:
//   df_colmass_kernel   one thread per column: colmass = (p(1) - p(nz)) / g * gridarea(jy)   (:135-143)
//   df_total_kernel     colmasstotal: ONE warp adds the columns in the reference's order (a float sum
//                       is order dependent; 2.6e5 adds, once per run)
//   df_count_kernel     ncolumn = nint(0.999 * npart(1) * colmass / colmasstotal)             (:158-159)
//   df_scan_*           exclusive scan of ncolumn = the first global particle index of each column
//                       (the reference's running numpart)
//   df_fill_kernel      one warp per column: pressure profile of the column in shared memory, lane
//                       l takes particles l, l+32, ...: pressure-equidistant (or, for thin columns,
//                       random) height, random horizontal position inside the cell, class, mass   (:166-260)
// Compiled with --fmad=false and explicit _rn operations where the order matters: with the
// reference's ran1 stream injected (FPB_RNG_REFERENCE) the particles are bit-identical to the
// sequential routine; the production modes draw from the particle's Philox counter stream.
#include "fpb_domainfill.cuh"

namespace {

constexpr int DF_BLOCK = 1024;
constexpr float R_AIR = 287.05f, GA = 9.81f; // par_mod

__device__ __forceinline__ void col_ij(const DomainfillArgs &a, int col, int &ix, int &jy) {
  jy = a.ny0 + col / a.ncolx;
  ix = a.nx0 + col % a.ncolx;
}

__device__ __forceinline__ float level_pressure(const DomainfillArgs &a, int ix, int jy, int kz /*1-based*/) {
  const size_t o = (size_t)(kz - 1) * a.cfg.nxd * a.cfg.nyd + (size_t)jy * a.cfg.nxd + ix;
  return __fmul_rn(__fmul_rn(__ldg(&a.A1[o]).w, R_AIR), __ldg(a.T1 + o)); // rho*r_air*tt
}

__global__ void __launch_bounds__(256) df_colmass_kernel(const DomainfillArgs a) {
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= a.ncols) return;
  int ix, jy;
  col_ij(a, col, ix, jy);
  const float p1 = level_pressure(a, ix, jy, 1), pn = level_pressure(a, ix, jy, a.cfg.nz);
  a.colmass[col] = __fmul_rn(__fdiv_rn(__fsub_rn(p1, pn), GA), a.gridarea[jy]);
}

// every lane ends with the same sequential sum
__global__ void __launch_bounds__(32) df_total_kernel(const DomainfillArgs a) {
  const int lane = threadIdx.x;
  float sum = 0.f;
  for (int base = 0; base < a.ncols; base += 32) {
    const float v = (base + lane < a.ncols) ? a.colmass[base + lane] : 0.f;
    const int n = min(32, a.ncols - base);
    for (int k = 0; k < n; k++) sum = __fadd_rn(sum, __shfl_sync(0xffffffffu, v, k));
  }
  if (lane == 0) *a.total = sum;
}

__global__ void __launch_bounds__(DF_BLOCK) df_count_kernel(const DomainfillArgs a) {
  __shared__ int warp_sum[32], warp_max[32];
  const int col = blockIdx.x * DF_BLOCK + threadIdx.x;
  int n = 0;
  if (col < a.ncols) {
    const float x = __fdiv_rn(__fmul_rn(__fmul_rn(0.999f, a.npart1), a.colmass[col]), *a.total);
    n = (int)roundf(x); // nint()
    a.ncolumn[col] = n;
  }
  int s = n, m = n;
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, d);
    m = max(m, __shfl_xor_sync(0xffffffffu, m, d));
  }
  if ((threadIdx.x & 31) == 0) { warp_sum[threadIdx.x >> 5] = s; warp_max[threadIdx.x >> 5] = m; }
  __syncthreads();
  if (threadIdx.x < 32) {
    s = warp_sum[threadIdx.x]; m = warp_max[threadIdx.x];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      s += __shfl_xor_sync(0xffffffffu, s, d);
      m = max(m, __shfl_xor_sync(0xffffffffu, m, d));
    }
    if (threadIdx.x == 0) {
      a.block_sums[blockIdx.x] = (unsigned)s;
      atomicAdd(a.out + 0, s);
      atomicMax(a.out + 1, m);
    }
  }
}

// exclusive scan of up to DF_BLOCK block sums by one block (ncols <= 1024 * 1024)
__global__ void __launch_bounds__(DF_BLOCK) df_scan_blocks_kernel(unsigned *v, int n) {
  __shared__ unsigned wtot[32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const unsigned x = (threadIdx.x < n) ? v[threadIdx.x] : 0u;
  unsigned inc = x;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const unsigned t = __shfl_up_sync(0xffffffffu, inc, d);
    if (lane >= d) inc += t;
  }
  if (lane == 31) wtot[w] = inc;
  __syncthreads();
  if (w == 0) {
    unsigned t = wtot[lane], ti = t;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const unsigned u = __shfl_up_sync(0xffffffffu, ti, d);
      if (lane >= d) ti += u;
    }
    wtot[lane] = ti - t;
  }
  __syncthreads();
  if (threadIdx.x < n) v[threadIdx.x] = wtot[w] + inc - x;
}

__global__ void __launch_bounds__(DF_BLOCK) df_colstart_kernel(const DomainfillArgs a) {
  __shared__ unsigned wtot[32];
  const int col = blockIdx.x * DF_BLOCK + threadIdx.x;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const unsigned x = (col < a.ncols) ? (unsigned)a.ncolumn[col] : 0u;
  unsigned inc = x;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const unsigned t = __shfl_up_sync(0xffffffffu, inc, d);
    if (lane >= d) inc += t;
  }
  if (lane == 31) wtot[w] = inc;
  __syncthreads();
  if (w == 0) {
    unsigned t = wtot[lane], ti = t;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const unsigned u = __shfl_up_sync(0xffffffffu, ti, d);
      if (lane >= d) ti += u;
    }
    wtot[lane] = ti - t;
  }
  __syncthreads();
  if (col < a.ncols) a.colstart[col] = a.block_sums[blockIdx.x] + wtot[w] + inc - x;
}

constexpr int FILL_WARPS = 8;

__global__ void __launch_bounds__(32 * FILL_WARPS) df_fill_kernel(const DomainfillArgs a) {
  __shared__ float pp_s[FILL_WARPS][FPB_MAXNZ + 1];
  __shared__ float hh[FPB_MAXNZ + 1];
  const DevCfg &c = a.cfg;
  const int nz = c.nz;
  for (int i = threadIdx.x; i < nz; i += blockDim.x) hh[i + 1] = a.height[i];
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  float *pp = pp_s[w]; // 1-based like the reference's pp(kz)
  const uint2 key = make_uint2((uint32_t)c.seed, (uint32_t)(c.seed >> 32));
  int live_max = 0;
  unsigned odd_columns = 0;

  for (int col = blockIdx.x * FILL_WARPS + w; col < a.ncols; col += gridDim.x * FILL_WARPS) {
    const int ncolumn = a.ncolumn[col];
    if (ncolumn == 0) continue; // warp-uniform
    int ix, jy;
    col_ij(a, col, ix, jy);
    __syncwarp();
    for (int kz = lane + 1; kz <= nz; kz += 32) pp[kz] = level_pressure(a, ix, jy, kz);
    __syncwarp();
    const float p1 = pp[1], pn = pp[nz];
    const float cm = a.colmass[col];
    const float mass = __fdiv_rn(cm, (float)ncolumn);
    const float deltacol = __fdiv_rn(__fsub_rn(p1, pn), (float)ncolumn);
    const unsigned g0 = a.colstart[col];
    const bool thin = !(ncolumn > 20);
    // draws per particle in the reference's ran1 order: [pnew] x [x at ix = 0] [x at ix = nxmin1] y class
    const int dpp = (thin ? 1 : 0) + 3 + (ix == 0 ? 1 : 0) + (ix == c.nxmin1 ? 1 : 0);
    float pbase = __fadd_rn(p1, __fdiv_rn(deltacol, 2.f)); // pnew before particle `chunk + 1`

    for (int chunk = 0; chunk < ncolumn; chunk += 32) {
      const int j = chunk + lane + 1; // 1-based particle of the column
      // pnew = pnew - deltacol, repeated: the reference's running subtraction, rounding included
      float pnew = pbase;
      for (int k = 0; k <= lane; k++) pnew = __fsub_rn(pnew, deltacol);
      pbase = __shfl_sync(0xffffffffu, pnew, 31);
      if (j > ncolumn) continue;
      const unsigned g = g0 + (unsigned)(j - 1); // global particle index (the reference's numpart + jj - 1)
      if ((int)(g % (unsigned)a.id_stride) != a.id_offset) continue;
      const int slot = (int)(g / (unsigned)a.id_stride);
      if (slot >= a.p.maxpart) continue; // (the host has checked the total)

      float u_p, u_x, u_x0, u_xn, u_y, u_c;
      if (a.uniforms) {
        const float *u = a.uniforms + a.u_off[col] + (size_t)(j - 1) * dpp;
        int k = 0;
        u_p = thin ? u[k++] : 0.f;
        u_x = u[k++];
        u_x0 = (ix == 0) ? u[k++] : 0.f;
        u_xn = (ix == c.nxmin1) ? u[k++] : 0.f;
        u_y = u[k++];
        u_c = u[k++];
      } else {
        const uint32_t pid = (uint32_t)g; // global id: the particles do not depend on the GPU count
        const uint4 r = philox4x32_10(make_uint4(pid, 0u, 32u, 0u), key);
        u_x = u01(r.x); u_y = u01(r.y); u_c = u01(r.z); u_p = u01(r.w);
        u_x0 = u_xn = 0.f;
        if (ix == 0 || ix == c.nxmin1) {
          const uint4 q = philox4x32_10(make_uint4(pid, 0u, 33u, 0u), key);
          u_x0 = u01(q.x); u_xn = u01(q.y);
        }
      }
      if (thin) pnew = __fsub_rn(p1, __fmul_rn(u_p, __fsub_rn(p1, pn)));

      int matches = 0;
      float z = 0.f;
      for (int kz = 1; kz <= nz - 1; kz++) {
        if ((pp[kz] >= pnew) && (pp[kz + 1] < pnew)) {
          const float dz1 = __fsub_rn(pp[kz], pnew), dz2 = __fsub_rn(pnew, pp[kz + 1]);
          const float dz = __fdiv_rn(1.f, __fadd_rn(dz1, dz2));
          z = __fmul_rn(__fadd_rn(__fmul_rn(hh[kz], dz2), __fmul_rn(hh[kz + 1], dz1)), dz);
          matches++;
        }
      }
      if (matches != 1) odd_columns++;
      if (matches == 0) continue; // no layer brackets pnew: the reference leaves the slot untouched
      if (z > hh[nz] - 0.5f) z = hh[nz] - 0.5f;
      double x = (double)__fadd_rn(__fsub_rn((float)ix, 0.5f), u_x);
      if (ix == 0) x = (double)u_x0;
      if (ix == c.nxmin1) x = (double)__fsub_rn((float)c.nxmin1, u_xn);
      const double y = (double)__fadd_rn(__fsub_rn((float)jy, 0.5f), u_y);
      const DevParticles &p = a.p;
      p.xtra1[slot] = x;
      p.ytra1[slot] = y;
      p.ztra1[slot] = z;
      const int nc = (int)__fmul_rn(u_c, (float)c.nclassunc) + 1;
      p.nclass[slot] = min(nc, c.nclassunc);
      p.npoint[slot] = (int)g + 1; // numparticlecount
      p.idt[slot] = c.mintime;
      p.itramem[slot] = 0;
      p.itrasplit[slot] = c.ldirect * a.itsplit;
      p.xmass1[slot] = mass;
      for (int ks = 1; ks < c.nspec; ks++) p.xmass1[(size_t)ks * p.maxpart + slot] = 0.f;
      p.uap[slot] = 0.f; p.ucp[slot] = 0.f; p.uzp[slot] = 0.f;
      p.us[slot] = 0.f; p.vs[slot] = 0.f; p.ws[slot] = 0.f;
      p.cbt[slot] = 1;
      p.slot[slot] = slot;
      // :266-271: particles outside the domain are terminated at once
      const bool inside = !((x < 0.) || (x >= (float)c.nxmin1) || (y < 0.) || (y >= (float)c.nymin1));
      p.itra1[slot] = inside ? 0 : FPB_ITRA_DEAD;
      if (inside) live_max = max(live_max, slot + 1);
    }
  }
  live_max = __reduce_max_sync(0xffffffffu, live_max);
  odd_columns = __reduce_add_sync(0xffffffffu, odd_columns);
  if (lane == 0) {
    if (live_max) atomicMax(a.out + 2, live_max);
    if (odd_columns) atomicAdd(a.out + 3, (int)odd_columns);
  }
}

} // namespace

void fpb_domainfill_launch(const DomainfillArgs &a, cudaStream_t st, int64_t *launches, int phase) {
  const int nb = (a.ncols + DF_BLOCK - 1) / DF_BLOCK;
  if (phase == 0) {
    df_colmass_kernel<<<(a.ncols + 255) / 256, 256, 0, st>>>(a);
    df_total_kernel<<<1, 32, 0, st>>>(a);
    df_count_kernel<<<nb, DF_BLOCK, 0, st>>>(a);
    df_scan_blocks_kernel<<<1, DF_BLOCK, 0, st>>>(a.block_sums, nb);
    df_colstart_kernel<<<nb, DF_BLOCK, 0, st>>>(a);
    *launches += 5;
  } else {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int want = (a.ncols + FILL_WARPS - 1) / FILL_WARPS;
    const int grid = want < sms * 8 ? want : sms * 8;
    df_fill_kernel<<<grid, 32 * FILL_WARPS, 0, st>>>(a);
    *launches += 1;
  }
}
