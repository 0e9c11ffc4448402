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

__global__ void __launch_bounds__(256) df_profiles_kernel(const DomainfillArgs a, const int2 *cols, int n, float *out) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n * a.cfg.nz) return;
  const int col = t / a.cfg.nz, kz = t % a.cfg.nz + 1;
  out[t] = level_pressure(a, cols[col].x, cols[col].y, kz);
}

// ------------------------------------------------------------------ boundcond_domainfill ----
// :59-72: particles that left the box are terminated
__global__ void __launch_bounds__(256) bc_terminate_kernel(const BoundcondArgs a) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.numpart_old) return;
  if (a.p.itra1[i] != a.cfg.itime) return;
  const double x = a.p.xtra1[i], y = a.p.ytra1[i];
  bool out = (y > (double)(float)a.ny1) || (y < (double)(float)a.ny0);
  if (a.check_x && ((x < (double)(float)a.nx0) || (x > (double)(float)a.nx1))) out = true;
  if (out) a.p.itra1[i] = FPB_ITRA_DEAD;
}

// :104-197 / :343-439: mass flux through one boundary location, accumulated mass, particles due
__global__ void __launch_bounds__(256) bc_flux_kernel(const BoundcondArgs a) {
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= a.nloc) return;
  const BcLoc q = a.loc[l];
  const DevCfg &c = a.cfg;
  const size_t plane = (size_t)c.nxd * c.nyd;
  const size_t o = (size_t)(q.indz - 1) * plane + (size_t)q.gy * c.nxd + q.gx;
  float windhl[2], rhohl[2];
#pragma unroll
  for (int m = 0; m < 2; m++) {
    const float4 lo = __ldg(a.A[m] + o), hi = __ldg(a.A[m] + o + plane);
    const float w1 = (q.flags & BC_WE) ? lo.x : lo.y, w2 = (q.flags & BC_WE) ? hi.x : hi.y;
    windhl[m] = __fmul_rn(__fadd_rn(__fmul_rn(q.dz2, w1), __fmul_rn(q.dz1, w2)), q.dz);
    rhohl[m] = __fmul_rn(__fadd_rn(__fmul_rn(q.dz2, lo.w), __fmul_rn(q.dz1, hi.w)), q.dz);
  }
  const float windx = __fmul_rn(__fadd_rn(__fmul_rn(windhl[0], a.dt2), __fmul_rn(windhl[1], a.dt1)), a.dtt);
  const float rhox = __fmul_rn(__fadd_rn(__fmul_rn(rhohl[0], a.dt2), __fmul_rn(rhohl[1], a.dt1)), a.dtt);
  const float flux = __fmul_rn(__fmul_rn(__fmul_rn(windx, rhox), q.boundarea), (float)c.lsynctime);
  float acc = a.acc_mass[l];
  if (!(q.flags & BC_K2)) acc = (flux >= 0.f) ? __fadd_rn(acc, flux) : 0.f;
  else acc = (flux <= 0.f) ? __fadd_rn(acc, fabsf(flux)) : 0.f;
  const float xm = a.xmassperparticle, half = __fdiv_rn(xm, 2.f);
  int mmass = 0;
  if (acc >= half) {
    mmass = (int)__fdiv_rn(__fadd_rn(acc, half), xm);
    acc = __fsub_rn(acc, __fmul_rn((float)mmass, xm));
  }
  a.acc_mass[l] = acc;
  a.mmass[l] = mmass;
}

__device__ __forceinline__ bool bc_slot_free(const BoundcondArgs &a, int s, int &row) {
  row = a.permuted ? a.row_of_slot[s] : s;
  return s >= a.numpart_old || a.p.itra1[row] != a.cfg.itime;
}

__global__ void __launch_bounds__(DF_BLOCK) bc_count_free_kernel(const BoundcondArgs a) {
  const int s = blockIdx.x * DF_BLOCK + threadIdx.x;
  int row;
  const int free_here = (s < a.p.maxpart) && bc_slot_free(a, s, row);
  const int n = __syncthreads_count(free_here);
  if (threadIdx.x == 0) a.block_counts[blockIdx.x] = (unsigned)n;
}

// exclusive scan of any number of block counts by one block; total -> *total
__global__ void __launch_bounds__(DF_BLOCK) bc_scan_kernel(unsigned *v, int n, int *total) {
  __shared__ unsigned wtot[32];
  __shared__ unsigned running;
  if (threadIdx.x == 0) running = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int base = 0; base < n; base += DF_BLOCK) {
    const int i = base + threadIdx.x;
    const unsigned x = (i < n) ? v[i] : 0u;
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
    const unsigned excl = running + wtot[w] + inc - x;
    if (i < n) v[i] = excl;
    __syncthreads();
    if (threadIdx.x == DF_BLOCK - 1) running = excl + x;
    __syncthreads();
  }
  if (threadIdx.x == 0) *total = (int)running;
}

// :199-313 / :441-541: the m-th particle this rank creates goes to its m-th free slot (the reference's
// minpart search: first slot whose itra1 differs from itime, minpart only grows)
__global__ void __launch_bounds__(DF_BLOCK) bc_create_kernel(const BoundcondArgs a) {
  __shared__ unsigned warp_cnt[32];
  const DevCfg &c = a.cfg;
  const int s = blockIdx.x * DF_BLOCK + threadIdx.x;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int row = 0;
  const bool free_here = (s < a.p.maxpart) && bc_slot_free(a, s, row);
  const unsigned bal = __ballot_sync(0xffffffffu, free_here);
  if (lane == 0) warp_cnt[w] = __popc(bal);
  __syncthreads();
  if (w == 0) {
    unsigned t = warp_cnt[lane], ti = t;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const unsigned u = __shfl_up_sync(0xffffffffu, ti, d);
      if (lane >= d) ti += u;
    }
    warp_cnt[lane] = ti - t;
  }
  __syncthreads();
  if (!free_here) return;
  const int m = (int)(a.block_counts[blockIdx.x] + warp_cnt[w] + __popc(bal & ((1u << lane) - 1u)));
  if (m >= a.n_mine) return;
  const int r = a.r0 + m * a.id_stride; // particle of the call, in the reference's creation order
  int lo = 0, hi = a.nloc - 1;          // its location: first[l] <= r < first[l + 1]
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (a.first[mid] <= r) lo = mid; else hi = mid - 1;
  }
  const BcLoc q = a.loc[lo];
  const int g = a.numparticlecount + r; // numparticlecount - 1 of this particle
  float u_h, u_z = 0.f, u_c;
  if (a.uniforms) {
    const int dpp = (q.flags & BC_ZDRAW) ? 3 : 2;
    const float *u = a.uniforms + a.u_off[lo] + (size_t)(r - a.first[lo]) * dpp;
    int k = 0;
    u_h = u[k++];
    if (q.flags & BC_ZDRAW) u_z = u[k++];
    u_c = u[k++];
  } else {
    const uint2 key = make_uint2((uint32_t)c.seed, (uint32_t)(c.seed >> 32));
    const uint4 v = philox4x32_10(make_uint4((uint32_t)g, (uint32_t)c.itime, 34u, 0u), key);
    u_h = u01(v.x); u_z = u01(v.y); u_c = u01(v.z);
  }
  // position along the boundary: half a cell at the two ends of the boundary, a whole cell else
  float along;
  if (q.flags & BC_EDGE_LOW) along = __fadd_rn((float)q.idx, __fmul_rn(0.5f, u_h));
  else if (q.flags & BC_EDGE_HIGH) along = __fsub_rn((float)q.idx, __fmul_rn(0.5f, u_h));
  else along = __fadd_rn((float)q.idx, __fsub_rn(u_h, .5f));
  const DevParticles &p = a.p;
  if (q.flags & BC_WE) { p.xtra1[row] = (double)(float)q.gx; p.ytra1[row] = (double)along; }
  else { p.ytra1[row] = (double)(float)q.gy; p.xtra1[row] = (double)along; }
  p.ztra1[row] = (q.flags & BC_ZDRAW) ? __fadd_rn(q.za, __fmul_rn(u_z, q.zb)) : q.za;
  const int nc = (int)__fmul_rn(u_c, (float)c.nclassunc) + 1;
  p.nclass[row] = min(nc, c.nclassunc);
  p.npoint[row] = g + 1;
  p.idt[row] = c.mintime;
  p.itra1[row] = c.itime;
  p.itramem[row] = c.itime;
  p.itrasplit[row] = c.itime + c.ldirect * a.itsplit;
  p.xmass1[row] = a.xmassperparticle;
  // (the reference leaves the other species and the velocity memory of the slot's last owner;
  //  initialize() sets the velocities and the CBL flag at the particle's first step)
  p.slot[row] = s;
  atomicMax(a.out, s + 1);
}

} // namespace

void fpb_domainfill_profiles(const DomainfillArgs &a, const int2 *cols, int n, float *out, cudaStream_t st) {
  if (n <= 0) return;
  df_profiles_kernel<<<(n * a.cfg.nz + 255) / 256, 256, 0, st>>>(a, cols, n, out);
}

void fpb_boundcond_launch(const BoundcondArgs &a, cudaStream_t st, int64_t *launches, int phase) {
  if (phase == 0) {
    if (a.numpart_old > 0) bc_terminate_kernel<<<(a.numpart_old + 255) / 256, 256, 0, st>>>(a);
    if (a.nloc > 0) bc_flux_kernel<<<(a.nloc + 255) / 256, 256, 0, st>>>(a);
    *launches += 2;
  } else {
    const int nb = (a.p.maxpart + DF_BLOCK - 1) / DF_BLOCK;
    bc_count_free_kernel<<<nb, DF_BLOCK, 0, st>>>(a);
    bc_scan_kernel<<<1, DF_BLOCK, 0, st>>>(a.block_counts, nb, a.out + 1);
    bc_create_kernel<<<nb, DF_BLOCK, 0, st>>>(a);
    *launches += 3;
  }
}

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
