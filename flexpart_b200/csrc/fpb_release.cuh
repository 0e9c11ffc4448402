// fpb_release.cuh -- releaseparticles(itime) on the device.
// Included inside the anonymous namespace of fpb_kernels.cu.
//
// The reference (src/releaseparticles.f90:139-378) walks the release points, and for each new
// particle searches the first slot from `minpart` on whose itra1 differs from itime; minpart only
// grows, so new particle #r of the call (points in order, particles in order) lands in the r-th
// such slot.  Here: count free slots per block, scan the block counts, then every free slot with
// rank r < M initialises particle r -- the same slots, no sequential search.  The host computes the
// per-point release counts (the rfraction / xmasssave recurrence, :89-123) and, in
// FPB_RNG_REFERENCE mode, the ran1 stream of the call (4 uniforms per particle in the reference's
// order); the Philox modes draw the four uniforms from the particle's own counter stream.
constexpr int REL_BLOCK = 1024;

__device__ __forceinline__ bool release_slot_free(const DevReleaseArgs &a, int s, int &row) {
  row = a.permuted ? a.row_of_slot[s] : s;
  return s >= a.numpart_old || a.p.itra1[row] != a.cfg.itime;
}

__global__ void __launch_bounds__(REL_BLOCK) release_count_kernel(const DevReleaseArgs a) {
  const int s = blockIdx.x * REL_BLOCK + threadIdx.x;
  int row;
  const int free_here = (s < a.p.maxpart) && release_slot_free(a, s, row);
  const int n = __syncthreads_count(free_here);
  if (threadIdx.x == 0) a.block_counts[blockIdx.x] = (unsigned)n;
}

// exclusive scan of the block counts (one block; the list is maxpart / 1024 long)
__global__ void __launch_bounds__(REL_BLOCK) release_scan_kernel(unsigned *block_counts, int *total, int nblocks) {
  __shared__ unsigned warp_tot[32];
  __shared__ unsigned running;
  if (threadIdx.x == 0) running = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int base = 0; base < nblocks; base += REL_BLOCK) {
    const int i = base + threadIdx.x;
    const unsigned v = (i < nblocks) ? block_counts[i] : 0u;
    unsigned inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const unsigned t = __shfl_up_sync(0xffffffffu, inc, d);
      if (lane >= d) inc += t;
    }
    if (lane == 31) warp_tot[w] = inc;
    __syncthreads();
    if (w == 0) {
      unsigned t = warp_tot[lane], ti = t;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const unsigned u = __shfl_up_sync(0xffffffffu, ti, d);
        if (lane >= d) ti += u;
      }
      warp_tot[lane] = ti - t; // exclusive over warps
    }
    __syncthreads();
    const unsigned excl = running + warp_tot[w] + inc - v;
    if (i < nblocks) block_counts[i] = excl;
    __syncthreads();
    if (threadIdx.x == REL_BLOCK - 1) running = excl + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) *total = (int)running;
}

__global__ void __launch_bounds__(REL_BLOCK) release_assign_kernel(const DevReleaseArgs a) {
  __shared__ unsigned warp_cnt[32];
  const DevCfg &c = a.cfg;
  const int s = blockIdx.x * REL_BLOCK + threadIdx.x;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int row = 0;
  const bool free_here = (s < a.p.maxpart) && release_slot_free(a, s, row);
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
  const int r = (int)(a.block_counts[blockIdx.x] + warp_cnt[w] + __popc(bal & ((1u << lane) - 1u)));
  if (r >= a.n_new) return;

  // release point of particle r: offsets[i] <= r < offsets[i + 1]
  int lo = 0, hi = a.numpoint - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (a.offsets[mid] <= r) lo = mid; else hi = mid - 1;
  }
  const int i = lo;

  float u1, u2, u3, u4;
  if (a.uniforms) { // the call's ran1 stream, drawn by the host in the reference's order
    const float4 u = reinterpret_cast<const float4 *>(a.uniforms)[r];
    u1 = u.x; u2 = u.y; u3 = u.z; u4 = u.w;
  } else {
    Rng rng;
    make_rng(c, nullptr, s, rng);
    u1 = rng.uniform(16u); u2 = rng.uniform(17u); u3 = rng.uniform(18u); u4 = rng.uniform(19u);
  }
  const float xaux = a.xpoint2[i] - a.xpoint1[i];
  const float yaux = a.ypoint2[i] - a.ypoint1[i];
  const float zaux = a.zpoint2[i] - a.zpoint1[i];
  // src/releaseparticles.f90:182-193 (single-precision sum widened on assignment)
  double x = (double)__fadd_rn(a.xpoint1[i], __fmul_rn(u1, xaux));
  if (c.xglobal) {
    if (x > (float)c.nxmin1) x = x - (float)c.nxmin1;
    if (x < 0.) x = x + (float)c.nxmin1;
  }
  a.p.xtra1[row] = x;
  a.p.ytra1[row] = (double)__fadd_rn(a.ypoint1[i], __fmul_rn(u2, yaux));
  for (int k = 0; k < c.nspec; k++) { // :195-203 (EMISVAR factors and the ind_rel density scaling at 1)
    a.p.xmass1[(size_t)k * a.p.maxpart + row] =
        __fdiv_rn(a.xmass[k * c.numpoint + i], (float)a.npart[i]) * 1.f / 1.f;
    if (c.drybkdep || c.wetbkdep) a.p.xscav_frac1[(size_t)k * a.p.maxpart + row] = -1.f;
  }
  const int nc = f_int(__fmul_rn(u3, (float)c.nclassunc)) + 1; // :209-210
  a.p.nclass[row] = min(nc, c.nclassunc);
  a.p.npoint[row] = i + 1;
  a.p.idt[row] = c.mintime;
  a.p.itra1[row] = c.itime;
  a.p.itramem[row] = c.itime;
  a.p.itrasplit[row] = c.itime + c.ldirect * a.itsplit;
  float z = __fadd_rn(a.zpoint1[i], __fmul_rn(u4, zaux)); // :226 (zkind 1: metres above ground)
  if (z < 1.e-6f) z = 1.e-6f;
  if (z > a.ztop - 0.5f) z = a.ztop - 0.5f;
  a.p.ztra1[row] = z;
  a.p.uap[row] = 0.f; a.p.ucp[row] = 0.f; a.p.uzp[row] = 0.f;
  a.p.us[row] = 0.f; a.p.vs[row] = 0.f; a.p.ws[row] = 0.f;
  a.p.cbt[row] = 1;
  a.p.slot[row] = s;
  atomicMax(a.out, s + 1);
}

// ---- particle splitting, src/timemanager.f90:472-503: every particle (slot j <= numpart) whose
// itrasplit has been reached is duplicated into slot numpart + (its rank among such particles), both
// halves carry half the mass; candidates beyond maxpart stay untouched.
__device__ __forceinline__ bool split_candidate(const DevSplitArgs &a, int s, int &row) {
  if (s >= a.numpart_old) return false;
  row = a.permuted ? a.row_of_slot[s] : s;
  return a.cfg.ldirect * a.cfg.itime >= a.cfg.ldirect * a.p.itrasplit[row];
}

__global__ void __launch_bounds__(REL_BLOCK) split_count_kernel(const DevSplitArgs a) {
  int row;
  const int n = __syncthreads_count(split_candidate(a, blockIdx.x * REL_BLOCK + threadIdx.x, row));
  if (threadIdx.x == 0) a.block_counts[blockIdx.x] = (unsigned)n;
}

__global__ void __launch_bounds__(REL_BLOCK) split_assign_kernel(const DevSplitArgs a) {
  __shared__ unsigned warp_cnt[32];
  const int s = blockIdx.x * REL_BLOCK + threadIdx.x;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int j = 0;
  const bool cand = split_candidate(a, s, j);
  const unsigned bal = __ballot_sync(0xffffffffu, cand);
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
  if (!cand) return;
  const int r = (int)(a.block_counts[blockIdx.x] + warp_cnt[w] + __popc(bal & ((1u << lane) - 1u)));
  const int n = a.numpart_old + r; // 0-based new slot = row (fresh rows are never permuted)
  if (n >= a.p.maxpart) return;
  const DevParticles &p = a.p;
  const int split = 2 * (p.itrasplit[j] - p.itramem[j]) + p.itramem[j];
  p.itrasplit[j] = split; p.itrasplit[n] = split;
  p.itramem[n] = p.itramem[j]; p.itra1[n] = p.itra1[j]; p.idt[n] = p.idt[j];
  p.npoint[n] = p.npoint[j]; p.nclass[n] = p.nclass[j];
  p.xtra1[n] = p.xtra1[j]; p.ytra1[n] = p.ytra1[j]; p.ztra1[n] = p.ztra1[j];
  p.uap[n] = p.uap[j]; p.ucp[n] = p.ucp[j]; p.uzp[n] = p.uzp[j];
  p.us[n] = p.us[j]; p.vs[n] = p.vs[j]; p.ws[n] = p.ws[j];
  p.cbt[n] = p.cbt[j];
  p.slot[n] = n;
  for (int ks = 0; ks < a.cfg.nspec; ks++) {
    const float m = p.xmass1[(size_t)ks * p.maxpart + j] / 2.f;
    p.xmass1[(size_t)ks * p.maxpart + j] = m;
    p.xmass1[(size_t)ks * p.maxpart + n] = m;
  }
}
