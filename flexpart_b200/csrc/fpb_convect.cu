// fpb_convect.cu -- convmix on the device (src/convmix.f90:60-196, mother grid).
//
// The reference sorts the particles by grid column (sort2), walks the sorted list and, whenever a new
// column starts, interpolates the column's sounding in time and calls calcmatrix (-> the Emanuel
// scheme); every particle of a convecting column is then moved by redist.  Here:
//   conv_keys_kernel     column key of every active row (igrid = nint(y)*nx + nint(x), :96-134)
//   [stable LSD radix sort of (key, row), fpb_scatter.cu]
//   conv_heads_* kernels run heads -> column index of every sorted position, first position and
//                        key of every column (block scan)
//   conv_pre_kernel      ONE BLOCK PER 32 COLUMNS x 8 LEVEL RESIDUES: sounding (:163-176), calcmatrix's pressures and
//                        saturation humidity, level by level, on the column's slice of a work pool
//   conv_column_kernel   ONE THREAD PER OCCUPIED COLUMN: the O(n) part of convect (fpb_convect.cuh:
//                        conv_convect_head).
//                        The 32 columns of a warp interleave their slices element by element (stride
//                        32): the lanes run the same loops, so a warp-wide access to "element e of my
//                        column" is one 128-byte line instead of 32 scattered sectors.
//   conv_mix_kernel      ONE BLOCK PER 32 COLUMNS x 8 ROW RESIDUES: the loops over level pairs (mixing
//                        fractions, normalisation), independent row by row;
//                        latency-bound, so the inner loops request 4-8 elements together before working
//                        through them in the reference's order
//   conv_assembly_kernel ONE BLOCK PER COLUMN: the O(n^3) sums of the flux assembly on the column's
//                        MENT in shared memory, thread t = level t + 2, each sum in the reference's order;
//                        then the redistribution matrix from the same copy, thread t = row t + 1
//   conv_column_tail_kernel  one thread per column again: subsidence, cbaseflux out, heights of the
//                        eta half levels
//   conv_redist_kernel   one thread per particle of the batch's columns: redist
// Columns are processed in batches (the work pool holds up to 65536 columns, 80-140 KB each at 138
// levels).  Compiled with --fmad=false; the column arithmetic is bit-comparable with the
// reference's routines in every math mode (see fpb_convect.cuh).
#include "fpb_convect.cuh"
#include "fpb_convmix.cuh"
#include "fpb_output.cuh"

namespace {
using namespace fpbconv;

constexpr int CB = 1024;

__global__ void __launch_bounds__(256) conv_keys_kernel(const ConvmixArgs a) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.nrows) return;
  const DevCfg &c = a.cfg;
  unsigned key = 0xffffffffu;
  if (a.p.itra1[i] == c.itime) {
    const float x = (float)a.p.xtra1[i], y = (float)a.p.ytra1[i]; // x = xtra1(ipart): real
    int ngrid = 0; // innermost nest the particle is in, src/convmix.f90:100-118
    const float eps = a.ecmwf_eps ? c.eps : 0.f;
    for (int j = c.numbnests; j >= 1; j--)
      if (x > c.xln[j - 1] + eps && x < c.xrn[j - 1] - eps && y > c.yln[j - 1] + eps && y < c.yrn[j - 1] - eps) {
        ngrid = j;
        break;
      }
    int ix, jy;
    if (ngrid > 0) {
      const float xtn = (x - c.xln[ngrid - 1]) * c.xresoln[ngrid - 1], ytn = (y - c.yln[ngrid - 1]) * c.yresoln[ngrid - 1];
      ix = (int)roundf(xtn); jy = (int)roundf(ytn); // nint
    } else {
      ix = (int)roundf(x); jy = (int)roundf(y);
    }
    key = ((unsigned)ngrid << a.col_bits) | (unsigned)(jy * a.gnx[ngrid] + ix);
  }
  a.keys[i] = key;
  a.ids[i] = (unsigned)i;
  if (a.key_by_slot) a.key_by_slot[a.p.slot[i]] = (int)key; // (grid, igrid - 1) of the particle, -1: not due
}

// heads of the runs of equal keys among the sorted keys
__global__ void __launch_bounds__(CB) conv_heads_count_kernel(const ConvmixArgs a, const unsigned *keys) {
  const int i = blockIdx.x * CB + threadIdx.x;
  const bool head = i < a.nrows && keys[i] != 0xffffffffu && (i == 0 || keys[i - 1] != keys[i]);
  const int n = __syncthreads_count(head);
  if (threadIdx.x == 0) a.block_counts[blockIdx.x] = (unsigned)n;
}

__global__ void __launch_bounds__(CB) conv_scan_blocks_kernel(unsigned *v, int n, int *total) {
  __shared__ unsigned wtot[32];
  __shared__ unsigned running;
  if (threadIdx.x == 0) running = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int base = 0; base < n; base += CB) {
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
    if (threadIdx.x == CB - 1) running = excl + x;
    __syncthreads();
  }
  if (threadIdx.x == 0) *total = (int)running;
}

__global__ void __launch_bounds__(CB) conv_heads_assign_kernel(const ConvmixArgs a, const unsigned *keys) {
  __shared__ unsigned wcnt[32];
  const int i = blockIdx.x * CB + threadIdx.x;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const bool live = i < a.nrows && keys[i] != 0xffffffffu;
  const bool head = live && (i == 0 || keys[i - 1] != keys[i]);
  const unsigned bal = __ballot_sync(0xffffffffu, head);
  if (lane == 0) wcnt[w] = __popc(bal);
  __syncthreads();
  if (w == 0) {
    unsigned t = wcnt[lane], ti = t;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const unsigned u = __shfl_up_sync(0xffffffffu, ti, d);
      if (lane >= d) ti += u;
    }
    wcnt[lane] = ti - t;
  }
  __syncthreads();
  if (!live) return;
  // heads up to and including this position
  const int incl = (int)(a.block_counts[blockIdx.x] + wcnt[w] + __popc(bal & ((2u << lane) - 1u)));
  const int c = incl - 1;
  a.colidx[i] = c;
  if (head) {
    a.col_key[c] = keys[i];
    a.col_start[c] = i;
  }
  if (i + 1 == a.nrows || keys[i + 1] == 0xffffffffu) a.col_start[incl] = i + 1; // end of the last column
}

// Per column and contiguous (pool2): the final MENT once more, element (i,j) at [i + lt*j] with the leading dimension of
// the flux assembly's shared-memory copy (conv_lt: odd, so that a walk along a row of the matrix is free of bank
// conflicts) -- the assembly fetches it as it lies with one bulk copy -- and FMASS, (i,j) at [i + ld*j].
__host__ __device__ inline int conv_lt(int ld) { return ld <= 65 ? 65 : ld <= 97 ? 97 : ld <= 129 ? 129 : ld; }
__host__ __device__ inline size_t conv_mentc_floats(int ld) { return ((size_t)conv_lt(ld) * conv_lt(ld) + 3) / 4 * 4; }
__host__ __device__ inline size_t conv_slab_floats(int ld) { return conv_mentc_floats(ld) + ((size_t)ld * ld + 3) / 4 * 4; }

// the column's slice of the work pool: block of 32 slices (q / 32), lane q % 32, elements interleaved
__device__ __forceinline__ void conv_column_work(const ConvmixArgs &a, int q, ConvWork &w) {
  const size_t nfl = conv_pool_floats(a.nuvz, a.nconvlev, false);
  conv_carve(w, a.pool + (size_t)(q / 32) * 32 * nfl + (q % 32), a.nuvz, a.nconvlev, 32, false);
  w.akz = a.akz; w.bkz = a.bkz; w.akm = a.akm; w.bkm = a.bkm;
  w.mentc = a.pool2 + (size_t)q * conv_slab_floats(w.ld);
  w.fmass = w.mentc + conv_mentc_floats(w.ld);
  w.fstride = 1;
}
__device__ __forceinline__ void conv_column_place(const ConvmixArgs &a, int c, int &g, size_t &o2, size_t &plane) {
  const unsigned key = a.col_key[c];
  g = (int)(key >> a.col_bits);
  const unsigned col = key & ((1u << a.col_bits) - 1u);
  const int jy = (int)(col / (unsigned)a.gnx[g]), ix = (int)(col - (unsigned)jy * a.gnx[g]);
  o2 = (size_t)jy * a.gnxd[g] + ix;
  plane = (size_t)a.gnxd[g] * a.gnyd[g];
}

// what every level of a column gets on its own -- the sounding interpolated in time (src/convmix.f90:163-176), pressures,
// saturation humidity (conv_calcmatrix_level) -- and the zeroing of the column's vectors, level-parallel: one block per
// group of 32 columns x PRE_ROWS level residues
constexpr int PRE_ROWS = 8;
__global__ void __launch_bounds__(32 * PRE_ROWS) conv_pre_kernel(const ConvmixArgs a, int c0, int c1) {
  const int c = c0 + blockIdx.x * 32 + threadIdx.x;
  const int y = threadIdx.y;
  const DevCfg &cf = a.cfg;
  const int nuvz = a.nuvz;
  ConvWork w;
  conv_column_work(a, c - c0, w);
  const size_t nvec = (size_t)(CONV_NVEC + 1) * (nuvz + 4);
  for (size_t k = y; k < nvec; k += PRE_ROWS) w.pconv[k * 32] = 0.f; // (the reference's zero-initialised locals; pconv = first vector)
  __syncthreads();
  if (c >= c1) return;
  int g;
  size_t o2, plane;
  conv_column_place(a, c, g, o2, plane);
  // src/convmix.f90:62-65,163-171 (nests: :221-233)
  const float dt1 = (float)(cf.itime - cf.memtime[0]), dt2 = (float)(cf.memtime[1] - cf.itime);
  const float dtt = 1.f / (dt1 + dt2);
  const float4 s1 = a.CS[g][0][o2], s2 = a.CS[g][1][o2];
  w.psconv = (s1.x * dt2 + s2.x * dt1) * dtt;
  for (int kz = 1 + y; kz <= nuvz - 1; kz += PRE_ROWS) {
    const float2 q1 = a.CT[g][0][(size_t)kz * plane + o2], q2 = a.CT[g][1][(size_t)kz * plane + o2]; // level kz+1
    w.tconv[(size_t)kz * w.stride] = (q1.x * dt2 + q2.x * dt1) * dtt;
    w.qconv[(size_t)kz * w.stride] = (q1.y * dt2 + q2.y * dt1) * dtt;
    conv_calcmatrix_level(w, kz);
  }
}

// first half: the O(n) part of the Emanuel scheme up to the loops over level pairs, one thread per column
__global__ void __launch_bounds__(32) conv_column_kernel(const ConvmixArgs a, int c0, int c1) {
  const int c = c0 + blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= c1) return;
  const DevCfg &cf = a.cfg;
  ConvWork w;
  conv_column_work(a, c - c0, w);
  int g;
  size_t o2, plane;
  conv_column_place(a, c, g, o2, plane);
  const float dt1 = (float)(cf.itime - cf.memtime[0]), dt2 = (float)(cf.memtime[1] - cf.itime);
  const float dtt = 1.f / (dt1 + dt2);
  const float4 s1 = a.CS[g][0][o2], s2 = a.CS[g][1][o2];
  w.psconv = (s1.x * dt2 + s2.x * dt1) * dtt;
  float cbmf = a.cbaseflux[g][o2];
  ConvState st;
  conv_calcmatrix_a(w, (float)abs(cf.lsynctime), cbmf, st, true, true);
  static_cast<ConvState *>(a.col_state)[c] = st;
}

// the loops over level pairs of the scheme (mixing fractions, normalisation: conv_mixnorm_row) with ONE BLOCK PER GROUP
// OF 32 COLUMNS: lane = column of the group (the interleaved layout: a warp still reads "element e of 32 columns" as
// one line), threadIdx.y = the rows i = icb + 1 + y, + MIX_ROWS, ... of every column.  The rows are independent, so the
// bits are the sequential loop's; what changes is that a column's chain is 1 / MIX_ROWS as long and MIX_ROWS times as
// many warps are there to hide the loads.  The eight vectors the walk along a row reads level by level (ft, tconv,
// qsconv, h, qconv, fq, lv, phconv_hpa) are staged in shared memory first: in the interleaved layout a vector of the
// group is one contiguous stretch, so each is ONE bulk asynchronous copy (cp.async.bulk, completion on an mbarrier),
// and the row code reads them through the same pointers, re-based.  After a block barrier the same threads write the
// contiguous copy of MENT (the staging area becomes the tiles of the transposing copy).
#ifndef FPB_MIX_ROWS
#define FPB_MIX_ROWS 32 // (A/B with the vectors staged: 16 rows x batch 8: 10.1 ms, 24 x 4: 9.4, 32 x 4: 9.5, 32 x 2: 9.2; gpurun_out/ab_conv_rows.txt)
#endif
#ifndef FPB_MIX_MINB
#define FPB_MIX_MINB 1
#endif
#ifndef FPB_TILE_UNROLL
#define FPB_TILE_UNROLL 16
#endif
constexpr int TILE_UNROLL = FPB_TILE_UNROLL;
constexpr int MIX_ROWS = FPB_MIX_ROWS;
constexpr int MIX_NSTAGE = 8; // vectors staged
constexpr size_t MIX_TILE_BYTES = (size_t)MIX_ROWS * 32 * 33 * sizeof(float);
// levels_max: levels the launch reserved staging room for (0: no staging)
__global__ void __launch_bounds__(32 * MIX_ROWS, FPB_MIX_MINB) conv_mix_kernel(const ConvmixArgs a, int c0, int c1, int levels_max) {
  extern __shared__ __align__(128) float mixsm[];
  __shared__ __align__(8) unsigned long long bar;
  const int c = c0 + blockIdx.x * 32 + threadIdx.x;
  const int y = threadIdx.y;
  ConvState st;
  st.go = 0;
  if (c < c1) st = static_cast<const ConvState *>(a.col_state)[c];
  ConvWork w;
  conv_column_work(a, c - c0, w); // (the pool is whole groups of 32 slices: valid for the idle lanes of the last group too)
  const unsigned gomask = __ballot_sync(0xffffffffu, st.go != 0);
  if (gomask == 0u) return; // (block-uniform: every warp sees the same 32 columns)
  const int imax = __reduce_max_sync(0xffffffffu, st.go ? st.inb : 0);
  const int hi = imax + 1; // levels 1 .. hi are read (phconv_hpa up to inb + 1)
  if (hi <= levels_max) {
    const unsigned bar_s = (unsigned)__cvta_generic_to_shared(&bar);
    float **vec[MIX_NSTAGE] = {&w.ft, &w.tconv, &w.qsconv, &w.h, &w.qconv, &w.fq, &w.lv, &w.phconv_hpa};
    if (threadIdx.x == 0 && y == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_s));
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      const unsigned bytes = (unsigned)hi * 32u * 4u; // per vector: element 1 .. hi of the 32 columns
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_s), "r"(bytes * MIX_NSTAGE) : "memory");
#pragma unroll
      for (int v = 0; v < MIX_NSTAGE; v++) {
        const unsigned dst_s = (unsigned)__cvta_generic_to_shared(mixsm + (size_t)v * hi * 32);
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_s),
                     "l"(*vec[v] + 32), "r"(bytes), "r"(bar_s)
                     : "memory");
      }
    }
    __syncthreads(); // (the barrier is initialised for everyone)
    unsigned done = 0;
    while (!done)
      asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                   : "=r"(done)
                   : "r"(bar_s), "r"(0u)
                   : "memory");
#pragma unroll
    for (int v = 0; v < MIX_NSTAGE; v++) *vec[v] = mixsm + (size_t)v * hi * 32 + threadIdx.x - 32; // element e at [e * 32]
  }
  if (st.go) {
    for (int i = st.icb + 1 + y; i <= st.inb; i += MIX_ROWS) {
      conv_mixnorm_row(w, st, i);
    }
  }
  __syncthreads();
  // the final MENT of the group's columns once more, contiguous per column, for the flux assembly: a transposing copy
  // through shared memory, tiles of 32 consecutive elements (rows i at one j) x 32 columns, so that both the reads
  // (element e of 32 columns = one line) and the writes (32 elements of one column = one line) are whole lines.  Every
  // column is copied over the union of the group's index ranges; what lies outside a column's own range is never read.
  float(*tile)[32][33] = reinterpret_cast<float(*)[32][33]>(mixsm);
  const int imin = __reduce_min_sync(0xffffffffu, st.go ? st.icb + 1 : 0x7fffffff);
  const int jmin = imin - 1, jmax = imax;
  const int nchunk = (imax - imin + 32) / 32;
  const int ld = w.ld;
  const float *gment = w.ment - threadIdx.x; // element e of column l of the group at [e * 32 + l]
  const int lt = conv_lt(ld);
  const size_t slab = conv_slab_floats(ld);
  float *gout = a.pool2 + (size_t)(blockIdx.x * 32) * slab;
  for (int t = y; t < (jmax - jmin + 1) * nchunk; t += MIX_ROWS) {
    const int j = jmin + t / nchunk, i0 = imin + 32 * (t % nchunk);
    const size_t e0 = (size_t)i0 + (size_t)ld * j;
#pragma unroll TILE_UNROLL
    for (int r = 0; r < 32; r++) tile[y][r][threadIdx.x] = (i0 + r <= imax) ? gment[(e0 + r) * 32 + threadIdx.x] : 0.0f;
    __syncwarp();
    if (i0 + (int)threadIdx.x <= imax) {
#pragma unroll 8
      for (int l = 0; l < 32; l++)
        if ((gomask >> l) & 1u) gout[(size_t)l * slab + (size_t)i0 + (size_t)lt * j + threadIdx.x] = tile[y][threadIdx.x][l];
    }
    __syncwarp();
  }
}

// the flux assembly (src/convect43c.f90:855-913; conv_flux_assembly is its definition) with ONE BLOCK PER COLUMN:
// thread t owns level i = 2 + t and runs that level's two sums in the reference's order, on the column's MENT in
// shared memory (the one-thread-per-column walk re-read the matrix from DRAM once per level: two thirds of
// fpb_convmix, profiles/convmix_column_r02.txt)
constexpr int ASM_THREADS = 128;
#ifndef FPB_ASM_U
#define FPB_ASM_U 8
#endif
constexpr int ASM_U = FPB_ASM_U; // elements of a row requested together
#define WV(a, i) w.a[(size_t)(i) * w.stride]
// LT: leading dimension of the shared-memory copy, a compile-time constant (odd: the walk along a row of the matrix is
// free of bank conflicts) so that the eight loads of an unrolled step are one address register plus immediates;
// 0: the column's own ld (any number of levels)
template <int LT> __global__ void __launch_bounds__(ASM_THREADS) conv_assembly_kernel(const ConvmixArgs a, int c0, int c1) {
  extern __shared__ __align__(128) float smem[];
  __shared__ int s_top;
  __shared__ __align__(8) unsigned long long bar; // mbarrier of the bulk copy
  const int c = c0 + blockIdx.x;
  ConvState *stp = static_cast<ConvState *>(a.col_state) + c;
  if (!stp->go) return; // block-uniform
  const unsigned bar_s = (unsigned)__cvta_generic_to_shared(&bar);
  if (threadIdx.x == 0) {
    s_top = 1;
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_s));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int inb = stp->inb, icb = stp->icb, nk = stp->nk;
  const float delti = stp->delti;
  ConvWork w;
  conv_column_work(a, c - c0, w);
  const int ld = w.ld;
  const int lt = LT ? LT : ld;
  float *T = smem;                       // MENT, (i,j) at [i + lt*j]: the image of the column's mentc
  float *mv = smem + conv_mentc_floats(ld); // m(1 .. inb+1)
  float *ph = mv + ld;                   // phconv_hpa(1 .. inb+2)
  // Columns icb .. inb of the matrix as they lie in mentc, with ONE bulk copy (cp.async.bulk, global -> shared memory,
  // completion on the mbarrier; the range widened to 16-byte boundaries).  The sums read rows icb+1 .. inb of columns
  // icb .. inb+1: the rows were written for these columns by conv_mix_kernel, column inb+1 (all 0) is set here.
  const int f0 = (lt * icb) & ~3, f1 = (lt * (inb + 1) + 3) & ~3;
  if (threadIdx.x == 0) {
    const unsigned bytes = (unsigned)(f1 - f0) * 4u;
    const unsigned dst_s = (unsigned)__cvta_generic_to_shared(T + f0);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_s), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_s),
                 "l"(w.mentc + f0), "r"(bytes), "r"(bar_s)
                 : "memory");
  }
  for (int i = threadIdx.x; i < ld; i += ASM_THREADS) { // (while the copy is in flight)
    mv[i] = (i >= 1 && i <= inb + 1) ? WV(m, i) : 0.0f;
    ph[i] = (i >= 1 && i <= inb + 2 && i < ld) ? WV(phconv_hpa, i) : 0.0f;
  }
  {
    unsigned done = 0;
    while (!done)
      asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                   : "=r"(done)
                   : "r"(bar_s), "r"(0u)
                   : "memory");
  }
  for (int i = icb + 1 + threadIdx.x; i <= inb; i += ASM_THREADS) T[i + lt * (inb + 1)] = 0.0f;
  __syncthreads();
  using namespace k;
  int flag4 = 0;
  if (threadIdx.x == 0) { // level 1 (:855-866)
    const float dpinv = 0.01f / (ph[1] - ph[2]);
    float am = 0.0f;
    if (nk == 1)
      for (int kq = 2; kq <= inb; kq++) am = am + mv[kq];
    WV(fup, 1) = am;
    if ((2.f * G * dpinv * am) >= delti) flag4 = 1;
  }
  for (int i = 2 + threadIdx.x; i <= inb; i += ASM_THREADS) {
    const float dpinv = 0.01f / (ph[i] - ph[i + 1]);
    float amp1 = 0.0f, ad = 0.0f;
    if (i >= nk)
      for (int kq = i + 1; kq <= inb + 1; kq++) amp1 = amp1 + mv[kq];
    // The two sums walk blocks of the same shape -- amp1: rows icb+1 .. i, columns i+1 .. inb+1 of MENT, row by row;
    // ad: columns icb .. i-1, rows max(i, icb+1) .. inb, column by column -- so for i > icb they run side by side, two
    // independent chains of additions per thread, each in the reference's order (i <= icb: both blocks are empty)
    if (i >= icb + 1) {
      const int n = inb + 1 - i;
      for (int o = 0; o < i - icb; o++) {
        const float *p = T + (icb + 1 + o) + lt * (i + 1);
        const float *q = T + i + lt * (icb + o);
        int r = n;
        for (; r >= ASM_U; r -= ASM_U, p += ASM_U * lt, q += ASM_U) {
          float v[ASM_U], x[ASM_U];
#pragma unroll
          for (int u = 0; u < ASM_U; u++) { v[u] = p[u * lt]; x[u] = q[u]; }
#pragma unroll
          for (int u = 0; u < ASM_U; u++) { amp1 = amp1 + v[u]; ad = ad + x[u]; }
        }
        if (r > 0) { // (the last 1 .. ASM_U-1 elements: requested together like a full step, added in order)
          float v[ASM_U - 1], x[ASM_U - 1];
#pragma unroll
          for (int u = 0; u < ASM_U - 1; u++) {
            v[u] = u < r ? p[u * lt] : 0.0f;
            x[u] = u < r ? q[u] : 0.0f;
          }
#pragma unroll
          for (int u = 0; u < ASM_U - 1; u++)
            if (u < r) { amp1 = amp1 + v[u]; ad = ad + x[u]; }
        }
      }
    }
    WV(fup, i) = amp1;
    if ((2.f * G * dpinv * amp1) >= delti) flag4 = 1;
    WV(fdown, i) = ad;
  }
  if (flag4) stp->iflag = 4; // (every writer writes the same value)
  // the redistribution matrix (conv_fmass_row: calcmatrix's scaling of FMASS = MENT (+ M in row NK)), thread t = rows
  // 1 + t, ..., from the shared-memory MENT into the column's contiguous FMASS: nconvtop and lconv do not depend on
  // what the sums above found (IFLAG 1 or 4)
  ConvState st = *stp;
  if (!((st.iflag == 1 || st.iflag == 4) && !(st.cbmf <= 0.f && st.cbmfold <= 0.f))) return; // lconv (conv_calcmatrix_b)
  int top = 0; // conv_convect_b's nconvtop, level by level
  for (int i = 1 + threadIdx.x; i <= inb + 1; i += ASM_THREADS) {
    float f = 0.0f;
    f = f + mv[i];
    if (f > EPSILON) top = i > nk ? i : nk;
    if (i >= icb + 1 && i <= inb) {
      const int t = WV(rowtop, i);
      top = top > t ? top : t;
    }
  }
  atomicMax(&s_top, top);
  __syncthreads();
  w.nconvtop = s_top + 1;
  const float delt = (float)abs(a.cfg.lsynctime);
  for (int kq = 1 + threadIdx.x; kq <= w.nconvtop; kq += ASM_THREADS)
    conv_fmass_row(w, st, delt, kq, [&](int i, int j) { return T[i + lt * j]; });
}
#undef WV

// second half: subsidence, cbaseflux out, heights of the half levels
__global__ void __launch_bounds__(32) conv_column_tail_kernel(const ConvmixArgs a, int c0, int c1) {
  const int c = c0 + blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= c1) return;
  const DevCfg &cf = a.cfg;
  ConvWork w;
  conv_column_work(a, c - c0, w);
  int g;
  size_t o2, plane;
  conv_column_place(a, c, g, o2, plane);
  // (psconv, tt2conv, td2conv of the first half: conv_uvzlev reads them)
  const float dt1 = (float)(cf.itime - cf.memtime[0]), dt2 = (float)(cf.memtime[1] - cf.itime);
  const float dtt = 1.f / (dt1 + dt2);
  const float4 s1 = a.CS[g][0][o2], s2 = a.CS[g][1][o2];
  w.psconv = (s1.x * dt2 + s2.x * dt1) * dtt;
  w.tt2conv = (s1.y * dt2 + s2.y * dt1) * dtt;
  w.td2conv = (s1.z * dt2 + s2.z * dt1) * dtt;
  const ConvState st = static_cast<const ConvState *>(a.col_state)[c];
  w.nconvtop = 0;
  float cbmf = st.cbmf;
  const bool lconv = conv_calcmatrix_b(w, (float)abs(cf.lsynctime), cbmf, st, false, true); // (rows: conv_mix_kernel)
  a.cbaseflux[g][o2] = cbmf;
  a.col_lconv[c] = lconv ? w.nconvtop : 0;
  if (lconv) conv_uvzlev(w);
}

// mode 0: mark the particles that will draw a uniform (reference RNG replay); 1: redistribute
__global__ void __launch_bounds__(128) conv_redist_kernel(const ConvmixArgs a, int c0, int i0, int i1, int mode) {
  const int i = i0 + blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= i1) return;
  const DevCfg &cf = a.cfg;
  const int c = a.colidx[i];
  const int nconvtop = a.col_lconv[c];
  if (nconvtop == 0) return; // the column does not convect
  ConvWork w;
  conv_column_work(a, c - c0, w);
  w.nconvtop = nconvtop;
  const int row = (int)a.sorted_ids[i];
  const int slot = a.p.slot[row];
  float z = a.p.ztra1[row];
  const int levold = conv_levold(w, z);
  if (mode == 0) {
    a.draws[slot] = levold > 0 ? 1 : 0;
    return;
  }
  if (levold > 0) {
    float rn;
    if (a.rn_by_slot) {
      rn = a.rn_by_slot[slot];
    } else { // counter stream of the particle (stream 48), src/redist.f90:140 `rn = ran3(iseed)`
      const uint4 r = philox4x32_10(make_uint4((uint32_t)(cf.part_id_offset + cf.part_id_stride * slot),
                                               (uint32_t)cf.itime, 48u, 0u),
                                    make_uint2((uint32_t)cf.seed, (uint32_t)(cf.seed >> 32)));
      rn = u01(r.x);
    }
    z = conv_redist(w, z, levold, rn, cf.ldirect, cf.lsynctime);
  }
  if (z > a.ztop - 0.5f) z = a.ztop - 0.5f; // label 90
  const float ztold = a.p.ztra1[row];
  a.p.ztra1[row] = z;
  if (a.flux) { // gross fluxes of the convective displacement, src/convmix.f90:205-218
    const int itage = abs(a.p.itra1[row] - a.p.itramem[row]);
    int nage;
    for (nage = 1; nage <= cf.nageclass; nage++)
      if (itage < cf.lage[nage - 1]) break;
    if (nage <= cf.nageclass) {
      const double xt = a.p.xtra1[row], yt = a.p.ytra1[row];
      const int kp = (cf.ioutputforeachrelease == 1 && cf.mdomainfill == 0) ? a.p.npoint[row] : 1;
      fpb_flux_particle(cf, a.flux, nage, kp, (float)xt, (float)yt, ztold, xt, yt, z,
                        [&](int k) { return a.p.xmass1[(size_t)(k - 1) * a.p.maxpart + row]; });
    }
  }
}

} // namespace

void fpb_convmix_keys(const ConvmixArgs &a, cudaStream_t st) {
  conv_keys_kernel<<<(a.nrows + 255) / 256, 256, 0, st>>>(a);
}
void fpb_convmix_heads(const ConvmixArgs &a, const unsigned *sorted_keys, int *total, cudaStream_t st) {
  const int nb = (a.nrows + CB - 1) / CB;
  conv_heads_count_kernel<<<nb, CB, 0, st>>>(a, sorted_keys);
  conv_scan_blocks_kernel<<<1, CB, 0, st>>>(a.block_counts, nb, total);
  conv_heads_assign_kernel<<<nb, CB, 0, st>>>(a, sorted_keys);
}
void fpb_convmix_columns(const ConvmixArgs &a, int c0, int c1, cudaStream_t st) {
  if (c1 <= c0) return;
  const int ld = a.nconvlev + 3;
  const int lt = conv_lt(ld);
  const size_t smem = (conv_mentc_floats(ld) + 2 * (size_t)ld) * sizeof(float);
  static bool attr_set_dev[64] = {}; // (67 / 135 KB of dynamic shared memory: above the default limit; per device)
  int dev = 0;
  cudaGetDevice(&dev);
  bool &attr_set = attr_set_dev[dev & 63];
  if (!attr_set) {
    cudaFuncSetAttribute(conv_assembly_kernel<65>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(conv_assembly_kernel<97>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(conv_assembly_kernel<129>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(conv_assembly_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(conv_mix_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    attr_set = true;
  }
  conv_pre_kernel<<<(c1 - c0 + 31) / 32, dim3(32, PRE_ROWS), 0, st>>>(a, c0, c1);
  conv_column_kernel<<<(c1 - c0 + 31) / 32, 32, 0, st>>>(a, c0, c1);
  // staging room of the level-pair kernel: eight vectors x levels 1 .. nconvlev + 1 x 32 columns, if that fits
  size_t mix_smem = (size_t)MIX_NSTAGE * (a.nconvlev + 1) * 32 * sizeof(float);
  int levels_max = a.nconvlev + 1;
  if (mix_smem > 200 * 1024) { mix_smem = 0; levels_max = 0; }
  if (mix_smem < MIX_TILE_BYTES) mix_smem = MIX_TILE_BYTES;
  conv_mix_kernel<<<(c1 - c0 + 31) / 32, dim3(32, MIX_ROWS), mix_smem, st>>>(a, c0, c1, levels_max);
  if (lt == 65) conv_assembly_kernel<65><<<c1 - c0, ASM_THREADS, smem, st>>>(a, c0, c1);
  else if (lt == 97) conv_assembly_kernel<97><<<c1 - c0, ASM_THREADS, smem, st>>>(a, c0, c1);
  else if (lt == 129) conv_assembly_kernel<129><<<c1 - c0, ASM_THREADS, smem, st>>>(a, c0, c1);
  else conv_assembly_kernel<0><<<c1 - c0, ASM_THREADS, smem, st>>>(a, c0, c1);
  conv_column_tail_kernel<<<(c1 - c0 + 31) / 32, 32, 0, st>>>(a, c0, c1);
}
size_t fpb_convmix_pool2_floats(int nconvlev) { return conv_slab_floats(nconvlev + 3); }
size_t fpb_convmix_state_bytes() { return sizeof(fpbconv::ConvState); }
void fpb_convmix_redist(const ConvmixArgs &a, int c0, int i0, int i1, int mode, cudaStream_t st) {
  if (i1 > i0) conv_redist_kernel<<<(i1 - i0 + 127) / 128, 128, 0, st>>>(a, c0, i0, i1, mode);
}
size_t fpb_convmix_pool_floats(int nuvz, int nconvlev) { return fpbconv::conv_pool_floats(nuvz, nconvlev, false); }
