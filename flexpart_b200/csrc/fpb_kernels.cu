// fpb_kernels.cu -- the particle-timestep kernels (sm_100a).
//
// Compiled twice into libfpb.so:
//   -DFPB_STRICT=0            : production; FMA contraction on, float libm
//   -DFPB_STRICT=1 --fmad=false: validation; every float operation rounded
//                               once, transcendentals evaluated in double and
//                               rounded once -> bit-comparable with oracle/.
//
// Kernels (fpb_step.cuh, fpb_release.cuh, fpb_cbl.cuh are included below):
//   fpb_init_kernel      initialize() for new particles           (src/initialize.f90:66-217)
//   fpb_bkdep_kernel     receptor scavenging of backward runs      (src/timemanager.f90:563-598)
//   fpb_pbl_kernel       persistent; the Langevin sub-step loop    (src/advance.f90:133-609) with
//                        interpol_all/misslev/vdep, hanna/hanna_short/hanna1, cbl, settling
//   fpb_finish_kernel    rest of advance() + rest of the loop body (src/advance.f90:629-985,
//                        src/timemanager.f90:630-707): interpol_wind(_short), windalign, cmapf,
//                        Petterssen, decay, dry-deposition split + drydepokernel(_nest), terminations
//   fpb_conccalc_kernel / fpb_conc_emit_kernel / fpb_receptor_kernel   (src/conccalc.f90:50-498)
//   fpb_wetdepo_kernel   wetdepo + get_wetscav + wetdepokernel(_nest)  (src/wetdepo.f90:70-147)
//   release_* / split_*  releaseparticles, particle splitting      (src/releaseparticles.f90:69-378)
//
// The reference keeps its scratch in module globals (interpol_mod, hanna_mod); here it is
// per-thread registers plus a per-lane shared-memory row: the nzmax-long profile cache
// (indzindicator, src/advance.f90:310-331) becomes a direct-mapped cache of 8 levels; a recomputed
// level gives the same bits because the horizontal and time weights are frozen for the whole call.
#include <cstdio>
#include <math.h>

#include "fpb_device.cuh"

#ifndef FPB_STRICT
#define FPB_STRICT 0
#endif
#ifndef FPB_PBL_MIN_BLOCKS
#define FPB_PBL_MIN_BLOCKS 5 // resident 128-thread CTAs per SM the sub-step kernel is tuned for
#endif

namespace {

// ---------------------------------------------------------------- math ----
#if FPB_STRICT
__device__ __forceinline__ float m_exp(float x) { return (float)exp((double)x); }
__device__ __forceinline__ float m_log(float x) { return (float)log((double)x); }
__device__ __forceinline__ float m_pow(float a, float b) { return (float)pow((double)a, (double)b); }
__device__ __forceinline__ float m_sin(float x) { return (float)sin((double)x); }
__device__ __forceinline__ float m_cos(float x) { return (float)cos((double)x); }
__device__ __forceinline__ float m_erf(float x) { return (float)erf((double)x); }
__device__ __forceinline__ float m_log10(float x) { return (float)log10((double)x); }
#else
// production: MUFU-based exp / pow (a > 0 at every call site); relative error
// ~1e-6, three orders of magnitude inside the 1e-5 position tolerance
__device__ __forceinline__ float m_exp(float x) { return __expf(x); }
__device__ __forceinline__ float m_log(float x) { return __logf(x); }
__device__ __forceinline__ float m_pow(float a, float b) { return exp2f(b * __log2f(a)); }
__device__ __forceinline__ float m_sin(float x) { return sinf(x); }
__device__ __forceinline__ float m_cos(float x) { return cosf(x); }
__device__ __forceinline__ float m_erf(float x) { return erff(x); }
__device__ __forceinline__ float m_log10(float x) { return __log10f(x); }
#endif
__device__ __forceinline__ float m_sqrt(float x) { return sqrtf(x); }
__device__ __forceinline__ int f_int(float x) { return (int)x; }   // Fortran int()
__device__ __forceinline__ int d_int(double x) { return (int)x; }
__device__ __forceinline__ int d_nint(double x) { return (int)round(x); } // nint()
__device__ __forceinline__ double d_modulo(double a, double p) {
  double r = fmod(a, p);
  if (r != 0.0 && ((r < 0.0) != (p < 0.0))) r += p;
  return r;
}

constexpr float PI_F = 3.14159265f;            // par_mod pi
constexpr float PI180 = PI_F / 180.f;          // par_mod pi180
constexpr float HREF = 15.f;                   // par_mod href
constexpr float EPS2 = 1.e-9f;                 // advance.f90:107
constexpr float EPS3 = 1.17549435e-38f;        // tiny(1.0)
constexpr float EPS_SIG = 1.0e-30f;            // interpol_*.f90 eps
constexpr float MINMASS = 0.0001f;             // par_mod minmass

// ----------------------------------------------------------------- RNG ----
// Philox4x32-10 and u01(): fpb_device.cuh (shared with fpb_domainfill.cu)

// 4 normals of one Philox block: Box-Muller, clipped to +-3 like gasdev1
// (src/random_mod.f90:86-89).  Out of line and by value so that the caller's
// generator state can stay in registers.
__device__ __noinline__ float4 philox_normals4(uint32_t pid, uint32_t tstep, uint32_t blk, uint2 key) {
  const uint4 r = philox4x32_10(make_uint4(pid, tstep, 0x52414e44u, blk), key);
  const float u1 = ((float)(r.x >> 8) + 0.5f) * (1.0f / 16777216.0f);
  const float u2 = u01(r.y);
  const float u3 = ((float)(r.z >> 8) + 0.5f) * (1.0f / 16777216.0f);
  const float u4 = u01(r.w);
  const float ra = sqrtf(-2.f * __logf(u1)), rb = sqrtf(-2.f * __logf(u3));
  float sa, ca, sb, cb;
  __sincosf(6.28318530718f * u2, &sa, &ca);
  __sincosf(6.28318530718f * u4, &sb, &cb);
  return make_float4(fminf(fmaxf(ra * ca, -3.f), 3.f), fminf(fmaxf(ra * sa, -3.f), 3.f),
                     fminf(fmaxf(rb * cb, -3.f), 3.f), fminf(fmaxf(rb * sb, -3.f), 3.f));
}

struct Rng {
  const float *tab; // rannumb, 0-based
  int maxrand;
  int mode;
  uint2 key;
  uint32_t pid, tstep;
  // FPB_RNG_PHILOX: cache of the last generated block of 4 normals
  int cblk;
  float4 cn;

  // index stream: replaces ran3 (src/advance.f90:153, src/initialize.f90:68)
  __device__ __forceinline__ float uniform(uint32_t stream) const {
    const uint4 r = philox4x32_10(make_uint4(pid, tstep, stream, 0u), key);
    return u01(r.x);
  }
  // Fortran rannumb(i)
  __device__ __forceinline__ float get(int i) {
    if (mode != FPB_RNG_PHILOX) return __ldg(tab + (i - 1));
    const int blk = i >> 2;
    if (blk != cblk) {
      cn = philox_normals4(pid, tstep, (uint32_t)blk, key);
      cblk = blk;
    }
    const int k = i & 3;
    return k == 0 ? cn.x : k == 1 ? cn.y : k == 2 ? cn.z : cn.w;
  }
};

// ------------------------------------------------------------- gathers ----
struct Lev { // one cached profile level (uprof(n) ... wsigprof(n))
  float u, v, w, rho, rhograd, usig, vsig, wsig;
};

struct Hz { // horizontal + temporal interpolation weights (interpol_mod)
  float p1, p2, p3, p4, dt1, dt2, dtt;
  int o00, o10, o01, o11; // 2-D offsets of the 4 corners
  int plane;              // elements per level of the grid in use
  int ngrid;
};

__device__ __forceinline__ void make_weights(const DevCfg &c, Hz &z, int itime_eff,
                                             float xt, float yt, int ix, int jy,
                                             int ixp, int jyp, int nxd, int nyd) {
  // src/interpol_all.f90:57-71 (identical in interpol_wind, interpol_wind_short)
  float ddx = xt - (float)ix, ddy = yt - (float)jy;
  float rddx = 1.f - ddx, rddy = 1.f - ddy;
  z.p1 = rddx * rddy;
  z.p2 = ddx * rddy;
  z.p3 = rddx * ddy;
  z.p4 = ddx * ddy;
  z.dt1 = (float)(itime_eff - c.memtime[0]);
  z.dt2 = (float)(c.memtime[1] - itime_eff);
  z.dtt = 1.f / (z.dt1 + z.dt2);
  z.o00 = ix + nxd * jy;
  z.o10 = ixp + nxd * jy;
  z.o01 = ix + nxd * jyp;
  z.o11 = ixp + nxd * jyp;
  z.plane = nxd * nyd;
}
__device__ __forceinline__ void make_weights(const DevCfg &c, Hz &z, int itime_eff,
                                             float xt, float yt, int ix, int jy,
                                             int ixp, int jyp) {
  make_weights(c, z, itime_eff, xt, yt, ix, jy, ixp, jyp, c.nxd, c.nyd);
}

// warp-uniform test in front of the polar (uupol/vvpol) loads: ptxas predicates them otherwise
// and every lane pays for loads only particles poleward of +-75 deg need
__device__ __forceinline__ bool warp_any_polar(const Hz &z) {
  return __any_sync(__activemask(), z.ngrid < 0);
}

__device__ __forceinline__ float bil(const Hz &z, float a, float b, float c, float d) {
  return z.p1 * a + z.p2 * b + z.p3 * c + z.p4 * d;
}

// first Fortran level index i in [2,nz] with height(i) > zt, minus 1
// (the linear searches at src/interpol_all.f90:118-125 etc.; height is
// strictly increasing so bisection returns the same index)
__device__ __forceinline__ int find_indz(const float *sh, int nz, float zt) {
  int lo = 2, hi = nz;
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (sh[mid - 1] > zt) hi = mid; else lo = mid + 1;
  }
  return lo - 1;
}

// same index, starting from a hint (the level of the previous search of this particle):
// one test when the particle is still between the same two levels
__device__ __forceinline__ int find_indz_near(const float *sh, int nz, float zt, int hint) {
  if (hint >= 1 && hint < nz && sh[hint] > zt && (hint == 1 || sh[hint - 1] <= zt)) return hint;
  return find_indz(sh, nz, zt);
}

// one profile level: src/interpol_all.f90:135-238 == src/interpol_misslev.f90:56-157.
// SIGMA=false leaves usigprof/vsigprof/wsigprof out (fpb_pbl_kernel: they are
// only read once, after the sub-step loop, and fpb_finish_kernel recomputes
// them for the final level pair with profile_sigma).
template <bool SIGMA>
__device__ __forceinline__ void profile_level(const DevCfg &c, const DevMetSlot *met,
                                              const Hz &z, int n, Lev &L) {
  const int base = (n - 1) * z.plane;
  const bool polar_any = true; // (a vote inside the sub-step kernel's divergent gather loop costs 3 %)
  float y1[2], y2[2], y3[2], r1[2], g1[2];
  float usl = 0.f, vsl = 0.f, wsl = 0.f, usq = 0.f, vsq = 0.f, wsq = 0.f;
#ifdef FPB_LEVEL_SLOT_LOOP
#pragma unroll 1
#else
#pragma unroll
#endif
  for (int m = 0; m < 2; m++) {
    const float4 *A = met[m].A + base;
    const float *G = met[m].G + base;
    float4 Aa = __ldg(A + z.o00), Ab = __ldg(A + z.o10), Ac = __ldg(A + z.o01), Ad = __ldg(A + z.o11);
    const float ga = __ldg(G + z.o00), gb = __ldg(G + z.o10), gc = __ldg(G + z.o01), gd = __ldg(G + z.o11);
    float ua, ub, uc, ud, va, vb, vc, vd;
    if (polar_any && z.ngrid < 0) {
      const float2 *P = met[m].P + base;
      const float2 Pa = __ldg(P + z.o00), Pb = __ldg(P + z.o10), Pc = __ldg(P + z.o01), Pd = __ldg(P + z.o11);
      ua = Pa.x; ub = Pb.x; uc = Pc.x; ud = Pd.x;
      va = Pa.y; vb = Pb.y; vc = Pc.y; vd = Pd.y;
    } else {
      ua = Aa.x; ub = Ab.x; uc = Ac.x; ud = Ad.x;
      va = Aa.y; vb = Ab.y; vc = Ac.y; vd = Ad.y;
    }
    y1[m] = bil(z, ua, ub, uc, ud);
    y2[m] = bil(z, va, vb, vc, vd);
    y3[m] = bil(z, Aa.z, Ab.z, Ac.z, Ad.z);
    g1[m] = bil(z, ga, gb, gc, gd);
    r1[m] = bil(z, Aa.w, Ab.w, Ac.w, Ad.w);
    if (SIGMA) {
      usl = usl + ua + ub + uc + ud;
      vsl = vsl + va + vb + vc + vd;
      usq = usq + ua * ua + ub * ub + uc * uc + ud * ud;
      vsq = vsq + va * va + vb * vb + vc * vc + vd * vd;
      wsl = wsl + Aa.z + Ab.z + Ac.z + Ad.z;
      wsq = wsq + Aa.z * Aa.z + Ab.z * Ab.z + Ac.z * Ac.z + Ad.z * Ad.z;
    }
  }
  L.u = (y1[0] * z.dt2 + y1[1] * z.dt1) * z.dtt;
  L.v = (y2[0] * z.dt2 + y2[1] * z.dt1) * z.dtt;
  L.w = (y3[0] * z.dt2 + y3[1] * z.dt1) * z.dtt;
  L.rho = (r1[0] * z.dt2 + r1[1] * z.dt1) * z.dtt;
  L.rhograd = (g1[0] * z.dt2 + g1[1] * z.dt1) * z.dtt;
  if (SIGMA) {
    float xaux = usq - usl * usl / 8.f;
    L.usig = (xaux < EPS_SIG) ? 0.f : m_sqrt(xaux / 7.f);
    xaux = vsq - vsl * vsl / 8.f;
    L.vsig = (xaux < EPS_SIG) ? 0.f : m_sqrt(xaux / 7.f);
    xaux = wsq - wsl * wsl / 8.f;
    L.wsig = (xaux < EPS_SIG) ? 0.f : m_sqrt(xaux / 7.f);
  }
}

// usigprof(n), vsigprof(n), wsigprof(n) alone: the standard deviations over the
// 8 surrounding values, src/interpol_all.f90:155-238 (same sums, same order)
__device__ __forceinline__ void profile_sigma(const DevCfg &c, const DevMetSlot *met, const Hz &z,
                                              int n, float &usig, float &vsig, float &wsig) {
  const int base = (n - 1) * z.plane;
  const bool polar_any = warp_any_polar(z);
  float usl = 0.f, vsl = 0.f, wsl = 0.f, usq = 0.f, vsq = 0.f, wsq = 0.f;
#pragma unroll
  for (int m = 0; m < 2; m++) {
    const float4 *A = met[m].A + base;
    float4 Aa = __ldg(A + z.o00), Ab = __ldg(A + z.o10), Ac = __ldg(A + z.o01), Ad = __ldg(A + z.o11);
    float ua, ub, uc, ud, va, vb, vc, vd;
    if (polar_any && z.ngrid < 0) {
      const float2 *P = met[m].P + base;
      const float2 Pa = __ldg(P + z.o00), Pb = __ldg(P + z.o10), Pc = __ldg(P + z.o01), Pd = __ldg(P + z.o11);
      ua = Pa.x; ub = Pb.x; uc = Pc.x; ud = Pd.x;
      va = Pa.y; vb = Pb.y; vc = Pc.y; vd = Pd.y;
    } else {
      ua = Aa.x; ub = Ab.x; uc = Ac.x; ud = Ad.x;
      va = Aa.y; vb = Ab.y; vc = Ac.y; vd = Ad.y;
    }
    usl = usl + ua + ub + uc + ud;
    vsl = vsl + va + vb + vc + vd;
    usq = usq + ua * ua + ub * ub + uc * uc + ud * ud;
    vsq = vsq + va * va + vb * vb + vc * vc + vd * vd;
    wsl = wsl + Aa.z + Ab.z + Ac.z + Ad.z;
    wsq = wsq + Aa.z * Aa.z + Ab.z * Ab.z + Ac.z * Ac.z + Ad.z * Ad.z;
  }
  float xaux = usq - usl * usl / 8.f;
  usig = (xaux < EPS_SIG) ? 0.f : m_sqrt(xaux / 7.f);
  xaux = vsq - vsl * vsl / 8.f;
  vsig = (xaux < EPS_SIG) ? 0.f : m_sqrt(xaux / 7.f);
  xaux = wsq - wsl * wsl / 8.f;
  wsig = (xaux < EPS_SIG) ? 0.f : m_sqrt(xaux / 7.f);
}

#ifndef FPB_PREFETCH_WIND
#define FPB_PREFETCH_WIND 0 // 0: off (measured: the 16 prefetches cost more LSU issue than the latency they hide, +15 % on the C5 finish kernel); 1: into L1, 2: into L2
#endif
__device__ __forceinline__ void prefetch_gl(const void *p) {
#if FPB_PREFETCH_WIND == 2
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
#else
  asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
#endif
}

__device__ __forceinline__ void ldg_pair(const MetPair *p, float4 &a, float4 &b) { // LDG.E.256
  asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w)
               : "l"(p));
}

// src/interpol_wind.f90:56-214 (SIGMA=true) / src/interpol_wind_short.f90:48-140
template <bool SIGMA>
__device__ __forceinline__ void interp_wind(const DevCfg &c, const DevMetSlot *met,
                                            const Hz &z, const float *sh, float zt,
                                            float &u, float &v, float &w, float &usig,
                                            float &vsig, float &wsig, int &indz_io) {
  // indz_io: in, a hint (0 = none); out, the level found
  const int indz = indz_io = find_indz_near(sh, c.nz, zt, indz_io);
  const float dz = 1.f / (sh[indz] - sh[indz - 1]);
  const float dz1 = (zt - sh[indz - 1]) * dz;
  const float dz2 = (sh[indz] - zt) * dz;
  const int plane = z.plane;
  const bool polar_any = warp_any_polar(z);
  // x-neighbour pairs in one aligned 32-byte word (MetPair, FPB_MET_PAIRS=1): measured SLOWER than the
  // two 16-byte loads (C5 finish kernel +5 %: twice the L2 footprint for the same sectors), so the pair
  // arrays are not built by default.  The two-way branch itself stays: with it ptxas issues the four loads
  // of one (time, level) pair back to back and consumes them before the next four, which is 3-4 % faster
  // than the fully hoisted schedule it picks for the branch-free loop (gpurun_out/ab_np.txt; compiler
  // barriers or `#pragma unroll 1` do not reproduce it: ab_b.txt, ab_u.txt).
#ifndef FPB_MET_PAIRS_CODE
#define FPB_MET_PAIRS_CODE 1
#endif
#ifndef FPB_WIND_BARRIER
#define FPB_WIND_BARRIER 0
#endif
#ifndef FPB_WIND_UNROLL
#define FPB_WIND_UNROLL 0
#endif
#if FPB_MET_PAIRS_CODE
  const bool pairs = met[0].AP != nullptr && z.o10 == z.o00 + 1 && z.o11 == z.o01 + 1;
#else
  constexpr bool pairs = false;
#endif
#if FPB_PREFETCH_WIND
  // The 16 corner words are independent, but at 64 registers the compiler keeps only four float4 loads in
  // flight: four dependent rounds of DRAM latency.  Prefetches need no destination register, so all 16
  // go out at once and the loads below find their sectors on the way (or in the cache).
#pragma unroll
  for (int m = 0; m < 2; m++) {
#pragma unroll
    for (int n = 0; n < 2; n++) {
      const float4 *A = met[m].A + (indz - 1 + n) * plane;
      prefetch_gl(A + z.o00); prefetch_gl(A + z.o10); prefetch_gl(A + z.o01); prefetch_gl(A + z.o11);
    }
  }
#endif
  float uh[2], vh[2], wh[2];
  float usl = 0.f, vsl = 0.f, wsl = 0.f, usq = 0.f, vsq = 0.f, wsq = 0.f;
#if FPB_WIND_UNROLL == 1
#pragma unroll 1
#else
#pragma unroll
#endif
  for (int m = 0; m < 2; m++) {
    float u1[2], v1[2], w1[2];
#if FPB_WIND_UNROLL == 2
#pragma unroll 1
#else
#pragma unroll
#endif
    for (int n = 0; n < 2; n++) {
      const int base = (indz - 1 + n) * plane;
      const float4 *A = met[m].A + base;
      float4 Aa, Ab, Ac, Ad;
      if (pairs) { // x-neighbours in one aligned 32-byte word: half the load instructions and sector requests
        ldg_pair(met[m].AP + base + z.o00, Aa, Ab);
        ldg_pair(met[m].AP + base + z.o01, Ac, Ad);
      } else {
        Aa = __ldg(A + z.o00); Ab = __ldg(A + z.o10); Ac = __ldg(A + z.o01); Ad = __ldg(A + z.o11);
      }
      float ua, ub, uc, ud, va, vb, vc, vd;
      if (polar_any && z.ngrid < 0) {
        const float2 *P = met[m].P + base;
        const float2 Pa = __ldg(P + z.o00), Pb = __ldg(P + z.o10), Pc = __ldg(P + z.o01), Pd = __ldg(P + z.o11);
        ua = Pa.x; ub = Pb.x; uc = Pc.x; ud = Pd.x;
        va = Pa.y; vb = Pb.y; vc = Pc.y; vd = Pd.y;
      } else {
        ua = Aa.x; ub = Ab.x; uc = Ac.x; ud = Ad.x;
        va = Aa.y; vb = Ab.y; vc = Ac.y; vd = Ad.y;
      }
      u1[n] = bil(z, ua, ub, uc, ud);
      v1[n] = bil(z, va, vb, vc, vd);
      w1[n] = bil(z, Aa.z, Ab.z, Ac.z, Ad.z);
      if (SIGMA) {
        usl = usl + ua + ub + uc + ud;
        vsl = vsl + va + vb + vc + vd;
        usq = usq + ua * ua + ub * ub + uc * uc + ud * ud;
        vsq = vsq + va * va + vb * vb + vc * vc + vd * vd;
        wsl = wsl + Aa.z + Ab.z + Ac.z + Ad.z;
        wsq = wsq + Aa.z * Aa.z + Ab.z * Ab.z + Ac.z * Ac.z + Ad.z * Ad.z;
      }
#if FPB_WIND_BARRIER == 1
      asm volatile("" ::: "memory");
#endif
    }
#if FPB_WIND_BARRIER == 2
    asm volatile("" ::: "memory");
#endif
    uh[m] = dz2 * u1[0] + dz1 * u1[1];
    vh[m] = dz2 * v1[0] + dz1 * v1[1];
    wh[m] = dz2 * w1[0] + dz1 * w1[1];
  }
  u = (uh[0] * z.dt2 + uh[1] * z.dt1) * z.dtt;
  v = (vh[0] * z.dt2 + vh[1] * z.dt1) * z.dtt;
  w = (wh[0] * z.dt2 + wh[1] * z.dt1) * z.dtt;
  if (SIGMA) {
    float xaux = usq - usl * usl / 16.f;
    usig = (xaux < EPS_SIG) ? 0.f : m_sqrt(xaux / 15.f);
    xaux = vsq - vsl * vsl / 16.f;
    vsig = (xaux < EPS_SIG) ? 0.f : m_sqrt(xaux / 15.f);
    xaux = wsq - wsl * wsl / 16.f;
    wsig = (xaux < EPS_SIG) ? 0.f : m_sqrt(xaux / 15.f);
  }
}

// --------------------------------------------------------- turbulence ----
struct Turb { // hanna_mod
  float ust, wst, ol, h, zeta, sigu, sigv, tlu, tlv, tlw, sigw, dsigwdz, dsigw2dz;
};

// src/interpol_all.f90:80-107
__device__ __forceinline__ void interp_surface(const DevMetSlot *met, const Hz &z, Turb &t) {
  float us1[2], ws1[2], ol1[2];
#pragma unroll
  for (int m = 0; m < 2; m++) {
    float4 a = __ldg(met[m].S + z.o00), b = __ldg(met[m].S + z.o10),
           c = __ldg(met[m].S + z.o01), d = __ldg(met[m].S + z.o11);
    us1[m] = bil(z, a.y, b.y, c.y, d.y);
    ws1[m] = bil(z, a.z, b.z, c.z, d.z);
    ol1[m] = bil(z, a.w, b.w, c.w, d.w);
  }
  t.ust = (us1[0] * z.dt2 + us1[1] * z.dt1) * z.dtt;
  t.wst = (ws1[0] * z.dt2 + ws1[1] * z.dt1) * z.dtt;
  float oliaux = (ol1[0] * z.dt2 + ol1[1] * z.dt1) * z.dtt;
  t.ol = (oliaux != 0.f) ? 1.f / oliaux : 99999.f;
}

// sigma_w, d(sigma_w)/dz and T_Lw of the unstable regime, shared by hanna and
// hanna_short (src/hanna.f90:66-88, src/hanna_short.f90:59-77).  The three
// T_Lw candidates are evaluated by every lane and selected (same arithmetic
// per candidate as the reference's if-chain, no divergent passes).
__device__ __forceinline__ void hanna_unstable_w(Turb &t, float z) {
#if FPB_STRICT
  const float z23 = m_pow(t.zeta, 0.66666f);
  const float zm13 = m_pow(fmaxf(t.zeta, 1.e-3f), -.33333f);
#else
  const float lz = __log2f(t.zeta); // one MUFU.LG2 feeds both powers
  const float z23 = exp2f(0.66666f * lz);
  const float zm13 = exp2f(-.33333f * fmaxf(lz, -9.965784284662087f)); // log2(1e-3)
#endif
  t.sigw = m_sqrt(1.2f * (t.wst * t.wst) * (1.f - .9f * t.zeta) * z23 +
                  (1.8f - 1.4f * t.zeta) * (t.ust * t.ust)) + 1.e-2f;
  t.dsigwdz = 0.5f / t.sigw / t.h *
              (-1.4f * (t.ust * t.ust) + (t.wst * t.wst) * (0.8f * zm13 - 1.8f * z23));
  const float tl_a = 0.1f * z / (t.sigw * (0.55f - 0.38f * fabsf(z / t.ol)));
  const float tl_b = 0.59f * z / t.sigw;
  const float tl_c = 0.15f * t.h / t.sigw * (1.f - m_exp(-5.f * t.zeta));
  t.tlw = (z < fabsf(t.ol)) ? tl_a : ((t.zeta < 0.1f) ? tl_b : tl_c);
}

// src/hanna.f90:42-106
// regime: 0 = h/|ol| < 1, 1 = ol < 0, 2 = otherwise (fixed for one advance() call)
__device__ __forceinline__ int hanna_regime(const Turb &t) {
  return (t.h / fabsf(t.ol) < 1.f) ? 0 : ((t.ol < 0.f) ? 1 : 2);
}

__device__ __forceinline__ void hanna(Turb &t, float z, int regime) {
  if (regime == 0) {
    t.ust = fmaxf(1.e-4f, t.ust);
    float corr = z / t.ust;
    t.sigu = 1.e-2f + 2.0f * t.ust * m_exp(-3.e-4f * corr);
    t.sigw = 1.3f * t.ust * m_exp(-2.e-4f * corr);
    t.dsigwdz = -2.e-4f * t.sigw;
    t.sigw = t.sigw + 1.e-2f;
    t.sigv = t.sigw;
    t.tlu = 0.5f * z / t.sigw / (1.f + 1.5e-3f * corr);
    t.tlv = t.tlu;
    t.tlw = t.tlu;
  } else if (regime == 1) {
    t.sigu = 1.e-2f + t.ust * m_pow(12.f - 0.5f * t.h / t.ol, 0.33333f);
    t.sigv = t.sigu;
    hanna_unstable_w(t, z);
    t.tlu = 0.15f * t.h / t.sigu;
    t.tlv = t.tlu;
  } else {
    t.sigu = 1.e-2f + 2.f * t.ust * (1.f - t.zeta);
    t.sigv = 1.e-2f + 1.3f * t.ust * (1.f - t.zeta);
    t.sigw = t.sigv;
    t.dsigwdz = -1.3f * t.ust / t.h;
    t.tlu = 0.15f * t.h / t.sigu * (m_sqrt(t.zeta));
    t.tlv = 0.467f * t.tlu;
    t.tlw = 0.1f * t.h / t.sigw * m_pow(t.zeta, 0.8f);
  }
  t.tlu = fmaxf(10.f, t.tlu);
  t.tlv = fmaxf(10.f, t.tlv);
  t.tlw = fmaxf(30.f, t.tlw);
  if (t.dsigwdz == 0.f) t.dsigwdz = 1.e-10f;
}

// src/hanna_short.f90:42-92
__device__ __forceinline__ void hanna_short(Turb &t, float z, int regime) {
  if (regime == 0) {
    t.ust = fmaxf(1.e-4f, t.ust);
    t.sigw = 1.3f * m_exp(-2.e-4f * z / t.ust);
    t.dsigwdz = -2.e-4f * t.sigw;
    t.sigw = t.sigw * t.ust + 1.e-2f;
    t.tlw = 0.5f * z / t.sigw / (1.f + 1.5e-3f * z / t.ust);
  } else if (regime == 1) {
    hanna_unstable_w(t, z);
  } else {
    t.sigw = 1.e-2f + 1.3f * t.ust * (1.f - t.zeta);
    t.dsigwdz = -1.3f * t.ust / t.h;
    t.tlw = 0.1f * t.h / t.sigw * m_pow(t.zeta, 0.8f);
  }
  t.tlu = fmaxf(10.f, t.tlu);
  t.tlv = fmaxf(10.f, t.tlv);
  t.tlw = fmaxf(30.f, t.tlw);
  if (t.dsigwdz == 0.f) t.dsigwdz = 1.e-10f;
}

// src/hanna1.f90:42-128
__device__ __forceinline__ void hanna1(Turb &t, float z, int regime) {
  if (regime == 0) {
    t.ust = fmaxf(1.e-4f, t.ust);
    t.sigu = 2.0f * t.ust * m_exp(-3.e-4f * z / t.ust);
    t.sigu = fmaxf(t.sigu, 1.e-5f);
    t.sigv = 1.3f * t.ust * m_exp(-2.e-4f * z / t.ust);
    t.sigv = fmaxf(t.sigv, 1.e-5f);
    t.sigw = t.sigv;
    t.dsigw2dz = -6.76e-4f * t.ust * m_exp(-4.e-4f * z / t.ust);
    t.tlu = 0.5f * z / t.sigw / (1.f + 1.5e-3f * z / t.ust);
    t.tlv = t.tlu;
    t.tlw = t.tlu;
  } else if (regime == 1) {
    t.sigu = t.ust * m_pow(12.f - 0.5f * t.h / t.ol, 0.33333f);
    t.sigu = fmaxf(t.sigu, 1.e-6f);
    t.sigv = t.sigu;
    if (t.zeta < 0.03f) {
      t.sigw = 0.96f * t.wst * m_pow(3.f * t.zeta - t.ol / t.h, 0.33333f);
      t.dsigw2dz = 1.8432f * t.wst * t.wst / t.h * m_pow(3.f * t.zeta - t.ol / t.h, -0.33333f);
    } else if (t.zeta < 0.4f) {
      float s1 = 0.96f * m_pow(3.f * t.zeta - t.ol / t.h, 0.33333f);
      float s2 = 0.763f * m_pow(t.zeta, 0.175f);
      if (s1 < s2) {
        t.sigw = t.wst * s1;
        t.dsigw2dz = 1.8432f * t.wst * t.wst / t.h * m_pow(3.f * t.zeta - t.ol / t.h, -0.33333f);
      } else {
        t.sigw = t.wst * s2;
        t.dsigw2dz = 0.203759f * t.wst * t.wst / t.h * m_pow(t.zeta, -0.65f);
      }
    } else if (t.zeta < 0.96f) {
      t.sigw = 0.722f * t.wst * m_pow(1.f - t.zeta, 0.207f);
      t.dsigw2dz = -.215812f * t.wst * t.wst / t.h * m_pow(1.f - t.zeta, -0.586f);
    } else { // zeta in [0.96,1]; the reference leaves zeta == 1 undefined
      t.sigw = 0.37f * t.wst;
      t.dsigw2dz = 0.f;
    }
    t.sigw = fmaxf(t.sigw, 1.e-6f);
    t.tlu = 0.15f * t.h / t.sigu;
    t.tlv = t.tlu;
    if (z < fabsf(t.ol))
      t.tlw = 0.1f * z / (t.sigw * (0.55f - 0.38f * fabsf(z / t.ol)));
    else if (t.zeta < 0.1f)
      t.tlw = 0.59f * z / t.sigw;
    else
      t.tlw = 0.15f * t.h / t.sigw * (1.f - m_exp(-5.f * t.zeta));
  } else {
    t.sigu = 2.f * t.ust * (1.f - t.zeta);
    t.sigv = 1.3f * t.ust * (1.f - t.zeta);
    t.sigu = fmaxf(t.sigu, 1.e-6f);
    t.sigv = fmaxf(t.sigv, 1.e-6f);
    t.sigw = t.sigv;
    t.dsigw2dz = 3.38f * t.ust * t.ust * (t.zeta - 1.f) / t.h;
    t.tlu = 0.15f * t.h / t.sigu * (m_sqrt(t.zeta));
    t.tlv = 0.467f * t.tlu;
    t.tlw = 0.1f * t.h / t.sigw * m_pow(t.zeta, 0.8f);
  }
  t.tlu = fmaxf(10.f, t.tlu);
  t.tlv = fmaxf(10.f, t.tlv);
  t.tlw = fmaxf(30.f, t.tlw);
}

// src/windalign.f90:36-54
__device__ __forceinline__ void windalign(float u, float v, float ffap, float ffcp,
                                          float &ux, float &vy) {
  float ffinv = 1.f / fmaxf(m_sqrt(u * u + v * v), 1.e-30f);
  float sinphi = v * ffinv;
  float vy1 = sinphi * ffap;
  float cosphi = u * ffinv;
  float ux1 = cosphi * ffap;
  float ux2 = -sinphi * ffcp;
  float vy2 = cosphi * ffcp;
  ux = ux1 + ux2;
  vy = vy1 + vy2;
}

// ------------------------------------------------- conformal map (poles) ----
// Taylor's CMAPF transformations as used by FLEXPART poleward of +-75 deg
// (src/cmapf_mod.f90: cspanf :494, cnllxy :310, cnxyll :367, cll2xy :295,
// cxy2ll :526, cgszll :190).  Which sub-expressions run in double follows
// the reference's declarations.
constexpr float CM_REARTH = 6371.2f, CM_ALMST1 = .9999999f;
constexpr float CM_PI = 3.14159265358979f;
constexpr float CM_RADPDG = CM_PI / 180.f, CM_DGPRAD = 180.f / CM_PI;

__device__ __forceinline__ float cspanf(float value, float begin, float end) {
  float first = fminf(begin, end), last = fmaxf(begin, end);
  float val = fmodf(value - first, last - first);
  return (val <= 0.f) ? val + last : val + first;
}

__device__ __noinline__ void cnllxy(const float *sc, float xlat, float xlong, float &xi, float &eta) {
  double gamma = sc[0];
  double dlat = xlat;
  double dlong = cspanf(xlong - sc[1], -180.f, 180.f);
  dlong = dlong * CM_RADPDG;
  float gdlong = (float)(gamma * dlong), sndgam, csdgam, rhog1;
  if (fabsf(gdlong) < .01f) {
    gdlong = gdlong * gdlong;
    sndgam = (float)(dlong * (1.f - 1.f / 6.f * gdlong * (1.f - 1.f / 20.f * gdlong * (1.f - 1.f / 42.f * gdlong))));
    csdgam = (float)(dlong * dlong * .5f * (1.f - 1.f / 12.f * gdlong * (1.f - 1.f / 30.f * gdlong * (1.f - 1.f / 56.f * gdlong))));
  } else {
    sndgam = (float)(m_sin(gdlong) / gamma);
    csdgam = (float)((1.f - m_cos(gdlong)) / gamma / gamma);
  }
  double slat = sin(CM_RADPDG * dlat);
  if ((slat >= CM_ALMST1) || (slat <= -CM_ALMST1)) {
    eta = 1.f / sc[0];
    xi = 0.f;
    return;
  }
  double mercy = .5f * log((1.f + slat) / (1.f - slat));
  double gmercy = gamma * mercy;
  if (fabs(gmercy) < .001f)
    rhog1 = (float)(mercy * (1.f - .5f * gmercy * (1.f - 1.f / 3.f * gmercy * (1.f - 1.f / 4.f * gmercy))));
  else
    rhog1 = (float)((1.f - exp(-gmercy)) / gamma);
  eta = (float)(rhog1 + (1.f - gamma * rhog1) * gamma * csdgam);
  xi = (float)((1.f - gamma * rhog1) * sndgam);
}

__device__ __noinline__ void cnxyll(const float *sc, double xi, double eta, float &xlat, float &xlong) {
  double gamma = sc[0], temp, ymerc, along;
  double arg2 = 2.f * eta - gamma * (xi * xi + eta * eta);
  double arg1 = gamma * arg2;
  if (fabs(arg1) < .01f) {
    temp = (arg1 / (2.f - arg1)) * (arg1 / (2.f - arg1));
    ymerc = arg2 / (2.f - arg1) * (1.f + temp * (1.f / 3.f + temp * (1.f / 5.f + temp * (1.f / 7.f))));
  } else {
    ymerc = -log(1.f - arg1) / 2.f / gamma;
  }
  temp = exp(-fabs(ymerc));
  xlat = (float)copysign(atan2((1.f - temp) * (1.f + temp), 2.f * temp), ymerc);
  double gxi = gamma * xi, cgeta = 1.f - gamma * eta;
  if (fabs(gxi) < .01f * cgeta) {
    temp = (gxi / cgeta) * (gxi / cgeta);
    along = xi / cgeta * (1.f - temp * (1.f / 3.f - temp * (1.f / 5.f - temp * (1.f / 7.f))));
  } else {
    along = atan2(gxi, cgeta) / gamma;
  }
  xlong = (float)(sc[1] + CM_DGPRAD * along);
  xlat = xlat * CM_DGPRAD;
}

__device__ __noinline__ void cll2xy(const float *sc, float xlat, float xlong, float &x, float &y) {
  float xi, eta;
  cnllxy(sc, xlat, xlong, xi, eta);
  x = sc[2] + CM_REARTH / sc[6] * (xi * sc[4] + eta * sc[5]);
  y = sc[3] + CM_REARTH / sc[6] * (eta * sc[4] - xi * sc[5]);
}

__device__ __noinline__ void cxy2ll(const float *sc, float x, float y, float &xlat, float &xlong) {
  double xi0 = (x - sc[2]) * sc[6] / CM_REARTH;
  double eta0 = (y - sc[3]) * sc[6] / CM_REARTH;
  double xi = xi0 * sc[4] - eta0 * sc[5];
  double eta = eta0 * sc[4] + xi0 * sc[5];
  cnxyll(sc, xi, eta, xlat, xlong);
  xlong = cspanf(xlong, -180.f, 180.f);
}

__device__ __noinline__ float cgszll(const float *sc, float xlat) {
  double slat, ymerc, efact;
  if (xlat > 89.985f) {
    if (sc[0] > 0.9999f) return 2.f * sc[6];
    efact = m_cos(CM_RADPDG * xlat);
    if (efact <= 0.) return 0.f;
    ymerc = -log(efact / (1.f + m_sin(CM_RADPDG * xlat)));
  } else if (xlat < -89.985f) {
    if (sc[0] < -0.9999f) return 2.f * sc[6];
    efact = m_cos(CM_RADPDG * xlat);
    if (efact <= 0.) return 0.f;
    ymerc = log(efact / (1.f - m_sin(CM_RADPDG * xlat)));
  } else {
    slat = m_sin(CM_RADPDG * xlat);
    ymerc = log((1.f + slat) / (1.f - slat)) / 2.f;
  }
  return (float)(sc[6] * m_cos(CM_RADPDG * xlat) * exp(sc[0] * ymerc));
}

// position update: src/advance.f90:750-778 (tfac = ldirect) and :923-951
// (tfac = ldt*ldirect)
__device__ __forceinline__ void move_horizontal(const DevCfg &c, int ngrid, double &xt,
                                                double &yt, float dxm, float dym, float tfac) {
  if (ngrid >= 0) {
#if FPB_STRICT
    float cosfact = (float)(c.dxconst / cos((yt * c.dy + c.ylat0) * PI180));
#else
    // fast: latitude formed in double, cosine in float (|lat| < 75 deg here: relative error < 5e-7
    // of one step's displacement)
    float cosfact = c.dxconst / cosf((float)((yt * c.dy + c.ylat0) * PI180));
#endif
    xt = xt + (double)(dxm * cosfact * tfac);
    yt = yt + (double)(dym * c.dyconst * tfac);
  } else {
    const float *map = (ngrid == -1) ? c.northpolemap : c.southpolemap;
    float xlon = (float)(c.xlon0 + xt * c.dx);
    float ylat = (float)(c.ylat0 + yt * c.dy);
    float xpol, ypol;
    cll2xy(map, ylat, xlon, xpol, ypol);
    float gridsize = 1000.f * cgszll(map, ylat);
    dxm = dxm / gridsize;
    dym = dym / gridsize;
    xpol = xpol + dxm * tfac;
    ypol = ypol + dym * tfac;
    cxy2ll(map, xpol, ypol, ylat, xlon);
    xt = (xlon - c.xlon0) / c.dx;
    yt = (ylat - c.ylat0) / c.dy;
  }
}

// cyclic boundary, pole crossing, domain exit: src/advance.f90:784-808
__device__ __forceinline__ bool wrap_and_check(const DevCfg &c, double &xt, double &yt) {
  const float eps = c.eps;
  if (c.xglobal) {
    if (xt >= (float)c.nxmin1) xt = xt - (float)c.nxmin1;
    if (xt < 0.) xt = xt + (float)c.nxmin1;
    if (xt <= eps) xt = eps;
    if (fabs(xt - (float)c.nxmin1) <= eps) xt = (float)c.nxmin1 - eps;
    if (yt < 0.) {
      xt = d_modulo(xt * c.dx + 180.f, 360.f) / c.dx;
      yt = -yt;
    } else if (yt > (float)c.nymin1) {
      xt = d_modulo(xt * c.dx + 180.f, 360.f) / c.dx;
      yt = 2 * (float)c.nymin1 - yt;
    }
  }
  return (xt < 0.) || (xt >= (float)c.nxmin1) || (yt < 0.) || (yt > (float)c.nymin1);
}

__device__ __forceinline__ int pole_grid(const DevCfg &c, double yt) {
  if (c.nglobal && (yt > c.switchnorthg)) return -1;
  if (c.sglobal && (yt < c.switchsouthg)) return -2;
  return 0;
}

// grid choice incl. the nesting level, src/advance.f90:161-175 (= :841-856)
__device__ __forceinline__ int choose_grid(const DevCfg &c, double xt, double yt) {
  const int ngrid = pole_grid(c, yt);
  if (ngrid != 0) return ngrid;
  for (int j = c.numbnests; j >= 1; j--)
    if ((xt > c.xln[j - 1] + c.eps) && (xt < c.xrn[j - 1] - c.eps) &&
        (yt > c.yln[j - 1] + c.eps) && (yt < c.yrn[j - 1] - c.eps))
      return j;
  return 0;
}

// The grid a call interpolates from: the mother grid, or nested grid ngrid
// with the particle at xtn = (xt-xln)*xresoln, ytn = (yt-yln)*yresoln
// (src/advance.f90:191-203).  The *_nests routines of the reference are the
// mother-grid routines on the nest's arrays (src/interpol_all_nests.f90 etc.).
struct GridSel {
  const DevMetSlot *met;  // [2]: memind(1), memind(2)
  const float *trop_lit1; // tropopause of Fortran slot 1 (literal in the reference)
  int nxd, nyd;
  float xf, yf;           // position in that grid's coordinates, real (f32)
  int ix, jy, nix, njy;
};

__device__ __forceinline__ GridSel select_grid(const DevStepArgs &a, int ngrid, double xt, double yt) {
  const DevCfg &c = a.cfg;
  GridSel g;
  if (ngrid > 0) {
    const int l = ngrid - 1;
    g.met = a.metn[l];
    g.trop_lit1 = a.tropn_lit1[l];
    g.nxd = c.nxdn[l];
    g.nyd = c.nydn[l];
    g.xf = (float)((xt - c.xln[l]) * c.xresoln[l]);
    g.yf = (float)((yt - c.yln[l]) * c.yresoln[l]);
    g.ix = f_int(g.xf);
    g.jy = f_int(g.yf);
    g.nix = (int)roundf(g.xf);
    g.njy = (int)roundf(g.yf);
  } else {
    g.met = a.met;
    g.trop_lit1 = a.met_lit1.trop;
    g.nxd = c.nxd;
    g.nyd = c.nyd;
    g.xf = (float)xt;
    g.yf = (float)yt;
    g.ix = d_int(xt);
    g.jy = d_int(yt);
    g.nix = d_nint(xt);
    g.njy = d_nint(yt);
  }
  return g;
}

// ------------------------------------------------------------ settling ----
// src/get_settling.f90:52-125 + dynamic_viscosity.f90; rho/tt from Fortran
// slot 1 (literal in the reference)
__device__ __noinline__ float get_settling(const DevCfg &c, const DevMetSlot &lit1, const float *sh,
                              float xt, float yt, float zt, int nsp) {
  const float ga = 9.81f;
  int nix = f_int(xt), njy = f_int(yt);
  int indz = find_indz(sh, c.nz, zt);
  float dz = 1.f / (sh[indz] - sh[indz - 1]);
  float dz1 = (zt - sh[indz - 1]) * dz;
  float dz2 = (sh[indz] - zt) * dz;
  int plane = c.nxd * c.nyd, o = nix + c.nxd * njy;
  float4 A0 = __ldg(lit1.A + (indz - 1) * plane + o), A1 = __ldg(lit1.A + indz * plane + o);
  const float t0 = __ldg(lit1.T + (indz - 1) * plane + o), t1 = __ldg(lit1.T + indz * plane + o);
  float temperature = dz2 * t0 + dz1 * t1;
  float airdens = dz2 * A0.w + dz1 * A1.w;
  const float cc = 120.f, t_0 = 291.15f, eta_0 = 1.827e-5f;
  float vis_dyn = eta_0 * (t_0 + cc) / (temperature + cc) * m_pow(temperature / t_0, 1.5f);
  float vis_kin = vis_dyn / airdens;
  float dq = c.dquer[nsp - 1], vsa = c.vsetaver[nsp - 1];
  float reynolds = dq / 1.e6f * fabsf(vsa) / vis_kin;
  float settling_old = vsa, settling = 0.f, c_d;
  for (int i = 1; i <= 20; i++) {
    if (reynolds < 1.917f) c_d = 24.f / reynolds;
    else if (reynolds < 500.f) c_d = 18.5f / m_pow(reynolds, 0.6f);
    else c_d = 0.44f;
    settling = -1.f * m_sqrt(4.f * ga * dq / 1.e6f * c.density[nsp - 1] * c.cunningham[nsp - 1] /
                             (3.f * c_d * airdens));
    if (fabsf((settling - settling_old) / settling) < 0.01f) break;
    reynolds = dq / 1.e6f * fabsf(settling) / vis_kin;
    settling_old = settling;
  }
  return settling;
}

// species choice + call, src/advance.f90:518-531
__device__ __forceinline__ float settling_term(const DevStepArgs &a, const float *sh,
                                               int nrelpoint, double xt, double yt, float zt) {
  const DevCfg &c = a.cfg;
  if (c.mdomainfill != 0 || !c.lsettling) return 0.f;
  int nsp;
  for (nsp = 1; nsp <= c.nspec; nsp++)
    if (__ldg(a.xmass + (nsp - 1) * c.numpoint + (nrelpoint - 1)) > EPS3) break;
  if (nsp > c.nspec) nsp = c.nspec;
  if (c.density[nsp - 1] > 0.f)
    return get_settling(c, a.met_lit1, sh, (float)xt, (float)yt, zt, nsp);
  return 0.f;
}

// ------------------------------------------------------------------ CBL ----
#include "fpb_cbl.cuh"

// --------------------------------------------------------- deposition ----
// src/drydepokernel.f90:41-116 and drydepokernel_nest.f90
__device__ __forceinline__ size_t didx(const DevCfg &c, int nxg, int nyg, int ix, int jy,
                                       int ks, int kp, int nc, int na) {
  size_t i = (size_t)(na - 1);
  i = i * c.nclassunc + (nc - 1);
  i = i * c.maxpointspec_act + (kp - 1);
  i = i * c.nspec + (ks - 1);
  i = i * nyg + jy;
  i = i * nxg + ix;
  return i;
}

// species-free cell key of a deposition grid: like cell_key() without the level
__device__ __forceinline__ unsigned dep_key(const DevCfg &c, int nxg, int nyg, int ix, int jy, int kp, int nc, int na) {
  unsigned i = (unsigned)(na - 1);
  i = i * c.nclassunc + (nc - 1);
  i = i * c.maxpointspec_act + (kp - 1);
  i = i * nyg + jy;
  i = i * nxg + ix;
  return i;
}
// one record of the deterministic path: the nspec values of corner `corner` of record slot `rslot`
__device__ __forceinline__ void dep_record(const DevCfg &c, const DevDepRecords &r, int g, int rslot, int corner,
                                           unsigned key, const float *deposit, float w, const int *specmask) {
  const size_t id = 4 * (size_t)rslot + corner;
  r.keys[g][id] = key;
  for (int ks = 0; ks < c.nspec; ks++)
    r.vals[g][(size_t)ks * r.nrec + id] = (specmask == nullptr || specmask[ks]) ? deposit[ks] * w : 0.f;
}

__device__ __noinline__ void drydepo_scatter(const DevCfg &c, float *grid, bool nest, int nunc,
                                const float *deposit, float x, float y, int nage, int kp,
                                const DevDepRecords &rec, int rslot) {
  const int nxg = nest ? c.numxgridn : c.numxgrid, nyg = nest ? c.numygridn : c.numygrid;
  float xl, yl;
  if (nest) {
    xl = (x * c.dx + c.xoutshiftn) / c.dxoutn;
    yl = (y * c.dy + c.youtshiftn) / c.dyoutn;
  } else {
    xl = (x * c.dx + c.xoutshift) / c.dxout;
    yl = (y * c.dy + c.youtshift) / c.dyout;
  }
  int ix = f_int(xl), jy = f_int(yl), ixp, jyp;
  float ddx = xl - (float)ix, ddy = yl - (float)jy, wx, wy;
  if (ddx > 0.5f) { ixp = ix + 1; wx = 1.5f - ddx; } else { ixp = ix - 1; wx = 0.5f + ddx; }
  if (ddy > 0.5f) { jyp = jy + 1; wy = 1.5f - ddy; } else { jyp = jy - 1; wy = 0.5f + ddy; }
  const bool in00 = (ix >= 0) && (jy >= 0) && (ix <= nxg - 1) && (jy <= nyg - 1);
  const bool in11 = (ixp >= 0) && (jyp >= 0) && (ixp <= nxg - 1) && (jyp <= nyg - 1);
  const bool in10 = (ixp >= 0) && (jy >= 0) && (ixp <= nxg - 1) && (jy <= nyg - 1);
  const bool in01 = (ix >= 0) && (jyp >= 0) && (ix <= nxg - 1) && (jyp <= nyg - 1);
  if (rec.keys[0]) { // deterministic: records, added in particle order after the kernel
    const int g = nest ? 1 : 0;
    float dep[FPB_MAXSPEC];
    bool any = false;
    for (int ks = 0; ks < c.nspec; ks++) {
      dep[ks] = ((fabsf(deposit[ks]) > 0.f) && c.drydepspec[ks]) ? deposit[ks] : 0.f;
      any = any || dep[ks] != 0.f;
    }
    if (!any) return;
    if (!nest && !c.lusekerneloutput) {
      if (in00) dep_record(c, rec, g, rslot, 0, dep_key(c, nxg, nyg, ix, jy, kp, nunc, nage), dep, 1.f, nullptr);
      return;
    }
    if (in00) dep_record(c, rec, g, rslot, 0, dep_key(c, nxg, nyg, ix, jy, kp, nunc, nage), dep, wx * wy, nullptr);
    if (in11) dep_record(c, rec, g, rslot, 1, dep_key(c, nxg, nyg, ixp, jyp, kp, nunc, nage), dep, (1.f - wx) * (1.f - wy), nullptr);
    if (in10) dep_record(c, rec, g, rslot, 2, dep_key(c, nxg, nyg, ixp, jy, kp, nunc, nage), dep, (1.f - wx) * wy, nullptr);
    if (in01) dep_record(c, rec, g, rslot, 3, dep_key(c, nxg, nyg, ix, jyp, kp, nunc, nage), dep, wx * (1.f - wy), nullptr);
    return;
  }
  for (int ks = 1; ks <= c.nspec; ks++) {
    float dep = deposit[ks - 1];
    if (!((fabsf(dep) > 0.f) && c.drydepspec[ks - 1])) continue;
    if (!nest && !c.lusekerneloutput) {
      if (in00) atomicAdd(grid + didx(c, nxg, nyg, ix, jy, ks, kp, nunc, nage), dep);
      continue;
    }
    if (in00) atomicAdd(grid + didx(c, nxg, nyg, ix, jy, ks, kp, nunc, nage), dep * (wx * wy));
    if (in11) atomicAdd(grid + didx(c, nxg, nyg, ixp, jyp, ks, kp, nunc, nage), dep * ((1.f - wx) * (1.f - wy)));
    if (in10) atomicAdd(grid + didx(c, nxg, nyg, ixp, jy, ks, kp, nunc, nage), dep * ((1.f - wx) * wy));
    if (in01) atomicAdd(grid + didx(c, nxg, nyg, ix, jyp, ks, kp, nunc, nage), dep * (wx * (1.f - wy)));
  }
}

// =========================================================== step kernel ===
struct PState { // the advance() in/out arguments of one particle
  double xt, yt;
  float zt, up, vp, wp, usigold, vsigold, wsigold;
  int ldt, icbt;
};

// src/initialize.f90:66-217
template <bool CBL>
__device__ void do_initialize(const DevStepArgs &a, const float *sh, Rng &rng, int nrand,
                              PState &s) {
  const DevCfg &c = a.cfg;
  s.icbt = 1;
  const int ngrid = pole_grid(c, s.yt); // "defined" behaviour, see DESIGN.md
  int ix = d_int(s.xt), jy = d_int(s.yt), ixp = ix + 1, jyp = jy + 1;
  if (jyp >= c.nymax) jyp = jyp - 1;
  Hz z;
  z.ngrid = ngrid;
  make_weights(c, z, c.itime, (float)s.xt, (float)s.yt, ix, jy, ixp, jyp);
  Turb t;
  {
    float h = __ldg(a.met[0].S + z.o00).x;
    h = fmaxf(h, __ldg(a.met[0].S + z.o10).x);
    h = fmaxf(h, __ldg(a.met[0].S + z.o01).x);
    h = fmaxf(h, __ldg(a.met[0].S + z.o11).x);
    h = fmaxf(h, __ldg(a.met[1].S + z.o00).x);
    h = fmaxf(h, __ldg(a.met[1].S + z.o10).x);
    h = fmaxf(h, __ldg(a.met[1].S + z.o01).x);
    h = fmaxf(h, __ldg(a.met[1].S + z.o11).x);
    t.h = h;
  }
  t.zeta = s.zt / t.h;
  float usig, vsig, wsig;
  if (t.zeta <= 1.f) {
    interp_surface(a.met, z, t);
    const int indz = find_indz(sh, c.nz, s.zt), indzp = indz + 1;
    Lev lo, hi;
    profile_level<true>(c, a.met, z, indz, lo);
    profile_level<true>(c, a.met, z, indzp, hi);
    // (u, v, w of initialize are dead: advance re-interpolates)
    if (c.turbswitch) hanna(t, s.zt, hanna_regime(t)); else hanna1(t, s.zt, hanna_regime(t));
    if (nrand + 2 > c.maxrand) nrand = 1;
    s.up = rng.get(nrand) * t.sigu;
    s.vp = rng.get(nrand + 1) * t.sigv;
    s.wp = rng.get(nrand + 2);
    if (!c.turbswitch) {
      s.wp = s.wp * t.sigw;
    } else if (CBL && c.cblflag == 1) {
      if (-t.h / t.ol > 5.f)
        s.wp = cbl_initial_velocity(c, (float)(nrand - 1) / (float)(c.maxrand - 1),
                                    rng.get(nrand + 3), s.zt, t.wst, t.h, t.sigw, t.ol);
      else
        s.wp = s.wp * t.sigw;
#ifdef FPB_DEBUG_NAN
      if (isnan(s.wp))
        printf("init nan: zt %g wst %g h %g sigw %g ol %g nrand %d r %g\n", s.zt, t.wst, t.h, t.sigw, t.ol, nrand, rng.get(nrand + 3));
#endif
    }
    if (c.turbswitch) {
      float q = fminf(t.tlw, t.h / fmaxf(2.f * fabsf(s.wp * t.sigw), 1.e-5f));
      q = fminf(q, 0.5f / fabsf(t.dsigwdz));
      q = fminf(q, 600.f);
      s.ldt = f_int(q * c.ctl);
    } else {
      float q = fminf(t.tlw, t.h / fmaxf(2.f * fabsf(s.wp), 1.e-5f));
      q = fminf(q, 600.f);
      s.ldt = f_int(q * c.ctl);
    }
    s.ldt = max(s.ldt, c.mintime);
    usig = (hi.usig + lo.usig) / 2.f;
    vsig = (hi.vsig + lo.vsig) / 2.f;
    wsig = (hi.wsig + lo.wsig) / 2.f;
  } else {
    float u, v, w;
    int indz = 0;
    interp_wind<true>(c, a.met, z, sh, s.zt, u, v, w, usig, vsig, wsig, indz);
    s.ldt = abs(c.lsynctime);
    if (nrand + 1 > c.maxrand) nrand = 1;
    s.up = rng.get(nrand) * 0.3f;
    s.vp = rng.get(nrand + 1) * 0.3f;
    nrand = nrand + 2;
    s.wp = 0.f;
  }
  if (nrand + 2 > c.maxrand) nrand = 1;
  s.usigold = rng.get(nrand) * usig;
  s.vsigold = rng.get(nrand + 1) * vsig;
  s.wsigold = rng.get(nrand + 2) * wsig;
}

__device__ __forceinline__ unsigned long long warp_sum(unsigned v) {
  return (unsigned long long)__reduce_add_sync(0xffffffffu, v);
}

__device__ __forceinline__ void make_rng(const DevCfg &c, const float *tab, int slot, Rng &rng) {
  rng.tab = tab;
  rng.maxrand = c.maxrand;
  rng.mode = c.rng_mode;
  rng.key = make_uint2((uint32_t)c.seed, (uint32_t)(c.seed >> 32));
  rng.pid = (uint32_t)(c.part_id_offset + c.part_id_stride * slot);
  rng.tstep = (uint32_t)c.itime;
  rng.cblk = -1;
}

// nrand at the top of advance(): the ran3 draw of src/advance.f90:153
__device__ __forceinline__ int advance_nrand(const DevStepArgs &a, int slot) {
  const DevCfg &c = a.cfg;
  if (c.rng_mode == FPB_RNG_REFERENCE) return a.nrand_adv[slot];
  if (c.rng_mode == FPB_RNG_PHILOX) return 64;
  Rng rng;
  make_rng(c, a.rannumb, slot, rng);
  return f_int(rng.uniform(2u) * (float)(c.maxrand - 1)) + 1;
}

// initialize() for the particles released this step (src/timemanager.f90:553-555).
// A separate launch keeps this once-per-lifetime code out of the step kernel.
template <bool CBL>
__global__ void __launch_bounds__(128)
fpb_init_kernel(const __grid_constant__ DevStepArgs a) {
  const DevCfg &c = a.cfg;
  __shared__ float sh[FPB_MAXNZ];
  for (int i = threadIdx.x; i < c.nz; i += blockDim.x) sh[i] = a.height[i];
  __syncthreads();
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int itime = c.itime;
  unsigned n_init = 0;
  if (j < c.numpart && a.p.itra1[j] == itime && ((a.p.itramem[j] == itime) || (itime == 0))) {
    n_init = 1;
    PState s;
    s.xt = a.p.xtra1[j];
    s.yt = a.p.ytra1[j];
    s.zt = a.p.ztra1[j];
    s.ldt = a.p.idt[j];
    const int slot = a.p.slot[j];
    Rng rng;
    make_rng(c, a.rannumb, slot, rng);
    int nrand;
    if (c.rng_mode == FPB_RNG_REFERENCE) nrand = a.nrand_init[slot];
    else if (c.rng_mode == FPB_RNG_PHILOX) nrand = 4;
    else nrand = f_int(rng.uniform(1u) * (float)(c.maxrand - 1)) + 1;
    do_initialize<CBL>(a, sh, rng, nrand, s);
    a.p.idt[j] = s.ldt;
    a.p.uap[j] = s.up; a.p.ucp[j] = s.vp; a.p.uzp[j] = s.wp;
    a.p.us[j] = s.usigold; a.p.vs[j] = s.vsigold; a.p.ws[j] = s.wsigold;
    a.p.cbt[j] = (int16_t)s.icbt;
  }
  if (a.stats) {
    unsigned long long v1 = warp_sum(n_init);
    if ((threadIdx.x & 31) == 0 && v1) atomicAdd(a.stats + 1, v1);
  }
}

#include "fpb_step.cuh"
#include "fpb_release.cuh"

// ======================================================= conccalc kernel ===
// A grid cell is named by a species-free key
//   key = ix + nxg*(jy + nyg*((kz-1) + nzg*((kp-1) + mps*((nc-1) + ncu*(na-1)))))
// and species ks lives at  inner + nxyz*((ks-1) + nspec*rest)  with
// inner = key % nxyz, rest = key / nxyz  (reference index order,
// src/outgrid_init.f90:192-193, packed to nspec).
__device__ __forceinline__ unsigned cell_key(const DevCfg &c, int nxg, int nyg, int ix, int jy,
                                             int kz, int kp, int nc, int na) {
  unsigned i = (unsigned)(na - 1);
  i = i * c.nclassunc + (nc - 1);
  i = i * c.maxpointspec_act + (kp - 1);
  i = i * c.numzgrid + (kz - 1);
  i = i * nyg + jy;
  i = i * nxg + ix;
  return i;
}

// red.global.add.f32 straight into the grid.  (Measured and rejected, round 2: aggregating the lanes
// that hit the same cell with match.any + shuffles before ONE atomic per group -- the rows are
// cell-sorted, so most of a warp shares its output cell -- is SLOWER on B200: conccalc 0.052 vs
// 0.034 ms at 1 M particles, 3.47 vs 2.13 ms at 100 M; the LSU already merges same-address reds of
// a warp.)
struct AtomicSink {
  float *grid[2];
  __device__ void add(const DevCfg &c, int nest, int /*slot*/, int nxyz, unsigned key,
                      const float *v) const {
    const size_t inner = key % (unsigned)nxyz, rest = key / (unsigned)nxyz;
    for (int ks = 0; ks < c.nspec; ks++)
      atomicAdd(grid[nest] + inner + (size_t)nxyz * (ks + (size_t)c.nspec * rest), v[ks]);
  }
  __device__ void skip(int, int) const {}
};

struct RecordSink { // (key, value) records for the deterministic path
  unsigned *keys; // [4*numpart], record id = 4*i + slot
  float *vals;    // [nspec][4*numpart]
  size_t nrec;
  int i;
  int nest_sel;
  __device__ void add(const DevCfg &c, int nest, int slot, int /*nxyz*/, unsigned key,
                      const float *v) const {
    if (nest != nest_sel) return;
    const size_t r = 4 * (size_t)i + slot;
    keys[r] = key;
    for (int ks = 0; ks < c.nspec; ks++) vals[(size_t)ks * nrec + r] = v[ks];
  }
};

// src/conccalc.f90:50-444 for particle i
template <class Sink>
__device__ __forceinline__ void conc_particle(const DevConcArgs &a, const float *sh, int i,
                                              const Sink &sink) {
  const DevCfg &c = a.cfg;
  const int itime = c.itime;
  const float weight = c.weight;

  const int itage = abs(itime - a.p.itramem[i]);
  int nage;
  for (nage = 1; nage <= c.nageclass; nage++)
    if (itage < c.lage[nage - 1]) break;

  const double xd = a.p.xtra1[i], yd = a.p.ytra1[i];
  const float zt = a.p.ztra1[i];

  float rhoi = 1.f;
  if (c.ind_samp == -1) { // conccalc.f90:80-121, density of memind(2)
    int ix = d_int(xd), jy = d_int(yd), ixp = ix + 1, jyp = jy + 1;
    float ddx = (float)(xd - (float)ix), ddy = (float)(yd - (float)jy);
    float rddx = 1.f - ddx, rddy = 1.f - ddy;
    float p1 = rddx * rddy, p2 = ddx * rddy, p3 = rddx * ddy, p4 = ddx * ddy;
    if (jyp >= c.nymax) jyp = jyp - 1;
    const int indz = find_indz(sh, c.nz, zt);
    const float dz1 = zt - sh[indz - 1], dz2 = sh[indz] - zt;
    const float dz = 1.f / (dz1 + dz2);
    const int plane = c.nxd * c.nyd;
    float rp[2];
#pragma unroll
    for (int n = 0; n < 2; n++) {
      const float4 *A = a.met[1].A + (indz - 1 + n) * plane;
      rp[n] = p1 * __ldg(A + ix + c.nxd * jy).w + p2 * __ldg(A + ixp + c.nxd * jy).w +
              p3 * __ldg(A + ix + c.nxd * jyp).w + p4 * __ldg(A + ixp + c.nxd * jyp).w;
    }
    rhoi = (dz1 * rp[1] + dz2 * rp[0]) * dz;
  }

  const int nrelpointer =
      ((c.ioutputforeachrelease == 0) || (c.mdomainfill == 1)) ? 1 : a.p.npoint[i];
  int kz;
  for (kz = 1; kz <= c.numzgrid; kz++)
    if (c.outheight[kz - 1] > zt) break;
  if (kz > c.numzgrid) return;

  const int nclass = a.p.nclass[i];
  const bool bk = c.drybkdep || c.wetbkdep;
  float val[FPB_MAXSPEC], vw[FPB_MAXSPEC]; // xmass1/rhoi*weight [*max(xscav,0)]
#pragma unroll
  for (int ks = 0; ks < FPB_MAXSPEC; ks++) {
    val[ks] = 0.f;
    if (ks < c.nspec) {
      float xm = a.p.xmass1[(size_t)ks * a.p.maxpart + i];
      val[ks] = xm / rhoi * weight;
      if (bk) val[ks] = val[ks] * fmaxf(a.p.xscav_frac1[(size_t)ks * a.p.maxpart + i], 0.0f);
    }
  }

  for (int nest = 0; nest <= (c.nested_output == 1 ? 1 : 0); nest++) {
    const int nxg = nest ? c.numxgridn : c.numxgrid, nyg = nest ? c.numygridn : c.numygrid;
    const int nxyz = nxg * nyg * c.numzgrid;
    float xl, yl;
    if (nest) {
      xl = (float)((xd * c.dx + c.xoutshiftn) / c.dxoutn);
      yl = (float)((yd * c.dy + c.youtshiftn) / c.dyoutn);
    } else {
      xl = (float)((xd * c.dx + c.xoutshift) / c.dxout);
      yl = (float)((yd * c.dy + c.youtshift) / c.dyout);
    }
    int ix = f_int(xl);
    if (xl < 0.f) ix = ix - 1;
    int jy = f_int(yl);
    if (yl < 0.f) jy = jy - 1;

    if ((!c.lusekerneloutput) || (itage < 10800) || (xl < 0.5f) || (yl < 0.5f) ||
        (xl > (float)(nxg - 1) - 0.5f) || (yl > (float)(nyg - 1) - 0.5f)) {
      if ((ix >= 0) && (jy >= 0) && (ix <= nxg - 1) && (jy <= nyg - 1)) {
        if (!bk && c.lparticlecountoutput) {
#pragma unroll
          for (int ks = 0; ks < FPB_MAXSPEC; ks++) vw[ks] = 1.f;
          sink.add(c, nest, 0, nxyz, cell_key(c, nxg, nyg, ix, jy, kz, nrelpointer, nclass, nage), vw);
        } else {
          sink.add(c, nest, 0, nxyz, cell_key(c, nxg, nyg, ix, jy, kz, nrelpointer, nclass, nage), val);
        }
      }
    } else {
      float ddx = xl - (float)ix, ddy = yl - (float)jy, wx, wy;
      int ixp, jyp;
      if (ddx > 0.5f) { ixp = ix + 1; wx = 1.5f - ddx; } else { ixp = ix - 1; wx = 0.5f + ddx; }
      if (ddy > 0.5f) { jyp = jy + 1; wy = 1.5f - ddy; } else { jyp = jy - 1; wy = 0.5f + ddy; }
      const bool xin = (ix >= 0) && (ix <= nxg - 1), xpin = (ixp >= 0) && (ixp <= nxg - 1);
      const bool yin = (jy >= 0) && (jy <= nyg - 1), ypin = (jyp >= 0) && (jyp <= nyg - 1);
      // the four cells in the reference's order, conccalc.f90:225-283
      if (xin && yin) {
        const float w = wx * wy;
#pragma unroll
        for (int ks = 0; ks < FPB_MAXSPEC; ks++) vw[ks] = val[ks] * w;
        sink.add(c, nest, 0, nxyz, cell_key(c, nxg, nyg, ix, jy, kz, nrelpointer, nclass, nage), vw);
      }
      if (xin && ypin) {
        const float w = wx * (1.f - wy);
#pragma unroll
        for (int ks = 0; ks < FPB_MAXSPEC; ks++) vw[ks] = val[ks] * w;
        sink.add(c, nest, 1, nxyz, cell_key(c, nxg, nyg, ix, jyp, kz, nrelpointer, nclass, nage), vw);
      }
      if (xpin && ypin) {
        const float w = (1.f - wx) * (1.f - wy);
#pragma unroll
        for (int ks = 0; ks < FPB_MAXSPEC; ks++) vw[ks] = val[ks] * w;
        sink.add(c, nest, 2, nxyz, cell_key(c, nxg, nyg, ixp, jyp, kz, nrelpointer, nclass, nage), vw);
      }
      if (xpin && yin) {
        const float w = (1.f - wx) * wy;
#pragma unroll
        for (int ks = 0; ks < FPB_MAXSPEC; ks++) vw[ks] = val[ks] * w;
        sink.add(c, nest, 3, nxyz, cell_key(c, nxg, nyg, ixp, jy, kz, nrelpointer, nclass, nage), vw);
      }
    }
  }
}

__global__ void __launch_bounds__(256)
fpb_conccalc_kernel(const __grid_constant__ DevConcArgs a) {
  const DevCfg &c = a.cfg;
  __shared__ float sh[FPB_MAXNZ];
  for (int i = threadIdx.x; i < c.nz; i += blockDim.x) sh[i] = a.height[i];
  __syncthreads();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= c.numpart) return;
  if (a.p.itra1[i] != c.itime) return;
  AtomicSink sink;
  sink.grid[0] = a.gridunc;
  sink.grid[1] = a.griduncn;
  conc_particle(a, sh, i, sink);
}

// deterministic path, stage 1: one (key, value) record per touched cell.
// keys must be pre-filled with 0xffffffff (= no record).
__global__ void __launch_bounds__(256)
fpb_conc_emit_kernel(const __grid_constant__ DevConcArgs a, int nest_sel, unsigned *keys,
                     float *vals, size_t nrec) {
  const DevCfg &c = a.cfg;
  __shared__ float sh[FPB_MAXNZ];
  for (int i = threadIdx.x; i < c.nz; i += blockDim.x) sh[i] = a.height[i];
  __syncthreads();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= c.numpart) return;
  if (a.p.itra1[i] != c.itime) return;
  RecordSink sink;
  sink.keys = keys;
  sink.vals = vals;
  sink.nrec = nrec;
  sink.i = a.p.slot[i] - a.slot_base; // record order = slot order = the reference's particle order
  sink.nest_sel = nest_sel;
  conc_particle(a, sh, i, sink);
}

// src/conccalc.f90:451-498: each thread evaluates its particle against every
// receptor; block partial sums go to crec_acc[n][ks] (= c(ks) of the reference).
__global__ void __launch_bounds__(256)
fpb_receptor_kernel(const __grid_constant__ DevConcArgs a) {
  const DevCfg &c = a.cfg;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const float factor = .596831f, hxmax = 6.0f, hymax = 4.0f, hzmax = 150.f;
  bool act = (i < c.numpart) && (a.p.itra1[i] == c.itime);
  double xd0 = 0., yd0 = 0.;
  float zt = 0.f, hz = 1.f, hx = 1.f, hy = 1.f, zd = 2.f;
  if (act) {
    const int itage = abs(c.itime - a.p.itramem[i]);
    xd0 = a.p.xtra1[i]; yd0 = a.p.ytra1[i]; zt = a.p.ztra1[i];
    hz = fminf(50.f + 0.3f * m_sqrt((float)itage), hzmax);
    zd = zt / hz;
    hx = fminf((0.29f + 2.222e-3f * m_sqrt((float)itage)) * c.dx + (float)itage * 1.2e-5f, hxmax);
    hy = fminf((0.18f + 1.389e-3f * m_sqrt((float)itage)) * c.dy + (float)itage * 7.5e-6f, hymax);
    if (zd > 1.f) act = false;
  }
  for (int n = 0; n < c.numreceptor; n++) {
    float kern = 0.f;
    if (act) {
      float xd = (float)((xd0 - c.xreceptor[n]) / hx);
      float yd = (float)((yd0 - c.yreceptor[n]) / hy);
      if (!(xd * xd > 1.f) && !(yd * yd > 1.f)) {
        float h = hx * hy * hz;
        float r2 = xd * xd + yd * yd + zd * zd;
        if (r2 < 1.f) kern = factor * (1.f - r2) / h;
      }
    }
    if (a.rec_vals) { // deterministic: the contributions are summed in particle order afterwards
      if (i < c.numpart && a.p.itra1[i] == c.itime) {
        const int sl = a.p.slot[i] - a.slot_base;
        for (int ks = 0; ks < c.nspec; ks++)
          a.rec_vals[(size_t)(n * c.nspec + ks) * a.rec_nslots + sl] =
              (kern != 0.f) ? a.p.xmass1[(size_t)ks * a.p.maxpart + i] * kern : 0.f;
      }
      continue;
    }
    for (int ks = 0; ks < c.nspec; ks++) {
      float v = (kern != 0.f) ? a.p.xmass1[(size_t)ks * a.p.maxpart + i] * kern : 0.f;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if ((threadIdx.x & 31) == 0 && v != 0.f) atomicAdd(a.crec_acc + n * c.nspec + ks, v);
    }
  }
}


// ======================================================== wet deposition ===
// src/wetdepo.f90:70-147, src/get_wetscav.f90:78-314, src/interpol_rain.f90:77-127,
// src/wetdepokernel.f90:38-108, src/wetdepokernel_nest.f90:38-105
__device__ __forceinline__ float wet_powi(float x, int m) { // real**integer as libgcc's __powisf2
  unsigned n = (unsigned)(m < 0 ? -m : m);
  float y = (n % 2) ? x : 1.f;
  while (n >>= 1) {
    x = x * x;
    if (n % 2) y = y * x;
  }
  return m < 0 ? 1.f / y : y;
}

// exponent of the Laakso / Kyro below-cloud polynomials (src/get_wetscav.f90:232-243):
// terms of order 1e2..1e3 cancel to order 1, so the sum is evaluated with explicitly
// rounded operations (no FMA contraction) in every math mode
__device__ __forceinline__ float wet_poly(const float *b, float lg, float sqrt_prec) {
  float e = __fadd_rn(b[0], __fmul_rn(b[1], wet_powi(lg, -4)));
  e = __fadd_rn(e, __fmul_rn(b[2], wet_powi(lg, -3)));
  e = __fadd_rn(e, __fmul_rn(b[3], wet_powi(lg, -2)));
  e = __fadd_rn(e, __fmul_rn(b[4], wet_powi(lg, -1)));
  e = __fadd_rn(e, __fmul_rn(b[5], sqrt_prec));
  return e;
}

// returns wetscav; grfraction1 is written only when scavenging is evaluated
__device__ float get_wetscav(const DevWetArgs &a, const float *sh, int i, int ks, float &grfraction1) {
  const DevCfg &c = a.cfg;
  const float lfr[5] = {0.5f, 0.65f, 0.8f, 0.9f, 0.95f};
  const float cfr[5] = {0.4f, 0.55f, 0.7f, 0.8f, 0.9f};
  const float bclr[6] = {274.35758f, 332839.59273f, 226656.57259f, 58005.91340f, 6588.38582f, 0.244984f};
  const float bcls[6] = {22.7f, 0.0f, 0.0f, 1321.0f, 381.0f, 0.0f};
  float wetscav = 0.f;
  const double xd = a.p.xtra1[i], yd = a.p.ytra1[i];
  int ngrid = 0;
  for (int j = c.numbnests; j >= 1; j--) // no eps margin here, src/get_wetscav.f90:82-90
    if ((xd > c.xln[j - 1]) && (xd < c.xrn[j - 1]) && (yd > c.yln[j - 1]) && (yd < c.yrn[j - 1])) {
      ngrid = j;
      break;
    }
  const DevMetSlot &M = (ngrid > 0) ? a.metn[ngrid - 1] : a.met;
  const int nxd = (ngrid > 0) ? c.nxdn[ngrid - 1] : c.nxd, nyd = (ngrid > 0) ? c.nydn[ngrid - 1] : c.nyd;
  const int nxu = (ngrid > 0) ? c.nxdn[ngrid - 1] : c.nx, nyu = (ngrid > 0) ? c.nydn[ngrid - 1] : c.ny;
  float xt, yt;
  int ix, jy;
  bool clouds_read;
  if (ngrid > 0) {
    xt = (float)((xd - c.xln[ngrid - 1]) * c.xresoln[ngrid - 1]);
    yt = (float)((yd - c.yln[ngrid - 1]) * c.yresoln[ngrid - 1]);
    ix = f_int(xt);
    jy = f_int(yt);
    clouds_read = c.readclouds_nest[ngrid - 1] != 0;
  } else {
    xt = (float)xd;
    yt = (float)yd;
    ix = d_int(xd);
    jy = d_int(yd);
    clouds_read = c.readclouds != 0;
  }
  float lsp, convp, cc;
  { // interpol_rain, level 1 of the chosen time level
    if (xt >= (float)(nxu - 1)) xt = (float)(nxu - 1) - 0.00001f;
    if (yt >= (float)(nyu - 1)) yt = (float)(nyu - 1) - 0.00001f;
    const int jx = f_int(xt), jj = f_int(yt), jxp = jx + 1, jjp = jj + 1;
    const float ddx = xt - (float)jx, ddy = yt - (float)jj, rddx = 1.f - ddx, rddy = 1.f - ddy;
    const float p1 = rddx * rddy, p2 = ddx * rddy, p3 = rddx * ddy, p4 = ddx * ddy;
    const float4 r1 = __ldg(M.R + jx + nxd * jj), r2 = __ldg(M.R + jxp + nxd * jj),
                 r3 = __ldg(M.R + jx + nxd * jjp), r4 = __ldg(M.R + jxp + nxd * jjp);
    lsp = p1 * r1.x + p2 * r2.x + p3 * r3.x + p4 * r4.x;
    convp = p1 * r1.y + p2 * r2.y + p3 * r3.y + p4 * r4.y;
    cc = p1 * r1.z + p2 * r2.z + p3 * r3.z + p4 * r4.z;
  }
  if ((lsp < 0.01f) && (convp < 0.01f)) return wetscav;

  const int hz = find_indz(sh, c.nz, a.p.ztra1[i]);
  const size_t i3 = (size_t)ix + (size_t)nxd * ((size_t)jy + (size_t)nyd * (size_t)(hz - 1));
  const int clouds_v = __ldg(M.C + i3);
  if (clouds_v <= 1) return wetscav;

  int li, lj;
  if (lsp > 20.f) li = 5; else if (lsp > 8.f) li = 4; else if (lsp > 3.f) li = 3; else if (lsp > 1.f) li = 2; else li = 1;
  if (convp > 20.f) lj = 5; else if (convp > 8.f) lj = 4; else if (convp > 3.f) lj = 3; else if (convp > 1.f) lj = 2; else lj = 1;
  grfraction1 = fmaxf(0.05f, cc * (lsp * lfr[li - 1] + convp * cfr[lj - 1]) / (lsp + convp));
  const float prec1 = (lsp + convp) / grfraction1;
  const float act_temp = __ldg(M.T + i3);

  if (clouds_v >= 4) { // below-cloud scavenging
    if ((c.dquer[ks] <= 0.f) && (c.weta_gas[ks] > 0.f || c.wetb_gas[ks] > 0.f)) {
      wetscav = c.weta_gas[ks] * m_pow(prec1, c.wetb_gas[ks]);
    } else if ((c.dquer[ks] > 0.f) && (c.crain_aero[ks] > 0.f || c.csnow_aero[ks] > 0.f)) {
      const float dquer_m = fminf(10.f, c.dquer[ks]) / 1000000.f;
      const float lg = m_log10(dquer_m);
      if (act_temp >= 273.f && c.crain_aero[ks] > 0.f) {
        wetscav = c.crain_aero[ks] * m_pow(10.f, wet_poly(bclr, lg, m_pow(prec1, 0.5f)));
      } else if (act_temp < 273.f && c.csnow_aero[ks] > 0.f) {
        wetscav = c.csnow_aero[ks] * m_pow(10.f, wet_poly(bcls, lg, m_pow(prec1, 0.5f)));
      }
    }
  }
  if (clouds_v < 4) { // in-cloud scavenging
    float ccn = c.ccn_aero[ks], in = c.in_aero[ks];
    if ((ccn > 0.f || in > 0.f) || (c.henry[ks] > 0.f && c.dquer[ks] <= 0.f)) {
      if (ccn < 0.f) ccn = 0.f;
      if (in < 0.f) in = 0.f;
      float cl;
      if (clouds_read) cl = __ldg(M.R + ix + nxd * jy).w * (grfraction1 / cc);
      else cl = (1.e6f * 2.e-7f) * m_pow(prec1, 0.36f);
      float liq_frac, ice_frac;
      if (act_temp <= 253.f) {
        liq_frac = 0.f; ice_frac = 1.f;
      } else if (act_temp >= 273.f) {
        liq_frac = 1.f; ice_frac = 0.f;
      } else {
        const float q = (act_temp - 273.f) / (273.f - 253.f);
#if FPB_STRICT
        ice_frac = (float)pow((double)q, 2.0);
#else
        ice_frac = q * q;
#endif
        liq_frac = fmaxf(0.f, 1.f - ice_frac);
      }
      const float frac_act = liq_frac * ccn + ice_frac * in;
      float S_i;
      if (c.dquer[ks] > 0.f) {
        S_i = frac_act / cl;
      } else {
        const float cle = (1.f - cl) / (c.henry[ks] * (287.05f / 3500.f) * act_temp) + cl;
        S_i = 1.f / cle;
      }
      wetscav = 6.2f * S_i * (prec1 / 3.6e6f); // incloud_ratio, src/par_mod.f90:82
    }
  }
  return wetscav;
}

__device__ void wetdepo_scatter(const DevCfg &c, float *grid, bool nest, int nunc, const float *deposit,
                                float x, float y, int nage, int kp, const DevDepRecords &rec, int rslot) {
  const int nxg = nest ? c.numxgridn : c.numxgrid, nyg = nest ? c.numygridn : c.numygrid;
  float xl, yl;
  int ix, jy, ixp, jyp;
  if (nest) { // floor(), src/wetdepokernel_nest.f90:43-44
    xl = (x * c.dx + c.xoutshiftn) / c.dxoutn;
    yl = (y * c.dy + c.youtshiftn) / c.dyoutn;
    ix = (int)floorf(xl);
    jy = (int)floorf(yl);
  } else {
    xl = (x * c.dx + c.xoutshift) / c.dxout;
    yl = (y * c.dy + c.youtshift) / c.dyout;
    ix = f_int(xl);
    jy = f_int(yl);
  }
  const float ddx = xl - (float)ix, ddy = yl - (float)jy;
  float wx, wy;
  if (ddx > 0.5f) { ixp = ix + 1; wx = 1.5f - ddx; } else { ixp = ix - 1; wx = 0.5f + ddx; }
  if (ddy > 0.5f) { jyp = jy + 1; wy = 1.5f - ddy; } else { jyp = jy - 1; wy = 0.5f + ddy; }
  const bool in00 = (ix >= 0) && (jy >= 0) && (ix <= nxg - 1) && (jy <= nyg - 1);
  const bool in11 = (ixp >= 0) && (jyp >= 0) && (ixp <= nxg - 1) && (jyp <= nyg - 1);
  const bool in10 = (ixp >= 0) && (jy >= 0) && (ixp <= nxg - 1) && (jy <= nyg - 1);
  const bool in01 = (ix >= 0) && (jyp >= 0) && (ix <= nxg - 1) && (jyp <= nyg - 1);
  if (rec.keys[0]) { // deterministic: records
    const int g = nest ? 1 : 0;
    if (!nest && !c.lusekerneloutput) {
      if (in00) dep_record(c, rec, g, rslot, 0, dep_key(c, nxg, nyg, ix, jy, kp, nunc, nage), deposit, 1.f, nullptr);
      return;
    }
    if (in00) dep_record(c, rec, g, rslot, 0, dep_key(c, nxg, nyg, ix, jy, kp, nunc, nage), deposit, wx * wy, nullptr);
    if (in11) dep_record(c, rec, g, rslot, 1, dep_key(c, nxg, nyg, ixp, jyp, kp, nunc, nage), deposit, (1.f - wx) * (1.f - wy), nullptr);
    if (in10) dep_record(c, rec, g, rslot, 2, dep_key(c, nxg, nyg, ixp, jy, kp, nunc, nage), deposit, (1.f - wx) * wy, nullptr);
    if (in01) dep_record(c, rec, g, rslot, 3, dep_key(c, nxg, nyg, ix, jyp, kp, nunc, nage), deposit, wx * (1.f - wy), nullptr);
    return;
  }
  for (int ks = 1; ks <= c.nspec; ks++) {
    const float dep = deposit[ks - 1];
    if (dep == 0.f) continue; // adding 0 changes nothing
    if (!nest && !c.lusekerneloutput) {
      if (in00) atomicAdd(grid + didx(c, nxg, nyg, ix, jy, ks, kp, nunc, nage), dep);
      continue;
    }
    if (in00) atomicAdd(grid + didx(c, nxg, nyg, ix, jy, ks, kp, nunc, nage), dep * (wx * wy));
    if (in11) atomicAdd(grid + didx(c, nxg, nyg, ixp, jyp, ks, kp, nunc, nage), dep * ((1.f - wx) * (1.f - wy)));
    if (in10) atomicAdd(grid + didx(c, nxg, nyg, ixp, jy, ks, kp, nunc, nage), dep * ((1.f - wx) * wy));
    if (in01) atomicAdd(grid + didx(c, nxg, nyg, ix, jyp, ks, kp, nunc, nage), dep * (wx * (1.f - wy)));
  }
}

__global__ void __launch_bounds__(256)
fpb_wetdepo_kernel(const __grid_constant__ DevWetArgs a) {
  const DevCfg &c = a.cfg;
  __shared__ float sh[FPB_MAXNZ];
  for (int i = threadIdx.x; i < c.nz; i += blockDim.x) sh[i] = a.height[i];
  __syncthreads();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= c.numpart) return;
  const int itra1 = a.p.itra1[i];
  if (itra1 == FPB_ITRA_DEAD) return;
  if (c.ldirect == 1) {
    if (itra1 > c.itime) return;
  } else {
    if (itra1 < c.itime) return;
  }
  const int itage = abs(itra1 - a.p.itramem[i]);
  int nage;
  for (nage = 1; nage <= c.nageclass; nage++)
    if (itage < c.lage[nage - 1]) break;

  float wetdeposit[FPB_MAXSPEC];
  float grfraction1 = 0.f;
  bool any = false;
#pragma unroll
  for (int ks = 0; ks < FPB_MAXSPEC; ks++) {
    wetdeposit[ks] = 0.f;
    if (ks < c.nspec && c.wetdepspec[ks]) {
      const float wetscav = get_wetscav(a, sh, i, ks, grfraction1);
      float *xm = a.p.xmass1 + (size_t)ks * a.p.maxpart + i;
      const float xm0 = *xm;
      if (wetscav > 0.f)
        wetdeposit[ks] = xm0 * (1.f - m_exp(-wetscav * (float)abs(a.ltsample))) * grfraction1;
      const float restmass = xm0 - wetdeposit[ks];
      *xm = (restmass > EPS3) ? restmass : 0.f;
      if (c.decay[ks] > 0.f) wetdeposit[ks] = wetdeposit[ks] * m_exp((float)abs(c.ldeltat) * c.decay[ks]);
      any = any || (wetdeposit[ks] != 0.f);
    }
  }
  if (c.ldirect == 1 && any) {
    const int kp = (c.ioutputforeachrelease == 1) ? a.p.npoint[i] : 1;
    const int nclass = a.p.nclass[i];
    const float x = (float)a.p.xtra1[i], y = (float)a.p.ytra1[i];
    const int rslot = a.p.slot[i] - a.dep.slot_base;
    wetdepo_scatter(c, a.wetgridunc, false, nclass, wetdeposit, x, y, nage, kp, a.dep, rslot);
    if (c.nested_output == 1) wetdepo_scatter(c, a.wetgriduncn, true, nclass, wetdeposit, x, y, nage, kp, a.dep, rslot);
  }
}

// ----------------------------------------------------------------------------
// RECEPTOR: dry/wet depovel (src/timemanager.f90:563-598): in a backward deposition run
// (IND_RECEPTOR 3 / 4) the scavenged fraction xscav_frac1 is determined once, right after the
// release (it was initialised negative) and before the particle moves: the deposition velocity at
// the particle (get_vdep_prob, src/get_vdep_prob.f90:41-124) or wetscav * release depth *
// grfraction (get_wetscav); a species that is not deposited loses its mass.
__global__ void __launch_bounds__(256)
fpb_bkdep_kernel(const __grid_constant__ DevBkdepArgs a) {
  const DevCfg &c = a.w.cfg;
  __shared__ float sh[FPB_MAXNZ];
  for (int i = threadIdx.x; i < c.nz; i += blockDim.x) sh[i] = a.w.height[i];
  __syncthreads();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= c.numpart) return;
  const DevParticles &p = a.w.p;
  if (p.itra1[i] != c.itime) return;
  if (c.drybkdep) {
    bool todo = false;
    for (int ks = 0; ks < c.nspec; ks++) todo = todo || (p.xscav_frac1[(size_t)ks * p.maxpart + i] < 0.f);
    if (todo) {
      const double xt = p.xtra1[i], yt = p.ytra1[i];
      const float zt = p.ztra1[i];
      const int ngrid = choose_grid(c, xt, yt);
      Hz z;
      z.ngrid = ngrid;
      const DevMetSlot *met = a.vmet;
      if (ngrid > 0) { // the weights of the grid the velocities are read from ("defined")
        const int l = ngrid - 1;
        const float xf = (float)((xt - c.xln[l]) * c.xresoln[l]), yf = (float)((yt - c.yln[l]) * c.yresoln[l]);
        const int ix = f_int(xf), jy = f_int(yf);
        make_weights(c, z, c.itime, xf, yf, ix, jy, ix + 1, min(jy + 1, c.nydn[l] - 1), c.nxdn[l], c.nydn[l]);
        met = a.vmetn[l];
      } else {
        const int ix = d_int(xt), jy = d_int(yt);
        make_weights(c, z, c.itime, (float)xt, (float)yt, ix, jy, ix + 1, min(jy + 1, c.nyd - 1));
      }
      for (int ks = 0; ks < c.nspec; ks++) {
        float *xs = p.xscav_frac1 + (size_t)ks * p.maxpart + i;
        if (!(*xs < 0.f)) continue;
        if (c.drydepspec[ks]) {
          float prob = 0.f;
          if (c.drydep && (zt < 2.f * HREF)) { // interpol_vdep, src/interpol_vdep.f90:39-54
            const int off = ks * z.plane;
            const float y0 = bil(z, __ldg(met[0].vdep + off + z.o00), __ldg(met[0].vdep + off + z.o10),
                                 __ldg(met[0].vdep + off + z.o01), __ldg(met[0].vdep + off + z.o11));
            const float y1 = bil(z, __ldg(met[1].vdep + off + z.o00), __ldg(met[1].vdep + off + z.o10),
                                 __ldg(met[1].vdep + off + z.o01), __ldg(met[1].vdep + off + z.o11));
            prob = (y0 * z.dt2 + y1 * z.dt1) * z.dtt;
          }
          *xs = prob;
        } else {
          p.xmass1[(size_t)ks * p.maxpart + i] = 0.f;
          *xs = 0.f;
        }
      }
    }
  }
  if (c.wetbkdep) {
    for (int ks = 0; ks < c.nspec; ks++) {
      float *xs = p.xscav_frac1 + (size_t)ks * p.maxpart + i;
      if (!(*xs < 0.f)) continue;
      float grfraction1 = 0.f;
      const float wetscav = get_wetscav(a.w, sh, i, ks, grfraction1);
      if (wetscav > 0.f) {
        const int np = p.npoint[i] - 1;
        *xs = wetscav * (a.zpoint2[np] - a.zpoint1[np]) * grfraction1;
      } else {
        p.xmass1[(size_t)ks * p.maxpart + i] = 0.f;
        *xs = 0.f;
      }
    }
  }
}

} // namespace

#if FPB_STRICT
#define FPB_SUF(name) name##_strict
#else
#define FPB_SUF(name) name##_fast
#endif

void FPB_SUF(fpbk_init)(const DevStepArgs &a, cudaStream_t st) {
  const int nb = (a.cfg.numpart + 127) / 128;
  if (nb == 0) return;
  if (a.cfg.cblflag == 1) fpb_init_kernel<true><<<nb, 128, 0, st>>>(a);
  else fpb_init_kernel<false><<<nb, 128, 0, st>>>(a);
}

void FPB_SUF(fpbk_step)(const DevStepArgs &a, cudaStream_t st) {
  if (a.cfg.numpart <= 0) return;
  // lean variant: table RNG, no dry deposition, no settling, no CBL, no nested input grids
  const bool full = a.cfg.drydep || a.cfg.cblflag == 1 || a.cfg.lsettling ||
                    a.cfg.rng_mode == FPB_RNG_PHILOX || a.cfg.numbnests > 0;
  // persistent grid: as many CTAs as can be resident (one wave), never more than the rows need
  // lean + the usual switches (turbswitch, method 1, IFINE 4, turbulence on) as compile-time constants
  const bool spec = !full && a.cfg.turbswitch && a.cfg.method == 1 && !a.cfg.turboff && a.cfg.ifine == 4;
  static int resident[16][4] = {}; // resident CTAs per kernel variant, per device of this process
#ifdef FPB_NO_EXTRA_NOCBL
  const int variant = full ? 1 : (spec ? 2 : 0);
#else
  // 3: the full feature set without the CBL scheme (its drift routine costs registers in every path)
  const int variant = full ? (a.cfg.cblflag == 1 ? 1 : 3) : (spec ? 2 : 0);
#endif
  int dev_now = 0;
  cudaGetDevice(&dev_now);
  int uncached = 0;
  int &res = (dev_now >= 0 && dev_now < 16) ? resident[dev_now][variant] : uncached;
  if (res == 0) {
    int dev = 0, sms = 0, per_sm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (variant == 1) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fpb_pbl_kernel<true, true, false>, PBL_THREADS, 0);
    else if (variant == 3) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fpb_pbl_kernel<true, false, false>, PBL_THREADS, 0);
    else if (variant == 2) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fpb_pbl_kernel<false, false, true>, PBL_THREADS, 0);
    else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fpb_pbl_kernel<false, false, false>, PBL_THREADS, 0);
    res = sms * (per_sm > 0 ? per_sm : 1);
  }
  const int want = (a.cfg.numpart + 127) / 128;
  const int want_pbl = (a.cfg.numpart + PBL_THREADS - 1) / PBL_THREADS;
  int cap = res; // persistent grid: one wave at most
  if (a.grid_frac > 0.f && a.grid_frac < 1.f) { // (fpb_step_host: the chunks' grids share the SMs)
    cap = (int)((float)res * a.grid_frac);
    if (cap < 1) cap = 1;
  }
  const int nb = want_pbl < cap ? want_pbl : cap;
  cudaMemsetAsync(a.work_counter, 0, sizeof(int), st);
  if (variant == 1) fpb_pbl_kernel<true, true, false><<<nb, PBL_THREADS, 0, st>>>(a);
  else if (variant == 3) fpb_pbl_kernel<true, false, false><<<nb, PBL_THREADS, 0, st>>>(a);
  else if (variant == 2) fpb_pbl_kernel<false, false, true><<<nb, PBL_THREADS, 0, st>>>(a);
  else fpb_pbl_kernel<false, false, false><<<nb, PBL_THREADS, 0, st>>>(a);
  // finish kernel: the variant without nests / settling / dry deposition / Philox-direct RNG when it applies
  if (!a.cfg.drydep && !a.cfg.lsettling && a.cfg.numbnests == 0 && a.cfg.rng_mode != FPB_RNG_PHILOX && !a.cfg.linit_cond)
    fpb_finish_kernel<true><<<want, 128, 0, st>>>(a);
  else
    fpb_finish_kernel<false><<<want, 128, 0, st>>>(a);
}

void FPB_SUF(fpbk_conccalc)(const DevConcArgs &a, cudaStream_t st) {
  const int nb = (a.cfg.numpart + 255) / 256;
  if (nb == 0) return;
  fpb_conccalc_kernel<<<nb, 256, 0, st>>>(a);
}

void FPB_SUF(fpbk_conc_emit)(const DevConcArgs &a, int nest_sel, unsigned *keys, float *vals,
                              size_t nrec, cudaStream_t st) {
  const int nb = (a.cfg.numpart + 255) / 256;
  if (nb == 0) return;
  fpb_conc_emit_kernel<<<nb, 256, 0, st>>>(a, nest_sel, keys, vals, nrec);
}

void FPB_SUF(fpbk_release)(const DevReleaseArgs &a, cudaStream_t st) {
  const int nb = (a.p.maxpart + REL_BLOCK - 1) / REL_BLOCK;
  release_count_kernel<<<nb, REL_BLOCK, 0, st>>>(a);
  release_scan_kernel<<<1, REL_BLOCK, 0, st>>>(a.block_counts, a.out + 1, nb);
  release_assign_kernel<<<nb, REL_BLOCK, 0, st>>>(a);
}

void FPB_SUF(fpbk_split)(const DevSplitArgs &a, cudaStream_t st) {
  const int nb = (a.numpart_old + REL_BLOCK - 1) / REL_BLOCK;
  if (nb == 0) return;
  split_count_kernel<<<nb, REL_BLOCK, 0, st>>>(a);
  release_scan_kernel<<<1, REL_BLOCK, 0, st>>>(a.block_counts, a.total, nb);
  split_assign_kernel<<<nb, REL_BLOCK, 0, st>>>(a);
}

void FPB_SUF(fpbk_bkdep)(const DevBkdepArgs &a, cudaStream_t st) {
  const int nb = (a.w.cfg.numpart + 255) / 256;
  if (nb == 0) return;
  fpb_bkdep_kernel<<<nb, 256, 0, st>>>(a);
}

void FPB_SUF(fpbk_wetdepo)(const DevWetArgs &a, cudaStream_t st) {
  const int nb = (a.cfg.numpart + 255) / 256;
  if (nb == 0) return;
  fpb_wetdepo_kernel<<<nb, 256, 0, st>>>(a);
}

void FPB_SUF(fpbk_receptor)(const DevConcArgs &a, cudaStream_t st) {
  const int nb = (a.cfg.numpart + 255) / 256;
  if (nb == 0 || a.cfg.numreceptor == 0) return;
  fpb_receptor_kernel<<<nb, 256, 0, st>>>(a);
}
