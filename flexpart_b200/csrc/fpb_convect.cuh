// fpb_convect.cuh -- the column part of FLEXPART's convective mixing, one grid column per thread:
//   conv_calcmatrix   src/calcmatrix.f90:45-135   pressures, saturation humidity, the call of the
//                                                 Emanuel scheme, the redistribution matrix fmassfrac
//   conv_convect      src/convect43c.f90:11-972   Emanuel (2002) convection scheme 4.3c as FLEXPART
//                                                 carries it: saturated up/downdraft mass fluxes
//                                                 between all level pairs (FMASS) and the
//                                                 compensating subsidence (SUB)
//   conv_tlift        src/convect43c.f90:974-1090 lifted-parcel temperature / condensate
//   conv_qvsat, conv_ew   src/qvsat.f90, src/ew.f90
//   conv_uvzlev       src/redist.f90:63-118       heights of the eta half levels of the column
//   conv_redist       src/redist.f90:120-237      new height of one particle: destination level drawn
//                                                 from the matrix row (column for backward runs),
//                                                 or displacement by the compensating subsidence
// The reference keeps the column in module conv_mod and convect's work arrays on the stack; here
// both live in a per-column slice of a global-memory pool (ConvWork).  All arithmetic is the
// reference's, statement by statement, evaluated like the validation build of the other kernels
// (no FMA contraction, exp/log/pow evaluated in double and rounded once), so the result is
// bit-comparable with the reference's own routine (oracle/_ref).  Written for nvcc and, for the
// CPU-side cross-check of tests/test_convection.py only, for g++ (FPB_CONV_HOST).
#pragma once
#include <math.h>
#include <stdint.h>

#ifdef __CUDACC__
#define FPB_HD __host__ __device__
#else
#define FPB_HD
#endif

#ifdef __CUDA_ARCH__
#define FPB_PRAGMA(x) _Pragma(#x)
#define FPB_UNROLL(n) FPB_PRAGMA(unroll n)
#else
#define FPB_UNROLL(n)
#endif

namespace fpbconv {

FPB_HD inline float c_exp(float x) { return (float)exp((double)x); }
FPB_HD inline float c_log(float x) { return (float)log((double)x); }
FPB_HD inline float c_pow(float a, float b) { return (float)pow((double)a, (double)b); }
FPB_HD inline float c_sqrt(float x) { return sqrtf(x); }
FPB_HD inline float c_max(float a, float b) { return a > b ? a : b; }
FPB_HD inline float c_min(float a, float b) { return a < b ? a : b; }
FPB_HD inline float c_powi(float x, int m) { // real**integer as libgcc's __powisf2 (what gfortran emits)
  unsigned n = (unsigned)(m < 0 ? -m : m);
  float y = (n % 2) ? x : 1.f;
  while (n >>= 1) {
    x = x * x;
    if (n % 2) y = y * x;
  }
  return m < 0 ? 1.f / y : y;
}

// One column: conv_mod's arrays + convect's work arrays.  Vectors are 1-based (element i at
// [i*stride]), matrices Fortran-ordered: (i,j) at [(i + ld*j)*stride].  stride = 32 on the device: the
// 32 columns of a warp interleave their work slices, so that lanes touching the same element of
// their own column -- the normal case, the lanes run the same loops -- share cache lines.
struct ConvWork {
  int nuvz, nconvlev, ld, stride;
  int fstride; // stride of FMASS alone (device: the matrix lives per column, contiguous, outside the interleaved slice)
  const float *akz, *bkz, *akm, *bkm; // 1-based hybrid coefficients (src/com_mod.f90, akz(nuvz) ...)
  // conv_mod
  float *pconv, *phconv, *dpr, *pconv_hpa, *phconv_hpa, *tconv, *qconv, *qsconv, *ft, *fq, *sub;
  float *fmass; // after conv_calcmatrix: fmassfrac (fmassfrac(k,kk) = delt*fmass(k,kk) + diagonal)
  float psconv, tt2conv, td2conv;
  int nconvtop;
  // convect locals
  float *fup, *fdown, *m, *mp, *tvp, *tv, *water, *qp, *ep, *th, *wt, *evap, *clw, *sigp, *tp, *cpn, *lv, *lvcp,
      *h, *hp, *gz, *hm, *uvzlev;
  int *nent;
  int *rowtop; // per row i of MENT: max(i, largest j with ment(i,j) > EPSILON), 0: none (conv_mixnorm_row)
  float *ment; // (ELIJ is only read by the precipitating downdraft, which is left out, SIJ only inside conv_mixnorm_row:
               //  not kept)
  float *mentc; // device: the column's final MENT once more, contiguous (element (i,j) at [i + ld*j]), for the
                // block-per-column flux assembly (conv_assembly_kernel; written by conv_mix_kernel)
};

// what the two halves of conv_convect / conv_calcmatrix hand over around the flux assembly
struct ConvState {
  int go;     // 1: the scheme got as far as the flux assembly (iflag so far in `iflag`), 0: `iflag` is final
  int iflag, inb, icb, nk;
  float delti, cbmf, cbmfold;
};

constexpr int CONV_NVEC = 35; // float vectors above (+ nent, stored as one more vector)

// floats of one column's slice
FPB_HD inline size_t conv_pool_floats(int nuvz, int nconvlev, bool with_fmass = true) {
  const size_t lv = (size_t)nuvz + 4, ld = (size_t)nconvlev + 3;
  return (CONV_NVEC + 1) * lv + (with_fmass ? 2 : 1) * ld * ld;
}

// carve the column's slice: `pool` = first float of the slice (for stride 32: of the warp's block of
// 32 slices, plus the lane)
// with_fmass = false: w.fmass / w.fstride are the caller's business
FPB_HD inline void conv_carve(ConvWork &w, float *pool, int nuvz, int nconvlev, int stride = 1, bool with_fmass = true) {
  const size_t lv = ((size_t)nuvz + 4) * stride;
  w.nuvz = nuvz; w.nconvlev = nconvlev; w.ld = nconvlev + 3; w.stride = stride;
  w.mentc = nullptr;
  float **vec[CONV_NVEC] = {&w.pconv, &w.phconv, &w.dpr, &w.pconv_hpa, &w.phconv_hpa, &w.tconv, &w.qconv, &w.qsconv,
                            &w.ft, &w.fq, &w.sub, &w.fup, &w.fdown, &w.m, &w.mp, &w.tvp, &w.tv, &w.water, &w.qp,
                            &w.ep, &w.th, &w.wt, &w.evap, &w.clw, &w.sigp, &w.tp, &w.cpn, &w.lv, &w.lvcp, &w.h,
                            &w.hp, &w.gz, &w.hm, &w.uvzlev, nullptr};
  float *p = pool;
  for (int k = 0; k < CONV_NVEC - 1; k++) { *vec[k] = p; p += lv; }
  w.nent = reinterpret_cast<int *>(p); p += lv;
  w.rowtop = reinterpret_cast<int *>(p); p += lv;
  const size_t ld2 = (size_t)w.ld * w.ld * stride;
  w.ment = p; p += ld2;
  w.fmass = with_fmass ? p : nullptr;
  w.fstride = stride;
}

#define CV(a, i) w.a[(size_t)(i) * w.stride]
#define CM(a, i, j) w.a[(size_t)((i) + w.ld * (j)) * w.stride]
#define CF(i, j) w.fmass[(size_t)((i) + w.ld * (j)) * w.fstride]

// src/ew.f90: saturation vapour pressure over water [Pa] (Goff-Gratch)
FPB_HD inline float conv_ew(float x) {
  float y = 373.16f / x;
  float a = -7.90298f * (y - 1.f);
  a = a + (5.02808f * 0.43429f * c_log(y));
  float c = (1.f - (1.f / y)) * 11.344f;
  c = -1.f + c_pow(10.f, c);
  c = -1.3816f * c / c_powi(10.f, 7);
  float d = (1.f - y) * 3.49149f;
  d = -1.f + c_pow(10.f, d);
  d = 8.1328f * d / c_powi(10.f, 3);
  y = a + c + d;
  return 101324.6f * c_pow(10.f, y);
}

// src/qvsat.f90: f_qvsat with f_esl / f_esi
FPB_HD inline float conv_qvsat(float p, float t) {
  const float rddrv = 287.0f / 461.0f;
  float fespt;
  if (t >= 253.15f) {
    const float f = 1.0007f + 3.46e-8f * p;
    fespt = f * 611.21f * c_exp(17.502f * (t - 273.15f) / (t - 32.18f));
  } else {
    const float f = 1.0003f + 4.18e-8f * p;
    fespt = f * 611.15f * c_exp(22.452f * (t - 273.15f) / (t - 0.6f));
  }
  if (p - (1.0f - rddrv) * fespt == 0.f) return 1.f;
  return rddrv * fespt / (p - (1.0f - rddrv) * fespt);
}

namespace k { // thermodynamic constants of convect / tlift, src/convect43c.f90:262-275
constexpr float CPD = 1005.7f, CPV = 1870.0f, CL = 2500.0f, RV = 461.5f, RD = 287.04f, LV0 = 2.501e6f, G = 9.81f,
                ROWL = 1000.0f, CPVMCL = CL - CPV, EPS0 = RD / RV, EPSI = 1.f / EPS0, GINV = 1.0f / G, EPSILON = 1.e-20f;
}

// src/convect43c.f90:974-1090
FPB_HD inline void conv_tlift(ConvWork &w, int icb, int nk, int nl, int kk) {
  using namespace k;
  const float qnk = CV(qconv, nk), tnk = CV(tconv, nk);
  const float ah0 = (CPD * (1.f - qnk) + CL * qnk) * tnk + qnk * (LV0 - CPVMCL * (tnk - 273.15f)) + CV(gz, nk);
  const float cpp = CPD * (1.f - qnk) + qnk * CPV;
  const float cpinv = 1.f / cpp;
  if (kk == 1) {
    for (int i = 1; i <= icb - 1; i++) CV(clw, i) = 0.0f;
    for (int i = nk; i <= icb - 1; i++) {
      CV(tp, i) = tnk - (CV(gz, i) - CV(gz, nk)) * cpinv;
      CV(tvp, i) = CV(tp, i) * (1.f + qnk * EPSI);
    }
  }
  int nst = icb, nsb = icb;
  if (kk == 2) {
    nst = nl;
    nsb = icb + 1;
  }
  for (int i = nsb; i <= nst; i++) { // (the level's inputs read once, tp / clw formed in registers)
    const float t_i = CV(tconv, i), gz_i = CV(gz, i), p_i = CV(pconv_hpa, i);
    float tg = t_i, qg = CV(qsconv, i);
    float alv = LV0 - CPVMCL * (t_i - 273.15f);
    for (int j = 1; j <= 2; j++) {
      float s = CPD + alv * alv * qg / (RV * t_i * t_i);
      s = 1.f / s;
      const float ahg = CPD * tg + (CL - CPD) * qnk * t_i + alv * qg + gz_i;
      tg = tg + s * (ah0 - ahg);
      tg = c_max(tg, 35.0f);
      const float tc = tg - 273.15f;
      const float denom = 243.5f + tc;
      float es;
      if (tc >= 0.0f) es = 6.112f * c_exp(17.67f * tc / denom);
      else es = c_exp(23.33086f - 6111.72784f / tg + 0.15215f * c_log(tg));
      qg = EPS0 * es / (p_i - es * (1.f - EPS0));
    }
    alv = LV0 - CPVMCL * (t_i - 273.15f);
    const float tp_i = (ah0 - (CL - CPD) * qnk * t_i - gz_i - alv * qg) / CPD;
    CV(tp, i) = tp_i;
    float clw_i = qnk - qg;
    clw_i = c_max(0.0f, clw_i);
    CV(clw, i) = clw_i;
    const float rg = qg / (1.f - qnk);
    CV(tvp, i) = tp_i * (1.f + rg * EPSI);
  }
}

// src/convect43c.f90:11-972 in two halves around the flux assembly (:855-913), so that the device can run that
// O(n^3) part with a warp per column (fpb_convect.cu).  nl = nconvlev; cbmf in/out.
// conv_convect_a: everything up to the flux assembly.  false: the scheme is over, st.iflag is its result.
// conv_convect_head: everything in front of the loops over level pairs (:277-631): the O(n) part, one column after
// the other in every build.  false: the scheme is over, st.iflag is its result; true: st.{iflag,icb,inb,nk,delti} set.
FPB_HD inline bool conv_convect_head(ConvWork &w, int nl, float delt, float &cbmf, ConvState &st) {
  using namespace k;
  const int MINORIG = 1;
  const float ELCRIT = .0011f, TLCRIT = -55.0f, ENTP = 1.5f, DTMAX = 0.9f, ALPHA = 0.025f, DAMP = 0.1f;
  int iflag;
  const float delti = 1.0f / delt;

  for (int i = 1; i <= nl + 1; i++) {
    CV(ft, i) = 0.0f; CV(fq, i) = 0.0f; CV(fdown, i) = 0.0f; CV(sub, i) = 0.0f; CV(fup, i) = 0.0f;
    CV(m, i) = 0.0f;
  }
  // (FMASS, MENT, ELIJ, SIJ: the reference zeroes them over (NL+1)^2 here and at :529-545; they are
  //  only ever read inside [1, INB+2]^2, so they are zeroed there, below, once INB is known)
  // (TH, :299-307, is only read by the dry adiabatic adjustment the reference has commented out; like MP, WATER, EVAP,
  //  WT, QP, SIGP and LVCP -- read by the precipitating downdraft and the tendencies alone -- it is not formed)
  iflag = 0;

  // geopotential, heat capacity, static energies (:413-437)
  CV(gz, 1) = 0.0f;
  CV(cpn, 1) = CPD * (1.f - CV(qconv, 1)) + CV(qconv, 1) * CPV;
  CV(h, 1) = CV(tconv, 1) * CV(cpn, 1);
  CV(lv, 1) = LV0 - CPVMCL * (CV(tconv, 1) - 273.15f);
  CV(hm, 1) = CV(lv, 1) * CV(qconv, 1);
  CV(tv, 1) = CV(tconv, 1) * (1.f + CV(qconv, 1) * EPSI - CV(qconv, 1));
  float ahmin = 1.0e12f;
  int ihmin = nl;
  { // (the inputs of four levels requested together; what the recurrence carries -- gz, hm and the level below's
    //  temperature and humidity -- stays in registers instead of being read back after its store)
    const float t_1 = CV(tconv, 1);
    float gz_m = 0.0f, hm_m = CV(hm, 1), t_m = t_1, q_m = CV(qconv, 1), p_m = CV(pconv_hpa, 1);
    for (int i0 = 2; i0 <= nl + 1; i0 += 4) {
      float t_[4], q_[4], p_[4], ph_[4];
FPB_UNROLL(4)
      for (int u = 0; u < 4; u++) {
        const int i = i0 + u <= nl + 1 ? i0 + u : nl + 1;
        t_[u] = CV(tconv, i); q_[u] = CV(qconv, i); p_[u] = CV(pconv_hpa, i); ph_[u] = CV(phconv_hpa, i);
      }
FPB_UNROLL(4)
      for (int u = 0; u < 4; u++) {
        const int i = i0 + u;
        if (i <= nl + 1) {
          const float t_i = t_[u], q_i = q_[u];
          const float tvx = t_i * (1.f + q_i * EPSI - q_i);
          const float tvy = t_m * (1.f + q_m * EPSI - q_m);
          const float gz_i = gz_m + 0.5f * RD * (tvx + tvy) * (p_m - p_[u]) / ph_[u];
          CV(gz, i) = gz_i;
          const float cpn_i = CPD * (1.f - q_i) + CPV * q_i;
          CV(cpn, i) = cpn_i;
          CV(h, i) = t_i * cpn_i + gz_i;
          const float lv_i = LV0 - CPVMCL * (t_i - 273.15f);
          CV(lv, i) = lv_i;
          const float hm_i = (CPD * (1.f - q_i) + CL * q_i) * (t_i - t_1) + lv_i * q_i + gz_i;
          CV(hm, i) = hm_i;
          CV(tv, i) = t_i * (1.f + q_i * EPSI - q_i);
          if (i >= MINORIG && hm_i < ahmin && hm_i < hm_m) {
            ahmin = hm_i;
            ihmin = i;
          }
          gz_m = gz_i; hm_m = hm_i; t_m = t_i; q_m = q_i; p_m = p_[u];
        }
      }
    }
  }
  ihmin = ihmin < nl - 1 ? ihmin : nl - 1;
  float ahmax = 0.0f;
  int nk = MINORIG;
  for (int i = MINORIG; i <= ihmin; i++)
    if (CV(hm, i) > ahmax) {
      nk = i;
      ahmax = CV(hm, i);
    }
  if (CV(tconv, nk) < 250.0f || CV(qconv, nk) <= 0.0f || ihmin == (nl - 1)) {
    cbmf = 0.0f;
    st.iflag = 0; return false;
  }
  // lifted condensation level (:464-471)
  const float rh = CV(qconv, nk) / CV(qsconv, nk);
  const float chi = CV(tconv, nk) / (1669.0f - 122.0f * rh - CV(tconv, nk));
  const float plcl = CV(pconv_hpa, nk) * c_pow(rh, chi);
  if (plcl < 200.0f || plcl >= 2000.0f) {
    cbmf = 0.0f;
    st.iflag = 2; return false;
  }
  int icb = nl - 1;
  for (int i = nk + 1; i <= nl; i++)
    if (CV(pconv_hpa, i) < plcl) icb = icb < i ? icb : i;
  if (icb >= (nl - 1)) {
    cbmf = 0.0f;
    st.iflag = 3; return false;
  }
  conv_tlift(w, icb, nk, nl, 1);
  for (int i = nk; i <= icb; i++) CV(tvp, i) = CV(tvp, i) - CV(tp, i) * CV(qconv, nk);
  if (cbmf == 0.0f && CV(tvp, icb) <= (CV(tv, icb) - DTMAX)) { st.iflag = 0; return false; }
  if (iflag != 4) iflag = 1;
  conv_tlift(w, icb, nk, nl, 2);
  // precipitation efficiencies (:503-520)
  for (int i = 1; i <= nk; i++) {
    CV(ep, i) = 0.0f;
  }
  for (int i0 = nk + 1; i0 <= nl; i0 += 4) { // (four levels requested together; ep(i) formed in a register)
    float tp_[4], cl_[4];
FPB_UNROLL(4)
    for (int u = 0; u < 4; u++) {
      const int i = i0 + u <= nl ? i0 + u : nl;
      tp_[u] = CV(tp, i); cl_[u] = CV(clw, i);
    }
FPB_UNROLL(4)
    for (int u = 0; u < 4; u++) {
      if (i0 + u <= nl) {
        const float tca = tp_[u] - 273.15f;
        float elacrit;
        if (tca >= 0.0f) elacrit = ELCRIT;
        else elacrit = ELCRIT * (1.0f - tca / TLCRIT);
        elacrit = c_max(elacrit, 0.0f);
        const float epmax = 0.999f;
        float ep_i = epmax * (1.0f - elacrit / c_max(cl_[u], 1.0e-8f));
        ep_i = c_max(ep_i, 0.0f);
        ep_i = c_min(ep_i, epmax);
        CV(ep, i0 + u) = ep_i;
      }
    }
  }
  { // (four levels requested together: a store into one vector keeps the compiler from moving the next level's loads
    //  ahead of it, and the column kernel is bound by the latency of its loads)
    const float q_nk = CV(qconv, nk);
    for (int i0 = icb + 1; i0 <= nl; i0 += 4) {
      float a_[4], b_[4];
FPB_UNROLL(4)
      for (int u = 0; u < 4; u++) {
        const int i = i0 + u <= nl ? i0 + u : nl;
        a_[u] = CV(tvp, i); b_[u] = CV(tp, i);
      }
FPB_UNROLL(4)
      for (int u = 0; u < 4; u++)
        if (i0 + u <= nl) CV(tvp, i0 + u) = a_[u] - b_[u] * q_nk;
    }
  }
  CV(tvp, nl + 1) = CV(tvp, nl) - (CV(gz, nl + 1) - CV(gz, nl)) / CPD;
  // initialise the work arrays (:529-545)
  for (int i = 1; i <= nl + 1; i++) {
    CV(hp, i) = CV(h, i);
    CV(nent, i) = 0;
  }
  // level of neutral buoyancy (:549-573)
  float cape = 0.0f, capem = 0.0f, byp = 0.0f;
  int inb = icb + 1, inb1 = inb;
  // (the buoyancy terms of five levels at a time, their elements requested together; BYP of level i is the same
  //  expression as BY of level i+1, so it is that value)
  for (int i0 = icb + 1; i0 <= nl - 1; i0 += 4) {
    float by_[5];
    {
      float a_[5], b_[5], p_[5], h_[6];
FPB_UNROLL(5)
      for (int u = 0; u < 5; u++) {
        const int i = i0 + u <= nl ? i0 + u : nl;
        a_[u] = CV(tvp, i); b_[u] = CV(tv, i); p_[u] = CV(pconv_hpa, i);
      }
FPB_UNROLL(6)
      for (int u = 0; u < 6; u++) h_[u] = CV(phconv_hpa, (i0 + u <= nl + 1 ? i0 + u : nl + 1));
FPB_UNROLL(5)
      for (int u = 0; u < 5; u++) by_[u] = (a_[u] - b_[u]) * (h_[u] - h_[u + 1]) / p_[u];
    }
FPB_UNROLL(4)
    for (int u = 0; u < 4; u++) {
      const int i = i0 + u;
      if (i <= nl - 1) {
        const float by = by_[u];
        cape = cape + by;
        if (by >= 0.0f) inb1 = i + 1;
        if (cape > 0.0f) {
          inb = i + 1;
          byp = by_[u + 1];
          capem = cape;
        }
      }
    }
  }
  inb = inb > inb1 ? inb : inb1;
  cape = capem + byp;
  float defrac = capem - cape;
  defrac = c_max(defrac, 0.001f);
  float frac = -cape / defrac;
  frac = c_min(frac, 1.0f);
  frac = c_max(frac, 0.0f);
  {
    const float h_nk = CV(h, nk);
    for (int i0 = icb; i0 <= inb; i0 += 4) {
      float l_[4], t_[4], e_[4], c_[4];
FPB_UNROLL(4)
      for (int u = 0; u < 4; u++) {
        const int i = i0 + u <= inb ? i0 + u : inb;
        l_[u] = CV(lv, i); t_[u] = CV(tconv, i); e_[u] = CV(ep, i); c_[u] = CV(clw, i);
      }
FPB_UNROLL(4)
      for (int u = 0; u < 4; u++)
        if (i0 + u <= inb) CV(hp, i0 + u) = h_nk + (l_[u] + (CPD - CPV) * t_[u]) * e_[u] * c_[u];
    }
  }
  // cloud base mass flux (:583-611)
  float dbosum = 0.0f;
  const float tvpplcl = CV(tvp, icb - 1) - RD * CV(tvp, icb - 1) * (CV(pconv_hpa, icb - 1) - plcl) /
                                               (CV(cpn, icb - 1) * CV(pconv_hpa, icb - 1));
  const float tvaplcl = CV(tv, icb) + (CV(tvp, icb) - CV(tvp, icb + 1)) * (plcl - CV(pconv_hpa, icb)) /
                                          (CV(pconv_hpa, icb) - CV(pconv_hpa, icb + 1));
  float dtpbl = 0.0f;
  for (int i = nk; i <= icb - 1; i++)
    dtpbl = dtpbl + (CV(tvp, i) - CV(tv, i)) * (CV(phconv_hpa, i) - CV(phconv_hpa, i + 1));
  dtpbl = dtpbl / (CV(phconv_hpa, nk) - CV(phconv_hpa, icb));
  const float dtmin = tvpplcl - tvaplcl + DTMAX + dtpbl;
  const float dtma = dtmin;
  const float cbmfold = cbmf;
  const float delt0 = delt / 3.f;
  const float damps = DAMP * delt / delt0;
  cbmf = (1.f - damps) * cbmf + 0.1f * ALPHA * dtma;
  cbmf = c_max(cbmf, 0.0f);
  if (cbmf == 0.0f && cbmfold == 0.0f) { st.iflag = iflag; return false; }
  // rates of mixing (:621-631)
  CV(m, icb) = 0.0f;
  for (int i = icb + 1; i <= inb; i++) {
    const int kq = i < inb1 ? i : inb1;
    const float dbo = fabsf(CV(tv, kq) - CV(tvp, kq)) + ENTP * 0.02f * (CV(phconv_hpa, kq) - CV(phconv_hpa, kq + 1));
    dbosum = dbosum + dbo;
    CV(m, i) = cbmf * dbo;
  }
  for (int i = icb + 1; i <= inb; i++) CV(m, i) = CV(m, i) / dbosum;
  // what depends on j alone in the loop over level pairs below -- bf2 and cwat of the reference's inner loop -- is
  // worked out once per level (into the unused ft / fq vectors: the same expressions, so the same bits)
  for (int j0 = icb; j0 <= inb; j0 += 4) {
    float l_[4], q_[4], t_[4], c_[4], e_[4];
FPB_UNROLL(4)
    for (int u = 0; u < 4; u++) {
      const int j = j0 + u <= inb ? j0 + u : inb;
      l_[u] = CV(lv, j); q_[u] = CV(qsconv, j); t_[u] = CV(tconv, j); c_[u] = CV(clw, j); e_[u] = CV(ep, j);
    }
FPB_UNROLL(4)
    for (int u = 0; u < 4; u++)
      if (j0 + u <= inb) {
        CV(ft, j0 + u) = 1.f + l_[u] * l_[u] * q_[u] / (RV * t_[u] * t_[u] * CPD);
        CV(fq, j0 + u) = c_[u] * (1.f - e_[u]);
      }
  }
  (void)frac;
  st.iflag = iflag; st.inb = inb; st.icb = icb; st.nk = nk; st.delti = delti;
  return true;
}

// The loops over level pairs (:636-682 mixing fractions, :686-746 normalisation) row by row: row i of
// SIJ / MENT / ELIJ and NENT(i) depend on the column's vectors and on row i alone, so the rows can be worked in any
// order -- or by different threads (conv_mix_kernel) -- with the reference's bits.
// Nothing is zeroed up front (the reference zeroes the four (NL+1)^2 matrices, :529-545): row i of MENT is written in
// full over the columns ICB..INB it can be set in (the value or 0), every other element of MENT is known to be 0 by its
// indices (conv_ment_set) and FMASS is written in full over [1, nconvtop]^2 by conv_fmass_row.
FPB_HD inline bool conv_ment_set(const ConvState &st, int i, int j) { // may MENT(i,j) differ from 0?
  return i >= st.icb + 1 && i <= st.inb && j >= st.icb && j <= st.inb;
}

// Row i of the entrained air mass flux: mixing fractions (:636-682) and normalisation (:686-746) in ONE walk along the
// row.  The normalisation of element j reads sij(i,j-1), sij(i,j), sij(i,j+1) and ment(i,j) of the mixing loop and
// nothing else of the matrices, and SIJ is read by nothing after it, so the mixing loop runs one element ahead and
// hands its values over in registers: SIJ is not stored at all, MENT(i,j) once.  (The reference normalises a row only
// when nent(i) != 0: with nent(i) == 0 no element of the row has 0 < sij < 0.9 and the walk changes nothing.)
// rowtop(i) notes how far the final row reaches above EPSILON (what conv_convect_b's search for nconvtop asks).
#ifndef FPB_MIX_BATCH
#define FPB_MIX_BATCH 2 // (the device stages these vectors in shared memory: two loads in flight and 64 registers beat eight and 128)
#endif
constexpr int CONV_MIX_BATCH = FPB_MIX_BATCH;
#ifndef FPB_SCALE_BATCH
#define FPB_SCALE_BATCH 8
#endif
constexpr int CONV_SCALE_BATCH = FPB_SCALE_BATCH;
FPB_HD inline void conv_mixnorm_row(ConvWork &w, const ConvState &st, int i) {
  using namespace k;
  const int icb = st.icb, inb = st.inb, nk = st.nk;
  const float qti = CV(qconv, nk) - CV(ep, i) * CV(clw, i);
  const float hp_i = CV(hp, i), h_i = CV(h, i), q_i = CV(qconv, i), m_i = CV(m, i);
  int nent_i = CV(nent, i);
  // scrit of the normalisation (:690-700; qp1 is the mixing loop's qti)
  float scrit;
  {
    const float lv_i = CV(lv, i), qs_i = CV(qsconv, i);
    const float anum = h_i - hp_i - lv_i * (qti - qs_i);
    float denom = h_i - hp_i + lv_i * (q_i - qti);
    if (fabsf(denom) < 0.01f) denom = 0.01f;
    scrit = anum / denom;
    const float alt = qti - qs_i + scrit * (q_i - qti);
    if (alt < 0.0f) scrit = 1.0f;
    scrit = c_max(scrit, 0.0f);
  }
  float asij = 0.0f, smin = 1.0f;
  float s_m1 = 0.0f, s_0 = 0.0f, m_0 = 0.0f; // sij(i,j-2), sij(i,j-1), ment(i,j-1) while element j is mixed
  float php = 0.0f;                          // phconv_hpa(j-1)
  // CONV_MIX_BATCH levels at a time: their vector elements are requested together (the loop is bound by the latency of
  // these loads), then worked through in order; element inb+1 only closes the window (sij(i,inb+1) = 0)
  for (int j0 = icb; j0 <= inb + 1; j0 += CONV_MIX_BATCH) {
    float bf2_[CONV_MIX_BATCH], t_[CONV_MIX_BATCH], qs_[CONV_MIX_BATCH], hj_[CONV_MIX_BATCH], qj_[CONV_MIX_BATCH],
        cw_[CONV_MIX_BATCH], lv_[CONV_MIX_BATCH], ph_[CONV_MIX_BATCH];
FPB_UNROLL(CONV_MIX_BATCH)
    for (int u = 0; u < CONV_MIX_BATCH; u++) {
      const int j = j0 + u <= inb ? j0 + u : inb;
      bf2_[u] = CV(ft, j); t_[u] = CV(tconv, j); qs_[u] = CV(qsconv, j); hj_[u] = CV(h, j); qj_[u] = CV(qconv, j);
      cw_[u] = CV(fq, j); lv_[u] = CV(lv, j);
      ph_[u] = CV(phconv_hpa, (j0 + u <= inb + 1 ? j0 + u : inb + 1));
    }
FPB_UNROLL(CONV_MIX_BATCH)
    for (int u = 0; u < CONV_MIX_BATCH; u++) {
      const int j = j0 + u;
      if (j <= inb + 1) {
        float s = 0.0f, mraw = 0.0f;
        if (j <= inb) { // mixing fraction of (i,j)
          const float bf2 = bf2_[u];
          const float t_j = t_[u], qs_j = qs_[u];
          float anum = hj_[u] - hp_i + (CPV - CPD) * t_j * (qti - qj_[u]);
          float denom = h_i - hp_i + (CPD - CPV) * (q_i - qti) * t_j;
          float dei = denom;
          if (fabsf(dei) < 0.01f) dei = 0.01f;
          s = anum / dei;
          if (j == i) s = 1.0f; // (sij(i,i) = 1)
          float altem = s * q_i + (1.f - s) * qti - qs_j;
          altem = altem / bf2;
          const float cwat = cw_[u];
          const float stemp = s;
          if ((stemp < 0.0f || stemp > 1.0f || altem > cwat) && j > i) {
            const float lv_j = lv_[u];
            anum = anum - lv_j * (qti - qs_j - cwat * bf2);
            denom = denom + lv_j * (q_i - qti);
            if (fabsf(denom) < 0.01f) denom = 0.01f;
            s = anum / denom;
          }
          if (s > 0.0f && s < 0.9f) { // (elij(i,j) = max(0, altem) is not kept)
            mraw = m_i / (1.f - s);
            nent_i = nent_i + 1;
          }
          s = c_max(0.0f, s);
          s = c_min(1.0f, s);
        }
        const int jn = j - 1; // normalisation step of (i,jn): window s_m1, s_0, s
        if (jn >= icb) {
          float out = m_0;
          if (s_0 > 0.0f && s_0 < 0.9f) {
            float smid, sjmax, sjmin;
            if (jn > i) {
              smid = c_min(s_0, scrit);
              sjmax = smid;
              sjmin = smid;
              if (smid < smin && s < smid) {
                smin = smid;
                sjmax = c_min(c_min(s, s_0), scrit);
                sjmin = c_max(s_m1, s_0);
                sjmin = c_min(sjmin, scrit);
              }
            } else {
              sjmax = c_max(s, scrit);
              smid = c_max(s_0, scrit);
              sjmin = 0.0f;
              if (jn > 1) sjmin = s_m1;
              sjmin = c_max(sjmin, scrit);
            }
            const float delp = fabsf(sjmax - smid);
            const float delm = fabsf(sjmin - smid);
            asij = asij + (delp + delm) * (php - ph_[u]);
            out = m_0 * (delp + delm) * (php - ph_[u]);
          }
          CM(ment, i, jn) = out;
        }
        s_m1 = s_0; s_0 = s; m_0 = mraw; php = ph_[u];
      }
    }
  }
  int top = 0;      // largest j != i with ment(i,j) > EPSILON
  float vii = m_i;  // ment(i,i)
  if (nent_i == 0) {
    CM(ment, i, i) = m_i; // (the rest of the row is 0)
  } else {
    asij = c_max(1.0e-21f, asij);
    asij = 1.0f / asij;
    float bsum = 0.0f; // (the reference's two loops -- scale the row, then add it up in the same order -- in one)
    for (int j0 = icb; j0 <= inb; j0 += CONV_SCALE_BATCH) { // (the row's elements requested CONV_SCALE_BATCH at a time)
      float mv[CONV_SCALE_BATCH];
FPB_UNROLL(CONV_SCALE_BATCH)
      for (int u = 0; u < CONV_SCALE_BATCH; u++) mv[u] = CM(ment, i, (j0 + u <= inb ? j0 + u : inb));
FPB_UNROLL(CONV_SCALE_BATCH)
      for (int u = 0; u < CONV_SCALE_BATCH; u++) {
        const int j = j0 + u;
        if (j <= inb) {
          const float v = mv[u] * asij;
          CM(ment, i, j) = v;
          bsum = bsum + v;
          if (j == i) vii = v;
          else if (v > EPSILON) top = j;
        }
      }
    }
    if (bsum < 1.0e-18f) {
      nent_i = 0;
      CM(ment, i, i) = m_i;
      vii = m_i;
    }
  }
  CV(nent, i) = nent_i;
  if (vii > EPSILON && i > top) top = i;
  CV(rowtop, i) = top > 0 ? (top > i ? top : i) : 0;
}

// conv_convect_a: everything up to the flux assembly, one column sequentially (host build, sequential path).
// The precipitating downdraft (:750-842) is left out like the tendencies FT, FQ (:867-872,914-934) it feeds: WATER,
// EVAP, MP, QP and PRECIP are read by nothing FLEXPART uses (FMASS, SUB, IFLAG, CBMF); IFLAG = 4 (the CFL condition
// on the subsidence, in the flux assembly) is kept.
FPB_HD inline bool conv_convect_a(ConvWork &w, int nl, float delt, float &cbmf, ConvState &st) {
  if (!conv_convect_head(w, nl, delt, cbmf, st)) return false;
  for (int i = st.icb + 1; i <= st.inb; i++) conv_mixnorm_row(w, st, i);
  st.go = 1;
  return true;
}

// the flux assembly in the reference's order (the definition of what conv_assembly_kernel computes; the host
// build and the sequential path run it as it stands)
FPB_HD inline void conv_flux_assembly(ConvWork &w, ConvState &st) {
  using namespace k;
  const int inb = st.inb, icb = st.icb, nk = st.nk;
  const float delti = st.delti;
  int iflag = st.iflag;
  float dpinv = 0.01f / (CV(phconv_hpa, 1) - CV(phconv_hpa, 2));
  float am = 0.0f;
  if (nk == 1)
    for (int kq = 2; kq <= inb; kq++) am = am + CV(m, kq);
  CV(fup, 1) = am;
  if ((2.f * G * dpinv * am) >= delti) iflag = 4;
  for (int i = 2; i <= inb; i++) {
    dpinv = 0.01f / (CV(phconv_hpa, i) - CV(phconv_hpa, i + 1));
    float amp1 = 0.0f, ad = 0.0f;
    if (i >= nk)
      for (int kq = i + 1; kq <= inb + 1; kq++) amp1 = amp1 + CV(m, kq);
    // (rows 1..ICB of MENT are zero -- only rows ICB+1..INB are ever set -- and x + 0.0 == x)
    for (int kq = icb + 1; kq <= i; kq++) {
FPB_UNROLL(8)
      for (int j = i + 1; j <= inb + 1; j++) amp1 = amp1 + (j <= inb ? CM(ment, kq, j) : 0.0f); // (column inb+1 is 0)
    }
    CV(fup, i) = amp1;
    if ((2.f * G * dpinv * amp1) >= delti) iflag = 4;
    for (int kq = icb; kq <= i - 1; kq++) { // (columns 1..ICB-1 of MENT are zero)
FPB_UNROLL(8)
      for (int j = (i > icb + 1 ? i : icb + 1); j <= inb; j++) ad = ad + CM(ment, j, kq);
    }
    CV(fdown, i) = ad;
  }
  st.iflag = iflag;
}

// conv_convect_b: mass displacement matrix and compensating subsidence (:972-989); returns iflag.  FMASS(j,i) =
// 0 (+ M(i) when j == NK) + MENT(j,i) itself is formed by conv_fmass_row together with calcmatrix's scaling; here only
// what the loop finds out about it: nconvtop = 1 + the largest index of an element above EPSILON (rows ICB+1..INB:
// rowtop; row NK < ICB+1: M(i) alone; every other row is 0).
FPB_HD inline int conv_convect_b(ConvWork &w, const ConvState &st, bool with_sub = true) {
  using namespace k;
  const int inb = st.inb, icb = st.icb, nk = st.nk;
  int nconvtop = 1;
  for (int i = 1; i <= inb + 1; i++) {
    float f = 0.0f;
    f = f + CV(m, i);
    if (f > EPSILON) {
      nconvtop = nconvtop > i ? nconvtop : i;
      nconvtop = nconvtop > nk ? nconvtop : nk;
    }
  }
  for (int i = icb + 1; i <= inb; i++) {
    const int t = CV(rowtop, i);
    nconvtop = nconvtop > t ? nconvtop : t;
  }
  if (with_sub) {
    CV(sub, 1) = 0.f;
    for (int i = 2; i <= inb + 1; i++) CV(sub, i) = CV(fup, i - 1) - CV(fdown, i);
  }
  w.nconvtop = nconvtop + 1;
  return st.iflag;
}

// the whole scheme, sequentially; returns iflag
FPB_HD inline int conv_convect(ConvWork &w, int nl, float delt, float &cbmf) {
  ConvState st;
  st.go = 0;
  if (!conv_convect_a(w, nl, delt, cbmf, st)) return st.iflag;
  conv_flux_assembly(w, st);
  return conv_convect_b(w, st);
}

// src/calcmatrix.f90:45-135 (ECMWF branch), like conv_convect in two halves around the flux assembly.
// cbmf = cbaseflux(ix,jy), in/out.  tconv(1..nuvz-1), qconv(1..nuvz-1) and psconv must be set.
// conv_calcmatrix_a: pressures, saturation humidity, the scheme up to the flux assembly (st.go: it is due)
// head_only: stop in front of the loops over level pairs (the device runs them in conv_mix_kernel)
// conv_calcmatrix_level: what calcmatrix (:66-89) works out for level kq = 1 .. nuvz-1, from tconv(kq) and psconv alone
// (the half-level pressure below is formed again by its own expression instead of being read back), so the levels can
// go to different threads (conv_pre_kernel)
FPB_HD inline void conv_calcmatrix_level(ConvWork &w, int kq) {
  const int kuvz = kq + 1;
  const float ph_lo = kq == 1 ? w.psconv : (w.akm[kq] + w.bkm[kq] * w.psconv);
  const float ph_hi = (w.akm[kuvz] + w.bkm[kuvz] * w.psconv);
  if (kq == 1) CV(phconv, 1) = ph_lo;
  CV(pconv, kq) = (w.akz[kuvz] + w.bkz[kuvz] * w.psconv);
  CV(phconv, kuvz) = ph_hi;
  CV(dpr, kq) = ph_lo - ph_hi;
  CV(qsconv, kq) = conv_qvsat(CV(pconv, kq), CV(tconv, kq));
  if (kq <= w.nconvlev + 1) {
    CV(pconv_hpa, kq) = CV(pconv, kq) / 100.f;
    CV(phconv_hpa, kq) = ph_lo / 100.f;
  }
}

// levels_done: conv_calcmatrix_level has run for every level
FPB_HD inline void conv_calcmatrix_a(ConvWork &w, float delt, float &cbmf, ConvState &st, bool head_only = false,
                                     bool levels_done = false) {
  const int nuvz = w.nuvz, nconvlev = w.nconvlev;
  if (!levels_done)
    for (int kq = 1; kq <= nuvz - 1; kq++) conv_calcmatrix_level(w, kq);
  st.cbmfold = cbmf;
  w.nconvtop = 0;
  st.go = 0;
  st.inb = st.icb = st.nk = 0;
  st.delti = 0.f;
  if (head_only ? conv_convect_head(w, nconvlev, delt, cbmf, st) : conv_convect_a(w, nconvlev, delt, cbmf, st)) st.go = 1;
  st.cbmf = cbmf;
}

// row kq of fmassfrac (src/calcmatrix.f90:118-131 on the FMASS of :972-985): fmassfrac(kq,kk) = delt*fmass(kq,kk), plus
// what stays in the level on the diagonal
// ment_at(i, j): MENT(i,j) where conv_ment_set(st, i, j) (the device reads the shared-memory copy of the flux assembly)
template <class MentAt>
FPB_HD inline void conv_fmass_row(ConvWork &w, const ConvState &st, float delt, int kq, MentAt ment_at) {
  const float ga = 9.81f;
  const int inb = st.inb, nk = st.nk;
  const float rlevmass = CV(dpr, kq) / ga;
  float summe = 0.f, vd = 0.f;
  for (int kk0 = 1; kk0 <= w.nconvtop; kk0 += 4) {
    float mv[4], ev[4];
FPB_UNROLL(4)
    for (int u = 0; u < 4; u++) { // (four elements requested together, then in order)
      const int kk = kk0 + u;
      ev[u] = conv_ment_set(st, kq, kk) ? ment_at(kq, kk) : 0.0f;
      mv[u] = (kq == nk && kk <= inb + 1) ? CV(m, kk) : 0.0f;
    }
FPB_UNROLL(4)
    for (int u = 0; u < 4; u++) {
      const int kk = kk0 + u;
      if (kk <= w.nconvtop) {
        float f = 0.0f;
        if (kq == nk && kk <= inb + 1) f = f + mv[u];
        f = f + ev[u];
        const float v = delt * f;
        CF(kq, kk) = v;
        summe = summe + v;
        if (kk == kq) vd = v;
      }
    }
  }
  CF(kq, kq) = vd + rlevmass - summe;
}

// conv_calcmatrix_b: the rest of the scheme (when the assembly ran) and the redistribution matrix.  Returns lconv.
// rows: also write the matrix (the device has its rows written by conv_mix_kernel); with_sub: also the subsidence
// (needs FUP / FDOWN of the flux assembly)
FPB_HD inline bool conv_calcmatrix_b(ConvWork &w, float delt, float &cbmf, const ConvState &st, bool rows = true,
                                     bool with_sub = true) {
  const float cbmfold = st.cbmfold;
  cbmf = st.cbmf;
  const int iflag = st.go ? conv_convect_b(w, st, with_sub) : st.iflag;
  if (iflag != 1 && iflag != 4) {
    cbmf = cbmfold;
    return false;
  }
  if (cbmf <= 0.f && cbmfold <= 0.f) {
    cbmf = cbmfold;
    return false;
  }
  if (rows)
    for (int kq = 1; kq <= w.nconvtop; kq++)
      conv_fmass_row(w, st, delt, kq, [&](int i, int j) { return CM(ment, i, j); });
  return true;
}

// the whole routine, sequentially
FPB_HD inline bool conv_calcmatrix(ConvWork &w, float delt, float &cbmf) {
  ConvState st;
  conv_calcmatrix_a(w, delt, cbmf, st);
  if (st.go) conv_flux_assembly(w, st);
  return conv_calcmatrix_b(w, delt, cbmf, st);
}

// src/redist.f90:63-118: heights above ground of the eta half levels of the column
FPB_HD inline void conv_uvzlev(ConvWork &w) {
  const float cnst = 287.05f / 9.81f; // r_air/ga
  float tvold = w.tt2conv * (1.f + 0.378f * conv_ew(w.td2conv) / w.psconv);
  float pold = w.psconv;
  CV(uvzlev, 1) = 0.f;
  float pint = CV(phconv, 2);
  float tv1 = CV(tconv, 1) * (1.f + 0.608f * CV(qconv, 1));
  float tv2 = CV(tconv, 2) * (1.f + 0.608f * CV(qconv, 2));
  float tv = tv1 + (tv2 - tv1) * (CV(pconv, 1) - CV(phconv, 2)) / (CV(pconv, 1) - CV(pconv, 2));
  if (fabsf(tv - tvold) > 0.2f) CV(uvzlev, 2) = CV(uvzlev, 1) + cnst * c_log(pold / pint) * (tv - tvold) / c_log(tv / tvold);
  else CV(uvzlev, 2) = CV(uvzlev, 1) + cnst * c_log(pold / pint) * tv;
  tvold = tv;
  tv1 = tv2;
  pold = pint;
  for (int kz = 3; kz <= w.nconvtop + 1; kz++) {
    pint = CV(phconv, kz);
    tv2 = CV(tconv, kz) * (1.f + 0.608f * CV(qconv, kz));
    tv = tv1 + (tv2 - tv1) * (CV(pconv, kz - 1) - CV(phconv, kz)) / (CV(pconv, kz - 1) - CV(pconv, kz));
    if (fabsf(tv - tvold) > 0.2f)
      CV(uvzlev, kz) = CV(uvzlev, kz - 1) + cnst * c_log(pold / pint) * (tv - tvold) / c_log(tv / tvold);
    else
      CV(uvzlev, kz) = CV(uvzlev, kz - 1) + cnst * c_log(pold / pint) * tv;
    tvold = tv;
    tv1 = tv2;
    pold = pint;
  }
}

// level of the particle in the column, src/redist.f90:122-131; 0: above the convective domain
FPB_HD inline int conv_levold(const ConvWork &w, float ztold) {
  for (int kz = 2; kz <= w.nconvtop; kz++)
    if (CV(uvzlev, kz) >= ztold) return kz - 1;
  return 0;
}

// src/redist.f90:120-237 for one particle (levold > 0): rn = the uniform of `ran3(iseed)`
FPB_HD inline float conv_redist(const ConvWork &w, float ztold, int levold, float rn, int ldirect, int lsynctime) {
  const float ga = 9.81f, r_air = 287.05f;
  float z = ztold, dlevfrac = 0.5f;
  int levnew = levold;
  float ffraction = 0.f;
  const float totlevmass = CV(dpr, levold) / ga;
  for (int kq = 1; kq <= w.nconvtop; kq++) {
    const float f = (ldirect == 1) ? CF(levold, kq) : CF(kq, levold);
    ffraction = ffraction + f / totlevmass;
    if (rn <= ffraction) {
      levnew = kq;
      if (ffraction > 1.e-20f) dlevfrac = (ffraction - rn) / f * totlevmass;
      else dlevfrac = 0.5f;
      break;
    }
  }
  if (levnew <= w.nconvtop) {
    if (levnew == levold) {
      z = ztold;
    } else {
      const float dlogp = (1.f - dlevfrac) * (c_log(CV(phconv, levnew + 1)) - c_log(CV(phconv, levnew)));
      const float pint = c_log(CV(phconv, levnew)) + dlogp;
      const float dz1 = pint - c_log(CV(phconv, levnew));
      const float dz2 = c_log(CV(phconv, levnew + 1)) - pint;
      const float dz = dz1 + dz2;
      z = (CV(uvzlev, levnew) * dz2 + CV(uvzlev, levnew + 1) * dz1) / dz;
      if (z < 0.f) z = -1.f * z;
    }
  }
  if (levnew <= w.nconvtop && levnew == levold) { // compensating subsidence
    const float zo = z;
    float wsub_lo, wsub_hi;
    if (levold > 1) {
      const float temp_levold = CV(tconv, levold - 1) + (CV(tconv, levold) - CV(tconv, levold - 1)) *
                                                            (CV(pconv, levold - 1) - CV(phconv, levold)) /
                                                            (CV(pconv, levold - 1) - CV(pconv, levold));
      const float sub_levold = CV(sub, levold) / (1.f - CV(sub, levold) / CV(dpr, levold) * ga);
      wsub_lo = -1.f * sub_levold * r_air * temp_levold / (CV(phconv, levold));
    } else {
      wsub_lo = 0.f;
    }
    const float temp_levold1 = CV(tconv, levold) + (CV(tconv, levold + 1) - CV(tconv, levold)) *
                                                       (CV(pconv, levold) - CV(phconv, levold + 1)) /
                                                       (CV(pconv, levold) - CV(pconv, levold + 1));
    const float sub_levold1 = CV(sub, levold + 1) / (1.f - CV(sub, levold + 1) / CV(dpr, levold + 1) * ga);
    wsub_hi = -1.f * sub_levold1 * r_air * temp_levold1 / (CV(phconv, levold + 1));
    const float dz1 = zo - CV(uvzlev, levold);
    const float dz2 = CV(uvzlev, levold + 1) - zo;
    const float dz = dz1 + dz2;
    const float wsubpart = (dz2 * wsub_lo + dz1 * wsub_hi) / dz;
    z = zo + wsubpart * (float)lsynctime;
    if (z < 0.f) z = -1.f * z;
  }
  return z;
}

#undef CV
#undef CM
#undef CF

} // namespace fpbconv
