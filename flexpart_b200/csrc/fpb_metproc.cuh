// fpb_metproc.cuh -- what getfields does to a freshly read wind field before the particle loop may
// use it (src/getfields.f90:126-129), one grid column per thread:
//   met_levels_column   src/verttransform_ecmwf.f90:198-231   heights of the eta levels (uvzlev)
//   met_interp_column   src/verttransform_ecmwf.f90:233-472,:530-542,:610-724
//                       u, v, T, q, PV, density on the terrain-following height levels, the vertical
//                       wind (eta-dot -> m/s, plus the slope term of the eta surfaces), the density
//                       gradient, polar-stereographic winds poleward of the switch latitudes, and
//                       the cloud / precipitation classes (parameterised from the humidity, or from the
//                       cloud water of the input: readclouds)
//   met_pole_level      src/verttransform_ecmwf.f90:474-527,:544-607   the pole rows
//   met_calcpar_column  src/calcpar.f90:78-258 with scalev.f90, obukhov.f90, richardson.f90:
//                       friction velocity, Obukhov length, mixing height, convective velocity scale,
//                       thermal tropopause
//   met_theta_column, met_calcpv_column, met_pv_pole_level   src/calcpv.f90:42-315: potential
//                       vorticity on the eta levels (isentropic differences of u and v)
// Arithmetic is the reference's, statement by statement, evaluated like the validation build of the
// other kernels (no FMA contraction; exp/log/pow/sin/cos/atan in double, rounded once), so the
// fields are bit-comparable with the reference's own routines (oracle/_ref).  Written for nvcc and,
// for the CPU-side cross-check of tests/test_metproc.py only, for g++.
#pragma once
#include <math.h>
#include <stdint.h>
#ifndef __CUDACC__
#include <vector_types.h>
#endif

#include "fpb_convect.cuh" // conv_ew, conv_qvsat, c_log, c_pow, ...

namespace fpbmet {
using fpbconv::c_log;
using fpbconv::c_max;
using fpbconv::c_min;
using fpbconv::c_pow;
using fpbconv::c_sqrt;
using fpbconv::conv_ew;
using fpbconv::conv_qvsat;

FPB_HD inline float m_sin(float x) { return (float)sin((double)x); }
FPB_HD inline float m_cos(float x) { return (float)cos((double)x); }
FPB_HD inline float m_atan(float x) { return (float)atan((double)x); }

constexpr float M_R_AIR = 287.05f, M_GA = 9.81f, M_CPA = 1004.6f, M_PI_F = 3.14159265f;
constexpr float M_KARMAN = 0.40f, M_CONVKE = 2.0f, M_HMIXMIN = 100.f, M_HMIXMAX = 4500.f;

struct MetGrid {
  int nx, ny, nz, nuvz, nwz; // grid points used (nx = nxmin1 + 1, ny = nymin1 + 1); nz = nuvz levels
  int nxd, nyd;              // row length / rows of the device arrays
  float dx, dy, xlon0, ylat0, dxconst, dyconst;
  int nglobal, sglobal, xglobal;
  float switchnorthg, switchsouthg;
  float northpolemap[9], southpolemap[9];
  int lsubgrid;
  int readclouds;            // (0: parameterised clouds)
  int nest;                  // 1: a nested input grid (calcpar_nests / verttransform_nests / calcpv_nests): the same
                             // column arithmetic with the nest's geometry, no poles, no wrap
  float xresol, yresol;      // nest: xresoln(l), yresoln(l) of the slope term (src/verttransform_nests.f90:322-323)
  const float *akz, *bkz, *akm, *bkm; // 1-based, [nuvz + 1]
  const float *height;       // height(k) = height[k - 1]
  const float *cosf;         // [ny] 1 / cos(latitude of row jy), rows 1 .. ny-2 (:406-408)
  // the wind field as readwind_ecmwf leaves it, [k][jy][ix]
  const float2 *UV;          // {uuh, vvh}, nuvz levels
  const float *W;            // wwh, nwz levels
  const float2 *TQ;          // {tth, qvh}
  float *PV;                 // pvh: given by the caller or computed by met_calcpv_column (null: pv = 0)
  float *theta;              // work (calcpv): potential temperature on the eta levels, or null
  const float4 *SF1;         // {ps, tt2, td2, sshf}
  const float4 *SF2;         // {surfstr, lsprec, convprec, tcc}
  const float *excessoro;    // lsubgrid = 1 only
  const float *CLW, *CIW;    // readclouds: clwch (and ciwch unless the input holds their sum), nuvz levels
  float *clw;                // readclouds work: cloud water of the column's layers, nz levels
  float *uvzlev;             // work: nuvz levels
  // the met slot the particle loop reads
  float4 *A;                 // {uu, vv, ww, rho}
  float *G, *T;              // drhodz, tt
  float2 *P;                 // {uupol, vvpol}
  float4 *S;                 // {hmix, ustar, wstar, oli}
  float *trop;
  float4 *R;                 // {lsprec, convprec, tcc, ctwc} or null
  int8_t *Cl;                // clouds or null
  float2 *Q;                 // {pv, qv} or null
};

FPB_HD inline size_t m_o2(const MetGrid &g, int ix, int jy) { return (size_t)jy * g.nxd + ix; }
FPB_HD inline size_t m_o3(const MetGrid &g, int ix, int jy, int k /*1-based*/) {
  return ((size_t)(k - 1) * g.nyd + jy) * g.nxd + ix;
}

// src/cmapf_mod.f90:23-52 (cc2gll) with cspanf :494-524
FPB_HD inline float m_cspanf(float value, float begin, float end) {
  const float first = c_min(begin, end), last = c_max(begin, end);
  const float val = fmodf(value - first, last - first);
  if (val <= 0.f) return val + last;
  return val + first;
}
FPB_HD inline void m_cc2gll(const float *strcmp, float xlat, float xlong, float ue, float vn, float &ug, float &vg) {
  const double radpdg = (double)(M_PI_F / 180.f); // a default-real parameter, src/cmapf_mod.f90:20
  const double along = (double)m_cspanf(xlong - strcmp[1], -180.f, 180.f);
  double rot;
  if (xlat > 89.985f) rot = (double)(-strcmp[0]) * along + (double)xlong - 180.;
  else if (xlat < -89.985f) rot = (double)(-strcmp[0]) * along - (double)xlong;
  else rot = (double)(-strcmp[0]) * along;
  const double slong = sin(radpdg * rot), clong = cos(radpdg * rot);
  const double xpolg = slong * (double)strcmp[4] + clong * (double)strcmp[5];
  const double ypolg = clong * (double)strcmp[4] - slong * (double)strcmp[5];
  ug = (float)(ypolg * (double)ue + xpolg * (double)vn);
  vg = (float)(ypolg * (double)vn - xpolg * (double)ue);
}

// ---- pass 1: heights of the eta levels of one column, :198-231 ----------------------------------
FPB_HD inline void met_levels_column(const MetGrid &g, int ix, int jy) {
  const float cnst = M_R_AIR / M_GA;
  const float4 sf = g.SF1[m_o2(g, ix, jy)];
  const float ps = sf.x;
  float tvold = sf.y * (1.f + 0.378f * conv_ew(sf.z) / ps);
  float pold = ps, zold = 0.f;
  g.uvzlev[m_o3(g, ix, jy, 1)] = 0.f;
  for (int kz = 2; kz <= g.nuvz; kz++) {
    const float pint = g.akz[kz] + g.bkz[kz] * ps;
    const float2 tq = g.TQ[m_o3(g, ix, jy, kz)];
    const float tv = tq.x * (1.f + 0.608f * tq.y);
    float z;
    if (fabsf(tv - tvold) > 0.2f) z = zold + cnst * c_log(pold / pint) * (tv - tvold) / c_log(tv / tvold);
    else z = zold + cnst * c_log(pold / pint) * tv;
    g.uvzlev[m_o3(g, ix, jy, kz)] = z;
    tvold = tv;
    pold = pint;
    zold = z;
  }
}

// helpers on the column's stored uvzlev
struct MetCol {
  const MetGrid &g;
  int ix, jy;
  float ps;
  FPB_HD float uvz(int kz) const { return g.uvzlev[m_o3(g, ix, jy, kz)]; }
  FPB_HD float wz(int kz) const { // :236-240
    if (kz == 1) return 0.f;
    if (kz < g.nwz) return (uvz(kz + 1) + uvz(kz)) / 2.f;
    return wz(g.nwz - 1) + uvz(g.nuvz) - uvz(g.nuvz - 1);
  }
  FPB_HD float pk(int k) const { return g.akz[k] + g.bkz[k] * ps; } // aknew = akz, bknew = bkz (gridcheck_ecmwf.f90:532-533)
  FPB_HD float pinmconv(int kz) const { // :244-255
    if (kz == 1) return uvz(2) / (pk(2) - pk(1));
    if (kz < g.nz) return (uvz(kz + 1) - uvz(kz - 1)) / (pk(kz + 1) - pk(kz - 1));
    return (uvz(g.nz) - uvz(g.nz - 1)) / (pk(g.nz) - pk(g.nz - 1));
  }
  FPB_HD float rhoh(int kz) const { // :211,:222 (level 1 from the 2 m values)
    if (kz == 1) {
      const float4 sf = g.SF1[m_o2(g, ix, jy)];
      const float tvold = sf.y * (1.f + 0.378f * conv_ew(sf.z) / ps);
      return ps / (M_R_AIR * tvold);
    }
    const float2 tq = g.TQ[m_o3(g, ix, jy, kz)];
    const float tv = tq.x * (1.f + 0.608f * tq.y);
    return pk(kz) / (M_R_AIR * tv);
  }
};

// ---- pass 2: everything a column can do on its own (needs the neighbours' uvzlev) -----------------
FPB_HD inline void met_interp_column(const MetGrid &g, int ix, int jy) {
  const int nz = g.nz, nuvz = g.nuvz, nwz = g.nwz;
  MetCol c{g, ix, jy, g.SF1[m_o2(g, ix, jy)].x};
  const bool polar_n = g.nglobal && jy >= (int)g.switchnorthg - 2;
  const bool polar_s = g.sglobal && jy <= (int)g.switchsouthg + 3;
  const float ylat = g.ylat0 + (float)jy * g.dy, xlon = g.xlon0 + (float)ix * g.dx;
  const bool interior = ix >= 1 && ix <= g.nx - 2 && jy >= 1 && jy <= g.ny - 2;
  const float uvztop = c.uvz(nuvz);

  // top level first: the levels above the highest eta level copy it (:290-306)
  const float2 uvn = g.UV[m_o3(g, ix, jy, nuvz)];
  const float2 tqn = g.TQ[m_o3(g, ix, jy, nuvz)];
  const float pvn = g.PV ? g.PV[m_o3(g, ix, jy, nuvz)] : 0.f;
  const float rhon = c.rhoh(nuvz);

  int idx = 2, idxw = 2, idxs = 2;
  float rho_m2 = 0.f, rho_m1 = 0.f; // rho(iz-2), rho(iz-1)
  for (int iz = 1; iz <= nz; iz++) {
    const float hz = g.height[iz - 1];
    float uu, vv, tt, qv, pv, rho, ww;
    if (iz == 1) { // :260-272
      const float2 uv = g.UV[m_o3(g, ix, jy, 1)];
      const float2 tq = g.TQ[m_o3(g, ix, jy, 1)];
      uu = uv.x; vv = uv.y; tt = tq.x; qv = tq.y;
      pv = g.PV ? g.PV[m_o3(g, ix, jy, 1)] : 0.f;
      rho = c.rhoh(1);
    } else if (iz == nz) { // :274-287
      uu = uvn.x; vv = uvn.y; tt = tqn.x; qv = tqn.y; pv = pvn; rho = rhon;
    } else if (hz > uvztop) {
      uu = uvn.x; vv = uvn.y; tt = tqn.x; qv = tqn.y; pv = pvn; rho = rhon;
    } else { // :308-355
      for (int kz = idx; kz <= nuvz; kz++)
        if (idx <= kz && hz > c.uvz(kz - 1) && hz <= c.uvz(kz)) { idx = kz; break; }
      const int kz = idx;
      const float dz1 = hz - c.uvz(kz - 1), dz2 = c.uvz(kz) - hz, dz = dz1 + dz2;
      const float2 uva = g.UV[m_o3(g, ix, jy, kz - 1)], uvb = g.UV[m_o3(g, ix, jy, kz)];
      const float2 tqa = g.TQ[m_o3(g, ix, jy, kz - 1)], tqb = g.TQ[m_o3(g, ix, jy, kz)];
      uu = (uva.x * dz2 + uvb.x * dz1) / dz;
      vv = (uva.y * dz2 + uvb.y * dz1) / dz;
      tt = (tqa.x * dz2 + tqb.x * dz1) / dz;
      qv = (tqa.y * dz2 + tqb.y * dz1) / dz;
      pv = g.PV ? (g.PV[m_o3(g, ix, jy, kz - 1)] * dz2 + g.PV[m_o3(g, ix, jy, kz)] * dz1) / dz : 0.f;
      rho = (c.rhoh(kz - 1) * dz2 + c.rhoh(kz) * dz1) / dz;
    }
    // vertical wind, :362-389
    if (iz == 1) {
      ww = g.W[m_o3(g, ix, jy, 1)] * c.pinmconv(1);
    } else {
      for (int kz = idxw; kz <= nwz; kz++)
        if (idxw <= kz && hz > c.wz(kz - 1) && hz <= c.wz(kz)) { idxw = kz; break; }
      const int kz = idxw;
      const float dz1 = hz - c.wz(kz - 1), dz2 = c.wz(kz) - hz, dz = dz1 + dz2;
      ww = (g.W[m_o3(g, ix, jy, kz - 1)] * c.pinmconv(kz - 1) * dz2 + g.W[m_o3(g, ix, jy, kz)] * c.pinmconv(kz) * dz1) / dz;
    }
    // slope of the eta levels in the wind direction, :410-447
    if (interior && iz >= 2 && iz <= nz - 1) {
      for (int kz = idxs; kz <= nz; kz++)
        if (idxs <= kz && hz > c.uvz(kz - 1) && hz <= c.uvz(kz)) { idxs = kz; break; }
      const int kz = idxs;
      const float dz1 = hz - c.uvz(kz - 1), dz2 = c.uvz(kz) - hz, dz = dz1 + dz2;
      const float dzdx1 = (g.uvzlev[m_o3(g, ix + 1, jy, kz - 1)] - g.uvzlev[m_o3(g, ix - 1, jy, kz - 1)]) / 2.f;
      const float dzdx2 = (g.uvzlev[m_o3(g, ix + 1, jy, kz)] - g.uvzlev[m_o3(g, ix - 1, jy, kz)]) / 2.f;
      const float dzdx = (dzdx1 * dz2 + dzdx2 * dz1) / dz;
      const float dzdy1 = (g.uvzlev[m_o3(g, ix, jy + 1, kz - 1)] - g.uvzlev[m_o3(g, ix, jy - 1, kz - 1)]) / 2.f;
      const float dzdy2 = (g.uvzlev[m_o3(g, ix, jy + 1, kz)] - g.uvzlev[m_o3(g, ix, jy - 1, kz)]) / 2.f;
      const float dzdy = (dzdy1 * dz2 + dzdy2 * dz1) / dz;
      if (g.nest) ww = ww + (dzdx * uu * g.dxconst * g.xresol * g.cosf[jy] + dzdy * vv * g.dyconst * g.yresol);
      else ww = ww + (dzdx * uu * g.dxconst * g.cosf[jy] + dzdy * vv * g.dyconst);
    }
    const size_t o = m_o3(g, ix, jy, iz);
    g.A[o] = make_float4(uu, vv, ww, rho);
    g.T[o] = tt;
    if (g.Q) g.Q[o] = make_float2(pv, qv);
    // density gradient, :393-400 (one level behind)
    if (iz == 2) g.G[m_o3(g, ix, jy, 1)] = (rho - rho_m1) / (g.height[1] - g.height[0]);
    if (iz >= 3) g.G[m_o3(g, ix, jy, iz - 1)] = (rho - rho_m2) / (g.height[iz - 1] - g.height[iz - 3]);
    if (iz == nz) g.G[o] = g.G[m_o3(g, ix, jy, nz - 1)];
    rho_m2 = rho_m1;
    rho_m1 = rho;
    // polar stereographic winds, :455-466 / :530-542 (the pole rows themselves: met_pole_level)
    if (polar_n) {
      float up, vp;
      m_cc2gll(g.northpolemap, ylat, xlon, uu, vv, up, vp);
      g.P[o] = make_float2(up, vp);
    }
    if (polar_s) {
      float up, vp;
      m_cc2gll(g.southpolemap, ylat, xlon, uu, vv, up, vp);
      g.P[o] = make_float2(up, vp);
    }
  }

  // rain fields pass through; cloud / scavenging classes where rh > 80 %, :683-724
  const float4 s2 = g.SF2[m_o2(g, ix, jy)];
  if (g.R) {
    float4 r = g.R[m_o2(g, ix, jy)];
    r.x = s2.y; r.y = s2.z; r.z = s2.w;
    g.R[m_o2(g, ix, jy)] = r;
  }
  if (g.Cl && g.readclouds) {
    // cloud water from the input (:610-681).  `cloudh_min` is a scalar the reference carries from column to
    // column; a column reads it only after setting it when it holds any cloud water -- here it starts
    // at 0 in every column (a precipitating column WITHOUT cloud water gets no below-cloud class).
    const float lsp = s2.y, convp = s2.z;
    float ctwc = 0.f, cloudh_min = 0.f;
    // clwc on the height levels: interpolated like qv (:270-272,:283-286,:297-301,:339-344), ice added (:620-622)
    {
      int idxc = 2;
      for (int iz = 1; iz <= nz; iz++) {
        const float hz = g.height[iz - 1];
        float cw;
        auto at = [&](int kz) {
          const size_t o = m_o3(g, ix, jy, kz);
          return g.CLW[o];
        };
        auto ice = [&](int kz) { return g.CIW ? g.CIW[m_o3(g, ix, jy, kz)] : 0.f; };
        float ci;
        if (iz == 1) { cw = at(1); ci = ice(1); }
        else if (iz == nz || hz > uvztop) { cw = at(nuvz); ci = ice(nuvz); }
        else {
          for (int kz = idxc; kz <= nuvz; kz++)
            if (idxc <= kz && hz > c.uvz(kz - 1) && hz <= c.uvz(kz)) { idxc = kz; break; }
          const int kz = idxc;
          const float dz1 = hz - c.uvz(kz - 1), dz2 = c.uvz(kz) - hz, dz = dz1 + dz2;
          cw = (at(kz - 1) * dz2 + at(kz) * dz1) / dz;
          ci = g.CIW ? (ice(kz - 1) * dz2 + ice(kz) * dz1) / dz : 0.f;
        }
        if (g.CIW) cw = cw + ci;
        g.clw[m_o3(g, ix, jy, iz)] = cw; // (clwc for now)
      }
    }
    for (int kz = 1; kz <= nz - 1; kz++) {
      const size_t o = m_o3(g, ix, jy, kz);
      const float clwc = g.clw[o];
      float w = 0.f;
      if (clwc > 0.f) {
        w = (clwc * g.A[o].w) * (g.height[kz] - g.height[kz - 1]);
        ctwc = ctwc + w;
        cloudh_min = c_min(g.height[kz], g.height[kz - 1]);
      }
      g.clw[o] = w;
    }
    g.clw[m_o3(g, ix, jy, nz)] = 0.f;
    for (int kz = 1; kz <= nz; kz++) g.Cl[m_o3(g, ix, jy, kz)] = 0;
    if ((lsp > 0.01f) || (convp > 0.01f)) {
      for (int kz = nz; kz >= 2; kz--) {
        const size_t o = m_o3(g, ix, jy, kz);
        int cl = 0;
        if (g.clw[o] > 0.f) cl = (lsp >= convp) ? 3 : 2;
        else if (cloudh_min >= g.height[kz - 1]) cl = (lsp >= convp) ? 5 : 4;
        if (g.height[kz - 1] >= 19000.f) cl = 0;
        g.Cl[o] = (int8_t)cl;
      }
    }
    if (g.R) {
      float4 r = g.R[m_o2(g, ix, jy)];
      r.w = ctwc;
      g.R[m_o2(g, ix, jy)] = r;
    }
  }
  if (g.Cl && !g.readclouds) {
    const float lsp = s2.y, convp = s2.z;
    int rain_cloud_above = 0;
    for (int kz_inv = 1; kz_inv <= nz - 1; kz_inv++) {
      const int kz = nz - kz_inv + 1;
      const size_t o = m_o3(g, ix, jy, kz);
      const float rho = g.A[o].w, tt = g.T[o];
      const float qv = g.Q ? g.Q[o].y : 0.f;
      const float pressure = rho * M_R_AIR * tt;
      const float rh = qv / conv_qvsat(pressure, tt);
      int cl = 0;
      if (rh > 0.8f) {
        if ((lsp > 0.01f) || (convp > 0.01f)) {
          rain_cloud_above = 1;
          cl = (lsp >= convp) ? 3 : 2;
        } else {
          cl = 1;
        }
      } else if (rain_cloud_above == 1) {
        cl = (lsp >= convp) ? 5 : 4;
      }
      g.Cl[o] = (int8_t)cl;
    }
  }
}

// ---- pass 3: the pole rows of one level, :474-527 (north) and :544-607 (south) -------------------
FPB_HD inline void met_pole_level(const MetGrid &g, int iz) {
  const float pi = M_PI_F;
  for (int south = 0; south < 2; south++) {
    if (south ? !g.sglobal : !g.nglobal) continue;
    const int jyp = south ? 0 : g.ny - 1;     // the pole row
    const int jyn = south ? 1 : g.ny - 2;     // its equatorward neighbour
    const int ixc = g.nx / 2 - 1;
    const float4 a = g.A[m_o3(g, ixc, jyp, iz)];
    const float uc = a.x, vc = a.y;
    float xlon = g.xlon0 + (float)ixc * g.dx;
    float xlonr = xlon * pi / 180.f;
    const float ffpol = c_sqrt(uc * uc + vc * vc);
    float ddpol;
    if (vc < 0.f) ddpol = south ? m_atan(uc / vc) + xlonr : m_atan(uc / vc) - xlonr;
    else if (vc > 0.f) ddpol = south ? pi + m_atan(uc / vc) + xlonr : pi + m_atan(uc / vc) - xlonr;
    else ddpol = pi / 2.f - xlonr;
    if (ddpol < 0.f) ddpol = 2.0f * pi + ddpol;
    if (ddpol > 2.0f * pi) ddpol = ddpol - 2.0f * pi;
    xlon = 180.0f;
    xlonr = xlon * pi / 180.f;
    float uuaux, vvaux, up, vp;
    if (!south) {
      uuaux = -ffpol * m_sin(xlonr + ddpol);
      vvaux = -ffpol * m_cos(xlonr + ddpol);
      m_cc2gll(g.northpolemap, 90.0f, xlon, uuaux, vvaux, up, vp);
    } else {
      uuaux = +ffpol * m_sin(xlonr - ddpol);
      vvaux = -ffpol * m_cos(xlonr - ddpol);
      m_cc2gll(g.northpolemap, -90.0f, xlon, uuaux, vvaux, up, vp); // (northpolemap: as in the reference, :591)
    }
    float wdummy = 0.f;
    for (int ix = 0; ix < g.nx; ix++) wdummy = wdummy + g.A[m_o3(g, ix, jyn, iz)].z;
    wdummy = wdummy / (float)g.nx;
    for (int ix = 0; ix < g.nx; ix++) {
      const size_t o = m_o3(g, ix, jyp, iz);
      g.P[o] = make_float2(up, vp);
      g.A[o].z = wdummy;
    }
  }
}

// ---- calcpar, one column ---------------------------------------------------------------------------
// src/richardson.f90:54-196 (ECMWF branch); the column's levels are read where they lie
FPB_HD inline void met_richardson(const MetGrid &g, int ix, int jy, float psurf, float ust, float hf, float tt2,
                                  float td2, float &h, float &wst, float &hmixplus) {
  const float cnst = M_R_AIR / M_GA, ric = 0.25f, b = 100.f, bs = 8.5f;
  const int itmax = 3, nuvz = g.nuvz;
  float excess = 0.0f;
  int iter = 0;
  const float2 uv2 = g.UV[m_o3(g, ix, jy, 2)];
  for (;;) {
    iter = iter + 1;
    float pold = psurf;
    float tvold = tt2 * (1.f + 0.378f * conv_ew(td2) / psurf);
    float zold = 2.0f;
    const float zref = zold;
    float rhold = conv_ew(td2) / conv_ew(tt2);
    const float thetaref = tvold * c_pow(100000.f / pold, M_R_AIR / M_CPA) + excess;
    float thetaold = thetaref;
    float z = 0.f, theta = 0.f, rh = 0.f, pint, tv;
    int k;
    bool found = false;
    for (k = 2; k <= nuvz; k++) {
      pint = g.akz[k] + g.bkz[k] * psurf;
      const float2 tq = g.TQ[m_o3(g, ix, jy, k)];
      const float2 uv = g.UV[m_o3(g, ix, jy, k)];
      tv = tq.x * (1.f + 0.608f * tq.y);
      if (fabsf(tv - tvold) > 0.2f) z = zold + cnst * c_log(pold / pint) * (tv - tvold) / c_log(tv / tvold);
      else z = zold + cnst * c_log(pold / pint) * tv;
      theta = tv * c_pow(100000.f / pint, M_R_AIR / M_CPA);
      rh = tq.y / conv_qvsat(pint, tq.x);
      const float du = uv.x - uv2.x, dv = uv.y - uv2.y;
      const float ri = M_GA / thetaref * (theta - thetaref) * (z - zref) / c_max((du * du + dv * dv + b * (ust * ust)), 0.1f);
      if (ri > ric && thetaold < theta) { found = true; break; }
      tvold = tv;
      pold = pint;
      rhold = rh;
      thetaold = theta;
      zold = z;
    }
    if (!found) k = k - 1; // (ticket #139)
    float zl1 = zold, theta1 = thetaold, zl = 0.f, ul = 0.f, vl = 0.f, zl2 = 0.f, theta2 = 0.f;
    const float2 uva = g.UV[m_o3(g, ix, jy, k - 1)], uvb = g.UV[m_o3(g, ix, jy, k)];
    for (int i = 1; i <= 20; i++) {
      const float f = (float)i / 20.f;
      zl = zold + f * (z - zold);
      ul = uva.x + f * (uvb.x - uva.x);
      vl = uva.y + f * (uvb.y - uva.y);
      const float thetal = thetaold + f * (theta - thetaold);
      const float rhl = rhold + f * (rh - rhold);
      (void)rhl;
      const float du = ul - uv2.x, dv = vl - uv2.y;
      const float ril = M_GA / thetaref * (thetal - thetaref) * (zl - zref) / c_max((du * du + dv * dv + b * (ust * ust)), 0.1f);
      zl2 = zl;
      theta2 = thetal;
      if (ril > ric) break;
      zl1 = zl;
      theta1 = thetal;
    }
    h = zl;
    const float thetam = 0.5f * (theta1 + theta2);
    const float wspeed = c_sqrt(ul * ul + vl * vl);
    const float bvfsq = (M_GA / thetam) * (theta2 - theta1) / (zl2 - zl1);
    if (bvfsq <= 0.f) hmixplus = 9999.f;
    else hmixplus = wspeed / c_sqrt(bvfsq) * M_CONVKE;
    if (hf < 0.f) {
      wst = c_pow(-h * M_GA / thetaref * hf / M_CPA, 0.333f);
      excess = -bs * hf / M_CPA / wst;
      if (iter < itmax) continue;
    } else {
      wst = 0.f;
    }
    break;
  }
}

FPB_HD inline void met_calcpar_column(const MetGrid &g, int ix, int jy) {
  const float cnst = M_R_AIR / M_GA;
  const int nuvz = g.nuvz;
  const float4 sf = g.SF1[m_o2(g, ix, jy)];
  const float ps = sf.x, tt2 = sf.y, td2 = sf.z, sshf = sf.w;
  const float surfstr = g.SF2[m_o2(g, ix, jy)].x;
  // tropopause search floor, src/calcpar.f90:84-95
  const float ylat = g.ylat0 + (float)jy * g.dy;
  float altmin;
  if ((ylat >= -20.f) && (ylat <= 20.f)) altmin = 5000.f;
  else if ((ylat > 20.f) && (ylat < 40.f)) altmin = 2500.f + (40.f - ylat) * 125.f;
  else if ((ylat > -40.f) && (ylat < -20.f)) altmin = 2500.f + (40.f + ylat) * 125.f;
  else altmin = 2500.f;

  // 1) friction velocity, src/scalev.f90
  float ustar;
  {
    const float e = conv_ew(td2);
    const float tv = tt2 * (1.f + 0.378f * e / ps);
    const float rhoa = ps / (M_R_AIR * tv);
    ustar = c_sqrt(fabsf(surfstr) / rhoa);
  }
  if (ustar <= 1.e-8f) ustar = 1.e-8f;
  // 2) Obukhov length, src/obukhov.f90 (ECMWF branch)
  float ol;
  {
    const float e = conv_ew(td2);
    const float tv = tt2 * (1.f + 0.378f * e / ps);
    const float rhoa = ps / (M_R_AIR * tv);
    const float ak1 = (g.akm[1] + g.akm[2]) / 2.f, bk1 = (g.bkm[1] + g.bkm[2]) / 2.f;
    const float plev = ak1 + bk1 * ps;
    const float theta = g.TQ[m_o3(g, ix, jy, 2)].x * c_pow(100000.f / plev, M_R_AIR / M_CPA);
    const float thetastar = sshf / (rhoa * M_CPA * ustar);
    if (fabsf(thetastar) > 1.e-10f) ol = theta * (ustar * ustar) / (M_KARMAN * M_GA * thetastar);
    else ol = 9999.f;
    if (ol > 9999.f) ol = 9999.f;
    if (ol < -9999.f) ol = -9999.f;
  }
  const float oli = (ol != 0.f) ? 1.f / ol : 99999.f;
  // 3) mixing height and convective velocity scale
  float hmix, wstar, hmixplus;
  met_richardson(g, ix, jy, ps, ustar, sshf, tt2, td2, hmix, wstar, hmixplus);
  const float subsceff = (g.lsubgrid == 1) ? c_min(g.excessoro[m_o2(g, ix, jy)], hmixplus) : 0.0f;
  hmix = hmix + subsceff;
  hmix = c_max(M_HMIXMIN, hmix);
  hmix = c_min(M_HMIXMAX, hmix);
  g.S[m_o2(g, ix, jy)] = make_float4(hmix, ustar, wstar, oli);

  // thermal tropopause (Hoinka 1997), :188-258: the level heights are recomputed on the fly
  // (zlev(kz) of the reference = the same recurrence as uvzlev, from z = 0)
  // 2) first level at or above altmin
  int kzmin = nuvz + 1;
  for (int kz = 2; kz <= nuvz; kz++)
    if (g.uvzlev[m_o3(g, ix, jy, kz)] >= altmin) { kzmin = kz; break; }
  // (level 1: zlev(1) is never assigned in the reference -- 0 by -finit-local-zero -- and altmin > 0)
  // 3) first layer, >= 2 km deep, with a lapse rate below 2 K/km
  for (int kz = kzmin; kz <= nuvz; kz++) {
    const float zk = g.uvzlev[m_o3(g, ix, jy, kz)];
    bool done = false;
    for (int lz = kz + 1; lz <= nuvz; lz++) {
      const float zl = g.uvzlev[m_o3(g, ix, jy, lz)];
      if ((zl - zk) > 2000.f) {
        if (((g.TQ[m_o3(g, ix, jy, kz)].x - g.TQ[m_o3(g, ix, jy, lz)].x) / (zl - zk)) < 0.002f) {
          g.trop[m_o2(g, ix, jy)] = zk;
          done = true;
        }
        break;
      }
    }
    if (done) break;
  }
  (void)cnst;
}

// ---- calcpv, src/calcpv.f90 ---------------------------------------------------------------------
FPB_HD inline void met_theta_column(const MetGrid &g, int ix, int jy) { // :57-66 (theta = tth * ppmk)
  const float kappa = 0.286f;
  const float ps = g.SF1[m_o2(g, ix, jy)].x;
  for (int kl = 1; kl <= g.nuvz; kl++) {
    const float ppml = g.akz[kl] + g.bkz[kl] * ps;
    const float ppmk = c_pow(100000.f / ppml, kappa);
    g.theta[m_o3(g, ix, jy, kl)] = g.TQ[m_o3(g, ix, jy, kl)].x * ppmk;
  }
}

// the level pair of column (cx, cy) that brackets `theta`, searched alternately upwards and
// downwards from klpt, at most nlck pairs (:118-166); the wind component interpolated to theta
template <bool U>
FPB_HD inline bool met_isentropic(const MetGrid &g, int cx, int cy, float theta, int klpt, int nlck, float &val) {
  const float eps = 1.e-5f;
  int kup = klpt - 1, kdn = klpt, kch = 0;
  for (;;) {
    kup = kup + 1;
    if (kch >= nlck) return false;
    for (int pass = 0; pass < 2; pass++) {
      int k;
      if (pass == 0) {
        if (kup >= g.nuvz) continue;
        kch = kch + 1;
        k = kup;
      } else {
        kdn = kdn - 1;
        if (kdn < 1) break;
        kch = kch + 1;
        k = kdn;
      }
      const float thdn = g.theta[m_o3(g, cx, cy, k)], thup = g.theta[m_o3(g, cx, cy, k + 1)];
      if (((thdn >= theta) && (thup <= theta)) || ((thdn <= theta) && (thup >= theta))) {
        float dt1 = fabsf(theta - thdn), dt2 = fabsf(theta - thup), dt = dt1 + dt2;
        if (dt < eps) { dt1 = 0.5f; dt2 = 0.5f; dt = 1.0f; }
        const float2 a = g.UV[m_o3(g, cx, cy, k)], b = g.UV[m_o3(g, cx, cy, k + 1)];
        val = U ? (a.x * dt2 + b.x * dt1) / dt : (a.y * dt2 + b.y * dt1) / dt;
        return true;
      }
    }
  }
}

FPB_HD inline void met_calcpv_column(const MetGrid &g, int ix, int jy) {
  if (g.sglobal && jy == 0) return;           // pole rows: met_pv_pole_level
  if (g.nglobal && jy == g.ny - 1) return;
  const int nuvz = g.nuvz, nx = g.nx, nxmin1 = g.nx - 1, nymin1 = g.ny - 1, nlck = nuvz / 3;
  const float pi = M_PI_F, r_earth = 6.371e6f;
  const float phi = (g.ylat0 + (float)jy * g.dy) * pi / 180.f;
  const float f = 0.00014585f * m_sin(phi);
  const float tanphi = (float)tan((double)phi), cosphi = m_cos(phi);
  int jyvp = jy + 1, jyvm = jy - 1;
  if (jy == 0) jyvm = 0;
  if (jy == nymin1) jyvp = nymin1;
  int jumpy = 2;
  if (jy == 0 || jy == nymin1) jumpy = 1;
  if (g.sglobal && jy == 1) { jyvm = 1; jumpy = 1; }
  if (g.nglobal && jy == g.ny - 2) { jyvp = g.ny - 2; jumpy = 1; }
  int ixvp = ix + 1, ixvm = ix - 1, jumpx = 2, ivrp, ivrm;
  if (g.xglobal) {
    ivrp = ixvp; ivrm = ixvm;
    if (ixvm < 0) ivrm = ixvm + nxmin1;
    if (ixvp >= nx) ivrp = ixvp - nx + 1;
  } else {
    if (ix == 0) ixvm = 0;
    if (ix == nxmin1) ixvp = nxmin1;
    ivrp = ixvp; ivrm = ixvm;
    if (ix == 0 || ix == nxmin1) jumpx = 1;
  }
  const float ps = g.SF1[m_o2(g, ix, jy)].x;
  for (int kl = 1; kl <= nuvz; kl++) {
    const float theta = g.theta[m_o3(g, ix, jy, kl)];
    int klvrp = kl + 1, klvrm = kl - 1;
    if (klvrp > nuvz) klvrp = nuvz;
    if (klvrm < 1) klvrm = 1;
    const float thetap = g.theta[m_o3(g, ix, jy, klvrp)], thetam = g.theta[m_o3(g, ix, jy, klvrm)];
    const float dthetadp = (thetap - thetam) / ((g.akz[klvrp] + g.bkz[klvrp] * ps) - (g.akz[klvrm] + g.bkz[klvrm] * ps));
    const float2 uv0 = g.UV[m_o3(g, ix, jy, kl)];
    // v on the isentropic surface at the x neighbours, :105-176
    int jux = jumpx, ii = 0;
    float vx[2] = {0.f, 0.f};
    for (int i = ixvm; i <= ixvp; i += jumpx) {
      int ivr = i;
      if (g.xglobal) {
        if (i < 0) ivr = ivr + nxmin1;
        if (i >= nx) ivr = ivr - nx + 1;
      }
      float v;
      if (!met_isentropic<false>(g, ivr, jy, theta, kl, nlck, v)) { v = uv0.y; jux = jux - 1; }
      if (ii < 2) vx[ii] = v;
      ii++;
    }
    float dvdx;
    if (jux > 0) {
      dvdx = (vx[1] - vx[0]) / (float)jux / (g.dx * pi / 180.f);
    } else {
      dvdx = g.UV[m_o3(g, ivrp, jy, kl)].y - g.UV[m_o3(g, ivrm, jy, kl)].y;
      dvdx = dvdx / (float)jumpx / (g.dx * pi / 180.f);
    }
    // u at the y neighbours, :183-252
    int juy = jumpy, jj = 0;
    float uy[2] = {0.f, 0.f};
    for (int j = jyvm; j <= jyvp; j += jumpy) {
      float u;
      if (!met_isentropic<true>(g, ix, j, theta, kl, nlck, u)) { u = uv0.x; juy = juy - 1; }
      if (jj < 2) uy[jj] = u;
      jj++;
    }
    float dudy;
    if (juy > 0) {
      dudy = (uy[1] - uy[0]) / (float)juy / (g.dy * pi / 180.f);
    } else {
      dudy = g.UV[m_o3(g, ix, jyvp, kl)].x - g.UV[m_o3(g, ix, jyvm, kl)].x;
      dudy = dudy / (float)jumpy / (g.dy * pi / 180.f);
    }
    g.PV[m_o3(g, ix, jy, kl)] = dthetadp * (f + (dvdx / cosphi - dudy + uv0.x * tanphi) / r_earth) * (-1.e6f) * 9.81f;
  }
}

// the pole rows take the zonal mean of the neighbouring row, :277-313
FPB_HD inline void met_pv_pole_level(const MetGrid &g, int kl) {
  for (int south = 0; south < 2; south++) {
    if (south ? !g.sglobal : !g.nglobal) continue;
    const int jyp = south ? 0 : g.ny - 1, jyn = south ? 1 : g.ny - 2;
    float pvavr = 0.f;
    for (int ix = 0; ix < g.nx; ix++) pvavr = pvavr + g.PV[m_o3(g, ix, jyn, kl)];
    pvavr = pvavr / (float)g.nx;
    for (int ix = 0; ix < g.nx; ix++) g.PV[m_o3(g, ix, jyp, kl)] = pvavr;
  }
}

} // namespace fpbmet
