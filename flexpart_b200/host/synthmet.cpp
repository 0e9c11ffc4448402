// synthmet.cpp -- synthetic "ECMWF-shaped" meteorology.
//
// Produces, for one time level, the arrays verttransform_ecmwf / calcpar
// would leave in com_mod (src/com_mod.f90:355-371,410-427,451), in the same
// padded Fortran layout: smooth analytic fields on a global lat-lon grid,
// drhodz by the reference's finite differences
// (src/verttransform_ecmwf.f90:392-398), uupol/vvpol by cc2gll poleward of the
// switch latitudes and the pole-row treatment of :459-607, ww at the pole rows
// replaced by the zonal mean of the neighbouring row.  Invariants the hot
// path relies on are honoured: hmix in [100,4500] (src/calcpar.f90:165-166),
// ustar >= 1e-8, rho > 0, heights strictly increasing with height(1)=0.
#include <cmath>
#include <thread>
#include <vector>

#include "fpbh_internal.h"

namespace {
const double PI = 3.14159265358979323846;
inline size_t i3(const fpb_config &c, int ix, int jy, int k) {
  return (size_t)ix + (size_t)c.nxmax * ((size_t)jy + (size_t)c.nymax * (size_t)k);
}
inline size_t i2(const fpb_config &c, int ix, int jy) { return (size_t)ix + (size_t)c.nxmax * (size_t)jy; }

template <class F>
void parallel_levels(int nk, F f) {
  unsigned nt = std::thread::hardware_concurrency();
  if (nt == 0) nt = 4;
  if (nt > 32) nt = 32;
  if ((int)nt > nk) nt = nk;
  std::vector<std::thread> th;
  for (unsigned t = 0; t < nt; t++)
    th.emplace_back([=]() {
      for (int k = (int)t; k < nk; k += (int)nt) f(k);
    });
  for (auto &x : th) x.join();
}
} // namespace

extern "C" int fpbh_synth_heights(int32_t nz, float *height) {
  if (nz < 2 || !height) return fpbh_fail("fpbh_synth_heights: bad argument");
  // first layer 10 m, geometric stretching to ~80 km at 138 levels
  const double top = 80000.0;
  // solve 10*(r^(nz-1)-1)/(r-1) = top by bisection
  double lo = 1.0001, hi = 2.0;
  for (int it = 0; it < 200; it++) {
    double r = 0.5 * (lo + hi);
    double s = 10.0 * (std::pow(r, nz - 1) - 1.0) / (r - 1.0);
    if (s > top) hi = r; else lo = r;
  }
  const double r = 0.5 * (lo + hi);
  double z = 0.0, dz = 10.0;
  height[0] = 0.f;
  for (int k = 1; k < nz; k++) {
    z += dz;
    dz *= r;
    height[k] = (float)z;
  }
  return 0;
}

extern "C" int fpbh_synth_met(const fpb_config *cp, const float *height, int32_t time_s,
                              const fpb_met_ptrs *o) {
  if (!cp || !height || !o) return fpbh_fail("fpbh_synth_met: null argument");
  const fpb_config &c = *cp;
  if (!o->uu || !o->vv || !o->ww || !o->rho || !o->drhodz || !o->hmix || !o->ustar || !o->wstar ||
      !o->oli || !o->tropopause)
    return fpbh_fail("fpbh_synth_met: mandatory output array is null");
  const int nx = c.nx, ny = c.ny, nz = c.nz;
  const double wt = 2.0 * PI * (double)time_s / 86400.0; // diurnal phase
  float *uu = (float *)o->uu, *vv = (float *)o->vv, *ww = (float *)o->ww, *rho = (float *)o->rho;
  float *drhodz = (float *)o->drhodz, *tt = (float *)o->tt;
  float *uupol = (float *)o->uupol, *vvpol = (float *)o->vvpol;

  std::vector<double> sl(nx), cl(nx), s2l(nx), s3l(nx), lam(nx);
  for (int ix = 0; ix < nx; ix++) {
    lam[ix] = ((double)c.xlon0 + (double)c.dx * ix) * PI / 180.0;
    sl[ix] = std::sin(lam[ix]); cl[ix] = std::cos(lam[ix]);
    s2l[ix] = std::sin(2 * lam[ix]); s3l[ix] = std::sin(3 * lam[ix]);
  }
  parallel_levels(nz, [&](int k) {
    const double z = height[k];
    const double rz = 1.225 * std::exp(-z / 8000.0);
    const double tz = 288.0 - 6.5e-3 * std::fmin(z, 11000.0);
    const double wz = 0.02 * std::sin(PI * z / 20000.0);
    const double shear = 1.0 + z / 12000.0;
    for (int jy = 0; jy < ny; jy++) {
      const double phi = ((double)c.ylat0 + (double)c.dy * jy) * PI / 180.0;
      const double cp_ = std::cos(phi), c2p = std::cos(2 * phi);
      for (int ix = 0; ix < nx; ix++) {
        const size_t i = i3(c, ix, jy, k);
        uu[i] = (float)(15.0 * cp_ * shear + 5.0 * std::sin(3 * lam[ix] + 0.5 * wt) * c2p);
        vv[i] = (float)(5.0 * std::sin(2 * lam[ix] + 0.3 * wt) * cp_);
        ww[i] = (float)(wz * std::sin(lam[ix] + wt) * cp_);
        // small horizontal density structure so that rho gathers matter
        rho[i] = (float)(rz * (1.0 + 0.02 * sl[ix] * cp_));
        if (tt) tt[i] = (float)(tz + 3.0 * cl[ix] * cp_);
      }
    }
  });
  // drhodz, src/verttransform_ecmwf.f90:392-398
  parallel_levels(nz, [&](int k) {
    for (int jy = 0; jy < ny; jy++)
      for (int ix = 0; ix < nx; ix++) {
        const size_t i = i3(c, ix, jy, k);
        if (k == 0)
          drhodz[i] = (rho[i3(c, ix, jy, 1)] - rho[i3(c, ix, jy, 0)]) / (height[1] - height[0]);
        else if (k < nz - 1)
          drhodz[i] = (rho[i3(c, ix, jy, k + 1)] - rho[i3(c, ix, jy, k - 1)]) / (height[k + 1] - height[k - 1]);
      }
  });
  for (int jy = 0; jy < ny; jy++)
    for (int ix = 0; ix < nx; ix++) drhodz[i3(c, ix, jy, nz - 1)] = drhodz[i3(c, ix, jy, nz - 2)];

  // surface / PBL parameters
  float *hmix = (float *)o->hmix, *ustar = (float *)o->ustar, *wstar = (float *)o->wstar;
  float *oli = (float *)o->oli, *trop = (float *)o->tropopause;
  for (int jy = 0; jy < ny; jy++) {
    const double phi = ((double)c.ylat0 + (double)c.dy * jy) * PI / 180.0;
    const double cp_ = std::cos(phi), sp = std::sin(phi);
    for (int ix = 0; ix < nx; ix++) {
      const size_t i = i2(c, ix, jy);
      const double day = std::sin(lam[ix] + wt); // >0 "day side"
      double h = 100.0 + 1400.0 * (1.0 + day) * cp_ * cp_;
      hmix[i] = (float)std::fmin(4500.0, std::fmax(100.0, h));
      ustar[i] = (float)(0.05 + 0.75 * 0.5 * (1.0 + std::sin(2 * lam[ix] + 0.7 * wt) * cp_));
      wstar[i] = (float)(day > 0 ? 2.5 * day * cp_ : 0.0);
      // 1/L: unstable (negative) by day, stable by night, |1/L| in [1e-3,0.1]
      const double mag = 1e-3 + 0.099 * std::fabs(day) * cp_;
      oli[i] = (float)(day > 0 ? -mag : mag);
      trop[i] = (float)(16000.0 - 8000.0 * sp * sp);
    }
  }
  if (o->vdep) {
    float *vdep = (float *)o->vdep;
    for (int ks = 0; ks < c.nspec; ks++)
      for (int jy = 0; jy < ny; jy++) {
        const double phi = ((double)c.ylat0 + (double)c.dy * jy) * PI / 180.0;
        for (int ix = 0; ix < nx; ix++)
          vdep[i3(c, ix, jy, ks)] =
              (float)((1e-3 + 4.5e-3 * (1.0 + std::sin(lam[ix] + wt + ks) * std::cos(phi))) * (1.0 + 0.5 * ks));
      }
  }

  // precipitation / cloud fields for wet deposition: rain bands that move with
  // the diurnal phase; clouds(ix,jy,k) carries verttransform_ecmwf's classes
  // (src/verttransform_ecmwf.f90:640-700): 0 none, 1 cloud without precipitation,
  // 2/3 in-cloud (convective / large-scale dominated), 4/5 below-cloud, 6 above
  if (o->lsprec && o->convprec && o->tcc) {
    float *lsprec = (float *)o->lsprec, *convprec = (float *)o->convprec, *tcc = (float *)o->tcc;
    float *ctwc = (float *)o->ctwc;
    int8_t *clouds = (int8_t *)o->clouds;
    for (int jy = 0; jy < ny; jy++) {
      const double phi = ((double)c.ylat0 + (double)c.dy * jy) * PI / 180.0;
      const double cp_ = std::cos(phi);
      for (int ix = 0; ix < nx; ix++) {
        const size_t i = i2(c, ix, jy);
        const double ls = 6.0 * std::sin(2 * lam[ix] + 0.4 * wt) * cp_ - 1.0;
        const double cv = 30.0 * std::sin(3 * lam[ix] - 0.6 * wt) * cp_ * cp_ - 8.0;
        lsprec[i] = (float)std::fmax(0.0, ls);
        convprec[i] = (float)std::fmax(0.0, cv);
        tcc[i] = (float)(0.25 + 0.7 * std::fabs(std::sin(lam[ix] + 0.2 * wt)));
        if (ctwc) ctwc[i] = (float)(2.0e-4 * (1.0 + 0.8 * std::sin(2 * lam[ix] + wt)) * cp_ + 1.0e-5);
        if (clouds) {
          const bool rain = (lsprec[i] + convprec[i]) > 0.f;
          const double base = 600.0 + 900.0 * (0.5 + 0.5 * std::sin(lam[ix])); // cloud base, m
          const double top = base + 2500.0 + 3000.0 * cp_;
          for (int k = 0; k < nz; k++) {
            const double z = height[k];
            int8_t v;
            if (z > top) v = rain ? 6 : 0;
            else if (z >= base) v = rain ? (convprec[i] > lsprec[i] ? 2 : 3) : 1;
            else v = rain ? (convprec[i] > lsprec[i] ? 4 : 5) : 0;
            clouds[i3(c, ix, jy, k)] = v;
          }
        }
      }
    }
  }

  // polar-stereographic winds, src/verttransform_ecmwf.f90:459-607
  if (uupol && vvpol) {
    const float pi_f = 3.14159265f;
    if (c.nglobal) {
      parallel_levels(nz, [&](int k) {
        for (int jy = (int)c.switchnorthg - 2; jy <= c.nymin1; jy++) {
          const float ylat = c.ylat0 + (float)jy * c.dy;
          for (int ix = 0; ix <= c.nxmin1; ix++) {
            const float xlon = c.xlon0 + (float)ix * c.dx;
            const size_t i = i3(c, ix, jy, k);
            cmap::cc2gll(c.northpolemap, ylat, xlon, uu[i], vv[i], uupol[i], vvpol[i]);
          }
        }
        // pole row: wind of the central grid point rotated to 180 deg
        const size_t ic = i3(c, nx / 2 - 1, c.nymin1, k);
        float xlon = c.xlon0 + (float)(nx / 2 - 1) * c.dx;
        float xlonr = xlon * pi_f / 180.f;
        const float ffpol = std::sqrt(uu[ic] * uu[ic] + vv[ic] * vv[ic]);
        float ddpol;
        if (vv[ic] < 0.f) ddpol = std::atan(uu[ic] / vv[ic]) - xlonr;
        else if (vv[ic] > 0.f) ddpol = pi_f + std::atan(uu[ic] / vv[ic]) - xlonr;
        else ddpol = pi_f / 2 - xlonr;
        if (ddpol < 0.f) ddpol = 2.0f * pi_f + ddpol;
        if (ddpol > 2.0f * pi_f) ddpol = ddpol - 2.0f * pi_f;
        xlon = 180.0f;
        xlonr = xlon * pi_f / 180.f;
        const float uuaux = -ffpol * std::sin(xlonr + ddpol), vvaux = -ffpol * std::cos(xlonr + ddpol);
        float up, vp;
        cmap::cc2gll(c.northpolemap, 90.0f, xlon, uuaux, vvaux, up, vp);
        float wsum = 0.f;
        for (int ix = 0; ix <= c.nxmin1; ix++) {
          uupol[i3(c, ix, c.nymin1, k)] = up;
          vvpol[i3(c, ix, c.nymin1, k)] = vp;
          wsum += ww[i3(c, ix, ny - 2, k)];
        }
        wsum = wsum / (float)nx;
        for (int ix = 0; ix <= c.nxmin1; ix++) ww[i3(c, ix, c.nymin1, k)] = wsum;
      });
    }
    if (c.sglobal) {
      parallel_levels(nz, [&](int k) {
        for (int jy = 0; jy <= (int)c.switchsouthg + 3; jy++) {
          const float ylat = c.ylat0 + (float)jy * c.dy;
          for (int ix = 0; ix <= c.nxmin1; ix++) {
            const float xlon = c.xlon0 + (float)ix * c.dx;
            const size_t i = i3(c, ix, jy, k);
            cmap::cc2gll(c.southpolemap, ylat, xlon, uu[i], vv[i], uupol[i], vvpol[i]);
          }
        }
        const size_t ic = i3(c, nx / 2 - 1, 0, k);
        float xlon = c.xlon0 + (float)(nx / 2 - 1) * c.dx;
        float xlonr = xlon * pi_f / 180.f;
        const float ffpol = std::sqrt(uu[ic] * uu[ic] + vv[ic] * vv[ic]);
        float ddpol;
        if (vv[ic] < 0.f) ddpol = std::atan(uu[ic] / vv[ic]) + xlonr;
        else if (vv[ic] > 0.f) ddpol = pi_f + std::atan(uu[ic] / vv[ic]) + xlonr;
        else ddpol = pi_f / 2 - xlonr;
        if (ddpol < 0.f) ddpol = 2.0f * pi_f + ddpol;
        if (ddpol > 2.0f * pi_f) ddpol = ddpol - 2.0f * pi_f;
        xlon = 180.0f;
        xlonr = xlon * pi_f / 180.f;
        const float uuaux = +ffpol * std::sin(xlonr - ddpol), vvaux = -ffpol * std::cos(xlonr - ddpol);
        float up, vp;
        // (the reference passes northpolemap here, src/verttransform_ecmwf.f90:578)
        cmap::cc2gll(c.northpolemap, -90.0f, xlon, uuaux, vvaux, up, vp);
        float wsum = 0.f;
        for (int ix = 0; ix <= c.nxmin1; ix++) {
          uupol[i3(c, ix, 0, k)] = up;
          vvpol[i3(c, ix, 0, k)] = vp;
          wsum += ww[i3(c, ix, 1, k)];
        }
        wsum = wsum / (float)nx;
        for (int ix = 0; ix <= c.nxmin1; ix++) ww[i3(c, ix, 0, k)] = wsum;
      });
    }
  }
  return 0;
}

// One time level of nested input grid `nest` (1-based): the same analytic
// fields sampled on the nest's own, finer grid (what readwind_nests +
// verttransform_nests + calcpar_nests leave in uun.. of src/com_mod.f90:501-529),
// padded to (nxmaxn, nymaxn, nzmax).  No polar-stereographic twins on nests.
extern "C" int fpbh_synth_met_nest(const fpb_config *cp, const float *height, int32_t time_s,
                                   int32_t nest, const fpb_met_ptrs *o) {
  if (!cp || !height || !o) return fpbh_fail("fpbh_synth_met_nest: null argument");
  if (nest < 1 || nest > cp->numbnests) return fpbh_fail("fpbh_synth_met_nest: nest out of range");
  fpb_config cn = *cp;
  const int l = nest - 1;
  cn.nx = cp->nxn[l]; cn.ny = cp->nyn[l];
  cn.nxmax = cp->nxmaxn; cn.nymax = cp->nymaxn;
  cn.dx = cp->dx / cp->xresoln[l]; cn.dy = cp->dy / cp->yresoln[l];
  cn.xlon0 = cp->xlon0 + cp->xln[l] * cp->dx;
  cn.ylat0 = cp->ylat0 + cp->yln[l] * cp->dy;
  cn.nxmin1 = cn.nx - 1; cn.nymin1 = cn.ny - 1;
  cn.nglobal = cn.sglobal = cn.xglobal = 0;
  fpb_met_ptrs on = *o;
  on.uupol = nullptr; on.vvpol = nullptr;
  return fpbh_synth_met(&cn, height, time_s, &on);
}

// src/mpi_mod.f90:2940-2973 (set_fields_synthetic): homogeneous test fields
extern "C" int fpbh_homogeneous_met(const fpb_config *cp, float u, float v, float w,
                                    const fpb_met_ptrs *o) {
  if (!cp || !o) return fpbh_fail("fpbh_homogeneous_met: null argument");
  const fpb_config &c = *cp;
  const size_t n3 = (size_t)c.nxmax * c.nymax * c.nzmax, n2 = (size_t)c.nxmax * c.nymax;
  auto fill = [](const float *p, size_t n, float val) {
    if (!p) return;
    float *q = (float *)p;
    for (size_t i = 0; i < n; i++) q[i] = val;
  };
  fill(o->uu, n3, u); fill(o->vv, n3, v); fill(o->ww, n3, w);
  fill(o->uupol, n3, u); fill(o->vvpol, n3, v);
  fill(o->rho, n3, 1.3f); fill(o->drhodz, n3, 0.f); fill(o->tt, n3, 300.f);
  fill(o->hmix, n2, 10000.f); fill(o->tropopause, n2, 10000.f);
  fill(o->ustar, n2, 1.f); fill(o->wstar, n2, 1.f); fill(o->oli, n2, 0.01f);
  fill(o->vdep, n2 * c.maxspec, 0.f);
  return 0;
}

// ---- model-level ("raw") fields: what readwind_ecmwf would leave for calcpar + verttransform -------
// L137-like hybrid coefficients as FLEXPART holds them (src/gridcheck_ecmwf.f90:470-566): akm, bkm
// on the half levels (index 1 = surface), akz, bkz on the layer centres (akz(1) = 0, bkz(1) = 1: the
// surface itself); all arrays (1:nuvz), 0-based here.  nconvlev as derived at :553-566.
extern "C" int fpbh_synth_hybrid_levels(int32_t nuvz, float *akm, float *bkm, float *akz, float *bkz, int32_t *nconvlev) {
  if (nuvz < 4 || !akm || !bkm || !akz || !bkz) return fpbh_fail("fpbh_synth_hybrid_levels: bad argument");
  for (int k = 1; k <= nuvz; k++) {
    const double eta = std::pow((double)(nuvz - k) / (nuvz - 1.0), 1.35);
    const double b = std::pow(eta, 2.2);
    akm[k - 1] = (float)(101325.0 * (eta - b) + 1.0 * (1.0 - eta)); // ~1 Pa at the top
    bkm[k - 1] = (float)b;
  }
  akz[0] = 0.f; bkz[0] = 1.f;
  for (int k = 2; k <= nuvz; k++) {
    akz[k - 1] = 0.5f * (akm[k - 2] + akm[k - 1]);
    bkz[k - 1] = 0.5f * (bkm[k - 2] + bkm[k - 1]);
  }
  if (nconvlev) {
    int n = nuvz - 2;
    for (int i = 1; i <= nuvz - 2; i++)
      if (akz[i - 1] + bkz[i - 1] * 101325.f < 5000.f) { n = i; break; }
    *nconvlev = n < nuvz - 2 ? n : nuvz - 2;
  }
  return 0;
}

// One time level of synthetic model-level fields in the reference's padded layout: a warm, moist
// tropical belt (part of the columns convect, part of them rain), jets aloft, two mountain massifs
// (surface pressure down to ~720 hPa), day/night surface heat flux.  `time_s` moves the phases.
extern "C" int fpbh_synth_rawmet(const fpb_config *cp, int32_t nuvz, const float *akz, const float *bkz, int32_t time_s,
                                 const fpb_rawmet_ptrs *o) {
  if (!cp || !akz || !bkz || !o) return fpbh_fail("fpbh_synth_rawmet: null argument");
  if (!o->uuh || !o->vvh || !o->tth || !o->qvh || !o->wwh || !o->ps || !o->tt2 || !o->td2 || !o->sshf || !o->surfstr)
    return fpbh_fail("fpbh_synth_rawmet: mandatory output array is null");
  const fpb_config &c = *cp;
  const int nx = c.nx, ny = c.ny;
  const double wt = 2.0 * PI * (double)time_s / 86400.0;
  float *uuh = (float *)o->uuh, *vvh = (float *)o->vvh, *tth = (float *)o->tth, *qvh = (float *)o->qvh, *wwh = (float *)o->wwh;
  float *ps = (float *)o->ps, *tt2 = (float *)o->tt2, *td2 = (float *)o->td2, *sshf = (float *)o->sshf;
  float *surfstr = (float *)o->surfstr, *lsprec = (float *)o->lsprec, *convprec = (float *)o->convprec, *tcc = (float *)o->tcc;
  std::vector<double> t0((size_t)nx * ny), rh0((size_t)nx * ny);
  for (int jy = 0; jy < ny; jy++) {
    const double lat = (double)c.ylat0 + (double)c.dy * jy, cphi = std::cos(lat * PI / 180.0);
    for (int ix = 0; ix < nx; ix++) {
      const double lon = (double)c.xlon0 + (double)c.dx * ix, lam = lon * PI / 180.0;
      const size_t i = i2(c, ix, jy), q = (size_t)ix + (size_t)nx * jy;
      const double bump = 28000.0 * std::exp(-(std::pow((lon - 85.0) / 25.0, 2) + std::pow((lat - 33.0) / 12.0, 2))) +
                          20000.0 * std::exp(-(std::pow((lon + 70.0) / 12.0, 2) + std::pow((lat + 20.0) / 25.0, 2)));
      ps[i] = (float)(100000.0 + 1500.0 * std::sin(2 * lam + 0.2 * wt) * cphi - bump);
      t0[q] = 272.0 + 32.0 * cphi * cphi + 2.0 * std::sin(3 * lam + 0.5 * wt);
      rh0[q] = std::fmin(0.98, std::fmax(0.2, 0.55 + 0.42 * cphi * cphi + 0.1 * std::sin(5 * lam + wt)));
      tt2[i] = (float)(t0[q] + 0.5);
      td2[i] = (float)(t0[q] + 0.5 - 2.0 - 4.0 * (1.0 - rh0[q]));
      const double day = std::sin(lam + wt);
      sshf[i] = (float)(-160.0 * cphi * day + 15.0); // upward (negative) by day
      surfstr[i] = (float)(0.02 + 0.25 * std::fabs(std::sin(2 * lam + 0.7 * wt) * cphi));
      if (lsprec) lsprec[i] = (float)std::fmax(0.0, 2.5 * std::sin(3 * lam + 0.35 + 0.3 * wt) * std::sin(2 * lat * PI / 180.0 + 0.17) - 0.8);
      if (convprec) convprec[i] = (float)std::fmax(0.0, 3.0 * std::pow(cphi, 6) * std::sin(5 * lam + 0.4 * wt) - 0.6);
      if (tcc) tcc[i] = (float)std::fmin(1.0, std::fmax(0.0, 0.5 + 0.5 * std::sin(2 * lam) * std::cos(3 * lat * PI / 180.0)));
    }
  }
  parallel_levels(nuvz, [&](int k0) {
    const int k = k0 + 1; // Fortran level
    const double eta = k > 1 ? ((double)akz[k - 1] + (double)bkz[k - 1] * 101325.0) / 101325.0 : 1.0;
    for (int jy = 0; jy < ny; jy++) {
      const double lat = (double)c.ylat0 + (double)c.dy * jy, phi = lat * PI / 180.0, cphi = std::cos(phi);
      const double ttrop = 205.0 + 8.0 * cphi * cphi;
      for (int ix = 0; ix < nx; ix++) {
        const double lon = (double)c.xlon0 + (double)c.dx * ix, lam = lon * PI / 180.0;
        const size_t i = i3(c, ix, jy, k0), q = (size_t)ix + (size_t)nx * jy;
        const double psd = ps[i2(c, ix, jy)];
        const double p = k > 1 ? (double)akz[k - 1] + (double)bkz[k - 1] * psd : psd;
        const double t = std::fmax(t0[q] * std::pow(p / psd, 0.19), ttrop) + 0.2 * std::sin(0.37 * k + lam);
        const double es = 611.2 * std::exp(17.67 * (t - 273.15) / (t - 29.65));
        const double qs = 0.622 * es / std::fmax(p - 0.378 * es, 1.0);
        const double rh = std::fmin(1.0, std::fmax(0.02, rh0[q] * std::pow(p / psd, 1.2)));
        tth[i] = (float)t;
        qvh[i] = (float)std::fmin(0.03, std::fmax(1e-7, rh * qs));
        const double jet = 35.0 * std::pow(1.0 - eta, 0.7) * cphi * cphi;
        uuh[i] = (float)(4.0 + jet * (1.0 + 0.3 * std::sin(3 * lam + 0.5 * wt)) + 2.0 * std::sin(0.21 * k + lam));
        vvh[i] = (float)(6.0 * std::sin(2 * lam + 0.7 * eta + 0.3 * wt) * cphi + 1.5 * std::cos(0.17 * k));
        wwh[i] = (float)(1.6 * eta * (1.0 - eta) * std::sin(4 * lam + wt) * std::sin(3 * phi) + 0.02 * std::sin(0.5 * k + phi));
      }
    }
  });
  // level 1 carries the 10 m wind and repeats the lowest layer's temperature and humidity
  for (int jy = 0; jy < ny; jy++)
    for (int ix = 0; ix < nx; ix++) {
      const size_t a = i3(c, ix, jy, 0), b = i3(c, ix, jy, 1);
      uuh[a] = 0.6f * uuh[b]; vvh[a] = 0.6f * vvh[b];
      tth[a] = tth[b]; qvh[a] = qvh[b];
    }
  return 0;
}
