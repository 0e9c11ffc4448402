// metinit.cpp -- the part of verttransform_ecmwf that runs once per model run, on the host: the
// Cartesian height levels (src/verttransform_ecmwf.f90:111-181).  The first wind field is searched
// for a grid point with a surface pressure above 1000 hPa, and the heights of that column's eta
// levels become `height(1:nuvz)` for the whole run.  fpb_init needs them (fpb_config::height), so a
// caller that lets the device do calcpar + verttransform (fpb_calcpar_verttransform) derives them
// here from the raw field it is about to hand over.
#include <cmath>

#include "fpbh_internal.h"

namespace {
inline float f_log(float x) { return (float)std::log((double)x); }
inline float f_pow(float a, float b) { return (float)std::pow((double)a, (double)b); }
inline float f_powi(float x, int m) { // real**integer as gfortran's __powisf2
  unsigned n = (unsigned)(m < 0 ? -m : m);
  float y = (n % 2) ? x : 1.f;
  while (n >>= 1) {
    x = x * x;
    if (n % 2) y = y * x;
  }
  return m < 0 ? 1.f / y : y;
}
// src/ew.f90 (Goff-Gratch)
float ew(float x) {
  float y = 373.16f / x;
  float a = -7.90298f * (y - 1.f);
  a = a + (5.02808f * 0.43429f * f_log(y));
  float c = (1.f - (1.f / y)) * 11.344f;
  c = -1.f + f_pow(10.f, c);
  c = -1.3816f * c / f_powi(10.f, 7);
  float d = (1.f - y) * 3.49149f;
  d = -1.f + f_pow(10.f, d);
  d = 8.1328f * d / f_powi(10.f, 3);
  y = a + c + d;
  return 101324.6f * f_pow(10.f, y);
}
} // namespace

// akz, bkz: the Fortran arrays (1:nuvz), 0-based here.  ps, tt2, td2 (nxmax, nymax); tth, qvh
// (nxmax, nymax, nuvzmax).  height[nuvz] out; *ixm, *jym (may be NULL) the reference column.
extern "C" int fpbh_verttransform_heights(const fpb_config *cp, int32_t nuvz, const float *akz, const float *bkz,
                                          const float *ps, const float *tt2, const float *td2, const float *tth,
                                          const float *qvh, float *height, int32_t *ixm_out, int32_t *jym_out) {
  if (!cp || !akz || !bkz || !ps || !tt2 || !td2 || !tth || !qvh || !height)
    return fpbh_fail("fpbh_verttransform_heights: null argument");
  const fpb_config &c = *cp;
  const float r_air = 287.05f, ga = 9.81f, cnst = r_air / ga;
  int ixm = -1, jym = -1;
  for (int jy = 0; jy <= c.nymin1 && ixm < 0; jy++)
    for (int ix = 0; ix <= c.nxmin1; ix++)
      if (ps[(size_t)ix + (size_t)c.nxmax * jy] > 100000.f) { ixm = ix; jym = jy; break; }
  if (ixm < 0) return fpbh_fail("fpbh_verttransform_heights: no grid point with a surface pressure above 1000 hPa "
                                "(the reference then uses an undefined column, src/verttransform_ecmwf.f90:131-140)");
  const size_t o2 = (size_t)ixm + (size_t)c.nxmax * jym;
  float tvold = tt2[o2] * (1.f + 0.378f * ew(td2[o2]) / ps[o2]);
  float pold = ps[o2];
  height[0] = 0.f;
  for (int kz = 2; kz <= nuvz; kz++) {
    const size_t o3 = o2 + (size_t)c.nxmax * c.nymax * (kz - 1);
    const float pint = akz[kz - 1] + bkz[kz - 1] * ps[o2];
    const float tv = tth[o3] * (1.f + 0.608f * qvh[o3]);
    if (std::fabs(tv - tvold) > 0.2f)
      height[kz - 1] = height[kz - 2] + cnst * f_log(pold / pint) * (tv - tvold) / f_log(tv / tvold);
    else
      height[kz - 1] = height[kz - 2] + cnst * f_log(pold / pint) * tv;
    tvold = tv;
    pold = pint;
  }
  if (ixm_out) *ixm_out = ixm;
  if (jym_out) *jym_out = jym;
  return 0;
}
