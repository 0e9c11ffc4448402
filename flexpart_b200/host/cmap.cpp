// cmap.cpp -- Taylor's conformal-map (CMAPF) routines FLEXPART uses to build
// and apply its polar-stereographic maps (src/cmapf_mod.f90: stlmbr :784,
// stcm2p :603, cll2xy :295, cxy2ll :526, cnllxy :310, cnxyll :367, cc2gll :23,
// cspanf :494).  Host-side twin of the device code in csrc/fpb_kernels.cu;
// the map arrays it produces are run constants handed to the engine.
// Single/double typing follows the reference's declarations.
#include <cmath>

#include "fpbh_internal.h"

namespace cmap {

static const float REARTH = 6371.2f, ALMST1 = .9999999f;
static const float PI_C = 3.14159265358979f;
static const float RADPDG = PI_C / 180.f, DGPRAD = 180.f / PI_C;

// float functions evaluated in double and rounded once (library convention)
static inline float sin_f(float x) { return (float)std::sin((double)x); }
static inline float cos_f(float x) { return (float)std::cos((double)x); }

float cspanf(float value, float begin, float end) {
  const float first = std::fmin(begin, end), last = std::fmax(begin, end);
  const float val = std::fmod(value - first, last - first);
  return (val <= 0.f) ? val + last : val + first;
}

void cnllxy(const float *m, float xlat, float xlong, float &xi, float &eta) {
  const double gamma = m[0], dlat = xlat;
  double dlong = cspanf(xlong - m[1], -180.f, 180.f);
  dlong = dlong * RADPDG;
  float gdlong = (float)(gamma * dlong), sndgam, csdgam, rhog1;
  if (std::fabs(gdlong) < .01f) {
    gdlong = gdlong * gdlong;
    sndgam = (float)(dlong * (1.f - 1.f / 6.f * gdlong * (1.f - 1.f / 20.f * gdlong * (1.f - 1.f / 42.f * gdlong))));
    csdgam = (float)(dlong * dlong * .5f * (1.f - 1.f / 12.f * gdlong * (1.f - 1.f / 30.f * gdlong * (1.f - 1.f / 56.f * gdlong))));
  } else {
    sndgam = (float)(sin_f(gdlong) / gamma);
    csdgam = (float)((1.f - cos_f(gdlong)) / gamma / gamma);
  }
  const double slat = std::sin(RADPDG * dlat);
  if (slat >= ALMST1 || slat <= -ALMST1) {
    eta = 1.f / m[0];
    xi = 0.f;
    return;
  }
  const double mercy = .5f * std::log((1.f + slat) / (1.f - slat));
  const double gmercy = gamma * mercy;
  if (std::fabs(gmercy) < .001f)
    rhog1 = (float)(mercy * (1.f - .5f * gmercy * (1.f - 1.f / 3.f * gmercy * (1.f - 1.f / 4.f * gmercy))));
  else
    rhog1 = (float)((1.f - std::exp(-gmercy)) / gamma);
  eta = (float)(rhog1 + (1.f - gamma * rhog1) * gamma * csdgam);
  xi = (float)((1.f - gamma * rhog1) * sndgam);
}

static void cnxyll(const float *m, double xi, double eta, float &xlat, float &xlong) {
  const double gamma = m[0];
  double temp, ymerc, along;
  const double arg2 = 2.f * eta - gamma * (xi * xi + eta * eta), arg1 = gamma * arg2;
  if (std::fabs(arg1) < .01f) {
    temp = (arg1 / (2.f - arg1)) * (arg1 / (2.f - arg1));
    ymerc = arg2 / (2.f - arg1) * (1.f + temp * (1.f / 3.f + temp * (1.f / 5.f + temp * (1.f / 7.f))));
  } else {
    ymerc = -std::log(1.f - arg1) / 2.f / gamma;
  }
  temp = std::exp(-std::fabs(ymerc));
  xlat = (float)std::copysign(std::atan2((1.f - temp) * (1.f + temp), 2.f * temp), ymerc);
  const double gxi = gamma * xi, cgeta = 1.f - gamma * eta;
  if (std::fabs(gxi) < .01f * cgeta) {
    temp = (gxi / cgeta) * (gxi / cgeta);
    along = xi / cgeta * (1.f - temp * (1.f / 3.f - temp * (1.f / 5.f - temp * (1.f / 7.f))));
  } else {
    along = std::atan2(gxi, cgeta) / gamma;
  }
  xlong = (float)(m[1] + DGPRAD * along);
  xlat = xlat * DGPRAD;
}

void cll2xy(const float *m, float xlat, float xlong, float &x, float &y) {
  float xi, eta;
  cnllxy(m, xlat, xlong, xi, eta);
  x = m[2] + REARTH / m[6] * (xi * m[4] + eta * m[5]);
  y = m[3] + REARTH / m[6] * (eta * m[4] - xi * m[5]);
}

void cxy2ll(const float *m, float x, float y, float &xlat, float &xlong) {
  const double xi0 = (x - m[2]) * m[6] / REARTH, eta0 = (y - m[3]) * m[6] / REARTH;
  const double xi = xi0 * m[4] - eta0 * m[5], eta = eta0 * m[4] + xi0 * m[5];
  cnxyll(m, xi, eta, xlat, xlong);
  xlong = cspanf(xlong, -180.f, 180.f);
}

void cc2gll(const float *m, float xlat, float xlong, float ue, float vn, float &ug, float &vg) {
  const double along = cspanf(xlong - m[1], -180.f, 180.f);
  double rot;
  if (xlat > 89.985f) rot = -m[0] * along + xlong - 180.f;
  else if (xlat < -89.985f) rot = -m[0] * along - xlong;
  else rot = -m[0] * along;
  const double slong = std::sin(RADPDG * rot), clong = std::cos(RADPDG * rot);
  const double xpolg = slong * m[4] + clong * m[5], ypolg = clong * m[4] - slong * m[5];
  ug = (float)(ypolg * ue + xpolg * vn);
  vg = (float)(ypolg * vn - xpolg * ue);
}

void stlmbr(float *m, float tnglat, float xlong) {
  float eta, xi;
  m[0] = sin_f(RADPDG * tnglat);
  m[1] = cspanf(xlong, -180.f, +180.f);
  m[2] = 0.f; m[3] = 0.f; m[4] = 1.f; m[5] = 0.f;
  m[6] = REARTH;
  cnllxy(m, 89.f, xlong, xi, eta);
  m[7] = 2.f * eta - m[0] * eta * eta;
  cnllxy(m, -89.f, xlong, xi, eta);
  m[8] = 2.f * eta - m[0] * eta * eta;
}

void stcm2p(float *m, float x1, float y1, float xlat1, float xlong1, float x2, float y2,
            float xlat2, float xlong2) {
  float x1a, y1a, x2a, y2a;
  m[2] = m[3] = m[5] = 0.f;
  m[4] = 1.f;
  m[6] = 1.f;
  cll2xy(m, xlat1, xlong1, x1a, y1a);
  cll2xy(m, xlat2, xlong2, x2a, y2a);
  const float den = std::sqrt((x1 - x2) * (x1 - x2) + (y1 - y2) * (y1 - y2));
  const float dena = std::sqrt((x1a - x2a) * (x1a - x2a) + (y1a - y2a) * (y1a - y2a));
  m[4] = ((x1a - x2a) * (x1 - x2) + (y1a - y2a) * (y1 - y2)) / den / dena;
  m[5] = ((y1a - y2a) * (x1 - x2) - (x1a - x2a) * (y1 - y2)) / den / dena;
  m[6] = m[6] * dena / den;
  cll2xy(m, xlat1, xlong1, x1a, y1a);
  m[2] = m[2] + x1 - x1a;
  m[3] = m[3] + y1 - y1a;
}

} // namespace cmap

extern "C" {
void fpbh_stlmbr(float *m, float tnglat, float xlong) { cmap::stlmbr(m, tnglat, xlong); }
void fpbh_stcm2p(float *m, float x1, float y1, float xlat1, float xlong1, float x2, float y2,
                 float xlat2, float xlong2) {
  cmap::stcm2p(m, x1, y1, xlat1, xlong1, x2, y2, xlat2, xlong2);
}
void fpbh_cc2gll(const float *m, float xlat, float xlong, float ue, float vn, float *ug, float *vg) {
  cmap::cc2gll(m, xlat, xlong, ue, vn, *ug, *vg);
}
void fpbh_cll2xy(const float *m, float xlat, float xlong, float *x, float *y) {
  cmap::cll2xy(m, xlat, xlong, *x, *y);
}
void fpbh_cxy2ll(const float *m, float x, float y, float *xlat, float *xlong) {
  cmap::cxy2ll(m, x, y, *xlat, *xlong);
}
}
