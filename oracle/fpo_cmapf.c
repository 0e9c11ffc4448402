/*
 * fpo_cmapf.c -- oracle restatement of the conformal-map routines FLEXPART
 * uses poleward of +-75 deg (test infrastructure).  Follows
 * src/cmapf_mod.f90: cspanf :494-524, cnllxy :310-365, cnxyll :367-425,
 * cll2xy :295-308, cxy2ll :526-543, cgszll :190-238, cc2gll :23-52,
 * stlmbr :784-814, stcm2p :603-640.  The mixed real/real(dp) typing of the
 * reference is kept: which operations run in double is part of the result.
 */
#include "fpo.h"
#include "fpo_math.h"

static const float rearth = 6371.2f, almst1 = .9999999f;
/* `real,parameter :: pi=3.14159265358979`: a single-precision constant */
static const float pi_c = 3.14159265358979f;
#define radpdg (pi_c / 180.f)
#define dgprad (180.f / pi_c)

/* src/cmapf_mod.f90:494-524 */
static float cspanf(float value, float begin, float end) {
  float first = fpo_minf(begin, end);
  float last = fpo_maxf(begin, end);
  float val = fmodf(value - first, last - first);
  if (val <= 0.f) return val + last;
  return val + first;
}

/* src/cmapf_mod.f90:310-365 */
static void cnllxy(const float *strcmp, float xlat, float xlong, float *xi,
                   float *eta) {
  float gdlong, sndgam, csdgam, rhog1;
  double gamma = strcmp[0];
  double dlat = xlat;
  double dlong = cspanf(xlong - strcmp[1], -180.f, 180.f);
  dlong = dlong * radpdg;
  gdlong = (float)(gamma * dlong);
  if (fabsf(gdlong) < .01f) {
    gdlong = gdlong * gdlong;
    sndgam = (float)(dlong * (1.f - 1.f / 6.f * gdlong *
                                        (1.f - 1.f / 20.f * gdlong *
                                                   (1.f - 1.f / 42.f * gdlong))));
    csdgam = (float)(dlong * dlong * .5f *
                     (1.f - 1.f / 12.f * gdlong *
                                (1.f - 1.f / 30.f * gdlong *
                                           (1.f - 1.f / 56.f * gdlong))));
  } else {
    sndgam = (float)(fpo_sinf(gdlong) / gamma);
    csdgam = (float)((1.f - fpo_cosf(gdlong)) / gamma / gamma);
  }
  double slat = sin(radpdg * dlat);
  if ((slat >= almst1) || (slat <= -almst1)) {
    *eta = 1.f / strcmp[0];
    *xi = 0.f;
    return;
  }
  double mercy = .5f * log((1.f + slat) / (1.f - slat));
  double gmercy = gamma * mercy;
  if (fabs(gmercy) < .001f) {
    rhog1 = (float)(mercy * (1.f - .5f * gmercy *
                                       (1.f - 1.f / 3.f * gmercy *
                                                  (1.f - 1.f / 4.f * gmercy))));
  } else {
    rhog1 = (float)((1.f - exp(-gmercy)) / gamma);
  }
  *eta = (float)(rhog1 + (1.f - gamma * rhog1) * gamma * csdgam);
  *xi = (float)((1.f - gamma * rhog1) * sndgam);
}

/* src/cmapf_mod.f90:367-425 (xi, eta are real(dp) here) */
static void cnxyll(const float *strcmp, double xi, double eta, float *xlat,
                   float *xlong) {
  double gamma = strcmp[0], temp, arg1, arg2, ymerc, along, gxi, cgeta;
  arg2 = 2.f * eta - gamma * (xi * xi + eta * eta);
  arg1 = gamma * arg2;
  if (fabs(arg1) < .01f) {
    temp = (arg1 / (2.f - arg1)) * (arg1 / (2.f - arg1));
    ymerc = arg2 / (2.f - arg1) *
            (1.f + temp * (1.f / 3.f + temp * (1.f / 5.f + temp * (1.f / 7.f))));
  } else {
    ymerc = -log(1.f - arg1) / 2.f / gamma;
  }
  temp = exp(-fabs(ymerc));
  *xlat = (float)copysign(atan2((1.f - temp) * (1.f + temp), 2.f * temp), ymerc);
  gxi = gamma * xi;
  cgeta = 1.f - gamma * eta;
  if (fabs(gxi) < .01f * cgeta) {
    temp = (gxi / cgeta) * (gxi / cgeta);
    along = xi / cgeta *
            (1.f - temp * (1.f / 3.f - temp * (1.f / 5.f - temp * (1.f / 7.f))));
  } else {
    along = atan2(gxi, cgeta) / gamma;
  }
  *xlong = (float)(strcmp[1] + dgprad * along);
  *xlat = *xlat * dgprad;
}

/* src/cmapf_mod.f90:295-308 */
void fpo_cll2xy(const float *strcmp, float xlat, float xlong, float *x,
                float *y) {
  float xi, eta;
  cnllxy(strcmp, xlat, xlong, &xi, &eta);
  *x = strcmp[2] + rearth / strcmp[6] * (xi * strcmp[4] + eta * strcmp[5]);
  *y = strcmp[3] + rearth / strcmp[6] * (eta * strcmp[4] - xi * strcmp[5]);
}

/* src/cmapf_mod.f90:526-543 */
void fpo_cxy2ll(const float *strcmp, float x, float y, float *xlat,
                float *xlong) {
  double xi0 = (x - strcmp[2]) * strcmp[6] / rearth;
  double eta0 = (y - strcmp[3]) * strcmp[6] / rearth;
  double xi = xi0 * strcmp[4] - eta0 * strcmp[5];
  double eta = eta0 * strcmp[4] + xi0 * strcmp[5];
  cnxyll(strcmp, xi, eta, xlat, xlong);
  *xlong = cspanf(*xlong, -180.f, 180.f);
}

/* src/cmapf_mod.f90:190-238 */
float fpo_cgszll(const float *strcmp, float xlat, float xlong) {
  (void)xlong;
  double slat, ymerc, efact;
  if (xlat > 89.985f) {
    if (strcmp[0] > 0.9999f) return 2.f * strcmp[6];
    efact = fpo_cosf(radpdg * xlat);
    if (efact <= 0.) return 0.f;
    ymerc = -log(efact / (1.f + fpo_sinf(radpdg * xlat)));
  } else if (xlat < -89.985f) {
    if (strcmp[0] < -0.9999f) return 2.f * strcmp[6];
    efact = fpo_cosf(radpdg * xlat);
    if (efact <= 0.) return 0.f;
    ymerc = log(efact / (1.f - fpo_sinf(radpdg * xlat)));
  } else {
    slat = fpo_sinf(radpdg * xlat);
    ymerc = log((1.f + slat) / (1.f - slat)) / 2.f;
  }
  return (float)(strcmp[6] * fpo_cosf(radpdg * xlat) * exp(strcmp[0] * ymerc));
}

/* src/cmapf_mod.f90:23-52 */
void fpo_cc2gll(const float *strcmp, float xlat, float xlong, float ue,
                float vn, float *ug, float *vg) {
  double along = cspanf(xlong - strcmp[1], -180.f, 180.f), rot;
  if (xlat > 89.985f)
    rot = -strcmp[0] * along + xlong - 180.f;
  else if (xlat < -89.985f)
    rot = -strcmp[0] * along - xlong;
  else
    rot = -strcmp[0] * along;
  double slong = sin(radpdg * rot);
  double clong = cos(radpdg * rot);
  double xpolg = slong * strcmp[4] + clong * strcmp[5];
  double ypolg = clong * strcmp[4] - slong * strcmp[5];
  *ug = (float)(ypolg * ue + xpolg * vn);
  *vg = (float)(ypolg * vn - xpolg * ue);
}

/* src/cmapf_mod.f90:784-814 */
void fpo_stlmbr(float *strcmp, float tnglat, float xlong) {
  float eta, xi;
  strcmp[0] = fpo_sinf(radpdg * tnglat);
  strcmp[1] = cspanf(xlong, -180.f, +180.f);
  strcmp[2] = 0.f;
  strcmp[3] = 0.f;
  strcmp[4] = 1.f;
  strcmp[5] = 0.f;
  strcmp[6] = rearth;
  cnllxy(strcmp, 89.f, xlong, &xi, &eta);
  strcmp[7] = 2.f * eta - strcmp[0] * eta * eta;
  cnllxy(strcmp, -89.f, xlong, &xi, &eta);
  strcmp[8] = 2.f * eta - strcmp[0] * eta * eta;
}

/* src/cmapf_mod.f90:603-640 */
void fpo_stcm2p(float *strcmp, float x1, float y1, float xlat1, float xlong1,
                float x2, float y2, float xlat2, float xlong2) {
  float x1a, y1a, x2a, y2a, den, dena;
  for (int k = 3; k <= 6; k++) strcmp[k - 1] = 0.f;
  strcmp[4] = 1.f;
  strcmp[6] = 1.f;
  fpo_cll2xy(strcmp, xlat1, xlong1, &x1a, &y1a);
  fpo_cll2xy(strcmp, xlat2, xlong2, &x2a, &y2a);
  den = fpo_sqrtf((x1 - x2) * (x1 - x2) + (y1 - y2) * (y1 - y2));
  dena = fpo_sqrtf((x1a - x2a) * (x1a - x2a) + (y1a - y2a) * (y1a - y2a));
  strcmp[4] = ((x1a - x2a) * (x1 - x2) + (y1a - y2a) * (y1 - y2)) / den / dena;
  strcmp[5] = ((y1a - y2a) * (x1 - x2) - (x1a - x2a) * (y1 - y2)) / den / dena;
  strcmp[6] = strcmp[6] * dena / den;
  fpo_cll2xy(strcmp, xlat1, xlong1, &x1a, &y1a);
  strcmp[2] = strcmp[2] + x1 - x1a;
  strcmp[3] = strcmp[3] + y1 - y1a;
}
