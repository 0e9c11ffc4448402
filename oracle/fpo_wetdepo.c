/*
 * fpo_wetdepo.c -- oracle restatement of the wet-deposition step
 * (test infrastructure):
 *   wetdepo             src/wetdepo.f90:55-147
 *   get_wetscav         src/get_wetscav.f90:78-314
 *   interpol_rain       src/interpol_rain.f90:77-127 (+ interpol_rain_nests.f90)
 *   wetdepokernel       src/wetdepokernel.f90:38-108
 *   wetdepokernel_nest  src/wetdepokernel_nest.f90:38-105
 *
 * Typing follows the Fortran declarations: default real = float; ix=int(xtra1)
 * truncates the double position; interpol_rain receives real(xtra1).
 * Integer powers x**(-n) are evaluated the way libgcc's __powisf2 (what
 * gfortran emits for real**integer) does: repeated squaring, then 1/y.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "fpo.h"
#include "fpo_math.h"

#define smallnum 1.17549435e-38f /* tiny(0.0) */
static const float incloud_ratio = 6.2f; /* src/par_mod.f90:82 */
static const float r_air = 287.05f;      /* src/par_mod.f90:59 */

static float powi(float x, int m) {
  unsigned n = (unsigned)(m < 0 ? -m : m);
  float y = (n % 2) ? x : 1.f;
  while (n >>= 1) {
    x = x * x;
    if (n % 2) y = y * x;
  }
  return m < 0 ? 1.f / y : y;
}

static inline float log10_f(float x) {
#ifdef FPO_LIBM_FLOAT
  return log10f(x);
#else
  return (float)log10((double)x);
#endif
}

/* src/interpol_rain.f90:77-127: bilinear, one time level (iwftouse), level 1 */
static void interpol_rain(const fpb_met_ptrs *M, size_t ldx, int nx, int ny, float xt,
                          float yt, float *yint1, float *yint2, float *yint3) {
  if (xt >= (float)(nx - 1)) xt = (float)(nx - 1) - 0.00001f;
  if (yt >= (float)(ny - 1)) yt = (float)(ny - 1) - 0.00001f;
  const int ix = fpo_int_f(xt), jy = fpo_int_f(yt), ixp = ix + 1, jyp = jy + 1;
  const float ddx = xt - (float)ix, ddy = yt - (float)jy;
  const float rddx = 1.f - ddx, rddy = 1.f - ddy;
  const float p1 = rddx * rddy, p2 = ddx * rddy, p3 = rddx * ddy, p4 = ddx * ddy;
  const size_t a = (size_t)ix + ldx * (size_t)jy, b = (size_t)ixp + ldx * (size_t)jy,
               c = (size_t)ix + ldx * (size_t)jyp, d = (size_t)ixp + ldx * (size_t)jyp;
  *yint1 = p1 * M->lsprec[a] + p2 * M->lsprec[b] + p3 * M->lsprec[c] + p4 * M->lsprec[d];
  *yint2 = p1 * M->convprec[a] + p2 * M->convprec[b] + p3 * M->convprec[c] + p4 * M->convprec[d];
  *yint3 = p1 * M->tcc[a] + p2 * M->tcc[b] + p3 * M->tcc[c] + p4 * M->tcc[d];
}

static float get_wetscav(fpo_state *S, int itime, int ltsample, int jpart, int ks, float *grfraction1);
float fpo_get_wetscav(fpo_state *S, int itime, int ltsample, int jpart, int ks, float *grfraction1) {
  return get_wetscav(S, itime, ltsample, jpart, ks, grfraction1);
}

/* src/get_wetscav.f90:78-314.  Returns wetscav; grfraction1 is only written
 * when scavenging is evaluated (as in the reference). */
static float get_wetscav(fpo_state *S, int itime, int ltsample, int jpart, int ks,
                         float *grfraction1) {
  const fpb_config *c = &S->c;
  static const float lfr[5] = {0.5f, 0.65f, 0.8f, 0.9f, 0.95f};
  static const float cfr[5] = {0.4f, 0.55f, 0.7f, 0.8f, 0.9f};
  static const float bclr[6] = {274.35758f, 332839.59273f, 226656.57259f, 58005.91340f, 6588.38582f, 0.244984f};
  static const float bcls[6] = {22.7f, 0.0f, 0.0f, 1321.0f, 381.0f, 0.0f};
  float wetscav = 0.f;
  const double xd = S->xtra1[jpart], yd = S->ytra1[jpart];
  int ngrid = 0, ix, jy, hz = 1, i, j, n;
  float xtn = 0.f, ytn = 0.f, lsp, convp, cc, prec1, act_temp, cl;
  int readclouds_this_nest = 0;

  for (j = c->numbnests; j >= 1; j--) /* :82-90, no eps margin here */
    if ((xd > c->xln[j - 1]) && (xd < c->xrn[j - 1]) && (yd > c->yln[j - 1]) && (yd < c->yrn[j - 1])) {
      ngrid = j;
      break;
    }
  if (ngrid > 0) {
    xtn = (float)((xd - c->xln[ngrid - 1]) * c->xresoln[ngrid - 1]);
    ytn = (float)((yd - c->yln[ngrid - 1]) * c->yresoln[ngrid - 1]);
    ix = fpo_int_f(xtn);
    jy = fpo_int_f(ytn);
    if (c->readclouds_nest[ngrid - 1]) readclouds_this_nest = 1;
  } else {
    ix = fpo_int_d(xd);
    jy = fpo_int_d(yd);
  }

  /* interpolated time refers to itime-0.5*ltsample, :113-117 */
  const int interp_time = fpo_nint_f((float)itime - 0.5f * (float)ltsample);
  n = S->memind[2];
  if (abs(S->memtime[1] - interp_time) < abs(S->memtime[2] - interp_time)) n = S->memind[1];

  const fpb_met_ptrs *M = (ngrid == 0) ? &S->met[n] : &S->metn[ngrid][n];
  const size_t ldx = (size_t)(ngrid == 0 ? c->nxmax : c->nxmaxn), ldy = (size_t)(ngrid == 0 ? c->nymax : c->nymaxn);
  if (ngrid == 0)
    interpol_rain(M, ldx, c->nx, c->ny, (float)xd, (float)yd, &lsp, &convp, &cc);
  else
    interpol_rain(M, ldx, c->nxn[ngrid - 1], c->nyn[ngrid - 1], xtn, ytn, &lsp, &convp, &cc);

  if ((lsp < 0.01f) && (convp < 0.01f)) return wetscav; /* :131 */

  for (int il = 2; il <= c->nz; il++)
    if (S->height[il] > S->ztra1[jpart]) {
      hz = il - 1;
      break;
    }
  const size_t i3 = (size_t)ix + ldx * ((size_t)jy + ldy * (size_t)(hz - 1));
  const int clouds_v = M->clouds[i3];
  if (clouds_v <= 1) return wetscav; /* :152 */

  if (lsp > 20.f) i = 5; else if (lsp > 8.f) i = 4; else if (lsp > 3.f) i = 3; else if (lsp > 1.f) i = 2; else i = 1;
  if (convp > 20.f) j = 5; else if (convp > 8.f) j = 4; else if (convp > 3.f) j = 3; else if (convp > 1.f) j = 2; else j = 1;

  *grfraction1 = fpo_maxf(0.05f, cc * (lsp * lfr[i - 1] + convp * cfr[j - 1]) / (lsp + convp)); /* :190 */
  prec1 = (lsp + convp) / *grfraction1;                                                       /* :194 */
  act_temp = M->tt[i3];                                                                       /* :200-204 */

  if (clouds_v >= 4) { /* below-cloud scavenging, :209-248 */
    if ((c->dquer[ks - 1] <= 0.f) && (c->weta_gas[ks - 1] > 0.f || c->wetb_gas[ks - 1] > 0.f)) {
      wetscav = c->weta_gas[ks - 1] * fpo_powf(prec1, c->wetb_gas[ks - 1]);
    } else if ((c->dquer[ks - 1] > 0.f) && (c->crain_aero[ks - 1] > 0.f || c->csnow_aero[ks - 1] > 0.f)) {
      const float dquer_m = fpo_minf(10.f, c->dquer[ks - 1]) / 1000000.f;
      const float lg = log10_f(dquer_m);
      if (act_temp >= 273.f && c->crain_aero[ks - 1] > 0.f) {
        wetscav = c->crain_aero[ks - 1] *
                  fpo_powf(10.f, bclr[0] + (bclr[1] * powi(lg, -4)) + (bclr[2] * powi(lg, -3)) +
                                     (bclr[3] * powi(lg, -2)) + (bclr[4] * powi(lg, -1)) +
                                     bclr[5] * fpo_powf(prec1, 0.5f));
      } else if (act_temp < 273.f && c->csnow_aero[ks - 1] > 0.f) {
        wetscav = c->csnow_aero[ks - 1] *
                  fpo_powf(10.f, bcls[0] + (bcls[1] * powi(lg, -4)) + (bcls[2] * powi(lg, -3)) +
                                     (bcls[3] * powi(lg, -2)) + (bcls[4] * powi(lg, -1)) +
                                     bcls[5] * fpo_powf(prec1, 0.5f));
      }
    }
  }

  if (clouds_v < 4) { /* in-cloud scavenging, :253-311 */
    float ccn = c->ccn_aero[ks - 1], in = c->in_aero[ks - 1];
    if ((ccn > 0.f || in > 0.f) || (c->henry[ks - 1] > 0.f && c->dquer[ks - 1] <= 0.f)) {
      float liq_frac, ice_frac, frac_act, S_i;
      if (ccn < 0.f) ccn = 0.f; /* the reference overwrites ccn_aero/in_aero: idempotent */
      if (in < 0.f) in = 0.f;
      if (ngrid > 0 && readclouds_this_nest)
        cl = M->ctwc[(size_t)ix + ldx * (size_t)jy] * (*grfraction1 / cc);
      else if (ngrid == 0 && c->readclouds)
        cl = M->ctwc[(size_t)ix + ldx * (size_t)jy] * (*grfraction1 / cc);
      else
        cl = (1.e6f * 2.e-7f) * fpo_powf(prec1, 0.36f);
      if (act_temp <= 253.f) {
        liq_frac = 0.f;
        ice_frac = 1.f;
      } else if (act_temp >= 273.f) {
        liq_frac = 1.f;
        ice_frac = 0.f;
      } else {
        ice_frac = fpo_powf((act_temp - 273.f) / (273.f - 253.f), 2.f);
        liq_frac = fpo_maxf(0.f, 1.f - ice_frac);
      }
      frac_act = liq_frac * ccn + ice_frac * in;
      if (c->dquer[ks - 1] > 0.f) {
        S_i = frac_act / cl;
      } else {
        const float cle = (1.f - cl) / (c->henry[ks - 1] * (r_air / 3500.f) * act_temp) + cl;
        S_i = 1.f / cle;
      }
      wetscav = incloud_ratio * S_i * (prec1 / 3.6e6f);
    }
  }
  return wetscav;
}

/* wetgridunc(0:nxg-1,0:nyg-1,maxspec,maxpointspec_act,nclassunc,maxageclass) */
static size_t widx(const fpb_config *c, int nxg, int nyg, int ix, int jy, int ks, int kp, int nc, int na) {
  size_t i = (size_t)(na - 1);
  i = i * c->nclassunc + (nc - 1);
  i = i * c->maxpointspec_act + (kp - 1);
  i = i * c->maxspec + (ks - 1);
  i = i * nyg + jy;
  i = i * nxg + ix;
  return i;
}

/* src/wetdepokernel.f90:38-108 (nest = 0) / src/wetdepokernel_nest.f90:38-105
 * (nest = 1: floor() instead of int(), always the 4-cell kernel) */
static void wetdepokernel(fpo_state *S, int nunc, const float *deposit, float x, float y,
                          int nage, int kp, int nest) {
  const fpb_config *c = &S->c;
  const int nxg = nest ? c->numxgridn : c->numxgrid, nyg = nest ? c->numygridn : c->numygrid;
  float *grid = nest ? S->wetgriduncn : S->wetgridunc;
  float xl, yl, ddx, ddy, wx, wy, w;
  int ix, jy, ixp, jyp;
  if (nest) {
    xl = (x * c->dx + c->xoutshiftn) / c->dxoutn;
    yl = (y * c->dy + c->youtshiftn) / c->dyoutn;
    ix = (int)floorf(xl);
    jy = (int)floorf(yl);
  } else {
    xl = (x * c->dx + c->xoutshift) / c->dxout;
    yl = (y * c->dy + c->youtshift) / c->dyout;
    ix = fpo_int_f(xl);
    jy = fpo_int_f(yl);
  }
  ddx = xl - (float)ix;
  ddy = yl - (float)jy;
  if (ddx > 0.5f) { ixp = ix + 1; wx = 1.5f - ddx; } else { ixp = ix - 1; wx = 0.5f + ddx; }
  if (ddy > 0.5f) { jyp = jy + 1; wy = 1.5f - ddy; } else { jyp = jy - 1; wy = 0.5f + ddy; }
  if (!nest && !c->lusekerneloutput) {
    for (int ks = 1; ks <= c->nspec; ks++)
      if ((ix >= 0) && (jy >= 0) && (ix <= nxg - 1) && (jy <= nyg - 1))
        grid[widx(c, nxg, nyg, ix, jy, ks, kp, nunc, nage)] += deposit[ks - 1];
    return;
  }
  for (int ks = 1; ks <= c->nspec; ks++) {
    if ((ix >= 0) && (jy >= 0) && (ix <= nxg - 1) && (jy <= nyg - 1)) {
      w = wx * wy;
      grid[widx(c, nxg, nyg, ix, jy, ks, kp, nunc, nage)] += deposit[ks - 1] * w;
    }
    if ((ixp >= 0) && (jyp >= 0) && (ixp <= nxg - 1) && (jyp <= nyg - 1)) {
      w = (1.f - wx) * (1.f - wy);
      grid[widx(c, nxg, nyg, ixp, jyp, ks, kp, nunc, nage)] += deposit[ks - 1] * w;
    }
    if ((ixp >= 0) && (jy >= 0) && (ixp <= nxg - 1) && (jy <= nyg - 1)) {
      w = (1.f - wx) * wy;
      grid[widx(c, nxg, nyg, ixp, jy, ks, kp, nunc, nage)] += deposit[ks - 1] * w;
    }
    if ((ix >= 0) && (jyp >= 0) && (ix <= nxg - 1) && (jyp <= nyg - 1)) {
      w = wx * (1.f - wy);
      grid[widx(c, nxg, nyg, ix, jyp, ks, kp, nunc, nage)] += deposit[ks - 1] * w;
    }
  }
}

/* src/wetdepo.f90:70-147; ldeltat as computed at :55-63 by the caller */
void fpo_wetdepo(fpo_state *S, int itime, int ltsample, int ldeltat) {
  const fpb_config *c = &S->c;
  float wetdeposit[FPB_MAXSPEC];
  float grfraction1 = 0.f;
  int kp = 1;
  memset(wetdeposit, 0, sizeof wetdeposit); /* -finit-local-zero */
  for (int jpart = 1; jpart <= S->numpart; jpart++) {
    int nage;
    if (S->itra1[jpart] == -999999999) continue;
    if (c->ldirect == 1) {
      if (S->itra1[jpart] > itime) continue;
    } else {
      if (S->itra1[jpart] < itime) continue;
    }
    const int itage = abs(S->itra1[jpart] - S->itramem[jpart]);
    for (nage = 1; nage <= c->nageclass; nage++)
      if (itage < c->lage[nage - 1]) break;

    for (int ks = 1; ks <= c->nspec; ks++) {
      if (!c->wetdepspec[ks - 1]) continue;
      const float wetscav = get_wetscav(S, itime, ltsample, jpart, ks, &grfraction1);
      float *xm = &S->xmass1[(size_t)jpart + (size_t)(S->maxpart + 1) * (ks - 1)];
      if (wetscav > 0.f)
        wetdeposit[ks - 1] = *xm * (1.f - fpo_expf(-wetscav * (float)abs(ltsample))) * grfraction1;
      else
        wetdeposit[ks - 1] = 0.f;
      const float restmass = *xm - wetdeposit[ks - 1];
      kp = (c->ioutputforeachrelease == 1) ? S->npoint[jpart] : 1;
      if (restmass > smallnum)
        *xm = restmass;
      else
        *xm = 0.f;
      if (c->decay[ks - 1] > 0.f)
        wetdeposit[ks - 1] = wetdeposit[ks - 1] * fpo_expf((float)abs(ldeltat) * c->decay[ks - 1]);
    }
    if (c->ldirect == 1) {
      wetdepokernel(S, S->nclass[jpart], wetdeposit, (float)S->xtra1[jpart], (float)S->ytra1[jpart], nage, kp, 0);
      if (c->nested_output == 1)
        wetdepokernel(S, S->nclass[jpart], wetdeposit, (float)S->xtra1[jpart], (float)S->ytra1[jpart], nage, kp, 1);
    }
  }
}

void fpo_fetch_wetgrids(fpo_state *S, float *wetgridunc, float *wetgriduncn) {
  const fpb_config *c = &S->c;
  const size_t per = (size_t)c->maxspec * c->maxpointspec_act * c->nclassunc * c->maxageclass;
  if (wetgridunc) memcpy(wetgridunc, S->wetgridunc, per * c->numxgrid * c->numygrid * sizeof(float));
  if (wetgriduncn && c->nested_output == 1)
    memcpy(wetgriduncn, S->wetgriduncn, per * c->numxgridn * c->numygridn * sizeof(float));
}
