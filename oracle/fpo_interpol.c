/*
 * fpo_interpol.c -- oracle restatement of the interpol_* family
 * (test infrastructure).  Scratch lives in fpo_state exactly as it lives in
 * interpol_mod / hanna_mod in the reference (src/interpol_mod.f90,
 * src/hanna_mod.f90), so the call order dependencies are preserved.
 */
#include "fpo.h"
#include "fpo_math.h"

/* Array extents and time level m of the grid in use: the mother grid
 * (uu(0:nxmax-1,0:nymax-1,nzmax,slot)) or, when S->ngrid > 0, nested grid
 * ngrid (uun(0:nxmaxn-1,0:nymaxn-1,nzmax,slot,ngrid)).  The *_nests routines
 * of the reference (src/interpol_all_nests.f90:59-170,
 * src/interpol_misslev_nests.f90, src/interpol_wind_nests.f90,
 * src/interpol_wind_short_nests.f90, src/interpol_vdep_nests.f90) are the
 * mother-grid routines with these arrays and without the polar branch. */
#define LDX(S) ((size_t)((S)->ngrid > 0 ? (S)->c.nxmaxn : (S)->c.nxmax))
#define LDY(S) ((size_t)((S)->ngrid > 0 ? (S)->c.nymaxn : (S)->c.nymax))
#define IDX3(S, i, j, k) ((size_t)(i) + LDX(S) * ((size_t)(j) + LDY(S) * (size_t)((k)-1)))
#define IDX2(S, i, j) ((size_t)(i) + LDX(S) * (size_t)(j))
static inline const fpb_met_ptrs *grid_met(const fpo_state *S, int m) {
  return (S->ngrid > 0) ? &S->metn[S->ngrid][S->memind[m]] : &S->met[S->memind[m]];
}

static const float EPS_SIG = 1.0e-30f;

/* src/interpol_all.f90:57-71 and the same block in interpol_wind(_short) */
static void weights(fpo_state *S, int itime, float xt, float yt) {
  S->ddx = xt - (float)S->ix;
  S->ddy = yt - (float)S->jy;
  S->rddx = 1.f - S->ddx;
  S->rddy = 1.f - S->ddy;
  S->p1 = S->rddx * S->rddy;
  S->p2 = S->ddx * S->rddy;
  S->p3 = S->rddx * S->ddy;
  S->p4 = S->ddx * S->ddy;
  S->dt1 = (float)(itime - S->memtime[1]);
  S->dt2 = (float)(S->memtime[2] - itime);
  S->dtt = 1.f / (S->dt1 + S->dt2);
}

void fpo_interpol_weights(fpo_state *S, int itime, float xt, float yt) { weights(S, itime, xt, yt); }

static inline float bilin(const fpo_state *S, const float *f, size_t a,
                          size_t b, size_t c, size_t d) {
  return S->p1 * f[a] + S->p2 * f[b] + S->p3 * f[c] + S->p4 * f[d];
}

/* one level n of the profile arrays: the body shared by
 * src/interpol_all.f90:135-238 and src/interpol_misslev.f90:56-157 */
static void profile_level(fpo_state *S, int n) {
  float y1[3], y2[3], y3[3], rho1[3], rhograd1[3];
  float usl = 0.f, vsl = 0.f, wsl = 0.f, usq = 0.f, vsq = 0.f, wsq = 0.f, xaux;
  for (int m = 1; m <= 2; m++) {
    const fpb_met_ptrs *M = grid_met(S, m);
    size_t a = IDX3(S, S->ix, S->jy, n), b = IDX3(S, S->ixp, S->jy, n),
           c = IDX3(S, S->ix, S->jyp, n), d = IDX3(S, S->ixp, S->jyp, n);
    const float *fu = (S->ngrid < 0) ? M->uupol : M->uu;
    const float *fv = (S->ngrid < 0) ? M->vvpol : M->vv;
    y1[m] = bilin(S, fu, a, b, c, d);
    y2[m] = bilin(S, fv, a, b, c, d);
    usl = usl + fu[a] + fu[b] + fu[c] + fu[d];
    vsl = vsl + fv[a] + fv[b] + fv[c] + fv[d];
    usq = usq + fu[a] * fu[a] + fu[b] * fu[b] + fu[c] * fu[c] + fu[d] * fu[d];
    vsq = vsq + fv[a] * fv[a] + fv[b] * fv[b] + fv[c] * fv[c] + fv[d] * fv[d];
    y3[m] = bilin(S, M->ww, a, b, c, d);
    rhograd1[m] = bilin(S, M->drhodz, a, b, c, d);
    rho1[m] = bilin(S, M->rho, a, b, c, d);
    wsl = wsl + M->ww[a] + M->ww[b] + M->ww[c] + M->ww[d];
    wsq = wsq + M->ww[a] * M->ww[a] + M->ww[b] * M->ww[b] +
          M->ww[c] * M->ww[c] + M->ww[d] * M->ww[d];
  }
  S->uprof[n] = (y1[1] * S->dt2 + y1[2] * S->dt1) * S->dtt;
  S->vprof[n] = (y2[1] * S->dt2 + y2[2] * S->dt1) * S->dtt;
  S->wprof[n] = (y3[1] * S->dt2 + y3[2] * S->dt1) * S->dtt;
  S->rhoprof[n] = (rho1[1] * S->dt2 + rho1[2] * S->dt1) * S->dtt;
  S->rhogradprof[n] = (rhograd1[1] * S->dt2 + rhograd1[2] * S->dt1) * S->dtt;
  S->indzindicator[n] = 0;

  /* standard deviations over the 8 surrounding values */
  xaux = usq - usl * usl / 8.f;
  S->usigprof[n] = (xaux < EPS_SIG) ? 0.f : fpo_sqrtf(xaux / 7.f);
  xaux = vsq - vsl * vsl / 8.f;
  S->vsigprof[n] = (xaux < EPS_SIG) ? 0.f : fpo_sqrtf(xaux / 7.f);
  xaux = wsq - wsl * wsl / 8.f;
  S->wsigprof[n] = (xaux < EPS_SIG) ? 0.f : fpo_sqrtf(xaux / 7.f);
}

/* src/interpol_all.f90:57-240 */
void fpo_interpol_all(fpo_state *S, int itime, float xt, float yt, float zt) {
  float ust1[3], wst1[3], oli1[3], oliaux;
  weights(S, itime, xt, yt);

  for (int m = 1; m <= 2; m++) {
    const fpb_met_ptrs *M = grid_met(S, m);
    size_t a = IDX2(S, S->ix, S->jy), b = IDX2(S, S->ixp, S->jy),
           c = IDX2(S, S->ix, S->jyp), d = IDX2(S, S->ixp, S->jyp);
    ust1[m] = bilin(S, M->ustar, a, b, c, d);
    wst1[m] = bilin(S, M->wstar, a, b, c, d);
    oli1[m] = bilin(S, M->oli, a, b, c, d);
  }
  S->ust = (ust1[1] * S->dt2 + ust1[2] * S->dt1) * S->dtt;
  S->wst = (wst1[1] * S->dt2 + wst1[2] * S->dt1) * S->dtt;
  oliaux = (oli1[1] * S->dt2 + oli1[2] * S->dt1) * S->dtt;
  if (oliaux != 0.f)
    S->ol = 1.f / oliaux;
  else
    S->ol = 99999.f;

  /* level search, src/interpol_all.f90:118-125 */
  for (int i = 2; i <= S->c.nz; i++) {
    if (S->height[i] > zt) {
      S->indz = i - 1;
      S->indzp = i;
      break;
    }
  }
  for (int n = S->indz; n <= S->indzp; n++) profile_level(S, n);
}

/* src/interpol_misslev.f90:56-159 */
void fpo_interpol_misslev(fpo_state *S, int n) { profile_level(S, n); }

/* src/interpol_wind.f90:56-214 (with_sigma) and
 * src/interpol_wind_short.f90:48-140 (without) */
static void wind_common(fpo_state *S, int itime, float xt, float yt, float zt,
                        int with_sigma) {
  float dz1, dz2, dz;
  float u1[3], v1[3], w1[3], uh[3], vh[3], wh[3];
  float usl = 0.f, vsl = 0.f, wsl = 0.f, usq = 0.f, vsq = 0.f, wsq = 0.f, xaux;
  weights(S, itime, xt, yt);

  for (int i = 2; i <= S->c.nz; i++) {
    if (S->height[i] > zt) {
      S->indz = i - 1;
      break;
    }
  }
  dz = 1.f / (S->height[S->indz + 1] - S->height[S->indz]);
  dz1 = (zt - S->height[S->indz]) * dz;
  dz2 = (S->height[S->indz + 1] - zt) * dz;

  for (int m = 1; m <= 2; m++) {
    const fpb_met_ptrs *M = grid_met(S, m);
    const float *fu = (S->ngrid < 0) ? M->uupol : M->uu;
    const float *fv = (S->ngrid < 0) ? M->vvpol : M->vv;
    for (int n = 1; n <= 2; n++) {
      int indzh = S->indz + n - 1;
      size_t a = IDX3(S, S->ix, S->jy, indzh), b = IDX3(S, S->ixp, S->jy, indzh),
             c = IDX3(S, S->ix, S->jyp, indzh),
             d = IDX3(S, S->ixp, S->jyp, indzh);
      u1[n] = bilin(S, fu, a, b, c, d);
      v1[n] = bilin(S, fv, a, b, c, d);
      w1[n] = bilin(S, M->ww, a, b, c, d);
      if (with_sigma) {
        usl = usl + fu[a] + fu[b] + fu[c] + fu[d];
        vsl = vsl + fv[a] + fv[b] + fv[c] + fv[d];
        usq = usq + fu[a] * fu[a] + fu[b] * fu[b] + fu[c] * fu[c] + fu[d] * fu[d];
        vsq = vsq + fv[a] * fv[a] + fv[b] * fv[b] + fv[c] * fv[c] + fv[d] * fv[d];
        wsl = wsl + M->ww[a] + M->ww[b] + M->ww[c] + M->ww[d];
        wsq = wsq + M->ww[a] * M->ww[a] + M->ww[b] * M->ww[b] +
              M->ww[c] * M->ww[c] + M->ww[d] * M->ww[d];
      }
    }
    uh[m] = dz2 * u1[1] + dz1 * u1[2];
    vh[m] = dz2 * v1[1] + dz1 * v1[2];
    wh[m] = dz2 * w1[1] + dz1 * w1[2];
  }
  S->u = (uh[1] * S->dt2 + uh[2] * S->dt1) * S->dtt;
  S->v = (vh[1] * S->dt2 + vh[2] * S->dt1) * S->dtt;
  S->w = (wh[1] * S->dt2 + wh[2] * S->dt1) * S->dtt;

  if (with_sigma) {
    xaux = usq - usl * usl / 16.f;
    S->usig = (xaux < EPS_SIG) ? 0.f : fpo_sqrtf(xaux / 15.f);
    xaux = vsq - vsl * vsl / 16.f;
    S->vsig = (xaux < EPS_SIG) ? 0.f : fpo_sqrtf(xaux / 15.f);
    xaux = wsq - wsl * wsl / 16.f;
    S->wsig = (xaux < EPS_SIG) ? 0.f : fpo_sqrtf(xaux / 15.f);
  }
}

void fpo_interpol_wind(fpo_state *S, int itime, float xt, float yt, float zt) {
  wind_common(S, itime, xt, yt, zt, 1);
}
void fpo_interpol_wind_short(fpo_state *S, int itime, float xt, float yt,
                             float zt) {
  wind_common(S, itime, xt, yt, zt, 0);
}

/* src/interpol_vdep.f90:39-54: reuses p1..p4, dt1, dt2, dtt of the last
 * interpol_all call. */
void fpo_interpol_vdep(fpo_state *S, int level, float *vdepo) {
  float y[3];
  for (int m = 1; m <= 2; m++) {
    const fpb_met_ptrs *M = grid_met(S, m);
    size_t a = IDX3(S, S->ix, S->jy, level), b = IDX3(S, S->ixp, S->jy, level),
           c = IDX3(S, S->ix, S->jyp, level),
           d = IDX3(S, S->ixp, S->jyp, level);
    y[m] = bilin(S, M->vdep, a, b, c, d);
  }
  *vdepo = (y[1] * S->dt2 + y[2] * S->dt1) * S->dtt;
  S->depoindicator[level] = 0;
}
