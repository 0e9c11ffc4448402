/*
 * fpo_advance.c -- oracle restatement of initialize() and advance()
 * (test infrastructure).  Follows src/initialize.f90:4-219 and
 * src/advance.f90:4-988 statement by statement, including the label
 * structure (100 = PBL time loop, 700 = above-PBL step, 99 = common tail).
 *
 * Typing: `real` = float, xt/yt = double; mixed expressions are evaluated
 * the way Fortran promotes them (see SURVEY.md 8c "arithmetic traps").
 *
 * strict_reference == 0 ("defined" behaviour, what the device implements):
 *   (1) a particle leaving the PBL on the last sub-step gets usig/vsig/wsig
 *       from its own cached levels (the reference reads stale module state,
 *       src/advance.f90:549-551 vs :604-606);
 *   (2) initialize() derives ngrid from the particle's own latitude (the
 *       reference reads the ngrid left by the previous advance call,
 *       src/interpol_all.f90:144);
 *   (3) the Petterssen step clamps jyp like the first stage does
 *       (src/advance.f90:228-231 vs :871-872).
 */
#include <math.h>
#include <stdlib.h>

#include "fpo.h"
#include "fpo_math.h"

static const float PI_F = 3.14159265f;
#define pi180 (PI_F / 180.f)
static const float href = 15.f;
static const float eps2 = 1.e-9f;
#define eps3 1.17549435e-38f /* tiny(1.0) */

#define IDX2(S, i, j) ((size_t)(i) + (size_t)(S)->c.nxmax * (size_t)(j))
#define IDX2N(S, i, j) ((size_t)(i) + (size_t)(S)->c.nxmaxn * (size_t)(j))

static int pole_grid(const fpo_state *S, double yt) {
  if (S->c.nglobal && (yt > S->c.switchnorthg)) return -1;
  if (S->c.sglobal && (yt < S->c.switchsouthg)) return -2;
  return 0;
}

/* grid choice incl. the nesting level, src/advance.f90:161-175 (= :841-856) */
static int choose_grid(const fpo_state *S, double xt, double yt) {
  const fpb_config *c = &S->c;
  const float eps = c->eps;
  int ngrid = pole_grid(S, yt);
  if (ngrid != 0) return ngrid;
  for (int j = c->numbnests; j >= 1; j--)
    if ((xt > c->xln[j - 1] + eps) && (xt < c->xrn[j - 1] - eps) &&
        (yt > c->yln[j - 1] + eps) && (yt < c->yrn[j - 1] - eps))
      return j;
  return 0;
}

/* src/get_vdep_prob.f90:41-124: the deposition velocity at a receptor particle's position (backward
 * dry-deposition runs, called once per particle right after its release, src/timemanager.f90:571-583).
 * The reference's interpol_vdep reuses the weights p1..p4, dt1, dt2, dtt the previous interpol_* call
 * left in interpol_mod -- the particle's own initialize() call on the mother grid; strict_reference
 * keeps that, the "defined" mode (the device) computes the weights of the grid it reads from. */
void fpo_get_vdep_prob(fpo_state *S, int itime, double xt, double yt, float zt, float *prob) {
  const fpb_config *c = &S->c;
  const float href = 15.f;
  float xf, yf, vdepo[FPB_MAXSPEC + 1];
  if (c->drydep)
    for (int ks = 1; ks <= c->nspec; ks++) {
      S->depoindicator[ks] = 1;
      prob[ks - 1] = 0.f;
    }
  S->ngrid = choose_grid(S, xt, yt);
  if (S->ngrid > 0) {
    xf = (float)((xt - c->xln[S->ngrid - 1]) * c->xresoln[S->ngrid - 1]);
    yf = (float)((yt - c->yln[S->ngrid - 1]) * c->yresoln[S->ngrid - 1]);
    S->ix = fpo_int_f(xf);
    S->jy = fpo_int_f(yf);
  } else {
    xf = (float)xt;
    yf = (float)yt;
    S->ix = fpo_int_d(xt);
    S->jy = fpo_int_d(yt);
  }
  S->ixp = S->ix + 1;
  S->jyp = S->jy + 1;
  if (!S->strict_reference) { /* (row ny exists in neither array: its weight is 0 there) */
    const int nyd = S->ngrid > 0 ? c->nyn[S->ngrid - 1] : ((c->ny < c->nymax) ? c->ny + 1 : c->ny);
    if (S->jyp > nyd - 1) S->jyp = nyd - 1;
    fpo_interpol_weights(S, itime, xf, yf);
  }
  if (c->drydep && (zt < 2.f * href))
    for (int ks = 1; ks <= c->nspec; ks++)
      if (c->drydepspec[ks - 1]) {
        if (S->depoindicator[ks]) fpo_interpol_vdep(S, ks, &vdepo[ks]);
        prob[ks - 1] = vdepo[ks];
      }
}

/* settling species choice, src/advance.f90:518-531 (and :686-699, :893-906) */
static void add_settling(fpo_state *S, int itime, int nrelpoint, double xt,
                         double yt, float zt) {
  if (S->c.mdomainfill == 0) {
    if (S->c.lsettling) {
      int nsp;
      for (nsp = 1; nsp <= S->c.nspec; nsp++)
        if (S->xmass[(nrelpoint - 1) + (size_t)S->c.numpoint * (nsp - 1)] > eps3) break;
      if (nsp > S->c.nspec) nsp = S->c.nspec;
      if (S->c.density[nsp - 1] > 0.f) {
        fpo_get_settling(S, itime, (float)xt, (float)yt, zt, nsp, &S->settling_saved);
        S->w = S->w + S->settling_saved;
      }
    }
  }
}

/* position update with the accumulated displacement (src/advance.f90:750-778)
 * or the Petterssen half-difference (src/advance.f90:923-951) */
static void move_horizontal(fpo_state *S, double *xt, double *yt, float ddx_m,
                            float ddy_m, float tfac) {
  const fpb_config *c = &S->c;
  if (S->ngrid >= 0) {
    float cosfact = (float)(c->dxconst / cos((*yt * c->dy + c->ylat0) * pi180));
    *xt = *xt + (double)(ddx_m * cosfact * tfac);
    *yt = *yt + (double)(ddy_m * c->dyconst * tfac);
  } else {
    const float *map = (S->ngrid == -1) ? c->northpolemap : c->southpolemap;
    float xlon = (float)(c->xlon0 + *xt * c->dx);
    float ylat = (float)(c->ylat0 + *yt * c->dy);
    float xpol, ypol;
    fpo_cll2xy(map, ylat, xlon, &xpol, &ypol);
    float gridsize = 1000.f * fpo_cgszll(map, ylat, xlon);
    ddx_m = ddx_m / gridsize;
    ddy_m = ddy_m / gridsize;
    xpol = xpol + ddx_m * tfac;
    ypol = ypol + ddy_m * tfac;
    fpo_cxy2ll(map, xpol, ypol, &ylat, &xlon);
    *xt = (xlon - c->xlon0) / c->dx;
    *yt = (ylat - c->ylat0) / c->dy;
  }
}

/* cyclic boundary + pole crossing + exit test, src/advance.f90:784-808.
 * returns 1 if the particle left the domain (nstop = 3). */
static int wrap_and_check(const fpo_state *S, double *xt, double *yt) {
  const fpb_config *c = &S->c;
  const float eps = c->eps;
  if (c->xglobal) {
    if (*xt >= (float)c->nxmin1) *xt = *xt - (float)c->nxmin1;
    if (*xt < 0.) *xt = *xt + (float)c->nxmin1;
    if (*xt <= eps) *xt = eps;
    if (fabs(*xt - (float)c->nxmin1) <= eps) *xt = (float)c->nxmin1 - eps;
    if (*yt < 0.) {
      *xt = fpo_modulo_d(*xt * c->dx + 180.f, 360.f) / c->dx;
      *yt = -*yt;
    } else if (*yt > (float)c->nymin1) {
      *xt = fpo_modulo_d(*xt * c->dx + 180.f, 360.f) / c->dx;
      *yt = 2 * (float)c->nymin1 - *yt;
    }
  }
  if ((*xt < 0.) || (*xt >= (float)c->nxmin1) || (*yt < 0.) ||
      (*yt > (float)c->nymin1))
    return 1;
  return 0;
}

/* src/initialize.f90:4-219 */
void fpo_initialize(fpo_state *S, int itime, int32_t *ldt, float *up, float *vp,
                    float *wp, float *usigold, float *vsigold, float *wsigold,
                    double xt, double yt, float zt, int16_t *icbt) {
  const fpb_config *c = &S->c;
  float dz, dz1, dz2;
  int nrand;
  const float *rn = S->rannumb;
  const int maxrand = S->maxrand;

  *icbt = 1;
  nrand = fpo_int_f(fpo_index_uniform(S, &S->idummy_initialize) * (float)(maxrand - 1)) + 1;

  if (!S->strict_reference) S->ngrid = pole_grid(S, yt);
  /* initialize() calls the mother-grid routines whatever ngrid a previous advance left:
   * interpol_all only tests ngrid < 0 (src/initialize.f90:102, src/interpol_all.f90:144) */
  if (S->ngrid > 0) S->ngrid = 0;

  S->ix = fpo_int_d(xt);
  S->jy = fpo_int_d(yt);
  S->ixp = S->ix + 1;
  S->jyp = S->jy + 1;
  if (!S->strict_reference && S->jyp >= c->nymax) S->jyp = S->jyp - 1;

  {
    const float *h1 = S->met[S->memind[1]].hmix, *h2 = S->met[S->memind[2]].hmix;
    float h = h1[IDX2(S, S->ix, S->jy)];
    h = fpo_maxf(h, h1[IDX2(S, S->ixp, S->jy)]);
    h = fpo_maxf(h, h1[IDX2(S, S->ix, S->jyp)]);
    h = fpo_maxf(h, h1[IDX2(S, S->ixp, S->jyp)]);
    h = fpo_maxf(h, h2[IDX2(S, S->ix, S->jy)]);
    h = fpo_maxf(h, h2[IDX2(S, S->ixp, S->jy)]);
    h = fpo_maxf(h, h2[IDX2(S, S->ix, S->jyp)]);
    h = fpo_maxf(h, h2[IDX2(S, S->ixp, S->jyp)]);
    S->h = h;
  }
  S->zeta = zt / S->h;

  if (S->zeta <= 1.f) {
    fpo_interpol_all(S, itime, (float)xt, (float)yt, zt);

    dz1 = zt - S->height[S->indz];
    dz2 = S->height[S->indzp] - zt;
    dz = 1.f / (dz1 + dz2);
    S->u = (dz1 * S->uprof[S->indzp] + dz2 * S->uprof[S->indz]) * dz;
    S->v = (dz1 * S->vprof[S->indzp] + dz2 * S->vprof[S->indz]) * dz;
    S->w = (dz1 * S->wprof[S->indzp] + dz2 * S->wprof[S->indz]) * dz;

    if (c->turbswitch)
      fpo_hanna(S, zt);
    else
      fpo_hanna1(S, zt);

    if (nrand + 2 > maxrand) nrand = 1;
    *up = rn[nrand] * S->sigu;
    *vp = rn[nrand + 1] * S->sigv;
    *wp = rn[nrand + 2];
    if (!c->turbswitch) {
      *wp = *wp * S->sigw;
    } else if (c->cblflag == 1) {
      if (-S->h / S->ol > 5.f) {
        if (S->strict_reference) {
          fpo_initialize_cbl_vel(S, &S->idummy_initialize, zt, S->ust, S->wst,
                                 S->h, S->sigw, wp, S->ol);
        } else {
          /* "defined" behaviour: the mode selector is the uniform that chose
           * the table index, the normal is the next unused table entry
           * (the reference pulls ran3 + gasdev from the sequential stream) */
          fpo_initialize_cbl_vel_defined(S, (float)(nrand - 1) / (float)(maxrand - 1),
                                         rn[nrand + 3], zt, S->wst, S->h, S->sigw,
                                         wp, S->ol);
        }
      } else
        *wp = *wp * S->sigw;
    }

    if (c->turbswitch) {
      float t = fpo_minf(S->tlw, S->h / fpo_maxf(2.f * fabsf(*wp * S->sigw), 1.e-5f));
      t = fpo_minf(t, 0.5f / fabsf(S->dsigwdz));
      t = fpo_minf(t, 600.f);
      *ldt = fpo_int_f(t * c->ctl);
    } else {
      float t = fpo_minf(S->tlw, S->h / fpo_maxf(2.f * fabsf(*wp), 1.e-5f));
      t = fpo_minf(t, 600.f);
      *ldt = fpo_int_f(t * c->ctl);
    }
    *ldt = (*ldt > c->mintime) ? *ldt : c->mintime;

    S->usig = (S->usigprof[S->indzp] + S->usigprof[S->indz]) / 2.f;
    S->vsig = (S->vsigprof[S->indzp] + S->vsigprof[S->indz]) / 2.f;
    S->wsig = (S->wsigprof[S->indzp] + S->wsigprof[S->indz]) / 2.f;
  } else {
    fpo_interpol_wind(S, itime, (float)xt, (float)yt, zt);
    *ldt = abs(c->lsynctime);
    if (nrand + 1 > maxrand) nrand = 1;
    *up = rn[nrand] * 0.3f;
    *vp = rn[nrand + 1] * 0.3f;
    nrand = nrand + 2;
    *wp = 0.f;
    S->sigw = 0.f;
  }

  if (nrand + 2 > maxrand) nrand = 1;
  *usigold = rn[nrand] * S->usig;
  *vsigold = rn[nrand + 1] * S->vsig;
  *wsigold = rn[nrand + 2] * S->wsig;
}

/* src/advance.f90:4-988 */
void fpo_advance(fpo_state *S, int itime, int nrelpoint, int32_t *ldt_io,
                 float *up_io, float *vp_io, float *wp_io, float *usigold_io,
                 float *vsigold_io, float *wsigold_io, int *nstop, double *xt_io,
                 double *yt_io, float *zt_io, float *prob, int16_t *icbt_io) {
  const fpb_config *c = &S->c;
  const float *rn = S->rannumb;
  const int maxrand = S->maxrand;
  const float eps = c->eps;
  const int nz = c->nz;
  const float *height = S->height;

  double xt = *xt_io, yt = *yt_io;
  float zt = *zt_io, up = *up_io, vp = *vp_io, wp = *wp_io;
  float usigold = *usigold_io, vsigold = *vsigold_io, wsigold = *wsigold_io;
  int ldt = *ldt_io;
  int icbt = *icbt_io;

  float xts, yts, xtn = 0.f, ytn = 0.f, weight;
  int itimec, i, nrand, loop, memindnext, ngr, nix, njy, ks;
  float dz, dz1, dz2;
  float ru, rv, rw, dt, ux = 0.f, vy = 0.f, tropop;
  float dxsave, dysave, dawsave, dcwsave;
  float r, rs, uold, vold, wold, vdepo[FPB_MAXSPEC + 1];
  float rhoa = 0.f, rhograd = 0.f, delz = 0.f, dtf, rhoaux, dtftlw, uxscale, wpscale;
  float ptot_lhh, Q_lhh, phi_lhh, ath, bth, old_wp_buf, del_test;
  int flagrein;
  int nsub = 0, took_pbl = 0, did_pett = 0;
  (void)memindnext;

  *nstop = 0;
  for (i = 1; i <= nz; i++) S->indzindicator[i] = 1; /* nmixz ~ all levels */

  if (c->drydep) {
    for (ks = 1; ks <= c->nspec; ks++) {
      S->depoindicator[ks] = 1;
      prob[ks - 1] = 0.f;
    }
  }

  dxsave = 0.f;
  dysave = 0.f;
  dawsave = 0.f;
  dcwsave = 0.f;

  itimec = itime;

  nrand = fpo_int_f(fpo_index_uniform(S, &S->idummy_advance) * (float)(maxrand - 1)) + 1;

  /* grid choice, :161-175 */
  S->ngrid = choose_grid(S, xt, yt);

  if (abs(itime - S->memtime[1]) < abs(itime - S->memtime[2]))
    memindnext = 1;
  else
    memindnext = 2;

  /* nested grid coordinates, :191-203 */
  if (S->ngrid > 0) {
    xtn = (float)((xt - c->xln[S->ngrid - 1]) * c->xresoln[S->ngrid - 1]);
    ytn = (float)((yt - c->yln[S->ngrid - 1]) * c->yresoln[S->ngrid - 1]);
    S->ix = fpo_int_f(xtn);
    S->jy = fpo_int_f(ytn);
    nix = fpo_nint_f(xtn);
    njy = fpo_nint_f(ytn);
  } else {
    S->ix = fpo_int_d(xt);
    S->jy = fpo_int_d(yt);
    nix = fpo_nint_d(xt);
    njy = fpo_nint_d(yt);
  }
  S->ixp = S->ix + 1;
  S->jyp = S->jy + 1;

  /* :211-225 (ddx etc. in advance itself mix dp and sp; they are
   * overwritten by interpol_* before use, except p1..p4 for interpolhmix
   * which is compile-time .false.) */
  S->dt1 = (float)(itime - S->memtime[1]);
  S->dt2 = (float)(S->memtime[2] - itime);
  S->dtt = 1.f / (S->dt1 + S->dt2);

  if (S->jyp >= c->nymax) S->jyp = S->jyp - 1; /* :228-231 */

  /* maximum mixing height around the particle, :236-264 */
  S->h = 0.f;
  if (S->ngrid <= 0) {
    for (int k = 1; k <= 2; k++) {
      const float *hm = S->met[S->memind[k]].hmix;
      for (int j = S->jy; j <= S->jyp; j++)
        for (int ii = S->ix; ii <= S->ixp; ii++)
          if (hm[IDX2(S, ii, j)] > S->h) S->h = hm[IDX2(S, ii, j)];
    }
    tropop = S->met[1].tropopause[IDX2(S, nix, njy)]; /* slot 1 literal, :253 */
  } else {
    for (int k = 1; k <= 2; k++) {
      const float *hm = S->metn[S->ngrid][S->memind[k]].hmix;
      for (int j = S->jy; j <= S->jyp; j++)
        for (int ii = S->ix; ii <= S->ixp; ii++)
          if (hm[IDX2N(S, ii, j)] > S->h) S->h = hm[IDX2N(S, ii, j)];
    }
    tropop = S->metn[S->ngrid][1].tropopause[IDX2N(S, nix, njy)]; /* :263 */
  }

  S->zeta = zt / S->h;

  if (S->zeta <= 1.f) {
    took_pbl = 1;
    loop = 0;
  L100:
    loop = loop + 1;
    nsub++;
    if (c->method == 1) {
      int rem = abs(c->lsynctime - itimec + itime);
      ldt = (ldt < rem) ? ldt : rem;
      itimec = itimec + ldt * c->ldirect;
    } else {
      ldt = abs(c->lsynctime);
      itimec = itime + c->lsynctime;
    }
    dt = (float)ldt;

    S->zeta = zt / S->h;

    if (loop == 1) {
      if (S->ngrid <= 0) {
        xts = (float)xt;
        yts = (float)yt;
        fpo_interpol_all(S, itime, xts, yts, zt);
      } else { /* interpol_all_nests(itime,xtn,ytn,zt), :300-302 */
        fpo_interpol_all(S, itime, xtn, ytn, zt);
      }
    } else {
      for (i = 2; i <= nz; i++) {
        if (height[i] > zt) {
          S->indz = i - 1;
          S->indzp = i;
          break;
        }
      }
      for (i = S->indz; i <= S->indzp; i++)
        if (S->indzindicator[i]) fpo_interpol_misslev(S, i);
    }

    /* vertical interpolation, :342-350 */
    dz = 1.f / (height[S->indzp] - height[S->indz]);
    dz1 = (zt - height[S->indz]) * dz;
    dz2 = (height[S->indzp] - zt) * dz;

    S->u = dz1 * S->uprof[S->indzp] + dz2 * S->uprof[S->indz];
    S->v = dz1 * S->vprof[S->indzp] + dz2 * S->vprof[S->indz];
    S->w = dz1 * S->wprof[S->indzp] + dz2 * S->wprof[S->indz];
    rhoa = dz1 * S->rhoprof[S->indzp] + dz2 * S->rhoprof[S->indz];
    rhograd = dz1 * S->rhogradprof[S->indzp] + dz2 * S->rhogradprof[S->indz];

    if (c->turbswitch)
      fpo_hanna(S, zt);
    else
      fpo_hanna1(S, zt);

    /* horizontal components, :371-384 */
    if (nrand + 1 > maxrand) nrand = 1;
    if (dt / S->tlu < .5f) {
      up = (1.f - dt / S->tlu) * up + rn[nrand] * S->sigu * fpo_sqrtf(2.f * dt / S->tlu);
    } else {
      ru = fpo_expf(-dt / S->tlu);
      up = ru * up + rn[nrand] * S->sigu * fpo_sqrtf(1.f - ru * ru);
    }
    if (dt / S->tlv < .5f) {
      vp = (1.f - dt / S->tlv) * vp + rn[nrand + 1] * S->sigv * fpo_sqrtf(2.f * dt / S->tlv);
    } else {
      rv = fpo_expf(-dt / S->tlv);
      vp = rv * vp + rn[nrand + 1] * S->sigv * fpo_sqrtf(1.f - rv * rv);
    }
    nrand = nrand + 2;

    if (nrand + c->ifine > maxrand) nrand = 1;
    rhoaux = rhograd / rhoa;
    dtf = dt * c->fine;
    dtftlw = dtf / S->tlw;

    /* ifine short steps for the vertical component, :396-498 */
    for (i = 1; i <= c->ifine; i++) {
      if (c->turbswitch) {
        if (dtftlw < .5f) {
          if (c->cblflag == 1) {
            if (-S->h / S->ol > 5.f) {
              flagrein = 0;
              nrand = nrand + 1;
              old_wp_buf = wp;
              fpo_cbl(S, wp, zt, S->ust, S->wst, S->h, rhoa, rhograd, S->sigw,
                      S->dsigwdz, S->tlw, &ptot_lhh, &Q_lhh, &phi_lhh, &ath, &bth,
                      S->ol, &flagrein);
              wp = (wp + ath * dtf + bth * rn[nrand] * fpo_sqrtf(dtf)) * (float)icbt;
              delz = wp * dtf;
              if (flagrein == 1) {
                fpo_re_initialize_particle(S, zt, S->ust, S->wst, S->h, S->sigw,
                                           &old_wp_buf, &nrand, S->ol);
                wp = old_wp_buf;
                delz = wp * dtf;
                S->nan_count++;
              }
            } else {
              nrand = nrand + 1;
              old_wp_buf = wp;
              ath = -wp / S->tlw + S->sigw * S->dsigwdz +
                    wp * wp / S->sigw * S->dsigwdz +
                    S->sigw * S->sigw / rhoa * rhograd;
              bth = S->sigw * rn[nrand] * fpo_sqrtf(2.f * dtftlw);
              wp = (wp + ath * dtf + bth) * (float)icbt;
              delz = wp * dtf;
              del_test = (1.f - wp) / wp;
              if (isnan(wp) || isnan(del_test)) {
                nrand = nrand + 1;
                wp = S->sigw * rn[nrand];
                delz = wp * dtf;
                S->nan_count2++;
              }
            }
          } else {
            wp = ((1.f - dtftlw) * wp + rn[nrand + i] * fpo_sqrtf(2.f * dtftlw) +
                  dtf * (S->dsigwdz + rhoaux * S->sigw)) *
                 (float)icbt;
            delz = wp * S->sigw * dtf;
          }
        } else {
          rw = fpo_expf(-dtftlw);
          wp = (rw * wp + rn[nrand + i] * fpo_sqrtf(1.f - rw * rw) +
                S->tlw * (1.f - rw) * (S->dsigwdz + rhoaux * S->sigw)) *
               (float)icbt;
          delz = wp * S->sigw * dtf;
        }
      } else {
        rw = fpo_expf(-dtftlw);
        wp = (rw * wp + rn[nrand + i] * fpo_sqrtf(1.f - rw * rw) * S->sigw +
              S->tlw * (1.f - rw) * (S->dsigw2dz + rhoaux * (S->sigw * S->sigw))) *
             (float)icbt;
        delz = wp * dtf;
      }

      if (c->turboff) {
        up = 0.0f;
        vp = 0.0f;
        wp = 0.0f;
        delz = 0.f;
      }

      if (fabsf(delz) > S->h) delz = fmodf(delz, S->h);

      if (delz < -zt) { /* reflection at ground */
        icbt = -1;
        zt = -zt - delz;
      } else if (delz > (S->h - zt)) { /* reflection at h */
        icbt = -1;
        zt = -zt - delz + 2.f * S->h;
      } else {
        icbt = 1;
        zt = zt + delz;
      }

      if (i != c->ifine) {
        S->zeta = zt / S->h;
        fpo_hanna_short(S, zt);
      }
    }
    /* after a Fortran DO the index is ifine+1 */
    if (c->cblflag != 1) nrand = nrand + (c->ifine + 1);

    /* time step for the next integration, :504-510 */
    if (c->turbswitch) {
      float t = fpo_minf(S->tlw, S->h / fpo_maxf(2.f * fabsf(wp * S->sigw), 1.e-5f));
      t = fpo_minf(t, 0.5f / fabsf(S->dsigwdz));
      ldt = fpo_int_f(t * c->ctl);
    } else {
      float t = fpo_minf(S->tlw, S->h / fpo_maxf(2.f * fabsf(wp), 1.e-5f));
      ldt = fpo_int_f(t * c->ctl);
    }
    ldt = (ldt > c->mintime) ? ldt : c->mintime;

    add_settling(S, itime, nrelpoint, xt, yt, zt);

    /* accumulate displacements, :539-547 */
    dxsave = dxsave + S->u * dt;
    dysave = dysave + S->v * dt;
    dawsave = dawsave + up * dt;
    dcwsave = dcwsave + vp * dt;
    zt = zt + S->w * dt * (float)c->ldirect;

    if (zt >= height[nz]) zt = height[nz] - 100.f * eps;

    if (zt > S->h) {
      if (itimec == itime + c->lsynctime) {
        if (!S->strict_reference) {
          S->usig = 0.5f * (S->usigprof[S->indzp] + S->usigprof[S->indz]);
          S->vsig = 0.5f * (S->vsigprof[S->indzp] + S->vsigprof[S->indz]);
          S->wsig = 0.5f * (S->wsigprof[S->indzp] + S->wsigprof[S->indz]);
        }
        goto L99;
      }
      goto L700;
    }

    /* probability of deposition, :582-599 */
    if (c->drydep && (zt < 2.f * href)) {
      for (ks = 1; ks <= c->nspec; ks++) {
        if (c->drydepspec[ks - 1]) {
          if (S->depoindicator[ks]) fpo_interpol_vdep(S, ks, &vdepo[ks]);
          prob[ks - 1] = 1.f + (prob[ks - 1] - 1.f) *
                                   fpo_expf(-vdepo[ks] * fabsf(dt) / (2.f * href));
        }
      }
    }

    if (zt < 0.f) zt = fpo_minf(S->h - eps2, -1.f * zt);

    if (itimec == (itime + c->lsynctime)) {
      S->usig = 0.5f * (S->usigprof[S->indzp] + S->usigprof[S->indz]);
      S->vsig = 0.5f * (S->vsigprof[S->indzp] + S->vsigprof[S->indz]);
      S->wsig = 0.5f * (S->wsigprof[S->indzp] + S->wsigprof[S->indz]);
      goto L99;
    }
    goto L100;
  }

  /* above the PBL: one step, :629-708 */
L700:
  if (S->ngrid <= 0) { /* :630-636 */
    xts = (float)xt;
    yts = (float)yt;
    fpo_interpol_wind(S, itime, xts, yts, zt);
  } else {
    fpo_interpol_wind(S, itime, xtn, ytn, zt);
  }

  ldt = abs(c->lsynctime - itimec + itime);
  dt = (float)ldt;

  if (zt < tropop) {
    uxscale = fpo_sqrtf(2.f * c->d_trop / dt);
    if (nrand + 1 > maxrand) nrand = 1;
    ux = rn[nrand] * uxscale;
    vy = rn[nrand + 1] * uxscale;
    nrand = nrand + 2;
    wp = 0.f;
  } else if (zt < tropop + 1000.f) {
    weight = (zt - tropop) / 1000.f;
    uxscale = fpo_sqrtf(2.f * c->d_trop / dt * (1.f - weight));
    if (nrand + 2 > maxrand) nrand = 1;
    ux = rn[nrand] * uxscale;
    vy = rn[nrand + 1] * uxscale;
    wpscale = fpo_sqrtf(2.f * c->d_strat / dt * weight);
    wp = rn[nrand + 2] * wpscale + c->d_strat / 1000.f;
    nrand = nrand + 3;
  } else {
    if (nrand > maxrand) nrand = 1;
    ux = 0.f;
    vy = 0.f;
    wpscale = fpo_sqrtf(2.f * c->d_strat / dt);
    wp = rn[nrand] * wpscale;
    nrand = nrand + 1;
  }

  if (c->turboff) {
    ux = 0.0f;
    vy = 0.0f;
    wp = 0.0f;
  }

  add_settling(S, itime, nrelpoint, xt, yt, zt);

  dxsave = dxsave + (S->u + ux) * dt;
  dysave = dysave + (S->v + vy) * dt;
  zt = zt + (S->w + wp) * dt * (float)c->ldirect;
  if (zt < 0.f) zt = fpo_minf(S->h - eps2, -1.f * zt);

L99:
  /* mesoscale fluctuations, :728-739 */
  r = fpo_expf(-2.f * (float)abs(c->lsynctime) / (float)S->lwindinterv);
  rs = fpo_sqrtf(1.f - r * r);
  if (nrand + 2 > maxrand) nrand = 1;
  usigold = r * usigold + rs * rn[nrand] * S->usig * c->turbmesoscale;
  vsigold = r * vsigold + rs * rn[nrand + 1] * S->vsig * c->turbmesoscale;
  wsigold = r * wsigold + rs * rn[nrand + 2] * S->wsig * c->turbmesoscale;

  dxsave = dxsave + usigold * (float)c->lsynctime;
  dysave = dysave + vsigold * (float)c->lsynctime;

  zt = zt + wsigold * (float)c->lsynctime;
  if (zt < 0.f) zt = -1.f * zt;

  /* along/cross wind -> x,y; new position, :747-778 */
  fpo_windalign(dxsave, dysave, dawsave, dcwsave, &ux, &vy);
  dxsave = dxsave + ux;
  dysave = dysave + vy;
  move_horizontal(S, &xt, &yt, dxsave, dysave, (float)c->ldirect);

  if (wrap_and_check(S, &xt, &yt)) {
    *nstop = 3;
    goto Lout;
  }

  if (zt >= height[nz]) zt = height[nz] - 100.f * eps;

  /* Petterssen corrector, :829-985 */
  if (ldt != abs(c->lsynctime)) goto Lout;
  if (abs(itime + ldt * c->ldirect) > abs(S->memtime[2])) goto Lout;

  ngr = choose_grid(S, xt, yt); /* :841-857 */
  if (ngr != S->ngrid) goto Lout;

  if (S->ngrid > 0) { /* :862-870 */
    xtn = (float)((xt - c->xln[S->ngrid - 1]) * c->xresoln[S->ngrid - 1]);
    ytn = (float)((yt - c->yln[S->ngrid - 1]) * c->yresoln[S->ngrid - 1]);
    S->ix = fpo_int_f(xtn);
    S->jy = fpo_int_f(ytn);
  } else {
    S->ix = fpo_int_d(xt);
    S->jy = fpo_int_d(yt);
  }
  S->ixp = S->ix + 1;
  S->jyp = S->jy + 1;
  if (!S->strict_reference && S->jyp >= c->nymax) S->jyp = S->jyp - 1;

  uold = S->u;
  vold = S->v;
  wold = S->w;

  if (S->ngrid <= 0) { /* :885-891 */
    xts = (float)xt;
    yts = (float)yt;
    fpo_interpol_wind_short(S, itime + ldt * c->ldirect, xts, yts, zt);
  } else {
    fpo_interpol_wind_short(S, itime + ldt * c->ldirect, xtn, ytn, zt);
  }
  did_pett = 1;

  add_settling(S, itime + ldt, nrelpoint, xt, yt, zt);

  S->u = (S->u - uold) / 2.f;
  S->v = (S->v - vold) / 2.f;
  S->w = (S->w - wold) / 2.f;

  zt = zt + S->w * (float)(ldt * c->ldirect);
  if (zt < 0.f) zt = fpo_minf(S->h - eps2, -1.f * zt);
  move_horizontal(S, &xt, &yt, S->u, S->v, (float)(ldt * c->ldirect));
  /* (the polar branch divides u,v by gridsize in place; they are dead after) */

  if (wrap_and_check(S, &xt, &yt)) {
    *nstop = 3;
    goto Lout;
  }

  if (zt >= height[nz]) zt = height[nz] - 100.f * eps;

Lout:
  *xt_io = xt;
  *yt_io = yt;
  *zt_io = zt;
  *up_io = up;
  *vp_io = vp;
  *wp_io = wp;
  *usigold_io = usigold;
  *vsigold_io = vsigold;
  *wsigold_io = wsigold;
  *ldt_io = ldt;
  *icbt_io = (int16_t)icbt;
  S->last.n_substeps += nsub;
  S->last.n_pbl += took_pbl;
  S->last.n_petterssen += did_pett;
}
