/*
 * fpo_step.c -- oracle restatement of timemanager's particle loop, conccalc,
 * drydepokernel(_nest) and the state plumbing (test infrastructure).
 *   particle loop : src/timemanager.f90:531-712
 *   conccalc      : src/conccalc.f90:50-498
 *   drydepokernel : src/drydepokernel.f90:41-116, drydepokernel_nest.f90
 *   grid layout   : src/outgrid_init.f90:192-201
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "fpo.h"
#include "fpo_math.h"

#define MAXRECEPTOR FPB_MAXRECEPTOR

/* gridunc(0:nxg-1,0:nyg-1,numzgrid,maxspec,maxpointspec_act,nclassunc,maxageclass) */
static inline size_t gidx(const fpb_config *c, int nxg, int nyg, int ix, int jy,
                          int kz, int ks, int kp, int nc, int na) {
  size_t i = (size_t)(na - 1);
  i = i * c->nclassunc + (nc - 1);
  i = i * c->maxpointspec_act + (kp - 1);
  i = i * c->maxspec + (ks - 1);
  i = i * c->numzgrid + (kz - 1);
  i = i * nyg + jy;
  i = i * nxg + ix;
  return i;
}
/* drygridunc(0:nxg-1,0:nyg-1,maxspec,maxpointspec_act,nclassunc,maxageclass) */
static inline size_t didx(const fpb_config *c, int nxg, int nyg, int ix, int jy,
                          int ks, int kp, int nc, int na) {
  size_t i = (size_t)(na - 1);
  i = i * c->nclassunc + (nc - 1);
  i = i * c->maxpointspec_act + (kp - 1);
  i = i * c->maxspec + (ks - 1);
  i = i * nyg + jy;
  i = i * nxg + ix;
  return i;
}
static size_t gsize(const fpb_config *c, int nxg, int nyg) {
  return (size_t)nxg * nyg * c->numzgrid * c->maxspec * c->maxpointspec_act *
         c->nclassunc * c->maxageclass;
}
static size_t dsize(const fpb_config *c, int nxg, int nyg) {
  return (size_t)nxg * nyg * c->maxspec * c->maxpointspec_act * c->nclassunc *
         c->maxageclass;
}

#define XM1(S, j, ks) (S)->xmass1[(size_t)(j) + (size_t)((S)->maxpart + 1) * ((ks)-1)]
#define XSC(S, j, ks) (S)->xscav_frac1[(size_t)(j) + (size_t)((S)->maxpart + 1) * ((ks)-1)]

fpo_state *fpo_create(const fpb_config *cfg, int strict_reference) {
  fpo_state *S = (fpo_state *)calloc(1, sizeof(fpo_state));
  S->c = *cfg;
  const fpb_config *c = &S->c;
  S->strict_reference = strict_reference;
  S->height = (float *)calloc((size_t)c->nzmax + 2, sizeof(float));
  memcpy(S->height + 1, cfg->height, (size_t)c->nz * sizeof(float));
  int np = c->numpoint > 0 ? c->numpoint : 1;
  S->npart = (int32_t *)calloc((size_t)np + 1, sizeof(int32_t));
  S->xmass = (float *)calloc((size_t)np * c->maxspec + 1, sizeof(float));
  if (cfg->npart) memcpy(S->npart + 1, cfg->npart, (size_t)c->numpoint * sizeof(int32_t));
  if (cfg->xmass) memcpy(S->xmass, cfg->xmass, (size_t)c->numpoint * c->maxspec * sizeof(float));
  S->c.height = NULL;
  S->c.npart = NULL;
  S->c.xmass = NULL;
  S->memind[1] = 1;
  S->memind[2] = 2;
  S->idummy_advance = -7;
  S->idummy_initialize = -7;
  S->idummy_release = -7;
  S->idummy_domainfill = -11;
  S->bc.idummy = -11;
  S->maxpart = c->maxpart;
  size_t n = (size_t)c->maxpart + 1;
  S->xtra1 = (double *)calloc(n, sizeof(double));
  S->ytra1 = (double *)calloc(n, sizeof(double));
  S->ztra1 = (float *)calloc(n, sizeof(float));
  S->itra1 = (int32_t *)calloc(n, sizeof(int32_t));
  S->npoint = (int32_t *)calloc(n, sizeof(int32_t));
  S->nclass = (int32_t *)calloc(n, sizeof(int32_t));
  S->idt = (int32_t *)calloc(n, sizeof(int32_t));
  S->itramem = (int32_t *)calloc(n, sizeof(int32_t));
  S->itrasplit = (int32_t *)calloc(n, sizeof(int32_t));
  S->uap = (float *)calloc(n, sizeof(float));
  S->ucp = (float *)calloc(n, sizeof(float));
  S->uzp = (float *)calloc(n, sizeof(float));
  S->us = (float *)calloc(n, sizeof(float));
  S->vs = (float *)calloc(n, sizeof(float));
  S->ws = (float *)calloc(n, sizeof(float));
  S->cbt = (int16_t *)calloc(n, sizeof(int16_t));
  S->xmass1 = (float *)calloc(n * c->maxspec, sizeof(float));
  S->xscav_frac1 = (float *)calloc(n * c->maxspec, sizeof(float));
  S->trace_nsub = (int32_t *)calloc(n, sizeof(int32_t));
  for (size_t j = 0; j < n; j++) S->itra1[j] = FPB_ITRA_DEAD; /* src/FLEXPART.f90:315-317 */
  size_t nzp = (size_t)c->nzmax + 2;
  S->uprof = (float *)calloc(nzp, sizeof(float));
  S->vprof = (float *)calloc(nzp, sizeof(float));
  S->wprof = (float *)calloc(nzp, sizeof(float));
  S->usigprof = (float *)calloc(nzp, sizeof(float));
  S->vsigprof = (float *)calloc(nzp, sizeof(float));
  S->wsigprof = (float *)calloc(nzp, sizeof(float));
  S->rhoprof = (float *)calloc(nzp, sizeof(float));
  S->rhogradprof = (float *)calloc(nzp, sizeof(float));
  S->indzindicator = (unsigned char *)calloc(nzp, 1);
  /* outgrid_init.f90:192-201, 305-338: allocate + zero */
  S->gridunc = (float *)calloc(gsize(c, c->numxgrid, c->numygrid) + 1, sizeof(float));
  S->drygridunc = (float *)calloc(dsize(c, c->numxgrid, c->numygrid) + 1, sizeof(float));
  S->wetgridunc = (float *)calloc(dsize(c, c->numxgrid, c->numygrid) + 1, sizeof(float));
  if (c->nested_output == 1) {
    S->griduncn = (float *)calloc(gsize(c, c->numxgridn, c->numygridn) + 1, sizeof(float));
    S->drygriduncn = (float *)calloc(dsize(c, c->numxgridn, c->numygridn) + 1, sizeof(float));
    S->wetgriduncn = (float *)calloc(dsize(c, c->numxgridn, c->numygridn) + 1, sizeof(float));
  }
  S->creceptor = (float *)calloc((size_t)MAXRECEPTOR * c->maxspec + 1, sizeof(float));
  return S;
}

void fpo_destroy(fpo_state *S) {
  if (!S) return;
  free(S->height); free(S->npart); free(S->xmass); free(S->rannumb);
  free(S->xtra1); free(S->ytra1); free(S->ztra1); free(S->itra1);
  free(S->npoint); free(S->nclass); free(S->idt); free(S->itramem);
  free(S->itrasplit); free(S->uap); free(S->ucp); free(S->uzp); free(S->us);
  free(S->vs); free(S->ws); free(S->cbt); free(S->xmass1);
  free(S->xscav_frac1); free(S->trace_nsub);
  free(S->uprof); free(S->vprof); free(S->wprof); free(S->usigprof);
  free(S->vsigprof); free(S->wsigprof); free(S->rhoprof); free(S->rhogradprof);
  free(S->indzindicator);
  free(S->gridunc); free(S->griduncn); free(S->drygridunc);
  free(S->drygriduncn); free(S->creceptor);
  free(S->wetgridunc); free(S->wetgriduncn);
  free(S->index_queue); free(S->zpoint1); free(S->zpoint2);
  free(S->bc.numcolumn_we); free(S->bc.numcolumn_sn); free(S->bc.zcolumn_we); free(S->bc.zcolumn_sn);
  free(S->bc.acc_mass_we); free(S->bc.acc_mass_sn);
  free(S);
}

void fpo_set_release_heights(fpo_state *S, const float *zpoint1, const float *zpoint2, int numpoint) {
  free(S->zpoint1); free(S->zpoint2);
  S->zpoint1 = (float *)calloc((size_t)numpoint + 1, sizeof(float));
  S->zpoint2 = (float *)calloc((size_t)numpoint + 1, sizeof(float));
  for (int i = 0; i < numpoint; i++) {
    S->zpoint1[i + 1] = zpoint1[i];
    S->zpoint2[i + 1] = zpoint2[i];
  }
}

void fpo_set_met(fpo_state *S, int slot, const fpb_met_ptrs *m) { S->met[slot] = *m; }
void fpo_set_met_nest(fpo_state *S, int slot, int nest, const fpb_met_ptrs *m) {
  if (nest >= 1 && nest <= FPB_MAXNESTS) S->metn[nest][slot] = *m;
}

void fpo_set_met_bracket(fpo_state *S, const int memind[2], const int memtime[2],
                         int lwindinterv) {
  S->memind[1] = memind[0];
  S->memind[2] = memind[1];
  S->memtime[1] = memtime[0];
  S->memtime[2] = memtime[1];
  S->lwindinterv = lwindinterv;
}

void fpo_set_numpart(fpo_state *S, int numpart) { S->numpart = numpart; }
int fpo_numpart(const fpo_state *S) { return S->numpart; }
int fpo_numparticlecount(const fpo_state *S) { return S->numparticlecount; }

/* sub-steps each particle took in the last fpo_step (1-based like the particle arrays) */
const int32_t *fpo_trace_nsub(const fpo_state *S) { return S->trace_nsub; }

void fpo_push_particles(fpo_state *S, int first, int count,
                        const fpb_particle_ptrs *p) {
  for (int i = 0; i < count; i++) {
    int s = first + i, j = s + 1;
    S->xtra1[j] = p->xtra1[s];
    S->ytra1[j] = p->ytra1[s];
    S->ztra1[j] = p->ztra1[s];
    S->itra1[j] = p->itra1[s];
    S->npoint[j] = p->npoint[s];
    S->nclass[j] = p->nclass[s];
    S->idt[j] = p->idt[s];
    S->itramem[j] = p->itramem[s];
    S->itrasplit[j] = p->itrasplit ? p->itrasplit[s] : 0;
    S->uap[j] = p->uap[s];
    S->ucp[j] = p->ucp[s];
    S->uzp[j] = p->uzp[s];
    S->us[j] = p->us[s];
    S->vs[j] = p->vs[s];
    S->ws[j] = p->ws[s];
    S->cbt[j] = p->cbt[s];
    for (int ks = 1; ks <= S->c.nspec; ks++) {
      XM1(S, j, ks) = p->xmass1[(size_t)s + (size_t)p->ld * (ks - 1)];
      if (p->xscav_frac1) XSC(S, j, ks) = p->xscav_frac1[(size_t)s + (size_t)p->ld * (ks - 1)];
    }
  }
  if (first + count > S->numpart) S->numpart = first + count;
}

void fpo_pull_particles(fpo_state *S, int first, int count,
                        const fpb_particle_ptrs *p) {
  for (int i = 0; i < count; i++) {
    int s = first + i, j = s + 1;
    if (p->xtra1) p->xtra1[s] = S->xtra1[j];
    if (p->ytra1) p->ytra1[s] = S->ytra1[j];
    if (p->ztra1) p->ztra1[s] = S->ztra1[j];
    if (p->itra1) p->itra1[s] = S->itra1[j];
    if (p->npoint) p->npoint[s] = S->npoint[j];
    if (p->nclass) p->nclass[s] = S->nclass[j];
    if (p->idt) p->idt[s] = S->idt[j];
    if (p->itramem) p->itramem[s] = S->itramem[j];
    if (p->itrasplit) p->itrasplit[s] = S->itrasplit[j];
    if (p->uap) p->uap[s] = S->uap[j];
    if (p->ucp) p->ucp[s] = S->ucp[j];
    if (p->uzp) p->uzp[s] = S->uzp[j];
    if (p->us) p->us[s] = S->us[j];
    if (p->vs) p->vs[s] = S->vs[j];
    if (p->ws) p->ws[s] = S->ws[j];
    if (p->cbt) p->cbt[s] = S->cbt[j];
    for (int ks = 1; ks <= S->c.nspec; ks++) {
      if (p->xmass1) p->xmass1[(size_t)s + (size_t)p->ld * (ks - 1)] = XM1(S, j, ks);
      if (p->xscav_frac1) p->xscav_frac1[(size_t)s + (size_t)p->ld * (ks - 1)] = XSC(S, j, ks);
    }
  }
}

/* src/drydepokernel.f90:41-116 (nest = 0) / drydepokernel_nest.f90 (nest = 1) */
static void drydepo_common(fpo_state *S, int nunc, const float *deposit, float x,
                           float y, int nage, int kp, int nest) {
  const fpb_config *c = &S->c;
  const int nxg = nest ? c->numxgridn : c->numxgrid;
  const int nyg = nest ? c->numygridn : c->numygrid;
  float *grid = nest ? S->drygriduncn : S->drygridunc;
  float xl, yl, ddx, ddy, wx, wy, w;
  int ix, jy, ixp, jyp;
  if (nest) {
    xl = (x * c->dx + c->xoutshiftn) / c->dxoutn;
    yl = (y * c->dy + c->youtshiftn) / c->dyoutn;
  } else {
    xl = (x * c->dx + c->xoutshift) / c->dxout;
    yl = (y * c->dy + c->youtshift) / c->dyout;
  }
  ix = fpo_int_f(xl);
  jy = fpo_int_f(yl);
  ddx = xl - (float)ix;
  ddy = yl - (float)jy;
  if (ddx > 0.5f) {
    ixp = ix + 1;
    wx = 1.5f - ddx;
  } else {
    ixp = ix - 1;
    wx = 0.5f + ddx;
  }
  if (ddy > 0.5f) {
    jyp = jy + 1;
    wy = 1.5f - ddy;
  } else {
    jyp = jy - 1;
    wy = 0.5f + ddy;
  }
  if (!nest && !c->lusekerneloutput) {
    for (int ks = 1; ks <= c->nspec; ks++)
      if ((fabsf(deposit[ks - 1]) > 0.f) && c->drydepspec[ks - 1])
        if ((ix >= 0) && (jy >= 0) && (ix <= nxg - 1) && (jy <= nyg - 1))
          grid[didx(c, nxg, nyg, ix, jy, ks, kp, nunc, nage)] += deposit[ks - 1];
    return;
  }
  for (int ks = 1; ks <= c->nspec; ks++) {
    if ((fabsf(deposit[ks - 1]) > 0.f) && c->drydepspec[ks - 1]) {
      if ((ix >= 0) && (jy >= 0) && (ix <= nxg - 1) && (jy <= nyg - 1)) {
        w = wx * wy;
        grid[didx(c, nxg, nyg, ix, jy, ks, kp, nunc, nage)] += deposit[ks - 1] * w;
      }
      if ((ixp >= 0) && (jyp >= 0) && (ixp <= nxg - 1) && (jyp <= nyg - 1)) {
        w = (1.f - wx) * (1.f - wy);
        grid[didx(c, nxg, nyg, ixp, jyp, ks, kp, nunc, nage)] += deposit[ks - 1] * w;
      }
      if ((ixp >= 0) && (jy >= 0) && (ixp <= nxg - 1) && (jy <= nyg - 1)) {
        w = (1.f - wx) * wy;
        grid[didx(c, nxg, nyg, ixp, jy, ks, kp, nunc, nage)] += deposit[ks - 1] * w;
      }
      if ((ix >= 0) && (jyp >= 0) && (ix <= nxg - 1) && (jyp <= nyg - 1)) {
        w = wx * (1.f - wy);
        grid[didx(c, nxg, nyg, ix, jyp, ks, kp, nunc, nage)] += deposit[ks - 1] * w;
      }
    }
  }
}

void fpo_drydepokernel(fpo_state *S, int nunc, const float *deposit, float x,
                       float y, int nage, int kp) {
  drydepo_common(S, nunc, deposit, x, y, nage, kp, 0);
}
void fpo_drydepokernel_nest(fpo_state *S, int nunc, const float *deposit,
                            float x, float y, int nage, int kp) {
  drydepo_common(S, nunc, deposit, x, y, nage, kp, 1);
}

/* src/timemanager.f90:531-712 */
void fpo_step(fpo_state *S, int itime, int ldeltat, fpb_step_stats *stats) {
  const fpb_config *c = &S->c;
  const float minmass = 0.0001f; /* par_mod.f90:214 */
  float prob[FPB_MAXSPEC], drydeposit[FPB_MAXSPEC], decfact, xmassfract;
  int nstop, kp, nage, itage;
  memset(&S->last, 0, sizeof(S->last));
  long nan0 = S->nan_count + S->nan_count2;
  for (int k = 0; k < FPB_MAXSPEC; k++) prob[k] = 0.f, drydeposit[k] = 0.f;

  for (int j = 1; j <= S->numpart; j++) {
    if (S->itra1[j] != itime) continue;
    S->last.n_active++;

    if (c->ioutputforeachrelease == 1)
      kp = S->npoint[j];
    else
      kp = 1;
    itage = abs(S->itra1[j] - S->itramem[j]);
    for (nage = 1; nage <= c->nageclass; nage++)
      if (itage < c->lage[nage - 1]) break;

    if ((S->itramem[j] == itime) || (itime == 0)) {
      fpo_initialize(S, itime, &S->idt[j], &S->uap[j], &S->ucp[j], &S->uzp[j],
                     &S->us[j], &S->vs[j], &S->ws[j], S->xtra1[j], S->ytra1[j],
                     S->ztra1[j], &S->cbt[j]);
      S->last.n_init++;
    }

    /* RECEPTOR: dry/wet depovel, src/timemanager.f90:563-598 -- once after release (xscav_frac1 was
     * initialised negative), before the particle is moved */
    if (c->drybkdep) {
      for (int ks = 1; ks <= c->nspec; ks++)
        if (XSC(S, j, ks) < 0.f) {
          float prob_rec[FPB_MAXSPEC];
          fpo_get_vdep_prob(S, itime, S->xtra1[j], S->ytra1[j], S->ztra1[j], prob_rec);
          if (c->drydepspec[ks - 1]) {
            XSC(S, j, ks) = prob_rec[ks - 1];
          } else {
            XM1(S, j, ks) = 0.f;
            XSC(S, j, ks) = 0.f;
          }
        }
    }
    if (c->wetbkdep) {
      for (int ks = 1; ks <= c->nspec; ks++)
        if (XSC(S, j, ks) < 0.f) {
          float grfraction1 = 0.f;
          const float wetscav = fpo_get_wetscav(S, itime, c->lsynctime, j, ks, &grfraction1);
          if (wetscav > 0.f) {
            const int np = S->npoint[j];
            XSC(S, j, ks) = wetscav * (S->zpoint2[np] - S->zpoint1[np]) * grfraction1;
          } else {
            XM1(S, j, ks) = 0.f;
            XSC(S, j, ks) = 0.f;
          }
        }
    }

    long nsub0 = S->last.n_substeps;
    fpo_advance(S, itime, S->npoint[j], &S->idt[j], &S->uap[j], &S->ucp[j],
                &S->uzp[j], &S->us[j], &S->vs[j], &S->ws[j], &nstop,
                &S->xtra1[j], &S->ytra1[j], &S->ztra1[j], prob, &S->cbt[j]);
    if (S->trace_nsub) S->trace_nsub[j] = (int32_t)(S->last.n_substeps - nsub0);

    /* not in the reference (it carries a NaN particle on and indexes out of bounds with it):
     * a position that is not finite terminates the particle, counted in n_nonfinite */
    if (!(isfinite(S->xtra1[j]) && isfinite(S->ytra1[j]) && isfinite(S->ztra1[j]))) {
      nstop = 4;
      S->last.n_nonfinite++;
    }
    if (nstop > 1) {
      S->itra1[j] = FPB_ITRA_DEAD;
      S->last.n_terminated++;
    } else {
      S->itra1[j] = itime + c->lsynctime;

      xmassfract = 0.f;
      for (int ks = 1; ks <= c->nspec; ks++) {
        if (c->decay[ks - 1] > 0.f)
          decfact = fpo_expf(-(float)abs(c->lsynctime) * c->decay[ks - 1]);
        else
          decfact = 1.f;

        if (c->drydepspec[ks - 1]) {
          drydeposit[ks - 1] = XM1(S, j, ks) * prob[ks - 1] * decfact;
          XM1(S, j, ks) = XM1(S, j, ks) * (1.f - prob[ks - 1]) * decfact;
          if (c->decay[ks - 1] > 0.f)
            drydeposit[ks - 1] = drydeposit[ks - 1] *
                                 fpo_expf((float)abs(ldeltat) * c->decay[ks - 1]);
        } else {
          XM1(S, j, ks) = XM1(S, j, ks) * decfact;
        }

        if (c->mdomainfill == 0 && c->mquasilag == 0) {
          float xm = S->xmass[(S->npoint[j] - 1) + (size_t)c->numpoint * (ks - 1)];
          if (xm > 0.f)
            xmassfract = fpo_maxf(xmassfract,
                                  (float)S->npart[S->npoint[j]] * XM1(S, j, ks) / xm);
        } else {
          xmassfract = 1.0f;
        }
      }

      int dead = 0;
      if (xmassfract < minmass) {
        S->itra1[j] = FPB_ITRA_DEAD;
        dead = 1;
      }

      if (c->drydep && (c->ldirect == 1)) {
        fpo_drydepokernel(S, S->nclass[j], drydeposit, (float)S->xtra1[j],
                          (float)S->ytra1[j], nage, kp);
        if (c->nested_output == 1)
          fpo_drydepokernel_nest(S, S->nclass[j], drydeposit, (float)S->xtra1[j],
                                 (float)S->ytra1[j], nage, kp);
      }

      if (abs(S->itra1[j] - S->itramem[j]) >= c->lage[c->nageclass - 1]) {
        S->itra1[j] = FPB_ITRA_DEAD;
        dead = 1;
      }
      S->last.n_terminated += dead;
    }
  }
  S->last.n_nan_cbl = (S->nan_count + S->nan_count2) - nan0;
  if (stats) *stats = S->last;
}

/* src/conccalc.f90:50-498 */
void fpo_conccalc(fpo_state *S, int itime, float weight) {
  const fpb_config *c = &S->c;
  int itage, ix, jy, ixp, jyp, kz, nage, indz = 1, indzp = 2, nrelpointer;
  float rddx, rddy, p1, p2, p3, p4, dz1, dz2, dz;
  float hx, hy, hz, h, xd, yd, zd, xkern, r2, cc[FPB_MAXSPEC], ddx, ddy;
  float rhoprof[3], rhoi = 1.f;
  float xl, yl, wx, wy, w;
  const float factor = .596831f, hxmax = 6.0f, hymax = 4.0f, hzmax = 150.f;
  const int bk = c->drybkdep || c->wetbkdep;

  for (int i = 1; i <= S->numpart; i++) {
    if (S->itra1[i] != itime) continue;

    itage = abs(S->itra1[i] - S->itramem[i]);
    for (nage = 1; nage <= c->nageclass; nage++)
      if (itage < c->lage[nage - 1]) break;

    if (c->ind_samp == -1) {
      ix = fpo_int_d(S->xtra1[i]);
      jy = fpo_int_d(S->ytra1[i]);
      ixp = ix + 1;
      jyp = jy + 1;
      ddx = (float)(S->xtra1[i] - (float)ix);
      ddy = (float)(S->ytra1[i] - (float)jy);
      rddx = 1.f - ddx;
      rddy = 1.f - ddy;
      p1 = rddx * rddy;
      p2 = ddx * rddy;
      p3 = rddx * ddy;
      p4 = ddx * ddy;
      if (jyp >= c->nymax) jyp = jyp - 1;
      for (int il = 2; il <= c->nz; il++)
        if (S->height[il] > S->ztra1[i]) {
          indz = il - 1;
          indzp = il;
          break;
        }
      dz1 = S->ztra1[i] - S->height[indz];
      dz2 = S->height[indzp] - S->ztra1[i];
      dz = 1.f / (dz1 + dz2);
      /* density from "the 2nd wind field": the serial routine mixes
       * memind(2) and the literal slot 2 (conccalc.f90:117-120); the MPI
       * routine uses memind(2) throughout (conccalc_mpi.f90:124-132).
       * strict_reference follows the serial text. */
      for (int ind = indz; ind <= indzp; ind++) {
        const float *r1 = S->met[S->memind[2]].rho;
        const float *r2p = S->strict_reference ? S->met[2].rho : r1;
        size_t st = (size_t)c->nxmax * c->nymax * (size_t)(ind - 1);
        rhoprof[ind - indz + 1] = p1 * r1[st + ix + (size_t)c->nxmax * jy] +
                                  p2 * r2p[st + ixp + (size_t)c->nxmax * jy] +
                                  p3 * r2p[st + ix + (size_t)c->nxmax * jyp] +
                                  p4 * r2p[st + ixp + (size_t)c->nxmax * jyp];
      }
      rhoi = (dz1 * rhoprof[2] + dz2 * rhoprof[1]) * dz;
    } else if (c->ind_samp == 0) {
      rhoi = 1.f;
    }

    if ((c->ioutputforeachrelease == 0) || (c->mdomainfill == 1))
      nrelpointer = 1;
    else
      nrelpointer = S->npoint[i];

    for (kz = 1; kz <= c->numzgrid; kz++)
      if (c->outheight[kz - 1] > S->ztra1[i]) break;
    if (kz > c->numzgrid) continue;

    for (int nest = 0; nest <= (c->nested_output == 1 ? 1 : 0); nest++) {
      const int nxg = nest ? c->numxgridn : c->numxgrid;
      const int nyg = nest ? c->numygridn : c->numygrid;
      float *grid = nest ? S->griduncn : S->gridunc;
      if (nest) {
        xl = (float)((S->xtra1[i] * c->dx + c->xoutshiftn) / c->dxoutn);
        yl = (float)((S->ytra1[i] * c->dy + c->youtshiftn) / c->dyoutn);
      } else {
        xl = (float)((S->xtra1[i] * c->dx + c->xoutshift) / c->dxout);
        yl = (float)((S->ytra1[i] * c->dy + c->youtshift) / c->dyout);
      }
      ix = fpo_int_f(xl);
      if (xl < 0.f) ix = ix - 1;
      jy = fpo_int_f(yl);
      if (yl < 0.f) jy = jy - 1;

/* the reference multiplies w and weight in a different order from cell to
 * cell in the backward-deposition branch (conccalc.f90:229,243,260,274);
 * WFIRST reproduces that. */
#define ADD(IX, JY, W, USEW, WFIRST)                                            \
  do {                                                                          \
    for (int ks = 1; ks <= c->nspec; ks++) {                                    \
      size_t g = gidx(c, nxg, nyg, (IX), (JY), kz, ks, nrelpointer, S->nclass[i], nage); \
      if (bk) {                                                                 \
        if ((USEW) && (WFIRST) && !nest)                                        \
          grid[g] = grid[g] + XM1(S, i, ks) / rhoi * (W)*weight *               \
                                  fpo_maxf(XSC(S, i, ks), 0.0f);                \
        else if (USEW)                                                          \
          grid[g] = grid[g] + XM1(S, i, ks) / rhoi * weight * (W) *             \
                                  fpo_maxf(XSC(S, i, ks), 0.0f);                \
        else                                                                    \
          grid[g] = grid[g] + XM1(S, i, ks) / rhoi * weight *                   \
                                  fpo_maxf(XSC(S, i, ks), 0.0f);                \
      } else if (!(USEW) && c->lparticlecountoutput) {                          \
        grid[g] = grid[g] + 1.f;                                                \
      } else if (USEW) {                                                        \
        grid[g] = grid[g] + XM1(S, i, ks) / rhoi * weight * (W);                \
      } else {                                                                  \
        grid[g] = grid[g] + XM1(S, i, ks) / rhoi * weight;                      \
      }                                                                         \
    }                                                                           \
  } while (0)

      if ((!c->lusekerneloutput) || (itage < 10800) || (xl < 0.5f) || (yl < 0.5f) ||
          (xl > (float)(nxg - 1) - 0.5f) || (yl > (float)(nyg - 1) - 0.5f)) {
        if ((ix >= 0) && (jy >= 0) && (ix <= nxg - 1) && (jy <= nyg - 1))
          ADD(ix, jy, 1.f, 0, 0);
      } else {
        ddx = xl - (float)ix;
        ddy = yl - (float)jy;
        if (ddx > 0.5f) {
          ixp = ix + 1;
          wx = 1.5f - ddx;
        } else {
          ixp = ix - 1;
          wx = 0.5f + ddx;
        }
        if (ddy > 0.5f) {
          jyp = jy + 1;
          wy = 1.5f - ddy;
        } else {
          jyp = jy - 1;
          wy = 0.5f + ddy;
        }
        if ((ix >= 0) && (ix <= nxg - 1)) {
          if ((jy >= 0) && (jy <= nyg - 1)) {
            w = wx * wy;
            ADD(ix, jy, w, 1, 1);
          }
          if ((jyp >= 0) && (jyp <= nyg - 1)) {
            w = wx * (1.f - wy);
            ADD(ix, jyp, w, 1, 0);
          }
        }
        if ((ixp >= 0) && (ixp <= nxg - 1)) {
          if ((jyp >= 0) && (jyp <= nyg - 1)) {
            w = (1.f - wx) * (1.f - wy);
            ADD(ixp, jyp, w, 1, 1);
          }
          if ((jy >= 0) && (jy <= nyg - 1)) {
            w = (1.f - wx) * wy;
            ADD(ixp, jy, w, 1, 0);
          }
        }
      }
#undef ADD
    }
  }

  /* receptor concentrations, src/conccalc.f90:451-498 */
  for (int n = 1; n <= c->numreceptor; n++) {
    for (int ks = 1; ks <= c->nspec; ks++) cc[ks - 1] = 0.f;
    for (int i = 1; i <= S->numpart; i++) {
      if (S->itra1[i] != itime) continue;
      itage = abs(S->itra1[i] - S->itramem[i]);
      hz = fpo_minf(50.f + 0.3f * fpo_sqrtf((float)itage), hzmax);
      zd = S->ztra1[i] / hz;
      if (zd > 1.f) continue;
      hx = fpo_minf((0.29f + 2.222e-3f * fpo_sqrtf((float)itage)) * c->dx +
                        (float)itage * 1.2e-5f,
                    hxmax);
      xd = (float)((S->xtra1[i] - c->xreceptor[n - 1]) / hx);
      if (xd * xd > 1.f) continue;
      hy = fpo_minf((0.18f + 1.389e-3f * fpo_sqrtf((float)itage)) * c->dy +
                        (float)itage * 7.5e-6f,
                    hymax);
      yd = (float)((S->ytra1[i] - c->yreceptor[n - 1]) / hy);
      if (yd * yd > 1.f) continue;
      h = hx * hy * hz;
      r2 = xd * xd + yd * yd + zd * zd;
      if (r2 < 1.f) {
        xkern = factor * (1.f - r2);
        for (int ks = 1; ks <= c->nspec; ks++)
          cc[ks - 1] = cc[ks - 1] + XM1(S, i, ks) * xkern / h;
      }
    }
    for (int ks = 1; ks <= c->nspec; ks++)
      S->creceptor[(n - 1) + (size_t)MAXRECEPTOR * (ks - 1)] +=
          2.f * weight * cc[ks - 1] / c->receptorarea[n - 1];
  }
}

void fpo_fetch_grids(fpo_state *S, float *gridunc, float *griduncn,
                     float *drygridunc, float *drygriduncn, float *creceptor,
                     int zero_conc) {
  const fpb_config *c = &S->c;
  size_t ng = gsize(c, c->numxgrid, c->numygrid), nd = dsize(c, c->numxgrid, c->numygrid);
  if (gridunc) memcpy(gridunc, S->gridunc, ng * sizeof(float));
  if (drygridunc) memcpy(drygridunc, S->drygridunc, nd * sizeof(float));
  if (creceptor) memcpy(creceptor, S->creceptor, (size_t)MAXRECEPTOR * c->maxspec * sizeof(float));
  if (c->nested_output == 1) {
    size_t ngn = gsize(c, c->numxgridn, c->numygridn), ndn = dsize(c, c->numxgridn, c->numygridn);
    if (griduncn) memcpy(griduncn, S->griduncn, ngn * sizeof(float));
    if (drygriduncn) memcpy(drygriduncn, S->drygriduncn, ndn * sizeof(float));
    if (zero_conc) memset(S->griduncn, 0, ngn * sizeof(float));
  }
  if (zero_conc) { /* src/concoutput.f90:719-720 */
    memset(S->gridunc, 0, ng * sizeof(float));
    memset(S->creceptor, 0, (size_t)MAXRECEPTOR * c->maxspec * sizeof(float));
  }
}

/* src/timemanager.f90:269-304 */
void fpo_scale_depgrids(fpo_state *S, const float *factor) {
  const fpb_config *c = &S->c;
  for (int nest = 0; nest <= (c->nested_output == 1 ? 1 : 0); nest++) {
    const int nxg = nest ? c->numxgridn : c->numxgrid;
    const int nyg = nest ? c->numygridn : c->numygrid;
    float *grid = nest ? S->drygriduncn : S->drygridunc;
    float *wgrid = nest ? S->wetgriduncn : S->wetgridunc;
    for (int ks = 1; ks <= c->nspec; ks++)
      for (int kp = 1; kp <= c->maxpointspec_act; kp++)
        for (int na = 1; na <= c->nageclass; na++)
          for (int l = 1; l <= c->nclassunc; l++)
            for (int jy = 0; jy < nyg; jy++)
              for (int ix = 0; ix < nxg; ix++) {
                wgrid[didx(c, nxg, nyg, ix, jy, ks, kp, l, na)] *= factor[ks - 1];
                grid[didx(c, nxg, nyg, ix, jy, ks, kp, l, na)] *= factor[ks - 1];
              }
  }
}
