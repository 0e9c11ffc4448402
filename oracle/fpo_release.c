/*
 * fpo_release.c -- oracle restatement of releaseparticles' integer semantics
 * (release counts, slot search) and its ran1 position stream (test
 * infrastructure).  Follows src/releaseparticles.f90:69-378.
 *
 * Outside this restatement (host-side configuration I/O):
 * the EMISVAR hour/day-of-week factors (point_hour, area_dow ...; all 1.0
 * without EMISVAR files), zkind 2/3 (needs orography / pressure), and the
 * ind_rel density scaling.  With those at their defaults the statements
 * below are the complete routine.
 */
#include <math.h>

#include "fpo.h"
#include "fpo_math.h"

#define XM1(S, j, ks) (S)->xmass1[(size_t)(j) + (size_t)((S)->maxpart + 1) * ((ks)-1)]
#define XSC(S, j, ks) (S)->xscav_frac1[(size_t)(j) + (size_t)((S)->maxpart + 1) * ((ks)-1)]

int fpo_releaseparticles(fpo_state *S, int itime, int numpoint,
                         const int32_t *ireleasestart,
                         const int32_t *ireleaseend, const float *xpoint1,
                         const float *ypoint1, const float *xpoint2,
                         const float *ypoint2, const float *zpoint1,
                         const float *zpoint2, float *xmasssave, int itsplit) {
  const fpb_config *c = &S->c;
  const float eps2 = 1.e-6f; /* releaseparticles.f90: eps2=1.e-6 */
  int minpart = 1, numrel, ipart;
  float rfraction, xaux, yaux, zaux;
  const float timecorrect = 1.f, average_timecorrect = 1.f;

  for (int i = 1; i <= numpoint; i++) {
    if (!((itime >= ireleasestart[i - 1]) && (itime <= ireleaseend[i - 1]))) continue;

    if (ireleasestart[i - 1] != ireleaseend[i - 1]) {
      rfraction = fabsf((float)S->npart[i] * (float)c->lsynctime /
                        (float)(ireleaseend[i - 1] - ireleasestart[i - 1]));
      if ((itime == ireleasestart[i - 1]) || (itime == ireleaseend[i - 1]))
        rfraction = rfraction / 2.f;
      rfraction = rfraction * average_timecorrect;
      rfraction = rfraction + xmasssave[i - 1];
      numrel = fpo_int_f(rfraction);
      xmasssave[i - 1] = rfraction - (float)numrel;
    } else {
      numrel = S->npart[i];
    }

    xaux = xpoint2[i - 1] - xpoint1[i - 1];
    yaux = ypoint2[i - 1] - ypoint1[i - 1];
    zaux = zpoint2[i - 1] - zpoint1[i - 1];
    for (int j = 1; j <= numrel; j++) {
      for (ipart = minpart; ipart <= S->maxpart; ipart++) {
        if (S->itra1[ipart] != itime) {
          S->xtra1[ipart] = xpoint1[i - 1] + fpo_ran1(S, &S->idummy_release) * xaux;
          if (c->xglobal) {
            if (S->xtra1[ipart] > (float)c->nxmin1)
              S->xtra1[ipart] = S->xtra1[ipart] - (float)c->nxmin1;
            if (S->xtra1[ipart] < 0.)
              S->xtra1[ipart] = S->xtra1[ipart] + (float)c->nxmin1;
          }
          S->ytra1[ipart] = ypoint1[i - 1] + fpo_ran1(S, &S->idummy_release) * yaux;
          for (int k = 1; k <= c->nspec; k++) {
            XM1(S, ipart, k) = S->xmass[(i - 1) + (size_t)c->numpoint * (k - 1)] /
                               (float)S->npart[i] * timecorrect / average_timecorrect;
            if (c->drybkdep || c->wetbkdep) XSC(S, ipart, k) = -1.f;
          }
          {
            int nc = fpo_int_f(fpo_ran1(S, &S->idummy_release) * (float)c->nclassunc) + 1;
            S->nclass[ipart] = nc < c->nclassunc ? nc : c->nclassunc;
          }
          S->npoint[ipart] = i; /* mquasilag == 0 */
          S->idt[ipart] = c->mintime;
          S->itra1[ipart] = itime;
          S->itramem[ipart] = S->itra1[ipart];
          S->itrasplit[ipart] = S->itra1[ipart] + c->ldirect * itsplit;
          S->ztra1[ipart] = zpoint1[i - 1] + fpo_ran1(S, &S->idummy_release) * zaux;
          if (S->ztra1[ipart] < eps2) S->ztra1[ipart] = eps2;
          if (S->ztra1[ipart] > S->height[c->nz] - 0.5f)
            S->ztra1[ipart] = S->height[c->nz] - 0.5f;
          if (ipart > S->numpart) S->numpart = ipart;
          break;
        }
      }
      if (ipart > S->maxpart) return 1; /* label 996: too many particles */
      minpart = ipart + 1;
    }
  }
  return 0;
}

/* particle splitting, src/timemanager.f90:472-503 (the caller keeps the outer itsplit test) */
void fpo_split_particles(fpo_state *S, int itime) {
  const fpb_config *c = &S->c;
  int n = S->numpart;
  for (int j = 1; j <= S->numpart; j++) {
    if (c->ldirect * itime >= c->ldirect * S->itrasplit[j]) {
      if (n < c->maxpart) {
        n = n + 1;
        S->itrasplit[j] = 2 * (S->itrasplit[j] - S->itramem[j]) + S->itramem[j];
        S->itrasplit[n] = S->itrasplit[j];
        S->itramem[n] = S->itramem[j];
        S->itra1[n] = S->itra1[j];
        S->idt[n] = S->idt[j];
        S->npoint[n] = S->npoint[j];
        S->nclass[n] = S->nclass[j];
        S->xtra1[n] = S->xtra1[j];
        S->ytra1[n] = S->ytra1[j];
        S->ztra1[n] = S->ztra1[j];
        S->uap[n] = S->uap[j];
        S->ucp[n] = S->ucp[j];
        S->uzp[n] = S->uzp[j];
        S->us[n] = S->us[j];
        S->vs[n] = S->vs[j];
        S->ws[n] = S->ws[j];
        S->cbt[n] = S->cbt[j];
        for (int ks = 1; ks <= c->nspec; ks++) {
          XM1(S, j, ks) = XM1(S, j, ks) / 2.f;
          XM1(S, n, ks) = XM1(S, j, ks);
        }
      }
    }
  }
  S->numpart = n;
}
