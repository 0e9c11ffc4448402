/*
 * fpo_domainfill.c -- oracle restatement of init_domainfill (test infrastructure):
 * src/init_domainfill.f90:55-283, the creation of the domain-filling particles (MDOMAINFILL = 1:
 * every particle carries the same share of the air mass of its column).
 *
 * Outside this restatement: MDOMAINFILL = 2 (stratospheric ozone tracer: the PV test and mass
 * scaling of :236-251), resuming from a particle dump (ipin = 1), and the second half of the
 * routine (:287-398), which prepares the inflow columns of boundcond_domainfill for a limited
 * domain -- for a global domain (gdomainfill, :62-68) boundcond_domainfill returns at once
 * (src/boundcond_domainfill.f90:54).
 */
#include <math.h>
#include <stdlib.h>

#include "fpo.h"
#include "fpo_math.h"

#define XM1(S, j, ks) (S)->xmass1[(size_t)(j) + (size_t)((S)->maxpart + 1) * ((ks)-1)]
#define M3(c, i, j, k) ((size_t)(i) + (size_t)(c)->nxmax * ((size_t)(j) + (size_t)(c)->nymax * (size_t)((k)-1)))

/* gridarea(jy) of src/init_domainfill.f90:81-130 for jy = ny_sn[0]..ny_sn[1] */
void fpo_domainfill_gridarea(const fpb_config *c, const int ny_sn[2], float *gridarea) {
  const float pi = 3.14159265f, r_earth = 6.371e6f, pih = pi / 180.f;
  float ylat, ylatp, ylatm, hzone, cosfactp, cosfactm;
  for (int jy = ny_sn[0]; jy <= ny_sn[1]; jy++) {
    ylat = c->ylat0 + (float)jy * c->dy;
    ylatp = ylat + 0.5f * c->dy;
    ylatm = ylat - 0.5f * c->dy;
    if ((ylatm < 0.f) && (ylatp > 0.f)) {
      hzone = 1.f / c->dyconst;
    } else {
      cosfactp = fpo_cosf(ylatp * pih) * r_earth;
      cosfactm = fpo_cosf(ylatm * pih) * r_earth;
      if (cosfactp < cosfactm)
        hzone = fpo_sqrtf(r_earth * r_earth - cosfactp * cosfactp) - fpo_sqrtf(r_earth * r_earth - cosfactm * cosfactm);
      else
        hzone = fpo_sqrtf(r_earth * r_earth - cosfactm * cosfactm) - fpo_sqrtf(r_earth * r_earth - cosfactp * cosfactp);
    }
    gridarea[jy] = 2.f * pi * r_earth * hzone * c->dx / 360.f;
  }
  if (c->sglobal) {
    ylat = c->ylat0;
    ylatp = ylat + 0.5f * c->dy;
    cosfactm = 0.f;
    cosfactp = fpo_cosf(ylatp * pih) * r_earth;
    hzone = fpo_sqrtf(r_earth * r_earth - cosfactm * cosfactm) - fpo_sqrtf(r_earth * r_earth - cosfactp * cosfactp);
    gridarea[0] = 2.f * pi * r_earth * hzone * c->dx / 360.f;
  }
  if (c->nglobal) {
    ylat = c->ylat0 + (float)c->nymin1 * c->dy;
    ylatm = ylat - 0.5f * c->dy;
    cosfactp = 0.f;
    cosfactm = fpo_cosf(ylatm * pih) * r_earth;
    hzone = fpo_sqrtf(r_earth * r_earth - cosfactp * cosfactp) - fpo_sqrtf(r_earth * r_earth - cosfactm * cosfactm);
    gridarea[c->nymin1] = 2.f * pi * r_earth * hzone * c->dx / 360.f;
  }
}

/* out[0..1] = nx_we, out[2..3] = ny_sn, out[4] = gdomainfill, out[5] = numcolumn, out[6] = numparttot;
 * fout[0] = colmasstotal, fout[1] = xmassperparticle.  Returns 1 when numpart would exceed maxpart. */
int fpo_init_domainfill(fpo_state *S, float xpoint1, float ypoint1, float xpoint2, float ypoint2,
                        int itsplit, int32_t *out, float *fout) {
  const fpb_config *c = &S->c;
  const float r_air = 287.05f, ga = 9.81f;
  const int nz = c->nz;
  const float *rho = S->met[1].rho, *tt = S->met[1].tt; /* slot 1, literal in the reference */
  int nx_we[2], ny_sn[2], gdomainfill = 0, numcolumn = 0, numparttot = 0, ncolumn, jj;
  float colmasstotal, deltacol, pnew, dz1, dz2, dz;
  float *gridarea = (float *)calloc((size_t)c->nymax + 1, sizeof(float));
  float *colmass = (float *)calloc((size_t)c->nxmax * c->nymax, sizeof(float));
  float *pp = (float *)calloc((size_t)nz + 2, sizeof(float));
  int rc = 0;

  nx_we[0] = fpo_int_f(xpoint1) > 0 ? fpo_int_f(xpoint1) : 0;
  nx_we[1] = (fpo_int_f(xpoint2) + 1) < c->nxmin1 ? (fpo_int_f(xpoint2) + 1) : c->nxmin1;
  ny_sn[0] = fpo_int_f(ypoint1) > 0 ? fpo_int_f(ypoint1) : 0;
  ny_sn[1] = (fpo_int_f(ypoint2) + 1) < c->nymin1 ? (fpo_int_f(ypoint2) + 1) : c->nymin1;
  if (c->xglobal && c->sglobal && c->nglobal)
    gdomainfill = (nx_we[0] == 0) && (nx_we[1] == c->nxmin1) && (ny_sn[0] == 0) && (ny_sn[1] == c->nymin1);
  if (c->xglobal) nx_we[1] = nx_we[1] < c->nx - 2 ? nx_we[1] : c->nx - 2;

  fpo_domainfill_gridarea(c, ny_sn, gridarea);

  colmasstotal = 0.f;
  for (int jy = ny_sn[0]; jy <= ny_sn[1]; jy++)
    for (int ix = nx_we[0]; ix <= nx_we[1]; ix++) {
      pp[1] = rho[M3(c, ix, jy, 1)] * r_air * tt[M3(c, ix, jy, 1)];
      pp[nz] = rho[M3(c, ix, jy, nz)] * r_air * tt[M3(c, ix, jy, nz)];
      colmass[ix + (size_t)c->nxmax * jy] = (pp[1] - pp[nz]) / ga * gridarea[jy];
      colmasstotal = colmasstotal + colmass[ix + (size_t)c->nxmax * jy];
    }

  S->numpart = 0; /* ipin == 0 */
  for (int jy = ny_sn[0]; jy <= ny_sn[1] && !rc; jy++) {
    for (int ix = nx_we[0]; ix <= nx_we[1]; ix++) {
      const float cm = colmass[ix + (size_t)c->nxmax * jy];
      ncolumn = fpo_nint_f(0.999f * (float)S->npart[1] * cm / colmasstotal);
      if (ncolumn == 0) continue;
      if (ncolumn > numcolumn) numcolumn = ncolumn;
      if ((long)S->numpart + ncolumn > S->maxpart) { rc = 1; break; } /* (the reference writes out of bounds) */
      for (int kz = 1; kz <= nz; kz++) pp[kz] = rho[M3(c, ix, jy, kz)] * r_air * tt[M3(c, ix, jy, kz)];
      deltacol = (pp[1] - pp[nz]) / (float)ncolumn;
      pnew = pp[1] + deltacol / 2.f;
      jj = 0;
      for (int j = 1; j <= ncolumn; j++) {
        jj = jj + 1;
        if (ncolumn > 20)
          pnew = pnew - deltacol;
        else
          pnew = pp[1] - fpo_ran1(S, &S->idummy_domainfill) * (pp[1] - pp[nz]);
        for (int kz = 1; kz <= nz - 1; kz++) {
          if ((pp[kz] >= pnew) && (pp[kz + 1] < pnew)) {
            const int n = S->numpart + jj;
            dz1 = pp[kz] - pnew;
            dz2 = pnew - pp[kz + 1];
            dz = 1.f / (dz1 + dz2);
            S->xtra1[n] = (float)ix - 0.5f + fpo_ran1(S, &S->idummy_domainfill);
            if (ix == 0) S->xtra1[n] = fpo_ran1(S, &S->idummy_domainfill);
            if (ix == c->nxmin1) S->xtra1[n] = (float)c->nxmin1 - fpo_ran1(S, &S->idummy_domainfill);
            S->ytra1[n] = (float)jy - 0.5f + fpo_ran1(S, &S->idummy_domainfill);
            S->ztra1[n] = (S->height[kz] * dz2 + S->height[kz + 1] * dz1) * dz;
            if (S->ztra1[n] > S->height[nz] - 0.5f) S->ztra1[n] = S->height[nz] - 0.5f;
            /* (the PV interpolation of :196-224 only feeds the MDOMAINFILL = 2 test) */
            {
              int nc = fpo_int_f(fpo_ran1(S, &S->idummy_domainfill) * (float)c->nclassunc) + 1;
              S->nclass[n] = nc < c->nclassunc ? nc : c->nclassunc;
            }
            S->numparticlecount = S->numparticlecount + 1;
            S->npoint[n] = S->numparticlecount;
            S->idt[n] = c->mintime;
            S->itra1[n] = 0;
            S->itramem[n] = 0;
            S->itrasplit[n] = S->itra1[n] + c->ldirect * itsplit;
            XM1(S, n, 1) = cm / (float)ncolumn;
          }
        }
      }
      numparttot = numparttot + ncolumn;
      S->numpart = S->numpart + jj;
    }
  }

  /* :266-271 */
  for (int j = 1; j <= S->numpart; j++)
    if ((S->xtra1[j] < 0.) || (S->xtra1[j] >= (float)c->nxmin1) || (S->ytra1[j] < 0.) ||
        (S->ytra1[j] >= (float)c->nymin1))
      S->itra1[j] = FPB_ITRA_DEAD;
  /* :391-397 */
  for (int i = S->numpart; i >= 1; i--) {
    if (S->itra1[i] == FPB_ITRA_DEAD) S->numpart = S->numpart - 1;
    else break;
  }

  out[0] = nx_we[0]; out[1] = nx_we[1]; out[2] = ny_sn[0]; out[3] = ny_sn[1];
  out[4] = gdomainfill; out[5] = numcolumn; out[6] = numparttot;
  fout[0] = colmasstotal;
  fout[1] = numparttot > 0 ? colmasstotal / (float)numparttot : 0.f;
  free(gridarea); free(colmass); free(pp);
  return rc;
}
