/*
 * fpo_domainfill.c -- oracle restatement of init_domainfill (test infrastructure):
 * src/init_domainfill.f90:55-283, the creation of the domain-filling particles (MDOMAINFILL = 1:
 * every particle carries the same share of the air mass of its column).
 *
 * and of boundcond_domainfill (src/boundcond_domainfill.f90:54-560): for a limited domain the
 * second half of init_domainfill (:287-389) memorises release heights at the four boundaries, and
 * every synchronisation step the routine terminates the particles that left the box and releases
 * new ones where the accumulated inflowing air mass reaches a particle's mass.
 *
 * Outside this restatement: MDOMAINFILL = 2 (stratospheric ozone tracer: the PV test and mass
 * scaling of :236-251), resuming from a particle dump (ipin = 1, boundcond.bin).
 *
 * Defined behaviour where the reference reads outside an array: for a boundary column with exactly
 * two release heights `zcolumn(k,i,j-2)` is element 0 of a 1-based dimension
 * (src/boundcond_domainfill.f90:115,353); it is 0 here.
 */
#include <math.h>
#include <stdlib.h>

#include "fpo.h"
#include "fpo_math.h"

#define MAXCOLUMN 3000 /* par_mod */
#define BCW(a, k, i, j) (a)[(size_t)((k)-1) + 2 * ((size_t)(i) + (size_t)S->c.nymax * (size_t)(j))]
#define BCS(a, k, i, j) (a)[(size_t)((k)-1) + 2 * ((size_t)(i) + (size_t)S->c.nxmax * (size_t)(j))]
#define M4(c, i, j, k) ((size_t)(i) + (size_t)(c)->nxmax * ((size_t)(j) + (size_t)(c)->nymax * (size_t)((k)-1)))
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
  /* :287-389: fewer release heights per column for the inflow boundaries */
  S->bc.nx_we[0] = nx_we[0]; S->bc.nx_we[1] = nx_we[1];
  S->bc.ny_sn[0] = ny_sn[0]; S->bc.ny_sn[1] = ny_sn[1];
  S->bc.gdomainfill = gdomainfill;
  S->bc.xmassperparticle = numparttot > 0 ? colmasstotal / (float)numparttot : 0.f;
  if (!gdomainfill && !rc) {
    float fractus = (float)numcolumn / (float)nz;
    fractus = fpo_sqrtf(fractus > 1.f ? fractus : 1.f) / 2.f;
    if (!S->bc.numcolumn_we) {
      S->bc.numcolumn_we = (int32_t *)calloc(2 * (size_t)c->nymax, sizeof(int32_t));
      S->bc.numcolumn_sn = (int32_t *)calloc(2 * (size_t)c->nxmax, sizeof(int32_t));
      S->bc.zcolumn_we = (float *)calloc(2 * (size_t)c->nymax * (MAXCOLUMN + 2), sizeof(float));
      S->bc.zcolumn_sn = (float *)calloc(2 * (size_t)c->nxmax * (MAXCOLUMN + 2), sizeof(float));
      S->bc.acc_mass_we = (float *)calloc(2 * (size_t)c->nymax * (MAXCOLUMN + 2), sizeof(float));
      S->bc.acc_mass_sn = (float *)calloc(2 * (size_t)c->nxmax * (MAXCOLUMN + 2), sizeof(float));
    }
    for (int jy = ny_sn[0]; jy <= ny_sn[1] && !rc; jy++)
      for (int ix = nx_we[0]; ix <= nx_we[1]; ix++) {
        ncolumn = fpo_nint_f(0.999f / fractus * (float)S->npart[1] * colmass[ix + (size_t)c->nxmax * jy] / colmasstotal);
        if (ncolumn > MAXCOLUMN) { rc = 2; break; } /* stop 'maxcolumn too small' */
        if (ncolumn == 0) continue;
        if (ix == nx_we[0]) S->bc.numcolumn_we[0 + 2 * jy] = ncolumn;
        if (ix == nx_we[1]) S->bc.numcolumn_we[1 + 2 * jy] = ncolumn;
        if (jy == ny_sn[0]) S->bc.numcolumn_sn[0 + 2 * ix] = ncolumn;
        if (jy == ny_sn[1]) S->bc.numcolumn_sn[1 + 2 * ix] = ncolumn;
        if (ix != nx_we[0] && ix != nx_we[1] && jy != ny_sn[0] && jy != ny_sn[1]) continue; /* (no effect) */
        for (int kz = 1; kz <= nz; kz++) pp[kz] = rho[M3(c, ix, jy, kz)] * r_air * tt[M3(c, ix, jy, kz)];
        deltacol = (pp[1] - pp[nz]) / (float)ncolumn;
        pnew = pp[1] + deltacol / 2.f;
        for (int j = 1; j <= ncolumn; j++) {
          pnew = pnew - deltacol;
          for (int kz = 1; kz <= nz - 1; kz++)
            if ((pp[kz] >= pnew) && (pp[kz + 1] < pnew)) {
              float zposition;
              dz1 = pp[kz] - pnew;
              dz2 = pnew - pp[kz + 1];
              dz = 1.f / (dz1 + dz2);
              zposition = (S->height[kz] * dz2 + S->height[kz + 1] * dz1) * dz;
              if (zposition > S->height[nz] - 0.5f) zposition = S->height[nz] - 0.5f;
              if (ix == nx_we[0]) BCW(S->bc.zcolumn_we, 1, jy, j) = zposition;
              if (ix == nx_we[1]) BCW(S->bc.zcolumn_we, 2, jy, j) = zposition;
              if (jy == ny_sn[0]) BCS(S->bc.zcolumn_sn, 1, ix, j) = zposition;
              if (jy == ny_sn[1]) BCS(S->bc.zcolumn_sn, 2, ix, j) = zposition;
              /* (acc_mass_* start from zero: the reference clears acc_mass_we/sn(1:2,jy,j) here) */
            }
        }
      }
  }

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


int fpo_boundcond_locations(fpo_state *S, double *accmass_sum) {
  int n = 0;
  double acc = 0.;
  if (!S->bc.numcolumn_we) return 0;
  for (int jy = S->bc.ny_sn[0]; jy <= S->bc.ny_sn[1]; jy++)
    for (int k = 1; k <= 2; k++)
      for (int j = 1; j <= S->bc.numcolumn_we[(k - 1) + 2 * jy]; j++) { n++; acc += BCW(S->bc.acc_mass_we, k, jy, j); }
  for (int ix = S->bc.nx_we[0]; ix <= S->bc.nx_we[1]; ix++)
    for (int k = 1; k <= 2; k++)
      for (int j = 1; j <= S->bc.numcolumn_sn[(k - 1) + 2 * ix]; j++) { n++; acc += BCS(S->bc.acc_mass_sn, k, ix, j); }
  if (accmass_sum) *accmass_sum = acc;
  return n;
}

/* one release location of the boundary loops, src/boundcond_domainfill.f90:104-317 (west/east,
 * we = 1) and :343-545 (south/north, we = 0): mass flux, accumulated mass, particle creation */
static int bc_location(fpo_state *S, int itime, int itsplit, int we, int k, int idx, int j, float cosfact,
                       float dt1, float dt2, float dtt, int *minpart, int *created) {
  const fpb_config *c = &S->c;
  const int nz = c->nz;
  const int *nx_we = S->bc.nx_we, *ny_sn = S->bc.ny_sn;
  const int ncol = we ? S->bc.numcolumn_we[(k - 1) + 2 * idx] : S->bc.numcolumn_sn[(k - 1) + 2 * idx];
  float *zc = we ? S->bc.zcolumn_we : S->bc.zcolumn_sn;
  float *accp = we ? &BCW(S->bc.acc_mass_we, k, idx, j) : &BCS(S->bc.acc_mass_sn, k, idx, j);
#define Z(jj) (we ? BCW(zc, k, idx, jj) : BCS(zc, k, idx, jj))
  float deltaz, boundarea, dz1, dz2, dz, windl[3], rhol[3], windhl[3], rhohl[3], windx, rhox, fluxofmass;
  int indz = 0, indzp = 0, mmass, ipart;
  const float xmpp = S->bc.xmassperparticle;
  /* grid point of the boundary location */
  const int gx = we ? nx_we[k - 1] : idx, gy = we ? idx : ny_sn[k - 1];

  if (j == 1) deltaz = (Z(2) + Z(1)) / 2.f;
  else if (j == ncol) deltaz = (Z(j) - Z(j - 2)) / 2.f;
  else deltaz = (Z(j + 1) - Z(j - 1)) / 2.f;
  if (we) {
    if ((idx == ny_sn[0]) || (idx == ny_sn[1])) boundarea = deltaz * 111198.5f / 2.f * c->dy;
    else boundarea = deltaz * 111198.5f * c->dy;
  } else {
    if ((idx == nx_we[0]) || (idx == nx_we[1])) boundarea = deltaz * 111198.5f / 2.f * cosfact * c->dx;
    else boundarea = deltaz * 111198.5f * cosfact * c->dx;
  }
  for (int i = 2; i <= nz; i++)
    if (S->height[i] > Z(j)) { indz = i - 1; indzp = i; break; }
  dz1 = Z(j) - S->height[indz];
  dz2 = S->height[indzp] - Z(j);
  dz = 1.f / (dz1 + dz2);
  for (int m = 1; m <= 2; m++) {
    const fpb_met_ptrs *M = &S->met[S->memind[m]];
    for (int in = 1; in <= 2; in++) {
      const int indzh = indz + in - 1;
      windl[in] = we ? M->uu[M4(c, gx, gy, indzh)] : M->vv[M4(c, gx, gy, indzh)];
      rhol[in] = M->rho[M4(c, gx, gy, indzh)];
    }
    windhl[m] = (dz2 * windl[1] + dz1 * windl[2]) * dz;
    rhohl[m] = (dz2 * rhol[1] + dz1 * rhol[2]) * dz;
  }
  windx = (windhl[1] * dt2 + windhl[2] * dt1) * dtt;
  rhox = (rhohl[1] * dt2 + rhohl[2] * dt1) * dtt;
  fluxofmass = windx * rhox * boundarea * (float)c->lsynctime;

  if (k == 1) {
    if (fluxofmass >= 0.f) *accp = *accp + fluxofmass; else *accp = 0.f;
  } else {
    if (fluxofmass <= 0.f) *accp = *accp + fabsf(fluxofmass); else *accp = 0.f;
  }
  if (*accp >= xmpp / 2.f) {
    mmass = fpo_int_f((*accp + xmpp / 2.f) / xmpp);
    *accp = *accp - (float)mmass * xmpp;
  } else {
    mmass = 0;
  }

  for (int m = 1; m <= mmass; m++) {
    for (ipart = *minpart; ipart <= S->maxpart; ipart++) {
      if (S->itra1[ipart] != itime) {
        if (we) {
          S->xtra1[ipart] = (float)nx_we[k - 1];
          if (idx == ny_sn[0]) S->ytra1[ipart] = (float)idx + 0.5f * fpo_ran1(S, &S->bc.idummy);
          else if (idx == ny_sn[1]) S->ytra1[ipart] = (float)idx - 0.5f * fpo_ran1(S, &S->bc.idummy);
          else S->ytra1[ipart] = (float)idx + (fpo_ran1(S, &S->bc.idummy) - .5f);
        } else {
          S->ytra1[ipart] = (float)ny_sn[k - 1];
          if (idx == nx_we[0]) S->xtra1[ipart] = (float)idx + 0.5f * fpo_ran1(S, &S->bc.idummy);
          else if (idx == nx_we[1]) S->xtra1[ipart] = (float)idx - 0.5f * fpo_ran1(S, &S->bc.idummy);
          else S->xtra1[ipart] = (float)idx + (fpo_ran1(S, &S->bc.idummy) - .5f);
        }
        if (j == 1) S->ztra1[ipart] = Z(1) + (Z(2) - Z(1)) / 4.f;
        else if (j == ncol) S->ztra1[ipart] = (2.f * Z(j) + Z(j - 1) + S->height[nz]) / 4.f;
        else S->ztra1[ipart] = Z(j - 1) + fpo_ran1(S, &S->bc.idummy) * (Z(j + 1) - Z(j - 1));
        /* (the PV interpolation of :205-243 only feeds the MDOMAINFILL = 2 test) */
        {
          int nc = fpo_int_f(fpo_ran1(S, &S->bc.idummy) * (float)c->nclassunc) + 1;
          S->nclass[ipart] = nc < c->nclassunc ? nc : c->nclassunc;
        }
        S->numparticlecount = S->numparticlecount + 1;
        S->npoint[ipart] = S->numparticlecount;
        S->idt[ipart] = c->mintime;
        S->itra1[ipart] = itime;
        S->itramem[ipart] = S->itra1[ipart];
        S->itrasplit[ipart] = S->itra1[ipart] + c->ldirect * itsplit;
        XM1(S, ipart, 1) = xmpp;
        /* (what a fresh particle needs beyond the reference's assignments: the turbulent velocity
         * memory of a slot is whatever its last owner left; initialize() overwrites it) */
        if (ipart > S->numpart) S->numpart = ipart;
        (*created)++;
        break;
      }
    }
    if (ipart > S->maxpart) return 1; /* 'too many particles required' */
    *minpart = ipart + 1;
  }
#undef Z
  return 0;
}

int fpo_boundcond_domainfill(fpo_state *S, int itime, int itsplit) {
  const fpb_config *c = &S->c;
  const float pi180 = 3.14159265f / 180.f;
  const int *nx_we = S->bc.nx_we, *ny_sn = S->bc.ny_sn;
  int minpart = 1, created = 0;
  if (S->bc.gdomainfill || !S->bc.numcolumn_we) return 0;
  for (int i = 1; i <= S->numpart; i++) {
    if (S->itra1[i] == itime) {
      if ((S->ytra1[i] > (float)ny_sn[1]) || (S->ytra1[i] < (float)ny_sn[0])) S->itra1[i] = FPB_ITRA_DEAD;
      if (((!c->xglobal) || (nx_we[1] != (c->nx - 2))) &&
          ((S->xtra1[i] < (float)nx_we[0]) || (S->xtra1[i] > (float)nx_we[1])))
        S->itra1[i] = FPB_ITRA_DEAD;
    }
  }
  const float dt1 = (float)(itime - S->memtime[1]), dt2 = (float)(S->memtime[2] - itime), dtt = 1.f / (dt1 + dt2);
  for (int jy = ny_sn[0]; jy <= ny_sn[1]; jy++)
    for (int k = 1; k <= 2; k++)
      for (int j = 1; j <= S->bc.numcolumn_we[(k - 1) + 2 * jy]; j++)
        if (bc_location(S, itime, itsplit, 1, k, jy, j, 0.f, dt1, dt2, dtt, &minpart, &created)) return -1;
  for (int ix = nx_we[0]; ix <= nx_we[1]; ix++)
    for (int k = 1; k <= 2; k++) {
      const float ylat = c->ylat0 + (float)ny_sn[k - 1] * c->dy;
      const float cosfact = fpo_cosf(ylat * pi180);
      for (int j = 1; j <= S->bc.numcolumn_sn[(k - 1) + 2 * ix]; j++)
        if (bc_location(S, itime, itsplit, 0, k, ix, j, cosfact, dt1, dt2, dtt, &minpart, &created)) return -1;
    }
  return created;
}
