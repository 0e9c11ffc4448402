/* fpo_hooks.c -- TEST INFRASTRUCTURE (the checker, never the product): sequential restatement of the optional
 * hooks of the particle loop, src/timemanager.f90:614-623,631,702,733-737:
 *   fpo_calcfluxes         src/calcfluxes.f90:43-166
 *   fpo_partpos_average    src/partpos_average.f90:47-186
 *   fpo_initial_cond_calc  src/initial_cond_calc.f90:49-204
 * Pinned against the reference's own routines (oracle/_ref) in tests/test_oracle_hooks.py.  Arrays are in the
 * reference's layouts: fields (0:nxmax-1,0:nymax-1,nzmax), flux(6,0:numxgrid-1,0:numygrid-1,numzgrid,nspec,
 * maxpointspec_act,nageclass), init_cond(0:numxgrid-1,0:numygrid-1,numzgrid,maxspec,maxpointspec_act). */
#include <math.h>
#include <stddef.h>

#include "fpo.h"
#include "fpo_math.h"

/* mass[k-1] = xmass1(jpart,k); nage 1..nageclass (src/timemanager.f90:544-547), npoint = npoint(jpart) */
void fpo_calcfluxes(const fpb_config *c, float *flux, int nage, int npoint, float xold, float yold, float zold,
                    double xtra1, double ytra1, float ztra1, const float *mass) {
  const int kp = ((c->ioutputforeachrelease == 1) && (c->mdomainfill == 0)) ? npoint : 1;
  const float xmean = (float)(((double)xold + xtra1) / 2.0);
  const float ymean = (float)(((double)yold + ytra1) / 2.0);
  const int ixave = (int)((xmean * c->dx + c->xoutshift) / c->dxout);
  const int jyave = (int)((ymean * c->dy + c->youtshift) / c->dyout);
  int kz, kzave;
  for (kz = 1; kz <= c->numzgrid; kz++)
    if (c->outheight[kz - 1] > ztra1) break;
  kzave = kz;
  const size_t nxg = c->numxgrid, nyg = c->numygrid, nzg = c->numzgrid;
#define FLUX(i, ix, jy, kz, k)                                                                              \
  flux[((i)-1) + 6 * ((ix) + nxg * ((jy) + nyg * (((kz)-1) + nzg * (((k)-1) + (size_t)c->nspec *              \
       ((kp - 1) + (size_t)c->maxpointspec_act * (nage - 1))))))]
#define HALF(kz) ((kz) == 1 ? c->outheight[0] / 2.f : (c->outheight[(kz)-2] + c->outheight[(kz)-1]) / 2.f) /* readoutgrid.f90:194-197 */
  if ((ixave >= 0) && (jyave >= 0) && (ixave <= c->numxgrid - 1) && (jyave <= c->numygrid - 1)) {
    for (kz = 1; kz <= c->numzgrid; kz++)
      if (HALF(kz) > zold) break;
    const int k1 = kz < c->numzgrid ? kz : c->numzgrid;
    for (kz = 1; kz <= c->numzgrid; kz++)
      if (HALF(kz) > ztra1) break;
    const int k2 = kz < c->numzgrid ? kz : c->numzgrid;
    for (int k = 1; k <= c->nspec; k++) {
      for (kz = k1; kz <= k2 - 1; kz++) FLUX(5, ixave, jyave, kz, k) = FLUX(5, ixave, jyave, kz, k) + mass[k - 1];
      for (kz = k2; kz <= k1 - 1; kz++) FLUX(6, ixave, jyave, kz, k) = FLUX(6, ixave, jyave, kz, k) + mass[k - 1];
    }
  }
  if ((kzave <= c->numzgrid) && (jyave >= 0) && (jyave <= c->numygrid - 1)) {
    if (fabs((double)xold - xtra1) < (double)((float)c->nx / 2.f)) {
      const int ix1 = (int)((xold * c->dx + c->xoutshift) / c->dxout + 0.5f);
      const int ix2 = (int)((xtra1 * (double)c->dx + (double)c->xoutshift) / (double)c->dxout + 0.5);
      for (int k = 1; k <= c->nspec; k++) {
        for (int ix = ix1; ix <= ix2 - 1; ix++)
          if ((ix >= 0) && (ix <= c->numxgrid - 1)) FLUX(1, ix, jyave, kzave, k) = FLUX(1, ix, jyave, kzave, k) + mass[k - 1];
        for (int ix = ix2; ix <= ix1 - 1; ix++)
          if ((ix >= 0) && (ix <= c->numxgrid - 1)) FLUX(2, ix, jyave, kzave, k) = FLUX(2, ix, jyave, kzave, k) + mass[k - 1];
      }
    } else {
      const int ixs = (int)((((float)c->nxmin1 - 1.0e5f) * c->dx + c->xoutshift) / c->dxout);
      if ((ixs >= 0) && (ixs <= c->numxgrid - 1)) {
        const int i = ((double)xold > xtra1) ? 1 : 2;
        for (int k = 1; k <= c->nspec; k++) FLUX(i, ixs, jyave, kzave, k) = FLUX(i, ixs, jyave, kzave, k) + mass[k - 1];
      }
    }
  }
  if ((kzave <= c->numzgrid) && (ixave >= 0) && (ixave <= c->numxgrid - 1)) {
    const int jy1 = (int)((yold * c->dy + c->youtshift) / c->dyout + 0.5f);
    const int jy2 = (int)((ytra1 * (double)c->dy + (double)c->youtshift) / (double)c->dyout + 0.5);
    for (int k = 1; k <= c->nspec; k++) {
      for (int jy = jy1; jy <= jy2 - 1; jy++)
        if ((jy >= 0) && (jy <= c->numygrid - 1)) FLUX(3, ixave, jy, kzave, k) = FLUX(3, ixave, jy, kzave, k) + mass[k - 1];
      for (int jy = jy2; jy <= jy1 - 1; jy++)
        if ((jy >= 0) && (jy <= c->numygrid - 1)) FLUX(4, ixave, jy, kzave, k) = FLUX(4, ixave, jy, kzave, k) + mass[k - 1];
    }
  }
#undef FLUX
#undef HALF
}

/* out[14]: the increments of part_av_cartx, carty, cartz, z, topo, pv, qv, tt, uu, vv, rho, tro, hmix, energy */
void fpo_partpos_average(const fpb_config *c, const float *height, int itime, const int32_t memtime[2], double xtra1,
                         double ytra1, float ztra1, const float *oro, const float *pv[2], const float *qv[2],
                         const float *tt[2], const float *uu[2], const float *vv[2], const float *rho[2],
                         const float *hmix[2], const float *tropopause[2], float out[14]) {
  const size_t nxm = c->nxmax, plane = (size_t)c->nxmax * c->nymax;
  const float dt1 = (float)(itime - memtime[0]), dt2 = (float)(memtime[1] - itime);
  const float dtt = 1.f / (dt1 + dt2);
  float xlon = (float)((double)c->xlon0 + xtra1 * (double)c->dx);
  float ylat = (float)((double)c->ylat0 + ytra1 * (double)c->dy);
  const int ix = (int)xtra1, jy = (int)ytra1;
  const int ixp = ix + 1;
  int jyp = jy + 1;
  const float ddx = (float)(xtra1 - (double)(float)ix), ddy = (float)(ytra1 - (double)(float)jy);
  const float rddx = 1.f - ddx, rddy = 1.f - ddy;
  const float p1 = rddx * rddy, p2 = ddx * rddy, p3 = rddx * ddy, p4 = ddx * ddy;
  if (jyp >= c->nymax) jyp = jyp - 1;
#define F2(f) (p1 * (f)[ix + nxm * jy] + p2 * (f)[ixp + nxm * jy] + p3 * (f)[ix + nxm * jyp] + p4 * (f)[ixp + nxm * jyp])
#define F3(f, k) (p1 * (f)[ix + nxm * jy + plane * ((k)-1)] + p2 * (f)[ixp + nxm * jy + plane * ((k)-1)] + \
                  p3 * (f)[ix + nxm * jyp + plane * ((k)-1)] + p4 * (f)[ixp + nxm * jyp + plane * ((k)-1)])
  const float topo = F2(oro);
  int indz = c->nz - 1;
  for (int il = 2; il <= c->nz; il++)
    if (height[il - 1] > ztra1) { indz = il - 1; break; }
  const int indzp = indz + 1;
  const float dz1 = ztra1 - height[indz - 1], dz2 = height[indzp - 1] - ztra1;
  const float dz = 1.f / (dz1 + dz2);
  float prof[6][2];
  for (int ind = indz; ind <= indzp; ind++) {
    const float *const *fld[6] = {pv, qv, tt, uu, vv, rho};
    for (int q = 0; q < 6; q++) {
      const float v1 = F3(fld[q][0], ind), v2 = F3(fld[q][1], ind);
      prof[q][ind - indz] = (v1 * dt2 + v2 * dt1) * dtt;
    }
  }
  float vi[6];
  for (int q = 0; q < 6; q++) vi[q] = (dz1 * prof[q][1] + dz2 * prof[q][0]) * dz;
  const float pvi = vi[0], qvi = vi[1], tti = vi[2], uui = vi[3], vvi = vi[4], rhoi = vi[5];
  float tr[2], hm[2];
  for (int m = 0; m < 2; m++) {
    tr[m] = F2(tropopause[m]);
    hm[m] = F2(hmix[m]);
  }
#undef F2
#undef F3
  const float hmixi = (hm[0] * dt2 + hm[1] * dt1) * dtt;
  const float tri = (tr[0] * dt2 + tr[1] * dt1) * dtt;
  const float energy = ((tti * 1004.6f + (ztra1 + topo) * 9.81f) + qvi * 2501000.f) + (uui * uui + vvi * vvi) / 2.f;
  const float pi180 = 3.14159265f / 180.f;
  xlon = xlon * pi180;
  ylat = ylat * pi180;
  const float cy = fpo_cosf(ylat), sy = fpo_sinf(ylat), cx = fpo_cosf(xlon), sx = fpo_sinf(xlon);
  out[0] = cy * sx;
  out[1] = -((1.0f * cy) * cx);
  out[2] = sy;
  out[3] = ztra1; out[4] = topo; out[5] = pvi; out[6] = qvi; out[7] = tti; out[8] = uui; out[9] = vvi;
  out[10] = rhoi; out[11] = tri; out[12] = hmixi; out[13] = energy;
}

/* rho2: rho(:,:,:,memind(2)) (linit_cond = 1 only); mass[k-1] = xmass1(i,k) */
void fpo_initial_cond_calc(const fpb_config *c, const float *height, float *init_cond, int linit_cond, double xtra1,
                           double ytra1, float ztra1, int npoint, const float *rho2, const float *mass) {
  const size_t nxm = c->nxmax, plane = (size_t)c->nxmax * c->nymax;
  float rhoi = 1.f;
  if (linit_cond == 1) {
    const int ix = (int)xtra1, jy = (int)ytra1, ixp = ix + 1, jyp = jy + 1;
    const float ddx = (float)(xtra1 - (double)(float)ix), ddy = (float)(ytra1 - (double)(float)jy);
    const float rddx = 1.f - ddx, rddy = 1.f - ddy;
    const float p1 = rddx * rddy, p2 = ddx * rddy, p3 = rddx * ddy, p4 = ddx * ddy;
    int indz = c->nz - 1;
    for (int il = 2; il <= c->nz; il++)
      if (height[il - 1] > ztra1) { indz = il - 1; break; }
    const int indzp = indz + 1;
    const float dz1 = ztra1 - height[indz - 1], dz2 = height[indzp - 1] - ztra1;
    const float dz = 1.f / (dz1 + dz2);
    float rhoprof[2];
    for (int ind = indz; ind <= indzp; ind++)
      rhoprof[ind - indz] = p1 * rho2[ix + nxm * jy + plane * (ind - 1)] + p2 * rho2[ixp + nxm * jy + plane * (ind - 1)] +
                            p3 * rho2[ix + nxm * jyp + plane * (ind - 1)] + p4 * rho2[ixp + nxm * jyp + plane * (ind - 1)];
    rhoi = (dz1 * rhoprof[1] + dz2 * rhoprof[0]) * dz;
  }
  const int nrelpointer = ((c->ioutputforeachrelease == 0) || (c->mdomainfill == 1)) ? 1 : npoint;
  int kz;
  for (kz = 1; kz <= c->numzgrid; kz++)
    if (c->outheight[kz - 1] > ztra1) break;
  if (kz > c->numzgrid) return;
  const float xl = (float)((xtra1 * (double)c->dx + (double)c->xoutshift) / (double)c->dxout);
  const float yl = (float)((ytra1 * (double)c->dy + (double)c->youtshift) / (double)c->dyout);
  int ix = (int)xl;
  if (xl < 0.f) ix = ix - 1;
  int jy = (int)yl;
  if (yl < 0.f) jy = jy - 1;
  const size_t nxg = c->numxgrid, nyg = c->numygrid, nzg = c->numzgrid;
#define IC(cx, cy, ks) init_cond[(cx) + nxg * ((cy) + nyg * ((kz - 1) + nzg * (((ks)-1) + (size_t)c->maxspec * (nrelpointer - 1))))]
  if ((xl < 0.5f) || (yl < 0.5f) || (xl > (float)(c->numxgrid - 1) - 0.5f) || (yl > (float)(c->numygrid - 1) - 0.5f)) {
    if ((ix >= 0) && (jy >= 0) && (ix <= c->numxgrid - 1) && (jy <= c->numygrid - 1))
      for (int ks = 1; ks <= c->nspec; ks++) IC(ix, jy, ks) = IC(ix, jy, ks) + mass[ks - 1] / rhoi;
  } else {
    const float ddx = xl - (float)ix, ddy = yl - (float)jy;
    float wx, wy, w;
    int ixp, jyp;
    if (ddx > 0.5f) { ixp = ix + 1; wx = 1.5f - ddx; } else { ixp = ix - 1; wx = 0.5f + ddx; }
    if (ddy > 0.5f) { jyp = jy + 1; wy = 1.5f - ddy; } else { jyp = jy - 1; wy = 0.5f + ddy; }
    if ((ix >= 0) && (ix <= c->numxgrid - 1)) {
      if ((jy >= 0) && (jy <= c->numygrid - 1)) {
        w = wx * wy;
        for (int ks = 1; ks <= c->nspec; ks++) IC(ix, jy, ks) = IC(ix, jy, ks) + mass[ks - 1] / rhoi * w;
      }
      if ((jyp >= 0) && (jyp <= c->numygrid - 1)) {
        w = wx * (1.f - wy);
        for (int ks = 1; ks <= c->nspec; ks++) IC(ix, jyp, ks) = IC(ix, jyp, ks) + mass[ks - 1] / rhoi * w;
      }
    }
    if ((ixp >= 0) && (ixp <= c->numxgrid - 1)) {
      if ((jyp >= 0) && (jyp <= c->numygrid - 1)) {
        w = (1.f - wx) * (1.f - wy);
        for (int ks = 1; ks <= c->nspec; ks++) IC(ixp, jyp, ks) = IC(ixp, jyp, ks) + mass[ks - 1] / rhoi * w;
      }
      if ((jy >= 0) && (jy <= c->numygrid - 1)) {
        w = (1.f - wx) * wy;
        for (int ks = 1; ks <= c->nspec; ks++) IC(ixp, jy, ks) = IC(ixp, jy, ks) + mass[ks - 1] / rhoi * w;
      }
    }
  }
#undef IC
}
