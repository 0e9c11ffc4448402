/* fpo_output.c -- TEST INFRASTRUCTURE (see fpo.h).  Sequential restatement of
 *   outgrid_init's cell areas and volumes   (src/outgrid_init.f90:48-100; nest: outgrid_init_nest.f90)
 *   concoutput's loop body for one (ks, kp, nage): class mean (src/mean_mod.f90:20-74),
 *   unit conversion and sparse dump          (src/concoutput.f90:225-235,287-475)
 * on grids in the reference layout (as fpb_fetch_grids / fpo_fetch_grids return them). */
#include <math.h>
#include <stdlib.h>

#include "fpo.h"
#include "fpo_math.h"

#define R_EARTH 6.371e6f
#define PI_F 3.14159265f
#define PI180 (PI_F / 180.f)

/* src/outgrid_init.f90:52-99 */
void fpo_outgrid_geometry(const fpb_config *c, int nest, float outlat0, float *area, float *volume) {
  const int nx = nest ? c->numxgridn : c->numxgrid, ny = nest ? c->numygridn : c->numygrid;
  const float lat0 = outlat0, dyo = nest ? c->dyoutn : c->dyout, dxo = nest ? c->dxoutn : c->dxout;
  for (int jy = 0; jy < ny; jy++) {
    float ylat = lat0 + ((float)jy + 0.5f) * dyo;
    float ylatp = ylat + 0.5f * dyo;
    float ylatm = ylat - 0.5f * dyo;
    float hzone;
    if ((ylatm < 0.f) && (ylatp > 0.f)) {
      hzone = dyo * R_EARTH * PI180;
    } else {
      float cosfactp = fpo_cosf(ylatp * PI180);
      float cosfactm = fpo_cosf(ylatm * PI180);
      if (cosfactp < cosfactm) {
        hzone = fpo_sqrtf(1.f - cosfactp * cosfactp) - fpo_sqrtf(1.f - cosfactm * cosfactm);
        hzone = hzone * R_EARTH;
      } else {
        hzone = fpo_sqrtf(1.f - cosfactm * cosfactm) - fpo_sqrtf(1.f - cosfactp * cosfactp);
        hzone = hzone * R_EARTH;
      }
    }
    float gridarea = 2.f * PI_F * R_EARTH * hzone * dxo / 360.f;
    for (int ix = 0; ix < nx; ix++) {
      area[ix + nx * jy] = gridarea;
      volume[ix + nx * jy] = area[ix + nx * jy] * c->outheight[0];
      for (int kz = 2; kz <= c->numzgrid; kz++)
        volume[ix + nx * (jy + (size_t)ny * (kz - 1))] =
            area[ix + nx * jy] * (c->outheight[kz - 1] - c->outheight[kz - 2]);
    }
  }
}

/* mean_sp, src/mean_mod.f90:20-74 */
static void mean_sp(const float *x, float *xm, float *xs, int number) {
  const float eps = 1.0e-30f;
  float xl = 0.f, xq = 0.f;
  for (int i = 0; i < number; i++) {
    xl = xl + x[i];
    xq = xq + x[i] * x[i];
  }
  *xm = xl / (float)number;
  float xaux = xq - xl * xl / (float)number;
  if (xaux < eps) *xs = 0.f;
  else *xs = fpo_sqrtf(xaux / (float)(number - 1));
}

/* densityoutgrid, src/concoutput.f90:164-190: rho(0:nxmax-1,0:nymax-1,nzmax) of time level memind(2) */
void fpo_density_outgrid(const fpb_config *c, const float *height, int nest, float outlon0, float outlat0,
                         const float *rho, float *densityoutgrid) {
  const int nx = nest ? c->numxgridn : c->numxgrid, ny = nest ? c->numygridn : c->numygrid;
  const float dxo = nest ? c->dxoutn : c->dxout, dyo = nest ? c->dyoutn : c->dyout;
  for (int kz = 1; kz <= c->numzgrid; kz++) {
    float halfheight;
    if (kz == 1) halfheight = c->outheight[0] / 2.f;
    else halfheight = (c->outheight[kz - 1] + c->outheight[kz - 2]) / 2.f;
    int kzz;
    for (kzz = 2; kzz <= c->nz; kzz++)
      if ((height[kzz - 2] < halfheight) && (height[kzz - 1] > halfheight)) break;
    kzz = kzz < c->nz ? kzz : c->nz;
    kzz = kzz > 2 ? kzz : 2;
    float dz1 = halfheight - height[kzz - 2];
    float dz2 = height[kzz - 1] - halfheight;
    float dz = dz1 + dz2;
    for (int jy = 0; jy < ny; jy++)
      for (int ix = 0; ix < nx; ix++) {
        float xl = outlon0 + (float)ix * dxo;
        float yl = outlat0 + (float)jy * dyo;
        xl = (xl - c->xlon0) / c->dx;
        yl = (yl - c->ylat0) / c->dy;
        int iix = fpo_nint_f(xl), jjy = fpo_nint_f(yl);
        iix = iix < c->nxmin1 ? iix : c->nxmin1; iix = iix > 0 ? iix : 0;
        jjy = jjy < c->nymin1 ? jjy : c->nymin1; jjy = jjy > 0 ? jjy : 0;
        const size_t plane = (size_t)c->nxmax * c->nymax;
        densityoutgrid[ix + (size_t)nx * (jy + (size_t)ny * (kz - 1))] =
            (rho[iix + (size_t)c->nxmax * jjy + plane * (kzz - 1)] * dz1 +
             rho[iix + (size_t)c->nxmax * jjy + plane * (kzz - 2)] * dz2) / dz;
      }
  }
}

/* which: 0 concentration (grid = gridunc, geom = volume), 1 dry / 2 wet deposition
 * (grid = drygridunc / wetgridunc, geom = area), 3 mixing ratio (grid = gridunc, geom = volume,
 * density = densityoutgrid, tot_mu = weightmolar(ks); src/concoutput.f90:563-593) */
void fpo_concoutput_sparse(const fpb_config *c, int nest, int which, const float *grid_ref,
                           const float *geom, const float *density, int ks, int kp, int nage, float outnum, float tot_mu,
                           int loutaver, int32_t *sp_count_i, int32_t *sparse_dump_i,
                           int32_t *sp_count_r, float *sparse_dump_r) {
  const int nx = nest ? c->numxgridn : c->numxgrid, ny = nest ? c->numygridn : c->numygrid;
  const int conc = which == 0 || which == 3;
  const int nzg = conc ? c->numzgrid : 1;
  const float smallnum = 1.17549435e-38f; /* tiny(0.0) */
  const size_t cells = (size_t)nx * ny * nzg;
  float *grid = (float *)malloc(cells * sizeof(float));
  float *aux = (float *)malloc((size_t)c->nclassunc * sizeof(float));
  /* :287-340: mean over the classes, times the number of classes */
  for (int jy = 0; jy < ny; jy++)
    for (int ix = 0; ix < nx; ix++)
      for (int kz = 1; kz <= nzg; kz++) {
        const size_t cell = ix + (size_t)nx * (jy + (size_t)ny * (kz - 1));
        for (int l = 1; l <= c->nclassunc; l++) {
          size_t o = (size_t)(nage - 1);
          o = o * c->nclassunc + (l - 1);
          o = o * c->maxpointspec_act + (kp - 1);
          o = o * c->maxspec + (ks - 1);
          aux[l - 1] = grid_ref[o * cells + cell];
        }
        float xm, xs;
        mean_sp(aux, &xm, &xs, c->nclassunc);
        grid[cell] = xm * (float)c->nclassunc;
      }
  /* :352-475 */
  int ci = 0, cr = 0;
  float sp_fact = -1.f;
  int sp_zer = 1;
  for (int kz = 1; kz <= nzg; kz++)
    for (int jy = 0; jy < ny; jy++)
      for (int ix = 0; ix < nx; ix++) {
        const size_t cell = ix + (size_t)nx * (jy + (size_t)ny * (kz - 1));
        if (grid[cell] > smallnum) {
          if (sp_zer) {
            ci++;
            sparse_dump_i[ci - 1] = conc ? ix + jy * nx + kz * nx * ny : ix + jy * nx;
            sp_zer = 0;
            sp_fact = sp_fact * (-1.f);
          }
          cr++;
          if (which == 0) {
            /* factor3d, :225-235 */
            float factor3d = (c->ldirect == 1) ? 1.e12f / geom[cell] / outnum
                                               : (float)abs(loutaver) / outnum;
            sparse_dump_r[cr - 1] = sp_fact * grid[cell] * factor3d / tot_mu;
          } else if (which == 3) {
            const float weightair = 28.97f;
            sparse_dump_r[cr - 1] = sp_fact * 1.e12f * grid[cell] / geom[cell] / outnum * weightair / tot_mu / density[cell];
          } else {
            sparse_dump_r[cr - 1] = sp_fact * 1.e12f * grid[cell] / geom[cell];
          }
        } else {
          sp_zer = 1;
        }
      }
  *sp_count_i = ci;
  *sp_count_r = cr;
  free(grid);
  free(aux);
}

/* partoutput's record of one particle, src/partoutput.f90:70-183.  Field arrays in the com_mod
 * layout (0:nxmax-1,0:nymax-1,nzmax) of the time levels memind(1), memind(2). */
void fpo_partoutput_record(const fpb_config *c, const float *height, int itime, const int32_t memtime[2],
                           double xtra1, double ytra1, float ztra1, const float *oro,
                           const float *pv[2], const float *qv[2], const float *tt[2], const float *rho[2],
                           const float *hmix[2], const float *tropopause[2], float out[9]) {
  const size_t nxm = c->nxmax, plane = (size_t)c->nxmax * c->nymax;
  float dt1 = (float)(itime - memtime[0]), dt2 = (float)(memtime[1] - itime);
  float dtt = 1.f / (dt1 + dt2);
  float xlon = (float)(c->xlon0 + xtra1 * c->dx);
  float ylat = (float)(c->ylat0 + ytra1 * c->dy);
  int ix = (int)xtra1, jy = (int)ytra1;
  int ixp = ix + 1, jyp = jy + 1;
  float ddx = (float)(xtra1 - (float)ix), ddy = (float)(ytra1 - (float)jy);
  float rddx = 1.f - ddx, rddy = 1.f - ddy;
  float p1 = rddx * rddy, p2 = ddx * rddy, p3 = rddx * ddy, p4 = ddx * ddy;
  if (jyp >= c->nymax) jyp = jyp - 1;
#define F2(f) (p1 * (f)[ix + nxm * jy] + p2 * (f)[ixp + nxm * jy] + p3 * (f)[ix + nxm * jyp] + p4 * (f)[ixp + nxm * jyp])
#define F3(f, k) (p1 * (f)[ix + nxm * jy + plane * ((k)-1)] + p2 * (f)[ixp + nxm * jy + plane * ((k)-1)] + \
                  p3 * (f)[ix + nxm * jyp + plane * ((k)-1)] + p4 * (f)[ixp + nxm * jyp + plane * ((k)-1)])
  float topo = F2(oro);
  int indz = c->nz - 1, indzp;
  for (int il = 2; il <= c->nz; il++)
    if (height[il - 1] > ztra1) { indz = il - 1; break; }
  indzp = indz + 1;
  float dz1 = ztra1 - height[indz - 1], dz2 = height[indzp - 1] - ztra1;
  float dz = 1.f / (dz1 + dz2);
  float pvprof[2], qvprof[2], ttprof[2], rhoprof[2];
  for (int ind = indz; ind <= indzp; ind++) {
    float pv1[2], qv1[2], tt1[2], rho1[2];
    for (int m = 0; m < 2; m++) {
      pv1[m] = F3(pv[m], ind);
      qv1[m] = F3(qv[m], ind);
      tt1[m] = F3(tt[m], ind);
      rho1[m] = F3(rho[m], ind);
    }
    pvprof[ind - indz] = (pv1[0] * dt2 + pv1[1] * dt1) * dtt;
    qvprof[ind - indz] = (qv1[0] * dt2 + qv1[1] * dt1) * dtt;
    ttprof[ind - indz] = (tt1[0] * dt2 + tt1[1] * dt1) * dtt;
    rhoprof[ind - indz] = (rho1[0] * dt2 + rho1[1] * dt1) * dtt;
  }
  float pvi = (dz1 * pvprof[1] + dz2 * pvprof[0]) * dz;
  float qvi = (dz1 * qvprof[1] + dz2 * qvprof[0]) * dz;
  float tti = (dz1 * ttprof[1] + dz2 * ttprof[0]) * dz;
  float rhoi = (dz1 * rhoprof[1] + dz2 * rhoprof[0]) * dz;
  float tr[2], hm[2];
  for (int m = 0; m < 2; m++) {
    tr[m] = F2(tropopause[m]);
    hm[m] = F2(hmix[m]);
  }
  float hmixi = (hm[0] * dt2 + hm[1] * dt1) * dtt;
  float tri = (tr[0] * dt2 + tr[1] * dt1) * dtt;
#undef F2
#undef F3
  out[0] = xlon; out[1] = ylat; out[2] = topo; out[3] = pvi; out[4] = qvi; out[5] = rhoi;
  out[6] = hmixi; out[7] = tri; out[8] = tti;
}
