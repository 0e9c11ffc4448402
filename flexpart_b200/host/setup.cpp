// setup.cpp -- run-constant derivation on the host: what gridcheck_ecmwf,
// readcommand and readoutgrid leave in com_mod / outg_mod for the hot path.
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>

#include "fpbh_internal.h"

static thread_local std::string g_err;
int fpbh_fail(const char *fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  return 1;
}
extern "C" const char *fpbh_last_error(void) { return g_err.c_str(); }

// src/gridcheck_ecmwf.f90:300-366; par_mod: r_earth=6.371e6, pi=3.14159265,
// switchnorth=75, switchsouth=-75 (src/par_mod.f90:63,123)
extern "C" int fpbh_gridcheck(fpb_config *c) {
  if (!c) return fpbh_fail("fpbh_gridcheck: null config");
  if (c->nx < 2 || c->ny < 2 || c->nz < 2) return fpbh_fail("fpbh_gridcheck: grid too small");
  if (c->nxmax < c->nx || c->nymax < c->ny || c->nzmax < c->nz)
    return fpbh_fail("fpbh_gridcheck: padded extents smaller than the grid");
  const float r_earth = 6.371e6f, pi = 3.14159265f;
  const float switchnorth = 75.f, switchsouth = -75.f;
  c->dxconst = 180.f / (c->dx * r_earth * pi);
  c->dyconst = 180.f / (c->dy * r_earth * pi);
  // cyclic if the last column repeats the first one (nx = nxfield+1)
  const float xspan = c->dx * (float)(c->nx - 1);
  c->xglobal = std::fabs(xspan - 360.f) < 0.001f ? 1 : 0;
  c->nxmin1 = c->nx - 1;
  c->nymin1 = c->ny - 1;
  if (c->xlon0 > 180.f) c->xlon0 = c->xlon0 - 360.f;
  const float yaux1 = c->ylat0, yaux2 = c->ylat0 + c->dy * (float)(c->ny - 1);
  if (c->xglobal && std::fabs(yaux1 + 90.f) < 0.001f) {
    c->sglobal = 1;
    const float sizesouth = 6.f * (switchsouth + 90.f) / c->dy;
    cmap::stlmbr(c->southpolemap, -90.f, 0.f);
    cmap::stcm2p(c->southpolemap, 0.f, 0.f, switchsouth, 0.f, sizesouth, sizesouth, switchsouth, 180.f);
    c->switchsouthg = (switchsouth - c->ylat0) / c->dy;
  } else {
    c->sglobal = 0;
    c->switchsouthg = 999999.f;
  }
  if (c->xglobal && std::fabs(yaux2 - 90.f) < 0.001f) {
    c->nglobal = 1;
    const float sizenorth = 6.f * (90.f - switchnorth) / c->dy;
    cmap::stlmbr(c->northpolemap, 90.f, 0.f);
    cmap::stcm2p(c->northpolemap, 0.f, 0.f, switchnorth, 0.f, sizenorth, sizenorth, switchnorth, 180.f);
    c->switchnorthg = (switchnorth - c->ylat0) / c->dy;
  } else {
    c->nglobal = 0;
    c->switchnorthg = 999999.f;
  }
  c->eps = (float)c->nxmax / 3.e5f; // src/advance.f90:107
  return 0;
}

// gridcheck_nests for nest l = c->numbnests+1 (src/gridcheck_nests.f90:359-389):
// resolution ratios and the nest's borders in mother-grid coordinates
extern "C" int fpbh_gridcheck_nest(fpb_config *c, float xlon0n, float ylat0n, int32_t nxn, int32_t nyn,
                                   float dxn, float dyn) {
  if (!c) return fpbh_fail("fpbh_gridcheck_nest: null config");
  if (c->numbnests >= FPB_MAXNESTS) return fpbh_fail("fpbh_gridcheck_nest: more than FPB_MAXNESTS nests");
  if (nxn < 2 || nyn < 2 || dxn <= 0.f || dyn <= 0.f) return fpbh_fail("fpbh_gridcheck_nest: bad nest grid");
  const int l = c->numbnests;
  c->xresoln[l] = c->dx / dxn;
  c->yresoln[l] = c->dy / dyn;
  const float xaux1 = xlon0n, xaux2 = xlon0n + (float)(nxn - 1) * dxn;
  const float yaux1 = ylat0n, yaux2 = ylat0n + (float)(nyn - 1) * dyn;
  c->xln[l] = (xaux1 - c->xlon0) / c->dx;
  c->xrn[l] = (xaux2 - c->xlon0) / c->dx;
  c->yln[l] = (yaux1 - c->ylat0) / c->dy;
  c->yrn[l] = (yaux2 - c->ylat0) / c->dy;
  if ((c->xln[l] < 0.f) || (c->yln[l] < 0.f) || (c->xrn[l] > (float)c->nxmin1) || (c->yrn[l] > (float)c->nymin1))
    return fpbh_fail("Nested domain does not fit into mother domain");
  c->nxn[l] = nxn;
  c->nyn[l] = nyn;
  if (nxn > c->nxmaxn) c->nxmaxn = nxn;
  if (nyn > c->nymaxn) c->nymaxn = nyn;
  c->numbnests = l + 1;
  return 0;
}

// src/readcommand.f90:244-272 (turbulence switches), :377-383 (method),
// :627-634 (backward runs); maxtl=1200 (src/com_mod.f90)
extern "C" int fpbh_readcommand(fpb_config *c) {
  if (!c) return fpbh_fail("fpbh_readcommand: null config");
  if (c->ldirect != 1 && c->ldirect != -1)
    return fpbh_fail("DIRECTION IN FILE \"COMMAND\" MUST BE EITHER -1 OR 1.");
  if (c->lsynctime <= 0) return fpbh_fail("fpbh_readcommand: LSYNCTIME must be given positive");
  if (c->ctl == 0.f) return fpbh_fail("fpbh_readcommand: CTL must not be 0");
  const int maxtl = 1200;
  c->ifine = c->ifine > 1 ? c->ifine : 1;
  if (c->cblflag == 1) {
    c->turbswitch = 1;
    if (c->lsynctime > maxtl) c->lsynctime = maxtl;
    if (c->ctl < 5.f) c->ctl = 5.f;
    if ((float)c->ifine * c->ctl < 50.f) c->ifine = (int)(50.f / c->ctl) + 1;
  } else if (c->ctl >= 0.1f) {
    c->turbswitch = 1;
  } else {
    c->turbswitch = 0;
    c->ifine = 1;
  }
  c->fine = 1.f / (float)c->ifine;
  c->ctl = 1.f / c->ctl;
  // method / mintime are derived while lsynctime is still positive (src/readcommand.f90:377-383);
  // the sign flip of a backward run comes later (:627-634), so mintime stays +|lsynctime|
  if (c->ctl > 0.f) {
    c->method = 1;
    c->mintime = 1; // par_mod minstep
  } else {
    c->method = 0;
    c->mintime = c->lsynctime;
  }
  if (c->ldirect == -1) c->lsynctime = -c->lsynctime;
  if (c->d_trop == 0.f) c->d_trop = 50.f;     // src/par_mod.f90:79
  if (c->d_strat == 0.f) c->d_strat = 0.1f;
  if (c->turbmesoscale == 0.f) c->turbmesoscale = 0.16f;
  return 0;
}

// src/readoutgrid.f90:199-200
extern "C" int fpbh_readoutgrid(fpb_config *c, float outlon0, float outlat0, int32_t numxgrid,
                                int32_t numygrid, float dxout, float dyout, const float *outheights,
                                int32_t numzgrid) {
  if (!c || !outheights) return fpbh_fail("fpbh_readoutgrid: null argument");
  if (numzgrid < 1 || numzgrid > FPB_MAXZGRID) return fpbh_fail("fpbh_readoutgrid: numzgrid out of range");
  c->numxgrid = numxgrid; c->numygrid = numygrid; c->numzgrid = numzgrid;
  c->dxout = dxout; c->dyout = dyout;
  for (int k = 0; k < numzgrid; k++) c->outheight[k] = outheights[k];
  c->xoutshift = c->xlon0 - outlon0;
  c->youtshift = c->ylat0 - outlat0;
  return 0;
}

// src/readoutgrid_nest.f90:111-112
extern "C" int fpbh_readoutgrid_nest(fpb_config *c, float outlon0n, float outlat0n, int32_t numxgridn,
                                     int32_t numygridn, float dxoutn, float dyoutn) {
  if (!c) return fpbh_fail("fpbh_readoutgrid_nest: null argument");
  c->numxgridn = numxgridn; c->numygridn = numygridn;
  c->dxoutn = dxoutn; c->dyoutn = dyoutn;
  c->xoutshiftn = c->xlon0 - outlon0n;
  c->youtshiftn = c->ylat0 - outlat0n;
  c->nested_output = 1;
  return 0;
}

// src/outgrid_init.f90:48-100 (area, volume; the wall areas are only used by the flux output)
extern "C" int fpbh_outgrid_geometry(const fpb_config *cp, int32_t nest, float outlat0, float *area, float *volume) {
  if (!cp || !area || !volume) return fpbh_fail("fpbh_outgrid_geometry: null argument");
  const fpb_config &c = *cp;
  if (nest && c.nested_output != 1) return fpbh_fail("fpbh_outgrid_geometry: no nested output grid");
  const float r_earth = 6.371e6f, pi = 3.14159265f, pi180 = pi / 180.f; // src/par_mod.f90:59-60
  const int nx = nest ? c.numxgridn : c.numxgrid, ny = nest ? c.numygridn : c.numygrid;
  const float dyo = nest ? c.dyoutn : c.dyout, dxo = nest ? c.dxoutn : c.dxout;
  for (int jy = 0; jy < ny; jy++) {
    const float ylat = outlat0 + ((float)jy + 0.5f) * dyo;
    const float ylatp = ylat + 0.5f * dyo, ylatm = ylat - 0.5f * dyo;
    float hzone;
    if (ylatm < 0.f && ylatp > 0.f) {
      hzone = dyo * r_earth * pi180;
    } else { // zone height between two latitude circles, M = 2*pi*R*h*dx/360
      const float cp_ = std::cos(ylatp * pi180), cm_ = std::cos(ylatm * pi180);
      hzone = (cp_ < cm_) ? std::sqrt(1.f - cp_ * cp_) - std::sqrt(1.f - cm_ * cm_)
                          : std::sqrt(1.f - cm_ * cm_) - std::sqrt(1.f - cp_ * cp_);
      hzone = hzone * r_earth;
    }
    const float gridarea = 2.f * pi * r_earth * hzone * dxo / 360.f;
    for (int ix = 0; ix < nx; ix++) {
      area[ix + (size_t)nx * jy] = gridarea;
      volume[ix + (size_t)nx * jy] = gridarea * c.outheight[0];
      for (int kz = 2; kz <= c.numzgrid; kz++)
        volume[ix + (size_t)nx * (jy + (size_t)ny * (kz - 1))] = gridarea * (c.outheight[kz - 1] - c.outheight[kz - 2]);
    }
  }
  return 0;
}
