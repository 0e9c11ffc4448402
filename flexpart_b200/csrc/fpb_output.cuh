// fpb_output.cuh -- concoutput's per-(species, release, age class) work on the device
// (SURVEY.md section 8f, rank 4): class mean, unit conversion and the sparse run-length dump.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

#include "fpb_device.cuh"

struct SparseDumpArgs {
  const float *grid;        // first class slice of (ks, kp, nage): [ncells]
  size_t class_stride;      // floats between the slices of consecutive uncertainty classes
  int nclassunc, ncells;
  const float *geom;        // which = 0, 3: volume[ncells]; 1, 2: area[ncells]
  const float *density;     // which = 3: densityoutgrid[ncells]
  int which;                // 0 concentration, 1 dry deposition, 2 wet deposition, 3 mixing ratio
  int ldirect;
  float outnum, tot_mu, loutaver_abs;
  int index_offset;         // added to the linear cell index (numxgrid*numygrid for concentrations)
  unsigned *block_counts;   // [2 * nblocks] scratch: non-zero cells, run starts
  int32_t *out_i;           // sparse_dump_i
  float *out_r;             // sparse_dump_r
  int *counts;              // [0] sp_count_i, [1] sp_count_r
};

void fpb_sparse_dump(const SparseDumpArgs &a, cudaStream_t st);

// densityoutgrid of concoutput (src/concoutput.f90:164-190): rho of time level memind(2) at the
// met cell nearest to each output cell, interpolated to the middle of the output layer
struct DensityArgs {
  const float4 *A;          // {uu, vv, ww, rho} of memind(2), [level][jy][ix], row stride nxd
  int nxd, plane;
  int numx, numy, numz;
  float outlon0, outlat0, dxout, dyout, xlon0, ylat0, dx, dy;
  int nxmin1, nymin1;
  int kzz[32];              // per output layer: met level above its middle (Fortran index)
  float dz1[32], dz2[32];
  float *density;           // [numx * numy * numz]
};
void fpb_density_outgrid(const DensityArgs &a, cudaStream_t st);

// partoutput (src/partoutput.f90:66-192): one record per particle with itra1 == itime, in slot
// order; the fields of the two time levels interpolated to the particle
struct PartoutArgs {
  DevCfg cfg;               // cfg.itime, cfg.memtime set
  DevMetSlot met[2];        // memind(1), memind(2)
  const float2 *Q[2];       // {pv, qv} of the same time levels
  const float *oro;         // [nyd][nxd]
  const float *height;
  DevParticles p;
  const int32_t *row_of_slot;
  int permuted, numpart;
  unsigned *block_counts;
  int *count;               // records written
  // outputs, compacted in slot order
  int32_t *npoint, *itramem;
  float *xlon, *ylat, *ztra1, *topo, *pvi, *qvi, *rhoi, *hmixi, *tri, *tti;
  float *xmass1;            // [nspec][maxpart]
};
void fpb_partoutput_launch(const PartoutArgs &a, cudaStream_t st);

// The two optional hooks of the particle loop (src/timemanager.f90:614-623), run by fpb_step /
// fpb_step_host around the step kernels when fpb_config.iflux == 1 / ipout == 3:
//   calcfluxes       (src/calcfluxes.f90)       gross mass fluxes through the faces of the output grid
//                    cells a particle crosses during the step, flux(6, numxgrid, numygrid, numzgrid,
//                    nspec, maxpointspec_act, nageclass): needs the position and the masses from BEFORE
//                    the step (advance moves the particle; decay and deposition change the masses after
//                    the hook), which fpb_hooks_pre memorises per row
//   partpos_average  (src/partpos_average.f90)  running sums of position (Cartesian), height, topography,
//                    pv, qv, tt, uu, vv, rho, tropopause, hmix and energy per particle (= slot)
struct HookArgs {
  DevCfg cfg;               // cfg.itime, cfg.memtime set; cfg.numpart = rows of the view
  DevMetSlot met[2];        // memind(1), memind(2)
  const float2 *Q[2];       // {pv, qv} of the same time levels (averages only)
  const float *oro;         // [nyd][nxd] (averages only)
  const float *height;
  DevParticles p;           // row view
  int iflux, ipout3;
  int linit;                // linit_cond: initial_cond_calc for the rows the step terminated (0: off)
  int final_pass;           // 1: fpb_initial_cond_final -- every row with itra1 == itime, nothing else
  const int32_t *flags;     // the step's scratch flags (SC_TERM_*)
  float *init_cond;         // (numxgrid, numygrid, numzgrid, maxspec, maxpointspec_act)
  int maxspec;
  uint8_t *adv;             // [rows] 1 = the row is advanced by this step (itra1 == itime before it)
  float *old;               // [3 + nspec][old_stride]: xold, yold, zold, xmass1(ks) of the row before the step
  size_t old_stride;
  float *flux;
  int32_t *npart_av;        // [maxpart] by slot
  float *av;                // [14][av_stride] by slot: cartx, carty, cartz, z, topo, pv, qv, tt, uu, vv, rho, tro, hmix, energy
  size_t av_stride;
};
void fpb_hooks_pre(const HookArgs &a, cudaStream_t st);
void fpb_hooks_post(const HookArgs &a, cudaStream_t st);

#ifdef __CUDACC__
// src/calcfluxes.f90:55-166 for one particle: old position (default reals) -> new position; mass(k) = xmass1(jpart, k).
// The caller has worked out the age class (1..nageclass) and kp.  Used by the loop hook (fpb_output.cu) and by
// convmix's own call after redist (src/convmix.f90:205-218, fpb_convect.cu); compile without contraction.
template <class MassFn>
__device__ __forceinline__ void fpb_flux_particle(const DevCfg &c, float *flux, int nage, int kp, float xold, float yold,
                                                  float zold, double xt, double yt, float zt, MassFn mass) {
  const float xmean = (float)(((double)xold + xt) / 2.0);
  const float ymean = (float)(((double)yold + yt) / 2.0);
  const int ixave = (int)((xmean * c.dx + c.xoutshift) / c.dxout);
  const int jyave = (int)((ymean * c.dy + c.youtshift) / c.dyout);
  int kzave;
  for (kzave = 1; kzave <= c.numzgrid; kzave++)
    if (c.outheight[kzave - 1] > zt) break;
  const size_t n1 = 6, nxg = c.numxgrid, nyg = c.numygrid, nzg = c.numzgrid;
  auto add = [&](int i, int ix, int jy, int kz, int k) { // flux(i, ix, jy, kz, k, kp, nage) += xmass1(jpart, k)
    const size_t o = (i - 1) + n1 * (ix + nxg * (jy + nyg * ((kz - 1) + nzg * ((k - 1) + (size_t)c.nspec *
                     ((kp - 1) + (size_t)c.maxpointspec_act * (nage - 1))))));
    atomicAdd(flux + o, mass(k));
  };
  auto half = [&](int kz) { // outheighthalf, src/readoutgrid.f90:194-197
    return kz == 1 ? c.outheight[0] / 2.f : (c.outheight[kz - 2] + c.outheight[kz - 1]) / 2.f;
  };
  // vertical fluxes
  if (ixave >= 0 && jyave >= 0 && ixave <= c.numxgrid - 1 && jyave <= c.numygrid - 1) {
    int kz;
    for (kz = 1; kz <= c.numzgrid; kz++)
      if (half(kz) > zold) break;
    const int k1 = min(c.numzgrid, kz);
    for (kz = 1; kz <= c.numzgrid; kz++)
      if (half(kz) > zt) break;
    const int k2 = min(c.numzgrid, kz);
    for (int k = 1; k <= c.nspec; k++) {
      for (kz = k1; kz <= k2 - 1; kz++) add(5, ixave, jyave, kz, k);
      for (kz = k2; kz <= k1 - 1; kz++) add(6, ixave, jyave, kz, k);
    }
  }
  // west-east fluxes
  if (kzave <= c.numzgrid && jyave >= 0 && jyave <= c.numygrid - 1) {
    if (fabs((double)xold - xt) < (double)((float)c.nx / 2.f)) {
      const int ix1 = (int)((xold * c.dx + c.xoutshift) / c.dxout + 0.5f);
      const int ix2 = (int)((xt * (double)c.dx + (double)c.xoutshift) / (double)c.dxout + 0.5);
      for (int k = 1; k <= c.nspec; k++) {
        for (int ix = ix1; ix <= ix2 - 1; ix++)
          if (ix >= 0 && ix <= c.numxgrid - 1) add(1, ix, jyave, kzave, k);
        for (int ix = ix2; ix <= ix1 - 1; ix++)
          if (ix >= 0 && ix <= c.numxgrid - 1) add(2, ix, jyave, kzave, k);
      }
    } else { // the particle crossed the date line of a global domain
      const int ixs = (int)((((float)c.nxmin1 - 1.0e5f) * c.dx + c.xoutshift) / c.dxout);
      if (ixs >= 0 && ixs <= c.numxgrid - 1) {
        const int i = ((double)xold > xt) ? 1 : 2;
        for (int k = 1; k <= c.nspec; k++) add(i, ixs, jyave, kzave, k);
      }
    }
  }
  // south-north fluxes
  if (kzave <= c.numzgrid && ixave >= 0 && ixave <= c.numxgrid - 1) {
    const int jy1 = (int)((yold * c.dy + c.youtshift) / c.dyout + 0.5f);
    const int jy2 = (int)((yt * (double)c.dy + (double)c.youtshift) / (double)c.dyout + 0.5);
    for (int k = 1; k <= c.nspec; k++) {
      for (int jy = jy1; jy <= jy2 - 1; jy++)
        if (jy >= 0 && jy <= c.numygrid - 1) add(3, ixave, jy, kzave, k);
      for (int jy = jy2; jy <= jy1 - 1; jy++)
        if (jy >= 0 && jy <= c.numygrid - 1) add(4, ixave, jy, kzave, k);
    }
  }
}
#endif
