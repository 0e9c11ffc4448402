// fpb_metproc.cu -- calcpar + verttransform_ecmwf on the device: the
// wind field as readwind leaves it becomes the met slot the particle loop reads, without the
// transformed fields ever existing on the host.  One thread per grid column (x fastest: the loads of
// a warp are 32 neighbouring columns of one level, coalesced), three passes because the slope term
// of the vertical wind reads the neighbours' level heights and the pole rows read whole rows:
//   met_levels_kernel    heights of the eta levels                       (2 x 138 planes read, 1 written)
//   met_calcpar_kernel   ustar, oli, hmix, wstar, tropopause             (column walks, ~3 passes)
//   met_theta_kernel / met_calcpv_kernel / met_pv_pole_kernel   potential vorticity on the eta levels
//   met_interp_kernel    the fields on the height levels + clouds        (raw field read once, slot written once)
//   met_pole_kernel      pole rows: one block per level
// Compiled with --fmad=false; see fpb_metproc.cuh for the arithmetic contract.
#include "fpb_metproc.cuh"

#include <cuda_runtime.h>

using namespace fpbmet;

namespace {
__global__ void __launch_bounds__(128) met_levels_kernel(const __grid_constant__ MetGrid g) {
  const int ix = blockIdx.x * blockDim.x + threadIdx.x, jy = blockIdx.y;
  if (ix < g.nx && jy < g.ny) met_levels_column(g, ix, jy);
}
__global__ void __launch_bounds__(128) met_calcpar_kernel(const __grid_constant__ MetGrid g) {
  const int ix = blockIdx.x * blockDim.x + threadIdx.x, jy = blockIdx.y;
  if (ix < g.nx && jy < g.ny) met_calcpar_column(g, ix, jy);
}
__global__ void __launch_bounds__(128) met_theta_kernel(const __grid_constant__ MetGrid g) {
  const int ix = blockIdx.x * blockDim.x + threadIdx.x, jy = blockIdx.y;
  if (ix < g.nx && jy < g.ny) met_theta_column(g, ix, jy);
}
__global__ void __launch_bounds__(128) met_calcpv_kernel(const __grid_constant__ MetGrid g) {
  const int ix = blockIdx.x * blockDim.x + threadIdx.x, jy = blockIdx.y;
  if (ix < g.nx && jy < g.ny) met_calcpv_column(g, ix, jy);
}
__global__ void __launch_bounds__(32) met_pv_pole_kernel(const __grid_constant__ MetGrid g) {
  const int kl = blockIdx.x * blockDim.x + threadIdx.x + 1;
  if (kl <= g.nuvz) met_pv_pole_level(g, kl);
}
__global__ void __launch_bounds__(128) met_interp_kernel(const __grid_constant__ MetGrid g) {
  const int ix = blockIdx.x * blockDim.x + threadIdx.x, jy = blockIdx.y;
  if (ix < g.nx && jy < g.ny) met_interp_column(g, ix, jy);
}
// the zonal sum of the reference is sequential in ix: one thread per level does the row
__global__ void __launch_bounds__(32) met_pole_kernel(const __grid_constant__ MetGrid g) {
  const int iz = blockIdx.x * blockDim.x + threadIdx.x + 1;
  if (iz <= g.nz) met_pole_level(g, iz);
}
} // namespace

void fpb_metproc_launch(const MetGrid &g, cudaStream_t st, int64_t *launches) {
  const dim3 grid((g.nx + 127) / 128, g.ny);
  met_levels_kernel<<<grid, 128, 0, st>>>(g);
  met_calcpar_kernel<<<grid, 128, 0, st>>>(g);
  if (g.theta) { // calcpv (src/calcpar.f90:266): pvh is computed here, not handed in
    met_theta_kernel<<<grid, 128, 0, st>>>(g);
    met_calcpv_kernel<<<grid, 128, 0, st>>>(g);
    if (g.nglobal || g.sglobal) met_pv_pole_kernel<<<(g.nuvz + 31) / 32, 32, 0, st>>>(g);
    if (launches) *launches += (g.nglobal || g.sglobal) ? 3 : 2;
  }
  met_interp_kernel<<<grid, 128, 0, st>>>(g);
  if (g.nglobal || g.sglobal) met_pole_kernel<<<(g.nz + 31) / 32, 32, 0, st>>>(g);
  if (launches) *launches += (g.nglobal || g.sglobal) ? 4 : 3;
}
