// CPU-side cross-check of flexpart_b200/csrc/fpb_metproc.cuh (the device code of calcpar +
// verttransform_ecmwf, compiled here for the host) against the reference's own routines
// (oracle/_ref/libflexref.so).  Test infrastructure: built and used by tests/test_metproc.py only.
#include <vector_types.h>
static inline float2 make_float2(float x, float y) { float2 v; v.x = x; v.y = y; return v; }
static inline float4 make_float4(float x, float y, float z, float w) { float4 v; v.x = x; v.y = y; v.z = z; v.w = w; return v; }

#include "../flexpart_b200/csrc/fpb_metproc.cuh"

using namespace fpbmet;

extern "C" unsigned long met_check_sizeof_grid() { return sizeof(MetGrid); }

// the passes in the order fpb_metproc_launch runs them
extern "C" void met_check_run(const MetGrid *g) {
  for (int jy = 0; jy < g->ny; jy++)
    for (int ix = 0; ix < g->nx; ix++) met_levels_column(*g, ix, jy);
  for (int jy = 0; jy < g->ny; jy++)
    for (int ix = 0; ix < g->nx; ix++) met_calcpar_column(*g, ix, jy);
  if (g->theta) {
    for (int jy = 0; jy < g->ny; jy++)
      for (int ix = 0; ix < g->nx; ix++) met_theta_column(*g, ix, jy);
    for (int jy = 0; jy < g->ny; jy++)
      for (int ix = 0; ix < g->nx; ix++) met_calcpv_column(*g, ix, jy);
    if (g->nglobal || g->sglobal)
      for (int kl = 1; kl <= g->nuvz; kl++) met_pv_pole_level(*g, kl);
  }
  for (int jy = 0; jy < g->ny; jy++)
    for (int ix = 0; ix < g->nx; ix++) met_interp_column(*g, ix, jy);
  if (g->nglobal || g->sglobal)
    for (int iz = 1; iz <= g->nz; iz++) met_pole_level(*g, iz);
}
