// CPU-side cross-check of flexpart_b200/csrc/fpb_convect.cuh (the device code of the convection
// column work, compiled here for the host) against the reference's own calcmatrix / convect / redist
// (oracle/_ref/libflexref.so).  Test infrastructure: built and used by tests/test_convection.py only.
#include <string.h>
#include <stdlib.h>

#include "../flexpart_b200/csrc/fpb_convect.cuh"

using namespace fpbconv;

extern "C" size_t conv_check_pool_floats(int nuvz, int nconvlev) { return conv_pool_floats(nuvz, nconvlev); }

// arrays are 1-based: element i at [i].  z[np] in/out; rn: uniforms consumed in particle order by
// the particles inside the convective domain (like ran3 in redist); *rn_used returns how many.
extern "C" int conv_check_column(int nuvz, int nconvlev, const float *akz, const float *bkz, const float *akm,
                                 const float *bkm, const float *tconv, const float *qconv, float psconv, float tt2conv,
                                 float td2conv, float delt, float *cbmf, int ldirect, int lsynctime, int np, float *z,
                                 const float *rn, int *rn_used, int *nconvtop, float *fmassfrac /* [ld*ld] */,
                                 float *sub, float *uvzlev, int *ld_out) {
  const size_t n = conv_pool_floats(nuvz, nconvlev);
  float *pool = (float *)calloc(n, sizeof(float));
  ConvWork w;
  conv_carve(w, pool, nuvz, nconvlev);
  w.akz = akz; w.bkz = bkz; w.akm = akm; w.bkm = bkm;
  for (int k = 1; k <= nuvz - 1; k++) { w.tconv[k] = tconv[k]; w.qconv[k] = qconv[k]; }
  w.psconv = psconv; w.tt2conv = tt2conv; w.td2conv = td2conv;
  const bool lconv = conv_calcmatrix(w, delt, *cbmf);
  *nconvtop = w.nconvtop;
  *ld_out = w.ld;
  int used = 0;
  if (lconv) {
    conv_uvzlev(w);
    for (int i = 0; i < np; i++) {
      const int levold = conv_levold(w, z[i]);
      if (levold > 0) z[i] = conv_redist(w, z[i], levold, rn[used++], ldirect, lsynctime);
    }
    memcpy(fmassfrac, w.fmass, sizeof(float) * w.ld * w.ld); // (stride 1 on the host)
    memcpy(sub, w.sub, sizeof(float) * (nuvz + 2));
    memcpy(uvzlev, w.uvzlev, sizeof(float) * (nuvz + 2));
  }
  *rn_used = used;
  free(pool);
  return lconv ? 1 : 0;
}
