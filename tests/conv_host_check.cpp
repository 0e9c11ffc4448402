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

// The same column with the loops over level pairs decomposed as conv_mix_kernel / conv_assembly_kernel decompose them:
// the O(n) head, then the rows of MENT in the order of `rows` residue classes taken LAST CLASS FIRST and each class from
// the top row down (any order must give the same bits: a row depends on the column's vectors and on itself alone), the
// flux assembly, nconvtop from the per-row tops, and the rows of the redistribution matrix again last-first.
// Returns lconv; fmassfrac / sub / nconvtop / cbmf as conv_check_column.
extern "C" int conv_check_column_rows(int nuvz, int nconvlev, const float *akz, const float *bkz, const float *akm,
                                      const float *bkm, const float *tconv, const float *qconv, float psconv, float delt,
                                      float *cbmf, int rows, int *nconvtop, float *fmassfrac /* [ld*ld] */, float *sub) {
  const size_t n = conv_pool_floats(nuvz, nconvlev);
  float *pool = (float *)malloc(n * sizeof(float));
  for (size_t k = 0; k < n; k++) pool[k] = -123.25f; // (nothing may depend on the pool's initial content ...)
  ConvWork w;
  conv_carve(w, pool, nuvz, nconvlev);
  const size_t nvec = (size_t)(CONV_NVEC + 1) * (nuvz + 4);
  for (size_t k = 0; k < nvec; k++) w.pconv[k] = 0.f; // (... but the vectors: the reference's zero-initialised locals)
  w.akz = akz; w.bkz = bkz; w.akm = akm; w.bkm = bkm;
  for (int k = 1; k <= nuvz - 1; k++) { w.tconv[k] = tconv[k]; w.qconv[k] = qconv[k]; }
  w.psconv = psconv; w.tt2conv = 280.f; w.td2conv = 275.f;
  ConvState st;
  conv_calcmatrix_a(w, delt, *cbmf, st, true);
  if (st.go) {
    for (int y = rows - 1; y >= 0; y--) {
      int last = st.icb + 1 + y;
      while (last + rows <= st.inb) last += rows;
      for (int i = last; i >= st.icb + 1; i -= rows) conv_mixnorm_row(w, st, i);
    }
    conv_flux_assembly(w, st);
  }
  const bool lconv = conv_calcmatrix_b(w, delt, *cbmf, st, false, true);
  *nconvtop = w.nconvtop;
  if (lconv) {
    for (int kq = w.nconvtop; kq >= 1; kq--)
      conv_fmass_row(w, st, delt, kq, [&](int i, int j) { return w.ment[(size_t)i + (size_t)w.ld * j]; });
    memcpy(fmassfrac, w.fmass, sizeof(float) * w.ld * w.ld);
    memcpy(sub, w.sub, sizeof(float) * (nuvz + 2));
  }
  free(pool);
  return lconv ? 1 : 0;
}
