/*
 * fpo_vtable.c -- adapters that give the oracle the fpb_* call contract
 * (int return, engine pointer first) so the tests can run the SAME host time
 * loop (fpbh_timemanager) once on the oracle and once on the CUDA engine
 * (test infrastructure).
 */
#include "fpo.h"

int fpo_vt_upload_met(void *self, int32_t slot, const fpb_met_ptrs *met) {
  fpo_set_met((fpo_state *)self, slot, met);
  return 0;
}
int fpo_vt_set_met_bracket(void *self, const int32_t memind[2], const int32_t memtime[2],
                           int32_t lwindinterv) {
  int mi[2] = {memind[0], memind[1]}, mt[2] = {memtime[0], memtime[1]};
  fpo_set_met_bracket((fpo_state *)self, mi, mt, lwindinterv);
  return 0;
}
int fpo_vt_push_particles(void *self, int32_t first, int32_t count, const fpb_particle_ptrs *p) {
  fpo_push_particles((fpo_state *)self, first, count, p);
  return 0;
}
int fpo_vt_pull_particles(void *self, int32_t first, int32_t count, const fpb_particle_ptrs *p) {
  fpo_pull_particles((fpo_state *)self, first, count, p);
  return 0;
}
int fpo_vt_set_numpart(void *self, int32_t numpart) {
  fpo_set_numpart((fpo_state *)self, numpart);
  return 0;
}
int fpo_vt_step(void *self, int32_t itime, int32_t ldeltat, fpb_step_stats *stats) {
  fpo_step((fpo_state *)self, itime, ldeltat, stats);
  return 0;
}
int fpo_vt_conccalc(void *self, int32_t itime, float weight) {
  fpo_conccalc((fpo_state *)self, itime, weight);
  return 0;
}
int fpo_vt_fetch_grids(void *self, float *gridunc, float *griduncn, float *drygridunc,
                       float *drygriduncn, float *creceptor, int32_t zero_conc) {
  fpo_fetch_grids((fpo_state *)self, gridunc, griduncn, drygridunc, drygriduncn, creceptor, zero_conc);
  return 0;
}
int fpo_vt_scale_depgrids(void *self, const float *factor) {
  fpo_scale_depgrids((fpo_state *)self, factor);
  return 0;
}
int fpo_vt_wetdepo(void *self, int32_t itime, int32_t ltsample, int32_t ldeltat) {
  fpo_wetdepo((fpo_state *)self, itime, ltsample, ldeltat);
  return 0;
}

/* domain filling and splitting: the oracle keeps the itsplit of init_domainfill for boundcond */
static int vt_itsplit = 99999999;

int fpo_vt_init_domainfill(void *self, float xpoint1, float ypoint1, float xpoint2, float ypoint2, int32_t itsplit,
                           int32_t *numpart, void *info) {
  fpo_state *S = (fpo_state *)self;
  int32_t out[8];
  float fout[2];
  (void)info;
  vt_itsplit = itsplit;
  if (fpo_init_domainfill(S, xpoint1, ypoint1, xpoint2, ypoint2, itsplit, out, fout)) return 1;
  if (numpart) *numpart = S->numpart;
  return 0;
}

int fpo_vt_boundcond_domainfill(void *self, int32_t itime, int32_t loutend, int32_t *numpart, int32_t *n_created) {
  fpo_state *S = (fpo_state *)self;
  (void)loutend;
  const int n = fpo_boundcond_domainfill(S, itime, vt_itsplit);
  if (n < 0) return 1;
  if (numpart) *numpart = S->numpart;
  if (n_created) *n_created = n;
  return 0;
}

int fpo_vt_split_particles(void *self, int32_t itime, int32_t *numpart) {
  fpo_state *S = (fpo_state *)self;
  fpo_split_particles(S, itime);
  if (numpart) *numpart = S->numpart;
  return 0;
}
