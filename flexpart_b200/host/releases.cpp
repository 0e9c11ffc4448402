// releases.cpp -- releaseparticles on the host (src/releaseparticles.f90:69-378)
// with its ran1 position stream (src/random_mod.f90:12-42).  Scope: zkind 1
// (metres above ground), EMISVAR factors 1, ind_rel 0 -- the shipped
// options/RELEASES case; the integer semantics (release counts, first-free
// slot search from minpart, numpart update) are the parity contract.
#include <cmath>
#include <vector>

#include "fpbh_internal.h"

// Park-Miller "minimal standard" generator with Bays-Durham shuffle (ran1 of
// Numerical Recipes), state kept per release stream
struct fpbh_release_state {
  int idum = -7; // SAVEd idummy of releaseparticles
  int iv[32];
  int iy = 0;
  std::vector<float> xmasssave;
  float ran1() {
    const int IA = 16807, IM = 2147483647, IQ = 127773, IR = 2836, NTAB = 32;
    const int NDIV = 1 + (IM - 1) / NTAB;
    const float AM = 1.f / (float)IM, RNMX = 1.f - 1.2e-7f;
    if (idum <= 0 || iy == 0) {
      idum = (-idum > 1) ? -idum : 1;
      for (int j = NTAB + 7; j >= 0; j--) {
        int k = idum / IQ;
        idum = IA * (idum - k * IQ) - IR * k;
        if (idum < 0) idum += IM;
        if (j < NTAB) iv[j] = idum;
      }
      iy = iv[0];
    }
    int k = idum / IQ;
    idum = IA * (idum - k * IQ) - IR * k;
    if (idum < 0) idum += IM;
    int j = iy / NDIV;
    iy = iv[j];
    iv[j] = idum;
    float r = AM * (float)iy;
    return r < RNMX ? r : RNMX;
  }
};

extern "C" fpbh_release_state *fpbh_release_state_new(int32_t numpoint) {
  auto *s = new fpbh_release_state();
  s->xmasssave.assign(numpoint > 0 ? numpoint : 1, 0.f);
  return s;
}
extern "C" void fpbh_release_state_free(fpbh_release_state *s) { delete s; }
extern "C" void fpbh_release_state_set_rank(fpbh_release_state *s, int32_t mp_pid) {
  if (!s || mp_pid <= 0) return;
  const long long v = (244LL * 181LL) * ((long long)(mp_pid - 83) * 359LL); // Fortran mod: sign of the dividend
  long long m = v % 104729LL;
  if (m < 0) m = -m;
  s->idum = s->idum + (int)(-m);
}

extern "C" int fpbh_releaseparticles(const fpb_config *cp, const float *height,
                                     const fpbh_releases *rel, fpbh_release_state *st, int32_t itime,
                                     const fpb_particle_ptrs *p, int32_t *numpart,
                                     int32_t *first_changed, int32_t *n_changed) {
  if (!cp || !height || !rel || !st || !p || !numpart) return fpbh_fail("fpbh_releaseparticles: null argument");
  const fpb_config &c = *cp;
  const float eps2 = 1.e-6f;
  int lo = c.maxpart, hi = -1;
  int minpart = 0; // 0-based first candidate slot
  for (int i = 0; i < rel->numpoint; i++) {
    const int t0 = rel->ireleasestart[i], t1 = rel->ireleaseend[i];
    if (!(itime >= t0 && itime <= t1)) continue;
    int numrel;
    if (t0 != t1) {
      float rfraction = std::fabs((float)c.npart[i] * (float)c.lsynctime / (float)(t1 - t0));
      if (itime == t0 || itime == t1) rfraction = rfraction / 2.f;
      rfraction = rfraction * 1.f; // average_timecorrect
      rfraction = rfraction + st->xmasssave[i];
      numrel = (int)rfraction;
      st->xmasssave[i] = rfraction - (float)numrel;
    } else {
      numrel = c.npart[i];
    }
    const float xaux = rel->xpoint2[i] - rel->xpoint1[i];
    const float yaux = rel->ypoint2[i] - rel->ypoint1[i];
    const float zaux = rel->zpoint2[i] - rel->zpoint1[i];
    for (int j = 0; j < numrel; j++) {
      int ipart = minpart;
      while (ipart < c.maxpart && p->itra1[ipart] == itime) ipart++;
      if (ipart >= c.maxpart)
        return fpbh_fail("RELEASEPARTICLES: TOTAL NUMBER OF PARTICLES REQUIRED EXCEEDS THE MAXIMUM ALLOWED NUMBER.");
      double x = rel->xpoint1[i] + st->ran1() * xaux;
      if (c.xglobal) {
        if (x > (float)c.nxmin1) x = x - (float)c.nxmin1;
        if (x < 0.) x = x + (float)c.nxmin1;
      }
      p->xtra1[ipart] = x;
      p->ytra1[ipart] = rel->ypoint1[i] + st->ran1() * yaux;
      for (int k = 0; k < c.nspec; k++) {
        p->xmass1[(size_t)ipart + (size_t)p->ld * k] =
            c.xmass[i + (size_t)c.numpoint * k] / (float)c.npart[i] * 1.f / 1.f;
        if ((c.drybkdep || c.wetbkdep) && p->xscav_frac1)
          p->xscav_frac1[(size_t)ipart + (size_t)p->ld * k] = -1.f;
      }
      int nc = (int)(st->ran1() * (float)c.nclassunc) + 1;
      p->nclass[ipart] = nc < c.nclassunc ? nc : c.nclassunc;
      p->npoint[ipart] = i + 1;
      p->idt[ipart] = c.mintime;
      p->itra1[ipart] = itime;
      p->itramem[ipart] = itime;
      if (p->itrasplit) p->itrasplit[ipart] = itime + c.ldirect * rel->itsplit;
      float z = rel->zpoint1[i] + st->ran1() * zaux;
      if (z < eps2) z = eps2;
      if (z > height[c.nz - 1] - 0.5f) z = height[c.nz - 1] - 0.5f;
      p->ztra1[ipart] = z;
      // fresh slots start with zero turbulent memory (com_mod arrays are
      // zero-initialised; initialize() overwrites them on the first step)
      p->uap[ipart] = p->ucp[ipart] = p->uzp[ipart] = 0.f;
      p->us[ipart] = p->vs[ipart] = p->ws[ipart] = 0.f;
      p->cbt[ipart] = 1;
      if (ipart + 1 > *numpart) *numpart = ipart + 1;
      if (ipart < lo) lo = ipart;
      if (ipart > hi) hi = ipart;
      minpart = ipart + 1;
    }
  }
  if (first_changed) *first_changed = (hi >= 0) ? lo : 0;
  if (n_changed) *n_changed = (hi >= 0) ? hi - lo + 1 : 0;
  return 0;
}
