/*
 * fpo_random.c -- oracle restatement of random_mod (test infrastructure).
 * Follows src/random_mod.f90 (ran1 :12-42, gasdev :45-67, gasdev1 :70-90,
 * ran3 :93-139; Numerical Recipes generators) and the table fill at
 * src/FLEXPART.f90:47,56-59.
 */
#include <stdlib.h>
#include <string.h>

#include "fpo.h"
#include "fpo_math.h"

/* src/random_mod.f90:12-42 -- Park-Miller with Bays-Durham shuffle */
float fpo_ran1(fpo_state *S, int *idum) {
  const int ia = 16807, im = 2147483647, iq = 127773, ir = 2836;
  const int ntab = 32, ndiv = 1 + (im - 1) / ntab;
  const float am = 1.f / (float)im, eps = 1.2e-7f, rnmx = 1.f - eps;
  int j, k;
  if (*idum <= 0 || S->r1_iy == 0) {
    *idum = (-*idum > 1) ? -*idum : 1;
    for (j = ntab + 8; j >= 1; j--) {
      k = *idum / iq;
      *idum = ia * (*idum - k * iq) - ir * k;
      if (*idum < 0) *idum += im;
      if (j <= ntab) S->r1_iv[j] = *idum;
    }
    S->r1_iy = S->r1_iv[1];
  }
  k = *idum / iq;
  *idum = ia * (*idum - k * iq) - ir * k;
  if (*idum < 0) *idum += im;
  j = 1 + S->r1_iy / ndiv;
  S->r1_iy = S->r1_iv[j];
  S->r1_iv[j] = *idum;
  float r = am * (float)S->r1_iy;
  return r < rnmx ? r : rnmx;
}

/* src/random_mod.f90:93-139 -- Knuth subtractive generator */
float fpo_ran3(fpo_state *S, int *idum) {
  const int mbig = 1000000000, mseed = 161803398, mz = 0;
  const float fac = 1.f / (float)mbig;
  int i, ii, k, mj, mk;
  int *ma = S->r3_ma;
  if (*idum < 0 || S->r3_iff == 0) {
    S->r3_iff = 1;
    mj = mseed - abs(*idum);
    mj = mj % mbig;
    ma[55] = mj;
    mk = 1;
    for (i = 1; i <= 54; i++) {
      ii = (21 * i) % 55;
      ma[ii] = mk;
      mk = mj - mk;
      if (mk < mz) mk += mbig;
      mj = ma[ii];
    }
    for (k = 1; k <= 4; k++)
      for (i = 1; i <= 55; i++) {
        ma[i] = ma[i] - ma[1 + (i + 30) % 55];
        if (ma[i] < mz) ma[i] += mbig;
      }
    S->r3_inext = 0;
    S->r3_inextp = 31;
    *idum = 1;
  }
  S->r3_inext++;
  if (S->r3_inext == 56) S->r3_inext = 1;
  S->r3_inextp++;
  if (S->r3_inextp == 56) S->r3_inextp = 1;
  mj = ma[S->r3_inext] - ma[S->r3_inextp];
  if (mj < mz) mj += mbig;
  ma[S->r3_inext] = mj;
  S->n_ran3_draws++;
  return (float)mj * fac;
}

/* the uniform that picks the rannumb index of an initialize / advance call: ran3, or the next
 * entry of the injected queue (validation hook, fpo.h) */
float fpo_index_uniform(fpo_state *S, int *idum) {
  if (S->index_queue && S->index_queue_pos < S->index_queue_n) return S->index_queue[S->index_queue_pos++];
  return fpo_ran3(S, idum);
}

void fpo_set_index_uniforms(fpo_state *S, const float *u, long n) {
  free(S->index_queue);
  S->index_queue = NULL;
  S->index_queue_n = S->index_queue_pos = 0;
  if (!u || n <= 0) return;
  S->index_queue = (float *)malloc((size_t)n * sizeof(float));
  memcpy(S->index_queue, u, (size_t)n * sizeof(float));
  S->index_queue_n = n;
}

/* src/random_mod.f90:45-67 */
float fpo_gasdev(fpo_state *S, int *idum) {
  if (S->gd_iset == 0) {
    float v1, v2, r, fac;
    do {
      v1 = 2.f * fpo_ran3(S, idum) - 1.f;
      v2 = 2.f * fpo_ran3(S, idum) - 1.f;
      r = v1 * v1 + v2 * v2;
    } while (r >= 1.0f || r == 0.0f);
    fac = fpo_sqrtf(-2.f * fpo_logf(r) / r);
    S->gd_gset = v1 * fac;
    S->gd_iset = 1;
    return v2 * fac;
  }
  S->gd_iset = 0;
  return S->gd_gset;
}

/* src/random_mod.f90:70-90 -- pair of normals clipped to [-3,3] */
void fpo_gasdev1(fpo_state *S, int *idum, float *random1, float *random2) {
  float v1, v2, r, fac;
  do {
    v1 = 2.f * fpo_ran3(S, idum) - 1.f;
    v2 = 2.f * fpo_ran3(S, idum) - 1.f;
    r = v1 * v1 + v2 * v2;
  } while (r >= 1.0f || r == 0.0f);
  fac = fpo_sqrtf(-2.f * fpo_logf(r) / r);
  *random1 = v1 * fac;
  *random2 = v2 * fac;
  if (*random1 < -3.f) *random1 = -3.f;
  if (*random2 < -3.f) *random2 = -3.f;
  if (*random1 > 3.f) *random1 = 3.f;
  if (*random2 > 3.f) *random2 = 3.f;
}

/* src/FLEXPART.f90:47,56-59: note the last call writes rannumb(maxrand) and
 * then rannumb(maxrand-1). */
void fpo_fill_rannumb(fpo_state *S, int maxrand, int idummy) {
  free(S->rannumb);
  S->rannumb = (float *)calloc((size_t)maxrand + 2, sizeof(float));
  S->maxrand = maxrand;
  for (int i = 1; i <= maxrand - 1; i += 2)
    fpo_gasdev1(S, &idummy, &S->rannumb[i], &S->rannumb[i + 1]);
  fpo_gasdev1(S, &idummy, &S->rannumb[maxrand], &S->rannumb[maxrand - 1]);
}

void fpo_set_rannumb(fpo_state *S, const float *tab, int n) {
  free(S->rannumb);
  S->rannumb = (float *)calloc((size_t)n + 2, sizeof(float));
  S->maxrand = n;
  memcpy(S->rannumb + 1, tab, (size_t)n * sizeof(float));
}

const float *fpo_rannumb(fpo_state *S) { return S->rannumb + 1; }
