/*
 * fpo_math.h -- transcendental layer of the CPU oracle (test infrastructure).
 *
 * The reference uses compiler intrinsics only (exp, log, sqrt, sin, cos,
 * atan2, **, erf); gfortran maps them to libm's float routines, whose last
 * bit depends on the libm version.  The oracle therefore defines them as the
 * correctly rounded float result (evaluate in double, round once), which the
 * device's strict mode reproduces bit for bit.  -DFPO_LIBM_FLOAT switches to
 * glibc's float routines (gfortran's own linkage) to measure the difference.
 */
#ifndef FPO_MATH_H
#define FPO_MATH_H
#include <math.h>

#ifdef FPO_LIBM_FLOAT
static inline float fpo_expf(float x) { return expf(x); }
static inline float fpo_logf(float x) { return logf(x); }
static inline float fpo_powf(float a, float b) { return powf(a, b); }
static inline float fpo_sinf(float x) { return sinf(x); }
static inline float fpo_cosf(float x) { return cosf(x); }
static inline float fpo_erff(float x) { return erff(x); }
#else
static inline float fpo_expf(float x) { return (float)exp((double)x); }
static inline float fpo_logf(float x) { return (float)log((double)x); }
static inline float fpo_powf(float a, float b) {
  return (float)pow((double)a, (double)b);
}
static inline float fpo_sinf(float x) { return (float)sin((double)x); }
static inline float fpo_cosf(float x) { return (float)cos((double)x); }
static inline float fpo_erff(float x) { return (float)erf((double)x); }
#endif
/* sqrt is correctly rounded in IEEE arithmetic on both sides */
static inline float fpo_sqrtf(float x) { return sqrtf(x); }

/* Fortran int(): truncate toward zero; nint(): round half away from zero */
static inline int fpo_int_f(float x) { return (int)x; }
static inline int fpo_int_d(double x) { return (int)x; }
static inline int fpo_nint_d(double x) { return (int)lround(x); }
static inline int fpo_nint_f(float x) { return (int)lroundf(x); }
static inline float fpo_maxf(float a, float b) { return a > b ? a : b; }
static inline float fpo_minf(float a, float b) { return a < b ? a : b; }
/* Fortran modulo(a,p) for reals: result has the sign of p */
static inline double fpo_modulo_d(double a, double p) {
  double r = fmod(a, p);
  if (r != 0.0 && ((r < 0.0) != (p < 0.0))) r += p;
  return r;
}
#endif
