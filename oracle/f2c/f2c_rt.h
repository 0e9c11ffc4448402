/* f2c_rt.h -- run-time support for the C that oracle/f2c/f90toc.py generates
 * from the reference's Fortran sources (TEST INFRASTRUCTURE).
 *
 * Generic intrinsics and the ** operator dispatch on the operand types with
 * C11 _Generic, so the C compiler does the typing Fortran's rules prescribe.
 * Transcendentals go through oracle/fpo_math.h: correctly rounded float
 * results, the same definition the hand-written oracle and the device's strict
 * mode use (gfortran would link libm's float routines; see that header).
 * real**integer follows libgcc's __powisf2/__powidf2 (what gfortran emits).
 */
#ifndef F2C_RT_H
#define F2C_RT_H
#include <float.h>
#include <math.h>
#include <stdlib.h>

#include "../fpo_math.h"

static inline void f2c_stop(void) { abort(); }

static inline float f2c_powi_f(float x, int m) {
  unsigned n = (unsigned)(m < 0 ? -m : m);
  float y = (n % 2) ? x : 1.f;
  while (n >>= 1) {
    x = x * x;
    if (n % 2) y = y * x;
  }
  return m < 0 ? 1.f / y : y;
}
static inline double f2c_powi_d(double x, int m) {
  unsigned n = (unsigned)(m < 0 ? -m : m);
  double y = (n % 2) ? x : 1.;
  while (n >>= 1) {
    x = x * x;
    if (n % 2) y = y * x;
  }
  return m < 0 ? 1. / y : y;
}
static inline int f2c_powi_i(int x, int m) {
  if (m < 0) return (x == 1) ? 1 : ((x == -1) ? ((m % 2) ? -1 : 1) : 0);
  int y = 1;
  while (m-- > 0) y *= x;
  return y;
}
static inline float f2c_pow_ff(float a, float b) { return fpo_powf(a, b); }
static inline double f2c_pow_dd(double a, double b) { return pow(a, b); }
static inline double f2c_pow_fd(float a, double b) { return pow((double)a, b); }
static inline double f2c_pow_df(double a, float b) { return pow(a, (double)b); }
static inline float f2c_pow_if(int a, float b) { return fpo_powf((float)a, b); }
static inline double f2c_pow_id(int a, double b) { return pow((double)a, b); }

#define F_POW(a, b)                                                                                 \
  _Generic((b),                                                                                     \
      int: _Generic((a), int: f2c_powi_i, float: f2c_powi_f, double: f2c_powi_d),                   \
      float: _Generic((a), int: f2c_pow_if, float: f2c_pow_ff, double: f2c_pow_df),                 \
      double: _Generic((a), int: f2c_pow_id, float: f2c_pow_fd, double: f2c_pow_dd))((a), (b))

static inline int f2c_abs_i(int x) { return x < 0 ? -x : x; }
static inline long long f2c_abs_l(long long x) { return x < 0 ? -x : x; }
static inline double f2c_log10_d(double x) { return log10(x); }
static inline float f2c_log10_f(float x) { return (float)log10((double)x); }
static inline float f2c_tan_f(float x) { return (float)tan((double)x); }
static inline float f2c_atan_f(float x) { return (float)atan((double)x); }
static inline float f2c_asin_f(float x) { return (float)asin((double)x); }
static inline float f2c_acos_f(float x) { return (float)acos((double)x); }
static inline float f2c_atan2_f(float y, float x) { return (float)atan2((double)y, (double)x); }
static inline double f2c_atan2_d(double y, double x) { return atan2(y, x); }

#define F_ABS(x) _Generic((x), int: f2c_abs_i, long long: f2c_abs_l, float: fabsf, double: fabs)(x)
#define F_SQRT(x) _Generic((x), float: fpo_sqrtf, double: sqrt)(x)
#define F_EXP(x) _Generic((x), float: fpo_expf, double: exp)(x)
#define F_LOG(x) _Generic((x), float: fpo_logf, double: log)(x)
#define F_LOG10(x) _Generic((x), float: f2c_log10_f, double: f2c_log10_d)(x)
#define F_SIN(x) _Generic((x), float: fpo_sinf, double: sin)(x)
#define F_COS(x) _Generic((x), float: fpo_cosf, double: cos)(x)
#define F_TAN(x) _Generic((x), float: f2c_tan_f, double: tan)(x)
#define F_ATAN(x) _Generic((x), float: f2c_atan_f, double: atan)(x)
#define F_ASIN(x) _Generic((x), float: f2c_asin_f, double: asin)(x)
#define F_ACOS(x) _Generic((x), float: f2c_acos_f, double: acos)(x)
#define F_ERF(x) _Generic((x), float: fpo_erff, double: erf)(x)
#define F_ATAN2(y, x) _Generic((y) + (x), float: f2c_atan2_f, double: f2c_atan2_d)((y), (x))

static inline int f2c_max_i(int a, int b) { return a > b ? a : b; }
static inline float f2c_max_f(float a, float b) { return a > b ? a : b; }
static inline double f2c_max_d(double a, double b) { return a > b ? a : b; }
static inline int f2c_min_i(int a, int b) { return a < b ? a : b; }
static inline float f2c_min_f(float a, float b) { return a < b ? a : b; }
static inline double f2c_min_d(double a, double b) { return a < b ? a : b; }
#define F_MAX(a, b) _Generic((a) + (b), int: f2c_max_i, float: f2c_max_f, double: f2c_max_d)((a), (b))
#define F_MIN(a, b) _Generic((a) + (b), int: f2c_min_i, float: f2c_min_f, double: f2c_min_d)((a), (b))

static inline int f2c_mod_i(int a, int p) { return a % p; }
static inline float f2c_mod_f(float a, float p) { return fmodf(a, p); }
static inline double f2c_mod_d(double a, double p) { return fmod(a, p); }
#define F_MOD(a, p) _Generic((a) + (p), int: f2c_mod_i, float: f2c_mod_f, double: f2c_mod_d)((a), (p))
static inline int f2c_modulo_i(int a, int p) {
  int r = a % p;
  return (r != 0 && ((r < 0) != (p < 0))) ? r + p : r;
}
static inline float f2c_modulo_f(float a, float p) {
  float r = fmodf(a, p);
  return (r != 0.f && ((r < 0.f) != (p < 0.f))) ? r + p : r;
}
static inline double f2c_modulo_d(double a, double p) {
  double r = fmod(a, p);
  return (r != 0. && ((r < 0.) != (p < 0.))) ? r + p : r;
}
#define F_MODULO(a, p) _Generic((a) + (p), int: f2c_modulo_i, float: f2c_modulo_f, double: f2c_modulo_d)((a), (p))

static inline int f2c_sign_i(int a, int b) { return b >= 0 ? f2c_abs_i(a) : -f2c_abs_i(a); }
static inline float f2c_sign_f(float a, float b) { return copysignf(fabsf(a), b); }
static inline double f2c_sign_d(double a, double b) { return copysign(fabs(a), b); }
#define F_SIGN(a, b) _Generic((a) + (b), int: f2c_sign_i, float: f2c_sign_f, double: f2c_sign_d)((a), (b))

static inline int f2c_nint_f(float x) { return (int)lroundf(x); }
static inline int f2c_nint_d(double x) { return (int)lround(x); }
#define F_NINT(x) _Generic((x), float: f2c_nint_f, double: f2c_nint_d)(x)
static inline int f2c_floor_f(float x) { return (int)floorf(x); }
static inline int f2c_floor_d(double x) { return (int)floor(x); }
#define F_FLOOR(x) _Generic((x), float: f2c_floor_f, double: f2c_floor_d)(x)
static inline int f2c_ceil_f(float x) { return (int)ceilf(x); }
static inline int f2c_ceil_d(double x) { return (int)ceil(x); }
#define F_CEILING(x) _Generic((x), float: f2c_ceil_f, double: f2c_ceil_d)(x)
#define F_AINT(x) _Generic((x), float: truncf, double: trunc)(x)
#define F_ANINT(x) _Generic((x), float: roundf, double: round)(x)
#define F_TINY(x) _Generic((x), float: FLT_MIN, double: DBL_MIN)
#define F_HUGE(x) _Generic((x), float: FLT_MAX, double: DBL_MAX, int: 2147483647)
#endif
