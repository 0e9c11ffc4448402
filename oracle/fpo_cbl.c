/*
 * fpo_cbl.c -- oracle restatement of the skewed convective-BL scheme
 * (test infrastructure): cbl (src/cbl.f90:4-214), cuberoot (:220-234),
 * re_initialize_particle (src/re_initialize_particle.f90:4-93),
 * initialize_cbl_vel (src/initialize_cbl_vel.f90:4-87).
 *
 * `erf` in cbl.f90 is declared `real :: erf` without EXTERNAL, so gfortran
 * binds the F2008 intrinsic; fpo_erff stands for it.
 * Integer powers (**2, **3) are repeated multiplication; real-valued
 * exponents (**2., **3., **1.5, **0.5 ...) go through fpo_powf, as gfortran
 * emits powf for them at -O2 without -ffast-math.
 */
#include "fpo.h"
#include "fpo_math.h"

static const float PI_F = 3.14159265f; /* par_mod pi */

static float cuberoot(float x) {
  const float third = 0.333333333f;
  return copysignf(fpo_powf(fabsf(x), third), x);
}

void fpo_cbl(fpo_state *S, float wp, float zp, float ust, float wst, float h,
             float rhoa, float rhograd, float sigmaw, float dsigmawdz,
             float tlw, float *ptot_o, float *Q_o, float *phi_o, float *ath_o,
             float *bth_o, float ol, int *flagrein) {
  (void)ust;
  const float usurad2 = 0.7071067812f, usurad2p = 0.3989422804f, C0 = 3.f,
              costluar4 = 0.66667f, eps = 0.000001f;
  float dens = rhoa, ddens = rhograd;
  float timedir = (float)S->c.ldirect;
  float z = (zp / h);
  float transition = 1.f;
  if (-h / ol < 15.f)
    transition = (fpo_sinf((((-h / ol) + 10.f) / 10.f) * PI_F)) / 2.f + 0.5f;
  float w2 = (sigmaw * sigmaw);
  float dw2 = (2.f * sigmaw * dsigmawdz);
  float alfa = 2.f * w2 / (C0 * tlw);
  float wold = timedir * wp;
  float w3 = ((1.2f * z * (fpo_powf(1.f - z, 3.f / 2.f))) + eps) * (wst * wst * wst) * transition;
  float dw3 = (1.2f * ((fpo_powf(1.f - z, 3.f / 2.f)) +
                       z * 1.5f * (fpo_powf(1.f - z, 1.f / 2.f)) * (-1.f))) *
              (wst * wst * wst) * (1.f / h) * transition;
  float skew = w3 / (fpo_powf(w2, 1.5f));
  float skew2 = skew * skew;
  float dskew = (dw3 * fpo_powf(w2, 1.5f) - w3 * 1.5f * fpo_powf(w2, 0.5f) * dw2) / (w2 * w2 * w2);
  float radw2 = fpo_powf(w2, 0.5f);
  float dradw2 = 0.5f * fpo_powf(w2, -0.5f) * dw2;
  float fluarw = costluar4 * (cuberoot(skew));
  float fluarw2 = fluarw * fluarw;
  float dfluarw, rluarw, xluarw, drluarw, dxluarw;
  if (skew != 0.f) {
    float a1 = 1.f + fluarw2, a3 = 3.f + fluarw2;
    dfluarw = costluar4 * (1.f / 3.f) * cuberoot(fpo_powf(skew, -2.f)) * dskew;
    rluarw = fpo_powf(a1, 3.f) * skew2 / (fpo_powf(a3, 2.f) * fluarw2);
    xluarw = fpo_powf(a1, 1.5f) * skew / (a3 * fluarw);
    drluarw = (((3.f * (a1 * a1) * (2.f * fluarw * dfluarw) * skew2) +
                (a1 * a1 * a1) * 2.f * skew * dskew) *
                   fpo_powf(a3, 2.f) * fluarw2 -
               (a1 * a1 * a1) * skew2 *
                   ((2.f * a3 * (2.f * fluarw * dfluarw) * fluarw2) +
                    (a3 * a3) * 2.f * fluarw * dfluarw)) /
              ((fpo_powf(a3, 2.f) * fluarw2) * (fpo_powf(a3, 2.f) * fluarw2));
    dxluarw = (((1.5f * fpo_powf(a1, 0.5f) * (2.f * fluarw * dfluarw) * skew) +
                fpo_powf(a1, 1.5f) * dskew) *
                   a3 * fluarw -
               fpo_powf(a1, 1.5f) * skew * (3.f * dfluarw + 3.f * fluarw2 * dfluarw)) /
              ((a3 * fluarw) * (a3 * fluarw));
  } else {
    dfluarw = 0.f;
    rluarw = 0.f;
    drluarw = 0.f;
    xluarw = 0.f;
    dxluarw = 0.f;
  }
  float aluarw = 0.5f * (1.f - xluarw / fpo_powf(4.f + rluarw, 0.5f));
  float bluarw = 1.f - aluarw;
  float daluarw = -0.5f *
                  ((dxluarw * fpo_powf(4.f + rluarw, 0.5f)) -
                   (0.5f * xluarw * fpo_powf(4.f + rluarw, -0.5f) * drluarw)) /
                  (4.f + rluarw);
  float dbluarw = -daluarw;
  float ra = bluarw / (aluarw * (1.f + fluarw2));
  float rb = aluarw / (bluarw * (1.f + fluarw2));
  float sigmawa = radw2 * fpo_powf(ra, 0.5f);
  float sigmawb = radw2 * fpo_powf(rb, 0.5f);
  float dsigmawa =
      dradw2 * fpo_powf(ra, 0.5f) +
      radw2 * ((0.5f * fpo_powf(ra, -0.5f)) *
               ((dbluarw * (aluarw * (1.f + fluarw2)) -
                 bluarw * (daluarw * (1.f + fluarw2) + aluarw * 2.f * fluarw * dfluarw)) /
                ((aluarw * (1.f + fluarw2)) * (aluarw * (1.f + fluarw2)))));
  float dsigmawb =
      dradw2 * fpo_powf(rb, 0.5f) +
      radw2 * ((0.5f * fpo_powf(rb, -0.5f)) *
               ((daluarw * (bluarw * (1.f + fluarw2)) -
                 aluarw * (dbluarw * (1.f + fluarw2) + bluarw * 2.f * fluarw * dfluarw)) /
                ((bluarw * (1.f + fluarw2)) * (bluarw * (1.f + fluarw2)))));
  float wa = (fluarw * sigmawa);
  float wb = (fluarw * sigmawb);
  float dwa = dfluarw * sigmawa + fluarw * dsigmawa;
  float dwb = dfluarw * sigmawb + fluarw * dsigmawb;
  float deltawa = wold - wa;
  float deltawb = wold + wb;
  float wold2 = wold * wold;
  float sigmawa2 = sigmawa * sigmawa;
  float sigmawb2 = sigmawb * sigmawb;
  if (fabsf(deltawa) > 6.f * sigmawa && fabsf(deltawb) > 6.f * sigmawb) *flagrein = 1;
  float pa = (usurad2p * (1.f / sigmawa)) *
             (fpo_expf(-(0.5f * (fpo_powf(deltawa / sigmawa, 2.f)))));
  float pb = (usurad2p * (1.f / sigmawb)) *
             (fpo_expf(-(0.5f * (fpo_powf(deltawb / sigmawb, 2.f)))));
  float ptot = dens * aluarw * pa + dens * bluarw * pb;
  float aperfa = deltawa * usurad2 / sigmawa;
  float aperfb = deltawb * usurad2 / sigmawb;
  float Phi =
      -0.5f * (aluarw * dens * dwa + dens * wa * daluarw + aluarw * wa * ddens) * fpo_erff(aperfa) +
      sigmawa *
          (aluarw * dens * dsigmawa * (wold2 / sigmawa2 + 1.f) +
           sigmawa * dens * daluarw + sigmawa * ddens * aluarw +
           aluarw * wold * dens / sigmawa2 * (sigmawa * dwa - wa * dsigmawa)) *
          pa +
      0.5f * (bluarw * dens * dwb + wb * dens * dbluarw + wb * bluarw * ddens) * fpo_erff(aperfb) +
      sigmawb *
          (bluarw * dens * dsigmawb * (wold2 / sigmawb2 + 1.f) +
           sigmawb * dens * dbluarw + sigmawb * ddens * bluarw +
           bluarw * wold * dens / sigmawb2 * (-sigmawb * dwb + wb * dsigmawb)) *
          pb;
  float Q = timedir * ((aluarw * dens * deltawa / sigmawa2) * pa +
                       (bluarw * dens * deltawb / sigmawb2) * pb);
  *ath_o = (1.f / ptot) * (-(C0 / 2.f) * alfa * Q + Phi);
  *bth_o = fpo_sqrtf(C0 * alfa);
  *ptot_o = ptot;
  *Q_o = Q;
  *phi_o = Phi;
}

/* shared moment closure of re_initialize_particle / initialize_cbl_vel */
static void lhh_split(float zp, float wst, float h, float sigmaw, float ol,
                      float *aluarw, float *sigmawa, float *sigmawb, float *wa,
                      float *wb) {
  const float costluar4 = 0.66667f, eps = 0.000001f;
  float z = zp / h;
  float transition = 1.f;
  if (-h / ol < 15.f)
    transition = (fpo_sinf((((-h / ol) + 10.f) / 10.f) * PI_F)) / 2.f + 0.5f;
  float w2 = sigmaw * sigmaw;
  float w3 = (((1.2f * z * (fpo_powf(1.f - z, 3.f / 2.f))) + eps) * (wst * wst * wst)) * transition;
  float skew = w3 / (fpo_powf(w2, 1.5f));
  float skew2 = skew * skew;
  float radw2 = fpo_sqrtf(w2);
  float fluarw = costluar4 * fpo_powf(skew, 0.333333333333333f);
  float fluarw2 = fluarw * fluarw;
  float rluarw = fpo_powf(1.f + fluarw2, 3.f) * skew2 / (fpo_powf(3.f + fluarw2, 2.f) * fluarw2);
  float xluarw = fpo_powf(rluarw, 0.5f);
  *aluarw = 0.5f * (1.f - xluarw / fpo_powf(4.f + rluarw, 0.5f));
  float bluarw = 1.f - *aluarw;
  *sigmawa = radw2 * fpo_powf(bluarw / (*aluarw * (1.f + fluarw2)), 0.5f);
  *sigmawb = radw2 * fpo_powf(*aluarw / (bluarw * (1.f + fluarw2)), 0.5f);
  *wa = (fluarw * *sigmawa);
  *wb = (fluarw * *sigmawb);
}

/* src/re_initialize_particle.f90:44-91 */
void fpo_re_initialize_particle(fpo_state *S, float zp, float ust, float wst,
                                float h, float sigmaw, float *wp, int *nrand,
                                float ol) {
  (void)ust;
  float aluarw, sigmawa, sigmawb, wa, wb;
  *nrand = *nrand + 1;
  float dcas1 = S->rannumb[*nrand];
  float timedir = (float)S->c.ldirect;
  lhh_split(zp, wst, h, sigmaw, ol, &aluarw, &sigmawa, &sigmawb, &wa, &wb);
  if ((copysignf(1.f, *wp) * timedir) > 0.f) { /* updraft */
    for (;;) {
      *wp = (dcas1 * sigmawa + wa);
      if (*wp < 0.f) {
        *nrand = *nrand + 1;
        dcas1 = S->rannumb[*nrand];
        continue;
      }
      break;
    }
    *wp = *wp * timedir;
  } else if ((copysignf(1.f, *wp) * timedir) < 0.f) { /* downdraft */
    for (;;) {
      *wp = (dcas1 * sigmawb - wb);
      if (*wp > 0.f) {
        *nrand = *nrand + 1;
        dcas1 = S->rannumb[*nrand];
        continue;
      }
      break;
    }
    *wp = *wp * timedir;
  }
}

/* src/initialize_cbl_vel.f90:46-84 */
void fpo_initialize_cbl_vel(fpo_state *S, int *idum, float zp, float ust,
                            float wst, float h, float sigmaw, float *wp,
                            float ol) {
  (void)ust;
  float aluarw, sigmawa, sigmawb, wa, wb;
  float timedir = (float)S->c.ldirect;
  lhh_split(zp, wst, h, sigmaw, ol, &aluarw, &sigmawa, &sigmawb, &wa, &wb);
  float dcas = fpo_ran3(S, idum);
  if (dcas <= aluarw) {
    float dcas1 = fpo_gasdev(S, idum);
    *wp = timedir * (dcas1 * sigmawa + wa);
  } else {
    float dcas1 = fpo_gasdev(S, idum);
    *wp = timedir * (dcas1 * sigmawb - wb);
  }
}

/* "defined" variant of initialize_cbl_vel (strict_reference == 0): same
 * closure, draws supplied by the caller */
void fpo_initialize_cbl_vel_defined(fpo_state *S, float dcas, float dcas1,
                                    float zp, float wst, float h, float sigmaw,
                                    float *wp, float ol) {
  float aluarw, sigmawa, sigmawb, wa, wb;
  float timedir = (float)S->c.ldirect;
  lhh_split(zp, wst, h, sigmaw, ol, &aluarw, &sigmawa, &sigmawb, &wa, &wb);
  if (dcas <= aluarw)
    *wp = timedir * (dcas1 * sigmawa + wa);
  else
    *wp = timedir * (dcas1 * sigmawb - wb);
}
