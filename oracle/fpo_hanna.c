/*
 * fpo_hanna.c -- oracle restatement of hanna / hanna1 / hanna_short /
 * windalign / get_settling (test infrastructure).
 */
#include "fpo.h"
#include "fpo_math.h"

/* src/hanna.f90:42-106 */
void fpo_hanna(fpo_state *S, float z) {
  float corr;
  if (S->h / fabsf(S->ol) < 1.f) { /* neutral */
    S->ust = fpo_maxf(1.e-4f, S->ust);
    corr = z / S->ust;
    S->sigu = 1.e-2f + 2.0f * S->ust * fpo_expf(-3.e-4f * corr);
    S->sigw = 1.3f * S->ust * fpo_expf(-2.e-4f * corr);
    S->dsigwdz = -2.e-4f * S->sigw;
    S->sigw = S->sigw + 1.e-2f;
    S->sigv = S->sigw;
    S->tlu = 0.5f * z / S->sigw / (1.f + 1.5e-3f * corr);
    S->tlv = S->tlu;
    S->tlw = S->tlu;
  } else if (S->ol < 0.f) { /* unstable */
    S->sigu = 1.e-2f + S->ust * fpo_powf(12.f - 0.5f * S->h / S->ol, 0.33333f);
    S->sigv = S->sigu;
    S->sigw = fpo_sqrtf(1.2f * (S->wst * S->wst) * (1.f - .9f * S->zeta) *
                            fpo_powf(S->zeta, 0.66666f) +
                        (1.8f - 1.4f * S->zeta) * (S->ust * S->ust)) +
              1.e-2f;
    S->dsigwdz =
        0.5f / S->sigw / S->h *
        (-1.4f * (S->ust * S->ust) +
         (S->wst * S->wst) *
             (0.8f * fpo_powf(fpo_maxf(S->zeta, 1.e-3f), -.33333f) -
              1.8f * fpo_powf(S->zeta, 0.66666f)));
    S->tlu = 0.15f * S->h / S->sigu;
    S->tlv = S->tlu;
    if (z < fabsf(S->ol))
      S->tlw = 0.1f * z / (S->sigw * (0.55f - 0.38f * fabsf(z / S->ol)));
    else if (S->zeta < 0.1f)
      S->tlw = 0.59f * z / S->sigw;
    else
      S->tlw = 0.15f * S->h / S->sigw * (1.f - fpo_expf(-5.f * S->zeta));
  } else { /* stable */
    S->sigu = 1.e-2f + 2.f * S->ust * (1.f - S->zeta);
    S->sigv = 1.e-2f + 1.3f * S->ust * (1.f - S->zeta);
    S->sigw = S->sigv;
    S->dsigwdz = -1.3f * S->ust / S->h;
    S->tlu = 0.15f * S->h / S->sigu * (fpo_sqrtf(S->zeta));
    S->tlv = 0.467f * S->tlu;
    S->tlw = 0.1f * S->h / S->sigw * fpo_powf(S->zeta, 0.8f);
  }
  S->tlu = fpo_maxf(10.f, S->tlu);
  S->tlv = fpo_maxf(10.f, S->tlv);
  S->tlw = fpo_maxf(30.f, S->tlw);
  if (S->dsigwdz == 0.f) S->dsigwdz = 1.e-10f;
}

/* src/hanna1.f90:42-128 */
void fpo_hanna1(fpo_state *S, float z) {
  float s1, s2;
  if (S->h / fabsf(S->ol) < 1.f) {
    S->ust = fpo_maxf(1.e-4f, S->ust);
    S->sigu = 2.0f * S->ust * fpo_expf(-3.e-4f * z / S->ust);
    S->sigu = fpo_maxf(S->sigu, 1.e-5f);
    S->sigv = 1.3f * S->ust * fpo_expf(-2.e-4f * z / S->ust);
    S->sigv = fpo_maxf(S->sigv, 1.e-5f);
    S->sigw = S->sigv;
    S->dsigw2dz = -6.76e-4f * S->ust * fpo_expf(-4.e-4f * z / S->ust);
    S->tlu = 0.5f * z / S->sigw / (1.f + 1.5e-3f * z / S->ust);
    S->tlv = S->tlu;
    S->tlw = S->tlu;
  } else if (S->ol < 0.f) {
    S->sigu = S->ust * fpo_powf(12.f - 0.5f * S->h / S->ol, 0.33333f);
    S->sigu = fpo_maxf(S->sigu, 1.e-6f);
    S->sigv = S->sigu;
    if (S->zeta < 0.03f) {
      S->sigw = 0.96f * S->wst * fpo_powf(3.f * S->zeta - S->ol / S->h, 0.33333f);
      S->dsigw2dz = 1.8432f * S->wst * S->wst / S->h *
                    fpo_powf(3.f * S->zeta - S->ol / S->h, -0.33333f);
    } else if (S->zeta < 0.4f) {
      s1 = 0.96f * fpo_powf(3.f * S->zeta - S->ol / S->h, 0.33333f);
      s2 = 0.763f * fpo_powf(S->zeta, 0.175f);
      if (s1 < s2) {
        S->sigw = S->wst * s1;
        S->dsigw2dz = 1.8432f * S->wst * S->wst / S->h *
                      fpo_powf(3.f * S->zeta - S->ol / S->h, -0.33333f);
      } else {
        S->sigw = S->wst * s2;
        S->dsigw2dz = 0.203759f * S->wst * S->wst / S->h * fpo_powf(S->zeta, -0.65f);
      }
    } else if (S->zeta < 0.96f) {
      S->sigw = 0.722f * S->wst * fpo_powf(1.f - S->zeta, 0.207f);
      S->dsigw2dz = -.215812f * S->wst * S->wst / S->h * fpo_powf(1.f - S->zeta, -0.586f);
    } else if (S->zeta < 1.00f || !S->strict_reference) {
      /* the reference leaves sigw/dsigw2dz untouched for zeta == 1 exactly
       * (stale module state); "defined" behaviour extends the last branch */
      S->sigw = 0.37f * S->wst;
      S->dsigw2dz = 0.f;
    }
    S->sigw = fpo_maxf(S->sigw, 1.e-6f);
    S->tlu = 0.15f * S->h / S->sigu;
    S->tlv = S->tlu;
    if (z < fabsf(S->ol))
      S->tlw = 0.1f * z / (S->sigw * (0.55f - 0.38f * fabsf(z / S->ol)));
    else if (S->zeta < 0.1f)
      S->tlw = 0.59f * z / S->sigw;
    else
      S->tlw = 0.15f * S->h / S->sigw * (1.f - fpo_expf(-5.f * S->zeta));
  } else {
    S->sigu = 2.f * S->ust * (1.f - S->zeta);
    S->sigv = 1.3f * S->ust * (1.f - S->zeta);
    S->sigu = fpo_maxf(S->sigu, 1.e-6f);
    S->sigv = fpo_maxf(S->sigv, 1.e-6f);
    S->sigw = S->sigv;
    S->dsigw2dz = 3.38f * S->ust * S->ust * (S->zeta - 1.f) / S->h;
    S->tlu = 0.15f * S->h / S->sigu * (fpo_sqrtf(S->zeta));
    S->tlv = 0.467f * S->tlu;
    S->tlw = 0.1f * S->h / S->sigw * fpo_powf(S->zeta, 0.8f);
  }
  S->tlu = fpo_maxf(10.f, S->tlu);
  S->tlv = fpo_maxf(10.f, S->tlv);
  S->tlw = fpo_maxf(30.f, S->tlw);
}

/* src/hanna_short.f90:42-92 */
void fpo_hanna_short(fpo_state *S, float z) {
  if (S->h / fabsf(S->ol) < 1.f) {
    S->ust = fpo_maxf(1.e-4f, S->ust);
    S->sigw = 1.3f * fpo_expf(-2.e-4f * z / S->ust);
    S->dsigwdz = -2.e-4f * S->sigw;
    S->sigw = S->sigw * S->ust + 1.e-2f;
    S->tlw = 0.5f * z / S->sigw / (1.f + 1.5e-3f * z / S->ust);
  } else if (S->ol < 0.f) {
    S->sigw = fpo_sqrtf(1.2f * (S->wst * S->wst) * (1.f - .9f * S->zeta) *
                            fpo_powf(S->zeta, 0.66666f) +
                        (1.8f - 1.4f * S->zeta) * (S->ust * S->ust)) +
              1.e-2f;
    S->dsigwdz =
        0.5f / S->sigw / S->h *
        (-1.4f * (S->ust * S->ust) +
         (S->wst * S->wst) *
             (0.8f * fpo_powf(fpo_maxf(S->zeta, 1.e-3f), -.33333f) -
              1.8f * fpo_powf(S->zeta, 0.66666f)));
    if (z < fabsf(S->ol))
      S->tlw = 0.1f * z / (S->sigw * (0.55f - 0.38f * fabsf(z / S->ol)));
    else if (S->zeta < 0.1f)
      S->tlw = 0.59f * z / S->sigw;
    else
      S->tlw = 0.15f * S->h / S->sigw * (1.f - fpo_expf(-5.f * S->zeta));
  } else {
    S->sigw = 1.e-2f + 1.3f * S->ust * (1.f - S->zeta);
    S->dsigwdz = -1.3f * S->ust / S->h;
    S->tlw = 0.1f * S->h / S->sigw * fpo_powf(S->zeta, 0.8f);
  }
  S->tlu = fpo_maxf(10.f, S->tlu);
  S->tlv = fpo_maxf(10.f, S->tlv);
  S->tlw = fpo_maxf(30.f, S->tlw);
  if (S->dsigwdz == 0.f) S->dsigwdz = 1.e-10f;
}

/* src/windalign.f90:36-54 */
void fpo_windalign(float u, float v, float ffap, float ffcp, float *ux,
                   float *vy) {
  const float eps = 1.e-30f;
  float ffinv = 1.f / fpo_maxf(fpo_sqrtf(u * u + v * v), eps);
  float sinphi = v * ffinv;
  float vy1 = sinphi * ffap;
  float cosphi = u * ffinv;
  float ux1 = cosphi * ffap;
  float ux2 = -sinphi * ffcp;
  float vy2 = cosphi * ffcp;
  *ux = ux1 + ux2;
  *vy = vy1 + vy2;
}

/* src/dynamic_viscosity.f90 */
static float viscosity(float t) {
  const float c = 120.f, t_0 = 291.15f, eta_0 = 1.827e-5f;
  return eta_0 * (t_0 + c) / (t + c) * fpo_powf(t / t_0, 1.5f);
}

/* src/get_settling.f90:52-125: nearest-neighbour rho, tt from SLOT 1
 * (literal), Reynolds iteration */
void fpo_get_settling(fpo_state *S, int itime, float xt, float yt, float zt,
                      int nsp, float *settling) {
  (void)itime;
  const float ga = 9.81f;
  int indz = 1, nix = fpo_int_f(xt), njy = fpo_int_f(yt);
  float rho1[3], tt1[3];
  for (int i = 2; i <= S->c.nz; i++)
    if (S->height[i] > zt) {
      indz = i - 1;
      break;
    }
  float dz = 1.f / (S->height[indz + 1] - S->height[indz]);
  float dz1 = (zt - S->height[indz]) * dz;
  float dz2 = (S->height[indz + 1] - zt) * dz;
  for (int n = 1; n <= 2; n++) {
    int indzh = indz + n - 1;
    size_t a = (size_t)nix + (size_t)S->c.nxmax * ((size_t)njy + (size_t)S->c.nymax * (size_t)(indzh - 1));
    rho1[n] = S->met[1].rho[a];
    tt1[n] = S->met[1].tt[a];
  }
  float temperature = dz2 * tt1[1] + dz1 * tt1[2];
  float airdens = dz2 * rho1[1] + dz1 * rho1[2];
  float vis_dyn = viscosity(temperature);
  float vis_kin = vis_dyn / airdens;
  float dq = S->c.dquer[nsp - 1], vsa = S->c.vsetaver[nsp - 1];
  float reynolds = dq / 1.e6f * fabsf(vsa) / vis_kin;
  float settling_old = vsa, c_d;
  for (int i = 1; i <= 20; i++) {
    if (reynolds < 1.917f)
      c_d = 24.f / reynolds;
    else if (reynolds < 500.f)
      c_d = 18.5f / fpo_powf(reynolds, 0.6f);
    else
      c_d = 0.44f;
    *settling = -1.f * fpo_sqrtf(4.f * ga * dq / 1.e6f * S->c.density[nsp - 1] *
                                 S->c.cunningham[nsp - 1] /
                                 (3.f * c_d * airdens));
    if (fabsf((*settling - settling_old) / *settling) < 0.01f) break;
    reynolds = dq / 1.e6f * fabsf(*settling) / vis_kin;
    settling_old = *settling;
  }
}
