// fpb_step.cuh -- advance() + the timemanager loop body as two kernels.
// Included inside the anonymous namespace of fpb_kernels.cu.
//
// The reference processes one particle start to finish (src/advance.f90:133-985
// called from src/timemanager.f90:531-712).  The number of Langevin sub-steps
// per call varies by an order of magnitude between particles (ldt follows the
// local Lagrangian time scale) while everything after the sub-step loop runs
// exactly once per particle, so the call is cut at label 99 / label 700:
//
//   fpb_pbl_kernel     persistent; a lane owns a particle only while it has
//                      sub-steps to do.  REFILL = src/advance.f90:133-276 (load
//                      state, grid choice, weights, mixing height, surface
//                      parameters); SUBSTEP = one pass of the label-100 loop,
//                      src/advance.f90:282-609.  A lane whose particle leaves
//                      the loop writes the values the rest of advance() needs
//                      to a scratch row and becomes idle; idle lanes pull the
//                      next rows from a global counter in batches.
//   fpb_finish_kernel  one thread per row, fully converged: label 700 (step
//                      above the PBL), label 99 (mesoscale term, windalign,
//                      position update, wrap/exit), Petterssen corrector, then
//                      the rest of the timemanager loop body
//                      (src/timemanager.f90:630-707) and the store.
//
// Per-particle arithmetic and its order are unchanged (floats pass through the
// scratch row bit for bit), so the strict build stays bit-identical to the
// oracle.

// SC_TERM_*: why the step terminated the particle (written only when linit_cond asks for initial_cond_calc)
enum : int { SC_PBL = 1, SC_ABOVE = 2, SC_TERM_NSTOP = 4, SC_TERM_AGE = 8 };

__device__ __noinline__ float rare_fmodf(float a, float b) { return fmodf(a, b); }

// ----------------------------------------------------------- PBL kernel ----
// Per-lane shared-memory rows (word w of lane t lives at ls[w * PBL_THREADS],
// ls = smem + t, so every access is conflict-free):
//   LS_P1..LS_P4, LS_O00, LS_O01   horizontal weights / corner offsets of the call
//   LS_TAG + e                     level held by cache entry e (0 = empty)
//   LS_LEV + 5*e + f               field f of that level: u, v, w, rho, rhograd
// The profile cache is direct-mapped by (level mod PBL_CACHE): it plays the
// role of the reference's nzmax-long uprof.. arrays + indzindicator
// (src/advance.f90:310-331) for the levels a particle visits during one call.
#ifndef FPB_PBL_THREADS
#define FPB_PBL_THREADS 128
#endif
constexpr int PBL_THREADS = FPB_PBL_THREADS;
#ifndef FPB_PBL_CACHE
#define FPB_PBL_CACHE 8
#endif
constexpr int PBL_CACHE = FPB_PBL_CACHE; // power of two
constexpr int LEV_WORDS = 5; // u, v, w, rho, rhograd
enum : int { LS_P1 = 0, LS_P2, LS_P3, LS_P4, LS_O00, LS_O01, LS_TAG, LS_LEV = LS_TAG + PBL_CACHE,
             LS_WORDS = LS_LEV + LEV_WORDS * PBL_CACHE };

// SPEC = true: forward or backward Hanna run with turbswitch on, method 1, IFINE = 4, turbulence
// not switched off -- the switches are compile-time constants (the launcher checks them).
template <bool EXTRA, bool CBL, bool SPEC = false>
struct PblTask {
  int j;
  bool running, first;
  int regime; // 0: h/|ol| < 1, 1: ol < 0, 2: stable  (hanna*.f90 branch, fixed for the call)
  int ngrid;
  float zt, up, vp, wp;
  int ldt, icbt;
  int nrand, itimec;
  float ust, wst, ol, h;
  int indz;
  float dxsave, dysave, dawsave, dcwsave;
  int nsub, nan_cbl;
  // EXTRA only
  float xtf, ytf;
  int npoint;
  Rng rng;
  float prob[EXTRA ? FPB_MAXSPEC : 1];
  float vdepo[EXTRA ? FPB_MAXSPEC : 1];
  unsigned depo_todo;

  __device__ __forceinline__ static float &lsf(float *ls, int w) { return ls[w * PBL_THREADS]; }
  __device__ __forceinline__ static int &lsi(float *ls, int w) {
    return reinterpret_cast<int *>(ls)[w * PBL_THREADS];
  }

  // interpol_mod weights of the call, rebuilt from the lane's shared row
  __device__ __forceinline__ void load_weights(const DevCfg &c, float *ls, Hz &z) const {
    z.p1 = lsf(ls, LS_P1); z.p2 = lsf(ls, LS_P2); z.p3 = lsf(ls, LS_P3); z.p4 = lsf(ls, LS_P4);
    z.o00 = lsi(ls, LS_O00); z.o10 = z.o00 + 1;
    z.o01 = lsi(ls, LS_O01); z.o11 = z.o01 + 1;
    z.dt1 = (float)(c.itime - c.memtime[0]);
    z.dt2 = (float)(c.memtime[1] - c.itime);
    z.dtt = 1.f / (z.dt1 + z.dt2);
    z.ngrid = ngrid;
    z.plane = (EXTRA && ngrid > 0) ? c.nxdn[ngrid - 1] * c.nydn[ngrid - 1] : c.nxd * c.nyd;
  }
  // time levels of the grid the call interpolates from (nests: EXTRA variant only)
  __device__ __forceinline__ const DevMetSlot *grid_met(const DevStepArgs &a) const {
    return (EXTRA && ngrid > 0) ? a.metn[ngrid - 1] : a.met;
  }

  __device__ __forceinline__ float normal(const DevStepArgs &a, int i) {
    if (EXTRA) return rng.get(i);
    return __ldg(a.rannumb + (i - 1));
  }

  // src/advance.f90:133-276.  Returns true when the row is active.
  __device__ __forceinline__ bool refill(const DevStepArgs &a, float *ls, int row, bool &pbl) {
    const DevCfg &c = a.cfg;
    pbl = false;
    if (a.p.itra1[row] != c.itime) return false;
    j = row;
    const double xt = a.p.xtra1[row], yt = a.p.ytra1[row];
    zt = a.p.ztra1[row];

    Hz z;
    z.ngrid = EXTRA ? choose_grid(c, xt, yt) : pole_grid(c, yt);
    ngrid = z.ngrid;
    const DevMetSlot *met = a.met;
    if (EXTRA && z.ngrid > 0) { // nested grid coordinates, advance.f90:191-203
      const GridSel g = select_grid(a, z.ngrid, xt, yt);
      int jyp = g.jy + 1;
      if (jyp >= c.nymax) jyp = jyp - 1;
      make_weights(c, z, c.itime, g.xf, g.yf, g.ix, g.jy, g.ix + 1, jyp, g.nxd, g.nyd);
      met = g.met;
    } else {
      const int ix = d_int(xt), jy = d_int(yt);
      int ixp = ix + 1, jyp = jy + 1;
      if (jyp >= c.nymax) jyp = jyp - 1;
      make_weights(c, z, c.itime, (float)xt, (float)yt, ix, jy, ixp, jyp);
    }

    // advance.f90:236-264 (max of hmix over 4 corners x 2 slots); the same
    // words carry ustar, wstar, oli for interpol_all.f90:80-107
    float4 s[2][4];
#pragma unroll
    for (int m = 0; m < 2; m++) {
      s[m][0] = __ldg(met[m].S + z.o00); s[m][1] = __ldg(met[m].S + z.o10);
      s[m][2] = __ldg(met[m].S + z.o01); s[m][3] = __ldg(met[m].S + z.o11);
    }
    float hh = 0.f;
#pragma unroll
    for (int m = 0; m < 2; m++)
#pragma unroll
      for (int q = 0; q < 4; q++)
        if (s[m][q].x > hh) hh = s[m][q].x;
    h = hh;
    const float zeta = zt / h;
    if (!(zeta <= 1.f)) { // above the PBL for the whole step: nothing to do here
      a.sc.flags[row] = SC_ABOVE;
      running = false;
      return true;
    }
    pbl = true;
    lsf(ls, LS_P1) = z.p1; lsf(ls, LS_P2) = z.p2; lsf(ls, LS_P3) = z.p3; lsf(ls, LS_P4) = z.p4;
    lsi(ls, LS_O00) = z.o00; lsi(ls, LS_O01) = z.o01;
#pragma unroll
    for (int e = 0; e < PBL_CACHE; e++) lsi(ls, LS_TAG + e) = 0;
    {
      float us1[2], ws1[2], ol1[2];
#pragma unroll
      for (int m = 0; m < 2; m++) {
        us1[m] = bil(z, s[m][0].y, s[m][1].y, s[m][2].y, s[m][3].y);
        ws1[m] = bil(z, s[m][0].z, s[m][1].z, s[m][2].z, s[m][3].z);
        ol1[m] = bil(z, s[m][0].w, s[m][1].w, s[m][2].w, s[m][3].w);
      }
      ust = (us1[0] * z.dt2 + us1[1] * z.dt1) * z.dtt;
      wst = (ws1[0] * z.dt2 + ws1[1] * z.dt1) * z.dtt;
      const float oliaux = (ol1[0] * z.dt2 + ol1[1] * z.dt1) * z.dtt;
      ol = (oliaux != 0.f) ? 1.f / oliaux : 99999.f;
    }
    regime = (h / fabsf(ol) < 1.f) ? 0 : ((ol < 0.f) ? 1 : 2);

    ldt = a.p.idt[row];
    up = a.p.uap[row]; vp = a.p.ucp[row]; wp = a.p.uzp[row];
    icbt = a.p.cbt[row];
    const int slot = a.p.slot[row];
    if (EXTRA) {
      make_rng(c, a.rannumb, slot, rng);
      xtf = (float)xt; ytf = (float)yt;
      npoint = a.p.npoint[row];
#pragma unroll
      for (int ks = 0; ks < FPB_MAXSPEC; ks++) { prob[ks] = 0.f; vdepo[ks] = 0.f; }
      depo_todo = 0xffu;
    }
    nrand = advance_nrand(a, slot);
    dxsave = 0.f; dysave = 0.f; dawsave = 0.f; dcwsave = 0.f;
    itimec = c.itime;
    nsub = 0; nan_cbl = 0;
    indz = 0;
    first = true;
    running = true;
    return true;
  }

  // leave the sub-step loop: hand the rest of advance() to fpb_finish_kernel
  __device__ __forceinline__ void finish(const DevStepArgs &a, bool above, float u, float v, float w) {
    const DevCfg &c = a.cfg;
    a.p.ztra1[j] = zt;
    a.p.uap[j] = up; a.p.ucp[j] = vp; a.p.uzp[j] = wp;
    a.p.idt[j] = ldt;
    a.p.cbt[j] = (int16_t)icbt;
    a.sc.s0[j] = make_float4(dxsave, dysave, dawsave, dcwsave);
    a.sc.s1[j] = make_float4(u, v, w, __int_as_float(indz));
    a.sc.s2[j] = make_int2(nrand, itimec);
    a.sc.flags[j] = SC_PBL | (above ? SC_ABOVE : 0);
    if (EXTRA && c.drydep) {
#pragma unroll
      for (int ks = 0; ks < FPB_MAXSPEC; ks++)
        if (ks < c.nspec) a.sc.prob[(size_t)ks * a.p.maxpart + j] = prob[ks];
    }
    running = false;
  }

  // one pass of the label-100 loop, src/advance.f90:282-609
  __device__ __forceinline__ void substep(const DevStepArgs &a, const float *sh, float *ls) {
    const DevCfg &c = a.cfg;
    const int itime = c.itime, nz = c.nz, maxrand = c.maxrand;
    const bool turbswitch = SPEC ? true : (c.turbswitch != 0), turboff = SPEC ? false : (c.turboff != 0);
    const bool method1 = SPEC ? true : (c.method == 1);
    const int ifine = SPEC ? 4 : c.ifine; // IFINE = 4: the shipped options/COMMAND value
    nsub++;
    if (method1) {
      ldt = min(ldt, abs(c.lsynctime - itimec + itime));
      itimec = itimec + ldt * c.ldirect;
    } else {
      ldt = abs(c.lsynctime);
      itimec = itime + c.lsynctime;
    }
    const float dt = (float)ldt;
    Turb t;
    t.ust = ust; t.wst = wst; t.ol = ol; t.h = h;
    t.zeta = zt / t.h;

    // level pair under the particle (src/advance.f90:310-331): search, then
    // compute the levels of the pair that this call has not computed yet
    int e_lo, e_hi;
    {
      int ni;
      if (first) {
        ni = find_indz(sh, nz, zt);
        first = false;
      } else { // walk from the previous level (same index as the reference's search from 2)
        // two branch-free moves (a sub-step rarely crosses more than one level), then the
        // rare remainder as loops
        ni = indz;
        const bool up = ni + 1 < nz && sh[ni] <= zt;
        const bool dn = !up && ni > 1 && sh[ni - 1] > zt;
        ni += up ? 1 : (dn ? -1 : 0);
        const bool up2 = up && ni + 1 < nz && sh[ni] <= zt;
        const bool dn2 = dn && ni > 1 && sh[ni - 1] > zt;
        ni += up2 ? 1 : (dn2 ? -1 : 0);
        if (up2) {
          while (ni + 1 < nz && sh[ni] <= zt) ni++;
        } else if (dn2) {
          while (ni > 1 && sh[ni - 1] > zt) ni--;
        }
      }
      indz = ni;
      e_lo = ni & (PBL_CACHE - 1);
      e_hi = (ni + 1) & (PBL_CACHE - 1);
      bool need_lo = lsi(ls, LS_TAG + e_lo) != ni;
      bool need_hi = lsi(ls, LS_TAG + e_hi) != ni + 1;
      // one converged call site; a lane that needs both levels goes round twice
#pragma unroll 1
      while (need_lo || need_hi) {
        const bool do_lo = need_lo;
        const int n = ni + (do_lo ? 0 : 1), e = do_lo ? e_lo : e_hi;
        Hz z;
        load_weights(c, ls, z);
        Lev L;
        profile_level<false>(c, grid_met(a), z, n, L);
        float *q = ls + (LS_LEV + LEV_WORDS * e) * PBL_THREADS;
        q[0 * PBL_THREADS] = L.u; q[1 * PBL_THREADS] = L.v; q[2 * PBL_THREADS] = L.w;
        q[3 * PBL_THREADS] = L.rho; q[4 * PBL_THREADS] = L.rhograd;
        lsi(ls, LS_TAG + e) = n;
        if (do_lo) need_lo = false; else need_hi = false;
      }
    }
    const float *lo = ls + (LS_LEV + LEV_WORDS * e_lo) * PBL_THREADS,
                *hi = ls + (LS_LEV + LEV_WORDS * e_hi) * PBL_THREADS;

    // advance.f90:342-350
    const float dz = 1.f / (sh[indz] - sh[indz - 1]);
    const float dz1 = (zt - sh[indz - 1]) * dz;
    const float dz2 = (sh[indz] - zt) * dz;
    const float u = dz1 * hi[0 * PBL_THREADS] + dz2 * lo[0 * PBL_THREADS];
    const float v = dz1 * hi[1 * PBL_THREADS] + dz2 * lo[1 * PBL_THREADS];
    float w = dz1 * hi[2 * PBL_THREADS] + dz2 * lo[2 * PBL_THREADS];
    const float rhoa = dz1 * hi[3 * PBL_THREADS] + dz2 * lo[3 * PBL_THREADS];
    const float rhograd = dz1 * hi[4 * PBL_THREADS] + dz2 * lo[4 * PBL_THREADS];

    if (turbswitch) hanna(t, zt, regime); else hanna1(t, zt, regime);

    // horizontal turbulent velocities, advance.f90:371-384
    // (requesting the normals before the level search costs 4 %: measured)
    if (nrand + 1 > maxrand) nrand = 1;
    const float r_up = normal(a, nrand), r_vp = normal(a, nrand + 1);
    nrand = nrand + 2;
    if (nrand + ifine > maxrand) nrand = 1;
    // first vertical normal requested early so its latency hides behind the u/v update
    float r_w = (CBL && c.cblflag == 1) ? 0.f : normal(a, nrand + 1);
    if (dt / t.tlu < .5f) {
      up = (1.f - dt / t.tlu) * up + r_up * t.sigu * m_sqrt(2.f * dt / t.tlu);
    } else {
      const float ru = m_exp(-dt / t.tlu);
      up = ru * up + r_up * t.sigu * m_sqrt(1.f - ru * ru);
    }
    if (dt / t.tlv < .5f) {
      vp = (1.f - dt / t.tlv) * vp + r_vp * t.sigv * m_sqrt(2.f * dt / t.tlv);
    } else {
      const float rv = m_exp(-dt / t.tlv);
      vp = rv * vp + r_vp * t.sigv * m_sqrt(1.f - rv * rv);
    }

    const float rhoaux = rhograd / rhoa;
    const float dtf = dt * c.fine;
    const float dtftlw = dtf / t.tlw;

    // vertical component in ifine short steps, advance.f90:396-498
#pragma unroll
    for (int i = 1; i <= ifine; i++) { // unrolled in the SPEC variant only (constant trip count)
      float delz;
      // next iteration's normal (table modes: plain loads, harmless past the end)
      const float r_w_next = (CBL && c.cblflag == 1) ? 0.f : normal(a, nrand + i + 1);
      if (turbswitch) {
        if (dtftlw < .5f) {
          if (CBL && c.cblflag == 1) {
            if (-t.h / t.ol > 5.f) {
              int flagrein = 0;
              nrand = nrand + 1;
              float old_wp_buf = wp, ath, bth;
              cbl_drift(c, wp, zt, t.wst, t.h, rhoa, rhograd, t.sigw, t.dsigwdz, t.tlw, t.ol,
                        ath, bth, flagrein);
#ifdef FPB_DEBUG_NAN
              if (isnan(ath) || isnan(bth))
                printf("cbl_drift nan: wp %g zt %g wst %g h %g rhoa %g rhograd %g sigw %g dsigwdz %g tlw %g ol %g -> ath %g bth %g flag %d\n",
                       wp, zt, t.wst, t.h, rhoa, rhograd, t.sigw, t.dsigwdz, t.tlw, t.ol, ath, bth, flagrein);
#endif
              wp = (wp + ath * dtf + bth * normal(a, nrand) * m_sqrt(dtf)) * (float)icbt;
              delz = wp * dtf;
              if (flagrein == 1) {
                cbl_reinitialize(c, rng, zt, t.wst, t.h, t.sigw, t.ol, old_wp_buf, nrand);
                wp = old_wp_buf;
                delz = wp * dtf;
                nan_cbl++;
              }
            } else {
              nrand = nrand + 1;
              const float ath = -wp / t.tlw + t.sigw * t.dsigwdz + wp * wp / t.sigw * t.dsigwdz +
                                t.sigw * t.sigw / rhoa * rhograd;
              const float bth = t.sigw * normal(a, nrand) * m_sqrt(2.f * dtftlw);
              wp = (wp + ath * dtf + bth) * (float)icbt;
              delz = wp * dtf;
              const float del_test = (1.f - wp) / wp;
              if (isnan(wp) || isnan(del_test)) {
                nrand = nrand + 1;
                wp = t.sigw * normal(a, nrand);
                delz = wp * dtf;
                nan_cbl++;
              }
            }
          } else {
            wp = ((1.f - dtftlw) * wp + r_w * m_sqrt(2.f * dtftlw) +
                  dtf * (t.dsigwdz + rhoaux * t.sigw)) * (float)icbt;
            delz = wp * t.sigw * dtf;
          }
        } else {
          const float rw = m_exp(-dtftlw);
          const float r_ = (CBL && c.cblflag == 1) ? normal(a, nrand + i) : r_w;
          wp = (rw * wp + r_ * m_sqrt(1.f - rw * rw) +
                t.tlw * (1.f - rw) * (t.dsigwdz + rhoaux * t.sigw)) * (float)icbt;
          delz = wp * t.sigw * dtf;
        }
      } else {
        const float rw = m_exp(-dtftlw);
        wp = (rw * wp + r_w * m_sqrt(1.f - rw * rw) * t.sigw +
              t.tlw * (1.f - rw) * (t.dsigw2dz + rhoaux * (t.sigw * t.sigw))) * (float)icbt;
        delz = wp * dtf;
      }
      r_w = r_w_next;
      if (turboff) { up = 0.f; vp = 0.f; wp = 0.f; delz = 0.f; }
#ifdef FPB_DEBUG_NAN
      if ((isnan(delz) || isnan(up) || isnan(vp)) && !isnan(zt) && !isnan(t.sigw))
        printf("substep nan: i %d wp %g up %g vp %g delz %g zt %g dt %g dtf %g tlw %g tlu %g sigu %g sigw %g dsigwdz %g h %g ol %g wst %g ust %g rhoa %g rhograd %g nrand %d\n",
               i, wp, up, vp, delz, zt, dt, dtf, t.tlw, t.tlu, t.sigu, t.sigw, t.dsigwdz, t.h, t.ol, t.wst, t.ust, rhoa, rhograd, nrand);
#endif

      if (fabsf(delz) > t.h) delz = rare_fmodf(delz, t.h); // almost never taken: keep fmodf's body out of line
      if (delz < -zt) {               // reflection at the ground
        icbt = -1;
        zt = -zt - delz;
      } else if (delz > (t.h - zt)) { // reflection at h
        icbt = -1;
        zt = -zt - delz + 2.f * t.h;
      } else {
        icbt = 1;
        zt = zt + delz;
      }
      if (i != ifine) {
        t.zeta = zt / t.h;
        hanna_short(t, zt, regime);
      }
    }
    if (!(CBL && c.cblflag == 1)) nrand = nrand + (ifine + 1);
    ust = t.ust; // hanna* may raise ust to 1e-4 (idempotent)

    // next time step, advance.f90:504-510
    if (turbswitch) {
      float q = fminf(t.tlw, t.h / fmaxf(2.f * fabsf(wp * t.sigw), 1.e-5f));
      q = fminf(q, 0.5f / fabsf(t.dsigwdz));
      ldt = f_int(q * c.ctl);
    } else {
      const float q = fminf(t.tlw, t.h / fmaxf(2.f * fabsf(wp), 1.e-5f));
      ldt = f_int(q * c.ctl);
    }
    ldt = max(ldt, c.mintime);

    if (EXTRA) w = w + settling_term(a, sh, npoint, xtf, ytf, zt);

    dxsave = dxsave + u * dt;
    dysave = dysave + v * dt;
    dawsave = dawsave + up * dt;
    dcwsave = dcwsave + vp * dt;
    zt = zt + w * dt * (float)c.ldirect;

    const float ztop = sh[nz - 1];
    if (zt >= ztop) zt = ztop - 100.f * c.eps;

    if (zt > t.h) {
      if (itimec == itime + c.lsynctime) {
        // "defined" behaviour for the stale-usig case
        finish(a, false, u, v, w);
      } else {
        finish(a, true, u, v, w);
      }
      return;
    }

    // dry-deposition probability, advance.f90:582-599
    if (EXTRA && c.drydep && (zt < 2.f * HREF)) {
#pragma unroll
      for (int ks = 0; ks < FPB_MAXSPEC; ks++) {
        if (ks < c.nspec && c.drydepspec[ks]) {
          if (depo_todo & (1u << ks)) { // interpol_vdep, src/interpol_vdep.f90:39-54
            Hz z;
            load_weights(c, ls, z);
            const int off = ks * z.plane;
            const DevMetSlot *met = grid_met(a);
            const float y0 = bil(z, __ldg(met[0].vdep + off + z.o00), __ldg(met[0].vdep + off + z.o10),
                                 __ldg(met[0].vdep + off + z.o01), __ldg(met[0].vdep + off + z.o11));
            const float y1 = bil(z, __ldg(met[1].vdep + off + z.o00), __ldg(met[1].vdep + off + z.o10),
                                 __ldg(met[1].vdep + off + z.o01), __ldg(met[1].vdep + off + z.o11));
            vdepo[ks] = (y0 * z.dt2 + y1 * z.dt1) * z.dtt;
            depo_todo &= ~(1u << ks);
          }
          prob[ks] = 1.f + (prob[ks] - 1.f) * m_exp(-vdepo[ks] * fabsf(dt) / (2.f * HREF));
        }
      }
    }

    if (zt < 0.f) zt = fminf(t.h - EPS2, -1.f * zt);

    if (itimec == (itime + c.lsynctime)) finish(a, false, u, v, w);
  }
};

// Persistent kernel: every warp pulls batches of particle rows from
// *a.work_counter; lanes run sub-steps until their particle leaves the loop.
template <bool EXTRA, bool CBL, bool SPEC>
#ifndef FPB_CBL_BLOCKS
#define FPB_CBL_BLOCKS 4
#endif
#ifndef FPB_EXTRA_NOCBL_BLOCKS
#define FPB_EXTRA_NOCBL_BLOCKS 3
#endif
__global__ void __launch_bounds__(PBL_THREADS, (EXTRA ? (CBL ? FPB_CBL_BLOCKS : FPB_EXTRA_NOCBL_BLOCKS) : FPB_PBL_MIN_BLOCKS) * (128 / PBL_THREADS))
fpb_pbl_kernel(const __grid_constant__ DevStepArgs a) {
  const DevCfg &c = a.cfg;
  __shared__ float sh[FPB_MAXNZ];
  __shared__ float lane_rows[LS_WORDS * PBL_THREADS];
  for (int i = threadIdx.x; i < c.nz; i += blockDim.x) sh[i] = a.height[i];
  __syncthreads();
  float *ls = lane_rows + threadIdx.x;

  constexpr unsigned FULL = 0xffffffffu;
#ifndef FPB_T_REFILL
#define FPB_T_REFILL 8
#endif
  constexpr int T_REFILL = FPB_T_REFILL; // refill when at least this many lanes are idle
  const int lane = threadIdx.x & 31;
  const unsigned lt_mask = (1u << lane) - 1u;
  const int nrows = c.numpart;

  PblTask<EXTRA, CBL, SPEC> task;
  task.running = false;
  task.j = -1;
  unsigned n_act = 0, n_pbl = 0, n_sub = 0, n_nan = 0;
  bool exhausted = false;

  // (Round 2, built, measured and removed again -- commit cd7639e has it: a variant of this loop that polls a
  // "rows ready" word and counts finished rows per chunk, so that ONE launch serves all chunks of fpb_step_host.
  // The host step got slower, DESIGN.md section 5, and the extra warp-uniform tests in this loop cost the
  // resident C2 step 3.6 %, profiles/ab_r02_pblloop.txt.)
  for (;;) {
    const unsigned idle = __ballot_sync(FULL, !task.running);
    if (!exhausted && __popc(idle) >= T_REFILL) {
      const int n = __popc(idle), leader = __ffs(idle) - 1;
      int base = 0;
      if (lane == leader) base = atomicAdd(a.work_counter, n);
      base = __shfl_sync(FULL, base, leader);
      if (!task.running) {
        const int row = base + __popc(idle & lt_mask);
        bool pbl = false;
        if (row < nrows && task.refill(a, ls, row, pbl)) {
          n_act++;
          if (pbl) n_pbl++;
        }
      }
      if (base + n >= nrows) exhausted = true;
      continue;
    }
    if (idle == FULL) break; // no rows left and nothing running
#ifdef FPB_TAIL_STATS // tools/tail_stats.py: warp iterations / running lanes before and after the rows ran out
    if (lane == 0) {
      atomicAdd(a.stats + (exhausted ? 7 : 6), 1ull);
      if (exhausted) atomicAdd(a.stats + 2, (unsigned long long)(32 - __popc(idle)));
    }
#endif
    if (task.running) {
      task.substep(a, sh, ls);
      if (!task.running) {
        n_sub += task.nsub;
        n_nan += task.nan_cbl;
      }
    }
  }

  if (a.stats) {
    const unsigned long long v0 = warp_sum(n_act), v3 = warp_sum(n_pbl), v4 = warp_sum(n_sub),
                             v6 = warp_sum(n_nan);
    if (lane == 0 && v0) {
      atomicAdd(a.stats + 0, v0);
      if (v3) atomicAdd(a.stats + 3, v3);
      if (v4) atomicAdd(a.stats + 4, v4);
      if (v6) atomicAdd(a.stats + 6, v6);
    }
  }
}

// -------------------------------------------------------- finish kernel ----
// label 700 + label 99 + Petterssen (src/advance.f90:629-985) and the rest of
// the timemanager loop body (src/timemanager.f90:630-707) for row j.
// SIMPLE = true: no nested input grids, no settling, no dry deposition, table RNG -- known at
// compile time (the launcher checks); the general variant covers everything else.
template <bool SIMPLE>
__device__ __forceinline__ void finish_row(const DevStepArgs &a, const float *sh, int j,
                                           unsigned &n_term, unsigned &n_pett) {
  const DevCfg &c = a.cfg;
  const int itime = c.itime, nz = c.nz, maxrand = c.maxrand;
  const float eps = c.eps;
  const float ztop = sh[nz - 1];
  const int npoint = a.p.npoint[j];
  const int slot = a.p.slot[j];
  const int flags = a.sc.flags[j];

  double xt = a.p.xtra1[j], yt = a.p.ytra1[j];
  float zt = a.p.ztra1[j];
  float wp = a.p.uzp[j];
  float usigold = a.p.us[j], vsigold = a.p.vs[j], wsigold = a.p.ws[j];
  int ldt = a.p.idt[j];

  Rng rng;
  make_rng(c, a.rannumb, slot, rng);
  auto normal = [&](int i) -> float { return SIMPLE ? __ldg(a.rannumb + (i - 1)) : rng.get(i); };

  float dxsave = 0.f, dysave = 0.f, dawsave = 0.f, dcwsave = 0.f;
  float u = 0.f, v = 0.f, w = 0.f, usig = 0.f, vsig = 0.f, wsig = 0.f;
  int nrand, itimec = itime, indz_last = 1;
  if (flags & SC_PBL) {
    const float4 s0 = a.sc.s0[j], s1 = a.sc.s1[j];
    const int2 s2 = a.sc.s2[j];
    dxsave = s0.x; dysave = s0.y; dawsave = s0.z; dcwsave = s0.w;
    u = s1.x; v = s1.y; w = s1.z;
    indz_last = __float_as_int(s1.w);
    nrand = s2.x;
    itimec = s2.y;
  } else {
    nrand = advance_nrand(a, slot);
  }

  // advance.f90:199-253 at the position the call started from
  Hz z;
  z.ngrid = SIMPLE ? pole_grid(c, yt) : choose_grid(c, xt, yt);
  const DevMetSlot *met;
  float h = 0.f, tropop;
  {
    const GridSel g = select_grid(a, SIMPLE ? min(z.ngrid, 0) : z.ngrid, xt, yt);
    met = g.met;
    int jyp = g.jy + 1;
    if (jyp >= c.nymax) jyp = jyp - 1;
    make_weights(c, z, itime, g.xf, g.yf, g.ix, g.jy, g.ix + 1, jyp, g.nxd, g.nyd);
#pragma unroll
    for (int m = 0; m < 2; m++) {
      const float v0 = __ldg(met[m].S + z.o00).x, v1 = __ldg(met[m].S + z.o10).x;
      const float v2 = __ldg(met[m].S + z.o01).x, v3 = __ldg(met[m].S + z.o11).x;
      if (v0 > h) h = v0;
      if (v1 > h) h = v1;
      if (v2 > h) h = v2;
      if (v3 > h) h = v3;
    }
    tropop = __ldg(g.trop_lit1 + g.nix + g.nxd * g.njy); // slot 1 literal, advance.f90:253,263
  }

  float ux = 0.f, vy = 0.f;
  int nstop = 0;

  if ((flags & (SC_PBL | SC_ABOVE)) == SC_PBL) {
    // usig = 0.5*(usigprof(indzp)+usigprof(indz)) etc. for the level pair of the last
    // sub-step, advance.f90:604-606 (and the "defined" stale-usig case)
    float a0, a1, a2, b0, b1, b2;
    profile_sigma(c, met, z, indz_last, a0, a1, a2);
    profile_sigma(c, met, z, indz_last + 1, b0, b1, b2);
    usig = 0.5f * (b0 + a0);
    vsig = 0.5f * (b1 + a1);
    wsig = 0.5f * (b2 + a2);
  }

  if (flags & SC_ABOVE) { // label 700, advance.f90:629-708
    indz_last = 0;
    interp_wind<true>(c, met, z, sh, zt, u, v, w, usig, vsig, wsig, indz_last);
    ldt = abs(c.lsynctime - itimec + itime);
    const float dt = (float)ldt;
    if (zt < tropop) {
      const float uxscale = m_sqrt(2.f * c.d_trop / dt);
      if (nrand + 1 > maxrand) nrand = 1;
      ux = normal(nrand) * uxscale;
      vy = normal(nrand + 1) * uxscale;
      nrand = nrand + 2;
      wp = 0.f;
    } else if (zt < tropop + 1000.f) {
      const float weight = (zt - tropop) / 1000.f;
      const float uxscale = m_sqrt(2.f * c.d_trop / dt * (1.f - weight));
      if (nrand + 2 > maxrand) nrand = 1;
      ux = normal(nrand) * uxscale;
      vy = normal(nrand + 1) * uxscale;
      const float wpscale = m_sqrt(2.f * c.d_strat / dt * weight);
      wp = normal(nrand + 2) * wpscale + c.d_strat / 1000.f;
      nrand = nrand + 3;
    } else {
      if (nrand > maxrand) nrand = 1;
      ux = 0.f;
      vy = 0.f;
      const float wpscale = m_sqrt(2.f * c.d_strat / dt);
      wp = normal(nrand) * wpscale;
      nrand = nrand + 1;
    }
    if (c.turboff) { ux = 0.f; vy = 0.f; wp = 0.f; }

    if (!SIMPLE) w = w + settling_term(a, sh, npoint, (float)xt, (float)yt, zt);

    dxsave = dxsave + (u + ux) * dt;
    dysave = dysave + (v + vy) * dt;
    zt = zt + (w + wp) * dt * (float)c.ldirect;
    if (zt < 0.f) zt = fminf(h - EPS2, -1.f * zt);
  }

  // label 99: mesoscale fluctuations, advance.f90:728-739
  {
    const float r = m_exp(-2.f * (float)abs(c.lsynctime) / (float)c.lwindinterv);
    const float rs = m_sqrt(1.f - r * r);
    if (nrand + 2 > maxrand) nrand = 1;
    usigold = r * usigold + rs * normal(nrand) * usig * c.turbmesoscale;
    vsigold = r * vsigold + rs * normal(nrand + 1) * vsig * c.turbmesoscale;
    wsigold = r * wsigold + rs * normal(nrand + 2) * wsig * c.turbmesoscale;
    dxsave = dxsave + usigold * (float)c.lsynctime;
    dysave = dysave + vsigold * (float)c.lsynctime;
    zt = zt + wsigold * (float)c.lsynctime;
    if (zt < 0.f) zt = -1.f * zt;
  }

  // advance.f90:747-778
  windalign(dxsave, dysave, dawsave, dcwsave, ux, vy);
  dxsave = dxsave + ux;
  dysave = dysave + vy;
  const int ngrid = z.ngrid;
  move_horizontal(c, ngrid, xt, yt, dxsave, dysave, (float)c.ldirect);

  bool done = false;
  if (wrap_and_check(c, xt, yt)) {
    nstop = 3;
    done = true;
  }
  if (!done) {
    if (zt >= ztop) zt = ztop - 100.f * eps;
    // Petterssen corrector, advance.f90:829-985
    if (ldt != abs(c.lsynctime)) done = true;
    else if (abs(itime + ldt * c.ldirect) > abs(c.memtime[1])) done = true;
    else if ((SIMPLE ? pole_grid(c, yt) : choose_grid(c, xt, yt)) != ngrid) done = true;
  }
  if (!done) {
    const GridSel g = select_grid(a, SIMPLE ? min(ngrid, 0) : ngrid, xt, yt); // advance.f90:862-870
    int jyp = g.jy + 1;
    if (jyp >= c.nymax) jyp = jyp - 1;
    const float uold = u, vold = v, wold = w;
    make_weights(c, z, itime + ldt * c.ldirect, g.xf, g.yf, g.ix, g.jy, g.ix + 1, jyp, g.nxd, g.nyd);
    float d0, d1, d2;
    interp_wind<false>(c, met, z, sh, zt, u, v, w, d0, d1, d2, indz_last); // hint: the level before the move
    n_pett++;
    if (!SIMPLE) w = w + settling_term(a, sh, npoint, (float)xt, (float)yt, zt);
    u = (u - uold) / 2.f;
    v = (v - vold) / 2.f;
    w = (w - wold) / 2.f;
    zt = zt + w * (float)(ldt * c.ldirect);
    if (zt < 0.f) zt = fminf(h - EPS2, -1.f * zt);
    move_horizontal(c, ngrid, xt, yt, u, v, (float)(ldt * c.ldirect));
    if (wrap_and_check(c, xt, yt)) {
      nstop = 3;
    } else if (zt >= ztop) {
      zt = ztop - 100.f * eps;
    }
  }

  // A position that is not finite (the CBL closure is singular where its transition factor
  // vanishes, src/initialize_cbl_vel.f90:50-63: the reference carries the NaN on) would index
  // outside every grid downstream: the particle is terminated and counted instead.
  if (!(isfinite(xt) && isfinite(yt) && isfinite(zt))) {
    nstop = 4;
    if (a.stats) atomicAdd(a.stats + 7, 1ull);
  }

  // ---- rest of the timemanager loop body, src/timemanager.f90:630-707
  const int itramem = a.p.itramem[j];
  int itra1;
  if (nstop > 1) {
    itra1 = FPB_ITRA_DEAD;
    n_term++;
    if (!SIMPLE && c.linit_cond) a.sc.flags[j] = flags | SC_TERM_NSTOP; // src/timemanager.f90:631 (launcher: general variant)
  } else {
    bool term = false;
    itra1 = itime + c.lsynctime;
    float xmassfract = 0.f;
    float drydeposit[SIMPLE ? 1 : FPB_MAXSPEC];
    for (int ks = 0; ks < c.nspec; ks++) {
      float xm1 = a.p.xmass1[(size_t)ks * a.p.maxpart + j];
      const float decfact = (c.decay[ks] > 0.f) ? m_exp(-(float)abs(c.lsynctime) * c.decay[ks]) : 1.f;
      if (!SIMPLE) drydeposit[ks] = 0.f;
      if (!SIMPLE && c.drydepspec[ks]) {
        const float pr = (c.drydep && (flags & SC_PBL)) ? a.sc.prob[(size_t)ks * a.p.maxpart + j] : 0.f;
        drydeposit[ks] = xm1 * pr * decfact;
        xm1 = xm1 * (1.f - pr) * decfact;
        if (c.decay[ks] > 0.f)
          drydeposit[ks] = drydeposit[ks] * m_exp((float)abs(c.ldeltat) * c.decay[ks]);
      } else {
        xm1 = xm1 * decfact;
      }
      a.p.xmass1[(size_t)ks * a.p.maxpart + j] = xm1;
      if (c.mdomainfill == 0 && c.mquasilag == 0) {
        const float xm = __ldg(a.xmass + ks * c.numpoint + (npoint - 1));
        if (xm > 0.f)
          xmassfract = fmaxf(xmassfract, (float)__ldg(a.npart + npoint - 1) * xm1 / xm);
      } else {
        xmassfract = 1.0f;
      }
    }
    if (xmassfract < MINMASS) { itra1 = FPB_ITRA_DEAD; term = true; }

    if (!SIMPLE && c.drydep && (c.ldirect == 1)) {
      const int kp = (c.ioutputforeachrelease == 1) ? npoint : 1;
      const int itage = abs(itime - itramem);
      int nage;
      for (nage = 1; nage <= c.nageclass; nage++)
        if (itage < c.lage[nage - 1]) break;
      const int nclass = a.p.nclass[j];
      const int rslot = slot - a.dep.slot_base;
      drydepo_scatter(c, a.drygridunc, false, nclass, drydeposit, (float)xt, (float)yt, nage, kp, a.dep, rslot);
      if (c.nested_output == 1)
        drydepo_scatter(c, a.drygriduncn, true, nclass, drydeposit, (float)xt, (float)yt, nage, kp, a.dep, rslot);
    }
    if (abs(itra1 - itramem) >= c.lage[c.nageclass - 1]) {
      if (!SIMPLE && c.linit_cond && itra1 != FPB_ITRA_DEAD) a.sc.flags[j] = flags | SC_TERM_AGE; // :702
      itra1 = FPB_ITRA_DEAD;
      term = true;
    }
    if (term) n_term++;
  }

  a.p.xtra1[j] = xt;
  a.p.ytra1[j] = yt;
  a.p.ztra1[j] = zt;
  a.p.itra1[j] = itra1;
  a.p.idt[j] = ldt;
  if (flags & SC_ABOVE) a.p.uzp[j] = wp; // uap, ucp, cbt: unchanged above the PBL
  a.p.us[j] = usigold; a.p.vs[j] = vsigold; a.p.ws[j] = wsigold;
}

#ifndef FPB_FINISH_MIN_BLOCKS
#define FPB_FINISH_MIN_BLOCKS 8
#endif
template <bool SIMPLE>
__global__ void __launch_bounds__(128, FPB_FINISH_MIN_BLOCKS)
fpb_finish_kernel(const __grid_constant__ DevStepArgs a) {
  const DevCfg &c = a.cfg;
  __shared__ float sh[FPB_MAXNZ];
  for (int i = threadIdx.x; i < c.nz; i += blockDim.x) sh[i] = a.height[i];
  __syncthreads();
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned n_term = 0, n_pett = 0;
  if (j < c.numpart && a.p.itra1[j] == c.itime) finish_row<SIMPLE>(a, sh, j, n_term, n_pett);
  if (a.stats) {
    const unsigned long long v2 = warp_sum(n_term), v5 = warp_sum(n_pett);
    if ((threadIdx.x & 31) == 0) {
      if (v2) atomicAdd(a.stats + 2, v2);
      if (v5) atomicAdd(a.stats + 5, v5);
    }
  }
}
