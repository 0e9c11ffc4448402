// timemanager.cpp -- the reference's time loop (src/timemanager.f90:152-729)
// with the engine in place of the per-particle loop and conccalc.
//
// Kept from the reference, in its order: getfields' two-slot rotation
// (src/getfields.f90:96-176), particle release, decay of deposited mass at
// loutnext, the sampling schedule with half weights at the window ends, the
// output + second conccalc at loutend, the exit at ideltas, ldeltat, the
// particle loop; wet deposition, domain filling (init_domainfill / boundcond_domainfill),
// convective mixing and particle splitting where the run asks for them.  Left out (outside the
// hot-path scope): OH reaction, flux and trajectory output, particle dumps.
#include <chrono>
#include <cmath>
#include <cstring>
#include <vector>

#include "fpbh_internal.h"

namespace {
struct MetStore { // com_mod met arrays, numwfmem = 2
  std::vector<float> f3[2][8], f2[2][5], vd[2], rain[2][4];
  std::vector<int8_t> cl[2];
  fpb_met_ptrs ptr[2];
  void alloc(const fpb_config &c) {
    const size_t n3 = (size_t)c.nxmax * c.nymax * c.nzmax, n2 = (size_t)c.nxmax * c.nymax;
    for (int s = 0; s < 2; s++) {
      for (auto &v : f3[s]) v.assign(n3, 0.f);
      for (auto &v : f2[s]) v.assign(n2, 0.f);
      vd[s].assign(n2 * c.maxspec, 0.f);
      fpb_met_ptrs &m = ptr[s];
      m = fpb_met_ptrs{};
      if (c.wetdep) { // lsprec, convprec, tcc, ctwc, clouds (src/com_mod.f90:376-395)
        for (auto &v : rain[s]) v.assign(n2, 0.f);
        cl[s].assign(n3, 0);
        m.lsprec = rain[s][0].data(); m.convprec = rain[s][1].data(); m.tcc = rain[s][2].data();
        m.ctwc = rain[s][3].data(); m.clouds = cl[s].data();
      }
      m.uu = f3[s][0].data(); m.vv = f3[s][1].data(); m.ww = f3[s][2].data(); m.rho = f3[s][3].data();
      m.drhodz = f3[s][4].data(); m.tt = f3[s][5].data(); m.uupol = f3[s][6].data(); m.vvpol = f3[s][7].data();
      m.hmix = f2[s][0].data(); m.ustar = f2[s][1].data(); m.wstar = f2[s][2].data();
      m.oli = f2[s][3].data(); m.tropopause = f2[s][4].data();
      m.vdep = vd[s].data();
    }
  }
};
struct RawStore { // one time level of model-level fields (readwind_ecmwf's uuh.. and com_mod's tth, ps, ..)
  std::vector<float> f3[5], f2[8];
  fpb_rawmet_ptrs ptr{};
  void alloc(const fpb_config &c) {
    const size_t n3 = (size_t)c.nxmax * c.nymax * c.nzmax, n2 = (size_t)c.nxmax * c.nymax;
    for (auto &v : f3) v.assign(n3, 0.f);
    for (auto &v : f2) v.assign(n2, 0.f);
    ptr.uuh = f3[0].data(); ptr.vvh = f3[1].data(); ptr.tth = f3[2].data(); ptr.qvh = f3[3].data(); ptr.wwh = f3[4].data();
    ptr.pvh = nullptr;
    ptr.ps = f2[0].data(); ptr.tt2 = f2[1].data(); ptr.td2 = f2[2].data(); ptr.sshf = f2[3].data();
    ptr.surfstr = f2[4].data(); ptr.lsprec = f2[5].data(); ptr.convprec = f2[6].data(); ptr.tcc = f2[7].data();
    ptr.excessoro = nullptr;
  }
};
double now() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
} // namespace

extern "C" int fpbh_timemanager(const fpb_config *cp, const float *height, const fpbh_releases *rel,
                                const fpbh_run *run, const fpbh_engine *eng, fpbh_output_fn out,
                                void *user, fpbh_run_result *res) {
  if (!cp || !height || !rel || !run || !eng) return fpbh_fail("fpbh_timemanager: null argument");
  const fpb_config &c = *cp;
  if (c.lsynctime == 0) return fpbh_fail("fpbh_timemanager: lsynctime == 0");
  if (run->met_interval <= 0) return fpbh_fail("fpbh_timemanager: met_interval <= 0");
  const int ldirect = c.ldirect, lsynctime = c.lsynctime;

  // host mirror of the particle arrays (com_mod.f90:675-695)
  const size_t mp = (size_t)c.maxpart;
  std::vector<double> xtra1(mp), ytra1(mp);
  std::vector<float> ztra1(mp), uap(mp), ucp(mp), uzp(mp), us(mp), vs(mp), ws(mp);
  std::vector<int32_t> itra1(mp, FPB_ITRA_DEAD), npoint(mp), nclass(mp), idt(mp), itramem(mp), itrasplit(mp);
  std::vector<int16_t> cbt(mp);
  std::vector<float> xmass1(mp * c.nspec), xscav(mp * c.nspec);
  fpb_particle_ptrs P{};
  P.xtra1 = xtra1.data(); P.ytra1 = ytra1.data(); P.ztra1 = ztra1.data();
  P.itra1 = itra1.data(); P.npoint = npoint.data(); P.nclass = nclass.data(); P.idt = idt.data();
  P.itramem = itramem.data(); P.itrasplit = itrasplit.data();
  P.uap = uap.data(); P.ucp = ucp.data(); P.uzp = uzp.data();
  P.us = us.data(); P.vs = vs.data(); P.ws = ws.data(); P.cbt = cbt.data();
  P.xmass1 = xmass1.data(); P.xscav_frac1 = xscav.data(); P.ld = c.maxpart;
  fpb_particle_ptrs Ponly_itra{};
  Ponly_itra.itra1 = itra1.data();
  Ponly_itra.ld = c.maxpart;
  int32_t numpart = 0;
  bool releases_set = false;

  fpbh_release_state *rst = fpbh_release_state_new(rel->numpoint);
  MetStore met;
  RawStore raw;
  std::vector<float> akm, bkm, akz, bkz;
  if (run->met_raw) {
    if (!eng->set_vertical || !eng->calcpar_verttransform)
    { fpbh_release_state_free(rst); return fpbh_fail("fpbh_timemanager: met_raw needs set_vertical / calcpar_verttransform in the engine table"); }
    raw.alloc(c);
    akm.resize(c.nz); bkm.resize(c.nz); akz.resize(c.nz); bkz.resize(c.nz);
  } else {
    met.alloc(c);
  }
  if (run->lconvection && (!run->met_raw || !eng->set_convection || !eng->convmix))
  { fpbh_release_state_free(rst); return fpbh_fail("fpbh_timemanager: lconvection needs met_raw and set_convection / convmix in the engine table"); }
  if (c.mdomainfill >= 1 && (!eng->init_domainfill || !eng->boundcond_domainfill))
  { fpbh_release_state_free(rst); return fpbh_fail("fpbh_timemanager: mdomainfill needs init_domainfill / boundcond_domainfill in the engine table"); }

  // grids handed to the output callback (reference layout, maxspec)
  const size_t outer = (size_t)c.maxspec * c.maxpointspec_act * c.nclassunc * c.maxageclass;
  std::vector<float> gridunc((size_t)c.numxgrid * c.numygrid * c.numzgrid * outer);
  std::vector<float> drygridunc((size_t)c.numxgrid * c.numygrid * outer);
  std::vector<float> griduncn, drygriduncn;
  if (c.nested_output == 1) {
    griduncn.resize((size_t)c.numxgridn * c.numygridn * c.numzgrid * outer);
    drygriduncn.resize((size_t)c.numxgridn * c.numygridn * outer);
  }
  std::vector<float> creceptor((size_t)FPB_MAXRECEPTOR * c.maxspec);

  fpbh_run_result R{};
  int rc = 0;
#define ENG(call)                                                    \
  do {                                                               \
    if ((call) != 0) {                                               \
      rc = fpbh_fail("fpbh_timemanager: engine call failed: " #call); \
      goto done;                                                     \
    }                                                                \
  } while (0)

  {
    // src/timemanager.f90:119-122
    int loutnext = run->loutstep / 2;
    float outnum = 0.f;
    int loutstart = loutnext - run->loutaver / 2;
    int loutend = loutnext + run->loutaver / 2;
    const float outstep = (float)std::abs(run->loutstep);

    int memind[2] = {1, 2};
    int memtime[2] = {999999999, 999999999};
    bool have_fields = false;
    const bool DEP = c.drydep != 0 || c.wetdep != 0; // src/readreleases.f90:389
    if (run->met_raw) { // gridcheck's vertical structure, then the engine's copies of it
      int32_t nconvlev = 0;
      if (fpbh_synth_hybrid_levels(c.nz, akm.data(), bkm.data(), akz.data(), bkz.data(), &nconvlev)) { rc = 1; goto done; }
      ENG(eng->set_vertical(eng->self, c.nz, c.nz, c.nzmax, c.nzmax, akm.data(), bkm.data(), akz.data(), bkz.data()));
      if (run->lconvection)
        ENG(eng->set_convection(eng->self, c.nz, c.nzmax, nconvlev, akz.data(), bkz.data(), akm.data(), bkm.data()));
    }
    auto do_convmix = [&](int itime) -> int {
      int32_t ncol = 0, nconv = 0;
      if (eng->convmix(eng->self, itime, &ncol, &nconv)) return 1;
      R.convmix_calls++;
      R.convecting_columns += nconv;
      return 0;
    };

    for (int itime = 0; ldirect * itime <= ldirect * run->ideltas; itime += lsynctime) {
      // ---- wet deposition, src/timemanager.f90:164-169: before new fields are read in,
      // never at the very beginning; ldeltat as in src/wetdepo.f90:55-63
      if (c.wetdep && itime != 0 && numpart > 0 && eng->wetdepo) {
        const int ldw = (itime <= loutnext) ? itime - (loutnext - run->loutstep) : itime - loutnext;
        ENG(eng->wetdepo(eng->self, itime, lsynctime, ldw));
      }

      // ---- convection for backward runs, src/timemanager.f90:183-193: before the next field is read
      if (ldirect == -1 && run->lconvection && itime < 0) ENG(do_convmix(itime));

      // ---- getfields, src/getfields.f90:96-176 (wind fields every
      // met_interval seconds, times counted in the run's direction)
      {
        const int mi = run->met_interval;
        auto synth = [&](int slot, int t) -> int {
          if (run->met_raw) { // readwind -> calcpar -> verttransform, the last two on the device
            if (fpbh_synth_rawmet(&c, c.nz, akz.data(), bkz.data(), t, &raw.ptr)) return 1;
            return eng->calcpar_verttransform(eng->self, slot, &raw.ptr, 0, nullptr);
          }
          if (run->met_homogeneous) {
            if (fpbh_homogeneous_met(&c, run->met_u, run->met_v, run->met_w, &met.ptr[slot - 1])) return 1;
          } else if (fpbh_synth_met(&c, height, t, &met.ptr[slot - 1])) {
            return 1;
          }
          return eng->upload_met(eng->self, slot, &met.ptr[slot - 1]);
        };
        if (have_fields && ldirect * memtime[0] <= ldirect * itime && ldirect * memtime[1] > ldirect * itime) {
          // fields in memory bracket itime
        } else if (have_fields && ldirect * memtime[1] <= ldirect * itime) {
          std::swap(memind[0], memind[1]);
          memtime[0] = memtime[1];
          // first field time strictly after itime (in run direction)
          int k = (int)std::floor((double)(ldirect * itime) / mi) + 1;
          memtime[1] = ldirect * k * mi;
          ENG(synth(memind[1], memtime[1]));
        } else {
          int k = (int)std::floor((double)(ldirect * itime) / mi);
          memind[0] = 1; memind[1] = 2;
          memtime[0] = ldirect * k * mi;
          memtime[1] = ldirect * (k + 1) * mi;
          ENG(synth(1, memtime[0]));
          ENG(synth(2, memtime[1]));
          have_fields = true;
        }
        const int lwindinterv = std::abs(memtime[1] - memtime[0]);
        ENG(eng->set_met_bracket(eng->self, memind, memtime, lwindinterv));
      }

      // ---- release particles, src/timemanager.f90:230-251
      if (c.mdomainfill >= 1) {
        if (itime == 0) {
          ENG(eng->init_domainfill(eng->self, rel->xpoint1[0], rel->ypoint1[0], rel->xpoint2[0], rel->ypoint2[0],
                                   rel->itsplit, &numpart, nullptr));
        } else {
          int32_t created = 0;
          ENG(eng->boundcond_domainfill(eng->self, itime, loutend, &numpart, &created));
          R.boundary_particles += created;
        }
      } else {
        bool due = false;
        for (int i = 0; i < rel->numpoint; i++)
          if (itime >= rel->ireleasestart[i] && itime <= rel->ireleaseend[i]) due = true;
        if (due && eng->releaseparticles) { // particles are created where they live
          if (!releases_set) {
            fpb_release_points rp{};
            rp.numpoint = rel->numpoint;
            rp.ireleasestart = rel->ireleasestart; rp.ireleaseend = rel->ireleaseend;
            rp.xpoint1 = rel->xpoint1; rp.ypoint1 = rel->ypoint1; rp.xpoint2 = rel->xpoint2;
            rp.ypoint2 = rel->ypoint2; rp.zpoint1 = rel->zpoint1; rp.zpoint2 = rel->zpoint2;
            rp.itsplit = rel->itsplit;
            rp.mp_pid = 0;
            ENG(eng->set_releases(eng->self, &rp));
            releases_set = true;
          }
          ENG(eng->releaseparticles(eng->self, itime, &numpart, nullptr));
        } else if (due) {
          if (numpart > 0) ENG(eng->pull_particles(eng->self, 0, numpart, &Ponly_itra));
          int32_t first = 0, n = 0;
          if (fpbh_releaseparticles(&c, height, rel, rst, itime, &P, &numpart, &first, &n)) {
            rc = 1;
            goto done;
          }
          // push the new rows only: with re-used slots [first, first + n) also spans live
          // particles, whose host copies are stale (only itra1 was refreshed above)
          for (int j = first; j < first + n;) {
            if (!(P.itra1[j] == itime && P.itramem[j] == itime)) { j++; continue; }
            int k = j;
            while (k < first + n && P.itra1[k] == itime && P.itramem[k] == itime) k++;
            ENG(eng->push_particles(eng->self, j, k - j, &P));
            j = k;
          }
          ENG(eng->set_numpart(eng->self, numpart));
        }
      }

      // ---- convective mixing for forward runs, src/timemanager.f90:258-263
      if (ldirect == 1 && run->lconvection) ENG(do_convmix(itime));

      // ---- decay of deposited mass, src/timemanager.f90:269-304
      if (DEP && itime == loutnext && ldirect > 0) {
        float f[FPB_MAXSPEC];
        bool any = false;
        for (int ks = 0; ks < c.nspec; ks++) {
          f[ks] = 1.f;
          if (c.decay[ks] > 0.f) {
            f[ks] = (float)std::exp((double)(-1.f * outstep * c.decay[ks]));
            any = true;
          }
        }
        if (any) ENG(eng->scale_depgrids(eng->self, f));
      }

      // ---- sampling, src/timemanager.f90:350-365
      if (ldirect * itime >= ldirect * loutstart && ldirect * itime <= ldirect * loutend) {
        if ((itime - loutstart) % run->loutsample == 0) {
          const float weight = (itime == loutstart || itime == loutend) ? 0.5f : 1.0f;
          outnum += weight;
          double t0 = now();
          ENG(eng->conccalc(eng->self, itime, weight));
          R.t_conc_s += now() - t0;
        }
        // ---- output, src/timemanager.f90:376-464
        if (itime == loutend && outnum > 0.f) {
          ENG(eng->fetch_grids(eng->self, gridunc.data(), griduncn.empty() ? nullptr : griduncn.data(),
                               drygridunc.data(), drygriduncn.empty() ? nullptr : drygriduncn.data(),
                               creceptor.data(), 1));
          if (out && out(user, itime, outnum, gridunc.data(), griduncn.empty() ? nullptr : griduncn.data(),
                         drygridunc.data(), drygriduncn.empty() ? nullptr : drygriduncn.data(),
                         creceptor.data())) {
            rc = fpbh_fail("fpbh_timemanager: output callback failed");
            goto done;
          }
          R.outputs++;
          outnum = 0.f;
          loutnext = loutnext + run->loutstep;
          loutstart = loutnext - run->loutaver / 2;
          loutend = loutnext + run->loutaver / 2;
          if (itime == loutstart) {
            const float weight = 0.5f;
            outnum += weight;
            double t0 = now();
            ENG(eng->conccalc(eng->self, itime, weight));
            R.t_conc_s += now() - t0;
          }
          // ---- particle splitting, src/timemanager.f90:472-503
          if (eng->split_particles && ldirect * itime >= ldirect * rel->itsplit) {
            if (!releases_set && eng->set_releases) { // (the engine's block-scan scratch)
              fpb_release_points rp{};
              rp.numpoint = rel->numpoint;
              rp.ireleasestart = rel->ireleasestart; rp.ireleaseend = rel->ireleaseend;
              rp.xpoint1 = rel->xpoint1; rp.ypoint1 = rel->ypoint1; rp.xpoint2 = rel->xpoint2;
              rp.ypoint2 = rel->ypoint2; rp.zpoint1 = rel->zpoint1; rp.zpoint2 = rel->zpoint2;
              rp.itsplit = rel->itsplit;
              rp.mp_pid = 0;
              ENG(eng->set_releases(eng->self, &rp));
              releases_set = true;
            }
            ENG(eng->split_particles(eng->self, itime, &numpart));
            R.split_calls++;
          }
        }
      }

      if (itime == run->ideltas) break; // src/timemanager.f90:509

      // src/timemanager.f90:514-518
      int ldeltat;
      if (itime < loutnext) ldeltat = itime - (loutnext - run->loutstep);
      else ldeltat = itime - loutnext;

      // ---- the particle loop, src/timemanager.f90:531-712
      {
        fpb_step_stats st{};
        double t0 = now();
        ENG(eng->step(eng->self, itime, ldeltat, &st));
        R.t_step_s += now() - t0;
        R.particle_steps += st.n_active;
        R.substeps += st.n_substeps;
        R.syncs++;
      }
      if (run->max_steps > 0 && R.syncs >= run->max_steps) break;
    }
  }
done:
  R.numpart_final = numpart;
  if (res) *res = R;
  fpbh_release_state_free(rst);
  return rc;
#undef ENG
}
