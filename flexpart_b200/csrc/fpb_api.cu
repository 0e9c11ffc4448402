// fpb_api.cu -- host side of the C ABI declared in include/fpb.h.
//
// Owns all device memory (met replica, particle SoA, grids), packs the
// Fortran-layout met arrays into the device layout, replays the reference's
// ran3 draw order in validation mode and launches the kernels of
// fpb_kernels.cu / fpb_scatter.cu on one CUDA stream.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "fpb_device.cuh"
#include "fpb_scatter.cuh"
#include "fpb_sort.cuh"
#include "fpb_output.cuh"
#include "fpb_domainfill.cuh"
#include "fpb_metproc.cuh"
#include "fpb_convmix.cuh"

// ------------------------------------------------------------ error state --
static thread_local std::string g_err;
static int fail(const char *fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  return 1;
}
#define CK(call)                                                              \
  do {                                                                        \
    cudaError_t e_ = (call);                                                  \
    if (e_ != cudaSuccess)                                                    \
      return fail("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_),     \
                  __FILE__, __LINE__);                                        \
  } while (0)

// ------------------------------------------------- host RNG (random_mod) --
// Knuth's subtractive generator as published in Numerical Recipes ("ran3"),
// which is what src/random_mod.f90:93-139 uses; needed on the host to fill
// rannumb (src/FLEXPART.f90:56-59) and to replay the per-call table index
// `nrand=int(ran3(idummy)*real(maxrand-1))+1` (src/advance.f90:153,
// src/initialize.f90:68) in the reference's particle order.
struct Ran3 {
  int ma[56];
  int inext = 0, inextp = 0;
  bool seeded = false;
  void seed(int idum) {
    const int MBIG = 1000000000, MSEED = 161803398;
    int mj = (MSEED - abs(idum)) % MBIG, mk = 1;
    ma[55] = mj;
    for (int i = 1; i <= 54; i++) {
      int ii = (21 * i) % 55;
      ma[ii] = mk;
      mk = mj - mk;
      if (mk < 0) mk += MBIG;
      mj = ma[ii];
    }
    for (int k = 0; k < 4; k++)
      for (int i = 1; i <= 55; i++) {
        ma[i] -= ma[1 + (i + 30) % 55];
        if (ma[i] < 0) ma[i] += MBIG;
      }
    inext = 0;
    inextp = 31;
    seeded = true;
  }
  // idum < 0 (or first use) re-seeds, then idum := 1
  float next(int &idum) {
    if (idum < 0 || !seeded) {
      seed(idum);
      idum = 1;
    }
    if (++inext == 56) inext = 1;
    if (++inextp == 56) inextp = 1;
    int mj = ma[inext] - ma[inextp];
    if (mj < 0) mj += 1000000000;
    ma[inext] = mj;
    return (float)mj * (1.f / 1000000000.f);
  }
};

// polar Box-Muller pair clipped to [-3,3]: src/random_mod.f90:70-90
static void gasdev1(Ran3 &g, int &idum, float &r1, float &r2) {
  float v1, v2, r;
  do {
    v1 = 2.f * g.next(idum) - 1.f;
    v2 = 2.f * g.next(idum) - 1.f;
    r = v1 * v1 + v2 * v2;
  } while (r >= 1.0f || r == 0.0f);
  // log evaluated in double and rounded once: the convention of the kernels'
  // strict mode and of the oracle
  float fac = sqrtf(-2.f * (float)log((double)r) / r);
  r1 = std::min(3.f, std::max(-3.f, v1 * fac));
  r2 = std::min(3.f, std::max(-3.f, v2 * fac));
}

// ------------------------------------------------------------------ handle --
struct fpb_handle {
  fpb_config cfg;
  std::vector<float> height, xmass;
  std::vector<int32_t> npart;
  DevCfg d;
  int device = 0;
  cudaStream_t stream = nullptr;

  // met: index = Fortran slot - 1
  // FPB_NSLOTS time levels: 1, 2 = the reference's memind slots; 3 = read-ahead slot, allocated on its
  // first upload (numwfmem = 3 of the reference's MPI build with a dedicated reader, src/par_mod.f90:226-227)
  float4 *A[FPB_NSLOTS] = {}, *S[FPB_NSLOTS] = {};
  MetPair *AP[FPB_NSLOTS] = {}; // x-neighbour pairs of A (interp_wind's 256-bit loads), or null
  bool use_pairs = false;
  float *G[FPB_NSLOTS] = {}, *T[FPB_NSLOTS] = {};
  float2 *P[FPB_NSLOTS] = {};
  float4 *R[FPB_NSLOTS] = {};  // wet deposition: {lsprec, convprec, tcc, ctwc}
  int8_t *Cl[FPB_NSLOTS] = {}; // wet deposition: clouds
  float4 *Rn[FPB_MAXNESTS][FPB_NSLOTS] = {};
  int8_t *Cln[FPB_MAXNESTS][FPB_NSLOTS] = {};
  float *Tn[FPB_MAXNESTS][FPB_NSLOTS] = {};
  bool slot_ready[FPB_NSLOTS] = {};
  cudaStream_t st_met = nullptr;   // uploads run here, next to the steps on `stream`
  cudaEvent_t ev_met = nullptr;
  bool met_in_flight = false;
  float met_upload_ms = 0.f;
  cudaEvent_t ev_met0 = nullptr;
  int8_t *stage8 = nullptr;
  size_t stage8_n = 0;
  float *wetgridunc = nullptr, *wetgriduncn = nullptr;
  // nested input grids [nest][slot] (no polar twins, no tt: settling reads the mother grid)
  float4 *An[FPB_MAXNESTS][FPB_NSLOTS] = {}, *Sn[FPB_MAXNESTS][FPB_NSLOTS] = {};
  float *Gn[FPB_MAXNESTS][FPB_NSLOTS] = {}, *tropn[FPB_MAXNESTS][FPB_NSLOTS] = {}, *vdepn[FPB_MAXNESTS][FPB_NSLOTS] = {};
  float *trop[FPB_NSLOTS] = {}, *vdep[FPB_NSLOTS] = {};
  float *stage = nullptr;
  size_t stage_n = 0;
  int memind[2] = {1, 2}, memtime[2] = {0, 0}, lwindinterv = 1;
  bool have_bracket = false;

  DevParticles p{};     // device rows
  DevParticles p_alt{}; // second buffer: sort target / slot-ordered staging
  int32_t *row_of_slot = nullptr;
  bool permuted = false; // rows != slots
  int numpart = 0;
  int active_rows = -1;  // live rows lead the arrays after a sort (-1: unknown)
  int steps_since_sort = 1 << 30;
  unsigned *d_nlive = nullptr;
  int *d_work = nullptr;
  // concoutput on the device
  struct Output {
    float *area[2] = {nullptr, nullptr}, *volume[2] = {nullptr, nullptr}; // [0] mother, [1] nest
    unsigned *block_counts = nullptr;
    int32_t *d_i = nullptr;
    float *d_r = nullptr, *d_density = nullptr;
    int *d_counts = nullptr;
    size_t cap = 0;
    float lon0[2] = {0.f, 0.f}, lat0[2] = {0.f, 0.f};
    bool have_origin = false;
    // partoutput
    float2 *Q[FPB_NSLOTS] = {};  // {pv, qv} per Fortran slot
    float *oro = nullptr;
    bool have_q[FPB_NSLOTS] = {};
    unsigned *po_counts = nullptr;
    int *po_count = nullptr;
    int32_t *po_i[2] = {nullptr, nullptr};
    float *po_f[10] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    float *po_mass = nullptr;
  } outp;
  // the optional hooks of the particle loop: calcfluxes / partpos_average (fpb_output.cuh)
  struct Hooks {
    uint8_t *adv = nullptr;
    float *old = nullptr, *flux = nullptr, *av = nullptr, *init_cond = nullptr;
    int32_t *npart_av = nullptr;
    size_t nflux = 0, ninit = 0;
  } hooks;
  // device-side releaseparticles
  struct Releases {
    int numpoint = 0, itsplit = 0;
    std::vector<int32_t> start, end;
    std::vector<float> xmasssave;
    float *d_pts[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr}; // x1 y1 x2 y2 z1 z2
    int32_t *d_offsets = nullptr;
    float *d_uniforms = nullptr;
    size_t uniforms_cap = 0;
    unsigned *d_block_counts = nullptr;
    int *d_out = nullptr;
    // ran1 of releaseparticles (SAVEd idummy = -7, src/releaseparticles.f90:56)
    int idum = -7, iv[32] = {0}, iy = 0;
    float ran1() { // src/random_mod.f90:40-68 (Numerical Recipes ran1)
      const int IA = 16807, IM = 2147483647, IQ = 127773, IR = 2836, NTAB = 32;
      const int NDIV = 1 + (IM - 1) / NTAB;
      const float AM = 1.f / (float)IM, RNMX = 1.f - 1.2e-7f;
      if (idum <= 0 || iy == 0) {
        idum = (-idum > 1) ? -idum : 1;
        for (int j = NTAB + 8; j >= 1; j--) {
          const int k = idum / IQ;
          idum = IA * (idum - k * IQ) - IR * k;
          if (idum < 0) idum += IM;
          if (j <= NTAB) iv[j - 1] = idum;
        }
        iy = iv[0];
      }
      const int k = idum / IQ;
      idum = IA * (idum - k * IQ) - IR * k;
      if (idum < 0) idum += IM;
      const int j = iy / NDIV;
      iy = iv[j];
      iv[j] = idum;
      const float t = AM * (float)iy;
      return t < RNMX ? t : RNMX;
    }
  } rel;
  // init_domainfill: ran1 with its own SAVEd idummy = -11 (src/init_domainfill.f90:47)
  struct Domainfill {
    bool done = false;
    int gdomainfill = 0;
    int nx_we[2] = {0, 0}, ny_sn[2] = {0, 0};
    int idum = -11, iv[32] = {0}, iy = 0;
    int idum_bc = -11;        // boundcond_domainfill's own SAVEd idummy (src/boundcond_domainfill.f90:49)
    int itsplit = 0, numparticlecount = 0;
    float xmassperparticle = 0.f;
    // inflow boundary of a limited box (fpb_boundcond_domainfill)
    int nloc = 0;
    std::vector<uint8_t> loc_draws; // ran1 draws per particle of each location (2 or 3)
    BcLoc *d_loc = nullptr;
    float *d_acc = nullptr;
    int32_t *d_mmass = nullptr, *d_first = nullptr, *d_uoff = nullptr;
    unsigned *d_blocks = nullptr;
    int *d_out = nullptr;
    float *d_uniforms = nullptr;
    size_t uniforms_cap = 0;
    float ran1(int &idum) { // src/random_mod.f90:40-68 (iv, iy are SAVEd in ran1, idum is the caller's)
      const int IA = 16807, IM = 2147483647, IQ = 127773, IR = 2836, NTAB = 32;
      const int NDIV = 1 + (IM - 1) / NTAB;
      const float AM = 1.f / (float)IM, RNMX = 1.f - 1.2e-7f;
      if (idum <= 0 || iy == 0) {
        idum = (-idum > 1) ? -idum : 1;
        for (int j = NTAB + 8; j >= 1; j--) {
          const int k = idum / IQ;
          idum = IA * (idum - k * IQ) - IR * k;
          if (idum < 0) idum += IM;
          if (j <= NTAB) iv[j - 1] = idum;
        }
        iy = iv[0];
      }
      const int k = idum / IQ;
      idum = IA * (idum - k * IQ) - IR * k;
      if (idum < 0) idum += IM;
      const int j = iy / NDIV;
      iy = iv[j];
      iv[j] = idum;
      const float t = AM * (float)iy;
      return t < RNMX ? t : RNMX;
    }
    float ran1() { return ran1(idum); }
  } dfill;
  // grid exchange over NCCL (fpb_comm_init / fpb_reduce_grids_begin / _end)
  struct Comm {
    void *nccl = nullptr;       // ncclComm_t
    int rank = 0, nranks = 1;
    cudaStream_t side = nullptr;
    cudaEvent_t ev_staged = nullptr, ev_t0 = nullptr, ev_done = nullptr;
    float *stage[7] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    size_t stage_n[7] = {0, 0, 0, 0, 0, 0, 0};
    bool in_flight = false, timed = false;
    // The NCCL enqueue (group start .. end, tens of microseconds of host time) runs on a worker
    // thread: every engine call is synchronous for the host, so host time between two calls is idle
    // GPU time; fpb_reduce_grids_begin only posts the job.
    std::thread worker;
    std::mutex mu;
    std::condition_variable cv;
    bool job = false, job_done = true, quit = false;
    int job_rc = 0;
    std::string job_err;
  } comm;
  // convective mixing (fpb_set_convection / fpb_upload_convmet / fpb_convmix)
  struct Conv {
    int nuvz = 0, nuvzmax = 0, nconvlev = 0;
    float *d_ab = nullptr;                 // akz, bkz, akm, bkm: 4 x (nuvz + 1), 1-based
    float2 *CT[FPB_MAXNESTS + 1][FPB_NSLOTS] = {};   // [0]: mother grid, [l]: nested input grid l
    float4 *CS[FPB_MAXNESTS + 1][FPB_NSLOTS] = {};
    bool have[FPB_MAXNESTS + 1][FPB_NSLOTS] = {};
    float *cbaseflux[FPB_MAXNESTS + 1] = {}, *cbase_bak[FPB_MAXNESTS + 1] = {};
    float *pool = nullptr, *pool2 = nullptr;
    uint8_t *col_state = nullptr;
    int pool_cols = 0;
    ScatterWork sw;
    unsigned *block_counts = nullptr, *col_key = nullptr;
    int32_t *colidx = nullptr, *col_start = nullptr, *col_lconv = nullptr, *key_by_slot = nullptr;
    int *d_total = nullptr;
    uint8_t *draws = nullptr;
    float *rn_by_slot = nullptr;
    size_t cap_rows = 0, cap_cols = 0;
    int iseed = -88;                       // SAVEd iseed of redist, src/redist.f90:58
  } conv;
  // calcpar + verttransform on the device (fpb_set_vertical / fpb_calcpar_verttransform)
  struct MetProc {
    int nuvz = 0, nwz = 0, nuvzmax = 0, nwzmax = 0;
    float *d_ab = nullptr;   // akz, bkz, akm, bkm: 4 x (nuvz + 1), 1-based
    float *d_cosf = nullptr; // [ny]
    float2 *UV = nullptr;
    float *W = nullptr, *PV = nullptr, *theta = nullptr, *excessoro = nullptr, *uvzlev = nullptr;
    float *CLW = nullptr, *CIW = nullptr, *clw = nullptr; // readclouds
    struct NestWork { // calcpar_nests / verttransform_nests: work arrays in the nest's extents
      float2 *UV = nullptr, *Q = nullptr;
      float *W = nullptr, *PV = nullptr, *theta = nullptr, *excessoro = nullptr, *uvzlev = nullptr, *T = nullptr, *cosf = nullptr;
      float4 *SF2 = nullptr;
    } nw[FPB_MAXNESTS];
    float4 *SF2 = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, evk = nullptr;
  } metproc;
  DevScratch sc{}; // fpb_pbl_kernel -> fpb_finish_kernel hand-over rows
  std::vector<int32_t> h_slot;
  DevCfg d_tmp;

  float *d_height = nullptr, *d_xmass = nullptr;
  int32_t *d_npart = nullptr;

  float *d_rannumb = nullptr;
  int maxrand = 0;
  Ran3 ran3;
  int idummy_init = -7, idummy_adv = -7; // SAVEd locals, src/advance.f90:120, src/initialize.f90:64
  int32_t *d_nrand_init = nullptr, *d_nrand_adv = nullptr;
  std::vector<int32_t> h_itra1, h_itramem, h_nrand_init, h_nrand_adv;

  float *gridunc = nullptr, *griduncn = nullptr, *drygridunc = nullptr, *drygriduncn = nullptr;
  float *creceptor = nullptr, *crec_acc = nullptr;
  size_t n_grid = 0, n_gridn = 0, n_dry = 0, n_dryn = 0, n_rec = 0;

  unsigned long long *d_stats = nullptr;
  int64_t launches = 0;
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr}; // step begin/end, conccalc begin/end
  bool timed_step = false, timed_conc = false;
  bool pending_init = true;
  bool have_init_until = false; // fpb_push_particles: latest itramem pushed (times ldirect)
  long long init_until = 0;
  ScatterWork scatter;
  void *sort_rec = nullptr; // packed records of the cell sort (sortk_permute_packed)
  size_t sort_rec_cap = 0;
  // deterministic deposition / receptor records (FPB_SCATTER_DETERMINISTIC), grown on demand
  struct DepStore {
    unsigned *keys[2] = {nullptr, nullptr};
    float *vals[2] = {nullptr, nullptr};
    size_t cap = 0;
    float *rec = nullptr; // receptor contributions [numreceptor*nspec][nslots]
    size_t cap_rec = 0;
  } depstore;

  // fpb_step_host pipeline lanes: chunk c runs on lane c % NLANES (own stream,
  // sort work area and work counter), so the copies of one chunk overlap the
  // kernels of another
  struct Lane {
    cudaStream_t st = nullptr;
    ScatterWork sw;
    DepStore dep;
    int *d_work = nullptr;
    unsigned *d_nlive = nullptr;
  };
  static constexpr int NLANES = 3;
  Lane lanes[NLANES];
  cudaStream_t st_in = nullptr;              // all host-to-device copies of fpb_step_host, in chunk order
  cudaEvent_t ev_in[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t ev_det[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t ev_dep[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t ev_ready = nullptr;
  bool lanes_ready = false;
  static constexpr int MAXCHUNKS = 32;
  cudaStream_t st_out = nullptr;      // FPB_HOST_DEFER_D2H: the copy-out stream
  cudaEvent_t ev_out[MAXCHUNKS] = {};
};

// ---- deterministic deposition records -----------------------------------------------------------
static void dep_free(fpb_handle::DepStore &s) {
  for (int g = 0; g < 2; g++) { cudaFree(s.keys[g]); cudaFree(s.vals[g]); s.keys[g] = nullptr; s.vals[g] = nullptr; }
  cudaFree(s.rec);
  s.rec = nullptr;
  s.cap = s.cap_rec = 0;
}
// record area for nslots particle slots starting at slot_base; keys cleared to "no record"
static int dep_begin(fpb_handle *h, fpb_handle::DepStore &s, int nslots, int slot_base, cudaStream_t st,
                     DevDepRecords &out) {
  const size_t nrec = 4 * (size_t)nslots;
  const int ngrids = h->cfg.nested_output == 1 ? 2 : 1;
  if (nrec > s.cap) {
    CK(cudaStreamSynchronize(st));
    for (int g = 0; g < 2; g++) { cudaFree(s.keys[g]); cudaFree(s.vals[g]); s.keys[g] = nullptr; s.vals[g] = nullptr; }
    for (int g = 0; g < ngrids; g++) {
      CK(cudaMalloc((void **)&s.keys[g], nrec * sizeof(unsigned)));
      CK(cudaMalloc((void **)&s.vals[g], nrec * h->cfg.nspec * sizeof(float)));
    }
    s.cap = nrec;
  }
  for (int g = 0; g < 2; g++) {
    out.keys[g] = g < ngrids ? s.keys[g] : nullptr;
    out.vals[g] = g < ngrids ? s.vals[g] : nullptr;
    if (g < ngrids) CK(cudaMemsetAsync(s.keys[g], 0xff, nrec * sizeof(unsigned), st));
  }
  out.nrec = nrec;
  out.slot_base = slot_base;
  return 0;
}
// add the records to the (dry or wet) deposition grids, every cell in slot order
static int dep_apply(fpb_handle *h, ScatterWork &sw, const DevDepRecords &r, float *grid, float *gridn, cudaStream_t st) {
  const fpb_config &c = h->cfg;
  const unsigned long long per = (unsigned long long)c.maxpointspec_act * c.nclassunc * c.maxageclass;
  if (scatter_records_deterministic(sw, r.keys[0], r.vals[0], r.nrec, c.nspec, grid, c.numxgrid * c.numygrid,
                                    per * c.numxgrid * c.numygrid, st, &h->launches))
    return fail("%s", scatter_error());
  if (c.nested_output == 1 &&
      scatter_records_deterministic(sw, r.keys[1], r.vals[1], r.nrec, c.nspec, gridn, c.numxgridn * c.numygridn,
                                    per * c.numxgridn * c.numygridn, st, &h->launches))
    return fail("%s", scatter_error());
  return 0;
}
// receptor contribution area [numreceptor*nspec][nslots], zeroed
static int rec_begin(fpb_handle *h, fpb_handle::DepStore &s, int nslots, cudaStream_t st, DevConcArgs &a) {
  const size_t n = (size_t)h->cfg.numreceptor * h->cfg.nspec * nslots;
  if (n > s.cap_rec) {
    CK(cudaStreamSynchronize(st));
    cudaFree(s.rec);
    s.rec = nullptr;
    CK(cudaMalloc((void **)&s.rec, n * sizeof(float)));
    s.cap_rec = n;
  }
  CK(cudaMemsetAsync(s.rec, 0, n * sizeof(float), st));
  a.rec_vals = s.rec;
  a.rec_nslots = nslots;
  return 0;
}

static void fill_devcfg(fpb_handle *h) {
  const fpb_config &c = h->cfg;
  DevCfg &d = h->d;
  memset(&d, 0, sizeof d);
  d.nx = c.nx; d.ny = c.ny; d.nz = c.nz;
  d.nxd = c.nx;
  d.nyd = (c.ny < c.nymax) ? c.ny + 1 : c.ny; // keep the row the reference's jyp can touch
  d.nymax = c.nymax;
  d.nxmin1 = c.nxmin1; d.nymin1 = c.nymin1;
  d.dx = c.dx; d.dy = c.dy; d.xlon0 = c.xlon0; d.ylat0 = c.ylat0;
  d.dxconst = c.dxconst; d.dyconst = c.dyconst;
  d.xglobal = c.xglobal; d.nglobal = c.nglobal; d.sglobal = c.sglobal;
  d.switchnorthg = c.switchnorthg; d.switchsouthg = c.switchsouthg;
  memcpy(d.northpolemap, c.northpolemap, sizeof d.northpolemap);
  memcpy(d.southpolemap, c.southpolemap, sizeof d.southpolemap);
  d.eps = c.eps;
  d.numbnests = c.numbnests;
  for (int l = 0; l < FPB_MAXNESTS; l++) {
    d.nxdn[l] = c.nxn[l]; d.nydn[l] = c.nyn[l];
    d.xln[l] = c.xln[l]; d.yln[l] = c.yln[l]; d.xrn[l] = c.xrn[l]; d.yrn[l] = c.yrn[l];
    d.xresoln[l] = c.xresoln[l]; d.yresoln[l] = c.yresoln[l];
  }
  d.ldirect = c.ldirect; d.lsynctime = c.lsynctime; d.method = c.method;
  d.mintime = c.mintime; d.ifine = c.ifine;
  d.turbswitch = c.turbswitch; d.cblflag = c.cblflag; d.mdomainfill = c.mdomainfill;
  d.mquasilag = c.mquasilag; d.lsettling = c.lsettling; d.turboff = c.turboff;
  d.ctl = c.ctl; d.fine = c.fine; d.d_trop = c.d_trop; d.d_strat = c.d_strat;
  d.turbmesoscale = c.turbmesoscale;
  d.ind_samp = c.ind_samp; d.ioutputforeachrelease = c.ioutputforeachrelease;
  d.lusekerneloutput = c.lusekerneloutput; d.lparticlecountoutput = c.lparticlecountoutput;
  d.drydep = c.drydep; d.drybkdep = c.drybkdep; d.wetbkdep = c.wetbkdep;
  d.nested_output = c.nested_output;
  d.linit_cond = c.linit_cond;
  d.nspec = c.nspec;
  for (int k = 0; k < FPB_MAXSPEC; k++) {
    d.decay[k] = c.decay[k]; d.drydepspec[k] = c.drydepspec[k]; d.density[k] = c.density[k];
    d.dquer[k] = c.dquer[k]; d.vsetaver[k] = c.vsetaver[k]; d.cunningham[k] = c.cunningham[k];
  }
  for (int k = 0; k < FPB_MAXSPEC; k++) {
    d.wetdepspec[k] = c.wetdepspec[k]; d.weta_gas[k] = c.weta_gas[k]; d.wetb_gas[k] = c.wetb_gas[k];
    d.crain_aero[k] = c.crain_aero[k]; d.csnow_aero[k] = c.csnow_aero[k]; d.ccn_aero[k] = c.ccn_aero[k];
    d.in_aero[k] = c.in_aero[k]; d.henry[k] = c.henry[k];
  }
  d.readclouds = c.readclouds;
  for (int l = 0; l < FPB_MAXNESTS; l++) d.readclouds_nest[l] = c.readclouds_nest[l];
  d.nageclass = c.nageclass;
  for (int k = 0; k < FPB_MAXAGECLASS; k++) d.lage[k] = c.lage[k];
  d.numxgrid = c.numxgrid; d.numygrid = c.numygrid; d.numzgrid = c.numzgrid;
  d.dxout = c.dxout; d.dyout = c.dyout; d.xoutshift = c.xoutshift; d.youtshift = c.youtshift;
  for (int k = 0; k < FPB_MAXZGRID; k++) d.outheight[k] = c.outheight[k];
  d.numxgridn = c.numxgridn; d.numygridn = c.numygridn;
  d.dxoutn = c.dxoutn; d.dyoutn = c.dyoutn; d.xoutshiftn = c.xoutshiftn; d.youtshiftn = c.youtshiftn;
  d.maxpointspec_act = c.maxpointspec_act; d.nclassunc = c.nclassunc; d.maxageclass = c.maxageclass;
  d.numreceptor = c.numreceptor;
  for (int k = 0; k < FPB_MAXRECEPTOR; k++) {
    d.xreceptor[k] = c.xreceptor[k]; d.yreceptor[k] = c.yreceptor[k]; d.receptorarea[k] = c.receptorarea[k];
  }
  d.numpoint = c.numpoint;
  d.rng_mode = c.rng_mode;
  d.seed = c.seed;
  d.part_id_stride = c.part_id_stride ? c.part_id_stride : 1;
  d.part_id_offset = c.part_id_offset;
}

// -------------------------------------------------------------- utilities --
extern "C" const char *fpb_last_error(void) { return g_err.c_str(); }
extern "C" int fpb_abi_version(void) { return FPB_ABI_VERSION; }
extern "C" size_t fpb_config_sizeof(void) { return sizeof(fpb_config); }

template <typename T>
static int dalloc(T **p, size_t n) {
  *p = nullptr;
  if (n == 0) n = 1;
  CK(cudaMalloc((void **)p, n * sizeof(T)));
  CK(cudaMemset(*p, 0, n * sizeof(T)));
  // the memset runs on the legacy default stream, which the engine's non-blocking streams do not
  // order against: without this a later cudaMemcpyAsync into the new buffer can be overtaken by it
  CK(cudaStreamSynchronize(cudaStreamLegacy));
  return 0;
}
#define DA(ptr, n)                      \
  do {                                  \
    if (dalloc(&(ptr), (n))) return 1;  \
  } while (0)

__global__ void fill_i32_kernel(int32_t *p, int32_t v, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

// Fortran-layout host arrays -> device layout.  Up to four arrays that end up interleaved in one
// device word ({uu,vv,ww,rho}, {hmix,ustar,wstar,oli}, ...) are copied into the staging buffer back
// to back (one H2D copy each, no host synchronisation in between) and ONE kernel then writes whole
// words:  dst[(k*nyd + jy)*nxd + ix] = {src_c[(k*nymax + jy)*nxmax + ix], c = 0..NC-1}.
template <int NC>
__global__ void __launch_bounds__(256)
pack_group_kernel(float *dst, const float *stage, size_t comp_stride, int nxd, int nyd, int nk, int nxmax, int nymax) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t n = (size_t)nxd * nyd * nk;
  if (i >= n) return;
  const int ix = (int)(i % nxd);
  const size_t r = i / nxd;
  const int jy = (int)(r % nyd), k = (int)(r / nyd);
  const size_t o = ((size_t)k * nymax + jy) * nxmax + ix;
  float v[NC];
#pragma unroll
  for (int c = 0; c < NC; c++) v[c] = (jy < nymax) ? stage[c * comp_stride + o] : 0.f;
  if (NC == 4) reinterpret_cast<float4 *>(dst)[i] = make_float4(v[0], v[1 % NC], v[2 % NC], v[3 % NC]);
  else if (NC == 2) reinterpret_cast<float2 *>(dst)[i] = make_float2(v[0], v[1 % NC]);
  else dst[i] = v[0];
}

__global__ void pack_i8_kernel(int8_t *dst, const int8_t *src, int nxd, int nyd, int nk, int nxmax, int nymax) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t n = (size_t)nxd * nyd * nk;
  if (i >= n) return;
  int ix = (int)(i % nxd);
  size_t r = i / nxd;
  int jy = (int)(r % nyd), k = (int)(r / nyd);
  dst[i] = (jy < nymax) ? src[((size_t)k * nymax + jy) * nxmax + ix] : (int8_t)0;
}

// stream-ordered, no host synchronisation (pageable sources make cudaMemcpyAsync block by themselves)
static int upload_i8(fpb_handle *h, cudaStream_t st, int8_t *dst, const int8_t *src, int nk, int nxd, int nyd,
                     int nxmax, int nymax) {
  const size_t nsrc = (size_t)nxmax * nymax * nk, n = (size_t)nxd * nyd * nk;
  if (nsrc > h->stage8_n) {
    CK(cudaStreamSynchronize(st));
    if (h->stage8) cudaFree(h->stage8);
    h->stage8 = nullptr;
    CK(cudaMalloc((void **)&h->stage8, nsrc));
    h->stage8_n = nsrc;
  }
  CK(cudaMemcpyAsync(h->stage8, src, nsrc, cudaMemcpyHostToDevice, st));
  pack_i8_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(dst, h->stage8, nxd, nyd, nk, nxmax, nymax);
  h->launches++;
  CK(cudaGetLastError());
  return 0;
}

static int upload_group(fpb_handle *h, cudaStream_t st, float *dst, int ncomp, const float *const *src, int nk,
                        int nxd, int nyd, int nxmax, int nymax) {
  const size_t nsrc = (size_t)nxmax * nymax * nk;
  if (nsrc * ncomp > h->stage_n) {
    CK(cudaStreamSynchronize(st));
    if (h->stage) cudaFree(h->stage);
    h->stage = nullptr;
    CK(cudaMalloc((void **)&h->stage, nsrc * ncomp * sizeof(float)));
    h->stage_n = nsrc * ncomp;
  }
  for (int c = 0; c < ncomp; c++) {
    if (src[c]) CK(cudaMemcpyAsync(h->stage + c * nsrc, src[c], nsrc * sizeof(float), cudaMemcpyHostToDevice, st));
    else CK(cudaMemsetAsync(h->stage + c * nsrc, 0, nsrc * sizeof(float), st));
  }
  const size_t n = (size_t)nxd * nyd * nk;
  const unsigned nb = (unsigned)((n + 255) / 256);
  if (ncomp == 4) pack_group_kernel<4><<<nb, 256, 0, st>>>(dst, h->stage, nsrc, nxd, nyd, nk, nxmax, nymax);
  else if (ncomp == 2) pack_group_kernel<2><<<nb, 256, 0, st>>>(dst, h->stage, nsrc, nxd, nyd, nk, nxmax, nymax);
  else if (ncomp == 1) pack_group_kernel<1><<<nb, 256, 0, st>>>(dst, h->stage, nsrc, nxd, nyd, nk, nxmax, nymax);
  else return fail("upload_group: ncomp = %d", ncomp);
  h->launches++;
  CK(cudaGetLastError());
  return 0;
}
static int upload_group(fpb_handle *h, cudaStream_t st, float *dst, int ncomp, const float *const *src, int nk) {
  return upload_group(h, st, dst, ncomp, src, nk, h->d.nxd, h->d.nyd, h->cfg.nxmax, h->cfg.nymax);
}

__global__ void __launch_bounds__(256) pair_kernel(const float4 *A, MetPair *AP, int nxd, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  MetPair q;
  q.a = A[i];
  q.b = ((int)(i % (size_t)nxd) < nxd - 1) ? A[i + 1] : q.a;
  AP[i] = q;
}
// after A of slot s has been written on stream st
static int build_pairs(fpb_handle *h, int s, cudaStream_t st) {
  if (!h->use_pairs) return 0;
  const size_t n3 = (size_t)h->d.nxd * h->d.nyd * h->cfg.nz;
  if (!h->AP[s] && cudaMalloc((void **)&h->AP[s], n3 * sizeof(MetPair)) != cudaSuccess) {
    cudaGetLastError();
    h->use_pairs = false; // (no room: the plain loads)
    for (auto &q : h->AP) { cudaFree(q); q = nullptr; }
    return 0;
  }
  pair_kernel<<<(unsigned)((n3 + 255) / 256), 256, 0, st>>>(h->A[s], h->AP[s], h->d.nxd, n3);
  h->launches++;
  CK(cudaGetLastError());
  return 0;
}

// device arrays of one time level (slot index s = Fortran slot - 1)
static int alloc_met_slot(fpb_handle *h, int s) {
  if (h->A[s]) return 0;
  const DevCfg &d = h->d;
  const fpb_config &c = h->cfg;
  const size_t n3 = (size_t)d.nxd * d.nyd * c.nz, n2 = (size_t)d.nxd * d.nyd;
  DA(h->A[s], n3); DA(h->G[s], n3); DA(h->T[s], n3); DA(h->P[s], n3); DA(h->S[s], n2);
  DA(h->trop[s], n2); DA(h->vdep[s], n2 * c.nspec);
  if (c.wetdep) { DA(h->R[s], n2); DA(h->Cl[s], n3); }
  for (int l = 0; l < c.numbnests; l++) {
    const size_t m2 = (size_t)c.nxn[l] * c.nyn[l], m3 = m2 * c.nz;
    DA(h->An[l][s], m3); DA(h->Gn[l][s], m3); DA(h->Sn[l][s], m2);
    DA(h->tropn[l][s], m2); DA(h->vdepn[l][s], m2 * c.nspec);
    if (c.wetdep) { DA(h->Rn[l][s], m2); DA(h->Cln[l][s], m3); DA(h->Tn[l][s], m3); }
  }
  return 0;
}

// ------------------------------------------------------------------- init --
extern "C" int fpb_init(const fpb_config *cfg, fpb_handle **out) {
  if (!cfg || !out) return fail("fpb_init: null argument");
  *out = nullptr;
  if (cfg->abi_version != FPB_ABI_VERSION)
    return fail("fpb_init: abi_version %d != %d", cfg->abi_version, FPB_ABI_VERSION);
  if (cfg->nz > FPB_MAXNZ || cfg->nz < 2) return fail("fpb_init: nz=%d out of range (max %d)", cfg->nz, FPB_MAXNZ);
  if (cfg->nspec < 1 || cfg->nspec > FPB_MAXSPEC) return fail("fpb_init: nspec=%d out of range", cfg->nspec);
  if (cfg->nageclass < 1 || cfg->nageclass > FPB_MAXAGECLASS) return fail("fpb_init: nageclass out of range");
  if (cfg->numzgrid < 1 || cfg->numzgrid > FPB_MAXZGRID) return fail("fpb_init: numzgrid out of range");
  if (cfg->numreceptor < 0 || cfg->numreceptor > FPB_MAXRECEPTOR) return fail("fpb_init: numreceptor out of range");
  if (cfg->numbnests < 0 || cfg->numbnests > FPB_MAXNESTS) return fail("fpb_init: numbnests=%d out of range (max %d)", cfg->numbnests, FPB_MAXNESTS);
  for (int l = 0; l < cfg->numbnests; l++)
    if (cfg->nxn[l] < 2 || cfg->nyn[l] < 2 || cfg->nxn[l] > cfg->nxmaxn || cfg->nyn[l] > cfg->nymaxn)
      return fail("fpb_init: nest %d extents %dx%d outside (2..nxmaxn=%d, 2..nymaxn=%d)", l + 1, cfg->nxn[l], cfg->nyn[l], cfg->nxmaxn, cfg->nymaxn);
  if (!cfg->height) return fail("fpb_init: height is null");
  if (cfg->maxpart < 1) return fail("fpb_init: maxpart < 1");
  if (cfg->numpoint < 1 || !cfg->npart || !cfg->xmass) return fail("fpb_init: releases (numpoint/npart/xmass) missing");
  if (cfg->ldirect != 1 && cfg->ldirect != -1) return fail("fpb_init: ldirect must be 1 or -1 (got %d)", cfg->ldirect);
  if ((cfg->drybkdep || cfg->wetbkdep) && cfg->ldirect != -1)
    return fail("fpb_init: drybkdep/wetbkdep (IND_RECEPTOR 3/4) are backward-run options (src/readcommand.f90:320-339)");
  if (cfg->linit_cond < 0 || cfg->linit_cond > 2) return fail("fpb_init: linit_cond must be 0, 1 or 2 (got %d)", cfg->linit_cond);
  if (cfg->linit_cond > 0 && cfg->ldirect == 1) return fail("fpb_init: linit_cond is a backward-run option (src/readcommand.f90:349)");
  if (cfg->drybkdep && !cfg->drydep) return fail("fpb_init: drybkdep needs drydep (a species with dry deposition)");
  if (cfg->wetbkdep && !cfg->wetdep) return fail("fpb_init: wetbkdep needs wetdep (a species with wet deposition)");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail("fpb_init: no CUDA device (%s); this library has no CPU path",
                e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
  if (cfg->device < 0 || cfg->device >= ndev) return fail("fpb_init: device %d not present", cfg->device);

  fpb_handle *h = new fpb_handle();
  h->cfg = *cfg;
  h->device = cfg->device;
  h->height.assign(cfg->height, cfg->height + cfg->nz);
  h->npart.assign(cfg->npart, cfg->npart + cfg->numpoint);
  // device xmass is packed [nspec][numpoint]
  h->xmass.resize((size_t)cfg->nspec * cfg->numpoint);
  for (int k = 0; k < cfg->nspec; k++)
    for (int i = 0; i < cfg->numpoint; i++)
      h->xmass[(size_t)k * cfg->numpoint + i] = cfg->xmass[i + (size_t)cfg->numpoint * k];
  h->cfg.height = nullptr; h->cfg.npart = nullptr; h->cfg.xmass = nullptr;
  fill_devcfg(h);
  if (const char *e = getenv("FPB_MET_PAIRS")) h->use_pairs = atoi(e) != 0; // (experiment knob)
  const DevCfg &d = h->d;
  const fpb_config &c = h->cfg;

  CK(cudaSetDevice(h->device));
  CK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  for (int k = 0; k < 4; k++) CK(cudaEventCreate(&h->ev[k]));

  CK(cudaStreamCreateWithFlags(&h->st_met, cudaStreamNonBlocking));
  CK(cudaEventCreate(&h->ev_met));
  CK(cudaEventCreate(&h->ev_met0));
  for (int s = 0; s < 2; s++)
    if (alloc_met_slot(h, s)) return 1;
  const size_t mp = (size_t)c.maxpart;
  for (DevParticles *q : {&h->p, &h->p_alt}) {
    DA(q->xtra1, mp); DA(q->ytra1, mp); DA(q->ztra1, mp);
    DA(q->itra1, mp); DA(q->npoint, mp); DA(q->nclass, mp); DA(q->idt, mp);
    DA(q->itramem, mp); DA(q->itrasplit, mp);
    DA(q->uap, mp); DA(q->ucp, mp); DA(q->uzp, mp);
    DA(q->us, mp); DA(q->vs, mp); DA(q->ws, mp); DA(q->cbt, mp);
    DA(q->xmass1, mp * c.nspec);
    DA(q->xscav_frac1, mp * c.nspec);
    DA(q->slot, mp);
    q->maxpart = c.maxpart;
    sortk_iota(q->slot, c.maxpart, h->stream);
  }
  DA(h->row_of_slot, mp);
  sortk_iota(h->row_of_slot, c.maxpart, h->stream);
  DA(h->d_nlive, 1);
  DA(h->d_work, 1);
  DA(h->sc.flags, mp); DA(h->sc.s0, mp); DA(h->sc.s1, mp); DA(h->sc.s2, mp);
  if (c.drydep) DA(h->sc.prob, mp * c.nspec);
  if (c.iflux == 1 || c.ipout == 3 || c.linit_cond > 0) DA(h->hooks.adv, mp);
  if (c.iflux == 1 || c.linit_cond > 0) DA(h->hooks.old, mp * (3 + (size_t)c.nspec));
  if (c.iflux == 1) {
    h->hooks.nflux = (size_t)6 * c.numxgrid * c.numygrid * c.numzgrid * c.nspec * c.maxpointspec_act * c.nageclass;
    DA(h->hooks.flux, h->hooks.nflux);
  }
  if (c.linit_cond > 0) {
    h->hooks.ninit = (size_t)c.numxgrid * c.numygrid * c.numzgrid * c.maxspec * c.maxpointspec_act;
    DA(h->hooks.init_cond, h->hooks.ninit);
  }
  if (c.ipout == 3) { DA(h->hooks.npart_av, mp); DA(h->hooks.av, mp * 14); }
  h->launches += 3;
  // itra1(:) = -999999999, src/FLEXPART.f90:315-317
  fill_i32_kernel<<<(unsigned)((mp + 255) / 256), 256, 0, h->stream>>>(h->p.itra1, FPB_ITRA_DEAD, c.maxpart);
  h->launches++;

  DA(h->d_height, c.nz); DA(h->d_npart, c.numpoint); DA(h->d_xmass, h->xmass.size());
  CK(cudaMemcpy(h->d_height, h->height.data(), c.nz * sizeof(float), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(h->d_npart, h->npart.data(), c.numpoint * sizeof(int32_t), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(h->d_xmass, h->xmass.data(), h->xmass.size() * sizeof(float), cudaMemcpyHostToDevice));
  DA(h->d_nrand_init, mp); DA(h->d_nrand_adv, mp);

  const size_t outer = (size_t)c.nspec * c.maxpointspec_act * c.nclassunc * c.maxageclass;
  h->n_grid = (size_t)c.numxgrid * c.numygrid * c.numzgrid * outer;
  h->n_dry = (size_t)c.numxgrid * c.numygrid * outer;
  DA(h->gridunc, h->n_grid); DA(h->drygridunc, h->n_dry);
  if (c.nested_output == 1) {
    h->n_gridn = (size_t)c.numxgridn * c.numygridn * c.numzgrid * outer;
    h->n_dryn = (size_t)c.numxgridn * c.numygridn * outer;
    DA(h->griduncn, h->n_gridn); DA(h->drygriduncn, h->n_dryn);
  }
  h->n_rec = (size_t)FPB_MAXRECEPTOR * c.nspec;
  if (c.wetdep) {
    DA(h->wetgridunc, h->n_dry);
    if (c.nested_output == 1) DA(h->wetgriduncn, h->n_dryn);
  }
  DA(h->creceptor, h->n_rec); DA(h->crec_acc, h->n_rec);
  DA(h->d_stats, 8);
  CK(cudaStreamSynchronize(h->stream));
  *out = h;
  return 0;
}

extern "C" int fpb_comm_finalize(fpb_handle *h);
extern "C" int fpb_finalize(fpb_handle *h) {
  if (!h) return 0;
  cudaSetDevice(h->device);
  cudaStreamSynchronize(h->stream);
  if (h->st_met) { cudaStreamSynchronize(h->st_met); cudaStreamDestroy(h->st_met); }
  if (h->ev_met) cudaEventDestroy(h->ev_met);
  if (h->ev_met0) cudaEventDestroy(h->ev_met0);
  for (int s = 0; s < FPB_NSLOTS; s++) {
    cudaFree(h->AP[s]);
    cudaFree(h->A[s]); cudaFree(h->G[s]); cudaFree(h->T[s]); cudaFree(h->P[s]); cudaFree(h->S[s]); cudaFree(h->trop[s]); cudaFree(h->vdep[s]);
  }
  for (int l = 0; l < FPB_MAXNESTS; l++)
    for (int s = 0; s < FPB_NSLOTS; s++) {
      cudaFree(h->Rn[l][s]); cudaFree(h->Cln[l][s]); cudaFree(h->Tn[l][s]);
      cudaFree(h->An[l][s]); cudaFree(h->Gn[l][s]); cudaFree(h->Sn[l][s]); cudaFree(h->tropn[l][s]); cudaFree(h->vdepn[l][s]);
    }
  cudaFree(h->stage); cudaFree(h->stage8); cudaFree(h->wetgridunc); cudaFree(h->wetgriduncn);
  for (int s = 0; s < FPB_NSLOTS; s++) { cudaFree(h->R[s]); cudaFree(h->Cl[s]); }
  for (DevParticles *q : {&h->p, &h->p_alt}) {
    cudaFree(q->xtra1); cudaFree(q->ytra1); cudaFree(q->ztra1); cudaFree(q->itra1);
    cudaFree(q->npoint); cudaFree(q->nclass); cudaFree(q->idt); cudaFree(q->itramem);
    cudaFree(q->itrasplit); cudaFree(q->uap); cudaFree(q->ucp); cudaFree(q->uzp);
    cudaFree(q->us); cudaFree(q->vs); cudaFree(q->ws); cudaFree(q->cbt);
    cudaFree(q->xmass1); cudaFree(q->xscav_frac1); cudaFree(q->slot);
  }
  cudaFree(h->row_of_slot); cudaFree(h->d_nlive); cudaFree(h->d_work);
  cudaFree(h->sc.flags); cudaFree(h->sc.s0); cudaFree(h->sc.s1); cudaFree(h->sc.s2); cudaFree(h->sc.prob);
  cudaFree(h->d_height); cudaFree(h->d_npart); cudaFree(h->d_xmass);
  for (int k = 0; k < 2; k++) { cudaFree(h->outp.area[k]); cudaFree(h->outp.volume[k]); }
  cudaFree(h->outp.block_counts); cudaFree(h->outp.d_i); cudaFree(h->outp.d_r); cudaFree(h->outp.d_counts);
  cudaFree(h->outp.d_density);
  for (auto &q : h->outp.Q) cudaFree(q);
  cudaFree(h->outp.oro); cudaFree(h->outp.po_counts);
  cudaFree(h->outp.po_count); cudaFree(h->outp.po_i[0]); cudaFree(h->outp.po_i[1]); cudaFree(h->outp.po_mass);
  for (auto &q : h->outp.po_f) cudaFree(q);
  for (auto &q : h->rel.d_pts) cudaFree(q);
  cudaFree(h->rel.d_offsets); cudaFree(h->rel.d_uniforms); cudaFree(h->rel.d_block_counts); cudaFree(h->rel.d_out);
  {
    auto &V = h->conv;
    cudaFree(V.d_ab); cudaFree(V.pool); cudaFree(V.block_counts);
    for (auto &q : V.cbaseflux) cudaFree(q);
    for (auto &q : V.cbase_bak) cudaFree(q);
    cudaFree(V.col_key); cudaFree(V.colidx); cudaFree(V.col_start); cudaFree(V.col_lconv); cudaFree(V.key_by_slot);
    cudaFree(V.pool2); cudaFree(V.col_state);
    cudaFree(V.d_total); cudaFree(V.draws); cudaFree(V.rn_by_slot);
    for (auto &g : V.CT) for (auto &q : g) cudaFree(q);
    for (auto &g : V.CS) for (auto &q : g) cudaFree(q);
    scatter_free(V.sw);
  }
  cudaFree(h->d_rannumb); cudaFree(h->d_nrand_init); cudaFree(h->d_nrand_adv);
  cudaFree(h->gridunc); cudaFree(h->griduncn); cudaFree(h->drygridunc); cudaFree(h->drygriduncn);
  cudaFree(h->creceptor); cudaFree(h->crec_acc); cudaFree(h->d_stats);
  scatter_free(h->scatter);
  cudaFree(h->sort_rec);
  dep_free(h->depstore);
  cudaFree(h->hooks.adv); cudaFree(h->hooks.old); cudaFree(h->hooks.flux); cudaFree(h->hooks.av); cudaFree(h->hooks.npart_av);
  cudaFree(h->hooks.init_cond);
  {
    auto &M = h->metproc;
    cudaFree(M.d_ab); cudaFree(M.d_cosf); cudaFree(M.UV); cudaFree(M.W); cudaFree(M.PV); cudaFree(M.theta); cudaFree(M.excessoro); cudaFree(M.CLW); cudaFree(M.CIW); cudaFree(M.clw);
    cudaFree(M.uvzlev); cudaFree(M.SF2);
    for (auto &w : M.nw) {
      cudaFree(w.UV); cudaFree(w.W); cudaFree(w.PV); cudaFree(w.theta); cudaFree(w.excessoro); cudaFree(w.uvzlev);
      cudaFree(w.T); cudaFree(w.cosf); cudaFree(w.SF2); cudaFree(w.Q);
    }
    if (M.ev0) cudaEventDestroy(M.ev0);
    if (M.ev1) cudaEventDestroy(M.ev1);
    if (M.evk) cudaEventDestroy(M.evk);
  }
  {
    auto &D = h->dfill;
    cudaFree(D.d_loc); cudaFree(D.d_acc); cudaFree(D.d_mmass); cudaFree(D.d_first); cudaFree(D.d_uoff);
    cudaFree(D.d_blocks); cudaFree(D.d_out); cudaFree(D.d_uniforms);
  }
  for (auto &L : h->lanes) {
    if (L.st) { cudaStreamSynchronize(L.st); cudaStreamDestroy(L.st); }
    scatter_free(L.sw);
    dep_free(L.dep);
    cudaFree(L.d_work); cudaFree(L.d_nlive);
  }
  if (h->ev_ready) cudaEventDestroy(h->ev_ready);
  if (h->st_out) { cudaStreamSynchronize(h->st_out); cudaStreamDestroy(h->st_out); }
  for (auto &e : h->ev_out) if (e) cudaEventDestroy(e);
  if (h->st_in) { cudaStreamSynchronize(h->st_in); cudaStreamDestroy(h->st_in); }
  for (auto &e : h->ev_in) if (e) cudaEventDestroy(e);
  for (auto &e : h->ev_det) if (e) cudaEventDestroy(e);
  for (auto &e : h->ev_dep) if (e) cudaEventDestroy(e);
  fpb_comm_finalize(h);
  for (int k = 0; k < 4; k++) cudaEventDestroy(h->ev[k]);
  cudaStreamDestroy(h->stream);
  delete h;
  return 0;
}

// -------------------------------------------------------------------- RNG --
extern "C" int fpb_set_rannumb(fpb_handle *h, const float *rannumb, int32_t n) {
  if (!h || !rannumb || n < 16) return fail("fpb_set_rannumb: bad argument");
  CK(cudaSetDevice(h->device));
  if (h->d_rannumb) cudaFree(h->d_rannumb);
  h->d_rannumb = nullptr;
  // a few guard entries: the reference's wrap tests are one entry short in
  // places (e.g. src/advance.f90:371 vs rannumb(nrand+1))
  CK(cudaMalloc((void **)&h->d_rannumb, ((size_t)n + 16) * sizeof(float)));
  CK(cudaMemset(h->d_rannumb, 0, ((size_t)n + 16) * sizeof(float)));
  CK(cudaMemcpy(h->d_rannumb, rannumb, (size_t)n * sizeof(float), cudaMemcpyHostToDevice));
  h->maxrand = n;
  return 0;
}

extern "C" int fpb_fill_rannumb(fpb_handle *h, int32_t maxrand, int32_t idummy) {
  if (!h || maxrand < 16) return fail("fpb_fill_rannumb: bad argument");
  std::vector<float> t((size_t)maxrand + 2);
  // 1-based like the Fortran loop, src/FLEXPART.f90:56-59
  for (int i = 1; i <= maxrand - 1; i += 2) gasdev1(h->ran3, idummy, t[i], t[i + 1]);
  gasdev1(h->ran3, idummy, t[maxrand], t[maxrand - 1]);
  return fpb_set_rannumb(h, t.data() + 1, maxrand);
}

// -------------------------------------------------------------------- met --
static int finish_met_upload(fpb_handle *h) {
  if (!h->met_in_flight) return 0;
  CK(cudaEventSynchronize(h->ev_met));
  CK(cudaEventElapsedTime(&h->met_upload_ms, h->ev_met0, h->ev_met));
  h->met_in_flight = false;
  return 0;
}

// One time level: 7 copies-and-packs instead of one per field; everything on the upload stream.
static int enqueue_met(fpb_handle *h, int s, const fpb_met_ptrs *m) {
  const fpb_config &c = h->cfg;
  cudaStream_t st = h->st_met;
  const float *a4[4] = {m->uu, m->vv, m->ww, m->rho}, *g1[1] = {m->drhodz}, *t1[1] = {m->tt};
  const float *p2[2] = {m->uupol, m->vvpol}, *s4[4] = {m->hmix, m->ustar, m->wstar, m->oli}, *tr[1] = {m->tropopause};
  if (upload_group(h, st, (float *)h->A[s], 4, a4, c.nz)) return 1;
  if (upload_group(h, st, h->G[s], 1, g1, c.nz)) return 1;
  if (upload_group(h, st, h->T[s], 1, t1, c.nz)) return 1;
  if (upload_group(h, st, (float *)h->P[s], 2, p2, c.nz)) return 1;
  if (upload_group(h, st, (float *)h->S[s], 4, s4, 1)) return 1;
  if (upload_group(h, st, h->trop[s], 1, tr, 1)) return 1;
  if (c.drydep) { // host vdep(nxmax,nymax,maxspec): species k is "level" k
    const float *vd[1] = {m->vdep};
    if (upload_group(h, st, h->vdep[s], 1, vd, c.nspec)) return 1;
  }
  if (c.wetdep) {
    const float *r4[4] = {m->lsprec, m->convprec, m->tcc, m->ctwc};
    if (upload_group(h, st, (float *)h->R[s], 4, r4, 1)) return 1;
    if (upload_i8(h, st, h->Cl[s], m->clouds, c.nz, h->d.nxd, h->d.nyd, c.nxmax, c.nymax)) return 1;
  }
  return 0;
}

extern "C" int fpb_upload_met_begin(fpb_handle *h, int32_t slot, const fpb_met_ptrs *m) {
  if (!h || !m) return fail("fpb_upload_met: null argument");
  if (slot < 1 || slot > FPB_NSLOTS) return fail("fpb_upload_met: slot must be 1..%d (got %d)", FPB_NSLOTS, slot);
  if (!m->uu || !m->vv || !m->ww || !m->rho || !m->drhodz || !m->hmix || !m->ustar || !m->wstar ||
      !m->oli || !m->tropopause)
    return fail("fpb_upload_met: a mandatory field pointer is null");
  const fpb_config &c = h->cfg;
  if ((c.nglobal || c.sglobal) && (!m->uupol || !m->vvpol))
    return fail("fpb_upload_met: uupol/vvpol required when a pole is in the domain");
  if (c.lsettling && !m->tt) return fail("fpb_upload_met: tt required when lsettling");
  if (c.mdomainfill && !m->tt) return fail("fpb_upload_met: tt required for domain-filling runs (column air mass)");
  if (c.drydep && !m->vdep) return fail("fpb_upload_met: vdep required when drydep");
  if (c.wetdep) {
    if (!m->lsprec || !m->convprec || !m->tcc || !m->clouds || !m->tt)
      return fail("fpb_upload_met: lsprec/convprec/tcc/clouds/tt required when wetdep");
    if (c.readclouds && !m->ctwc) return fail("fpb_upload_met: ctwc required when readclouds");
  }
  CK(cudaSetDevice(h->device));
  if (finish_met_upload(h)) return 1;
  const int s = slot - 1;
  if (alloc_met_slot(h, s)) return 1;
  CK(cudaEventRecord(h->ev_met0, h->st_met));
  if (enqueue_met(h, s, m)) return 1;
  if (build_pairs(h, s, h->st_met)) return 1;
  CK(cudaEventRecord(h->ev_met, h->st_met));
  h->met_in_flight = true;
  h->slot_ready[s] = true;
  return 0;
}

extern "C" int fpb_upload_met_end(fpb_handle *h, float *upload_ms) {
  if (!h) return fail("fpb_upload_met_end: null handle");
  CK(cudaSetDevice(h->device));
  if (finish_met_upload(h)) return 1;
  if (upload_ms) *upload_ms = h->met_upload_ms;
  return 0;
}

extern "C" int fpb_upload_met(fpb_handle *h, int32_t slot, const fpb_met_ptrs *m) {
  if (fpb_upload_met_begin(h, slot, m)) return 1;
  return fpb_upload_met_end(h, nullptr); // the caller may reuse its arrays
}

extern "C" int fpb_upload_met_nest(fpb_handle *h, int32_t slot, int32_t nest, const fpb_met_ptrs *m) {
  if (!h || !m) return fail("fpb_upload_met_nest: null argument");
  if (slot < 1 || slot > FPB_NSLOTS) return fail("fpb_upload_met_nest: slot must be 1..%d (got %d)", FPB_NSLOTS, slot);
  const fpb_config &c = h->cfg;
  if (nest < 1 || nest > c.numbnests) return fail("fpb_upload_met_nest: nest %d outside 1..numbnests=%d", nest, c.numbnests);
  if (!m->uu || !m->vv || !m->ww || !m->rho || !m->drhodz || !m->hmix || !m->ustar || !m->wstar ||
      !m->oli || !m->tropopause)
    return fail("fpb_upload_met_nest: a mandatory field pointer is null");
  if (c.drydep && !m->vdep) return fail("fpb_upload_met_nest: vdep required when drydep");
  if (c.wetdep) {
    if (!m->lsprec || !m->convprec || !m->tcc || !m->clouds || !m->tt)
      return fail("fpb_upload_met_nest: lsprec/convprec/tcc/clouds/tt required when wetdep");
    if (c.readclouds_nest[nest - 1] && !m->ctwc) return fail("fpb_upload_met_nest: ctwc required when readclouds_nest");
  }
  CK(cudaSetDevice(h->device));
  if (finish_met_upload(h)) return 1;
  const int s = slot - 1, l = nest - 1;
  if (alloc_met_slot(h, s)) return 1;
  const int nx = c.nxn[l], ny = c.nyn[l], mx = c.nxmaxn, my = c.nymaxn;
  cudaStream_t st = h->st_met;
  const float *a4[4] = {m->uu, m->vv, m->ww, m->rho}, *g1[1] = {m->drhodz};
  const float *s4[4] = {m->hmix, m->ustar, m->wstar, m->oli}, *tr[1] = {m->tropopause};
  if (upload_group(h, st, (float *)h->An[l][s], 4, a4, c.nz, nx, ny, mx, my)) return 1;
  if (upload_group(h, st, h->Gn[l][s], 1, g1, c.nz, nx, ny, mx, my)) return 1;
  if (upload_group(h, st, (float *)h->Sn[l][s], 4, s4, 1, nx, ny, mx, my)) return 1;
  if (upload_group(h, st, h->tropn[l][s], 1, tr, 1, nx, ny, mx, my)) return 1;
  if (c.drydep) {
    const float *vd[1] = {m->vdep};
    if (upload_group(h, st, h->vdepn[l][s], 1, vd, c.nspec, nx, ny, mx, my)) return 1;
  }
  if (c.wetdep) {
    const float *r4[4] = {m->lsprec, m->convprec, m->tcc, m->ctwc}, *t1[1] = {m->tt};
    if (upload_group(h, st, (float *)h->Rn[l][s], 4, r4, 1, nx, ny, mx, my)) return 1;
    if (upload_group(h, st, h->Tn[l][s], 1, t1, c.nz, nx, ny, mx, my)) return 1;
    if (upload_i8(h, st, h->Cln[l][s], m->clouds, c.nz, nx, ny, mx, my)) return 1;
  }
  CK(cudaStreamSynchronize(st));
  return 0;
}

// page-lock the host's met / particle arrays so that fpb_upload_met_begin and fpb_step_host copy
// asynchronously at full PCIe rate (a Fortran host need not link the CUDA runtime for this)
extern "C" int fpb_host_register(void *p, size_t bytes) {
  if (!p || !bytes) return fail("fpb_host_register: null argument");
  CK(cudaHostRegister(p, bytes, cudaHostRegisterDefault));
  return 0;
}
extern "C" int fpb_host_unregister(void *p) {
  if (!p) return fail("fpb_host_unregister: null argument");
  CK(cudaHostUnregister(p));
  return 0;
}

extern "C" int fpb_set_met_bracket(fpb_handle *h, const int32_t memind[2], const int32_t memtime[2],
                                   int32_t lwindinterv) {
  if (!h || !memind || !memtime) return fail("fpb_set_met_bracket: null argument");
  if (memind[0] < 1 || memind[0] > FPB_NSLOTS || memind[1] < 1 || memind[1] > FPB_NSLOTS || memind[0] == memind[1])
    return fail("fpb_set_met_bracket: memind must be two different slots of 1..%d", FPB_NSLOTS);
  if (!h->slot_ready[memind[0] - 1] || !h->slot_ready[memind[1] - 1])
    return fail("fpb_set_met_bracket: slot %d or %d has not been uploaded", memind[0], memind[1]);
  cudaSetDevice(h->device);
  if (finish_met_upload(h)) return 1; // a read-ahead upload into one of these slots must have landed
  if (memtime[0] == memtime[1]) return fail("fpb_set_met_bracket: memtime(1) == memtime(2)");
  if (lwindinterv == 0) return fail("fpb_set_met_bracket: lwindinterv == 0");
  h->memind[0] = memind[0]; h->memind[1] = memind[1];
  h->memtime[0] = memtime[0]; h->memtime[1] = memtime[1];
  h->lwindinterv = lwindinterv;
  h->have_bracket = true;
  return 0;
}

static void per_step_cfg(fpb_handle *h, DevCfg &d, int itime, int ldeltat);

// -------------------------------------------------------------- particles --
#define H2D(dst, src, T)                                                             \
  do {                                                                               \
    if (!(src)) return fail("fpb_push_particles: array " #src " is null");          \
    CK(cudaMemcpyAsync((dst) + first, (src) + first, (size_t)count * sizeof(T),     \
                       cudaMemcpyHostToDevice, st));                                 \
  } while (0)
#define D2H(dst, src, T)                                                             \
  do {                                                                               \
    if (dst)                                                                         \
      CK(cudaMemcpyAsync((dst) + first, (src) + first, (size_t)count * sizeof(T),   \
                         cudaMemcpyDeviceToHost, h->stream));                        \
  } while (0)

// loop_only: just the arrays the particle loop and conccalc read (fpb_step_host): itrasplit is
// not among them, xscav_frac1 only in backward deposition runs
static int copy_rows_h2d(fpb_handle *h, const DevParticles &d, int first, int count,
                         const fpb_particle_ptrs *p, cudaStream_t st, bool loop_only = false) {
  H2D(d.xtra1, p->xtra1, double); H2D(d.ytra1, p->ytra1, double); H2D(d.ztra1, p->ztra1, float);
  H2D(d.itra1, p->itra1, int32_t); H2D(d.npoint, p->npoint, int32_t);
  H2D(d.nclass, p->nclass, int32_t); H2D(d.idt, p->idt, int32_t);
  H2D(d.itramem, p->itramem, int32_t);
  if (p->itrasplit && !loop_only) H2D(d.itrasplit, p->itrasplit, int32_t);
  H2D(d.uap, p->uap, float); H2D(d.ucp, p->ucp, float); H2D(d.uzp, p->uzp, float);
  H2D(d.us, p->us, float); H2D(d.vs, p->vs, float); H2D(d.ws, p->ws, float);
  H2D(d.cbt, p->cbt, int16_t);
  if (!p->xmass1) return fail("fpb_push_particles: xmass1 is null");
  for (int k = 0; k < h->cfg.nspec; k++) {
    CK(cudaMemcpyAsync(d.xmass1 + (size_t)k * h->cfg.maxpart + first,
                       p->xmass1 + (size_t)k * p->ld + first, (size_t)count * sizeof(float),
                       cudaMemcpyHostToDevice, st));
    if (p->xscav_frac1 && (!loop_only || h->cfg.drybkdep || h->cfg.wetbkdep))
      CK(cudaMemcpyAsync(d.xscav_frac1 + (size_t)k * h->cfg.maxpart + first,
                         p->xscav_frac1 + (size_t)k * p->ld + first, (size_t)count * sizeof(float),
                         cudaMemcpyHostToDevice, st));
  }
  return 0;
}

extern "C" int fpb_push_particles(fpb_handle *h, int32_t first, int32_t count, const fpb_particle_ptrs *p) {
  if (!h || !p) return fail("fpb_push_particles: null argument");
  if (count == 0) return 0;
  if (first < 0 || count < 0 || (int64_t)first + count > h->cfg.maxpart)
    return fail("fpb_push_particles: rows [%d,%d) outside capacity %d", first, first + count, h->cfg.maxpart);
  CK(cudaSetDevice(h->device));
  if (h->permuted && first == 0 && count >= h->numpart) {
    // every row is replaced: drop the permutation
    sortk_iota(h->p.slot, h->cfg.maxpart, h->stream);
    sortk_iota(h->row_of_slot, h->cfg.maxpart, h->stream);
    h->launches += 2;
    h->permuted = false;
  }
  if (!h->permuted) {
    if (copy_rows_h2d(h, h->p, first, count, p, h->stream)) return 1;
  } else {
    // slot-ordered staging, then scatter to the rows the slots live in
    if (copy_rows_h2d(h, h->p_alt, first, count, p, h->stream)) return 1;
    sortk_scatter_from_staging(h->p_alt, h->p, h->row_of_slot, first, count, h->cfg.nspec,
                               p->itrasplit != nullptr, p->xscav_frac1 != nullptr, h->stream);
    h->launches++;
    CK(cudaGetLastError());
  }
  CK(cudaStreamSynchronize(h->stream));
  if (first + count > h->numpart) h->numpart = first + count;
  h->pending_init = true;
  // rows staged ahead of their release (itramem later than the next step, in the run's direction) get
  // their initialize() at that later step: remember until when the init kernel has to be launched
  if (p->itramem) {
    const int ld = h->cfg.ldirect;
    for (int32_t i = first; i < first + count; i++) {
      const long long t = (long long)ld * p->itramem[i];
      if (!h->have_init_until || t > h->init_until) { h->init_until = t; h->have_init_until = true; }
    }
  }
  h->active_rows = -1;
  return 0;
}

extern "C" int fpb_set_numpart(fpb_handle *h, int32_t numpart) {
  if (!h || numpart < 0 || numpart > h->cfg.maxpart) return fail("fpb_set_numpart: bad argument");
  h->numpart = numpart;
  return 0;
}

extern "C" int fpb_pull_particles(fpb_handle *h, int32_t first, int32_t count, const fpb_particle_ptrs *p) {
  if (!h || !p) return fail("fpb_pull_particles: null argument");
  if (count == 0) return 0;
  if (first < 0 || count < 0 || (int64_t)first + count > h->cfg.maxpart)
    return fail("fpb_pull_particles: rows [%d,%d) outside capacity %d", first, first + count, h->cfg.maxpart);
  CK(cudaSetDevice(h->device));
  const DevParticles *src = &h->p;
  if (h->permuted) {
    sortk_gather_to_staging(h->p, h->p_alt, h->row_of_slot, first, count, h->cfg.nspec, h->stream);
    h->launches++;
    CK(cudaGetLastError());
    src = &h->p_alt;
  }
  D2H(p->xtra1, src->xtra1, double); D2H(p->ytra1, src->ytra1, double); D2H(p->ztra1, src->ztra1, float);
  D2H(p->itra1, src->itra1, int32_t); D2H(p->npoint, src->npoint, int32_t);
  D2H(p->nclass, src->nclass, int32_t); D2H(p->idt, src->idt, int32_t);
  D2H(p->itramem, src->itramem, int32_t); D2H(p->itrasplit, src->itrasplit, int32_t);
  D2H(p->uap, src->uap, float); D2H(p->ucp, src->ucp, float); D2H(p->uzp, src->uzp, float);
  D2H(p->us, src->us, float); D2H(p->vs, src->vs, float); D2H(p->ws, src->ws, float);
  D2H(p->cbt, src->cbt, int16_t);
  for (int k = 0; k < h->cfg.nspec; k++) {
    if (p->xmass1)
      CK(cudaMemcpyAsync(p->xmass1 + (size_t)k * p->ld + first,
                         src->xmass1 + (size_t)k * h->cfg.maxpart + first, (size_t)count * sizeof(float),
                         cudaMemcpyDeviceToHost, h->stream));
    if (p->xscav_frac1)
      CK(cudaMemcpyAsync(p->xscav_frac1 + (size_t)k * p->ld + first,
                         src->xscav_frac1 + (size_t)k * h->cfg.maxpart + first,
                         (size_t)count * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
  }
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}

// ------------------------------------------------------------------- sort --
static void per_step_cfg(fpb_handle *h, DevCfg &d, int itime, int ldeltat);
static DevMetSlot slot_view(const fpb_handle *h, int fslot);

// itime_valid: the sort runs at the top of fpb_step(itime), so the key can carry
// the turbulence regime of that step
static int do_sort(fpb_handle *h, bool itime_valid, int itime) {
  const int n = h->numpart;
  if (n <= 1) return 0;
  if (scatter_reserve(h->scatter, (size_t)n, 1)) return fail("%s", scatter_error());
  per_step_cfg(h, h->d_tmp, itime, 0);
  const bool regime = itime_valid && h->have_bracket;
  DevMetSlot met[2];
  if (regime) {
    met[0] = slot_view(h, h->memind[0]);
    met[1] = slot_view(h, h->memind[1]);
  }
  sortk_build_keys(h->d_tmp, h->p, h->d_height, n, h->scatter.keys[0], h->scatter.ids[0], h->d_nlive,
                   h->stream, regime ? met : nullptr);
  int bits = ((sortk_key_bits(h->d_tmp, regime) + 7) / 8) * 8;
  if (bits > 32) bits = 32;
  int cur = 0;
  if (scatter_sort_pairs(h->scatter, (size_t)n, bits, h->stream, &h->launches, &cur)) return fail("%s", scatter_error());
  // The first sort of a large unsorted set (after init_domainfill or a big push) is a random gather: 18
  // separate arrays pay a 32-byte sector per value (measured at 12.5 M rows: 26 GB read, 4.2 ms), so it
  // goes through packed records (2.2 GB, 1.0 ms).  Every later sort finds the rows nearly in order and
  // the plain gather is coalesced again (0.75 ms against 0.95 ms packed).  FPB_SORT_PACKED=0/1 overrides.
  size_t rec_bytes = sortk_packed_bytes(n, h->cfg.nspec);
  bool packed = rec_bytes != 0 && n >= 2000000 && !h->permuted, packed_used = false;
  if (const char *e = getenv("FPB_SORT_PACKED")) packed = rec_bytes != 0 && atoi(e) != 0;
  if (packed && rec_bytes > h->sort_rec_cap) {
    cudaFree(h->sort_rec); h->sort_rec = nullptr; h->sort_rec_cap = 0;
    rec_bytes = sortk_packed_bytes(h->cfg.maxpart, h->cfg.nspec);
    if (cudaMalloc(&h->sort_rec, rec_bytes) != cudaSuccess) { cudaGetLastError(); packed = false; } // (no room: plain gather)
    else h->sort_rec_cap = rec_bytes;
  }
  if (packed) {
    sortk_permute_packed(h->p, h->p_alt, h->scatter.ids[cur], n, h->cfg.nspec, h->stream, h->row_of_slot, h->sort_rec);
    h->launches += 1;
    packed_used = true;
  } else {
    sortk_permute(h->p, h->p_alt, h->scatter.ids[cur], n, h->cfg.nspec, h->stream, h->row_of_slot);
  }
  std::swap(h->p, h->p_alt);
  h->launches += 2;
  unsigned nlive = 0;
  CK(cudaMemcpyAsync(&nlive, h->d_nlive, sizeof nlive, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(h->stream));
  if (packed_used && !getenv("FPB_SORT_PACKED")) { // a one-time buffer: give the memory back
    cudaFree(h->sort_rec); h->sort_rec = nullptr; h->sort_rec_cap = 0;
  }
  h->permuted = true;
  h->active_rows = (int)nlive;
  h->steps_since_sort = 0;
  return 0;
}

extern "C" int fpb_sort_particles(fpb_handle *h) {
  if (!h) return fail("fpb_sort_particles: null handle");
  CK(cudaSetDevice(h->device));
  return do_sort(h, false, 0);
}

// ------------------------------------------------------------------- step --
static void per_step_cfg(fpb_handle *h, DevCfg &d, int itime, int ldeltat) {
  d = h->d;
  d.itime = itime;
  d.ldeltat = ldeltat;
  d.memtime[0] = h->memtime[0];
  d.memtime[1] = h->memtime[1];
  d.lwindinterv = h->lwindinterv;
  d.maxrand = (h->cfg.rng_mode == FPB_RNG_PHILOX) ? (1 << 30) : h->maxrand;
  d.numpart = h->numpart;
}

static void nest_views(const fpb_handle *h, DevStepArgs &a) {
  for (int l = 0; l < FPB_MAXNESTS; l++) {
    for (int m = 0; m < 2; m++) {
      const int s = h->memind[m] - 1;
      DevMetSlot v{};
      v.A = h->An[l][s]; v.AP = nullptr; v.G = h->Gn[l][s]; v.S = h->Sn[l][s]; v.trop = h->tropn[l][s]; v.vdep = h->vdepn[l][s];
      a.metn[l][m] = v;
    }
    a.tropn_lit1[l] = h->tropn[l][0];
  }
}

static DevMetSlot slot_view(const fpb_handle *h, int fslot) {
  DevMetSlot m;
  const int s = fslot - 1;
  m.A = h->A[s]; m.AP = h->AP[s]; m.G = h->G[s]; m.T = h->T[s]; m.P = h->P[s]; m.S = h->S[s]; m.R = h->R[s]; m.C = h->Cl[s]; m.trop = h->trop[s]; m.vdep = h->vdep[s];
  return m;
}

// Validation RNG: replay `nrand=int(ran3(idummy)*real(maxrand-1))+1` in the
// reference's order: one draw per initialize call and one per advance call,
// particles in slot order (src/timemanager.f90:531-611).  Each routine owns a
// SAVEd idummy=-7, so the FIRST call of each re-seeds the shared generator.
static int replay_ran3_indices(fpb_handle *h, int itime) {
  const int n = h->numpart;
  h->h_itra1.resize(n); h->h_itramem.resize(n); h->h_slot.resize(n);
  h->h_nrand_init.assign(n, 1); h->h_nrand_adv.assign(n, 1);
  CK(cudaMemcpyAsync(h->h_itra1.data(), h->p.itra1, (size_t)n * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaMemcpyAsync(h->h_itramem.data(), h->p.itramem, (size_t)n * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
  if (h->permuted)
    CK(cudaMemcpyAsync(h->h_slot.data(), h->row_of_slot, (size_t)n * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  const float scale = (float)(h->maxrand - 1);
  for (int s = 0; s < n; s++) { // slot order = the reference's particle order
    const int r = h->permuted ? h->h_slot[s] : s;
    if (h->h_itra1[r] != itime) continue;
    if (h->h_itramem[r] == itime || itime == 0)
      h->h_nrand_init[s] = (int)(h->ran3.next(h->idummy_init) * scale) + 1;
    h->h_nrand_adv[s] = (int)(h->ran3.next(h->idummy_adv) * scale) + 1;
  }
  CK(cudaMemcpyAsync(h->d_nrand_init, h->h_nrand_init.data(), (size_t)n * sizeof(int32_t), cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemcpyAsync(h->d_nrand_adv, h->h_nrand_adv.data(), (size_t)n * sizeof(int32_t), cudaMemcpyHostToDevice, h->stream));
  return 0;
}

// RECEPTOR block of the particle loop (src/timemanager.f90:563-598) for the rows in `rows`
static int launch_bkdep(fpb_handle *h, const DevCfg &cfg, const DevParticles &rows, cudaStream_t st) {
  const fpb_config &c = h->cfg;
  if (!c.drybkdep && !c.wetbkdep) return 0;
  if (c.wetbkdep && h->rel.numpoint == 0)
    return fail("fpb_step: wetbkdep needs the release heights (fpb_set_releases) for xscav_frac1 = wetscav * "
                "(zpoint2 - zpoint1) * grfraction, src/timemanager.f90:590-591");
  DevBkdepArgs a;
  a.w.cfg = cfg;
  // time level closest to itime - lsynctime/2, src/get_wetscav.f90:113-117 (ltsample = lsynctime)
  const int interp_time = (int)lroundf((float)cfg.itime - 0.5f * (float)c.lsynctime);
  int n = h->memind[1];
  if (abs(h->memtime[0] - interp_time) < abs(h->memtime[1] - interp_time)) n = h->memind[0];
  a.w.met = slot_view(h, n);
  for (int l = 0; l < FPB_MAXNESTS; l++) {
    DevMetSlot v{};
    v.R = h->Rn[l][n - 1]; v.C = h->Cln[l][n - 1]; v.T = h->Tn[l][n - 1];
    a.w.metn[l] = v;
    for (int m = 0; m < 2; m++) {
      DevMetSlot q{};
      q.vdep = h->vdepn[l][h->memind[m] - 1];
      a.vmetn[l][m] = q;
    }
  }
  a.w.p = rows;
  a.w.height = h->d_height;
  a.w.wetgridunc = nullptr; a.w.wetgriduncn = nullptr;
  a.w.dep = DevDepRecords{};
  a.w.ltsample = c.lsynctime;
  a.vmet[0] = slot_view(h, h->memind[0]);
  a.vmet[1] = slot_view(h, h->memind[1]);
  a.zpoint1 = h->rel.d_pts[4]; a.zpoint2 = h->rel.d_pts[5];
  if (c.math_mode == FPB_MATH_STRICT) fpbk_bkdep_strict(a, st); else fpbk_bkdep_fast(a, st);
  h->launches++;
  return 0;
}

// calcfluxes / partpos_average around the step kernels (src/timemanager.f90:614-623); rows = the view the step
// kernels work on, row0 = its first row in the engine's arrays
static bool hooks_on(const fpb_handle *h) { return h->cfg.iflux == 1 || h->cfg.ipout == 3 || h->cfg.linit_cond > 0; }
static int hooks_args(fpb_handle *h, HookArgs &k, const DevCfg &cfg, const DevParticles &rows, int row0) {
  const fpb_config &c = h->cfg;
  if (c.ipout == 3 && (!h->outp.oro || !h->outp.have_q[h->memind[0] - 1] || !h->outp.have_q[h->memind[1] - 1]))
    return fail("ipout = 3 (partpos_average): fpb_set_orography / fpb_upload_pvqv of both time levels are missing");
  if (c.ipout == 3 && (!h->T[h->memind[0] - 1] || !h->T[h->memind[1] - 1]))
    return fail("ipout = 3 (partpos_average): tt of both time levels is missing (fpb_upload_met)");
  k.cfg = cfg;
  k.met[0] = slot_view(h, h->memind[0]); k.met[1] = slot_view(h, h->memind[1]);
  k.Q[0] = h->outp.Q[h->memind[0] - 1]; k.Q[1] = h->outp.Q[h->memind[1] - 1];
  k.oro = h->outp.oro;
  k.height = h->d_height;
  k.p = rows;
  k.iflux = c.iflux == 1; k.ipout3 = c.ipout == 3; k.linit = c.linit_cond;
  k.flags = h->sc.flags + row0;
  k.init_cond = h->hooks.init_cond;
  k.maxspec = c.maxspec;
  k.final_pass = 0;
  k.adv = h->hooks.adv + row0;
  k.old = h->hooks.old ? h->hooks.old + row0 : nullptr;
  k.old_stride = (size_t)c.maxpart;
  k.flux = h->hooks.flux;
  k.npart_av = h->hooks.npart_av;
  k.av = h->hooks.av;
  k.av_stride = (size_t)c.maxpart;
  return 0;
}

extern "C" int fpb_step(fpb_handle *h, int32_t itime, int32_t ldeltat, fpb_step_stats *stats) {
  if (!h) return fail("fpb_step: null handle");
  if (!h->have_bracket) return fail("fpb_step: fpb_set_met_bracket has not been called");
  CK(cudaSetDevice(h->device));
  if (h->cfg.rng_mode != FPB_RNG_PHILOX && !h->d_rannumb) {
    if (fpb_fill_rannumb(h, 1000000, -320)) return 1; // par_mod maxrand, FLEXPART.f90:47
  }
  if (h->cfg.cblflag == 1 && h->cfg.rng_mode == FPB_RNG_REFERENCE && h->cfg.math_mode != FPB_MATH_STRICT) {
    // allowed, but CBL initial velocities use the "defined" draw
  }
  if (h->numpart == 0) {
    if (stats) memset(stats, 0, sizeof *stats);
    return 0;
  }
  if (h->cfg.sort_interval > 0 && h->steps_since_sort >= h->cfg.sort_interval) {
    if (do_sort(h, true, itime)) return 1;
  }
  h->steps_since_sort++;
  if (h->cfg.rng_mode == FPB_RNG_REFERENCE && replay_ran3_indices(h, itime)) return 1;

  DevStepArgs a;
  per_step_cfg(h, a.cfg, itime, ldeltat);
  if (h->active_rows >= 0) a.cfg.numpart = h->active_rows; // dead rows trail after a sort
  a.met[0] = slot_view(h, h->memind[0]);
  a.met[1] = slot_view(h, h->memind[1]);
  a.met_lit1 = slot_view(h, 1);
  nest_views(h, a);
  a.p = h->p;
  a.height = h->d_height;
  a.rannumb = h->d_rannumb;
  a.npart = h->d_npart;
  a.xmass = h->d_xmass;
  a.nrand_init = h->d_nrand_init;
  a.nrand_adv = h->d_nrand_adv;
  a.drygridunc = h->drygridunc;
  a.drygriduncn = h->drygriduncn;
  a.stats = stats ? h->d_stats : nullptr;
  a.work_counter = h->d_work;
  a.sc = h->sc;
  a.dep = DevDepRecords{};
  a.grid_frac = 0.f;
  const bool det_dry = h->cfg.scatter_mode == FPB_SCATTER_DETERMINISTIC && h->cfg.drydep;
  if (det_dry && dep_begin(h, h->depstore, h->numpart, 0, h->stream, a.dep)) return 1;
  if (stats) CK(cudaMemsetAsync(h->d_stats, 0, 8 * sizeof(unsigned long long), h->stream));
  // initialize() can only be due for rows pushed since the last step, or at itime 0
  if (h->pending_init || itime == 0 || (h->have_init_until && (long long)h->cfg.ldirect * itime <= h->init_until)) {
    if (h->cfg.math_mode == FPB_MATH_STRICT) fpbk_init_strict(a, h->stream);
    else fpbk_init_fast(a, h->stream);
    h->launches++;
    h->pending_init = false;
  }
  if (launch_bkdep(h, a.cfg, a.p, h->stream)) return 1;
  HookArgs hk;
  if (hooks_on(h)) {
    if (hooks_args(h, hk, a.cfg, a.p, 0)) return 1;
    fpb_hooks_pre(hk, h->stream);
    h->launches++;
  }
  CK(cudaEventRecord(h->ev[0], h->stream));
  if (h->cfg.math_mode == FPB_MATH_STRICT) fpbk_step_strict(a, h->stream);
  else fpbk_step_fast(a, h->stream);
  CK(cudaEventRecord(h->ev[1], h->stream));
  if (hooks_on(h)) {
    fpb_hooks_post(hk, h->stream);
    h->launches++;
  }
  h->timed_step = true;
  h->launches += 2; // fpb_pbl_kernel + fpb_finish_kernel
  if (det_dry && dep_apply(h, h->scatter, a.dep, h->drygridunc, h->drygriduncn, h->stream)) return 1;
  CK(cudaGetLastError());
  if (stats) {
    unsigned long long hs[8];
    CK(cudaMemcpyAsync(hs, h->d_stats, sizeof hs, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    stats->n_active = (int64_t)hs[0]; stats->n_init = (int64_t)hs[1]; stats->n_terminated = (int64_t)hs[2];
    stats->n_pbl = (int64_t)hs[3]; stats->n_substeps = (int64_t)hs[4]; stats->n_petterssen = (int64_t)hs[5];
    stats->n_nan_cbl = (int64_t)hs[6];
    stats->n_nonfinite = (int64_t)hs[7];
  } else {
    CK(cudaStreamSynchronize(h->stream));
  }
  return 0;
}

// --------------------------------------------------------------- conccalc --
__global__ void receptor_finalize_kernel(float *creceptor, float *acc, int numreceptor, int nspec,
                                         float weight, const float *area /*unused*/, DevCfg c) {
  int t = threadIdx.x;
  if (t < numreceptor * nspec) {
    int n = t / nspec, ks = t % nspec;
    // creceptor(n,ks) += 2*weight*c(ks)/receptorarea(n), src/conccalc.f90:494-496
    creceptor[n + FPB_MAXRECEPTOR * ks] += 2.f * weight * acc[n * nspec + ks] / c.receptorarea[n];
    acc[n * nspec + ks] = 0.f;
  }
}

extern "C" int fpb_conccalc(fpb_handle *h, int32_t itime, float weight) {
  if (!h) return fail("fpb_conccalc: null handle");
  if (h->cfg.ind_samp == -1 && !h->have_bracket) return fail("fpb_conccalc: no met bracket set");
  if (h->numpart == 0) return 0;
  CK(cudaSetDevice(h->device));
  DevConcArgs a;
  per_step_cfg(h, a.cfg, itime, 0);
  a.cfg.weight = weight;
  const int nslots = h->numpart;
  a.met[0] = slot_view(h, h->memind[0]);
  a.met[1] = slot_view(h, h->memind[1]);
  a.p = h->p;
  a.height = h->d_height;
  a.gridunc = h->gridunc;
  a.griduncn = h->griduncn;
  a.crec_acc = h->crec_acc;
  a.slot_base = 0;
  a.rec_vals = nullptr; a.rec_nslots = 0;
  const bool strict = h->cfg.math_mode == FPB_MATH_STRICT;
  const bool det_rec = h->cfg.scatter_mode == FPB_SCATTER_DETERMINISTIC && h->cfg.numreceptor > 0;
  if (det_rec && rec_begin(h, h->depstore, nslots, h->stream, a)) return 1;
  CK(cudaEventRecord(h->ev[2], h->stream));
  if (h->cfg.scatter_mode == FPB_SCATTER_DETERMINISTIC) {
    if (scatter_conccalc_deterministic(h->scatter, a, strict, h->stream, &h->launches)) return fail("%s", scatter_error());
  } else {
    if (strict) fpbk_conccalc_strict(a, h->stream); else fpbk_conccalc_fast(a, h->stream);
    h->launches++;
  }
  if (h->cfg.numreceptor > 0) {
    if (strict) fpbk_receptor_strict(a, h->stream); else fpbk_receptor_fast(a, h->stream);
    if (det_rec) {
      if (scatter_receptor_ordered(a.rec_vals, h->cfg.numreceptor * h->cfg.nspec, nslots, h->crec_acc, h->stream))
        return fail("%s", scatter_error());
      h->launches++;
    }
    receptor_finalize_kernel<<<1, 256, 0, h->stream>>>(h->creceptor, h->crec_acc, h->cfg.numreceptor,
                                                      h->cfg.nspec, weight, nullptr, a.cfg);
    h->launches += 2;
  }
  CK(cudaEventRecord(h->ev[3], h->stream));
  h->timed_conc = true;
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}

// -------------------------------------------------------------- output --
extern "C" int fpb_set_outgrid_geometry(fpb_handle *h, const float *area, const float *volume,
                                        const float *arean, const float *volumen) {
  if (!h || !area || !volume) return fail("fpb_set_outgrid_geometry: null argument");
  const fpb_config &c = h->cfg;
  CK(cudaSetDevice(h->device));
  auto up = [&](float *&d, const float *src, size_t n) -> int {
    if (!d) DA(d, n);
    CK(cudaMemcpy(d, src, n * sizeof(float), cudaMemcpyHostToDevice));
    return 0;
  };
  const size_t n2 = (size_t)c.numxgrid * c.numygrid;
  if (up(h->outp.area[0], area, n2) || up(h->outp.volume[0], volume, n2 * c.numzgrid)) return 1;
  size_t cap = n2 * c.numzgrid;
  if (c.nested_output == 1 && arean && volumen) {
    const size_t m2 = (size_t)c.numxgridn * c.numygridn;
    if (up(h->outp.area[1], arean, m2) || up(h->outp.volume[1], volumen, m2 * c.numzgrid)) return 1;
    cap = std::max(cap, m2 * c.numzgrid);
  }
  if (cap > h->outp.cap) {
    cudaFree(h->outp.block_counts); cudaFree(h->outp.d_i); cudaFree(h->outp.d_r); cudaFree(h->outp.d_density);
    h->outp.block_counts = nullptr; h->outp.d_i = nullptr; h->outp.d_r = nullptr; h->outp.d_density = nullptr;
    DA(h->outp.block_counts, 2 * ((cap + 1023) / 1024));
    DA(h->outp.d_i, cap); DA(h->outp.d_r, cap); DA(h->outp.d_density, cap);
    h->outp.cap = cap;
  }
  if (!h->outp.d_counts) DA(h->outp.d_counts, 2);
  return 0;
}

extern "C" int fpb_set_outgrid_origin(fpb_handle *h, float outlon0, float outlat0, float outlon0n, float outlat0n) {
  if (!h) return fail("fpb_set_outgrid_origin: null handle");
  h->outp.lon0[0] = outlon0; h->outp.lat0[0] = outlat0;
  h->outp.lon0[1] = outlon0n; h->outp.lat0[1] = outlat0n;
  h->outp.have_origin = true;
  return 0;
}

extern "C" int fpb_concoutput_sparse(fpb_handle *h, int32_t nest, int32_t which, int32_t ks, int32_t kp,
                                     int32_t nage, float outnum, float tot_mu, int32_t loutaver,
                                     int32_t *sp_count_i, int32_t *sparse_dump_i, int32_t *sp_count_r,
                                     float *sparse_dump_r) {
  if (!h || !sp_count_i || !sparse_dump_i || !sp_count_r || !sparse_dump_r)
    return fail("fpb_concoutput_sparse: null argument");
  const fpb_config &c = h->cfg;
  if (nest < 0 || nest > 1 || (nest == 1 && c.nested_output != 1)) return fail("fpb_concoutput_sparse: no such output grid %d", nest);
  if (which < 0 || which > 3) return fail("fpb_concoutput_sparse: which = %d", which);
  if (which == 3 && (!h->outp.have_origin || !h->have_bracket))
    return fail("fpb_concoutput_sparse: the mixing-ratio record needs fpb_set_outgrid_origin and a met bracket");
  if (ks < 1 || ks > c.nspec || kp < 1 || kp > c.maxpointspec_act || nage < 1 || nage > c.nageclass)
    return fail("fpb_concoutput_sparse: (ks, kp, nage) = (%d, %d, %d) out of range", ks, kp, nage);
  if (!h->outp.area[nest] || !h->outp.volume[nest]) return fail("fpb_concoutput_sparse: fpb_set_outgrid_geometry has not been called");
  if (which == 2 && !c.wetdep) return fail("fpb_concoutput_sparse: the run has no wet deposition");
  CK(cudaSetDevice(h->device));
  const int nxg = nest ? c.numxgridn : c.numxgrid, nyg = nest ? c.numygridn : c.numygrid;
  const size_t n2 = (size_t)nxg * nyg;
  SparseDumpArgs a;
  a.which = which;
  const bool conc = which == 0 || which == 3;
  a.ncells = (int)(conc ? n2 * c.numzgrid : n2);
  const float *g = conc ? (nest ? h->griduncn : h->gridunc)
                 : which == 1 ? (nest ? h->drygriduncn : h->drygridunc) : (nest ? h->wetgriduncn : h->wetgridunc);
  // device layout: [nage][class][kp][ks][cells]
  const size_t inner = (size_t)a.ncells;
  a.class_stride = (size_t)c.maxpointspec_act * c.nspec * inner;
  a.grid = g + ((((size_t)(nage - 1) * c.nclassunc) * c.maxpointspec_act + (kp - 1)) * c.nspec + (ks - 1)) * inner;
  a.nclassunc = c.nclassunc;
  a.geom = conc ? h->outp.volume[nest] : h->outp.area[nest];
  a.density = nullptr;
  if (which == 3) { // densityoutgrid, src/concoutput.f90:164-190
    if (c.numzgrid > 32) return fail("fpb_concoutput_sparse: numzgrid > 32");
    DensityArgs d;
    d.A = slot_view(h, h->memind[1]).A;
    d.nxd = h->d.nxd; d.plane = h->d.nxd * h->d.nyd;
    d.numx = nxg; d.numy = nyg; d.numz = c.numzgrid;
    d.outlon0 = h->outp.lon0[nest]; d.outlat0 = h->outp.lat0[nest];
    d.dxout = nest ? c.dxoutn : c.dxout; d.dyout = nest ? c.dyoutn : c.dyout;
    d.xlon0 = c.xlon0; d.ylat0 = c.ylat0; d.dx = c.dx; d.dy = c.dy;
    d.nxmin1 = c.nxmin1; d.nymin1 = c.nymin1;
    for (int kz = 1; kz <= c.numzgrid; kz++) {
      const float halfheight = (kz == 1) ? c.outheight[0] / 2.f : (c.outheight[kz - 1] + c.outheight[kz - 2]) / 2.f;
      int kzz = 2;
      for (; kzz <= c.nz; kzz++)
        if (h->height[kzz - 2] < halfheight && h->height[kzz - 1] > halfheight) break;
      kzz = std::max(std::min(kzz, (int)c.nz), 2);
      d.kzz[kz - 1] = kzz;
      d.dz1[kz - 1] = halfheight - h->height[kzz - 2];
      d.dz2[kz - 1] = h->height[kzz - 1] - halfheight;
    }
    d.density = h->outp.d_density;
    fpb_density_outgrid(d, h->stream);
    h->launches++;
    a.density = h->outp.d_density;
  }
  a.ldirect = c.ldirect;
  a.outnum = outnum; a.tot_mu = tot_mu; a.loutaver_abs = (float)std::abs(loutaver);
  a.index_offset = conc ? (int)n2 : 0; // kz is 1-based in ix+jy*numxgrid+kz*numxgrid*numygrid
  a.block_counts = h->outp.block_counts;
  a.out_i = h->outp.d_i; a.out_r = h->outp.d_r; a.counts = h->outp.d_counts;
  fpb_sparse_dump(a, h->stream);
  h->launches += 3;
  int counts[2];
  CK(cudaMemcpyAsync(counts, a.counts, sizeof counts, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(h->stream));
  *sp_count_i = counts[0]; *sp_count_r = counts[1];
  if (counts[0] > 0) CK(cudaMemcpyAsync(sparse_dump_i, a.out_i, (size_t)counts[0] * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
  if (counts[1] > 0) CK(cudaMemcpyAsync(sparse_dump_r, a.out_r, (size_t)counts[1] * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}

// ---------------------------------------------------------- partoutput --
extern "C" int fpb_set_orography(fpb_handle *h, const float *oro) {
  if (!h || !oro) return fail("fpb_set_orography: null argument");
  CK(cudaSetDevice(h->device));
  if (!h->outp.oro) DA(h->outp.oro, (size_t)h->d.nxd * h->d.nyd);
  if (finish_met_upload(h)) return 1;
  const float *o1[1] = {oro};
  if (upload_group(h, h->st_met, h->outp.oro, 1, o1, 1)) return 1;
  CK(cudaStreamSynchronize(h->st_met));
  return 0;
}

extern "C" int fpb_upload_pvqv(fpb_handle *h, int32_t slot, const float *pv, const float *qv) {
  if (!h || !pv || !qv) return fail("fpb_upload_pvqv: null argument");
  if (slot < 1 || slot > FPB_NSLOTS) return fail("fpb_upload_pvqv: slot %d", slot);
  CK(cudaSetDevice(h->device));
  const int s = slot - 1;
  if (!h->outp.Q[s]) DA(h->outp.Q[s], (size_t)h->d.nxd * h->d.nyd * h->cfg.nz);
  if (finish_met_upload(h)) return 1;
  const float *q2[2] = {pv, qv};
  if (upload_group(h, h->st_met, reinterpret_cast<float *>(h->outp.Q[s]), 2, q2, h->cfg.nz)) return 1;
  CK(cudaStreamSynchronize(h->st_met));
  h->outp.have_q[s] = true;
  return 0;
}

extern "C" int fpb_fetch_fluxes(fpb_handle *h, float *flux, int32_t zero) {
  if (!h) return fail("fpb_fetch_fluxes: null handle");
  if (h->cfg.iflux != 1) return fail("fpb_fetch_fluxes: the engine was created with iflux = %d (no flux calculation)", h->cfg.iflux);
  CK(cudaSetDevice(h->device));
  if (flux) CK(cudaMemcpyAsync(flux, h->hooks.flux, h->hooks.nflux * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
  if (zero) CK(cudaMemsetAsync(h->hooks.flux, 0, h->hooks.nflux * sizeof(float), h->stream)); // src/fluxoutput.f90:288-303
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}

extern "C" int fpb_fetch_init_cond(fpb_handle *h, float *init_cond, int32_t zero) {
  if (!h) return fail("fpb_fetch_init_cond: null handle");
  if (h->cfg.linit_cond <= 0) return fail("fpb_fetch_init_cond: the engine was created with linit_cond = 0");
  CK(cudaSetDevice(h->device));
  if (init_cond) CK(cudaMemcpyAsync(init_cond, h->hooks.init_cond, h->hooks.ninit * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
  if (zero) CK(cudaMemsetAsync(h->hooks.init_cond, 0, h->hooks.ninit * sizeof(float), h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}

// src/timemanager.f90:733-737: initial_cond_calc for every particle still active at the end of the run
extern "C" int fpb_initial_cond_final(fpb_handle *h, int32_t itime) {
  if (!h) return fail("fpb_initial_cond_final: null handle");
  if (h->cfg.linit_cond <= 0) return fail("fpb_initial_cond_final: the engine was created with linit_cond = 0");
  if (h->cfg.linit_cond == 1 && !h->have_bracket) return fail("fpb_initial_cond_final: fpb_set_met_bracket has not been called");
  if (h->numpart == 0) return 0;
  CK(cudaSetDevice(h->device));
  HookArgs k;
  DevCfg cfg;
  per_step_cfg(h, cfg, itime, 0);
  if (h->active_rows >= 0) cfg.numpart = h->active_rows;
  const int ipout = h->cfg.ipout;
  h->cfg.ipout = 0; // (no average fields needed here)
  const int rc = hooks_args(h, k, cfg, h->p, 0);
  h->cfg.ipout = ipout;
  if (rc) return 1;
  k.iflux = 0; k.ipout3 = 0; k.final_pass = 1;
  fpb_hooks_post(k, h->stream);
  h->launches++;
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}

extern "C" int fpb_fetch_partpos_average(fpb_handle *h, int32_t numpart, const fpb_partav_ptrs *out, int32_t zero) {
  if (!h || !out) return fail("fpb_fetch_partpos_average: null argument");
  if (h->cfg.ipout != 3) return fail("fpb_fetch_partpos_average: the engine was created with ipout = %d", h->cfg.ipout);
  if (numpart < 0 || numpart > h->cfg.maxpart) return fail("fpb_fetch_partpos_average: numpart %d outside capacity", numpart);
  CK(cudaSetDevice(h->device));
  const size_t mp = (size_t)h->cfg.maxpart, nb = (size_t)numpart * sizeof(float);
  float *dst[14] = {out->cartx, out->carty, out->cartz, out->z, out->topo, out->pv, out->qv, out->tt, out->uu, out->vv,
                    out->rho, out->tro, out->hmix, out->energy};
  if (numpart > 0) {
    if (out->npart_av) CK(cudaMemcpyAsync(out->npart_av, h->hooks.npart_av, nb, cudaMemcpyDeviceToHost, h->stream));
    for (int q = 0; q < 14; q++)
      if (dst[q]) CK(cudaMemcpyAsync(dst[q], h->hooks.av + q * mp, nb, cudaMemcpyDeviceToHost, h->stream));
  }
  if (zero) { // src/partoutput_average.f90:171-187
    CK(cudaMemsetAsync(h->hooks.npart_av, 0, mp * sizeof(int32_t), h->stream));
    CK(cudaMemsetAsync(h->hooks.av, 0, mp * 14 * sizeof(float), h->stream));
  }
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}

extern "C" int fpb_partoutput(fpb_handle *h, int32_t itime, int32_t *nrecords, const fpb_partout_ptrs *out) {
  if (!h || !nrecords || !out) return fail("fpb_partoutput: null argument");
  if (!h->have_bracket) return fail("fpb_partoutput: fpb_set_met_bracket has not been called");
  if (!h->outp.oro || !h->outp.have_q[h->memind[0] - 1] || !h->outp.have_q[h->memind[1] - 1])
    return fail("fpb_partoutput: fpb_set_orography / fpb_upload_pvqv of both time levels are missing");
  *nrecords = 0;
  if (h->numpart == 0) return 0;
  CK(cudaSetDevice(h->device));
  auto &O = h->outp;
  const size_t mp = h->cfg.maxpart;
  if (!O.po_count) {
    DA(O.po_count, 1);
    DA(O.po_counts, 2 * ((mp + 1023) / 1024));
    DA(O.po_i[0], mp); DA(O.po_i[1], mp);
    for (auto &q : O.po_f) DA(q, mp);
    DA(O.po_mass, mp * h->cfg.nspec);
  }
  PartoutArgs a;
  per_step_cfg(h, a.cfg, itime, 0);
  a.met[0] = slot_view(h, h->memind[0]); a.met[1] = slot_view(h, h->memind[1]);
  a.Q[0] = O.Q[h->memind[0] - 1]; a.Q[1] = O.Q[h->memind[1] - 1];
  a.oro = O.oro;
  a.height = h->d_height;
  a.p = h->p;
  a.row_of_slot = h->row_of_slot;
  a.permuted = h->permuted ? 1 : 0;
  a.numpart = h->numpart;
  a.block_counts = O.po_counts; a.count = O.po_count;
  a.npoint = O.po_i[0]; a.itramem = O.po_i[1];
  a.xlon = O.po_f[0]; a.ylat = O.po_f[1]; a.ztra1 = O.po_f[2]; a.topo = O.po_f[3]; a.pvi = O.po_f[4];
  a.qvi = O.po_f[5]; a.rhoi = O.po_f[6]; a.hmixi = O.po_f[7]; a.tri = O.po_f[8]; a.tti = O.po_f[9];
  a.xmass1 = O.po_mass;
  fpb_partoutput_launch(a, h->stream);
  h->launches += 3;
  int n = 0;
  CK(cudaMemcpyAsync(&n, O.po_count, sizeof n, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(h->stream));
  *nrecords = n;
  if (n == 0) return 0;
  auto get = [&](void *dst, const void *src, size_t bytes) -> int {
    if (dst) CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, h->stream));
    return 0;
  };
  const size_t nb = (size_t)n * 4;
  if (get(out->npoint, a.npoint, nb) || get(out->itramem, a.itramem, nb) || get(out->xlon, a.xlon, nb) ||
      get(out->ylat, a.ylat, nb) || get(out->ztra1, a.ztra1, nb) || get(out->topo, a.topo, nb) ||
      get(out->pvi, a.pvi, nb) || get(out->qvi, a.qvi, nb) || get(out->rhoi, a.rhoi, nb) ||
      get(out->hmixi, a.hmixi, nb) || get(out->tri, a.tri, nb) || get(out->tti, a.tti, nb))
    return 1;
  if (out->xmass1)
    for (int ks = 0; ks < h->cfg.nspec; ks++)
      if (get(out->xmass1 + (size_t)ks * out->ld, a.xmass1 + (size_t)ks * mp, nb)) return 1;
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}

// ------------------------------------------------------------ releases --
extern "C" int fpb_set_releases(fpb_handle *h, const fpb_release_points *r) {
  if (!h || !r) return fail("fpb_set_releases: null argument");
  if (r->numpoint != h->cfg.numpoint)
    return fail("fpb_set_releases: numpoint %d differs from the configuration's %d", r->numpoint, h->cfg.numpoint);
  const float *src[6] = {r->xpoint1, r->ypoint1, r->xpoint2, r->ypoint2, r->zpoint1, r->zpoint2};
  for (auto q : src) if (!q) return fail("fpb_set_releases: a coordinate array is null");
  if (!r->ireleasestart || !r->ireleaseend) return fail("fpb_set_releases: release times are null");
  CK(cudaSetDevice(h->device));
  auto &R = h->rel;
  const int n = r->numpoint;
  for (int k = 0; k < 6; k++) {
    if (!R.d_pts[k]) DA(R.d_pts[k], n);
    CK(cudaMemcpy(R.d_pts[k], src[k], (size_t)n * sizeof(float), cudaMemcpyHostToDevice));
  }
  if (!R.d_offsets) DA(R.d_offsets, n + 1);
  if (!R.d_block_counts) DA(R.d_block_counts, (size_t)(h->cfg.maxpart + 1023) / 1024);
  if (!R.d_out) DA(R.d_out, 2);
  R.numpoint = n;
  R.itsplit = r->itsplit;
  R.start.assign(r->ireleasestart, r->ireleasestart + n);
  R.end.assign(r->ireleaseend, r->ireleaseend + n);
  R.xmasssave.assign(n, 0.f);
  R.idum = -7; R.iy = 0;
  if (r->mp_pid > 0) { // src/mpi_mod.f90:331-335, src/releaseparticles_mpi.f90:65
    long long m = ((244LL * 181LL) * ((long long)(r->mp_pid - 83) * 359LL)) % 104729LL;
    if (m < 0) m = -m;
    R.idum += (int)(-m);
  }
  return 0;
}

extern "C" int fpb_releaseparticles(fpb_handle *h, int32_t itime, int32_t *numpart, int32_t *n_released) {
  if (!h) return fail("fpb_releaseparticles: null handle");
  auto &R = h->rel;
  if (R.numpoint == 0) return fail("fpb_releaseparticles: fpb_set_releases has not been called");
  const fpb_config &c = h->cfg;
  if (numpart) *numpart = h->numpart;
  if (n_released) *n_released = 0;
  // release counts of this call, src/releaseparticles.f90:89-123
  std::vector<int32_t> offsets(R.numpoint + 1, 0);
  for (int i = 0; i < R.numpoint; i++) {
    const int t0 = R.start[i], t1 = R.end[i];
    int numrel = 0;
    if (itime >= t0 && itime <= t1) {
      if (t0 != t1) {
        float rfraction = std::fabs((float)h->npart[i] * (float)c.lsynctime / (float)(t1 - t0));
        if (itime == t0 || itime == t1) rfraction = rfraction / 2.f;
        rfraction = rfraction * 1.f; // average_timecorrect (no EMISVAR files)
        rfraction = rfraction + R.xmasssave[i];
        numrel = (int)rfraction;
        R.xmasssave[i] = rfraction - (float)numrel;
      } else {
        numrel = h->npart[i];
      }
    }
    offsets[i + 1] = offsets[i] + numrel;
  }
  const int m = offsets[R.numpoint];
  if (m == 0) return 0;
  CK(cudaSetDevice(h->device));
  DevReleaseArgs a;
  per_step_cfg(h, a.cfg, itime, 0);
  a.p = h->p;
  a.row_of_slot = h->row_of_slot;
  a.permuted = h->permuted ? 1 : 0;
  a.numpart_old = h->numpart;
  a.numpoint = R.numpoint; a.n_new = m; a.itsplit = R.itsplit;
  a.ztop = h->height[c.nz - 1];
  a.xpoint1 = R.d_pts[0]; a.ypoint1 = R.d_pts[1]; a.xpoint2 = R.d_pts[2]; a.ypoint2 = R.d_pts[3];
  a.zpoint1 = R.d_pts[4]; a.zpoint2 = R.d_pts[5];
  a.offsets = R.d_offsets;
  a.uniforms = nullptr;
  a.xmass = h->d_xmass; a.npart = h->d_npart;
  a.block_counts = R.d_block_counts;
  a.out = R.d_out;
  CK(cudaMemcpyAsync(R.d_offsets, offsets.data(), offsets.size() * sizeof(int32_t), cudaMemcpyHostToDevice, h->stream));
  std::vector<float> u;
  if (c.rng_mode == FPB_RNG_REFERENCE) { // the reference's ran1 stream: x, y, nclass, z per particle
    u.resize((size_t)4 * m);
    for (auto &v : u) v = R.ran1();
    if (R.uniforms_cap < u.size()) {
      cudaFree(R.d_uniforms); R.d_uniforms = nullptr;
      DA(R.d_uniforms, u.size());
      R.uniforms_cap = u.size();
    }
    CK(cudaMemcpyAsync(R.d_uniforms, u.data(), u.size() * sizeof(float), cudaMemcpyHostToDevice, h->stream));
    a.uniforms = R.d_uniforms;
  }
  const int out0[2] = {h->numpart, 0};
  CK(cudaMemcpyAsync(R.d_out, out0, sizeof out0, cudaMemcpyHostToDevice, h->stream));
  if (c.math_mode == FPB_MATH_STRICT) fpbk_release_strict(a, h->stream); else fpbk_release_fast(a, h->stream);
  h->launches += 3;
  int out[2];
  CK(cudaMemcpyAsync(out, R.d_out, sizeof out, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(h->stream));
  if (out[1] < m)
    return fail("RELEASEPARTICLES: TOTAL NUMBER OF PARTICLES REQUIRED (%d new, %d free slots) EXCEEDS THE "
                "MAXIMUM ALLOWED NUMBER %d", m, out[1], c.maxpart);
  h->numpart = out[0];
  h->pending_init = true;
  h->active_rows = -1;
  if (numpart) *numpart = h->numpart;
  if (n_released) *n_released = m;
  return 0;
}

extern "C" int fpb_split_particles(fpb_handle *h, int32_t itime, int32_t *numpart) {
  if (!h) return fail("fpb_split_particles: null handle");
  if (numpart) *numpart = h->numpart;
  if (h->numpart == 0) return 0;
  CK(cudaSetDevice(h->device));
  auto &R = h->rel;
  if (!R.d_block_counts) DA(R.d_block_counts, (size_t)(h->cfg.maxpart + 1023) / 1024);
  if (!R.d_out) DA(R.d_out, 2);
  DevSplitArgs a;
  per_step_cfg(h, a.cfg, itime, 0);
  a.p = h->p;
  a.row_of_slot = h->row_of_slot;
  a.permuted = h->permuted ? 1 : 0;
  a.numpart_old = h->numpart;
  a.block_counts = R.d_block_counts;
  a.total = R.d_out;
  if (h->cfg.math_mode == FPB_MATH_STRICT) fpbk_split_strict(a, h->stream); else fpbk_split_fast(a, h->stream);
  h->launches += 3;
  int total = 0;
  CK(cudaMemcpyAsync(&total, R.d_out, sizeof total, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(h->stream));
  const int room = h->cfg.maxpart - h->numpart;
  h->numpart += total < room ? total : room;
  h->active_rows = -1;
  if (numpart) *numpart = h->numpart;
  return 0;
}

// ------------------------------------------------------------ domain fill --
// gridarea(jy) of src/init_domainfill.f90:81-130 (cos evaluated in double and rounded once: the
// convention of the strict kernels and of the oracle)
static void domainfill_gridarea(const fpb_config &c, const int ny_sn[2], std::vector<float> &ga) {
  const float pi = 3.14159265f, r_earth = 6.371e6f, pih = pi / 180.f;
  auto cosf_cr = [](float x) { return (float)cos((double)x); };
  auto zone = [&](float cp, float cm) {
    return (cp < cm) ? sqrtf(r_earth * r_earth - cp * cp) - sqrtf(r_earth * r_earth - cm * cm)
                     : sqrtf(r_earth * r_earth - cm * cm) - sqrtf(r_earth * r_earth - cp * cp);
  };
  for (int jy = ny_sn[0]; jy <= ny_sn[1]; jy++) {
    const float ylat = c.ylat0 + (float)jy * c.dy, ylatp = ylat + 0.5f * c.dy, ylatm = ylat - 0.5f * c.dy;
    float hzone;
    if ((ylatm < 0.f) && (ylatp > 0.f)) hzone = 1.f / c.dyconst;
    else hzone = zone(cosf_cr(ylatp * pih) * r_earth, cosf_cr(ylatm * pih) * r_earth);
    ga[jy] = 2.f * pi * r_earth * hzone * c.dx / 360.f;
  }
  if (c.sglobal) {
    const float cp = cosf_cr((c.ylat0 + 0.5f * c.dy) * pih) * r_earth;
    ga[0] = 2.f * pi * r_earth * (sqrtf(r_earth * r_earth - 0.f * 0.f) - sqrtf(r_earth * r_earth - cp * cp)) * c.dx / 360.f;
  }
  if (c.nglobal) {
    const float ylat = c.ylat0 + (float)c.nymin1 * c.dy;
    const float cm = cosf_cr((ylat - 0.5f * c.dy) * pih) * r_earth;
    ga[c.nymin1] = 2.f * pi * r_earth * (sqrtf(r_earth * r_earth - 0.f * 0.f) - sqrtf(r_earth * r_earth - cm * cm)) * c.dx / 360.f;
  }
}

// Second half of init_domainfill (src/init_domainfill.f90:287-389) for a limited box: fewer release
// heights per boundary column ("on the order of nz"), memorised for boundcond_domainfill.  A few
// thousand columns, once per run: the pressure profiles come from the device, the sequential
// pnew recurrence and everything that stays constant per release location are worked out here.
static int domainfill_boundary_setup(fpb_handle *h, const DomainfillArgs &a, const float *d_colmass, float colmasstotal,
                                     int numcolumn) {
  const fpb_config &c = h->cfg;
  auto &D = h->dfill;
  const int nz = c.nz, nx0 = a.nx0, nx1 = a.nx1, ny0 = a.ny0, ny1 = a.ny1;
  const int MAXCOLUMN = 3000; // par_mod
  std::vector<float> colmass((size_t)a.ncols);
  CK(cudaMemcpy(colmass.data(), d_colmass, colmass.size() * sizeof(float), cudaMemcpyDeviceToHost));
  float fractus = (float)numcolumn / (float)nz;
  fractus = sqrtf(fractus > 1.f ? fractus : 1.f) / 2.f;
  std::vector<int2> cols;
  for (int jy = ny0; jy <= ny1; jy++)
    for (int ix = nx0; ix <= nx1; ix++)
      if (ix == nx0 || ix == nx1 || jy == ny0 || jy == ny1) cols.push_back(make_int2(ix, jy));
  int2 *d_cols = nullptr;
  float *d_pp = nullptr;
  std::vector<float> pps(cols.size() * nz);
  if (dalloc(&d_pp, pps.size())) return 1;
  if (cudaMalloc((void **)&d_cols, cols.size() * sizeof(int2)) != cudaSuccess) { cudaFree(d_pp); return fail("cudaMalloc failed"); }
  cudaMemcpyAsync(d_cols, cols.data(), cols.size() * sizeof(int2), cudaMemcpyHostToDevice, h->stream);
  fpb_domainfill_profiles(a, d_cols, (int)cols.size(), d_pp, h->stream);
  h->launches++;
  cudaError_t e = cudaMemcpyAsync(pps.data(), d_pp, pps.size() * sizeof(float), cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  cudaFree(d_cols); cudaFree(d_pp);
  if (e != cudaSuccess) return fail("fpb_init_domainfill: boundary profiles: %s", cudaGetErrorString(e));

  // zcolumn_we(k, jy, :) / zcolumn_sn(k, ix, :) with element 0 and the element after the last = 0
  std::vector<std::vector<float>> zwe[2], zsn[2];
  for (int k = 0; k < 2; k++) { zwe[k].assign(c.ny + 1, {}); zsn[k].assign(c.nx + 1, {}); }
  for (size_t q = 0; q < cols.size(); q++) {
    const int ix = cols[q].x, jy = cols[q].y;
    const float cm = colmass[(size_t)(jy - ny0) * a.ncolx + (ix - nx0)];
    const int ncolumn = (int)roundf(0.999f / fractus * (float)h->npart[0] * cm / colmasstotal); // nint()
    if (ncolumn > MAXCOLUMN) return fail("fpb_init_domainfill: maxcolumn too small (%d release heights in a boundary column)", ncolumn);
    if (ncolumn == 0) continue;
    const float *pp = pps.data() + q * nz - 1; // pp[1..nz]
    std::vector<float> z((size_t)ncolumn + 2, 0.f);
    const float deltacol = (pp[1] - pp[nz]) / (float)ncolumn;
    float pnew = pp[1] + deltacol / 2.f;
    for (int j = 1; j <= ncolumn; j++) {
      pnew = pnew - deltacol;
      for (int kz = 1; kz <= nz - 1; kz++)
        if ((pp[kz] >= pnew) && (pp[kz + 1] < pnew)) {
          const float dz1 = pp[kz] - pnew, dz2 = pnew - pp[kz + 1];
          const float dz = 1.f / (dz1 + dz2);
          float zposition = (h->height[kz - 1] * dz2 + h->height[kz] * dz1) * dz;
          if (zposition > h->height[nz - 1] - 0.5f) zposition = h->height[nz - 1] - 0.5f;
          z[j] = zposition;
        }
    }
    if (ix == nx0) zwe[0][jy] = z;
    if (ix == nx1) zwe[1][jy] = z;
    if (jy == ny0) zsn[0][ix] = z;
    if (jy == ny1) zsn[1][ix] = z;
  }

  // the release locations in the order boundcond_domainfill visits them
  std::vector<BcLoc> loc;
  D.loc_draws.clear();
  const float ztop = h->height[nz - 1];
  auto add_column = [&](bool we, int k /*0,1*/, int idx, const std::vector<float> &Z, float cosfact) {
    const int ncol = (int)Z.size() - 2;
    for (int j = 1; j <= ncol; j++) {
      BcLoc q;
      float deltaz;
      if (j == 1) deltaz = (Z[2] + Z[1]) / 2.f;
      else if (j == ncol) deltaz = (Z[j] - Z[j - 2]) / 2.f;
      else deltaz = (Z[j + 1] - Z[j - 1]) / 2.f;
      const bool low = we ? idx == ny0 : idx == nx0, high = we ? idx == ny1 : idx == nx1;
      if (we) q.boundarea = (low || high) ? deltaz * 111198.5f / 2.f * c.dy : deltaz * 111198.5f * c.dy;
      else q.boundarea = (low || high) ? deltaz * 111198.5f / 2.f * cosfact * c.dx : deltaz * 111198.5f * cosfact * c.dx;
      int indz = 1;
      for (int i = 2; i <= nz; i++)
        if (h->height[i - 1] > Z[j]) { indz = i - 1; break; }
      q.indz = indz;
      q.dz1 = Z[j] - h->height[indz - 1];
      q.dz2 = h->height[indz] - Z[j];
      q.dz = 1.f / (q.dz1 + q.dz2);
      q.flags = (we ? BC_WE : 0) | (k ? BC_K2 : 0) | (low ? BC_EDGE_LOW : (high ? BC_EDGE_HIGH : 0));
      q.zb = 0.f;
      if (j == 1) q.za = Z[1] + (Z[2] - Z[1]) / 4.f;
      else if (j == ncol) q.za = (2.f * Z[j] + Z[j - 1] + ztop) / 4.f;
      else { q.za = Z[j - 1]; q.zb = Z[j + 1] - Z[j - 1]; q.flags |= BC_ZDRAW; }
      q.idx = idx;
      q.gx = we ? (k ? nx1 : nx0) : idx;
      q.gy = we ? idx : (k ? ny1 : ny0);
      loc.push_back(q);
      D.loc_draws.push_back((q.flags & BC_ZDRAW) ? 3 : 2);
    }
  };
  for (int jy = ny0; jy <= ny1; jy++)
    for (int k = 0; k < 2; k++)
      if (zwe[k][jy].size() > 2) add_column(true, k, jy, zwe[k][jy], 0.f);
  for (int ix = nx0; ix <= nx1; ix++)
    for (int k = 0; k < 2; k++) {
      const float ylat = c.ylat0 + (float)(k ? ny1 : ny0) * c.dy;
      const float cosfact = (float)cos((double)(ylat * (3.14159265f / 180.f)));
      if (zsn[k][ix].size() > 2) add_column(false, k, ix, zsn[k][ix], cosfact);
    }
  D.nloc = (int)loc.size();
  cudaFree(D.d_loc); cudaFree(D.d_acc); cudaFree(D.d_mmass); cudaFree(D.d_first); cudaFree(D.d_uoff);
  cudaFree(D.d_blocks); cudaFree(D.d_out);
  D.d_loc = nullptr; D.d_acc = nullptr; D.d_mmass = nullptr; D.d_first = nullptr; D.d_uoff = nullptr;
  D.d_blocks = nullptr; D.d_out = nullptr;
  const size_t n = loc.size() ? loc.size() : 1;
  CK(cudaMalloc((void **)&D.d_loc, n * sizeof(BcLoc)));
  DA(D.d_acc, n); DA(D.d_mmass, n); DA(D.d_first, n + 1); DA(D.d_uoff, n);
  DA(D.d_blocks, (size_t)(c.maxpart + 1023) / 1024 + 1); DA(D.d_out, 2);
  if (!loc.empty()) CK(cudaMemcpy(D.d_loc, loc.data(), loc.size() * sizeof(BcLoc), cudaMemcpyHostToDevice));
  return 0;
}

extern "C" int fpb_init_domainfill(fpb_handle *h, float xpoint1, float ypoint1, float xpoint2, float ypoint2,
                                   int32_t itsplit, int32_t *numpart, fpb_domainfill_info *info) {
  if (!h) return fail("fpb_init_domainfill: null handle");
  const fpb_config &c = h->cfg;
  if (c.mdomainfill != 1)
    return fail("fpb_init_domainfill: mdomainfill = %d; only the air-mass tracer (MDOMAINFILL = 1) is built "
                "(2 = stratospheric ozone needs the PV test of src/init_domainfill.f90:236-251)", c.mdomainfill);
  if (h->numpart != 0) return fail("fpb_init_domainfill: the particle set is not empty (ipin = 1 restarts are host-side)");
  CK(cudaSetDevice(h->device));
  auto &D = h->dfill;
  // :55-76
  D.nx_we[0] = std::max((int)xpoint1, 0);
  D.nx_we[1] = std::min((int)xpoint2 + 1, (int)c.nxmin1);
  D.ny_sn[0] = std::max((int)ypoint1, 0);
  D.ny_sn[1] = std::min((int)ypoint2 + 1, (int)c.nymin1);
  D.gdomainfill = 0;
  if (c.xglobal && c.sglobal && c.nglobal)
    D.gdomainfill = (D.nx_we[0] == 0 && D.nx_we[1] == c.nxmin1 && D.ny_sn[0] == 0 && D.ny_sn[1] == c.nymin1) ? 1 : 0;
  if (c.xglobal) D.nx_we[1] = std::min(D.nx_we[1], (int)c.nx - 2);
  if (D.nx_we[1] < D.nx_we[0] || D.ny_sn[1] < D.ny_sn[0]) return fail("fpb_init_domainfill: empty domain box");

  DomainfillArgs a;
  per_step_cfg(h, a.cfg, 0, 0);
  a.p = h->p;
  a.A1 = h->A[0]; a.T1 = h->T[0];
  a.height = h->d_height;
  a.nx0 = D.nx_we[0]; a.nx1 = D.nx_we[1]; a.ny0 = D.ny_sn[0]; a.ny1 = D.ny_sn[1];
  a.ncolx = a.nx1 - a.nx0 + 1;
  a.ncols = a.ncolx * (a.ny1 - a.ny0 + 1);
  if (a.ncols > 1024 * 1024) return fail("fpb_init_domainfill: more than 2^20 columns");
  a.npart1 = (float)h->npart[0];
  a.itsplit = itsplit;
  a.id_stride = h->d.part_id_stride; a.id_offset = h->d.part_id_offset;
  a.uniforms = nullptr; a.u_off = nullptr;
  std::vector<float> ga(c.ny + 1, 0.f);
  domainfill_gridarea(c, D.ny_sn, ga);
  float *d_ga = nullptr, *d_colmass = nullptr, *d_total = nullptr, *d_uniforms = nullptr;
  int32_t *d_ncolumn = nullptr;
  unsigned *d_colstart = nullptr, *d_bsums = nullptr;
  unsigned long long *d_uoff = nullptr;
  int *d_out = nullptr;
  auto cleanup = [&]() {
    cudaFree(d_ga); cudaFree(d_colmass); cudaFree(d_total); cudaFree(d_uniforms); cudaFree(d_ncolumn);
    cudaFree(d_colstart); cudaFree(d_bsums); cudaFree(d_uoff); cudaFree(d_out);
  };
#define DFA(ptr, n) do { if (dalloc(&(ptr), (n))) { cleanup(); return 1; } } while (0)
#define DFK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { cleanup(); \
    return fail("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); } } while (0)
  DFA(d_ga, ga.size()); DFA(d_colmass, (size_t)a.ncols); DFA(d_total, 1); DFA(d_ncolumn, (size_t)a.ncols);
  DFA(d_colstart, (size_t)a.ncols); DFA(d_bsums, (size_t)(a.ncols + 1023) / 1024 + 1); DFA(d_out, 4);
  DFK(cudaMemcpyAsync(d_ga, ga.data(), ga.size() * sizeof(float), cudaMemcpyHostToDevice, h->stream));
  a.gridarea = d_ga; a.colmass = d_colmass; a.total = d_total; a.ncolumn = d_ncolumn; a.colstart = d_colstart;
  a.block_sums = d_bsums; a.out = d_out;
  // a fresh, unpermuted particle set
  sortk_iota(h->p.slot, c.maxpart, h->stream);
  sortk_iota(h->row_of_slot, c.maxpart, h->stream);
  fill_i32_kernel<<<(unsigned)((c.maxpart + 255) / 256), 256, 0, h->stream>>>(h->p.itra1, FPB_ITRA_DEAD, c.maxpart);
  h->launches += 3;
  h->permuted = false;
  fpb_domainfill_launch(a, h->stream, &h->launches, 0);
  int out[4] = {0, 0, 0, 0};
  float total = 0.f;
  DFK(cudaMemcpyAsync(out, d_out, sizeof out, cudaMemcpyDeviceToHost, h->stream));
  DFK(cudaMemcpyAsync(&total, d_total, sizeof total, cudaMemcpyDeviceToHost, h->stream));
  DFK(cudaGetLastError());
  DFK(cudaStreamSynchronize(h->stream));
  const long long numparttot = out[0];
  const long long mine = numparttot > a.id_offset ? (numparttot - a.id_offset + a.id_stride - 1) / a.id_stride : 0;
  if (mine > c.maxpart) {
    cleanup();
    return fail("fpb_init_domainfill: %lld particles for this rank exceed maxpart = %d "
                "(numpart too large: src/init_domainfill.f90:254-257)", mine, c.maxpart);
  }
  if (c.rng_mode == FPB_RNG_REFERENCE) {
    // the reference's ran1 stream in its draw order: columns jy-major, particles in order, per
    // particle [pnew when ncolumn <= 20] x [x again at ix = 0] [x again at ix = nxmin1] y class
    std::vector<int32_t> ncol((size_t)a.ncols);
    DFK(cudaMemcpy(ncol.data(), d_ncolumn, ncol.size() * sizeof(int32_t), cudaMemcpyDeviceToHost));
    std::vector<unsigned long long> uoff((size_t)a.ncols);
    unsigned long long pos = 0;
    for (int col = 0; col < a.ncols; col++) {
      const int ix = a.nx0 + col % a.ncolx;
      const int dpp = (ncol[col] > 20 ? 0 : 1) + 3 + (ix == 0 ? 1 : 0) + (ix == c.nxmin1 ? 1 : 0);
      uoff[col] = pos;
      pos += (unsigned long long)ncol[col] * dpp;
    }
    std::vector<float> u((size_t)pos);
    for (auto &v : u) v = D.ran1();
    DFA(d_uniforms, u.size()); DFA(d_uoff, uoff.size());
    DFK(cudaMemcpy(d_uniforms, u.data(), u.size() * sizeof(float), cudaMemcpyHostToDevice));
    DFK(cudaMemcpy(d_uoff, uoff.data(), uoff.size() * sizeof(unsigned long long), cudaMemcpyHostToDevice));
    a.uniforms = d_uniforms; a.u_off = d_uoff;
  }
  fpb_domainfill_launch(a, h->stream, &h->launches, 1);
  DFK(cudaMemcpyAsync(out, d_out, sizeof out, cudaMemcpyDeviceToHost, h->stream));
  DFK(cudaGetLastError());
  DFK(cudaStreamSynchronize(h->stream));
  D.itsplit = itsplit;
  D.numparticlecount = (int)numparttot;
  D.xmassperparticle = numparttot > 0 ? total / (float)numparttot : 0.f;
  if (!D.gdomainfill && domainfill_boundary_setup(h, a, d_colmass, total, out[1])) { cleanup(); return 1; }
#undef DFA
#undef DFK
  cleanup();
  if (c.rng_mode == FPB_RNG_REFERENCE && out[3] != 0)
    return fail("fpb_init_domainfill: %d particles found no or several pressure layers (non-monotonic rho*tt "
                "profile); the reference's ran1 draw order cannot be replayed for them", out[3]);
  h->numpart = out[2]; // :391-397: dead particles at the end of the arrays are dropped
  h->pending_init = true;
  h->active_rows = -1;
  D.done = true;
  if (numpart) *numpart = h->numpart;
  if (info) {
    info->nx_we[0] = D.nx_we[0]; info->nx_we[1] = D.nx_we[1];
    info->ny_sn[0] = D.ny_sn[0]; info->ny_sn[1] = D.ny_sn[1];
    info->gdomainfill = D.gdomainfill;
    info->numcolumn = out[1];
    info->numparttot = (int32_t)numparttot;
    info->colmasstotal = total;
    info->xmassperparticle = numparttot > 0 ? total / (float)numparttot : 0.f;
  }
  return 0;
}

extern "C" int fpb_boundcond_domainfill(fpb_handle *h, int32_t itime, int32_t loutend, int32_t *numpart,
                                        int32_t *n_created) {
  (void)loutend; // (the dump of the accumulated masses to boundcond.bin at loutend stays with the caller)
  if (!h) return fail("fpb_boundcond_domainfill: null handle");
  auto &D = h->dfill;
  if (!D.done) return fail("fpb_boundcond_domainfill: fpb_init_domainfill has not been called");
  if (numpart) *numpart = h->numpart;
  if (n_created) *n_created = 0;
  if (D.gdomainfill) return 0; // src/boundcond_domainfill.f90:54: nothing to do for a global domain
  if (!h->have_bracket) return fail("fpb_boundcond_domainfill: fpb_set_met_bracket has not been called");
  const fpb_config &c = h->cfg;
  CK(cudaSetDevice(h->device));
  BoundcondArgs a;
  per_step_cfg(h, a.cfg, itime, 0);
  a.p = h->p;
  a.row_of_slot = h->row_of_slot;
  a.permuted = h->permuted ? 1 : 0;
  a.numpart_old = h->numpart;
  a.A[0] = slot_view(h, h->memind[0]).A;
  a.A[1] = slot_view(h, h->memind[1]).A;
  a.dt1 = (float)(itime - h->memtime[0]);
  a.dt2 = (float)(h->memtime[1] - itime);
  a.dtt = 1.f / (a.dt1 + a.dt2);
  a.nx0 = D.nx_we[0]; a.nx1 = D.nx_we[1]; a.ny0 = D.ny_sn[0]; a.ny1 = D.ny_sn[1];
  a.check_x = (!c.xglobal || D.nx_we[1] != c.nx - 2) ? 1 : 0;
  a.loc = D.d_loc; a.nloc = D.nloc;
  a.acc_mass = D.d_acc; a.mmass = D.d_mmass; a.first = D.d_first;
  a.uniforms = nullptr; a.u_off = D.d_uoff;
  a.n_new = 0; a.n_mine = 0; a.r0 = 0;
  a.numparticlecount = D.numparticlecount;
  a.xmassperparticle = D.xmassperparticle;
  a.itsplit = D.itsplit;
  a.id_stride = h->d.part_id_stride;
  a.block_counts = D.d_blocks; a.out = D.d_out;
  fpb_boundcond_launch(a, h->stream, &h->launches, 0);
  std::vector<int32_t> mmass((size_t)D.nloc), first((size_t)D.nloc + 1, 0), uoff((size_t)D.nloc, 0);
  CK(cudaMemcpyAsync(mmass.data(), D.d_mmass, mmass.size() * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(h->stream));
  long long ndraws = 0;
  for (int l = 0; l < D.nloc; l++) {
    first[l + 1] = first[l] + mmass[l];
    uoff[l] = (int32_t)ndraws;
    ndraws += (long long)mmass[l] * D.loc_draws[l];
  }
  const int n_new = first[D.nloc];
  if (n_new == 0) return 0;
  const int stride = h->d.part_id_stride, offset = h->d.part_id_offset;
  a.n_new = n_new;
  a.r0 = (int)((((long long)offset - D.numparticlecount) % stride + stride) % stride);
  a.n_mine = n_new > a.r0 ? (n_new - a.r0 + stride - 1) / stride : 0;
  CK(cudaMemcpyAsync(D.d_first, first.data(), first.size() * sizeof(int32_t), cudaMemcpyHostToDevice, h->stream));
  std::vector<float> u;
  if (c.rng_mode == FPB_RNG_REFERENCE) { // the reference's ran1 stream of the call: [along] [height] class
    u.resize((size_t)ndraws);
    for (auto &v : u) v = D.ran1(D.idum_bc);
    if (D.uniforms_cap < u.size()) {
      cudaFree(D.d_uniforms); D.d_uniforms = nullptr;
      DA(D.d_uniforms, u.size());
      D.uniforms_cap = u.size();
    }
    CK(cudaMemcpyAsync(D.d_uniforms, u.data(), u.size() * sizeof(float), cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(D.d_uoff, uoff.data(), uoff.size() * sizeof(int32_t), cudaMemcpyHostToDevice, h->stream));
    a.uniforms = D.d_uniforms;
  }
  D.numparticlecount += n_new;
  if (a.n_mine > 0) {
    const int out0[2] = {h->numpart, 0};
    CK(cudaMemcpyAsync(D.d_out, out0, sizeof out0, cudaMemcpyHostToDevice, h->stream));
    fpb_boundcond_launch(a, h->stream, &h->launches, 1);
    int out[2];
    CK(cudaMemcpyAsync(out, D.d_out, sizeof out, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(h->stream));
    if (out[1] < a.n_mine)
      return fail("boundcond_domainfill: too many particles required (%d new, %d free slots, maxpart %d)", a.n_mine,
                  out[1], c.maxpart);
    h->numpart = out[0];
    h->pending_init = true;
    h->active_rows = -1;
  } else {
    CK(cudaStreamSynchronize(h->stream));
  }
  if (numpart) *numpart = h->numpart;
  if (n_created) *n_created = a.n_mine;
  return 0;
}

// ------------------------------------------------- calcpar + verttransform --
void fpb_metproc_launch(const fpbmet::MetGrid &g, cudaStream_t st, int64_t *launches); // fpb_metproc.cu

extern "C" int fpb_set_vertical(fpb_handle *h, int32_t nuvz, int32_t nwz, int32_t nuvzmax, int32_t nwzmax,
                                const float *akm, const float *bkm, const float *akz, const float *bkz) {
  if (!h || !akm || !bkm || !akz || !bkz) return fail("fpb_set_vertical: null argument");
  const fpb_config &c = h->cfg;
  if (nuvz != c.nz || nwz != nuvz)
    return fail("fpb_set_vertical: nuvz = %d, nwz = %d, nz = %d: only the ECMWF layout nz = nuvz = nwz "
                "(src/gridcheck_ecmwf.f90:434-436,516-526) is built", nuvz, nwz, c.nz);
  if (nuvzmax < nuvz || nwzmax < nwz) return fail("fpb_set_vertical: nuvzmax/nwzmax smaller than nuvz/nwz");
  CK(cudaSetDevice(h->device));
  auto &M = h->metproc;
  M.nuvz = nuvz; M.nwz = nwz; M.nuvzmax = nuvzmax; M.nwzmax = nwzmax;
  cudaFree(M.d_ab); M.d_ab = nullptr;
  cudaFree(M.d_cosf); M.d_cosf = nullptr;
  const size_t n1 = (size_t)nuvz + 1;
  DA(M.d_ab, 4 * n1);
  std::vector<float> ab(4 * n1, 0.f);
  const float *src[4] = {akz, bkz, akm, bkm};
  for (int q = 0; q < 4; q++)
    for (int k = 0; k < nuvz; k++) ab[q * n1 + k + 1] = src[q][k];
  CK(cudaMemcpy(M.d_ab, ab.data(), ab.size() * sizeof(float), cudaMemcpyHostToDevice));
  // cosf(jy) = 1./cos((real(jy)*dy+ylat0)*pi180), src/verttransform_ecmwf.f90:406-408
  std::vector<float> cosf((size_t)c.ny, 0.f);
  const float pi180 = 3.14159265f / 180.f;
  for (int jy = 0; jy < c.ny; jy++) cosf[jy] = 1.f / (float)cos((double)(((float)jy * c.dy + c.ylat0) * pi180));
  DA(M.d_cosf, cosf.size());
  CK(cudaMemcpy(M.d_cosf, cosf.data(), cosf.size() * sizeof(float), cudaMemcpyHostToDevice));
  if (!M.ev0) { CK(cudaEventCreate(&M.ev0)); CK(cudaEventCreate(&M.ev1)); CK(cudaEventCreate(&M.evk)); }
  return 0;
}

extern "C" int fpb_upload_vdep(fpb_handle *h, int32_t slot, const float *vdep) {
  if (!h || !vdep) return fail("fpb_upload_vdep: null argument");
  if (slot < 1 || slot > FPB_NSLOTS) return fail("fpb_upload_vdep: slot %d", slot);
  if (!h->cfg.drydep) return fail("fpb_upload_vdep: the run has no dry deposition (drydep = 0)");
  CK(cudaSetDevice(h->device));
  if (finish_met_upload(h)) return 1;
  if (alloc_met_slot(h, slot - 1)) return 1;
  const float *vd[1] = {vdep};
  if (upload_group(h, h->st_met, h->vdep[slot - 1], 1, vd, h->cfg.nspec)) return 1;
  CK(cudaStreamSynchronize(h->st_met));
  return 0;
}

extern "C" int fpb_calcpar_verttransform(fpb_handle *h, int32_t slot, const fpb_rawmet_ptrs *m, int32_t lsubgrid,
                                         float *device_ms) {
  if (!h || !m) return fail("fpb_calcpar_verttransform: null argument");
  auto &M = h->metproc;
  auto &V = h->conv;
  const fpb_config &c = h->cfg;
  if (M.nuvz == 0) return fail("fpb_calcpar_verttransform: fpb_set_vertical has not been called");
  if (slot < 1 || slot > FPB_NSLOTS) return fail("fpb_calcpar_verttransform: slot %d", slot);
  if (!m->uuh || !m->vvh || !m->tth || !m->qvh || !m->wwh || !m->ps || !m->tt2 || !m->td2 || !m->sshf || !m->surfstr)
    return fail("fpb_calcpar_verttransform: a mandatory field pointer is null");
  if (c.wetdep && (!m->lsprec || !m->convprec || !m->tcc))
    return fail("fpb_calcpar_verttransform: lsprec/convprec/tcc required when wetdep");
  const bool rdcl = c.wetdep && c.readclouds;
  if (rdcl && !m->clwch) return fail("fpb_calcpar_verttransform: clwch required when readclouds");
  if (lsubgrid == 1 && !m->excessoro) return fail("fpb_calcpar_verttransform: excessoro required when lsubgrid = 1");
  CK(cudaSetDevice(h->device));
  if (finish_met_upload(h)) return 1;
  const int s = slot - 1;
  if (alloc_met_slot(h, s)) return 1;
  const int nxd = h->d.nxd, nyd = h->d.nyd, nuvz = M.nuvz;
  const size_t n2 = (size_t)nxd * nyd, n3 = n2 * nuvz;
  if (!M.UV) { DA(M.UV, n3); DA(M.W, n3); DA(M.uvzlev, n3); DA(M.SF2, n2); }
  if (!M.PV) DA(M.PV, n3);
  if (!m->pvh && !M.theta) DA(M.theta, n3); // calcpv on the device
  if (lsubgrid == 1 && !M.excessoro) DA(M.excessoro, n2);
  if (rdcl && !M.CLW) { DA(M.CLW, n3); DA(M.clw, n3); }
  if (rdcl && m->ciwch && !M.CIW) DA(M.CIW, n3);
  if (!V.CT[0][s]) { DA(V.CT[0][s], n3); DA(V.CS[0][s], n2); }
  if (!h->outp.Q[s]) DA(h->outp.Q[s], n3);
  cudaStream_t st = h->st_met;
  CK(cudaEventRecord(M.ev0, st));
  {
    const float *uv[2] = {m->uuh, m->vvh}, *w1[1] = {m->wwh}, *tq[2] = {m->tth, m->qvh}, *pv[1] = {m->pvh};
    const float *s1[4] = {m->ps, m->tt2, m->td2, m->sshf}, *s2[4] = {m->surfstr, m->lsprec, m->convprec, m->tcc};
    const float *ex[1] = {m->excessoro};
    if (upload_group(h, st, (float *)M.UV, 2, uv, nuvz)) return 1;
    if (upload_group(h, st, M.W, 1, w1, M.nwz)) return 1;
    if (upload_group(h, st, (float *)V.CT[0][s], 2, tq, nuvz)) return 1;
    if (m->pvh && upload_group(h, st, M.PV, 1, pv, nuvz)) return 1;
    if (upload_group(h, st, (float *)V.CS[0][s], 4, s1, 1)) return 1;
    if (upload_group(h, st, (float *)M.SF2, 4, s2, 1)) return 1;
    if (lsubgrid == 1 && upload_group(h, st, M.excessoro, 1, ex, 1)) return 1;
    const float *cw[1] = {m->clwch}, *ci[1] = {m->ciwch};
    if (rdcl && upload_group(h, st, M.CLW, 1, cw, nuvz)) return 1;
    if (rdcl && m->ciwch && upload_group(h, st, M.CIW, 1, ci, nuvz)) return 1;
  }
  fpbmet::MetGrid g{};
  g.nx = c.nx; g.ny = c.ny; g.nz = c.nz; g.nuvz = nuvz; g.nwz = M.nwz;
  g.nxd = nxd; g.nyd = nyd;
  g.dx = c.dx; g.dy = c.dy; g.xlon0 = c.xlon0; g.ylat0 = c.ylat0; g.dxconst = c.dxconst; g.dyconst = c.dyconst;
  g.nglobal = c.nglobal; g.sglobal = c.sglobal; g.xglobal = c.xglobal;
  g.switchnorthg = c.switchnorthg; g.switchsouthg = c.switchsouthg;
  for (int k = 0; k < 9; k++) { g.northpolemap[k] = c.northpolemap[k]; g.southpolemap[k] = c.southpolemap[k]; }
  g.lsubgrid = lsubgrid; g.readclouds = rdcl ? 1 : 0;
  g.CLW = rdcl ? M.CLW : nullptr; g.CIW = (rdcl && m->ciwch) ? M.CIW : nullptr; g.clw = rdcl ? M.clw : nullptr;
  const size_t n1 = (size_t)nuvz + 1;
  g.akz = M.d_ab; g.bkz = M.d_ab + n1; g.akm = M.d_ab + 2 * n1; g.bkm = M.d_ab + 3 * n1;
  g.height = h->d_height; g.cosf = M.d_cosf;
  g.UV = M.UV; g.W = M.W; g.TQ = V.CT[0][s]; g.PV = M.PV; g.theta = m->pvh ? nullptr : M.theta;
  g.SF1 = V.CS[0][s]; g.SF2 = M.SF2; g.excessoro = M.excessoro; g.uvzlev = M.uvzlev;
  g.A = h->A[s]; g.G = h->G[s]; g.T = h->T[s]; g.P = h->P[s]; g.S = h->S[s]; g.trop = h->trop[s];
  g.R = c.wetdep ? h->R[s] : nullptr; g.Cl = c.wetdep ? h->Cl[s] : nullptr; g.Q = h->outp.Q[s];
  CK(cudaEventRecord(M.evk, st));
  fpb_metproc_launch(g, st, &h->launches);
  if (build_pairs(h, s, st)) return 1;
  CK(cudaEventRecord(M.ev1, st));
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(st));
  if (device_ms) {
    CK(cudaEventElapsedTime(device_ms, M.ev0, M.ev1));
    CK(cudaEventElapsedTime(device_ms + 1, M.evk, M.ev1));
  }
  h->slot_ready[s] = true;
  V.have[0][s] = (V.nuvz == nuvz);
  return 0;
}

// calcpar_nests + verttransform_nests (+ calcpv_nests) for nested input grid `nest` (src/getfields.f90:131-134):
// the same column code as the mother grid with the nest's geometry, no pole rows, no wrap in x
extern "C" int fpb_calcpar_verttransform_nest(fpb_handle *h, int32_t slot, int32_t nest, const fpb_rawmet_ptrs *m,
                                              int32_t lsubgrid, float dxn, float dyn, float xlon0n, float ylat0n,
                                              float *device_ms) {
  if (!h || !m) return fail("fpb_calcpar_verttransform_nest: null argument");
  auto &M = h->metproc;
  auto &V = h->conv;
  const fpb_config &c = h->cfg;
  if (M.nuvz == 0) return fail("fpb_calcpar_verttransform_nest: fpb_set_vertical has not been called");
  if (slot < 1 || slot > FPB_NSLOTS) return fail("fpb_calcpar_verttransform_nest: slot %d", slot);
  if (nest < 1 || nest > c.numbnests) return fail("fpb_calcpar_verttransform_nest: nest %d outside 1..numbnests=%d", nest, c.numbnests);
  if (!m->uuh || !m->vvh || !m->tth || !m->qvh || !m->wwh || !m->ps || !m->tt2 || !m->td2 || !m->sshf || !m->surfstr)
    return fail("fpb_calcpar_verttransform_nest: a mandatory field pointer is null");
  if (c.wetdep && (!m->lsprec || !m->convprec || !m->tcc))
    return fail("fpb_calcpar_verttransform_nest: lsprec/convprec/tcc required when wetdep");
  if (c.wetdep && c.readclouds_nest[nest - 1])
    return fail("fpb_calcpar_verttransform_nest: cloud water from the input (readclouds_nest) is not built");
  if (lsubgrid == 1 && !m->excessoro) return fail("fpb_calcpar_verttransform_nest: excessoro required when lsubgrid = 1");
  if (!(dxn > 0.f) || !(dyn > 0.f)) return fail("fpb_calcpar_verttransform_nest: dxn, dyn must be positive");
  CK(cudaSetDevice(h->device));
  if (finish_met_upload(h)) return 1;
  const int s = slot - 1, l = nest - 1;
  if (alloc_met_slot(h, s)) return 1;
  const int nxd = c.nxn[l], nyd = c.nyn[l], mx = c.nxmaxn, my = c.nymaxn, nuvz = M.nuvz;
  const size_t n2 = (size_t)nxd * nyd, n3 = n2 * nuvz;
  auto &W = M.nw[l];
  if (!W.UV) {
    DA(W.UV, n3); DA(W.W, n3); DA(W.uvzlev, n3); DA(W.SF2, n2); DA(W.PV, n3);
    // cosf(jy) = 1./cos((real(jy)*dyn(l)+ylat0n(l))*pi180), src/verttransform_nests.f90:283-286
    std::vector<float> cosf((size_t)nyd, 0.f);
    const float pi180 = 3.14159265f / 180.f;
    for (int jy = 0; jy < nyd; jy++) cosf[jy] = 1.f / (float)cos((double)(((float)jy * dyn + ylat0n) * pi180));
    DA(W.cosf, cosf.size());
    CK(cudaMemcpy(W.cosf, cosf.data(), cosf.size() * sizeof(float), cudaMemcpyHostToDevice));
  }
  if (!m->pvh && !W.theta) DA(W.theta, n3);
  if (lsubgrid == 1 && !W.excessoro) DA(W.excessoro, n2);
  if (!c.wetdep && !W.T) DA(W.T, n3); // (ttn is kept only for wet deposition, src/com_mod.f90:501-529)
  if (c.wetdep && !W.Q) DA(W.Q, n3);  // (qvn: read by the cloud classes)
  if (!V.CT[nest][s]) { DA(V.CT[nest][s], n3); DA(V.CS[nest][s], n2); }
  cudaStream_t st = h->st_met;
  CK(cudaEventRecord(M.ev0, st));
  {
    const float *uv[2] = {m->uuh, m->vvh}, *w1[1] = {m->wwh}, *tq[2] = {m->tth, m->qvh}, *pv[1] = {m->pvh};
    const float *s1[4] = {m->ps, m->tt2, m->td2, m->sshf}, *s2[4] = {m->surfstr, m->lsprec, m->convprec, m->tcc};
    const float *ex[1] = {m->excessoro};
    if (upload_group(h, st, (float *)W.UV, 2, uv, nuvz, nxd, nyd, mx, my)) return 1;
    if (upload_group(h, st, W.W, 1, w1, M.nwz, nxd, nyd, mx, my)) return 1;
    if (upload_group(h, st, (float *)V.CT[nest][s], 2, tq, nuvz, nxd, nyd, mx, my)) return 1;
    if (m->pvh && upload_group(h, st, W.PV, 1, pv, nuvz, nxd, nyd, mx, my)) return 1;
    if (upload_group(h, st, (float *)V.CS[nest][s], 4, s1, 1, nxd, nyd, mx, my)) return 1;
    if (upload_group(h, st, (float *)W.SF2, 4, s2, 1, nxd, nyd, mx, my)) return 1;
    if (lsubgrid == 1 && upload_group(h, st, W.excessoro, 1, ex, 1, nxd, nyd, mx, my)) return 1;
  }
  fpbmet::MetGrid g{};
  g.nx = nxd; g.ny = nyd; g.nz = c.nz; g.nuvz = nuvz; g.nwz = M.nwz;
  g.nxd = nxd; g.nyd = nyd;
  g.dx = dxn; g.dy = dyn; g.xlon0 = xlon0n; g.ylat0 = ylat0n; g.dxconst = c.dxconst; g.dyconst = c.dyconst;
  g.nest = 1; g.xresol = c.xresoln[l]; g.yresol = c.yresoln[l];
  g.lsubgrid = lsubgrid;
  const size_t n1 = (size_t)nuvz + 1;
  g.akz = M.d_ab; g.bkz = M.d_ab + n1; g.akm = M.d_ab + 2 * n1; g.bkm = M.d_ab + 3 * n1;
  g.height = h->d_height; g.cosf = W.cosf;
  g.UV = W.UV; g.W = W.W; g.TQ = V.CT[nest][s]; g.PV = W.PV; g.theta = m->pvh ? nullptr : W.theta;
  g.SF1 = V.CS[nest][s]; g.SF2 = W.SF2; g.excessoro = W.excessoro; g.uvzlev = W.uvzlev;
  g.A = h->An[l][s]; g.G = h->Gn[l][s]; g.T = c.wetdep ? h->Tn[l][s] : W.T; g.S = h->Sn[l][s]; g.trop = h->tropn[l][s];
  g.R = c.wetdep ? h->Rn[l][s] : nullptr; g.Cl = c.wetdep ? h->Cln[l][s] : nullptr; g.Q = c.wetdep ? W.Q : nullptr;
  CK(cudaEventRecord(M.evk, st));
  fpb_metproc_launch(g, st, &h->launches);
  CK(cudaEventRecord(M.ev1, st));
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(st));
  if (device_ms) {
    CK(cudaEventElapsedTime(device_ms, M.ev0, M.ev1));
    CK(cudaEventElapsedTime(device_ms + 1, M.evk, M.ev1));
  }
  V.have[nest][s] = (V.nuvz == nuvz);
  return 0;
}

static int fetch_met_impl(fpb_handle *h, int32_t slot, int32_t nest, const fpb_met_out_ptrs *o) {
  if (!h || !o) return fail("fpb_fetch_met: null argument");
  if (slot < 1 || slot > FPB_NSLOTS || !h->slot_ready[slot - 1]) return fail("fpb_fetch_met: slot %d holds no field", slot);
  const fpb_config &c = h->cfg;
  if (nest < 0 || nest > c.numbnests) return fail("fpb_fetch_met_nest: nest %d outside 1..numbnests=%d", nest, c.numbnests);
  CK(cudaSetDevice(h->device));
  if (finish_met_upload(h)) return 1;
  const int s = slot - 1, l = nest - 1, nz = c.nz;
  const int nxd = nest ? c.nxn[l] : h->d.nxd, nyd = nest ? c.nyn[l] : h->d.nyd;
  const int nxu = nest ? c.nxn[l] : c.nx, nyu = nest ? c.nyn[l] : c.ny, mx = nest ? c.nxmaxn : c.nxmax, my = nest ? c.nymaxn : c.nymax;
  const size_t n2 = (size_t)nxd * nyd, n3 = n2 * nz;
  std::vector<float> buf;
  // device [k][jy][ix][ncomp] -> host (nxmax, nymax, nk) of component `comp`
  auto fetch = [&](const void *dev, int ncomp, int nk, float *const *dst) -> int {
    if (!dev) return 0;
    bool any = false;
    for (int q = 0; q < ncomp; q++) any = any || dst[q];
    if (!any) return 0;
    buf.resize((size_t)nxd * nyd * nk * ncomp);
    CK(cudaMemcpy(buf.data(), dev, buf.size() * sizeof(float), cudaMemcpyDeviceToHost));
    for (int q = 0; q < ncomp; q++) {
      if (!dst[q]) continue;
      for (int k = 0; k < nk; k++)
        for (int jy = 0; jy < nyu; jy++)
          for (int ix = 0; ix < nxu; ix++)
            dst[q][(size_t)ix + (size_t)mx * ((size_t)jy + (size_t)my * k)] =
                buf[(((size_t)k * nyd + jy) * nxd + ix) * ncomp + q];
    }
    return 0;
  };
  float *a4[4] = {o->uu, o->vv, o->ww, o->rho}, *g1[1] = {o->drhodz}, *t1[1] = {o->tt}, *p2[2] = {o->uupol, o->vvpol};
  float *q2[2] = {o->pv, o->qv}, *s4[4] = {o->hmix, o->ustar, o->wstar, o->oli}, *tr[1] = {o->tropopause};
  float *r4[4] = {nullptr, nullptr, nullptr, o->ctwc};
  const int8_t *Cl;
  if (nest) {
    if (fetch(h->An[l][s], 4, nz, a4) || fetch(h->Gn[l][s], 1, nz, g1) || fetch(h->Tn[l][s], 1, nz, t1) ||
        fetch(h->Sn[l][s], 4, 1, s4) || fetch(h->tropn[l][s], 1, 1, tr))
      return 1;
    if (o->ctwc && h->Rn[l][s] && fetch(h->Rn[l][s], 4, 1, r4)) return 1;
    Cl = h->Cln[l][s];
  } else {
    if (fetch(h->A[s], 4, nz, a4) || fetch(h->G[s], 1, nz, g1) || fetch(h->T[s], 1, nz, t1) || fetch(h->P[s], 2, nz, p2) ||
        fetch(h->outp.Q[s], 2, nz, q2) || fetch(h->S[s], 4, 1, s4) || fetch(h->trop[s], 1, 1, tr))
      return 1;
    if (o->ctwc && h->R[s] && fetch(h->R[s], 4, 1, r4)) return 1;
    Cl = h->Cl[s];
  }
  if (o->clouds && Cl) {
    std::vector<int8_t> cb(n3);
    CK(cudaMemcpy(cb.data(), Cl, n3, cudaMemcpyDeviceToHost));
    for (int k = 0; k < nz; k++)
      for (int jy = 0; jy < nyu; jy++)
        for (int ix = 0; ix < nxu; ix++)
          o->clouds[(size_t)ix + (size_t)mx * ((size_t)jy + (size_t)my * k)] = cb[((size_t)k * nyd + jy) * nxd + ix];
  }
  return 0;
}
extern "C" int fpb_fetch_met(fpb_handle *h, int32_t slot, const fpb_met_out_ptrs *o) { return fetch_met_impl(h, slot, 0, o); }
extern "C" int fpb_fetch_met_nest(fpb_handle *h, int32_t slot, int32_t nest, const fpb_met_out_ptrs *o) {
  if (nest < 1) return fail("fpb_fetch_met_nest: nest %d", nest);
  return fetch_met_impl(h, slot, nest, o);
}

// ----------------------------------------------------------- convection --
extern "C" int fpb_set_convection(fpb_handle *h, int32_t nuvz, int32_t nuvzmax, int32_t nconvlev, const float *akz,
                                  const float *bkz, const float *akm, const float *bkm) {
  if (!h || !akz || !bkz || !akm || !bkm) return fail("fpb_set_convection: null argument");
  if (nuvz < 4 || nuvz > nuvzmax || nconvlev < 2 || nconvlev > nuvz - 2)
    return fail("fpb_set_convection: nuvz = %d (max %d), nconvlev = %d out of range", nuvz, nuvzmax, nconvlev);
  CK(cudaSetDevice(h->device));
  auto &V = h->conv;
  V.nuvz = nuvz; V.nuvzmax = nuvzmax; V.nconvlev = nconvlev;
  cudaFree(V.d_ab); V.d_ab = nullptr;
  const size_t n1 = (size_t)nuvz + 1;
  DA(V.d_ab, 4 * n1);
  std::vector<float> ab(4 * n1, 0.f);
  const float *src[4] = {akz, bkz, akm, bkm};
  for (int q = 0; q < 4; q++)
    for (int k = 0; k < nuvz; k++) ab[q * n1 + k + 1] = src[q][k];
  CK(cudaMemcpy(V.d_ab, ab.data(), ab.size() * sizeof(float), cudaMemcpyHostToDevice));
  for (int g = 0; g <= h->cfg.numbnests; g++) { // cbaseflux(n) = 0 at the start
    const size_t n2 = g ? (size_t)h->cfg.nxn[g - 1] * h->cfg.nyn[g - 1] : (size_t)h->d.nxd * h->d.nyd;
    if (!V.cbaseflux[g]) { DA(V.cbaseflux[g], n2); DA(V.cbase_bak[g], n2); }
  }
  return 0;
}

static int upload_convmet(fpb_handle *h, int32_t slot, int32_t nest, const fpb_conv_ptrs *m) {
  if (!h || !m) return fail("fpb_upload_convmet: null argument");
  auto &V = h->conv;
  const fpb_config &c = h->cfg;
  if (V.nuvz == 0) return fail("fpb_upload_convmet: fpb_set_convection has not been called");
  if (slot < 1 || slot > FPB_NSLOTS) return fail("fpb_upload_convmet: slot %d", slot);
  if (nest < 0 || nest > c.numbnests) return fail("fpb_upload_convmet: nest %d outside 0..numbnests=%d", nest, c.numbnests);
  if (!m->ps || !m->tt2 || !m->td2 || !m->tth || !m->qvh) return fail("fpb_upload_convmet: a field pointer is null");
  CK(cudaSetDevice(h->device));
  if (finish_met_upload(h)) return 1;
  const int s = slot - 1;
  const int nxd = nest ? c.nxn[nest - 1] : h->d.nxd, nyd = nest ? c.nyn[nest - 1] : h->d.nyd;
  const int mx = nest ? c.nxmaxn : c.nxmax, my = nest ? c.nymaxn : c.nymax;
  const size_t n2 = (size_t)nxd * nyd;
  if (!V.CT[nest][s]) { DA(V.CT[nest][s], n2 * V.nuvz); DA(V.CS[nest][s], n2); }
  // tth, qvh: (nxmax, nymax, nuvzmax) -> {tth, qvh}[k][jy][ix]: the first nuvz of the nuvzmax levels
  const float *q2[2] = {m->tth, m->qvh}, *s4[4] = {m->ps, m->tt2, m->td2, nullptr};
  if (upload_group(h, h->st_met, (float *)V.CT[nest][s], 2, q2, V.nuvz, nxd, nyd, mx, my)) return 1;
  if (upload_group(h, h->st_met, (float *)V.CS[nest][s], 4, s4, 1, nxd, nyd, mx, my)) return 1;
  CK(cudaStreamSynchronize(h->st_met));
  V.have[nest][s] = true;
  return 0;
}
extern "C" int fpb_upload_convmet(fpb_handle *h, int32_t slot, const fpb_conv_ptrs *m) {
  return upload_convmet(h, slot, 0, m);
}
extern "C" int fpb_upload_convmet_nest(fpb_handle *h, int32_t slot, int32_t nest, const fpb_conv_ptrs *m) {
  if (nest < 1) return fail("fpb_upload_convmet_nest: nest %d", nest);
  return upload_convmet(h, slot, nest, m);
}

// the reference's sort2 (src/sort2.f90, Numerical Recipes' quicksort of arr with brr alongside): the
// order in which the particles of a column -- and the columns -- are visited decides which particle
// gets which ran3 uniform, so the replay of the reference stream needs this very permutation
static void sort2_reference(int n, int32_t *arr, int32_t *brr) {
  const int M = 7, NSTACK = 50;
  int istack[NSTACK + 1];
  int jstack = 0, l = 1, ir = n, i, j, k;
  int32_t a, b;
  arr--; brr--; // 1-based
  for (;;) {
    if (ir - l < M) {
      for (j = l + 1; j <= ir; j++) {
        a = arr[j]; b = brr[j];
        for (i = j - 1; i >= 1; i--) {
          if (arr[i] <= a) break;
          arr[i + 1] = arr[i]; brr[i + 1] = brr[i];
        }
        arr[i + 1] = a; brr[i + 1] = b;
      }
      if (jstack == 0) return;
      ir = istack[jstack]; l = istack[jstack - 1]; jstack -= 2;
    } else {
      k = (l + ir) / 2;
      std::swap(arr[k], arr[l + 1]); std::swap(brr[k], brr[l + 1]);
      if (arr[l + 1] > arr[ir]) { std::swap(arr[l + 1], arr[ir]); std::swap(brr[l + 1], brr[ir]); }
      if (arr[l] > arr[ir]) { std::swap(arr[l], arr[ir]); std::swap(brr[l], brr[ir]); }
      if (arr[l + 1] > arr[l]) { std::swap(arr[l + 1], arr[l]); std::swap(brr[l + 1], brr[l]); }
      i = l + 1; j = ir;
      a = arr[l]; b = brr[l];
      for (;;) {
        do i++; while (arr[i] < a);
        do j--; while (arr[j] > a);
        if (j < i) break;
        std::swap(arr[i], arr[j]); std::swap(brr[i], brr[j]);
      }
      arr[l] = arr[j]; arr[j] = a;
      brr[l] = brr[j]; brr[j] = b;
      jstack += 2;
      if (jstack > NSTACK) return; // ('nstack too small in sort2': 2^25 elements and more)
      if (ir - i + 1 >= j - l) {
        istack[jstack] = ir; istack[jstack - 1] = i; ir = j - 1;
      } else {
        istack[jstack] = j - 1; istack[jstack - 1] = l; l = i;
      }
    }
  }
}

extern "C" int fpb_convmix(fpb_handle *h, int32_t itime, int32_t *ncolumns, int32_t *nconvecting) {
  if (!h) return fail("fpb_convmix: null handle");
  auto &V = h->conv;
  if (ncolumns) *ncolumns = 0;
  if (nconvecting) *nconvecting = 0;
  if (V.nuvz == 0) return fail("fpb_convmix: fpb_set_convection has not been called");
  if (!h->have_bracket) return fail("fpb_convmix: fpb_set_met_bracket has not been called");
  for (int g = 0; g <= h->cfg.numbnests; g++)
    if (!V.have[g][h->memind[0] - 1] || !V.have[g][h->memind[1] - 1])
      return fail("fpb_convmix: fpb_upload_convmet%s of both time levels of the bracket is missing (grid %d)", g ? "_nest" : "", g);
  if (h->numpart <= 0) return 0; // src/convmix.f90:75
  CK(cudaSetDevice(h->device));
  const fpb_config &c = h->cfg;
  const int n = h->active_rows >= 0 ? h->active_rows : h->numpart;
  if (n <= 0) return 0;
  const bool refrng = c.rng_mode == FPB_RNG_REFERENCE;
  // the column kernels want as many columns at once as possible (fpb_convect.cu): as many as a quarter of
  // the free device memory holds (80-140 KB of interleaved work slice + 130 KB of per-column matrices each at
  // 138 levels), at most 65536
  const size_t pf = fpb_convmix_pool_floats(V.nuvz, V.nconvlev);
  if (!V.pool) {
    size_t free_b = 0, total_b = 0;
    CK(cudaMemGetInfo(&free_b, &total_b));
    const size_t ld2 = fpb_convmix_pool2_floats(V.nconvlev); // + the contiguous MENT of the flux assembly and FMASS
    size_t cols = (free_b / 4) / ((pf + ld2) * sizeof(float));
    cols = std::max<size_t>(256, std::min<size_t>(cols, 65536)) / 32 * 32; // whole blocks of 32 interleaved slices
    DA(V.pool, pf * cols);
    DA(V.pool2, ld2 * cols);
    V.pool_cols = (int)cols;
  }
  if ((size_t)n > V.cap_rows) {
    cudaFree(V.block_counts); cudaFree(V.colidx); cudaFree(V.col_key); cudaFree(V.col_start); cudaFree(V.col_lconv);
    cudaFree(V.col_state);
    V.block_counts = nullptr; V.colidx = nullptr; V.col_key = nullptr; V.col_start = nullptr; V.col_lconv = nullptr;
    V.col_state = nullptr;
    DA(V.block_counts, (size_t)(n + 1023) / 1024 + 1); DA(V.colidx, (size_t)n);
    DA(V.col_key, (size_t)n + 1); DA(V.col_start, (size_t)n + 2); DA(V.col_lconv, (size_t)n + 1);
    DA(V.col_state, ((size_t)n + 1) * fpb_convmix_state_bytes());
    V.cap_rows = n;
  }
  if (!V.d_total) DA(V.d_total, 1);
  if (refrng && !V.key_by_slot) {
    DA(V.key_by_slot, (size_t)c.maxpart); DA(V.draws, (size_t)c.maxpart); DA(V.rn_by_slot, (size_t)c.maxpart);
  }
  if (scatter_reserve(V.sw, (size_t)n, 1)) return fail("%s", scatter_error());

  ConvmixArgs a;
  per_step_cfg(h, a.cfg, itime, 0);
  a.p = h->p;
  a.nrows = n;
  a.nuvz = V.nuvz; a.nconvlev = V.nconvlev;
  const size_t n1 = (size_t)V.nuvz + 1;
  a.akz = V.d_ab; a.bkz = V.d_ab + n1; a.akm = V.d_ab + 2 * n1; a.bkm = V.d_ab + 3 * n1;
  long long maxcol = (long long)c.nx * c.ny;
  for (int g = 0; g <= FPB_MAXNESTS; g++) {
    for (int m = 0; m < 2; m++) { a.CT[g][m] = V.CT[g][h->memind[m] - 1]; a.CS[g][m] = V.CS[g][h->memind[m] - 1]; }
    a.cbaseflux[g] = V.cbaseflux[g];
    a.gnx[g] = g ? c.nxn[g - 1] : c.nx;
    a.gnxd[g] = g ? c.nxn[g - 1] : h->d.nxd;
    a.gnyd[g] = g ? c.nyn[g - 1] : h->d.nyd;
    if (g >= 1 && g <= c.numbnests) maxcol = std::max(maxcol, (long long)c.nxn[g - 1] * c.nyn[g - 1]);
  }
  a.col_bits = 1;
  while ((1ll << a.col_bits) < maxcol + 1) a.col_bits++;
  a.ecmwf_eps = 1;
  a.ztop = h->height[c.nz - 1];
  a.keys = V.sw.keys[0]; a.ids = V.sw.ids[0];
  a.key_by_slot = refrng ? V.key_by_slot : nullptr;
  a.block_counts = V.block_counts; a.colidx = V.colidx; a.col_key = V.col_key; a.col_start = V.col_start;
  a.col_lconv = V.col_lconv; a.pool = V.pool; a.pool2 = V.pool2; a.col_state = V.col_state;
  a.draws = V.draws; a.rn_by_slot = nullptr; a.sorted_ids = nullptr;
  a.flux = (c.iflux == 1) ? h->hooks.flux : nullptr;
  if (refrng) CK(cudaMemsetAsync(V.key_by_slot, 0xff, (size_t)h->numpart * sizeof(int32_t), h->stream)); // -1
  fpb_convmix_keys(a, h->stream);
  int gbits = 0;
  while ((1 << gbits) < c.numbnests + 1) gbits++;
  int bits = ((a.col_bits + gbits + 1 + 7) / 8) * 8; // + the all-ones key of the rows that are not due
  if (bits > 32) bits = 32;
  int cur = 0;
  if (scatter_sort_pairs(V.sw, (size_t)n, bits, h->stream, &h->launches, &cur)) return fail("%s", scatter_error());
  a.sorted_ids = V.sw.ids[cur];
  fpb_convmix_heads(a, V.sw.keys[cur], V.d_total, h->stream);
  h->launches += 4;
  int ncols = 0;
  CK(cudaMemcpyAsync(&ncols, V.d_total, sizeof ncols, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(h->stream));
  if (ncolumns) *ncolumns = ncols;
  if (ncols == 0) return 0;
  std::vector<int32_t> col_start((size_t)ncols + 1);
  CK(cudaMemcpy(col_start.data(), V.col_start, col_start.size() * sizeof(int32_t), cudaMemcpyDeviceToHost));

  // reference RNG: which particles draw is only known once their column's matrix exists, and the
  // uniforms must be handed out in the reference's visiting order: pass 0 marks, the host replays
  // ran3 in sort2 order, pass 1 moves.  Philox: one pass.
  for (int pass = refrng ? 0 : 1; pass <= 1; pass++) {
    if (pass == 1 && refrng) {
      const int np = h->numpart;
      std::vector<int32_t> igrid(np), ipoint(np);
      std::vector<uint8_t> draws(np);
      std::vector<float> rn(np, 0.f);
      CK(cudaMemcpy(igrid.data(), V.key_by_slot, (size_t)np * sizeof(int32_t), cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(draws.data(), V.draws, (size_t)np, cudaMemcpyDeviceToHost));
      // the mother grid first, then every nest with its own sort2 (src/convmix.f90:150-196,203-281)
      std::vector<int32_t> ig(np);
      for (int g = 0; g <= c.numbnests; g++) {
        for (int i = 0; i < np; i++) {
          ipoint[i] = i;
          const int32_t kk = igrid[i];
          ig[i] = (kk != -1 && (kk >> a.col_bits) == g) ? (kk & ((1 << a.col_bits) - 1)) + 1 : -1;
        }
        sort2_reference(np, ig.data(), ipoint.data());
        for (int kq = 0; kq < np; kq++) {
          if (ig[kq] == -1) continue;
          const int ip = ipoint[kq];
          if (draws[ip]) rn[ip] = h->ran3.next(V.iseed);
        }
      }
      CK(cudaMemcpy(V.rn_by_slot, rn.data(), (size_t)np * sizeof(float), cudaMemcpyHostToDevice));
      a.rn_by_slot = V.rn_by_slot;
      // the columns are computed once more below: cbaseflux must not be advanced twice
      for (int g = 0; g <= c.numbnests; g++)
        CK(cudaMemcpyAsync(V.cbaseflux[g], V.cbase_bak[g], (size_t)a.gnxd[g] * a.gnyd[g] * sizeof(float),
                           cudaMemcpyDeviceToDevice, h->stream));
    }
    if (pass == 0) {
      CK(cudaMemsetAsync(V.draws, 0, (size_t)h->numpart, h->stream));
      for (int g = 0; g <= c.numbnests; g++)
        CK(cudaMemcpyAsync(V.cbase_bak[g], V.cbaseflux[g], (size_t)a.gnxd[g] * a.gnyd[g] * sizeof(float),
                           cudaMemcpyDeviceToDevice, h->stream));
    }
    for (int c0 = 0; c0 < ncols; c0 += V.pool_cols) {
      const int c1 = std::min(ncols, c0 + V.pool_cols);
      fpb_convmix_columns(a, c0, c1, h->stream);
      fpb_convmix_redist(a, c0, col_start[c0], col_start[c1], pass, h->stream);
      h->launches += 6; // levels, column head, level-pair rows, flux assembly, tail, redist
    }
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(h->stream));
  }
  if (nconvecting) {
    std::vector<int32_t> lc((size_t)ncols);
    CK(cudaMemcpy(lc.data(), V.col_lconv, lc.size() * sizeof(int32_t), cudaMemcpyDeviceToHost));
    int k = 0;
    for (int v : lc) k += v > 0;
    *nconvecting = k;
  }
  return 0;
}

// ------------------------------------------------------- host-buffer step --
static DevParticles rows_view(const DevParticles &p, int c0) {
  DevParticles v = p;
  v.xtra1 += c0; v.ytra1 += c0; v.ztra1 += c0;
  v.itra1 += c0; v.npoint += c0; v.nclass += c0; v.idt += c0; v.itramem += c0; v.itrasplit += c0;
  v.uap += c0; v.ucp += c0; v.uzp += c0; v.us += c0; v.vs += c0; v.ws += c0;
  v.cbt += c0;
  v.xmass1 += c0;      // species stride stays maxpart
  v.xscav_frac1 += c0;
  v.slot += c0;
  return v;
}

static DevScratch scratch_view(const DevScratch &s, int c0) {
  DevScratch v = s;
  v.flags += c0; v.s0 += c0; v.s1 += c0; v.s2 += c0;
  if (v.prob) v.prob += c0;
  return v;
}

static int ensure_lanes(fpb_handle *h) {
  if (h->lanes_ready) return 0;
  for (auto &L : h->lanes) {
    CK(cudaStreamCreateWithFlags(&L.st, cudaStreamNonBlocking));
    DA(L.d_work, 1);
    DA(L.d_nlive, 1);
  }
  CK(cudaEventCreateWithFlags(&h->ev_ready, cudaEventDisableTiming));
  CK(cudaStreamCreateWithFlags(&h->st_in, cudaStreamNonBlocking));
  for (auto &e : h->ev_in) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  for (auto &e : h->ev_det) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  for (auto &e : h->ev_dep) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  CK(cudaStreamCreateWithFlags(&h->st_out, cudaStreamNonBlocking));
  for (auto &e : h->ev_out) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  h->lanes_ready = true;
  return 0;
}

#define D2HS(dst, src, T)                                                                       \
  CK(cudaMemcpyAsync((dst) + c0, (src) + c0, (size_t)n * sizeof(T), cudaMemcpyDeviceToHost, post))

// One synchronisation interval for a host that keeps the particle arrays
// (include/fpb.h).  Rows are cut into chunks; chunk c runs H2D -> sort ->
// conccalc -> initialize -> particle loop -> D2H on lane c % NLANES, so the
// PCIe copies of one chunk overlap the kernels of the others.  Particles are
// independent within a step, so chunking changes nothing
// but the order of the float atomics into the grids.
static int step_host_impl(fpb_handle *h, int32_t itime, int32_t ldeltat, int32_t numpart, const fpb_particle_ptrs *p,
                          float conc_weight, fpb_step_stats *stats);
extern "C" int fpb_step_host(fpb_handle *h, int32_t itime, int32_t ldeltat, int32_t numpart,
                             const fpb_particle_ptrs *p, float conc_weight, fpb_step_stats *stats) {
  const int rc = step_host_impl(h, itime, ldeltat, numpart, p, conc_weight, stats);
  if (rc && h && h->lanes_ready) {
    // an error in the middle of the chunk loop leaves copies into the caller's arrays in flight on the
    // lanes: let them land before the caller may touch (or free) its buffers
    cudaSetDevice(h->device);
    for (auto &L : h->lanes) if (L.st) cudaStreamSynchronize(L.st);
    if (h->st_in) cudaStreamSynchronize(h->st_in);
    if (h->st_out) cudaStreamSynchronize(h->st_out);
    cudaStreamSynchronize(h->stream);
    cudaGetLastError();
  }
  return rc;
}
static int step_host_impl(fpb_handle *h, int32_t itime, int32_t ldeltat, int32_t numpart, const fpb_particle_ptrs *p,
                          float conc_weight, fpb_step_stats *stats) {
  if (!h || !p) return fail("fpb_step_host: null argument");
  if (!h->have_bracket) return fail("fpb_step_host: fpb_set_met_bracket has not been called");
  if (numpart < 0 || numpart > h->cfg.maxpart) return fail("fpb_step_host: numpart %d outside capacity %d", numpart, h->cfg.maxpart);
  if ((h->cfg.drybkdep || h->cfg.wetbkdep) && !p->xscav_frac1) return fail("fpb_step_host: xscav_frac1 is null");
  CK(cudaSetDevice(h->device));
  if (h->cfg.rng_mode != FPB_RNG_PHILOX && !h->d_rannumb) {
    if (fpb_fill_rannumb(h, 1000000, -320)) return 1;
  }
  if (stats) memset(stats, 0, sizeof *stats);
  if (numpart == 0) return 0;
  if (ensure_lanes(h)) return 1;
  const fpb_config &c = h->cfg;
  const bool strict = c.math_mode == FPB_MATH_STRICT;

  // staging rows are in slot order
  sortk_iota(h->p_alt.slot, numpart, h->stream);
  h->launches++;
  if (c.rng_mode == FPB_RNG_REFERENCE) {
    // the reference's ran3 draws in particle order, from the host arrays (see replay_ran3_indices)
    if (!p->itra1 || !p->itramem) return fail("fpb_step_host: itra1/itramem are null");
    h->h_nrand_init.assign(numpart, 1); h->h_nrand_adv.assign(numpart, 1);
    const float scale = (float)(h->maxrand - 1);
    for (int s = 0; s < numpart; s++) {
      if (p->itra1[s] != itime) continue;
      if (p->itramem[s] == itime || itime == 0)
        h->h_nrand_init[s] = (int)(h->ran3.next(h->idummy_init) * scale) + 1;
      h->h_nrand_adv[s] = (int)(h->ran3.next(h->idummy_adv) * scale) + 1;
    }
    CK(cudaMemcpyAsync(h->d_nrand_init, h->h_nrand_init.data(), (size_t)numpart * sizeof(int32_t), cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(h->d_nrand_adv, h->h_nrand_adv.data(), (size_t)numpart * sizeof(int32_t), cudaMemcpyHostToDevice, h->stream));
  }
  CK(cudaMemsetAsync(h->d_stats, 0, 8 * sizeof(unsigned long long), h->stream));
  CK(cudaEventRecord(h->ev_ready, h->stream));

  const bool regime = true;
  int bits = ((sortk_key_bits(h->d, regime) + 7) / 8) * 8;
  if (bits > 32) bits = 32;
  DevMetSlot met[2] = {slot_view(h, h->memind[0]), slot_view(h, h->memind[1])};

  // Equal row chunks of >= ~250k rows (multiples of 128 rows).  Each chunk costs ~20 launches and
  // one tail of the persistent sub-step kernel (~0.18 ms, FPB_HOST_TIMING=1 shows the timeline), so
  // there are few of them: 4 at 1M rows, worse beyond 6.
  // (Round 2, built, measured and removed again -- commit cd7639e has it, DESIGN.md section 5: ONE persistent sub-step
  // launch that consumes the chunks as their copies land.  3.4-3.9 ms per step against 2.96 chunk by chunk: what a
  // chunk costs after its copy has landed is the chain of its slowest particle, not the kernel's tail.)
  const bool dbg = getenv("FPB_HOST_DEBUG") != nullptr;
  std::vector<int> bounds; // chunk c = rows [bounds[c], bounds[c+1])
  {
    int nchunk = numpart / 250000;
    nchunk = nchunk < 1 ? 1 : (nchunk > 6 ? 6 : nchunk);
    if (const char *e = getenv("FPB_HOST_CHUNKS")) { // tuning knob
      const int v = atoi(e);
      if (v >= 1 && v <= fpb_handle::MAXCHUNKS) nchunk = v;
    }
    const int per_eq = (((numpart + nchunk - 1) / nchunk) + 127) / 128 * 128;
    // From four chunks on the first ones are larger and the last one holds 0.6 of an equal share: what is left to do when the last copy has landed
    // is shorter (with the copy-out deferred behind the last upload, below: -2.8 %, profiles/ab_r02_hostplan.txt).
    // FPB_HOST_PLAN (tuning knob): chunk sizes as fractions, e.g. "0.3,0.3,0.25,0.15"; "equal": equal chunks.
    const char *plan = getenv("FPB_HOST_PLAN");
    if (plan && !strcmp(plan, "equal")) plan = nullptr;
    else if (!plan && !getenv("FPB_HOST_CHUNKS") && nchunk >= 4) plan = "auto";
    if (plan) {
      std::vector<double> fr;
      if (!strcmp(plan, "auto")) {
        fr.assign((size_t)nchunk, 1.0); // 4 chunks: 0.30, 0.30, 0.25, 0.15
        for (int k = 0; k < nchunk / 2; k++) fr[k] = 1.2;
        fr.back() = 0.6;
      } else {
        for (const char *q = plan; *q;) {
          char *end = nullptr;
          const double v = strtod(q, &end);
          if (end == q) break;
          if (v > 0.) fr.push_back(v);
          q = (*end == ',') ? end + 1 : end;
        }
      }
      if (fr.empty() || (int)fr.size() > fpb_handle::MAXCHUNKS) fr = {0.5, 0.5};
      double tot = 0., acc = 0.;
      for (double v : fr) tot += v;
      bounds.push_back(0);
      for (size_t k = 0; k + 1 < fr.size(); k++) {
        acc += fr[k] / tot;
        const int b = (int)(acc * numpart) / 128 * 128;
        if (b > bounds.back() && b < numpart) bounds.push_back(b);
      }
      bounds.push_back(numpart);
    } else {
      for (int c0 = 0; c0 < numpart; c0 += per_eq) bounds.push_back(c0);
      bounds.push_back(numpart);
    }
  }
  // The persistent sub-step grid of a chunk takes 0.6 of a resident wave, so that the next chunk's
  // kernels start under its draining tail (measured at 1 M rows, gpurun_out/ab_host.txt:
  // 3 chunks x full grid 3.17 ms, 4 x 0.6: 2.95 ms, 4 x 0.4: 2.97, 6 x 0.6: 3.03, 8 x 0.6: 3.21)
  float host_grid_frac = bounds.size() > 2 ? 0.6f : 0.f;
  if (const char *e = getenv("FPB_HOST_GRID_FRAC")) host_grid_frac = (float)atof(e); // tuning knob
  // FPB_HOST_TIMING=1: per-chunk timeline (ms since the call started) on stderr
  const bool timing = getenv("FPB_HOST_TIMING") != nullptr;
  std::vector<cudaEvent_t> tev;
  const auto t_host0 = std::chrono::steady_clock::now();
  auto mark = [&](cudaStream_t st) {
    if (!timing) return;
    cudaEvent_t e;
    cudaEventCreate(&e);
    cudaEventRecord(e, st);
    tev.push_back(e);
  };
  mark(h->stream);
  // FPB_HOST_DEBUG=1: synchronise after every stage so that a device fault is attributed to it
#define STAGE(name)                                                                              \
  do {                                                                                           \
    if (dbg) {                                                                                   \
      cudaError_t e_ = cudaStreamSynchronize(L.st);                                              \
      if (e_ != cudaSuccess) return fail("fpb_step_host: %s in stage %s of chunk %d (rows %d+%d)", \
                                         cudaGetErrorString(e_), name, ci, c0, n);               \
    }                                                                                            \
  } while (0)
  int per = 0; // largest chunk: size of the lanes' sort work areas
  for (size_t k = 0; k + 1 < bounds.size(); k++) per = std::max(per, bounds[k + 1] - bounds[k]);

  // Every chunk's copy back to the host waits for the LAST chunk's upload, on one stream of its own: the two directions
  // then do not share the link while the uploads (the critical path) run.  FPB_HOST_DEFER_D2H=0 (tuning knob): each
  // chunk copies out as soon as it is done.
  // (Not with several ranks on one host: their copies share the host's PCIe root ports and memory anyway, and bunching
  // the copy-out at the end of the step makes that worse -- 7.5 against 6.6-7.0 ms per step and rank at 8 ranks.)
  const char *defer_env = getenv("FPB_HOST_DEFER_D2H");
  const bool defer_d2h = !dbg && !timing && bounds.size() > 2 &&
                         (defer_env ? atoi(defer_env) != 0 : h->comm.nranks <= 1);
  auto d2h_rows = [&](int c0, int n, cudaStream_t post) -> int {
    D2HS(p->xtra1, h->p_alt.xtra1, double); D2HS(p->ytra1, h->p_alt.ytra1, double);
    D2HS(p->ztra1, h->p_alt.ztra1, float); D2HS(p->itra1, h->p_alt.itra1, int32_t);
    D2HS(p->idt, h->p_alt.idt, int32_t);
    D2HS(p->uap, h->p_alt.uap, float); D2HS(p->ucp, h->p_alt.ucp, float); D2HS(p->uzp, h->p_alt.uzp, float);
    D2HS(p->us, h->p_alt.us, float); D2HS(p->vs, h->p_alt.vs, float); D2HS(p->ws, h->p_alt.ws, float);
    D2HS(p->cbt, h->p_alt.cbt, int16_t);
    for (int k = 0; k < c.nspec; k++) {
      CK(cudaMemcpyAsync(p->xmass1 + (size_t)k * p->ld + c0, h->p_alt.xmass1 + (size_t)k * c.maxpart + c0,
                         (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, post));
      if (c.drybkdep || c.wetbkdep) // set once after the release by the receptor block of the loop
        CK(cudaMemcpyAsync(p->xscav_frac1 + (size_t)k * p->ld + c0, h->p_alt.xscav_frac1 + (size_t)k * c.maxpart + c0,
                           (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, post));
    }
    return 0;
  };
  auto fill_step_args = [&](DevStepArgs &a, int c0, int n) {
    per_step_cfg(h, a.cfg, itime, ldeltat);
    a.cfg.numpart = n;
    a.met[0] = met[0]; a.met[1] = met[1];
    a.met_lit1 = slot_view(h, 1);
    nest_views(h, a);
    a.p = rows_view(h->p, c0);
    a.height = h->d_height;
    a.rannumb = h->d_rannumb;
    a.npart = h->d_npart;
    a.xmass = h->d_xmass;
    a.nrand_init = h->d_nrand_init; // indexed by (global) slot
    a.nrand_adv = h->d_nrand_adv;
    a.drygridunc = h->drygridunc;
    a.drygriduncn = h->drygriduncn;
    a.stats = h->d_stats;
    a.sc = scratch_view(h->sc, c0);
    a.dep = DevDepRecords{};
    a.grid_frac = host_grid_frac;
  };
  for (int ci = 0; ci + 1 < (int)bounds.size(); ci++) {
    const int c0 = bounds[ci], n = bounds[ci + 1] - bounds[ci];
    if (n <= 0) continue;
    fpb_handle::Lane &L = h->lanes[ci % fpb_handle::NLANES];
    if (scatter_reserve(L.sw, (size_t)per, 1)) return fail("%s", scatter_error());
    // all uploads go through one stream, in chunk order: chunk 0 is complete (and its kernels
    // start) after a third of the transfer instead of sharing the link with the later chunks
    if (ci == 0) CK(cudaStreamWaitEvent(h->st_in, h->ev_ready, 0));
    mark(h->st_in);
    if (copy_rows_h2d(h, h->p_alt, c0, n, p, h->st_in, true)) return 1;
    mark(h->st_in);
    cudaEvent_t ev_in = h->ev_in[ci % 8];
    CK(cudaEventRecord(ev_in, h->st_in));
    CK(cudaStreamWaitEvent(L.st, h->ev_ready, 0));
    CK(cudaStreamWaitEvent(L.st, ev_in, 0));
    STAGE("h2d");
    const DevParticles stg = rows_view(h->p_alt, c0), rows = rows_view(h->p, c0);

    DevStepArgs a;
    per_step_cfg(h, a.cfg, itime, ldeltat);
    a.cfg.numpart = n;
    sortk_build_keys(a.cfg, stg, h->d_height, n, L.sw.keys[0], L.sw.ids[0], L.d_nlive, L.st,
                     regime ? met : nullptr);
    int cur = 0;
    if (scatter_sort_pairs(L.sw, (size_t)n, bits, L.st, &h->launches, &cur)) return fail("%s", scatter_error());
    sortk_permute(stg, rows, L.sw.ids[cur], n, c.nspec, L.st, h->row_of_slot, c0);
    h->launches += 2;
    STAGE("sort");

    if (conc_weight > 0.f) {
      DevConcArgs q;
      q.cfg = a.cfg;
      q.cfg.weight = conc_weight;
      q.met[0] = met[0]; q.met[1] = met[1];
      q.p = rows;
      q.height = h->d_height;
      q.gridunc = h->gridunc; q.griduncn = h->griduncn; q.crec_acc = h->crec_acc;
      q.slot_base = c0;
      q.rec_vals = nullptr; q.rec_nslots = 0;
      const bool det = c.scatter_mode == FPB_SCATTER_DETERMINISTIC;
      if (det) {
        // every cell must receive its contributions in slot order: the chunks hold ascending slot
        // ranges, so chunk ci adds after chunk ci-1 has (an event chain across the lanes)
        if (ci > 0) CK(cudaStreamWaitEvent(L.st, h->ev_det[(ci - 1) % 8], 0));
        if (scatter_conccalc_deterministic(L.sw, q, strict, L.st, &h->launches)) return fail("%s", scatter_error());
      } else {
        if (strict) fpbk_conccalc_strict(q, L.st); else fpbk_conccalc_fast(q, L.st);
        h->launches++;
      }
      if (c.numreceptor > 0) {
        if (det && rec_begin(h, L.dep, n, L.st, q)) return 1;
        if (strict) fpbk_receptor_strict(q, L.st); else fpbk_receptor_fast(q, L.st);
        h->launches++;
        if (det) { // the running sums continue from the previous chunk's
          if (scatter_receptor_ordered(q.rec_vals, c.numreceptor * c.nspec, n, h->crec_acc, L.st))
            return fail("%s", scatter_error());
          h->launches++;
        }
      }
      if (det) CK(cudaEventRecord(h->ev_det[ci % 8], L.st));
    }

    STAGE("conccalc");
    fill_step_args(a, c0, n);
    a.work_counter = L.d_work;
    const bool det_dry = c.scatter_mode == FPB_SCATTER_DETERMINISTIC && c.drydep;
    if (det_dry && dep_begin(h, L.dep, n, c0, L.st, a.dep)) return 1;
    if (strict) fpbk_init_strict(a, L.st); else fpbk_init_fast(a, L.st);
    if (launch_bkdep(h, a.cfg, rows, L.st)) return 1;
    STAGE("initialize");
    cudaStream_t post = L.st; // the stream the rest of the chunk runs on
    HookArgs hk;
    if (hooks_on(h)) {
      if (hooks_args(h, hk, a.cfg, rows, c0)) return 1;
      fpb_hooks_pre(hk, L.st);
    }
    if (strict) fpbk_step_strict(a, L.st); else fpbk_step_fast(a, L.st);
    if (hooks_on(h)) {
      fpb_hooks_post(hk, L.st);
      h->launches += 2;
    }
    h->launches += 3;
    STAGE("step");
    if (det_dry) { // deposition in slot order across the chunks, like the concentration grid
      if (ci > 0) CK(cudaStreamWaitEvent(L.st, h->ev_dep[(ci - 1) % 8], 0));
      if (dep_apply(h, L.sw, a.dep, h->drygridunc, h->drygriduncn, L.st)) return 1;
      CK(cudaEventRecord(h->ev_dep[ci % 8], L.st));
      STAGE("drydep records");
    }

    sortk_scatter_back(rows, h->p_alt, n, c.nspec, post, c.drybkdep || c.wetbkdep);
    h->launches++;
    STAGE("scatter_back");
    if (defer_d2h) { // (copied out after the last upload, below)
      CK(cudaEventRecord(h->ev_out[ci], post));
      CK(cudaGetLastError());
      continue;
    }
    mark(post);
    if (d2h_rows(c0, n, post)) return 1;
    mark(post);
    CK(cudaGetLastError());
  }
#undef STAGE
  if (defer_d2h) {
    const int nch = (int)bounds.size() - 1;
    CK(cudaStreamWaitEvent(h->st_out, h->ev_in[(nch - 1) % 8], 0));
    for (int ci = 0; ci < nch; ci++) {
      CK(cudaStreamWaitEvent(h->st_out, h->ev_out[ci], 0));
      if (d2h_rows(bounds[ci], bounds[ci + 1] - bounds[ci], h->st_out)) return 1;
    }
    CK(cudaStreamSynchronize(h->st_out));
  }
  const double host_submit_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_host0).count();
  for (auto &L : h->lanes) CK(cudaStreamSynchronize(L.st));
  if (timing) {
    for (size_t k = 1; k + 3 < tev.size() + 1; k += 4) {
      float t[4];
      for (int q = 0; q < 4; q++) cudaEventElapsedTime(&t[q], tev[0], tev[k + q]);
      fprintf(stderr, "fpb_step_host chunk %zu (%d rows): h2d %.3f-%.3f  kernels -%.3f  d2h -%.3f ms\n", (k - 1) / 4,
              bounds[(k - 1) / 4 + 1] - bounds[(k - 1) / 4], t[0], t[1], t[2], t[3]);
    }
    fprintf(stderr, "fpb_step_host: all chunks submitted after %.3f ms of host time\n", host_submit_ms);
    for (auto e : tev) cudaEventDestroy(e);
  }
  if (conc_weight > 0.f && c.numreceptor > 0) {
    DevCfg d;
    per_step_cfg(h, d, itime, ldeltat);
    receptor_finalize_kernel<<<1, 256, 0, h->stream>>>(h->creceptor, h->crec_acc, c.numreceptor, c.nspec,
                                                      conc_weight, nullptr, d);
    h->launches++;
  }
  unsigned long long hs[8];
  CK(cudaMemcpyAsync(hs, h->d_stats, sizeof hs, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  if (stats) {
    stats->n_active = (int64_t)hs[0]; stats->n_init = (int64_t)hs[1]; stats->n_terminated = (int64_t)hs[2];
    stats->n_pbl = (int64_t)hs[3]; stats->n_substeps = (int64_t)hs[4]; stats->n_petterssen = (int64_t)hs[5];
    stats->n_nan_cbl = (int64_t)hs[6];
    stats->n_nonfinite = (int64_t)hs[7];
  }
  // the device rows stay valid (sorted inside each chunk) for resident-mode calls
  h->numpart = numpart;
  h->permuted = true;
  h->active_rows = -1;
  h->pending_init = false;
  h->steps_since_sort = 0;
  return 0;
}

// ---------------------------------------------------------- wet deposition --
extern "C" int fpb_wetdepo(fpb_handle *h, int32_t itime, int32_t ltsample, int32_t ldeltat) {
  if (!h) return fail("fpb_wetdepo: null handle");
  if (!h->cfg.wetdep) return fail("fpb_wetdepo: the run was initialised without wet deposition (wetdep = 0)");
  if (!h->have_bracket) return fail("fpb_wetdepo: fpb_set_met_bracket has not been called");
  if (h->numpart == 0) return 0;
  CK(cudaSetDevice(h->device));
  DevWetArgs a;
  per_step_cfg(h, a.cfg, itime, ldeltat);
  // time level closest to itime - ltsample/2, src/get_wetscav.f90:113-117
  const int interp_time = (int)lroundf((float)itime - 0.5f * (float)ltsample);
  int n = h->memind[1];
  if (abs(h->memtime[0] - interp_time) < abs(h->memtime[1] - interp_time)) n = h->memind[0];
  a.met = slot_view(h, n);
  for (int l = 0; l < FPB_MAXNESTS; l++) {
    DevMetSlot v{};
    v.R = h->Rn[l][n - 1]; v.C = h->Cln[l][n - 1]; v.T = h->Tn[l][n - 1];
    a.metn[l] = v;
  }
  a.p = h->p;
  a.height = h->d_height;
  a.wetgridunc = h->wetgridunc;
  a.wetgriduncn = h->wetgriduncn;
  a.ltsample = ltsample;
  a.dep = DevDepRecords{};
  const bool det = h->cfg.scatter_mode == FPB_SCATTER_DETERMINISTIC;
  if (det && dep_begin(h, h->depstore, h->numpart, 0, h->stream, a.dep)) return 1;
  if (h->cfg.math_mode == FPB_MATH_STRICT) fpbk_wetdepo_strict(a, h->stream);
  else fpbk_wetdepo_fast(a, h->stream);
  h->launches++;
  if (det && dep_apply(h, h->scatter, a.dep, h->wetgridunc, h->wetgriduncn, h->stream)) return 1;
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}

extern "C" int fpb_kernel_times(fpb_handle *h, float *step_ms, float *conccalc_ms) {
  if (!h) return fail("fpb_kernel_times: null handle");
  CK(cudaSetDevice(h->device));
  if (step_ms) {
    *step_ms = 0.f;
    if (h->timed_step) {
      CK(cudaEventSynchronize(h->ev[1]));
      CK(cudaEventElapsedTime(step_ms, h->ev[0], h->ev[1]));
    }
  }
  if (conccalc_ms) {
    *conccalc_ms = 0.f;
    if (h->timed_conc) {
      CK(cudaEventSynchronize(h->ev[3]));
      CK(cudaEventElapsedTime(conccalc_ms, h->ev[2], h->ev[3]));
    }
  }
  return 0;
}

extern "C" int fpb_get_rannumb(fpb_handle *h, float *out, int32_t n) {
  if (!h || !out) return fail("fpb_get_rannumb: null argument");
  if (!h->d_rannumb || n > h->maxrand) return fail("fpb_get_rannumb: table has %d entries", h->maxrand);
  CK(cudaSetDevice(h->device));
  CK(cudaMemcpy(out, h->d_rannumb, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost));
  return 0;
}

// ------------------------------------------------------------------ grids --
// device layout: species extent nspec; host layout: species extent maxspec
static int fetch_one(fpb_handle *h, float *host, const float *dev, size_t inner, size_t ndev,
                     std::vector<float> &tmp) {
  if (!host || !dev) return 0;
  const fpb_config &c = h->cfg;
  tmp.resize(ndev);
  CK(cudaMemcpyAsync(tmp.data(), dev, ndev * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  const size_t outer = (size_t)c.maxpointspec_act * c.nclassunc * c.maxageclass;
  for (size_t o = 0; o < outer; o++) {
    memcpy(host + o * c.maxspec * inner, tmp.data() + o * c.nspec * inner, (size_t)c.nspec * inner * sizeof(float));
    if (c.maxspec > c.nspec)
      memset(host + (o * c.maxspec + c.nspec) * inner, 0, (size_t)(c.maxspec - c.nspec) * inner * sizeof(float));
  }
  return 0;
}

extern "C" int fpb_zero_conc_grids(fpb_handle *h) {
  if (!h) return fail("fpb_zero_conc_grids: null handle");
  CK(cudaSetDevice(h->device));
  CK(cudaMemsetAsync(h->gridunc, 0, h->n_grid * sizeof(float), h->stream));
  if (h->griduncn) CK(cudaMemsetAsync(h->griduncn, 0, h->n_gridn * sizeof(float), h->stream));
  CK(cudaMemsetAsync(h->creceptor, 0, h->n_rec * sizeof(float), h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}

extern "C" int fpb_fetch_grids(fpb_handle *h, float *gridunc, float *griduncn, float *drygridunc,
                               float *drygriduncn, float *creceptor, int32_t zero_conc) {
  if (!h) return fail("fpb_fetch_grids: null handle");
  const fpb_config &c = h->cfg;
  if (c.maxspec < c.nspec) return fail("fpb_fetch_grids: maxspec < nspec");
  CK(cudaSetDevice(h->device));
  std::vector<float> tmp;
  if (fetch_one(h, gridunc, h->gridunc, (size_t)c.numxgrid * c.numygrid * c.numzgrid, h->n_grid, tmp)) return 1;
  if (fetch_one(h, drygridunc, h->drygridunc, (size_t)c.numxgrid * c.numygrid, h->n_dry, tmp)) return 1;
  if (c.nested_output == 1) {
    if (fetch_one(h, griduncn, h->griduncn, (size_t)c.numxgridn * c.numygridn * c.numzgrid, h->n_gridn, tmp)) return 1;
    if (fetch_one(h, drygriduncn, h->drygriduncn, (size_t)c.numxgridn * c.numygridn, h->n_dryn, tmp)) return 1;
  }
  if (creceptor) {
    tmp.resize(h->n_rec);
    CK(cudaMemcpyAsync(tmp.data(), h->creceptor, h->n_rec * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    memset(creceptor, 0, (size_t)FPB_MAXRECEPTOR * c.maxspec * sizeof(float));
    memcpy(creceptor, tmp.data(), h->n_rec * sizeof(float)); // (maxreceptor, ks) with ks < nspec
  }
  if (zero_conc) return fpb_zero_conc_grids(h);
  return 0;
}

extern "C" int fpb_fetch_wetgrids(fpb_handle *h, float *wetgridunc, float *wetgriduncn) {
  if (!h) return fail("fpb_fetch_wetgrids: null handle");
  const fpb_config &c = h->cfg;
  if (!c.wetdep) return fail("fpb_fetch_wetgrids: the run was initialised without wet deposition");
  CK(cudaSetDevice(h->device));
  std::vector<float> tmp;
  if (fetch_one(h, wetgridunc, h->wetgridunc, (size_t)c.numxgrid * c.numygrid, h->n_dry, tmp)) return 1;
  if (c.nested_output == 1 && wetgriduncn)
    if (fetch_one(h, wetgriduncn, h->wetgriduncn, (size_t)c.numxgridn * c.numygridn, h->n_dryn, tmp)) return 1;
  return 0;
}

__global__ void scale_dep_kernel(float *g, size_t n, size_t inner, int nspec, DevCfg c, const float f0,
                                 const float f1, const float f2, const float f3, const float f4,
                                 const float f5, const float f6, const float f7) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int ks = (int)((i / inner) % nspec);
  const float f[8] = {f0, f1, f2, f3, f4, f5, f6, f7};
  g[i] *= f[ks];
}

extern "C" int fpb_scale_depgrids(fpb_handle *h, const float *factor) {
  if (!h || !factor) return fail("fpb_scale_depgrids: null argument");
  CK(cudaSetDevice(h->device));
  float f[8] = {1, 1, 1, 1, 1, 1, 1, 1};
  for (int k = 0; k < h->cfg.nspec; k++) f[k] = factor[k];
  const fpb_config &c = h->cfg;
  scale_dep_kernel<<<(unsigned)((h->n_dry + 255) / 256), 256, 0, h->stream>>>(
      h->drygridunc, h->n_dry, (size_t)c.numxgrid * c.numygrid, c.nspec, h->d, f[0], f[1], f[2], f[3], f[4], f[5], f[6], f[7]);
  h->launches++;
  if (h->drygriduncn) {
    scale_dep_kernel<<<(unsigned)((h->n_dryn + 255) / 256), 256, 0, h->stream>>>(
        h->drygriduncn, h->n_dryn, (size_t)c.numxgridn * c.numygridn, c.nspec, h->d, f[0], f[1], f[2], f[3], f[4], f[5], f[6], f[7]);
    h->launches++;
  }
  if (h->wetgridunc) { // wetgridunc decays with drygridunc, src/timemanager.f90:276-282
    scale_dep_kernel<<<(unsigned)((h->n_dry + 255) / 256), 256, 0, h->stream>>>(
        h->wetgridunc, h->n_dry, (size_t)c.numxgrid * c.numygrid, c.nspec, h->d, f[0], f[1], f[2], f[3], f[4], f[5], f[6], f[7]);
    h->launches++;
  }
  if (h->wetgriduncn) {
    scale_dep_kernel<<<(unsigned)((h->n_dryn + 255) / 256), 256, 0, h->stream>>>(
        h->wetgriduncn, h->n_dryn, (size_t)c.numxgridn * c.numygridn, c.nspec, h->d, f[0], f[1], f[2], f[3], f[4], f[5], f[6], f[7]);
    h->launches++;
  }
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}

extern "C" int fpb_grid_device_ptr(fpb_handle *h, int32_t which, void **dptr, size_t *nfloats) {
  if (!h || !dptr || !nfloats) return fail("fpb_grid_device_ptr: null argument");
  switch (which) {
    case 0: *dptr = h->gridunc; *nfloats = h->n_grid; break;
    case 1: *dptr = h->griduncn; *nfloats = h->n_gridn; break;
    case 2: *dptr = h->drygridunc; *nfloats = h->n_dry; break;
    case 3: *dptr = h->drygriduncn; *nfloats = h->n_dryn; break;
    case 4: *dptr = h->creceptor; *nfloats = h->n_rec; break;
    default: return fail("fpb_grid_device_ptr: which=%d", which);
  }
  return 0;
}


extern "C" void *fpb_stream(fpb_handle *h) { return h ? (void *)h->stream : nullptr; }
extern "C" int64_t fpb_launch_count(fpb_handle *h) { return h ? h->launches : 0; }


// ------------------------------------------------------------ collective --
// The one collective of the path: the sum of the accumulation grids over the ranks at each output
// interval (mpif_tm_reduce_grid, src/mpi_mod.f90:2395-2579, called at src/timemanager_mpi.f90:468-485),
// as an NCCL reduce over NVLink.  NCCL is bound at run time (dlopen of libnccl.so.2, or the copy the
// host process has already loaded), so libfpb.so has no link-time dependency on it and single-GPU
// hosts need no NCCL at all.
namespace {
struct NcclUid { char b[128]; }; // ncclUniqueId (NCCL_UNIQUE_ID_BYTES = 128), passed by value
struct NcclApi {
  void *lib = nullptr;
  int (*GetUniqueId)(void *) = nullptr;
  int (*CommInitRank)(void **, int, NcclUid, int) = nullptr;
  int (*CommDestroy)(void *) = nullptr;
  int (*Reduce)(const void *, void *, size_t, int, int, int, void *, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(int) = nullptr;
  int (*GetVersion)(int *) = nullptr;
  bool ok = false;
} g_nccl;
constexpr int NCCL_FLOAT32 = 7, NCCL_SUM = 0; // ncclFloat32, ncclSum (nccl.h)

int nccl_load() {
  if (g_nccl.ok) return 0;
  void *lib = nullptr;
  if (dlsym(RTLD_DEFAULT, "ncclCommInitRank")) lib = dlopen(nullptr, RTLD_NOW); // already in the process
  if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!lib) return fail("fpb_comm: libnccl.so.2 not found (%s); the multi-GPU grid reduce needs NCCL", dlerror());
  g_nccl.lib = lib;
#define NSYM(field, name)                                                          \
  *(void **)(&g_nccl.field) = dlsym(lib, name);                                    \
  if (!g_nccl.field) return fail("fpb_comm: symbol %s missing in the NCCL library", name);
  NSYM(GetUniqueId, "ncclGetUniqueId") NSYM(CommInitRank, "ncclCommInitRank") NSYM(CommDestroy, "ncclCommDestroy")
  NSYM(Reduce, "ncclReduce") NSYM(GroupStart, "ncclGroupStart") NSYM(GroupEnd, "ncclGroupEnd")
  NSYM(GetErrorString, "ncclGetErrorString") NSYM(GetVersion, "ncclGetVersion")
#undef NSYM
  g_nccl.ok = true;
  return 0;
}
#define NK(call)                                                                              \
  do {                                                                                        \
    int r_ = (call);                                                                          \
    if (r_ != 0) return fail("%s failed: %s", #call, g_nccl.GetErrorString(r_));              \
  } while (0)
} // namespace

// worker: records the start event, enqueues the reduce group on the side stream, records the end event
static void comm_worker(fpb_handle *h) {
  auto &Q = h->comm;
  cudaSetDevice(h->device);
  for (;;) {
    std::unique_lock<std::mutex> lk(Q.mu);
    Q.cv.wait(lk, [&] { return Q.job || Q.quit; });
    if (Q.quit) return;
    Q.job = false;
    lk.unlock();
    int rc = 0;
    std::string err;
    auto nk = [&](int r, const char *what) {
      if (r != 0 && rc == 0) { rc = 1; err = std::string(what) + " failed: " + g_nccl.GetErrorString(r); }
    };
    if (cudaEventRecord(Q.ev_t0, Q.side) != cudaSuccess) { rc = 1; err = "cudaEventRecord failed"; }
    nk(g_nccl.GroupStart(), "ncclGroupStart");
    for (int k = 0; k < 7; k++)
      if (Q.stage_n[k])
        nk(g_nccl.Reduce(Q.stage[k], Q.stage[k], Q.stage_n[k], NCCL_FLOAT32, NCCL_SUM, 0, Q.nccl, Q.side), "ncclReduce");
    nk(g_nccl.GroupEnd(), "ncclGroupEnd");
    if (cudaEventRecord(Q.ev_done, Q.side) != cudaSuccess && rc == 0) { rc = 1; err = "cudaEventRecord failed"; }
    lk.lock();
    Q.job_rc = rc; Q.job_err = err; Q.job_done = true;
    lk.unlock();
    Q.cv.notify_all();
  }
}
// the posted job has been enqueued (its events are recorded); returns its status
static int comm_wait_enqueued(fpb_handle *h) {
  auto &Q = h->comm;
  std::unique_lock<std::mutex> lk(Q.mu);
  Q.cv.wait(lk, [&] { return Q.job_done; });
  if (Q.job_rc) {
    Q.job_rc = 0;
    return fail("fpb_reduce_grids: %s", Q.job_err.c_str());
  }
  return 0;
}

extern "C" int fpb_comm_unique_id(void *id128) {
  if (!id128) return fail("fpb_comm_unique_id: null argument");
  if (nccl_load()) return 1;
  NK(g_nccl.GetUniqueId(id128));
  return 0;
}

extern "C" int fpb_comm_init(fpb_handle *h, const void *id128, int32_t rank, int32_t nranks) {
  if (!h || !id128) return fail("fpb_comm_init: null argument");
  if (nranks < 1 || rank < 0 || rank >= nranks) return fail("fpb_comm_init: rank %d of %d", rank, nranks);
  if (h->comm.nccl) return fail("fpb_comm_init: the handle already has a communicator");
  CK(cudaSetDevice(h->device));
  auto &Q = h->comm;
  Q.rank = rank; Q.nranks = nranks;
  if (nranks > 1) {
    if (nccl_load()) return 1;
    NcclUid uid;
    memcpy(uid.b, id128, sizeof uid.b);
    NK(g_nccl.CommInitRank(&Q.nccl, nranks, uid, rank));
    Q.worker = std::thread(comm_worker, h);
  }
  int lo = 0, hi = 0;
  CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
  CK(cudaStreamCreateWithPriority(&Q.side, cudaStreamNonBlocking, hi)); // ahead of the step kernels
  CK(cudaEventCreateWithFlags(&Q.ev_staged, cudaEventDisableTiming));
  CK(cudaEventCreate(&Q.ev_t0));
  CK(cudaEventCreate(&Q.ev_done));
  // staging copies of the grids: the interval's sums leave through them while the next interval
  // accumulates into the (zeroed) grids
  const float *src[7] = {h->gridunc, h->griduncn, h->drygridunc, h->drygriduncn, h->wetgridunc, h->wetgriduncn,
                         h->cfg.numreceptor > 0 ? h->creceptor : nullptr};
  const size_t n[7] = {h->n_grid, h->n_gridn, h->cfg.drydep ? h->n_dry : 0, h->cfg.drydep ? h->n_dryn : 0,
                       h->n_dry, h->n_dryn, h->n_rec};  // (grids nothing is ever added to stay at home)
  for (int k = 0; k < 7; k++) {
    Q.stage_n[k] = src[k] ? n[k] : 0;
    if (Q.stage_n[k]) DA(Q.stage[k], Q.stage_n[k]);
  }
  return 0;
}

// Start the exchange of the interval that just ended: copy the grids to the staging buffers and
// zero the concentration grids (src/concoutput.f90:719-720) on the engine's stream, then sum the
// staging buffers to rank 0 on a high-priority side stream.  Returns at once; the engine may step on.
extern "C" int fpb_reduce_grids_begin(fpb_handle *h) {
  if (!h) return fail("fpb_reduce_grids_begin: null handle");
  auto &Q = h->comm;
  if (!Q.side) return fail("fpb_reduce_grids_begin: fpb_comm_init has not been called");
  CK(cudaSetDevice(h->device));
  if (comm_wait_enqueued(h)) return 1;
  if (Q.in_flight) CK(cudaStreamWaitEvent(h->stream, Q.ev_done, 0)); // the staging buffers are free again
  const float *src[7] = {h->gridunc, h->griduncn, h->drygridunc, h->drygriduncn, h->wetgridunc, h->wetgriduncn,
                         h->creceptor};
  for (int k = 0; k < 7; k++)
    if (Q.stage_n[k])
      CK(cudaMemcpyAsync(Q.stage[k], src[k], Q.stage_n[k] * sizeof(float), cudaMemcpyDeviceToDevice, h->stream));
  CK(cudaMemsetAsync(h->gridunc, 0, h->n_grid * sizeof(float), h->stream));
  if (h->griduncn) CK(cudaMemsetAsync(h->griduncn, 0, h->n_gridn * sizeof(float), h->stream));
  CK(cudaMemsetAsync(h->creceptor, 0, h->n_rec * sizeof(float), h->stream));
  CK(cudaEventRecord(Q.ev_staged, h->stream));
  CK(cudaStreamWaitEvent(Q.side, Q.ev_staged, 0));
  if (Q.nranks > 1) {
    { // post the enqueue of the reduce group to the worker
      std::lock_guard<std::mutex> lk(Q.mu);
      Q.job = true;
      Q.job_done = false;
    }
    Q.cv.notify_all();
    h->launches++;
  } else {
    CK(cudaEventRecord(Q.ev_t0, Q.side));
    CK(cudaEventRecord(Q.ev_done, Q.side));
  }
  Q.in_flight = true;
  Q.timed = true;
  return 0;
}

// Finish the exchange: wait for the sums; on rank 0 copy them into the caller's arrays (reference
// layout, like fpb_fetch_grids / fpb_fetch_wetgrids; NULL pointers are skipped).  Other ranks pass NULLs.
extern "C" int fpb_reduce_grids_end(fpb_handle *h, float *gridunc, float *griduncn, float *drygridunc,
                                    float *drygriduncn, float *wetgridunc, float *wetgriduncn, float *creceptor) {
  if (!h) return fail("fpb_reduce_grids_end: null handle");
  auto &Q = h->comm;
  if (!Q.in_flight) return fail("fpb_reduce_grids_end: no exchange in flight");
  CK(cudaSetDevice(h->device));
  if (comm_wait_enqueued(h)) return 1;
  CK(cudaEventSynchronize(Q.ev_done));
  Q.in_flight = false;
  if (Q.rank != 0) return 0;
  const fpb_config &c = h->cfg;
  std::vector<float> tmp;
  float *dst[6] = {gridunc, griduncn, drygridunc, drygriduncn, wetgridunc, wetgriduncn};
  const size_t inner[6] = {(size_t)c.numxgrid * c.numygrid * c.numzgrid, (size_t)c.numxgridn * c.numygridn * c.numzgrid,
                           (size_t)c.numxgrid * c.numygrid, (size_t)c.numxgridn * c.numygridn,
                           (size_t)c.numxgrid * c.numygrid, (size_t)c.numxgridn * c.numygridn};
  for (int k = 0; k < 6; k++)
    if (dst[k] && Q.stage_n[k] && fetch_one(h, dst[k], Q.stage[k], inner[k], Q.stage_n[k], tmp)) return 1;
  if (creceptor && Q.stage_n[6]) { // (maxreceptor, maxspec), species beyond nspec zero
    tmp.resize(h->n_rec);
    CK(cudaMemcpyAsync(tmp.data(), Q.stage[6], h->n_rec * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    memset(creceptor, 0, (size_t)FPB_MAXRECEPTOR * c.maxspec * sizeof(float));
    memcpy(creceptor, tmp.data(), h->n_rec * sizeof(float));
  }
  return 0;
}

// device view of a reduced staging buffer (valid on rank 0 between _begin and the next _begin, after
// the engine's stream or the host has waited for the exchange) and the duration of the last reduce
extern "C" int fpb_reduce_grids_device(fpb_handle *h, int32_t which, void **dptr, size_t *nfloats, float *reduce_ms) {
  if (!h) return fail("fpb_reduce_grids_device: null handle");
  auto &Q = h->comm;
  if (which < 0 || which > 6) return fail("fpb_reduce_grids_device: which=%d", which);
  if (!Q.side) return fail("fpb_reduce_grids_device: fpb_comm_init has not been called");
  CK(cudaSetDevice(h->device));
  if (comm_wait_enqueued(h)) return 1;
  if (Q.in_flight) CK(cudaEventSynchronize(Q.ev_done));
  if (dptr) *dptr = Q.stage[which];
  if (nfloats) *nfloats = Q.stage_n[which];
  if (reduce_ms) {
    *reduce_ms = 0.f;
    if (Q.timed) CK(cudaEventElapsedTime(reduce_ms, Q.ev_t0, Q.ev_done));
  }
  return 0;
}

extern "C" int fpb_comm_finalize(fpb_handle *h) {
  if (!h) return 0;
  auto &Q = h->comm;
  cudaSetDevice(h->device);
  if (Q.worker.joinable()) {
    comm_wait_enqueued(h);
    { std::lock_guard<std::mutex> lk(Q.mu); Q.quit = true; }
    Q.cv.notify_all();
    Q.worker.join();
  }
  if (Q.side) { cudaStreamSynchronize(Q.side); }
  if (Q.nccl) { g_nccl.CommDestroy(Q.nccl); Q.nccl = nullptr; }
  for (auto &p : Q.stage) { cudaFree(p); p = nullptr; }
  if (Q.ev_staged) cudaEventDestroy(Q.ev_staged);
  if (Q.ev_t0) cudaEventDestroy(Q.ev_t0);
  if (Q.ev_done) cudaEventDestroy(Q.ev_done);
  if (Q.side) cudaStreamDestroy(Q.side);
  Q.side = nullptr; Q.ev_staged = Q.ev_t0 = Q.ev_done = nullptr;
  Q.in_flight = Q.timed = false; Q.quit = false; Q.job = false; Q.job_done = true;
  for (auto &v : Q.stage_n) v = 0;
  return 0;
}
