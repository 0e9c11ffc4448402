// fpb_output.cu -- see fpb_output.cuh.  Compiled with --fmad=false: the sums of mean_sp and the
// unit conversion keep the reference's operation order and rounding.
//
// The reference (src/concoutput.f90:287-475) walks the cells of one (ks, kp, nage) grid in
// storage order.  A cell is written when its value exceeds tiny(0.0); the first cell of every run
// of such cells also records its linear index, and the values of a run carry the sign
// (-1)**(run number - 1): `sp_fact` flips at each run start.  Both lists are a stream compaction:
// two block-level scans (non-zero cells, run starts) give every cell its place and its sign.
#include "fpb_output.cuh"

namespace {
constexpr int OUT_BLOCK = 1024;
constexpr float SMALLNUM = 1.17549435e-38f; // tiny(0.0), src/concoutput.f90:83

// mean_sp over the uncertainty classes times nclassunc (src/mean_mod.f90:20-74,
// src/concoutput.f90:326-331); the standard deviation is not part of the dump
__device__ __forceinline__ float cell_total(const SparseDumpArgs &a, int i) {
  float xl = 0.f;
  for (int l = 0; l < a.nclassunc; l++) xl = xl + a.grid[(size_t)l * a.class_stride + i];
  const float xm = xl / (float)a.nclassunc;
  return xm * (float)a.nclassunc;
}

// magnitude written to sparse_dump_r (the sign is the run's)
__device__ __forceinline__ float cell_value(const SparseDumpArgs &a, int i, float g) {
  if (a.which == 0) { // src/concoutput.f90:225-235,451-454
    const float factor3d = (a.ldirect == 1) ? 1.e12f / a.geom[i] / a.outnum : a.loutaver_abs / a.outnum;
    return g * factor3d / a.tot_mu;
  }
  if (a.which == 3) // mixing ratio, :579-583 (tot_mu carries weightmolar(ks); weightair = 28.97)
    return 1.e12f * g / a.geom[i] / a.outnum * 28.97f / a.tot_mu / a.density[i];
  return 1.e12f * g / a.geom[i]; // :370-372, :402-405
}

// exclusive scan over a 1024-thread block of two counters at once
__device__ __forceinline__ void block_scan2(unsigned va, unsigned vb, unsigned &ea, unsigned &eb,
                                            unsigned &ta, unsigned &tb) {
  __shared__ unsigned wa[32], wb[32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  unsigned ia = va, ib = vb;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const unsigned xa = __shfl_up_sync(0xffffffffu, ia, d), xb = __shfl_up_sync(0xffffffffu, ib, d);
    if (lane >= d) { ia += xa; ib += xb; }
  }
  if (lane == 31) { wa[w] = ia; wb[w] = ib; }
  __syncthreads();
  if (w == 0) {
    unsigned sa = wa[lane], sb = wb[lane], ja = sa, jb = sb;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const unsigned xa = __shfl_up_sync(0xffffffffu, ja, d), xb = __shfl_up_sync(0xffffffffu, jb, d);
      if (lane >= d) { ja += xa; jb += xb; }
    }
    wa[lane] = ja - sa; wb[lane] = jb - sb;
  }
  __syncthreads();
  ea = wa[w] + ia - va;
  eb = wb[w] + ib - vb;
  __shared__ unsigned tot[2];
  if (threadIdx.x == OUT_BLOCK - 1) { tot[0] = ea + va; tot[1] = eb + vb; }
  __syncthreads();
  ta = tot[0]; tb = tot[1];
  __syncthreads();
}

__device__ __forceinline__ void cell_flags(const SparseDumpArgs &a, int i, float &g, unsigned &nz, unsigned &rs) {
  nz = 0; rs = 0; g = 0.f;
  if (i < a.ncells) {
    g = cell_total(a, i);
    nz = g > SMALLNUM;
    if (nz) rs = (i == 0) || !(cell_total(a, i - 1) > SMALLNUM);
  }
}

__global__ void __launch_bounds__(OUT_BLOCK) sparse_count_kernel(const SparseDumpArgs a) {
  const int i = blockIdx.x * OUT_BLOCK + threadIdx.x;
  float g;
  unsigned nz, rs;
  cell_flags(a, i, g, nz, rs);
  const int cn = __syncthreads_count(nz), cr = __syncthreads_count(rs);
  if (threadIdx.x == 0) { a.block_counts[2 * blockIdx.x] = cn; a.block_counts[2 * blockIdx.x + 1] = cr; }
}

__global__ void __launch_bounds__(OUT_BLOCK) sparse_scan_kernel(const SparseDumpArgs a, int nblocks) {
  __shared__ unsigned run[2];
  if (threadIdx.x == 0) { run[0] = 0; run[1] = 0; }
  __syncthreads();
  for (int base = 0; base < nblocks; base += OUT_BLOCK) {
    const int b = base + threadIdx.x;
    const unsigned va = (b < nblocks) ? a.block_counts[2 * b] : 0u, vb = (b < nblocks) ? a.block_counts[2 * b + 1] : 0u;
    unsigned ea, eb, ta, tb;
    block_scan2(va, vb, ea, eb, ta, tb);
    if (b < nblocks) { a.block_counts[2 * b] = run[0] + ea; a.block_counts[2 * b + 1] = run[1] + eb; }
    __syncthreads();
    if (threadIdx.x == 0) { run[0] += ta; run[1] += tb; }
    __syncthreads();
  }
  if (threadIdx.x == 0) { a.counts[1] = (int)run[0]; a.counts[0] = (int)run[1]; }
}

__global__ void __launch_bounds__(OUT_BLOCK) sparse_write_kernel(const SparseDumpArgs a) {
  const int i = blockIdx.x * OUT_BLOCK + threadIdx.x;
  float g;
  unsigned nz, rs;
  cell_flags(a, i, g, nz, rs);
  unsigned en, er, tn, tr;
  block_scan2(nz, rs, en, er, tn, tr);
  if (!nz) return;
  const unsigned kn = a.block_counts[2 * blockIdx.x] + en;          // place in sparse_dump_r
  const unsigned kr = a.block_counts[2 * blockIdx.x + 1] + er + rs; // number of this cell's run (1-based)
  if (rs) a.out_i[kr - 1] = i + a.index_offset;
  const float v = cell_value(a, i, g);
  a.out_r[kn] = (kr & 1u) ? v : -v;
}
__global__ void __launch_bounds__(256) density_outgrid_kernel(const DensityArgs a) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int n2 = a.numx * a.numy;
  if (i >= n2 * a.numz) return;
  const int kz = i / n2, jy = (i - kz * n2) / a.numx, ix = i - kz * n2 - jy * a.numx;
  float xl = a.outlon0 + (float)ix * a.dxout;
  float yl = a.outlat0 + (float)jy * a.dyout;
  xl = (xl - a.xlon0) / a.dx;
  yl = (yl - a.ylat0) / a.dy;
  const int iix = max(min((int)roundf(xl), a.nxmin1), 0); // nint(): half away from zero
  const int jjy = max(min((int)roundf(yl), a.nymin1), 0);
  const int kzz = a.kzz[kz];
  const float dz1 = a.dz1[kz], dz2 = a.dz2[kz], dz = dz1 + dz2;
  const float r1 = a.A[(size_t)(kzz - 1) * a.plane + (size_t)jjy * a.nxd + iix].w;
  const float r0 = a.A[(size_t)(kzz - 2) * a.plane + (size_t)jjy * a.nxd + iix].w;
  a.density[i] = (r1 * dz1 + r0 * dz2) / dz;
}
// ---- partoutput
__device__ __forceinline__ bool partout_active(const PartoutArgs &a, int s, int &row) {
  if (s >= a.numpart) return false;
  row = a.permuted ? a.row_of_slot[s] : s;
  return a.p.itra1[row] == a.cfg.itime;
}

__global__ void __launch_bounds__(OUT_BLOCK) partout_count_kernel(const PartoutArgs a) {
  int row;
  const int n = __syncthreads_count(partout_active(a, blockIdx.x * OUT_BLOCK + threadIdx.x, row));
  if (threadIdx.x == 0) { a.block_counts[2 * blockIdx.x] = n; a.block_counts[2 * blockIdx.x + 1] = 0; }
}

__global__ void __launch_bounds__(OUT_BLOCK) partout_scan_kernel(const PartoutArgs a, int nblocks) {
  __shared__ unsigned run;
  if (threadIdx.x == 0) run = 0;
  __syncthreads();
  for (int base = 0; base < nblocks; base += OUT_BLOCK) {
    const int b = base + threadIdx.x;
    const unsigned v = (b < nblocks) ? a.block_counts[2 * b] : 0u;
    unsigned ea, eb, ta, tb;
    block_scan2(v, 0u, ea, eb, ta, tb);
    if (b < nblocks) a.block_counts[2 * b] = run + ea;
    __syncthreads();
    if (threadIdx.x == 0) run += ta;
    __syncthreads();
  }
  if (threadIdx.x == 0) a.count[0] = (int)run;
}

__global__ void __launch_bounds__(OUT_BLOCK) partout_write_kernel(const PartoutArgs a) {
  const DevCfg &c = a.cfg;
  const int s = blockIdx.x * OUT_BLOCK + threadIdx.x;
  int row = 0;
  const unsigned act = partout_active(a, s, row);
  unsigned e, e2, t, t2;
  block_scan2(act, 0u, e, e2, t, t2);
  if (!act) return;
  const unsigned k = a.block_counts[2 * blockIdx.x] + e;
  const double xt = a.p.xtra1[row], yt = a.p.ytra1[row];
  const float zt = a.p.ztra1[row];
  a.xlon[k] = (float)(c.xlon0 + xt * c.dx);   // src/partoutput.f90:72-73
  a.ylat[k] = (float)(c.ylat0 + yt * c.dy);
  const int ix = (int)xt, jy = (int)yt;
  const int ixp = ix + 1;
  int jyp = jy + 1;
  const float ddx = (float)(xt - (float)ix), ddy = (float)(yt - (float)jy);
  const float rddx = 1.f - ddx, rddy = 1.f - ddy;
  const float p1 = rddx * rddy, p2 = ddx * rddy, p3 = rddx * ddy, p4 = ddx * ddy;
  if (jyp >= c.nymax) jyp = jyp - 1;
  const int o00 = ix + c.nxd * jy, o10 = ixp + c.nxd * jy, o01 = ix + c.nxd * jyp, o11 = ixp + c.nxd * jyp;
  a.topo[k] = p1 * a.oro[o00] + p2 * a.oro[o10] + p3 * a.oro[o01] + p4 * a.oro[o11];
  int indz = c.nz - 1; // :105-112 (height is increasing: the first level above the particle)
  for (int il = 2; il <= c.nz; il++)
    if (a.height[il - 1] > zt) { indz = il - 1; break; }
  const int indzp = indz + 1;
  const float dz1 = zt - a.height[indz - 1], dz2 = a.height[indzp - 1] - zt;
  const float dz = 1.f / (dz1 + dz2);
  const float dt1 = (float)(c.itime - c.memtime[0]), dt2 = (float)(c.memtime[1] - c.itime);
  const float dtt = 1.f / (dt1 + dt2);
  const int plane = c.nxd * c.nyd;
  float pvprof[2], qvprof[2], ttprof[2], rhoprof[2];
#pragma unroll
  for (int n = 0; n < 2; n++) {
    const int base = (indz - 1 + n) * plane;
    float pv1[2], qv1[2], tt1[2], rho1[2];
#pragma unroll
    for (int m = 0; m < 2; m++) {
      const float2 qa = a.Q[m][base + o00], qb = a.Q[m][base + o10], qc = a.Q[m][base + o01], qd = a.Q[m][base + o11];
      pv1[m] = p1 * qa.x + p2 * qb.x + p3 * qc.x + p4 * qd.x;
      qv1[m] = p1 * qa.y + p2 * qb.y + p3 * qc.y + p4 * qd.y;
      const float *T = a.met[m].T + base;
      tt1[m] = p1 * T[o00] + p2 * T[o10] + p3 * T[o01] + p4 * T[o11];
      const float4 *A = a.met[m].A + base;
      rho1[m] = p1 * A[o00].w + p2 * A[o10].w + p3 * A[o01].w + p4 * A[o11].w;
    }
    pvprof[n] = (pv1[0] * dt2 + pv1[1] * dt1) * dtt;
    qvprof[n] = (qv1[0] * dt2 + qv1[1] * dt1) * dtt;
    ttprof[n] = (tt1[0] * dt2 + tt1[1] * dt1) * dtt;
    rhoprof[n] = (rho1[0] * dt2 + rho1[1] * dt1) * dtt;
  }
  a.pvi[k] = (dz1 * pvprof[1] + dz2 * pvprof[0]) * dz;
  a.qvi[k] = (dz1 * qvprof[1] + dz2 * qvprof[0]) * dz;
  a.tti[k] = (dz1 * ttprof[1] + dz2 * ttprof[0]) * dz;
  a.rhoi[k] = (dz1 * rhoprof[1] + dz2 * rhoprof[0]) * dz;
  float tr[2], hm[2];
#pragma unroll
  for (int m = 0; m < 2; m++) {
    const float *tp = a.met[m].trop;
    tr[m] = p1 * tp[o00] + p2 * tp[o10] + p3 * tp[o01] + p4 * tp[o11];
    const float4 *S = a.met[m].S;
    hm[m] = p1 * S[o00].x + p2 * S[o10].x + p3 * S[o01].x + p4 * S[o11].x;
  }
  a.hmixi[k] = (hm[0] * dt2 + hm[1] * dt1) * dtt;
  a.tri[k] = (tr[0] * dt2 + tr[1] * dt1) * dtt;
  a.npoint[k] = a.p.npoint[row];
  a.itramem[k] = a.p.itramem[row];
  a.ztra1[k] = zt;
  for (int ks = 0; ks < c.nspec; ks++)
    a.xmass1[(size_t)ks * a.p.maxpart + k] = a.p.xmass1[(size_t)ks * a.p.maxpart + row];
}

// ---- loop hooks: calcfluxes / partpos_average
__global__ void __launch_bounds__(256) hooks_pre_kernel(const HookArgs a) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= a.cfg.numpart) return;
  const bool adv = a.p.itra1[j] == a.cfg.itime;
  a.adv[j] = adv ? 1 : 0;
  if (!adv || !(a.iflux || a.linit)) return;
  // src/timemanager.f90:559-561: xold, yold, zold are default reals
  a.old[j] = (float)a.p.xtra1[j];
  a.old[a.old_stride + j] = (float)a.p.ytra1[j];
  a.old[2 * a.old_stride + j] = a.p.ztra1[j];
  for (int ks = 0; ks < a.cfg.nspec; ks++)
    a.old[(3 + (size_t)ks) * a.old_stride + j] = a.p.xmass1[(size_t)ks * a.p.maxpart + j];
}

// src/calcfluxes.f90:49-166 for row j (nage: src/timemanager.f90:544-547)
__device__ __forceinline__ void flux_row(const HookArgs &a, int j) {
  const DevCfg &c = a.cfg;
  const float xold = a.old[j], yold = a.old[a.old_stride + j], zold = a.old[2 * a.old_stride + j];
  const int itage = abs(c.itime - a.p.itramem[j]);
  int nage;
  for (nage = 1; nage <= c.nageclass; nage++)
    if (itage < c.lage[nage - 1]) break;
  if (nage > c.nageclass) return; // (the reference would index past the array)
  const int kp = (c.ioutputforeachrelease == 1 && c.mdomainfill == 0) ? a.p.npoint[j] : 1;
  fpb_flux_particle(c, a.flux, nage, kp, xold, yold, zold, a.p.xtra1[j], a.p.ytra1[j], a.p.ztra1[j],
                    [&](int k) { return a.old[(3 + (size_t)(k - 1)) * a.old_stride + j]; });
}

// src/partpos_average.f90:47-186 for row j
__device__ __forceinline__ void average_row(const HookArgs &a, int j) {
  const DevCfg &c = a.cfg;
  const double xt = a.p.xtra1[j], yt = a.p.ytra1[j];
  const float zt = a.p.ztra1[j];
  float xlon = (float)((double)c.xlon0 + xt * (double)c.dx), ylat = (float)((double)c.ylat0 + yt * (double)c.dy);
  const int ix = (int)xt, jy = (int)yt;
  const int ixp = ix + 1;
  int jyp = jy + 1;
  const float ddx = (float)(xt - (double)(float)ix), ddy = (float)(yt - (double)(float)jy);
  const float rddx = 1.f - ddx, rddy = 1.f - ddy;
  const float p1 = rddx * rddy, p2 = ddx * rddy, p3 = rddx * ddy, p4 = ddx * ddy;
  if (jyp >= c.nymax) jyp = jyp - 1;
  const int o00 = ix + c.nxd * jy, o10 = ixp + c.nxd * jy, o01 = ix + c.nxd * jyp, o11 = ixp + c.nxd * jyp;
  const float topo = p1 * a.oro[o00] + p2 * a.oro[o10] + p3 * a.oro[o01] + p4 * a.oro[o11];
  int indz = c.nz - 1;
  for (int il = 2; il <= c.nz; il++)
    if (a.height[il - 1] > zt) { indz = il - 1; break; }
  const int indzp = indz + 1;
  const float dz1 = zt - a.height[indz - 1], dz2 = a.height[indzp - 1] - zt;
  const float dz = 1.f / (dz1 + dz2);
  const float dt1 = (float)(c.itime - c.memtime[0]), dt2 = (float)(c.memtime[1] - c.itime);
  const float dtt = 1.f / (dt1 + dt2);
  const int plane = c.nxd * c.nyd;
  float pvprof[2], qvprof[2], ttprof[2], uuprof[2], vvprof[2], rhoprof[2];
#pragma unroll
  for (int n = 0; n < 2; n++) {
    const int base = (indz - 1 + n) * plane;
    float pv1[2], qv1[2], tt1[2], uu1[2], vv1[2], rho1[2];
#pragma unroll
    for (int m = 0; m < 2; m++) {
      const float2 qa = a.Q[m][base + o00], qb = a.Q[m][base + o10], qc = a.Q[m][base + o01], qd = a.Q[m][base + o11];
      pv1[m] = p1 * qa.x + p2 * qb.x + p3 * qc.x + p4 * qd.x;
      qv1[m] = p1 * qa.y + p2 * qb.y + p3 * qc.y + p4 * qd.y;
      const float *T = a.met[m].T + base;
      tt1[m] = p1 * T[o00] + p2 * T[o10] + p3 * T[o01] + p4 * T[o11];
      const float4 *A = a.met[m].A + base;
      const float4 wa = A[o00], wb = A[o10], wc = A[o01], wd = A[o11];
      uu1[m] = p1 * wa.x + p2 * wb.x + p3 * wc.x + p4 * wd.x;
      vv1[m] = p1 * wa.y + p2 * wb.y + p3 * wc.y + p4 * wd.y;
      rho1[m] = p1 * wa.w + p2 * wb.w + p3 * wc.w + p4 * wd.w;
    }
    pvprof[n] = (pv1[0] * dt2 + pv1[1] * dt1) * dtt;
    qvprof[n] = (qv1[0] * dt2 + qv1[1] * dt1) * dtt;
    ttprof[n] = (tt1[0] * dt2 + tt1[1] * dt1) * dtt;
    uuprof[n] = (uu1[0] * dt2 + uu1[1] * dt1) * dtt;
    vvprof[n] = (vv1[0] * dt2 + vv1[1] * dt1) * dtt;
    rhoprof[n] = (rho1[0] * dt2 + rho1[1] * dt1) * dtt;
  }
  const float pvi = (dz1 * pvprof[1] + dz2 * pvprof[0]) * dz;
  const float qvi = (dz1 * qvprof[1] + dz2 * qvprof[0]) * dz;
  const float tti = (dz1 * ttprof[1] + dz2 * ttprof[0]) * dz;
  const float uui = (dz1 * uuprof[1] + dz2 * uuprof[0]) * dz;
  const float vvi = (dz1 * vvprof[1] + dz2 * vvprof[0]) * dz;
  const float rhoi = (dz1 * rhoprof[1] + dz2 * rhoprof[0]) * dz;
  float tr[2], hm[2];
#pragma unroll
  for (int m = 0; m < 2; m++) {
    const float *tp = a.met[m].trop;
    tr[m] = p1 * tp[o00] + p2 * tp[o10] + p3 * tp[o01] + p4 * tp[o11];
    const float4 *S = a.met[m].S;
    hm[m] = p1 * S[o00].x + p2 * S[o10].x + p3 * S[o01].x + p4 * S[o11].x;
  }
  const float hmixi = (hm[0] * dt2 + hm[1] * dt1) * dtt;
  const float tri = (tr[0] * dt2 + tr[1] * dt1) * dtt;
  const float energy = ((tti * 1004.6f + (zt + topo) * 9.81f) + qvi * 2501000.f) + (uui * uui + vvi * vvi) / 2.f;
  const float pi180 = 3.14159265f / 180.f;
  xlon = xlon * pi180;
  ylat = ylat * pi180;
  const float cy = (float)cos((double)ylat), sy = (float)sin((double)ylat);
  const float cx = (float)cos((double)xlon), sx = (float)sin((double)xlon);
  const float x = cy * sx, y = -((1.0f * cy) * cx), z = sy;
  const int s = a.p.slot[j];
  a.npart_av[s] = a.npart_av[s] + 1;
  const float v[14] = {x, y, z, zt, topo, pvi, qvi, tti, uui, vvi, rhoi, tri, hmixi, energy};
#pragma unroll
  for (int q = 0; q < 14; q++) a.av[q * a.av_stride + s] = a.av[q * a.av_stride + s] + v[q];
}

// src/initial_cond_calc.f90:49-204 for row j; old_mass: the masses from before the step (termination by nstop)
__device__ __forceinline__ void init_cond_row(const HookArgs &a, int j, bool old_mass) {
  const DevCfg &c = a.cfg;
  const double xt = a.p.xtra1[j], yt = a.p.ytra1[j];
  const float zt = a.p.ztra1[j];
  float rhoi = 1.f;
  if (a.linit == 1) { // mass unit: rho of memind(2) at the particle
    int ix = (int)xt, jy = (int)yt;
    const float ddx = (float)(xt - (double)(float)ix), ddy = (float)(yt - (double)(float)jy);
    const float rddx = 1.f - ddx, rddy = 1.f - ddy;
    const float p1 = rddx * rddy, p2 = ddx * rddy, p3 = rddx * ddy, p4 = ddx * ddy;
    // (a particle that has left the domain: the reference reads past its arrays; nearest grid point here)
    ix = max(0, min(ix, c.nxd - 1)); jy = max(0, min(jy, c.nyd - 1));
    const int ixp = min(ix + 1, c.nxd - 1), jyp = min(jy + 1, c.nyd - 1);
    int indz = c.nz - 1;
    for (int il = 2; il <= c.nz; il++)
      if (a.height[il - 1] > zt) { indz = il - 1; break; }
    const int indzp = indz + 1;
    const float dz1 = zt - a.height[indz - 1], dz2 = a.height[indzp - 1] - zt;
    const float dz = 1.f / (dz1 + dz2);
    const int plane = c.nxd * c.nyd;
    float rhoprof[2];
#pragma unroll
    for (int n = 0; n < 2; n++) {
      const float4 *A = a.met[1].A + (indz - 1 + n) * plane;
      rhoprof[n] = p1 * A[ix + c.nxd * jy].w + p2 * A[ixp + c.nxd * jy].w + p3 * A[ix + c.nxd * jyp].w + p4 * A[ixp + c.nxd * jyp].w;
    }
    rhoi = (dz1 * rhoprof[1] + dz2 * rhoprof[0]) * dz;
  }
  const int nrelpointer = ((c.ioutputforeachrelease == 0) || (c.mdomainfill == 1)) ? 1 : a.p.npoint[j];
  int kz;
  for (kz = 1; kz <= c.numzgrid; kz++)
    if (c.outheight[kz - 1] > zt) break;
  if (kz > c.numzgrid) return;
  const float xl = (float)((xt * (double)c.dx + (double)c.xoutshift) / (double)c.dxout);
  const float yl = (float)((yt * (double)c.dy + (double)c.youtshift) / (double)c.dyout);
  int ix = (int)xl;
  if (xl < 0.f) ix = ix - 1;
  int jy = (int)yl;
  if (yl < 0.f) jy = jy - 1;
  const size_t nxg = c.numxgrid, nyg = c.numygrid, nzg = c.numzgrid;
  auto mass = [&](int ks) {
    return old_mass ? a.old[(3 + (size_t)(ks - 1)) * a.old_stride + j] : a.p.xmass1[(size_t)(ks - 1) * a.p.maxpart + j];
  };
  auto add = [&](int cx, int cy, int ks, float v) { // init_cond(cx, cy, kz, ks, nrelpointer) += v
    atomicAdd(a.init_cond + cx + nxg * (cy + nyg * ((kz - 1) + nzg * ((ks - 1) + (size_t)a.maxspec * (nrelpointer - 1)))), v);
  };
  if ((xl < 0.5f) || (yl < 0.5f) || (xl > (float)(c.numxgrid - 1) - 0.5f) || (yl > (float)(c.numygrid - 1) - 0.5f)) {
    if ((ix >= 0) && (jy >= 0) && (ix <= c.numxgrid - 1) && (jy <= c.numygrid - 1))
      for (int ks = 1; ks <= c.nspec; ks++) add(ix, jy, ks, mass(ks) / rhoi);
  } else {
    const float ddx = xl - (float)ix, ddy = yl - (float)jy;
    float wx, wy;
    int ixp, jyp;
    if (ddx > 0.5f) { ixp = ix + 1; wx = 1.5f - ddx; } else { ixp = ix - 1; wx = 0.5f + ddx; }
    if (ddy > 0.5f) { jyp = jy + 1; wy = 1.5f - ddy; } else { jyp = jy - 1; wy = 0.5f + ddy; }
    if ((ix >= 0) && (ix <= c.numxgrid - 1)) {
      if ((jy >= 0) && (jy <= c.numygrid - 1)) {
        const float w = wx * wy;
        for (int ks = 1; ks <= c.nspec; ks++) add(ix, jy, ks, mass(ks) / rhoi * w);
      }
      if ((jyp >= 0) && (jyp <= c.numygrid - 1)) {
        const float w = wx * (1.f - wy);
        for (int ks = 1; ks <= c.nspec; ks++) add(ix, jyp, ks, mass(ks) / rhoi * w);
      }
    }
    if ((ixp >= 0) && (ixp <= c.numxgrid - 1)) {
      if ((jyp >= 0) && (jyp <= c.numygrid - 1)) {
        const float w = (1.f - wx) * (1.f - wy);
        for (int ks = 1; ks <= c.nspec; ks++) add(ixp, jyp, ks, mass(ks) / rhoi * w);
      }
      if ((jy >= 0) && (jy <= c.numygrid - 1)) {
        const float w = (1.f - wx) * wy;
        for (int ks = 1; ks <= c.nspec; ks++) add(ixp, jy, ks, mass(ks) / rhoi * w);
      }
    }
  }
}

__global__ void __launch_bounds__(256) hooks_post_kernel(const HookArgs a) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= a.cfg.numpart) return;
  if (a.final_pass) { // src/timemanager.f90:733-737
    if (a.p.itra1[j] == a.cfg.itime) init_cond_row(a, j, false);
    return;
  }
  if (!a.adv[j]) return;
  if (a.linit) {
    const int f = a.flags[j];
    if (f & 4) init_cond_row(a, j, true);        // SC_TERM_NSTOP, src/timemanager.f90:631
    else if (f & 8) init_cond_row(a, j, false);  // SC_TERM_AGE, :702
  }
  if (a.iflux) flux_row(a, j);
  // (the reference also averages a particle that the step has just terminated -- at a position that may lie
  // outside the fields; such a particle is never written by partoutput_average, so it is left out)
  if (a.ipout3 && a.p.itra1[j] != FPB_ITRA_DEAD) average_row(a, j);
}
} // namespace

void fpb_hooks_pre(const HookArgs &a, cudaStream_t st) {
  const int nb = (a.cfg.numpart + 255) / 256;
  if (nb) hooks_pre_kernel<<<nb, 256, 0, st>>>(a);
}
void fpb_hooks_post(const HookArgs &a, cudaStream_t st) {
  const int nb = (a.cfg.numpart + 255) / 256;
  if (nb) hooks_post_kernel<<<nb, 256, 0, st>>>(a);
}

void fpb_partoutput_launch(const PartoutArgs &a, cudaStream_t st) {
  const int nb = (a.numpart + OUT_BLOCK - 1) / OUT_BLOCK;
  if (nb == 0) return;
  partout_count_kernel<<<nb, OUT_BLOCK, 0, st>>>(a);
  partout_scan_kernel<<<1, OUT_BLOCK, 0, st>>>(a, nb);
  partout_write_kernel<<<nb, OUT_BLOCK, 0, st>>>(a);
}

void fpb_density_outgrid(const DensityArgs &a, cudaStream_t st) {
  const int n = a.numx * a.numy * a.numz;
  density_outgrid_kernel<<<(n + 255) / 256, 256, 0, st>>>(a);
}

void fpb_sparse_dump(const SparseDumpArgs &a, cudaStream_t st) {
  const int nb = (a.ncells + OUT_BLOCK - 1) / OUT_BLOCK;
  sparse_count_kernel<<<nb, OUT_BLOCK, 0, st>>>(a);
  sparse_scan_kernel<<<1, OUT_BLOCK, 0, st>>>(a, nb);
  sparse_write_kernel<<<nb, OUT_BLOCK, 0, st>>>(a);
}
