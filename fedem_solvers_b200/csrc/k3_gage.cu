// k3_gage.cu -- strain rosettes / strain gages on sm_100a (the fedem_gage path, config 5).
//
// Reference: InitStrainRosette builds, per rosette, Bcart(3 x ndim) = Teps . B_el(z) . T_el . H_el
// (src/vpmStress/strainRosetteModule.f90:587-812, H_el from ElDispFromSupElDisp,
// displacementModule.f90:1096-1202) and then evaluates, every time step and one rosette after the
// other, epsC = Bcart . finit + epsCInit, sigmaC = C . epsC, the gage legs, Mohr's circle and the
// von Mises values (calcRosetteStrains :251-324, evaluateStrainGages :225-248), pushing
// sigmaP(1) and every leg stress into a per-gage std::vector for rainflow counting
// (AddFatiguePoints, strainGageModule.f90:691-716).
//
// Here:  * setup (once): the small geometric factor bscr(3 x <=24) per rosette is formed on the
//          host exactly like InitStrainRosette does; Bcart = bscr . R[rows of the rosette nodes]
//          is formed on the GPU from the row operator R that fsr_set_recovery left there (R is
//          H_el for every node at once);
//        * per tile of time steps: ALL rosettes at once as one FP64 tensor-core GEMM
//          eps[3 nros x T] = Bcart[3 nros x ndim] . Q[ndim x T] (the K1 DMMA kernel, TMA staged),
//          then one thread per (rosette, step) finishes sigmaC, gage legs, principal values,
//          angles, von Mises and writes the fatigue series (max principal + legs, scaled to MPa)
//          gage-major, which the K3 rainflow kernels (k3_fatigue.cu) consume tile by tile.
// Nothing per-step is kept on the host and no history is ever stored whole.
#include <algorithm>
#include <cmath>

#include "common.cuh"

#define FSR_GAGE_NVAL_ 24

struct fsr_gages {
  fsr_part* part = nullptr;
  int device = 0, nros = 0, ndim = 0, ldk = 0;
  int nrows_pad = 0;   // 3*nros padded to the K1 row tile
  int tile = 0;        // steps per device batch
  double* Bcart = nullptr;    // [nrows_pad][ldk] row-major, zero padded (row = 3*r + component)
  double* Qt = nullptr;       // [tile][ldk]
  double* eps = nullptr;      // [nrows_pad][tile]
  double* hist = nullptr;     // [4*nros][tile] fatigue series, gage-major
  double* values = nullptr;   // [tile][nros][NVAL] staging for host output
  double* Qstage = nullptr; size_t Qstage_cap = 0;
  double* cmat = nullptr;     // [nros][4]: C11, C12, C33, strain coat SCF (0 = strain gage: fatigue series of sigmaP(1))
  double* tg = nullptr;       // [nros][9] Teps_NfromC of up to three legs
  double* eps0 = nullptr;     // [nros][3] epsCInit
  int* ngage = nullptr;       // [nros]
  int* zero_init = nullptr;   // [nros] 1 = epsCInit still to be taken from the first step
  bool zero_pending = false;
  std::vector<fsr_rosette> ros;
  fsr_fatigue_state* fat = nullptr;  // streaming rainflow state of 4*nros series
  double to_mpa = 1.0;
  // strain coat summary (k3 coat kernels below): running envelopes, angle bins [bin][rosette], biaxiality sums
  int coat_nbin = 0;
  double coat_gate = 0.0;
  double* coat_env = nullptr;     // [8][nros]: epsMax, epsMin, sigMax, sigMin, gammaMax, tauMax, vmeMax, vmsMax
  double* coat_bsum = nullptr;    // [2][nros]: sum and sum of squares of the biaxiality ratio
  int* coat_nbiax = nullptr;      // [nros]
  int* coat_nval = nullptr;       // [nros][nbin] hit counts
  unsigned long long* coat_bin = nullptr;   // [nros][4][nbin]: sigMax, sigMin, epsMax, epsMin of every bin, order-preserving encoding
  cudaStream_t stream = nullptr;
};

namespace fsr {

// Bcart[3r+j][c] = sum_i bscr[r][j][i] * R[row(r,i)][c]; one block per rosette, threads over c
__global__ void gage_bcart_kernel(double* __restrict__ Bcart, int ldk, const double* __restrict__ R,
                                  const double* __restrict__ bscr /* [nros][3][24] */,
                                  const int* __restrict__ rows /* [nros][24] 0-based, -1 = unused */, int nros)
{
  const int r = blockIdx.x;
  if (r >= nros) return;
  for (int c = threadIdx.x; c < ldk; c += blockDim.x) {
    double a0 = 0.0, a1 = 0.0, a2 = 0.0;
    for (int i = 0; i < 24; ++i) {
      const int row = rows[r * 24 + i];
      if (row < 0) break;
      const double h = R[(size_t)row * ldk + c];
      a0 += bscr[(r * 3 + 0) * 24 + i] * h;
      a1 += bscr[(r * 3 + 1) * 24 + i] * h;
      a2 += bscr[(r * 3 + 2) * 24 + i] * h;
    }
    Bcart[(size_t)(3 * r + 0) * ldk + c] = a0;
    Bcart[(size_t)(3 * r + 1) * ldk + c] = a1;
    Bcart[(size_t)(3 * r + 2) * ldk + c] = a2;
  }
}

// epsCInit = -Bcart . finit(first step) for rosettes with zeroInit (calcZeroStartRosetteStrains,
// strainRosetteModule.f90:327-353); eps holds Bcart . Q of the current tile, column 0 = first step
__global__ void gage_zero_init_kernel(const double* __restrict__ eps, size_t ldu, int nros, int* __restrict__ zero_init,
                                      double* __restrict__ eps0)
{
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= nros || !zero_init[r]) return;
  for (int j = 0; j < 3; ++j) eps0[3 * r + j] = -eps[(size_t)(3 * r + j) * ldu];
  zero_init[r] = 0;
}

// one thread per (step, rosette), steps fastest: coalesced reads of eps rows and writes of hist rows
// One (rosette, step): strains -> stresses -> gage legs -> Mohr circle; the fatigue series go to hist, the full result record
// (FSR_GAGE_NVAL values, only when WANT) to v[].
template <bool WANT>
__device__ __forceinline__ void gage_point(const double* __restrict__ eps, size_t ldu, int r, int t, const double* __restrict__ cmat,
                                           const double* __restrict__ tg, const double* __restrict__ eps0,
                                           const int* __restrict__ ngage, double to_mpa, double* __restrict__ hist, size_t ld_hist,
                                           double* v)
{
  double e[3], s[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) e[j] = eps[(size_t)(3 * r + j) * ldu + t] + eps0[3 * r + j];
  const double C11 = cmat[4 * r], C12 = cmat[4 * r + 1], C33 = cmat[4 * r + 2];
  s[0] = C11 * e[0] + C12 * e[1];
  s[1] = C12 * e[0] + C11 * e[1];
  s[2] = C33 * e[2];
  const int ng = ngage[r];
  double eg[3] = {0, 0, 0}, sg[3] = {0, 0, 0};
  for (int i = 0; i < ng; ++i) {
    const double* T = tg + 9 * r + 3 * i;
    eg[i] = T[0] * e[0] + T[1] * e[1] + T[2] * e[2];
    sg[i] = T[0] * s[0] + T[1] * s[1] + T[2] * s[2];
  }
  // PrincipleStresses2D (strainAndStressUtils.f90:57-98)
  double origo = (s[0] + s[1]) * 0.5, d12 = s[0] - s[1];
  double radius = sqrt(d12 * d12 + 4.0 * s[2] * s[2]) * 0.5;
  const double sp1 = origo + radius, sp2 = origo - radius, tmax = radius;
  // fatigue series: sigmaP(1) and the leg stresses in MPa (strainGageModule.f90:711-716)
  // strain coats: the signed abs-max principal stress in MPa times the stress concentration factor (strainCoatModule.f90:385-400)
  const double scf = cmat[4 * r + 3];
  hist[(size_t)(4 * r) * ld_hist + t] = scf != 0.0 ? (fabs(sp1) > fabs(sp2) ? sp1 : sp2) * to_mpa * scf : sp1 * to_mpa;
#pragma unroll
  for (int i = 0; i < 3; ++i) hist[(size_t)(4 * r + 1 + i) * ld_hist + t] = sg[i] * to_mpa;
  if (WANT) {
    // PrincipleStrains2D (strainAndStressUtils.f90:14-55): only the per-step result record needs the principal strains and
    // the two atan2 angles, the fatigue pass does not pay for them
    origo = (e[0] + e[1]) * 0.5; d12 = e[0] - e[1];
    const double exy = e[2] * 0.5;
    radius = sqrt(d12 * d12 + e[2] * e[2]) * 0.5;
    const double ep1 = origo + radius, ep2 = origo - radius, gmax = radius * 2.0;
    double alpha1 = 0.0, alphaG = 0.0;
    if (fabs(exy) > kEpsDiv0 || fabs(d12) > kEpsDiv0) {
      alpha1 = atan2(exy, d12) * 0.5;
      alphaG = atan2(d12, exy) * 0.5;
    }
    v[0] = e[0]; v[1] = e[1]; v[2] = e[2];
    v[3] = ep1; v[4] = ep2; v[5] = fabs(ep1) > fabs(ep2) ? ep1 : ep2;
    v[6] = gmax; v[7] = sqrt(ep1 * ep1 + ep2 * ep2 - ep1 * ep2);
    v[8] = alpha1; v[9] = alphaG;
    v[10] = s[0]; v[11] = s[1]; v[12] = s[2];
    v[13] = sp1; v[14] = sp2; v[15] = fabs(sp1) > fabs(sp2) ? sp1 : sp2;
    v[16] = tmax; v[17] = sqrt(sp1 * sp1 + sp2 * sp2 - sp1 * sp2);
    v[18] = eg[0]; v[19] = eg[1]; v[20] = eg[2];
    v[21] = sg[0]; v[22] = sg[1]; v[23] = sg[2];
  }
}

// fatigue pass: no result record.  One thread per (rosette, PAIR of steps), steps fastest: the three strain rows are read and the
// four series rows written as 16-byte accesses, the rosette's constants (13 doubles) are fetched once per pair; same arithmetic
// per step as gage_point.  An odd last step takes the scalar path.
__global__ void __launch_bounds__(256)
gage_post_kernel(const double* __restrict__ eps, size_t ldu, int nros, int nsteps,
                 const double* __restrict__ cmat, const double* __restrict__ tg,
                 const double* __restrict__ eps0, const int* __restrict__ ngage, double to_mpa,
                 double* __restrict__ hist, size_t ld_hist)
{
  const unsigned npair = (unsigned)(nsteps + 1) >> 1;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)nros * npair) return;
  const int r = (int)(idx / npair), t = 2 * (int)(idx - (size_t)r * npair);
  if (t + 1 >= nsteps || ((ldu | ld_hist) & 1)) {
    gage_point<false>(eps, ldu, r, t, cmat, tg, eps0, ngage, to_mpa, hist, ld_hist, nullptr);
    if (t + 1 < nsteps) gage_point<false>(eps, ldu, r, t + 1, cmat, tg, eps0, ngage, to_mpa, hist, ld_hist, nullptr);
    return;
  }
  double2 e[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const double2 x = *reinterpret_cast<const double2*>(eps + (size_t)(3 * r + j) * ldu + t);
    const double e0 = __ldg(eps0 + 3 * r + j);
    e[j] = make_double2(x.x + e0, x.y + e0);
  }
  const double C11 = __ldg(cmat + 4 * r), C12 = __ldg(cmat + 4 * r + 1), C33 = __ldg(cmat + 4 * r + 2), scf = __ldg(cmat + 4 * r + 3);
  const int ng = __ldg(ngage + r);
  double2 s[3];
  s[0] = make_double2(C11 * e[0].x + C12 * e[1].x, C11 * e[0].y + C12 * e[1].y);
  s[1] = make_double2(C12 * e[0].x + C11 * e[1].x, C12 * e[0].y + C11 * e[1].y);
  s[2] = make_double2(C33 * e[2].x, C33 * e[2].y);
  double2 sg[3] = {{0, 0}, {0, 0}, {0, 0}};
#pragma unroll
  for (int i = 0; i < 3; ++i)
    if (i < ng) {
      const double T0 = __ldg(tg + 9 * r + 3 * i), T1 = __ldg(tg + 9 * r + 3 * i + 1), T2 = __ldg(tg + 9 * r + 3 * i + 2);
      sg[i] = make_double2(T0 * s[0].x + T1 * s[1].x + T2 * s[2].x, T0 * s[0].y + T1 * s[1].y + T2 * s[2].y);
    }
  double2 h0;
  {
    const double origo = (s[0].x + s[1].x) * 0.5, d12 = s[0].x - s[1].x;
    const double radius = sqrt(d12 * d12 + 4.0 * s[2].x * s[2].x) * 0.5;
    const double sp1 = origo + radius, sp2 = origo - radius;
    h0.x = scf != 0.0 ? (fabs(sp1) > fabs(sp2) ? sp1 : sp2) * to_mpa * scf : sp1 * to_mpa;
  }
  {
    const double origo = (s[0].y + s[1].y) * 0.5, d12 = s[0].y - s[1].y;
    const double radius = sqrt(d12 * d12 + 4.0 * s[2].y * s[2].y) * 0.5;
    const double sp1 = origo + radius, sp2 = origo - radius;
    h0.y = scf != 0.0 ? (fabs(sp1) > fabs(sp2) ? sp1 : sp2) * to_mpa * scf : sp1 * to_mpa;
  }
  *reinterpret_cast<double2*>(hist + (size_t)(4 * r) * ld_hist + t) = h0;
#pragma unroll
  for (int i = 0; i < 3; ++i)
    *reinterpret_cast<double2*>(hist + (size_t)(4 * r + 1 + i) * ld_hist + t) = make_double2(sg[i].x * to_mpa, sg[i].y * to_mpa);
}

// result-record pass: a block takes 32 steps x 4 rosettes, the records are staged in shared memory and written out as contiguous
// runs.  LAYOUT 0: values[t][r][NVAL] (the host API's per-step record order: 4 x NVAL doubles per step are contiguous);
// LAYOUT 1: values[r][t][NVAL] (the strain coat kernel's order: 32 x NVAL doubles per rosette are contiguous).
template <int LAYOUT>
__global__ void __launch_bounds__(128)
gage_values_kernel(const double* __restrict__ eps, size_t ldu, int nros, int nsteps, const double* __restrict__ cmat,
                   const double* __restrict__ tg, const double* __restrict__ eps0, const int* __restrict__ ngage, double to_mpa,
                   double* __restrict__ hist, size_t ld_hist, double* __restrict__ values, size_t ld_t /* steps per rosette, LAYOUT 1 */)
{
  __shared__ double sv[4][32][FSR_GAGE_NVAL_];
  const int tl = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int r0 = blockIdx.x * 4, t0 = blockIdx.y * 32;
  const int r = r0 + rl, t = t0 + tl;
  if (r < nros && t < nsteps) gage_point<true>(eps, ldu, r, t, cmat, tg, eps0, ngage, to_mpa, hist, ld_hist, sv[rl][tl]);
  __syncthreads();
  const int nr = min(4, nros - r0), nt = min(32, nsteps - t0);
  if (LAYOUT == 0) {
    for (int k = 0; k < nt; ++k) {
      double* dst = values + ((size_t)(t0 + k) * nros + r0) * FSR_GAGE_NVAL_;
      for (int j = threadIdx.x; j < nr * FSR_GAGE_NVAL_; j += 128) dst[j] = sv[j / FSR_GAGE_NVAL_][k][j % FSR_GAGE_NVAL_];
    }
  } else {
    for (int k = 0; k < nr; ++k) {
      double* dst = values + ((size_t)(r0 + k) * ld_t + t0) * FSR_GAGE_NVAL_;
      const double* src = &sv[k][0][0];
      for (int j = threadIdx.x; j < nt * FSR_GAGE_NVAL_; j += 128) dst[j] = src[j];
    }
  }
}

// ---- host geometry (InitStrainRosette up to bscr, strainRosetteModule.f90:630-724) ------------
struct H3 { double x, y, z; };
static H3 hsub(H3 a, H3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
static H3 hcross(H3 a, H3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
static double hdot(H3 a, H3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static bool hnormalize(H3& a)
{
  double l2 = hdot(a, a);
  if (!(l2 > kEpsDiv0 * kEpsDiv0)) return false;
  double l = std::sqrt(l2);
  a = {a.x / l, a.y / l, a.z / l};
  return true;
}

// bscr[3][24] (row j, column i), rows[24] (0-based nodal DOF, -1 padded)
static int rosette_bscr(const fsr_part* p, const fsr_rosette& ro, double* bscr, int* rows)
{
  const int nn = ro.numnod;
  if (nn != 3 && nn != 4) { set_error("rosette %d: %d nodes (3 or 4 expected)", ro.id, nn); return FSR_ERR_ARG; }
  H3 X[4];
  int nd[4], nNDof = 0;
  for (int i = 0; i < nn; ++i) {
    const int n = ro.nodes[i];
    if (n < 1 || n > p->nnod) { set_error("rosette %d: node %d out of range", ro.id, n); return FSR_ERR_ARG; }
    X[i] = {p->xyz_host[3 * (size_t)(n - 1)], p->xyz_host[3 * (size_t)(n - 1) + 1], p->xyz_host[3 * (size_t)(n - 1) + 2]};
    nd[i] = p->madof_host[n] - p->madof_host[n - 1];
    nNDof = std::max(nNDof, nd[i]);
  }
  if (nNDof > 6 || nNDof < 3) { set_error("rosette %d: unsupported nodal DOF count %d", ro.id, nNDof); return FSR_ERR_ARG; }
  // element axes (getShellElementAxes, strainAndStressUtils.f90:339-434)
  H3 ex, ey, ez;
  if (nn == 3) { ex = hsub(X[1], X[0]); ez = hcross(ex, hsub(X[2], X[0])); }
  else ez = hcross(hsub(X[2], X[0]), hsub(X[3], X[1]));
  if (!hnormalize(ez)) { set_error("rosette %d: degenerate element normal", ro.id); return FSR_ERR_ARG; }
  if (nn == 4) { ex = hsub(X[1], X[0]); ey = hcross(ez, ex); ex = hcross(ey, ez); }
  if (!hnormalize(ex)) { set_error("rosette %d: degenerate element x-axis", ro.id); return FSR_ERR_ARG; }
  ey = hcross(ez, ex);
  const double T[3][3] = {{ex.x, ex.y, ex.z}, {ey.x, ey.y, ey.z}, {ez.x, ez.y, ez.z}};
  // c = T_el(1:2,:) . posInGl(:,1:2); Teps rotates element-axes strains to rosette axes (:655-664)
  double c[2][2];
  for (int i = 0; i < 2; ++i)
    for (int j = 0; j < 2; ++j) c[i][j] = T[i][0] * ro.rpos[3 * j] + T[i][1] * ro.rpos[3 * j + 1] + T[i][2] * ro.rpos[3 * j + 2];
  const double Te[3][3] = {{c[0][0] * c[0][0], c[1][0] * c[1][0], c[0][0] * c[1][0]},
                           {c[0][1] * c[0][1], c[1][1] * c[1][1], c[0][1] * c[1][1]},
                           {2.0 * c[0][0] * c[0][1], 2.0 * c[1][1] * c[1][0], c[0][0] * c[1][1] + c[0][1] * c[1][0]}};
  // in-plane shape-function gradients at the gage position (StrainDispCST / StrainDispQuad4 at xi=eta=0)
  double sx[4], sy[4];
  if (nn == 3) {
    double xl[3][3] = {{0}}, yl[3][3] = {{0}};
    for (int a = 0; a < 3; ++a)
      for (int b = 0; b < 3; ++b)
        if (a != b) { H3 d = hsub(X[a], X[b]); xl[a][b] = hdot(ex, d); yl[a][b] = hdot(ey, d); }
    const double a2 = xl[1][0] * yl[2][0] - xl[2][0] * yl[1][0];
    sx[0] = yl[1][2] / a2; sx[1] = yl[2][0] / a2; sx[2] = yl[0][1] / a2;
    sy[0] = -xl[1][2] / a2; sy[1] = -xl[2][0] / a2; sy[2] = -xl[0][1] / a2;
  } else {
    double xl[4] = {0, 0, 0, 0}, yl[4] = {0, 0, 0, 0};
    for (int k = 1; k < 4; ++k) { H3 d = hsub(X[k], X[0]); xl[k] = hdot(ex, d); yl[k] = hdot(ey, d); }
    const double dxi[4] = {-0.25, 0.25, 0.25, -0.25}, det[4] = {-0.25, -0.25, 0.25, 0.25};
    double j11 = 0, j12 = 0, j21 = 0, j22 = 0;
    for (int k = 0; k < 4; ++k) { j11 += dxi[k] * xl[k]; j12 += dxi[k] * yl[k]; j21 += det[k] * xl[k]; j22 += det[k] * yl[k]; }
    const double dj = j11 * j22 - j21 * j12;
    const double i11 = j22 / dj, i22 = j11 / dj, i12 = -j12 / dj, i21 = -j21 / dj;
    for (int k = 0; k < 4; ++k) { sx[k] = i11 * dxi[k] + i12 * det[k]; sy[k] = i21 * dxi[k] + i22 * det[k]; }
  }
  const bool bend = nNDof >= 5 && std::fabs(ro.zpos) > kEpsDiv0;
  for (int k = 0; k < 3 * 24; ++k) bscr[k] = 0.0;
  for (int k = 0; k < 24; ++k) rows[k] = -1;
  int col = 0;
  for (int n = 0; n < nn; ++n) {
    // local B block of node n: columns (u, v, w, rx, ry, rz) in element axes
    double Bl[3][6] = {{0}};
    Bl[0][0] = sx[n]; Bl[2][0] = sy[n]; Bl[1][1] = sy[n]; Bl[2][1] = sx[n];
    if (bend)
      for (int r = 0; r < 3; ++r) { Bl[r][3] = -ro.zpos * Bl[r][1]; Bl[r][4] = ro.zpos * Bl[r][0]; }
    // to global DOF directions, three DOFs at a time: B . T_el (:697-707)
    for (int j = 0; j + 3 <= nd[n]; j += 3) {
      for (int b = 0; b < 3; ++b) {
        double g[3];
        for (int r = 0; r < 3; ++r) g[r] = Bl[r][j] * T[0][b] + Bl[r][j + 1] * T[1][b] + Bl[r][j + 2] * T[2][b];
        for (int r = 0; r < 3; ++r) bscr[r * 24 + col + j + b] = Te[r][0] * g[0] + Te[r][1] * g[1] + Te[r][2] * g[2];
      }
    }
    for (int d = 0; d < nd[n]; ++d) rows[col + d] = p->madof_host[ro.nodes[n] - 1] - 1 + d;
    col += nd[n];
    if (col > 24) { set_error("rosette %d: more than 24 element DOFs", ro.id); return FSR_ERR_ARG; }
  }
  return FSR_OK;
}

// Teps_NfromC of leg i (InitStrainGages, strainGageModule.f90:645-661; vec_to_mat through the
// quaternion of rotationModule.f90:393-428,478-497)
static void gage_direction(const fsr_rosette& ro, int leg, double* out)
{
  const double* X = ro.rpos; const double* Y = ro.rpos + 3; const double* Z = ro.rpos + 6;
  double rv[3];
  for (int k = 0; k < 3; ++k) rv[k] = leg * ro.alpha_gages * Z[k];
  const double eps_th = 0.0005;
  const double thh = 0.5 * std::sqrt(rv[0] * rv[0] + rv[1] * rv[1] + rv[2] * rv[2]);
  double fac;
  if (thh < eps_th) { const double f1 = thh / eps_th; fac = f1 * std::sin(eps_th) / eps_th + 1.0 - f1; }
  else fac = std::sin(thh) / thh;
  double q[4] = {std::cos(thh), rv[0] * fac * 0.5, rv[1] * fac * 0.5, rv[2] * fac * 0.5};
  for (int pass = 0; pass < 2; ++pass) {
    const double nq = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    for (double& v : q) v /= nq;
  }
  const double Rm[3][3] = {
      {2.0 * (q[1] * q[1] + q[0] * q[0]) - 1.0, 2.0 * (q[1] * q[2] - q[3] * q[0]), 2.0 * (q[1] * q[3] + q[2] * q[0])},
      {2.0 * (q[2] * q[1] + q[3] * q[0]), 2.0 * (q[2] * q[2] + q[0] * q[0]) - 1.0, 2.0 * (q[2] * q[3] - q[1] * q[0])},
      {2.0 * (q[3] * q[1] - q[2] * q[0]), 2.0 * (q[3] * q[2] + q[1] * q[0]), 2.0 * (q[3] * q[3] + q[0] * q[0]) - 1.0}};
  double v[3];
  for (int k = 0; k < 3; ++k) v[k] = Rm[k][0] * X[0] + Rm[k][1] * X[1] + Rm[k][2] * X[2];
  const double c = v[0] * X[0] + v[1] * X[1] + v[2] * X[2];
  const double s = v[0] * Y[0] + v[1] * Y[1] + v[2] * Y[2];
  out[0] = c * c; out[1] = s * s; out[2] = c * s;
}

static int gage_buffers(fsr_gages* g, bool want_values)
{
  if (!g->Qt) FSR_CUDA(cudaMalloc(&g->Qt, sizeof(double) * (size_t)g->tile * g->ldk));
  if (!g->eps) FSR_CUDA(cudaMalloc(&g->eps, sizeof(double) * ((size_t)g->nrows_pad * g->tile + 64)));
  if (!g->hist) FSR_CUDA(cudaMalloc(&g->hist, sizeof(double) * (size_t)4 * std::max(g->nros, 1) * g->tile));
  if (want_values && !g->values)
    FSR_CUDA(cudaMalloc(&g->values, sizeof(double) * (size_t)g->tile * std::max(g->nros, 1) * FSR_GAGE_NVAL_));
  return FSR_OK;
}

// Q tile (device) -> eps -> per-step values + fatigue series
static int gage_tile(fsr_gages* g, const double* Q_dev, int ldq, int nsteps, double* values_dev, cudaStream_t s, int values_layout = 0)
{
  const int nsteps_pad = (nsteps + 63) / 64 * 64;
  int rc;
  if ((rc = launch_pack_q_raw(g->Qt, g->ldk, Q_dev, ldq, g->ndim, nsteps, nsteps_pad, s))) return rc;
  if ((rc = launch_k1_raw(g->Bcart, g->Qt, g->eps, g->ldk, g->nrows_pad, nsteps_pad, (size_t)g->tile, s))) return rc;
  if (g->zero_pending) {
    gage_zero_init_kernel<<<(g->nros + 127) / 128, 128, 0, s>>>(g->eps, (size_t)g->tile, g->nros, g->zero_init, g->eps0);
    FSR_LAUNCH_CHECK();
    g->zero_pending = false;
  }
  const size_t total = (size_t)g->nros * nsteps;
  if (total > 0 && !values_dev) {
    const size_t pairs = (size_t)g->nros * ((nsteps + 1) / 2);
    gage_post_kernel<<<(unsigned)((pairs + 255) / 256), 256, 0, s>>>(g->eps, (size_t)g->tile, g->nros, nsteps, g->cmat,
                                                                    g->tg, g->eps0, g->ngage, g->to_mpa, g->hist,
                                                                    (size_t)g->tile);
    FSR_LAUNCH_CHECK();
  } else if (total > 0) {
    const dim3 grid((unsigned)((g->nros + 3) / 4), (unsigned)((nsteps + 31) / 32));
    if (values_layout == 0)
      gage_values_kernel<0><<<grid, 128, 0, s>>>(g->eps, (size_t)g->tile, g->nros, nsteps, g->cmat, g->tg, g->eps0, g->ngage, g->to_mpa, g->hist,
                                                 (size_t)g->tile, values_dev, 0);
    else
      gage_values_kernel<1><<<grid, 128, 0, s>>>(g->eps, (size_t)g->tile, g->nros, nsteps, g->cmat, g->tg, g->eps0, g->ngage, g->to_mpa, g->hist,
                                                 (size_t)g->tile, values_dev, (size_t)nsteps);
    FSR_LAUNCH_CHECK();
  }
  return FSR_OK;
}

static int stage_q(fsr_gages* g, const double* Q, int ldq, int nsteps, cudaStream_t s)
{
  const size_t qbytes = sizeof(double) * (size_t)ldq * nsteps;
  if (g->Qstage_cap < qbytes) {
    cudaFree(g->Qstage); g->Qstage = nullptr; g->Qstage_cap = 0;
    FSR_CUDA(cudaMalloc(&g->Qstage, std::max<size_t>(qbytes, 8)));
    g->Qstage_cap = qbytes;
  }
  FSR_CUDA(cudaMemcpyAsync(g->Qstage, Q, qbytes, cudaMemcpyHostToDevice, s));
  return FSR_OK;
}

}  // namespace fsr

using namespace fsr;

extern "C" {

void fsr_gage_destroy(fsr_gages* g)
{
  if (!g) return;
  cudaSetDevice(g->device);
  if (g->stream) cudaStreamSynchronize(g->stream);
  if (g->fat) fsr_fatigue_destroy(g->fat);
  cudaFree(g->Bcart); cudaFree(g->Qt); cudaFree(g->eps); cudaFree(g->hist); cudaFree(g->values);
  cudaFree(g->Qstage); cudaFree(g->cmat); cudaFree(g->tg); cudaFree(g->eps0); cudaFree(g->ngage);
  cudaFree(g->zero_init);
  cudaFree(g->coat_env); cudaFree(g->coat_bsum); cudaFree(g->coat_nbiax); cudaFree(g->coat_nval); cudaFree(g->coat_bin);
  if (g->stream) cudaStreamDestroy(g->stream);
  delete g;
}

int fsr_gage_create(fsr_gages** out, fsr_part* p, const fsr_rosette* ros, int nros)
{
  if (!out || !p || (nros > 0 && !ros) || nros < 0) { set_error("fsr_gage_create: bad arguments"); return FSR_ERR_ARG; }
  *out = nullptr;
  if (!p->have_R) { set_error("fsr_gage_create: call fsr_set_recovery first (Bcart needs the B and E matrices)"); return FSR_ERR_STATE; }
  FSR_CUDA(cudaSetDevice(p->device));
  fsr_gages* g = new fsr_gages();
  g->part = p; g->device = p->device; g->nros = nros; g->ndim = p->ndim; g->ldk = p->ldk;
  g->nrows_pad = (3 * nros + 127) / 128 * 128;
  g->tile = 512;
  g->ros.assign(ros, ros + nros);
  auto fail = [&](int code) { fsr_gage_destroy(g); return code; };
  if (cudaStreamCreateWithFlags(&g->stream, cudaStreamNonBlocking) != cudaSuccess) { set_error("cudaStreamCreate failed"); return fail(FSR_ERR_CUDA); }

  const size_t nr = (size_t)std::max(nros, 1);
  std::vector<double> bscr(nr * 72), cmat(nr * 4, 0.0), tg(nr * 9, 0.0), eps0(nr * 3, 0.0);
  std::vector<int> rows(nr * 24, -1), ngage(nr, 0), zinit(nr, 0);
  for (int r = 0; r < nros; ++r) {
    const fsr_rosette& ro = ros[r];
    int rc = rosette_bscr(p, ro, &bscr[(size_t)r * 72], &rows[(size_t)r * 24]);
    if (rc) return fail(rc);
    if (ro.ngage < 0 || ro.ngage > 3) { set_error("rosette %d: %d gages (0..3 expected)", ro.id, ro.ngage); return fail(FSR_ERR_ARG); }
    ngage[r] = ro.ngage;
    zinit[r] = ro.zero_init ? 1 : 0;
    if (ro.zero_init) g->zero_pending = true;
    // isoMat2D (isoMatModule.f90:21-38)
    cmat[4 * (size_t)r] = ro.emod / (1.0 - ro.nu * ro.nu);
    cmat[4 * (size_t)r + 1] = ro.nu * cmat[4 * (size_t)r];
    cmat[4 * (size_t)r + 2] = 0.5 * ro.emod / (1.0 + ro.nu);
    for (int i = 0; i < ro.ngage; ++i) gage_direction(ro, i, &tg[9 * (size_t)r + 3 * i]);
  }
  double* d_bscr = nullptr; int* d_rows = nullptr;
  bool ok = cudaMalloc(&g->Bcart, sizeof(double) * (size_t)g->nrows_pad * g->ldk) == cudaSuccess &&
            cudaMalloc(&g->cmat, sizeof(double) * cmat.size()) == cudaSuccess &&
            cudaMalloc(&g->tg, sizeof(double) * tg.size()) == cudaSuccess &&
            cudaMalloc(&g->eps0, sizeof(double) * eps0.size()) == cudaSuccess &&
            cudaMalloc(&g->ngage, sizeof(int) * ngage.size()) == cudaSuccess &&
            cudaMalloc(&g->zero_init, sizeof(int) * zinit.size()) == cudaSuccess &&
            cudaMalloc(&d_bscr, sizeof(double) * bscr.size()) == cudaSuccess &&
            cudaMalloc(&d_rows, sizeof(int) * rows.size()) == cudaSuccess;
  if (!ok) { set_error("fsr_gage_create: device allocation failed"); cudaFree(d_bscr); cudaFree(d_rows); return fail(FSR_ERR_ALLOC); }
  cudaStream_t s = g->stream;
  cudaMemcpyAsync(g->cmat, cmat.data(), sizeof(double) * cmat.size(), cudaMemcpyHostToDevice, s);
  cudaMemcpyAsync(g->tg, tg.data(), sizeof(double) * tg.size(), cudaMemcpyHostToDevice, s);
  cudaMemcpyAsync(g->eps0, eps0.data(), sizeof(double) * eps0.size(), cudaMemcpyHostToDevice, s);
  cudaMemcpyAsync(g->ngage, ngage.data(), sizeof(int) * ngage.size(), cudaMemcpyHostToDevice, s);
  cudaMemcpyAsync(g->zero_init, zinit.data(), sizeof(int) * zinit.size(), cudaMemcpyHostToDevice, s);
  cudaMemcpyAsync(d_bscr, bscr.data(), sizeof(double) * bscr.size(), cudaMemcpyHostToDevice, s);
  cudaMemcpyAsync(d_rows, rows.data(), sizeof(int) * rows.size(), cudaMemcpyHostToDevice, s);
  cudaMemsetAsync(g->Bcart, 0, sizeof(double) * (size_t)g->nrows_pad * g->ldk, s);
  if (nros > 0) {
    gage_bcart_kernel<<<nros, 128, 0, s>>>(g->Bcart, g->ldk, p->R, d_bscr, d_rows, nros);
    ++g_launches;
  }
  cudaError_t e = cudaStreamSynchronize(s);
  cudaFree(d_bscr); cudaFree(d_rows);
  if (e != cudaSuccess) { set_error("fsr_gage_create: %s", cudaGetErrorString(e)); return fail(FSR_ERR_CUDA); }

  // ElDispFromSupElDisp puts the unit response of external DOF i at equation meqn2(i)
  // (displacementModule.f90:1170) whereas calcIntDisplacements (and therefore R) associates
  // finit(i) with meqn2(dofPosIn2(i)).  Identical whenever meqn2 is in nodal order; otherwise move
  // the unit entries so that Bcart equals the reference's.
  bool permuted = false;
  for (size_t j = 0; j < p->extcol.size(); ++j) permuted = permuted || p->extcol[j] != (int)j;
  if (permuted && nros > 0 && !p->ext_rowptr.empty()) {
    std::vector<double> hb((size_t)g->nrows_pad * g->ldk);
    FSR_CUDA(cudaMemcpy(hb.data(), g->Bcart, sizeof(double) * hb.size(), cudaMemcpyDeviceToHost));
    for (int r = 0; r < nros; ++r)
      for (int i = 0; i < 24 && rows[(size_t)r * 24 + i] >= 0; ++i) {
        const int d = rows[(size_t)r * 24 + i];
        for (int ip = p->ext_rowptr[d]; ip < p->ext_rowptr[d + 1]; ++ip) {
          const int j = p->ext_j[ip];
          const double w = p->ext_w[ip];
          for (int c3 = 0; c3 < 3; ++c3) {
            const double b = bscr[(size_t)r * 72 + c3 * 24 + i] * w;
            hb[(size_t)(3 * r + c3) * g->ldk + j] += b;
            hb[(size_t)(3 * r + c3) * g->ldk + p->extcol[j]] -= b;
          }
        }
      }
    FSR_CUDA(cudaMemcpy(g->Bcart, hb.data(), sizeof(double) * hb.size(), cudaMemcpyHostToDevice));
  }
  *out = g;
  return FSR_OK;
}

int fsr_gage_num_series(const fsr_gages* g) { return g ? 4 * g->nros : FSR_ERR_ARG; }

int fsr_gage_get_bcart(fsr_gages* g, double* bcart)
{
  if (!g || !bcart) { set_error("fsr_gage_get_bcart: bad arguments"); return FSR_ERR_ARG; }
  FSR_CUDA(cudaSetDevice(g->device));
  std::vector<double> hb((size_t)g->nrows_pad * g->ldk);
  FSR_CUDA(cudaMemcpy(hb.data(), g->Bcart, sizeof(double) * hb.size(), cudaMemcpyDeviceToHost));
  for (int r = 0; r < g->nros; ++r)
    for (int c = 0; c < g->ndim; ++c)
      for (int j = 0; j < 3; ++j) bcart[(size_t)r * 3 * g->ndim + (size_t)c * 3 + j] = hb[(size_t)(3 * r + j) * g->ldk + c];
  return FSR_OK;
}

int fsr_gage_recover_dev(fsr_gages* g, const double* Q_dev, int ldq, int nsteps, double* values_dev, void* stream)
{
  if (!g || !Q_dev || nsteps < 0 || ldq < g->ndim) { set_error("fsr_gage_recover_dev: bad arguments"); return FSR_ERR_ARG; }
  FSR_CUDA(cudaSetDevice(g->device));
  int rc = gage_buffers(g, false);
  if (rc) return rc;
  cudaStream_t s = stream ? (cudaStream_t)stream : g->stream;
  for (int t0 = 0; t0 < nsteps; t0 += g->tile) {
    const int nt = std::min(g->tile, nsteps - t0);
    rc = gage_tile(g, Q_dev + (size_t)t0 * ldq, ldq, nt,
                   values_dev ? values_dev + (size_t)t0 * g->nros * FSR_GAGE_NVAL_ : nullptr, s);
    if (rc) return rc;
  }
  return FSR_OK;
}

int fsr_gage_recover(fsr_gages* g, const double* Q, int ldq, int nsteps, double* values)
{
  if (!g || !Q || nsteps < 0 || ldq < g->ndim) { set_error("fsr_gage_recover: bad arguments"); return FSR_ERR_ARG; }
  FSR_CUDA(cudaSetDevice(g->device));
  int rc = gage_buffers(g, values != nullptr);
  if (rc) return rc;
  cudaStream_t s = g->stream;
  if ((rc = stage_q(g, Q, ldq, nsteps, s))) return rc;
  for (int t0 = 0; t0 < nsteps; t0 += g->tile) {
    const int nt = std::min(g->tile, nsteps - t0);
    if ((rc = gage_tile(g, g->Qstage + (size_t)t0 * ldq, ldq, nt, values ? g->values : nullptr, s))) return rc;
    if (values) {
      FSR_CUDA(cudaMemcpyAsync(values + (size_t)t0 * g->nros * FSR_GAGE_NVAL_, g->values,
                               sizeof(double) * (size_t)nt * g->nros * FSR_GAGE_NVAL_, cudaMemcpyDeviceToHost, s));
      FSR_CUDA(cudaStreamSynchronize(s));
    }
  }
  FSR_CUDA(cudaStreamSynchronize(s));
  return FSR_OK;
}

int fsr_gage_set_coat_fatigue(fsr_gages* g, const double* scf)
{
  if (!g) { set_error("fsr_gage_set_coat_fatigue: bad arguments"); return FSR_ERR_ARG; }
  FSR_CUDA(cudaSetDevice(g->device));
  std::vector<double> z((size_t)std::max(g->nros, 1), 0.0);
  if (scf) std::copy(scf, scf + g->nros, z.begin());
  FSR_CUDA(cudaStreamSynchronize(g->stream));
  if (g->nros > 0)
    FSR_CUDA(cudaMemcpy2D(g->cmat + 3, 4 * sizeof(double), z.data(), sizeof(double), sizeof(double), (size_t)g->nros, cudaMemcpyHostToDevice));
  return FSR_OK;
}

int fsr_gage_fatigue_begin(fsr_gages* g, double to_mpa, double default_gate, const double* default_curve,
                           double bin_size, int nbins, int stack_cap)
{
  if (!g || !default_curve) { set_error("fsr_gage_fatigue_begin: bad arguments"); return FSR_ERR_ARG; }
  FSR_CUDA(cudaSetDevice(g->device));
  if (g->fat) { fsr_fatigue_destroy(g->fat); g->fat = nullptr; }
  g->to_mpa = to_mpa;
  int rc = fsr_fatigue_create(&g->fat, g->device, 4 * g->nros, default_gate, default_curve, bin_size, nbins, stack_cap);
  if (rc) return rc;
  // per-rosette gate / S-N data override the defaults when given (> 0), reportDamage
  // (strainGageModule.f90:806-812): m2 has no per-rosette default rule, it is taken as given
  std::vector<double> gate((size_t)4 * std::max(g->nros, 1)), curve((size_t)16 * std::max(g->nros, 1));
  for (int r = 0; r < g->nros; ++r) {
    const fsr_rosette& ro = g->ros[r];
    double c[4];
    for (int k = 0; k < 4; ++k) c[k] = ro.sncurve[k];
    for (int k = 0; k < 3; ++k) if (c[k] <= 0.0) c[k] = default_curve[k];
    if (c[3] <= 0.0) c[3] = default_curve[3];
    for (int k = 0; k < 4; ++k) {
      gate[4 * (size_t)r + k] = ro.gate > 0.0 ? ro.gate : default_gate;
      for (int m = 0; m < 4; ++m) curve[16 * (size_t)r + 4 * k + m] = c[m];
    }
  }
  if (g->nros > 0) rc = fsr_fatigue_set_gage_params(g->fat, gate.data(), curve.data());
  return rc;
}

int fsr_gage_fatigue_feed_dev(fsr_gages* g, const double* Q_dev, int ldq, int step0, int nsteps, int mode,
                              int* n_pending, void* stream)
{
  if (!g || !g->fat || !Q_dev || nsteps < 0 || ldq < g->ndim) { set_error("fsr_gage_fatigue_feed_dev: bad arguments / no fsr_gage_fatigue_begin"); return FSR_ERR_ARG; }
  FSR_CUDA(cudaSetDevice(g->device));
  int rc = gage_buffers(g, false);
  if (rc) return rc;
  cudaStream_t s = stream ? (cudaStream_t)stream : g->stream;
  int pend_total = 0;
  for (int t0 = 0; t0 < nsteps; t0 += g->tile) {
    const int nt = std::min(g->tile, nsteps - t0);
    if ((rc = gage_tile(g, Q_dev + (size_t)t0 * ldq, ldq, nt, nullptr, s))) return rc;
    if (mode == 0) {
      int pend = 0;
      rc = fsr_fatigue_locate_dev(g->fat, g->hist, (size_t)g->tile, FSR_HIST_GAGE_MAJOR, step0 + t0, nt,
                                  n_pending ? &pend : nullptr, s);
      pend_total = pend;
      if (rc) return rc;
      if (n_pending && pend == 0) break;
    } else if ((rc = fsr_fatigue_feed_dev(g->fat, g->hist, (size_t)g->tile, FSR_HIST_GAGE_MAJOR, step0 + t0, nt, s)))
      return rc;
  }
  if (n_pending) *n_pending = pend_total;
  return FSR_OK;
}

int fsr_gage_fatigue_end(fsr_gages* g, double* damage, int* ncycles, int* bins, int* status)
{
  if (!g || !g->fat) { set_error("fsr_gage_fatigue_end: no fsr_gage_fatigue_begin"); return FSR_ERR_ARG; }
  FSR_CUDA(cudaSetDevice(g->device));
  FSR_CUDA(cudaStreamSynchronize(g->stream));
  return fsr_fatigue_finish(g->fat, damage, ncycles, bins, status);
}

int fsr_gage_fatigue(fsr_gages* g, const double* Q, int ldq, int nsteps, double to_mpa, double gate,
                     const double* curve, double bin_size, int nbins, double* damage, int* ncycles, int* bins,
                     int* status)
{
  if (!g || !Q || nsteps < 0 || ldq < g->ndim || !curve) { set_error("fsr_gage_fatigue: bad arguments"); return FSR_ERR_ARG; }
  FSR_CUDA(cudaSetDevice(g->device));
  int rc = fsr_gage_fatigue_begin(g, to_mpa, gate, curve, bin_size, nbins, std::min(nsteps + 8, 1 << 16));
  if (rc) return rc;
  cudaStream_t s = g->stream;
  if ((rc = stage_q(g, Q, ldq, nsteps, s))) return rc;
  // zero-start strains must come from the first step in BOTH passes: remember the flags
  int pend = 0;
  if ((rc = fsr_gage_fatigue_feed_dev(g, g->Qstage, ldq, 0, nsteps, 0, &pend, s))) return rc;
  if ((rc = fsr_gage_fatigue_feed_dev(g, g->Qstage, ldq, 0, nsteps, 1, nullptr, s))) return rc;
  FSR_CUDA(cudaStreamSynchronize(s));
  return fsr_fatigue_finish(g->fat, damage, ncycles, bins, status);
}

}  // extern "C"

// ---- strain coat summary ---------------------------------------------------------------------------------------------------
// calcStrainCoatData (src/vpmStress/strainCoatModule.f90:315-480) for every rosette as one coat result point: the running
// envelopes (updateMax / updateMin, :349-364,410-420; max starts at 0, min at hugeVal, nullifyResults :142-170), the angle bins
// of the principal directions (updateAngBin :436-478: nBin = -angleBins - 1 bins over 180 degrees; the bin of the largest
// principal value counts the hit and tracks the range of sigmaP(1) / epsP(1), the bin of the smallest one, 90 degrees away, tracks
// sigmaP(2) / epsP(2) without counting) and the biaxiality sums (updateBiAxial :421-434, gated on the signed abs-max principal
// stress).  Sequential per point over the time steps, independent between points: one thread per rosette, state in HBM laid out
// [bin][rosette].  calcAngleData (:481-547) and BiAxMean / BiAxStdDev (:672-704) finish it.
namespace fsr {

// doubles as order-preserving unsigned integers: max / min of the encodings = encoding of the max / min, so that the bin updates
// (which commute: counts, max, min) can be done with native 64-bit integer atomics by the 32 lanes of a warp on different steps
__device__ __forceinline__ unsigned long long coat_enc(double x)
{
  const long long b = __double_as_longlong(x);
  return (unsigned long long)(b ^ ((b >> 63) | (long long)0x8000000000000000ULL));
}
__device__ __forceinline__ double coat_dec(unsigned long long u)
{
  const long long b = (long long)u;
  return __longlong_as_double(b ^ (((~b) >> 63) | (long long)0x8000000000000000ULL));
}
constexpr unsigned long long kCoatMaxInit = 0ULL, kCoatMinInit = ~0ULL;   // below / above the encoding of every double

// One warp per result point; the point's bin block (nbin counts + 4 x nbin encoded doubles, contiguous in HBM) is streamed
// through shared memory once per tile of steps: lanes take the steps t = lane, lane + 32, ... and update the bins with shared
// memory atomics.  The envelopes and the biaxiality sums are lane-local and folded with shuffles at the end.
__global__ void coat_update_kernel(const double* __restrict__ values, int nros, int nsteps, int nbin, double gate,
                                   double* __restrict__ env, double* __restrict__ bsum, int* __restrict__ nbiax,
                                   int* __restrict__ nval, unsigned long long* __restrict__ bin)
{
  extern __shared__ __align__(16) unsigned long long coat_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  const int r = blockIdx.x * wpb + warp;
  if (r >= nros) return;
  const int nbp = (nbin + 1) & ~1;                            // counts padded to a multiple of 8 bytes
  unsigned long long* sB = coat_smem + (size_t)warp * (4 * nbin + nbp / 2);
  int* sN = reinterpret_cast<int*>(sB + 4 * nbin);
  const unsigned long long* gB = bin + (size_t)r * 4 * nbin;
  int* gN = nval + (size_t)r * nbin;
  for (int i = lane; i < 4 * nbin; i += 32) sB[i] = gB[i];
  for (int i = lane; i < nbin; i += 32) sN[i] = gN[i];
  __syncwarp();
  const double pi = 3.141592653589793238;
  double e[8] = {0.0, kHuge, 0.0, kHuge, 0.0, 0.0, 0.0, 0.0};   // neutral w.r.t. the stored envelopes (max from 0, min from hugeVal)
  double s1 = 0.0, s2 = 0.0;
  int nb = 0;
  for (int t = lane; t < nsteps; t += 32) {
    const double* v = values + ((size_t)r * nsteps + t) * FSR_GAGE_NVAL_;   // [rosette][step][NVAL] (gage_values_kernel<1>)
    const double epsP1 = v[3], epsP2 = v[4], sigP1 = v[13], sigP2 = v[14], sigP3 = v[15];
    e[0] = fmax(e[0], epsP1); e[1] = fmin(e[1], epsP2); e[2] = fmax(e[2], sigP1); e[3] = fmin(e[3], sigP2);
    e[4] = fmax(e[4], v[6]); e[5] = fmax(e[5], v[16]); e[6] = fmax(e[6], v[7]); e[7] = fmax(e[7], v[17]);
    const double angle = v[8];
    int iAng = (int)llround((angle / pi + 0.5) * nbin);   // nint: half away from zero
    int jAng = (int)llround((angle / pi + 1.0) * nbin);
    if (iAng < 1) iAng = nbin;
    if (jAng > nbin) jAng = jAng - nbin;
    --iAng; --jAng;
    atomicAdd(&sN[iAng], 1);
    const unsigned long long s1e = coat_enc(sigP1), e1e = coat_enc(epsP1), s2e = coat_enc(sigP2), e2e = coat_enc(epsP2);
    atomicMax(&sB[iAng], s1e); atomicMin(&sB[nbin + iAng], s1e); atomicMax(&sB[2 * nbin + iAng], e1e); atomicMin(&sB[3 * nbin + iAng], e1e);
    atomicMax(&sB[jAng], s2e); atomicMin(&sB[nbin + jAng], s2e); atomicMax(&sB[2 * nbin + jAng], e2e); atomicMin(&sB[3 * nbin + jAng], e2e);
    if (sigP3 > gate) {
      const double biaxial = fabs(sigP1) > fabs(sigP2) ? sigP2 / sigP1 : sigP1 / sigP2;
      s1 = s1 + biaxial;
      s2 = s2 + __dmul_rn(biaxial, biaxial);
      ++nb;
    }
  }
  __syncwarp();
  unsigned long long* wB = bin + (size_t)r * 4 * nbin;
  for (int i = lane; i < 4 * nbin; i += 32) wB[i] = sB[i];
  for (int i = lane; i < nbin; i += 32) gN[i] = sN[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const double other = __shfl_xor_sync(0xffffffffu, e[k], o);
      e[k] = (k == 1 || k == 3) ? fmin(e[k], other) : fmax(e[k], other);
    }
    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    nb += __shfl_xor_sync(0xffffffffu, nb, o);
  }
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const double old = env[(size_t)k * nros + r];
      env[(size_t)k * nros + r] = (k == 1 || k == 3) ? fmin(old, e[k]) : fmax(old, e[k]);
    }
    bsum[r] += s1; bsum[(size_t)nros + r] += s2;
    nbiax[r] += nb;
  }
}

// calcAngleData with useOldRange = .false. + BiAxMean / BiAxStdDev; out [6][nros]: sRange(1), sRange(2), popAng, angSpd,
// biaxial mean, biaxial standard deviation.  A bin is allocated once anything was folded into it (its max left the initial value);
// bins that only ever held the smaller principal value (count 0) are freed before the gap search, as in the reference.
__global__ void coat_finish_kernel(int nros, int nbin, const int* __restrict__ nval, const unsigned long long* __restrict__ bin,
                                   const double* __restrict__ bsum, const int* __restrict__ nbiax, double* __restrict__ out)
{
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= nros) return;
  const unsigned long long* B = bin + (size_t)r * 4 * nbin;
  const int* N = nval + (size_t)r * nbin;
  const double binSize = 180.0 / nbin;
  double sr1 = 0.0, sr2 = 0.0;
  int iGap = 0, mVal = 0;
  for (int i = 1; i <= nbin; ++i) {
    if (B[i - 1] == kCoatMaxInit) continue;
    sr1 = fmax(sr1, coat_dec(B[i - 1]) - coat_dec(B[nbin + i - 1]));
    sr2 = fmax(sr2, coat_dec(B[2 * nbin + i - 1]) - coat_dec(B[3 * nbin + i - 1]));
    if (N[i - 1] > mVal) { iGap = i; mVal = N[i - 1]; }
  }
  const double popAng = __dsub_rn(__dmul_rn((double)iGap, binSize), 90.0);   // unfused, like the reference's arithmetic
  int firstGap = 0, maxGap = 0;
  iGap = 0;
  for (int i = 1; i <= nbin; ++i) {
    if (N[i - 1] > 0) {
      if (firstGap == 0) firstGap = i;
      else if (iGap > 0) { maxGap = max(maxGap, i - iGap + 1); iGap = 0; }
    } else if (iGap == 0 && firstGap > 0)
      iGap = i;
  }
  if (iGap > 0) firstGap = firstGap + nbin - iGap + 1;
  if (firstGap > maxGap) maxGap = firstGap;
  out[r] = sr1; out[(size_t)nros + r] = sr2; out[(size_t)2 * nros + r] = popAng;
  out[(size_t)3 * nros + r] = __dsub_rn(180.0, __dmul_rn((double)maxGap, binSize));
  const int nb = nbiax[r];
  const double s1 = bsum[r], s2 = bsum[(size_t)nros + r];
  out[(size_t)4 * nros + r] = s1 / (double)max(1, nb);
  double sd = 0.0;
  const double dnum = (double)nb;
  if (dnum > 1.0) {
    const double mean = s1 / dnum, dvar = __dsub_rn(s2 / dnum, __dmul_rn(mean, mean));
    if (dvar > 0.0) sd = sqrt(dvar * dnum / (dnum - 1.0));
  }
  out[(size_t)5 * nros + r] = sd;
}

__global__ void coat_init_kernel(int nros, int nbin, double* env, double* bsum, int* nbiax, int* nval, unsigned long long* bin)
{
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < (size_t)nbin * nros) {
    nval[i] = 0;
    const size_t r = i / nbin, b = i % nbin;
    unsigned long long* B = bin + r * 4 * nbin;
    B[b] = kCoatMaxInit; B[nbin + b] = kCoatMinInit; B[2 * (size_t)nbin + b] = kCoatMaxInit; B[3 * (size_t)nbin + b] = kCoatMinInit;
  }
  if (i < (size_t)nros) {
    for (int k = 0; k < 8; ++k) env[(size_t)k * nros + i] = (k == 1 || k == 3) ? kHuge : 0.0;
    bsum[i] = 0.0; bsum[(size_t)nros + i] = 0.0;
    nbiax[i] = 0;
  }
}

}  // namespace fsr

extern "C" {

int fsr_coat_begin(fsr_gages* g, int angle_bins, double biaxial_gate)
{
  if (!g || angle_bins < 3) { set_error("fsr_coat_begin: bad arguments"); return FSR_ERR_ARG; }
  FSR_CUDA(cudaSetDevice(g->device));
  const int nbin = angle_bins - 1;   // allocate(angBin(angBinSize-1)), strainCoatModule.f90:297-299
  const size_t nr = (size_t)std::max(g->nros, 1);
  if (nbin != g->coat_nbin) {
    cudaFree(g->coat_nval); cudaFree(g->coat_bin); g->coat_nval = nullptr; g->coat_bin = nullptr;
  }
  g->coat_nbin = nbin; g->coat_gate = biaxial_gate;
  if (!g->coat_env) FSR_CUDA(cudaMalloc(&g->coat_env, sizeof(double) * 8 * nr));
  if (!g->coat_bsum) FSR_CUDA(cudaMalloc(&g->coat_bsum, sizeof(double) * 2 * nr));
  if (!g->coat_nbiax) FSR_CUDA(cudaMalloc(&g->coat_nbiax, sizeof(int) * nr));
  if (!g->coat_nval) FSR_CUDA(cudaMalloc(&g->coat_nval, sizeof(int) * nr * nbin));
  if (!g->coat_bin) FSR_CUDA(cudaMalloc(&g->coat_bin, sizeof(unsigned long long) * 4 * nr * nbin));
  const size_t n = nr * nbin;
  fsr::coat_init_kernel<<<(unsigned)((n + 255) / 256), 256, 0, g->stream>>>(g->nros, nbin, g->coat_env, g->coat_bsum, g->coat_nbiax, g->coat_nval, g->coat_bin);
  FSR_LAUNCH_CHECK();
  return FSR_OK;
}

// the next nsteps steps of the reduced history (device): rosette strains / stresses (K1 GEMM + Mohr circle) -> coat update
int fsr_coat_feed_dev(fsr_gages* g, const double* Q_dev, int ldq, int nsteps, void* stream)
{
  if (!g || !Q_dev || nsteps < 0 || ldq < g->ndim || g->coat_nbin < 1) { set_error("fsr_coat_feed_dev: bad arguments (fsr_coat_begin first)"); return FSR_ERR_ARG; }
  FSR_CUDA(cudaSetDevice(g->device));
  int rc = gage_buffers(g, true);
  if (rc) return rc;
  cudaStream_t s = stream ? (cudaStream_t)stream : g->stream;
  for (int t0 = 0; t0 < nsteps; t0 += g->tile) {
    const int nt = std::min(g->tile, nsteps - t0);
    if ((rc = gage_tile(g, Q_dev + (size_t)t0 * ldq, ldq, nt, g->values, s, 1))) return rc;   // records as [rosette][step][NVAL]
    if (g->nros > 0) {
      const int nbin = g->coat_nbin;
      const size_t per_warp = sizeof(unsigned long long) * (4 * (size_t)nbin + ((nbin + 1) & ~1) / 2);
      int wpb = 4;
      while (wpb > 1 && per_warp * wpb > 100 * 1024) wpb >>= 1;      // two blocks per SM when they fit
      if (per_warp * wpb > 227 * 1024) { set_error("fsr_coat_feed: %d angle bins do not fit the shared memory of an SM", nbin + 1); return FSR_ERR_ARG; }
      FSR_CUDA(cudaFuncSetAttribute(fsr::coat_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(per_warp * wpb)));
      fsr::coat_update_kernel<<<(g->nros + wpb - 1) / wpb, wpb * 32, per_warp * wpb, s>>>(g->values, g->nros, nt, nbin, g->coat_gate, g->coat_env,
                                                                                       g->coat_bsum, g->coat_nbiax, g->coat_nval, g->coat_bin);
      FSR_LAUNCH_CHECK();
    }
  }
  return FSR_OK;
}

int fsr_coat_feed(fsr_gages* g, const double* Q, int ldq, int nsteps)
{
  if (!g || !Q || nsteps < 0 || ldq < g->ndim) { set_error("fsr_coat_feed: bad arguments"); return FSR_ERR_ARG; }
  FSR_CUDA(cudaSetDevice(g->device));
  int rc = stage_q(g, Q, ldq, nsteps, g->stream);
  if (rc) return rc;
  if ((rc = fsr_coat_feed_dev(g, g->Qstage, ldq, nsteps, g->stream))) return rc;
  FSR_CUDA(cudaStreamSynchronize(g->stream));
  return FSR_OK;
}

// env [8][nros] (epsMax, epsMin, sigMax, sigMin, gammaMax, tauMax, vmeMax, vmsMax), summary [6][nros] (stress range, strain range,
// most popular angle [deg], angle spread [deg], biaxiality mean, biaxiality standard deviation), nbiax [nros]; any may be NULL
int fsr_coat_end(fsr_gages* g, double* env, double* summary, int* nbiax)
{
  if (!g || g->coat_nbin < 1) { set_error("fsr_coat_end: bad arguments (fsr_coat_begin first)"); return FSR_ERR_ARG; }
  FSR_CUDA(cudaSetDevice(g->device));
  const size_t nr = (size_t)std::max(g->nros, 1);
  double* out = nullptr;
  FSR_CUDA(cudaMalloc(&out, sizeof(double) * 6 * nr));
  if (g->nros > 0) {
    fsr::coat_finish_kernel<<<(g->nros + 127) / 128, 128, 0, g->stream>>>(g->nros, g->coat_nbin, g->coat_nval, g->coat_bin, g->coat_bsum,
                                                                        g->coat_nbiax, out);
    ++fsr::g_launches;
  }
  cudaError_t e = cudaSuccess;
  if (env) e = cudaMemcpyAsync(env, g->coat_env, sizeof(double) * 8 * g->nros, cudaMemcpyDeviceToHost, g->stream);
  if (e == cudaSuccess && summary) e = cudaMemcpyAsync(summary, out, sizeof(double) * 6 * g->nros, cudaMemcpyDeviceToHost, g->stream);
  if (e == cudaSuccess && nbiax) e = cudaMemcpyAsync(nbiax, g->coat_nbiax, sizeof(int) * g->nros, cudaMemcpyDeviceToHost, g->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(g->stream);
  cudaFree(out);
  if (e != cudaSuccess) { set_error("fsr_coat_end: %s", cudaGetErrorString(e)); return FSR_ERR_CUDA; }
  return FSR_OK;
}

}  // extern "C"
