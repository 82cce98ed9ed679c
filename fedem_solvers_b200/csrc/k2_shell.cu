// k2_shell.cu -- K2 for thin shells on sm_100a: ANDES quadrilateral (type 24) and triangle (23).
//
// The reference rebuilds, for every element and EVERY time step, the rigid-body projector, the
// element axes and eight strain-displacement matrices (quad: STR24 -> pMatStiff -> STR22a,
// src/vpmStress/elStressModule.f90:1005-1078,738-849, src/Femlib/pmatStiff.f90:23-127) or two
// hybrid-element flexibility inverses (triangle: STR23 -> FTSA31/FTSA32,
// elStressModule.f90:901-999).  All of that is time-invariant and linear in the element
// displacement vector, so this file
//   (1) builds, once per part, one stress operator per element:  sigma[ncmp*8] = S_e . v_e,
//       rows ordered component-major (xx at the 8 result points, then yy, then xy), already in
//       the reference's "globalized-X" output system (strainAndStressUtils.f90:437-481), stored
//       directly as FP64 MMA A-fragments;
//   (2) applies it to the step-batched displacements U[dof][t] with DMMA.8x8x4: one warp per
//       element, operator fragments resident in registers for the whole step tile, the element's
//       DOF rows of U gathered as 64-byte segments, von Mises (FFaTensorTransforms.C:33-36)
//       evaluated in the accumulator registers, running max/min envelope
//       (strainCoatModule.f90:159-166,410-420) fused into the same loop.
// HBM-bound by design: per element and step 8*nedof bytes of displacements in, 8*nstrp bytes of
// von Mises out; the operator (4.6 KB per quad) is read once per step tile.
#include "common.cuh"

namespace fsr {

// ------------------------------------------------------------------------------------------
// small device helpers
// ------------------------------------------------------------------------------------------
struct V3 { double x, y, z; };
__device__ __forceinline__ V3 vsub(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ V3 vcross(V3 a, V3 b)
{
  return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
__device__ __forceinline__ double vdot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ V3 vscale(V3 a, double s) { return {a.x * s, a.y * s, a.z * s}; }

// X-axis of the stress output system: global X projected onto the shell plane, or via global Y
// when the normal is (nearly) parallel to X (getGlobalizedX, strainAndStressUtils.f90:297-336).
__device__ bool globalized_x(V3 n, V3& v1)
{
  if (fabs(n.y) > 0.01 || fabs(n.z) > 0.01)
    v1 = {n.y * n.y + n.z * n.z, -n.x * n.y, -n.x * n.z};
  else {
    V3 v2 = {-n.y * n.x, n.x * n.x + n.z * n.z, -n.y * n.z};
    v1 = vcross(v2, n);
  }
  double l2 = vdot(v1, v1);
  if (l2 > kEpsDiv0 * kEpsDiv0) { v1 = vscale(v1, 1.0 / sqrt(l2)); return true; }
  v1 = {0, 0, 0};
  return false;
}

// cos/sin of the in-plane rotation from the element x-axis to the output x-axis
// (getShellStressTrans, strainAndStressUtils.f90:437-481)
__device__ bool stress_rotation(V3 ex, V3 ez, double& ca, double& sa)
{
  V3 xo;
  if (!globalized_x(ez, xo)) return false;
  V3 c = vcross(ex, xo);
  ca = vdot(xo, ex);
  double s = sqrt(vdot(c, c));
  sa = vdot(c, ez) >= 0.0 ? s : -s;
  return true;
}

// Congruence rotation of a 2-D symmetric tensor (t11, t22, t12) by the matrix the reference
// passes to tratensor: eX = (ca, -sa), eY = (sa, ca)  (FFaTensorTransforms.C:335-361).
__device__ __forceinline__ void rot2d(double& t11, double& t22, double& t12, double ca, double sa)
{
  double a11 = ca * t11 + sa * t12, a12 = ca * t12 + sa * t22;
  double a21 = -sa * t11 + ca * t12, a22 = -sa * t12 + ca * t22;
  t11 = a11 * ca + a12 * sa;
  t22 = a21 * (-sa) + a22 * ca;
  t12 = a11 * (-sa) + a12 * ca;
}

// where entry (row, col) of an operator with KT k-tiles lives in the A-fragment stream
__device__ __forceinline__ size_t frag_index(int row, int col, int KT)
{
  return ((size_t)((row >> 3) * KT + (col >> 2)) << 5) + ((row & 7) << 2) + (col & 3);
}

// ------------------------------------------------------------------------------------------
// Operator build: ANDES quadrilateral (type 24)
// ------------------------------------------------------------------------------------------
// One thread per element.  Column j of S_e is the stress response to the j-th unit nodal
// displacement after rigid-body projection, P e_j = e_j - Rt Rt' e_j - Rr G^-1 Rr' e_j, where
// Rt/Rr are the normalised rigid translation/rotation modes about the nodal centroid
// (pMatStiff.f90:59-125; only the rotational 3x3 Gram block G needs inverting).
__global__ void build_quad_ops_kernel(int nelt, const int* __restrict__ elem,
                                      const int* __restrict__ conn /* [nelt][4] 0-based nodes */,
                                      const double* __restrict__ xyz,
                                      const double* __restrict__ emod, const double* __restrict__ rny,
                                      const double* __restrict__ thk, double* __restrict__ Sfrag,
                                      unsigned char* __restrict__ failed, double* __restrict__ aux)
{
  const int KT = 6;
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nelt) return;
  int e = elem[i];
  double* S = Sfrag + (size_t)i * 3 * KT * 32;
  V3 X[4];
  for (int k = 0; k < 4; ++k) {
    int n = conn[i * 4 + k];
    X[k] = {xyz[3 * n], xyz[3 * n + 1], xyz[3 * n + 2]};
  }
  const double E = emod[e], nu = rny[e], t = thk[e];
  const double C11 = E / (1.0 - nu * nu), C12 = nu * C11, C33 = 0.5 * E / (1.0 + nu);
  aux[i * 4 + 0] = E; aux[i * 4 + 1] = nu; aux[i * 4 + 2] = t; aux[i * 4 + 3] = 0.0;
  bool ok = true;

  // element axes (getShellElementAxes, strainAndStressUtils.f90:382-424): normal from the
  // diagonals, x from edge 1-2 projected into the plane
  V3 ez = vcross(vsub(X[2], X[0]), vsub(X[3], X[1]));
  double l2 = vdot(ez, ez);
  if (l2 > kEpsDiv0 * kEpsDiv0) ez = vscale(ez, 1.0 / sqrt(l2)); else ok = false;
  V3 ex = vsub(X[1], X[0]);
  V3 ey = vcross(ez, ex);
  ex = vcross(ey, ez);
  l2 = vdot(ex, ex);
  if (l2 > kEpsDiv0 * kEpsDiv0) ex = vscale(ex, 1.0 / sqrt(l2)); else ok = false;
  ey = vcross(ez, ex);
  double ca = 1.0, sa = 0.0;
  if (ok) ok = stress_rotation(ex, ez, ca, sa);

  // rigid-body modes about the nodal centroid
  V3 cen = {0.25 * (X[0].x + X[1].x + X[2].x + X[3].x), 0.25 * (X[0].y + X[1].y + X[2].y + X[3].y),
            0.25 * (X[0].z + X[1].z + X[2].z + X[3].z)};
  V3 r[4];
  double nrm[3] = {4.0, 4.0, 4.0};  // |rotation mode|^2 = 4 (unit rotations) + lever arms
  for (int k = 0; k < 4; ++k) {
    r[k] = vsub(X[k], cen);
    nrm[0] += r[k].z * r[k].z + r[k].y * r[k].y;
    nrm[1] += r[k].z * r[k].z + r[k].x * r[k].x;
    nrm[2] += r[k].y * r[k].y + r[k].x * r[k].x;
  }
  const double in0 = 1.0 / sqrt(nrm[0]), in1 = 1.0 / sqrt(nrm[1]), in2 = 1.0 / sqrt(nrm[2]);
  // Gram matrix of the three normalised rotation modes and its inverse
  double G[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
  for (int k = 0; k < 4; ++k) {
    // rotation-mode rows at node k: translations (3) and rotations (3)
    double m0[6] = {0, -r[k].z * in0, r[k].y * in0, in0, 0, 0};
    double m1[6] = {r[k].z * in1, 0, -r[k].x * in1, 0, in1, 0};
    double m2[6] = {-r[k].y * in2, r[k].x * in2, 0, 0, 0, in2};
    for (int d = 0; d < 6; ++d) {
      G[0][0] += m0[d] * m0[d]; G[0][1] += m0[d] * m1[d]; G[0][2] += m0[d] * m2[d];
      G[1][1] += m1[d] * m1[d]; G[1][2] += m1[d] * m2[d]; G[2][2] += m2[d] * m2[d];
    }
  }
  G[1][0] = G[0][1]; G[2][0] = G[0][2]; G[2][1] = G[1][2];
  double det = G[0][0] * (G[1][1] * G[2][2] - G[2][1] * G[1][2]) -
               G[0][1] * (G[1][0] * G[2][2] - G[2][0] * G[1][2]) +
               G[0][2] * (G[1][0] * G[2][1] - G[2][0] * G[1][1]);
  double Gi[3][3];
  if (fabs(det) < kEpsDiv0) { ok = false; det = 1.0; }
  Gi[0][0] = (G[1][1] * G[2][2] - G[2][1] * G[1][2]) / det;
  Gi[0][1] = -(G[0][1] * G[2][2] - G[2][1] * G[0][2]) / det;
  Gi[0][2] = (G[0][1] * G[1][2] - G[1][1] * G[0][2]) / det;
  Gi[1][0] = -(G[1][0] * G[2][2] - G[2][0] * G[1][2]) / det;
  Gi[1][1] = (G[0][0] * G[2][2] - G[2][0] * G[0][2]) / det;
  Gi[1][2] = -(G[0][0] * G[1][2] - G[1][0] * G[0][2]) / det;
  Gi[2][0] = (G[1][0] * G[2][1] - G[2][0] * G[1][1]) / det;
  Gi[2][1] = -(G[0][0] * G[2][1] - G[2][0] * G[0][1]) / det;
  Gi[2][2] = (G[0][0] * G[1][1] - G[1][0] * G[0][1]) / det;

  // in-plane node coordinates relative to node 1 and the bilinear shape-function gradients at
  // the 2x2 Gauss points (StrainDispQuad4 / Quad4ShapeDer, strainAndStressUtils.f90:165-294);
  // gp index = 2*ixi + ieta like the reference's (i,j) loops
  double xl[4], yl[4];
  for (int k = 0; k < 4; ++k) {
    V3 d = vsub(X[k], X[0]);
    xl[k] = vdot(ex, d);
    yl[k] = vdot(ey, d);
  }
  xl[0] = 0.0; yl[0] = 0.0;
  const double gq = 1.0 / sqrt(3.0);
  double sx[4][4], sy[4][4];
  for (int gp = 0; gp < 4; ++gp) {
    double xi = (gp >> 1) ? gq : -gq, eta = (gp & 1) ? gq : -gq;
    double dxi[4] = {-(1.0 - eta) * 0.25, (1.0 - eta) * 0.25, (1.0 + eta) * 0.25, -(1.0 + eta) * 0.25};
    double det_[4] = {-(1.0 - xi) * 0.25, -(1.0 + xi) * 0.25, (1.0 + xi) * 0.25, (1.0 - xi) * 0.25};
    double j11 = 0, j12 = 0, j21 = 0, j22 = 0;
    for (int k = 0; k < 4; ++k) {
      j11 += dxi[k] * xl[k]; j12 += dxi[k] * yl[k];
      j21 += det_[k] * xl[k]; j22 += det_[k] * yl[k];
    }
    double dj = j11 * j22 - j21 * j12;
    double i11 = j22 / dj, i22 = j11 / dj, i12 = -j12 / dj, i21 = -j21 / dj;
    for (int k = 0; k < 4; ++k) {
      sx[gp][k] = i11 * dxi[k] + i12 * det_[k];
      sy[gp][k] = i21 * dxi[k] + i22 * det_[k];
    }
  }
  const double hh = (t + t + t + t) / 8.0;  // half thickness, sum(THK)/(2*nenod)
  const double f1 = 0.5 + 0.5 * sqrt(3.0), f2 = 0.5 - 0.5 * sqrt(3.0);
  // Gauss point closest to / farthest from node n: (iClose,jClose),(iFar,jFar) of STR22a
  const int gclose[4] = {0, 2, 3, 1}, gfar[4] = {3, 1, 0, 2};

  for (int col = 0; col < 24; ++col) {
    // ---- v = P e_col ----
    const int kn = col / 6, kd = col % 6;
    // coefficients of e_col on the 6 normalised rigid modes
    double ct[3] = {0, 0, 0}, cr[3];
    if (kd < 3) ct[kd] = 0.5;  // translation modes have entries 1/sqrt(4)
    {
      double m0[6] = {0, -r[kn].z * in0, r[kn].y * in0, in0, 0, 0};
      double m1[6] = {r[kn].z * in1, 0, -r[kn].x * in1, 0, in1, 0};
      double m2[6] = {-r[kn].y * in2, r[kn].x * in2, 0, 0, 0, in2};
      double b0 = m0[kd], b1 = m1[kd], b2 = m2[kd];
      // (Rr G^-1)[col,:] as the reference forms it: rsmat(3+j) = sum_k rmat(i,3+k)*subinv(k,j)
      cr[0] = b0 * Gi[0][0] + b1 * Gi[1][0] + b2 * Gi[2][0];
      cr[1] = b0 * Gi[0][1] + b1 * Gi[1][1] + b2 * Gi[2][1];
      cr[2] = b0 * Gi[0][2] + b1 * Gi[1][2] + b2 * Gi[2][2];
    }
    // local (element-axes) nodal translations u,v and rotations about x,y of the projected vector
    double ul[4], vl[4], tx[4], ty[4];
    for (int k = 0; k < 4; ++k) {
      double m0[6] = {0, -r[k].z * in0, r[k].y * in0, in0, 0, 0};
      double m1[6] = {r[k].z * in1, 0, -r[k].x * in1, 0, in1, 0};
      double m2[6] = {-r[k].y * in2, r[k].x * in2, 0, 0, 0, in2};
      double v[6];
      for (int d = 0; d < 6; ++d) {
        double tr = (d < 3) ? 0.5 * ct[d] : 0.0;
        v[d] = ((k == kn && d == kd) ? 1.0 : 0.0) - tr - (cr[0] * m0[d] + cr[1] * m1[d] + cr[2] * m2[d]);
      }
      ul[k] = ex.x * v[0] + ex.y * v[1] + ex.z * v[2];
      vl[k] = ey.x * v[0] + ey.y * v[1] + ey.z * v[2];
      tx[k] = ex.x * v[3] + ex.y * v[4] + ex.z * v[5];
      ty[k] = ey.x * v[3] + ey.y * v[4] + ey.z * v[5];
    }
    // ---- strains at the Gauss points, top (+hh) and bottom (-hh), rotated to output axes ----
    double et[4][3], eb[4][3];
    for (int gp = 0; gp < 4; ++gp) {
      double m0 = 0, m1 = 0, m2 = 0, k0 = 0, k1 = 0, k2 = 0;
      for (int k = 0; k < 4; ++k) {
        m0 += sx[gp][k] * ul[k];
        m1 += sy[gp][k] * vl[k];
        m2 += sy[gp][k] * ul[k] + sx[gp][k] * vl[k];
        k0 += sx[gp][k] * ty[k];
        k1 -= sy[gp][k] * tx[k];
        k2 += sy[gp][k] * ty[k] - sx[gp][k] * tx[k];
      }
      double a0 = m0 + hh * k0, a1 = m1 + hh * k1, a2 = 0.5 * (m2 + hh * k2);
      double b0 = m0 - hh * k0, b1 = m1 - hh * k1, b2 = 0.5 * (m2 - hh * k2);
      rot2d(a0, a1, a2, ca, sa);
      rot2d(b0, b1, b2, ca, sa);
      et[gp][0] = a0; et[gp][1] = a1; et[gp][2] = 2.0 * a2;
      eb[gp][0] = b0; eb[gp][1] = b1; eb[gp][2] = 2.0 * b2;
    }
    // ---- extrapolate to the nodes, sigma = C eps; rows: comp*8 + point (top 0-3, bottom 4-7) ----
    for (int n = 0; n < 4; ++n) {
      double e0 = f1 * et[gclose[n]][0] + f2 * et[gfar[n]][0];
      double e1 = f1 * et[gclose[n]][1] + f2 * et[gfar[n]][1];
      double e2 = f1 * et[gclose[n]][2] + f2 * et[gfar[n]][2];
      S[frag_index(0 + n, col, KT)] = ok ? C11 * e0 + C12 * e1 : 0.0;
      S[frag_index(8 + n, col, KT)] = ok ? C12 * e0 + C11 * e1 : 0.0;
      S[frag_index(16 + n, col, KT)] = ok ? C33 * e2 : 0.0;
      e0 = f1 * eb[gclose[n]][0] + f2 * eb[gfar[n]][0];
      e1 = f1 * eb[gclose[n]][1] + f2 * eb[gfar[n]][1];
      e2 = f1 * eb[gclose[n]][2] + f2 * eb[gfar[n]][2];
      S[frag_index(4 + n, col, KT)] = ok ? C11 * e0 + C12 * e1 : 0.0;
      S[frag_index(12 + n, col, KT)] = ok ? C12 * e0 + C11 * e1 : 0.0;
      S[frag_index(20 + n, col, KT)] = ok ? C33 * e2 : 0.0;
    }
  }
  failed[i] = ok ? 0 : 1;
}

// ------------------------------------------------------------------------------------------
// K2 apply: von Mises + envelope for shell families (3 m-tiles: xx, yy, xy at 8 points)
// ------------------------------------------------------------------------------------------
// The FP64 tensor pipe is the shared resource here (DMMA and scalar FP64 issue to the same pipe and
// a DMMA holds the dispatch port for its 16 cycles), so every non-DMMA instruction in the step loop
// costs wall time: the loop body is written for minimum instruction count -- row pointers bumped
// once per four tiles with immediate offsets in between, register double-buffering without moves,
// a Newton square root on MUFU.RSQ64H instead of the IEEE slow path, compare/select envelopes.

// sqrt(x) for x >= 0 to < 1 ulp-ish (two Newton steps on the 2^-22 hardware seed); 0 for x < 1e-290
__device__ __forceinline__ double sqrt_pos(double x)
{
  double r;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double y = x * r, h = 0.5 * r;
  double e = fma(-h, y, 0.5);
  y = fma(y, e, y);
  h = fma(h, e, h);
  e = fma(-y, y, x);
  y = fma(e, h, y);
  return x > 1.0e-290 ? y : 0.0;
}

template <int KT, bool WRITE_VM, bool GUARD>
__device__ __forceinline__ void shell_tile(const double (&a)[3][KT], const double (&b)[KT], double*& vmp0,
                                           double*& vmp1, size_t ld8, int t0, int nsteps, double& emax,
                                           double& emin)
{
  double c[3][2] = {{0, 0}, {0, 0}, {0, 0}};
#pragma unroll
  for (int j = 0; j < KT; ++j)
#pragma unroll
    for (int m = 0; m < 3; ++m) dmma884(c[m][0], c[m][1], a[m][j], b[j]);
  // von Mises, FFaTensorTransforms.C:33-36: sqrt(s11^2 + s22^2 - s11*s22 + 3*s12^2)
  double q0 = fma(c[2][0] * 3.0, c[2][0], fma(-c[0][0], c[1][0], fma(c[1][0], c[1][0], c[0][0] * c[0][0])));
  double q1 = fma(c[2][1] * 3.0, c[2][1], fma(-c[0][1], c[1][1], fma(c[1][1], c[1][1], c[0][1] * c[0][1])));
  double v0 = sqrt_pos(q0), v1 = sqrt_pos(q1);
  if (GUARD) {
    if (t0 < nsteps) {
      if (WRITE_VM) *vmp0 = v0;
      emax = v0 > emax ? v0 : emax;
      emin = v0 < emin ? v0 : emin;
    }
    if (t0 + 1 < nsteps) {
      if (WRITE_VM) *vmp1 = v1;
      emax = v1 > emax ? v1 : emax;
      emin = v1 < emin ? v1 : emin;
    }
  } else {
    if (WRITE_VM) { *vmp0 = v0; *vmp1 = v1; }
    const bool p = v0 > v1;
    const double hi = p ? v0 : v1, lo = p ? v1 : v0;
    emax = hi > emax ? hi : emax;
    emin = lo < emin ? lo : emin;
  }
  if (WRITE_VM) { vmp0 += ld8; vmp1 += ld8; }
}

template <int KT, bool WRITE_VM>
__global__ void __launch_bounds__(256, 2)
k2_shell_vm_kernel(const double* __restrict__ U, size_t ldu, int nsteps, int nsteps_pad,
                   const double* __restrict__ Sfrag, const int* __restrict__ edof,
                   const int* __restrict__ ptoff, const unsigned char* __restrict__ failed,
                   int nelt, int nstrp, double* __restrict__ vm, size_t ld_vm,
                   double* __restrict__ env_max, double* __restrict__ env_min)
{
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, t4 = lane & 3;
  const int i = blockIdx.x * (blockDim.x >> 5) + warp;
  if (i >= nelt) return;
  const bool live = g < nstrp;
  const size_t pt = (size_t)ptoff[i] + (live ? g : 0);

  if (failed[i]) {  // operator build failed: hugeVal results (stressRoutines.f90:264-268)
    if (live) {
      if (WRITE_VM)
        for (int t = t4; t < nsteps; t += 4) vm[(size_t)t * ld_vm + pt] = kHuge;
      if (t4 == 0 && nsteps > 0) { env_max[pt] = kHuge; if (kHuge < env_min[pt]) env_min[pt] = kHuge; }
    }
    return;
  }

  // operator fragments: resident for the whole step tile
  double a[3][KT];
  const double* sf = Sfrag + (size_t)i * 3 * KT * 32 + lane;
#pragma unroll
  for (int m = 0; m < 3; ++m)
#pragma unroll
    for (int j = 0; j < KT; ++j) a[m][j] = __ldg(sf + (size_t)(m * KT + j) * 32);

  // this lane's DOF row of U for each k-tile (B fragment: row k = t4, column n = g)
  const double* up[KT];
#pragma unroll
  for (int j = 0; j < KT; ++j) up[j] = U + (size_t)__ldg(edof + (size_t)i * KT * 4 + j * 4 + t4) * ldu + g;

  double emax = 0.0, emin = kHuge;  // neutral w.r.t. the stored envelope (max starts at 0)
  // lane owns point g at steps t0 = 8*tile + 2*t4 and t0 + 1; padded lanes (g >= nstrp) write to a
  // valid dummy location guarded below by `live`
  double* vmp0 = WRITE_VM ? vm + (size_t)(2 * t4) * ld_vm + pt : nullptr;
  double* vmp1 = WRITE_VM ? vmp0 + ld_vm : nullptr;
  const size_t ld8 = ld_vm * 8;

  const int ntiles = nsteps_pad >> 3;            // multiple of 8
  const int nfull = ((nsteps >> 3) >> 2) << 2;   // tiles (in groups of 4) with all 8 steps valid
  double b0[KT], b1[KT];
#pragma unroll
  for (int j = 0; j < KT; ++j) b0[j] = up[j][0];

  if (live) {
    int nt = 0;
    for (; nt < nfull; nt += 4) {
#pragma unroll
      for (int j = 0; j < KT; ++j) b1[j] = up[j][8];
      shell_tile<KT, WRITE_VM, false>(a, b0, vmp0, vmp1, ld8, 0, 0, emax, emin);
#pragma unroll
      for (int j = 0; j < KT; ++j) b0[j] = up[j][16];
      shell_tile<KT, WRITE_VM, false>(a, b1, vmp0, vmp1, ld8, 0, 0, emax, emin);
#pragma unroll
      for (int j = 0; j < KT; ++j) b1[j] = up[j][24];
      shell_tile<KT, WRITE_VM, false>(a, b0, vmp0, vmp1, ld8, 0, 0, emax, emin);
#pragma unroll
      for (int j = 0; j < KT; ++j) { up[j] += 32; b0[j] = up[j][0]; }  // U rows carry 64 doubles of slack
      shell_tile<KT, WRITE_VM, false>(a, b1, vmp0, vmp1, ld8, 0, 0, emax, emin);
    }
    for (; nt < ntiles && nt * 8 < nsteps; ++nt) {  // ragged tail: guarded stores
#pragma unroll
      for (int j = 0; j < KT; ++j) { up[j] += 8; b1[j] = up[j][0]; }
      shell_tile<KT, WRITE_VM, true>(a, b0, vmp0, vmp1, ld8, nt * 8 + 2 * t4, nsteps, emax, emin);
#pragma unroll
      for (int j = 0; j < KT; ++j) b0[j] = b1[j];
    }
  } else {
    // lanes of padded result points (triangles: g = 6,7) still feed the MMAs
    int nt = 0;
    double dmax = 0.0, dmin = 0.0;
    double *d0 = nullptr, *d1 = nullptr;
    for (; nt < ntiles && nt * 8 < nsteps; ++nt) {
#pragma unroll
      for (int j = 0; j < KT; ++j) { up[j] += 8; b1[j] = up[j][0]; }
      shell_tile<KT, false, false>(a, b0, d0, d1, 0, 0, 0, dmax, dmin);
#pragma unroll
      for (int j = 0; j < KT; ++j) b0[j] = b1[j];
    }
  }
  // combine the four lanes that share a result point, then fold into the stored envelope
  emax = fmax(emax, __shfl_xor_sync(0xffffffffu, emax, 1));
  emin = fmin(emin, __shfl_xor_sync(0xffffffffu, emin, 1));
  emax = fmax(emax, __shfl_xor_sync(0xffffffffu, emax, 2));
  emin = fmin(emin, __shfl_xor_sync(0xffffffffu, emin, 2));
  if (live && t4 == 0 && nsteps > 0) {
    if (emax > env_max[pt]) env_max[pt] = emax;
    if (emin < env_min[pt]) env_min[pt] = emin;
  }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
static int upload_family(FamilyData& f, const std::vector<int>& elem, const std::vector<int>& conn,
                         const std::vector<int>& edof, const std::vector<int>& ptoff, int naux,
                         int** d_conn, cudaStream_t s)
{
  f.nelt = (int)elem.size();
  f.naux = naux;
  if (f.nelt == 0) return FSR_OK;
  FSR_CUDA(cudaMalloc(&f.elem, sizeof(int) * elem.size()));
  FSR_CUDA(cudaMalloc(&f.edof, sizeof(int) * edof.size()));
  FSR_CUDA(cudaMalloc(&f.ptoff, sizeof(int) * ptoff.size()));
  FSR_CUDA(cudaMalloc(&f.failed, f.nelt));
  FSR_CUDA(cudaMalloc(&f.Sfrag, sizeof(double) * (size_t)f.nelt * f.MT * f.KT * 32));
  FSR_CUDA(cudaMalloc(&f.aux, sizeof(double) * (size_t)f.nelt * naux));
  FSR_CUDA(cudaMalloc(d_conn, sizeof(int) * conn.size()));
  FSR_CUDA(cudaMemcpyAsync(f.elem, elem.data(), sizeof(int) * elem.size(), cudaMemcpyHostToDevice, s));
  FSR_CUDA(cudaMemcpyAsync(f.edof, edof.data(), sizeof(int) * edof.size(), cudaMemcpyHostToDevice, s));
  FSR_CUDA(cudaMemcpyAsync(f.ptoff, ptoff.data(), sizeof(int) * ptoff.size(), cudaMemcpyHostToDevice, s));
  FSR_CUDA(cudaMemcpyAsync(*d_conn, conn.data(), sizeof(int) * conn.size(), cudaMemcpyHostToDevice, s));
  FSR_CUDA(cudaMemsetAsync(f.Sfrag, 0, sizeof(double) * (size_t)f.nelt * f.MT * f.KT * 32, s));
  return FSR_OK;
}

// Collects the active elements of one type: connectivity, the U row of every element DOF
// (extractEV, elStressModule.f90:84-95: min(nodal DOFs, nndof) per node), result-point offsets.
static int gather_family(const fsr_part* p, const fsr_sam* sam, const fsr_elmdata* elm, int type,
                         int nenod, int nndof, int KT, std::vector<int>& elem,
                         std::vector<int>& conn, std::vector<int>& edof, std::vector<int>& ptoff)
{
  for (int e = 0; e < sam->nel; ++e) {
    if (sam->melcon[e] != type) continue;
    if (elm->elmid && elm->elmid[e] < 1) continue;
    int ip0 = sam->mpmnpc[e] - 1, nn = sam->mpmnpc[e + 1] - sam->mpmnpc[e];
    if (nn != nenod) { set_error("element %d of type %d has %d nodes, expected %d", e + 1, type, nn, nenod); return FSR_ERR_ARG; }
    elem.push_back(e);
    ptoff.push_back(p->ptoff_host[e]);
    size_t base = edof.size();
    edof.resize(base + (size_t)KT * 4, 0);
    for (int k = 0; k < nenod; ++k) {
      int n = sam->mmnpc[ip0 + k] - 1;
      if (n < 0 || n >= sam->nnod) { set_error("element %d: node index out of range", e + 1); return FSR_ERR_ARG; }
      conn.push_back(n);
      int js = sam->madof[n] - 1, nd = sam->madof[n + 1] - sam->madof[n];
      if (nd < nndof) { set_error("element %d: node %d has %d DOFs, element needs %d", e + 1, n + 1, nd, nndof); return FSR_ERR_ARG; }
      for (int d = 0; d < nndof; ++d) edof[base + (size_t)k * nndof + d] = js + d;
    }
  }
  return FSR_OK;
}

int build_shell_operators(fsr_part* p, const fsr_sam* sam, const fsr_elmdata* elm)
{
  cudaStream_t s = p->stream;
  // ---- quads (type 24) ----
  {
    FamilyData& f = p->fam[FAM_QUAD];
    f.nenod = 4; f.nndof = 6; f.nstrp = 8; f.ncmp = 3; f.MT = 3; f.KT = 6;
    std::vector<int> elem, conn, edof, ptoff;
    int rc = gather_family(p, sam, elm, 24, 4, 6, f.KT, elem, conn, edof, ptoff);
    if (rc) return rc;
    int* d_conn = nullptr;
    rc = upload_family(f, elem, conn, edof, ptoff, 4, &d_conn, s);
    if (rc) return rc;
    if (f.nelt > 0) {
      build_quad_ops_kernel<<<(f.nelt + 63) / 64, 64, 0, s>>>(f.nelt, f.elem, d_conn, p->xyz, p->emod,
                                                            p->rny, p->thk, f.Sfrag, f.failed, f.aux);
      FSR_LAUNCH_CHECK();
      FSR_CUDA(cudaStreamSynchronize(s));
      cudaFree(d_conn);
    }
  }
  return FSR_OK;
}

int launch_k2_shell_vm(fsr_part* p, int nsteps, int nsteps_pad, double* vm_dev, size_t ld_vm, cudaStream_t s)
{
  const int warps = 8;
  {
    FamilyData& f = p->fam[FAM_QUAD];
    if (f.nelt > 0) {
      if (vm_dev)
        k2_shell_vm_kernel<6, true><<<(f.nelt + warps - 1) / warps, warps * 32, 0, s>>>(
            p->U, (size_t)p->step_tile, nsteps, nsteps_pad, f.Sfrag, f.edof, f.ptoff, f.failed, f.nelt,
            f.nstrp, vm_dev, ld_vm, p->env_max, p->env_min);
      else
        k2_shell_vm_kernel<6, false><<<(f.nelt + warps - 1) / warps, warps * 32, 0, s>>>(
            p->U, (size_t)p->step_tile, nsteps, nsteps_pad, f.Sfrag, f.edof, f.ptoff, f.failed, f.nelt,
            f.nstrp, vm_dev, ld_vm, p->env_max, p->env_min);
      FSR_LAUNCH_CHECK();
    }
  }
  return FSR_OK;
}

}  // namespace fsr
