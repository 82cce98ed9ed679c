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
#include <cstdlib>

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
                                      unsigned char* __restrict__ failed, double* __restrict__ aux, int ngauss)
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
  // ngauss = 1 (-ffqStressForm 1 of the legacy FFQ shell): STR22a evaluates at the centroid only and copies it to every node
  const double gq = ngauss == 1 ? 0.0 : 1.0 / sqrt(3.0);
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
  const double f1 = ngauss == 1 ? 1.0 : 0.5 + 0.5 * sqrt(3.0), f2 = ngauss == 1 ? 0.0 : 0.5 - 0.5 * sqrt(3.0);
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
// Operator build: ANDES triangle (type 23)
// ------------------------------------------------------------------------------------------
// In-place inverse of a column-major N x N matrix: LU with partial pivoting, then the inverse
// from the factors -- the job DINV12 (src/Femlib/dinv12.f:9-25) gives to LAPACK DGETRF/DGETRI
// (an un-vendored system library in the reference; pivot-order rounding is O(1e-15) relative).
template <int N>
__device__ bool lu_invert(double* a)
{
  int piv[N];
  double inv[N * N], col[N];
  for (int k = 0; k < N; ++k) {
    int p = k;
    double big = fabs(a[k + N * k]);
    for (int i = k + 1; i < N; ++i)
      if (fabs(a[i + N * k]) > big) { big = fabs(a[i + N * k]); p = i; }
    piv[k] = p;
    if (a[p + N * k] == 0.0) return false;
    if (p != k)
      for (int j = 0; j < N; ++j) { double t = a[k + N * j]; a[k + N * j] = a[p + N * j]; a[p + N * j] = t; }
    for (int i = k + 1; i < N; ++i) a[i + N * k] /= a[k + N * k];
    for (int j = k + 1; j < N; ++j)
      for (int i = k + 1; i < N; ++i) a[i + N * j] -= a[i + N * k] * a[k + N * j];
  }
  for (int c = 0; c < N; ++c) {
    for (int i = 0; i < N; ++i) col[i] = (i == c) ? 1.0 : 0.0;
    for (int k = 0; k < N; ++k)
      if (piv[k] != k) { double t = col[k]; col[k] = col[piv[k]]; col[piv[k]] = t; }
    for (int i = 0; i < N; ++i)
      for (int k = 0; k < i; ++k) col[i] -= a[i + N * k] * col[k];
    for (int i = N - 1; i >= 0; --i) {
      for (int k = i + 1; k < N; ++k) col[i] -= a[i + N * k] * col[k];
      col[i] /= a[i + N * i];
    }
    for (int i = 0; i < N; ++i) inv[i + N * c] = col[i];
  }
  for (int k = 0; k < N * N; ++k) a[k] = inv[k];
  return true;
}

// Membrane "kappa" matrix of the hybrid triangle, AKM(7 x 9) = F^-1 A with the IOP = 1 edge
// displacement variant (HLST31, src/Femlib/hlst.f:102-243).  x, y: local corner coordinates,
// Ei: inverse of the 3x3 constitutive matrix (column-major), thk: mean thickness.
__device__ bool tri_membrane_kappa(double* AKM, const double* Ei, const double* x, const double* y, double thk)
{
  const int nxt[3] = {1, 2, 0};
  double c[3], s[3], sl[3], xl[3], yl[3];
  double area = 0.5 * (x[0] * y[1] + x[1] * y[2] + x[2] * y[0] - x[0] * y[2] - x[1] * y[0] - x[2] * y[1]);
  if (area < 0.0) return false;
  for (int i = 0; i < 3; ++i) {
    const int j = nxt[i];
    sl[i] = sqrt((x[j] - x[i]) * (x[j] - x[i]) + (y[j] - y[i]) * (y[j] - y[i]));
    s[i] = (x[i] - x[j]) / sl[i];
    c[i] = (y[j] - y[i]) / sl[i];
  }
  const double x0 = (x[0] + x[1] + x[2]) / 3., y0 = (y[0] + y[1] + y[2]) / 3.;
  for (int i = 0; i < 3; ++i) { xl[i] = x[i] - x0; yl[i] = y[i] - y0; }
  double f = area / (12. * thk);
  const double P20 = f * (xl[0] * xl[0] + xl[1] * xl[1] + xl[2] * xl[2]);
  const double P11 = f * (xl[0] * yl[0] + xl[1] * yl[1] + xl[2] * yl[2]);
  const double P02 = f * (yl[0] * yl[0] + yl[1] * yl[1] + yl[2] * yl[2]);
#define EI(i, j) Ei[(i - 1) + 3 * (j - 1)]
#define FM(i, j) F7[(i - 1) + 7 * (j - 1)]
#define AM(i, j) A[(i - 1) + 7 * (j - 1)]
  double F7[49], A[63];
  for (int k = 0; k < 49; ++k) F7[k] = 0.0;
  for (int k = 0; k < 63; ++k) A[k] = 0.0;
  f = area / thk;
  FM(1, 1) = f * EI(1, 1); FM(1, 2) = f * EI(1, 2); FM(2, 2) = f * EI(2, 2);
  FM(1, 3) = f * EI(1, 3); FM(2, 3) = f * EI(2, 3); FM(3, 3) = f * EI(3, 3);
  FM(4, 4) = EI(1, 1) * P20 - 2. * EI(1, 3) * P11 + EI(3, 3) * P02;
  FM(4, 5) = EI(1, 2) * P20 - EI(2, 3) * P11;
  FM(5, 5) = EI(2, 2) * P20;
  FM(4, 6) = EI(1, 1) * P11 - EI(1, 3) * P02;
  FM(5, 6) = EI(1, 2) * P11;
  FM(6, 6) = EI(1, 1) * P02;
  FM(4, 7) = -EI(1, 3) * P20 + (EI(1, 2) + EI(3, 3)) * P11 - EI(2, 3) * P02;
  FM(5, 7) = EI(2, 2) * P11 - EI(2, 3) * P20;
  FM(6, 7) = EI(1, 2) * P02 - EI(1, 3) * P11;
  FM(7, 7) = EI(2, 2) * P02 - 2. * EI(2, 3) * P11 + EI(3, 3) * P20;
  for (int i = 1; i <= 7; ++i)
    for (int j = i; j <= 7; ++j) FM(j, i) = FM(i, j);
  if (!lu_invert<7>(F7)) return false;
  for (int K = 0; K < 3; ++K) {
    const double ck = c[K], sk = s[K], l = sl[K];
    const double C2 = ck * ck, S2 = sk * sk, SC = sk * ck;
    const double CX = ck * xl[K], CY = ck * yl[K], SX = sk * xl[K], SY = sk * yl[K];
    const double l2 = l * l, A1 = l / 2.;
    for (int end = 0; end < 2; ++end) {
      // end 0: the edge's first node (K), end 1: its second node
      const int J = 3 * (end == 0 ? K : nxt[K]) + 1;
      const double sgn = end == 0 ? 1.0 : -1.0;
      const double A4 = -sgn * ck * l2 / 12., A5 = -sgn * sk * l2 / 12.;
      const double B4 = end == 0 ? -ck * l2 * l / 30. : ck * l2 * l / 20.;
      const double B5 = end == 0 ? -sk * l2 * l / 30. : sk * l2 * l / 20.;
      const double B1 = end == 0 ? l2 / 6. : l2 / 3., B2 = 0., B3 = B1;
      AM(1, J) += A1 * ck;
      AM(3, J) += A1 * sk;
      AM(4, J) += A1 * (CX - SY) - 2. * B1 * SC - B2 * C2;
      AM(5, J) += -B2 * S2;
      AM(6, J) += A1 * CY + B1 * C2;
      AM(7, J) += -A1 * SX + B1 * S2 + 2. * B2 * SC;
      AM(2, J + 1) += A1 * sk;
      AM(3, J + 1) += A1 * ck;
      AM(4, J + 1) += -A1 * CY - 2. * B2 * SC - B3 * C2;
      AM(5, J + 1) += A1 * SX - B3 * S2;
      AM(6, J + 1) += B2 * C2;
      AM(7, J + 1) += A1 * (SY - CX) + B2 * S2 + 2. * B3 * SC;
      AM(1, J + 2) += A4 * ck;
      AM(2, J + 2) += A5 * sk;
      AM(3, J + 2) += A4 * sk + A5 * ck;
      AM(4, J + 2) += A4 * (CX - SY) - A5 * CY - 2. * B4 * SC - B5 * C2;
      AM(5, J + 2) += A5 * SX - B5 * S2;
      AM(6, J + 2) += A4 * CY + B4 * C2;
      AM(7, J + 2) += -A4 * SX + A5 * (SY - CX) + B4 * S2 + 2. * B5 * SC;
    }
  }
  for (int i = 1; i <= 7; ++i)
    for (int j = 1; j <= 9; ++j) {
      double acc = FM(i, 1) * AM(1, j);
      for (int k = 2; k <= 7; ++k) acc += FM(i, k) * AM(k, j);
      AKM[(i - 1) + 7 * (j - 1)] = acc;
    }
#undef FM
#undef AM
  return true;
}

// Bending "kappa" matrix AKB(9 x 9) = Fb^-1 G T (TEBA31, src/Femlib/nyteba.f:7-331).  The 7-point
// rule is kept in the REAL*4 precision of the reference's DATA statements (nyteba.f:35-43).
__device__ bool tri_bending_kappa(double* AKB, const double* Ei, const double* x, const double* y, const double* th)
{
  const int nxt[3] = {1, 2, 0};
  const double za = (double)0.33333333f, zb = (double)0.05971587f, zc = (double)0.47014206f,
               zd = (double)0.79742699f, ze = (double)0.10128651f;
  const double Z1[7] = {za, zb, zc, zc, zd, ze, ze}, Z2[7] = {za, zc, zb, zc, ze, zd, ze},
               Z3[7] = {za, zc, zc, zb, ze, ze, zd};
  const double wa = (double)0.225f, wb = (double)0.13239415f, wc = (double)0.12593918f;
  const double W[7] = {wa, wb, wb, wb, wc, wc, wc};
  double c[3], s[3], sl[3];
  const double area = 0.5 * (x[0] * y[1] + x[1] * y[2] + x[2] * y[0] - x[0] * y[2] - x[1] * y[0] - x[2] * y[1]);
  if (area <= 0.0) return false;
  for (int i = 0; i < 3; ++i) {
    const int j = nxt[i];
    sl[i] = sqrt((x[j] - x[i]) * (x[j] - x[i]) + (y[j] - y[i]) * (y[j] - y[i]));
    s[i] = (x[i] - x[j]) / sl[i];
    c[i] = (y[j] - y[i]) / sl[i];
  }
  const double RL11 = 0.5 * (y[1] - y[2]) / area, RL12 = 0.5 * (y[2] - y[0]) / area;
  const double RL21 = 0.5 * (x[2] - x[1]) / area, RL22 = 0.5 * (x[0] - x[2]) / area;
  double Bq[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (int k = 0; k < 7; ++k) {
    const double t = th[0] * Z1[k] + th[1] * Z2[k] + th[2] * Z3[k];
    if (t <= 0.0) return false;
    const double f = 12. * area * W[k] / (t * t * t);
    const double z[3] = {Z1[k], Z2[k], Z3[k]};
    for (int i = 0; i < 3; ++i)
      for (int j = i; j < 3; ++j) Bq[i + 3 * j] += f * z[i] * z[j];
  }
  for (int i = 0; i < 3; ++i)
    for (int j = i; j < 3; ++j) Bq[j + 3 * i] = Bq[i + 3 * j];
  double EK[81], G[108], T[108], GT[81];
#define EKM(i, j) EK[(i - 1) + 9 * (j - 1)]
#define GM(i, j) G[(i - 1) + 9 * (j - 1)]
#define TM(i, j) T[(i - 1) + 12 * (j - 1)]
  for (int II = 1; II <= 3; ++II)
    for (int JJ = II; JJ <= 3; ++JJ)
      for (int i = 1; i <= 3; ++i)
        for (int j = 1; j <= 3; ++j) EKM(3 * II - 3 + i, 3 * JJ - 3 + j) = EI(II, JJ) * Bq[(i - 1) + 3 * (j - 1)];
  for (int i = 1; i <= 9; ++i)
    for (int j = i; j <= 9; ++j) EKM(j, i) = EKM(i, j);
  if (!lu_invert<9>(EK)) return false;
  for (int k = 0; k < 108; ++k) { G[k] = 0.0; T[k] = 0.0; }
  for (int I = 1; I <= 3; ++I) {
    const int K = nxt[nxt[I - 1]] + 1;
    GM(I, I) = s[K - 1] * c[K - 1] - s[I - 1] * c[I - 1];
    GM(I + 3, I) = -GM(I, I);
    GM(I + 6, I) = c[I - 1] * c[I - 1] - s[I - 1] * s[I - 1] - c[K - 1] * c[K - 1] + s[K - 1] * s[K - 1];
  }
  for (int I = 1; I <= 3; ++I) {
    const int J = I + 3;
    const double ci = c[I - 1], si = s[I - 1];
    const double SS = si * si * ci, CC = ci * ci * si, CS = ci * ci - si * si;
    GM(1, J) = (ci + SS) * RL11 - CC * RL21;
    GM(2, J) = (ci + SS) * RL12 - CC * RL22;
    GM(3, J) = -(ci + SS) * (RL11 + RL12) + CC * (RL21 + RL22);
    GM(4, J) = (si + CC) * RL21 - SS * RL11;
    GM(5, J) = (si + CC) * RL22 - SS * RL12;
    GM(6, J) = -(si + CC) * (RL21 + RL22) + SS * (RL11 + RL12);
    GM(7, J) = si * RL11 + ci * RL21 - CS * (si * RL11 - ci * RL21);
    GM(8, J) = ci * RL22 + si * RL12 - CS * (si * RL12 - ci * RL22);
    GM(9, J) = -si * (RL11 + RL12) - ci * (RL21 + RL22);
    GM(9, J) = GM(9, J) + CS * (si * (RL11 + RL12) - ci * (RL21 + RL22));
  }
  for (int I = 1; I <= 3; ++I) {
    const int J = nxt[I - 1] + 1, N = 2 * I + 5;
    const double ci = c[I - 1], si = s[I - 1];
    GM(I, N) = ci * ci; GM(I + 3, N) = si * si; GM(I + 6, N) = 2. * si * ci;
    GM(J, N + 1) = ci * ci; GM(J + 3, N + 1) = si * si; GM(J + 6, N + 1) = 2. * si * ci;
  }
  for (int I = 1; I <= 3; ++I) {
    const int J = nxt[I - 1] + 1;
    const double l = sl[I - 1], SS = s[I - 1] * l, CC = c[I - 1] * l;
    TM(I, 3 * I - 2) = 1.;
    TM(I + 3, 3 * I - 2) = l / 2.;
    TM(I + 3, 3 * I - 1) = -l * SS / 12.;
    TM(I + 3, 3 * I) = l * CC / 12.;
    TM(I + 3, 3 * J - 2) = l / 2.;
    TM(I + 3, 3 * J - 1) = l * SS / 12.;
    TM(I + 3, 3 * J) = -l * CC / 12.;
    TM(2 * I + 5, 3 * I - 1) = -CC / 3.;
    TM(2 * I + 5, 3 * I) = -SS / 3.;
    TM(2 * I + 5, 3 * J - 1) = -CC / 6.;
    TM(2 * I + 5, 3 * J) = -SS / 6.;
    TM(2 * I + 6, 3 * I - 1) = -CC / 6.;
    TM(2 * I + 6, 3 * I) = -SS / 6.;
    TM(2 * I + 6, 3 * J - 1) = -CC / 3.;
    TM(2 * I + 6, 3 * J) = -SS / 3.;
  }
  // (w, dw/dx, dw/dy) -> (w, theta_x, theta_y): swap the two rotation columns with a sign change
  for (int I = 1; I <= 3; ++I) {
    const int J = 3 * I - 1, K = J + 1;
    for (int M = 1; M <= 12; ++M) { const double f = TM(M, J); TM(M, J) = TM(M, K); TM(M, K) = -f; }
  }
  for (int I = 1; I <= 9; ++I)
    for (int J = 1; J <= 9; ++J) {
      double acc = 0.;
      for (int K = 1; K <= 12; ++K) acc += GM(I, K) * TM(K, J);
      GT[(I - 1) + 9 * (J - 1)] = acc;
    }
  for (int I = 1; I <= 9; ++I)
    for (int J = 1; J <= 9; ++J) {
      double acc = 0.;
      for (int K = 1; K <= 9; ++K) acc += EKM(I, K) * GT[(K - 1) + 9 * (J - 1)];
      AKB[(I - 1) + 9 * (J - 1)] = acc;
    }
#undef EKM
#undef GM
#undef TM
#undef EI
  return true;
}

// ---- legacy FFT3 shell with -fftStressForm 0 / 2 (STR21 -> FTS31 / FTS32, fts.f:7-292) ---------------------------------
// The membrane part is the Bergan / Felippa triangle of tmrf.f: TMRF31 / SM3MH (:7-55,151-303) deliver the higher-order
// strain-displacement relation HH(3,9) through a 9 x 9 Crout factorisation with implicit row scaling (LUFACT :304-388, LUSOLV
// :389-473 in its row-storage form), TMRF32 (:474-634) the centroid stress matrix DM (L'/A + BH HH).
__device__ static bool tmrf_lufact9(double* A, int* perm, double* V)
{
#define A_(i, j) A[(i) + 9 * (j)]
  for (int i = 0; i < 9; ++i) {
    double y = 0.0;
    for (int j = 0; j < 9; ++j) y += A_(i, j) * A_(i, j);
    V[i] = y > 0.0 ? sqrt(1.0 / y) : 0.0;
  }
  for (int k = 0; k < 9; ++k) {
    perm[k] = k;
    if (V[k] <= 0.0) continue;
    int l = k;
    double x = 0.0;
    for (int i = k; i < 9; ++i) {
      double y = 0.0;
      for (int j = 0; j < k; ++j) y += A_(i, j) * A_(j, k);
      A_(i, k) -= y;
      y = fabs(V[i] * A_(i, k));
      if (y > x) { x = y; l = i; }
    }
    if (l != k) {
      for (int j = 0; j < 9; ++j) { const double y = A_(k, j); A_(k, j) = A_(l, j); A_(l, j) = y; }
      V[l] = V[k];
      perm[k] = l;
    }
    if (x <= 2.0e-16) return false;   // MACTOL
    x = 1.0 / A_(k, k);
    A_(k, k) = x;
    for (int j = k + 1; j < 9; ++j) {
      double y = 0.0;
      for (int i = 0; i < k; ++i) y += A_(k, i) * A_(i, j);
      A_(k, j) = (A_(k, j) - y) * x;
    }
  }
  return true;
}

// HH [3][9] (row r, column j at HH[r + 3 j]; columns u1 v1 u2 v2 u3 v3 th1 th2 th3) and the centroid matrix SMM[3][9] of FTS32,
// which is handed the plane stress matrix Dm where TMRF32 expects the membrane rigidity and adds the columns of HH to those of
// the lumping matrix (node order u1 v1 th1 ..) as they come -- both as in the reference.
__device__ static bool tri_legacy_membrane(double (*SMM)[9], const double* Dm /* 3x3 column-major */, const double* X, const double* Y,
                                           double alpha)
{
  const double area2 = (Y[1] - Y[0]) * (X[0] - X[2]) - (X[1] - X[0]) * (Y[0] - Y[2]);
  if (area2 <= 1.0e-16) return false;
  const double x0 = (X[0] + X[1] + X[2]) / 3.0, y0 = (Y[0] + Y[1] + Y[2]) / 3.0, area = 0.5 * area2, c = 1. / sqrt(area);
  double xc[3], yc[3], xm[3], ym[3], GT[81], HH[27], T[9], BH[9];
  int perm[9];
  for (int i = 0; i < 3; ++i) { xc[i] = c * (X[i] - x0); yc[i] = c * (Y[i] - y0); }
  xm[0] = 0.5 * (xc[1] + xc[2]); xm[1] = 0.5 * (xc[2] + xc[0]); xm[2] = 0.5 * (xc[0] + xc[1]);
  ym[0] = 0.5 * (yc[1] + yc[2]); ym[1] = 0.5 * (yc[2] + yc[0]); ym[2] = 0.5 * (yc[0] + yc[1]);
  for (int i = 0; i < 81; ++i) GT[i] = 0.0;
  for (int i = 0; i < 27; ++i) HH[i] = 0.0;
#define G_(i, j) GT[(i) + 9 * (j)]
  for (int j = 0; j < 3; ++j) {
    const double dx = xm[j] - xc[j], dy = ym[j] - yc[j], dl = sqrt(dx * dx + dy * dy), cj = dx / dl, sj = dy / dl;
    const double a1 = -0.5 * sj * (cj * cj), a2 = 0.5 * (cj * cj * cj), b2 = -0.5 * (sj * sj * sj), b3 = 0.5 * (sj * sj) * cj;
    const double a3 = -(b2 + a1 + a1), b1 = -(b3 + b3 + a2);
    G_(0, 2 * j) = 1.; G_(1, 2 * j + 1) = 1.;
    G_(2, 2 * j) = -yc[j]; G_(2, 2 * j + 1) = xc[j]; G_(2, j + 6) = c;
    G_(3, 2 * j) = xc[j]; G_(5, 2 * j) = yc[j]; G_(4, 2 * j + 1) = yc[j]; G_(5, 2 * j + 1) = xc[j];
    HH[j + 3 * (j + 6)] = 1.;
    for (int i = 0; i < 3; ++i) {
      const double xi = xc[i], yi = yc[i];
      G_(j + 6, 2 * i) = a1 * xi * xi + 2. * a2 * xi * yi + a3 * yi * yi;
      G_(j + 6, 2 * i + 1) = b1 * xi * xi + 2. * b2 * xi * yi + b3 * yi * yi;
      G_(j + 6, i + 6) = -c * (cj * xi + sj * yi);
    }
    BH[0 + 3 * j] = c * (2 * a1 * xc[j] + a2 * yc[j]);
    BH[1 + 3 * j] = c * (b2 * xc[j] + 2 * b3 * yc[j]);
    BH[2 + 3 * j] = c * (-4 * b3 * xc[j] - 4 * a1 * yc[j]);
  }
#undef G_
  if (!tmrf_lufact9(GT, perm, T)) return false;
  // LUSOLV(GT,9,9,IPERM,HH,3,-3): the three right hand sides are the rows of HH
  for (int i = 0; i < 9; ++i) {
    const int k = perm[i];
    if (k != i) for (int r = 0; r < 3; ++r) { const double t = HH[r + 3 * i]; HH[r + 3 * i] = HH[r + 3 * k]; HH[r + 3 * k] = t; }
  }
  for (int r = 0; r < 3; ++r) {
    HH[r] = HH[r] * GT[0];
    for (int i = 0; i < 8; ++i) {
      double sum = 0.0;
      for (int k = 0; k <= i; ++k) sum -= GT[(i + 1) + 9 * k] * HH[r + 3 * k];
      HH[r + 3 * (i + 1)] = (HH[r + 3 * (i + 1)] + sum) * GT[(i + 1) + 9 * (i + 1)];
    }
    for (int i = 7; i >= 0; --i) {
      double sum = 0.0;
      for (int k = 1; k <= 8 - i; ++k) sum -= GT[i + 9 * (i + k)] * HH[r + 3 * (i + k)];
      HH[r + 3 * i] += sum;
    }
  }
#undef A_
  // TMRF32: lumping matrix L (9 x 3) in node order, TM = L'/A + BH HH, SM = Dm TM
  const double x21 = X[1] - X[0], x12 = -x21, x32 = X[2] - X[1], x23 = -x32, x13 = X[0] - X[2], x31 = -x13;
  const double y21 = Y[1] - Y[0], y12 = -y21, y32 = Y[2] - Y[1], y23 = -y32, y13 = Y[0] - Y[2], y31 = -y13;
  double L[9][3] = {{0.5 * y23, 0.0, 0.5 * x32}, {0.0, 0.5 * x32, 0.5 * y23}, {0.0, 0.0, 0.0},
                    {0.5 * y31, 0.0, 0.5 * x13}, {0.0, 0.5 * x13, 0.5 * y31}, {0.0, 0.0, 0.0},
                    {0.5 * y12, 0.0, 0.5 * x21}, {0.0, 0.5 * x21, 0.5 * y12}, {0.0, 0.0, 0.0}};
  if (alpha > 0.0) {
    L[2][0] = 0.5 * (y23 * (y13 - y21) * alpha / 6); L[2][1] = 0.5 * (x32 * (x31 - x12) * alpha / 6); L[2][2] = 0.5 * ((x31 * y13 - x12 * y21) * alpha / 3);
    L[5][0] = 0.5 * (y31 * (y21 - y32) * alpha / 6); L[5][1] = 0.5 * (x13 * (x12 - x23) * alpha / 6); L[5][2] = 0.5 * ((x12 * y21 - x23 * y32) * alpha / 3);
    L[8][0] = 0.5 * (y12 * (y32 - y13) * alpha / 6); L[8][1] = 0.5 * (x21 * (x23 - x31) * alpha / 6); L[8][2] = 0.5 * ((x23 * y32 - x31 * y13) * alpha / 3);
  }
  for (int i = 0; i < 9; ++i) {
    double tm[3];
    for (int j = 0; j < 3; ++j) {
      double sh = 0.0;
      for (int k = 0; k < 3; ++k) sh += BH[j + 3 * k] * HH[k + 3 * i];
      tm[j] = (1. / area) * L[i][j] + sh;
    }
    for (int j = 0; j < 3; ++j) SMM[j][i] = Dm[j] * tm[0] + Dm[j + 3] * tm[1] + Dm[j + 6] * tm[2];
  }
  return true;
}

// One thread per triangle: STR23 (src/vpmStress/elStressModule.f90:901-999) as a 6-point x 3-component
// x 18-DOF operator.  FTSA31 (ftsa.f:53-104) gives the two kappa matrices, FTSA32 (:196-260) their
// centroid stress matrices with ZZ = 1./3. in REAL*4, FTS38 (fts.f:446-490) the split of the nodal
// vector into membrane (u, v, rz) and bending (w, rx, ry) parts in the DIRC30 axes
// (beamaux.f:209-260); top / bottom stresses = (N +- 6 M / t) / t at every node.
__global__ void build_tri_ops_kernel(int nelt, const int* __restrict__ elem, const int* __restrict__ conn,
                                     const double* __restrict__ xyz, const double* __restrict__ emod,
                                     const double* __restrict__ rny, const double* __restrict__ thk,
                                     double* __restrict__ Sfrag, unsigned char* __restrict__ failed,
                                     double* __restrict__ aux, int legacy)
{
  const int KT = 5;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nelt) return;
  const int e = elem[i];
  double* S = Sfrag + (size_t)i * 3 * KT * 32;
  V3 X[3];
  for (int k = 0; k < 3; ++k) {
    const int n = conn[i * 3 + k];
    X[k] = {xyz[3 * n], xyz[3 * n + 1], xyz[3 * n + 2]};
  }
  const double E = emod[e], nu = rny[e], t = thk[e];
  aux[i * 4 + 0] = E; aux[i * 4 + 1] = nu; aux[i * 4 + 2] = t; aux[i * 4 + 3] = 0.0;
  bool ok = true;
  // isoMat2D and its LU inverse (HLST31/TEBA31 invert the matrix they are given)
  const double C11 = E / (1.0 - nu * nu);
  double Ei[9] = {C11, nu * C11, 0., nu * C11, C11, 0., 0., 0., 0.5 * E / (1.0 + nu)};
  ok = lu_invert<3>(Ei);
  // local corner coordinates (ftsa.f:60-80)
  const V3 d21 = vsub(X[1], X[0]), d31 = vsub(X[2], X[0]);
  const double L21 = sqrt(vdot(d21, d21)), L31 = sqrt(vdot(d31, d31));
  const double cosg = vdot(d31, d21) / (L31 * L21);
  const double a1 = 1. - cosg * cosg;
  const double sing = a1 <= 0.0 ? 0. : sqrt(a1);
  const double xl[3] = {0., L21, L31 * cosg}, yl[3] = {0., 0., L31 * sing};
  const double th[3] = {t, t, t};
  double AKM[63], AKB[81], SMM[3][9], SMB[3][9];
  if (ok && !legacy) ok = tri_membrane_kappa(AKM, Ei, xl, yl, (t + t + t) / 3.);
  if (ok && legacy) {
    // legacy FFT3 formulation (-fftStressForm 0 / 2): membrane matrix of TMRF32 on the plane stress matrix E itself
    // (elStressModule.f90:589-591 passes E where FTS32 expects t * E)
    const double Dm[9] = {C11, nu * C11, 0., nu * C11, C11, 0., 0., 0., 0.5 * E / (1.0 + nu)};
    ok = tri_legacy_membrane(SMM, Dm, xl, yl, 1.5);
  }
  if (ok) ok = tri_bending_kappa(AKB, Ei, xl, yl, th);
  // output-system rotation from the triangle axes (x along 1->2, not projected)
  V3 ex = d21, ez = vcross(d21, d31);
  double l2 = vdot(ez, ez);
  if (l2 > kEpsDiv0 * kEpsDiv0) ez = vscale(ez, 1.0 / sqrt(l2)); else ok = false;
  l2 = vdot(ex, ex);
  if (l2 > kEpsDiv0 * kEpsDiv0) ex = vscale(ex, 1.0 / sqrt(l2)); else ok = false;
  double ca = 1.0, sa = 0.0;
  if (ok) ok = stress_rotation(ex, ez, ca, sa);
  if (!ok) { failed[i] = 1; return; }  // Sfrag was zeroed by the caller

  // centroid stress matrices (HLST32, hlst.f:330-352; TEBA32, nyteba.f:333-364), ZZ = REAL*4 1./3.
  // legacy FFT3 formulation: FTS32 with ZZ = 1/3 in double precision
  const double zz = legacy ? 1.0 / 3.0 : (double)(1.f / 3.f);
  const double x0 = (xl[0] + xl[1] + xl[2]) / 3., y0 = (yl[0] + yl[1] + yl[2]) / 3.;
  double rx = 0., ry = 0.;
  for (int k = 0; k < 3; ++k) { rx += (xl[k] - x0) * zz; ry += (yl[k] - y0) * zz; }
  for (int j = 0; j < 9; ++j) {
    if (!legacy) {
      const double* a = AKM + 7 * j;
      SMM[0][j] = a[0] + rx * a[3] + ry * a[5];
      SMM[1][j] = a[1] + rx * a[4] + ry * a[6];
      SMM[2][j] = a[2] - ry * a[3] - rx * a[6];
    }
    for (int r = 0; r < 3; ++r) SMB[r][j] = zz * AKB[3 * r + 9 * j] + zz * AKB[3 * r + 1 + 9 * j] + zz * AKB[3 * r + 2 + 9 * j];
  }
  // direction cosines (DIRC30): rows = local x (1->2), y, z; every row renormalised
  V3 cz = vcross(ex, d31);
  cz = vscale(cz, 1.0 / sqrt(vdot(cz, cz)));
  V3 cy = vcross(cz, ex);
  cy = vscale(cy, 1.0 / sqrt(vdot(cy, cy)));
  const double Cd[3][3] = {{ex.x, ex.y, ex.z}, {cy.x, cy.y, cy.z}, {cz.x, cz.y, cz.z}};
  for (int n = 0; n < 3; ++n)
    for (int d = 0; d < 6; ++d) {
      const int k = d % 3;
      double N[3], M[3];
      for (int r = 0; r < 3; ++r) {
        if (d < 3) {  // translation: (u, v) -> membrane, w -> bending
          N[r] = SMM[r][3 * n] * Cd[0][k] + SMM[r][3 * n + 1] * Cd[1][k];
          M[r] = SMB[r][3 * n] * Cd[2][k];
        } else {      // rotation: (rx, ry) -> bending, rz -> membrane
          M[r] = SMB[r][3 * n + 1] * Cd[0][k] + SMB[r][3 * n + 2] * Cd[1][k];
          N[r] = SMM[r][3 * n + 2] * Cd[2][k];
        }
      }
      rot2d(N[0], N[1], N[2], ca, sa);
      rot2d(M[0], M[1], M[2], ca, sa);
      const int col = 6 * n + d;
      for (int r = 0; r < 3; ++r) {
        const double top = (N[r] + M[r] * 6.0 / t) / t, bot = (N[r] - M[r] * 6.0 / t) / t;
        for (int pnt = 0; pnt < 3; ++pnt) {
          S[frag_index(8 * r + pnt, col, KT)] = top;
          S[frag_index(8 * r + 3 + pnt, col, KT)] = bot;
        }
      }
    }
  failed[i] = 0;
}

// ------------------------------------------------------------------------------------------
// K2 apply: von Mises + envelope for shell families (3 m-tiles: xx, yy, xy at 8 points)
// ------------------------------------------------------------------------------------------
// The FP64 tensor pipe is the shared resource here (DMMA and scalar FP64 issue to the same pipe and
// a DMMA holds the dispatch port for its 16 cycles), so every non-DMMA instruction in the step loop
// costs wall time: the loop body is written for minimum instruction count -- row pointers bumped
// once per four tiles with immediate offsets in between, register double-buffering without moves,
// a Newton square root on MUFU.RSQ64H instead of the IEEE slow path, compare/select envelopes.

// `live` = this lane's result point exists (triangles use 6 of the 8 rows of an m-tile).  Every lane
// of the warp must reach every mma.sync, so dead lanes run the same loop and only their stores and
// envelope updates are predicated off (ALL_LIVE = true for quads compiles the predicate away).
template <int KT, bool WRITE_VM, bool GUARD, bool ALL_LIVE, class OUT_T, bool ENV>
__device__ __forceinline__ void shell_tile(const double (&a)[3][KT], const double (&b)[KT], OUT_T*& vmp0,
                                           OUT_T*& vmp1, size_t ld8, int t0, int nsteps, double& emax,
                                           double& emin, bool live)
{
  double c[3][2] = {{0, 0}, {0, 0}, {0, 0}};
#pragma unroll
  for (int j = 0; j < KT; ++j)
#pragma unroll
    for (int m = 0; m < 3; ++m) dmma884(c[m][0], c[m][1], a[m][j], b[j]);
  // von Mises, FFaTensorTransforms.C:33-36: sqrt(s11^2 + s22^2 - s11*s22 + 3*s12^2).  Envelope only (WRITE_VM = false):
  // the running max / min are kept on the radicand, which orders like its root; the root is taken once at the end.
  double q0 = fma(c[2][0] * 3.0, c[2][0], fma(-c[0][0], c[1][0], fma(c[1][0], c[1][0], c[0][0] * c[0][0])));
  double q1 = fma(c[2][1] * 3.0, c[2][1], fma(-c[0][1], c[1][1], fma(c[1][1], c[1][1], c[0][1] * c[0][1])));
  const double v0 = WRITE_VM ? sqrt_pos(q0) : q0, v1 = WRITE_VM ? sqrt_pos(q1) : q1;
  const bool wr = ALL_LIVE || live;
  // compare / select (measured: 64-bit integer max / min on the ALU pipe costs more issue slots than three DSETP here)
  if (GUARD) {
    if (t0 < nsteps) {
      if (WRITE_VM && wr) *vmp0 = (OUT_T)v0;
      if (ENV) { emax = v0 > emax ? v0 : emax; emin = v0 < emin ? v0 : emin; }
    }
    if (t0 + 1 < nsteps) {
      if (WRITE_VM && wr) *vmp1 = (OUT_T)v1;
      if (ENV) { emax = v1 > emax ? v1 : emax; emin = v1 < emin ? v1 : emin; }
    }
  } else {
    if (WRITE_VM && wr) { *vmp0 = (OUT_T)v0; *vmp1 = (OUT_T)v1; }
    if (ENV) {
      const bool p = v0 > v1;
      const double hi = p ? v0 : v1, lo = p ? v1 : v0;
      emax = hi > emax ? hi : emax;
      emin = lo < emin ? lo : emin;
    }
  }
  if (WRITE_VM) { vmp0 += ld8; vmp1 += ld8; }
}

// OUT_T / ENV / roff: the same kernel writes the von Mises values of a `-vmStress` results database straight into the
// step records (io_rdb.cu): OUT_T = float or double, column = record slot roff[i] + point instead of the result-point
// number, no envelope.
template <int KT, bool WRITE_VM, bool ALL_LIVE, class OUT_T = double, bool ENV = true>
__global__ void __launch_bounds__(256, 2)
k2_shell_vm_kernel(const double* __restrict__ U, size_t ldu, int nsteps, int nsteps_pad,
                   const double* __restrict__ Sfrag, const int* __restrict__ edof,
                   const int* __restrict__ ptoff, const unsigned char* __restrict__ failed,
                   int nelt, int nstrp, OUT_T* __restrict__ vm, size_t ld_vm,
                   double* __restrict__ env_max, double* __restrict__ env_min, const long long* __restrict__ roff = nullptr,
                   const int* __restrict__ list = nullptr /* family elements to process, NULL = all */)
{
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, t4 = lane & 3;
  const int il = blockIdx.x * (blockDim.x >> 5) + warp;
  if (il >= nelt) return;  // whole warp
  const int i = list ? __ldg(list + il) : il;
  const bool live = ALL_LIVE || g < nstrp;
  size_t pt;
  if (ENV)
    pt = (size_t)ptoff[i] + (live ? g : 0);
  else {
    const long long r0 = roff[i];
    if (r0 < 0) return;   // element not in the results database; whole warp
    pt = (size_t)r0 + (live ? g : 0);
  }

  if (failed[i]) {  // operator build failed: hugeVal results (stressRoutines.f90:264-268); whole warp
    if (live) {
      if (WRITE_VM)
        for (int t = t4; t < nsteps; t += 4) vm[(size_t)t * ld_vm + pt] = (OUT_T)kHuge;
      if (ENV && t4 == 0 && nsteps > 0) { env_max[pt] = kHuge; if (kHuge < env_min[pt]) env_min[pt] = kHuge; }
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
  // lane owns point g at steps t0 = 8*tile + 2*t4 and t0 + 1
  OUT_T* vmp0 = WRITE_VM ? vm + (size_t)(2 * t4) * ld_vm + pt : nullptr;
  OUT_T* vmp1 = WRITE_VM ? vmp0 + ld_vm : nullptr;
  const size_t ld8 = ld_vm * 8;

  const int ntiles = nsteps_pad >> 3;            // multiple of 8
  const int nfull = ((nsteps >> 3) >> 2) << 2;   // tiles (in groups of 4) with all 8 steps valid
  // two register buffers, loads of tile n+1 in flight while tile n is multiplied (a third buffer was
  // measured slower on B200: 27.9 vs 26.0 ms per 5e8 element.steps, the extra registers spill)
  double b0[KT], b1[KT];
#pragma unroll
  for (int j = 0; j < KT; ++j) b0[j] = up[j][0];

  int nt = 0;
  for (; nt < nfull; nt += 4) {
#pragma unroll
    for (int j = 0; j < KT; ++j) b1[j] = up[j][8];
    shell_tile<KT, WRITE_VM, false, ALL_LIVE, OUT_T, ENV>(a, b0, vmp0, vmp1, ld8, 0, 0, emax, emin, live);
#pragma unroll
    for (int j = 0; j < KT; ++j) b0[j] = up[j][16];
    shell_tile<KT, WRITE_VM, false, ALL_LIVE, OUT_T, ENV>(a, b1, vmp0, vmp1, ld8, 0, 0, emax, emin, live);
#pragma unroll
    for (int j = 0; j < KT; ++j) b1[j] = up[j][24];
    shell_tile<KT, WRITE_VM, false, ALL_LIVE, OUT_T, ENV>(a, b0, vmp0, vmp1, ld8, 0, 0, emax, emin, live);
#pragma unroll
    for (int j = 0; j < KT; ++j) { up[j] += 32; b0[j] = up[j][0]; }  // U rows carry 64 doubles of slack
    shell_tile<KT, WRITE_VM, false, ALL_LIVE, OUT_T, ENV>(a, b1, vmp0, vmp1, ld8, 0, 0, emax, emin, live);
  }
  for (; nt < ntiles && nt * 8 < nsteps; ++nt) {  // ragged tail: guarded stores
#pragma unroll
    for (int j = 0; j < KT; ++j) { up[j] += 8; b1[j] = up[j][0]; }
    shell_tile<KT, WRITE_VM, true, ALL_LIVE, OUT_T, ENV>(a, b0, vmp0, vmp1, ld8, nt * 8 + 2 * t4, nsteps, emax, emin, live);
#pragma unroll
    for (int j = 0; j < KT; ++j) b0[j] = b1[j];
  }
  // combine the four lanes that share a result point, then fold into the stored envelope
  emax = fmax(emax, __shfl_xor_sync(0xffffffffu, emax, 1));
  emin = fmin(emin, __shfl_xor_sync(0xffffffffu, emin, 1));
  emax = fmax(emax, __shfl_xor_sync(0xffffffffu, emax, 2));
  emin = fmin(emin, __shfl_xor_sync(0xffffffffu, emin, 2));
  if (!WRITE_VM) {   // radicand -> von Mises (kHuge: no step seen)
    emax = sqrt_pos(emax);
    emin = emin == kHuge ? kHuge : sqrt_pos(emin);
  }
  if (ENV && live && t4 == 0 && nsteps > 0) {
    if (emax > env_max[pt]) env_max[pt] = emax;
    if (emin < env_min[pt]) env_min[pt] = emin;
  }
}

// ------------------------------------------------------------------------------------------
// Flat quadrilaterals: membrane / bending split of the operator
// ------------------------------------------------------------------------------------------
// For a flat element the rigid-body projector changes nothing (the strain-displacement matrices annihilate the rigid modes by
// themselves), the membrane strains see only the nodal TRANSLATIONS and the curvatures only the nodal ROTATIONS.  With
// top / bottom = membrane +- bending the 24 x 24 operator splits into two 12-row x 12-column blocks,
//     M = (S_top + S_bot) / 2  on the translation columns,   B = (S_top - S_bot) / 2  on the rotation columns,
// and the off-diagonal blocks are rounding noise (checked per element against 1e-12 of the block they would add to; an
// element that is not flat to that level, or warped, keeps the dense operator).  12 DMMA per 8 steps instead of 18.
// Fragment layout per element: [M | B][m-tile][k-tile][lane]; m-tile 0 rows = xx at the 4 nodes, yy at the 4 nodes;
// m-tile 1 rows = xy at the 4 nodes, twice; column k of a block = 3 * node + component.
__global__ void build_quad_flat_kernel(int nelt, const double* __restrict__ Sfrag, const unsigned char* __restrict__ failed,
                                       double* __restrict__ Ffrag, unsigned char* __restrict__ flat)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nelt) return;
  const double* S = Sfrag + (size_t)i * 3 * 6 * 32;
  double* F = Ffrag + (size_t)i * 12 * 32;
  double keep[2] = {0.0, 0.0}, drop[2] = {0.0, 0.0};   // [0] translation columns, [1] rotation columns
  for (int c = 0; c < 3; ++c)
    for (int n = 0; n < 4; ++n)
      for (int col = 0; col < 24; ++col) {
        const double top = S[frag_index(c * 8 + n, col, 6)], bot = S[frag_index(c * 8 + 4 + n, col, 6)];
        const double m = 0.5 * (top + bot), b = 0.5 * (top - bot);
        const int rot = (col % 6) >= 3, k = 3 * (col / 6) + (col % 3);
        if (!rot) { keep[0] = fmax(keep[0], fabs(m)); drop[0] = fmax(drop[0], fabs(b)); }
        else      { keep[1] = fmax(keep[1], fabs(b)); drop[1] = fmax(drop[1], fabs(m)); }
        const double v = rot ? b : m;
        double* blk = F + (size_t)(rot ? 6 : 0) * 32;
        if (c < 2) blk[frag_index(4 * c + n, k, 3)] = v;
        else { blk[frag_index(8 + n, k, 3)] = v; blk[frag_index(12 + n, k, 3)] = v; }
      }
  flat[i] = !failed[i] && drop[0] <= 1.0e-12 * keep[0] && drop[1] <= 1.0e-12 * keep[1];
}

template <bool WRITE_VM, bool GUARD, int NK = 3>
__device__ __forceinline__ void quad_flat_tile(const double (&am)[2][NK], const double (&ab)[2][NK], const double (&bm)[NK],
                                               const double (&bb)[NK], double sgn, double*& vmp0, double*& vmp1, size_t ld8, int t0,
                                               int nsteps, double& emax, double& emin)
{
  double cm[2][2] = {{0, 0}, {0, 0}}, cb[2][2] = {{0, 0}, {0, 0}};
#pragma unroll
  for (int j = 0; j < NK; ++j)
#pragma unroll
    for (int m = 0; m < 2; ++m) { dmma884(cm[m][0], cm[m][1], am[m][j], bm[j]); dmma884(cb[m][0], cb[m][1], ab[m][j], bb[j]); }
  double v[2];
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    // lanes g < 4 hold xx of node g (top point g), lanes g >= 4 yy of node g - 4 (bottom point g): the partner lane ^ 16
    // holds the other direct stress of the same node; von Mises is symmetric in the two
    const double own = fma(sgn, cb[0][q], cm[0][q]);
    const double oth = fma(sgn, __shfl_xor_sync(0xffffffffu, cb[0][q], 16), __shfl_xor_sync(0xffffffffu, cm[0][q], 16));
    const double sxy = fma(sgn, cb[1][q], cm[1][q]);
    const double rad = fma(sxy * 3.0, sxy, fma(-own, oth, fma(oth, oth, own * own)));
    v[q] = WRITE_VM ? sqrt_pos(rad) : rad;
  }
  if (GUARD) {
    if (t0 < nsteps) { if (WRITE_VM) *vmp0 = v[0]; emax = v[0] > emax ? v[0] : emax; emin = v[0] < emin ? v[0] : emin; }
    if (t0 + 1 < nsteps) { if (WRITE_VM) *vmp1 = v[1]; emax = v[1] > emax ? v[1] : emax; emin = v[1] < emin ? v[1] : emin; }
  } else {
    if (WRITE_VM) { *vmp0 = v[0]; *vmp1 = v[1]; }
    const bool p = v[0] > v[1];
    const double hi = p ? v[0] : v[1], lo = p ? v[1] : v[0];
    emax = hi > emax ? hi : emax;
    emin = lo < emin ? lo : emin;
  }
  if (WRITE_VM) { vmp0 += ld8; vmp1 += ld8; }
}

template <bool WRITE_VM>
__global__ void __launch_bounds__(256, 2)
k2_quad_flat_vm_kernel(const double* __restrict__ U, size_t ldu, int nsteps, int nsteps_pad, const double* __restrict__ Ffrag,
                       const int* __restrict__ edof, const int* __restrict__ ptoff, int nlist, const int* __restrict__ list,
                       double* __restrict__ vm, size_t ld_vm, double* __restrict__ env_max, double* __restrict__ env_min)
{
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, t4 = lane & 3;
  const int il = blockIdx.x * (blockDim.x >> 5) + warp;
  if (il >= nlist) return;  // whole warp
  const int i = __ldg(list + il);
  const size_t pt = (size_t)ptoff[i] + g;
  const double sgn = g < 4 ? 1.0 : -1.0;   // top = membrane + bending, bottom = membrane - bending

  double am[2][3], ab[2][3];
  const double* ff = Ffrag + (size_t)i * 12 * 32 + lane;
#pragma unroll
  for (int m = 0; m < 2; ++m)
#pragma unroll
    for (int j = 0; j < 3; ++j) { am[m][j] = __ldg(ff + (size_t)(m * 3 + j) * 32); ab[m][j] = __ldg(ff + (size_t)(6 + m * 3 + j) * 32); }
  // B operands: column k = 4 j + t4 of a block = component k % 3 of node k / 3
  const double *upm[3], *upb[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const int k = 4 * j + t4, d = 6 * (k / 3) + k % 3;
    upm[j] = U + (size_t)__ldg(edof + (size_t)i * 24 + d) * ldu + g;
    upb[j] = U + (size_t)__ldg(edof + (size_t)i * 24 + d + 3) * ldu + g;
  }
  double emax = 0.0, emin = kHuge;
  double* vmp0 = WRITE_VM ? vm + (size_t)(2 * t4) * ld_vm + pt : nullptr;
  double* vmp1 = WRITE_VM ? vmp0 + ld_vm : nullptr;
  const size_t ld8 = ld_vm * 8;
  const int ntiles = nsteps_pad >> 3;
  const int nfull = ((nsteps >> 3) >> 1) << 1;   // tiles (in pairs) with all 8 steps valid
  double m0[3], r0[3], m1[3], r1[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) { m0[j] = upm[j][0]; r0[j] = upb[j][0]; }
  int nt = 0;
  for (; nt < nfull; nt += 2) {
#pragma unroll
    for (int j = 0; j < 3; ++j) { m1[j] = upm[j][8]; r1[j] = upb[j][8]; }
    quad_flat_tile<WRITE_VM, false>(am, ab, m0, r0, sgn, vmp0, vmp1, ld8, 0, 0, emax, emin);
#pragma unroll
    for (int j = 0; j < 3; ++j) { upm[j] += 16; upb[j] += 16; m0[j] = upm[j][0]; r0[j] = upb[j][0]; }   // rows carry 64 doubles of slack
    quad_flat_tile<WRITE_VM, false>(am, ab, m1, r1, sgn, vmp0, vmp1, ld8, 0, 0, emax, emin);
  }
  for (; nt < ntiles && nt * 8 < nsteps; ++nt) {  // ragged tail: guarded stores
#pragma unroll
    for (int j = 0; j < 3; ++j) { upm[j] += 8; upb[j] += 8; m1[j] = upm[j][0]; r1[j] = upb[j][0]; }
    quad_flat_tile<WRITE_VM, true>(am, ab, m0, r0, sgn, vmp0, vmp1, ld8, nt * 8 + 2 * t4, nsteps, emax, emin);
#pragma unroll
    for (int j = 0; j < 3; ++j) { m0[j] = m1[j]; r0[j] = r1[j]; }
  }
  emax = fmax(emax, __shfl_xor_sync(0xffffffffu, emax, 1));
  emin = fmin(emin, __shfl_xor_sync(0xffffffffu, emin, 1));
  emax = fmax(emax, __shfl_xor_sync(0xffffffffu, emax, 2));
  emin = fmin(emin, __shfl_xor_sync(0xffffffffu, emin, 2));
  if (!WRITE_VM) {   // radicand -> von Mises (kHuge: no step seen)
    emax = sqrt_pos(emax);
    emin = emin == kHuge ? kHuge : sqrt_pos(emin);
  }
  if (t4 == 0 && nsteps > 0) {
    if (emax > env_max[pt]) env_max[pt] = emax;
    if (emin < env_min[pt]) env_min[pt] = emin;
  }
}

// ------------------------------------------------------------------------------------------
// Flat quadrilaterals in flat regions: the in-plane form
// ------------------------------------------------------------------------------------------
// A flat element never sees the translation along its normal nor the rotation about it.  Where ALL flat quadrilaterals around
// a node lie in one plane (decks, webs, panels: most nodes of a plated structure), the node gets a frame (a1, a2 in the plane)
// and the recovery operator is rotated into it once: four rows (u, v, theta1, theta2) of Rp = W . R replace the six global
// ones, so K1 expands 2/3 of the rows for such nodes and the element operator shrinks to two 12 x 8 blocks
//     M' = M . [a1 a2] per node (translations),   B' = B . [a1 a2] per node (rotations):
// 8 DMMA per 8 steps instead of 12 and four U rows per node instead of six.  The dropped columns (M . a3, B . a3) are checked
// per element against 1e-12 of the kept ones; an element that fails, or has a node on a fold line, keeps the flat form above
// on the global rows.  Fragment layout per element: [M' | B'][m-tile][k-tile][lane], column k of a block = 2 * node + (u|v).
__global__ void build_quad_planar_kernel(int ncand, const int* __restrict__ celem, const int* __restrict__ cnode,
                                         const double* __restrict__ frames, const double* __restrict__ Ffrag,
                                         double* __restrict__ Pfrag, unsigned char* __restrict__ ok)
{
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ncand) return;
  const double* F = Ffrag + (size_t)celem[c] * 12 * 32;
  double* P = Pfrag + (size_t)c * 8 * 32;
  double keep[2] = {0.0, 0.0}, drop[2] = {0.0, 0.0};
  for (int n = 0; n < 4; ++n) {
    const double* fr = frames + (size_t)cnode[(size_t)c * 4 + n] * 6;
    const double a1[3] = {fr[0], fr[1], fr[2]}, a2[3] = {fr[3], fr[4], fr[5]};
    const double a3[3] = {a1[1] * a2[2] - a1[2] * a2[1], a1[2] * a2[0] - a1[0] * a2[2], a1[0] * a2[1] - a1[1] * a2[0]};
    for (int blk = 0; blk < 2; ++blk)
      for (int r = 0; r < 16; ++r) {
        double v[3];
        for (int k = 0; k < 3; ++k) v[k] = F[(size_t)blk * 6 * 32 + frag_index(r, 3 * n + k, 3)];
        const double pu = v[0] * a1[0] + v[1] * a1[1] + v[2] * a1[2];
        const double pv = v[0] * a2[0] + v[1] * a2[1] + v[2] * a2[2];
        const double pw = v[0] * a3[0] + v[1] * a3[1] + v[2] * a3[2];
        P[(size_t)blk * 4 * 32 + frag_index(r, 2 * n, 2)] = pu;
        P[(size_t)blk * 4 * 32 + frag_index(r, 2 * n + 1, 2)] = pv;
        keep[blk] = fmax(keep[blk], fmax(fabs(pu), fabs(pv)));
        drop[blk] = fmax(drop[blk], fabs(pw));
      }
  }
  ok[c] = drop[0] <= 1.0e-12 * keep[0] && drop[1] <= 1.0e-12 * keep[1];
}

// edof2: [candidate][20] = rows of Up (u, v, theta1, theta2) of the four nodes, [16] = first result point of the element
// (three CTAs per SM when the history is written: 80 registers without spills, measured 14.7 vs 16.0 ms; the envelope-only
// variant would spill there and keeps two)
constexpr int kPlStages = 3;            // staged variant: 16-step pairs in flight per warp
constexpr int kPlRow = 20;              // doubles per staged row: 16 steps + 4 of padding (conflict-free 8-byte fragment reads)
__device__ __forceinline__ void pl_cp_async16(void* smem_dst, const void* gsrc)
{
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}

template <bool WRITE_VM, bool STAGED = false>
__global__ void __launch_bounds__(256, (WRITE_VM || STAGED) ? 3 : 2)
k2_quad_planar_vm_kernel(const double* __restrict__ Up, size_t ldu, int nsteps, int nsteps_pad, const double* __restrict__ Pfrag,
                         const int* __restrict__ edof2, int nlist, const int* __restrict__ list, double* __restrict__ vm,
                         size_t ld_vm, double* __restrict__ env_max, double* __restrict__ env_min)
{
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, t4 = lane & 3;
  const int il = blockIdx.x * (blockDim.x >> 5) + warp;
  if (il >= nlist) return;  // whole warp
  const int c = __ldg(list + il);
  const int* ed = edof2 + (size_t)c * 20;
  const size_t pt = (size_t)__ldg(ed + 16) + g;
  const double sgn = g < 4 ? 1.0 : -1.0;   // top = membrane + bending, bottom = membrane - bending

  double am[2][2], ab[2][2];
  const double* ff = Pfrag + (size_t)c * 8 * 32 + lane;
#pragma unroll
  for (int m = 0; m < 2; ++m)
#pragma unroll
    for (int j = 0; j < 2; ++j) { am[m][j] = __ldg(ff + (size_t)(m * 2 + j) * 32); ab[m][j] = __ldg(ff + (size_t)(4 + m * 2 + j) * 32); }
  // B operands: column k = 4 j + t4 of a block = in-plane component k % 2 of node k / 2
  const double *upm[2], *upb[2];
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const int k = 4 * j + t4;
    upm[j] = Up + (size_t)__ldg(ed + 4 * (k >> 1) + (k & 1)) * ldu;
    upb[j] = Up + (size_t)__ldg(ed + 4 * (k >> 1) + 2 + (k & 1)) * ldu;
  }
  double emax = 0.0, emin = kHuge;
  const int ntiles = nsteps_pad >> 3;
  const int nfull = ((nsteps >> 3) >> 1) << 1;   // tiles (in pairs) with all 8 steps valid
  if (STAGED) {
    // The sixteen U rows of the element, 16 steps at a time, go through shared memory with cp.async (kPlStages pairs in
    // flight per warp, no registers tied up by the prefetch).  Staged row (blk * 2 + j) * 4 + t4 = the row lane t4 reads as
    // B operand of k-tile j of block blk (0 membrane, 1 bending).
    extern __shared__ __align__(16) double pl_smem[];
    double* st = pl_smem + (size_t)warp * kPlStages * 16 * kPlRow;
    // this lane copies four 16-byte pieces per pair: piece q = lane + 32 i -> staged row q / 8, steps 2 (q % 8)
    const double* src[4];
    int dst[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int q = lane + 32 * i, sr = q >> 3, c2 = (q & 7) * 2;
      const int blk = sr >> 3, k = 4 * ((sr >> 2) & 1) + (sr & 3);
      src[i] = Up + (size_t)__ldg(ed + 4 * (k >> 1) + 2 * blk + (k & 1)) * ldu + c2;
      dst[i] = sr * kPlRow + c2;
    }
    const int npairs = nfull >> 1;
    auto issue = [&](int pr) {
      if (pr < npairs) {
        double* d = st + (size_t)(pr % kPlStages) * 16 * kPlRow;
#pragma unroll
        for (int i = 0; i < 4; ++i) pl_cp_async16(d + dst[i], src[i] + (size_t)pr * 16);
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
#pragma unroll
    for (int pr = 0; pr < kPlStages - 1; ++pr) issue(pr);
    double* vmp0 = WRITE_VM ? vm + (size_t)(2 * t4) * ld_vm + pt : nullptr;
    double* vmp1 = WRITE_VM ? vmp0 + ld_vm : nullptr;
    const size_t ld8 = ld_vm * 8;
    for (int pr = 0; pr < npairs; ++pr) {
      issue(pr + kPlStages - 1);
      asm volatile("cp.async.wait_group %0;" ::"n"(kPlStages - 1) : "memory");
      __syncwarp();
      const double* d = st + (size_t)(pr % kPlStages) * 16 * kPlRow + t4 * kPlRow + g;
      double ma[2], ra[2], mb[2], rb[2];
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        ma[j] = d[(size_t)j * 4 * kPlRow];     mb[j] = d[(size_t)j * 4 * kPlRow + 8];
        ra[j] = d[(size_t)(2 + j) * 4 * kPlRow]; rb[j] = d[(size_t)(2 + j) * 4 * kPlRow + 8];
      }
      quad_flat_tile<WRITE_VM, false, 2>(am, ab, ma, ra, sgn, vmp0, vmp1, ld8, 0, 0, emax, emin);
      quad_flat_tile<WRITE_VM, false, 2>(am, ab, mb, rb, sgn, vmp0, vmp1, ld8, 0, 0, emax, emin);
      __syncwarp();   // the stage is refilled by the next iteration's issue
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
  } else {
    double* vmp0 = WRITE_VM ? vm + (size_t)(2 * t4) * ld_vm + pt : nullptr;
    double* vmp1 = WRITE_VM ? vmp0 + ld_vm : nullptr;
    const size_t ld8 = ld_vm * 8;
    double m0[2], r0[2], m1[2], r1[2];
    const double *qm[2] = {upm[0] + g, upm[1] + g}, *qb[2] = {upb[0] + g, upb[1] + g};
#pragma unroll
    for (int j = 0; j < 2; ++j) { m0[j] = qm[j][0]; r0[j] = qb[j][0]; }
    for (int nt = 0; nt < nfull; nt += 2) {
#pragma unroll
      for (int j = 0; j < 2; ++j) { m1[j] = qm[j][8]; r1[j] = qb[j][8]; }
      quad_flat_tile<WRITE_VM, false, 2>(am, ab, m0, r0, sgn, vmp0, vmp1, ld8, 0, 0, emax, emin);
#pragma unroll
      for (int j = 0; j < 2; ++j) { qm[j] += 16; qb[j] += 16; m0[j] = qm[j][0]; r0[j] = qb[j][0]; }
      quad_flat_tile<WRITE_VM, false, 2>(am, ab, m1, r1, sgn, vmp0, vmp1, ld8, 0, 0, emax, emin);
    }
  }
  {   // ragged tail: one tile at a time, guarded stores
    double* vmp0 = WRITE_VM ? vm + (size_t)(nfull * 8 + 2 * t4) * ld_vm + pt : nullptr;
    double* vmp1 = WRITE_VM ? vmp0 + ld_vm : nullptr;
    for (int nt = nfull; nt < ntiles && nt * 8 < nsteps; ++nt) {
      double m0[2], r0[2];
#pragma unroll
      for (int j = 0; j < 2; ++j) { m0[j] = upm[j][(size_t)nt * 8 + g]; r0[j] = upb[j][(size_t)nt * 8 + g]; }
      quad_flat_tile<WRITE_VM, true, 2>(am, ab, m0, r0, sgn, vmp0, vmp1, ld_vm * 8, nt * 8 + 2 * t4, nsteps, emax, emin);
    }
  }
  emax = fmax(emax, __shfl_xor_sync(0xffffffffu, emax, 1));
  emin = fmin(emin, __shfl_xor_sync(0xffffffffu, emin, 1));
  emax = fmax(emax, __shfl_xor_sync(0xffffffffu, emax, 2));
  emin = fmin(emin, __shfl_xor_sync(0xffffffffu, emin, 2));
  if (!WRITE_VM) {   // radicand -> von Mises (kHuge: no step seen)
    emax = sqrt_pos(emax);
    emin = emin == kHuge ? kHuge : sqrt_pos(emin);
  }
  if (t4 == 0 && nsteps > 0) {
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
  for (int e : elements_of_type(p, sam, elm, type)) {
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

// Which flat quadrilaterals take the in-plane form, the node frames, the rows of Rp / Up and the row tiles of R the von Mises
// path still has to expand for everything else.  lst[0] = flat elements on entry; the ones that qualify move to lst[2]
// (positions in the candidate arrays fast2 / edof2).
static int setup_planar_quads(fsr_part* p, const fsr_sam* sam, const fsr_elmdata* elm, FamilyData& f, const std::vector<int>& conn,
                              const std::vector<int>& ptoff, std::vector<int> (&lst)[3], cudaStream_t s)
{
  const int nnod = sam->nnod;
  const double* X = elm->xyz;
  // the plane of every node: the normal of its first flat quadrilateral; -1 = its flat quadrilaterals are not coplanar
  std::vector<signed char> state((size_t)nnod, 0);
  std::vector<double> nrm((size_t)nnod * 3, 0.0);
  for (int i : lst[0]) {
    const int* nd = &conn[(size_t)i * 4];
    double d1[3], d2[3], n[3];
    for (int k = 0; k < 3; ++k) {
      d1[k] = X[3 * (size_t)nd[2] + k] - X[3 * (size_t)nd[0] + k];
      d2[k] = X[3 * (size_t)nd[3] + k] - X[3 * (size_t)nd[1] + k];
    }
    n[0] = d1[1] * d2[2] - d1[2] * d2[1]; n[1] = d1[2] * d2[0] - d1[0] * d2[2]; n[2] = d1[0] * d2[1] - d1[1] * d2[0];
    const double len = std::sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
    if (!(len > 0.0)) { for (int k = 0; k < 4; ++k) state[(size_t)nd[k]] = -1; continue; }
    int kmax = 0;   // one sign for both orientations of the plane (element blocks of a sharded part see other neighbours)
    for (int k = 1; k < 3; ++k) if (std::fabs(n[k]) > std::fabs(n[kmax])) kmax = k;
    const double sg = n[kmax] < 0.0 ? -1.0 / len : 1.0 / len;
    for (int k = 0; k < 3; ++k) n[k] *= sg;
    for (int k = 0; k < 4; ++k) {
      const size_t j = (size_t)nd[k];
      if (state[j] == 0) { state[j] = 1; for (int d = 0; d < 3; ++d) nrm[3 * j + d] = n[d]; }
      else if (state[j] == 1) {
        const double c = n[0] * nrm[3 * j] + n[1] * nrm[3 * j + 1] + n[2] * nrm[3 * j + 2];
        if (1.0 - std::fabs(c) > 1.0e-10) state[j] = -1;
      }
    }
  }
  std::vector<int> celem, cnode;   // candidates: family index, nodes
  for (int i : lst[0]) {
    const int* nd = &conn[(size_t)i * 4];
    if (state[(size_t)nd[0]] == 1 && state[(size_t)nd[1]] == 1 && state[(size_t)nd[2]] == 1 && state[(size_t)nd[3]] == 1) {
      celem.push_back(i);
      cnode.insert(cnode.end(), nd, nd + 4);
    }
  }
  const int ncand = (int)celem.size();
  if (ncand == 0) return FSR_OK;
  // node frames: a1 = the global axis least aligned with the normal, projected into the plane; a2 = a3 x a1
  std::vector<double> frames((size_t)nnod * 6, 0.0);
  for (int j = 0; j < nnod; ++j) {
    if (state[(size_t)j] != 1) continue;
    const double* a3 = &nrm[3 * (size_t)j];
    int k0 = 0;
    for (int k = 1; k < 3; ++k) if (std::fabs(a3[k]) < std::fabs(a3[k0])) k0 = k;
    double a1[3] = {-a3[k0] * a3[0], -a3[k0] * a3[1], -a3[k0] * a3[2]};
    a1[k0] += 1.0;
    const double len = std::sqrt(a1[0] * a1[0] + a1[1] * a1[1] + a1[2] * a1[2]);
    for (int k = 0; k < 3; ++k) a1[k] /= len;
    double* fr = &frames[6 * (size_t)j];
    fr[0] = a1[0]; fr[1] = a1[1]; fr[2] = a1[2];
    fr[3] = a3[1] * a1[2] - a3[2] * a1[1]; fr[4] = a3[2] * a1[0] - a3[0] * a1[2]; fr[5] = a3[0] * a1[1] - a3[1] * a1[0];
  }
  int *d_celem = nullptr, *d_cnode = nullptr;
  double* d_frames = nullptr;
  unsigned char* d_ok = nullptr;
  FSR_CUDA(cudaMalloc(&d_celem, sizeof(int) * celem.size()));
  FSR_CUDA(cudaMalloc(&d_cnode, sizeof(int) * cnode.size()));
  FSR_CUDA(cudaMalloc(&d_frames, sizeof(double) * frames.size()));
  FSR_CUDA(cudaMalloc(&d_ok, (size_t)ncand));
  FSR_CUDA(cudaMalloc(&f.fast2, sizeof(double) * (size_t)ncand * 8 * 32));
  FSR_CUDA(cudaMemcpyAsync(d_celem, celem.data(), sizeof(int) * celem.size(), cudaMemcpyHostToDevice, s));
  FSR_CUDA(cudaMemcpyAsync(d_cnode, cnode.data(), sizeof(int) * cnode.size(), cudaMemcpyHostToDevice, s));
  FSR_CUDA(cudaMemcpyAsync(d_frames, frames.data(), sizeof(double) * frames.size(), cudaMemcpyHostToDevice, s));
  FSR_CUDA(cudaMemsetAsync(f.fast2, 0, sizeof(double) * (size_t)ncand * 8 * 32, s));
  build_quad_planar_kernel<<<(ncand + 63) / 64, 64, 0, s>>>(ncand, d_celem, d_cnode, d_frames, f.fast, f.fast2, d_ok);
  FSR_LAUNCH_CHECK();
  std::vector<unsigned char> ok((size_t)ncand);
  FSR_CUDA(cudaMemcpyAsync(ok.data(), d_ok, (size_t)ncand, cudaMemcpyDeviceToHost, s));
  FSR_CUDA(cudaStreamSynchronize(s));
  cudaFree(d_celem); cudaFree(d_cnode); cudaFree(d_frames); cudaFree(d_ok);
  // rows of Rp / Up: four per node that some in-plane element reads, in node order
  std::vector<int> prow((size_t)nnod, -1);
  std::vector<unsigned char> is_planar((size_t)f.nelt, 0);
  int nok = 0;
  for (int c = 0; c < ncand; ++c) {
    if (!ok[(size_t)c]) continue;
    ++nok;
    is_planar[(size_t)celem[(size_t)c]] = 1;
    for (int k = 0; k < 4; ++k) prow[(size_t)cnode[(size_t)c * 4 + k]] = 0;
  }
  if (nok == 0) { cudaFree(f.fast2); f.fast2 = nullptr; return FSR_OK; }
  int nrows = 0;
  for (int j = 0; j < nnod; ++j)
    if (prow[(size_t)j] == 0) { prow[(size_t)j] = nrows; nrows += 4; }
  std::vector<int> src((size_t)nrows * 3);
  std::vector<double> w((size_t)nrows * 3);
  for (int j = 0; j < nnod; ++j) {
    if (prow[(size_t)j] < 0) continue;
    const int js = sam->madof[j] - 1;
    for (int q = 0; q < 4; ++q)
      for (int k = 0; k < 3; ++k) {
        src[((size_t)prow[(size_t)j] + q) * 3 + k] = js + (q >= 2 ? 3 : 0) + k;
        w[((size_t)prow[(size_t)j] + q) * 3 + k] = frames[6 * (size_t)j + (q & 1) * 3 + k];
      }
  }
  std::vector<int> edof2((size_t)ncand * 20, 0), planar_list, flat_rest;
  for (int c = 0; c < ncand; ++c) {
    if (!ok[(size_t)c]) continue;
    planar_list.push_back(c);
    for (int k = 0; k < 4; ++k)
      for (int q = 0; q < 4; ++q) edof2[(size_t)c * 20 + 4 * k + q] = prow[(size_t)cnode[(size_t)c * 4 + k]] + q;
    edof2[(size_t)c * 20 + 16] = ptoff[(size_t)celem[(size_t)c]];
  }
  for (int i : lst[0])
    if (!is_planar[(size_t)i]) flat_rest.push_back(i);
  // row tiles of R that anything but the in-plane quadrilaterals reads (every other active element of the part)
  const int ntile = p->nrows_pad / 128;
  std::vector<unsigned char> need((size_t)ntile, 0);
  std::vector<unsigned char> planar_sam((size_t)sam->nel, 0);
  {
    std::vector<int> elem_host((size_t)f.nelt);
    FSR_CUDA(cudaMemcpy(elem_host.data(), f.elem, sizeof(int) * (size_t)f.nelt, cudaMemcpyDeviceToHost));
    for (int i = 0; i < f.nelt; ++i) planar_sam[(size_t)elem_host[(size_t)i]] = is_planar[(size_t)i];
  }
  for (int e = 0; e < sam->nel; ++e) {
    if (planar_sam[(size_t)e] || (elm->elmid && elm->elmid[e] < 1)) continue;
    for (int ip = sam->mpmnpc[e] - 1; ip < sam->mpmnpc[e + 1] - 1; ++ip) {
      const int n = sam->mmnpc[ip] - 1;
      if (n < 0 || n >= nnod) continue;
      for (int d = sam->madof[n] - 1; d < sam->madof[n + 1] - 1; ++d) need[(size_t)(d / 128)] = 1;
    }
  }
  std::vector<int> tiles;
  for (int t = 0; t < ntile; ++t)
    if (need[(size_t)t]) tiles.push_back(t);
  // Is it worth it?  The in-plane form saves ~14 ps of K2 per quadrilateral and step (24 instead of 38 on the global rows,
  // profiles/R4) and costs the expansion of its own rows, ~0.068 ps per row, step and reduced DOF, minus the row tiles of R
  // nobody reads any more.  A plate of quadrilaterals frees every tile (C2: 4 M rows instead of 6 M); quadrilaterals
  // scattered among triangles free none and K1 would expand 4 + 6 rows per node (config 4's mixed plate: K1 at 30 instead
  // of 18 ps per element.step).  FSR_QUAD_PLANAR=2 takes the in-plane form wherever it applies.
  {
    const double extra_rows = (double)nrows - 128.0 * (double)(ntile - (int)tiles.size());
    const double k1_cost = 0.068 * (double)std::max(p->ndim, 1) * extra_rows, k2_gain = 14.0 * (double)nok;
    const bool force = getenv("FSR_QUAD_PLANAR") && atoi(getenv("FSR_QUAD_PLANAR")) == 2;
    if (!force && k1_cost > k2_gain) { cudaFree(f.fast2); f.fast2 = nullptr; return FSR_OK; }
  }
  lst[0].swap(flat_rest);
  lst[2].swap(planar_list);
  p->planar = true;
  p->np_rows = nrows;
  p->np_rows_pad = (nrows + 127) / 128 * 128;
  p->n_k1_tiles = (int)tiles.size();
  FSR_CUDA(cudaMalloc(&p->prow_src, sizeof(int) * src.size()));
  FSR_CUDA(cudaMalloc(&p->prow_w, sizeof(double) * w.size()));
  FSR_CUDA(cudaMalloc(&f.edof2, sizeof(int) * edof2.size()));
  FSR_CUDA(cudaMemcpy(p->prow_src, src.data(), sizeof(int) * src.size(), cudaMemcpyHostToDevice));
  FSR_CUDA(cudaMemcpy(p->prow_w, w.data(), sizeof(double) * w.size(), cudaMemcpyHostToDevice));
  FSR_CUDA(cudaMemcpy(f.edof2, edof2.data(), sizeof(int) * edof2.size(), cudaMemcpyHostToDevice));
  if (!tiles.empty()) {
    FSR_CUDA(cudaMalloc(&p->k1_tiles, sizeof(int) * tiles.size()));
    FSR_CUDA(cudaMemcpy(p->k1_tiles, tiles.data(), sizeof(int) * tiles.size(), cudaMemcpyHostToDevice));
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
                                                            p->rny, p->thk, f.Sfrag, f.failed, f.aux, p->quad_ngauss);
      FSR_LAUNCH_CHECK();
      FSR_CUDA(cudaStreamSynchronize(s));
      cudaFree(d_conn);
      // flat elements: membrane / bending split (k2_quad_flat_vm_kernel); FSR_QUAD_FLAT=0 keeps the dense operator for all
      if (p->quad_ngauss == 2 && !(getenv("FSR_QUAD_FLAT") && atoi(getenv("FSR_QUAD_FLAT")) == 0)) {
        unsigned char* d_flat = nullptr;
        FSR_CUDA(cudaMalloc(&f.fast, sizeof(double) * (size_t)f.nelt * 12 * 32));
        FSR_CUDA(cudaMalloc(&d_flat, f.nelt));
        build_quad_flat_kernel<<<(f.nelt + 63) / 64, 64, 0, s>>>(f.nelt, f.Sfrag, f.failed, f.fast, d_flat);
        FSR_LAUNCH_CHECK();
        std::vector<unsigned char> h((size_t)f.nelt);
        FSR_CUDA(cudaMemcpyAsync(h.data(), d_flat, f.nelt, cudaMemcpyDeviceToHost, s));
        FSR_CUDA(cudaStreamSynchronize(s));
        cudaFree(d_flat);
        std::vector<int> lst[3];
        for (int i = 0; i < f.nelt; ++i) lst[h[(size_t)i] ? 0 : 1].push_back(i);
        if (lst[0].empty()) { cudaFree(f.fast); f.fast = nullptr; }
        else {
          // flat regions: in-plane rows and operators (FSR_QUAD_PLANAR=0 keeps the global rows for all)
          if (!(getenv("FSR_QUAD_PLANAR") && atoi(getenv("FSR_QUAD_PLANAR")) == 0))
            if ((rc = setup_planar_quads(p, sam, elm, f, conn, ptoff, lst, s))) return rc;
          for (int k = 0; k < 3; ++k) {
            f.nsub[k] = (int)lst[k].size();
            if (f.nsub[k] == 0) continue;
            FSR_CUDA(cudaMalloc(&f.sub[k], sizeof(int) * lst[k].size()));
            FSR_CUDA(cudaMemcpy(f.sub[k], lst[k].data(), sizeof(int) * lst[k].size(), cudaMemcpyHostToDevice));
          }
        }
      }
    }
  }
  // ---- triangles (type 23) ----
  {
    FamilyData& f = p->fam[FAM_TRI];
    f.nenod = 3; f.nndof = 6; f.nstrp = 6; f.ncmp = 3; f.MT = 3; f.KT = 5;
    std::vector<int> elem, conn, edof, ptoff;
    int rc = gather_family(p, sam, elm, 23, 3, 6, f.KT, elem, conn, edof, ptoff);
    if (rc) return rc;
    int* d_conn = nullptr;
    rc = upload_family(f, elem, conn, edof, ptoff, 4, &d_conn, s);
    if (rc) return rc;
    if (f.nelt > 0) {
      build_tri_ops_kernel<<<(f.nelt + 31) / 32, 32, 0, s>>>(f.nelt, f.elem, d_conn, p->xyz, p->emod, p->rny,
                                                           p->thk, f.Sfrag, f.failed, f.aux, p->tri_legacy);
      FSR_LAUNCH_CHECK();
      FSR_CUDA(cudaStreamSynchronize(s));
      cudaFree(d_conn);
    }
  }
  return FSR_OK;
}

template <int KT, bool ALL_LIVE>
static int launch_shell_family(fsr_part* p, FamilyData& f, int nsteps, int nsteps_pad, double* vm_dev, size_t ld_vm,
                               cudaStream_t s)
{
  const int warps = 8;
  if (f.nelt == 0) return FSR_OK;
  int ngen = f.nelt;
  const int* gen_list = nullptr;
  if (KT == 6 && f.fast2 && f.nsub[2] > 0) {  // quads of flat regions: in-plane rows
    const unsigned grid = (unsigned)((f.nsub[2] + warps - 1) / warps);
    // U rows through shared memory with cp.async (default; measured 14.58 vs 14.78 ms with the history, 11.6 vs 11.9 without);
    // FSR_PLANAR_STAGED=0 keeps the register prefetch
    static const bool staged = !(getenv("FSR_PLANAR_STAGED") && atoi(getenv("FSR_PLANAR_STAGED")) == 0);
    const size_t st_smem = sizeof(double) * warps * kPlStages * 16 * kPlRow;
    if (staged && vm_dev) {
      if (int rc = smem_opt_in((const void*)k2_quad_planar_vm_kernel<true, true>, st_smem)) return rc;
      k2_quad_planar_vm_kernel<true, true><<<grid, warps * 32, st_smem, s>>>(p->Up, (size_t)p->step_tile, nsteps, nsteps_pad, f.fast2, f.edof2,
                                                                             f.nsub[2], f.sub[2], vm_dev, ld_vm, p->env_max, p->env_min);
    } else if (staged) {
      if (int rc = smem_opt_in((const void*)k2_quad_planar_vm_kernel<false, true>, st_smem)) return rc;
      k2_quad_planar_vm_kernel<false, true><<<grid, warps * 32, st_smem, s>>>(p->Up, (size_t)p->step_tile, nsteps, nsteps_pad, f.fast2, f.edof2,
                                                                              f.nsub[2], f.sub[2], vm_dev, ld_vm, p->env_max, p->env_min);
    } else if (vm_dev)
      k2_quad_planar_vm_kernel<true><<<grid, warps * 32, 0, s>>>(p->Up, (size_t)p->step_tile, nsteps, nsteps_pad, f.fast2, f.edof2, f.nsub[2],
                                                                 f.sub[2], vm_dev, ld_vm, p->env_max, p->env_min);
    else
      k2_quad_planar_vm_kernel<false><<<grid, warps * 32, 0, s>>>(p->Up, (size_t)p->step_tile, nsteps, nsteps_pad, f.fast2, f.edof2, f.nsub[2],
                                                                  f.sub[2], vm_dev, ld_vm, p->env_max, p->env_min);
    FSR_LAUNCH_CHECK();
  }
  if (KT == 6 && f.fast && (f.nsub[0] > 0 || f.nsub[2] > 0)) {   // the other flat ones take the membrane / bending split on global rows
    ngen = f.nsub[1];
    gen_list = f.sub[1];
  }
  if (KT == 6 && f.fast && f.nsub[0] > 0) {
    const unsigned grid = (unsigned)((f.nsub[0] + warps - 1) / warps);
    if (vm_dev)
      k2_quad_flat_vm_kernel<true><<<grid, warps * 32, 0, s>>>(p->U, (size_t)p->step_tile, nsteps, nsteps_pad, f.fast, f.edof, f.ptoff,
                                                               f.nsub[0], f.sub[0], vm_dev, ld_vm, p->env_max, p->env_min);
    else
      k2_quad_flat_vm_kernel<false><<<grid, warps * 32, 0, s>>>(p->U, (size_t)p->step_tile, nsteps, nsteps_pad, f.fast, f.edof, f.ptoff,
                                                                f.nsub[0], f.sub[0], vm_dev, ld_vm, p->env_max, p->env_min);
    FSR_LAUNCH_CHECK();
  }
  if (ngen == 0) return FSR_OK;
  if (vm_dev)
    k2_shell_vm_kernel<KT, true, ALL_LIVE><<<(ngen + warps - 1) / warps, warps * 32, 0, s>>>(
        p->U, (size_t)p->step_tile, nsteps, nsteps_pad, f.Sfrag, f.edof, f.ptoff, f.failed, ngen, f.nstrp, vm_dev,
        ld_vm, p->env_max, p->env_min, nullptr, gen_list);
  else
    k2_shell_vm_kernel<KT, false, ALL_LIVE><<<(ngen + warps - 1) / warps, warps * 32, 0, s>>>(
        p->U, (size_t)p->step_tile, nsteps, nsteps_pad, f.Sfrag, f.edof, f.ptoff, f.failed, ngen, f.nstrp, vm_dev,
        ld_vm, p->env_max, p->env_min, nullptr, gen_list);
  FSR_LAUNCH_CHECK();
  return FSR_OK;
}

// von Mises of the family's elements into float / double step records: out[t * ld_out + roff[i] + point]
template <class OUT_T>
int launch_k2_shell_rec(fsr_part* p, int fam, const double* U, int nsteps, int nsteps_pad, const long long* roff, OUT_T* out, size_t ld_out,
                        cudaStream_t s)
{
  FamilyData& f = p->fam[fam];
  const int warps = 8;
  if (f.nelt == 0) return FSR_OK;
  const unsigned grid = (unsigned)((f.nelt + warps - 1) / warps);
  if (fam == FAM_QUAD)
    k2_shell_vm_kernel<6, true, true, OUT_T, false><<<grid, warps * 32, 0, s>>>(U, (size_t)p->step_tile, nsteps, nsteps_pad, f.Sfrag, f.edof,
                                                                                f.ptoff, f.failed, f.nelt, f.nstrp, out, ld_out, nullptr, nullptr, roff);
  else
    k2_shell_vm_kernel<5, true, false, OUT_T, false><<<grid, warps * 32, 0, s>>>(U, (size_t)p->step_tile, nsteps, nsteps_pad, f.Sfrag, f.edof,
                                                                                 f.ptoff, f.failed, f.nelt, f.nstrp, out, ld_out, nullptr, nullptr, roff);
  FSR_LAUNCH_CHECK();
  return FSR_OK;
}
template int launch_k2_shell_rec<float>(fsr_part*, int, const double*, int, int, const long long*, float*, size_t, cudaStream_t);
template int launch_k2_shell_rec<double>(fsr_part*, int, const double*, int, int, const long long*, double*, size_t, cudaStream_t);

int launch_k2_shell_vm(fsr_part* p, int nsteps, int nsteps_pad, double* vm_dev, size_t ld_vm, cudaStream_t s)
{
  int rc = launch_shell_family<5, false>(p, p->fam[FAM_TRI], nsteps, nsteps_pad, vm_dev, ld_vm, s);
  if (rc) return rc;
  return launch_shell_family<6, true>(p, p->fam[FAM_QUAD], nsteps, nsteps_pad, vm_dev, ld_vm, s);
}

}  // namespace fsr
