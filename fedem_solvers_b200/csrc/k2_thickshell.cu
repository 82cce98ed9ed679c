// k2_thickshell.cu -- thick shells: 6-noded triangle (type 31) and 8-noded quadrilateral (type 32).
//
// Reference: STR31 / STR32 (src/vpmStress/elStressModule.f90:1083-1180, 1186-1302) -> SCTS30 / SCTS32 (src/Femlib/scts.f:1241-1292,
// 1806-2052), SCQS30 / SCQS32 (src/Femlib/scqs.f:770-824, 1345-1604), DNI630 / DNI830, RCOS30, LNCS30, JACO30, CHQT30 / CHQA30,
// tratensor -> FFaTensorTransforms::rotate3D (FFaTensorTransforms.C:369-405).  The reference rebuilds, per element and per time step,
// the node systems, 6 (8) stress matrices SIG(5 x 5*nenod) with their Jacobians, multiplies them with the element vector, rotates the
// local stresses to the global axes and extrapolates to the nodes.  All of that is linear in the element displacement vector, so here it
// is folded ONCE into two dense per-element operators
//     sigma(6 x nstrp) = S[6*nstrp x 6*nenod] . v      epsil = Es[6*nstrp x 6*nenod] . v     (row = point*6 + component)
// (72 x 36 for the triangle, 96 x 48 for the quadrilateral; Es already carries ElStress's tensorial-shear factor, elStressModule.f90:
// 248-252), stored in DMMA A-fragment order.  The per-step work is then a batched small GEMM over 8-step tiles on the FP64 tensor
// cores with the von Mises / envelope epilogue fused (k2_dense6_vm_kernel).
//
// Fortran 77 default-REAL literals of the Femlib sources ("1.2", ".333333333", "1.E-06", "1.E-10", "1.E-03") are REAL*4 constants;
// they are reproduced with float casts exactly as in the CPU checker.
#include "common.cuh"

namespace fsr {

namespace {

__device__ __forceinline__ size_t frag_ix(int row, int col, int KT)
{
  return ((size_t)((row >> 3) * KT + (col >> 2)) << 5) + ((row & 7) << 2) + (col & 3);
}

struct Mat3 { double m[9]; __device__ double& operator()(int i, int j) { return m[(i - 1) + 3 * (j - 1)]; } };  // column-major, 1-based

// LNCS30: x', y' axes from the normal held in column 3
__device__ void lncs30(Mat3& R, int icodir, int& ierr)
{
  int stage = icodir;
  double da;
  if (icodir < 1 || icodir > 3) { ierr = -1; return; }
  bool done = false;
  if (stage == 1) {
    da = sqrt(R(2, 3) * R(2, 3) + R(3, 3) * R(3, 3));
    if (da - (double)1.E-03f < 0.0) { ierr = 1; stage = 2; }
    else {
      R(1, 2) = 0.; R(2, 2) = R(3, 3); R(3, 2) = -R(2, 3);
      R(1, 1) = R(3, 3) * R(3, 3) + R(2, 3) * R(2, 3); R(2, 1) = -R(2, 3) * R(1, 3); R(3, 1) = -R(1, 3) * R(3, 3);
      done = true;
    }
  }
  if (!done && stage == 2) {
    da = sqrt(R(1, 3) * R(1, 3) + R(3, 3) * R(3, 3));
    if (da - (double)1.E-03f < 0.0) { ierr = 2; stage = 3; }
    else {
      R(1, 2) = -R(3, 3); R(2, 2) = 0.; R(3, 2) = R(1, 3);
      R(1, 1) = -R(1, 3) * R(2, 3); R(2, 1) = R(3, 3) * R(3, 3) + R(1, 3) * R(1, 3); R(3, 1) = -R(3, 3) * R(2, 3);
      done = true;
    }
  }
  if (!done) {
    da = sqrt(R(1, 3) * R(1, 3) + R(2, 3) * R(2, 3));
    if (da - (double)1.E-03f < 0.0) { ierr = -3; return; }
    R(1, 2) = R(2, 3); R(2, 2) = -R(1, 3); R(3, 2) = 0.;
    R(1, 1) = -R(1, 3) * R(3, 3); R(2, 1) = -R(2, 3) * R(3, 3); R(3, 1) = R(2, 3) * R(2, 3) + R(1, 3) * R(1, 3);
  }
  const double rl1 = sqrt(R(1, 1) * R(1, 1) + R(2, 1) * R(2, 1) + R(3, 1) * R(3, 1));
  const double rl2 = sqrt(R(1, 2) * R(1, 2) + R(2, 2) * R(2, 2) + R(3, 2) * R(3, 2));
  for (int i = 1; i <= 3; ++i) { R(i, 1) = R(i, 1) / rl1; R(i, 2) = R(i, 2) / rl2; }
}

// RCOS30 with ICODIR > 0: node system n; lambi[n][r][c]
__device__ void rcos30(const double* X, const double* Y, const double* Z, const double* d1, const double* d2, double (*lambi)[3][3],
                       int n, int mek, int icodir, int& ierr)
{
  double a1 = 0., b1 = 0., c1 = 0., a2 = 0., b2 = 0., c2 = 0.;
  for (int k = 0; k < mek; ++k) {
    a1 = a1 + X[k] * d1[k]; b1 = b1 + Y[k] * d1[k]; c1 = c1 + Z[k] * d1[k];
    a2 = a2 + X[k] * d2[k]; b2 = b2 + Y[k] * d2[k]; c2 = c2 + Z[k] * d2[k];
  }
  const double v0 = b1 * c2 - b2 * c1, v1 = a2 * c1 - a1 * c2, v2 = a1 * b2 - a2 * b1;
  double r = sqrt(v0 * v0 + v1 * v1 + v2 * v2);
  if (r - 1.0e-15 <= 0.0) { ierr = -1; return; }
  r = 1 / r;
  Mat3 T;
  for (int k = 0; k < 9; ++k) T.m[k] = 0.0;
  T(1, 3) = v0 * r; T(2, 3) = v1 * r; T(3, 3) = v2 * r;
  lncs30(T, icodir, ierr);
  if (ierr < 0) return;
  for (int c = 0; c < 3; ++c) { lambi[n][0][c] = T.m[c]; lambi[n][1][c] = T.m[3 + c]; lambi[n][2][c] = T.m[6 + c]; }
}

__device__ void dni630(double* d1, double* d2, double* nl, double rl1, double rl2, int lin)
{
  if (lin > 1) {
    nl[0] = 2.0 * rl1 * rl1 - rl1; nl[1] = 2.0 * rl2 * rl2 - rl2;
    nl[2] = 2.0 * rl1 * rl1 + 2.0 * rl2 * rl2 + 4.0 * rl1 * rl2 - 3.0 * rl1 - 3.0 * rl2 + 1.0;
    nl[3] = 4.0 * rl1 * rl2; nl[4] = 4.0 * rl2 - 4.0 * rl1 * rl2 - 4.0 * rl2 * rl2; nl[5] = 4.0 * rl1 - 4.0 * rl1 * rl1 - 4.0 * rl1 * rl2;
  } else if (lin == 1) {
    nl[0] = 0.6 * rl1 - 0.2; nl[1] = 0.6 * rl2 - 0.2; nl[2] = -0.6 * rl1 - 0.6 * rl2 + 0.4;
    nl[3] = 0.8 * rl1 + 0.8 * rl2 - 0.2; nl[4] = -0.8 * rl1 + 0.6; nl[5] = -0.8 * rl2 + 0.6;
  }
  d1[0] = 4.0 * rl1 - 1.0; d2[0] = 0.0; d1[1] = 0.0; d2[1] = 4.0 * rl2 - 1.0;
  d1[2] = 4.0 * rl1 + 4.0 * rl2 - 3.0; d2[2] = d1[2];
  d1[3] = 4.0 * rl2; d2[3] = 4.0 * rl1; d1[4] = -4.0 * rl2; d2[4] = 4.0 - 4.0 * rl1 - 8.0 * rl2;
  d1[5] = 4.0 - 8.0 * rl1 - 4.0 * rl2; d2[5] = -4.0 * rl1;
}

__device__ void dni830(double* dx, double* de, double* nn, double xi, double et, const double* xii, const double* eti)
{
  for (int i = 0; i < 8; ++i) {
    const double xixi = xii[i] * xi, eteti = eti[i] * et;
    double ets, xis;
    if ((i & 1) == 0) {
      ets = .25 * (1. + eteti); xis = .25 * (1. + xixi);
      dx[i] = xii[i] * (2. * xixi + eteti) * ets; de[i] = eti[i] * xis * (2. * eteti + xixi); nn[i] = xis * (1. + eteti) * (xixi + eteti - 1.);
    } else if (i == 1 || i == 5) {
      ets = 1. + eteti; xis = .5 * (1. - xi * xi);
      dx[i] = -xi * ets; de[i] = eti[i] * xis; nn[i] = ets * xis;
    } else {
      ets = (1. - et * et) * .5; xis = 1. + xixi;
      dx[i] = xii[i] * ets; de[i] = -xis * et; nn[i] = xis * ets;
    }
  }
}

// CHQT30 / CHQA30: mid-side nodes inside the middle half of their edges
__device__ bool midside_ok(const double* X, const double* Y, const double* Z, bool tri)
{
  const int nedge = tri ? 3 : 4;
  for (int i = 0; i < nedge; ++i) {
    const int a = tri ? i : 2 * i, m = tri ? 3 + i : 2 * i + 1, b = tri ? (i + 1) % 3 : (2 * i + 2) % 8;
    double dx = X[m] - X[a], dy = Y[m] - Y[a], dz = Z[m] - Z[a];
    const double rl1 = sqrt(dx * dx + dy * dy + dz * dz);
    dx = X[b] - X[m]; dy = Y[b] - Y[m]; dz = Z[b] - Z[m];
    const double rl2 = sqrt(dx * dx + dy * dy + dz * dz);
    if (rl1 - (double)1.E-06f <= 0.0 || rl2 - (double)1.E-06f <= 0.0) return false;
    const double rl = rl1 / rl2;
    if (tri ? (rl - (double).333333333f < 0.0) : (rl - (double).333333333f <= 0.0)) return false;
    if (rl - 3. >= 0.0) return false;
  }
  return true;
}

__device__ void rotate3d(double* S, const double* R)
{
  const double *eX = R, *eY = R + 3, *eZ = R + 6;
  const double TS11 = eX[0] * S[0] + eY[0] * S[3] + eZ[0] * S[4], TS12 = eX[0] * S[3] + eY[0] * S[1] + eZ[0] * S[5],
               TS13 = eX[0] * S[4] + eY[0] * S[5] + eZ[0] * S[2], TS21 = eX[1] * S[0] + eY[1] * S[3] + eZ[1] * S[4],
               TS22 = eX[1] * S[3] + eY[1] * S[1] + eZ[1] * S[5], TS23 = eX[1] * S[4] + eY[1] * S[5] + eZ[1] * S[2],
               TS31 = eX[2] * S[0] + eY[2] * S[3] + eZ[2] * S[4], TS32 = eX[2] * S[3] + eY[2] * S[1] + eZ[2] * S[5],
               TS33 = eX[2] * S[4] + eY[2] * S[5] + eZ[2] * S[2];
  S[0] = TS11 * eX[0] + TS12 * eY[0] + TS13 * eZ[0];
  S[1] = TS21 * eX[1] + TS22 * eY[1] + TS23 * eZ[1];
  S[2] = TS31 * eX[2] + TS32 * eY[2] + TS33 * eZ[2];
  S[3] = TS11 * eX[1] + TS12 * eY[1] + TS13 * eZ[1];
  S[4] = TS11 * eX[2] + TS12 * eY[2] + TS13 * eZ[2];
  S[5] = TS21 * eX[2] + TS22 * eY[2] + TS23 * eZ[2];
}

// The common body of SCTS32 / SCQS32 after the shape function derivatives: sig (5 x 5*mek) column-major, LAMP.
__device__ bool stress_matrix(double* sig, Mat3& L, const double* X, const double* Y, const double* Z, double th, double young,
                              double rny, const double* d1, const double* d2, const double* nn, double ze, double (*lambi)[3][3],
                              int icodir, int mek)
{
  const double D11 = young / (1. - rny * rny), D12 = D11 * rny, D33 = D11 * (1. - rny) * .5, D44 = D33 / (double)1.2f;
  Mat3 J, I;
  for (int k = 0; k < 9; ++k) J.m[k] = 0.0;
  for (int i = 0; i < mek; ++i) {   // JACO30
    const double zet = ze * th * 0.5;
    const double f1 = X[i] + zet * lambi[i][2][0], f2 = Y[i] + zet * lambi[i][2][1], f3 = Z[i] + zet * lambi[i][2][2];
    const double f4 = nn[i] * .5 * th;
    J(1, 1) = J(1, 1) + d1[i] * f1; J(1, 2) = J(1, 2) + d1[i] * f2; J(1, 3) = J(1, 3) + d1[i] * f3;
    J(2, 1) = J(2, 1) + d2[i] * f1; J(2, 2) = J(2, 2) + d2[i] * f2; J(2, 3) = J(2, 3) + d2[i] * f3;
    J(3, 1) = J(3, 1) + f4 * lambi[i][2][0]; J(3, 2) = J(3, 2) + f4 * lambi[i][2][1]; J(3, 3) = J(3, 3) + f4 * lambi[i][2][2];
  }
  {
    const double f1 = J(2, 2) * J(3, 3) - J(2, 3) * J(3, 2), f2 = J(2, 3) * J(3, 1) - J(2, 1) * J(3, 3),
                 f3 = J(2, 1) * J(3, 2) - J(2, 2) * J(3, 1);
    const double detj = J(1, 1) * f1 + J(1, 2) * f2 + J(1, 3) * f3;
    if (fabs(detj) - (double)1.E-10f <= 0.0) return false;
    const double di = 1 / detj;
    I(1, 1) = di * f1; I(2, 1) = di * f2; I(3, 1) = di * f3;
    I(1, 2) = di * (J(3, 2) * J(1, 3) - J(3, 3) * J(1, 2));
    I(1, 3) = di * (J(1, 2) * J(2, 3) - J(1, 3) * J(2, 2));
    I(2, 2) = di * (J(1, 1) * J(3, 3) - J(1, 3) * J(3, 1));
    I(2, 3) = di * (J(2, 1) * J(1, 3) - J(2, 3) * J(1, 1));
    I(3, 2) = di * (J(3, 1) * J(1, 2) - J(3, 2) * J(1, 1));
    I(3, 3) = di * (J(1, 1) * J(2, 2) - J(1, 2) * J(2, 1));
  }
  L(1, 3) = J(1, 2) * J(2, 3) - J(2, 2) * J(1, 3);
  L(2, 3) = J(2, 1) * J(1, 3) - J(1, 1) * J(2, 3);
  L(3, 3) = J(1, 1) * J(2, 2) - J(2, 1) * J(1, 2);
  const double rl = sqrt(L(1, 3) * L(1, 3) + L(2, 3) * L(2, 3) + L(3, 3) * L(3, 3));
  L(1, 3) = L(1, 3) / rl; L(2, 3) = L(2, 3) / rl; L(3, 3) = L(3, 3) / rl;
  int ierr = 0;
  lncs30(L, icodir, ierr);
  if (ierr < 0) return false;
  double A[3][3];
  A[0][0] = L(1, 1) * I(1, 1) + L(2, 1) * I(2, 1) + L(3, 1) * I(3, 1);
  A[0][1] = L(1, 1) * I(1, 2) + L(2, 1) * I(2, 2) + L(3, 1) * I(3, 2);
  A[1][0] = L(1, 2) * I(1, 1) + L(2, 2) * I(2, 1) + L(3, 2) * I(3, 1);
  A[1][1] = L(1, 2) * I(1, 2) + L(2, 2) * I(2, 2) + L(3, 2) * I(3, 2);
  A[2][0] = L(1, 3) * I(1, 1) + L(2, 3) * I(2, 1) + L(3, 3) * I(3, 1);
  A[2][1] = L(1, 3) * I(1, 2) + L(2, 3) * I(2, 2) + L(3, 3) * I(3, 2);
  A[2][2] = L(1, 3) * I(1, 3) + L(2, 3) * I(2, 3) + L(3, 3) * I(3, 3);
  for (int i = 0; i < mek; ++i) {
    const double B0 = A[0][0] * d1[i] + A[0][1] * d2[i], B1 = A[1][0] * d1[i] + A[1][1] * d2[i], B2 = A[2][0] * d1[i] + A[2][1] * d2[i];
    const double C = A[2][2] * nn[i];
    double A1[5][3], A2[5][2], A4[2][2];
    for (int c = 1; c <= 3; ++c) {
      A1[0][c - 1] = L(c, 1) * B0;
      A1[1][c - 1] = L(c, 2) * B1;
      A1[2][c - 1] = L(c, 1) * B1 + L(c, 2) * B0;
      A1[3][c - 1] = L(c, 3) * B0 + L(c, 1) * B2;
      A1[4][c - 1] = L(c, 3) * B1 + L(c, 2) * B2;
    }
    for (int r = 0; r < 5; ++r) {
      A2[r][1] = lambi[i][0][0] * A1[r][0] + lambi[i][0][1] * A1[r][1] + lambi[i][0][2] * A1[r][2];
      A2[r][0] = -lambi[i][1][0] * A1[r][0] - lambi[i][1][1] * A1[r][1] - lambi[i][1][2] * A1[r][2];
    }
    for (int r = 0; r < 2; ++r) {   // rows 4, 5 of A3 = C * LAMP(:, r+1); A4 = A3 * FI
      const double a0 = L(1, r + 1) * C, a1 = L(2, r + 1) * C, a2 = L(3, r + 1) * C;
      A4[r][1] = lambi[i][0][0] * a0 + lambi[i][0][1] * a1 + lambi[i][0][2] * a2;
      A4[r][0] = -lambi[i][1][0] * a0 - lambi[i][1][1] * a1 - lambi[i][1][2] * a2;
    }
    const double thh = .5 * th;
    double* s = sig + 25 * i;   // columns 5i .. 5i+4, 5 rows each
    for (int c = 0; c < 3; ++c) {
      s[0 + 5 * c] = A1[0][c] * D11 + A1[1][c] * D12;
      s[1 + 5 * c] = A1[0][c] * D12 + A1[1][c] * D11;
      s[2 + 5 * c] = A1[2][c] * D33;
      s[3 + 5 * c] = A1[3][c] * D44;
      s[4 + 5 * c] = A1[4][c] * D44;
    }
    for (int c = 0; c < 2; ++c) {
      const double A7[5] = {A2[0][c] * D11 + A2[1][c] * D12, A2[0][c] * D12 + A2[1][c] * D11, A2[2][c] * D33, A2[3][c] * D44, A2[4][c] * D44};
      for (int r = 0; r < 5; ++r) {
        double v = thh * ze * A7[r];
        if (r >= 3) v = v + thh * (A4[r - 3][c] * D44);
        s[r + 5 * (3 + c)] = v;
      }
    }
  }
  return true;
}

// extrapolation weights of STR31 (:1163-1168) / STR32 (:1275-1289): result point p of a surface = sum_g W[p][g] * sampling point g
template <int NEN>
__device__ void extrapolation_weights(double (*W)[NEN == 6 ? 3 : 4])
{
  constexpr int NG = NEN == 6 ? 3 : 4;
  if constexpr (NEN == 6) {
    const double w[6][3] = {{1, -1, 1}, {1, 1, -1}, {-1, 1, 1}, {1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    for (int p = 0; p < 6; ++p) for (int g = 0; g < 3; ++g) W[p][g] = w[p][g];
  } else {
    const double sq3 = sqrt(3.0), f1 = 0.5 + 0.5 * sq3, f2 = 0.5 - 0.5 * sq3;
    for (int p = 0; p < NEN; ++p) for (int g = 0; g < NG; ++g) W[p][g] = 0.0;
    // sampling point g = (j-1)*2 + (i-1) of the (xi_i, eta_j) loop
    W[0][0] = f1; W[0][3] = f2; W[2][1] = f1; W[2][2] = f2; W[4][3] = f1; W[4][0] = f2; W[6][2] = f1; W[6][1] = f2;
    for (int p = 1; p < NEN; p += 2) for (int g = 0; g < NG; ++g) W[p][g] = 0.5 * (W[p - 1][g] + W[(p + 1) % NEN][g]);
  }
}

// One thread per element: node systems, the 2 x NG sampling-point stress matrices, global rotation, extrapolation weights.
template <int NEN>
__global__ void build_thickshell_ops_kernel(int nelt, const int* __restrict__ elem, const int* __restrict__ conn,
                                            const double* __restrict__ xyz, const double* __restrict__ emod,
                                            const double* __restrict__ rny, const double* __restrict__ thk,
                                            double* __restrict__ Sfrag, double* __restrict__ Efrag, double* __restrict__ Gfrag,
                                            unsigned char* __restrict__ failed, double* __restrict__ aux)
{
  constexpr bool TRI = NEN == 6;
  constexpr int NG = TRI ? 3 : 4, NCOL = 6 * NEN, NROW = 12 * NEN;
  constexpr int MT = (NROW + 7) / 8, KT = (NCOL + 3) / 4, MTG = (12 * NG + 7) / 8;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nelt) return;
  const int e = elem[i];
  double* S = Sfrag + (size_t)i * MT * KT * 32;
  double* Es = Efrag + (size_t)i * MT * KT * 32;
  double* Gs = Gfrag + (size_t)i * MTG * KT * 32;   // global stresses at the 2 x NG sampling points: row = (k*NG + g)*6 + component
  double X[NEN], Y[NEN], Z[NEN];
  for (int k = 0; k < NEN; ++k) {
    const int n = conn[i * NEN + k];
    X[k] = xyz[3 * n]; Y[k] = xyz[3 * n + 1]; Z[k] = xyz[3 * n + 2];
  }
  const double E = emod[e], nu = rny[e], th = thk[e];
  aux[i * 3] = E; aux[i * 3 + 1] = nu; aux[i * 3 + 2] = th;
  // inverse constitutive matrix of STR31 / STR32 (isoMat2Dinv + the 1.2 shear factor, in double there)
  const double C11 = 1.0 / E, C12 = -nu / E, C33 = 2.0 * (1.0 + nu) / E, C44 = C33 * 1.2;
  const double XII[8] = {-1.0, 0.0, 1.0, 1.0, 1.0, 0.0, -1.0, -1.0}, ETI[8] = {-1.0, -1.0, -1.0, 0.0, 1.0, 1.0, 1.0, 0.0};
  const double RL1n[6] = {1.0, 0.0, 0.0, 0.5, 0.0, 0.5}, RL2n[6] = {0.0, 1.0, 0.0, 0.5, 0.5, 0.0};
  double lambi[NEN][3][3], d1[NEN], d2[NEN], nn[NEN];
  bool ok = true;
  int ierr = 0;
  for (int n = 0; n < NEN && ok; ++n) {   // SCTS30 / SCQS30
    if constexpr (TRI) dni630(d1, d2, nn, RL1n[n], RL2n[n], 0);
    else dni830(d1, d2, nn, XII[n], ETI[n], XII, ETI);
    rcos30(X, Y, Z, d1, d2, lambi, n, NEN, 1, ierr);
    if (ierr < 0) ok = false;
  }
  ok = ok && midside_ok(X, Y, Z, TRI);
  // extrapolation weights: result point p of a surface = sum_g W[p][g] * sampling point g
  double W[NEN][NG];
  extrapolation_weights<NEN>(W);
  double sig[5 * 5 * NEN];
  const double sq3 = sqrt(3.0);
  for (int k = 0; k < 2 && ok; ++k) {
    const double zeta = (double)(1 - 2 * k);
    for (int g = 0; g < NG && ok; ++g) {
      if constexpr (TRI) {
        const double L1[3] = {0.5, 0.0, 0.5}, L2[3] = {0.5, 0.5, 0.0};
        dni630(d1, d2, nn, L1[g], L2[g], 1);
      } else {
        const double xi = (double)(2 * (g & 1) - 1) / sq3, eta = (double)(2 * (g >> 1) - 1) / sq3;
        dni830(d1, d2, nn, xi, eta, XII, ETI);
      }
      Mat3 L;
      if (!stress_matrix(sig, L, X, Y, Z, th, E, nu, d1, d2, nn, zeta, lambi, 1, NEN)) { ok = false; break; }
      for (int j = 0; j < NCOL; ++j) {   // element DOF j = 6*node + d: column of SIG . (local 5-DOF transformation)
        const int n = j / 6, d = j % 6;
        double s5[5];
        for (int r = 0; r < 5; ++r)
          s5[r] = d < 3 ? sig[r + 5 * (5 * n + d)]
                        : sig[r + 5 * (5 * n + 3)] * lambi[n][0][d - 3] + sig[r + 5 * (5 * n + 4)] * lambi[n][1][d - 3];
        double s6[6] = {s5[0], s5[1], 0.0, s5[2], s5[3], s5[4]};
        double e6[6] = {C11 * s5[0] + C12 * s5[1], C12 * s5[0] + C11 * s5[1], 0.0, C33 * s5[2], C44 * s5[3], C44 * s5[4]};
        rotate3d(s6, L.m);
        rotate3d(e6, L.m);
        e6[3] *= 0.5; e6[4] *= 0.5; e6[5] *= 0.5;   // tensorial shear strain (ElStress, elStressModule.f90:248-252)
        for (int c = 0; c < 6; ++c) Gs[frag_ix((k * NG + g) * 6 + c, j, KT)] = s6[c];
        for (int p = 0; p < NEN; ++p) {
          const double w = W[p][g];
          if (w == 0.0) continue;
          for (int c = 0; c < 6; ++c) {
            const size_t ix = frag_ix((k * NEN + p) * 6 + c, j, KT);
            S[ix] += w * s6[c];
            Es[ix] += w * e6[c];
          }
        }
      }
    }
  }
  if (!ok) {
    for (int k = 0; k < MT * KT * 32; ++k) { S[k] = 0.0; Es[k] = 0.0; }
    for (int k = 0; k < MTG * KT * 32; ++k) Gs[k] = 0.0;
  }
  failed[i] = ok ? 0 : 1;
}

// Dense apply for elements with 6-component stresses at NPT points and NCOL element DOFs: one block per element, the operator
// fragments resident in shared memory, 8-step tiles of U staged in 64-byte segments, von Mises + envelope fused.
template <int NPT, int NCOL>
__global__ void __launch_bounds__(128)
k2_dense6_vm_kernel(const double* __restrict__ U, size_t ldu, int nsteps, int nsteps_pad, const double* __restrict__ Sfrag,
                    const int* __restrict__ edof, const int* __restrict__ ptoff, const unsigned char* __restrict__ failed, int nelt,
                    double* __restrict__ vm, size_t ld_vm, double* __restrict__ env_max, double* __restrict__ env_min)
{
  constexpr int NROW = 6 * NPT;
  constexpr int MT = (NROW + 7) / 8, KT = (NCOL + 3) / 4;
  constexpr int NITEM = NPT * 8, IPT = (NITEM + 127) / 128;
  extern __shared__ __align__(16) double smem[];
  double* sS = smem;                       // [MT][KT][32]
  double* sU = sS + MT * KT * 32;          // [KT*4][8]
  double* sSig = sU + KT * 4 * 8;          // [MT*8][8]
  int* sDof = reinterpret_cast<int*>(sSig + MT * 8 * 8);
  const int i = blockIdx.x;
  if (i >= nelt) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t4 = lane & 3;
  const double* src = Sfrag + (size_t)i * MT * KT * 32;
  for (int k = tid; k < MT * KT * 32; k += 128) sS[k] = src[k];
  for (int k = tid; k < KT * 4; k += 128) sDof[k] = edof[(size_t)i * KT * 4 + k];
  const bool bad = failed[i] != 0;
  const size_t pt0 = (size_t)ptoff[i];
  double emax[IPT], emin[IPT];
#pragma unroll
  for (int r = 0; r < IPT; ++r) { emax[r] = 0.0; emin[r] = kHuge; }
  __syncthreads();
  const int ntiles = nsteps_pad >> 3;
  for (int nt = 0; nt < ntiles && nt * 8 < nsteps; ++nt) {
    for (int k = tid; k < KT * 4 * 8; k += 128) {
      const int row = k >> 3, s = k & 7;
      sU[k] = row < NCOL ? U[(size_t)sDof[row] * ldu + (size_t)nt * 8 + s] : 0.0;
    }
    __syncthreads();
    for (int m = warp; m < MT; m += 4) {
      double c0 = 0.0, c1 = 0.0;
#pragma unroll 4
      for (int j = 0; j < KT; ++j) dmma884(c0, c1, sS[(m * KT + j) * 32 + lane], sU[(4 * j + t4) * 8 + g]);
      *reinterpret_cast<double2*>(sSig + (m * 8 + g) * 8 + 2 * t4) = make_double2(c0, c1);
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < IPT; ++r) {
      const int idx = tid + 128 * r;
      if (idx < NITEM) {
        const int p = idx >> 3, s = idx & 7, t = nt * 8 + s;
        const double* s6 = sSig + (p * 6) * 8 + s;
        const double s11 = s6[0], s22 = s6[8], s33 = s6[16], s12 = s6[24], s13 = s6[32], s23 = s6[40];
        double v = sqrt(s11 * s11 + s22 * s22 + s33 * s33 - s11 * s22 - s22 * s33 - s33 * s11 + 3.0 * (s12 * s12 + s13 * s13 + s23 * s23));
        if (bad) v = kHuge;
        if (t < nsteps) {
          if (vm) vm[(size_t)t * ld_vm + pt0 + p] = v;
          emax[r] = fmax(emax[r], v);
          emin[r] = fmin(emin[r], v);
        }
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int r = 0; r < IPT; ++r) {
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) {
      emax[r] = fmax(emax[r], __shfl_xor_sync(0xffffffffu, emax[r], o));
      emin[r] = fmin(emin[r], __shfl_xor_sync(0xffffffffu, emin[r], o));
    }
    const int idx = tid + 128 * r;
    if (idx < NITEM && (idx & 7) == 0 && nsteps > 0) {
      const int p = idx >> 3;
      if (emax[r] > env_max[pt0 + p]) env_max[pt0 + p] = emax[r];
      if (emin[r] < env_min[pt0 + p]) env_min[pt0 + p] = emin[r];
    }
  }
}


// The throughput kernel: von Mises + envelope from the SAMPLING-POINT operator (2 x NG points, 36 / 48 rows instead of the 72 / 96
// rows of the nodal operator), the extrapolation to the nodes (STR31 :1163-1168, STR32 :1275-1289) done in the epilogue in the order
// the reference does it (stress at the sampling points first, then the weighted sums).  One block of 4 warps per element, operator
// fragments resident in shared memory (13 / 18 KB), 8-step tiles of U; every warp keeps the B fragment of a k-tile for its two m-tiles.
template <int NEN>
__global__ void __launch_bounds__(128)
k2_thick_gauss_vm_kernel(const double* __restrict__ U, size_t ldu, int nsteps, int nsteps_pad, const double* __restrict__ Gfrag,
                         const int* __restrict__ edof, const int* __restrict__ ptoff, const unsigned char* __restrict__ failed, int nelt,
                         double* __restrict__ vm, size_t ld_vm, double* __restrict__ env_max, double* __restrict__ env_min)
{
  constexpr int NG = NEN == 6 ? 3 : 4, NCOL = 6 * NEN, NPT = 2 * NEN;
  constexpr int MTG = (12 * NG + 7) / 8, KT = (NCOL + 3) / 4;
  constexpr int NITEM = NPT * 8;   // 96 / 128 (point, step) pairs per tile
  extern __shared__ __align__(16) double smem[];
  double* sS = smem;                       // [MTG][KT][32]
  double* sU = sS + MTG * KT * 32;         // [2][KT*4][8]   double buffered displacement tile
  double* sSig = sU + 2 * KT * 4 * 8;      // [MTG*8][8]
  double* sW = sSig + MTG * 8 * 8;         // [NEN][NG]
  int* sDof = reinterpret_cast<int*>(sW + NEN * NG);
  const int i = blockIdx.x;
  if (i >= nelt) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t4 = lane & 3;
  const double* src = Gfrag + (size_t)i * MTG * KT * 32;
  for (int k = tid; k < MTG * KT * 32; k += 128) sS[k] = src[k];
  for (int k = tid; k < KT * 4; k += 128) sDof[k] = edof[(size_t)i * KT * 4 + k];
  if (tid == 0) extrapolation_weights<NEN>(reinterpret_cast<double (*)[NG]>(sW));
  const bool bad = failed[i] != 0;
  const size_t pt0 = (size_t)ptoff[i];
  double emax = 0.0, emin = kHuge;
  __syncthreads();
  const int ntiles = min(nsteps_pad >> 3, (nsteps + 7) >> 3);
  constexpr int NLD = (KT * 4 * 8 + 127) / 128;   // staged doubles per thread and tile
  double r[NLD];
  auto fetch = [&](int nt) {   // global -> registers; consumed by put() after the DMMAs so that the loads fly under them
#pragma unroll
    for (int q = 0; q < NLD; ++q) {
      const int k = tid + 128 * q, row = k >> 3, st = k & 7;
      r[q] = (k < KT * 4 * 8 && row < NCOL) ? U[(size_t)sDof[row] * ldu + (size_t)nt * 8 + st] : 0.0;
    }
  };
  auto put = [&](int buf) {
    double* d = sU + buf * KT * 4 * 8;
#pragma unroll
    for (int q = 0; q < NLD; ++q) { const int k = tid + 128 * q; if (k < KT * 4 * 8) d[k] = r[q]; }
  };
  if (ntiles > 0) { fetch(0); put(0); }
  __syncthreads();
  const int p = tid >> 3, s = tid & 7;       // epilogue item of this thread
  const int ksurf = p / NEN, pn = p % NEN;
  for (int nt = 0; nt < ntiles; ++nt) {
    const double* u = sU + (nt & 1) * KT * 4 * 8;
    if (nt + 1 < ntiles) fetch(nt + 1);
    const int m0 = warp, m1 = warp + 4;                     // MTG <= 6: at most two m-tiles per warp
    double c00 = 0.0, c01 = 0.0, c10 = 0.0, c11 = 0.0;
#pragma unroll 3
    for (int j = 0; j < KT; ++j) {
      const double b = u[(4 * j + t4) * 8 + g];
      dmma884(c00, c01, sS[(m0 * KT + j) * 32 + lane], b);
      if (m1 < MTG) dmma884(c10, c11, sS[(m1 * KT + j) * 32 + lane], b);   // warp-uniform
    }
    *reinterpret_cast<double2*>(sSig + (m0 * 8 + g) * 8 + 2 * t4) = make_double2(c00, c01);
    if (m1 < MTG) *reinterpret_cast<double2*>(sSig + (m1 * 8 + g) * 8 + 2 * t4) = make_double2(c10, c11);
    if (nt + 1 < ntiles) put((nt + 1) & 1);   // the other buffer: last read two barriers ago
    __syncthreads();
    if (tid < NITEM) {
      double sg[6] = {0, 0, 0, 0, 0, 0};
#pragma unroll
      for (int q = 0; q < NG; ++q) {
        const double w = sW[pn * NG + q];
        const double* r = sSig + ((ksurf * NG + q) * 6) * 8 + s;
#pragma unroll
        for (int c = 0; c < 6; ++c) sg[c] += w * r[c * 8];
      }
      double v = sqrt(sg[0] * sg[0] + sg[1] * sg[1] + sg[2] * sg[2] - sg[0] * sg[1] - sg[1] * sg[2] - sg[2] * sg[0] +
                      3.0 * (sg[3] * sg[3] + sg[4] * sg[4] + sg[5] * sg[5]));
      if (bad) v = kHuge;
      const int t = nt * 8 + s;
      if (t < nsteps) {
        if (vm) vm[(size_t)t * ld_vm + pt0 + p] = v;
        emax = fmax(emax, v);
        emin = fmin(emin, v);
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int o = 1; o < 8; o <<= 1) {
    emax = fmax(emax, __shfl_xor_sync(0xffffffffu, emax, o));
    emin = fmin(emin, __shfl_xor_sync(0xffffffffu, emin, o));
  }
  if (tid < NITEM && s == 0 && nsteps > 0) {
    if (emax > env_max[pt0 + p]) env_max[pt0 + p] = emax;
    if (emin < env_min[pt0 + p]) env_min[pt0 + p] = emin;
  }
}

template <int NEN>
constexpr size_t thick_gauss_smem()
{
  constexpr int NG = NEN == 6 ? 3 : 4, KT = (6 * NEN + 3) / 4, MTG = (12 * NG + 7) / 8;
  return sizeof(double) * (MTG * KT * 32 + 2 * KT * 4 * 8 + MTG * 8 * 8 + NEN * NG) + sizeof(int) * KT * 4;
}

template <int NPT, int NCOL>
constexpr size_t dense6_smem()
{
  return sizeof(double) * (((6 * NPT + 7) / 8) * ((NCOL + 3) / 4) * 32 + ((NCOL + 3) / 4) * 4 * 8 + ((6 * NPT + 7) / 8) * 8 * 8) +
         sizeof(int) * ((NCOL + 3) / 4) * 4;
}

template <int NEN>
int build_family(fsr_part* p, const fsr_sam* sam, const fsr_elmdata* elm, int fam, int type, const char* name)
{
  cudaStream_t s = p->stream;
  FamilyData& f = p->fam[fam];
  f.nenod = NEN; f.nndof = 6; f.nstrp = 2 * NEN; f.ncmp = 6; f.MT = (12 * NEN + 7) / 8; f.KT = (6 * NEN + 3) / 4; f.naux = 3;
  std::vector<int> elem, conn, edof, ptoff;
  const int estride = f.KT * 4;
  for (int e : elements_of_type(p, sam, elm, type)) {
    const int ip0 = sam->mpmnpc[e] - 1, nn = sam->mpmnpc[e + 1] - sam->mpmnpc[e];
    if (nn != NEN) { set_error("%s element %d has %d nodes", name, e + 1, nn); return FSR_ERR_ARG; }
    elem.push_back(e);
    ptoff.push_back(p->ptoff_host[e]);
    const size_t base = edof.size();
    edof.resize(base + estride, 0);
    for (int k = 0; k < NEN; ++k) {
      const int n = sam->mmnpc[ip0 + k] - 1;
      if (n < 0 || n >= sam->nnod) { set_error("element %d: node index out of range", e + 1); return FSR_ERR_ARG; }
      conn.push_back(n);
      const int js = sam->madof[n] - 1, nd = sam->madof[n + 1] - sam->madof[n];
      if (nd < 6) { set_error("element %d: node %d has %d DOFs, a shell needs 6", e + 1, n + 1, nd); return FSR_ERR_ARG; }
      for (int d = 0; d < 6; ++d) edof[base + (size_t)k * 6 + d] = js + d;
    }
  }
  f.nelt = (int)elem.size();
  if (f.nelt == 0) return FSR_OK;
  const size_t opsz = sizeof(double) * (size_t)f.nelt * f.MT * f.KT * 32;
  int* d_conn = nullptr;
  FSR_CUDA(cudaMalloc(&f.elem, sizeof(int) * elem.size()));
  FSR_CUDA(cudaMalloc(&f.edof, sizeof(int) * edof.size()));
  FSR_CUDA(cudaMalloc(&f.ptoff, sizeof(int) * ptoff.size()));
  FSR_CUDA(cudaMalloc(&f.failed, f.nelt));
  FSR_CUDA(cudaMalloc(&f.Sfrag, opsz));
  FSR_CUDA(cudaMalloc(&f.Efrag, opsz));
  const size_t gsz = sizeof(double) * (size_t)f.nelt * ((12 * (NEN == 6 ? 3 : 4) + 7) / 8) * f.KT * 32;
  FSR_CUDA(cudaMalloc(&f.Gfrag, gsz));
  FSR_CUDA(cudaMemsetAsync(f.Gfrag, 0, gsz, s));
  FSR_CUDA(cudaMalloc(&f.aux, sizeof(double) * (size_t)f.nelt * f.naux));
  FSR_CUDA(cudaMalloc(&d_conn, sizeof(int) * conn.size()));
  FSR_CUDA(cudaMemcpyAsync(f.elem, elem.data(), sizeof(int) * elem.size(), cudaMemcpyHostToDevice, s));
  FSR_CUDA(cudaMemcpyAsync(f.edof, edof.data(), sizeof(int) * edof.size(), cudaMemcpyHostToDevice, s));
  FSR_CUDA(cudaMemcpyAsync(f.ptoff, ptoff.data(), sizeof(int) * ptoff.size(), cudaMemcpyHostToDevice, s));
  FSR_CUDA(cudaMemcpyAsync(d_conn, conn.data(), sizeof(int) * conn.size(), cudaMemcpyHostToDevice, s));
  FSR_CUDA(cudaMemsetAsync(f.Sfrag, 0, opsz, s));
  FSR_CUDA(cudaMemsetAsync(f.Efrag, 0, opsz, s));
  build_thickshell_ops_kernel<NEN><<<(f.nelt + 31) / 32, 32, 0, s>>>(f.nelt, f.elem, d_conn, p->xyz, p->emod, p->rny, p->thk, f.Sfrag,
                                                                   f.Efrag, f.Gfrag, f.failed, f.aux);
  FSR_LAUNCH_CHECK();
  FSR_CUDA(cudaStreamSynchronize(s));
  cudaFree(d_conn);
  return FSR_OK;
}

template <int NPT, int NCOL>
int launch_family(fsr_part* p, int fam, int nsteps, int nsteps_pad, double* vm_dev, size_t ld_vm, cudaStream_t s)
{
  FamilyData& f = p->fam[fam];
  if (f.nelt == 0) return FSR_OK;
  // FSR_THICK_DENSE=1 selects the nodal (dense 72x36 / 96x48) operator kernel (A/B timing, cross-check)
  static const bool dense = getenv("FSR_THICK_DENSE") && atoi(getenv("FSR_THICK_DENSE")) != 0;
  if (!dense) {
    constexpr int NEN = NPT / 2;
    constexpr size_t gsmem = thick_gauss_smem<NEN>();
    if (int rc = smem_opt_in((const void*)k2_thick_gauss_vm_kernel<NEN>, gsmem)) return rc;
    k2_thick_gauss_vm_kernel<NEN><<<f.nelt, 128, gsmem, s>>>(p->U, (size_t)p->step_tile, nsteps, nsteps_pad, f.Gfrag, f.edof, f.ptoff,
                                                            f.failed, f.nelt, vm_dev, ld_vm, p->env_max, p->env_min);
    FSR_LAUNCH_CHECK();
    return FSR_OK;
  }
  constexpr size_t smem = dense6_smem<NPT, NCOL>();
  if (int rc = smem_opt_in((const void*)k2_dense6_vm_kernel<NPT, NCOL>, smem)) return rc;
  k2_dense6_vm_kernel<NPT, NCOL><<<f.nelt, 128, smem, s>>>(p->U, (size_t)p->step_tile, nsteps, nsteps_pad, f.Sfrag, f.edof, f.ptoff,
                                                          f.failed, f.nelt, vm_dev, ld_vm, p->env_max, p->env_min);
  FSR_LAUNCH_CHECK();
  return FSR_OK;
}

}  // namespace

int build_thickshell_operators(fsr_part* p, const fsr_sam* sam, const fsr_elmdata* elm)
{
  int rc = build_family<6>(p, sam, elm, FAM_TRI6, 31, "TRI6");
  if (rc) return rc;
  return build_family<8>(p, sam, elm, FAM_QUAD8, 32, "QUAD8");
}

int launch_k2_thickshell_vm(fsr_part* p, int nsteps, int nsteps_pad, double* vm_dev, size_t ld_vm, cudaStream_t s)
{
  int rc = launch_family<12, 36>(p, FAM_TRI6, nsteps, nsteps_pad, vm_dev, ld_vm, s);
  if (rc) return rc;
  return launch_family<16, 48>(p, FAM_QUAD8, nsteps, nsteps_pad, vm_dev, ld_vm, s);
}

}  // namespace fsr
