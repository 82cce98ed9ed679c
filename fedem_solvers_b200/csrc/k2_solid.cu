// k2_solid.cu -- K2 for the 10-node tetrahedron (type 41) on sm_100a.
//
// Reference: STR41 -> ITET32 -> DN1031 / JACI31 (src/vpmStress/elStressModule.f90:1308-1375,
// src/Femlib/itet.f:7-87,701-997, src/Femlib/jaci31.f) re-evaluates ten 3x3 Jacobian inverses per
// element per step.  Here the 60x30 stress operator sigma(6,10) = S_e . v(3,10) is built once
// (default -stressForm 0: direct evaluation at the nodes; otherwise 4 Gauss points with the
// reference's REAL*4 abscissae, extrapolated with alpha_p/beta_p) and applied per step tile with
// DMMA.8x8x4: rows packed densely (row = 6*node + component, 60 of 64 used, K = 30 of 32), the
// accumulators are transposed through shared memory so that one lane sees all six components of
// a node for the von Mises evaluation (FFaTensorTransforms.C:38-43) and the fused envelope.
//
// The von Mises + envelope kernel (the throughput path) does not apply that 60x30 operator.  sigma = D.B.v is
// D . sym(grad u), and grad u at the 10 result points is [30 x 10] . [10 x 3]: the displacement-gradient
// operator G (rows = node x derivative direction, columns = nodes; for Gauss-point extrapolation the
// extrapolation weights are folded in, everything stays linear) applied to the nodal displacements as three
// right-hand sides u, v, w.  That is 2*30*10*3 = 1,800 flops per element.step instead of 3,600, 12 A-fragments
// in registers instead of 64, and the isotropic D (E, nu) is applied in the epilogue on the accumulators.
// Rows are ordered so that lane (g, t4) of the DMMA accumulator layout owns all nine gradient entries of
// node g at its two steps (m-tile d holds d/dx_d of nodes 0..7); nodes 8 and 9 share the fourth m-tile and
// are completed with two warp shuffles.  The 60x30 operator is kept for the full-result / record kernels.
#include <cstdlib>

#include "common.cuh"

namespace fsr {

__device__ __forceinline__ size_t frag_index8(int row, int col, int KT)
{
  return ((size_t)((row >> 3) * KT + (col >> 2)) << 5) + ((row & 7) << 2) + (col & 3);
}

// shape-function derivatives w.r.t. the volume coordinates L1..L3 (L4 eliminated), itet.f:48-84
__device__ void tet10_dn(double L1, double L2, double L3, double L4, double d1[10], double d2[10],
                         double d3[10])
{
  d1[0] = 4. * L1 - 1.; d1[1] = 4. * L2; d1[2] = 0.; d1[3] = 0.; d1[4] = 0.; d1[5] = 4. * L3;
  d1[6] = 4. * (L4 - L1); d1[7] = -4. * L2; d1[8] = -4. * L3; d1[9] = -4. * L4 + 1.;
  d2[0] = 0.; d2[1] = 4. * L1; d2[2] = 4. * L2 - 1.; d2[3] = 4. * L3; d2[4] = 0.; d2[5] = 0.;
  d2[6] = -4. * L1; d2[7] = 4. * (L4 - L2); d2[8] = -4. * L3; d2[9] = -4. * L4 + 1.;
  d3[0] = 0.; d3[1] = 0.; d3[2] = 0.; d3[3] = 4. * L2; d3[4] = 4. * L3 - 1.; d3[5] = 4. * L1;
  d3[6] = -4. * L1; d3[7] = -4. * L2; d3[8] = 4. * L4 - 4. * L3; d3[9] = -4. * L4 + 1.;
}

// row of the gradient operator for d/dx_d at result point p: m-tile d, row p for the first eight points;
// points 8 and 9 share m-tile 3 (rows 3*(p-8)+d)
__host__ __device__ __forceinline__ int tet10_grad_row(int p, int d) { return p < 8 ? d * 8 + p : 24 + (p - 8) * 3 + d; }

struct Tet10Points {
  int npt;            // evaluation points (10 nodes, or 4 Gauss points)
  double L[10][3];    // volume coordinates L1..L3 of each evaluation point
  double W[10][10];   // sigma(node p) = sum_g W[p][g] * sigma(evaluation point g)
};

__global__ void build_tet10_ops_kernel(int nelt, const int* __restrict__ elem,
                                       const int* __restrict__ conn /* [nelt][10] */,
                                       const double* __restrict__ xyz, const double* __restrict__ emod,
                                       const double* __restrict__ rny, const Tet10Points* __restrict__ pts,
                                       double* __restrict__ Sfrag, unsigned char* __restrict__ failed,
                                       double* __restrict__ aux, double* __restrict__ Gfrag,
                                       double* __restrict__ fast /* [nelt][10]: J^-1 (row d, column j) + flag */,
                                       double* __restrict__ fastJ /* [nelt][10][9]: J^-1 of every nodal evaluation point, or NULL */)
{
  const int KT = 8;
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nelt) return;
  const int e = elem[i];
  double* S = Sfrag + (size_t)i * 8 * KT * 32;
  double* G = Gfrag + (size_t)i * 12 * 32;
  double X[10], Y[10], Z[10];
  for (int k = 0; k < 10; ++k) {
    int n = conn[i * 10 + k];
    X[k] = xyz[3 * n]; Y[k] = xyz[3 * n + 1]; Z[k] = xyz[3 * n + 2];
  }
  const double E = emod[e], nu = rny[e];
  aux[i * 2] = E; aux[i * 2 + 1] = nu;
  const double D = E * (1. - nu) / ((1. + nu) * (1. - 2. * nu));
  const double D1 = D * nu / (1. - nu);
  const double D2 = D * (1. - 2. * nu) / (2. * (1. - nu));
  bool ok = true;
  const int npt = pts->npt;

  for (int gpt = 0; gpt < npt; ++gpt) {
    const double L1 = pts->L[gpt][0], L2 = pts->L[gpt][1], L3 = pts->L[gpt][2];
    const double L4 = 1.0 - L1 - L2 - L3;
    double d1[10], d2[10], d3[10];
    tet10_dn(L1, L2, L3, L4, d1, d2, d3);
    double J[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
    for (int k = 0; k < 10; ++k) {
      J[0][0] += d1[k] * X[k]; J[0][1] += d1[k] * Y[k]; J[0][2] += d1[k] * Z[k];
      J[1][0] += d2[k] * X[k]; J[1][1] += d2[k] * Y[k]; J[1][2] += d2[k] * Z[k];
      J[2][0] += d3[k] * X[k]; J[2][1] += d3[k] * Y[k]; J[2][2] += d3[k] * Z[k];
    }
    double det = J[0][0] * (J[1][1] * J[2][2] - J[1][2] * J[2][1]) +
                 J[0][1] * (J[1][2] * J[2][0] - J[1][0] * J[2][2]) +
                 J[0][2] * (J[1][0] * J[2][1] - J[1][1] * J[2][0]);
    if (fabs(det) <= 2.2250738585072014e-308 * 100.0) { ok = false; det = 1.0; }
    double I[3][3];
    I[0][0] = (J[1][1] * J[2][2] - J[1][2] * J[2][1]) / det;
    I[0][1] = (J[2][1] * J[0][2] - J[2][2] * J[0][1]) / det;
    I[0][2] = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) / det;
    I[1][0] = (J[2][0] * J[1][2] - J[2][2] * J[1][0]) / det;
    I[1][1] = (J[0][0] * J[2][2] - J[0][2] * J[2][0]) / det;
    I[1][2] = (J[1][0] * J[0][2] - J[1][2] * J[0][0]) / det;
    I[2][0] = (J[1][0] * J[2][1] - J[1][1] * J[2][0]) / det;
    I[2][1] = (J[2][0] * J[0][1] - J[2][1] * J[0][0]) / det;
    I[2][2] = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) / det;
    if (gpt == 0)
      for (int d = 0; d < 3; ++d)
        for (int j = 0; j < 3; ++j) fast[(size_t)i * 10 + 3 * d + j] = I[d][j];
    if (fastJ)
      for (int d = 0; d < 3; ++d)
        for (int j = 0; j < 3; ++j) fastJ[((size_t)i * 10 + gpt) * 9 + 3 * d + j] = I[d][j];
    for (int j = 0; j < 10; ++j) {
      const double bx = I[0][0] * d1[j] + I[0][1] * d2[j] + I[0][2] * d3[j];
      const double by = I[1][0] * d1[j] + I[1][1] * d2[j] + I[1][2] * d3[j];
      const double bz = I[2][0] * d1[j] + I[2][1] * d2[j] + I[2][2] * d3[j];
      // D*B block of node j: rows xx,yy,zz,xy,xz,yz ; columns u,v,w  (itet.f:917-934)
      const double db[6][3] = {{D * bx, D1 * by, D1 * bz},  {D1 * bx, D * by, D1 * bz},
                               {D1 * bx, D1 * by, D * bz},  {D2 * by, D2 * bx, 0.0},
                               {D2 * bz, 0.0, D2 * bx},     {0.0, D2 * bz, D2 * by}};
      const double bd[3] = {bx, by, bz};
      for (int p = 0; p < 10; ++p) {
        const double w = pts->W[p][gpt];
        if (w == 0.0) continue;
        for (int c = 0; c < 6; ++c)
          for (int d = 0; d < 3; ++d) S[frag_index8(p * 6 + c, 3 * j + d, KT)] += w * db[c][d];
        for (int d = 0; d < 3; ++d) G[frag_index8(tet10_grad_row(p, d), j, 3)] += w * bd[d];
      }
    }
  }
  if (!ok) {
    for (int k = 0; k < 8 * KT * 32; ++k) S[k] = 0.0;
    for (int k = 0; k < 12 * 32; ++k) G[k] = 0.0;
  }
  failed[i] = ok ? 0 : 1;
  // straight-sided element: every mid-edge node sits at the midpoint of its two corners (to 1e-13 of the edge length), so
  // the Jacobian is constant and the displacement gradient is linear over the element (k2_tet10_affine_vm_kernel); only for
  // the nodal evaluation (-stressForm 0), where evaluation point 0 is corner 0
  const int mid[6][3] = {{1, 0, 2}, {3, 2, 4}, {5, 4, 0}, {6, 0, 9}, {7, 2, 9}, {8, 4, 9}};
  bool affine = ok && npt == 10;
  for (int m = 0; m < 6 && affine; ++m) {
    const int q = mid[m][0], a = mid[m][1], b = mid[m][2];
    const double ex = X[b] - X[a], ey = Y[b] - Y[a], ez = Z[b] - Z[a];
    const double dx = X[q] - 0.5 * (X[a] + X[b]), dy = Y[q] - 0.5 * (Y[a] + Y[b]), dz = Z[q] - 0.5 * (Z[a] + Z[b]);
    if (dx * dx + dy * dy + dz * dz > 1.0e-26 * (ex * ex + ey * ey + ez * ez)) affine = false;
  }
  // 1 = straight-sided, 0 = curved (scalar kernel with ten inverses), -1 = operator build failed (general kernel: hugeVal)
  fast[(size_t)i * 10 + 9] = affine ? 1.0 : ok ? 0.0 : -1.0;
}

// one warp per element; 8 m-tiles x 8 k-tiles of operator fragments live in registers
__global__ void __launch_bounds__(256, 1)
k2_tet10_vm_kernel(const double* __restrict__ U, size_t ldu, int nsteps, int nsteps_pad,
                   const double* __restrict__ Sfrag, const int* __restrict__ edof,
                   const int* __restrict__ ptoff, const unsigned char* __restrict__ failed, int nelt,
                   double* __restrict__ vm, size_t ld_vm, double* __restrict__ env_max,
                   double* __restrict__ env_min)
{
  constexpr int KT = 8, MT = 8;
  __shared__ __align__(16) double sig_s[8][64 * 8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, t4 = lane & 3;
  const int i = blockIdx.x * (blockDim.x >> 5) + warp;
  if (i >= nelt) return;
  double* sig = sig_s[warp];

  double a[MT][KT];
  const double* sf = Sfrag + (size_t)i * MT * KT * 32 + lane;
#pragma unroll
  for (int m = 0; m < MT; ++m)
#pragma unroll
    for (int j = 0; j < KT; ++j) a[m][j] = __ldg(sf + (size_t)(m * KT + j) * 32);
  const double* urow[KT];
#pragma unroll
  for (int j = 0; j < KT; ++j) urow[j] = U + (size_t)__ldg(edof + (size_t)i * KT * 4 + j * 4 + t4) * ldu + g;

  const bool bad = failed[i] != 0;
  const size_t pt0 = (size_t)ptoff[i];
  // the (node, step-in-tile) pairs this lane evaluates: idx = lane + 32 r -> node = idx/8, step = idx%8
  const int st = lane & 7, pl = lane >> 3;
  double emax[3] = {0.0, 0.0, 0.0}, emin[3] = {kHuge, kHuge, kHuge};

  const int ntiles = nsteps_pad >> 3;
  double b[KT], bn[KT];
#pragma unroll
  for (int j = 0; j < KT; ++j) b[j] = urow[j][0];
  for (int nt = 0; nt < ntiles; ++nt) {
    if (nt + 1 < ntiles) {
#pragma unroll
      for (int j = 0; j < KT; ++j) bn[j] = urow[j][(nt + 1) * 8];
    }
    double c[MT][2];
#pragma unroll
    for (int m = 0; m < MT; ++m) c[m][0] = c[m][1] = 0.0;
#pragma unroll
    for (int j = 0; j < KT; ++j)
#pragma unroll
      for (int m = 0; m < MT; ++m) dmma884(c[m][0], c[m][1], a[m][j], b[j]);
    // transpose through shared memory: sig[row][step]
#pragma unroll
    for (int m = 0; m < MT; ++m)
      *reinterpret_cast<double2*>(sig + (m * 8 + g) * 8 + 2 * t4) = make_double2(c[m][0], c[m][1]);
    __syncwarp();
    const int t = nt * 8 + st;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const int p = pl + 4 * r;
      if (p < 10) {
        const double* s6 = sig + (p * 6) * 8 + st;
        const double s11 = s6[0], s22 = s6[8], s33 = s6[16], s12 = s6[24], s13 = s6[32], s23 = s6[40];
        double v = sqrt(s11 * s11 + s22 * s22 + s33 * s33 - s11 * s22 - s22 * s33 - s33 * s11 +
                        3.0 * (s12 * s12 + s13 * s13 + s23 * s23));
        if (bad) v = kHuge;
        if (t < nsteps) {
          if (vm) vm[(size_t)t * ld_vm + pt0 + p] = v;
          emax[r] = fmax(emax[r], v);
          emin[r] = fmin(emin[r], v);
        }
      }
    }
    __syncwarp();
#pragma unroll
    for (int j = 0; j < KT; ++j) b[j] = bn[j];
  }
#pragma unroll
  for (int r = 0; r < 3; ++r) {
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) {
      emax[r] = fmax(emax[r], __shfl_xor_sync(0xffffffffu, emax[r], o));
      emin[r] = fmin(emin[r], __shfl_xor_sync(0xffffffffu, emin[r], o));
    }
    const int p = pl + 4 * r;
    if (st == 0 && p < 10 && nsteps > 0) {
      if (emax[r] > env_max[pt0 + p]) env_max[pt0 + p] = emax[r];
      if (emin[r] < env_min[pt0 + p]) env_min[pt0 + p] = emin[r];
    }
  }
}


// von Mises of an isotropic solid from the displacement gradient H[c][d] = d u_c / d x_d.  With
// sigma = lambda tr(eps) I + 2 mu eps (the D, D1, D2 of itet.f:917-934: D - D1 = 2 mu, D2 = mu) the hydrostatic
// part drops out of von Mises: vm = 2 mu sqrt( ((exx-eyy)^2 + (eyy-ezz)^2 + (ezz-exx)^2)/2 + 3/4 (gxy^2 + gxz^2 + gyz^2) ),
// the same number as FFaTensorTransforms::vonMises of the stress tensor up to rounding, at a third of the
// FP64 instructions (DMMA and scalar FP64 share one pipe: every instruction here is wall time).
__device__ __forceinline__ double solid_vm2_from_gradient(const double (&H)[3][3])   // (vm / 2 mu)^2 >= 0
{
  const double a = H[0][0] - H[1][1], b = H[1][1] - H[2][2], c = H[2][2] - H[0][0];
  const double gxy = H[0][1] + H[1][0], gxz = H[0][2] + H[2][0], gyz = H[1][2] + H[2][1];
  const double dev = 0.5 * fma(a, a, fma(b, b, c * c));
  const double shr = fma(gxy, gxy, fma(gxz, gxz, gyz * gyz));
  return fma(0.75, shr, dev);
}
__device__ __forceinline__ double solid_vm_from_gradient(const double (&H)[3][3], double mu2)
{
  return mu2 * sqrt_pos(solid_vm2_from_gradient(H));
}
// the same from the six engineering strains (exx, eyy, ezz, gxy, gxz, gyz), which are linear in the nodal displacements
// like the gradient but only six numbers to pass between lanes
__device__ __forceinline__ double solid_vm2_from_strain(const double (&e)[6])
{
  const double a = e[0] - e[1], b = e[1] - e[2], c = e[2] - e[0];
  const double dev = 0.5 * fma(a, a, fma(b, b, c * c));
  const double shr = fma(e[3], e[3], fma(e[4], e[4], e[5] * e[5]));
  return fma(0.75, shr, dev);
}


// one warp per element: 4 m-tiles x 3 k-tiles of the gradient operator in registers, 36 DMMA per 8 steps.
// Two 8-step tiles per loop trip: result points 0..7 are evaluated by their accumulator lanes after each tile,
// result points 8 and 9 of both tiles (2 x 2 x 8 = 32 evaluations) are spread over the 32 lanes once per trip.
template <int MINB>
__global__ void __launch_bounds__(256, MINB)
k2_tet10_grad_vm_kernel(const double* __restrict__ U, size_t ldu, int nsteps, int nsteps_pad,
                        const double* __restrict__ Gfrag, const double* __restrict__ aux, const int* __restrict__ edof,
                        const int* __restrict__ ptoff, const unsigned char* __restrict__ failed, int nelt,
                        double* __restrict__ vm, size_t ld_vm, double* __restrict__ env_max, double* __restrict__ env_min,
                        const int* __restrict__ list /* family elements to process, NULL = all */)
{
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, t4 = lane & 3;
  const int il = blockIdx.x * (blockDim.x >> 5) + warp;
  if (il >= nelt) return;
  const int i = list ? __ldg(list + il) : il;

  double a[4][3];
  const double* gf = Gfrag + (size_t)i * 12 * 32 + lane;
#pragma unroll
  for (int m = 0; m < 4; ++m)
#pragma unroll
    for (int j = 0; j < 3; ++j) a[m][j] = __ldg(gf + (size_t)(m * 3 + j) * 32);
  // B operand: k = element node 4*j + t4 (nodes 10, 11 are padding: operator columns are zero), n = step g
  int urow[3][3];      // row of U per (component, k-tile); addresses are formed at the load (saves 9 registers)
  const int* ed = edof + (size_t)i * 32;
  const double* Ug = U + g;
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const int node = 4 * j + t4;
#pragma unroll
    for (int c = 0; c < 3; ++c) urow[c][j] = node < 10 ? __ldg(ed + 3 * node + c) : 0;
  }
  const double E = __ldg(aux + (size_t)i * 2), nu = __ldg(aux + (size_t)i * 2 + 1);
  const double mu2 = E / (1.0 + nu);            // 2 mu = D - D1
  const bool bad = failed[i] != 0;
  const size_t pt0 = (size_t)ptoff[i];
  // the (tile parity, point 8/9, step) evaluation of this lane in the second phase and its source lanes
  const int par2 = lane >> 4, node2 = (lane >> 3) & 1, step2 = lane & 7;
  const int src0 = ((3 * node2) << 2) | (step2 >> 1);     // lane holding d/dx of that point at that step pair
  const bool odd2 = step2 & 1;
  double emax = 0.0, emin = kHuge, emax2 = 0.0, emin2 = kHuge;

  const int ntiles = nsteps_pad >> 3;
  double b[3][3], bn[3][3];
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int j = 0; j < 3; ++j) b[c][j] = Ug[(size_t)urow[c][j] * ldu];
  for (int nt = 0; nt < ntiles; nt += 2) {
    double x3[2][3][2];   // m-tile 3 accumulators (points 8, 9) of the two tiles of this trip
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int tile = nt + half;
      if (tile + 1 < ntiles) {
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
          for (int j = 0; j < 3; ++j) bn[c][j] = Ug[(size_t)urow[c][j] * ldu + (tile + 1) * 8];
      }
      double acc[4][3][2];
#pragma unroll
      for (int m = 0; m < 4; ++m)
#pragma unroll
        for (int c = 0; c < 3; ++c) acc[m][c][0] = acc[m][c][1] = 0.0;
#pragma unroll
      for (int j = 0; j < 3; ++j)
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
          for (int m = 0; m < 4; ++m) dmma884(acc[m][c][0], acc[m][c][1], a[m][j], b[c][j]);
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int t = tile * 8 + 2 * t4 + q;
        double H[3][3];
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
          for (int d = 0; d < 3; ++d) H[c][d] = acc[d][c][q];
        double v = solid_vm_from_gradient(H, mu2);
        if (bad) v = kHuge;
        if (t < nsteps) {
          if (vm) vm[(size_t)t * ld_vm + pt0 + g] = v;
          emax = fmax(emax, v);
          emin = fmin(emin, v);
        }
      }
#pragma unroll
      for (int c = 0; c < 3; ++c) { x3[half][c][0] = acc[3][c][0]; x3[half][c][1] = acc[3][c][1]; }
#pragma unroll
      for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int j = 0; j < 3; ++j) b[c][j] = bn[c][j];
    }
    // points 8 and 9: rows 24 + 3 (p - 8) + d of the operator = accumulator lanes g = 3 (p - 8) + d
    {
      double H[3][3];
#pragma unroll
      for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int d = 0; d < 3; ++d) {
          const int src = src0 + 4 * d;
          const double a0 = __shfl_sync(0xffffffffu, x3[0][c][0], src), a1 = __shfl_sync(0xffffffffu, x3[0][c][1], src);
          const double b0 = __shfl_sync(0xffffffffu, x3[1][c][0], src), b1 = __shfl_sync(0xffffffffu, x3[1][c][1], src);
          H[c][d] = par2 ? (odd2 ? b1 : b0) : (odd2 ? a1 : a0);
        }
      double v = solid_vm_from_gradient(H, mu2);
      if (bad) v = kHuge;
      const int t = (nt + par2) * 8 + step2;
      if (t < nsteps) {
        if (vm) vm[(size_t)t * ld_vm + pt0 + 8 + node2] = v;
        emax2 = fmax(emax2, v);
        emin2 = fmin(emin2, v);
      }
    }
  }
  // fold the four step-lanes of points 0..7, and the 16 (parity, step) lanes of points 8, 9
#pragma unroll
  for (int o = 1; o < 4; o <<= 1) {
    emax = fmax(emax, __shfl_xor_sync(0xffffffffu, emax, o));
    emin = fmin(emin, __shfl_xor_sync(0xffffffffu, emin, o));
  }
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    if (o == 8) continue;
    emax2 = fmax(emax2, __shfl_xor_sync(0xffffffffu, emax2, o));
    emin2 = fmin(emin2, __shfl_xor_sync(0xffffffffu, emin2, o));
  }
  if (nsteps > 0) {
    if (t4 == 0) {
      if (emax > env_max[pt0 + g]) env_max[pt0 + g] = emax;
      if (emin < env_min[pt0 + g]) env_min[pt0 + g] = emin;
    }
    if ((lane & 0x17) == 0) {
      if (emax2 > env_max[pt0 + 8 + node2]) env_max[pt0 + 8 + node2] = emax2;
      if (emin2 < env_min[pt0 + 8 + node2]) env_min[pt0 + 8 + node2] = emin2;
    }
  }
}

// Straight-sided TET10 (constant Jacobian J, the common case away from curved boundaries): grad u is linear over the
// element, so it is evaluated at the four corners only and averaged for the six mid-edge points -- exactly what the nodal
// evaluation of ITET32 (itet.f:822-934) gives for such an element, up to rounding.  The natural derivatives at a corner are
// three-point differences along the edges (rows of DN1031, itet.f:48-84, at L_a = 1):
//     du/dL_j |corner a  =  g_j - g_4,   g_a = 3 u_a,   g_b = 4 u_mid(a,b) - u_b  (b != a)
// and grad u = J^-1 . D with ONE 3x3 inverse per element.  No operator is streamed (80 bytes of constants per element instead
// of 3 KB of fragments) and no tensor-core work is padded: ~500 FP64 operations per element.step instead of ~1,400.
// One warp per element, lane = (corner k = lane / 8, step s = lane % 8): the lane evaluates corner k at step s, then the
// mid-edge points in two rounds, fetching the partner corner's gradient with warp shuffles.
// The element's 30 displacement rows of a tile of 8 steps (30 x 64 bytes) are staged through shared memory with cp.async,
// the tile after the current one always in flight (two buffers per warp): the lanes read their 24 operands from shared
// memory, so no load latency sits inside the arithmetic and only 8 index registers are needed.
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc)
{
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

constexpr int kAffRows = 36;   // 12 node slots x 3 components, 64 bytes each

// WRITE_VM = false (envelope only): the envelope is taken over the radicand (vm / 2 mu)^2, which orders like vm, and the
// square root is taken once per result point at the end instead of once per step.
// CURVED = true: the same kernel for curved elements (nodal evaluation).  u is quadratic in the volume coordinates, so its
// NATURAL derivatives D = du/dL are linear in them whatever the geometry: the three-point differences at the corners stay
// exact and D at a mid-edge point is still the average of its two corners; only J^-1 differs from point to point
// (fastJ: the ten inverses of ITET32's nodal evaluation, itet.f:822-934).  The lanes exchange D (nine numbers) instead of
// the strains and keep the inverses of their three result points in registers, hence the smaller blocks.
template <bool WRITE_VM, bool CURVED, int NW>
__global__ void __launch_bounds__(NW * 32, CURVED ? 3 : 2)
k2_tet10_affine_vm_kernel(const double* __restrict__ U, size_t ldu, int nsteps, int nsteps_pad, const double* __restrict__ fast,
                          const double* __restrict__ aux, const int* __restrict__ edof, const int* __restrict__ ptoff, int nlist,
                          const int* __restrict__ list, double* __restrict__ vm, size_t ld_vm, double* __restrict__ env_max,
                          double* __restrict__ env_min, const double* __restrict__ fastJ = nullptr)
{
  __shared__ __align__(16) double sU_all[NW][2][kAffRows * 8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int il = blockIdx.x * NW + warp;
  if (il >= nlist) return;   // whole warp
  const int i = __ldg(list + il);
  const int k = lane >> 3, s = lane & 7;
  double (*sU)[kAffRows * 8] = sU_all[warp];
  // element nodes (0-based) of the corners L1..L4 and of the mid-edge node between two corners; mid(a,a) = corner a makes
  // g_a = 4 u_a - u_a = 3 u_a fall out of the same expression.  Shared-memory slot of a node: the two nodes that lanes of
  // one half-warp (corners k, k+1) read with the same instruction sit in slots of different parity = different bank halves.
  const int cn[4] = {0, 2, 4, 9};
  const int mdn[4][4] = {{0, 1, 5, 6}, {1, 2, 3, 7}, {5, 3, 4, 8}, {6, 7, 8, 9}};
  const int slot[10] = {0, 1, 2, 3, 8, 4, 5, 6, 7, 10};   // even: nodes 0 2 5 7 4 9, odd: nodes 1 3 6 8
  const int* ed = edof + (size_t)i * 32;
  // staging: chunk q = lane + 32 r (r < 4, q < 120) is 16 bytes: 2 steps of row q / 4 (node q / 12, component (q / 4) % 3)
  const double* csrc[4];
  int cdst[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int q = lane + 32 * r, row = q >> 2;
    const int node = row / 3, comp = row - 3 * node;
    csrc[r] = q < 120 ? U + (size_t)__ldg(ed + row) * ldu + 2 * (q & 3) : nullptr;
    cdst[r] = q < 120 ? (slot[node] * 3 + comp) * 8 + 2 * (q & 3) : 0;
  }
  int ms[4];   // shared-memory offsets (doubles) of mid(k, ci), component 0, step s
#pragma unroll
  for (int ci = 0; ci < 4; ++ci) ms[ci] = slot[mdn[k][ci]] * 24 + s;
  const double E = __ldg(aux + (size_t)i * 2), nu = __ldg(aux + (size_t)i * 2 + 1);
  const double mu2 = E / (1.0 + nu), mu1 = 0.5 * mu2;   // mid-edge points: vm(1/2 (Ha + Hb)) = 1/2 vm(Ha + Hb)
  const size_t pt0 = (size_t)ptoff[i];
  // result points of this lane: corner k; round 1 the mid node of edge (k, p1[k]); round 2 (k = 1, 2) of edge (k, 3)
  const int p1 = k == 0 ? 1 : k == 1 ? 2 : 0;
  const int pc = cn[k], pm1 = mdn[k][p1], pm2 = mdn[k][3];
  const bool r2 = k == 1 || k == 2;
  const int src1 = p1 * 8 + s, src2 = 24 + s;
  double Ji[3][3], Jm1[3][3], Jm2[3][3];   // J^-1 (row d, column j): of the element, or of this lane's three result points
#pragma unroll
  for (int d = 0; d < 3; ++d)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      if (CURVED) {
        Ji[d][j] = __ldg(fastJ + ((size_t)i * 10 + pc) * 9 + 3 * d + j);
        Jm1[d][j] = __ldg(fastJ + ((size_t)i * 10 + pm1) * 9 + 3 * d + j);
        Jm2[d][j] = __ldg(fastJ + ((size_t)i * 10 + pm2) * 9 + 3 * d + j);
      } else
        Ji[d][j] = __ldg(fast + (size_t)i * 10 + 3 * d + j);
    }
  double emax[3] = {0.0, 0.0, 0.0}, emin[3] = {kHuge, kHuge, kHuge};

  auto stage = [&](int buf, int t0) {
#pragma unroll
    for (int r = 0; r < 4; ++r)
      if (r < 3 || lane < 24) cp_async16(&sU[buf][cdst[r]], csrc[r] + t0);
    cp_async_commit();
  };
  stage(0, 0);
  int buf = 0;
  for (int t0 = 0; t0 < nsteps_pad; t0 += 8, buf ^= 1) {
    if (t0 + 8 < nsteps_pad) { stage(buf ^ 1, t0 + 8); cp_async_wait<1>(); }
    else cp_async_wait<0>();
    __syncwarp();
    const double* su = sU[buf];
    double H[3][3], D[3][3];   // D[c][j] = d u_c / d L_j at corner k
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      double gq[4];
#pragma unroll
      for (int ci = 0; ci < 4; ++ci) gq[ci] = fma(4.0, su[ms[ci] + c * 8], -su[(slot[cn[ci]] * 3 + c) * 8 + s]);
      D[c][0] = gq[0] - gq[3]; D[c][1] = gq[1] - gq[3]; D[c][2] = gq[2] - gq[3];
#pragma unroll
      for (int d = 0; d < 3; ++d) H[c][d] = fma(Ji[d][2], D[c][2], fma(Ji[d][1], D[c][1], Ji[d][0] * D[c][0]));
    }
    const int t = t0 + s;
    const bool live = t < nsteps;
    // only the symmetric part of the gradient enters von Mises: six strains per corner go between the lanes, not nine
    const double e[6] = {H[0][0], H[1][1], H[2][2], H[0][1] + H[1][0], H[0][2] + H[2][0], H[1][2] + H[2][1]};
    double rad = solid_vm2_from_strain(e);
    double v = WRITE_VM ? mu2 * sqrt_pos(rad) : rad;
    if (live) {
      if (WRITE_VM) vm[(size_t)t * ld_vm + pt0 + pc] = v;
      emax[0] = max_nonneg(emax[0], v); emin[0] = min_nonneg(emin[0], v);
    }
    double em[6];
    // strains of a mid-edge point: straight-sided = sum of the two corners' strains (the 1/2 sits in mu1); curved = its own
    // J^-1 applied to the sum of the two corners' natural derivatives
    auto mid_strains = [&](int src, const double (&Jm)[3][3]) {
      if (!CURVED) {
#pragma unroll
        for (int c = 0; c < 6; ++c) em[c] = e[c] + __shfl_sync(0xffffffffu, e[c], src);
      } else {
        double Hm[3][3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const double d0 = D[c][0] + __shfl_sync(0xffffffffu, D[c][0], src), d1 = D[c][1] + __shfl_sync(0xffffffffu, D[c][1], src),
                       d2 = D[c][2] + __shfl_sync(0xffffffffu, D[c][2], src);
#pragma unroll
          for (int d = 0; d < 3; ++d) Hm[c][d] = fma(Jm[d][2], d2, fma(Jm[d][1], d1, Jm[d][0] * d0));
        }
        em[0] = Hm[0][0]; em[1] = Hm[1][1]; em[2] = Hm[2][2];
        em[3] = Hm[0][1] + Hm[1][0]; em[4] = Hm[0][2] + Hm[2][0]; em[5] = Hm[1][2] + Hm[2][1];
      }
    };
    mid_strains(src1, Jm1);
    rad = solid_vm2_from_strain(em);
    v = WRITE_VM ? mu1 * sqrt_pos(rad) : rad;
    if (live) {
      if (WRITE_VM) vm[(size_t)t * ld_vm + pt0 + pm1] = v;
      emax[1] = max_nonneg(emax[1], v); emin[1] = min_nonneg(emin[1], v);
    }
    mid_strains(src2, Jm2);
    if (r2) {
      rad = solid_vm2_from_strain(em);
      v = WRITE_VM ? mu1 * sqrt_pos(rad) : rad;
      if (live) {
        if (WRITE_VM) vm[(size_t)t * ld_vm + pt0 + pm2] = v;
        emax[2] = max_nonneg(emax[2], v); emin[2] = min_nonneg(emin[2], v);
      }
    }
    __syncwarp();   // everybody is done with this buffer before the next iteration refills it
  }
  // fold the eight step lanes of each point, then into the stored envelope
#pragma unroll
  for (int r = 0; r < 3; ++r) {
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) {
      emax[r] = max_nonneg(emax[r], __shfl_xor_sync(0xffffffffu, emax[r], o));
      emin[r] = min_nonneg(emin[r], __shfl_xor_sync(0xffffffffu, emin[r], o));
    }
    if (!WRITE_VM) {   // radicand -> von Mises; kHuge = no step seen
      const double m = r == 0 ? mu2 : mu1;
      emax[r] = m * sqrt_pos(emax[r]);
      emin[r] = emin[r] == kHuge ? kHuge : m * sqrt_pos(emin[r]);
    }
  }
  if (s == 0 && nsteps > 0) {
    const int pr[3] = {pc, pm1, pm2};
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      if (r == 2 && !r2) continue;
      if (emax[r] > env_max[pt0 + pr[r]]) env_max[pt0 + pr[r]] = emax[r];
      if (emin[r] < env_min[pt0 + pr[r]]) env_min[pt0 + pr[r]] = emin[r];
    }
  }
}

// The same scalar formulation with lane = TIME STEP (tiles of 32 steps): every lane holds the 30 nodal displacements of its
// own step in registers, so a row of U is read as ONE 256-byte segment per warp, nothing goes through shared memory or
// shuffles on the way to the strains (the (corner, step) kernel above spends 24 LDS + 18 SHFL per lane and tile and is
// LSU-bound), and the envelope of the ten result points lives in registers.  When the history is written, the 10 x 32
// values of a tile are turned through shared memory so that each step record receives its ten points as one 80-byte piece.
// Used for tiles of at least 32 steps; shorter tiles stay on the (corner, step) kernel.
template <bool WRITE_VM, bool CURVED, int NW>
__global__ void __launch_bounds__(NW * 32, CURVED ? 3 : 4)
k2_tet10_steplane_vm_kernel(const double* __restrict__ U, size_t ldu, int nsteps, const double* __restrict__ fast,
                            const double* __restrict__ aux, const int* __restrict__ edof, const int* __restrict__ ptoff, int nlist,
                            const int* __restrict__ list, double* __restrict__ vm, size_t ld_vm, double* __restrict__ env_max,
                            double* __restrict__ env_min, const double* __restrict__ fastJ)
{
  __shared__ double sv_all[NW][10 * 33];
  __shared__ unsigned so_all[CURVED ? 1 : NW][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int il = blockIdx.x * NW + warp;
  if (il >= nlist) return;   // whole warp
  const int i = __ldg(list + il);
  double* sv = sv_all[warp];
  constexpr int cn[4] = {0, 2, 4, 9};
  constexpr int mdn[4][4] = {{0, 1, 5, 6}, {1, 2, 3, 7}, {5, 3, 4, 8}, {6, 7, 8, 9}};
  constexpr int edge[6][3] = {{0, 1, 1}, {1, 2, 3}, {0, 2, 5}, {0, 3, 6}, {1, 3, 7}, {2, 3, 8}};   // corners a, b -> node
  const int* ed = edof + (size_t)i * 32;
  // row offsets in units of 64 doubles (ldu is a multiple of 64): 32 bits are enough for any U, one LEA pair per load
  const unsigned ldu64 = (unsigned)(ldu >> 6);
  // (kept in shared memory by the straight-sided variant: one broadcast LDS + one IMAD.WIDE per load; the curved one
  // re-reads the row numbers)
  unsigned* eoff = so_all[CURVED ? 0 : warp];
  if (!CURVED) {
    if (lane < 30) eoff[lane] = (unsigned)__ldg(ed + lane) * ldu64;
    __syncwarp();
  }
  const double E = __ldg(aux + (size_t)i * 2), nu = __ldg(aux + (size_t)i * 2 + 1);
  const double mu2 = E / (1.0 + nu), mu1 = 0.5 * mu2;
  const size_t pt0 = (size_t)ptoff[i];
  double Ji[3][3];
  if (!CURVED) {
#pragma unroll
    for (int d = 0; d < 3; ++d)
#pragma unroll
      for (int j = 0; j < 3; ++j) Ji[d][j] = __ldg(fast + (size_t)i * 10 + 3 * d + j);
  }
  const double* fJ = CURVED ? fastJ + (size_t)i * 90 : nullptr;
  double emax = 0.0, emin = kHuge;   // lanes 0..29: result point lane % 10, steps 11 (lane / 10) .. + 10 of every tile

  for (int t0 = 0; t0 < nsteps; t0 += 32) {
    // lanes past the last step repeat it: no predicates, no zero fill, and a repeated value does not move an envelope
    const double* Ut = U + min(t0 + lane, nsteps - 1);
    double u[10][3];
    if (!CURVED) {
#pragma unroll
      for (int n = 0; n < 10; ++n)
#pragma unroll
        for (int c = 0; c < 3; ++c) u[n][c] = __ldg(Ut + ((size_t)eoff[3 * n + c] << 6));
    } else {   // predicated loads keep ptxas from front-loading the 90 inverses on top of these (it spills 1 KB otherwise)
      const bool live = t0 + lane < nsteps;
#pragma unroll
      for (int n = 0; n < 10; ++n)
#pragma unroll
        for (int c = 0; c < 3; ++c) u[n][c] = live ? __ldg(Ut + ((size_t)((unsigned)__ldg(ed + 3 * n + c) * ldu64) << 6)) : 0.0;
    }
    double v[10];
    // D[a][c][j] = d u_c / d L_j at corner a (three-point differences, see above)
    double D[4][3][3];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        double gq[4];
#pragma unroll
        for (int ci = 0; ci < 4; ++ci) gq[ci] = fma(4.0, u[mdn[a][ci]][c], -u[cn[ci]][c]);
        D[a][c][0] = gq[0] - gq[3]; D[a][c][1] = gq[1] - gq[3]; D[a][c][2] = gq[2] - gq[3];
      }
    auto strains = [&](const double (&Dp)[3][3], const double (&J)[3][3], double (&e)[6]) {
      double H[3][3];
#pragma unroll
      for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int d = 0; d < 3; ++d) H[c][d] = fma(J[d][2], Dp[c][2], fma(J[d][1], Dp[c][1], J[d][0] * Dp[c][0]));
      e[0] = H[0][0]; e[1] = H[1][1]; e[2] = H[2][2];
      e[3] = H[0][1] + H[1][0]; e[4] = H[0][2] + H[2][0]; e[5] = H[1][2] + H[2][1];
    };
    auto loadJ = [&](int node, double (&J)[3][3]) {
#pragma unroll
      for (int d = 0; d < 3; ++d)
#pragma unroll
        for (int j = 0; j < 3; ++j) J[d][j] = __ldg(fJ + node * 9 + 3 * d + j);
    };
    if (!CURVED) {
      double e[4][6];
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        strains(D[a], Ji, e[a]);
        v[cn[a]] = solid_vm2_from_strain(e[a]);
      }
#pragma unroll
      for (int q = 0; q < 6; ++q) {
        double em[6];
#pragma unroll
        for (int c = 0; c < 6; ++c) em[c] = e[edge[q][0]][c] + e[edge[q][1]][c];
        v[edge[q][2]] = solid_vm2_from_strain(em);
      }
    } else {
      double J[3][3], e[6];
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        loadJ(cn[a], J);
        strains(D[a], J, e);
        v[cn[a]] = solid_vm2_from_strain(e);
      }
#pragma unroll
      for (int q = 0; q < 6; ++q) {
        double Dm[3][3];
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
          for (int j = 0; j < 3; ++j) Dm[c][j] = D[edge[q][0]][c][j] + D[edge[q][1]][c][j];
        loadJ(edge[q][2], J);
        strains(Dm, J, e);
        v[edge[q][2]] = solid_vm2_from_strain(e);
      }
    }
    // corners carry 2 mu, mid-edge points (sums of two corners) mu.  The 10 x 32 values of the tile are turned through shared
    // memory: three lanes per result point fold its envelope (11 steps each), the history goes out as 80-byte pieces of the
    // step records
#pragma unroll
    for (int p = 0; p < 10; ++p) {
      const bool corner = p == 0 || p == 2 || p == 4 || p == 9;
      sv[p * 33 + lane] = WRITE_VM ? (corner ? mu2 : mu1) * sqrt_pos(v[p]) : v[p];
    }
    __syncwarp();
    const int ns1 = min(32, nsteps - t0) - 1;   // last live step of the tile: the scan repeats it instead of reading past it
    if (lane < 30) {
      const int pl = lane % 10, s0 = 11 * (lane / 10);
      const double* row = sv + pl * 33;
#pragma unroll
      for (int q = 0; q < 11; ++q) {
        const double x = row[min(s0 + q, ns1)];
        emax = max_nonneg(emax, x); emin = min_nonneg(emin, x);
      }
    }
    if (WRITE_VM) {
#pragma unroll
      for (int r = 0; r < 10; ++r) {
        const int idx = lane + 32 * r, s = idx / 10, p = idx - 10 * s;
        if (s <= ns1) vm[(size_t)(t0 + s) * ld_vm + pt0 + p] = sv[p * 33 + s];
      }
    }
    __syncwarp();
  }
  emax = max_nonneg(emax, max_nonneg(__shfl_down_sync(0xffffffffu, emax, 10), __shfl_down_sync(0xffffffffu, emax, 20)));
  emin = min_nonneg(emin, min_nonneg(__shfl_down_sync(0xffffffffu, emin, 10), __shfl_down_sync(0xffffffffu, emin, 20)));
  if (lane < 10 && nsteps > 0) {
    if (!WRITE_VM) {   // radicand -> von Mises
      const double m = (lane == 0 || lane == 2 || lane == 4 || lane == 9) ? mu2 : mu1;
      emax = m * sqrt_pos(emax);
      emin = m * sqrt_pos(emin);
    }
    if (emax > env_max[pt0 + lane]) env_max[pt0 + lane] = emax;
    if (emin < env_min[pt0 + lane]) env_min[pt0 + lane] = emin;
  }
}

int build_solid_operators(fsr_part* p, const fsr_sam* sam, const fsr_elmdata* elm)
{
  cudaStream_t s = p->stream;
  FamilyData& f = p->fam[FAM_TET10];
  f.nenod = 10; f.nndof = 3; f.nstrp = 10; f.ncmp = 6; f.MT = 8; f.KT = 8;
  std::vector<int> elem, conn, edof, ptoff;
  for (int e : elements_of_type(p, sam, elm, 41)) {
    int ip0 = sam->mpmnpc[e] - 1, nn = sam->mpmnpc[e + 1] - sam->mpmnpc[e];
    if (nn != 10) { set_error("TET10 element %d has %d nodes", e + 1, nn); return FSR_ERR_ARG; }
    elem.push_back(e);
    ptoff.push_back(p->ptoff_host[e]);
    size_t base = edof.size();
    edof.resize(base + 32, 0);
    for (int k = 0; k < 10; ++k) {
      int n = sam->mmnpc[ip0 + k] - 1;
      if (n < 0 || n >= sam->nnod) { set_error("element %d: node index out of range", e + 1); return FSR_ERR_ARG; }
      conn.push_back(n);
      int js = sam->madof[n] - 1, nd = sam->madof[n + 1] - sam->madof[n];
      if (nd < 3) { set_error("element %d: node %d has %d DOFs, solid needs 3", e + 1, n + 1, nd); return FSR_ERR_ARG; }
      for (int d = 0; d < 3; ++d) edof[base + (size_t)k * 3 + d] = js + d;
    }
  }
  f.nelt = (int)elem.size();
  f.naux = 2;
  if (f.nelt == 0) return FSR_OK;

  // evaluation points and node extrapolation weights (itet.f:822-877, elStressModule.f90:1357-1368)
  Tet10Points h;
  memset(&h, 0, sizeof(h));
  if (p->stressForm == 0) {
    h.npt = 10;
    const double L[10][3] = {{1, 0, 0}, {.5, .5, 0}, {0, 1, 0}, {0, .5, .5}, {0, 0, 1},
                             {.5, 0, .5}, {.5, 0, 0}, {0, .5, 0}, {0, 0, .5}, {0, 0, 0}};
    memcpy(h.L, L, sizeof(L));
    for (int q = 0; q < 10; ++q) h.W[q][q] = 1.0;
  } else {
    h.npt = 4;
    const double al = (double)0.585410196625f, be = (double)0.138196601125f;  // REAL*4 literals
    const double L[4][3] = {{al, be, be}, {be, al, be}, {be, be, al}, {be, be, be}};
    memcpy(h.L, L, sizeof(L));
    const double ap = 1.927051062810166, bp = -0.309017015969668;
    const int corner[4] = {0, 2, 4, 9};
    for (int c = 0; c < 4; ++c)
      for (int gq = 0; gq < 4; ++gq) h.W[corner[c]][gq] = (c == gq) ? ap : bp;
    const int mid[6][3] = {{1, 0, 2}, {3, 2, 4}, {5, 4, 0}, {6, 0, 9}, {7, 2, 9}, {8, 4, 9}};
    for (auto& m : mid)
      for (int gq = 0; gq < 4; ++gq) h.W[m[0]][gq] = 0.5 * (h.W[m[1]][gq] + h.W[m[2]][gq]);
  }
  Tet10Points* d_pts = nullptr;
  int* d_conn = nullptr;
  FSR_CUDA(cudaMalloc(&d_pts, sizeof(Tet10Points)));
  FSR_CUDA(cudaMemcpyAsync(d_pts, &h, sizeof(h), cudaMemcpyHostToDevice, s));
  FSR_CUDA(cudaMalloc(&f.elem, sizeof(int) * elem.size()));
  FSR_CUDA(cudaMalloc(&f.edof, sizeof(int) * edof.size()));
  FSR_CUDA(cudaMalloc(&f.ptoff, sizeof(int) * ptoff.size()));
  FSR_CUDA(cudaMalloc(&f.failed, f.nelt));
  FSR_CUDA(cudaMalloc(&f.Sfrag, sizeof(double) * (size_t)f.nelt * f.MT * f.KT * 32));
  FSR_CUDA(cudaMalloc(&f.aux, sizeof(double) * (size_t)f.nelt * f.naux));
  FSR_CUDA(cudaMalloc(&f.Gfrag, sizeof(double) * (size_t)f.nelt * 12 * 32));
  FSR_CUDA(cudaMemsetAsync(f.Gfrag, 0, sizeof(double) * (size_t)f.nelt * 12 * 32, s));
  FSR_CUDA(cudaMalloc(&f.fast, sizeof(double) * (size_t)f.nelt * 10));
  FSR_CUDA(cudaMemsetAsync(f.fast, 0, sizeof(double) * (size_t)f.nelt * 10, s));
  if (p->stressForm == 0) {   // nodal evaluation: the ten inverses of every element, for the curved ones
    FSR_CUDA(cudaMalloc(&f.fast2, sizeof(double) * (size_t)f.nelt * 90));
    FSR_CUDA(cudaMemsetAsync(f.fast2, 0, sizeof(double) * (size_t)f.nelt * 90, s));
  }
  FSR_CUDA(cudaMalloc(&d_conn, sizeof(int) * conn.size()));
  FSR_CUDA(cudaMemcpyAsync(f.elem, elem.data(), sizeof(int) * elem.size(), cudaMemcpyHostToDevice, s));
  FSR_CUDA(cudaMemcpyAsync(f.edof, edof.data(), sizeof(int) * edof.size(), cudaMemcpyHostToDevice, s));
  FSR_CUDA(cudaMemcpyAsync(f.ptoff, ptoff.data(), sizeof(int) * ptoff.size(), cudaMemcpyHostToDevice, s));
  FSR_CUDA(cudaMemcpyAsync(d_conn, conn.data(), sizeof(int) * conn.size(), cudaMemcpyHostToDevice, s));
  FSR_CUDA(cudaMemsetAsync(f.Sfrag, 0, sizeof(double) * (size_t)f.nelt * f.MT * f.KT * 32, s));
  build_tet10_ops_kernel<<<(f.nelt + 63) / 64, 64, 0, s>>>(f.nelt, f.elem, d_conn, p->xyz, p->emod, p->rny,
                                                         d_pts, f.Sfrag, f.failed, f.aux, f.Gfrag, f.fast, f.fast2);
  FSR_LAUNCH_CHECK();
  FSR_CUDA(cudaStreamSynchronize(s));
  cudaFree(d_conn);
  cudaFree(d_pts);
  // split the family into straight-sided elements (fast path) and the rest; FSR_TET10_AFFINE=0 sends all to the general kernel
  {
    std::vector<double> h((size_t)f.nelt * 10);
    FSR_CUDA(cudaMemcpy(h.data(), f.fast, sizeof(double) * h.size(), cudaMemcpyDeviceToHost));
    const bool use = !(getenv("FSR_TET10_AFFINE") && atoi(getenv("FSR_TET10_AFFINE")) == 0);
    // FSR_TET10_CURVED=0 sends the curved ones to the DMMA gradient kernel instead of the scalar one (A/B, cross-check)
    const bool use_curved = use && f.fast2 && !(getenv("FSR_TET10_CURVED") && atoi(getenv("FSR_TET10_CURVED")) == 0);
    std::vector<int> lst[3];
    for (int i = 0; i < f.nelt; ++i) {
      const double fl = h[(size_t)i * 10 + 9];
      lst[use && fl > 0.0 ? 0 : (use_curved && fl == 0.0) ? 2 : 1].push_back(i);
    }
    for (int k = 0; k < 3; ++k) {
      f.nsub[k] = (int)lst[k].size();
      if (f.nsub[k] == 0) continue;
      FSR_CUDA(cudaMalloc(&f.sub[k], sizeof(int) * lst[k].size()));
      FSR_CUDA(cudaMemcpy(f.sub[k], lst[k].data(), sizeof(int) * lst[k].size(), cudaMemcpyHostToDevice));
    }
  }
  return FSR_OK;
}

int launch_k2_tet10_vm(fsr_part* p, int nsteps, int nsteps_pad, double* vm_dev, size_t ld_vm, cudaStream_t s)
{
  FamilyData& f = p->fam[FAM_TET10];
  if (f.nelt == 0) return FSR_OK;
  const int warps = 8;
  // FSR_TET10_DENSE=1 selects the dense 60x30 formulation (kept for A/B timing and as a cross-check)
  static const bool dense = getenv("FSR_TET10_DENSE") && atoi(getenv("FSR_TET10_DENSE")) != 0;
  if (dense)
    k2_tet10_vm_kernel<<<(f.nelt + warps - 1) / warps, warps * 32, 0, s>>>(
        p->U, (size_t)p->step_tile, nsteps, nsteps_pad, f.Sfrag, f.edof, f.ptoff, f.failed, f.nelt, vm_dev,
        ld_vm, p->env_max, p->env_min);
  else {
    // tiles of 32 steps and more: lane = step (FSR_TET10_STEPLANE=0 keeps the (corner, step) kernels, A/B and cross-check)
    const bool steplane = !(getenv("FSR_TET10_STEPLANE") && atoi(getenv("FSR_TET10_STEPLANE")) == 0);
    const bool sl = steplane && nsteps >= 32;
    if (sl && f.nsub[0] > 0) {
      if (vm_dev)
        k2_tet10_steplane_vm_kernel<true, false, 4><<<(f.nsub[0] + 3) / 4, 128, 0, s>>>(
            p->U, (size_t)p->step_tile, nsteps, f.fast, f.aux, f.edof, f.ptoff, f.nsub[0], f.sub[0], vm_dev, ld_vm, p->env_max, p->env_min, nullptr);
      else
        k2_tet10_steplane_vm_kernel<false, false, 4><<<(f.nsub[0] + 3) / 4, 128, 0, s>>>(
            p->U, (size_t)p->step_tile, nsteps, f.fast, f.aux, f.edof, f.ptoff, f.nsub[0], f.sub[0], vm_dev, ld_vm, p->env_max, p->env_min, nullptr);
      FSR_LAUNCH_CHECK();
    }
    if (sl && f.nsub[2] > 0) {
      if (vm_dev)
        k2_tet10_steplane_vm_kernel<true, true, 4><<<(f.nsub[2] + 3) / 4, 128, 0, s>>>(
            p->U, (size_t)p->step_tile, nsteps, f.fast, f.aux, f.edof, f.ptoff, f.nsub[2], f.sub[2], vm_dev, ld_vm, p->env_max, p->env_min, f.fast2);
      else
        k2_tet10_steplane_vm_kernel<false, true, 4><<<(f.nsub[2] + 3) / 4, 128, 0, s>>>(
            p->U, (size_t)p->step_tile, nsteps, f.fast, f.aux, f.edof, f.ptoff, f.nsub[2], f.sub[2], vm_dev, ld_vm, p->env_max, p->env_min, f.fast2);
      FSR_LAUNCH_CHECK();
    }
    if (!sl && f.nsub[0] > 0) {
      if (vm_dev)
        k2_tet10_affine_vm_kernel<true, false, 8><<<(f.nsub[0] + 7) / 8, 256, 0, s>>>(
            p->U, (size_t)p->step_tile, nsteps, nsteps_pad, f.fast, f.aux, f.edof, f.ptoff, f.nsub[0], f.sub[0], vm_dev, ld_vm, p->env_max,
            p->env_min);
      else
        k2_tet10_affine_vm_kernel<false, false, 8><<<(f.nsub[0] + 7) / 8, 256, 0, s>>>(
            p->U, (size_t)p->step_tile, nsteps, nsteps_pad, f.fast, f.aux, f.edof, f.ptoff, f.nsub[0], f.sub[0], vm_dev, ld_vm, p->env_max,
            p->env_min);
      FSR_LAUNCH_CHECK();
    }
    if (!sl && f.nsub[2] > 0) {   // curved elements, nodal evaluation: the scalar kernel with the ten inverses
      if (vm_dev)
        k2_tet10_affine_vm_kernel<true, true, 4><<<(f.nsub[2] + 3) / 4, 128, 0, s>>>(
            p->U, (size_t)p->step_tile, nsteps, nsteps_pad, f.fast, f.aux, f.edof, f.ptoff, f.nsub[2], f.sub[2], vm_dev, ld_vm, p->env_max,
            p->env_min, f.fast2);
      else
        k2_tet10_affine_vm_kernel<false, true, 4><<<(f.nsub[2] + 3) / 4, 128, 0, s>>>(
            p->U, (size_t)p->step_tile, nsteps, nsteps_pad, f.fast, f.aux, f.edof, f.ptoff, f.nsub[2], f.sub[2], vm_dev, ld_vm, p->env_max,
            p->env_min, f.fast2);
      FSR_LAUNCH_CHECK();
    }
    if (f.nsub[1] > 0)
      k2_tet10_grad_vm_kernel<2><<<(f.nsub[1] + warps - 1) / warps, warps * 32, 0, s>>>(
          p->U, (size_t)p->step_tile, nsteps, nsteps_pad, f.Gfrag, f.aux, f.edof, f.ptoff, f.failed, f.nsub[1], vm_dev,
          ld_vm, p->env_max, p->env_min, f.sub[1]);
    else
      return FSR_OK;
  }
  FSR_LAUNCH_CHECK();
  return FSR_OK;
}

}  // namespace fsr
