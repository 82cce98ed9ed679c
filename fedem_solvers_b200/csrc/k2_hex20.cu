// k2_hex20.cu -- K2 for the 20-node hexahedron (type 43) on sm_100a, and the generic "large solid"
// apply kernel (operator too big for registers: 120 x 60 for HEX20).
//
// Reference: STR43 -> IHEX32 -> DN2031 / JACI31 (src/vpmStress/elStressModule.f90:1472-1584,
// src/Femlib/ihex.f:224-560,2433-2545, src/Femlib/jaci31.f) evaluates 27 (or 8) Jacobian inverses
// per element per step.  Here the stress operator sigma(6,20) = S_e . v(3,20) is built once:
// -stressForm 0 = direct evaluation at the 20 nodes; otherwise the 2x2x2 Gauss points (abscissa in
// the reference's REAL*4 precision, ihex.f:364-365) extrapolated tri-linearly with
// (1 +- sqrt3 xi_n)(1 +- sqrt3 eta_n)(1 +- sqrt3 zeta_n)/8 (elStressModule.f90:1563-1577).
// Apply: one CTA (4 warps) per element; the operator lives in shared memory in DMMA A-fragment
// order (57.6 KB), the element's 60 rows of U are staged per 8-step tile, each warp owns a share
// of the 15 m-tiles, accumulators are transposed through shared memory so that one thread sees the
// six components of a node for von Mises (FFaTensorTransforms.C:38-43) and the fused envelope.
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "common.cuh"

namespace fsr {

__device__ __forceinline__ size_t frag_index_g(int row, int col, int KT)
{
  return ((size_t)((row >> 3) * KT + (col >> 2)) << 5) + ((row & 7) << 2) + (col & 3);
}

__constant__ double c_hx[20] = {-1, 0, 1, 1, 1, 0, -1, -1, -1, 1, 1, -1, -1, 0, 1, 1, 1, 0, -1, -1};
__constant__ double c_he[20] = {-1, -1, -1, 0, 1, 1, 1, 0, -1, -1, 1, 1, -1, -1, -1, 0, 1, 1, 1, 0};
__constant__ double c_hz[20] = {-1, -1, -1, -1, -1, -1, -1, -1, 0, 0, 0, 0, 1, 1, 1, 1, 1, 1, 1, 1};

// serendipity shape-function derivatives w.r.t. (xi, eta, zeta), node order of DN2031
__device__ void hex20_dn(double xi, double et, double ze, double* dx, double* de, double* dz)
{
  for (int i = 0; i < 20; ++i) {
    const double a = c_hx[i], b = c_he[i], c = c_hz[i];
    if (a != 0.0 && b != 0.0 && c != 0.0) {           // corner
      dx[i] = .125 * a * (1. + et * b) * (1. + ze * c) * (2. * xi * a + et * b + ze * c - 1.);
      de[i] = .125 * b * (1. + xi * a) * (1. + ze * c) * (xi * a + 2. * et * b + ze * c - 1.);
      dz[i] = .125 * c * (1. + xi * a) * (1. + et * b) * (xi * a + et * b + 2. * ze * c - 1.);
    } else if (a == 0.0) {                             // mid-edge along xi
      dx[i] = -.5 * xi * (1. + et * b) * (1. + ze * c);
      de[i] = .25 * b * (1. - xi * xi) * (1. + ze * c);
      dz[i] = .25 * c * (1. - xi * xi) * (1. + et * b);
    } else if (b == 0.0) {                             // mid-edge along eta
      dx[i] = .25 * a * (1. - et * et) * (1. + ze * c);
      de[i] = -.5 * et * (1. + xi * a) * (1. + ze * c);
      dz[i] = .25 * c * (1. + xi * a) * (1. - et * et);
    } else {                                           // mid-edge along zeta
      dx[i] = .25 * a * (1. + et * b) * (1. - ze * ze);
      de[i] = .25 * b * (1. + xi * a) * (1. - ze * ze);
      dz[i] = -.5 * ze * (1. + xi * a) * (1. + et * b);
    }
  }
}

__global__ void build_hex20_ops_kernel(int nelt, const int* __restrict__ elem, const int* __restrict__ conn,
                                       const double* __restrict__ xyz, const double* __restrict__ emod,
                                       const double* __restrict__ rny, int stressForm, double* __restrict__ Sfrag,
                                       unsigned char* __restrict__ failed, double* __restrict__ aux, double* __restrict__ Gfrag,
                                       double* __restrict__ fastJ /* [nelt][20][9]: J^-1 of the nodal evaluation points, or NULL */)
{
  constexpr int KT = 15, MT = 15;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nelt) return;
  const int e = elem[i];
  double* S = Sfrag + (size_t)i * MT * KT * 32;
  double* G = Gfrag + (size_t)i * 9 * 5 * 32;   // displacement-gradient operator, see k2_hex20_grad_vm_kernel
  double X[20], Y[20], Z[20];
  for (int k = 0; k < 20; ++k) {
    const int n = conn[i * 20 + k];
    X[k] = xyz[3 * n]; Y[k] = xyz[3 * n + 1]; Z[k] = xyz[3 * n + 2];
  }
  const double E = emod[e], nu = rny[e];
  aux[i * 2] = E; aux[i * 2 + 1] = nu;
  const double D = E * (1. - nu) / ((1. + nu) * (1. - 2. * nu));
  const double D1 = D * nu / (1. - nu);
  const double D2 = D * (1. - 2. * nu) / (2. * (1. - nu));
  const double gp = (double)0.577350269189626f;  // REAL*4 literal of ihex.f:364-365
  const double s3 = sqrt(3.0);
  const int npt = stressForm == 0 ? 20 : 8;
  bool ok = true;
  for (int q = 0; q < npt && ok; ++q) {
    double xi, et, ze;
    if (stressForm == 0) { xi = c_hx[q]; et = c_he[q]; ze = c_hz[q]; }
    else { xi = (q & 1) ? gp : -gp; et = (q & 2) ? gp : -gp; ze = (q & 4) ? gp : -gp; }
    double dx[20], de[20], dz[20];
    hex20_dn(xi, et, ze, dx, de, dz);
    double J[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
    for (int k = 0; k < 20; ++k) {
      J[0][0] += dx[k] * X[k]; J[0][1] += dx[k] * Y[k]; J[0][2] += dx[k] * Z[k];
      J[1][0] += de[k] * X[k]; J[1][1] += de[k] * Y[k]; J[1][2] += de[k] * Z[k];
      J[2][0] += dz[k] * X[k]; J[2][1] += dz[k] * Y[k]; J[2][2] += dz[k] * Z[k];
    }
    const double det = J[0][0] * (J[1][1] * J[2][2] - J[1][2] * J[2][1]) + J[0][1] * (J[1][2] * J[2][0] - J[1][0] * J[2][2]) +
                       J[0][2] * (J[1][0] * J[2][1] - J[1][1] * J[2][0]);
    if (fabs(det) <= 2.2250738585072014e-308 * 100.0) { ok = false; break; }
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
    if (fastJ && stressForm == 0)
      for (int d = 0; d < 3; ++d)
        for (int j = 0; j < 3; ++j) fastJ[((size_t)i * 20 + q) * 9 + 3 * d + j] = I[d][j];
    for (int j = 0; j < 20; ++j) {
      const double bx = I[0][0] * dx[j] + I[0][1] * de[j] + I[0][2] * dz[j];
      const double by = I[1][0] * dx[j] + I[1][1] * de[j] + I[1][2] * dz[j];
      const double bz = I[2][0] * dx[j] + I[2][1] * de[j] + I[2][2] * dz[j];
      // D*B block of node j: rows xx,yy,zz,xy,xz,yz ; columns u,v,w  (ihex.f:430-450)
      const double db[6][3] = {{D * bx, D1 * by, D1 * bz}, {D1 * bx, D * by, D1 * bz}, {D1 * bx, D1 * by, D * bz},
                               {D2 * by, D2 * bx, 0.0},    {D2 * bz, 0.0, D2 * bx},    {0.0, D2 * bz, D2 * by}};
      const double bd[3] = {bx, by, bz};
      if (stressForm == 0) {
        for (int c = 0; c < 6; ++c)
          for (int d = 0; d < 3; ++d) S[frag_index_g(q * 6 + c, 3 * j + d, KT)] = db[c][d];
        for (int d = 0; d < 3; ++d) G[frag_index_g(((q >> 3) * 3 + d) * 8 + (q & 7), j, 5)] = bd[d];
      } else {
        for (int n = 0; n < 20; ++n) {
          const double w = (1.0 + ((q & 1) ? 1.0 : -1.0) * (s3 * c_hx[n])) * (1.0 + ((q & 2) ? 1.0 : -1.0) * (s3 * c_he[n])) *
                           (1.0 + ((q & 4) ? 1.0 : -1.0) * (s3 * c_hz[n])) * 0.125;
          for (int c = 0; c < 6; ++c)
            for (int d = 0; d < 3; ++d) S[frag_index_g(n * 6 + c, 3 * j + d, KT)] += db[c][d] * w;
          for (int d = 0; d < 3; ++d) G[frag_index_g(((n >> 3) * 3 + d) * 8 + (n & 7), j, 5)] += bd[d] * w;
        }
      }
    }
  }
  if (!ok) {
    for (int k = 0; k < MT * KT * 32; ++k) S[k] = 0.0;
    for (int k = 0; k < 9 * 5 * 32; ++k) G[k] = 0.0;
    if (fastJ)
      for (int k = 0; k < 180; ++k) fastJ[(size_t)i * 180 + k] = 0.0;
  }
  failed[i] = ok ? 0 : 1;
}

// Displacement-gradient form of the HEX20 von Mises kernel (same idea as k2_tet10_grad_vm_kernel, k2_solid.cu):
// sigma = D . sym(grad u) and grad u at the 20 result points = [60 x 20] . [20 x 3].  The result points are taken in
// three blocks of 8 DMMA rows (the last one holds 4); per block 3 m-tiles (one per derivative direction) x 5 k-tiles,
// so lane (g, t4) owns all nine gradient entries of point 8 pb + g at its two steps and no transposition is needed.
// 135 DMMA per 8 steps instead of 225 for the dense 120 x 60 operator, which stays for the full-result path.
// One warp per element; the 15 B fragments (u, v, w of the 20 nodes, 8 steps) stay in registers for the three
// blocks, the A fragments of a block stream from L1/L2 (5.6 KB per element, re-read every tile).
// The same kernel serves the 15-node wedge (type 42): NPB = 2 blocks of result points, KT = 4 k-tiles (90 DMMA).
template <int NPB, int KT, int NP, int NN>
__global__ void __launch_bounds__(128, 3)
k2_bigsolid_grad_vm_kernel(const double* __restrict__ U, size_t ldu, int nsteps, int nsteps_pad, const double* __restrict__ Gfrag,
                           const double* __restrict__ aux, const int* __restrict__ edof, const int* __restrict__ ptoff,
                           const unsigned char* __restrict__ failed, int nelt, double* __restrict__ vm, size_t ld_vm,
                           double* __restrict__ env_max, double* __restrict__ env_min)
{
  constexpr int ESTRIDE = ((3 * NN + 3) / 4) * 4;   // edof row stride = 4 * (k-tiles of the dense operator)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, t4 = lane & 3;
  const int i = blockIdx.x * (blockDim.x >> 5) + warp;
  if (i >= nelt) return;
  const double* gf = Gfrag + (size_t)i * 3 * NPB * KT * 32 + lane;
  const int* ed = edof + (size_t)i * ESTRIDE;
  const double* urow[3][KT];
#pragma unroll
  for (int j = 0; j < KT; ++j) {
    const int node = 4 * j + t4;   // nodes >= NN are padding: their operator columns are zero
#pragma unroll
    for (int c = 0; c < 3; ++c) urow[c][j] = U + (size_t)(node < NN ? __ldg(ed + 3 * node + c) : 0) * ldu + g;
  }
  const double E = __ldg(aux + (size_t)i * 2), nu = __ldg(aux + (size_t)i * 2 + 1);
  const double mu2 = E / (1.0 + nu);
  const bool bad = failed[i] != 0;
  const size_t pt0 = (size_t)ptoff[i];
  double emax[NPB], emin[NPB];
#pragma unroll
  for (int pb = 0; pb < NPB; ++pb) { emax[pb] = 0.0; emin[pb] = kHuge; }
  const int ntiles = nsteps_pad >> 3;
  double b[3][KT], bn[3][KT];
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int j = 0; j < KT; ++j) b[c][j] = urow[c][j][0];
  for (int nt = 0; nt < ntiles; ++nt) {
    if (nt + 1 < ntiles) {
#pragma unroll
      for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int j = 0; j < KT; ++j) bn[c][j] = urow[c][j][(nt + 1) * 8];
    }
#pragma unroll
    for (int pb = 0; pb < NPB; ++pb) {
      double a[3][KT];
#pragma unroll
      for (int m = 0; m < 3; ++m)
#pragma unroll
        for (int j = 0; j < KT; ++j) a[m][j] = __ldg(gf + (size_t)((pb * 3 + m) * KT + j) * 32);
      double acc[3][3][2];
#pragma unroll
      for (int m = 0; m < 3; ++m)
#pragma unroll
        for (int c = 0; c < 3; ++c) acc[m][c][0] = acc[m][c][1] = 0.0;
#pragma unroll
      for (int j = 0; j < KT; ++j)
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
          for (int m = 0; m < 3; ++m) dmma884(acc[m][c][0], acc[m][c][1], a[m][j], b[c][j]);
      const bool live = 8 * pb + g < NP;
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int t = nt * 8 + 2 * t4 + q;
        const double da = acc[0][0][q] - acc[1][1][q], db = acc[1][1][q] - acc[2][2][q], dc = acc[2][2][q] - acc[0][0][q];
        const double gxy = acc[1][0][q] + acc[0][1][q], gxz = acc[2][0][q] + acc[0][2][q], gyz = acc[2][1][q] + acc[1][2][q];
        const double dev = 0.5 * fma(da, da, fma(db, db, dc * dc));
        const double shr = fma(gxy, gxy, fma(gxz, gxz, gyz * gyz));
        double v = mu2 * sqrt_pos(fma(0.75, shr, dev));
        if (bad) v = kHuge;
        if (live && t < nsteps) {
          if (vm) vm[(size_t)t * ld_vm + pt0 + 8 * pb + g] = v;
          emax[pb] = fmax(emax[pb], v);
          emin[pb] = fmin(emin[pb], v);
        }
      }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
      for (int j = 0; j < KT; ++j) b[c][j] = bn[c][j];
  }
#pragma unroll
  for (int pb = 0; pb < NPB; ++pb) {
#pragma unroll
    for (int o = 1; o < 4; o <<= 1) {
      emax[pb] = fmax(emax[pb], __shfl_xor_sync(0xffffffffu, emax[pb], o));
      emin[pb] = fmin(emin[pb], __shfl_xor_sync(0xffffffffu, emin[pb], o));
    }
    if (t4 == 0 && nsteps > 0 && 8 * pb + g < NP) {
      const size_t pt = pt0 + 8 * pb + g;
      if (emax[pb] > env_max[pt]) env_max[pt] = emax[pb];
      if (emin[pb] < env_min[pt]) env_min[pt] = emin[pb];
    }
  }
}

// ---- nodal evaluation in natural coordinates, lane = time step ------------------------------------------------------------
// At a NODE of the serendipity element most shape-function derivatives vanish: a corner sees the three nodes of each of
// its three edges, a mid-edge node the two ends of its edge and the eight nodes of the two faces that meet there -- 288
// non-zero entries of the 20 x 3 x 20 table instead of 1,200.  So for -stressForm 0 the NATURAL derivatives
// D[c][j] = d u_c / d xi_j are formed from those entries with compile-time coefficients (-3/2, 2, -1/2, +-1/2, 1) and the
// inverse Jacobian of the point (fastJ [elem][20][9], row d, column j, from the operator builder) turns them into the
// gradient: 864 + 540 FMA per element.step against 3,600 in the dense gradient operator (4,320 issued with its padding),
// and 1.4 KB of constants per element instead of 11.5 KB of fragments.  Lane = time step (tiles of 32): the 60 nodal
// displacements of a step sit in registers, every row of U is read as one 256-byte segment per warp, no shared-memory
// staging.  The 20 x 32 values of a tile are turned through shared memory: lanes 0..19 fold the envelope of one result
// point each, and the history goes out as 160-byte pieces of the step records.
__host__ __device__ constexpr int hex20_nat(int i, int axis)
{
  constexpr int hx[20] = {-1, 0, 1, 1, 1, 0, -1, -1, -1, 1, 1, -1, -1, 0, 1, 1, 1, 0, -1, -1};
  constexpr int he[20] = {-1, -1, -1, 0, 1, 1, 1, 0, -1, -1, 1, 1, -1, -1, -1, 0, 1, 1, 1, 0};
  constexpr int hz[20] = {-1, -1, -1, -1, -1, -1, -1, -1, 0, 0, 0, 0, 1, 1, 1, 1, 1, 1, 1, 1};
  return axis == 0 ? hx[i] : axis == 1 ? he[i] : hz[i];
}
// d N_j / d xi_d at node p (DN2031, ihex.f:2433-2545, evaluated at the node coordinates)
__host__ __device__ constexpr double hex20_dn_node(int p, int d, int j)
{
  const double xi = hex20_nat(p, 0), et = hex20_nat(p, 1), ze = hex20_nat(p, 2);
  const double a = hex20_nat(j, 0), b = hex20_nat(j, 1), c = hex20_nat(j, 2);
  if (a != 0.0 && b != 0.0 && c != 0.0)
    return d == 0 ? .125 * a * (1. + et * b) * (1. + ze * c) * (2. * xi * a + et * b + ze * c - 1.)
         : d == 1 ? .125 * b * (1. + xi * a) * (1. + ze * c) * (xi * a + 2. * et * b + ze * c - 1.)
                  : .125 * c * (1. + xi * a) * (1. + et * b) * (xi * a + et * b + 2. * ze * c - 1.);
  if (a == 0.0)
    return d == 0 ? -.5 * xi * (1. + et * b) * (1. + ze * c) : d == 1 ? .25 * b * (1. - xi * xi) * (1. + ze * c) : .25 * c * (1. - xi * xi) * (1. + et * b);
  if (b == 0.0)
    return d == 0 ? .25 * a * (1. - et * et) * (1. + ze * c) : d == 1 ? -.5 * et * (1. + xi * a) * (1. + ze * c) : .25 * c * (1. + xi * a) * (1. - et * et);
  return d == 0 ? .25 * a * (1. + et * b) * (1. - ze * ze) : d == 1 ? .25 * b * (1. + xi * a) * (1. - ze * ze) : -.5 * ze * (1. + xi * a) * (1. + et * b);
}

template <int P>
struct Hex20NodeRow {   // the coefficients of result point P as compile-time constants
  double c[3][20];
  constexpr Hex20NodeRow() : c{}
  {
    for (int d = 0; d < 3; ++d)
      for (int j = 0; j < 20; ++j) c[d][j] = hex20_dn_node(P, d, j);
  }
};

template <int P>
__device__ __forceinline__ double hex20_point_vm2(const double (&u)[20][3], const double (&J)[3][3])
{
  constexpr Hex20NodeRow<P> R{};
  double D[3][3];   // D[c][j]
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    bool first = true;
#pragma unroll
    for (int n = 0; n < 20; ++n) {
      if (R.c[j][n] != 0.0) {
#pragma unroll
        for (int c = 0; c < 3; ++c) D[c][j] = first ? R.c[j][n] * u[n][c] : fma(R.c[j][n], u[n][c], D[c][j]);
        first = false;
      }
    }
  }
  double H[3][3];
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int d = 0; d < 3; ++d) H[c][d] = fma(J[d][2], D[c][2], fma(J[d][1], D[c][1], J[d][0] * D[c][0]));
  const double da = H[0][0] - H[1][1], db = H[1][1] - H[2][2], dc = H[2][2] - H[0][0];
  const double gxy = H[1][0] + H[0][1], gxz = H[2][0] + H[0][2], gyz = H[2][1] + H[1][2];
  const double dev = 0.5 * fma(da, da, fma(db, db, dc * dc));
  const double shr = fma(gxy, gxy, fma(gxz, gxz, gyz * gyz));
  return fma(0.75, shr, dev);   // (vm / 2 mu)^2
}

__device__ __forceinline__ void hex20_load_J(const double* __restrict__ fJ, int p, double (&J)[3][3])
{
#pragma unroll
  for (int d = 0; d < 3; ++d)
#pragma unroll
    for (int j = 0; j < 3; ++j) J[d][j] = __ldg(fJ + p * 9 + 3 * d + j);
}

// the inverse of point P + 1 is fetched (warp-uniform loads, L1) before the arithmetic of point P
template <int P, bool WRITE_VM>
__device__ __forceinline__ void hex20_points(const double (&u)[20][3], const double* __restrict__ fJ, const double (&J)[3][3], double mu2,
                                             bool bad, double* sv)
{
  if constexpr (P < 20) {
    double Jn[3][3];
    if constexpr (P + 1 < 20) hex20_load_J(fJ, P + 1, Jn);
    double v = hex20_point_vm2<P>(u, J);
    if (WRITE_VM) v = mu2 * sqrt_pos(v);
    if (bad) v = kHuge;
    sv[P * 33] = v;
    hex20_points<P + 1, WRITE_VM>(u, fJ, Jn, mu2, bad, sv);
  }
}

template <bool WRITE_VM, int NW, int MINB>
__global__ void __launch_bounds__(NW * 32, MINB)
k2_hex20_steplane_vm_kernel(const double* __restrict__ U, size_t ldu, int nsteps, const double* __restrict__ fastJ,
                            const double* __restrict__ aux, const int* __restrict__ edof, const int* __restrict__ ptoff,
                            const unsigned char* __restrict__ failed, int nelt, double* __restrict__ vm, size_t ld_vm,
                            double* __restrict__ env_max, double* __restrict__ env_min)
{
  __shared__ double sv_all[NW][20 * 33];
  __shared__ unsigned so_all[NW][64];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int i = blockIdx.x * NW + warp;
  if (i >= nelt) return;   // whole warp
  double* sv = sv_all[warp];
  // row offsets of the element's 60 DOFs in units of 64 doubles (ldu is a multiple of 64: 32 bits are enough for any U),
  // kept in shared memory: one broadcast LDS + one IMAD.WIDE per load
  unsigned* so = so_all[warp];
  const unsigned ldu64 = (unsigned)(ldu >> 6);
  if (lane < 30) {
    const int2 e2 = __ldg(reinterpret_cast<const int2*>(edof + (size_t)i * 60) + lane);
    so[2 * lane] = (unsigned)e2.x * ldu64;
    so[2 * lane + 1] = (unsigned)e2.y * ldu64;
  }
  __syncwarp();
  const double E = __ldg(aux + (size_t)i * 2), nu = __ldg(aux + (size_t)i * 2 + 1);
  const double mu2 = E / (1.0 + nu);
  const bool bad = failed[i] != 0;
  const size_t pt0 = (size_t)ptoff[i];
  const double* fJ = fastJ + (size_t)i * 180;
  double emax = 0.0, emin = kHuge;   // lanes 0..19: result point = lane
  for (int t0 = 0; t0 < nsteps; t0 += 32) {
    // lanes past the last step repeat it: no predicates, no zero fill (their values are not scanned below)
    const double* Ut = U + min(t0 + lane, nsteps - 1);
    double u[20][3];
#pragma unroll
    for (int n = 0; n < 20; ++n)
#pragma unroll
      for (int c = 0; c < 3; ++c) u[n][c] = __ldg(Ut + ((size_t)so[3 * n + c] << 6));
    double J0[3][3];
    hex20_load_J(fJ, 0, J0);
    hex20_points<0, WRITE_VM>(u, fJ, J0, mu2, bad, sv + lane);
    __syncwarp();
    const int ns = min(32, nsteps - t0);
    if (lane < 20) {
      const double* row = sv + lane * 33;
      for (int s = 0; s < ns; ++s) {
        const double v = row[s];
        emax = max_nonneg(emax, v); emin = min_nonneg(emin, v);
      }
    }
    if (WRITE_VM) {
#pragma unroll
      for (int r = 0; r < 20; ++r) {
        const int idx = lane + 32 * r, s = idx / 20, p = idx - 20 * s;
        if (s < ns) vm[(size_t)(t0 + s) * ld_vm + pt0 + p] = sv[p * 33 + s];
      }
    }
    __syncwarp();
  }
  if (lane < 20 && nsteps > 0) {
    if (!WRITE_VM && !bad) {   // radicand -> von Mises
      emax = mu2 * sqrt_pos(emax);
      emin = mu2 * sqrt_pos(emin);
    }
    if (emax > env_max[pt0 + lane]) env_max[pt0 + lane] = emax;
    if (emin < env_min[pt0 + lane]) env_min[pt0 + lane] = emin;
  }
}

// Generic solid apply: NEN nodes, 6 stress components per node, 3 DOFs per node.
template <int NEN>
__global__ void __launch_bounds__(128)
k2_solid_smem_vm_kernel(const double* __restrict__ U, size_t ldu, int nsteps, int nsteps_pad,
                        const double* __restrict__ Sfrag, const int* __restrict__ edof, const int* __restrict__ ptoff,
                        const unsigned char* __restrict__ failed, int nelt, double* __restrict__ vm, size_t ld_vm,
                        double* __restrict__ env_max, double* __restrict__ env_min)
{
  constexpr int NROW = 6 * NEN, NCOL = 3 * NEN;
  constexpr int MT = (NROW + 7) / 8, KT = (NCOL + 3) / 4;
  constexpr int NITEM = NEN * 8;                       // (node, step-in-tile) pairs per tile
  constexpr int IPT = (NITEM + 127) / 128;             // items per thread
  extern __shared__ __align__(16) double smem[];
  double* sS = smem;                                   // [MT][KT][32] operator fragments
  double* sU = sS + MT * KT * 32;                      // [KT*4][8]   displacement tile
  double* sSig = sU + KT * 4 * 8;                      // [MT*8][8]   stresses of the tile
  int* sDof = reinterpret_cast<int*>(sSig + MT * 8 * 8);  // [KT*4]
  const int i = blockIdx.x;
  if (i >= nelt) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t4 = lane & 3;
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
    // stage the element's rows of U for these 8 steps (64-byte segments)
    for (int k = tid; k < KT * 4 * 8; k += 128) {
      const int row = k >> 3, s = k & 7;
      sU[k] = row < NCOL ? U[(size_t)sDof[row] * ldu + (size_t)nt * 8 + s] : 0.0;
    }
    __syncthreads();
    for (int m = warp; m < MT; m += 4) {               // warp-uniform: every lane reaches each mma
      double c0 = 0.0, c1 = 0.0;
#pragma unroll 5
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
    __syncthreads();
  }
  // the 8 lanes that share a node are consecutive: reduce, lane with step 0 folds into the envelope
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

template <int NEN>
static size_t solid_smem_bytes()
{
  constexpr int MT = (6 * NEN + 7) / 8, KT = (3 * NEN + 3) / 4;
  return sizeof(double) * (MT * KT * 32 + KT * 4 * 8 + MT * 8 * 8) + sizeof(int) * KT * 4;
}

int build_hex20_operators(fsr_part* p, const fsr_sam* sam, const fsr_elmdata* elm)
{
  cudaStream_t s = p->stream;
  FamilyData& f = p->fam[FAM_HEX20];
  f.nenod = 20; f.nndof = 3; f.nstrp = 20; f.ncmp = 6; f.MT = 15; f.KT = 15;
  std::vector<int> elem, conn, edof, ptoff;
  for (int e : elements_of_type(p, sam, elm, 43)) {
    const int ip0 = sam->mpmnpc[e] - 1, nn = sam->mpmnpc[e + 1] - sam->mpmnpc[e];
    if (nn != 20) { set_error("HEX20 element %d has %d nodes", e + 1, nn); return FSR_ERR_ARG; }
    elem.push_back(e);
    ptoff.push_back(p->ptoff_host[e]);
    const size_t base = edof.size();
    edof.resize(base + 60, 0);
    for (int k = 0; k < 20; ++k) {
      const int n = sam->mmnpc[ip0 + k] - 1;
      if (n < 0 || n >= sam->nnod) { set_error("element %d: node index out of range", e + 1); return FSR_ERR_ARG; }
      conn.push_back(n);
      const int js = sam->madof[n] - 1, nd = sam->madof[n + 1] - sam->madof[n];
      if (nd < 3) { set_error("element %d: node %d has %d DOFs, solid needs 3", e + 1, n + 1, nd); return FSR_ERR_ARG; }
      for (int d = 0; d < 3; ++d) edof[base + (size_t)k * 3 + d] = js + d;
    }
  }
  f.nelt = (int)elem.size();
  f.naux = 2;
  if (f.nelt == 0) return FSR_OK;
  int* d_conn = nullptr;
  FSR_CUDA(cudaMalloc(&f.elem, sizeof(int) * elem.size()));
  FSR_CUDA(cudaMalloc(&f.edof, sizeof(int) * edof.size()));
  FSR_CUDA(cudaMalloc(&f.ptoff, sizeof(int) * ptoff.size()));
  FSR_CUDA(cudaMalloc(&f.failed, f.nelt));
  FSR_CUDA(cudaMalloc(&f.Sfrag, sizeof(double) * (size_t)f.nelt * f.MT * f.KT * 32));
  FSR_CUDA(cudaMalloc(&f.aux, sizeof(double) * (size_t)f.nelt * f.naux));
  FSR_CUDA(cudaMalloc(&f.Gfrag, sizeof(double) * (size_t)f.nelt * 9 * 5 * 32));
  FSR_CUDA(cudaMemsetAsync(f.Gfrag, 0, sizeof(double) * (size_t)f.nelt * 9 * 5 * 32, s));
  if (p->stressForm == 0) {   // nodal evaluation: the twenty inverses of every element for the natural-coordinate kernel
    FSR_CUDA(cudaMalloc(&f.fast2, sizeof(double) * (size_t)f.nelt * 180));
    FSR_CUDA(cudaMemsetAsync(f.fast2, 0, sizeof(double) * (size_t)f.nelt * 180, s));
  }
  FSR_CUDA(cudaMalloc(&d_conn, sizeof(int) * conn.size()));
  FSR_CUDA(cudaMemcpyAsync(f.elem, elem.data(), sizeof(int) * elem.size(), cudaMemcpyHostToDevice, s));
  FSR_CUDA(cudaMemcpyAsync(f.edof, edof.data(), sizeof(int) * edof.size(), cudaMemcpyHostToDevice, s));
  FSR_CUDA(cudaMemcpyAsync(f.ptoff, ptoff.data(), sizeof(int) * ptoff.size(), cudaMemcpyHostToDevice, s));
  FSR_CUDA(cudaMemcpyAsync(d_conn, conn.data(), sizeof(int) * conn.size(), cudaMemcpyHostToDevice, s));
  FSR_CUDA(cudaMemsetAsync(f.Sfrag, 0, sizeof(double) * (size_t)f.nelt * f.MT * f.KT * 32, s));
  build_hex20_ops_kernel<<<(f.nelt + 31) / 32, 32, 0, s>>>(f.nelt, f.elem, d_conn, p->xyz, p->emod, p->rny, p->stressForm,
                                                         f.Sfrag, f.failed, f.aux, f.Gfrag, f.fast2);
  FSR_LAUNCH_CHECK();
  FSR_CUDA(cudaStreamSynchronize(s));
  cudaFree(d_conn);
  return FSR_OK;
}

int launch_k2_hex20_vm(fsr_part* p, int nsteps, int nsteps_pad, double* vm_dev, size_t ld_vm, cudaStream_t s)
{
  FamilyData& f = p->fam[FAM_HEX20];
  if (f.nelt == 0) return FSR_OK;
  // FSR_HEX20_DENSE=1 selects the dense 120x60 shared-memory formulation (A/B timing, cross-check)
  static const bool dense = getenv("FSR_HEX20_DENSE") && atoi(getenv("FSR_HEX20_DENSE")) != 0;
  // tiles of 32 steps and more, nodal evaluation: natural-coordinate form, lane = step (FSR_HEX20_STEPLANE=0: A/B, cross-check)
  const bool steplane = !(getenv("FSR_HEX20_STEPLANE") && atoi(getenv("FSR_HEX20_STEPLANE")) == 0);
  if (!dense && steplane && f.fast2 && nsteps >= 32) {
    constexpr int NW = 4;
    const int grid = (f.nelt + NW - 1) / NW;
    const bool b3 = getenv("FSR_HEX20_MINB") && atoi(getenv("FSR_HEX20_MINB")) == 3;   // A/B: 168 registers, 12 warps per SM
#define FSR_HEX20_SL(W, B)                                                                                                      \
  k2_hex20_steplane_vm_kernel<W, NW, B><<<grid, NW * 32, 0, s>>>(p->U, (size_t)p->step_tile, nsteps, f.fast2, f.aux, f.edof, f.ptoff, \
                                                                 f.failed, f.nelt, vm_dev, ld_vm, p->env_max, p->env_min)
    if (vm_dev) { if (b3) FSR_HEX20_SL(true, 3); else FSR_HEX20_SL(true, 2); }
    else { if (b3) FSR_HEX20_SL(false, 3); else FSR_HEX20_SL(false, 2); }
#undef FSR_HEX20_SL
    FSR_LAUNCH_CHECK();
    return FSR_OK;
  }
  if (!dense) {
    const int warps = 4;
    k2_bigsolid_grad_vm_kernel<3, 5, 20, 20><<<(f.nelt + warps - 1) / warps, warps * 32, 0, s>>>(p->U, (size_t)p->step_tile, nsteps, nsteps_pad, f.Gfrag,
                                                                               f.aux, f.edof, f.ptoff, f.failed, f.nelt, vm_dev,
                                                                               ld_vm, p->env_max, p->env_min);
    FSR_LAUNCH_CHECK();
    return FSR_OK;
  }
  const size_t smem = solid_smem_bytes<20>();
  if (int rc = smem_opt_in((const void*)k2_solid_smem_vm_kernel<20>, smem)) return rc;
  k2_solid_smem_vm_kernel<20><<<f.nelt, 128, smem, s>>>(p->U, (size_t)p->step_tile, nsteps, nsteps_pad, f.Sfrag, f.edof,
                                                       f.ptoff, f.failed, f.nelt, vm_dev, ld_vm, p->env_max, p->env_min);
  FSR_LAUNCH_CHECK();
  return FSR_OK;
}

// ---- 15-node wedge (type 42) -----------------------------------------------------------------------------------
// Reference: STR42 -> IPRI32 -> DN1531 / JACI31 (src/vpmStress/elStressModule.f90:1381-1465, src/Femlib/ipri.f:391-767,
// 2867-2971).  -stressForm 0 evaluates at the 15 nodes; otherwise at 3 mid-side points x 2 Gauss levels (abscissa in the
// reference's REAL*4 precision, ipri.f "ZE(1)=-.577350269189626") with the linear extrapolation of STR42 (:1441-1462).
struct Wedg15Points {
  int npt;
  double L[18][4];     // L1, L2, L3, zeta of each evaluation point
  double W[15][18];    // result point p = sum_g W[p][g] * evaluation point g
};

__device__ void wedg15_dn(double RL1, double RL2, double RL3, double ZE, double* DNL1, double* DNL2, double* DNZE)
{
  const double RL1RL1 = RL1 * RL1, RL1RL2 = RL1 * RL2, RL1RL3 = RL1 * RL3, RL1ZE = RL1 * ZE, RL2RL2 = RL2 * RL2, RL2RL3 = RL2 * RL3,
               RL2ZE = RL2 * ZE, RL3RL3 = RL3 * RL3, RL3ZE = RL3 * ZE, ZEZE = ZE * ZE;
  DNL1[0] = -1. + 0.5 * ZE + 0.5 * ZEZE + 2. * RL1 - 2. * RL1ZE; DNL1[1] = 2. * RL2 - 2. * RL2ZE; DNL1[2] = 0.;
  DNL1[3] = -2. * RL2 + 2. * RL2ZE; DNL1[4] = 1. - 0.5 * ZE - 0.5 * ZEZE - 2. * RL3 + 2. * RL3ZE; DNL1[5] = 2. * (1. - ZE) * (RL3 - RL1);
  DNL1[6] = 1. - ZEZE; DNL1[7] = 0.; DNL1[8] = -1. + ZEZE;
  DNL1[9] = -1. - 0.5 * ZE + 0.5 * ZEZE + 2. * RL1 + 2. * RL1ZE; DNL1[10] = 2. * RL2 + 2. * RL2ZE; DNL1[11] = 0.;
  DNL1[12] = -2. * RL2 - 2. * RL2ZE; DNL1[13] = 1. + 0.5 * ZE - 0.5 * ZEZE - 2. * RL3 - 2. * RL3ZE; DNL1[14] = 2. * (1. + ZE) * (RL3 - RL1);
  DNL2[0] = 0.; DNL2[1] = 2. * RL1 - 2. * RL1ZE; DNL2[2] = -1. + 0.5 * ZE + 0.5 * ZEZE + 2. * RL2 - 2. * RL2ZE;
  DNL2[3] = 2. * (RL3 - RL2) * (1. - ZE); DNL2[4] = 1. - 0.5 * ZE - 0.5 * ZEZE - 2. * RL3 + 2. * RL3ZE; DNL2[5] = -2. * RL1 + 2. * RL1ZE;
  DNL2[6] = 0.; DNL2[7] = 1. - ZEZE; DNL2[8] = -1. + ZEZE;
  DNL2[9] = 0.; DNL2[10] = 2. * RL1 + 2. * RL1ZE; DNL2[11] = -1. - 0.5 * ZE + 0.5 * ZEZE + 2. * RL2 + 2. * RL2ZE;
  DNL2[12] = 2. * (RL3 - RL2) * (1. + ZE); DNL2[13] = 1. + 0.5 * ZE - 0.5 * ZEZE - 2. * RL3 - 2. * RL3ZE; DNL2[14] = -2. * RL1 - 2. * RL1ZE;
  DNZE[0] = 0.5 * RL1 + RL1ZE - RL1RL1; DNZE[1] = -2. * RL1RL2; DNZE[2] = 0.5 * RL2 + RL2ZE - RL2RL2;
  DNZE[3] = -2. * RL2RL3; DNZE[4] = 0.5 * RL3 + RL3ZE - RL3RL3; DNZE[5] = -2. * RL1RL3;
  DNZE[6] = -2. * RL1ZE; DNZE[7] = -2. * RL2ZE; DNZE[8] = -2. * RL3ZE;
  DNZE[9] = -0.5 * RL1 + RL1ZE + RL1RL1; DNZE[10] = 2. * RL1RL2; DNZE[11] = -0.5 * RL2 + RL2ZE + RL2RL2;
  DNZE[12] = 2. * RL2RL3; DNZE[13] = -0.5 * RL3 + RL3ZE + RL3RL3; DNZE[14] = 2. * RL1RL3;
}

__global__ void build_wedg15_ops_kernel(int nelt, const int* __restrict__ elem, const int* __restrict__ conn,
                                        const double* __restrict__ xyz, const double* __restrict__ emod,
                                        const double* __restrict__ rny, const Wedg15Points* __restrict__ pts,
                                        double* __restrict__ Sfrag, double* __restrict__ Gfrag, unsigned char* __restrict__ failed,
                                        double* __restrict__ aux)
{
  constexpr int MT = 12, KT = 12, KTG = 4;   // dense 90 x 45 operator; gradient operator 2 blocks x 3 x 8 rows, 16 columns
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nelt) return;
  const int e = elem[i];
  double* S = Sfrag + (size_t)i * MT * KT * 32;
  double* G = Gfrag + (size_t)i * 6 * KTG * 32;
  double X[15], Y[15], Z[15];
  for (int k = 0; k < 15; ++k) {
    const int n = conn[i * 15 + k];
    X[k] = xyz[3 * n]; Y[k] = xyz[3 * n + 1]; Z[k] = xyz[3 * n + 2];
  }
  const double E = emod[e], nu = rny[e];
  aux[i * 2] = E; aux[i * 2 + 1] = nu;
  const double D = E * (1. - nu) / ((1. + nu) * (1. - 2. * nu));
  const double D1 = D * nu / (1. - nu);
  const double D2 = D * (1. - 2. * nu) / (2. * (1. - nu));
  bool ok = true;
  for (int q = 0; q < pts->npt && ok; ++q) {
    double dx[15], de[15], dz[15];
    wedg15_dn(pts->L[q][0], pts->L[q][1], pts->L[q][2], pts->L[q][3], dx, de, dz);
    double J[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
    for (int k = 0; k < 15; ++k) {
      J[0][0] += dx[k] * X[k]; J[0][1] += dx[k] * Y[k]; J[0][2] += dx[k] * Z[k];
      J[1][0] += de[k] * X[k]; J[1][1] += de[k] * Y[k]; J[1][2] += de[k] * Z[k];
      J[2][0] += dz[k] * X[k]; J[2][1] += dz[k] * Y[k]; J[2][2] += dz[k] * Z[k];
    }
    const double det = J[0][0] * (J[1][1] * J[2][2] - J[1][2] * J[2][1]) + J[0][1] * (J[1][2] * J[2][0] - J[1][0] * J[2][2]) +
                       J[0][2] * (J[1][0] * J[2][1] - J[1][1] * J[2][0]);
    if (fabs(det) <= 2.2250738585072014e-308 * 100.0) { ok = false; break; }
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
    for (int j = 0; j < 15; ++j) {
      const double bd[3] = {I[0][0] * dx[j] + I[0][1] * de[j] + I[0][2] * dz[j], I[1][0] * dx[j] + I[1][1] * de[j] + I[1][2] * dz[j],
                            I[2][0] * dx[j] + I[2][1] * de[j] + I[2][2] * dz[j]};
      const double bx = bd[0], by = bd[1], bz = bd[2];
      const double db[6][3] = {{D * bx, D1 * by, D1 * bz}, {D1 * bx, D * by, D1 * bz}, {D1 * bx, D1 * by, D * bz},
                               {D2 * by, D2 * bx, 0.0},    {D2 * bz, 0.0, D2 * bx},    {0.0, D2 * bz, D2 * by}};
      for (int n = 0; n < 15; ++n) {
        const double w = pts->W[n][q];
        if (w == 0.0) continue;
        for (int c = 0; c < 6; ++c)
          for (int d = 0; d < 3; ++d) S[frag_index_g(n * 6 + c, 3 * j + d, KT)] += db[c][d] * w;
        for (int d = 0; d < 3; ++d) G[frag_index_g(((n >> 3) * 3 + d) * 8 + (n & 7), j, KTG)] += bd[d] * w;
      }
    }
  }
  if (!ok) {
    for (int k = 0; k < MT * KT * 32; ++k) S[k] = 0.0;
    for (int k = 0; k < 6 * KTG * 32; ++k) G[k] = 0.0;
  }
  failed[i] = ok ? 0 : 1;
}

int build_wedg15_operators(fsr_part* p, const fsr_sam* sam, const fsr_elmdata* elm)
{
  cudaStream_t s = p->stream;
  FamilyData& f = p->fam[FAM_WEDG15];
  f.nenod = 15; f.nndof = 3; f.nstrp = 15; f.ncmp = 6; f.MT = 12; f.KT = 12; f.naux = 2;
  std::vector<int> elem, conn, edof, ptoff;
  for (int e : elements_of_type(p, sam, elm, 42)) {
    const int ip0 = sam->mpmnpc[e] - 1, nn = sam->mpmnpc[e + 1] - sam->mpmnpc[e];
    if (nn != 15) { set_error("WEDG15 element %d has %d nodes", e + 1, nn); return FSR_ERR_ARG; }
    elem.push_back(e);
    ptoff.push_back(p->ptoff_host[e]);
    const size_t base = edof.size();
    edof.resize(base + 48, 0);
    for (int k = 0; k < 15; ++k) {
      const int n = sam->mmnpc[ip0 + k] - 1;
      if (n < 0 || n >= sam->nnod) { set_error("element %d: node index out of range", e + 1); return FSR_ERR_ARG; }
      conn.push_back(n);
      const int js = sam->madof[n] - 1, nd = sam->madof[n + 1] - sam->madof[n];
      if (nd < 3) { set_error("element %d: node %d has %d DOFs, solid needs 3", e + 1, n + 1, nd); return FSR_ERR_ARG; }
      for (int d = 0; d < 3; ++d) edof[base + (size_t)k * 3 + d] = js + d;
    }
  }
  f.nelt = (int)elem.size();
  if (f.nelt == 0) return FSR_OK;
  Wedg15Points h;
  memset(&h, 0, sizeof(h));
  const double RL[6][3] = {{1, 0, 0}, {.5, .5, 0}, {0, 1, 0}, {0, .5, .5}, {0, 0, 1}, {.5, 0, .5}};
  if (p->stressForm == 0) {   // the 15 nodes: 6 at zeta = -1, corners at zeta = 0, 6 at zeta = +1
    h.npt = 15;
    for (int n = 0; n < 15; ++n) {
      const int k = n < 6 ? n : n < 9 ? 2 * (n - 6) : n - 9;
      for (int c = 0; c < 3; ++c) h.L[n][c] = RL[k][c];
      h.L[n][3] = n < 6 ? -1.0 : n < 9 ? 0.0 : 1.0;
      h.W[n][n] = 1.0;
    }
  } else {                    // IPRI32 with NSTRP = 3, NSTRPZ = 2, then the extrapolation of STR42 applied to unit vectors
    h.npt = 6;
    const double ze[2] = {(double)-.577350269189626f, (double).577350269189626f};
    const int mid[3] = {1, 3, 5};   // RL1(1..3) of the NSTRP = 3 branch are the mid-side points
    for (int l = 0; l < 2; ++l)
      for (int k = 0; k < 3; ++k) {
        for (int c = 0; c < 3; ++c) h.L[3 * l + k][c] = RL[mid[k]][c];
        h.L[3 * l + k][3] = ze[l];
      }
    const double zm1 = 0.5 * sqrt(3.0) - 0.5, zp1 = zm1 + 1.0;
    for (int g = 0; g < 6; ++g) {
      double SG[6] = {0, 0, 0, 0, 0, 0}, EP[6], SI[15];
      SG[g] = 1.0;
      EP[0] = SG[2] + SG[0] - SG[1]; EP[1] = SG[1] + SG[0] - SG[2]; EP[2] = SG[1] + SG[2] - SG[0];
      EP[3] = SG[5] + SG[3] - SG[4]; EP[4] = SG[4] + SG[3] - SG[5]; EP[5] = SG[4] + SG[5] - SG[3];
      SI[0] = zp1 * EP[0] - zm1 * EP[3]; SI[2] = zp1 * EP[1] - zm1 * EP[4]; SI[4] = zp1 * EP[2] - zm1 * EP[5];
      SI[9] = zp1 * EP[3] - zm1 * EP[0]; SI[11] = zp1 * EP[4] - zm1 * EP[1]; SI[13] = zp1 * EP[5] - zm1 * EP[2];
      SI[1] = 0.5 * (SI[0] + SI[2]); SI[3] = 0.5 * (SI[2] + SI[4]); SI[5] = 0.5 * (SI[4] + SI[0]);
      SI[6] = 0.5 * (SI[0] + SI[9]); SI[7] = 0.5 * (SI[2] + SI[11]); SI[8] = 0.5 * (SI[4] + SI[13]);
      SI[10] = 0.5 * (SI[9] + SI[11]); SI[12] = 0.5 * (SI[11] + SI[13]); SI[14] = 0.5 * (SI[13] + SI[9]);
      for (int n = 0; n < 15; ++n) h.W[n][g] = SI[n];
    }
  }
  Wedg15Points* d_pts = nullptr;
  int* d_conn = nullptr;
  FSR_CUDA(cudaMalloc(&d_pts, sizeof(h)));
  FSR_CUDA(cudaMemcpyAsync(d_pts, &h, sizeof(h), cudaMemcpyHostToDevice, s));
  FSR_CUDA(cudaMalloc(&f.elem, sizeof(int) * elem.size()));
  FSR_CUDA(cudaMalloc(&f.edof, sizeof(int) * edof.size()));
  FSR_CUDA(cudaMalloc(&f.ptoff, sizeof(int) * ptoff.size()));
  FSR_CUDA(cudaMalloc(&f.failed, f.nelt));
  FSR_CUDA(cudaMalloc(&f.Sfrag, sizeof(double) * (size_t)f.nelt * f.MT * f.KT * 32));
  FSR_CUDA(cudaMalloc(&f.Gfrag, sizeof(double) * (size_t)f.nelt * 6 * 4 * 32));
  FSR_CUDA(cudaMalloc(&f.aux, sizeof(double) * (size_t)f.nelt * f.naux));
  FSR_CUDA(cudaMalloc(&d_conn, sizeof(int) * conn.size()));
  FSR_CUDA(cudaMemcpyAsync(f.elem, elem.data(), sizeof(int) * elem.size(), cudaMemcpyHostToDevice, s));
  FSR_CUDA(cudaMemcpyAsync(f.edof, edof.data(), sizeof(int) * edof.size(), cudaMemcpyHostToDevice, s));
  FSR_CUDA(cudaMemcpyAsync(f.ptoff, ptoff.data(), sizeof(int) * ptoff.size(), cudaMemcpyHostToDevice, s));
  FSR_CUDA(cudaMemcpyAsync(d_conn, conn.data(), sizeof(int) * conn.size(), cudaMemcpyHostToDevice, s));
  FSR_CUDA(cudaMemsetAsync(f.Sfrag, 0, sizeof(double) * (size_t)f.nelt * f.MT * f.KT * 32, s));
  FSR_CUDA(cudaMemsetAsync(f.Gfrag, 0, sizeof(double) * (size_t)f.nelt * 6 * 4 * 32, s));
  build_wedg15_ops_kernel<<<(f.nelt + 31) / 32, 32, 0, s>>>(f.nelt, f.elem, d_conn, p->xyz, p->emod, p->rny, d_pts, f.Sfrag, f.Gfrag,
                                                          f.failed, f.aux);
  FSR_LAUNCH_CHECK();
  FSR_CUDA(cudaStreamSynchronize(s));
  cudaFree(d_conn);
  cudaFree(d_pts);
  return FSR_OK;
}

int launch_k2_wedg15_vm(fsr_part* p, int nsteps, int nsteps_pad, double* vm_dev, size_t ld_vm, cudaStream_t s)
{
  FamilyData& f = p->fam[FAM_WEDG15];
  if (f.nelt == 0) return FSR_OK;
  const int warps = 4;
  k2_bigsolid_grad_vm_kernel<2, 4, 15, 15><<<(f.nelt + warps - 1) / warps, warps * 32, 0, s>>>(
      p->U, (size_t)p->step_tile, nsteps, nsteps_pad, f.Gfrag, f.aux, f.edof, f.ptoff, f.failed, f.nelt, vm_dev, ld_vm, p->env_max,
      p->env_min);
  FSR_LAUNCH_CHECK();
  return FSR_OK;
}

}  // namespace fsr
