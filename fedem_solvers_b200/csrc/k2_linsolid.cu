// k2_linsolid.cu -- K2 for the linear solid elements: 8-node hexahedron (type 44), 4-node tetrahedron (45),
// 6-node wedge (46), on sm_100a.
//
// Reference: STR44 -> HEXA32 / JABN30 / LINHEX (src/vpmStress/elStressModule.f90:1590-1689, src/Femlib/hexa.f:278-413,
// 802-880,1116-1156; compatible element, -useIncompatibleModes off), STR45 -> CSTetStrain / cstetbmat / pdvcoor
// (elStressModule.f90:1695-1727, src/Femlib/cstetra.f90:23-116,528-615), STR46 -> Ipri6Strain / ipri6bmat / pdvn /
// ipri6extrapolH (elStressModule.f90:1733-1764, src/Femlib/ipri6.f90:392-533,614-816) + JACI31 (jaci31.f): per element
// per step a handful of Jacobian inversions and a B-matrix product.
//
// Same design as the TET10 kernel (k2_solid.cu): for an isotropic solid sigma = D . sym(grad u), and grad u at the
// result points is a small dense operator G (rows = derivative direction x result point, columns = element nodes;
// Gauss-point evaluation and extrapolation to the nodes -- -stressForm 1/2 -- are linear and folded into G) applied to
// the nodal displacements as three right-hand sides u, v, w.  All three types have at most 8 result points, i.e. one
// block of 8 DMMA rows per derivative direction: 3 m-tiles x KT k-tiles (KT = 1 TET4, 2 WEDG6/HEX8) in registers,
// 9 KT DMMA per 8 time steps, lane (g, t4) owns all nine gradient entries of result point g at its two steps, D and
// the deviatoric von Mises are applied on the accumulators, the envelope is fused.  One kernel template serves the
// three types.  The dense 6 nstrp x 3 nenod operator (component order of the reference: HEX8 xx,yy,zz,xy,xz,yz;
// TET4 and WEDG6 xx,yy,zz,xy,yz,zx) is built alongside for the full-result / record kernels.
#include <cmath>
#include <cstring>

#include "common.cuh"

namespace fsr {

namespace {

struct LinSolidSpec {
  int type, nn, neval, shear;     // shear: 0 = (xy,xz,yz), 1 = (xy,yz,zx)
  double pt[8][4];                // evaluation points: (xi, eta, zeta) or, for the wedge, (xi1, xi2, xi3, zeta)
  double W[8][8];                 // result point p = sum_g W[p][g] * evaluation point g
  int volume_average;             // HEX8 -stressForm 1: W[p][g] = detJ_g / sum detJ (geometry dependent)
};

__device__ __forceinline__ size_t fragidx(int row, int col, int KT)
{
  return ((size_t)((row >> 3) * KT + (col >> 2)) << 5) + ((row & 7) << 2) + (col & 3);
}

// inverse Jacobian from the shape-function derivatives (JACI31): returns false when singular
__device__ bool jac_inverse(int nn, const double* d1, const double* d2, const double* d3, const double* X, const double* Y,
                            const double* Z, double I[3][3], double& adet)
{
  double J[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
  for (int k = 0; k < nn; ++k) {
    J[0][0] += d1[k] * X[k]; J[0][1] += d1[k] * Y[k]; J[0][2] += d1[k] * Z[k];
    J[1][0] += d2[k] * X[k]; J[1][1] += d2[k] * Y[k]; J[1][2] += d2[k] * Z[k];
    J[2][0] += d3[k] * X[k]; J[2][1] += d3[k] * Y[k]; J[2][2] += d3[k] * Z[k];
  }
  const double det = J[0][0] * (J[1][1] * J[2][2] - J[1][2] * J[2][1]) + J[0][1] * (J[1][2] * J[2][0] - J[1][0] * J[2][2]) +
                     J[0][2] * (J[1][0] * J[2][1] - J[1][1] * J[2][0]);
  if (fabs(det) <= 2.2250738585072014e-308 * 100.0) return false;
  I[0][0] = (J[1][1] * J[2][2] - J[1][2] * J[2][1]) / det;
  I[0][1] = (J[2][1] * J[0][2] - J[2][2] * J[0][1]) / det;
  I[0][2] = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) / det;
  I[1][0] = (J[2][0] * J[1][2] - J[2][2] * J[1][0]) / det;
  I[1][1] = (J[0][0] * J[2][2] - J[0][2] * J[2][0]) / det;
  I[1][2] = (J[1][0] * J[0][2] - J[1][2] * J[0][0]) / det;
  I[2][0] = (J[1][0] * J[2][1] - J[1][1] * J[2][0]) / det;
  I[2][1] = (J[2][0] * J[0][1] - J[2][1] * J[0][0]) / det;
  I[2][2] = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) / det;
  adet = fabs(det);
  return true;
}

// Cartesian shape-function gradients b[d][k] = dN_k/dx_d at one evaluation point; false = element failed
__device__ bool lin_gradients(const LinSolidSpec& sp, int gpt, const double* X, const double* Y, const double* Z, double b[3][8],
                              double& detj)
{
  detj = 1.0;
  if (sp.type == 45) {   // cstetbmat: constant gradients a, b, c / (6 V), V from the triple product, must be positive
    const double s12[3] = {X[1] - X[0], Y[1] - Y[0], Z[1] - Z[0]}, s13[3] = {X[2] - X[0], Y[2] - Y[0], Z[2] - Z[0]},
                 s14[3] = {X[3] - X[0], Y[3] - Y[0], Z[3] - Z[0]};
    const double cr[3] = {s12[1] * s13[2] - s12[2] * s13[1], s12[2] * s13[0] - s12[0] * s13[2], s12[0] * s13[1] - s12[1] * s13[0]};
    const double vol = (cr[0] * s14[0] + cr[1] * s14[1] + cr[2] * s14[2]) / 6.0;
    if (!(vol > kEpsDiv0)) return false;
    const double f = 1.0 / (6.0 * vol);
    b[0][0] = ((Y[1] - Y[2]) * (Z[3] - Z[1]) - (Y[3] - Y[1]) * (Z[1] - Z[2])) * f;
    b[0][1] = ((Y[3] - Y[2]) * (Z[0] - Z[2]) - (Y[0] - Y[2]) * (Z[3] - Z[2])) * f;
    b[0][2] = ((Y[3] - Y[0]) * (Z[1] - Z[3]) - (Y[1] - Y[3]) * (Z[3] - Z[0])) * f;
    b[0][3] = ((Y[1] - Y[0]) * (Z[2] - Z[0]) - (Y[2] - Y[0]) * (Z[1] - Z[0])) * f;
    b[1][0] = ((Z[1] - Z[2]) * (X[3] - X[1]) - (Z[3] - Z[1]) * (X[1] - X[2])) * f;
    b[1][1] = ((Z[3] - Z[2]) * (X[0] - X[2]) - (Z[0] - Z[2]) * (X[3] - X[2])) * f;
    b[1][2] = ((Z[3] - Z[0]) * (X[1] - X[3]) - (Z[1] - Z[3]) * (X[3] - X[0])) * f;
    b[1][3] = ((Z[1] - Z[0]) * (X[2] - X[0]) - (Z[2] - Z[0]) * (X[1] - X[0])) * f;
    b[2][0] = ((X[1] - X[2]) * (Y[3] - Y[1]) - (X[3] - X[1]) * (Y[1] - Y[2])) * f;
    b[2][1] = ((X[3] - X[2]) * (Y[0] - Y[2]) - (X[0] - X[2]) * (Y[3] - Y[2])) * f;
    b[2][2] = ((X[3] - X[0]) * (Y[1] - Y[3]) - (X[1] - X[3]) * (Y[3] - Y[0])) * f;
    b[2][3] = ((X[1] - X[0]) * (Y[2] - Y[0]) - (X[2] - X[0]) * (Y[1] - Y[0])) * f;
    return true;
  }
  double d1[8], d2[8], d3[8], I[3][3];
  if (sp.type == 46) {   // pdvn (ipri6.f90:490-533)
    const double zeta = sp.pt[gpt][3];
    d1[0] = (1.0 - zeta) * 0.5; d1[3] = (1.0 + zeta) * 0.5; d1[2] = -d1[0]; d1[5] = -d1[3]; d1[1] = 0.0; d1[4] = 0.0;
    d2[1] = d1[0]; d2[4] = d1[3]; d2[2] = -d1[0]; d2[5] = -d1[3]; d2[0] = 0.0; d2[3] = 0.0;
    for (int i = 0; i < 3; ++i) { d3[i] = -sp.pt[gpt][i] * 0.5; d3[3 + i] = sp.pt[gpt][i] * 0.5; }
    if (!jac_inverse(6, d1, d2, d3, X, Y, Z, I, detj)) return false;
  } else {               // trilinear hexahedron, node table of HEXA32; JABN30 rejects a non-positive determinant
    const double cxi[8] = {-1., 1., 1., -1., -1., 1., 1., -1.}, cet[8] = {-1., -1., 1., 1., -1., -1., 1., 1.},
                 cze[8] = {-1., -1., -1., -1., 1., 1., 1., 1.};
    const double xi = sp.pt[gpt][0], et = sp.pt[gpt][1], ze = sp.pt[gpt][2];
    for (int k = 0; k < 8; ++k) {
      d1[k] = 0.125 * cxi[k] * (1. + et * cet[k]) * (1. + ze * cze[k]);
      d2[k] = 0.125 * cet[k] * (1. + xi * cxi[k]) * (1. + ze * cze[k]);
      d3[k] = 0.125 * cze[k] * (1. + xi * cxi[k]) * (1. + et * cet[k]);
    }
    double J[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
    for (int k = 0; k < 8; ++k) {
      J[0][0] += d1[k] * X[k]; J[0][1] += d1[k] * Y[k]; J[0][2] += d1[k] * Z[k];
      J[1][0] += d2[k] * X[k]; J[1][1] += d2[k] * Y[k]; J[1][2] += d2[k] * Z[k];
      J[2][0] += d3[k] * X[k]; J[2][1] += d3[k] * Y[k]; J[2][2] += d3[k] * Z[k];
    }
    const double det = J[0][0] * (J[1][1] * J[2][2] - J[1][2] * J[2][1]) - J[0][1] * (J[1][0] * J[2][2] - J[1][2] * J[2][0]) +
                       J[0][2] * (J[1][0] * J[2][1] - J[1][1] * J[2][0]);
    if (!(det > 0.0)) return false;   // (JABN30 leaves BJ = 0 for det = 0 and HEXA32 flags det < 0: both are unusable elements)
    if (!jac_inverse(8, d1, d2, d3, X, Y, Z, I, detj)) return false;
  }
  for (int k = 0; k < sp.nn; ++k)
    for (int d = 0; d < 3; ++d) b[d][k] = I[d][0] * d1[k] + I[d][1] * d2[k] + I[d][2] * d3[k];
  return true;
}

__global__ void build_linsolid_ops_kernel(int nelt, const int* __restrict__ elem, const int* __restrict__ conn,
                                          const double* __restrict__ xyz, const double* __restrict__ emod,
                                          const double* __restrict__ rny, const LinSolidSpec* __restrict__ spec, int MT, int KT,
                                          int KTG, double* __restrict__ Sfrag, double* __restrict__ Gfrag,
                                          unsigned char* __restrict__ failed, double* __restrict__ aux)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nelt) return;
  const LinSolidSpec& sp = *spec;
  const int e = elem[i], nn = sp.nn;
  double* S = Sfrag + (size_t)i * MT * KT * 32;
  double* G = Gfrag + (size_t)i * 3 * KTG * 32;
  double X[8], Y[8], Z[8];
  for (int k = 0; k < nn; ++k) {
    const int n = conn[i * nn + k];
    X[k] = xyz[3 * n]; Y[k] = xyz[3 * n + 1]; Z[k] = xyz[3 * n + 2];
  }
  const double E = emod[e], nu = rny[e];
  aux[i * 2] = E; aux[i * 2 + 1] = nu;
  // isoMat3D (isoMatModule.f90:63-91)
  const double fac = E / ((1.0 + nu) * (1.0 - nu - nu));
  const double D = (1.0 - nu) * fac, D1 = nu * fac, D2 = (0.5 - nu) * fac;
  bool ok = true;
  double wsum = 0.0, wdet[8];
  if (sp.volume_average) {
    for (int g = 0; g < sp.neval && ok; ++g) {
      double b[3][8];
      ok = lin_gradients(sp, g, X, Y, Z, b, wdet[g]);
      wsum += wdet[g];
    }
  }
  for (int g = 0; g < sp.neval && ok; ++g) {
    double b[3][8], detj;
    ok = lin_gradients(sp, g, X, Y, Z, b, detj);
    if (!ok) break;
    for (int j = 0; j < nn; ++j) {
      const double bx = b[0][j], by = b[1][j], bz = b[2][j];
      // rows xx,yy,zz,xy, then (xz,yz) or (yz,zx); columns u,v,w of node j
      double db[6][3] = {{D * bx, D1 * by, D1 * bz}, {D1 * bx, D * by, D1 * bz}, {D1 * bx, D1 * by, D * bz},
                         {D2 * by, D2 * bx, 0.0},    {D2 * bz, 0.0, D2 * bx},    {0.0, D2 * bz, D2 * by}};
      if (sp.shear == 1) {   // B rows 5, 6 of cstetbmat / ipri6bmat: gamma_yz, gamma_zx
        db[4][0] = 0.0; db[4][1] = D2 * bz; db[4][2] = D2 * by;
        db[5][0] = D2 * bz; db[5][1] = 0.0; db[5][2] = D2 * bx;
      }
      for (int p = 0; p < nn; ++p) {
        const double w = sp.volume_average ? wdet[g] / wsum : sp.W[p][g];
        if (w == 0.0) continue;
        for (int c = 0; c < 6; ++c)
          for (int d = 0; d < 3; ++d) S[fragidx(p * 6 + c, 3 * j + d, KT)] += w * db[c][d];
        for (int d = 0; d < 3; ++d) G[fragidx(d * 8 + p, j, KTG)] += w * b[d][j];
      }
    }
  }
  if (!ok) {
    for (int k = 0; k < MT * KT * 32; ++k) S[k] = 0.0;
    for (int k = 0; k < 3 * KTG * 32; ++k) G[k] = 0.0;
  }
  failed[i] = ok ? 0 : 1;
}

// one warp per element; NP result points (<= 8) in one block of DMMA rows per derivative direction
template <int KT>
__global__ void __launch_bounds__(256, 2)
k2_linsolid_grad_vm_kernel(const double* __restrict__ U, size_t ldu, int nsteps, int nsteps_pad, int nn, int np, int estride,
                           const double* __restrict__ Gfrag, const double* __restrict__ aux, const int* __restrict__ edof,
                           const int* __restrict__ ptoff, const unsigned char* __restrict__ failed, int nelt,
                           double* __restrict__ vm, size_t ld_vm, double* __restrict__ env_max, double* __restrict__ env_min)
{
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, t4 = lane & 3;
  const int i = blockIdx.x * (blockDim.x >> 5) + warp;
  if (i >= nelt) return;
  double a[3][KT];
  const double* gf = Gfrag + (size_t)i * 3 * KT * 32 + lane;
#pragma unroll
  for (int m = 0; m < 3; ++m)
#pragma unroll
    for (int j = 0; j < KT; ++j) a[m][j] = __ldg(gf + (size_t)(m * KT + j) * 32);
  const double* urow[3][KT];
  const int* ed = edof + (size_t)i * estride;   // [node][3] rows of U
#pragma unroll
  for (int j = 0; j < KT; ++j) {
    const int node = 4 * j + t4;
#pragma unroll
    for (int c = 0; c < 3; ++c) urow[c][j] = U + (size_t)(node < nn ? __ldg(ed + 3 * node + c) : 0) * ldu + g;
  }
  const double E = __ldg(aux + (size_t)i * 2), nu = __ldg(aux + (size_t)i * 2 + 1);
  const double mu2 = E / (1.0 + nu);
  const bool bad = failed[i] != 0, live = g < np;
  const size_t pt0 = (size_t)ptoff[i];
  double emax = 0.0, emin = kHuge;
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
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int t = nt * 8 + 2 * t4 + q;
      // deviatoric von Mises from the displacement gradient H[c][d] = acc[d][c] (see k2_solid.cu)
      const double da = acc[0][0][q] - acc[1][1][q], db = acc[1][1][q] - acc[2][2][q], dc = acc[2][2][q] - acc[0][0][q];
      const double gxy = acc[1][0][q] + acc[0][1][q], gxz = acc[2][0][q] + acc[0][2][q], gyz = acc[2][1][q] + acc[1][2][q];
      const double dev = 0.5 * fma(da, da, fma(db, db, dc * dc));
      const double shr = fma(gxy, gxy, fma(gxz, gxz, gyz * gyz));
      double v = mu2 * sqrt_pos(fma(0.75, shr, dev));
      if (bad) v = kHuge;
      if (live && t < nsteps) {
        if (vm) vm[(size_t)t * ld_vm + pt0 + g] = v;
        emax = fmax(emax, v);
        emin = fmin(emin, v);
      }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
      for (int j = 0; j < KT; ++j) b[c][j] = bn[c][j];
  }
#pragma unroll
  for (int o = 1; o < 4; o <<= 1) {
    emax = fmax(emax, __shfl_xor_sync(0xffffffffu, emax, o));
    emin = fmin(emin, __shfl_xor_sync(0xffffffffu, emin, o));
  }
  if (live && t4 == 0 && nsteps > 0) {
    if (emax > env_max[pt0 + g]) env_max[pt0 + g] = emax;
    if (emin < env_min[pt0 + g]) env_min[pt0 + g] = emin;
  }
}

LinSolidSpec make_spec(int type, int stressForm)
{
  LinSolidSpec s;
  memset(&s, 0, sizeof(s));
  s.type = type;
  if (type == 45) {   // constant strain: one evaluation, copied to the four nodes
    s.nn = 4; s.neval = 1; s.shear = 1;
    for (int p = 0; p < 4; ++p) s.W[p][0] = 1.0;
  } else if (type == 46) {
    s.nn = 6; s.neval = 6; s.shear = 1;
    const int code = stressForm == 1 ? 3 : stressForm;   // elStressModule.f90:1757
    const double r3 = 1.0 / std::sqrt(3.0);
    double zeta[2], xi_d, xi_o;
    switch (code) {
      case 1: zeta[0] = -r3; zeta[1] = r3; xi_d = 2.0 / 3.0; xi_o = 1.0 / 6.0; break;
      case 2: zeta[0] = -r3; zeta[1] = r3; xi_d = 0.0; xi_o = 0.5; break;
      case 3: zeta[0] = zeta[1] = 0.0; xi_d = 0.0; xi_o = 0.5; break;
      default: zeta[0] = -1.0; zeta[1] = 1.0; xi_d = 1.0; xi_o = 0.0; break;
    }
    for (int zp = 0; zp < 2; ++zp)
      for (int xp = 0; xp < 3; ++xp) {
        for (int k = 0; k < 3; ++k) s.pt[3 * zp + xp][k] = k == xp ? xi_d : xi_o;
        s.pt[3 * zp + xp][3] = zeta[zp];
      }
    // in-plane extrapolation T (3x3) within each triangle, then ipri6extrapolH between the two triangles
    double T[3][3];
    for (int a = 0; a < 3; ++a)
      for (int b2 = 0; b2 < 3; ++b2)
        T[a][b2] = code == 1 ? (a == b2 ? 5.0 / 3.0 : -1.0 / 3.0) : (code == 2 || code == 3) ? (a == b2 ? -1.0 : 1.0) : (a == b2 ? 1.0 : 0.0);
    const double zm1 = 0.5 * (std::sqrt(3.0) - 1.0), zp1 = 0.5 * (std::sqrt(3.0) + 1.0);
    for (int a = 0; a < 3; ++a)
      for (int b2 = 0; b2 < 3; ++b2) {
        if (code == 1 || code == 2) {
          s.W[a][b2] = zp1 * T[a][b2]; s.W[a][3 + b2] = -zm1 * T[a][b2];
          s.W[3 + a][3 + b2] = zp1 * T[a][b2]; s.W[3 + a][b2] = -zm1 * T[a][b2];
        } else if (code == 3) {   // only the first triangle is evaluated; both faces get its extrapolation
          s.W[a][b2] = T[a][b2]; s.W[3 + a][b2] = T[a][b2];
        } else {
          s.W[a][b2] = T[a][b2]; s.W[3 + a][3 + b2] = T[a][b2];
        }
      }
    if (code == 3) s.neval = 3;
  } else {            // 44
    s.nn = 8; s.neval = 8; s.shear = 0;
    const double cxi[8] = {-1., 1., 1., -1., -1., 1., 1., -1.}, cet[8] = {-1., -1., 1., 1., -1., -1., 1., 1.},
                 cze[8] = {-1., -1., -1., -1., 1., 1., 1., 1.};
    const double abc = stressForm == 0 ? 1.0 : std::sqrt(1.0 / 3.0);   // LOP = 0 or NIP - 1 = 2
    for (int n = 0; n < 8; ++n) { s.pt[n][0] = abc * cxi[n]; s.pt[n][1] = abc * cet[n]; s.pt[n][2] = abc * cze[n]; }
    if (stressForm == 0) for (int n = 0; n < 8; ++n) s.W[n][n] = 1.0;
    else if (stressForm == 1) s.volume_average = 1;
    else {   // extrapolation with LINHEX at -sqrt(3) * node position (elStressModule.f90:1671-1684, hexa.f:1137-1151)
      const double lx[8] = {1., -1., -1., 1., 1., -1., -1., 1.}, ly[8] = {1., 1., -1., -1., 1., 1., -1., -1.},
                   lz[8] = {1., 1., 1., 1., -1., -1., -1., -1.};
      const double r3 = std::sqrt(3.0);
      for (int n = 0; n < 8; ++n)
        for (int g = 0; g < 8; ++g)
          s.W[n][g] = 0.125 * (1.0 - cxi[n] * r3 * lx[g]) * (1.0 - cet[n] * r3 * ly[g]) * (1.0 - cze[n] * r3 * lz[g]);
    }
  }
  return s;
}

}  // namespace

int build_linsolid_operators(fsr_part* p, const fsr_sam* sam, const fsr_elmdata* elm)
{
  cudaStream_t s = p->stream;
  const struct { int fam, type, nn; } kinds[3] = {{FAM_HEX8, 44, 8}, {FAM_TET4, 45, 4}, {FAM_WEDG6, 46, 6}};
  for (const auto& k : kinds) {
    FamilyData& f = p->fam[k.fam];
    f.nenod = k.nn; f.nndof = 3; f.nstrp = k.nn; f.ncmp = 6; f.naux = 2;
    f.MT = (6 * k.nn + 7) / 8; f.KT = (3 * k.nn + 3) / 4;
    const int KTG = (k.nn + 3) / 4;
    std::vector<int> elem, conn, edof, ptoff;
    for (int e : elements_of_type(p, sam, elm, k.type)) {
      const int ip0 = sam->mpmnpc[e] - 1, nn = sam->mpmnpc[e + 1] - sam->mpmnpc[e];
      if (nn != k.nn) { set_error("element %d of type %d has %d nodes, expected %d", e + 1, k.type, nn, k.nn); return FSR_ERR_ARG; }
      elem.push_back(e);
      ptoff.push_back(p->ptoff_host[e]);
      const size_t base = edof.size();
      edof.resize(base + (size_t)f.KT * 4, 0);
      for (int q = 0; q < k.nn; ++q) {
        const int n = sam->mmnpc[ip0 + q] - 1;
        if (n < 0 || n >= sam->nnod) { set_error("element %d: node index out of range", e + 1); return FSR_ERR_ARG; }
        conn.push_back(n);
        const int js = sam->madof[n] - 1, nd = sam->madof[n + 1] - sam->madof[n];
        if (nd < 3) { set_error("element %d: node %d has %d DOFs, solid needs 3", e + 1, n + 1, nd); return FSR_ERR_ARG; }
        for (int d = 0; d < 3; ++d) edof[base + (size_t)q * 3 + d] = js + d;
      }
    }
    f.nelt = (int)elem.size();
    if (f.nelt == 0) continue;
    const LinSolidSpec h = make_spec(k.type, p->stressForm);
    LinSolidSpec* d_spec = nullptr;
    int* d_conn = nullptr;
    FSR_CUDA(cudaMalloc(&d_spec, sizeof(h)));
    FSR_CUDA(cudaMemcpyAsync(d_spec, &h, sizeof(h), cudaMemcpyHostToDevice, s));
    FSR_CUDA(cudaMalloc(&f.elem, sizeof(int) * elem.size()));
    FSR_CUDA(cudaMalloc(&f.edof, sizeof(int) * edof.size()));
    FSR_CUDA(cudaMalloc(&f.ptoff, sizeof(int) * ptoff.size()));
    FSR_CUDA(cudaMalloc(&f.failed, f.nelt));
    FSR_CUDA(cudaMalloc(&f.Sfrag, sizeof(double) * (size_t)f.nelt * f.MT * f.KT * 32));
    FSR_CUDA(cudaMalloc(&f.Gfrag, sizeof(double) * (size_t)f.nelt * 3 * KTG * 32));
    FSR_CUDA(cudaMalloc(&f.aux, sizeof(double) * (size_t)f.nelt * f.naux));
    FSR_CUDA(cudaMalloc(&d_conn, sizeof(int) * conn.size()));
    FSR_CUDA(cudaMemcpyAsync(f.elem, elem.data(), sizeof(int) * elem.size(), cudaMemcpyHostToDevice, s));
    FSR_CUDA(cudaMemcpyAsync(f.edof, edof.data(), sizeof(int) * edof.size(), cudaMemcpyHostToDevice, s));
    FSR_CUDA(cudaMemcpyAsync(f.ptoff, ptoff.data(), sizeof(int) * ptoff.size(), cudaMemcpyHostToDevice, s));
    FSR_CUDA(cudaMemcpyAsync(d_conn, conn.data(), sizeof(int) * conn.size(), cudaMemcpyHostToDevice, s));
    FSR_CUDA(cudaMemsetAsync(f.Sfrag, 0, sizeof(double) * (size_t)f.nelt * f.MT * f.KT * 32, s));
    FSR_CUDA(cudaMemsetAsync(f.Gfrag, 0, sizeof(double) * (size_t)f.nelt * 3 * KTG * 32, s));
    build_linsolid_ops_kernel<<<(f.nelt + 63) / 64, 64, 0, s>>>(f.nelt, f.elem, d_conn, p->xyz, p->emod, p->rny, d_spec, f.MT, f.KT,
                                                               KTG, f.Sfrag, f.Gfrag, f.failed, f.aux);
    FSR_LAUNCH_CHECK();
    FSR_CUDA(cudaStreamSynchronize(s));
    cudaFree(d_conn);
    cudaFree(d_spec);
  }
  return FSR_OK;
}

int launch_k2_linsolid_vm(fsr_part* p, int nsteps, int nsteps_pad, double* vm_dev, size_t ld_vm, cudaStream_t s)
{
  const int warps = 8;
  for (int fam : {FAM_HEX8, FAM_TET4, FAM_WEDG6}) {
    FamilyData& f = p->fam[fam];
    if (f.nelt == 0) continue;
    const int grid = (f.nelt + warps - 1) / warps;
    if (f.nenod <= 4)
      k2_linsolid_grad_vm_kernel<1><<<grid, warps * 32, 0, s>>>(p->U, (size_t)p->step_tile, nsteps, nsteps_pad, f.nenod, f.nstrp,
                                                               f.KT * 4, f.Gfrag, f.aux, f.edof, f.ptoff, f.failed, f.nelt, vm_dev, ld_vm,
                                                               p->env_max, p->env_min);
    else
      k2_linsolid_grad_vm_kernel<2><<<grid, warps * 32, 0, s>>>(p->U, (size_t)p->step_tile, nsteps, nsteps_pad, f.nenod, f.nstrp,
                                                               f.KT * 4, f.Gfrag, f.aux, f.edof, f.ptoff, f.failed, f.nelt, vm_dev, ld_vm,
                                                               p->env_max, p->env_min);
    FSR_LAUNCH_CHECK();
  }
  return FSR_OK;
}

}  // namespace fsr
