// io_rdb.cu -- the stress results database (.frs) of fedem_stress, written straight from the GPU.
//
// Reference: writeStressHeader / writeElementsHeader / writeBeamHeader / writeShellHeader / writeSolidHeader
// (src/vpmStress/saveStressModule.f90:120-247,625-752,764-966,978-1186,1198-1347) build the text header in
// three scratch files (variables, item groups, data blocks; src/vpmCommon/rdbModule.f90:191-251,418-463),
// then calcStresses (src/vpmStress/stressRoutines.f90:169-331) writes, per time step and per element in
// SAM order: the stress resultants SR(6,nenod) when -SR is on (writeStressDB, saveStressModule.f90:1527-1567)
// and, per result point, [stress tensor][strain tensor][the selected ones of vmStress, maxP, minP, maxShear,
// vmStrain, maxP, minP, maxShear] (writeStrMeasureDB :1579-1633), each value as float unless -double.
//
// Here the record of a whole tile of time steps is formed on the device: K1 expands the tile, the record
// kernels evaluate every element result point for every step of the tile and drop the selected values at
// their slot of the step record (slot-major, step fastest: coalesced reads of U and coalesced writes), one
// tiled transpose turns that into step-major float/double records, and the host only adds the 12-byte
// step key (int32 step number + float64 time, writeTimeStepDB rdbModule.f90:669-736) in front of each.
#include <algorithm>
#include <cerrno>
#include <chrono>
#include <condition_variable>
#include <cstdarg>
#include <cstdlib>
#include <ctime>
#include <deque>
#include <mutex>
#include <string>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/statvfs.h>
#include <sys/uio.h>
#include <thread>
#include <unistd.h>
#include <vector>

#include "common.cuh"
#include "invariants.cuh"
#include "io_tagged.cuh"

namespace fsr {

int ensure_batch_buffers(fsr_part* p, bool need_vm_tile);

struct RecLayout {
  int sr, stress, strain;   // 1 = written
  int mask;                 // bit j = resMat row j+1 written
  int nsel;                 // popcount(mask)
  int def;                  // 0 = no nodal output, 1 = deformational displacements, 3 = + total displacements
};

__device__ __forceinline__ size_t frag_at2(int row, int col, int KT)
{
  return ((size_t)((row >> 3) * KT + (col >> 2)) << 5) + ((row & 7) << 2) + (col & 3);
}

// The record kernel of every element family with stress points: one warp per element, a tile of 8 time steps at a time.
//   (1) sigma rows (thick shells: also the strain rows of Efrag) = S_e . v_e with DMMA.8x8x4; the operator fragments are
//       streamed from L1/L2 (one coalesced 256-byte load per fragment, the same bytes for all step tiles of the element),
//       the element's DOF rows of U[dof][t] are read as 64-byte segments like in the von Mises kernels;
//   (2) the accumulators go to shared memory as [row][step];
//   (3) the lanes take (step, point) pairs with the point running fastest, form the strain tensor, von Mises, the principal
//       values (FFa::cubicSolve branches, invariants.cuh) and max shear of whatever is selected and put the float/double
//       values at their place of the element's piece of the step record -- in shared memory;
//   (4) the eight record pieces (contiguous in the STEP-MAJOR records) go out with fully coalesced stores: no slot-major
//       staging in HBM, no transpose kernel.
// layout: 0 = thin shells (operator row = comp*8 + point), 1 = solids (row = point*ncmp + comp),
// 2 = thick shells (rows as solids, strain from its own operator Efrag, zero stress resultants)
constexpr int kRecWarps = 4;
constexpr int kRecLds = 9;   // shared-memory row stride in doubles (8 steps + 1: spreads the rows over the banks)

// von Mises / principal values of one result point with the component count known at compile time (everything stays in
// registers).  2-D: FFaTensorTransforms.C:33-36 and the quadratic branch of FFa::cubicSolve (FFaMath.C:117-133: P > 0 two
// roots, P > -1e-64 a double root, else no result -> the zeros the caller put in, like the reference's stale values);
// 3-D: the shared routines of invariants.cuh (trigonometric cubic with the reference's case analysis).
template <int NCMP>
__device__ __forceinline__ void rec_invariants(const double (&S)[NCMP], bool want_vm, bool want_p, double& vm, double& pmax, double& pmin)
{
  vm = pmax = pmin = 0.0;
  if (NCMP == 3) {
    if (want_vm) vm = sqrt_pos(fma(3.0 * S[2], S[2], fma(-S[0], S[1], fma(S[1], S[1], S[0] * S[0]))));
    if (want_p) {
      const double Cq = -(S[0] + S[1]), Dq = S[0] * S[1] - S[2] * S[2];
      const double Pq = Cq * Cq - 4.0 * Dq;
      if (Pq > 0.0) { const double Q = sqrt_pos(Pq); pmax = 0.5 * (Q - Cq); pmin = 0.5 * (-Cq - Q); }
      else if (Pq > -1.0e-64) pmax = pmin = -0.5 * Cq;
    }
  } else {
    if (want_vm) vm = von_mises(NCMP, S);
    if (want_p) { double P[3] = {0.0, 0.0, 0.0}; principal_values(NCMP, S, P); pmax = P[0]; pmin = P[2]; }
  }
}

template <int KT, int LAYOUT, class OUT_T>
__global__ void __launch_bounds__(kRecWarps * 32)
record_points_dmma_kernel(const double* __restrict__ U, size_t ldu, int nt, const double* __restrict__ Sfrag,
                          const double* __restrict__ Efrag, const int* __restrict__ edof, const long long* __restrict__ roff,
                          const unsigned char* __restrict__ failed, const double* __restrict__ aux, int naux, int nelt, int nstrp,
                          int MT, int nenod, RecLayout L, OUT_T* __restrict__ out, size_t ld_out)
{
  constexpr int NCMP = LAYOUT == 0 ? 3 : 6;
  extern __shared__ double rec_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, t4 = lane & 3;
  const int i = blockIdx.x * kRecWarps + warp;
  if (i >= nelt) return;   // whole warp
  const long long base = roff[i];
  if (base < 0) return;
  const int nrow = MT * 8;
  const int srsize = L.sr && LAYOUT != 1 ? 6 * nenod : 0;
  const int ptsize = (L.stress ? NCMP : 0) + (L.strain ? NCMP : 0) + L.nsel;
  const int nval = srsize + nstrp * ptsize;          // values of this element in one step record
  const int nval_pad = (nval + 1) & ~1;              // keeps the double rows 8-byte aligned behind float pieces
  // per warp: [sig rows][eps rows (thick shells)] as doubles, then 8 record pieces of nval values
  const size_t warp_doubles = (size_t)(LAYOUT == 2 ? 2 : 1) * nrow * kRecLds + ((size_t)8 * nval_pad * sizeof(OUT_T) + 7) / 8;
  double* sig_s = rec_smem + (size_t)warp * warp_doubles;
  double* eps_s = sig_s + (size_t)nrow * kRecLds;   // thick shells only
  OUT_T* rec_s = reinterpret_cast<OUT_T*>(sig_s + (size_t)(LAYOUT == 2 ? 2 : 1) * nrow * kRecLds);
  const bool bad = failed[i] != 0;
  const double* S = Sfrag + (size_t)i * MT * KT * 32 + lane;
  const double* Es = LAYOUT == 2 ? Efrag + (size_t)i * MT * KT * 32 + lane : nullptr;
  const double E = aux[(size_t)i * naux], nu = aux[(size_t)i * naux + 1];
  const double th = LAYOUT == 0 ? aux[(size_t)i * naux + 2] : 0.0;
  const double iE = 1.0 / E, g1 = (1.0 + nu) * iE;   // isoMat2Dinv / isoMat3Dinv (isoMatModule.f90:41-57,95-120); tensorial shear
  const double srn = 0.5 * th, srm = 0.5 * th * th / 6.0;
  const bool want_s = (L.mask & 0x0f) != 0, want_e = (L.mask & 0xf0) != 0 || L.strain;
  // B operand rows of this lane: element DOF 4*k + t4 (padding columns point at row 0, their operator entries are zero)
  const double* up[KT];
#pragma unroll
  for (int k = 0; k < KT; ++k) up[k] = U + (size_t)__ldg(edof + (size_t)i * KT * 4 + k * 4 + t4) * ldu + g;
  const int npair = 8 * nstrp;

  for (int t0 = 0; t0 < nt; t0 += 8) {
    if (!bad) {
      double b[KT];
#pragma unroll
      for (int k = 0; k < KT; ++k) b[k] = up[k][t0];   // U rows carry slack up to the padded tile
      // three m-tiles at a time (three independent accumulator chains); the fragment loads of a group carry no branches
      // (the last group repeats the last m-tile instead of running short), so they are all in flight before the first DMMA
      for (int m0 = 0; m0 < MT; m0 += 3) {
        const double* Sm[3];
        const double* Em[3];
#pragma unroll
        for (int mm = 0; mm < 3; ++mm) {
          const int mi = min(m0 + mm, MT - 1);
          Sm[mm] = S + (size_t)mi * KT * 32;
          Em[mm] = LAYOUT == 2 ? Es + (size_t)mi * KT * 32 : nullptr;
        }
        double c[3][2] = {{0, 0}, {0, 0}, {0, 0}};
#pragma unroll
        for (int k = 0; k < KT; ++k)
#pragma unroll
          for (int mm = 0; mm < 3; ++mm) dmma884(c[mm][0], c[mm][1], __ldg(Sm[mm] + k * 32), b[k]);
#pragma unroll
        for (int mm = 0; mm < 3; ++mm)
          if (m0 + mm < MT) {
            double* q = sig_s + ((m0 + mm) * 8 + g) * kRecLds + 2 * t4;
            q[0] = c[mm][0]; q[1] = c[mm][1];
          }
        if (LAYOUT == 2) {
          double d[3][2] = {{0, 0}, {0, 0}, {0, 0}};
#pragma unroll
          for (int k = 0; k < KT; ++k)
#pragma unroll
            for (int mm = 0; mm < 3; ++mm) dmma884(d[mm][0], d[mm][1], __ldg(Em[mm] + k * 32), b[k]);
#pragma unroll
          for (int mm = 0; mm < 3; ++mm)
            if (m0 + mm < MT) {
              double* q = eps_s + ((m0 + mm) * 8 + g) * kRecLds + 2 * t4;
              q[0] = d[mm][0]; q[1] = d[mm][1];
            }
        }
      }
    }
    __syncwarp();
    for (int idx = lane; idx < npair; idx += 32) {
      const int st = idx / nstrp, pnt = idx - st * nstrp;
      const int t = t0 + st;
      if (t >= nt) continue;
      OUT_T* o = rec_s + (size_t)st * nval_pad + srsize + pnt * ptsize;
      OUT_T* so = rec_s + (size_t)st * nval_pad + 6 * pnt;
      if (bad) {   // stressRoutines.f90:237-241,264-268: hugeVal for everything that is written
        for (int k = 0; k < ptsize; ++k) o[k] = (OUT_T)kHuge;
        if (srsize && pnt < nenod) for (int k = 0; k < 6; ++k) so[k] = (OUT_T)kHuge;
        continue;
      }
      double sig[NCMP], eps[NCMP];
#pragma unroll
      for (int c = 0; c < NCMP; ++c) sig[c] = sig_s[(LAYOUT == 0 ? c * 8 + pnt : pnt * NCMP + c) * kRecLds + st];
      if (LAYOUT == 2) {
#pragma unroll
        for (int c = 0; c < NCMP; ++c) eps[c] = eps_s[(pnt * NCMP + c) * kRecLds + st];
      } else if (NCMP == 3) {
        eps[0] = (sig[0] - nu * sig[1]) * iE;
        eps[1] = (sig[1] - nu * sig[0]) * iE;
        eps[2] = g1 * sig[2];
      } else {
        eps[0] = (sig[0] - nu * (sig[1] + sig[2])) * iE;
        eps[1] = (sig[1] - nu * (sig[0] + sig[2])) * iE;
        eps[2] = (sig[2] - nu * (sig[0] + sig[1])) * iE;
#pragma unroll
        for (int c = 3; c < NCMP; ++c) eps[c] = g1 * sig[c];
      }
      int k = 0;
      if (L.stress) {
#pragma unroll
        for (int c = 0; c < NCMP; ++c) o[k + c] = (OUT_T)sig[c];
        k += NCMP;
      }
      if (L.strain) {
#pragma unroll
        for (int c = 0; c < NCMP; ++c) o[k + c] = (OUT_T)eps[c];
        k += NCMP;
      }
      if (L.nsel) {
        double r[8];
        if (want_s) rec_invariants<NCMP>(sig, (L.mask & 0x01) != 0, (L.mask & 0x0e) != 0, r[0], r[1], r[2]);
        if (L.mask & 0xf0) rec_invariants<NCMP>(eps, (L.mask & 0x10) != 0, (L.mask & 0xe0) != 0, r[4], r[5], r[6]);
        r[3] = 0.5 * (r[1] - r[2]); r[7] = 0.5 * (r[5] - r[6]);   // maxshearvalue_ (FFaTensorTransforms.C:295)
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (L.mask & (1 << j)) o[k++] = (OUT_T)r[j];
      }
      if (srsize && pnt < nenod) {
        if (LAYOUT == 2) {   // thick shells: SR = 0 (STR31 / STR32, elStressModule.f90:1174-1176, 1296-1298)
#pragma unroll
          for (int c = 0; c < 6; ++c) so[c] = (OUT_T)0.0;
        } else if (LAYOUT == 0) {   // thin shells: from the top and bottom stresses of the node (STR22a :826-831, STR23 :976-979)
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            const double bot = sig_s[(c * 8 + nenod + pnt) * kRecLds + st];
            so[c] = (OUT_T)((sig[c] + bot) * srn);
            so[3 + c] = (OUT_T)((sig[c] - bot) * srm);
          }
        }
      }
    }
    __syncwarp();
    // the element's piece of each of the 8 step records: contiguous, coalesced
    for (int st = 0; st < 8 && t0 + st < nt; ++st) {
      OUT_T* dst = out + (size_t)(t0 + st) * ld_out + base;
      const OUT_T* src = rec_s + (size_t)st * nval_pad;
      for (int j = lane; j < nval; j += 32) dst[j] = src[j];
    }
    __syncwarp();
  }
  (void)want_e;
}

struct BeamOp12 { double S[12][12]; };

// beam section forces SF(6,2) (STR11, elStressModule.f90:402-515): 12 values per beam and step, written at their record
// place; x = (beam, row) with the row fastest (contiguous record bytes), y = step
template <class OUT_T>
__global__ void record_beams_kernel(const double* __restrict__ U, size_t ldu, int nt, const BeamOp12* __restrict__ ops,
                                    const int* __restrict__ edof, const long long* __restrict__ roff,
                                    const unsigned char* __restrict__ failed, int nelt, OUT_T* __restrict__ out, size_t ld_out)
{
  const long long ir = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int t = blockIdx.y * blockDim.y + threadIdx.y;
  if (t >= nt || ir >= (long long)nelt * 12) return;
  const int i = (int)(ir / 12), r = (int)(ir % 12);
  if (roff[i] < 0) return;
  double acc = 0.0;
  for (int c = 0; c < 12; ++c) acc += ops[i].S[r][c] * U[(size_t)edof[i * 12 + c] * ldu + t];
  out[(size_t)t * ld_out + roff[i] + r] = (OUT_T)(failed[i] ? kHuge : acc);
}

// ---- rotation utilities of src/vpmUtilities/rotationModule.f90 (vec_to_quat :393-428, quat_to_mat :478-497,
// mat_to_quat :441-470, quat_to_vec :505-533), matrices column-major a[i + 3*j] like the Fortran arrays
__host__ __device__ inline void rot_vec_to_mat(const double* v, double* R)
{
  const double eps = 0.0005;
  const double thh = 0.5 * sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
  const double sthh = sin(thh), cthh = cos(thh);
  double q[4];
  bool ok = true;
  if (thh > 1.0e6 && fabs(1.0 - cthh * cthh - sthh * sthh) > 0.00001) { q[0] = 1.0; q[1] = q[2] = q[3] = 0.0; ok = false; }
  if (ok) {
    double fac;
    if (thh < eps) { const double f1 = thh / eps; fac = f1 * sin(eps) / eps + 1.0 - f1; }
    else fac = sthh / thh;
    q[0] = cthh; q[1] = v[0] * fac * 0.5; q[2] = v[1] * fac * 0.5; q[3] = v[2] * fac * 0.5;
    const double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    for (int i = 0; i < 4; ++i) q[i] /= n;
  }
  const double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  for (int i = 0; i < 4; ++i) q[i] /= n;
  R[0] = 2.0 * (q[1] * q[1] + q[0] * q[0]) - 1.0;
  R[4] = 2.0 * (q[2] * q[2] + q[0] * q[0]) - 1.0;
  R[8] = 2.0 * (q[3] * q[3] + q[0] * q[0]) - 1.0;
  R[3] = 2.0 * (q[1] * q[2] - q[3] * q[0]);
  R[6] = 2.0 * (q[1] * q[3] + q[2] * q[0]);
  R[7] = 2.0 * (q[2] * q[3] - q[1] * q[0]);
  R[1] = 2.0 * (q[2] * q[1] + q[3] * q[0]);
  R[2] = 2.0 * (q[3] * q[1] - q[2] * q[0]);
  R[5] = 2.0 * (q[3] * q[2] + q[1] * q[0]);
}

__host__ __device__ inline void rot_mat_to_vec(const double* R, double* v)
{
  const double eps = 0.0005;
  double q[4];
  const double trace = R[0] + R[4] + R[8];
  int imax = 0;
  if (R[4] > R[4 * imax]) imax = 1;
  if (R[8] > R[4 * imax]) imax = 2;
  if (trace > R[4 * imax]) {
    q[0] = sqrt(1.0 + trace) * 0.5;
    q[1] = (R[2 + 3 * 1] - R[1 + 3 * 2]) / (4.0 * q[0]);
    q[2] = (R[0 + 3 * 2] - R[2 + 3 * 0]) / (4.0 * q[0]);
    q[3] = (R[1 + 3 * 0] - R[0 + 3 * 1]) / (4.0 * q[0]);
  } else {
    const int i = imax, j = (imax + 1) % 3, k = (imax + 2) % 3;
    q[i + 1] = sqrt(R[i + 3 * i] * 0.5 + (1.0 - trace) * 0.25);
    q[0] = (R[k + 3 * j] - R[j + 3 * k]) / (4.0 * q[i + 1]);
    q[j + 1] = (R[j + 3 * i] + R[i + 3 * j]) / (4.0 * q[i + 1]);
    q[k + 1] = (R[k + 3 * i] + R[i + 3 * k]) / (4.0 * q[i + 1]);
  }
  const double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  for (int i = 0; i < 4; ++i) q[i] /= n;
  const double cthh = q[0], sthh = sqrt(q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  const double thh = sthh < 0.7 ? asin(sthh) : acos(cthh);
  double fac;
  if (thh < eps) { const double f1 = thh / eps; fac = f1 * eps / sin(eps) + 1.0 - f1; }
  else if (sthh >= 1.0) fac = thh;
  else fac = thh / sthh;
  for (int i = 0; i < 3; ++i) v[i] = q[i + 1] * fac * 2.0;
}

// calcTotalNodalDisplacement (src/vpmStress/displacementModule.f90:1694-1745): x0 = nodal coordinates, u = the
// node's 3 or 6 deformational displacements, T = current and T0 = initial 3x4 position matrix of the part
__host__ __device__ inline void total_nodal_displacement(const double* x0, const double* u, int nd, const double* T,
                                                         const double* T0, double* utot)
{
  for (int i = 0; i < 3; ++i) {
    const double a = T[i] * (x0[0] + u[0]) + T[i + 3] * (x0[1] + u[1]) + T[i + 6] * (x0[2] + u[2]) + T[i + 9];
    const double b = T0[i] * x0[0] + T0[i + 3] * x0[1] + T0[i + 6] * x0[2] + T0[i + 9];
    utot[i] = a - b;
  }
  if (nd < 6) return;
  double dR[9], A[9], M[9];
  rot_vec_to_mat(u + 3, dR);
  for (int i = 0; i < 3; ++i)       // A = dR . T(:,1:3)
    for (int j = 0; j < 3; ++j) A[i + 3 * j] = dR[i] * T[3 * j] + dR[i + 3] * T[1 + 3 * j] + dR[i + 6] * T[2 + 3 * j];
  for (int i = 0; i < 3; ++i)       // deltaRot(T0, A) = mat_to_vec(A . T0^T)
    for (int j = 0; j < 3; ++j) M[i + 3 * j] = A[i] * T0[j] + A[i + 3] * T0[j + 3] + A[i + 6] * T0[j + 6];
  rot_mat_to_vec(M, utot + 3);
}

// writeDisplacementDB (saveStressModule.f90:1437-1515): per node the deformational displacements sv(j:k) and,
// with -deformation in the current reference (iDef = 3), the total displacements.  One thread per (node, step).
__global__ void record_nodes_kernel(const double* __restrict__ U, size_t ldu, int nt, int nnod, const int* __restrict__ madof,
                                    const long long* __restrict__ nslot, const double* __restrict__ xyz,
                                    const double* __restrict__ supTr, const double* __restrict__ supTr0, int total,
                                    double* __restrict__ rec, size_t ldt)
{
  const int t = blockIdx.y * blockDim.x + threadIdx.x;
  const int n = blockIdx.x * blockDim.y + threadIdx.y;
  if (t >= nt || n >= nnod) return;
  const long long base = nslot[n];
  if (base < 0) return;
  const int j = madof[n] - 1, nd = madof[n + 1] - madof[n] > 5 ? 6 : 3;
  double u[6] = {0, 0, 0, 0, 0, 0}, ut[6];
  for (int d = 0; d < nd; ++d) { u[d] = U[(size_t)(j + d) * ldu + t]; rec[(size_t)(base + d) * ldt + t] = u[d]; }
  if (!total) return;
  double T[12], T0[12];
  for (int i = 0; i < 12; ++i) { T[i] = supTr[(size_t)t * 12 + i]; T0[i] = supTr0[i]; }
  total_nodal_displacement(xyz + 3 * (size_t)n, u, nd, T, T0, ut);
  for (int d = 0; d < nd; ++d) rec[(size_t)(base + nd + d) * ldt + t] = ut[d];
}

// rec[slot][t] (double) -> out[t * ld_out + slot] as float or double: the nodal part of the step records (the
// displacement rows of U are step-fastest, the record is slot-fastest)
template <class T>
__global__ void record_transpose_kernel(const double* __restrict__ rec, size_t ldt, long long nslot, int nt, T* __restrict__ out,
                                        size_t ld_out)
{
  __shared__ double tile[32][33];
  const long long s0 = (long long)blockIdx.x * 32;
  const int t0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const long long s = s0 + j;
    const int t = t0 + threadIdx.x;
    if (s < nslot && t < nt) tile[j][threadIdx.x] = rec[(size_t)s * ldt + t];
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int t = t0 + j;
    const long long s = s0 + threadIdx.x;
    if (s < nslot && t < nt) out[(size_t)t * ld_out + s] = (T)tile[threadIdx.x][j];
  }
}

static void appendf(std::string& s, const char* fmt, ...)
{
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  s += buf;
}

}  // namespace fsr

using namespace fsr;

// Two record buffers per device, two pinned ones on the host: while tile n is computed (stream of the part), tile n-1
// crosses PCIe (copy stream) and tile n-2 goes to the file (writer thread + helpers, pwritev of whole batches of step
// records).  With a group of element blocks (several GPUs) every device fills the slots of ITS elements -- a contiguous
// range of the step record, blocks being contiguous element ranges -- and a strided device-to-host copy drops them at
// their place of the full records in the shared pinned buffer: the merge costs nothing.
struct RdbJob { int buf = 0, nt = 0; std::vector<char> keys; };   // keys: nt x (int32 step number, float64 time)

struct RdbDev {
  fsr_part* part = nullptr;
  long long slot0 = 0, nslot = 0;      // this device's slot range of the step record
  long long* roff[fsr::FAM_COUNT] = {};// device: record slot (relative to slot0) of each family element, -1 = not written
  void* out[2] = {nullptr, nullptr};   // [tile][nslot] float/double
  double* dQ = nullptr;                // device Q tile
  cudaEvent_t ev[2][5] = {};           // per buffer: compute start / done (part stream), copy start / done (copy stream),
                                       // [4] = expansion done (part stream; = start for the later record tiles of a chunk)
  cudaEvent_t ev_q[2] = {};            // the H2D copy out of Qpin[k] has finished
  cudaStream_t copy_stream = nullptr;
};

struct fsr_rdb {
  std::vector<RdbDev> devs;
  fsr_part* part = nullptr;            // devs[0].part
  FILE* f = nullptr;
  int fd = -1;
  int mfd = -1;                        // second descriptor (O_RDWR) for the mapped writes, -1 = pwritev path
  long long data_pos = 0;              // file offset of the next step record
  std::string header, path;
  RecLayout L{};
  int dbl = 0;
  long long nslot = 0;                 // values per step record
  long long nslot_nodes = 0;           // the leading nodal values of the record (writeDisplacementDB; single device only)
  double* rec = nullptr;               // [nslot_nodes][tile] slot-major staging of the nodal values
  void* host[2] = {nullptr, nullptr};  // pinned step records
  double* Qpin[2] = {nullptr, nullptr};// pinned staging of the caller's Q tile
  int ldq_cap = 0;
  int tile = 0, next_buf = 0;          // record tile: steps per pinned buffer
  int ktile = 0, next_q = 0;           // expansion chunk: steps per K1 launch (>= tile; several record tiles read one U)
  long long steps_written = 0;
  long long* node_slot = nullptr;      // device [nnod] record slot of each node's displacements (-1 = none)
  int* madof = nullptr;                // device [nnod+1]
  double* supTr0 = nullptr;            // device [12] initial part position (total displacements)
  double* supTr = nullptr;             // device [tile][12]
  double* supPin[2] = {nullptr, nullptr};
  // writer thread
  std::thread writer;
  std::mutex mtx;
  std::condition_variable cv;
  std::deque<RdbJob> jobs;
  bool busy[2] = {false, false};       // buffer handed to the writer and not yet on file
  bool quit = false;
  std::string werr;                    // first error of the writer thread
  int nwriters = 1;
  // accounting (fsr_rdb_flush)
  double ms_compute = 0.0, ms_copy = 0.0, ms_disk = 0.0, ms_k1 = 0.0;
  long long bytes_written = 0, tiles = 0;

  void writer_main();
  void drain()
  {
    std::unique_lock<std::mutex> lk(mtx);
    cv.wait(lk, [&] { return jobs.empty() && !busy[0] && !busy[1]; });
  }
  ~fsr_rdb()
  {
    if (writer.joinable()) {
      { std::lock_guard<std::mutex> lk(mtx); quit = true; }
      cv.notify_all();
      writer.join();
    }
    if (mfd >= 0) close(mfd);
    if (f) fclose(f);
    for (RdbDev& d : devs) {
      cudaSetDevice(d.part->device);
      for (auto& r : d.roff) cudaFree(r);
      cudaFree(d.dQ);
      for (int b = 0; b < 2; ++b) {
        cudaFree(d.out[b]);
        for (auto& e : d.ev[b]) if (e) cudaEventDestroy(e);
        if (d.ev_q[b]) cudaEventDestroy(d.ev_q[b]);
      }
      if (d.copy_stream) cudaStreamDestroy(d.copy_stream);
    }
    if (part) cudaSetDevice(part->device);
    cudaFree(rec); cudaFree(node_slot); cudaFree(madof); cudaFree(supTr0); cudaFree(supTr);
    for (int b = 0; b < 2; ++b) {
      if (host[b]) cudaFreeHost(host[b]);
      if (Qpin[b]) cudaFreeHost(Qpin[b]);
      if (supPin[b]) cudaFreeHost(supPin[b]);
    }
  }
};

// steps [t0, t1) of a pinned buffer -> file at `pos`: 12-byte key + payload per step (writeTimeStepDB,
// rdbModule.f90:669-736), up to 512 steps per pwritev
static std::string write_step_range(int fd, long long pos, const char* keys, const char* payload, size_t rec_bytes, int t0, int t1,
                                    const std::string& path)
{
  std::vector<iovec> iov;
  for (int t = t0; t < t1;) {
    const int n = std::min(512, t1 - t);
    iov.resize(2 * (size_t)n);
    size_t want = 0;
    for (int k = 0; k < n; ++k) {
      iov[2 * k] = {(void*)(keys + 12 * (size_t)(t + k)), 12};
      iov[2 * k + 1] = {(void*)(payload + rec_bytes * (size_t)(t + k)), rec_bytes};
      want += 12 + rec_bytes;
    }
    size_t first = 0;
    long long off = pos + (long long)(t - t0) * (long long)(12 + rec_bytes);
    while (want > 0) {   // a short write continues where it stopped
      const ssize_t w = pwritev(fd, iov.data() + first, (int)std::min<size_t>(iov.size() - first, 1024), off);
      if (w < 0) { if (errno == EINTR) continue; return path + ": write error: " + strerror(errno); }
      want -= (size_t)w;
      off += w;
      size_t left = (size_t)w;
      while (left > 0 && first < iov.size()) {
        if (left >= iov[first].iov_len) { left -= iov[first].iov_len; ++first; }
        else { iov[first].iov_base = (char*)iov[first].iov_base + left; iov[first].iov_len -= left; left = 0; }
      }
    }
    t += n;
  }
  return std::string();
}

// The same records through a shared mapping of the file: the file is extended to the end of the tile, the window is mapped and
// `nthreads` threads copy equal BYTE ranges of the record stream (12-byte key + payload per step) into it.  write() / pwritev()
// hold the inode lock while they copy into the page cache, so concurrent writers to one file do not add up; page faults on a
// shared mapping do (per-VMA locks), and the copy is what the file stage consists of.  Returns "!" when the mapping cannot be
// set up (the caller falls back to pwritev), an error text, or nothing.
static std::string write_steps_mapped(int mfd, long long pos, const char* keys, const char* payload, size_t rec_bytes, int nt,
                                      int nthreads, const std::string& path)
{
  const long long step_bytes = 12 + (long long)rec_bytes, total = step_bytes * nt;
  if (total == 0) return std::string();
  // a full file system would show up as SIGBUS inside the copy (ftruncate only makes a hole): ask first, and leave a tile
  // that may not fit -- or a file system that does not say -- to pwritev, which reports ENOSPC properly
  struct statvfs vfs;
  if (fstatvfs(mfd, &vfs) != 0 || (unsigned long long)vfs.f_bavail * vfs.f_frsize < (unsigned long long)total + (1ull << 20)) return "!";
  if (ftruncate(mfd, pos + total) != 0) return "!";
  const long long page = sysconf(_SC_PAGESIZE), map0 = pos / page * page;
  const size_t len = (size_t)(pos + total - map0);
  char* m = (char*)mmap(nullptr, len, PROT_READ | PROT_WRITE, MAP_SHARED, mfd, (off_t)map0);
  if (m == MAP_FAILED) return "!";
  char* base = m + (pos - map0);
  auto copy_range = [&](long long a, long long b) {   // bytes [a, b) of the record stream
    while (a < b) {
      const long long t = a / step_bytes, o = a - t * step_bytes;
      const long long n = std::min(b - a, (o < 12 ? 12 : step_bytes) - o);
      memcpy(base + a, o < 12 ? keys + 12 * t + o : payload + rec_bytes * (size_t)t + (o - 12), (size_t)n);
      a += n;
    }
  };
  const int nw = (int)std::max<long long>(1, std::min<long long>(nthreads, total >> 20));   // at least 1 MiB per thread
  std::vector<std::thread> th;
  for (int w = 0; w < nw; ++w) {
    const long long a = total * w / nw, b = total * (w + 1) / nw;
    if (w + 1 < nw) th.emplace_back(copy_range, a, b); else copy_range(a, b);
  }
  for (std::thread& t : th) t.join();
  if (munmap(m, len) != 0) return path + ": write error: " + strerror(errno);
  return std::string();
}

// The writer: waits for the device-to-host copies of a buffer, then appends its step records; the copy into the page cache
// is what takes the time, so the steps of a tile are split over a few helper threads writing at their own file offsets.
void fsr_rdb::writer_main()
{
  const size_t vb = dbl ? 8 : 4, rec_bytes = vb * (size_t)nslot;
  for (;;) {
    RdbJob job;
    {
      std::unique_lock<std::mutex> lk(mtx);
      cv.wait(lk, [&] { return quit || !jobs.empty(); });
      if (jobs.empty()) return;
      job = std::move(jobs.front());
      jobs.pop_front();
    }
    std::string err;
    float a = 0.f, c = 0.f, k1 = 0.f;
    for (RdbDev& d : devs) {
      if (d.nslot == 0) continue;
      cudaSetDevice(d.part->device);
      if (cudaEventSynchronize(d.ev[job.buf][3]) != cudaSuccess) { err = std::string("device error while recovering a tile of steps: ") + cudaGetErrorString(cudaGetLastError()); break; }
      float ai = 0.f, ci = 0.f, ki = 0.f;
      cudaEventElapsedTime(&ai, d.ev[job.buf][0], d.ev[job.buf][1]);
      cudaEventElapsedTime(&ci, d.ev[job.buf][2], d.ev[job.buf][3]);
      cudaEventElapsedTime(&ki, d.ev[job.buf][0], d.ev[job.buf][4]);
      a = std::max(a, ai); c = std::max(c, ci); k1 = std::max(k1, ki);
    }
    const auto t0 = std::chrono::steady_clock::now();
    bool mapped = false;
    if (err.empty() && mfd >= 0) {
      const std::string e = write_steps_mapped(mfd, data_pos, job.keys.data(), (const char*)host[job.buf], rec_bytes, job.nt, nwriters, path);
      if (e == "!") { close(mfd); mfd = -1; }   // no mapping on this file system: pwritev from here on
      else { mapped = true; err = e; if (err.empty()) data_pos += (12 + (long long)rec_bytes) * job.nt; }
    }
    if (err.empty() && !mapped) {
      const int nw = std::max(1, std::min(nwriters, job.nt));
      std::vector<std::string> errs((size_t)nw);
      std::vector<std::thread> th;
      const long long step_bytes = 12 + (long long)rec_bytes;
      for (int w = 0; w < nw; ++w) {
        const int ta = (int)((long long)job.nt * w / nw), tb = (int)((long long)job.nt * (w + 1) / nw);
        auto work = [&, w, ta, tb] { errs[(size_t)w] = write_step_range(fd, data_pos + step_bytes * ta, job.keys.data(), (const char*)host[job.buf], rec_bytes, ta, tb, path); };
        if (w + 1 < nw) th.emplace_back(work); else work();
      }
      for (std::thread& t : th) t.join();
      for (const std::string& e : errs) if (!e.empty() && err.empty()) err = e;
      data_pos += step_bytes * job.nt;
    }
    const double disk = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    {
      std::lock_guard<std::mutex> lk(mtx);
      if (!err.empty() && werr.empty()) werr = err;
      if (err.empty()) {
        ms_compute += a; ms_copy += c; ms_disk += disk; ms_k1 += k1; ++tiles;
        bytes_written += (long long)job.nt * (12 + (long long)rec_bytes);
        steps_written += job.nt;
      }
      busy[job.buf] = false;
    }
    cv.notify_all();
  }
}

namespace {

// The three scratch files of rdbModule.f90 (ivard, iitem, idatd) as strings + the id bookkeeping of HeaderId
struct HeaderBuilder {
  std::string vard, item, datd;
  int nvar = 0, nig = 0, nbit = 32;
  int idSR1 = 0, idSR2 = 0, idSF1 = 0, idSF2 = 0, idStress[4] = {0, 0, 0, 0}, idStrain[4] = {0, 0, 0, 0}, idMeasure[8] = {};

  void vardef(int& id, const char* name, const char* unit, const char* type, int n, const char* comps)
  {
    if (id) return;
    id = ++nvar;   // format 601: '<',i3,';"',a,'";',a,';FLOAT;',i2,';',a,';(',i1,');((',a,'))>'
    appendf(vard, "<%3d;\"%s\";%s;FLOAT;%2d;%s;(%1d);((%s))>\n", id, name, unit, nbit, type, n, comps);
  }
  void scalardef(int& id, const char* name, const char* unit)
  {
    if (id) return;
    id = ++nvar;   // format 605
    appendf(vard, "<%3d;\"%s\";%s;FLOAT;%2d;SCALAR>\n", id, name, unit, nbit);
  }
  void tensor(bool strain, int dim)
  {
    int& id = strain ? idStrain[dim] : idStress[dim];
    const char* nm = strain ? "Strain" : "Stress";
    const char* un = strain ? "NONE" : "FORCE/AREA";
    if (dim == 2)
      vardef(id, nm, un, "TENSOR2", 3, strain ? "\"epsilon_xx\",\"epsilon_yy\",\"epsilon_xy\"" : "\"sigma_xx\",\"sigma_yy\",\"sigma_xy\"");
    else
      vardef(id, nm, un, "TENSOR3", 6,
             strain ? "\"epsilon_xx\",\"epsilon_yy\",\"epsilon_zz\",\"epsilon_xy\",\"epsilon_xz\",\"epsilon_yz\""
                    : "\"sigma_xx\",\"sigma_yy\",\"sigma_zz\",\"sigma_xy\",\"sigma_xz\",\"sigma_yz\"");
  }
  // the per-point variable list shared by shells and solids (writeShellHeader :1048-1166, writeSolidHeader :1221-1330)
  std::string point_vars(const RecLayout& L, int dim)
  {
    static const char* names[8] = {"Von Mises stress", "Max principal stress", "Min principal stress", "Max shear stress",
                                   "Von Mises strain", "Max principal strain", "Min principal strain", "Max shear strain"};
    std::string v;
    if (L.stress) { tensor(false, dim); appendf(v, "<%3d>", idStress[dim]); }
    if (L.strain) { tensor(true, dim); appendf(v, "<%3d>", idStrain[dim]); }
    for (int j = 0; j < 8; ++j)
      if (L.mask & (1 << j)) { scalardef(idMeasure[j], names[j], j < 4 ? "FORCE/AREA" : "NONE"); appendf(v, "<%3d>", idMeasure[j]); }
    return v;
  }
  void node_lines(int n, const std::string& vars) { for (int i = 1; i <= n; ++i) appendf(item, "      [;%2d;%s]\n", i, vars.c_str()); }

  int beam(const RecLayout& L)
  {
    const int ig = ++nig;
    appendf(item, "[%3d;\"BEAM2\";\n", ig);
    if (L.sr) {
      vardef(idSF1, "Beam sectional force", "FORCE", "VEC3", 3, "\"N\",\"V_y\",\"V_z\"");
      vardef(idSF2, "Beam sectional moment", "FORCE*LENGTH", "VEC3", 3, "\"M_x\",\"M_y\",\"M_z\"");
      std::string v;
      appendf(v, "<%3d><%3d>", idSF1, idSF2);
      item += "  [;\"Element nodes\";\n    [;\"Basic\";\n";
      node_lines(2, v);
      item += "    ]\n  ]\n";
    }
    item += "]\n";
    return ig;
  }
  int shell(const RecLayout& L, const char* type, int nelnod)
  {
    const int ig = ++nig;
    appendf(item, "[%3d;\"%s\";\n  [;\"Element nodes\";\n", ig, type);
    if (L.sr) {
      vardef(idSR1, "Shell stress resultant force", "FORCE/LENGTH", "TENSOR2", 3, "\"n_xx\",\"n_yy\",\"n_xy\"");
      vardef(idSR2, "Shell stress resultant moment", "FORCE*LENGTH/LENGTH", "TENSOR2", 3, "\"m_xx\",\"m_yy\",\"m_xy\"");
      std::string v;
      appendf(v, "<%3d><%3d>", idSR1, idSR2);
      item += "    [;\"Basic\";\n";
      node_lines(nelnod, v);
      item += "    ]\n";
    }
    if (L.stress || L.strain || L.nsel) {
      const std::string v = point_vars(L, nelnod > 4 ? 3 : 2);
      for (const char* side : {"Top", "Bottom"}) {
        appendf(item, "    [;\"%s\";\n", side);
        node_lines(nelnod, v);
        item += "    ]\n";
      }
    }
    item += "  ]\n]\n";
    return ig;
  }
  int solid(const RecLayout& L, const char* type, int nelnod)
  {
    const int ig = ++nig;
    appendf(item, "[%3d;\"%s\";\n  [;\"Element nodes\";\n    [;\"Basic\";\n", ig, type);
    node_lines(nelnod, point_vars(L, 3));
    item += "    ]\n  ]\n]\n";
    return ig;
  }
};

}  // namespace

namespace fsr {
// openHeaderFiles (rdbModule.f90:191-251): the meta data lines that precede VARIABLES:
std::string rdb_file_preamble(const char* module, const char* model_file, const char* link_file, const char* info)
{
  std::string t;
  char host[96] = "unknown", date[32] = "";
  gethostname(host, sizeof(host) - 1);
  const char* user = getenv("USER");
  time_t now = time(nullptr);
  strftime(date, sizeof(date), "%d %b %Y %H:%M:%S", localtime(&now));
  if (model_file && *model_file) appendf(t, " AssociatedModelFileName = %s;\n", model_file);
  if (link_file && *link_file) appendf(t, " ModelName               = %s;\n", link_file);
  appendf(t, " InformationText         = %s;\n", info ? info : "response data base file");
  appendf(t, " User                    = %s;\n", user ? user : "unknown");
  appendf(t, " Computer                = %s;\n", host);
  appendf(t, " DateTime                = %s;\n", date);
  t += " UsedTime                = 00:00:00.00;\n";
  appendf(t, " Module                  = %s;\n", module);
  t += " ModuleVersion           = B200 1.0;\n";
  return t;
}
}  // namespace fsr

namespace {

// header text + record slot of every element (-1 = not written); returns the number of values per step
static long long build_header(int nnod, const int* madof, int nel, const int* melcon, const int* active,
                              const fsr_rdb_options* o, const RecLayout& L, std::string& header, std::vector<long long>& slot,
                              std::vector<long long>& node_slot)
{
  HeaderBuilder hb;
  hb.nbit = o->double_precision ? 64 : 32;
  hb.vard = rdb_file_preamble(o->module_name ? o->module_name : "fedem_stress", o->model_file, o->link_file) + "VARIABLES:\n";
  // writeTimeStepHeader (rdbModule.f90:633-650)
  hb.nvar = 2;
  hb.vard += "<1;\"Time step number\";NONE;INT;32;NUMBER>\n<2;\"Physical time\";TIME;FLOAT;64;SCALAR>\n";
  hb.datd = "DATABLOCKS:\n<1><2>\n";
  // writeIdHeader('Part',sup%id,idatd,.true.) (idTypeModule.f90:103-143)
  hb.datd += "{\"Part\";";
  if (o->part_base_id > 0) appendf(hb.datd, "%d;", o->part_base_id); else hb.datd += ";";
  if (o->part_user_id > 0) appendf(hb.datd, "%d;", o->part_user_id); else hb.datd += ";";
  if (o->part_descr && *o->part_descr) appendf(hb.datd, "\"%s\";\n", o->part_descr); else hb.datd += ";\n";
  long long nslot = 0;
  node_slot.assign((size_t)std::max(nnod, 1), -1);
  if (L.def) {   // writeNodesHeader (saveStressModule.f90:438-537), deformations as nodal data
    int idDis[2] = {0, 0}, idRot[2] = {0, 0}, id3 = 0, id6 = 0;
    const bool tot = L.def > 1;
    hb.datd += "  [;\"Nodes\";\n";
    for (int n = 0; n < nnod; ++n) {
      const int nd = madof[n + 1] - madof[n];
      if (nd < 3) continue;
      hb.vardef(idDis[0], "Translational deformation", "LENGTH", "VEC3", 3, "\"d_x\",\"d_y\",\"d_z\"");
      if (tot) hb.vardef(idDis[1], "Total translation", "LENGTH", "VEC3", 3, "\"u_x\",\"u_y\",\"u_z\"");
      int ig = 0, nv = 0;
      if (nd > 5) {
        hb.vardef(idRot[0], "Angular deformation", "ANGLE", "ROT3", 3, "\"theta_x\",\"theta_y\",\"theta_z\"");
        if (tot) hb.vardef(idRot[1], "Total rotation", "ANGLE", "ROT3", 3, "\"theta_x\",\"theta_y\",\"theta_z\"");
        if (!id6) {
          id6 = ++hb.nig;
          if (tot) appendf(hb.item, "[%3d;\"Dynamic response\";<%3d><%3d><%3d><%3d>]\n", id6, idDis[0], idRot[0], idDis[1], idRot[1]);
          else appendf(hb.item, "[%3d;\"Dynamic response\";<%3d><%3d>]\n", id6, idDis[0], idRot[0]);
        }
        ig = id6; nv = tot ? 12 : 6;
      } else {
        if (!id3) {
          id3 = ++hb.nig;
          if (tot) appendf(hb.item, "[%3d;\"Dynamic response\";<%3d><%3d>]\n", id3, idDis[0], idDis[1]);
          else appendf(hb.item, "[%3d;\"Dynamic response\";<%3d>]\n", id3, idDis[0]);
        }
        ig = id3; nv = tot ? 6 : 3;
      }
      appendf(hb.datd, "    [;%8d;[%3d]]\n", o->minex ? o->minex[n] : n + 1, ig);
      node_slot[(size_t)n] = nslot;
      nslot += nv;
    }
    hb.datd += "  ]\n";
  }
  // writeElementsHeader (saveStressModule.f90:625-752) + the record slot of every element
  slot.assign((size_t)std::max(nel, 1), -1);
  int ig[64] = {};
  const bool elements = L.sr || L.stress || L.strain || L.nsel;
  if (elements) hb.datd += "  [;\"Elements\";\n";
  if (!elements) nel = 0;
  for (int e = 0; e < nel; ++e) {
    const int elmno = o->elmid ? o->elmid[e] : e + 1;
    if (elmno <= 0 || (active && !active[e])) continue;
    const int t = melcon[e];
    int nelnod = 0, ncmp = 0;
    bool shell = false;
    switch (t) {
      case 11: if (!L.sr) continue; if (!ig[t]) ig[t] = hb.beam(L); nelnod = 2; break;
      case 21: case 23: if (!ig[21]) ig[21] = hb.shell(L, "TRI3", 3); nelnod = 3; ncmp = 3; shell = true; break;
      case 22: case 24: if (!ig[22]) ig[22] = hb.shell(L, "QUAD4", 4); nelnod = 4; ncmp = 3; shell = true; break;
      case 31: if (!ig[t]) ig[t] = hb.shell(L, "TRI6", 6); nelnod = 6; ncmp = 6; shell = true; break;
      case 32: if (!ig[t]) ig[t] = hb.shell(L, "QUAD8", 8); nelnod = 8; ncmp = 6; shell = true; break;
      case 41: if (!ig[t]) ig[t] = hb.solid(L, "TET10", 10); nelnod = 10; ncmp = 6; break;
      case 42: if (!ig[t]) ig[t] = hb.solid(L, "WEDG15", 15); nelnod = 15; ncmp = 6; break;
      case 43: if (!ig[t]) ig[t] = hb.solid(L, "HEX20", 20); nelnod = 20; ncmp = 6; break;
      case 44: if (!ig[t]) ig[t] = hb.solid(L, "HEX8", 8); nelnod = 8; ncmp = 6; break;
      case 45: if (!ig[t]) ig[t] = hb.solid(L, "TET4", 4); nelnod = 4; ncmp = 6; break;
      case 46: if (!ig[t]) ig[t] = hb.solid(L, "WEDG6", 6); nelnod = 6; ncmp = 6; break;
      default: continue;   // element types without a stress operator in this library
    }
    const int key = t == 23 ? 21 : t == 24 ? 22 : t;
    appendf(hb.datd, "    [;%8d;[%3d]]\n", elmno, ig[key]);
    slot[(size_t)e] = nslot;
    if (t == 11) nslot += 12;
    else {
      const int nstrp = shell ? 2 * nelnod : nelnod;
      nslot += (L.sr && shell ? 6 * nelnod : 0) + (long long)nstrp * ((L.stress ? ncmp : 0) + (L.strain ? ncmp : 0) + L.nsel);
    }
  }
  if (elements) hb.datd += "  ]\n";
  hb.datd += "}\n";
  header = hb.vard + hb.item + hb.datd;
  return nslot;
}

static int layout_from_options(const fsr_rdb_options* o, RecLayout& L)
{
  L = RecLayout{};
  L.sr = (o->out_mask & FSR_OUT_SR) ? 1 : 0;
  L.stress = (o->out_mask & FSR_OUT_STRESS) ? 1 : 0;
  L.strain = (o->out_mask & FSR_OUT_STRAIN) ? 1 : 0;
  L.mask = (int)(o->out_mask & 0xff);
  for (int j = 0; j < 8; ++j) L.nsel += (L.mask >> j) & 1;
  if (o->out_mask & FSR_OUT_DEFORMATION) L.def = o->sup_tr_init ? 3 : 1;
  if (!(L.sr || L.stress || L.strain || L.nsel || L.def)) { set_error("no result output requested"); return FSR_ERR_ARG; }
  return FSR_OK;
}

}  // namespace

template <class OUT_T>
static int launch_record_kernels(fsr_rdb* r, RdbDev& d, int ts, int nt, OUT_T* out, cudaStream_t s)
{
  fsr_part* p = d.part;
  const size_t ld_out = (size_t)d.nslot;
  const double* U = p->U + ts;   // steps [ts, ts + nt) of the expanded chunk
  for (int fi = 0; fi < FAM_COUNT; ++fi) {
    FamilyData& f = p->fam[fi];
    if (f.nelt == 0 || !d.roff[fi]) continue;
    if (fi == FAM_BEAM) {
      if (!r->L.sr) continue;
      dim3 blk(96, 4), grd((unsigned)(((long long)f.nelt * 12 + 95) / 96), (nt + 3) / 4);
      record_beams_kernel<OUT_T><<<grd, blk, 0, s>>>(U, (size_t)p->step_tile, nt, reinterpret_cast<const BeamOp12*>(f.Sfrag), f.edof,
                                                     d.roff[fi], f.failed, f.nelt, out, ld_out);
    } else {
      if (f.nstrp == 0) continue;
      if (!(r->L.stress || r->L.strain || r->L.nsel || (r->L.sr && f.ncmp == 3) || (r->L.sr && f.Efrag))) continue;
      const int layout = (fi == FAM_QUAD || fi == FAM_TRI) ? 0 : (fi == FAM_TRI6 || fi == FAM_QUAD8) ? 2 : 1;
      // `-vmStress` alone on thin shells: the tuned von Mises kernel (operator fragments in registers) writes the records
      if (layout == 0 && r->L.mask == 0x01 && !r->L.sr && !r->L.stress && !r->L.strain && !getenv("FSR_RDB_GENERIC")) {
        if (int rc = launch_k2_shell_rec<OUT_T>(p, fi, U, nt, (nt + 7) / 8 * 8, d.roff[fi], out, ld_out, s)) return rc;
        continue;
      }
      const int srsz = r->L.sr && layout != 1 ? 6 * f.nenod : 0;
      const int nvl = srsz + f.nstrp * ((r->L.stress ? f.ncmp : 0) + (r->L.strain ? f.ncmp : 0) + r->L.nsel);
      const size_t smem = sizeof(double) * kRecWarps * ((layout == 2 ? 2 : 1) * (size_t)f.MT * 8 * kRecLds +
                                                        ((size_t)8 * ((nvl + 1) & ~1) * sizeof(OUT_T) + 7) / 8);
      const unsigned grid = (unsigned)((f.nelt + kRecWarps - 1) / kRecWarps);
#define FSR_REC_LAUNCH(KTV, LAY)                                                                                                   \
  if (f.KT == KTV && layout == LAY) {                                                                                              \
    if (smem > 48 * 1024)                                                                                                          \
      if (int rc = smem_opt_in((const void*)record_points_dmma_kernel<KTV, LAY, OUT_T>, 200 * 1024)) return rc;                   \
    record_points_dmma_kernel<KTV, LAY, OUT_T><<<grid, kRecWarps * 32, smem, s>>>(U, (size_t)p->step_tile, nt, f.Sfrag, f.Efrag,   \
                                                                                   f.edof, d.roff[fi], f.failed, f.aux, f.naux,    \
                                                                                   f.nelt, f.nstrp, f.MT, f.nenod, r->L, out, ld_out); \
    launched = true;                                                                                                               \
  }
      // the (k-tile count, row layout) pairs of the element families (common.cuh): loops over both are unrolled at compile time
      bool launched = false;
      FSR_REC_LAUNCH(6, 0) FSR_REC_LAUNCH(5, 0)                                                   // ANDES quad, triangle
      FSR_REC_LAUNCH(8, 1) FSR_REC_LAUNCH(15, 1) FSR_REC_LAUNCH(12, 1)                            // TET10, HEX20, WEDG15
      FSR_REC_LAUNCH(6, 1) FSR_REC_LAUNCH(3, 1) FSR_REC_LAUNCH(5, 1)                              // HEX8, TET4, WEDG6
      FSR_REC_LAUNCH(9, 2) FSR_REC_LAUNCH(12, 2)                                                  // TRI6, QUAD8 thick shells
      if (!launched) { set_error("internal: no record kernel for an operator with %d k-tiles, layout %d (family %d)", f.KT, layout, fi); return FSR_ERR_LIMIT; }
#undef FSR_REC_LAUNCH
    }
    FSR_LAUNCH_CHECK();
  }
  return FSR_OK;
}

// common part of fsr_rdb_create / fsr_rdb_create_group: parts = the element blocks (one: the whole part), e_cut their
// element ranges in the parent, the remaining arguments describe the parent
static int rdb_create(fsr_rdb** out, const std::vector<fsr_part*>& parts, const std::vector<int>& e_cut, int nnod, const int* madof,
                      int nel, const int* melcon, const int* active, const char* path, const fsr_rdb_options* o)
{
  *out = nullptr;
  RecLayout L;
  int rc0 = layout_from_options(o, L);
  if (rc0) return rc0;
  if (L.def && parts.size() > 1) { set_error("nodal deformation output is written by a single device: create the results database on one part"); return FSR_ERR_ARG; }
  std::string header;
  std::vector<long long> slot, node_slot;
  const long long nslot = build_header(nnod, madof, nel, melcon, active, o, L, header, slot, node_slot);
  if (nslot == 0) { set_error("fsr_rdb_create: none of the active elements has the requested results"); return FSR_ERR_ARG; }

  fsr_rdb* r = new fsr_rdb;
  fsr_part* p = parts[0];
  r->part = p; r->L = L; r->dbl = o->double_precision ? 1 : 0; r->nslot = nslot;
  r->header = header;
  // openRDBfile (rdbModule.f90:268-403): <name>_<rdbinc>.<ext>
  r->path = path;
  if (o->rdbinc > 0) {
    const size_t dot = r->path.rfind('.');
    const size_t sep = r->path.rfind('/');
    char inc[16];
    snprintf(inc, sizeof(inc), "_%d", o->rdbinc);
    if (dot != std::string::npos && dot > 0 && (sep == std::string::npos || dot > sep)) r->path.insert(dot, inc);
    else r->path += inc;
  }
  TaggedFile tf;
  int rc = tf.open_write(r->path.c_str(), "#FEDEM response data", 0u);
  if (rc) { delete r; return rc; }
  if (fputs(r->header.c_str(), tf.f) < 0 || fputs("DATA:", tf.f) < 0) { set_error("%s: write error", r->path.c_str()); delete r; return FSR_ERR_ARG; }
  r->f = tf.f;
  tf.f = nullptr;
  // slot range of every block: blocks are contiguous element ranges and slots ascend with the element number
  r->devs.resize(parts.size());
  const long long first_elem_slot = [&] { for (int e = 0; e < nel; ++e) if (slot[(size_t)e] >= 0) return slot[(size_t)e]; return nslot; }();
  for (size_t ib = 0; ib < parts.size(); ++ib) {
    RdbDev& d = r->devs[ib];
    d.part = parts[ib];
    long long s0 = -1, s1 = nslot;
    for (int e = e_cut[ib]; e < e_cut[ib + 1] && s0 < 0; ++e) if (slot[(size_t)e] >= 0) s0 = slot[(size_t)e];
    for (int e = e_cut[ib + 1]; e < nel; ++e) if (slot[(size_t)e] >= 0) { s1 = slot[(size_t)e]; break; }
    d.slot0 = s0 < 0 ? s1 : s0;
    d.nslot = s0 < 0 ? 0 : s1 - s0;
    if (parts.size() == 1) { d.slot0 = 0; d.nslot = nslot; }   // the single device also owns the nodal values in front
  }
  (void)first_elem_slot;
  // per device and family: record slot of each of its elements, relative to the device's range
  for (size_t ib = 0; ib < parts.size(); ++ib) {
    RdbDev& d = r->devs[ib];
    if (cudaSetDevice(d.part->device) != cudaSuccess) { set_error("cudaSetDevice failed"); delete r; return FSR_ERR_CUDA; }
    for (int fi = 0; fi < FAM_COUNT; ++fi) {
      FamilyData& f = d.part->fam[fi];
      if (f.nelt == 0) continue;
      std::vector<int> elem((size_t)f.nelt);
      if (cudaMemcpy(elem.data(), f.elem, sizeof(int) * f.nelt, cudaMemcpyDeviceToHost) != cudaSuccess) { set_error("fsr_rdb_create: cudaMemcpy failed"); delete r; return FSR_ERR_CUDA; }
      std::vector<long long> ro((size_t)f.nelt);
      for (int i = 0; i < f.nelt; ++i) {
        const long long sl = slot[(size_t)e_cut[ib] + (size_t)elem[(size_t)i]];
        ro[(size_t)i] = sl < 0 ? -1 : sl - d.slot0;
      }
      if (cudaMalloc(&d.roff[fi], sizeof(long long) * f.nelt) != cudaSuccess ||
          cudaMemcpy(d.roff[fi], ro.data(), sizeof(long long) * f.nelt, cudaMemcpyHostToDevice) != cudaSuccess) {
        set_error("fsr_rdb_create: device allocation failed"); delete r; return FSR_ERR_ALLOC;
      }
    }
  }
  cudaSetDevice(p->device);
  if (L.def) {
    if (cudaMalloc(&r->node_slot, sizeof(long long) * std::max(p->nnod, 1)) != cudaSuccess ||
        cudaMalloc(&r->madof, sizeof(int) * (p->nnod + 1)) != cudaSuccess ||
        cudaMemcpy(r->node_slot, node_slot.data(), sizeof(long long) * p->nnod, cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(r->madof, p->madof_host.data(), sizeof(int) * (p->nnod + 1), cudaMemcpyHostToDevice) != cudaSuccess) {
      set_error("fsr_rdb_create: device allocation failed"); delete r; return FSR_ERR_ALLOC;
    }
    if (L.def > 1 && (cudaMalloc(&r->supTr0, sizeof(double) * 12) != cudaSuccess ||
                      cudaMemcpy(r->supTr0, o->sup_tr_init, sizeof(double) * 12, cudaMemcpyHostToDevice) != cudaSuccess)) {
      set_error("fsr_rdb_create: device allocation failed"); delete r; return FSR_ERR_ALLOC;
    }
  }
  // step tile of the record buffers: bounded by the parts' step tile, by 256 MiB per pinned host buffer (two of them; pinning
  // costs about a second per GiB at program start, and the file, not the device, paces the pipeline) and by ~1/4 of the free
  // device memory
  r->nslot_nodes = 0;
  if (L.def)
    for (int n = 0; n < p->nnod; ++n)
      if (node_slot[(size_t)n] >= 0) {
        const int nd = p->madof_host[(size_t)n + 1] - p->madof_host[(size_t)n] > 5 ? 6 : 3;
        r->nslot_nodes = std::max(r->nslot_nodes, node_slot[(size_t)n] + (L.def > 1 ? 2 : 1) * nd);
      }
  const size_t vb = r->dbl ? 8 : 4;
  long long tile = (long long)((double)(1u << 28) / ((double)nslot * (double)vb));
  for (RdbDev& d : r->devs) {
    size_t free_b = 0, total_b = 0;
    cudaSetDevice(d.part->device);
    cudaMemGetInfo(&free_b, &total_b);
    const double per_step = 2.0 * (double)std::max<long long>(d.nslot, 1) * (double)vb + 8.0 * (double)r->nslot_nodes;
    tile = std::min<long long>(tile, (long long)(0.25 * (double)free_b / per_step));
    tile = std::min<long long>(tile, d.part->step_tile);
  }
  if (const char* e = getenv("FSR_RDB_TILE")) tile = std::min<long long>(std::max(1, atoi(e)), tile);   // tests: several tiles per call
  tile = std::max<long long>(getenv("FSR_RDB_TILE") ? 1 : 8, tile);   // the record kernels work on 8 steps at a time
  if (tile >= 8) tile = tile / 8 * 8;
  r->tile = (int)tile;
  // the expansion (K1) works on chunks of the parts' step tile; the record tiles of a chunk read the same U
  r->ktile = 1 << 30;
  for (RdbDev& d : r->devs) r->ktile = std::min(r->ktile, d.part->step_tile);
  r->ktile = std::max(r->ktile / r->tile, 1) * r->tile;
  bool ok = true;
  for (RdbDev& d : r->devs) {
    cudaSetDevice(d.part->device);
    ok = ok && cudaStreamCreateWithFlags(&d.copy_stream, cudaStreamNonBlocking) == cudaSuccess;
    for (int b = 0; b < 2 && ok; ++b) {
      ok = cudaMalloc(&d.out[b], vb * (size_t)std::max<long long>(d.nslot, 1) * r->tile) == cudaSuccess;
      for (auto& e : d.ev[b]) ok = ok && cudaEventCreate(&e) == cudaSuccess;
      ok = ok && cudaEventCreateWithFlags(&d.ev_q[b], cudaEventDisableTiming) == cudaSuccess;
    }
  }
  cudaSetDevice(p->device);
  for (int b = 0; b < 2 && ok; ++b)
    ok = cudaHostAlloc(&r->host[b], vb * (size_t)nslot * r->tile, cudaHostAllocPortable) == cudaSuccess &&
         (L.def <= 1 || cudaHostAlloc((void**)&r->supPin[b], sizeof(double) * 12 * r->ktile, cudaHostAllocPortable) == cudaSuccess);
  ok = ok && (r->nslot_nodes == 0 || cudaMalloc(&r->rec, sizeof(double) * (size_t)r->nslot_nodes * r->tile) == cudaSuccess) &&
       (L.def <= 1 || cudaMalloc(&r->supTr, sizeof(double) * 12 * r->ktile) == cudaSuccess);
  if (!ok) {
    set_error("fsr_rdb_create: cannot allocate the record buffers (%lld values x %d steps): %s", nslot, r->tile,
              cudaGetErrorString(cudaGetLastError()));
    delete r;
    return FSR_ERR_ALLOC;
  }
  // the header went through stdio; the step records go through the descriptor at explicit offsets
  if (fflush(r->f) != 0) { set_error("%s: write error", r->path.c_str()); delete r; return FSR_ERR_ARG; }
  r->fd = fileno(r->f);
  r->data_pos = (long long)ftello(r->f);
  const unsigned hc = std::thread::hardware_concurrency();
  r->nwriters = (int)std::max(1u, std::min(8u, hc / 2));
  if (const char* e = getenv("FSR_RDB_WRITERS")) r->nwriters = std::max(1, atoi(e));
  // step records copied into a shared mapping of the file by the helper threads (measured: 16 GB in 2.2 s on tmpfs, 3.4 s on the
  // container's disk, against 4.2 s / 4.1 s with pwritev; MADV_POPULATE_WRITE and 16 threads measured slower);
  // FSR_RDB_MMAP=0 keeps pwritev, which is also the fallback when the file cannot be mapped
  if (!(getenv("FSR_RDB_MMAP") && atoi(getenv("FSR_RDB_MMAP")) == 0)) r->mfd = open(r->path.c_str(), O_RDWR);
  r->writer = std::thread(&fsr_rdb::writer_main, r);
  *out = r;
  return FSR_OK;
}

extern "C" {

// Host only: the header text and record size for a part given by its element type codes (SAM melcon) and
// external element ids.  header may be NULL / cap 0 to query the length.  Returns the header length.
int fsr_rdb_build_header(int nnod, const int* madof, int nel, const int* melcon, const fsr_rdb_options* o, char* header, int cap,
                         long long* step_bytes)
{
  if (nel < 0 || !melcon || !o || ((o->out_mask & FSR_OUT_DEFORMATION) && (nnod < 0 || !madof))) { set_error("fsr_rdb_build_header: bad arguments"); return FSR_ERR_ARG; }
  RecLayout L;
  int rc = layout_from_options(o, L);
  if (rc) return rc;
  std::string h;
  std::vector<long long> slot, node_slot;
  const long long nslot = build_header(nnod, madof, nel, melcon, nullptr, o, L, h, slot, node_slot);
  if (step_bytes) *step_bytes = 12 + nslot * (o->double_precision ? 8 : 4);
  if (header && cap > 0) { strncpy(header, h.c_str(), (size_t)cap - 1); header[cap - 1] = 0; }
  return (int)h.size();
}

int fsr_rdb_create(fsr_rdb** out, fsr_part* p, const char* path, const fsr_rdb_options* o)
{
  if (!out || !p || !path || !o) { set_error("fsr_rdb_create: bad arguments"); return FSR_ERR_ARG; }
  FSR_CUDA(cudaSetDevice(p->device));
  return rdb_create(out, {p}, {0, p->nel}, p->nnod, p->madof_host.data(), p->nel, p->melcon_host.data(), p->active_host.data(), path, o);
}

// the results database of a part recovered by a group of element blocks on several GPUs (no nodal deformation output)
int fsr_rdb_create_group(fsr_rdb** out, fsr_group* g, const char* path, const fsr_rdb_options* o)
{
  if (!out || !g || !path || !o) { set_error("fsr_rdb_create_group: bad arguments"); return FSR_ERR_ARG; }
  return rdb_create(out, g->parts, g->e_cut, g->nnod, g->madof_host.data(), g->nel, g->melcon_host.data(), g->active_host.data(), path, o);
}

long long fsr_rdb_step_bytes(const fsr_rdb* r) { return r ? 12 + r->nslot * (r->dbl ? 8 : 4) : 0; }

int fsr_rdb_header(const fsr_rdb* r, char* buf, int cap)
{
  if (!r) return FSR_ERR_ARG;
  if (buf && cap > 0) { strncpy(buf, r->header.c_str(), (size_t)cap - 1); buf[cap - 1] = 0; }
  return (int)r->header.size();
}

int fsr_rdb_path(const fsr_rdb* r, char* buf, int cap)
{
  if (!r) return FSR_ERR_ARG;
  if (buf && cap > 0) { strncpy(buf, r->path.c_str(), (size_t)cap - 1); buf[cap - 1] = 0; }
  return (int)r->path.size();
}

static int rdb_write(fsr_rdb* r, const double* Q, int ldq, const double* sv_hist, int nsteps, const int* stepno, const double* time,
                     const double* sup_tr);

int fsr_rdb_write_steps(fsr_rdb* r, const double* Q, int ldq, int nsteps, const int* stepno, const double* time,
                        const double* sup_tr)
{
  if (!Q) { set_error("fsr_rdb_write_steps: bad arguments"); return FSR_ERR_ARG; }
  return rdb_write(r, Q, ldq, nullptr, nsteps, stepno, time, sup_tr);
}

// the same from nodal displacements that are already there (fsr_recover_displacements): sv_hist [nsteps x ndof] step-major
int fsr_rdb_write_steps_displacements(fsr_rdb* r, const double* sv_hist, int nsteps, const int* stepno, const double* time,
                                      const double* sup_tr)
{
  if (!sv_hist) { set_error("fsr_rdb_write_steps_displacements: bad arguments"); return FSR_ERR_ARG; }
  if (r && r->devs.size() > 1) { set_error("fsr_rdb_write_steps_displacements: one device only (the element blocks of a group hold their own nodes)"); return FSR_ERR_ARG; }
  return rdb_write(r, nullptr, 0, sv_hist, nsteps, stepno, time, sup_tr);
}

static int rdb_write(fsr_rdb* r, const double* Q, int ldq, const double* sv_hist, int nsteps, const int* stepno, const double* time,
                     const double* sup_tr)
{
  if (!r || !r->f || nsteps < 0 || !stepno || !time) { set_error("fsr_rdb_write_steps: bad arguments"); return FSR_ERR_ARG; }
  if (r->L.def > 1 && !sup_tr) { set_error("fsr_rdb_write_steps: total displacements need the part position matrix of every step"); return FSR_ERR_ARG; }
  fsr_part* p = r->part;
  if (Q && ldq < p->ndim) { set_error("fsr_rdb_write_steps: ldq < ndim"); return FSR_ERR_ARG; }
  if (!Q) ldq = 1;
  int rc;
  for (RdbDev& d : r->devs) {
    if (Q && !d.part->have_R) { set_error("fsr_rdb_write_steps: call fsr_set_recovery first"); return FSR_ERR_STATE; }
    FSR_CUDA(cudaSetDevice(d.part->device));
    if ((rc = ensure_batch_buffers(d.part, false))) return rc;
  }
  const size_t vb = r->dbl ? 8 : 4;
  if (ldq > r->ldq_cap) {   // Q staging, sized on first use (ldq is the caller's)
    r->drain();
    for (RdbDev& d : r->devs) {
      FSR_CUDA(cudaSetDevice(d.part->device));
      FSR_CUDA(cudaStreamSynchronize(d.part->stream));
      cudaFree(d.dQ); d.dQ = nullptr;
      FSR_CUDA(cudaMalloc(&d.dQ, sizeof(double) * (size_t)ldq * r->ktile));
    }
    for (int b = 0; b < 2; ++b) {
      if (r->Qpin[b]) cudaFreeHost(r->Qpin[b]);
      r->Qpin[b] = nullptr;
      FSR_CUDA(cudaHostAlloc((void**)&r->Qpin[b], sizeof(double) * (size_t)ldq * r->ktile, cudaHostAllocPortable));
    }
    r->ldq_cap = ldq;
  }
  for (int c0 = 0; c0 < nsteps; c0 += r->ktile) {   // expansion chunks
    const int nc = std::min(r->ktile, nsteps - c0), nc_pad = (nc + 63) / 64 * 64;
    const int qb = r->next_q;
    for (RdbDev& d : r->devs) {   // the H2D copy of the chunk before the previous one has left this staging buffer
      FSR_CUDA(cudaSetDevice(d.part->device));
      FSR_CUDA(cudaEventSynchronize(d.ev_q[qb]));
    }
    if (Q) memcpy(r->Qpin[qb], Q + (size_t)c0 * ldq, sizeof(double) * (size_t)ldq * nc);
    if (r->L.def > 1) memcpy(r->supPin[qb], sup_tr + (size_t)c0 * 12, sizeof(double) * 12 * nc);
    bool first = true;
    for (int ts = 0; ts < nc; ts += r->tile) {   // record tiles of the chunk
      const int nt = std::min(r->tile, nc - ts);
      const int b = r->next_buf;
      {   // buffer b is free again once the writer has put its previous content on file
        std::unique_lock<std::mutex> lk(r->mtx);
        r->cv.wait(lk, [&] { return !r->busy[b]; });
        if (!r->werr.empty()) { set_error("%s", r->werr.c_str()); return FSR_ERR_ARG; }
      }
      for (RdbDev& d : r->devs) {
        if (d.nslot == 0) continue;
        fsr_part* dp = d.part;
        cudaStream_t s = dp->stream;
        FSR_CUDA(cudaSetDevice(dp->device));
        FSR_CUDA(cudaEventRecord(d.ev[b][0], s));
        if (first) {
          if (Q) FSR_CUDA(cudaMemcpyAsync(d.dQ, r->Qpin[qb], sizeof(double) * (size_t)ldq * nc, cudaMemcpyHostToDevice, s));
          if (r->L.def > 1) FSR_CUDA(cudaMemcpyAsync(r->supTr, r->supPin[qb], sizeof(double) * 12 * nc, cudaMemcpyHostToDevice, s));
          FSR_CUDA(cudaEventRecord(d.ev_q[qb], s));
          if (Q) {
            if ((rc = launch_pack_q(dp, d.dQ, ldq, nc, nc_pad, s)) || (rc = launch_k1(dp, nc_pad, s))) return rc;
          } else if ((rc = upload_displacements(dp, sv_hist + (size_t)c0 * dp->ndof, nc, s)))   // pageable source: the copy is staged before the call returns
            return rc;
        }
        FSR_CUDA(cudaEventRecord(d.ev[b][4], s));
        if (r->L.def) {   // nodal values: slot-major staging, then one tiled transpose into the leading part of the records
          const size_t ldt = (size_t)r->tile;
          dim3 blk(32, 8), grd((dp->nnod + 7) / 8, (nt + 31) / 32);
          record_nodes_kernel<<<grd, blk, 0, s>>>(dp->U + ts, (size_t)dp->step_tile, nt, dp->nnod, r->madof, r->node_slot, dp->xyz,
                                                  r->supTr ? r->supTr + (size_t)12 * ts : nullptr, r->supTr0, r->L.def > 1, r->rec, ldt);
          FSR_LAUNCH_CHECK();
          dim3 tg((unsigned)((r->nslot_nodes + 31) / 32), (nt + 31) / 32);
          if (r->dbl) record_transpose_kernel<double><<<tg, blk, 0, s>>>(r->rec, ldt, r->nslot_nodes, nt, (double*)d.out[b], (size_t)d.nslot);
          else record_transpose_kernel<float><<<tg, blk, 0, s>>>(r->rec, ldt, r->nslot_nodes, nt, (float*)d.out[b], (size_t)d.nslot);
          FSR_LAUNCH_CHECK();
        }
        rc = r->dbl ? launch_record_kernels<double>(r, d, ts, nt, (double*)d.out[b], s) : launch_record_kernels<float>(r, d, ts, nt, (float*)d.out[b], s);
        if (rc) return rc;
        FSR_CUDA(cudaEventRecord(d.ev[b][1], s));
        FSR_CUDA(cudaStreamWaitEvent(d.copy_stream, d.ev[b][1], 0));
        FSR_CUDA(cudaEventRecord(d.ev[b][2], d.copy_stream));
        // this device's slot range of every step record of the tile
        FSR_CUDA(cudaMemcpy2DAsync((char*)r->host[b] + vb * (size_t)d.slot0, vb * (size_t)r->nslot, d.out[b], vb * (size_t)d.nslot,
                                   vb * (size_t)d.nslot, (size_t)nt, cudaMemcpyDeviceToHost, d.copy_stream));
        FSR_CUDA(cudaEventRecord(d.ev[b][3], d.copy_stream));
      }
      first = false;
      RdbJob job;
      job.buf = b; job.nt = nt;
      job.keys.resize(12 * (size_t)nt);
      for (int t = 0; t < nt; ++t) {
        memcpy(job.keys.data() + 12 * (size_t)t, &stepno[c0 + ts + t], 4);
        memcpy(job.keys.data() + 12 * (size_t)t + 4, &time[c0 + ts + t], 8);
      }
      {
        std::lock_guard<std::mutex> lk(r->mtx);
        r->busy[b] = true;
        r->jobs.push_back(std::move(job));
      }
      r->cv.notify_all();
      r->next_buf ^= 1;
    }
    r->next_q ^= 1;
  }
  return FSR_OK;
}

// Waits until every record handed over so far is on file.  t (may be NULL): [0] device time of the tiles (H2D of Q, K1,
// record kernels), [1] device-to-host copies, [2] file writes, all in ms and summed over the tiles (they overlap in wall
// time), [3] bytes written, [4] tiles, [5] the part of [0] spent on the H2D copy + the expansion (K1).  Returns the number
// of entries written or a negative error.
int fsr_rdb_flush(fsr_rdb* r, double* t, int n)
{
  if (!r) { set_error("fsr_rdb_flush: null handle"); return FSR_ERR_ARG; }
  r->drain();
  std::lock_guard<std::mutex> lk(r->mtx);
  if (!r->werr.empty()) { set_error("%s", r->werr.c_str()); return FSR_ERR_ARG; }
  const double v[6] = {r->ms_compute, r->ms_copy, r->ms_disk, (double)r->bytes_written, (double)r->tiles, r->ms_k1};
  const int m = t ? std::min(n, 6) : 0;
  for (int i = 0; i < m; ++i) t[i] = v[i];
  return m;
}

void fsr_total_nodal_displacement(const double* x0, const double* u, int nd, const double* T, const double* T0, double* utot)
{
  total_nodal_displacement(x0, u, nd, T, T0, utot);
}

int fsr_rdb_close(fsr_rdb* r)
{
  if (!r) return FSR_ERR_ARG;
  int rc = FSR_OK;
  r->drain();
  if (!r->werr.empty()) { set_error("%s", r->werr.c_str()); rc = FSR_ERR_ARG; }
  {
    std::lock_guard<std::mutex> lk(r->mtx);
    r->quit = true;
  }
  r->cv.notify_all();
  if (r->writer.joinable()) r->writer.join();
  if (r->f && fclose(r->f) != 0 && rc == FSR_OK) { set_error("%s: close error", r->path.c_str()); rc = FSR_ERR_ARG; }
  r->f = nullptr;
  delete r;
  return rc;
}

}  // extern "C"
