// k2_beam.cu -- K2 for the 2-node beam (type 11): section forces SF(6,2) = S_e(12 x 12) . v_e.
//
// Reference: STR11 (src/vpmStress/elStressModule.f90:402-515) rebuilds, per beam and per step, the
// 12x12 stiffness in global axes (BEAM31 -> BELS31 / DCOS30 / MPRO30 / TRIX30,
// src/Femlib/beam.f:11-120,619-770, beamaux.f:48-208), multiplies its first six rows with the nodal
// displacements, moves the moments to the element end, rotates to the local axes and obtains end 2
// by equilibrium.  Beams have no stress points (nstrp = 0, :511-513): they never contribute von
// Mises, only section forces.  All of it is linear in v_e and time-invariant, so one 12x12 operator
// per beam is built once on the GPU and applied per step.
#include "common.cuh"

namespace fsr {

struct BeamOp { double S[12][12]; };

__device__ void beam_dcos(double T[3][3], const double* X, const double* Y, const double* Z)
{
  // rows: local x (1->2), y = (z-point - 1) x x ... exactly DCOS30's sequence of cross products
  double cx = X[1] - X[0], cy = Y[1] - Y[0], cz = Z[1] - Z[0];
  double ab = sqrt(cx * cx + cy * cy + cz * cz);
  T[0][0] = cx / ab; T[0][1] = cy / ab; T[0][2] = cz / ab;
  cx = T[0][2] * (Y[2] - Y[0]) - T[0][1] * (Z[2] - Z[0]);
  cy = T[0][0] * (Z[2] - Z[0]) - T[0][2] * (X[2] - X[0]);
  cz = T[0][1] * (X[2] - X[0]) - T[0][0] * (Y[2] - Y[0]);
  ab = sqrt(cx * cx + cy * cy + cz * cz);
  T[1][0] = cx / ab; T[1][1] = cy / ab; T[1][2] = cz / ab;
  cx = T[0][1] * T[1][2] - T[0][2] * T[1][1];
  cy = T[0][2] * T[1][0] - T[0][0] * T[1][2];
  cz = T[0][0] * T[1][1] - T[0][1] * T[1][0];
  ab = sqrt(cx * cx + cy * cy + cz * cz);
  T[2][0] = cx / ab; T[2][1] = cy / ab; T[2][2] = cz / ab;
}

// K <- F^T K F where F = identity with the 3x3 block T at the diagonal positions listed in `blocks`
__device__ void beam_congruence(double K[12][12], const double T[3][3], const bool blocks[4])
{
  double W[12][12];
  // W = K F: column block b of W = K[:, b] . T (if selected)
  for (int i = 0; i < 12; ++i)
    for (int b = 0; b < 4; ++b)
      for (int c = 0; c < 3; ++c) {
        double acc;
        if (blocks[b]) acc = K[i][3 * b] * T[0][c] + K[i][3 * b + 1] * T[1][c] + K[i][3 * b + 2] * T[2][c];
        else acc = K[i][3 * b + c];
        W[i][3 * b + c] = acc;
      }
  // K = F^T W
  for (int b = 0; b < 4; ++b)
    for (int r = 0; r < 3; ++r)
      for (int j = 0; j < 12; ++j) {
        double acc;
        if (blocks[b]) acc = T[0][r] * W[3 * b][j] + T[1][r] * W[3 * b + 1][j] + T[2][r] * W[3 * b + 2][j];
        else acc = W[3 * b + r][j];
        K[3 * b + r][j] = acc;
      }
}

__global__ void build_beam_ops_kernel(int nelt, const double* __restrict__ beam /* [nelt][32] gathered */,
                                      BeamOp* __restrict__ ops, unsigned char* __restrict__ failed)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nelt) return;
  const double* b = beam + (size_t)i * FSR_NBEAM;
  double X[5], Y[5], Z[5], BS[14];
  for (int k = 0; k < 5; ++k) { X[k] = b[k]; Y[k] = b[5 + k]; Z[k] = b[10 + k]; }
  for (int k = 0; k < 14; ++k) BS[k] = b[15 + k];
  const int ipa = (int)b[29], ipb = (int)b[30];
  BeamOp& op = ops[i];
  for (int r = 0; r < 12; ++r)
    for (int c = 0; c < 12; ++c) op.S[r][c] = 0.0;
  failed[i] = 1;

  // ---- BEAM31 ----
  const double E = BS[1], G = BS[2], A = BS[3], RIY = BS[4], RIZ = BS[5], RIT = BS[6];
  const double CAY = BS[8], CAZ = BS[9];
  double YS = BS[10], ZS = BS[11];
  const double efflen = BS[12], phi = BS[13];
  double bx = X[1] - X[0], by = Y[1] - Y[0], bz = Z[1] - Z[0];
  const double len = sqrt(bx * bx + by * by + bz * bz);
  const double BL = efflen > 0.0 ? efflen : len;
  const double ba = 1.0e-6 * (fabs(X[0]) + fabs(X[1]) + fabs(Y[0]) + fabs(Y[1]) + fabs(Z[0]) + fabs(Z[1]));
  if (BL - ba <= 0.0) return;
  bx = X[2] - X[0]; by = Y[2] - Y[0]; bz = Z[2] - Z[0];
  if (sqrt(bx * bx + by * by + bz * bz) - ba <= 0.0) return;
  if (G <= 1.0e-16 || A <= 1.0e-16) return;
  double T0[3][3], T2[3][3];
  beam_dcos(T0, X, Y, Z);
  for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) T2[r][c] = T0[r][c];
  if (fabs(phi) > 1.0e-6) {
    const double fi = phi * atan(1.0) / 4.5e1, cf = cos(fi), sf = sin(fi);
    for (int c = 0; c < 3; ++c) {
      const double a = cf * T2[1][c] + sf * T2[2][c], bb = cf * T2[2][c] - sf * T2[1][c];
      T2[1][c] = a; T2[2][c] = bb;
    }
    const double a = cf * YS + sf * ZS, bb = cf * ZS - sf * YS;
    YS = a; ZS = bb;
  }
  // ---- BELS31: local stiffness ----
  double EA, EIY, EIZ, ALY, ALZ, GIT;
  if (E < -1.0e-16) {
    EA = A; EIY = RIY; EIZ = RIZ; GIT = RIT;
    ALY = CAY > 1.0e-16 ? 12.0 * EIY / (CAY * BL * BL) : 0.0;
    ALZ = CAZ > 1.0e-16 ? 12.0 * EIZ / (CAZ * BL * BL) : 0.0;
  } else {
    EA = E * A; EIY = E * RIY; EIZ = E * RIZ; GIT = G * RIT;
    ALY = 12.0 * CAY * EIY / (A * G * BL * BL);
    ALZ = 12.0 * CAZ * EIZ / (A * G * BL * BL);
  }
  double K[12][12];
  for (int r = 0; r < 12; ++r) for (int c = 0; c < 12; ++c) K[r][c] = 0.0;
  const double k11 = EA / BL, k22 = 12. * EIY / (BL * BL * BL * (1. + ALY)), k33 = 12. * EIZ / (BL * BL * BL * (1. + ALZ));
  const double k44 = GIT / BL, k35 = -.5 * BL * k33, k26 = .5 * BL * k22;
  const double k55 = EIZ * (4. + ALZ) / (BL * (1. + ALZ)), k66 = EIY * (4. + ALY) / (BL * (1. + ALY));
  K[0][0] = k11; K[1][1] = k22; K[2][2] = k33; K[3][3] = k44; K[2][4] = k35; K[1][5] = k26; K[4][4] = k55; K[5][5] = k66;
  K[0][6] = -k11; K[6][6] = k11; K[1][7] = -k22; K[5][7] = -k26; K[7][7] = k22; K[2][8] = -k33; K[4][8] = -k35;
  K[8][8] = k33; K[3][9] = -k44; K[9][9] = k44; K[2][10] = k35; K[4][10] = EIZ * (2. - ALZ) / (BL * (1. + ALZ));
  K[8][10] = -k35; K[10][10] = k55; K[1][11] = k26; K[5][11] = EIY * (2. - ALY) / (BL * (1. + ALY));
  K[7][11] = -k26; K[11][11] = k66;
  for (int r = 0; r < 12; ++r) for (int c = 0; c < r; ++c) K[r][c] = K[c][r];
  if ((double)1.0e-5f * sqrt(A) - (fabs(YS) + fabs(ZS)) < 0.0) {  // single or no symmetry: shear-centre coupling
    for (int c = 0; c < 12; ++c) {
      K[3][c] = K[3][c] - ZS * K[1][c] + YS * K[2][c];
      K[9][c] = K[9][c] - ZS * K[7][c] + YS * K[8][c];
    }
    for (int r = 0; r < 12; ++r) {
      K[r][3] = K[r][3] - K[r][1] * ZS + K[r][2] * YS;
      K[r][9] = K[r][9] - K[r][7] * ZS + K[r][8] * YS;
    }
  }
  // ---- to global axes (MPRO30, or MATTRA on the un-pinned end only) ----
  bool blocks[4] = {false, false, false, false};
  if (ipa <= 0 && ipb <= 0) blocks[0] = blocks[1] = blocks[2] = blocks[3] = true;
  else if (ipa <= 0) blocks[0] = blocks[1] = true;
  else if (ipb <= 0) blocks[2] = blocks[3] = true;
  beam_congruence(K, T2, blocks);
  // ---- end points -> eccentric nodes (TRIX30) ----
  double e1[3] = {X[3] - X[0], Y[3] - Y[0], Z[3] - Z[0]}, e2[3] = {X[4] - X[1], Y[4] - Y[1], Z[4] - Z[1]};
  if (ipa > 0) {
    const double w[3] = {T2[0][0] * e1[0] + T2[0][1] * e1[1] + T2[0][2] * e1[2], T2[1][0] * e1[0] + T2[1][1] * e1[1] + T2[1][2] * e1[2],
                         T2[2][0] * e1[0] + T2[2][1] * e1[1] + T2[2][2] * e1[2]};
    e1[0] = w[0]; e1[1] = w[1]; e1[2] = w[2];
  }
  if (ipb > 0) {
    const double w[3] = {T2[0][0] * e2[0] + T2[0][1] * e2[1] + T2[0][2] * e2[2], T2[1][0] * e2[0] + T2[1][1] * e2[1] + T2[1][2] * e2[2],
                         T2[2][0] * e2[0] + T2[2][1] * e2[1] + T2[2][2] * e2[2]};
    e2[0] = w[0]; e2[1] = w[1]; e2[2] = w[2];
  }
  for (int c = 0; c < 12; ++c) {
    K[3][c] = K[3][c] + e1[2] * K[1][c] - e1[1] * K[2][c];
    K[4][c] = K[4][c] - e1[2] * K[0][c] + e1[0] * K[2][c];
    K[5][c] = K[5][c] + e1[1] * K[0][c] - e1[0] * K[1][c];
    K[9][c] = K[9][c] + e2[2] * K[7][c] - e2[1] * K[8][c];
    K[10][c] = K[10][c] - e2[2] * K[6][c] + e2[0] * K[8][c];
    K[11][c] = K[11][c] + e2[1] * K[6][c] - e2[0] * K[7][c];
  }
  for (int r = 0; r < 12; ++r) {
    K[r][3] = K[r][3] + e1[2] * K[r][1] - e1[1] * K[r][2];
    K[r][4] = K[r][4] - e1[2] * K[r][0] + e1[0] * K[r][2];
    K[r][5] = K[r][5] + e1[1] * K[r][0] - e1[0] * K[r][1];
    K[r][9] = K[r][9] + e2[2] * K[r][7] - e2[1] * K[r][8];
    K[r][10] = K[r][10] - e2[2] * K[r][6] + e2[0] * K[r][8];
    K[r][11] = K[r][11] + e2[1] * K[r][6] - e2[0] * K[r][7];
  }
  // ---- STR11: section forces as a linear operator on v_e ----
  const double ex = X[3] - X[0], ey = Y[3] - Y[0], ez = Z[3] - Z[0];
  for (int c = 0; c < 12; ++c) {
    double Sg[6];
    for (int r = 0; r < 6; ++r) Sg[r] = K[r][c];
    Sg[3] = Sg[3] - ez * Sg[1] + ey * Sg[2];
    Sg[4] = Sg[4] + ez * Sg[0] - ex * Sg[2];
    Sg[5] = Sg[5] - ey * Sg[0] + ex * Sg[1];
    double SN[3], SM[3];
    for (int r = 0; r < 3; ++r) {
      SN[r] = T0[r][0] * Sg[0] + T0[r][1] * Sg[1] + T0[r][2] * Sg[2];
      SM[r] = T0[r][0] * Sg[3] + T0[r][1] * Sg[4] + T0[r][2] * Sg[5];
    }
    // BSEC(10), BSEC(11) as the reference reads them after BEAM31 (XS(1) possibly rotated by PHI)
    const double sf[6] = {-SN[0], SN[1], SN[2], -SM[0] + CAZ * SN[2] - YS * SN[1], SM[1], SM[2]};
    for (int r = 0; r < 6; ++r) { op.S[r][c] = sf[r]; op.S[6 + r][c] = sf[r]; }
    op.S[6 + 4][c] += SN[2] * len;
    op.S[6 + 5][c] += SN[1] * len;
  }
  failed[i] = 0;
}

// one thread per (beam, section-force row): sres[24*elem + row] for the step in column 0 of U
__global__ void beam_apply_kernel(int nelt, const BeamOp* __restrict__ ops, const int* __restrict__ edof,
                                  const int* __restrict__ elem, const unsigned char* __restrict__ failed,
                                  const double* __restrict__ U, size_t ldu, double* __restrict__ sres)
{
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= nelt * 12) return;
  const int i = idx / 12, r = idx % 12;
  double acc = 0.0;
  for (int c = 0; c < 12; ++c) acc += ops[i].S[r][c] * U[(size_t)edof[i * 12 + c] * ldu];
  sres[(size_t)24 * elem[i] + r] = failed[i] ? kHuge : acc;
}

int build_beam_operators(fsr_part* p, const fsr_sam* sam, const fsr_elmdata* elm)
{
  FamilyData& f = p->fam[FAM_BEAM];
  f.nenod = 2; f.nndof = 6; f.nstrp = 0; f.ncmp = 1; f.MT = 0; f.KT = 3;
  std::vector<int> elem, edof;
  std::vector<double> bd;
  for (int e = 0; e < sam->nel; ++e) {
    if (sam->melcon[e] != 11) continue;
    if (elm->elmid && elm->elmid[e] < 1) continue;
    const int ip0 = sam->mpmnpc[e] - 1, nn = sam->mpmnpc[e + 1] - sam->mpmnpc[e];
    if (nn != 2) { set_error("beam element %d has %d nodes", e + 1, nn); return FSR_ERR_ARG; }
    if (!elm->beam) { set_error("part has beam elements but fsr_elmdata.beam is NULL"); return FSR_ERR_ARG; }
    elem.push_back(e);
    for (int k = 0; k < 2; ++k) {
      const int n = sam->mmnpc[ip0 + k] - 1;
      if (n < 0 || n >= sam->nnod) { set_error("element %d: node index out of range", e + 1); return FSR_ERR_ARG; }
      const int js = sam->madof[n] - 1, nd = sam->madof[n + 1] - sam->madof[n];
      if (nd < 6) { set_error("beam element %d: node %d has %d DOFs, beam needs 6", e + 1, n + 1, nd); return FSR_ERR_ARG; }
      for (int d = 0; d < 6; ++d) edof.push_back(js + d);
    }
    bd.insert(bd.end(), elm->beam + (size_t)FSR_NBEAM * e, elm->beam + (size_t)FSR_NBEAM * (e + 1));
  }
  f.nelt = (int)elem.size();
  if (f.nelt == 0) return FSR_OK;
  cudaStream_t s = p->stream;
  double* d_bd = nullptr;
  FSR_CUDA(cudaMalloc(&f.elem, sizeof(int) * elem.size()));
  FSR_CUDA(cudaMalloc(&f.edof, sizeof(int) * edof.size()));
  FSR_CUDA(cudaMalloc(&f.failed, f.nelt));
  FSR_CUDA(cudaMalloc(&f.Sfrag, sizeof(BeamOp) * (size_t)f.nelt));  // plain 12x12 operators, not MMA fragments
  FSR_CUDA(cudaMalloc(&d_bd, sizeof(double) * bd.size()));
  FSR_CUDA(cudaMemcpyAsync(f.elem, elem.data(), sizeof(int) * elem.size(), cudaMemcpyHostToDevice, s));
  FSR_CUDA(cudaMemcpyAsync(f.edof, edof.data(), sizeof(int) * edof.size(), cudaMemcpyHostToDevice, s));
  FSR_CUDA(cudaMemcpyAsync(d_bd, bd.data(), sizeof(double) * bd.size(), cudaMemcpyHostToDevice, s));
  build_beam_ops_kernel<<<(f.nelt + 63) / 64, 64, 0, s>>>(f.nelt, d_bd, reinterpret_cast<BeamOp*>(f.Sfrag), f.failed);
  FSR_LAUNCH_CHECK();
  FSR_CUDA(cudaStreamSynchronize(s));
  cudaFree(d_bd);
  return FSR_OK;
}

int launch_beam_full(fsr_part* p, double* sres, cudaStream_t s)
{
  FamilyData& f = p->fam[FAM_BEAM];
  if (f.nelt == 0 || !sres) return FSR_OK;
  beam_apply_kernel<<<(f.nelt * 12 + 127) / 128, 128, 0, s>>>(f.nelt, reinterpret_cast<const BeamOp*>(f.Sfrag), f.edof,
                                                            f.elem, f.failed, p->U, (size_t)p->step_tile, sres);
  FSR_LAUNCH_CHECK();
  return FSR_OK;
}

}  // namespace fsr
