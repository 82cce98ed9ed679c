// k2_full.cu -- the full per-step result set (one step, every derived measure), generic over the
// element families.  This is the B200 counterpart of what fedem_stress computes per element when
// all of -SR -stress -strain -vmStress -maxPStress ... are switched on
// (reference src/vpmStress/stressRoutines.f90:234-331): stress and strain tensors at every result
// point, von Mises / max & min principal / max shear of both (calcVonMises, calcPrincipalVals,
// src/vpmStress/strainAndStressUtils.f90:484-555 -> FFaTensorTransforms.C:33-67,242-296 ->
// FFa::cubicSolve, FFaMath.C:61-142, same branch structure), and the shell stress resultants.
// Not the throughput path (that is the von Mises + envelope kernel of each family); one thread
// per result point, operator rows read back from the A-fragment stream.
#include "common.cuh"
#include "invariants.cuh"

namespace fsr {

__device__ __forceinline__ size_t frag_at(int row, int col, int KT)
{
  return ((size_t)((row >> 3) * KT + (col >> 2)) << 5) + ((row & 7) << 2) + (col & 3);
}

// layout: 0 = shells (row = comp*8 + point), 1 = solids (row = point*ncmp + comp)
__global__ void k2_full_kernel(const double* __restrict__ U, size_t ldu, const double* __restrict__ Sfrag,
                               const int* __restrict__ edof, const int* __restrict__ ptoff,
                               const int* __restrict__ elem, const unsigned char* __restrict__ failed,
                               const double* __restrict__ aux, int naux, const double* __restrict__ Efrag, int nelt, int nstrp, int ncmp,
                               int nedof, int MT, int KT, int layout, int nenod,
                               double* __restrict__ resmat, double* __restrict__ stress,
                               double* __restrict__ strain, double* __restrict__ sres)
{
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= nelt * nstrp) return;
  const int i = idx / nstrp, pnt = idx % nstrp;
  const size_t pt = (size_t)ptoff[i] + pnt;
  const double* S = Sfrag + (size_t)i * MT * KT * 32;
  const int* ed = edof + (size_t)i * KT * 4;
  double sig[6] = {0, 0, 0, 0, 0, 0}, eps[6] = {0, 0, 0, 0, 0, 0};
  if (failed[i]) {
    for (int k = 0; k < 8; ++k) resmat[8 * pt + k] = kHuge;
    for (int k = 0; k < 6; ++k) { stress[6 * pt + k] = kHuge; strain[6 * pt + k] = kHuge; }
    if ((ncmp == 3 && pnt < nenod) || (Efrag && pnt < 4)) for (int k = 0; k < 6; ++k) sres[(size_t)24 * elem[i] + 6 * pnt + k] = kHuge;
    return;
  }
  for (int c = 0; c < ncmp; ++c) {
    const int row = layout == 0 ? c * 8 + pnt : pnt * ncmp + c;
    double s = 0.0;
    for (int col = 0; col < nedof; ++col) s += S[frag_at(row, col, KT)] * U[(size_t)ed[col] * ldu];
    sig[c] = s;
  }
  const double E = aux[(size_t)i * naux], nu = aux[(size_t)i * naux + 1];
  if (Efrag) {
    // thick shells: the strain has its own operator (local inverse constitutive matrix applied before the rotation to the
    // global axes, STR31 / STR32 elStressModule.f90:1148-1160, 1260-1272; tensorial shear already folded in)
    const double* Es = Efrag + (size_t)i * MT * KT * 32;
    for (int c = 0; c < ncmp; ++c) {
      double s = 0.0;
      for (int col = 0; col < nedof; ++col) s += Es[frag_at(pnt * ncmp + c, col, KT)] * U[(size_t)ed[col] * ldu];
      eps[c] = s;
    }
    if (pnt < 4) for (int k = 0; k < 6; ++k) sres[(size_t)24 * elem[i] + 6 * pnt + k] = 0.0;   // SR = 0 "maybe later" (:1174-1176)
  } else if (ncmp == 3) {
    // isoMat2Dinv (isoMatModule.f90:41-57), then tensorial shear (elStressModule.f90:244-248)
    eps[0] = sig[0] / E - nu / E * sig[1];
    eps[1] = -nu / E * sig[0] + sig[1] / E;
    eps[2] = 0.5 * (2.0 * (1.0 + nu) / E * sig[2]);
  } else {
    // isoMat3Dinv (isoMatModule.f90:95-120), then tensorial shear (elStressModule.f90:249-251)
    eps[0] = (sig[0] - nu * (sig[1] + sig[2])) / E;
    eps[1] = (sig[1] - nu * (sig[0] + sig[2])) / E;
    eps[2] = (sig[2] - nu * (sig[0] + sig[1])) / E;
    const double g2 = 2.0 * (1.0 + nu) / E;
    eps[3] = 0.5 * g2 * sig[3]; eps[4] = 0.5 * g2 * sig[4]; eps[5] = 0.5 * g2 * sig[5];
  }
  for (int k = 0; k < 6; ++k) { stress[6 * pt + k] = sig[k]; strain[6 * pt + k] = eps[k]; }
  const int np = ncmp == 3 ? 2 : 3;
  double P[3] = {0, 0, 0};
  resmat[8 * pt + 0] = von_mises(ncmp, sig);
  principal_values(ncmp, sig, P);
  resmat[8 * pt + 1] = P[0]; resmat[8 * pt + 2] = P[np - 1]; resmat[8 * pt + 3] = 0.5 * (P[0] - P[np - 1]);
  resmat[8 * pt + 4] = von_mises(ncmp, eps);
  principal_values(ncmp, eps, P);
  resmat[8 * pt + 5] = P[0]; resmat[8 * pt + 6] = P[np - 1]; resmat[8 * pt + 7] = 0.5 * (P[0] - P[np - 1]);

  // shell stress resultants at node pnt from the top (pnt) and bottom (nenod+pnt) stresses
  // (STR22a, elStressModule.f90:826-831; STR23 :976-979 is the same relation inverted)
  if (ncmp == 3 && pnt < nenod) {
    const double t = aux[(size_t)i * naux + 2];
    double bot[3];
    for (int c = 0; c < 3; ++c) {
      const int row = c * 8 + nenod + pnt;
      double s = 0.0;
      for (int col = 0; col < nedof; ++col) s += S[frag_at(row, col, KT)] * U[(size_t)ed[col] * ldu];
      bot[c] = s;
    }
    double* sr = sres + (size_t)24 * elem[i] + 6 * pnt;
    for (int c = 0; c < 3; ++c) {
      sr[c] = (sig[c] + bot[c]) * 0.5 * t;
      sr[3 + c] = (sig[c] - bot[c]) * 0.5 * t * t / 6.0;
    }
  }
}

int launch_k2_full(fsr_part* p, double* resmat, double* stress, double* strain, double* sres, cudaStream_t s)
{
  for (int fi = 0; fi < FAM_COUNT; ++fi) {
    FamilyData& f = p->fam[fi];
    if (f.nelt == 0 || f.nstrp == 0) continue;
    int total = f.nelt * f.nstrp;
    int layout = (fi == FAM_QUAD || fi == FAM_TRI) ? 0 : 1;
    k2_full_kernel<<<(total + 127) / 128, 128, 0, s>>>(p->U, (size_t)p->step_tile, f.Sfrag, f.edof, f.ptoff,
                                                      f.elem, f.failed, f.aux, f.naux, f.Efrag, f.nelt, f.nstrp, f.ncmp,
                                                      f.nenod * f.nndof, f.MT, f.KT, layout, f.nenod, resmat,
                                                      stress, strain, sres);
    FSR_LAUNCH_CHECK();
  }
  return launch_beam_full(p, sres, s);
}

}  // namespace fsr
