// k3_fatigue.cu -- K3: peak-valley extraction, rainflow counting and Miner damage (placeholder
// until the kernels land; the entry points fail loudly).
#include "common.cuh"
using namespace fsr;
extern "C" {
int fsr_fatigue(int, const double*, int, int, double, const double*, double, int, double*, int*, int*)
{
  set_error("fsr_fatigue: not built yet");
  return FSR_ERR_STATE;
}
int fsr_fatigue_dev(int, const double*, size_t, int, int, double, const double*, double, int, double*, int*,
                    int*, void*)
{
  set_error("fsr_fatigue_dev: not built yet");
  return FSR_ERR_STATE;
}
}
