// fatigue_core_host.cpp -- TEST INFRASTRUCTURE.  Compiles the K3 state machines of the product
// (fedem_solvers_b200/csrc/fatigue_core.cuh, the code the CUDA kernels run per thread) with g++ so
// that tests/ can fuzz them against the reference's own C++ (oracle/_ref) on millions of samples
// without a GPU.  It is never loaded by the product package.
#include <vector>
#include "../fedem_solvers_b200/csrc/fatigue_core.cuh"

extern "C" int core_fatigue(const double* data, int n, double gate, const double* curve, double bin_size,
                            int nbins, int cap, double* damage, int* ncycles, int* bins, double* ranges,
                            int* first_out, double* turns, int* nturns)
{
  using namespace fsr;
  FatigueParams p;
  p.gate = gate; p.loga1 = curve[0]; p.loga2 = curve[1]; p.m1 = curve[2]; p.m2 = curve[3];
  p.logN0 = (p.m2 * p.loga1 - p.m1 * p.loga2) / (p.m2 - p.m1);
  p.bin_size = bin_size; p.nbins = nbins;
  std::vector<double> edges(nbins + 2, 0.0);
  for (int k = 1; k <= nbins; ++k) edges[k] = edges[k - 1] + bin_size;
  for (int k = 0; k < nbins; ++k) bins[k] = 0;
  PvxLocate loc; loc.init();
  for (int i = 0; i < n; ++i) if (loc.feed(data[i], gate)) break;
  *first_out = loc.first;
  std::vector<double> A(cap > 0 ? cap : 1), B(cap + 8);
  PvxStream pv; pv.init();
  Rainflow rf; rf.init();
  CycleSink sink; sink.init();
  int nr = 0, nt = 0;
  auto count = [&](double a, double b) {
    count_cycle(a, b, p, sink, nbins > 0 ? bins : nullptr, 1, edges.data());
    if (ranges) ranges[nr++] = fabs(a - b);
  };
  auto emit = [&](double v) {
    if (turns) turns[nt++] = v;
    rf.push(v, gate, A.data(), 1, cap, count);
  };
  if (loc.first >= 0)
    for (int i = loc.first; i < n; ++i) pv.feed(i, loc.first, data[i], gate, emit);
  pv.finish(gate, emit);
  int ok = rf.overflow ? 0 : rainflow_finish(rf, gate, A.data(), B.data(), 1, count);
  *damage = sink.damage; *ncycles = sink.ncycles; *nturns = nt;
  for (int k = 0; k < nbins; ++k)
    if (sink.ncycles == 0 || edges[k] > sink.max_range) bins[k] = -1;
  if (rf.overflow) return -2;
  return ok;
}
