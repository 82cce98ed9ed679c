// fatigue_core.cuh -- K3 per-series state machines: first-turning-point search, gated peak-valley
// extraction (PVX), 4-point rainflow counting, Miner damage on a two-slope S-N curve.
//
// Replaces, for ONE scalar history, the reference's
//   FFpPVXprocessor::locateFirstTP / ::process   (fedem-foundation/src/FFpLib/FFpFatigue/FFpFatigue.C:77-163)
//   FFpRainFlowCycleCounter::processTPList / ::processFinish              (FFpFatigue.C:201-320)
//   FFpFatigue::getDamage + FFpSNCurveNorSok::getValue       (FFpFatigue.C:381-396, FFpSNCurve.C:10-33)
//   ffp_getnumcycles bin counts                                               (FFpFatigue_F.C:127-141)
// The reference keeps every sample, every turning point (std::vector) and a std::list it sweeps
// repeatedly; here everything is STREAMING so that one GPU thread owns one gage and the history
// never has to exist as a whole (config 5: 4e5 series x 1e5 steps = 320 GB):
//   * PvxLocate   -- locateFirstTP with values instead of indices (same branch order, same
//                    quirks: initial gradient = data[0], stale iTP after a min/max registration);
//   * PvxStream   -- the main loop of process(); it starts at the first turning point, which is
//                    why the driver runs the (early-exiting) locate pass first;
//   * Rainflow    -- the 4-point rules applied to the top of a stack as each turning point
//                    arrives.  The reference's sweeps repeat until no rule applies anywhere; the
//                    rules are confluent in the multiset of counted ranges and in the residue
//                    values (a tie |r0| == |r1| only decides which of two equal-range pairs is
//                    taken), so the stack formulation counts the same cycles.  tests/ checks this
//                    against the reference's own compiled C++ on adversarial series (plateaus,
//                    ties, monotone, below-gate, growing/shrinking sawtooth);
//   * the residue closure of processFinish: rotate at the first max-|value| point, re-run, the
//                    last three points give the final (ungated) cycle, anything else is the
//                    reference's failure return (cycles counted so far are kept, as ffp_getdamage
//                    ignores the return value).
// Host+device header: the CUDA kernels (k3_fatigue.cu) are the product; tests/ additionally
// compile this header with g++ to fuzz the state machines against the reference on the CPU.
#pragma once
#include <math.h>

#ifdef __CUDACC__
#define FSR_HD __host__ __device__ __forceinline__
#else
#define FSR_HD inline
#endif

namespace fsr {

struct FatigueParams {
  double gate;
  double loga1, loga2, m1, m2, logN0;  // logN0 = (m2*loga1 - m1*loga2)/(m2 - m1), FFpSNCurve.C:12-15
  double bin_size;                     // reportDamage bins [k*binSize, (k+1)*binSize), 0 = no bins
  int nbins;
};

// FFpSNCurveNorSok::getValue (FFpSNCurve.C:22-33): cycles to failure at stress range s
FSR_HD double sn_norsok(double s, const FatigueParams& p)
{
  double logN = p.loga1 - p.m1 * log10(s);
  if (logN < p.logN0) return pow(10.0, logN);
  logN = p.loga2 - p.m2 * log10(s);
  return pow(10.0, logN);
}

// Damage of one cycle, 1 / getValue(s) (FFpFatigue.C:381-396).  On the device the general pow(10, x) (an extended-precision
// log + exp, ~3x the instructions of everything else a cycle costs) is replaced by exp10(-x): one logarithm, one exponential,
// no division.  The branch is decided on the same logN as the reference; the value differs from 1 / pow(10, logN) by a few
// ulp (every term of the Miner sum is positive, so the sum keeps that relative error; the tests hold 1e-10).
FSR_HD double sn_norsok_damage(double s, const FatigueParams& p)
{
#ifdef __CUDA_ARCH__
  const double ls = log10(s);
  double logN = p.loga1 - p.m1 * ls;
  if (!(logN < p.logN0)) logN = p.loga2 - p.m2 * ls;
  return exp10(-logN);
#else
  return 1.0 / sn_norsok(s, p);
#endif
}

// ---- locateFirstTP (FFpFatigue.C:129-163), one sample at a time -----------------------------
struct PvxLocate {
  double deltaTP, vTP, vMin, vMax, xprev;
  int iTP, iMin, iMax, n;  // n = samples consumed so far
  int first;               // >= 0: index of the first turning point; -1: not found (yet)

  FSR_HD void init()
  {
    n = 0; first = -1; iTP = iMin = iMax = 0;
    deltaTP = vTP = vMin = vMax = xprev = 0.0;
  }

  // returns true once the first turning point is known
  FSR_HD bool feed(double x, double gate)
  {
    if (first >= 0) return true;
    if (n == 0) {
      deltaTP = x;  // sic: the reference seeds the gradient with data[0]
      vTP = vMin = vMax = x;
    } else if ((x - vTP) * deltaTP > 0.0) {
      iTP = n; vTP = x;
    } else if (xprev - vMin > gate)
      first = iMin;
    else if (vMax - xprev > gate)
      first = iMax;
    else if (vTP > vMax) {
      iMax = iTP; vMax = vTP; deltaTP = x - vTP;
    } else if (vTP < vMin) {
      iMin = iTP; vMin = vTP; deltaTP = x - vTP;
    } else {
      iTP = n; vTP = x; deltaTP = 0.0;  // deltaTP = data[i] - data[iTP] with iTP = i
    }
    xprev = x;
    ++n;
    return first >= 0;
  }
};

// ---- rainflow stack: the top three points live in registers, the rest in a caller-provided spill
// array (element k of this series at spill[k*stride]) ---------------------------------------------
struct CycleSink {
  double damage, max_range;
  int ncycles;
  FSR_HD void init() { damage = 0.0; max_range = 0.0; ncycles = 0; }
};

struct Rainflow {
  double s0, s1, s2;  // s2 = top of stack
  int n;              // points on the stack (registers + spill)
  int overflow;       // spill capacity exceeded

  FSR_HD void init() { s0 = s1 = s2 = 0.0; n = 0; overflow = 0; }

  FSR_HD double get(int j, const double* spill, size_t stride) const
  {
    const int nreg = n < 3 ? n : 3;
    const int base = n - nreg;  // number of spilled points
    if (j < base) return spill[(size_t)j * stride];
    const int r = j - base + (3 - nreg);  // 0 -> s0, 1 -> s1, 2 -> s2
    return r == 0 ? s0 : (r == 1 ? s1 : s2);
  }

  // registers hold the top min(n,3) points right-aligned: n=1: s2; n=2: s1,s2; n>=3: s0,s1,s2
  FSR_HD void append(double v, double* spill, size_t stride, int cap)
  {
    if (n >= 3) {
      if (n - 3 >= cap) { overflow = 1; return; }
      spill[(size_t)(n - 3) * stride] = s0;
    }
    s0 = s1; s1 = s2; s2 = v;
    ++n;
  }

  // drop `k` (1 or 2) of the register-held points below the top element(s) that survive;
  // `keep_lo`, `keep_hi` are the survivors in stack order.  Refill from the spill.
  template <class Count>
  FSR_HD void push(double v, double gate, double* spill, size_t stride, int cap, Count&& count)
  {
    while (n >= 3) {
      const double r0 = s1 - s0, r1 = s2 - s1, r2 = v - s2;
      if (r0 * r1 > 0.0) {            // point 1 is not a turning point (FFpFatigue.C:227-237)
        s1 = s0;                      // survivors: s0, s2
        --n;
        s0 = n >= 3 ? spill[(size_t)(n - 3) * stride] : 0.0;
      } else if (r1 * r2 > 0.0) {     // point 2 is not a turning point (:238-247)
        s2 = s1; s1 = s0;             // survivors: s0, s1
        --n;
        s0 = n >= 3 ? spill[(size_t)(n - 3) * stride] : 0.0;
      } else if (fabs(r0) >= fabs(r1) && fabs(r2) >= fabs(r1)) {  // a genuine cycle (:248-264)
        if (fabs(r1) > gate) count(s1, s2);
        s2 = s0;                      // survivor: s0
        n -= 2;
        s1 = n >= 2 ? spill[(size_t)(n - 2) * stride] : 0.0;
        s0 = n >= 3 ? spill[(size_t)(n - 3) * stride] : 0.0;
      } else
        break;
    }
    append(v, spill, stride, cap);
  }
};

// ---- PVX main loop (FFpFatigue.C:86-106 with isFirstData handled by the caller) ----------------
struct PvxStream {
  double P, D, lastTP;  // myPossibleTP.second, myDeltaTP, turns.back()
  int started;          // 0: before the first turning point, 1: first TP taken, gradient pending, 2: running

  FSR_HD void init() { P = D = lastTP = 0.0; started = 0; }

  // i = global sample index, first = index of the first turning point.  emit(v) receives each
  // turning point in order.
  template <class Emit>
  FSR_HD void feed(int i, int first, double x, double gate, Emit&& emit)
  {
    if (i < first) return;
    if (i == first) {
      P = x; lastTP = x; started = 1;
      emit(x);
      return;
    }
    if (started == 1) { D = x - P; started = 2; }  // myDeltaTP = data[iFirst+1] - data[iFirst]
    const double delta = x - P;
    if (delta * D <= 0.0) {
      if (fabs(delta) > gate) { emit(P); lastTP = P; }
      else return;  // ignore small ranges
    }
    P = x;
    D = x - lastTP;
  }

  // end of data: the last possible turning point (FFpFatigue.C:108-109)
  template <class Emit>
  FSR_HD void finish(double gate, Emit&& emit)
  {
    if (started >= 1 && fabs(D) > gate) emit(P);
  }
};

// ---- per-cycle accounting: Miner sum, counts, histogram -------------------------------------
// bins: this series' histogram, element k at bins[k*bstride]; edges[0..nbins] are the bin limits
// accumulated exactly like reportDamage does (strainGageModule.f90:827-848: s0 = s1; s1 = s0 +
// binSize); a cycle belongs to bin k when edges[k] <= range < edges[k+1] (lower_bound semantics of
// ffp_getnumcycles, FFpFatigue_F.C:127-141).
FSR_HD void count_cycle(double a, double b, const FatigueParams& p, CycleSink& sink, int* bins, size_t bstride,
                        const double* edges)
{
  const double range = fabs(a - b);  // FFpCycle::range with toMPaScale = 1
  ++sink.ncycles;
  if (range > sink.max_range) sink.max_range = range;
  sink.damage += sn_norsok_damage(range, p);
  if (bins && p.nbins > 0) {
    double q = range / p.bin_size;
    int k = q < (double)p.nbins ? (int)q : p.nbins;
    while (k > 0 && range < edges[k]) --k;
    while (k < p.nbins && range >= edges[k + 1]) ++k;
    if (k < p.nbins) ++bins[(size_t)k * bstride];
  }
}

// ---- processTPList (FFpFatigue.C:201-271) as ONE literal left-to-right sweep over an array -----
// a[0..m) with element j at a[j*stride]; compacts in place, returns the new length, adds the number
// of removed points to *removed.  Same window moves as the reference's four list iterators:
// rule A drops p1, rule B drops p2, rule C counts (p1,p2) and continues with (p0,p3,p4,p5), else
// the window slides by one.  Used for the residue closure, where the sweep ORDER matters as soon as
// the list holds equal neighbours (zero ranges), e.g. the duplicated first turning point that
// FFpPVXprocessor emits for a history that starts on a plateau.
template <class Count>
FSR_HD int rainflow_sweep(double* a, size_t stride, int m, double gate, int* removed, Count&& count)
{
  if (m < 4) return m;
  int out = 0, rd = 4;
  double w0 = a[0], w1 = a[stride], w2 = a[2 * stride], w3 = a[3 * stride];
  int nw = 4;  // valid window entries
  while (nw == 4) {
    const double r0 = w1 - w0, r1 = w2 - w1, r2 = w3 - w2;
    if (r0 * r1 > 0.0) {
      ++*removed;
      w1 = w2; w2 = w3;
      if (rd < m) w3 = a[(size_t)rd++ * stride]; else nw = 3;
    } else if (r1 * r2 > 0.0) {
      ++*removed;
      w2 = w3;
      if (rd < m) w3 = a[(size_t)rd++ * stride]; else nw = 3;
    } else if (fabs(r0) >= fabs(r1) && fabs(r2) >= fabs(r1)) {
      *removed += 2;
      if (fabs(r1) > gate) count(w1, w2);
      w1 = w3;
      if (rd + 1 < m) { w2 = a[(size_t)rd * stride]; w3 = a[(size_t)(rd + 1) * stride]; rd += 2; }
      else if (rd < m) { w2 = a[(size_t)rd * stride]; ++rd; nw = 3; }
      else nw = 2;
    } else {
      a[(size_t)out++ * stride] = w0;
      w0 = w1; w1 = w2; w2 = w3;
      if (rd < m) w3 = a[(size_t)rd++ * stride]; else nw = 3;
    }
  }
  a[(size_t)out++ * stride] = w0;
  if (nw >= 2) a[(size_t)out++ * stride] = w1;
  if (nw >= 3) a[(size_t)out++ * stride] = w2;
  return out;
}

// ---- processFinish (FFpFatigue.C:274-320) ------------------------------------------------------
// `rf` holds the residue of the streaming phase (spill A); spill B (cap+4 entries) receives the
// rotated list v[kmax..n-1], v[0..kmax-1], v[kmax] and is swept literally until nothing changes.
// Returns 1 = ok, 0 = the reference's failure return (not exactly three points left).
template <class Count>
FSR_HD int rainflow_finish(Rainflow& rf, double gate, const double* spillA, double* spillB, size_t stride,
                           Count&& count)
{
  const int n = rf.n;
  if (n <= 1) return 1;
  int kmax = 0;
  double vmax = fabs(rf.get(0, spillA, stride));
  for (int j = 1; j < n; ++j) {
    const double a = fabs(rf.get(j, spillA, stride));
    if (a > vmax) { vmax = a; kmax = j; }
  }
  int m = 0;
  for (int j = kmax; j < n; ++j) spillB[(size_t)m++ * stride] = rf.get(j, spillA, stride);
  for (int j = 0; j < kmax; ++j) spillB[(size_t)m++ * stride] = rf.get(j, spillA, stride);
  spillB[(size_t)m++ * stride] = rf.get(kmax, spillA, stride);
  for (;;) {
    int removed = 0;
    m = rainflow_sweep(spillB, stride, m, gate, &removed, count);
    if (removed == 0) break;
  }
  if (m != 3) return 0;
  count(spillB[0], spillB[stride]);  // the last cycle is counted without a gate check (:316-317)
  return 1;
}

}  // namespace fsr
