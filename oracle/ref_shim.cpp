// ref_shim.cpp -- TEST INFRASTRUCTURE (see oracle.h).  A thin extern "C" face over the
// reference's OWN, unmodified C++ sources (compiled where they lie under /root/reference by
// oracle/Makefile into oracle/_ref/libfedem_ref.so): tensor invariants
// (FFaLib/FFaAlgebra/FFaTensorTransforms.C, FFaMath.C) and fatigue
// (FFpLib/FFpFatigue/FFpFatigue.C, FFpCycle.C, FFpSNCurve.C).  It reproduces the call sequence
// of ffp_getdamage / ffp_getnumcycles (FFpFatigue_F.C:81-141) without the FFaCmdLineArg and
// S-N curve library singletons those wrappers drag in.  No reference source is copied.
#include <algorithm>
#include <sstream>
#include <vector>

#include "FFaLib/FFaAlgebra/FFaMath.H"
#include "FFaLib/FFaAlgebra/FFaTensorTransforms.H"
#include "FFpLib/FFpFatigue/FFpFatigue.H"
#include "FFpLib/FFpFatigue/FFpSNCurve.H"
#include "FFpLib/FFpFatigue/FFpSNCurveLib.H"

// FFpFatigue.C references the S-N curve library singleton (file-based curve tables, GUI side).
// It is never reached from the ffp_getdamage path; satisfy the linker with inert definitions.
FFpSNCurveLib::~FFpSNCurveLib() {}
bool FFpSNCurveLib::readSNCurves(const std::string&) { return false; }
FFpSNCurve* FFpSNCurveLib::getCurve(long int, long int) const { return NULL; }

static std::vector<FFpCycle> g_cycles; // sorted cycles of the last ref_get_damage call

static void cycle_ends(const FFpCycle& c, double& a, double& b)
{
  std::ostringstream os;
  os.precision(17);
  os << c; // "first second" (FFpCycle.C operator<<)
  std::istringstream is(os.str());
  is >> a >> b;
}

extern "C" {

double ref_von_mises(int N, const double* S) { return FFaTensorTransforms::vonMises(N, S); }

int ref_principal_values(int N, const double* S, double* P)
{
  return FFaTensorTransforms::principalValues(N, S, P) ? 1 : 0;
}

double ref_max_shear_value(double pmax, double pmin)
{
  return FFaTensorTransforms::maxShearValue(pmax, pmin);
}

void ref_rotate2d(const double* S, const double* T, double* out)
{
  FFaTensorTransforms::rotate2D(S, T, out);
}

int ref_cubic_solve(double A, double B, double C, double D, double* X)
{
  return FFa::cubicSolve(A, B, C, D, X);
}

// FFpPVXprocessor::process(times,data,...) exactly as ffp_getdamage drives it
int ref_pvx(const double* data, int n, double gate, double* turns)
{
  std::vector<double> times(n > 0 ? n : 1);
  for (int i = 0; i < n; i++) times[i] = i;
  FFpPVXprocessor pvx(gate);
  std::vector<FFpPoint> tp;
  pvx.process(&times.front(), data, n, tp, true);
  for (size_t i = 0; i < tp.size(); i++) turns[i] = tp[i].second;
  return (int)tp.size();
}

// FFpRainFlowCycleCounter::process(turns,cycles,true); returns -1 on the reference's failure
int ref_rainflow(const double* turns, int nturns, double gate, double* cf, double* cs)
{
  std::vector<FFpPoint> tp(nturns);
  for (int i = 0; i < nturns; i++) tp[i] = FFpPoint((double)i, turns[i]);
  FFpCycles cycles;
  FFpRainFlowCycleCounter cyc(gate);
  bool ok = cyc.process(tp, cycles, true);
  for (size_t i = 0; i < cycles.size(); i++) cycle_ends(cycles[i], cf[i], cs[i]);
  return ok ? (int)cycles.size() : -1;
}

// ffp_getdamage (FFpFatigue_F.C:81-124): PVX -> rainflow -> sort -> Miner sum on a NorSok curve
double ref_get_damage(const double* data, int n, double gate, const double* curve)
{
  std::vector<double> times(n > 0 ? n : 1);
  for (int i = 0; i < n; i++) times[i] = i;
  FFpPVXprocessor pvx(gate);
  std::vector<FFpPoint> turns;
  pvx.process(&times.front(), data, n, turns, true);
  g_cycles.clear();
  FFpRainFlowCycleCounter cyc(gate);
  cyc.process(turns, g_cycles, true);
  std::sort(g_cycles.begin(), g_cycles.end());
  FFpSNCurveNorSok snCurve(curve[0], curve[1], curve[2], curve[3]);
  return FFpFatigue::getDamage(g_cycles, snCurve);
}

// ffp_getnumcycles (FFpFatigue_F.C:127-141) on the cycles of the last ref_get_damage call
int ref_get_num_cycles(double low, double high)
{
  if (g_cycles.empty()) return -1;
  const std::vector<FFpCycle>& cyc = g_cycles;
  std::vector<FFpCycle>::const_iterator ilow, ihigh;
  ilow = std::lower_bound(cyc.begin(), cyc.end(), low);
  if (ilow == cyc.end()) return -1;
  ihigh = std::lower_bound(ilow, cyc.end(), high);
  return (int)(ihigh - ilow);
}

int ref_num_cycles_total() { return (int)g_cycles.size(); }

double ref_sn_norsok(double s, double loga1, double loga2, double m1, double m2)
{
  FFpSNCurveNorSok c(loga1, loga2, m1, m2);
  return c.getValue(s);
}
}
