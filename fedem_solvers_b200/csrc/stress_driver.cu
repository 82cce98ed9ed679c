// stress_driver.cu -- the fedem_stress program flow on top of the library's own entry points (host code).
//
// Reference: main() of src/vpmStress/stressmain.C:16-83 (the option table), initSolverArgs / solveStress
// (src/vpmStress/stressInterface.C:86-116 -> cmdLineArgInit, readOptionFiles) and subroutine stress
// (src/vpmStress/stress.f90:17-493).  The reference walks the time steps one by one (readSupElDisplacements,
// calcIntDisplacements, calcStresses, a few hundred small fwrites per step); here the time steps selected by
// -statm/-stotm/-tinc are collected first (ffr_getnextstep semantics), their reduced displacements are read
// and assembled in windows (readSupElDisplacements + BuildFinit batched), and each window makes one trip to the
// GPU (K1 expansion + record kernels) from which the finished .frs records come back.
//
// Exported with the reference's names so that the same launcher / fedempy code can drive it:
//   initSolverArgs(argc, argv), solveStress().
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <chrono>
#include <string>
#include <thread>
#include <unistd.h>
#include <vector>

#include "cmdline.hpp"
#include "common.cuh"

using namespace fsr;

namespace {

CmdLine g_cmd;
bool g_stress_options_defined = false, g_gage_options_defined = false, g_modes_options_defined = false, g_fpp_options_defined = false;

std::string strip_ext_add(const std::string& link, const char* suffix)
{
  if (link.empty()) return std::string("fedem") + suffix;
  const size_t dot = link.rfind('.');
  return (dot == std::string::npos ? link : link.substr(0, dot)) + suffix;
}

// getFileName (src/vpmStress/displacementModule.f90:87-117)
std::string file_name(const char* option, const char* suffix)
{
  std::string name = g_cmd.get_string(option);
  if (name.empty() && suffix) name = strip_ext_add(g_cmd.get_string("linkfile"), suffix);
  return name;
}

// FFaTokenizer(files,'<','>',','): "<a,b,c>" -> a b c
std::vector<std::string> file_list(const std::string& s)
{
  std::vector<std::string> out;
  if (s.empty()) return out;
  if (s[0] != '<') { out.push_back(s); return out; }
  std::string tok;
  for (size_t i = 1; i < s.size() && s[i] != '>'; ++i) {
    if (s[i] == ',') { if (!tok.empty()) out.push_back(tok); tok.clear(); }
    else if (!isspace((unsigned char)s[i]) && s[i] != '"') tok += s[i];
  }
  if (!tok.empty()) out.push_back(tok);
  return out;
}

struct Log {
  FILE* f = nullptr;
  clock_t t0 = clock();
  void open(const std::string& path, const char* heading)
  {
    f = fopen(path.c_str(), "w");
    if (f) fprintf(f, "\n     %s (B200)\n\n", heading);
  }
  void line(const char* fmt, ...)
  {
    va_list ap;
    va_start(ap, fmt);
    if (f) { va_list cp; va_copy(cp, ap); vfprintf(f, fmt, cp); va_end(cp); fputc('\n', f); }
    vprintf(fmt, ap);
    putchar('\n');
    va_end(ap);
  }
  ~Log() { if (f) fclose(f); }
};

}  // namespace

extern "C" {

// ffr_getnextstep over a sorted key list (fedem-foundation/src/FFrLib/FFrExtractorInterface.f90:134-170,
// ffr_setposition = first key >= wanted - FLT_EPSILON, clamped to the ends, FFrResultContainer.C:953-1010):
// the indices of the time steps the stress loop visits.  Returns their number (out may be NULL).
int fsr_select_steps(const double* times, int n, double start, double stop, double tinc, int* out, int cap)
{
  if (n < 0 || (n > 0 && !times)) { set_error("fsr_select_steps: bad arguments"); return FSR_ERR_ARG; }
  const double tol = 1.0e-12, huge = 1.7976931348623157e308, feps = 1.1920928955078125e-07;
  auto position = [&](double wanted) -> int {
    if (n == 0) return -1;
    if (times[0] > wanted) return 0;
    if (wanted > times[n - 1]) return n - 1;
    return (int)(std::upper_bound(times, times + n, wanted - feps) - times);
  };
  int count = 0, idx = -1;
  double curr = start - 1.0, last = -huge;
  for (;;) {
    if (curr > stop - tol) break;
    if (curr < start - tol) {
      idx = position(start);
      if (idx >= 0) { curr = times[idx]; start = curr; }
      last = -huge;
    } else if (tinc < tol) {
      idx = idx + 1 < n ? idx + 1 : -1;
      if (idx >= 0) curr = times[idx];
    } else {
      idx = position(curr + tinc);
      if (idx >= 0) curr = times[idx];
    }
    const bool ok = curr < stop + tol && idx >= 0 && curr > last + tol;
    last = curr;
    if (!ok) break;
    if (out && count < cap) out[count] = idx;
    ++count;
  }
  return count;
}

// ---- the option parser behind a C face (tests compare it with the reference's FFaCmdLineArg) -------------
void fsr_cmdline_reset(void) { g_cmd = CmdLine(); g_stress_options_defined = false; }
void fsr_cmdline_add_bool(const char* n, int v) { g_cmd.add(n, v != 0, ""); }
void fsr_cmdline_add_int(const char* n, int v) { g_cmd.add(n, v, ""); }
void fsr_cmdline_add_double(const char* n, double v) { g_cmd.add(n, v, ""); }
void fsr_cmdline_add_string(const char* n, const char* v) { g_cmd.add(n, v, ""); }
void fsr_cmdline_init(int argc, char** argv) { g_cmd.init(argc, argv); }
int fsr_cmdline_read_file(const char* path) { return g_cmd.read_options_file(path) ? 1 : 0; }
int fsr_cmdline_get_bool(const char* n) { return g_cmd.get_bool(n) ? 1 : 0; }
int fsr_cmdline_get_int(const char* n) { return g_cmd.get_int(n); }
double fsr_cmdline_get_double(const char* n) { return g_cmd.get_double(n); }
int fsr_cmdline_get_string(const char* n, char* out, int cap)
{
  const std::string s = g_cmd.get_string(n);
  if (out && cap > 0) { strncpy(out, s.c_str(), (size_t)cap - 1); out[cap - 1] = 0; }
  return (int)s.size();
}
int fsr_cmdline_is_set(const char* n) { return g_cmd.is_set(n) ? 1 : 0; }

// The option table of fedem_stress: cmdLineArgInitStd + cmdLineArgInit (src/vpmCommon/cmdLineArgInitStd.C:72-90,
// cmdLineArgInit.C:109-113) and stressmain.C:22-79, same names, defaults and help texts' meaning.
void fsr_stress_define_options(void)
{
  CmdLine& c = g_cmd;
  c.add("fao", "", "Read additional options from this file");
  c.add("fco", "", "Read calculation options from this file");
  c.add("fop", "", "Read output options from this file");
  c.add("cwd", "", "Change working directory");
  c.add("help", false, "Print out this help text");
  c.add("helpAll", false, "Print out this help text\nincluding the private options, if any", false);
  c.add("version", false, "Print out program version");
  c.add("debug", 0, "Debug print switch");
  c.add("terminal", 6, "File unit number for terminal output");
  c.add("consolemsg", false, "Output error messages to console");
  c.add("Bramsize", -1, "In-core size (MB) of displacement recovery matrix\n< 0: Use the same as in the reducer (default)\n= 0: Store full matrix in core");
  c.add("dmramsize", -1, "Same as -Bramsize but in terms of double words", false);
  c.add("linkId", 0, "Link base-ID number");
  c.add("linkfile", "", "Name of link input file");
  c.add("Bmatfile", "", "Name of B-matrix file");
  c.add("eigfile", "", "Name of eigenvector file");
  c.add("dispfile", "", "Name of gravitation displacement file");
  c.add("resfile", "", "Name of result output file");
  c.add("samfile", "", "Name of SAM data file");
  c.add("fsifile", "fedem_solver.fsi", "Name of solver input file");
  c.add("resStressFile", "", "Name of residual stress input file");
  c.add("resStressSet", "", "Name of residual stress set");
  c.add("frsfile", "", "Name of solver results database file");
  c.add("rdbfile", "", "Name of stress results database file");
  c.add("rdbinc", 1, "Increment number for the results database file");
  c.add("VTFfile", "", "Name of VTF output file");
  c.add("VTFoffset", 0, "VTF result block id offset");
  c.add("VTFparts", 0, "Number of parts in VTF-file");
  c.add("VTFavgelm", true, "Write averaged element results to VTF-file");
  c.add("VTFinit", false, "Write initial state to VTF-file");
  c.add("VTFdscale", 1.0, "Deformation scaling factor for VTF output");
  c.add("double", false, "Save all results in double precision");
  c.add("group", "", "List of element groups to do calculations for");
  c.add("nodalForces", false, "Compute and print nodal forces");
  c.add("SR", false, "Save stress resultants to results database");
  c.add("stress", false, "Save stress tensors to results database");
  c.add("strain", false, "Save strain tensors to results database");
  c.add("vmStress", false, "Save von Mises stress to results database");
  c.add("vmStrain", false, "Save von Mises strain to results database");
  c.add("maxPStress", false, "Save max principal stress to results database");
  c.add("maxPStrain", false, "Save max principal strain to results database");
  c.add("minPStress", false, "Save min principal stress to results database");
  c.add("minPStrain", false, "Save min principal strain to results database");
  c.add("maxSStress", false, "Save max shear stress to results database");
  c.add("maxSStrain", false, "Save max shear strain to results database");
  c.add("deformation", false, "Save deformations to results database");
  c.add("dumpDefNas", false, "Save deformations to Nastran bulk data files");
  c.add("write_nodes", true, "Save deformations as nodal data");
  c.add("write_vector", false, "Save deformations as vector data");
  c.add("statm", 0.0, "Start time");
  c.add("stotm", 1.0, "Stop time");
  c.add("tinc", 0.1, "Time increment (= 0.0: process all time steps)");
  c.add("stressForm", 0, "General stress formulation option\n= 0: Direct evaluation in nodes\n= 1: Volume averaged or mid-point evaluation\n= 2: Extrapolation from Gauss integration points", false);
  c.add("ffqStressForm", 2, "Stress formulation for the FFQ shell", false);
  c.add("fftStressForm", 1, "Stress formulation for the FFT shell", false);
  c.add("useIncompatibleModes", false, "Linear hexahedron option", false);
  // B200 additions
  c.add("device", 0, "CUDA device ordinal");
  c.add("stepTile", 0, "Time steps per device batch (0 = from free device memory)");
  c.add("gpus", 0, "Number of GPUs the elements of the part are spread over\n= 0: automatic (one per 250,000 elements for an envelope-only run,\n     one per 2,000,000 when a results database is written)");
  g_stress_options_defined = true;
}

// The option table of fedem_gage: the standard options + gagemain.C:22-61.
void fsr_gage_define_options(void)
{
  CmdLine& c = g_cmd;
  c.add("fao", "", "Read additional options from this file");
  c.add("fco", "", "Read calculation options from this file");
  c.add("fop", "", "Read output options from this file");
  c.add("cwd", "", "Change working directory");
  c.add("help", false, "Print out this help text");
  c.add("helpAll", false, "Print out this help text\nincluding the private options, if any", false);
  c.add("version", false, "Print out program version");
  c.add("debug", 0, "Debug print switch");
  c.add("terminal", 6, "File unit number for terminal output");
  c.add("consolemsg", false, "Output error messages to console");
  c.add("Bramsize", -1, "In-core size (MB) of displacement recovery matrix\n< 0: Use the same as in the reducer (default)\n= 0: Store full matrix in core");
  c.add("dmramsize", -1, "Same as -Bramsize but in terms of double words", false);
  c.add("linkId", 0, "Link base-ID number");
  c.add("linkfile", "", "Name of link input file");
  c.add("Bmatfile", "", "Name of B-matrix file");
  c.add("eigfile", "", "Name of eigenvector file");
  c.add("dispfile", "", "Name of gravitation displacement file");
  c.add("resfile", "", "Name of results output file");
  c.add("rdbfile", "", "Name of strain gage results database file");
  c.add("rdbinc", 1, "Increment number for the results database file");
  c.add("samfile", "", "Name of SAM data file");
  c.add("fsifile", "fedem_solver.fsi", "Name of solver input file");
  c.add("frsfile", "", "Name of solver results database file");
  c.add("rosfile", "", "Name of strain rosette input file");
  c.add("writeAsciiFiles", false, "Write rosette results to ASCII files");
  c.add("deformation", false, "Save nodal deformations to results database");
  c.add("nullify_start_rosettestrains", false, "Set start strains to zero for the rosettes");
  c.add("statm", 0.0, "Start time");
  c.add("stotm", 1.0, "Stop time");
  c.add("tinc", 0.0, "Time increment (= 0.0: process all time steps)");
  c.add("dac_sampleinc", 0.001, "Sampling increment for dac output files");
  c.add("flushinc", -1.0, "Time between each database file flush\n< 0.0: Do not flush results database (let the OS decide)\n= 0.0: Flush at each time step, no external buffers\n> 0.0: Flush at specified time interval, use external buffers");
  c.add("fatigue", 0, "Perform damage calculation on the gage stresses");
  c.add("stressToMPaScale", 1.0e-6, "Scale factor scaling stresses to MPa");
  c.add("gate", 25.0, "Stress gate value for the damage calculation [MPa]");
  c.add("binSize", 10.0, "Bin size for stress cycle counting [MPa]");
  c.add("loga1", 15.117, "Parameter log(a1) of the S-N curve");
  c.add("loga2", 17.146, "Parameter log(a2) of the S-N curve");
  c.add("m1", 4.0, "Parameter m1 of the S-N curve");
  c.add("littleEndian", false, "Use Little Endian formatting of DAC files");
  // B200 additions
  c.add("device", 0, "CUDA device ordinal");
  c.add("stepTile", 0, "Time steps per device batch (0 = from free device memory)");
  g_gage_options_defined = true;
}

// The option table of fedem_modes: the standard options + modesmain.C:22-52.
void fsr_modes_define_options(void)
{
  CmdLine& c = g_cmd;
  c.add("fao", "", "Read additional options from this file");
  c.add("fco", "", "Read calculation options from this file");
  c.add("fop", "", "Read output options from this file");
  c.add("cwd", "", "Change working directory");
  c.add("help", false, "Print out this help text");
  c.add("helpAll", false, "Print out this help text\nincluding the private options, if any", false);
  c.add("version", false, "Print out program version");
  c.add("debug", 0, "Debug print switch");
  c.add("terminal", 6, "File unit number for terminal output");
  c.add("consolemsg", false, "Output error messages to console");
  c.add("Bramsize", -1, "In-core size (MB) of displacement recovery matrix\n< 0: Use the same as in the reducer (default)\n= 0: Store full matrix in core");
  c.add("dmramsize", -1, "Same as -Bramsize but in terms of double words", false);
  c.add("linkId", 0, "Link base-ID number");
  c.add("linkfile", "", "Name of link input file");
  c.add("Bmatfile", "", "Name of B-matrix file");
  c.add("eigfile", "", "Name of eigenvector file");
  c.add("dispfile", "", "Name of gravitation displacement file");
  c.add("resfile", "", "Name of results output file");
  c.add("samfile", "", "Name of SAM data file");
  c.add("fsifile", "fedem_solver.fsi", "Name of solver input file");
  c.add("frsfile", "", "Name of solver results database file");
  c.add("rdbfile", "", "Name of modes results database file");
  c.add("rdbinc", 1, "Increment number for the results database file");
  c.add("VTFfile", "", "Name of VTF output file");
  c.add("VTFoffset", 0, "VTF result block id offset");
  c.add("VTFparts", 0, "Number of parts in VTF-file");
  c.add("VTFexpress", false, "Write express VTF-files (one file per mode)");
  c.add("VTFdscale", 1.0, "Deformation scaling factor for VTF output");
  c.add("double", false, "Save all results in double precision");
  c.add("damped", false, "Complex modes are calculated");
  c.add("recover_modes", "", "List of mode numbers to expand");
  c.add("write_nodes", false, "Save results as nodal data");
  c.add("write_vector", true, "Save results as vector data");
  c.add("energy_density", false, "Save scaled strain energy density");
  c.add("stressForm", 0, "General stress formulation option", false);
  c.add("ffqStressForm", 2, "Stress formulation for the FFQ shell", false);
  c.add("fftStressForm", 1, "Stress formulation for the FFT shell", false);
  c.add("useIncompatibleModes", false, "Linear hexahedron option", false);
  c.add("device", 0, "CUDA device ordinal");
  c.add("stepTile", 0, "Time steps per device batch (0 = from free device memory)");
  g_modes_options_defined = true;
}

// The option table of fedem_fpp: the standard options + fppmain.C:21-68.
void fsr_fpp_define_options(void)
{
  CmdLine& c = g_cmd;
  c.add("fao", "", "Read additional options from this file");
  c.add("fco", "", "Read calculation options from this file");
  c.add("fop", "", "Read output options from this file");
  c.add("cwd", "", "Change working directory");
  c.add("help", false, "Print out this help text");
  c.add("helpAll", false, "Print out this help text\nincluding the private options, if any", false);
  c.add("version", false, "Print out program version");
  c.add("debug", 0, "Debug print switch");
  c.add("terminal", 6, "File unit number for terminal output");
  c.add("consolemsg", false, "Output error messages to console");
  c.add("Bramsize", -1, "In-core size (MB) of displacement recovery matrix\n< 0: Use the same as in the reducer (default)\n= 0: Store full matrix in core");
  c.add("dmramsize", -1, "Same as -Bramsize but in terms of double words", false);
  c.add("linkId", 0, "Link base-ID number");
  c.add("linkfile", "", "Name of link input file");
  c.add("Bmatfile", "", "Name of B-matrix file");
  c.add("eigfile", "", "Name of eigenvector file");
  c.add("dispfile", "", "Name of gravitation displacement file");
  c.add("resfile", "", "Name of result output file");
  c.add("samfile", "", "Name of SAM data file");
  c.add("fsifile", "fedem_solver.fsi", "Name of solver input file");
  c.add("resStressFile", "", "Name of residual stress input file");
  c.add("resStressSet", "", "Name of residual stress set");
  c.add("frsfile", "", "Name of solver results database file");
  c.add("fppfile", "", "Name of fpp output file");
  c.add("rdbfile", "", "Name of strain coat results database file");
  c.add("rdbinc", 1, "Increment number for the results database file");
  c.add("double", false, "Save results in double precision");
  c.add("writeHistory", false, "Write history frs-files instead", false);
  c.add("oldRange", false, "Use old stress/strain range meassures", false);
  c.add("group", "", "List of element groups to do calculations for");
  c.add("blockSize", 2000, "Max number of elements processed together");
  c.add("BufSizeInc", 20, "Buffer increment size");
  c.add("PVXGate", 10.0, "Gate value for the Peak Valley extraction\n(MPa or microns depending on HistDataType)");
  c.add("biAxialGate", 10.0, "Gate value for the biaxiality calculation");
  c.add("angleBins", 541, "Number of bins in search for most popular angle");
  c.add("HistXMin", -100.0, "Histogram min X-value");
  c.add("HistXMax", 100.0, "Histogram max X-value");
  c.add("HistYMin", -100.0, "Histogram min Y-value");
  c.add("HistYMax", 100.0, "Histogram max Y-value");
  c.add("HistXBins", 64, "Histogram number of X-bins");
  c.add("HistYBins", 64, "Histogram number of Y-bins");
  c.add("HistDataType", 0, "Histogram data type\n= 0: None\n= 1: Signed abs max stress\n= 2: Signed abs max strain");
  c.add("surface", 0, "Surface selection option\n= 0: All element surfaces\n= 1: Bottom shell surfaces only\n= 2: Middle shell surfaces only\n= 3: Top shell surfaces only");
  c.add("SNfile", "", "Name of SN-curve definition file");
  c.add("stressToMPaScale", 1.0e-6, "Stress convertion factor to MPa");
  c.add("statm", 0.0, "Start time");
  c.add("stotm", 1.0, "Stop time");
  c.add("tinc", 0.0, "Time increment (= 0.0: process all time steps)");
  // B200 additions
  c.add("device", 0, "CUDA device ordinal");
  c.add("stepTile", 0, "Time steps per device batch (0 = from free device memory)");
  g_fpp_options_defined = true;
}

void initSolverArgs(int argc, char** argv)
{
  g_cmd = CmdLine();
  g_cmd.init(argc, argv);
  g_gage_options_defined = false;
  g_modes_options_defined = false;
  g_fpp_options_defined = false;
  fsr_stress_define_options();
}

#define FAIL(...) do { set_error(__VA_ARGS__); log.line(" *** Error: %s", fsr_last_error()); log.line("\n    %s failed :-(", what); return -1; } while (0)
#define CHECK(call) do { const int rc_ = (call); if (rc_ < 0) { log.line(" *** Error: %s", fsr_last_error()); log.line("\n    %s failed :-(", what); return rc_; } } while (0)

}  // extern "C"

// subroutine stress (stress.f90:17-493) and subroutine gage (gage.f90:8-431) share everything up to the opened B and E
// matrices and the time step selection: one body, `gage` switches the program specific parts.
static int gage_part(CmdLine& c, Log& log, const char* what, fsr_part* part, fsr_ftl* ftl, fsr_frs* db, int isup, int sup_user_id, const char* model_file,
                     const std::string& linkfile, const std::vector<int>& minex, const std::vector<double>& xyz, int ndof2, int ngen,
                     int ntriads, const std::vector<int>& tb, const std::vector<int>& tnd, const std::vector<int>& tfd,
                     const std::vector<double>& tru, int gen_first, const std::vector<int>& sel, int nsel, const std::vector<int>& stepno,
                     const std::vector<double>& times, bool lgrav, const double* grv, const std::vector<int>& madof);

static int fpp_part(CmdLine& c, Log& log, const char* what, fsr_part* part, fsr_ftl* ftl, fsr_frs* db, int isup, int user_id, const char* descr,
                    const char* model_file, const std::string& linkfile, const std::vector<double>& xyz, int ndof2, int ngen, int ntriads,
                    const std::vector<int>& tb, const std::vector<int>& tnd, const std::vector<int>& tfd, const std::vector<double>& tru,
                    int gen_first, const std::vector<int>& sel, int nsel, const std::vector<int>& stepno, const std::vector<double>& times,
                    bool lgrav, const double* grv, const std::vector<int>& madof);
static int modes_part(CmdLine& c, Log& log, const char* what, fsr_part* part, fsr_frs* db, int isup, int user_id, const char* descr,
                      const char* model_file, const std::string& linkfile, const std::vector<int>& madof, const std::vector<int>& minex,
                      int ndof2, int ngen, int ntriads, const std::vector<int>& tb, const std::vector<int>& tnd, const std::vector<int>& tfd,
                      const std::vector<double>& tru, int gen_first, const std::vector<int>& stepno, const std::vector<double>& times, bool lgrav,
                      const double* grv, const std::vector<int>& melcon, const std::vector<int>& elmid);

static int run_program(int which)
{
  const bool gage = which == 1, modes = which == 2, fpp = which == 3;
  const char* what = gage ? "Strain gage recovery" : modes ? "Modal recovery" : fpp ? "Strain coat calculation" : "Stress calculation";
  const char* prog = gage ? "fedem_gage" : modes ? "fedem_modes" : fpp ? "fedem_fpp" : "fedem_stress";
  if (gage) {
    if (!g_gage_options_defined) { g_cmd.clear_options(); g_stress_options_defined = g_modes_options_defined = g_fpp_options_defined = false; fsr_gage_define_options(); }
  } else if (modes) {
    if (!g_modes_options_defined) { g_cmd.clear_options(); g_stress_options_defined = g_gage_options_defined = g_fpp_options_defined = false; fsr_modes_define_options(); }
  } else if (fpp) {
    if (!g_fpp_options_defined) { g_cmd.clear_options(); g_stress_options_defined = g_gage_options_defined = g_modes_options_defined = false; fsr_fpp_define_options(); }
  } else if (!g_stress_options_defined)
    fsr_stress_define_options();
  CmdLine& c = g_cmd;
  // readOptionFilesStd (cmdLineArgInitStd.C:103-135)
  const std::string cwd = c.get_string("cwd");
  if (!cwd.empty() && chdir(cwd.c_str())) { perror((std::string(prog) + ": " + cwd).c_str()); return 1; }
  for (const char* o : {"fao", "fco", "fop"}) { const std::string f = c.get_string(o); if (!f.empty()) c.read_options_file(f); }
  if (c.get_bool("help") || c.get_bool("helpAll")) { printf("%s", c.help_text(c.get_bool("helpAll")).c_str()); return 0; }
  if (c.get_bool("version")) { printf("%s B200 1.0\n", prog); return 0; }

  // CUDA context creation (a few hundred ms on a B200) runs beside the parsing of the input files
  using clk = std::chrono::steady_clock;
  const clk::time_point t_start = clk::now();
  auto since = [](clk::time_point t) { return std::chrono::duration<double>(clk::now() - t).count(); };
  double t_context = 0.0;
  const int device = c.get_int("device");
  std::thread ctx_thread([&] { const clk::time_point t = clk::now(); if (cudaSetDevice(device) == cudaSuccess) cudaFree(nullptr); t_context = since(t); });
  struct CtxJoin { std::thread& t; ~CtxJoin() { if (t.joinable()) t.join(); } } ctx_join{ctx_thread};

  Log log;
  log.open(file_name("resfile", gage ? "_gage.res" : modes ? "_modes.res" : fpp ? "_fpp.res" : "_stress.res"),
           gage ? "Strain Gage Recovery" : modes ? "Modal Recovery" : fpp ? "Damage Recovery" : "Stress Recovery");
  log.line("\n           ================> START OF PROGRAM %s <================", gage ? "GAGE" : modes ? "MODES" : fpp ? "FPP" : "STRESS");
  if (gage) {
    if (c.get_bool("writeAsciiFiles")) log.line("  ** Note: ASCII / DAC rosette result files (-writeAsciiFiles) are not part of this build; ignored");
  } else
  if (!modes && c.is_set("resStressFile")) log.line("  ** Note: residual stress import (-resStressFile) is not part of this build; ignored");
  if (!fpp && c.is_set("VTFfile") && !c.get_string("VTFfile").empty()) log.line("  ** Note: VTF export (-VTFfile) is not part of this build; ignored");
  if (!fpp && c.get_bool("dumpDefNas")) log.line("  ** Note: Nastran deformation dump (-dumpDefNas) is not part of this build; ignored");
  if (which == 0 && c.get_bool("nodalForces")) log.line("  ** Note: nodal force print-out (-nodalForces, stressRoutines.f90:128) is not part of this build; ignored");

  // --- Read the link file (ffl_init)
  const std::string linkfile = c.get_string("linkfile");
  if (linkfile.empty()) FAIL("FE data file must be specified through -linkfile");
  log.line("           --> Reading link files");
  fsr_ftl* ftl = nullptr;
  CHECK(fsr_ftl_open(&ftl, linkfile.c_str()));
  struct FtlGuard { fsr_ftl* p; ~FtlGuard() { fsr_ftl_close(p); } } ftl_guard{ftl};
  const std::string groups = modes ? std::string() : c.get_string("group");
  if (!groups.empty()) CHECK(fsr_ftl_activate_groups(ftl, groups.c_str()));
  int fsz[12];
  const int nael = fsr_ftl_sizes(ftl, fsz);

  // --- Establish the SAM datastructure (initiateSAM, samStressModule.f90:39-263)
  const std::string samfile = file_name("samfile", ".fsm");
  int mpar[64] = {0}, cs_sam = 0;
  CHECK(fsr_fsm_read_mpar(samfile.c_str(), &cs_sam, mpar, 64));
  const int nnod = mpar[0], nel = mpar[1], ndof = mpar[2], ndof1 = mpar[3], ndof2 = mpar[4], nceq = mpar[6], neq = mpar[10],
            nmmnpc = mpar[14], nmmceq = mpar[15], ngen = mpar[21];
  if (fsz[0] != nnod || fsz[1] != nel || fsz[2] != ndof || fsz[3] < nmmnpc || fsz[5] != mpar[22])
    FAIL("The FE data file %s does not match the SAM file %s (nnod %d/%d, nel %d/%d, ndof %d/%d, nmmnpc %d/%d, nxnod %d/%d)",
         linkfile.c_str(), samfile.c_str(), fsz[0], nnod, fsz[1], nel, fsz[2], ndof, fsz[3], nmmnpc, fsz[5], mpar[22]);
  std::vector<int> madof(nnod + 1), minex(std::max(nnod, 1)), mnnn(std::max(nnod, 1)), msc(std::max(ndof, 1)), mpmnpc(nel + 1),
      mmnpc(std::max(nmmnpc, 1)), melcon(std::max(nel, 1)), mpmceq(nceq + 1), mmceq(std::max(nmmceq, 1)), meqn(std::max(ndof, 1)),
      meqn1(std::max(ndof1, 1)), meqn2(std::max(ndof2, 1));
  std::vector<double> ttcc(std::max(nmmceq, 1));
  CHECK(fsr_fsm_read(samfile.c_str(), madof.data(), minex.data(), mnnn.data(), msc.data(), mpmnpc.data(), mmnpc.data(), melcon.data(),
                     mpmceq.data(), mmceq.data(), ttcc.data(), meqn.data(), meqn1.data(), meqn2.data()));
  log.line("           --> FE part: %d nodes, %d elements (%d active), %d DOFs (%d internal, %d external, %d component modes)", nnod, nel,
           nael, ndof, ndof1, ndof2, ngen);

  // --- element data in SAM order
  std::vector<double> xyz(3 * (size_t)std::max(nnod, 1)), emod(std::max(nel, 1)), rny(std::max(nel, 1)), rho(std::max(nel, 1)),
      thk(std::max(nel, 1)), beam((size_t)FSR_NBEAM * std::max(nel, 1));
  std::vector<int> elmid(std::max(nel, 1)), estat(std::max(nel, 1));
  CHECK(fsr_ftl_get_nodes(ftl, nullptr, nullptr, nullptr, nullptr, xyz.data()));
  const int nbad = fsr_ftl_get_elmdata(ftl, emod.data(), rny.data(), rho.data(), thk.data(), elmid.data(), beam.data(), estat.data());
  if (nbad > 0) log.line("  ** Warning: %d elements lack material / thickness / cross section data", nbad);
  if (which == 0) {   // legacy thin shells are recovered for the default stress formulations only (fsr_part_create): say so loudly otherwise
    const int ffq = c.get_int("ffqStressForm"), fft = c.get_int("fftStressForm");
    int nq = 0, nt = 0;
    bool has23 = false;
    for (int e = 0; e < nel; ++e) has23 = has23 || melcon[(size_t)e] == 23;
    for (int e = 0; e < nel; ++e) {
      if (elmid[(size_t)e] < 1) continue;
      if (melcon[(size_t)e] == 22 && ffq != 2 && ffq != 1) ++nq;
      if (melcon[(size_t)e] == 21 && fft != 1 && has23) ++nt;   // FTS31 / FTS32 is served for parts whose triangles are all FFT3
    }
    if (nq) log.line("  ** Warning: %d FFQ shells (type 22) get NO results: -ffqStressForm %d is not supported by this build (only 1 and the default, 2)", nq, ffq);
    if (nt) log.line("  ** Warning: %d FFT shells (type 21) get NO results: -fftStressForm %d next to ANDES triangles (type 23) in one part is not supported by this build", nt, fft);
  }

  // --- Read superelement data from the solver input file (readSolverData)
  int isup = c.get_int("linkId");
  if (isup < 1) isup = mpar[17];
  log.line("           --> Process Part; baseID (isup) = %d", isup);
  int user_id = 0, ntriads = 0, ngen_fsi = 0, gen_first = 1;
  char descr[256] = "", model_file[1024] = "";
  double sup_pos[12] = {1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0}, grv[3] = {0, 0, 0};
  std::vector<int> tb, tu, tnd, tfd;
  std::vector<double> tru;
  bool have_fsi = c.is_set("fsifile");
  if (have_fsi) {
    fsr_fsi* fsi = nullptr;
    CHECK(fsr_fsi_open(&fsi, c.get_string("fsifile").c_str(), isup));
    fsr_fsi_part(fsi, &user_id, descr, sizeof(descr), &ntriads, &ngen_fsi, sup_pos, grv, model_file, sizeof(model_file));
    tb.resize(std::max(ntriads, 1)); tu.resize(tb.size()); tnd.resize(tb.size()); tfd.resize(tb.size()); tru.resize(12 * tb.size());
    gen_first = fsr_fsi_triads(fsi, tb.data(), tu.data(), tnd.data(), tfd.data(), tru.data(), nullptr);
    fsr_fsi_close(fsi);
    if (gen_first - 1 != ndof2 || ngen_fsi != ngen)
      FAIL("Part %d: the solver input file gives %d triad DOFs + %d component modes, the SAM file %d + %d", isup, gen_first - 1, ngen_fsi, ndof2, ngen);
  } else if (which != 0)
    FAIL("%s needs the solver input file (-fsifile)", prog);
  else {
    // stress.f90:131-135: no supernodes, a direct solution is assumed to be on the results files
    // ("Vectors|Dynamic response|Displacement" of the part, readDisplPointer / readIntDisplacements)
    log.line("           --> No solver input file: the nodal displacements are read from the results database");
    snprintf(descr, sizeof(descr), "%s", linkfile.c_str());
  }
  const bool direct = !have_fsi;

  // --- gravitation displacement modes (stress.f90:184-206): folded in as three extra "component modes"
  const bool want_grav = std::sqrt(grv[0] * grv[0] + grv[1] * grv[1] + grv[2] * grv[2]) > 1.0e-8 && c.is_set("dispfile");
  std::string dispfile = want_grav ? file_name("dispfile", "_V.fmx") : std::string();
  bool lgrav = false;
  if (want_grav) { FILE* t = fopen(dispfile.c_str(), "rb"); if (t) { lgrav = true; fclose(t); } }
  // gage.f90:175-194 adds the strains of the static gravitation deflection vgii = dis1Expand(V . g), g the gravitation vector of the
  // solver input file (NOT turned with the part, unlike stress.f90:412): same three extra modes, constant amplitudes
  const int nmodes = ngen + (lgrav ? 3 : 0);

  // --- the part on the device
  const double t_parse = since(t_start);
  if (ctx_thread.joinable()) ctx_thread.join();
  const clk::time_point t_setup0 = clk::now();
  fsr_sam sam;
  memset(&sam, 0, sizeof(sam));
  sam.nnod = nnod; sam.nel = nel; sam.ndof = ndof; sam.ndof1 = ndof1; sam.ndof2 = ndof2; sam.ngen = nmodes; sam.neq = neq; sam.nceq = nceq;
  sam.nmmnpc = nmmnpc; sam.nmmceq = nmmceq;
  sam.madof = madof.data(); sam.msc = msc.data(); sam.mpmnpc = mpmnpc.data(); sam.mmnpc = mmnpc.data(); sam.melcon = melcon.data();
  sam.mpmceq = mpmceq.data(); sam.mmceq = mmceq.data(); sam.ttcc = ttcc.data(); sam.meqn = meqn.data(); sam.meqn1 = meqn1.data();
  sam.meqn2 = meqn2.data();
  fsr_elmdata ed;
  ed.xyz = xyz.data(); ed.emod = emod.data(); ed.rny = rny.data(); ed.thk = thk.data(); ed.elmid = elmid.data(); ed.beam = beam.data();
  fsr_options po;
  memset(&po, 0, sizeof(po));
  po.device = c.get_int("device"); po.stressForm = c.get_int("stressForm"); po.step_tile = c.get_int("stepTile");
  if (!gage && !fpp) { po.reserved[1] = c.get_int("ffqStressForm") + 1; po.reserved[2] = c.get_int("fftStressForm") + 1; }
  if (fpp) po.stressForm = 0;
  // fedem_stress spreads the elements of the part over several GPUs (element blocks, sharded.cu) when the part is large;
  // nodal deformation output needs all nodes on one device
  int ngpu = 1;
  if (which == 0) {
    int nvis = 1;
    cudaGetDeviceCount(&nvis);
    ngpu = c.get_int("gpus");
    // automatic: the results database is paced by the file (one GPU fills ~80 GB/s of records), so only very large parts gain;
    // an envelope-only run scales with the GPUs
    bool any_out = false;
    for (const char* o : {"vmStress", "maxPStress", "minPStress", "maxSStress", "vmStrain", "maxPStrain", "minPStrain", "maxSStrain", "stress", "strain", "SR"})
      any_out = any_out || c.get_bool(o);
    if (ngpu <= 0) ngpu = std::max(1, std::min(nvis, nael / (any_out ? 2000000 : 250000)));
    if (ngpu > nvis) { log.line("  ** Note: -gpus %d but only %d GPU(s) visible", ngpu, nvis); ngpu = nvis; }
    if (ngpu > 1 && c.get_bool("deformation")) { log.line("  ** Note: -deformation is written by one GPU: -gpus %d ignored", ngpu); ngpu = 1; }
    if (ngpu > 1 && direct) { log.line("  ** Note: a direct solution is recovered by one GPU: -gpus %d ignored", ngpu); ngpu = 1; }
  }
  fsr_part* part = nullptr;
  fsr_group* group = nullptr;
  int nfail;
  if (ngpu > 1) {
    std::vector<int> devs((size_t)ngpu);
    for (int i = 0; i < ngpu; ++i) devs[(size_t)i] = (po.device + i) % ngpu;
    nfail = fsr_group_create(&group, &sam, &ed, &po, devs.data(), ngpu);
    if (nfail >= 0) log.line("           --> %d element blocks on %d GPUs", fsr_group_num_blocks(group), ngpu);
  } else
    nfail = fsr_part_create(&part, &sam, &ed, &po);
  if (nfail < 0) { log.line(" *** Error: %s", fsr_last_error()); log.line("\n    Stress calculation failed :-("); return nfail; }
  struct PartGuard { fsr_part*& p; fsr_group*& g; ~PartGuard() { if (p) fsr_part_destroy(p); if (g) fsr_group_destroy(g); } } part_guard{part, group};
  if (nfail > 0) log.line("  ** Warning: the stress operator of %d elements could not be formed; they get %g", nfail, kHuge);

  // --- Open the B-matrix and the generalized modes files (openBandEmatrices); not needed for a direct solution
  if (!direct && !c.is_set("Bmatfile") && ndof2 > 0) log.line("  ** Note: -Bmatfile not given, using %s", file_name("Bmatfile", "_B.fmx").c_str());
  if (!direct) {
    std::vector<double> B((size_t)std::max(ndof1, 1) * std::max(ndof2, 1)), E((size_t)std::max(ndof1, 1) * std::max(nmodes, 1));
    char tag[64];
    int cs = 0, sp = 0;
    if (ndof1 > 0 && ndof2 > 0) {
      CHECK(fsr_fmx_read(file_name("Bmatfile", "_B.fmx").c_str(), tag, sizeof(tag), &cs, &sp, B.data(), (long long)ndof1 * ndof2));
      if (cs != cs_sam) log.line("  ** Warning: checksum of the B-matrix file (%d) differs from the SAM file (%d)", cs, cs_sam);
    }
    if (ndof1 > 0 && ngen > 0) {
      CHECK(fsr_fmx_read(file_name("eigfile", "_E.fmx").c_str(), tag, sizeof(tag), &cs, &sp, E.data(), (long long)ndof1 * ngen));
      if (strcmp(tag, "#FEDEM generalized modes") != 0) FAIL("Invalid disk matrix file %s: file tag '%s'", file_name("eigfile", "_E.fmx").c_str(), tag);
    }
    if (lgrav) {   // readDoubleDB(chname,'displacement matrix',ndof1*3,vii) (stress.f90:199)
      CHECK(fsr_fmx_read(dispfile.c_str(), tag, sizeof(tag), &cs, &sp, E.data() + (size_t)ndof1 * ngen, (long long)ndof1 * 3));
      if (strcmp(tag, "#FEDEM displacement matrix") != 0) FAIL("%s is not a displacement matrix file, tag=%s", dispfile.c_str(), tag);
      log.line("           --> Gravitation displacement modes read from %s", dispfile.c_str());
    }
    if (group) CHECK(fsr_group_set_recovery(group, ndof2 > 0 ? B.data() : nullptr, ndof1, nmodes > 0 ? E.data() : nullptr, ndof1));
    else CHECK(fsr_set_recovery(part, ndof2 > 0 ? B.data() : nullptr, ndof1, nmodes > 0 ? E.data() : nullptr, ndof1));
  }

  const double t_setup = since(t_setup0);
  // --- Open the solver results database (ffr_init) and select the time steps (ffr_getnextstep loop)
  log.line("           --> Reading solver result files");
  const std::vector<std::string> frs_files = file_list(c.get_string("frsfile"));
  if (frs_files.empty()) FAIL("No results database files specified");
  std::vector<const char*> fp;
  for (const std::string& s : frs_files) fp.push_back(s.c_str());
  fsr_frs* db = nullptr;
  CHECK(fsr_frs_open(&db, fp.data(), (int)fp.size()));
  struct FrsGuard { fsr_frs* p; ~FrsGuard() { fsr_frs_close(p); } } frs_guard{db};
  const int nall = fsr_frs_num_steps(db);
  std::vector<int> stepno(std::max(nall, 1));
  std::vector<double> times(std::max(nall, 1));
  fsr_frs_get_steps(db, stepno.data(), times.data(), nall);
  const double statm = c.get_double("statm"), stotm = c.get_double("stotm"), tinc = c.get_double("tinc");
  std::vector<int> sel(std::max(nall, 1));
  const int nsel = fsr_select_steps(times.data(), nall, statm, stotm, tinc, sel.data(), nall);
  if (!modes) log.line("           --> %d of %d time steps selected in [%g, %g], increment %g", nsel, nall, statm, stotm, tinc);
  if (modes)
    return modes_part(c, log, what, part, db, isup, user_id, descr[0] ? descr : linkfile.c_str(), model_file, linkfile, madof, minex, ndof2, ngen, ntriads,
                      tb, tnd, tfd, tru, gen_first, stepno, times, lgrav, grv,
                      std::vector<int>(melcon.begin(), melcon.begin() + nel), std::vector<int>(elmid.begin(), elmid.begin() + nel));
  if (fpp)
    return fpp_part(c, log, what, part, ftl, db, isup, user_id, descr[0] ? descr : linkfile.c_str(), model_file, linkfile, xyz, ndof2, ngen, ntriads, tb, tnd,
                    tfd, tru, gen_first, sel, nsel, stepno, times, lgrav, grv, madof);
  if (gage)
    return gage_part(c, log, what, part, ftl, db, isup, user_id, model_file, linkfile, minex, xyz, ndof2, ngen, ntriads, tb, tnd, tfd, tru, gen_first,
                     sel, nsel, stepno, times, lgrav, grv, madof);

  // --- Initialize the stress results database (writeStressHeader)
  fsr_rdb_options ro;
  memset(&ro, 0, sizeof(ro));
  static const struct { const char* opt; unsigned bit; } kOut[] = {
      {"vmStress", FSR_OUT_VMSTRESS}, {"maxPStress", FSR_OUT_MAXPSTRESS}, {"minPStress", FSR_OUT_MINPSTRESS},
      {"maxSStress", FSR_OUT_MAXSSTRESS}, {"vmStrain", FSR_OUT_VMSTRAIN}, {"maxPStrain", FSR_OUT_MAXPSTRAIN},
      {"minPStrain", FSR_OUT_MINPSTRAIN}, {"maxSStrain", FSR_OUT_MAXSSTRAIN}, {"stress", FSR_OUT_STRESS},
      {"strain", FSR_OUT_STRAIN}, {"SR", FSR_OUT_SR}, {"deformation", FSR_OUT_DEFORMATION}};
  for (const auto& k : kOut) if (c.get_bool(k.opt)) ro.out_mask |= k.bit;
  if ((ro.out_mask & FSR_OUT_DEFORMATION) && (!c.get_bool("write_nodes") || c.get_bool("write_vector")))
    log.line("  ** Note: deformations are written as nodal data (-write_vector is not part of this build)");
  fsr_rdb* rdb = nullptr;
  const std::string descr_s = descr[0] ? descr : linkfile;
  if (ro.out_mask) {
    log.line("           --> Writing result database headers");
    ro.double_precision = c.get_bool("double") ? 1 : 0;
    ro.rdbinc = c.get_int("rdbinc");
    ro.part_base_id = isup; ro.part_user_id = user_id; ro.part_descr = descr_s.c_str();
    ro.model_file = model_file; ro.link_file = linkfile.c_str();
    ro.elmid = elmid.data(); ro.minex = minex.data(); ro.sup_tr_init = sup_pos;
    if (group) CHECK(fsr_rdb_create_group(&rdb, group, file_name("rdbfile", ".frs").c_str(), &ro));
    else CHECK(fsr_rdb_create(&rdb, part, file_name("rdbfile", ".frs").c_str(), &ro));
    char path[1024];
    fsr_rdb_path(rdb, path, sizeof(path));
    log.line("           --> Results database file: %s (%lld bytes per time step)", path, fsr_rdb_step_bytes(rdb));
  } else
    log.line("  ** Note: no result output requested, only the von Mises envelope is computed");
  struct RdbGuard { fsr_rdb*& p; ~RdbGuard() { if (p) fsr_rdb_close(p); } } rdb_guard{rdb};

  // --- Time loop, in windows of consecutive selected steps
  log.line("           --> Starting time loop");
  const int ndim = ndof2 + nmodes, window = 256;
  std::vector<double> Q((size_t)ndim * window), supTr(12 * (size_t)window), Qs((size_t)std::max(ndof2 + ngen, 1)), t_w(window);
  std::vector<int> s_w(window);
  std::vector<double> sup_all;
  const int hsup = fsr_frs_find(db, "Position matrix", "Part", isup);
  for (int k = 0; k < window; ++k)   // part position when the results files hold none: where the modelling put it
    for (int j = 0; j < 12; ++j) supTr[12 * (size_t)k + j] = sup_pos[j];
  int hdis = -1;
  std::vector<double> SV;
  if (direct) {
    hdis = fsr_frs_find(db, "Vectors|Dynamic response|Displacement", "Part", isup);
    if (hdis < 0) FAIL("No solver input file (-fsifile) and no nodal displacements of Part %d (Vectors|Dynamic response|Displacement) on the results files", isup);
    if (fsr_frs_var_size(db, hdis) != ndof)
      FAIL("Mismatch between length of wanted array: %d and actual variable size: %d (nodal displacements of Part %d)", ndof, fsr_frs_var_size(db, hdis), isup);
    SV.resize((size_t)ndof * window);
  }
  const clk::time_point t_loop0 = clk::now();
  double t_hist = 0.0;
  bool warned_no_response = false;
  for (int w0 = 0; w0 < nsel; w0 += window) {
    const int nw = std::min(window, nsel - w0);
    const clk::time_point t_h0 = clk::now();
    for (int k = 0; k < nw;) {   // runs of consecutive steps on file are read with one call
      int run = 1;
      while (k + run < nw && sel[w0 + k + run] == sel[w0 + k] + run) ++run;
      if (direct)   // readIntDisplacements (displacementModule.f90:865-904)
        CHECK(fsr_frs_read(db, hdis, sel[w0 + k], run, SV.data() + (size_t)k * ndof, ndof, ndof));
      else {
        double* q0 = Q.data() + (size_t)k * ndim;
        const int rch = fsr_frs_reduced_history(db, isup, ntriads, tb.data(), tnd.data(), tfd.data(), tru.data(), ngen, gen_first, sel[w0 + k], run, q0, ndim);
        CHECK(rch);
        if (rch > 0 && !warned_no_response) { log.line("  ** Warning: %s", fsr_last_error()); warned_no_response = true; }
      }
      if (hsup >= 0) CHECK(fsr_frs_read(db, hsup, sel[w0 + k], run, supTr.data() + 12 * (size_t)k, 12, 12));
      k += run;
    }
    for (int k = 0; k < nw; ++k) {
      s_w[k] = stepno[sel[w0 + k]];
      t_w[k] = times[sel[w0 + k]];
      if (lgrav && !direct) {   // g = matmul(grv, sup%supTr(:,1:3)) (stress.f90:412)
        const double* T = supTr.data() + 12 * (size_t)k;
        double* g = Q.data() + (size_t)k * ndim + ndof2 + ngen;
        for (int j = 0; j < 3; ++j) g[j] = grv[0] * T[3 * j] + grv[1] * T[3 * j + 1] + grv[2] * T[3 * j + 2];
      }
    }
    t_hist += since(t_h0);
    // queued: the device, the PCIe copy and the file writer work on this window while the next one is read
    if (direct) {
      if (rdb) CHECK(fsr_rdb_write_steps_displacements(rdb, SV.data(), nw, s_w.data(), t_w.data(), supTr.data()));
      else CHECK(fsr_recover_displacements(part, SV.data(), nw, nullptr));
    } else if (rdb) CHECK(fsr_rdb_write_steps(rdb, Q.data(), ndim, nw, s_w.data(), t_w.data(), supTr.data()));
    else if (group) CHECK(fsr_group_recover(group, Q.data(), ndim, nw, nullptr));
    else CHECK(fsr_recover(part, Q.data(), ndim, nw, nullptr));
    log.line("           --> ......Simulation time : %12.5E  (%d of %d steps done)", t_w[nw - 1], w0 + nw, nsel);
  }
  if (!rdb && nsel > 0) {
    std::vector<double> mx(std::max(group ? fsr_group_num_result_points(group) : fsr_num_result_points(part), 1)), mn(mx.size());
    if (group) CHECK(fsr_group_get_envelope(group, mx.data(), mn.data()));
    else CHECK(fsr_get_envelope(part, mx.data(), mn.data()));
    log.line("           --> largest von Mises stress over all result points and steps: %g", *std::max_element(mx.begin(), mx.end()));
  }
  double tp[5] = {0, 0, 0, 0, 0};
  if (rdb) CHECK(fsr_rdb_flush(rdb, tp, 5));
  const double t_loop = since(t_loop0);
  log.line("           --> Time loop done. Closing database files");
  if (rdb) { fsr_rdb* r = rdb; rdb = nullptr; CHECK(fsr_rdb_close(r)); }
  // where the wall time went; the three pipeline stages overlap with each other and with the reading of the history
  log.line("           --> Wall time %.3f s: input files %.3f (CUDA context %.3f beside it), part on device + B/E %.3f, time loop %.3f",
           since(t_start), t_parse, t_context, t_setup, t_loop);
  log.line("           --> Time loop: history read %.3f s | device (K1 + record kernels) %.3f s | device-to-host %.3f s | file %.3f s (%.1f MB in %d tiles)",
           t_hist, 1e-3 * tp[0], 1e-3 * tp[1], 1e-3 * tp[2], 1e-6 * tp[3], (int)tp[4]);
  log.line("           ================>  END OF PROGRAM STRESS  <================");
  log.line("\n    Stress calculation successfully completed :-)  (%.2f s CPU)", (double)(clock() - log.t0) / CLOCKS_PER_SEC);
  return 0;
}

// getShellElementAxes (src/vpmStress/strainAndStressUtils.f90:339-434) for 3 or 4 nodes X[k][3]; axes V1, V2, V3.  Returns 0 if ok.
static int shell_element_axes(int n, const double (*X)[3], double* V1, double* V2, double* V3, bool globalize = false)
{
  auto cross = [](const double* a, const double* b, double* c) { c[0] = a[1] * b[2] - a[2] * b[1]; c[1] = a[2] * b[0] - a[0] * b[2]; c[2] = a[0] * b[1] - a[1] * b[0]; };
  auto dot = [](const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; };
  const double eps2 = kEpsDiv0 * kEpsDiv0;
  if (n == 3) {
    for (int k = 0; k < 3; ++k) { V1[k] = X[1][k] - X[0][k]; V2[k] = X[2][k] - X[0][k]; }
  } else if (n == 4) {
    for (int k = 0; k < 3; ++k) { V1[k] = X[2][k] - X[0][k]; V2[k] = X[3][k] - X[1][k]; }
  } else
    return -1;
  cross(V1, V2, V3);
  double vn = dot(V3, V3);
  if (vn <= eps2) return 2;
  vn = std::sqrt(vn);
  for (int k = 0; k < 3; ++k) V3[k] /= vn;
  if (globalize) {   // getGlobalizedX (:297-336): the global X axis (or Y when the normal is close to it) projected into the element plane
    if (std::fabs(V3[1]) > 0.01 || std::fabs(V3[2]) > 0.01) {
      V1[0] = V3[1] * V3[1] + V3[2] * V3[2]; V1[1] = -V3[0] * V3[1]; V1[2] = -V3[0] * V3[2];
    } else {
      V2[0] = -V3[1] * V3[0]; V2[1] = V3[0] * V3[0] + V3[2] * V3[2]; V2[2] = -V3[1] * V3[2];
      cross(V2, V3, V1);
    }
    const double l2 = dot(V1, V1);
    for (int k = 0; k < 3; ++k) V1[k] = l2 > eps2 ? V1[k] / std::sqrt(l2) : 0.0;
  } else if (n == 4) {
    for (int k = 0; k < 3; ++k) V1[k] = X[1][k] - X[0][k];
    cross(V3, V1, V2);
    cross(V2, V3, V1);
  }
  vn = dot(V1, V1);
  if (vn <= eps2) return 3;
  vn = std::sqrt(vn);
  for (int k = 0; k < 3; ++k) V1[k] /= vn;
  cross(V3, V1, V2);
  return 0;
}

// FFa_glbEulerZYX -> FaMat33::getEulerZYX (fedem-foundation/src/FFaLib/FFaAlgebra/FFaMat33.C:327-349), a = 3x3 column-major
static void glb_euler_zyx(const double* a, double* ang)
{
  auto atan3 = [](double y, double x) { return std::fabs(y) > 1.0e-15 || std::fabs(x) > 1.0e-15 ? std::atan2(y, x) : 0.0; };   // FFaMath.H:44-47
  const double a11 = a[0], a21 = a[1], a31 = a[2], a32 = a[5], a33 = a[8];
  ang[2] = atan3(a21, a11);
  ang[1] = -atan3(a31, std::hypot(a11, a21));
  ang[0] = atan3(a32, a33);
}

// --------------------------------------------------------------------------------------------------------------------
// The fedem_gage specific part (gage.f90:136-150,196-247,278-400): rosette input, Bcart on the GPU, results database
// (saveStrainGageModule.f90:23-268), fatigue report (reportDamage, strainGageModule.f90:778-862).
static int gage_part(CmdLine& c, Log& log, const char* what, fsr_part* part, fsr_ftl* ftl, fsr_frs* db, int isup, int sup_user_id, const char* model_file,
                     const std::string& linkfile, const std::vector<int>& minex, const std::vector<double>& xyz, int ndof2, int ngen,
                     int ntriads, const std::vector<int>& tb, const std::vector<int>& tnd, const std::vector<int>& tfd,
                     const std::vector<double>& tru, int gen_first, const std::vector<int>& sel, int nsel, const std::vector<int>& stepno,
                     const std::vector<double>& times, bool lgrav, const double* grv, const std::vector<int>& madof)
{
  (void)minex;
  // --- Initializing strain rosettes (readStrainGageData, checkRosette)
  log.line("           --> Initializing strain rosettes");
  const std::string rosfile = c.get_string("rosfile");
  if (rosfile.empty()) FAIL("No strain rosette input file (-rosfile)");
  const bool fsi_format = rosfile.size() >= 4 && rosfile.compare(rosfile.size() - 4, 4, ".fsi") == 0;   // gage.f90:138
  constexpr int kDescr = 128;
  std::vector<fsr_rosette> ros;
  std::vector<int> ruser;
  std::vector<char> rdescr;
  int nros = 0;
  if (fsi_format) {
    nros = fsr_fsi_read_rosettes(rosfile.c_str(), isup, nullptr, nullptr, nullptr, 0, 0);
    if (nros < 0) CHECK(nros);
    log.line("               Number of &STRAIN_ROSETTE =%6d", nros);
  } else {
    // The old rosette definition file (ReadStrainGageOldData, strainGageModule.f90:246-476): free-format records
    //   id type link nnod node_1..node_nnod zPos Xx Xy Xz Zx Zy Zz Emod nu      ('#' starts a comment, END / EOF ends the file)
    // of which the ones of this part (link == the part's user id) are kept; type 1 single gage, 2 double (90 deg),
    // 3 triple 60 deg, 4 triple 45 deg; a dummy base id idIn + link*1000 + rdbinc*10000000 names the rosette on the frs file.
    FILE* fp = fopen(rosfile.c_str(), "r");
    if (!fp) FAIL("Could not open the rosette definition file %s", rosfile.c_str());
    std::vector<std::string> tok;
    {
      char lbuf[4096];
      bool done = false;
      while (!done && fgets(lbuf, sizeof(lbuf), fp)) {
        if (char* hash = strchr(lbuf, '#')) *hash = 0;
        for (char* t = strtok(lbuf, " \t\r\n,"); t; t = strtok(nullptr, " \t\r\n,")) {
          if (!strncasecmp(t, "END", 3) || !strncasecmp(t, "EOF", 3)) { done = true; break; }
          tok.push_back(t);
        }
      }
      fclose(fp);
    }
    const int rdbinc = c.get_int("rdbinc");
    size_t it = 0;
    bool bad_file = false;
    auto next_i = [&](int& v) { if (it >= tok.size()) { bad_file = true; v = 0; return; } char* e = nullptr; v = (int)strtol(tok[it].c_str(), &e, 10); if (*e) bad_file = true; ++it; };
    auto next_d = [&](double& v) {
      if (it >= tok.size()) { bad_file = true; v = 0.0; return; }
      std::string t = tok[it++];
      for (char& ch : t) if (ch == 'D' || ch == 'd') ch = 'e';   // Fortran exponents
      char* e = nullptr; v = strtod(t.c_str(), &e); if (*e) bad_file = true;
    };
    while (it < tok.size() && !bad_file) {
      int id, type, link, nn;
      next_i(id); next_i(type); next_i(link); next_i(nn);
      if (bad_file || nn < 3 || nn > 4) { bad_file = true; break; }
      fsr_rosette R;
      memset(&R, 0, sizeof(R));
      R.numnod = nn;
      for (int k = 0; k < nn; ++k) next_i(R.nodes[k]);
      double xv[3], zv[3];
      next_d(R.zpos);
      for (int k = 0; k < 3; ++k) next_d(xv[k]);
      for (int k = 0; k < 3; ++k) next_d(zv[k]);
      next_d(R.emod); next_d(R.nu);
      if (bad_file) break;
      if (link != sup_user_id) continue;   // skip for all other links
      const double lx = std::sqrt(xv[0] * xv[0] + xv[1] * xv[1] + xv[2] * xv[2]), lz = std::sqrt(zv[0] * zv[0] + zv[1] * zv[1] + zv[2] * zv[2]);
      if (lx < 1000.0 * 2.2250738585072014e-308) FAIL("Undefined x-direction for rosette number:%8d", id);
      if (lz < 1000.0 * 2.2250738585072014e-308) FAIL("Undefined z-direction for rosette number:%8d", id);
      for (int k = 0; k < 3; ++k) { R.rpos[k] = xv[k] / lx; R.rpos[6 + k] = zv[k] / lz; }   // posInGl(:,1), posInGl(:,3); (:,2) and (:,4) follow below
      const double pi = 3.14159265358979323846;
      switch (type) {
        case 1: R.ngage = 1; R.alpha_gages = 0.0; break;
        case 2: R.ngage = 2; R.alpha_gages = pi / 2.0; break;
        case 3: R.ngage = 3; R.alpha_gages = pi / 3.0; break;
        case 4: R.ngage = 3; R.alpha_gages = pi / 4.0; break;
        default: FAIL("Undefined rosette-type:%8d", type);
      }
      ++nros;
      R.id = nros + link * 1000 + rdbinc * 10000000;
      ros.push_back(R);
      ruser.push_back(id);
    }
    if (bad_file) FAIL("Failed to read rosette definition file %s", rosfile.c_str());
    rdescr.assign((size_t)std::max(nros, 1) * kDescr, 0);
    log.line("               Number of strain rosettes on this link =%6d", nros);
  }
  if (nros == 0) {
    log.line("  ** Note: No strain rosettes on this link");
    log.line("\n    %s successfully completed :-)", what);
    return 0;
  }
  if (fsi_format) {
    ros.resize((size_t)nros);
    ruser.resize((size_t)nros);
    rdescr.resize((size_t)nros * kDescr);
    CHECK(fsr_fsi_read_rosettes(rosfile.c_str(), isup, ros.data(), ruser.data(), rdescr.data(), kDescr, nros));
  }
  for (int r = 0; r < nros; ++r) {
    fsr_rosette& R = ros[(size_t)r];
    int bad = 0;
    for (int k = 0; k < R.numnod; ++k) {
      R.nodes[k] = fsr_ftl_ext2int(ftl, 1, R.nodes[k]);
      if (R.nodes[k] < 1) ++bad;
    }
    if (bad) FAIL("%d invalid node number(s) for Rosette %d", bad, R.id);
    // checkOrientation (strainGageModule.f90:524-570): the element normal of the first three nodes must follow the rosette Z axis
    const double* X[3];
    for (int k = 0; k < 3; ++k) X[k] = &xyz[3 * (size_t)(R.nodes[k] - 1)];
    const double a[3] = {X[1][0] - X[0][0], X[1][1] - X[0][1], X[1][2] - X[0][2]}, b[3] = {X[2][0] - X[0][0], X[2][1] - X[0][1], X[2][2] - X[0][2]};
    const double vn[3] = {a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]};
    if (R.rpos[6] * vn[0] + R.rpos[7] * vn[1] + R.rpos[8] * vn[2] < 0.0) {
      for (int k = 0; k < R.numnod / 2; ++k) std::swap(R.nodes[k], R.nodes[R.numnod - 1 - k]);
      log.line("  ** Note: Nodal ordering for Rosette %d has been swapped", R.id);
    }
  }
  if (!fsi_format)
    // calcElmCoordSystem with useElCoordSys = .false. (strainRosetteModule.f90:506-583, gage.f90:216-229): the rosette sits in
    // the centroid of its nodes, takes the element's Z axis and the given X direction projected into the element plane
    for (int r = 0; r < nros; ++r) {
      fsr_rosette& R = ros[(size_t)r];
      double X[4][3], T[9];
      for (int k = 0; k < R.numnod; ++k) for (int d = 0; d < 3; ++d) X[k][d] = xyz[3 * (size_t)(R.nodes[k] - 1) + d];
      if (shell_element_axes(R.numnod, X, T, T + 3, T + 6)) FAIL("Could not calculate coordinate system for Rosette %d. Check the rosette definition.", ruser[(size_t)r]);
      for (int d = 0; d < 3; ++d) { double sum = 0.0; for (int k = 0; k < R.numnod; ++k) sum += X[k][d]; R.rpos[9 + d] = sum / R.numnod; }
      double* P = R.rpos;   // columns X, Y, Z, origin
      if (T[6] * P[6] + T[7] * P[7] + T[8] * P[8] < kEpsDiv0) FAIL("Could not calculate coordinate system for Rosette %d. Check the rosette definition.", ruser[(size_t)r]);
      for (int d = 0; d < 3; ++d) P[6 + d] = T[6 + d];
      P[3] = P[7] * P[2] - P[8] * P[1]; P[4] = P[8] * P[0] - P[6] * P[2]; P[5] = P[6] * P[1] - P[7] * P[0];   // Y = Z x X
      P[0] = P[4] * P[8] - P[5] * P[7]; P[1] = P[5] * P[6] - P[3] * P[8]; P[2] = P[3] * P[7] - P[4] * P[6];   // X = Y x Z
      // orthoNorm3 = mat_to_quat + quat_to_mat (rotationModule.f90:441-497,549-558); rten(i,j) = P[3 (j-1) + (i-1)]
      auto M = [&](int i, int j) -> double { return P[3 * (j - 1) + (i - 1)]; };
      double q[5] = {0, 0, 0, 0, 0};   // q[1..4]
      const double trace = M(1, 1) + M(2, 2) + M(3, 3);
      int imax = 1;
      if (M(2, 2) > M(imax, imax)) imax = 2;
      if (M(3, 3) > M(imax, imax)) imax = 3;
      if (trace > M(imax, imax)) {
        q[1] = std::sqrt(1.0 + trace) * 0.5;
        q[2] = (M(3, 2) - M(2, 3)) / (4.0 * q[1]);
        q[3] = (M(1, 3) - M(3, 1)) / (4.0 * q[1]);
        q[4] = (M(2, 1) - M(1, 2)) / (4.0 * q[1]);
      } else {
        const int i = imax, j = imax % 3 + 1, k = (imax + 1) % 3 + 1;
        q[i + 1] = std::sqrt(M(i, i) * 0.5 + (1.0 - trace) * 0.25);
        q[1] = (M(k, j) - M(j, k)) / (4.0 * q[i + 1]);
        q[j + 1] = (M(j, i) + M(i, j)) / (4.0 * q[i + 1]);
        q[k + 1] = (M(k, i) + M(i, k)) / (4.0 * q[i + 1]);
      }
      const double qn = std::sqrt(q[1] * q[1] + q[2] * q[2] + q[3] * q[3] + q[4] * q[4]);
      for (int k = 1; k <= 4; ++k) q[k] /= qn;
      double Rm[3][3];
      Rm[0][0] = 2.0 * (q[2] * q[2] + q[1] * q[1]) - 1.0;
      Rm[1][1] = 2.0 * (q[3] * q[3] + q[1] * q[1]) - 1.0;
      Rm[2][2] = 2.0 * (q[4] * q[4] + q[1] * q[1]) - 1.0;
      Rm[0][1] = 2.0 * (q[2] * q[3] - q[4] * q[1]);
      Rm[0][2] = 2.0 * (q[2] * q[4] + q[3] * q[1]);
      Rm[1][2] = 2.0 * (q[3] * q[4] - q[2] * q[1]);
      Rm[1][0] = 2.0 * (q[3] * q[2] + q[4] * q[1]);
      Rm[2][0] = 2.0 * (q[4] * q[2] - q[3] * q[1]);
      Rm[2][1] = 2.0 * (q[4] * q[3] + q[2] * q[1]);
      for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) P[3 * j + i] = Rm[i][j];
    }
  if (c.get_bool("nullify_start_rosettestrains"))   // gage.f90:312-320: every rosette starts from zero strain at the first step
    for (fsr_rosette& R : ros) R.zero_init = 1;
  fsr_gages* gages = nullptr;
  CHECK(fsr_gage_create(&gages, part, ros.data(), nros));
  struct GageGuard { fsr_gages* p; ~GageGuard() { fsr_gage_destroy(p); } } gage_guard{gages};

  // --- Writing result database headers (writeStrainGageHeader / writeRosetteHeader)
  log.line("           --> Writing result database headers");
  std::string hdr = rdb_file_preamble("fedem_gage", model_file, linkfile.c_str());
  hdr += "VARIABLES:\n<1;\"Time step number\";NONE;INT;32;NUMBER>\n<2;\"Physical time\";TIME;FLOAT;64;SCALAR>\n"
         "<3;\"Angle of maximum principal strain/stress\";ANGLE;FLOAT;32;SCALAR>\n<4;\"Angle of maximum shear\";ANGLE;FLOAT;32;SCALAR>\n"
         "<5;\"Strain tensor\";NONE;FLOAT;32;TENSOR2;(3);((\"epsilon_x\",\"epsilon_y\",\"epsilon_xy\"))>\n"
         "<6;\"Stress tensor\";FORCE/AREA;FLOAT;32;TENSOR2;(3);((\"sigma_x\",\"sigma_y\",\"sigma_xy\"))>\n"
         "<7;\"Gage strain\";NONE;FLOAT;32;SCALAR>\n<8;\"Gage stress\";NONE;FLOAT;32;SCALAR>\n";
  // .fsi format: InitStrainRosette is called with calcDisp = .true. (gage.f90:203-214), so every rosette carries its node
  // deformations (written with -deformation) and its position + Euler angles in the global system (always written)
  // (the old rosette definition format goes without calcDisp: none of the three)
  const bool lDisp = fsi_format;
  const bool lDef = lDisp && c.get_bool("deformation");
  int nv = 8;
  const int idDef = lDef ? ++nv : 0, idPos = lDisp ? ++nv : 0, idAng = lDisp ? ++nv : 0;
  if (lDef) { char b[128]; snprintf(b, sizeof(b), "<%d;\"Deformation\";LENGTH;FLOAT;32;VEC3;(3);((\"d_x\",\"d_y\",\"d_z\"))>\n", idDef); hdr += b; }
  if (lDisp) { char b[256]; snprintf(b, sizeof(b), "<%d;\"Position\";LENGTH;FLOAT;32;VEC3;(3);((\"x\",\"y\",\"z\"))>\n<%d;\"Euler angles\";ANGLE;FLOAT;32;ROT3;(3);((\"eps_x\",\"eps_y\",\"eps_z\"))>\n", idPos, idAng); hdr += b; }
  int maxg = 0;
  for (const fsr_rosette& R : ros) maxg = std::max(maxg, R.ngage);
  for (int j = 1; j <= maxg; ++j) { char b[64]; snprintf(b, sizeof(b), "[%d;\"Gage %d\";<7><8>]\n", j, j); hdr += b; }
  int nig = maxg;
  std::vector<std::string> node_groups((size_t)nros);
  if (lDef)   // writeItGDef(rdb,idNode,iFile,'Node'//StrId(globalNodes(j))) with idNode = 0: a new item group per rosette node
    for (int r = 0; r < nros; ++r)
      for (int k = 0; k < ros[(size_t)r].numnod; ++k) {
        char b[96];
        snprintf(b, sizeof(b), "[%d;\"Node%d\";<%d>]\n", ++nig, ros[(size_t)r].nodes[k], idDef);
        hdr += b;
        snprintf(b, sizeof(b), "[%d]", nig);
        node_groups[(size_t)r] += b;
      }
  hdr += "DATABLOCKS:\n<1><2>\n";
  long long nval = 0;
  for (int r = 0; r < nros; ++r) {
    const fsr_rosette& R = ros[(size_t)r];
    char b[512];
    std::string line = "{\"Strain rosette\";";
    if (R.id > 0) { snprintf(b, sizeof(b), "%d;", R.id); line += b; } else line += ";";
    if (ruser[(size_t)r] > 0) { snprintf(b, sizeof(b), "%d;", ruser[(size_t)r]); line += b; } else line += ";";
    const char* d = &rdescr[(size_t)r * kDescr];
    if (*d) { snprintf(b, sizeof(b), "\"%s\";", d); line += b; } else line += ";";
    line += "<3><4><5><6>";
    for (int j = 1; j <= R.ngage; ++j) { snprintf(b, sizeof(b), "[%d]", j); line += b; }
    line += node_groups[(size_t)r];
    if (lDisp) { snprintf(b, sizeof(b), "<%d><%d>", idPos, idAng); line += b; }
    hdr += line + "}\n";
    nval += 8 + 2 * R.ngage + (lDef ? 3 * R.numnod : 0) + (lDisp ? 6 : 0);
  }
  // X0 and T0 of every rosette (InitStrainRosette, strainRosetteModule.f90:753-764): node coordinates and initial element axes
  std::vector<double> X0((size_t)nros * 12), T0((size_t)nros * 9);
  for (int r = 0; r < nros; ++r) {
    const fsr_rosette& R = ros[(size_t)r];
    double X[4][3];
    for (int k = 0; k < R.numnod; ++k) for (int d = 0; d < 3; ++d) X[k][d] = X0[(size_t)r * 12 + 3 * k + d] = xyz[3 * (size_t)(R.nodes[k] - 1) + d];
    double* T = &T0[(size_t)r * 9];      // columns = V1, V2, V3 (T0 = transpose(T_el), T_el rows = the element axes)
    if (shell_element_axes(R.numnod, X, T, T + 3, T + 6)) FAIL("Could not calculate coordinate system for Rosette %d. Check the rosette definition.", R.id);
  }
  std::string path = file_name("rdbfile", ".frs");
  {   // openRDBfile (rdbModule.f90:268-403): <name>_<rdbinc>.<ext>
    const int inc = c.get_int("rdbinc");
    if (inc > 0) {
      const size_t dot = path.rfind('.'), sep = path.rfind('/');
      char b[16];
      snprintf(b, sizeof(b), "_%d", inc);
      if (dot != std::string::npos && dot > 0 && (sep == std::string::npos || dot > sep)) path.insert(dot, b);
      else path += b;
    }
  }
  fsr_frs_writer* w = nullptr;
  CHECK(fsr_frs_create(&w, path.c_str(), 0, hdr.c_str(), 4 * nval));
  struct WGuard { fsr_frs_writer*& p; ~WGuard() { if (p) fsr_frs_finish(p); } } w_guard{w};
  log.line("           --> Results database file: %s (%lld bytes per time step)", path.c_str(), 12 + 4 * nval);

  // --- Time loop in windows: reduced history (readSupElDisplacements + BuildFinit), rosette strains on the GPU
  log.line("           --> Starting time loop");
  const int ndof = madof.back() - 1;
  const int ndim = ndof2 + ngen + (lgrav ? 3 : 0);
  const int window = 256;
  (void)ndof;
  // the translational DOFs of the rosette nodes: the only rows of the expansion CalcRosetteDisplacements needs
  std::vector<int> urows;
  for (const fsr_rosette& R : ros)
    for (int i = 0; i < 4; ++i)
      for (int j = 0; j < 3; ++j) urows.push_back(i < R.numnod ? madof[(size_t)(R.nodes[i] - 1)] - 1 + j : 0);
  const int nur = (int)urows.size();   // 12 per rosette
  const int iFatigue = c.get_int("fatigue");
  std::vector<double> Uw((size_t)window * nur), supTr(12 * (size_t)window);
  for (int k = 0; k < window; ++k) { double* T = &supTr[12 * (size_t)k]; for (int j = 0; j < 12; ++j) T[j] = (j == 0 || j == 4 || j == 8) ? 1.0 : 0.0; }
  const int hsup = fsr_frs_find(db, "Position matrix", "Part", isup);
  std::vector<double> Qall(iFatigue > 0 ? (size_t)ndim * std::max(nsel, 1) : 0), Q((size_t)ndim * window),
      vals((size_t)window * nros * FSR_GAGE_NVAL);
  std::vector<float> recbuf((size_t)std::max<long long>(nval, 1));
  for (int w0 = 0; w0 < nsel; w0 += window) {
    const int nw = std::min(window, nsel - w0);
    for (int k = 0; k < nw;) {
      int run = 1;
      while (k + run < nw && sel[(size_t)(w0 + k + run)] == sel[(size_t)(w0 + k)] + run) ++run;
      {
        const int rch = fsr_frs_reduced_history(db, isup, ntriads, tb.data(), tnd.data(), tfd.data(), tru.data(), ngen, gen_first, sel[(size_t)(w0 + k)], run,
                                    Q.data() + (size_t)k * ndim, ndim);
        CHECK(rch);
        if (rch > 0 && w0 == 0 && k == 0) log.line("  ** Warning: %s", fsr_last_error());
      }
      if (hsup >= 0) CHECK(fsr_frs_read(db, hsup, sel[(size_t)(w0 + k)], run, supTr.data() + 12 * (size_t)k, 12, 12));
      k += run;
    }
    if (lgrav)
      for (int k = 0; k < nw; ++k) for (int j = 0; j < 3; ++j) Q[(size_t)k * ndim + ndof2 + ngen + j] = grv[j];
    if (iFatigue > 0) memcpy(Qall.data() + (size_t)w0 * ndim, Q.data(), sizeof(double) * (size_t)nw * ndim);
    CHECK(fsr_gage_recover(gages, Q.data(), ndim, nw, vals.data()));
    if (lDisp) CHECK(fsr_expand_rows(part, Q.data(), ndim, nw, urows.data(), nur, Uw.data()));   // node displacements for CalcRosetteDisplacements
    for (int k = 0; k < nw; ++k) {   // writeStrainGageDB (saveStrainGageModule.f90:196-262)
      size_t n = 0;
      for (int r = 0; r < nros; ++r) {
        const double* v = vals.data() + ((size_t)k * nros + r) * FSR_GAGE_NVAL;
        recbuf[n++] = (float)v[8]; recbuf[n++] = (float)v[9];
        recbuf[n++] = (float)v[0]; recbuf[n++] = (float)v[1]; recbuf[n++] = (float)(0.5 * v[2]);
        recbuf[n++] = (float)v[10]; recbuf[n++] = (float)v[11]; recbuf[n++] = (float)v[12];
        for (int j = 0; j < ros[(size_t)r].ngage; ++j) { recbuf[n++] = (float)v[18 + j]; recbuf[n++] = (float)v[21 + j]; }
        // CalcRosetteDisplacements (strainRosetteModule.f90:814-873): node deformations, position and Euler angles of the rosette
        if (!lDisp) continue;   // old rosette definition format: no rosette displacement state
        const fsr_rosette& R = ros[(size_t)r];
        const double* u = Uw.data() + (size_t)k * nur + 12 * (size_t)r;
        const double* S = &supTr[12 * (size_t)k];
        double Xn[4][3], posR[3] = {0, 0, 0};
        for (int i = 0; i < R.numnod; ++i) {
          const double* d = u + 3 * i;
          for (int j = 0; j < 3; ++j) {
            if (lDef) recbuf[n++] = (float)d[j];
            posR[j] += d[j];
            Xn[i][j] = X0[(size_t)r * 12 + 3 * i + j] + d[j];
          }
        }
        for (int j = 0; j < 3; ++j) posR[j] = R.rpos[9 + j] + posR[j] / (double)R.numnod;
        for (int j = 0; j < 3; ++j) recbuf[n++] = (float)(S[j] * posR[0] + S[3 + j] * posR[1] + S[6 + j] * posR[2] + S[9 + j]);   // matmul34
        double Tn[9], Ti[9], Tg[9], ang[3];
        if (shell_element_axes(R.numnod, Xn, Tn, Tn + 3, Tn + 6)) FAIL("Could not calculate the deformed coordinate system of Rosette %d", R.id);
        const double* T0r = &T0[(size_t)r * 9];
        for (int a = 0; a < 3; ++a) for (int b2 = 0; b2 < 3; ++b2) {   // Tinc = Tn . T0^T
          double t = 0.0;
          for (int q = 0; q < 3; ++q) t += Tn[a + 3 * q] * T0r[b2 + 3 * q];
          Ti[a + 3 * b2] = t;
        }
        for (int a = 0; a < 3; ++a) for (int b2 = 0; b2 < 3; ++b2) {   // Tinc . supTr(:,1:3)
          double t = 0.0;
          for (int q = 0; q < 3; ++q) t += Ti[a + 3 * q] * S[q + 3 * b2];
          Tg[a + 3 * b2] = t;
        }
        glb_euler_zyx(Tg, ang);
        for (int j = 0; j < 3; ++j) recbuf[n++] = (float)ang[j];
      }
      CHECK(fsr_frs_write_step(w, stepno[(size_t)sel[(size_t)(w0 + k)]], times[(size_t)sel[(size_t)(w0 + k)]], recbuf.data()));
    }
    log.line("           --> ......Simulation time : %12.5E  (%d of %d steps done)", times[(size_t)sel[(size_t)(w0 + nw - 1)]], w0 + nw, nsel);
  }

  // --- fatigue (AddFatiguePoints during the loop + reportDamage at the end)
  if (iFatigue > 0 && nsel > 0) {
    log.line("           --> Time loop done. Performing fatigue calculation");
    const double to_mpa = c.get_double("stressToMPaScale"), gate = c.get_double("gate"), bin_size = c.get_double("binSize");
    // m2: the reference has no command-line default for it -- reportDamage hands snCurve(4) of the rosette record to
    // ffp_getdamage as it is, 0 when the record does not set it (strainGageModule.f90:157,811-816); same here, and the
    // report below prints the value the computation used
    const double curve[4] = {c.get_double("loga1"), c.get_double("loga2"), c.get_double("m1"), 0.0};
    const int nser = 4 * nros;
    std::vector<double> damage((size_t)nser);
    std::vector<int> ncyc((size_t)nser), status((size_t)nser), bins;
    int nbins = 64;
    for (;;) {   // reportDamage walks the bins until every series answers -1 (no cycles left above the bin)
      bins.assign((size_t)nser * nbins, 0);
      CHECK(fsr_gage_fatigue(gages, Qall.data(), ndim, nsel, to_mpa, gate, curve, bin_size, nbins, damage.data(), ncyc.data(), bins.data(),
                             status.data()));
      bool open_end = false;
      for (int s = 0; s < nser; ++s) if (bins[(size_t)s * nbins + nbins - 1] >= 0) open_end = true;
      if (!open_end || nbins >= 65536) break;
      nbins *= 4;
    }
    // status 1 = the reference's own closure failure (ffp_getdamage ignores it too, the cycles counted so far stand);
    // status 2 = the turning-point stack of the series overflowed: its damage and cycle counts are NOT valid
    int nclosure = 0, noverflow = 0;
    for (int r = 0; r < nros; ++r)
      for (int j = 0; j <= ros[(size_t)r].ngage; ++j) {
        const int st = status[(size_t)(4 * r + j)];
        if (st == 2) {
          ++noverflow;
          log.line(" *** Error: Rosette %d, %s: the rainflow residue exceeded the stack capacity; no damage / cycle counts for this series",
                   ros[(size_t)r].id, j == 0 ? "max principal stress" : (std::string("gage ") + std::to_string(j)).c_str());
        } else if (st == 1) {
          ++nclosure;
          log.line("  ** Warning: Rosette %d, %s: the rainflow residue did not close on three points (as in the reference, the cycles "
                   "counted so far are kept)", ros[(size_t)r].id, j == 0 ? "max principal stress" : (std::string("gage ") + std::to_string(j)).c_str());
        }
      }
    if (noverflow) log.line(" *** Error: %d series without fatigue results (marked -1 below)", noverflow);
    for (int r = 0; r < nros; ++r) {
      const fsr_rosette& R = ros[(size_t)r];
      const double g = R.gate > 0.0 ? R.gate : gate;
      const double sn[4] = {R.sncurve[0] > 0.0 ? R.sncurve[0] : curve[0], R.sncurve[1] > 0.0 ? R.sncurve[1] : curve[1],
                            R.sncurve[2] > 0.0 ? R.sncurve[2] : curve[2], R.sncurve[3]};
      log.line("\n     ===== Computed damage in strain rosette =====\n           gate value :%12.5E\n           S-N curve  : log(a1) =%7.3f log(a2) =%7.3f m1 =%7.3f m2 =%7.3f"
               "\n\n     Strain Rosette                      Max princ.  Gage 1      Gage 2  ...", g, sn[0], sn[1], sn[2], sn[3]);
      char idtxt[64], row[256];
      snprintf(idtxt, sizeof(idtxt), " [%d] %s", ruser[(size_t)r], &rdescr[(size_t)r * kDescr]);
      int n = snprintf(row, sizeof(row), "    %-36.36s", idtxt);
      for (int j = 0; j <= R.ngage; ++j) n += snprintf(row + n, sizeof(row) - (size_t)n, "%12.5E", damage[(size_t)(4 * r + j)]);
      log.line("%s\n", row);
      for (int b = 0; b < nbins; ++b) {
        bool any_pos = false, any_open = false;
        n = 0;
        for (int j = 0; j <= R.ngage; ++j) {
          const int cnt = bins[(size_t)(4 * r + j) * nbins + b];
          if (cnt >= 0) { any_open = true; n += snprintf(row + n, sizeof(row) - (size_t)n, "%12d", cnt); if (cnt > 0) any_pos = true; }
          else n += snprintf(row + n, sizeof(row) - (size_t)n, "%12s", "");
        }
        if (!any_open) break;
        if (any_pos) log.line("     Stress cycles %7.2f -%7.2f%s", b * bin_size, (b + 1) * bin_size, row);
      }
      log.line("     =============================================\n");
    }
    log.line("           --> Closing database files");
  } else
    log.line("           --> Time loop done. Closing database files");
  { fsr_frs_writer* x = w; w = nullptr; CHECK(fsr_frs_finish(x)); }
  log.line("           ================>  END OF PROGRAM GAGE  <================");
  log.line("\n    %s successfully completed :-)  (%.2f s CPU)", what, (double)(clock() - log.t0) / CLOCKS_PER_SEC);
  return 0;
}

// --------------------------------------------------------------------------------------------------------------------
// The fedem_modes specific part (modes.f90:236-460, stressInterface.C:30-83): the dynamic response and the requested eigenmodes at
// the requested times, expanded with ONE K1 call per time (columns = response, mode 1 Re, mode 1 Im, mode 2 ...), written as vector
// data to one modal results file (writeModesHeader :356-426, writeNodesHeader :541-595, writeDisplacementDB :1437-1515).
static std::vector<std::string> bracket_tokens(const std::string& s)   // FFaTokenizer(s,'<','>',','), one nesting level kept
{
  std::vector<std::string> out;
  size_t i = 0;
  while (i < s.size() && isspace((unsigned char)s[i])) ++i;
  if (i >= s.size()) return out;
  if (s[i] != '<') { out.push_back(s.substr(i)); return out; }
  int depth = 0;
  std::string tok;
  for (; i < s.size(); ++i) {
    const char ch = s[i];
    if (ch == '<') { if (depth++ > 0) tok += ch; }
    else if (ch == '>') { if (--depth > 0) tok += ch; else { if (!tok.empty()) out.push_back(tok); break; } }
    else if (ch == ',' && depth == 1) { if (!tok.empty()) out.push_back(tok); tok.clear(); }
    else if (!isspace((unsigned char)ch) || depth > 1) tok += ch;
  }
  return out;
}

static int modes_part(CmdLine& c, Log& log, const char* what, fsr_part* part, fsr_frs* db, int isup, int user_id, const char* descr,
                      const char* model_file, const std::string& linkfile, const std::vector<int>& madof, const std::vector<int>& minex,
                      int ndof2, int ngen, int ntriads, const std::vector<int>& tb, const std::vector<int>& tnd, const std::vector<int>& tfd,
                      const std::vector<double>& tru, int gen_first, const std::vector<int>& stepno, const std::vector<double>& times, bool lgrav,
                      const double* grv, const std::vector<int>& melcon, const std::vector<int>& elmid)
{
  // --- -recover_modes <<t1,m1,m2,..>,<t2,..>> (getModesToExpand, stressInterface.C:30-83): the union of the mode numbers in order
  // of first appearance, processThisMode[step][mode]
  const std::string opt = c.get_string("recover_modes");
  const std::vector<std::string> tt = bracket_tokens(opt);
  if (tt.empty()) FAIL("fedem_modes: No time steps, check option -recover_modes \"%s\"", opt.c_str());
  const int nStep = (int)tt.size();
  std::vector<double> tsteps((size_t)nStep, 0.0);
  std::vector<int> modeNum;
  std::vector<std::vector<char>> process;   // [mode][step]
  for (int i = 0; i < nStep; ++i) {
    const std::vector<std::string> mt = bracket_tokens(tt[(size_t)i]);
    if (mt.size() < 2) FAIL("fedem_modes: No modes at time step %d check sub-token \"%s\" in -recover_modes", i, tt[(size_t)i].c_str());
    tsteps[(size_t)i] = atof(mt[0].c_str());
    for (size_t j = 1; j < mt.size(); ++j) {
      const int jMode = atoi(mt[j].c_str());
      size_t k = 0;
      while (k < modeNum.size() && modeNum[k] != jMode) ++k;
      if (k == modeNum.size()) { modeNum.push_back(jMode); process.emplace_back((size_t)nStep, (char)0); }
      process[k][(size_t)i] = 1;
    }
  }
  const int nMode = (int)modeNum.size();
  if (!c.get_string("VTFfile").empty()) log.line("  ** Note: VTF export (-VTFfile) is not part of this build; ignored");
  const bool lComplex = c.get_bool("damped"), lDouble = c.get_bool("double"), lEnergy = c.get_bool("energy_density");
  const int ncomp = lComplex ? 2 : 1, nnod = (int)madof.size() - 1, ndof = madof.back() - 1;

  // --- initWriteDisp (saveStressModule.f90:79-108): vector form if asked for and possible, nodal form otherwise or on request
  bool lVector = c.get_bool("write_vector"), lNodes = true;
  for (int i = 0; i < nnod && lVector; ++i) if (madof[(size_t)i + 1] < madof[(size_t)i] + 3) lVector = false;
  if (lVector) lNodes = c.get_bool("write_nodes");
  bool all_processed = true;
  for (const std::vector<char>& pm : process) for (char f : pm) all_processed = all_processed && f;
  const bool multiFiles = lNodes || lEnergy || !all_processed;   // modes.f90:258-261
  int ntra = 0, nrot = 0;
  for (int i = 0; i < nnod; ++i) {
    if (minex[(size_t)i] < 0) continue;   // internal beam nodes
    ntra += 3;
    if (madof[(size_t)i + 1] >= madof[(size_t)i] + 6) nrot += 3;
  }

  // --- result pointers of the eigenvectors (readModesPointers, modesRoutines.f90:47-119)
  std::vector<int> hm((size_t)nMode * (ntriads + 1), -1);
  for (int j = 0; j < nMode; ++j) {
    char path[64];
    snprintf(path, sizeof(path), "Eigenvectors|Mode%3d", modeNum[(size_t)j]);
    for (int i = 0; i < ntriads; ++i)
      if ((hm[(size_t)j * (ntriads + 1) + i] = fsr_frs_find(db, path, "Triad", tb[(size_t)i])) < 0)
        FAIL("Can not find eigenvector components for Triad {%d} and Mode%3d", tb[(size_t)i], modeNum[(size_t)j]);
    if (ngen > 0 && (hm[(size_t)j * (ntriads + 1) + ntriads] = fsr_frs_find(db, path, "Part", isup)) < 0)
      FAIL("Can not find eigenvector components for Part {%d} (component modes) and Mode%3d", isup, modeNum[(size_t)j]);
  }
  const int hsup = fsr_frs_find(db, "Position matrix", "Part", isup);

  // --- Writing result database headers (writeModesHeader :356-426: all shapes as vectors in one file; writeModeHeader :259-343: one
  //     file for the dynamic response and one per mode, with the nodal form and / or the scaled strain energy density)
  log.line("           --> Writing result database headers");
  const int nbit = lDouble ? 64 : 32, nel = (int)melcon.size();
  std::vector<int> ptoff((size_t)nel + 1, 0);
  if (lEnergy) CHECK(fsr_result_point_offsets(part, ptoff.data()));
  // one header: comps = 1 (dynamic response) or ncomp; vectors = the names of the [;"Vectors"; entries (single-file form: all of them);
  // returns the number of values of a step record
  auto make_header = [&](const std::vector<std::string>& names, const std::vector<bool>& compl_, bool nodes, bool vectors_wrapped, bool energy,
                         std::string& hdr) -> long long {
    std::string vard = rdb_file_preamble("fedem_modes", model_file, linkfile.c_str(), "modes data base file");
    vard += "VARIABLES:\n<1;\"Time step number\";NONE;INT;32;NUMBER>\n<2;\"Physical time\";TIME;FLOAT;64;SCALAR>\n";
    std::string item, datd = "DATABLOCKS:\n<1><2>\n{\"Part\";";
    char b[512];
    if (isup > 0) { snprintf(b, sizeof(b), "%d;", isup); datd += b; } else datd += ";";
    if (user_id > 0) { snprintf(b, sizeof(b), "%d;", user_id); datd += b; } else datd += ";";
    if (descr && *descr) { snprintf(b, sizeof(b), "\"%s\";\n", descr); datd += b; } else datd += ";\n";
    int nvar = 2, nig = 0;
    long long nval = 0;
    const bool cx = compl_[0];
    if (nodes) {   // writeNodesHeader :459-537 (iDef = 1)
      int idDis = 0, idRot = 0, id3 = 0, id6 = 0;
      datd += "  [;\"Nodes\";\n";
      for (int i = 0; i < nnod; ++i) {
        const int nd = madof[(size_t)i + 1] - madof[(size_t)i];
        if (nd < 3) continue;
        if (!idDis) { idDis = ++nvar; snprintf(b, sizeof(b), "<%3d;\"Translational deformation\";LENGTH;FLOAT;%2d;VEC3;(3);((\"d_x\",\"d_y\",\"d_z\"))>\n", idDis, nbit); vard += b; }
        int ig;
        if (nd > 5) {
          if (!idRot) { idRot = ++nvar; snprintf(b, sizeof(b), "<%3d;\"Angular deformation\";ANGLE;FLOAT;%2d;ROT3;(3);((\"theta_x\",\"theta_y\",\"theta_z\"))>\n", idRot, nbit); vard += b; }
          if (!id6) {
            id6 = ++nig;
            if (cx) snprintf(b, sizeof(b), "[%3d;\"%s\";[;\"Re\";<%3d><%3d>][;\"Im\";<%3d><%3d>]]\n", id6, names[0].c_str(), idDis, idRot, idDis, idRot);
            else snprintf(b, sizeof(b), "[%3d;\"%s\";<%3d><%3d>]\n", id6, names[0].c_str(), idDis, idRot);
            item += b;
          }
          ig = id6; nval += cx ? 12 : 6;
        } else {
          if (!id3) {
            id3 = ++nig;
            if (cx) snprintf(b, sizeof(b), "[%3d;\"%s\";[;\"Re\";<%3d>][;\"Im\";<%3d>]]\n", id3, names[0].c_str(), idDis, idDis);
            else snprintf(b, sizeof(b), "[%3d;\"%s\";<%3d>]\n", id3, names[0].c_str(), idDis);
            item += b;
          }
          ig = id3; nval += cx ? 6 : 3;
        }
        snprintf(b, sizeof(b), "    [;%8d;[%3d]]\n", minex[(size_t)i], ig); datd += b;
      }
      datd += "  ]\n";
    }
    if (lVector) {   // :538-588
      int idDis = 0, idRot = 0;
      if (vectors_wrapped) datd += "  [;\"Vectors\";\n";
      for (size_t k = 0; k < names.size(); ++k) {
        if (!idDis) { idDis = ++nvar; snprintf(b, sizeof(b), "<%3d;\"Translational deformation\";LENGTH;FLOAT;%2d;VECTOR;(%8d)>\n", idDis, nbit, ntra); vard += b; }
        if (nrot > 0 && !idRot) { idRot = ++nvar; snprintf(b, sizeof(b), "<%3d;\"Angular deformation\";LENGTH;FLOAT;%2d;VECTOR;(%8d)>\n", idRot, nbit, nrot); vard += b; }
        const char* name = names[k].c_str();
        if (nrot > 0) {
          if (compl_[k]) snprintf(b, sizeof(b), "    [;\"%s\";[;\"Re\";<%3d><%3d>][;\"Im\";<%3d><%3d>]]\n", name, idDis, idRot, idDis, idRot);
          else snprintf(b, sizeof(b), "    [;\"%s\";<%3d><%3d>]\n", name, idDis, idRot);
        } else {
          if (compl_[k]) snprintf(b, sizeof(b), "    [;\"%s\";[;\"Re\";<%3d>][;\"Im\";<%3d>]]\n", name, idDis, idDis);
          else snprintf(b, sizeof(b), "    [;\"%s\";<%3d>]\n", name, idDis);
        }
        datd += b;
        nval += (long long)(ntra + nrot) * (compl_[k] ? 2 : 1);
      }
      if (vectors_wrapped) datd += "  ]\n";
    }
    if (energy) {   // writeElementsHeader with lStrRes(9) only (:625-752, shell groups :1158-1175, solid groups :1327-1338)
      int idE = 0, ig[64] = {};
      datd += "  [;\"Elements\";\n";
      for (int e = 0; e < nel; ++e) {
        if (elmid[(size_t)e] <= 0 || ptoff[(size_t)e + 1] == ptoff[(size_t)e]) continue;
        const int t = melcon[(size_t)e], key = t == 23 ? 21 : t == 24 ? 22 : t;
        const char* tn = nullptr;
        int nelnod = 0;
        bool shell = false;
        switch (key) {
          case 21: tn = "TRI3"; nelnod = 3; shell = true; break;
          case 22: tn = "QUAD4"; nelnod = 4; shell = true; break;
          case 31: tn = "TRI6"; nelnod = 6; shell = true; break;
          case 32: tn = "QUAD8"; nelnod = 8; shell = true; break;
          case 41: tn = "TET10"; nelnod = 10; break;
          case 42: tn = "WEDG15"; nelnod = 15; break;
          case 43: tn = "HEX20"; nelnod = 20; break;
          case 44: tn = "HEX8"; nelnod = 8; break;
          case 45: tn = "TET4"; nelnod = 4; break;
          case 46: tn = "WEDG6"; nelnod = 6; break;
          default: continue;
        }
        if (!ig[key]) {
          ig[key] = ++nig;
          if (!idE) { idE = ++nvar; snprintf(b, sizeof(b), "<%3d;\"Scaled strain energy density\";FORCE/AREA;FLOAT;%2d;SCALAR>\n", idE, nbit); vard += b; }
          snprintf(b, sizeof(b), "[%3d;\"%s\";\n  [;\"Element nodes\";\n", ig[key], tn); item += b;
          for (const char* side : {"Top", "Bottom", "Basic"}) {
            if (shell == (side[1] == 'a')) continue;   // shells: Top + Bottom, solids: Basic
            snprintf(b, sizeof(b), "    [;\"%s\";\n", side); item += b;
            for (int i = 1; i <= nelnod; ++i) { snprintf(b, sizeof(b), "      [;%2d;<%3d>]\n", i, idE); item += b; }
            item += "    ]\n";
          }
          item += "  ]\n]\n";
        }
        snprintf(b, sizeof(b), "    [;%8d;[%3d]]\n", elmid[(size_t)e], ig[key]); datd += b;
        nval += ptoff[(size_t)e + 1] - ptoff[(size_t)e];
      }
      datd += "  ]\n";
    }
    datd += "}\n";
    hdr = vard + item + datd;
    return nval;
  };
  // openRDBfile (rdbModule.f90:300-318): every new file takes the next increment number
  int next_inc = c.get_int("rdbinc");
  if (multiFiles && next_inc <= 0) next_inc = 1;
  const std::string base_path = file_name("rdbfile", ".frs");
  auto next_path = [&]() {
    std::string path = base_path;
    if (next_inc > 0) {
      const size_t dot = path.rfind('.'), sep = path.rfind('/');
      char t[16];
      snprintf(t, sizeof(t), "_%d", next_inc++);
      if (dot != std::string::npos && dot > 0 && (sep == std::string::npos || dot > sep)) path.insert(dot, t);
      else path += t;
    }
    return path;
  };
  struct OutFile { fsr_frs_writer* w = nullptr; long long nval = 0; };
  std::vector<OutFile> files(multiFiles ? (size_t)nMode + 1 : 1);
  struct WGuard { std::vector<OutFile>& f; ~WGuard() { for (OutFile& x : f) if (x.w) fsr_frs_finish(x.w); } } w_guard{files};
  auto mode_name = [&](int j) { char nm[32]; snprintf(nm, sizeof(nm), "Mode%3d", modeNum[(size_t)j]); return std::string(nm); };
  if (multiFiles) {
    for (int f = 0; f <= nMode; ++f) {
      if (f > 0 && std::find(process[(size_t)f - 1].begin(), process[(size_t)f - 1].end(), (char)1) == process[(size_t)f - 1].end()) {
        log.line("  ** Note: Eigenmode%3d will not be expanded", modeNum[(size_t)f - 1]);
        continue;
      }
      std::string hdr;
      files[(size_t)f].nval = make_header({f == 0 ? std::string("Dynamic response") : mode_name(f - 1)}, {f > 0 && lComplex}, lNodes, true, f > 0 && lEnergy, hdr);
      const std::string path = next_path();
      CHECK(fsr_frs_create_tagged(&files[(size_t)f].w, path.c_str(), "#FEDEM modal data", 0, hdr.c_str(), files[(size_t)f].nval * (nbit / 8)));
      log.line("           --> Results database file: %s (%lld bytes per time step)", path.c_str(), 12 + files[(size_t)f].nval * (nbit / 8));
    }
  } else {
    std::vector<std::string> names{"Dynamic response"};
    std::vector<bool> cx{false};
    for (int j = 0; j < nMode; ++j) { names.push_back(mode_name(j)); cx.push_back(lComplex); }
    std::string hdr;
    files[0].nval = make_header(names, cx, false, true, false, hdr);
    const std::string path = next_path();
    CHECK(fsr_frs_create_tagged(&files[0].w, path.c_str(), "#FEDEM modal data", 0, hdr.c_str(), files[0].nval * (nbit / 8)));
    log.line("           --> Results database file: %s (%lld bytes per time step)", path.c_str(), 12 + files[0].nval * (nbit / 8));
  }

  // --- Time step loop
  log.line("           --> Starting time loop");
  const int nmodes_g = ngen + (lgrav ? 3 : 0), ndim = ndof2 + nmodes_g, ncols = 1 + nMode * ncomp, nall = (int)times.size();
  std::vector<double> Q((size_t)ndim * ncols), sv((size_t)ncols * ndof), supTr(12), eig;
  int ndofs_tot = 0;
  for (int i = 0; i < ntriads; ++i) ndofs_tot += tnd[(size_t)i];
  eig.resize((size_t)std::max(ndofs_tot, 1) * ncomp);
  std::vector<double> geig((size_t)std::max(ngen, 1) * ncomp), rec;
  std::vector<float> rec_f;
  const int npts = lEnergy ? fsr_num_result_points(part) : 0;
  std::vector<double> sig(lEnergy ? (size_t)6 * std::max(npts, 1) : 0), eps(sig.size());
  // writeDisplacementDB (:1437-1515) for the columns [col0, col0 + nc) of sv
  auto put_displacements = [&](int col0, int nc, bool nodes) {
    if (nodes)
      for (int i = 0; i < nnod; ++i) {
        const int j0 = madof[(size_t)i] - 1, nd = madof[(size_t)i + 1] - madof[(size_t)i];
        if (nd < 3) continue;
        for (int l = 0; l < nc; ++l) {
          const double* u = sv.data() + (size_t)(col0 + l) * ndof + j0;
          rec.insert(rec.end(), u, u + (nd > 5 ? 6 : 3));
        }
      }
    if (lVector)
      for (int l = 0; l < nc; ++l) {   // all translations, then all rotations, per component
        const double* u = sv.data() + (size_t)(col0 + l) * ndof;
        for (int i = 0; i < nnod; ++i) {
          if (minex[(size_t)i] < 0) continue;
          const int j0 = madof[(size_t)i] - 1;
          rec.insert(rec.end(), u + j0, u + j0 + 3);
        }
        for (int i = 0; i < nnod; ++i) {
          if (minex[(size_t)i] < 0 || madof[(size_t)i + 1] < madof[(size_t)i] + 6) continue;
          const int j0 = madof[(size_t)i] - 1 + 3;
          rec.insert(rec.end(), u + j0, u + j0 + 3);
        }
      }
  };
  auto write_record = [&](OutFile& f, int step, double time) -> int {
    if ((long long)rec.size() != f.nval) { set_error("internal: modal record of %zu values, header says %lld", rec.size(), f.nval); return FSR_ERR_STATE; }
    if (lDouble) return fsr_frs_write_step(f.w, step, time, rec.data());
    rec_f.assign(rec.begin(), rec.end());
    return fsr_frs_write_step(f.w, step, time, rec_f.data());
  };
  int nerr = 0;
  for (int it = 0; it < nStep; ++it) {
    // ffr_setposition: the first key >= wanted - FLT_EPSILON, clamped to the ends (FFrResultContainer.C:953-1010)
    int idx = -1;
    if (nall > 0) {
      const double wanted = tsteps[(size_t)it];
      if (times[0] > wanted) idx = 0;
      else if (wanted > times[(size_t)nall - 1]) idx = nall - 1;
      else idx = (int)(std::upper_bound(times.begin(), times.begin() + nall, wanted - 1.1920928955078125e-07) - times.begin());
    }
    if (idx < 0 || idx >= nall) { ++nerr; log.line(" *** Error: Error searching for results at time =%12.5E", tsteps[(size_t)it]); continue; }
    std::fill(Q.begin(), Q.end(), 0.0);
    CHECK(fsr_frs_reduced_history(db, isup, ntriads, tb.data(), tnd.data(), tfd.data(), tru.data(), ngen, gen_first, idx, 1, Q.data(), ndim));
    for (int j = 0; j < 12; ++j) supTr[(size_t)j] = (j == 0 || j == 4 || j == 8) ? 1.0 : 0.0;
    if (hsup >= 0) CHECK(fsr_frs_read(db, hsup, idx, 1, supTr.data(), 12, 12));
    if (lgrav)   // g = matmul(grv, sup%supTr(:,1:3)) (modes.f90:364-367)
      for (int j = 0; j < 3; ++j) Q[(size_t)ndof2 + ngen + j] = grv[0] * supTr[3 * j] + grv[1] * supTr[3 * j + 1] + grv[2] * supTr[3 * j + 2];
    int nproc = 0;
    for (int j = 0; j < nMode; ++j) {   // readSupElModes
      if (!process[(size_t)j][(size_t)it]) continue;
      ++nproc;
      size_t off = 0;
      for (int i = 0; i < ntriads; ++i) {
        const int n = tnd[(size_t)i] * ncomp;
        CHECK(fsr_frs_read(db, hm[(size_t)j * (ntriads + 1) + i], idx, 1, eig.data() + off, n, n));
        off += (size_t)n;
      }
      if (ngen > 0) CHECK(fsr_frs_read(db, hm[(size_t)j * (ntriads + 1) + ntriads], idx, 1, geig.data(), ngen * ncomp, ngen * ncomp));
      CHECK(fsr_build_mode_finit(ntriads, supTr.data(), tnd.data(), tfd.data(), eig.data(), ngen, gen_first, geig.data(), ncomp,
                                 Q.data() + (size_t)ndim * (1 + j * ncomp), ndim));
    }
    CHECK(fsr_expand(part, Q.data(), ndim, ncols, sv.data()));
    const int step = stepno[(size_t)idx];
    const double time = times[(size_t)idx];
    if (!multiFiles) {
      rec.clear();
      put_displacements(0, ncols, false);
      CHECK(write_record(files[0], step, time));
    } else {
      rec.clear();
      put_displacements(0, 1, lNodes);
      CHECK(write_record(files[0], step, time));
      for (int j = 0; j < nMode; ++j) {
        if (!process[(size_t)j][(size_t)it]) continue;
        rec.clear();
        put_displacements(1 + j * ncomp, ncomp, lNodes);
        if (lEnergy) {
          // calcStrainEnergyDensity (modesRoutines.f90:219-305): sum(sigma*epsilon) with the shear terms counted twice, of the (real part of
          // the) mode shape, at every result point of every element in SAM order
          CHECK(fsr_recover_step_full(part, Q.data() + (size_t)ndim * (1 + j * ncomp), nullptr, sig.data(), eps.data(), nullptr, nullptr));
          for (int e = 0; e < nel; ++e) {
            if (elmid[(size_t)e] <= 0) continue;
            const int t = melcon[(size_t)e];
            const bool plane = t >= 21 && t <= 24;
            if (!plane && !(t == 31 || t == 32 || (t >= 41 && t <= 46))) continue;
            for (int n = ptoff[(size_t)e]; n < ptoff[(size_t)e + 1]; ++n) {
              const double *s = sig.data() + 6 * (size_t)n, *x = eps.data() + 6 * (size_t)n;
              double v;
              if (std::fabs(s[0]) >= 1.0e300) v = kHuge;   // stress calculation failed, write hugeVal
              else if (plane) v = s[0] * x[0] + s[1] * x[1] + s[2] * x[2] + s[2] * x[2];
              else v = s[0] * x[0] + s[1] * x[1] + s[2] * x[2] + 2.0 * (s[3] * x[3] + s[4] * x[4] + s[5] * x[5]);
              rec.push_back(v);
            }
          }
        }
        CHECK(write_record(files[(size_t)j + 1], step, time));
      }
    }
    log.line("           --> ......Simulation time : %12.5E  (step %d, %d modes expanded)", time, step, nproc);
  }
  log.line("           --> Done time loop. Closing database files");
  for (OutFile& f : files) if (f.w) { fsr_frs_writer* x = f.w; f.w = nullptr; CHECK(fsr_frs_finish(x)); }
  if (nerr) { log.line("\n    %s failed :-(", what); return -nerr; }
  log.line("           ================>  END OF PROGRAM MODES  <================");
  log.line("\n    %s successfully completed :-)  (%.2f s CPU)", what, (double)(clock() - log.t0) / CLOCKS_PER_SEC);
  return 0;
}

// --------------------------------------------------------------------------------------------------------------------
// S-N curve library file (FFpSNCurveLib::readSNCurves / read, FFpSNCurveLib.C:132-277)
namespace {
struct SnCurve {
  std::string name;
  int std_id = 0;
  std::vector<double> loga, m, logN0;
  double thk_exp = 0.0;
};
// FFaTokenizer::createTokens (FFaTokenizer.C:132-189) with '<' '>' ',' on a string that starts with the entry-begin character
std::vector<std::string> ffa_tokens(const std::string& s)
{
  std::vector<std::string> out;
  std::string token;
  int sub = 0;
  bool text = false;
  for (size_t i = 0; i < s.size(); ++i) {
    const int ch = (unsigned char)s[i];
    if (!text && (ch == '<' || ch == '[' || ch == '{')) ++sub;
    if (ch == '"') text = !text;
    if (!(sub == 1 && ch == '"')) {
      if (text) token += (char)ch;
      else if (sub > 1 || (ch != ',' && ch != '<' && ch != '>')) { if (!isspace(ch)) token += (char)ch; }
    }
    if (!text && sub == 1 && (ch == ',' || ch == '>')) { out.push_back(token); token.clear(); }
    if (!text && (ch == '>' || ch == ']' || ch == '}')) --sub;
    if (sub == 0) break;
  }
  return out;
}
}  // namespace

struct fsr_sn_lib {
  std::vector<std::pair<std::string, std::vector<SnCurve>>> stds;
  const SnCurve* curve(int is, int ic) const
  {
    if (is < 0 || (size_t)is >= stds.size() || ic < 0 || (size_t)ic >= stds[(size_t)is].second.size()) return nullptr;
    return &stds[(size_t)is].second[(size_t)ic];
  }
};

extern "C" {

int fsr_sn_read(fsr_sn_lib** lib, const char* path)
{
  if (!lib || !path) { set_error("fsr_sn_read: bad arguments"); return FSR_ERR_ARG; }
  *lib = nullptr;
  FILE* f = fopen(path, "r");
  if (!f) { set_error("Can't open S-N curves file %s", path); return FSR_ERR_ARG; }
  std::string text;
  { char buf[4096]; size_t n; while ((n = fread(buf, 1, sizeof(buf), f)) > 0) text.append(buf, n); }
  fclose(f);
  fsr_sn_lib* L = new fsr_sn_lib;
  size_t i = 0;
  auto skip_line = [&] { while (i < text.size() && text[i] != '\n') ++i; if (i < text.size()) ++i; };
  while (i < text.size()) {
    while (i < text.size() && isspace((unsigned char)text[i])) ++i;   // ignore leading white-space
    if (i >= text.size()) break;
    if (text[i] == '#') { skip_line(); continue; }                   // comment lines
    if (text[i] != '<') { skip_line(); continue; }                   // "Invalid leading character": the rest of the line is ignored
    // one entry: up to the matching end character (quoted text does not count)
    size_t j = i;
    int depth = 0;
    bool quoted = false;
    for (; j < text.size(); ++j) {
      const char ch = text[j];
      if (ch == '"') quoted = !quoted;
      if (quoted) continue;
      if (ch == '<' || ch == '[' || ch == '{') ++depth;
      else if (ch == '>' || ch == ']' || ch == '}') { if (--depth == 0) break; }
    }
    const std::vector<std::string> st = ffa_tokens(text.substr(i, j + 1 - i));
    i = j < text.size() ? j + 1 : j;
    if (st.empty()) { delete L; set_error("%s: empty S-N curve standard entry", path); return FSR_ERR_ARG; }
    std::vector<SnCurve> curves;
    const int std_id = st.size() > 1 ? atoi(st[1].c_str()) : -1;
    for (size_t k = 2; k < st.size(); ++k) {
      const std::vector<std::string> ct = ffa_tokens(st[k]);
      SnCurve cv;
      cv.std_id = std_id;
      if (std_id == 0) {        // NorSok: name, <values>, thickness exponent
        if (ct.size() != 3) continue;
        cv.thk_exp = atof(ct[2].c_str());
      } else if (std_id == 1) { // British: name, <values>
        if (ct.size() != 2) continue;
      } else
        break;                  // unknown curve standard: the rest of the entry is skipped
      cv.name = ct[0];
      const std::vector<std::string> vt = ffa_tokens(ct[1]);
      bool valid = vt.size() % 2 == 0;
      double loga0 = 0.0, m0 = 0.0, logN0 = 0.0;
      for (size_t q = 0; q < vt.size() && valid; q += 2) {
        const double loga1 = atof(vt[q].c_str()), m1 = atof(vt[q + 1].c_str());
        if (m1 < 0.0) valid = false;
        else if (q > 0 && std_id == 0) {   // intersection between the last two line segments
          if (m1 == m0) valid = false;
          else if (loga1 == loga0) continue;   // coincident lines are ignored
          else {
            const double logN1 = (m1 * loga0 - m0 * loga1) / (m1 - m0);
            if (logN1 > loga1 || (q > 2 && logN1 < logN0)) valid = false;
            cv.logN0.push_back(logN1);
            logN0 = logN1;
          }
        }
        cv.loga.push_back(loga1);
        cv.m.push_back(m1);
        loga0 = loga1;
        m0 = m1;
      }
      if (valid) curves.push_back(cv);
    }
    if (!curves.empty()) L->stds.push_back({st[0], curves});
  }
  *lib = L;
  return FSR_OK;
}

void fsr_sn_free(fsr_sn_lib* lib) { delete lib; }
int fsr_sn_num_standards(const fsr_sn_lib* lib) { return lib ? (int)lib->stds.size() : 0; }
int fsr_sn_num_curves(const fsr_sn_lib* lib, int is) { return lib && is >= 0 && (size_t)is < lib->stds.size() ? (int)lib->stds[(size_t)is].second.size() : 0; }

int fsr_sn_get(const fsr_sn_lib* lib, int is, int ic, int* std_id, double* loga, double* m, double* logN0, int cap)
{
  const SnCurve* cv = lib ? lib->curve(is, ic) : nullptr;
  if (!cv) { set_error("S-N curve index (%d, %d) is out of range", is, ic); return FSR_ERR_ARG; }
  if (std_id) *std_id = cv->std_id;
  const int n = (int)cv->loga.size();
  for (int k = 0; k < n && k < cap; ++k) {
    if (loga) loga[k] = cv->loga[(size_t)k];
    if (m) m[k] = cv->m[(size_t)k];
    if (logN0 && (size_t)k < cv->logN0.size()) logN0[k] = cv->logN0[(size_t)k];
  }
  return n;
}

// FFpSNCurveNorSok::getValue / FFpSNCurveBritish::getValue (FFpSNCurve.C:22-47)
double fsr_sn_value(const fsr_sn_lib* lib, int is, int ic, double sr)
{
  const SnCurve* cv = lib ? lib->curve(is, ic) : nullptr;
  if (!cv || cv->loga.empty()) return -1.0;
  if (cv->std_id == 1) return cv->loga.size() > 1 ? pow(10.0, cv->loga[0] - cv->loga[1] * cv->m[1] - cv->m[0] * log10(sr)) : -1.0;
  const size_t n = cv->logN0.size();
  if (cv->loga.size() <= n) return -1.0;
  for (size_t k = 0; k < n; ++k) {
    const double logN = cv->loga[k] - cv->m[k] * log10(sr);
    if (logN < cv->logN0[k]) return pow(10.0, logN);
  }
  return pow(10.0, cv->loga[n] - cv->m[n] * log10(sr));
}

}  // extern "C"

// --------------------------------------------------------------------------------------------------------------------
// The fedem_fpp specific part (fpp.f90:127-142,206-520): strain coat elements of the FE part as rosettes in their element systems,
// running summary (and rainflow damage) on the GPU, the strain coat results database (saveStrainCoatModule.f90).
static int fpp_part(CmdLine& c, Log& log, const char* what, fsr_part* part, fsr_ftl* ftl, fsr_frs* db, int isup, int user_id, const char* descr,
                    const char* model_file, const std::string& linkfile, const std::vector<double>& xyz, int ndof2, int ngen, int ntriads,
                    const std::vector<int>& tb, const std::vector<int>& tnd, const std::vector<int>& tfd, const std::vector<double>& tru,
                    int gen_first, const std::vector<int>& sel, int nsel, const std::vector<int>& stepno, const std::vector<double>& times,
                    bool lgrav, const double* grv, const std::vector<int>& madof)
{
  (void)madof;
  const int iprint = c.get_int("debug"), iSurface = c.get_int("surface"), angBinSize = c.get_int("angleBins"), fppType = c.get_int("HistDataType");
  const bool oldRange = c.get_bool("oldRange"), writeHistory = c.get_bool("writeHistory"), lDouble = c.get_bool("double");
  const double pvxGate = (double)(float)c.get_double("PVXGate"), biAxialGate = c.get_double("biAxialGate"), toMPa = c.get_double("stressToMPaScale");
  const double startTime = c.get_double("statm"), stopTime = c.get_double("stotm");
  if (fppType < 0) FAIL("-HistDataType %d asks for the nCode FPP plug-in (-fppfile), which is not part of this build", fppType);
  if (angBinSize < 2) FAIL("Invalid value on option -angleBins: %d", angBinSize);

  const int nStrainCoatTotal = fsr_ftl_num_strain_coats(ftl);
  if (nStrainCoatTotal < 1) FAIL("Link does not contain any strain coat elements");
  log.line("               Number of strain coats =%6d", nStrainCoatTotal);
  std::vector<fsr_strain_coat> coats((size_t)nStrainCoatTotal);
  CHECK(fsr_ftl_get_strain_coats(ftl, coats.data(), nStrainCoatTotal));

  // fatigueInit (fatigueModule.f90:14-37): the S-N curve library
  fsr_sn_lib* sn = nullptr;
  struct SnGuard { fsr_sn_lib*& p; ~SnGuard() { fsr_sn_free(p); } } sn_guard{sn};
  if (fppType > 0) {
    if (fsr_sn_read(&sn, c.get_string("SNfile").c_str()) < 0) {
      log.line(" *** Error: %s", fsr_last_error());
      FAIL("Failure in fatigue initialization");
    }
  }

  // --- initiateStrainCoats (strainCoatModule.f90:172-312) + calcElmCoordSystem with useElCoordSys (strainRosetteModule.f90:506-567):
  //     one rosette per kept result set, in the centroid of the coat nodes, axes = the globalized element axes
  struct Point { int coat, set, mat, sn[2]; double scf, zpos; bool fat; };
  std::vector<fsr_rosette> ros;
  std::vector<Point> pts;
  std::vector<int> first_pt((size_t)nStrainCoatTotal + 1, 0), coat_fpp((size_t)nStrainCoatTotal, 0);
  bool haveSNdata = false;
  int ndegenerate = 0;
  for (int i = 0; i < nStrainCoatTotal; ++i) {
    const fsr_strain_coat& k = coats[(size_t)i];
    first_pt[(size_t)i] = (int)pts.size();
    if (k.nnod < 3 || k.nnod > 4) FAIL("Invalid strain coat definition. Too many nodes or points: %d %d (element %d)", k.nnod, k.npts, k.id);
    for (int j = 0; j < k.nnod; ++j) if (k.nodes[j] < 1) FAIL("Non-existing node referenced by strain coat element %d", k.id);
    int snmin = 0;
    for (int j = 0; j < k.npts; ++j) snmin = std::min(snmin, std::min(k.sn_curve[j][0], k.sn_curve[j][1]));
    coat_fpp[(size_t)i] = (k.npts > 0 && snmin >= 0) ? fppType : 0;
    double X[4][3], T[9];
    for (int j = 0; j < k.nnod; ++j) for (int d = 0; d < 3; ++d) X[j][d] = xyz[3 * (size_t)(k.nodes[j] - 1) + d];
    if (shell_element_axes(k.nnod, X, T, T + 3, T + 6, true)) { ++ndegenerate; continue; }   // degenerated strain coat element, just ignore it
    for (int j = 0; j < k.npts; ++j) {
      if (iSurface > 0 && iSurface != k.res_set[j]) continue;
      fsr_rosette R;
      memset(&R, 0, sizeof(R));
      R.id = i + 1;
      R.numnod = k.nnod;
      for (int q = 0; q < k.nnod; ++q) R.nodes[q] = k.nodes[q];
      for (int d = 0; d < 3; ++d) {
        R.rpos[d] = T[d]; R.rpos[3 + d] = T[3 + d]; R.rpos[6 + d] = T[6 + d];
        double sum = 0.0;
        for (int q = 0; q < k.nnod; ++q) sum += X[q][d];
        R.rpos[9 + d] = sum / k.nnod;
      }
      R.zpos = k.zpos[j]; R.emod = k.emod[j]; R.nu = k.nu[j];
      Point P;
      P.coat = i; P.set = k.res_set[j]; P.mat = k.mat_id[j]; P.sn[0] = k.sn_curve[j][0]; P.sn[1] = k.sn_curve[j][1]; P.scf = k.scf[j]; P.zpos = k.zpos[j];
      P.fat = coat_fpp[(size_t)i] > 0;
      if (P.fat) {
        haveSNdata = true;
        // ffp_calcDamage (FFpFatigue_F.C:53-78): the curve of the S-N library; one or two line segments (British: one line)
        int std_id = 0;
        double la[2], mm[2], ln0[2];
        const int nseg = fsr_sn_get(sn, P.sn[0], P.sn[1], &std_id, la, mm, ln0, 2);
        if (nseg < 1) { log.line(" *** Error: %s", fsr_last_error()); FAIL("Failure in damage calculation: Strain coat %d has no valid S-N curve", k.id); }
        if (std_id == 1) {
          if (nseg < 2) FAIL("Failure in damage calculation: British S-N curve (%d, %d) lacks its second parameter pair", P.sn[0], P.sn[1]);
          R.sncurve[0] = R.sncurve[1] = la[0] - la[1] * mm[1]; R.sncurve[2] = R.sncurve[3] = mm[0];
        } else if (nseg == 1) { R.sncurve[0] = R.sncurve[1] = la[0]; R.sncurve[2] = R.sncurve[3] = mm[0]; }
        else if (nseg == 2) { R.sncurve[0] = la[0]; R.sncurve[1] = la[1]; R.sncurve[2] = mm[0]; R.sncurve[3] = mm[1]; }
        else FAIL("S-N curve (%d, %d) has %d line segments; curves with more than two are not part of this build", P.sn[0], P.sn[1], nseg);
      }
      ros.push_back(R);
      pts.push_back(P);
    }
  }
  first_pt[(size_t)nStrainCoatTotal] = (int)pts.size();
  const int npt = (int)pts.size();
  if (ndegenerate) log.line("  ** Warning: %d degenerated strain coat elements are ignored", ndegenerate);
  if (npt == 0) FAIL("No strain coat result points to process (check the -surface option)");
  static const char* kSet[4] = {"Basic", "Bottom", "Mid", "Top"};
  if (iprint > 0) {   // printStrainCoatInput (strainCoatModule.f90:550-596)
    log.line("\n\n     STRAIN COAT PROPERTY SUMMARY\n     --------------------------------------------------------------------------");
    log.line("     Strain Coat  Result set  Z-position  Material Group  S-N curve  SCF factor");
    for (int i = 0; i < nStrainCoatTotal; ++i)
      for (int q = first_pt[(size_t)i]; q < first_pt[(size_t)i + 1]; ++q) {
        const Point& P = pts[(size_t)q];
        if (q == first_pt[(size_t)i]) log.line("%8d%8d%10s%14.5E%10d%11d%3d%12.3f", i + 1, coats[(size_t)i].id, kSet[P.set & 3], P.zpos, P.mat, P.sn[0], P.sn[1], P.scf);
        else log.line("              %12s%14.5E%10d%11d%3d%12.3f", kSet[P.set & 3], P.zpos, P.mat, P.sn[0], P.sn[1], P.scf);
      }
    log.line("    ---------------------------------------------------------------------------\n");
  }

  // --- the reduced history of the selected steps, read once (readSupElDisplacements + BuildFinit); every block of coats runs over it
  log.line("           --> Reading the reduced history of %d time steps", nsel);
  const int ndim = ndof2 + ngen + (lgrav ? 3 : 0);
  std::vector<double> Qall((size_t)ndim * std::max(nsel, 1), 0.0);
  for (int k = 0; k < nsel;) {
    int run = 1;
    while (k + run < nsel && sel[(size_t)(k + run)] == sel[(size_t)k] + run) ++run;
    const int rch = fsr_frs_reduced_history(db, isup, ntriads, tb.data(), tnd.data(), tfd.data(), tru.data(), ngen, gen_first, sel[(size_t)k], run,
                                            Qall.data() + (size_t)k * ndim, ndim);
    CHECK(rch);
    if (rch > 0 && k == 0) log.line("  ** Warning: %s", fsr_last_error());
    k += run;
  }
  if (lgrav)   // the strains of the static gravitation deflection vgii = dis1Expand(V . g) (fpp.f90:170-192), as in fedem_gage
    for (int k = 0; k < nsel; ++k) for (int j = 0; j < 3; ++j) Qall[(size_t)k * ndim + ndof2 + ngen + j] = grv[j];

  // --- Element block loop: blocks of whole coats, bounded by the device memory of the angle bins (5 x 8 bytes per bin and point)
  std::vector<double> env((size_t)8 * npt), summary((size_t)6 * npt), damage((size_t)npt, 0.0);
  std::vector<int> nbiax((size_t)npt, 0);
  size_t free_b = 0, total_b = 0;
  cudaMemGetInfo(&free_b, &total_b);
  const long long per_pt = 44LL * (angBinSize - 1) + 4096;
  const int max_pts = (int)std::max<long long>(1024, std::min<long long>(1 << 20, (long long)(free_b / 4) / per_pt));
  const int nbit = lDouble ? 64 : 32;
  int next_inc = c.get_int("rdbinc");
  const std::string base_path = file_name("rdbfile", ".frs");
  auto next_path = [&]() {   // openRDBfile (rdbModule.f90:300-318)
    std::string path = base_path;
    if (next_inc > 0) {
      const size_t dot = path.rfind('.'), sep = path.rfind('/');
      char t[16];
      snprintf(t, sizeof(t), "_%d", next_inc++);
      if (dot != std::string::npos && dot > 0 && (sep == std::string::npos || dot > sep)) path.insert(dot, t);
      else path += t;
    }
    return path;
  };
  // the header files of rdbModule (ivard, iitem, idatd) with the ids of saveStrainCoatModule (id(15), idSTRC(4,0:3))
  struct Header {
    std::string vard, item, datd;
    int nvar = 2, nig = 0, id[16] = {}, idSTRC[5][4] = {};
    long long nval = 0;
  };
  auto begin_header = [&](Header& H) {
    H = Header();
    H.vard = rdb_file_preamble("fedem_fpp", model_file, linkfile.c_str(), "strain coat data base file");
    H.vard += "VARIABLES:\n<1;\"Time step number\";NONE;INT;32;NUMBER>\n<2;\"Physical time\";TIME;FLOAT;64;SCALAR>\n";
    H.datd = "DATABLOCKS:\n<1><2>\n{\"Part\";";
    char b[512];
    if (isup > 0) { snprintf(b, sizeof(b), "%d;", isup); H.datd += b; } else H.datd += ";";
    if (user_id > 0) { snprintf(b, sizeof(b), "%d;", user_id); H.datd += b; } else H.datd += ";";
    if (descr && *descr) { snprintf(b, sizeof(b), "\"%s\";\n", descr); H.datd += b; } else H.datd += ";\n";
    H.datd += "  [;\"Elements\";\n";
  };
  // writeStrainCoatHeader (saveStrainCoatModule.f90:172-330) for the coats [c0, c1)
  auto add_coats = [&](Header& H, int c0, int c1, bool hist) -> int {
    char b[256];
    auto vardef = [&](int k, const char* name, const char* unit) {
      int& idv = H.id[k];
      if (idv > 0) return;
      if (idv < 0) { idv = -idv; return; }
      idv = ++H.nvar;
      snprintf(b, sizeof(b), "<%d;\"%s\";%s;FLOAT;%2d;SCALAR>\n", idv, name, unit, nbit); H.vard += b;
    };
    auto itemgroup = [&](int& idg, int nvar, int rset) {
      H.nval += nvar;
      if (idg > 0) return;
      idg = ++H.nig;
      snprintf(b, sizeof(b), "[%d;\"%s\";", idg, kSet[rset]); H.item += b;
      for (int k = 1; k <= nvar; ++k) if (H.id[k] > 0) { snprintf(b, sizeof(b), "<%d>", H.id[k]); H.item += b; }
      H.item += "]\n";
    };
    for (int i = c0; i < c1; ++i) {
      const int q0 = first_pt[(size_t)i], q1 = first_pt[(size_t)i + 1];
      if (q1 == q0) continue;   // empty strain coat (switched off with -surface, or degenerated)
      std::string vars;
      for (int q = q0; q < q1; ++q) {
        const Point& P = pts[(size_t)q];
        if (P.set < 0 || P.set > 3) { set_error("writeStrainCoatHeader: invalid result set"); return FSR_ERR_STATE; }
        int itemG = 1;
        if (hist) {
          static const char* nm[10] = {"Max principal stress", "Min principal stress", "Signed abs max stress", "Max shear stress", "Von Mises stress",
                                       "Max principal strain", "Min principal strain", "Signed abs max strain", "Max shear strain", "Von Mises strain"};
          for (int k = 0; k < 10; ++k) vardef(k + 1, nm[k], k < 5 ? "FORCE/AREA" : "NONE");
          itemgroup(H.idSTRC[1][P.set], 10, P.set);
        } else {
          vardef(1, "Max principal stress", "FORCE/AREA"); vardef(2, "Max shear stress", "FORCE/AREA");
          vardef(3, "Max stress range", "FORCE/AREA"); vardef(4, "Max von Mises stress", "FORCE/AREA");
          vardef(5, "Max principal strain", "NONE"); vardef(6, "Max shear strain", "NONE");
          vardef(7, "Max strain range", "NONE"); vardef(8, "Max von Mises strain", "NONE");
          if (coat_fpp[(size_t)i] > 0) {
            itemG = 2;
            vardef(9, "Damage", "NONE"); vardef(10, "Life (equnits)", "TIME"); vardef(11, "Life (repeats)", "NONE");
          } else {
            for (int k = 9; k <= 11; ++k) H.id[k] = -std::abs(H.id[k]);
            H.nval -= 3;
          }
          vardef(12, "Angle spread", "ANGLE"); vardef(13, "Most popular angle", "ANGLE");
          if (nbiax[(size_t)q] > 0) {
            itemG += 2;
            vardef(14, "Mean bi-axiality", "NONE"); vardef(15, "Biaxiality standard deviation", "NONE");
            itemgroup(H.idSTRC[itemG][P.set], 15, P.set);
          } else
            itemgroup(H.idSTRC[itemG][P.set], 13, P.set);
        }
        snprintf(b, sizeof(b), "[%d]", H.idSTRC[itemG][P.set]); vars += b;
      }
      snprintf(b, sizeof(b), "    [;%8d;[;\"%s\";[;\"Element\";%s]]]\n", coats[(size_t)i].id, coats[(size_t)i].nnod == 3 ? "STRCT3" : "STRCQ4", vars.c_str());
      H.datd += b;
    }
    return FSR_OK;
  };
  Header Hsum;
  if (!writeHistory) { log.line("           --> Initializing result database headers"); begin_header(Hsum); }
  std::vector<double> record;   // the one summary record, block after block

  std::vector<double> vals;
  for (int c0 = 0; c0 < nStrainCoatTotal;) {
    int c1 = c0;
    while (c1 < nStrainCoatTotal && (c1 == c0 || first_pt[(size_t)c1 + 1] - first_pt[(size_t)c0] <= max_pts)) ++c1;
    const int p0 = first_pt[(size_t)c0], np = first_pt[(size_t)c1] - p0;
    log.line("               Processing elements%8d  to%8d  of total%8d elements", c0 + 1, c1, nStrainCoatTotal);
    if (np > 0) {
      log.line("           --> Computing strain-displ matrices");
      fsr_gages* gages = nullptr;
      CHECK(fsr_gage_create(&gages, part, ros.data() + p0, np));
      struct GagesGuard { fsr_gages* p; ~GagesGuard() { fsr_gage_destroy(p); } } gages_guard{gages};
      fsr_frs_writer* hw = nullptr;
      struct HGuard { fsr_frs_writer*& p; ~HGuard() { if (p) fsr_frs_finish(p); } } h_guard{hw};
      if (writeHistory) {   // writeHistoryHeader (saveStrainCoatModule.f90:27-62): one file per block
        log.line("           --> Writing result database headers");
        Header H;
        begin_header(H);
        CHECK(add_coats(H, c0, c1, true));
        H.datd += "  ]\n}\n";
        const std::string path = next_path();
        CHECK(fsr_frs_create_tagged(&hw, path.c_str(), "#FEDEM strain coat data", 0, (H.vard + H.item + H.datd).c_str(), H.nval * (nbit / 8)));
        log.line("           --> Results database file: %s (%lld bytes per time step)", path.c_str(), 12 + H.nval * (nbit / 8));
      }
      log.line("           --> Starting time loop");
      CHECK(fsr_coat_begin(gages, angBinSize, biAxialGate));
      if (!writeHistory) CHECK(fsr_coat_feed(gages, Qall.data(), ndim, nsel));
      else {
        // calcStrainCoatData + writeHistoryDB / writeRosetteDB (:332-398): sigmaP(3), tauMax, sigmaVM, epsP(3), gammaMax, epsVM per result point
        const int window = 256;
        vals.resize((size_t)window * np * FSR_GAGE_NVAL);
        std::vector<double> rec_d((size_t)10 * np);
        std::vector<float> rec_f((size_t)10 * np);
        for (int w0 = 0; w0 < nsel; w0 += window) {
          const int nw = std::min(window, nsel - w0);
          CHECK(fsr_coat_feed(gages, Qall.data() + (size_t)w0 * ndim, ndim, nw));
          CHECK(fsr_gage_recover(gages, Qall.data() + (size_t)w0 * ndim, ndim, nw, vals.data()));
          for (int k = 0; k < nw; ++k) {
            size_t n = 0;
            for (int r = 0; r < np; ++r) {
              const double* v = vals.data() + ((size_t)k * np + r) * FSR_GAGE_NVAL;
              const double out[10] = {v[13], v[14], v[15], v[16], v[17], v[3], v[4], v[5], v[6], v[7]};
              for (double x : out) { rec_d[n] = x; rec_f[n] = (float)x; ++n; }
            }
            CHECK(fsr_frs_write_step(hw, stepno[(size_t)sel[(size_t)(w0 + k)]], times[(size_t)sel[(size_t)(w0 + k)]],
                                     lDouble ? (const void*)rec_d.data() : (const void*)rec_f.data()));
          }
        }
      }
      // fsr_coat_end returns [8][np] / [6][np] blocks: fetch them block-wise and scatter into the part-wide arrays
      {
        std::vector<double> e8((size_t)8 * np), s6((size_t)6 * np);
        std::vector<int> nb((size_t)np);
        CHECK(fsr_coat_end(gages, e8.data(), s6.data(), nb.data()));
        for (int r = 0; r < np; ++r) {
          for (int k = 0; k < 8; ++k) env[(size_t)k * npt + p0 + r] = e8[(size_t)k * np + r];
          for (int k = 0; k < 6; ++k) summary[(size_t)k * npt + p0 + r] = s6[(size_t)k * np + r];
          nbiax[(size_t)(p0 + r)] = nb[(size_t)r];
        }
      }
      log.line("           --> Time loop done. Closing fpp processors");
      if (hw) { fsr_frs_writer* x = hw; hw = nullptr; CHECK(fsr_frs_finish(x)); }
      bool any_fat = false;
      for (int r = 0; r < np; ++r) any_fat = any_fat || pts[(size_t)(p0 + r)].fat;
      if (any_fat && nsel > 0) {
        // fatigueAddPoint + fatigueDamage (fatigueModule.f90:40-118): PVX + rainflow + Miner sum of fatValue on the point's S-N curve
        std::vector<double> scf((size_t)np);
        for (int r = 0; r < np; ++r) scf[(size_t)r] = pts[(size_t)(p0 + r)].fat ? pts[(size_t)(p0 + r)].scf : 1.0;
        CHECK(fsr_gage_set_coat_fatigue(gages, scf.data()));
        const double curve[4] = {15.117, 17.146, 4.0, 5.0};
        std::vector<double> dmg((size_t)4 * np);
        std::vector<int> ncyc((size_t)4 * np), status((size_t)4 * np);
        const int nw = fsr_gage_fatigue(gages, Qall.data(), ndim, nsel, toMPa, pvxGate, curve, 1.0, 0, dmg.data(), ncyc.data(), nullptr, status.data());
        CHECK(nw);
        for (int r = 0; r < np; ++r) {
          const Point& P = pts[(size_t)(p0 + r)];
          if (!P.fat) continue;
          if (status[4 * (size_t)r] == 2) FAIL("Failure in damage calculation: Strain coat %d (turning point stack exhausted)", coats[(size_t)P.coat].id);
          damage[(size_t)(p0 + r)] = P.scf == 0.0 ? 0.0 : dmg[4 * (size_t)r];
        }
      }
    }
    if (!writeHistory) {
      // writeElementsHeader + writeElementsDB / writeStrainCoatDB (:141-170,400-558) of this block
      CHECK(add_coats(Hsum, c0, c1, false));
      const double time = stopTime - startTime;
      for (int q = p0; q < p0 + np; ++q) {
        const Point& P = pts[(size_t)q];
        auto E = [&](int k) { return env[(size_t)k * npt + q]; };   // epsMax, epsMin, sigMax, sigMin, gammaMax, tauMax, vmeMax, vmsMax
        auto S = [&](int k) { return summary[(size_t)k * npt + q]; };
        record.push_back(std::fabs(E(2)) > std::fabs(E(3)) ? E(2) : E(3));
        record.push_back(E(5));
        record.push_back(oldRange ? E(2) - E(3) : S(0));
        record.push_back(E(7));
        record.push_back(std::fabs(E(0)) > std::fabs(E(1)) ? E(0) : E(1));
        record.push_back(E(4));
        record.push_back(oldRange ? E(0) - E(1) : S(1));
        record.push_back(E(6));
        if (coat_fpp[(size_t)P.coat] > 0) {
          const double d = damage[(size_t)q];
          if (d > 0.0) { record.push_back(d); record.push_back(time / d); record.push_back(1.0 / d); }
          else { record.push_back(1.0e20); record.push_back(1.0e20); record.push_back(1.0e20); }   // no damage = infinite life
        }
        record.push_back(S(3));
        record.push_back(S(2));
        if (nbiax[(size_t)q] > 0) { record.push_back(S(4)); record.push_back(S(5)); }
      }
    }
    c0 = c1;
  }
  if (iprint > 0) {   // printStrainCoatData (strainCoatModule.f90:599-666)
    log.line("\n\n     STRAIN COAT RECOVERY SUMMARY\n     ------------------------   --------------- Stress --------------   --------------- Strain --------------"
             "   ------ Biaxiality ------   - Principal dir. angle -");
    log.line("     Strain Coat  Result set      Max P1      Max shear  Max von Mises    Max P1      Max shear  Max von Mises     Mean      Std. Dev."
             "      Most popular    spread");
    for (int i = 0; i < nStrainCoatTotal; ++i) {
      for (int q = first_pt[(size_t)i]; q < first_pt[(size_t)i + 1]; ++q) {
        const Point& P = pts[(size_t)q];
        auto E = [&](int k) { return env[(size_t)k * npt + q]; };
        auto S = [&](int k) { return summary[(size_t)k * npt + q]; };
        char head[40];
        if (q == first_pt[(size_t)i]) snprintf(head, sizeof(head), "%8d%8d%10s", i + 1, coats[(size_t)i].id, kSet[P.set & 3]);
        else snprintf(head, sizeof(head), "              %12s", kSet[P.set & 3]);
        log.line("%s    %13.5E%13.5E%13.5E %13.5E%13.5E%13.5E %13.5E%13.5E %13.5f%13.5f", head, E(2), E(5), E(7), E(0), E(4), E(6), S(4), S(5), S(2), S(3));
      }
      if (coat_fpp[(size_t)i] > 0 && first_pt[(size_t)i + 1] > first_pt[(size_t)i]) {
        std::string l = "                    Damage =  ";
        char b[32];
        for (int q = first_pt[(size_t)i]; q < first_pt[(size_t)i + 1]; ++q) { snprintf(b, sizeof(b), "%13.5E", damage[(size_t)q]); l += b; }
        log.line("%s", l.c_str());
      }
    }
    log.line("    ---------------------------------------------------------------------------------------------------------------------------------------------------------------\n");
  }
  log.line("           --> Block loop done. Closing database files");
  if (!haveSNdata && fppType > 0)
    log.line("  ** Warning: None of the element groups processed were assigned an S-N curve.\n"
             "              Damage and Life contour plots are therefore not created for this part.");
  if (!writeHistory) {   // finalizeStrainCoatHeader (:118-138) + the one record (writeTimeStepDB(rdb,1,stopTime), fpp.f90:222)
    if ((long long)record.size() != Hsum.nval) FAIL("internal: strain coat summary record of %zu values, header says %lld", record.size(), Hsum.nval);
    Hsum.datd += "  ]\n}\n";
    const std::string path = next_path();
    fsr_frs_writer* w = nullptr;
    CHECK(fsr_frs_create_tagged(&w, path.c_str(), "#FEDEM strain coat data", 0, (Hsum.vard + Hsum.item + Hsum.datd).c_str(), Hsum.nval * (nbit / 8)));
    int rcw;
    if (lDouble) rcw = fsr_frs_write_step(w, 1, stopTime, record.data());
    else { std::vector<float> rf(record.begin(), record.end()); rcw = fsr_frs_write_step(w, 1, stopTime, rf.data()); }
    const int rcf = fsr_frs_finish(w);
    CHECK(rcw);
    CHECK(rcf);
    log.line("           --> Results database file: %s (%lld values)", path.c_str(), Hsum.nval);
  }
  log.line("           ================>  END OF PROGRAM FPP  <================");
  log.line("\n    %s successfully completed :-)  (%.2f s CPU)", what, (double)(clock() - log.t0) / CLOCKS_PER_SEC);
  return 0;
}

extern "C" {

int solveStress(void) { return run_program(0); }
int solveGage(void) { return run_program(1); }
int solveModes(void) { return run_program(2); }
int solveFpp(void) { return run_program(3); }

}  // extern "C"
