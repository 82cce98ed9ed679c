// ref_shim_ffl.cpp -- TEST INFRASTRUCTURE (see oracle.h).  A thin extern "C" face over the reference's
// OWN, unmodified FE-model reader and Fortran accessor layer: fedem-foundation/src/FFlLib (link handler,
// FE parts, the .ftl reader FFlFedemReader.C) and FFlLinkHandler_F.C -- the very ffl_getsize / ffl_getnodes /
// ffl_gettopol / ffl_getcoor / ffl_getmat / ffl_getthick / ffl_getbeamsection / ffl_getpinflags /
// ffl_getelmid entry points fedem_stress calls per element -- plus the reference's command-line parser
// (FFaCmdLineArg), compiled where they lie under /root/reference by oracle/Makefile into
// oracle/_ref/libfedem_ref_ffl.so.  No reference source is copied.  Used only by tests/ to check the
// product's .ftl reader (csrc/io_ftl.cu) and option parser (csrc/cmdline.hpp) value by value.
//
// The few symbols FFlLib references from parts of the reference that are not on this path (VTF export,
// the other FE file formats, the build-stamp module) are stubbed here.
#include <cstring>
#include <string>
#include <vector>

#include "FFaLib/FFaCmdLineArg/FFaCmdLineArg.H"
#include "FFlLib/FFlIOAdaptors/FFlFedemReader.H"
#include "FFlLib/FFlIOAdaptors/FFlReaders.H"
#include "FFlLib/FFlIOAdaptors/FFlVTFWriter.H"
#include "FFlLib/FFlLinkHandler.H"
#include "FFpLib/FFpFatigue/FFpSNCurve.H"
#include "FFpLib/FFpFatigue/FFpSNCurveLib.H"

// ---- stubs for off-path parts --------------------------------------------------------------------
namespace FedemAdmin { const char* getCopyrightString() { return "checker build"; } }
namespace FFl {
void initAllReaders() { static bool done = false; if (!done) FFlFedemReader::init(); done = true; }
void releaseAllReaders() {}
}
FFlVTFWriter::FFlVTFWriter(const FFlLinkHandler* l) : FFlWriterBase(l), myFile(NULL) {}
FFlVTFWriter::~FFlVTFWriter() {}
bool FFlVTFWriter::write(const std::string&, const std::string&, int, int) { return false; }

// the reference's Fortran-callable accessors (FFlLinkHandler_F.C); gfortran name mangling
extern "C" {
void ffl_full_init_(const char* linkFile, const char* elmGroups, int& ierr, const int ncharF, const int ncharG);
void ffl_release_(const int& removeSingletons);
}

extern "C" {

// defines the options FFlLinkHandler_F.C queries and parses `argv` with the reference's own parser
void ref_cmdline_init(int argc, char** argv)
{
  FFaCmdLineArg::removeInstance();
  FFaCmdLineArg::init(argc, argv);
  FFaCmdLineArg::instance()->addOption("useANDESformulation", false, "use ANDES shells");
  FFaCmdLineArg::instance()->addOption("linkfile", std::string(""), "link file");
  FFaCmdLineArg::instance()->addOption("group", std::string(""), "element groups");
}

void ref_cmdline_add_int(const char* name, int v) { FFaCmdLineArg::instance()->addOption(name, v, "int option"); }
void ref_cmdline_add_double(const char* name, double v) { FFaCmdLineArg::instance()->addOption(name, v, "double option"); }
void ref_cmdline_add_bool(const char* name, int v) { FFaCmdLineArg::instance()->addOption(name, v != 0, "bool option"); }
void ref_cmdline_add_string(const char* name, const char* v) { FFaCmdLineArg::instance()->addOption(name, std::string(v), "string option"); }
int ref_cmdline_get_int(const char* name) { int v = 0; FFaCmdLineArg::instance()->getValue(name, v); return v; }
double ref_cmdline_get_double(const char* name) { double v = 0; FFaCmdLineArg::instance()->getValue(name, v); return v; }
int ref_cmdline_get_bool(const char* name) { bool v = false; FFaCmdLineArg::instance()->getValue(name, v); return v ? 1 : 0; }
int ref_cmdline_get_string(const char* name, char* out, int cap)
{
  std::string v;
  FFaCmdLineArg::instance()->getValue(name, v);
  strncpy(out, v.c_str(), cap - 1);
  out[cap - 1] = 0;
  return (int)v.size();
}
int ref_cmdline_read_file(const char* path) { return FFaCmdLineArg::instance()->readOptionsFile(path) ? 1 : 0; }
int ref_cmdline_is_set(const char* name) { return FFaCmdLineArg::instance()->isOptionSetOnCmdLine(name) ? 1 : 0; }

// ffl_full_init: read the .ftl file and activate the -group selection (stress.f90:111 -> ffl_init)
int ref_ffl_load(const char* path, const char* groups)
{
  if (FFaCmdLineArg::empty()) ref_cmdline_init(0, NULL);
  int ierr = 0;
  ffl_full_init_(path, groups ? groups : "", ierr, (int)strlen(path), groups ? (int)strlen(groups) : 0);
  return ierr;
}

void ref_ffl_release() { const int all = 0; ffl_release_(all); }

// the reference's S-N curve library reader (FFpSNCurveLib.C), as ffp_initfatigue / ffp_calcdamage use it
int ref_sn_read(const char* path) { return FFpSNCurveLib::instance()->readSNCurves(path) ? 1 : 0; }
int ref_sn_num_standards() { return (int)FFpSNCurveLib::instance()->getNoCurveStds(); }
int ref_sn_num_curves(int is) { return (int)FFpSNCurveLib::instance()->getNoCurves(is); }
int ref_sn_get(int is, int ic, int* std_id, double* loga, double* m, int cap)
{
  FFpSNCurve* c = FFpSNCurveLib::instance()->getCurve((long int)is, (long int)ic);
  if (!c) return -1;
  *std_id = (int)c->getStdId();
  for (size_t k = 0; k < c->loga.size() && (int)k < cap; k++) { loga[k] = c->loga[k]; m[k] = c->m[k]; }
  return (int)c->loga.size();
}
double ref_sn_value(int is, int ic, double s)
{
  FFpSNCurve* c = FFpSNCurveLib::instance()->getCurve((long int)is, (long int)ic);
  return c ? c->getValue(s) : -1.0;
}

}  // extern "C"
