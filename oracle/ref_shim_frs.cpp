// ref_shim_frs.cpp -- TEST INFRASTRUCTURE (see oracle.h).  A thin extern "C" face over the reference's
// OWN, unmodified .frs reader (fedem-foundation/src/FFrLib/*.C and the FFaLib files it needs, compiled
// where they lie under /root/reference by oracle/Makefile into oracle/_ref/libfedem_ref_frs.so).  It
// reproduces the call sequence of ffr_init / ffr_findptr / ffr_setposition / ffr_increment /
// ffr_getdata (FFrExtractor_F.C:33-263) without the FFaCmdLineArg singleton.  No reference source is
// copied.  Used only by tests/ to check the product's .frs reader (csrc/io_frs.cu) value by value.
#include <cfloat>
#include <set>
#include <string>
#include <vector>

#include "FFaLib/FFaDefinitions/FFaResultDescription.H"
#include "FFrLib/FFrExtractor.H"

extern "C" {

void* ref_frs_open(const char* const* files, int n)
{
  FFrExtractor* rdb = new FFrExtractor("checker");
  for (int i = 0; i < n; ++i)
    if (!rdb->addFile(files[i], true)) { delete rdb; return NULL; }
  return rdb;
}

void ref_frs_close(void* h) { delete static_cast<FFrExtractor*>(h); }

// all physical-time keys (sorted); returns their number
int ref_frs_keys(void* h, const char* const* files, int n, double* keys, int cap)
{
  FFrExtractor* rdb = static_cast<FFrExtractor*>(h);
  std::set<double> k;
  std::set<std::string> names;
  for (int i = 0; i < n; ++i) names.insert(files[i]);
  rdb->getValidKeys(k, names);
  int i = 0;
  for (double t : k) { if (i < cap) keys[i] = t; ++i; }
  return i;
}

// ffr_findptr + the stress time loop: reads nw values of the variable at every key in [keys[0..nkeys)).
// Returns the number of steps for which exactly nw values were read, or -1 when the search fails.
int ref_frs_read(void* h, const char* path, const char* og_type, int base_id, const double* keys, int nkeys, int nw,
                 double* out)
{
  FFrExtractor* rdb = static_cast<FFrExtractor*>(h);
  FFaResultDescription entry;
  entry.baseId = base_id;
  entry.OGType = og_type;
  std::string p(path);
  size_t a = 0;
  while (a != std::string::npos) {
    size_t b = p.find('|', a);
    entry.varDescrPath.push_back(p.substr(a, b == std::string::npos ? b : b - a));
    a = b == std::string::npos ? b : b + 1;
  }
  FFrEntryBase* ptr = rdb->search(entry);
  if (!ptr) return -1;
  int ok = 0;
  for (int s = 0; s < nkeys; ++s) {
    double found = 0.0;
    if (!rdb->positionRDB(keys[s], found, true)) continue;
    if (found != keys[s]) continue;
    if (rdb->getSingleTimeStepData(ptr, out + (size_t)nw * s, nw) == nw) ++ok;
  }
  return ok;
}

// ffr_setposition / ffr_increment (FFrExtractor_F.C:254-308) without the step-number read: the two moves the
// Fortran ffr_getNextStep loop (FFrExtractorInterface.f90:134-170) is built from.  Return 0 or -1 (istep < 0).
int ref_frs_setposition(void* h, double atime, double* btime)
{
  return static_cast<FFrExtractor*>(h)->positionRDB(atime, *btime, true) ? 0 : -1;
}

int ref_frs_increment(void* h, double* btime)
{
  FFrExtractor* rdb = static_cast<FFrExtractor*>(h);
  const bool ok = rdb->incrementRDB();
  *btime = rdb->getCurrentRDBPhysTime();
  return ok ? 0 : -1;
}

}  // extern "C"
