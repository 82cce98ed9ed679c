// io_fsi.cu -- the solver input file (.fsi) as far as fedem_stress reads it, host code only.
//
// readSolverData (src/vpmStress/displacementModule.f90:138-229) reads, from the Fortran namelist file the
// dynamics solver was run with: &HEADING (modelFile), &ENVIRONMENT (gravity), the &SUP_EL record whose id is the
// base id of the part (InitiateSupEls1, src/vpmStress/initiateTriadAndSupElTypeModule.f90:121-246: id, extId,
// extDescr, numTriads, triadIds, numGenDOFs, supPos), the numTriads &TRIAD_UNDPOS records that follow it
// (undPosInSupElSystem -> sup%TrUndeformed) and the &TRIAD records of those triads (InitiateTriads :34-103:
// id, extId, nDOFs, ur); InitiateSupEls2 (:262-319) then numbers the reduced DOFs: triads in the order of
// triadIds, nDOFs each, followed by the generalized DOFs.  That is everything BuildFinit needs.
//
// The namelist reader below handles what the solver input files contain: `&GROUP ... /` records, `key = v v v`
// with blanks, commas or line breaks between values, several assignments on one line, quoted strings (either
// quote, doubled quotes inside), r*v repeat counts, d/D exponents and `!` comments.
#include <algorithm>
#include <cctype>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "common.cuh"

using namespace fsr;

namespace {

struct Record {
  std::string group;                                        // upper case, without the '&'
  std::map<std::string, std::vector<std::string>> values;   // lower-case key -> raw value tokens
  const std::vector<std::string>* get(const char* k) const
  {
    auto it = values.find(k);
    return it == values.end() ? nullptr : &it->second;
  }
  bool ints(const char* k, std::vector<int>& out) const
  {
    out.clear();
    const std::vector<std::string>* v = get(k);
    if (!v) return false;
    for (const std::string& s : *v) { char* e; long l = strtol(s.c_str(), &e, 10); if (*e) return false; out.push_back((int)l); }
    return true;
  }
  bool reals(const char* k, std::vector<double>& out) const
  {
    out.clear();
    const std::vector<std::string>* v = get(k);
    if (!v) return false;
    for (std::string s : *v) {
      for (char& c : s) if (c == 'd' || c == 'D') c = 'e';
      char* e; double d = strtod(s.c_str(), &e); if (*e) return false; out.push_back(d);
    }
    return true;
  }
  int int1(const char* k, int dflt) const { std::vector<int> v; return ints(k, v) && !v.empty() ? v[0] : dflt; }
  std::string str(const char* k) const { const std::vector<std::string>* v = get(k); return v && !v->empty() ? (*v)[0] : std::string(); }
};

static bool read_namelists(const char* path, std::vector<Record>& recs)
{
  FILE* f = fopen(path, "r");
  if (!f) { set_error("Unable to open solver input file %s", path); return false; }
  std::string text;
  char buf[1 << 14];
  size_t n;
  while ((n = fread(buf, 1, sizeof(buf), f)) > 0) text.append(buf, n);
  fclose(f);
  size_t i = 0;
  const size_t N = text.size();
  auto skip_blank = [&]() {
    while (i < N) {
      if (text[i] == '!') { while (i < N && text[i] != '\n') ++i; }
      else if (isspace((unsigned char)text[i]) || text[i] == ',') ++i;
      else break;
    }
  };
  while (i < N) {
    while (i < N && text[i] != '&') { if (text[i] == '!') while (i < N && text[i] != '\n') ++i; else ++i; }
    if (i >= N) break;
    ++i;
    Record r;
    while (i < N && (isalnum((unsigned char)text[i]) || text[i] == '_')) r.group += (char)toupper((unsigned char)text[i++]);
    if (r.group == "END") continue;
    std::vector<std::string>* cur = nullptr;
    for (;;) {
      skip_blank();
      if (i >= N) { set_error("%s: namelist &%s is not terminated", path, r.group.c_str()); return false; }
      if (text[i] == '/') { ++i; break; }
      if (text[i] == '&') {   // "&END" terminator of old-style namelists
        if (N - i >= 4 && strncasecmp(text.c_str() + i, "&end", 4) == 0) { i += 4; break; }
        set_error("%s: namelist &%s is not terminated", path, r.group.c_str());
        return false;
      }
      if (text[i] == '\'' || text[i] == '"') {
        const char q = text[i++];
        std::string s;
        while (i < N) {
          if (text[i] == q) { if (i + 1 < N && text[i + 1] == q) { s += q; i += 2; continue; } ++i; break; }
          s += text[i++];
        }
        if (cur) cur->push_back(s);
        continue;
      }
      // a bare token: either "name" followed by '=' (possibly with an index) or a value
      size_t j = i;
      while (j < N && !isspace((unsigned char)text[j]) && text[j] != ',' && text[j] != '=' && text[j] != '/' && text[j] != '!' &&
             text[j] != '\'' && text[j] != '"') ++j;
      std::string tok = text.substr(i, j - i);
      size_t k = j;
      while (k < N && (text[k] == ' ' || text[k] == '\t')) ++k;
      if (k < N && text[k] == '=' && !tok.empty() && (isalpha((unsigned char)tok[0]) || tok[0] == '_')) {
        for (char& c : tok) c = (char)tolower((unsigned char)c);
        const size_t par = tok.find('(');
        if (par != std::string::npos) tok.resize(par);   // a(2) = ... appends to a: good enough for this file type
        cur = &r.values[tok];
        i = k + 1;
        continue;
      }
      if (tok.empty()) { ++i; continue; }
      i = j;
      if (!cur) continue;
      const size_t star = tok.find('*');
      if (star != std::string::npos && star > 0 && std::all_of(tok.begin(), tok.begin() + star, [](char c) { return isdigit((unsigned char)c); })) {
        const int rep = atoi(tok.substr(0, star).c_str());
        for (int q = 0; q < rep; ++q) cur->push_back(tok.substr(star + 1));
      } else
        cur->push_back(tok);
    }
    recs.push_back(r);
  }
  return true;
}

}  // namespace

struct fsr_fsi {
  int base_id = 0, user_id = 0, ngen = 0;
  std::string descr, model_file;
  double sup_pos[12] = {1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0};   // column-major 3x4
  double gravity[3] = {0, 0, 0};
  std::vector<int> triad_base, triad_user, ndofs, first_dof;
  std::vector<double> tr_undef, triad_ur;                       // [ntriads][12] column-major
};

// the file stores 3x4 matrices row by row (the Fortran code transposes after reading a (4,3) array)
static bool rowwise_to_colmajor(const std::vector<double>& v, double* out)
{
  if (v.size() != 12) return false;
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 4; ++j) out[i + 3 * j] = v[4 * i + j];
  return true;
}

extern "C" {

int fsr_fsi_open(fsr_fsi** out, const char* path, int part_base_id)
{
  if (!out || !path) { set_error("fsr_fsi_open: bad arguments"); return FSR_ERR_ARG; }
  *out = nullptr;
  std::vector<Record> recs;
  if (!read_namelists(path, recs)) return FSR_ERR_ARG;
  fsr_fsi* h = new fsr_fsi;
  std::vector<double> r;
  std::vector<int> ids;
  size_t isup = recs.size();
  for (size_t k = 0; k < recs.size(); ++k) {
    const Record& rec = recs[k];
    if (rec.group == "HEADING") h->model_file = rec.str("modelfile");
    else if (rec.group == "ENVIRONMENT") { if (rec.reals("gravity", r) && r.size() == 3) for (int i = 0; i < 3; ++i) h->gravity[i] = r[i]; }
    else if (rec.group == "SUP_EL" && isup == recs.size() && rec.int1("id", -1) == part_base_id) isup = k;
  }
  if (isup == recs.size()) { set_error("baseID %d was not found on the solver input file %s", part_base_id, path); delete h; return FSR_ERR_ARG; }
  const Record& sup = recs[isup];
  h->base_id = part_base_id;
  h->user_id = sup.int1("extid", 0);
  h->descr = sup.str("extdescr");
  h->ngen = sup.int1("numgendofs", 0);
  const int nt = sup.int1("numtriads", 0);
  if (!sup.ints("triadids", ids) || (int)ids.size() < nt) { set_error("%s: &SUP_EL %d: triadIds has fewer than numTriads = %d entries", path, part_base_id, nt); delete h; return FSR_ERR_ARG; }
  ids.resize((size_t)nt);
  if (sup.reals("suppos", r) && !rowwise_to_colmajor(r, h->sup_pos)) { set_error("%s: &SUP_EL %d: supPos needs 12 values", path, part_base_id); delete h; return FSR_ERR_ARG; }
  h->triad_base = ids;
  h->triad_user.assign((size_t)nt, 0);
  h->ndofs.assign((size_t)nt, -1);
  h->first_dof.assign((size_t)nt, 0);
  h->tr_undef.assign((size_t)nt * 12, 0.0);
  h->triad_ur.assign((size_t)nt * 12, 0.0);
  // the numTriads &TRIAD_UNDPOS records following the &SUP_EL record
  int found = 0;
  for (size_t k = isup + 1; k < recs.size() && found < nt; ++k) {
    const Record& rec = recs[k];
    if (rec.group != "TRIAD_UNDPOS") continue;
    ++found;
    const int sid = rec.int1("supelid", -1), tid = rec.int1("triadid", -1);
    if (sid != part_base_id) { set_error("%s: &TRIAD_UNDPOS of Triad %d belongs to Part %d, expected Part %d", path, tid, sid, part_base_id); delete h; return FSR_ERR_ARG; }
    const auto it = std::find(ids.begin(), ids.end(), tid);
    if (it == ids.end()) { set_error("%s: Triad %d is not connected to Part %d", path, tid, part_base_id); delete h; return FSR_ERR_ARG; }
    if (!rec.reals("undposinsupelsystem", r) || !rowwise_to_colmajor(r, &h->tr_undef[12 * (size_t)(it - ids.begin())])) {
      set_error("%s: &TRIAD_UNDPOS of Triad %d: undPosInSupElSystem needs 12 values", path, tid); delete h; return FSR_ERR_ARG;
    }
  }
  if (found < nt) { set_error("%s: Part %d has %d triads but only %d &TRIAD_UNDPOS records follow it", path, part_base_id, nt, found); delete h; return FSR_ERR_ARG; }
  // the &TRIAD records of those triads
  for (const Record& rec : recs) {
    if (rec.group != "TRIAD") continue;
    const auto it = std::find(ids.begin(), ids.end(), rec.int1("id", -1));
    if (it == ids.end()) continue;
    const size_t j = (size_t)(it - ids.begin());
    h->triad_user[j] = rec.int1("extid", 0);
    h->ndofs[j] = rec.int1("ndofs", 0);
    if (rec.reals("ur", r)) rowwise_to_colmajor(r, &h->triad_ur[12 * j]);
  }
  int ndof = 0;   // InitiateSupEls2
  for (int j = 0; j < nt; ++j) {
    if (h->ndofs[(size_t)j] < 0) { set_error("%s: Triad %d of Part %d has no &TRIAD record", path, ids[(size_t)j], part_base_id); delete h; return FSR_ERR_ARG; }
    h->first_dof[(size_t)j] = ndof + 1;
    ndof += h->ndofs[(size_t)j];
  }
  *out = h;
  return FSR_OK;
}

void fsr_fsi_close(fsr_fsi* h) { delete h; }

// sup%id, numTriads, numGenDOFs, sup%supTr (= supTrInit, column-major 3x4), gravity, modelFile.  Any may be NULL.
int fsr_fsi_part(const fsr_fsi* h, int* user_id, char* descr, int dcap, int* ntriads, int* ngen, double* sup_pos,
                 double* gravity, char* model_file, int mcap)
{
  if (!h) { set_error("fsr_fsi_part: bad arguments"); return FSR_ERR_ARG; }
  if (user_id) *user_id = h->user_id;
  if (descr && dcap > 0) { strncpy(descr, h->descr.c_str(), (size_t)dcap - 1); descr[dcap - 1] = 0; }
  if (ntriads) *ntriads = (int)h->triad_base.size();
  if (ngen) *ngen = h->ngen;
  if (sup_pos) memcpy(sup_pos, h->sup_pos, sizeof(h->sup_pos));
  if (gravity) memcpy(gravity, h->gravity, sizeof(h->gravity));
  if (model_file && mcap > 0) { strncpy(model_file, h->model_file.c_str(), (size_t)mcap - 1); model_file[mcap - 1] = 0; }
  return h->base_id;
}

// per triad of the part, in the order of triadIds: base id, user id, nDOFs, first reduced DOF (1-based),
// TrUndeformed and the initial position ur (column-major 3x4 each).  Returns the position of the first
// generalized DOF (sup%genDOFs%firstDOF).
int fsr_fsi_triads(const fsr_fsi* h, int* base_id, int* user_id, int* ndofs, int* first_dof, double* tr_undef, double* ur)
{
  if (!h) { set_error("fsr_fsi_triads: bad arguments"); return FSR_ERR_ARG; }
  const size_t nt = h->triad_base.size();
  int ndof = 0;
  for (size_t j = 0; j < nt; ++j) {
    if (base_id) base_id[j] = h->triad_base[j];
    if (user_id) user_id[j] = h->triad_user[j];
    if (ndofs) ndofs[j] = h->ndofs[j];
    if (first_dof) first_dof[j] = h->first_dof[j];
    ndof += h->ndofs[j];
  }
  if (tr_undef) memcpy(tr_undef, h->tr_undef.data(), sizeof(double) * 12 * nt);
  if (ur) memcpy(ur, h->triad_ur.data(), sizeof(double) * 12 * nt);
  return ndof + 1;
}

// ReadStrainGages (src/vpmStress/strainGageModule.f90:78-237) with thisLinkNumber present: every &STRAIN_ROSETTE record
// of the file (id, extId, extDescr, linkId, type, zeroInit, numnod, nodes, rPos, zPos, Emod, nu, gateVal, snCurve);
// a record whose linkId is not `link_base_id` is an input error there ("The part base ID does not match"), and so is
// an unknown rosette type.  ros[k].nodes come back as the EXTERNAL node numbers of the file: the caller maps them with
// ffl_ext2int and swaps them when the normal points the wrong way (checkRosette :479-521).  rPos(4,3) is filled in
// Fortran order, i.e. the file lists posInGl row by row; rpos is posInGl(3,4) column-major.
// user_id [cap], descr [cap][descr_stride] may be NULL.  Returns the number of records (ros may be NULL to count).
int fsr_fsi_read_rosettes(const char* path, int link_base_id, fsr_rosette* ros, int* user_id, char* descr, int descr_stride, int cap)
{
  if (!path) { set_error("fsr_fsi_read_rosettes: bad arguments"); return FSR_ERR_ARG; }
  std::vector<Record> recs;
  if (!read_namelists(path, recs)) return FSR_ERR_ARG;
  static const struct { const char* name; int ngage; double alpha; } kTypes[] = {
      {"SINGLE_GAGE", 1, 0.0}, {"DOUBLE_GAGE_90", 2, 1.5707963267948966}, {"TRIPLE_GAGE_60", 3, 1.0471975511965976},
      {"TRIPLE_GAGE_45", 3, 0.7853981633974483}};
  int n = 0;
  std::vector<int> iv;
  std::vector<double> dv;
  for (const Record& r : recs) {
    if (r.group != "STRAIN_ROSETTE") continue;
    if (ros && n < cap) {
      fsr_rosette& o = ros[n];
      memset(&o, 0, sizeof(o));
      o.id = r.int1("id", 0);
      const int link = r.int1("linkid", 0);
      if (link != link_base_id) {
        set_error("%s: STRAIN_ROSETTE %d (record %d): the part base ID %d does not match %d", path, o.id, n + 1, link, link_base_id);
        return FSR_ERR_ARG;
      }
      std::string type = r.str("type");
      while (!type.empty() && type.back() == ' ') type.pop_back();
      int it = -1;
      for (int k = 0; k < 4; ++k) if (strcasecmp(type.c_str(), kTypes[k].name) == 0) it = k;
      if (it < 0) { set_error("%s: invalid rosette-type %s for Rosette %d", path, type.c_str(), o.id); return FSR_ERR_ARG; }
      o.ngage = kTypes[it].ngage; o.alpha_gages = kTypes[it].alpha;
      o.zero_init = r.int1("zeroinit", 0) > 0 ? 1 : 0;
      o.numnod = r.int1("numnod", 0);
      if (o.numnod < 3 || o.numnod > 4) { set_error("%s: Rosette %d has %d nodes (3 or 4 expected)", path, o.id, o.numnod); return FSR_ERR_ARG; }
      if (!r.ints("nodes", iv) || (int)iv.size() < o.numnod) { set_error("%s: Rosette %d: nodes missing", path, o.id); return FSR_ERR_ARG; }
      for (int k = 0; k < o.numnod; ++k) o.nodes[k] = iv[(size_t)k];
      if (!r.reals("rpos", dv) || dv.size() < 12) { set_error("%s: Rosette %d: rPos needs 12 values", path, o.id); return FSR_ERR_ARG; }
      for (int i = 0; i < 3; ++i) for (int j = 0; j < 4; ++j) o.rpos[i + 3 * j] = dv[(size_t)(4 * i + j)];
      o.zpos = r.reals("zpos", dv) && !dv.empty() ? dv[0] : 0.0;
      o.emod = r.reals("emod", dv) && !dv.empty() ? dv[0] : 0.0;
      o.nu = r.reals("nu", dv) && !dv.empty() ? dv[0] : 0.0;
      o.gate = r.reals("gateval", dv) && !dv.empty() ? dv[0] : 0.0;
      if (r.reals("sncurve", dv)) for (size_t k = 0; k < 4 && k < dv.size(); ++k) o.sncurve[k] = dv[k];
      if (user_id) { user_id[n] = r.ints("extid", iv) && !iv.empty() ? iv[0] : 0; }
      if (descr && descr_stride > 0) {
        const std::string d = r.str("extdescr");
        strncpy(descr + (size_t)n * descr_stride, d.c_str(), (size_t)descr_stride - 1);
        descr[(size_t)n * descr_stride + descr_stride - 1] = 0;
      }
    }
    ++n;
  }
  return n;
}

}  // extern "C"
