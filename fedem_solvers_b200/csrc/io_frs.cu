// io_frs.cu -- the .frs results database on the drop-in surface (host code only, no device code).
//
// Reader = what the recovery path uses of FFrExtractor (fedem-foundation/src/FFrLib): header grammar of
// FFrResultContainer::readFileHeader / readVariables (FFrResultContainer.C:234-528), variable
// descriptions FFrVariable::fillObject (FFrVariable.C:105-150), item groups and object groups with
// references, inlined definitions and nesting (FFrItemGroup.C:66-172, FFrObjectGroup.C:68-92,
// FFrFieldEntryBase.C:50-140), the per-step binary record whose layout is the depth-first traversal of
// the DATABLOCKS section (buildAndResolveHierarchy, FFrResultContainer.C:534-585), the physical-time key
// of every record (readTimeStepInformation, :714-860) and the path search of FFrExtractor::search
// (FFrExtractor.C:346-403) as driven by ffr_findptr / ffr_getdata (FFrExtractor_F.C:113-135,254-263).
// On top of it: readSupElDisplacements (src/vpmStress/displacementModule.f90:434-524) for a window of
// steps -> the columns of Q via fsr_build_finit.
//
// Writer = the subset of src/vpmCommon/rdbModule.f90 the stress module needs: text header
// (openRDBfile :268-403, writeVarDef :489-554, writeItGDef :576-619), "DATA:" marker, then one record
// per step = int32 step number, double time (writeTimeStepDB :669-736) and the caller's payload.
#include <cctype>
#include <cmath>
#include <cstdlib>
#include <map>
#include <memory>

#include "io_tagged.cuh"

namespace fsr {

struct FrsVar {
  std::string name, unit, cls;
  bool is_int = false;
  int bits = 0, repeats = 1;
};

struct FrsNode {     // a field of an object group or item group
  int var = -1;      // >= 0: variable reference (leaf)
  std::string name;  // item-group description (name, or its number as text)
  std::vector<int> kids;
  long long bits = -1;  // size of the sub-tree in the record
};

struct FrsObj {
  std::string type, descr;
  int base_id = 0, user_id = 0;
  std::vector<int> kids;
  long long bit_off = 0;
};

struct FrsTopVar {
  int var;
  long long bit_off;
};

struct FrsFile {
  std::string path, module;
  FILE* f = nullptr;
  bool swap = false;
  long long header_size = 0, step_size = 0;
  std::vector<FrsVar> vars;
  std::map<int, int> var_by_id, ig_by_id;
  std::vector<FrsNode> nodes;
  std::vector<FrsObj> objs;
  std::map<int, int> obj_by_base;
  std::vector<FrsTopVar> top;
  std::vector<double> times;
  std::vector<int> stepno;
  std::map<double, int> time_index;
  ~FrsFile() { if (f) fclose(f); }
};

// ---------------------------------------------------------------------------------------------
// header grammar

struct FrsParser {
  const std::string& s;
  size_t p;
  FrsFile& F;
  std::string err;
  FrsParser(const std::string& text, size_t pos, FrsFile& file) : s(text), p(pos), F(file) {}

  static char closer(char c) { return c == '<' ? '>' : c == '[' ? ']' : c == '{' ? '}' : ')'; }

  // s[a] is an opening bracket; returns the index of its matching closer (quotes respected)
  static size_t match(const std::string& t, size_t a)
  {
    std::vector<char> st;
    bool q = false;
    for (size_t i = a; i < t.size(); ++i) {
      const char c = t[i];
      if (c == '"') q = !q;
      if (q) continue;
      if (c == '<' || c == '[' || c == '{' || c == '(') st.push_back(closer(c));
      else if (!st.empty() && c == st.back()) {
        st.pop_back();
        if (st.empty()) return i;
      }
    }
    return std::string::npos;
  }

  // splits t at ';' outside quotes and brackets (FFaTokenizer semantics), trimming blanks and quotes
  static std::vector<std::string> split(const std::string& t)
  {
    std::vector<std::string> out;
    std::string cur;
    int depth = 0;
    bool q = false;
    for (char c : t) {
      if (c == '"') q = !q;
      if (!q) {
        if (c == '<' || c == '[' || c == '{' || c == '(') ++depth;
        else if (c == '>' || c == ']' || c == '}' || c == ')') --depth;
        else if (c == ';' && depth == 0) { out.push_back(cur); cur.clear(); continue; }
      }
      cur += c;
    }
    out.push_back(cur);
    for (std::string& x : out) {
      size_t a = 0, b = x.size();
      while (a < b && isspace((unsigned char)x[a])) ++a;
      while (b > a && isspace((unsigned char)x[b - 1])) --b;
      x = x.substr(a, b - a);
      if (x.size() >= 2 && x.front() == '"' && x.back() == '"') x = x.substr(1, x.size() - 2);
    }
    return out;
  }

  int make_var(const std::vector<std::string>& tk)  // FFrVariable::fillObject
  {
    if (tk.size() < 6) { err = "fewer than 6 fields in a variable description"; return -1; }
    FrsVar v;
    v.name = tk[1];
    v.unit = tk[2];
    v.is_int = tk[3] == "INT";
    v.bits = atoi(tk[4].c_str());
    v.cls = tk[5];
    v.repeats = 1;
    if (tk.size() > 6) {
      const std::string& d = tk[6];
      for (size_t i = 0; i < d.size();) {
        if (isdigit((unsigned char)d[i])) {
          char* e;
          v.repeats *= (int)strtol(d.c_str() + i, &e, 10);
          i = (size_t)(e - d.c_str());
        } else
          ++i;
      }
    }
    F.vars.push_back(v);
    return (int)F.vars.size() - 1;
  }

  // children of an item/object group: a run of <...> and [...] entries (FFrFieldEntryBase::resolve)
  bool parse_kids(const std::string& t, std::vector<int>& kids)
  {
    for (size_t i = 0; i < t.size();) {
      const char c = t[i];
      if (c != '<' && c != '[') { ++i; continue; }
      const size_t e = match(t, i);
      if (e == std::string::npos) { err = "unbalanced brackets in a data field list"; return false; }
      const std::vector<std::string> tk = split(t.substr(i + 1, e - i - 1));
      if (c == '<') {
        int var;
        if (tk.size() == 1) {
          auto it = F.var_by_id.find(atoi(tk[0].c_str()));
          if (it == F.var_by_id.end()) { err = "undefined variable reference " + tk[0]; return false; }
          var = it->second;
        } else if ((var = make_var(tk)) < 0)
          return false;
        FrsNode n;
        n.var = var;
        n.name = F.vars[(size_t)var].name;
        F.nodes.push_back(n);
        kids.push_back((int)F.nodes.size() - 1);
      } else {
        if (tk.size() == 1) {
          auto it = F.ig_by_id.find(atoi(tk[0].c_str()));
          if (it == F.ig_by_id.end()) { err = "undefined item group reference " + tk[0]; return false; }
          kids.push_back(it->second);
        } else {
          const int ig = make_group(tk);
          if (ig < 0) return false;
          kids.push_back(ig);
        }
      }
      i = e + 1;
    }
    return true;
  }

  int make_group(const std::vector<std::string>& tk)  // FFrItemGroup::fillObject
  {
    if (tk.size() < 3) { err = "fewer than 3 fields in an item group description"; return -1; }
    FrsNode n;
    n.name = tk[1];
    std::vector<int> kids;
    std::string rest = tk[2];
    for (size_t k = 3; k < tk.size(); ++k) rest += tk[k];
    if (!parse_kids(rest, kids)) return -1;
    n.kids.swap(kids);
    F.nodes.push_back(n);
    return (int)F.nodes.size() - 1;
  }

  long long size_of(int node)
  {
    FrsNode& n = F.nodes[(size_t)node];
    if (n.bits >= 0) return n.bits;
    long long b = 0;
    if (n.var >= 0)
      b = (long long)F.vars[(size_t)n.var].bits * F.vars[(size_t)n.var].repeats;
    else
      for (int k : n.kids) b += size_of(k);
    return F.nodes[(size_t)node].bits = b;
  }

  // VARIABLES: or DATABLOCKS: section (FFrResultContainer::readVariables)
  bool section(bool datablocks, long long& bitpos)
  {
    while (p < s.size()) {
      while (p < s.size() && isspace((unsigned char)s[p])) ++p;
      if (p >= s.size()) break;
      const char c = s[p];
      if (c == '#') { while (p < s.size() && s[p] != '\n') ++p; continue; }
      if (c != '<' && c != '[' && c != '{') break;  // next label
      const size_t e = match(s, p);
      if (e == std::string::npos) { err = "unbalanced brackets in the file header"; return false; }
      const std::vector<std::string> tk = split(s.substr(p + 1, e - p - 1));
      p = e + 1;
      if (c == '<') {
        if (tk.size() == 1 && datablocks) {
          auto it = F.var_by_id.find(atoi(tk[0].c_str()));
          if (it == F.var_by_id.end()) { err = "undefined variable " + tk[0]; return false; }
          F.top.push_back({it->second, bitpos});
          bitpos += (long long)F.vars[(size_t)it->second].bits * F.vars[(size_t)it->second].repeats;
        } else {
          const int v = make_var(tk);
          if (v < 0) return false;
          const int id = atoi(tk[0].c_str());
          if (id > 0) F.var_by_id[id] = v;
          if (datablocks && id > 0) {
            F.top.push_back({v, bitpos});
            bitpos += (long long)F.vars[(size_t)v].bits * F.vars[(size_t)v].repeats;
          }
        }
      } else if (c == '[') {
        int ig;
        if (tk.size() == 1 && datablocks) {
          auto it = F.ig_by_id.find(atoi(tk[0].c_str()));
          if (it == F.ig_by_id.end()) { err = "undefined item group " + tk[0]; return false; }
          ig = it->second;
        } else {
          if ((ig = make_group(tk)) < 0) return false;
          const int id = atoi(tk[0].c_str());
          if (id > 0) F.ig_by_id[id] = ig;
          else if (!datablocks) { err = "item group with no ID in the variable section"; return false; }
          else continue;  // inlined, not a top-level entry
          if (!datablocks) continue;
        }
        // a top-level item group: owner-less fields, kept as an object with base id 0
        FrsObj o;
        o.type = F.nodes[(size_t)ig].name;
        o.kids.push_back(ig);
        o.bit_off = bitpos;
        bitpos += size_of(ig);
        F.objs.push_back(o);
      } else {
        if (tk.size() < 5) { err = "fewer than 5 fields in an object group description"; return false; }
        FrsObj o;
        o.type = tk[0];
        o.base_id = atoi(tk[1].c_str());
        o.user_id = atoi(tk[2].c_str());
        o.descr = tk[3];
        std::string rest = tk[4];
        for (size_t k = 5; k < tk.size(); ++k) rest += tk[k];
        if (!parse_kids(rest, o.kids)) return false;
        o.bit_off = bitpos;
        for (int k : o.kids) bitpos += size_of(k);
        F.objs.push_back(o);
        if (o.base_id > 0 && !F.obj_by_base.count(o.base_id)) F.obj_by_base[o.base_id] = (int)F.objs.size() - 1;
      }
    }
    return true;
  }
};

static int frs_open_file(const char* path, FrsFile& F)
{
  F.path = path;
  TaggedFile tf;
  int rc = tf.open_read(path, true);
  if (rc) return rc;
  // FFrResultContainer accepts every tag (FFaTag::read); the solvers write "#FEDEM response data",
  // fedem_modes "#FEDEM modal data".  Matrix and SAM files are refused here.
  if (tf.tag.compare(0, 6, "#FEDEM") != 0 || tf.tag.find(" data") == std::string::npos) { set_error("%s is not a results database file, tag=%s", path, tf.tag.c_str()); return FSR_ERR_ARG; }
  F.swap = tf.swap;
  // read the text header up to the "DATA:" label; the binary records start right after the colon
  // (copyHeaderToBinaryFile, rdbModule.f90:418-463, writes 'DATA:' without a line end)
  std::string text;
  bool found = false, have_blocks = false;
  long long hdr_end = 0;
  {
    const long long start = (long long)ftello(tf.f);
    std::string line;
    int c;
    while (!found && (c = fgetc(tf.f)) != EOF) {
      line += (char)c;
      if (line.size() == 5 && line == "DATA:") { found = true; break; }
      if (c == '\n') {
        if (!line.compare(0, 11, "DATABLOCKS:")) have_blocks = true;
        text += line;
        line.clear();
      }
    }
    hdr_end = found ? (long long)ftello(tf.f) : start;
  }
  if (!found) { set_error("%s: could not find the DATA: field (incomplete header)", path); return FSR_ERR_ARG; }
  if (!have_blocks) { set_error("%s: could not find the DATABLOCKS: field", path); return FSR_ERR_ARG; }
  F.header_size = hdr_end;
  // heading lines "Label = value;" until VARIABLES:
  const size_t pv = text.find("VARIABLES:"), pd = text.find("DATABLOCKS:");
  if (pd == std::string::npos) { set_error("%s: malformed header", path); return FSR_ERR_ARG; }
  {
    const std::string head = text.substr(0, pv == std::string::npos ? pd : pv);
    size_t m = head.find("Module ");
    if (m != std::string::npos) {
      size_t eq = head.find('=', m), sc = head.find(';', m);
      if (eq != std::string::npos && sc != std::string::npos && sc > eq) {
        F.module = head.substr(eq + 1, sc - eq - 1);
        while (!F.module.empty() && F.module.front() == ' ') F.module.erase(0, 1);
      }
    }
  }
  FrsParser P(text, 0, F);
  long long bitpos = 0;
  if (pv != std::string::npos && pv < pd) {
    P.p = pv + 10;
    if (!P.section(false, bitpos)) { set_error("%s: %s", path, P.err.c_str()); return FSR_ERR_ARG; }
  }
  P.p = pd + 11;
  if (!P.section(true, bitpos)) { set_error("%s: %s", path, P.err.c_str()); return FSR_ERR_ARG; }
  F.step_size = bitpos >> 3;
  if (F.step_size < 1) { set_error("%s: empty time step record", path); return FSR_ERR_ARG; }
  // physical time and step number of every record
  long long t_off = -1, n_off = -1;
  for (const FrsTopVar& t : F.top) {
    const FrsVar& v = F.vars[(size_t)t.var];
    if (v.name == "Physical time" && t_off < 0) t_off = t.bit_off >> 3;
    if (v.name == "Time step number" && n_off < 0) n_off = t.bit_off >> 3;
  }
  if (t_off < 0) { set_error("%s: no time step data found", path); return FSR_ERR_ARG; }
  fseeko(tf.f, 0, SEEK_END);
  const long long fsize = (long long)ftello(tf.f);
  const long long nst = (fsize - F.header_size) / F.step_size;
  F.times.resize((size_t)nst);
  F.stepno.assign((size_t)nst, 0);
  for (long long i = 0; i < nst; ++i) {
    double t;
    fseeko(tf.f, (off_t)(F.header_size + i * F.step_size + t_off), SEEK_SET);
    if (fread(&t, 8, 1, tf.f) != 1) { set_error("%s: error reading the physical time of step %lld", path, i); return FSR_ERR_ARG; }
    if (F.swap) swap_bytes(&t, 8, 1);
    F.times[(size_t)i] = t;
    F.time_index[t] = (int)i;
    if (n_off >= 0) {
      int n;
      fseeko(tf.f, (off_t)(F.header_size + i * F.step_size + n_off), SEEK_SET);
      if (fread(&n, 4, 1, tf.f) == 1) {
        if (F.swap) swap_bytes(&n, 4, 1);
        F.stepno[(size_t)i] = n;
      }
    }
  }
  F.f = tf.f;
  tf.f = nullptr;
  return FSR_OK;
}

struct FrsLoc {
  int file, var;
  long long byte_off;
  int count = 0;   // > 0: an item group read as one array (ffr_findptr on a group path, e.g. "Eigenvectors|Mode  1"): values of all
                   // its variables, which share one number format; 0: the variable's own size
};

}  // namespace fsr

using namespace fsr;

struct fsr_frs {
  std::vector<std::unique_ptr<FrsFile>> files;
  std::vector<double> times;  // sorted union of the time keys of all files
  std::vector<int> stepno;
  std::vector<std::vector<FrsLoc>> handles;
};

namespace fsr {

// leaf variables below an item group, in record order; false when they do not share one number format
static bool frs_group_span(const FrsFile& F, const FrsNode& n, int& first_var, int& count)
{
  if (n.var >= 0) {
    const FrsVar& v = F.vars[(size_t)n.var];
    if (first_var < 0) first_var = n.var;
    else if (F.vars[(size_t)first_var].bits != v.bits || F.vars[(size_t)first_var].is_int != v.is_int) return false;
    count += v.repeats;
    return true;
  }
  for (int k : n.kids)
    if (!frs_group_span(F, F.nodes[(size_t)k], first_var, count)) return false;
  return true;
}

static bool frs_find_in_file(const FrsFile& F, const std::vector<std::string>& path, const std::string& og_type, int base_id,
                             int& var, long long& bit_off, int& count)
{
  count = 0;
  if (og_type.empty()) {  // top-level variable
    if (path.size() != 1) return false;
    for (const FrsTopVar& t : F.top)
      if (F.vars[(size_t)t.var].name == path[0]) { var = t.var; bit_off = t.bit_off; return true; }
    return false;
  }
  auto it = F.obj_by_base.find(base_id);
  if (it == F.obj_by_base.end()) return false;
  const FrsObj& o = F.objs[(size_t)it->second];
  const std::vector<int>* kids = &o.kids;
  long long off = o.bit_off;
  for (size_t lev = 0; lev < path.size(); ++lev) {
    int hit = -1;
    long long o2 = off;
    for (int k : *kids) {
      const FrsNode& n = F.nodes[(size_t)k];
      if (n.name == path[lev]) { hit = k; break; }
      o2 += n.bits;
    }
    if (hit < 0) return false;
    off = o2;
    const FrsNode& n = F.nodes[(size_t)hit];
    if (lev + 1 == path.size()) {
      bit_off = off;
      if (n.var < 0) {   // the path ends on an item group: all its values as one array (FFrExtractor reads a group that way)
        int first = -1, cnt = 0;
        if (!frs_group_span(F, n, first, cnt) || first < 0) return false;
        var = first;
        count = cnt;
        return true;
      }
      var = n.var;
      return true;
    }
    if (n.var >= 0) return false;
    kids = &n.kids;
  }
  return false;
}

static int frs_read_values(fsr_frs* db, int handle, int gstep, double* out, int nw)
{
  const double key = db->times[(size_t)gstep];
  for (const FrsLoc& L : db->handles[(size_t)handle]) {
    FrsFile& F = *db->files[(size_t)L.file];
    auto it = F.time_index.find(key);
    if (it == F.time_index.end()) continue;
    const FrsVar& v = F.vars[(size_t)L.var];
    const int n = std::min(nw, L.count > 0 ? L.count : v.repeats), nb = v.bits / 8;
    unsigned char buf[8];
    fseeko(F.f, (off_t)(F.header_size + (long long)it->second * F.step_size + L.byte_off), SEEK_SET);
    for (int i = 0; i < n; ++i) {
      if (nb < 1 || nb > 8 || fread(buf, 1, (size_t)nb, F.f) != (size_t)nb) return i;
      if (F.swap && nb > 1) swap_bytes(buf, (size_t)nb, 1);
      double x;
      if (v.is_int) {
        if (nb == 1) x = (double)*reinterpret_cast<signed char*>(buf);
        else if (nb == 2) { int16_t t; memcpy(&t, buf, 2); x = t; }
        else if (nb == 4) { int32_t t; memcpy(&t, buf, 4); x = t; }
        else { int64_t t; memcpy(&t, buf, 8); x = (double)t; }
      } else if (nb == 4) { float t; memcpy(&t, buf, 4); x = t; }
      else if (nb == 8) memcpy(&x, buf, 8);
      else return i;
      out[i] = x;
    }
    return n;
  }
  return 0;
}

}  // namespace fsr

extern "C" {

int fsr_frs_open(fsr_frs** db, const char* const* paths, int nfiles)
{
  if (!db || !paths || nfiles < 1) { set_error("fsr_frs_open: bad arguments"); return FSR_ERR_ARG; }
  std::unique_ptr<fsr_frs> d(new fsr_frs);
  for (int i = 0; i < nfiles; ++i) {
    std::unique_ptr<FrsFile> F(new FrsFile);
    const int rc = frs_open_file(paths[i], *F);
    if (rc) return rc;
    d->files.push_back(std::move(F));
  }
  std::map<double, int> keys;
  for (auto& F : d->files)
    for (size_t i = 0; i < F->times.size(); ++i) keys.emplace(F->times[i], F->stepno[i]);
  for (auto& kv : keys) { d->times.push_back(kv.first); d->stepno.push_back(kv.second); }
  *db = d.release();
  return FSR_OK;
}

void fsr_frs_close(fsr_frs* db) { delete db; }

int fsr_frs_num_steps(const fsr_frs* db) { return db ? (int)db->times.size() : FSR_ERR_ARG; }

int fsr_frs_get_steps(const fsr_frs* db, int* stepno, double* time, int cap)
{
  if (!db) { set_error("fsr_frs_get_steps: null handle"); return FSR_ERR_ARG; }
  const int n = std::min<int>(cap, (int)db->times.size());
  for (int i = 0; i < n; ++i) {
    if (stepno) stepno[i] = db->stepno[(size_t)i];
    if (time) time[i] = db->times[(size_t)i];
  }
  return (int)db->times.size();
}

int fsr_frs_find(fsr_frs* db, const char* var_path, const char* og_type, int base_id)
{
  if (!db || !var_path) { set_error("fsr_frs_find: bad arguments"); return FSR_ERR_ARG; }
  std::vector<std::string> path;
  {
    std::string p = var_path;
    size_t a = 0;
    while (true) {
      const size_t b = p.find('|', a);
      path.push_back(p.substr(a, b == std::string::npos ? b : b - a));
      if (b == std::string::npos) break;
      a = b + 1;
    }
  }
  std::vector<FrsLoc> locs;
  for (size_t i = 0; i < db->files.size(); ++i) {
    int var, count;
    long long off;
    if (frs_find_in_file(*db->files[i], path, og_type ? og_type : "", base_id, var, off, count)) locs.push_back({(int)i, var, off >> 3, count});
  }
  if (locs.empty()) return -1;  // like a null pointer from ffr_findptr: not an error by itself
  db->handles.push_back(locs);
  return (int)db->handles.size() - 1;
}

int fsr_frs_var_size(const fsr_frs* db, int handle)
{
  if (!db || handle < 0 || handle >= (int)db->handles.size()) { set_error("fsr_frs_var_size: invalid handle"); return FSR_ERR_ARG; }
  const FrsLoc& L = db->handles[(size_t)handle][0];
  return L.count > 0 ? L.count : db->files[(size_t)L.file]->vars[(size_t)L.var].repeats;
}

int fsr_frs_read(fsr_frs* db, int handle, int step0, int nsteps, double* data, int nw, int ld)
{
  if (!db || handle < 0 || handle >= (int)db->handles.size() || !data || nw < 1 || ld < nw || step0 < 0 || nsteps < 0 ||
      step0 + nsteps > (int)db->times.size()) { set_error("fsr_frs_read: bad arguments"); return FSR_ERR_ARG; }
  for (int s = 0; s < nsteps; ++s) {
    const int got = frs_read_values(db, handle, step0 + s, data + (size_t)ld * s, nw);
    if (got != nw) {  // the message of readSupElDisplacements (displacementModule.f90:600)
      set_error("Mismatch between length of wanted array %d and actual variable size %d on the results file (step %d)", nw, got, step0 + s);
      return FSR_ERR_ARG;
    }
  }
  return FSR_OK;
}

int fsr_frs_reduced_history(fsr_frs* db, int sup_base_id, int ntriads, const int* triad_base_id, const int* ndofs,
                            const int* first_dof, const double* tr_undef, int ngen, int gen_first_dof, int step0,
                            int nsteps, double* Q, int ldq)
{
  if (!db || ntriads < 0 || (ntriads > 0 && (!triad_base_id || !ndofs || !first_dof || !tr_undef)) || !Q || nsteps < 0) {
    set_error("fsr_frs_reduced_history: bad arguments");
    return FSR_ERR_ARG;
  }
  // initiateTriadAndSupElTypeModule / displacementModule.f90:262-292: result pointers of the triads and the part
  // readResponsePointers (displacementModule.f90:246-328): all found = go on; NONE found = a warning (ierr = 1) and the
  // recovery is "based on local deformations relative to the modelling configuration of the part", i.e. the triads stay
  // where the solver input file put them and finit = 0; only a PARTIAL set is an error
  std::vector<int> h((size_t)ntriads + 2, -1);
  int nfound = 0, nfixed = 0;
  for (int i = 0; i < ntriads; ++i) {
    if (ndofs[i] == 6) h[(size_t)i] = fsr_frs_find(db, "Position matrix", "Triad", triad_base_id[i]);
    else if (ndofs[i] == 3) h[(size_t)i] = fsr_frs_find(db, "Position", "Triad", triad_base_id[i]);
    else { ++nfixed; continue; }
    nfound += h[(size_t)i] >= 0;
  }
  nfound += (h[(size_t)ntriads] = fsr_frs_find(db, "Position matrix", "Part", sup_base_id)) >= 0;
  if (ngen > 0) nfound += (h[(size_t)ntriads + 1] = fsr_frs_find(db, "Generalized displacement", "Part", sup_base_id)) >= 0;
  if (nfound == 0) {
    for (int s = 0; s < nsteps; ++s)
      for (int k = 0; k < ldq; ++k) Q[(size_t)s * ldq + k] = 0.0;
    set_error("No system-level response variables found. Stress recovery will be based on local deformations relative to the "
              "modelling configuration of the part.");
    return 1;   // warning, like ierr = 1 of readResponsePointers
  }
  if (nfound + nfixed != ntriads + 1 + (ngen > 0 ? 1 : 0)) {
    for (int i = 0; i < ntriads; ++i)
      if (h[(size_t)i] < 0 && (ndofs[i] == 6 || ndofs[i] == 3)) { set_error("Cannot find position for Triad {%d} on the results file", triad_base_id[i]); return FSR_ERR_ARG; }
    if (h[(size_t)ntriads] < 0) { set_error("Cannot find position matrix for Part {%d} on the results file", sup_base_id); return FSR_ERR_ARG; }
    set_error("Cannot find generalized displacements for Part {%d} on the results file", sup_base_id);
    return FSR_ERR_ARG;
  }
  std::vector<double> sup((size_t)nsteps * 12), tri((size_t)nsteps * std::max(ntriads, 1) * 12, 0.0), gen((size_t)nsteps * std::max(ngen, 1));
  int rc;
  if ((rc = fsr_frs_read(db, h[(size_t)ntriads], step0, nsteps, sup.data(), 12, 12))) return rc;
  for (int i = 0; i < ntriads; ++i) {
    if (h[(size_t)i] < 0) continue;
    // 6-DOF triads: the whole 3x4 matrix; 3-DOF triads: the position column only (ur(:,4))
    if (ndofs[i] == 6) rc = fsr_frs_read(db, h[(size_t)i], step0, nsteps, tri.data() + 12 * (size_t)i, 12, 12 * ntriads);
    else rc = fsr_frs_read(db, h[(size_t)i], step0, nsteps, tri.data() + 12 * (size_t)i + 9, 3, 12 * ntriads);
    if (rc) return rc;
  }
  if (ngen > 0 && (rc = fsr_frs_read(db, h[(size_t)ntriads + 1], step0, nsteps, gen.data(), ngen, ngen))) return rc;
  return fsr_build_finit(nsteps, ntriads, sup.data(), tri.data(), tr_undef, ndofs, first_dof, ngen, gen.data(), gen_first_dof, Q, ldq);
}

// ---------------------------------------------------------------------------------------------
// writer

struct fsr_frs_writer {
  FILE* f = nullptr;
  long long payload_bytes = 0;
  int nsteps = 0;
  ~fsr_frs_writer() { if (f) fclose(f); }
};

int fsr_frs_create(fsr_frs_writer** w, const char* path, int checksum, const char* header_text, long long payload_bytes)
{
  return fsr_frs_create_tagged(w, path, "#FEDEM response data", checksum, header_text, payload_bytes);
}

int fsr_frs_create_tagged(fsr_frs_writer** w, const char* path, const char* tag, int checksum, const char* header_text, long long payload_bytes)
{
  if (!w || !path || !tag || !header_text || payload_bytes < 0) { set_error("fsr_frs_create: bad arguments"); return FSR_ERR_ARG; }
  TaggedFile tf;
  int rc = tf.open_write(path, tag, (unsigned int)checksum);
  if (rc) return rc;
  if (fputs(header_text, tf.f) < 0 || fputs("DATA:", tf.f) < 0) { set_error("%s: write error", path); return FSR_ERR_ARG; }
  fsr_frs_writer* x = new fsr_frs_writer;
  x->f = tf.f;
  tf.f = nullptr;
  x->payload_bytes = payload_bytes;
  *w = x;
  return FSR_OK;
}

int fsr_frs_write_step(fsr_frs_writer* w, int stepno, double time, const void* payload)
{
  if (!w || !w->f || (w->payload_bytes > 0 && !payload)) { set_error("fsr_frs_write_step: bad arguments"); return FSR_ERR_ARG; }
  if (fwrite(&stepno, 4, 1, w->f) != 1 || fwrite(&time, 8, 1, w->f) != 1 ||
      (w->payload_bytes > 0 && fwrite(payload, 1, (size_t)w->payload_bytes, w->f) != (size_t)w->payload_bytes)) {
    set_error("fsr_frs_write_step: write error");
    return FSR_ERR_ARG;
  }
  return ++w->nsteps;
}

int fsr_frs_finish(fsr_frs_writer* w)
{
  if (!w) return FSR_ERR_ARG;
  int rc = FSR_OK;
  if (w->f && fclose(w->f) != 0) { set_error("fsr_frs_finish: close error"); rc = FSR_ERR_ARG; }
  w->f = nullptr;
  delete w;
  return rc;
}

}  // extern "C"
