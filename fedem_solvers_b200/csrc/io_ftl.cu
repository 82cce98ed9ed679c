// io_ftl.cu -- the FE part (.ftl) side of the drop-in surface, host code only.
//
// fedem_stress gets everything it knows about an element through the Fortran accessors of the FE-model
// singleton (fedem-foundation/src/FFlLib/FFlLinkHandler_F.C: ffl_getsize :367, ffl_getnodes :459,
// ffl_gettopol :558, ffl_getelmid :699, ffl_getcoor :745, ffl_getmat :851, ffl_getthick :1057,
// ffl_getpinflags :1086, ffl_getbeamsection :1131), once per element per time step.  This file reads the
// same .ftl text file (grammar: FFlLib/FFlIOAdaptors/FFlFedemReader.C:456-763) and delivers the same
// numbers once, as flat arrays in SAM order, ready for fsr_part_create:
//   * nodes sorted by id, DOF-less (loose) nodes dropped, 3 or 6 DOFs per node from the element types
//     that use the node (FFlFENodeRefs.C:198-228; FFlNode::pushDOFs), status codes 2/1/0 for
//     external/free/fixed DOFs, extra nodes for pinned beam ends (ffl_getnodes :510-556);
//   * finite elements sorted by id, strain-coat elements skipped, dangling RGD/WAVGM/CMASS ignored
//     (FFlLinkHandler.C:804-880), the SAM element type codes of ffl_gettopol;
//   * E, nu, rho, thickness, beam coordinates with orientation point and eccentricities, the 14 beam
//     section values with the inverted shear factors, pin flags, and the external element id negated
//     for elements outside the -group selection (FFlUtils.C:18-61).
// Checked value by value against the reference's own FFlLib (oracle/_ref/libfedem_ref_ffl.so) in
// tests/test_ftl_cpu.py.
#include <algorithm>
#include <cctype>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "common.cuh"

using namespace fsr;

namespace {

enum Cat { SOLID, SHELL, BEAM, CONSTRAINT, OTHER, STRC };
struct ElmType { const char* name; int sam; int nnod; int ndofs; Cat cat; };
// SAM codes: ffl_gettopol's typeMap (:574-594); node counts / DOFs / categories: FFlFEParts/FFl*.C init()
static const ElmType kTypes[] = {
    {"BEAM2", 11, 2, 6, BEAM},     {"BEAM3", 0, 3, 6, BEAM},      {"TRI3", 21, 3, 6, SHELL},
    {"QUAD4", 22, 4, 6, SHELL},    {"TRI6", 31, 6, 6, SHELL},     {"QUAD8", 32, 8, 6, SHELL},
    {"TET10", 41, 10, 3, SOLID},   {"WEDG15", 42, 15, 3, SOLID},  {"HEX20", 43, 20, 3, SOLID},
    {"HEX8", 44, 8, 3, SOLID},     {"TET4", 45, 4, 3, SOLID},     {"WEDG6", 46, 6, 3, SOLID},
    {"CMASS", 51, 1, 3, OTHER},    {"RGD", 61, 0, 6, CONSTRAINT}, {"RBAR", 62, 2, 6, CONSTRAINT},
    {"WAVGM", 63, 0, 0, CONSTRAINT}, {"SPRING", 71, 2, 3, OTHER}, {"RSPRING", 71, 2, 6, OTHER},
    {"BUSH", 72, 2, 6, OTHER},     {"STRCT3", 0, 3, 0, STRC},     {"STRCQ4", 0, 4, 0, STRC},
    {"STRCT6", 0, 6, 0, STRC},     {"STRCQ8", 0, 8, 0, STRC}};

struct Node {
  int id = 0, status = 0, dofs = 0;
  double x[3] = {0, 0, 0};
};
struct Elem {
  int id = 0;
  const ElmType* type = nullptr;
  std::vector<int> nodes;             // external ids as read; after resolve: indices into nodes_
  std::map<std::string, int> attr;    // attribute type -> id
  std::vector<int> pstrc;             // strain coat elements: every {PSTRC id} in file order (multiple references allowed)
  int fe = 0;                         // strain coat elements: {FE id}, the underlying finite element
  bool calc = true;
};
struct Field {                         // one LABEL{entries {REF id opts} ...} record
  std::string label;
  std::vector<std::string> entries;
  std::vector<std::pair<std::string, std::pair<std::vector<int>, std::vector<std::string>>>> refs;
};

static bool parse_int(int& v, const std::string& s)
{
  if (s.empty()) return true;
  char* e = nullptr;
  const long l = strtol(s.c_str(), &e, 10);
  if (*e == 0) { v = (int)l; return true; }
  // "12.000" is accepted as an integer without changing the value (parseNumericField :44-52)
  bool dot = false;
  for (const char* p = e; *p; ++p)
    if (!dot && *p == '.') dot = true;
    else if (!(dot && *p == '0')) return false;
  return true;
}

static bool parse_double(double& v, std::string s)
{
  if (s.empty()) return true;
  // Nastran-style exponents "1.5-3" get their 'E' (parseNumericField :67-78)
  for (int i = (int)s.size() - 1; i > 0; --i)
    if (s[i] == '-' || s[i] == '+') {
      if (s[i - 1] != 'e' && s[i - 1] != 'E') s.insert((size_t)i, 1, 'E');
      break;
    }
  char* e = nullptr;
  v = strtod(s.c_str(), &e);
  return *e == 0;
}

}  // namespace

struct fsr_ftl {
  std::vector<Node> nodes;                 // sorted by id
  std::vector<Elem> elems;                 // sorted by id
  std::map<std::string, std::map<int, std::vector<double>>> attrs;
  std::map<std::string, std::map<int, std::map<std::string, int>>> attr_refs;   // attribute -> the attributes it refers to
  std::map<int, std::string> pstrc_name;                                         // PSTRC id -> "Bottom" / "Mid" / "Top"
  std::map<int, std::vector<int>> groups;  // id -> element ids
  std::vector<int> fe_nodes;               // indices of nodes with DOFs        (internal node number - 1)
  std::vector<int> fe_elems;               // indices of the finite elements    (internal element number - 1)
  std::map<int, int> ext2int_node;
  int version = 0;
  unsigned long checksum = 0;

  // ---- tokenizer: the state machine of FFlFedemReader::getNextField ---------------------------------
  static bool next_field(FILE* f, Field& fl, unsigned long& cs, std::string& err)
  {
    fl.label.clear(); fl.entries.clear(); fl.refs.clear();
    int c;
    auto skip_line = [&]() { while ((c = fgetc(f)) != EOF && c != '\n') {} };
    auto token = [&](int& ch, bool upper) {
      std::string t;
      while (ch != EOF && isgraph(ch) && ch != '#' && ch != '{' && ch != '}' && ch != '"') {
        t += (char)(upper ? toupper(ch) : ch);
        ch = fgetc(f);
      }
      return t;
    };
    // label
    for (;;) {
      do c = fgetc(f); while (c != EOF && isspace(c));
      if (c == EOF) return false;
      if (c == '#') {
        std::string line;
        while ((c = fgetc(f)) != EOF && c != '\n') line += (char)c;
        if (line.size() > 15 && line.compare(0, 15, " File checksum:") == 0) cs = strtoul(line.c_str() + 15, nullptr, 10);
        continue;
      }
      while (c != EOF && isalnum(c)) { fl.label += (char)toupper(c); c = fgetc(f); }
      while (c != EOF && isspace(c)) c = fgetc(f);
      if (c == '#') { skip_line(); fl.label.clear(); continue; }  // label error: start over
      if (c != '{' || fl.label.empty()) { if (c != EOF && c != '\n') skip_line(); fl.label.clear(); if (c == EOF) return false; continue; }
      break;
    }
    if (fl.label == "FFT3") fl.label = "TRI3";
    else if (fl.label == "FFQ4") fl.label = "QUAD4";
    // entries and {REF id options} groups until the closing brace
    c = fgetc(f);
    for (;;) {
      while (c != EOF && isspace(c)) c = fgetc(f);
      if (c == EOF) { err = "premature end-of-file in " + fl.label; return false; }
      if (c == '#') { skip_line(); c = fgetc(f); continue; }
      if (c == '}') return true;
      if (c == '"') {
        std::string s;
        while ((c = fgetc(f)) != EOF && c != '"') s += (char)c;
        fl.entries.push_back(s);
        c = fgetc(f);
        continue;
      }
      if (c == '{') {
        c = fgetc(f);
        while (c != EOF && isspace(c)) c = fgetc(f);
        std::string name = token(c, true);
        if (name.empty()) { err = "unexpected character while reading a reference in " + fl.label; return false; }
        fl.refs.push_back({name, {}});
        auto& ref = fl.refs.back().second;
        for (;;) {   // ids, then free options
          while (c != EOF && isspace(c)) c = fgetc(f);
          if (c == '#') { skip_line(); c = fgetc(f); continue; }
          if (c != EOF && isdigit(c) && ref.second.empty()) {
            int id = 0;
            while (c != EOF && isdigit(c)) { id = id * 10 + (c - '0'); c = fgetc(f); }
            ref.first.push_back(id);
            continue;
          }
          if (c == '"') {
            std::string s;
            while ((c = fgetc(f)) != EOF && c != '"') s += (char)c;
            ref.second.push_back(s);
            c = fgetc(f);
            continue;
          }
          if (c == '}') { c = fgetc(f); break; }
          if (c == EOF || c == '{') { err = "malformed reference in " + fl.label; return false; }
          std::string o = token(c, true);
          if (!o.empty()) ref.second.push_back(o);
        }
        continue;
      }
      std::string e = token(c, true);
      if (!e.empty()) fl.entries.push_back(e);
      else { err = "unexpected character in " + fl.label; return false; }
    }
  }

  int find_node(int id) const
  {
    auto it = std::lower_bound(nodes.begin(), nodes.end(), id, [](const Node& n, int i) { return n.id < i; });
    return it != nodes.end() && it->id == id ? (int)(it - nodes.begin()) : -1;
  }
  const std::vector<double>* attribute(const Elem& e, const char* type) const
  {
    auto r = e.attr.find(type);
    if (r == e.attr.end()) return nullptr;
    auto t = attrs.find(type);
    if (t == attrs.end()) return nullptr;
    auto a = t->second.find(r->second);
    return a == t->second.end() ? nullptr : &a->second;
  }

  int read(const char* path)
  {
    FILE* f = fopen(path, "r");
    if (!f) { set_error("Can not open FE data file %s", path); return FSR_ERR_ARG; }
    Field fl;
    std::string err;
    int nerr = 0;
    while (next_field(f, fl, checksum, err)) {
      const ElmType* et = nullptr;
      for (const ElmType& t : kTypes) if (fl.label == t.name) et = &t;
      if (fl.label == "FTLVERSION") {
        if (fl.entries.empty() || !parse_int(version, fl.entries[0])) ++nerr;
      } else if (fl.label == "NODE") {
        Node n;
        fl.entries.resize(std::max<size_t>(fl.entries.size(), 5));
        if (!parse_int(n.id, fl.entries[0]) || !parse_int(n.status, fl.entries[1]) || !parse_double(n.x[0], fl.entries[2]) ||
            !parse_double(n.x[1], fl.entries[3]) || !parse_double(n.x[2], fl.entries[4])) ++nerr;
        nodes.push_back(n);
      } else if (et) {
        Elem e;
        e.type = et;
        if (fl.entries.empty() || !parse_int(e.id, fl.entries[0])) ++nerr;
        for (size_t i = 1; i < fl.entries.size(); ++i) {
          int n = 0;
          if (parse_int(n, fl.entries[i])) e.nodes.push_back(n); else ++nerr;
        }
        for (auto& r : fl.refs)
          if (!r.second.first.empty() && r.first[0] != 'V' && r.first != "FE") {
            std::string key = r.first == "PBEAMORIENT" || r.first == "PBUSHORIENT" ? "PORIENT" : r.first;
            e.attr[key] = r.second.first.front();
            if (key == "PSTRC") e.pstrc.push_back(r.second.first.front());
          } else if (r.first == "FE" && !r.second.first.empty())
            e.fe = r.second.first.front();
        elems.push_back(e);
      } else if (fl.label == "GROUP") {
        int id = 0;
        if (fl.entries.empty() || !parse_int(id, fl.entries[0])) ++nerr;
        std::vector<int>& g = groups[id];
        for (size_t i = 1; i < fl.entries.size(); ++i) {
          int el = 0;
          if (parse_int(el, fl.entries[i])) g.push_back(el); else ++nerr;
        }
      } else if (!fl.label.empty() && fl.label[0] == 'P') {   // attribute record: id + numeric fields
        int id = 0;
        if (fl.entries.empty() || !parse_int(id, fl.entries[0])) ++nerr;
        std::string key = fl.label == "PBEAMORIENT" || fl.label == "PBUSHORIENT" ? "PORIENT" : fl.label;
        std::vector<double> v;
        for (size_t i = 1; i < fl.entries.size(); ++i) {
          double d = 0.0;
          if (parse_double(d, fl.entries[i])) v.push_back(d);   // text fields (names, types) are not needed here
          else v.push_back(0.0);
        }
        attrs[key][id] = v;
        for (auto& r : fl.refs) if (!r.second.first.empty()) attr_refs[key][id][r.first] = r.second.first.front();
        if (key == "PSTRC" && fl.entries.size() > 1) pstrc_name[id] = fl.entries[1];
      }
      // visuals, loads, coordinate systems etc. are not on the recovery path and are skipped
    }
    fclose(f);
    if (!err.empty()) { set_error("%s: %s. The FE data file is corrupt.", path, err.c_str()); return FSR_ERR_ARG; }
    if (nerr) { set_error("%s: %d syntax errors. The FE data file is corrupt.", path, nerr); return FSR_ERR_ARG; }
    return resolve(path);
  }

  // FFlLinkHandler::resolve (:1709-1870) + buildFiniteElementVec (:826-880), restricted to what decides
  // the node / element numbering
  int resolve(const char* path)
  {
    std::stable_sort(nodes.begin(), nodes.end(), [](const Node& a, const Node& b) { return a.id < b.id; });
    std::stable_sort(elems.begin(), elems.end(), [](const Elem& a, const Elem& b) { return a.id < b.id; });
    nodes.erase(std::unique(nodes.begin(), nodes.end(), [](const Node& a, const Node& b) { return a.id == b.id; }), nodes.end());
    elems.erase(std::unique(elems.begin(), elems.end(), [](const Elem& a, const Elem& b) { return a.id == b.id; }), elems.end());
    if (nodes.empty()) { set_error("%s: No nodes!", path); return FSR_ERR_ARG; }
    for (Elem& e : elems) {
      if (e.type->nnod > 0 && (int)e.nodes.size() != e.type->nnod) {
        set_error("%s: %s element %d has %d nodes, expected %d", path, e.type->name, e.id, (int)e.nodes.size(), e.type->nnod);
        return FSR_ERR_ARG;
      }
      int local = 0;
      for (int& nid : e.nodes) {
        const int k = find_node(nid);
        if (k < 0) { set_error("%s: Resolving %s element %d failed (node %d)", path, e.type->name, e.id, nid); return FSR_ERR_ARG; }
        nid = k;
        ++local;
        const bool rgd = e.type->sam == 61, wavgm = e.type->sam == 63;
        const int dofs = rgd ? (local == 1 ? 6 : 0) : e.type->ndofs;       // FFlRGD.H:30
        nodes[k].dofs = std::max(nodes[k].dofs, dofs);
        if (wavgm && local == 1) nodes[k].status = 3;                      // reference node, cannot be external
        else if (e.type->sam == 62) nodes[k].status = 2;                   // RBAR: both nodes are slaves
      }
      for (auto& a : e.attr) {
        auto t = attrs.find(a.first);
        if (t == attrs.end() || !t->second.count(a.second)) {
          set_error("%s: Resolving %s element %d failed (%s %d)", path, e.type->name, e.id, a.first.c_str(), a.second);
          return FSR_ERR_ARG;
        }
      }
    }
    for (auto& g : groups)   // FFlGroup::resolveElemRefs (FFlGroup.C:120-133): a group may only name existing elements
      for (int el : g.second) {
        auto e = std::lower_bound(elems.begin(), elems.end(), el, [](const Elem& a, int i) { return a.id < i; });
        if (e == elems.end() || e->id != el) {
          set_error("%s: Invalid element Id %d. Resolving element group %d failed", path, el, g.first);
          return FSR_ERR_ARG;
        }
      }
    // WAVGM elements: loose master nodes are dropped, elements left with the reference node only are erased
    std::vector<Elem> kept;
    kept.reserve(elems.size());
    for (Elem& e : elems) {
      if (e.nodes.empty()) continue;
      if (e.type->sam == 63) {
        std::vector<int> nn(1, e.nodes[0]);
        for (size_t i = 1; i < e.nodes.size(); ++i) {
          Node& n = nodes[e.nodes[i]];
          if (n.dofs < 1 && n.status == 1) n.dofs = 6;
          if (n.dofs >= 1) nn.push_back(e.nodes[i]);
        }
        if (nn.size() < 2) continue;
        nodes[e.nodes[0]].dofs = std::max(nodes[e.nodes[0]].dofs, 6);
        e.nodes.swap(nn);
      }
      kept.push_back(e);
    }
    elems.swap(kept);
    for (size_t i = 0; i < nodes.size(); ++i)
      if (nodes[i].dofs >= 1) { fe_nodes.push_back((int)i); ext2int_node[nodes[i].id] = (int)fe_nodes.size(); }
    for (size_t i = 0; i < elems.size(); ++i) {
      const Elem& e = elems[i];
      if (e.type->cat == STRC) continue;
      int nelnod = 0, lerr = 0;
      for (int k : e.nodes)
        if (nodes[k].dofs >= 1) ++nelnod;
        else if (nelnod == 0 && (e.type->sam == 63 || e.type->sam == 51)) break;
        else ++lerr;
      if (e.type->sam == 61 && nelnod < 2) continue;
      if ((e.type->sam == 63 && nelnod < 2 && lerr == 0) || (e.type->sam == 51 && nelnod < 1)) continue;
      fe_elems.push_back((int)i);
    }
    return FSR_OK;
  }
};

extern "C" {

int fsr_ftl_open(fsr_ftl** ftl, const char* path)
{
  if (!ftl || !path) { set_error("fsr_ftl_open: bad arguments"); return FSR_ERR_ARG; }
  fsr_ftl* h = new fsr_ftl;
  const int rc = h->read(path);
  if (rc) { delete h; *ftl = nullptr; return rc; }
  *ftl = h;
  return FSR_OK;
}

void fsr_ftl_close(fsr_ftl* ftl) { delete ftl; }

// FFl::activateElmGroups (FFlUtils.C:18-61): "55", "<33,22,44>", "<PMAT 33, PTHICK 55>" or a mixture
int fsr_ftl_activate_groups(fsr_ftl* h, const char* groups)
{
  if (!h) { set_error("fsr_ftl_activate_groups: bad arguments"); return FSR_ERR_ARG; }
  std::string g = groups ? groups : "";
  for (Elem& e : h->elems) e.calc = g.empty();
  if (g.empty()) return FSR_OK;
  std::vector<int> ids;
  std::vector<std::pair<std::string, int>> implicit;
  if (g[0] == '<') {
    std::string tok;
    auto flush = [&]() {
      if (tok.empty()) return;
      if (isdigit((unsigned char)tok[0])) ids.push_back(atoi(tok.c_str()));
      else {
        size_t s = tok.size() + 1;
        while (--s > 0 && isdigit((unsigned char)tok[s - 1])) {}
        if (s > 1) implicit.push_back({tok.substr(0, s), atoi(tok.c_str() + s)});
      }
      tok.clear();
    };
    for (size_t i = 1; i < g.size() && g[i] != '>'; ++i)
      if (g[i] == ',') flush();
      else if (!isspace((unsigned char)g[i])) tok += (char)toupper((unsigned char)g[i]);
    flush();
  } else
    ids.push_back(atoi(g.c_str()));
  if (ids.empty() && implicit.empty()) { set_error("invalid element group specification '%s'", g.c_str()); return FSR_ERR_ARG; }
  int warnings = 0;
  for (int id : ids) {
    auto it = h->groups.find(id);
    if (it == h->groups.end()) { ++warnings; continue; }   // "Non-existing element group ignored"
    for (int el : it->second) {
      auto e = std::lower_bound(h->elems.begin(), h->elems.end(), el, [](const Elem& a, int i) { return a.id < i; });
      if (e != h->elems.end() && e->id == el) e->calc = true;
    }
  }
  for (auto& p : implicit)
    for (Elem& e : h->elems) {
      auto a = e.attr.find(p.first);
      if (a != e.attr.end() && a->second == p.second && h->attrs.count(p.first) && h->attrs[p.first].count(p.second)) e.calc = true;
    }
  return warnings;
}

// ffl_getsize: sz[0..12) = nnod, nel, ndof, nmnpc, nmat, nxnod, npbeam, nrgd, nrbar, nwavgm, nprop, ncons;
// returns the number of finite elements with the calculation flag on
int fsr_ftl_sizes(const fsr_ftl* h, int* sz)
{
  if (!h || !sz) { set_error("fsr_ftl_sizes: bad arguments"); return FSR_ERR_ARG; }
  int nnod = 0, ndof = 0, nmnpc = 0, nxnod = 0, npbeam = 0, nrgd = 0, nrbar = 0, nwavgm = 0, nael = 0;
  for (int k : h->fe_nodes) { ++nnod; ndof += h->nodes[k].dofs; }
  for (int i : h->fe_elems) {
    const Elem& e = h->elems[i];
    nmnpc += (int)e.nodes.size();
    if (e.type->sam == 11) {
      if (const std::vector<double>* pin = h->attribute(e, "PBEAMPIN")) {
        ++npbeam;
        if (pin->size() > 0 && (*pin)[0] > 0) ++nxnod;
        if (pin->size() > 1 && (*pin)[1] > 0) ++nxnod;
      }
    } else if (e.type->sam == 61) ++nrgd;
    else if (e.type->sam == 62) ++nrbar;
    else if (e.type->sam == 63) ++nwavgm;
  }
  for (const Elem& e : h->elems) if (e.calc) ++nael;   // getElementCount(FFL_ALL, true): strain coats included
  auto count = [&](const char* t) { auto it = h->attrs.find(t); return it == h->attrs.end() ? 0 : (int)it->second.size(); };
  sz[0] = nnod + nxnod; sz[1] = (int)h->fe_elems.size(); sz[2] = ndof + 6 * nxnod; sz[3] = nmnpc; sz[4] = count("PMAT");
  sz[5] = nxnod; sz[6] = npbeam; sz[7] = nrgd; sz[8] = nrbar; sz[9] = nwavgm;
  sz[10] = count("PTHICK") + count("PBEAMSECTION") + count("PNSM"); sz[11] = nrgd + nrbar + nwavgm;
  return nael;
}

static int pin_status(int flag, int* msc)   // resolvePinFlag (:436-451)
{
  if (flag <= 0) return 0;
  int n = 6;
  while (flag > 0) {
    const int l = flag % 10;
    flag /= 10;
    while (n > l) msc[--n] = 0;
    msc[--n] = 1;
  }
  while (n > 0) msc[--n] = 0;
  return 6;
}

// ffl_getnodes: madof [nnod+1], minex [nnod], mnode [nnod] (2 external / 1 internal), msc [ndof],
// xyz [nnod][3]; any may be NULL.  Returns nnod.
int fsr_ftl_get_nodes(const fsr_ftl* h, int* madof, int* minex, int* mnode, int* msc, double* xyz)
{
  if (!h) { set_error("fsr_ftl_get_nodes: bad arguments"); return FSR_ERR_ARG; }
  int inod = 0, ndof = 0;
  if (madof) madof[0] = 1;
  auto put = [&](const Node& n, int id, int code, int nd, const int* sc) {
    if (minex) minex[inod] = id;
    if (mnode) mnode[inod] = code;
    if (xyz) { xyz[3 * inod] = n.x[0]; xyz[3 * inod + 1] = n.x[1]; xyz[3 * inod + 2] = n.x[2]; }
    if (msc) for (int i = 0; i < nd; ++i) msc[ndof + i] = sc[i];
    ndof += nd;
    ++inod;
    if (madof) madof[inod] = ndof + 1;
  };
  for (int k : h->fe_nodes) {
    const Node& n = h->nodes[k];
    if (n.dofs != 3 && n.dofs != 6) { set_error("Invalid DOFs for node %d : %d", n.id, n.dofs); return FSR_ERR_ARG; }
    const int code = n.status == 1 ? 2 : 1;
    int sc[6];
    for (int i = 0; i < n.dofs; ++i) sc[i] = (n.status < 0 && ((-n.status) & (1 << i))) ? 0 : code;
    put(n, n.id, code, n.dofs, sc);
  }
  for (int i : h->fe_elems) {
    const Elem& e = h->elems[i];
    if (e.type->sam != 11) continue;
    const std::vector<double>* pin = h->attribute(e, "PBEAMPIN");
    if (!pin) continue;
    for (int end = 0; end < 2; ++end) {
      int sc[6];
      if (pin_status(end < (int)pin->size() ? (int)(*pin)[end] : 0, sc)) put(h->nodes[e.nodes[end]], -inod - 1, 1, 6, sc);
    }
  }
  return inod;
}

// ffl_gettopol: melcon [nel], mpmnpc [nel+1], mmnpc [nmnpc] (internal node numbers).  Returns nel.
int fsr_ftl_get_topology(const fsr_ftl* h, int use_andes, int* melcon, int* mpmnpc, int* mmnpc)
{
  if (!h || !melcon || !mpmnpc || !mmnpc) { set_error("fsr_ftl_get_topology: bad arguments"); return FSR_ERR_ARG; }
  int nel = 0, n = 0;
  mpmnpc[0] = 1;
  for (int i : h->fe_elems) {
    const Elem& e = h->elems[i];
    int code = e.type->sam;
    if ((code == 21 || code == 22) && use_andes) code += 2;
    else if (code == 51 && !e.attr.count("PMASS")) code = 50;
    else if (code == 72 && !e.attr.count("PBUSHCOEFF")) code = 70;
    melcon[nel] = code;
    for (int k : e.nodes) {
      auto it = h->ext2int_node.find(h->nodes[k].id);
      if (it != h->ext2int_node.end()) mmnpc[n++] = it->second;   // DOF-less nodes are removed from the topology
    }
    if (code == 31) {   // mid-side nodes last: 1-2-3-4-5-6 -> 1-3-5-2-4-6
      std::swap(mmnpc[n - 5], mmnpc[n - 4]);
      std::swap(mmnpc[n - 4], mmnpc[n - 2]);
      std::swap(mmnpc[n - 3], mmnpc[n - 2]);
    }
    mpmnpc[++nel] = n + 1;
  }
  return nel;
}

// FaMat34::makeGlobalizedCS(p1, p2)[VZ] (FFaMat33.C:204-237): the default beam Z-axis
static void globalized_z(const double* d, double* ez)
{
  double ex[3] = {d[0], d[1], d[2]};
  double len = sqrt(ex[0] * ex[0] + ex[1] * ex[1] + ex[2] * ex[2]);
  if (len < 1.0e-15) { ex[0] = 1.0; ex[1] = ex[2] = 0.0; }
  else for (double& v : ex) v /= len;
  auto normalize = [](double* v) {
    const double l = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    if (l < 1.0e-15) { v[0] = 1.0; v[1] = v[2] = 0.0; }
    else for (int i = 0; i < 3; ++i) v[i] /= l;
  };
  if (fabs(ex[2]) > fabs(ex[1])) {
    double ey[3] = {-ex[1] * ex[0], ex[0] * ex[0] + ex[2] * ex[2], -ex[1] * ex[2]};
    normalize(ey);
    ez[0] = ex[1] * ey[2] - ex[2] * ey[1];
    ez[1] = ex[2] * ey[0] - ex[0] * ey[2];
    ez[2] = ex[0] * ey[1] - ex[1] * ey[0];
  } else {
    ez[0] = -ex[2] * ex[0]; ez[1] = -ex[2] * ex[1]; ez[2] = ex[0] * ex[0] + ex[1] * ex[1];
    normalize(ez);
  }
}

// Per finite element, in SAM order: what ffl_getmat / ffl_getthick / ffl_getelmid / ffl_getcoor (beams) /
// ffl_getbeamsection / ffl_getpinflags return.  emod, rny, rho, thk [nel]; elmid [nel] (negative = outside the
// -group selection); beam [nel][FSR_NBEAM] in the layout of fsr_elmdata.beam; status [nel]: 0 ok, -2 no
// material, -3 Poisson's ratio outside [0,0.5> or no beam section, -4 no thickness on a shell.  Any may be
// NULL.  Returns the number of elements with status != 0.
int fsr_ftl_get_elmdata(const fsr_ftl* h, double* emod, double* rny, double* rho, double* thk, int* elmid, double* beam,
                        int* status)
{
  if (!h) { set_error("fsr_ftl_get_elmdata: bad arguments"); return FSR_ERR_ARG; }
  int iel = 0, bad = 0;
  for (int i : h->fe_elems) {
    const Elem& e = h->elems[i];
    int st = 0;
    double E = 0, nu = 0, r = 0, t = 0;
    const std::vector<double>* mat = h->attribute(e, "PMAT");
    auto val = [](const std::vector<double>* v, size_t k) { return v && k < v->size() ? (*v)[k] : 0.0; };
    const bool structural = e.type->cat == SOLID || e.type->cat == SHELL || e.type->cat == BEAM;
    if (mat) { E = val(mat, 0); nu = val(mat, 2); r = val(mat, 3); if (!(nu >= 0.0 && nu < 0.5)) st = -3; }
    else if (structural) st = -2;
    if (e.type->cat == SHELL) {
      const std::vector<double>* th = h->attribute(e, "PTHICK");
      if (th) t = val(th, 0); else if (!st) st = -4;
    }
    if (beam) {
      double* b = beam + (size_t)iel * FSR_NBEAM;
      for (int k = 0; k < FSR_NBEAM; ++k) b[k] = 0.0;
      if (e.type->sam == 11) {
        double* X = b; double* Y = b + 5; double* Z = b + 10; double* sec = b + 15;
        for (int k = 0; k < 2; ++k) { const Node& n = h->nodes[e.nodes[k]]; X[k] = n.x[0]; Y[k] = n.x[1]; Z[k] = n.x[2]; }
        double zax[3] = {0, 0, 0};
        if (const std::vector<double>* o = h->attribute(e, "PORIENT")) for (int k = 0; k < 3; ++k) zax[k] = val(o, k);
        if (fabs(zax[0]) <= 1.0e-10 && fabs(zax[1]) <= 1.0e-10 && fabs(zax[2]) <= 1.0e-10) {
          const double d[3] = {X[1] - X[0], Y[1] - Y[0], Z[1] - Z[0]};
          globalized_z(d, zax);
        }
        X[2] = X[0] + zax[0]; Y[2] = Y[0] + zax[1]; Z[2] = Z[0] + zax[2];
        X[3] = X[0]; Y[3] = Y[0]; Z[3] = Z[0];
        X[4] = X[1]; Y[4] = Y[1]; Z[4] = Z[1];
        if (const std::vector<double>* ec = h->attribute(e, "PBEAMECCENT")) {
          X[0] += val(ec, 0); Y[0] += val(ec, 1); Z[0] += val(ec, 2);
          X[1] += val(ec, 3); Y[1] += val(ec, 4); Z[1] += val(ec, 5);
          X[2] += val(ec, 0); Y[2] += val(ec, 1); Z[2] += val(ec, 2);
        }
        const std::vector<double>* s = h->attribute(e, "PBEAMSECTION");
        if (!s && !st) st = -3;
        if (mat && s) {
          sec[0] = val(mat, 3); sec[1] = val(mat, 0); sec[2] = val(mat, 1); sec[3] = val(s, 0);
          sec[4] = val(s, 1); sec[5] = val(s, 2); sec[6] = val(s, 3);
          const double ixx = sec[4] + sec[5];
          sec[7] = ixx > 0.0 ? ixx : sec[6];
          sec[8] = val(s, 4) > 0.0 ? 1.0 / val(s, 4) : 0.0;   // the file stores As/A, the beam routine wants A/As
          sec[9] = val(s, 5) > 0.0 ? 1.0 / val(s, 5) : 0.0;
          sec[10] = val(s, 6); sec[11] = val(s, 7); sec[13] = val(s, 8);
          sec[12] = val(h->attribute(e, "PEFFLENGTH"), 0);
        }
        if (const std::vector<double>* pin = h->attribute(e, "PBEAMPIN")) { b[29] = val(pin, 0); b[30] = val(pin, 1); }
      }
    }
    if (emod) emod[iel] = E;
    if (rny) rny[iel] = nu;
    if (rho) rho[iel] = r;
    if (thk) thk[iel] = t;
    if (elmid) elmid[iel] = e.calc ? e.id : -e.id;
    if (status) status[iel] = st;
    if (st) ++bad;
    ++iel;
  }
  return bad;
}

// ffl_ext2int (:716-737): internal number of a node (is_node != 0) or finite element; 0/-1 when absent
// ffl_getnostrc (FFlLinkHandler_F.C:1753-1762): strain coat elements whose calculation flag is on
int fsr_ftl_num_strain_coats(const fsr_ftl* h)
{
  if (!h) return FSR_ERR_ARG;
  int n = 0;
  for (const Elem& e : h->elems) if (e.type->cat == STRC && e.calc) ++n;
  return n;
}

// ffl_getstraincoat (:1640-1701) + getStrainCoatAttributes (:1587-1630) for all strain coat elements in element order
int fsr_ftl_get_strain_coats(const fsr_ftl* h, fsr_strain_coat* out, int cap)
{
  if (!h || (cap > 0 && !out)) { set_error("fsr_ftl_get_strain_coats: bad arguments"); return FSR_ERR_ARG; }
  auto ref_of = [&](const char* type, int id, const char* to) -> int {
    auto t = h->attr_refs.find(type);
    if (t == h->attr_refs.end()) return 0;
    auto a = t->second.find(id);
    if (a == t->second.end()) return 0;
    auto r = a->second.find(to);
    return r == a->second.end() ? 0 : r->second;
  };
  auto values = [&](const char* type, int id) -> const std::vector<double>* {
    auto t = h->attrs.find(type);
    if (t == h->attrs.end()) return nullptr;
    auto a = t->second.find(id);
    return a == t->second.end() ? nullptr : &a->second;
  };
  int n = 0;
  for (const Elem& e : h->elems) {
    if (e.type->cat != STRC || !e.calc) continue;
    if (n < cap) {
      fsr_strain_coat& c = out[n];
      memset(&c, 0, sizeof(c));
      c.id = e.id;
      // every second node of the 6- and 8-noded elements is skipped
      for (size_t i = 0; i < e.nodes.size(); i += e.nodes.size() > 4 ? 2 : 1) {
        auto it = h->ext2int_node.find(h->nodes[(size_t)e.nodes[i]].id);
        c.nodes[c.nnod++] = it == h->ext2int_node.end() ? -1 : it->second;
      }
      const std::vector<double>* fat = nullptr;
      { auto f = e.attr.find("PFATIGUE"); if (f != e.attr.end()) fat = values("PFATIGUE", f->second); }
      for (int ps : e.pstrc) {
        if (c.npts >= 3) { set_error("Invalid strain coat definition. Too many points: element %d", e.id); return FSR_ERR_ARG; }
        const int k = c.npts++;
        c.sn_curve[k][0] = c.sn_curve[k][1] = -1;
        auto nm = h->pstrc_name.find(ps);
        std::string name = nm == h->pstrc_name.end() ? "" : nm->second;
        for (char& ch : name) ch = (char)toupper((unsigned char)ch);
        c.res_set[k] = name == "BOTTOM" ? 1 : name == "MID" ? 2 : name == "TOP" ? 3 : 0;
        if (const int mid = ref_of("PSTRC", ps, "PMAT"))
          if (const std::vector<double>* m = values("PMAT", mid)) {
            c.mat_id[k] = mid;
            c.emod[k] = m->size() > 0 ? (*m)[0] : 0.0;
            c.nu[k] = m->size() > 2 ? (*m)[2] : 0.0;
          }
        if (const int hid = ref_of("PSTRC", ps, "PHEIGHT")) {
          if (const std::vector<double>* hv = values("PHEIGHT", hid)) c.zpos[k] = hv->empty() ? 0.0 : (*hv)[0];
        } else if (const int tid = ref_of("PSTRC", ps, "PTHICKREF")) {
          const std::vector<double>* tr = values("PTHICKREF", tid);
          const int thid = ref_of("PTHICKREF", tid, "PTHICK");
          const std::vector<double>* th = thid ? values("PTHICK", thid) : nullptr;
          if (tr && th && !th->empty()) c.zpos[k] = (*th)[0] * (tr->empty() ? 0.0 : (*tr)[0]);
        }
        if (fat) {
          c.sn_curve[k][0] = fat->size() > 0 ? (int)(*fat)[0] : 0;
          c.sn_curve[k][1] = fat->size() > 1 ? (int)(*fat)[1] : 0;
          c.scf[k] = fat->size() > 2 ? (*fat)[2] : 0.0;
        }
      }
      c.elm_id = e.fe;
    }
    ++n;
  }
  return n;
}

int fsr_ftl_ext2int(const fsr_ftl* h, int is_node, int id)
{
  if (!h || id <= 0) return 0;
  if (is_node) { auto it = h->ext2int_node.find(id); return it == h->ext2int_node.end() ? -1 : it->second; }
  for (size_t k = 0; k < h->fe_elems.size(); ++k) if (h->elems[h->fe_elems[k]].id == id) return (int)k + 1;
  return 0;
}

int fsr_ftl_version(const fsr_ftl* h) { return h ? h->version : 0; }

}  // extern "C"
