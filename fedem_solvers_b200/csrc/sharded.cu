// sharded.cu -- element-block sharding of one FE part over the GPUs of a box, behind the C ABI.
//
// The reference recovers one part per process in one serial element loop (src/vpmStress/stress.f90:126-128,
// src/vpmStress/stressRoutines.f90:169; in the solver, src/vpmSolver/stressRecoveryModule.f90:1021-1061 loops over the
// parts).  Elements are independent given the nodal displacements and result points belong to elements, so here a part is
// cut into contiguous element blocks of equal cost; a block is a self-contained part handle (own SamType arrays, the B / E
// rows of every node it touches, seam nodes duplicated, all external DOFs so that the reduced history Q is the parent's).
//   fsr_split_elements / fsr_part_create_block / fsr_set_recovery_parent : the cut, native (no host-language helper needed)
//   fsr_group_*  : one process, several GPUs -- the blocks of a part on the listed devices, Q broadcast with ncclBroadcast,
//                  per-block envelopes gathered with ncclSend / ncclRecv over NVLink into the parent's result-point order
//   fsr_comm_*   : one process per GPU (MPI / torchrun style hosts): the same two collectives on a communicator built from
//                  an ncclUniqueId that the host passes around
// NCCL is bound at run time (dlopen of libnccl.so.2: the copy already in the process, e.g. PyTorch's, else the system's), so
// single-GPU users of the library do not need it.
#include <dlfcn.h>

#include <algorithm>
#include <cmath>
#include <mutex>
#include <numeric>

#include "common.cuh"

namespace fsr {

// ---- NCCL, bound lazily ---------------------------------------------------------------------------------
struct NcclApi {
  void* h = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  ncclResult_t (*GetVersion)(int*) = nullptr;
};

static NcclApi* nccl_api()
{
  static NcclApi api;
  static std::once_flag once;
  static bool ok = false;
  std::call_once(once, [] {
    for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
      api.h = dlopen(name, RTLD_NOW | RTLD_LOCAL);
      if (api.h) break;
    }
    if (!api.h) return;
    bool all = true;
    auto sym = [&](const char* n) { void* p = dlsym(api.h, n); if (!p) all = false; return p; };
    api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
    api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
    api.CommInitAll = (decltype(api.CommInitAll))sym("ncclCommInitAll");
    api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
    api.Broadcast = (decltype(api.Broadcast))sym("ncclBroadcast");
    api.Send = (decltype(api.Send))sym("ncclSend");
    api.Recv = (decltype(api.Recv))sym("ncclRecv");
    api.GroupStart = (decltype(api.GroupStart))sym("ncclGroupStart");
    api.GroupEnd = (decltype(api.GroupEnd))sym("ncclGroupEnd");
    api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
    api.GetVersion = (decltype(api.GetVersion))sym("ncclGetVersion");
    ok = all;
  });
  if (!ok) { set_error("NCCL is not available (dlopen of libnccl.so.2 failed: %s)", dlerror() ? dlerror() : "missing symbols"); return nullptr; }
  return &api;
}

#define FSR_NCCL(api, call)                                                                                  \
  do {                                                                                                       \
    ncclResult_t r_ = (call);                                                                                \
    if (r_ != ncclSuccess) {                                                                                 \
      set_error("%s failed at %s:%d: %s", #call, __FILE__, __LINE__, (api)->GetErrorString(r_));             \
      return FSR_ERR_CUDA;                                                                                   \
    }                                                                                                        \
  } while (0)

// ---- the cut ----------------------------------------------------------------------------------------------
// relative cost of one element.step = the K1 rows it brings (nodal DOFs x n_red, shared with the neighbours) + its K2 kernel,
// in measured picoseconds per element.step on B200 (profiles/R4_bench.json, R4_bench_configs.json, the per-piece times of
// R5_bench_c4_n8.json); the K1 share is quoted at n_red = 98 and scales with the reduced dimension f = n_red / 98.  Flat quads
// and straight-sided TET10, the common case; the same table as partition.py (ELEMENT_K1 / ELEMENT_K2).
static double element_cost(int type, double f)
{
  switch (type) {
    case 24: case 22: return 24.0 + 26.0 * f;   // flat regions on in-plane rows; 38 + 39 f on six global rows
    case 23: case 21: return 37.0 + 20.0 * f;
    case 41: return 31.0 + 23.0 * f;     // step-lane kernel
    case 42: return 170.0 + 70.0 * f;
    case 43: return 149.0 + 63.0 * f;    // step-lane kernel
    case 44: return 67.0 + 14.0 * f;
    case 45: return 36.0 + 8.0 * f;
    case 46: return 57.0 + 12.0 * f;
    case 11: return 2.0 + 1.0 * f;
    case 31: return 160.0 + 56.0 * f;
    case 32: return 230.0 + 78.0 * f;
    default: return 0.0;
  }
}

struct BlockArrays {
  std::vector<int> madof, msc, mpmnpc, mmnpc, melcon, mpmceq, mmceq, meqn, meqn1, meqn2, elmid, nodes, rows1;
  std::vector<double> ttcc, xyz, emod, rny, thk, beam;
  fsr_sam sam;
  fsr_elmdata elm;
  int pt0 = 0, parent_npts = 0;
};

// The element block [e0, e1) of a part as SAM arrays of its own: the block's nodes, all external nodes (ndof2 and Q stay
// the parent's) and, by closure, the master nodes of every constraint equation a kept DOF depends on.  Equation and
// constraint numbers are renumbered in ascending order, which keeps the order of meqn1 (the rows of B and E).
static int make_block(const fsr_sam* s, const fsr_elmdata* el, const int* melcon_eff, int e0, int e1, BlockArrays& o)
{
  const int nnod = s->nnod, ndof = s->ndof;
  std::vector<char> keep((size_t)nnod + 1, 0);
  std::vector<int> node_of_dof((size_t)ndof);
  for (int n = 0; n < nnod; ++n)
    for (int d = s->madof[n] - 1; d < s->madof[n + 1] - 1; ++d) node_of_dof[(size_t)d] = n + 1;
  const int ip0 = s->mpmnpc[e0] - 1, ip1 = s->mpmnpc[e1] - 1;
  for (int ip = ip0; ip < ip1; ++ip) {
    const int n = s->mmnpc[ip];
    if (n < 1 || n > nnod) { set_error("element connectivity: node %d out of range", n); return FSR_ERR_ARG; }
    keep[(size_t)n] = 1;
  }
  for (int d = 0; d < ndof; ++d) if (s->msc[d] == 2) keep[(size_t)node_of_dof[(size_t)d]] = 1;
  if (s->nceq > 0) {
    std::vector<int> stack;
    for (int d = 0; d < ndof; ++d) if (keep[(size_t)node_of_dof[(size_t)d]] && s->meqn[d] < 0) stack.push_back(d);
    while (!stack.empty()) {
      const int d = stack.back();
      stack.pop_back();
      const int ic = -s->meqn[d];
      if (ic < 1 || ic > s->nceq) continue;
      for (int ip = s->mpmceq[ic - 1] + 1; ip <= s->mpmceq[ic] - 1; ++ip) {
        const int m = s->mmceq[ip - 1];
        if (m < 1 || m > ndof) continue;
        const int nn = node_of_dof[(size_t)m - 1];
        if (keep[(size_t)nn]) continue;
        keep[(size_t)nn] = 1;
        for (int dd = s->madof[nn - 1] - 1; dd < s->madof[nn] - 1; ++dd) if (s->meqn[dd] < 0) stack.push_back(dd);
      }
    }
  }
  std::vector<int> newnode((size_t)nnod + 1, 0), newdof((size_t)ndof + 1, 0);
  o.madof.assign(1, 1);
  for (int n = 1; n <= nnod; ++n) {
    if (!keep[(size_t)n]) continue;
    o.nodes.push_back(n);
    newnode[(size_t)n] = (int)o.nodes.size();
    for (int d = s->madof[n - 1]; d < s->madof[n]; ++d) { newdof[(size_t)d] = (int)o.msc.size() + 1; o.msc.push_back(s->msc[d - 1]); }
    o.madof.push_back((int)o.msc.size() + 1);
    for (int k = 0; k < 3; ++k) o.xyz.push_back(el->xyz[3 * (size_t)(n - 1) + k]);
  }
  const int ndof_b = (int)o.msc.size();
  // equations and constraint equations of the kept DOFs, renumbered in ascending order of their old numbers
  std::vector<int> neweq((size_t)s->neq + 1, 0), newceq((size_t)s->nceq + 1, 0);
  for (int d = 1; d <= ndof; ++d) {
    if (!newdof[(size_t)d]) continue;
    const int q = s->meqn[d - 1];
    if (q > 0 && q <= s->neq) neweq[(size_t)q] = 1;
    else if (q < 0 && -q <= s->nceq) newceq[(size_t)-q] = 1;
  }
  int neq_b = 0, nceq_b = 0;
  for (int q = 1; q <= s->neq; ++q) if (neweq[(size_t)q]) neweq[(size_t)q] = ++neq_b;
  o.mpmceq.assign(1, 1);
  for (int c = 1; c <= s->nceq; ++c) {
    if (!newceq[(size_t)c]) continue;
    newceq[(size_t)c] = ++nceq_b;
    for (int ip = s->mpmceq[c - 1]; ip < s->mpmceq[c]; ++ip) {   // incl. the leading (dependent, c0) entry
      const int m = s->mmceq[ip - 1];
      o.mmceq.push_back(m > 0 && m <= ndof ? newdof[(size_t)m] : 0);
      o.ttcc.push_back(s->ttcc[ip - 1]);
    }
    o.mpmceq.push_back((int)o.mmceq.size() + 1);
  }
  o.meqn.assign((size_t)ndof_b, 0);
  for (int d = 1; d <= ndof; ++d) {
    const int nd = newdof[(size_t)d];
    if (!nd) continue;
    const int q = s->meqn[d - 1];
    o.meqn[(size_t)nd - 1] = q > 0 && q <= s->neq ? neweq[(size_t)q] : q < 0 && -q <= s->nceq ? -newceq[(size_t)-q] : 0;
  }
  for (int k = 0; k < s->ndof1; ++k) {
    const int q = s->meqn1[k];
    if (q >= 1 && q <= s->neq && neweq[(size_t)q]) { o.rows1.push_back(k); o.meqn1.push_back(neweq[(size_t)q]); }
  }
  for (int j = 0; j < s->ndof2; ++j) {
    const int q = s->meqn2[j];
    if (q < 1 || q > s->neq || !neweq[(size_t)q]) { set_error("external DOF %d lost in the element block", j + 1); return FSR_ERR_ARG; }
    o.meqn2.push_back(neweq[(size_t)q]);
  }
  // elements
  const int nel_b = e1 - e0;
  o.mpmnpc.resize((size_t)nel_b + 1);
  for (int e = e0; e <= e1; ++e) o.mpmnpc[(size_t)(e - e0)] = s->mpmnpc[e] - s->mpmnpc[e0] + 1;
  o.mmnpc.resize((size_t)(ip1 - ip0));
  for (int ip = ip0; ip < ip1; ++ip) o.mmnpc[(size_t)(ip - ip0)] = newnode[(size_t)s->mmnpc[ip]];
  o.melcon.assign(melcon_eff + e0, melcon_eff + e1);
  o.emod.assign(el->emod + e0, el->emod + e1);
  o.rny.assign(el->rny + e0, el->rny + e1);
  o.thk.assign(el->thk + e0, el->thk + e1);
  if (el->elmid) o.elmid.assign(el->elmid + e0, el->elmid + e1);
  if (el->beam) o.beam.assign(el->beam + (size_t)FSR_NBEAM * e0, el->beam + (size_t)FSR_NBEAM * e1);
  // result points of the parent before / in / after the block
  int npt = 0;
  for (int e = 0; e < s->nel; ++e) {
    if (e == e0) o.pt0 = npt;
    if (el->elmid && el->elmid[e] < 1) continue;
    if (supported_type(melcon_eff[e])) npt += nstrp_of(melcon_eff[e]);
  }
  if (e0 == s->nel) o.pt0 = npt;
  o.parent_npts = npt;
  auto pad = [](std::vector<int>& v) { if (v.empty()) v.push_back(0); };
  pad(o.mmceq); pad(o.meqn1); pad(o.meqn2); pad(o.mmnpc); pad(o.melcon);
  if (o.ttcc.empty()) o.ttcc.push_back(0.0);
  memset(&o.sam, 0, sizeof(o.sam));
  o.sam.nnod = (int)o.nodes.size(); o.sam.nel = nel_b; o.sam.ndof = ndof_b; o.sam.ndof1 = (int)o.rows1.size(); o.sam.ndof2 = s->ndof2;
  o.sam.ngen = s->ngen; o.sam.neq = neq_b; o.sam.nceq = nceq_b; o.sam.nmmnpc = ip1 - ip0; o.sam.nmmceq = o.mpmceq.back() - 1;
  o.sam.madof = o.madof.data(); o.sam.msc = o.msc.data(); o.sam.mpmnpc = o.mpmnpc.data(); o.sam.mmnpc = o.mmnpc.data();
  o.sam.melcon = o.melcon.data(); o.sam.mpmceq = o.mpmceq.data(); o.sam.mmceq = o.mmceq.data(); o.sam.ttcc = o.ttcc.data();
  o.sam.meqn = o.meqn.data(); o.sam.meqn1 = o.meqn1.data(); o.sam.meqn2 = o.meqn2.data();
  if (o.emod.empty()) { o.emod.push_back(0.0); o.rny.push_back(0.0); o.thk.push_back(0.0); }
  o.elm.xyz = o.xyz.data(); o.elm.emod = o.emod.data(); o.elm.rny = o.rny.data(); o.elm.thk = o.thk.data();
  o.elm.elmid = el->elmid ? (o.elmid.empty() ? nullptr : o.elmid.data()) : nullptr;
  o.elm.beam = el->beam && !o.beam.empty() ? o.beam.data() : nullptr;
  return FSR_OK;
}

static int split_elements(const fsr_sam* sam, const fsr_elmdata* elm, int nblocks, int* e_cut)
{
  const int nel = sam->nel;
  // a quadrilateral that shares a node with any other element type leaves the in-plane path (such nodes keep their six global
  // rows): 38 + 39 f instead of 24 + 26 f (config 4's mixed plate; the same rule as partition.py)
  const double fred = (double)(sam->ndof2 + sam->ngen) / 98.0;
  std::vector<char> mixed_node;
  bool any_quad = false, any_other = false;
  for (int e = 0; e < nel; ++e) {
    const bool q = sam->melcon[e] == 24 || sam->melcon[e] == 22;
    any_quad = any_quad || q; any_other = any_other || !q;
  }
  if (any_quad && any_other && sam->mpmnpc && sam->mmnpc) {
    mixed_node.assign((size_t)sam->nnod + 1, 0);
    for (int e = 0; e < nel; ++e)
      if (sam->melcon[e] != 24 && sam->melcon[e] != 22)
        for (int ip = sam->mpmnpc[e] - 1; ip < sam->mpmnpc[e + 1] - 1; ++ip) {
          const int n = sam->mmnpc[ip];
          if (n >= 1 && n <= sam->nnod) mixed_node[(size_t)n] = 1;
        }
  }
  auto cost_of = [&](int e) {
    const int t = sam->melcon[e];
    if (!mixed_node.empty() && (t == 24 || t == 22))
      for (int ip = sam->mpmnpc[e] - 1; ip < sam->mpmnpc[e + 1] - 1; ++ip) {
        const int n = sam->mmnpc[ip];
        if (n >= 1 && n <= sam->nnod && mixed_node[(size_t)n]) return 38.0 + 39.0 * fred;
      }
    return element_cost(t, fred);
  };
  std::vector<double> cum((size_t)nel + 1, 0.0);
  for (int e = 0; e < nel; ++e)
    cum[(size_t)e + 1] = cum[(size_t)e] + ((elm && elm->elmid && elm->elmid[e] < 1) ? 0.0 : cost_of(e));
  const double total = cum[(size_t)nel];
  e_cut[0] = 0;
  for (int b = 1; b < nblocks; ++b) {
    const double want = total * b / nblocks;
    int e = (int)(std::lower_bound(cum.begin(), cum.end(), want) - cum.begin());   // first e with cum[e] >= want
    e_cut[b] = std::max(e_cut[b - 1], std::min(e, nel));
  }
  e_cut[nblocks] = nel;
  return FSR_OK;
}

}  // namespace fsr

using namespace fsr;

struct fsr_blockdef { BlockArrays a; int e0 = 0, e1 = 0, parent_ndof1 = 0, parent_nel = 0; };

struct fsr_comm {
  ncclComm_t comm = nullptr;
  int rank = 0, world = 1, device = 0;
};

extern "C" {

int fsr_split_elements(const fsr_sam* sam, const fsr_elmdata* elm, int nblocks, int* e_cut)
{
  if (!sam || !sam->melcon || nblocks < 1 || !e_cut) { set_error("fsr_split_elements: bad arguments"); return FSR_ERR_ARG; }
  return split_elements(sam, elm, nblocks, e_cut);
}

// host only: the SAM arrays of an element block, for hosts that keep their own copy (and for the CPU tests)
int fsr_blockdef_create(fsr_blockdef** out, const fsr_sam* sam, const fsr_elmdata* elm, const fsr_options* opt, int e0, int e1)
{
  if (!out || !sam || !elm || !sam->melcon) { set_error("fsr_blockdef_create: null argument"); return FSR_ERR_ARG; }
  *out = nullptr;
  if (e0 < 0 || e1 < e0 || e1 > sam->nel) { set_error("fsr_blockdef_create: element range [%d, %d) outside 0..%d", e0, e1, sam->nel); return FSR_ERR_ARG; }
  std::vector<int> melcon_eff;
  int quad_ngauss = 2;
  effective_element_types(sam, opt, melcon_eff, quad_ngauss);
  fsr_blockdef* d = new fsr_blockdef;
  d->e0 = e0; d->e1 = e1; d->parent_ndof1 = sam->ndof1; d->parent_nel = sam->nel;
  const int rc = make_block(sam, elm, melcon_eff.data(), e0, e1, d->a);
  if (rc) { delete d; return rc; }
  // hand out the caller's element type codes, not the mapped ones (fsr_part_create maps them again)
  d->a.melcon.assign(sam->melcon + e0, sam->melcon + e1);
  if (d->a.melcon.empty()) d->a.melcon.push_back(0);
  d->a.sam.melcon = d->a.melcon.data();
  *out = d;
  return FSR_OK;
}
const fsr_sam* fsr_blockdef_sam(const fsr_blockdef* d) { return d ? &d->a.sam : nullptr; }
const fsr_elmdata* fsr_blockdef_elm(const fsr_blockdef* d) { return d ? &d->a.elm : nullptr; }
int fsr_blockdef_info(const fsr_blockdef* d, int* info, int* rows1, int* nodes)
{
  if (!d) { set_error("fsr_blockdef_info: null handle"); return FSR_ERR_ARG; }
  if (info) {
    info[0] = d->e0; info[1] = d->e1; info[2] = d->a.pt0; info[3] = 0; info[4] = d->a.sam.nnod; info[5] = d->a.sam.ndof1;
    info[6] = d->parent_ndof1; info[7] = d->a.parent_npts; info[8] = d->a.sam.ndof; info[9] = d->parent_nel;
    for (int e = 0; e < d->a.sam.nel; ++e)   // result points of the block (default formulations)
      if (!(d->a.elm.elmid && d->a.elm.elmid[e] < 1)) {
        int t = d->a.melcon[(size_t)e];
        t = t == 21 ? 23 : t == 22 ? 24 : t;
        if (supported_type(t)) info[3] += nstrp_of(t);
      }
  }
  if (rows1) std::copy(d->a.rows1.begin(), d->a.rows1.end(), rows1);
  if (nodes) std::copy(d->a.nodes.begin(), d->a.nodes.end(), nodes);
  return FSR_OK;
}
void fsr_blockdef_destroy(fsr_blockdef* d) { delete d; }

int fsr_part_create_block(fsr_part** out, const fsr_sam* sam, const fsr_elmdata* elm, const fsr_options* opt, int e0, int e1)
{
  if (!out || !sam || !elm) { set_error("fsr_part_create_block: null argument"); return FSR_ERR_ARG; }
  *out = nullptr;
  if (e0 < 0 || e1 < e0 || e1 > sam->nel) { set_error("fsr_part_create_block: element range [%d, %d) outside 0..%d", e0, e1, sam->nel); return FSR_ERR_ARG; }
  if (!sam->madof || !sam->msc || !sam->mpmnpc || !sam->mmnpc || !sam->melcon || !sam->meqn || !elm->xyz || !elm->emod || !elm->rny ||
      !elm->thk || (sam->ndof1 > 0 && !sam->meqn1) || (sam->ndof2 > 0 && !sam->meqn2) ||
      (sam->nceq > 0 && (!sam->mpmceq || !sam->mmceq || !sam->ttcc))) {
    set_error("fsr_part_create_block: incomplete SAM / element data");
    return FSR_ERR_ARG;
  }
  std::vector<int> melcon_eff;
  int quad_ngauss = 2;
  effective_element_types(sam, opt, melcon_eff, quad_ngauss);   // on the PARENT: a block must not change the formulation
  BlockArrays b;
  int rc = make_block(sam, elm, melcon_eff.data(), e0, e1, b);
  if (rc) return rc;
  fsr_part* p = nullptr;
  rc = part_create_mapped(&p, &b.sam, &b.elm, opt, quad_ngauss);
  if (rc < 0) return rc;
  p->is_block = true;
  p->blk_e0 = e0; p->blk_e1 = e1; p->blk_pt0 = b.pt0; p->parent_ndof1 = sam->ndof1; p->parent_nel = sam->nel; p->parent_npts = b.parent_npts;
  p->blk_nodes.swap(b.nodes);
  p->blk_rows1.swap(b.rows1);
  *out = p;
  return rc;
}

int fsr_block_info(const fsr_part* p, int* info)
{
  if (!p || !info) { set_error("fsr_block_info: bad arguments"); return FSR_ERR_ARG; }
  info[0] = p->is_block ? p->blk_e0 : 0;
  info[1] = p->is_block ? p->blk_e1 : p->nel;
  info[2] = p->is_block ? p->blk_pt0 : 0;
  info[3] = p->npts;
  info[4] = p->nnod;
  info[5] = p->ndof1;
  info[6] = p->is_block ? p->parent_ndof1 : p->ndof1;
  info[7] = p->is_block ? p->parent_npts : p->npts;
  info[8] = p->ndof;
  info[9] = p->is_block ? p->parent_nel : p->nel;
  return FSR_OK;
}

int fsr_block_rows(const fsr_part* p, int* rows1, int* nodes)
{
  if (!p) { set_error("fsr_block_rows: null handle"); return FSR_ERR_ARG; }
  if (rows1) { if (p->is_block) std::copy(p->blk_rows1.begin(), p->blk_rows1.end(), rows1); else std::iota(rows1, rows1 + p->ndof1, 0); }
  if (nodes) { if (p->is_block) std::copy(p->blk_nodes.begin(), p->blk_nodes.end(), nodes); else std::iota(nodes, nodes + p->nnod, 1); }
  return FSR_OK;
}

int fsr_set_recovery_parent(fsr_part* p, const double* B, int ldB, const double* E, int ldE)
{
  if (!p) { set_error("fsr_set_recovery_parent: null handle"); return FSR_ERR_ARG; }
  if (!p->is_block) return fsr_set_recovery(p, B, ldB, E, ldE);
  if (p->ndof2 > 0 && p->parent_ndof1 > 0 && (!B || ldB < p->parent_ndof1)) { set_error("fsr_set_recovery_parent: bad B / ldB"); return FSR_ERR_ARG; }
  if (p->ngen > 0 && p->parent_ndof1 > 0 && (!E || ldE < p->parent_ndof1)) { set_error("fsr_set_recovery_parent: bad E / ldE"); return FSR_ERR_ARG; }
  const size_t n1 = (size_t)p->ndof1;
  std::vector<double> Bb(std::max<size_t>(n1 * p->ndof2, 1)), Eb(std::max<size_t>(n1 * p->ngen, 1));
  for (int c = 0; c < p->ndof2; ++c)
    for (size_t k = 0; k < n1; ++k) Bb[(size_t)c * n1 + k] = B[(size_t)c * ldB + p->blk_rows1[k]];
  for (int c = 0; c < p->ngen; ++c)
    for (size_t k = 0; k < n1; ++k) Eb[(size_t)c * n1 + k] = E[(size_t)c * ldE + p->blk_rows1[k]];
  return fsr_set_recovery(p, p->ndof2 > 0 ? Bb.data() : nullptr, (int)n1, p->ngen > 0 ? Eb.data() : nullptr, (int)n1);
}

// ---- group ---------------------------------------------------------------------------------------------
void fsr_group_destroy(fsr_group* g)
{
  if (!g) return;
  NcclApi* api = g->comms.empty() ? nullptr : nccl_api();
  for (int b = 0; b < g->nblk; ++b) {
    cudaSetDevice(g->devices[(size_t)b]);
    if (b < (int)g->streams.size() && g->streams[(size_t)b]) cudaStreamSynchronize(g->streams[(size_t)b]);
    if (api && b < (int)g->comms.size() && g->comms[(size_t)b]) api->CommDestroy(g->comms[(size_t)b]);
    if (b < (int)g->Qdev.size()) cudaFree(g->Qdev[(size_t)b]);
    if (b < (int)g->env_blk.size()) cudaFree(g->env_blk[(size_t)b]);
    if (b < (int)g->vm_blk.size()) cudaFree(g->vm_blk[(size_t)b]);
    if (b == 0) { if (g->q_ev) cudaEventDestroy(g->q_ev); cudaFree(g->env_root); if (g->env_pin) cudaFreeHost(g->env_pin); if (g->Qpin) cudaFreeHost(g->Qpin); if (g->vm_pin) cudaFreeHost(g->vm_pin); }
    if (b < (int)g->parts.size()) fsr_part_destroy(g->parts[(size_t)b]);
  }
  delete g;
}

int fsr_group_create(fsr_group** out, const fsr_sam* sam, const fsr_elmdata* elm, const fsr_options* opt, const int* devices, int ndev)
{
  if (!out || !sam || !elm) { set_error("fsr_group_create: null argument"); return FSR_ERR_ARG; }
  *out = nullptr;
  int nvis = 0;
  if (cudaGetDeviceCount(&nvis) != cudaSuccess || nvis < 1) { set_error("no CUDA device available: this library has no CPU fallback"); return FSR_ERR_CUDA; }
  std::vector<int> devs;
  if (ndev <= 0 || !devices) { for (int d = 0; d < (ndev > 0 ? std::min(ndev, nvis) : nvis); ++d) devs.push_back(d); }
  else devs.assign(devices, devices + ndev);
  for (int d : devs) if (d < 0 || d >= nvis) { set_error("fsr_group_create: device %d out of range (0..%d)", d, nvis - 1); return FSR_ERR_ARG; }
  fsr_group* g = new fsr_group;
  g->nblk = (int)devs.size();
  g->devices = devs;
  g->ndim = sam->ndof2 + sam->ngen;
  g->nel = sam->nel;
  g->nnod = sam->nnod;
  if (!sam->madof || !sam->melcon) { delete g; set_error("fsr_group_create: incomplete SAM data"); return FSR_ERR_ARG; }
  g->madof_host.assign(sam->madof, sam->madof + sam->nnod + 1);
  {
    int qg = 2;
    effective_element_types(sam, opt, g->melcon_host, qg);
    g->active_host.resize((size_t)sam->nel);
    for (int e = 0; e < sam->nel; ++e) g->active_host[(size_t)e] = (elm->elmid && elm->elmid[e] < 1) ? 0 : 1;
  }
  g->e_cut.resize((size_t)g->nblk + 1);
  split_elements(sam, elm, g->nblk, g->e_cut.data());
  int nfail = 0;
  for (int b = 0; b < g->nblk; ++b) {
    fsr_options o;
    memset(&o, 0, sizeof(o));
    if (opt) o = *opt;
    o.device = devs[(size_t)b];
    fsr_part* p = nullptr;
    const int rc = fsr_part_create_block(&p, sam, elm, &o, g->e_cut[(size_t)b], g->e_cut[(size_t)b + 1]);
    if (rc < 0) { fsr_group_destroy(g); return rc; }
    nfail += rc;
    g->parts.push_back(p);
    g->streams.push_back(p->stream);
    g->npts = p->parent_npts;
  }
  g->Qdev.assign((size_t)g->nblk, nullptr);
  g->env_blk.assign((size_t)g->nblk, nullptr);
  g->vm_blk.assign((size_t)g->nblk, nullptr);
  if (g->nblk > 1) {
    NcclApi* api = nccl_api();
    if (!api) { fsr_group_destroy(g); return FSR_ERR_CUDA; }
    g->comms.assign((size_t)g->nblk, nullptr);
    ncclResult_t r = api->CommInitAll(g->comms.data(), g->nblk, devs.data());
    if (r != ncclSuccess) { set_error("ncclCommInitAll failed: %s", api->GetErrorString(r)); g->comms.clear(); fsr_group_destroy(g); return FSR_ERR_CUDA; }
  }
  for (int b = 0; b < g->nblk; ++b) {
    cudaSetDevice(devs[(size_t)b]);
    if (cudaMalloc(&g->env_blk[(size_t)b], sizeof(double) * 2 * std::max(g->parts[(size_t)b]->npts, 1)) != cudaSuccess) { set_error("fsr_group_create: device allocation failed"); fsr_group_destroy(g); return FSR_ERR_ALLOC; }
  }
  cudaSetDevice(devs[0]);
  if (cudaMalloc(&g->env_root, sizeof(double) * 2 * std::max(g->npts, 1)) != cudaSuccess ||
      cudaMallocHost(&g->env_pin, sizeof(double) * 2 * std::max(g->npts, 1)) != cudaSuccess) { set_error("fsr_group_create: allocation of the envelope buffers failed"); fsr_group_destroy(g); return FSR_ERR_ALLOC; }
  *out = g;
  return nfail;
}

int fsr_group_num_blocks(const fsr_group* g) { return g ? g->nblk : FSR_ERR_ARG; }
int fsr_group_num_result_points(const fsr_group* g) { return g ? g->npts : FSR_ERR_ARG; }
int fsr_group_ndim(const fsr_group* g) { return g ? g->ndim : FSR_ERR_ARG; }
fsr_part* fsr_group_block(fsr_group* g, int b) { return g && b >= 0 && b < g->nblk ? g->parts[(size_t)b] : nullptr; }

int fsr_group_set_recovery(fsr_group* g, const double* B, int ldB, const double* E, int ldE)
{
  if (!g) { set_error("fsr_group_set_recovery: null handle"); return FSR_ERR_ARG; }
  for (fsr_part* p : g->parts) {
    const int rc = fsr_set_recovery_parent(p, B, ldB, E, ldE);
    if (rc < 0) return rc;
  }
  return FSR_OK;
}

int fsr_group_reset_envelope(fsr_group* g)
{
  if (!g) return FSR_ERR_ARG;
  for (fsr_part* p : g->parts) { const int rc = fsr_reset_envelope(p); if (rc < 0) return rc; }
  return FSR_OK;
}

// Q (host) -> device of block 0 -> ncclBroadcast to every block; K1 + K2 of all blocks run concurrently on their GPUs.
// vm_hist: optional [nsteps x npts] step-major history in the PARENT's result-point order.
int fsr_group_recover(fsr_group* g, const double* Q, int ldq, int nsteps, double* vm_hist)
{
  if (!g || !Q || nsteps < 0 || ldq < g->ndim) { set_error("fsr_group_recover: bad arguments"); return FSR_ERR_ARG; }
  if (nsteps == 0) return FSR_OK;
  const size_t nq = (size_t)ldq * nsteps;
  if (g->q_cap < nq) {
    for (int b = 0; b < g->nblk; ++b) {
      FSR_CUDA(cudaSetDevice(g->devices[(size_t)b]));
      FSR_CUDA(cudaStreamSynchronize(g->streams[(size_t)b]));
      cudaFree(g->Qdev[(size_t)b]); g->Qdev[(size_t)b] = nullptr;
      FSR_CUDA(cudaMalloc(&g->Qdev[(size_t)b], sizeof(double) * nq));
    }
    g->q_cap = nq;
  }
  FSR_CUDA(cudaSetDevice(g->devices[0]));
  if (!g->q_ev) FSR_CUDA(cudaEventCreateWithFlags(&g->q_ev, cudaEventDisableTiming));
  FSR_CUDA(cudaEventSynchronize(g->q_ev));   // the previous window's H2D out of Qpin is done (the device may still compute)
  if (g->qpin_cap < nq) {
    if (g->Qpin) cudaFreeHost(g->Qpin);
    g->Qpin = nullptr; g->qpin_cap = 0;
    FSR_CUDA(cudaMallocHost(&g->Qpin, sizeof(double) * nq));
    g->qpin_cap = nq;
  }
  memcpy(g->Qpin, Q, sizeof(double) * nq);
  FSR_CUDA(cudaMemcpyAsync(g->Qdev[0], g->Qpin, sizeof(double) * nq, cudaMemcpyHostToDevice, g->streams[0]));
  FSR_CUDA(cudaEventRecord(g->q_ev, g->streams[0]));
  if (g->nblk > 1) {
    NcclApi* api = nccl_api();
    if (!api) return FSR_ERR_CUDA;
    FSR_NCCL(api, api->GroupStart());
    for (int b = 0; b < g->nblk; ++b)
      FSR_NCCL(api, api->Broadcast(g->Qdev[0], g->Qdev[(size_t)b], nq, ncclDouble, 0, g->comms[(size_t)b], g->streams[(size_t)b]));
    FSR_NCCL(api, api->GroupEnd());
  }
  if (!vm_hist) {
    for (int b = 0; b < g->nblk; ++b) {
      const int rc = fsr_recover_dev(g->parts[(size_t)b], g->Qdev[(size_t)b], ldq, nsteps, nullptr, 0, nullptr);
      if (rc < 0) return rc;
    }
    return FSR_OK;
  }
  // with history: tiles of steps, every block writes its columns of the parent's [step][point] rows
  int tile = 1 << 30;
  for (fsr_part* p : g->parts) tile = std::min(tile, p->step_tile);
  size_t maxpts = 1;
  for (fsr_part* p : g->parts) maxpts = std::max(maxpts, (size_t)p->npts);
  if (g->vm_cap < (size_t)tile * maxpts) {
    for (int b = 0; b < g->nblk; ++b) {
      FSR_CUDA(cudaSetDevice(g->devices[(size_t)b]));
      cudaFree(g->vm_blk[(size_t)b]); g->vm_blk[(size_t)b] = nullptr;
      FSR_CUDA(cudaMalloc(&g->vm_blk[(size_t)b], sizeof(double) * (size_t)tile * maxpts));
    }
    g->vm_cap = (size_t)tile * maxpts;
  }
  for (int t0 = 0; t0 < nsteps; t0 += tile) {
    const int nt = std::min(tile, nsteps - t0);
    for (int b = 0; b < g->nblk; ++b) {
      fsr_part* p = g->parts[(size_t)b];
      if (p->npts == 0) { const int rc = fsr_recover_dev(p, g->Qdev[(size_t)b] + (size_t)t0 * ldq, ldq, nt, nullptr, 0, nullptr); if (rc < 0) return rc; continue; }
      const int rc = fsr_recover_dev(p, g->Qdev[(size_t)b] + (size_t)t0 * ldq, ldq, nt, g->vm_blk[(size_t)b], (size_t)p->npts, nullptr);
      if (rc < 0) return rc;
      FSR_CUDA(cudaSetDevice(g->devices[(size_t)b]));
      FSR_CUDA(cudaMemcpy2DAsync(vm_hist + (size_t)t0 * g->npts + p->blk_pt0, sizeof(double) * (size_t)g->npts, g->vm_blk[(size_t)b],
                                 sizeof(double) * (size_t)p->npts, sizeof(double) * (size_t)p->npts, (size_t)nt, cudaMemcpyDeviceToHost,
                                 g->streams[(size_t)b]));
    }
    for (int b = 0; b < g->nblk; ++b) {   // vm_blk is reused by the next tile
      FSR_CUDA(cudaSetDevice(g->devices[(size_t)b]));
      FSR_CUDA(cudaStreamSynchronize(g->streams[(size_t)b]));
    }
  }
  return FSR_OK;
}

int fsr_group_synchronize(fsr_group* g)
{
  if (!g) return FSR_ERR_ARG;
  for (int b = 0; b < g->nblk; ++b) {
    FSR_CUDA(cudaSetDevice(g->devices[(size_t)b]));
    FSR_CUDA(cudaStreamSynchronize(g->streams[(size_t)b]));
  }
  return FSR_OK;
}

// per-block envelopes -> device of block 0 with ncclSend / ncclRecv (blocks are contiguous element ranges, so block b
// lands at its pt0 and the concatenation IS the parent's result-point order) -> host
int fsr_group_get_envelope(fsr_group* g, double* vm_max, double* vm_min)
{
  if (!g) { set_error("fsr_group_get_envelope: null handle"); return FSR_ERR_ARG; }
  const size_t np = (size_t)g->npts;
  for (int b = 0; b < g->nblk; ++b) {
    fsr_part* p = g->parts[(size_t)b];
    const int rc = fsr_copy_envelope_dev(p, g->env_blk[(size_t)b], g->env_blk[(size_t)b] + p->npts, nullptr);
    if (rc < 0) return rc;
  }
  if (g->nblk > 1) {
    NcclApi* api = nccl_api();
    if (!api) return FSR_ERR_CUDA;
    FSR_NCCL(api, api->GroupStart());
    for (int b = 1; b < g->nblk; ++b) {
      fsr_part* p = g->parts[(size_t)b];
      if (p->npts == 0) continue;
      FSR_NCCL(api, api->Send(g->env_blk[(size_t)b], (size_t)p->npts, ncclDouble, 0, g->comms[(size_t)b], g->streams[(size_t)b]));
      FSR_NCCL(api, api->Send(g->env_blk[(size_t)b] + p->npts, (size_t)p->npts, ncclDouble, 0, g->comms[(size_t)b], g->streams[(size_t)b]));
      FSR_NCCL(api, api->Recv(g->env_root + p->blk_pt0, (size_t)p->npts, ncclDouble, b, g->comms[0], g->streams[0]));
      FSR_NCCL(api, api->Recv(g->env_root + np + p->blk_pt0, (size_t)p->npts, ncclDouble, b, g->comms[0], g->streams[0]));
    }
    FSR_NCCL(api, api->GroupEnd());
  }
  FSR_CUDA(cudaSetDevice(g->devices[0]));
  fsr_part* p0 = g->parts[0];
  if (p0->npts > 0) {
    FSR_CUDA(cudaMemcpyAsync(g->env_root + p0->blk_pt0, g->env_blk[0], sizeof(double) * p0->npts, cudaMemcpyDeviceToDevice, g->streams[0]));
    FSR_CUDA(cudaMemcpyAsync(g->env_root + np + p0->blk_pt0, g->env_blk[0] + p0->npts, sizeof(double) * p0->npts, cudaMemcpyDeviceToDevice, g->streams[0]));
  }
  FSR_CUDA(cudaMemcpyAsync(g->env_pin, g->env_root, sizeof(double) * 2 * np, cudaMemcpyDeviceToHost, g->streams[0]));
  FSR_CUDA(cudaStreamSynchronize(g->streams[0]));
  for (int b = 1; b < g->nblk; ++b) {
    FSR_CUDA(cudaSetDevice(g->devices[(size_t)b]));
    FSR_CUDA(cudaStreamSynchronize(g->streams[(size_t)b]));
  }
  if (vm_max) memcpy(vm_max, g->env_pin, sizeof(double) * np);
  if (vm_min) memcpy(vm_min, g->env_pin + np, sizeof(double) * np);
  return FSR_OK;
}

// device time of the slowest block since the last fsr_group_timing_reset: t[0] = K1, t[1] = K2 (ms), t[2] = tiles,
// t[3] = fastest / slowest block (load balance)
int fsr_group_last_timing(fsr_group* g, double* t, int n)
{
  if (!g || !t) return FSR_ERR_ARG;
  double worst[3] = {0, 0, 0}, lo = 1e300, hi = 0.0;
  for (fsr_part* p : g->parts) {
    double v[3] = {0, 0, 0};
    const int rc = fsr_last_timing(p, v, 3);
    if (rc < 0) return rc;
    const double s = v[0] + v[1];
    if (s > hi) { hi = s; worst[0] = v[0]; worst[1] = v[1]; worst[2] = v[2]; }
    lo = std::min(lo, s);
  }
  const double v[4] = {worst[0], worst[1], worst[2], hi > 0.0 ? lo / hi : 1.0};
  const int m = std::min(n, 4);
  for (int i = 0; i < m; ++i) t[i] = v[i];
  return m;
}

int fsr_group_timing_reset(fsr_group* g)
{
  if (!g) return FSR_ERR_ARG;
  for (fsr_part* p : g->parts) fsr_timing_reset(p);
  return FSR_OK;
}

// ---- one process per GPU -----------------------------------------------------------------------------------
int fsr_comm_unique_id(char* id, int cap)
{
  if (!id || cap < (int)sizeof(ncclUniqueId)) { set_error("fsr_comm_unique_id: the id buffer must hold %d bytes", (int)sizeof(ncclUniqueId)); return FSR_ERR_ARG; }
  NcclApi* api = nccl_api();
  if (!api) return FSR_ERR_CUDA;
  ncclUniqueId u;
  FSR_NCCL(api, api->GetUniqueId(&u));
  memcpy(id, &u, sizeof(u));
  return (int)sizeof(u);
}

int fsr_comm_init_rank(fsr_comm** out, const char* id, int rank, int world, int device)
{
  if (!out || !id || world < 1 || rank < 0 || rank >= world) { set_error("fsr_comm_init_rank: bad arguments"); return FSR_ERR_ARG; }
  *out = nullptr;
  NcclApi* api = nccl_api();
  if (!api) return FSR_ERR_CUDA;
  FSR_CUDA(cudaSetDevice(device));
  ncclUniqueId u;
  memcpy(&u, id, sizeof(u));
  fsr_comm* c = new fsr_comm;
  c->rank = rank; c->world = world; c->device = device;
  ncclResult_t r = api->CommInitRank(&c->comm, world, u, rank);
  if (r != ncclSuccess) { set_error("ncclCommInitRank failed: %s", api->GetErrorString(r)); delete c; return FSR_ERR_CUDA; }
  *out = c;
  return FSR_OK;
}

void fsr_comm_destroy(fsr_comm* c)
{
  if (!c) return;
  NcclApi* api = nccl_api();
  cudaSetDevice(c->device);
  if (api && c->comm) api->CommDestroy(c->comm);
  delete c;
}

// the reduced history of a window of steps, from the root's device buffer to everybody's (in place on the root)
int fsr_comm_broadcast(fsr_comm* c, double* buf_dev, long long count, int root, void* stream)
{
  if (!c || !buf_dev || count < 0) { set_error("fsr_comm_broadcast: bad arguments"); return FSR_ERR_ARG; }
  NcclApi* api = nccl_api();
  if (!api) return FSR_ERR_CUDA;
  FSR_CUDA(cudaSetDevice(c->device));
  FSR_NCCL(api, api->Broadcast(buf_dev, buf_dev, (size_t)count, ncclDouble, root, c->comm, (cudaStream_t)stream));
  return FSR_OK;
}

// Envelopes of this rank's element block to the root: the root passes device buffers [parent npts] each and receives
// block r at pt0[r] (pt0 / npts per rank as fsr_block_info reports them, identical arrays on all ranks); the others send.
int fsr_comm_gather_envelope(fsr_comm* c, fsr_part* block, const int* pt0, const int* npts, double* vm_max_root_dev,
                             double* vm_min_root_dev, int root, void* stream)
{
  if (!c || !block || !pt0 || !npts) { set_error("fsr_comm_gather_envelope: bad arguments"); return FSR_ERR_ARG; }
  NcclApi* api = nccl_api();
  if (!api) return FSR_ERR_CUDA;
  FSR_CUDA(cudaSetDevice(c->device));
  cudaStream_t s = (cudaStream_t)stream;
  double *emax = nullptr, *emin = nullptr;
  fsr_envelope_dev(block, &emax, &emin);
  if (c->rank == root && (!vm_max_root_dev || !vm_min_root_dev)) { set_error("fsr_comm_gather_envelope: the root needs the two destination buffers"); return FSR_ERR_ARG; }
  FSR_NCCL(api, api->GroupStart());
  if (c->rank != root) {
    if (block->npts > 0) {
      FSR_NCCL(api, api->Send(emax, (size_t)block->npts, ncclDouble, root, c->comm, s));
      FSR_NCCL(api, api->Send(emin, (size_t)block->npts, ncclDouble, root, c->comm, s));
    }
  } else {
    for (int r = 0; r < c->world; ++r) {
      if (r == root || npts[r] <= 0) continue;
      FSR_NCCL(api, api->Recv(vm_max_root_dev + pt0[r], (size_t)npts[r], ncclDouble, r, c->comm, s));
      FSR_NCCL(api, api->Recv(vm_min_root_dev + pt0[r], (size_t)npts[r], ncclDouble, r, c->comm, s));
    }
  }
  FSR_NCCL(api, api->GroupEnd());
  if (c->rank == root && block->npts > 0) {
    FSR_CUDA(cudaMemcpyAsync(vm_max_root_dev + pt0[root], emax, sizeof(double) * block->npts, cudaMemcpyDeviceToDevice, s));
    FSR_CUDA(cudaMemcpyAsync(vm_min_root_dev + pt0[root], emin, sizeof(double) * block->npts, cudaMemcpyDeviceToDevice, s));
  }
  return FSR_OK;
}

int fsr_nccl_version(void)
{
  NcclApi* api = nccl_api();
  if (!api) return FSR_ERR_CUDA;
  int v = 0;
  api->GetVersion(&v);
  return v;
}

}  // extern "C"
