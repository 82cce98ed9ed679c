// common.cuh -- shared declarations of the B200 stress-recovery library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include <nccl.h>   // types only: the library binds NCCL at run time (sharded.cu)

#include "../../include/fedem_b200.h"

namespace fsr {

void set_error(const char* fmt, ...);
extern long long g_launches;
// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device (primary context) attribute: set it once per
// (current device, kernel), thread-safe (api.cu).  Returns FSR_OK or FSR_ERR_CUDA with the error text set.
int smem_opt_in(const void* kernel, size_t bytes);

#define FSR_CUDA(call)                                                              \
  do {                                                                              \
    cudaError_t e_ = (call);                                                        \
    if (e_ != cudaSuccess) {                                                        \
      fsr::set_error("%s failed at %s:%d: %s", #call, __FILE__, __LINE__,           \
                     cudaGetErrorString(e_));                                       \
      return FSR_ERR_CUDA;                                                          \
    }                                                                               \
  } while (0)

#define FSR_LAUNCH_CHECK()                                                          \
  do {                                                                              \
    ++fsr::g_launches;                                                              \
    cudaError_t e_ = cudaGetLastError();                                            \
    if (e_ != cudaSuccess) {                                                        \
      fsr::set_error("kernel launch failed at %s:%d: %s", __FILE__, __LINE__,       \
                     cudaGetErrorString(e_));                                       \
      return FSR_ERR_CUDA;                                                          \
    }                                                                               \
  } while (0)

constexpr double kHuge = 1.7976931348623157e308;  // hugeVal_p = huge(1.0_dp)
constexpr double kEpsDiv0 = 2.220446049250313e-16; // epsDiv0_p = epsilon(1.0_dp)

// One FP64 tensor-core MMA: D(8x8) += A(8x4, row) * B(4x8, col).  SASS: DMMA.8x8x4.
// Fragment ownership (g = lane>>2, t = lane&3): a = A[g][t], b = B[t][g], c = C[g][2t..2t+1].
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b)
{
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// sqrt(x) for x >= 0 to < 1 ulp-ish (two Newton steps on the 2^-22 hardware seed); 0 for x < 1e-290
__device__ __forceinline__ double sqrt_pos(double x)
{
  double r;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double y = x * r, h = 0.5 * r;
  double e = fma(-h, y, 0.5);
  y = fma(y, e, y);
  h = fma(h, e, h);
  e = fma(-y, y, x);
  y = fma(e, h, y);
  return x > 1.0e-290 ? y : 0.0;
}

// max / min of NON-NEGATIVE doubles (von Mises values, their radicands) on the integer pipe: IEEE order = integer order
// there, and the FP64 pipe is the busy one in every K2 kernel
__device__ __forceinline__ double max_nonneg(double a, double b)
{
  return __longlong_as_double(max(__double_as_longlong(a), __double_as_longlong(b)));
}
__device__ __forceinline__ double min_nonneg(double a, double b)
{
  return __longlong_as_double(min(__double_as_longlong(a), __double_as_longlong(b)));
}

// Element families handled by the K2 kernels.  A family fixes the operator shape:
// MT m-tiles of 8 rows, KT k-tiles of 4 element DOFs.
enum Family { FAM_QUAD = 0, FAM_TRI = 1, FAM_TET10 = 2, FAM_BEAM = 3, FAM_HEX20 = 4, FAM_HEX8 = 5, FAM_TET4 = 6, FAM_WEDG6 = 7, FAM_WEDG15 = 8, FAM_TRI6 = 9, FAM_QUAD8 = 10, FAM_COUNT = 11 };

struct FamilyData {
  int nelt = 0;          // elements of this family (active only)
  int nenod = 0, nndof = 0, nstrp = 0, ncmp = 0;
  int MT = 0, KT = 0;    // operator tiles
  int* elem = nullptr;   // [nelt] 0-based element index in SAM order
  int* edof = nullptr;   // [nelt][KT*4] 0-based row of U for each element DOF (padding -> 0)
  int* ptoff = nullptr;  // [nelt] first result point (or element slot for beams)
  double* Sfrag = nullptr;      // [nelt][MT][KT][32] operator in DMMA A-fragment order
  unsigned char* failed = nullptr; // [nelt] 1 = operator build failed -> hugeVal results
  double* aux = nullptr; // per-element scalars needed by the full-output kernels
  int naux = 0;
  double* Gfrag = nullptr;      // solids: displacement-gradient operator in A-fragment order (k2_solid.cu)
  double* Efrag = nullptr;      // thick shells: strain operator, same shape as Sfrag (strain is not an isotropic function of
                                // the global stress there, k2_thickshell.cu); NULL for every other family
  // geometry fast path of a family (TET10: straight-sided elements, constant Jacobian): the family elements are split
  // into sub[0] (fast path) and sub[1] (general kernel), indices into the family arrays; fast[] = per-element constants
  int* sub[3] = {nullptr, nullptr, nullptr};   // TET10: [2] = curved elements on the scalar kernel
  int nsub[3] = {0, 0, 0};
  double* fast = nullptr;
  double* fast2 = nullptr;      // TET10: J^-1 of the ten nodal evaluation points; quads: operators of the in-plane form
  int* edof2 = nullptr;         // quads, in-plane form: [nsub[2]][16] rows of Up (u, v, theta1, theta2 of the four nodes)
};

}  // namespace fsr

namespace fsr {
// Host copies of the SAM index maps that fsr_set_recovery needs after fsr_part_create returned
// (the caller may free its arrays in between).
struct SamKeep {
  bool valid = false;
  int nnod = 0, nel = 0, ndof = 0, ndof1 = 0, ndof2 = 0, ngen = 0, neq = 0, nceq = 0;
  std::vector<int> msc, meqn, meqn1, meqn2, mpmceq, mmceq;
  std::vector<double> ttcc;
  void keep(const fsr_sam* s)
  {
    nnod = s->nnod; nel = s->nel; ndof = s->ndof; ndof1 = s->ndof1; ndof2 = s->ndof2;
    ngen = s->ngen; neq = s->neq; nceq = s->nceq;
    msc.assign(s->msc, s->msc + ndof);
    meqn.assign(s->meqn, s->meqn + ndof);
    if (ndof1 > 0) meqn1.assign(s->meqn1, s->meqn1 + ndof1);
    if (ndof2 > 0) meqn2.assign(s->meqn2, s->meqn2 + ndof2);
    if (nceq > 0) {
      mpmceq.assign(s->mpmceq, s->mpmceq + nceq + 1);
      int nm = s->nmmceq > 0 ? s->nmmceq : mpmceq[nceq] - 1;
      mmceq.assign(s->mmceq, s->mmceq + nm);
      ttcc.assign(s->ttcc, s->ttcc + nm);
    }
    valid = true;
  }
  fsr_sam view() const
  {
    fsr_sam v;
    memset(&v, 0, sizeof(v));
    v.nnod = nnod; v.nel = nel; v.ndof = ndof; v.ndof1 = ndof1; v.ndof2 = ndof2; v.ngen = ngen;
    v.neq = neq; v.nceq = nceq; v.nmmceq = (int)mmceq.size();
    v.msc = msc.data(); v.meqn = meqn.data(); v.meqn1 = meqn1.data(); v.meqn2 = meqn2.data();
    v.mpmceq = mpmceq.data(); v.mmceq = mmceq.data(); v.ttcc = ttcc.data();
    return v;
  }
};
}  // namespace fsr

struct fsr_part {
  int device = 0;
  int nnod = 0, nel = 0, ndof = 0, ndof1 = 0, ndof2 = 0, ngen = 0, neq = 0, nceq = 0, ndim = 0;
  int ldk = 0;         // padded reduced dimension (multiple of 4, == 4 mod 8: conflict-free smem)
  int nrows_pad = 0;   // ndof padded to the K1 row tile
  int npts = 0;        // result points
  int stressForm = 0;
  int tri_legacy = 0;   // 1 = the triangles are legacy FFT3 shells recovered with -fftStressForm 0 / 2 (FTS31 / FTS32 instead of FTSA31 / FTSA32)
  int quad_ngauss = 2;  // Gauss points per direction of the quad shell stress evaluation (1 only for legacy FFQ with -ffqStressForm 1)
  int elem_order = 0;  // 0 = elements processed in Morton order of their centroids (L2 reuse of shared
                       // nodes), 1 = SAM order.  Outputs are always in SAM order.
  int step_tile = 0;   // steps per device batch
  std::vector<int> ptoff_host;  // [nel+1]
  std::vector<int> melcon_host;
  std::vector<int> madof_host;      // [nnod+1] (gage setup, in-core vms layout)
  std::vector<int> nenod_host;      // [nel] nodes per element
  std::vector<int> active_host;     // [nel] 1 = element takes part (elmid >= 1 or no elmid given)
  std::vector<double> xyz_host;     // [3*nnod] (gage setup)
  // external-DOF sources of every nodal DOF row of R: row d receives w * finit(extcol[j]) for each
  // (j, w) in [ext_rowptr[d], ext_rowptr[d+1]); kept for the gage operator (k3_gage.cu)
  std::vector<int> ext_rowptr, ext_j, extcol;
  std::vector<double> ext_w;
  fsr::SamKeep sam_keep;
  // device model data
  double* xyz = nullptr;    // [3*nnod]
  double* emod = nullptr;   // [nel]
  double* rny = nullptr;    // [nel]
  double* thk = nullptr;    // [nel]
  // recovery operator
  double* R = nullptr;      // [nrows_pad][ldk] row-major, zero padded
  bool have_R = false;
  // per-batch buffers
  double* Qt = nullptr;     // [step_tile][ldk]
  double* U = nullptr;      // [nrows_pad][step_tile]  (row = nodal DOF, t fastest)
  // Flat shell regions (k2_shell.cu, "in-plane form"): the nodes whose flat quadrilaterals all lie in one plane get four rows
  // (u, v, theta1, theta2) in the axes of that plane instead of six global ones: Rp = W . R, Up = Rp . Q.  The von Mises
  // path expands only these rows plus the 128-row tiles of R that some other element still reads (k1_tiles).
  bool planar = false;
  int np_rows = 0, np_rows_pad = 0;   // rows of Rp / Up
  int* prow_src = nullptr;            // [np_rows][3] rows of R / U feeding a row of Rp / Up
  double* prow_w = nullptr;           // [np_rows][3] their weights (a unit vector of the plane)
  double* Rp = nullptr;               // [np_rows_pad][ldk]
  double* Up = nullptr;               // [np_rows_pad][step_tile]
  int* k1_tiles = nullptr;            // row tiles of R the von Mises path still needs
  int n_k1_tiles = 0;
  double* vm_tile = nullptr;// [step_tile][npts] staging for host output
  double* Qstage = nullptr; // device copy of the caller's Q (host API)
  size_t Qstage_cap = 0;
  double* env_max = nullptr;
  double* env_min = nullptr;
  fsr::FamilyData fam[fsr::FAM_COUNT];
  int nfailed = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = true;
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t evring[256][3] = {};  // per-tile event triplets: start, after K1, after K2
  int ntimed = 0;
  double* pinned = nullptr; size_t pinned_cap = 0;
  // asynchronous envelope read-back (fsr_get_envelope_async): snapshot of the envelopes, copy stream, copy-finished event
  double* env_snap = nullptr;
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t ev_snap = nullptr, ev_copied = nullptr;
  // element block of a larger part (sharded.cu): the parent's element range, the block's first result point in the parent's
  // result-point order, the parent's nodes (1-based) of the block's nodes and the rows (0-based) of the parent's B / E it keeps
  bool is_block = false;
  int blk_e0 = 0, blk_e1 = 0, blk_pt0 = 0, parent_ndof1 = 0, parent_nel = 0, parent_npts = 0;
  std::vector<int> blk_nodes, blk_rows1;
};

// ---- one process, several GPUs ------------------------------------------------------------------------
struct fsr_group {
  int nblk = 0, ndim = 0, npts = 0, nel = 0, nnod = 0;
  std::vector<int> madof_host, melcon_host, active_host;   // the parent part (results database header)
  std::vector<fsr_part*> parts;
  std::vector<int> devices, e_cut;
  std::vector<ncclComm_t> comms;       // empty when nblk == 1
  std::vector<cudaStream_t> streams;   // the blocks' own streams
  std::vector<double*> Qdev;           // per device: the broadcast Q window
  size_t q_cap = 0;
  double* Qpin = nullptr;              // pinned staging of the caller's Q
  size_t qpin_cap = 0;
  cudaEvent_t q_ev = nullptr;          // the H2D copy out of Qpin has finished
  double* env_root = nullptr;          // device 0: [2][npts] gathered envelopes in the parent's result-point order
  double* env_pin = nullptr;           // pinned copy
  std::vector<double*> env_blk;        // per device: [2][npts_b] contiguous copy of the block's envelopes (send buffer)
  std::vector<double*> vm_blk;         // per device: von Mises history staging [tile][npts_b]
  size_t vm_cap = 0;
  double* vm_pin = nullptr;            // pinned [nblk slices] for the host history
  size_t vm_pin_cap = 0;
};


namespace fsr {
// k1_expand.cu
int build_row_operator(fsr_part* p, const fsr_sam* sam, const double* B, int ldB, const double* E,
                       int ldE);
int launch_pack_q(fsr_part* p, const double* Q_dev, int ldq, int nsteps, int nsteps_pad,
                  cudaStream_t s);
int launch_k1(fsr_part* p, int nsteps_pad, cudaStream_t s);
// the expansion of the von Mises path: with in-plane rows (fsr_part::planar) Up = Rp . Q and only the row tiles of U that are
// still read; full_u = true expands all of U as well (the solver step hands the nodal displacements out)
int launch_k1_vm(fsr_part* p, int nsteps_pad, cudaStream_t s, bool full_u);
int build_planar_rows(fsr_part* p);                                   // Rp = W . R (after build_row_operator)
int planar_rows_from_u(fsr_part* p, int nsteps_pad, cudaStream_t s);  // Up = W . U (displacements given, no expansion)
// the same GEMM on any row-major operator: U[nrows_pad x ldu] = R[nrows_pad x ldk] . Qt[nsteps_pad x ldk]^T
int launch_k1_raw(const double* R, const double* Qt, double* U, int ldk, int nrows_pad, int nsteps_pad, size_t ldu,
                  cudaStream_t s, const int* tiles = nullptr, int ntiles = 0);
int launch_pack_q_raw(double* Qt, int ldk, const double* Q_dev, int ldq, int ndim, int nsteps, int nsteps_pad,
                      cudaStream_t s);
// api.cu: result points per element type, types with a stress operator, the legacy shell type mapping of fsr_part_create
int nstrp_of(int type);
bool supported_type(int type);
void effective_element_types(const fsr_sam* sam, const fsr_options* opt, std::vector<int>& melcon_eff, int& quad_ngauss);
int part_create_mapped(fsr_part** out, const fsr_sam* sam, const fsr_elmdata* elm, const fsr_options* opt, int quad_ngauss);
// api.cu: nodal displacements of nt steps (host, [nt][ndof]) -> U[dof][t] (direct solution on the results files)
int upload_displacements(fsr_part* p, const double* sv_host, int nt, cudaStream_t s);
// api.cu: one solver step (solver_state.cu)
int step_enqueue(fsr_part* p, const double* q, double* sv_host, double* vm_host);
// api.cu: active elements of one type, in processing order (see fsr_part::elem_order)
std::vector<int> elements_of_type(const fsr_part* p, const fsr_sam* sam, const fsr_elmdata* elm, int type);
// io_rdb.cu: the meta data lines of a results database header (openHeaderFiles)
std::string rdb_file_preamble(const char* module, const char* model_file, const char* link_file, const char* info = nullptr);
// k2_*.cu
int build_shell_operators(fsr_part* p, const fsr_sam* sam, const fsr_elmdata* elm);
int build_solid_operators(fsr_part* p, const fsr_sam* sam, const fsr_elmdata* elm);
int build_beam_operators(fsr_part* p, const fsr_sam* sam, const fsr_elmdata* elm);
int build_hex20_operators(fsr_part* p, const fsr_sam* sam, const fsr_elmdata* elm);
int build_wedg15_operators(fsr_part* p, const fsr_sam* sam, const fsr_elmdata* elm);
int launch_k2_wedg15_vm(fsr_part* p, int nsteps, int nsteps_pad, double* vm_dev, size_t ld_vm, cudaStream_t s);
int build_thickshell_operators(fsr_part* p, const fsr_sam* sam, const fsr_elmdata* elm);
int launch_k2_thickshell_vm(fsr_part* p, int nsteps, int nsteps_pad, double* vm_dev, size_t ld_vm, cudaStream_t s);
int build_linsolid_operators(fsr_part* p, const fsr_sam* sam, const fsr_elmdata* elm);
int launch_k2_linsolid_vm(fsr_part* p, int nsteps, int nsteps_pad, double* vm_dev, size_t ld_vm, cudaStream_t s);
int launch_k2_hex20_vm(fsr_part* p, int nsteps, int nsteps_pad, double* vm_dev, size_t ld_vm, cudaStream_t s);
int launch_k2_shell_vm(fsr_part* p, int nsteps, int nsteps_pad, double* vm_dev, size_t ld_vm,
                       cudaStream_t s);
int launch_k2_tet10_vm(fsr_part* p, int nsteps, int nsteps_pad, double* vm_dev, size_t ld_vm,
                       cudaStream_t s);
// k2_shell.cu: von Mises of a thin-shell family straight into float / double step records (results database, -vmStress only)
template <class OUT_T>
int launch_k2_shell_rec(fsr_part* p, int fam, const double* U, int nsteps, int nsteps_pad, const long long* roff, OUT_T* out, size_t ld_out,
                        cudaStream_t s);
int launch_beam_full(fsr_part* p, double* sres, cudaStream_t s);
int launch_k2_full(fsr_part* p, double* resmat, double* stress, double* strain, double* sres,
                   cudaStream_t s);
}  // namespace fsr
