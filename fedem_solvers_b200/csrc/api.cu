// api.cu -- the extern "C" boundary of libfedem_b200.so (see include/fedem_b200.h) and the host
// orchestration of the step-tiled pipeline:  pack Q -> K1 (DMMA expansion) -> K2 (element
// kernels with fused envelope) per tile of steps.
#include <cstdarg>
#include <algorithm>
#include <mutex>
#include <set>
#include <utility>

#include "common.cuh"

namespace fsr {

static thread_local char g_err[1024] = "";
long long g_launches = 0;

void set_error(const char* fmt, ...)
{
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int smem_opt_in(const void* kernel, size_t bytes)
{
  static std::mutex mtx;
  static std::set<std::pair<int, const void*>> done;
  int dev = 0;
  FSR_CUDA(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lock(mtx);
  if (done.count({dev, kernel})) return FSR_OK;
  FSR_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  done.insert({dev, kernel});
  return FSR_OK;
}

int nstrp_of(int type)
{
  switch (type) {
    case 21: case 23: return 6;
    case 22: case 24: return 8;
    case 31: return 12;
    case 32: return 16;
    case 41: return 10;
    case 42: return 15;
    case 43: return 20;
    case 44: return 8;
    case 45: return 4;
    case 46: return 6;
    default: return 0;  // beams (11) carry section forces only: nstrp = 0 (elStressModule.f90:161-164)
  }
}

bool supported_type(int type) { return type == 24 || type == 23 || type == 31 || type == 32 || (type >= 41 && type <= 46) || type == 11; }

static int round_up(int a, int b) { return (a + b - 1) / b * b; }

static inline unsigned spread10(unsigned v)  // 10 bits -> every third bit
{
  v &= 0x3ffu;
  v = (v | (v << 16)) & 0x030000ffu;
  v = (v | (v << 8)) & 0x0300f00fu;
  v = (v | (v << 4)) & 0x030c30c3u;
  v = (v | (v << 2)) & 0x09249249u;
  return v;
}

// The reference walks the elements in SAM order (stressRoutines.f90:169).  Elements that share nodes
// read the same rows of U, so the K2 kernels process them in Morton order of their centroids: the
// rows a warp needs were just touched by its neighbours and are still in L2.  Results keep their SAM
// positions (ptoff), only the order of evaluation changes.
std::vector<int> elements_of_type(const fsr_part* p, const fsr_sam* sam, const fsr_elmdata* elm, int type)
{
  std::vector<int> el;
  for (int e = 0; e < sam->nel; ++e)
    if (sam->melcon[e] == type && !(elm->elmid && elm->elmid[e] < 1)) el.push_back(e);
  if (p->elem_order != 0 || el.size() < 2) return el;
  double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
  for (int n = 0; n < sam->nnod; ++n)
    for (int k = 0; k < 3; ++k) {
      lo[k] = std::min(lo[k], elm->xyz[3 * (size_t)n + k]);
      hi[k] = std::max(hi[k], elm->xyz[3 * (size_t)n + k]);
    }
  double inv[3];
  for (int k = 0; k < 3; ++k) inv[k] = hi[k] > lo[k] ? 1023.999 / (hi[k] - lo[k]) : 0.0;
  std::vector<std::pair<unsigned, int>> key(el.size());
  for (size_t i = 0; i < el.size(); ++i) {
    const int e = el[i], ip0 = sam->mpmnpc[e] - 1, nn = sam->mpmnpc[e + 1] - sam->mpmnpc[e];
    double c[3] = {0, 0, 0};
    for (int k = 0; k < nn; ++k) {
      const int n = sam->mmnpc[ip0 + k] - 1;
      if (n < 0 || n >= sam->nnod) continue;  // reported by the family builder
      for (int d = 0; d < 3; ++d) c[d] += elm->xyz[3 * (size_t)n + d];
    }
    unsigned q[3];
    for (int d = 0; d < 3; ++d) q[d] = (unsigned)((c[d] / std::max(nn, 1) - lo[d]) * inv[d]);
    key[i] = {spread10(q[0]) | (spread10(q[1]) << 1) | (spread10(q[2]) << 2), e};
  }
  std::stable_sort(key.begin(), key.end());
  for (size_t i = 0; i < el.size(); ++i) el[i] = key[i].second;
  return el;
}

static void free_family(FamilyData& f)
{
  cudaFree(f.elem); cudaFree(f.edof); cudaFree(f.ptoff); cudaFree(f.Sfrag); cudaFree(f.failed); cudaFree(f.Gfrag); cudaFree(f.Efrag);
  cudaFree(f.aux); cudaFree(f.sub[0]); cudaFree(f.sub[1]); cudaFree(f.sub[2]); cudaFree(f.fast); cudaFree(f.fast2); cudaFree(f.edof2);
  f = FamilyData();
}

int ensure_batch_buffers(fsr_part* p, bool need_vm_tile)
{
  if (!p->Qt) FSR_CUDA(cudaMalloc(&p->Qt, sizeof(double) * (size_t)p->step_tile * p->ldk));
  if (!p->U) FSR_CUDA(cudaMalloc(&p->U, sizeof(double) * ((size_t)p->nrows_pad * p->step_tile + 64)));  // +64: prefetch slack
  if (p->planar && !p->Up) FSR_CUDA(cudaMalloc(&p->Up, sizeof(double) * ((size_t)p->np_rows_pad * p->step_tile + 64)));
  if (need_vm_tile && !p->vm_tile && p->npts > 0)
    FSR_CUDA(cudaMalloc(&p->vm_tile, sizeof(double) * (size_t)p->step_tile * p->npts));
  return FSR_OK;
}

static int choose_step_tile(fsr_part* p, int requested)
{
  if (requested > 0) return round_up(requested, 64);
  size_t free_b = 0, total_b = 0;
  if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) return 512;
  // per step: U row slice + vm staging; keep half of the free memory for everything else
  double per_step = 8.0 * ((double)p->nrows_pad + (double)p->npts + p->ldk);
  long long t = (long long)(0.5 * (double)free_b / per_step);
  if (t > 1024) t = 1024;
  if (t < 64) t = 64;
  return (int)(t / 64 * 64);
}

}  // namespace fsr

using namespace fsr;

extern "C" {

const char* fsr_last_error(void) { return g_err; }

long long fsr_kernel_launches(int reset)
{
  long long n = g_launches;
  if (reset) g_launches = 0;
  return n;
}

int fsr_part_create(fsr_part** out, const fsr_sam* sam_in, const fsr_elmdata* elm, const fsr_options* opt)
{
  if (!out || !sam_in || !elm) { set_error("fsr_part_create: null argument"); return FSR_ERR_ARG; }
  *out = nullptr;
  std::vector<int> melcon_eff;
  int quad_ngauss = 2;
  fsr_sam sam_eff = *sam_in;
  if (sam_in->melcon && sam_in->nel > 0) {
    effective_element_types(sam_in, opt, melcon_eff, quad_ngauss);
    sam_eff.melcon = melcon_eff.data();
  }
  return part_create_mapped(out, &sam_eff, elm, opt, quad_ngauss);
}

}  // extern "C"

namespace fsr {

void effective_element_types(const fsr_sam* sam_in, const fsr_options* opt, std::vector<int>& melcon_eff, int& quad_ngauss)
{
  // Legacy thin shells (types 21 FFT3 and 22 FFQ4, parts reduced with -useANDESformulation-): with the default stress
  // formulations of fedem_stress (stressmain.C:72-78: -fftStressForm 1, -ffqStressForm 2) STR21 runs exactly the statements of
  // STR23 (FTSA31 / FTSA32 / FTS38, elStressModule.f90:559-562,586-587 vs :935,953) and STR22 exactly those of STR24
  // (pMatStiff projection + STR22a with 2 x 2 Gauss points, :675-686,717-722 vs :1039-1076), so they join those families.
  // Of the other (private, non-default) formulations -ffqStressForm 1 and -fftStressForm 0 / 2 are served below; FFQ elements of a
  // run with -ffqStressForm 0 (STR22b: the Femlib nodal evaluation FQS32) get NO results, like any unsupported type, and
  // fedem_stress says so (stress_driver.cu).
  const int ffq = opt && opt->reserved[1] ? opt->reserved[1] - 1 : 2, fft = opt && opt->reserved[2] ? opt->reserved[2] - 1 : 1;
  // -ffqStressForm 1 is STR22a with one Gauss point (:761-768,806-809): the quad operator builder takes the point count, so it
  // is served as well as long as the part has no ANDES quads (which always use 2 x 2) next to the FFQ ones.
  quad_ngauss = 2;
  melcon_eff.assign(sam_in->melcon, sam_in->melcon + sam_in->nel);
  bool has24 = false, has22 = false;
  for (int t : melcon_eff) { has24 |= t == 24; has22 |= t == 22; }
  const bool ffq1 = ffq == 1 && has22 && !has24;
  if (ffq1) quad_ngauss = 1;
  // -fftStressForm 0 / 2 is STR21 on FTS31 / FTS32 (:559-566,586-592): the triangle operator builder has that membrane formulation
  // too, for a part whose triangles are all FFT3 (bit 8 of quad_ngauss carries the switch to part_create_mapped)
  bool has23 = false, has21 = false;
  for (int t : melcon_eff) { has23 |= t == 23; has21 |= t == 21; }
  const bool fft_legacy = fft != 1 && has21 && !has23;
  if (fft_legacy) quad_ngauss |= 0x100;
  for (int& t : melcon_eff) {
    if (t == 21) t = (fft == 1 || fft_legacy) ? 23 : 0;
    else if (t == 22) t = (ffq == 2 || ffq1) ? 24 : 0;
  }
}

// fsr_part_create after the legacy shell types were mapped (an element block takes the mapping of its parent part)
int part_create_mapped(fsr_part** out, const fsr_sam* sam, const fsr_elmdata* elm, const fsr_options* opt, int quad_ngauss)
{
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) {
    set_error("no CUDA device available: this library has no CPU fallback");
    return FSR_ERR_CUDA;
  }
  int dev = opt ? opt->device : 0;
  if (dev < 0 || dev >= ndev) { set_error("device %d out of range (0..%d)", dev, ndev - 1); return FSR_ERR_ARG; }
  FSR_CUDA(cudaSetDevice(dev));
  cudaDeviceProp prop;
  FSR_CUDA(cudaGetDeviceProperties(&prop, dev));
  if (prop.major != 10) {
    set_error("device %d is sm_%d%d; this library is built for sm_100a (B200) only", dev, prop.major, prop.minor);
    return FSR_ERR_CUDA;
  }
  if (sam->nnod < 1 || sam->nel < 0 || sam->ndof < 1 || !sam->madof || !sam->mpmnpc || !sam->mmnpc ||
      !sam->melcon || !sam->meqn || !sam->msc || !elm->xyz || !elm->emod || !elm->rny || !elm->thk) {
    set_error("fsr_part_create: incomplete SAM / element data");
    return FSR_ERR_ARG;
  }
  if (sam->nceq > 0 && (!sam->mpmceq || !sam->mmceq || !sam->ttcc)) {
    set_error("fsr_part_create: nceq > 0 but constraint arrays missing");
    return FSR_ERR_ARG;
  }

  fsr_part* p = new fsr_part();
  p->device = dev;
  p->nnod = sam->nnod; p->nel = sam->nel; p->ndof = sam->ndof; p->ndof1 = sam->ndof1;
  p->ndof2 = sam->ndof2; p->ngen = sam->ngen; p->neq = sam->neq; p->nceq = sam->nceq;
  p->ndim = sam->ndof2 + sam->ngen;
  p->stressForm = opt ? opt->stressForm : 0;
  p->quad_ngauss = quad_ngauss & 0xff;
  p->tri_legacy = (quad_ngauss >> 8) & 1;
  p->elem_order = opt ? opt->reserved[0] : 0;
  // ldk: multiple of 4 with ldk % 8 == 4 (bank-conflict-free fragment loads in K1)
  p->ldk = round_up(std::max(p->ndim, 1), 4);
  if (p->ldk % 8 == 0) p->ldk += 4;
  p->nrows_pad = round_up(p->ndof, 128);

  // result point offsets in SAM (processing) order, stressRoutines.f90:169-331
  p->ptoff_host.assign((size_t)sam->nel + 1, 0);
  p->melcon_host.assign(sam->melcon, sam->melcon + sam->nel);
  p->sam_keep.keep(sam);
  p->madof_host.assign(sam->madof, sam->madof + sam->nnod + 1);
  p->xyz_host.assign(elm->xyz, elm->xyz + (size_t)3 * sam->nnod);
  p->nenod_host.resize((size_t)sam->nel);
  p->active_host.resize((size_t)sam->nel);
  for (int e = 0; e < sam->nel; ++e) {
    p->nenod_host[e] = sam->mpmnpc[e + 1] - sam->mpmnpc[e];
    p->active_host[e] = (elm->elmid && elm->elmid[e] < 1) ? 0 : 1;
  }
  int n = 0, nskipped = 0;
  for (int e = 0; e < sam->nel; ++e) {
    p->ptoff_host[e] = n;
    if (elm->elmid && elm->elmid[e] < 1) continue;
    if (!supported_type(sam->melcon[e])) { ++nskipped; continue; }
    n += nstrp_of(sam->melcon[e]);
  }
  p->ptoff_host[sam->nel] = n;
  p->npts = n;
  (void)nskipped;

  int rc = FSR_OK;
  auto fail = [&](int code) { fsr_part_destroy(p); return code; };
  if (cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking) != cudaSuccess) { set_error("cudaStreamCreate failed"); return fail(FSR_ERR_CUDA); }
  for (int i = 0; i < 4; ++i)
    if (cudaEventCreate(&p->ev[i]) != cudaSuccess) { set_error("cudaEventCreate failed"); return fail(FSR_ERR_CUDA); }

  auto up = [&](double** dst, const double* src, size_t cnt) -> int {
    FSR_CUDA(cudaMalloc(dst, sizeof(double) * std::max<size_t>(cnt, 1)));
    FSR_CUDA(cudaMemcpyAsync(*dst, src, sizeof(double) * cnt, cudaMemcpyHostToDevice, p->stream));
    return FSR_OK;
  };
  if ((rc = up(&p->xyz, elm->xyz, (size_t)3 * sam->nnod))) return fail(rc);
  if ((rc = up(&p->emod, elm->emod, (size_t)sam->nel))) return fail(rc);
  if ((rc = up(&p->rny, elm->rny, (size_t)sam->nel))) return fail(rc);
  if ((rc = up(&p->thk, elm->thk, (size_t)sam->nel))) return fail(rc);

  if (cudaMalloc(&p->R, sizeof(double) * (size_t)p->nrows_pad * p->ldk) != cudaSuccess ||
      cudaMalloc(&p->env_max, sizeof(double) * std::max(p->npts, 1)) != cudaSuccess ||
      cudaMalloc(&p->env_min, sizeof(double) * std::max(p->npts, 1)) != cudaSuccess) {
    set_error("device allocation failed (R: %zu bytes)", sizeof(double) * (size_t)p->nrows_pad * p->ldk);
    return fail(FSR_ERR_ALLOC);
  }
  if ((rc = build_shell_operators(p, sam, elm))) return fail(rc);
  if ((rc = build_solid_operators(p, sam, elm))) return fail(rc);
  if ((rc = build_beam_operators(p, sam, elm))) return fail(rc);
  if ((rc = build_hex20_operators(p, sam, elm))) return fail(rc);
  if ((rc = build_linsolid_operators(p, sam, elm))) return fail(rc);
  if ((rc = build_wedg15_operators(p, sam, elm))) return fail(rc);
  if ((rc = build_thickshell_operators(p, sam, elm))) return fail(rc);
  if ((rc = fsr_reset_envelope(p))) return fail(rc);

  // count failed elements (they get hugeVal results, the run continues)
  p->nfailed = 0;
  for (int f = 0; f < FAM_COUNT; ++f) {
    FamilyData& fd = p->fam[f];
    if (fd.nelt == 0) continue;
    std::vector<unsigned char> h(fd.nelt);
    if (cudaMemcpy(h.data(), fd.failed, fd.nelt, cudaMemcpyDeviceToHost) != cudaSuccess) { set_error("copy of element status failed: %s", cudaGetErrorString(cudaGetLastError())); return fail(FSR_ERR_CUDA); }
    for (unsigned char c : h) p->nfailed += c ? 1 : 0;
  }
  p->step_tile = choose_step_tile(p, opt ? opt->step_tile : 0);
  *out = p;
  return p->nfailed;
}

}  // namespace fsr

extern "C" {

void fsr_part_destroy(fsr_part* p)
{
  if (!p) return;
  cudaSetDevice(p->device);
  if (p->stream) cudaStreamSynchronize(p->stream);
  cudaFree(p->xyz); cudaFree(p->emod); cudaFree(p->rny); cudaFree(p->thk);
  cudaFree(p->R); cudaFree(p->Qt); cudaFree(p->U); cudaFree(p->vm_tile); cudaFree(p->Qstage);
  cudaFree(p->Rp); cudaFree(p->Up); cudaFree(p->prow_src); cudaFree(p->prow_w); cudaFree(p->k1_tiles);
  cudaFree(p->env_max); cudaFree(p->env_min); cudaFree(p->env_snap);
  if (p->copy_stream) { cudaStreamSynchronize(p->copy_stream); cudaStreamDestroy(p->copy_stream); }
  if (p->ev_snap) cudaEventDestroy(p->ev_snap);
  if (p->ev_copied) cudaEventDestroy(p->ev_copied);
  for (int f = 0; f < FAM_COUNT; ++f) free_family(p->fam[f]);
  if (p->pinned) cudaFreeHost(p->pinned);
  for (int i = 0; i < 4; ++i) if (p->ev[i]) cudaEventDestroy(p->ev[i]);
  for (auto& tr : p->evring) for (int i = 0; i < 3; ++i) if (tr[i]) cudaEventDestroy(tr[i]);
  if (p->own_stream && p->stream) cudaStreamDestroy(p->stream);
  delete p;
}

int fsr_set_recovery(fsr_part* p, const double* B, int ldB, const double* E, int ldE)
{
  if (!p) { set_error("fsr_set_recovery: null handle"); return FSR_ERR_ARG; }
  if (p->ndof1 > 0 && p->ndof2 > 0 && (!B || ldB < p->ndof1)) { set_error("fsr_set_recovery: bad B / ldB"); return FSR_ERR_ARG; }
  if (p->ndof1 > 0 && p->ngen > 0 && (!E || ldE < p->ndof1)) { set_error("fsr_set_recovery: bad E / ldE"); return FSR_ERR_ARG; }
  FSR_CUDA(cudaSetDevice(p->device));
  if (!p->sam_keep.valid) { set_error("fsr_set_recovery: SAM maps missing"); return FSR_ERR_STATE; }
  fsr_sam sam = p->sam_keep.view();
  if (int rc = build_row_operator(p, &sam, B, ldB, E, ldE)) return rc;
  return build_planar_rows(p);
}

int fsr_set_stream(fsr_part* p, void* stream)
{
  if (!p) return FSR_ERR_ARG;
  cudaSetDevice(p->device);
  if (p->stream) cudaStreamSynchronize(p->stream);
  if (p->own_stream && p->stream) cudaStreamDestroy(p->stream);
  p->stream = (cudaStream_t)stream;
  p->own_stream = false;
  return FSR_OK;
}

int fsr_num_result_points(const fsr_part* p) { return p ? p->npts : FSR_ERR_ARG; }
int fsr_ndim(const fsr_part* p) { return p ? p->ndim : FSR_ERR_ARG; }
int fsr_result_point_offsets(const fsr_part* p, int* off)
{
  if (!p || !off) return FSR_ERR_ARG;
  std::copy(p->ptoff_host.begin(), p->ptoff_host.end(), off);
  return FSR_OK;
}

int fsr_reset_envelope(fsr_part* p)
{
  if (!p) return FSR_ERR_ARG;
  FSR_CUDA(cudaSetDevice(p->device));
  std::vector<double> hmin((size_t)std::max(p->npts, 1), kHuge);
  FSR_CUDA(cudaMemsetAsync(p->env_max, 0, sizeof(double) * std::max(p->npts, 1), p->stream));
  FSR_CUDA(cudaMemcpyAsync(p->env_min, hmin.data(), sizeof(double) * hmin.size(), cudaMemcpyHostToDevice, p->stream));
  FSR_CUDA(cudaStreamSynchronize(p->stream));
  return FSR_OK;
}

int fsr_get_envelope(fsr_part* p, double* vm_max, double* vm_min)
{
  if (!p) return FSR_ERR_ARG;
  FSR_CUDA(cudaSetDevice(p->device));
  FSR_CUDA(cudaStreamSynchronize(p->stream));
  if (vm_max) FSR_CUDA(cudaMemcpy(vm_max, p->env_max, sizeof(double) * p->npts, cudaMemcpyDeviceToHost));
  if (vm_min) FSR_CUDA(cudaMemcpy(vm_min, p->env_min, sizeof(double) * p->npts, cudaMemcpyDeviceToHost));
  return FSR_OK;
}

// The envelopes as they are after the work queued so far, delivered to the host without stopping the pipeline: a
// device-to-device snapshot in stream order, then the PCIe copy on a second stream while the next tiles compute.
int fsr_get_envelope_async(fsr_part* p, double* vm_max, double* vm_min)
{
  if (!p || (!vm_max && !vm_min)) { set_error("fsr_get_envelope_async: bad arguments"); return FSR_ERR_ARG; }
  FSR_CUDA(cudaSetDevice(p->device));
  const size_t np = (size_t)std::max(p->npts, 1);
  if (!p->copy_stream) {
    FSR_CUDA(cudaStreamCreateWithFlags(&p->copy_stream, cudaStreamNonBlocking));
    FSR_CUDA(cudaEventCreateWithFlags(&p->ev_snap, cudaEventDisableTiming));
    FSR_CUDA(cudaEventCreateWithFlags(&p->ev_copied, cudaEventDisableTiming));
    FSR_CUDA(cudaMalloc(&p->env_snap, sizeof(double) * 2 * np));
    FSR_CUDA(cudaEventRecord(p->ev_copied, p->copy_stream));
  }
  cudaStream_t s = p->stream;
  FSR_CUDA(cudaStreamWaitEvent(s, p->ev_copied, 0));   // the previous read-back has left the snapshot buffer
  FSR_CUDA(cudaMemcpyAsync(p->env_snap, p->env_max, sizeof(double) * p->npts, cudaMemcpyDeviceToDevice, s));
  FSR_CUDA(cudaMemcpyAsync(p->env_snap + np, p->env_min, sizeof(double) * p->npts, cudaMemcpyDeviceToDevice, s));
  FSR_CUDA(cudaEventRecord(p->ev_snap, s));
  FSR_CUDA(cudaStreamWaitEvent(p->copy_stream, p->ev_snap, 0));
  if (vm_max) FSR_CUDA(cudaMemcpyAsync(vm_max, p->env_snap, sizeof(double) * p->npts, cudaMemcpyDeviceToHost, p->copy_stream));
  if (vm_min) FSR_CUDA(cudaMemcpyAsync(vm_min, p->env_snap + np, sizeof(double) * p->npts, cudaMemcpyDeviceToHost, p->copy_stream));
  FSR_CUDA(cudaEventRecord(p->ev_copied, p->copy_stream));
  return FSR_OK;
}

// waits for everything queued on the handle: tiles of steps (fsr_recover_async / fsr_recover_dev) and read-backs
int fsr_synchronize(fsr_part* p)
{
  if (!p) { set_error("fsr_synchronize: null handle"); return FSR_ERR_ARG; }
  FSR_CUDA(cudaSetDevice(p->device));
  FSR_CUDA(cudaStreamSynchronize(p->stream));
  if (p->copy_stream) FSR_CUDA(cudaStreamSynchronize(p->copy_stream));
  return FSR_OK;
}

// waits for the read-backs only (the tiles queued after them keep running)
int fsr_envelope_wait(fsr_part* p)
{
  if (!p) { set_error("fsr_envelope_wait: null handle"); return FSR_ERR_ARG; }
  FSR_CUDA(cudaSetDevice(p->device));
  if (p->ev_copied) FSR_CUDA(cudaEventSynchronize(p->ev_copied));
  return FSR_OK;
}

int fsr_copy_envelope_dev(fsr_part* p, double* vm_max_dst_dev, double* vm_min_dst_dev, void* stream)
{
  if (!p) return FSR_ERR_ARG;
  FSR_CUDA(cudaSetDevice(p->device));
  cudaStream_t s = stream ? (cudaStream_t)stream : p->stream;
  if (vm_max_dst_dev)
    FSR_CUDA(cudaMemcpyAsync(vm_max_dst_dev, p->env_max, sizeof(double) * p->npts, cudaMemcpyDeviceToDevice, s));
  if (vm_min_dst_dev)
    FSR_CUDA(cudaMemcpyAsync(vm_min_dst_dev, p->env_min, sizeof(double) * p->npts, cudaMemcpyDeviceToDevice, s));
  return FSR_OK;
}

int fsr_envelope_dev(fsr_part* p, double** vm_max_dev, double** vm_min_dev)
{
  if (!p) return FSR_ERR_ARG;
  if (vm_max_dev) *vm_max_dev = p->env_max;
  if (vm_min_dev) *vm_min_dev = p->env_min;
  return FSR_OK;
}

// sv[t][dof] (step-major, what ffr_getData delivers per step) -> U[dof][t]
__global__ void load_u_kernel(double* __restrict__ U, size_t ldu, const double* __restrict__ sv, int ndof, int nt)
{
  __shared__ double tile[32][33];
  const int d0 = blockIdx.x * 32, t0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int t = t0 + j, d = d0 + threadIdx.x;
    if (t < nt && d < ndof) tile[j][threadIdx.x] = sv[(size_t)t * ndof + d];
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int d = d0 + j, t = t0 + threadIdx.x;
    if (t < nt && d < ndof) U[(size_t)d * ldu + t] = tile[threadIdx.x][j];
  }
}

}  // extern "C"

namespace fsr {
// nodal displacements of nt steps (host, step-major [nt][ndof]) into U[dof][t] of the part: the place K1 would fill
int upload_displacements(fsr_part* p, const double* sv_host, int nt, cudaStream_t s)
{
  int rc = ensure_batch_buffers(p, false);
  if (rc) return rc;
  const size_t need = sizeof(double) * (size_t)p->ndof * p->step_tile;
  if (p->Qstage_cap < need) {
    FSR_CUDA(cudaStreamSynchronize(s));
    cudaFree(p->Qstage); p->Qstage = nullptr; p->Qstage_cap = 0;
    FSR_CUDA(cudaMalloc(&p->Qstage, need));
    p->Qstage_cap = need;
  }
  FSR_CUDA(cudaMemcpyAsync(p->Qstage, sv_host, sizeof(double) * (size_t)p->ndof * nt, cudaMemcpyHostToDevice, s));
  dim3 blk(32, 8), grd((p->ndof + 31) / 32, (nt + 31) / 32);
  load_u_kernel<<<grd, blk, 0, s>>>(p->U, (size_t)p->step_tile, p->Qstage, p->ndof, nt);
  FSR_LAUNCH_CHECK();
  return planar_rows_from_u(p, nt, s);   // the in-plane rows of flat shell regions
}
}  // namespace fsr

extern "C" {

static int run_k2(fsr_part* p, int nsteps, int nsteps_pad, double* vm_dev, size_t ld_vm, cudaStream_t s)
{
  int rc;
  if ((rc = launch_k2_shell_vm(p, nsteps, nsteps_pad, vm_dev, ld_vm, s))) return rc;
  if ((rc = launch_k2_tet10_vm(p, nsteps, nsteps_pad, vm_dev, ld_vm, s))) return rc;
  if ((rc = launch_k2_hex20_vm(p, nsteps, nsteps_pad, vm_dev, ld_vm, s))) return rc;
  if ((rc = launch_k2_linsolid_vm(p, nsteps, nsteps_pad, vm_dev, ld_vm, s))) return rc;
  if ((rc = launch_k2_wedg15_vm(p, nsteps, nsteps_pad, vm_dev, ld_vm, s))) return rc;
  return launch_k2_thickshell_vm(p, nsteps, nsteps_pad, vm_dev, ld_vm, s);
}

// One tile of steps, everything on `s`.  vm_dev may be NULL (envelope only).
static int run_tile(fsr_part* p, const double* Q_dev, int ldq, int nsteps, double* vm_dev, size_t ld_vm,
                    cudaStream_t s, bool timed, bool full_u = false)
{
  int nsteps_pad = round_up(nsteps, 64);
  int rc;
  cudaEvent_t* ev = nullptr;
  if (timed && p->ntimed < (int)(sizeof(p->evring) / sizeof(p->evring[0]))) {
    ev = p->evring[p->ntimed];
    for (int i = 0; i < 3; ++i)
      if (!ev[i]) FSR_CUDA(cudaEventCreate(&ev[i]));
    ++p->ntimed;
  }
  if (ev) cudaEventRecord(ev[0], s);
  if ((rc = launch_pack_q(p, Q_dev, ldq, nsteps, nsteps_pad, s))) return rc;
  if ((rc = launch_k1_vm(p, nsteps_pad, s, full_u))) return rc;
  if (ev) cudaEventRecord(ev[1], s);
  if ((rc = run_k2(p, nsteps, nsteps_pad, vm_dev, ld_vm, s))) return rc;
  if (ev) cudaEventRecord(ev[2], s);
  return FSR_OK;
}

// calcStresses on nodal displacements that are already there (stress.f90:397, readIntDisplacements: the direct solution of a
// linear analysis on the results files; no B / E matrices, no expansion): sv_hist [nsteps x ndof] step-major in nodal DOF order.
// vm_hist as fsr_recover; the envelopes accumulate.
int fsr_recover_displacements(fsr_part* p, const double* sv_hist, int nsteps, double* vm_hist)
{
  if (!p || !sv_hist || nsteps < 0) { set_error("fsr_recover_displacements: bad arguments"); return FSR_ERR_ARG; }
  FSR_CUDA(cudaSetDevice(p->device));
  int rc = ensure_batch_buffers(p, vm_hist != nullptr);
  if (rc) return rc;
  cudaStream_t s = p->stream;
  for (int t0 = 0; t0 < nsteps; t0 += p->step_tile) {
    const int nt = std::min(p->step_tile, nsteps - t0);
    if ((rc = upload_displacements(p, sv_hist + (size_t)t0 * p->ndof, nt, s))) return rc;
    if ((rc = run_k2(p, nt, round_up(nt, 64), vm_hist ? p->vm_tile : nullptr, (size_t)p->npts, s))) return rc;
    if (vm_hist) {
      FSR_CUDA(cudaMemcpyAsync(vm_hist + (size_t)t0 * p->npts, p->vm_tile, sizeof(double) * (size_t)nt * p->npts, cudaMemcpyDeviceToHost, s));
      FSR_CUDA(cudaStreamSynchronize(s));
    }
  }
  FSR_CUDA(cudaStreamSynchronize(s));
  return FSR_OK;
}

}  // extern "C"

namespace fsr {
// One converged step of the dynamics solver: q = [finit; vg] (host) -> expanded displacements and the von Mises stress
// of every result point, queued on the part's stream (the running envelope takes the step along).  sv_host [ndof] and
// vm_host [npts] should be page-locked; the caller synchronises (fsr_synchronize) before reading them.
int step_enqueue(fsr_part* p, const double* q, double* sv_host, double* vm_host)
{
  if (!p->have_R) { set_error("recovery step: call fsr_set_recovery first"); return FSR_ERR_STATE; }
  FSR_CUDA(cudaSetDevice(p->device));
  int rc = ensure_batch_buffers(p, true);
  if (rc) return rc;
  cudaStream_t s = p->stream;
  if (p->Qstage_cap < sizeof(double) * (size_t)p->ndim) {
    FSR_CUDA(cudaStreamSynchronize(s));
    cudaFree(p->Qstage); p->Qstage = nullptr; p->Qstage_cap = 0;
    FSR_CUDA(cudaMalloc(&p->Qstage, sizeof(double) * (size_t)std::max(p->ndim, 1)));
    p->Qstage_cap = sizeof(double) * (size_t)p->ndim;
  }
  FSR_CUDA(cudaMemcpyAsync(p->Qstage, q, sizeof(double) * p->ndim, cudaMemcpyHostToDevice, s));
  if ((rc = run_tile(p, p->Qstage, p->ndim, 1, p->vm_tile, (size_t)p->npts, s, false, sv_host != nullptr))) return rc;
  if (sv_host)   // column t = 0 of U[dof][t]
    FSR_CUDA(cudaMemcpy2DAsync(sv_host, sizeof(double), p->U, sizeof(double) * p->step_tile, sizeof(double), (size_t)p->ndof,
                               cudaMemcpyDeviceToHost, s));
  if (vm_host && p->npts > 0) FSR_CUDA(cudaMemcpyAsync(vm_host, p->vm_tile, sizeof(double) * p->npts, cudaMemcpyDeviceToHost, s));
  return FSR_OK;
}
}  // namespace fsr

extern "C" {

int fsr_recover_dev(fsr_part* p, const double* Q_dev, int ldq, int nsteps, double* vm_hist_dev,
                    size_t ld_vm, void* stream)
{
  if (!p || !Q_dev || nsteps < 0 || ldq < p->ndim) { set_error("fsr_recover_dev: bad arguments"); return FSR_ERR_ARG; }
  if (!p->have_R) { set_error("fsr_recover_dev: call fsr_set_recovery first"); return FSR_ERR_STATE; }
  if (vm_hist_dev && ld_vm < (size_t)p->npts) { set_error("fsr_recover_dev: ld_vm < number of result points"); return FSR_ERR_ARG; }
  FSR_CUDA(cudaSetDevice(p->device));
  int rc = ensure_batch_buffers(p, false);
  if (rc) return rc;
  cudaStream_t s = stream ? (cudaStream_t)stream : p->stream;
  for (int t0 = 0; t0 < nsteps; t0 += p->step_tile) {
    int nt = std::min(p->step_tile, nsteps - t0);
    rc = run_tile(p, Q_dev + (size_t)t0 * ldq, ldq, nt, vm_hist_dev ? vm_hist_dev + (size_t)t0 * ld_vm : nullptr,
                  ld_vm, s, true);
    if (rc) return rc;
  }
  return FSR_OK;
}

static int recover_host(fsr_part* p, const double* Q, int ldq, int nsteps, double* vm_hist, bool wait);

int fsr_recover(fsr_part* p, const double* Q, int ldq, int nsteps, double* vm_hist) { return recover_host(p, Q, ldq, nsteps, vm_hist, true); }

// fsr_recover without the history and without waiting: the window is queued (H2D of Q, K1, K2 + envelope) and the call
// returns; Q should be page-locked for the copy to be asynchronous and must stay untouched until the copy has happened
// (fsr_synchronize, or the next fsr_recover* call, which reuses the staging buffer in stream order).
int fsr_recover_async(fsr_part* p, const double* Q, int ldq, int nsteps) { return recover_host(p, Q, ldq, nsteps, nullptr, false); }

static int recover_host(fsr_part* p, const double* Q, int ldq, int nsteps, double* vm_hist, bool wait)
{
  if (!p || !Q || nsteps < 0 || ldq < p->ndim) { set_error("fsr_recover: bad arguments"); return FSR_ERR_ARG; }
  if (!p->have_R) { set_error("fsr_recover: call fsr_set_recovery first"); return FSR_ERR_STATE; }
  FSR_CUDA(cudaSetDevice(p->device));
  int rc = ensure_batch_buffers(p, vm_hist != nullptr);
  if (rc) return rc;
  cudaStream_t s = p->stream;
  size_t qbytes = sizeof(double) * (size_t)ldq * nsteps;
  if (p->Qstage_cap < qbytes) {
    FSR_CUDA(cudaStreamSynchronize(s));   // earlier windows may still read the old staging buffer
    cudaFree(p->Qstage); p->Qstage = nullptr; p->Qstage_cap = 0;
    FSR_CUDA(cudaMalloc(&p->Qstage, std::max<size_t>(qbytes, 8)));
    p->Qstage_cap = qbytes;
  }
  p->ntimed = 0;
  FSR_CUDA(cudaMemcpyAsync(p->Qstage, Q, qbytes, cudaMemcpyHostToDevice, s));
  for (int t0 = 0; t0 < nsteps; t0 += p->step_tile) {
    int nt = std::min(p->step_tile, nsteps - t0);
    rc = run_tile(p, p->Qstage + (size_t)t0 * ldq, ldq, nt, vm_hist ? p->vm_tile : nullptr, (size_t)p->npts, s, true);
    if (rc) return rc;
    if (vm_hist)
      FSR_CUDA(cudaMemcpyAsync(vm_hist + (size_t)t0 * p->npts, p->vm_tile, sizeof(double) * (size_t)nt * p->npts,
                               cudaMemcpyDeviceToHost, s));
    if (vm_hist) FSR_CUDA(cudaStreamSynchronize(s));  // vm_tile is reused by the next tile
  }
  if (wait) FSR_CUDA(cudaStreamSynchronize(s));
  return FSR_OK;
}

int fsr_expand(fsr_part* p, const double* Q, int ldq, int nsteps, double* U_host)
{
  if (!p || !Q || !U_host || nsteps < 0 || ldq < p->ndim) { set_error("fsr_expand: bad arguments"); return FSR_ERR_ARG; }
  if (!p->have_R) { set_error("fsr_expand: call fsr_set_recovery first"); return FSR_ERR_STATE; }
  FSR_CUDA(cudaSetDevice(p->device));
  int rc = ensure_batch_buffers(p, false);
  if (rc) return rc;
  cudaStream_t s = p->stream;
  size_t qbytes = sizeof(double) * (size_t)ldq * nsteps;
  double* dQ = nullptr;
  FSR_CUDA(cudaMalloc(&dQ, std::max<size_t>(qbytes, 8)));
  FSR_CUDA(cudaMemcpyAsync(dQ, Q, qbytes, cudaMemcpyHostToDevice, s));
  std::vector<double> tile;
  for (int t0 = 0; t0 < nsteps; t0 += p->step_tile) {
    int nt = std::min(p->step_tile, nsteps - t0);
    int nsteps_pad = round_up(nt, 64);
    if ((rc = launch_pack_q(p, dQ + (size_t)t0 * ldq, ldq, nt, nsteps_pad, s))) { cudaFree(dQ); return rc; }
    if ((rc = launch_k1(p, nsteps_pad, s))) { cudaFree(dQ); return rc; }
    // U is [dof][t]; hand back step-major [t][dof]
    tile.resize((size_t)p->ndof * nt);   // only the nt live columns of every DOF row cross the bus
    FSR_CUDA(cudaMemcpy2DAsync(tile.data(), sizeof(double) * nt, p->U, sizeof(double) * p->step_tile, sizeof(double) * nt, (size_t)p->ndof,
                               cudaMemcpyDeviceToHost, s));
    FSR_CUDA(cudaStreamSynchronize(s));
    for (int t = 0; t < nt; ++t)
      for (int d = 0; d < p->ndof; ++d) U_host[(size_t)(t0 + t) * p->ndof + d] = tile[(size_t)d * nt + t];
  }
  cudaFree(dQ);
  return FSR_OK;
}

// out[t][k] = U[rows[k]][t]: the expanded displacements of a few DOFs only (rosette nodes, monitored nodes)
__global__ void gather_rows_kernel(const double* __restrict__ U, size_t ldu, const int* __restrict__ rows, int nrows, int nt,
                                   double* __restrict__ out)
{
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)nrows * nt) return;
  const int t = (int)(i % nt), k = (int)(i / nt);      // t fastest: coalesced reads of U
  out[(size_t)t * nrows + k] = U[(size_t)rows[k] * ldu + t];
}

int fsr_expand_rows(fsr_part* p, const double* Q, int ldq, int nsteps, const int* rows, int nrows, double* out)
{
  if (!p || !Q || nsteps < 0 || ldq < p->ndim || nrows < 0 || (nrows > 0 && (!rows || !out))) { set_error("fsr_expand_rows: bad arguments"); return FSR_ERR_ARG; }
  if (!p->have_R) { set_error("fsr_expand_rows: call fsr_set_recovery first"); return FSR_ERR_STATE; }
  for (int k = 0; k < nrows; ++k)
    if (rows[k] < 0 || rows[k] >= p->ndof) { set_error("fsr_expand_rows: DOF %d out of range (0..%d)", rows[k], p->ndof - 1); return FSR_ERR_ARG; }
  if (nrows == 0 || nsteps == 0) return FSR_OK;
  FSR_CUDA(cudaSetDevice(p->device));
  int rc = ensure_batch_buffers(p, false);
  if (rc) return rc;
  cudaStream_t s = p->stream;
  const size_t qbytes = sizeof(double) * (size_t)ldq * nsteps;
  double *dQ = nullptr, *dout = nullptr;
  int* drows = nullptr;
  auto done = [&](int code) { cudaFree(dQ); cudaFree(dout); cudaFree(drows); return code; };
  if (cudaMalloc(&dQ, qbytes) != cudaSuccess || cudaMalloc(&drows, sizeof(int) * nrows) != cudaSuccess ||
      cudaMalloc(&dout, sizeof(double) * (size_t)nrows * p->step_tile) != cudaSuccess) { set_error("fsr_expand_rows: device allocation failed"); return done(FSR_ERR_ALLOC); }
  cudaMemcpyAsync(dQ, Q, qbytes, cudaMemcpyHostToDevice, s);
  cudaMemcpyAsync(drows, rows, sizeof(int) * nrows, cudaMemcpyHostToDevice, s);
  for (int t0 = 0; t0 < nsteps; t0 += p->step_tile) {
    const int nt = std::min(p->step_tile, nsteps - t0);
    if ((rc = launch_pack_q(p, dQ + (size_t)t0 * ldq, ldq, nt, round_up(nt, 64), s))) return done(rc);
    if ((rc = launch_k1(p, round_up(nt, 64), s))) return done(rc);
    const size_t n = (size_t)nrows * nt;
    gather_rows_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(p->U, (size_t)p->step_tile, drows, nrows, nt, dout);
    ++g_launches;
    if (cudaMemcpyAsync(out + (size_t)t0 * nrows, dout, sizeof(double) * n, cudaMemcpyDeviceToHost, s) != cudaSuccess ||
        cudaStreamSynchronize(s) != cudaSuccess) { set_error("fsr_expand_rows: %s", cudaGetErrorString(cudaGetLastError())); return done(FSR_ERR_CUDA); }
  }
  return done(FSR_OK);
}

int fsr_recover_step_full(fsr_part* p, const double* q, double* resmat, double* stress, double* strain,
                          double* sres, double* sv)
{
  if (!p || !q) { set_error("fsr_recover_step_full: bad arguments"); return FSR_ERR_ARG; }
  if (!p->have_R) { set_error("fsr_recover_step_full: call fsr_set_recovery first"); return FSR_ERR_STATE; }
  FSR_CUDA(cudaSetDevice(p->device));
  int rc = ensure_batch_buffers(p, false);
  if (rc) return rc;
  cudaStream_t s = p->stream;
  double* dq = nullptr;
  FSR_CUDA(cudaMalloc(&dq, sizeof(double) * p->ndim));
  FSR_CUDA(cudaMemcpyAsync(dq, q, sizeof(double) * p->ndim, cudaMemcpyHostToDevice, s));
  if ((rc = launch_pack_q(p, dq, p->ndim, 1, 64, s))) { cudaFree(dq); return rc; }
  if ((rc = launch_k1(p, 64, s))) { cudaFree(dq); return rc; }
  size_t np = (size_t)std::max(p->npts, 1);
  double *d_res = nullptr, *d_sig = nullptr, *d_eps = nullptr, *d_sr = nullptr;
  FSR_CUDA(cudaMalloc(&d_res, sizeof(double) * 8 * np));
  FSR_CUDA(cudaMalloc(&d_sig, sizeof(double) * 6 * np));
  FSR_CUDA(cudaMalloc(&d_eps, sizeof(double) * 6 * np));
  FSR_CUDA(cudaMalloc(&d_sr, sizeof(double) * 24 * (size_t)std::max(p->nel, 1)));
  FSR_CUDA(cudaMemsetAsync(d_res, 0, sizeof(double) * 8 * np, s));
  FSR_CUDA(cudaMemsetAsync(d_sig, 0, sizeof(double) * 6 * np, s));
  FSR_CUDA(cudaMemsetAsync(d_eps, 0, sizeof(double) * 6 * np, s));
  FSR_CUDA(cudaMemsetAsync(d_sr, 0, sizeof(double) * 24 * (size_t)std::max(p->nel, 1), s));
  rc = launch_k2_full(p, d_res, d_sig, d_eps, d_sr, s);
  if (rc == FSR_OK) {
    if (resmat) cudaMemcpyAsync(resmat, d_res, sizeof(double) * 8 * p->npts, cudaMemcpyDeviceToHost, s);
    if (stress) cudaMemcpyAsync(stress, d_sig, sizeof(double) * 6 * p->npts, cudaMemcpyDeviceToHost, s);
    if (strain) cudaMemcpyAsync(strain, d_eps, sizeof(double) * 6 * p->npts, cudaMemcpyDeviceToHost, s);
    if (sres) cudaMemcpyAsync(sres, d_sr, sizeof(double) * 24 * p->nel, cudaMemcpyDeviceToHost, s);
    if (sv) {
      // column t = 0 of U[dof][t]
      cudaMemcpy2DAsync(sv, sizeof(double), p->U, sizeof(double) * p->step_tile, sizeof(double), p->ndof,
                        cudaMemcpyDeviceToHost, s);
    }
    if (cudaStreamSynchronize(s) != cudaSuccess) { set_error("fsr_recover_step_full: %s", cudaGetErrorString(cudaGetLastError())); rc = FSR_ERR_CUDA; }
  }
  cudaFree(dq); cudaFree(d_res); cudaFree(d_sig); cudaFree(d_eps); cudaFree(d_sr);
  return rc;
}

int fsr_vms_size(const fsr_part* p)
{
  if (!p) return FSR_ERR_ARG;
  int n = 0;
  for (int e = 0; e < p->nel; ++e) {
    const int nstrp = p->ptoff_host[e + 1] - p->ptoff_host[e];
    if (nstrp > 0) n += 3 + nstrp;
  }
  return n;
}

int fsr_get_vms(fsr_part* p, const double* q, double* vms, int nvms)
{
  if (!p || !q || !vms) { set_error("fsr_get_vms: bad arguments"); return FSR_ERR_ARG; }
  if (nvms < fsr_vms_size(p)) { set_error("fsr_get_vms: array too small (%d < %d)", nvms, fsr_vms_size(p)); return FSR_ERR_ARG; }
  std::vector<double> vm((size_t)std::max(p->npts, 1));
  // the envelope belongs to the batched history: keep it untouched by this single-step query
  std::vector<double> emax((size_t)std::max(p->npts, 1)), emin((size_t)std::max(p->npts, 1));
  int rc = fsr_get_envelope(p, emax.data(), emin.data());
  if (rc < 0) return rc;
  rc = fsr_recover(p, q, p->ndim, 1, vm.data());
  if (rc < 0) return rc;
  FSR_CUDA(cudaMemcpy(p->env_max, emax.data(), sizeof(double) * p->npts, cudaMemcpyHostToDevice));
  FSR_CUDA(cudaMemcpy(p->env_min, emin.data(), sizeof(double) * p->npts, cudaMemcpyHostToDevice));
  size_t k = 0;
  for (int e = 0; e < p->nel; ++e) {
    const int nstrp = p->ptoff_host[e + 1] - p->ptoff_host[e];
    if (nstrp <= 0) continue;
    vms[k++] = (double)(e + 1);
    vms[k++] = (double)p->nenod_host[e];
    vms[k++] = (double)nstrp;
    for (int i = 0; i < nstrp; ++i) vms[k++] = vm[(size_t)p->ptoff_host[e] + i];
  }
  return FSR_OK;
}

// counts[3 * f + 0..2] = elements of family f (enum Family of common.cuh: 0 quad, 1 tri, 2 TET10, 3 beam, 4 HEX20, 5 HEX8,
// 6 TET4, 7 WEDG6, 8 WEDG15, 9 TRI6, 10 QUAD8), of which on the geometry fast path (flat quads, straight-sided TET10), of
// which on the general kernel.  Returns the number of families.
int fsr_family_counts(const fsr_part* p, int* counts, int cap)
{
  if (!p || !counts) { set_error("fsr_family_counts: bad arguments"); return FSR_ERR_ARG; }
  for (int f = 0; f < FAM_COUNT && 3 * f + 2 < cap; ++f) {
    const FamilyData& fd = p->fam[f];
    const bool split = fd.nsub[0] + fd.nsub[1] + fd.nsub[2] > 0;
    counts[3 * f] = fd.nelt;
    const int fast = f == FAM_QUAD ? fd.nsub[0] + fd.nsub[2] : fd.nsub[0];   // quads: flat, on global or in-plane rows
    counts[3 * f + 1] = split ? fast : 0;
    counts[3 * f + 2] = split ? fd.nsub[0] + fd.nsub[1] + fd.nsub[2] - fast : fd.nelt;
  }
  return FAM_COUNT;
}

// What the von Mises path of the part expands and where its quadrilaterals go: info[0] = rows of K1 per step tile (all nodal
// DOFs, or with in-plane rows of flat shell regions: those rows + the 128-row tiles of R still read), [1] = nodal DOFs,
// [2] = in-plane rows, [3] = row tiles of R still expanded, [4] = quadrilaterals in the in-plane form, [5] = flat
// quadrilaterals on global rows, [6] = quadrilaterals on the dense operator.  Returns the number of entries filled.
int fsr_vm_path_info(const fsr_part* p, long long* info, int cap)
{
  if (!p || !info) { set_error("fsr_vm_path_info: bad arguments"); return FSR_ERR_ARG; }
  const FamilyData& q = p->fam[FAM_QUAD];
  const bool split = q.nsub[0] + q.nsub[1] + q.nsub[2] > 0;
  const long long v[7] = {p->planar ? (long long)p->np_rows_pad + 128LL * p->n_k1_tiles : (long long)p->nrows_pad, p->ndof,
                          p->planar ? p->np_rows : 0, p->planar ? p->n_k1_tiles : p->nrows_pad / 128, q.nsub[2], q.nsub[0],
                          split ? q.nsub[1] : q.nelt};
  const int n = std::min(cap, 7);
  for (int i = 0; i < n; ++i) info[i] = v[i];
  return n;
}

int fsr_last_timing(fsr_part* p, double* t_ms, int n)
{
  if (!p || !t_ms) return FSR_ERR_ARG;
  cudaSetDevice(p->device);
  double v[3] = {0.0, 0.0, (double)p->ntimed};
  for (int k = 0; k < p->ntimed; ++k) {
    cudaEvent_t* ev = p->evring[k];
    if (cudaEventSynchronize(ev[2]) != cudaSuccess) { set_error("fsr_last_timing: event sync failed"); return FSR_ERR_CUDA; }
    float a = 0, b = 0;
    cudaEventElapsedTime(&a, ev[0], ev[1]);
    cudaEventElapsedTime(&b, ev[1], ev[2]);
    v[0] += a; v[1] += b;
  }
  int m = std::min(n, 3);
  for (int i = 0; i < m; ++i) t_ms[i] = v[i];
  return m;
}

int fsr_timing_reset(fsr_part* p)
{
  if (!p) return FSR_ERR_ARG;
  p->ntimed = 0;
  return FSR_OK;
}

}  // extern "C"
