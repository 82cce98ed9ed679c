// k3_fatigue.cu -- K3 on sm_100a: gated peak-valley extraction, rainflow counting, Miner damage and
// cycle histograms for many independent scalar histories (strain-gage legs and rosette max
// principal stresses), one GPU thread per history, STREAMING over tiles of time steps.
//
// Replaces ffp_addpoint / ffp_getdamage / ffp_getnumcycles
// (fedem-foundation/src/FFpLib/FFpFatigue/FFpFatigue_F.C:37-141) and the C++ behind them
// (FFpFatigue.C:77-320,381-396, FFpSNCurve.C:10-33); the per-series state machines are in
// fatigue_core.cuh.  The reference appends every sample to a std::vector per gage and processes at
// the end; here nothing but a ~160-byte state and the rainflow residue stack is kept per gage, so
// 4e5 histories x 1e5 steps (config 5) never exist in memory at once.
//
// Data movement: a history tile is read exactly once per pass.  Step-major tiles (hist[t*ld + g],
// what the rosette kernel writes) are read coalesced by consecutive threads; gage-major tiles
// (hist[g*ld + t], what a host caller holds) are staged through shared memory with cp.async
// (LDGSTS) as 32-step x 128-gage blocks, each gage row a 256-byte segment (one buffer per block, five blocks per SM), so
// that global reads stay coalesced although every thread walks its own row.  The work per sample is
// a divergent state machine: the kernel is latency / issue bound, not HBM bound (8 B per sample).
#include <algorithm>

#include "common.cuh"
#include "fatigue_core.cuh"

namespace fsr {

struct GageState {
  PvxLocate loc;
  PvxStream pv;
  Rainflow rf;
  CycleSink sink;
  int status;  // 0 ok, 1 = reference's closure failure, 2 = residue stack overflow
  int done;    // finish already applied
};

constexpr int K3_THREADS = 128;
constexpr int K3_CHUNK = 32;   // steps staged per pass of a block (FSR_K3_CHUNK=16: half the tile, more blocks per SM)
constexpr int K3_TPCAP = 8;  // turning points a gage may queue between two convergent rainflow passes
constexpr int K3_QCAP = 4;   // closed cycles a gage may queue between two convergent damage evaluations

}  // namespace fsr

struct fsr_fatigue_state {
  int device = 0;
  int ngage = 0, nbins = 0, cap = 0;
  double bin_size = 0.0;
  fsr::GageState* st = nullptr;   // [ngage]
  double* gate = nullptr;         // [ngage]
  double* curve = nullptr;        // [ngage][5]: loga1, loga2, m1, m2, logN0
  double* edges = nullptr;        // [nbins+2]
  double* spillA = nullptr;       // [cap][ngage]
  double* spillB = nullptr;       // [cap+4][ngage], allocated by finish
  int* bins = nullptr;            // [nbins][ngage]
  int* pending = nullptr;         // device counter of gages without a first turning point
  double* out_damage = nullptr;   // [ngage] finish outputs
  int* out_ncyc = nullptr;
  int* out_status = nullptr;
  int* out_bins = nullptr;        // [ngage][nbins] gage-major for the caller
  cudaStream_t stream = nullptr;
};

namespace fsr {

__device__ __forceinline__ void cp_async8(void* smem, const void* gmem)
{
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__global__ void k3_init_kernel(GageState* st, int ngage, int* bins, int nbins)
{
  int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= ngage) return;
  GageState s;
  s.loc.init(); s.pv.init(); s.rf.init(); s.sink.init();
  s.status = 0; s.done = 0;
  st[g] = s;
  for (int k = 0; k < nbins; ++k) bins[(size_t)k * ngage + g] = 0;
}

// Stages the CHUNK-step x 128-gage block starting at (g0, t0) of a gage-major history into
// tile[step][gage] (row padded to 129 doubles: conflict-free both ways).  A warp instruction copies CHUNK consecutive
// steps of 32 / CHUNK gages.
template <int CHUNK>
__device__ __forceinline__ void stage_tile(double* tile, const double* __restrict__ hist, size_t ld, int g0,
                                           int ngage, int t0, int t1)
{
  constexpr int GPI = 32 / CHUNK;   // gages per instruction
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int ts = lane & (CHUNK - 1), t = t0 + ts, sub = lane / CHUNK;
#pragma unroll 8
  for (int r = 0; r < 32 / GPI; ++r) {
    const int gl = warp * 32 + r * GPI + sub, g = g0 + gl;
    if (g < ngage && t < t1) cp_async8(tile + ts * (K3_THREADS + 1) + gl, hist + (size_t)g * ld + t);
  }
}

// MODE 0: locate the first turning point (early exit once every gage of the block has one);
// MODE 1: PVX main loop + rainflow + damage.  hist points at the sample of global step `step0`.
template <int LAYOUT, int MODE, int CHUNK>
__global__ void __launch_bounds__(K3_THREADS)
k3_stream_kernel(GageState* __restrict__ st, const double* __restrict__ hist, size_t ld, int ngage, int step0,
                 int nsteps, const double* __restrict__ gate_g, const double* __restrict__ curve_g,
                 const double* __restrict__ edges, double bin_size, int nbins, double* __restrict__ spill, int cap,
                 int* __restrict__ bins, int* __restrict__ pending)
{
  extern __shared__ double tiles[];  // LAYOUT 0: [32][129]
  const int g0 = blockIdx.x * K3_THREADS;
  const int g = g0 + threadIdx.x;
  const bool active = g < ngage;
  GageState s;
  FatigueParams p;
  p.gate = 0.0; p.loga1 = p.loga2 = p.m1 = p.m2 = p.logN0 = 0.0; p.bin_size = bin_size; p.nbins = nbins;
  if (active) {
    s = st[g];
    p.gate = gate_g[g];
    if (MODE == 1) {
      p.loga1 = curve_g[5 * (size_t)g]; p.loga2 = curve_g[5 * (size_t)g + 1]; p.m1 = curve_g[5 * (size_t)g + 2];
      p.m2 = curve_g[5 * (size_t)g + 3]; p.logN0 = curve_g[5 * (size_t)g + 4];
    }
  }
  bool idle = !active || s.status == 2 || s.done;
  if (MODE == 0) idle = idle || s.loc.first >= 0;
  if (MODE == 1) idle = idle || s.loc.first < 0 || s.loc.first >= step0 + nsteps;
  const bool was_pending = MODE == 0 && !idle;
  double* myspill = spill + (active ? g : 0);
  int* mybins = (bins && nbins > 0) ? bins + (active ? g : 0) : nullptr;
  const size_t stride = (size_t)ngage;

  // Closed cycles are not evaluated where the rainflow rules find them: the S-N damage (log10 + pow) and the histogram
  // search are by far the most expensive part of a sample, and inside the state machine they would run once per step in
  // which ANY lane of the warp closes a cycle (~80 % of the steps at one cycle per 20 samples).  They are queued per gage
  // (FIFO, so the order of the Miner sum is unchanged) and evaluated warp-convergently once per 32-step chunk: the warp
  // then pays max-over-lanes evaluations per chunk instead of one per step.
  __shared__ double q_a[MODE == 1 ? K3_QCAP : 1][K3_THREADS], q_b[MODE == 1 ? K3_QCAP : 1][K3_THREADS];
  int qn = 0;
  auto count = [&](double a, double b) {
    if (qn == K3_QCAP) {   // more than K3_QCAP cycles since the last flush: drain in order (divergent, rare)
      for (int k = 0; k < K3_QCAP; ++k) count_cycle(q_a[k][threadIdx.x], q_b[k][threadIdx.x], p, s.sink, mybins, stride, edges);
      qn = 0;
    }
    q_a[qn][threadIdx.x] = a; q_b[qn][threadIdx.x] = b; ++qn;
  };
  auto flush = [&]() {     // reached by every lane of the warp
    int m = qn;
#pragma unroll
    for (int o = 16; o; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    for (int k = 0; k < m; ++k)
      if (k < qn) count_cycle(q_a[k][threadIdx.x], q_b[k][threadIdx.x], p, s.sink, mybins, stride, edges);
    qn = 0;
  };
  // The same separation one level up: the peak-valley filter only queues its turning points (FIFO per gage); the
  // rainflow stack consumes them at the end of the chunk in a loop every lane of the warp runs together, so the three
  // state machines (PVX, rainflow, damage) are three tight loops instead of one interleaved, divergent one.
  __shared__ double q_tp[MODE == 1 ? K3_TPCAP : 1][K3_THREADS];
  int qt = 0;
  auto emit = [&](double v) {
    if (qt == K3_TPCAP) {   // queue full: drain in order (divergent, rare)
      for (int k = 0; k < K3_TPCAP; ++k) s.rf.push(q_tp[k][threadIdx.x], p.gate, myspill, stride, cap, count);
      qt = 0;
    }
    q_tp[qt][threadIdx.x] = v; ++qt;
  };
  auto drain = [&]() {      // reached by every lane of the warp
    int m = qt;
#pragma unroll
    for (int o = 16; o; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    for (int k = 0; k < m; ++k)
      if (k < qt) s.rf.push(q_tp[k][threadIdx.x], p.gate, myspill, stride, cap, count);
    qt = 0;
    flush();
    if (s.rf.overflow) { s.status = 2; idle = true; }
  };
  auto consume = [&](int i, double x) {
    if (MODE == 0) {
      if (s.loc.feed(x, p.gate)) idle = true;
    } else {
      s.pv.feed(i, s.loc.first, x, p.gate, emit);
      if (s.rf.overflow) { s.status = 2; idle = true; }
    }
  };

  if (LAYOUT == 0) {
    const int nchunks = (nsteps + CHUNK - 1) / CHUNK;
    // One staging buffer per block: the kernel is latency bound (a divergent state machine per lane), so the shared
    // memory goes to occupancy -- five blocks per SM hide each other's staging waits -- rather than to double buffering
    // inside a block (r01l profile: 2 blocks/SM, 12 % of the warp slots active with two 33 KB buffers).
    for (int c = 0; c < nchunks; ++c) {
      stage_tile<CHUNK>(tiles, hist, ld, g0, ngage, c * CHUNK, nsteps);
      cp_async_commit();
      cp_async_wait<0>();
      __syncthreads();
      const int tn = min(CHUNK, nsteps - c * CHUNK);
      const double* col = tiles + threadIdx.x;
      if (!idle)
        for (int k = 0; k < tn && !idle; ++k) consume(step0 + c * CHUNK + k, col[k * (K3_THREADS + 1)]);
      if (MODE == 1) drain();
      // every gage of the block located: nothing left to read in this pass
      if (MODE == 0 && __syncthreads_and(idle)) break;
      if (MODE != 0) __syncthreads();
    }
    cp_async_wait<0>();
  } else {
    const double* hp = hist + (active ? g : 0);
    int t = 0;
    for (; t + 8 <= nsteps; t += 8) {   // warp-uniform trip count: the flush below needs every lane
      if (!idle) {
        double xb[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) xb[k] = __ldg(hp + (size_t)(t + k) * ld);
#pragma unroll
        for (int k = 0; k < 8; ++k)
          if (!idle) consume(step0 + t + k, xb[k]);
      }
      if (MODE == 1 && (t & 31) == 24) drain();
      if (MODE == 0 && __all_sync(0xffffffffu, idle)) break;
    }
    if (!idle)
      for (; t < nsteps && !idle; ++t) consume(step0 + t, __ldg(hp + (size_t)t * ld));
    if (MODE == 1) drain();
  }
  if (active && (MODE == 1 || was_pending)) st[g] = s;
  if (MODE == 0 && pending) {
    // gages still without a first turning point after this tile
    const int still = (active && s.status != 2 && !s.done && s.loc.first < 0) ? 1 : 0;
    const int cnt = __syncthreads_count(still);
    if (threadIdx.x == 0 && cnt) atomicAdd(pending, cnt);
  }
}

// End of data: last possible turning point, residue closure, results.
__global__ void __launch_bounds__(K3_THREADS)
k3_finish_kernel(GageState* __restrict__ st, int ngage, const double* __restrict__ gate_g,
                 const double* __restrict__ curve_g, const double* __restrict__ edges, double bin_size, int nbins,
                 double* __restrict__ spillA, double* __restrict__ spillB, int cap, int* __restrict__ bins,
                 double* __restrict__ damage, int* __restrict__ ncyc, int* __restrict__ status,
                 int* __restrict__ bins_out)
{
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= ngage) return;
  GageState s = st[g];
  FatigueParams p;
  p.gate = gate_g[g];
  p.loga1 = curve_g[5 * (size_t)g]; p.loga2 = curve_g[5 * (size_t)g + 1]; p.m1 = curve_g[5 * (size_t)g + 2];
  p.m2 = curve_g[5 * (size_t)g + 3]; p.logN0 = curve_g[5 * (size_t)g + 4];
  p.bin_size = bin_size; p.nbins = nbins;
  const size_t stride = (size_t)ngage;
  int* mybins = (bins && nbins > 0) ? bins + g : nullptr;
  auto count = [&](double a, double b) { count_cycle(a, b, p, s.sink, mybins, stride, edges); };
  if (!s.done && s.status != 2) {
    auto emit = [&](double v) { s.rf.push(v, p.gate, spillA + g, stride, cap, count); };
    s.pv.finish(p.gate, emit);
    if (s.rf.overflow) s.status = 2;
    else if (!rainflow_finish(s.rf, p.gate, spillA + g, spillB + g, stride, count)) s.status = 1;
    s.done = 1;
    st[g] = s;
  }
  damage[g] = s.status == 2 ? -1.0 : s.sink.damage;
  ncyc[g] = s.status == 2 ? -1 : s.sink.ncycles;
  status[g] = s.status;
  if (bins_out)
    for (int k = 0; k < nbins; ++k) {
      // ffp_getnumcycles: -1 when there are no cycles or the bin starts beyond the largest range
      int v = mybins[(size_t)k * stride];
      if (s.status == 2 || s.sink.ncycles == 0 || edges[k] > s.sink.max_range) v = -1;
      bins_out[(size_t)g * nbins + k] = v;
    }
}

static size_t k3_smem(int layout, int chunk) { return layout == 0 ? sizeof(double) * chunk * (K3_THREADS + 1) : 0; }

template <int MODE>
static int launch_stream(fsr_fatigue_state* f, const double* hist, size_t ld, int layout, int step0, int nsteps,
                         cudaStream_t s)
{
  if (nsteps <= 0 || f->ngage == 0) return FSR_OK;
  const unsigned blocks = (unsigned)((f->ngage + K3_THREADS - 1) / K3_THREADS);
  static const int chunk = (getenv("FSR_K3_CHUNK") && atoi(getenv("FSR_K3_CHUNK")) == 32) ? 32 : 16;
  if (layout == 0 && chunk == 16) {
    if (int rc = smem_opt_in((const void*)k3_stream_kernel<0, MODE, 16>, k3_smem(0, 16))) return rc;
    k3_stream_kernel<0, MODE, 16><<<blocks, K3_THREADS, k3_smem(0, 16), s>>>(f->st, hist, ld, f->ngage, step0, nsteps, f->gate,
                                                                             f->curve, f->edges, f->bin_size, f->nbins,
                                                                             f->spillA, f->cap, f->bins, f->pending);
  } else if (layout == 0) {
    if (int rc = smem_opt_in((const void*)k3_stream_kernel<0, MODE, 32>, k3_smem(0, 32))) return rc;
    k3_stream_kernel<0, MODE, 32><<<blocks, K3_THREADS, k3_smem(0, 32), s>>>(f->st, hist, ld, f->ngage, step0, nsteps, f->gate,
                                                                             f->curve, f->edges, f->bin_size, f->nbins,
                                                                             f->spillA, f->cap, f->bins, f->pending);
  } else
    k3_stream_kernel<1, MODE, 32><<<blocks, K3_THREADS, 0, s>>>(f->st, hist, ld, f->ngage, step0, nsteps, f->gate, f->curve,
                                                                f->edges, f->bin_size, f->nbins, f->spillA, f->cap, f->bins,
                                                                f->pending);
  FSR_LAUNCH_CHECK();
  return FSR_OK;
}

}  // namespace fsr

using namespace fsr;

extern "C" {

void fsr_fatigue_destroy(fsr_fatigue_state* f)
{
  if (!f) return;
  cudaSetDevice(f->device);
  if (f->stream) { cudaStreamSynchronize(f->stream); cudaStreamDestroy(f->stream); }
  cudaFree(f->st); cudaFree(f->gate); cudaFree(f->curve); cudaFree(f->edges); cudaFree(f->spillA);
  cudaFree(f->spillB); cudaFree(f->bins); cudaFree(f->pending); cudaFree(f->out_damage); cudaFree(f->out_ncyc);
  cudaFree(f->out_status); cudaFree(f->out_bins);
  delete f;
}

int fsr_fatigue_reset(fsr_fatigue_state* f)
{
  if (!f) { set_error("fsr_fatigue_reset: null handle"); return FSR_ERR_ARG; }
  FSR_CUDA(cudaSetDevice(f->device));
  if (f->ngage > 0) {
    k3_init_kernel<<<(f->ngage + 255) / 256, 256, 0, f->stream>>>(f->st, f->ngage, f->bins, f->nbins);
    FSR_LAUNCH_CHECK();
  }
  FSR_CUDA(cudaStreamSynchronize(f->stream));
  return FSR_OK;
}

int fsr_fatigue_set_gage_params(fsr_fatigue_state* f, const double* gate, const double* curve)
{
  if (!f) { set_error("fsr_fatigue_set_gage_params: null handle"); return FSR_ERR_ARG; }
  FSR_CUDA(cudaSetDevice(f->device));
  if (gate) FSR_CUDA(cudaMemcpy(f->gate, gate, sizeof(double) * f->ngage, cudaMemcpyHostToDevice));
  if (curve) {
    std::vector<double> c5((size_t)5 * f->ngage);
    for (int g = 0; g < f->ngage; ++g) {
      const double* c = curve + 4 * (size_t)g;
      if (c[3] == c[2] && c[0] != c[1]) { set_error("gage %d: S-N slopes m1 == m2", g); return FSR_ERR_ARG; }
      for (int k = 0; k < 4; ++k) c5[5 * (size_t)g + k] = c[k];
      // FFpSNCurve.C:12-15; twice the same line = a one-segment curve of the S-N library: always the second segment
      c5[5 * (size_t)g + 4] = c[3] == c[2] ? -kHuge : (c[3] * c[0] - c[2] * c[1]) / (c[3] - c[2]);
    }
    FSR_CUDA(cudaMemcpy(f->curve, c5.data(), sizeof(double) * c5.size(), cudaMemcpyHostToDevice));
  }
  return FSR_OK;
}

int fsr_fatigue_create(fsr_fatigue_state** out, int device, int ngage, double gate, const double* curve, double bin_size,
                       int nbins, int stack_cap)
{
  if (!out || ngage < 0 || !curve || nbins < 0 || (nbins > 0 && !(bin_size > 0.0))) {
    set_error("fsr_fatigue_create: bad arguments");
    return FSR_ERR_ARG;
  }
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) {
    set_error("no CUDA device available: this library has no CPU fallback");
    return FSR_ERR_CUDA;
  }
  if (device < 0 || device >= ndev) { set_error("device %d out of range", device); return FSR_ERR_ARG; }
  FSR_CUDA(cudaSetDevice(device));
  fsr_fatigue_state* f = new fsr_fatigue_state();
  f->device = device; f->ngage = ngage; f->nbins = nbins; f->bin_size = bin_size;
  f->cap = stack_cap > 0 ? stack_cap : 1024;
  const size_t ng = (size_t)std::max(ngage, 1);
  bool ok = cudaStreamCreateWithFlags(&f->stream, cudaStreamNonBlocking) == cudaSuccess &&
            cudaMalloc(&f->st, sizeof(GageState) * ng) == cudaSuccess &&
            cudaMalloc(&f->gate, sizeof(double) * ng) == cudaSuccess &&
            cudaMalloc(&f->curve, sizeof(double) * 5 * ng) == cudaSuccess &&
            cudaMalloc(&f->edges, sizeof(double) * (nbins + 2)) == cudaSuccess &&
            cudaMalloc(&f->spillA, sizeof(double) * ng * f->cap) == cudaSuccess &&
            cudaMalloc(&f->bins, sizeof(int) * ng * std::max(nbins, 1)) == cudaSuccess &&
            cudaMalloc(&f->pending, sizeof(int)) == cudaSuccess &&
            cudaMalloc(&f->out_damage, sizeof(double) * ng) == cudaSuccess &&
            cudaMalloc(&f->out_ncyc, sizeof(int) * ng) == cudaSuccess &&
            cudaMalloc(&f->out_status, sizeof(int) * ng) == cudaSuccess &&
            cudaMalloc(&f->out_bins, sizeof(int) * ng * std::max(nbins, 1)) == cudaSuccess;
  if (!ok) {
    set_error("fsr_fatigue_create: device allocation failed (%d gages, stack %d): %s", ngage, f->cap,
              cudaGetErrorString(cudaGetLastError()));
    fsr_fatigue_destroy(f);
    return FSR_ERR_ALLOC;
  }
  std::vector<double> edges((size_t)nbins + 2, 0.0);
  for (int k = 1; k <= nbins + 1; ++k) edges[k] = edges[k - 1] + bin_size;  // s0 = s1; s1 = s0 + binSize
  std::vector<double> g(ng, gate), c((size_t)4 * ng);
  for (size_t i = 0; i < ng; ++i) for (int k = 0; k < 4; ++k) c[4 * i + k] = curve[k];
  int rc = FSR_OK;
  if (cudaMemcpy(f->edges, edges.data(), sizeof(double) * edges.size(), cudaMemcpyHostToDevice) != cudaSuccess) rc = FSR_ERR_CUDA;
  if (rc == FSR_OK && ngage > 0) rc = fsr_fatigue_set_gage_params(f, g.data(), c.data());
  if (rc == FSR_OK) rc = fsr_fatigue_reset(f);
  if (rc != FSR_OK) { fsr_fatigue_destroy(f); return rc; }
  *out = f;
  return FSR_OK;
}

int fsr_fatigue_locate_dev(fsr_fatigue_state* f, const double* hist_dev, size_t ld, int layout, int step0, int nsteps,
                           int* n_pending, void* stream)
{
  if (!f || !hist_dev || nsteps < 0 || (layout != 0 && layout != 1)) { set_error("fsr_fatigue_locate_dev: bad arguments"); return FSR_ERR_ARG; }
  FSR_CUDA(cudaSetDevice(f->device));
  cudaStream_t s = stream ? (cudaStream_t)stream : f->stream;
  FSR_CUDA(cudaMemsetAsync(f->pending, 0, sizeof(int), s));
  int rc = launch_stream<0>(f, hist_dev, ld, layout, step0, nsteps, s);
  if (rc) return rc;
  if (n_pending) {
    FSR_CUDA(cudaMemcpyAsync(n_pending, f->pending, sizeof(int), cudaMemcpyDeviceToHost, s));
    FSR_CUDA(cudaStreamSynchronize(s));
  }
  return FSR_OK;
}

int fsr_fatigue_feed_dev(fsr_fatigue_state* f, const double* hist_dev, size_t ld, int layout, int step0, int nsteps,
                         void* stream)
{
  if (!f || !hist_dev || nsteps < 0 || (layout != 0 && layout != 1)) { set_error("fsr_fatigue_feed_dev: bad arguments"); return FSR_ERR_ARG; }
  FSR_CUDA(cudaSetDevice(f->device));
  cudaStream_t s = stream ? (cudaStream_t)stream : f->stream;
  return launch_stream<1>(f, hist_dev, ld, layout, step0, nsteps, s);
}

int fsr_fatigue_finish_dev(fsr_fatigue_state* f, void* stream)
{
  if (!f) { set_error("fsr_fatigue_finish_dev: null handle"); return FSR_ERR_ARG; }
  FSR_CUDA(cudaSetDevice(f->device));
  cudaStream_t s = stream ? (cudaStream_t)stream : f->stream;
  if (!f->spillB) FSR_CUDA(cudaMalloc(&f->spillB, sizeof(double) * (size_t)std::max(f->ngage, 1) * (f->cap + 4)));
  if (f->ngage > 0) {
    k3_finish_kernel<<<(f->ngage + K3_THREADS - 1) / K3_THREADS, K3_THREADS, 0, s>>>(
        f->st, f->ngage, f->gate, f->curve, f->edges, f->bin_size, f->nbins, f->spillA, f->spillB, f->cap, f->bins,
        f->out_damage, f->out_ncyc, f->out_status, f->nbins > 0 ? f->out_bins : nullptr);
    FSR_LAUNCH_CHECK();
  }
  return FSR_OK;
}

int fsr_fatigue_results_dev(fsr_fatigue_state* f, double** damage_dev, int** ncycles_dev, int** bins_dev, int** status_dev)
{
  if (!f) return FSR_ERR_ARG;
  if (damage_dev) *damage_dev = f->out_damage;
  if (ncycles_dev) *ncycles_dev = f->out_ncyc;
  if (bins_dev) *bins_dev = f->out_bins;
  if (status_dev) *status_dev = f->out_status;
  return FSR_OK;
}

int fsr_fatigue_finish(fsr_fatigue_state* f, double* damage, int* ncycles, int* bins, int* status)
{
  int rc = fsr_fatigue_finish_dev(f, nullptr);
  if (rc) return rc;
  cudaStream_t s = f->stream;
  const size_t ng = (size_t)f->ngage;
  if (damage) FSR_CUDA(cudaMemcpyAsync(damage, f->out_damage, sizeof(double) * ng, cudaMemcpyDeviceToHost, s));
  if (ncycles) FSR_CUDA(cudaMemcpyAsync(ncycles, f->out_ncyc, sizeof(int) * ng, cudaMemcpyDeviceToHost, s));
  if (status) FSR_CUDA(cudaMemcpyAsync(status, f->out_status, sizeof(int) * ng, cudaMemcpyDeviceToHost, s));
  if (bins && f->nbins > 0) FSR_CUDA(cudaMemcpyAsync(bins, f->out_bins, sizeof(int) * ng * f->nbins, cudaMemcpyDeviceToHost, s));
  FSR_CUDA(cudaStreamSynchronize(s));
  // warning count = gages whose closure failed like the reference's or whose stack overflowed -- also when the caller
  // did not ask for the status array
  std::vector<int> st;
  if (!status) {
    st.resize(ng);
    FSR_CUDA(cudaMemcpy(st.data(), f->out_status, sizeof(int) * ng, cudaMemcpyDeviceToHost));
    status = st.data();
  }
  int nwarn = 0;
  for (size_t g = 0; g < ng; ++g) nwarn += status[g] != 0;
  return nwarn;
}

int fsr_fatigue_dev(int device, const double* hist_dev, size_t ld_hist, int ngage, int nsteps, double gate,
                    const double* curve, double bin_size, int nbins, double* damage_dev, int* ncycles_dev,
                    int* bins_dev, void* stream)
{
  if (!hist_dev || ngage < 0 || nsteps < 0 || ld_hist < (size_t)nsteps) { set_error("fsr_fatigue_dev: bad arguments"); return FSR_ERR_ARG; }
  fsr_fatigue_state* f = nullptr;
  int rc = fsr_fatigue_create(&f, device, ngage, gate, curve, bin_size, nbins, std::min(nsteps + 8, 1 << 16));
  if (rc) return rc;
  cudaStream_t s = stream ? (cudaStream_t)stream : f->stream;
  rc = fsr_fatigue_locate_dev(f, hist_dev, ld_hist, 0, 0, nsteps, nullptr, s);
  if (!rc) rc = fsr_fatigue_feed_dev(f, hist_dev, ld_hist, 0, 0, nsteps, s);
  if (!rc) rc = fsr_fatigue_finish_dev(f, s);
  if (!rc) {
    cudaError_t e = cudaSuccess;
    if (damage_dev) e = cudaMemcpyAsync(damage_dev, f->out_damage, sizeof(double) * ngage, cudaMemcpyDeviceToDevice, s);
    if (e == cudaSuccess && ncycles_dev) e = cudaMemcpyAsync(ncycles_dev, f->out_ncyc, sizeof(int) * ngage, cudaMemcpyDeviceToDevice, s);
    if (e == cudaSuccess && bins_dev && nbins > 0)
      e = cudaMemcpyAsync(bins_dev, f->out_bins, sizeof(int) * (size_t)ngage * nbins, cudaMemcpyDeviceToDevice, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) { set_error("fsr_fatigue_dev: %s", cudaGetErrorString(e)); rc = FSR_ERR_CUDA; }
  }
  fsr_fatigue_destroy(f);
  return rc;
}

int fsr_fatigue(int device, const double* hist, int ngage, int nsteps, double gate, const double* curve,
                double bin_size, int nbins, double* damage, int* ncycles, int* bins)
{
  if (!hist || ngage < 0 || nsteps < 0 || !curve) { set_error("fsr_fatigue: bad arguments"); return FSR_ERR_ARG; }
  fsr_fatigue_state* f = nullptr;
  int rc = fsr_fatigue_create(&f, device, ngage, gate, curve, bin_size, nbins, std::min(nsteps + 8, 1 << 16));
  if (rc) return rc;
  // the host history goes through the device in windows of steps; two passes (locate, then count)
  const size_t budget = (size_t)1 << 30;  // bytes per window
  int win = (int)std::min<size_t>((size_t)std::max(nsteps, 1), std::max<size_t>(budget / (sizeof(double) * std::max(ngage, 1)), 64));
  double* d = nullptr;
  if (cudaMalloc(&d, sizeof(double) * (size_t)std::max(ngage, 1) * win) != cudaSuccess) {
    set_error("fsr_fatigue: device allocation failed");
    fsr_fatigue_destroy(f);
    return FSR_ERR_ALLOC;
  }
  cudaStream_t s = f->stream;
  for (int pass = 0; pass < 2 && rc == FSR_OK; ++pass)
    for (int t0 = 0; t0 < nsteps && rc == FSR_OK; t0 += win) {
      const int nt = std::min(win, nsteps - t0);
      if (cudaMemcpy2DAsync(d, sizeof(double) * win, hist + t0, sizeof(double) * nsteps, sizeof(double) * nt, ngage,
                            cudaMemcpyHostToDevice, s) != cudaSuccess) {
        set_error("fsr_fatigue: upload failed: %s", cudaGetErrorString(cudaGetLastError()));
        rc = FSR_ERR_CUDA;
        break;
      }
      if (pass == 0) {
        int pend = 0;
        rc = fsr_fatigue_locate_dev(f, d, win, 0, t0, nt, &pend, s);
        if (rc == FSR_OK && pend == 0) break;  // every gage has its first turning point
      } else
        rc = fsr_fatigue_feed_dev(f, d, win, 0, t0, nt, s);
    }
  std::vector<int> status((size_t)std::max(ngage, 1));
  if (rc == FSR_OK) rc = fsr_fatigue_finish(f, damage, ncycles, bins, status.data());
  cudaFree(d);
  fsr_fatigue_destroy(f);
  return rc;
}

}  // extern "C"
