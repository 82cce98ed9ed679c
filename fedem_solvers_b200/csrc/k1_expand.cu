// k1_expand.cu -- K1: expansion of the reduced history to nodal displacements on sm_100a.
//
// Replaces, for a whole batch of steps, calcIntDisplacements + disExpand
// (reference src/vpmStress/displacementModule.f90:931-1024,1226-1259) and the two dmMatTimesVec
// column-AXPY sweeps (src/vpmUtilities/diskMatrixModule.f90:993-1045).
//
//   build_row_operator : folds dofPosIn2 (samModule.f90:960-970), the meqn1/meqn2 scatters and
//                        the constraint equations of disExpand into ONE constant row operator
//                        R[ndof x ndim] in nodal DOF order.
//   k1_expand_kernel   : U[dof][t] = R[dof][:] . Q[:, t] as an FP64 tensor-core GEMM
//                        (mma.sync m8n8k4 f64 -> SASS DMMA.8x8x4; tcgen05 has no f64 kind),
//                        operands staged in shared memory by 1-D bulk TMA (cp.async.bulk ->
//                        SASS UBLKCP) completing on mbarriers, whole K resident.
//
// Layouts: R row-major [nrows_pad][ldk], Qt step-major [nsteps_pad][ldk], both zero padded,
// ldk == 4 (mod 8) so that the 8x4 fragment loads (row stride ldk doubles) hit 32 distinct
// banks per half-warp.  U row-major [nrows_pad][ldu] with t fastest: the K2 kernels read 8
// consecutive steps of one DOF as one 64-byte segment.
#include <algorithm>
#include <cstdlib>

#include "common.cuh"

namespace fsr {

// ------------------------------------------------------------------------------------------
// Row operator assembly
// ------------------------------------------------------------------------------------------

// One thread per (dof row, reduced column); rows fastest so that reads of the column-major
// B / E are coalesced.  rowptr/src/w: CSR over the sources of each DOF row; src >= 0 is a row
// of [B|E] (position in meqn1), src < 0 is external DOF j = -src-1 (position in meqn2).
__global__ void build_R_kernel(double* __restrict__ R, int ldk, int ndof, int ndim, int ndof2,
                               const double* __restrict__ B, size_t ldB,
                               const double* __restrict__ E, size_t ldE,
                               const int* __restrict__ rowptr, const int* __restrict__ src,
                               const double* __restrict__ w,
                               const int* __restrict__ bcol,   // [ndof2] column of B feeding finit(c)
                               const int* __restrict__ extcol) // [ndof2] finit index of external j
{
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t total = (size_t)ndof * ndim;
  if (idx >= total) return;
  int d = (int)(idx % ndof);
  int c = (int)(idx / ndof);
  double acc = 0.0;
  for (int ip = rowptr[d]; ip < rowptr[d + 1]; ++ip) {
    int s = src[ip];
    double v;
    if (s >= 0)
      v = (c < ndof2) ? B[(size_t)bcol[c] * ldB + s] : E[(size_t)(c - ndof2) * ldE + s];
    else
      v = (extcol[-s - 1] == c) ? 1.0 : 0.0;
    acc += w[ip] * v;
  }
  R[(size_t)d * ldk + c] = acc;
}

int build_row_operator(fsr_part* p, const fsr_sam* sam, const double* B, int ldB,
                       const double* E, int ldE)
{
  const int ndof = sam->ndof, ndof1 = sam->ndof1, ndof2 = sam->ndof2, neq = sam->neq;
  const int ngen = sam->ngen;

  // equation -> position in meqn1 (>=0) / meqn2 (encoded -j-1); INT_MIN-like = none
  const int NONE = 0x7fffffff;
  std::vector<int> eqsrc((size_t)neq + 1, NONE);
  for (int k = 0; k < ndof1; ++k) {
    int eq = sam->meqn1[k];
    if (eq < 1 || eq > neq) { set_error("meqn1(%d)=%d out of range", k + 1, eq); return FSR_ERR_ARG; }
    eqsrc[eq] = k;
  }
  for (int j = 0; j < ndof2; ++j) {
    int eq = sam->meqn2[j];
    if (eq < 1 || eq > neq) { set_error("meqn2(%d)=%d out of range", j + 1, eq); return FSR_ERR_ARG; }
    eqsrc[eq] = -j - 1;
  }
  // dofPosIn2 (samModule.f90:964-970): i2-th status-2 DOF in nodal order -> position in meqn2
  std::vector<int> bcol(ndof2 > 0 ? ndof2 : 1, 0), extcol(ndof2 > 0 ? ndof2 : 1, -1);
  {
    int i2 = 0;
    for (int d = 0; d < ndof; ++d)
      if (sam->msc[d] == 2) {
        if (i2 >= ndof2) { set_error("more status-2 DOFs than ndof2"); return FSR_ERR_ARG; }
        int eq = sam->meqn[d];
        int s = (eq >= 1 && eq <= neq) ? eqsrc[eq] : NONE;
        if (s == NONE || s >= 0) { set_error("external DOF %d not found in meqn2", d + 1); return FSR_ERR_ARG; }
        int j = -s - 1;
        bcol[i2] = j;     // ve(dofPosIn2(i2)) = finit(i2): column j of B multiplies finit(i2)
        extcol[j] = i2;
        ++i2;
      }
    if (i2 != ndof2) { set_error("found %d status-2 DOFs, expected ndof2=%d", i2, ndof2); return FSR_ERR_ARG; }
  }
  // CSR of row sources (disExpand, displacementModule.f90:1239-1257)
  std::vector<int> rowptr((size_t)ndof + 1, 0), src;
  std::vector<double> w;
  src.reserve(ndof);
  w.reserve(ndof);
  for (int d = 0; d < ndof; ++d) {
    int ieq = sam->meqn[d];
    int iceq = -ieq;
    if (ieq > 0 && ieq <= neq) {
      if (eqsrc[ieq] != NONE) { src.push_back(eqsrc[ieq]); w.push_back(1.0); }
    } else if (iceq > 0 && iceq <= sam->nceq) {
      for (int ip = sam->mpmceq[iceq - 1] + 1; ip <= sam->mpmceq[iceq] - 1; ++ip) {
        int m = sam->mmceq[ip - 1];
        if (m > 0 && m <= ndof) {
          int jeq = sam->meqn[m - 1];
          if (jeq > 0 && jeq <= neq && eqsrc[jeq] != NONE) {
            src.push_back(eqsrc[jeq]);
            w.push_back(sam->ttcc[ip - 1]);
          }
        }
      }
    }
    rowptr[d + 1] = (int)src.size();
  }
  // external sources per row, for the gage operator (ElDispFromSupElDisp puts the unit of external
  // DOF i at meqn2(i), displacementModule.f90:1170)
  p->ext_rowptr.assign((size_t)ndof + 1, 0);
  p->ext_j.clear(); p->ext_w.clear();
  for (int d = 0; d < ndof; ++d) {
    for (int ip = rowptr[d]; ip < rowptr[d + 1]; ++ip)
      if (src[ip] < 0) { p->ext_j.push_back(-src[ip] - 1); p->ext_w.push_back(w[ip]); }
    p->ext_rowptr[d + 1] = (int)p->ext_j.size();
  }
  p->extcol.assign(extcol.begin(), extcol.end());
  if (src.empty()) { src.push_back(0); w.push_back(0.0); }

  int *d_rowptr = nullptr, *d_src = nullptr, *d_bcol = nullptr, *d_extcol = nullptr;
  double *d_w = nullptr, *d_B = nullptr, *d_E = nullptr;
  cudaStream_t s = p->stream;
  FSR_CUDA(cudaMalloc(&d_rowptr, sizeof(int) * rowptr.size()));
  FSR_CUDA(cudaMalloc(&d_src, sizeof(int) * src.size()));
  FSR_CUDA(cudaMalloc(&d_w, sizeof(double) * w.size()));
  FSR_CUDA(cudaMalloc(&d_bcol, sizeof(int) * bcol.size()));
  FSR_CUDA(cudaMalloc(&d_extcol, sizeof(int) * extcol.size()));
  FSR_CUDA(cudaMemcpyAsync(d_rowptr, rowptr.data(), sizeof(int) * rowptr.size(), cudaMemcpyHostToDevice, s));
  FSR_CUDA(cudaMemcpyAsync(d_src, src.data(), sizeof(int) * src.size(), cudaMemcpyHostToDevice, s));
  FSR_CUDA(cudaMemcpyAsync(d_w, w.data(), sizeof(double) * w.size(), cudaMemcpyHostToDevice, s));
  FSR_CUDA(cudaMemcpyAsync(d_bcol, bcol.data(), sizeof(int) * bcol.size(), cudaMemcpyHostToDevice, s));
  FSR_CUDA(cudaMemcpyAsync(d_extcol, extcol.data(), sizeof(int) * extcol.size(), cudaMemcpyHostToDevice, s));
  size_t nB = (size_t)ldB * ndof2, nE = (size_t)ldE * ngen;
  if (ndof1 > 0 && ndof2 > 0) {
    FSR_CUDA(cudaMalloc(&d_B, sizeof(double) * nB));
    FSR_CUDA(cudaMemcpyAsync(d_B, B, sizeof(double) * nB, cudaMemcpyHostToDevice, s));
  }
  if (ndof1 > 0 && ngen > 0) {
    FSR_CUDA(cudaMalloc(&d_E, sizeof(double) * nE));
    FSR_CUDA(cudaMemcpyAsync(d_E, E, sizeof(double) * nE, cudaMemcpyHostToDevice, s));
  }
  FSR_CUDA(cudaMemsetAsync(p->R, 0, sizeof(double) * (size_t)p->nrows_pad * p->ldk, s));
  size_t total = (size_t)ndof * p->ndim;
  if (total > 0) {
    unsigned blocks = (unsigned)((total + 255) / 256);
    build_R_kernel<<<blocks, 256, 0, s>>>(p->R, p->ldk, ndof, p->ndim, ndof2, d_B, (size_t)ldB, d_E,
                                          (size_t)ldE, d_rowptr, d_src, d_w, d_bcol, d_extcol);
    FSR_LAUNCH_CHECK();
  }
  FSR_CUDA(cudaStreamSynchronize(s));
  cudaFree(d_rowptr); cudaFree(d_src); cudaFree(d_w); cudaFree(d_bcol); cudaFree(d_extcol);
  cudaFree(d_B); cudaFree(d_E);
  p->have_R = true;
  return FSR_OK;
}

// ------------------------------------------------------------------------------------------
// Q packing: caller's Q (ndim x nsteps, column-major, ldq) -> Qt[nsteps_pad][ldk], zero padded
// ------------------------------------------------------------------------------------------
__global__ void pack_q_kernel(double* __restrict__ Qt, int ldk, const double* __restrict__ Q,
                              size_t ldq, int ndim, int nsteps, int nsteps_pad)
{
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t total = (size_t)nsteps_pad * ldk;
  if (idx >= total) return;
  int k = (int)(idx % ldk);
  int s = (int)(idx / ldk);
  Qt[idx] = (k < ndim && s < nsteps) ? Q[(size_t)s * ldq + k] : 0.0;
}

int launch_pack_q_raw(double* Qt, int ldk, const double* Q_dev, int ldq, int ndim, int nsteps, int nsteps_pad,
                      cudaStream_t s)
{
  size_t total = (size_t)nsteps_pad * ldk;
  unsigned blocks = (unsigned)((total + 255) / 256);
  pack_q_kernel<<<blocks, 256, 0, s>>>(Qt, ldk, Q_dev, (size_t)ldq, ndim, nsteps, nsteps_pad);
  FSR_LAUNCH_CHECK();
  return FSR_OK;
}

int launch_pack_q(fsr_part* p, const double* Q_dev, int ldq, int nsteps, int nsteps_pad,
                  cudaStream_t s)
{
  return launch_pack_q_raw(p->Qt, p->ldk, Q_dev, ldq, p->ndim, nsteps, nsteps_pad, s);
}

// ------------------------------------------------------------------------------------------
// The DMMA GEMM
// ------------------------------------------------------------------------------------------
constexpr int K1_BM = 128;      // rows (nodal DOFs) per CTA
constexpr int K1_BN = 64;       // steps per shared-memory chunk
constexpr int K1_THREADS = 256; // 8 warps: 4 along M x 2 along N, warp tile 32 x 32

__device__ __forceinline__ uint32_t smem_u32(const void* p)
{
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  }
}
// 1-D bulk TMA: global -> shared, completion counted in bytes on an mbarrier.
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar)
{
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

__global__ void __launch_bounds__(K1_THREADS, 1)
k1_expand_kernel(const double* __restrict__ R, const double* __restrict__ Qt, double* __restrict__ U,
                 int ldk, int nsteps_pad, size_t ldu, const int* __restrict__ tiles)
{
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* sR = reinterpret_cast<double*>(smem_raw);   // [K1_BM][ldk]
  double* sQ = sR + (size_t)K1_BM * ldk;              // 2 x [K1_BN][ldk]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sQ + (size_t)2 * K1_BN * ldk); // [0]=R, [1..2]=Q

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t4 = lane & 3;
  const int wm = warp & 3, wn = warp >> 2;
  const size_t row0 = (size_t)(tiles ? __ldg(tiles + blockIdx.x) : (int)blockIdx.x) * K1_BM;
  const int nchunks = nsteps_pad / K1_BN;
  const uint32_t bytesR = (uint32_t)(K1_BM * ldk * sizeof(double));
  const uint32_t bytesQ = (uint32_t)(K1_BN * ldk * sizeof(double));

  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_init(&bars[2], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (tid == 0) {
    // the row tile of R is one contiguous block of memory: issue it as four bulk copies
    mbar_expect_tx(&bars[0], bytesR);
    const uint32_t q = bytesR / 4;  // K1_BM*ldk*8/4 = 32*ldk*8, a multiple of 16
    const unsigned char* srcR = reinterpret_cast<const unsigned char*>(R + row0 * ldk);
    for (int i = 0; i < 4; ++i)
      tma_load_1d(reinterpret_cast<unsigned char*>(sR) + (size_t)i * q, srcR + (size_t)i * q, q, &bars[0]);
    mbar_expect_tx(&bars[1], bytesQ);
    tma_load_1d(sQ, Qt, bytesQ, &bars[1]);
  }

  for (int c = 0; c < nchunks; ++c) {
    const int buf = c & 1;
    if (tid == 0 && c + 1 < nchunks) {
      // buffer buf^1 was released by the __syncthreads that closed iteration c-1
      mbar_expect_tx(&bars[1 + (buf ^ 1)], bytesQ);
      tma_load_1d(sQ + (size_t)(buf ^ 1) * K1_BN * ldk, Qt + (size_t)(c + 1) * K1_BN * ldk, bytesQ,
                  &bars[1 + (buf ^ 1)]);
    }
    if (c == 0) mbar_wait(&bars[0], 0);
    mbar_wait(&bars[1 + buf], (uint32_t)((c >> 1) & 1));

    const double* a_base = sR + (size_t)(wm * 32 + g) * ldk + t4;
    const double* b_base = sQ + (size_t)buf * K1_BN * ldk + (size_t)(wn * 32 + g) * ldk + t4;

    double acc[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    const int ktiles = ldk >> 2;
#pragma unroll 2
    for (int kt = 0; kt < ktiles; ++kt) {
      double a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = a_base[(size_t)i * 8 * ldk + kt * 4];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = b_base[(size_t)j * 8 * ldk + kt * 4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }

    // epilogue: each lane owns 2 consecutive steps of one DOF row per 8x8 tile -> 16-byte stores,
    // four lanes complete a 64-byte segment
    double* u_base = U + (row0 + wm * 32 + g) * ldu + (size_t)c * K1_BN + wn * 32 + 2 * t4;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        double2 v = make_double2(acc[i][j][0], acc[i][j][1]);
        *reinterpret_cast<double2*>(u_base + (size_t)i * 8 * ldu + j * 8) = v;
      }
    __syncthreads();
  }
}

// The same kernel without a CTA-wide barrier per step chunk: a warp that has finished chunk c arrives on the chunk buffer's
// "empty" mbarrier and goes straight on to chunk c + 1 (already in shared memory); thread 0 alone waits for the eight arrivals
// before it refills the buffer with chunk c + 2.  The accumulators of chunk c are kept in a second register set and stored
// between the k-tiles of chunk c + 1, so the tensor pipe does not drain while the results go out (190 registers; one CTA per
// SM has 255 to spend).
__device__ __forceinline__ void mbar_arrive(uint64_t* bar)
{
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__global__ void __launch_bounds__(K1_THREADS, 1)
k1_expand_pipe_kernel(const double* __restrict__ R, const double* __restrict__ Qt, double* __restrict__ U,
                      int ldk, int nsteps_pad, size_t ldu, const int* __restrict__ tiles)
{
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* sR = reinterpret_cast<double*>(smem_raw);   // [K1_BM][ldk]
  double* sQ = sR + (size_t)K1_BM * ldk;              // 2 x [K1_BN][ldk]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sQ + (size_t)2 * K1_BN * ldk); // [0]=R, [1..2]=Q full, [3..4]=Q empty

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t4 = lane & 3;
  const int wm = warp & 3, wn = warp >> 2;
  const size_t row0 = (size_t)(tiles ? __ldg(tiles + blockIdx.x) : (int)blockIdx.x) * K1_BM;
  const int nchunks = nsteps_pad / K1_BN;
  const uint32_t bytesR = (uint32_t)(K1_BM * ldk * sizeof(double));
  const uint32_t bytesQ = (uint32_t)(K1_BN * ldk * sizeof(double));

  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_init(&bars[2], 1);
    mbar_init(&bars[3], K1_THREADS / 32);
    mbar_init(&bars[4], K1_THREADS / 32);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (tid == 0) {
    mbar_expect_tx(&bars[0], bytesR);
    const uint32_t q = bytesR / 4;
    const unsigned char* srcR = reinterpret_cast<const unsigned char*>(R + row0 * ldk);
    for (int i = 0; i < 4; ++i)
      tma_load_1d(reinterpret_cast<unsigned char*>(sR) + (size_t)i * q, srcR + (size_t)i * q, q, &bars[0]);
    for (int c = 0; c < 2 && c < nchunks; ++c) {
      mbar_expect_tx(&bars[1 + c], bytesQ);
      tma_load_1d(sQ + (size_t)c * K1_BN * ldk, Qt + (size_t)c * K1_BN * ldk, bytesQ, &bars[1 + c]);
    }
  }
  mbar_wait(&bars[0], 0);

  const double* a_base = sR + (size_t)(wm * 32 + g) * ldk + t4;
  double* u_row = U + (row0 + wm * 32 + g) * ldu + wn * 32 + 2 * t4;
  const int ktiles = ldk >> 2, kq = ktiles >> 2;
  double acc[4][4][2], prev[4][4][2];

  for (int c = 0; c < nchunks; ++c) {
    const int buf = c & 1;
    mbar_wait(&bars[1 + buf], (uint32_t)((c >> 1) & 1));
    const double* b_base = sQ + (size_t)buf * K1_BN * ldk + (size_t)(wn * 32 + g) * ldk + t4;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    double* u_prev = u_row + (size_t)(c - 1) * K1_BN;
#pragma unroll
    for (int part = 0; part < 4; ++part) {
      const int k_end = part == 3 ? ktiles : (part + 1) * kq;
#pragma unroll 2
      for (int kt = part * kq; kt < k_end; ++kt) {
        double a[4], b[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = a_base[(size_t)i * 8 * ldk + kt * 4];
#pragma unroll
        for (int j = 0; j < 4; ++j) b[j] = b_base[(size_t)j * 8 * ldk + kt * 4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
      }
      if (c > 0) {   // a quarter of the previous chunk's results goes out between the k-tiles of this one
#pragma unroll
        for (int j = 0; j < 4; ++j)
          *reinterpret_cast<double2*>(u_prev + (size_t)part * 8 * ldu + j * 8) = make_double2(prev[part][j][0], prev[part][j][1]);
      }
    }
    // this warp is done with the chunk buffer
    __syncwarp();
    if (lane == 0) mbar_arrive(&bars[3 + buf]);
    if (tid == 0 && c + 2 < nchunks) {
      mbar_wait(&bars[3 + buf], (uint32_t)((c >> 1) & 1));   // all eight warps have left it
      mbar_expect_tx(&bars[1 + buf], bytesQ);
      tma_load_1d(sQ + (size_t)buf * K1_BN * ldk, Qt + (size_t)(c + 2) * K1_BN * ldk, bytesQ, &bars[1 + buf]);
    }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) { prev[i][j][0] = acc[i][j][0]; prev[i][j][1] = acc[i][j][1]; }
  }
  if (nchunks > 0) {
    double* u_prev = u_row + (size_t)(nchunks - 1) * K1_BN;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j)
        *reinterpret_cast<double2*>(u_prev + (size_t)i * 8 * ldu + j * 8) = make_double2(prev[i][j][0], prev[i][j][1]);
  }
}

// The same GEMM for reduced dimensions that do not fit the resident-K kernel (ldk > 108: superelements with many triads or
// component modes): the K dimension is cut into slabs of 52 columns; a stage = the slab of the R row tile and of one 64-step
// Q chunk, copied row by row with bulk TMA into one of two shared-memory buffers while the other is multiplied; the
// accumulators of a chunk stay in registers across its slabs.  R is re-read once per step chunk (from L2 for the most part).
constexpr int K1S_LS = 52;   // slab width, == 4 (mod 8) like ldk

__global__ void __launch_bounds__(K1_THREADS, 1)
k1_expand_slab_kernel(const double* __restrict__ R, const double* __restrict__ Qt, double* __restrict__ U, int ldk, int nsteps_pad,
                      size_t ldu, const int* __restrict__ tiles)
{
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* sbuf = reinterpret_cast<double*>(smem_raw);                       // 2 x [K1_BM + K1_BN][K1S_LS]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sbuf + (size_t)2 * (K1_BM + K1_BN) * K1S_LS);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t4 = lane & 3;
  const int wm = warp & 3, wn = warp >> 2;
  const size_t row0 = (size_t)(tiles ? __ldg(tiles + blockIdx.x) : (int)blockIdx.x) * K1_BM;
  const int nchunks = nsteps_pad / K1_BN;
  const int nslab = (ldk + K1S_LS - 1) / K1S_LS;
  const int nstage = nchunks * nslab;

  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  auto issue = [&](int st) {   // warp 0 only
    const int buf = st & 1, c = st / nslab, sl = st - c * nslab;
    const int k0 = sl * K1S_LS, w = min(K1S_LS, ldk - k0);
    double* dst = sbuf + (size_t)buf * (K1_BM + K1_BN) * K1S_LS;
    if (lane == 0) mbar_expect_tx(&bars[buf], (uint32_t)((K1_BM + K1_BN) * w * sizeof(double)));
    __syncwarp();
    for (int r = lane; r < K1_BM + K1_BN; r += 32) {
      const double* src = r < K1_BM ? R + (row0 + r) * ldk + k0 : Qt + ((size_t)c * K1_BN + (r - K1_BM)) * ldk + k0;
      tma_load_1d(dst + (size_t)r * K1S_LS, src, (uint32_t)(w * sizeof(double)), &bars[buf]);
    }
  };
  if (warp == 0) {
    issue(0);
    if (nstage > 1) issue(1);
  }

  double acc[4][4][2];
  for (int st = 0; st < nstage; ++st) {
    const int buf = st & 1, c = st / nslab, sl = st - c * nslab;
    const int w = min(K1S_LS, ldk - sl * K1S_LS);
    mbar_wait(&bars[buf], (uint32_t)((st >> 1) & 1));
    if (sl == 0) {
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    }
    const double* sR = sbuf + (size_t)buf * (K1_BM + K1_BN) * K1S_LS;
    const double* a_base = sR + (size_t)(wm * 32 + g) * K1S_LS + t4;
    const double* b_base = sR + (size_t)(K1_BM + wn * 32 + g) * K1S_LS + t4;
    const int ktiles = w >> 2;
#pragma unroll 2
    for (int kt = 0; kt < ktiles; ++kt) {
      double a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = a_base[(size_t)i * 8 * K1S_LS + kt * 4];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = b_base[(size_t)j * 8 * K1S_LS + kt * 4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
    if (sl == nslab - 1) {
      double* u_base = U + (row0 + wm * 32 + g) * ldu + (size_t)c * K1_BN + wn * 32 + 2 * t4;
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
          *reinterpret_cast<double2*>(u_base + (size_t)i * 8 * ldu + j * 8) = make_double2(acc[i][j][0], acc[i][j][1]);
    }
    __syncthreads();   // the buffer is free again
    if (warp == 0 && st + 2 < nstage) issue(st + 2);
  }
}

static size_t k1_slab_smem_bytes() { return (size_t)2 * (K1_BM + K1_BN) * K1S_LS * sizeof(double) + 2 * sizeof(uint64_t) + 64; }

static size_t k1_smem_bytes(int ldk)
{
  return ((size_t)K1_BM * ldk + (size_t)2 * K1_BN * ldk) * sizeof(double) + 3 * sizeof(uint64_t) + 64;
}

int launch_k1_raw(const double* R, const double* Qt, double* U, int ldk, int nrows_pad, int nsteps_pad, size_t ldu,
                  cudaStream_t s, const int* tiles, int ntiles)
{
  size_t smem = k1_smem_bytes(ldk);
  if (nsteps_pad % K1_BN != 0 || nrows_pad % K1_BM != 0) {
    set_error("internal: K1 tile mismatch (%d rows, %d steps)", nrows_pad, nsteps_pad);
    return FSR_ERR_ARG;
  }
  unsigned blocks = tiles ? (unsigned)ntiles : (unsigned)(nrows_pad / K1_BM);   // tiles: only these row tiles
  if (blocks == 0) return FSR_OK;
  // FSR_K1_SLAB=1 forces the K-slab kernel (tests); it is the only one for ldk > 108
  static const bool force_slab = getenv("FSR_K1_SLAB") && atoi(getenv("FSR_K1_SLAB")) != 0;
  if (smem > 227 * 1024 || force_slab) {
    if (int rc = smem_opt_in((const void*)k1_expand_slab_kernel, k1_slab_smem_bytes())) return rc;
    k1_expand_slab_kernel<<<blocks, K1_THREADS, k1_slab_smem_bytes(), s>>>(R, Qt, U, ldk, nsteps_pad, ldu, tiles);
    FSR_LAUNCH_CHECK();
    return FSR_OK;
  }
  // FSR_K1_PIPE=0: the version with a CTA barrier per step chunk (kept for comparison)
  static const bool pipe = !(getenv("FSR_K1_PIPE") && atoi(getenv("FSR_K1_PIPE")) == 0);
  if (pipe) {
    if (int rc = smem_opt_in((const void*)k1_expand_pipe_kernel, 227 * 1024)) return rc;
    k1_expand_pipe_kernel<<<blocks, K1_THREADS, smem + 2 * sizeof(uint64_t), s>>>(R, Qt, U, ldk, nsteps_pad, ldu, tiles);
    FSR_LAUNCH_CHECK();
    return FSR_OK;
  }
  if (int rc = smem_opt_in((const void*)k1_expand_kernel, 227 * 1024)) return rc;
  k1_expand_kernel<<<blocks, K1_THREADS, smem, s>>>(R, Qt, U, ldk, nsteps_pad, ldu, tiles);
  FSR_LAUNCH_CHECK();
  return FSR_OK;
}

int launch_k1(fsr_part* p, int nsteps_pad, cudaStream_t s)
{
  return launch_k1_raw(p->R, p->Qt, p->U, p->ldk, p->nrows_pad, nsteps_pad, (size_t)p->step_tile, s);
}

// ------------------------------------------------------------------------------------------
// In-plane rows of flat shell regions (fsr_part::planar, set up by k2_shell.cu)
// ------------------------------------------------------------------------------------------
// out[r][c] = sum_j w[r][j] * in[src[r][j]][c]: the rows of R (once) or of U (displacements given) in the axes of the plane
__global__ void combine_rows_kernel(double* __restrict__ out, size_t ld_out, const double* __restrict__ in, size_t ld_in,
                                    const int* __restrict__ src, const double* __restrict__ w, int nrows, int ncols)
{
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)nrows * ncols) return;
  const int r = (int)(idx / ncols), c = (int)(idx % ncols);
  double acc = 0.0;
#pragma unroll
  for (int j = 0; j < 3; ++j) acc = fma(w[(size_t)r * 3 + j], in[(size_t)src[(size_t)r * 3 + j] * ld_in + c], acc);
  out[(size_t)r * ld_out + c] = acc;
}

int build_planar_rows(fsr_part* p)
{
  if (!p->planar) return FSR_OK;
  cudaStream_t s = p->stream;
  if (!p->Rp) FSR_CUDA(cudaMalloc(&p->Rp, sizeof(double) * (size_t)p->np_rows_pad * p->ldk));
  FSR_CUDA(cudaMemsetAsync(p->Rp, 0, sizeof(double) * (size_t)p->np_rows_pad * p->ldk, s));
  const size_t total = (size_t)p->np_rows * p->ldk;
  combine_rows_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(p->Rp, (size_t)p->ldk, p->R, (size_t)p->ldk, p->prow_src,
                                                                      p->prow_w, p->np_rows, p->ldk);
  FSR_LAUNCH_CHECK();
  FSR_CUDA(cudaStreamSynchronize(s));
  return FSR_OK;
}

int planar_rows_from_u(fsr_part* p, int nsteps_pad, cudaStream_t s)
{
  if (!p->planar) return FSR_OK;
  const size_t total = (size_t)p->np_rows * nsteps_pad;
  combine_rows_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(p->Up, (size_t)p->step_tile, p->U, (size_t)p->step_tile,
                                                                      p->prow_src, p->prow_w, p->np_rows, nsteps_pad);
  FSR_LAUNCH_CHECK();
  return FSR_OK;
}

int launch_k1_vm(fsr_part* p, int nsteps_pad, cudaStream_t s, bool full_u)
{
  if (!p->planar) return launch_k1(p, nsteps_pad, s);
  int rc = launch_k1_raw(p->Rp, p->Qt, p->Up, p->ldk, p->np_rows_pad, nsteps_pad, (size_t)p->step_tile, s);
  if (rc) return rc;
  if (full_u) return launch_k1(p, nsteps_pad, s);
  if (p->n_k1_tiles == 0) return FSR_OK;
  return launch_k1_raw(p->R, p->Qt, p->U, p->ldk, p->nrows_pad, nsteps_pad, (size_t)p->step_tile, s, p->k1_tiles, p->n_k1_tiles);
}

}  // namespace fsr
