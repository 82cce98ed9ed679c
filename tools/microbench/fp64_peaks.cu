// FP64 throughput probes for B200 (sm_100a): DMMA.8x8x4 issue rate, DFMA rate,
// shared-memory fed DMMA. Prints TFLOP/s; used once to pick the K1/K2 design and
// to give the FP64 roofline denominator. Not part of the product path.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__); exit(1);} }while(0)

template<int NACC>
__global__ void __launch_bounds__(256) dmma_loop(double* out, int iters, double seed){
  double a = seed + threadIdx.x*1e-9, b = seed*0.5 + threadIdx.x*1e-9;
  double c[NACC][2];
  #pragma unroll
  for(int i=0;i<NACC;i++){c[i][0]=0;c[i][1]=0;}
  for(int it=0; it<iters; ++it){
    #pragma unroll
    for(int i=0;i<NACC;i++)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   :"+d"(c[i][0]),"+d"(c[i][1]):"d"(a),"d"(b));
  }
  double s=0;
  #pragma unroll
  for(int i=0;i<NACC;i++) s+=c[i][0]+c[i][1];
  if(s==123.456) out[0]=s;
}

template<int NACC>
__global__ void __launch_bounds__(256) dfma_loop(double* out, int iters, double seed){
  double a = seed + threadIdx.x*1e-9, b = seed*0.5;
  double c[NACC];
  #pragma unroll
  for(int i=0;i<NACC;i++) c[i]=i;
  for(int it=0; it<iters; ++it){
    #pragma unroll
    for(int i=0;i<NACC;i++) c[i]=fma(c[i],a,b);
  }
  double s=0;
  #pragma unroll
  for(int i=0;i<NACC;i++) s+=c[i];
  if(s==123.456) out[0]=s;
}

int main(){
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p,0));
  printf("device %s sms %d clock %d kHz\n", p.name, p.multiProcessorCount, p.clockRate);
  double* out; CK(cudaMalloc(&out,8));
  cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  int nsm=p.multiProcessorCount;
  for(int wpb=1; wpb<=8; wpb*=2){
    for(int bps=1;bps<=4;bps*=2){
      int iters=20000; float ms;
      dmma_loop<8><<<nsm*bps, 32*wpb>>>(out, 100, 1.0);
      CK(cudaDeviceSynchronize());
      cudaEventRecord(e0);
      dmma_loop<8><<<nsm*bps, 32*wpb>>>(out, iters, 1.0);
      cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms,e0,e1);
      double fl = 2.0*8*8*4*8.0*iters*wpb*bps*nsm;
      printf("DMMA.8x8x4 warps/blk %d blk/SM %d : %.2f TFLOP/s (%.3f ms)\n", wpb,bps, fl/ms*1e-9, ms);
    }
  }
  // dependent-chain probe: one warp per SM, NACC independent accumulators -> DMMA latency
  {
    int iters=20000; float ms;
    #define LAT(N) dmma_loop<N><<<nsm,32>>>(out,100,1.0); CK(cudaDeviceSynchronize()); cudaEventRecord(e0); \
      dmma_loop<N><<<nsm,32>>>(out,iters,1.0); cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms,e0,e1); \
      printf("DMMA chain probe NACC=%d: %.1f cycles per DMMA per warp (clock %d kHz)\n", N, ms*1e-3*p.clockRate*1e3/(double)(iters*N), p.clockRate);
    LAT(1) LAT(2) LAT(3) LAT(4) LAT(6) LAT(8)
  }
  for(int wpb=4; wpb<=8; wpb*=2){
    for(int bps=1;bps<=4;bps*=2){
      int iters=20000; float ms;
      dfma_loop<16><<<nsm*bps, 32*wpb>>>(out, 100, 1.0);
      CK(cudaDeviceSynchronize());
      cudaEventRecord(e0);
      dfma_loop<16><<<nsm*bps, 32*wpb>>>(out, iters, 1.0);
      cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms,e0,e1);
      double fl = 2.0*16*32.0*iters*wpb*bps*nsm;
      printf("DFMA warps/blk %d blk/SM %d : %.2f TFLOP/s (%.3f ms)\n", wpb,bps, fl/ms*1e-9, ms);
    }
  }
  return 0;
}
