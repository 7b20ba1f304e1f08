// FP64 pipe microbenchmark for sm_100a: DFMA vs DMUL+DADD vs DMMA.8x8x4 issue rates.
// Used to establish the FP64 roofline denominator next to cuBLAS DGEMM (see bench.py).
#include <cstdio>
#include <cuda_runtime.h>

#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1;}}while(0)

template<int ILP>
__global__ void k_dfma(double* out, double a, double b, int iters){
  double acc[ILP];
  #pragma unroll
  for(int i=0;i<ILP;i++) acc[i]=threadIdx.x+i;
  for(int it=0; it<iters; it++){
    #pragma unroll
    for(int i=0;i<ILP;i++) acc[i]=fma(acc[i], a, b);
  }
  double s=0;
  #pragma unroll
  for(int i=0;i<ILP;i++) s+=acc[i];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}

template<int ILP>
__global__ void k_dmuladd(double* out, double a, double b, int iters){
  double acc[ILP];
  #pragma unroll
  for(int i=0;i<ILP;i++) acc[i]=threadIdx.x+i;
  for(int it=0; it<iters; it++){
    #pragma unroll
    for(int i=0;i<ILP;i++) acc[i]=__dadd_rn(__dmul_rn(acc[i], a), b);
  }
  double s=0;
  #pragma unroll
  for(int i=0;i<ILP;i++) s+=acc[i];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}

template<int NACC>
__global__ void k_dmma(double* out, double a, double b, int iters){
  double c[NACC][2];
  #pragma unroll
  for(int i=0;i<NACC;i++){c[i][0]=0;c[i][1]=0;}
  double av=a+threadIdx.x, bv=b-threadIdx.x;
  for(int it=0; it<iters; it++){
    #pragma unroll
    for(int i=0;i<NACC;i++)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                   :"+d"(c[i][0]),"+d"(c[i][1]):"d"(av),"d"(bv));
  }
  double s=0;
  #pragma unroll
  for(int i=0;i<NACC;i++) s+=c[i][0]+c[i][1];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}

// mixed: DMMA with interleaved DFMA (do they share the pipe?)
template<int NACC, int NF>
__global__ void k_mixed(double* out, double a, double b, int iters){
  double c[NACC][2]; double f[NF];
  #pragma unroll
  for(int i=0;i<NACC;i++){c[i][0]=0;c[i][1]=0;}
  #pragma unroll
  for(int i=0;i<NF;i++) f[i]=threadIdx.x+i;
  double av=a+threadIdx.x, bv=b-threadIdx.x;
  for(int it=0; it<iters; it++){
    #pragma unroll
    for(int i=0;i<NACC;i++)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                   :"+d"(c[i][0]),"+d"(c[i][1]):"d"(av),"d"(bv));
    #pragma unroll
    for(int i=0;i<NF;i++) f[i]=fma(f[i], a, b);
  }
  double s=0;
  #pragma unroll
  for(int i=0;i<NACC;i++) s+=c[i][0]+c[i][1];
  #pragma unroll
  for(int i=0;i<NF;i++) s+=f[i];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}

template<typename F>
float timeit(F f, int reps){
  cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  float best=1e30f;
  for(int r=0;r<reps;r++){
    cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms,e0,e1); if(ms<best) best=ms;
  }
  return best;
}

int main(){
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p,0));
  int sms=p.multiProcessorCount;
  printf("device %s SMs %d clock %d kHz\n", p.name, sms, p.clockRate);
  double* out; CK(cudaMalloc(&out, sizeof(double)*sms*32*1024));
  const int iters=20000;
  for(int warps=4; warps<=32; warps*=2){
    for(int cps=1; cps<=2; cps++){
      int threads=warps*32, blocks=sms*cps;
      if(threads*cps>2048) continue;
      float ms;
      ms=timeit([&]{k_dfma<8><<<blocks,threads>>>(out,1.0000001,1e-9,iters);},5);
      double tf=2.0*8*iters*(double)threads*blocks/ms/1e9;
      printf("DFMA    warps/CTA %2d CTAs/SM %d : %.3f ms  %.2f TFLOP/s\n",warps,cps,ms,tf);
      ms=timeit([&]{k_dmuladd<8><<<blocks,threads>>>(out,1.0000001,1e-9,iters);},5);
      tf=2.0*8*iters*(double)threads*blocks/ms/1e9;
      printf("DMUL+DADD warps/CTA %2d CTAs/SM %d : %.3f ms  %.2f Tops/s (mul+add counted 2)\n",warps,cps,ms,tf);
      ms=timeit([&]{k_dmma<8><<<blocks,threads>>>(out,1.0000001,1e-9,iters);},5);
      tf=2.0*256*8*iters*(double)warps*blocks/ms/1e9;
      printf("DMMA884 x8acc warps/CTA %2d CTAs/SM %d : %.3f ms  %.2f TFLOP/s\n",warps,cps,ms,tf);
      ms=timeit([&]{k_dmma<16><<<blocks,threads>>>(out,1.0000001,1e-9,iters/2);},5);
      tf=2.0*256*16*(iters/2)*(double)warps*blocks/ms/1e9;
      printf("DMMA884 x16acc warps/CTA %2d CTAs/SM %d : %.3f ms  %.2f TFLOP/s\n",warps,cps,ms,tf);
      ms=timeit([&]{k_dmma<2><<<blocks,threads>>>(out,1.0000001,1e-9,iters);},5);
      tf=2.0*256*2*iters*(double)warps*blocks/ms/1e9;
      printf("DMMA884 x2acc warps/CTA %2d CTAs/SM %d : %.3f ms  %.2f TFLOP/s\n",warps,cps,ms,tf);
      ms=timeit([&]{k_mixed<8,8><<<blocks,threads>>>(out,1.0000001,1e-9,iters);},5);
      tf=(2.0*256*8*warps + 2.0*8*threads)*iters*(double)blocks/ms/1e9;
      printf("MIXED 8 DMMA + 8 DFMA warps/CTA %2d CTAs/SM %d : %.3f ms  %.2f TFLOP/s (sum)\n",warps,cps,ms,tf);
    }
  }
  return 0;
}
