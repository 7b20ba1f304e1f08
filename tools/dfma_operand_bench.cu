// Does the FP64 pipe of sm_100a sustain its peak when every DFMA reads three distinct,
// non-reused 64-bit register operands?  (k_triples: most DFMAs do.)  Two loops with the same
// instruction count: (A) c[i] = fma(a[i], b[i], c[i]) - three distinct register pairs per
// instruction, nothing shared between neighbours; (B) c[i] = fma(a0, b0, c[i]) - two operands
// shared by all instructions (operand-reuse cache hits).
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256) k(double *out, int iters, double seed)
{
    double a[12], b[12], c[12];
#pragma unroll
    for (int i = 0; i < 12; i++) { a[i] = seed + 1e-9 * (threadIdx.x + i); b[i] = 1.0 - 1e-9 * i; c[i] = 1e-3 * i; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 4; r++)
#pragma unroll
            for (int i = 0; i < 12; i++) {
                if (MODE == 0) c[i] = fma(a[i], b[i], c[i]);
                else if (MODE == 1) c[i] = fma(a[0], b[0], c[i]);
                else c[i] = fma(a[i], b[(i + 5) % 12], c[i]);      // distinct, and b shuffled against a
            }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 12; i++) s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char *name, int warps_per_sm)
{
    double *out;
    cudaMalloc(&out, sizeof(double) * 148 * 1024);
    const int iters = 20000, threads = 32 * warps_per_sm;
    k<MODE><<<148, threads>>>(out, 10, 0.5);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<MODE><<<148, threads>>>(out, iters, 0.5);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double flops = 2.0 * 48 * (double)iters * threads * 148;
    printf("%-44s %2d warps/SM: %.3f ms  %.2f TFLOP/s\n", name, warps_per_sm, ms, flops / ms / 1e9);
    cudaFree(out);
}

int main()
{
    for (int w : {4, 8, 12, 16}) {
        run<0>("3 distinct operands c[i]=fma(a[i],b[i],c[i])", w);
        run<2>("3 distinct operands, b shuffled", w);
        run<1>("2 shared operands  c[i]=fma(a0,b0,c[i])", w);
    }
    return 0;
}
