// Isolated DMMA k-loop of k_fast_pairs: smem-resident tiles, NW consumer warps, no producers,
// trivial epilogue.  Measures what the k-loop alone can sustain (fraction of the FP64 pipe).
#include <cstdio>
#include <cuda_runtime.h>
#define S1 132
#define S2 36
#define MP 108

template <int VARIANT>
__global__ void __launch_bounds__(256, 1) kloop(double *out, int tiles)
{
    extern __shared__ double sm[];
    double *D1s = sm, *D2s = sm + MP * S1;
    for (int i = threadIdx.x; i < MP * S1 + 3 * MP * S2; i += blockDim.x) sm[i] = 1e-3 * (i % 97);
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t4 = lane & 3;
    const int wrow = warp * 16;
    double tot = 0.0;
    for (int jt = 0; jt < tiles; jt++) {
        const int st = jt % 3;
        double acc[2][4][2];
#pragma unroll
        for (int mt = 0; mt < 2; mt++)
#pragma unroll
            for (int nt = 0; nt < 4; nt++) { acc[mt][nt][0] = 0.0; acc[mt][nt][1] = 0.0; }
        const double *A_ = D1s + t4 * S1 + wrow + g;
        const double *B_ = D2s + st * MP * S2 + t4 * S2 + g;
        if (VARIANT == 0) {
#pragma unroll 3
            for (int ks = 0; ks < MP / 4; ks++) {
                double af[2], bf[4];
#pragma unroll
                for (int mt = 0; mt < 2; mt++) af[mt] = A_[ks * 4 * S1 + 8 * mt];
#pragma unroll
                for (int nt = 0; nt < 4; nt++) bf[nt] = B_[ks * 4 * S2 + 8 * nt];
#pragma unroll
                for (int mt = 0; mt < 2; mt++)
#pragma unroll
                    for (int nt = 0; nt < 4; nt++)
                        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                                     : "+d"(acc[mt][nt][0]), "+d"(acc[mt][nt][1]) : "d"(af[mt]), "d"(bf[nt]));
            }
        } else if (VARIANT == 1) {
            // register double-buffered fragments: loads of step k+1 issued before the DMMAs of step k
            double af[2][2], bf[2][4];
#pragma unroll
            for (int mt = 0; mt < 2; mt++) af[0][mt] = A_[8 * mt];
#pragma unroll
            for (int nt = 0; nt < 4; nt++) bf[0][nt] = B_[8 * nt];
#pragma unroll
            for (int ks = 0; ks < MP / 4; ks++) {
                const int cur = ks & 1, nxt = cur ^ 1;
                if (ks + 1 < MP / 4) {
#pragma unroll
                    for (int mt = 0; mt < 2; mt++) af[nxt][mt] = A_[(ks + 1) * 4 * S1 + 8 * mt];
#pragma unroll
                    for (int nt = 0; nt < 4; nt++) bf[nxt][nt] = B_[(ks + 1) * 4 * S2 + 8 * nt];
                }
#pragma unroll
                for (int mt = 0; mt < 2; mt++)
#pragma unroll
                    for (int nt = 0; nt < 4; nt++)
                        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                                     : "+d"(acc[mt][nt][0]), "+d"(acc[mt][nt][1]) : "d"(af[cur][mt]), "d"(bf[cur][nt]));
            }
        } else {
            // no smem loads at all (register operands): upper bound of the DMMA stream
            double af[2] = {1.0 + lane, 2.0}, bf[4] = {1.0, 2.0, 3.0, 4.0 + lane};
#pragma unroll 3
            for (int ks = 0; ks < MP / 4; ks++) {
#pragma unroll
                for (int mt = 0; mt < 2; mt++)
#pragma unroll
                    for (int nt = 0; nt < 4; nt++)
                        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                                     : "+d"(acc[mt][nt][0]), "+d"(acc[mt][nt][1]) : "d"(af[mt]), "d"(bf[nt]));
            }
        }
#pragma unroll
        for (int mt = 0; mt < 2; mt++)
#pragma unroll
            for (int nt = 0; nt < 4; nt++) tot += acc[mt][nt][0] + acc[mt][nt][1];
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = tot;
}

template <int V>
void run(const char *name, int threads)
{
    double *out;
    cudaMalloc(&out, sizeof(double) * 148 * 512);
    size_t smem = sizeof(double) * (MP * S1 + 3 * MP * S2);
    cudaFuncSetAttribute(kloop<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int tiles = 2000;
    kloop<V><<<148, threads, smem>>>(out, 10);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    kloop<V><<<148, threads, smem>>>(out, tiles);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    double flops = 2.0 * 256 * 8 * 27 * (threads / 32) * (double)tiles * 148;
    printf("%-40s threads %d: %.3f ms  %.2f TFLOP/s  err=%s\n", name, threads, ms, flops / ms / 1e9,
           cudaGetErrorString(cudaGetLastError()));
    cudaFree(out);
}

int main()
{
    run<0>("k-loop as in k_fast_pairs", 256);
    run<1>("register double-buffered fragments", 256);
    run<2>("register operands only", 256);
    run<0>("k-loop as in k_fast_pairs", 128);
    run<1>("register double-buffered fragments", 128);
    run<2>("register operands only", 128);
    return 0;
}
