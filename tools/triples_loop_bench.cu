// Inner loop of k_triples in isolation (shared-memory resident operands, no global traffic, the
// vote never fires): which thread tile / CTA shape / vote layout lets the FP64 pipe of sm_100a
// run closest to its peak?  Prints executed FP64 operations per second against 37.1 TFLOP/s
// (18.55 T FP64 instructions-lanes per second).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/triples_loop_bench tools/triples_loop_bench.cu
#include <cstdio>
#include <cuda_runtime.h>

// P x Q pairs per thread, W1: sign of W1 in the vote, VS: i3 steps per vote, CSF: gain test only
template <int P, int Q, int W1, int VS, int CSF, int THREADS, int CTAS>
__global__ void __launch_bounds__(THREADS, CTAS) k(double *out, int nsteps, int TXT, int rowlen, int KC, double seed)
{
    extern __shared__ __align__(16) double smem[];
    const int tid = threadIdx.x;
    const int tx = tid % TXT, ty = tid / TXT;
    const int T1 = P * TXT;
    for (int i = tid; i < KC * rowlen + KC; i += THREADS)
        smem[i] = 0.4 * sin(seed * (i + 1) * 0.37 + blockIdx.x);
    const double *z3s = smem + KC * rowlen;
    double r12[P * Q], c33[P * Q], U1p[P * Q], U2p[P * Q], Tq[P * Q], z1r[P];
    const double c0t = 1e-13;
#pragma unroll
    for (int p = 0; p < P; p++) z1r[p] = 0.3 + 0.01 * p + 1e-4 * tx;
#pragma unroll
    for (int e = 0; e < P * Q; e++) {
        r12[e] = 0.3 * sin(seed * (tid * 8 + e));
        c33[e] = fma(-r12[e], r12[e], 1.0);
        U1p[e] = 0.2 + 1e-3 * e; U2p[e] = 0.25 - 1e-3 * e;
        Tq[e] = 50.0 + tid;                       // far above any gain: the vote never passes
    }
    __syncthreads();
    int hits = 0;
    for (int s0 = 0; s0 < nsteps; s0 += VS) {
        int sany = -1;
#pragma unroll
        for (int s = 0; s < VS; s++) {
            const double *rowp = smem + (size_t)((s0 + s) & (KC - 1)) * rowlen;
            double r13[P], r23[Q];
#pragma unroll
            for (int p = 0; p < P; p += 2) {
                if (P >= 2) { const double2 v = *reinterpret_cast<const double2 *>(rowp + P * tx + p); r13[p] = v.x; r13[p + 1 < P ? p + 1 : p] = v.y; }
                else r13[p] = rowp[tx];
            }
#pragma unroll
            for (int q = 0; q < Q; q += 2) {
                const double2 v = *reinterpret_cast<const double2 *>(rowp + T1 + Q * ty + q);
                r23[q] = v.x; r23[q + 1 < Q ? q + 1 : q] = v.y;
            }
            const double z3 = z3s[(s0 + s) & (KC - 1)];
            int sall = -1;
#pragma unroll
            for (int p = 0; p < P; p++) {
                const double m13 = fma(-r13[p], r13[p], 1.0);
                const double d13 = fma(-r13[p], z1r[p], z3);
#pragma unroll
                for (int q = 0; q < Q; q++) {
                    const int e = p * Q + q;
                    const double q2 = fma(-r12[e], r13[p], r23[q]);
                    const double S = fma(-q2, q2, c33[e] * m13);
                    const double dl = fma(-q2, U2p[e], d13);
                    const double t = fma(-Tq[e], S, fma(dl, dl, c0t));
                    if (CSF) {
                        sall &= __double2hiint(t);
                    } else {
                        const double W2 = fma(-q2, dl, U2p[e] * S);
                        if (W1) {
                            const double q1 = fma(-r12[e], r23[q], r13[p]);
                            const double w1 = fma(-q1, dl, U1p[e] * S);
                            sall &= (__double2hiint(w1) | __double2hiint(W2) | __double2hiint(dl)) | __double2hiint(t);
                        } else {
                            sall &= (__double2hiint(W2) | __double2hiint(dl)) | __double2hiint(t);
                        }
                    }
                }
            }
            sany &= sall;
        }
        if (__any_sync(0xffffffffu, sany >= 0)) hits++;
    }
    double s = hits;
#pragma unroll
    for (int e = 0; e < P * Q; e++) s += r12[e] + U1p[e] + Tq[e] * 1e-30;
    out[blockIdx.x * THREADS + tid] = s;
}

template <int P, int Q, int W1, int VS, int CSF, int THREADS, int CTAS>
void run(const char *name)
{
    double *out;
    cudaMalloc(&out, sizeof(double) * 148 * CTAS * THREADS);
    // thread layout: TXT x TYT with TXT = 32 when possible
    const int TXT = THREADS >= 256 ? 32 : 16, TYT = THREADS / TXT;
    const int rowlen = P * TXT + Q * TYT, KC = 64;
    const size_t smem = sizeof(double) * (KC * rowlen + KC);
    auto kern = k<P, Q, W1, VS, CSF, THREADS, CTAS>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncAttributes fa;
    cudaFuncGetAttributes(&fa, kern);
    const int nsteps = 40000;
    kern<<<148 * CTAS, THREADS, smem>>>(out, 64, TXT, rowlen, KC, 0.77);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    kern<<<148 * CTAS, THREADS, smem>>>(out, nsteps, TXT, rowlen, KC, 0.77);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double per_tuple = (CSF ? 6.0 : (W1 ? 11.0 : 8.0)) + 2.0 / Q;
    const double tuples = (double)nsteps * P * Q * THREADS * 148 * CTAS;
    printf("%-34s P%dxQ%d W1=%d VS=%d CSF=%d %3d thr x %d CTA, %3d regs, spill %zu B: %.2f ms, %.3f T tuples/s, %.1f ops/tuple, %.1f%% of the FP64 pipe\n",
           name, P, Q, W1, VS, CSF, THREADS, CTAS, fa.numRegs, (size_t)fa.localSizeBytes, ms, tuples / ms / 1e9,
           per_tuple, 100.0 * tuples * per_tuple * 2.0 / ms / 1e9 / 37.1);
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) printf("  CUDA error: %s\n", cudaGetErrorString(err));
    cudaFree(out);
}


// ---- the same loop (2 x 4 tile, W1 sign in the vote) written stage by stage over groups of WD
// independent tuples, so that dependent FP64 instructions are WD instructions apart.  PIN: every
// FP64 operation is an asm volatile statement (the compiler keeps their order).
__device__ __forceinline__ double pfma(double a, double b, double c)
{
    double d;
    asm volatile("fma.rn.f64 %0, %1, %2, %3;" : "=d"(d) : "d"(a), "d"(b), "d"(c));
    return d;
}
__device__ __forceinline__ double pmul(double a, double b)
{
    double d;
    asm volatile("mul.rn.f64 %0, %1, %2;" : "=d"(d) : "d"(a), "d"(b));
    return d;
}
template <int WD, int PIN, int THREADS>
__global__ void __launch_bounds__(THREADS, 1) kstage(double *out, int nsteps, int TXT, int rowlen, int KC, double seed)
{
    extern __shared__ __align__(16) double smem[];
    const int tid = threadIdx.x;
    const int tx = tid % TXT, ty = tid / TXT;
    const int T1 = 2 * TXT;
    for (int i = tid; i < KC * rowlen + KC; i += THREADS)
        smem[i] = 0.4 * sin(seed * (i + 1) * 0.37 + blockIdx.x);
    const double *z3s = smem + KC * rowlen;
    double r12[8], c33[8], U1p[8], U2p[8], Tq[8], z1r[2];
    const double c0t = 1e-13;
#pragma unroll
    for (int p = 0; p < 2; p++) z1r[p] = 0.3 + 0.01 * p + 1e-4 * tx;
#pragma unroll
    for (int e = 0; e < 8; e++) {
        r12[e] = 0.3 * sin(seed * (tid * 8 + e));
        c33[e] = fma(-r12[e], r12[e], 1.0);
        U1p[e] = 0.2 + 1e-3 * e; U2p[e] = 0.25 - 1e-3 * e;
        Tq[e] = 50.0 + tid;
    }
    __syncthreads();
    int hits = 0;
#define FMA_(a, b, c) (PIN ? pfma(a, b, c) : fma(a, b, c))
#define MUL_(a, b) (PIN ? pmul(a, b) : (a) * (b))
    for (int s0 = 0; s0 < nsteps; s0 += 4) {
        int sany = -1;
#pragma unroll
        for (int s = 0; s < 4; s++) {
            const double *rowp = smem + (size_t)((s0 + s) & (KC - 1)) * rowlen;
            const double2 r13v = *reinterpret_cast<const double2 *>(rowp + 2 * tx);
            const double2 r23a = *reinterpret_cast<const double2 *>(rowp + T1 + 4 * ty);
            const double2 r23b = *reinterpret_cast<const double2 *>(rowp + T1 + 4 * ty + 2);
            const double z3 = z3s[(s0 + s) & (KC - 1)];
            const double r13[2] = {r13v.x, r13v.y};
            const double r23[4] = {r23a.x, r23a.y, r23b.x, r23b.y};
            double m13[2], d13[2];
#pragma unroll
            for (int p = 0; p < 2; p++) { m13[p] = FMA_(-r13[p], r13[p], 1.0); d13[p] = FMA_(-r13[p], z1r[p], z3); }
            int sall = -1;
#pragma unroll
            for (int g0 = 0; g0 < 8; g0 += WD) {
                double q2[WD], q1[WD], cm[WD], S[WD], dl[WD], t1[WD], u1[WD], u2[WD], t[WD], W1[WD], W2[WD];
#pragma unroll
                for (int i = 0; i < WD; i++) { const int e = g0 + i; q2[i] = FMA_(-r12[e], r13[e >> 2], r23[e & 3]); }
#pragma unroll
                for (int i = 0; i < WD; i++) { const int e = g0 + i; cm[i] = MUL_(c33[e], m13[e >> 2]); }
#pragma unroll
                for (int i = 0; i < WD; i++) { const int e = g0 + i; q1[i] = FMA_(-r12[e], r23[e & 3], r13[e >> 2]); }
#pragma unroll
                for (int i = 0; i < WD; i++) { S[i] = FMA_(-q2[i], q2[i], cm[i]); }
#pragma unroll
                for (int i = 0; i < WD; i++) { const int e = g0 + i; dl[i] = FMA_(-q2[i], U2p[e], d13[e >> 2]); }
#pragma unroll
                for (int i = 0; i < WD; i++) { const int e = g0 + i; u1[i] = MUL_(U1p[e], S[i]); }
#pragma unroll
                for (int i = 0; i < WD; i++) { const int e = g0 + i; u2[i] = MUL_(U2p[e], S[i]); }
#pragma unroll
                for (int i = 0; i < WD; i++) { t1[i] = FMA_(dl[i], fabs(dl[i]), c0t); }
#pragma unroll
                for (int i = 0; i < WD; i++) { W1[i] = FMA_(-q1[i], dl[i], u1[i]); }
#pragma unroll
                for (int i = 0; i < WD; i++) { W2[i] = FMA_(-q2[i], dl[i], u2[i]); }
#pragma unroll
                for (int i = 0; i < WD; i++) { const int e = g0 + i; t[i] = FMA_(-Tq[e], S[i], t1[i]); }
#pragma unroll
                for (int i = 0; i < WD; i++) sall &= (__double2hiint(W1[i]) | __double2hiint(W2[i])) | __double2hiint(t[i]);
            }
            sany &= sall;
        }
        if (__any_sync(0xffffffffu, sany >= 0)) hits++;
    }
    double sres = hits;
#pragma unroll
    for (int e = 0; e < 8; e++) sres += r12[e] + U1p[e] + Tq[e] * 1e-30;
    out[blockIdx.x * THREADS + tid] = sres;
}

template <int WD, int PIN, int THREADS>
void runstage(const char *name)
{
    double *out;
    cudaMalloc(&out, sizeof(double) * 148 * THREADS);
    const int TXT = 32, TYT = THREADS / TXT;
    const int rowlen = 2 * TXT + 4 * TYT, KC = 64;
    const size_t smem = sizeof(double) * (KC * rowlen + KC);
    auto kern = kstage<WD, PIN, THREADS>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncAttributes fa;
    cudaFuncGetAttributes(&fa, kern);
    const int nsteps = 40000;
    kern<<<148, THREADS, smem>>>(out, 64, TXT, rowlen, KC, 0.77);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    kern<<<148, THREADS, smem>>>(out, nsteps, TXT, rowlen, KC, 0.77);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double tuples = (double)nsteps * 8 * THREADS * 148;
    printf("%-28s width %d pinned %d %3d thr, %3d regs, spill %zu B: %.2f ms, %.3f T tuples/s, %.1f%% of the FP64 pipe at 11.5 ops\n",
           name, WD, PIN, THREADS, fa.numRegs, (size_t)fa.localSizeBytes, ms, tuples / ms / 1e9,
           100.0 * tuples * 23.0 / ms / 1e9 / 37.1);
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) printf("  CUDA error: %s\n", cudaGetErrorString(err));
    cudaFree(out);
}

// dependent DFMA chains: ILP independent chains per warp, W warps per SM
template <int ILP>
__global__ void kchain(double *out, int iters, double a, double b)
{
    double c[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) c[i] = 1e-3 * (i + threadIdx.x);
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 8; r++)
#pragma unroll
            for (int i = 0; i < ILP; i++) c[i] = fma(c[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int ILP>
void runchain(int warps)
{
    double *out;
    cudaMalloc(&out, sizeof(double) * 148 * 1024);
    const int iters = 20000;
    kchain<ILP><<<148, 32 * warps>>>(out, 10, 0.999, 1e-3);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    kchain<ILP><<<148, 32 * warps>>>(out, iters, 0.999, 1e-3);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double inst = (double)iters * 8 * ILP * warps / 4;      // warp instructions per SM sub-partition
    printf("chains: ILP %d, %2d warps/SM: %.2f TFLOP/s, %.2f clk per DFMA per sub-partition (1.965 GHz)\n", ILP, warps,
           2.0 * iters * 8 * ILP * 32 * warps * 148 / ms / 1e9, ms * 1e-3 * 1.965e9 / inst);
    cudaFree(out);
}

int main()
{
    for (int w : {12}) { runchain<1>(w); runchain<2>(w); runchain<4>(w); }
    runstage<4, 0, 384>("stage-wise"); runstage<8, 0, 384>("stage-wise"); runstage<2, 0, 384>("stage-wise");
    runstage<4, 1, 384>("stage-wise pinned"); runstage<8, 1, 384>("stage-wise pinned"); runstage<2, 1, 384>("stage-wise pinned");
    runstage<4, 1, 256>("stage-wise pinned"); runstage<8, 1, 256>("stage-wise pinned");
    run<2, 4, 1, 4, 0, 384, 1>("shipped shape");
    run<2, 4, 1, 4, 0, 192, 2>("2 CTAs");
    run<2, 4, 1, 8, 0, 384, 1>("8 steps per vote");
    run<2, 4, 1, 2, 0, 384, 1>("2 steps per vote");
    run<2, 4, 1, 1, 0, 384, 1>("1 step per vote");
    run<2, 4, 0, 4, 0, 384, 1>("no W1 sign");
    run<2, 4, 0, 4, 1, 384, 1>("CSF (gain only)");
    run<2, 2, 1, 4, 0, 768, 1>("2x2, 24 warps");
    run<2, 2, 1, 4, 0, 384, 2>("2x2, 24 warps, 2 CTAs");
    run<2, 2, 1, 8, 0, 768, 1>("2x2, 24 warps, 8 steps");
    run<2, 2, 1, 4, 0, 512, 1>("2x2, 16 warps");
    run<2, 2, 0, 4, 0, 768, 1>("2x2, 24 warps, no W1");
    run<1, 4, 1, 4, 0, 768, 1>("1x4, 24 warps");
    run<4, 4, 1, 2, 0, 256, 1>("4x4, 8 warps");
    run<2, 6, 1, 4, 0, 256, 1>("2x6, 8 warps");
    run<2, 4, 1, 4, 0, 256, 1>("2x4, 8 warps");
    run<2, 4, 1, 4, 0, 512, 1>("2x4, 16 warps (128 regs)");
    return 0;
}
