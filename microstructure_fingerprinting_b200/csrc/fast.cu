// fast.cu -- the screening ("fast") tier of libmfb200 for 2-fascicle voxels, sm_100a.
//
// For a voxel with two fascicles (optionally + the CSF column) the reference
// (mf_utils.py:288-392 `_2`, 470-607 `_3`) forms the N x N cross-Gram of the two rotated
// sub-dictionaries over the M measurements and solves a closed-form 2- or 3-variable NNLS
// per atom pair.  Here:
//   k_fast_prep   rotates each sub-dictionary once per voxel to get the per-atom
//                 statistics (|a|^2, a.y, a.csf), stores the interpolation plan, and
//                 reduces the best single-atom / atom+CSF gains (the pair-independent
//                 branches of the NNLS);
//   k_fast_pairs  one CTA per (voxel, 128-atom i1 tile): the rotated, CSF-projected and
//                 normalised i1 tile stays resident in shared memory, i2 tiles of 32 atoms
//                 are gathered from the L2-resident lookup table (register-prefetched,
//                 double-buffered), the correlation tile is formed with FP64 tensor-core
//                 DMMA (mma.sync.m8n8k4.f64) and consumed in registers by a division-free
//                 closed-form NNLS + argmax epilogue;
//   k_fast_select merges the tiles of a voxel and decides whether the winner is certain.
// Screening works on gains (|y|^2 - residual) in a different summation order than the
// reference, so it only *selects*: the winning tuple is re-evaluated in the reference's
// arithmetic by the exact tier's evaluate kernel, and every voxel whose winner is not
// separated from the runner-up (or from a pair-independent branch) by more than the
// screening error bound is handed to the exact tier.
#include <climits>
#include <cstdlib>
#include <cstring>

#include "common.cuh"

namespace mfb {

#define FT_TI 128
#define FT_TJ 32
#define FT_CONS 256         // consumer threads (8 DMMA warps)
#define FT_PROD 128         // producer threads (4 gather warps)
#define FT_THREADS (FT_CONS + FT_PROD)
#define FT_NS 3             // stages of the i2-tile ring
#define FT_S1 (FT_TI + 4)  // row strides: == 4 mod 16 doubles -> conflict-free fragment loads
#define FT_S2 (FT_TJ + 4)
#define FT_NPAR 7          // per-atom parameters: scale, alpha, z, beta, kappa, gamma, zu

// 1 - rho^2 below which a pair is tracked as ill-conditioned (its screening error bound
// c0 / det is no longer small against typical gaps between competing pairs)
static constexpr double kIllDet = 1e-4;

struct FastArgs {
    DevPlan p;         // table source: rotation plan; explicit source: only p.M is used
    int src;           // 0: sub-dictionaries rotated from the lookup table (fit path)
                       // 1: explicit dictionaries A (mfb_solve_batch)
    int N1, N2;        // atoms of the two searched blocks
    const double *A;   // explicit source: voxel row r reads A + r*strideA, (M, lda) row-major
    int64_t lda, strideA;
    int start1, start2, start3;  // first column of block 1, block 2 and of the third (1-column) block
    int a_by_local;    // explicit source: the dictionary of local voxel v is A + v*strideA (rows index y / tuple only)
    double *Dn;        // k_gemm_pairs: normalised (and CSF-projected) copy, [v][Mp2][ldn], zero padded
    int64_t dn_stride;
    int ldn, N1pad, Mp2;
    int32_t *redo_local;  // local indices of the voxels handed to the exact tier
    int csf;
    int Mp;            // M padded to a multiple of 4
    int Npad;          // max(N1, N2) padded to a multiple of FT_TJ
    int ntI;           // i1 tiles per voxel
    int debug;            // timing experiments only (MFB_FAST_DEBUG): 1 skip epilogue, 2 skip gathers
    const int32_t *vox_list;
    const double *peaks;
    int peaks_ld;
    const double *y;
    int *ip_rows;      // [v][2][M][2]  rl, rh
    double *ip_w;      // [v][2][M][2]  wl, wh
    double *colp;      // [v][2][FT_NPAR][Npad]
    double *voxp;      // [v][8]  y_sq, A33, Y3, gain_c, c0, Gpre(fasc 0), Gpre(fasc 1), -
    double *cta_gain;  // [v][ntI]
    double *cta_tol;   // [v][ntI]
    int *cta_idx;      // [v][ntI]
    int *cta_flag;     // [v][ntI]
    double *cta_ill;   // [v][ntI]  best optimistic gain among ill-conditioned pairs
    long long *tuple;  // [row]
    int32_t *redo_list;  // voxels handed to the exact tier
    int32_t *redo_count;
    int32_t *reasons;    // [4] why voxels were handed over: no pair, ill-conditioned, near tie, pair-independent branch
};

__device__ __forceinline__ int search_left(const double *xs, int n, double x)
{
    int lo = 0, hi = n;
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (xs[mid] < x) lo = mid + 1; else hi = mid;
    }
    return lo;
}

__device__ __forceinline__ double block_max(double v, double *sm)
{
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
    __syncthreads();
    double r = sm[0];
    for (int w = 1; w < (int)(blockDim.x >> 5); w++) r = fmax(r, sm[w]);
    return r;
}

// Gain (|y|^2 - residual) of the 2-variable NNLS in Gram form (screening precision).
__device__ __forceinline__ double nnls2_gain(double A11, double A12, double A22, double Y1, double Y2)
{
    double w1d = A22 * Y1 - A12 * Y2, w2d = A11 * Y2 - A12 * Y1;
    double g1 = Y1 > 0 ? Y1 * Y1 / A11 : 0.0, g2 = Y2 > 0 ? Y2 * Y2 / A22 : 0.0;
    if (w1d > 0 && w2d > 0) {
        double det = A11 * A22 - A12 * A12;
        if (det > 0) return fmax((Y1 * w1d + Y2 * w2d) / det, fmax(g1, g2));
    }
    return fmax(g1, g2);
}

// ---------------------------------------------------------------------------------
// prepass: grid (V, 2), block 256
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_fast_prep(FastArgs a)
{
    extern __shared__ double sm[];
    const DevPlan &p = a.p;
    const int M = p.M;
    double *wl = sm, *wh = sm + M, *ys = sm + 2 * M, *cs = sm + 3 * M, *red = sm + 4 * M;
    int *rl = (int *)(red + 32), *rh = rl + M;
    const int64_t v = blockIdx.x;
    const int k = blockIdx.y;
    const int Nk = k ? a.N2 : a.N1;
    const int64_t row = a.vox_list ? a.vox_list[v] : v;
    const double *Ar = a.src ? a.A + (a.a_by_local ? v : row) * a.strideA : nullptr;
    if (!a.src) {
        const double *u = a.peaks + row * a.peaks_ld + 3 * k;
        const double ux = u[0], uy = u[1], uz = u[2];
        for (int m = threadIdx.x; m < M; m += blockDim.x) {
            // same expression order as the exact tier (exact.cu dir_dot / shell_lerp)
            double x = fabs(__dadd_rn(__dadd_rn(__dmul_rn(p.gdir[3 * m], ux), __dmul_rn(p.gdir[3 * m + 1], uy)),
                                      __dmul_rn(p.gdir[3 * m + 2], uz)));
            int s = p.shell_lo[m];
            const double *xs = p.nodes + p.off[s];
            int n = p.off[s + 1] - p.off[s];
            int j = search_left(xs, n, x);
            j = j < 1 ? 1 : (j > n - 1 ? n - 1 : j);
            double den = __dsub_rn(xs[j], xs[j - 1]);
            rl[m] = p.off[s] + j - 1;
            rh[m] = p.off[s] + j;
            wh[m] = __ddiv_rn(__dsub_rn(x, xs[j - 1]), den);
            wl[m] = __ddiv_rn(__dsub_rn(xs[j], x), den);
            int64_t o = ((v * 2 + k) * M + m) * 2;
            a.ip_rows[o] = rl[m]; a.ip_rows[o + 1] = rh[m];
            a.ip_w[o] = wl[m]; a.ip_w[o + 1] = wh[m];
        }
    }
    for (int m = threadIdx.x; m < M; m += blockDim.x) {
        ys[m] = a.y[row * M + m];
        cs[m] = !a.csf ? 0.0 : (a.src ? Ar[(size_t)m * a.lda + a.start3] : p.sig_csf[m]);
    }
    __syncthreads();
    double y_sq = 0.0, A33 = 0.0, Y3 = 0.0;
    for (int m = 0; m < M; m++) {
        y_sq = fma(ys[m], ys[m], y_sq);
        A33 = fma(cs[m], cs[m], A33);
        Y3 = fma(cs[m], ys[m], Y3);
    }
    const double gain_c = (a.csf && Y3 > 0) ? Y3 * Y3 / A33 : 0.0;
    double *cp = a.colp + (v * 2 + k) * (int64_t)FT_NPAR * a.Npad;
    double gbest = 0.0;
    for (int i = threadIdx.x; i < a.Npad; i += blockDim.x) {
        double par[FT_NPAR] = {0, 0, 0, 0, 0, 0, 0};
        if (i < Nk) {
            double sq = 0.0, dy = 0.0, d3 = 0.0;
            const double *Ac = a.src ? Ar + (k ? a.start2 : a.start1) + i : nullptr;
            for (int m = 0; m < M; m++) {
                double d = a.src ? Ac[(size_t)m * a.lda]
                                 : fma(wh[m], p.table[(size_t)rh[m] * p.N + i], wl[m] * p.table[(size_t)rl[m] * p.N + i]);
                sq = fma(d, d, sq);
                dy = fma(d, ys[m], dy);
                d3 = fma(d, cs[m], d3);
            }
            const double r = rsqrt(sq);
            if (!a.csf) {
                par[0] = r;            // scale
                par[2] = dy * r;       // z
                gbest = fmax(gbest, dy > 0 ? dy * dy / sq : 0.0);
            } else {
                const double gam = d3 * r * rsqrt(A33);      // corr(atom, csf)
                const double kap2 = fmax(1.0 - gam * gam, 1e-300);
                const double kap = sqrt(kap2);
                const double rp = r / kap;                   // 1/|a projected|
                const double alpha = d3 / A33;
                par[0] = rp;
                par[1] = alpha;
                par[2] = (dy - alpha * Y3) * rp;             // z'
                par[3] = d3 * rp;                            // beta
                par[4] = kap;
                par[5] = gam;
                par[6] = dy * r;                             // zu
                gbest = fmax(gbest, nnls2_gain(sq, d3, A33, dy, Y3));
            }
        }
        for (int q = 0; q < FT_NPAR; q++) cp[(size_t)q * a.Npad + i] = par[q];
    }
    gbest = block_max(gbest, red);
    if (threadIdx.x == 0) {
        double *vp = a.voxp + v * 8;
        vp[5 + k] = fmax(gbest, gain_c);
        if (k == 0) {
            vp[0] = y_sq; vp[1] = A33; vp[2] = Y3; vp[3] = gain_c;
            vp[4] = 4.0 * (M + 8) * 2.2204e-16 * y_sq;   // c0: screening error scale
        }
    }
}

// ---------------------------------------------------------------------------------
// pair scan
// ---------------------------------------------------------------------------------
// Screening quantities of one (i1, i2) pair from its correlation rho (registers only).
// Returns false when the pair has no both-positive closed form worth tracking.
template <int CSF>
__device__ __forceinline__ bool pair_gain(double rho, double z1, double z2, double b1, double b2,
                                          double k1, double k2, double g1, double g2, double zu1,
                                          double zu2, double Y3, double gain_c, double &num,
                                          double &det)
{
    const double w1 = fma(-rho, z2, z1);
    const double w2 = fma(-rho, z1, z2);
    det = fma(-rho, rho, 1.0);
    num = fma(z1, w1, z2 * w2);
    bool pos = min(__double2hiint(w1), __double2hiint(w2)) > 0;
    if (CSF) {
        const double w3 = fma(-b2, w2, fma(-b1, w1, Y3 * det));
        pos = pos && (__double2hiint(w3) > 0);
        num = fma(gain_c, det, num);
        if (!pos) {
            // best of the 2-column sub-problems: only the fascicle pair depends on (i1, i2);
            // the atom + CSF ones are pair-independent and live in gpre
            const double r = fma(rho * k1, k2, g1 * g2);
            const double v1 = fma(-r, zu2, zu1);
            const double v2 = fma(-r, zu1, zu2);
            pos = min(__double2hiint(v1), __double2hiint(v2)) > 0;
            det = fma(-r, r, 1.0);
            num = fma(zu1, v1, zu2 * v2);
        }
    }
    return pos;
}

// ---- mbarrier helpers (producer / consumer ring) ----
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count)
{
    unsigned addr = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(addr), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(unsigned long long *bar)
{
    unsigned addr = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(addr) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity)
{
    unsigned addr = (unsigned)__cvta_generic_to_shared(bar);
    unsigned ok;
    do {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
                     " selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void consumer_sync()
{
    asm volatile("bar.sync 1, %0;" ::"n"(FT_CONS) : "memory");
}

// Warp-specialised pair scan.  Warps 0..7 (consumers): DMMA correlation tile + closed-form
// epilogue; warps 8..9 (producers): gather the next i2 tiles from the L2-resident lookup
// table, rotate / project / normalise them and fill a 3-stage shared-memory ring.  full[] /
// empty[] mbarriers are the only synchronisation inside the tile loop, so consumer warps
// drift apart and one warp's scalar epilogue overlaps another warp's DMMA stream.
template <int CSF, int SRC>
__global__ void __launch_bounds__(FT_THREADS, 1) k_fast_pairs(FastArgs a)
{
    extern __shared__ __align__(16) double smem[];
    const DevPlan &p = a.p;
    const int M = p.M, N = p.N, Mp = a.Mp;
    const int N1 = a.N1, N2 = a.N2;
    double *D1s = smem;                                   // [Mp][FT_S1]
    double *D2s = D1s + (size_t)Mp * FT_S1;               // [FT_NS][Mp][FT_S2]
    double *colq = D2s + (size_t)FT_NS * Mp * FT_S2;      // [FT_NS][5][FT_TJ]  z, beta, kappa, gamma, zu
    double *w1l = colq + FT_NS * 5 * FT_TJ;               // [Mp] plan of fascicle 1
    double *w1h = w1l + Mp;
    double *w2l = w1h + Mp;                               // [Mp] plan of fascicle 2
    double *w2h = w2l + Mp;
    double *cs = w2h + Mp;                                // [Mp] csf column
    double *red = cs + Mp;                                // [64]
    int *r1l = (int *)(red + 64);
    int *r1h = r1l + Mp;
    int *r2l = r1h + Mp;
    int *r2h = r2l + Mp;
    __shared__ unsigned long long s_thr;                  // CTA-wide lower bound on the winning gain
    __shared__ unsigned long long s_full[FT_NS], s_empty[FT_NS];
    __shared__ double s_tolG;
    __shared__ int s_flag;

    const int64_t v = blockIdx.y;
    const int tI = blockIdx.x;
    const int i0 = tI * FT_TI;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const double *vp = a.voxp + v * 8;
    const double gain_c = vp[3], c0 = vp[4], Y3 = vp[2];
    const double gpre = fmax(vp[5], vp[6]);
    const double *cp1 = a.colp + (v * 2 + 0) * (int64_t)FT_NPAR * a.Npad;
    const double *cp2 = a.colp + (v * 2 + 1) * (int64_t)FT_NPAR * a.Npad;
    const int ntJ = (N2 + FT_TJ - 1) / FT_TJ;
    const int64_t row = a.vox_list ? a.vox_list[v] : v;
    const double *Ar = SRC ? a.A + (a.a_by_local ? v : row) * a.strideA : nullptr;

    for (int m = tid; m < Mp; m += FT_THREADS) {
        if (m < M && SRC) {
            // explicit dictionaries: rows are read directly (weights 1 / 0, row index = m)
            r1l[m] = r1h[m] = r2l[m] = r2h[m] = m;
            w1l[m] = w2l[m] = 0.0; w1h[m] = w2h[m] = 1.0;
            cs[m] = CSF ? Ar[(size_t)m * a.lda + a.start3] : 0.0;
        } else if (m < M) {
            int64_t o = ((v * 2 + 0) * M + m) * 2, o2 = ((v * 2 + 1) * M + m) * 2;
            r1l[m] = a.ip_rows[o]; r1h[m] = a.ip_rows[o + 1];
            w1l[m] = a.ip_w[o]; w1h[m] = a.ip_w[o + 1];
            r2l[m] = a.ip_rows[o2]; r2h[m] = a.ip_rows[o2 + 1];
            w2l[m] = a.ip_w[o2]; w2h[m] = a.ip_w[o2 + 1];
            cs[m] = CSF ? p.sig_csf[m] : 0.0;
        } else {
            r1l[m] = r1h[m] = r2l[m] = r2h[m] = 0;
            w1l[m] = w1h[m] = w2l[m] = w2h[m] = 0.0; cs[m] = 0.0;
        }
    }
    if (tid == 0) {
        s_thr = (unsigned long long)__double_as_longlong(fmax(gpre - c0, 0.0));
        s_flag = 0;
        for (int st = 0; st < FT_NS; st++) { mbar_init(&s_full[st], FT_PROD); mbar_init(&s_empty[st], FT_CONS); }
    }
    __syncthreads();

    if (tid >= FT_CONS) {
        // =========================== producers ===========================
        const int pt = tid - FT_CONS;
        const int jj = pt & (FT_TJ - 1);
        const int mrow0 = pt / FT_TJ;                     // 0..FT_PROD/32-1
        constexpr int RS = FT_PROD / FT_TJ;               // rows per pass
        constexpr int UB = 14;                            // (lo, hi) pairs in flight per thread
        for (int jt = 0; jt < ntJ; jt++) {
            const int st = jt % FT_NS;
            if (jt >= FT_NS) mbar_wait(&s_empty[st], (unsigned)((jt / FT_NS) - 1) & 1u);
            if ((a.debug & 2) && jt >= FT_NS) { mbar_arrive(&s_full[st]); continue; }
            const int j = jt * FT_TJ + jj;
            const bool ok = j < N2;
            const double csc = ok ? __ldg(cp2 + j) : 0.0;
            const double cal = (CSF && ok) ? __ldg(cp2 + (size_t)a.Npad + j) : 0.0;
            // source rows: lookup table (stride N) or this voxel's dictionary (stride lda)
            const double *Tc = SRC ? Ar + a.start2 + (ok ? j : 0) : p.table + (ok ? j : 0);
            const size_t rs = SRC ? (size_t)a.lda : (size_t)N;
            double *dst = D2s + (size_t)st * Mp * FT_S2 + jj;
            for (int mb = mrow0; mb < Mp; mb += RS * UB) {
                double lo[UB], hi[UB];
#pragma unroll
                for (int q = 0; q < UB; q++) {
                    const int m = min(mb + RS * q, Mp - 1);   // rows >= M carry zero weights
                    if (SRC) {                                 // explicit dictionary: row m itself
                        lo[q] = 0.0;
                        hi[q] = __ldg(Tc + (size_t)min(m, M - 1) * rs);
                    } else {
                        lo[q] = __ldg(Tc + (size_t)r2l[m] * rs);
                        hi[q] = __ldg(Tc + (size_t)r2h[m] * rs);
                    }
                }
                // three passes of independent FP64 ops (the FP64 pipe is shared with the
                // consumers' DMMA stream: dependent chains would serialise on its latency)
                if (!SRC) {
#pragma unroll
                    for (int q = 0; q < UB; q++) lo[q] *= w2l[min(mb + RS * q, Mp - 1)];
                }
#pragma unroll
                for (int q = 0; q < UB; q++) {
                    const int m = min(mb + RS * q, Mp - 1);
                    hi[q] = SRC ? w2h[m] * hi[q] : fma(w2h[m], hi[q], lo[q]);   // w2h = 0 on padding rows
                    if (CSF) hi[q] = fma(-cal, cs[m], hi[q]);
                }
#pragma unroll
                for (int q = 0; q < UB; q++) hi[q] *= csc;
#pragma unroll
                for (int q = 0; q < UB; q++) {
                    const int m = mb + RS * q;
                    if (m < Mp) dst[(size_t)m * FT_S2] = hi[q];
                }
            }
            for (int e = pt; e < 5 * FT_TJ; e += FT_PROD)
                colq[(st * 5 + e / FT_TJ) * FT_TJ + (e % FT_TJ)] =
                    __ldg(cp2 + (size_t)(e / FT_TJ + 2) * a.Npad + jt * FT_TJ + (e % FT_TJ));
            mbar_arrive(&s_full[st]);
        }
        return;
    }

    // =============================== consumers ===============================
    const int g = lane >> 2, t4 = lane & 3;
    const double wide = 4.0 * c0 / kIllDet;
    const double negc0 = -c0;
    // ---- resident i1 tile: rotate, project out the CSF column, normalise ----
    {
        const int ii = tid & (FT_TI - 1);
        const int i = i0 + ii;
        const bool ok = i < N1;
        const double sc = ok ? cp1[i] : 0.0;
        const double al = (CSF && ok) ? cp1[(size_t)a.Npad + i] : 0.0;
        const double *Tc = SRC ? Ar + a.start1 + (ok ? i : 0) : p.table + (ok ? i : 0);
        const size_t rs = SRC ? (size_t)a.lda : (size_t)N;
        constexpr int RS = FT_CONS / FT_TI;               // rows per pass (2)
        constexpr int UB = 18;                             // loads in flight per thread: 2*UB
        for (int mb = tid / FT_TI; mb < Mp; mb += RS * UB) {
            double lo[UB], hi[UB];
#pragma unroll
            for (int q = 0; q < UB; q++) {
                const int m = mb + RS * q;
                lo[q] = 0.0; hi[q] = 0.0;
                if (ok && m < M) {
                    if (!SRC) lo[q] = __ldg(Tc + (size_t)r1l[m] * rs);
                    hi[q] = __ldg(Tc + (size_t)r1h[m] * rs);
                }
            }
#pragma unroll
            for (int q = 0; q < UB; q++) {
                const int m = mb + RS * q;
                if (m < Mp) {
                    double d = fma(w1h[m], hi[q], w1l[m] * lo[q]);
                    if (CSF) d = fma(-al, cs[m], d);
                    D1s[(size_t)m * FT_S1 + ii] = d * sc;
                }
            }
        }
    }
    consumer_sync();

    // ---- per-thread row constants (rows g and g+8 of the warp's 16-row slab) ----
    const int wrow = warp * 16;
    double z1[2], b1[2], k1[2], g1[2], zu1[2];
#pragma unroll
    for (int mt = 0; mt < 2; mt++) {
        const int i = i0 + wrow + 8 * mt + g;
        const bool ok = i < N1;
        z1[mt] = ok ? cp1[(size_t)2 * a.Npad + i] : 0.0;
        b1[mt] = (CSF && ok) ? cp1[(size_t)3 * a.Npad + i] : 0.0;
        k1[mt] = (CSF && ok) ? cp1[(size_t)4 * a.Npad + i] : 0.0;
        g1[mt] = (CSF && ok) ? cp1[(size_t)5 * a.Npad + i] : 0.0;
        zu1[mt] = (CSF && ok) ? cp1[(size_t)6 * a.Npad + i] : 0.0;
    }

    const int mtv = max(0, min(2, (N1 - (i0 + wrow) + 7) >> 3));   // valid 8-row blocks of this warp

    // thread-local best (central gain gb, tolerance tb) and the shared screening threshold
    double gb = -1.0, tb = 0.0, thr = fmax(gpre - c0, 0.0), gill = -1.0;
    int bidx = -1, flag = 0;

    for (int jt = 0; jt < ntJ; jt++) {
        const int st = jt % FT_NS;
        mbar_wait(&s_full[st], (unsigned)(jt / FT_NS) & 1u);
        thr = fmax(thr, __longlong_as_double((long long)s_thr));

        // ---- correlation tile: 16 x 32 per warp, DMMA m8n8k4 over k ----
        double acc[2][4][2];
#pragma unroll
        for (int mt = 0; mt < 2; mt++)
#pragma unroll
            for (int nt = 0; nt < 4; nt++) { acc[mt][nt][0] = 0.0; acc[mt][nt][1] = 0.0; }
        const double *A_ = D1s + (size_t)t4 * FT_S1 + wrow + g;
        const double *B_ = D2s + (size_t)st * Mp * FT_S2 + (size_t)t4 * FT_S2 + g;
        // 8-atom blocks of this tile that hold real atoms (warp-uniform): the last i1 / i2
        // tiles of a dictionary whose size is not a multiple of the tile are partly empty
        const int ntv = min(4, (N2 - jt * FT_TJ + 7) >> 3);
        if (ntv == 4 && mtv == 2) {
#pragma unroll 3
            for (int ks = 0; ks < Mp / 4; ks++) {
                double af[2], bf[4];
#pragma unroll
                for (int mt = 0; mt < 2; mt++) af[mt] = A_[(size_t)ks * 4 * FT_S1 + 8 * mt];
#pragma unroll
                for (int nt = 0; nt < 4; nt++) bf[nt] = B_[(size_t)ks * 4 * FT_S2 + 8 * nt];
#pragma unroll
                for (int mt = 0; mt < 2; mt++)
#pragma unroll
                    for (int nt = 0; nt < 4; nt++)
                        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                                     : "+d"(acc[mt][nt][0]), "+d"(acc[mt][nt][1])
                                     : "d"(af[mt]), "d"(bf[nt]));
            }
        } else {
#pragma unroll 1
            for (int ks = 0; ks < Mp / 4; ks++) {
                double af[2], bf[4];
#pragma unroll
                for (int mt = 0; mt < 2; mt++) af[mt] = A_[(size_t)ks * 4 * FT_S1 + 8 * mt];
#pragma unroll
                for (int nt = 0; nt < 4; nt++) bf[nt] = B_[(size_t)ks * 4 * FT_S2 + 8 * nt];
#pragma unroll
                for (int mt = 0; mt < 2; mt++)
#pragma unroll
                    for (int nt = 0; nt < 4; nt++)
                        if (mt < mtv && nt < ntv)
                            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                                         : "+d"(acc[mt][nt][0]), "+d"(acc[mt][nt][1])
                                         : "d"(af[mt]), "d"(bf[nt]));
            }
        }

        // ---- closed-form NNLS screening, branch-free over the thread's 16 pairs ----
        // (num + c0)/det >= thr  <=>  fma(-thr, det, num) >= -c0
        const double *cq = colq + st * 5 * FT_TJ;
        unsigned hit = 0;
        if (a.debug & 1) {
            double sacc = 0.0;
#pragma unroll
            for (int mt = 0; mt < 2; mt++)
#pragma unroll
                for (int nt = 0; nt < 4; nt++) sacc += acc[mt][nt][0] + acc[mt][nt][1];
            if (sacc == 1.2345e300) hit = 1;
        } else if (!CSF) {
#pragma unroll
            for (int nt = 0; nt < 4; nt++) {
                const double2 z2v = *reinterpret_cast<const double2 *>(cq + 8 * nt + 2 * t4);
#pragma unroll
                for (int e = 0; e < 2; e++)
#pragma unroll
                    for (int mt = 0; mt < 2; mt++) {
                        const double rho = acc[mt][nt][e], z2 = e ? z2v.y : z2v.x;
                        const double w1 = fma(-rho, z2, z1[mt]);
                        const double w2 = fma(-rho, z1[mt], z2);
                        const double det = fma(-rho, rho, 1.0);
                        const double num = fma(z1[mt], w1, z2 * w2);
                        const bool pos = min(__double2hiint(w1), __double2hiint(w2)) > 0;
                        if (pos && fma(-thr, det, num) >= negc0) hit |= 1u << (nt * 4 + e * 2 + mt);
                    }
            }
        } else {
            // pass 1: three-compartment closed form (fascicle pair projected off the CSF column)
            unsigned fb = 0;   // pairs whose 3-variable solution has a non-positive weight
#pragma unroll
            for (int nt = 0; nt < 4; nt++) {
                const double2 z2v = *reinterpret_cast<const double2 *>(cq + 8 * nt + 2 * t4);
                const double2 b2v = *reinterpret_cast<const double2 *>(cq + FT_TJ + 8 * nt + 2 * t4);
#pragma unroll
                for (int e = 0; e < 2; e++)
#pragma unroll
                    for (int mt = 0; mt < 2; mt++) {
                        const double rho = acc[mt][nt][e], z2 = e ? z2v.y : z2v.x, b2 = e ? b2v.y : b2v.x;
                        const double w1 = fma(-rho, z2, z1[mt]);
                        const double w2 = fma(-rho, z1[mt], z2);
                        const double det = fma(-rho, rho, 1.0);
                        const double w3 = fma(-b2, w2, fma(-b1[mt], w1, Y3 * det));
                        const double num = fma(gain_c, det, fma(z1[mt], w1, z2 * w2));
                        const bool pos = min(min(__double2hiint(w1), __double2hiint(w2)), __double2hiint(w3)) > 0;
                        const unsigned bit = 1u << (nt * 4 + e * 2 + mt);
                        if (!pos) fb |= bit;
                        else if (fma(-thr, det, num) >= negc0) hit |= bit;
                    }
            }
            // pass 2: best 2-column sub-problem of the pairs that fell back (only the fascicle
            // pair depends on (i1, i2); the atom + CSF ones are pair-independent, in gpre)
            if (__any_sync(0xffffffffu, fb != 0)) {
#pragma unroll
                for (int nt = 0; nt < 4; nt++) {
                    const double2 k2v = *reinterpret_cast<const double2 *>(cq + 2 * FT_TJ + 8 * nt + 2 * t4);
                    const double2 g2v = *reinterpret_cast<const double2 *>(cq + 3 * FT_TJ + 8 * nt + 2 * t4);
                    const double2 zu2v = *reinterpret_cast<const double2 *>(cq + 4 * FT_TJ + 8 * nt + 2 * t4);
#pragma unroll
                    for (int e = 0; e < 2; e++)
#pragma unroll
                        for (int mt = 0; mt < 2; mt++) {
                            const double rho = acc[mt][nt][e];
                            const double zu2 = e ? zu2v.y : zu2v.x;
                            const double r = fma(rho * k1[mt], e ? k2v.y : k2v.x, g1[mt] * (e ? g2v.y : g2v.x));
                            const double v1 = fma(-r, zu2, zu1[mt]);
                            const double v2 = fma(-r, zu1[mt], zu2);
                            const double det = fma(-r, r, 1.0);
                            const double num = fma(zu1[mt], v1, zu2 * v2);
                            const bool pos = min(__double2hiint(v1), __double2hiint(v2)) > 0;
                            const unsigned bit = 1u << (nt * 4 + e * 2 + mt);
                            if ((fb & bit) && pos && fma(-thr, det, num) >= negc0) hit |= bit;
                        }
                }
            }
        }
        // ---- rare: some lane of the warp has a competitive pair ----
        if (__any_sync(0xffffffffu, hit != 0)) {
            if (hit) {
                // off the hot path: one instantiation of the closed form, accumulators read
                // back through a (local-memory) copy indexed at run time
                double rcopy[16];
#pragma unroll
                for (int nt = 0; nt < 4; nt++)
#pragma unroll
                    for (int e = 0; e < 2; e++)
#pragma unroll
                        for (int mt = 0; mt < 2; mt++) rcopy[nt * 4 + e * 2 + mt] = acc[mt][nt][e];
#pragma unroll 1
                for (int q = 0; q < 16; q++) {
                    if (!(hit & (1u << q))) continue;
                    const int nt = q >> 2, e = (q >> 1) & 1, mt = q & 1;
                    const int c = 8 * nt + 2 * t4 + e;
                    double num, det;
                    pair_gain<CSF>(rcopy[q], mt ? z1[1] : z1[0], cq[c], mt ? b1[1] : b1[0], cq[FT_TJ + c],
                                   mt ? k1[1] : k1[0], cq[2 * FT_TJ + c], mt ? g1[1] : g1[0],
                                   cq[3 * FT_TJ + c], mt ? zu1[1] : zu1[0], cq[4 * FT_TJ + c], Y3, gain_c,
                                   num, det);
                    if (!(det > 1e-12)) { gill = INFINITY; continue; }   // numerically singular
                    const double gq = num / det, tq = c0 / det;
                    if (det < kIllDet) gill = fmax(gill, gq + tq);  // ill-conditioned: optimistic gain
                    if (gq > gb) {
                        flag = (bidx >= 0 && !(gq > gb + wide)) ? 1 : 0;
                        gb = gq; tb = tq;
                        bidx = (i0 + wrow + 8 * mt + g) * N2 + jt * FT_TJ + c;
                    } else if (!(gb > gq + wide)) {
                        flag = 1;
                    }
                }
            }
            double lb = bidx >= 0 ? gb - tb : 0.0;               // certified lower bound
            for (int o = 16; o > 0; o >>= 1) lb = fmax(lb, __shfl_xor_sync(0xffffffffu, lb, o));
            if (lb > thr) {
                thr = lb;
                if (lane == 0) atomicMax(&s_thr, (unsigned long long)__double_as_longlong(lb));
            }
        }
        mbar_arrive(&s_empty[st]);
    }

    // ---- reduction over the consumer threads: best gain, tie -> lower index ----
    const double gt = bidx >= 0 ? gb : -1.0;
    const double tolt = bidx >= 0 ? tb : 0.0;
    double gm = gt;
    int im = bidx >= 0 ? bidx : INT_MAX;
    for (int o = 16; o > 0; o >>= 1) {
        double og = __shfl_xor_sync(0xffffffffu, gm, o);
        int oi = __shfl_xor_sync(0xffffffffu, im, o);
        if (og > gm || (og == gm && oi < im)) { gm = og; im = oi; }
    }
    for (int o = 16; o > 0; o >>= 1) gill = fmax(gill, __shfl_xor_sync(0xffffffffu, gill, o));
    double *redg = red, *redl = red + 16;
    int *redi = (int *)(red + 8);
    if (lane == 0) { redg[warp] = gm; redi[warp] = im; redl[warp] = gill; }
    consumer_sync();
    double G = redg[0], Gill = redl[0];
    int I = redi[0];
    for (int w = 1; w < FT_CONS / 32; w++) {
        if (redg[w] > G || (redg[w] == G && redi[w] < I)) { G = redg[w]; I = redi[w]; }
        Gill = fmax(Gill, redl[w]);
    }
    if (bidx >= 0 && bidx == I) s_tolG = tolt;
    consumer_sync();
    const double tolG = I != INT_MAX ? s_tolG : 0.0;
    if (bidx >= 0) {
        const bool winner = bidx == I;
        const bool close = gt + tolt >= G - tolG;
        if ((winner && flag) || (!winner && close)) atomicOr(&s_flag, 1);
    }
    consumer_sync();
    if (tid == 0) {
        const int64_t o = v * a.ntI + tI;
        a.cta_gain[o] = G;
        a.cta_tol[o] = tolG;
        a.cta_idx[o] = I == INT_MAX ? -1 : I;
        a.cta_flag[o] = s_flag;
        a.cta_ill[o] = Gill;
    }
}

// ---------------------------------------------------------------------------------
// General-M variant.  When the M x 128 i1 tile does not fit in shared memory (M > 112) the
// dictionaries are first normalised / CSF-projected into a zero-padded copy Dn (k_normalize),
// and k_gemm_pairs streams BOTH operands through a k-chunked shared-memory ring filled by
// TMA bulk copies (cp.async.bulk + mbarrier complete_tx, one producer warp, no FP64 work on
// the producer side); eight consumer warps accumulate a 16 x 64 correlation tile each over
// all k chunks with DMMA and then run the same closed-form screening as k_fast_pairs.
// ---------------------------------------------------------------------------------
#define GP_TI 128
#define GP_TJ 64
#define GP_KC 32
#define GP_NS 3
#define GP_S1 (GP_TI + 4)
#define GP_S2 (GP_TJ + 4)
#define GP_STAGE (GP_KC * (GP_S1 + GP_S2))
#define GP_THREADS (FT_CONS + 32)
#define GP_NROWCHUNK 64

// grid (V or 1, 2, row chunks): block k of voxel v -> Dn[v][:, off_k : off_k + Nkpad]
__global__ void __launch_bounds__(256) k_normalize(FastArgs a)
{
    const int64_t v = blockIdx.x;
    const int k = blockIdx.y;
    const int64_t row = a.vox_list ? a.vox_list[v] : v;
    const double *Ar = a.A + (a.a_by_local ? v : row) * a.strideA;
    const int Nk = k ? a.N2 : a.N1;
    const int Nkpad = k ? a.ldn - a.N1pad : a.N1pad;
    const int coff = k ? a.N1pad : 0, start = k ? a.start2 : a.start1;
    const double *cp = a.colp + (v * 2 + k) * (int64_t)FT_NPAR * a.Npad;
    double *dst = a.Dn + v * a.dn_stride + coff;
    const int M = a.p.M;
    const int m0 = blockIdx.z * GP_NROWCHUNK, m1 = min(a.Mp2, m0 + GP_NROWCHUNK);
    for (int i = threadIdx.x; i < Nkpad; i += blockDim.x) {
        const bool ok = i < Nk;
        const double sc = ok ? cp[i] : 0.0;
        const double al = (a.csf && ok) ? cp[(size_t)a.Npad + i] : 0.0;
#pragma unroll 4
        for (int m = m0; m < m1; m++) {
            double val = 0.0;
            if (ok && m < M) {
                double d = Ar[(size_t)m * a.lda + start + i];
                if (a.csf) d = fma(-al, Ar[(size_t)m * a.lda + a.start3], d);
                val = d * sc;
            }
            dst[(size_t)m * a.ldn + i] = val;
        }
    }
}

__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes)
{
    unsigned addr = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(addr), "r"(bytes) : "memory");
}
// TMA 1-D bulk copy global -> shared, completion counted in bytes on an mbarrier
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, unsigned long long *bar)
{
    unsigned d = (unsigned)__cvta_generic_to_shared(dst);
    unsigned b = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(d), "l"(__cvta_generic_to_global(src)), "r"(bytes), "r"(b) : "memory");
}

template <int CSF>
__global__ void __launch_bounds__(GP_THREADS, 1) k_gemm_pairs(FastArgs a)
{
    extern __shared__ __align__(16) double smem[];
    double *stages = smem;                                 // [GP_NS][GP_STAGE]: D1 chunk | D2 chunk
    double *colq = stages + (size_t)GP_NS * GP_STAGE;      // [8 warps][5][GP_TJ] per-warp copies
    double *red = colq + 8 * 5 * GP_TJ;                    // [64]
    __shared__ unsigned long long s_thr;
    __shared__ unsigned long long s_full[GP_NS], s_empty[GP_NS];
    __shared__ double s_tolG;
    __shared__ int s_flag;

    const int N1 = a.N1, N2 = a.N2;
    const int64_t v = blockIdx.y;
    const int tI = blockIdx.x;
    const int i0 = tI * GP_TI;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const double *vp = a.voxp + v * 8;
    const double gain_c = vp[3], c0 = vp[4], Y3 = vp[2];
    const double gpre = fmax(vp[5], vp[6]);
    const double *cp1 = a.colp + (v * 2 + 0) * (int64_t)FT_NPAR * a.Npad;
    const double *cp2 = a.colp + (v * 2 + 1) * (int64_t)FT_NPAR * a.Npad;
    const int ntJ = (N2 + GP_TJ - 1) / GP_TJ;
    const int nch = a.Mp2 / GP_KC;
    const int total = ntJ * nch;
    const double *Dv = a.Dn + v * a.dn_stride;

    if (tid == 0) {
        s_thr = (unsigned long long)__double_as_longlong(fmax(gpre - c0, 0.0));
        s_flag = 0;
        for (int st = 0; st < GP_NS; st++) { mbar_init(&s_full[st], 1); mbar_init(&s_empty[st], FT_CONS / 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (tid >= FT_CONS) {
        // ============ producer warp: one D1 row (1 KB) and one D2 row (512 B) per lane ============
        for (int s = 0; s < total; s++) {
            const int st = s % GP_NS, jt = s / nch, ch = s - jt * nch;
            if (s >= GP_NS) mbar_wait(&s_empty[st], (unsigned)((s / GP_NS) - 1) & 1u);
            double *d1 = stages + (size_t)st * GP_STAGE;
            double *d2 = d1 + GP_KC * GP_S1;
            if (lane == 0) mbar_expect_tx(&s_full[st], GP_KC * (GP_TI + GP_TJ) * (unsigned)sizeof(double));
            __syncwarp();
            const double *src = Dv + (size_t)(ch * GP_KC + lane) * a.ldn;
            bulk_g2s(d1 + (size_t)lane * GP_S1, src + i0, GP_TI * sizeof(double), &s_full[st]);
            bulk_g2s(d2 + (size_t)lane * GP_S2, src + a.N1pad + jt * GP_TJ, GP_TJ * sizeof(double), &s_full[st]);
        }
        return;
    }

    // ===================================== consumers =====================================
    const int g = lane >> 2, t4 = lane & 3;
    const double wide = 4.0 * c0 / kIllDet;
    const double negc0 = -c0;
    const int wrow = warp * 16;
    double *cq = colq + warp * 5 * GP_TJ;
    double z1[2], b1[2], k1[2], g1[2], zu1[2];
#pragma unroll
    for (int mt = 0; mt < 2; mt++) {
        const int i = i0 + wrow + 8 * mt + g;
        const bool ok = i < N1;
        z1[mt] = ok ? cp1[(size_t)2 * a.Npad + i] : 0.0;
        b1[mt] = (CSF && ok) ? cp1[(size_t)3 * a.Npad + i] : 0.0;
        k1[mt] = (CSF && ok) ? cp1[(size_t)4 * a.Npad + i] : 0.0;
        g1[mt] = (CSF && ok) ? cp1[(size_t)5 * a.Npad + i] : 0.0;
        zu1[mt] = (CSF && ok) ? cp1[(size_t)6 * a.Npad + i] : 0.0;
    }
    const int mtv = max(0, min(2, (N1 - (i0 + wrow) + 7) >> 3));
    double gb = -1.0, tb = 0.0, thr = fmax(gpre - c0, 0.0), gill = -1.0;
    int bidx = -1, flag = 0;

    int s = 0;
    for (int jt = 0; jt < ntJ; jt++) {
        // this warp's copy of the i2 tile's per-atom parameters (z, beta, kappa, gamma, zu);
        // Npad is a multiple of GP_TJ here, so the reads stay inside colp
        __syncwarp();
        for (int e = lane; e < (CSF ? 5 : 1) * GP_TJ; e += 32)
            cq[e] = __ldg(cp2 + (size_t)(e / GP_TJ + 2) * a.Npad + jt * GP_TJ + (e % GP_TJ));
        double acc[2][8][2];
#pragma unroll
        for (int mt = 0; mt < 2; mt++)
#pragma unroll
            for (int nt = 0; nt < 8; nt++) { acc[mt][nt][0] = 0.0; acc[mt][nt][1] = 0.0; }
        const int ntv = max(0, min(8, (N2 - jt * GP_TJ + 7) >> 3));
        for (int ch = 0; ch < nch; ch++, s++) {
            const int st = s % GP_NS;
            mbar_wait(&s_full[st], (unsigned)(s / GP_NS) & 1u);
            const double *A_ = stages + (size_t)st * GP_STAGE + (size_t)t4 * GP_S1 + wrow + g;
            const double *B_ = stages + (size_t)st * GP_STAGE + GP_KC * GP_S1 + (size_t)t4 * GP_S2 + g;
            if (ntv == 8 && mtv == 2) {
#pragma unroll 2
                for (int ks = 0; ks < GP_KC / 4; ks++) {
                    double af[2], bf[8];
#pragma unroll
                    for (int mt = 0; mt < 2; mt++) af[mt] = A_[(size_t)ks * 4 * GP_S1 + 8 * mt];
#pragma unroll
                    for (int nt = 0; nt < 8; nt++) bf[nt] = B_[(size_t)ks * 4 * GP_S2 + 8 * nt];
#pragma unroll
                    for (int mt = 0; mt < 2; mt++)
#pragma unroll
                        for (int nt = 0; nt < 8; nt++)
                            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                                         : "+d"(acc[mt][nt][0]), "+d"(acc[mt][nt][1])
                                         : "d"(af[mt]), "d"(bf[nt]));
                }
            } else {
#pragma unroll 1
                for (int ks = 0; ks < GP_KC / 4; ks++) {
                    double af[2], bf[8];
#pragma unroll
                    for (int mt = 0; mt < 2; mt++) af[mt] = A_[(size_t)ks * 4 * GP_S1 + 8 * mt];
#pragma unroll
                    for (int nt = 0; nt < 8; nt++) bf[nt] = B_[(size_t)ks * 4 * GP_S2 + 8 * nt];
#pragma unroll
                    for (int mt = 0; mt < 2; mt++)
#pragma unroll
                        for (int nt = 0; nt < 8; nt++)
                            if (mt < mtv && nt < ntv)
                                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                                             : "+d"(acc[mt][nt][0]), "+d"(acc[mt][nt][1])
                                             : "d"(af[mt]), "d"(bf[nt]));
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_empty[st]);
        }

        // ---- closed-form screening of the thread's 32 pairs (see k_fast_pairs) ----
        __syncwarp();
        thr = fmax(thr, __longlong_as_double((long long)s_thr));
        unsigned hit = 0;
        if (!CSF) {
#pragma unroll
            for (int nt = 0; nt < 8; nt++) {
                const double2 z2v = *reinterpret_cast<const double2 *>(cq + 8 * nt + 2 * t4);
#pragma unroll
                for (int e = 0; e < 2; e++)
#pragma unroll
                    for (int mt = 0; mt < 2; mt++) {
                        const double rho = acc[mt][nt][e], z2 = e ? z2v.y : z2v.x;
                        const double w1 = fma(-rho, z2, z1[mt]);
                        const double w2 = fma(-rho, z1[mt], z2);
                        const double det = fma(-rho, rho, 1.0);
                        const double num = fma(z1[mt], w1, z2 * w2);
                        const bool pos = min(__double2hiint(w1), __double2hiint(w2)) > 0;
                        if (pos && fma(-thr, det, num) >= negc0) hit |= 1u << (nt * 4 + e * 2 + mt);
                    }
            }
        } else {
            unsigned fb = 0;
#pragma unroll
            for (int nt = 0; nt < 8; nt++) {
                const double2 z2v = *reinterpret_cast<const double2 *>(cq + 8 * nt + 2 * t4);
                const double2 b2v = *reinterpret_cast<const double2 *>(cq + GP_TJ + 8 * nt + 2 * t4);
#pragma unroll
                for (int e = 0; e < 2; e++)
#pragma unroll
                    for (int mt = 0; mt < 2; mt++) {
                        const double rho = acc[mt][nt][e], z2 = e ? z2v.y : z2v.x, b2 = e ? b2v.y : b2v.x;
                        const double w1 = fma(-rho, z2, z1[mt]);
                        const double w2 = fma(-rho, z1[mt], z2);
                        const double det = fma(-rho, rho, 1.0);
                        const double w3 = fma(-b2, w2, fma(-b1[mt], w1, Y3 * det));
                        const double num = fma(gain_c, det, fma(z1[mt], w1, z2 * w2));
                        const bool pos = min(min(__double2hiint(w1), __double2hiint(w2)), __double2hiint(w3)) > 0;
                        const unsigned bit = 1u << (nt * 4 + e * 2 + mt);
                        if (!pos) fb |= bit;
                        else if (fma(-thr, det, num) >= negc0) hit |= bit;
                    }
            }
            if (__any_sync(0xffffffffu, fb != 0)) {
#pragma unroll
                for (int nt = 0; nt < 8; nt++) {
                    const double2 k2v = *reinterpret_cast<const double2 *>(cq + 2 * GP_TJ + 8 * nt + 2 * t4);
                    const double2 g2v = *reinterpret_cast<const double2 *>(cq + 3 * GP_TJ + 8 * nt + 2 * t4);
                    const double2 zu2v = *reinterpret_cast<const double2 *>(cq + 4 * GP_TJ + 8 * nt + 2 * t4);
#pragma unroll
                    for (int e = 0; e < 2; e++)
#pragma unroll
                        for (int mt = 0; mt < 2; mt++) {
                            const double rho = acc[mt][nt][e];
                            const double zu2 = e ? zu2v.y : zu2v.x;
                            const double r = fma(rho * k1[mt], e ? k2v.y : k2v.x, g1[mt] * (e ? g2v.y : g2v.x));
                            const double v1 = fma(-r, zu2, zu1[mt]);
                            const double v2 = fma(-r, zu1[mt], zu2);
                            const double det = fma(-r, r, 1.0);
                            const double num = fma(zu1[mt], v1, zu2 * v2);
                            const bool pos = min(__double2hiint(v1), __double2hiint(v2)) > 0;
                            const unsigned bit = 1u << (nt * 4 + e * 2 + mt);
                            if ((fb & bit) && pos && fma(-thr, det, num) >= negc0) hit |= bit;
                        }
                }
            }
        }
        if (__any_sync(0xffffffffu, hit != 0)) {
            if (hit) {
                double rcopy[32];
#pragma unroll
                for (int nt = 0; nt < 8; nt++)
#pragma unroll
                    for (int e = 0; e < 2; e++)
#pragma unroll
                        for (int mt = 0; mt < 2; mt++) rcopy[nt * 4 + e * 2 + mt] = acc[mt][nt][e];
#pragma unroll 1
                for (int q = 0; q < 32; q++) {
                    if (!(hit & (1u << q))) continue;
                    const int nt = q >> 2, e = (q >> 1) & 1, mt = q & 1;
                    const int c = 8 * nt + 2 * t4 + e;
                    double num, det;
                    pair_gain<CSF>(rcopy[q], mt ? z1[1] : z1[0], cq[c], mt ? b1[1] : b1[0],
                                   CSF ? cq[GP_TJ + c] : 0.0, mt ? k1[1] : k1[0], CSF ? cq[2 * GP_TJ + c] : 0.0,
                                   mt ? g1[1] : g1[0], CSF ? cq[3 * GP_TJ + c] : 0.0, mt ? zu1[1] : zu1[0],
                                   CSF ? cq[4 * GP_TJ + c] : 0.0, Y3, gain_c, num, det);
                    if (!(det > 1e-12)) { gill = INFINITY; continue; }
                    const double gq = num / det, tq = c0 / det;
                    if (det < kIllDet) gill = fmax(gill, gq + tq);
                    if (gq > gb) {
                        flag = (bidx >= 0 && !(gq > gb + wide)) ? 1 : 0;
                        gb = gq; tb = tq;
                        bidx = (i0 + wrow + 8 * mt + g) * N2 + jt * GP_TJ + c;
                    } else if (!(gb > gq + wide)) {
                        flag = 1;
                    }
                }
            }
            double lb = bidx >= 0 ? gb - tb : 0.0;
            for (int o = 16; o > 0; o >>= 1) lb = fmax(lb, __shfl_xor_sync(0xffffffffu, lb, o));
            if (lb > thr) {
                thr = lb;
                if (lane == 0) atomicMax(&s_thr, (unsigned long long)__double_as_longlong(lb));
            }
        }
    }

    // ---- reduction over the consumer threads ----
    const double gt = bidx >= 0 ? gb : -1.0;
    const double tolt = bidx >= 0 ? tb : 0.0;
    double gm = gt;
    int im = bidx >= 0 ? bidx : INT_MAX;
    for (int o = 16; o > 0; o >>= 1) {
        double og = __shfl_xor_sync(0xffffffffu, gm, o);
        int oi = __shfl_xor_sync(0xffffffffu, im, o);
        if (og > gm || (og == gm && oi < im)) { gm = og; im = oi; }
    }
    for (int o = 16; o > 0; o >>= 1) gill = fmax(gill, __shfl_xor_sync(0xffffffffu, gill, o));
    double *redg = red, *redl = red + 16;
    int *redi = (int *)(red + 8);
    if (lane == 0) { redg[warp] = gm; redi[warp] = im; redl[warp] = gill; }
    consumer_sync();
    double G = redg[0], Gill = redl[0];
    int I = redi[0];
    for (int w = 1; w < FT_CONS / 32; w++) {
        if (redg[w] > G || (redg[w] == G && redi[w] < I)) { G = redg[w]; I = redi[w]; }
        Gill = fmax(Gill, redl[w]);
    }
    if (bidx >= 0 && bidx == I) s_tolG = tolt;
    consumer_sync();
    const double tolG = I != INT_MAX ? s_tolG : 0.0;
    if (bidx >= 0) {
        const bool winner = bidx == I;
        const bool close = gt + tolt >= G - tolG;
        if ((winner && flag) || (!winner && close)) atomicOr(&s_flag, 1);
    }
    consumer_sync();
    if (tid == 0) {
        const int64_t o = v * a.ntI + tI;
        a.cta_gain[o] = G;
        a.cta_tol[o] = tolG;
        a.cta_idx[o] = I == INT_MAX ? -1 : I;
        a.cta_flag[o] = s_flag;
        a.cta_ill[o] = Gill;
    }
}

// ---------------------------------------------------------------------------------
// select: one thread per voxel
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_fast_select(FastArgs a, int64_t V)
{
    int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= V) return;
    const double *vp = a.voxp + v * 8;
    const double c0 = vp[4];
    const double gpre = fmax(vp[5], vp[6]);
    double G = -1.0, tolG = 0.0;
    int I = -1, best_t = -1;
    for (int t = 0; t < a.ntI; t++) {
        const int64_t o = v * a.ntI + t;
        if (a.cta_idx[o] >= 0 && (a.cta_gain[o] > G || (a.cta_gain[o] == G && a.cta_idx[o] < I))) {
            G = a.cta_gain[o]; tolG = a.cta_tol[o]; I = a.cta_idx[o]; best_t = t;
        }
    }
    bool certain = I >= 0;
    int reason = certain ? -1 : 0;
    for (int t = 0; t < a.ntI && certain; t++) {
        const int64_t o = v * a.ntI + t;
        if (a.cta_ill[o] >= G - tolG) { certain = false; reason = 1; }
        if (t == best_t) { if (a.cta_flag[o]) { certain = false; reason = 2; } continue; }
        if (a.cta_idx[o] >= 0 && a.cta_gain[o] + a.cta_tol[o] >= G - tolG) { certain = false; reason = 2; }
    }
    // pair-independent branches (single atoms, atom + CSF, CSF alone) must be clearly worse
    if (certain && !(G - tolG > gpre + 16.0 * c0)) { certain = false; reason = 3; }
    if (reason >= 0 && a.reasons) atomicAdd(&a.reasons[reason], 1);
    const int64_t row = a.vox_list ? a.vox_list[v] : v;
    if (certain) {
        a.tuple[row] = (long long)I;
    } else {
        int pos = atomicAdd(a.redo_count, 1);
        a.redo_list[pos] = (int32_t)row;
        if (a.redo_local) a.redo_local[pos] = (int32_t)v;
    }
}

// ---------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------
static size_t al256(size_t x) { return (x + 255) & ~(size_t)255; }

bool fast_supported(const DevPlan &p, int K, int csf, int ear)
{
    const int Mp = (p.M + 3) & ~3;
    return K == 2 && !ear && !p.has_between && Mp <= 112 && p.N >= 8 && p.N <= 46000 &&
           (csf == 0 || p.sig_csf);
}

bool fast_supported_materialised(const DevPlan &p, int K, int csf, int ear)
{
    return K == 2 && !ear && p.N >= 8 && p.N <= 46000 && p.M <= 16384 && (csf == 0 || p.sig_csf);
}

// Explicit dictionaries (mfb_solve_batch): two searched blocks, optionally a third block of
// exactly one column (the CSF-like compartment).
bool fast_supported_explicit(int M, const BlockSpec &bs)
{
    if (M > 16384 || bs.nb < 2 || bs.nb > 3) return false;
    if (bs.nb == 3 && bs.size[2] != 1) return false;
    if (bs.size[0] < 8 || bs.size[1] < 8) return false;
    return (long long)bs.size[0] * bs.size[1] < 2000000000LL;
}

// Geometry shared by fast_scratch_bytes and launch_fast_search.
struct FastGeom {
    int Mp, Mp2, Npad, ntI, N1pad, ldn;
    bool gemm;      // general-M path (k_normalize + k_gemm_pairs)
};
static FastGeom fast_geom(int M, int N1, int N2)
{
    FastGeom g;
    g.Mp = (M + 3) & ~3;
    g.gemm = g.Mp > 112;
    const int Nmax = N1 > N2 ? N1 : N2;
    const int padto = g.gemm ? GP_TI : FT_TJ;
    g.Npad = (Nmax + padto - 1) / padto * padto;
    g.ntI = g.gemm ? (N1 + GP_TI - 1) / GP_TI : (N1 + FT_TI - 1) / FT_TI;
    g.Mp2 = (M + GP_KC - 1) / GP_KC * GP_KC;
    g.N1pad = (N1 + GP_TI - 1) / GP_TI * GP_TI;
    g.ldn = g.N1pad + (N2 + GP_TJ - 1) / GP_TJ * GP_TJ;
    return g;
}

size_t fast_scratch_bytes(int M, int N1, int N2, int64_t V, int src, int shared_dict)
{
    const FastGeom g = fast_geom(M, N1, N2);
    size_t s = 0;
    if (!src) {
        s += al256(sizeof(int) * V * 2 * M * 2);
        s += al256(sizeof(double) * V * 2 * M * 2);
    }
    s += al256(sizeof(double) * V * 2 * FT_NPAR * g.Npad);
    s += al256(sizeof(double) * V * 8);
    s += 3 * al256(sizeof(double) * V * g.ntI);
    s += 2 * al256(sizeof(int) * V * g.ntI);
    if (g.gemm) s += al256(sizeof(double) * (shared_dict ? 1 : V) * (size_t)g.Mp2 * g.ldn);
    return s;
}

int launch_fast_search(const DevPlan &p, const FastProblem &fp, int64_t V, const int32_t *vox_list,
                       const double *peaks, int peaks_ld, const double *y, void *scratch,
                       long long *tuple, int32_t *redo_list, int32_t *redo_count, int32_t *reasons,
                       cudaStream_t st, cudaEvent_t *ev)
{
    if (V == 0) return MFB_OK;
    const FastGeom g = fast_geom(p.M, fp.N1, fp.N2);
    if (g.gemm && !fp.src) {
        set_error("fast tier: M > 112 needs explicit (materialised) dictionaries");
        return MFB_EUNSUPPORTED;
    }
    FastArgs a;
    memset(&a, 0, sizeof(a));
    a.p = p; a.csf = fp.csf;
    a.src = fp.src; a.N1 = fp.N1; a.N2 = fp.N2;
    a.A = fp.A; a.lda = fp.lda; a.strideA = fp.strideA;
    a.start1 = fp.start1; a.start2 = fp.start2; a.start3 = fp.start3;
    a.a_by_local = fp.a_by_local; a.redo_local = fp.redo_local;
    a.Mp = g.Mp; a.Npad = g.Npad; a.ntI = g.ntI;
    a.Mp2 = g.Mp2; a.N1pad = g.N1pad; a.ldn = g.ldn;
    const bool shared_dict = fp.src && fp.strideA == 0;
    a.dn_stride = shared_dict ? 0 : (int64_t)g.Mp2 * g.ldn;
    {
        const char *d = getenv("MFB_FAST_DEBUG");
        a.debug = d ? atoi(d) : 0;
    }
    a.vox_list = vox_list; a.peaks = peaks; a.peaks_ld = peaks_ld; a.y = y;
    char *q = (char *)scratch;
    if (!fp.src) {
        a.ip_rows = (int *)q; q += al256(sizeof(int) * V * 2 * p.M * 2);
        a.ip_w = (double *)q; q += al256(sizeof(double) * V * 2 * p.M * 2);
    }
    a.colp = (double *)q; q += al256(sizeof(double) * V * 2 * FT_NPAR * a.Npad);
    a.voxp = (double *)q; q += al256(sizeof(double) * V * 8);
    a.cta_gain = (double *)q; q += al256(sizeof(double) * V * a.ntI);
    a.cta_tol = (double *)q; q += al256(sizeof(double) * V * a.ntI);
    a.cta_ill = (double *)q; q += al256(sizeof(double) * V * a.ntI);
    a.cta_idx = (int *)q; q += al256(sizeof(int) * V * a.ntI);
    a.cta_flag = (int *)q; q += al256(sizeof(int) * V * a.ntI);
    a.Dn = g.gemm ? (double *)q : nullptr;
    a.tuple = tuple; a.redo_list = redo_list; a.redo_count = redo_count; a.reasons = reasons;

    const size_t smem_prep = sizeof(double) * (4 * p.M + 32) + sizeof(int) * 2 * p.M;
    if (smem_prep > 200 * 1024) {
        set_error("fast tier: too many measurements");
        return MFB_EUNSUPPORTED;
    }
    if (smem_prep > 48 * 1024)
        MFB_CUDA_TRY(cudaFuncSetAttribute(k_fast_prep, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_prep));
    MFB_LAUNCH(k_fast_prep, dim3((unsigned)V, 2), 256, smem_prep, st, a);

    const int64_t maxy = 65535;
    // per-voxel scratch is indexed by the local voxel: shift every per-voxel pointer
    auto shifted = [&](int64_t v0) {
        FastArgs b = a;
        if (!fp.src) { b.ip_rows += v0 * 2 * p.M * 2; b.ip_w += v0 * 2 * p.M * 2; }
        b.colp += v0 * 2 * FT_NPAR * a.Npad; b.voxp += v0 * 8;
        b.cta_gain += v0 * a.ntI; b.cta_tol += v0 * a.ntI; b.cta_ill += v0 * a.ntI;
        b.cta_idx += v0 * a.ntI; b.cta_flag += v0 * a.ntI;
        if (b.Dn) b.Dn += v0 * a.dn_stride;
        if (vox_list) b.vox_list = vox_list + v0;
        else { b.y = y + v0 * p.M; b.tuple = tuple + v0; }
        if (fp.A && (!vox_list || fp.a_by_local)) b.A = fp.A + v0 * fp.strideA;
        return b;
    };
    if (g.gemm) {
        const size_t smem = sizeof(double) * ((size_t)GP_NS * GP_STAGE + 8 * 5 * GP_TJ + 64);
        void (*kern)(FastArgs) = fp.csf ? k_gemm_pairs<1> : k_gemm_pairs<0>;
        MFB_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const unsigned zc = (unsigned)((g.Mp2 + GP_NROWCHUNK - 1) / GP_NROWCHUNK);
        if (shared_dict) {
            MFB_LAUNCH(k_normalize, dim3(1, 2, zc), 256, 0, st, a);
        } else {
            for (int64_t v0 = 0; v0 < V; v0 += maxy) {
                const int64_t nv = V - v0 < maxy ? V - v0 : maxy;
                MFB_LAUNCH(k_normalize, dim3((unsigned)nv, 2, zc), 256, 0, st, shifted(v0));
            }
        }
        if (ev) MFB_CUDA_TRY(cudaEventRecord(ev[0], st));
        for (int64_t v0 = 0; v0 < V; v0 += maxy) {
            const int64_t nv = V - v0 < maxy ? V - v0 : maxy;
            MFB_LAUNCH(kern, dim3(a.ntI, (unsigned)nv), GP_THREADS, smem, st, shifted(v0));
        }
        if (ev) MFB_CUDA_TRY(cudaEventRecord(ev[1], st));
    } else {
        const size_t smem = sizeof(double) * ((size_t)a.Mp * FT_S1 + (size_t)FT_NS * a.Mp * FT_S2 + FT_NS * 5 * FT_TJ +
                                              5 * a.Mp + 64) + sizeof(int) * 4 * a.Mp;
        if (smem + 64 > 227 * 1024) {
            set_error("fast tier: tile does not fit in shared memory");
            return MFB_EUNSUPPORTED;
        }
        // per device / context attribute: set on every launch (microseconds)
        void (*kern)(FastArgs) = fp.csf ? (fp.src ? k_fast_pairs<1, 1> : k_fast_pairs<1, 0>)
                                        : (fp.src ? k_fast_pairs<0, 1> : k_fast_pairs<0, 0>);
        MFB_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        if (ev) MFB_CUDA_TRY(cudaEventRecord(ev[0], st));
        for (int64_t v0 = 0; v0 < V; v0 += maxy) {
            const int64_t nv = V - v0 < maxy ? V - v0 : maxy;
            MFB_LAUNCH(kern, dim3(a.ntI, (unsigned)nv), FT_THREADS, smem, st, shifted(v0));
        }
        if (ev) MFB_CUDA_TRY(cudaEventRecord(ev[1], st));
    }
    MFB_LAUNCH(k_fast_select, (unsigned)((V + 127) / 128), 128, 0, st, a, V);
    return MFB_OK;
}

}  // namespace mfb
