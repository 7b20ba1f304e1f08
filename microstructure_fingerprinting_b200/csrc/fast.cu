// fast.cu -- the screening ("fast") tier of libmfb200, sm_100a.
//
// For a voxel with two fascicles (optionally + the CSF column) the reference
// (mf_utils.py:288-392 `_2`, 470-607 `_3`) forms the N x N cross-Gram of the two rotated
// sub-dictionaries over the M measurements and solves a closed-form 2- or 3-variable NNLS
// per atom pair; with three searched blocks it does so per atom triple.  Here:
//   k_fast_prep    rotates each sub-dictionary once per voxel to get the per-atom
//                  statistics (|a|^2, a.y, a.csf), stores the interpolation plan, and
//                  reduces the best single-atom / atom+CSF gains (the pair-independent
//                  branches of the NNLS) and the atoms holding them;
//   k_fast_seed    seeds the voxel-wide screening threshold with one real pair;
//   k_fast_tiles   M <= 111.  One CTA per (voxel, 128-atom i1 tile): the rotated, CSF-projected and
//                  normalised i1 tile is built once in shared memory; the i2 tiles of 32 atoms come
//                  from a tile-major copy prepared by k_fast_prep and are moved by ONE elected
//                  thread with TMA bulk copies (cp.async.bulk + mbarrier complete_tx) through a
//                  3-stage ring; eight consumer warps form the correlation tile with FP64
//                  tensor-core DMMA (mma.sync.m8n8k4.f64) and screen it in registers: the
//                  unconstrained-gain test is folded into the DMMA stream (row M of the tiles), the
//                  closed-form NNLS + argmax epilogue runs only for warps with a passing pair;
//   k_normalize + k_gemm_pairs   any M.  Both operands streamed through a k-chunked ring by
//                  TMA bulk copies (cp.async.bulk + mbarrier complete_tx) from a normalised
//                  copy of the dictionaries; same epilogue.  With STORE it also writes the
//                  correlation matrices for the triple scan;
//   k_triple_seed + k_triples   three searched blocks (permuted so that the largest is the
//                  streamed one): 8.5 FP64 operations per tuple from the correlation matrices in
//                  the first-level warp vote (every four steps, sign-bit tests), the last weight
//                  sign in a straight-line second level, chunk ring filled by TMA bulk copies;
//   k_fast_select / k_select3   merge the tiles of a voxel and decide whether the winner is
//                  certain; k_fast_select also restricts the reference-order search of voxels
//                  won by a one-atom solution to the rows / columns that can hold the minimum.
// Screening works on gains (|y|^2 - residual) in a different summation order than the
// reference, so it only *selects*: the winning tuple is re-evaluated in the reference's
// arithmetic by the exact tier's evaluate kernel, and every voxel whose winner is not
// separated from the runner-up (or from a solution with fewer active columns) by more than
// the screening error bounds is handed to the exact tier.
#include <climits>
#include <cstdlib>
#include <cstring>

#include "common.cuh"

namespace mfb {

#define FT_TI 128
#define FT_TJ 32
#define FT_CONS 256         // consumer threads (8 DMMA warps)
#define FT_NS 3             // stages of the i2-tile ring
#define FT_S1 (FT_TI + 4)  // row strides: == 4 mod 16 doubles -> conflict-free fragment loads
#define FT_S2 (FT_TJ + 4)
#define FT_NQ 6            // rows of a colq slot: z, beta, kappa, gamma, zu, alpha2 (folded screen)
#define FT_NPAR 8          // per-atom parameters: scale, alpha, z, beta, kappa, gamma, zu, single-solution gain
#define FT_VP 16           // per-voxel scalars (voxp): see FastArgs

// A competitive pair / tuple whose (refined) error bound exceeds kIllTol * c0 is tracked as
// ill-conditioned: it can win only through the exact tier.
static constexpr double kIllTol = 64.0;
// c0 = kC0 (M + 8) eps |y|^2: bound on the error of the screening quantity
// T = z1^2 + z2^2 - 2 rho z1 z2 - thr (1 - rho^2) caused by the rounding of its inputs (rho, z from
// normalised FMA / DMMA dot products of M terms) and of its own evaluation.
static constexpr double kC0 = 8.0;
// Experiment switches (timing ablations; some produce invalid results) exist only in builds
// with -DMFB_EXPERIMENTS; the default library has no environment-dependent behaviour.
#ifdef MFB_EXPERIMENTS
#define FT_DEBUG(a, bit) (((a).debug & (bit)) != 0)
#else
#define FT_DEBUG(a, bit) false
#endif
// The screening threshold starts kPreMargin * c0 below the best solution with fewer active
// columns (single atom, atom + CSF): a voxel without any competitive pair is then certified to
// be won by such a solution with that margin, and its reference-order search can be
// restricted to the rows / columns of the atoms that can hold it (k_fast_select).
static constexpr double kPreMargin = 32.0;
// The reference's three-block solver accepts the unconstrained Cramer solution when its
// numerators D_i >= -tol with an ABSOLUTE tol = 2.2204e-14 (mf_utils.py:562): on data of tiny
// magnitude (|a|^4 |a.y| approaching tol) it takes solutions with negative weights, which the
// screening tier (strictly positive weights) would never certify.  Voxels whose numerators can
// be that small are handed to the reference-order tier, which reproduces the tolerance.
static constexpr double kCramerScaleMin = 2.2204e-14 * 1e6;
// The screen is folded into the DMMA stream only while the (CSF-reduced) threshold is at least
// 1 / kFoldMax of the (projected) signal energy: the folded quantities grow like that ratio.
static constexpr double kFoldMax = 8.0;
// evaluation error of the folded form, in units of (thr' + |y'|^2)^2 / thr' 
static constexpr double kFoldEps = 16.0 * 2.2204e-16;

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
    uint8_t *redo_mask;   // [redo position][2][mask_ld]: atoms whose rows / columns the exact tier must scan
    int mask_ld;
    int nblk;          // searched blocks (2, or 3 for the triple scan)
    int Nb[3], startb[3], dnoff[3];   // atoms, first column in A, first column in Dn of each block
    int nsplit;        // k_gemm_pairs: CTAs sharing one i1 tile, each scanning a slice of the i2 tiles
    int njobs;         // k_gemm_pairs jobs per voxel (1: the pair scan; 3: the correlation matrices of a triple scan)
    int job_rb[3], job_cb[3];         // row / column block of each job
    double *R[3];      // job outputs (STORE): [v][rows padded][ldr], normalised correlations
    int ldr[3];
    int64_t r_stride[3];
    // triple scan (k_triples): thread layout, tiles, per-tile results, voxel-wide threshold
    int tr_txt, tr_tyt, tr_nt1, tr_ntiles, tr_kc;
    int tr_perm[3];              // scan block k = caller's block tr_perm[k] (the scan streams its largest block)
    double *t_gain, *t_tol, *t_ill;
    long long *t_idx;
    int *t_flag;
    unsigned long long *vthr;
    int csf;
    int Mp;            // M + 1 (the folded-screen row) padded to a multiple of 4
    int Npad;          // max(N1, N2) padded to a multiple of FT_TJ
    int ntI;           // i1 tiles per voxel
    int debug;            // experiments only (MFB_FAST_DEBUG): 1 skip epilogue, 2 skip gathers (timing, results
                          // invalid); 4 keep the Cramer-form error bound (no refinement)
    const int32_t *vox_list;
    const double *peaks;
    int peaks_ld;
    const double *y;
    int *ip_rows;      // [v][2][M][2]  rl, rh
    double *ip_w;      // [v][2][M][2]  wl, wh
    // tile-major prepared copy of the STREAMED block for k_fast_tiles (written by k_fast_prep):
    // per voxel, the rotated / CSF-projected / normalised atoms of block 2 laid out exactly as
    // the kernel's shared-memory tiles, so that one TMA bulk copy moves a whole tile; the 8 CTAs
    // of a voxel read it through the L2.  (The resident i1 tile is used by ONE CTA: it is built
    // in place by that CTA, a prepared copy would only add HBM traffic.)
    double *D2c;       // [v][i2 tile][Mp][FT_S2] + [FT_NQ][FT_TJ] per-atom parameters
    int64_t d2c_stride;   // doubles per voxel
    int nt2;           // i2 tiles per voxel
    double *colp;      // [v][2][FT_NPAR][Npad]
    double *voxp;      // [v][FT_VP]  (slots 12..14: magnitude of the Cramer numerators of block 0..2, see kCramerScaleMin)
                       //             y_sq, A33, Y3, gain_c3 (CSF projection), c0, Gpre(block 0..2), best single atom
                       //             of block 0..2, gain of the CSF-only solution
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
    const int Nk = a.Nb[k];
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
    // gain_c3: projection of y on the CSF column, the CSF share of the gain of every solution
    // that contains the column with its UNCONSTRAINED weight (valid whatever the sign of Y3: the
    // three-compartment weight w3 can be positive although csf.y is not); gain_c: the CSF-only
    // NNLS solution, which needs Y3 > 0
    const double gain_c3 = a.csf ? Y3 * Y3 / A33 : 0.0;
    const double gain_c = (a.csf && Y3 > 0) ? gain_c3 : 0.0;
    double *cp = a.colp + (v * a.nblk + k) * (int64_t)FT_NPAR * a.Npad;
    double gbest = 0.0, sqmax = 0.0, dymax = 0.0;
    int ibest = 0;
    // atoms up to the end of the block's last tile (the prepared copies hold whole tiles)
    const int iend = (a.D2c && k == 1) ? a.nt2 * FT_TJ : a.Npad;
    for (int i = threadIdx.x; i < iend; i += blockDim.x) {
        double par[FT_NPAR] = {0, 0, 0, 0, 0, 0, 0, 0};
        if (i < Nk) {
            double sq = 0.0, dy = 0.0, d3 = 0.0;
            const double *Ac = a.src ? Ar + a.startb[k] + i : nullptr;
            for (int m = 0; m < M; m++) {
                double d = a.src ? Ac[(size_t)m * a.lda]
                                 : fma(wh[m], p.table[(size_t)rh[m] * p.N + i], wl[m] * p.table[(size_t)rl[m] * p.N + i]);
                sq = fma(d, d, sq);
                dy = fma(d, ys[m], dy);
                d3 = fma(d, cs[m], d3);
            }
            sqmax = fmax(sqmax, sq);
            dymax = fmax(dymax, fabs(dy));
            const double r = rsqrt(sq);
            if (!a.csf) {
                par[0] = r;            // scale
                par[2] = dy * r;       // z
                const double gi = dy > 0 ? dy * dy / sq : 0.0;
                par[7] = gi;
                if (gi > gbest) { gbest = gi; ibest = i; }
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
                const double gi = nnls2_gain(sq, d3, A33, dy, Y3);
                par[7] = gi;
                if (gi > gbest) { gbest = gi; ibest = i; }
            }
        }
        if (i < a.Npad)
            for (int q = 0; q < FT_NPAR; q++) cp[(size_t)q * a.Npad + i] = par[q];
        if (a.D2c && k == 1) {
            // second pass over the measurements: the atom as the pair kernel multiplies it
            // (rotated, off the CSF column, unit norm), written into its tile; rows M.. of a
            // tile: the folded-screen row (-z) and zero padding; then the tile's parameters
            const size_t rec = (size_t)a.Mp * FT_S2 + FT_NQ * FT_TJ;
            double *dc = a.D2c + v * a.d2c_stride + (size_t)(i / FT_TJ) * rec + (i % FT_TJ);
            const double *Ac = a.src ? Ar + a.startb[k] + i : nullptr;
            const double scl = par[0], al = par[1];
            for (int m = 0; m < M; m++) {
                double val = 0.0;
                if (i < Nk) {
                    double d = a.src ? Ac[(size_t)m * a.lda]
                                     : fma(wh[m], p.table[(size_t)rh[m] * p.N + i], wl[m] * p.table[(size_t)rl[m] * p.N + i]);
                    if (a.csf) d = fma(-al, cs[m], d);
                    val = d * scl;
                }
                dc[(size_t)m * FT_S2] = val;
            }
            for (int m = M; m < a.Mp; m++) dc[(size_t)m * FT_S2] = m == M ? -par[2] : 0.0;
            double *pq = dc + (size_t)a.Mp * FT_S2;      // [FT_NQ][FT_TJ]: z, beta, kappa, gamma, zu, z^2
            for (int q = 0; q < 5; q++) pq[q * FT_TJ] = par[2 + q];
            pq[5 * FT_TJ] = par[2] * par[2];
        }
    }
    const double gmine = gbest;
    gbest = block_max(gbest, red);
    sqmax = block_max(sqmax, red);
    dymax = block_max(dymax, red);
    // atom holding the best single-column gain (lowest index on ties): the pair scan starts there
    __shared__ int s_ibest;
    if (threadIdx.x == 0) s_ibest = INT_MAX;
    __syncthreads();
    if (gmine == gbest) atomicMin(&s_ibest, ibest);
    __syncthreads();
    if (threadIdx.x == 0) {
        double *vp = a.voxp + v * FT_VP;
        vp[5 + k] = fmax(gbest, gain_c);   // slots 5, 6, 7: blocks 0, 1, 2
        vp[8 + k] = (double)s_ibest;
        vp[12 + k] = fmax(sqmax, a.csf ? A33 : 0.0) * fmax(sqmax, a.csf ? A33 : 0.0) * fmax(dymax, a.csf ? fabs(Y3) : 0.0);
        if (k == 0) {
            vp[0] = y_sq; vp[1] = A33; vp[2] = Y3; vp[3] = gain_c3; vp[11] = gain_c;
            vp[4] = kC0 * (M + 8) * 2.2204e-16 * y_sq;   // c0: screening error scale
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
                                          double &det, double &re, double &za, double &zb, double &gadd)
{
    const double w1 = fma(-rho, z2, z1);
    const double w2 = fma(-rho, z1, z2);
    det = fma(-rho, rho, 1.0);
    num = fma(z1, w1, z2 * w2);
    re = rho; za = z1; zb = z2; gadd = 0.0;     // the 2 x 2 system the gain belongs to
    bool pos = min(__double2hiint(w1), __double2hiint(w2)) > 0;
    if (CSF) {
        const double w3 = fma(-b2, w2, fma(-b1, w1, Y3 * det));
        pos = pos && (__double2hiint(w3) > 0);
        num = fma(gain_c, det, num);
        gadd = gain_c;
        if (!pos) {
            // best of the 2-column sub-problems: only the fascicle pair depends on (i1, i2);
            // the atom + CSF ones are pair-independent and live in gpre
            const double r = fma(rho * k1, k2, g1 * g2);
            const double v1 = fma(-r, zu2, zu1);
            const double v2 = fma(-r, zu1, zu2);
            pos = min(__double2hiint(v1), __double2hiint(v2)) > 0;
            det = fma(-r, r, 1.0);
            num = fma(zu1, v1, zu2 * v2);
            re = r; za = zu1; zb = zu2; gadd = 0.0;
        }
    }
    return pos;
}

// Rare path: sharpen (gain, error bound) of a competitive pair.  The Cramer form num / det
// loses c0 / det to cancellation; the stationary form 2 w.z - w'Gw evaluated at the Cramer
// weights is second order in their error, so its bound only carries the input errors of the
// correlations and projections (relative c1 = c0 / |y|^2) times (|w|_1^2 + |w|_1 |y|).
__device__ __forceinline__ void refine_pair(double re, double za, double zb, double gadd, double rdet,
                                            double c0, double c1, double y_sq, double &gq, double &tq)
{
    const double w1 = fma(-re, zb, za) * rdet, w2 = fma(-re, za, zb) * rdet;
    const double gr = 2.0 * fma(w1, za, w2 * zb) - fma(w1, w1, fma(w2, w2, 2.0 * re * w1 * w2)) + gadd;
    const double sw = fabs(w1) + fabs(w2), rel = c1 * rdet;
    // |w|_1 |y| <= (|w|_1^2 + |y|^2) / 2
    const double tr = fma(1.5 * c1, sw * sw, fma(4.0 * rel * rel, y_sq, (gadd != 0.0 ? 1.5 : 0.5) * c0));
    if (tr < tq) { gq = gr; tq = tr; }
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
__device__ __forceinline__ void consumer_sync()
{
    asm volatile("bar.sync 1, %0;" ::"n"(FT_CONS) : "memory");
}

// Seed of the voxel-wide screening threshold: the certified lower bound of the gain of ONE
// real pair, (best single atom of block 1, best single atom of block 2).  Without it every
// warp's first tile floods the competitive-pair path (any pair beats the single-atom bound).
// grid V, one warp.
__global__ void __launch_bounds__(32) k_fast_seed(FastArgs a)
{
    const DevPlan &p = a.p;
    const int M = p.M, lane = threadIdx.x;
    const int64_t v = blockIdx.x;
    const double *vp = a.voxp + v * FT_VP;
    const int i1 = max(0, min(a.N1 - 1, (int)vp[8])), i2 = max(0, min(a.N2 - 1, (int)vp[9]));
    const double *cp1 = a.colp + (v * 2 + 0) * (int64_t)FT_NPAR * a.Npad;
    const double *cp2 = a.colp + (v * 2 + 1) * (int64_t)FT_NPAR * a.Npad;
    const int64_t row = a.vox_list ? a.vox_list[v] : v;
    const double *Ar = a.src ? a.A + (a.a_by_local ? v : row) * a.strideA : nullptr;
    const double sc1 = cp1[i1], sc2 = cp2[i2];
    const double al1 = a.csf ? cp1[(size_t)a.Npad + i1] : 0.0, al2 = a.csf ? cp2[(size_t)a.Npad + i2] : 0.0;
    double acc = 0.0;
    for (int m = lane; m < M; m += 32) {
        double d1, d2;
        if (a.src) {
            d1 = Ar[(size_t)m * a.lda + a.start1 + i1];
            d2 = Ar[(size_t)m * a.lda + a.start2 + i2];
        } else {
            const int64_t o1 = ((v * 2 + 0) * M + m) * 2, o2 = ((v * 2 + 1) * M + m) * 2;
            d1 = fma(a.ip_w[o1 + 1], p.table[(size_t)a.ip_rows[o1 + 1] * p.N + i1],
                     a.ip_w[o1] * p.table[(size_t)a.ip_rows[o1] * p.N + i1]);
            d2 = fma(a.ip_w[o2 + 1], p.table[(size_t)a.ip_rows[o2 + 1] * p.N + i2],
                     a.ip_w[o2] * p.table[(size_t)a.ip_rows[o2] * p.N + i2]);
        }
        if (a.csf) {
            const double c = a.src ? Ar[(size_t)m * a.lda + a.start3] : p.sig_csf[m];
            d1 = fma(-al1, c, d1);
            d2 = fma(-al2, c, d2);
        }
        acc = fma(d1 * sc1, d2 * sc2, acc);
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) {
        const double c0 = vp[4];
        double num, det, re, za, zb, gadd;
        bool pos;
        if (a.csf)
            pos = pair_gain<1>(acc, cp1[(size_t)2 * a.Npad + i1], cp2[(size_t)2 * a.Npad + i2],
                               cp1[(size_t)3 * a.Npad + i1], cp2[(size_t)3 * a.Npad + i2],
                               cp1[(size_t)4 * a.Npad + i1], cp2[(size_t)4 * a.Npad + i2],
                               cp1[(size_t)5 * a.Npad + i1], cp2[(size_t)5 * a.Npad + i2],
                               cp1[(size_t)6 * a.Npad + i1], cp2[(size_t)6 * a.Npad + i2], vp[2], vp[3],
                               num, det, re, za, zb, gadd);
        else
            pos = pair_gain<0>(acc, cp1[(size_t)2 * a.Npad + i1], cp2[(size_t)2 * a.Npad + i2], 0, 0, 0, 0, 0, 0,
                               0, 0, 0, 0, num, det, re, za, zb, gadd);
        double lb = 0.0;
        if (pos && det > 1e-6) lb = fmax((num - c0) / det - 2.0 * c0, 0.0);
        a.vthr[v] = (unsigned long long)__double_as_longlong(lb);
    }
}

// ---------------------------------------------------------------------------------
// k_fast_tiles: the pair scan for M <= 111.  One CTA per (voxel, 128-atom i1 tile); the i1 tile
// (rotated from the L2-resident lookup table or read from explicit dictionaries, projected off
// the CSF column, unit-normalised) is built once in shared memory by the CTA that uses it; the
// 32-atom i2 tiles are streamed through a 3-stage ring; eight consumer warps form a 16 x 32
// correlation tile each with FP64 tensor-core DMMA (m8n8k4, 27 k-steps at M = 105).
//   * Nothing is gathered or computed on the producer side: ONE elected thread moves every i2
//     tile and its per-atom parameters with one TMA bulk copy (cp.async.bulk + mbarrier
//     complete_tx) from the tile-major copy k_fast_prep wrote, which the 8 CTAs of a voxel
//     share through the L2.  (The round-1 kernel gathered and blended the i2 tiles with four
//     producer warps in every CTA: a select added to their store pass cost 12 % of the kernel,
//     DESIGN.md section 8 -- the producers, not the DMMA stream, set its pace.)
//   * The screen is folded into the DMMA stream with a WARP-PRIVATE threshold: row M of the
//     streamed tiles holds -z2_j, and every consumer warp keeps z1_i / thr' in ITS 16 columns
//     of row M of the resident tile (no other warp reads those columns), rewriting them
//     whenever it adopts a higher threshold.  The accumulators then deliver
//     rho~ = rho - z1 z2 / thr' and, with alpha = 1 - z^2 / thr',
//         z1^2 + z2^2 - 2 rho z1 z2 - thr' (1 - rho^2)  =  thr' (rho~^2 - alpha1 alpha2),
//     the unconstrained least-squares gain test of the pair's columns (+ CSF), which is
//     NECESSARY for any solution on a subset of those columns to reach the threshold: about
//     2.5 FP64 operations and a sign test per pair (the round-1 screen: 7 to 11).  Only warps in
//     which some pair passes run the closed-form NNLS screen (signs of the weights, 2-column
//     sub-problems, current threshold) on the recovered rho, and from there the
//     competitive-pair path (refined error bounds, certified lower bound -> threshold).
// ---------------------------------------------------------------------------------
#define FT2_THREADS (FT_CONS + 32)
#define FT2_REC(Mp) ((size_t)(Mp) * FT_S2 + FT_NQ * FT_TJ)      // doubles per streamed tile record

template <int CSF, int SRC>
__global__ void __launch_bounds__(FT2_THREADS, 1) k_fast_tiles(FastArgs a)
{
    extern __shared__ __align__(16) double smem[];
    const DevPlan &p = a.p;
    const int M = p.M, N = p.N, Mp = a.Mp;
    const int N1 = a.N1, N2 = a.N2;
    const size_t rec = FT2_REC(Mp);
    double *D1s = smem;                                   // [Mp][FT_S1]
    double *D2s = D1s + (size_t)Mp * FT_S1;               // [FT_NS] records: [Mp][FT_S2] | [FT_NQ][FT_TJ]
    double *w1l = D2s + (size_t)FT_NS * rec;              // [Mp] plan of fascicle 1
    double *w1h = w1l + Mp;
    double *cs = w1h + Mp;                                // [Mp] csf column
    double *red = cs + Mp;                                // [64]
    int *r1l = (int *)(red + 64);
    int *r1h = r1l + Mp;
    __shared__ unsigned long long s_thr;                  // CTA-wide lower bound on the winning gain
    __shared__ unsigned long long s_full[FT_NS], s_empty[FT_NS];
    __shared__ double s_tolG;
    __shared__ int s_flag;

    const int64_t v = blockIdx.y;
    const int tI = blockIdx.x;
    const int i0 = tI * FT_TI;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const double *vp = a.voxp + v * FT_VP;
    const double gain_c = vp[3], c0 = vp[4], Y3 = vp[2];
    const double ysq_p = fmax(vp[0] - gain_c, 0.0);       // energy of y off the CSF column
    const double gpre = fmax(vp[5], vp[6]);
    const double *cp1 = a.colp + (v * 2 + 0) * (int64_t)FT_NPAR * a.Npad;
    const int ntJ = a.nt2;
    const int jt0 = min(ntJ - 1, max(0, (int)vp[9]) / FT_TJ);   // scan starts at block 2's best single atom
    unsigned long long *vthr = a.vthr + v;
    const int64_t row = a.vox_list ? a.vox_list[v] : v;
    const double *Ar = SRC ? a.A + (a.a_by_local ? v : row) * a.strideA : nullptr;

    for (int m = tid; m < Mp; m += FT2_THREADS) {
        if (m < M && SRC) {
            // explicit dictionaries: rows are read directly (weights 1 / 0, row index = m)
            r1l[m] = r1h[m] = m;
            w1l[m] = 0.0; w1h[m] = 1.0;
            cs[m] = CSF ? Ar[(size_t)m * a.lda + a.start3] : 0.0;
        } else if (m < M) {
            const int64_t o = ((v * 2 + 0) * M + m) * 2;
            r1l[m] = a.ip_rows[o]; r1h[m] = a.ip_rows[o + 1];
            w1l[m] = a.ip_w[o]; w1h[m] = a.ip_w[o + 1];
            cs[m] = CSF ? p.sig_csf[m] : 0.0;
        } else {
            r1l[m] = r1h[m] = 0;
            w1l[m] = w1h[m] = 0.0; cs[m] = 0.0;
        }
    }
    if (tid == 0) {
        s_thr = max((unsigned long long)__double_as_longlong(fmax(gpre - kPreMargin * c0, 0.0)),
                    *(volatile unsigned long long *)vthr);
        s_flag = 0;
        for (int st = 0; st < FT_NS; st++) { mbar_init(&s_full[st], 1); mbar_init(&s_empty[st], FT_CONS / 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (tid >= FT_CONS) {
        // ============ producer: one thread, one bulk copy per tile ============
        if (lane == 0) {
            const double *src2 = a.D2c + v * a.d2c_stride;
            for (int jt = 0; jt < ntJ; jt++) {
                const int st = jt % FT_NS;
                if (jt >= FT_NS) mbar_wait(&s_empty[st], (unsigned)((jt / FT_NS) - 1) & 1u);
                const int jr = jt + jt0 < ntJ ? jt + jt0 : jt + jt0 - ntJ;
                atomicMax(&s_thr, *(volatile unsigned long long *)vthr);   // thresholds of the voxel's other CTAs
                mbar_expect_tx(&s_full[st], (unsigned)(rec * sizeof(double)));
                bulk_g2s(D2s + (size_t)st * rec, src2 + (size_t)jr * rec, (unsigned)(rec * sizeof(double)), &s_full[st]);
            }
        }
        return;
    }

    // =============================== consumers ===============================
    const int g = lane >> 2, t4 = lane & 3;
    const double wide = 4.0 * kIllTol * c0;
    const double c1 = kC0 * (a.p.M + 8) * 2.2204e-16;     // c0 / |y|^2 (k_fast_prep)
    const double negc0 = -c0;
    // ---- resident i1 tile: rotate, project out the CSF column, normalise (rows >= M: zero; the
    // warps fill their entries of row M when they adopt a threshold) ----
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
    // z of the row whose fold entry this lane maintains (lanes 0..15: rows wrow + lane)
    const double z1own = (lane < 16 && i0 + wrow + lane < N1) ? cp1[(size_t)2 * a.Npad + i0 + wrow + lane] : 0.0;
    const int mtv = max(0, min(2, (N1 - (i0 + wrow) + 7) >> 3));   // valid 8-row blocks of this warp

    double gb = -1.0, tb = 0.0, thr = fmax(gpre - kPreMargin * c0, 0.0), gill = -1.0;
    int bidx = -1, flag = 0;
    // the warp's folded threshold: thr_f = thr' of the fold entries in shared memory (0: unfolded)
    double thr_w = -1.0, thr_f = 0.0, rthr = 0.0, cm = 0.0, al1[2] = {1.0, 1.0};

    for (int jt = 0; jt < ntJ; jt++) {
        const int st = jt % FT_NS;
        const int jr = jt + jt0 < ntJ ? jt + jt0 : jt + jt0 - ntJ;
        thr = fmax(thr, __longlong_as_double((long long)*(volatile unsigned long long *)&s_thr));
        if (thr != thr_w) {
            // adopt the threshold: this warp's 16 entries of row M of the resident tile
            thr_w = thr;
            const double thrp = thr - gain_c;
            const bool fold = thrp * kFoldMax > ysq_p && thrp > 0.0;
            rthr = fold ? 1.0 / thrp : 0.0;
            thr_f = fold ? thrp : 0.0;
            if (lane < 16) D1s[(size_t)M * FT_S1 + wrow + lane] = z1own * rthr;
            const double sum = thrp + ysq_p;
            cm = fma(kFoldEps * sum, sum * rthr, c0) * rthr;               // margin / thr'
#pragma unroll
            for (int mt = 0; mt < 2; mt++) al1[mt] = fma(-z1[mt] * z1[mt], rthr, 1.0);
            __syncwarp();
        }
        mbar_wait(&s_full[st], (unsigned)(jt / FT_NS) & 1u);

        // ---- correlation tile: 16 x 32 per warp, DMMA m8n8k4 over k ----
        double acc[2][4][2];
#pragma unroll
        for (int mt = 0; mt < 2; mt++)
#pragma unroll
            for (int nt = 0; nt < 4; nt++) { acc[mt][nt][0] = 0.0; acc[mt][nt][1] = 0.0; }
        const double *A_ = D1s + (size_t)t4 * FT_S1 + wrow + g;
        const double *B_ = D2s + (size_t)st * rec + (size_t)t4 * FT_S2 + g;
        const int ntv = min(4, (N2 - jr * FT_TJ + 7) >> 3);
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

        const double *cq = D2s + (size_t)st * rec + (size_t)Mp * FT_S2;     // the tile's per-atom parameters
        unsigned hit = 0, hit1 = ~0u;
        if (thr_f > 0.0) {
            // ---- level 1: folded screen, thr' (rho~^2 - alpha1 alpha2) >= -margin ----
            hit1 = 0;
#pragma unroll
            for (int nt = 0; nt < 4; nt++) {
                const double2 zq = *reinterpret_cast<const double2 *>(cq + 5 * FT_TJ + 8 * nt + 2 * t4);
                const double a2[2] = {fma(-zq.x, rthr, 1.0), fma(-zq.y, rthr, 1.0)};
#pragma unroll
                for (int e = 0; e < 2; e++)
#pragma unroll
                    for (int mt = 0; mt < 2; mt++) {
                        const double rt = acc[mt][nt][e];
                        const double t = fma(-al1[mt], a2[e], fma(rt, rt, cm));
                        if (__double2hiint(t) >= 0) hit1 |= 1u << (nt * 4 + e * 2 + mt);
                    }
            }
        }
        if (__any_sync(0xffffffffu, hit1 != 0)) {
            // ---- level 2: closed-form NNLS screen on rho recovered with the very products the
            // fold rows hold, current threshold, straight-line over the thread's 16 pairs ----
            if (thr_f > 0.0) {
                const double bz[2] = {z1[0] * rthr, z1[1] * rthr};
#pragma unroll
                for (int nt = 0; nt < 4; nt++) {
                    const double2 z2v = *reinterpret_cast<const double2 *>(cq + 8 * nt + 2 * t4);
#pragma unroll
                    for (int e = 0; e < 2; e++)
#pragma unroll
                        for (int mt = 0; mt < 2; mt++)
                            acc[mt][nt][e] = fma(bz[mt], e ? z2v.y : z2v.x, acc[mt][nt][e]);
                }
            }
            if (!CSF) {
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
                        double num, det, re, za, zb, gadd;
                        if (!pair_gain<CSF>(rcopy[q], mt ? z1[1] : z1[0], cq[c], mt ? b1[1] : b1[0], cq[FT_TJ + c],
                                            mt ? k1[1] : k1[0], cq[2 * FT_TJ + c], mt ? g1[1] : g1[0],
                                            cq[3 * FT_TJ + c], mt ? zu1[1] : zu1[0], cq[4 * FT_TJ + c], Y3, gain_c,
                                            num, det, re, za, zb, gadd))
                            continue;
                        if (!(det > 1e-12)) { gill = INFINITY; continue; }   // numerically singular
                        const double rdet = 1.0 / det;
                        double gq = num * rdet, tq = c0 * rdet;
                        if (!(gq + tq >= thr)) continue;
                        refine_pair(re, za, zb, gadd, rdet, c0, c1, vp[0], gq, tq);
                        if (tq > kIllTol * c0) gill = fmax(gill, gq + tq);  // ill-conditioned: optimistic gain
                        if (gq > gb) {
                            flag = (bidx >= 0 && !(gq > gb + wide)) ? 1 : 0;
                            gb = gq; tb = tq;
                            bidx = (i0 + wrow + 8 * mt + g) * N2 + jr * FT_TJ + c;
                        } else if (!(gb > gq + wide)) {
                            flag = 1;
                        }
                    }
                }
                double lb = bidx >= 0 ? gb - tb : 0.0;               // certified lower bound
                for (int o = 16; o > 0; o >>= 1) lb = fmax(lb, __shfl_xor_sync(0xffffffffu, lb, o));
                if (lb > thr) {
                    thr = lb;
                    if (lane == 0) {
                        atomicMax(&s_thr, (unsigned long long)__double_as_longlong(lb));
                        atomicMax(vthr, (unsigned long long)__double_as_longlong(lb));
                    }
                }
            }
        }
        // the record (tile + parameters) is consumed
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_empty[st]);
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
// all k chunks with DMMA and then run the closed-form screening (level 2 of k_fast_tiles).
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
    const int Nk = a.Nb[k];
    const int coff = a.dnoff[k], start = a.startb[k];
    const int Nkpad = (k + 1 < a.nblk ? a.dnoff[k + 1] : a.ldn) - coff;
    const double *cp = a.colp + (v * a.nblk + k) * (int64_t)FT_NPAR * a.Npad;
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


// STORE: also write the normalised correlation tile to R[job] (the triple scan reads it).
template <int CSF, int STORE>
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

    const int job = blockIdx.z;
    const int rb = a.job_rb[job], cb = a.job_cb[job];
    const int N1 = a.Nb[rb], N2 = a.Nb[cb];
    const int64_t v = blockIdx.y;
    // blockIdx.x = i1 tile + (tiles) * slice: the CTAs of a voxel are adjacent in launch order,
    // so the voxels in flight (and their Dn copies) stay few enough for the L2
    const int ntI1 = a.ntI / a.nsplit;
    const int tI = blockIdx.x % ntI1, sp = blockIdx.x / ntI1;
    const int i0 = tI * GP_TI;
    if (i0 >= N1) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const double *vp = a.voxp + v * FT_VP;
    const double gain_c = vp[3], c0 = vp[4], Y3 = vp[2];
    const double gpre = fmax(vp[5 + rb], vp[5 + cb]);
    const double *cp1 = a.colp + (v * a.nblk + rb) * (int64_t)FT_NPAR * a.Npad;
    const double *cp2 = a.colp + (v * a.nblk + cb) * (int64_t)FT_NPAR * a.Npad;
    const int doff1 = a.dnoff[rb] + i0, doff2 = a.dnoff[cb];
    // STORE covers the whole zero-padded column block so that R has no unwritten entries
    const int ntJ = STORE ? ((cb + 1 < a.nblk ? a.dnoff[cb + 1] : a.ldn) - a.dnoff[cb]) / GP_TJ
                          : (N2 + GP_TJ - 1) / GP_TJ;
    const int nch = a.Mp2 / GP_KC;
    const int jt_lo = (int)((long long)sp * ntJ / a.nsplit), jt_hi = (int)((long long)(sp + 1) * ntJ / a.nsplit);
    const int total = (jt_hi - jt_lo) * nch;
    const double *Dv = a.Dn + v * a.dn_stride;

    if (tid == 0) {
        s_thr = (unsigned long long)__double_as_longlong(fmax(gpre - kPreMargin * c0, 0.0));
        s_flag = 0;
        for (int st = 0; st < GP_NS; st++) { mbar_init(&s_full[st], 1); mbar_init(&s_empty[st], FT_CONS / 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (tid >= FT_CONS) {
        // ============ producer warp: one D1 row (1 KB) and one D2 row (512 B) per lane ============
        for (int s = 0; s < total; s++) {
            const int st = s % GP_NS, jt = jt_lo + s / nch, ch = s % nch;
            if (s >= GP_NS) mbar_wait(&s_empty[st], (unsigned)((s / GP_NS) - 1) & 1u);
            double *d1 = stages + (size_t)st * GP_STAGE;
            double *d2 = d1 + GP_KC * GP_S1;
            if (lane == 0) mbar_expect_tx(&s_full[st], GP_KC * (GP_TI + GP_TJ) * (unsigned)sizeof(double));
            __syncwarp();
            const double *src = Dv + (size_t)(ch * GP_KC + lane) * a.ldn;
            bulk_g2s(d1 + (size_t)lane * GP_S1, src + doff1, GP_TI * sizeof(double), &s_full[st]);
            bulk_g2s(d2 + (size_t)lane * GP_S2, src + doff2 + jt * GP_TJ, GP_TJ * sizeof(double), &s_full[st]);
        }
        return;
    }

    // ===================================== consumers =====================================
    const int g = lane >> 2, t4 = lane & 3;
    const double wide = 4.0 * kIllTol * c0;
    const double c1 = kC0 * (a.p.M + 8) * 2.2204e-16;     // c0 / |y|^2 (k_fast_prep)
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
    double gb = -1.0, tb = 0.0, thr = fmax(gpre - kPreMargin * c0, 0.0), gill = -1.0;
    int bidx = -1, flag = 0;

    int s = 0;
    for (int jt = jt_lo; jt < jt_hi; jt++) {
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
            // the rows of the last chunk beyond M are zero padding: their k-steps are skipped
            // (M = 105: 3 instead of 8 k-steps in the fourth chunk, 16 % of the kernel's DMMA work)
            const int nks = ch == nch - 1 ? (a.p.M - ch * GP_KC + 3) >> 2 : GP_KC / 4;
            if (ntv == 8 && mtv == 2) {
#pragma unroll 2
                for (int ks = 0; ks < nks; ks++) {
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
                for (int ks = 0; ks < nks; ks++) {
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

        if (STORE) {
            double *Rv = a.R[job] + v * a.r_stride[job] + (size_t)(i0 + wrow + g) * a.ldr[job] + jt * GP_TJ + 2 * t4;
#pragma unroll
            for (int mt = 0; mt < 2; mt++)
#pragma unroll
                for (int nt = 0; nt < 8; nt++)
                    *reinterpret_cast<double2 *>(Rv + (size_t)(8 * mt) * a.ldr[job] + 8 * nt) =
                        make_double2(acc[mt][nt][0], acc[mt][nt][1]);
        }
        // ---- closed-form screening of the thread's 32 pairs (see k_fast_tiles, level 2) ----
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
                    double num, det, re, za, zb, gadd;
                    pair_gain<CSF>(rcopy[q], mt ? z1[1] : z1[0], cq[c], mt ? b1[1] : b1[0],
                                   CSF ? cq[GP_TJ + c] : 0.0, mt ? k1[1] : k1[0], CSF ? cq[2 * GP_TJ + c] : 0.0,
                                   mt ? g1[1] : g1[0], CSF ? cq[3 * GP_TJ + c] : 0.0, mt ? zu1[1] : zu1[0],
                                   CSF ? cq[4 * GP_TJ + c] : 0.0, Y3, gain_c, num, det, re, za, zb, gadd);
                    if (!(det > 1e-12)) { gill = INFINITY; continue; }
                    const double rdet = 1.0 / det;
                    double gq = num * rdet, tq = c0 * rdet;
                    if (!(gq + tq >= thr)) continue;
                    if (!FT_DEBUG(a, 4)) refine_pair(re, za, zb, gadd, rdet, c0, c1, vp[0], gq, tq);
                    if (tq > kIllTol * c0) gill = fmax(gill, gq + tq);
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
        const int64_t o = (v * a.njobs + job) * a.ntI + blockIdx.x;
        a.cta_gain[o] = G;
        a.cta_tol[o] = tolG;
        a.cta_idx[o] = I == INT_MAX ? -1 : I;
        a.cta_flag[o] = s_flag;
        a.cta_ill[o] = Gill;
    }
}

// ---------------------------------------------------------------------------------
// Triple scan: three searched blocks [N1, N2, N3] (three fascicles, or two fascicles + an
// E-column compartment; reference `_3`, mf_utils.py:470-607).  The three normalised
// correlation matrices of a voxel come from k_gemm_pairs<0, 1> (DMMA, stored to R; its
// screening pass also yields the best 2-column solutions, i.e. every branch of `_3` that
// does not depend on the third index).  k_triples then enumerates all N1*N2*N3 tuples on the
// FP64 pipe: a CTA owns a (T1 x T2) tile of (i1, i2) pairs, every thread keeps 2 x 4 pairs in
// registers and streams over i3 through a double-buffered shared-memory ring of R13^T / R23^T rows
// filled by TMA bulk copies.  Everything is written in the basis that eliminates atom 1 first, so that
// what depends on (i1, i3) only is shared by the thread's four i2 columns.  With
// c33 = 1 - r12^2, U2 = z2 - r12 z1 (per pair), m13 = 1 - r13^2, d13 = z3 - r13 z1 (per (i1, i3),
// 2 operations for 4 tuples), q2 = r23 - r12 r13, q1 = r13 - r12 r23:
//     S     = c33 m13 - q2^2                (3x3 determinant of the correlation matrix)
//     delta = d13 - q2 (U2 / c33)           (Cramer numerator of w3 over c33)
//     W1 / c33 = (U1 / c33) S - q1 delta,  W2 / c33 = (U2 / c33) S - q2 delta
//     gain >= thr  <=>  delta^2 >= S (thr c33 - n2) / c33^2,   n2 = z1 U1 + z2 U2
// No division; signs are read off the high words with integer instructions.  The vote over 4 i3
// steps (32 tuples per thread) tests gain + error bound >= thr (in the delta |delta| form, which
// also asks for w3 > 0) and w2 > 0: 8.5 FP64 operations per tuple (6.5 for the CSF-projected scan,
// which tests the gain only).  Only when some lane passes does the warp evaluate the sign of W1
// as well -- 3 more operations, q1 is needed for nothing else -- straight-line for the flagged
// steps, and only what passes both levels is walked tuple by tuple.  (Round 1 evaluated S, D3, W1,
// W2 in the symmetric form, all in the vote: 13 / 9 operations; with W1 in the first level the
// scan ran at 28.0 k instead of 34.6 k voxels/s at [300, 300, 300].)  The isolated loop (tools/triples_loop_bench.cu) runs
// at 1.10 T tuples/s whatever the thread tile or the number of warps: FP64 instructions hold the
// issue port for ~2.2 cycles each and every other instruction (sign logic, shared-memory loads)
// adds its own cycle, so the count of instructions per tuple is what sets the pace.
// ---------------------------------------------------------------------------------
#define TR_KC_MAX 128        // i3 steps per chunk: as many as fit in shared memory, at most this
#define TR_CSF_CTAS 3         // CSF-projected scan: CTAs per SM.  It keeps the caller's block order (336 CTAs of 12 steps
                              // per voxel at N = 1000, E = 10) and is bound by the prologue / reduction of its CTAs: three
                              // small CTAs per SM overlap them (1 / 2 / 3 / 4 CTAs: 17.1 k / 19.1 k / 20.5 k / 19.3 k voxels/s)
#define TR_MAXTHREADS 384     // per SM: one CTA of up to 384 threads, or two of up to 192 (k_triples<CSF, 2>) when
                              // the tiling cannot use more than 256 threads in one CTA (a short tiled block: T1 = 2 x 5
                              // for the 10-atom EAR block leaves 240 threads) -- [N,N,E] +6 %, [300,300,300] -1 %
#ifndef TR_W1VOTE
#define TR_W1VOTE 0          // 1: the sign of W1 is part of the first-level vote (11.5 operations per tuple)
#endif

#ifdef TR_COUNT
__device__ unsigned long long g_tr_counters[4];   // votes, CSF tuples passing the gain test, votes that hit, competitive tuples
#endif

struct TripleGeom {
    int txt, tyt, T1, T2, nt1, nt2, threads, ctas;
};

// CSF: the three searched blocks are projected off a fourth, single-column block (two
// fascicles + CSF + the EAR block of MFModel.fit, reference `_4up`).  The scan then tests the
// UNCONSTRAINED gain of the projected triple against the threshold minus the CSF share -- a
// necessary condition for any non-negative solution on the tuple's four columns -- and the
// tuples that pass are solved in closed form on the two supports the pair jobs do not cover.
template <int CSF, int CTAS>
__global__ void __launch_bounds__(TR_MAXTHREADS / CTAS, CTAS) k_triples(FastArgs a)
{
    extern __shared__ __align__(16) double smem[];
    const int TXT = a.tr_txt, TYT = a.tr_tyt;
    const int T1 = 2 * TXT, T2 = 4 * TYT, rowlen = T1 + T2;
    const int KC = a.tr_kc;                               // i3 steps per shared-memory chunk (multiple of 4)
    const int N1 = a.Nb[0], N2 = a.Nb[1], N3 = a.Nb[2];
    double *z3s = smem + (size_t)2 * KC * rowlen;       // [N3]
    double *red = z3s + ((N3 + 3) & ~3);                   // [64]
    __shared__ unsigned long long s_thr;
    __shared__ unsigned long long s_cfull[2];            // chunk ring: bytes landed (TMA bulk copies)
    __shared__ double s_tolG;
    __shared__ int s_flag;

    const int64_t v = blockIdx.y;
    const int t1 = blockIdx.x % a.tr_nt1, t2 = blockIdx.x / a.tr_nt1;
    const int i10 = t1 * T1, i20 = t2 * T2;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    const bool active = tid < TXT * TYT;
    const int tx = active ? tid % TXT : 0, ty = active ? tid / TXT : 0;
    const double *vp = a.voxp + v * FT_VP;
    const double c0 = vp[4];
    const double gshift = CSF ? vp[3] : 0.0;      // CSF share of the gain: the projected scan works on thr - gshift
    const double *cpz1 = a.colp + ((v * 3 + 0) * (int64_t)FT_NPAR + 2) * a.Npad;
    const double *cpz2 = a.colp + ((v * 3 + 1) * (int64_t)FT_NPAR + 2) * a.Npad;
    const double *cpz3 = a.colp + ((v * 3 + 2) * (int64_t)FT_NPAR + 2) * a.Npad;
    const double *R12 = a.R[0] + v * a.r_stride[0];
    const double *R13T = a.R[1] + v * a.r_stride[1];
    const double *R23T = a.R[2] + v * a.r_stride[2];
    const int ld12 = a.ldr[0], ld13 = a.ldr[1], ld23 = a.ldr[2];
    unsigned long long *vthr = a.vthr + v;

    // ---- voxel-wide starting threshold: k_triple_seed left the best certified lower bound of the
    // 2-column solutions and of the seed lines there; other CTAs of the voxel keep raising it ----
    if (tid == 0) {
        s_thr = *(volatile unsigned long long *)vthr;
        s_flag = 0;
        mbar_init(&s_cfull[0], 1);
        mbar_init(&s_cfull[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    // padding steps (i3 >= N3): zero correlations and a hugely negative z3 give delta |delta| = -1e300,
    // never a candidate (CSF: the sign of delta is not looked at; z3 = 0 and the competitive path
    // skips the step)
    for (int i = tid; i < ((N3 + 3) & ~3); i += blockDim.x) z3s[i] = i < N3 ? cpz3[i] : (CSF ? 0.0 : -1e150);

    // ---- chunk loader: rows i3 of R13^T[:, i1 tile] | R23^T[:, i2 tile], two TMA bulk copies per
    // row issued by the row's thread, completion counted in bytes on the buffer's mbarrier (the
    // cp.async loader of the first version was 5 % of the kernel in address arithmetic) ----
    const int nchunks = (N3 + KC - 1) / KC;
    const int w1 = min(T1, ld13 - i10), w2 = min(T2, ld23 - i20);     // tile columns inside the padded R
    auto load_chunk = [&](int c, double *dst, unsigned long long *bar) {
        const int rows = min(KC, ((N3 + 3) & ~3) - c * KC), real = min(KC, N3 - c * KC);
        if (tid == 0) mbar_expect_tx(bar, (unsigned)(real * (w1 + w2) * sizeof(double)));
        for (int r = tid; r < rows; r += blockDim.x) {
            const int i3 = c * KC + r;
            double *drow = dst + (size_t)r * rowlen;
            if (r < real) {
                bulk_g2s(drow, R13T + (size_t)i3 * ld13 + i10, (unsigned)(w1 * sizeof(double)), bar);
                bulk_g2s(drow + T1, R23T + (size_t)i3 * ld23 + i20, (unsigned)(w2 * sizeof(double)), bar);
                for (int col = w1; col < T1; col++) drow[col] = 0.0;
                for (int col = w2; col < T2; col++) drow[T1 + col] = 0.0;
            } else {
                for (int col = 0; col < rowlen; col++) drow[col] = 0.0;
            }
        }
    };
    load_chunk(0, smem, &s_cfull[0]);

    // ---- the thread's 2 x 4 pairs ----
    // per pair: r12, c33, U1 / c33, U2 / c33 and Tq = (thr' c33 - n2) / c33^2; per thread: z1 of
    // its two rows and c0t = max over its pairs of c0 / c33^2 (the screen's error margin in the
    // scaled test; the largest one keeps the test necessary for every pair of the thread).
    // z2 is only needed when the threshold moves or a tuple is competitive: re-read from colp.
    double r12[8], c33[8], U1p[8], U2p[8], Tq[8], z1r[2];
    double c0t = 0.0;
    unsigned valid = 0;
    bool illpair = false;
    auto zrow = [&](int p) { const int i = i10 + 2 * tx + p; return (active && i < N1) ? __ldg(cpz1 + i) : 0.0; };
    auto zcol = [&](int q) { const int j = i20 + 4 * ty + q; return (active && j < N2) ? __ldg(cpz2 + j) : 0.0; };
    auto retarget = [&](double th) {
        double z2[4];
#pragma unroll
        for (int q = 0; q < 4; q++) z2[q] = zcol(q);
#pragma unroll
        for (int e = 0; e < 8; e++) {
            // n2 / c33^2 = z1 U1p / c33 + z2 U2p / c33
            const double ic = 1.0 / c33[e];
            Tq[e] = ic * ((th - gshift) - fma(z1r[e >> 2], U1p[e], z2[e & 3] * U2p[e]));
        }
    };
    __syncthreads();
    double thr = __longlong_as_double((long long)s_thr);
#pragma unroll
    for (int p = 0; p < 2; p++) z1r[p] = zrow(p);
#pragma unroll
    for (int p = 0; p < 2; p++)
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int i = i10 + 2 * tx + p, j = i20 + 4 * ty + q, e = p * 4 + q;
            bool ok = active && i < N1 && j < N2;
            const double zz2 = zcol(q);
            double rr = ok ? R12[(size_t)i * ld12 + j] : 0.0;
            // (numerically) identical atoms in blocks 1 and 2, e.g. two fascicles along the same
            // peak: every tuple of the pair is singular -> the voxel goes to the exact tier
            if (ok && fma(-rr, rr, 1.0) < 1e-10) { illpair = true; ok = false; rr = 0.0; }
            if (ok) valid |= 1u << e;
            r12[e] = rr;
            c33[e] = fma(-rr, rr, 1.0);
            const double ic = 1.0 / c33[e];
            U1p[e] = ok ? fma(-rr, zz2, z1r[p]) * ic : 0.0;
            U2p[e] = ok ? fma(-rr, z1r[p], zz2) * ic : 0.0;
            c0t = fmax(c0t, c0 * ic * ic);
        }
    retarget(thr);

    double gb = -1.0, tb = 0.0, gill = -1.0;
    long long bidx = -1;
    int flag = 0;
    if (illpair) gill = INFINITY;

    for (int c = 0; c < nchunks; c++) {
        if (c + 1 < nchunks) load_chunk(c + 1, smem + (size_t)((c + 1) & 1) * KC * rowlen, &s_cfull[(c + 1) & 1]);
        const double *B = smem + (size_t)(c & 1) * KC * rowlen;
        mbar_wait(&s_cfull[c & 1], (unsigned)(c >> 1) & 1u);
        const int rows = min(KC, N3 - c * KC);
        {   // pick up the voxel-wide threshold raised by other CTAs / warps
            double tn = fmax(__longlong_as_double((long long)s_thr),
                             __longlong_as_double((long long)*(volatile unsigned long long *)vthr));
            if (tn > thr) { thr = tn; retarget(thr); }
        }
#pragma unroll 1
        for (int r0 = 0; r0 < rows; r0 += 4) {
            // four i3 steps per vote: one basic block of 32 independent tuples for the scheduler.
            // The gain test is t = delta |delta| + c0t - Tq S >= 0: the |delta| form makes a negative
            // w3 fail (unless the pair alone already reaches the threshold) without a sign word of its
            // own.  sgn[s] keeps the AND of the tuples' sign words of step s: bit 31 clear <=> some
            // tuple of that step has positive weights and gain + bound >= thr.
            int sgn[4];
#pragma unroll
            for (int s4 = 0; s4 < 4; s4++) {
                const double *rowp = B + (size_t)(r0 + s4) * rowlen;
                const double2 r13v = *reinterpret_cast<const double2 *>(rowp + 2 * tx);
                const double2 r23a = *reinterpret_cast<const double2 *>(rowp + T1 + 4 * ty);
                const double2 r23b = *reinterpret_cast<const double2 *>(rowp + T1 + 4 * ty + 2);
                const double z3 = z3s[c * KC + r0 + s4];
                const double r13[2] = {r13v.x, r13v.y};
                const double r23[4] = {r23a.x, r23a.y, r23b.x, r23b.y};
                int sall = -1;
#pragma unroll
                for (int p = 0; p < 2; p++) {
                    const double m13 = fma(-r13[p], r13[p], 1.0);
                    const double d13 = fma(-r13[p], z1r[p], z3);
#pragma unroll
                    for (int q = 0; q < 4; q++) {
                        const int e = p * 4 + q;
                        const double q2 = fma(-r12[e], r13[p], r23[q]);
                        const double S = fma(-q2, q2, c33[e] * m13);
                        const double dl = fma(-q2, U2p[e], d13);
                        const double t = fma(-Tq[e], S, fma(dl, CSF ? dl : fabs(dl), c0t));
                        if (CSF) {
                            sall &= __double2hiint(t);
                        } else {
                            const double W2 = fma(-q2, dl, U2p[e] * S);
#if TR_W1VOTE
                            const double q1 = fma(-r12[e], r23[q], r13[p]);
                            const double W1 = fma(-q1, dl, U1p[e] * S);
                            sall &= (__double2hiint(W1) | __double2hiint(W2)) | __double2hiint(t);
#else
                            sall &= __double2hiint(W2) | __double2hiint(t);
#endif
                        }
                    }
                }
                sgn[s4] = sall;
            }
            int sany = (sgn[0] & sgn[1]) & (sgn[2] & sgn[3]);
            bool hit = __any_sync(0xffffffffu, sany >= 0);
#if !TR_W1VOTE
            if (!CSF && hit) {
                // second level, only for the steps in which some lane passed the first (warp-uniform
                // mask): the same test with the sign of W1 as well, straight-line over the thread's 8
                // tuples, before any lane walks its tuples one by one
                const unsigned m1 = __reduce_or_sync(0xffffffffu, (sgn[0] >= 0 ? 1u : 0u) | (sgn[1] >= 0 ? 2u : 0u) |
                                                                      (sgn[2] >= 0 ? 4u : 0u) | (sgn[3] >= 0 ? 8u : 0u));
#pragma unroll 1
                for (int s4 = 0; s4 < 4; s4++) {
                    if (!(m1 >> s4 & 1u)) continue;
                    const double *rowp = B + (size_t)(r0 + s4) * rowlen;
                    const double2 r13v = *reinterpret_cast<const double2 *>(rowp + 2 * tx);
                    const double2 r23a = *reinterpret_cast<const double2 *>(rowp + T1 + 4 * ty);
                    const double2 r23b = *reinterpret_cast<const double2 *>(rowp + T1 + 4 * ty + 2);
                    const double z3 = z3s[c * KC + r0 + s4];
                    const double r13[2] = {r13v.x, r13v.y};
                    const double r23[4] = {r23a.x, r23a.y, r23b.x, r23b.y};
                    int sall = -1;
#pragma unroll
                    for (int p = 0; p < 2; p++) {
                        const double m13 = fma(-r13[p], r13[p], 1.0);
                        const double d13 = fma(-r13[p], z1r[p], z3);
#pragma unroll
                        for (int q = 0; q < 4; q++) {
                            const int e = p * 4 + q;
                            const double q2 = fma(-r12[e], r13[p], r23[q]);
                            const double S = fma(-q2, q2, c33[e] * m13);
                            const double dl = fma(-q2, U2p[e], d13);
                            const double t = fma(-Tq[e], S, fma(dl, fabs(dl), c0t));
                            const double W2 = fma(-q2, dl, U2p[e] * S);
                            const double q1 = fma(-r12[e], r23[q], r13[p]);
                            const double W1 = fma(-q1, dl, U1p[e] * S);
                            sall &= (__double2hiint(W1) | __double2hiint(W2)) | __double2hiint(t);
                        }
                    }
                    if (s4 == 0) sgn[0] = sall; else if (s4 == 1) sgn[1] = sall; else if (s4 == 2) sgn[2] = sall; else sgn[3] = sall;
                }
                sany = (sgn[0] & sgn[1]) & (sgn[2] & sgn[3]);
                hit = __any_sync(0xffffffffu, sany >= 0);
            }
#endif
#ifdef TR_COUNT
            if (lane == 0) { atomicAdd(&g_tr_counters[0], 1ull); if (hit) atomicAdd(&g_tr_counters[2], 1ull); }
#endif
            if (hit) {
                if (sany >= 0) {
                    const double y_sq = vp[0];
                    const double c1 = y_sq > 0 ? c0 / y_sq : 0.0, ynorm = sqrt(y_sq);
                    const double tmax = 16.0 * c0, wide = 4.0 * tmax;
                    double pc[8];
#pragma unroll
                    for (int e = 0; e < 8; e++) pc[e] = r12[e];
                    const int smask = (sgn[0] >= 0 ? 1 : 0) | (sgn[1] >= 0 ? 2 : 0) | (sgn[2] >= 0 ? 4 : 0) | (sgn[3] >= 0 ? 8 : 0);
#pragma unroll 1
                    for (int s4 = 0; s4 < 4; s4++) {
                        if (!(smask >> s4 & 1)) continue;
                        const int r = r0 + s4;
                        const double *rowp = B + (size_t)r * rowlen;
                        const double z3 = z3s[c * KC + r];
#pragma unroll 1
                        for (int e = 0; e < 8; e++) {
                            if (!(valid >> e & 1u)) continue;
                            const int p = e >> 2, q = e & 3;
                            const double zz1 = zrow(p), zz2 = zcol(q);
                            const double a12 = pc[e], k33 = fma(-a12, a12, 1.0);
                            const double u1 = fma(-a12, zz2, zz1), u2 = fma(-a12, zz1, zz2);
                            const double a13 = rowp[2 * tx + p], a23 = rowp[T1 + 4 * ty + q];
                            const double q1 = fma(-a12, a23, a13), q2 = fma(-a12, a13, a23);
                            const double S = fma(-a23, q2, fma(-a13, q1, k33));
                            const double D3 = fma(-a23, u2, fma(-a13, u1, k33 * z3));
                            const double W1 = fma(-q1, D3, u1 * S), W2 = fma(-q2, D3, u2 * S);
                            if (CSF) {
                                // Four columns.  The tuple's non-negative optimum is the best of the supports whose
                                // unconstrained solution is positive.  Supports without the third block's atom, or
                                // with only one fascicle atom, belong to the pair jobs (bounded by g2); what is left:
                                // the full support (all four positive) and {atom 1, atom 2, atom 3} without the CSF
                                // column -- two closed forms instead of a 15-support enumeration.
                                if (c * KC + r >= N3) continue;
                                if (!(fma(-fma(thr - gshift, k33, -fma(zz1, u1, zz2 * u2)), S, fma(D3, D3, c0)) >= 0.0)) continue;
#ifdef TR_COUNT
                                atomicAdd(&g_tr_counters[1], 1ull);
#endif
                                const int i1 = i10 + 2 * tx + p, i2 = i20 + 4 * ty + q, i3 = c * KC + r;
                                const double *P1 = a.colp + (v * 3 + 0) * (int64_t)FT_NPAR * a.Npad;
                                const double *P2 = a.colp + (v * 3 + 1) * (int64_t)FT_NPAR * a.Npad;
                                const double *P3 = a.colp + (v * 3 + 2) * (int64_t)FT_NPAR * a.Npad;
                                const double ka = P1[(size_t)4 * a.Npad + i1], ga = P1[(size_t)5 * a.Npad + i1];
                                const double kb = P2[(size_t)4 * a.Npad + i2], gbb = P2[(size_t)5 * a.Npad + i2];
                                const double kc = P3[(size_t)4 * a.Npad + i3], gc = P3[(size_t)5 * a.Npad + i3];
                                const double y_sq = vp[0];
                                const double c1 = y_sq > 0 ? c0 / y_sq : 0.0, ynorm = sqrt(y_sq);
                                const double zc = vp[2] * rsqrt(vp[1]);              // csf . y on the unit CSF column
                                const double zua = P1[(size_t)6 * a.Npad + i1], zub = P2[(size_t)6 * a.Npad + i2],
                                             zuc = P3[(size_t)6 * a.Npad + i3];
                                double gq = -1.0, tq = 0.0;
                                const double dd = k33 * S;
                                if (W1 > 0.0 && W2 > 0.0 && D3 > 0.0 && dd > 1e-13 && S > 0.0) {
                                    // full support: weights on the unit columns from the projected solution
                                    const double wa = W1 / (dd * ka), wb = W2 / (dd * kb), wcc = D3 / (S * kc);
                                    const double wcsf = zc - fma(wa, ga, fma(wb, gbb, wcc * gc));
                                    if (wcsf > 0.0) {
                                        gq = gshift + fma(fma(zz1, u1, zz2 * u2), S, D3 * D3) / dd;
                                        const double sw = wa + wb + wcc + wcsf, rel = c1 / dd;
                                        tq = fmin(4.0 * c0 / dd, 2.0 * c1 * fma(sw, sw, sw * ynorm) + 4.0 * rel * rel * y_sq);
                                    }
                                }
                                {   // support {1, 2, 3} without the CSF column: unprojected 3 x 3 Cramer on unit columns
                                    const double r12 = fma(ka * kb, a12, ga * gbb), r13 = fma(ka * kc, a13, ga * gc),
                                                 r23 = fma(kb * kc, a23, gbb * gc);
                                    const double m1 = fma(-r23, r23, 1.0), m2 = fma(-r13, r23, r12), m3 = fma(r12, r23, -r13);
                                    const double det = fma(-r12, m2, fma(r13, m3, m1));
                                    if (det > 1e-13) {
                                        const double e1 = fma(zua, m1, fma(-zub, m2, zuc * m3));
                                        const double e2 = fma(-zua, m2, fma(zub, fma(-r13, r13, 1.0), -zuc * fma(-r12, r13, r23)));
                                        const double e3 = fma(zua, m3, fma(-zub, fma(-r12, r13, r23), zuc * fma(-r12, r12, 1.0)));
                                        if (e1 > 0.0 && e2 > 0.0 && e3 > 0.0) {
                                            const double wa = e1 / det, wb = e2 / det, wcc = e3 / det;
                                            // stationary form 2 w.z - w'Gw (second order in the weight error)
                                            const double quad = fma(wa, wa, fma(wb, wb, wcc * wcc)) +
                                                                2.0 * fma(wa * wb, r12, fma(wa * wcc, r13, wb * wcc * r23));
                                            const double g3 = 2.0 * fma(wa, zua, fma(wb, zub, wcc * zuc)) - quad;
                                            const double sw = wa + wb + wcc, rel = c1 / det;
                                            const double t3 = 2.0 * c1 * fma(sw, sw, sw * ynorm) + 4.0 * rel * rel * y_sq;
                                            if (g3 > gq) { gq = g3; tq = t3; }
                                        }
                                    } else {
                                        gill = INFINITY;         // numerically singular triple
                                    }
                                }
                                if (gq < 0.0) continue;           // the tuple's optimum lies on a support of the pair jobs
                                if (!(gq + tq >= thr)) continue;
#ifdef TR_COUNT
                                atomicAdd(&g_tr_counters[3], 1ull);
#endif
                                // (all of these margins are ~1e-11 of |y|^2, far below the gaps between tuples)
                                const double tmax4 = 256.0 * c0, wide4 = 4.0 * tmax4;
                                if (tq > tmax4) gill = fmax(gill, gq + tq);
                                if (gq > gb) {
                                    flag = (bidx >= 0 && !(gq > gb + wide4)) ? 1 : 0;
                                    gb = gq; tb = tq;
                                    bidx = ((long long)i3 * N1 + i1) * N2 + i2;
                                } else if (!(gb > gq + wide4)) {
                                    flag = 1;
                                }
                                continue;
                            }
                            if (!(W1 > 0.0 && W2 > 0.0 && D3 > 0.0)) continue;
                            const double dd = k33 * S;
                            if (!(dd > 1e-13 && S > 0.0)) { gill = INFINITY; continue; }   // numerically singular
                            // Cramer-form gain and its (pessimistic) evaluation error bound
                            const double n2 = fma(zz1, u1, zz2 * u2);
                            double gq = fma(n2, S, D3 * D3) / dd, tq = c0 / dd;
                            if (!(gq + tq >= thr)) continue;
                            // refined: weights from Cramer, gain from the stationary form 2 w.z - w'Gw,
                            // whose error is second order in the weight error
                            const double w1 = W1 / dd, w2 = W2 / dd, w3 = D3 / S;
                            const double quad = fma(w1, w1, fma(w2, w2, w3 * w3)) +
                                                2.0 * fma(w1 * w2, a12, fma(w1 * w3, a13, w2 * w3 * a23));
                            const double gr = 2.0 * fma(w1, zz1, fma(w2, zz2, w3 * z3)) - quad;
                            const double sw = w1 + w2 + w3, rel = c1 / dd;
                            const double tr = c1 * fma(sw, sw, sw * ynorm) + 4.0 * rel * rel * y_sq;
                            if (tr < tq) { gq = gr; tq = tr; }
                            if (tq > tmax) gill = fmax(gill, gq + tq);
                            if (gq > gb) {
                                flag = (bidx >= 0 && !(gq > gb + wide)) ? 1 : 0;
                                gb = gq; tb = tq;
                                bidx = ((long long)(c * KC + r) * N1 + (i10 + 2 * tx + p)) * N2 + (i20 + 4 * ty + q);
                            } else if (!(gb > gq + wide)) {
                                flag = 1;
                            }
                        }
                    }
                }
                double lb = bidx >= 0 ? gb - tb : 0.0;               // certified lower bound
                for (int o = 16; o > 0; o >>= 1) lb = fmax(lb, __shfl_xor_sync(0xffffffffu, lb, o));
                if (lb > thr) {
                    if (lane == 0) {
                        atomicMax(&s_thr, (unsigned long long)__double_as_longlong(lb));
                        atomicMax(vthr, (unsigned long long)__double_as_longlong(lb));
                    }
                }
                const double tn = fmax(lb, __longlong_as_double((long long)*(volatile unsigned long long *)&s_thr));
                if (tn > thr) { thr = tn; retarget(thr); }
            }
        }
        __syncthreads();       // every thread is done with this buffer before its next load is issued
    }

    // ---- reduction over the CTA: best gain, tie -> lower loop index ----
    const double gt = bidx >= 0 ? gb : -1.0;
    const double tolt = bidx >= 0 ? tb : 0.0;
    double gm = gt;
    long long im = bidx >= 0 ? bidx : LLONG_MAX;
    for (int o = 16; o > 0; o >>= 1) {
        double og = __shfl_xor_sync(0xffffffffu, gm, o);
        long long oi = __shfl_xor_sync(0xffffffffu, im, o);
        if (og > gm || (og == gm && oi < im)) { gm = og; im = oi; }
    }
    for (int o = 16; o > 0; o >>= 1) gill = fmax(gill, __shfl_xor_sync(0xffffffffu, gill, o));
    double *redg = red, *redl = red + 16;
    long long *redi = (long long *)(red + 32);
    if (lane == 0) { redg[warp] = gm; redi[warp] = im; redl[warp] = gill; }
    __syncthreads();
    double G = redg[0], Gill = redl[0];
    long long I = redi[0];
    for (int w = 1; w < nwarps; w++) {
        if (redg[w] > G || (redg[w] == G && redi[w] < I)) { G = redg[w]; I = redi[w]; }
        Gill = fmax(Gill, redl[w]);
    }
    if (bidx >= 0 && bidx == I) s_tolG = tolt;
    __syncthreads();
    const double tolG = I != LLONG_MAX ? s_tolG : 0.0;
    if (bidx >= 0) {
        const bool winner = bidx == I;
        const bool close = gt + tolt >= G - tolG;
        if ((winner && flag) || (!winner && close)) atomicOr(&s_flag, 1);
    }
    __syncthreads();
    if (tid == 0) {
        const int64_t o = v * a.tr_ntiles + blockIdx.x;
        a.t_gain[o] = G;
        a.t_tol[o] = tolG;
        a.t_idx[o] = I == LLONG_MAX ? -1 : I;
        a.t_flag[o] = s_flag;
        a.t_ill[o] = Gill;
    }
}

// Seed of the triple scan's voxel-wide threshold.  The pair jobs give the best two-block
// solutions; the optimum triple almost always contains one of those pairs, so the three LINES of
// tuples through them (best pair of blocks 1-2 with every atom of block 3, best pair 3-1 with
// every atom of block 2, best pair 3-2 with every atom of block 1) are evaluated first and their
// best certified lower bound becomes the starting threshold: without it every CTA of k_triples
// starts from the two-block bound and floods its competitive path until the first good triple is
// met.  grid V, 128 threads.
__global__ void __launch_bounds__(128) k_triple_seed(FastArgs a)
{
    __shared__ double s_red[4];
    const int64_t v = blockIdx.x;
    const double *vp = a.voxp + v * FT_VP;
    const double c0 = vp[4], gshift = a.csf ? vp[3] : 0.0;
    const double *R12 = a.R[0] + v * a.r_stride[0], *R13T = a.R[1] + v * a.r_stride[1], *R23T = a.R[2] + v * a.r_stride[2];
    const double *P[3];
    for (int k = 0; k < 3; k++) P[k] = a.colp + (v * 3 + k) * (int64_t)FT_NPAR * a.Npad;
    const double zc = a.csf ? vp[2] * rsqrt(vp[1]) : 0.0;
    double best = 0.0;
    for (int j = 0; j < 3; j++) {
        // best pair of job j
        const int rb = a.job_rb[j], cb = a.job_cb[j], ob = 3 - rb - cb;
        const int ntj = (a.Nb[rb] + GP_TI - 1) / GP_TI;
        double G = -1.0;
        int I = -1;
        for (int t = 0; t < ntj; t++) {
            const int64_t o = (v * 3 + j) * a.ntI + t;
            if (a.cta_idx[o] >= 0 && a.cta_gain[o] > G) { G = a.cta_gain[o]; I = a.cta_idx[o]; }
        }
        if (I < 0) continue;
        int iw[3];
        iw[rb] = I / a.Nb[cb]; iw[cb] = I - iw[rb] * a.Nb[cb];
        for (int io = threadIdx.x; io < a.Nb[ob]; io += blockDim.x) {
            iw[ob] = io;
            const double r12 = R12[(size_t)iw[0] * a.ldr[0] + iw[1]];
            const double r13 = R13T[(size_t)iw[2] * a.ldr[1] + iw[0]];
            const double r23 = R23T[(size_t)iw[2] * a.ldr[2] + iw[1]];
            const double z1 = P[0][(size_t)2 * a.Npad + iw[0]], z2 = P[1][(size_t)2 * a.Npad + iw[1]],
                         z3 = P[2][(size_t)2 * a.Npad + iw[2]];
            const double c33 = fma(-r12, r12, 1.0), U1 = fma(-r12, z2, z1), U2 = fma(-r12, z1, z2);
            const double q1 = fma(-r12, r23, r13), q2 = fma(-r12, r13, r23);
            const double S = fma(-r23, q2, fma(-r13, q1, c33));
            const double D3 = fma(-r23, U2, fma(-r13, U1, c33 * z3));
            const double W1 = fma(-q1, D3, U1 * S), W2 = fma(-q2, D3, U2 * S);
            const double dd = c33 * S;
            if (!(W1 > 0.0 && W2 > 0.0 && D3 > 0.0 && S > 0.0 && dd > 1e-9)) continue;
            if (a.csf) {    // the CSF weight of the four-column solution must be positive too
                const double wa = W1 / (dd * P[0][(size_t)4 * a.Npad + iw[0]]), wb = W2 / (dd * P[1][(size_t)4 * a.Npad + iw[1]]),
                             wc = D3 / (S * P[2][(size_t)4 * a.Npad + iw[2]]);
                const double wcsf = zc - fma(wa, P[0][(size_t)5 * a.Npad + iw[0]],
                                             fma(wb, P[1][(size_t)5 * a.Npad + iw[1]], wc * P[2][(size_t)5 * a.Npad + iw[2]]));
                if (!(wcsf > 0.0)) continue;
            }
            const double gq = gshift + fma(fma(z1, U1, z2 * U2), S, D3 * D3) / dd;
            best = fmax(best, gq - 4.0 * c0 / dd - c0);        // certified lower bound (Cramer-form error bound)
        }
    }
    for (int o = 16; o > 0; o >>= 1) best = fmax(best, __shfl_xor_sync(0xffffffffu, best, o));
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = best;
    __syncthreads();
    if (threadIdx.x == 0) {
        best = fmax(fmax(s_red[0], s_red[1]), fmax(s_red[2], s_red[3]));
        // certified lower bounds of the 1- and 2-column solutions (pair jobs)
        double t0 = fmax(fmax(vp[5], vp[6]), vp[7]);
        for (int j = 0; j < 3; j++) {
            const int ntj = (a.Nb[a.job_rb[j]] + GP_TI - 1) / GP_TI;
            for (int t = 0; t < ntj; t++) {
                const int64_t o = (v * 3 + j) * a.ntI + t;
                if (a.cta_idx[o] >= 0) t0 = fmax(t0, a.cta_gain[o] - a.cta_tol[o]);
            }
        }
        best = fmax(best, fmax(t0 - c0, 0.0));
        if (best > 0.0) atomicMax(a.vthr + v, (unsigned long long)__double_as_longlong(best));
    }
}

// select for the triple scan: one thread per voxel
__global__ void __launch_bounds__(128) k_select3(FastArgs a, int64_t V)
{
    int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= V) return;
    const double *vp = a.voxp + v * FT_VP;
    const double c0 = vp[4];
    // optimistic gain of every solution with at most two active columns
    double g2 = fmax(fmax(vp[5], vp[6]), vp[7]);
    for (int j = 0; j < 3; j++) {
        const int ntj = (a.Nb[a.job_rb[j]] + GP_TI - 1) / GP_TI;
        for (int t = 0; t < ntj; t++) {
            const int64_t o = (v * 3 + j) * a.ntI + t;
            if (a.cta_idx[o] >= 0) g2 = fmax(g2, a.cta_gain[o] + a.cta_tol[o]);
            g2 = fmax(g2, a.cta_ill[o]);
        }
    }
    double G = -1.0, tolG = 0.0;
    long long I = -1;
    int best_t = -1;
    for (int t = 0; t < a.tr_ntiles; t++) {
        const int64_t o = v * a.tr_ntiles + t;
        if (a.t_idx[o] >= 0 && (a.t_gain[o] > G || (a.t_gain[o] == G && a.t_idx[o] < I))) {
            G = a.t_gain[o]; tolG = a.t_tol[o]; I = a.t_idx[o]; best_t = t;
        }
    }
    bool certain = I >= 0;
    int reason = certain ? -1 : 0;
    for (int t = 0; t < a.tr_ntiles && certain; t++) {
        const int64_t o = v * a.tr_ntiles + t;
        if (a.t_ill[o] >= G - tolG) { certain = false; reason = 1; }
        if (t == best_t) { if (a.t_flag[o]) { certain = false; reason = 2; } continue; }
        if (a.t_idx[o] >= 0 && a.t_gain[o] + a.t_tol[o] >= G - tolG) { certain = false; reason = 2; }
    }
    if (certain && !(G - tolG > g2 + 16.0 * c0)) { certain = false; reason = 3; }
    if (certain && fmax(fmax(vp[12], vp[13]), vp[14]) < kCramerScaleMin) { certain = false; reason = 3; }
    if (!certain && !(fmax(fmax(vp[12], vp[13]), vp[14]) < kCramerScaleMin)) {
        // One of the three blocks may simply be inactive: the best two-block solution (job 0: blocks
        // 1-2, job 1: blocks 3-1, job 2: blocks 3-2; both weights positive, with the CSF column when
        // there is one) wins when it is certain inside its own scan and clearly above (i) every
        // solution of the other two jobs and every one-atom solution, (ii) every tuple in which all
        // three blocks are active (the tuples the scan did not look at lie below its threshold, which
        // started at the jobs' certified bounds).  The reference then returns the FIRST tuple of its
        // loop order that contains the pair: the inactive block's index is 0 (`_3`: the 2-column
        // sub-problem is the same for every atom of the inactive block; `_4up`: same active set).
        double Gj[3], tolj[3];
        int Ij[3], tbj[3], jb = 0;
        for (int j = 0; j < 3; j++) {
            Gj[j] = -1.0; tolj[j] = 0.0; Ij[j] = -1; tbj[j] = -1;
            const int ntj = (a.Nb[a.job_rb[j]] + GP_TI - 1) / GP_TI;
            for (int t = 0; t < ntj; t++) {
                const int64_t o = (v * 3 + j) * a.ntI + t;
                if (a.cta_idx[o] >= 0 && (a.cta_gain[o] > Gj[j] || (a.cta_gain[o] == Gj[j] && a.cta_idx[o] < Ij[j]))) {
                    Gj[j] = a.cta_gain[o]; tolj[j] = a.cta_tol[o]; Ij[j] = a.cta_idx[o]; tbj[j] = t;
                }
            }
            if (Gj[j] > Gj[jb]) jb = j;
        }
        const int I0 = Ij[jb];
        bool okb = I0 >= 0;
        const double lower = Gj[jb] - tolj[jb], sep = 16.0 * c0;
        for (int j = 0; j < 3 && okb; j++) {
            const int ntj = (a.Nb[a.job_rb[j]] + GP_TI - 1) / GP_TI;
            for (int t = 0; t < ntj && okb; t++) {
                const int64_t o = (v * 3 + j) * a.ntI + t;
                if (j == jb) {      // inside the winner's own scan: the usual certainty test
                    if (a.cta_ill[o] >= lower) okb = false;
                    if (t == tbj[jb]) { if (a.cta_flag[o]) okb = false; continue; }
                    if (a.cta_idx[o] >= 0 && a.cta_gain[o] + a.cta_tol[o] >= lower) okb = false;
                } else {
                    if (a.cta_idx[o] >= 0 && a.cta_gain[o] + a.cta_tol[o] + sep >= lower) okb = false;
                    if (a.cta_ill[o] + sep >= lower) okb = false;
                }
            }
        }
        if (okb && fmax(fmax(vp[5], vp[6]), vp[7]) + sep >= lower) okb = false;
        for (int t = 0; t < a.tr_ntiles && okb; t++) {
            const int64_t o = v * a.tr_ntiles + t;
            if (a.t_idx[o] >= 0 && a.t_gain[o] + a.t_tol[o] + sep >= lower) okb = false;
            if (a.t_ill[o] + sep >= lower) okb = false;
        }
        // the winner's atoms: job (rb, cb) index = i_rb * N_cb + i_cb; the third block is inactive
        int iw[3] = {0, 0, 0};
        const int rbw = a.job_rb[jb], cbw = a.job_cb[jb], obw = 3 - rbw - cbw;
        if (okb) { iw[rbw] = I0 / a.Nb[cbw]; iw[cbw] = I0 - iw[rbw] * a.Nb[cbw]; }
        if (okb) {
            // `_3` takes the unconstrained 3-column solution whenever its Cramer numerators are
            // >= -tol (mfu:562), also when one is zero to rounding (noise-free data whose weight on
            // the inactive block is exactly 0): the residuals of the tuples (pair, any atom of that
            // block) then differ only by rounding noise and the first minimum cannot be predicted;
            // the support enumeration of four blocks has the same degenerate tie (a weight of
            // +1e-16).  Require, for every atom of the inactive block, a clearly negative numerator of
            // the (CSF-projected) three-column solution: the tuple's optimum then lies on a
            // sub-support, all of which are accounted for above.
            const double *Zb[3];
            for (int k = 0; k < 3; k++) Zb[k] = a.colp + ((v * 3 + k) * (int64_t)FT_NPAR + 2) * a.Npad;
            const double *R12 = a.R[0] + v * a.r_stride[0], *R13T = a.R[1] + v * a.r_stride[1], *R23T = a.R[2] + v * a.r_stride[2];
            const double delta = 1e-9 * sqrt(vp[0]);
            for (int io = 0; io < a.Nb[obw] && okb; io++) {
                iw[obw] = io;
                const double r12 = R12[(size_t)iw[0] * a.ldr[0] + iw[1]];
                const double r13 = R13T[(size_t)iw[2] * a.ldr[1] + iw[0]];
                const double r23 = R23T[(size_t)iw[2] * a.ldr[2] + iw[1]];
                const double z1 = Zb[0][iw[0]], z2 = Zb[1][iw[1]], z3 = Zb[2][iw[2]];
                const double D1 = z1 * (1.0 - r23 * r23) - z2 * (r12 - r13 * r23) + z3 * (r12 * r23 - r13);
                const double D2 = -z1 * (r12 - r13 * r23) + z2 * (1.0 - r13 * r13) - z3 * (r23 - r12 * r13);
                const double D3 = z1 * (r12 * r23 - r13) - z2 * (r23 - r12 * r13) + z3 * (1.0 - r12 * r12);
                if (!(fmin(D1, fmin(D2, D3)) < -delta)) okb = false;
            }
            iw[obw] = 0;
        }
        // `_3` loop index (i3 outermost), re-encoded below for four blocks
        if (okb) { certain = true; reason = -1; I = ((long long)iw[2] * a.Nb[0] + iw[0]) * a.Nb[1] + iw[1]; }
    }
    if (reason >= 0 && a.reasons) atomicAdd(&a.reasons[reason], 1);
    const int64_t row = a.vox_list ? a.vox_list[v] : v;
    if (certain) {
        // scan order -> the caller's blocks -> loop index of the reference
        const long long n1 = a.Nb[0], n2 = a.Nb[1];
        long long is[3] = {(I / n2) % n1, I % n2, I / (n1 * n2)}, ic[3], nc[3];
        for (int k = 0; k < 3; k++) { ic[a.tr_perm[k]] = is[k]; nc[a.tr_perm[k]] = a.Nb[k]; }
        if (a.csf)      // four blocks [N1, N2, 1, N3]: product loop order of the reference's `_4up`
            a.tuple[row] = (ic[0] * nc[1] + ic[1]) * nc[2] + ic[2];
        else            // `_3`: i3 outermost
            a.tuple[row] = (ic[2] * nc[0] + ic[0]) * nc[1] + ic[1];
    } else {
        int pos = atomicAdd(a.redo_count, 1);
        a.redo_list[pos] = (int32_t)row;
        if (a.redo_local) a.redo_local[pos] = (int32_t)v;
    }
}

// ---------------------------------------------------------------------------------
// select: one thread per voxel
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_fast_select(FastArgs a, int64_t V)
{
    int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= V) return;
    const double *vp = a.voxp + v * FT_VP;
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
    // three-block voxels of tiny magnitude: the reference's absolute tolerance decides (kCramerScaleMin)
    if (certain && a.csf && fmax(vp[12], vp[13]) < kCramerScaleMin) { certain = false; reason = 3; }
    if (reason >= 0 && a.reasons) atomicAdd(&a.reasons[reason], 1);
    const int64_t row = a.vox_list ? a.vox_list[v] : v;
    if (certain) {
        a.tuple[row] = (long long)I;
    } else {
        int pos = atomicAdd(a.redo_count, 1);
        a.redo_list[pos] = (int32_t)row;
        if (a.redo_local) a.redo_local[pos] = (int32_t)v;
        if (a.redo_mask) {
            // No competitive pair at all: the winner is a solution with one fascicle atom
            // (alone or with the CSF column).  Every tuple that can reach the minimum then lies
            // in the row of an atom of block 1, or the column of an atom of block 2, whose
            // single-solution gain is within the margin of the best: only those rows / columns
            // need the reference-order search.  Not applicable when the CSF-only or the
            // all-zero solution could win (their first tuple in loop order can be anywhere).
            uint8_t *mk = a.redo_mask + (size_t)pos * 2 * a.mask_ld;
            const double gain_c = vp[11], margin = kPreMargin * c0;
            const bool restricted = reason == 0 && gpre > margin && gain_c < gpre - margin;
            for (int t = 0; t < 2 * a.mask_ld; t++) mk[t] = restricted ? 0 : 1;
            if (restricted)
                for (int k = 0; k < 2; k++) {
                    const double *gs = a.colp + ((v * a.nblk + k) * (int64_t)FT_NPAR + 7) * a.Npad;
                    const int Nk = a.Nb[k];
                    for (int i = 0; i < Nk; i++)
                        if (gs[i] >= gpre - margin) mk[k * a.mask_ld + i] = 1;
                }
        }
    }
}

// ---------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------
static size_t al256(size_t x) { return (x + 255) & ~(size_t)255; }

bool fast_supported(const DevPlan &p, int K, int csf, int ear)
{
    const int Mp = (p.M + 4) & ~3;     // + the folded-screen row
    return K == 2 && !ear && !p.has_between && Mp <= 112 && p.N >= 8 && p.N <= 46000 &&
           (csf == 0 || p.sig_csf);
}

bool fast_supported_materialised(const DevPlan &p, int K, int csf, int ear)
{
    return K == 2 && !ear && p.N >= 8 && p.N <= 46000 && p.M <= 16384 && (csf == 0 || p.sig_csf);
}

// two fascicles + the EAR block ([N, N, E]): triple scan on materialised dictionaries
bool fast3_supported_materialised(const DevPlan &p, int K, int csf, int ear)
{
    return K == 2 && ear && (!csf || p.sig_csf) && p.sig_ear && p.E >= 2 && p.N >= 2 && p.N <= 4096 && p.M <= 16384;
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
    int Mp, Mp2, Npad, ntI, N1pad, ldn, nsplit;
    bool gemm;      // general-M path (k_normalize + k_gemm_pairs)
};
static FastGeom fast_geom(int M, int N1, int N2)
{
    FastGeom g;
    g.Mp = (M + 4) & ~3;               // M + 1 (the folded-screen row of k_fast_tiles) padded to 4
    g.gemm = g.Mp > 112;
#ifdef MFB_EXPERIMENTS
    if (getenv("MFB_FORCE_GEMM")) g.gemm = true;
#endif
    const int Nmax = N1 > N2 ? N1 : N2;
    const int padto = g.gemm ? GP_TI : FT_TJ;
    g.Npad = (Nmax + padto - 1) / padto * padto;
    g.ntI = g.gemm ? (N1 + GP_TI - 1) / GP_TI : (N1 + FT_TI - 1) / FT_TI;
    g.nsplit = 1;
    g.Mp2 = (M + GP_KC - 1) / GP_KC * GP_KC;
    g.N1pad = (N1 + GP_TI - 1) / GP_TI * GP_TI;
    g.ldn = g.N1pad + (N2 + GP_TJ - 1) / GP_TJ * GP_TJ;
    if (g.gemm) {
        // one voxel's normalised copy is Mp2 * ldn * 8 bytes; when a few of them exceed the L2
        // (126 MB) the i1 tile, re-read once per i2 tile, would come from HBM every time: spread
        // a voxel over ~128 CTAs (i1 tiles x i2 slices) so that ~1 voxel is in flight per wave
        const double mb = (double)g.Mp2 * g.ldn * 8.0 / 1e6;
        const int ntJ = (N2 + GP_TJ - 1) / GP_TJ;
        if (mb * 148.0 / g.ntI > 100.0) {
            int ns = (128 + g.ntI - 1) / g.ntI;
            g.nsplit = ns < 1 ? 1 : (ns > ntJ ? ntJ : ns);
            g.ntI *= g.nsplit;
        }
    }
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
    s += al256(sizeof(double) * V * FT_VP);
    s += al256(sizeof(unsigned long long) * V);
    s += 3 * al256(sizeof(double) * V * g.ntI);
    s += 2 * al256(sizeof(int) * V * g.ntI);
    if (g.gemm) s += al256(sizeof(double) * (shared_dict ? 1 : V) * (size_t)g.Mp2 * g.ldn);
    else        // tile-major prepared copy of the streamed block (k_fast_tiles)
        s += al256(sizeof(double) * V * (size_t)((N2 + FT_TJ - 1) / FT_TJ) * FT2_REC(g.Mp));
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
    a.redo_mask = fp.redo_mask; a.mask_ld = fp.mask_ld;
    a.Mp = g.Mp; a.Npad = g.Npad; a.ntI = g.ntI;
    a.Mp2 = g.Mp2; a.N1pad = g.N1pad; a.ldn = g.ldn;
    a.nblk = 2; a.njobs = 1; a.job_rb[0] = 0; a.job_cb[0] = 1; a.nsplit = g.nsplit;
    a.Nb[0] = fp.N1; a.Nb[1] = fp.N2; a.startb[0] = fp.start1; a.startb[1] = fp.start2;
    a.dnoff[0] = 0; a.dnoff[1] = g.N1pad;
    const bool shared_dict = fp.src && fp.strideA == 0;
    a.dn_stride = shared_dict ? 0 : (int64_t)g.Mp2 * g.ldn;
    a.debug = 0;
#ifdef MFB_EXPERIMENTS
    if (const char *d = getenv("MFB_FAST_DEBUG")) a.debug = atoi(d);
#endif
    a.vox_list = vox_list; a.peaks = peaks; a.peaks_ld = peaks_ld; a.y = y;
    char *q = (char *)scratch;
    if (!fp.src) {
        a.ip_rows = (int *)q; q += al256(sizeof(int) * V * 2 * p.M * 2);
        a.ip_w = (double *)q; q += al256(sizeof(double) * V * 2 * p.M * 2);
    }
    a.colp = (double *)q; q += al256(sizeof(double) * V * 2 * FT_NPAR * a.Npad);
    a.voxp = (double *)q; q += al256(sizeof(double) * V * FT_VP);
    a.vthr = (unsigned long long *)q; q += al256(sizeof(unsigned long long) * V);
    a.cta_gain = (double *)q; q += al256(sizeof(double) * V * a.ntI);
    a.cta_tol = (double *)q; q += al256(sizeof(double) * V * a.ntI);
    a.cta_ill = (double *)q; q += al256(sizeof(double) * V * a.ntI);
    a.cta_idx = (int *)q; q += al256(sizeof(int) * V * a.ntI);
    a.cta_flag = (int *)q; q += al256(sizeof(int) * V * a.ntI);
    a.Dn = g.gemm ? (double *)q : nullptr;
    a.nt2 = (fp.N2 + FT_TJ - 1) / FT_TJ;
    if (!g.gemm) {
        a.d2c_stride = (int64_t)a.nt2 * (int64_t)FT2_REC(g.Mp);
        a.D2c = (double *)q; q += al256(sizeof(double) * V * (size_t)a.d2c_stride);
    }
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
        b.colp += v0 * 2 * FT_NPAR * a.Npad; b.voxp += v0 * FT_VP; b.vthr += v0;
        b.cta_gain += v0 * a.ntI; b.cta_tol += v0 * a.ntI; b.cta_ill += v0 * a.ntI;
        b.cta_idx += v0 * a.ntI; b.cta_flag += v0 * a.ntI;
        if (b.Dn) b.Dn += v0 * a.dn_stride;
        if (b.D2c) b.D2c += v0 * a.d2c_stride;
        if (vox_list) b.vox_list = vox_list + v0;
        else { b.y = y + v0 * p.M; b.tuple = tuple + v0; }
        if (fp.A && (!vox_list || fp.a_by_local)) b.A = fp.A + v0 * fp.strideA;
        return b;
    };
    if (g.gemm) {
        const size_t smem = sizeof(double) * ((size_t)GP_NS * GP_STAGE + 8 * 5 * GP_TJ + 64);
        void (*kern)(FastArgs) = fp.csf ? k_gemm_pairs<1, 0> : k_gemm_pairs<0, 0>;
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
        const size_t smem = sizeof(double) * ((size_t)a.Mp * FT_S1 + (size_t)FT_NS * FT2_REC(a.Mp) + 3 * a.Mp + 64) +
                            sizeof(int) * 2 * a.Mp;
        if (smem + 128 > 227 * 1024) {
            set_error("fast tier: tile does not fit in shared memory");
            return MFB_EUNSUPPORTED;
        }
        // per device / context attribute: set on every launch (microseconds)
        void (*kern)(FastArgs) = fp.csf ? (fp.src ? k_fast_tiles<1, 1> : k_fast_tiles<1, 0>)
                                        : (fp.src ? k_fast_tiles<0, 1> : k_fast_tiles<0, 0>);
        MFB_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        MFB_LAUNCH(k_fast_seed, (unsigned)V, 32, 0, st, a);
        if (ev) MFB_CUDA_TRY(cudaEventRecord(ev[0], st));
        for (int64_t v0 = 0; v0 < V; v0 += maxy) {
            const int64_t nv = V - v0 < maxy ? V - v0 : maxy;
            MFB_LAUNCH(kern, dim3(a.ntI, (unsigned)nv), FT2_THREADS, smem, st, shifted(v0));
        }
        if (ev) MFB_CUDA_TRY(cudaEventRecord(ev[1], st));
    }
    MFB_LAUNCH(k_fast_select, (unsigned)((V + 127) / 128), 128, 0, st, a, V);
    return MFB_OK;
}


// ------------------------------ triple scan, host side ------------------------------
bool fast3_supported_explicit(int M, const BlockSpec &bs)
{
    if (bs.nb != 3 || M > 16384) return false;
    for (int b = 0; b < 3; b++)
        if (bs.size[b] < 2 || bs.size[b] > 4096) return false;
    return true;
}

// thread layout (txt x tyt threads, 2 x 4 pairs each) wasting the fewest lanes and tile slots.
// cost: CTAs x (steps over the streamed block + the fixed cost of a CTA -- first chunk, pair
// constants, reduction -- expressed in steps) x lanes x tile slots per useful pair
#define TR_CTA_OVERHEAD_STEPS 40
static TripleGeom triple_geom1(int N1, int N2, int N3, int maxthr, double *cost_out)
{
    TripleGeom best;
    memset(&best, 0, sizeof(best));
    double best_cost = 1e300;
    for (int txt = 4; txt <= 64; txt++)
        for (int tyt = 2; tyt <= 48; tyt++) {
            const int nthr = txt * tyt;
            if (nthr < maxthr / 2 || nthr > maxthr) continue;
            const int T1 = 2 * txt, T2 = 4 * tyt;
            const int nt1 = (N1 + T1 - 1) / T1, nt2 = (N2 + T2 - 1) / T2;
            const int threads = (nthr + 31) / 32 * 32;
            // mild preference for full CTAs
            const double cost = (double)nt1 * T1 * nt2 * T2 * threads / nthr * (1.0 + 0.02 * (maxthr - threads) / 32) *
                                (((N3 + 3) & ~3) + TR_CTA_OVERHEAD_STEPS);
            if (cost < best_cost) {
                best_cost = cost;
                best.txt = txt; best.tyt = tyt; best.T1 = T1; best.T2 = T2; best.nt1 = nt1; best.nt2 = nt2;
                best.threads = threads;
            }
        }
    if (cost_out) *cost_out = best_cost;
    return best;
}

// one CTA per SM, or two when one cannot use more than 256 threads
static TripleGeom triple_geom(int N1, int N2, int N3, double *cost_out, int force_ctas = 0)
{
    TripleGeom g = triple_geom1(N1, N2, N3, TR_MAXTHREADS, cost_out);
    g.ctas = 1;
    if (g.threads <= 256 || force_ctas > 1) {
        const int nc = force_ctas > 1 ? force_ctas : 2;
        double c2;
        TripleGeom g2 = triple_geom1(N1, N2, N3, TR_MAXTHREADS / nc, &c2);
        if (g2.threads > 0) { g = g2; g.ctas = nc; if (cost_out) *cost_out = c2; }
    }
    return g;
}

// The scan tiles two blocks over the CTAs and streams the third: with a short third block (the
// EAR compartment of MFModel.fit: ~10 atoms against 1000 per fascicle) a CTA would spend its life
// in its prologue.  The blocks are therefore permuted so that the scan is cheapest by the cost
// model above; k_select3 maps the winner back to the caller's loop order.
static BlockSpec triple_permute(const BlockSpec &bs, int csf, int perm[3], TripleGeom *tg)
{
    static const int P[6][3] = {{0, 1, 2}, {1, 0, 2}, {0, 2, 1}, {2, 0, 1}, {1, 2, 0}, {2, 1, 0}};
    double best_cost = 1e300;
    int bp = 0;
    TripleGeom bg;
    memset(&bg, 0, sizeof(bg));
    // (CSF-projected scan: the caller's order is kept -- its competitive path, not the stream length,
    // sets the pace, and measured 10-20 % slower with the short block tiled)
    for (int k = 0; k < (csf ? 1 : 6); k++) {
        double c;
        const TripleGeom g = triple_geom(bs.size[P[k][0]], bs.size[P[k][1]], bs.size[P[k][2]], &c, csf ? TR_CSF_CTAS : 0);
        if (c < best_cost * (1.0 - 1e-9)) { best_cost = c; bp = k; bg = g; }   // ties keep the caller's order
    }
    BlockSpec out = bs;
    for (int k = 0; k < 3; k++) { perm[k] = P[bp][k]; out.size[k] = bs.size[perm[k]]; out.start[k] = bs.start[perm[k]]; }
    if (tg) *tg = bg;
    return out;
}

struct Fast3Layout {
    int Mp2, Npad, Np[3], ldn, ntI;
    TripleGeom tg;
    size_t off_colp, off_voxp, off_gain, off_tol, off_ill, off_idx, off_flag, off_dn, off_r[3];
    size_t off_tgain, off_ttol, off_till, off_tidx, off_tflag, off_vthr, total;
    size_t r_elems[3];
};

static Fast3Layout fast3_layout(int M, const BlockSpec &bs_caller, int64_t V, int shared_dict, int csf)
{
    Fast3Layout L;
    int perm[3];
    const BlockSpec bs = triple_permute(bs_caller, csf, perm, &L.tg);
    L.Mp2 = (M + GP_KC - 1) / GP_KC * GP_KC;
    int nmax = 0;
    for (int b = 0; b < 3; b++) { L.Np[b] = (bs.size[b] + GP_TI - 1) / GP_TI * GP_TI; nmax = nmax > L.Np[b] ? nmax : L.Np[b]; }
    L.Npad = nmax;
    L.ldn = L.Np[0] + L.Np[1] + L.Np[2];
    L.ntI = nmax / GP_TI;
    const int64_t Vd = shared_dict ? 1 : V;
    const size_t ntile = (size_t)L.tg.nt1 * L.tg.nt2;
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t r = o; o += al256(bytes); return r; };
    L.off_colp = take(sizeof(double) * V * 3 * FT_NPAR * L.Npad);
    L.off_voxp = take(sizeof(double) * V * FT_VP);
    L.off_gain = take(sizeof(double) * V * 3 * L.ntI);
    L.off_tol = take(sizeof(double) * V * 3 * L.ntI);
    L.off_ill = take(sizeof(double) * V * 3 * L.ntI);
    L.off_idx = take(sizeof(int) * V * 3 * L.ntI);
    L.off_flag = take(sizeof(int) * V * 3 * L.ntI);
    L.off_dn = take(sizeof(double) * Vd * (size_t)L.Mp2 * L.ldn);
    L.r_elems[0] = (size_t)L.Np[0] * L.Np[1];   // R12   [N1p][N2p]
    L.r_elems[1] = (size_t)L.Np[2] * L.Np[0];   // R13^T [N3p][N1p]
    L.r_elems[2] = (size_t)L.Np[2] * L.Np[1];   // R23^T [N3p][N2p]
    for (int j = 0; j < 3; j++) L.off_r[j] = take(sizeof(double) * Vd * L.r_elems[j]);
    L.off_tgain = take(sizeof(double) * V * ntile);
    L.off_ttol = take(sizeof(double) * V * ntile);
    L.off_till = take(sizeof(double) * V * ntile);
    L.off_tidx = take(sizeof(long long) * V * ntile);
    L.off_tflag = take(sizeof(int) * V * ntile);
    L.off_vthr = take(sizeof(unsigned long long) * V);
    L.total = o;
    return L;
}

size_t fast3_scratch_bytes(int M, const BlockSpec &bs, int64_t V, int shared_dict)
{
    // (the CSF-projected scan keeps the caller's block order: a different tiling, never a larger R)
    const size_t t0 = fast3_layout(M, bs, V, shared_dict, 0).total, t1 = fast3_layout(M, bs, V, shared_dict, 1).total;
    return t0 > t1 ? t0 : t1;
}

int launch_fast_search3(int M, const BlockSpec &bs_caller, const double *A, int64_t lda, int64_t strideA,
                        int64_t V, const double *y, void *scratch, long long *tuple,
                        int32_t *redo_list, int32_t *redo_count, int32_t *reasons, cudaStream_t st,
                        cudaEvent_t *ev, const int32_t *vox_list, int a_by_local, int32_t *redo_local, int csf_col)
{
    if (V == 0) return MFB_OK;
    if (V > 65535) { set_error("triple scan: at most 65535 voxels per launch"); return MFB_EINVAL; }
    const int shared_dict = strideA == 0;
    const Fast3Layout L = fast3_layout(M, bs_caller, V, shared_dict, csf_col >= 0);
    FastArgs a;
    memset(&a, 0, sizeof(a));
    const BlockSpec bs = triple_permute(bs_caller, csf_col >= 0, a.tr_perm, nullptr);
    a.p.M = M; a.src = 1;
    a.csf = csf_col >= 0 ? 1 : 0;        // the three searched blocks are projected off this column
    a.start3 = csf_col >= 0 ? csf_col : 0;
    a.A = A; a.lda = lda; a.strideA = strideA;
    a.nblk = 3; a.njobs = 3; a.nsplit = 1;
    int off = 0;
    for (int b = 0; b < 3; b++) { a.Nb[b] = bs.size[b]; a.startb[b] = bs.start[b]; a.dnoff[b] = off; off += L.Np[b]; }
    a.N1 = bs.size[0]; a.N2 = bs.size[1];
    a.start1 = bs.start[0]; a.start2 = bs.start[1];
    a.Mp = (M + 3) & ~3; a.Mp2 = L.Mp2; a.Npad = L.Npad; a.ntI = L.ntI; a.ldn = L.ldn; a.N1pad = L.Np[0];
    a.dn_stride = shared_dict ? 0 : (int64_t)L.Mp2 * L.ldn;
    a.job_rb[0] = 0; a.job_cb[0] = 1; a.ldr[0] = L.Np[1];
    a.job_rb[1] = 2; a.job_cb[1] = 0; a.ldr[1] = L.Np[0];
    a.job_rb[2] = 2; a.job_cb[2] = 1; a.ldr[2] = L.Np[1];
    char *q = (char *)scratch;
    a.colp = (double *)(q + L.off_colp); a.voxp = (double *)(q + L.off_voxp);
    a.cta_gain = (double *)(q + L.off_gain); a.cta_tol = (double *)(q + L.off_tol);
    a.cta_ill = (double *)(q + L.off_ill); a.cta_idx = (int *)(q + L.off_idx); a.cta_flag = (int *)(q + L.off_flag);
    a.Dn = (double *)(q + L.off_dn);
    for (int j = 0; j < 3; j++) { a.R[j] = (double *)(q + L.off_r[j]); a.r_stride[j] = shared_dict ? 0 : (int64_t)L.r_elems[j]; }
    a.t_gain = (double *)(q + L.off_tgain); a.t_tol = (double *)(q + L.off_ttol); a.t_ill = (double *)(q + L.off_till);
    a.t_idx = (long long *)(q + L.off_tidx); a.t_flag = (int *)(q + L.off_tflag);
    a.vthr = (unsigned long long *)(q + L.off_vthr);
    a.tr_txt = L.tg.txt; a.tr_tyt = L.tg.tyt; a.tr_nt1 = L.tg.nt1; a.tr_ntiles = L.tg.nt1 * L.tg.nt2;
    a.y = y; a.tuple = tuple; a.redo_list = redo_list; a.redo_count = redo_count; a.reasons = reasons;
    a.vox_list = vox_list; a.a_by_local = a_by_local; a.redo_local = redo_local;

    MFB_CUDA_TRY(cudaMemsetAsync(a.vthr, 0, sizeof(unsigned long long) * V, st));
    const size_t smem_prep = sizeof(double) * (4 * M + 32) + sizeof(int) * 2 * M;
    if (smem_prep > 200 * 1024) { set_error("triple scan: too many measurements"); return MFB_EUNSUPPORTED; }
    if (smem_prep > 48 * 1024)
        MFB_CUDA_TRY(cudaFuncSetAttribute(k_fast_prep, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_prep));
    MFB_LAUNCH(k_fast_prep, dim3((unsigned)V, 3), 256, smem_prep, st, a);
    const unsigned zc = (unsigned)((L.Mp2 + GP_NROWCHUNK - 1) / GP_NROWCHUNK);
    MFB_LAUNCH(k_normalize, dim3((unsigned)(shared_dict ? 1 : V), 3, zc), 256, 0, st, a);
    if (ev) MFB_CUDA_TRY(cudaEventRecord(ev[0], st));
    {
        const size_t smem = sizeof(double) * ((size_t)GP_NS * GP_STAGE + 8 * 5 * GP_TJ + 64);
        void (*kern)(FastArgs) = a.csf ? k_gemm_pairs<1, 1> : k_gemm_pairs<0, 1>;
        MFB_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        MFB_LAUNCH(kern, dim3(a.ntI, (unsigned)V, 3), GP_THREADS, smem, st, a);
    }
    {
        const int rowlen = L.tg.T1 + L.tg.T2;
        const size_t fixed = sizeof(double) * (((bs.size[2] + 3) & ~3) + 64);
        // ctas CTAs per SM: each gets its share of the 227 KB (1 KB per CTA is reserved)
        int kc = (int)(((size_t)(224 / L.tg.ctas - 2) * 1024 - fixed) / (sizeof(double) * 2 * rowlen)) & ~3;
        kc = kc > TR_KC_MAX ? TR_KC_MAX : kc;
        if (kc < 4) { set_error("triple scan: third block too large for shared memory"); return MFB_EUNSUPPORTED; }
        a.tr_kc = kc;
        const size_t smem = sizeof(double) * (size_t)2 * kc * rowlen + fixed;
        MFB_LAUNCH(k_triple_seed, (unsigned)V, 128, 0, st, a);
        void (*ktr)(FastArgs) = L.tg.ctas == 2 ? (a.csf ? k_triples<1, 2> : k_triples<0, 2>)
                                : L.tg.ctas == 3 ? k_triples<1, 3>
                                               : (a.csf ? k_triples<1, 1> : k_triples<0, 1>);
        MFB_CUDA_TRY(cudaFuncSetAttribute(ktr, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        MFB_LAUNCH(ktr, dim3((unsigned)a.tr_ntiles, (unsigned)V), L.tg.threads, smem, st, a);
    }
    if (ev) MFB_CUDA_TRY(cudaEventRecord(ev[1], st));
    MFB_LAUNCH(k_select3, (unsigned)((V + 127) / 128), 128, 0, st, a, V);
#ifdef TR_COUNT
    {
        unsigned long long h[4];
        cudaStreamSynchronize(st);
        cudaMemcpyFromSymbol(h, g_tr_counters, sizeof(h));
        fprintf(stderr, "k_triples votes %llu, hits %llu (%.4f), CSF: tuples passing the gain test %llu (%.2f per hit), competitive %llu\n", h[0], h[2],
                (double)h[2] / (double)(h[0] ? h[0] : 1), h[1], (double)h[1] / (double)(h[2] ? h[2] : 1), h[3]);
        memset(h, 0, sizeof(h));
        cudaMemcpyToSymbol(g_tr_counters, h, sizeof(h));
    }
#endif
    return MFB_OK;
}

}  // namespace mfb
