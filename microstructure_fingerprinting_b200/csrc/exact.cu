// exact.cu -- the reference-order ("exact") tier of libmfb200, sm_100a.
//
// Every floating-point operation that decides a result goes through __dmul_rn /
// __dadd_rn / __dsub_rn / __ddiv_rn (never contracted into FMA; the file is also built
// with -fmad=false), in the summation order of the reference's Numba kernels, so that
// weights, indices and objectives are bit-identical to a strict IEEE evaluation of
//   mfu.solve_exhaustive_posweights_1/_2/_3 (mf_utils.py:225-278, 288-392, 470-607),
//   mfu.lsqnonneg_2var_opt (mf_utils.py:404-459) and
//   mfu.interp_PGSE_from_multishell, fast mode (mf_utils.py:1693-1956, with the
//   scipy interp1d._call_linear two-weight lerp).
// Loop-order tie-breaks ("first strict minimum") are reproduced by reducing on the pair
// (residual, loop index).
#include <algorithm>
#include <climits>

#include "common.cuh"

namespace mfb {

#define DM(a, b) __dmul_rn((a), (b))
#define DA(a, b) __dadd_rn((a), (b))
#define DS(a, b) __dsub_rn((a), (b))
#define DD(a, b) __ddiv_rn((a), (b))

// ---------------------------------------------------------------------------------
// classification
// ---------------------------------------------------------------------------------
__global__ void k_classify(int64_t V, const int32_t *K, const uint8_t *csf, const uint8_t *ear,
                           int maxfasc, uint8_t *type, uint8_t *nbv, int32_t *lists,
                           int32_t *counts)
{
    int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= V) return;
    int k = K[v];
    k = k < 0 ? 0 : (k > maxfasc ? maxfasc : k);
    int c = csf ? (csf[v] != 0) : 0;
    int e = ear ? (ear[v] != 0) : 0;
    int t = k + 3 * c + 6 * e;
    type[v] = (uint8_t)t;
    nbv[v] = (uint8_t)(k + c + e);
    int pos = atomicAdd(&counts[t], 1);
    lists[(int64_t)t * V + pos] = (int32_t)v;
}

int launch_classify(int64_t V, const int32_t *K, const uint8_t *csf, const uint8_t *ear,
                    int maxfasc, uint8_t *type, uint8_t *nbv, int32_t *lists, int32_t *counts,
                    cudaStream_t st)
{
    MFB_CUDA_TRY(cudaMemsetAsync(counts, 0, 12 * sizeof(int32_t), st));
    MFB_LAUNCH(k_classify, (unsigned)((V + 255) / 256), 256, 0, st, V, K, csf, ear, maxfasc,
               type, nbv, lists, counts);
    return MFB_OK;
}

// ---------------------------------------------------------------------------------
// rotation (interp_PGSE_from_multishell, fast mode)
// ---------------------------------------------------------------------------------
struct Lerp {
    int rl, rh;
    double wl, wh;
};

__device__ __forceinline__ Lerp shell_lerp(const DevPlan &p, int s, double x)
{
    const double *xs = p.nodes + p.off[s];
    int n = p.off[s + 1] - p.off[s];
    int lo = 0, hi = n;  // searchsorted(side='left')
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (xs[mid] < x) lo = mid + 1; else hi = mid;
    }
    int j = lo < 1 ? 1 : (lo > n - 1 ? n - 1 : lo);  // clip(1, n-1)
    double x_lo = xs[j - 1], x_hi = xs[j];
    double den = DS(x_hi, x_lo);
    Lerp L;
    L.rl = p.off[s] + j - 1;
    L.rh = p.off[s] + j;
    L.wh = DD(DS(x, x_lo), den);
    L.wl = DD(DS(x_hi, x), den);
    return L;
}

__device__ __forceinline__ double dir_dot(const DevPlan &p, int m, double ux, double uy,
                                          double uz)
{
    // mfu:1810  x = |g . newdir|, left-to-right, separately rounded
    double d = DA(DA(DM(p.gdir[3 * m], ux), DM(p.gdir[3 * m + 1], uy)),
                  DM(p.gdir[3 * m + 2], uz));
    return fabs(d);
}

// One rotated entry from the lookup table (exact-G or between-shell).
__device__ __forceinline__ double rot_entry(const DevPlan &p, const Lerp &a, const Lerp &b,
                                            bool between, double gwl, double gwh, int j)
{
    const double *T = p.table;
    double d_l = DA(DM(a.wh, T[(size_t)a.rh * p.N + j]), DM(a.wl, T[(size_t)a.rl * p.N + j]));
    if (!between) return d_l;
    double d_h = DA(DM(b.wh, T[(size_t)b.rh * p.N + j]), DM(b.wl, T[(size_t)b.rl * p.N + j]));
    return DA(DM(gwh, d_h), DM(gwl, d_l));
}

// grid (nvox, K + (csf|ear)); block 256.  Dynamic smem: per-measurement interpolation plan.
__global__ void __launch_bounds__(256)
k_rotate_assemble(DevPlan p, const int32_t *vox_list, const double *peaks, int peaks_ld, int K,
                  int csf, int ear, double *A, int64_t lda, int64_t strideA)
{
    extern __shared__ double sm[];
    const int M = p.M, N = p.N;
    double *wl = sm, *wh = sm + M, *wl2 = sm + 2 * M, *wh2 = sm + 3 * M;
    int *rl = (int *)(sm + 4 * M), *rh = rl + M, *rl2 = rh + M, *rh2 = rl2 + M;
    const int64_t v = blockIdx.x;
    const int64_t row = vox_list ? vox_list[v] : v;
    double *Av = A + v * strideA;
    const int k = blockIdx.y;
    if (k >= K) {  // iso columns (mf:401-408)
        for (int e = threadIdx.x; e < M * (csf + ear * p.E); e += blockDim.x) {
            int m = e / (csf + ear * p.E), c = e % (csf + ear * p.E);
            double val = (csf && c == 0) ? p.sig_csf[m] : p.sig_ear[(size_t)m * p.E + (c - csf)];
            Av[(size_t)m * lda + (size_t)K * N + c] = val;
        }
        return;
    }
    const double ux = peaks[row * peaks_ld + 3 * k], uy = peaks[row * peaks_ld + 3 * k + 1],
                 uz = peaks[row * peaks_ld + 3 * k + 2];
    for (int m = threadIdx.x; m < M; m += blockDim.x) {
        double x = dir_dot(p, m, ux, uy, uz);
        Lerp a = shell_lerp(p, p.shell_lo[m], x);
        rl[m] = a.rl; rh[m] = a.rh; wl[m] = a.wl; wh[m] = a.wh;
        if (p.shell_hi[m] != p.shell_lo[m]) {
            Lerp b = shell_lerp(p, p.shell_hi[m], x);
            rl2[m] = b.rl; rh2[m] = b.rh; wl2[m] = b.wl; wh2[m] = b.wh;
        } else {
            rl2[m] = -1;
        }
    }
    __syncthreads();
    double *out = Av + (size_t)k * N;
    // one warp per measurement row; lanes stream the row with 16-byte accesses when the row
    // starts of the table and of the output are 16-byte aligned (N, lda, k*N even)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    const bool vec = (N % 2 == 0) && (lda % 2 == 0) && ((((size_t)k * N) & 1) == 0) &&
                     ((reinterpret_cast<size_t>(p.table) & 15) == 0) && ((reinterpret_cast<size_t>(Av) & 15) == 0);
    for (int m = warp; m < M; m += nwarp) {
        Lerp a, b;
        a.rl = rl[m]; a.rh = rh[m]; a.wl = wl[m]; a.wh = wh[m];
        const bool between = rl2[m] >= 0;
        double gwl = 0.0, gwh = 0.0;
        b = a;
        if (between) {
            b.rl = rl2[m]; b.rh = rh2[m]; b.wl = wl2[m]; b.wh = wh2[m];
            gwl = p.gw_lo[m]; gwh = p.gw_hi[m];
        }
        double *orow = out + (size_t)m * lda;
        if (vec) {
            const double2 *Tl = reinterpret_cast<const double2 *>(p.table + (size_t)a.rl * N);
            const double2 *Th = reinterpret_cast<const double2 *>(p.table + (size_t)a.rh * N);
            const double2 *Tl2 = reinterpret_cast<const double2 *>(p.table + (size_t)b.rl * N);
            const double2 *Th2 = reinterpret_cast<const double2 *>(p.table + (size_t)b.rh * N);
            double2 *o2 = reinterpret_cast<double2 *>(orow);
            const int n2 = N >> 1;
#pragma unroll 4
            for (int j = lane; j < n2; j += 32) {
                const double2 lo = __ldg(Tl + j), hi = __ldg(Th + j);
                double2 d;
                d.x = DA(DM(a.wh, hi.x), DM(a.wl, lo.x));
                d.y = DA(DM(a.wh, hi.y), DM(a.wl, lo.y));
                if (between) {
                    const double2 lo2 = __ldg(Tl2 + j), hi2 = __ldg(Th2 + j);
                    const double hx = DA(DM(b.wh, hi2.x), DM(b.wl, lo2.x));
                    const double hy = DA(DM(b.wh, hi2.y), DM(b.wl, lo2.y));
                    d.x = DA(DM(gwh, hx), DM(gwl, d.x));
                    d.y = DA(DM(gwh, hy), DM(gwl, d.y));
                }
                o2[j] = d;
            }
        } else {
            for (int j = lane; j < N; j += 32) orow[j] = rot_entry(p, a, b, between, gwl, gwh, j);
        }
    }
}

int launch_rotate_assemble(const DevPlan &p, int64_t nvox, const int32_t *vox_list,
                           const double *peaks, int peaks_ld, int K, int csf, int ear,
                           double *A, int64_t lda, int64_t strideA, cudaStream_t st)
{
    if (nvox == 0) return MFB_OK;
    size_t smem = (size_t)p.M * (4 * sizeof(double) + 4 * sizeof(int));
    const int64_t maxgrid = 1 << 20;
    for (int64_t v0 = 0; v0 < nvox; v0 += maxgrid) {
        int64_t nv = nvox - v0 < maxgrid ? nvox - v0 : maxgrid;
        dim3 grid((unsigned)nv, (unsigned)(K + ((csf || ear) ? 1 : 0)));
        MFB_LAUNCH(k_rotate_assemble, grid, 256, smem, st, p, vox_list ? vox_list + v0 : nullptr,
                   vox_list ? peaks : peaks + v0 * peaks_ld, peaks_ld, K, csf, ear,
                   A + v0 * strideA, lda, strideA);
    }
    return MFB_OK;
}

// ---------------------------------------------------------------------------------
// explicit-plan row lerp (rotate_atom / rotate_atom_2Dprotocol): grid (M, V)
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_lerp_rows(int M, int N, const double *table, const int32_t *row_lo, const int32_t *row_hi,
            const double *w_lo, const double *w_hi, const double *scale, double *out, int64_t ldd)
{
    const int64_t v = blockIdx.y;
    const int m = blockIdx.x;
    const int64_t o = v * M + m;
    const double *tl = table + (size_t)row_lo[o] * N, *th = table + (size_t)row_hi[o] * N;
    const double wl = w_lo[o], wh = w_hi[o];
    double *dst = out + (v * M + m) * ldd;
    const double sc = scale ? scale[o] : 1.0;
    const bool vec = (N % 2 == 0) && (ldd % 2 == 0) && ((reinterpret_cast<size_t>(table) & 15) == 0) &&
                     ((reinterpret_cast<size_t>(out) & 15) == 0);
    if (vec) {      // 16-byte accesses
        const double2 *tl2 = reinterpret_cast<const double2 *>(tl), *th2 = reinterpret_cast<const double2 *>(th);
        double2 *d2 = reinterpret_cast<double2 *>(dst);
        for (int j = threadIdx.x; j < (N >> 1); j += blockDim.x) {
            const double2 lo = __ldg(tl2 + j), hi = __ldg(th2 + j);
            double2 r;
            r.x = DA(DM(wh, hi.x), DM(wl, lo.x));
            r.y = DA(DM(wh, hi.y), DM(wl, lo.y));
            if (scale) { r.x = DM(sc, r.x); r.y = DM(sc, r.y); }
            d2[j] = r;
        }
    } else if (scale) {
        for (int j = threadIdx.x; j < N; j += blockDim.x) dst[j] = DM(sc, DA(DM(wh, th[j]), DM(wl, tl[j])));
    } else {
        for (int j = threadIdx.x; j < N; j += blockDim.x) dst[j] = DA(DM(wh, th[j]), DM(wl, tl[j]));
    }
}

int launch_lerp_rows(int64_t V, int M, int N, const double *table, const int32_t *row_lo,
                     const int32_t *row_hi, const double *w_lo, const double *w_hi,
                     const double *scale, double *out, int64_t ldd, cudaStream_t st)
{
    if (V == 0 || M == 0) return MFB_OK;
    for (int64_t v0 = 0; v0 < V; v0 += 65535) {
        const int64_t nv = std::min<int64_t>(65535, V - v0);
        MFB_LAUNCH(k_lerp_rows, dim3((unsigned)M, (unsigned)nv), 256, 0, st, M, N, table, row_lo + v0 * M,
                   row_hi + v0 * M, w_lo + v0 * M, w_hi + v0 * M, scale ? scale + v0 * M : nullptr,
                   out + v0 * M * ldd, ldd);
    }
    return MFB_OK;
}

// ---------------------------------------------------------------------------------
// rotate_atom_2Dprotocol (mfu:1440-1690): expansion of the per-direction, per-class decisions of
// the host (which gradient line of the reference a class of sequences interpolates along, its sign,
// vanished perpendicular gradients) into the per-sequence interpolation plan that k_lerp_rows
// consumes.  One thread per (direction, sequence); the arithmetic is the host's, operation by
// operation (this file is compiled without FMA contraction): G_perp = G nrm, G_par = |g_z| G,
// S_par = exp(-(gamma delta G_par)^2 (Delta - delta/3) D), scipy's two-weight lerp in the signed
// perpendicular gradient with the interval index clipped to [1, n-1].
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_plan2d(int64_t total, int M, int U, int C, const int32_t *m_class, const int32_t *m_lab, const uint8_t *m_isb0,
         const int32_t *m_b0row, const double *m_G, const double *m_gd, const double *m_tt, double DIFF,
         const double *nrm, const double *gz, const uint8_t *kind, const int32_t *line, const double *sgn,
         const uint8_t *ok, const int32_t *line_off, const double *line_nodes, const int32_t *line_rows,
         int32_t *row_lo, int32_t *row_hi, double *w_lo, double *w_hi, double *scale)
{
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int64_t v = idx / M;
    const int m = (int)(idx - v * M);
    const int u = m_lab[m];
    const double G = m_G[m];
    const double a = DM(m_gd[m], DM(gz[v * U + u], G));
    const double S_par = exp(DM(DM(-DM(a, a), m_tt[m]), DIFF));
    int rl = m, rh = m;
    double wl = 1.0, wh = 0.0;
    bool covered = m_isb0[m] != 0;
    if (!covered) {
        const int c = m_class[m];
        const int k = kind[v * C + c];
        if (k == 2) {                       // gradient parallel to the new fascicle: mean b0 signal of the shell
            rl = rh = m_b0row[m];
            covered = true;
        } else if (k == 3) {                // interpolate along the closest reference line
            const int li = line[v * C + c];
            const int o = line_off[li], n = line_off[li + 1] - o;
            const double *xs = line_nodes + o;
            const double x = DM(DM(G, nrm[v * U + u]), sgn[v * C + c]);
            int lo = 0, hi = n;             // np.searchsorted(xs, x, side='left')
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (xs[mid] < x) lo = mid + 1; else hi = mid;
            }
            const int j = max(1, min(n - 1, lo));
            const double x_lo = xs[j - 1], x_hi = xs[j];
            wl = DD(DS(x_hi, x), DS(x_hi, x_lo));
            wh = DD(DS(x, x_lo), DS(x_hi, x_lo));
            rl = line_rows[o + j - 1];
            rh = line_rows[o + j];
            covered = true;
        }
    }
    row_lo[idx] = rl; row_hi[idx] = rh; w_lo[idx] = wl; w_hi[idx] = wh;
    scale[idx] = (covered && ok[v]) ? S_par : 0.0;
}

int launch_plan2d(int64_t V, int M, int U, int C, const int32_t *m_class, const int32_t *m_lab,
                  const uint8_t *m_isb0, const int32_t *m_b0row, const double *m_G, const double *m_gd,
                  const double *m_tt, double DIFF, const double *nrm, const double *gz, const uint8_t *kind,
                  const int32_t *line, const double *sgn, const uint8_t *ok, const int32_t *line_off,
                  const double *line_nodes, const int32_t *line_rows, int32_t *row_lo, int32_t *row_hi,
                  double *w_lo, double *w_hi, double *scale, cudaStream_t st)
{
    const int64_t total = V * M;
    if (total == 0) return MFB_OK;
    MFB_LAUNCH(k_plan2d, (unsigned)((total + 255) / 256), 256, 0, st, total, M, U, C, m_class, m_lab, m_isb0, m_b0row,
               m_G, m_gd, m_tt, DIFF, nrm, gz, kind, line, sgn, ok, line_off, line_nodes, line_rows, row_lo, row_hi,
               w_lo, w_hi, scale);
    return MFB_OK;
}

// ---------------------------------------------------------------------------------
// closed forms
// ---------------------------------------------------------------------------------
// mfu:404-459 lsqnonneg_2var_opt (also inlined at mfu:331-381)
__device__ __forceinline__ double lsq2(double y_sq, double A11, double A12, double A22,
                                       double Y1, double Y2, double &w0, double &w1)
{
    double w1d = DS(DM(A22, Y1), DM(A12, Y2));
    double w2d = DS(DM(A11, Y2), DM(A12, Y1));
    double res = y_sq;
    w0 = 0.0; w1 = 0.0;
    if (w1d > 0.0 && w2d > 0.0) {
        double Det = DS(DM(A11, A22), DM(A12, A12));
        w0 = DD(w1d, Det);
        w1 = DD(w2d, Det);
        double t1 = DA(DA(res, DM(DM(w0, w0), A11)), DM(DM(w1, w1), A22));
        double t2 = DS(DS(DM(DM(w0, w1), A12), DM(w0, Y1)), DM(w1, Y2));
        res = DA(t1, DM(2.0, t2));
    } else if (w1d >= 0.0 && w2d <= 0.0) {
        if (Y1 >= 0.0) { w0 = DD(Y1, A11); res = DS(res, DM(Y1, w0)); }
    } else if (w1d <= 0.0 && w2d >= 0.0) {
        if (Y2 >= 0.0) { w1 = DD(Y2, A22); res = DS(res, DM(Y2, w1)); }
    } else if (w1d < 0.0 && w2d < 0.0) {
        if (Y1 > 0.0) { w0 = DD(Y1, A11); res = DS(res, DM(Y1, w0)); }
        else if (Y2 > 0.0) { w1 = DD(Y2, A22); res = DS(res, DM(Y2, w1)); }
    }
    return res;
}

// mfu:556-567 Cramer numerators and determinant; returns true when all three numerators
// are >= -tol (the branch whose residual the reference evaluates directly, mfu:562-573).
__device__ __forceinline__ bool cramer3(double a11, double a12, double a13, double a22,
                                        double a23, double a33, double Y1, double Y2, double Y3,
                                        double &w0, double &w1, double &w2)
{
    const double tol = 100 * 2.2204e-16;  // mfu:480-481
    double m2233 = DS(DM(a22, a33), DM(a23, a23));
    double m1233 = DS(DM(a12, a33), DM(a23, a13));
    double m1223 = DS(DM(a12, a23), DM(a22, a13));
    double D1 = DA(DS(DM(Y1, m2233), DM(Y2, m1233)), DM(Y3, m1223));
    double n1233 = DS(DM(a12, a33), DM(a13, a23));
    double m1133 = DS(DM(a11, a33), DM(a13, a13));
    double m1123 = DS(DM(a11, a23), DM(a12, a13));
    double D2 = DS(DA(DM(-Y1, n1233), DM(Y2, m1133)), DM(Y3, m1123));
    double n1223 = DS(DM(a12, a23), DM(a13, a22));
    double m1122 = DS(DM(a11, a22), DM(a12, a12));
    double D3 = DA(DS(DM(Y1, n1223), DM(Y2, m1123)), DM(Y3, m1122));
    if (D1 >= -tol && D2 >= -tol && D3 >= -tol) {
        double D = DA(DS(DM(a11, m2233), DM(a12, m1233)), DM(a13, m1223));
        w0 = DD(D1, D); w1 = DD(D2, D); w2 = DD(D3, D);
        return true;
    }
    return false;
}

// mfu:578-593: best of the three 2-column sub-problems (12, 13, 23; strict <).
__device__ __forceinline__ double fallback3(double y_sq, double a11, double a12, double a13,
                                            double a22, double a23, double a33, double Y1,
                                            double Y2, double Y3, double &w0, double &w1,
                                            double &w2)
{
    double u, v;
    double res = lsq2(y_sq, a11, a12, a22, Y1, Y2, u, v);
    w0 = u; w1 = v; w2 = 0.0;
    double r = lsq2(y_sq, a11, a13, a33, Y1, Y3, u, v);
    if (r < res) { w0 = u; w1 = 0.0; w2 = v; res = r; }
    r = lsq2(y_sq, a22, a23, a33, Y2, Y3, u, v);
    if (r < res) { w0 = 0.0; w1 = u; w2 = v; res = r; }
    return res;
}

// ---------------------------------------------------------------------------------
// (res, loop index) argmin helpers
// ---------------------------------------------------------------------------------
struct Best {
    double res;
    long long idx;
};
__device__ __forceinline__ void best_take(Best &b, double res, long long idx)
{
    if (res < b.res || (res == b.res && idx < b.idx)) { b.res = res; b.idx = idx; }
}
__device__ __forceinline__ Best best_block_reduce(Best b, Best *sm /* >= 32 */)
{
    for (int o = 16; o > 0; o >>= 1) {
        double r = __shfl_xor_sync(0xffffffffu, b.res, o);
        long long i = __shfl_xor_sync(0xffffffffu, b.idx, o);
        best_take(b, r, i);
    }
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
    if (lane == 0) sm[warp] = b;
    __syncthreads();
    if (warp == 0) {
        Best c;
        c.res = INFINITY; c.idx = LLONG_MAX;
        if (lane < nw) c = sm[lane];
        for (int o = 16; o > 0; o >>= 1) {
            double r = __shfl_xor_sync(0xffffffffu, c.res, o);
            long long i = __shfl_xor_sync(0xffffffffu, c.idx, o);
            best_take(c, r, i);
        }
        b = c;
    }
    return b;  // valid in thread 0
}

// ---------------------------------------------------------------------------------
// exact search: column statistics
// ---------------------------------------------------------------------------------
struct ExactArgs {
    int M;
    BlockSpec bs;
    const double *A;
    int64_t lda, strideA;
    const double *y;
    int64_t y_ld;
    const int32_t *vox_list;
    double *colsq, *ady, *ysq, *cross13, *cross23, *tile_res;
    long long *tile_idx;
    int ntiles;
    const int32_t *a_list;   // optional: dictionary index of voxel v (default: v itself)
    const uint8_t *tile_mask;  // optional: [v][2][mask_ld] atoms of block 1 / 2 whose rows / columns are scanned (others are skipped)
    int mask_ld;
};

__device__ __forceinline__ const double *ex_A(const ExactArgs &a, int64_t v)
{
    const int64_t r = a.a_list ? (int64_t)a.a_list[v] : v;
    return a.A + r * a.strideA;
}

// grid (ceil(ntot/128), V); dynamic smem M doubles (y)
__global__ void __launch_bounds__(128) k_colstats(ExactArgs a)
{
    extern __shared__ double ys[];
    const int64_t v = blockIdx.y;
    const int64_t row = a.vox_list ? a.vox_list[v] : v;
    const double *y = a.y + row * a.y_ld;
    for (int k = threadIdx.x; k < a.M; k += blockDim.x) ys[k] = y[k];
    __syncthreads();
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const double *A = ex_A(a, v);
    if (c < a.bs.ntot) {
        double sq = 0.0, dy = 0.0;
        for (int k = 0; k < a.M; k++) {
            double x = A[(size_t)k * a.lda + c];
            sq = DA(sq, DM(x, x));          // mfu:310, 513
            dy = DA(dy, DM(ys[k], x));      // mfu:325, 535
        }
        a.colsq[v * a.bs.ntot + c] = sq;
        a.ady[v * a.bs.ntot + c] = dy;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        double s = 0.0;
        for (int k = 0; k < a.M; k++) s = DA(s, DM(ys[k], ys[k]));  // mfu:323, 533
        a.ysq[v] = s;
    }
}

// cross13[v][i1][i3], cross23[v][i2][i3] (mfu:524-531); grid (ceil((N1+N2)*N3/128), V)
__global__ void __launch_bounds__(128) k_cross3(ExactArgs a)
{
    const int64_t v = blockIdx.y;
    const int N1 = a.bs.size[0], N2 = a.bs.size[1], N3 = a.bs.size[2];
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)(N1 + N2) * N3) return;
    const int i = (int)(t / N3), i3 = (int)(t % N3);
    const double *A = ex_A(a, v);
    const int c3 = a.bs.start[2] + i3;
    const int c = i < N1 ? a.bs.start[0] + i : a.bs.start[1] + (i - N1);
    double s = 0.0;
    for (int k = 0; k < a.M; k++)
        s = DA(s, DM(A[(size_t)k * a.lda + c], A[(size_t)k * a.lda + c3]));
    if (i < N1) a.cross13[(v * N1 + i) * N3 + i3] = s;
    else a.cross23[(v * N2 + (i - N1)) * N3 + i3] = s;
}

// ---------------------------------------------------------------------------------
// exact search, 1 block (mfu:225-278): grid (ceil(N1/256), V)
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_single(ExactArgs a)
{
    __shared__ Best red[32];
    const int64_t v = blockIdx.y;
    const int i1 = blockIdx.x * blockDim.x + threadIdx.x;
    const double y_sq = a.ysq[v];
    Best b;
    b.res = INFINITY; b.idx = LLONG_MAX;
    if (i1 < a.bs.size[0]) {
        double adoty = a.ady[v * a.bs.ntot + i1];
        if (adoty >= 0) {
            double w = DD(adoty, a.colsq[v * a.bs.ntot + i1]);
            double res = DS(y_sq, DM(w, adoty));
            if (res < y_sq) best_take(b, res, i1);
        }
    }
    b = best_block_reduce(b, red);
    if (threadIdx.x == 0) {
        a.tile_res[v * a.ntiles + blockIdx.x] = b.res;
        a.tile_idx[v * a.ntiles + blockIdx.x] = b.idx;
    }
}

// ---------------------------------------------------------------------------------
// exact search, 2 and 3 blocks.  CTA = 64x64 tile of (i1, i2) pairs, 256 threads, each
// thread 4x4 pairs; the k loop runs in reference order through KC-row smem chunks.
// grid (tilesJ, tilesI, V).
// ---------------------------------------------------------------------------------
#define EX_T 64
#define EX_KC 32

template <int NB>
__global__ void __launch_bounds__(256) k_pairs(ExactArgs a)
{
    __shared__ double s1[EX_KC][EX_T];
    __shared__ double s2[EX_KC][EX_T];
    __shared__ double s3[EX_KC];
    __shared__ double sy[EX_KC];
    __shared__ Best red[32];
    const int64_t v = blockIdx.z;
    const int tI = blockIdx.y * EX_T, tJ = blockIdx.x * EX_T;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    // `mine`: some pair of this thread lies in a row / column the search is restricted to
    bool mine = true;
    if (a.tile_mask) {
        // the screening tier certified that the minimum lies in the rows / columns of a few atoms
        // (one byte per atom, mask_ld a multiple of the tile size): tiles without such an atom are
        // skipped, and inside a tile only the threads that own one of those rows / columns work
        const uint8_t *mk = a.tile_mask + (size_t)v * 2 * a.mask_ld;
        const int t = threadIdx.x;
        const bool any = t < EX_T ? (tI + t < a.mask_ld && mk[tI + t])
                                  : (t < 2 * EX_T && tJ + t - EX_T < a.mask_ld && mk[a.mask_ld + tJ + t - EX_T]);
        if (!__syncthreads_or(any)) {
            if (threadIdx.x == 0) {
                const int tile = blockIdx.y * gridDim.x + blockIdx.x;
                a.tile_res[v * a.ntiles + tile] = INFINITY;
                a.tile_idx[v * a.ntiles + tile] = LLONG_MAX;
            }
            return;
        }
        mine = false;
#pragma unroll
        for (int r = 0; r < 4; r++) {
            if (tI + ty + 16 * r < a.mask_ld && mk[tI + ty + 16 * r]) mine = true;
            if (tJ + tx + 16 * r < a.mask_ld && mk[a.mask_ld + tJ + tx + 16 * r]) mine = true;
        }
    }
    const int64_t row = a.vox_list ? a.vox_list[v] : v;
    const int N1 = a.bs.size[0], N2 = a.bs.size[1];
    const double *A = ex_A(a, v);
    const double *B1 = A + a.bs.start[0], *B2 = A + a.bs.start[1];
    const double *yv = a.y + row * a.y_ld;
    const int M = a.M;

    double acc[4][4];
#pragma unroll
    for (int r = 0; r < 4; r++)
#pragma unroll
        for (int c = 0; c < 4; c++) acc[r][c] = 0.0;

    for (int k0 = 0; k0 < M; k0 += EX_KC) {
        const int kc = min(EX_KC, M - k0);
        for (int e = threadIdx.x; e < EX_KC * EX_T; e += 256) {
            int kk = e / EX_T, cc = e % EX_T;
            double x1 = 0.0, x2 = 0.0;
            if (kk < kc) {
                if (tI + cc < N1) x1 = B1[(size_t)(k0 + kk) * a.lda + tI + cc];
                if (tJ + cc < N2) x2 = B2[(size_t)(k0 + kk) * a.lda + tJ + cc];
            }
            s1[kk][cc] = x1;
            s2[kk][cc] = x2;
        }
        __syncthreads();
        for (int kk = 0; mine && kk < kc; kk++) {
            double av[4], bv[4];
#pragma unroll
            for (int r = 0; r < 4; r++) av[r] = s1[kk][ty + 16 * r];
#pragma unroll
            for (int c = 0; c < 4; c++) bv[c] = s2[kk][tx + 16 * c];
#pragma unroll
            for (int r = 0; r < 4; r++)
#pragma unroll
                for (int c = 0; c < 4; c++) acc[r][c] = DA(acc[r][c], DM(av[r], bv[c]));  // mfu:319, 523
        }
        __syncthreads();
    }

    const double y_sq = a.ysq[v];
    const double *colsq = a.colsq + v * a.bs.ntot, *ady = a.ady + v * a.bs.ntot;
    Best b;
    b.res = INFINITY; b.idx = LLONG_MAX;
    double A11[4], Y1[4], A22[4], Y2[4];
#pragma unroll
    for (int r = 0; r < 4; r++) {
        int i1 = tI + ty + 16 * r;
        A11[r] = i1 < N1 ? colsq[a.bs.start[0] + i1] : 1.0;
        Y1[r] = i1 < N1 ? ady[a.bs.start[0] + i1] : 0.0;
    }
#pragma unroll
    for (int c = 0; c < 4; c++) {
        int i2 = tJ + tx + 16 * c;
        A22[c] = i2 < N2 ? colsq[a.bs.start[1] + i2] : 1.0;
        Y2[c] = i2 < N2 ? ady[a.bs.start[1] + i2] : 0.0;
    }

    if (NB == 2) {
#pragma unroll
        for (int r = 0; r < 4; r++)
#pragma unroll
            for (int c = 0; c < 4; c++) {
                int i1 = tI + ty + 16 * r, i2 = tJ + tx + 16 * c;
                if (mine && i1 < N1 && i2 < N2) {
                    double w0, w1;
                    double res = lsq2(y_sq, A11[r], acc[r][c], A22[c], Y1[r], Y2[c], w0, w1);
                    if (res < y_sq) best_take(b, res, (long long)i1 * N2 + i2);
                }
            }
    } else {
        const int N3 = a.bs.size[2];
        const double *B3 = A + a.bs.start[2];
        for (int i3 = 0; i3 < N3; i3++) {
            const double a33 = colsq[a.bs.start[2] + i3], Y3 = ady[a.bs.start[2] + i3];
            // two half-batches of 8 pairs keep the register footprint bounded
            for (int half = 0; half < 2; half++) {
                double w0[8], w1[8], w2[8], res[8];
                unsigned posmask = 0;
#pragma unroll
                for (int rr = 0; rr < 2; rr++)
#pragma unroll
                    for (int c = 0; c < 4; c++) {
                        const int r = half * 2 + rr, q = rr * 4 + c;
                        const int i1 = tI + ty + 16 * r, i2 = tJ + tx + 16 * c;
                        res[q] = 0.0; w0[q] = w1[q] = w2[q] = 0.0;
                        if (mine && i1 < N1 && i2 < N2) {
                            double a13 = a.cross13[(v * N1 + i1) * N3 + i3];
                            double a23 = a.cross23[(v * N2 + i2) * N3 + i3];
                            if (cramer3(A11[r], acc[r][c], a13, A22[c], a23, a33, Y1[r], Y2[c], Y3,
                                        w0[q], w1[q], w2[q]))
                                posmask |= 1u << q;
                        }
                    }
                // direct residual for the all-positive pairs (mfu:569-573)
                if (__syncthreads_or(posmask != 0)) {
                    for (int k0 = 0; k0 < M; k0 += EX_KC) {
                        const int kc = min(EX_KC, M - k0);
                        for (int e = threadIdx.x; e < EX_KC * EX_T; e += 256) {
                            int kk = e / EX_T, cc = e % EX_T;
                            double x1 = 0.0, x2 = 0.0;
                            if (kk < kc) {
                                if (tI + cc < N1) x1 = B1[(size_t)(k0 + kk) * a.lda + tI + cc];
                                if (tJ + cc < N2) x2 = B2[(size_t)(k0 + kk) * a.lda + tJ + cc];
                            }
                            s1[kk][cc] = x1;
                            s2[kk][cc] = x2;
                        }
                        if (threadIdx.x < kc) {
                            s3[threadIdx.x] = B3[(size_t)(k0 + threadIdx.x) * a.lda + i3];
                            sy[threadIdx.x] = yv[k0 + threadIdx.x];
                        }
                        __syncthreads();
                        if (posmask) {
                            for (int kk = 0; kk < kc; kk++) {
                                const double a3 = s3[kk], yk = sy[kk];
#pragma unroll
                                for (int rr = 0; rr < 2; rr++)
#pragma unroll
                                    for (int c = 0; c < 4; c++) {
                                        const int r = half * 2 + rr, q = rr * 4 + c;
                                        if (posmask & (1u << q)) {
                                            double a1 = s1[kk][ty + 16 * r], a2 = s2[kk][tx + 16 * c];
                                            double d = DS(DA(DA(DM(w0[q], a1), DM(w1[q], a2)),
                                                             DM(w2[q], a3)), yk);
                                            res[q] = DA(res[q], DM(d, d));
                                        }
                                    }
                            }
                        }
                        __syncthreads();
                    }
                }
#pragma unroll
                for (int rr = 0; rr < 2; rr++)
#pragma unroll
                    for (int c = 0; c < 4; c++) {
                        const int r = half * 2 + rr, q = rr * 4 + c;
                        const int i1 = tI + ty + 16 * r, i2 = tJ + tx + 16 * c;
                        if (mine && i1 < N1 && i2 < N2) {
                            double rs = res[q];
                            if (!(posmask & (1u << q))) {
                                double a13 = a.cross13[(v * N1 + i1) * N3 + i3];
                                double a23 = a.cross23[(v * N2 + i2) * N3 + i3];
                                double u0, u1, u2;
                                rs = fallback3(y_sq, A11[r], acc[r][c], a13, A22[c], a23, a33,
                                               Y1[r], Y2[c], Y3, u0, u1, u2);
                            }
                            if (rs < y_sq)
                                best_take(b, rs, ((long long)i3 * N1 + i1) * N2 + i2);
                        }
                    }
            }
        }
    }
    b = best_block_reduce(b, red);
    if (threadIdx.x == 0) {
        int tile = blockIdx.y * gridDim.x + blockIdx.x;
        a.tile_res[v * a.ntiles + tile] = b.res;
        a.tile_idx[v * a.ntiles + tile] = b.idx;
    }
}

// ---------------------------------------------------------------------------------
// 4 and 5 blocks (mfu:612-657 solve_exhaustive_posweights_4up).  The reference calls
// scipy.optimize.nnls on every index tuple (itertools.product order) and keeps the first
// strict minimum of rnorm^2.  Here the tiny NNLS is solved in closed form from the Gram
// entries by enumerating the 2^nb - 1 supports (normal equations on the support, all
// weights > 0, best gain): same optimum as Lawson-Hanson up to rounding, so
// parity for these shapes is to rounding, not bitwise.
// ---------------------------------------------------------------------------------
__device__ double nnls_enum(int nb, const double *G /* 5x5 */, const double *Y, double *wbest)
{
    double best = 0.0;
    for (int b = 0; b < kMaxBlocks; b++) wbest[b] = 0.0;
    for (int mask = 1; mask < (1 << nb); mask++) {
        int id[kMaxBlocks], n = 0;
        for (int b = 0; b < nb; b++)
            if (mask & (1 << b)) id[n++] = b;
        // Cholesky of the support's Gram block
        double L[kMaxBlocks][kMaxBlocks], z[kMaxBlocks], w[kMaxBlocks];
        bool ok = true;
        for (int i = 0; i < n && ok; i++) {
            for (int j = 0; j <= i; j++) {
                double sacc = G[id[i] * kMaxBlocks + id[j]];
                for (int k = 0; k < j; k++) sacc -= L[i][k] * L[j][k];
                if (i == j) {
                    if (!(sacc > 0.0)) { ok = false; break; }
                    L[i][i] = sqrt(sacc);
                } else {
                    L[i][j] = sacc / L[j][j];
                }
            }
        }
        if (!ok) continue;
        for (int i = 0; i < n; i++) {
            double sacc = Y[id[i]];
            for (int k = 0; k < i; k++) sacc -= L[i][k] * z[k];
            z[i] = sacc / L[i][i];
        }
        for (int i = n - 1; i >= 0; i--) {
            double sacc = z[i];
            for (int k = i + 1; k < n; k++) sacc -= L[k][i] * w[k];
            w[i] = sacc / L[i][i];
        }
        bool pos = true;
        double gain = 0.0;
        for (int i = 0; i < n; i++) { pos = pos && (w[i] > 0.0); gain += Y[id[i]] * w[i]; }
        if (pos && gain > best) {
            best = gain;
            for (int b = 0; b < kMaxBlocks; b++) wbest[b] = 0.0;
            for (int i = 0; i < n; i++) wbest[id[i]] = w[i];
        }
    }
    return best;
}

struct MultiArgs {
    ExactArgs e;
    int S;                 // total columns of blocks 3..nb
    long long T;           // number of trailing tuples = prod(size[2:])
    const double *crossS;  // [V][N1+N2][S]   (a_i . a_s) for i in blocks 1-2, s in blocks 3..
    const double *crossSS; // [V][S][S]
};

// grid (ceil((N1+N2+S)*S/128), V): cross terms with the trailing blocks, reference-style sums
__global__ void __launch_bounds__(128) k_cross_small(MultiArgs ma, double *crossS, double *crossSS)
{
    const ExactArgs &a = ma.e;
    const int64_t v = blockIdx.y;
    const int N12 = a.bs.size[0] + a.bs.size[1], S = ma.S;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)(N12 + S) * S) return;
    const int i = (int)(t / S), sidx = (int)(t % S);
    const double *A = ex_A(a, v);
    const int cs = a.bs.start[2] + sidx;
    const int c = i < N12 ? i : a.bs.start[2] + (i - N12);
    double acc = 0.0;
    for (int k = 0; k < a.M; k++) acc = DA(acc, DM(A[(size_t)k * a.lda + c], A[(size_t)k * a.lda + cs]));
    if (i < N12) crossS[(v * N12 + i) * S + sidx] = acc;
    else crossSS[(v * S + (i - N12)) * S + sidx] = acc;
}

__global__ void __launch_bounds__(256) k_pairs_multi(MultiArgs ma)
{
    const ExactArgs &a = ma.e;
    __shared__ double s1[EX_KC][EX_T];
    __shared__ double s2[EX_KC][EX_T];
    __shared__ Best red[32];
    const int64_t v = blockIdx.z;
    const int N1 = a.bs.size[0], N2 = a.bs.size[1], nb = a.bs.nb, S = ma.S;
    const int tI = blockIdx.y * EX_T, tJ = blockIdx.x * EX_T;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const double *A = ex_A(a, v);
    const double *B1 = A + a.bs.start[0], *B2 = A + a.bs.start[1];
    const int M = a.M;
    double acc[4][4];
#pragma unroll
    for (int r = 0; r < 4; r++)
#pragma unroll
        for (int c = 0; c < 4; c++) acc[r][c] = 0.0;
    for (int k0 = 0; k0 < M; k0 += EX_KC) {
        const int kc = min(EX_KC, M - k0);
        for (int e = threadIdx.x; e < EX_KC * EX_T; e += 256) {
            int kk = e / EX_T, cc = e % EX_T;
            double x1 = 0.0, x2 = 0.0;
            if (kk < kc) {
                if (tI + cc < N1) x1 = B1[(size_t)(k0 + kk) * a.lda + tI + cc];
                if (tJ + cc < N2) x2 = B2[(size_t)(k0 + kk) * a.lda + tJ + cc];
            }
            s1[kk][cc] = x1;
            s2[kk][cc] = x2;
        }
        __syncthreads();
        for (int kk = 0; kk < kc; kk++) {
            double av[4], bv[4];
#pragma unroll
            for (int r = 0; r < 4; r++) av[r] = s1[kk][ty + 16 * r];
#pragma unroll
            for (int c = 0; c < 4; c++) bv[c] = s2[kk][tx + 16 * c];
#pragma unroll
            for (int r = 0; r < 4; r++)
#pragma unroll
                for (int c = 0; c < 4; c++) acc[r][c] = DA(acc[r][c], DM(av[r], bv[c]));
        }
        __syncthreads();
    }
    const double y_sq = a.ysq[v];
    const double *colsq = a.colsq + v * a.bs.ntot, *ady = a.ady + v * a.bs.ntot;
    const double *cS = ma.crossS + v * (int64_t)(N1 + N2) * S;
    const double *cSS = ma.crossSS + v * (int64_t)S * S;
    Best b;
    b.res = INFINITY; b.idx = LLONG_MAX;
#pragma unroll 1
    for (int q = 0; q < 16; q++) {
        const int r = q >> 2, c = q & 3;
        const int i1 = tI + ty + 16 * r, i2 = tJ + tx + 16 * c;
        if (i1 >= N1 || i2 >= N2) continue;
        double a12 = 0.0;
#pragma unroll
        for (int rr = 0; rr < 4; rr++)
#pragma unroll
            for (int cc = 0; cc < 4; cc++)
                if (rr == r && cc == c) a12 = acc[rr][cc];
        double G[kMaxBlocks * kMaxBlocks], Y[kMaxBlocks], w[kMaxBlocks];
        G[0] = colsq[a.bs.start[0] + i1]; Y[0] = ady[a.bs.start[0] + i1];
        G[kMaxBlocks + 1] = colsq[a.bs.start[1] + i2]; Y[1] = ady[a.bs.start[1] + i2];
        G[1] = G[kMaxBlocks] = a12;
        for (long long t = 0; t < ma.T; t++) {
            int idx[kMaxBlocks];
            long long rem = t;
            for (int bb = nb - 1; bb >= 2; bb--) { idx[bb] = (int)(rem % a.bs.size[bb]); rem /= a.bs.size[bb]; }
            for (int bb = 2; bb < nb; bb++) {
                const int sb = a.bs.start[bb] - a.bs.start[2] + idx[bb];   // column among the trailing ones
                G[bb * kMaxBlocks + bb] = colsq[a.bs.start[bb] + idx[bb]];
                Y[bb] = ady[a.bs.start[bb] + idx[bb]];
                G[bb] = G[bb * kMaxBlocks] = cS[(int64_t)i1 * S + sb];
                G[kMaxBlocks + bb] = G[bb * kMaxBlocks + 1] = cS[(int64_t)(N1 + i2) * S + sb];
                for (int b2 = 2; b2 < bb; b2++) {
                    const int sb2 = a.bs.start[b2] - a.bs.start[2] + idx[b2];
                    G[b2 * kMaxBlocks + bb] = G[bb * kMaxBlocks + b2] = cSS[(int64_t)sb2 * S + sb];
                }
            }
            const double gain = nnls_enum(nb, G, Y, w);
            const double res = y_sq - gain;
            if (gain > 0.0 && res < y_sq)
                best_take(b, res, ((long long)i1 * N2 + i2) * ma.T + t);   // product order
        }
    }
    b = best_block_reduce(b, red);
    if (threadIdx.x == 0) {
        int tile = blockIdx.y * gridDim.x + blockIdx.x;
        a.tile_res[v * a.ntiles + tile] = b.res;
        a.tile_idx[v * a.ntiles + tile] = b.idx;
    }
}

// per voxel: min over tiles; one warp per voxel
__global__ void __launch_bounds__(128) k_reduce_tiles(int64_t V, int ntiles, const double *tile_res,
                                                      const long long *tile_idx,
                                                      const int32_t *vox_list, long long *tuple)
{
    int64_t v = (int64_t)blockIdx.x * 4 + (threadIdx.x >> 5);
    if (v >= V) return;
    int lane = threadIdx.x & 31;
    Best b;
    b.res = INFINITY; b.idx = LLONG_MAX;
    for (int t = lane; t < ntiles; t += 32) best_take(b, tile_res[v * ntiles + t], tile_idx[v * ntiles + t]);
    for (int o = 16; o > 0; o >>= 1) {
        double r = __shfl_xor_sync(0xffffffffu, b.res, o);
        long long i = __shfl_xor_sync(0xffffffffu, b.idx, o);
        best_take(b, r, i);
    }
    if (lane == 0) {
        int64_t row = vox_list ? vox_list[v] : v;
        tuple[row] = (b.idx == LLONG_MAX) ? kNoTuple : b.idx;
    }
}

static int exact_ntiles(const BlockSpec &bs)
{
    if (bs.nb == 1) return (bs.size[0] + 255) / 256;
    return ((bs.size[0] + EX_T - 1) / EX_T) * ((bs.size[1] + EX_T - 1) / EX_T);
}

static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

size_t exact_scratch_bytes(int64_t V, const BlockSpec &bs)
{
    size_t s = 0;
    s += 2 * align256(sizeof(double) * V * bs.ntot);
    s += align256(sizeof(double) * V);
    if (bs.nb == 3) {
        s += align256(sizeof(double) * V * bs.size[0] * bs.size[2]);
        s += align256(sizeof(double) * V * bs.size[1] * bs.size[2]);
    }
    if (bs.nb >= 4) {
        const size_t S = bs.ntot - bs.start[2];
        s += align256(sizeof(double) * V * (bs.size[0] + bs.size[1]) * S);
        s += align256(sizeof(double) * V * S * S);
    }
    s += align256(sizeof(double) * V * exact_ntiles(bs));
    s += align256(sizeof(long long) * V * exact_ntiles(bs));
    return s;
}

int launch_exact_search(int64_t V, int M, const BlockSpec &bs, const double *A, int64_t lda,
                        int64_t strideA, const double *y, int64_t y_ld, const int32_t *vox_list,
                        void *scratch, long long *tuple_out, cudaStream_t st, cudaEvent_t *ev,
                        const int32_t *a_list, const uint8_t *tile_mask, int mask_ld)
{
    if (V == 0) return MFB_OK;
    if (bs.nb < 1 || bs.nb > kMaxBlocks) {
        set_error("exact search supports 1-5 blocks");
        return MFB_EUNSUPPORTED;
    }
    if (V > 65535) {
        set_error("exact search: at most 65535 voxels per launch");
        return MFB_EINVAL;
    }
    ExactArgs a;
    a.M = M; a.bs = bs; a.A = A; a.lda = lda; a.strideA = strideA; a.y = y; a.y_ld = y_ld;
    a.vox_list = vox_list;
    a.a_list = a_list;
    a.tile_mask = (bs.nb == 2 || bs.nb == 3) ? tile_mask : nullptr;
    a.mask_ld = mask_ld;
    a.ntiles = exact_ntiles(bs);
    char *p = (char *)scratch;
    a.colsq = (double *)p; p += align256(sizeof(double) * V * bs.ntot);
    a.ady = (double *)p; p += align256(sizeof(double) * V * bs.ntot);
    a.ysq = (double *)p; p += align256(sizeof(double) * V);
    a.cross13 = a.cross23 = nullptr;
    if (bs.nb == 3) {
        a.cross13 = (double *)p; p += align256(sizeof(double) * V * bs.size[0] * bs.size[2]);
        a.cross23 = (double *)p; p += align256(sizeof(double) * V * bs.size[1] * bs.size[2]);
    }
    double *crossS = nullptr, *crossSS = nullptr;
    if (bs.nb >= 4) {
        const size_t S = bs.ntot - bs.start[2];
        crossS = (double *)p; p += align256(sizeof(double) * V * (bs.size[0] + bs.size[1]) * S);
        crossSS = (double *)p; p += align256(sizeof(double) * V * S * S);
    }
    a.tile_res = (double *)p; p += align256(sizeof(double) * V * a.ntiles);
    a.tile_idx = (long long *)p;

    MFB_LAUNCH(k_colstats, dim3((bs.ntot + 127) / 128, (unsigned)V), 128, M * sizeof(double), st, a);
    if (bs.nb == 1) {
        MFB_LAUNCH(k_single, dim3((bs.size[0] + 255) / 256, (unsigned)V), 256, 0, st, a);
    } else {
        dim3 grid((bs.size[1] + EX_T - 1) / EX_T, (bs.size[0] + EX_T - 1) / EX_T, (unsigned)V);
        if (bs.nb == 3) {
            long long n = (long long)(bs.size[0] + bs.size[1]) * bs.size[2];
            MFB_LAUNCH(k_cross3, dim3((unsigned)((n + 127) / 128), (unsigned)V), 128, 0, st, a);
        }
        MultiArgs ma;
        if (bs.nb >= 4) {
            ma.e = a;
            ma.S = bs.ntot - bs.start[2];
            ma.T = 1;
            for (int b = 2; b < bs.nb; b++) ma.T *= bs.size[b];
            ma.crossS = crossS; ma.crossSS = crossSS;
            long long n = (long long)(bs.size[0] + bs.size[1] + ma.S) * ma.S;
            MFB_LAUNCH(k_cross_small, dim3((unsigned)((n + 127) / 128), (unsigned)V), 128, 0, st, ma, crossS, crossSS);
        }
        if (ev) MFB_CUDA_TRY(cudaEventRecord(ev[0], st));
        if (bs.nb == 2) MFB_LAUNCH(k_pairs<2>, grid, 256, 0, st, a);
        else if (bs.nb == 3) MFB_LAUNCH(k_pairs<3>, grid, 256, 0, st, a);
        else MFB_LAUNCH(k_pairs_multi, grid, 256, 0, st, ma);
        if (ev) MFB_CUDA_TRY(cudaEventRecord(ev[1], st));
    }
    MFB_LAUNCH(k_reduce_tiles, (unsigned)((V + 3) / 4), 128, 0, st, V, a.ntiles, a.tile_res,
               a.tile_idx, vox_list, tuple_out);
    return MFB_OK;
}

// ---------------------------------------------------------------------------------
// one-fascicle voxels ([N], [N,1], [N,E], [N,1,E]): rotation, Gram terms and closed forms
// fused in one kernel, nothing materialised.  One CTA per voxel, one thread per atom; the
// atom's rotated column is produced on the fly from the lookup table (two coalesced row
// reads per measurement) and every sum runs in the reference's order.
// ---------------------------------------------------------------------------------
#define SF_MAXISO 12

struct SfArgs {
    DevPlan p;
    const int32_t *vox_list;
    const double *peaks;
    int peaks_ld;
    const double *y;
    int csf, ear;
    long long *tuple;
};

__global__ void __launch_bounds__(256) k_single_fascicle(SfArgs a)
{
    extern __shared__ double sm[];
    const DevPlan &p = a.p;
    const int M = p.M, N = p.N;
    const int nIso = a.csf + a.ear * p.E;
    double *wl = sm, *wh = sm + M, *wl2 = sm + 2 * M, *wh2 = sm + 3 * M, *ys = sm + 4 * M;
    double *iso = ys + M;                       // [nIso][M] iso columns (csf first)
    double *isoSq = iso + (size_t)nIso * M;     // [nIso]
    double *isoY = isoSq + SF_MAXISO;           // [nIso]
    double *isoX = isoY + SF_MAXISO;            // [nIso] csf . ear[e]
    int *rl = (int *)(isoX + SF_MAXISO), *rh = rl + M, *rl2 = rh + M, *rh2 = rl2 + M;
    __shared__ Best red[32];
    __shared__ double s_ysq;
    const int64_t v = blockIdx.x;
    const int64_t row = a.vox_list ? a.vox_list[v] : v;
    const double ux = a.peaks[row * a.peaks_ld], uy = a.peaks[row * a.peaks_ld + 1],
                 uz = a.peaks[row * a.peaks_ld + 2];
    for (int m = threadIdx.x; m < M; m += blockDim.x) {
        double x = dir_dot(p, m, ux, uy, uz);
        Lerp l1 = shell_lerp(p, p.shell_lo[m], x);
        rl[m] = l1.rl; rh[m] = l1.rh; wl[m] = l1.wl; wh[m] = l1.wh;
        if (p.shell_hi[m] != p.shell_lo[m]) {
            Lerp l2 = shell_lerp(p, p.shell_hi[m], x);
            rl2[m] = l2.rl; rh2[m] = l2.rh; wl2[m] = l2.wl; wh2[m] = l2.wh;
        } else {
            rl2[m] = -1;
        }
        ys[m] = a.y[row * M + m];
        for (int e = 0; e < nIso; e++)
            iso[(size_t)e * M + m] = (a.csf && e == 0) ? p.sig_csf[m] : p.sig_ear[(size_t)m * p.E + (e - a.csf)];
    }
    __syncthreads();
    if (threadIdx.x < nIso) {  // statistics of the iso columns, reference order
        const int e = threadIdx.x;
        double sq = 0.0, dy = 0.0, cx = 0.0;
        for (int k = 0; k < M; k++) {
            double c = iso[(size_t)e * M + k];
            sq = DA(sq, DM(c, c));
            dy = DA(dy, DM(ys[k], c));
            if (a.csf && e > 0) cx = DA(cx, DM(iso[k], c));   // A23 = csf . ear (mfu:528-531)
        }
        isoSq[e] = sq; isoY[e] = dy; isoX[e] = cx;
    }
    if (threadIdx.x == 32) {
        double s = 0.0;
        for (int k = 0; k < M; k++) s = DA(s, DM(ys[k], ys[k]));
        s_ysq = s;
    }
    __syncthreads();
    const double y_sq = s_ysq;
    const int nb = 1 + a.csf + a.ear;

    auto column = [&](int m, int i) -> double {
        Lerp la, lb;
        la.rl = rl[m]; la.rh = rh[m]; la.wl = wl[m]; la.wh = wh[m];
        const bool between = rl2[m] >= 0;
        double gwl = 0.0, gwh = 0.0;
        lb = la;
        if (between) { lb.rl = rl2[m]; lb.rh = rh2[m]; lb.wl = wl2[m]; lb.wh = wh2[m]; gwl = p.gw_lo[m]; gwh = p.gw_hi[m]; }
        return rot_entry(p, la, lb, between, gwl, gwh, i);
    };

    Best b;
    b.res = INFINITY; b.idx = LLONG_MAX;
    for (int i1 = threadIdx.x; i1 < N; i1 += blockDim.x) {
        double sq = 0.0, dy = 0.0, cr[SF_MAXISO];
#pragma unroll
        for (int e = 0; e < SF_MAXISO; e++) cr[e] = 0.0;
        for (int k = 0; k < M; k++) {
            const double d = column(k, i1);
            sq = DA(sq, DM(d, d));
            dy = DA(dy, DM(ys[k], d));
#pragma unroll
            for (int e = 0; e < SF_MAXISO; e++)
                if (e < nIso) cr[e] = DA(cr[e], DM(d, iso[(size_t)e * M + k]));
        }
        if (nb == 1) {  // mfu:252-273
            if (dy >= 0) {
                double w = DD(dy, sq);
                double res = DS(y_sq, DM(w, dy));
                if (res < y_sq) best_take(b, res, i1);
            }
        } else if (nb == 2) {  // mfu:329-386, second block = the iso columns
#pragma unroll
            for (int e = 0; e < SF_MAXISO; e++)
                if (e < nIso) {
                    double w0, w1;
                    double res = lsq2(y_sq, sq, cr[e], isoSq[e], dy, isoY[e], w0, w1);
                    if (res < y_sq) best_take(b, res, (long long)i1 * nIso + e);
                }
        } else {  // [N, 1, E]: loop order i3 (EAR), i1, i2 = 0 (mfu:540-601)
#pragma unroll
            for (int e = 1; e < SF_MAXISO; e++)
                if (e < nIso) {
                    const int i3 = e - 1;
                    double w0, w1, w2, res;
                    if (cramer3(sq, cr[0], cr[e], isoSq[0], isoX[e], isoSq[e], dy, isoY[0], isoY[e], w0, w1, w2)) {
                        res = 0.0;
                        for (int k = 0; k < M; k++) {
                            double dd = DS(DA(DA(DM(w0, column(k, i1)), DM(w1, iso[k])),
                                              DM(w2, iso[(size_t)e * M + k])), ys[k]);
                            res = DA(res, DM(dd, dd));
                        }
                    } else {
                        res = fallback3(y_sq, sq, cr[0], cr[e], isoSq[0], isoX[e], isoSq[e], dy, isoY[0],
                                        isoY[e], w0, w1, w2);
                    }
                    if (res < y_sq) best_take(b, res, (long long)i3 * N + i1);
                }
        }
    }
    b = best_block_reduce(b, red);
    if (threadIdx.x == 0) a.tuple[row] = (b.idx == LLONG_MAX) ? kNoTuple : b.idx;
}

static size_t single_fascicle_smem(const DevPlan &p, int csf, int ear)
{
    const int nIso = csf + ear * p.E;
    return sizeof(double) * ((size_t)5 * p.M + (size_t)nIso * p.M + 3 * SF_MAXISO) + sizeof(int) * 4 * p.M;
}

// Long protocols (shared-memory plan above 200 KB, M beyond ~3000) are not an error: the
// caller falls through to the materialising exact tier, which has no limit on M.
bool single_fascicle_supported(const DevPlan &p, int K, int csf, int ear)
{
    return K == 1 && csf + ear * p.E <= SF_MAXISO && single_fascicle_smem(p, csf, ear) <= 200 * 1024;
}

int launch_single_fascicle(const DevPlan &p, int64_t nvox, const int32_t *vox_list,
                           const double *peaks, int peaks_ld, const double *y, int csf, int ear,
                           long long *tuple, cudaStream_t st)
{
    if (nvox == 0) return MFB_OK;
    SfArgs a;
    a.p = p; a.vox_list = vox_list; a.peaks = peaks; a.peaks_ld = peaks_ld; a.y = y;
    a.csf = csf; a.ear = ear; a.tuple = tuple;
    const size_t smem = single_fascicle_smem(p, csf, ear);
    if (smem > 200 * 1024) { set_error("single-fascicle kernel: protocol too long for shared memory"); return MFB_EUNSUPPORTED; }
    if (smem > 48 * 1024)  // per device / context attribute
        MFB_CUDA_TRY(cudaFuncSetAttribute(k_single_fascicle, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t maxgrid = 1 << 20;
    for (int64_t v0 = 0; v0 < nvox; v0 += maxgrid) {
        SfArgs b = a;
        b.vox_list = vox_list + v0;
        MFB_LAUNCH(k_single_fascicle, (unsigned)std::min<int64_t>(maxgrid, nvox - v0), 256, smem, st, b);
    }
    return MFB_OK;
}

// ---------------------------------------------------------------------------------
// tuple decode + gather of the winning columns
// ---------------------------------------------------------------------------------
__device__ __forceinline__ void decode_tuple(const BlockSpec &bs, long long L, int *idx)
{
    for (int b = 0; b < kMaxBlocks; b++) idx[b] = 0;
    if (L < 0) return;
    if (bs.nb == 1) {
        idx[0] = (int)L;
    } else if (bs.nb == 2) {
        idx[0] = (int)(L / bs.size[1]); idx[1] = (int)(L % bs.size[1]);
    } else if (bs.nb == 3) {  // loop order i3, i1, i2 (mfu:540-547)
        idx[1] = (int)(L % bs.size[1]); L /= bs.size[1];
        idx[0] = (int)(L % bs.size[0]); idx[2] = (int)(L / bs.size[0]);
    } else {  // itertools.product order (mfu:637): last block fastest
        for (int b = bs.nb - 1; b >= 0; b--) { idx[b] = (int)(L % bs.size[b]); L /= bs.size[b]; }
    }
}

__global__ void __launch_bounds__(128)
k_gather_A(int64_t V, int M, BlockSpec bs, const double *A, int64_t lda, int64_t strideA,
           const long long *tuple, const int32_t *vox_list, double *Asmall, int32_t *idx_sub)
{
    const int64_t v = blockIdx.x;
    const int64_t row = vox_list ? vox_list[v] : v;
    int idx[kMaxBlocks];
    decode_tuple(bs, tuple[row], idx);
    const double *Av = A + v * strideA;
    for (int e = threadIdx.x; e < M * kMaxBlocks; e += blockDim.x) {
        int k = e / kMaxBlocks, b = e % kMaxBlocks;
        Asmall[(row * M + k) * kMaxBlocks + b] =
            b < bs.nb ? Av[(size_t)k * lda + bs.start[b] + idx[b]] : 0.0;
    }
    if (threadIdx.x < kMaxBlocks) idx_sub[row * kMaxBlocks + threadIdx.x] = idx[threadIdx.x];
}

int launch_gather_from_A(int64_t V, int M, const BlockSpec &bs, const double *A, int64_t lda,
                         int64_t strideA, const long long *tuple, const int32_t *vox_list,
                         double *Asmall, int32_t *idx_sub, cudaStream_t st)
{
    if (V == 0) return MFB_OK;
    MFB_LAUNCH(k_gather_A, (unsigned)V, 128, 0, st, V, M, bs, A, lda, strideA, tuple, vox_list,
               Asmall, idx_sub);
    return MFB_OK;
}

// Fast tier: rotate only the selected atoms (one column per fascicle) from the table.
__global__ void __launch_bounds__(128)
k_gather_table(DevPlan p, int64_t V, BlockSpec bs, int K, int csf, int ear, const double *peaks,
               int peaks_ld, const long long *tuple, const int32_t *vox_list, double *Asmall,
               int32_t *idx_sub)
{
    const int64_t v = blockIdx.x;
    const int64_t row = vox_list ? vox_list[v] : v;
    int idx[kMaxBlocks];
    decode_tuple(bs, tuple[row], idx);
    for (int m = threadIdx.x; m < p.M; m += blockDim.x) {
        double *out = Asmall + (row * p.M + m) * kMaxBlocks;
        for (int b = 0; b < kMaxBlocks; b++) out[b] = 0.0;
        for (int k = 0; k < K; k++) {
            const double *u = peaks + row * peaks_ld + 3 * k;
            double x = dir_dot(p, m, u[0], u[1], u[2]);
            Lerp a = shell_lerp(p, p.shell_lo[m], x), bb = a;
            bool between = p.shell_hi[m] != p.shell_lo[m];
            double gwl = 0.0, gwh = 0.0;
            if (between) { bb = shell_lerp(p, p.shell_hi[m], x); gwl = p.gw_lo[m]; gwh = p.gw_hi[m]; }
            out[k] = rot_entry(p, a, bb, between, gwl, gwh, idx[k]);
        }
        if (csf) out[K] = p.sig_csf[m];
        if (ear) out[K + csf] = p.sig_ear[(size_t)m * p.E + idx[K + csf]];
    }
    if (threadIdx.x < kMaxBlocks) idx_sub[row * kMaxBlocks + threadIdx.x] = idx[threadIdx.x];
}

int launch_gather_from_table(const DevPlan &p, int64_t V, int K, int csf, int ear,
                             const double *peaks, int peaks_ld, const long long *tuple,
                             const int32_t *vox_list, double *Asmall, int32_t *idx_sub,
                             cudaStream_t st)
{
    if (V == 0) return MFB_OK;
    BlockSpec bs;
    bs.nb = 0; bs.ntot = 0;
    for (int k = 0; k < K; k++) { bs.size[bs.nb] = p.N; bs.start[bs.nb++] = bs.ntot; bs.ntot += p.N; }
    if (csf) { bs.size[bs.nb] = 1; bs.start[bs.nb++] = bs.ntot; bs.ntot += 1; }
    if (ear) { bs.size[bs.nb] = p.E; bs.start[bs.nb++] = bs.ntot; bs.ntot += p.E; }
    MFB_LAUNCH(k_gather_table, (unsigned)V, 128, 0, st, p, V, bs, K, csf, ear, peaks, peaks_ld,
               tuple, vox_list, Asmall, idx_sub);
    return MFB_OK;
}

// ---------------------------------------------------------------------------------
// exact evaluation of one tuple (sizes [1]*nb): literal reference logic on the compact
// column matrix.  One thread per voxel.
// ---------------------------------------------------------------------------------
__device__ void eval_tuple(int M, int nb, const double *As /* M x kMaxBlocks */, const double *y,
                           int64_t ys, double *w, double &obj)
{
    double sq[3] = {0, 0, 0}, dy[3] = {0, 0, 0}, c12 = 0, c13 = 0, c23 = 0, y_sq = 0;
    for (int k = 0; k < M; k++) {
        const double yk = y[k * ys];
        const double *r = As + (size_t)k * kMaxBlocks;
        y_sq = DA(y_sq, DM(yk, yk));
        for (int b = 0; b < 3; b++)
            if (b < nb) { sq[b] = DA(sq[b], DM(r[b], r[b])); dy[b] = DA(dy[b], DM(yk, r[b])); }
        if (nb >= 2) c12 = DA(c12, DM(r[0], r[1]));
        if (nb >= 3) { c13 = DA(c13, DM(r[0], r[2])); c23 = DA(c23, DM(r[1], r[2])); }
    }
    for (int b = 0; b < kMaxBlocks; b++) w[b] = 0.0;
    obj = y_sq;
    if (nb == 1) {
        if (dy[0] >= 0) {
            double ww = DD(dy[0], sq[0]);
            double res = DS(y_sq, DM(ww, dy[0]));
            if (res < y_sq) { w[0] = ww; obj = res; }
        }
    } else if (nb == 2) {
        double w0, w1;
        double res = lsq2(y_sq, sq[0], c12, sq[1], dy[0], dy[1], w0, w1);
        if (res < y_sq) { w[0] = w0; w[1] = w1; obj = res; }
    } else if (nb == 3) {
        double w0, w1, w2, res;
        if (cramer3(sq[0], c12, c13, sq[1], c23, sq[2], dy[0], dy[1], dy[2], w0, w1, w2)) {
            res = 0.0;
            for (int k = 0; k < M; k++) {
                const double *r = As + (size_t)k * kMaxBlocks;
                double d = DS(DA(DA(DM(w0, r[0]), DM(w1, r[1])), DM(w2, r[2])), y[k * ys]);
                res = DA(res, DM(d, d));
            }
        } else {
            res = fallback3(y_sq, sq[0], c12, c13, sq[1], c23, sq[2], dy[0], dy[1], dy[2], w0,
                            w1, w2);
        }
        if (res < y_sq) { w[0] = w0; w[1] = w1; w[2] = w2; obj = res; }
    } else {
        // 4-5 blocks: NNLS by support enumeration on the tuple's Gram matrix; the objective is
        // the direct residual, like scipy's rnorm^2 (mfu:640-641)
        double G[kMaxBlocks * kMaxBlocks], Y[kMaxBlocks], ww[kMaxBlocks];
        for (int i = 0; i < nb; i++) {
            double dyi = 0.0;
            for (int k = 0; k < M; k++) dyi = DA(dyi, DM(y[k * ys], As[(size_t)k * kMaxBlocks + i]));
            Y[i] = dyi;
            for (int j = 0; j <= i; j++) {
                double g = 0.0;
                for (int k = 0; k < M; k++)
                    g = DA(g, DM(As[(size_t)k * kMaxBlocks + i], As[(size_t)k * kMaxBlocks + j]));
                G[i * kMaxBlocks + j] = G[j * kMaxBlocks + i] = g;
            }
        }
        const double gain = nnls_enum(nb, G, Y, ww);
        if (gain > 0.0) {
            double res = 0.0;
            for (int k = 0; k < M; k++) {
                double r = -y[k * ys];
                for (int i = 0; i < nb; i++) r += As[(size_t)k * kMaxBlocks + i] * ww[i];
                res += r * r;
            }
            if (res < y_sq) { for (int i = 0; i < nb; i++) w[i] = ww[i]; obj = res; }
        }
    }
}

__global__ void __launch_bounds__(128)
k_evaluate(int64_t V, int M, int nb_fixed, const uint8_t *nbv, const double *Asmall,
           const double *y, int64_t y_ld, const long long *tuple, double *w, double *obj,
           double *y_rec, int32_t *idx_sub)
{
    int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= V) return;
    const int nb = nb_fixed > 0 ? nb_fixed : nbv[v];
    const double *As = Asmall + v * M * kMaxBlocks;
    const double *yv = y + v * y_ld;
    double ww[kMaxBlocks], o;
    if (nb == 0 || tuple[v] < 0) {
        for (int b = 0; b < kMaxBlocks; b++) ww[b] = 0.0;
        double s = 0.0;
        for (int k = 0; k < M; k++) s = DA(s, DM(yv[k], yv[k]));
        o = nb == 0 ? 0.0 : s;
    } else {
        eval_tuple(M, nb, As, yv, 1, ww, o);
        bool zero = true;
        for (int b = 0; b < nb; b++) zero = zero && (ww[b] == 0.0);
        // the reference keeps indices 0 when nothing beats w = 0 (mfu:246-249, 296-297)
        if (zero && idx_sub)
            for (int b = 0; b < kMaxBlocks; b++) idx_sub[v * kMaxBlocks + b] = 0;
    }
    for (int b = 0; b < kMaxBlocks; b++) w[v * kMaxBlocks + b] = ww[b];
    obj[v] = o;
    if (y_rec) {
        // y_recons = A[:, ind] . w (mfu:277, 391, 606), left to right
        for (int k = 0; k < M; k++) {
            double s = 0.0;
            for (int b = 0; b < nb; b++) s = DA(s, DM(As[(size_t)k * kMaxBlocks + b], ww[b]));
            y_rec[v * M + k] = s;
        }
    }
}

int launch_evaluate(int64_t V, int M, int nb_fixed, const uint8_t *nbv, const double *Asmall,
                    const double *y, int64_t y_ld, const long long *tuple, double *w, double *obj,
                    double *y_rec, int32_t *idx_sub, cudaStream_t st)
{
    if (V == 0) return MFB_OK;
    MFB_LAUNCH(k_evaluate, (unsigned)((V + 127) / 128), 128, 0, st, V, M, nb_fixed, nbv, Asmall, y,
               y_ld, tuple, w, obj, y_rec, idx_sub);
    return MFB_OK;
}

// ---------------------------------------------------------------------------------
// params row (mf:420-450).  One thread per voxel.
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
k_finalize(int64_t V, int M, int maxfasc, int csf_on, int ear_on, const int32_t *Kv,
           const uint8_t *csf, const uint8_t *ear, const double *y, const double *w,
           const int32_t *idx_sub, const double *obj, const double *y_rec, double *params)
{
    int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= V) return;
    const int P = 1 + 2 * maxfasc + csf_on + 2 * ear_on + 2;
    double *row = params + v * P;
    for (int i = 0; i < P; i++) row[i] = 0.0;
    int K = Kv[v];
    K = K < 0 ? 0 : (K > maxfasc ? maxfasc : K);
    const int c = csf ? (csf[v] != 0) : 0, e = ear ? (ear[v] != 0) : 0;
    const int nb = K + c + e;
    if (nb == 0) return;  // mf:387-388
    const double *wv = w + v * kMaxBlocks;
    const int32_t *iv = idx_sub + v * kMaxBlocks;
    double M0 = 0.0;  // np.sum of <= 5 values: left to right
    for (int b = 0; b < nb; b++) M0 = DA(M0, wv[b]);
    double nu[kMaxBlocks];
    for (int b = 0; b < nb; b++) nu[b] = fabs(M0) > 0 ? DD(wv[b], M0) : wv[b];
    row[0] = M0;
    for (int k = 0; k < K; k++) { row[1 + k] = nu[k]; row[1 + maxfasc + k] = (double)iv[k]; }
    if (c) row[2 * maxfasc + 1] = nu[K];
    if (e) {
        row[2 * maxfasc + csf_on + 1] = nu[K + c];
        row[2 * maxfasc + csf_on + 2] = (double)iv[K + c];
    }
    row[P - 2] = DD(obj[v], (double)M);  // mf:446
    // R2 = corrcoef(y, y_rec)[0,1]**2 if both have positive std (mf:449-450)
    if (M > 1) {
        const double *yv = y + v * M, *rv = y_rec + v * M;
        double my = 0.0, mr = 0.0;
        for (int k = 0; k < M; k++) { my += yv[k]; mr += rv[k]; }
        my /= M; mr /= M;
        double syy = 0.0, srr = 0.0, syr = 0.0;
        for (int k = 0; k < M; k++) {
            double a = yv[k] - my, b = rv[k] - mr;
            syy += a * a; srr += b * b; syr += a * b;
        }
        if (syy > 0.0 && srr > 0.0) {
            double cc = syr / (M - 1) / sqrt(syy / (M - 1)) / sqrt(srr / (M - 1));
            cc = fmin(1.0, fmax(-1.0, cc));
            row[P - 1] = cc * cc;
        }
    }
}

int launch_finalize(int64_t V, int M, int maxfasc, int csf_on, int ear_on, const int32_t *K,
                    const uint8_t *csf, const uint8_t *ear, const double *y, const double *w,
                    const int32_t *idx_sub, const double *obj, const double *y_rec, double *params,
                    cudaStream_t st)
{
    if (V == 0) return MFB_OK;
    MFB_LAUNCH(k_finalize, (unsigned)((V + 127) / 128), 128, 0, st, V, M, maxfasc, csf_on, ear_on, K,
               csf, ear, y, w, idx_sub, obj, y_rec, params);
    return MFB_OK;
}

__global__ void k_unpack(int64_t V, int nb, const double *w5, const int32_t *idx5, double *w,
                         int32_t *idx)
{
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= V * nb) return;
    int64_t v = t / nb;
    int b = (int)(t % nb);
    w[t] = w5[v * kMaxBlocks + b];
    idx[t] = idx5[v * kMaxBlocks + b];
}

int launch_unpack_solution(int64_t V, int nb, const double *w5, const int32_t *idx5, double *w,
                           int32_t *idx, cudaStream_t st)
{
    if (V == 0) return MFB_OK;
    MFB_LAUNCH(k_unpack, (unsigned)((V * nb + 255) / 256), 256, 0, st, V, nb, w5, idx5, w, idx);
    return MFB_OK;
}

}  // namespace mfb
