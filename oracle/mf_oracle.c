/*
 * mf_oracle.c -- TEST INFRASTRUCTURE, NOT A PRODUCT PATH.
 *
 * CPU restatement (plain C, strict IEEE-754 double, sequential summation, no
 * FMA contraction: build with -ffp-contract=off) of the reference's per-voxel
 * exhaustive dictionary fit.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load this file's .so.
 * The shipped package never does: its CUDA extension has no CPU fallback.
 *
 * Parity status: PINNED.  oracle/make_golden.py runs the unmodified reference
 * (imported from /root/reference in the build container) on seeded inputs and
 * on the reference's own known-answer tests, and tests/test_oracle_golden.py
 * checks this restatement against those vectors (bit-identical for the
 * solvers given the same A; a few ulp for the rotation because the reference
 * evaluates |g.u| through BLAS gemv).
 *
 * Reference citations are path:line in rensonnetg/microstructure_fingerprinting
 * (mfu = microstructure_fingerprinting/mf_utils.py, mf = .../mf.py).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ */
/* Rotation: interp_PGSE_from_multishell, fast (pre-initialised) mode   */
/* mfu:1693-1737, 1785-1840, 1921-1956 with the per-shell interpolators */
/* of init_PGSE_multishell_interp (mfu:1959-2085) flattened to a table. */
/* The lerp is scipy.interpolate.interp1d._call_linear (scipy 1.18.1):  */
/*   idx = clip(searchsorted_left(x, xn), 1, n-1)                       */
/*   y = ((xn-x_lo)/(x_hi-x_lo))*y_hi + ((x_hi-xn)/(x_hi-x_lo))*y_lo    */
/* ------------------------------------------------------------------ */

static int searchsorted_left(const double *x, int n, double v)
{
    int lo = 0, hi = n; /* first index with x[idx] >= v */
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (x[mid] < v) lo = mid + 1; else hi = mid;
    }
    return lo;
}

static void shell_weights(const double *nodes, const int32_t *off, int s,
                          double x, int *row_lo, int *row_hi,
                          double *w_lo, double *w_hi)
{
    const double *xs = nodes + off[s];
    int n = off[s + 1] - off[s];
    int j = searchsorted_left(xs, n, x);
    if (j < 1) j = 1;
    if (j > n - 1) j = n - 1;
    double x_lo = xs[j - 1], x_hi = xs[j];
    *row_lo = off[s] + j - 1;
    *row_hi = off[s] + j;
    *w_hi = (x - x_lo) / (x_hi - x_lo);
    *w_lo = (x_hi - x) / (x_hi - x_lo);
}

/* D_out is (M, ldd) row-major; columns [0,N) are written. */
void orc_rotate_multishell(int M, int N, int n_shells, const int32_t *off,
                           const double *nodes, const double *table,
                           const double *gdir, const int32_t *shell_lo,
                           const int32_t *shell_hi, const double *gw_lo,
                           const double *gw_hi, const double *u,
                           double *D_out, int ldd)
{
    (void)n_shells;
    for (int m = 0; m < M; m++) {
        /* mfu:1810 x = |g . newdir| (left-to-right, separately rounded) */
        double x = fabs(gdir[3 * m] * u[0] + gdir[3 * m + 1] * u[1]
                        + gdir[3 * m + 2] * u[2]);
        int rl, rh;
        double wl, wh;
        double *out = D_out + (size_t)m * ldd;
        shell_weights(nodes, off, shell_lo[m], x, &rl, &rh, &wl, &wh);
        const double *tl = table + (size_t)rl * N, *th = table + (size_t)rh * N;
        if (shell_hi[m] == shell_lo[m]) {
            /* identical G: mfu:1931-1935 */
            for (int j = 0; j < N; j++) out[j] = wh * th[j] + wl * tl[j];
        } else {
            /* between-shell: mfu:1938-1955 (lerp over G of two shell results) */
            int rl2, rh2;
            double wl2, wh2;
            shell_weights(nodes, off, shell_hi[m], x, &rl2, &rh2, &wl2, &wh2);
            const double *tl2 = table + (size_t)rl2 * N;
            const double *th2 = table + (size_t)rh2 * N;
            for (int j = 0; j < N; j++) {
                double d_l = wh * th[j] + wl * tl[j];
                double d_h = wh2 * th2[j] + wl2 * tl2[j];
                out[j] = gw_hi[m] * d_h + gw_lo[m] * d_l;
            }
        }
    }
}

/* ------------------------------------------------------------------ */
/* Solvers.  A is (M, lda) row-major (NumPy C order).                   */
/* ------------------------------------------------------------------ */

/* mfu:225-278 solve_exhaustive_posweights_1 */
void orc_solve_1(int M, int N, const double *A, int lda, const double *y,
                 double *w_out, int32_t *idx_out, double *obj_out)
{
    double w_nneg = 0.0, y_sq = 0.0;
    int32_t ind = 0;
    for (int k = 0; k < M; k++) y_sq += y[k] * y[k];
    double min_obj = y_sq;
    for (int i1 = 0; i1 < N; i1++) {
        double adoty = 0.0, w = 0.0, resnorm = y_sq;
        for (int k = 0; k < M; k++) adoty = adoty + A[(size_t)k * lda + i1] * y[k];
        if (adoty >= 0) {
            double asq = 0.0;
            for (int k = 0; k < M; k++)
                asq = asq + A[(size_t)k * lda + i1] * A[(size_t)k * lda + i1];
            w = adoty / asq;
            resnorm -= w * adoty;
        }
        if (resnorm < min_obj) { ind = i1; min_obj = resnorm; w_nneg = w; }
    }
    *w_out = w_nneg; *idx_out = ind; *obj_out = min_obj;
}

/* mfu:404-459 lsqnonneg_2var_opt (also inlined at mfu:331-381) */
static double lsq2(double y_sq, double A11, double A12, double A22, double Y1,
                   double Y2, double *w)
{
    double w1d = A22 * Y1 - A12 * Y2;
    double w2d = A11 * Y2 - A12 * Y1;
    double resnorm = y_sq;
    w[0] = 0.0; w[1] = 0.0;
    if (w1d > 0.0 && w2d > 0.0) {
        double Det = A11 * A22 - A12 * A12;
        w[0] = w1d / Det;
        w[1] = w2d / Det;
        resnorm = ((resnorm + w[0] * w[0] * A11) + w[1] * w[1] * A22)
                  + 2 * ((w[0] * w[1] * A12 - w[0] * Y1) - w[1] * Y2);
    } else if (w1d >= 0.0 && w2d <= 0.0) {
        if (Y1 >= 0.0) { w[0] = Y1 / A11; resnorm = resnorm - Y1 * w[0]; }
    } else if (w1d <= 0.0 && w2d >= 0.0) {
        if (Y2 >= 0.0) { w[1] = Y2 / A22; resnorm = resnorm - Y2 * w[1]; }
    } else if (w1d < 0.0 && w2d < 0.0) {
        if (Y1 > 0) { w[0] = Y1 / A11; resnorm -= Y1 * w[0]; }
        else if (Y2 > 0) { w[1] = Y2 / A22; resnorm -= Y2 * w[1]; }
    }
    return resnorm;
}

/* Gram pieces in the reference's summation order (mfu:307-325, 503-535). */
static void col_sq(int M, int N, const double *A, int lda, double *out)
{
    for (int i = 0; i < N; i++) {
        double s = 0.0;
        for (int k = 0; k < M; k++) s += A[(size_t)k * lda + i] * A[(size_t)k * lda + i];
        out[i] = s;
    }
}
static void cross(int M, int Na, int Nb, const double *A, const double *B,
                  int lda, double *out /* (Na,Nb) */)
{
    for (int i = 0; i < Na; i++)
        for (int j = 0; j < Nb; j++) {
            double s = 0.0;
            for (int k = 0; k < M; k++) s += A[(size_t)k * lda + i] * B[(size_t)k * lda + j];
            out[(size_t)i * Nb + j] = s;
        }
}
static double adoty_ysq(int M, int Nt, const double *A, int lda, const double *y,
                        double *Adoty)
{
    double y_sq = 0.0;
    for (int i = 0; i < Nt; i++) Adoty[i] = 0.0;
    for (int k = 0; k < M; k++) {
        y_sq += y[k] * y[k];
        for (int i = 0; i < Nt; i++) Adoty[i] += y[k] * A[(size_t)k * lda + i];
    }
    return y_sq;
}

/* mfu:288-392 solve_exhaustive_posweights_2 */
int orc_solve_2(int M, int N1, int N2, const double *A, int lda, const double *y,
                double *w_out, int32_t *idx_out, double *obj_out)
{
    double *A11 = malloc(sizeof(double) * N1), *A22 = malloc(sizeof(double) * N2);
    double *A12 = malloc(sizeof(double) * (size_t)N1 * N2);
    double *Ady = malloc(sizeof(double) * (N1 + N2));
    if (!A11 || !A22 || !A12 || !Ady) return -1;
    col_sq(M, N1, A, lda, A11);
    col_sq(M, N2, A + N1, lda, A22);
    cross(M, N1, N2, A, A + N1, lda, A12);
    double y_sq = adoty_ysq(M, N1 + N2, A, lda, y, Ady);
    double min_obj = y_sq, wb[2] = {0, 0};
    int32_t ib[2] = {0, 0};
    for (int i1 = 0; i1 < N1; i1++)
        for (int i2 = 0; i2 < N2; i2++) {
            double w[2];
            double res = lsq2(y_sq, A11[i1], A12[(size_t)i1 * N2 + i2], A22[i2],
                              Ady[i1], Ady[N1 + i2], w);
            if (res < min_obj) {
                ib[0] = i1; ib[1] = i2; min_obj = res; wb[0] = w[0]; wb[1] = w[1];
            }
        }
    w_out[0] = wb[0]; w_out[1] = wb[1];
    idx_out[0] = ib[0]; idx_out[1] = ib[1];
    *obj_out = min_obj;
    free(A11); free(A22); free(A12); free(Ady);
    return 0;
}

/* mfu:470-607 solve_exhaustive_posweights_3 */
int orc_solve_3(int M, int N1, int N2, int N3, const double *A, int lda,
                const double *y, double *w_out, int32_t *idx_out, double *obj_out)
{
    const double eps = 2.2204e-16, tol = 100 * eps; /* mfu:480-481 */
    double *A11 = malloc(sizeof(double) * N1), *A22 = malloc(sizeof(double) * N2);
    double *A33 = malloc(sizeof(double) * N3);
    double *A12 = malloc(sizeof(double) * (size_t)N1 * N2);
    double *A13 = malloc(sizeof(double) * (size_t)N1 * N3);
    double *A23 = malloc(sizeof(double) * (size_t)N2 * N3);
    double *Ady = malloc(sizeof(double) * (N1 + N2 + N3));
    if (!A11 || !A22 || !A33 || !A12 || !A13 || !A23 || !Ady) return -1;
    const double *B1 = A, *B2 = A + N1, *B3 = A + N1 + N2;
    col_sq(M, N1, B1, lda, A11);
    col_sq(M, N2, B2, lda, A22);
    col_sq(M, N3, B3, lda, A33);
    cross(M, N1, N2, B1, B2, lda, A12);
    cross(M, N1, N3, B1, B3, lda, A13);
    cross(M, N2, N3, B2, B3, lda, A23);
    double y_sq = adoty_ysq(M, N1 + N2 + N3, A, lda, y, Ady);
    double min_obj = y_sq, wb[3] = {0, 0, 0};
    int32_t ib[3] = {0, 0, 0};
    for (int i3 = 0; i3 < N3; i3++) {
        double a33 = A33[i3], Y3 = Ady[N1 + N2 + i3];
        for (int i1 = 0; i1 < N1; i1++) {
            double a11 = A11[i1], a13 = A13[(size_t)i1 * N3 + i3], Y1 = Ady[i1];
            for (int i2 = 0; i2 < N2; i2++) {
                double a12 = A12[(size_t)i1 * N2 + i2], a22 = A22[i2];
                double a23 = A23[(size_t)i2 * N3 + i3], Y2 = Ady[N1 + i2];
                double w[3], res;
                /* mfu:556-561, association exactly as written */
                double D1 = (Y1 * (a22 * a33 - a23 * a23) - Y2 * (a12 * a33 - a23 * a13))
                            + Y3 * (a12 * a23 - a22 * a13);
                double D2 = (-Y1 * (a12 * a33 - a13 * a23) + Y2 * (a11 * a33 - a13 * a13))
                            - Y3 * (a11 * a23 - a12 * a13);
                double D3 = (Y1 * (a12 * a23 - a13 * a22) - Y2 * (a11 * a23 - a12 * a13))
                            + Y3 * (a11 * a22 - a12 * a12);
                if (D1 >= -tol && D2 >= -tol && D3 >= -tol) {
                    double D = (a11 * (a22 * a33 - a23 * a23) - a12 * (a12 * a33 - a23 * a13))
                               + a13 * (a12 * a23 - a22 * a13);
                    w[0] = D1 / D; w[1] = D2 / D; w[2] = D3 / D;
                    res = 0.0;
                    for (int k = 0; k < M; k++) { /* mfu:569-573 direct residual */
                        double r = ((w[0] * B1[(size_t)k * lda + i1]
                                     + w[1] * B2[(size_t)k * lda + i2])
                                    + w[2] * B3[(size_t)k * lda + i3]) - y[k];
                        res += r * r;
                    }
                } else {
                    double w2[2], r2;
                    res = lsq2(y_sq, a11, a12, a22, Y1, Y2, w2);
                    w[0] = w2[0]; w[1] = w2[1]; w[2] = 0.0;
                    r2 = lsq2(y_sq, a11, a13, a33, Y1, Y3, w2);
                    if (r2 < res) { w[0] = w2[0]; w[1] = 0.0; w[2] = w2[1]; res = r2; }
                    r2 = lsq2(y_sq, a22, a23, a33, Y2, Y3, w2);
                    if (r2 < res) { w[0] = 0.0; w[1] = w2[0]; w[2] = w2[1]; res = r2; }
                }
                if (res < min_obj) {
                    ib[0] = i1; ib[1] = i2; ib[2] = i3; min_obj = res;
                    wb[0] = w[0]; wb[1] = w[1]; wb[2] = w[2];
                }
            }
        }
    }
    for (int i = 0; i < 3; i++) { w_out[i] = wb[i]; idx_out[i] = ib[i]; }
    *obj_out = min_obj;
    free(A11); free(A22); free(A33); free(A12); free(A13); free(A23); free(Ady);
    return 0;
}

/* y_recons = A[:, ind_atoms_totdic] @ w (mfu:277, 391, 606): the reference
 * uses np.dot on a gathered (M,K) copy; for K<=3 columns that is a
 * left-to-right sum per row up to BLAS' FMA usage, so callers compare y_rec
 * with a tolerance, never bitwise. */
void orc_reconstruct(int M, int nb, const double *A, int lda, const int32_t *tot_idx,
                     const double *w, double *y_rec)
{
    for (int k = 0; k < M; k++) {
        double s = 0.0;
        for (int b = 0; b < nb; b++) s += A[(size_t)k * lda + tot_idx[b]] * w[b];
        y_rec[k] = s;
    }
}

/* monte_carlo_average (mfu:2758-2812): per sequence, the spins of its reference
 * sequence are visited in order; the phase is the left-to-right sum over the
 * gradient components of separately rounded products; the signal is the running
 * sum of cos(Dscaling * phase) divided by the number of spins. */
#include <math.h>
void orc_mc_average(long long n_seq, int dim, const double *sim_phases, const long long *delta_mapping,
                    const double *gscaling, double Dscaling, long long num_spins, double *signal)
{
    for (long long iseq = 0; iseq < n_seq; iseq++) {
        const long long start = delta_mapping[iseq] * num_spins;
        double s = 0.0;
        for (long long l = 0; l < num_spins; l++) {
            double ph = 0.0;
            for (int d = 0; d < dim; d++) ph += gscaling[iseq * dim + d] * sim_phases[(start + l) * dim + d];
            s += cos(Dscaling * ph);
        }
        signal[iseq] = s / (double)num_spins;
    }
}
