/*
 * mfb200.h -- C ABI of libmfb200.so: the B200-native (sm_100a) replacement for the
 * per-voxel hot path of rensonnetg/microstructure_fingerprinting.
 *
 * What it replaces (reference path:line; mfu = microstructure_fingerprinting/
 * mf_utils.py, mf = microstructure_fingerprinting/mf.py):
 *
 *   mfb_plan_create / mfb_plan_destroy
 *       the per-study state the reference keeps in the `sm` dict handed to every
 *       voxel task (mf:955-970): the pre-initialised multi-shell interpolator
 *       (mfu:1959-2085, flattened to one lookup table), the subject scheme
 *       (mf:821-846) and the CSF / EAR columns (mf:918-925).
 *   mfb_rotate_multishell
 *       mfu.interp_PGSE_from_multishell(sch_mat, newdir, msinterp=...) in fast
 *       mode (mfu:1693-1737, 1785-1840, 1921-1956), batched over directions.
 *   mfb_solve_batch
 *       mfu.solve_exhaustive_posweights(A, y, dicsizes) (mfu:115-214) and the
 *       Numba kernels behind it (_1 mfu:225, _2 mfu:288, _3 mfu:470, _4up
 *       mfu:612), batched over voxels.
 *   mfb_mc_average
 *       mfu.monte_carlo_average (mfu:2758-2812), the spin average behind
 *       mfu.get_PGSE_from_phases (mfu:2815-3015): dictionary generation.
 *   mfb_fit
 *       the voxel loop of MFModel.fit (mf:978-1028) over mf._fit_voxel
 *       (mf:340-461): rotate, assemble, solve, M0/nu/MSE/R2, pack params row.
 *
 * Conventions: plain pointers and sizes, no C++/torch types.  Every function
 * returns 0 on success or a negative MFB_E* code and never throws; the message
 * of the last failure on the calling thread is mfb_last_error().  Pointers
 * documented "device" must be CUDA device memory on the plan's device; "host"
 * pointers are ordinary (pageable or pinned) host memory.  `stream` is a
 * cudaStream_t passed as void* (NULL = legacy default stream).  One plan per
 * GPU; calls on different plans may run from different host threads.
 * There is no CPU fallback: without a CUDA device every entry point fails.
 */
#ifndef MFB200_H
#define MFB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MFB_OK 0
#define MFB_EINVAL (-1)      /* bad argument (shape, NULL pointer, range)       */
#define MFB_ECUDA (-2)       /* CUDA runtime error (see mfb_last_error)          */
#define MFB_ENOMEM (-3)      /* device or host allocation failed                 */
#define MFB_EUNSUPPORTED (-4) /* shape outside what the kernels cover            */

typedef struct mfb_plan mfb_plan;

/* ABI version (bumped on any signature change); mfb_version() returns the value the
 * library was built with, the Python binding refuses a library whose version differs. */
#define MFB_ABI_VERSION 4
int mfb_version(void);

/* element types of the host volume handed to mfb_fit_volume */
#define MFB_F64 0
#define MFB_F32 1

/* Message of the last error raised on this thread ("" if none). */
const char *mfb_last_error(void);

/* Number of kernels launched by this library since load (all threads); used by
 * bench.py's "gpu_launches" and by the tests that prove the CUDA path ran. */
int64_t mfb_launch_count(void);

/*
 * Create the per-GPU plan.  All pointers are HOST pointers; contents are copied.
 *   M, N          measurements per voxel, atoms per fascicle sub-dictionary
 *   R, n_shells   rows of the lookup table, number of dense shells
 *   shell_row_offset[n_shells+1]   first table row of each shell
 *   nodes[R]      sorted |g.ordir| nodes of every shell, concatenated
 *   table[R*N]    row-major signal rows at those nodes
 *   gdir[M*3]     subject gradient directions
 *   shell_lo/hi[M], gw_lo/hi[M]    dense shell(s) of each measurement and the
 *                 between-shell weights (hi==lo, gw unused, when G matches a
 *                 dense shell exactly)
 *   sig_csf[M] or NULL; sig_ear[M*E] row-major or NULL (then E = 0)
 * Returns NULL on failure (see mfb_last_error).
 */
mfb_plan *mfb_plan_create(int device, int M, int N, int R, int n_shells,
                          const int32_t *shell_row_offset, const double *nodes,
                          const double *table, const double *gdir,
                          const int32_t *shell_lo, const int32_t *shell_hi,
                          const double *gw_lo, const double *gw_hi,
                          const double *sig_csf, const double *sig_ear, int E);

void mfb_plan_destroy(mfb_plan *plan);

/*
 * Rotate the fascicle dictionary along V directions.
 *   dirs   device, V*3
 *   D_out  device, V * M * ldd doubles, voxel-major then row-major (M rows of
 *          ldd >= N doubles); columns [0,N) of every row are written.
 */
int mfb_rotate_multishell(mfb_plan *plan, int64_t V, const double *dirs,
                          double *D_out, int64_t ldd, void *stream);

/*
 * Row-lerp primitive behind mfu.rotate_atom (mf_utils.py:1205-1437) and
 * mfu.rotate_atom_2Dprotocol (mf_utils.py:1440-1690): the per-shell / per-line
 * scipy interp1d calls (mf_utils.py:1423-1426, 1678-1684) evaluated for a batch of
 * directions from an explicit interpolation plan.  Device pointers:
 *   table R*N row-major; row_lo,row_hi,w_lo,w_hi V*M; scale V*M or NULL
 *   out[v][m][0..N) = scale * (w_hi*table[row_hi] + w_lo*table[row_lo]),
 *   separately rounded products and sum (scipy's two-weight form); rows of ldd doubles.
 */
int mfb_lerp_rows(int device, int64_t V, int M, int N, const double *table,
                  const int32_t *row_lo, const int32_t *row_hi, const double *w_lo,
                  const double *w_hi, const double *scale, double *out, int64_t ldd,
                  void *stream);

/*
 * Interpolation plan of rotate_atom_2Dprotocol (reference mf_utils.py:1440-1690) for V
 * directions, expanded on the device from the host's per-direction decisions; replaces the
 * per-sequence part of the reference's per-direction loop (mf_utils.py:1557-1686).  Outputs
 * feed mfb_lerp_rows.  All pointers are device pointers.
 *   per sequence m (M):   m_class (class = (Delta, delta) pair x laboratory direction of a
 *       b > 0 sequence, -1 for b0 sequences, which keep their own table row), m_lab (unique
 *       laboratory direction, < U), m_isb0, m_b0row (table row of the pair's mean b0 signal),
 *       m_G, m_gd = gamma*delta, m_tt = Delta - delta/3; DIFF = free diffusivity
 *   per direction v:      nrm, gz (V x U: in-plane norm and |g_z| of the rotated laboratory
 *       directions), ok (V: 0 = the direction breaks the protocol's assumptions, scale = 0);
 *       kind, line, sgn (V x C: 0 = no rule reaches the class -> zero signal, 2 = gradient
 *       parallel to the fascicle -> b0 row, 3 = interpolate along reference line `line` with
 *       sign `sgn` of the perpendicular gradient)
 *   reference lines:      line_off (L+1), line_nodes (sorted signed G), line_rows (table rows)
 *   outputs (V x M each): row_lo, row_hi, w_lo, w_hi (scipy's two-weight linear form, interval
 *       index clipped to [1, n-1]), scale = exp(-(gamma delta |g_z| G)^2 (Delta - delta/3) DIFF)
 */
int mfb_plan2d(int device, int64_t V, int M, int U, int C, const int32_t *m_class,
               const int32_t *m_lab, const uint8_t *m_isb0, const int32_t *m_b0row,
               const double *m_G, const double *m_gd, const double *m_tt, double DIFF,
               const double *nrm, const double *gz, const uint8_t *kind, const int32_t *line,
               const double *sgn, const uint8_t *ok, const int32_t *line_off,
               const double *line_nodes, const int32_t *line_rows, int32_t *row_lo,
               int32_t *row_hi, double *w_lo, double *w_hi, double *scale, void *stream);

/*
 * Batched solve_exhaustive_posweights on explicit dictionaries.
 *   nblocks in [1,5]; sizes[nblocks] (host) > 0, Ntot = sum(sizes)
 *   A       device; voxel v's dictionary is the row-major (M, lda) matrix at
 *           A + v*strideA (strideA = 0 shares one dictionary between voxels)
 *   y       device, V*M
 * Outputs (device): w V*nblocks, idx_sub V*nblocks (index inside each block),
 *   min_obj V, y_rec V*M (may be NULL).
 * 1-3 blocks reproduce the reference's arithmetic exactly (summation order,
 * no FMA, loop-order tie-break); 4-5 blocks use closed-form support
 * enumeration, which agrees with scipy.optimize.nnls to rounding.
 * flags: bit 0 = force the exact (reference-order) tier for every voxel (verification;
 * results are the same by construction).
 */
int mfb_solve_batch(int device, int64_t V, int M, int nblocks,
                    const int64_t *sizes, const double *A, int64_t lda,
                    int64_t strideA, const double *y, double *w,
                    int32_t *idx_sub, double *min_obj, double *y_rec,
                    int flags, void *stream);

/*
 * Monte-Carlo signal synthesis, mfu.monte_carlo_average (mf_utils.py:2758-2812):
 *   signal[i] = (1/num_spins) * sum_l cos(Dscaling * sum_d gscaling[i,d] *
 *               sim_phases[(delta_mapping[i]*num_spins + l)*dim + d])
 * Device pointers: sim_phases n_entries*dim (row-major), delta_mapping n_seq (int64),
 * gscaling n_seq*dim, signal n_seq (output).  dim in [1,8].  A delta_mapping entry that
 * points outside sim_phases yields NaN for that sequence.
 */
int mfb_mc_average(int device, int64_t n_entries, int dim, const double *sim_phases,
                   int64_t n_seq, const int64_t *delta_mapping, const double *gscaling,
                   double Dscaling, int64_t num_spins, double *signal, void *stream);

/* mfb_solve_batch keeps its device workspace between calls; mfb_trim frees the workspace of
 * `device` (a later call allocates it again). */
int mfb_trim(int device);

/* Process-wide counters of mfb_solve_batch, out[6]: voxels decided by the screening tier,
 * voxels redone by the reference-order tier, and why they were handed over ([2] no
 * candidate, [3] ill-conditioned competitor, [4] near tie, [5] a solution with fewer active
 * columns was competitive).  reset != 0 clears them after reading. */
int mfb_solve_stats(int64_t *out, int n, int reset);

/*
 * The voxel loop of MFModel.fit.  Device pointers:
 *   y V*M, peaks V*3*maxfasc, K V (0..maxfasc), csf V, ear V (0/1 bytes)
 *   params_out V*P with P = 1 + 2*maxfasc + csf_on + 2*ear_on + 2, row layout of
 *   mf:375-381: [M0 | nu_k | ID_k | nu_csf | nu_ear ID_ear | MSE | R2].
 * flags: bit 0 = force the exact (reference-order) tier for every voxel;
 *        bit 1 = bracket the dominant kernel with CUDA events (mfb_fit_stats).
 */
int mfb_fit(mfb_plan *plan, int64_t V, const double *y, const double *peaks,
            const int32_t *K, const uint8_t *csf, const uint8_t *ear,
            int maxfasc, int csf_on, int ear_on, double *params_out,
            int flags, void *stream);

/* Same with HOST buffers (pageable or page-locked): chunks of voxels are staged into the
 * plan's page-locked slots by a helper thread, uploaded, searched and downloaded on three
 * streams so that staging, H2D, compute and D2H overlap.  This is the call a non-PyTorch
 * host binds.  Returns when params_out is complete. */
int mfb_fit_host(mfb_plan *plan, int64_t V, const double *y, const double *peaks,
                 const int32_t *K, const uint8_t *csf, const uint8_t *ear,
                 int maxfasc, int csf_on, int ear_on, double *params_out,
                 int flags);

/*
 * The voxel loop of MFModel.fit over ROI voxels scattered in a host volume: replaces
 * `data_arr[mask > 0]` (mf:644) + the loop (mf:978-1028) without materialising the
 * (ROI_size, M) copy.  Measurement m of voxel v is element
 *     voxel_offset[v] + m * meas_stride        (voxel_offset != NULL)
 *     v * row_stride + m * meas_stride         (voxel_offset == NULL)
 * of the host array `data` of element type `dtype` (MFB_F64 / MFB_F32; float32 volumes are
 * widened on the fly).  The gather runs in the library's helper thread, overlapped with the
 * GPU.  peaks / K / csf / ear / params_out are host arrays in ROI order as for mfb_fit_host.
 */
int mfb_fit_volume(mfb_plan *plan, int64_t V, const void *data, int dtype,
                   const int64_t *voxel_offset, int64_t row_stride, int64_t meas_stride,
                   const double *peaks, const int32_t *K, const uint8_t *csf,
                   const uint8_t *ear, int maxfasc, int csf_on, int ear_on,
                   double *params_out, int flags);

/* Per-plan counters of the last mfb_fit / mfb_fit_host call, out[8]:
 *   [0] voxels accepted by the fast (DMMA screening) tier, [1] voxels solved by the exact
 *   tier, [2] ms in the dominant kernel (flags bit 1), [3] its launches, [4] voxels it
 *   covered, [5] one-fascicle voxels (fused kernel), [6] fast-tier hand-overs caused by an
 *   ill-conditioned competitor, [7] by a near tie. */
int mfb_fit_stats(mfb_plan *plan, double *out, int n);

#ifdef __cplusplus
}
#endif
#endif /* MFB200_H */
