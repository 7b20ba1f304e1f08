// Shared declarations of libmfb200 (sm_100a).  Internal header, not part of the C ABI.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdio>
#include <string>

#include "../../include/mfb200.h"

namespace mfb {

extern std::atomic<long long> g_launches;
void set_error(const std::string &msg);

#define MFB_CUDA_TRY(expr)                                                              \
    do {                                                                                \
        cudaError_t e__ = (expr);                                                       \
        if (e__ != cudaSuccess) {                                                       \
            ::mfb::set_error(std::string(#expr) + ": " + cudaGetErrorString(e__) +      \
                             " (" + __FILE__ + ":" + std::to_string(__LINE__) + ")");   \
            return MFB_ECUDA;                                                           \
        }                                                                               \
    } while (0)

#define MFB_TRY(expr)                                                                   \
    do {                                                                                \
        int rc__ = (expr);                                                              \
        if (rc__ != MFB_OK) return rc__;                                                \
    } while (0)

// Count every kernel launch (mfb_launch_count) and check the launch itself.
#define MFB_LAUNCH(kernel, grid, block, smem, stream, ...)                              \
    do {                                                                                \
        kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);                     \
        ::mfb::g_launches.fetch_add(1, std::memory_order_relaxed);                      \
        MFB_CUDA_TRY(cudaGetLastError());                                               \
    } while (0)

// ---------------------------------------------------------------------------------
// Which sub-dictionaries a voxel's dictionary is made of (mf._fit_voxel, mf:371-408):
// K fascicle blocks of N atoms, then CSF (1 column), then EAR (E columns).
// ---------------------------------------------------------------------------------
struct BlockSpec {
    int nb;        // number of blocks (1..5)
    int size[5];   // atoms per block
    int start[5];  // first column of each block in the assembled dictionary
    int ntot;      // total number of columns
};

// Rotation lookup table and subject scheme resident on the device (one per plan).
struct DevPlan {
    int M, N, R, n_shells, E;
    int has_between;            // some measurement lies between two dense shells
    const int32_t *off;         // [n_shells+1]
    const double *nodes;        // [R]
    const double *table;        // [R*N]
    const double *gdir;         // [M*3]
    const int32_t *shell_lo;    // [M]
    const int32_t *shell_hi;    // [M]
    const double *gw_lo;        // [M]
    const double *gw_hi;        // [M]
    const double *sig_csf;      // [M] or null
    const double *sig_ear;      // [M*E] or null
};

// Loop index (reference loop order) of the winning tuple, or kNoTuple when the
// all-zero solution wins (min_obj stays y_sq, mfu:249/327/537).
static const long long kNoTuple = -1;

// Width of the compact per-voxel matrix holding the winning tuple's columns.
static const int kMaxBlocks = 5;

// ------------------------- exact tier (exact.cu) ----------------------------------
// Fascicle blocks k < K rotated along peaks[v, 3k:3k+3]; CSF / EAR columns appended.
// vox_list (may be null = identity) maps the local voxel to its row in peaks.
int launch_rotate_assemble(const DevPlan &p, int64_t nvox, const int32_t *vox_list,
                           const double *peaks, int peaks_ld, int K, int csf, int ear,
                           double *A, int64_t lda, int64_t strideA, cudaStream_t st);

int launch_lerp_rows(int64_t V, int M, int N, const double *table, const int32_t *row_lo,
                     const int32_t *row_hi, const double *w_lo, const double *w_hi,
                     const double *scale, double *out, int64_t ldd, cudaStream_t st);

int launch_plan2d(int64_t V, int M, int U, int C, const int32_t *m_class, const int32_t *m_lab,
                  const uint8_t *m_isb0, const int32_t *m_b0row, const double *m_G, const double *m_gd,
                  const double *m_tt, double DIFF, const double *nrm, const double *gz, const uint8_t *kind,
                  const int32_t *line, const double *sgn, const uint8_t *ok, const int32_t *line_off,
                  const double *line_nodes, const int32_t *line_rows, int32_t *row_lo, int32_t *row_hi,
                  double *w_lo, double *w_hi, double *scale, cudaStream_t st);

size_t exact_scratch_bytes(int64_t V, const BlockSpec &bs);

// Exhaustive search in the reference's arithmetic on explicit dictionaries.
// Voxel v reads A + v*strideA and y + row(v)*y_ld with row(v) = vox_list ? vox_list[v] : v;
// the winning tuple's loop index is written to tuple_out[row(v)].
int launch_exact_search(int64_t V, int M, const BlockSpec &bs, const double *A, int64_t lda,
                        int64_t strideA, const double *y, int64_t y_ld,
                        const int32_t *vox_list, void *scratch, long long *tuple_out,
                        cudaStream_t st, cudaEvent_t *ev = nullptr, const int32_t *a_list = nullptr,
                        const uint8_t *tile_mask = nullptr, int mask_ld = 0);
inline int exact_mask_ld(const BlockSpec &bs)
{
    const int n = bs.size[0] > bs.size[1] ? bs.size[0] : bs.size[1];
    return (n + 63) / 64 * 64;     // one byte per atom, padded to the 64-atom tiles of k_pairs
}

// Copy the winning tuple's columns into Asmall[row(v)] (M x kMaxBlocks, row-major) and
// decode the per-block indices into idx_sub[row(v)*kMaxBlocks + b].
int launch_gather_from_A(int64_t V, int M, const BlockSpec &bs, const double *A, int64_t lda,
                         int64_t strideA, const long long *tuple, const int32_t *vox_list,
                         double *Asmall, int32_t *idx_sub, cudaStream_t st);

// Same, but rotating the selected atoms straight from the lookup table (fast tier).
int launch_gather_from_table(const DevPlan &p, int64_t V, int K, int csf, int ear,
                             const double *peaks, int peaks_ld, const long long *tuple,
                             const int32_t *vox_list, double *Asmall, int32_t *idx_sub,
                             cudaStream_t st);

// Evaluate the tuple in the reference's arithmetic -> w[row*kMaxBlocks+b], obj, y_rec.
// nb_fixed > 0: every voxel has nb_fixed blocks; else nb per voxel = nbv[row].
int launch_evaluate(int64_t V, int M, int nb_fixed, const uint8_t *nbv, const double *Asmall,
                    const double *y, int64_t y_ld, const long long *tuple, double *w,
                    double *obj, double *y_rec, int32_t *idx_sub, cudaStream_t st);

// Pack the params row (mf:420-450).
int launch_finalize(int64_t V, int M, int maxfasc, int csf_on, int ear_on, const int32_t *K,
                    const uint8_t *csf, const uint8_t *ear, const double *y,
                    const double *w, const int32_t *idx_sub, const double *obj,
                    const double *y_rec, double *params, cudaStream_t st);

// Voxel type code K + 3*csf + 6*ear and per-type index lists.
int launch_classify(int64_t V, const int32_t *K, const uint8_t *csf, const uint8_t *ear,
                    int maxfasc, uint8_t *type, uint8_t *nbv, int32_t *lists /*12*V*/,
                    int32_t *counts /*12*/, cudaStream_t st);

// One-fascicle voxels: rotation + Gram terms + closed forms fused (reference arithmetic).
bool single_fascicle_supported(const DevPlan &p, int K, int csf, int ear);
int launch_single_fascicle(const DevPlan &p, int64_t nvox, const int32_t *vox_list,
                           const double *peaks, int peaks_ld, const double *y, int csf, int ear,
                           long long *tuple, cudaStream_t st);

// ------------------------- fast tier (fast.cu) -------------------------------------
// DMMA screening for 2-fascicle voxels ([N,N] and [N,N,1]); voxels whose winner is not
// certain are appended to redo_list (exact tier).
struct FastProblem {
    int src;            // 0: rotate from the plan's lookup table; 1: explicit dictionaries
    int N1, N2;         // atoms of the two searched blocks
    const double *A;    // explicit: voxel row r reads A + r*strideA, (M, lda) row-major
    int64_t lda, strideA;
    int start1, start2, start3;
    int csf;            // a third, single-column block is present
    int a_by_local;     // explicit + vox_list: local voxel v reads A + v*strideA (vox_list maps y / tuple rows only)
    int32_t *redo_local;  // optional: local indices of the voxels handed to the exact tier
    uint8_t *redo_mask;   // optional: [redo position][2][mask_ld] atoms of block 1 / block 2 whose rows / columns
    int mask_ld;          // the exact tier must scan for that voxel (see k_fast_select)
};
bool fast_supported(const DevPlan &p, int K, int csf, int ear);
bool fast_supported_explicit(int M, const BlockSpec &bs);
// fit path through materialised dictionaries (between-shell protocols, M > 112)
bool fast_supported_materialised(const DevPlan &p, int K, int csf, int ear);
size_t fast_scratch_bytes(int M, int N1, int N2, int64_t V, int src, int shared_dict);
int launch_fast_search(const DevPlan &p, const FastProblem &fp, int64_t V, const int32_t *vox_list,
                       const double *peaks, int peaks_ld, const double *y, void *scratch,
                       long long *tuple, int32_t *redo_list, int32_t *redo_count, int32_t *reasons,
                       cudaStream_t st, cudaEvent_t *ev);

// Triple scan ([N1, N2, N3], reference `_3`) on explicit dictionaries: DMMA correlation
// matrices + FP64 closed-form enumeration; uncertain voxels are appended to redo_list.
bool fast3_supported_explicit(int M, const BlockSpec &bs);
size_t fast3_scratch_bytes(int M, const BlockSpec &bs, int64_t V, int shared_dict);
int launch_fast_search3(int M, const BlockSpec &bs, const double *A, int64_t lda, int64_t strideA,
                        int64_t V, const double *y, void *scratch, long long *tuple,
                        int32_t *redo_list, int32_t *redo_count, int32_t *reasons, cudaStream_t st,
                        cudaEvent_t *ev, const int32_t *vox_list = nullptr, int a_by_local = 0,
                        int32_t *redo_local = nullptr, int csf_col = -1);
// csf_col >= 0: the three searched blocks of `bs` are projected off that single column of A (two
// fascicles + CSF + EAR, reference `_4up`); tuples are returned in the product loop order of the
// four blocks [N1, N2, 1, N3].
bool fast3_supported_materialised(const DevPlan &p, int K, int csf, int ear);

// ------------------------- Monte-Carlo average (mc.cu) ------------------------------
int mc_nsplit(int64_t n_seq, int64_t num_spins);
int launch_mc_average(int64_t n_entries, int dim, const double *phases, int64_t n_seq,
                      const long long *delta_mapping, const double *gscaling, double Dscaling,
                      int64_t num_spins, int nsplit, double *partial, double *signal, cudaStream_t st);

// solve_batch helpers
int launch_unpack_solution(int64_t V, int nb, const double *w5, const int32_t *idx5,
                           double *w, int32_t *idx, cudaStream_t st);

}  // namespace mfb
