// api.cu -- C ABI of libmfb200 (see include/mfb200.h) and the host-side orchestration of
// the voxel loop: classify voxels by dictionary composition, run the fast (DMMA screening)
// tier where it applies and the exact (reference-order) tier elsewhere, evaluate the
// winning tuple in the reference's arithmetic, pack the params rows.
#include <algorithm>
#include <condition_variable>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

#include "common.cuh"

namespace mfb {

std::atomic<long long> g_launches{0};
static thread_local std::string t_error;
void set_error(const std::string &msg) { t_error = msg; }

// growable device buffer
struct Buf {
    void *p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes)
    {
        if (bytes <= cap) return MFB_OK;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 8;
        if (cudaMalloc(&p, want) != cudaSuccess) {
            cudaGetLastError();
            if (cudaMalloc(&p, bytes) != cudaSuccess) {
                cudaGetLastError();
                set_error("device allocation of " + std::to_string(bytes) + " bytes failed");
                return MFB_ENOMEM;
            }
            want = bytes;
        }
        cap = want;
        return MFB_OK;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <typename T> T *as() { return (T *)p; }
};

// page-locked host buffer (staging slots of the host pipeline)
struct PinBuf {
    void *p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes)
    {
        if (bytes <= cap) return MFB_OK;
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) {
            cudaGetLastError();
            p = nullptr;
            set_error("page-locked host allocation of " + std::to_string(bytes) + " bytes failed");
            return MFB_ENOMEM;
        }
        cap = bytes;
        return MFB_OK;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

// Every entry point runs on its plan's (or the requested) device and puts the caller's
// current device back on return: a library call must not move the host thread to another GPU.
struct DeviceGuard {
    int prev = -1;
    cudaError_t err = cudaSuccess;
    explicit DeviceGuard(int device)
    {
        if (cudaGetDevice(&prev) != cudaSuccess) { cudaGetLastError(); prev = -1; }
        if (prev != device) err = cudaSetDevice(device); else prev = -1;
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};
#define MFB_ON_DEVICE(dev)                 \
    ::mfb::DeviceGuard guard__(dev);       \
    MFB_CUDA_TRY(guard__.err)

static BlockSpec make_spec(int nb, const int64_t *sizes)
{
    BlockSpec bs;
    memset(&bs, 0, sizeof(bs));
    bs.nb = nb;
    int s = 0;
    for (int b = 0; b < nb; b++) { bs.size[b] = (int)sizes[b]; bs.start[b] = s; s += (int)sizes[b]; }
    bs.ntot = s;
    return bs;
}

static BlockSpec fit_spec(int N, int E, int K, int csf, int ear)
{
    int64_t sizes[5];
    int nb = 0;
    for (int k = 0; k < K; k++) sizes[nb++] = N;
    if (csf) sizes[nb++] = 1;
    if (ear) sizes[nb++] = E;
    return make_spec(nb, sizes);
}

}  // namespace mfb

using namespace mfb;

struct mfb_plan {
    int device = 0;
    DevPlan dp;
    std::vector<void *> owned;
    // per-chunk workspace
    Buf type, nbv, lists, counts, tuple, asmall, idx5, w5, obj, yrec, abuf, scratch, fscratch, redo, redomask;
    // host pipeline of mfb_fit_host / mfb_fit_volume: page-locked and device slots, streams, events
    PinBuf h_in[3], h_out[2];
    Buf d_in[2], d_out[2];
    cudaStream_t s_copy = nullptr, s_d2h = nullptr;
    cudaEvent_t pipe_ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    cudaStream_t stream = nullptr;
    std::vector<cudaEvent_t> events;  // pairs bracketing the dominant kernel (flags bit 1)
    size_t events_used = 0;
    double stats[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    size_t exact_budget = (size_t)3 << 30;  // bytes of materialised dictionaries per sub-chunk
};

// per-device workspace of mfb_solve_batch
static const int kMaxDevices = 64;
struct SolveCache {
    std::mutex mu;
    Buf scratch, tuple, asmall, idx5, w5, redo, redomask;
};
static SolveCache g_solve_cache[kMaxDevices];

extern "C" int mfb_trim(int device)
{
    if (device < 0 || device >= kMaxDevices) { set_error("mfb_trim: device index out of range"); return MFB_EINVAL; }
    SolveCache &c = g_solve_cache[device];
    std::lock_guard<std::mutex> lock(c.mu);
    if (c.scratch.p || c.tuple.p) {
        MFB_ON_DEVICE(device);
        Buf *bufs[] = {&c.scratch, &c.tuple, &c.asmall, &c.idx5, &c.w5, &c.redo, &c.redomask};
        for (Buf *b : bufs) b->release();
    }
    return MFB_OK;
}

// counters of the mfb_solve_batch calls of this process: voxels decided by the screening
// tier, voxels redone in reference order, hand-over reasons (no candidate, ill-conditioned
// competitor, near tie, branch with fewer active columns)
static std::atomic<long long> g_solve_stats[6];

extern "C" int mfb_version(void) { return MFB_ABI_VERSION; }
extern "C" int mfb_solve_stats(int64_t *out, int n, int reset)
{
    if (!out && n > 0) { set_error("mfb_solve_stats: invalid argument"); return MFB_EINVAL; }
    for (int i = 0; i < n && i < 6; i++) out[i] = (int64_t)g_solve_stats[i].load();
    if (reset) for (int i = 0; i < 6; i++) g_solve_stats[i] = 0;
    return MFB_OK;
}
extern "C" const char *mfb_last_error(void) { return t_error.c_str(); }
extern "C" int64_t mfb_launch_count(void) { return (int64_t)g_launches.load(); }

template <typename T>
static int upload(mfb_plan *pl, const T *host, size_t n, const T **dev)
{
    void *d = nullptr;
    MFB_CUDA_TRY(cudaMalloc(&d, std::max<size_t>(n, 1) * sizeof(T)));
    pl->owned.push_back(d);
    if (n) MFB_CUDA_TRY(cudaMemcpy(d, host, n * sizeof(T), cudaMemcpyHostToDevice));
    *dev = (const T *)d;
    return MFB_OK;
}

static int plan_build(mfb_plan *pl, int device, int M, int N, int R, int n_shells,
                      const int32_t *off, const double *nodes, const double *table,
                      const double *gdir, const int32_t *shell_lo, const int32_t *shell_hi,
                      const double *gw_lo, const double *gw_hi, const double *sig_csf,
                      const double *sig_ear, int E)
{
    cudaDeviceProp prop;
    MFB_CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) {
        set_error(std::string("libmfb200 is built for sm_100a only; device is ") + prop.name);
        return MFB_EUNSUPPORTED;
    }
    pl->device = device;
    DevPlan &dp = pl->dp;
    memset(&dp, 0, sizeof(dp));
    dp.M = M; dp.N = N; dp.R = R; dp.n_shells = n_shells; dp.E = sig_ear ? E : 0;
    dp.has_between = 0;
    for (int m = 0; m < M; m++) {
        if (shell_lo[m] < 0 || shell_lo[m] >= n_shells || shell_hi[m] < 0 || shell_hi[m] >= n_shells) {
            set_error("shell index out of range");
            return MFB_EINVAL;
        }
        if (shell_hi[m] != shell_lo[m]) dp.has_between = 1;
    }
    for (int s = 0; s < n_shells; s++)
        if (off[s + 1] - off[s] < 2 || off[s + 1] > R) {
            set_error("every shell needs at least 2 nodes");
            return MFB_EINVAL;
        }
    MFB_TRY(upload(pl, off, (size_t)n_shells + 1, &dp.off));
    MFB_TRY(upload(pl, nodes, (size_t)R, &dp.nodes));
    MFB_TRY(upload(pl, table, (size_t)R * N, &dp.table));
    MFB_TRY(upload(pl, gdir, (size_t)M * 3, &dp.gdir));
    MFB_TRY(upload(pl, shell_lo, (size_t)M, &dp.shell_lo));
    MFB_TRY(upload(pl, shell_hi, (size_t)M, &dp.shell_hi));
    MFB_TRY(upload(pl, gw_lo, (size_t)M, &dp.gw_lo));
    MFB_TRY(upload(pl, gw_hi, (size_t)M, &dp.gw_hi));
    if (sig_csf) MFB_TRY(upload(pl, sig_csf, (size_t)M, &dp.sig_csf));
    if (sig_ear) MFB_TRY(upload(pl, sig_ear, (size_t)M * E, &dp.sig_ear));
    MFB_CUDA_TRY(cudaStreamCreateWithFlags(&pl->stream, cudaStreamNonBlocking));
    return MFB_OK;
}

extern "C" mfb_plan *mfb_plan_create(int device, int M, int N, int R, int n_shells,
                                     const int32_t *shell_row_offset, const double *nodes,
                                     const double *table, const double *gdir,
                                     const int32_t *shell_lo, const int32_t *shell_hi,
                                     const double *gw_lo, const double *gw_hi,
                                     const double *sig_csf, const double *sig_ear, int E)
{
    if (M <= 0 || N <= 0 || R <= 1 || n_shells <= 0 || !shell_row_offset || !nodes || !table ||
        !gdir || !shell_lo || !shell_hi || !gw_lo || !gw_hi || (sig_ear && E <= 0)) {
        set_error("mfb_plan_create: invalid argument");
        return nullptr;
    }
    DeviceGuard guard(device);
    if (guard.err != cudaSuccess) {
        set_error(std::string("mfb_plan_create: cudaSetDevice: ") + cudaGetErrorString(guard.err));
        cudaGetLastError();
        return nullptr;
    }
    mfb_plan *pl = new (std::nothrow) mfb_plan();
    if (!pl) { set_error("out of host memory"); return nullptr; }
    int rc = plan_build(pl, device, M, N, R, n_shells, shell_row_offset, nodes, table, gdir,
                        shell_lo, shell_hi, gw_lo, gw_hi, sig_csf, sig_ear, E);
    if (rc != MFB_OK) { mfb_plan_destroy(pl); return nullptr; }
    return pl;
}

extern "C" void mfb_plan_destroy(mfb_plan *pl)
{
    if (!pl) return;
    DeviceGuard guard(pl->device);
    for (void *p : pl->owned) cudaFree(p);
    Buf *bufs[] = {&pl->type, &pl->nbv, &pl->lists, &pl->counts, &pl->tuple, &pl->asmall,
                   &pl->idx5, &pl->w5, &pl->obj, &pl->yrec, &pl->abuf, &pl->scratch, &pl->fscratch, &pl->redo, &pl->redomask,
                   &pl->d_in[0], &pl->d_in[1], &pl->d_out[0], &pl->d_out[1]};
    for (Buf *b : bufs) b->release();
    for (PinBuf &b : pl->h_in) b.release();
    for (PinBuf &b : pl->h_out) b.release();
    for (cudaEvent_t e : pl->pipe_ev) if (e) cudaEventDestroy(e);
    if (pl->s_copy) cudaStreamDestroy(pl->s_copy);
    if (pl->s_d2h) cudaStreamDestroy(pl->s_d2h);
    for (cudaEvent_t e : pl->events) cudaEventDestroy(e);
    if (pl->stream) cudaStreamDestroy(pl->stream);
    delete pl;
}

extern "C" int mfb_rotate_multishell(mfb_plan *pl, int64_t V, const double *dirs, double *D_out,
                                     int64_t ldd, void *stream)
{
    if (!pl || V < 0 || (V > 0 && (!dirs || !D_out)) || ldd < pl->dp.N) {
        set_error("mfb_rotate_multishell: invalid argument");
        return MFB_EINVAL;
    }
    MFB_ON_DEVICE(pl->device);
    return launch_rotate_assemble(pl->dp, V, nullptr, dirs, 3, 1, 0, 0, D_out, ldd,
                                  (int64_t)pl->dp.M * ldd, (cudaStream_t)stream);
}

extern "C" int mfb_lerp_rows(int device, int64_t V, int M, int N, const double *table,
                             const int32_t *row_lo, const int32_t *row_hi, const double *w_lo,
                             const double *w_hi, const double *scale, double *out, int64_t ldd,
                             void *stream)
{
    if (V < 0 || M < 0 || N <= 0 || ldd < N || (V > 0 && M > 0 && (!table || !row_lo || !row_hi || !w_lo || !w_hi || !out))) {
        set_error("mfb_lerp_rows: invalid argument");
        return MFB_EINVAL;
    }
    MFB_ON_DEVICE(device);
    return launch_lerp_rows(V, M, N, table, row_lo, row_hi, w_lo, w_hi, scale, out, ldd, (cudaStream_t)stream);
}

extern "C" int mfb_plan2d(int device, int64_t V, int M, int U, int C, const int32_t *m_class,
                          const int32_t *m_lab, const uint8_t *m_isb0, const int32_t *m_b0row,
                          const double *m_G, const double *m_gd, const double *m_tt, double DIFF,
                          const double *nrm, const double *gz, const uint8_t *kind, const int32_t *line,
                          const double *sgn, const uint8_t *ok, const int32_t *line_off,
                          const double *line_nodes, const int32_t *line_rows, int32_t *row_lo,
                          int32_t *row_hi, double *w_lo, double *w_hi, double *scale, void *stream)
{
    if (V < 0 || M <= 0 || U <= 0 || C < 0 ||
        (V > 0 && (!m_class || !m_lab || !m_isb0 || !m_b0row || !m_G || !m_gd || !m_tt || !nrm || !gz || !ok ||
                   !line_off || !row_lo || !row_hi || !w_lo || !w_hi || !scale)) ||
        (V > 0 && C > 0 && (!kind || !line || !sgn || !line_nodes || !line_rows))) {
        set_error("mfb_plan2d: invalid argument");
        return MFB_EINVAL;
    }
    MFB_ON_DEVICE(device);
    return launch_plan2d(V, M, U, C, m_class, m_lab, m_isb0, m_b0row, m_G, m_gd, m_tt, DIFF, nrm, gz, kind, line,
                         sgn, ok, line_off, line_nodes, line_rows, row_lo, row_hi, w_lo, w_hi, scale,
                         (cudaStream_t)stream);
}

extern "C" int mfb_mc_average(int device, int64_t n_entries, int dim, const double *sim_phases,
                              int64_t n_seq, const int64_t *delta_mapping, const double *gscaling,
                              double Dscaling, int64_t num_spins, double *signal, void *stream)
{
    if (n_entries < 0 || dim < 1 || dim > 8 || n_seq < 0 || num_spins <= 0 ||
        (n_seq > 0 && (!sim_phases || !delta_mapping || !gscaling || !signal))) {
        set_error("mfb_mc_average: invalid argument");
        return MFB_EINVAL;
    }
    if (n_seq == 0) return MFB_OK;
    MFB_ON_DEVICE(device);
    cudaStream_t st = (cudaStream_t)stream;
    const int nsplit = mc_nsplit(n_seq, num_spins);
    double *partial = nullptr;
    MFB_CUDA_TRY(cudaMallocAsync((void **)&partial, sizeof(double) * n_seq * nsplit, st));
    int rc = launch_mc_average(n_entries, dim, sim_phases, n_seq, (const long long *)delta_mapping, gscaling,
                               Dscaling, num_spins, nsplit, partial, signal, st);
    cudaFreeAsync(partial, st);
    return rc;
}

// ---------------------------------------------------------------------------------
// batched solve on explicit dictionaries
// ---------------------------------------------------------------------------------
extern "C" int mfb_solve_batch(int device, int64_t V, int M, int nblocks, const int64_t *sizes,
                               const double *A, int64_t lda, int64_t strideA, const double *y,
                               double *w, int32_t *idx_sub, double *min_obj, double *y_rec,
                               int flags, void *stream)
{
    if (V < 0 || M <= 0 || nblocks < 1 || nblocks > 5 || !sizes || (V > 0 && (!A || !y || !w || !idx_sub || !min_obj))) {
        set_error("mfb_solve_batch: invalid argument");
        return MFB_EINVAL;
    }
    for (int b = 0; b < nblocks; b++)
        if (sizes[b] <= 0) { set_error("mfb_solve_batch: sizes must be > 0"); return MFB_EINVAL; }
    BlockSpec bs = make_spec(nblocks, sizes);
    if (lda < bs.ntot) { set_error("mfb_solve_batch: lda < sum(sizes)"); return MFB_EINVAL; }
    if (device < 0 || device >= kMaxDevices) { set_error("mfb_solve_batch: device index out of range"); return MFB_EINVAL; }
    if (V == 0) return MFB_OK;
    MFB_ON_DEVICE(device);
    cudaStream_t st = (cudaStream_t)stream;
    // sub-batches keep the scratch bounded (the general-M fast path holds a normalised copy
    // of every dictionary of the sub-batch)
    const bool no_fast = (flags & 1) != 0;
    const bool fast = fast_supported_explicit(M, bs) && !no_fast;
    // three searched blocks, or [N1, N2, 1, N4]: the triple scan on blocks 1, 2, 4 projected off the
    // single column of block 3 (two fascicles + CSF + EAR, reference `_4up`)
    BlockSpec bs3 = bs;
    int csf_col = -1;
    if (bs.nb == 4 && bs.size[2] == 1) {
        bs3.nb = 3; bs3.size[2] = bs.size[3]; bs3.start[2] = bs.start[3];
        csf_col = bs.start[2];
    }
    const bool fast3 = !fast && fast3_supported_explicit(M, bs3) && !no_fast;
    const int shared_dict = strideA == 0;
    size_t per_vox = exact_scratch_bytes(1, bs) + 4096;
    size_t fixed = 0;
    if (fast) {
        fixed = shared_dict ? fast_scratch_bytes(M, bs.size[0], bs.size[1], 0, 1, 1) : 0;
        per_vox = std::max(per_vox, fast_scratch_bytes(M, bs.size[0], bs.size[1], 1, 1, shared_dict) - fixed + 4096);
    } else if (fast3) {
        fixed = shared_dict ? fast3_scratch_bytes(M, bs3, 0, 1) : 0;
        per_vox = std::max(per_vox, fast3_scratch_bytes(M, bs3, 1, shared_dict) - fixed + 4096);
    }
    // (sub-chunks of ~1000 voxels left 10 % on the table at [800, 800]: launch tails, one preparation pass per sub-chunk)
    const size_t budget = per_vox > ((size_t)1 << 18) ? (size_t)6 << 30 : (size_t)1 << 30;
    int64_t sub = std::max<int64_t>(1, std::min<int64_t>((fast || fast3) ? 8192 : 65535, budget / per_vox));
    sub = std::min(sub, V);
    // workspace of the device, kept between calls (mfb_trim releases it): allocating and
    // freeing gigabytes per call costs milliseconds and synchronises the device
    SolveCache &cache = g_solve_cache[device];
    std::lock_guard<std::mutex> lock(cache.mu);
    Buf &scratch = cache.scratch, &tuple = cache.tuple, &asmall = cache.asmall, &idx5 = cache.idx5, &w5 = cache.w5,
        &redo = cache.redo, &redomask = cache.redomask;
    const int mask_ld = exact_mask_ld(bs);
    int rc = MFB_OK;
    auto cleanup = [&]() {};
    size_t sbytes = exact_scratch_bytes(sub, bs);
    if (fast) sbytes = std::max(sbytes, fast_scratch_bytes(M, bs.size[0], bs.size[1], sub, 1, shared_dict));
    if (fast3) sbytes = std::max(sbytes, fast3_scratch_bytes(M, bs3, sub, shared_dict));
    if ((rc = scratch.ensure(sbytes)) || (rc = tuple.ensure(sizeof(long long) * sub)) ||
        (rc = asmall.ensure(sizeof(double) * sub * M * kMaxBlocks)) ||
        (rc = idx5.ensure(sizeof(int32_t) * sub * kMaxBlocks)) || (rc = w5.ensure(sizeof(double) * sub * kMaxBlocks)) ||
        (rc = redo.ensure(sizeof(int32_t) * (sub + 8))) || (fast && (rc = redomask.ensure((size_t)sub * 2 * mask_ld)))) {
        cleanup();
        return rc;
    }
    DevPlan dummy;
    memset(&dummy, 0, sizeof(dummy));
    dummy.M = M;
    for (int64_t v0 = 0; v0 < V && rc == MFB_OK; v0 += sub) {
        int64_t nv = std::min(sub, V - v0);
        const double *Av = A + v0 * strideA;
        const double *yv = y + v0 * M;
        if (fast || fast3) {
            // DMMA screening on the explicit dictionaries; uncertain voxels fall through to the
            // reference-order search below
            int32_t *redo_count = redo.as<int32_t>(), *reasons = redo_count + 1, *redo_list = redo_count + 8;
            if (cudaMemsetAsync(redo_count, 0, 8 * sizeof(int32_t), st) != cudaSuccess ||
                cudaMemsetAsync(tuple.p, 0xff, sizeof(long long) * nv, st) != cudaSuccess) { rc = MFB_ECUDA; break; }
            if (fast) {
                FastProblem fp;
                memset(&fp, 0, sizeof(fp));
                fp.src = 1; fp.N1 = bs.size[0]; fp.N2 = bs.size[1]; fp.A = Av; fp.lda = lda; fp.strideA = strideA;
                fp.start1 = bs.start[0]; fp.start2 = bs.start[1]; fp.start3 = bs.nb == 3 ? bs.start[2] : 0;
                fp.csf = bs.nb == 3;
                fp.redo_mask = redomask.as<uint8_t>(); fp.mask_ld = mask_ld;
                rc = launch_fast_search(dummy, fp, nv, nullptr, nullptr, 0, yv, scratch.p, tuple.as<long long>(),
                                        redo_list, redo_count, reasons, st, nullptr);
            } else {
                rc = launch_fast_search3(M, bs3, Av, lda, strideA, nv, yv, scratch.p, tuple.as<long long>(),
                                         redo_list, redo_count, reasons, st, nullptr, nullptr, 0, nullptr, csf_col);
            }
            if (rc) break;
            int32_t head[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            if (cudaMemcpyAsync(head, redo_count, sizeof(head), cudaMemcpyDeviceToHost, st) != cudaSuccess ||
                cudaStreamSynchronize(st) != cudaSuccess) { rc = MFB_ECUDA; break; }
            const int32_t n_redo = head[0];
            g_solve_stats[0] += nv - n_redo; g_solve_stats[1] += n_redo;
            for (int i = 0; i < 4; i++) g_solve_stats[2 + i] += head[1 + i];
            if (n_redo > 0)
                rc = launch_exact_search(n_redo, M, bs, Av, lda, strideA, yv, M, redo_list, scratch.p,
                                         tuple.as<long long>(), st, nullptr, redo_list,
                                         fast ? redomask.as<uint8_t>() : nullptr, mask_ld);
        } else {
            rc = launch_exact_search(nv, M, bs, Av, lda, strideA, yv, M, nullptr, scratch.p,
                                     tuple.as<long long>(), st);
        }
        if (rc) break;
        rc = launch_gather_from_A(nv, M, bs, Av, lda, strideA, tuple.as<long long>(), nullptr,
                                  asmall.as<double>(), idx5.as<int32_t>(), st);
        if (rc) break;
        rc = launch_evaluate(nv, M, bs.nb, nullptr, asmall.as<double>(), yv, M, tuple.as<long long>(),
                             w5.as<double>(), min_obj + v0, y_rec ? y_rec + v0 * M : nullptr,
                             idx5.as<int32_t>(), st);
        if (rc) break;
        rc = launch_unpack_solution(nv, bs.nb, w5.as<double>(), idx5.as<int32_t>(), w + v0 * bs.nb,
                                    idx_sub + v0 * bs.nb, st);
    }
    cudaError_t e = cudaStreamSynchronize(st);
    cleanup();
    if (rc == MFB_OK && e != cudaSuccess) {
        set_error(std::string("mfb_solve_batch: ") + cudaGetErrorString(e));
        return MFB_ECUDA;
    }
    return rc;
}

// ---------------------------------------------------------------------------------
// the voxel loop of MFModel.fit
// ---------------------------------------------------------------------------------
static int fit_chunk(mfb_plan *pl, int64_t nv, const double *y, const double *peaks,
                     const int32_t *K, const uint8_t *csf, const uint8_t *ear, int maxfasc,
                     int csf_on, int ear_on, double *params, int flags, cudaStream_t st)
{
    const DevPlan &dp = pl->dp;
    const int M = dp.M;
    const int pld = 3 * maxfasc;
    MFB_TRY(pl->type.ensure(nv));
    MFB_TRY(pl->nbv.ensure(nv));
    MFB_TRY(pl->lists.ensure(sizeof(int32_t) * 12 * nv));
    MFB_TRY(pl->counts.ensure(sizeof(int32_t) * 12));
    MFB_TRY(pl->tuple.ensure(sizeof(long long) * nv));
    MFB_TRY(pl->asmall.ensure(sizeof(double) * nv * M * kMaxBlocks));
    MFB_TRY(pl->idx5.ensure(sizeof(int32_t) * nv * kMaxBlocks));
    MFB_TRY(pl->w5.ensure(sizeof(double) * nv * kMaxBlocks));
    MFB_TRY(pl->obj.ensure(sizeof(double) * nv));
    MFB_TRY(pl->yrec.ensure(sizeof(double) * nv * M));

    MFB_TRY(launch_classify(nv, K, csf, ear, maxfasc, pl->type.as<uint8_t>(), pl->nbv.as<uint8_t>(),
                            pl->lists.as<int32_t>(), pl->counts.as<int32_t>(), st));
    int32_t counts[12];
    MFB_CUDA_TRY(cudaMemcpyAsync(counts, pl->counts.p, sizeof(counts), cudaMemcpyDeviceToHost, st));
    MFB_CUDA_TRY(cudaMemsetAsync(pl->tuple.p, 0xff, sizeof(long long) * nv, st));
    MFB_CUDA_TRY(cudaMemsetAsync(pl->idx5.p, 0, sizeof(int32_t) * nv * kMaxBlocks, st));
    MFB_CUDA_TRY(cudaStreamSynchronize(st));

    for (int t = 0; t < 12; t++) {
        const int64_t cnt = counts[t];
        if (cnt == 0) continue;
        const int Kt = t % 3, ct = (t / 3) % 2, et = t / 6;
        if (Kt + ct + et == 0) continue;
        if ((ct && !dp.sig_csf) || (et && !dp.sig_ear)) {
            set_error("mfb_fit: CSF/EAR compartment requested but the plan holds no such column");
            return MFB_EINVAL;
        }
        const BlockSpec bs = fit_spec(dp.N, dp.E, Kt, ct, et);
        const int32_t *list = pl->lists.as<int32_t>() + (int64_t)t * nv;
        const bool timed = (flags & 2) && bs.nb >= 2 && Kt == 2;
        auto next_events = [&]() -> cudaEvent_t * {
            if (pl->events_used + 2 > pl->events.size()) {
                cudaEvent_t e0, e1;
                if (cudaEventCreate(&e0) != cudaSuccess || cudaEventCreate(&e1) != cudaSuccess) return nullptr;
                pl->events.push_back(e0);
                pl->events.push_back(e1);
            }
            cudaEvent_t *ev = &pl->events[pl->events_used];
            pl->events_used += 2;
            return ev;
        };
        // exact tier: materialise the dictionaries of a sub-chunk, search in reference order
        auto run_exact = [&](const int32_t *lst, int64_t n, bool time_it, const uint8_t *mask = nullptr,
                             int mask_ld = 0) -> int {
            const int64_t lda = (bs.ntot + 1) & ~(int64_t)1;
            const size_t per_vox = (size_t)M * lda * sizeof(double);
            int64_t sub = std::max<int64_t>(1, std::min<int64_t>(65535, pl->exact_budget / per_vox));
            sub = std::min(sub, n);
            MFB_TRY(pl->abuf.ensure(per_vox * sub));
            MFB_TRY(pl->scratch.ensure(exact_scratch_bytes(sub, bs)));
            for (int64_t s0 = 0; s0 < n; s0 += sub) {
                const int64_t ns = std::min(sub, n - s0);
                MFB_TRY(launch_rotate_assemble(dp, ns, lst + s0, peaks, pld, Kt, ct, et,
                                               pl->abuf.as<double>(), lda, (int64_t)M * lda, st));
                cudaEvent_t *ev = time_it ? next_events() : nullptr;
                MFB_TRY(launch_exact_search(ns, M, bs, pl->abuf.as<double>(), lda, (int64_t)M * lda, y, M,
                                            lst + s0, pl->scratch.p, pl->tuple.as<long long>(), st, ev, nullptr,
                                            mask ? mask + (size_t)s0 * 2 * mask_ld : nullptr, mask_ld));
                if (ev) { pl->stats[3] += 1; pl->stats[4] += (double)ns; }
                MFB_TRY(launch_gather_from_A(ns, M, bs, pl->abuf.as<double>(), lda, (int64_t)M * lda,
                                             pl->tuple.as<long long>(), lst + s0, pl->asmall.as<double>(),
                                             pl->idx5.as<int32_t>(), st));
            }
            pl->stats[1] += (double)n;
            return MFB_OK;
        };
        if (!(flags & 1) && fast_supported(dp, Kt, ct, et)) {
            // fast tier: DMMA screening; uncertain voxels are redone by the exact tier
            const int64_t fsub = std::min<int64_t>(cnt, 8192);
            MFB_TRY(pl->fscratch.ensure(fast_scratch_bytes(dp.M, dp.N, dp.N, fsub, 0, 0)));
            FastProblem fp;
            memset(&fp, 0, sizeof(fp));
            fp.src = 0; fp.N1 = fp.N2 = dp.N; fp.csf = ct;
            const int mask_ld = exact_mask_ld(bs);
            MFB_TRY(pl->redomask.ensure((size_t)cnt * 2 * mask_ld));
            fp.redo_mask = pl->redomask.as<uint8_t>(); fp.mask_ld = mask_ld;
            MFB_TRY(pl->redo.ensure(sizeof(int32_t) * (nv + 8)));
            int32_t *redo_count = pl->redo.as<int32_t>();   // [count, reasons[4], -, -, -, list...]
            int32_t *reasons = redo_count + 1;
            int32_t *redo_list = redo_count + 8;
            MFB_CUDA_TRY(cudaMemsetAsync(redo_count, 0, 8 * sizeof(int32_t), st));
            for (int64_t s0 = 0; s0 < cnt; s0 += fsub) {
                const int64_t ns = std::min(fsub, cnt - s0);
                cudaEvent_t *ev = timed ? next_events() : nullptr;
                MFB_TRY(launch_fast_search(dp, fp, ns, list + s0, peaks, pld, y, pl->fscratch.p,
                                           pl->tuple.as<long long>(), redo_list, redo_count, reasons, st, ev));
                if (ev) { pl->stats[3] += 1; pl->stats[4] += (double)ns; }
            }
            MFB_TRY(launch_gather_from_table(dp, cnt, Kt, ct, et, peaks, pld, pl->tuple.as<long long>(),
                                             list, pl->asmall.as<double>(), pl->idx5.as<int32_t>(), st));
            int32_t head[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            MFB_CUDA_TRY(cudaMemcpyAsync(head, redo_count, sizeof(head), cudaMemcpyDeviceToHost, st));
            MFB_CUDA_TRY(cudaStreamSynchronize(st));
            const int32_t n_redo = head[0];
#ifdef MFB_EXPERIMENTS
            if (getenv("MFB_FAST_DEBUG") && (atoi(getenv("MFB_FAST_DEBUG")) & 8))
                fprintf(stderr, "[mfb] %lld voxels: %d rare-path warp entries, %d competitive pairs, %d level-2 warp entries\n",
                        (long long)cnt, head[5], head[6], head[7]);
#endif
            pl->stats[0] += (double)(cnt - n_redo);
            pl->stats[6] += head[2];          // ill-conditioned competitor
            pl->stats[7] += head[3] + 1e-6 * head[4] ;  // near ties (+ 1e-6 * pair-independent branch)
            if (n_redo > 0) MFB_TRY(run_exact(redo_list, n_redo, false, pl->redomask.as<uint8_t>(), mask_ld));
        } else if (!(flags & 1) && (fast_supported_materialised(dp, Kt, ct, et) || fast3_supported_materialised(dp, Kt, ct, et))) {
            // between-shell protocols, M > 112 and [N, N, E] voxels: materialise the rotated
            // dictionaries of a sub-chunk (k_rotate_assemble), screen them with the
            // explicit-source fast tier (pair scan, or triple scan when the EAR block is the
            // third searched block), redo the uncertain voxels on the same dictionaries in
            // reference order
            const bool triple = et != 0;
            const int64_t lda = (bs.ntot + 1) & ~(int64_t)1;
            const int64_t strideA = (int64_t)M * lda;
            // the blocks the triple scan searches: with the CSF column ([N, N, 1, E]) the fascicle and
            // EAR blocks, projected off that column
            BlockSpec bs3 = bs;
            const int csf_col = (triple && ct) ? bs.start[2] : -1;
            if (csf_col >= 0) {
                bs3.nb = 3;
                bs3.size[2] = bs.size[3]; bs3.start[2] = bs.start[3];
            }
            auto fbytes = [&](int64_t n) {
                return triple ? fast3_scratch_bytes(M, bs3, n, 0) : fast_scratch_bytes(M, dp.N, dp.N, n, 1, 0);
            };
            const size_t per_vox = (size_t)strideA * sizeof(double) + fbytes(1);
            // (triple scan: ~14 MB per voxel at N = 1000 -- the three correlation matrices; with 6 GB a sub-chunk
            // is 440 voxels and the scan's 6 CTAs per voxel make 9 waves: a budget of 24 GB keeps the tail small)
            int64_t sub = std::max<int64_t>(1, std::min<int64_t>(8192, (triple ? 8 : 2) * pl->exact_budget / per_vox));
            sub = std::min(sub, cnt);
            MFB_TRY(pl->abuf.ensure((size_t)strideA * sizeof(double) * sub));
            MFB_TRY(pl->fscratch.ensure(fbytes(sub)));
            MFB_TRY(pl->redo.ensure(sizeof(int32_t) * (2 * sub + 8)));
            int32_t *redo_count = pl->redo.as<int32_t>(), *reasons = redo_count + 1;
            int32_t *redo_list = redo_count + 8, *redo_local = redo_list + sub;
            FastProblem fp;
            memset(&fp, 0, sizeof(fp));
            fp.src = 1; fp.N1 = fp.N2 = dp.N; fp.csf = ct;
            fp.A = pl->abuf.as<double>(); fp.lda = lda; fp.strideA = strideA;
            fp.start1 = 0; fp.start2 = dp.N; fp.start3 = ct ? 2 * dp.N : 0;
            fp.a_by_local = 1; fp.redo_local = redo_local;
            const int mask_ld = exact_mask_ld(bs);
            if (!triple) {
                MFB_TRY(pl->redomask.ensure((size_t)sub * 2 * mask_ld));
                fp.redo_mask = pl->redomask.as<uint8_t>(); fp.mask_ld = mask_ld;
            }
            for (int64_t s0 = 0; s0 < cnt; s0 += sub) {
                const int64_t ns = std::min(sub, cnt - s0);
                MFB_CUDA_TRY(cudaMemsetAsync(redo_count, 0, 8 * sizeof(int32_t), st));
                MFB_TRY(launch_rotate_assemble(dp, ns, list + s0, peaks, pld, Kt, ct, et,
                                               pl->abuf.as<double>(), lda, strideA, st));
                cudaEvent_t *ev = timed ? next_events() : nullptr;
                if (triple)
                    MFB_TRY(launch_fast_search3(M, bs3, pl->abuf.as<double>(), lda, strideA, ns, y, pl->fscratch.p,
                                                pl->tuple.as<long long>(), redo_list, redo_count, reasons, st, ev,
                                                list + s0, 1, redo_local, csf_col));
                else
                    MFB_TRY(launch_fast_search(dp, fp, ns, list + s0, peaks, pld, y, pl->fscratch.p,
                                               pl->tuple.as<long long>(), redo_list, redo_count, reasons, st, ev));
                if (ev) { pl->stats[3] += 1; pl->stats[4] += (double)ns; }
                int32_t head[8] = {0, 0, 0, 0, 0, 0, 0, 0};
                MFB_CUDA_TRY(cudaMemcpyAsync(head, redo_count, sizeof(head), cudaMemcpyDeviceToHost, st));
                MFB_CUDA_TRY(cudaStreamSynchronize(st));
                const int32_t n_redo = head[0];
                pl->stats[0] += (double)(ns - n_redo);
                pl->stats[1] += (double)n_redo;
                pl->stats[6] += head[2];
                pl->stats[7] += head[3] + 1e-6 * head[4];
                if (n_redo > 0) {
                    MFB_TRY(pl->scratch.ensure(exact_scratch_bytes(n_redo, bs)));
                    MFB_TRY(launch_exact_search(n_redo, M, bs, pl->abuf.as<double>(), lda, strideA, y, M,
                                                redo_list, pl->scratch.p, pl->tuple.as<long long>(), st,
                                                nullptr, redo_local, triple ? nullptr : pl->redomask.as<uint8_t>(),
                                                mask_ld));
                }
                MFB_TRY(launch_gather_from_A(ns, M, bs, pl->abuf.as<double>(), lda, strideA,
                                             pl->tuple.as<long long>(), list + s0, pl->asmall.as<double>(),
                                             pl->idx5.as<int32_t>(), st));
            }
        } else if (!(flags & 1) && single_fascicle_supported(dp, Kt, ct, et) && list) {
            // one fascicle: fused rotation + closed forms in the reference's arithmetic
            MFB_TRY(launch_single_fascicle(dp, cnt, list, peaks, pld, y, ct, et, pl->tuple.as<long long>(), st));
            MFB_TRY(launch_gather_from_table(dp, cnt, Kt, ct, et, peaks, pld, pl->tuple.as<long long>(),
                                             list, pl->asmall.as<double>(), pl->idx5.as<int32_t>(), st));
            pl->stats[5] += (double)cnt;
        } else {
            MFB_TRY(run_exact(list, cnt, timed));
        }
    }
    MFB_TRY(launch_evaluate(nv, M, 0, pl->nbv.as<uint8_t>(), pl->asmall.as<double>(), y, M,
                            pl->tuple.as<long long>(), pl->w5.as<double>(), pl->obj.as<double>(),
                            pl->yrec.as<double>(), pl->idx5.as<int32_t>(), st));
    MFB_TRY(launch_finalize(nv, M, maxfasc, csf_on, ear_on, K, csf, ear, y, pl->w5.as<double>(),
                            pl->idx5.as<int32_t>(), pl->obj.as<double>(), pl->yrec.as<double>(),
                            params, st));
    if (pl->events_used) {  // dominant-kernel time of this chunk
        MFB_CUDA_TRY(cudaStreamSynchronize(st));
        for (size_t i = 0; i + 1 < pl->events_used; i += 2) {
            float ms = 0.f;
            MFB_CUDA_TRY(cudaEventElapsedTime(&ms, pl->events[i], pl->events[i + 1]));
            pl->stats[2] += ms;
        }
        pl->events_used = 0;
    }
    return MFB_OK;
}

static const int64_t kFitChunk = 32768;

extern "C" int mfb_fit(mfb_plan *pl, int64_t V, const double *y, const double *peaks,
                       const int32_t *K, const uint8_t *csf, const uint8_t *ear, int maxfasc,
                       int csf_on, int ear_on, double *params_out, int flags, void *stream)
{
    if (!pl || V < 0 || maxfasc < 0 || maxfasc > 2 || (V > 0 && (!y || !K || !params_out)) ||
        (maxfasc > 0 && V > 0 && !peaks)) {
        set_error("mfb_fit: invalid argument");
        return MFB_EINVAL;
    }
    MFB_ON_DEVICE(pl->device);
    cudaStream_t st = (cudaStream_t)stream;
    const int P = 1 + 2 * maxfasc + (csf_on ? 1 : 0) + 2 * (ear_on ? 1 : 0) + 2;
    for (int i = 0; i < 8; i++) pl->stats[i] = 0;
    for (int64_t v0 = 0; v0 < V; v0 += kFitChunk) {
        const int64_t nv = std::min(kFitChunk, V - v0);
        MFB_TRY(fit_chunk(pl, nv, y + v0 * pl->dp.M, peaks ? peaks + v0 * 3 * maxfasc : nullptr, K + v0,
                          csf ? csf + v0 : nullptr, ear ? ear + v0 : nullptr, maxfasc, csf_on ? 1 : 0,
                          ear_on ? 1 : 0, params_out + v0 * P, flags, st));
    }
    return MFB_OK;
}

// ---------------------------------------------------------------------------------
// Host pipeline of mfb_fit_host / mfb_fit_volume.  Chunks of kFitChunk voxels flow through
//   helper thread : gather the chunk's signals (and peaks / K / csf / ear) into a page-locked slot
//   copy stream   : one H2D of the slot into a device slot
//   compute stream: fit_chunk
//   d2h stream    : params rows into a page-locked slot, drained to the caller's array by the
//                   calling thread one chunk later
// so that the gather of chunk c + 2, the upload of chunk c + 1, the search of chunk c and the
// download of chunk c - 1 overlap.  Slots and streams belong to the plan and are reused by
// later calls.
// ---------------------------------------------------------------------------------
namespace {

struct SlotLayout {
    size_t y, peaks, K, csf, ear, total;
    SlotLayout(int64_t C, int M, int maxfasc)
    {
        auto up = [](size_t x) { return (x + 255) & ~(size_t)255; };
        y = 0;
        peaks = up(sizeof(double) * C * M);
        K = peaks + up(sizeof(double) * C * std::max(1, 3 * maxfasc));
        csf = K + up(sizeof(int32_t) * C);
        ear = csf + up((size_t)C);
        total = ear + up((size_t)C);
    }
};

struct GatherJob {
    int64_t V, C;
    int M, maxfasc, dtype;
    const void *data;
    const int64_t *offsets;   // element offset of each voxel's first measurement, or null (v * row_stride)
    int64_t row_stride, meas_stride;
    const double *peaks;
    const int32_t *K;
    const uint8_t *csf, *ear;
};

template <typename T>
static void gather_rows(const GatherJob &j, int64_t v0, int64_t nv, double *dst)
{
    const T *base = (const T *)j.data;
    const int M = j.M;
    for (int64_t v = 0; v < nv; v++) {
        const T *src = base + (j.offsets ? j.offsets[v0 + v] : (v0 + v) * j.row_stride);
        double *d = dst + v * M;
        if (j.meas_stride == 1) {
            if (sizeof(T) == sizeof(double)) memcpy(d, src, sizeof(double) * M);
            else for (int m = 0; m < M; m++) d[m] = (double)src[m];
        } else {
            for (int m = 0; m < M; m++) d[m] = (double)src[(int64_t)m * j.meas_stride];
        }
    }
}

static void gather_chunk(const GatherJob &j, int64_t v0, int64_t nv, char *slot, const SlotLayout &L)
{
    double *y = (double *)(slot + L.y);
    if (j.dtype == MFB_F32) gather_rows<float>(j, v0, nv, y);
    else gather_rows<double>(j, v0, nv, y);
    if (j.maxfasc > 0) memcpy(slot + L.peaks, j.peaks + v0 * 3 * j.maxfasc, sizeof(double) * nv * 3 * j.maxfasc);
    memcpy(slot + L.K, j.K + v0, sizeof(int32_t) * nv);
    if (j.csf) memcpy(slot + L.csf, j.csf + v0, nv);
    if (j.ear) memcpy(slot + L.ear, j.ear + v0, nv);
}

struct PipeSync {
    std::mutex mu;
    std::condition_variable cv;
    int64_t gathered = 0;   // chunks staged by the helper thread
    int64_t uploaded = 0;   // chunks whose H2D copy has completed (their host slot is free again)
    bool abort = false;
};

static void CUDART_CB slot_uploaded(void *arg)
{
    PipeSync *ps = (PipeSync *)arg;
    { std::lock_guard<std::mutex> lk(ps->mu); ps->uploaded++; }
    ps->cv.notify_all();
}

}  // namespace

static const int kHostSlots = 3;

static int fit_pipeline(mfb_plan *pl, const GatherJob &job, int csf_on, int ear_on, double *params_out, int flags)
{
    const int M = pl->dp.M, maxfasc = job.maxfasc;
    const int P = 1 + 2 * maxfasc + (csf_on ? 1 : 0) + 2 * (ear_on ? 1 : 0) + 2;
    const int64_t V = job.V, C = job.C;
    const int64_t nchunks = (V + C - 1) / C;
    for (int i = 0; i < 8; i++) pl->stats[i] = 0;
    if (V == 0) return MFB_OK;
    const SlotLayout L(C, M, maxfasc);
    for (int i = 0; i < kHostSlots; i++) MFB_TRY(pl->h_in[i].ensure(L.total));
    for (int i = 0; i < 2; i++) {
        MFB_TRY(pl->d_in[i].ensure(L.total));
        MFB_TRY(pl->d_out[i].ensure(sizeof(double) * C * P));
        MFB_TRY(pl->h_out[i].ensure(sizeof(double) * C * P));
    }
    if (!pl->s_copy) {
        MFB_CUDA_TRY(cudaStreamCreateWithFlags(&pl->s_copy, cudaStreamNonBlocking));
        MFB_CUDA_TRY(cudaStreamCreateWithFlags(&pl->s_d2h, cudaStreamNonBlocking));
        for (int i = 0; i < 6; i++) MFB_CUDA_TRY(cudaEventCreateWithFlags(&pl->pipe_ev[i], cudaEventDisableTiming));
    }
    cudaEvent_t *ev_h2d = pl->pipe_ev, *ev_done = pl->pipe_ev + 2, *ev_d2h = pl->pipe_ev + 4;
    cudaStream_t st = pl->stream;

    PipeSync ps;
    std::thread helper([&]() {
        for (int64_t c = 0; c < nchunks; c++) {
            {
                std::unique_lock<std::mutex> lk(ps.mu);
                ps.cv.wait(lk, [&] { return ps.abort || c < ps.uploaded + kHostSlots; });
                if (ps.abort) return;
            }
            const int64_t v0 = c * C, nv = std::min(C, V - v0);
            gather_chunk(job, v0, nv, (char *)pl->h_in[c % kHostSlots].p, L);
            { std::lock_guard<std::mutex> lk(ps.mu); ps.gathered = c + 1; }
            ps.cv.notify_all();
        }
    });

    auto upload = [&](int64_t c) -> int {
        {
            // the helper waits for free slots, which only upload completions release: a failed
            // device would otherwise leave both threads waiting on each other
            std::unique_lock<std::mutex> lk(ps.mu);
            while (!ps.cv.wait_for(lk, std::chrono::milliseconds(200), [&] { return ps.gathered > c; })) {
                lk.unlock();
                cudaError_t q = cudaStreamQuery(pl->s_copy);
                if (q != cudaSuccess && q != cudaErrorNotReady) {
                    set_error(std::string("mfb_fit_host: upload stream: ") + cudaGetErrorString(q));
                    return MFB_ECUDA;
                }
                lk.lock();
            }
        }
        const int d = (int)(c & 1);
        // the device slot was last read by the search of chunk c - 2
        if (c >= 2) MFB_CUDA_TRY(cudaStreamWaitEvent(pl->s_copy, ev_done[d], 0));
        const int64_t nv = std::min(C, V - c * C);
        // one copy up to the end of the last array in use
        const size_t bytes = job.ear ? L.ear + nv : (job.csf ? L.csf + nv : L.K + sizeof(int32_t) * nv);
        MFB_CUDA_TRY(cudaMemcpyAsync(pl->d_in[d].p, pl->h_in[c % kHostSlots].p, bytes, cudaMemcpyHostToDevice, pl->s_copy));
        MFB_CUDA_TRY(cudaEventRecord(ev_h2d[d], pl->s_copy));
        MFB_CUDA_TRY(cudaLaunchHostFunc(pl->s_copy, slot_uploaded, &ps));
        return MFB_OK;
    };
    auto drain = [&](int64_t c) -> int {
        const int d = (int)(c & 1);
        MFB_CUDA_TRY(cudaEventSynchronize(ev_d2h[d]));
        const int64_t nv = std::min(C, V - c * C);
        memcpy(params_out + c * C * P, pl->h_out[d].p, sizeof(double) * nv * P);
        return MFB_OK;
    };
    auto body = [&]() -> int {
        MFB_TRY(upload(0));
        for (int64_t c = 0; c < nchunks; c++) {
            if (c + 1 < nchunks) MFB_TRY(upload(c + 1));
            const int d = (int)(c & 1);
            const int64_t nv = std::min(C, V - c * C);
            MFB_CUDA_TRY(cudaStreamWaitEvent(st, ev_h2d[d], 0));
            if (c >= 2) MFB_CUDA_TRY(cudaStreamWaitEvent(st, ev_d2h[d], 0));   // its params slot has been downloaded
            char *din = (char *)pl->d_in[d].p;
            MFB_TRY(fit_chunk(pl, nv, (const double *)(din + L.y), (const double *)(din + L.peaks),
                              (const int32_t *)(din + L.K), job.csf ? (const uint8_t *)(din + L.csf) : nullptr,
                              job.ear ? (const uint8_t *)(din + L.ear) : nullptr, maxfasc, csf_on ? 1 : 0,
                              ear_on ? 1 : 0, pl->d_out[d].as<double>(), flags, st));
            MFB_CUDA_TRY(cudaEventRecord(ev_done[d], st));
            MFB_CUDA_TRY(cudaStreamWaitEvent(pl->s_d2h, ev_done[d], 0));
            MFB_CUDA_TRY(cudaMemcpyAsync(pl->h_out[d].p, pl->d_out[d].p, sizeof(double) * nv * P,
                                         cudaMemcpyDeviceToHost, pl->s_d2h));
            MFB_CUDA_TRY(cudaEventRecord(ev_d2h[d], pl->s_d2h));
            if (c >= 1) MFB_TRY(drain(c - 1));
        }
        MFB_TRY(drain(nchunks - 1));
        return MFB_OK;
    };
    int rc = body();
    if (rc != MFB_OK) {
        { std::lock_guard<std::mutex> lk(ps.mu); ps.abort = true; }
        ps.cv.notify_all();
    }
    helper.join();
    // nothing of this call may still be in flight when the slots are reused or PipeSync dies
    cudaError_t e1 = cudaStreamSynchronize(pl->s_copy), e2 = cudaStreamSynchronize(st),
                e3 = cudaStreamSynchronize(pl->s_d2h);
    if (rc == MFB_OK && (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess)) {
        set_error(std::string("mfb_fit_host: ") + cudaGetErrorString(e1 != cudaSuccess ? e1 : (e2 != cudaSuccess ? e2 : e3)));
        rc = MFB_ECUDA;
    }
    return rc;
}

extern "C" int mfb_fit_volume(mfb_plan *pl, int64_t V, const void *data, int dtype,
                              const int64_t *voxel_offset, int64_t row_stride, int64_t meas_stride,
                              const double *peaks, const int32_t *K, const uint8_t *csf,
                              const uint8_t *ear, int maxfasc, int csf_on, int ear_on,
                              double *params_out, int flags)
{
    if (!pl || V < 0 || maxfasc < 0 || maxfasc > 2 || (V > 0 && (!data || !K || !params_out)) ||
        (maxfasc > 0 && V > 0 && !peaks) || (dtype != MFB_F64 && dtype != MFB_F32)) {
        set_error("mfb_fit_volume: invalid argument");
        return MFB_EINVAL;
    }
    MFB_ON_DEVICE(pl->device);
    GatherJob job;
    job.V = V; job.C = std::min<int64_t>(kFitChunk, std::max<int64_t>(V, 1));
    job.M = pl->dp.M; job.maxfasc = maxfasc; job.dtype = dtype;
    job.data = data; job.offsets = voxel_offset; job.row_stride = row_stride; job.meas_stride = meas_stride;
    job.peaks = peaks; job.K = K; job.csf = csf; job.ear = ear;
    return fit_pipeline(pl, job, csf_on, ear_on, params_out, flags);
}

extern "C" int mfb_fit_host(mfb_plan *pl, int64_t V, const double *y, const double *peaks,
                            const int32_t *K, const uint8_t *csf, const uint8_t *ear, int maxfasc,
                            int csf_on, int ear_on, double *params_out, int flags)
{
    if (!pl) { set_error("mfb_fit_host: invalid argument"); return MFB_EINVAL; }
    return mfb_fit_volume(pl, V, y, MFB_F64, nullptr, pl->dp.M, 1, peaks, K, csf, ear, maxfasc, csf_on,
                          ear_on, params_out, flags);
}

extern "C" int mfb_fit_stats(mfb_plan *pl, double *out, int n)
{
    if (!pl || !out) { set_error("mfb_fit_stats: invalid argument"); return MFB_EINVAL; }
    for (int i = 0; i < n && i < 8; i++) out[i] = pl->stats[i];
    return MFB_OK;
}
