// mc.cu -- Monte-Carlo dictionary generation of libmfb200, sm_100a.
//
// Reference: mf_utils.py:2758-2812 `monte_carlo_average` (the other Numba kernel of the
// repository): the DW-MRI signal of sequence i is the spin average of cos(Dscaling * phi),
// phi = sum_d gscaling[i, d] * sim_phases[delta_mapping[i] * num_spins + l, d].
// One CTA per (sequence, spin slice): the phases of a reference sequence are streamed from
// HBM / L2 with coalesced loads, the cosine runs on the FP64 pipe, partial sums are reduced
// with warp shuffles and finished in a fixed order by a second kernel (deterministic).
#include "common.cuh"

namespace mfb {

#define MC_THREADS 256

template <int DIM>
__global__ void __launch_bounds__(MC_THREADS)
k_mc_partial(int64_t n_entries, int dim_rt, const double *__restrict__ phases, int64_t n_seq,
             const long long *__restrict__ delta_mapping, const double *__restrict__ gscaling,
             double Dscaling, int64_t num_spins, int nsplit, double *__restrict__ partial)
{
    const int64_t iseq = blockIdx.x;
    const int split = blockIdx.y;
    const int dim = DIM > 0 ? DIM : dim_rt;
    const long long iref = delta_mapping[iseq];
    __shared__ double red[MC_THREADS / 32];
    const bool ok = iref >= 0 && (iref + 1) * num_spins <= n_entries;
    double gs[8];
    for (int d = 0; d < dim; d++) gs[d] = gscaling[iseq * dim + d];
    const int64_t per = (num_spins + nsplit - 1) / nsplit;
    const int64_t l0 = split * per, l1 = min(num_spins, l0 + per);
    const double *base = phases + (ok ? iref * num_spins * dim : 0);
    double acc = 0.0;
    if (ok)
        for (int64_t l = l0 + threadIdx.x; l < l1; l += MC_THREADS) {
            // same expression order as the reference: ph += g[d] * phase[d], separately rounded
            double ph = 0.0;
#pragma unroll
            for (int d = 0; d < (DIM > 0 ? DIM : 8); d++)
                if (d < dim) ph = __dadd_rn(ph, __dmul_rn(gs[d], __ldg(base + l * dim + d)));
            acc += cos(__dmul_rn(Dscaling, ph));
        }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < MC_THREADS / 32; w++) s += red[w];
        partial[iseq * nsplit + split] = ok ? s : __longlong_as_double(0x7ff8000000000000LL);
    }
}

__global__ void __launch_bounds__(128)
k_mc_finish(int64_t n_seq, int nsplit, int64_t num_spins, const double *__restrict__ partial,
            double *__restrict__ signal)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_seq) return;
    double s = 0.0;
    for (int k = 0; k < nsplit; k++) s += partial[i * nsplit + k];
    signal[i] = s / (double)num_spins;
}

int launch_mc_average(int64_t n_entries, int dim, const double *phases, int64_t n_seq,
                      const long long *delta_mapping, const double *gscaling, double Dscaling,
                      int64_t num_spins, int nsplit, double *partial, double *signal, cudaStream_t st)
{
    if (n_seq == 0) return MFB_OK;
    dim3 grid((unsigned)n_seq, (unsigned)nsplit);
    if (dim == 2)
        MFB_LAUNCH(k_mc_partial<2>, grid, MC_THREADS, 0, st, n_entries, dim, phases, n_seq, delta_mapping,
                   gscaling, Dscaling, num_spins, nsplit, partial);
    else if (dim == 3)
        MFB_LAUNCH(k_mc_partial<3>, grid, MC_THREADS, 0, st, n_entries, dim, phases, n_seq, delta_mapping,
                   gscaling, Dscaling, num_spins, nsplit, partial);
    else
        MFB_LAUNCH(k_mc_partial<0>, grid, MC_THREADS, 0, st, n_entries, dim, phases, n_seq, delta_mapping,
                   gscaling, Dscaling, num_spins, nsplit, partial);
    MFB_LAUNCH(k_mc_finish, (unsigned)((n_seq + 127) / 128), 128, 0, st, n_seq, nsplit, num_spins, partial,
               signal);
    return MFB_OK;
}

int mc_nsplit(int64_t n_seq, int64_t num_spins)
{
    // enough CTAs for ~4 waves of 148 SMs, at least 4096 spins per CTA
    int64_t want = (4 * 148 + n_seq - 1) / (n_seq > 0 ? n_seq : 1);
    int64_t cap = (num_spins + 4095) / 4096;
    int64_t ns = want < cap ? want : cap;
    return (int)(ns < 1 ? 1 : (ns > 1024 ? 1024 : ns));
}

}  // namespace mfb
