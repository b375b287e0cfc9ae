// Sufficient statistics reduced on device (the only quantities that cross GPUs):
//   transition counts N_ij           (jax_moseq.utils.transitions.count_transitions)
//   per-state AR Gram matrices       (the einsums of jax_moseq.models.arhmm.gibbs._resample_regression_params)
// Frames are stably counting-sorted by state so every state's Gram is one dense, deterministic
// accumulation in double over contiguous work.
//
// Gram feature order: f = [x_{t-L} .. x_{t-1} (n) | x_t (d) | 1],  F = n + d + 1,  out (K, F, F) double.
#include "common.cuh"
#include "../../include/kpms_b200.h"

namespace kpms {

constexpr int SORT_TILE = 1024;
constexpr int GRAM_SPLIT = 8;
constexpr int GRAM_FT = 64;

// valid frame t' of chain nn: mask[nn][L + t'] != 0
__global__ void __launch_bounds__(256)
transition_count_kernel(const int* __restrict__ z, const int* __restrict__ mask, int N, int T, int L, int K,
                        int* __restrict__ counts) {
    extern __shared__ int hist[];
    const int Tp = T - L;
    for (int i = threadIdx.x; i < K * K; i += blockDim.x) hist[i] = 0;
    __syncthreads();
    const long long total = (long long)N * (Tp - 1);
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
         e += (long long)gridDim.x * blockDim.x) {
        const int nn = (int)(e / (Tp - 1)), t = (int)(e % (Tp - 1));
        const int* mk = mask + (size_t)nn * T + L + t;
        if (mk[0] != 0 && mk[1] != 0) {
            const int a = z[(size_t)nn * Tp + t], b = z[(size_t)nn * Tp + t + 1];
            atomicAdd(&hist[a * K + b], 1);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < K * K; i += blockDim.x)
        if (hist[i]) atomicAdd(&counts[i], hist[i]);
}

__global__ void __launch_bounds__(256)
tile_hist_kernel(const int* __restrict__ z, const int* __restrict__ mask, int N, int T, int L, int K,
                 int* __restrict__ tile_hist) {
    extern __shared__ int hist[];
    const int Tp = T - L;
    const long long total = (long long)N * Tp;
    for (int i = threadIdx.x; i < K; i += blockDim.x) hist[i] = 0;
    __syncthreads();
    const long long base = (long long)blockIdx.x * SORT_TILE;
    for (int e = threadIdx.x; e < SORT_TILE; e += blockDim.x) {
        const long long f = base + e;
        if (f < total) {
            const int nn = (int)(f / Tp), t = (int)(f % Tp);
            if (mask[(size_t)nn * T + L + t] != 0) atomicAdd(&hist[z[f]], 1);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < K; i += blockDim.x) tile_hist[(size_t)blockIdx.x * K + i] = hist[i];
}

// one block: per-state exclusive scan over tiles, then exclusive scan over states
__global__ void tile_scan_kernel(int* __restrict__ tile_hist, int n_tiles, int K, int* __restrict__ state_start) {
    extern __shared__ int tot[];
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
        int run = 0;
        for (int t = 0; t < n_tiles; ++t) {
            int c = tile_hist[(size_t)t * K + k];
            tile_hist[(size_t)t * K + k] = run;
            run += c;
        }
        tot[k] = run;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int run = 0;
        for (int k = 0; k < K; ++k) { state_start[k] = run; run += tot[k]; }
        state_start[K] = run;
    }
}

__global__ void __launch_bounds__(128)
tile_scatter_kernel(const int* __restrict__ z, const int* __restrict__ mask, const int* __restrict__ tile_off,
                    const int* __restrict__ state_start, int N, int T, int L, int K, int* __restrict__ order) {
    __shared__ int zs[SORT_TILE];
    const int Tp = T - L;
    const long long total = (long long)N * Tp;
    const long long base = (long long)blockIdx.x * SORT_TILE;
    for (int e = threadIdx.x; e < SORT_TILE; e += blockDim.x) {
        const long long f = base + e;
        int val = -1;
        if (f < total) {
            const int nn = (int)(f / Tp), t = (int)(f % Tp);
            if (mask[(size_t)nn * T + L + t] != 0) val = z[f];
        }
        zs[e] = val;
    }
    __syncthreads();
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
        int pos = state_start[k] + tile_off[(size_t)blockIdx.x * K + k];
        for (int e = 0; e < SORT_TILE; ++e)
            if (zs[e] == k) order[pos++] = (int)(base + e);
    }
}

template <typename R>
__global__ void __launch_bounds__(256)
gram_partial_kernel(const R* __restrict__ x, const int* __restrict__ order, const int* __restrict__ state_start,
                    int T, int d, int L, double* __restrict__ partial) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int n = d * L, F = n + d + 1, NP = F * (F + 1) / 2, Tp = T - L;
    double* vt = reinterpret_cast<double*>(smem_raw);                 // GRAM_FT x F
    const int k = blockIdx.y, sp = blockIdx.x;
    const int s0 = state_start[k], cnt = state_start[k + 1] - s0;
    const int per = (cnt + GRAM_SPLIT - 1) / GRAM_SPLIT;
    const int lo = s0 + min(sp * per, cnt), hi = s0 + min((sp + 1) * per, cnt);
    constexpr int PPT = 8;                                            // pairs per thread (<= 8*256 = 2048 pairs)
    double acc[PPT];
    int pa[PPT], pb[PPT];
#pragma unroll
    for (int q = 0; q < PPT; ++q) {
        acc[q] = 0.0;
        int p = threadIdx.x + q * 256;
        int a = 0, b = 0;
        if (p < NP) {
            a = (int)((sqrtf(8.0f * (float)p + 1.0f) - 1.0f) * 0.5f);
            while ((a + 1) * (a + 2) / 2 <= p) ++a;
            while (a * (a + 1) / 2 > p) --a;
            b = p - a * (a + 1) / 2;
        }
        pa[q] = a;
        pb[q] = b;
    }
    for (int f0 = lo; f0 < hi; f0 += GRAM_FT) {
        const int nf = min(GRAM_FT, hi - f0);
        __syncthreads();
        for (int idx = threadIdx.x; idx < nf * F; idx += blockDim.x) {
            const int fr = idx / F, e = idx % F;
            const int f = order[f0 + fr];
            const int nn = f / Tp, t = f % Tp;
            vt[idx] = (e < n + d) ? (double)x[((size_t)nn * T + t) * d + e] : 1.0;
        }
        __syncthreads();
#pragma unroll
        for (int q = 0; q < PPT; ++q) {
            if (threadIdx.x + q * 256 < NP) {
                double a = acc[q];
                for (int fr = 0; fr < nf; ++fr) a = fma(vt[fr * F + pa[q]], vt[fr * F + pb[q]], a);
                acc[q] = a;
            }
        }
    }
    double* out = partial + ((size_t)k * GRAM_SPLIT + sp) * F * F;
#pragma unroll
    for (int q = 0; q < PPT; ++q) {
        if (threadIdx.x + q * 256 < NP) {
            out[pa[q] * F + pb[q]] = acc[q];
            out[pb[q] * F + pa[q]] = acc[q];
        }
    }
}

__global__ void gram_reduce_kernel(const double* __restrict__ partial, int FF, double* __restrict__ out) {
    const int k = blockIdx.x;
    for (int e = threadIdx.x; e < FF; e += blockDim.x) {
        double a = 0.0;
        for (int sp = 0; sp < GRAM_SPLIT; ++sp) a += partial[((size_t)k * GRAM_SPLIT + sp) * FF + e];
        out[(size_t)k * FF + e] = a;
    }
}

static void stats_ws_layout(int N, int T, int d, int L, int K, size_t off[5]) {
    const size_t total = (size_t)N * (T - L);
    const size_t n_tiles = (total + SORT_TILE - 1) / SORT_TILE;
    const size_t F = (size_t)d * L + d + 1;
    off[0] = 0;
    off[1] = off[0] + align_up(n_tiles * K * sizeof(int), 256);                 // tile_hist / offsets
    off[2] = off[1] + align_up((size_t)(K + 1) * sizeof(int), 256);             // state_start
    off[3] = off[2] + align_up(total * sizeof(int), 256);                       // order
    off[4] = off[3] + align_up((size_t)K * GRAM_SPLIT * F * F * sizeof(double), 256);  // partial
}

template <typename R>
static int ar_suffstats_impl(const void* x, const int* z, const int* mask, int N, int T, int d, int L, int K,
                             double* gram, void* ws, cudaStream_t st) {
    if (T - L < 1) return set_error(-3, "ar_suffstats: T (%d) must exceed nlags (%d)", T, L);
    const int F = d * L + d + 1;
    if (F * (F + 1) / 2 > 8 * 256) return set_error(-3, "ar_suffstats: feature dimension %d too large", F);
    size_t off[5];
    stats_ws_layout(N, T, d, L, K, off);
    char* base = reinterpret_cast<char*>(ws);
    int* tile_hist = reinterpret_cast<int*>(base + off[0]);
    int* state_start = reinterpret_cast<int*>(base + off[1]);
    int* order = reinterpret_cast<int*>(base + off[2]);
    double* partial = reinterpret_cast<double*>(base + off[3]);
    const long long total = (long long)N * (T - L);
    const int n_tiles = (int)((total + SORT_TILE - 1) / SORT_TILE);
    { KPMS_LAUNCH("sort_tile_hist", st); tile_hist_kernel<<<n_tiles, 256, K * sizeof(int), st>>>(z, mask, N, T, L, K, tile_hist); }
    { KPMS_LAUNCH("sort_tile_scan", st); tile_scan_kernel<<<1, 128, K * sizeof(int), st>>>(tile_hist, n_tiles, K, state_start); }
    { KPMS_LAUNCH("sort_tile_scatter", st); tile_scatter_kernel<<<n_tiles, 128, 0, st>>>(z, mask, tile_hist, state_start, N, T, L, K, order); }
    dim3 grid(GRAM_SPLIT, K);
    size_t smem = (size_t)GRAM_FT * F * sizeof(double);
    { KPMS_LAUNCH("gram_partial", st); gram_partial_kernel<R><<<grid, 256, smem, st>>>((const R*)x, order, state_start, T, d, L, partial); }
    { KPMS_LAUNCH("gram_reduce", st); gram_reduce_kernel<<<K, 256, 0, st>>>(partial, F * F, gram); }
    return check_launch("ar_suffstats");
}

}  // namespace kpms

using namespace kpms;

extern "C" {

int kpms_transition_counts(const int32_t* z, const int32_t* mask, int N, int T, int L, int K, int32_t* counts,
                           void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if ((size_t)K * K * sizeof(int) > 200 * 1024) return set_error(-3, "transition_counts: num_states %d too large", K);
    cudaMemsetAsync(counts, 0, (size_t)K * K * sizeof(int), st);
    if (T - L < 2) return 0;
    const long long total = (long long)N * (T - L - 1);
    int blocks = (int)max(1LL, min((long long)148 * 4, (total + 255) / 256));
    size_t smem = (size_t)K * K * sizeof(int);
    cudaFuncSetAttribute(transition_count_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    { KPMS_LAUNCH("transition_counts", st); transition_count_kernel<<<blocks, 256, smem, st>>>(z, mask, N, T, L, K, counts); }
    return check_launch("transition_counts");
}

size_t kpms_ar_suffstats_workspace_bytes(int N, int T, int d, int L, int K) {
    size_t off[5];
    stats_ws_layout(N, T, d, L, K, off);
    return off[4];
}

int kpms_ar_suffstats(int dtype, const void* x, const int32_t* z, const int32_t* mask, int N, int T, int d, int L,
                      int K, double* gram, void* ws, void* stream) {
    return KPMS_DISPATCH_DTYPE(dtype, ar_suffstats_impl, x, z, mask, N, T, d, L, K, gram, ws, (cudaStream_t)stream);
}

}  // extern "C"
