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

// num_states > 226: the K x K histogram no longer fits shared memory; transitions are sparse (sticky chains), so
// global atomics on the L2-resident table are enough
__global__ void __launch_bounds__(256)
transition_count_global_kernel(const int* __restrict__ z, const int* __restrict__ mask, int N, int T, int L, int K,
                               int* __restrict__ counts) {
    const int Tp = T - L;
    const long long total = (long long)N * (Tp - 1);
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
         e += (long long)gridDim.x * blockDim.x) {
        const int nn = (int)(e / (Tp - 1)), t = (int)(e % (Tp - 1));
        const int* mk = mask + (size_t)nn * T + L + t;
        if (mk[0] != 0 && mk[1] != 0) atomicAdd(&counts[z[(size_t)nn * Tp + t] * K + z[(size_t)nn * Tp + t + 1]], 1);
    }
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
        for (int t0 = 0; t0 < n_tiles; t0 += 16) {           // 16 independent loads in flight per round
            int c[16];
#pragma unroll
            for (int q = 0; q < 16; ++q) c[q] = (t0 + q < n_tiles) ? tile_hist[(size_t)(t0 + q) * K + k] : 0;
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                if (t0 + q < n_tiles) tile_hist[(size_t)(t0 + q) * K + k] = run;
                run += c[q];
            }
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
    __shared__ int rowoff[SORT_TILE];                  // nn * T + t: row of x where the frame's features start
    const int Tp = T - L;
    const long long total = (long long)N * Tp;
    const long long base = (long long)blockIdx.x * SORT_TILE;
    for (int e = threadIdx.x; e < SORT_TILE; e += blockDim.x) {
        const long long f = base + e;
        int val = -1, ro = 0;
        if (f < total) {
            const int nn = (int)(f / Tp), t = (int)(f % Tp);
            ro = nn * T + t;
            if (mask[(size_t)nn * T + L + t] != 0) val = z[f];
        }
        zs[e] = val;
        rowoff[e] = ro;
    }
    __syncthreads();
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
        int pos = state_start[k] + tile_off[(size_t)blockIdx.x * K + k];
        for (int e = 0; e < SORT_TILE; ++e)
            if (zs[e] == k) order[pos++] = rowoff[e];
    }
}

// Per-state Gram matrices on the FP64 tensor pipe (mma.sync m8n8k4 f64).  G_k = sum_f f f' over the
// frames of state k is a GEMM F' F with the frames as the contraction dimension, 4 frames per k-step.
// The A fragment of feature tile ta and the B fragment of feature tile tb are the same numbers
// (lane l: feature 8*tile + l/4 of frame l%4), read straight from x through the sorted row offsets,
// so a k-step costs NT loads and NT(NT+1)/2 DMMAs (lower tile triangle).  Grid (GRAM_SPLIT, K), four
// warps per CTA on disjoint frame ranges, combined in warp order (deterministic).
__device__ __forceinline__ void dmma884_stats(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <typename R, int D_, int L_>
__global__ void __launch_bounds__(128)
gram_partial_kernel(const R* __restrict__ x, const int* __restrict__ order, const int* __restrict__ state_start,
                    double* __restrict__ partial) {
    constexpr int n = D_ * L_, NF = n + D_, F = NF + 1, NT = (F + 7) / 8, NPAIR = NT * (NT + 1) / 2;
    __shared__ double red[NPAIR * 64];
    const int k = blockIdx.y, sp = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int fr = lane & 3, fe = lane >> 2;
    const int s0 = state_start[k], cnt = state_start[k + 1] - s0;
    const int per = (cnt + GRAM_SPLIT - 1) / GRAM_SPLIT;
    const int lo = s0 + min(sp * per, cnt), hi = s0 + min((sp + 1) * per, cnt);
    const int wper = ((hi - lo + 3) / 4 + 3) / 4 * 4;               // frames per warp, whole k-steps
    const int wlo = min(lo + warp * wper, hi), whi = min(wlo + wper, hi);
    double acc[NPAIR][2];
#pragma unroll
    for (int p = 0; p < NPAIR; ++p) acc[p][0] = acc[p][1] = 0.0;
    auto load_row = [&](int g0) -> int { return (g0 + fr < whi) ? order[g0 + fr] : -1; };
    auto load_feat = [&](int ro, double (&fv)[NT]) {
        const R* xr = x + (size_t)max(ro, 0) * D_;
#pragma unroll
        for (int ti = 0; ti < NT; ++ti) {
            const int e = 8 * ti + fe;
            double v = 0.0;
            if (ro >= 0) v = e < NF ? (double)xr[e] : (e == NF ? 1.0 : 0.0);
            fv[ti] = v;
        }
    };
    double cur[NT], nxt[NT];
    int ro_next = -1;
    if (wlo < whi) {
        load_feat(load_row(wlo), cur);
        ro_next = load_row(wlo + 4);
    }
    for (int g0 = wlo; g0 < whi; g0 += 4) {
        load_feat(ro_next, nxt);                                    // features of the next k-step
        ro_next = load_row(g0 + 8);                                 // row offsets two k-steps ahead
#pragma unroll
        for (int ta = 0; ta < NT; ++ta)
#pragma unroll
            for (int tb = 0; tb <= ta; ++tb) dmma884_stats(acc[ta * (ta + 1) / 2 + tb][0], acc[ta * (ta + 1) / 2 + tb][1], cur[ta], cur[tb]);
#pragma unroll
        for (int ti = 0; ti < NT; ++ti) cur[ti] = nxt[ti];
    }
    // combine the four warps in order
    for (int w = 0; w < 4; ++w) {
        if (warp == w) {
#pragma unroll
            for (int p = 0; p < NPAIR; ++p) {
                double2* slot = reinterpret_cast<double2*>(red) + p * 32 + lane;
                if (w == 0) *slot = make_double2(acc[p][0], acc[p][1]);
                else { double2 v = *slot; v.x += acc[p][0]; v.y += acc[p][1]; *slot = v; }
            }
        }
        __syncthreads();
    }
    double* out = partial + ((size_t)k * GRAM_SPLIT + sp) * F * F;
    for (int idx = threadIdx.x; idx < NPAIR * 64; idx += blockDim.x) {
        const int p = idx / 64, l = (idx % 64) / 2, c = idx % 2;
        int ta = 0;
        while ((ta + 1) * (ta + 2) / 2 <= p) ++ta;
        const int tb = p - ta * (ta + 1) / 2;
        const int a = 8 * ta + l / 4, b = 8 * tb + 2 * (l % 4) + c;
        if (a < F && b < F && (ta != tb || a >= b)) {
            const double v = red[idx];
            out[a * F + b] = v;
            out[b * F + a] = v;
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
    bool launched = false;
#define X(DD, LL)                                                                                              \
    if (d == DD && L == LL) {                                                                                  \
        KPMS_LAUNCH("gram_partial", st);                                                                       \
        gram_partial_kernel<R, DD, LL><<<grid, 128, 0, st>>>((const R*)x, order, state_start, partial);        \
        launched = true;                                                                                       \
    }
    KPMS_FOR_EACH_DL(X)
#undef X
    if (!launched) return set_error(-3, "ar_suffstats: unsupported (latent_dim, nlags) = (%d, %d)", d, L);
    { KPMS_LAUNCH("gram_reduce", st); gram_reduce_kernel<<<K, 256, 0, st>>>(partial, F * F, gram); }
    return check_launch("ar_suffstats");
}

}  // namespace kpms

using namespace kpms;

extern "C" {

int kpms_transition_counts(const int32_t* z, const int32_t* mask, int N, int T, int L, int K, int32_t* counts,
                           void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (K < 1 || K > 512) return set_error(-3, "transition_counts: num_states %d outside 1..512", K);
    cudaMemsetAsync(counts, 0, (size_t)K * K * sizeof(int), st);
    if (T - L < 2) return 0;
    const long long total = (long long)N * (T - L - 1);
    int blocks = (int)max(1LL, min((long long)148 * 4, (total + 255) / 256));
    size_t smem = (size_t)K * K * sizeof(int);
    if (smem > 200 * 1024) {
        KPMS_LAUNCH("transition_counts", st);
        transition_count_global_kernel<<<blocks, 256, 0, st>>>(z, mask, N, T, L, K, counts);
        return check_launch("transition_counts");
    }
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
