// Parameter updates from (all-reduced) sufficient statistics, always in float64:
//   AR parameters  (Ab, Q) ~ MNIW posterior      jax_moseq.models.arhmm.gibbs.resample_ar_params / sample_mniw
//   transitions    (betas, pi), sticky HDP-HMM   jax_moseq.utils.transitions.resample_hdp_transitions
//   obs variance   sigmasq                       jax_moseq.models.keypoint_slds.gibbs.resample_obs_variance
// Every rank draws the same parameters from the same statistics and the same Philox key, so no
// broadcast follows the all-reduce.
#include "common.cuh"
#include "../../include/kpms_b200.h"

namespace kpms {

// ---- warp-level float64 dense helpers on shared-memory matrices (row-major, leading dim ld) ----
__device__ inline void w_chol(double* A, int n, int ld, int lane) {      // lower, in place
    for (int j = 0; j < n; ++j) {
        for (int r = j + lane; r < n; r += 32) {
            double s = A[r * ld + j];
            for (int p = 0; p < j; ++p) s -= A[r * ld + p] * A[j * ld + p];
            A[r * ld + j] = s;
        }
        __syncwarp();
        const double piv = sqrt(A[j * ld + j]);
        __syncwarp();
        for (int r = j + lane; r < n; r += 32) A[r * ld + j] = (r == j) ? piv : A[r * ld + j] / piv;
        __syncwarp();
    }
    for (int idx = lane; idx < n * n; idx += 32) { int r = idx / n, c = idx % n; if (c > r) A[r * ld + c] = 0.0; }
    __syncwarp();
}

__device__ inline void w_tri_inv(const double* Lm, double* Li, int n, int ld, int lane) {   // Li = Lm^-1 (lower)
    for (int c = lane; c < n; c += 32) {
        for (int r = 0; r < n; ++r) {
            if (r < c) { Li[r * ld + c] = 0.0; continue; }
            double s = (r == c) ? 1.0 : 0.0;
            for (int p = c; p < r; ++p) s -= Lm[r * ld + p] * Li[p * ld + c];
            Li[r * ld + c] = s / Lm[r * ld + r];
        }
    }
    __syncwarp();
}

// C (m x p) = A (m x q) * B (q x p), optional transposes via strides
__device__ inline void w_matmul(const double* A, int ars, int acs, const double* B, int brs, int bcs, double* C,
                                int ldc, int m, int q, int p, int lane) {
    for (int idx = lane; idx < m * p; idx += 32) {
        int r = idx / p, c = idx % p;
        double s = 0.0;
        for (int e = 0; e < q; ++e) s = fma(A[r * ars + e * acs], B[e * brs + c * bcs], s);
        C[r * ldc + c] = s;
    }
    __syncwarp();
}

// ---------------------------------------------------------------------------
// AR parameters: one warp per state
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(32)
ar_params_kernel(const double* __restrict__ gram, const double* __restrict__ K_0, const double* __restrict__ M_0,
                 const double* __restrict__ S_0, double nu_0, const double* __restrict__ w_G,
                 const double* __restrict__ w_B, const double* __restrict__ g_chi, SeedArg seed, int d, int L,
                 double* __restrict__ Ab, double* __restrict__ Q) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int n = d * L, p = n + 1, F = n + d + 1;
    const int k = blockIdx.x, lane = threadIdx.x;
    double* M1 = reinterpret_cast<double*>(smem_raw);   // p x p
    double* M2 = M1 + p * p;                            // p x p
    double* M3 = M2 + p * p;                            // p x p
    double* K0i = M3 + p * p;                           // p x p
    double* Mn = K0i + p * p;                           // d x p
    double* Tm = Mn + d * p;                            // d x p
    double* Sn = Tm + d * p;                            // d x d
    double* Zm = Sn + d * d;                            // d x d
    double* Zi = Zm + d * d;                            // d x d
    double* Qs = Zi + d * d;                            // d x d
    double* Gn = Qs + d * d;                            // d x p (normals)
    const double* G = gram + (size_t)k * F * F;
    // feature order [phi (n) | y (d) | 1]: regression inputs are in(j) = j < n ? j : n + d
    auto in = [&](int j) { return j < n ? j : n + d; };
    // K0i = K_0^-1
    for (int idx = lane; idx < p * p; idx += 32) M1[idx] = K_0[idx];
    __syncwarp();
    w_chol(M1, p, p, lane);
    w_tri_inv(M1, M2, p, p, lane);
    w_matmul(M2, 1, p, M2, p, 1, K0i, p, p, p, p, lane);          // Li' Li
    // Kni = K0i + Sxx  (M1); Kn = Kni^-1 (M3)
    for (int idx = lane; idx < p * p; idx += 32) {
        int r = idx / p, c = idx % p;
        M1[idx] = K0i[idx] + G[in(r) * F + in(c)];
    }
    __syncwarp();
    for (int idx = lane; idx < p * p; idx += 32) M3[idx] = M1[idx];   // keep Kni in M3 for S_n
    __syncwarp();
    w_chol(M1, p, p, lane);
    w_tri_inv(M1, M2, p, p, lane);
    w_matmul(M2, 1, p, M2, p, 1, M1, p, p, p, p, lane);           // Kn in M1
    for (int idx = lane; idx < p * p; idx += 32) {                // symmetrise
        int r = idx / p, c = idx % p;
        if (c < r) { double a = 0.5 * (M1[r * p + c] + M1[c * p + r]); M1[r * p + c] = a; M1[c * p + r] = a; }
    }
    __syncwarp();
    // Tm = M_0 K0i + Syx ; Mn = Tm Kn
    for (int idx = lane; idx < d * p; idx += 32) {
        int r = idx / p, c = idx % p;
        double s = G[(n + r) * F + in(c)];
        for (int e = 0; e < p; ++e) s = fma(M_0[r * p + e], K0i[e * p + c], s);
        Tm[idx] = s;
    }
    __syncwarp();
    w_matmul(Tm, p, 1, M1, p, 1, Mn, p, d, p, p, lane);
    // Sn = S_0 + Syy + M_0 K0i M_0' - Mn Kni Mn'
    w_matmul(Mn, p, 1, M3, p, 1, Tm, p, d, p, p, lane);           // Tm = Mn Kni
    for (int idx = lane; idx < d * p; idx += 32) {                // Gn = M_0 K0i (scratch)
        int r = idx / p, c = idx % p;
        double s = 0.0;
        for (int e = 0; e < p; ++e) s = fma(M_0[r * p + e], K0i[e * p + c], s);
        Gn[idx] = s;
    }
    __syncwarp();
    for (int idx = lane; idx < d * d; idx += 32) {
        int r = idx / d, c = idx % d;
        double s = S_0[idx] + G[(n + r) * F + (n + c)];
        for (int e = 0; e < p; ++e) s += Gn[r * p + e] * M_0[c * p + e] - Tm[r * p + e] * Mn[c * p + e];
        Sn[idx] = s;
    }
    __syncwarp();
    for (int idx = lane; idx < d * d; idx += 32) {
        int r = idx / d, c = idx % d;
        if (c < r) { double a = 0.5 * (Sn[r * d + c] + Sn[c * d + r]); Sn[r * d + c] = a; Sn[c * d + r] = a; }
    }
    __syncwarp();
    const double nu = nu_0 + G[(n + d) * F + (n + d)];
    // Bartlett factor Z (lower), Q = Ls Z^-T Z^-1 Ls'
    for (int idx = lane; idx < d * d; idx += 32) {
        int r = idx / d, c = idx % d;
        double val = 0.0;
        if (c < r) {
            if (w_B) val = w_B[(size_t)k * d * d + idx];
            else { Philox gen(seed, KPMS_STREAM_AR_B, (uint64_t)k * d * d + idx); double a0, a1; philox_normal2(gen, a0, a1); val = a0; }
        } else if (c == r) {
            Philox gen(seed, KPMS_STREAM_AR_CHI, (uint64_t)k * d + r);
            double gm = gamma_draw<double>(0.5 * (nu - r), g_chi ? g_chi + ((size_t)k * d + r) * KPMS_GAMMA_TAPE : nullptr, gen);
            val = sqrt(2.0 * gm);
        }
        Zm[idx] = val;
    }
    for (int idx = lane; idx < d * p; idx += 32) {
        if (w_G) Gn[idx] = w_G[(size_t)k * d * p + idx];
        else { Philox gen(seed, KPMS_STREAM_AR_G, (uint64_t)k * d * p + idx); double a0, a1; philox_normal2(gen, a0, a1); Gn[idx] = a0; }
    }
    __syncwarp();
    w_tri_inv(Zm, Zi, d, d, lane);
    w_chol(Sn, d, d, lane);                                        // Ls
    w_matmul(Sn, d, 1, Zi, 1, d, Zm, d, d, d, d, lane);            // Zm = Ls Zi'
    w_matmul(Zm, d, 1, Zm, 1, d, Qs, d, d, d, d, lane);            // Qs = T T'
    for (int idx = lane; idx < d * d; idx += 32) {
        int r = idx / d, c = idx % d;
        if (c < r) { double a = 0.5 * (Qs[r * d + c] + Qs[c * d + r]); Qs[r * d + c] = a; Qs[c * d + r] = a; }
    }
    __syncwarp();
    for (int idx = lane; idx < d * d; idx += 32) Q[(size_t)k * d * d + idx] = Qs[idx];
    __syncwarp();
    // Ab = Mn + chol(Q) Gn chol(Kn)'
    w_chol(Qs, d, d, lane);
    w_chol(M1, p, p, lane);
    w_matmul(Qs, d, 1, Gn, p, 1, Tm, p, d, d, p, lane);            // Tm = Lq Gn
    for (int idx = lane; idx < d * p; idx += 32) {
        int r = idx / p, c = idx % p;
        double s = Mn[idx];
        for (int e = 0; e <= c; ++e) s = fma(Tm[r * p + e], M1[c * p + e], s);
        Ab[(size_t)k * d * p + idx] = s;
    }
}

// ---------------------------------------------------------------------------
// transitions
// ---------------------------------------------------------------------------
// exclusive prefix of the flattened counts (tape offsets) and of the diagonal; one block of 1024 threads
__global__ void __launch_bounds__(1024)
count_prefix_kernel(const int* __restrict__ counts, int K, long long* __restrict__ starts,
                    long long* __restrict__ dstarts) {
    __shared__ long long part[1024];
    const int tid = threadIdx.x, total = K * K;
    const int per = (total + 1023) / 1024;
    const int lo = min(tid * per, total), hi = min(lo + per, total);
    long long sum = 0;
    for (int i = lo; i < hi; ++i) sum += counts[i];
    part[tid] = sum;
    __syncthreads();
    for (int off = 1; off < 1024; off <<= 1) {
        long long add = (tid >= off) ? part[tid - off] : 0;
        __syncthreads();
        part[tid] += add;
        __syncthreads();
    }
    long long run = part[tid] - sum;
    for (int i = lo; i < hi; ++i) { starts[i] = run; run += counts[i]; }
    if (tid == 0) {
        long long r2 = 0;
        for (int i = 0; i < K; ++i) { dstarts[i] = r2; r2 += counts[i * K + i]; }
    }
}

// CRP table counts: one block per row, one warp per column (strided)
__global__ void __launch_bounds__(256)
crp_tables_kernel(const int* __restrict__ counts, const double* __restrict__ betas, double alpha, double kappa,
                  const long long* __restrict__ starts, const double* __restrict__ u_crp, SeedArg seed, int K,
                  int* __restrict__ m) {
    const int i = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    constexpr int BIG = 4096;                  // entries with more customers than this are drawn by the whole CTA
    __shared__ int red[32];
    auto draw = [&](int j, int r, double conc) -> int {
        double u;
        if (u_crp) u = u_crp[starts[i * K + j] + r];
        else { Philox gen(seed, KPMS_STREAM_CRP, ((uint64_t)(i * K + j) << 32) | (uint32_t)r); double u2; philox_uniform2(gen, u, u2); }
        return (u < conc / ((double)r + conc)) ? 1 : 0;
    };
    // the sticky diagonal holds most of the transitions (tens of thousands per state on a large cohort)
    for (int j = 0; j < K; ++j) {
        const int nn = counts[i * K + j];
        if (nn <= BIG) continue;               // uniform across the CTA
        const double conc = alpha * betas[j] + (i == j ? kappa : 0.0);
        int cnt = 0;
        for (int r = threadIdx.x; r < nn; r += blockDim.x) cnt += draw(j, r, conc);
        cnt = block_sum(cnt, red);
        if (threadIdx.x == 0) m[i * K + j] = cnt;
        __syncthreads();
    }
    for (int j = warp; j < K; j += nw) {
        const int nn = counts[i * K + j];
        if (nn > BIG) continue;
        const double conc = alpha * betas[j] + (i == j ? kappa : 0.0);
        int cnt = 0;
        for (int r = lane; r < nn; r += 32) cnt += draw(j, r, conc);
        cnt = warp_sum(cnt);
        if (lane == 0) m[i * K + j] = cnt;
    }
}

// overrides w_i ~ Bin(m_ii, rho / (rho + beta_i (1 - rho))): one warp per state, written into m's diagonal
__global__ void __launch_bounds__(32)
overrides_kernel(int* __restrict__ m, const double* __restrict__ betas_in, double alpha, double kappa,
                 const long long* __restrict__ dstarts, const double* __restrict__ u_bin, SeedArg seed, int K,
                 int* __restrict__ wov) {
    const int i = blockIdx.x, lane = threadIdx.x;
    const double rho = kappa / (alpha + kappa);
    const int mii = m[i * K + i];
    const double pov = rho / (rho + betas_in[i] * (1.0 - rho));
    int w = 0;
    for (int r = lane; r < mii; r += 32) {
        double u;
        if (u_bin) u = u_bin[dstarts[i] + r];
        else { Philox gen(seed, KPMS_STREAM_BIN, ((uint64_t)i << 32) | (uint32_t)r); double u2; philox_uniform2(gen, u, u2); }
        w += (u < pov) ? 1 : 0;
    }
    w = warp_sum(w);
    if (lane == 0) wov[i] = w;
}

// betas ~ Dir(gamma/K + colsum(m) - w); single block of >= K threads
__global__ void betas_kernel(const int* __restrict__ m, const int* __restrict__ wov, double gamma,
                             const double* __restrict__ g_beta, SeedArg seed, int K,
                             double* __restrict__ betas_out) {
    extern __shared__ double sh[];
    double* gb = sh;             // K gammas
    __shared__ double red[32];
    const int tid = threadIdx.x;
    double g = 0.0;
    if (tid < K) {
        double colsum = 0.0;
        for (int i = 0; i < K; ++i) colsum += (double)m[i * K + tid];
        colsum -= (double)wov[tid];
        Philox gen(seed, KPMS_STREAM_BETA, (uint64_t)tid);
        g = gamma_draw<double>(gamma / K + colsum, g_beta ? g_beta + (size_t)tid * KPMS_GAMMA_TAPE : nullptr, gen);
        gb[tid] = g;
    }
    double tot = block_sum(g, red);
    if (tid < K) betas_out[tid] = gb[tid] / tot;
}

// pi_i ~ Dir(alpha betas + kappa e_i + N_i.); one block per row
__global__ void pi_rows_kernel(const int* __restrict__ counts, const double* __restrict__ betas, double alpha,
                               double kappa, const double* __restrict__ g_pi, SeedArg seed, int K,
                               double* __restrict__ pi) {
    __shared__ double red[32];
    const int i = blockIdx.x, j = threadIdx.x;
    double g = 0.0;
    if (j < K) {
        const double a = alpha * betas[j] + (i == j ? kappa : 0.0) + (double)counts[i * K + j];
        Philox gen(seed, KPMS_STREAM_PI, (uint64_t)i * K + j);
        g = gamma_draw<double>(a, g_pi ? g_pi + ((size_t)i * K + j) * KPMS_GAMMA_TAPE : nullptr, gen);
    }
    double tot = block_sum(g, red);
    if (j < K) pi[(size_t)i * K + j] = g / tot;
}

__global__ void sigmasq_kernel(const double* __restrict__ stats, double nu_sigma, double sigmasq_0, int Dk,
                               const double* __restrict__ g_sig, SeedArg seed, int k, double* __restrict__ sigmasq) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= k) return;
    const double degs = nu_sigma + KPMS_OBS_DOF(Dk) * stats[k];
    const double variance = stats[j] + nu_sigma * sigmasq_0;
    Philox gen(seed, KPMS_STREAM_SIGMA, (uint64_t)j);
    const double g = gamma_draw<double>(0.5 * degs, g_sig ? g_sig + (size_t)j * KPMS_GAMMA_TAPE : nullptr, gen);
    sigmasq[j] = variance / (2.0 * g);
}

}  // namespace kpms

using namespace kpms;

extern "C" {

int kpms_resample_ar_params(const double* gram, const double* K_0, const double* M_0, const double* S_0, double nu_0,
                            const double* w_G, const double* w_B, const double* g_chi, uint64_t seed_, const uint64_t* seed_dev, int K,
                            int d, int L, double* Ab, double* Q, void* stream) {
    const SeedArg seed(seed_, seed_dev);
    const int n = d * L, p = n + 1;
    size_t smem = ((size_t)4 * p * p + 3 * d * p + 4 * d * d) * sizeof(double);
    if (smem > 220 * 1024) return set_error(-3, "resample_ar_params: (latent_dim, nlags) = (%d, %d) too large", d, L);
    cudaFuncSetAttribute(ar_params_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    { cudaStream_t st = (cudaStream_t)stream; KPMS_LAUNCH("ar_params", st);
    ar_params_kernel<<<K, 32, smem, st>>>(gram, K_0, M_0, S_0, nu_0, w_G, w_B, g_chi, seed, d, L, Ab, Q); }
    return check_launch("resample_ar_params");
}

size_t kpms_transitions_workspace_bytes(int K) {
    return kpms::align_up((size_t)K * K * sizeof(long long), 256) + kpms::align_up((size_t)K * sizeof(long long), 256) +
           kpms::align_up((size_t)K * K * sizeof(int), 256) + kpms::align_up((size_t)K * sizeof(int), 256);
}

int kpms_resample_hdp_transitions(const int32_t* counts, const double* betas_in, double alpha, double kappa,
                                  double gamma, const double* u_crp, const double* u_bin, const double* g_beta,
                                  const double* g_pi, uint64_t seed_, const uint64_t* seed_dev, int K, double* betas_out, double* pi,
                                  void* ws, void* stream) {
    const SeedArg seed(seed_, seed_dev);
    cudaStream_t st = (cudaStream_t)stream;
    if (K > 1024) return set_error(-3, "resample_hdp_transitions: num_states %d > 1024", K);
    char* base = reinterpret_cast<char*>(ws);
    long long* starts = reinterpret_cast<long long*>(base);
    long long* dstarts = reinterpret_cast<long long*>(base + kpms::align_up((size_t)K * K * sizeof(long long), 256));
    int* m = reinterpret_cast<int*>(base + kpms::align_up((size_t)K * K * sizeof(long long), 256) +
                                    kpms::align_up((size_t)K * sizeof(long long), 256));
    int* wov = m + kpms::align_up((size_t)K * K * sizeof(int), 256) / sizeof(int);
    { KPMS_LAUNCH("trans_prefix", st); count_prefix_kernel<<<1, 1024, 0, st>>>(counts, K, starts, dstarts); }
    { KPMS_LAUNCH("trans_crp", st); crp_tables_kernel<<<K, 256, 0, st>>>(counts, betas_in, alpha, kappa, starts, u_crp, seed, K, m); }
    int threads = (K + 31) / 32 * 32;
    { KPMS_LAUNCH("trans_overrides", st); overrides_kernel<<<K, 32, 0, st>>>(m, betas_in, alpha, kappa, dstarts, u_bin, seed, K, wov); }
    { KPMS_LAUNCH("trans_betas", st);
    betas_kernel<<<1, threads, K * sizeof(double), st>>>(m, wov, gamma, g_beta, seed, K, betas_out); }
    { KPMS_LAUNCH("trans_pi", st); pi_rows_kernel<<<K, threads, 0, st>>>(counts, betas_out, alpha, kappa, g_pi, seed, K, pi); }
    return check_launch("resample_hdp_transitions");
}

int kpms_resample_obs_variance(const double* stats, double nu_sigma, double sigmasq_0, int Dk, const double* g_sig,
                               uint64_t seed_, const uint64_t* seed_dev, int k, double* sigmasq, void* stream) {
    const SeedArg seed(seed_, seed_dev);
    { cudaStream_t st = (cudaStream_t)stream; KPMS_LAUNCH("obs_variance", st);
    sigmasq_kernel<<<(k + 63) / 64, 64, 0, st>>>(stats, nu_sigma, sigmasq_0, Dk, g_sig, seed, k, sigmasq); }
    return check_launch("resample_obs_variance");
}

}  // extern "C"
