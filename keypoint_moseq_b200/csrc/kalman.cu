// Continuous-state resampling: Kalman forward filter + backward sampling over the
// (latent_dim x nlags) augmented AR state.  Replaces
// jax_moseq.models.keypoint_slds.resample_continuous_stateseqs -> utils.kalman.kalman_sample
// (reached from keypoint_moseq/fitting.py:25; flags at fitting.py:47-60, :260-261).
//
// The FFBS is split so that only the covariance recursion is serial:
//   K1a obs_info      (parallel over frames)  J_t = C~' R_t^-1 C~ -> chol, r_t = C~' R_t^-1 (y~_t - d~)
//   K1b forward       (serial, CTA per chain) covariance-form filter, stashes (m_t, P_t)
//   K1c backprep      (parallel, warp/frame)  G_t = P_t A' P'^-1, chol(Sigma_t), h_t = m_t - G_t(A m_t + b) + L_t w_t
//   K1d affine        (serial, warp per chain) xi_t = G_t xi_{t+1} + h_t
// which draws exactly the sample mu_t + chol(Sigma_t) w_t of the sequential sampler.
//
// Frame index i = t - (L-1), i in [0, Tx), Tx = T - L + 1.  z[i] governs the transition i -> i+1.
// Layouts: info (N,Tx,REC) REC = d(d+1)/2 + d;  stash_m (N,Tx,n);  stash_S (N,Tx,n(n+1)/2) packed lower;
//          GH (N,Tx,RECS): [GT (n*n) with GT[c][r] = G[r][c] | h (n) | pad], RECS*sizeof(R) % 16 == 0.
#include "common.cuh"
#include "../../include/kpms_b200.h"

namespace kpms {

__device__ __forceinline__ void tri_unpack(int q, int& i, int& j) {
    i = (int)((sqrtf(8.0f * (float)q + 1.0f) - 1.0f) * 0.5f);
    while ((i + 1) * (i + 2) / 2 <= q) ++i;
    while (i * (i + 1) / 2 > q) --i;
    j = q - i * (i + 1) / 2;
}

// ---------------------------------------------------------------------------
// K1a: per-frame observation information
// ---------------------------------------------------------------------------
template <typename R, int D_, int DK>
__global__ void __launch_bounds__(128)
obs_info_kernel(const R* __restrict__ Y, const int* __restrict__ mask, const R* __restrict__ v,
                const R* __restrict__ h, const R* __restrict__ s, const R* __restrict__ sigmasq,
                const R* __restrict__ Ct, int N, int T, int k, int L, R* __restrict__ info) {
    constexpr int NP = D_ * (D_ + 1) / 2, REC = NP + D_;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    R* Cs = reinterpret_cast<R*>(smem_raw);            // (k*DK) x (D_+1)
    R* sg = Cs + (size_t)k * DK * (D_ + 1);            // k
    for (int i = threadIdx.x; i < k * DK * (D_ + 1); i += blockDim.x) Cs[i] = Ct[i];
    for (int i = threadIdx.x; i < k; i += blockDim.x) sg[i] = sigmasq[i];
    __syncthreads();
    const int Tx = T - L + 1;
    long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= (long long)N * Tx) return;
    const int nn = (int)(g / Tx), i = (int)(g % Tx);
    const int t = i + L - 1;
    const size_t ft = (size_t)nn * T + t;
    R* out = info + (size_t)g * REC;
    if (mask[ft] == 0) {
#pragma unroll 1
        for (int q = 0; q < REC; ++q) out[q] = (R)0;
        return;
    }
    R J[NP], r[D_];
#pragma unroll
    for (int q = 0; q < NP; ++q) J[q] = (R)0;
#pragma unroll
    for (int q = 0; q < D_; ++q) r[q] = (R)0;
    R vv[DK];
#pragma unroll
    for (int c = 0; c < DK; ++c) vv[c] = v[ft * DK + c];
    R sn, cs;
    sincos_r<R>(h[ft], sn, cs);
    for (int j = 0; j < k; ++j) {
        const R w = (R)1 / (s[ft * k + j] * sg[j]);
        R yc[DK];
#pragma unroll
        for (int c = 0; c < DK; ++c) yc[c] = Y[(ft * k + j) * DK + c] - vv[c];
        R y0 = cs * yc[0] + sn * yc[1], y1 = -sn * yc[0] + cs * yc[1];
        yc[0] = y0;
        yc[1] = y1;
#pragma unroll
        for (int c = 0; c < DK; ++c) {
            const R* crow = Cs + (size_t)(j * DK + c) * (D_ + 1);
            R cr[D_];
#pragma unroll
            for (int a = 0; a < D_; ++a) cr[a] = crow[a];
            const R wres = w * (yc[c] - crow[D_]);
#pragma unroll
            for (int a = 0; a < D_; ++a) {
                r[a] = fma(wres, cr[a], r[a]);
                const R wa = w * cr[a];
#pragma unroll
                for (int b = 0; b <= a; ++b) J[a * (a + 1) / 2 + b] = fma(wa, cr[b], J[a * (a + 1) / 2 + b]);
            }
        }
    }
    // in-register Cholesky J = Lj Lj'; a non-positive pivot (rank-deficient C~) zeroes its column
#pragma unroll
    for (int c = 0; c < D_; ++c) {
        R sdiag = J[c * (c + 1) / 2 + c];
#pragma unroll
        for (int p = 0; p < c; ++p) sdiag -= J[c * (c + 1) / 2 + p] * J[c * (c + 1) / 2 + p];
        const bool ok = sdiag > (R)0;
        const R inv = ok ? rsqrt_r<R>(sdiag) : (R)0;
        J[c * (c + 1) / 2 + c] = ok ? sdiag * inv : (R)0;
#pragma unroll
        for (int a = c + 1; a < D_; ++a) {
            R val = J[a * (a + 1) / 2 + c];
#pragma unroll
            for (int p = 0; p < c; ++p) val -= J[a * (a + 1) / 2 + p] * J[c * (c + 1) / 2 + p];
            J[a * (a + 1) / 2 + c] = val * inv;
        }
    }
#pragma unroll
    for (int q = 0; q < NP; ++q) out[q] = J[q];
#pragma unroll
    for (int q = 0; q < D_; ++q) out[NP + q] = r[q];
}

// ---------------------------------------------------------------------------
// K1b: serial covariance-form filter, one CTA per chain.
// Measurement update through the well-conditioned d x d matrix B = I + Lj' P_nn Lj:
//   U = P[:,new] Lj, V = U chol(B)^-T, P+ = P - V V', m+ = m + P+[:,new] (r - J m_new)
// Predict with the companion structure (shift + d dense rows).
// ---------------------------------------------------------------------------
template <typename R, int D_, int L_>
struct FwdSmem {
    static constexpr int n = D_ * L_, LD = n | 1, NP = D_ * (D_ + 1) / 2, REC = NP + D_,
                         NP2 = n * (n + 1) / 2, NA = D_ * (n + 1);
    static constexpr size_t elems = 2 * n * LD + 2 * n + 2 * n * D_ + D_ * LD + D_ * D_ + 2 * D_ +
                                    2 * REC + 2 * NA + 2 * D_ * D_;
    static constexpr size_t bytes = elems * sizeof(R) + NP2 * sizeof(unsigned short) + 16;
};

template <typename R, int D_, int L_>
__global__ void __launch_bounds__(256)
kalman_forward_kernel(const R* __restrict__ info, const int* __restrict__ mask, const int* __restrict__ z,
                      const R* __restrict__ Ab, const R* __restrict__ Q, R jitter, int T,
                      R* __restrict__ stash_m, R* __restrict__ stash_S) {
    typedef FwdSmem<R, D_, L_> SM;
    constexpr int n = SM::n, LD = SM::LD, NP = SM::NP, REC = SM::REC, NP2 = SM::NP2, NA = SM::NA;
    constexpr int NT = 256;
    constexpr int NO = n - D_;                       // first index of the newest block
    constexpr int NPRE = (REC + NA + D_ * D_ + NT - 1) / NT;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    R* P0 = reinterpret_cast<R*>(smem_raw);          // predicted covariance
    R* P1 = P0 + n * LD;                             // filtered covariance
    R* m0 = P1 + n * LD;
    R* m1 = m0 + n;
    R* U = m1 + n;
    R* V = U + n * D_;
    R* Tm = V + n * D_;                              // D_ x LD
    R* Bm = Tm + D_ * LD;
    R* cvec = Bm + D_ * D_;
    R* tmp1 = cvec + D_;
    R* inf = tmp1 + D_;                              // 2 x REC
    R* Az = inf + 2 * REC;                           // 2 x NA
    R* Qz = Az + 2 * NA;                             // 2 x D_*D_
    unsigned short* ij = reinterpret_cast<unsigned short*>(Qz + 2 * D_ * D_);
    const int nn = blockIdx.x, tid = threadIdx.x;
    const int Tx = T - L_ + 1;
    const R* inf_g = info + (size_t)nn * Tx * REC;
    const int* mk = mask + (size_t)nn * T + (L_ - 1);
    const int* zz = z + (size_t)nn * (Tx - 1);
    R* sm_g = stash_m + (size_t)nn * Tx * n;
    R* sS_g = stash_S + (size_t)nn * Tx * NP2;
    const R eps = (R)KPMS_EPS_SHIFT + jitter;

    for (int q = tid; q < NP2; q += NT) { int i, j; tri_unpack(q, i, j); ij[q] = (unsigned short)((i << 8) | j); }
    for (int w = tid; w < n * n; w += NT) { int i = w / n, j = w % n; P0[i * LD + j] = (i == j) ? (R)KPMS_X_PRIOR_VAR : (R)0; }
    for (int w = tid; w < n; w += NT) m0[w] = (R)0;

    // staged operands for step i live in buffer (i & 1)
    auto stage_load = [&](int i, int zi, R* pre) {
#pragma unroll
        for (int q = 0; q < NPRE; ++q) {
            int w = tid + q * NT;
            R val = (R)0;
            if (w < REC) val = inf_g[(size_t)i * REC + w];
            else if (w < REC + NA) { if (zi >= 0) val = Ab[(size_t)zi * NA + (w - REC)]; }
            else if (w < REC + NA + D_ * D_) { if (zi >= 0) val = Q[(size_t)zi * D_ * D_ + (w - REC - NA)]; }
            pre[q] = val;
        }
    };
    auto stage_store = [&](int b, const R* pre) {
#pragma unroll
        for (int q = 0; q < NPRE; ++q) {
            int w = tid + q * NT;
            if (w < REC) inf[b * REC + w] = pre[q];
            else if (w < REC + NA) Az[b * NA + (w - REC)] = pre[q];
            else if (w < REC + NA + D_ * D_) Qz[b * D_ * D_ + (w - REC - NA)] = pre[q];
        }
    };
    R pre[NPRE];
    int z_next = (Tx > 1) ? zz[0] : -1;              // z for step i+1 while running step i
    stage_load(0, z_next, pre);
    stage_store(0, pre);
    z_next = (Tx > 2) ? zz[1] : -1;
    int mk_cur = mk[0];
    __syncthreads();

    for (int i = 0; i < Tx; ++i) {
        const int b = i & 1;
        const bool last = (i == Tx - 1);
        const int mk_next = last ? 0 : mk[i + 1];
        if (!last) stage_load(i + 1, z_next, pre);
        const int z_next2 = (i + 2 < Tx - 1) ? zz[i + 2] : -1;
        const R* fi = inf + b * REC;
        const R* A = Az + b * NA;
        const R* Qs = Qz + b * D_ * D_;
        if (mk_cur != 0) {
            // ---- A: U = P[:,new] Lj ; tmp1 = Lj' m_new
            for (int w = tid; w < n * D_ + D_; w += NT) {
                if (w < n * D_) {
                    int r = w / D_, c = w % D_;
                    R acc = 0;
                    for (int e = c; e < D_; ++e) acc = fma(P0[r * LD + NO + e], fi[e * (e + 1) / 2 + c], acc);
                    U[w] = acc;
                } else {
                    int c = w - n * D_;
                    R acc = 0;
                    for (int e = c; e < D_; ++e) acc = fma(fi[e * (e + 1) / 2 + c], m0[NO + e], acc);
                    tmp1[c] = acc;
                }
            }
            __syncthreads();
            // ---- B: Bm = I + Lj' U[new,:] ; cvec = r - Lj tmp1
            for (int w = tid; w < D_ * D_ + D_; w += NT) {
                if (w < D_ * D_) {
                    int a = w / D_, c = w % D_;
                    R acc = (a == c) ? (R)1 : (R)0;
                    for (int e = a; e < D_; ++e) acc = fma(fi[e * (e + 1) / 2 + a], U[(NO + e) * D_ + c], acc);
                    Bm[w] = acc;
                } else {
                    int a = w - D_ * D_;
                    R acc = fi[NP + a];
                    for (int c = 0; c <= a; ++c) acc = fma(-fi[a * (a + 1) / 2 + c], tmp1[c], acc);
                    cvec[a] = acc;
                }
            }
            __syncthreads();
            // ---- C: every lane of warp 0 factors B redundantly in registers, then V = U Lb^-T by rows
            if (tid < 32) {
                R Lb[NP];
#pragma unroll
                for (int a = 0; a < D_; ++a)
#pragma unroll
                    for (int c = 0; c <= a; ++c) Lb[a * (a + 1) / 2 + c] = Bm[a * D_ + c];
#pragma unroll
                for (int c = 0; c < D_; ++c) {
                    R sd = Lb[c * (c + 1) / 2 + c];
#pragma unroll
                    for (int p = 0; p < c; ++p) sd -= Lb[c * (c + 1) / 2 + p] * Lb[c * (c + 1) / 2 + p];
                    const R inv = rsqrt_r<R>(sd);
                    Lb[c * (c + 1) / 2 + c] = inv;               // inverse diagonal
#pragma unroll
                    for (int a = c + 1; a < D_; ++a) {
                        R val = Lb[a * (a + 1) / 2 + c];
#pragma unroll
                        for (int p = 0; p < c; ++p) val -= Lb[a * (a + 1) / 2 + p] * Lb[c * (c + 1) / 2 + p];
                        Lb[a * (a + 1) / 2 + c] = val * inv;
                    }
                }
                for (int r = tid; r < n; r += 32) {
                    R vr[D_];
#pragma unroll
                    for (int c = 0; c < D_; ++c) {
                        R val = U[r * D_ + c];
#pragma unroll
                        for (int p = 0; p < c; ++p) val -= Lb[c * (c + 1) / 2 + p] * vr[p];
                        vr[c] = val * Lb[c * (c + 1) / 2 + c];
                        V[r * D_ + c] = vr[c];
                    }
                }
            }
            __syncthreads();
            // ---- D: P+ = P - V V' (lower computed, mirrored), stash packed lower
            for (int q = tid; q < NP2; q += NT) {
                int r = ij[q] >> 8, c = ij[q] & 255;
                R acc = P0[r * LD + c];
#pragma unroll
                for (int e = 0; e < D_; ++e) acc = fma(-V[r * D_ + e], V[c * D_ + e], acc);
                P1[r * LD + c] = acc;
                P1[c * LD + r] = acc;
                sS_g[(size_t)i * NP2 + q] = acc;
            }
            __syncthreads();
            // ---- E: m+ = m + P+[:,new] cvec ; Tm = A P+
            for (int w = tid; w < n + (last ? 0 : D_ * n); w += NT) {
                if (w < n) {
                    R acc = m0[w];
#pragma unroll
                    for (int c = 0; c < D_; ++c) acc = fma(P1[w * LD + NO + c], cvec[c], acc);
                    m1[w] = acc;
                    sm_g[(size_t)i * n + w] = acc;
                } else {
                    int a = (w - n) / n, c = (w - n) % n;
                    R acc = 0;
                    for (int e = 0; e < n; ++e) acc = fma(A[a * (n + 1) + e], P1[e * LD + c], acc);
                    Tm[a * LD + c] = acc;
                }
            }
            __syncthreads();
            // ---- F: predict into P0 / m0
            if (!last) {
                for (int w = tid; w < NO * NO; w += NT) {
                    int r = w / NO, c = w % NO;
                    P0[r * LD + c] = P1[(r + D_) * LD + c + D_] + ((r == c) ? eps : (R)0);
                }
                for (int w = tid; w < D_ * NO; w += NT) {
                    int a = w / NO, c = w % NO;
                    R val = Tm[a * LD + c + D_];
                    P0[(NO + a) * LD + c] = val;
                    P0[c * LD + NO + a] = val;
                }
                for (int w = tid; w < NP; w += NT) {
                    int a = ij[w] >> 8, c = ij[w] & 255;
                    R acc = Qs[a * D_ + c] + ((a == c) ? jitter : (R)0);
                    for (int e = 0; e < n; ++e) acc = fma(Tm[a * LD + e], A[c * (n + 1) + e], acc);
                    P0[(NO + a) * LD + NO + c] = acc;
                    P0[(NO + c) * LD + NO + a] = acc;
                }
                for (int w = tid; w < n; w += NT) {
                    if (w < NO) m0[w] = m1[w + D_];
                    else {
                        int a = w - NO;
                        R acc = A[a * (n + 1) + n];
                        for (int e = 0; e < n; ++e) acc = fma(A[a * (n + 1) + e], m1[e], acc);
                        m0[w] = acc;
                    }
                }
            }
        } else if (last) {
            for (int q = tid; q < NP2; q += NT) {
                int r = ij[q] >> 8, c = ij[q] & 255;
                sS_g[(size_t)i * NP2 + q] = P0[r * LD + c];
            }
            for (int w = tid; w < n; w += NT) sm_g[(size_t)i * n + w] = m0[w];
        }
        if (!last) stage_store(b ^ 1, pre);
        z_next = z_next2;
        mk_cur = mk_next;
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------
// warp-level dense kernels on shared-memory matrices (leading dimension LD odd)
// ---------------------------------------------------------------------------
template <typename R, int n, int LD>
__device__ inline void warp_cholesky(R* A, R* invdiag, int lane) {
    for (int j = 0; j < n; ++j) {
        for (int r = lane; r < n; r += 32) {
            if (r >= j) {
                R sacc = A[r * LD + j];
                for (int p = 0; p < j; ++p) sacc = fma(-A[r * LD + p], A[j * LD + p], sacc);
                A[r * LD + j] = sacc;
            }
        }
        __syncwarp();
        const R piv = A[j * LD + j];
        const R inv = rsqrt_r<R>(piv);
        __syncwarp();
        for (int r = lane; r < n; r += 32) {
            if (r > j) A[r * LD + j] *= inv;
            else if (r == j) { A[j * LD + j] = piv * inv; invdiag[j] = inv; }
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------
// K1c: backward preparation, one warp per frame
// ---------------------------------------------------------------------------
template <typename R, int D_, int L_>
struct PrepSmem {
    static constexpr int n = D_ * L_, LD = n | 1;
    static constexpr size_t per_warp = 3 * n * LD + 5 * n;
    // one backward record per frame: [GT | h | pad] in 16-byte units
    static constexpr int RECS = ((n * n + n) * (int)sizeof(R) + 15) / 16 * 16 / (int)sizeof(R);
};

template <typename R, int D_, int L_, int WARPS>
__global__ void __launch_bounds__(32 * WARPS)
kalman_backprep_kernel(const R* __restrict__ stash_m, const R* __restrict__ stash_S,
                       const int* __restrict__ mask, const int* __restrict__ z, const R* __restrict__ Ab,
                       const R* __restrict__ Q, R jitter, const R* __restrict__ w_tape, uint64_t seed, int N,
                       int T, R* __restrict__ GH) {
    typedef PrepSmem<R, D_, L_> SM;
    constexpr int n = SM::n, LD = SM::LD, NP2 = n * (n + 1) / 2, NO = n - D_, NA1 = n + 1;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    R* S = reinterpret_cast<R*>(smem_raw) + (size_t)warp * SM::per_warp;
    R* Pp = S + n * LD;
    R* Wt = Pp + n * LD;
    R* mv = Wt + n * LD;
    R* mp = mv + n;
    R* wv = mp + n;
    R* idP = wv + n;
    R* idS = idP + n;
    const int Tx = T - L_ + 1;
    const long long g = (long long)blockIdx.x * WARPS + warp;
    if (g >= (long long)N * Tx) return;
    const int nn = (int)(g / Tx), i = (int)(g % Tx);
    const bool last = (i == Tx - 1);
    R* Gout = GH + (size_t)g * SM::RECS;
    R* hout = Gout + n * n;
    if (!last && mask[(size_t)nn * T + (L_ - 1) + i] == 0) {
        for (int w = lane; w < n * n; w += 32) Gout[w] = ((w / n) == (w % n)) ? (R)1 : (R)0;
        for (int w = lane; w < n; w += 32) hout[w] = (R)0;
        return;
    }
    // operands
    const R* Sg = stash_S + (size_t)g * NP2;
    for (int q = lane; q < NP2; q += 32) {
        int r, c;
        tri_unpack(q, r, c);
        R val = Sg[q];
        S[r * LD + c] = val;
        S[c * LD + r] = val;
    }
    for (int w = lane; w < n; w += 32) {
        mv[w] = stash_m[(size_t)g * n + w];
        R wn;
        if (w_tape) wn = w_tape[(size_t)g * n + w];
        else {
            Philox gen(seed, KPMS_STREAM_X, (uint64_t)g * n + w);
            double a0, a1;
            philox_normal2(gen, a0, a1);
            wn = (R)a0;
        }
        wv[w] = wn;
    }
    __syncwarp();
    if (last) {
        warp_cholesky<R, n, LD>(S, idS, lane);
        for (int r = lane; r < n; r += 32) {
            R acc = mv[r];
            for (int c = 0; c <= r; ++c) acc = fma(S[r * LD + c], wv[c], acc);
            hout[r] = acc;
        }
        return;
    }
    const int zi = z[(size_t)nn * (Tx - 1) + i];
    const R* A = Ab + (size_t)zi * D_ * NA1;
    const R* Qk = Q + (size_t)zi * D_ * D_;
    const R eps = (R)KPMS_EPS_SHIFT + jitter;
    // Wt = Aaug S   (lane <-> column)
    for (int c = lane; c < n; c += 32) {
        for (int r = 0; r < NO; ++r) Wt[r * LD + c] = S[(r + D_) * LD + c];
        for (int a = 0; a < D_; ++a) {
            R acc = 0;
            for (int e = 0; e < n; ++e) acc = fma(__ldg(A + a * NA1 + e), S[e * LD + c], acc);
            Wt[(NO + a) * LD + c] = acc;
        }
    }
    // mp = Aaug m + b
    for (int r = lane; r < n; r += 32) {
        if (r < NO) mp[r] = mv[r + D_];
        else {
            int a = r - NO;
            R acc = __ldg(A + a * NA1 + n);
            for (int e = 0; e < n; ++e) acc = fma(__ldg(A + a * NA1 + e), mv[e], acc);
            mp[r] = acc;
        }
    }
    __syncwarp();
    // Pp = Wt Aaug' + Qaug   (lane <-> row)
    for (int r = lane; r < n; r += 32) {
        for (int c = 0; c < NO; ++c) Pp[r * LD + c] = Wt[r * LD + c + D_] + ((r == c) ? eps : (R)0);
        for (int a = 0; a < D_; ++a) {
            R acc = (r >= NO) ? (__ldg(Qk + (r - NO) * D_ + a) + ((r - NO == a) ? jitter : (R)0)) : (R)0;
            for (int e = 0; e < n; ++e) acc = fma(Wt[r * LD + e], __ldg(A + a * NA1 + e), acc);
            Pp[r * LD + NO + a] = acc;
        }
    }
    __syncwarp();
    warp_cholesky<R, n, LD>(Pp, idP, lane);
    // V = Lp^-1 Wt  (forward substitution, lane <-> column, in place)
    for (int c = lane; c < n; c += 32) {
        R col[n];
#pragma unroll
        for (int r = 0; r < n; ++r) {
            R acc = Wt[r * LD + c];
#pragma unroll
            for (int e = 0; e < r; ++e) acc = fma(-Pp[r * LD + e], col[e], acc);
            col[r] = acc * idP[r];
            Wt[r * LD + c] = col[r];
        }
    }
    __syncwarp();
    // Sigma = S - V'V (lower, in place over S; lane <-> column b, rows a >= b)
    for (int bcol = lane; bcol < n; bcol += 32) {
        R col[n];
#pragma unroll
        for (int e = 0; e < n; ++e) col[e] = Wt[e * LD + bcol];
        for (int a = bcol; a < n; ++a) {
            R acc = S[a * LD + bcol];
#pragma unroll
            for (int e = 0; e < n; ++e) acc = fma(-Wt[e * LD + a], col[e], acc);
            S[a * LD + bcol] = acc;
        }
    }
    __syncwarp();
    warp_cholesky<R, n, LD>(S, idS, lane);
    // X = Lp^-T V  (back substitution, lane <-> column, in place); GT[c][r] = X[c][r]
    for (int c = lane; c < n; c += 32) {
        R col[n];
#pragma unroll
        for (int r = n - 1; r >= 0; --r) {
            R acc = Wt[r * LD + c];
#pragma unroll
            for (int e = r + 1; e < n; ++e) acc = fma(-Pp[e * LD + r], col[e], acc);
            col[r] = acc * idP[r];
            Wt[r * LD + c] = col[r];
        }
    }
    __syncwarp();
    for (int w = lane; w < n * n; w += 32) Gout[w] = Wt[(w / n) * LD + (w % n)];
    // h = m - G mp + Ls w,  G[r][c] = X[c][r]
    for (int r = lane; r < n; r += 32) {
        R acc = mv[r];
        for (int c = 0; c < n; ++c) acc = fma(-Wt[c * LD + r], mp[c], acc);
        for (int c = 0; c <= r; ++c) acc = fma(S[r * LD + c], wv[c], acc);
        hout[r] = acc;
    }
}

// ---------------------------------------------------------------------------
// K1d: serial affine recursion, one warp per chain, operands streamed through a
// cp.async ring in shared memory
// ---------------------------------------------------------------------------
template <typename R, int D_, int L_, int STAGES>
__global__ void __launch_bounds__(32)
kalman_affine_kernel(const R* __restrict__ GH, int T, R* __restrict__ x) {
    constexpr int n = D_ * L_, NN = n * n;
    constexpr int RECP = PrepSmem<R, D_, L_>::RECS;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    R* ring = reinterpret_cast<R*>(smem_raw);        // STAGES x RECP
    R* xi = ring + (size_t)STAGES * RECP;            // n
    const int nn = blockIdx.x, lane = threadIdx.x;
    const int Tx = T - L_ + 1;
    const R* Gn = GH + (size_t)nn * Tx * RECP;
    R* xn = x + (size_t)nn * T * D_;
    auto issue = [&](int i) {
        if (i >= 0) {
            constexpr int CHUNKS = RECP * (int)sizeof(R) / 16;
            char* dst = reinterpret_cast<char*>(ring + (size_t)(i % STAGES) * RECP);
            const char* src = reinterpret_cast<const char*>(Gn + (size_t)i * RECP);
            for (int c = lane; c < CHUNKS; c += 32) {
                unsigned d32 = (unsigned)__cvta_generic_to_shared(dst + 16 * c);
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d32), "l"(src + 16 * c));
            }
        }
        asm volatile("cp.async.commit_group;\n" ::);
    };
    // terminal state
    for (int r = lane; r < n; r += 32) xi[r] = Gn[(size_t)(Tx - 1) * RECP + NN + r];
    for (int s = 0; s < STAGES - 1; ++s) issue(Tx - 2 - s);
    __syncwarp();
    if (Tx == 1) { for (int r = lane; r < n; r += 32) xn[r] = xi[r]; }
    else { for (int r = lane; r < D_; r += 32) xn[(size_t)(Tx - 1 + L_ - 1) * D_ + r] = xi[(n - D_) + r]; }
    for (int i = Tx - 2; i >= 0; --i) {
        issue(i - (STAGES - 1));
        asm volatile("cp.async.wait_group %0;\n" ::"n"(STAGES - 1));
        __syncwarp();
        const R* Gs = ring + (size_t)(i % STAGES) * RECP;
        R nv[(n + 31) / 32];
#pragma unroll
        for (int q = 0; q < (n + 31) / 32; ++q) {
            int r = lane + 32 * q;
            R a0 = 0, a1 = 0;
            if (r < n) {
                a0 = Gs[NN + r];
#pragma unroll 4
                for (int c = 0; c + 1 < n; c += 2) {
                    a0 = fma(Gs[c * n + r], xi[c], a0);
                    a1 = fma(Gs[(c + 1) * n + r], xi[c + 1], a1);
                }
                if (n & 1) a0 = fma(Gs[(n - 1) * n + r], xi[n - 1], a0);
            }
            nv[q] = a0 + a1;
        }
        __syncwarp();
#pragma unroll
        for (int q = 0; q < (n + 31) / 32; ++q) {
            int r = lane + 32 * q;
            if (r < n) {
                xi[r] = nv[q];
                if (i == 0) xn[r] = nv[q];                                  // frames 0..L-1 from xi_0
                else if (r >= n - D_) xn[(size_t)(i + L_ - 1) * D_ + (r - (n - D_))] = nv[q];
            }
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
template <typename R>
static void kalman_ws_layout(int N, int T, int d, int L, size_t off[5]) {
    const size_t n = (size_t)d * L, Tx = T - L + 1, fr = (size_t)N * Tx;
    const size_t rec = (size_t)d * (d + 1) / 2 + d, np2 = n * (n + 1) / 2;
    const size_t recs = ((n * n + n) * sizeof(R) + 15) / 16 * 16 / sizeof(R);
    off[0] = 0;
    off[1] = off[0] + align_up(fr * rec * sizeof(R), 256);     // info
    off[2] = off[1] + align_up(fr * n * sizeof(R), 256);       // stash_m
    off[3] = off[2] + align_up(fr * np2 * sizeof(R), 256);     // stash_S
    off[4] = off[3] + align_up(fr * recs * sizeof(R), 256);    // GH
}

template <typename R, int D_, int L_>
static int kalman_launch(const R* Y, const int* mask, const R* v, const R* h, const R* s, const int* z,
                         const R* Ct, const R* sigmasq, const R* Ab, const R* Q, double jitter,
                         const R* w_tape, uint64_t seed, int N, int T, int k, int Dk, R* x, void* ws,
                         cudaStream_t st) {
    constexpr int n = D_ * L_;
    const int Tx = T - L_ + 1;
    size_t off[5];
    kalman_ws_layout<R>(N, T, D_, L_, off);
    char* base = reinterpret_cast<char*>(ws);
    R* info = reinterpret_cast<R*>(base + off[0]);
    R* stash_m = reinterpret_cast<R*>(base + off[1]);
    R* stash_S = reinterpret_cast<R*>(base + off[2]);
    R* GH = reinterpret_cast<R*>(base + off[3]);
    const long long frames = (long long)N * Tx;
    {
        size_t smem = ((size_t)k * Dk * (D_ + 1) + k) * sizeof(R);
        int blocks = (int)((frames + 127) / 128);
        KPMS_LAUNCH("kalman_obs_info", st);
        if (Dk == 2)
            obs_info_kernel<R, D_, 2><<<blocks, 128, smem, st>>>(Y, mask, v, h, s, sigmasq, Ct, N, T, k, L_, info);
        else
            obs_info_kernel<R, D_, 3><<<blocks, 128, smem, st>>>(Y, mask, v, h, s, sigmasq, Ct, N, T, k, L_, info);
        int rc = check_launch("kalman obs_info");
        if (rc) return rc;
    }
    {
        auto kern = kalman_forward_kernel<R, D_, L_>;
        size_t smem = FwdSmem<R, D_, L_>::bytes;
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        { KPMS_LAUNCH("kalman_forward", st); kern<<<N, 256, smem, st>>>(info, mask, z, Ab, Q, (R)jitter, T, stash_m, stash_S); }
        int rc = check_launch("kalman forward");
        if (rc) return rc;
    }
    {
        constexpr int WARPS = 4;
        auto kern = kalman_backprep_kernel<R, D_, L_, WARPS>;
        size_t smem = PrepSmem<R, D_, L_>::per_warp * WARPS * sizeof(R);
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        int blocks = (int)((frames + WARPS - 1) / WARPS);
        { KPMS_LAUNCH("kalman_backprep", st); kern<<<blocks, 32 * WARPS, smem, st>>>(stash_m, stash_S, mask, z, Ab, Q, (R)jitter, w_tape, seed, N, T, GH); }
        int rc = check_launch("kalman backprep");
        if (rc) return rc;
    }
    {
        constexpr int STAGES = 4;
        auto kern = kalman_affine_kernel<R, D_, L_, STAGES>;
        size_t smem = ((size_t)STAGES * PrepSmem<R, D_, L_>::RECS + n) * sizeof(R);
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        { KPMS_LAUNCH("kalman_affine", st); kern<<<N, 32, smem, st>>>(GH, T, x); }
        int rc = check_launch("kalman affine");
        if (rc) return rc;
    }
    return 0;
}

template <typename R>
static int kalman_impl(const void* Y, const int* mask, const void* v, const void* h, const void* s, const int* z,
                       const void* Ct, const void* sigmasq, const void* Ab, const void* Q, double jitter,
                       const void* w_tape, uint64_t seed, int N, int T, int k, int Dk, int d, int L, void* x,
                       void* ws, cudaStream_t st) {
    if (Dk != 2 && Dk != 3) return set_error(-3, "kalman_sample: keypoint dimension must be 2 or 3, got %d", Dk);
    if (T < L) return set_error(-3, "kalman_sample: T (%d) < nlags (%d)", T, L);
#define X(DD, LL)                                                                                            \
    if (d == DD && L == LL)                                                                                  \
        return kalman_launch<R, DD, LL>((const R*)Y, mask, (const R*)v, (const R*)h, (const R*)s, z,         \
                                        (const R*)Ct, (const R*)sigmasq, (const R*)Ab, (const R*)Q, jitter,  \
                                        (const R*)w_tape, seed, N, T, k, Dk, (R*)x, ws, st);
    KPMS_FOR_EACH_DL(X)
#undef X
    return set_error(-3, "kalman_sample: unsupported (latent_dim, nlags) = (%d, %d)", d, L);
}

}  // namespace kpms

using namespace kpms;

extern "C" {

size_t kpms_kalman_workspace_bytes(int dtype, int N, int T, int d, int L) {
    size_t off[5];
    if (dtype == 0) kalman_ws_layout<float>(N, T, d, L, off);
    else kalman_ws_layout<double>(N, T, d, L, off);
    return off[4];
}

int kpms_kalman_sample(int dtype, const void* Y, const int* mask, const void* v, const void* h, const void* s,
                       const int* z, const void* Ct, const void* sigmasq, const void* Ab, const void* Q,
                       double jitter, const void* w_tape, uint64_t seed, int N, int T, int k, int Dk, int d,
                       int L, void* x, void* ws, void* stream) {
    return KPMS_DISPATCH_DTYPE(dtype, kalman_impl, Y, mask, v, h, s, z, Ct, sigmasq, Ab, Q, jitter, w_tape,
                               seed, N, T, k, Dk, d, L, x, ws, (cudaStream_t)stream);
}

}  // extern "C"
