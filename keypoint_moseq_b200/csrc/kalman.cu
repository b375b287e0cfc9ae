// Continuous-state resampling: Kalman forward filter + backward sampling over the
// (latent_dim x nlags) augmented AR state.  Replaces
// jax_moseq.models.keypoint_slds.resample_continuous_stateseqs -> utils.kalman.kalman_sample
// (reached from keypoint_moseq/fitting.py:25; flags at fitting.py:47-60, :260-261).
//
// The FFBS is split so that only the covariance recursion is serial:
//   K1a obs_info      (parallel over frames)  J_t = C~' R_t^-1 C~ -> chol, r_t = C~' R_t^-1 (y~_t - d~)
//   K1b forward       (serial per (chain, time chunk)) covariance-form filter, stashes (m_t, P_t): one warp with a
//                     covariance row per lane for n <= 32, a two-warp team for n <= 64 in float32
//                     (kalman_rows_wide.cuh), the 256-thread shared-memory kernel otherwise
//   K1c backprep      (parallel over frames)  G_t = P_t A' P'^-1, chol(Sigma_t), h_t = m_t - G_t(A m_t + b) + L_t w_t,
//                     in the two-stage form of kalman_split.cuh (nlags = 1: kalman_rows2.cuh)
//   K1d affine        (serial, warp per (chain, chunk)) xi_t = G_t xi_{t+1} + h_t
// which draws exactly the sample mu_t + chol(Sigma_t) w_t of the sequential sampler.
//
// Frame index i = t - (L-1), i in [0, Tx), Tx = T - L + 1.  z[i] governs the transition i -> i+1.
// Layouts: info (N,Tx,REC) REC = d(d+1)/2 + d;  stash_m (N,Tx,n);  stash_S (N,Tx,n(n+1)/2) packed lower;
//          GH (N,Tx,RECS): [GT (n*n) with GT[c][r] = G[r][c] | h (n) | pad], RECS*sizeof(R) % 16 == 0.
#include <algorithm>
#include <type_traits>
#include <cstdlib>
#include <string>
#ifndef KPMS_DL_GROUP
#define KPMS_DL_GROUP 0
#endif
#include "common.cuh"
#include "../../include/kpms_b200.h"

namespace kpms {

__device__ __forceinline__ void tri_unpack(int q, int& i, int& j) {
    i = (int)((sqrtf(8.0f * (float)q + 1.0f) - 1.0f) * 0.5f);
    while ((i + 1) * (i + 2) / 2 <= q) ++i;
    while (i * (i + 1) / 2 > q) --i;
    j = q - i * (i + 1) / 2;
}

// filter -> backward-pass records, padded to 16 bytes so they can be moved with 16-byte cp.async
__host__ __device__ constexpr int stash_m_stride(int n) { return (n + 3) / 4 * 4; }
__host__ __device__ constexpr int stash_S_stride(int n) { return (n * (n + 1) / 2 + 3) / 4 * 4; }
// offset of column c in a lower triangle packed by columns (column c holds rows c..n-1)
__host__ __device__ constexpr int col_start(int n, int c) { return c * n - c * (c - 1) / 2; }
// packing of the filter's covariance record: by columns from the row-per-lane filters, by rows from the shared-memory one
template <typename R, int n>
constexpr bool stash_rowpack() { return n > 32 && !(sizeof(R) == 4 && n <= 64); }
// per-frame observation record: [Lj packed lower (d(d+1)/2) | r (d) | Lj^-1 r (d) | pad]
__host__ __device__ constexpr int info_stride(int d) { return (d * (d + 1) / 2 + 2 * d + 3) / 4 * 4; }

// ---------------------------------------------------------------------------
// K1a: per-frame observation information
// ---------------------------------------------------------------------------
// One thread per frame; the records (REC numbers per frame) are staged in shared memory with an odd stride and leave
// the CTA as one contiguous, coalesced block: a thread writing its own 300-byte record word by word costs one
// 32-byte sector transaction per word (measured: 0.38 ms at C2 for 244 MB of records).
template <int D_>
struct ObsInfoCfg {
    static constexpr int REC = info_stride(D_), RS = REC | 1;          // padded record stride (odd: conflict-free)
    template <typename R>
    static constexpr int threads() { return (size_t)128 * RS * sizeof(R) <= 100 * 1024 ? 128 : 64; }
};

template <typename R, int D_, int DK, int TPB>
__global__ void __launch_bounds__(TPB)
obs_info_kernel(const R* __restrict__ Y, const int* __restrict__ mask, const R* __restrict__ v,
                const R* __restrict__ h, const R* __restrict__ s, const R* __restrict__ sigmasq,
                const R* __restrict__ Ct, int N, int T, int k, int L, R* __restrict__ info) {
    constexpr int NP = D_ * (D_ + 1) / 2, REC = info_stride(D_), RS = ObsInfoCfg<D_>::RS;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    R* recs = reinterpret_cast<R*>(smem_raw);          // TPB x RS
    R* Cs = recs + (size_t)TPB * RS;                   // (k*DK) x (D_+1)
    R* sg = Cs + (size_t)k * DK * (D_ + 1);            // k
    for (int i = threadIdx.x; i < k * DK * (D_ + 1); i += TPB) Cs[i] = Ct[i];
    for (int i = threadIdx.x; i < k; i += TPB) sg[i] = sigmasq[i];
    __syncthreads();
    const int Tx = T - L + 1;
    const long long total = (long long)N * Tx, base = (long long)blockIdx.x * TPB;
    const long long g = base + threadIdx.x;
    R* out = recs + (size_t)threadIdx.x * RS;
    bool on = false;
    size_t ft = 0;
    if (g < total) {
        const int nn = (int)(g / Tx), i = (int)(g % Tx);
        ft = (size_t)nn * T + (i + L - 1);
        on = mask[ft] != 0;
    }
    if (!on) {
#pragma unroll 1
        for (int q = 0; q < REC; ++q) out[q] = (R)0;
    } else {
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
    R invp[D_];
#pragma unroll
    for (int c = 0; c < D_; ++c) {
        R sdiag = J[c * (c + 1) / 2 + c];
#pragma unroll
        for (int p = 0; p < c; ++p) sdiag -= J[c * (c + 1) / 2 + p] * J[c * (c + 1) / 2 + p];
        const bool ok = sdiag > (R)0;
        const R inv = ok ? rsqrt_r<R>(sdiag) : (R)0;
        invp[c] = inv;
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
    // whitened pseudo-observation y~ = Lj^-1 r (unit noise, observation matrix Lj'); zero on zeroed pivots
#pragma unroll
    for (int c = 0; c < D_; ++c) {
        R val = r[c];
#pragma unroll
        for (int p = 0; p < c; ++p) val -= J[c * (c + 1) / 2 + p] * r[p];
        r[c] = val * invp[c];
        out[NP + D_ + c] = r[c];
    }
#pragma unroll
    for (int q = NP + 2 * D_; q < REC; ++q) out[q] = (R)0;      // padding of the record
    }
    __syncthreads();
    // the CTA's records are contiguous in global memory
    const int count = (int)min((long long)TPB, total - base) * REC;
    R* dst = info + (size_t)base * REC;
    for (int w = threadIdx.x; w < count; w += TPB) dst[w] = recs[(w / REC) * RS + (w % REC)];
}

// ---------------------------------------------------------------------------
// K1b: serial covariance-form filter, one CTA (256 threads) per chain, 4 barriers per step.
//
// With P the predicted covariance, Lj = chol(J_t) and B = I + Lj' P_nn Lj (eigenvalues >= 1):
//   U = P[:,new] Lj,  V = U chol(B)^-T,  P+ = P - V V',  m+ = m + P[:,new] c - V (V_new' c),  c = r - J m_new
// and the next prediction is assembled directly from products of the PREDICTED covariance,
//   A P+ A' = A P A' - (A V)(A V)',   A P+ = A P - (A V) V',
// so the dense products A P, A P A', A U do not wait for the serial d x d factorisation: they run on
// the other warps while warps 0/1 factor B redundantly in registers (no shuffles, no barriers).
//   phase 1: U, Lj' m_new, A P (two half-range register tiles), A m + b
//   phase 2: A P (sum), B, c, A U
//   phase 3: warps 0-1: chol(B), V / A V rows, V_new' c   |   warps 2-7: A P A', P[:,new] c
//   phase 4: P+ -> stash, next predicted covariance and mean
// ---------------------------------------------------------------------------
template <typename R, int D_, int L_>
struct FwdSmem {
    static constexpr int n = D_ * L_, LD = n | 1, NP = D_ * (D_ + 1) / 2, REC = info_stride(D_),
                         NP2 = n * (n + 1) / 2, VEC = 16 / (int)sizeof(R),
                         DP = (D_ + VEC - 1) / VEC * VEC, AP = (n + 1 + VEC - 1) / VEC * VEC;
    static constexpr size_t elems = 2 * n * LD + 2 * n + 2 * n * DP + 2 * D_ * DP + 3 * D_ * LD + 2 * D_ * D_ +
                                    3 * D_ + 2 * n + 2 * REC + 2 * D_ * AP + 2 * D_ * D_;
    static constexpr size_t bytes = (elems + 8) * sizeof(R) + NP2 * sizeof(unsigned short) + 16;
};

template <typename R, int D_, int L_>
__global__ void __launch_bounds__(256, 3)
kalman_forward_kernel(const R* __restrict__ info, const int* __restrict__ mask, const int* __restrict__ z,
                      const R* __restrict__ Ab, const R* __restrict__ Q, R jitter, int T,
                      R* __restrict__ stash_m, R* __restrict__ stash_S, int C, int W,
                      const int* __restrict__ vlen, const int* __restrict__ dirty, R* __restrict__ bnd_warm,
                      R* __restrict__ bnd_end) {
    typedef FwdSmem<R, D_, L_> SM;
    typedef typename Vec16<R>::type VecT;
    constexpr int n = SM::n, LD = SM::LD, NP = SM::NP, REC = SM::REC, NP2 = SM::NP2, VEC = SM::VEC,
                  DP = SM::DP, AP = SM::AP;
    constexpr int NT = 256;
    constexpr int NO = n - D_;                       // first index of the newest block
    constexpr int NA = D_ * (n + 1);
    constexpr int NPRE = (REC + NA + D_ * D_ + NT - 1) / NT;
    constexpr int GA = (D_ % 5 == 0) ? 5 : ((D_ % 4 == 0) ? 4 : ((D_ % 2 == 0) ? 2 : 1));   // A-row tile
    constexpr int NG = D_ / GA;
    constexpr int ES = ((n / 2) + VEC - 1) / VEC * VEC;                                      // split of the e range
    constexpr int DV = DP / VEC;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    R* base = reinterpret_cast<R*>(smem_raw);
    R* Az = base;                                    // 2 x D_ x AP   (16-byte aligned rows: [A | b | 0])
    R* U = Az + 2 * D_ * AP;                         // n x DP
    R* V = U + n * DP;                               // n x DP
    R* AU = V + n * DP;                              // D_ x DP
    R* AV = AU + D_ * DP;                            // D_ x DP
    R* Pb = AV + D_ * DP;                            // 2 x n x LD  (predicted covariance, ping-pong)
    R* TPh = Pb + 2 * n * LD;                        // 2 x D_ x LD (half-range partial A P)
    R* TP = TPh + 2 * D_ * LD;                       // D_ x LD
    R* APA = TP + D_ * LD;                           // D_ x D_
    R* Bm = APA + D_ * D_;                           // D_ x D_
    R* cvec = Bm + D_ * D_;
    R* tmp1 = cvec + D_;
    R* tvec = tmp1 + D_;
    R* pc = tvec + D_;                               // n
    R* mp = pc + n;                                  // n
    R* mb = mp + n;                                  // 2 x n (predicted mean, ping-pong)
    R* inf = mb + 2 * n;                             // 2 x REC
    R* Qz = inf + 2 * REC;                           // 2 x D_*D_
    unsigned short* ij = reinterpret_cast<unsigned short*>(Qz + 2 * D_ * D_ + 4);
    const int nn = blockIdx.x, tid = threadIdx.x, ck = blockIdx.y;
    const int Tx = T - L_ + 1;
    if (dirty && dirty[nn] == 0) return;             // sequential re-run of flagged chains only
    const ChunkRange cr = chunk_range(vlen ? vlen[nn] : Tx, Tx, C, W, ck);
    if (cr.empty) return;
    constexpr int BREC = n + n * n;                  // boundary record: mean | covariance
    const R* inf_g = info + (size_t)nn * Tx * REC;
    const int* mk = mask + (size_t)nn * T + (L_ - 1);
    const int* zz = z + (size_t)nn * (Tx - 1);
    constexpr int SMS = stash_m_stride(n), SSS = stash_S_stride(n);
    R* sm_g = stash_m + (size_t)nn * Tx * SMS;
    R* sS_g = stash_S + (size_t)nn * Tx * SSS;
    const R eps = (R)KPMS_EPS_SHIFT + jitter;

    for (int q = tid; q < NP2; q += NT) { int i, j; tri_unpack(q, i, j); ij[q] = (unsigned short)((i << 8) | j); }
    for (int w = tid; w < n * LD; w += NT) { int i = w / LD, j = w % LD; Pb[w] = (i == j) ? (R)KPMS_X_PRIOR_VAR : (R)0; }
    for (int w = tid; w < n; w += NT) mb[w] = (R)0;
    for (int w = tid; w < 2 * D_ * AP; w += NT) Az[w] = (R)0;
    for (int w = tid; w < 2 * n * DP; w += NT) U[w] = (R)0;          // U and V (contiguous), padding columns stay 0
    for (int w = tid; w < 2 * D_ * DP; w += NT) AU[w] = (R)0;        // AU and AV
    __syncthreads();

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
            else if (w < REC + NA) { int e = w - REC; Az[b * D_ * AP + (e / (n + 1)) * AP + (e % (n + 1))] = pre[q]; }
            else if (w < REC + NA + D_ * D_) Qz[b * D_ * D_ + (w - REC - NA)] = pre[q];
        }
    };
    R pre[NPRE];
    const int i0 = cr.start, i1 = cr.end;            // steps [i0, i1); outputs for i >= cr.begin
    int z_next = (i0 < Tx - 1) ? zz[i0] : -1;        // z for step i+1 while running step i
    stage_load(i0, z_next, pre);
    stage_store(i0 & 1, pre);
    z_next = (i0 + 1 < Tx - 1) ? zz[i0 + 1] : -1;
    int mk_cur = mk[i0];
    int pb = 0;                                      // which Pb / mb buffer holds the current prediction
    __syncthreads();

    for (int i = i0; i < i1; ++i) {
        const int b = i & 1;
        const bool last = (i == Tx - 1);
        const bool keep = (i >= cr.begin);           // warm-up steps store nothing
        const int mk_next = (i + 1 < Tx) ? mk[i + 1] : 0;
        if (i + 1 < i1) stage_load(i + 1, z_next, pre);
        const int z_next2 = (i + 2 < Tx - 1) ? zz[i + 2] : -1;
        const R* fi = inf + b * REC;
        const R* A = Az + b * D_ * AP;
        const R* Qs = Qz + b * D_ * D_;
        const R* P0 = Pb + pb * n * LD;
        R* Pn = Pb + (pb ^ 1) * n * LD;
        const R* m0 = mb + pb * n;
        R* mn = mb + (pb ^ 1) * n;
        if (ck > 0 && i == cr.begin) {               // the state this chunk arrived with
            R* bw = bnd_warm + ((size_t)nn * C + ck) * BREC;
            for (int w = tid; w < BREC; w += NT) bw[w] = (w < n) ? m0[w] : P0[((w - n) / n) * LD + (w - n) % n];
        }
        if (mk_cur != 0) {
            // ---------------- phase 1
            constexpr int TPI = NG * n * 2;
            const int n1 = (last ? 0 : TPI);
            for (int w = tid; w < n1 + n * D_ + D_ + (last ? 0 : n); w += NT) {
                if (w < n1) {                                   // register tile: GA rows of A x one column of P
                    const int hh = w / (NG * n), g = (w / n) % NG, j = w % n;
                    R acc[GA];
#pragma unroll
                    for (int q = 0; q < GA; ++q) acc[q] = (R)0;
                    const int e0 = hh ? ES : 0;
#pragma unroll
                    for (int ev = 0; ev < (hh ? (n - ES + VEC - 1) / VEC : ES / VEC); ++ev) {
                        R pv[VEC];
#pragma unroll
                        for (int c = 0; c < VEC; ++c) { const int e = e0 + ev * VEC + c; pv[c] = (e < n) ? P0[e * LD + j] : (R)0; }
#pragma unroll
                        for (int q = 0; q < GA; ++q) {
                            const VecT av = *reinterpret_cast<const VecT*>(A + (g * GA + q) * AP + e0 + ev * VEC);
                            const R* ae = reinterpret_cast<const R*>(&av);
#pragma unroll
                            for (int c = 0; c < VEC; ++c) acc[q] = fma(ae[c], pv[c], acc[q]);
                        }
                    }
#pragma unroll
                    for (int q = 0; q < GA; ++q) TPh[hh * D_ * LD + (g * GA + q) * LD + j] = acc[q];
                } else if (w < n1 + n * D_) {                   // U = P[:,new] Lj
                    const int u = w - n1, r = u / D_, c = u % D_;
                    R acc = 0;
#pragma unroll
                    for (int e = 0; e < D_; ++e)
                        if (e >= c) acc = fma(P0[r * LD + NO + e], fi[e * (e + 1) / 2 + c], acc);
                    U[r * DP + c] = acc;
                } else if (w < n1 + n * D_ + D_) {              // tmp1 = Lj' m_new
                    const int c = w - n1 - n * D_;
                    R acc = 0;
#pragma unroll
                    for (int e = 0; e < D_; ++e)
                        if (e >= c) acc = fma(fi[e * (e + 1) / 2 + c], m0[NO + e], acc);
                    tmp1[c] = acc;
                } else {                                        // mp = Aaug m + b
                    const int r = w - n1 - n * D_ - D_;
                    if (r < NO) mp[r] = m0[r + D_];
                    else {
                        const int a = r - NO;
                        R acc = A[a * AP + n];
#pragma unroll 6
                        for (int e = 0; e < n; ++e) acc = fma(A[a * AP + e], m0[e], acc);
                        mp[r] = acc;
                    }
                }
            }
            __syncthreads();
            // ---------------- phase 2
            {
                constexpr int AUI = D_ * DV;                    // A U tiles: one A row x VEC columns of U
                const int nau = last ? 0 : AUI;
                const int ntp = last ? 0 : D_ * n;
                for (int w = tid; w < nau + D_ * D_ + D_ + ntp; w += NT) {
                    if (w < nau) {
                        const int a = w / DV, cv = w % DV;
                        R acc[VEC];
#pragma unroll
                        for (int c = 0; c < VEC; ++c) acc[c] = (R)0;
#pragma unroll 6
                        for (int e = 0; e < n; ++e) {
                            const R ae = A[a * AP + e];
                            const VecT uv = *reinterpret_cast<const VecT*>(U + e * DP + cv * VEC);
                            const R* ue = reinterpret_cast<const R*>(&uv);
#pragma unroll
                            for (int c = 0; c < VEC; ++c) acc[c] = fma(ae, ue[c], acc[c]);
                        }
#pragma unroll
                        for (int c = 0; c < VEC; ++c) AU[a * DP + cv * VEC + c] = acc[c];
                    } else if (w < nau + D_ * D_) {             // B = I + Lj' U[new,:]
                        const int u = w - nau, a = u / D_, c = u % D_;
                        R acc = (a == c) ? (R)1 : (R)0;
#pragma unroll
                        for (int e = 0; e < D_; ++e)
                            if (e >= a) acc = fma(fi[e * (e + 1) / 2 + a], U[(NO + e) * DP + c], acc);
                        Bm[u] = acc;
                    } else if (w < nau + D_ * D_ + D_) {        // c = r - Lj tmp1
                        const int a = w - nau - D_ * D_;
                        R acc = fi[NP + a];
#pragma unroll
                        for (int c = 0; c < D_; ++c)
                            if (c <= a) acc = fma(-fi[a * (a + 1) / 2 + c], tmp1[c], acc);
                        cvec[a] = acc;
                    } else {                                    // A P = sum of the two half-range partials
                        const int u = w - nau - D_ * D_ - D_, a = u / n, j = u % n;
                        TP[a * LD + j] = TPh[a * LD + j] + TPh[D_ * LD + a * LD + j];
                    }
                }
            }
            __syncthreads();
            // ---------------- phase 3
            if (tid < 64) {
                // right-looking Cholesky of B in registers, redundantly in every lane of warps 0 and 1
                R Lb[NP];
#pragma unroll
                for (int a = 0; a < D_; ++a)
#pragma unroll
                    for (int c = 0; c <= a; ++c) Lb[a * (a + 1) / 2 + c] = Bm[a * D_ + c];
#pragma unroll
                for (int c = 0; c < D_; ++c) {
                    const R inv = rsqrt_r<R>(Lb[c * (c + 1) / 2 + c]);
                    Lb[c * (c + 1) / 2 + c] = inv;               // inverse diagonal
#pragma unroll
                    for (int a = c + 1; a < D_; ++a) Lb[a * (a + 1) / 2 + c] *= inv;
#pragma unroll
                    for (int a = c + 1; a < D_; ++a)
#pragma unroll
                        for (int bb = c + 1; bb <= a; ++bb)
                            Lb[a * (a + 1) / 2 + bb] = fma(-Lb[a * (a + 1) / 2 + c], Lb[bb * (bb + 1) / 2 + c], Lb[a * (a + 1) / 2 + bb]);
                }
                // rows of V = U Lb^-T (warp 0) and of A V = (A U) Lb^-T (warp 1)
                const int lane = tid & 31;
                const bool w1 = tid >= 32;
                const R* src = w1 ? AU : U;
                R* dst = w1 ? AV : V;
                const int rows = w1 ? (last ? 0 : D_) : n;
                for (int r = lane; r < rows; r += 32) {
                    R vr[D_];
#pragma unroll
                    for (int c = 0; c < D_; ++c) {
                        R val = src[r * DP + c];
#pragma unroll
                        for (int p2 = 0; p2 < c; ++p2) val = fma(-Lb[c * (c + 1) / 2 + p2], vr[p2], val);
                        vr[c] = val * Lb[c * (c + 1) / 2 + c];
                        dst[r * DP + c] = vr[c];
                    }
                }
                if (!w1) {
                    __syncwarp();
                    if (lane < D_) {                            // tvec = V_new' c
                        R acc = 0;
#pragma unroll
                        for (int r = 0; r < D_; ++r) acc = fma(V[(NO + r) * DP + lane], cvec[r], acc);
                        tvec[lane] = acc;
                    }
                }
            } else {
                const int t2 = tid - 64;
                const int napa = last ? 0 : NP;
                for (int w = t2; w < napa + n; w += NT - 64) {
                    if (w < napa) {                             // A P A' (lower)
                        const int a = ij[w] >> 8, c = ij[w] & 255;
                        R a0 = 0, a1 = 0;
#pragma unroll 5
                        for (int e = 0; e + 1 < n; e += 2) {
                            a0 = fma(TP[a * LD + e], A[c * AP + e], a0);
                            a1 = fma(TP[a * LD + e + 1], A[c * AP + e + 1], a1);
                        }
                        if (n & 1) a0 = fma(TP[a * LD + n - 1], A[c * AP + n - 1], a0);
                        APA[a * D_ + c] = a0 + a1;
                    } else {                                    // pc = P[:,new] c
                        const int r = w - napa;
                        R acc = 0;
#pragma unroll
                        for (int c = 0; c < D_; ++c) acc = fma(P0[r * LD + NO + c], cvec[c], acc);
                        pc[r] = acc;
                    }
                }
            }
            __syncthreads();
            // ---------------- phase 4
            {
                auto rowdot = [&](const R* x, const R* y) {
                    R acc = 0;
#pragma unroll
                    for (int cv = 0; cv < DV; ++cv) {
                        const VecT xa = *reinterpret_cast<const VecT*>(x + cv * VEC);
                        const VecT ya = *reinterpret_cast<const VecT*>(y + cv * VEC);
                        const R* xe = reinterpret_cast<const R*>(&xa);
                        const R* ye = reinterpret_cast<const R*>(&ya);
#pragma unroll
                        for (int c = 0; c < VEC; ++c) acc = fma(xe[c], ye[c], acc);
                    }
                    return acc;
                };
                const int nmn = last ? 0 : D_;                  // new-block mean rows (heaviest items first)
                const int nno = last ? 0 : D_ * NO;
                const int nnn = last ? 0 : NP;
                const int nmo = last ? 0 : NO;
                for (int w = tid; w < nmn + NP2 + nno + nnn + n + nmo; w += NT) {
                    int u = w;
                    if (u < nmn) {                              // m'[new] = mp + A pc - (A V) tvec
                        const int a = u;
                        R acc = mp[NO + a];
#pragma unroll 6
                        for (int e = 0; e < n; ++e) acc = fma(A[a * AP + e], pc[e], acc);
#pragma unroll
                        for (int c = 0; c < D_; ++c) acc = fma(-AV[a * DP + c], tvec[c], acc);
                        mn[NO + a] = acc;
                        continue;
                    }
                    u -= nmn;
                    if (u < NP2) {                              // P+ (stash) and the shifted old-old block
                        const int r = ij[u] >> 8, c = ij[u] & 255;
                        const R val = P0[r * LD + c] - rowdot(V + r * DP, V + c * DP);
                        if (keep) sS_g[(size_t)i * SSS + u] = val;
                        if (!last && c >= D_) {
                            const R v2 = val + ((r == c) ? eps : (R)0);
                            Pn[(r - D_) * LD + (c - D_)] = v2;
                            Pn[(c - D_) * LD + (r - D_)] = v2;
                        }
                        continue;
                    }
                    u -= NP2;
                    if (u < nno) {                              // new-old block
                        const int a = u / NO, j = u % NO;
                        const R val = TP[a * LD + j + D_] - rowdot(AV + a * DP, V + (j + D_) * DP);
                        Pn[(NO + a) * LD + j] = val;
                        Pn[j * LD + NO + a] = val;
                        continue;
                    }
                    u -= nno;
                    if (u < nnn) {                              // new-new block
                        const int a = ij[u] >> 8, c = ij[u] & 255;
                        const R val = APA[a * D_ + c] - rowdot(AV + a * DP, AV + c * DP) + Qs[a * D_ + c] +
                                      ((a == c) ? jitter : (R)0);
                        Pn[(NO + a) * LD + NO + c] = val;
                        Pn[(NO + c) * LD + NO + a] = val;
                        continue;
                    }
                    u -= nnn;
                    if (u < n) {                                // m+ (stash)
                        R acc = m0[u] + pc[u];
#pragma unroll
                        for (int c = 0; c < D_; ++c) acc = fma(-V[u * DP + c], tvec[c], acc);
                        if (keep) sm_g[(size_t)i * SMS + u] = acc;
                        continue;
                    }
                    u -= n;
                    {                                           // m'[old] = mp + delta[shifted]
                        R acc = mp[u] + pc[u + D_];
#pragma unroll
                        for (int c = 0; c < D_; ++c) acc = fma(-V[(u + D_) * DP + c], tvec[c], acc);
                        mn[u] = acc;
                    }
                }
            }
            if (!last) pb ^= 1;
        } else if (last) {
            for (int q = tid; q < NP2; q += NT) {
                int r = ij[q] >> 8, c = ij[q] & 255;
                sS_g[(size_t)i * SSS + q] = P0[r * LD + c];
            }
            for (int w = tid; w < n; w += NT) sm_g[(size_t)i * SMS + w] = m0[w];
        }
        if (i + 1 < i1) stage_store(b ^ 1, pre);
        z_next = z_next2;
        mk_cur = mk_next;
        __syncthreads();
    }
    if (i1 < Tx) {                                   // the state handed to the next chunk
        R* be = bnd_end + ((size_t)nn * C + ck + 1) * BREC;
        const R* P0 = Pb + pb * n * LD;
        const R* m0 = mb + pb * n;
        for (int w = tid; w < BREC; w += NT) be[w] = (w < n) ? m0[w] : P0[((w - n) / n) * LD + (w - n) % n];
    }
}

// ---------------------------------------------------------------------------
// K1b (n <= 32): covariance-form filter with one warp per (chain, time chunk) and one row of the
// predicted covariance per lane in registers (the matrix is symmetric, so the row is also the
// column).  Per step, with Lj = chol(J_t), y~ = Lj^-1 r_t (unit-noise pseudo-observation of Lj' x_new):
//   U = P[:,new] Lj                     own row, Lj broadcast from the prefetched record
//   B = I + Lj' U[new,:], nu = y~ - Lj' m_new      55 + 10 entries spread over the lanes, through smem
//   Lb = chol(B)                        redundantly in every lane's registers (no communication)
//   V = U Lb^-T, w = Lb^-1 nu,  m+ = m + V w,  P+ = P - V V'     V rows published, read as broadcasts
//   A P+ (lane j: column j),  m' = A m+ + b,  A P+ A' by L partial sums per lane group + shuffles,
//   shifted blocks by shuffles from lane r + d
// The transition parameters of the current state sit in shared memory and are reloaded only when
// z changes; the per-frame records arrive through a cp.async ring.
// stash_S here is the lower triangle packed by COLUMNS (column c holds rows c..n-1), which makes
// both this kernel's stores and the backward preparation's loads contiguous across lanes.
// ---------------------------------------------------------------------------
template <typename R, int D_, int L_>
struct FwdRowsSmem {
    static constexpr int n = D_ * L_, NP = D_ * (D_ + 1) / 2, NPP = (NP + 3) / 4 * 4, RECI = info_stride(D_);
    static constexpr int QO = (n + 1 + 3) / 4 * 4;                 // offset of the Q row inside an A row
    static constexpr int AS = QO + (D_ + 3) / 4 * 4;               // row: [A (n) | b | pad | Q row (d)]
    static constexpr int VS = (D_ + 3) / 4 * 4, APS = 36, STAGES = 4;
    static constexpr int BS = NPP + (D_ + 3) / 4 * 4;
    static constexpr size_t per_warp = STAGES * RECI + 2 * D_ * AS + 32 * VS + D_ * APS + BS + 32;
};

template <typename R, int D_, int L_, int WARPS>
__global__ void __launch_bounds__(32 * WARPS, (sizeof(R) == 4 ? 2 : 1))
kalman_forward_rows_kernel(const R* __restrict__ info, const int* __restrict__ mask, const int* __restrict__ z,
                           const R* __restrict__ Ab, const R* __restrict__ Q, R jitter, int N, int T,
                           R* __restrict__ stash_m, R* __restrict__ stash_S, int C, int W,
                           const int* __restrict__ vlen, const int* __restrict__ vend, const int* __restrict__ dirty,
                           R* __restrict__ bnd_warm, R* __restrict__ bnd_end) {
    typedef FwdRowsSmem<R, D_, L_> SM;
    typedef typename Vec16<R>::type VecT;
    constexpr int n = SM::n, NO = n - D_, NP = SM::NP, NPP = SM::NPP, RECI = SM::RECI, AS = SM::AS, QO = SM::QO,
                  VS = SM::VS, APS = SM::APS, STAGES = SM::STAGES;
    constexpr int VEC = 16 / (int)sizeof(R), NV = (n + VEC - 1) / VEC, DV = (D_ + VEC - 1) / VEC;
    constexpr int SMS = stash_m_stride(n), SSS = stash_S_stride(n), BREC = n + n * n;
    static_assert(n <= 32, "one covariance row per lane");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long task = (long long)blockIdx.x * WARPS + warp;
    if (task >= (long long)N * C) return;
    const int nn = (int)(task / C), ck = (int)(task % C);
    const int Tx = T - L_ + 1;
    if (dirty && dirty[nn] == 0) return;
    const ChunkRange cr = chunk_range(vlen ? vlen[nn] : Tx, Tx, C, W, ck);
    if (cr.empty) return;
    R* ring = reinterpret_cast<R*>(smem_raw) + (size_t)warp * SM::per_warp;   // STAGES x RECI
    R* Asb = ring + STAGES * RECI;                   // 2 x D_ x AS
    R* Vs = Asb + 2 * D_ * AS;                       // 32 x VS   (U rows, then V rows)
    R* APs = Vs + 32 * VS;                           // D_ x APS  (A P+)
    R* Bs = APs + D_ * APS;                          // [B lower packed | nu]
    R* ms = Bs + SM::BS;                             // 32
    const bool act = lane < n;
    const int row = act ? lane : n - 1;              // idle lanes shadow the last row
    const R* inf_g = info + (size_t)nn * Tx * RECI;
    const int* mk = mask + (size_t)nn * T + (L_ - 1);
    const int* zz = z + (size_t)nn * (Tx - 1);
    R* sm_g = stash_m + (size_t)nn * Tx * SMS;
    R* sS_g = stash_S + (size_t)nn * Tx * SSS;
    const R eps = (R)KPMS_EPS_SHIFT + jitter;
    // entries of B (lower, packed by rows) computed by this lane: NQ = ceil(d(d+1)/2 / 32) per lane
    constexpr int NQ = (NP + 31) / 32;
    int ba[NQ], bc[NQ];
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
        int a = 0, c = 0;
        const int idx = lane + 32 * q;
        if (idx < NP) tri_unpack(idx, a, c);
        ba[q] = a;
        bc[q] = c;
    }
    // frames from vend[nn] on are all masked: they carry (m, P) unchanged, so the walk stops there and the state is
    // stored once at the terminal frame (a short row's last chunk used to step through thousands of padded frames)
    const int i0 = cr.start, i1 = cr.end, i_stop = min(i1, vend[nn]);
    auto issue_info = [&](int i) {
        if (i < i_stop)
            for (int c = lane; c < RECI * (int)sizeof(R) / 16; c += 32)
                cp_async_16(reinterpret_cast<char*>(ring + (i % STAGES) * RECI) + 16 * c,
                            reinterpret_cast<const char*>(inf_g + (size_t)i * RECI) + 16 * c);
        asm volatile("cp.async.commit_group;\n" ::);
    };
    auto load_A = [&](int zi, int buf) {             // joins the next committed group
        R* dst = Asb + buf * D_ * AS;
        const R* A = Ab + (size_t)zi * D_ * (n + 1);
        for (int w = lane; w < D_ * (n + 1); w += 32) cp_async_elem(dst + (w / (n + 1)) * AS + (w % (n + 1)), A + w);
        const R* Qk = Q + (size_t)zi * D_ * D_;
        for (int w = lane; w < D_ * D_; w += 32) cp_async_elem(dst + (w / D_) * AS + QO + (w % D_), Qk + w);
    };
    R p[n], m = (R)0;
#pragma unroll
    for (int c = 0; c < n; ++c) p[c] = (c == row) ? (R)KPMS_X_PRIOR_VAR : (R)0;
    for (int w = lane; w < 2 * D_ * AS; w += 32) Asb[w] = (R)0;     // padding is multiplied by masked zeros
    __syncwarp();
    int cur = 0;
    int zc = (i0 < Tx - 1) ? zz[i0] : -1;
    if (zc >= 0) load_A(zc, 0);
    for (int s2 = 0; s2 < STAGES - 1; ++s2) issue_info(i0 + s2);
    int mk_cur = mk[i0];
    bool changed_prev = false;
    for (int i = i0; i < i_stop; ++i) {
        const bool last = (i == Tx - 1);
        const bool keep = (i >= cr.begin);
        const int mk_next = (i + 1 < Tx) ? mk[i + 1] : 0;
        const int z_next = (i + 1 < Tx - 1) ? zz[i + 1] : -1;
        const bool change = (z_next >= 0 && z_next != zc);
        if (change) load_A(z_next, cur ^ 1);
        issue_info(i + STAGES - 1);
        if (changed_prev) asm volatile("cp.async.wait_all;\n" ::);
        else asm volatile("cp.async.wait_group %0;\n" ::"n"(STAGES - 1));
        __syncwarp();
        if (ck > 0 && i == cr.begin && act) {        // the state this chunk arrived with
            R* bw = bnd_warm + ((size_t)nn * C + ck) * BREC;
            bw[lane] = m;
#pragma unroll
            for (int c = 0; c < n; ++c) bw[n + lane * n + c] = p[c];
        }
        const R* fi = ring + (i % STAGES) * RECI;
        const R* A = Asb + cur * D_ * AS;
        if (mk_cur != 0) {
            // ---- U = P[:,new] Lj (own row); publish U rows and the mean
            R u[D_];
            {
                R lj[NPP];
#pragma unroll
                for (int cv = 0; cv < NPP / VEC; ++cv) {
                    const VecT lv = *reinterpret_cast<const VecT*>(fi + cv * VEC);
                    const R* le = reinterpret_cast<const R*>(&lv);
#pragma unroll
                    for (int q = 0; q < VEC; ++q) lj[cv * VEC + q] = le[q];
                }
#pragma unroll
                for (int c = 0; c < D_; ++c) {
                    R acc = 0;
#pragma unroll
                    for (int e = c; e < D_; ++e) acc = fma(p[NO + e], lj[e * (e + 1) / 2 + c], acc);
                    u[c] = acc;
                }
            }
            ms[lane] = m;
#pragma unroll
            for (int cv = 0; cv < DV; ++cv) {
                VecT ov;
                R* oe = reinterpret_cast<R*>(&ov);
#pragma unroll
                for (int q = 0; q < VEC; ++q) oe[q] = (cv * VEC + q < D_) ? u[cv * VEC + q] : (R)0;
                *reinterpret_cast<VecT*>(Vs + lane * VS + cv * VEC) = ov;
            }
            __syncwarp();
            // ---- B = I + Lj' U[new,:] (lower) and nu = y~ - Lj' m_new, spread over the lanes
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
                const int a = ba[q], c = bc[q];
                R acc = (a == c) ? (R)1 : (R)0;
#pragma unroll
                for (int e = 0; e < D_; ++e)
                    if (e >= a) acc = fma(fi[e * (e + 1) / 2 + a], Vs[(NO + e) * VS + c], acc);
                if (lane + 32 * q < NP) Bs[lane + 32 * q] = acc;
            }
            if (lane < D_) {
                R acc = fi[NP + D_ + lane];
#pragma unroll
                for (int e = 0; e < D_; ++e)
                    if (e >= lane) acc = fma(-fi[e * (e + 1) / 2 + lane], ms[NO + e], acc);
                Bs[NPP + lane] = acc;
            }
            __syncwarp();
            // ---- Lb = chol(B) in registers (inverse pivots on the diagonal); V row, w, m+
            R v[D_];
            {
                R Lb[NPP], nu[DV * VEC];
#pragma unroll
                for (int cv = 0; cv < NPP / VEC; ++cv) {
                    const VecT lv = *reinterpret_cast<const VecT*>(Bs + cv * VEC);
                    const R* le = reinterpret_cast<const R*>(&lv);
#pragma unroll
                    for (int q = 0; q < VEC; ++q) Lb[cv * VEC + q] = le[q];
                }
#pragma unroll
                for (int cv = 0; cv < DV; ++cv) {
                    const VecT lv = *reinterpret_cast<const VecT*>(Bs + NPP + cv * VEC);
                    const R* le = reinterpret_cast<const R*>(&lv);
#pragma unroll
                    for (int q = 0; q < VEC; ++q) nu[cv * VEC + q] = le[q];
                }
#pragma unroll
                for (int c = 0; c < D_; ++c) {
                    const R inv = rsqrt_fast<R>(Lb[c * (c + 1) / 2 + c]);
                    Lb[c * (c + 1) / 2 + c] = inv;
#pragma unroll
                    for (int a = c + 1; a < D_; ++a) Lb[a * (a + 1) / 2 + c] *= inv;
#pragma unroll
                    for (int a = c + 1; a < D_; ++a)
#pragma unroll
                        for (int bb = c + 1; bb <= a; ++bb)
                            Lb[a * (a + 1) / 2 + bb] = fma(-Lb[a * (a + 1) / 2 + c], Lb[bb * (bb + 1) / 2 + c], Lb[a * (a + 1) / 2 + bb]);
                }
                R dm = 0;
#pragma unroll
                for (int c = 0; c < D_; ++c) {
                    R val = u[c], wv = nu[c];
#pragma unroll
                    for (int p2 = 0; p2 < c; ++p2) {
                        val = fma(-Lb[c * (c + 1) / 2 + p2], v[p2], val);
                        wv = fma(-Lb[c * (c + 1) / 2 + p2], nu[p2], wv);
                    }
                    v[c] = val * Lb[c * (c + 1) / 2 + c];
                    nu[c] = wv * Lb[c * (c + 1) / 2 + c];
                    dm = fma(v[c], nu[c], dm);
                }
                m += dm;
            }
#pragma unroll
            for (int cv = 0; cv < DV; ++cv) {
                VecT ov;
                R* oe = reinterpret_cast<R*>(&ov);
#pragma unroll
                for (int q = 0; q < VEC; ++q) oe[q] = (cv * VEC + q < D_) ? v[cv * VEC + q] : (R)0;
                *reinterpret_cast<VecT*>(Vs + lane * VS + cv * VEC) = ov;
            }
            __syncwarp();
            // ---- P+ = P - V V' (own row); packed FMAs over pairs of the d contraction terms
            {
                R nv[DV * VEC];
#pragma unroll
                for (int q = 0; q < DV * VEC; ++q) nv[q] = (q < D_) ? -v[q] : (R)0;
#pragma unroll
                for (int c = 0; c < n; ++c) {
                    R acc0 = p[c], acc1 = 0;
#pragma unroll
                    for (int cv = 0; cv < DV; ++cv) {
                        const VecT vv = *reinterpret_cast<const VecT*>(Vs + c * VS + cv * VEC);
                        const R* ve = reinterpret_cast<const R*>(&vv);
#pragma unroll
                        for (int q = 0; q < VEC; q += 2) {
                            if (cv * VEC + q + 1 < D_) fma2<R>(acc0, acc1, nv[cv * VEC + q], nv[cv * VEC + q + 1], ve[q], ve[q + 1]);
                            else if (cv * VEC + q < D_) acc0 = fma(nv[cv * VEC + q], ve[q], acc0);
                        }
                    }
                    p[c] = acc0 + acc1;
                }
            }
            if (keep && act) {
                sm_g[(size_t)i * SMS + lane] = m;
                R* so = sS_g + (size_t)i * SSS + lane;
#pragma unroll
                for (int c = 0; c < n; ++c)
                    if (c <= lane) so[col_start(n, c) - c] = p[c];
            }
            if (!last) {
                // ---- A P+ (column `row`), published by rows of A
                R ap[D_];
#pragma unroll
                for (int a = 0; a < D_; ++a) {
                    R acc0 = 0, acc1 = 0;
#pragma unroll
                    for (int cv = 0; cv < NV; ++cv) {
                        const VecT av = *reinterpret_cast<const VecT*>(A + a * AS + cv * VEC);
                        const R* ae = reinterpret_cast<const R*>(&av);
#pragma unroll
                        for (int q = 0; q < VEC; q += 2) {
                            const int e = cv * VEC + q;
                            if (e + 1 < n) fma2<R>(acc0, acc1, ae[q], ae[q + 1], p[e], p[e + 1]);
                            else if (e < n) acc0 = fma(ae[q], p[e], acc0);
                        }
                    }
                    ap[a] = acc0 + acc1;
                    APs[a * APS + lane] = ap[a];
                }
                ms[lane] = m;
                __syncwarp();
                // ---- next mean: shifted blocks by shuffle, newest block = A m+ + b
                const int arow = (row >= NO) ? row - NO : 0;
                const int grp = row / D_;                         // lane group = block of the augmented state
                {
                    const R* Ar = A + arow * AS;
                    R acc0 = Ar[n], acc1 = 0;
#pragma unroll
                    for (int cv = 0; cv < NV; ++cv) {
                        const VecT av = *reinterpret_cast<const VecT*>(Ar + cv * VEC);
                        const VecT mv = *reinterpret_cast<const VecT*>(ms + cv * VEC);
                        const R* ae = reinterpret_cast<const R*>(&av);
                        const R* me = reinterpret_cast<const R*>(&mv);
#pragma unroll
                        for (int q = 0; q < VEC; ++q) {
                            const int e = cv * VEC + q;
                            if (e < n) { if (e & 1) acc1 = fma(ae[q], me[q], acc1); else acc0 = fma(ae[q], me[q], acc0); }
                        }
                    }
                    const R mshift = __shfl_down_sync(0xffffffffu, m, D_);
                    m = (row < NO) ? mshift : (acc0 + acc1);
                }
                // ---- A P+ A': lane (grp, a' = row - grp d) sums its block of the contraction for every
                //      column a; the L partial sums meet in the last lane group by shuffles
                R apa[D_];
                {
                    constexpr int WN = (D_ + 2 * VEC - 2) / VEC;         // aligned window covering any block
                    const int e0 = grp * D_;
                    const int w0 = e0 / VEC * VEC;
                    const R* aprow = APs + (row - e0) * APS + w0;
                    R aw[WN * VEC];
#pragma unroll
                    for (int cv = 0; cv < WN; ++cv) {
                        const VecT lv = *reinterpret_cast<const VecT*>(aprow + cv * VEC);
                        const R* le = reinterpret_cast<const R*>(&lv);
#pragma unroll
                        for (int q = 0; q < VEC; ++q) {
                            const int e = w0 + cv * VEC + q;
                            aw[cv * VEC + q] = (e >= e0 && e < e0 + D_) ? le[q] : (R)0;
                        }
                    }
#pragma unroll
                    for (int a = 0; a < D_; ++a) {
                        R acc = 0, acc1 = 0;
#pragma unroll
                        for (int cv = 0; cv < WN; ++cv) {
                            const VecT av = *reinterpret_cast<const VecT*>(A + a * AS + w0 + cv * VEC);
                            const R* ae = reinterpret_cast<const R*>(&av);
#pragma unroll
                            for (int q = 0; q < VEC; q += 2) fma2<R>(acc, acc1, aw[cv * VEC + q], aw[cv * VEC + q + 1], ae[q], ae[q + 1]);
                        }
                        acc += acc1;
                        R tot = acc;
#pragma unroll
                        for (int gq = 1; gq < L_; ++gq) tot += __shfl_up_sync(0xffffffffu, acc, D_ * gq);
                        apa[a] = tot;                                   // complete in the last group (rows >= NO)
                    }
                    // Keep the covariance symmetric to the last bit: every other block of P' is mirrored
                    // by construction, this one is computed twice in different orders.  An antisymmetric
                    // rounding residue is not contracted by the measurement update and grows with |A| > 1.
                    if (act && row >= NO) {                  // idle lanes hold partial sums of the wrong lanes
#pragma unroll
                        for (int a = 0; a < D_; ++a) Vs[arow * VS + a] = apa[a];
                    }
                    __syncwarp();
#pragma unroll
                    for (int a = 0; a < D_; ++a) apa[a] = (R)0.5 * (apa[a] + Vs[a * VS + arow]);
                }
                // ---- next predicted covariance: rows < NO are shifted rows of P+ / columns of A P+
                //      (from lane r + d), rows >= NO are A P+ and A P+ A' + Q
                {
                    const R* aprow = APs + arow * APS;
                    R nv[NV * VEC];
#pragma unroll
                    for (int cv = 0; cv < NV; ++cv) {
                        const VecT lv = *reinterpret_cast<const VecT*>(aprow + cv * VEC);
                        const R* le = reinterpret_cast<const R*>(&lv);
#pragma unroll
                        for (int q = 0; q < VEC; ++q) nv[cv * VEC + q] = le[q];
                    }
                    R pn[n];
#pragma unroll
                    for (int c = 0; c < NO; ++c) {
                        const R sh = __shfl_down_sync(0xffffffffu, p[c + D_], D_);
                        pn[c] = (row < NO) ? sh + ((row == c) ? eps : (R)0) : nv[c + D_];
                    }
                    const R* qrow = A + arow * AS + QO;
#pragma unroll
                    for (int a = 0; a < D_; ++a) {
                        const R sh = __shfl_down_sync(0xffffffffu, ap[a], D_);
                        const R nw = apa[a] + qrow[a] + ((arow == a) ? jitter : (R)0);
                        pn[NO + a] = (row < NO) ? sh : nw;
                    }
#pragma unroll
                    for (int c = 0; c < n; ++c) p[c] = pn[c];
                }
            }
        } else if (last && keep && act) {
            sm_g[(size_t)i * SMS + lane] = m;
            R* so = sS_g + (size_t)i * SSS + lane;
#pragma unroll
            for (int c = 0; c < n; ++c)
                if (c <= lane) so[col_start(n, c) - c] = p[c];
        }
        if (change) { cur ^= 1; zc = z_next; }
        changed_prev = change;
        mk_cur = mk_next;
        __syncwarp();
    }
    asm volatile("cp.async.wait_all;\n" ::);
    if (i1 == Tx && i_stop < Tx && act) {            // masked terminal frame: the carried state (Tx - 1 >= cr.begin)
        sm_g[(size_t)(Tx - 1) * SMS + lane] = m;
        R* so = sS_g + (size_t)(Tx - 1) * SSS + lane;
#pragma unroll
        for (int c = 0; c < n; ++c)
            if (c <= lane) so[col_start(n, c) - c] = p[c];
    }
    if (i1 < Tx && act) {                            // the state handed to the next chunk
        R* be = bnd_end + ((size_t)nn * C + ck + 1) * BREC;
        be[lane] = m;
#pragma unroll
        for (int c = 0; c < n; ++c) be[n + lane * n + c] = p[c];
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
// K1c: backward preparation (kalman_split.cuh; nlags = 1: kalman_rows2.cuh)
// ---------------------------------------------------------------------------
template <typename R, int D_, int L_>
struct PrepSmem {
    static constexpr int n = D_ * L_, LD = n | 1;
    // one backward record per frame: [GT | h | pad] in 16-byte units
    static constexpr int RECS = ((n * n + n) * (int)sizeof(R) + 15) / 16 * 16 / (int)sizeof(R);
};

#include "kalman_rows2.cuh"
#include "kalman_split.cuh"
#include "kalman_rows_wide.cuh"

// ---------------------------------------------------------------------------
// K1d: serial affine recursion, one warp per chain, operands streamed through a
// cp.async ring in shared memory
// ---------------------------------------------------------------------------
template <typename R, int D_, int L_, int STAGES>
__global__ void __launch_bounds__(32)
kalman_affine_kernel(const R* __restrict__ GH, const int* __restrict__ mask, int T, R* __restrict__ x, int C, int W,
                     const int* __restrict__ vlen, const int* __restrict__ dirty, R* __restrict__ bx_warm,
                     R* __restrict__ bx_exact) {
    constexpr int n = D_ * L_, NN = n * n;
    constexpr int RECP = PrepSmem<R, D_, L_>::RECS;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    R* ring = reinterpret_cast<R*>(smem_raw);        // STAGES x RECP
    R* xi = ring + (size_t)STAGES * RECP;            // n
    const int nn = blockIdx.x, ck = blockIdx.y, lane = threadIdx.x;
    const int Tx = T - L_ + 1;
    if (dirty && dirty[nn] == 0) return;
    const ChunkRange cr = chunk_range(vlen ? vlen[nn] : Tx, Tx, C, W, ck);
    if (cr.empty) return;
    const R* Gn = GH + (size_t)nn * Tx * RECP;
    R* xn = x + (size_t)nn * T * D_;
    // The recursion runs downward from `top`.  The last chunk starts at the chain's terminal draw
    // xi_{Tx-1} = h_{Tx-1}; the others start W steps above their range from an arbitrary value
    // (warm-up) and publish the xi they reach at their upper boundary for the check.
    const bool exact_top = (cr.end == Tx);
    int top = exact_top ? Tx - 1 : min(cr.end + W, Tx - 1);
    const int lowest = cr.begin;
    auto issue = [&](int i) {
        if (i >= lowest) {
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
    auto emit = [&](int i, int r, R val) {           // x frames carried by xi_i
        if (i == 0) xn[r] = val;                                            // frames 0..L-1 from xi_0
        else if (r >= n - D_) xn[(size_t)(i + L_ - 1) * D_ + (r - (n - D_))] = val;
    };
    const bool top_known = (top == Tx - 1);
    for (int r = lane; r < n; r += 32) xi[r] = top_known ? Gn[(size_t)top * RECP + NN + r] : (R)0;
    if (exact_top) {
        // Masked steps carry the state through (identity records): the padded tail below the terminal
        // draw is emitted 32 steps at a time without touching its records.
        __syncwarp();
        const int* mk = mask + (size_t)nn * T + (L_ - 1);
        for (int r = lane; r < n; r += 32) emit(top, r, xi[r]);
        int i = top - 1;
        while (i >= lowest) {
            const int s = i - lane;
            const bool masked = s >= lowest && mk[s] == 0;
            const unsigned mm = __ballot_sync(0xffffffffu, masked);
            const int run = (mm == 0xffffffffu) ? 32 : __ffs(~mm) - 1;      // masked steps i, i-1, .. i-run+1
            for (int e = lane; e < run * D_; e += 32) {
                const int st = i - e / D_, c = e % D_;
                if (st == 0) {
                    for (int r = c; r < n; r += D_) xn[r] = xi[r];          // frames 0..L-1 from xi_0 (c-th components)
                } else {
                    xn[(size_t)(st + L_ - 1) * D_ + c] = xi[n - D_ + c];
                }
            }
            i -= run;
            if (run < 32) break;
        }
        top = i + 1;                       // the recursion proper resumes below step `top` (xi is the state at `top`)
        if (top - 1 < lowest) {            // nothing left: the whole chunk was masked
            if (ck > 0)
                for (int r = lane; r < n; r += 32) bx_exact[((size_t)nn * C + ck) * n + r] = xi[r];
            return;
        }
    }
    const bool emitted_top = exact_top;
    for (int s = 0; s < STAGES - 1; ++s) issue(top - 1 - s);
    __syncwarp();
    for (int r = lane; r < n; r += 32) {
        if (top < cr.end && !emitted_top) emit(top, r, xi[r]);
        if (top == cr.end) bx_warm[((size_t)nn * C + ck + 1) * n + r] = xi[r];
        if (top == lowest && ck > 0) bx_exact[((size_t)nn * C + ck) * n + r] = xi[r];
    }
    for (int i = top - 1; i >= lowest; --i) {
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
                R a2 = 0, a3 = 0;
#pragma unroll
                for (int c = 0; c + 3 < n; c += 4) {
                    a0 = fma(Gs[c * n + r], xi[c], a0);
                    a1 = fma(Gs[(c + 1) * n + r], xi[c + 1], a1);
                    a2 = fma(Gs[(c + 2) * n + r], xi[c + 2], a2);
                    a3 = fma(Gs[(c + 3) * n + r], xi[c + 3], a3);
                }
#pragma unroll
                for (int c = n / 4 * 4; c < n; ++c) a0 = fma(Gs[c * n + r], xi[c], a0);
                a0 += a2;
                a1 += a3;
            }
            nv[q] = a0 + a1;
        }
        __syncwarp();
#pragma unroll
        for (int q = 0; q < (n + 31) / 32; ++q) {
            int r = lane + 32 * q;
            if (r < n) {
                xi[r] = nv[q];
                if (i < cr.end) emit(i, r, nv[q]);
                if (i == cr.end) bx_warm[((size_t)nn * C + ck + 1) * n + r] = nv[q];
                if (i == lowest && ck > 0) bx_exact[((size_t)nn * C + ck) * n + r] = nv[q];
            }
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
// workspace: [diagnostics 256 B | vlen | dirty fwd | dirty bwd | info | stash_m | stash_S | GH |
//             boundary records of the forward filter (warm, end) and of the backward recursion (warm, exact)]
// diagnostics (unsigned[4]): max forward boundary discrepancy (float bits), forward chains re-run
// sequentially, max backward discrepancy (float bits), backward chains re-run.
enum { KW_DIAG, KW_VLEN, KW_DIRTY_F, KW_DIRTY_B, KW_INFO, KW_SM, KW_SS, KW_GH, KW_BFW, KW_BFE, KW_BBW, KW_BBE, KW_OPS, KW_WN, KW_END };

template <typename R>
static void kalman_ws_layout(int N, int T, int d, int L, int K, int C, int Cb, size_t off[KW_END + 1]) {
    const size_t n = (size_t)d * L, Tx = T - L + 1, fr = (size_t)N * Tx;
    const size_t rec = info_stride(d);
    const size_t recs = ((n * n + n) * sizeof(R) + 15) / 16 * 16 / sizeof(R);
    const size_t brec = n + n * n, nb = (size_t)N * (C + 1), nbb = (size_t)N * (Cb + 1);
    const size_t vec = 16 / sizeof(R), dp = (d + vec - 1) / vec * vec;
    const size_t ops = n * dp + dp + d * dp;             // PrepSplit::OPS: per-state operator block of the backward preparation
    size_t sz[KW_END] = {256, (size_t)N * 8, (size_t)N * 4, (size_t)N * 4, fr * rec * sizeof(R), fr * stash_m_stride((int)n) * sizeof(R),
                         fr * stash_S_stride((int)n) * sizeof(R), fr * recs * sizeof(R), nb * brec * sizeof(R), nb * brec * sizeof(R),
                         nbb * n * sizeof(R), nbb * n * sizeof(R), (size_t)K * ops * sizeof(R), (fr * n + 4) * sizeof(R)};
    off[0] = 0;
    for (int i = 0; i < KW_END; ++i) off[i + 1] = off[i] + align_up(sz[i], 256);
}

// chunks per chain for the Kalman recursions: enough chunk-CTAs to fill the device once
// float32 filters: 16 chunk-warps per SM (n <= 32), 4 two-warp teams per SM (n <= 64); shared-memory filter: 3 CTAs
static int kalman_chunks(int N, int T, int d, int L, bool backward, bool f32) {
    const int n = d * L;
    const int fwd = n <= 32 ? 16 : ((f32 && n <= 64) ? 4 : 3);
    return chunks_for(N, KPMS_SM_COUNT * (backward ? 12 : fwd), T - L + 1, chunk_config().warmup);
}

template <typename R, int D_, int L_>
static int kalman_launch(const R* Y, const int* mask, const R* v, const R* h, const R* s, const int* z,
                         const R* Ct, const R* sigmasq, const R* Ab, const R* Q, double jitter,
                         const R* w_tape, SeedArg seed, int N, int T, int k, int Dk, int K, R* x, void* ws,
                         cudaStream_t st, int stage) {
    // stage 0: the whole sampler; 1: only the per-frame observation records (kpms_kalman_obs_info, which may run on
    // another stream beside the discrete-state kernels); 2: everything but those records
    constexpr int n = D_ * L_;
    const int Tx = T - L_ + 1;
    const ChunkConfig cfg = chunk_config();
    const int C = kalman_chunks(N, T, D_, L_, false, sizeof(R) == 4), Cb = kalman_chunks(N, T, D_, L_, true, sizeof(R) == 4), W = cfg.warmup;
    const R tol = (R)(sizeof(R) == 4 ? cfg.tol32 : cfg.tol64);
    size_t off[KW_END + 1];
    kalman_ws_layout<R>(N, T, D_, L_, K, C, Cb, off);
    char* base = reinterpret_cast<char*>(ws);
    unsigned* diag = reinterpret_cast<unsigned*>(base + off[KW_DIAG]);
    int* vlen = reinterpret_cast<int*>(base + off[KW_VLEN]);
    int* dirty_f = reinterpret_cast<int*>(base + off[KW_DIRTY_F]);
    int* dirty_b = reinterpret_cast<int*>(base + off[KW_DIRTY_B]);
    R* info = reinterpret_cast<R*>(base + off[KW_INFO]);
    R* stash_m = reinterpret_cast<R*>(base + off[KW_SM]);
    R* stash_S = reinterpret_cast<R*>(base + off[KW_SS]);
    R* GH = reinterpret_cast<R*>(base + off[KW_GH]);
    R* bfw = reinterpret_cast<R*>(base + off[KW_BFW]);
    R* bfe = reinterpret_cast<R*>(base + off[KW_BFE]);
    R* bbw = reinterpret_cast<R*>(base + off[KW_BBW]);
    R* bbe = reinterpret_cast<R*>(base + off[KW_BBE]);
    R* ops = reinterpret_cast<R*>(base + off[KW_OPS]);
    R* wbuf = reinterpret_cast<R*>(base + off[KW_WN]);
    const long long frames = (long long)N * Tx;
    if (stage != 1) {
        cudaMemsetAsync(diag, 0, 256, st);
        { KPMS_LAUNCH("valid_len", st);
          valid_len_kernel<<<N, 256, 0, st>>>(mask, T, L_ - 1, Tx, vlen); }
    }
    if (stage != 2) {
        constexpr int TPB = ObsInfoCfg<D_>::template threads<R>();
        size_t smem = ((size_t)TPB * ObsInfoCfg<D_>::RS + (size_t)k * Dk * (D_ + 1) + k) * sizeof(R);
        if (smem > 220 * 1024) return set_error(-3, "kalman_sample: %d keypoints exceed the shared-memory staging of the observation records", k);
        int blocks = (int)((frames + TPB - 1) / TPB);
        KPMS_LAUNCH("kalman_obs_info", st);
        if (Dk == 2) {
            auto kern = obs_info_kernel<R, D_, 2, TPB>;
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            kern<<<blocks, TPB, smem, st>>>(Y, mask, v, h, s, sigmasq, Ct, N, T, k, L_, info);
        } else {
            auto kern = obs_info_kernel<R, D_, 3, TPB>;
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            kern<<<blocks, TPB, smem, st>>>(Y, mask, v, h, s, sigmasq, Ct, N, T, k, L_, info);
        }
        int rc = check_launch("kalman obs_info");
        if (rc) return rc;
    }
    if (stage == 1) return 0;
    if constexpr (n <= 32) {
        constexpr int WARPS = 8;
        auto kern = kalman_forward_rows_kernel<R, D_, L_, WARPS>;
        size_t smem = FwdRowsSmem<R, D_, L_>::per_warp * WARPS * sizeof(R);
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (C > 1) {
            { KPMS_LAUNCH("kalman_forward", st);
              kern<<<(int)(((long long)N * C + WARPS - 1) / WARPS), 32 * WARPS, smem, st>>>(
                  info, mask, z, Ab, Q, (R)jitter, N, T, stash_m, stash_S, C, W, vlen, vlen + N, nullptr, bfw, bfe); }
            { KPMS_LAUNCH("kalman_forward_check", st);
              cudaMemsetAsync(dirty_f, 0, (size_t)N * sizeof(int), st);
              boundary_check_kernel<R><<<dim3(C - 1, N), 128, 0, st>>>(bfw, bfe, vlen, Tx, C, W, 1, n, n + n * n, tol, dirty_f, diag); }
            { KPMS_LAUNCH("kalman_forward_rerun", st);       // exits at once for chains whose boundaries agree
              kern<<<(N + WARPS - 1) / WARPS, 32 * WARPS, smem, st>>>(info, mask, z, Ab, Q, (R)jitter, N, T, stash_m,
                                                                    stash_S, 1, 0, nullptr, vlen + N, dirty_f, bfw, bfe); }
        } else {
            KPMS_LAUNCH("kalman_forward", st);
            kern<<<(N + WARPS - 1) / WARPS, 32 * WARPS, smem, st>>>(info, mask, z, Ab, Q, (R)jitter, N, T, stash_m,
                                                                  stash_S, 1, 0, nullptr, vlen + N, nullptr, bfw, bfe);
        }
        int rc = check_launch("kalman forward");
        if (rc) return rc;
    } else if constexpr (n <= 64 && sizeof(R) == 4) {
        // two warps per (chain, chunk), one covariance row per lane (kalman_rows_wide.cuh)
        constexpr int TEAMS = 4;
        auto kern = kalman_forward_rows2w_kernel<R, D_, L_, TEAMS>;
        size_t smem = FwdRows2wSmem<R, D_, L_>::per_team * TEAMS * sizeof(R);
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (C > 1) {
            { KPMS_LAUNCH("kalman_forward", st);
              kern<<<(int)(((long long)N * C + TEAMS - 1) / TEAMS), 64 * TEAMS, smem, st>>>(
                  info, mask, z, Ab, Q, (R)jitter, N, T, stash_m, stash_S, C, W, vlen, vlen + N, nullptr, bfw, bfe); }
            { KPMS_LAUNCH("kalman_forward_check", st);
              cudaMemsetAsync(dirty_f, 0, (size_t)N * sizeof(int), st);
              boundary_check_kernel<R><<<dim3(C - 1, N), 128, 0, st>>>(bfw, bfe, vlen, Tx, C, W, 1, n, n + n * n, tol, dirty_f, diag); }
            { KPMS_LAUNCH("kalman_forward_rerun", st);       // exits at once for chains whose boundaries agree
              kern<<<(N + TEAMS - 1) / TEAMS, 64 * TEAMS, smem, st>>>(info, mask, z, Ab, Q, (R)jitter, N, T, stash_m,
                                                                    stash_S, 1, 0, nullptr, vlen + N, dirty_f, bfw, bfe); }
        } else {
            KPMS_LAUNCH("kalman_forward", st);
            kern<<<(N + TEAMS - 1) / TEAMS, 64 * TEAMS, smem, st>>>(info, mask, z, Ab, Q, (R)jitter, N, T, stash_m,
                                                                  stash_S, 1, 0, nullptr, vlen + N, nullptr, bfw, bfe);
        }
        int rc = check_launch("kalman forward");
        if (rc) return rc;
    } else {
        auto kern = kalman_forward_kernel<R, D_, L_>;
        size_t smem = FwdSmem<R, D_, L_>::bytes;
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (C > 1) {
            { KPMS_LAUNCH("kalman_forward", st);
              kern<<<dim3(N, C), 256, smem, st>>>(info, mask, z, Ab, Q, (R)jitter, T, stash_m, stash_S, C, W, vlen,
                                                  nullptr, bfw, bfe); }
            { KPMS_LAUNCH("kalman_forward_check", st);
              cudaMemsetAsync(dirty_f, 0, (size_t)N * sizeof(int), st);
              boundary_check_kernel<R><<<dim3(C - 1, N), 128, 0, st>>>(bfw, bfe, vlen, Tx, C, W, 1, n, n + n * n, tol, dirty_f, diag); }
            { KPMS_LAUNCH("kalman_forward_rerun", st);       // exits at once for chains whose boundaries agree
              kern<<<dim3(N, 1), 256, smem, st>>>(info, mask, z, Ab, Q, (R)jitter, T, stash_m, stash_S, 1, 0, nullptr,
                                                  dirty_f, bfw, bfe); }
        } else {
            KPMS_LAUNCH("kalman_forward", st);
            kern<<<dim3(N, 1), 256, smem, st>>>(info, mask, z, Ab, Q, (R)jitter, T, stash_m, stash_S, 1, 0, nullptr,
                                                nullptr, bfw, bfe);
        }
        int rc = check_launch("kalman forward");
        if (rc) return rc;
    }
    // standard normals of the backward draw: the tape in verification mode, else one Philox pass (four per call)
    if (!w_tape) {
        const long long count = frames * n;
        KPMS_LAUNCH("kalman_normals", st);
        fill_normal_kernel<R><<<(int)((count / 4 + 256) / 256), 256, 0, st>>>(wbuf, count, seed, KPMS_STREAM_X);
        w_tape = wbuf;
    }
    if (frames >= (1LL << 31)) return set_error(-3, "kalman_sample: %lld frame slots on one device exceed 2^31", frames);
    if constexpr (L_ >= 2) {
        // two-stage backward preparation (kalman_split.cuh): d lanes per frame, 32/d frames per warp
        typedef PrepSplit<R, D_, L_> PS;
        { KPMS_LAUNCH("kalman_backprep_ops", st);
          backprep_ops_kernel<R, D_, L_><<<K, 128, 0, st>>>(Ab, Q, (R)jitter, ops); }
        { KPMS_LAUNCH("kalman_backprep_special", st);
          backprep_special_kernel<R, D_, L_, stash_rowpack<R, n>()><<<N, 128, 0, st>>>(stash_m, stash_S, mask, w_tape, N, T, GH); }
        // warps x CTAs per SM: 12 warps per SM as 6 x 2 when shared memory allows (one barrier per tile keeps the
        // six warps of a CTA on the same instruction-cache lines), else the largest CTA that fits
        constexpr size_t per_warp = PS::FPW * PS::frame_bytes + 16;
        constexpr int FITW = (int)((220 * 1024) / per_warp);                 // warps per SM by shared memory
        static_assert(FITW >= 1, "one warp of the two-stage backward preparation must fit in shared memory");
        constexpr int WARPS = FITW >= 12 ? 6 : (FITW >= 6 ? (FITW / 2 < 6 ? FITW / 2 : 6) : FITW);
        constexpr int MINB = sizeof(R) == 8 ? 1 : (FITW / WARPS >= 2 ? 2 : 1);
        auto launch = [&](auto kern, int warps, int minb) {
            const size_t smem = (size_t)warps * PS::FPW * PS::frame_bytes + warps * 2 * sizeof(uint64_t);
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            const long long tiles = (frames + warps * PS::FPW - 1) / (warps * PS::FPW);
            const int blocks = (int)std::min<long long>(tiles, (long long)KPMS_SM_COUNT * minb);
            KPMS_LAUNCH("kalman_backprep", st);
            kern<<<blocks, 32 * warps, smem, st>>>(stash_m, stash_S, mask, z, ops, (R)(KPMS_EPS_SHIFT + jitter), w_tape, N, T, GH);
        };
        bool done = false;
        if constexpr (sizeof(R) == 4 && D_ == 10 && L_ == 3) {
            // A/B switch: KPMS_BP_CFG=4x3 is the round-2 layout without the barrier (5.47 ms at C2 against 4.56 ms; two
            // or three barriers per tile: 4.66 / 4.64 ms; 5 warps x 2 CTAs with 204 registers: 5.23 ms)
            static const std::string cfg = [] { const char* e = getenv("KPMS_BP_CFG"); return std::string(e ? e : ""); }();
            if (cfg == "4x3") { launch(kalman_backprep_split_kernel<R, D_, L_, 4, 3, 0>, 4, 3); done = true; }
        }
        if (!done) launch(kalman_backprep_split_kernel<R, D_, L_, WARPS, MINB, 1>, WARPS, MINB);
        int rc = check_launch("kalman backprep (two-stage)");
        if (rc) return rc;
    } else {
        // nlags = 1 has no shifted block to condition on first: one-stage form, two rows per lane (kalman_rows2.cuh)
        static_assert(n <= 32, "the one-stage backward preparation holds one frame per half warp");
        typedef PrepRows2<R, D_, L_> P2;
        constexpr int WARPS = sizeof(R) == 4 ? 8 : 4;
        auto kern = kalman_backprep_rows2_kernel<R, D_, L_, WARPS>;
        size_t smem = (size_t)P2::per_group * P2::FPW * WARPS * sizeof(R);
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        int blocks = (int)std::min<long long>((frames + WARPS * P2::FPW - 1) / (WARPS * P2::FPW), (long long)KPMS_SM_COUNT);
        { KPMS_LAUNCH("kalman_backprep", st); kern<<<blocks, 32 * WARPS, smem, st>>>(stash_m, stash_S, mask, z, Ab, Q, (R)jitter, w_tape, seed, N, T, GH); }
        int rc = check_launch("kalman backprep");
        if (rc) return rc;
    }
    {
        constexpr int STAGES = 4;
        auto kern = kalman_affine_kernel<R, D_, L_, STAGES>;
        size_t smem = ((size_t)STAGES * PrepSmem<R, D_, L_>::RECS + n) * sizeof(R);
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (Cb > 1) {
            { KPMS_LAUNCH("kalman_affine", st);
              kern<<<dim3(N, Cb), 32, smem, st>>>(GH, mask, T, x, Cb, W, vlen, nullptr, bbw, bbe); }
            { KPMS_LAUNCH("kalman_affine_check", st);
              cudaMemsetAsync(dirty_b, 0, (size_t)N * sizeof(int), st);
              boundary_check_kernel<R><<<dim3(Cb - 1, N), 128, 0, st>>>(bbw, bbe, vlen, Tx, Cb, W, 1, n, n, tol, dirty_b, diag + 2); }
            { KPMS_LAUNCH("kalman_affine_rerun", st);
              kern<<<dim3(N, 1), 32, smem, st>>>(GH, mask, T, x, 1, 0, nullptr, dirty_b, bbw, bbe); }
        } else {
            KPMS_LAUNCH("kalman_affine", st);
            kern<<<dim3(N, 1), 32, smem, st>>>(GH, mask, T, x, 1, 0, nullptr, nullptr, bbw, bbe);
        }
        int rc = check_launch("kalman affine");
        if (rc) return rc;
    }
    return 0;
}

// (latent_dim, nlags) pairs of this translation unit's group (common.cuh: KPMS_DL_GROUP_g, -DKPMS_DL_GROUP=g)
template <typename R>
static int kalman_group_impl(const void* Y, const int* mask, const void* v, const void* h, const void* s, const int* z,
                             const void* Ct, const void* sigmasq, const void* Ab, const void* Q, double jitter,
                             const void* w_tape, SeedArg seed, int N, int T, int k, int Dk, int d, int L, int K, void* x,
                             void* ws, cudaStream_t st, int stage) {
#define X(DD, LL)                                                                                            \
    if (d == DD && L == LL)                                                                                  \
        return kalman_launch<R, DD, LL>((const R*)Y, mask, (const R*)v, (const R*)h, (const R*)s, z,         \
                                        (const R*)Ct, (const R*)sigmasq, (const R*)Ab, (const R*)Q, jitter,  \
                                        (const R*)w_tape, seed, N, T, k, Dk, K, (R*)x, ws, st, stage);
    KPMS_FOR_GROUP_DL(X)
#undef X
    return KPMS_NOT_IN_GROUP;
}

#define KPMS_KALMAN_GROUP_ARGS                                                                                    \
    int dtype, const void *Y, const int *mask, const void *v, const void *h, const void *s, const int *z,         \
        const void *Ct, const void *sigmasq, const void *Ab, const void *Q, double jitter, const void *w_tape,    \
        SeedArg seed, int N, int T, int k, int Dk, int d, int L, int K, void *x, void *ws, cudaStream_t st, int stage

int KPMS_CAT(kalman_group_, KPMS_DL_GROUP)(KPMS_KALMAN_GROUP_ARGS) {
    return KPMS_DISPATCH_DTYPE(dtype, kalman_group_impl, Y, mask, v, h, s, z, Ct, sigmasq, Ab, Q, jitter, w_tape, seed, N, T,
                               k, Dk, d, L, K, x, ws, st, stage);
}

#if KPMS_DL_GROUP == 0
int kalman_group_1(KPMS_KALMAN_GROUP_ARGS);
int kalman_group_2(KPMS_KALMAN_GROUP_ARGS);
int kalman_group_3(KPMS_KALMAN_GROUP_ARGS);
int kalman_group_4(KPMS_KALMAN_GROUP_ARGS);
int kalman_group_5(KPMS_KALMAN_GROUP_ARGS);
int kalman_group_6(KPMS_KALMAN_GROUP_ARGS);
int kalman_group_7(KPMS_KALMAN_GROUP_ARGS);
static_assert(KPMS_DL_GROUPS == 8, "one dispatcher per group");

static int kalman_dispatch(KPMS_KALMAN_GROUP_ARGS) {
    if (dtype != 0 && dtype != 1) return set_error(-2, "dtype must be 0 (f32) or 1 (f64), got %d", dtype);
    if (K < 1) return set_error(-3, "kalman_sample: num_states must be positive, got %d", K);
    if (Dk != 2 && Dk != 3) return set_error(-3, "kalman_sample: keypoint dimension must be 2 or 3, got %d", Dk);
    if (T < L) return set_error(-3, "kalman_sample: T (%d) < nlags (%d)", T, L);
    typedef int (*GroupFn)(KPMS_KALMAN_GROUP_ARGS);
    static const GroupFn groups[KPMS_DL_GROUPS] = {kalman_group_0, kalman_group_1, kalman_group_2, kalman_group_3,
                                                   kalman_group_4, kalman_group_5, kalman_group_6, kalman_group_7};
    for (int g = 0; g < KPMS_DL_GROUPS; ++g) {
        const int rc = groups[g](dtype, Y, mask, v, h, s, z, Ct, sigmasq, Ab, Q, jitter, w_tape, seed, N, T, k, Dk, d, L, K,
                                 x, ws, st, stage);
        if (rc != KPMS_NOT_IN_GROUP) return rc;
    }
    return set_error(-3, "kalman_sample: unsupported (latent_dim, nlags) = (%d, %d); see kpms_supported_dims", d, L);
}
#endif

}  // namespace kpms

#if KPMS_DL_GROUP == 0
using namespace kpms;

extern "C" {

size_t kpms_kalman_workspace_bytes(int dtype, int N, int T, int d, int L, int K) {
    size_t off[KW_END + 1];
    const int C = kalman_chunks(N, T, d, L, false, dtype == 0), Cb = kalman_chunks(N, T, d, L, true, dtype == 0);
    if (dtype == 0) kalman_ws_layout<float>(N, T, d, L, K, C, Cb, off);
    else kalman_ws_layout<double>(N, T, d, L, K, C, Cb, off);
    return off[KW_END];
}

int kpms_kalman_sample(int dtype, const void* Y, const int* mask, const void* v, const void* h, const void* s,
                       const int* z, const void* Ct, const void* sigmasq, const void* Ab, const void* Q,
                       double jitter, const void* w_tape, uint64_t seed, const uint64_t* seed_dev, int N, int T, int k, int Dk, int d,
                       int L, int K, int info_ready, void* x, void* ws, void* stream) {
    return kalman_dispatch(dtype, Y, mask, v, h, s, z, Ct, sigmasq, Ab, Q, jitter, w_tape, SeedArg(seed, seed_dev), N, T, k,
                           Dk, d, L, K, x, ws, (cudaStream_t)stream, info_ready ? 2 : 0);
}

int kpms_kalman_obs_info(int dtype, const void* Y, const int* mask, const void* v, const void* h, const void* s,
                         const void* Ct, const void* sigmasq, int N, int T, int k, int Dk, int d, int L, int K, void* ws,
                         void* stream) {
    return kalman_dispatch(dtype, Y, mask, v, h, s, nullptr, Ct, sigmasq, nullptr, nullptr, 0.0, nullptr, SeedArg(), N, T, k,
                           Dk, d, L, K, nullptr, ws, (cudaStream_t)stream, 1);
}

/* (latent_dim, nlags) pairs the library was built for: writes up to `cap` pairs as d0, L0, d1, L1, ... and returns
 * how many pairs exist. */
int kpms_supported_dims(int* pairs, int cap) {
    static const int table[][2] = {
#define X(DD, LL) {DD, LL},
        KPMS_FOR_EACH_DL(X)
#undef X
    };
    const int count = (int)(sizeof(table) / sizeof(table[0]));
    for (int i = 0; i < count && i < cap; ++i) { pairs[2 * i] = table[i][0]; pairs[2 * i + 1] = table[i][1]; }
    return count;
}

}  // extern "C"
#endif
