// AR-HMM path: per-frame AR log-likelihood weights (K2) and HMM forward filter /
// backward sampler (K3).  Replaces jax_moseq.models.arhmm.resample_discrete_stateseqs,
// marginal_log_likelihood and stateseq_marginals, reached from
// keypoint_moseq/fitting.py:25, :536-538, :667-673.
//
// Layouts (row-major):
//   x     (N, T, d)            latent trajectories
//   mask  (N, T) int32
//   W     float32: (N, K, ldT), W[n][k][t'] = exp(ll[n][t'][k] - mx[n][t']),  t' = t - L
//         float64: (N, T', 8*ceil(K/8)) state-contiguous (operand layout of the tensor-pipe kernels, hmm_f64.cuh)
//   mx    (N, ldT)             per-frame max log-likelihood (0 on masked frames)
//   filt  (N, T', ldK)         filtered state probabilities
//   z     (N, T') int32
#include <algorithm>
#ifndef KPMS_DL_GROUP
#define KPMS_DL_GROUP 0
#endif
#include "common.cuh"
#include "../../include/kpms_b200.h"

namespace kpms {

// ---------------------------------------------------------------------------
// per-state whitened regression operator: G_k = Lq_k^{-1} [ -A_k | I | -b_k ]  (d x F),
// F = n + d + 1, features f = [x_{t-L} .. x_{t-1} | x_t | 1];  c_k = -sum log diag Lq - d/2 log 2pi
// ---------------------------------------------------------------------------
template <typename R, int D_, int L_>
__global__ void __launch_bounds__(32)
ar_prep_kernel(const R* __restrict__ Ab, const R* __restrict__ Q, int K,
               R* __restrict__ G, R* __restrict__ cst, int Fp) {
    // one warp per state: the d x d factor and its inverse redundantly per lane (a few hundred
    // flops), then the lanes share the columns of G
    constexpr int n = D_ * L_;
    const int k = blockIdx.x, lane = threadIdx.x;
    if (k >= K) return;
    double Lq[D_][D_], Li[D_][D_];
    for (int i = 0; i < D_; ++i)
        for (int j = 0; j < D_; ++j) { Lq[i][j] = (double)Q[(k * D_ + i) * D_ + j]; Li[i][j] = 0.0; }
    double logdet = 0.0;
    for (int j = 0; j < D_; ++j) {
        double s = Lq[j][j];
        for (int p = 0; p < j; ++p) s -= Lq[j][p] * Lq[j][p];
        s = sqrt(s);
        Lq[j][j] = s;
        logdet += log(s);
        for (int i = j + 1; i < D_; ++i) {
            double v = Lq[i][j];
            for (int p = 0; p < j; ++p) v -= Lq[i][p] * Lq[j][p];
            Lq[i][j] = v / s;
        }
    }
    for (int c = 0; c < D_; ++c) {       // Li = Lq^{-1}, column by column
        for (int i = c; i < D_; ++i) {
            double v = (i == c) ? 1.0 : 0.0;
            for (int p = c; p < i; ++p) v -= Lq[i][p] * Li[p][c];
            Li[i][c] = v / Lq[i][i];
        }
    }
    const R* A = Ab + (size_t)k * D_ * (n + 1);
    R* g = G + (size_t)k * D_ * Fp;
    for (int j = lane; j <= n; j += 32) {
        double col[D_];
        for (int p = 0; p < D_; ++p) col[p] = (double)A[p * (n + 1) + j];
        for (int i = 0; i < D_; ++i) {
            double v = 0.0;
            for (int p = 0; p <= i; ++p) v += Li[i][p] * col[p];
            if (j < n) g[i * Fp + j] = (R)(-v);
            else g[i * Fp + n + D_] = (R)(-v);
        }
    }
    for (int e = lane; e < D_ * D_; e += 32) g[(e / D_) * Fp + n + (e % D_)] = (R)Li[e / D_][e % D_];
    for (int i = lane; i < D_; i += 32)
        for (int j = n + D_ + 1; j < Fp; ++j) g[i * Fp + j] = (R)0;
    if (lane == 0) cst[k] = (R)(-logdet - 0.5 * D_ * 1.8378770664093453);
}

// ---------------------------------------------------------------------------
// K2: log-likelihood weights.  One thread owns FPT frames (feature vectors in
// registers); the whitened operators stream through shared memory in state chunks
// and are read as broadcasts.  Output is written time-contiguous per state so both
// this kernel's stores and the filter's per-column streams are coalesced.
// ---------------------------------------------------------------------------
template <typename R, int D_, int L_, int FPT, int KC>
__global__ void __launch_bounds__(128)
ar_loglik_kernel(const R* __restrict__ x, const int* __restrict__ mask, const R* __restrict__ G,
                 const R* __restrict__ cst, int N, int T, int K, int Fp, int ldT,
                 R* __restrict__ W, R* __restrict__ mx) {
    constexpr int n = D_ * L_;
    constexpr int NF = n + D_;
    constexpr int FR = 128 * FPT;
    constexpr int VEC = 16 / sizeof(R);
    typedef typename Vec16<R>::type VecT;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    R* xs = reinterpret_cast<R*>(smem_raw);                       // (FR + L) * d
    R* Gs = xs + align_up((size_t)(FR + L_) * D_, 4);             // KC * d * Fp
    R* cs = Gs + (size_t)KC * D_ * Fp;                            // KC
    const int nn = blockIdx.y;
    const int Tp = T - L_;
    const int t0 = blockIdx.x * FR;                               // first t' of this tile
    const int tid = threadIdx.x;
    const R* xrow = x + (size_t)nn * T * D_;
    const int tile_vals = min(FR + L_, T - t0) * D_;
    for (int i = tid; i < tile_vals; i += 128) xs[i] = xrow[(size_t)t0 * D_ + i];
    __syncthreads();
    R f[FPT][NF];
    bool valid[FPT], on[FPT];
    R best[FPT];
#pragma unroll
    for (int q = 0; q < FPT; ++q) {
        int lt = tid + q * 128;
        valid[q] = (t0 + lt) < Tp;
        on[q] = valid[q] && mask[(size_t)nn * T + t0 + lt + L_] != 0;
        best[q] = (R)(-INFINITY);
#pragma unroll
        for (int j = 0; j < NF; ++j) f[q][j] = valid[q] ? xs[lt * D_ + j] : (R)0;
    }
    R* Wn = W + (size_t)nn * K * ldT;
    for (int k0 = 0; k0 < K; k0 += KC) {
        const int kc = min(KC, K - k0);
        __syncthreads();
        for (int i = tid; i < kc * D_ * Fp; i += 128) Gs[i] = G[(size_t)k0 * D_ * Fp + i];
        for (int i = tid; i < kc; i += 128) cs[i] = cst[k0 + i];
        __syncthreads();
        for (int kk = 0; kk < kc; ++kk) {
            R acc[FPT];
#pragma unroll
            for (int q = 0; q < FPT; ++q) acc[q] = (R)0;
#pragma unroll 2
            for (int i = 0; i < D_; ++i) {
                // 16-byte broadcast loads of the operator row; index NF is the bias column
                const VecT* g = reinterpret_cast<const VecT*>(Gs + (size_t)(kk * D_ + i) * Fp);
                R r[FPT];
#pragma unroll
                for (int q = 0; q < FPT; ++q) r[q] = (R)0;
#pragma unroll
                for (int jv = 0; jv < (NF + VEC) / VEC; ++jv) {
                    VecT gv = g[jv];
                    const R* ge = reinterpret_cast<const R*>(&gv);
#pragma unroll
                    for (int c = 0; c < VEC; ++c) {
                        const int j = jv * VEC + c;
                        if (j < NF) {
#pragma unroll
                            for (int q = 0; q < FPT; ++q) r[q] = fma(ge[c], f[q][j], r[q]);
                        } else if (j == NF) {
#pragma unroll
                            for (int q = 0; q < FPT; ++q) r[q] += ge[c];
                        }
                    }
                }
#pragma unroll
                for (int q = 0; q < FPT; ++q) acc[q] = fma(r[q], r[q], acc[q]);
            }
#pragma unroll
            for (int q = 0; q < FPT; ++q) {
                if (valid[q]) {
                    R ll = on[q] ? (R)(-0.5) * acc[q] + cs[kk] : (R)0;
                    best[q] = ll > best[q] ? ll : best[q];
                    Wn[(size_t)(k0 + kk) * ldT + t0 + tid + q * 128] = ll;
                }
            }
        }
    }
    // second pass over this thread's own stores: W = exp(ll - max)
#pragma unroll
    for (int q = 0; q < FPT; ++q) {
        if (!valid[q]) continue;
        int tp = t0 + tid + q * 128;
        mx[(size_t)nn * ldT + tp] = best[q];
        for (int k = 0; k < K; ++k) {
            R* p = Wn + (size_t)k * ldT + tp;
            *p = exp(*p - best[q]);
        }
    }
}

// ---------------------------------------------------------------------------
// K3 forward: scaled filter.  One CTA advances M independent (chain, time chunk) tasks in lockstep,
// 4 threads per state column with the transition-matrix column slices resident in registers and
// shared by the M tasks; one barrier per step; the weight rows arrive through a cp.async ring,
// 8 steps at a time.
//
// Time chunks (common.cuh).  The unmasked prefix of a chain, rounded up to 8 steps (vb), is cut into
// C chunks that start Wm steps early from the uniform prior and are checked against their
// neighbours (1e-12: far below the spacing of the inverse-CDF thresholds, so the labels stay those
// of the sequential filter).  Padded frames carry no likelihood: there the prediction obeys
// p_{t+TL} = (pi^TL)' p_t exactly, so the padded tail is cut into TL-step chunks whose starting
// predictions are propagated with the precomputed power of pi (pass 1) - no warm-up, no check.
// pass 0: prefix chunks, pass 1: tail chunks, pass 2: whole chains flagged dirty, sequentially.
// ---------------------------------------------------------------------------
constexpr int HMM_TL = 128;         // steps per padded-tail chunk (power of two)

struct HmmTask { int nn, begin, end, start, slot; bool on, given; };

__device__ inline HmmTask hmm_task(long long id, int pass, int N, int Tp, int C, int CT, int Wm,
                                   const int* __restrict__ vb, const int* __restrict__ dirty) {
    HmmTask t;
    t.on = false; t.given = false; t.nn = 0; t.begin = t.end = t.start = 0; t.slot = 0;
    if (pass == 0) {
        const int nn = (int)(id / C), ck = (int)(id % C);
        if (nn >= N) return t;
        const int v = vb ? vb[nn] : Tp;
        const ChunkRange cr = chunk_range(v, v, C, Wm, ck, 8);
        if (cr.empty || cr.begin >= cr.end) return t;
        t.on = true; t.nn = nn; t.begin = cr.begin; t.end = cr.end; t.start = cr.start; t.slot = ck;
    } else if (pass == 1) {
        const int nn = (int)(id / CT), k = (int)(id % CT);
        if (nn >= N || dirty[nn] != 0) return t;
        const int b = vb[nn] + k * HMM_TL;
        if (b >= Tp) return t;
        t.on = true; t.given = true; t.nn = nn; t.begin = t.start = b; t.end = min(b + HMM_TL, Tp); t.slot = k;
    } else if (pass == 3) {
        // refinement: the prefix chunks of the chains still flagged, restarted (no warm-up) from the end
        // state their predecessor produced in the previous pass; chains with mask holes stay sequential
        const int nn = (int)(id / C), ck = (int)(id % C);
        if (nn >= N || dirty[nn] == 0 || vb[N + nn] != 0) return t;
        const int v = vb[nn];
        const ChunkRange cr = chunk_range(v, v, C, Wm, ck, 8);
        if (cr.empty || cr.begin >= cr.end) return t;
        t.on = true; t.given = ck > 0; t.nn = nn; t.begin = t.start = cr.begin; t.end = cr.end; t.slot = ck;
    } else {
        const int nn = (int)id;
        if (nn >= N || (dirty && dirty[nn] == 0)) return t;
        t.on = true; t.nn = nn; t.begin = t.start = 0; t.end = Tp; t.slot = 0;
    }
    return t;
}

// After a refinement pass: boundary c of a refined chain compares the start chunk c used (`cur`, the end state
// chunk c-1 had produced one pass earlier) with the end state chunk c-1 produced in this pass (`fresh`),
// component by component (see boundary_check_kernel<COMPONENTWISE>), then makes the fresh state current.
// When every boundary of a chain agrees, its chunks all started from what their predecessors now end in, so
// the concatenation IS the sequential filter to that tolerance.  Grid (C, N); boundary C feeds the padded tail.
template <typename R>
__global__ void __launch_bounds__(128)
hmm_refine_check_kernel(R* __restrict__ cur, const R* __restrict__ fresh, const int* __restrict__ vb, int N, int Tp,
                        int C, int Wm, int K, R tol, const int* __restrict__ dirty_in, int* __restrict__ dirty_out,
                        unsigned* __restrict__ left) {
    __shared__ R red[4];
    const int nn = blockIdx.y, c = blockIdx.x + 1, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (dirty_in[nn] == 0) return;
    if (vb[N + nn] != 0) {                                   // mask holes: sequential re-run, as before
        if (c == 1 && threadIdx.x == 0 && atomicExch(&dirty_out[nn], 1) == 0) atomicAdd(left, 1u);
        return;
    }
    const int v = vb[nn];
    const ChunkRange prev = chunk_range(v, v, C, Wm, c - 1, 8);
    if (prev.empty || prev.begin >= prev.end || prev.end >= Tp) return;       // no end state was written
    bool compare = false;
    if (c < C) {
        const ChunkRange me = chunk_range(v, v, C, Wm, c, 8);
        compare = !me.empty && me.begin < me.end;
    }
    R* a = cur + ((size_t)nn * C + c) * K;
    const R* b = fresh + ((size_t)nn * C + c) * K;
    R worst = 0;
    int bad = 0;
    for (int e = threadIdx.x; e < K; e += blockDim.x) {
        const R x = a[e], y = b[e];
        if (compare) {
            R dd = fabs(x - y);
            if (!(dd < (R)INFINITY)) bad = 1;
            const R big = fmax(fabs(x), fabs(y));
            worst = fmax(worst, big > (R)0 ? dd / big : (R)0);
        }
        a[e] = y;
    }
    worst = warp_max(worst);
    if (lane == 0) red[warp] = worst;
    const int any_bad = __syncthreads_or(bad);
    if (threadIdx.x == 0 && compare) {
        worst = fmax(fmax(red[0], red[1]), fmax(red[2], red[3]));
        if ((any_bad || !(worst <= tol)) && atomicExch(&dirty_out[nn], 1) == 0) atomicAdd(left, 1u);
    }
}

// ROWW: the weights are (N, Tp, ldW) state-contiguous (float64 path) instead of (N, K, ldT).
template <typename R, int RPT, int M, bool ROWW = false>
__global__ void __launch_bounds__(4 * 128)
hmm_forward_kernel(const R* __restrict__ W, const R* __restrict__ mx, const R* __restrict__ pi, int N, int K,
                   int Tp, int ldT, int ldK, R* __restrict__ filt, double* __restrict__ logZ,
                   double* __restrict__ logZ_part, int pass, int C, int CT, int Wm,
                   const int* __restrict__ vb, const int* __restrict__ dirty, R* __restrict__ bnd_warm,
                   R* __restrict__ bnd_end, const R* __restrict__ tail_start) {
    constexpr int VEC = 16 / sizeof(R);
    constexpr int RPTP = (RPT + VEC - 1) / VEC * VEC;
    constexpr int CH = 8;                                    // steps per staged group (ldT % 8 == 0)
    constexpr int KC = ROWW ? (4 * RPT + 7) / 8 * 8 : 4 * RPT;   // columns covered by the thread grid / row stride
    extern __shared__ __align__(16) unsigned char smem_raw[];
    R* wbuf = reinterpret_cast<R*>(smem_raw);                // 2 x M x KC x CH  (ROWW: 2 x M x CH x KC)
    R* qbuf = wbuf + 2 * M * KC * CH;                        // 2 x M x 4*RPTP
    __shared__ double red[32];
    const int tid = threadIdx.x;
    const int j = tid >> 2, p = tid & 3;
    const bool col = j < K;
    HmmTask tk[M];
    int maxlen = 0;
#pragma unroll
    for (int m = 0; m < M; ++m) {
        tk[m] = hmm_task((long long)blockIdx.x * M + m, pass, N, Tp, C, CT, Wm, vb, dirty);
        if (tk[m].on) maxlen = max(maxlen, tk[m].end - tk[m].start);
    }
    if (maxlen == 0) return;
    R pic[RPT];
#pragma unroll
    for (int r = 0; r < RPT; ++r) {
        int i = p * RPT + r;
        pic[r] = (col && i < K) ? pi[(size_t)i * K + j] : (R)0;
    }
    for (int i = tid; i < 2 * M * 4 * RPTP; i += blockDim.x) qbuf[i] = (R)0;
    // sums of the per-frame maxima (part of the log-normaliser), one block reduction per task
    double msum[M];
#pragma unroll
    for (int m = 0; m < M; ++m) {
        double acc = 0.0;
        if (tk[m].on)
            for (int t = tk[m].begin + tid; t < tk[m].end; t += blockDim.x) acc += (double)mx[(size_t)tk[m].nn * ldT + t];
        msum[m] = block_sum(acc, red);
    }
    const int qslot = (j / RPT) * RPTP + (j % RPT);
    R pred[M], inv_s[M], qprev[M];
    double lz[M], lzp[M];
    int lze[M];
#pragma unroll
    for (int m = 0; m < M; ++m) {
        pred[m] = (R)1 / (R)K;
        if (tk[m].on && tk[m].given && col) pred[m] = tail_start[((size_t)tk[m].nn * CT + tk[m].slot) * K + j];
        inv_s[m] = (R)1;
        qprev[m] = (R)0;
        lz[m] = 0.0;
        lzp[m] = 1.0;
        lze[m] = 0;
    }
    // staging of the weights: thread (j, p) moves the p-th 16-byte piece of column j's 8-step group
    auto stage = [&](int r0, int sbuf) {
        if (ROWW) {                                          // one 16-byte piece of the 8 x KC tile per thread
            constexpr int PPR = KC / VEC;                    // pieces per row
            const int rr = tid / PPR, pc = tid % PPR;
            if (rr < CH) {
#pragma unroll
                for (int m = 0; m < M; ++m) {
                    const int t0 = tk[m].start + r0 + rr;
                    if (tk[m].on && t0 < tk[m].end)
                        cp_async_16(wbuf + ((size_t)(sbuf * M + m) * CH + rr) * KC + pc * VEC,
                                    W + ((size_t)tk[m].nn * Tp + t0) * KC + pc * VEC);
                }
            }
        } else if (col && p < CH / VEC) {
#pragma unroll
            for (int m = 0; m < M; ++m) {
                const int t0 = tk[m].start + r0;
                if (tk[m].on && t0 < tk[m].end)
                    cp_async_16(wbuf + ((size_t)(sbuf * M + m) * KC + j) * CH + p * VEC,
                                W + ((size_t)tk[m].nn * K + j) * ldT + t0 + p * VEC);
            }
        }
        asm volatile("cp.async.commit_group;\n" ::);
    };
    stage(0, 0);
    int buf = 0;
    for (int r0 = 0; r0 < maxlen; r0 += CH) {
        const int sbuf = (r0 / CH) & 1;
        stage(r0 + CH, sbuf ^ 1);
        asm volatile("cp.async.wait_group 1;\n" ::);
        __syncthreads();
#pragma unroll
        for (int c = 0; c < CH; ++c) {
            const int r = r0 + c;
            if (r >= maxlen) break;
            R qv[M];
#pragma unroll
            for (int m = 0; m < M; ++m) {
                const int t = tk[m].start + r;
                if (tk[m].on && t < tk[m].end) {
                    if (col && p == 0 && !tk[m].given && tk[m].slot > 0 && t == tk[m].begin && pass == 0)
                        bnd_warm[((size_t)tk[m].nn * C + tk[m].slot) * K + j] = pred[m] * inv_s[m];
                    const R w = !col ? (R)0 : ROWW ? wbuf[((size_t)(sbuf * M + m) * CH + c) * KC + j]
                                                   : wbuf[((size_t)(sbuf * M + m) * KC + j) * CH + c];
                    qv[m] = pred[m] * inv_s[m] * w;
                    if (col && p == 0) {
                        qbuf[(buf * M + m) * 4 * RPTP + qslot] = qv[m];
                        if (t > tk[m].begin) filt[((size_t)tk[m].nn * Tp + t - 1) * ldK + j] = qprev[m] * inv_s[m];
                    }
                }
            }
            __syncthreads();
#pragma unroll
            for (int m = 0; m < M; ++m) {
                const int t = tk[m].start + r;
                if (tk[m].on && t < tk[m].end) {
                    const R* qs = qbuf + (buf * M + m) * 4 * RPTP + p * RPTP;
                    R a0 = 0, a1 = 0, s0 = 0, s1 = 0;
#pragma unroll
                    for (int rr = 0; rr + 1 < RPT; rr += 2) {
                        R q0 = qs[rr], q1 = qs[rr + 1];
                        a0 = fma(pic[rr], q0, a0);
                        a1 = fma(pic[rr + 1], q1, a1);
                        s0 += q0;
                        s1 += q1;
                    }
                    if (RPT & 1) { R q0 = qs[RPT - 1]; a0 = fma(pic[RPT - 1], q0, a0); s0 += q0; }
                    R a = a0 + a1, s = s0 + s1;
                    a += __shfl_xor_sync(0xffffffffu, a, 1);
                    s += __shfl_xor_sync(0xffffffffu, s, 1);
                    a += __shfl_xor_sync(0xffffffffu, a, 2);
                    s += __shfl_xor_sync(0xffffffffu, s, 2);
                    pred[m] = a;
                    inv_s[m] = rcp_fast<R>(s);              // any common scale of a step's row is immaterial
                    qprev[m] = qv[m];
                    if (tid == 4 * m && t >= tk[m].begin) {    // log s accumulated as mantissa product + exponent
                        int ex;
                        lzp[m] *= frexp((double)s, &ex);
                        lze[m] += ex;
                        if ((r & 7) == 7) { lz[m] += log(lzp[m]); lzp[m] = 1.0; }
                    }
                }
            }
            buf ^= 1;
        }
    }
#pragma unroll
    for (int m = 0; m < M; ++m) {
        if (!tk[m].on) continue;
        if (col && p == 0) {
            filt[((size_t)tk[m].nn * Tp + tk[m].end - 1) * ldK + j] = qprev[m] * inv_s[m];
            if (pass == 0 && tk[m].end < Tp)               // handed to the next chunk / to the padded tail
                bnd_end[((size_t)tk[m].nn * C + tk[m].slot + 1) * K + j] = pred[m] * inv_s[m];
        }
        if (tid == 4 * m) {
            const double val = lz[m] + log(lzp[m]) + 0.6931471805599453094 * (double)lze[m] + msum[m];
            if (pass == 2) logZ[tk[m].nn] = val;
            else logZ_part[(size_t)tk[m].nn * (C + CT) + (pass == 1 ? C + tk[m].slot : tk[m].slot)] = val;
        }
    }
}

#include "hmm_f64.cuh"
#include "hmm_wide.cuh"

#if KPMS_DL_GROUP == 0
// logZ[nn] = ordered sum of the chunks' parts (zero-initialised; chains flagged dirty are
// overwritten by the sequential pass afterwards)
__global__ void logz_sum_kernel(const double* __restrict__ part, int N, int parts, double* __restrict__ logZ) {
    const int nn = blockIdx.x * blockDim.x + threadIdx.x;
    if (nn >= N) return;
    double acc = 0.0;
    for (int c = 0; c < parts; ++c) acc += part[(size_t)nn * parts + c];
    logZ[nn] = acc;
}
#endif

// vb[nn] = length of the leading unmasked run of steps, rounded up to 8 (capped at Tp);
// holes[nn] = 1 when a valid frame follows a masked one (such chains run sequentially)
static __global__ void __launch_bounds__(256)
hmm_prefix_kernel(const int* __restrict__ mask, int T, int L, int Tp, int* __restrict__ vb, int* __restrict__ holes) {
    __shared__ int first, late;
    const int nn = blockIdx.x;
    if (threadIdx.x == 0) { first = Tp; late = 0; }
    __syncthreads();
    const int* mk = mask + (size_t)nn * T + L;
    for (int i = threadIdx.x; i < Tp; i += blockDim.x)
        if (mk[i] == 0) { atomicMin(&first, i); break; }
    __syncthreads();
    const int f = first;
    for (int i = f + threadIdx.x; i < Tp; i += blockDim.x)
        if (mk[i] != 0) { late = 1; break; }
    __syncthreads();
    if (threadIdx.x == 0) { vb[nn] = min(Tp, (f + 7) / 8 * 8); holes[nn] = late; }
}

// out = A A (K x K, row-major); used to form pi^(2^k)
template <typename R>
__global__ void mat_square_kernel(const R* __restrict__ A, int K, R* __restrict__ out) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= K * K) return;
    const int i = idx / K, jn = idx % K;
    R acc = 0;
    for (int k = 0; k < K; ++k) acc = fma(A[i * K + k], A[k * K + jn], acc);
    out[idx] = acc;
}

// predictions at the start of every padded-tail chunk: start[0] = state handed over by the last prefix
// chunk (uniform when the chain has no valid frame), start[k] = (pi^TL)' start[k-1].  One CTA per chain.
template <typename R>
__global__ void __launch_bounds__(128)
hmm_tail_starts_kernel(const R* __restrict__ PTL, const R* __restrict__ bnd_end, const int* __restrict__ vb,
                       const int* __restrict__ dirty, int K, int Tp, int C, int CT, int Wm,
                       R* __restrict__ tail_start) {
    __shared__ R pcur[128];
    const int nn = blockIdx.x, jn = threadIdx.x;
    const int v = vb[nn];
    if (dirty[nn] != 0 || v >= Tp) return;
    int last = -1;                                          // last non-empty prefix chunk
    for (int c = 0; c < C; ++c) {
        const ChunkRange cr = chunk_range(v, v, C, Wm, c, 8);
        if (!cr.empty && cr.begin < cr.end) last = c;
    }
    R val = (R)1 / (R)K;
    if (last >= 0 && jn < K) val = bnd_end[((size_t)nn * C + last + 1) * K + jn];
    const int nk = (Tp - v + HMM_TL - 1) / HMM_TL;
    for (int k = 0; k < nk; ++k) {
        if (jn < K) tail_start[((size_t)nn * CT + k) * K + jn] = val;
        __syncthreads();
        pcur[jn] = (jn < K) ? val : (R)0;
        __syncthreads();
        R acc = 0;
        if (jn < K)
            for (int i = 0; i < K; ++i) acc = fma(pcur[i], PTL[(size_t)i * K + jn], acc);
        val = acc;
    }
}

// ---------------------------------------------------------------------------
// K3 backward: one warp per chain, VPL states per lane, pi^T resident in shared memory so the
// column selected by z_{t+1} is a contiguous row; filtered rows arrive through a cp.async ring
// and the uniforms (tape or pre-generated Philox draws) are fetched 32 steps at a time.
// ---------------------------------------------------------------------------
template <typename R>
__global__ void fill_uniform_kernel(R* __restrict__ u, long long count, SeedArg seed, uint32_t stream) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= count) return;
    Philox g(seed, stream, (uint64_t)e);
    double u0, u1;
    philox_uniform2(g, u0, u1);
    u[e] = (R)u0;
}

// ---------------------------------------------------------------------------
// K3 backward: z_{Tp-1} ~ Cat(filt_{Tp-1}),  z_t ~ Cat(filt_t * pi[:, z_{t+1}]) by inverse CDF,
//   z = #{ i : c_i < (1 - u_t) c_{K-1} },  c_i = sum_{i' <= i} filt_t[i'] pi[i'][z_{t+1}].
// One warp per (chain, time chunk); O(K) work per step.
//
// Walker.  Syllables persist, so the walker tests the hypothesis "the label stays j" for 8
// steps at once: lane group g (4 lanes) takes step t-g and forms A = c_{j-1}, B = c_j and the total
// from its quarter of the row (pi[:, j] lives in registers, once plain and once masked to i < j);
// the label stays iff A < r <= B.  The longest run of passing steps is accepted in one go; the
// first failing step is resolved by a full inverse CDF (running sums piece by piece inside every
// 4-lane group) and the walk continues with the new label.  Rows of filt and the uniforms arrive
// through a per-warp cp.async ring, four 8-step groups in flight; pi^T is resident in shared memory.
//
// Exact time parallelism.  With the uniforms fixed, backward sampling is a deterministic map
// z_{t+1} -> z_t, and two paths that meet stay together.  Chunk c (steps [begin, end)) starts W
// steps above its range from a draw of the filtered marginal alone, walks down without storing
// until it reaches `end` - the label it holds there (zwarm) is the one it conditions on - and then
// stores its range.  Chunks cut the unmasked prefix only: the top chunk starts from the terminal
// draw and takes the whole padded tail (paths do not merge where the filter carries no
// information, and the walker crosses such stretches at a few tens of cycles per step).  A repair
// pass (one warp per chain, top down) compares every zwarm with the label the chunk above actually
// produced; on a mismatch it re-walks the chunk below from the true label until the new path meets
// the stored one.  The result is the sequential sampler's path whatever the merging time;
// mismatches only cost time (diag word 2 = mismatched boundaries, word 3 = re-walked steps).
// ---------------------------------------------------------------------------
template <typename R>
__global__ void transpose_pi_kernel(const R* __restrict__ pi, int K, int ldK, R* __restrict__ piT) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= K * ldK) return;
    const int jn = idx / ldK, i = idx % ldK;      // piT[jn][i] = pi[i][jn]
    piT[idx] = (i < K) ? pi[(size_t)i * K + jn] : (R)0;
}

constexpr int HMM_BW_WARPS = 4;      // warps per CTA (one CTA per SM: pi^T and the rings fill shared memory)
[[maybe_unused]] constexpr int HMM_BW_RING = 32;      // ring slots per warp = four groups of 8 steps

// NPC = 16-byte pieces of a row per lane of a 4-lane group: ceil(ldK * sizeof(R) / 64)
template <typename R, int NPC>
__global__ void __launch_bounds__(32 * HMM_BW_WARPS, 1)
hmm_backward_walk_kernel(const R* __restrict__ filt, const R* __restrict__ piT, const R* __restrict__ u, int N,
                         int K, int Tp, int ldK, int Cb, int Wm, const int* __restrict__ vb, int repair,
                         int* __restrict__ z, int* __restrict__ zwarm, unsigned* __restrict__ diag) {
    typedef typename Vec16<R>::type VecT;
    constexpr int VEC = 16 / (int)sizeof(R);
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, q = lane & 3;
    const int slot_len = ldK + VEC;                              // row + uniform, 16-byte multiple
    R* piTs = reinterpret_cast<R*>(smem_raw);                    // K x ldK, piTs[j][i] = pi[i][j]
    for (int idx = threadIdx.x; idx < K * ldK / VEC; idx += blockDim.x)
        reinterpret_cast<VecT*>(piTs)[idx] = __ldg(reinterpret_cast<const VecT*>(piT) + idx);
    __syncthreads();
    R* ring = piTs + (size_t)K * ldK + (size_t)warp * HMM_BW_RING * slot_len;
    const long long id = (long long)blockIdx.x * (blockDim.x >> 5) + warp;
    const int nn = repair ? (int)id : (int)(id / Cb);
    if (nn >= N) return;
    const int v = vb ? vb[nn] : Tp;
    const R* fl = filt + (size_t)nn * Tp * ldK;
    const R* un = u + (size_t)nn * Tp;
    int* zn = z + (size_t)nn * Tp;
    const int pieces = ldK / VEC;                                // 16-byte pieces per row

    // Walk from step t_hi down to t_lo.  init < 0: the first step is drawn from the filtered marginal
    // alone; otherwise init is the label at t_hi + 1.  Labels of steps < store_hi are stored, the label
    // at store_hi goes to *warm_out.  merge: stop as soon as the new label equals the stored one.
    // Returns the number of steps whose stored label changed (merge mode) / was written.
    auto walk = [&](int t_hi, int t_lo, int init, int store_hi, int* warm_out, bool merge) -> unsigned {
        unsigned written = 0;
        int next_group = 0;
        auto issue_group = [&](int k) {                          // rows t_hi - 8k - (0..7), clipped at t_lo
            const int t = t_hi - 8 * k - g;                      // lane group g brings row g of the group
            if (t >= t_lo) {
                R* dst = ring + (size_t)((8 * k + g) & (HMM_BW_RING - 1)) * slot_len;
                const R* src = fl + (size_t)t * ldK;
#pragma unroll
                for (int r = 0; r < NPC; ++r) {
                    const int pc = q + 4 * r;
                    if (pc < pieces) cp_async_16(dst + pc * VEC, src + pc * VEC);
                }
                if (q == 0) cp_async_elem(dst + ldK, un + t);
            }
            asm volatile("cp.async.commit_group;\n" ::);
        };
        __syncwarp();
        for (; next_group < 4; ++next_group) issue_group(next_group);
        int t = t_hi, j = init;
        R pa[NPC][VEC], pl[NPC][VEC], pjj = (R)0;
        auto load_column = [&](int jj) {
            const R* row = piTs + (size_t)jj * ldK;
#pragma unroll
            for (int r = 0; r < NPC; ++r) {
                const int pc = q + 4 * r;
                if (pc < pieces) *reinterpret_cast<VecT*>(pa[r]) = *reinterpret_cast<const VecT*>(row + pc * VEC);
                else {
#pragma unroll
                    for (int e = 0; e < VEC; ++e) pa[r][e] = (R)0;
                }
#pragma unroll
                for (int e = 0; e < VEC; ++e) pl[r][e] = (pc * VEC + e < jj) ? pa[r][e] : (R)0;
            }
            pjj = row[jj];
        };
        // Inverse CDF at step tt, every 4-lane group redundantly in its own layout (piece q + 4r of
        // the row per lane): running sums piece by piece, left to right.  with_pi: weights pa (the
        // column of the current label), else the filtered marginal alone.  `pl` is scratch here
        // (it is reloaded with the new label's column afterwards).
        auto full_draw = [&](int tt, bool with_pi) -> int {
            const R* row = ring + (size_t)((t_hi - tt) & (HMM_BW_RING - 1)) * slot_len;
            R run = (R)0;
#pragma unroll
            for (int r = 0; r < NPC; ++r) {
                const int pc = q + 4 * r;
                R pv[VEC];
                if (pc < pieces) {
                    const VecT fv = *reinterpret_cast<const VecT*>(row + pc * VEC);
                    const R* fe = reinterpret_cast<const R*>(&fv);
#pragma unroll
                    for (int e = 0; e < VEC; ++e) pv[e] = with_pi ? fe[e] * pa[r][e] : fe[e];
                } else {
#pragma unroll
                    for (int e = 0; e < VEC; ++e) pv[e] = (R)0;
                }
#pragma unroll
                for (int e = 1; e < VEC; ++e) pv[e] += pv[e - 1];
                const R pt = pv[VEC - 1];
                R s1 = __shfl_up_sync(0xffffffffu, pt, 1);
                s1 = q >= 1 ? pt + s1 : pt;
                R s2 = __shfl_up_sync(0xffffffffu, s1, 2);
                s2 = q >= 2 ? s1 + s2 : s1;
                R ex = __shfl_up_sync(0xffffffffu, s2, 1);
                ex = q >= 1 ? ex : (R)0;
                const R tot = __shfl_sync(0xffffffffu, s2, (lane & ~3) | 3);
                const R base = run + ex;
#pragma unroll
                for (int e = 0; e < VEC; ++e) pl[r][e] = base + pv[e];
                run += tot;
            }
            const R thr = run * ((R)1 - row[ldK]);
            int c = 0;
#pragma unroll
            for (int r = 0; r < NPC; ++r)
#pragma unroll
                for (int e = 0; e < VEC; ++e) c += (q + 4 * r < pieces && pl[r][e] < thr) ? 1 : 0;
            c += __shfl_xor_sync(0xffffffffu, c, 1);
            c += __shfl_xor_sync(0xffffffffu, c, 2);
            return min(c, K - 1);
        };
        auto commit_label = [&](int tt, int lab) -> bool {       // true = merged with the stored path
            if (merge && zn[tt] == lab) return true;
            __syncwarp();
            if (lane == 0) {
                if (tt < store_hi) zn[tt] = lab;
                else if (tt == store_hi && warm_out) *warm_out = lab;
            }
            ++written;
            return false;
        };
        auto advance_groups = [&]() {                            // keep four groups in flight below t
            const int kc = (t_hi - t) >> 3;
            if (next_group < kc + 4) {
                __syncwarp();
                for (; next_group < kc + 4; ++next_group) issue_group(next_group);
            }
        };
        if (init < 0) {
            asm volatile("cp.async.wait_group 2;\n" ::);
            __syncwarp();
            j = full_draw(t, false);
            if (commit_label(t, j)) return written;
            --t;
        }
        if (t >= t_lo) load_column(j);
        while (t >= t_lo) {
            advance_groups();
            asm volatile("cp.async.wait_group 2;\n" ::);
            __syncwarp();
            const int s = t - g;
            const bool in_range = s >= t_lo;
            const R* row = ring + (size_t)((t_hi - s) & (HMM_BW_RING - 1)) * slot_len;
            R all0 = 0, all1 = 0, lo0 = 0, lo1 = 0;
            if (in_range) {
#pragma unroll
                for (int r = 0; r < NPC; ++r) {
                    const int pc = q + 4 * r;
                    if (pc < pieces) {
                        const VecT fv = *reinterpret_cast<const VecT*>(row + pc * VEC);
                        const R* fe = reinterpret_cast<const R*>(&fv);
#pragma unroll
                        for (int e = 0; e < VEC; ++e) {
                            if ((r + e) & 1) { all1 = fma(fe[e], pa[r][e], all1); lo1 = fma(fe[e], pl[r][e], lo1); }
                            else { all0 = fma(fe[e], pa[r][e], all0); lo0 = fma(fe[e], pl[r][e], lo0); }
                        }
                    }
                }
            }
            R all = all0 + all1, lo = lo0 + lo1;
            all += __shfl_xor_sync(0xffffffffu, all, 1);
            lo += __shfl_xor_sync(0xffffffffu, lo, 1);
            all += __shfl_xor_sync(0xffffffffu, all, 2);
            lo += __shfl_xor_sync(0xffffffffu, lo, 2);
            bool fail = false;
            int old = -1;
            if (in_range) {
                const R rr = all * ((R)1 - row[ldK]);
                const R hi = lo + row[j] * pjj;
                fail = !((lo < rr) && !(hi < rr));
                if (merge) old = zn[s];
            }
            const unsigned fmask = __ballot_sync(0xffffffffu, fail);
            const int first = fmask ? ((__ffs(fmask) - 1) >> 2) : 8;
            const int nacc = min(first, min(8, t - t_lo + 1));
            if (merge) {                                         // first accepted step whose stored label is already j
                const unsigned mm = __ballot_sync(0xffffffffu, in_range && g < nacc && old == j);
                if (mm) {
                    const int gm = (__ffs(mm) - 1) >> 2;
                    if (q == 0 && g < gm) zn[s] = j;
                    return written + (unsigned)gm;
                }
            }
            if (q == 0 && g < nacc) {
                if (s < store_hi) zn[s] = j;
                else if (s == store_hi && warm_out) *warm_out = j;
            }
            written += (unsigned)nacc;
            t -= nacc;
            if (first < 8 && t >= t_lo && nacc == first) {       // the label changes at step t
                const int jn = full_draw(t, true);
                if (commit_label(t, jn)) return written;
                j = jn;
                --t;
                if (t >= t_lo) load_column(j);
            }
        }
        return written;
    };

    if (!repair) {
        const int c = (int)(id % Cb);
        const ChunkRange cr = chunk_range(v, Tp, Cb, Wm, c);
        if (cr.empty || cr.begin >= cr.end) return;
        const bool top = cr.end >= Tp;
        const int t0 = top ? Tp - 1 : min(cr.end - 1 + max(Wm, 1), Tp - 1);
        walk(t0, cr.begin, -1, cr.end, top ? nullptr : zwarm + (size_t)nn * Cb + c, false);
        return;
    }
    // repair: boundaries top down
    int Cn = 0;
    for (int c = 0; c < Cb; ++c) {
        const ChunkRange cr = chunk_range(v, Tp, Cb, Wm, c);
        if (!cr.empty && cr.begin < cr.end) Cn = c + 1;
    }
    unsigned mism = 0, steps = 0;
    for (int c = Cn - 2; c >= 0; --c) {
        const ChunkRange cr = chunk_range(v, Tp, Cb, Wm, c);
        __syncwarp();
        const int zc = zn[cr.end];
        if (zwarm[(size_t)nn * Cb + c] == zc) continue;
        ++mism;
        steps += walk(cr.end - 1, cr.begin, zc, Tp, nullptr, true);
        asm volatile("cp.async.wait_group 0;\n" ::);
    }
    if (lane == 0 && mism) { atomicAdd(&diag[2], mism); atomicAdd(&diag[3], steps); }
}

// ---------------------------------------------------------------------------
// smoothed marginals (stateseq_marginals): backward recursion over the stored filter
// ---------------------------------------------------------------------------
template <typename R>
__global__ void hmm_smooth_kernel(const R* __restrict__ filt, const R* __restrict__ pi, int K, int Tp,
                                  int ldK, R* __restrict__ marg) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    R* pis = reinterpret_cast<R*>(smem_raw);           // K*K
    R* sm = pis + (size_t)K * K;                       // K smoothed at t+1
    R* ratio = sm + K;                                 // K
    R* fcur = ratio + K;                               // K
    __shared__ R red[32];
    const int nn = blockIdx.x, tid = threadIdx.x;
    for (int i = tid; i < K * K; i += blockDim.x) pis[i] = pi[i];
    const R* fl = filt + (size_t)nn * Tp * ldK;
    R* mg = marg + (size_t)nn * Tp * K;
    for (int i = tid; i < K; i += blockDim.x) { R v = fl[(size_t)(Tp - 1) * ldK + i]; sm[i] = v; mg[(size_t)(Tp - 1) * K + i] = v; }
    __syncthreads();
    for (int t = Tp - 2; t >= 0; --t) {
        for (int i = tid; i < K; i += blockDim.x) fcur[i] = fl[(size_t)t * ldK + i];
        __syncthreads();
        for (int jn = tid; jn < K; jn += blockDim.x) {
            R pred = 0;
            for (int i = 0; i < K; ++i) pred = fma(fcur[i], pis[i * K + jn], pred);
            ratio[jn] = pred > (R)0 ? sm[jn] / pred : (R)0;
        }
        __syncthreads();
        R part = 0, mine = 0;
        if (tid < K) {
            R acc = 0;
            for (int jn = 0; jn < K; ++jn) acc = fma(pis[tid * K + jn], ratio[jn], acc);
            mine = fcur[tid] * acc;
            part = mine;
        }
        R tot = block_sum(part, red);
        if (tid < K) { R v = mine / tot; sm[tid] = v; mg[(size_t)t * K + tid] = v; }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------
// host launchers
// ---------------------------------------------------------------------------
static inline int fp_of(int n, int d, size_t esz) { int F = n + d + 1; int v = 16 / (int)esz; return (F + v - 1) / v * v; }

// workspace shared by the three HMM entry points of one call sequence
enum { HW_DIAG, HW_G, HW_GF, HW_CST, HW_VLEN, HW_DIRTY, HW_BW, HW_BE, HW_LZP, HW_TS, HW_PW, HW_PIT, HW_ZB, HW_END };

// float64 path: state tiles of 8 columns (one warp each) for the tensor-pipe kernels; 0 = unsupported
// (K <= 128: one of four compiled tile counts; up to 512 states: run-time tile count of the wide kernels, hmm_wide.cuh)
static inline int state_tiles(int K) {
    return K <= 32 ? 4 : K <= 56 ? 7 : K <= 104 ? 13 : K <= 128 ? 16 : K <= 512 ? (K + 7) / 8 : 0;
}
static inline bool hmm_wide(int K) { return K > 128; }
static int hmm_check_states(const char* who, int K, size_t esz) {
    if (K < 1 || K > 512) return set_error(-3, "%s: num_states %d outside 1..512", who, K);
    if (hmm_wide(K) && esz != 8)
        return set_error(-3, "%s: num_states %d > 128 needs the float64 discrete-state path (hmm_dtype=float64)", who, K);
    return 0;
}
constexpr int HMM_MT = 1;           // 8-task tiles per CTA of the float64 forward kernel

static int hmm_chunks(int N, int Tp, bool f64) {
    // forward-filter time chunks: multiples of 8 steps (vector loads of the weight rows)
    const int W = (chunk_config().warmup + 7) / 8 * 8;
    return chunks_for(N, f64 ? KPMS_SM_COUNT * 8 * HMM_MT : KPMS_SM_COUNT * 4, Tp, W);
}

template <typename R>
static void hmm_ws_layout(int N, int T, int K, int d, int L, size_t off[HW_END + 1]) {
    const int Fp = fp_of(d * L, d, sizeof(R));
    const int Tp = T - L, C = hmm_chunks(N, Tp, sizeof(R) == 8);
    const int CT = (Tp + HMM_TL - 1) / HMM_TL;
    size_t sz[HW_END] = {256,
                         (size_t)K * d * Fp * sizeof(R),
                         (size_t)state_tiles(K) * (d * ((d * L + d + 3) / 4) * 32 + d * 8 + 8) * sizeof(double),
                         (size_t)K * sizeof(R),
                         (size_t)N * 8,
                         (size_t)N * 8,
                         (size_t)N * (C + 1) * K * sizeof(R),
                         (size_t)N * (C + 1) * K * sizeof(R),
                         (size_t)N * (C + CT) * sizeof(double),
                         (size_t)N * CT * K * sizeof(R),
                         (size_t)2 * K * K * sizeof(R),
                         (size_t)K * ((K + 3) / 4 * 4) * sizeof(R),
                         (size_t)N * KPMS_MAX_CHUNKS * 4};
    off[0] = 0;
    for (int i = 0; i < HW_END; ++i) off[i + 1] = off[i] + align_up(sz[i], 256);
}

template <typename R, int D_, int L_>
static int ar_loglik_launch(const R* x, const int* mask, const R* Ab, const R* Q, int N, int T, int K,
                            int ldT, R* W, R* mx, void* ws, cudaStream_t st) {
    constexpr int n = D_ * L_;
    constexpr int FPT = sizeof(R) == 4 ? 2 : 1;
    constexpr int KC = sizeof(R) == 4 ? 32 : 16;
    if (int rc = hmm_check_states("ar_loglik", K, sizeof(R))) return rc;
    int Fp = fp_of(n, D_, sizeof(R));
    size_t off[HW_END + 1];
    hmm_ws_layout<R>(N, T, K, D_, L_, off);
    R* G = reinterpret_cast<R*>(reinterpret_cast<char*>(ws) + off[HW_G]);
    R* cst = reinterpret_cast<R*>(reinterpret_cast<char*>(ws) + off[HW_CST]);
    { KPMS_LAUNCH("ar_prep", st);
    ar_prep_kernel<R, D_, L_><<<K, 32, 0, st>>>(Ab, Q, K, G, cst, Fp); }
    {   // unmasked prefix of every chain (time chunks of the forward filter)
        int* vb = reinterpret_cast<int*>(reinterpret_cast<char*>(ws) + off[HW_VLEN]);
        KPMS_LAUNCH("hmm_prefix", st);
        hmm_prefix_kernel<<<N, 256, 0, st>>>(mask, T, L_, T - L_, vb, vb + N);
    }
    if constexpr (sizeof(R) == 8) {
        // tensor-pipe path: operators repacked in fragment order, W (N, Tp, 8*KT) states-contiguous
        typedef ArFrag<D_, L_> AF;
        const int KT = state_tiles(K);
        double* Gf = reinterpret_cast<double*>(reinterpret_cast<char*>(ws) + off[HW_GF]);
        { KPMS_LAUNCH("ar_pack_frag", st);
          ar_pack_frag_kernel<D_, L_><<<ceil_div(KT * AF::CHUNK, 256), 256, 0, st>>>((const double*)G, (const double*)cst, K, Fp, KT, Gf); }
        constexpr int FRD = 8 * AR_WARPS;
        const size_t smem = (align_up((size_t)(FRD + L_) * D_, 2) + (size_t)2 * AF::CHUNK) * sizeof(double);
        dim3 grid(ceil_div(T - L_, FRD), N);
#define ARL(KT_)                                                                                              \
        {                                                                                                     \
            auto kern = ar_loglik_dmma_kernel<D_, L_, KT_>;                                                   \
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);               \
            KPMS_LAUNCH("ar_loglik", st);                                                                     \
            kern<<<grid, 32 * AR_WARPS, smem, st>>>((const double*)x, mask, Gf, N, T, K, ldT, (double*)W, (double*)mx); \
        }
        if (hmm_wide(K)) {
            auto kern = ar_loglik_dmma_wide_kernel<D_, L_>;
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            KPMS_LAUNCH("ar_loglik", st);
            kern<<<grid, 32 * AR_WARPS, smem, st>>>((const double*)x, mask, Gf, N, T, K, KT, ldT, (double*)W, (double*)mx);
        }
        else if (KT == 4) ARL(4) else if (KT == 7) ARL(7) else if (KT == 13) ARL(13) else ARL(16)
#undef ARL
        return check_launch("ar_loglik");
    }
    constexpr int FR = 128 * FPT;
    size_t smem = (align_up((size_t)(FR + L_) * D_, 4) + (size_t)KC * D_ * Fp + KC) * sizeof(R);
    auto kern = ar_loglik_kernel<R, D_, L_, FPT, KC>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    dim3 grid(ceil_div(T - L_, FR), N);
    { KPMS_LAUNCH("ar_loglik", st); kern<<<grid, 128, smem, st>>>(x, mask, G, cst, N, T, K, Fp, ldT, W, mx); }
    return check_launch("ar_loglik");
}

// (latent_dim, nlags) pairs of this translation unit's group (common.cuh: KPMS_DL_GROUP_g, -DKPMS_DL_GROUP=g)
template <typename R>
static int ar_loglik_group_impl(const void* x, const int* mask, const void* Ab, const void* Q, int N, int T, int d,
                                int L, int K, int ldT, void* W, void* mx, void* ws, cudaStream_t st) {
#define X(DD, LL)                                                                                         \
    if (d == DD && L == LL)                                                                               \
        return ar_loglik_launch<R, DD, LL>((const R*)x, mask, (const R*)Ab, (const R*)Q, N, T, K, ldT,    \
                                           (R*)W, (R*)mx, ws, st);
    KPMS_FOR_GROUP_DL(X)
#undef X
    return KPMS_NOT_IN_GROUP;
}

#define KPMS_ARLL_GROUP_ARGS                                                                                      \
    int dtype, const void *x, const int *mask, const void *Ab, const void *Q, int N, int T, int d, int L, int K,  \
        int ldT, void *W, void *mx, void *ws, cudaStream_t st

int KPMS_CAT(ar_loglik_group_, KPMS_DL_GROUP)(KPMS_ARLL_GROUP_ARGS) {
    return KPMS_DISPATCH_DTYPE(dtype, ar_loglik_group_impl, x, mask, Ab, Q, N, T, d, L, K, ldT, W, mx, ws, st);
}

#if KPMS_DL_GROUP == 0
int ar_loglik_group_1(KPMS_ARLL_GROUP_ARGS);
int ar_loglik_group_2(KPMS_ARLL_GROUP_ARGS);
int ar_loglik_group_3(KPMS_ARLL_GROUP_ARGS);
int ar_loglik_group_4(KPMS_ARLL_GROUP_ARGS);
int ar_loglik_group_5(KPMS_ARLL_GROUP_ARGS);
int ar_loglik_group_6(KPMS_ARLL_GROUP_ARGS);
int ar_loglik_group_7(KPMS_ARLL_GROUP_ARGS);
static_assert(KPMS_DL_GROUPS == 8, "one dispatcher per group");

static int ar_loglik_dispatch(KPMS_ARLL_GROUP_ARGS) {
    if (dtype != 0 && dtype != 1) return set_error(-2, "dtype must be 0 (f32) or 1 (f64), got %d", dtype);
    if (T <= L) return set_error(-3, "ar_loglik: T (%d) must exceed nlags (%d)", T, L);
    if (ldT < T - L || ldT % 8) return set_error(-3, "ar_loglik: ldT (%d) must be a multiple of 8 and >= T-L", ldT);
    typedef int (*GroupFn)(KPMS_ARLL_GROUP_ARGS);
    static const GroupFn groups[KPMS_DL_GROUPS] = {ar_loglik_group_0, ar_loglik_group_1, ar_loglik_group_2,
                                                   ar_loglik_group_3, ar_loglik_group_4, ar_loglik_group_5,
                                                   ar_loglik_group_6, ar_loglik_group_7};
    for (int g = 0; g < KPMS_DL_GROUPS; ++g) {
        const int rc = groups[g](dtype, x, mask, Ab, Q, N, T, d, L, K, ldT, W, mx, ws, st);
        if (rc != KPMS_NOT_IN_GROUP) return rc;
    }
    return set_error(-3, "ar_loglik: unsupported (latent_dim, nlags) = (%d, %d); see kpms_supported_dims", d, L);
}

template <typename R>
static int hmm_forward_impl(const void* W, const void* mx, const void* pi, int N, int K, int Tp, int ldT,
                            void* filt, double* logZ, void* ws, int d, int L, cudaStream_t st) {
    const int ldK = (K + 3) / 4 * 4;
    const int Kpad = (K + 7) / 8 * 8;
    if (int rc = hmm_check_states("hmm_forward", K, sizeof(R))) return rc;
    constexpr int M = sizeof(R) == 8 ? 8 * HMM_MT : 4;       // tasks per CTA
    static_assert(8 * HMM_MT == HMM_WIDE_M, "the wide filter shares the task grid of the tensor-pipe filter");
    const bool wide = hmm_wide(K);
    const int Kw = 8 * state_tiles(K);                       // row stride of W (float64 layout)
    const int wide_threads = (K + 31) / 32 * 32;
    const size_t wide_smem = ((size_t)2 * Kw * HMM_WIDE_M + (size_t)2 * HMM_WIDE_M * (wide_threads / 32)) * sizeof(double);
    if (wide) cudaFuncSetAttribute(hmm_forward_wide_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wide_smem);
    size_t off[HW_END + 1];
    hmm_ws_layout<R>(N, Tp + L, K, d, L, off);
    char* base = reinterpret_cast<char*>(ws);
    unsigned* diag = reinterpret_cast<unsigned*>(base + off[HW_DIAG]);
    int* vb = reinterpret_cast<int*>(base + off[HW_VLEN]);
    int* holes = vb + N;
    int* dirty = reinterpret_cast<int*>(base + off[HW_DIRTY]);
    R* bw = reinterpret_cast<R*>(base + off[HW_BW]);
    R* be = reinterpret_cast<R*>(base + off[HW_BE]);
    double* lzp = reinterpret_cast<double*>(base + off[HW_LZP]);
    R* tstart = reinterpret_cast<R*>(base + off[HW_TS]);
    R* pw = reinterpret_cast<R*>(base + off[HW_PW]);
    const ChunkConfig cfg = chunk_config();
    static const int cfg_refine = [] {
        const char* e = getenv("KPMS_HMM_REFINE");
        const int v = e ? atoi(e) : 3;                           // default: three passes (0 switches them off)
        return v < 0 ? 0 : v > 8 ? 8 : v;
    }();
    const int C = hmm_chunks(N, Tp, sizeof(R) == 8), CT = (Tp + HMM_TL - 1) / HMM_TL, Wm = (cfg.warmup + 7) / 8 * 8;
    const dim3 block(4 * Kpad);
#define FWD(RPT, GRID, PASS, VB, DIRTY)                                                                       \
    {                                                                                                         \
        auto kern = hmm_forward_kernel<R, RPT, M>;                                                            \
        constexpr int RPTP_ = (RPT + 16 / (int)sizeof(R) - 1) / (16 / (int)sizeof(R)) * (16 / (int)sizeof(R)); \
        const size_t smem = ((size_t)2 * M * 4 * RPT * 8 + (size_t)2 * M * 4 * RPTP_) * sizeof(R);            \
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);                   \
        kern<<<GRID, block, smem, st>>>((const R*)W, (const R*)mx, (const R*)pi, N, K, Tp, ldT, ldK, (R*)filt, \
                                        logZ, lzp, PASS, C, CT, Wm, VB, DIRTY, bw, be, tstart);               \
    }
#define FWD64(KT_, GRID, PASS, VB, DIRTY)                                                                     \
    hmm_forward_dmma_kernel<KT_, HMM_MT><<<GRID, 32 * KT_, 0, st>>>(                                          \
        (const double*)W, (const double*)mx, (const double*)pi, N, K, Tp, ldT, ldK, (double*)filt, logZ, lzp, \
        PASS, C, CT, Wm, VB, DIRTY, (double*)bw, (PASS) == 3 ? (double*)nullptr : (double*)be,                \
        (PASS) == 3 ? (const double*)be : (const double*)tstart);
#define FWDW(GRID, PASS, VB, DIRTY)                                                                           \
    hmm_forward_wide_kernel<<<GRID, wide_threads, wide_smem, st>>>(                                           \
        (const double*)W, (const double*)mx, (const double*)pi, N, K, Kw, Tp, ldT, ldK, (double*)filt, logZ,  \
        lzp, PASS, C, CT, Wm, VB, DIRTY, (double*)bw, (PASS) == 3 ? (double*)nullptr : (double*)be,           \
        (PASS) == 3 ? (const double*)be : (const double*)tstart);
#define FWD_K(GRID, PASS, VB, DIRTY)                                             \
    if (wide) FWDW(GRID, PASS, VB, DIRTY)                                        \
    else if (sizeof(R) == 8) {                                                   \
        if (K <= 32) FWD64(4, GRID, PASS, VB, DIRTY)                             \
        else if (K <= 56) FWD64(7, GRID, PASS, VB, DIRTY)                        \
        else if (K <= 104) FWD64(13, GRID, PASS, VB, DIRTY)                      \
        else FWD64(16, GRID, PASS, VB, DIRTY)                                    \
    }                                                                            \
    else if (K <= 28) FWD(7, GRID, PASS, VB, DIRTY)                              \
    else if (K <= 52) FWD(13, GRID, PASS, VB, DIRTY)                             \
    else if (K <= 100) FWD(25, GRID, PASS, VB, DIRTY)                            \
    else FWD(32, GRID, PASS, VB, DIRTY)
    cudaMemsetAsync(diag, 0, 256, st);
    if (C <= 1) {                                        // sequential: every chain is one task
        KPMS_LAUNCH("hmm_forward", st);
        FWD_K((N + M - 1) / M, 2, (const int*)nullptr, (const int*)nullptr)
        return check_launch("hmm_forward");
    }
    const R tol = (R)(sizeof(R) == 4 ? cfg.tol32 : cfg.tol_hmm);
    cudaMemsetAsync(lzp, 0, (size_t)N * (C + CT) * sizeof(double), st);
    { KPMS_LAUNCH("hmm_forward", st); FWD_K((int)(((long long)N * C + M - 1) / M), 0, vb, (const int*)nullptr) }
    { KPMS_LAUNCH("hmm_forward_check", st);
      cudaMemsetAsync(dirty, 0, (size_t)N * sizeof(int), st);
      boundary_check_kernel<R, true><<<dim3(C - 1, N), 128, 0, st>>>(bw, be, vb, Tp, C, Wm, 8, K, K, tol, dirty, diag, holes); }
    if (sizeof(R) == 8 && cfg_refine > 0) {
        // Refinement passes (KPMS_HMM_REFINE=<passes>, default 3): chains whose warm-up did not forget are not
        // handed to the sequential kernel at once; their chunks restart from the predecessor's last end state,
        // the boundary error contracts by the chunk-long forgetting factor per pass, and the same
        // component-wise check certifies convergence.  Chains still flagged afterwards fall back as before.
        int* dirty2 = dirty + N;
        int* din = dirty;
        int* dout = dirty2;
        cudaMemsetAsync(diag + 4, cfg_refine, 1, st);            // low byte of word 4 = passes enabled
        for (int r = 0; r < cfg_refine; ++r) {
            cudaMemsetAsync(dout, 0, (size_t)N * sizeof(int), st);
            cudaMemsetAsync(diag + 5, 0, sizeof(unsigned), st);
            { KPMS_LAUNCH("hmm_forward_refine", st);
              FWD_K((int)(((long long)N * C + M - 1) / M), 3, vb, din) }
            { KPMS_LAUNCH("hmm_forward_check", st);
              hmm_refine_check_kernel<R><<<dim3(C, N), 128, 0, st>>>(be, bw, vb, N, Tp, C, Wm, K, tol, din, dout, diag + 5); }
            int* tmp = din; din = dout; dout = tmp;
        }
        dirty = din;
    }
    // pi^TL by repeated squaring, then the starting predictions of the padded-tail chunks
    {
        const R* src = (const R*)pi;
        int which = 0;
        for (int e = 1; e < HMM_TL; e <<= 1) {
            KPMS_LAUNCH("hmm_pi_power", st);
            mat_square_kernel<R><<<ceil_div(K * K, 128), 128, 0, st>>>(src, K, pw + (size_t)which * K * K);
            src = pw + (size_t)which * K * K;
            which ^= 1;
        }
        KPMS_LAUNCH("hmm_tail_starts", st);
        if (wide) hmm_tail_starts_wide_kernel<<<N, HMM_WIDE_MAX, 0, st>>>((const double*)src, (const double*)be, vb, dirty, K, Tp, C, CT, Wm, (double*)tstart);
        else hmm_tail_starts_kernel<R><<<N, 128, 0, st>>>(src, be, vb, dirty, K, Tp, C, CT, Wm, tstart);
    }
    { KPMS_LAUNCH("hmm_forward_tail", st); FWD_K((int)(((long long)N * CT + M - 1) / M), 1, vb, dirty) }
    { KPMS_LAUNCH("hmm_logz_sum", st); logz_sum_kernel<<<ceil_div(N, 128), 128, 0, st>>>(lzp, N, C + CT, logZ); }
    if (wide) {
        KPMS_LAUNCH("hmm_forward_rerun", st); FWDW((N + M - 1) / M, 2, (const int*)nullptr, dirty)
    } else if (sizeof(R) == 8) {
        // Chains whose boundaries failed the check (slow forgetting: parameters far from the data) are
        // re-run sequentially, one chain per CTA on the latency-lean DFMA kernel (a DMMA step costs the
        // same pipe time for one task as for eight).
        KPMS_LAUNCH("hmm_forward_rerun", st);
#define SEQ64(RPT_)                                                                                           \
        {                                                                                                     \
            auto kern = hmm_forward_kernel<double, RPT_, 1, true>;                                            \
            constexpr int KC_ = (4 * RPT_ + 7) / 8 * 8, RPTP_ = (RPT_ + 1) / 2 * 2;                           \
            const size_t smem = ((size_t)2 * KC_ * 8 + (size_t)2 * 4 * RPTP_) * sizeof(double);               \
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);               \
            kern<<<N, 4 * KC_, smem, st>>>((const double*)W, (const double*)mx, (const double*)pi, N, K, Tp,  \
                                           ldT, ldK, (double*)filt, logZ, lzp, 2, C, CT, Wm,                  \
                                           (const int*)nullptr, dirty, (double*)bw, (double*)be,              \
                                           (const double*)tstart);                                            \
        }
        const int KTs = state_tiles(K);
        if (KTs == 4) SEQ64(8) else if (KTs == 7) SEQ64(14) else if (KTs == 13) SEQ64(26) else SEQ64(32)
#undef SEQ64
    } else {
        KPMS_LAUNCH("hmm_forward_rerun", st); FWD_K((N + M - 1) / M, 2, (const int*)nullptr, dirty)
    }
#undef FWD_K
#undef FWDW
#undef FWD64
#undef FWD
    return check_launch("hmm_forward");
}

static int hmm_backward_chunks(int N, int Tp) {
    const int W = chunk_config().warmup;
    return chunks_for(N, KPMS_SM_COUNT * HMM_BW_WARPS, Tp, W);
}

template <typename R>
static int hmm_backward_impl(const void* filt, const void* pi, const void* u, void* u_scratch, SeedArg seed, int N,
                             int K, int Tp, int* z, void* ws, int d, int L, cudaStream_t st) {
    const int ldK = (K + 3) / 4 * 4;
    if (int rc = hmm_check_states("hmm_backward", K, sizeof(R))) return rc;
    size_t off[HW_END + 1];
    hmm_ws_layout<R>(N, Tp + L, K, d, L, off);
    char* base = reinterpret_cast<char*>(ws);
    unsigned* diag = reinterpret_cast<unsigned*>(base + off[HW_DIAG]);
    const int* vb = reinterpret_cast<const int*>(base + off[HW_VLEN]);     // written by kpms_ar_loglik
    int* zwarm = reinterpret_cast<int*>(base + off[HW_ZB]);
    const R* usrc = (const R*)u;
    if (!usrc) {
        if (!u_scratch) return set_error(-3, "hmm_backward: u_scratch (N*Tp reals) is required when no tape is given");
        const long long count = (long long)N * Tp;
        KPMS_LAUNCH("hmm_uniforms", st);
        fill_uniform_kernel<R><<<(int)((count + 255) / 256), 256, 0, st>>>((R*)u_scratch, count, seed, KPMS_STREAM_Z);
        usrc = (const R*)u_scratch;
    }
    const int Cb = hmm_backward_chunks(N, Tp), Wm = chunk_config().warmup;
    R* piT = reinterpret_cast<R*>(base + off[HW_PIT]);
    { KPMS_LAUNCH("hmm_transpose_pi", st);
      transpose_pi_kernel<R><<<ceil_div(K * ldK, 256), 256, 0, st>>>((const R*)pi, K, ldK, piT); }
    cudaMemsetAsync(diag + 2, 0, 8, st);
    if (hmm_wide(K)) {
        constexpr int WPC = 4;
#define BWDW(EPL_)                                                                                            \
        {                                                                                                     \
            { KPMS_LAUNCH("hmm_backward", st);                                                                \
              hmm_backward_wide_kernel<EPL_><<<(int)(((long long)N * Cb + WPC - 1) / WPC), 32 * WPC, 0, st>>>( \
                  (const double*)filt, (const double*)piT, (const double*)usrc, N, K, Tp, ldK, Cb, Wm, vb, 0, z, zwarm, diag); } \
            if (Cb > 1) {                                                                                     \
                KPMS_LAUNCH("hmm_backward_repair", st);                                                       \
                hmm_backward_wide_kernel<EPL_><<<(N + WPC - 1) / WPC, 32 * WPC, 0, st>>>(                     \
                    (const double*)filt, (const double*)piT, (const double*)usrc, N, K, Tp, ldK, Cb, Wm, vb, 1, z, zwarm, diag); \
            }                                                                                                 \
        }
        if (ldK <= 256) BWDW(8) else BWDW(16)
#undef BWDW
        return check_launch("hmm_backward");
    }
    const size_t warp_ring = (size_t)HMM_BW_RING * (ldK + 16 / sizeof(R)) * sizeof(R);
    const size_t pit_bytes = (size_t)K * ldK * sizeof(R);
    const int wpc = pit_bytes + HMM_BW_WARPS * warp_ring <= 220 * 1024 ? HMM_BW_WARPS : HMM_BW_WARPS / 2;   // warps per CTA
    const size_t smem = pit_bytes + wpc * warp_ring;
    const int npc = (int)((ldK * sizeof(R) + 63) / 64);
#define BWD(NPC_)                                                                                             \
    {                                                                                                         \
        auto kern = hmm_backward_walk_kernel<R, NPC_>;                                                        \
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);                   \
        { KPMS_LAUNCH("hmm_backward", st);                                                                    \
          kern<<<(int)(((long long)N * Cb + wpc - 1) / wpc), 32 * wpc, smem, st>>>(                           \
              (const R*)filt, piT, usrc, N, K, Tp, ldK, Cb, Wm, vb, 0, z, zwarm, diag); }                     \
        if (Cb > 1) {                                                                                         \
            KPMS_LAUNCH("hmm_backward_repair", st);                                                           \
            kern<<<(N + wpc - 1) / wpc, 32 * wpc, smem, st>>>(                                                \
                (const R*)filt, piT, usrc, N, K, Tp, ldK, Cb, Wm, vb, 1, z, zwarm, diag);                     \
        }                                                                                                     \
    }
    if (npc <= 2) BWD(2) else if (npc <= 4) BWD(4) else if (npc <= 7) BWD(7) else if (npc <= 8) BWD(8)
    else if (npc <= 13) BWD(13) else BWD(16)
#undef BWD
    return check_launch("hmm_backward");
}

template <typename R>
static int hmm_smooth_impl(const void* filt, const void* pi, int N, int K, int Tp, void* marg, cudaStream_t st) {
    int ldK = (K + 3) / 4 * 4;
    if (int rc = hmm_check_states("hmm_smooth", K, sizeof(R))) return rc;
    if (hmm_wide(K)) {
        KPMS_LAUNCH("hmm_smooth", st);
        hmm_smooth_wide_kernel<<<N, (K + 31) / 32 * 32, 0, st>>>((const double*)filt, (const double*)pi, K, Tp, ldK, (double*)marg);
        return check_launch("hmm_smooth");
    }
    size_t smem = ((size_t)K * K + 3 * K) * sizeof(R);
    auto kern = hmm_smooth_kernel<R>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int threads = (K + 31) / 32 * 32;
    { KPMS_LAUNCH("hmm_smooth", st); kern<<<N, threads, smem, st>>>((const R*)filt, (const R*)pi, K, Tp, ldK, (R*)marg); }
    return check_launch("hmm_smooth");
}

#endif  // KPMS_DL_GROUP == 0

}  // namespace kpms

#if KPMS_DL_GROUP == 0
using namespace kpms;

extern "C" {

size_t kpms_hmm_workspace_bytes(int dtype, int N, int T, int K, int d, int L) {
    size_t off[HW_END + 1];
    if (dtype == 0) hmm_ws_layout<float>(N, T, K, d, L, off);
    else hmm_ws_layout<double>(N, T, K, d, L, off);
    return off[HW_END];
}

size_t kpms_hmm_weights_bytes(int dtype, int N, int T, int K, int L) {
    const size_t Tp = T > L ? T - L : 0, ldT = (Tp + 7) / 8 * 8;
    if (dtype == 0) return (size_t)N * K * ldT * sizeof(float);
    return (size_t)N * Tp * 8 * (state_tiles(K) ? state_tiles(K) : (K + 7) / 8) * sizeof(double);
}

int kpms_ar_loglik(int dtype, const void* x, const int* mask, const void* Ab, const void* Q, int N, int T,
                   int d, int L, int K, int ldT, void* W, void* mx, void* ws, void* stream) {
    return ar_loglik_dispatch(dtype, x, mask, Ab, Q, N, T, d, L, K, ldT, W, mx, ws, (cudaStream_t)stream);
}

int kpms_hmm_forward(int dtype, const void* W, const void* mx, const void* pi, int N, int K, int Tp, int ldT,
                     void* filt, double* logZ, void* ws, int d, int L, void* stream) {
    return KPMS_DISPATCH_DTYPE(dtype, hmm_forward_impl, W, mx, pi, N, K, Tp, ldT, filt, logZ, ws, d, L,
                               (cudaStream_t)stream);
}

int kpms_hmm_backward_sample(int dtype, const void* filt, const void* pi, const void* u_tape, void* u_scratch,
                             uint64_t seed, const uint64_t* seed_dev, int N, int K, int Tp, int* z, void* ws, int d, int L,
                             void* stream) {
    return KPMS_DISPATCH_DTYPE(dtype, hmm_backward_impl, filt, pi, u_tape, u_scratch, SeedArg(seed, seed_dev), N, K, Tp, z, ws, d, L,
                               (cudaStream_t)stream);
}

int kpms_hmm_smooth(int dtype, const void* filt, const void* pi, int N, int K, int Tp, void* marg, void* stream) {
    return KPMS_DISPATCH_DTYPE(dtype, hmm_smooth_impl, filt, pi, N, K, Tp, marg, (cudaStream_t)stream);
}

}  // extern "C"
#endif
