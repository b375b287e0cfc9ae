// AR-HMM path: per-frame AR log-likelihood weights (K2) and HMM forward filter /
// backward sampler (K3).  Replaces jax_moseq.models.arhmm.resample_discrete_stateseqs,
// marginal_log_likelihood and stateseq_marginals, reached from
// keypoint_moseq/fitting.py:25, :536-538, :667-673.
//
// Layouts (row-major):
//   x     (N, T, d)            latent trajectories
//   mask  (N, T) int32
//   W     (N, K, ldT)          W[n][k][t'] = exp(ll[n][t'][k] - mx[n][t']),  t' = t - L
//   mx    (N, ldT)             per-frame max log-likelihood (0 on masked frames)
//   filt  (N, T', ldK)         filtered state probabilities
//   z     (N, T') int32
#include "common.cuh"
#include "../../include/kpms_b200.h"

namespace kpms {

// ---------------------------------------------------------------------------
// per-state whitened regression operator: G_k = Lq_k^{-1} [ -A_k | I | -b_k ]  (d x F),
// F = n + d + 1, features f = [x_{t-L} .. x_{t-1} | x_t | 1];  c_k = -sum log diag Lq - d/2 log 2pi
// ---------------------------------------------------------------------------
template <typename R, int D_, int L_>
__global__ void ar_prep_kernel(const R* __restrict__ Ab, const R* __restrict__ Q, int K,
                               R* __restrict__ G, R* __restrict__ cst, int Fp) {
    constexpr int n = D_ * L_;
    int k = blockIdx.x * blockDim.x + threadIdx.x;
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
    for (int i = 0; i < D_; ++i) {
        for (int j = 0; j <= n; ++j) {
            double v = 0.0;
            for (int p = 0; p <= i; ++p) v += Li[i][p] * (double)A[p * (n + 1) + j];
            if (j < n) g[i * Fp + j] = (R)(-v);
            else g[i * Fp + n + D_] = (R)(-v);
        }
        for (int j = 0; j < D_; ++j) g[i * Fp + n + j] = (R)Li[i][j];
        for (int j = n + D_ + 1; j < Fp; ++j) g[i * Fp + j] = (R)0;
    }
    cst[k] = (R)(-logdet - 0.5 * D_ * 1.8378770664093453);
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
// K3 forward: scaled filter, one CTA per chain, 4 threads per state column with the
// transition-matrix column slices resident in registers; one barrier per step.
// ---------------------------------------------------------------------------
template <typename R, int RPT>
__global__ void __launch_bounds__(4 * 128)
hmm_forward_kernel(const R* __restrict__ W, const R* __restrict__ mx, const R* __restrict__ pi,
                   int K, int Tp, int ldT, int ldK, R* __restrict__ filt, double* __restrict__ logZ,
                   int C, int Wm, const int* __restrict__ vlen, const int* __restrict__ dirty,
                   R* __restrict__ bnd_warm, R* __restrict__ bnd_end, double* __restrict__ logZ_part) {
    constexpr int VEC = 16 / sizeof(R);
    constexpr int RPTP = (RPT + VEC - 1) / VEC * VEC;
    constexpr int CH = 8;                                    // steps per prefetched chunk (ldT % 8 == 0)
    __shared__ __align__(16) R qbuf[2][4 * RPTP];
    __shared__ double red[32];
    const int nn = blockIdx.x, ck = blockIdx.y;
    const int tid = threadIdx.x;
    if (dirty && dirty[nn] == 0) return;
    // time chunk (common.cuh): outputs for [cr.begin, cr.end), recursion from cr.start with the
    // uniform prior; all three are multiples of CH except the chain end
    const ChunkRange cr = chunk_range(vlen ? vlen[nn] : Tp, Tp, C, Wm, ck, 8);
    if (cr.empty) return;
    const int j = tid >> 2, p = tid & 3;
    const bool col = j < K;
    R pic[RPT];
#pragma unroll
    for (int r = 0; r < RPT; ++r) {
        int i = p * RPT + r;
        pic[r] = (col && i < K) ? pi[(size_t)i * K + j] : (R)0;
    }
    for (int i = tid; i < 2 * 4 * RPTP; i += blockDim.x) (&qbuf[0][0])[i] = (R)0;
    // sum of per-frame maxima (part of the log-normaliser)
    double msum = 0.0;
    for (int t = cr.begin + tid; t < cr.end; t += blockDim.x) msum += (double)mx[(size_t)nn * ldT + t];
    msum = block_sum(msum, red);
    const R* Wc = W + ((size_t)nn * K + (col ? j : 0)) * ldT;
    R* fl = filt + (size_t)nn * Tp * ldK;
    const int qslot = (j / RPT) * RPTP + (j % RPT);
    R cur[CH], nxt[CH];
    typedef typename Vec16<R>::type VecT;
    auto load_chunk = [&](int t0, R* dst) {
        if (t0 < ldT) {
#pragma unroll
            for (int v = 0; v < CH / VEC; ++v) {
                const VecT val = *reinterpret_cast<const VecT*>(Wc + t0 + v * VEC);
                const R* ve = reinterpret_cast<const R*>(&val);
#pragma unroll
                for (int c = 0; c < VEC; ++c) dst[v * VEC + c] = ve[c];
            }
        } else {
#pragma unroll
            for (int c = 0; c < CH; ++c) dst[c] = (R)0;
        }
    };
    load_chunk(cr.start, cur);
    R pred = (R)1 / (R)K;
    R inv_s = (R)1;
    R qprev = (R)0;
    double lz = 0.0;
    int buf = 0;
    for (int t0 = cr.start; t0 < cr.end; t0 += CH) {
        load_chunk(t0 + CH, nxt);
#pragma unroll
        for (int c = 0; c < CH; ++c) {
            const int t = t0 + c;
            if (t >= cr.end) break;
            if (ck > 0 && t == cr.begin && col && p == 0)          // the prediction this chunk arrived with
                bnd_warm[((size_t)nn * C + ck) * K + j] = pred * inv_s;
            R qv = pred * inv_s * cur[c];
            if (col && p == 0) {
                qbuf[buf][qslot] = qv;
                if (t > cr.begin) fl[(size_t)(t - 1) * ldK + j] = qprev * inv_s;
            }
            __syncthreads();
            const R* qs = &qbuf[buf][p * RPTP];
            R a0 = 0, a1 = 0, s0 = 0, s1 = 0;
#pragma unroll
            for (int r = 0; r + 1 < RPT; r += 2) {
                R q0 = qs[r], q1 = qs[r + 1];
                a0 = fma(pic[r], q0, a0);
                a1 = fma(pic[r + 1], q1, a1);
                s0 += q0;
                s1 += q1;
            }
            if (RPT & 1) { R q0 = qs[RPT - 1]; a0 = fma(pic[RPT - 1], q0, a0); s0 += q0; }
            R a = a0 + a1, s = s0 + s1;
            a += __shfl_xor_sync(0xffffffffu, a, 1);
            s += __shfl_xor_sync(0xffffffffu, s, 1);
            a += __shfl_xor_sync(0xffffffffu, a, 2);
            s += __shfl_xor_sync(0xffffffffu, s, 2);
            pred = a;
            inv_s = (R)1 / s;
            qprev = qv;
            if (tid == 0 && t >= cr.begin) lz += log((double)s);
            buf ^= 1;
        }
#pragma unroll
        for (int c = 0; c < CH; ++c) cur[c] = nxt[c];
    }
    if (col && p == 0) {
        fl[(size_t)(cr.end - 1) * ldK + j] = qprev * inv_s;
        if (cr.end < Tp) bnd_end[((size_t)nn * C + ck + 1) * K + j] = pred * inv_s;   // handed to the next chunk
    }
    if (tid == 0) {
        if (logZ_part) logZ_part[(size_t)nn * C + ck] = lz + msum;
        else logZ[nn] = lz + msum;
    }
}

// logZ[nn] = ordered sum of the chunks' parts (chains flagged dirty are overwritten by the re-run)
__global__ void logz_sum_kernel(const double* __restrict__ part, const int* __restrict__ vlen, int N, int Tp, int C,
                                int Wm, double* __restrict__ logZ) {
    const int nn = blockIdx.x * blockDim.x + threadIdx.x;
    if (nn >= N) return;
    double acc = 0.0;
    for (int c = 0; c < C; ++c) {
        if (chunk_range(vlen[nn], Tp, C, Wm, c, 8).empty) break;
        acc += part[(size_t)nn * C + c];
    }
    logZ[nn] = acc;
}

// ---------------------------------------------------------------------------
// K3 backward: one warp per chain, VPL states per lane, pi^T resident in shared memory so the
// column selected by z_{t+1} is a contiguous row; filtered rows arrive through a cp.async ring
// and the uniforms (tape or pre-generated Philox draws) are fetched 32 steps at a time.
// ---------------------------------------------------------------------------
template <typename R>
__global__ void fill_uniform_kernel(R* __restrict__ u, long long count, uint64_t seed, uint32_t stream) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= count) return;
    Philox g(seed, stream, (uint64_t)e);
    double u0, u1;
    philox_uniform2(g, u0, u1);
    u[e] = (R)u0;
}

template <typename R, int VPL, int STAGES>
__global__ void __launch_bounds__(32)
hmm_backward_kernel(const R* __restrict__ filt, const R* __restrict__ piT, const R* __restrict__ u_src,
                    int K, int Tp, int ldK, int* __restrict__ z, int C, int Wm, const int* __restrict__ vlen,
                    const int* __restrict__ dirty, int* __restrict__ bz_warm, int* __restrict__ bz_exact) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    R* pis = reinterpret_cast<R*>(smem_raw);                 // K x ldK
    R* ring = pis + (size_t)K * ldK;                         // STAGES x ldK
    const int nn = blockIdx.x, ck = blockIdx.y;
    const int lane = threadIdx.x;
    if (dirty && dirty[nn] == 0) return;
    // Time chunk: labels for [cr.begin, cr.end).  The last chunk starts from the chain's terminal
    // draw; the others start Wm steps above their range from an unconditioned draw and, because
    // every step is the same deterministic map of (z_{t+1}, u_t), merge with the sequential
    // sampler's path as soon as the two agree once.  The label used at the upper boundary is
    // published and compared with the neighbour's own label (exact integer check).
    const ChunkRange cr = chunk_range(vlen ? vlen[nn] : Tp, Tp, C, Wm, ck, 8);
    if (cr.empty) return;
    const int top = (cr.end == Tp) ? Tp : min(cr.end + Wm, Tp);   // first step taken is t = top - 1
    for (int i = lane; i < K * ldK; i += 32) pis[i] = piT[i];
    const R* fl = filt + (size_t)nn * Tp * ldK;
    const R* un = u_src + (size_t)nn * Tp;
    int* zn = z + (size_t)nn * Tp;
    const int chunks = ldK * (int)sizeof(R) / 16;
    auto issue = [&](int t) {
        if (t >= cr.begin) {
            char* dst = reinterpret_cast<char*>(ring + (size_t)(t % STAGES) * ldK);
            const char* src = reinterpret_cast<const char*>(fl + (size_t)t * ldK);
            for (int c = lane; c < chunks; c += 32) {
                unsigned d32 = (unsigned)__cvta_generic_to_shared(dst + 16 * c);
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d32), "l"(src + 16 * c));
            }
        }
        asm volatile("cp.async.commit_group;\n" ::);
    };
    for (int s2 = 0; s2 < STAGES - 1; ++s2) issue(top - 1 - s2);
    // uniforms: block b covers steps t = top-1-32b-lane
    auto load_u = [&](int blk) {
        const int t = top - 1 - 32 * blk - lane;
        return (t >= cr.begin) ? un[t] : (R)0.5;
    };
    R ucur = load_u(0), unext = load_u(1);
    __syncwarp();
    int znext = -1;
    for (int t = top - 1; t >= cr.begin; --t) {
        const int step = top - 1 - t;
        if (step > 0 && (step & 31) == 0) { ucur = unext; unext = load_u((step >> 5) + 1); }
        const R u = __shfl_sync(0xffffffffu, ucur, step & 31);
        issue(t - (STAGES - 1));
        asm volatile("cp.async.wait_group %0;\n" ::"n"(STAGES - 1));
        __syncwarp();
        const R* row = ring + (size_t)(t % STAGES) * ldK;
        R v[VPL];
#pragma unroll
        for (int c = 0; c < VPL; ++c) {
            const int i = lane * VPL + c;
            v[c] = (i < K) ? row[i] : (R)0;
        }
        if (znext >= 0) {
            const R* prow = pis + (size_t)znext * ldK;
#pragma unroll
            for (int c = 0; c < VPL; ++c) {
                const int i = lane * VPL + c;
                v[c] = (i < K) ? v[c] * prow[i] : (R)0;
            }
        }
        R c[VPL];
        R run = 0;
#pragma unroll
        for (int q = 0; q < VPL; ++q) { run += v[q]; c[q] = run; }
        R incl = run;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            R y = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += y;
        }
        const R excl = incl - run;
        const R total = __shfl_sync(0xffffffffu, incl, 31);
        const R r = total * ((R)1 - u);
        int cnt = 0;
#pragma unroll
        for (int q = 0; q < VPL; ++q) {
            const int i = lane * VPL + q;
            cnt += (i < K && (excl + c[q]) < r) ? 1 : 0;
        }
        cnt = __reduce_add_sync(0xffffffffu, cnt);
        znext = min(cnt, K - 1);
        if (lane == 0) {
            if (t < cr.end) zn[t] = znext;
            if (t == cr.end) bz_warm[(size_t)nn * C + ck + 1] = znext;
            if (t == cr.begin && ck > 0) bz_exact[(size_t)nn * C + ck] = znext;
        }
        __syncwarp();
    }
}

// exact boundary check for the label paths
__global__ void label_check_kernel(const int* __restrict__ warm, const int* __restrict__ exact,
                                   const int* __restrict__ vlen, int N, int Tp, int C, int Wm,
                                   int* __restrict__ dirty, unsigned* __restrict__ stats) {
    const int nn = blockIdx.x * blockDim.x + threadIdx.x;
    if (nn >= N) return;
    int bad = 0;
    for (int c = 1; c < C; ++c) {
        if (chunk_range(vlen[nn], Tp, C, Wm, c, 8).empty) break;
        bad |= warm[(size_t)nn * C + c] != exact[(size_t)nn * C + c];
    }
    dirty[nn] = bad;
    if (bad) atomicAdd(&stats[1], 1u);
}

template <typename R>
__global__ void transpose_pi_kernel(const R* __restrict__ pi, int K, int ldK, R* __restrict__ piT) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= K * ldK) return;
    int jn = idx / ldK, i = idx % ldK;      // piT[jn][i] = pi[i][jn]
    piT[idx] = (i < K) ? pi[(size_t)i * K + jn] : (R)0;
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

template <typename R>
static size_t hmm_ws_bytes(int K, int d, int L) {
    int Fp = fp_of(d * L, d, sizeof(R));
    int ldK = (K + 3) / 4 * 4;
    size_t b = 0;
    b += align_up((size_t)K * d * Fp * sizeof(R), 256);   // G
    b += align_up((size_t)K * sizeof(R), 256);            // cst
    b += align_up((size_t)K * ldK * sizeof(R), 256);      // piT
    return b;
}

template <typename R, int D_, int L_>
static int ar_loglik_launch(const R* x, const int* mask, const R* Ab, const R* Q, int N, int T, int K,
                            int ldT, R* W, R* mx, void* ws, cudaStream_t st) {
    constexpr int n = D_ * L_;
    constexpr int FPT = sizeof(R) == 4 ? 2 : 1;
    constexpr int KC = sizeof(R) == 4 ? 32 : 16;
    int Fp = fp_of(n, D_, sizeof(R));
    R* G = reinterpret_cast<R*>(ws);
    R* cst = reinterpret_cast<R*>(reinterpret_cast<char*>(ws) + align_up((size_t)K * D_ * Fp * sizeof(R), 256));
    { KPMS_LAUNCH("ar_prep", st);
    ar_prep_kernel<R, D_, L_><<<ceil_div(K, 64), 64, 0, st>>>(Ab, Q, K, G, cst, Fp); }
    constexpr int FR = 128 * FPT;
    size_t smem = (align_up((size_t)(FR + L_) * D_, 4) + (size_t)KC * D_ * Fp + KC) * sizeof(R);
    auto kern = ar_loglik_kernel<R, D_, L_, FPT, KC>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    dim3 grid(ceil_div(T - L_, FR), N);
    { KPMS_LAUNCH("ar_loglik", st); kern<<<grid, 128, smem, st>>>(x, mask, G, cst, N, T, K, Fp, ldT, W, mx); }
    return check_launch("ar_loglik");
}

template <typename R>
static int ar_loglik_impl(const void* x, const int* mask, const void* Ab, const void* Q, int N, int T, int d,
                          int L, int K, int ldT, void* W, void* mx, void* ws, cudaStream_t st) {
    if (T <= L) return set_error(-3, "ar_loglik: T (%d) must exceed nlags (%d)", T, L);
    if (ldT < T - L || ldT % 8) return set_error(-3, "ar_loglik: ldT (%d) must be a multiple of 8 and >= T-L", ldT);
#define X(DD, LL)                                                                                         \
    if (d == DD && L == LL)                                                                               \
        return ar_loglik_launch<R, DD, LL>((const R*)x, mask, (const R*)Ab, (const R*)Q, N, T, K, ldT,    \
                                           (R*)W, (R*)mx, ws, st);
    KPMS_FOR_EACH_DL(X)
#undef X
    return set_error(-3, "ar_loglik: unsupported (latent_dim, nlags) = (%d, %d)", d, L);
}

template <typename R>
static int hmm_forward_impl(const void* W, const void* mx, const void* pi, int N, int K, int Tp, int ldT,
                            void* filt, double* logZ, cudaStream_t st) {
    int ldK = (K + 3) / 4 * 4;
    int Kpad = (K + 7) / 8 * 8;
    dim3 grid(N), block(4 * Kpad);
#define LAUNCH(RPT)                                                                                      \
    { KPMS_LAUNCH("hmm_forward", st);                                                                  \
    hmm_forward_kernel<R, RPT><<<grid, block, 0, st>>>((const R*)W, (const R*)mx, (const R*)pi, K, Tp,  \
                                                       ldT, ldK, (R*)filt, logZ, 1, 0, nullptr, nullptr,  \
                                                       nullptr, nullptr, nullptr); }
    if (K <= 28) { LAUNCH(7); }
    else if (K <= 52) { LAUNCH(13); }
    else if (K <= 100) { LAUNCH(25); }
    else if (K <= 128) { LAUNCH(32); }
    else return set_error(-3, "hmm_forward: num_states %d > 128 not supported", K);
#undef LAUNCH
    return check_launch("hmm_forward");
}

template <typename R>
static int hmm_backward_impl(const void* filt, const void* pi, const void* u, void* u_scratch, uint64_t seed, int N,
                             int K, int Tp, int* z, void* ws, int d, int L, cudaStream_t st) {
    int ldK = (K + 3) / 4 * 4;
    int Fp = fp_of(d * L, d, sizeof(R));
    char* base = reinterpret_cast<char*>(ws);
    R* piT = reinterpret_cast<R*>(base + align_up((size_t)K * d * Fp * sizeof(R), 256) + align_up((size_t)K * sizeof(R), 256));
    if (K > 128) return set_error(-3, "hmm_backward: num_states %d > 128 not supported", K);
    const R* usrc = (const R*)u;
    if (!usrc) {
        if (!u_scratch) return set_error(-3, "hmm_backward: u_scratch (N*Tp reals) is required when no tape is given");
        const long long count = (long long)N * Tp;
        KPMS_LAUNCH("hmm_uniforms", st);
        fill_uniform_kernel<R><<<(int)((count + 255) / 256), 256, 0, st>>>((R*)u_scratch, count, seed, KPMS_STREAM_Z);
        usrc = (const R*)u_scratch;
    }
    { KPMS_LAUNCH("transpose_pi", st); transpose_pi_kernel<R><<<ceil_div(K * ldK, 256), 256, 0, st>>>((const R*)pi, K, ldK, piT); }
    constexpr int STAGES = 8;
    size_t smem = ((size_t)K * ldK + (size_t)STAGES * ldK) * sizeof(R);
    auto kern = hmm_backward_kernel<R, 4, STAGES>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    { KPMS_LAUNCH("hmm_backward", st); kern<<<N, 32, smem, st>>>((const R*)filt, piT, usrc, K, Tp, ldK, z, 1, 0, nullptr, nullptr, nullptr, nullptr); }
    return check_launch("hmm_backward");
}

template <typename R>
static int hmm_smooth_impl(const void* filt, const void* pi, int N, int K, int Tp, void* marg, cudaStream_t st) {
    int ldK = (K + 3) / 4 * 4;
    size_t smem = ((size_t)K * K + 3 * K) * sizeof(R);
    auto kern = hmm_smooth_kernel<R>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int threads = (K + 31) / 32 * 32;
    { KPMS_LAUNCH("hmm_smooth", st); kern<<<N, threads, smem, st>>>((const R*)filt, (const R*)pi, K, Tp, ldK, (R*)marg); }
    return check_launch("hmm_smooth");
}

}  // namespace kpms

using namespace kpms;

extern "C" {

size_t kpms_hmm_workspace_bytes(int dtype, int K, int d, int L) {
    return dtype == 0 ? hmm_ws_bytes<float>(K, d, L) : hmm_ws_bytes<double>(K, d, L);
}

int kpms_ar_loglik(int dtype, const void* x, const int* mask, const void* Ab, const void* Q, int N, int T,
                   int d, int L, int K, int ldT, void* W, void* mx, void* ws, void* stream) {
    return KPMS_DISPATCH_DTYPE(dtype, ar_loglik_impl, x, mask, Ab, Q, N, T, d, L, K, ldT, W, mx, ws,
                               (cudaStream_t)stream);
}

int kpms_hmm_forward(int dtype, const void* W, const void* mx, const void* pi, int N, int K, int Tp, int ldT,
                     void* filt, double* logZ, void* stream) {
    return KPMS_DISPATCH_DTYPE(dtype, hmm_forward_impl, W, mx, pi, N, K, Tp, ldT, filt, logZ,
                               (cudaStream_t)stream);
}

int kpms_hmm_backward_sample(int dtype, const void* filt, const void* pi, const void* u_tape, void* u_scratch,
                             uint64_t seed, int N, int K, int Tp, int* z, void* ws, int d, int L, void* stream) {
    return KPMS_DISPATCH_DTYPE(dtype, hmm_backward_impl, filt, pi, u_tape, u_scratch, seed, N, K, Tp, z, ws, d, L,
                               (cudaStream_t)stream);
}

int kpms_hmm_smooth(int dtype, const void* filt, const void* pi, int N, int K, int Tp, void* marg, void* stream) {
    return KPMS_DISPATCH_DTYPE(dtype, hmm_smooth_impl, filt, pi, N, K, Tp, marg, (cudaStream_t)stream);
}

}  // extern "C"
