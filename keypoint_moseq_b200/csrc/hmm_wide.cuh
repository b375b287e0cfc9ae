// Discrete-state kernels for 128 < num_states <= 512 (float64 only).  num_states is user configuration
// (keypoint_moseq/io.py:72-83; BASELINE config 5 names 25 - 500 states).
//
// The kernels for K <= 128 keep a column slice of pi in registers (filter) or pi' in shared memory (backward
// walker); a 500 x 500 float64 matrix is 2 MB and fits neither, so here pi stays in L2 and every kernel streams
// the part it needs:
//   ar_loglik   the state loop runs over a run-time number of 8-state tiles; raw log-likelihoods go to W while
//               the frame maximum is tracked, a second pass over the thread's own entries forms exp(ll - max)
//   forward     one CTA advances 8 (chain, chunk) tasks in lockstep, thread j owns state column j for all 8 tasks:
//               pred[m][j] = sum_i q[m][i] pi[i][j], pi rows read coalesced from L2 (one pass per step), q
//               broadcast from shared memory; 8 K^2 DFMA and K^2 x 8 bytes of L2 traffic per CTA step
//   backward    one warp per (chain, chunk): per step the filtered row times the column of pi selected by the
//               label above, lane-local running sums + warp scan, inverse CDF by counting
// Task scheme, boundary records, passes and the repair step are the ones of the K <= 128 kernels (hmm.cu,
// hmm_f64.cuh); so are the draws: z = #{ i : c_i < (1 - u) c_{K-1} }.
// Included by hmm.cu inside namespace kpms.
#pragma once

constexpr int HMM_WIDE_MAX = 512;      // largest num_states
constexpr int HMM_WIDE_M = 8;          // tasks per CTA of the wide filter

// ---------------------------------------------------------------------------
// K2, run-time number of state tiles (layout and operators as ar_loglik_dmma_kernel)
// ---------------------------------------------------------------------------
template <int D_, int L_>
__global__ void __launch_bounds__(32 * AR_WARPS, 3)
ar_loglik_dmma_wide_kernel(const double* __restrict__ x, const int* __restrict__ mask, const double* __restrict__ Gf,
                           int N, int T, int K, int KT, int ldT, double* __restrict__ W, double* __restrict__ mx) {
    typedef ArFrag<D_, L_> A;
    constexpr int FR = 8 * AR_WARPS, NT = 32 * AR_WARPS, KK = A::KK, NF = A::NF;
    const int ldKw = 8 * KT;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* xs = reinterpret_cast<double*>(smem_raw);                 // (FR + L) * D_
    double* gs = xs + align_up((size_t)(FR + L_) * D_, 2);            // 2 x CHUNK
    const int nn = blockIdx.y, Tp = T - L_, t0 = blockIdx.x * FR;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, p = lane & 3;
    const double* xrow = x + (size_t)nn * T * D_;
    const int tile_vals = min(FR + L_, T - t0) * D_;
    for (int i = tid; i < (FR + L_) * D_; i += NT) xs[i] = i < tile_vals ? xrow[(size_t)t0 * D_ + i] : 0.0;
    auto stage = [&](int kt, int buf) {
        const double* src = Gf + (size_t)kt * A::CHUNK;
        double* dst = gs + (size_t)buf * A::CHUNK;
        for (int i = tid * 2; i < A::CHUNK; i += 2 * NT) cp_async_16(dst + i, src + i);
        asm volatile("cp.async.commit_group;\n" ::);
    };
    stage(0, 0);
    const int lt = warp * 8 + g;
    const int tp = t0 + lt;
    const bool valid = tp < Tp;
    const bool on = valid && mask[(size_t)nn * T + tp + L_] != 0;
    const bool warp_on = __any_sync(0xffffffffu, on);
    __syncthreads();
    double a[KK];
#pragma unroll
    for (int kk = 0; kk < KK; ++kk) a[kk] = (4 * kk + p < NF) ? xs[lt * D_ + 4 * kk + p] : 0.0;
    double* wrow = W + ((size_t)nn * Tp + (valid ? tp : 0)) * ldKw;
    double best = -INFINITY;
    for (int kt = 0; kt < KT; ++kt) {
        asm volatile("cp.async.wait_group 0;\n" ::);
        __syncthreads();
        if (kt + 1 < KT) stage(kt + 1, (kt + 1) & 1);
        const double* gb = gs + (size_t)(kt & 1) * A::CHUNK;
        double acc0 = 0.0, acc1 = 0.0;
        if (warp_on) {
#pragma unroll 2
            for (int i = 0; i < D_; ++i) {
                const double2 bias = *reinterpret_cast<const double2*>(gb + A::FRAG + i * 8 + 2 * p);
                double c0 = bias.x, c1 = bias.y, e0 = 0.0, e1 = 0.0;
                const double* bf = gb + (size_t)i * KK * 32 + lane;
#pragma unroll
                for (int kk = 0; kk < KK; ++kk) {
                    const double b = bf[kk * 32];
                    if (kk & 1) dmma884(e0, e1, a[kk], b);
                    else dmma884(c0, c1, a[kk], b);
                }
                c0 += e0;
                c1 += e1;
                acc0 = fma(c0, c0, acc0);
                acc1 = fma(c1, c1, acc1);
            }
        }
        const double2 cs = *reinterpret_cast<const double2*>(gb + A::FRAG + D_ * 8 + 2 * p);
        const double l0 = on ? fma(-0.5, acc0, cs.x) : 0.0, l1 = on ? fma(-0.5, acc1, cs.y) : 0.0;
        const int s0 = 8 * kt + 2 * p;
        if (s0 < K) best = fmax(best, l0);
        if (s0 + 1 < K) best = fmax(best, l1);
        if (valid) *reinterpret_cast<double2*>(wrow + s0) = make_double2(l0, l1);      // raw, rescaled below
    }
    best = fmax(best, __shfl_xor_sync(0xffffffffu, best, 1));
    best = fmax(best, __shfl_xor_sync(0xffffffffu, best, 2));
    if (valid) {
        for (int kt = 0; kt < KT; ++kt) {                  // this thread's own entries: no ordering issue
            const int s0 = 8 * kt + 2 * p;
            const double2 l = *reinterpret_cast<const double2*>(wrow + s0);
            double2 o;
            o.x = s0 < K ? exp(l.x - best) : 0.0;
            o.y = s0 + 1 < K ? exp(l.y - best) : 0.0;
            *reinterpret_cast<double2*>(wrow + s0) = o;
        }
        if (p == 0) mx[(size_t)nn * ldT + tp] = best;
    }
}

#if KPMS_DL_GROUP == 0      // not (latent_dim, nlags)-dependent: compiled once
// ---------------------------------------------------------------------------
// K3 forward, wide.  Arguments and passes as hmm_forward_dmma_kernel; Kp = row stride of W (8 * tiles).
// Shared memory: q[2][Kp][8] (state-major so that the mat-vec reads the 8 tasks of a state as four 16-byte
// broadcasts) | part[2][8][warps] | tasks.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(HMM_WIDE_MAX, 1)
hmm_forward_wide_kernel(const double* __restrict__ W, const double* __restrict__ mx, const double* __restrict__ pi,
                        int N, int K, int Kp, int Tp, int ldT, int ldK, double* __restrict__ filt,
                        double* __restrict__ logZ, double* __restrict__ logZ_part, int pass, int C, int CT, int Wm,
                        const int* __restrict__ vb, const int* __restrict__ dirty, double* __restrict__ bnd_warm,
                        double* __restrict__ bnd_end, const double* __restrict__ tail_start) {
    constexpr int M = HMM_WIDE_M;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    double* qbuf = reinterpret_cast<double*>(smem_raw);                 // 2 x Kp x M
    double* part = qbuf + (size_t)2 * Kp * M;                           // 2 x M x nwarps
    __shared__ HmmTask tks[M];
    __shared__ double msum_s[M];
    __shared__ int maxlen_s;
    if (tid < M) tks[tid] = hmm_task((long long)blockIdx.x * M + tid, pass, N, Tp, C, CT, Wm, vb, dirty);
    __syncthreads();
    if (tid == 0) {
        int ml = 0;
        for (int m = 0; m < M; ++m)
            if (tks[m].on) ml = max(ml, tks[m].end - tks[m].start);
        maxlen_s = ml;
    }
    __syncthreads();
    const int maxlen = maxlen_s;
    if (maxlen == 0) return;
    for (int m = warp; m < M; m += nwarps) {               // sums of the per-frame maxima (part of the log-normaliser)
        const HmmTask t = tks[m];
        double acc = 0.0;
        if (t.on)
            for (int tt = t.begin + lane; tt < t.end; tt += 32) acc += mx[(size_t)t.nn * ldT + tt];
        acc = warp_sum(acc);
        if (lane == 0) msum_s[m] = acc;
    }
    const int j = tid;                                     // this thread's state column
    const bool col = j < K;
    double pr[M], lz = 0.0, lzp = 1.0;                      // lz*: task `tid` (threads 0..M-1)
    int lze = 0;
    const double* wp[M];
#pragma unroll
    for (int m = 0; m < M; ++m) {
        const HmmTask t = tks[m];
        pr[m] = col ? 1.0 / (double)K : 0.0;
        if (t.on && t.given && col) pr[m] = tail_start[((size_t)t.nn * (pass == 3 ? C : CT) + t.slot) * K + j];
        wp[m] = W + ((size_t)t.nn * Tp + t.start) * Kp + j;
    }
    double wa[M];
    auto fetch = [&](int m, int r) -> double {
        const HmmTask& t = tks[m];
        return (t.on && j < Kp && t.start + r < t.end) ? wp[m][(size_t)r * Kp] : 0.0;
    };
#pragma unroll
    for (int m = 0; m < M; ++m) wa[m] = fetch(m, 0);
    int buf = 0;
    for (int r = 0; r < maxlen; ++r) {
        double q[M];
        double* qb = qbuf + (size_t)buf * Kp * M;
        double* pb = part + (size_t)buf * M * nwarps;
#pragma unroll
        for (int m = 0; m < M; ++m) {
            const HmmTask& t = tks[m];
            const int tt = t.start + r;
            const bool act = t.on && tt < t.end;
            const double w = wa[m];
            wa[m] = fetch(m, r + 1);
            if (act && pass == 0 && !t.given && t.slot > 0 && tt == t.begin && col)
                bnd_warm[((size_t)t.nn * C + t.slot) * K + j] = pr[m];
            q[m] = act ? pr[m] * w : 0.0;
            const double ps = warp_sum(q[m]);
            if (lane == 0) pb[m * nwarps + warp] = ps;
        }
        if (j < Kp) {
#pragma unroll
            for (int m = 0; m < M; m += 2) *reinterpret_cast<double2*>(qb + (size_t)j * M + m) = make_double2(q[m], q[m + 1]);
        }
        __syncthreads();
        double inv_s[M];
#pragma unroll
        for (int m = 0; m < M; ++m) {
            const HmmTask& t = tks[m];
            const int tt = t.start + r;
            const bool act = t.on && tt < t.end;
            double s = 0.0;
            for (int w2 = 0; w2 < nwarps; ++w2) s += pb[m * nwarps + w2];
            inv_s[m] = act ? rcp_fast<double>(s) : 1.0;
            if (act && tt >= t.begin) {
                if (j < ldK) filt[((size_t)t.nn * Tp + tt) * ldK + j] = q[m] * inv_s[m];
                if (tid == m) {                          // log s accumulated as mantissa product + exponent
                    int ex;
                    lzp *= frexp(s, &ex);
                    lze += ex;
                    if ((r & 7) == 7) { lz += log(lzp); lzp = 1.0; }
                }
            }
        }
        double acc[M];
#pragma unroll
        for (int m = 0; m < M; ++m) acc[m] = 0.0;
        if (col) {
            const double* pcol = pi + j;
#pragma unroll 4
            for (int i = 0; i < K; ++i) {
                const double pv = __ldg(pcol + (size_t)i * K);
                const double2* qi = reinterpret_cast<const double2*>(qb + (size_t)i * M);
#pragma unroll
                for (int m = 0; m < M; m += 2) {
                    const double2 qq = qi[m / 2];
                    acc[m] = fma(pv, qq.x, acc[m]);
                    acc[m + 1] = fma(pv, qq.y, acc[m + 1]);
                }
            }
        }
#pragma unroll
        for (int m = 0; m < M; ++m) {
            const HmmTask& t = tks[m];
            if (t.on && t.start + r < t.end) pr[m] = acc[m] * inv_s[m];
        }
        buf ^= 1;
    }
#pragma unroll
    for (int m = 0; m < M; ++m) {
        const HmmTask& t = tks[m];
        if (!t.on) continue;
        if ((pass == 0 || pass == 3) && t.end < Tp && col)       // handed to the next chunk / to the padded tail
            (pass == 3 ? bnd_warm : bnd_end)[((size_t)t.nn * C + t.slot + 1) * K + j] = pr[m];
        if (tid == m) {
            const double val = lz + log(lzp) + 0.6931471805599453094 * (double)lze + msum_s[m];
            if (pass == 2) logZ[t.nn] = val;
            else logZ_part[(size_t)t.nn * (C + CT) + (pass == 1 ? C + t.slot : t.slot)] = val;
        }
    }
}

// predictions at the start of every padded-tail chunk, any K (see hmm_tail_starts_kernel); blockDim >= K
__global__ void hmm_tail_starts_wide_kernel(const double* __restrict__ PTL, const double* __restrict__ bnd_end,
                                            const int* __restrict__ vb, const int* __restrict__ dirty, int K, int Tp,
                                            int C, int CT, int Wm, double* __restrict__ tail_start) {
    __shared__ double pcur[HMM_WIDE_MAX];
    const int nn = blockIdx.x, jn = threadIdx.x;
    const int v = vb[nn];
    if (dirty[nn] != 0 || v >= Tp) return;
    int last = -1;
    for (int c = 0; c < C; ++c) {
        const ChunkRange cr = chunk_range(v, v, C, Wm, c, 8);
        if (!cr.empty && cr.begin < cr.end) last = c;
    }
    double val = 1.0 / (double)K;
    if (last >= 0 && jn < K) val = bnd_end[((size_t)nn * C + last + 1) * K + jn];
    const int nk = (Tp - v + HMM_TL - 1) / HMM_TL;
    for (int k = 0; k < nk; ++k) {
        if (jn < K) tail_start[((size_t)nn * CT + k) * K + jn] = val;
        __syncthreads();
        pcur[jn] = (jn < K) ? val : 0.0;
        __syncthreads();
        double acc = 0.0;
        if (jn < K)
            for (int i = 0; i < K; ++i) acc = fma(pcur[i], PTL[(size_t)i * K + jn], acc);
        val = acc;
    }
}

// ---------------------------------------------------------------------------
// K3 backward, wide: one warp per (chain, chunk), EPL consecutive states per lane (32 * EPL >= ldK), the column of
// pi selected by the label above read as a row of pi' from L2.  Same chunk / zwarm / repair protocol and the same
// inverse CDF as hmm_backward_walk_kernel; every step is a full draw (no stay test).
// ---------------------------------------------------------------------------
template <int EPL>
__global__ void __launch_bounds__(128)
hmm_backward_wide_kernel(const double* __restrict__ filt, const double* __restrict__ piT, const double* __restrict__ u,
                         int N, int K, int Tp, int ldK, int Cb, int Wm, const int* __restrict__ vb, int repair,
                         int* __restrict__ z, int* __restrict__ zwarm, unsigned* __restrict__ diag) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long id = (long long)blockIdx.x * (blockDim.x >> 5) + warp;
    const int nn = repair ? (int)id : (int)(id / Cb);
    if (nn >= N) return;
    const int v = vb ? vb[nn] : Tp;
    const double* fl = filt + (size_t)nn * Tp * ldK;
    const double* un = u + (size_t)nn * Tp;
    int* zn = z + (size_t)nn * Tp;
    const int e0 = lane * EPL;

    auto load_row = [&](const double* row, double (&out)[EPL]) {
#pragma unroll
        for (int e = 0; e < EPL; e += 2) {
            double2 val = make_double2(0.0, 0.0);
            if (e0 + e < ldK) val = *reinterpret_cast<const double2*>(row + e0 + e);      // ldK is a multiple of 4
            out[e] = val.x;
            out[e + 1] = val.y;
        }
    };
    // label at step tt given the label `above` at tt + 1 (above < 0: filtered marginal alone)
    auto draw = [&](const double (&f)[EPL], double ut, int above) -> int {
        double c[EPL];
        if (above >= 0) {
            double pc[EPL];
            load_row(piT + (size_t)above * ldK, pc);
#pragma unroll
            for (int e = 0; e < EPL; ++e) c[e] = f[e] * pc[e];
        } else {
#pragma unroll
            for (int e = 0; e < EPL; ++e) c[e] = f[e];
        }
#pragma unroll
        for (int e = 1; e < EPL; ++e) c[e] += c[e - 1];
        double incl = c[EPL - 1];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const double up = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += up;
        }
        const double total = __shfl_sync(0xffffffffu, incl, 31);
        const double base = incl - c[EPL - 1];
        const double thr = total * (1.0 - ut);
        int cnt = 0;
#pragma unroll
        for (int e = 0; e < EPL; ++e) cnt += (e0 + e < K && base + c[e] < thr) ? 1 : 0;
        cnt = warp_sum(cnt);
        return min(cnt, K - 1);
    };
    // walk from t_hi down to t_lo (see hmm_backward_walk_kernel::walk)
    auto walk = [&](int t_hi, int t_lo, int init, int store_hi, int* warm_out, bool merge) -> unsigned {
        unsigned written = 0;
        int jcur = init;
        double f[EPL], fn[EPL];
        load_row(fl + (size_t)t_hi * ldK, f);
#pragma unroll
        for (int e = 0; e < EPL; ++e) fn[e] = 0.0;
        double ut = un[t_hi];
        for (int t = t_hi; t >= t_lo; --t) {
            double utn = 0.0;
            if (t > t_lo) { load_row(fl + (size_t)(t - 1) * ldK, fn); utn = un[t - 1]; }      // next row in flight
            const int lab = draw(f, ut, jcur);
            if (merge && zn[t] == lab) return written;
            __syncwarp();
            if (lane == 0) {
                if (t < store_hi) zn[t] = lab;
                else if (t == store_hi && warm_out) *warm_out = lab;
            }
            ++written;
            jcur = lab;
#pragma unroll
            for (int e = 0; e < EPL; ++e) f[e] = fn[e];
            ut = utn;
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
    }
    if (lane == 0 && mism) { atomicAdd(&diag[2], mism); atomicAdd(&diag[3], steps); }
}

// ---------------------------------------------------------------------------
// smoothed marginals, wide: as hmm_smooth_kernel with pi read from L2, both products with coalesced rows
// (the prediction column-per-thread, the backward product row-per-warp)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(HMM_WIDE_MAX)
hmm_smooth_wide_kernel(const double* __restrict__ filt, const double* __restrict__ pi, int K, int Tp, int ldK,
                       double* __restrict__ marg) {
    __shared__ double sm[HMM_WIDE_MAX], ratio[HMM_WIDE_MAX], fcur[HMM_WIDE_MAX], mine_s[HMM_WIDE_MAX];
    __shared__ double red[32];
    const int nn = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    const double* fl = filt + (size_t)nn * Tp * ldK;
    double* mg = marg + (size_t)nn * Tp * K;
    if (tid < K) { const double val = fl[(size_t)(Tp - 1) * ldK + tid]; sm[tid] = val; mg[(size_t)(Tp - 1) * K + tid] = val; }
    __syncthreads();
    for (int t = Tp - 2; t >= 0; --t) {
        if (tid < K) fcur[tid] = fl[(size_t)t * ldK + tid];
        __syncthreads();
        if (tid < K) {
            double pred = 0.0;
            for (int i = 0; i < K; ++i) pred = fma(fcur[i], __ldg(pi + (size_t)i * K + tid), pred);
            ratio[tid] = pred > 0.0 ? sm[tid] / pred : 0.0;
        }
        __syncthreads();
        for (int row = warp; row < K; row += nwarps) {
            double acc = 0.0;
            for (int jn = lane; jn < K; jn += 32) acc = fma(__ldg(pi + (size_t)row * K + jn), ratio[jn], acc);
            acc = warp_sum(acc);
            if (lane == 0) mine_s[row] = fcur[row] * acc;
        }
        __syncthreads();
        const double mine = tid < K ? mine_s[tid] : 0.0;
        const double tot = block_sum(mine, red);
        if (tid < K) { const double val = mine / tot; sm[tid] = val; mg[(size_t)t * K + tid] = val; }
        __syncthreads();
    }
}
#endif  // KPMS_DL_GROUP == 0
