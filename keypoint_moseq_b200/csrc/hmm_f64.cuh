// Float64 AR-HMM kernels on the FP64 tensor pipe (DMMA.8x8x4, mma.sync m8n8k4 f64; measured
// 37.1 TFLOP/s on B200 against 34 TFLOP/s of plain DFMA at one eighth of the issue slots,
// tools/micro/dmma_peak.cu).  The two dense contractions of the discrete-state path are real
// GEMMs once frames / lock-stepped chains are batched eight at a time:
//   ar_loglik    [frames x (n+d)] x [(n+d) x K*d]   whitened residuals, squared and summed per state
//   hmm_forward  [8 tasks x K]   x [K x K]          one filter step of eight (chain, chunk) tasks
// Included by hmm.cu inside namespace kpms (uses HmmTask / hmm_task / chunk helpers from there).
//
// Fragment layout of mma.m8n8k4 (lane l): A[row l/4][col l%4], B[row l%4][col l/4],
// C[row l/4][cols 2(l%4), 2(l%4)+1].
#pragma once

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// ---------------------------------------------------------------------------
// Whitened regression operators in fragment order, one chunk per tile of 8 states:
//   chunk kt = [ frag (D_ x KK x 32) | bias (D_ x 8) | cst (8) ] doubles,
//   frag[i][kk][lane] = G_{8kt + lane/4}[i][4kk + lane%4]   (B operand of the k-step kk),
//   bias[i][s] = G_{8kt+s}[i][bias column],  cst[s] = -sum log diag Lq - d/2 log 2pi.
// States >= K and features >= n+d are zero.
// ---------------------------------------------------------------------------
template <int D_, int L_>
struct ArFrag {
    static constexpr int n = D_ * L_;
    static constexpr int NF = n + D_;                 // features without the bias
    static constexpr int KK = (NF + 3) / 4;           // k-steps
    static constexpr int FRAG = D_ * KK * 32;
    static constexpr int CHUNK = FRAG + D_ * 8 + 8;   // doubles per state tile
};

template <int D_, int L_>
__global__ void ar_pack_frag_kernel(const double* __restrict__ G, const double* __restrict__ cst, int K, int Fp,
                                    int KT, double* __restrict__ Gf) {
    typedef ArFrag<D_, L_> A;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= KT * A::CHUNK) return;
    const int kt = idx / A::CHUNK, e = idx % A::CHUNK;
    double v = 0.0;
    if (e < A::FRAG) {
        const int lane = e % 32, kk = (e / 32) % A::KK, i = e / (32 * A::KK);
        const int st = 8 * kt + lane / 4, f = 4 * kk + lane % 4;
        if (st < K && f < A::NF) v = G[((size_t)st * D_ + i) * Fp + f];
    } else if (e < A::FRAG + D_ * 8) {
        const int s = (e - A::FRAG) % 8, i = (e - A::FRAG) / 8;
        const int st = 8 * kt + s;
        if (st < K) v = G[((size_t)st * D_ + i) * Fp + A::NF];
    } else {
        const int st = 8 * kt + (e - A::FRAG - D_ * 8);
        if (st < K) v = cst[st];
    }
    Gf[idx] = v;
}

// ---------------------------------------------------------------------------
// K2 on the tensor pipe.  CTA = AR_WARPS warps x 8 frames, three CTAs per SM so that one CTA's
// prologue / exp epilogue overlaps the other's DMMA stream; the operator chunks stream through a
// two-deep cp.async ring; every warp keeps its 8 frames' features as A fragments in registers and
// the log-likelihoods of all states until the frame maximum is known.
// W (N, Tp, ldKw) <- exp(ll - max), states contiguous per frame (ldKw = 8*KT, pad columns 0);
// mx (N, ldT) <- max.  Masked frames: W = 1, mx = 0.
// ---------------------------------------------------------------------------
constexpr int AR_WARPS = 8;

template <int D_, int L_, int KT>
__global__ void __launch_bounds__(32 * AR_WARPS, 3)
ar_loglik_dmma_kernel(const double* __restrict__ x, const int* __restrict__ mask, const double* __restrict__ Gf,
                      int N, int T, int K, int ldT, double* __restrict__ W, double* __restrict__ mx) {
    typedef ArFrag<D_, L_> A;
    constexpr int FR = 8 * AR_WARPS, NT = 32 * AR_WARPS, KK = A::KK, NF = A::NF, ldKw = 8 * KT;
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
    const int lt = warp * 8 + g;                       // this lane's frame inside the tile (C row)
    const int tp = t0 + lt;
    const bool valid = tp < Tp;
    const bool on = valid && mask[(size_t)nn * T + tp + L_] != 0;
    const bool warp_on = __any_sync(0xffffffffu, on);
    __syncthreads();
    double a[KK];
#pragma unroll
    for (int kk = 0; kk < KK; ++kk) a[kk] = (4 * kk + p < NF) ? xs[lt * D_ + 4 * kk + p] : 0.0;
    double ll[KT][2];
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
        ll[kt][0] = on ? fma(-0.5, acc0, cs.x) : 0.0;
        ll[kt][1] = on ? fma(-0.5, acc1, cs.y) : 0.0;
    }
    double best = -INFINITY;
#pragma unroll
    for (int kt = 0; kt < KT; ++kt) {
        const int s0 = 8 * kt + 2 * p;
        if (s0 < K) best = fmax(best, ll[kt][0]);
        if (s0 + 1 < K) best = fmax(best, ll[kt][1]);
    }
    best = fmax(best, __shfl_xor_sync(0xffffffffu, best, 1));
    best = fmax(best, __shfl_xor_sync(0xffffffffu, best, 2));
    if (valid) {
        double* wrow = W + ((size_t)nn * Tp + tp) * ldKw;
#pragma unroll
        for (int kt = 0; kt < KT; ++kt) {
            const int s0 = 8 * kt + 2 * p;
            double2 o;
            o.x = s0 < K ? exp(ll[kt][0] - best) : 0.0;
            o.y = s0 + 1 < K ? exp(ll[kt][1] - best) : 0.0;
            *reinterpret_cast<double2*>(wrow + s0) = o;
        }
        if (p == 0) mx[(size_t)nn * ldT + tp] = best;
    }
}

// ---------------------------------------------------------------------------
// K3 forward on the tensor pipe.  One CTA advances M = 8*MT (chain, time chunk) tasks in lockstep;
// warp w owns the state columns [8w, 8w+8) with its slice of pi resident in registers as B
// fragments; the filtered vectors of the 8 tasks of a tile are the A operand, exchanged through a
// double-buffered shared tile; one barrier per step.  Scaled filter (same recursion as the
// float kernel):  q_t = pred_t * w_t,  s_t = sum q_t,  filt_t = q_t / s_t,  pred_{t+1} = pi' filt_t.
// Task scheme (prefix chunks / padded-tail chunks / sequential re-run) as in hmm_forward_kernel, plus pass 3:
// refinement of flagged chains (prefix chunks restarted from `tail_start` = the current end states, fresh end
// states written to `bnd_warm`; see hmm_refine_check_kernel).
// ---------------------------------------------------------------------------
template <int KT, int MT>
__global__ void __launch_bounds__(32 * KT, 1)
hmm_forward_dmma_kernel(const double* __restrict__ W, const double* __restrict__ mx, const double* __restrict__ pi,
                        int N, int K, int Tp, int ldT, int ldK, double* __restrict__ filt,
                        double* __restrict__ logZ, double* __restrict__ logZ_part, int pass, int C, int CT, int Wm,
                        const int* __restrict__ vb, const int* __restrict__ dirty, double* __restrict__ bnd_warm,
                        double* __restrict__ bnd_end, const double* __restrict__ tail_start) {
    constexpr int Kp = 8 * KT, KS = Kp / 4, QS = Kp + 4, M = 8 * MT, PS = KT + 1;
    __shared__ __align__(16) double qbuf[2][M][QS];
    __shared__ double part[2][M][PS];
    __shared__ double msum_s[M];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, p = lane & 3;
    const int c0 = 8 * warp + 2 * p;                         // this lane's two state columns
    // tasks: lane group g of tile mt works on task blockIdx.x*M + mt*8 + g
    int maxlen = 0;
    for (int m = 0; m < M; ++m) {
        const HmmTask t = hmm_task((long long)blockIdx.x * M + m, pass, N, Tp, C, CT, Wm, vb, dirty);
        if (t.on) maxlen = max(maxlen, t.end - t.start);
    }
    if (maxlen == 0) return;
    HmmTask tk[MT];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
        tk[mt] = hmm_task((long long)blockIdx.x * M + mt * 8 + g, pass, N, Tp, C, CT, Wm, vb, dirty);
    double pib[KS];
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
        const int i = 4 * ks + p, j = 8 * warp + g;
        pib[ks] = (i < K && j < K) ? pi[(size_t)i * K + j] : 0.0;
    }
    // sums of the per-frame maxima (part of the log-normaliser): warp w reduces tasks w, w+KT, ...
    for (int m = warp; m < M; m += KT) {
        const HmmTask t = hmm_task((long long)blockIdx.x * M + m, pass, N, Tp, C, CT, Wm, vb, dirty);
        double acc = 0.0;
        if (t.on)
            for (int tt = t.begin + lane; tt < t.end; tt += 32) acc += mx[(size_t)t.nn * ldT + tt];
        acc = warp_sum(acc);
        if (lane == 0) msum_s[m] = acc;
    }
    double pr[MT][2], lz[MT], lzp[MT];
    int lze[MT];
    const double* wp[MT];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
        pr[mt][0] = c0 < K ? 1.0 / (double)K : 0.0;
        pr[mt][1] = c0 + 1 < K ? 1.0 / (double)K : 0.0;
        if (tk[mt].on && tk[mt].given) {
            // padded-tail chunks start from pi-power predictions; refinement chunks (pass 3) from the end
            // state of their predecessor, kept per (chain, chunk) like the boundary records
            const double* ts = tail_start + ((size_t)tk[mt].nn * (pass == 3 ? C : CT) + tk[mt].slot) * K;
            pr[mt][0] = c0 < K ? ts[c0] : 0.0;
            pr[mt][1] = c0 + 1 < K ? ts[c0 + 1] : 0.0;
        }
        lz[mt] = 0.0; lzp[mt] = 1.0; lze[mt] = 0;
        wp[mt] = W + ((size_t)tk[mt].nn * Tp + tk[mt].start) * Kp + c0;
    }
    // weights two steps ahead in registers
    double2 wa[MT], wb[MT];
    auto fetch = [&](int mt, int r) -> double2 {
        double2 v = make_double2(0.0, 0.0);
        if (tk[mt].on && tk[mt].start + r < tk[mt].end) v = *reinterpret_cast<const double2*>(wp[mt] + (size_t)r * Kp);
        return v;
    };
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) { wa[mt] = fetch(mt, 0); wb[mt] = fetch(mt, 1); }
    int buf = 0;
    for (int r = 0; r < maxlen; ++r) {
        double q[MT][2];
        bool act[MT];
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
            const int t = tk[mt].start + r;
            act[mt] = tk[mt].on && t < tk[mt].end;
            const double2 w = wa[mt];
            wa[mt] = wb[mt];
            wb[mt] = fetch(mt, r + 2);
            if (act[mt] && pass == 0 && !tk[mt].given && tk[mt].slot > 0 && t == tk[mt].begin) {
                double* bw = bnd_warm + ((size_t)tk[mt].nn * C + tk[mt].slot) * K;
                if (c0 < K) bw[c0] = pr[mt][0];
                if (c0 + 1 < K) bw[c0 + 1] = pr[mt][1];
            }
            q[mt][0] = act[mt] ? pr[mt][0] * w.x : 0.0;
            q[mt][1] = act[mt] ? pr[mt][1] * w.y : 0.0;
            *reinterpret_cast<double2*>(&qbuf[buf][mt * 8 + g][c0]) = make_double2(q[mt][0], q[mt][1]);
            double ps = q[mt][0] + q[mt][1];
            ps += __shfl_xor_sync(0xffffffffu, ps, 1);
            ps += __shfl_xor_sync(0xffffffffu, ps, 2);
            if (p == 0) part[buf][mt * 8 + g][warp] = ps;
        }
        __syncthreads();
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
            const double* pp = part[buf][mt * 8 + g];
            double s0 = 0.0, s1 = 0.0;
#pragma unroll
            for (int w = 0; w + 1 < KT; w += 2) { s0 += pp[w]; s1 += pp[w + 1]; }
            if (KT & 1) s0 += pp[KT - 1];
            const double s = s0 + s1;
            const double inv_s = act[mt] ? rcp_fast<double>(s) : 1.0;
            const int t = tk[mt].start + r;
            if (act[mt] && t >= tk[mt].begin) {
                if (c0 < ldK)
                    *reinterpret_cast<double2*>(filt + ((size_t)tk[mt].nn * Tp + t) * ldK + c0) =
                        make_double2(q[mt][0] * inv_s, q[mt][1] * inv_s);
                if (warp == 0 && p == 0) {               // log s accumulated as mantissa product + exponent
                    int ex;
                    lzp[mt] *= frexp(s, &ex);
                    lze[mt] += ex;
                    if ((r & 7) == 7) { lz[mt] += log(lzp[mt]); lzp[mt] = 1.0; }
                }
            }
            double a0 = 0.0, a1 = 0.0, b0 = 0.0, b1 = 0.0, d0 = 0.0, d1 = 0.0, e0 = 0.0, e1 = 0.0;
            const double* qa = &qbuf[buf][mt * 8 + g][p];
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) {
                const double av = qa[4 * ks];
                if ((ks & 3) == 0) dmma884(a0, a1, av, pib[ks]);
                else if ((ks & 3) == 1) dmma884(b0, b1, av, pib[ks]);
                else if ((ks & 3) == 2) dmma884(d0, d1, av, pib[ks]);
                else dmma884(e0, e1, av, pib[ks]);
            }
            if (act[mt]) {
                pr[mt][0] = ((a0 + b0) + (d0 + e0)) * inv_s;
                pr[mt][1] = ((a1 + b1) + (d1 + e1)) * inv_s;
            }
        }
        buf ^= 1;
    }
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
        if (!tk[mt].on) continue;
        if ((pass == 0 || pass == 3) && tk[mt].end < Tp) {    // handed to the next chunk / to the padded tail
            double* be = (pass == 3 ? bnd_warm : bnd_end) + ((size_t)tk[mt].nn * C + tk[mt].slot + 1) * K;
            if (c0 < K) be[c0] = pr[mt][0];
            if (c0 + 1 < K) be[c0 + 1] = pr[mt][1];
        }
        if (warp == 0 && p == 0) {
            const double val = lz[mt] + log(lzp[mt]) + 0.6931471805599453094 * (double)lze[mt] + msum_s[mt * 8 + g];
            if (pass == 2) logZ[tk[mt].nn] = val;
            else logZ_part[(size_t)tk[mt].nn * (C + CT) + (pass == 1 ? C + tk[mt].slot : tk[mt].slot)] = val;
        }
    }
}
