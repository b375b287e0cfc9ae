// K1c, two-stage form: backward preparation that uses the companion structure of the lag-augmented AR state.
//
// With xi_t = [a; c] (a = oldest d-block, c = the other n-d coordinates) the next state is
//   xi_{t+1} = [ c + e1 ;  A xi_t + b + e2 ],   e1 ~ N(0, eps1 I),  e2 ~ N(0, Q'),  eps1 = EPS_SHIFT + jitter, Q' = Q + jitter I,
// so conditioning N(m, S) on xi_{t+1} is a measurement update with the n-d shifted coordinates (isotropic noise)
// followed by one with the d new coordinates.  Stage 1 has a closed form in Z = (S_cc + eps1 I)^-1:
//   T = S_ac Z,  F = I - eps1 Z,  K1 = [T; F]  (gain of stage 1, n x (n-d)),
//   Sigma1 = [[S_aa - T S_ca, eps1 T], [eps1 T', eps1 F]]          (no cancellation in the c rows and columns)
// and stage 2 is a d x d problem:
//   W = A Sigma1 = [Wa | eps1 Wc'],  Wc' = A K1,  B2 = W A' + Q' = L2 L2',  V2 = W' L2^-T,  K2 = V2 L2^-1,
//   Sigma = Sigma1 - V2 V2',  Ls = chol(Sigma),  G = [K1 - K2 Wc' | K2],
//   h = m - K1 m_c - K2 (A m + b - Wc' m_c) + Ls w.
// About 1.3 n^3 multiply-adds per frame instead of the 2.2 - 3 n^3 of the one-stage form (A S A' + Q, two n x n
// Cholesky factorisations, two triangular solves), and smaller rounding errors (tools/numerics_backprep_study.py:
// G 1.1e-7 against 2.7e-7 median, chol(Sigma) 5e-8 against 2.6e-7, relative to float64).
// It draws the same xi_t = G xi_{t+1} + h as the sequential sampler of utils.kalman.kalman_sample.
//
// Mapping: d lanes per frame, 32/d frames per warp; lane gl owns row gl of the a block and rows gl + s d of the
// c block (one row of every d-row block), held in registers.  Shared memory carries only what other lanes must
// see, always as whole rows read back as 16-byte broadcasts:
//   phase B  Gauss-Jordan on [S_cc + eps1 I | S_ca] (pivot row published per step) -> [Z | T']
//   phase D  Sigma1 a-row from the rows of T'
//   phase E  W' rows = Sigma1' rows x A' (A' rows broadcast from the per-state operator block)
//   phase G  B2 row, t = A m + b - Wc' m_c
//   phase H  chol(B2) across the d lanes by shuffles; V2, K2 rows by substitution in registers
//   phase J  G1' rows (chunked, straight to global memory) against the rows of K2'
//   phase I  Sigma rows in place (lower triangle only), then right-looking Cholesky through a two-row ring
// Inputs arrive by three 1-D bulk copies per frame (cp.async.bulk + mbarrier): the packed covariance, the mean
// and the state's operator block, which a prep kernel lays out exactly as this kernel reads it.
// Included by kalman.cu inside namespace kpms.
#pragma once

template <typename R, int D_, int L_>
struct PrepSplit {
    static constexpr int n = D_ * L_, NO = n - D_, LA = L_ - 1;
    static constexpr int FPW = 32 / D_;                                   // frames per warp
    static constexpr int VEC = 16 / (int)sizeof(R);
    static constexpr int DP = (D_ + VEC - 1) / VEC * VEC;                 // padded d-wide row
    static constexpr int NP = (n + VEC - 1) / VEC * VEC;                  // padded n-wide row
    static constexpr int SB = stash_S_stride(n), SMS = stash_m_stride(n);
    // operator block of one state: At (n x DP, At[e][a] = A[a][e]) | b (DP) | Q' (D_ x DP)
    static constexpr int OPS = n * DP + DP + D_ * DP;
    // region X of a frame's shared-memory block, reused phase by phase:
    //   [0, SB) packed S                       (B, D)   then  W' rows, later V2 rows (n x DP)   (E..I)
    //   [X_TT, +NO*DP) T' rows                 (B..E)
    //   [X_PIV, +2*NP) pivot rows / Cholesky ring
    //   [X_K2T, +D_*NP) K2' rows (H..J),  [X_B2, +D_*DP) B2 / L2 rows (G, H)
    static constexpr int X_TT = SB, X_PIV = SB + NO * DP, X_K2T = n * DP, X_B2 = X_K2T + D_ * NP;
    static constexpr int XA = X_PIV + 2 * NP, XB = X_B2 + D_ * DP;
    static constexpr int X_END = (XA > XB ? XA : XB);
    static_assert(n * DP <= SB + NO * DP, "W' rows must not reach the pivot rows");
    static constexpr int O_OPS = X_END, O_M = O_OPS + OPS, O_W = O_M + 2 * SMS, O_T = O_W + 2 * NP, RAW = O_T + DP;   // mean, normals: ping-pong
    // frames of one warp sit 8 banks apart (mod 32) so that broadcasts with FPW distinct addresses do not collide
    static constexpr int PER_FRAME = FPW > 1 ? (RAW + 31) / 32 * 32 + 8 : RAW;
    static constexpr size_t frame_bytes = (size_t)PER_FRAME * sizeof(R);
    static constexpr unsigned TX_S = (unsigned)((SB + SMS) * sizeof(R)), TX_A = (unsigned)(OPS * sizeof(R));   // bytes per frame
};

// ---- mbarrier / bulk-copy primitives (sm_90+; SASS: SYNCS / UBLKCP) ----
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }

// Per-state operator blocks in the layout PrepSplit reads: At | b | Q + jitter I.  One CTA per state.
template <typename R, int D_, int L_>
__global__ void __launch_bounds__(128)
backprep_ops_kernel(const R* __restrict__ Ab, const R* __restrict__ Q, R jitter, R* __restrict__ ops) {
    typedef PrepSplit<R, D_, L_> P;
    constexpr int n = P::n, DP = P::DP;
    const int k = blockIdx.x;
    const R* A = Ab + (size_t)k * D_ * (n + 1);
    const R* Qk = Q + (size_t)k * D_ * D_;
    R* o = ops + (size_t)k * P::OPS;
    for (int w = threadIdx.x; w < P::OPS; w += blockDim.x) {
        R val = (R)0;
        if (w < n * DP) { const int e = w / DP, a = w % DP; if (a < D_) val = A[a * (n + 1) + e]; }
        else if (w < n * DP + DP) { const int a = w - n * DP; if (a < D_) val = A[a * (n + 1) + n]; }
        else { const int u = w - n * DP - DP, a = u / DP, c = u % DP; if (c < D_) val = Qk[a * D_ + c] + (a == c ? jitter : (R)0); }
        o[w] = val;
    }
}

// standard normals for the backward sampler, one per (frame, coordinate): four per Philox call
template <typename R>
__global__ void fill_normal_kernel(R* __restrict__ w, long long count, SeedArg seed, uint32_t stream) {
    const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (q * 4 >= count) return;
    Philox gen(seed, stream, (uint64_t)q);
    R out[4];
    if (sizeof(R) == 4) {
        const uint4 r = gen.next4();
        const float u0 = ((float)(r.x >> 8) + 0.5f) * (1.0f / 16777216.0f), u1 = ((float)(r.y >> 8) + 0.5f) * (1.0f / 16777216.0f);
        const float u2 = ((float)(r.z >> 8) + 0.5f) * (1.0f / 16777216.0f), u3 = ((float)(r.w >> 8) + 0.5f) * (1.0f / 16777216.0f);
        const float ra = sqrtf(-2.0f * logf(u0)), rb = sqrtf(-2.0f * logf(u2));
        float s0, c0, s1, c1;
        sincospif(2.0f * u1, &s0, &c0);
        sincospif(2.0f * u3, &s1, &c1);
        out[0] = (R)(ra * c0); out[1] = (R)(ra * s0); out[2] = (R)(rb * c1); out[3] = (R)(rb * s1);
    } else {
        double a0, a1, b0, b1;
        philox_normal2(gen, a0, a1);
        philox_normal2(gen, b0, b1);
        out[0] = (R)a0; out[1] = (R)a1; out[2] = (R)b0; out[3] = (R)b1;
    }
#pragma unroll
    for (int c = 0; c < 4; ++c)
        if (q * 4 + c < count) w[q * 4 + c] = out[c];
}

// Masked and terminal frames of a chain (one CTA per chain): identity records for masked frames below the last
// unmasked one (the recursion walks through them; the padded tail above is never read), and the terminal draw
// h = m + chol(S) w of the last frame.  Generic in n (shared-memory Cholesky by warp 0).
template <typename R, int D_, int L_, bool ROWPACK>
__global__ void __launch_bounds__(128)
backprep_special_kernel(const R* __restrict__ stash_m, const R* __restrict__ stash_S, const int* __restrict__ mask,
                        const R* __restrict__ wbuf, int N, int T, R* __restrict__ GH) {
    constexpr int n = D_ * L_, LD = n | 1, NP2 = n * (n + 1) / 2;
    constexpr int RECS = PrepSmem<R, D_, L_>::RECS, SMS = stash_m_stride(n), SSS = stash_S_stride(n);
    __shared__ R Ssh[n * LD];
    __shared__ R idS[n];
    __shared__ int last_valid;
    const int nn = blockIdx.x, tid = threadIdx.x, Tx = T - L_ + 1;
    const int* mk = mask + (size_t)nn * T + (L_ - 1);
    if (tid == 0) last_valid = -1;
    __syncthreads();
    for (int i = tid; i < Tx; i += blockDim.x)
        if (mk[i] != 0) atomicMax(&last_valid, i);
    __syncthreads();
    const int lv = last_valid;
    const long long g0 = (long long)nn * Tx;
    for (int i = tid; i < lv; i += blockDim.x) {       // holes below the last unmasked frame (none in padded batches)
        if (mk[i] != 0 || i == Tx - 1) continue;
        R* Gout = GH + (size_t)(g0 + i) * RECS;
        for (int w = 0; w < n * n + n; ++w) Gout[w] = (w < n * n && (w / n) == (w % n)) ? (R)1 : (R)0;
    }
    if (tid >= 32) return;
    // terminal frame: always drawn from the filter marginal (the filter carries masked steps through)
    const long long g = g0 + Tx - 1;
    const R* Sg = stash_S + (size_t)g * SSS;
    for (int q = tid; q < NP2; q += 32) {
        int r, c;
        tri_unpack(q, r, c);                           // q = r(r+1)/2 + c
        const R val = ROWPACK ? Sg[q] : Sg[col_start(n, c) + r - c];
        Ssh[r * LD + c] = val;
        Ssh[c * LD + r] = val;
    }
    __syncwarp();
    warp_cholesky<R, n, LD>(Ssh, idS, tid);
    R* hout = GH + (size_t)g * RECS + n * n;
    for (int r = tid; r < n; r += 32) {
        R acc = stash_m[(size_t)g * SMS + r];
        for (int c = 0; c <= r; ++c) acc = fma(Ssh[r * LD + c], wbuf[(size_t)g * n + c], acc);
        hout[r] = acc;
    }
}

template <typename R, int D_, int L_, int WARPS, int MINB, int LOCKSTEP = 1>
__global__ void __launch_bounds__(32 * WARPS, MINB)
kalman_backprep_split_kernel(const R* __restrict__ stash_m, const R* __restrict__ stash_S, const int* __restrict__ mask,
                             const int* __restrict__ z, const R* __restrict__ ops, R eps1, const R* __restrict__ wbuf,
                             int N, int T, R* __restrict__ GH) {
    typedef PrepSplit<R, D_, L_> P;
    typedef typename Vec16<R>::type VecT;
    constexpr int n = P::n, NO = P::NO, LA = P::LA, FPW = P::FPW, VEC = P::VEC, DP = P::DP, NP = P::NP;
    constexpr int DV = DP / VEC, NV = NP / VEC;
    constexpr int RECS = PrepSmem<R, D_, L_>::RECS;
    constexpr bool ROWPACK = stash_rowpack<R, n>();   // packing of the filter's covariance record (see kalman_rows2.cuh)
    static_assert(L_ >= 2 && D_ <= 32, "two-stage form needs shifted blocks and one warp per frame group");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int grp0 = lane / D_;
    const bool live = grp0 < FPW;
    const int grp = live ? grp0 : FPW - 1;
    const int gl = live ? lane - grp0 * D_ : D_ - 1;  // spare lanes shadow the last lane of the last group
    const int gbase = grp * D_;
    R* fb = reinterpret_cast<R*>(smem_raw) + (size_t)(warp * FPW + grp) * P::PER_FRAME;
    uint64_t* barS = reinterpret_cast<uint64_t*>(reinterpret_cast<R*>(smem_raw) + (size_t)WARPS * FPW * P::PER_FRAME) + 2 * warp;
    uint64_t* barA = barS + 1;
    R* Sb = fb;
    R* TTs = fb + P::X_TT;
    R* piv = fb + P::X_PIV;
    R* WTs = fb;
    R* K2T = fb + P::X_K2T;
    R* B2s = fb + P::X_B2;
    R* Ats = fb + P::O_OPS;
    R* bvec = Ats + n * DP;
    R* Qs = bvec + DP;
    R* mvb = fb + P::O_M;                         // 2 x SMS
    R* wvb = fb + P::O_W;                         // 2 x NP
    R* tv = fb + P::O_T;
    const int Tx = T - L_ + 1;
    const long long frames = (long long)N * Tx;
    const long long stride = (long long)gridDim.x * WARPS * FPW;
    const unsigned FULL = 0xffffffffu;

    // accessors of the packed covariance for the rows this lane owns: S[r][c] = c <= r ? pL[offL(c)] : pU[offU(c)]
    auto offL = [](int c) { return ROWPACK ? c : col_start(n, c) - c; };
    auto offU = [](int c) { return ROWPACK ? c * (c + 1) / 2 : c; };
    if (lane == 0) {
        mbar_init(barS, 1);
        mbar_init(barA, 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    __syncwarp();
    unsigned parS = 0, parA = 0;

    // Frame of this lane group in the iteration that starts at frame cb: its index and the raw words that decide
    // whether it is a regular frame.  Plain loads with clamped addresses and no branch on their values, so that they
    // can be issued at the top of an iteration and consumed much later (the host guarantees frames < 2^31).
    const unsigned uframes = (unsigned)frames, ustride = (unsigned)stride, uTx = (unsigned)Tx;
    auto peek = [&](unsigned cb, unsigned& g, int& i, int& mk, int& zr) {
        g = cb + (unsigned)(warp * FPW + grp);
        const unsigned gc = (cb < uframes && g < uframes) ? g : 0u;
        const unsigned nn = gc / uTx;
        i = (int)(gc - nn * uTx);
        mk = mask[(size_t)nn * T + (L_ - 1) + i];
        zr = Tx > 1 ? z[(size_t)nn * (Tx - 1) + min(i, Tx - 2)] : 0;
        if (!(cb < uframes && g < uframes)) i = Tx;                   // out of range: never a regular frame
    };
    // Bulk copies of a frame's inputs (every lane calls; generic-proxy accesses of the target are fenced first).
    // The covariance lands in the place the W' / V2 rows used, the operator block in its own, the mean in buffer `buf`.
    auto request_S = [&](unsigned g, bool on, unsigned onmask, int buf) {
        fence_proxy_async();
        __syncwarp();
        if (onmask == 0) return;
        if (lane == 0) mbar_expect_tx(barS, (unsigned)__popc(onmask) * P::TX_S);
        __syncwarp();
        if (on && live && gl == 0) {
            bulk_g2s(Sb, stash_S + (size_t)g * P::SB, P::SB * (unsigned)sizeof(R), barS);
            bulk_g2s(mvb + buf * P::SMS, stash_m + (size_t)g * P::SMS, P::SMS * (unsigned)sizeof(R), barS);
        }
    };
    auto request_A = [&](int zi, bool on, unsigned onmask) {
        fence_proxy_async();
        __syncwarp();
        if (onmask == 0) return;
        if (lane == 0) mbar_expect_tx(barA, (unsigned)__popc(onmask) * P::TX_A);
        __syncwarp();
        if (on && live && gl == 0) bulk_g2s(Ats, ops + (size_t)zi * P::OPS, P::OPS * (unsigned)sizeof(R), barA);
    };

    unsigned g, cbase = (unsigned)blockIdx.x * WARPS * FPW;
    int zi, buf = 0;
    bool on;
    {
        int i0, mk0;
        peek(cbase, g, i0, mk0, zi);
        on = i0 < Tx - 1 && mk0 != 0;
    }
    unsigned onmask = __ballot_sync(FULL, on && live && gl == 0);
    request_S(g, on, onmask, 0);
    request_A(zi, on, onmask);
    if (on) {
#pragma unroll
        for (int s = 0; s < L_; ++s) wvb[gl + s * D_] = wbuf[(size_t)g * n + gl + s * D_];
    }
    // Every iteration requests the NEXT frame's inputs while the current one is factored: the operator block as soon
    // as phase G has consumed it, covariance and mean once the V2 rows are consumed (before the Cholesky of Sigma).
    // The words that describe the next frame (mask, state label, normals) are loaded here and first used there.
    for (; cbase < uframes; cbase += ustride, buf ^= 1) {
        // LOCKSTEP: the unrolled body is ~150 KB of code, more than the instruction cache holds; warps of a CTA that
        // start every tile together walk the same cache lines (measured at C2: 5.47 ms without, 4.57 ms with one
        // barrier per tile, ncu `no_instruction` stall 2.3 -> see profiles/).  The trip count is uniform over the CTA
        // (cbase does not depend on the warp); LOCKSTEP = 2, 3 add barriers after phases E and J, matched by the
        // same number in the branch of a warp without regular frames.
        if (LOCKSTEP >= 1) __syncthreads();
        unsigned gN;
        int iN, mkN, ziN;
        peek(cbase + ustride, gN, iN, mkN, ziN);
        R wN[L_];
        auto load_wN = [&]() {                        // normals of the next frame (stored into the other buffer at the end)
            const size_t gw = (size_t)(iN < Tx ? gN : 0u) * n;
#pragma unroll
            for (int s = 0; s < L_; ++s) wN[s] = wbuf[gw + gl + s * D_];
        };
        bool onN;
        unsigned onmaskN;
        if (onmask == 0) {
            load_wN();
            onN = iN < Tx - 1 && mkN != 0;
            onmaskN = __ballot_sync(FULL, onN && live && gl == 0);
            request_A(ziN, onN, onmaskN);
            request_S(gN, onN, onmaskN, buf ^ 1);
            if (LOCKSTEP >= 2) __syncthreads();
            if (LOCKSTEP >= 3) __syncthreads();
        } else {
        const R* mv = mvb + buf * P::SMS;
        const R* wv = wvb + buf * NP;
        mbar_wait(barS, parS);
        parS ^= 1;
        __syncwarp();
        R* Gout = GH + (size_t)(on ? g : 0u) * RECS;

        // ---- phase B: Gauss-Jordan on [S_cc + eps1 I | S_ca], rows rc = gl + s d  ->  [Z | T']
        R m[LA][n];                                   // m[s][e] e < NO: c columns; m[s][NO + a]: a columns
#pragma unroll
        for (int s = 0; s < LA; ++s) {
            const int rmin = D_ + s * D_, rmax = rmin + D_ - 1, r = rmin + gl;
            const R* pLs_ = ROWPACK ? Sb + r * (r + 1) / 2 : Sb + r;
            const R* pUs_ = ROWPACK ? Sb + r : Sb + col_start(n, r) - r;
#pragma unroll
            for (int e = 0; e < n; ++e) {
                const int c = e < NO ? D_ + e : e - NO;             // column of S
                R val;
                if (c <= rmin) val = pLs_[offL(c)];
                else if (c > rmax) val = pUs_[offU(c)];
                else val = (c <= r) ? pLs_[offL(c)] : pUs_[offU(c)];
                if (c >= rmin && c <= rmax) val += (c == r) ? eps1 : (R)0;
                m[s][e] = val;
            }
        }
#pragma unroll
        for (int j = 0; j < NO; ++j) {
            constexpr int dummy = 0; (void)dummy;
            const int sj = j / D_, oj = j % D_;
            R* pr_ = piv + (j & 1) * NP;
            if (gl == oj) {
#pragma unroll
                for (int cv = 0; cv < NV; ++cv) {
                    VecT ov;
                    R* oe = reinterpret_cast<R*>(&ov);
#pragma unroll
                    for (int q = 0; q < VEC; ++q) oe[q] = (cv * VEC + q < n) ? m[sj][cv * VEC + q] : (R)0;
                    *reinterpret_cast<VecT*>(pr_ + cv * VEC) = ov;
                }
            }
            __syncwarp();
            R prow[NP];
#pragma unroll
            for (int cv = 0; cv < NV; ++cv) {
                const VecT lv = *reinterpret_cast<const VecT*>(pr_ + cv * VEC);
                const R* le = reinterpret_cast<const R*>(&lv);
#pragma unroll
                for (int q = 0; q < VEC; ++q) prow[cv * VEC + q] = le[q];
            }
            const R ip = rcp_fast<R>(prow[j]);
#pragma unroll
            for (int s = 0; s < LA; ++s) {
                const R coef = m[s][j] * ip;              // ordinary row: m -= (f ip) prow, m[j] = -f ip
                const R ncoef = -coef;
#pragma unroll
                for (int c = 0; c + 1 < n; c += 2) {
                    if (c == j || c + 1 == j) {
                        if (c != j) m[s][c] = fma(ncoef, prow[c], m[s][c]);
                        if (c + 1 != j) m[s][c + 1] = fma(ncoef, prow[c + 1], m[s][c + 1]);
                    } else {
                        fma2<R>(m[s][c], m[s][c + 1], ncoef, ncoef, prow[c], prow[c + 1]);
                    }
                }
                if ((n & 1) && n - 1 != j) m[s][n - 1] = fma(ncoef, prow[n - 1], m[s][n - 1]);
                m[s][j] = ncoef;
            }
            if (gl == oj) {                               // the pivot row itself: prow / pivot, 1 / pivot on the diagonal
#pragma unroll
                for (int c = 0; c + 1 < n; c += 2) mul2<R>(m[sj][c], m[sj][c + 1], prow[c], prow[c + 1], ip);
                if (n & 1) m[sj][n - 1] = prow[n - 1] * ip;
                m[sj][j] = ip;
            }
        }
        // m[s] = [Z row | T' row]; turn the Z part into F = I - eps1 Z
#pragma unroll
        for (int s = 0; s < LA; ++s) {
#pragma unroll
            for (int e = 0; e < NO; ++e) {
                const bool diag_blk = (e >= s * D_ && e < (s + 1) * D_);
                const R one = diag_blk ? ((e - s * D_ == gl) ? (R)1 : (R)0) : (R)0;
                m[s][e] = fma(-eps1, m[s][e], one);
            }
        }
        // publish T' rows
#pragma unroll
        for (int s = 0; s < LA; ++s) {
            R* dst = TTs + (gl + s * D_) * DP;
#pragma unroll
            for (int cv = 0; cv < DV; ++cv) {
                VecT ov;
                R* oe = reinterpret_cast<R*>(&ov);
#pragma unroll
                for (int q = 0; q < VEC; ++q) oe[q] = (cv * VEC + q < D_) ? m[s][NO + cv * VEC + q] : (R)0;
                *reinterpret_cast<VecT*>(dst + cv * VEC) = ov;
            }
        }
        __syncwarp();

        // ---- phase D: a-row of Sigma1: [S_aa - S_ac T' | eps1 T row]; u1 = K1 m_c for the owned rows
        R s1a[D_], u1[L_];
        const R* pL0 = ROWPACK ? Sb + gl * (gl + 1) / 2 : Sb + gl;
        const R* pU0 = ROWPACK ? Sb + gl : Sb + col_start(n, gl) - gl;
#pragma unroll
        for (int a = 0; a < D_; ++a) s1a[a] = (a <= gl) ? pL0[offL(a)] : pU0[offU(a)];
        u1[0] = (R)0;
#pragma unroll
        for (int e = 0; e < NO; ++e) {
            const R sac = pU0[offU(D_ + e)];                        // S[gl][d + e], always above the diagonal
            const R tg = TTs[e * DP + gl];                          // T[gl][e]
            u1[0] = fma(tg, mv[D_ + e], u1[0]);
            const R nsac = -sac;
#pragma unroll
            for (int cv = 0; cv < DV; ++cv) {
                const VecT lv = *reinterpret_cast<const VecT*>(TTs + e * DP + cv * VEC);
                const R* le = reinterpret_cast<const R*>(&lv);
#pragma unroll
                for (int q = 0; q < VEC; q += 2) {
                    const int a = cv * VEC + q;
                    if (a + 1 < D_) fma2<R>(s1a[a], s1a[a + 1], nsac, nsac, le[q], le[q + 1]);
                    else if (a < D_) s1a[a] = fma(nsac, le[q], s1a[a]);
                }
            }
        }
        __syncwarp();                                 // every lane is done with the packed covariance

        mbar_wait(barA, parA);
        parA ^= 1;
        // ---- phase E: W' rows = [Sigma1 a-row ; K1' rows] x A'   (wt[0] = Wa' row, wt[s+1] = Wc'' rows)
        R wt[L_][D_];
#pragma unroll
        for (int s = 0; s < L_; ++s)
#pragma unroll
            for (int a = 0; a < D_; ++a) wt[s][a] = (R)0;
#pragma unroll
        for (int e = 0; e < n; ++e) {
            R arow[DP];
#pragma unroll
            for (int cv = 0; cv < DV; ++cv) {
                const VecT lv = *reinterpret_cast<const VecT*>(Ats + e * DP + cv * VEC);
                const R* le = reinterpret_cast<const R*>(&lv);
#pragma unroll
                for (int q = 0; q < VEC; ++q) arow[cv * VEC + q] = le[q];
            }
#pragma unroll
            for (int s = 0; s < L_; ++s) {
                // coefficient Sigma1'[row][e] in the column order [a | c]
                // (the c part of the a-row, eps1 T[gl][:], is read back from the T' rows instead of being kept)
                const R cf = (s == 0) ? (e < D_ ? s1a[e < D_ ? e : 0] : eps1 * TTs[(e >= D_ ? e - D_ : 0) * DP + gl])
                                      : (e < D_ ? m[s > 0 ? s - 1 : 0][NO + (e < D_ ? e : 0)] : m[s > 0 ? s - 1 : 0][e >= D_ ? e - D_ : 0]);
#pragma unroll
                for (int a = 0; a + 1 < D_; a += 2) fma2<R>(wt[s][a], wt[s][a + 1], cf, cf, arow[a], arow[a + 1]);
                if (D_ & 1) wt[s][D_ - 1] = fma(cf, arow[D_ - 1], wt[s][D_ - 1]);
            }
        }
#pragma unroll
        for (int s = 0; s < L_; ++s) {
            R* dst = WTs + (gl + s * D_) * DP;
#pragma unroll
            for (int cv = 0; cv < DV; ++cv) {
                VecT ov;
                R* oe = reinterpret_cast<R*>(&ov);
#pragma unroll
                for (int q = 0; q < VEC; ++q) oe[q] = (cv * VEC + q < D_) ? wt[s][cv * VEC + q] : (R)0;
                *reinterpret_cast<VecT*>(dst + cv * VEC) = ov;
            }
        }
        __syncwarp();
        if (LOCKSTEP >= 2) __syncthreads();

        // ---- phase G: B2 row gl = Q'[gl] + sum_r W[gl][r] A'[r];  t[gl] = (A m + b)[gl] - (Wc' m_c)[gl]
        R b2[D_];
#pragma unroll
        for (int cv = 0; cv < DV; ++cv) {
            const VecT lv = *reinterpret_cast<const VecT*>(Qs + gl * DP + cv * VEC);
            const R* le = reinterpret_cast<const R*>(&lv);
#pragma unroll
            for (int q = 0; q < VEC; ++q)
                if (cv * VEC + q < D_) b2[cv * VEC + q] = le[q];
        }
        R tacc = bvec[gl];
#pragma unroll
        for (int r = 0; r < n; ++r) {
            const R wcol = WTs[r * DP + gl];                        // W'[r][gl]
            const R mr = mv[r];
            tacc = fma(Ats[r * DP + gl], mr, tacc);
            if (r >= D_) tacc = fma(-wcol, mr, tacc);
            const R wv_ = (r < D_) ? wcol : eps1 * wcol;            // W[gl][r]
#pragma unroll
            for (int cv = 0; cv < DV; ++cv) {
                const VecT lv = *reinterpret_cast<const VecT*>(Ats + r * DP + cv * VEC);
                const R* le = reinterpret_cast<const R*>(&lv);
#pragma unroll
                for (int q = 0; q < VEC; q += 2) {
                    const int a = cv * VEC + q;
                    if (a + 1 < D_) fma2<R>(b2[a], b2[a + 1], wv_, wv_, le[q], le[q + 1]);
                    else if (a < D_) b2[a] = fma(wv_, le[q], b2[a]);
                }
            }
        }
        tv[gl] = tacc;

        // ---- phase H: L2 = chol(B2) across the d lanes of the frame (row gl per lane), inverse pivots in every lane
        R invd[D_];
#pragma unroll
        for (int j = 0; j < D_; ++j) {
            const R dj = __shfl_sync(FULL, b2[j], gbase + j);
            const R inv = rsqrt_fast<R>(dj);
            invd[j] = inv;
            const R l = (gl >= j) ? b2[j] * inv : (R)0;
            b2[j] = l;
#pragma unroll
            for (int c = j + 1; c < D_; ++c) {
                const R lc = __shfl_sync(FULL, l, gbase + c);       // L2[c][j]
                b2[c] = fma(-l, lc, b2[c]);
            }
        }
        {
            R* dst = B2s + gl * DP;
#pragma unroll
            for (int cv = 0; cv < DV; ++cv) {
                VecT ov;
                R* oe = reinterpret_cast<R*>(&ov);
#pragma unroll
                for (int q = 0; q < VEC; ++q) oe[q] = (cv * VEC + q < D_) ? b2[cv * VEC + q] : (R)0;
                *reinterpret_cast<VecT*>(dst + cv * VEC) = ov;
            }
        }
        __syncwarp();                                 // L2 rows, t and every lane's reads of the W' columns are complete
        load_wN();
        onN = iN < Tx - 1 && mkN != 0;               // first use of the words loaded at the top of the iteration
        onmaskN = __ballot_sync(FULL, onN && live && gl == 0);
        request_A(ziN, onN, onmaskN);                 // the operator block is consumed: fetch the next frame's
        // V2 rows = W rows L2^-T (forward substitution), K2 rows = V2 rows L2^-1 (back substitution); the K2 rows go
        // straight to K2', to the last d rows of G' and into the offset of the mean
        R hk[L_];
        {
        R v2[L_][D_];
#pragma unroll
        for (int c = 0; c < D_; ++c) {                // forward substitution, every owned row at once
            R acc[L_];
#pragma unroll
            for (int s = 0; s < L_; ++s) acc[s] = (s == 0) ? wt[s][c] : eps1 * wt[s][c];
#pragma unroll
            for (int p2 = 0; p2 < c; ++p2) {
                const R nl = -B2s[c * DP + p2];
#pragma unroll
                for (int s = 0; s + 1 < L_; s += 2) fma2<R>(acc[s], acc[s + 1], nl, nl, v2[s][p2], v2[s + 1][p2]);
                if (L_ & 1) acc[L_ - 1] = fma(nl, v2[L_ - 1][p2], acc[L_ - 1]);
            }
#pragma unroll
            for (int s = 0; s < L_; ++s) v2[s][c] = acc[s] * invd[c];
        }
#pragma unroll
        for (int s0 = 0; s0 < L_; s0 += 2) {          // back substitution, two owned rows at a time (packed)
            constexpr int dummy2 = 0; (void)dummy2;
            const bool two = s0 + 1 < L_;
            R k2[2][D_];
#pragma unroll
            for (int c = D_ - 1; c >= 0; --c) {
                R a0 = v2[s0][c], a1 = two ? v2[two ? s0 + 1 : s0][c] : (R)0;
#pragma unroll
                for (int p2 = c + 1; p2 < D_; ++p2) {
                    const R nl = -B2s[p2 * DP + c];
                    if (two) fma2<R>(a0, a1, nl, nl, k2[0][p2], k2[1][p2]);
                    else a0 = fma(nl, k2[0][p2], a0);
                }
                k2[0][c] = a0 * invd[c];
                k2[1][c] = a1 * invd[c];
            }
#pragma unroll
            for (int t = 0; t < 2; ++t) {
                if (t == 1 && !two) continue;
                const int sidx = s0 + t, r = gl + sidx * D_;
                R acc = 0;
#pragma unroll
                for (int e = 0; e < D_; ++e) {
                    K2T[e * NP + r] = k2[t][e];
                    acc = fma(k2[t][e], tv[e], acc);
                    if (on && live) Gout[(size_t)(NO + e) * n + r] = k2[t][e];
                }
                hk[sidx] = acc;
            }
        }
        // V2 rows over the W' rows (their columns were consumed before the barrier above)
#pragma unroll
        for (int s = 0; s < L_; ++s) {
            R* dst = WTs + (gl + s * D_) * DP;
#pragma unroll
            for (int cv = 0; cv < DV; ++cv) {
                VecT ov;
                R* oe = reinterpret_cast<R*>(&ov);
#pragma unroll
                for (int q = 0; q < VEC; ++q) oe[q] = (cv * VEC + q < D_) ? v2[s][cv * VEC + q] : (R)0;
                *reinterpret_cast<VecT*>(dst + cv * VEC) = ov;
            }
        }
        }                                             // v2 leaves the registers here: phase I reads the owned rows back
        __syncwarp();

        // ---- phase J: G1' rows rc: K1'[rc][:] - Wc''[rc][:] K2'  (column chunks, straight to global memory)
        {
            constexpr int CH = 2 * VEC;               // columns per chunk
#pragma unroll
            for (int c0 = 0; c0 < n; c0 += CH) {
                R acc[LA][CH];
#pragma unroll
                for (int s = 0; s < LA; ++s)
#pragma unroll
                    for (int q = 0; q < CH; ++q) {
                        const int c = c0 + q;       // column of G1' = coordinate of xi_t, order [a | c]
                        acc[s][q] = (c < n) ? (c < D_ ? m[s][NO + (c < D_ ? c : 0)] : m[s][c >= D_ && c < n ? c - D_ : 0]) : (R)0;
                    }
#pragma unroll
                for (int e = 0; e < D_; ++e) {
                    R kv[CH];
#pragma unroll
                    for (int cv = 0; cv < 2; ++cv) {
                        if (c0 + cv * VEC < NP) {
                            const VecT lv = *reinterpret_cast<const VecT*>(K2T + e * NP + c0 + cv * VEC);
                            const R* le = reinterpret_cast<const R*>(&lv);
#pragma unroll
                            for (int q = 0; q < VEC; ++q) kv[cv * VEC + q] = le[q];
                        } else {
#pragma unroll
                            for (int q = 0; q < VEC; ++q) kv[cv * VEC + q] = (R)0;
                        }
                    }
#pragma unroll
                    for (int s = 0; s < LA; ++s) {
                        const R nw = -wt[s + 1][e];
#pragma unroll
                        for (int q = 0; q < CH; q += 2) fma2<R>(acc[s][q], acc[s][q + 1], nw, nw, kv[q], kv[q + 1]);
                    }
                }
                if (on && live) {
#pragma unroll
                    for (int s = 0; s < LA; ++s) {
                        R* dst = Gout + (size_t)(gl + s * D_) * n + c0;
#pragma unroll
                        for (int q = 0; q < CH; ++q)
                            if (c0 + q < n) dst[q] = acc[s][q];
                    }
                }
            }
        }

        if (LOCKSTEP >= 3) __syncthreads();
#pragma unroll
        for (int s = 0; s < LA; ++s) {                // u1 = F m_c for the owned c rows (F still unscaled)
            R acc = 0;
#pragma unroll
            for (int e = 0; e < NO; ++e) acc = fma(m[s][e], mv[D_ + e], acc);
            u1[s + 1] = acc;
        }
        // ---- phase I: Sigma rows (lower triangle only: slot s needs columns < (s+1) d), natural column order [a | c]
        R sg[L_][n];
#pragma unroll
        for (int c = 0; c < D_; ++c) sg[0][c] = s1a[c];
#pragma unroll
        for (int s = 1; s < L_; ++s)
#pragma unroll
            for (int c = 0; c < (s + 1) * D_; ++c) sg[s][c] = eps1 * (c < D_ ? m[s - 1][NO + (c < D_ ? c : 0)] : m[s - 1][c >= D_ ? c - D_ : 0]);
        R v2[L_][D_];                                 // owned V2 rows, back from shared memory
#pragma unroll
        for (int s = 0; s < L_; ++s)
#pragma unroll
            for (int cv = 0; cv < DV; ++cv) {
                const VecT lv = *reinterpret_cast<const VecT*>(WTs + (gl + s * D_) * DP + cv * VEC);
                const R* le = reinterpret_cast<const R*>(&lv);
#pragma unroll
                for (int q = 0; q < VEC; ++q)
                    if (cv * VEC + q < D_) v2[s][cv * VEC + q] = le[q];
            }
#pragma unroll
        for (int c = 0; c < n; ++c) {
            R vrow[DP];
#pragma unroll
            for (int cv = 0; cv < DV; ++cv) {
                const VecT lv = *reinterpret_cast<const VecT*>(WTs + c * DP + cv * VEC);
                const R* le = reinterpret_cast<const R*>(&lv);
#pragma unroll
                for (int q = 0; q < VEC; ++q) vrow[cv * VEC + q] = le[q];
            }
#pragma unroll
            for (int s = 0; s < L_; ++s) {
                if (c < (s + 1) * D_) {
                    R a0 = 0, a1 = 0;
#pragma unroll
                    for (int e = 0; e + 1 < D_; e += 2) fma2<R>(a0, a1, v2[s][e], v2[s][e + 1], vrow[e], vrow[e + 1]);
                    if (D_ & 1) a0 = fma(v2[s][D_ - 1], vrow[D_ - 1], a0);
                    sg[s][c] -= a0 + a1;
                }
            }
        }
        __syncwarp();                                 // V2 rows consumed: the ring below reuses the pivot rows' place
        request_S(gN, onN, onmaskN, buf ^ 1);         // ... and the next frame's covariance takes theirs
        // ---- Ls = chol(Sigma), right-looking; column j is published in a two-row ring and read back by every lane
#pragma unroll
        for (int j = 0; j < n; ++j) {
            const int sj = j / D_, oj = j % D_;
            const R dj = __shfl_sync(FULL, sg[sj][j], gbase + oj);
            const R inv = rsqrt_fast<R>(dj);
            R* ring = piv + (j & 1) * NP;
            R l[L_];
#pragma unroll
            for (int s = 0; s < L_; ++s) {
                l[s] = (R)0;
                if (s >= sj) {
                    const int r = gl + s * D_;
                    l[s] = (r >= j) ? sg[s][j] * inv : (R)0;
                    sg[s][j] = l[s];
                    ring[r] = l[s];
                }
            }
            if (j + 1 < n) {
                R nl[L_];
#pragma unroll
                for (int s = 0; s < L_; ++s) nl[s] = -l[s];
                __syncwarp();
#pragma unroll
                for (int cv = (j + 1) / VEC; cv < NV; ++cv) {
                    const VecT lv = *reinterpret_cast<const VecT*>(ring + cv * VEC);
                    const R* le = reinterpret_cast<const R*>(&lv);
#pragma unroll
                    for (int q = 0; q < VEC; q += 2) {
                        const int c = cv * VEC + q;
#pragma unroll
                        for (int s = 0; s < L_; ++s) {
                            if (s < sj) continue;
                            const bool ok0 = c > j && c < (s + 1) * D_, ok1 = c + 1 > j && c + 1 < (s + 1) * D_;
                            if (ok0 && ok1) fma2<R>(sg[s][c], sg[s][c + 1], nl[s], nl[s], le[q], le[q + 1]);
                            else if (ok0) sg[s][c] = fma(nl[s], le[q], sg[s][c]);
                            else if (ok1) sg[s][c + 1] = fma(nl[s], le[q + 1], sg[s][c + 1]);
                        }
                    }
                }
            }
        }
        // ---- h = m - K1 m_c - K2 t + Ls w
#pragma unroll
        for (int s = 0; s < L_; ++s) {
            const int r = gl + s * D_;
            R acc = 0;
#pragma unroll
            for (int c = 0; c < (s + 1) * D_; ++c) acc = fma((c <= r) ? sg[s][c] : (R)0, wv[c], acc);
            if (on && live) Gout[(size_t)n * n + r] = mv[r] - u1[s] - hk[s] + acc;
        }
        }                                             // onmask != 0
        if (onN) {
#pragma unroll
            for (int s = 0; s < L_; ++s) wvb[(buf ^ 1) * NP + gl + s * D_] = wN[s];
        }
        __syncwarp();
        g = gN;
        on = onN;
        zi = ziN;
        onmask = onmaskN;
    }
}
