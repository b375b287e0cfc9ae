// K1b for 32 < n <= 64 (latent_dim 11 .. 16 at nlags 3, latent_dim 10 at nlags 4), float32: the row-per-lane filter of
// kalman.cu (kalman_forward_rows_kernel) with TWO warps per (chain, time chunk).  Lane tl = 32 * (warp in team) + lane
// owns row tl of the predicted covariance in registers; the step has the same phases and the same arithmetic order,
// with three differences forced by the team being wider than a warp:
//   * the team synchronises on a named barrier (bar.sync id, 64) instead of __syncwarp;
//   * what the one-warp kernel moves by shuffles crosses warps here and goes through shared memory: the shifted blocks
//     of P+ (a row tile Ps, n x (n | 1)), the per-block partial sums of A P+ A' (Pt) and the shifted mean (ms);
//   * chol(J_t) is read from the record in shared memory where it is used instead of being copied to registers first
//     (d (d + 1) / 2 = 136 numbers at latent_dim 16).
// Same covariance-form update as the one-warp kernel - P+ = P - V V' from the current prediction, then A P+ and
// A P+ A' from P+ - and not the shared-memory filter's A P A' - (A V)(A V)', whose cancellation costs float32 a digit
// (tests/test_gpu_parity.py::test_full_sweep_every_compiled_pair).  The covariance record is packed by COLUMNS like the
// one-warp kernel's (the float64 path for n > 32 keeps the shared-memory filter and its row packing).
// Included by kalman.cu inside namespace kpms.
#pragma once

template <typename R, int D_, int L_>
struct FwdRows2wSmem {
    static constexpr int n = D_ * L_, NP = D_ * (D_ + 1) / 2, NPP = (NP + 3) / 4 * 4, RECI = info_stride(D_);
    static constexpr int TL = 64;                                  // lanes of a team
    static constexpr int QO = (n + 1 + 3) / 4 * 4;                 // offset of the Q row inside an A row
    static constexpr int AS = QO + (D_ + 3) / 4 * 4;               // row: [A (n) | b | pad | Q row (d)]
    static constexpr int VS = (D_ + 3) / 4 * 4, APS = TL + 4, STAGES = 4;
    static constexpr int BS = NPP + VS;
    static constexpr int PST = n | 1;                              // row stride of the P+ tile (odd: conflict-free)
    static constexpr int DPT = D_ | 1;                             // row stride of the partial sums
    static constexpr size_t per_team = (STAGES * RECI + 2 * D_ * AS + TL * VS + D_ * APS + BS + TL + n * PST + n * DPT + 3) / 4 * 4;
};

__device__ __forceinline__ void team_sync(int id) { asm volatile("bar.sync %0, 64;\n" ::"r"(id) : "memory"); }

template <typename R, int D_, int L_, int TEAMS>
__global__ void __launch_bounds__(64 * TEAMS, 1)
kalman_forward_rows2w_kernel(const R* __restrict__ info, const int* __restrict__ mask, const int* __restrict__ z,
                             const R* __restrict__ Ab, const R* __restrict__ Q, R jitter, int N, int T,
                             R* __restrict__ stash_m, R* __restrict__ stash_S, int C, int W,
                             const int* __restrict__ vlen, const int* __restrict__ vend, const int* __restrict__ dirty,
                             R* __restrict__ bnd_warm, R* __restrict__ bnd_end) {
    typedef FwdRows2wSmem<R, D_, L_> SM;
    typedef typename Vec16<R>::type VecT;
    constexpr int n = SM::n, NO = n - D_, NP = SM::NP, NPP = SM::NPP, RECI = SM::RECI, AS = SM::AS, QO = SM::QO,
                  VS = SM::VS, APS = SM::APS, STAGES = SM::STAGES, TL = SM::TL, PST = SM::PST, DPT = SM::DPT;
    constexpr int VEC = 16 / (int)sizeof(R), NV = (n + VEC - 1) / VEC, DV = (D_ + VEC - 1) / VEC;
    constexpr int SMS = stash_m_stride(n), SSS = stash_S_stride(n), BREC = n + n * n;
    static_assert(n > 32 && n <= 64 && sizeof(R) == 4, "two warps, one covariance row per lane, float32");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int team = threadIdx.x >> 6, tl = threadIdx.x & 63;
    const int bar = 1 + team;                        // named barrier of this team (0 is __syncthreads)
    const long long task = (long long)blockIdx.x * TEAMS + team;
    if (task >= (long long)N * C) return;
    const int nn = (int)(task / C), ck = (int)(task % C);
    const int Tx = T - L_ + 1;
    if (dirty && dirty[nn] == 0) return;
    const ChunkRange cr = chunk_range(vlen ? vlen[nn] : Tx, Tx, C, W, ck);
    if (cr.empty) return;
    R* ring = reinterpret_cast<R*>(smem_raw) + (size_t)team * SM::per_team;   // STAGES x RECI
    R* Asb = ring + STAGES * RECI;                   // 2 x D_ x AS
    R* Vs = Asb + 2 * D_ * AS;                       // TL x VS   (U rows, then V rows)
    R* APs = Vs + TL * VS;                           // D_ x APS  (A P+)
    R* Bs = APs + D_ * APS;                          // [B lower packed | nu]
    R* ms = Bs + SM::BS;                             // TL
    R* Ps = ms + TL;                                 // n x PST   (P+ rows, for the block shift)
    R* Pt = Ps + n * PST;                            // n x DPT   (partial sums of A P+ A' per row)
    const bool act = tl < n;
    const int row = act ? tl : n - 1;                // idle lanes shadow the last row
    const R* inf_g = info + (size_t)nn * Tx * RECI;
    const int* mk = mask + (size_t)nn * T + (L_ - 1);
    const int* zz = z + (size_t)nn * (Tx - 1);
    R* sm_g = stash_m + (size_t)nn * Tx * SMS;
    R* sS_g = stash_S + (size_t)nn * Tx * SSS;
    const R eps = (R)KPMS_EPS_SHIFT + jitter;
    constexpr int NQ = (NP + TL - 1) / TL;           // entries of B (lower, packed by rows) computed by this lane
    int ba[NQ], bc[NQ];
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
        int a = 0, c = 0;
        const int idx = tl + TL * q;
        if (idx < NP) tri_unpack(idx, a, c);
        ba[q] = a;
        bc[q] = c;
    }
    // frames from vend[nn] on are all masked: they carry (m, P) unchanged, so the walk stops there and the state is
    // stored once at the terminal frame (a short row's last chunk used to step through thousands of padded frames)
    const int i0 = cr.start, i1 = cr.end, i_stop = min(i1, vend[nn]);
    auto issue_info = [&](int i) {
        if (i < i_stop)
            for (int c = tl; c < RECI * (int)sizeof(R) / 16; c += TL)
                cp_async_16(reinterpret_cast<char*>(ring + (i % STAGES) * RECI) + 16 * c,
                            reinterpret_cast<const char*>(inf_g + (size_t)i * RECI) + 16 * c);
        asm volatile("cp.async.commit_group;\n" ::);
    };
    auto load_A = [&](int zi, int buf) {             // joins the next committed group
        R* dst = Asb + buf * D_ * AS;
        const R* A = Ab + (size_t)zi * D_ * (n + 1);
        for (int w = tl; w < D_ * (n + 1); w += TL) cp_async_elem(dst + (w / (n + 1)) * AS + (w % (n + 1)), A + w);
        const R* Qk = Q + (size_t)zi * D_ * D_;
        for (int w = tl; w < D_ * D_; w += TL) cp_async_elem(dst + (w / D_) * AS + QO + (w % D_), Qk + w);
    };
    R p[n], m = (R)0;
#pragma unroll
    for (int c = 0; c < n; ++c) p[c] = (c == row) ? (R)KPMS_X_PRIOR_VAR : (R)0;
    for (int w = tl; w < 2 * D_ * AS; w += TL) Asb[w] = (R)0;     // padding is multiplied by masked zeros
    team_sync(bar);
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
        team_sync(bar);
        if (ck > 0 && i == cr.begin && act) {        // the state this chunk arrived with
            R* bw = bnd_warm + ((size_t)nn * C + ck) * BREC;
            bw[tl] = m;
#pragma unroll
            for (int c = 0; c < n; ++c) bw[n + tl * n + c] = p[c];
        }
        const R* fi = ring + (i % STAGES) * RECI;
        const R* A = Asb + cur * D_ * AS;
        if (mk_cur != 0) {
            // ---- U = P[:,new] Lj (own row); publish U rows and the mean
            R u[D_];
#pragma unroll
            for (int c = 0; c < D_; ++c) {
                R acc = 0;
#pragma unroll
                for (int e = c; e < D_; ++e) acc = fma(p[NO + e], fi[e * (e + 1) / 2 + c], acc);
                u[c] = acc;
            }
            ms[tl] = m;
#pragma unroll
            for (int cv = 0; cv < DV; ++cv) {
                VecT ov;
                R* oe = reinterpret_cast<R*>(&ov);
#pragma unroll
                for (int q = 0; q < VEC; ++q) oe[q] = (cv * VEC + q < D_) ? u[cv * VEC + q] : (R)0;
                *reinterpret_cast<VecT*>(Vs + tl * VS + cv * VEC) = ov;
            }
            team_sync(bar);
            // ---- B = I + Lj' U[new,:] (lower) and nu = y~ - Lj' m_new, spread over the lanes
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
                const int a = ba[q], c = bc[q];
                R acc = (a == c) ? (R)1 : (R)0;
#pragma unroll
                for (int e = 0; e < D_; ++e)
                    if (e >= a) acc = fma(fi[e * (e + 1) / 2 + a], Vs[(NO + e) * VS + c], acc);
                if (tl + TL * q < NP) Bs[tl + TL * q] = acc;
            }
            if (tl < D_) {
                R acc = fi[NP + D_ + tl];
#pragma unroll
                for (int e = 0; e < D_; ++e)
                    if (e >= tl) acc = fma(-fi[e * (e + 1) / 2 + tl], ms[NO + e], acc);
                Bs[NPP + tl] = acc;
            }
            team_sync(bar);
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
                *reinterpret_cast<VecT*>(Vs + tl * VS + cv * VEC) = ov;
            }
            team_sync(bar);
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
                sm_g[(size_t)i * SMS + tl] = m;
                R* so = sS_g + (size_t)i * SSS + tl;
#pragma unroll
                for (int c = 0; c < n; ++c)
                    if (c <= tl) so[col_start(n, c) - c] = p[c];
            }
            if (!last) {
                // ---- A P+ (column `row`), published by rows of A; P+ rows published for the block shift
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
                    APs[a * APS + tl] = ap[a];
                }
                ms[tl] = m;
                if (act && row >= D_) {              // rows below the oldest block are read by row - d, columns >= d only
#pragma unroll
                    for (int c = D_; c < n; ++c) Ps[row * PST + c] = p[c];
                }
                team_sync(bar);
                // ---- next mean: shifted blocks from the published means, newest block = A m+ + b
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
                    const R mshift = ms[row < NO ? row + D_ : row];
                    m = (row < NO) ? mshift : (acc0 + acc1);
                }
                // ---- A P+ A': lane (grp, a' = row - grp d) sums its block of the contraction for every column a and
                //      parks the partial sums in Pt; the lanes of the last block add the L partial sums of their row
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
                        if (act) Pt[row * DPT + a] = acc + acc1;
                    }
                    team_sync(bar);
#pragma unroll
                    for (int a = 0; a < D_; ++a) {
                        R tot = Pt[(NO + arow) * DPT + a];              // own block first, then the older ones (shuffle order of the one-warp kernel)
#pragma unroll
                        for (int gq = 1; gq < L_; ++gq) tot += Pt[(NO - gq * D_ + arow) * DPT + a];
                        apa[a] = tot;                                   // meaningful for rows >= NO
                    }
                    // Keep the covariance symmetric to the last bit (see the one-warp kernel)
                    if (act && row >= NO) {
#pragma unroll
                        for (int a = 0; a < D_; ++a) Vs[arow * VS + a] = apa[a];
                    }
                    team_sync(bar);
#pragma unroll
                    for (int a = 0; a < D_; ++a) apa[a] = (R)0.5 * (apa[a] + Vs[a * VS + arow]);
                }
                // ---- next predicted covariance: rows < NO are shifted rows of P+ / columns of A P+ (row r + d),
                //      rows >= NO are A P+ and A P+ A' + Q
                {
                    const R* aprow = APs + arow * APS;
                    const R* prow = Ps + (row < NO ? row + D_ : row) * PST;
                    R pn[n];
#pragma unroll
                    for (int c = 0; c < NO; ++c) {
                        const R sh = prow[c + D_];
                        pn[c] = (row < NO) ? sh + ((row == c) ? eps : (R)0) : aprow[c + D_];
                    }
                    const R* qrow = A + arow * AS + QO;
#pragma unroll
                    for (int a = 0; a < D_; ++a) {
                        const R sh = APs[a * APS + (row < NO ? row + D_ : row)];
                        const R nw = apa[a] + qrow[a] + ((arow == a) ? jitter : (R)0);
                        pn[NO + a] = (row < NO) ? sh : nw;
                    }
#pragma unroll
                    for (int c = 0; c < n; ++c) p[c] = pn[c];
                }
            }
        } else if (last && keep && act) {
            sm_g[(size_t)i * SMS + tl] = m;
            R* so = sS_g + (size_t)i * SSS + tl;
#pragma unroll
            for (int c = 0; c < n; ++c)
                if (c <= tl) so[col_start(n, c) - c] = p[c];
        }
        if (change) { cur ^= 1; zc = z_next; }
        changed_prev = change;
        mk_cur = mk_next;
        team_sync(bar);
    }
    asm volatile("cp.async.wait_all;\n" ::);
    if (i1 == Tx && i_stop < Tx && act) {            // masked terminal frame: the carried state (Tx - 1 >= cr.begin)
        sm_g[(size_t)(Tx - 1) * SMS + tl] = m;
        R* so = sS_g + (size_t)(Tx - 1) * SSS + tl;
#pragma unroll
        for (int c = 0; c < n; ++c)
            if (c <= tl) so[col_start(n, c) - c] = p[c];
    }
    if (i1 < Tx && act) {                            // the state handed to the next chunk
        R* be = bnd_end + ((size_t)nn * C + ck + 1) * BREC;
        be[tl] = m;
#pragma unroll
        for (int c = 0; c < n; ++c) be[n + tl * n + c] = p[c];
    }
}
