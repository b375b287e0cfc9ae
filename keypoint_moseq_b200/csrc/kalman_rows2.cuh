// K1c, register-blocked one-stage form: backward preparation with TWO matrix rows per lane.  Round 1's default for
// every shape; since round 2 the two-stage kernel (kalman_split.cuh) handles nlags >= 2 and this one nlags = 1.
//
// The one-row-per-lane kernel it replaced (kalman_backprep_rows_kernel, removed) was bound by shared-memory operand
// traffic: every FMA consumed one operand that another lane published, delivered by 16-byte
// broadcasts at two floats per lane per wavefront, so the shared-memory pipe (1 wavefront / clk)
// saturated at half the FMA rate (ncu: 1 750 wavefronts per frame, 60 % of the run time).  Here a lane
// owns rows gl and gl + NH (NH = ceil(n/2)) of its frame, so every broadcast operand feeds two FMAs,
// and a warp carries FPW = 32 / GW frames (GW = lanes per frame, the power of two >= NH): n = 30 ->
// 15 active lanes per frame, two frames per warp; n = 12 -> four frames per warp; n = 48 -> one.
// Same algebra, phases and record layout as the one-row kernel:
//   Wt = Aaug S, Pp = Wt Aaug' + Qaug, Lp = chol(Pp) fused with V = Lp^-1 Wt, Sigma = S - V'V,
//   Ls = chol(Sigma), GT = Lp^-T V, h = m - G (Aaug m + b) + Ls w.
// Included by kalman.cu inside namespace kpms.
#pragma once

template <typename R, int D_, int L_>
struct PrepRows2 {
    static constexpr int n = D_ * L_, NH = (n + 1) / 2;
    static constexpr int GW = NH <= 4 ? 4 : NH <= 8 ? 8 : NH <= 16 ? 16 : 32;   // lanes per frame
    static constexpr int FPW = 32 / GW;                                          // frames per warp
    static constexpr int LS0 = (n + 1 + 3) / 4 * 4;            // the [A | b] rows need n + 1 entries
    static constexpr int LS = ((LS0 / 4) & 1) ? LS0 : LS0 + 4;    // row stride: LS/4 odd -> conflict-free 16-byte row stores
    static constexpr int NP = (n + 3) / 4 * 4;
    static constexpr int SB = stash_S_stride(n);
    static constexpr int RAW = 2 * n * LS + D_ * LS + SB + 6 * NP;
    // frames of one warp sit 32/FPW banks apart, so a broadcast with FPW distinct addresses is conflict-free
    static constexpr int SHIFT = FPW > 1 ? 32 / FPW : 0;
    static constexpr int per_group = FPW > 1 ? RAW + ((SHIFT - RAW % 32) + 32) % 32 : RAW;
};

// Right-looking Cholesky of the matrix whose rows r0 / r1 are in a0[] / a1[]; column j is published
// as row j of LT.  On return a0[c] (c <= r0) holds L[r0][c], likewise a1.  With SOLVE, the columns
// c0 / c1 are forward-substituted in the same sweep and the inverse pivots are kept in invd.
template <typename R, int n, int NH, int LS, bool SOLVE>
__device__ __forceinline__ void chol_rows2(R (&a0)[n], R (&a1)[n], R* LT, R (&c0)[n], R (&c1)[n], R* invd,
                                           int r0, int r1, int gbase) {
    typedef typename Vec16<R>::type VecT;
    constexpr int VEC = 16 / (int)sizeof(R), NV = (n + VEC - 1) / VEC;
#pragma unroll
    for (int j = 0; j < n; ++j) {
        const R dj = __shfl_sync(0xffffffffu, (j < NH) ? a0[j] : a1[j], gbase + (j < NH ? j : j - NH));
        const R inv = rsqrt_fast<R>(dj);
        const R l0 = (r0 >= j) ? a0[j] * inv : (R)0;
        const R l1 = (r1 >= j) ? a1[j] * inv : (R)0;
        a0[j] = l0;
        a1[j] = l1;
        LT[j * LS + r0] = l0;
        LT[j * LS + r1] = l1;
        R v0 = 0, v1 = 0;
        if (SOLVE) {
            invd[j] = inv;
            v0 = c0[j] * inv;
            v1 = c1[j] * inv;
            c0[j] = v0;
            c1[j] = v1;
        }
        __syncwarp();
        const R nl0 = -l0, nl1 = -l1, nv0 = -v0, nv1 = -v1;
#pragma unroll
        for (int cv = (j + 1) / VEC; cv < NV; ++cv) {
            const VecT lv = *reinterpret_cast<const VecT*>(LT + j * LS + cv * VEC);
            const R* le = reinterpret_cast<const R*>(&lv);
#pragma unroll
            for (int q = 0; q < VEC; q += 2) {               // columns c, c + 1 (c even) as one packed FMA
                const int c = cv * VEC + q;
                if (c > j && c + 1 < n) {
                    fma2<R>(a0[c], a0[c + 1], nl0, nl0, le[q], le[q + 1]);
                    fma2<R>(a1[c], a1[c + 1], nl1, nl1, le[q], le[q + 1]);
                    if (SOLVE) {
                        fma2<R>(c0[c], c0[c + 1], le[q], le[q + 1], nv0, nv0);
                        fma2<R>(c1[c], c1[c + 1], le[q], le[q + 1], nv1, nv1);
                    }
                } else {
#pragma unroll
                    for (int t = 0; t < 2; ++t) {
                        const int cc = c + t;
                        if (cc > j && cc < n) {
                            a0[cc] = fma(nl0, le[q + t], a0[cc]);
                            a1[cc] = fma(nl1, le[q + t], a1[cc]);
                            if (SOLVE) {
                                c0[cc] = fma(le[q + t], nv0, c0[cc]);
                                c1[cc] = fma(le[q + t], nv1, c1[cc]);
                            }
                        }
                    }
                }
            }
        }
    }
}

template <typename R, int D_, int L_, int WARPS>
__global__ void __launch_bounds__(32 * WARPS, 1)
kalman_backprep_rows2_kernel(const R* __restrict__ stash_m, const R* __restrict__ stash_S,
                             const int* __restrict__ mask, const int* __restrict__ z, const R* __restrict__ Ab,
                             const R* __restrict__ Q, R jitter, const R* __restrict__ w_tape, SeedArg seed,
                             int N, int T, R* __restrict__ GH) {
    typedef PrepRows2<R, D_, L_> SM;
    typedef typename Vec16<R>::type VecT;
    constexpr int n = SM::n, NH = SM::NH, GW = SM::GW, FPW = SM::FPW, LS = SM::LS, NP = SM::NP;
    constexpr int NO = n - D_, NA1 = n + 1;
    constexpr int VEC = 16 / (int)sizeof(R), NV = (n + VEC - 1) / VEC;
    constexpr int SMS = stash_m_stride(n), SSS = stash_S_stride(n);
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, grp = lane / GW, gl = lane % GW;
    const int gbase = lane - gl;
    R* T1 = reinterpret_cast<R*>(smem_raw) + (size_t)(warp * FPW + grp) * SM::per_group;   // Lp' rows (n x LS)
    R* T2 = T1 + n * LS;                     // Wt' / V' / Ls' rows (n x LS)
    R* As = T2 + n * LS;                     // D_ rows [A | b]
    R* Sb = As + D_ * LS;                    // packed lower triangle of S
    R* mvb = Sb + SM::SB;                    // 2 x NP filtered means (ping-pong)
    R* wvb = mvb + 2 * NP;                   // 2 x NP normals (ping-pong)
    R* mp = wvb + 2 * NP;
    R* invd = mp + NP;
    const int Tx = T - L_ + 1;
    const long long frames = (long long)N * Tx;
    const long long stride = (long long)gridDim.x * WARPS * FPW;
    const bool act = gl < NH;
    const int r0 = act ? gl : NH - 1;                    // idle lanes shadow a valid row (same values, same addresses)
    const bool has1 = act && (gl + NH < n);
    const int r1 = has1 ? gl + NH : n - 1;
    // S is the lower triangle, packed by columns by the row-per-lane filter (n <= 32):
    //   S[r][c] = c <= r ? Sb[col_start(c) + r - c] : Sb[col_start(r) + c - r],
    // and by rows by the generic filter (n > 32): S[r][c] = c <= r ? Sb[r(r+1)/2 + c] : Sb[c(c+1)/2 + r].
    constexpr bool ROWPACK = n > 32;
    auto Sat = [&](int r, int c) -> R {
        if (ROWPACK) return (c <= r) ? Sb[r * (r + 1) / 2 + c] : Sb[c * (c + 1) / 2 + r];
        return (c <= r) ? Sb[col_start(n, c) + r - c] : Sb[col_start(n, r) + c - r];
    };
    const R eps = (R)KPMS_EPS_SHIFT + jitter;

    // frame status: -1 = none, 0 = masked (identity record), 1 = regular, 2 = last frame of its chain
    auto status = [&](long long g) {
        if (g >= frames) return -1;
        const int nn = (int)(g / Tx), i = (int)(g % Tx);
        if (i == Tx - 1) return 2;
        return mask[(size_t)nn * T + (L_ - 1) + i] != 0 ? 1 : 0;
    };
    auto issue_A = [&](long long g) {
        if (status(g) == 1) {
            const int nn = (int)(g / Tx), i = (int)(g % Tx);
            const R* A = Ab + (size_t)z[(size_t)nn * (Tx - 1) + i] * D_ * NA1;
            for (int w = gl; w < D_ * NA1; w += GW) cp_async_elem(As + (w / NA1) * LS + (w % NA1), A + w);
        }
        asm volatile("cp.async.commit_group;\n" ::);
    };
    auto issue_S = [&](long long g, int buf) {
        if (status(g) > 0) {
            const char* Sg = reinterpret_cast<const char*>(stash_S + (size_t)g * SSS);
            for (int c = gl; c < SSS * (int)sizeof(R) / 16; c += GW) cp_async_16(reinterpret_cast<char*>(Sb) + 16 * c, Sg + 16 * c);
            const char* mg = reinterpret_cast<const char*>(stash_m + (size_t)g * SMS);
            for (int c = gl; c < SMS * (int)sizeof(R) / 16; c += GW) cp_async_16(reinterpret_cast<char*>(mvb + NP * buf) + 16 * c, mg + 16 * c);
            if (w_tape)
                for (int c = gl; c < n; c += GW) cp_async_elem(wvb + NP * buf + c, w_tape + (size_t)g * n + c);
        }
        asm volatile("cp.async.commit_group;\n" ::);
    };
    auto publish_rows = [&](R* dst, const R (&u0)[n], const R (&u1)[n]) {
#pragma unroll
        for (int cv = 0; cv < NV; ++cv) {
            VecT o0, o1;
            R* e0 = reinterpret_cast<R*>(&o0);
            R* e1 = reinterpret_cast<R*>(&o1);
#pragma unroll
            for (int q = 0; q < VEC; ++q) {
                e0[q] = (cv * VEC + q < n) ? u0[cv * VEC + q] : (R)0;
                e1[q] = (cv * VEC + q < n) ? u1[cv * VEC + q] : (R)0;
            }
            *reinterpret_cast<VecT*>(dst + r0 * LS + cv * VEC) = o0;
            *reinterpret_cast<VecT*>(dst + r1 * LS + cv * VEC) = o1;
        }
    };

    int buf = 0;
    {
        const long long g0 = ((long long)blockIdx.x * WARPS + warp) * FPW + grp;
        issue_A(g0);
        issue_S(g0, buf);
    }
    // All warps walk their frames in lockstep phases (shared instruction fetch); every lane executes every
    // phase - frames that are absent, masked or terminal only skip the memory traffic.
    for (long long base = (long long)blockIdx.x * WARPS * FPW; base < frames; base += stride, buf ^= 1) {
        const long long g = base + (long long)warp * FPW + grp, gn = g + stride;
        const int stat = status(g);
        const bool on = (stat == 1), term = (stat == 2);
        const int nn = on || term || stat == 0 ? (int)(g / Tx) : 0, i = stat >= 0 ? (int)(g % Tx) : 0;
        R* Gout = GH + (size_t)(stat >= 0 ? g : 0) * PrepSmem<R, D_, L_>::RECS;
        R* hout = Gout + n * n;
        asm volatile("cp.async.wait_all;\n" ::);
        __syncwarp();
        const R* mv = mvb + NP * buf;
        R* wv = wvb + NP * buf;
        if (stat > 0 && !w_tape) {
            if (act) {
                Philox gen(seed, KPMS_STREAM_X, (uint64_t)g * n + r0);
                double a0, a1;
                philox_normal2(gen, a0, a1);
                wv[r0] = (R)a0;
            }
            if (has1) {
                Philox gen(seed, KPMS_STREAM_X, (uint64_t)g * n + r1);
                double a0, a1;
                philox_normal2(gen, a0, a1);
                wv[r1] = (R)a0;
            }
        }
        if (stat == 0) {
            for (int w = gl; w < n * n; w += GW) Gout[w] = ((w / n) == (w % n)) ? (R)1 : (R)0;
            for (int w = gl; w < n; w += GW) hout[w] = (R)0;
        }
        const int zi = on ? z[(size_t)nn * (Tx - 1) + i] : 0;
        R wt0[n], wt1[n], pp0[n], pp1[n];
        // ---- phase 1: Wt columns r0, r1 = Aaug S[:, r]; mp = Aaug m + b
        {
            R s0[n], s1[n];
#pragma unroll
            for (int c = 0; c < n; ++c) {
                s0[c] = Sat(r0, c);
                s1[c] = Sat(r1, c);
            }
#pragma unroll
            for (int r = 0; r < NO; ++r) { wt0[r] = s0[r + D_]; wt1[r] = s1[r + D_]; }
            // the D_ dense rows of Aaug: a rolled loop (code size), results parked in this lane's own rows
            // of T2 and read back with static indices
#pragma unroll 1
            for (int a = 0; a < D_; ++a) {
                R x0 = 0, x1 = 0, y0 = 0, y1 = 0;
#pragma unroll
                for (int cv = 0; cv < NV; ++cv) {
                    const VecT av = *reinterpret_cast<const VecT*>(As + a * LS + cv * VEC);
                    const R* ae = reinterpret_cast<const R*>(&av);
#pragma unroll
                    for (int q = 0; q < VEC; q += 2) {
                        const int e = cv * VEC + q;
                        if (e + 1 < n) {
                            fma2<R>(x0, x1, ae[q], ae[q + 1], s0[e], s0[e + 1]);
                            fma2<R>(y0, y1, ae[q], ae[q + 1], s1[e], s1[e + 1]);
                        } else if (e < n) {
                            x0 = fma(ae[q], s0[e], x0);
                            y0 = fma(ae[q], s1[e], y0);
                        }
                    }
                }
                T2[r0 * LS + NO + a] = x0 + x1;
                T2[r1 * LS + NO + a] = y0 + y1;
            }
            __syncwarp();
#pragma unroll
            for (int a = 0; a < D_; ++a) { wt0[NO + a] = T2[r0 * LS + NO + a]; wt1[NO + a] = T2[r1 * LS + NO + a]; }
            R m0, m1;
            if (r0 < NO) m0 = mv[r0 + D_];
            else {
                const R* arow = As + (r0 - NO) * LS;
                m0 = arow[n];
#pragma unroll 6
                for (int e = 0; e < n; ++e) m0 = fma(arow[e], mv[e], m0);
            }
            if (r1 < NO) m1 = mv[r1 + D_];
            else {
                const R* arow = As + (r1 - NO) * LS;
                m1 = arow[n];
#pragma unroll 6
                for (int e = 0; e < n; ++e) m1 = fma(arow[e], mv[e], m1);
            }
            mp[r0] = m0;
            mp[r1] = m1;
            publish_rows(T2, wt0, wt1);               // T2 = Wt'
        }
        __syncthreads();
        // ---- phase 2: Pp rows r0, r1 = Wt[r, :] Aaug' + Qaug
        {
            R w0[n], w1[n];
#pragma unroll
            for (int e = 0; e < n; ++e) { w0[e] = T2[e * LS + r0]; w1[e] = T2[e * LS + r1]; }
#pragma unroll
            for (int c = 0; c < NO; ++c) {
                pp0[c] = w0[c + D_] + ((r0 == c) ? eps : (R)0);
                pp1[c] = w1[c + D_] + ((r1 == c) ? eps : (R)0);
            }
            const R* Qk = Q + (size_t)zi * D_ * D_;
#pragma unroll 1
            for (int a = 0; a < D_; ++a) {
                R x0 = (on && r0 >= NO) ? (__ldg(Qk + (r0 - NO) * D_ + a) + ((r0 - NO == a) ? jitter : (R)0)) : (R)0;
                R y0 = (on && r1 >= NO) ? (__ldg(Qk + (r1 - NO) * D_ + a) + ((r1 - NO == a) ? jitter : (R)0)) : (R)0;
                R x1 = 0, y1 = 0;
#pragma unroll
                for (int cv = 0; cv < NV; ++cv) {
                    const VecT av = *reinterpret_cast<const VecT*>(As + a * LS + cv * VEC);
                    const R* ae = reinterpret_cast<const R*>(&av);
#pragma unroll
                    for (int q = 0; q < VEC; q += 2) {
                        const int e = cv * VEC + q;
                        if (e + 1 < n) {
                            fma2<R>(x0, x1, ae[q], ae[q + 1], w0[e], w0[e + 1]);
                            fma2<R>(y0, y1, ae[q], ae[q + 1], w1[e], w1[e + 1]);
                        } else if (e < n) {
                            x0 = fma(ae[q], w0[e], x0);
                            y0 = fma(ae[q], w1[e], y0);
                        }
                    }
                }
                T1[r0 * LS + a] = x0 + x1;                 // T1 is free until phase 3
                T1[r1 * LS + a] = y0 + y1;
            }
            __syncwarp();
#pragma unroll
            for (int a = 0; a < D_; ++a) { pp0[NO + a] = T1[r0 * LS + a]; pp1[NO + a] = T1[r1 * LS + a]; }
            __syncwarp();                            // As and T2 (Wt') fully consumed
            issue_A(gn);
        }
        __syncthreads();
        // ---- phase 3: Lp = chol(Pp), V = Lp^-1 Wt
        chol_rows2<R, n, NH, LS, true>(pp0, pp1, T1, wt0, wt1, invd, r0, r1, gbase);   // T1 = Lp', wt = V[:, r]
        publish_rows(T2, wt0, wt1);                  // T2 = V'
        __syncthreads();
        // ---- phase 4: Sigma rows = S[r, :] - V[:, r]' V   (terminal frame: Sigma = S)
        // rolled over a (code size): row a of V' is consumed by iteration a only, so Sigma[a][:] takes its
        // place in T2; afterwards every lane reads its own two rows back with static indices
#pragma unroll 1
        for (int a = 0; a < n; ++a) {
            R x0 = 0, x1 = 0, y0 = 0, y1 = 0;
            if (!term) {
#pragma unroll
                for (int cv = 0; cv < NV; ++cv) {
                    const VecT vv = *reinterpret_cast<const VecT*>(T2 + a * LS + cv * VEC);
                    const R* ve = reinterpret_cast<const R*>(&vv);
#pragma unroll
                    for (int q = 0; q < VEC; q += 2) {
                        const int e = cv * VEC + q;
                        if (e + 1 < n) {
                            fma2<R>(x0, x1, ve[q], ve[q + 1], wt0[e], wt0[e + 1]);
                            fma2<R>(y0, y1, ve[q], ve[q + 1], wt1[e], wt1[e + 1]);
                        } else if (e < n) {
                            x0 = fma(ve[q], wt0[e], x0);
                            y0 = fma(ve[q], wt1[e], y0);
                        }
                    }
                }
            }
            const R sx = Sat(r0, a) - (x0 + x1), sy = Sat(r1, a) - (y0 + y1);
            __syncwarp();
            T2[a * LS + r0] = sx;
            T2[a * LS + r1] = sy;
        }
        __syncwarp();
#pragma unroll
        for (int cv = 0; cv < NV; ++cv) {
            const VecT u0 = *reinterpret_cast<const VecT*>(T2 + r0 * LS + cv * VEC);
            const VecT u1 = *reinterpret_cast<const VecT*>(T2 + r1 * LS + cv * VEC);
            const R* e0 = reinterpret_cast<const R*>(&u0);
            const R* e1 = reinterpret_cast<const R*>(&u1);
#pragma unroll
            for (int q = 0; q < VEC; ++q)
                if (cv * VEC + q < n) { pp0[cv * VEC + q] = e0[q]; pp1[cv * VEC + q] = e1[q]; }   // pp now holds the Sigma rows
        }
        __syncwarp();                                // Sb and T2 (V') fully consumed
        issue_S(gn, buf ^ 1);
        __syncthreads();
        // ---- phase 5: Ls = chol(Sigma)
        {
            R d0[n], d1[n];                          // unused by the non-solving instantiation
            chol_rows2<R, n, NH, LS, false>(pp0, pp1, T2, d0, d1, invd, r0, r1, gbase);   // pp[c <= r] = Ls[r][c]
        }
        __syncthreads();
        // ---- phase 6: X = Lp^-T V (columns r0, r1, in place over wt), records out
        {
#pragma unroll
            for (int r = n - 1; r >= 0; --r) {
                R x0 = 0, x1 = 0, y0 = 0, y1 = 0;
#pragma unroll
                for (int cv = (r + 1) / VEC; cv < NV; ++cv) {
                    const VecT lv = *reinterpret_cast<const VecT*>(T1 + r * LS + cv * VEC);
                    const R* le = reinterpret_cast<const R*>(&lv);
#pragma unroll
                    for (int q = 0; q < VEC; q += 2) {
                        const int e = cv * VEC + q;
                        if (e > r && e + 1 < n) {
                            fma2<R>(x0, x1, le[q], le[q + 1], wt0[e], wt0[e + 1]);
                            fma2<R>(y0, y1, le[q], le[q + 1], wt1[e], wt1[e + 1]);
                        } else {
#pragma unroll
                            for (int t = 0; t < 2; ++t)
                                if (e + t > r && e + t < n) {
                                    x0 = fma(le[q + t], wt0[e + t], x0);
                                    y0 = fma(le[q + t], wt1[e + t], y0);
                                }
                        }
                    }
                }
                const R iv = invd[r];
                wt0[r] = (wt0[r] - (x0 + x1)) * iv;
                wt1[r] = (wt1[r] - (y0 + y1)) * iv;
            }
            if (on) {
#pragma unroll
                for (int r = 0; r < n; ++r) {
                    if (act) Gout[r * n + r0] = wt0[r];
                    if (has1) Gout[r * n + r1] = wt1[r];
                }
            }
            // h = m - X' mp + Ls w   (terminal frame: h = m + Ls w)
            R h0 = mv[r0], h1 = mv[r1], g0 = 0, g1 = 0, k0 = 0, k1 = 0;
#pragma unroll
            for (int c = 0; c < n; ++c) {
                const R mpc = mp[c], wc = wv[c];
                g0 = fma(-wt0[c], mpc, g0);
                g1 = fma(-wt1[c], mpc, g1);
                k0 = fma((c <= r0) ? pp0[c] : (R)0, wc, k0);
                k1 = fma((c <= r1) ? pp1[c] : (R)0, wc, k1);
            }
            if (on || term) {
                if (act) hout[r0] = h0 + (on ? g0 : (R)0) + k0;
                if (has1) hout[r1] = h1 + (on ? g1 : (R)0) + k1;
            }
        }
        __syncthreads();                             // also orders this frame's shared-memory reads before the next one's writes
    }
    asm volatile("cp.async.wait_all;\n" ::);
}
