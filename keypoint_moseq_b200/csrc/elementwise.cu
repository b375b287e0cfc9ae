// Per-frame / per-keypoint resamplers of the keypoint-SLDS sweep:
//   resample_scales, resample_heading, resample_location, resample_obs_variance statistics
// (jax_moseq.models.keypoint_slds.gibbs, reached from keypoint_moseq/fitting.py:25).
//
// Ct (k*Dk, d+1) is the lifted observation operator (Gamma kron I) Cd, so that the centred,
// aligned pose is Ybar[j][c] = Ct[j*Dk+c][:d] . x + Ct[j*Dk+c][d].
#include "common.cuh"
#include "../../include/kpms_b200.h"

namespace kpms {

// ---------------------------------------------------------------------------
// K4: s[n,t,j] ~ ScaledInvChi2; one thread per (frame, keypoint), fully coalesced on Y/s/prior
// ---------------------------------------------------------------------------
template <typename R, int DK>
__global__ void __launch_bounds__(256)
scales_kernel(const R* __restrict__ Y, const R* __restrict__ x, const R* __restrict__ v,
              const R* __restrict__ h, const R* __restrict__ Ct, const R* __restrict__ sigmasq,
              const R* __restrict__ prior, double nu_s, GammaShape shape, const R* __restrict__ g_tape, SeedArg seed,
              long long frames, int k, int d, R* __restrict__ s_out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    R* Cs = reinterpret_cast<R*>(smem_raw);
    double* isg = reinterpret_cast<double*>(Cs + align_up((size_t)k * DK * (d + 1), 2));     // 1 / sigmasq per keypoint
    for (int i = threadIdx.x; i < k * DK * (d + 1); i += blockDim.x) Cs[i] = Ct[i];
    for (int i = threadIdx.x; i < k; i += blockDim.x) isg[i] = 1.0 / (double)sigmasq[i];
    __syncthreads();
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= frames * k) return;
    const long long ft = e / k;
    const int j = (int)(e - ft * k);
    R sn, cs;
    sincos_r<R>(h[ft], sn, cs);
    R yb[DK];
#pragma unroll
    for (int c = 0; c < DK; ++c) {
        const R* crow = Cs + (size_t)(j * DK + c) * (d + 1);
        R acc = crow[d];
        for (int a = 0; a < d; ++a) acc = fma(crow[a], x[ft * d + a], acc);
        yb[c] = acc;
    }
    R r0 = cs * yb[0] - sn * yb[1], r1 = sn * yb[0] + cs * yb[1];
    yb[0] = r0;
    yb[1] = r1;
    R sq = 0;
#pragma unroll
    for (int c = 0; c < DK; ++c) {
        R df = Y[e * DK + c] - yb[c] - v[ft * DK + c];
        sq = fma(df, df, sq);
    }
    // (the Marsaglia-Tsang constants of the shared shape arrive as a kernel argument and 1 / sigmasq sits in shared
    // memory: one double division per element, the final one, instead of three and a square root)
    const double variance = fma((double)sq, isg[j], nu_s * (double)prior[e]);
    Philox gen(seed, KPMS_STREAM_S, (uint64_t)e);
    const double gam = gamma_draw<R>(shape, g_tape ? g_tape + e * KPMS_GAMMA_TAPE : nullptr, gen);
    s_out[e] = (R)(variance / (2.0 * gam));
}

// ---------------------------------------------------------------------------
// K9: sum_t mask * sqerr / s per keypoint (+ number of valid frames); deterministic two-stage sum
// ---------------------------------------------------------------------------
template <typename R, int DK>
__global__ void __launch_bounds__(128)
obsvar_partial_kernel(const R* __restrict__ Y, const int* __restrict__ mask, const R* __restrict__ x,
                      const R* __restrict__ v, const R* __restrict__ h, const R* __restrict__ s,
                      const R* __restrict__ Ct, long long frames, int k, int d, double* __restrict__ partial) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    R* Cs = reinterpret_cast<R*>(smem_raw);
    __shared__ double red[32];
    for (int i = threadIdx.x; i < k * DK * (d + 1); i += blockDim.x) Cs[i] = Ct[i];
    __syncthreads();
    const long long ft = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool on = ft < frames && mask[ft] != 0;
    R sn = 0, cs = 1;
    if (on) sincos_r<R>(h[ft], sn, cs);
    for (int j = 0; j < k; ++j) {
        double val = 0.0;
        if (on) {
            R yb[DK];
#pragma unroll
            for (int c = 0; c < DK; ++c) {
                const R* crow = Cs + (size_t)(j * DK + c) * (d + 1);
                R acc = crow[d];
                for (int a = 0; a < d; ++a) acc = fma(crow[a], x[ft * d + a], acc);
                yb[c] = acc;
            }
            R r0 = cs * yb[0] - sn * yb[1], r1 = sn * yb[0] + cs * yb[1];
            yb[0] = r0;
            yb[1] = r1;
            R sq = 0;
#pragma unroll
            for (int c = 0; c < DK; ++c) {
                R df = Y[(ft * k + j) * DK + c] - yb[c] - v[ft * DK + c];
                sq = fma(df, df, sq);
            }
            val = (double)sq / (double)s[ft * k + j];
        }
        double tot = block_sum(val, red);
        if (threadIdx.x == 0) partial[(size_t)blockIdx.x * (k + 1) + j] = tot;
    }
    double cnt = block_sum(on ? 1.0 : 0.0, red);
    if (threadIdx.x == 0) partial[(size_t)blockIdx.x * (k + 1) + k] = cnt;
}

__global__ void column_sum_kernel(const double* __restrict__ partial, int rows, int cols, double* __restrict__ out) {
    __shared__ double red[32];
    const int c = blockIdx.x;
    double acc = 0.0;
    for (int r = threadIdx.x; r < rows; r += blockDim.x) acc += partial[(size_t)r * cols + c];
    acc = block_sum(acc, red);
    if (threadIdx.x == 0) out[c] = acc;
}

// ---------------------------------------------------------------------------
// K5 + K6a: heading draw and the centroid pseudo-observation (mu_t, gamma_t^2) in one pass
// ---------------------------------------------------------------------------
template <typename R, int DK>
__global__ void __launch_bounds__(128)
heading_location_kernel(const R* __restrict__ Y, const R* __restrict__ x, const R* __restrict__ v,
                        const R* __restrict__ h_in, const R* __restrict__ s, const R* __restrict__ Ct,
                        const R* __restrict__ sigmasq, int fix_heading, const R* __restrict__ u_tape,
                        SeedArg seed, long long frames, int k, int d, R* __restrict__ h_out,
                        R* __restrict__ mu, R* __restrict__ gsq, R* __restrict__ wbuf) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    R* Cs = reinterpret_cast<R*>(smem_raw);
    R* sg = Cs + (size_t)k * DK * (d + 1);
    for (int i = threadIdx.x; i < k * DK * (d + 1); i += blockDim.x) Cs[i] = Ct[i];
    for (int i = threadIdx.x; i < k; i += blockDim.x) sg[i] = sigmasq[i];
    __syncthreads();
    const long long ft = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (ft >= frames) return;
    R vv[DK], sumY[DK], sumB[DK];
#pragma unroll
    for (int c = 0; c < DK; ++c) { vv[c] = v[ft * DK + c]; sumY[c] = 0; sumB[c] = 0; }
    R sumw = 0, kc = 0, ks = 0;
    for (int j = 0; j < k; ++j) {
        const R w = (R)1 / (s[ft * k + j] * sg[j]);
        R yb[DK], yy[DK];
#pragma unroll
        for (int c = 0; c < DK; ++c) {
            const R* crow = Cs + (size_t)(j * DK + c) * (d + 1);
            R acc = crow[d];
            for (int a = 0; a < d; ++a) acc = fma(crow[a], x[ft * d + a], acc);
            yb[c] = acc;
            yy[c] = Y[(ft * k + j) * DK + c];
            sumY[c] = fma(w, yy[c], sumY[c]);
            sumB[c] = fma(w, yb[c], sumB[c]);
        }
        const R y0 = yy[0] - vv[0], y1 = yy[1] - vv[1];
        kc = fma(w, yb[0] * y0 + yb[1] * y1, kc);
        ks = fma(w, yb[0] * y1 - yb[1] * y0, ks);
        sumw += w;
    }
    double hn;
    if (fix_heading) hn = (double)h_in[ft];
    else {
        Philox gen(seed, KPMS_STREAM_H, (uint64_t)ft);
        const double kcd = (double)kc, ksd = (double)ks;
        hn = vonmises_draw<R>(atan2(ksd, kcd), sqrt(kcd * kcd + ksd * ksd),
                              u_tape ? u_tape + ft * (KPMS_VM_R * 3) : nullptr, gen);
    }
    const R hr = (R)hn;
    h_out[ft] = hr;
    R sn, cs;
    sincos_r<R>(hr, sn, cs);
    const R g = (R)1 / sumw;
    R rb[DK];
#pragma unroll
    for (int c = 0; c < DK; ++c) rb[c] = sumB[c];
    rb[0] = cs * sumB[0] - sn * sumB[1];
    rb[1] = sn * sumB[0] + cs * sumB[1];
#pragma unroll
    for (int c = 0; c < DK; ++c) mu[ft * DK + c] = (sumY[c] - rb[c]) * g;
    gsq[ft] = g;
    if (wbuf) {          // standard normals for the centroid FFBS, generated here so the serial scan only loads
        Philox gen(seed, KPMS_STREAM_V, (uint64_t)ft);
        double a0, a1;
        philox_normal2(gen, a0, a1);
        wbuf[ft * DK + 0] = (R)a0;
        wbuf[ft * DK + 1] = (R)a1;
        if (DK > 2) { philox_normal2(gen, a0, a1); wbuf[ft * DK + 2] = (R)a0; }
    }
}

// ---------------------------------------------------------------------------
// K6b: random-walk FFBS of the centroid, parallel in time.  Covariances are scalar multiples of
// I, so the three recursions of the sampler are scans over tiny associative elements:
//   predicted variance   Pp' = ((g + sl) Pp + sl g) / (Pp + g)          (Moebius map, 2x2 matrix)
//   filtered mean        m'  = (1 - kg) m + kg mu                        (affine map)
//   backward sample      v_t = gain v_{t+1} + (1 - gain) m_t + sd w_t    (affine map, reversed)
// One CTA per chain; every thread owns a contiguous run of frames, composes its run serially,
// the CTA scans the per-thread composites (warp shuffles + one shared-memory hop), and each thread
// then replays its run from the exact incoming value with the same arithmetic as the sequential
// recursion.  All Moebius entries are positive, so the composition has no cancellation.
// v_out doubles as the filtered-mean stash; fP is (N,T) scratch.
// ---------------------------------------------------------------------------
template <typename R, int NV>
struct Affine {      // y = a x + b,  b has NV components
    R a, b[NV];
};
template <typename R, int NV>
__device__ __forceinline__ Affine<R, NV> affine_after(const Affine<R, NV>& second, const Affine<R, NV>& first) {
    Affine<R, NV> o;
    o.a = second.a * first.a;
#pragma unroll
    for (int c = 0; c < NV; ++c) o.b[c] = fma(second.a, first.b[c], second.b[c]);
    return o;
}
template <typename R>
struct Moebius {     // P' = (a P + b) / (c P + d), entries > 0, scaled so the largest is 1
    R a, b, c, d;
};
template <typename R>
__device__ __forceinline__ Moebius<R> moebius_after(const Moebius<R>& s, const Moebius<R>& f) {
    Moebius<R> o;
    o.a = fma(s.a, f.a, s.b * f.c);
    o.b = fma(s.a, f.b, s.b * f.d);
    o.c = fma(s.c, f.a, s.d * f.c);
    o.d = fma(s.c, f.b, s.d * f.d);
    const R m = (R)1 / fmax(fmax(o.a, o.b), fmax(o.c, o.d));
    o.a *= m; o.b *= m; o.c *= m; o.d *= m;
    return o;
}

// Inclusive scan over the CTA of per-thread elements (thread order = application order), returned
// as the EXCLUSIVE prefix (composition of all earlier threads' elements; identity for thread 0).
// E is a struct of NW 32-bit or 64-bit words of type R.
template <typename R, typename E, typename F>
__device__ inline E block_exclusive_scan(E mine, const E& identity, F after, E* sh /* >= 32 */) {
    constexpr int NW = sizeof(E) / sizeof(R);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    auto shfl_up = [&](const E& e, int o) {
        E r;
        const R* src = reinterpret_cast<const R*>(&e);
        R* dst = reinterpret_cast<R*>(&r);
#pragma unroll
        for (int q = 0; q < NW; ++q) dst[q] = __shfl_up_sync(0xffffffffu, src[q], o);
        return r;
    };
    E incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        E prev = shfl_up(incl, o);
        if (lane >= o) incl = after(incl, prev);
    }
    if (lane == 31) sh[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        E w = (lane < nwarps) ? sh[lane] : identity;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            E prev = shfl_up(w, o);
            if (lane >= o) w = after(w, prev);
        }
        sh[lane] = w;                               // inclusive over warps
    }
    __syncthreads();
    E excl = shfl_up(incl, 1);
    if (lane == 0) excl = identity;
    if (warp > 0) excl = after(excl, sh[warp - 1]);
    __syncthreads();
    return excl;
}

template <typename R, int DK, int NT>
__global__ void __launch_bounds__(NT)
location_ffbs_kernel(const R* __restrict__ mu, const R* __restrict__ gsq, const int* __restrict__ mask,
                     double sigmasq_loc, const R* __restrict__ w_tape, int N, int T,
                     R* __restrict__ fP, R* __restrict__ v_out) {
    typedef Affine<R, DK> Aff;
    typedef Moebius<R> Moe;
    __shared__ Moe sh_m[32];
    __shared__ Aff sh_a[32];
    const int nn = blockIdx.x, tid = threadIdx.x;
    const R* mun = mu + (size_t)nn * T * DK;
    const R* gn = gsq + (size_t)nn * T;
    const int* mk = mask + (size_t)nn * T;
    const R* wn = w_tape + (size_t)nn * T * DK;
    R* fPn = fP + (size_t)nn * T;
    R* vn = v_out + (size_t)nn * T * DK;
    const R sl = (R)sigmasq_loc;
    const int per = (T + NT - 1) / NT;
    const int lo = min(tid * per, T), hi = min(lo + per, T);
    Moe idm; idm.a = 1; idm.b = 0; idm.c = 0; idm.d = 1;
    Aff ida; ida.a = 1;
#pragma unroll
    for (int c = 0; c < DK; ++c) ida.b[c] = 0;

    // ---- pass 1: predicted variance at the start of every run
    Moe run = idm;
    for (int t = lo; t < hi; ++t) {
        if (mk[t]) {
            const R g = gn[t];
            Moe e; e.a = g + sl; e.b = sl * g; e.c = 1; e.d = g;
            run = moebius_after(e, run);
        }
    }
    const Moe pre = block_exclusive_scan<R>(run, idm, moebius_after<R>, sh_m);
    const R P0 = (R)KPMS_V_PRIOR_VAR;
    R Pp = fma(pre.a, P0, pre.b) / fma(pre.c, P0, pre.d);
    // replay the run: filtered variances (stash) and the run's affine map of the mean
    Aff runa = ida;
    for (int t = lo; t < hi; ++t) {
        if (mk[t]) {
            const R g = gn[t];
            const R Pc = Pp * g / (Pp + g);
            const R kg = Pc / g;
            Aff e; e.a = (R)1 - kg;
#pragma unroll
            for (int c = 0; c < DK; ++c) e.b[c] = kg * mun[(size_t)t * DK + c];
            runa = affine_after(e, runa);
            fPn[t] = Pc;
            Pp = Pc + sl;
        } else {
            fPn[t] = Pp;
        }
    }
    // ---- pass 2: filtered means (prior mean 0, so the incoming mean is the prefix offset)
    const Aff prea = block_exclusive_scan<R>(runa, ida, affine_after<R, DK>, sh_a);
    R mp[DK];
#pragma unroll
    for (int c = 0; c < DK; ++c) mp[c] = prea.b[c];
    for (int t = lo; t < hi; ++t) {
        if (mk[t]) {
            const R kg = fPn[t] / gn[t];
#pragma unroll
            for (int c = 0; c < DK; ++c) mp[c] = fma(kg, mun[(size_t)t * DK + c] - mp[c], mp[c]);
        }
#pragma unroll
        for (int c = 0; c < DK; ++c) vn[(size_t)t * DK + c] = mp[c];
    }
    __syncthreads();
    // ---- pass 3: backward sampling; thread order reversed so that the scan runs from T-1 down.
    // Frame T-1 draws from its filter marginal: a constant map v = m + sqrt(P) w.
    const int rt = NT - 1 - tid;
    const int lo2 = min(rt * per, T), hi2 = min(lo2 + per, T);
    auto back_elem = [&](int t) {
        Aff e;
        if (t == T - 1) {
            const R sd = sqrt(fPn[t]);
            e.a = 0;
#pragma unroll
            for (int c = 0; c < DK; ++c) e.b[c] = fma(sd, wn[(size_t)t * DK + c], vn[(size_t)t * DK + c]);
        } else if (mk[t]) {
            const R p = fPn[t];
            const R gain = p / (p + sl);
            const R sd = sqrt(gain * sl);
            e.a = gain;
#pragma unroll
            for (int c = 0; c < DK; ++c) {
                const R m = vn[(size_t)t * DK + c];
                e.b[c] = fma(-gain, m, m) + sd * wn[(size_t)t * DK + c];
            }
        } else {
            e = ida;
        }
        return e;
    };
    Aff runb = ida;
    for (int t = hi2 - 1; t >= lo2; --t) runb = affine_after(back_elem(t), runb);
    const Aff preb = block_exclusive_scan<R>(runb, ida, affine_after<R, DK>, sh_a);
    R vc[DK];
#pragma unroll
    for (int c = 0; c < DK; ++c) vc[c] = preb.b[c];           // v_{hi2}: the map so far applied to anything
    for (int t = hi2 - 1; t >= lo2; --t) {
        if (t == T - 1) {
            const R sd = sqrt(fPn[t]);
#pragma unroll
            for (int c = 0; c < DK; ++c) vc[c] = fma(sd, wn[(size_t)t * DK + c], vn[(size_t)t * DK + c]);
        } else if (mk[t]) {
            const R p = fPn[t];
            const R gain = p / (p + sl);
            const R sd = sqrt(gain * sl);
#pragma unroll
            for (int c = 0; c < DK; ++c) {
                const R m = vn[(size_t)t * DK + c];
                vc[c] = fma(gain, vc[c] - m, m) + sd * wn[(size_t)t * DK + c];
            }
        }
#pragma unroll
        for (int c = 0; c < DK; ++c) vn[(size_t)t * DK + c] = vc[c];
    }
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
template <typename R>
static int scales_impl(const void* Y, const void* x, const void* v, const void* h, const void* Ct,
                       const void* sigmasq, const void* prior, double nu_s, const void* g_tape, SeedArg seed,
                       int N, int T, int k, int Dk, int d, void* s_out, cudaStream_t st) {
    const long long frames = (long long)N * T;
    const long long elems = frames * k;
    size_t smem = align_up((size_t)k * Dk * (d + 1), 2) * sizeof(R) + (size_t)k * sizeof(double);
    int blocks = (int)((elems + 255) / 256);
    KPMS_LAUNCH("resample_scales", st);
    if (Dk == 2)
        scales_kernel<R, 2><<<blocks, 256, smem, st>>>((const R*)Y, (const R*)x, (const R*)v, (const R*)h, (const R*)Ct,
                                                       (const R*)sigmasq, (const R*)prior, nu_s,
                                                       GammaShape(0.5 * (nu_s + KPMS_OBS_DOF(2))), (const R*)g_tape, seed,
                                                       frames, k, d, (R*)s_out);
    else if (Dk == 3)
        scales_kernel<R, 3><<<blocks, 256, smem, st>>>((const R*)Y, (const R*)x, (const R*)v, (const R*)h, (const R*)Ct,
                                                       (const R*)sigmasq, (const R*)prior, nu_s,
                                                       GammaShape(0.5 * (nu_s + KPMS_OBS_DOF(3))), (const R*)g_tape, seed,
                                                       frames, k, d, (R*)s_out);
    else return set_error(-3, "resample_scales: keypoint dimension must be 2 or 3, got %d", Dk);
    return check_launch("resample_scales");
}

template <typename R>
static int obsvar_impl(const void* Y, const int* mask, const void* x, const void* v, const void* h, const void* s,
                       const void* Ct, int N, int T, int k, int Dk, int d, double* out, void* ws, cudaStream_t st) {
    const long long frames = (long long)N * T;
    size_t smem = (size_t)k * Dk * (d + 1) * sizeof(R);
    int blocks = (int)((frames + 127) / 128);
    double* partial = reinterpret_cast<double*>(ws);
    {
    KPMS_LAUNCH("obsvar_partial", st);
    if (Dk == 2)
        obsvar_partial_kernel<R, 2><<<blocks, 128, smem, st>>>((const R*)Y, mask, (const R*)x, (const R*)v, (const R*)h,
                                                               (const R*)s, (const R*)Ct, frames, k, d, partial);
    else if (Dk == 3)
        obsvar_partial_kernel<R, 3><<<blocks, 128, smem, st>>>((const R*)Y, mask, (const R*)x, (const R*)v, (const R*)h,
                                                               (const R*)s, (const R*)Ct, frames, k, d, partial);
    else return set_error(-3, "obsvar_suffstats: keypoint dimension must be 2 or 3, got %d", Dk);
    }
    { KPMS_LAUNCH("obsvar_reduce", st); column_sum_kernel<<<k + 1, 256, 0, st>>>(partial, blocks, k + 1, out); }
    return check_launch("obsvar_suffstats");
}

template <typename R>
static int headloc_impl(const void* Y, const int* mask, const void* x, const void* v_in, const void* h_in,
                        const void* s, const void* Ct, const void* sigmasq, double sigmasq_loc, int fix_heading,
                        const void* u_tape, const void* w_tape, SeedArg seed, int N, int T, int k, int Dk, int d,
                        void* h_out, void* v_out, void* ws, cudaStream_t st) {
    const long long frames = (long long)N * T;
    size_t smem = ((size_t)k * Dk * (d + 1) + k) * sizeof(R);
    int blocks = (int)((frames + 127) / 128);
    char* base = reinterpret_cast<char*>(ws);
    R* mu = reinterpret_cast<R*>(base);
    R* gsq = reinterpret_cast<R*>(base + align_up((size_t)frames * Dk * sizeof(R), 256));
    R* fP = reinterpret_cast<R*>(base + align_up((size_t)frames * Dk * sizeof(R), 256) + align_up((size_t)frames * sizeof(R), 256));
    R* wbuf = w_tape ? nullptr : reinterpret_cast<R*>(base + align_up((size_t)frames * Dk * sizeof(R), 256) +
                                                      2 * align_up((size_t)frames * sizeof(R), 256));
    const R* wsrc = w_tape ? (const R*)w_tape : wbuf;
#define LAUNCH(DK)                                                                                             \
    { KPMS_LAUNCH("heading_location", st);                                                                    \
    heading_location_kernel<R, DK><<<blocks, 128, smem, st>>>(                                                 \
        (const R*)Y, (const R*)x, (const R*)v_in, (const R*)h_in, (const R*)s, (const R*)Ct, (const R*)sigmasq, \
        fix_heading, (const R*)u_tape, seed, frames, k, d, (R*)h_out, mu, gsq, wbuf); }                              \
    KPMS_LAUNCH("location_ffbs", st);                                                                          \
    location_ffbs_kernel<R, DK, 1024><<<N, 1024, 0, st>>>(mu, gsq, mask, sigmasq_loc, wsrc, N, T, fP, (R*)v_out)
    if (Dk == 2) { LAUNCH(2); }
    else if (Dk == 3) { LAUNCH(3); }
    else return set_error(-3, "resample_heading_location: keypoint dimension must be 2 or 3, got %d", Dk);
#undef LAUNCH
    return check_launch("resample_heading_location");
}

}  // namespace kpms

using namespace kpms;

extern "C" {

int kpms_resample_scales(int dtype, const void* Y, const void* x, const void* v, const void* h, const void* Ct,
                         const void* sigmasq, const void* noise_prior, double nu_s, const void* g_tape,
                         uint64_t seed, const uint64_t* seed_dev, int N, int T, int k, int Dk, int d, void* s_out,
                         void* stream) {
    return KPMS_DISPATCH_DTYPE(dtype, scales_impl, Y, x, v, h, Ct, sigmasq, noise_prior, nu_s, g_tape, SeedArg(seed, seed_dev), N, T,
                               k, Dk, d, s_out, (cudaStream_t)stream);
}

size_t kpms_obsvar_workspace_bytes(int N, int T, int k) {
    return (size_t)(((long long)N * T + 127) / 128) * (k + 1) * sizeof(double);
}

int kpms_obsvar_suffstats(int dtype, const void* Y, const int32_t* mask, const void* x, const void* v,
                          const void* h, const void* s, const void* Ct, int N, int T, int k, int Dk, int d,
                          double* out, void* ws, void* stream) {
    return KPMS_DISPATCH_DTYPE(dtype, obsvar_impl, Y, mask, x, v, h, s, Ct, N, T, k, Dk, d, out, ws,
                               (cudaStream_t)stream);
}

size_t kpms_heading_location_workspace_bytes(int dtype, int N, int T, int Dk) {
    size_t esz = dtype == 0 ? 4 : 8;
    size_t frames = (size_t)N * T;
    return 2 * kpms::align_up(frames * Dk * esz, 256) + 2 * kpms::align_up(frames * esz, 256);
}

int kpms_resample_heading_location(int dtype, const void* Y, const int32_t* mask, const void* x, const void* v_in,
                                   const void* h_in, const void* s, const void* Ct, const void* sigmasq,
                                   double sigmasq_loc, int fix_heading, const void* u_tape, const void* w_tape,
                                   uint64_t seed, const uint64_t* seed_dev, int N, int T, int k, int Dk, int d, void* h_out, void* v_out,
                                   void* ws, void* stream) {
    return KPMS_DISPATCH_DTYPE(dtype, headloc_impl, Y, mask, x, v_in, h_in, s, Ct, sigmasq, sigmasq_loc,
                               fix_heading, u_tape, w_tape, SeedArg(seed, seed_dev), N, T, k, Dk, d, h_out, v_out, ws,
                               (cudaStream_t)stream);
}

}  // extern "C"
