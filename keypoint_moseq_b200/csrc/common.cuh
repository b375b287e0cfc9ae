// Shared device helpers for the keypoint-SLDS Gibbs kernels (sm_100a).
//
// Randomness contract (DESIGN.md "draws"): every sampler either consumes an
// injected tape (verification mode, pointer != nullptr) or generates its draws
// from Philox4x32-10 keyed by the 64-bit sweep seed, with the counter built from
// (element index, stream id, draw index).  The transforms (Marsaglia-Tsang gamma,
// Best-Fisher von Mises, inverse-CDF categorical) are the same in both modes and
// are evaluated in double.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

#define KPMS_GAMMA_R 6
#define KPMS_VM_R 8
#define KPMS_GAMMA_TAPE (2 * KPMS_GAMMA_R + 1)
#define KPMS_EPS_SHIFT 1e-2
#define KPMS_X_PRIOR_VAR 10.0
#define KPMS_V_PRIOR_VAR 1e6
/* degrees of freedom one keypoint of dimension DK adds to the scaled-inverse-chi-square posteriors of the noise
 * scales and of sigmasq (oracle: SCALE_DOF / OBSVAR_DOF = None); upstream may hard-code 3 - change both together */
#define KPMS_OBS_DOF(DK) (DK)
#define KPMS_SM_COUNT 148     /* B200; the library is built for sm_100a only */
#define KPMS_MAX_CHUNKS 1024   /* time chunks per chain at most (a single 10^6-frame chain still fills the device) */

// stream ids for Philox counters (one per sampler)
enum KpmsStream : uint32_t {
    KPMS_STREAM_Z = 1, KPMS_STREAM_X = 2, KPMS_STREAM_S = 3, KPMS_STREAM_H = 4,
    KPMS_STREAM_V = 5, KPMS_STREAM_AR_G = 6, KPMS_STREAM_AR_B = 7, KPMS_STREAM_AR_CHI = 8,
    KPMS_STREAM_CRP = 9, KPMS_STREAM_BIN = 10, KPMS_STREAM_BETA = 11, KPMS_STREAM_PI = 12,
    KPMS_STREAM_SIGMA = 13
};

namespace kpms {

int set_error(int code, const char* fmt, ...);
int check_launch(const char* what);

// Every kernel launch is bracketed by a LaunchScope: it counts launches and, when profiling is
// enabled (kpms_profile_enable), records CUDA events on the launching stream around the kernel.
struct LaunchScope {
    int slot;
    cudaStream_t st;
    LaunchScope(const char* name, cudaStream_t st);
    ~LaunchScope();
};
#define KPMS_CAT2(a, b) a##b
#define KPMS_CAT(a, b) KPMS_CAT2(a, b)
#define KPMS_LAUNCH(name, st) kpms::LaunchScope KPMS_CAT(_launch_scope_, __LINE__)(name, st)

__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
__host__ __device__ inline size_t align_up(size_t a, size_t b) { return (a + b - 1) / b * b; }

// ---------------------------------------------------------------------------
// Philox4x32-10
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        uint32_t hi0 = __umulhi(M0, c.x), lo0 = M0 * c.x;
        uint32_t hi1 = __umulhi(M1, c.z), lo1 = M1 * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += 0x9E3779B9u;
        k.y += 0xBB67AE85u;
    }
    return c;
}

// Philox key of a sampler call: `salt` by value, optionally XORed with a 64-bit key read from device memory.
// A captured CUDA graph freezes by-value arguments, so a replayed sweep takes its per-sweep key from `dev`
// (advanced on the device by kpms_advance_seed) and keeps only the constant salt (e.g. the rank mix) by value.
struct SeedArg {
    uint64_t salt;
    const uint64_t* dev;
    __host__ __device__ SeedArg(uint64_t s = 0, const uint64_t* d = nullptr) : salt(s), dev(d) {}
    __device__ __forceinline__ operator uint64_t() const { return dev ? (__ldg(dev) ^ salt) : salt; }
};

// One logical generator per (seed, stream, element); `draw` advances inside it.
struct Philox {
    uint2 key;
    uint32_t stream;
    uint64_t elem;
    uint32_t draw;
    __device__ Philox(uint64_t seed, uint32_t stream_, uint64_t elem_)
        : key(make_uint2((uint32_t)seed, (uint32_t)(seed >> 32))), stream(stream_), elem(elem_), draw(0) {}
    __device__ __forceinline__ uint4 next4() {
        uint4 c = make_uint4((uint32_t)elem, (uint32_t)(elem >> 32), stream, draw++);
        return philox4x32_10(c, key);
    }
};

__device__ __forceinline__ double u32x2_to_unit(uint32_t a, uint32_t b) {
    // 53-bit uniform strictly inside (0,1)
    uint64_t m = (((uint64_t)a << 32) | b) >> 11;
    return ((double)m + 0.5) * (1.0 / 9007199254740992.0);
}

__device__ __forceinline__ void philox_uniform2(Philox& g, double& u0, double& u1) {
    uint4 r = g.next4();
    u0 = u32x2_to_unit(r.x, r.y);
    u1 = u32x2_to_unit(r.z, r.w);
}

__device__ __forceinline__ void philox_normal2(Philox& g, double& n0, double& n1) {
    double u0, u1;
    philox_uniform2(g, u0, u1);
    double r = sqrt(-2.0 * log(u0));
    double s, c;
    sincospi(2.0 * u1, &s, &c);
    n0 = r * c;
    n1 = r * s;
}

// ---------------------------------------------------------------------------
// Gamma(a,1), Marsaglia-Tsang.  tape: KPMS_GAMMA_TAPE values [normals R | uniforms R | boost]
// (verification mode, bounded attempts, fallback d) or nullptr (Philox, up to 64 attempts).
// The acceptance test is log u < x^2/2 + d(1 - v + log v); Marsaglia and Tsang's squeeze u < 1 - 0.0331 x^4 is a
// lower bound of the same acceptance function, so testing it first takes the same decisions and spares both
// logarithms on ~92 % of the attempts.  In Philox mode with float32 states the proposal normal comes from a
// float32 Box-Muller (one Philox call per attempt: two words for the normal, two for a 53-bit uniform); the test
// itself is evaluated in double in every mode.
// ---------------------------------------------------------------------------
__device__ __forceinline__ bool gamma_accept(double x, double u, double dd, double c, double& out) {
    const double vv = 1.0 + c * x;
    if (!(vv > 0)) return false;
    const double v3 = vv * vv * vv, x2 = x * x;
    if (u < 1.0 - 0.0331 * x2 * x2 || log(u) < 0.5 * x2 + dd - dd * v3 + dd * log(v3)) { out = dd * v3; return true; }
    return false;
}

// Marsaglia-Tsang constants of a shape parameter (a kernel whose draws all share one shape computes them once)
struct GammaShape {
    double a, dd, c;
    bool boost;
    __host__ __device__ explicit GammaShape(double a_) : a(a_), boost(a_ < 1.0) {
        dd = (boost ? a_ + 1.0 : a_) - 1.0 / 3.0;
        c = 1.0 / sqrt(9.0 * dd);
    }
};

template <typename R>
__device__ inline double gamma_draw(const GammaShape& sh, const R* tape, Philox& g) {
    const bool boost = sh.boost;
    const double a = sh.a, dd = sh.dd, c = sh.c;
    double out = dd;
    if (tape) {
#pragma unroll 1
        for (int r = 0; r < KPMS_GAMMA_R; ++r)
            if (gamma_accept((double)tape[r], (double)tape[KPMS_GAMMA_R + r], dd, c, out)) break;
        if (boost) out *= pow((double)tape[2 * KPMS_GAMMA_R], 1.0 / a);
    } else {
#pragma unroll 1
        for (int r = 0; r < 64; ++r) {
            double x, u;
            if (sizeof(R) == 4) {
                const uint4 w = g.next4();
                const float u0 = ((float)(w.x >> 8) + 0.5f) * (1.0f / 16777216.0f);
                const float u1 = ((float)(w.y >> 8) + 0.5f) * (1.0f / 16777216.0f);
                float sn, cs;
                sincospif(2.0f * u1, &sn, &cs);
                x = (double)(sqrtf(-2.0f * logf(u0)) * cs);
                u = u32x2_to_unit(w.z, w.w);
            } else {
                double x2, u2;
                philox_normal2(g, x, x2);
                philox_uniform2(g, u, u2);
            }
            if (gamma_accept(x, u, dd, c, out)) break;
        }
        if (boost) {
            double u, u2;
            philox_uniform2(g, u, u2);
            out *= pow(u, 1.0 / a);
        }
    }
    return out;
}

template <typename R>
__device__ inline double gamma_draw(double a, const R* tape, Philox& g) {
    return gamma_draw<R>(GammaShape(a), tape, g);
}

// ---------------------------------------------------------------------------
// von Mises(mu, kappa), Best-Fisher.  tape: KPMS_VM_R x 3 uniforms, or nullptr.
// ---------------------------------------------------------------------------
template <typename R>
__device__ inline double vonmises_draw(double mu, double kappa, const R* tape, Philox& g) {
    const double PI = 3.14159265358979323846;
    double dev = 0.0;
    if (kappa < 1e-8) {
        double u;
        if (tape) u = (double)tape[0];
        else { double u2; philox_uniform2(g, u, u2); }
        dev = PI * (2.0 * u - 1.0);
    } else {
        double tau = 1.0 + sqrt(1.0 + 4.0 * kappa * kappa);
        double rho = (tau - sqrt(2.0 * tau)) / (2.0 * kappa);
        double r = (1.0 + rho * rho) / (2.0 * rho);
        const int R_MAX = tape ? KPMS_VM_R : 64;
#pragma unroll 1
        for (int a = 0; a < R_MAX; ++a) {
            double u1, u2, u3;
            if (tape) { u1 = (double)tape[3 * a]; u2 = (double)tape[3 * a + 1]; u3 = (double)tape[3 * a + 2]; }
            else { double u4; philox_uniform2(g, u1, u2); philox_uniform2(g, u3, u4); }
            double zc = cos(PI * u1);
            double f = (1.0 + r * zc) / (r + zc);
            double c = kappa * (r - f);
            if ((c * (2.0 - c) - u2 > 0) || (log(c / u2) + 1.0 - c >= 0)) {
                double fc = fmin(1.0, fmax(-1.0, f));
                dev = (u3 > 0.5 ? 1.0 : -1.0) * acos(fc);
                break;
            }
        }
    }
    double out = mu + dev;
    return out - 2.0 * PI * floor((out + PI) / (2.0 * PI));
}

// ---------------------------------------------------------------------------
// warp helpers
// ---------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
template <typename T>
__device__ __forceinline__ T warp_max(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { T w = __shfl_xor_sync(0xffffffffu, v, o); v = w > v ? w : v; }
    return v;
}

// block-wide sum via shared scratch of >= 32 elements; result valid in all threads
template <typename T>
__device__ inline T block_sum(T v, T* scratch) {
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) scratch[w] = v;
    __syncthreads();
    T r = (lane < nw) ? scratch[lane] : T(0);
    r = warp_sum(r);
    return r;
}

// 16-byte vector of R (one LDS.128 / LDG.128)
template <typename R> struct Vec16;
template <> struct Vec16<float> { typedef float4 type; };
template <> struct Vec16<double> { typedef double2 type; };

template <typename R> __device__ __forceinline__ R rsqrt_r(R x);
template <> __device__ __forceinline__ float rsqrt_r<float>(float x) { return rsqrtf(x); }
template <> __device__ __forceinline__ double rsqrt_r<double>(double x) { return rsqrt(x); }

template <typename R> __device__ __forceinline__ R rsqrt_fast(R x);
template <> __device__ __forceinline__ float rsqrt_fast<float>(float x) {
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
template <> __device__ __forceinline__ double rsqrt_fast<double>(double x) { return rsqrt(x); }

// reciprocal to ~1 ulp without the IEEE division sequence
template <typename R> __device__ __forceinline__ R rcp_fast(R x);
template <> __device__ __forceinline__ float rcp_fast<float>(float x) {
    float y;                                   // MUFU.RCP + one Newton step: < 1 ulp, no slow-path branch
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return fmaf(y, fmaf(-x, y, 1.0f), y);
}
template <> __device__ __forceinline__ double rcp_fast<double>(double x) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    y = fma(y, fma(-x, y, 1.0), y);
    y = fma(y, fma(-x, y, 1.0), y);
    return y;
}

// Two independent FMAs in one issue slot: c0 += a0*b0, c1 += a1*b1.  float: fma.rn.f32x2 (SASS FFMA2,
// sm_100+; same rounding as two fma.rn, measured 68 TFLOP/s = the scalar FFMA peak at half the issue
// slots, tools/micro/dmma_peak.cu); the pack / unpack moves vanish when the operands sit in aligned
// register pairs.  double: two DFMAs.
template <typename R>
__device__ __forceinline__ void fma2(R& c0, R& c1, R a0, R a1, R b0, R b1) {
    c0 = fma(a0, b0, c0);
    c1 = fma(a1, b1, c1);
}
template <>
__device__ __forceinline__ void fma2<float>(float& c0, float& c1, float a0, float a1, float b0, float b1) {
    unsigned long long a, b, c;
    asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "f"(a0), "f"(a1));
    asm("mov.b64 %0, {%1, %2};" : "=l"(b) : "f"(b0), "f"(b1));
    asm("mov.b64 %0, {%1, %2};" : "=l"(c) : "f"(c0), "f"(c1));
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(c) : "l"(a), "l"(b));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(c0), "=f"(c1) : "l"(c));
}

// Two independent products in one issue slot: c0 = a0*b, c1 = a1*b (mul.rn.f32x2, sm_100+).
template <typename R>
__device__ __forceinline__ void mul2(R& c0, R& c1, R a0, R a1, R b) {
    c0 = a0 * b;
    c1 = a1 * b;
}
template <>
__device__ __forceinline__ void mul2<float>(float& c0, float& c1, float a0, float a1, float b) {
    unsigned long long a, bb, c;
    asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "f"(a0), "f"(a1));
    asm("mov.b64 %0, {%1, %1};" : "=l"(bb) : "f"(b));
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(c) : "l"(a), "l"(bb));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(c0), "=f"(c1) : "l"(c));
}

// asynchronous global -> shared copies (one element, or one 16-byte chunk)
template <typename R>
__device__ __forceinline__ void cp_async_elem(R* dst, const R* src) {
    const unsigned d32 = (unsigned)__cvta_generic_to_shared(dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], %2;\n" ::"r"(d32), "l"(src), "n"((int)sizeof(R)));
}
__device__ __forceinline__ void cp_async_16(void* dst, const void* src) {
    const unsigned d32 = (unsigned)__cvta_generic_to_shared(dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d32), "l"(src));
}

template <typename R> __device__ __forceinline__ void sincos_r(R x, R& s, R& c);
template <> __device__ __forceinline__ void sincos_r<float>(float x, float& s, float& c) { sincosf(x, &s, &c); }
template <> __device__ __forceinline__ void sincos_r<double>(double x, double& s, double& c) { sincos(x, &s, &c); }

// ---------------------------------------------------------------------------
// Time-parallel execution of the serial recursions ("speculative chunks with verified boundaries").
//
// A filter forgets its initial condition geometrically, so a chain is cut into chunks that run
// concurrently: chunk c > 0 starts W steps before its first frame from the chain's own prior
// (warm-up, nothing stored), and by the time it reaches its first frame its state agrees with the
// sequential recursion's to rounding.  Nothing is assumed about the forgetting rate: every chunk
// publishes the state it arrived with ("warm") and the state it handed to its neighbour ("exact"),
// a check kernel compares the two at every boundary, and any chain with a discrepancy above the
// tolerance is re-run by the sequential kernel (same code, one chunk) before anything consumes the
// result.  Masked steps carry no information, so chunks only cut the leading run of unmasked
// steps (vlen); the last chunk takes whatever follows.
// Slot c of the boundary buffers is the boundary between chunks c-1 and c, for forward and
// backward recursions alike.
// ---------------------------------------------------------------------------
struct ChunkRange {
    int begin, end, start;     // outputs for steps [begin, end); forward recursions start at `start` <= begin
    bool empty;
};
__host__ __device__ inline ChunkRange chunk_range(int vlen, int len, int C, int W, int c, int align = 1) {
    ChunkRange r;
    int Lc = (vlen + C - 1) / C;
    if (Lc < 4 * W) Lc = 4 * W;
    if (Lc < 1) Lc = 1;
    Lc = (Lc + align - 1) / align * align;           // chunk starts on a multiple of `align` (W must be one too)
    int Cn = (vlen + Lc - 1) / Lc;
    if (Cn < 1) Cn = 1;
    r.empty = c >= Cn;
    r.begin = c * Lc;
    r.end = (c == Cn - 1) ? len : (c + 1) * Lc;
    r.start = c > 0 ? r.begin - W : 0;
    return r;
}

// vlen[nn] = number of leading steps i in [0, len) with mask[nn][off + i] != 0;
// vlen[gridDim.x + nn] = one past the LAST unmasked step (= vlen for a chain whose only masked steps are its tail)
static __global__ void __launch_bounds__(256)
valid_len_kernel(const int* __restrict__ mask, int T, int off, int len, int* __restrict__ vlen) {
    __shared__ int first, last;
    const int nn = blockIdx.x;
    if (threadIdx.x == 0) { first = len; last = 0; }
    __syncthreads();
    const int* mk = mask + (size_t)nn * T + off;
    int mine_first = len, mine_last = 0;
    for (int i = threadIdx.x; i < len; i += blockDim.x) {
        if (mk[i] == 0) { if (mine_first == len) mine_first = i; }
        else mine_last = i + 1;
    }
    atomicMin(&first, mine_first);
    atomicMax(&last, mine_last);
    __syncthreads();
    if (threadIdx.x == 0) { vlen[nn] = first; vlen[gridDim.x + nn] = last; }
}

// Boundary check for real-valued states of `rec` numbers per boundary, in two blocks with their
// own scales: [0, n_mean) (difference relative to 1 + max|exact|) and [n_mean, rec) (relative to
// max|exact|).  One CTA per (boundary, chain); dirty[nn] (zeroed by the caller) is raised when the
// discrepancy exceeds tol or is not finite, or when pre[nn] != 0.
// COMPONENTWISE (HMM filter): every entry is compared relative to itself, |x - y| <= tol max(|x|, |y|).
// A probability vector that agrees component by component is within 2 tol in Hilbert's projective
// metric, under which the filter step is non-expansive, so every later filtered probability keeps that
// relative accuracy; an absolute comparison would let a negligible state with a large relative error
// pass and be amplified if the likelihood later favours it.
// stats[0] = max discrepancy seen (float bits, atomicMax), stats[1] += number of dirty chains.
template <typename R, bool COMPONENTWISE = false>
__global__ void __launch_bounds__(128)
boundary_check_kernel(const R* __restrict__ warm, const R* __restrict__ exact, const int* __restrict__ vlen,
                      int len, int C, int W, int align, int n_mean, int rec, R tol, int* __restrict__ dirty,
                      unsigned* __restrict__ stats, const int* __restrict__ pre = nullptr) {
    __shared__ R red[2][4];
    const int nn = blockIdx.y, c = blockIdx.x + 1, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    R worst = 0;
    int bad = 0;
    const bool flagged = (c == 1 && pre && pre[nn] != 0);
    const ChunkRange r = chunk_range(vlen[nn], len, C, W, c, align);
    if (c < C && !r.empty) {
        const R* a = warm + ((size_t)nn * C + c) * rec;
        const R* b = exact + ((size_t)nn * C + c) * rec;
        for (int part = 0; part < 2; ++part) {
            const int lo = part ? n_mean : 0, hi = part ? rec : n_mean;
            if (hi <= lo) continue;
            R diff = 0, scale = 0;
            for (int e = lo + threadIdx.x; e < hi; e += blockDim.x) {
                const R x = a[e], y = b[e];
                R dd = fabs(x - y);
                if (!(dd < (R)INFINITY)) bad = 1;
                if (COMPONENTWISE) {
                    const R big = fmax(fabs(x), fabs(y));
                    dd = big > (R)0 ? dd / big : (R)0;          // both exactly zero: equal
                }
                diff = fmax(diff, dd);
                scale = fmax(scale, fabs(y));
            }
            diff = warp_max(diff);
            scale = warp_max(scale);
            __syncthreads();
            if (lane == 0) { red[0][warp] = diff; red[1][warp] = scale; }
            __syncthreads();
            const R dmax = fmax(fmax(red[0][0], red[0][1]), fmax(red[0][2], red[0][3]));
            const R smax = fmax(fmax(red[1][0], red[1][1]), fmax(red[1][2], red[1][3]));
            worst = fmax(worst, COMPONENTWISE ? dmax : dmax / ((part ? (R)0 : (R)1) + smax + (R)1e-30));
        }
    }
    const int any_bad = __syncthreads_or(bad);
    if (threadIdx.x == 0) {
        const bool d = any_bad || !(worst <= tol) || flagged;
        if (worst > 0 || any_bad) atomicMax(&stats[0], __float_as_uint(any_bad ? INFINITY : (float)worst));
        if (d && atomicExch(&dirty[nn], 1) == 0) atomicAdd(&stats[1], 1u);
    }
}

// process-wide time-chunking configuration (kpms_set_time_chunking); see capi.cu
struct ChunkConfig { int chunks; int warmup; double tol32, tol64, tol_hmm; };
ChunkConfig chunk_config();
// chunks per chain for N chains when `slots` chunk-CTAs fit on the device at once
int chunks_for(int N, int slots, int len, int warmup);

}  // namespace kpms

// dtype dispatch for the C-ABI: 0 = f32, 1 = f64
#define KPMS_DISPATCH_DTYPE(dtype, FN, ...)                                            \
    ((dtype) == 0 ? FN<float>(__VA_ARGS__)                                             \
                  : ((dtype) == 1 ? FN<double>(__VA_ARGS__)                            \
                                  : kpms::set_error(-2, "dtype must be 0 (f32) or 1 (f64), got %d", (int)(dtype))))

// compile-time (latent_dim, nlags) instantiations.  The unrolled Kalman and likelihood kernels cost 20 - 60 s of
// nvcc time per pair, so the pairs are dealt into KPMS_DL_GROUPS groups and kalman.cu / ar_loglik.cu are compiled
// once per group (-DKPMS_DL_GROUP=g, build.py) in parallel; KPMS_FOR_EACH_DL lists them all for the cheap kernels
// and for the host-side table (kpms_supported_dims).  latent_dim 2..16 at the reference's default nlags = 3, the
// even dimensions at nlags 2, and 4..10 at nlags 1, 4 (and 4, 6 at nlags 5).
#define KPMS_DL_GROUPS 8
#define KPMS_DL_GROUP_0(X) X(10, 3) X(2, 2) X(2, 3) X(4, 1)
#define KPMS_DL_GROUP_1(X) X(16, 3) X(3, 3) X(4, 2) X(6, 1)
#define KPMS_DL_GROUP_2(X) X(15, 3) X(4, 3) X(6, 2) X(8, 1) X(4, 4)
#define KPMS_DL_GROUP_3(X) X(14, 3) X(5, 3) X(8, 2) X(10, 1) X(4, 5)
#define KPMS_DL_GROUP_4(X) X(13, 3) X(6, 3) X(10, 2) X(6, 4)
#define KPMS_DL_GROUP_5(X) X(12, 3) X(7, 3) X(12, 2) X(6, 5)
#define KPMS_DL_GROUP_6(X) X(11, 3) X(8, 3) X(16, 2) X(8, 4)
#define KPMS_DL_GROUP_7(X) X(9, 3) X(10, 4)
#define KPMS_FOR_EACH_DL(X)                                                                                       \
    KPMS_DL_GROUP_0(X) KPMS_DL_GROUP_1(X) KPMS_DL_GROUP_2(X) KPMS_DL_GROUP_3(X) KPMS_DL_GROUP_4(X) KPMS_DL_GROUP_5(X) \
    KPMS_DL_GROUP_6(X) KPMS_DL_GROUP_7(X)
#ifdef KPMS_DL_GROUP
#define KPMS_FOR_GROUP_DL(X) KPMS_CAT(KPMS_DL_GROUP_, KPMS_DL_GROUP)(X)
#endif
// return code of a group's dispatcher for a pair that belongs to another group
#define KPMS_NOT_IN_GROUP (-1000)
