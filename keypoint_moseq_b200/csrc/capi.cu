// Error reporting and version for the C-ABI (include/kpms_b200.h).
#include <algorithm>
#include <cstdlib>
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>
#include "common.cuh"
#include "../../include/kpms_b200.h"

namespace kpms {

static thread_local char g_err[512] = "";

int set_error(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_error((int)e, "%s: %s", what, cudaGetErrorString(e));
    return 0;
}

// ---- launch accounting / optional per-kernel timing ----
struct ProfEntry { const char* name; cudaEvent_t a, b; };
static std::mutex g_mu;
static std::vector<ProfEntry> g_prof;
static std::atomic<long long> g_launches{0};
static bool g_profile = false;

LaunchScope::LaunchScope(const char* name, cudaStream_t st_) : slot(-1), st(st_) {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    if (g_profile) {
        ProfEntry e;
        e.name = name;
        cudaEventCreate(&e.a);
        cudaEventCreate(&e.b);
        cudaEventRecord(e.a, st);
        std::lock_guard<std::mutex> lk(g_mu);
        slot = (int)g_prof.size();
        g_prof.push_back(e);
    }
}

LaunchScope::~LaunchScope() {
    if (slot >= 0) {
        std::lock_guard<std::mutex> lk(g_mu);
        cudaEventRecord(g_prof[slot].b, st);
    }
}

// ---- time-chunking configuration (common.cuh) ----
// KPMS_WARMUP=<steps> overrides the default warm-up of the verified time chunks (experiments; kpms_set_time_chunking wins)
static ChunkConfig g_chunk = [] {
    ChunkConfig c = {0, 64, 2e-5, 1e-10, 1e-12};
    if (const char* e = getenv("KPMS_WARMUP")) { const int w = atoi(e); if (w >= 0 && w <= 1024) c.warmup = w; }
    return c;
}();
ChunkConfig chunk_config() { return g_chunk; }
// Chunks per chain for N chains when `slots` chunk tasks are resident on the device at once.  Tasks are equally long,
// so the kernel runs in ceil(N C / slots) waves of (len / C + warm-up) steps: C minimises that product.  With fewer
// chains than slots this is the old rule (one wave, C = slots / N: C2 -> 29); with about as many or more chains than
// slots (C4 on one GPU: 1200 chains, 2368 filter slots, 1184 HMM slots) one chunk per chain would leave half of a
// wave empty or spill a few tasks into a second full-length wave, and several shorter waves win.
int chunks_for(int N, int slots, int len, int warmup) {
    if (g_chunk.chunks > 0) return g_chunk.chunks > KPMS_MAX_CHUNKS ? KPMS_MAX_CHUNKS : g_chunk.chunks;
    if (N < 1) N = 1;
    if (slots < 1) slots = 1;
    int cmax = warmup > 0 ? len / (4 * warmup) : len;
    if (cmax > KPMS_MAX_CHUNKS) cmax = KPMS_MAX_CHUNKS;
    if (cmax < 1) cmax = 1;
    auto cost_of = [&](int C) {
        const long long waves = ((long long)N * C + slots - 1) / slots;
        const long long steps = (len + C - 1) / C + (C > 1 ? warmup : 0);
        return waves * steps;
    };
    long long best_cost = cost_of(1);
    for (int C = 2; C <= cmax && (long long)N * C <= 64LL * slots; ++C) best_cost = std::min(best_cost, cost_of(C));
    // the smallest C within 3 % of the optimum: the wave model ignores tails, so fewer, longer waves win a near-tie
    for (int C = 1; C <= cmax; ++C)
        if (cost_of(C) * 100 <= best_cost * 103) return C;
    return 1;
}

}  // namespace kpms

extern "C" {
void kpms_set_time_chunking(int chunks, int warmup, double tol32, double tol64) {
    if (chunks >= 0) kpms::g_chunk.chunks = chunks;
    if (warmup >= 0) kpms::g_chunk.warmup = warmup;
    if (tol32 > 0) kpms::g_chunk.tol32 = tol32;
    if (tol64 > 0) kpms::g_chunk.tol64 = tol64;
}

int kpms_plan_chunks(int N, int slots, int len, int warmup) { return kpms::chunks_for(N, slots, len, warmup); }

long long kpms_launch_count(void) { return kpms::g_launches.load(); }

void kpms_profile_enable(int on) { kpms::g_profile = on != 0; }

// Synchronises, then writes "name total_ms count\n" lines into buf and clears the records.
int kpms_profile_report(char* buf, size_t cap) {
    using namespace kpms;
    std::lock_guard<std::mutex> lk(g_mu);
    std::map<std::string, std::pair<double, int>> acc;
    for (auto& e : g_prof) {
        cudaEventSynchronize(e.b);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e.a, e.b);
        auto& slot = acc[e.name];
        slot.first += ms;
        slot.second += 1;
        cudaEventDestroy(e.a);
        cudaEventDestroy(e.b);
    }
    g_prof.clear();
    std::string out;
    for (auto& kv : acc) {
        char line[160];
        snprintf(line, sizeof(line), "%s %.6f %d\n", kv.first.c_str(), kv.second.first, kv.second.second);
        out += line;
    }
    if (out.size() + 1 > cap) return set_error(-4, "profile report needs %zu bytes", out.size() + 1);
    memcpy(buf, out.c_str(), out.size() + 1);
    return 0;
}

// splitmix64 step on the device-resident sweep key (same map as gibbs.advance_seed on the host)
static __global__ void advance_seed_kernel(uint64_t* seed) {
    uint64_t v = seed[0] + 0x9E3779B97F4A7C15ull;
    v ^= v >> 30;
    v *= 0xBF58476D1CE4E5B9ull;
    v ^= v >> 27;
    v *= 0x94D049BB133111EBull;
    v ^= v >> 31;
    seed[0] = v;
}

int kpms_advance_seed(uint64_t* seed_dev, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    { KPMS_LAUNCH("advance_seed", st); advance_seed_kernel<<<1, 1, 0, st>>>(seed_dev); }
    return kpms::check_launch("advance_seed");
}

int kpms_version(void) { return 101; }
const char* kpms_last_error(void) { return kpms::g_err; }
}
