// Error reporting and version for the C-ABI (include/kpms_b200.h).
#include <cstdarg>
#include <cstdio>
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

}  // namespace kpms

extern "C" {
int kpms_version(void) { return 100; }
const char* kpms_last_error(void) { return kpms::g_err; }
}
