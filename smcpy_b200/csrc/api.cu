// api.cu -- library-level entry points: version, error text, workspace sizing.
#include <stdarg.h>

#include "common.cuh"

namespace smcb {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int check_launch(const char* what) {
    const cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) return SMCB_OK;
    set_error("%s: %s", what, cudaGetErrorString(e));
    return SMCB_ERR_CUDA;
}

}  // namespace smcb

extern "C" int smcb_version(void) { return SMCB_VERSION; }

extern "C" const char* smcb_last_error(void) { return smcb::g_err; }

extern "C" size_t smcb_workspace_bytes(int64_t n, int32_t d) {
    (void)d;
    if (n < 1) n = 1;
    // tickets + block partials + small scalars + max(scan tile state, 1 flag byte per particle)
    const size_t tiles = (size_t)((n + 2047) / 2048) * 8;
    const size_t large = tiles > (size_t)n ? tiles : (size_t)n;
    return smcb::kWsLarge + large + 4096;
}
