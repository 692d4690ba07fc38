// api.cu -- library-level entry points: version, error text, workspace sizing.
#include <stdarg.h>
#include <stdlib.h>

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

static int env_int(const char* name, int dflt) {
    const char* e = getenv(name);
    return e ? atoi(e) : dflt;
}

Options& options() {
    static Options o = [] {
        Options v;
        v.no_tma = getenv("SMCB_NO_TMA") != nullptr;
        v.strict_proposal = env_int("SMCB_STRICT_PROPOSAL", 0);
        v.linear_cfg = env_int("SMCB_LINEAR_CFG", -1);
        v.persist = env_int("SMCB_PERSIST", -1);
        v.no_dmma = getenv("SMCB_NO_DMMA") != nullptr;
        v.cov_two_pass = getenv("SMCB_COV_TWO_PASS") != nullptr;
        v.no_coop = getenv("SMCB_NO_COOP") != nullptr;
        v.coop_tpb = env_int("SMCB_COOP_TPB", 0);
        v.coop_bps = env_int("SMCB_COOP_BPS", 0);
        v.bisect_plain = env_int("SMCB_BISECT_PLAIN", 0);
        v.spring_ppt = env_int("SMCB_SPRING_PPT", -1);
        v.no_mvn3 = env_int("SMCB_NO_MVN3", 0);
        v.no_pref = env_int("SMCB_NO_PREF", 0);
        v.xchg_timeout_s = env_int("SMCB_XCHG_TIMEOUT_S", 0);
        v.no_fused_resample = env_int("SMCB_NO_FUSED_RESAMPLE", 0);
        v.no_pdl = env_int("SMCB_NO_PDL", 0);
        v.spring_form = env_int("SMCB_SPRING_FORM", 1);
        return v;
    }();
    return o;
}

static thread_local char g_variant[256] = "";
void note_mh_variant(const char* text) {
    strncpy(g_variant, text, sizeof(g_variant) - 1);
    g_variant[sizeof(g_variant) - 1] = 0;
}

struct OptEntry {
    const char* name;
    int Options::*field;
};
static const OptEntry kOptTable[] = {
    {"no_tma", &Options::no_tma},           {"strict_proposal", &Options::strict_proposal},
    {"linear_cfg", &Options::linear_cfg},   {"persist", &Options::persist},
    {"no_dmma", &Options::no_dmma},         {"cov_two_pass", &Options::cov_two_pass},
    {"no_coop", &Options::no_coop},         {"coop_tpb", &Options::coop_tpb},
    {"coop_bps", &Options::coop_bps},       {"bisect_plain", &Options::bisect_plain},
    {"spring_ppt", &Options::spring_ppt},   {"xchg_timeout_s", &Options::xchg_timeout_s},
    {"no_fused_resample", &Options::no_fused_resample},
    {"no_mvn3", &Options::no_mvn3},
    {"no_pref", &Options::no_pref},
    {"no_pdl", &Options::no_pdl},
    {"spring_form", &Options::spring_form},
};

}  // namespace smcb

extern "C" int smcb_set_option(const char* name, int64_t value) {
    if (name)
        for (const auto& e : smcb::kOptTable)
            if (strcmp(e.name, name) == 0) {
                smcb::options().*(e.field) = (int)value;
                return SMCB_OK;
            }
    smcb::set_error("smcb_set_option: unknown option '%s'", name ? name : "(null)");
    return SMCB_ERR_ARG;
}

extern "C" int64_t smcb_get_option(const char* name) {
    if (name)
        for (const auto& e : smcb::kOptTable)
            if (strcmp(e.name, name) == 0) return smcb::options().*(e.field);
    return INT64_MIN;
}

extern "C" const char* smcb_last_mh_variant(void) { return smcb::g_variant; }

// FP64 FMA throughput probe: 8 independent FMA chains per thread (measurement aid for the
// FP64-pipe-bound kernels; MEASURED_PEAKS.json has no FP64 figure).
__global__ void __launch_bounds__(256) fp64_probe_kernel(double* out, int iters) {
    const double t = 1e-9 * (double)(threadIdx.x + 1);
    double a0 = t, a1 = 2 * t, a2 = 3 * t, a3 = 4 * t, a4 = 5 * t, a5 = 6 * t, a6 = 7 * t, a7 = 8 * t;
    const double m = 0.999999, c = 1e-7;
    for (int i = 0; i < iters; ++i) {
        a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
        a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
    }
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
}

extern "C" int smcb_fp64_probe(double* out, int32_t blocks, int32_t iters, void* stream) {
    if (!out || blocks < 1 || iters < 1) {
        smcb::set_error("smcb_fp64_probe: bad argument");
        return SMCB_ERR_ARG;
    }
    fp64_probe_kernel<<<blocks, 256, 0, smcb::as_stream(stream)>>>(out, iters);
    return smcb::check_launch("fp64_probe_kernel");
}

// FP64 tensor-core (DMMA, mma.sync.m8n8k4.f64) throughput probe: 4 independent accumulator pairs per warp,
// 4 x 4 DMMAs per iteration and warp, 512 flops each.  The larger of the two probes is the FP64 roof
// bench.py reports against (tools/probe/dmma_probe.cu: the two share one pipe).
__global__ void __launch_bounds__(256) dmma_probe_kernel(double* out, int iters, double a, double b) {
    double m0[4], m1[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        m0[i] = threadIdx.x * 1e-3 + i;
        m1[i] = i;
    }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                             : "+d"(m0[i]), "+d"(m1[i])
                             : "d"(a), "d"(b));
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) s += m0[i] + m1[i];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

extern "C" int smcb_dmma_probe(double* out, int32_t blocks, int32_t iters, void* stream) {
    if (!out || blocks < 1 || iters < 1) {
        smcb::set_error("smcb_dmma_probe: bad argument");
        return SMCB_ERR_ARG;
    }
    dmma_probe_kernel<<<blocks, 256, 0, smcb::as_stream(stream)>>>(out, iters, 0.999, 1e-9);
    return smcb::check_launch("dmma_probe_kernel");
}

__global__ void exp_probe_kernel(const double* x, double* out, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = smcb::exp_nonpos(x[i]);
}

extern "C" int smcb_exp_probe(const double* x, double* out, int64_t n, void* stream) {
    if (!x || !out || n < 1) {
        smcb::set_error("smcb_exp_probe: bad argument");
        return SMCB_ERR_ARG;
    }
    exp_probe_kernel<<<(unsigned)((n + 255) / 256), 256, 0, smcb::as_stream(stream)>>>(x, out, n);
    return smcb::check_launch("exp_probe_kernel");
}

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
