// usermodel.cu -- user models registered as CUDA source through the C ABI (BASELINE.json north_star:
// "User models plug in either as a torch callable on device tensors or as a registered __device__
// model through the C-ABI").
//
// The user supplies the body of
//     __device__ double smcb_user_model(const double* theta, int j, double x, double* state)
// = model output at data point j (x = grid value or 0; `state` = 8 doubles per particle that
// persist over j, zero at j = 0: enough to step an ODE).  It is NVRTC-compiled for sm_100a together
// with a small kernel that evaluates the Normal log-likelihood (smcpy/log_likelihoods.py:22-49) of
// every particle: model outputs and residuals stay in registers.  The kernel replaces the
// "user model + smcb_normal_loglike_rows" pair of the un-fused Metropolis route
// (smcb_mh_propose -> log-likelihood -> smcb_mh_accept).
//
// libnvrtc and the driver API are dlopen'ed on first use: the library has no hard dependency on them.
#include <cuda.h>
#include <dlfcn.h>
#include <nvrtc.h>

#include <mutex>
#include <string>
#include <vector>

#include "common.cuh"

namespace smcb {

static const char* kUserKernelSrc = R"SRC(
extern "C" __global__ void __launch_bounds__(128)
smcb_user_ll(const double* __restrict__ params, long long stride, long long n, int d_model, int m,
             const double* __restrict__ data, const double* __restrict__ xgrid,
             const double* __restrict__ sigma_col, double sigma, const double* __restrict__ lp,
             double* __restrict__ ll, int* flags) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double ninf = __longlong_as_double(0xfff0000000000000LL);
    if (lp && lp[i] == ninf) {            // rows with a -inf prior keep 0 (vector_mcmc.py:205-210)
        ll[i] = 0.0;
        return;
    }
    double theta[SMCB_USER_D];
#pragma unroll
    for (int c = 0; c < SMCB_USER_D; ++c) theta[c] = (c < d_model) ? params[(long long)c * stride + i] : 0.0;
    double state[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    double ssq = 0.0;
    for (int j = 0; j < m; ++j) {
        const double out = smcb_user_model(theta, j, xgrid ? xgrid[j] : 0.0, state);
        const double dd = out - data[j];
        ssq = fma(dd, dd, ssq);
    }
    const double s = sigma_col ? sigma_col[i] : sigma;
    const double var = s * s;
    const double term1 = -log(2 * 3.141592653589793 * var) * ((double)m / 2.0);
    const double term2 = -1 / 2.0 * ssq / var;
    const double v = term1 + term2;
    ll[i] = v;
    if (v != v) atomicOr(flags, 2);       // SMCB_FLAG_MODEL_NAN
}
)SRC";

struct Nvrtc {
    void* h = nullptr;
    nvrtcResult (*create)(nvrtcProgram*, const char*, const char*, int, const char* const*, const char* const*);
    nvrtcResult (*compile)(nvrtcProgram, int, const char* const*);
    nvrtcResult (*cubin_size)(nvrtcProgram, size_t*);
    nvrtcResult (*cubin)(nvrtcProgram, char*);
    nvrtcResult (*log_size)(nvrtcProgram, size_t*);
    nvrtcResult (*log)(nvrtcProgram, char*);
    nvrtcResult (*destroy)(nvrtcProgram*);
};
struct Driver {
    void* h = nullptr;
    CUresult (*load)(CUmodule*, const void*);
    CUresult (*get_fn)(CUfunction*, CUmodule, const char*);
    CUresult (*launch)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, CUstream,
                       void**, void**);
    CUresult (*unload)(CUmodule);
};
struct UserModel {
    CUmodule mod;
    CUfunction fn;
    int d_model;
};

static Nvrtc g_nvrtc;
static Driver g_drv;
static std::vector<UserModel> g_models;
static std::mutex g_mu;

template <class F>
static bool sym(void* h, const char* name, F& f) {
    f = reinterpret_cast<F>(dlsym(h, name));
    return f != nullptr;
}

static int load_libs() {
    if (!g_nvrtc.h) {
        for (const char* n : {"libnvrtc.so.12", "libnvrtc.so"}) {
            g_nvrtc.h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (g_nvrtc.h) break;
        }
        if (!g_nvrtc.h) { set_error("libnvrtc not found: %s", dlerror()); return SMCB_ERR_UNSUPPORTED; }
        bool ok = sym(g_nvrtc.h, "nvrtcCreateProgram", g_nvrtc.create) && sym(g_nvrtc.h, "nvrtcCompileProgram", g_nvrtc.compile) &&
                  sym(g_nvrtc.h, "nvrtcGetCUBINSize", g_nvrtc.cubin_size) && sym(g_nvrtc.h, "nvrtcGetCUBIN", g_nvrtc.cubin) &&
                  sym(g_nvrtc.h, "nvrtcGetProgramLogSize", g_nvrtc.log_size) && sym(g_nvrtc.h, "nvrtcGetProgramLog", g_nvrtc.log) &&
                  sym(g_nvrtc.h, "nvrtcDestroyProgram", g_nvrtc.destroy);
        if (!ok) { set_error("libnvrtc lacks a required symbol"); g_nvrtc.h = nullptr; return SMCB_ERR_UNSUPPORTED; }
    }
    if (!g_drv.h) {
        for (const char* n : {"libcuda.so.1", "libcuda.so"}) {
            g_drv.h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (g_drv.h) break;
        }
        if (!g_drv.h) { set_error("libcuda not found: %s", dlerror()); return SMCB_ERR_UNSUPPORTED; }
        bool ok = sym(g_drv.h, "cuModuleLoadData", g_drv.load) && sym(g_drv.h, "cuModuleGetFunction", g_drv.get_fn) &&
                  sym(g_drv.h, "cuLaunchKernel", g_drv.launch) && sym(g_drv.h, "cuModuleUnload", g_drv.unload);
        if (!ok) { set_error("libcuda lacks a required symbol"); g_drv.h = nullptr; return SMCB_ERR_UNSUPPORTED; }
    }
    return SMCB_OK;
}

}  // namespace smcb

using namespace smcb;

extern "C" int smcb_model_register_nvrtc(const char* source, int32_t n_model_params, int32_t* handle_out) {
    SMCB_REQUIRE(source && handle_out && n_model_params >= 1 && n_model_params <= SMCB_MAX_PARAMS, "bad argument");
    std::lock_guard<std::mutex> lock(g_mu);
    int rc = load_libs();
    if (rc) return rc;
    std::string src = "#define SMCB_USER_D " + std::to_string(n_model_params) + "\n";
    src += "__device__ double smcb_user_model(const double* theta, int j, double x, double* state);\n";
    src += source;
    src += "\n";
    src += kUserKernelSrc;
    nvrtcProgram prog;
    if (g_nvrtc.create(&prog, src.c_str(), "smcb_user_model.cu", 0, nullptr, nullptr) != NVRTC_SUCCESS) {
        set_error("nvrtcCreateProgram failed");
        return SMCB_ERR_CUDA;
    }
    const char* opts[] = {"--gpu-architecture=sm_100a", "--std=c++17", "--fmad=true", "-lineinfo"};
    const nvrtcResult cr = g_nvrtc.compile(prog, 4, opts);
    if (cr != NVRTC_SUCCESS) {
        size_t ls = 0;
        g_nvrtc.log_size(prog, &ls);
        std::string log(ls + 1, '\0');
        if (ls) g_nvrtc.log(prog, &log[0]);
        set_error("NVRTC compile failed: %.400s", log.c_str());
        g_nvrtc.destroy(&prog);
        return SMCB_ERR_ARG;
    }
    size_t cs = 0;
    g_nvrtc.cubin_size(prog, &cs);
    std::vector<char> cubin(cs);
    g_nvrtc.cubin(prog, cubin.data());
    g_nvrtc.destroy(&prog);
    cudaFree(0);                                      // make the primary context current
    UserModel um;
    um.d_model = n_model_params;
    CUresult r = g_drv.load(&um.mod, cubin.data());
    if (r != CUDA_SUCCESS) { set_error("cuModuleLoadData failed (%d)", (int)r); return SMCB_ERR_CUDA; }
    r = g_drv.get_fn(&um.fn, um.mod, "smcb_user_ll");
    if (r != CUDA_SUCCESS) { set_error("cuModuleGetFunction failed (%d)", (int)r); return SMCB_ERR_CUDA; }
    g_models.push_back(um);
    *handle_out = (int32_t)g_models.size();
    return SMCB_OK;
}

extern "C" int smcb_user_loglike(int32_t handle, const double* params, int64_t stride, int64_t n, int32_t m,
                                 const double* data, const double* xgrid, const double* sigma_col, double sigma,
                                 const double* lp, double* ll_out, int32_t* flags, void* stream) {
    SMCB_REQUIRE(params && data && ll_out && flags && n > 0 && m > 0 && stride >= n, "bad argument");
    UserModel um;
    {
        std::lock_guard<std::mutex> lock(g_mu);
        SMCB_REQUIRE(handle >= 1 && handle <= (int32_t)g_models.size(), "unknown user-model handle");
        um = g_models[handle - 1];
    }
    long long stride_ll = stride, n_ll = n;
    int d_model = um.d_model, mm = m;
    void* args[] = {&params, &stride_ll, &n_ll, &d_model, &mm, &data, &xgrid, &sigma_col, &sigma, &lp, &ll_out, &flags};
    const unsigned grid = (unsigned)((n + 127) / 128);
    const CUresult r = g_drv.launch(um.fn, grid, 1, 1, 128, 1, 1, 0, (CUstream)stream, args, nullptr);
    if (r != CUDA_SUCCESS) { set_error("cuLaunchKernel(smcb_user_ll) failed (%d)", (int)r); return SMCB_ERR_CUDA; }
    return SMCB_OK;
}
