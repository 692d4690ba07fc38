// xchg.cu -- small all-gather / all-reduce of per-rank records over NVLink peer memory, one kernel.
//
// The hot path's cross-GPU couplings are a handful of scalars or one small matrix per SMC step
// (weight totals, covariance partial sums 1 + d + d^2 <= 1057 doubles, counts, error
// flags).  The reference moves them with MPI allgather / bcast (smcpy/mcmc/parallel_vector_mcmc.py:36-47,
// smcpy/utils/mpi_utils.py); a NCCL collective costs a launch plus a protocol round trip each.  Here
// every rank stores its record straight into a slot of EVERY peer's symmetric buffer (coalesced P2P
// stores), fences at system scope, stores a flag = generation, then waits for the `world` flags in its
// OWN memory and reads the records in rank order: the sum (reduce = 1) is the same add chain on every
// rank, i.e. bit-identical everywhere.  Slots are double-buffered by generation parity: a rank can be
// at most one call ahead of the slowest (it needs that rank's record to finish a call).
#include "common.cuh"

namespace smcb {

constexpr int kXgMaxVals = 1152;
constexpr int kXgRec = kXgMaxVals + 8;                      // payload + flag word, 64-byte padded
constexpr int kXgSlotWords = 2 * SMCB_MAX_RANKS * kXgRec;

__device__ __forceinline__ unsigned long long xg_ld(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void xg_st(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long xg_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

__global__ void __launch_bounds__(256) xchg_allgather_kernel(const smcb_xchg_t x, const double* __restrict__ src,
                                                             int n_vals, double* __restrict__ dst, int reduce,
                                                             int* status, unsigned long long timeout_ns) {
    __shared__ int s_ok;
    const int tid = threadIdx.x;
    const size_t off = (size_t)(x.gen & 1) * SMCB_MAX_RANKS * kXgRec;
    if (tid == 0) s_ok = 1;
    for (int r = 0; r < x.world; ++r) {
        unsigned long long* slot = reinterpret_cast<unsigned long long*>(x.peer_slots[r]) + off + (size_t)x.rank * kXgRec;
        for (int k = tid; k < n_vals; k += blockDim.x) xg_st(slot + k, (unsigned long long)__double_as_longlong(src[k]));
    }
    __threadfence_system();
    __syncthreads();
    if (tid < x.world)
        xg_st(reinterpret_cast<unsigned long long*>(x.peer_slots[tid]) + off + (size_t)x.rank * kXgRec + kXgMaxVals, x.gen);
    const unsigned long long* own = reinterpret_cast<const unsigned long long*>(x.peer_slots[x.rank]) + off;
    if (tid < x.world) {
        const unsigned long long t0 = xg_ns();
        unsigned int spins = 0;
        while (xg_ld(own + (size_t)tid * kXgRec + kXgMaxVals) != x.gen) {
            __nanosleep(40);
            if ((++spins & 1023u) == 0u && xg_ns() - t0 > timeout_ns) {
                s_ok = 0;
                break;
            }
        }
    }
    __threadfence_system();
    __syncthreads();
    if (!s_ok) {
        if (tid == 0 && status) atomicOr(status, 1);
        return;
    }
    if (reduce) {
        for (int k = tid; k < n_vals; k += blockDim.x) {
            double s = 0.0;
            for (int r = 0; r < x.world; ++r) s += __longlong_as_double((long long)xg_ld(own + (size_t)r * kXgRec + k));
            dst[k] = s;
        }
    } else {
        for (int r = 0; r < x.world; ++r)
            for (int k = tid; k < n_vals; k += blockDim.x)
                dst[(size_t)r * n_vals + k] = __longlong_as_double((long long)xg_ld(own + (size_t)r * kXgRec + k));
    }
}

}  // namespace smcb

using namespace smcb;

extern "C" int smcb_xchg_big_slot_words(void) { return kXgSlotWords; }
extern "C" int smcb_xchg_max_vals(void) { return kXgMaxVals; }

extern "C" int smcb_xchg_allgather(const smcb_xchg_t* x, const double* src, int32_t n_vals, double* dst,
                                   int32_t reduce, int32_t* status, void* stream) {
    SMCB_REQUIRE(x && src && dst && n_vals > 0 && n_vals <= kXgMaxVals, "bad argument");
    SMCB_REQUIRE(x->world >= 1 && x->world <= SMCB_MAX_RANKS && x->rank >= 0 && x->rank < x->world && x->gen > 0,
                 "bad exchange descriptor");
    for (int r = 0; r < x->world; ++r) SMCB_REQUIRE(x->peer_slots[r] != nullptr, "peer slot pointer is NULL");
    const int s = options().xchg_timeout_s;
    const unsigned long long tmo = (unsigned long long)(s > 0 ? s : 30) * 1000000000ull;
    xchg_allgather_kernel<<<1, 256, 0, as_stream(stream)>>>(*x, src, n_vals, dst, reduce, status, tmo);
    return check_launch("xchg_allgather_kernel");
}
