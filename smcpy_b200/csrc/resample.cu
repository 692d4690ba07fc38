// resample.cu -- kernel (c): resampling.
//   uniforms (standard / stratified / systematic)      smcpy/resampler_rngs.py:4-9
//   CDF = inclusive scan of the normalised weights      np.cumsum, smcpy/smc/updater.py:152
//   idx_k = #{j : cdf_j <= u_k}                         np.digitize, updater.py:152
//   gather of the particle columns                      updater.py:140-143
//   distinct-ancestor count                             updater.py:154-165
//
// Two scans: SEQUENTIAL reproduces np.cumsum's single fp64 add chain bit for bit (parity
// mode: bit-exact indices for the same weights and uniforms); LOOKBACK is a single-pass
// decoupled look-back scan whose cross-tile carries are 2^-61 fixed-point integers, so the
// result does not depend on the order in which tiles publish (deterministic run to run).
// Algorithmic bytes per particle: scan 16, search 16 (+ CDF probes, L2 resident), gather
// 16 (d + 2) + 8.
#include "common.cuh"

namespace smcb {

// ---- uniforms ------------------------------------------------------------------------------
__global__ void resample_uniforms_kernel(int kind, int64_t n_global, int64_t first, int64_t count,
                                         uint64_t seed, uint32_t stream_id, double* u) {
    const Philox rng(seed);
    const double inv_n = 1.0 / (double)n_global;          // resampler_rngs.py:9: 1 / N * (...)
    double u_sys = 0.0;
    if (kind == SMCB_RESAMPLE_SYSTEMATIC) {
        const uint4 r = rng(0, 0, stream_id);
        u_sys = u01_53(r.x, r.y);
    }
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride) {
        const int64_t k = first + i;
        double v;
        if (kind == SMCB_RESAMPLE_SYSTEMATIC) {
            v = __dmul_rn(inv_n, __dadd_rn((double)k, u_sys));
        } else {
            const uint4 r = rng((uint64_t)k, 0, stream_id);
            const double x = u01_53(r.x, r.y);
            v = (kind == SMCB_RESAMPLE_STRATIFIED) ? __dmul_rn(inv_n, __dadd_rn((double)k, x)) : x;
        }
        u[i] = v;
    }
}

// ---- sorted multinomial uniforms by exponential spacings ------------------------------------------
// N i.i.d. uniforms in sorted order are U_(k) = S_k / S_(N+1), S_k = E_1 + ... + E_k with i.i.d.
// unit exponentials: draw E per global slot (Philox keyed by the slot id: independent of the GPU
// count), prefix-scan them like the weights (local look-back scan + rank offsets), divide.  The
// slots come out sorted, so a sharded resampler needs no sort and its routing is a contiguous split.
// The exponentials are scaled by 1 / (2 N) so that every running sum stays below 1 (the look-back
// scan carries tile sums in 2^-61 fixed point); the ratio in the last step does not see the scale.
__global__ void resample_exponentials_kernel(int64_t first, int64_t count, int64_t n_global, uint64_t seed,
                                             uint32_t stream_id, double* e) {
    const Philox rng(seed);
    const double scale = 0.5 / (double)n_global;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride) {
        const uint4 r = rng((uint64_t)(first + i), 0xE590u, stream_id);
        e[i] = -log1p(-u01_53(r.x, r.y)) * scale;         // u in [0, 1): E in [0, 36.8]
    }
}
__global__ void resample_spacings_finish_kernel(const double* __restrict__ t, int64_t n, const double* offset,
                                                const double* total, int64_t n_global, uint64_t seed,
                                                uint32_t stream_id, double* __restrict__ u) {
    const Philox rng(seed);
    const uint4 r = rng((uint64_t)n_global, 0xE590u, stream_id);      // E_(N+1): same value in every thread / rank
    const double s_all = *total + -log1p(-u01_53(r.x, r.y)) * (0.5 / (double)n_global);
    const double off = *offset;
    const double below_one = 1.0 - 0x1p-53;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        u[i] = fmin((off + t[i]) / s_all, below_one);
}

// ---- sequential scan: bit-equal to np.cumsum ----------------------------------------------------
constexpr int kSeqChunk = 2048;
__global__ void __launch_bounds__(256) scan_sequential_kernel(const double* __restrict__ w,
                                                              double* __restrict__ cdf, int64_t n) {
    __shared__ double buf[kSeqChunk];
    __shared__ double carry;
    if (threadIdx.x == 0) carry = 0.0;
    for (int64_t base = 0; base < n; base += kSeqChunk) {
        const int m = (int)((n - base < kSeqChunk) ? (n - base) : kSeqChunk);
        __syncthreads();
        for (int j = threadIdx.x; j < m; j += blockDim.x) buf[j] = w[base + j];
        __syncthreads();
        if (threadIdx.x == 0) {
            double c = carry;
            const bool first = (base == 0);
            for (int j = 0; j < m; ++j) {
                // np.cumsum: out[0] = w[0]; out[i] = out[i-1] + w[i]
                c = (first && j == 0) ? buf[0] : __dadd_rn(c, buf[j]);
                buf[j] = c;
            }
            carry = c;
        }
        __syncthreads();
        for (int j = threadIdx.x; j < m; j += blockDim.x) cdf[base + j] = buf[j];
    }
}

// ---- decoupled look-back scan -------------------------------------------------------------------
constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanThreads * kScanItems;
constexpr double kFix = 2305843009213693952.0;          // 2^61
constexpr double kFixInv = 1.0 / 2305843009213693952.0;
constexpr unsigned long long kFlagAgg = 1ull, kFlagPrefix = 2ull;

__device__ __forceinline__ unsigned long long ld_state(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_state(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__global__ void scan_reset_kernel(unsigned long long* state, int64_t n_tiles, unsigned int* counter) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_tiles; i += stride) state[i] = 0ull;
    if (blockIdx.x == 0 && threadIdx.x == 0) *counter = 0u;
}

// Scan inputs.  LoadArr: an array of weights.  LoadRw: the normalised weights recomputed from the resident
// log-likelihood (and log-prior / previous log-weight) columns with the scalars of the reweight pass --
// w = exp((un_lw - max) - log S1), exactly rw_finalize_body's arithmetic -- so a step that resamples never
// writes or re-reads log_w / w.  LoadExp: the scaled unit exponentials of resample_exponentials_kernel,
// generated in registers.
struct LoadArr {
    const double* w;
    __device__ __forceinline__ void load8(int64_t base, int64_t n, double (&v)[8]) const {
        if (base + 8 <= n) {
            const double2* p = reinterpret_cast<const double2*>(w + base);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const double2 t = p[j];
                v[2 * j] = t.x;
                v[2 * j + 1] = t.y;
            }
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = (base + j < n) ? w[base + j] : 0.0;
        }
    }
};
struct LoadRw {
    RwIn in;
    const double* sc;        // scalars of the reweight pass: [0] max, [5] log S1
    __device__ __forceinline__ void load8(int64_t base, int64_t n, double (&v)[8]) const {
        const double m = sc[0], ls = sc[5];
        double x[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) x[j] = (base + j < n) ? rw_value(in, base + j) : kNegInf;
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = (base + j < n) ? exp((x[j] - m) - ls) : 0.0;
    }
};
struct LoadExp {
    int64_t first, n_global;
    uint64_t seed;
    uint32_t stream_id;
    __device__ __forceinline__ void load8(int64_t base, int64_t n, double (&v)[8]) const {
        const Philox rng(seed);
        const double scale = 0.5 / (double)n_global;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            double e = 0.0;
            if (base + j < n) {
                const uint4 r = rng((uint64_t)(first + base + j), 0xE590u, stream_id);
                e = -log1p(-u01_53(r.x, r.y)) * scale;
            }
            v[j] = e;
        }
    }
};

template <class Loader>
__global__ void __launch_bounds__(kScanThreads) scan_lookback_kernel(const Loader ld,
                                                                     double* __restrict__ cdf, int64_t n,
                                                                     unsigned long long* state,
                                                                     unsigned int* counter, double* total_out) {
    __shared__ double sh_warp[kScanThreads / 32];
    __shared__ double sh_wmax[kScanThreads / 32];
    __shared__ unsigned int sh_tile;
    __shared__ double sh_prefix, sh_prefix_incl;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;

    // tiles are handed out in launch order so every predecessor is already resident
    if (tid == 0) sh_tile = atomicAdd(counter, 1u);
    __syncthreads();
    const int64_t tile = sh_tile;
    const int64_t base = tile * kScanTile + (int64_t)tid * kScanItems;

    static_assert(kScanItems == 8, "loaders fetch eight values per thread");
    double v[kScanItems];
    ld.load8(base, n, v);
    // thread-local inclusive scan
#pragma unroll
    for (int j = 1; j < kScanItems; ++j) v[j] += v[j - 1];
    const double tsum = v[kScanItems - 1];
    // warp inclusive scan of thread totals
    double x = tsum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const double y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
    }
    if (lane == 31) sh_warp[wid] = x;
    double xprev = __shfl_up_sync(0xffffffffu, x, 1);
    if (lane == 0) xprev = 0.0;
    __syncthreads();
    double warp_off = 0.0, tile_agg = 0.0;
#pragma unroll
    for (int k = 0; k < kScanThreads / 32; ++k) {
        const double s = sh_warp[k];
        if (k < wid) warp_off += s;
        tile_agg += s;
    }
    const double thread_excl = warp_off + xprev;

    // cross-tile carry by decoupled look-back on fixed-point aggregates
    if (wid == 0) {
        const unsigned long long agg_fx = __double2ull_rn(tile_agg * kFix);
        unsigned long long excl = 0ull;
        if (tile == 0) {
            if (lane == 0) st_state(state, (agg_fx << 2) | kFlagPrefix);
        } else {
            if (lane == 0) st_state(state + tile, (agg_fx << 2) | kFlagAgg);
            int64_t look = tile - 1;
            while (true) {
                const int64_t idx = look - lane;
                unsigned long long word = kFlagPrefix;            // virtual predecessor: prefix 0
                if (idx >= 0) {
                    word = ld_state(state + idx);
                    while ((word & 3ull) == 0ull) word = ld_state(state + idx);
                }
                const unsigned pmask = __ballot_sync(0xffffffffu, (word & 3ull) == kFlagPrefix);
                const int firstp = pmask ? (__ffs(pmask) - 1) : 31;  // nearest predecessor with a prefix
                const unsigned long long val = (lane <= firstp) ? (word >> 2) : 0ull;
                excl += warp_sum_ull(val);
                if (pmask) break;                                  // (always true once idx < 0 appears)
                look -= 32;
            }
            if (lane == 0) st_state(state + tile, ((excl + agg_fx) << 2) | kFlagPrefix);
        }
        if (lane == 0) {
            sh_prefix = (double)excl * kFixInv;
            sh_prefix_incl = (double)(excl + agg_fx) * kFixInv;
        }
    }
    __syncthreads();
    // candidate values, clamped into [tile prefix, next tile prefix] so tiles are ordered
    const double off = sh_prefix + thread_excl;
    const double hi = sh_prefix_incl;
#pragma unroll
    for (int j = 0; j < kScanItems; ++j) v[j] = fmin(off + v[j], hi);
    // exact (associative) running maximum across threads makes the CDF monotone by construction
    double m = v[kScanItems - 1];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const double y = __shfl_up_sync(0xffffffffu, m, o);
        if (lane >= o) m = fmax(m, y);
    }
    if (lane == 31) sh_wmax[wid] = m;
    double lower = __shfl_up_sync(0xffffffffu, m, 1);
    if (lane == 0) lower = 0.0;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kScanThreads / 32; ++k)
        if (k < wid) lower = fmax(lower, sh_wmax[k]);
#pragma unroll
    for (int j = 0; j < kScanItems; ++j) v[j] = fmax(v[j], lower);
    if (total_out && base <= n - 1 && n - 1 < base + kScanItems) {       // the running sum of the whole shard
#pragma unroll
        for (int j = 0; j < kScanItems; ++j)
            if (base + j == n - 1) *total_out = v[j];
    }
    if (base + kScanItems <= n) {
        double2* q = reinterpret_cast<double2*>(cdf + base);
#pragma unroll
        for (int j = 0; j < kScanItems / 2; ++j) q[j] = make_double2(v[2 * j], v[2 * j + 1]);
    } else {
#pragma unroll
        for (int j = 0; j < kScanItems; ++j)
            if (base + j < n) cdf[base + j] = v[j];
    }
}

// ---- search -----------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) searchsorted_kernel(const double* __restrict__ cdf, int64_t n_cdf,
                                                           const double* __restrict__ offset_ptr,
                                                           const double* __restrict__ u, int64_t n_u,
                                                           int64_t* __restrict__ idx) {
    const double off = offset_ptr ? *offset_ptr : 0.0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n_u; k += stride) {
        const double x = u[k];
        int64_t lo = 0, hi = n_cdf;                 // first j with cdf_j > x
        while (lo < hi) {
            const int64_t mid = (lo + hi) >> 1;
            const double c = offset_ptr ? __dadd_rn(off, __ldg(cdf + mid)) : __ldg(cdf + mid);
            if (c <= x) lo = mid + 1; else hi = mid;
        }
        idx[k] = (lo < n_cdf) ? lo : (n_cdf - 1);   // clamp (quirk Q4)
    }
}

// ---- gather -----------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gather_kernel(const double* __restrict__ src, int64_t src_stride,
                                                     double* __restrict__ dst, int64_t dst_stride,
                                                     const int64_t* __restrict__ idx, int64_t n_out,
                                                     int n_cols) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n_out; k += stride) {
        const int64_t a = idx[k];
        int c = 0;
        for (; c + 4 <= n_cols; c += 4) {           // four independent loads in flight
            const double v0 = __ldg(src + (int64_t)(c + 0) * src_stride + a);
            const double v1 = __ldg(src + (int64_t)(c + 1) * src_stride + a);
            const double v2 = __ldg(src + (int64_t)(c + 2) * src_stride + a);
            const double v3 = __ldg(src + (int64_t)(c + 3) * src_stride + a);
            dst[(int64_t)(c + 0) * dst_stride + k] = v0;
            dst[(int64_t)(c + 1) * dst_stride + k] = v1;
            dst[(int64_t)(c + 2) * dst_stride + k] = v2;
            dst[(int64_t)(c + 3) * dst_stride + k] = v3;
        }
        for (; c < n_cols; ++c) dst[(int64_t)c * dst_stride + k] = __ldg(src + (int64_t)c * src_stride + a);
    }
}

// ---- fused gather + all-to-all over peer memory ---------------------------------------------------------
struct PeerDst {
    double* p[SMCB_MAX_RANKS];
};
__global__ void __launch_bounds__(256) gather_p2p_kernel(const double* __restrict__ src, int64_t src_stride,
                                                         const int64_t* __restrict__ idx, int64_t n_req, int n_cols,
                                                         int world, const int64_t* __restrict__ block_off,
                                                         const int64_t* __restrict__ slot_off, PeerDst dst,
                                                         int64_t dst_stride) {
    __shared__ int64_t s_boff[SMCB_MAX_RANKS + 1], s_soff[SMCB_MAX_RANKS];
    if (threadIdx.x <= world) s_boff[threadIdx.x] = block_off[threadIdx.x];
    if (threadIdx.x < world) s_soff[threadIdx.x] = slot_off[threadIdx.x];
    __syncthreads();
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n_req; j += stride) {
        int q = 0;
        while (q + 1 < world && j >= s_boff[q + 1]) ++q;            // requester rank of request j
        const int64_t row = s_soff[q] + (j - s_boff[q]);
        const int64_t a = idx[j];
        double* out = dst.p[q] + row;
        int c = 0;
        for (; c + 4 <= n_cols; c += 4) {
            const double v0 = __ldg(src + (int64_t)(c + 0) * src_stride + a);
            const double v1 = __ldg(src + (int64_t)(c + 1) * src_stride + a);
            const double v2 = __ldg(src + (int64_t)(c + 2) * src_stride + a);
            const double v3 = __ldg(src + (int64_t)(c + 3) * src_stride + a);
            out[(int64_t)(c + 0) * dst_stride] = v0;
            out[(int64_t)(c + 1) * dst_stride] = v1;
            out[(int64_t)(c + 2) * dst_stride] = v2;
            out[(int64_t)(c + 3) * dst_stride] = v3;
        }
        for (; c < n_cols; ++c) out[(int64_t)c * dst_stride] = __ldg(src + (int64_t)c * src_stride + a);
    }
}

// ---- distinct ancestors -------------------------------------------------------------------------------
__global__ void mark_kernel(const int64_t* __restrict__ idx, int64_t n_idx, uint8_t* flags) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n_idx; k += stride) flags[idx[k]] = 1;
}

__global__ void __launch_bounds__(256) count_bytes_kernel(const uint8_t* __restrict__ flags, int64_t n,
                                                          unsigned long long* acc, unsigned int* ticket,
                                                          int64_t* out) {
    unsigned long long c = 0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride) c += (flags[k] != 0);
    c = warp_sum_ull(c);
    __shared__ unsigned long long sh[8];
    __shared__ bool is_last;
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long t = 0;
        for (int k = 0; k < (int)(blockDim.x >> 5); ++k) t += sh[k];
        atomicAdd(acc, t);                          // integer: order-independent
        __threadfence();
        is_last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
        if (is_last) {
            *out = (int64_t)atomicAdd(acc, 0ull);
            *acc = 0ull;
            *ticket = 0u;
        }
    }
}

// ---- fused resampling for sorted uniforms (native generator) -------------------------------------------------
// All three built-in resamplers produce their uniforms in non-decreasing slot order (stratified and
// systematic by construction, multinomial by exponential spacings), so
//   * slot k's ancestor index is found by a binary search confined to the CDF window that the first
//     slots of this tile and of the next one select (merge-path style partition, rs_tile_bounds_kernel):
//     ~log2(tile) probes that stay in L1 instead of ~23 probes through L2;
//   * the uniforms themselves are evaluated on the fly (Philox keyed by the global slot, or the running
//     sums of the exponentials), the ancestor rows are gathered in the same pass, and the number of distinct
//     ancestors is the number of index changes between neighbouring slots.
// One rank of `world` (particles sharded in equal contiguous blocks): the rank OWNS the global slots whose
// uniform falls into its segment [O_r, O_r+1) of the global CDF -- a contiguous slot range [K_lo, K_hi)
// that it derives itself from the all-gathered rank totals (rs_plan_kernel) -- searches its own CDF and
// writes every selected row column by column straight into the SoA buffer of the rank that holds the slot
// (peer-mapped pointers over NVLink; its own buffer when world == 1).  No request routing, no host
// round trip.
constexpr int kRsThreads = 256;
constexpr int kRsPpt = 4;
constexpr int kRsTile = kRsThreads * kRsPpt;

struct RsDev {
    int32_t kind, world, rank, n_cols;
    int64_t n_local, n_global;
    const double* cdf;
    const double* totals;                    // world x 2
    const double* t_peer[SMCB_MAX_RANKS];
    const double* src;
    int64_t src_stride;
    double* dst_peer[SMCB_MAX_RANKS];
    int64_t dst_stride;
    int64_t* idx_out;
    int64_t* n_unique_out;
    uint64_t seed;
    uint32_t stream_id;
    // workspace
    double* w_off;                           // world + 1 CDF segment offsets
    double* e_off;                           // world + 1 offsets of the exponential running sums, [world + 1] = S_(N+1)
    int64_t* krange;                         // [0] K_lo, [1] K_hi
    int64_t* tb;                             // per tile: ancestor index of its first slot (+ one closing entry)
    int64_t* tp;                             // per tile: ancestor index of the slot before its first
    unsigned long long* uniq;
    unsigned int* ticket;
};

__device__ __forceinline__ double rs_slot_u(const RsDev& a, const Philox& rng, double u_sys, double s_all, int64_t k) {
    const double inv_n = 1.0 / (double)a.n_global;          // resampler_rngs.py:9: 1 / N * (...)
    if (a.kind == SMCB_RESAMPLE_SYSTEMATIC) return __dmul_rn(inv_n, __dadd_rn((double)k, u_sys));
    if (a.kind == SMCB_RESAMPLE_STRATIFIED) {
        const uint4 r = rng((uint64_t)k, 0, a.stream_id);
        return __dmul_rn(inv_n, __dadd_rn((double)k, u01_53(r.x, r.y)));
    }
    const int64_t q = k / a.n_local, j = k - q * a.n_local;  // standard: U_(k) = S_k / S_(N+1)
    return fmin((a.e_off[q] + a.t_peer[q][j]) / s_all, 1.0 - 0x1p-53);
}
__device__ __forceinline__ double rs_u_sys(const RsDev& a, const Philox& rng) {
    if (a.kind != SMCB_RESAMPLE_SYSTEMATIC) return 0.0;
    const uint4 r = rng(0, 0, a.stream_id);
    return u01_53(r.x, r.y);
}
// first j in [lo, hi) with off + cdf_j > x; hi if none
__device__ __forceinline__ int64_t rs_search(const double* __restrict__ cdf, double off, double x, int64_t lo, int64_t hi) {
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        const double c = __dadd_rn(off, __ldg(cdf + mid));
        if (c <= x) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// offsets from the gathered totals (rank order, one add chain: identical on every rank), S_(N+1), and the
// slot range this rank owns
__global__ void rs_plan_kernel(const RsDev a) {
    const int t = threadIdx.x;
    if (t == 0) {
        double o = 0.0, e = 0.0;
        a.w_off[0] = 0.0;
        a.e_off[0] = 0.0;
        for (int r = 0; r < a.world; ++r) {
            // single GPU without a totals record: the running sums' last entries
            const double tw = a.totals ? a.totals[2 * r] : a.cdf[a.n_local - 1];
            const double te = a.totals ? a.totals[2 * r + 1]
                                       : (a.kind == SMCB_RESAMPLE_STANDARD ? a.t_peer[0][a.n_local - 1] : 0.0);
            o = __dadd_rn(o, tw);
            e = __dadd_rn(e, te);
            a.w_off[r + 1] = o;
            a.e_off[r + 1] = e;
        }
        const Philox rng(a.seed);
        const uint4 r4 = rng((uint64_t)a.n_global, 0xE590u, a.stream_id);      // E_(N+1): same value on every rank
        a.e_off[a.world + 1] = e + -log1p(-u01_53(r4.x, r4.y)) * (0.5 / (double)a.n_global);
    }
    __syncthreads();
    if (t < 2) {
        const Philox rng(a.seed);
        const double u_sys = rs_u_sys(a, rng);
        const double s_all = a.e_off[a.world + 1];
        const int seg = a.rank + t;                            // boundary O_seg: slots with u < O_seg lie below it
        int64_t k;
        if (seg == 0) k = 0;
        else if (seg == a.world) k = a.n_global;
        else {
            const double bound = a.w_off[seg];
            int64_t lo = 0, hi = a.n_global;
            while (lo < hi) {
                const int64_t mid = (lo + hi) >> 1;
                if (rs_slot_u(a, rng, u_sys, s_all, mid) < bound) lo = mid + 1; else hi = mid;
            }
            k = lo;
        }
        a.krange[t] = k;
    }
}

__global__ void __launch_bounds__(256) rs_tile_bounds_kernel(const RsDev a, int64_t max_tiles) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t k_lo = a.krange[0], k_hi = a.krange[1];
    const int64_t cnt = k_hi - k_lo;
    const int64_t n_tiles = (cnt + kRsTile - 1) / kRsTile;
    if (cnt <= 0 || t > n_tiles || t > max_tiles) return;
    const Philox rng(a.seed);
    const double u_sys = rs_u_sys(a, rng);
    const double s_all = a.e_off[a.world + 1];
    const double off = a.w_off[a.rank];
    int64_t slot = k_lo + t * kRsTile;
    if (slot >= k_hi) slot = k_hi - 1;
    a.tb[t] = rs_search(a.cdf, off, rs_slot_u(a, rng, u_sys, s_all, slot), 0, a.n_local);
    int64_t prev = -1;                                         // the slot before the range belongs to another owner
    if (t > 0 && t < n_tiles) {
        prev = rs_search(a.cdf, off, rs_slot_u(a, rng, u_sys, s_all, k_lo + t * kRsTile - 1), 0, a.n_local);
        if (prev > a.n_local - 1) prev = a.n_local - 1;
    }
    a.tp[t] = prev;
}

__global__ void __launch_bounds__(kRsThreads) rs_fused_kernel(const RsDev a) {
    __shared__ int64_t sidx[kRsTile];
    __shared__ unsigned int s_cnt;
    const int tid = threadIdx.x;
    const int64_t k_lo = a.krange[0], k_hi = a.krange[1];
    const int64_t cnt = k_hi - k_lo;
    const int64_t n_tiles = cnt > 0 ? (cnt + kRsTile - 1) / kRsTile : 0;
    const Philox rng(a.seed);
    const double u_sys = rs_u_sys(a, rng);
    const double s_all = a.e_off[a.world + 1];
    const double off = a.w_off[a.rank];
    unsigned int my_unique = 0;
    if (tid == 0) s_cnt = 0;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t lo = a.tb[tile], hi = a.tb[tile + 1];
        const int64_t base = k_lo + tile * kRsTile;
        int64_t idx[kRsPpt];
        bool valid[kRsPpt];
        double* dstp[kRsPpt];
#pragma unroll
        for (int p = 0; p < kRsPpt; ++p) {
            const int s = p * kRsThreads + tid;
            const int64_t k = base + s;
            valid[p] = k < k_hi;
            int64_t j = -2;
            dstp[p] = nullptr;
            if (valid[p]) {
                const double x = rs_slot_u(a, rng, u_sys, s_all, k);
                j = rs_search(a.cdf, off, x, lo, hi);
                if (j > a.n_local - 1) j = a.n_local - 1;       // clamp (quirk Q4)
                const int64_t q = k / a.n_local;
                dstp[p] = a.dst_peer[q] + (k - q * a.n_local);
                if (a.idx_out) a.idx_out[k - (int64_t)a.rank * a.n_local] = j;    // world == 1 only
            }
            idx[p] = j;
            sidx[s] = j;
        }
        __syncthreads();
#pragma unroll
        for (int p = 0; p < kRsPpt; ++p) {
            const int s = p * kRsThreads + tid;
            const int64_t prev = (s == 0) ? a.tp[tile] : sidx[s - 1];
            if (valid[p] && idx[p] != prev) ++my_unique;
        }
        // gather: two columns x kRsPpt slots in flight
        int c = 0;
        for (; c + 2 <= a.n_cols; c += 2) {
            double v0[kRsPpt], v1[kRsPpt];
            const double* s0 = a.src + (int64_t)c * a.src_stride;
            const double* s1 = s0 + a.src_stride;
#pragma unroll
            for (int p = 0; p < kRsPpt; ++p) {
                v0[p] = valid[p] ? __ldg(s0 + idx[p]) : 0.0;
                v1[p] = valid[p] ? __ldg(s1 + idx[p]) : 0.0;
            }
#pragma unroll
            for (int p = 0; p < kRsPpt; ++p) {
                if (valid[p]) {
                    dstp[p][(int64_t)c * a.dst_stride] = v0[p];
                    dstp[p][(int64_t)(c + 1) * a.dst_stride] = v1[p];
                }
            }
        }
        if (c < a.n_cols) {
            const double* s0 = a.src + (int64_t)c * a.src_stride;
#pragma unroll
            for (int p = 0; p < kRsPpt; ++p)
                if (valid[p]) dstp[p][(int64_t)c * a.dst_stride] = __ldg(s0 + idx[p]);
        }
        __syncthreads();                                        // sidx is reused by the next tile
    }
    // distinct ancestors: block sum, one integer atomic per block; the last block publishes and resets
    my_unique = __reduce_add_sync(0xffffffffu, my_unique);
    if ((tid & 31) == 0 && my_unique) atomicAdd(&s_cnt, my_unique);
    __syncthreads();
    if (tid == 0) {
        if (s_cnt) atomicAdd(a.uniq, (unsigned long long)s_cnt);
        __threadfence();
        if (atomicAdd(a.ticket, 1u) == gridDim.x - 1) {
            *a.n_unique_out = (int64_t)atomicAdd(a.uniq, 0ull);
            *a.uniq = 0ull;
            *a.ticket = 0u;
        }
    }
}

}  // namespace smcb

using namespace smcb;

extern "C" int smcb_resample_uniforms(int32_t kind, int64_t n_global, int64_t first, int64_t count,
                                      uint64_t seed, uint64_t stream_id, double* u_out, void* stream) {
    SMCB_REQUIRE(kind >= 0 && kind <= 2 && n_global > 0 && count > 0 && u_out, "bad argument");
    const int grid = grid_for(count, 256, 4, 8);
    resample_uniforms_kernel<<<grid, 256, 0, as_stream(stream)>>>(kind, n_global, first, count, seed,
                                                                   (uint32_t)stream_id, u_out);
    return check_launch("resample_uniforms_kernel");
}

extern "C" int smcb_resample_exponentials(int64_t first, int64_t count, int64_t n_global, uint64_t seed,
                                          uint64_t stream_id, double* e_out, void* stream) {
    SMCB_REQUIRE(first >= 0 && count > 0 && n_global >= first + count && e_out, "bad argument");
    resample_exponentials_kernel<<<grid_for(count, 256, 4, 8), 256, 0, as_stream(stream)>>>(
        first, count, n_global, seed, (uint32_t)stream_id, e_out);
    return check_launch("resample_exponentials_kernel");
}

extern "C" int smcb_resample_spacings_finish(const double* t, int64_t n, const double* offset, const double* total,
                                             int64_t n_global, uint64_t seed, uint64_t stream_id, double* u_out,
                                             void* stream) {
    SMCB_REQUIRE(t && n > 0 && offset && total && n_global >= n && u_out, "bad argument");
    resample_spacings_finish_kernel<<<grid_for(n, 256, 4, 8), 256, 0, as_stream(stream)>>>(
        t, n, offset, total, n_global, seed, (uint32_t)stream_id, u_out);
    return check_launch("resample_spacings_finish_kernel");
}

extern "C" int smcb_cdf_scan(const double* w, double* cdf, int64_t n, int32_t mode, void* ws, void* stream) {
    SMCB_REQUIRE(w && cdf && n > 0 && ws, "bad argument");
    if (mode == SMCB_SCAN_SEQUENTIAL) {
        scan_sequential_kernel<<<1, 256, 0, as_stream(stream)>>>(w, cdf, n);
        return check_launch("scan_sequential_kernel");
    }
    SMCB_REQUIRE(mode == SMCB_SCAN_LOOKBACK, "unknown scan mode");
    SMCB_REQUIRE((((uintptr_t)w | (uintptr_t)cdf) & 15) == 0, "w and cdf must be 16-byte aligned");
    char* base = (char*)ws;
    unsigned long long* state = (unsigned long long*)(base + kWsLarge);
    unsigned int* counter = (unsigned int*)(base + kWsTickets) + 1;
    const int64_t n_tiles = (n + kScanTile - 1) / kScanTile;
    scan_reset_kernel<<<grid_for(n_tiles, 256, 1, 4), 256, 0, as_stream(stream)>>>(state, n_tiles, counter);
    scan_lookback_kernel<LoadArr><<<(unsigned)n_tiles, kScanThreads, 0, as_stream(stream)>>>(LoadArr{w}, cdf, n, state,
                                                                                               counter, nullptr);
    return check_launch("scan_lookback_kernel");
}

extern "C" int smcb_searchsorted(const double* cdf, int64_t n_cdf, const double* u, int64_t n_u,
                                 int64_t* idx_out, void* stream) {
    SMCB_REQUIRE(cdf && u && idx_out && n_cdf > 0 && n_u > 0, "bad argument");
    searchsorted_kernel<<<grid_for(n_u, 256, 1, 8), 256, 0, as_stream(stream)>>>(cdf, n_cdf, nullptr, u, n_u, idx_out);
    return check_launch("searchsorted_kernel");
}

extern "C" int smcb_searchsorted_offset(const double* cdf, int64_t n_cdf, const double* offset, const double* u,
                                        int64_t n_u, int64_t* idx_out, void* stream) {
    SMCB_REQUIRE(cdf && offset && u && idx_out && n_cdf > 0 && n_u > 0, "bad argument");
    searchsorted_kernel<<<grid_for(n_u, 256, 1, 8), 256, 0, as_stream(stream)>>>(cdf, n_cdf, offset, u, n_u, idx_out);
    return check_launch("searchsorted_kernel");
}

extern "C" int smcb_gather(const double* src, int64_t src_stride, double* dst, int64_t dst_stride,
                           const int64_t* idx, int64_t n_out, int32_t n_cols, void* stream) {
    SMCB_REQUIRE(src && dst && idx && n_out > 0 && n_cols > 0, "bad argument");
    gather_kernel<<<grid_for(n_out, 256, 1, 8), 256, 0, as_stream(stream)>>>(src, src_stride, dst, dst_stride,
                                                                              idx, n_out, n_cols);
    return check_launch("gather_kernel");
}

extern "C" int smcb_gather_p2p(const double* src, int64_t src_stride, const int64_t* idx, int64_t n_req,
                               int32_t n_cols, int32_t world, const int64_t* block_off, const int64_t* slot_off,
                               double* const* dst, int64_t dst_stride, void* stream) {
    SMCB_REQUIRE(src && idx && block_off && slot_off && dst && n_req > 0 && n_cols > 0, "bad argument");
    SMCB_REQUIRE(world >= 1 && world <= SMCB_MAX_RANKS, "world out of range");
    PeerDst d;
    for (int r = 0; r < SMCB_MAX_RANKS; ++r) d.p[r] = (r < world) ? dst[r] : nullptr;
    for (int r = 0; r < world; ++r) SMCB_REQUIRE(d.p[r] != nullptr, "peer destination pointer is NULL");
    gather_p2p_kernel<<<grid_for(n_req, 256, 1, 8), 256, 0, as_stream(stream)>>>(src, src_stride, idx, n_req, n_cols,
                                                                                  world, block_off, slot_off, d,
                                                                                  dst_stride);
    return check_launch("gather_p2p_kernel");
}

extern "C" int smcb_count_unique(const int64_t* idx, int64_t n_idx, int64_t n_src, int64_t* count_out,
                                 void* ws, void* stream) {
    SMCB_REQUIRE(idx && count_out && ws && n_idx > 0 && n_src > 0, "bad argument");
    char* base = (char*)ws;
    uint8_t* flags = (uint8_t*)(base + kWsLarge);
    cudaError_t e = cudaMemsetAsync(flags, 0, (size_t)n_src, as_stream(stream));
    if (e != cudaSuccess) {
        set_error("smcb_count_unique: memset: %s", cudaGetErrorString(e));
        return SMCB_ERR_CUDA;
    }
    mark_kernel<<<grid_for(n_idx, 256, 1, 8), 256, 0, as_stream(stream)>>>(idx, n_idx, flags);
    count_bytes_kernel<<<grid_for(n_src, 256, 8, 4), 256, 0, as_stream(stream)>>>(
        flags, n_src, (unsigned long long*)(base + kWsSmall + 64), (unsigned int*)(base + kWsTickets) + 2,
        count_out);
    return check_launch("count_bytes_kernel");
}

extern "C" int smcb_count_moved(const uint8_t* moved, int64_t n, int64_t* count_out, void* ws, void* stream) {
    SMCB_REQUIRE(moved && count_out && ws && n > 0, "bad argument");
    char* base = (char*)ws;
    count_bytes_kernel<<<grid_for(n, 256, 8, 4), 256, 0, as_stream(stream)>>>(
        moved, n, (unsigned long long*)(base + kWsSmall + 64), (unsigned int*)(base + kWsTickets) + 2,
        count_out);
    return check_launch("count_bytes_kernel");
}

extern "C" int smcb_cdf_scan_weights(int64_t n, const double* ll, const double* lp, double lp_const, const double* lq,
                                     const double* log_w_in, int64_t n_global, const smcb_path_t* path,
                                     const double* scalars, double* cdf, double* total_out, void* ws, void* stream) {
    SMCB_REQUIRE(n > 0 && ll && path && scalars && cdf && ws, "bad argument");
    SMCB_REQUIRE(((uintptr_t)cdf & 15) == 0, "cdf must be 16-byte aligned");
    char* base = (char*)ws;
    unsigned long long* state = (unsigned long long*)(base + kWsLarge);
    unsigned int* counter = (unsigned int*)(base + kWsTickets) + 1;
    const int64_t n_tiles = (n + kScanTile - 1) / kScanTile;
    LoadRw ld;
    ld.in = make_rw_in(ll, lp, lq, log_w_in, n_global, path, lp_const);
    ld.sc = scalars;
    scan_reset_kernel<<<grid_for(n_tiles, 256, 1, 4), 256, 0, as_stream(stream)>>>(state, n_tiles, counter);
    scan_lookback_kernel<LoadRw><<<(unsigned)n_tiles, kScanThreads, 0, as_stream(stream)>>>(ld, cdf, n, state, counter,
                                                                                              total_out);
    return check_launch("scan_lookback_kernel<LoadRw>");
}

extern "C" int smcb_resample_exponential_sums(int64_t first, int64_t count, int64_t n_global, uint64_t seed,
                                              uint64_t stream_id, double* t_out, double* total_out, void* ws,
                                              void* stream) {
    SMCB_REQUIRE(first >= 0 && count > 0 && n_global >= first + count && t_out && ws, "bad argument");
    SMCB_REQUIRE(((uintptr_t)t_out & 15) == 0, "t_out must be 16-byte aligned");
    char* base = (char*)ws;
    unsigned long long* state = (unsigned long long*)(base + kWsLarge);
    unsigned int* counter = (unsigned int*)(base + kWsTickets) + 1;
    const int64_t n_tiles = (count + kScanTile - 1) / kScanTile;
    LoadExp ld{first, n_global, seed, (uint32_t)stream_id};
    scan_reset_kernel<<<grid_for(n_tiles, 256, 1, 4), 256, 0, as_stream(stream)>>>(state, n_tiles, counter);
    scan_lookback_kernel<LoadExp><<<(unsigned)n_tiles, kScanThreads, 0, as_stream(stream)>>>(ld, t_out, count, state,
                                                                                               counter, total_out);
    return check_launch("scan_lookback_kernel<LoadExp>");
}

extern "C" int smcb_resample_fused(const smcb_resample_t* r, void* ws, void* stream) {
    SMCB_REQUIRE(r && ws, "bad argument");
    SMCB_REQUIRE(r->kind >= 0 && r->kind <= 2, "unknown resampler kind");
    SMCB_REQUIRE(r->world >= 1 && r->world <= SMCB_MAX_RANKS && r->rank >= 0 && r->rank < r->world, "bad ranks");
    SMCB_REQUIRE(r->n_local > 0 && r->n_global == r->n_local * r->world, "shards must be equal: n_global = world * n_local");
    SMCB_REQUIRE(r->cdf && r->src && r->n_cols > 0 && r->n_unique_out, "NULL argument");
    SMCB_REQUIRE(r->totals || r->world == 1, "totals are required on several GPUs");
    SMCB_REQUIRE(r->world == 1 || r->idx_out == nullptr, "idx_out is single-GPU only");
    RsDev a;
    memset(&a, 0, sizeof(a));
    a.kind = r->kind;
    a.world = r->world;
    a.rank = r->rank;
    a.n_cols = r->n_cols;
    a.n_local = r->n_local;
    a.n_global = r->n_global;
    a.cdf = r->cdf;
    a.totals = r->totals;
    a.src = r->src;
    a.src_stride = r->src_stride;
    a.dst_stride = r->dst_stride;
    a.idx_out = r->idx_out;
    a.n_unique_out = r->n_unique_out;
    a.seed = r->seed;
    a.stream_id = (uint32_t)r->stream_id;
    for (int q = 0; q < r->world; ++q) {
        SMCB_REQUIRE(r->dst_peer[q] != nullptr, "destination pointer is NULL");
        SMCB_REQUIRE(r->kind != SMCB_RESAMPLE_STANDARD || r->t_peer[q] != nullptr, "running-sum pointer is NULL");
        a.dst_peer[q] = r->dst_peer[q];
        a.t_peer[q] = r->t_peer[q];
    }
    char* base = (char*)ws;
    a.w_off = (double*)(base + kWsSmall + 1024);
    a.e_off = a.w_off + 16;
    a.krange = (int64_t*)(a.e_off + 16);
    a.uniq = (unsigned long long*)(base + kWsSmall + 64);
    a.ticket = (unsigned int*)(base + kWsTickets) + 2;
    const int64_t max_tiles = (r->n_global + kRsTile - 1) / kRsTile;
    a.tb = (int64_t*)(base + kWsLarge);
    a.tp = a.tb + (max_tiles + 2);
    cudaStream_t s = as_stream(stream);
    rs_plan_kernel<<<1, 32, 0, s>>>(a);
    rs_tile_bounds_kernel<<<(unsigned)((max_tiles + 1 + 255) / 256), 256, 0, s>>>(a, max_tiles);
    const int64_t cap = (int64_t)kNumSMs * 8;
    const int64_t local_tiles = (r->n_local + kRsTile - 1) / kRsTile;
    // an owner's slot range is ~n_local slots on average but may be all of n_global: persistent tiles
    int64_t grid = r->world == 1 ? local_tiles : 2 * local_tiles;
    if (grid > cap) grid = cap;
    if (grid < 1) grid = 1;
    rs_fused_kernel<<<(unsigned)grid, kRsThreads, 0, s>>>(a);
    return check_launch("rs_fused_kernel");
}
