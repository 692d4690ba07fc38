// moments.cu -- weighted mean / covariance of the particle cloud, Cholesky factor, layout.
//
// Reference: Particles.compute_covariance = np.cov(params.T, ddof=0, aweights=w)
// (smcpy/smc/particles.py:189-198): two passes,
//     mean = sum(w x) / sum(w);   C = sum(w (x-mean)(x-mean)^T) * (1 / sum(w)).
// np.linalg.cholesky of the result (smcpy/mcmc/vector_mcmc.py:225).
// Algorithmic bytes per particle: pass 1 reads d+1 doubles, pass 2 reads d+1 doubles.
#include <stdlib.h>

#include "common.cuh"

namespace smcb {

constexpr int kMomThreads = 256;

// ---- pass 1: sum w, sum w x_c for a chunk of DC columns ---------------------------------------
template <int DC>
__global__ void __launch_bounds__(kMomThreads) wsums1_kernel(const double* __restrict__ params, int64_t pstride,
                                                             const double* __restrict__ w, int64_t n, int d,
                                                             double* partials /* [grid.x][1+d] */) {
    __shared__ double sh[32];
    const int c0 = blockIdx.y * DC;
    double aw = 0.0, ax[DC];
#pragma unroll
    for (int j = 0; j < DC; ++j) ax[j] = 0.0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const double wi = w[i];
        aw += wi;
#pragma unroll
        for (int j = 0; j < DC; ++j)
            if (c0 + j < d) ax[j] += __dmul_rn(params[(int64_t)(c0 + j) * pstride + i], wi);
    }
    double* out = partials + (int64_t)blockIdx.x * (1 + d);
    const double tw = block_sum(aw, sh);
    if (threadIdx.x == 0 && blockIdx.y == 0) out[0] = tw;
#pragma unroll
    for (int j = 0; j < DC; ++j) {
        const double t = block_sum(ax[j], sh);
        if (threadIdx.x == 0 && c0 + j < d) out[1 + c0 + j] = t;
    }
}

// out[j] = sum over blocks of partials[b][j]: one warp per output value, lane l adds blocks
// l, l+32, ... in order, then a fixed shuffle tree (deterministic)
__global__ void __launch_bounds__(128) reduce_partials_kernel(const double* partials, int n_blocks, int n_vals,
                                                              double* out) {
    const int j = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (j >= n_vals) return;
    double s = 0.0;
    for (int b = lane; b < n_blocks; b += 32) s += partials[(int64_t)b * n_vals + j];
    s = warp_sum(s);
    if (lane == 0) out[j] = s;
}

__global__ void mean_from_sums_kernel(const double* sums, int d, double* mean) {
    const int j = threadIdx.x;
    if (j < d) mean[j] = sums[1 + j] / sums[0];
}

// ---- pass 2: sum w (x_i - m_i)(x_j - m_j), smem-tiled, 4x4 register blocks ----------------------
// Tile of P particles x d columns is staged (centred) in shared memory; thread (slice, blk)
// accumulates one upper-triangular 4x4 block over its slice of the tile.
constexpr int kTileP = 128;

struct Blk { unsigned char bi, bj; };

__global__ void __launch_bounds__(kMomThreads) wsums2_kernel(const double* __restrict__ params, int64_t pstride,
                                                             const double* __restrict__ w, int64_t n, int d,
                                                             const double* __restrict__ mean,
                                                             double* partials /* [grid][nbu*16] */) {
    extern __shared__ double smem[];
    const int nb = (d + 3) >> 2;                 // 4-wide column blocks
    const int dp = nb * 4;                       // padded column count
    const int nbu = nb * (nb + 1) / 2;           // upper-triangular blocks
    const int nslices = kMomThreads / nbu;
    constexpr int ld = kTileP + 1;               // odd leading dimension: conflict-free column reads
    double* xc = smem;                           // [dp][ld] centred columns
    double* ws = smem + dp * ld;                 // [kTileP] weights

    const int t = threadIdx.x;
    const int blk = t % nbu, slice = t / nbu;
    // decode blk -> (bi <= bj)
    int bi = 0, rem = blk;
    while (rem >= nb - bi) { rem -= nb - bi; ++bi; }
    const int bj = bi + rem;
    const bool active = slice < nslices;

    double acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = 0.0;

    const int64_t n_tiles = (n + kTileP - 1) / kTileP;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t base = tile * kTileP;
        __syncthreads();
        for (int e = t; e < dp * kTileP; e += kMomThreads) {
            const int c = e / kTileP, p = e - c * kTileP;
            double v = 0.0;
            if (c < d && base + p < n) v = params[(int64_t)c * pstride + base + p] - mean[c];
            xc[c * ld + p] = v;
        }
        for (int p = t; p < kTileP; p += kMomThreads) ws[p] = (base + p < n) ? w[base + p] : 0.0;
        __syncthreads();
        if (active) {
            for (int p = slice; p < kTileP; p += nslices) {
                const double wp = ws[p];
                double xi[4], xj[4];
#pragma unroll
                for (int a = 0; a < 4; ++a) {
                    xi[a] = __dmul_rn(xc[(bi * 4 + a) * ld + p], wp);
                    xj[a] = xc[(bj * 4 + a) * ld + p];
                }
#pragma unroll
                for (int a = 0; a < 4; ++a)
#pragma unroll
                    for (int b = 0; b < 4; ++b) acc[a][b] = fma(xi[a], xj[b], acc[a][b]);
            }
        }
    }
    // reduce the slices of this block in slice order (deterministic)
    __syncthreads();
    double* red = smem;                           // reuse: [kMomThreads][16]
    if (active) {
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) red[(slice * nbu + blk) * 16 + a * 4 + b] = acc[a][b];
    }
    __syncthreads();
    for (int e = t; e < nbu * 16; e += kMomThreads) {
        double s = 0.0;
        for (int sl = 0; sl < nslices; ++sl) s += red[sl * nbu * 16 + e];
        partials[(int64_t)blockIdx.x * (nbu * 16) + e] = s;
    }
}

// ---- pass 2 for d > 8: FP64 tensor cores (DMMA m8n8k4) + TMA-prefetched tiles ----------------------
// sum_p w_p (x_p - m)(x_p - m)^T is a SYRK: C (d x d) += XW (d x 4) * XC^T (4 x d) for every
// group of 4 particles.  One mma.sync.m8n8k4.f64 does 256 FMAs from two 8-byte fragment loads
// per lane, so shared-memory bandwidth is no longer the limit (a 4x4 register-blocked SYRK
// needs 8 loads per 16 FMAs and is smem-bound at ~1/3 of the FP64 rate, profiles/r01_cov.md).
// Per CTA: the raw tile (d columns + weights, 128 particles) is bulk-copied by TMA
// (cp.async.bulk + mbarrier); all threads centre and weight it into xc / xw; the raw stage is
// re-armed with the next tile at once; each of the 8 warps then takes 4 of the tile's 32
// k-steps over all upper-triangular 8x8 blocks, accumulating in registers across tiles.
constexpr int kWP = 128;            // particles per tile
constexpr int kWLd = 132;           // row length: == 4 (mod 16) doubles -> conflict-free fragment loads
constexpr int kDmmaWarps = 8;
constexpr int kWStages = 3;        // one-pass moments: TMA stages per CTA (3 x 34 KB x 2 CTAs per SM at d = 32)

__device__ __forceinline__ void dmma_m8n8k4(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

template <int NB8>
__global__ void __launch_bounds__(kDmmaWarps * 32, 2)
wsums2_dmma_kernel(const double* __restrict__ params, int64_t pstride, const double* __restrict__ w, int64_t n,
                   int d, const double* __restrict__ mean, double* partials /* [grid][NBU8*64] */) {
    constexpr int NBU8 = NB8 * (NB8 + 1) / 2;
    constexpr int DP = NB8 * 8;
    extern __shared__ __align__(16) double smem[];
    __shared__ __align__(8) uint64_t mbar;
    double* raw = smem;                          // [DP + 1][kWP]  (last row: weights)
    double* xc = raw + (DP + 1) * kWP;           // [DP][kWLd] centred
    double* xw = xc + DP * kWLd;                 // [DP][kWLd] centred * weight
    double* smean = xw + DP * kWLd;              // [DP]
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int frow = lane >> 2, fk = lane & 3;   // fragment row / k index of this lane

    double acc[NBU8][2];
#pragma unroll
    for (int q = 0; q < NBU8; ++q) acc[q][0] = acc[q][1] = 0.0;

    for (int c = tid; c < DP; c += blockDim.x) smean[c] = (c < d) ? mean[c] : 0.0;
    const int64_t n_tiles = (n + kWP - 1) / kWP;
    auto issue = [&](int64_t tile) {             // one elected thread
        const int64_t base = tile * kWP;
        const uint32_t cnt = (uint32_t)((n - base < kWP) ? (n - base) : kWP);
        const uint32_t bytes = cnt * 8u;
        fence_proxy_async();
        mbar_expect_tx(&mbar, bytes * (uint32_t)(d + 1));
        for (int c = 0; c < d; ++c) tma_load_1d(raw + c * kWP, params + (int64_t)c * pstride + base, bytes, &mbar);
        tma_load_1d(raw + DP * kWP, w + base, bytes, &mbar);
    };
    if (tid == 0) mbar_init(&mbar, 1);
    __syncthreads();
    if (tid == 0 && (int64_t)blockIdx.x < n_tiles) issue(blockIdx.x);
    uint32_t phase = 0;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t base = tile * kWP;
        const int cnt = (int)((n - base < kWP) ? (n - base) : kWP);
        mbar_wait(&mbar, phase);
        phase ^= 1u;
        for (int e = tid; e < DP * kWP; e += blockDim.x) {
            const int c = e / kWP, p = e - c * kWP;
            double v = 0.0, vw = 0.0;
            if (c < d && p < cnt) {
                v = raw[c * kWP + p] - smean[c];
                vw = __dmul_rn(v, raw[DP * kWP + p]);
            }
            xc[c * kWLd + p] = v;
            xw[c * kWLd + p] = vw;
        }
        __syncthreads();                          // raw drained, xc / xw complete
        if (tid == 0 && tile + gridDim.x < n_tiles) issue(tile + gridDim.x);
#pragma unroll
        for (int ks = 0; ks < kWP / 4 / kDmmaWarps; ++ks) {
            const int p0 = (wid + ks * kDmmaWarps) * 4 + fk;
            double fa[NB8], fb[NB8];
#pragma unroll
            for (int I = 0; I < NB8; ++I) {
                fa[I] = xw[(I * 8 + frow) * kWLd + p0];
                fb[I] = xc[(I * 8 + frow) * kWLd + p0];
            }
            int q = 0;
#pragma unroll
            for (int I = 0; I < NB8; ++I)
#pragma unroll
                for (int J = I; J < NB8; ++J) {
                    dmma_m8n8k4(acc[q][0], acc[q][1], fa[I], fb[J]);
                    ++q;
                }
        }
        __syncthreads();                          // xc / xw free for the next tile
    }
    // warps -> CTA partial in warp order (deterministic); reuse xc as [warp][NBU8*64]
    double* red = xc;
#pragma unroll
    for (int q = 0; q < NBU8; ++q) {
        red[(wid * NBU8 + q) * 64 + frow * 8 + fk * 2 + 0] = acc[q][0];
        red[(wid * NBU8 + q) * 64 + frow * 8 + fk * 2 + 1] = acc[q][1];
    }
    __syncthreads();
    for (int e = tid; e < NBU8 * 64; e += blockDim.x) {
        double sum = 0.0;
        for (int wq = 0; wq < kDmmaWarps; ++wq) sum += red[wq * NBU8 * 64 + e];
        partials[(int64_t)blockIdx.x * (NBU8 * 64) + e] = sum;
    }
}

// ---- one pass for d > 8: shifted first and second moments together ------------------------------------
// W = sum w, S1 = sum w (x - c), S2 = sum w (x - c)(x - c)^T about a shift c close to the mean (the
// previous SMC step's mean): mean = c + S1 / W, cov = S2 / W - (S1 / W)(S1 / W)^T.  The cancellation
// error is eps * |mean - c|^2 / var, i.e. nil when c is within a few standard deviations of the mean;
// the finalisation flags SMCB_FLAG_COV_RESHIFT otherwise and the caller falls back to the two passes.
// Halves the HBM traffic of the covariance (one read of the particles instead of two).
// Tiles go from TMA straight into rows padded to kWLd (conflict-free fragment loads); centring and
// weighting happen in registers at fragment-load time, so there is no staging pass; a ring of kWStages
// raw tiles per CTA keeps the loads ahead of the DMMAs (with two stages the refill, issued only when all
// eight warps have left a tile, did not cover the HBM latency: 45 % of the warp time at the tile barrier).
template <int NB8>
__global__ void __launch_bounds__(kDmmaWarps * 32, 2)
wmoments_dmma_kernel(const double* __restrict__ params, int64_t pstride, const double* __restrict__ w, int64_t n,
                     int d, const double* __restrict__ shift, double* partials /* [grid][NBU8*64 + DP + 1] */) {
    constexpr int NBU8 = NB8 * (NB8 + 1) / 2;
    constexpr int DP = NB8 * 8;
    constexpr int kStage = (DP + 1) * kWLd;
    constexpr int kOut = NBU8 * 64 + DP + 1;
    extern __shared__ __align__(16) double smem[];
    __shared__ __align__(8) uint64_t mbar[kWStages];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int frow = lane >> 2, fk = lane & 3;   // fragment row / k index of this lane

    double acc[NBU8][2];
#pragma unroll
    for (int q = 0; q < NBU8; ++q) acc[q][0] = acc[q][1] = 0.0;
    double s1[NB8], sh[NB8], sw = 0.0;
#pragma unroll
    for (int I = 0; I < NB8; ++I) {
        s1[I] = 0.0;
        sh[I] = (I * 8 + frow < d) ? shift[I * 8 + frow] : 0.0;
    }
    const int64_t n_tiles = (n + kWP - 1) / kWP;
    auto issue = [&](int stage, int64_t tile) {  // one elected thread
        double* raw = smem + stage * kStage;
        const int64_t base = tile * kWP;
        const uint32_t cnt = (uint32_t)((n - base < kWP) ? (n - base) : kWP);
        const uint32_t bytes = cnt * 8u;
        fence_proxy_async();
        mbar_expect_tx(&mbar[stage], bytes * (uint32_t)(d + 1));
#pragma unroll 1
        for (int c = 0; c < d; ++c) tma_load_1d(raw + c * kWLd, params + (int64_t)c * pstride + base, bytes, &mbar[stage]);
        tma_load_1d(raw + DP * kWLd, w + base, bytes, &mbar[stage]);
    };
    if (tid == 0) {
#pragma unroll
        for (int st = 0; st < kWStages; ++st) mbar_init(&mbar[st], 1);
    }
    __syncthreads();
    if (tid == 0) {
#pragma unroll
        for (int st = 0; st < kWStages; ++st)
            if ((int64_t)blockIdx.x + (int64_t)st * gridDim.x < n_tiles) issue(st, (int64_t)blockIdx.x + (int64_t)st * gridDim.x);
    }
    int stage = 0;
    uint32_t phase = 0;                          // parity of the current trip round the stage ring
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const double* raw = smem + stage * kStage;
        const int64_t base = tile * kWP;
        const int cnt = (int)((n - base < kWP) ? (n - base) : kWP);
        mbar_wait(&mbar[stage], phase);
#pragma unroll
        for (int ks = 0; ks < kWP / 4 / kDmmaWarps; ++ks) {
            const int p0 = (wid + ks * kDmmaWarps) * 4 + fk;
            const bool valid = p0 < cnt;
            const double wv = valid ? raw[DP * kWLd + p0] : 0.0;
            double fa[NB8], fb[NB8];
#pragma unroll
            for (int I = 0; I < NB8; ++I) {
                const bool on = valid && (I * 8 + frow < d);
                fb[I] = on ? raw[(I * 8 + frow) * kWLd + p0] - sh[I] : 0.0;
                fa[I] = __dmul_rn(fb[I], wv);
                s1[I] += fa[I];
            }
            if (frow == 0) sw += wv;
            int q = 0;
#pragma unroll
            for (int I = 0; I < NB8; ++I)
#pragma unroll
                for (int J = I; J < NB8; ++J) {
                    dmma_m8n8k4(acc[q][0], acc[q][1], fa[I], fb[J]);
                    ++q;
                }
        }
        __syncthreads();                          // every warp is done with this stage
        if (tid == 0 && tile + kWStages * (int64_t)gridDim.x < n_tiles) issue(stage, tile + kWStages * (int64_t)gridDim.x);
        if (++stage == kWStages) {
            stage = 0;
            phase ^= 1u;
        }
    }
    __syncthreads();
    // lanes -> warp: the four k-lanes of a fragment row hold different particles
#pragma unroll
    for (int I = 0; I < NB8; ++I) {
        s1[I] += __shfl_xor_sync(0xffffffffu, s1[I], 1);
        s1[I] += __shfl_xor_sync(0xffffffffu, s1[I], 2);
    }
    sw += __shfl_xor_sync(0xffffffffu, sw, 1);
    sw += __shfl_xor_sync(0xffffffffu, sw, 2);
    // warps -> CTA partial in warp order (deterministic); the stages are free now
    double* red = smem;                           // [warp][kOut]
#pragma unroll
    for (int q = 0; q < NBU8; ++q) {
        red[wid * kOut + q * 64 + frow * 8 + fk * 2 + 0] = acc[q][0];
        red[wid * kOut + q * 64 + frow * 8 + fk * 2 + 1] = acc[q][1];
    }
    if (fk == 0) {
#pragma unroll
        for (int I = 0; I < NB8; ++I) red[wid * kOut + NBU8 * 64 + I * 8 + frow] = s1[I];
    }
    if (lane == 0) red[wid * kOut + NBU8 * 64 + DP] = sw;
    __syncthreads();
    for (int e = tid; e < kOut; e += blockDim.x) {
        double sum = 0.0;
        for (int wq = 0; wq < kDmmaWarps; ++wq) sum += red[wq * kOut + e];
        partials[(int64_t)blockIdx.x * kOut + e] = sum;
    }
}

// packed (S2 blocks, S1[DP], W) -> sums layout [W, S1[d], S2[d x d]]
__global__ void moments_unpack8_kernel(const double* packed, int d, int nb8, double* sums) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    const int nbu8 = nb8 * (nb8 + 1) / 2;
    if (e == 0) sums[0] = packed[nbu8 * 64 + nb8 * 8];
    if (e < d) sums[1 + e] = packed[nbu8 * 64 + e];
    if (e >= d * d) return;
    int i = e / d, j = e - i * d;
    if (i > j) { const int tmp = i; i = j; j = tmp; }
    const int bi = i >> 3, bj = j >> 3;
    const int blk = bi * nb8 - bi * (bi - 1) / 2 + (bj - bi);
    sums[1 + d + e] = packed[blk * 64 + (i & 7) * 8 + (j & 7)];
}

// mean = c + S1 / W; cov = S2 / W - (S1 / W)(S1 / W)^T; c <- mean for the next call.
// One block of >= d*d threads... (d <= 32: 1024 threads)
__global__ void cov_from_shifted_sums_kernel(const double* sums, double* shift, int d, double* mean, double* cov,
                                             int* flags) {
    __shared__ double m1[SMCB_MAX_PARAMS];
    const int e = threadIdx.x;
    const double fact = 1.0 / sums[0];
    if (e < d) m1[e] = __dmul_rn(sums[1 + e], fact);
    __syncthreads();
    if (e < d * d) {
        const int i = e / d, j = e - i * d;
        const double c = __dmul_rn(sums[1 + d + e], fact) - m1[i] * m1[j];
        cov[e] = c;
        if (i == j && !(m1[i] * m1[i] <= 1e4 * c)) atomicOr(flags, SMCB_FLAG_COV_RESHIFT);
    }
    if (e < d) {
        const double mu = shift[e] + m1[e];
        mean[e] = mu;
        shift[e] = mu;
    }
}

// expand the 8x8-block-packed upper triangle to a full symmetric d x d matrix
__global__ void sums2_unpack8_kernel(const double* packed, int d, int nb8, double* full) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= d * d) return;
    int i = e / d, j = e - i * d;
    if (i > j) { const int tmp = i; i = j; j = tmp; }
    const int bi = i >> 3, bj = j >> 3;
    const int blk = bi * nb8 - bi * (bi - 1) / 2 + (bj - bi);
    full[e] = packed[blk * 64 + (i & 7) * 8 + (j & 7)];
}

// expand the block-packed upper triangle to a full symmetric d x d matrix
__global__ void sums2_unpack_kernel(const double* packed, int d, double* full) {
    const int nb = (d + 3) >> 2;
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= d * d) return;
    int i = e / d, j = e - i * d;
    if (i > j) { const int tmp = i; i = j; j = tmp; }
    const int bi = i >> 2, bj = j >> 2;
    // index of block (bi, bj) in the upper-triangular enumeration
    const int blk = bi * nb - bi * (bi - 1) / 2 + (bj - bi);
    full[e] = packed[blk * 16 + (i & 3) * 4 + (j & 3)];
}

__global__ void cov_from_sums_kernel(const double* wsum, const double* sums2, int d, double* cov) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= d * d) return;
    const double fact = 1.0 / wsum[0];            // np.cov: c *= true_divide(1, fact)
    cov[e] = __dmul_rn(sums2[e], fact);
}

// ---- Cholesky (lower), one warp; LAPACK dpotrf failure rule: pivot <= 0 or NaN --------------------
__global__ void cholesky_kernel(const double* cov, int d, double* L, int* flags) {
    __shared__ double a[SMCB_MAX_PARAMS * (SMCB_MAX_PARAMS + 1)];
    const int lane = threadIdx.x;
    const int ld = d + 1;
    for (int e = lane; e < d * d; e += 32) a[(e / d) * ld + (e % d)] = cov[e];
    __syncwarp();
    bool fail = false;
    for (int j = 0; j < d; ++j) {
        // row `lane` (>= j): s = a[i][j] - sum_k<j L[i][k] L[j][k]
        double s = 0.0;
        const int i = lane;
        if (i >= j && i < d) {
            s = a[i * ld + j];
            for (int k = 0; k < j; ++k) s -= a[i * ld + k] * a[j * ld + k];
        }
        const double pivot = __shfl_sync(0xffffffffu, s, j);
        if (!(pivot > 0.0)) { fail = true; break; }
        const double r = sqrt(pivot);
        if (i >= j && i < d) a[i * ld + j] = (i == j) ? r : s / r;
        __syncwarp();
    }
    for (int e = lane; e < d * d; e += 32) {
        const int i = e / d, j = e % d;
        L[e] = (j <= i && !fail) ? a[i * ld + j] : 0.0;
    }
    if (fail && lane == 0) atomicOr(flags, SMCB_FLAG_CHOL_FAIL);
}

// ---- layout -----------------------------------------------------------------------------------------
// AoS (n, d) row-major <-> SoA (d, stride); staged through shared memory so both sides coalesce
template <bool kToSoa>
__global__ void __launch_bounds__(256) transpose_kernel(const double* __restrict__ src, double* __restrict__ dst,
                                                        int64_t n, int d, int64_t stride, int pb) {
    extern __shared__ double tile[];
    const int ldp = d | 1;                         // odd row length
    const int64_t n_chunks = (n + pb - 1) / pb;
    for (int64_t ch = blockIdx.x; ch < n_chunks; ch += gridDim.x) {
        const int64_t base = ch * pb;
        const int m = (int)((n - base < pb) ? (n - base) : pb);
        __syncthreads();
        if (kToSoa) {
            for (int e = threadIdx.x; e < m * d; e += blockDim.x) tile[(e / d) * ldp + (e % d)] = src[base * d + e];
            __syncthreads();
            for (int e = threadIdx.x; e < m * d; e += blockDim.x) {
                const int c = e / m, p = e - c * m;
                dst[(int64_t)c * stride + base + p] = tile[p * ldp + c];
            }
        } else {
            for (int e = threadIdx.x; e < m * d; e += blockDim.x) {
                const int c = e / m, p = e - c * m;
                tile[p * ldp + c] = src[(int64_t)c * stride + base + p];
            }
            __syncthreads();
            for (int e = threadIdx.x; e < m * d; e += blockDim.x) dst[base * d + e] = tile[(e / d) * ldp + (e % d)];
        }
    }
}

__global__ void fill_kernel(double* dst, int64_t n, double v) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = v;
}

static int sums2_grid(int64_t n) {
    const int64_t tiles = (n + kTileP - 1) / kTileP;
    const int64_t cap = (int64_t)kNumSMs * 4;
    return (int)(tiles < cap ? tiles : cap);
}
static size_t sums2_smem(int d) {
    const int dp = ((d + 3) / 4) * 4;
    size_t a = (size_t)(dp * (kTileP + 1) + kTileP) * sizeof(double);
    size_t b = (size_t)kMomThreads * 16 * sizeof(double);
    return a > b ? a : b;
}

}  // namespace smcb

using namespace smcb;

extern "C" int smcb_weighted_sums1(const double* params, int64_t stride, const double* w, int64_t n, int32_t d,
                                   double* sums_out, void* ws, void* stream) {
    SMCB_REQUIRE(params && w && sums_out && ws && n > 0 && d > 0 && d <= 64, "bad argument");
    double* partials = (double*)((char*)ws + kWsPartials);
    int grid = grid_for(n, kMomThreads, 4, 2);
    const int maxb = (int)(kWsPartialsBytes / sizeof(double) / (1 + d));
    if (grid > maxb) grid = maxb;
    cudaStream_t s = as_stream(stream);
    if (d <= 1) wsums1_kernel<1><<<dim3(grid, 1), kMomThreads, 0, s>>>(params, stride, w, n, d, partials);
    else if (d <= 2) wsums1_kernel<2><<<dim3(grid, 1), kMomThreads, 0, s>>>(params, stride, w, n, d, partials);
    else if (d <= 4) wsums1_kernel<4><<<dim3(grid, 1), kMomThreads, 0, s>>>(params, stride, w, n, d, partials);
    else wsums1_kernel<8><<<dim3(grid, (d + 7) / 8), kMomThreads, 0, s>>>(params, stride, w, n, d, partials);
    reduce_partials_kernel<<<(1 + d + 3) / 4, 128, 0, s>>>(partials, grid, 1 + d, sums_out);
    return check_launch("wsums1_kernel");
}

extern "C" int smcb_mean_from_sums(const double* sums, int32_t d, double* mean_out, void* stream) {
    SMCB_REQUIRE(sums && mean_out && d > 0 && d <= 64, "bad argument");
    mean_from_sums_kernel<<<1, 64, 0, as_stream(stream)>>>(sums, d, mean_out);
    return check_launch("mean_from_sums_kernel");
}

extern "C" int smcb_weighted_sums2(const double* params, int64_t stride, const double* w, int64_t n, int32_t d,
                                   const double* mean, double* sums_out, void* ws, void* stream) {
    SMCB_REQUIRE(params && w && mean && sums_out && ws && n > 0 && d > 0 && d <= SMCB_MAX_PARAMS, "bad argument");
    static bool attr_set = false;
    const size_t smem = sums2_smem(d);
    if (!attr_set) {
        cudaFuncSetAttribute(wsums2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
        attr_set = true;
    }
    const int nb = (d + 3) / 4, nbu = nb * (nb + 1) / 2;
    double* partials = (double*)((char*)ws + kWsPartials);
    cudaStream_t s = as_stream(stream);
    int grid;
    const bool no_dmma = options().no_dmma != 0;                        // tuning / test aid
    const bool aligned = ((((uintptr_t)params | (uintptr_t)w) & 15) == 0) && (stride % 2 == 0) && (n % 2 == 0);
    if (!no_dmma && d > 8 && aligned && n >= (int64_t)kNumSMs * kWP) {
        const int nb8 = (d + 7) / 8, nbu8 = nb8 * (nb8 + 1) / 2, dp = nb8 * 8;
        const size_t smem_w = (size_t)((dp + 1) * kWP + 2 * dp * kWLd + dp) * sizeof(double);
        grid = kNumSMs * 2;
        if (nb8 == 2) {
            static bool at = false;
            if (!at) { cudaFuncSetAttribute(wsums2_dmma_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024); at = true; }
            wsums2_dmma_kernel<2><<<grid, kDmmaWarps * 32, smem_w, s>>>(params, stride, w, n, d, mean, partials);
        } else if (nb8 == 3) {
            static bool at = false;
            if (!at) { cudaFuncSetAttribute(wsums2_dmma_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024); at = true; }
            wsums2_dmma_kernel<3><<<grid, kDmmaWarps * 32, smem_w, s>>>(params, stride, w, n, d, mean, partials);
        } else {
            static bool at = false;
            if (!at) { cudaFuncSetAttribute(wsums2_dmma_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024); at = true; }
            wsums2_dmma_kernel<4><<<grid, kDmmaWarps * 32, smem_w, s>>>(params, stride, w, n, d, mean, partials);
        }
        double* packed8 = partials + (int64_t)grid * nbu8 * 64;
        reduce_partials_kernel<<<(nbu8 * 64 + 3) / 4, 128, 0, s>>>(partials, grid, nbu8 * 64, packed8);
        sums2_unpack8_kernel<<<(d * d + 63) / 64, 64, 0, s>>>(packed8, d, nb8, sums_out);
        return check_launch("wsums2_dmma_kernel");
    }
    grid = sums2_grid(n);
    {
        const int maxb = (int)(kWsPartialsBytes / sizeof(double) / (nbu * 16 + 0)) - 1;
        if (grid > maxb) grid = maxb;
        wsums2_kernel<<<grid, kMomThreads, smem, s>>>(params, stride, w, n, d, mean, partials);
    }
    // packed block sums live after the per-block partials
    double* packed = partials + (int64_t)grid * nbu * 16;
    reduce_partials_kernel<<<(nbu * 16 + 3) / 4, 128, 0, s>>>(partials, grid, nbu * 16, packed);
    sums2_unpack_kernel<<<(d * d + 63) / 64, 64, 0, s>>>(packed, d, sums_out);
    return check_launch("wsums2_kernel");
}

extern "C" int smcb_weighted_moments_shifted(const double* params, int64_t stride, const double* w, int64_t n,
                                             int32_t d, const double* shift, double* sums_out, void* ws,
                                             void* stream) {
    SMCB_REQUIRE(params && w && shift && sums_out && ws && n > 0 && d > 0 && d <= SMCB_MAX_PARAMS, "bad argument");
    const bool aligned = ((((uintptr_t)params | (uintptr_t)w) & 15) == 0) && (stride % 2 == 0) && (n % 2 == 0);
    if (!(d > 8 && aligned && n >= (int64_t)kNumSMs * kWP)) {
        set_error("smcb_weighted_moments_shifted: needs d > 8, 16-byte aligned even-length columns and n >= %d",
                  kNumSMs * kWP);
        return SMCB_ERR_UNSUPPORTED;
    }
    const int nb8 = (d + 7) / 8, nbu8 = nb8 * (nb8 + 1) / 2, dp = nb8 * 8;
    const int k_out = nbu8 * 64 + dp + 1;
    size_t smem_w = (size_t)kWStages * (dp + 1) * kWLd * sizeof(double);
    const size_t red = (size_t)kDmmaWarps * k_out * sizeof(double);
    if (red > smem_w) smem_w = red;
    const int grid = kNumSMs * 2;
    double* partials = (double*)((char*)ws + kWsPartials);
    cudaStream_t s = as_stream(stream);
    if (nb8 == 2) {
        static bool at = false;
        if (!at) { cudaFuncSetAttribute(wmoments_dmma_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024); at = true; }
        wmoments_dmma_kernel<2><<<grid, kDmmaWarps * 32, smem_w, s>>>(params, stride, w, n, d, shift, partials);
    } else if (nb8 == 3) {
        static bool at = false;
        if (!at) { cudaFuncSetAttribute(wmoments_dmma_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024); at = true; }
        wmoments_dmma_kernel<3><<<grid, kDmmaWarps * 32, smem_w, s>>>(params, stride, w, n, d, shift, partials);
    } else {
        static bool at = false;
        if (!at) { cudaFuncSetAttribute(wmoments_dmma_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024); at = true; }
        wmoments_dmma_kernel<4><<<grid, kDmmaWarps * 32, smem_w, s>>>(params, stride, w, n, d, shift, partials);
    }
    double* packed = partials + (int64_t)grid * k_out;
    reduce_partials_kernel<<<(k_out + 3) / 4, 128, 0, s>>>(partials, grid, k_out, packed);
    moments_unpack8_kernel<<<(d * d + 63) / 64, 64, 0, s>>>(packed, d, nb8, sums_out);
    return check_launch("wmoments_dmma_kernel");
}

extern "C" int smcb_cov_from_shifted_sums(const double* sums, double* shift, int32_t d, double* mean_out,
                                          double* cov_out, int32_t* flags, void* stream) {
    SMCB_REQUIRE(sums && shift && mean_out && cov_out && flags && d > 0 && d <= SMCB_MAX_PARAMS, "bad argument");
    cov_from_shifted_sums_kernel<<<1, 1024, 0, as_stream(stream)>>>(sums, shift, d, mean_out, cov_out, flags);
    return check_launch("cov_from_shifted_sums_kernel");
}

extern "C" int smcb_cov_from_sums(const double* wsum, const double* sums2, int32_t d, double* cov_out,
                                  void* stream) {
    SMCB_REQUIRE(wsum && sums2 && cov_out && d > 0 && d <= SMCB_MAX_PARAMS, "bad argument");
    cov_from_sums_kernel<<<(d * d + 63) / 64, 64, 0, as_stream(stream)>>>(wsum, sums2, d, cov_out);
    return check_launch("cov_from_sums_kernel");
}

extern "C" int smcb_weighted_cov(const double* params, int64_t stride, const double* w, int64_t n, int32_t d,
                                 double* mean_out, double* cov_out, double* sums_out, void* ws, void* stream) {
    SMCB_REQUIRE(sums_out != nullptr, "sums_out required (1 + d + d*d doubles)");
    int rc = smcb_weighted_sums1(params, stride, w, n, d, sums_out, ws, stream);
    if (rc) return rc;
    rc = smcb_mean_from_sums(sums_out, d, mean_out, stream);
    if (rc) return rc;
    rc = smcb_weighted_sums2(params, stride, w, n, d, mean_out, sums_out + 1 + d, ws, stream);
    if (rc) return rc;
    return smcb_cov_from_sums(sums_out, sums_out + 1 + d, d, cov_out, stream);
}

// the same two passes for a particle set sharded over several GPUs, enqueued by ONE call: the first and second
// moment sums are each all-reduced over the ranks by the peer-memory exchange kernel (rank-order add chain:
// bit-identical on every rank) on generations x->gen and x->gen + 1.  Six separate library calls with two Python
// exchange wrappers between them cost ~50 us of host enqueue time per SMC step around ~60 us of kernels.
extern "C" int smcb_weighted_cov_xchg(const double* params, int64_t stride, const double* w, int64_t n, int32_t d,
                                      double* mean_out, double* cov_out, double* sums_out, void* ws,
                                      const smcb_xchg_t* x, double* xchg_scratch, int32_t* status, void* stream) {
    SMCB_REQUIRE(sums_out != nullptr && x != nullptr && xchg_scratch != nullptr, "sums_out, x and xchg_scratch required");
    SMCB_REQUIRE(d > 0 && d <= SMCB_MAX_PARAMS && d * d <= smcb_xchg_max_vals(), "record too large for the peer exchange");
    // local sums go to the scratch record, the exchange kernel leaves the rank-order totals in sums_out
    int rc = smcb_weighted_sums1(params, stride, w, n, d, xchg_scratch, ws, stream);
    if (rc) return rc;
    smcb_xchg_t xv = *x;
    rc = smcb_xchg_allgather(&xv, xchg_scratch, 1 + d, sums_out, 1, status, stream);
    if (rc) return rc;
    rc = smcb_mean_from_sums(sums_out, d, mean_out, stream);
    if (rc) return rc;
    double* s2 = sums_out + 1 + d;
    rc = smcb_weighted_sums2(params, stride, w, n, d, mean_out, xchg_scratch, ws, stream);
    if (rc) return rc;
    xv.gen = x->gen + 1;
    rc = smcb_xchg_allgather(&xv, xchg_scratch, d * d, s2, 1, status, stream);
    if (rc) return rc;
    return smcb_cov_from_sums(sums_out, s2, d, cov_out, stream);
}

extern "C" int smcb_cholesky(const double* cov, int32_t d, double* chol_out, int32_t* flags, void* stream) {
    SMCB_REQUIRE(cov && chol_out && flags && d > 0 && d <= SMCB_MAX_PARAMS, "bad argument");
    cholesky_kernel<<<1, 32, 0, as_stream(stream)>>>(cov, d, chol_out, flags);
    return check_launch("cholesky_kernel");
}

static int transpose_launch(bool to_soa, const double* src, double* dst, int64_t n, int d, int64_t stride,
                            void* stream) {
    int pb = 4096 / (d | 1);
    if (pb > 256) pb = 256;
    if (pb < 1) pb = 1;
    const size_t smem = (size_t)pb * (d | 1) * sizeof(double);
    const int64_t chunks = (n + pb - 1) / pb;
    const int grid = (int)(chunks < (int64_t)kNumSMs * 8 ? chunks : (int64_t)kNumSMs * 8);
    if (to_soa) transpose_kernel<true><<<grid, 256, smem, as_stream(stream)>>>(src, dst, n, d, stride, pb);
    else transpose_kernel<false><<<grid, 256, smem, as_stream(stream)>>>(src, dst, n, d, stride, pb);
    return check_launch("transpose_kernel");
}

extern "C" int smcb_aos_to_soa(const double* aos, double* soa, int64_t n, int32_t d, int64_t stride,
                               void* stream) {
    SMCB_REQUIRE(aos && soa && n > 0 && d > 0 && d <= 4095 && stride >= n, "bad argument");
    return transpose_launch(true, aos, soa, n, d, stride, stream);
}

extern "C" int smcb_soa_to_aos(const double* soa, double* aos, int64_t n, int32_t d, int64_t stride,
                               void* stream) {
    SMCB_REQUIRE(aos && soa && n > 0 && d > 0 && d <= 4095 && stride >= n, "bad argument");
    return transpose_launch(false, soa, aos, n, d, stride, stream);
}

extern "C" int smcb_fill_f64(double* dst, int64_t n, double value, void* stream) {
    SMCB_REQUIRE(dst && n > 0, "bad argument");
    fill_kernel<<<grid_for(n, 256, 4, 8), 256, 0, as_stream(stream)>>>(dst, n, value);
    return check_launch("fill_kernel");
}
