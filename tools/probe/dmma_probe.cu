// Microbenchmark: does the FP64 tensor-core path (mma.sync.m8n8k4.f64, DMMA) run concurrently with
// the vector FP64 pipe (DFMA) on B200?  Prints time for DFMA-only, DMMA-only and the interleaved mix.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/probe/dmma_probe tools/probe/dmma_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

template <int NF, int NM>
__global__ void __launch_bounds__(256) probe(double* out, int iters, double a, double b) {
    double f[NF > 0 ? NF : 1];
    double m0[NM > 0 ? NM : 1], m1[NM > 0 ? NM : 1];
#pragma unroll
    for (int i = 0; i < NF; ++i) f[i] = threadIdx.x * 1e-3 + i;
#pragma unroll
    for (int i = 0; i < NM; ++i) { m0[i] = threadIdx.x * 1e-3 + i; m1[i] = i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
            for (int i = 0; i < NF; ++i) f[i] = fma(f[i], a, b);
#pragma unroll
            for (int i = 0; i < NM; ++i) dmma(m0[i], m1[i], a, b);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NF; ++i) s += f[i];
#pragma unroll
    for (int i = 0; i < NM; ++i) s += m0[i] + m1[i];
    if (s == 12345.678) out[0] = s;
}

template <int NF, int NM>
float run(double* out, int iters, int blocks) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    probe<NF, NM><<<blocks, 256>>>(out, 10, 0.999, 1e-9);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    probe<NF, NM><<<blocks, 256>>>(out, iters, 0.999, 1e-9);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    return ms;
}

int main() {
    double* out; cudaMalloc(&out, 8);
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int blocks = p.multiProcessorCount * 4, iters = 4096;
    const double warps = (double)blocks * 8, reps = (double)iters * 4;
    float t;
    t = run<8, 0>(out, iters, blocks);
    printf("DFMA x8        : %.3f ms  %.2f TFLOP/s (fma=2)\n", t, warps * reps * 8 * 32 * 2 / t / 1e9);
    t = run<0, 4>(out, iters, blocks);
    printf("DMMA x4        : %.3f ms  %.2f TFLOP/s (512/instr)  %.3f DMMA/clk/SM @1.965GHz\n", t, warps * reps * 4 * 512 / t / 1e9,
           warps * reps * 4 / (t * 1e-3) / p.multiProcessorCount / 1.965e9);
    t = run<0, 8>(out, iters, blocks);
    printf("DMMA x8        : %.3f ms  %.2f TFLOP/s\n", t, warps * reps * 8 * 512 / t / 1e9);
    t = run<8, 4>(out, iters, blocks);
    printf("DFMA x8+DMMA x4: %.3f ms  (sum of parts if serialized)\n", t);
    t = run<8, 1>(out, iters, blocks);
    printf("DFMA x8+DMMA x1: %.3f ms\n", t);
    t = run<8, 2>(out, iters, blocks);
    printf("DFMA x8+DMMA x2: %.3f ms\n", t);
    t = run<4, 4>(out, iters, blocks);
    printf("DFMA x4+DMMA x4: %.3f ms\n", t);
    return 0;
}
