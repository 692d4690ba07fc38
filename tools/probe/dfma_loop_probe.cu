// Microbenchmark: what limits the linear + Normal inner loop (per data point and particle r = fma(a, u, v);
// acc = fma(r, r, acc)) on B200?  Variants, all 128-thread blocks x 6 per SM, 4 particles per thread:
//   0  constant-operand chains a = fma(a, m, c)                    (the bench's FP64 peak probe)
//   1  the loop's DFMA pattern, (u, v) from registers rotated by adds (no memory)
//   2  the loop's DFMA pattern, (u, v) by one broadcast LDS.128 per point (what the kernel does)
//   3  like 2 with 8 particles per thread
//   4, 5  (u, v) read from the constant bank (no LDS), 4 / 8 particles per thread
// Measured on B200 (1965 MHz): 0: 34.9 TFLOP/s; 2: 31.7; 3: 33.7; 4: 22.9 (two indexed LDC.64 per point); 5: 28.4.
// Variant 1 is folded by the compiler (period-2 operands) and is not a rate.  The kernel's broadcast LDS.128 form is
// the best of the memory-fed forms: 91 % of the constant-operand chain rate.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/probe/dfma_loop_probe tools/probe/dfma_loop_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

__constant__ double2 cdata[1024];

template <int MODE, int PPT>
__global__ void __launch_bounds__(128, (MODE == 3 || MODE == 5 ? 4 : 6)) probe(double* out, const double2* data, int m, int reps) {
    __shared__ double2 s[1024];
    for (int j = threadIdx.x; j < m; j += blockDim.x) s[j] = data[j];
    __syncthreads();
    double a[PPT], acc0[PPT], acc1[PPT];
#pragma unroll
    for (int p = 0; p < PPT; ++p) { a[p] = 1.0 + 1e-3 * (threadIdx.x + p); acc0[p] = 0.0; acc1[p] = 0.0; }
    for (int r = 0; r < reps; ++r) {
        if (MODE == 0) {
#pragma unroll 4
            for (int j = 0; j < m; j += 2) {
#pragma unroll
                for (int p = 0; p < PPT; ++p) {
                    acc0[p] = fma(acc0[p], 0.999999, 1e-7); acc1[p] = fma(acc1[p], 0.999999, 1e-7);
                    a[p] = fma(a[p], 0.999999, 1e-7); acc0[p] = fma(acc0[p], 0.999998, 1e-7);
                }
            }
        } else if (MODE == 1) {
            double u0 = 1e-3, v0 = 2e-3, u1 = 3e-3, v1 = 4e-3;
#pragma unroll 8
            for (int j = 0; j < m; j += 2) {
#pragma unroll
                for (int p = 0; p < PPT; ++p) {
                    const double d0 = fma(a[p], u0, v0), d1 = fma(a[p], u1, v1);
                    acc0[p] = fma(d0, d0, acc0[p]);
                    acc1[p] = fma(d1, d1, acc1[p]);
                }
                u0 = -u0; v1 = -v1;            // cheap non-FP64 change so the loop is not hoisted
            }
        } else if (MODE == 4 || MODE == 5) {
#pragma unroll 8
            for (int j = 0; j + 2 <= m; j += 2) {
                const double2 p0 = cdata[j], p1 = cdata[j + 1];
#pragma unroll
                for (int p = 0; p < PPT; ++p) {
                    const double d0 = fma(a[p], p0.x, p0.y), d1 = fma(a[p], p1.x, p1.y);
                    acc0[p] = fma(d0, d0, acc0[p]);
                    acc1[p] = fma(d1, d1, acc1[p]);
                }
            }
        } else {
#pragma unroll 8
            for (int j = 0; j + 2 <= m; j += 2) {
                const double2 p0 = s[j], p1 = s[j + 1];
#pragma unroll
                for (int p = 0; p < PPT; ++p) {
                    const double d0 = fma(a[p], p0.x, p0.y), d1 = fma(a[p], p1.x, p1.y);
                    acc0[p] = fma(d0, d0, acc0[p]);
                    acc1[p] = fma(d1, d1, acc1[p]);
                }
            }
        }
    }
    double t = 0;
#pragma unroll
    for (int p = 0; p < PPT; ++p) t += acc0[p] + acc1[p] + a[p];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = t;
}

template <int MODE, int PPT>
void run(const char* name, double* out, const double2* data, int m, int reps, int blocks) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    probe<MODE, PPT><<<blocks, 128>>>(out, data, m, reps);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    probe<MODE, PPT><<<blocks, 128>>>(out, data, m, reps);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double fmas = (double)blocks * 128 * PPT * (double)m * 2.0 * reps;     // 2 DFMA per point and particle (4 per 2 points in mode 0)
    printf("%-44s %8.3f ms  %6.2f TFLOP/s (%s)\n", name, ms, 2.0 * fmas / ms * 1e-9, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    const int m = 1000, reps = 40;
    double* out; double2* data;
    cudaMalloc(&out, sizeof(double) * 148 * 6 * 128);
    cudaMalloc(&data, sizeof(double2) * 1024);
    cudaMemset(data, 0, sizeof(double2) * 1024);
    run<0, 4>("0 constant-operand chains", out, data, m, reps, 148 * 6);
    run<1, 4>("1 loop pattern, register operands, PPT 4", out, data, m, reps, 148 * 6);
    run<2, 4>("2 loop pattern, LDS.128 per point, PPT 4", out, data, m, reps, 148 * 6);
    run<3, 8>("3 loop pattern, LDS.128 per point, PPT 8", out, data, m, reps, 148 * 4);
    run<4, 4>("4 loop pattern, constant bank operands, PPT 4", out, data, m, reps, 148 * 6);
    run<5, 8>("5 loop pattern, constant bank operands, PPT 8", out, data, m, reps, 148 * 4);
    return 0;
}
