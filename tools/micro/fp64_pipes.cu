// Do DMMA (FP64 tensor) and DFMA (FP64 vector) share issue bandwidth on sm_100a?  Times DMMA only, DFMA only, and both in one warp.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_pipes fp64_pipes.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

template <int NM, int NF>
__global__ void __launch_bounds__(512, 1) mix(double* out, int iters, double seed) {
    double m[8][2], f[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { m[i][0] = seed * i; m[i][1] = seed; f[i] = seed + i; }
    const double a = seed * 0.5, b = seed * 0.25;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
            for (int i = 0; i < NM; ++i) dmma884(m[i][0], m[i][1], a, b);
#pragma unroll
            for (int i = 0; i < NF; ++i) asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(f[i]) : "d"(a), "d"(b));
        }
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += m[i][0] + m[i][1] + f[i];
    if (s == 123.456) out[0] = s;
}

template <int NM, int NF>
void run(const char* name, double* d) {
    const int iters = 20000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    mix<NM, NF><<<148, 512>>>(d, 100, 1e-300);
    cudaEventRecord(e0);
    mix<NM, NF><<<148, 512>>>(d, iters, 1e-300);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double warps = 148.0 * 16, n = (double)iters * 4;
    const double tf_m = warps * n * NM * 512 / (ms * 1e-3) / 1e12, tf_f = warps * n * NF * 64 / (ms * 1e-3) / 1e12;
    printf("%-28s %8.3f ms  DMMA %6.2f TFLOP/s  DFMA %6.2f TFLOP/s  sum %6.2f\n", name, ms, tf_m, tf_f, tf_m + tf_f);
}

int main() {
    double* d; cudaMalloc(&d, 8);
    run<8, 0>("DMMA only (8 per round)", d);
    run<0, 8>("DFMA only (8 per round)", d);
    run<8, 8>("8 DMMA + 8 DFMA", d);
    run<8, 4>("8 DMMA + 4 DFMA", d);
    run<4, 8>("4 DMMA + 8 DFMA", d);
    run<6, 2>("6 DMMA + 2 DFMA", d);
    run<2, 8>("2 DMMA + 8 DFMA", d);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
