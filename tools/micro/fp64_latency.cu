// Dependent-issue latency of DMMA.8x8x4, DFMA, DMUL and a shared-memory round trip on sm_100a (one warp, clock64).
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma884a(double& d0, double& d1, double& a, double b) {   // D feeds A of the next one (the forward step)
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%4, %5};" : "=d"(d0), "=d"(d1) : "d"(a), "d"(b), "d"(0.0), "d"(0.0));
    a = d0;
}
__global__ void lat(long long* out, double seed, int warps) {
    __shared__ double sm[64];
    const int N = 4096;
    double d0 = seed, d1 = seed, a = seed, b = seed;
    long long t0, t1;
    if ((threadIdx.x >> 5) >= warps) return;
    t0 = clock64();
    for (int i = 0; i < N; ++i) dmma884(d0, d1, a, b);
    t1 = clock64();
    if (threadIdx.x == 0) out[0] = (t1 - t0) / N;
    t0 = clock64();
    for (int i = 0; i < N; ++i) dmma884a(d0, d1, a, b);
    t1 = clock64();
    if (threadIdx.x == 0) out[1] = (t1 - t0) / N;
    t0 = clock64();
    for (int i = 0; i < N; ++i) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d0) : "d"(a), "d"(b));
    t1 = clock64();
    if (threadIdx.x == 0) out[2] = (t1 - t0) / N;
    t0 = clock64();
    for (int i = 0; i < N; ++i) asm volatile("mul.rn.f64 %0, %0, %1;" : "+d"(d1) : "d"(b));
    t1 = clock64();
    if (threadIdx.x == 0) out[3] = (t1 - t0) / N;
    sm[threadIdx.x & 63] = d0;
    __syncwarp();
    t0 = clock64();
    for (int i = 0; i < N; ++i) {
        asm volatile("st.shared.f64 [%0], %1;" :: "r"((unsigned)__cvta_generic_to_shared(sm + (threadIdx.x & 31))), "d"(d0) : "memory");
        __syncwarp();
        asm volatile("ld.shared.f64 %0, [%1];" : "=d"(d0) : "r"((unsigned)__cvta_generic_to_shared(sm + ((threadIdx.x + 1) & 31))) : "memory");
        __syncwarp();
    }
    t1 = clock64();
    if (threadIdx.x == 0) out[4] = (t1 - t0) / N;
    t0 = clock64();
    for (int i = 0; i < N; ++i) d1 = __shfl_xor_sync(0xffffffffu, d1, 1);
    t1 = clock64();
    if (threadIdx.x == 0) out[5] = (t1 - t0) / N;
    if (d0 + d1 + a == 123.0) out[7] = 1;
}
int main() {
    long long* d; cudaMalloc(&d, 64);
    long long h[8];
    for (int warps : {1, 4, 8, 16}) {
        lat<<<1, 512>>>(d, 1e-300, warps);
        cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
        printf("warps %2d: DMMA acc-chain %lld clk, DMMA D->A chain %lld, DFMA %lld, DMUL %lld, STS-sync-LDS-sync %lld, SHFL.f64 %lld\n", warps, h[0], h[1], h[2], h[3], h[4], h[5]);
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
}
