// FP64 pipe microbenchmarks for sm_100a (B200): numbers that decide the forward-kernel design.
//   * DFMA throughput vs warps/SMSP and ILP (-> pipe width and dependent-issue latency)
//   * DMMA (mma.sync f64) m8n8k4 / m16n8k4 / m16n8k8 / m16n8k16 throughput
//   * broadcast LDS.64 / LDS.128 throughput (T staged in shared memory, warp-uniform address)
//   * SHFL throughput
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o build/fp64_micro tools/microbench/fp64_micro.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

template <int ILP>
__global__ void k_dfma(double* out, int iters, double a, double b) {
    double acc[ILP];
#pragma unroll
    for (int k = 0; k < ILP; ++k) acc[k] = threadIdx.x * 1e-9 + k;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
            for (int k = 0; k < ILP; ++k) acc[k] = fma(acc[k], a, b);
    }
    double s = 0;
#pragma unroll
    for (int k = 0; k < ILP; ++k) s += acc[k];
    if (s == 12345.678) out[0] = s;
}

// ---- DMMA ---------------------------------------------------------------------
__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma1684(double* d, const double* a, double b) {
    asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                 : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3]) : "d"(a[0]), "d"(a[1]), "d"(b));
}
__device__ __forceinline__ void dmma1688(double* d, const double* a, const double* b) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__device__ __forceinline__ void dmma16816(double* d, const double* a, const double* b) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                 : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                   "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

template <int SHAPE, int ILP>
__global__ void k_dmma(double* out, int iters) {
    double d[ILP][4];
    double a[8], b[4];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = 1e-3 * (threadIdx.x + i);
#pragma unroll
    for (int i = 0; i < 4; ++i) b[i] = 1e-3 * (threadIdx.x * 3 + i);
#pragma unroll
    for (int k = 0; k < ILP; ++k)
#pragma unroll
        for (int i = 0; i < 4; ++i) d[k][i] = k + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int k = 0; k < ILP; ++k) {
                if (SHAPE == 0) dmma884(d[k][0], d[k][1], a[0], b[0]);
                if (SHAPE == 1) dmma1684(d[k], a, b[0]);
                if (SHAPE == 2) dmma1688(d[k], a, b);
                if (SHAPE == 3) dmma16816(d[k], a, b);
            }
    }
    double s = 0;
#pragma unroll
    for (int k = 0; k < ILP; ++k)
#pragma unroll
        for (int i = 0; i < 4; ++i) s += d[k][i];
    if (s == 12345.678) out[0] = s;
}

// ---- broadcast LDS ---------------------------------------------------------------
template <int WIDTH>  // 8 or 16 bytes
__global__ void k_lds(double* out, int iters) {
    __shared__ __align__(16) double sm[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = i;
    __syncthreads();
    unsigned long long acc0 = 0, acc1 = 0;
    int base = 0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            int idx = (base + u * 2) & 1022;  // warp-uniform address, changes every access
            if (WIDTH == 8) {
                unsigned long long v;
                asm volatile("ld.shared.u64 %0, [%1];" : "=l"(v) : "r"((unsigned)__cvta_generic_to_shared(&sm[idx])));
                acc0 ^= v;
            } else {
                unsigned long long v0, v1;
                asm volatile("ld.shared.v2.u64 {%0,%1}, [%2];" : "=l"(v0), "=l"(v1) : "r"((unsigned)__cvta_generic_to_shared(&sm[idx])));
                acc0 ^= v0; acc1 ^= v1;
            }
        }
        base += 32;
    }
    if ((acc0 ^ acc1) == 0x123456789ull) out[0] = (double)acc0;
}

// DFMA fed by broadcast LDS: RATIO dfma per loaded double (register tiling factor), LDS.128 = 2 doubles
template <int R, int WIDTH>
__global__ void k_dfma_lds(double* out, int iters) {
    __shared__ __align__(16) double sm[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = 1.0 + 1e-9 * i;
    __syncthreads();
    double acc[R][8];
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[r][k] = r + k + threadIdx.x * 1e-9;
    double x[R];
#pragma unroll
    for (int r = 0; r < R; ++r) x[r] = 1.0 + 1e-7 * (threadIdx.x + r);
    int base = 0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            // 8 matrix elements (one "row" of 8 outputs) shared by the R chains of this thread
            double t[8];
            const double* p = &sm[(base + u * 8) & 1016];
            if (WIDTH == 8) {
#pragma unroll
                for (int k = 0; k < 8; ++k) t[k] = p[k];
            } else {
#pragma unroll
                for (int k = 0; k < 8; k += 2) { double2 v = *reinterpret_cast<const double2*>(p + k); t[k] = v.x; t[k + 1] = v.y; }
            }
#pragma unroll
            for (int r = 0; r < R; ++r)
#pragma unroll
                for (int k = 0; k < 8; ++k) acc[r][k] = fma(x[r], t[k], acc[r][k]);
        }
        base += 32;
    }
    double s = 0;
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
        for (int k = 0; k < 8; ++k) s += acc[r][k];
    if (s == 12345.678) out[0] = s;
}

__global__ void k_shfl(double* out, int iters) {
    unsigned v = threadIdx.x, w = threadIdx.x * 7;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            v = __shfl_xor_sync(0xffffffffu, v, 1) + 1;
            w = __shfl_xor_sync(0xffffffffu, w, 1) + 3;
        }
    }
    if (v + w == 0x12345u) out[0] = v;
}

struct Timer {
    cudaEvent_t a, b;
    Timer() { cudaEventCreate(&a); cudaEventCreate(&b); }
    void start() { cudaEventRecord(a); }
    float stop() { cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b); return ms; }
};

template <typename F>
float best_ms(F launch, int reps = 5) {
    Timer t; float best = 1e30f;
    launch();  // warm
    CK(cudaDeviceSynchronize());
    for (int r = 0; r < reps; ++r) { t.start(); launch(); float ms = t.stop(); if (ms < best) best = ms; }
    CK(cudaGetLastError());
    return best;
}

int main() {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    int sms = prop.multiProcessorCount;
    printf("device %s sms %d clock %d kHz\n", prop.name, sms, prop.clockRate);
    double* out; CK(cudaMalloc(&out, 64));
    const int iters = 20000;

    printf("\n== DFMA: TFLOP/s (2 flop per lane-FMA) ==\nwarps/SMSP  ILP1    ILP2    ILP4    ILP8    ILP16\n");
    for (int wps : {1, 2, 4, 8}) {
        int threads = wps * 4 * 32;  // one CTA per SM
        printf("%9d", wps);
        auto run = [&](auto kern, int ilp) {
            float ms = best_ms([&] { kern<<<sms, threads>>>(out, iters, 1.0000001, 1e-9); });
            double flop = 2.0 * ilp * 8.0 * iters * (double)threads * sms;
            printf("  %6.2f", flop / ms / 1e9);
        };
        run(k_dfma<1>, 1); run(k_dfma<2>, 2); run(k_dfma<4>, 4); run(k_dfma<8>, 8); run(k_dfma<16>, 16);
        printf("\n");
    }

    printf("\n== DMMA: TFLOP/s ==\nshape      warps/SMSP  ILP1    ILP2    ILP4    ILP8\n");
    const char* names[4] = {"m8n8k4", "m16n8k4", "m16n8k8", "m16n8k16"};
    const double flops_per[4] = {2.0 * 8 * 8 * 4, 2.0 * 16 * 8 * 4, 2.0 * 16 * 8 * 8, 2.0 * 16 * 8 * 16};
    for (int shape = 0; shape < 4; ++shape)
        for (int wps : {1, 2, 4}) {
            int threads = wps * 4 * 32;
            printf("%-10s %9d", names[shape], wps);
            auto run = [&](auto kern, int ilp) {
                float ms = best_ms([&] { kern<<<sms, threads>>>(out, iters / 4); });
                double flop = flops_per[shape] * ilp * 4.0 * (iters / 4) * (double)(threads / 32) * sms;
                printf("  %6.2f", flop / ms / 1e9);
            };
            if (shape == 0) { run(k_dmma<0, 1>, 1); run(k_dmma<0, 2>, 2); run(k_dmma<0, 4>, 4); run(k_dmma<0, 8>, 8); }
            if (shape == 1) { run(k_dmma<1, 1>, 1); run(k_dmma<1, 2>, 2); run(k_dmma<1, 4>, 4); run(k_dmma<1, 8>, 8); }
            if (shape == 2) { run(k_dmma<2, 1>, 1); run(k_dmma<2, 2>, 2); run(k_dmma<2, 4>, 4); run(k_dmma<2, 8>, 8); }
            if (shape == 3) { run(k_dmma<3, 1>, 1); run(k_dmma<3, 2>, 2); run(k_dmma<3, 4>, 4); run(k_dmma<3, 8>, 8); }
            printf("\n");
        }

    printf("\n== broadcast LDS: warp-instr per clk per SM (clock = %d kHz nominal; uses measured time) ==\n", prop.clockRate);
    for (int wps : {1, 2, 4, 8}) {
        int threads = wps * 4 * 32;
        float ms8 = best_ms([&] { k_lds<8><<<sms, threads>>>(out, iters); });
        float ms16 = best_ms([&] { k_lds<16><<<sms, threads>>>(out, iters); });
        double n = 16.0 * iters * (threads / 32);  // warp-instr per SM
        printf("warps/SMSP %d: LDS.64 %.3f Ginstr/s/SM  LDS.128 %.3f Ginstr/s/SM\n", wps, n / ms8 / 1e6, n / ms16 / 1e6);
    }

    printf("\n== DFMA fed from broadcast LDS (TFLOP/s): R chains per thread ==\n");
    for (int wps : {1, 2, 4}) {
        int threads = wps * 4 * 32;
        auto run = [&](auto kern, int R, const char* nm) {
            float ms = best_ms([&] { kern<<<sms, threads>>>(out, iters / 4); });
            double flop = 2.0 * R * 8 * 4.0 * (iters / 4) * (double)threads * sms;
            printf("  %s %6.2f", nm, flop / ms / 1e9);
        };
        printf("warps/SMSP %d:", wps);
        run(k_dfma_lds<1, 8>, 1, "R1/LDS64"); run(k_dfma_lds<1, 16>, 1, "R1/LDS128");
        run(k_dfma_lds<2, 8>, 2, "R2/LDS64"); run(k_dfma_lds<2, 16>, 2, "R2/LDS128");
        run(k_dfma_lds<4, 16>, 4, "R4/LDS128");
        printf("\n");
    }

    printf("\n== SHFL.32: warp-instr per s per SM ==\n");
    for (int wps : {1, 2, 4, 8}) {
        int threads = wps * 4 * 32;
        float ms = best_ms([&] { k_shfl<<<sms, threads>>>(out, iters); });
        double n = 16.0 * iters * (threads / 32);
        printf("warps/SMSP %d: %.3f Ginstr/s/SM\n", wps, n / ms / 1e6);
    }
    return 0;
}
