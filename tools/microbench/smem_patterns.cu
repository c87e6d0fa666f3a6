// How many shared-memory wavefronts (SM cycles) does one LDS.128 / STS.64 cost for a given pattern of active
// lanes and addresses?  4 warps per SM (one per scheduler), 16 independent accesses per loop iteration.
// cycles per warp-instruction = SM cycles / (4 warps * instructions per warp).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1);} } while (0)

struct Pattern { const char* name; unsigned mask; int addr16[32]; };   // address of lane in 16-byte units (relative), -1 inactive

__global__ void __launch_bounds__(128) k_lds128(const int* addr16, unsigned mask, int iters, long long* cycles, double* sink) {
    extern __shared__ __align__(128) double sm[];
    for (int i = threadIdx.x; i < 8192; i += blockDim.x) sm[i] = 1.0 + i;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool act = (mask >> lane) & 1u;
    const unsigned base = (unsigned)__cvta_generic_to_shared(sm) + (unsigned)(warp * 8192 + (act ? addr16[lane] : 0) * 16);
    double acc0 = 0.0, acc1 = 0.0;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            double x, y;
            asm volatile("{ .reg .pred p; setp.ne.u32 p, %3, 0; mov.f64 %0, 0d0000000000000000; mov.f64 %1, 0d0000000000000000;\n"
                         "@p ld.shared.v2.f64 {%0, %1}, [%2]; }" : "=d"(x), "=d"(y) : "r"(base + u * 512), "r"((int)act));
            acc0 += x; acc1 += y;
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    if (acc0 + acc1 == 12345.0) sink[0] = acc0;
}

__global__ void __launch_bounds__(128) k_lds64(const int* addr8, unsigned mask, int iters, long long* cycles, double* sink) {
    extern __shared__ __align__(128) double sm[];
    for (int i = threadIdx.x; i < 8192; i += blockDim.x) sm[i] = 1.0 + i;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool act = (mask >> lane) & 1u;
    const unsigned base = (unsigned)__cvta_generic_to_shared(sm) + (unsigned)(warp * 8192 + (act ? addr8[lane] : 0) * 8);
    double acc0 = 0.0;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            double x;
            asm volatile("{ .reg .pred p; setp.ne.u32 p, %2, 0; mov.f64 %0, 0d0000000000000000;\n"
                         "@p ld.shared.f64 %0, [%1]; }" : "=d"(x) : "r"(base + u * 512), "r"((int)act));
            acc0 += x;
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    if (acc0 == 12345.0) sink[0] = acc0;
}

__global__ void __launch_bounds__(128) k_sts64(const int* addr8, unsigned mask, int iters, long long* cycles, double* sink) {
    extern __shared__ __align__(128) double sm[];
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool act = (mask >> lane) & 1u;
    const unsigned base = (unsigned)__cvta_generic_to_shared(sm) + (unsigned)(warp * 8192 + (act ? addr8[lane] : 0) * 8);
    double v = lane;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 16; ++u)
            asm volatile("{ .reg .pred p; setp.ne.u32 p, %2, 0; @p st.shared.f64 [%0], %1; }" :: "r"(base + u * 512), "d"(v), "r"((int)act) : "memory");
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    if (v == 12345.0) sink[0] = sm[0];
}


__global__ void __launch_bounds__(128) k_sts128(const int* addr16, unsigned mask, int iters, long long* cycles, double* sink) {
    extern __shared__ __align__(128) double sm[];
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool act = (mask >> lane) & 1u;
    const unsigned base = (unsigned)__cvta_generic_to_shared(sm) + (unsigned)(warp * 8192 + (act ? addr16[lane] : 0) * 16);
    double v = lane;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 16; ++u)
            asm volatile("{ .reg .pred p; setp.ne.u32 p, %2, 0; @p st.shared.v2.f64 [%0], {%1, %1}; }" :: "r"(base + u * 512), "d"(v), "r"((int)act) : "memory");
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    if (v == 12345.0) sink[0] = sm[0];
}

template <typename F>
static void fill(int* a, F f) { for (int l = 0; l < 32; ++l) a[l] = f(l); }

int main() {
    int* d_addr; long long* d_cyc; double* d_sink;
    CK(cudaMalloc(&d_addr, 32 * sizeof(int))); CK(cudaMalloc(&d_cyc, 148 * sizeof(long long))); CK(cudaMalloc(&d_sink, 8));
    CK(cudaFuncSetAttribute(k_lds128, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 8192 + 65536));
    CK(cudaFuncSetAttribute(k_sts64, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 8192 + 65536));
    const int iters = 2000;
    CK(cudaFuncSetAttribute(k_sts128, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 8192 + 65536));
    CK(cudaFuncSetAttribute(k_lds64, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 8192 + 65536));
    auto run = [&](int store, const char* name, unsigned mask, const int* addr) {
        CK(cudaMemcpy(d_addr, addr, 32 * sizeof(int), cudaMemcpyHostToDevice));
        for (int rep = 0; rep < 2; ++rep) {
            if (store == 3) k_sts128<<<148, 128, 4 * 8192 + 65536>>>(d_addr, mask, iters, d_cyc, d_sink);
            else if (store == 1) k_sts64<<<148, 128, 4 * 8192 + 65536>>>(d_addr, mask, iters, d_cyc, d_sink);
            else if (store == 2) k_lds64<<<148, 128, 4 * 8192 + 65536>>>(d_addr, mask, iters, d_cyc, d_sink);
            else k_lds128<<<148, 128, 4 * 8192 + 65536>>>(d_addr, mask, iters, d_cyc, d_sink);
            CK(cudaDeviceSynchronize());
        }
        long long c[148];
        CK(cudaMemcpy(c, d_cyc, sizeof c, cudaMemcpyDeviceToHost));
        double avg = 0; for (int i = 0; i < 148; ++i) avg += (double)c[i]; avg /= 148;
        printf("%-6s %-58s mask %08x  %.2f SM-cycles per warp-instruction\n", store == 3 ? "STS128" : (store == 1 ? "STS.64" : (store == 2 ? "LDS.64" : "LDS128")), name, mask, avg / (4.0 * iters * 16));
    };
    int a[32];
    // ---- LDS.128: addresses in 16-byte units; a quarter-warp = lanes 8g..8g+7; matrices of the 4 quarters 40 units apart (same bank alignment)
    fill(a, [](int l) { return (l >> 3) * 80 + (l & 7); });                       run(0, "all lanes, 4 x 128 B contiguous (conflict-free)", 0xffffffffu, a);
    fill(a, [](int l) { return (l & 7); });                                      run(0, "all lanes, every quarter the same 128 B (broadcast x4)", 0xffffffffu, a);
    fill(a, [](int l) { return (l >> 3) * 11; });                                run(0, "all lanes, quarter-broadcast, 4 distinct 16 B (groups 0,3,6,1)", 0xffffffffu, a);
    fill(a, [](int l) { return (l >> 3); });                                     run(0, "all lanes, quarter-broadcast, 4 adjacent 16 B", 0xffffffffu, a);
    fill(a, [](int l) { return 0; });                                            run(0, "all lanes, one 16 B (full broadcast)", 0xffffffffu, a);
    fill(a, [](int l) { return (l >> 3) * 80 + (l & 7); });                       run(0, "lanes 0,1 of each quarter (groups 0,1 x4: conflict)", 0x03030303u, a);
    fill(a, [](int l) { return (l >> 3) * 80 + (l & 7); });                       run(0, "lanes 2g,2g+1 of quarter g (8 distinct groups)", 0xc0300c03u, a);
    fill(a, [](int l) { return (l >> 3) * 80 + (l & 7); });                       run(0, "quarter 0 only, 8 lanes", 0x000000ffu, a);
    fill(a, [](int l) { return (l >> 3) * 80 + (l & 7); });                       run(0, "lanes 0-3 of q0 and 4-7 of q1 (8 distinct groups)", 0x0000f00fu, a);
    fill(a, [](int l) { return (l >> 3) * 80 + (l & 7); });                       run(0, "lanes 0-3 of q0 and 4-7 of q2", 0x00f0000fu, a);
    fill(a, [](int l) { return (l >> 3) * 80 + (l & 7); });                       run(0, "lanes 0,1 q0; 2,3 q1 only", 0x00000c03u, a);
    fill(a, [](int l) { return (l >> 3) * 80 + (l & 7); });                       run(0, "lanes 0,1 q0; 2,3 q2 only", 0x000c0003u, a);
    fill(a, [](int l) { return (l >> 3) * 80 + (l & 7); });                       run(0, "lanes 0-3 of each quarter (groups 0-3 x4)", 0x0f0f0f0fu, a);
    fill(a, [](int l) { return (l >> 3) * 80 + (l & 7); });                       run(0, "lanes 0-3 of q0,q2; 4-7 of q1,q3", 0xf00ff00fu, a);
    fill(a, [](int l) { return (l >> 3) * 80 + (l & 7); });                       run(0, "lanes 0-3 of q0,q1; 4-7 of q2,q3", 0xf0f00f0fu, a);
    fill(a, [](int l) { return (l >> 3) * 80 + ((l & 7) + 2 * (l >> 3)) % 8; });  run(0, "lanes 0,1 of each quarter, address rotated to groups 2g,2g+1", 0x03030303u, a);
    fill(a, [](int l) { return (l >> 3) * 80 + ((l & 7) + 4 * ((l >> 3) & 1)) % 8; }); run(0, "lanes 0-3 of each quarter, address groups 0-3 / 4-7 alternating", 0x0f0f0f0fu, a);
    fill(a, [](int l) { return (l >> 2) * 80 + (l & 3) + 4 * ((l >> 2) & 1); });  run(0, "4 lanes per matrix: 8 matrices, groups 0-3/4-7 alternate (conflict-free per quarter)", 0xffffffffu, a);
    fill(a, [](int l) { return (l >> 2) * 80 + (l & 3); });                       run(0, "4 lanes per matrix: 8 matrices, all groups 0-3 (2-way conflict)", 0xffffffffu, a);
    // ---- STS.64: addresses in 8-byte units
    fill(a, [](int l) { return l; });                                            run(1, "all lanes, 256 B contiguous", 0xffffffffu, a);
    fill(a, [](int l) { return (l >> 3) * 22 + (l & 7); });                       run(1, "4 groups of 64 B, stride 176 B (current K=10 layout)", 0xffffffffu, a);
    fill(a, [](int l) { return (l >> 3) * 24 + (l & 7); });                       run(1, "4 groups of 64 B, stride 192 B", 0xffffffffu, a);
    fill(a, [](int l) { return (l >> 3) * 22 + 8 + (l & 1); });                   run(1, "lanes 0,1 of each quarter, stride 176 B", 0x03030303u, a);
    fill(a, [](int l) { return (l >> 3) * 22 + 8 + (l & 1); });                   run(1, "lanes 2g,2g+1 of quarter g, stride 176 B", 0xc0300c03u, a);
    fill(a, [](int l) { return (l >> 3) * 8 + (l & 7); });                        run(1, "4 groups of 64 B contiguous (=256 B)", 0xffffffffu, a);
    // ---- LDS.64: addresses in 8-byte units
    fill(a, [](int l) { return l; });                                            run(2, "all lanes, 256 B contiguous", 0xffffffffu, a);
    fill(a, [](int l) { return (l >> 3) * 24 + (l >> 4) * 2; });                  run(2, "quarter-broadcast, 4 distinct 8 B", 0xffffffffu, a);
    fill(a, [](int l) { return 0; });                                            run(2, "full broadcast", 0xffffffffu, a);
    fill(a, [](int l) { return (l >> 3) * 160 + 2 * (l & 7); });                  run(2, "lanes 2g,2g+1 of quarter g, 16 B apart (8 distinct 8 B)", 0xc0300c03u, a);
    fill(a, [](int l) { return (l >> 3) * 160 + 2 * (l & 7); });                  run(2, "lanes 0,1 of every quarter", 0x03030303u, a);
    fill(a, [](int l) { return (l >> 3) * 160 + (l & 7); });                      run(2, "4 x 64 B contiguous, same banks (conflict)", 0xffffffffu, a);
    fill(a, [](int l) { return (l >> 3) * 168 + (l & 7); });                      run(2, "4 x 64 B contiguous, banks 0/64 B alternate", 0xffffffffu, a);
    // ---- 5 lanes per chain (K = 10): lane L = 5c + j, six chains with different matrices (56 units apart x random id)
    {
        static const int d[7] = {3, 0, 6, 1, 4, 2, 5};
        for (int t : {0, 3, 5, 7}) {
            fill(a, [&](int l) { const int c = l / 5, j = l % 5; return 56 * d[c] + 5 * ((t + c) % 8) + j; });
            char nm[96]; snprintf(nm, sizeof nm, "5 lanes/chain ring-8 matrix load t=%d, 30 lanes", t);          run(0, strdup(nm), 0x3fffffffu, a);
        }
        fill(a, [&](int l) { const int c = l / 5, j = l % 5; return 56 * d[c] + 5 * ((0 + c) % 8) + j; });       run(0, "5 lanes/chain ring-8 t=0, all 32 lanes (chain 6 = 2 lanes)", 0xffffffffu, a);
        for (int t : {0, 1}) {
            fill(a, [&](int l) { const int c = l / 5, j = l % 5; return 56 * d[c] + 40 + 5 * ((t + c) & 1) + j; });
            char nm[96]; snprintf(nm, sizeof nm, "5 lanes/chain ring-2 matrix load t=%d (2-way conflicts)", t); run(0, strdup(nm), 0x3fffffffu, a);
        }
        fill(a, [&](int l) { const int c = l / 5, j = l % 5; return 56 * d[c] + 5 * ((9 + c) % 10) + j; });      run(0, "5 lanes/chain ring-10 t=9 (wrap: groups off by 2)", 0x3fffffffu, a);
        fill(a, [&](int l) { const int c = l / 5; return 5 * c + 2; });                                         run(0, "5 lanes/chain read-back: 7 distinct 16 B, groups 5c+u", 0xffffffffu, a);
        fill(a, [&](int l) { const int c = l / 5; return 8 * c + 2; });                                         run(0, "5 lanes/chain read-back: 7 distinct 16 B, all in group 2", 0xffffffffu, a);
        fill(a, [&](int l) { const int c = l / 5, j = l % 5; return 5 * c + (j < 4 ? (j + 4 - c % 4) % 4 : 4); }); run(3, "5 lanes/chain store: 16 B per lane, 5c + perm(j)", 0xffffffffu, a);
        fill(a, [&](int l) { return l; });                                                                      run(3, "all lanes, 512 B contiguous", 0xffffffffu, a);
    }
    printf("done\n");
    return 0;
}
