// Second round: emission strategies on top of the 10-accumulator ordering, with realistic symbol streams
// (5% non-zero, per-lane independent).  Numbers: % of 64 FP64 lanes/clk/SM counting only the 50 mat-vec
// FP64 instructions per position as useful (so strategies are directly comparable).
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <vector>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1);} } while (0)
constexpr int K = 10, H = 5;

__device__ __forceinline__ double shfl_xor_f64(double v, int mask) {
    int lo = __double2loint(v), hi = __double2hiint(v);
    lo = __shfl_xor_sync(0xffffffffu, lo, mask);
    hi = __shfl_xor_sync(0xffffffffu, hi, mask);
    return __hiloint2double(hi, lo);
}

template <int ACC2>
__device__ __forceinline__ void matvec(const double (&al)[K], const double (&To)[H][H], const double (&Tx)[H][H], double (&b)[H]) {
    if (!ACC2) {
#pragma unroll
        for (int jj = 0; jj < H; ++jj) b[jj] = al[0] * To[0][jj];
#pragma unroll
        for (int ii = 1; ii < H; ++ii)
#pragma unroll
            for (int jj = 0; jj < H; ++jj) b[jj] = fma(al[ii], To[ii][jj], b[jj]);
#pragma unroll
        for (int ii = 0; ii < H; ++ii)
#pragma unroll
            for (int jj = 0; jj < H; ++jj) b[jj] = fma(al[H + ii], Tx[ii][jj], b[jj]);
    } else {
        double c[H];
#pragma unroll
        for (int jj = 0; jj < H; ++jj) { b[jj] = al[0] * To[0][jj]; c[jj] = al[H] * Tx[0][jj]; }
#pragma unroll
        for (int ii = 1; ii < H; ++ii)
#pragma unroll
            for (int jj = 0; jj < H; ++jj) { b[jj] = fma(al[ii], To[ii][jj], b[jj]); c[jj] = fma(al[H + ii], Tx[ii][jj], c[jj]); }
#pragma unroll
        for (int jj = 0; jj < H; ++jj) b[jj] += c[jj];
    }
}

// EM: 0 none, 1 LDS row always, 2 register select-3 always, 3 fold: vote + register ratio select + predicated mul,
//     4 fold: vote + LDS ratio row for lanes that need it, 5 LDS row prefetched one position ahead
template <int ACC2, int EM>
__global__ void __launch_bounds__(128) k_step(double* out, const double* Tg, const uint32_t* words, int nwords, int rowstride) {
    extern __shared__ __align__(16) double sm[];
    const int lane = threadIdx.x & 31;
    double* erow = sm + (size_t)threadIdx.x * rowstride;
    for (int i = 0; i < 24; ++i) erow[i] = 1.0 - 1e-9 * i;
    double To[H][H], Tx[H][H];
#pragma unroll
    for (int ii = 0; ii < H; ++ii)
#pragma unroll
        for (int jj = 0; jj < H; ++jj) { To[ii][jj] = Tg[(ii * H + jj + lane) % 64]; Tx[ii][jj] = Tg[(ii * H + jj + 25 + lane) % 64]; }
    double R0[H], R1[H], R2[H];
#pragma unroll
    for (int jj = 0; jj < H; ++jj) { R0[jj] = 1.0 - 1e-9 * jj; R1[jj] = 0.9 + 1e-9 * jj; R2[jj] = 1.0 + 1e-8 * (jj + lane); }
    double al[K];
#pragma unroll
    for (int k = 0; k < K; ++k) al[k] = 0.1 + 1e-3 * k + 1e-6 * lane;
    const uint32_t* wp = words + (blockIdx.x * 128 + threadIdx.x) / 2 % 4096;
    uint32_t word = wp[0];
    for (int w = 0; w < nwords; ++w) {
        const uint32_t next = wp[(size_t)min(w + 1, nwords - 1) * 4096];
        uint32_t ww = word;
        double epre[H];
        if (EM == 5) {
            const double* e = erow + (ww & 3u) * 6;
#pragma unroll
            for (int jj = 0; jj < H; ++jj) epre[jj] = e[jj];
        }
#pragma unroll 4
        for (int t = 0; t < 16; ++t) {
            const int o = ww & 3u;
            ww >>= 2;
            double b[H];
            matvec<ACC2>(al, To, Tx, b);
            if (EM == 1) {
                const double* e = erow + o * 6;
#pragma unroll
                for (int jj = 0; jj < H; ++jj) b[jj] *= e[jj];
            } else if (EM == 2) {
#pragma unroll
                for (int jj = 0; jj < H; ++jj) { const double e = o == 0 ? R0[jj] : (o == 1 ? R1[jj] : R2[jj]); b[jj] *= e; }
            } else if (EM == 3) {
                if (__any_sync(0xffffffffu, o != 0)) {
#pragma unroll
                    for (int jj = 0; jj < H; ++jj) { const double e = o == 1 ? R1[jj] : R2[jj]; if (o != 0) b[jj] *= e; }
                }
            } else if (EM == 4) {
                if (__any_sync(0xffffffffu, o != 0)) {
                    if (o != 0) {
                        const double* e = erow + o * 6;
#pragma unroll
                        for (int jj = 0; jj < H; ++jj) b[jj] *= e[jj];
                    }
                }
            } else if (EM == 5) {
#pragma unroll
                for (int jj = 0; jj < H; ++jj) b[jj] *= epre[jj];
                const double* e = erow + (ww & 3u) * 6;   // next position's row (garbage past the word: harmless here)
#pragma unroll
                for (int jj = 0; jj < H; ++jj) epre[jj] = e[jj];
            }
#pragma unroll
            for (int jj = 0; jj < H; ++jj) { al[jj] = b[jj]; al[H + jj] = shfl_xor_f64(b[jj], 1); }
        }
        double sum = 0;
#pragma unroll
        for (int k = 0; k < K; ++k) sum += al[k];
        const int e = ((__double2hiint(sum) >> 20) & 0x7ff) - 1023;
        const double f = __hiloint2double((1023 - e) << 20, 0);
#pragma unroll
        for (int k = 0; k < K; ++k) al[k] *= f;
        word = next;
    }
    double s = 0;
#pragma unroll
    for (int k = 0; k < K; ++k) s += al[k];
    if (s == 12345.678) out[0] = s;
}

template <typename F>
float best_ms(F launch) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    launch(); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) { cudaEventRecord(a); launch(); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms; }
    CK(cudaGetLastError());
    return best;
}

int main() {
    double *out, *Tg; CK(cudaMalloc(&out, 64)); CK(cudaMalloc(&Tg, 64 * 8));
    double h[64]; for (int i = 0; i < 64; ++i) h[i] = 0.01 + 0.001 * i; CK(cudaMemcpy(Tg, h, sizeof h, cudaMemcpyHostToDevice));
    const int nwords = 4000;
    std::vector<uint32_t> hw((size_t)nwords * 4096);
    uint64_t st = 88172645463325252ull;
    for (auto& w : hw) {
        uint32_t v = 0;
        for (int t = 0; t < 16; ++t) {
            st ^= st << 13; st ^= st >> 7; st ^= st << 17;
            const unsigned r = (unsigned)(st >> 33) % 100;
            const unsigned o = r < 95 ? 0 : (r < 96 ? 1 : 2);
            v |= o << (2 * t);
        }
        w = v;
    }
    uint32_t* dw; CK(cudaMalloc(&dw, hw.size() * 4)); CK(cudaMemcpy(dw, hw.data(), hw.size() * 4, cudaMemcpyHostToDevice));
    const int rowstride = 26;
    const size_t smem = 128 * rowstride * 8;
    printf("useful = 50 FP64 instr/position; %% of 64 lanes/clk/SM @1.965GHz; symbols 95/1/4%% iid per chain\n");
    for (int cps : {2, 3}) {
        auto report = [&](const char* name, float ms) {
            double lane_ops = 50.0 * 16.0 * nwords * 128.0 * cps;
            printf("  ctas/SM %d  %-44s %8.3f ms  %5.1f%%\n", cps, name, ms, lane_ops / (ms * 1e-3 * 1.965e9 * 64.0) * 100.0);
        };
#define RUN(A, E, NAME) report(NAME, best_ms([&] { k_step<A, E><<<148 * cps, 128, smem>>>(out, Tg, dw, nwords, rowstride); }))
        RUN(0, 0, "5acc  no emission");
        RUN(1, 0, "10acc no emission");
        RUN(0, 1, "5acc  LDS row always (current unfolded)");
        RUN(1, 1, "10acc LDS row always");
        RUN(1, 2, "10acc register select-3 always");
        RUN(1, 3, "10acc fold: vote + reg ratio select");
        RUN(0, 3, "5acc  fold: vote + reg ratio select");
        RUN(1, 4, "10acc fold: vote + LDS ratio (current fold)");
        RUN(1, 5, "10acc LDS row prefetched 1 ahead");
    }
    return 0;
}
