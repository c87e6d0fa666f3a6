// Compressed ("zip") forward kernel for sm_100a: the GPU form of ziphmm.zip_forward
// (/root/reference/src/IMCoalHMM/hmm.py:20-21) over the token streams produced by tokenizer.inl.
//
//   C_s      = diag(E[:,s]) T^T                      base symbol matrices (alpha' = C_s alpha)
//   C_(a,b)  = C_b C_a                               dictionary entry for "a then b", built level by level
//   alpha_0  = pi o E[:,o_0];  alpha <- C_tok alpha  one mat-vec per token;  logL = log sum(alpha) + exponents
//
// One CTA = one parameter point x one share of the chunks.  Phase 1 builds the point's dictionary (M matrices,
// each scaled by an exact power of two so that its largest entry is in [1,2), exponent kept aside) in shared
// memory.  Phase 2 runs the chains: a chain is owned by EIGHT lanes (a quarter-warp); lane q owns output rows
// q, q+8, ... and reads them with LDS.128 from a layout in which the 8 lanes of a quarter-warp always touch 8
// consecutive 16-byte units of ONE matrix, so every load is bank-conflict free no matter which matrices the four
// chains of a warp are using.  The new state is exchanged through a small double-buffered shared-memory
// buffer (one STS.64 per owned row, K/2 quarter-broadcast LDS.128 back), after which every lane holds all of
// alpha again.  The bound is the shared-memory pipe: K*K*8 bytes of matrix per chain-step against 128 B/clk/SM,
// i.e. at most 1/4 of the FP64 rate -- times the compression ratio (50-150x on the benchmark alignments).
#pragma once
#include "forward_kernels.cuh"

namespace imc {

struct ZipChunk {
    long long tok_off;   // byte offset of the chunk's first token (16-byte aligned, 16 readable bytes past the end)
    int ntok;            // tokens (the chunk's symbols 1..L-1 after compression)
    int first_sym;       // symbol at position 0
    int out_index;       // column of chain_out this chunk writes
    int pad;
};

struct ZipArgs {
    const uint8_t* tokens;
    const ZipChunk* chunks;      // sorted by ntok, descending
    int nchunks;
    int J;                       // CTAs per parameter point; CTA j owns chunks j, j+J, j+2J, ...
    const uint8_t* pairs;        // [M][2] (left, right) in level order; entries < S unused
    const int* level_start;      // [nlevels + 1]
    int nlevels;
    int M;
    int N, K, S;
    const double* pi;            // [N][K]
    const double* T;             // [N][K][K]
    const double* E;             // [N][K][S]
    double* chain_out;           // [N][out_stride]
    int out_stride;
};

template <int K>
struct ZipCfg {
    static constexpr int KP = (K + 1) & ~1;            // columns padded to an even count (16-byte units)
    static constexpr int CP = KP / 2;                  // units per row
    static constexpr int RPL = (K + 7) / 8;            // rows per lane
    static constexpr int STRIDE_D = RPL * CP * 16;     // doubles per dictionary matrix
    // Exchange buffer per chain: 2 x KP doubles (+2 of skew room), stride = 64 (mod 128) bytes, chains 2,3 of a warp
    // skewed by 16 bytes: in 16-byte bank groups the four chains then start at 0,4,1,5 (mod 8), so the two chains of
    // each half-warp store their 64-byte row blocks into disjoint bank halves (STS.64: 2 wavefronts instead of 4,
    // tools/microbench/smem_patterns.cu) AND the quarter-broadcast read-back touches four different bank groups.
    static constexpr int GS = ((2 * KP * 8 + 16 + 63) / 128 * 128 + 64) / 8;
    static constexpr int UNROLL = K <= 12 ? 4 : 1;
    static constexpr int FULL = K / 8;                 // slots in which all 8 lanes own a row
    static constexpr int REM = K % 8;                  // rows of the last, partial slot
    // The partial slot is stored REP times side by side (lane positions f*REM .. f*REM+REM-1 hold rows 8*FULL ..):
    // chain slot g of a warp lets lanes of copy g % REP own the remainder rows, so that the four quarter-warps
    // of one LDS.128 touch different bank groups instead of all hitting groups 0..REM-1.
    static constexpr int REP = REM ? 8 / REM : 1;
    __host__ __device__ static constexpr int se_doubles(int S) { return (K * S + 1) & ~1; }   // keeps what follows 16-byte aligned
    // element (row r, column c): slot r/8, unit (slot*CP + c/2)*8 + r%8 (first copy of the partial slot)
    __host__ __device__ static constexpr int off(int r, int c) {
        return (((r >> 3) * CP + (c >> 1)) * 8 + (r & 7)) * 2 + (c & 1);
    }
    __device__ static __forceinline__ void store(double* D, int r, int c, double v) {
        const int o = off(r, c);
        D[o] = v;
        if (REM && r >= 8 * FULL) {
#pragma unroll
            for (int f = 1; f < REP; ++f) D[o + 2 * f * REM] = v;
        }
    }
    static size_t smem_bytes(int M, int S, int threads) {
        size_t d = (size_t)M * STRIDE_D + (size_t)se_doubles(S) + KP + (size_t)(threads / 8) * GS;
        return d * sizeof(double) + ((size_t)M + 4) * sizeof(int);
    }
    static int max_entries(size_t budget, int S, int threads) {
        const size_t fixed = smem_bytes(0, S, threads);
        if (budget <= fixed) return 0;
        const size_t m = (budget - fixed) / (STRIDE_D * sizeof(double) + sizeof(int));
        return (int)(m > 256 ? 256 : m);
    }
};

// one token for one chain: acc = (rows of C_id owned by lane q) . al ; exchange ; al = new state
template <int K, bool PRED>
__device__ __forceinline__ void zip_step(double (&al)[ZipCfg<K>::KP], const double* dict, const int* dexp, int id,
                                         double* sb, int q, int rem_row, long long& scale, bool active) {
    using C = ZipCfg<K>;
    if (!PRED || active) {
        const double2* mp = reinterpret_cast<const double2*>(dict + (size_t)id * C::STRIDE_D) + q;
#pragma unroll
        for (int k = 0; k < C::RPL; ++k) {
            if (k < C::FULL || rem_row >= 0) {     // rem_row: row of the partial slot owned by this lane, or -1
                double s0 = 0.0, s1 = 0.0;
#pragma unroll
                for (int cp = 0; cp < C::CP; ++cp) {
                    const double2 m = mp[(k * C::CP + cp) * 8];
                    s0 = fma(m.x, al[2 * cp], s0);
                    s1 = fma(m.y, al[2 * cp + 1], s1);
                }
                sb[k < C::FULL ? q + 8 * k : rem_row] = s0 + s1;
            }
        }
        scale += dexp[id];
    }
    __syncwarp();
    if (!PRED || active) {
#ifdef IMC_ZIP_READBACK64
#pragma unroll
        for (int k = 0; k < C::KP; ++k) al[k] = sb[k];
#else
#pragma unroll
        for (int cp = 0; cp < C::CP; ++cp) {
            const double2 v = reinterpret_cast<const double2*>(sb)[cp];
            al[2 * cp] = v.x;
            al[2 * cp + 1] = v.y;
        }
#endif
    }
}

template <int K>
__device__ __forceinline__ void zip_rescale(double (&al)[ZipCfg<K>::KP], long long& scale, bool& dead, bool& isnan) {
    double sum = 0.0;
#pragma unroll
    for (int k = 0; k < K; ++k) sum += al[k];
    if (sum > 0.0 && sum < 1.7e308) {
        const int e = exponent_of(sum);
        const double f = pow2_neg(e);
#pragma unroll
        for (int k = 0; k < K; ++k) al[k] *= f;
        scale += e;
    } else {
        dead = true;
        isnan = isnan || (sum != sum) || (sum > 0.0);
    }
}

template <int K, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) zip_forward_kernel(ZipArgs a) {
    using C = ZipCfg<K>;
    constexpr int KP = C::KP, NW = THREADS / 32;
    extern __shared__ __align__(128) unsigned char zsm_raw[];
    double* dict = reinterpret_cast<double*>(zsm_raw);
    const int M = a.M, S = a.S;
    double* sE = dict + (size_t)M * C::STRIDE_D;      // [K][S]
    double* spi = sE + C::se_doubles(S);              // [KP]
    double* sbuf = spi + KP;                          // [THREADS/8][GS]
    int* dexp = reinterpret_cast<int*>(sbuf + (THREADS / 8) * C::GS);   // [M]
    int* s_next = dexp + M;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = blockIdx.x / a.J, j = blockIdx.x % a.J;
    const double* Tg = a.T + (size_t)n * K * K;
    const double* Eg = a.E + (size_t)n * K * S;
    const double* pig = a.pi + (size_t)n * K;

    // ---- phase 1: dictionary ------------------------------------------------------------------------
    for (int x = tid; x < M * C::STRIDE_D; x += THREADS) dict[x] = 0.0;
    for (int x = tid; x < K * S; x += THREADS) sE[x] = Eg[x];
    for (int x = tid; x < KP; x += THREADS) spi[x] = x < K ? pig[x] : 0.0;
    for (int x = tid; x < (THREADS / 8) * C::GS; x += THREADS) sbuf[x] = 0.0;   // the padding column of odd K stays 0
    if (tid == 0) *s_next = 0;
    __syncthreads();
    for (int lv = -1; lv < a.nlevels; ++lv) {
        const int lo = lv < 0 ? 0 : a.level_start[lv], hi = lv < 0 ? S : a.level_start[lv + 1];
        for (int e = lo + warp; e < hi; e += NW) {
            double* D = dict + (size_t)e * C::STRIDE_D;
            double mx = 0.0;
            bool bad = false;
            int ebase = 0;
            if (lv < 0) {     // C_s[r][c] = E[r][s] * T[c][r]
                for (int x = lane; x < K * K; x += 32) {
                    const int r = x / K, c = x % K;
                    const double v = sE[r * S + e] * Tg[c * K + r];
                    C::store(D, r, c, v);
                    mx = fmax(mx, fabs(v));
                    bad = bad || !(fabs(v) < 1.7e308);
                }
            } else {          // C_(l,r) = C_r C_l
                const int il = a.pairs[2 * e], ir = a.pairs[2 * e + 1];
                const double* A = dict + (size_t)il * C::STRIDE_D;
                const double* B = dict + (size_t)ir * C::STRIDE_D;
                ebase = dexp[il] + dexp[ir];
                for (int x = lane; x < K * K; x += 32) {
                    const int r = x / K, c = x % K;
                    double acc = 0.0;
#pragma unroll 4
                    for (int k = 0; k < K; ++k) acc = fma(B[C::off(r, k)], A[C::off(k, c)], acc);
                    C::store(D, r, c, acc);
                    mx = fmax(mx, fabs(acc));
                    bad = bad || !(fabs(acc) < 1.7e308);
                }
            }
#pragma unroll
            for (int m = 16; m >= 1; m >>= 1) mx = fmax(mx, shfl_xor_f64(mx, m));
            bad = __any_sync(0xffffffffu, bad);
            int ex = 0;
            if (mx > 0.0 && !bad) {
                ex = exponent_of(mx);
                if (ex < -1000) ex = -1000;
                const double f = pow2_neg(ex);
                for (int x = lane; x < K * K; x += 32) C::store(D, x / K, x % K, D[C::off(x / K, x % K)] * f);
            }
            if (lane == 0) dexp[e] = ebase + ex;
        }
        __syncthreads();
    }

    // ---- phase 2: chains ----------------------------------------------------------------------------
    const int q = lane & 7, grp = lane >> 3;
    int rem_row = -1;
    if (C::REM) {
        const int p = q - (grp % C::REP) * C::REM;
        if (p >= 0 && p < C::REM) rem_row = 8 * C::FULL + p;
    }
    double* sb0 = sbuf + (size_t)(warp * 4 + grp) * C::GS + (grp >> 1) * 2;
    const int cnt = (a.nchunks - j + a.J - 1) / a.J;       // chunks owned by this CTA
    for (;;) {
        int base = 0;
        if (lane == 0) base = atomicAdd(s_next, 4);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base >= cnt) break;
        const int ci = base + grp;
        const bool have = ci < cnt;
        const ZipChunk ch = a.chunks[j + (have ? ci : base) * a.J];
        const int nt = have ? ch.ntok : 0;
        const uint4* tp = reinterpret_cast<const uint4*>(a.tokens + ch.tok_off);
        int maxnt = nt;
#pragma unroll
        for (int m = 16; m >= 8; m >>= 1) maxnt = max(maxnt, __shfl_xor_sync(0xffffffffu, maxnt, m));

        double al[KP];
#pragma unroll
        for (int k = 0; k < KP; ++k) al[k] = k < K ? spi[k] * sE[k * S + ch.first_sym] : 0.0;
        long long scale = 0;
        bool dead = false, isnan = false;
        int buf = 0;
        uint4 cur = make_uint4(0, 0, 0, 0);
        if (nt > 0) cur = tp[0];
        for (int blk = 0; blk * 16 < maxnt; ++blk) {
            uint4 nxt = make_uint4(0, 0, 0, 0);
            if ((blk + 1) * 16 < nt) nxt = tp[blk + 1];
            const int rem = nt - blk * 16;
            const uint32_t w[4] = {cur.x, cur.y, cur.z, cur.w};
            if (__all_sync(0xffffffffu, rem >= 16)) {
#pragma unroll
                for (int wi = 0; wi < 4; ++wi) {
                    uint32_t wv = w[wi];
#pragma unroll C::UNROLL
                    for (int b = 0; b < 4; ++b) {
                        const int id = wv & 0xffu;
                        wv >>= 8;
                        zip_step<K, false>(al, dict, dexp, id, sb0 + buf * KP, q, rem_row, scale, true);
                        buf ^= 1;
                    }
                    if (wi & 1) zip_rescale<K>(al, scale, dead, isnan);
                }
            } else {
#pragma unroll
                for (int wi = 0; wi < 4; ++wi) {
                    uint32_t wv = w[wi];
#pragma unroll 1
                    for (int b = 0; b < 4; ++b) {
                        const int id = wv & 0xffu;
                        wv >>= 8;
                        zip_step<K, true>(al, dict, dexp, id, sb0 + buf * KP, q, rem_row, scale, wi * 4 + b < rem);
                        buf ^= 1;
                    }
                    if (wi & 1) zip_rescale<K>(al, scale, dead, isnan);
                }
            }
            cur = nxt;
        }
        double sum = 0.0;
#pragma unroll
        for (int k = 0; k < K; ++k) sum += al[k];
        double result;
        if (dead || !(sum > 0.0)) result = (isnan || sum != sum) ? __longlong_as_double(0x7ff8000000000000LL) : -INFINITY;
        else result = log(sum) + (double)scale * LN2;
        if (have && q == 0) a.chain_out[(size_t)n * a.out_stride + ch.out_index] = result;
    }
}

}  // namespace imc
