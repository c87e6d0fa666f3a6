// Forward log-likelihood kernels for sm_100a (B200).
//
// What they compute (reference semantics: /root/reference/src/IMCoalHMM/hmm.py:19-21 -> ziphmm.zip_forward,
// summed over forwarders at likelihood.py:33):
//     alpha_0 = pi o E[:,o_0];  alpha_t = (T^T alpha_{t-1}) o E[:,o_t];  logL = log sum_j alpha_L-1[j]
// for every chain = (parameter point n, sequence stream s).  Scaling is by exact powers of two (the
// exponent of sum(alpha) is moved into an integer accumulator), so no rounding is introduced by the
// rescale and logL = ln2 * sum(exponents) + log(sum(alpha_final)).
//
// Measured on this pool's B200 (tools/microbench/fp64_micro.cu, profiles/r01_fp64_micro.txt):
//   DFMA 34.2 TFLOP/s, DMMA.8x8x4 37.1 TFLOP/s, broadcast LDS only 8 B/clk/SM.  Hence T lives in REGISTERS:
//   * fwd_pair_kernel<K>   (small even K, e.g. 10): two lanes per chain, each owns K/2 output states and the
//                          matching K x K/2 slice of T in registers; 10 SHFL per step exchange the halves.
//   * fwd_dmma_kernel<K,MT> (K = 16..64): chains are rows of an m8n8k4 FP64 MMA, T is held as B fragments
//                          (K*K/32 doubles per lane); the C fragment of step t is the A fragment of step t+1
//                          under a fixed permutation of the state order, so no data moves between steps.
//   * fwd_generic_kernel   any K <= 128: state in shared memory, rescale every step.  Correctness baseline
//                          and the path for parameter points whose per-step decay could underflow 16 steps.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace imc {

struct StreamInfo {
    long long base;  // index of word 0 of this stream in the packed word array; word w is at base + 32*w
    int nwords;      // ceil(len / 16)
    int len;         // symbols
};

struct FwdArgs {
    const uint32_t* words;
    const StreamInfo* streams;
    int nstreams;
    int N, K, S;
    const double* pi;  // [N][K]
    const double* T;   // [N][K][K]
    const double* E;   // [N][K][S]
    double* chain_out; // [N][nstreams]
    long long nchains; // N * nstreams
    int fold_sym;      // most frequent symbol of the set (its emission column is folded into T where possible)
};

constexpr double LN2 = 0.693147180559945309417232121458;

__device__ __forceinline__ int exponent_of(double x) {
    return ((__double2hiint(x) >> 20) & 0x7ff) - 1023;
}
__device__ __forceinline__ double pow2_neg(int e) {  // 2^-e for e in [-1023, 1023]
    return __hiloint2double((1023 - e) << 20, 0);
}
// result of a finished chain: log(sum) + scale*ln2; -inf when the sequence is impossible under the model
__device__ __forceinline__ double finish_logl(double sum, int scale) {
    if (sum == 0.0) return -INFINITY;
    return log(sum) + (double)scale * LN2;
}
__device__ __forceinline__ double shfl_xor_f64(double v, int mask) {
    int lo = __double2loint(v), hi = __double2hiint(v);
    lo = __shfl_xor_sync(0xffffffffu, lo, mask);
    hi = __shfl_xor_sync(0xffffffffu, hi, mask);
    return __hiloint2double(hi, lo);
}

// ------------------------------------------------------------------------------------------------
// generic kernel: one thread per chain, one parameter point per CTA (blockIdx.y), state in shared memory
// ------------------------------------------------------------------------------------------------
__global__ void fwd_generic_kernel(FwdArgs a) {
    extern __shared__ double sm[];
    const int K = a.K, S = a.S, nt = blockDim.x, tid = threadIdx.x;
    const int theta = blockIdx.y;
    double* Ts = sm;               // [K][K]
    double* Es = Ts + K * K;       // [4][K], row 3 (padding code) = 1
    double* A = Es + 4 * K;        // [K][nt]
    double* B = A + (size_t)K * nt;
    const double* Tg = a.T + (size_t)theta * K * K;
    const double* Eg = a.E + (size_t)theta * K * S;
    const double* pig = a.pi + (size_t)theta * K;
    for (int x = tid; x < K * K; x += nt) Ts[x] = Tg[x];
    for (int x = tid; x < 4 * K; x += nt) {
        const int s = x / K, j = x % K;
        Es[x] = s < S ? Eg[j * S + s] : 1.0;
    }
    __syncthreads();
    const int stream = blockIdx.x * nt + tid;
    if (stream >= a.nstreams) return;
    const StreamInfo si = a.streams[stream];
    const uint32_t* wp = a.words + si.base;
    int scale = 0;
    bool dead = false;
    for (int w = 0; w < si.nwords; ++w) {
        uint32_t word = wp[(long long)w * 32];
        const int hi = min(16, si.len - 16 * w);
        for (int t = 0; t < hi; ++t) {
            const int o = word & 3u;
            word >>= 2;
            double sum = 0.0;
            if (w == 0 && t == 0) {
                for (int j = 0; j < K; ++j) {
                    const double v = pig[j] * Es[o * K + j];
                    A[j * nt + tid] = v;
                    sum += v;
                }
            } else {
                for (int j = 0; j < K; ++j) {
                    double acc = 0.0;
                    for (int i = 0; i < K; ++i) acc = fma(A[i * nt + tid], Ts[i * K + j], acc);
                    acc *= Es[o * K + j];
                    B[j * nt + tid] = acc;
                    sum += acc;
                }
                double* tmp = A; A = B; B = tmp;
            }
            if (!(sum > 0.0)) { dead = dead || (sum == 0.0); continue; }
            const int e = exponent_of(sum);
            const double f = pow2_neg(e);
            for (int j = 0; j < K; ++j) A[j * nt + tid] *= f;
            scale += e;
        }
    }
    double sum = 0.0;
    for (int j = 0; j < K; ++j) sum += A[j * nt + tid];
    a.chain_out[(size_t)theta * a.nstreams + stream] = dead ? -INFINITY : finish_logl(sum, scale);
}

// ------------------------------------------------------------------------------------------------
// lane-pair DFMA kernel: 16 chains per warp, T slice in registers
//
// ncu (profiles/r01_pair10_v1.txt) showed the first version limited by the shared-memory data pipe:
// 14 LDS wavefronts (per-lane emission rows) + 10 SHFL wavefronts per warp-step against 27.5 SM-cycles
// of FP64 work.  This version folds the emission column of the most frequent symbol s0 into the register
// copy of T (T0[i][j] = T[i][j] * E[j][s0]); a step whose symbol is s0 is then pure DFMA + the 10 SHFL, and
// other symbols multiply by the ratio row E[:,o]/E[:,s0] (one LDS per lane that needs it, under a warp
// vote).  Parameter points with E[j][s0] == 0 cannot be folded and take the unfolded loop.
// ------------------------------------------------------------------------------------------------
template <int K>
struct PairCfg {
    static constexpr int H = K / 2;
    static constexpr int Hp = (H + 1) & ~1;        // padded to an even count (16-byte rows)
    // doubles of E table per lane: [4 symbols][Hp] + 2 of padding so that the per-lane stride is an ODD
    // multiple of 16 bytes: the 8 lanes of a quarter-warp then fall into 8 different 16-byte bank groups
    // (with a 192-byte stride they fell into 2 and every LDS.128 replayed 4x).
    static constexpr int ROW = 4 * Hp + 2;
    static constexpr int THREADS = 128;
    static constexpr size_t smem_bytes() { return (size_t)THREADS * ROW * sizeof(double); }
};

// b = (T slice)^T alpha for this lane's H output states
template <int K>
__device__ __forceinline__ void pair_matvec(const double (&al)[K], const double (&To)[K / 2][K / 2],
                                            const double (&Tx)[K / 2][K / 2], double (&b)[K / 2]) {
    constexpr int H = K / 2;
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
}

template <int K>
__device__ __forceinline__ void pair_exchange(double (&al)[K], const double (&b)[K / 2]) {
    constexpr int H = K / 2;
#pragma unroll
    for (int jj = 0; jj < H; ++jj) {
        al[jj] = b[jj];
        al[H + jj] = shfl_xor_f64(b[jj], 1);
    }
}

// FOLD: To/Tx already carry E[:,s0]; erow rows hold E[:,o]/E[:,s0] (row s0 is never read)
template <int K, bool FOLD>
__device__ __forceinline__ void pair_step(double (&al)[K], const double (&To)[K / 2][K / 2],
                                          const double (&Tx)[K / 2][K / 2], const double* esm, int o, int s0) {
    constexpr int H = K / 2, Hp = PairCfg<K>::Hp;
    double b[H];
    pair_matvec<K>(al, To, Tx, b);
    if (FOLD) {
        if (__any_sync(0xffffffffu, o != s0)) {
            if (o != s0) {
                const double* erow = esm + o * Hp;
#pragma unroll
                for (int jj = 0; jj < H; ++jj) b[jj] *= erow[jj];
            }
        }
    } else {
        const double* erow = esm + o * Hp;
#pragma unroll
        for (int jj = 0; jj < H; ++jj) b[jj] *= erow[jj];
    }
    pair_exchange<K>(al, b);
}

struct PairChain {
    const uint32_t* wp;
    int nw, len, maxnw;
    int scale;
    bool dead, isnan;
    double result;
};

template <int K, bool FOLD>
__device__ __forceinline__ void pair_main_loop(double (&al)[K], const double (&To)[K / 2][K / 2],
                                               const double (&Tx)[K / 2][K / 2], const double* esm, int s0,
                                               PairChain& c, uint32_t word) {
    for (int w = 0; w < c.maxnw; ++w) {
        const uint32_t next = c.wp[(long long)min(w + 1, c.nw - 1) * 32];
        const int lo = (w == 0) ? 1 : 0;
        const int hi = max(0, min(16, c.len - 16 * w));
        const bool full = (lo == 0) && (hi == 16);
        if (__all_sync(0xffffffffu, full)) {
            uint32_t ww = word;
#pragma unroll 4
            for (int t = 0; t < 16; ++t) {
                const int o = ww & 3u;
                ww >>= 2;
                pair_step<K, FOLD>(al, To, Tx, esm, o, s0);
            }
        } else {
            uint32_t ww = word >> (2 * lo);
#pragma unroll 1
            for (int t = lo; t < 16; ++t) {
                const int o = ww & 3u;
                ww >>= 2;
                double old[K];
#pragma unroll
                for (int k = 0; k < K; ++k) old[k] = al[k];
                pair_step<K, FOLD>(al, To, Tx, esm, o, s0);
                const bool act = t < hi;
#pragma unroll
                for (int k = 0; k < K; ++k) al[k] = act ? al[k] : old[k];
            }
        }
        double sum = 0.0;
#pragma unroll
        for (int k = 0; k < K; ++k) sum += al[k];
        if (sum > 0.0) {
            const int e = exponent_of(sum);
            const double f = pow2_neg(e);
#pragma unroll
            for (int k = 0; k < K; ++k) al[k] *= f;
            c.scale += e;
            sum *= f;
        } else if (w < c.nw) {
            c.dead = true;                     // sum == 0: impossible observation; NaN: bad input
            c.isnan = c.isnan || (sum != sum);
        }
        if (w == c.nw - 1)
            c.result = c.dead ? (c.isnan ? __longlong_as_double(0x7ff8000000000000LL) : -INFINITY) : finish_logl(sum, c.scale);
        word = next;
    }
}

template <int K>
__global__ void __launch_bounds__(PairCfg<K>::THREADS) fwd_pair_kernel(FwdArgs a) {
    using C = PairCfg<K>;
    constexpr int H = C::H, Hp = C::Hp;
    extern __shared__ __align__(16) double esm_all[];
    const int lane = threadIdx.x & 31, h = lane & 1, wib = threadIdx.x >> 5;
    double* esm = esm_all + (size_t)(wib * 32 + lane) * C::ROW;
    long long chain = ((long long)blockIdx.x * (C::THREADS / 32) + wib) * 16 + (lane >> 1);
    const bool valid = chain < a.nchains;
    if (!valid) chain = a.nchains - 1;
    const int theta = (int)(chain / a.nstreams), s = (int)(chain % a.nstreams);
    const StreamInfo si = a.streams[s];
    const double* Tg = a.T + (size_t)theta * K * K;
    const double* Eg = a.E + (size_t)theta * K * a.S;
    const double* pig = a.pi + (size_t)theta * K;
    const int s0 = a.fold_sym;

    // can the emission of symbol s0 be folded into T for this parameter point?  (needs E[j][s0] > 0 and finite ratios)
    bool fold_ok = s0 >= 0 && s0 < a.S;
    if (fold_ok) {
#pragma unroll
        for (int jj = 0; jj < H; ++jj) {
            const double e0 = Eg[(h * H + jj) * a.S + s0];
            fold_ok = fold_ok && (e0 > 0.0);
            for (int sym = 0; sym < a.S; ++sym) {
                const double r = Eg[(h * H + jj) * a.S + sym] / e0;
                fold_ok = fold_ok && (r == r) && (r < 1e300);
            }
        }
    }
    const bool fold = __all_sync(0xffffffffu, fold_ok);

    // T slice: this lane produces output states h*H .. h*H+H-1.  To: inputs from its own half, Tx: from the partner's.
    double To[H][H], Tx[H][H];
#pragma unroll
    for (int ii = 0; ii < H; ++ii)
#pragma unroll
        for (int jj = 0; jj < H; ++jj) {
            const double e0 = fold ? Eg[(h * H + jj) * a.S + s0] : 1.0;
            To[ii][jj] = Tg[(h * H + ii) * K + h * H + jj] * e0;
            Tx[ii][jj] = Tg[((1 - h) * H + ii) * K + h * H + jj] * e0;
        }
#pragma unroll
    for (int sym = 0; sym < 4; ++sym)
#pragma unroll
        for (int jj = 0; jj < Hp; ++jj) {
            double v = 1.0;                                   // padding code 3 and padded columns
            if (sym < a.S && jj < H) {
                v = Eg[(h * H + jj) * a.S + sym];
                if (fold) v /= Eg[(h * H + jj) * a.S + s0];
            }
            esm[sym * Hp + jj] = v;
        }
    __syncwarp();

    PairChain c;
    c.wp = a.words + si.base;
    c.nw = si.nwords;
    c.len = si.len;
    c.maxnw = c.nw;
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) c.maxnw = max(c.maxnw, __shfl_xor_sync(0xffffffffu, c.maxnw, m));
    c.scale = 0; c.dead = false; c.isnan = false; c.result = 0.0;

    double al[K];  // [own half | partner half]
    const uint32_t word = c.wp[0];
    {   // position 0: alpha = pi o E[:, o_0]  (true emission, not the ratio)
        const int o = word & 3u;
        double b[H];
#pragma unroll
        for (int jj = 0; jj < H; ++jj) b[jj] = pig[h * H + jj] * (o < a.S ? Eg[(h * H + jj) * a.S + o] : 1.0);
        pair_exchange<K>(al, b);
    }
    if (fold) pair_main_loop<K, true>(al, To, Tx, esm, s0, c, word);
    else pair_main_loop<K, false>(al, To, Tx, esm, s0, c, word);
    if (valid && h == 0) a.chain_out[chain] = c.result;
}

// ------------------------------------------------------------------------------------------------
// DMMA kernel: chains are rows of m8n8k4 tiles; T as B fragments in registers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

template <int K>
struct DmmaCfg {
    static constexpr int NT = (K + 7) / 8;                 // output tiles of 8 states
    static constexpr int REM = K - 8 * (NT - 1);           // valid states in the last tile (1..8)
    static constexpr bool LAST_ONE = REM <= 4;             // last tile packed into its r=0 registers only
    static constexpr int KS = 2 * (NT - 1) + (LAST_ONE ? 1 : 2);  // k-slices of 4 states
    // state held at column n (0..7) of C tile t; -1 = padding
    __host__ __device__ static constexpr int state_of(int t, int n) {
        if (t < NT - 1 || !LAST_ONE) { return (8 * t + n < K) ? 8 * t + n : -1; }
        return ((n & 1) == 0 && (n >> 1) < REM) ? 8 * t + (n >> 1) : -1;
    }
};

template <int K, int MT>
struct DmmaState {
    double c[MT][DmmaCfg<K>::NT][2];
};

// one forward step for MT tiles of 8 chains; o[mt] = symbol of this lane's chain in tile mt
template <int K, int MT>
__device__ __forceinline__ void dmma_step(DmmaState<K, MT>& st, const double (&Bf)[DmmaCfg<K>::KS][DmmaCfg<K>::NT],
                                          const double* esm, const int (&o)[MT], int q) {
    using C = DmmaCfg<K>;
    constexpr int NT = C::NT, KS = C::KS;
    double n[MT][NT][2];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int t = 0; t < NT; ++t) { n[mt][t][0] = 0.0; n[mt][t][1] = 0.0; }
#pragma unroll
    for (int sl = 0; sl < KS; ++sl)
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
            for (int t = 0; t < NT; ++t) dmma884(n[mt][t][0], n[mt][t][1], st.c[mt][sl >> 1][sl & 1], Bf[sl][t]);
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int t = 0; t < NT; ++t) {
            const double2 e = *reinterpret_cast<const double2*>(esm + ((o[mt] * NT + t) * 4 + q) * 2);
            st.c[mt][t][0] = n[mt][t][0] * e.x;
            st.c[mt][t][1] = n[mt][t][1] * e.y;
        }
}

template <int K, int MT>
__global__ void __launch_bounds__(128) fwd_dmma_kernel(FwdArgs a) {
    using C = DmmaCfg<K>;
    constexpr int NT = C::NT, KS = C::KS;
    __shared__ __align__(16) double esm[4 * NT * 8];   // [sym][tile][q][r], sym 3 (padding) = 1
    const int lane = threadIdx.x & 31, g = lane >> 2, q = lane & 3, wib = threadIdx.x >> 5;
    const int theta = blockIdx.y;
    const double* Tg = a.T + (size_t)theta * K * K;
    const double* Eg = a.E + (size_t)theta * K * a.S;
    const double* pig = a.pi + (size_t)theta * K;
    for (int x = threadIdx.x; x < 4 * NT * 8; x += blockDim.x) {
        const int r = x & 1, qq = (x >> 1) & 3, t = (x >> 3) % NT, sym = (x >> 3) / NT;
        const int stt = C::state_of(t, 2 * qq + r);
        esm[x] = (stt >= 0 && sym < a.S) ? Eg[stt * a.S + sym] : (stt >= 0 ? 1.0 : 0.0);
    }
    // B fragments: lane holds row kq = lane&3 (input state) and column n = lane>>2 (output state)
    double Bf[KS][NT];
#pragma unroll
    for (int sl = 0; sl < KS; ++sl) {
        const int sin = C::state_of(sl >> 1, 2 * q + (sl & 1));
#pragma unroll
        for (int t = 0; t < NT; ++t) {
            const int sout = C::state_of(t, g);
            Bf[sl][t] = (sin >= 0 && sout >= 0) ? Tg[sin * K + sout] : 0.0;
        }
    }
    __syncthreads();

    const int tile0 = (blockIdx.x * (blockDim.x >> 5) + wib) * MT;     // first M-tile of this warp
    const uint32_t* wp[MT];
    int nw[MT], len[MT];
    bool valid[MT];
    int maxnw = 0;
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
        int s = (tile0 + mt) * 8 + g;
        valid[mt] = s < a.nstreams;
        if (!valid[mt]) s = a.nstreams - 1;
        const StreamInfo si = a.streams[s];
        wp[mt] = a.words + si.base;
        nw[mt] = si.nwords;
        len[mt] = si.len;
        maxnw = max(maxnw, nw[mt]);
    }
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) maxnw = max(maxnw, __shfl_xor_sync(0xffffffffu, maxnw, m));
    if (tile0 * 8 >= a.nstreams) return;   // whole warp past the end (warp-uniform)

    DmmaState<K, MT> st;
    int scale[MT];
    bool dead[MT];
    double result[MT];
    uint32_t word[MT];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
        scale[mt] = 0; dead[mt] = false; result[mt] = 0.0;
        word[mt] = wp[mt][0];
        const int o = word[mt] & 3u;
#pragma unroll
        for (int t = 0; t < NT; ++t)
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const int stt = C::state_of(t, 2 * q + r);
                st.c[mt][t][r] = stt >= 0 ? pig[stt] * esm[((o * NT + t) * 4 + q) * 2 + r] : 0.0;
            }
    }
    for (int w = 0; w < maxnw; ++w) {
        uint32_t next[MT];
        int hi[MT];
        bool full = true;
        const int lo = (w == 0) ? 1 : 0;
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
            next[mt] = wp[mt][(long long)min(w + 1, nw[mt] - 1) * 32];
            hi[mt] = max(0, min(16, len[mt] - 16 * w));
            full = full && (lo == 0) && (hi[mt] == 16);
        }
        if (__all_sync(0xffffffffu, full)) {
            uint32_t ww[MT];
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) ww[mt] = word[mt];
#pragma unroll 2
            for (int t = 0; t < 16; ++t) {
                int o[MT];
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) { o[mt] = ww[mt] & 3u; ww[mt] >>= 2; }
                dmma_step<K, MT>(st, Bf, esm, o, q);
            }
        } else {
            uint32_t ww[MT];
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) ww[mt] = word[mt] >> (2 * lo);
#pragma unroll 1
            for (int t = lo; t < 16; ++t) {
                int o[MT];
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) { o[mt] = ww[mt] & 3u; ww[mt] >>= 2; }
                DmmaState<K, MT> old = st;
                dmma_step<K, MT>(st, Bf, esm, o, q);
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) {
                    const bool act = t < hi[mt];
#pragma unroll
                    for (int tt = 0; tt < NT; ++tt) {
                        st.c[mt][tt][0] = act ? st.c[mt][tt][0] : old.c[mt][tt][0];
                        st.c[mt][tt][1] = act ? st.c[mt][tt][1] : old.c[mt][tt][1];
                    }
                }
            }
        }
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
            double sum = 0.0;
#pragma unroll
            for (int t = 0; t < NT; ++t) sum += st.c[mt][t][0] + st.c[mt][t][1];
            sum += shfl_xor_f64(sum, 1);
            sum += shfl_xor_f64(sum, 2);
            if (sum > 0.0) {
                const int e = exponent_of(sum);
                const double f = pow2_neg(e);
#pragma unroll
                for (int t = 0; t < NT; ++t) { st.c[mt][t][0] *= f; st.c[mt][t][1] *= f; }
                scale[mt] += e;
                sum *= f;
            } else if (w < nw[mt]) {
                dead[mt] = true;
                if (sum != sum) result[mt] = sum;
            }
            if (w == nw[mt] - 1)
                result[mt] = dead[mt] ? (result[mt] != result[mt] ? result[mt] : -INFINITY) : finish_logl(sum, scale[mt]);
            word[mt] = next[mt];
        }
    }
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
        const int s = (tile0 + mt) * 8 + g;
        if (valid[mt] && q == 0) a.chain_out[(size_t)theta * a.nstreams + s] = result[mt];
    }
}

// ------------------------------------------------------------------------------------------------
// out[n] = sum_s chain_out[n][s] with a fixed-shape tree (bitwise reproducible run to run)
// ------------------------------------------------------------------------------------------------
__global__ void reduce_chains_kernel(const double* chain_out, int nstreams, double* out) {
    __shared__ double sh[256];
    const int n = blockIdx.x, tid = threadIdx.x;
    const double* src = chain_out + (size_t)n * nstreams;
    double acc = 0.0;
    for (int s = tid; s < nstreams; s += 256) acc += src[s];
    sh[tid] = acc;
    __syncthreads();
    for (int m = 128; m >= 1; m >>= 1) {
        if (tid < m) sh[tid] += sh[tid + m];
        __syncthreads();
    }
    if (tid == 0) out[n] = sh[0];
}

// ------------------------------------------------------------------------------------------------
// The same reduction fused with the all-reduce over ranks (one process per GPU, chunks sharded): every block
// pushes its point's partial sum straight into the mailbox of EVERY rank with peer-to-peer stores over NVLink; the last
// block of the grid to deliver raises this rank's flag in all mailboxes, waits for the other ranks' flags and adds the
// rows up in rank order (bitwise the same result on every rank).  Mailbox of a rank: double val[2][nranks][nmax] (two
// epochs, so that a rank that runs ahead never overwrites rows a slower rank is still reading) followed by one
// 128-byte flag line per source rank holding the epoch it has delivered.  No NCCL call, no extra launch.
// ------------------------------------------------------------------------------------------------
constexpr int MAX_PEERS = 32;
struct PeerReduceArgs {
    int nranks, rank, nmax;
    unsigned epoch;                 // 1, 2, ... one per collective call, the same sequence on every rank
    double* box[MAX_PEERS];         // val array of rank r's mailbox as mapped into this process
    unsigned* flag[MAX_PEERS];      // flag lines of rank r's mailbox: flag[r][32 * source]
    unsigned* counter;              // local: blocks of this launch that have delivered (left at 0)
    unsigned* status;               // host-visible word: 0, or 0x80000000 | (rank that did not deliver in time)
    unsigned long long timeout_ns;  // how long to wait for a peer's flag before giving up
};

__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

__global__ void __launch_bounds__(256) reduce_chains_peer_kernel(const double* chain_out, int nstreams, double* out, PeerReduceArgs pa) {
    __shared__ double sh[256];
    __shared__ int s_last;
    const int n = blockIdx.x, tid = threadIdx.x, N = gridDim.x;
    const double* src = chain_out + (size_t)n * nstreams;
    double acc = 0.0;
    for (int s = tid; s < nstreams; s += 256) acc += src[s];
    sh[tid] = acc;
    __syncthreads();
    for (int m = 128; m >= 1; m >>= 1) {
        if (tid < m) sh[tid] += sh[tid + m];
        __syncthreads();
    }
    const size_t row = ((size_t)(pa.epoch & 1u) * pa.nranks) * pa.nmax;
    if (tid == 0) {
        const double v = sh[0];
        for (int r = 0; r < pa.nranks; ++r)
            *((volatile double*)(pa.box[r] + row + (size_t)pa.rank * pa.nmax + n)) = v;
        __threadfence_system();                                   // the rows are out before this block counts as delivered
        s_last = atomicAdd(pa.counter, 1u) == (unsigned)(N - 1);
    }
    __syncthreads();
    if (!s_last) return;
    if (tid == 0) {
        __threadfence_system();                                   // every block's rows (seen through the counter) before the flag
        for (int r = 0; r < pa.nranks; ++r)
            if (r != pa.rank) *((volatile unsigned*)(pa.flag[r] + 32 * pa.rank)) = pa.epoch;
    }
    // Every thread waits for every source itself -- but not for ever: a peer that died, or that issued a different
    // sequence of calls, would otherwise hang every GPU of the job.  On a timeout the results are NaN and the status word
    // (host-visible) names the rank; the host turns that into IMC_ERR_CUDA and marks the communicator broken.
    bool timed_out = false;
    const unsigned long long t0 = global_timer_ns();
    for (int r = 0; r < pa.nranks && !timed_out; ++r) {
        if (r == pa.rank) continue;
        const volatile unsigned* f = pa.flag[pa.rank] + 32 * r;
        unsigned spins = 0;
        while ((int)(*f - pa.epoch) < 0) {
            __nanosleep(64);
            if ((++spins & 1023u) == 0u && global_timer_ns() - t0 > pa.timeout_ns) {
                timed_out = true;
                if (tid == 0) *((volatile unsigned*)pa.status) = 0x80000000u | (unsigned)r;
                break;
            }
        }
    }
    if (__syncthreads_or(timed_out ? 1 : 0)) {
        for (int i = tid; i < N; i += 256) out[i] = __longlong_as_double(0x7ff8000000000000LL);
        if (tid == 0) { *pa.counter = 0u; __threadfence_system(); }
        return;
    }
    __threadfence_system();
    const double* mine = pa.box[pa.rank] + row;
    for (int i = tid; i < N; i += 256) {
        double s = 0.0;
        for (int r = 0; r < pa.nranks; ++r) s += *((const volatile double*)(mine + (size_t)r * pa.nmax + i));
        out[i] = s;
    }
    if (tid == 0) *pa.counter = 0u;
}

// ------------------------------------------------------------------------------------------------
// FP64 peak probes (the roofline denominator is measured in the same run: MEASURED_PEAKS.json has no FP64 entry)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(512) peak_dfma_kernel(double* out, int iters, double a, double b) {
    double acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = threadIdx.x * 1e-9 + k;
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
            for (int k = 0; k < 8; ++k) acc[k] = fma(acc[k], a, b);
    double s = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += acc[k];
    if (s == 12345.678) out[0] = s;
}

__global__ void __launch_bounds__(512) peak_dmma_kernel(double* out, int iters) {
    double d[4][2];
    const double a = 1e-3 * threadIdx.x, b = 1e-3 * (threadIdx.x * 3 + 1);
#pragma unroll
    for (int k = 0; k < 4; ++k) { d[k][0] = k; d[k][1] = k + 0.5; }
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
            for (int k = 0; k < 4; ++k) dmma884(d[k][0], d[k][1], a, b);
    double s = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) s += d[k][0] + d[k][1];
    if (s == 12345.678) out[0] = s;
}

}  // namespace imc
