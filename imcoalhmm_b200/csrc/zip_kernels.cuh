// Compressed ("zip") forward kernel for sm_100a: the GPU form of ziphmm.zip_forward
// (/root/reference/src/IMCoalHMM/hmm.py:20-21) over the token streams produced by tokenizer.inl.
//
//   C_s      = diag(E[:,s]) T^T                      base symbol matrices (alpha' = C_s alpha)
//   C_(a,b)  = C_b C_a                               dictionary entry for "a then b", built level by level
//   alpha_0  = pi o E[:,o_0];  alpha <- C_tok alpha  one mat-vec per token;  logL = log sum(alpha) + exponents
//
// Persistent CTAs; a CTA serves one parameter point at a time.  zip_build_dictionary puts the point's dictionary (M
// matrices, each scaled by an exact power of two so that its largest entry is in [1,2), exponent kept aside) into shared
// memory; the warps then claim warp-loads of chunks of that point (work stealing across CTAs once a CTA's own points are
// done) and run the chains.  One chain's mat-vec is spread over G lanes (three decompositions below: 8, 4 or 32 lanes),
// each lane reading its rows with LDS.128 from a layout that is bank-conflict free no matter which matrices the chains
// of a warp are using.  The new state is exchanged through a small shared-memory buffer (one STS.64 per owned row, K/2
// broadcast LDS.128 back), after which every lane holds all of alpha again.  The bound is the shared-memory pipe: 8 K^2
// bytes of matrix per chain-step against 128 B/clk/SM, i.e. at most 1/4 of the FP64 rate -- times the compression ratio
// (50-150x on the benchmark alignments).  Launches with few work units per warp walk every chunk in ordered pieces that
// hand their state on through global memory (ZipArgs::nseg, bit-identical), chain-scarce calls use one warp per chain or
// the segmented (scan) form below.  DESIGN.md section 4.1 has the measurements behind every choice made here.
#pragma once
#include "forward_kernels.cuh"
#include <type_traits>
#include <cassert>

// -DIMC_DEBUG_BOUNDS: device-side checks of every data-dependent index (token ids, power-table rows, chunk and list
// indices).  compute-sanitizer is closed on the development pool, so this build + tools/sanitize_case.py is the memory-safety
// evidence (profiles/r02_bounds_checked_run.txt).
#ifdef IMC_DEBUG_BOUNDS
#define IMC_ASSERT(x) assert(x)
#else
#define IMC_ASSERT(x) ((void)0)
#endif

namespace imc {

// Run tokens of the spectral form (tokenizer.inl): word = id | n << 8, n = a + (b << RUN_LO_BITS) sites of the run symbol
// after entry id; Lambda^n = ptab[a] o ptab[RUN_LO_ROWS + b] (two rows of the per-point power table).
constexpr int RUN_LO_BITS = 6, RUN_HI_BITS = 6;
constexpr int RUN_LO_ROWS = 1 << RUN_LO_BITS, RUN_LO_MASK = RUN_LO_ROWS - 1;
constexpr int RUN_ROWS = RUN_LO_ROWS + (1 << RUN_HI_BITS);
constexpr int RUN_MAX = (1 << (RUN_LO_BITS + RUN_HI_BITS)) - 1;
constexpr int RUN_HI_MASK = (1 << RUN_HI_BITS) - 1;
constexpr uint32_t RUN_NOP_BIT = 1u << 21;         // word bit: no token of this chain in this warp-step (aligned form, zip_align_quad)
constexpr uint32_t RUN_TABLE2_BIT = 1u << 20;      // word bit: the run of this token is of the SECOND run symbol (two-run form)

struct ZipChunk {
    long long tok_off;   // byte offset of the chunk's first token (16-byte aligned, 16 readable bytes past the end)
    int ntok;            // tokens (the chunk's symbols 1..L-1 after compression)
    int first_sym;       // >= 0: symbol at position 0 (alpha_0 = pi o E[:,first_sym]);  < 0: the chain starts from the
                         // unit vector e_c, c = -1 - first_sym (column c of a segment's transfer matrix, see zip_fold_kernel)
    int out_index;       // column of chain_out this chunk writes
    int first_run;       // spectral form: sites of the run symbol between position 0 and the first token (<= RUN_MAX)
    int run_sites;       // spectral form: first_run + the runs of all tokens; the result gets run_sites * ln(lambda_max)
    int run2_sites;      // two-run form: sites in the spectrally treated runs of the second run symbol (x ln(lambda2_max))
    int pad;
    int continues;       // > 0: this is a later part of a chunk cut across ranks -- no start of its own (position 0 is an ordinary
                         // site, all of its segments run from unit vectors); see zip_fold_parts_kernel
};

struct ZipArgs {
    const uint8_t* tokens;
    const ZipChunk* chunks;      // sorted by ntok, descending
    int nchunks;
    int* point_next;             // [N] warp-loads of chunks already claimed per parameter point (zeroed before the launch)
    const uint8_t* pairs;        // [M][2] (left, right) in level order; entries < S unused
    const int* level_start;      // [nlevels + 1]
    int nlevels;
    int M;
    int N, K, S;                 // K = actual number of states (<= the tile size the kernel was instantiated for)
    const double* pi;            // [N][K]
    const double* T;             // [N][K][K]
    const double* E;             // [N][K][S]
    double* chain_out;           // [N][out_stride]  log-likelihood per chain (vec_out == NULL)
    int out_stride;
    int active_warps;            // warps per CTA that claim work (scarce work is spread over the SMs, one chain group per warp)
    double* vec_out;             // segmented mode: [N][nchunks][vec_stride] final vector (K doubles, unnormalised) + exponent
    int vec_stride;
    // pipelined mode (nseg > 1): every chunk is walked in nseg pieces of seglen tokens that are separate work units; a
    // piece starts from the state its predecessor left in `carry` once `progress` says so.  Same arithmetic, same bits,
    // but work units nseg times smaller, so the SMs run dry together at the end of a launch with few units per warp.
    int nseg, seglen;
    double* carry;               // [N][nchunks][carry_stride]: state registers of the chain's lanes, exponent, flags
    int carry_stride;
    int* progress;               // [N][nchunks] pieces completed (zeroed before the launch)
    // Subset of the points served by this launch (NULL: all N).  The spectral form serves the points whose C_r could be
    // diagonalised, the plain form the others; both lists are written by zip_spectral_kernel, so the split never
    // touches the host.  point_next is indexed by position in the list, everything else by the point itself.
    const int* plist;
    const int* pcount;
    // spectral form (SPEC kernels): tokens are 32-bit run words (tokenizer.inl), per point spec_stride doubles:
    // lambda[K], wsum[K], b0[S][K], R[S][K][K], ln(lambda_max)  (see zip_spectral_kernel)
    const double* spec;
    int spec_stride;
    int hot_id;                  // MMA form: the most frequent entry of the streams (its matrix lives in registers)
    int nbase;                   // spectral form: base entries read from `spec` (S, or S + 2 in the two-run form: + B^-1, B)
    int run2;                    // two-run form: a second power table (of the second run symbol's eigenvalues) is in use
    int sched;                        // MMA form: aligned streams (one entry per warp-step, no-op words; zip_align_quad in zip_host.inl)
    unsigned long long* mma_passes;   // MMA form: passes (of KT x NT DMMAs each) executed, for the roofline (one atomic per work unit)
};

// Segmented mode (chain-scarce calls: few chunks x few points).  A long chunk is cut into segments of `seglen` tokens.
// Segment 0 runs as usual from pi; segment s > 0 is run K times from the unit vectors e_0..e_{K-1}, which yields the
// columns of its transfer matrix P_s = C_tok[last] ... C_tok[first] (K times the arithmetic, but K * #segments times
// the parallelism).  zip_fold_kernel then folds alpha <- P_s alpha over the segments of each chunk.

// ------------------------------------------------------------------------------------------------
// Lane decompositions.  A config class C describes how one chain's mat-vec is spread over G lanes:
//   C::K, C::KP (state registers per lane), C::G, C::CPW = 32 / G chains per warp,
//   C::STRIDE_D doubles per dictionary matrix, C::off / C::store (matrix layout), C::GS (exchange buffer stride),
//   C::Lane (per-lane constants), C::step (one token), C::state_of (which state register k of a lane holds).
// Costs measured on B200 (tools/microbench/smem_patterns.cu): LDS.128 with 512 distinct bytes 4.2 clk, sparse or
// quarter-broadcast LDS.128 ~3 clk, STS.64 2 clk when bank-disjoint per half-warp.
// ------------------------------------------------------------------------------------------------

// G = 8: lane q owns rows q, q+8, ...; every LDS.128 of a quarter-warp reads 8 consecutive units of ONE matrix.
template <int K_>
struct ZipCfg8 {
    static constexpr int K = K_;
    static constexpr bool MMA = false;
    __host__ __device__ static constexpr int slot(int k) { return k; }      // where state k sits in the per-chain arrays (b0, wsum, power table)
    static constexpr int G = 8, CPW = 4;
    static constexpr int KP = (K + 1) & ~1;            // columns padded to an even count (16-byte units)
    static constexpr int PT = KP;                      // row stride of the power table (spectral form)
    static constexpr int CP = KP / 2;                  // units per row
    static constexpr int RPL = (K + 7) / 8;            // rows per lane
    static constexpr int STRIDE_D = RPL * CP * 16;     // doubles per dictionary matrix
    // Exchange buffer per chain: 2 x KP doubles (+2 of skew room), stride = 64 (mod 128) bytes, chains 2,3 of a warp
    // skewed by 16 bytes: in 16-byte bank groups the four chains then start at 0,4,1,5 (mod 8), so the two chains of
    // each half-warp store their 64-byte row blocks into disjoint bank halves (STS.64: 2 wavefronts instead of 4)
    // AND the quarter-broadcast read-back touches four different bank groups.
    // Large K: a single buffer and a second warp barrier per step instead -- the shared memory saved buys one more
    // dictionary matrix (K = 40: 17 instead of 16 entries, ~5 % fewer tokens), and a barrier is nothing against a step.
    static constexpr int NBUF = K >= 32 ? 1 : 2;
    static constexpr int GS = ((NBUF * KP * 8 + 63) / 128 * 128 + 64) / 8;
    static constexpr int SBUF_PER_WARP = 4 * GS + 2;   // doubles: four chains + the 16-byte skew of chains 2,3
    static constexpr int UNROLL = K <= 12 ? 4 : 1;
    static constexpr int FULL = K / 8;                 // slots in which all 8 lanes own a row
    static constexpr int REM = K % 8;                  // rows of the last, partial slot
    // The partial slot is stored REP times side by side (lane positions f*REM .. f*REM+REM-1 hold rows 8*FULL ..):
    // chain slot g of a warp lets lanes of copy g % REP own the remainder rows, so that the four quarter-warps
    // of one LDS.128 touch different bank groups instead of all hitting groups 0..REM-1.
    static constexpr int REP = REM ? 8 / REM : 1;
    // Two remainder rows (the K = 10 tile): their REP = 4 copies let ONE full load fetch the remainder rows of FOUR
    // consecutive tokens -- lane pair j reads copy j of the matrix of token j of a 4-token word -- instead of one
    // partly filled load per token (5 full loads per 4 tokens in place of 20 partial ones: -19 % shared-memory wavefronts).
    // Pair j then adds the two rows at step j from registers.  Used on the fast path (all chains of the warp inside a
    // 16-token block); the tail falls back to per-token partial loads.
#ifdef IMC_ZIP_NO_GROUP4
    static constexpr bool GROUP4 = false;
#else
    static constexpr bool GROUP4 = (REM == 2 && FULL >= 1);
#endif
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
    struct Lane {
        int q, grp, rem_row;
        double* sb0;
        __device__ __forceinline__ Lane(int lane, int warp, double* sbuf) {
            q = lane & 7; grp = lane >> 3;
            rem_row = -1;
            if (REM) {
                const int p = q - (grp % REP) * REM;
                if (p >= 0 && p < REM) rem_row = 8 * FULL + p;
            }
            sb0 = sbuf + (size_t)warp * SBUF_PER_WARP + grp * GS + (grp >> 1) * 2;
        }
        __device__ __forceinline__ bool writer() const { return q == 0; }
    };
    __device__ static __forceinline__ int state_of(const Lane&, int k) { return k; }
    // four tokens on the fast path of the GROUP4 shape; w[b] = id | run << 8 (run == 0 in the plain form)
    template <bool SPEC>
    __device__ static __forceinline__ void word4(double (&al)[KP], const double* dict, const long long* dexp, const double* ptab,
                                                 const uint32_t (&w)[4], const Lane& L, int& buf, long long& scale) {
        const int cj = L.q >> 1, pr = L.q & 1;          // this lane serves token cj of the four, remainder row pr
        const uint32_t myw = cj == 0 ? w[0] : (cj == 1 ? w[1] : (cj == 2 ? w[2] : w[3]));
        const int myid = myw & 0xffu;
        IMC_ASSERT(!SPEC || (myw >> 20) == 0u);
        const double2* rp = reinterpret_cast<const double2*>(dict + (size_t)myid * STRIDE_D) + FULL * CP * 8 + L.q;
        double2 rm[CP];
#pragma unroll
        for (int cp = 0; cp < CP; ++cp) rm[cp] = rp[cp * 8];
        scale += dexp[myid];       // every lane pair books the exponent of ITS token; zip_run_unit adds the four pairs up
        double frem = 1.0;
        if (SPEC) {
            const int ra = (myw >> 8) & RUN_LO_MASK, rb = RUN_LO_ROWS + ((myw >> (8 + RUN_LO_BITS)) & RUN_HI_MASK);
            frem = ptab[ra * PT + 8 * FULL + pr] * ptab[rb * PT + 8 * FULL + pr];
        }
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int id = w[b] & 0xffu;
            double* sb = L.sb0 + buf * KP;
            const double2* mp = reinterpret_cast<const double2*>(dict + (size_t)id * STRIDE_D) + L.q;
            const double* pa = ptab + ((w[b] >> 8) & RUN_LO_MASK) * PT + L.q;
            const double* pb = ptab + (RUN_LO_ROWS + ((w[b] >> (8 + RUN_LO_BITS)) & RUN_HI_MASK)) * PT + L.q;
#pragma unroll
            for (int k = 0; k < FULL; ++k) {
                double s0 = 0.0, s1 = 0.0;
#pragma unroll
                for (int cp = 0; cp < CP; ++cp) {
                    const double2 m = mp[(k * CP + cp) * 8];
                    s0 = fma(m.x, al[2 * cp], s0);
                    s1 = fma(m.y, al[2 * cp + 1], s1);
                }
                sb[L.q + 8 * k] = SPEC ? (s0 + s1) * (pa[8 * k] * pb[8 * k]) : s0 + s1;
            }
            if (cj == b) {
                double s0 = 0.0, s1 = 0.0;
#pragma unroll
                for (int cp = 0; cp < CP; ++cp) {
                    s0 = fma(rm[cp].x, al[2 * cp], s0);
                    s1 = fma(rm[cp].y, al[2 * cp + 1], s1);
                }
                sb[8 * FULL + pr] = SPEC ? (s0 + s1) * frem : s0 + s1;
            }
            __syncwarp();
#pragma unroll
            for (int cp = 0; cp < CP; ++cp) {
                const double2 v = reinterpret_cast<const double2*>(sb)[cp];
                al[2 * cp] = v.x;
                al[2 * cp + 1] = v.y;
            }
            buf ^= 1;
        }
    }
    template <bool PRED, bool SPEC>
    __device__ static __forceinline__ void step(double (&al)[KP], const double* dict, const long long* dexp, const double* ptab,
                                                uint32_t w, const Lane& L, int buf, long long& scale, bool active) {
        double* sb = L.sb0 + (NBUF == 2 ? buf * KP : 0);
        if (!PRED || active) {
            const int id = w & 0xffu;
            const int ra = (w >> 8) & RUN_LO_MASK, rb = RUN_LO_ROWS + ((w >> (8 + RUN_LO_BITS)) & RUN_HI_MASK);
            IMC_ASSERT((w >> 20) == 0u || !SPEC);
            const double* pa = ptab + ra * PT;
            const double* pb = ptab + rb * PT;
            const double2* mp = reinterpret_cast<const double2*>(dict + (size_t)id * STRIDE_D) + L.q;
#pragma unroll
            for (int k = 0; k < RPL; ++k) {
                if (k < FULL || L.rem_row >= 0) {     // rem_row: row of the partial slot owned by this lane, or -1
                    double s0 = 0.0, s1 = 0.0;
#pragma unroll
                    for (int cp = 0; cp < CP; ++cp) {
                        const double2 m = mp[(k * CP + cp) * 8];
                        s0 = fma(m.x, al[2 * cp], s0);
                        s1 = fma(m.y, al[2 * cp + 1], s1);
                    }
                    const int row = k < FULL ? L.q + 8 * k : L.rem_row;
                    sb[row] = SPEC ? (s0 + s1) * (pa[row] * pb[row]) : s0 + s1;
                }
            }
            if (!GROUP4 || L.q < 2) {                       // GROUP4: token exponents are per-pair partial sums
                scale += dexp[id];
            }
        }
        __syncwarp();
        if (!PRED || active) {
#pragma unroll
            for (int cp = 0; cp < CP; ++cp) {
                const double2 v = reinterpret_cast<const double2*>(sb)[cp];
                al[2 * cp] = v.x;
                al[2 * cp + 1] = v.y;
            }
        }
        if (NBUF == 1) __syncwarp();       // everybody has read the buffer before the next step overwrites it
    }
};

// G = 4: the state count is padded to a multiple of 4, lane g owns rows g, g+4, ...  A quarter-warp holds TWO chains
// (slots c = 0, 1) that use different matrices, so the layout splits every 128-byte line into an even half (bank
// groups 0-3) and an odd half (groups 4-7): line l holds pieces 2l and 2l+1 of each lane's row data (piece p =
// row g + 4*(p / CP), column pair p % CP).  Slot 0 reads piece t at instruction t, slot 1 reads piece t^1, so the
// two chains of a quarter-warp always sit in opposite halves -- conflict free whatever the two matrices are.
// CP is even, so t^1 is the neighbouring column pair of the same row: slot-1 lanes simply keep their state registers
// with neighbouring pairs swapped (they read the exchange buffer at cp^1).  Half as many lanes per chain halves the
// cost of the state exchange per chain-step and leaves no partially filled loads.
template <int K_>
struct ZipCfg4 {
    static constexpr int K = K_;
    static constexpr bool MMA = false;
    __host__ __device__ static constexpr int slot(int k) { return k; }      // where state k sits in the per-chain arrays (b0, wsum, power table)
    static constexpr int G = 4, CPW = 8;
    static constexpr bool GROUP4 = false;
    static constexpr int KP = (K + 3) & ~3;            // states padded to a multiple of 4
    static constexpr int PT = KP;
    static constexpr int CP = KP / 2;                  // column pairs per row (even)
    static constexpr int RPL = KP / 4;                 // rows per lane
    static constexpr int STRIDE_D = KP * KP;           // dense
    // exchange buffer per chain: 2 x KP doubles, stride = 32 (mod 128) bytes: the four chains of a half-warp store
    // their 32-byte row blocks into four different bank quarters (STS.64 in 2 wavefronts)
    static constexpr int NBUF = K >= 16 ? 1 : 2;        // larger K: single buffer + a second warp barrier (more dictionary)
    static constexpr int GS = ((NBUF * KP * 8 + 95) / 128 * 128 + 32) / 8;
    static constexpr int SBUF_PER_WARP = 8 * GS;
    static constexpr int UNROLL = K <= 12 ? 4 : 1;
    __host__ __device__ static constexpr int off(int r, int c) {
        const int p = (r >> 2) * CP + (c >> 1);
        return (((p >> 1) * 8 + (p & 1) * 4 + (r & 3)) * 2) + (c & 1);
    }
    __device__ static __forceinline__ void store(double* D, int r, int c, double v) { D[off(r, c)] = v; }
    struct Lane {
        int g, c, grp;
        int off_even, off_odd;     // byte offsets of this lane's even / odd pieces inside a matrix
        double* sb0;
        __device__ __forceinline__ Lane(int lane, int warp, double* sbuf) {
            g = lane & 3; grp = lane >> 2; c = grp & 1;
            off_even = (c * 4 + g) * 16;
            off_odd = ((1 - c) * 4 + g) * 16;
            sb0 = sbuf + (size_t)(warp * 8 + grp) * GS;
        }
        __device__ __forceinline__ bool writer() const { return g == 0; }
    };
    // register k of a slot-c lane holds state 2*((k/2)^c) + k%2
    __device__ static __forceinline__ int state_of(const Lane& L, int k) { return 2 * ((k >> 1) ^ L.c) + (k & 1); }
    template <bool PRED, bool SPEC>
    __device__ static __forceinline__ void step(double (&al)[KP], const double* dict, const long long* dexp, const double* ptab,
                                                uint32_t w, const Lane& L, int buf, long long& scale, bool active) {
        double* sb = L.sb0 + (NBUF == 2 ? buf * KP : 0);
        if (!PRED || active) {
            const int id = w & 0xffu;
            const int ra = (w >> 8) & RUN_LO_MASK, rb = RUN_LO_ROWS + ((w >> (8 + RUN_LO_BITS)) & RUN_HI_MASK);
            IMC_ASSERT((w >> 20) == 0u || !SPEC);
            const double* pa = ptab + ra * PT + L.g;
            const double* pb = ptab + rb * PT + L.g;
            const char* mb = reinterpret_cast<const char*>(dict + (size_t)id * STRIDE_D);
            const char* me = mb + L.off_even;
            const char* mo = mb + L.off_odd;
#pragma unroll
            for (int k = 0; k < RPL; ++k) {
                double s0 = 0.0, s1 = 0.0;
#pragma unroll
                for (int cp = 0; cp < CP; ++cp) {
                    const int t = k * CP + cp;
                    const double2 m = *reinterpret_cast<const double2*>(((t & 1) ? mo : me) + (t >> 1) * 128);
                    s0 = fma(m.x, al[2 * cp], s0);
                    s1 = fma(m.y, al[2 * cp + 1], s1);
                }
                sb[L.g + 4 * k] = SPEC ? (s0 + s1) * (pa[4 * k] * pb[4 * k]) : s0 + s1;
            }
            scale += dexp[id];
        }
        __syncwarp();
        if (!PRED || active) {
#pragma unroll
            for (int cp = 0; cp < CP; ++cp) {
                const double2 v = reinterpret_cast<const double2*>(sb)[cp ^ L.c];
                al[2 * cp] = v.x;
                al[2 * cp + 1] = v.y;
            }
        }
        if (NBUF == 1) __syncwarp();       // everybody has read the buffer before the next step overwrites it
    }
};

// G = 32: one chain per warp, lane q owns rows q and q+32.  Dense layout [column pair][row]: the lanes of a warp read
// consecutive 16-byte units, so loads are conflict free; with K rows only K lanes are busy, which wastes issue slots
// but makes ONE chain-step as short as it can be (K=20: 10 LDS.128 + 20 DFMA per lane instead of 50 + 100 with 4
// lanes).  Used for chain-scarce calls (a single theta on few chunks), where latency per step is what counts.
template <int K_>
struct ZipCfg32 {
    static constexpr int K = K_;
    static constexpr bool MMA = false;
    __host__ __device__ static constexpr int slot(int k) { return k; }      // where state k sits in the per-chain arrays (b0, wsum, power table)
    static constexpr int G = 32, CPW = 1;
    static constexpr bool GROUP4 = false;
    static constexpr int KP = (K + 1) & ~1;
    static constexpr int PT = KP;
    static constexpr int CP = KP / 2;
    static constexpr int RPL = (K + 31) / 32;
    static constexpr int STRIDE_D = CP * K * 2;        // dense: K units of 16 bytes per column pair
    static constexpr int GS = 2 * KP + 2;              // one chain per warp: no cross-chain bank concerns
    static constexpr int SBUF_PER_WARP = GS;
    static constexpr int UNROLL = K <= 12 ? 4 : 1;
    __host__ __device__ static constexpr int off(int r, int c) { return ((c >> 1) * K + r) * 2 + (c & 1); }
    __device__ static __forceinline__ void store(double* D, int r, int c, double v) { D[off(r, c)] = v; }
    struct Lane {
        int q, grp;
        double* sb0;
        __device__ __forceinline__ Lane(int lane, int warp, double* sbuf) {
            q = lane; grp = 0;
            sb0 = sbuf + (size_t)warp * GS;
        }
        __device__ __forceinline__ bool writer() const { return q == 0; }
    };
    __device__ static __forceinline__ int state_of(const Lane&, int k) { return k; }
    template <bool PRED, bool SPEC>
    __device__ static __forceinline__ void step(double (&al)[KP], const double* dict, const long long* dexp, const double* ptab,
                                                uint32_t w, const Lane& L, int buf, long long& scale, bool active) {
        double* sb = L.sb0 + buf * KP;
        if (!PRED || active) {
            const int id = w & 0xffu;
            const int ra = (w >> 8) & RUN_LO_MASK, rb = RUN_LO_ROWS + ((w >> (8 + RUN_LO_BITS)) & RUN_HI_MASK);
            IMC_ASSERT((w >> 20) == 0u || !SPEC);
            const double* pa = ptab + ra * PT + L.q;
            const double* pb = ptab + rb * PT + L.q;
            const double2* mp = reinterpret_cast<const double2*>(dict + (size_t)id * STRIDE_D) + L.q;
#pragma unroll
            for (int k = 0; k < RPL; ++k) {
                if (L.q + 32 * k < K) {
                    double s0 = 0.0, s1 = 0.0;
#pragma unroll
                    for (int cp = 0; cp < CP; ++cp) {
                        const double2 m = mp[cp * K + 32 * k];
                        s0 = fma(m.x, al[2 * cp], s0);
                        s1 = fma(m.y, al[2 * cp + 1], s1);
                    }
                    sb[L.q + 32 * k] = SPEC ? (s0 + s1) * (pa[32 * k] * pb[32 * k]) : s0 + s1;
                }
            }
            scale += dexp[id];
        }
        __syncwarp();
        if (!PRED || active) {
#pragma unroll
            for (int cp = 0; cp < CP; ++cp) {
                const double2 v = reinterpret_cast<const double2*>(sb)[cp];
                al[2 * cp] = v.x;
                al[2 * cp + 1] = v.y;
            }
        }
    }
};

// MMA form (spectral form only): the 8 chains of a warp are the ROWS of mma.sync.m8n8k4.f64 tiles, alpha' = alpha M^T,
// so chains that apply the SAME matrix share it as the B operand.  In run-token streams one entry -- the lone
// mismatch "1" followed by its run -- is 75-90 % of all tokens: its matrix (the "hot" entry) is held as B fragments in
// REGISTERS (KT x NT doubles per lane), and a step of the warp is KT x NT DMMAs with no shared-memory traffic for the
// matrix and no state exchange at all: the D fragment of one step is the A fragment of the next under the state
// order below (lane q = l % 4 of chain l / 4 holds states 8t + 4e + q in D[t][e], which is column q of k-tile 2t + e).
// Chains whose token is another entry are served by extra passes with B fragments read from shared memory (entries
// are stored in fragment order: one conflict-free LDS.64 per tile), one pass per distinct cold entry in the warp.
// The bound moves from the shared-memory pipe to the FP64 tensor pipe.
template <int K_>
struct ZipCfgM {
    static constexpr int K = K_;
    static constexpr bool MMA = true;
    static constexpr int G = 4, CPW = 8;
    static constexpr bool GROUP4 = false;
    static constexpr int KT = (K + 3) / 4, NT = (K + 7) / 8;     // k-tiles of 4 input states, n-tiles of 8 output states
    static constexpr int KP = NT * 8;                            // slots per chain, lane order: t*8 + q*2 + e
    // power-table rows start at alternating halves of a 128-byte line, so that the two chains of a quarter-warp (64 bytes
    // each per load) collide only when their rows have the same parity
    static constexpr int PT = (KP % 16 == 0) ? KP + 8 : KP;
    static constexpr int STRIDE_D = KT * NT * 32;                // doubles per dictionary entry (fragment order)
    static constexpr int SBUF_PER_WARP = 0;
    static constexpr int UNROLL = 1;
    __host__ __device__ static constexpr int slot(int s) { return (s >> 3) * 8 + (s & 3) * 2 + ((s >> 2) & 1); }
    // element (row r = output state, column c = input state): tile (u = c/4, t = r/8), lane = n-column * 4 + k-row
    __host__ __device__ static constexpr int off(int r, int c) {
        return ((c >> 2) * NT + (r >> 3)) * 32 + (2 * (r & 3) + ((r >> 2) & 1)) * 4 + (c & 3);
    }
    __device__ static __forceinline__ void store(double* D, int r, int c, double v) { D[off(r, c)] = v; }
    struct Lane {
        int q, grp;
        __device__ __forceinline__ Lane(int lane, int, double*) { q = lane & 3; grp = lane >> 2; }
        __device__ __forceinline__ bool writer() const { return q == 0; }
    };
};

template <class C, bool SPEC>
struct ZipSmem {
    // doubles in front of the exchange buffers: plain form E[K][S], spectral form the start vectors b0[S][KP]
    __host__ __device__ static constexpr int se_doubles(int S) { return ((SPEC ? C::KP : C::K) * S + 1) & ~1; }   // keeps what follows 16-byte aligned
    __host__ __device__ static constexpr int tab_doubles() { return SPEC ? RUN_ROWS * C::PT : 0; }
    static size_t bytes(int M, int S, int threads, bool run2 = false) {
        size_t d = (size_t)M * C::STRIDE_D + (size_t)se_doubles(S) + C::KP + (size_t)(threads / 32) * C::SBUF_PER_WARP + tab_doubles() * (run2 ? 2 : 1);
        return d * sizeof(double) + (size_t)M * sizeof(long long) + 4 * sizeof(int);   // dexp[M], s_point[2] + s_best
    }
    static int max_entries(size_t budget, int S, int threads, bool run2 = false) {
        const size_t fixed = bytes(0, S, threads, run2);
        if (budget <= fixed) return 0;
        const size_t m = (budget - fixed) / (C::STRIDE_D * sizeof(double) + sizeof(long long));
        return (int)(m > 256 ? 256 : m);
    }
};

// Take the binary exponent of the state out of it (exactly).  Plain form: the entries are probabilities, their sum
// measures the state.  Spectral form: the entries are coordinates in the eigenbasis of C_r and may be negative; the
// largest magnitude measures the state, the sum of magnitudes only serves to notice NaN / infinity.
template <class C, bool SPEC>
__device__ __forceinline__ void zip_rescale(double (&al)[C::KP], long long& scale, bool& dead, bool& isnan) {
    double sum = 0.0, mx = 0.0;
#pragma unroll
    for (int k = 0; k < C::KP; ++k) {       // padding registers hold 0
        if (SPEC) { const double v = fabs(al[k]); sum += v; mx = fmax(mx, v); }
        else sum += al[k];
    }
    if (sum > 0.0 && sum < 1.7e308) {
        const int e = exponent_of(SPEC ? mx : sum);
        const double f = pow2_neg(e);
#pragma unroll
        for (int k = 0; k < C::KP; ++k) al[k] *= f;
        scale += e;
    } else {
        dead = true;
        isnan = isnan || (sum != sum) || (sum > 0.0);
    }
}

// Build the dictionary of parameter point n in shared memory (all threads of the CTA).
//   plain form:     base entries C_s[r][c] = E[r][s] T[c][r];  sE = E, spi = pi
//   spectral form:  base entries R_s = V^-1 C_s V from zip_spectral_kernel;  sE = start vectors b0[s] = V^-1 (pi o E[:,s]),
//                   spi = wsum = V^T 1 (so that sum(alpha) = wsum . beta), and the power table
//                   ptab[a]      = (lambda / lambda_max)^a        (a < 2^RUN_LO_BITS)
//                   ptab[LO + b] = (lambda / lambda_max)^(b << RUN_LO_BITS)
template <class C, int THREADS, bool SPEC>
__device__ __forceinline__ void zip_build_dictionary(const ZipArgs& a, int n, double* dict, double* sE, double* spi, long long* dexp,
                                                     double* ptab) {
    constexpr int KP = C::KP, NW = THREADS / 32;
    const int K = a.K;          // actual state count <= C::K (the kernel's tile); rows / columns beyond it stay zero
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int M = a.M, S = a.S;
    const double* Tg = a.T + (size_t)n * K * K;
    const double* Eg = a.E + (size_t)n * K * S;
    const double* pig = a.pi + (size_t)n * K;
    const double* sp = SPEC ? a.spec + (size_t)n * a.spec_stride : nullptr;     // lambda[K], wsum[K], b0[S][K], R[S][K][K]
    for (int x = tid; x < M * C::STRIDE_D; x += THREADS) dict[x] = 0.0;
    if (SPEC) {
        for (int x = tid; x < S * KP; x += THREADS) sE[x] = 0.0;
        for (int x = tid; x < KP; x += THREADS) spi[x] = 0.0;
        for (int x = tid; x < RUN_ROWS * C::PT; x += THREADS) ptab[x] = 0.0;
        __syncthreads();
        for (int x = tid; x < S * K; x += THREADS) { const int s = x / K, k = x - s * K; sE[s * KP + C::slot(k)] = sp[2 * K + x]; }
        for (int x = tid; x < K; x += THREADS) spi[C::slot(x)] = sp[K + x];
        // power table: rows hold rho_k^p with rho_k = lambda_k / lambda_max (|rho| <= 1, the dominant entry exactly 1); the
        // common factor lambda_max^(all run sites of the chunk) is added to the result analytically (ZipChunk::run_sites)
        double lmax = 0.0;
        for (int k = 0; k < K; ++k) lmax = fmax(lmax, fabs(sp[k]));
        for (int x = tid; x < RUN_ROWS * K; x += THREADS) {
            const int row = x / K, k = x - row * K;
            const double p = row < RUN_LO_ROWS ? (double)row : (double)((row - RUN_LO_ROWS) << RUN_LO_BITS);
            ptab[row * C::PT + C::slot(k)] = p == 0.0 ? 1.0 : pow(sp[k] / lmax, p);
        }
        if (a.run2) {          // the same table for the eigenvalues of the second run symbol, right behind the first
            const double* l2 = sp + 2 * K + S * K + (size_t)a.nbase * K * K;
            double* ptab2 = ptab + RUN_ROWS * C::PT;
            double l2max = 0.0;
            for (int k = 0; k < K; ++k) l2max = fmax(l2max, fabs(l2[k]));
            for (int x = tid; x < RUN_ROWS * C::PT; x += THREADS) ptab2[x] = 0.0;
            __syncthreads();
            for (int x = tid; x < RUN_ROWS * K; x += THREADS) {
                const int row = x / K, k = x - row * K;
                const double p = row < RUN_LO_ROWS ? (double)row : (double)((row - RUN_LO_ROWS) << RUN_LO_BITS);
                ptab2[row * C::PT + C::slot(k)] = p == 0.0 ? 1.0 : pow(l2[k] / l2max, p);
            }
        }
    } else {
        for (int x = tid; x < K * S; x += THREADS) sE[x] = Eg[x];
        for (int x = tid; x < KP; x += THREADS) spi[x] = x < K ? pig[x] : 0.0;
    }
    __syncthreads();
    for (int lv = -1; lv < a.nlevels; ++lv) {
        const int lo = lv < 0 ? 0 : a.level_start[lv], hi = lv < 0 ? (SPEC ? a.nbase : S) : a.level_start[lv + 1];
        for (int e = lo + warp; e < hi; e += NW) {
            double* D = dict + (size_t)e * C::STRIDE_D;
            double mx = 0.0;
            bool bad = false;
            long long ebase = 0;
            if (lv < 0) {     // plain: C_s[r][c] = E[r][s] * T[c][r];  spectral: R_s[r][c]
                const double* Rg = SPEC ? sp + 2 * K + S * K + (size_t)e * K * K : nullptr;
                for (int x = lane; x < K * K; x += 32) {
                    const int r = x % K, c = x / K;      // rows fastest: consecutive lanes touch consecutive 16-byte units
                    const double v = SPEC ? Rg[r * K + c] : sE[r * S + e] * Tg[c * K + r];
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
                    const int r = x % K, c = x / K;      // B[r][k]: conflict-free across lanes; A[k][c]: (near) broadcast
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
                for (int x = lane; x < K * K; x += 32) C::store(D, x % K, x / K, D[C::off(x % K, x / K)] * f);
            }
            if (lane == 0) dexp[e] = ebase + ex;
        }
        __syncthreads();
    }
}

// One warp-load of chains (C::CPW of them, one per lane group) of parameter point n: the chunks
// unit*CPW .. unit*CPW + CPW-1 of the sorted chunk list.
template <class C, bool SPEC>
__device__ __forceinline__ void zip_run_unit(const ZipArgs& a, int n, int unit, const double* dict, const double* sE,
                                             const double* spi, const long long* dexp, const double* ptab,
                                             const typename C::Lane& L) {
    constexpr int KP = C::KP;
    constexpr int BLK = SPEC ? 8 : 16;        // tokens per block (two / one 16-byte loads); the state is rescaled every 8 tokens
    const int K = a.K, S = a.S;
    int seg = 0, quad = unit;
    if (a.nseg > 1) {          // pipelined mode: units are numbered piece-major, so a chunk's pieces are claimed in order
        const int nquads = (a.nchunks + C::CPW - 1) / C::CPW;
        seg = unit / nquads;
        quad = unit - seg * nquads;
    }
    const int ci = quad * C::CPW + L.grp;
    const bool have = ci < a.nchunks;
    IMC_ASSERT(n >= 0 && n < a.N && quad * C::CPW < a.nchunks);
    const ZipChunk ch = a.chunks[have ? ci : quad * C::CPW];
    const int tok0 = seg * a.seglen;
    const int nt = have ? (a.nseg > 1 ? max(0, min(ch.ntok - tok0, a.seglen)) : ch.ntok) : 0;
    IMC_ASSERT(ch.ntok >= 0 && ch.first_sym < a.S && ch.first_sym >= -a.K && ch.first_run <= RUN_MAX);
    const uint4* tp = reinterpret_cast<const uint4*>(a.tokens + ch.tok_off + (size_t)tok0 * (SPEC ? 4 : 1));
    int maxnt = nt;
#pragma unroll
    for (int m = 16; m >= C::G; m >>= 1) maxnt = max(maxnt, __shfl_xor_sync(0xffffffffu, maxnt, m));

    double al[KP];
    long long scale = 0;       // exponents taken out by zip_rescale (the same in every lane of the chain)
    long long tscale = 0;      // exponents of the dictionary matrices applied (GROUP4: partial sum per lane pair)
    bool dead = false, isnan = false;
    double* carry = a.carry + ((size_t)n * a.nchunks + (have ? ci : 0)) * a.carry_stride;
    int* prog = a.progress + (size_t)n * a.nchunks + (have ? ci : 0);
    if (seg == 0) {
#pragma unroll
        for (int k = 0; k < KP; ++k) {
            const int st = C::state_of(L, k);
            if (ch.first_sym >= 0) {
                if (SPEC) al[k] = sE[ch.first_sym * KP + st];
                else al[k] = st < K ? spi[st] * sE[st * S + ch.first_sym] : 0.0;
            } else al[k] = st == -1 - ch.first_sym ? 1.0 : 0.0;
        }
        if (SPEC && ch.first_run > 0) {       // run-symbol sites between position 0 and the first token
            const int ra = ch.first_run & RUN_LO_MASK, rb = RUN_LO_ROWS + (ch.first_run >> RUN_LO_BITS);
#pragma unroll
            for (int k = 0; k < KP; ++k) {
                const int st = C::state_of(L, k);
                al[k] *= ptab[ra * C::PT + st] * ptab[rb * C::PT + st];
            }
        }
    } else if (have) {         // every lane of the chain waits for the predecessor piece itself, then reads what it left
        while (*((volatile int*)prog) < seg) __nanosleep(200);
        __threadfence();
#pragma unroll
        for (int k = 0; k < KP; ++k) al[k] = __ldcg(carry + k);
        scale = (long long)__ldcg(carry + KP);
        const int fl = (int)__ldcg(carry + KP + 1);
        dead = fl & 1;
        isnan = fl & 2;
    } else {
#pragma unroll
        for (int k = 0; k < KP; ++k) al[k] = 0.0;
    }
    int buf = 0;
    uint4 cur = make_uint4(0, 0, 0, 0), cur2 = cur;
    if (nt > 0) { cur = tp[0]; if (SPEC) cur2 = tp[1]; }        // 16 readable bytes of slack behind every stream
    for (int blk = 0; blk * BLK < maxnt; ++blk) {
        uint4 nxt = make_uint4(0, 0, 0, 0), nxt2 = nxt;
        if ((blk + 1) * BLK < nt) {
            if (SPEC) { nxt = tp[2 * blk + 2]; nxt2 = tp[2 * blk + 3]; }
            else nxt = tp[blk + 1];
        }
        const int rem = nt - blk * BLK;
        if (SPEC) {
            const uint32_t w[8] = {cur.x, cur.y, cur.z, cur.w, cur2.x, cur2.y, cur2.z, cur2.w};
            if (__all_sync(0xffffffffu, rem >= BLK)) {
                if constexpr (C::GROUP4) {
                    const uint32_t w0[4] = {w[0], w[1], w[2], w[3]}, w1[4] = {w[4], w[5], w[6], w[7]};
                    C::template word4<true>(al, dict, dexp, ptab, w0, L, buf, tscale);
                    C::template word4<true>(al, dict, dexp, ptab, w1, L, buf, tscale);
                } else {
#pragma unroll C::UNROLL
                    for (int b = 0; b < 8; ++b) {
                        C::template step<false, true>(al, dict, dexp, ptab, w[b], L, buf, tscale, true);
                        buf ^= 1;
                    }
                }
            } else {
#pragma unroll 1
                for (int b = 0; b < 8; ++b) {
                    C::template step<true, true>(al, dict, dexp, ptab, w[b], L, buf, tscale, b < rem);
                    buf ^= 1;
                }
            }
            zip_rescale<C, true>(al, scale, dead, isnan);
        } else {
            const uint32_t w[4] = {cur.x, cur.y, cur.z, cur.w};
            if (__all_sync(0xffffffffu, rem >= 16)) {
#pragma unroll
                for (int wi = 0; wi < 4; ++wi) {
                    uint32_t wv = w[wi];
                    if constexpr (C::GROUP4) {
                        const uint32_t w4[4] = {wv & 0xffu, (wv >> 8) & 0xffu, (wv >> 16) & 0xffu, wv >> 24};
                        C::template word4<false>(al, dict, dexp, ptab, w4, L, buf, tscale);
                    } else {
#pragma unroll C::UNROLL
                        for (int b = 0; b < 4; ++b) {
                            IMC_ASSERT((int)(wv & 0xffu) < a.M);
                            C::template step<false, false>(al, dict, dexp, ptab, wv & 0xffu, L, buf, tscale, true);
                            wv >>= 8;
                            buf ^= 1;
                        }
                    }
                    if (wi & 1) zip_rescale<C, false>(al, scale, dead, isnan);
                }
            } else {
#pragma unroll
                for (int wi = 0; wi < 4; ++wi) {
                    uint32_t wv = w[wi];
#pragma unroll 1
                    for (int b = 0; b < 4; ++b) {
                        C::template step<true, false>(al, dict, dexp, ptab, wv & 0xffu, L, buf, tscale, wi * 4 + b < rem);
                        wv >>= 8;
                        buf ^= 1;
                    }
                    if (wi & 1) zip_rescale<C, false>(al, scale, dead, isnan);
                }
            }
        }
        cur = nxt;
        cur2 = nxt2;
    }
    if constexpr (C::GROUP4) {      // add the four lane pairs' partial sums (lanes q and q^1 hold the same value)
        tscale += __shfl_xor_sync(0xffffffffu, tscale, 2);
        tscale += __shfl_xor_sync(0xffffffffu, tscale, 4);
    }
    scale += tscale;
    if (a.nseg > 1 && seg < a.nseg - 1) {       // hand the state on to the next piece (the same lane group of some warp)
        if (have && L.writer()) {
#pragma unroll
            for (int k = 0; k < KP; ++k) __stcg(carry + k, al[k]);
            __stcg(carry + KP, (double)scale);
            __stcg(carry + KP + 1, (double)((dead ? 1 : 0) | (isnan ? 2 : 0)));
            __threadfence();
            *((volatile int*)prog) = seg + 1;
        }
        return;
    }
    double sum = 0.0, mag = 0.0;
#pragma unroll
    for (int k = 0; k < KP; ++k) {
        if (SPEC) { sum = fma(spi[C::state_of(L, k)], al[k], sum); mag += fabs(al[k]); }      // sum(alpha) = wsum . beta
        else sum += al[k];
    }
    const bool bad = isnan || sum != sum || (SPEC && mag != mag);
    if (a.vec_out) {       // segmented mode: hand the vector on (a dead chain hands on zeros; a unit-vector chain may
        if (have && L.writer()) {                 // legitimately end at zero, which is not an error by itself)
            double* out = a.vec_out + ((size_t)n * a.nchunks + ch.out_index) * a.vec_stride;
#pragma unroll
            for (int k = 0; k < KP; ++k) {
                const int st = C::state_of(L, k);
                if (st < K) out[st] = bad ? __longlong_as_double(0x7ff8000000000000LL) : (dead ? 0.0 : al[k]);
            }
            out[K] = (double)scale;
        }
        return;
    }
    double result;
    if (dead || !(sum > 0.0)) result = bad ? __longlong_as_double(0x7ff8000000000000LL) : -INFINITY;
    else {
        result = log(sum) + (double)scale * LN2;
        if (SPEC) result += (double)ch.run_sites * __ldg(a.spec + (size_t)n * a.spec_stride + a.spec_stride - 1);     // ln(lambda_max) per run site
    }
    if (have && L.writer()) a.chain_out[(size_t)n * a.out_stride + ch.out_index] = result;
}

__device__ __forceinline__ double2 lds_f64x2(uint32_t addr) {
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ double lds_f64(uint32_t addr) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ long long lds_s64(uint32_t addr) {
    long long v;
    asm volatile("ld.shared.s64 %0, [%1];" : "=l"(v) : "r"(addr));
    return v;
}

// The same work unit in the MMA form (ZipCfgM, spectral form only): 8 chains per warp, state in the D / A fragments.
template <class C, bool SCHED>
__device__ __forceinline__ void zip_run_unit_mma(const ZipArgs& a, int n, int unit, const double* dict, const double* sE,
                                                 const double* spi, const long long* dexp, const double* ptab,
                                                 const double (&Bh)[C::KT][C::NT], const typename C::Lane& L) {
    constexpr int KP = C::KP, KT = C::KT, NT = C::NT, PT = C::PT;
    const int K = a.K, lane = threadIdx.x & 31, hot = a.hot_id;
    int seg = 0, quad = unit;
    if (a.nseg > 1) {
        const int nquads = (a.nchunks + C::CPW - 1) / C::CPW;
        seg = unit / nquads;
        quad = unit - seg * nquads;
    }
    const int ci = quad * C::CPW + L.grp;
    const bool have = ci < a.nchunks;
    IMC_ASSERT(n >= 0 && n < a.N && quad * C::CPW < a.nchunks);
    const ZipChunk ch = a.chunks[have ? ci : quad * C::CPW];
    const int tok0 = seg * a.seglen;
    // (aligned form: lanes without a chain follow the words of the quad's first chain, as no-ops, to know every step's entry)
    const int nt = (have || SCHED) ? (a.nseg > 1 ? max(0, min(ch.ntok - tok0, a.seglen)) : ch.ntok) : 0;
    IMC_ASSERT(ch.ntok >= 0 && ch.first_sym < a.S && ch.first_sym >= -a.K && ch.first_run <= RUN_MAX && hot < a.M);
    const uint4* tp = reinterpret_cast<const uint4*>(a.tokens + ch.tok_off + (size_t)tok0 * 4);
    int maxnt = nt;
#pragma unroll
    for (int m = 16; m >= C::G; m >>= 1) maxnt = max(maxnt, __shfl_xor_sync(0xffffffffu, maxnt, m));

    double D[NT][2];           // D[t][e]: state 8t + 4e + q of this lane's chain (slot t*8 + 2q + e of the per-chain arrays)
    long long scale = 0;
    bool dead = false, isnan = false;
    double* carry = a.carry + ((size_t)n * a.nchunks + (have ? ci : 0)) * a.carry_stride;
    int* prog = a.progress + (size_t)n * a.nchunks + (have ? ci : 0);
    if (seg == 0) {
#pragma unroll
        for (int t = 0; t < NT; ++t) {
            if (ch.first_sym >= 0) {
                const double2 v = *reinterpret_cast<const double2*>(sE + ch.first_sym * KP + t * 8 + 2 * L.q);
                D[t][0] = v.x; D[t][1] = v.y;
            } else {
                const int c = -1 - ch.first_sym;
                D[t][0] = 8 * t + L.q == c ? 1.0 : 0.0;
                D[t][1] = 8 * t + 4 + L.q == c ? 1.0 : 0.0;
            }
        }
        if (ch.first_run > 0) {
            const int ra = ch.first_run & RUN_LO_MASK, rb = RUN_LO_ROWS + (ch.first_run >> RUN_LO_BITS);
#pragma unroll
            for (int t = 0; t < NT; ++t) {
                const double2 fa = *reinterpret_cast<const double2*>(ptab + ra * PT + t * 8 + 2 * L.q);
                const double2 fb = *reinterpret_cast<const double2*>(ptab + rb * PT + t * 8 + 2 * L.q);
                D[t][0] *= fa.x * fb.x; D[t][1] *= fa.y * fb.y;
            }
        }
    } else if (have) {
        while (*((volatile int*)prog) < seg) __nanosleep(200);
        __threadfence();
#pragma unroll
        for (int t = 0; t < NT; ++t) { D[t][0] = __ldcg(carry + t * 8 + 2 * L.q); D[t][1] = __ldcg(carry + t * 8 + 2 * L.q + 1); }
        scale = (long long)__ldcg(carry + KP);
        const int fl = (int)__ldcg(carry + KP + 1);
        dead = fl & 1;
        isnan = fl & 2;
    } else {
#pragma unroll
        for (int t = 0; t < NT; ++t) { D[t][0] = 0.0; D[t][1] = 0.0; }
    }
    // shared-memory addresses of what a step reads, as 32-bit registers the compiler cannot rematerialise from the kernel
    // arguments every step (it did: ~25 integer instructions per step in an issue-bound loop)
    uint32_t ptab_s = (uint32_t)__cvta_generic_to_shared(ptab) + 16u * L.q;      // this lane's column pair of a table row
    uint32_t dict_s = (uint32_t)__cvta_generic_to_shared(dict) + 8u * lane;      // this lane's element of a fragment tile
    uint32_t dexp_s = (uint32_t)__cvta_generic_to_shared(dexp);
    asm volatile("" : "+r"(ptab_s), "+r"(dict_s), "+r"(dexp_s));
    unsigned passes = 0;
    // one token of every chain of the warp; ALL: every chain has one (no predicates)
    auto step = [&](uint32_t wb, bool active, auto all_tag) {
        constexpr bool ALL = decltype(all_tag)::value;
        if (!ALL && !active) wb = 0u;
        const int id = wb & 0xffu;
        IMC_ASSERT(id < a.M && (wb >> 22) == 0u && (a.run2 || !(wb & RUN_TABLE2_BIT)));
        const uint32_t tab = ptab_s + ((wb >> 20) & 1u) * (RUN_ROWS * PT * 8);      // the first or the second run symbol's table
        const uint32_t pa = tab + ((wb >> 8) & RUN_LO_MASK) * (PT * 8), pb = tab + (RUN_LO_ROWS + ((wb >> (8 + RUN_LO_BITS)) & RUN_HI_MASK)) * (PT * 8);
        double2 fa[NT], fb[NT];            // (lambda / lambda_max)^n of this token for the lane's states: independent of the products below
#pragma unroll
        for (int t = 0; t < NT; ++t) { fa[t] = lds_f64x2(pa + t * 64); fb[t] = lds_f64x2(pb + t * 64); }
        const long long ex = lds_s64(dexp_s + id * 8);
        const bool is_cold = (ALL || active) && id != hot;
        unsigned cold = __ballot_sync(0xffffffffu, is_cold);
        ++passes;
        double N[NT][2];
#pragma unroll
        for (int t = 0; t < NT; ++t) { N[t][0] = 0.0; N[t][1] = 0.0; }
#pragma unroll
        for (int u = 0; u < KT; ++u)       // hot pass, unconditionally: a warp without a single hot chain is rare
#pragma unroll
            for (int t = 0; t < NT; ++t) dmma884(N[t][0], N[t][1], D[u >> 1][u & 1], Bh[u][t]);
        while (cold) {                     // one pass per distinct cold entry among the warp's chains
            const int idc = __shfl_sync(0xffffffffu, id, __ffs(cold) - 1);
            const bool mine = is_cold && id == idc;
            cold &= ~__ballot_sync(0xffffffffu, mine);
            ++passes;
            const uint32_t bc = dict_s + idc * (C::STRIDE_D * 8);
            double Cc[NT][2];
#pragma unroll
            for (int t = 0; t < NT; ++t) { Cc[t][0] = 0.0; Cc[t][1] = 0.0; }
#pragma unroll
            for (int u = 0; u < KT; ++u)
#pragma unroll
                for (int t = 0; t < NT; ++t) dmma884(Cc[t][0], Cc[t][1], D[u >> 1][u & 1], lds_f64(bc + (u * NT + t) * 256));
            if (mine) {
#pragma unroll
                for (int t = 0; t < NT; ++t) { N[t][0] = Cc[t][0]; N[t][1] = Cc[t][1]; }
            }
        }
        if (ALL || active) {
#pragma unroll
            for (int t = 0; t < NT; ++t) { D[t][0] = N[t][0] * (fa[t].x * fb[t].x); D[t][1] = N[t][1] * (fa[t].y * fb[t].y); }
            scale += ex;
        }
    };
    // aligned form: ONE entry per warp-step, the same for every chain that takes part; the others hold a no-op word that still
    // names the step's entry (so the choice of the B operand needs no vote; the padding of a stream names the hot entry)
    auto step_aligned = [&](uint32_t wb) {
        const int ids = wb & 0xffu;
        const bool active = have && !(wb & RUN_NOP_BIT);
        IMC_ASSERT(ids < a.M && (wb >> 22) == 0u && __all_sync(0xffffffffu, ids == __shfl_sync(0xffffffffu, ids, 0)));
        // (a no-op word carries no run: rows 0 of the tables)
        const uint32_t tab = ptab_s + ((wb >> 20) & 1u) * (RUN_ROWS * PT * 8);
        const uint32_t pa = tab + ((wb >> 8) & RUN_LO_MASK) * (PT * 8), pb = tab + (RUN_LO_ROWS + ((wb >> (8 + RUN_LO_BITS)) & RUN_HI_MASK)) * (PT * 8);
        double f[NT][2];                   // (lambda / lambda_max)^n of this token for the lane's states: independent of the products below
#pragma unroll
        for (int t = 0; t < NT; ++t) {
            const double2 fa = lds_f64x2(pa + t * 64), fb = lds_f64x2(pb + t * 64);
            f[t][0] = fa.x * fb.x; f[t][1] = fa.y * fb.y;
        }
        const long long ex = lds_s64(dexp_s + ids * 8);
        ++passes;
        double N[NT][2];
#pragma unroll
        for (int t = 0; t < NT; ++t) { N[t][0] = 0.0; N[t][1] = 0.0; }
        if (ids == hot) {
#pragma unroll
            for (int u = 0; u < KT; ++u)
#pragma unroll
                for (int t = 0; t < NT; ++t) dmma884(N[t][0], N[t][1], D[u >> 1][u & 1], Bh[u][t]);
        } else {
            const uint32_t bc = dict_s + ids * (C::STRIDE_D * 8);
#pragma unroll
            for (int u = 0; u < KT; ++u)
#pragma unroll
                for (int t = 0; t < NT; ++t) dmma884(N[t][0], N[t][1], D[u >> 1][u & 1], lds_f64(bc + (u * NT + t) * 256));
        }
        // chains that sit this step out keep their state: selects, not a branch (the lanes of a warp differ here, and a divergent
        // region around eight DMULs costs more than the DMULs)
#pragma unroll
        for (int t = 0; t < NT; ++t) {
            const double d0 = N[t][0] * f[t][0], d1 = N[t][1] * f[t][1];
            D[t][0] = active ? d0 : D[t][0];
            D[t][1] = active ? d1 : D[t][1];
        }
        scale += active ? ex : 0ll;
    };
    uint4 cur = make_uint4(0, 0, 0, 0), cur2 = cur;
    if (nt > 0) { cur = tp[0]; cur2 = tp[1]; }
    for (int blk = 0; blk * 8 < maxnt; ++blk) {
        uint4 nxt = make_uint4(0, 0, 0, 0), nxt2 = nxt;
        if ((blk + 1) * 8 < nt) { nxt = tp[2 * blk + 2]; nxt2 = tp[2 * blk + 3]; }
        const int rem = nt - blk * 8;
        const uint32_t w[8] = {cur.x, cur.y, cur.z, cur.w, cur2.x, cur2.y, cur2.z, cur2.w};
        if (SCHED) {          // streams of a quad have one length, a multiple of 8; lanes without a chain hold no-op words throughout
#pragma unroll
            for (int b = 0; b < 8; ++b) step_aligned(w[b]);
        } else if (__all_sync(0xffffffffu, rem >= 8)) {
#pragma unroll
            for (int b = 0; b < 8; ++b) step(w[b], true, std::true_type());
        } else {
#pragma unroll 1
            for (int b = 0; b < 8; ++b) {
                const uint32_t wb = b == 0 ? w[0] : (b == 1 ? w[1] : (b == 2 ? w[2] : (b == 3 ? w[3] : (b == 4 ? w[4] : (b == 5 ? w[5] : (b == 6 ? w[6] : w[7]))))));
                step(wb, b < rem, std::false_type());
            }
        }
        {   // exact power-of-two rescale of every chain (4 lanes each)
            double sum = 0.0, mx = 0.0;
#pragma unroll
            for (int t = 0; t < NT; ++t) {
                const double v0 = fabs(D[t][0]), v1 = fabs(D[t][1]);
                sum += v0 + v1;
                mx = fmax(mx, fmax(v0, v1));
            }
            sum += shfl_xor_f64(sum, 1); sum += shfl_xor_f64(sum, 2);
            mx = fmax(mx, shfl_xor_f64(mx, 1)); mx = fmax(mx, shfl_xor_f64(mx, 2));
            if (sum > 0.0 && sum < 1.7e308) {
                const int e = exponent_of(mx);
                const double f = pow2_neg(e);
#pragma unroll
                for (int t = 0; t < NT; ++t) { D[t][0] *= f; D[t][1] *= f; }
                scale += e;
            } else if (rem > 0) {
                dead = true;
                isnan = isnan || (sum != sum) || (sum > 0.0);
            }
        }
        cur = nxt;
        cur2 = nxt2;
    }
    if (lane == 0 && a.mma_passes) atomicAdd(a.mma_passes, (unsigned long long)passes);
    if (a.nseg > 1 && seg < a.nseg - 1) {       // hand the state on to the next piece
        if (have) {
#pragma unroll
            for (int t = 0; t < NT; ++t) { __stcg(carry + t * 8 + 2 * L.q, D[t][0]); __stcg(carry + t * 8 + 2 * L.q + 1, D[t][1]); }
            if (L.writer()) {
                __stcg(carry + KP, (double)scale);
                __stcg(carry + KP + 1, (double)((dead ? 1 : 0) | (isnan ? 2 : 0)));
            }
            __threadfence();
        }
        __syncwarp();                            // all four lanes of a chain have published their part
        if (have && L.writer()) *((volatile int*)prog) = seg + 1;
        return;
    }
    double sum = 0.0, mag = 0.0;
#pragma unroll
    for (int t = 0; t < NT; ++t) {
        const double2 wv = *reinterpret_cast<const double2*>(spi + t * 8 + 2 * L.q);
        sum = fma(wv.x, D[t][0], fma(wv.y, D[t][1], sum));
        mag += fabs(D[t][0]) + fabs(D[t][1]);
    }
    sum += shfl_xor_f64(sum, 1); sum += shfl_xor_f64(sum, 2);
    mag += shfl_xor_f64(mag, 1); mag += shfl_xor_f64(mag, 2);
    const bool bad = isnan || sum != sum || mag != mag;
    if (a.vec_out) {
        if (have) {
            double* out = a.vec_out + ((size_t)n * a.nchunks + ch.out_index) * a.vec_stride;
#pragma unroll
            for (int t = 0; t < NT; ++t)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int st = 8 * t + 4 * e + L.q;
                    if (st < K) out[st] = bad ? __longlong_as_double(0x7ff8000000000000LL) : (dead ? 0.0 : D[t][e]);
                }
            if (L.writer()) out[K] = (double)scale;
        }
        return;
    }
    double result;
    if (dead || !(sum > 0.0)) result = bad ? __longlong_as_double(0x7ff8000000000000LL) : -INFINITY;
    else result = log(sum) + (double)scale * LN2 + (double)ch.run_sites * __ldg(a.spec + (size_t)n * a.spec_stride + a.spec_stride - 1)
                  + (double)ch.run2_sites * __ldg(a.spec + (size_t)n * a.spec_stride + a.spec_stride - 2);
    if (have && L.writer()) a.chain_out[(size_t)n * a.out_stride + ch.out_index] = result;
}

// Persistent CTAs.  Work unit = (parameter point, warp-load of C::CPW chunks); point_next[i] counts the units of
// the i-th point of the launch already claimed (zeroed before the launch).  A CTA first serves the points blockIdx.x,
// blockIdx.x + gridDim.x, ... (building each point's dictionary once and letting its warps claim units), then helps
// whichever point still has unclaimed units, so that the SMs finish together no matter how points and chunks divide
// among them.
template <class C, int THREADS, int MINB, bool SPEC>
__global__ void __launch_bounds__(THREADS, MINB) zip_forward_kernel(ZipArgs a) {
    constexpr int KP = C::KP;
    extern __shared__ __align__(128) unsigned char zsm_raw[];
    double* dict = reinterpret_cast<double*>(zsm_raw);
    const int M = a.M, S = a.S;
    double* sE = dict + (size_t)M * C::STRIDE_D;      // plain: E[K][S];  spectral: b0[S][KP]
    double* spi = sE + ZipSmem<C, SPEC>::se_doubles(S);     // [KP] pi / wsum
    double* sbuf = spi + KP;                          // [THREADS/32][SBUF_PER_WARP]
    double* ptab = sbuf + (THREADS / 32) * C::SBUF_PER_WARP;            // spectral: [RUN_ROWS][KP]
    long long* dexp = reinterpret_cast<long long*>(ptab + ZipSmem<C, SPEC>::tab_doubles() * (SPEC && a.run2 ? 2 : 1));   // [M] (64-bit: an entry can span millions of sites)
    int* s_point = reinterpret_cast<int*>(dexp + M);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const typename C::Lane L(lane, warp, sbuf);
    const int NP = a.pcount ? *a.pcount : a.N;         // points served by this launch
    if (NP <= 0) return;                               // (the plain-form pass of a spectral call is usually empty)
    for (int x = tid; x < (THREADS / 32) * C::SBUF_PER_WARP; x += THREADS) sbuf[x] = 0.0;   // padding entries stay 0
    const int nunits = (a.nchunks + C::CPW - 1) / C::CPW * (a.nseg > 1 ? a.nseg : 1);
    int primary = blockIdx.x;          // next point of this CTA's own share
    int scan = NP > 0 ? (int)(((long long)blockIdx.x * 7919) % NP) : 0;   // where the search for points to help starts
    unsigned long long* s_best = reinterpret_cast<unsigned long long*>(s_point + 2);
    for (;;) {
        // ---- choose a point: own share first, then the point with the MOST unclaimed warp-loads (ties: the first one
        // after this CTA's own scan start, so that helpers spread out).  s_point[0] = point, -1 = none left anywhere,
        // -2 = own point already finished by helpers.  All decisions go through shared memory so that they are uniform
        // over the CTA.
        __syncthreads();               // everybody is done with the previous point's dictionary and s_point[0]
        if (primary < NP) {
            if (tid == 0) s_point[0] = *((volatile int*)(a.point_next + primary)) < nunits ? primary : -2;
            primary += gridDim.x;
        } else {
            if (tid == 0) { s_point[0] = -1; *s_best = 0ull; }
            __syncthreads();
            for (int idx = tid; idx < NP; idx += THREADS) {
                const int left = nunits - *((volatile int*)(a.point_next + (scan + idx) % NP));
                if (left > 0) atomicMax(s_best, ((unsigned long long)left << 32) | (0xffffffffu - (unsigned)idx));
            }
            __syncthreads();
            if (tid == 0 && *s_best != 0ull) s_point[0] = (scan + (int)(0xffffffffu - (unsigned)(*s_best & 0xffffffffull))) % NP;
        }
        __syncthreads();
        const int slot = s_point[0];
        if (slot == -1) break;
        if (slot == -2) continue;
        const int n = a.plist ? a.plist[slot] : slot;
        IMC_ASSERT(slot >= 0 && slot < NP && n >= 0 && n < a.N);
        zip_build_dictionary<C, THREADS, SPEC>(a, n, dict, sE, spi, dexp, ptab);
        if constexpr (C::MMA) {
            double Bh[C::KT][C::NT];       // the hot entry's B fragments
#pragma unroll
            for (int u = 0; u < C::KT; ++u)
#pragma unroll
                for (int t = 0; t < C::NT; ++t) Bh[u][t] = dict[(size_t)a.hot_id * C::STRIDE_D + (u * C::NT + t) * 32 + lane];
            for (; warp < a.active_warps;) {
                int unit = 0;
                if (lane == 0) unit = atomicAdd(a.point_next + slot, 1);
                unit = __shfl_sync(0xffffffffu, unit, 0);
                if (unit >= nunits) break;
                if (a.sched) zip_run_unit_mma<C, true>(a, n, unit, dict, sE, spi, dexp, ptab, Bh, L);
                else zip_run_unit_mma<C, false>(a, n, unit, dict, sE, spi, dexp, ptab, Bh, L);
            }
        } else {
            for (; warp < a.active_warps;) {
                int unit = 0;
                if (lane == 0) unit = atomicAdd(a.point_next + slot, 1);
                unit = __shfl_sync(0xffffffffu, unit, 0);
                if (unit >= nunits) break;
                zip_run_unit<C, SPEC>(a, n, unit, dict, sE, spi, dexp, ptab, L);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Spectral preparation, one CTA per parameter point.  With r the run symbol, J = diag(pi) T symmetric (every model
// of the reference builds T from a symmetric joint matrix, transitions.py:231-246) and w = sqrt(E[:,r] o pi):
//     C_r = diag(E[:,r]) T^T = W A W^-1,   A[i][j] = sqrt(E[i,r] E[j,r]) J[i][j] / sqrt(pi_i pi_j)   symmetric,
//     A = Q Lambda Q^T  (cyclic Jacobi, K <= 64),   V = W Q,  V^-1 = Q^T W^-1,
//     C_s = diag(E[:,s] / E[:,r]) C_r   =>   R_s = V^-1 C_s V = (Q^T diag(E[:,s]/E[:,r]) Q) Lambda,   R_r = Lambda.
// Written per point (spec_stride doubles): lambda[K], wsum[K] = V^T 1 = Q^T w, b0[S][K] = V^-1 (pi o E[:,s]), R[S][K][K],
// ln(lambda_max) in the last slot.
// A point is served by the spectral kernel only if all of this is sound: pi, E[:,r] > 0, J symmetric to 1e-13, Jacobi
// converged, lambda_max > 0; such points are appended to ok_list, the others to bad_list (plain form of the kernel).
// ------------------------------------------------------------------------------------------------
struct ZipSpecArgs {
    int N, K, S, run_sym;
    int run_sym2;           // two-run form: the second run symbol, or -1
    const double* pi; const double* T; const double* E;
    double* spec; int spec_stride;
    int* ok_list; int* bad_list;
    int* counts;            // [2]: points in ok_list, bad_list (zeroed before the launch)
    int* okflag;            // [N]: 1 where the point went to ok_list
    int force_bad;          // experiment switch: send every point to the plain form
    int point_base;         // index of this launch's first point in the caller's batch (the lists hold batch indices)
};

constexpr int SPEC_THREADS = 128;
static inline int zip_spec_stride(int K, int S, bool run2) { return 2 * K + S * K + (S + (run2 ? 2 : 0)) * K * K + K + 2; }
static inline size_t zip_spec_smem(int K, bool run2) {
    return sizeof(double) * ((size_t)(run2 ? 4 : 2) * K * (K + 1) + 7 * (size_t)K + 16) + sizeof(int) * 4;
}

// Symmetric form A = sqrt(E_i E_j) J_ij / sqrt(pi_i pi_j) of C_sym = diag(E[:,sym]) T^T in shared memory, its checks, and its
// diagonalisation A = Q diag(A) Q^T by cyclic Jacobi.  All threads of the CTA; flag[0] is cleared where the point does not qualify.
__device__ void zip_spec_diagonalise(const ZipSpecArgs& s, const double* pig, const double* Tg, const double* Eg, int sym,
                                     double* A, double* Q, double* w, double* rot, double* red, int* flag) {
    const int K = s.K, S = s.S, LD = K + 1, tid = threadIdx.x;
    if (tid == 0) { flag[1] = 0; red[0] = 0.0; red[1] = 0.0; }
    __syncthreads();
    for (int i = tid; i < K; i += SPEC_THREADS) {
        const double v = pig[i] * Eg[i * S + sym];
        if (!(v > 0.0) || !(v < 1.7e308) || !(pig[i] > 0.0)) flag[0] = 0;
        w[i] = sqrt(v);
    }
    __syncthreads();
    double amax = 0.0, asym = 0.0;
    for (int x = tid; x < K * K; x += SPEC_THREADS) {
        const int i = x / K, j = x - i * K;
        const double gi = Eg[i * S + sym] / w[i], gj = Eg[j * S + sym] / w[j];          // sqrt(E_i / pi_i)
        const double aij = gi * gj * (pig[i] * Tg[i * K + j]), aji = gi * gj * (pig[j] * Tg[j * K + i]);
        A[i * LD + j] = 0.5 * (aij + aji);
        Q[i * LD + j] = i == j ? 1.0 : 0.0;
        amax = fmax(amax, fabs(aij));
        asym = fmax(asym, fabs(aij - aji));
        if (!(fabs(aij) < 1.7e308)) flag[0] = 0;
    }
    for (int m = 16; m >= 1; m >>= 1) { amax = fmax(amax, shfl_xor_f64(amax, m)); asym = fmax(asym, shfl_xor_f64(asym, m)); }
    if ((tid & 31) == 0) {
        atomicMax(reinterpret_cast<unsigned long long*>(red), (unsigned long long)__double_as_longlong(amax));     // non-negative doubles order like integers
        atomicMax(reinterpret_cast<unsigned long long*>(red + 1), (unsigned long long)__double_as_longlong(asym));
    }
    __syncthreads();
    if (tid == 0 && !(red[1] <= 1e-13 * red[0])) flag[0] = 0;
    __syncthreads();
    if (!flag[0]) return;
    // ---- cyclic Jacobi with the round-robin ordering: m - 1 rounds of m / 2 disjoint rotations per sweep.  Warp w takes
    // the pairs w, w + nwarps, ...; lane j the column (row) j, j + 32: no integer division in the loops.
    const int m = (K + 1) & ~1, half = m / 2, nwarps = SPEC_THREADS / 32, warp = tid >> 5, lane = tid & 31;
    const double tiny = 1e-300;
    auto pair_of = [&](int round, int pr, int& p, int& q) {     // (m-1, round) for pr == 0, else ((round + pr), (round - pr)) mod (m-1)
        int a1 = round + pr, a2 = round - pr;
        if (a1 >= m - 1) a1 -= m - 1;
        if (a2 < 0) a2 += m - 1;
        p = pr == 0 ? m - 1 : a1;
        q = pr == 0 ? round : a2;
        if (p > q) { const int t = p; p = q; q = t; }
    };
    for (int sweep = 0; sweep < 30; ++sweep) {
        double off = 0.0;
        for (int i = warp; i < K; i += nwarps)
            for (int j = lane; j < K; j += 32) if (i != j) off = fmax(off, fabs(A[i * LD + j]));
        for (int mm = 16; mm >= 1; mm >>= 1) off = fmax(off, shfl_xor_f64(off, mm));
        if (tid == 0) red[2] = 0.0;
        __syncthreads();
        if (lane == 0) atomicMax(reinterpret_cast<unsigned long long*>(red + 2), (unsigned long long)__double_as_longlong(off));
        __syncthreads();
        // off-diagonal residual r: A = Q Lambda Q^T holds to r, which moves logL by about r x sites (1e-16 x 1e6 sites = 1e-10
        // absolute, 5e-15 of a typical |logL|); rounding noise keeps r near 1e-17 however many sweeps follow
        if (red[2] <= 1e-16 * red[0]) { if (tid == 0) flag[1] = 1; break; }
        for (int round = 0; round < m - 1; ++round) {
            if (tid < half) {
                int p, q;
                pair_of(round, tid, p, q);
                double c = 1.0, sn = 0.0;
                if (q < K) {
                    const double apq = A[p * LD + q], app = A[p * LD + p], aqq = A[q * LD + q];
                    if (fabs(apq) > tiny) {
                        const double th = (aqq - app) / (2.0 * apq);
                        const double t = (th >= 0.0 ? 1.0 : -1.0) / (fabs(th) + sqrt(th * th + 1.0));
                        c = rsqrt(t * t + 1.0);
                        sn = t * c;
                    }
                }
                rot[2 * tid] = c; rot[2 * tid + 1] = sn;
            }
            __syncthreads();
            for (int pr = warp; pr < half; pr += nwarps) {                 // rows: A <- J^T A
                int p, q;
                pair_of(round, pr, p, q);
                if (q >= K) continue;
                const double c = rot[2 * pr], sn = rot[2 * pr + 1];
                if (sn == 0.0) continue;
                for (int j = lane; j < K; j += 32) {
                    const double ap = A[p * LD + j], aq = A[q * LD + j];
                    A[p * LD + j] = c * ap - sn * aq;
                    A[q * LD + j] = sn * ap + c * aq;
                }
            }
            __syncthreads();
            for (int pr = warp; pr < half; pr += nwarps) {                 // columns: A <- A J, Q <- Q J
                int p, q;
                pair_of(round, pr, p, q);
                if (q >= K) continue;
                const double c = rot[2 * pr], sn = rot[2 * pr + 1];
                if (sn == 0.0) continue;
                for (int i = lane; i < K; i += 32) {
                    const double ap = A[i * LD + p], aq = A[i * LD + q];
                    double np = c * ap - sn * aq, nq = sn * ap + c * aq;
                    if (i == p) nq = 0.0;                                  // the annihilated pair, exactly
                    if (i == q) np = 0.0;
                    A[i * LD + p] = np;
                    A[i * LD + q] = nq;
                    const double vp = Q[i * LD + p], vq = Q[i * LD + q];
                    Q[i * LD + p] = c * vp - sn * vq;
                    Q[i * LD + q] = sn * vp + c * vq;
                }
            }
            __syncthreads();
        }
    }
    __syncthreads();
    if (tid == 0) {
        double lmax = 0.0, lpos = 0.0;
        for (int k = 0; k < K; ++k) { lmax = fmax(lmax, fabs(A[k * LD + k])); lpos = fmax(lpos, A[k * LD + k]); }
        if (!flag[1] || !(lmax > 0.0) || !(lpos >= lmax) || !(lmax < 1.7e308)) flag[0] = 0;     // the dominant eigenvalue must be the positive one
        red[3] = lmax;
    }
    __syncthreads();
}

// ------------------------------------------------------------------------------------------------
// Spectral preparation, one CTA per parameter point.  With r the run symbol, J = diag(pi) T symmetric (every model
// of the reference builds T from a symmetric joint matrix, transitions.py:231-246) and w = sqrt(E[:,r] o pi):
//     C_r = diag(E[:,r]) T^T = W A W^-1,   A[i][j] = sqrt(E[i,r] E[j,r]) J[i][j] / sqrt(pi_i pi_j)   symmetric,
//     A = Q Lambda Q^T  (cyclic Jacobi, K <= 64),   V = W Q,  V^-1 = Q^T W^-1,
//     C_s = diag(E[:,s] / E[:,r]) C_r   =>   R_s = V^-1 C_s V = (Q^T diag(E[:,s]/E[:,r]) Q) Lambda,   R_r = Lambda.
// Written per point (spec_stride doubles): lambda[K], wsum[K] = V^T 1 = Q^T w, b0[S][K] = V^-1 (pi o E[:,s]), R[nbase][K][K],
// lambda2[K], ln(lambda2_max), ln(lambda_max) (the last two slots).
// Two-run form (run_sym2 >= 0, nbase = S + 2): the second run symbol r2 (missing data in a pairwise alignment) is
// diagonalised the same way, C_r2 = V2 Lambda2 V2^-1; in the basis of r a run of m sites of r2 is B Lambda2^m B^-1 with
// B = V^-1 V2 = Q^T diag(w2 / w) Q2, so the two extra base entries R[S] = B^-1 and R[S+1] = B turn ANY such run into two
// fixed matrices and a diagonal.
// A point is served by the spectral kernel only if all of this is sound: pi, E[:,r] > 0, J symmetric to 1e-13, Jacobi
// converged, lambda_max > 0; such points are appended to ok_list, the others to bad_list (plain form of the kernel).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(SPEC_THREADS) zip_spectral_kernel(ZipSpecArgs s) {
    extern __shared__ __align__(16) unsigned char spsm_raw[];
    const int K = s.K, S = s.S, LD = K + 1, tid = threadIdx.x, n = blockIdx.x;
    const bool run2 = s.run_sym2 >= 0;
    const int nbase = S + (run2 ? 2 : 0);
    double* A = reinterpret_cast<double*>(spsm_raw);     // [K][LD]
    double* Q = A + (size_t)K * LD;                      // [K][LD]
    double* A2 = Q + (size_t)K * LD;                     // two-run form: the same for the second run symbol
    double* Q2 = A2 + (run2 ? (size_t)K * LD : 0);
    double* w = Q2 + (run2 ? (size_t)K * LD : 0);        // [K]
    double* w2 = w + K;                                  // [K]
    double* rot = w2 + K;                                // [K/2+1][2] cosine, sine of the round's rotations
    double* dsc = rot + 2 * (K / 2 + 1);                 // [K] E[:,s] / E[:,r]
    double* red = dsc + K;                               // [4] reductions
    int* flag = reinterpret_cast<int*>(red + 4 + 2 * K); // [2]
    const double* Tg = s.T + (size_t)n * K * K;
    const double* Eg = s.E + (size_t)n * K * S;
    const double* pig = s.pi + (size_t)n * K;
    double* out = s.spec + (size_t)n * s.spec_stride;
    const int r = s.run_sym;
    if (tid == 0) flag[0] = s.force_bad ? 0 : 1;
    __syncthreads();
    if (flag[0]) zip_spec_diagonalise(s, pig, Tg, Eg, r, A, Q, w, rot, red, flag);
    __syncthreads();
    double* lam2_out = out + 2 * K + (size_t)S * K + (size_t)nbase * K * K;
    if (flag[0]) {
        double* lam = out;
        double* wsum = out + K;
        double* b0 = out + 2 * K;
        double* R = out + 2 * K + (size_t)S * K;
        if (tid == 0) { out[s.spec_stride - 1] = log(red[3]); out[s.spec_stride - 2] = 0.0; }
        for (int k = tid; k < K; k += SPEC_THREADS) {
            lam[k] = A[k * LD + k];
            lam2_out[k] = 1.0;
            double acc = 0.0;
            for (int i = 0; i < K; ++i) acc = fma(Q[i * LD + k], w[i], acc);
            wsum[k] = acc;
        }
        for (int x = tid; x < S * K; x += SPEC_THREADS) {
            const int sy = x / K, k = x - sy * K;
            double acc = 0.0;
            for (int i = 0; i < K; ++i) acc = fma(Q[i * LD + k], pig[i] * Eg[i * S + sy] / w[i], acc);
            b0[x] = acc;
        }
        for (int sy = 0; sy < S; ++sy) {
            __syncthreads();
            for (int i = tid; i < K; i += SPEC_THREADS) dsc[i] = Eg[i * S + sy] / Eg[i * S + r];
            __syncthreads();
            double* Rs = R + (size_t)sy * K * K;
            for (int x = tid; x < K * K; x += SPEC_THREADS) {
                const int i = x / K, j = x - i * K;
                double v;
                if (sy == r) v = i == j ? A[i * LD + i] : 0.0;
                else {
                    double acc = 0.0;
                    for (int k = 0; k < K; ++k) acc = fma(Q[k * LD + i] * dsc[k], Q[k * LD + j], acc);
                    v = acc * A[j * LD + j];
                }
                Rs[x] = v;
            }
        }
    }
    __syncthreads();
    if (flag[0] && run2) {
        zip_spec_diagonalise(s, pig, Tg, Eg, s.run_sym2, A2, Q2, w2, rot, red, flag);
        __syncthreads();
        if (flag[0]) {
            double* Binv = out + 2 * K + (size_t)S * K + (size_t)S * K * K;
            double* B = Binv + (size_t)K * K;
            if (tid == 0) out[s.spec_stride - 2] = log(red[3]);
            for (int k = tid; k < K; k += SPEC_THREADS) lam2_out[k] = A2[k * LD + k];
            for (int x = tid; x < K * K; x += SPEC_THREADS) {
                const int i = x / K, j = x - i * K;
                double b = 0.0, bi = 0.0;
                for (int k = 0; k < K; ++k) {
                    b = fma(Q[k * LD + i] * (w2[k] / w[k]), Q2[k * LD + j], b);        // B = Q^T diag(w2 / w) Q2
                    bi = fma(Q2[k * LD + i] * (w[k] / w2[k]), Q[k * LD + j], bi);      // B^-1 = Q2^T diag(w / w2) Q
                }
                B[x] = b;
                Binv[x] = bi;
            }
        }
    }
    __syncthreads();
    if (tid == 0) {
        if (flag[0]) s.ok_list[atomicAdd(s.counts, 1)] = s.point_base + n;
        else s.bad_list[atomicAdd(s.counts + 1, 1)] = s.point_base + n;
        s.okflag[s.point_base + n] = flag[0] ? 1 : 0;
    }
}

// Fold alpha <- P_s alpha over a run of consecutive segments.  One CTA of 64 threads per (item, parameter point), K <= 64.
// An item starts from vector `start` of its source buffer, folds in `nfold` segments whose column c sits at index
// first_cols + i*K + c, and writes either a log-likelihood (dst == 0: chain_out[n][out_index]) or the folded vector
// (dst == 1: vec2[n][out_index]).  Long chunks are folded in two levels (groups of ~sqrt(#segments) segments in
// parallel, then the groups), so the sequential depth stays ~2 sqrt(#segments) whatever the chunk length.
struct ZipFoldItem {
    int start, first_cols, nfold, out_index;
    int src, dst;        // src: 0 = vec (kernel output), 1 = vec2 (level-1 output)
    int run_sites;       // dst == 0, spectral form: run sites of the whole chunk (see ZipChunk)
    int run2_sites;
};

// Spectral form (spec != NULL): the vectors are coordinates in the eigenbasis of C_r -- entries of either sign, measured
// by their largest magnitude, and sum(alpha) = wsum . beta at the end.  plist / pcount: the points of this launch
// (list_base + blockIdx.y indexes the list), as in zip_forward_kernel; without a list blockIdx.y is the point.
__global__ void __launch_bounds__(64) zip_fold_kernel(const double* vec, int nvec, double* vec2, int nvec2, int vec_stride,
                                                      const ZipFoldItem* items, int K, double* chain_out, int out_stride,
                                                      const double* spec, int spec_stride, const int* plist, const int* pcount, int list_base) {
    __shared__ double w[64];
    __shared__ double red[2];
    __shared__ int s_e[2];
    if (pcount && list_base + (int)blockIdx.y >= *pcount) return;
    const int n = plist ? plist[list_base + blockIdx.y] : blockIdx.y, j = threadIdx.x;
    const ZipFoldItem it = items[blockIdx.x];
    IMC_ASSERT(it.nfold >= 0 && it.start >= 0 && it.first_cols >= 0 && (it.src == 0 ? it.first_cols + it.nfold * K <= nvec : it.first_cols + it.nfold * K <= nvec2));
    const double* base = it.src == 0 ? vec + (size_t)n * nvec * vec_stride : vec2 + (size_t)n * nvec2 * vec_stride;
    const double* v0 = base + (size_t)it.start * vec_stride;
    double alpha = j < K ? v0[j] : 0.0;
    double scale = v0[K];
    for (int s = 0; s < it.nfold; ++s) {
        const double* cols = base + (size_t)(it.first_cols + s * K) * vec_stride;
        // column c carries 2^e_c; bring all columns to the largest exponent among those that matter
        const double ec = j < K ? cols[(size_t)j * vec_stride + K] : 0.0;
        bool matters = j < K && alpha != 0.0;
        if (matters) {                                 // an all-zero column (impossible continuation) carries no scale
            bool any = false;
            for (int r = 0; r < K; ++r) any = any || cols[(size_t)j * vec_stride + r] != 0.0;
            matters = any;
        }
        int e = matters ? (int)ec : -0x40000000;
        for (int m = 16; m >= 1; m >>= 1) e = max(e, __shfl_xor_sync(0xffffffffu, e, m));
        if ((j & 31) == 0) s_e[j >> 5] = e;
        __syncthreads();
        const int emax = max(s_e[0], s_e[1]);
        const int d = (int)ec - emax;                  // <= 0 where it matters
        w[j] = matters ? (d < -1000 ? 0.0 : alpha * pow2_neg(-d)) : 0.0;
        __syncthreads();
        double acc = 0.0;
        if (j < K)
            for (int c = 0; c < K; ++c) acc = fma(w[c], cols[(size_t)c * vec_stride + j], acc);
        double sum = spec ? fabs(acc) : acc;
        for (int m = 16; m >= 1; m >>= 1) sum = spec ? fmax(sum, shfl_xor_f64(sum, m)) : sum + shfl_xor_f64(sum, m);
        __syncthreads();                               // w and s_e are free again
        if ((j & 31) == 0) red[j >> 5] = sum;
        __syncthreads();
        sum = spec ? fmax(red[0], red[1]) : red[0] + red[1];
        int en = 0;
        if (sum > 0.0 && sum < 1.7e308) { en = exponent_of(sum); acc *= pow2_neg(en); }
        alpha = acc;
        scale += (emax > -0x40000000 ? (double)emax : 0.0) + (double)en;
        __syncthreads();
    }
    if (it.dst == 1) {
        double* out = vec2 + ((size_t)n * nvec2 + it.out_index) * vec_stride;
        if (j < K) out[j] = alpha;
        if (j == 0) out[K] = scale;
        return;
    }
    double sum = spec ? (j < K ? spec[(size_t)n * spec_stride + K + j] * alpha : 0.0) : alpha;
    for (int m = 16; m >= 1; m >>= 1) sum += shfl_xor_f64(sum, m);
    if ((j & 31) == 0) red[j >> 5] = sum;
    __syncthreads();
    if (j == 0) {
        sum = red[0] + red[1];
        double r;
        if (sum != sum) r = sum;
        else if (!(sum > 0.0)) r = -INFINITY;
        else r = log(sum) + scale * LN2 + (spec ? (double)it.run_sites * spec[(size_t)n * spec_stride + spec_stride - 1]
                                                   + (double)it.run2_sites * spec[(size_t)n * spec_stride + spec_stride - 2] : 0.0);
        chain_out[(size_t)n * out_stride + it.out_index] = r;
    }
}

// ------------------------------------------------------------------------------------------------
// One long chunk cut into consecutive PARTS that live on different ranks (fewer chunks than GPUs): every part is
// run in segmented mode; part 0 folds its segments into a vector, every later part into the K columns of its transfer
// matrix (zip_fold_kernel, dst == 1).  The per-rank blocks are all-gathered and this kernel folds alpha <- P_part alpha over
// the parts in order, for one parameter point per CTA (64 threads, K <= 64).
//   gathered: [rank][block] with block = double vec[N][n_local * K][K + 1], then run_sites[n_local]
//   part p lives on rank p / n_local as local part p % n_local; its vector c of point n is vec[n][(p % n_local) * K + c]
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(64) zip_fold_parts_kernel(const double* gathered, size_t block_doubles, int N, int K, int n_local,
                                                            int n_parts, const double* spec, int spec_stride, const int* okflag,
                                                            double* out) {
    __shared__ double w[64];
    __shared__ double red[2];
    __shared__ int s_e[2];
    const int n = blockIdx.x, j = threadIdx.x, vs = K + 1;
    const bool sp = spec && (!okflag || okflag[n]);
    auto vec_of = [&](int part, int c) {
        return gathered + (size_t)(part / n_local) * block_doubles + ((size_t)n * n_local * K + (size_t)(part % n_local) * K + c) * vs;
    };
    const double* v0 = vec_of(0, 0);
    double alpha = j < K ? v0[j] : 0.0;
    double scale = v0[K];
    double run_sites = 0.0;
    for (int p = 0; p < n_parts; ++p) run_sites += gathered[(size_t)(p / n_local) * block_doubles + (size_t)N * n_local * K * vs + (p % n_local)];
    for (int p = 1; p < n_parts; ++p) {
        const double* cj = vec_of(p, j < K ? j : 0);
        const double ec = j < K ? cj[K] : 0.0;
        bool matters = j < K && alpha != 0.0;
        if (matters) {
            bool any = false;
            for (int r = 0; r < K; ++r) any = any || cj[r] != 0.0;
            matters = any;
        }
        int e = matters ? (int)ec : -0x40000000;
        for (int m = 16; m >= 1; m >>= 1) e = max(e, __shfl_xor_sync(0xffffffffu, e, m));
        if ((j & 31) == 0) s_e[j >> 5] = e;
        __syncthreads();
        const int emax = max(s_e[0], s_e[1]);
        const int d = (int)ec - emax;
        w[j] = matters ? (d < -1000 ? 0.0 : alpha * pow2_neg(-d)) : 0.0;
        __syncthreads();
        double acc = 0.0;
        if (j < K)
            for (int c = 0; c < K; ++c) acc = fma(w[c], vec_of(p, c)[j], acc);
        double sum = sp ? fabs(acc) : acc;
        for (int m = 16; m >= 1; m >>= 1) sum = sp ? fmax(sum, shfl_xor_f64(sum, m)) : sum + shfl_xor_f64(sum, m);
        __syncthreads();
        if ((j & 31) == 0) red[j >> 5] = sum;
        __syncthreads();
        sum = sp ? fmax(red[0], red[1]) : red[0] + red[1];
        int en = 0;
        if (sum > 0.0 && sum < 1.7e308) { en = exponent_of(sum); acc *= pow2_neg(en); }
        alpha = acc;
        scale += (emax > -0x40000000 ? (double)emax : 0.0) + (double)en;
        __syncthreads();
    }
    double sum = sp ? (j < K ? spec[(size_t)n * spec_stride + K + j] * alpha : 0.0) : alpha;
    for (int m = 16; m >= 1; m >>= 1) sum += shfl_xor_f64(sum, m);
    if ((j & 31) == 0) red[j >> 5] = sum;
    __syncthreads();
    if (j == 0) {
        sum = red[0] + red[1];
        double r;
        if (sum != sum) r = sum;
        else if (!(sum > 0.0)) r = -INFINITY;
        else r = log(sum) + scale * LN2 + (sp ? run_sites * spec[(size_t)n * spec_stride + spec_stride - 1] : 0.0);
        out[n] = r;
    }
}

}  // namespace imc
