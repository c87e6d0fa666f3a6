// Batched model build on the GPU: theta[N][P] -> (pi[N][K], T[N][K][K], E[N][K][3]).
//
// Reference path (all host Python + scipy in the reference):
//   model.py:44-49                       build_hidden_markov_model = ctmc system -> (pi, T) -> E
//   <model>.py build_ctmc_system         rates tables, break points, through / upto / between matrices
//   CTMC.py:12-51                        rate matrix from the edge list; expm(Q * dt)
//   transitions.py:34-76, 204-248        upto / between products; joint matrix J -> pi, T
//   emissions.py:11-100                  truncated-exponential coalescence points, Jukes-Cantor emissions
//   break_points.py:9-30,60-78,81-108    exp / uniform / psmc break points
//
// Three kernels, no host round trip between them and the forward kernels:
//   model_params_kernel  one thread per parameter point: validity, break points, per-interval rates and dt, E
//   model_expm_kernel    one CTA per (parameter point, interval): P = expm(Q dt) by uniformisation
//                        (S = I + Q/q is stochastic, exp(Q dt) = e^-l sum_k l^k/k! S^k, all terms non-negative)
//                        with scaling and squaring; optional 0/1 projection into the next interval's space
//   model_chain_kernel   one CTA per parameter point: u_i = u_{i-1} P_{i-1}, L-block products, J, pi, T
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace imc {

enum { SP_ISO = 0, SP_SINGLE = 1, SP_MIG = 2, N_SPACES = 3 };
enum { MODEL_ISOLATION = 0, MODEL_IM = 1, MODEL_PSMC_ISO = 2, MODEL_VARMIG = 3, MODEL_IM_EPOCHS = 4 };
enum { LBL_C1 = 0, LBL_C2 = 1, LBL_R = 2, LBL_M12 = 3, LBL_M21 = 4, N_LABELS = 5 };
constexpr int MAX_STATES = 94;
constexpr int MAX_L = 16;

struct SpaceDev {
    int n, n_edges;
    const int* edges;      // [n_edges][3] = (src, dst, label)
    int nB, nL, nE;
    const int* B; const int* L; const int* E;
};

struct ModelDev {
    int kind, K, P;
    int n_mig, n_anc, n_epochs, est_split, initial_state;
    SpaceDev space[N_SPACES];
    const int* interval_space;   // [K]   state space used in interval i
    const int* interval_epoch;   // [K]   epoch of interval i (variable-rate models)
    const int* proj_iso_single;  // [4]
    const int* proj_iso_mig;     // [4]
    const int* proj_mig_single;  // [94]
    const double* c_exp_a;       // expon.ppf(i / n) constants for the ancestral / only exp break points
    const double* c_psmc;        // psmc_break_points(K) constants (offset 0)
    const long long* p_off;      // [K]   offset (in doubles) of P_i in the per-point matrix buffer
    long long p_stride;          // doubles per parameter point in the matrix buffer
    int pre_space_n;             // 4 if the model has an isolation pre-phase, 0 otherwise
};

// per parameter point, written by model_params_kernel
struct PointParams {
    // layout in a double array: [0] t0, [1..5] pre-phase rates, then per interval i: 5 rates + dt  -> 6 doubles,
    // then K doubles rep[i]: the first interval whose transition matrix interval i shares (see model_params_kernel)
    static __host__ __device__ int size(int K) { return 6 + 7 * K; }
    static __host__ __device__ int rep_off(int K) { return 6 + 6 * K; }
};

__device__ __forceinline__ double trunc_exp_mid(double t1, double t2, double rate) {   // emissions.py:11-25
    const double d = t2 - t1;
    return t1 + 1.0 / rate - (d * exp(-d * rate)) / (1.0 - exp(-d * rate));
}

__device__ void write_emissions(const double* bp, const double* rates, int K, double* E) {   // emissions.py:44-100
    for (int i = 0; i < K; ++i) {
        const double m = (i + 1 < K) ? trunc_exp_mid(bp[i], bp[i + 1], rates[i]) : bp[K - 1] + 1.0 / rates[K - 1];
        const double x = exp(-4.0 / 3 * (2 * m));
        E[i * 3 + 0] = 0.25 + 0.75 * x;
        E[i * 3 + 1] = 0.75 - 0.75 * x;
        E[i * 3 + 2] = 1.0;
    }
}

// ------------------------------------------------------------------------------------------------
// theta -> per-interval rates, durations, emission matrix.  One thread per parameter point.
// scratch: per point 3*K doubles (break points, emission break points, emission rates)
// ------------------------------------------------------------------------------------------------
__global__ void model_params_kernel(ModelDev m, int N, const double* theta, double* params, double* E,
                                    int* status, double* scratch) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const int K = m.K, P = m.P;
    const double* th = theta + (size_t)n * P;
    double* pp = params + (size_t)n * PointParams::size(K);
    double* bp = scratch + (size_t)n * 3 * K;
    double* ebp = bp + K;
    double* er = ebp + K;
    bool ok = true;
    for (int p = 0; p < P; ++p) ok = ok && (th[p] > 0.0);          // model.py:32-42: all(parameters > 0)
    status[n] = ok ? 0 : 1;
    if (!ok) {   // keep the downstream kernels finite; the result is overwritten with -inf
        for (int x = 0; x < PointParams::size(K); ++x) pp[x] = 0.0;
        for (int i = 0; i < K; ++i) pp[PointParams::rep_off(K) + i] = i;
        for (int x = 0; x < K * 3; ++x) E[(size_t)n * K * 3 + x] = 1.0;
        return;
    }
    double* pre = pp;            // [0] t0, [1..5] rates
    double* iv = pp + 6;         // [i][0..4] rates, [i][5] dt
    for (int x = 0; x < 6 + 6 * K; ++x) pp[x] = 0.0;

    if (m.kind == MODEL_ISOLATION) {                                 // isolation_model.py:106-122
        const double split = th[0], coal = th[1], rec = th[2];
        for (int i = 0; i < K; ++i) { bp[i] = m.c_exp_a[i] / coal + split; er[i] = coal; ebp[i] = bp[i]; }
        pre[0] = bp[0]; pre[1 + LBL_C1] = coal; pre[1 + LBL_C2] = coal; pre[1 + LBL_R] = rec;
        for (int i = 0; i < K; ++i) { iv[i * 6 + LBL_C1] = coal; iv[i * 6 + LBL_R] = rec; }
    } else if (m.kind == MODEL_IM) {                                 // isolation_with_migration_model.py:131-164
        const double tau1 = th[0], tau2 = th[0] + th[1], coal = th[2], rec = th[3], mig = th[4];
        for (int i = 0; i < m.n_mig; ++i) bp[i] = ((double)i / m.n_mig) * (tau2 - tau1) + tau1;
        for (int i = 0; i < m.n_anc; ++i) bp[m.n_mig + i] = m.c_exp_a[i] / coal + tau2;
        for (int i = 0; i < K; ++i) { er[i] = coal; ebp[i] = bp[i]; }
        pre[0] = bp[0]; pre[1 + LBL_C1] = coal; pre[1 + LBL_C2] = coal; pre[1 + LBL_R] = rec;
        for (int i = 0; i < K; ++i) {
            iv[i * 6 + LBL_C1] = coal; iv[i * 6 + LBL_R] = rec;
            if (i < m.n_mig) { iv[i * 6 + LBL_C2] = coal; iv[i * 6 + LBL_M12] = mig; iv[i * 6 + LBL_M21] = mig; }
        }
    } else if (m.kind == MODEL_PSMC_ISO) {                           // variable_coalescence_rate_isolation_model.py:110-178
        const int e = m.n_epochs;
        const double split = m.est_split ? th[0] : 0.0;
        const double* cr = th + (m.est_split ? 1 : 0);
        const double rec = th[P - 1];
        bp[0] = split;
        for (int i = 1; i < K; ++i) bp[i] = split + m.c_psmc[i];
        for (int i = 0; i < K; ++i) { er[i] = cr[m.interval_epoch[i]]; ebp[i] = bp[i]; }
        pre[0] = bp[0]; pre[1 + LBL_C1] = cr[0]; pre[1 + LBL_C2] = cr[0]; pre[1 + LBL_R] = rec;
        for (int i = 0; i < K; ++i) { iv[i * 6 + LBL_C1] = cr[m.interval_epoch[i]]; iv[i * 6 + LBL_R] = rec; }
        (void)e;
    } else if (m.kind == MODEL_VARMIG) {                             // variable_migration_model.py:116-181
        const int e = m.n_epochs;
        const double rec = th[P - 1];
        bp[0] = 0.0;
        for (int i = 1; i < K; ++i) bp[i] = m.c_psmc[i];
        for (int i = 0; i < K; ++i) {
            const int ep = m.interval_epoch[i];
            const double c1 = th[ep], c2 = th[e + ep], m12 = th[2 * e + ep], m21 = th[3 * e + ep];
            er[i] = (c1 + c2) / 2.0; ebp[i] = bp[i];
            // make_rates_table_migration(c1, c2, m12, m21, recomb) is called positionally against the signature
            // (coal_1, coal_2, recomb, mig_12, mig_21) -- variable_migration_model.py:172-174 vs state_spaces.py:119-120.
            // The slip is reproduced: recombination gets m12, migration 1->2 gets m21, migration 2->1 gets recomb.
            iv[i * 6 + LBL_C1] = c1; iv[i * 6 + LBL_C2] = c2; iv[i * 6 + LBL_R] = m12;
            iv[i * 6 + LBL_M12] = m21; iv[i * 6 + LBL_M21] = rec;
        }
    } else {                                                         // isolation_with_migration_model_epochs.py:150-211
        const int e = m.n_epochs, nm = e * m.n_mig, na = e * m.n_anc;
        const double tau1 = th[0], tau2 = th[0] + th[1], rec = th[2];
        const double* cr = th + 3;            // 2e+1 coalescence rates
        const double* mr = th + 3 + 2 * e + 1;  // e migration rates
        double s_anc = 0.0, s_all = 0.0;
        for (int k = e + 1; k < 2 * e + 1; ++k) s_anc += cr[k];
        for (int k = 0; k < 2 * e + 1; ++k) s_all += cr[k];
        const double coal_anc = s_anc / e, coal_all = s_all / (2 * e + 1);
        for (int i = 0; i < nm; ++i) { bp[i] = ((double)i / nm) * (tau2 - tau1) + tau1; ebp[i] = bp[i]; }
        for (int i = 0; i < na; ++i) {
            bp[nm + i] = m.c_exp_a[i] / coal_anc + tau2;           // build_ctmc_system: mean of the ancestral rates
            ebp[nm + i] = m.c_exp_a[i] / coal_all + tau2;          // emission_points: mean of ALL rates (:160-165)
        }
        for (int i = 0; i < K; ++i) er[i] = coal_all;
        pre[0] = bp[0]; pre[1 + LBL_C1] = cr[0]; pre[1 + LBL_C2] = cr[0]; pre[1 + LBL_R] = rec;
        for (int i = 0; i < K; ++i) {
            iv[i * 6 + LBL_R] = rec;
            if (i < nm) {
                const int ep = i / m.n_mig;
                iv[i * 6 + LBL_C1] = cr[ep + 1]; iv[i * 6 + LBL_C2] = cr[ep + 1];
                iv[i * 6 + LBL_M12] = mr[ep]; iv[i * 6 + LBL_M21] = mr[ep];
            } else {
                iv[i * 6 + LBL_C1] = cr[(i - nm) / m.n_anc + e + 1];
            }
        }
    }
    for (int i = 0; i + 1 < K; ++i) iv[i * 6 + 5] = bp[i + 1] - bp[i];   // duration of interval i (the last is the pseudo interval)
    // Consecutive intervals with the same state space, the same rates and the same duration have the same transition
    // matrix (the reference memoises expm on the duration, CTMC.py:39-51; uniform break points give durations that are
    // equal up to the rounding of t_{i+1} - t_i, so "same" is taken as 1e-13 relative).  Only the first is computed.
    // An interval that hands over to a different state space stores a projected matrix and stays on its own.
    double* rep = pp + PointParams::rep_off(K);
    for (int i = 0; i < K; ++i) {
        int r = i;
        if (i > 0 && i + 1 < K) {
            const int j = (int)rep[i - 1];
            const bool plain_i = m.interval_space[i + 1] == m.interval_space[i];
            const bool plain_j = m.interval_space[j + 1] == m.interval_space[j];
            bool same = plain_i && plain_j && m.interval_space[i] == m.interval_space[j];
            for (int l = 0; l < 5; ++l) same = same && (iv[i * 6 + l] == iv[j * 6 + l]);
            same = same && fabs(iv[i * 6 + 5] - iv[j * 6 + 5]) <= 1e-13 * fabs(iv[j * 6 + 5]);
            if (same) r = j;
        }
        rep[i] = r;
    }
    write_emissions(ebp, er, K, E + (size_t)n * K * 3);
}

// ------------------------------------------------------------------------------------------------
// expm by uniformisation + scaling and squaring.  One CTA (256 threads) per (point, job).
// job 0 = pre-phase (isolation space), job 1+i = interval i (i < K-1).
// Shared memory: three n x n matrices.
// ------------------------------------------------------------------------------------------------
template <int TS>
__device__ __forceinline__ void smem_matmul(const double* A, const double* B, double* C, int n) {
    // C = A * B, 16x16 threads, TS x TS outputs each
    const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
    double acc[TS][TS];
#pragma unroll
    for (int r = 0; r < TS; ++r)
#pragma unroll
        for (int c = 0; c < TS; ++c) acc[r][c] = 0.0;
    for (int k = 0; k < n; ++k) {
        double a[TS], b[TS];
#pragma unroll
        for (int r = 0; r < TS; ++r) { const int row = ty + 16 * r; a[r] = row < n ? A[row * n + k] : 0.0; }
#pragma unroll
        for (int c = 0; c < TS; ++c) { const int col = tx + 16 * c; b[c] = col < n ? B[k * n + col] : 0.0; }
#pragma unroll
        for (int r = 0; r < TS; ++r)
#pragma unroll
            for (int c = 0; c < TS; ++c) acc[r][c] = fma(a[r], b[c], acc[r][c]);
    }
#pragma unroll
    for (int r = 0; r < TS; ++r)
#pragma unroll
        for (int c = 0; c < TS; ++c) {
            const int row = ty + 16 * r, col = tx + 16 * c;
            if (row < n && col < n) C[row * n + col] = acc[r][c];
        }
}

__device__ __forceinline__ void dmma884_acc(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// C = A * B for 16 < n <= 96 on the FP64 tensor path (mma.sync.m8n8k4.f64): 8 warps, warp w owns the 3 x 6 block of
// 8x8 output tiles at tile rows 3 (w/2).., tile columns 6 (w%2)..; per k-slice of 4 it loads 3 A fragments and 6 B
// fragments (one double per lane each) for 18 DMMAs, i.e. 0.5 bytes of shared memory per FMA, so the FP64 pipe and
// not the shared-memory pipe is the limit (the 6x6 register-tiled DFMA version moves 2.7 B/FMA and ran at ~22 % of peak).
__device__ __forceinline__ void smem_matmul_dmma(const double* A, const double* B, double* C, int n) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, q = lane & 3;                  // fragment coordinates
    const int tr0 = (warp >> 1) * 3, tc0 = (warp & 1) * 6;
    double acc[3][6][2];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 6; ++j) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }
    for (int k0 = 0; k0 < n; k0 += 4) {
        const int kk = k0 + q;
        double a[3], b[6];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const int row = 8 * (tr0 + i) + g;
            a[i] = (row < n && kk < n) ? A[row * n + kk] : 0.0;
        }
#pragma unroll
        for (int j = 0; j < 6; ++j) {
            const int col = 8 * (tc0 + j) + g;
            b[j] = (col < n && kk < n) ? B[kk * n + col] : 0.0;
        }
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 6; ++j) dmma884_acc(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 6; ++j) {
            const int row = 8 * (tr0 + i) + g, col = 8 * (tc0 + j) + 2 * q;
            if (row < n) {
                if (col < n) C[row * n + col] = acc[i][j][0];
                if (col + 1 < n) C[row * n + col + 1] = acc[i][j][1];
            }
        }
}

// SMALL: the launch holds only state spaces of at most 16 states (isolation-type models).  That instantiation carries no
// tensor-path code, needs a quarter of the registers and fits 8 CTAs per SM instead of 2 (0.13 -> 0.03 ms per 256 points).
// C = A * B for n <= 16 with any number of threads (the lean instantiation runs 64 per CTA)
__device__ __forceinline__ void smem_matmul_small(const double* A, const double* B, double* C, int n) {
    for (int x = threadIdx.x; x < n * n; x += blockDim.x) {
        const int row = x / n, col = x - row * n;
        double acc = 0.0;
        for (int k = 0; k < n; ++k) acc = fma(A[row * n + k], B[k * n + col], acc);
        C[x] = acc;
    }
}

template <bool SMALL>
__device__ __forceinline__ void smem_matmul_n(const double* A, const double* B, double* C, int n) {
    if (SMALL) smem_matmul_small(A, B, C, n);
    else if (n <= 16) smem_matmul<1>(A, B, C, n);
    else if constexpr (!SMALL) {
#ifdef IMC_EXPM_DFMA
        smem_matmul<6>(A, B, C, n);
#else
        smem_matmul_dmma(A, B, C, n);
#endif
    }
}

constexpr int EXPM_TAYLOR_DEGREE = 14;      // lambda <= 0.5: tail < 0.5^15/15! = 2.3e-17
constexpr double EXPM_LAMBDA_MAX = 0.5;

template <bool SMALL>
__global__ void __launch_bounds__(256) model_expm_kernel(ModelDev m, const double* params, const int* status,
                                                         double* pbuf, double* prebuf) {
    extern __shared__ double sm[];
    const int n_pt = blockIdx.y, job = blockIdx.x, K = m.K;
    if (status[n_pt] != 0) return;
    const double* pp = params + (size_t)n_pt * PointParams::size(K);
    int sp; const double* rates; double dt; int interval = job - 1;
    if (job == 0) {
        if (m.pre_space_n == 0) return;
        sp = SP_ISO; rates = pp + 1; dt = pp[0];
    } else {
        if ((int)pp[PointParams::rep_off(K) + interval] != interval) return;   // shares an earlier interval's matrix
        sp = m.interval_space[interval]; rates = pp + 6 + interval * 6; dt = rates[5];
    }
    const SpaceDev& S = m.space[sp];
    const int n = S.n, nn = n * n;
    double* Sm = sm; double* R = sm + nn; double* Tm = sm + 2 * nn;
    __shared__ double s_q;
    // rate matrix (CTMC.py:28-35): assignment per edge, diagonal = -row sum
    for (int x = threadIdx.x; x < nn; x += blockDim.x) Sm[x] = 0.0;
    __syncthreads();
    for (int e = threadIdx.x; e < S.n_edges; e += blockDim.x)
        Sm[S.edges[e * 3] * n + S.edges[e * 3 + 1]] = rates[S.edges[e * 3 + 2]];
    __syncthreads();
    if (threadIdx.x < n) {
        double rs = 0.0;
        for (int c = 0; c < n; ++c) rs += Sm[threadIdx.x * n + c];
        Tm[threadIdx.x] = rs;           // total exit rate of the state
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double q = 0.0;
        for (int r = 0; r < n; ++r) q = fmax(q, Tm[r]);
        s_q = q;
    }
    __syncthreads();
    const double q = s_q;
    const double lam_total = q * dt;
    int sq = 0;
    double lam = lam_total;
    while (lam > EXPM_LAMBDA_MAX && sq < 64) { lam *= 0.5; ++sq; }
    if (!(lam <= EXPM_LAMBDA_MAX)) lam = __longlong_as_double(0x7ff8000000000000LL);   // inf / NaN duration: poison the result
    // S = I + Q/q  (stochastic, non-negative); R = I
    {
        const double inv_q = q > 0.0 ? 1.0 / q : 0.0;
        double exit_rate[1];
        (void)exit_rate;
        for (int x = threadIdx.x; x < nn; x += blockDim.x) {
            const int r = x / n, c = x % n;
            double v = Sm[x] * inv_q;
            if (r == c) v = 1.0 - Tm[r] * inv_q;
            R[x] = (r == c) ? 1.0 : 0.0;
            Sm[x] = v;
        }
    }
    __syncthreads();
    if (lam > 0.0 || lam != lam) {
        // Horner: R <- I + (lam/k) S R, k = m..1
        for (int k = EXPM_TAYLOR_DEGREE; k >= 1; --k) {
            smem_matmul_n<SMALL>(Sm, R, Tm, n);
            __syncthreads();
            const double f = lam / k;
            for (int x = threadIdx.x; x < nn; x += blockDim.x) R[x] = Tm[x] * f + ((x / n == x % n) ? 1.0 : 0.0);
            __syncthreads();
        }
        const double el = exp(-lam);
        for (int x = threadIdx.x; x < nn; x += blockDim.x) R[x] *= el;
        __syncthreads();
        for (int s = 0; s < sq; ++s) {
            smem_matmul_n<SMALL>(R, R, Tm, n);
            __syncthreads();
            for (int x = threadIdx.x; x < nn; x += blockDim.x) R[x] = Tm[x];
            __syncthreads();
        }
    }
    // write out, projecting into the next interval's state space where the spaces differ
    if (job == 0) {
        double* out = prebuf + (size_t)n_pt * 16;
        for (int x = threadIdx.x; x < 16; x += blockDim.x) out[x] = R[x];
    } else {
        double* out = pbuf + (size_t)n_pt * m.p_stride + m.p_off[interval];
        const int sp_next = m.interval_space[min(interval + 1, K - 1)];
        if (sp_next == sp) {
            for (int x = threadIdx.x; x < nn; x += blockDim.x) out[x] = R[x];
        } else {   // migration -> single projection (isolation_with_migration_model.py:39-49)
            const int n2 = m.space[sp_next].n;
            for (int r = threadIdx.x; r < n; r += blockDim.x) {
                double* row = out + (size_t)r * n2;
                for (int c = 0; c < n2; ++c) row[c] = 0.0;
                for (int b = 0; b < n; ++b) row[m.proj_mig_single[b]] += R[r * n + b];
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// u chain, L-block products, joint matrix J -> pi, T   (transitions.py:204-248).  One CTA per point.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) model_chain_kernel(ModelDev m, const double* params, const double* pbuf,
                                                          const double* prebuf, int* status, double* pi, double* T, int stage) {
    extern __shared__ double sm[];
    const int n_pt = blockIdx.x, K = m.K, tid = threadIdx.x, nt = blockDim.x;
    if (status[n_pt] != 0) {   // invalid point: a harmless HMM; its log-likelihood is replaced by -inf afterwards
        for (int i = tid; i < K; i += nt) {
            pi[(size_t)n_pt * K + i] = 1.0 / K;
            for (int j = 0; j < K; ++j) T[((size_t)n_pt * K + i) * K + j] = (i == j) ? 1.0 : 0.0;
        }
        return;
    }
    double* u = sm;                          // [MAX_STATES]
    double* u2 = u + MAX_STATES;             // [MAX_STATES]
    double* J = u2 + MAX_STATES;             // [K][K]
    double* upth = J + K * K;                // [K][MAX_L]   up_through_i on L(space(i+1))
    double* esum = upth + K * MAX_L;         // [K][MAX_L]   sum over E(space(j+1)) of P_j[l][e], l in L(space(j))
    double* red = esum + K * MAX_L;          // [nt]
    // small models (stage != 0): all of the point's interval matrices are copied into shared memory first -- the chains below
    // are serial and would otherwise wait for L2 at every link
    const double* P = pbuf + (size_t)n_pt * m.p_stride;
    if (stage) {
        double* Ps = red + nt;
        for (long long x = tid; x < m.p_stride; x += nt) Ps[x] = P[x];
        P = Ps;
    }
    const double* rep = params + (size_t)n_pt * PointParams::size(K) + PointParams::rep_off(K);
    for (int x = tid; x < K * K; x += nt) J[x] = 0.0;
    // u_0: row `initial` of upto0
    {
        const SpaceDev& S0 = m.space[m.interval_space[0]];
        for (int x = tid; x < S0.n; x += nt) u[x] = 0.0;
        __syncthreads();
        if (m.pre_space_n == 0) {
            if (tid == 0) u[m.initial_state] = 1.0;                 // variable_migration_model.py:74-75: identity
        } else if (tid == 0) {
            const double* pre = prebuf + (size_t)n_pt * 16 + m.initial_state * 4;
            const int* proj = (m.interval_space[0] == SP_SINGLE) ? m.proj_iso_single : m.proj_iso_mig;
            for (int b = 0; b < 4; ++b) u[proj[b]] += pre[b];
        }
        __syncthreads();
    }
    for (int i = 0; i < K; ++i) {
        const int sp = m.interval_space[i], spn = m.interval_space[min(i + 1, K - 1)];
        const SpaceDev& S = m.space[sp];
        const SpaceDev& Sn = m.space[spn];
        if (i == K - 1) {
            // J[K-1][K-1] = sum_b u[b]   (transitions.py:227-228); pseudo through: esum = 1
            if (tid == 0) {
                double s = 0.0;
                for (int b = 0; b < S.nB; ++b) s += u[S.B[b]];
                J[i * K + i] = s;
            }
            for (int l = tid; l < S.nL; l += nt) esum[i * MAX_L + l] = 1.0;
            break;
        }
        const double* Pi = P + m.p_off[(int)rep[i]];
        const int n1 = S.n, n2 = Sn.n;
        // esum_i[l] = sum_{e in E(next)} P_i[L[l]][e]
        for (int l = tid; l < S.nL; l += nt) {
            double s = 0.0;
            for (int e = 0; e < Sn.nE; ++e) s += Pi[(size_t)S.L[l] * n2 + Sn.E[e]];
            esum[i * MAX_L + l] = s;
        }
        // J[i][i] = sum_b u[b] * sum_e P_i[b][e]   (1 <= i < K-1; J[0][0] is overwritten below)
        {
            double part = 0.0;
            for (int b = tid; b < S.nB; b += nt) {
                double s = 0.0;
                for (int e = 0; e < Sn.nE; ++e) s += Pi[(size_t)S.B[b] * n2 + Sn.E[e]];
                part += u[S.B[b]] * s;
            }
            // fixed-shape tree (warp shuffles, then the warps' sums in order): deterministic, and not 128 serial reads by thread 0
#pragma unroll
            for (int mk = 16; mk >= 1; mk >>= 1) {
                int lo = __double2loint(part), hi = __double2hiint(part);
                lo = __shfl_xor_sync(0xffffffffu, lo, mk);
                hi = __shfl_xor_sync(0xffffffffu, hi, mk);
                part += __hiloint2double(hi, lo);
            }
            if ((tid & 31) == 0) red[tid >> 5] = part;
        }
        __syncthreads();
        if (tid == 0) {
            double s = 0.0;
            for (int x = 0; x < (nt >> 5); ++x) s += red[x];
            J[i * K + i] = s;
        }
        // up_through_i[l] = sum_b u[b] P_i[b][Ln[l]]
        for (int l = tid; l < Sn.nL; l += nt) {
            double s = 0.0;
            for (int b = 0; b < S.nB; ++b) s += u[S.B[b]] * Pi[(size_t)S.B[b] * n2 + Sn.L[l]];
            upth[i * MAX_L + l] = s;
        }
        // u_{i+1} = u_i P_i
        for (int c = tid; c < n2; c += nt) {
            double s = 0.0;
            for (int r = 0; r < n1; ++r) s += u[r] * Pi[(size_t)r * n2 + c];
            u2[c] = s;
        }
        __syncthreads();
        for (int c = tid; c < n2; c += nt) u[c] = u2[c];
        __syncthreads();
        if (i == 0 && tid == 0) {
            // joint[0,0] = up_to(1)[initial, end_states(0)].sum()  (transitions.py:222) -- end states of interval 0's
            // space indexed into interval 1's vector; the model constructors guarantee both intervals share a space.
            double s = 0.0;
            for (int e = 0; e < S.nE; ++e) s += u[S.E[e]];
            J[0] = s;
        }
        __syncthreads();
    }
    __syncthreads();
    // i < j: w_{i,i+1} = up_through_i; J[i][j] = w . esum_j; w <- w P_j[L(j), L(j+1)]
    const int warp = tid >> 5, lane = tid & 31, nwarps = nt >> 5;
    for (int i = warp; i < K - 1; i += nwarps) {
        double w = lane < MAX_L ? upth[i * MAX_L + lane] : 0.0;
        int nl = m.space[m.interval_space[i + 1]].nL;
        if (lane >= nl) w = 0.0;
        for (int j = i + 1; j < K; ++j) {
            const SpaceDev& S = m.space[m.interval_space[j]];
            double t = lane < S.nL ? w * esum[j * MAX_L + lane] : 0.0;
#pragma unroll
            for (int mk = 16; mk >= 1; mk >>= 1) t += __shfl_xor_sync(0xffffffffu, t, mk);
            if (lane == 0) { J[i * K + j] = t; J[j * K + i] = t; }
            if (j + 1 < K) {
                const SpaceDev& Sn = m.space[m.interval_space[j + 1]];
                const double* Pj = P + m.p_off[(int)rep[j]];
                double nw = 0.0;
                for (int l = 0; l < S.nL; ++l) {
                    const double wl = __shfl_sync(0xffffffffu, w, l);
                    if (lane < Sn.nL) nw += wl * Pj[(size_t)S.L[l] * Sn.n + Sn.L[lane]];
                }
                w = lane < Sn.nL ? nw : 0.0;
            }
        }
    }
    __syncthreads();
    // joint.sum() ~ 1 to 7 decimals (transitions.py:239), pi and T (transitions.py:241-246)
    if (tid == 0) {
        double tot = 0.0;
        for (int x = 0; x < K * K; ++x) tot += J[x];
        if (!(fabs(tot - 1.0) < 1.5e-7)) status[n_pt] = 2;
    }
    for (int i = tid; i < K; i += nt) {
        double s = 0.0;
        for (int j = 0; j < K; ++j) s += J[i * K + j];
        pi[(size_t)n_pt * K + i] = s;
        for (int j = 0; j < K; ++j) T[((size_t)n_pt * K + i) * K + j] = J[i * K + j] / s;
    }
}

// out[n] = -inf for invalid parameter points (likelihood.py:29-30), NaN when the joint matrix failed its check
__global__ void model_status_fixup_kernel(const int* status, int N, double* out) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    if (status[n] == 1) out[n] = -INFINITY;
    else if (status[n] == 2) out[n] = __longlong_as_double(0x7ff8000000000000LL);
}

}  // namespace imc
