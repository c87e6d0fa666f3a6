/*
 * oracle/forward_oracle.c -- CPU restatement of the reference's forward log-likelihood.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (imcoalhmm_b200/, include/, the CUDA
 * library) may link, import or call this file; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs use it, and only as the checker or the timed CPU baseline.
 *
 * PARITY UNPINNED at the forward boundary: the reference delegates the arithmetic to the external,
 * un-vendored, un-pinned `ziphmm` module (birc-aeh/mini-ziphmm; named bare in
 * /root/reference/requirements.txt:4 and setup.py:27-30), which is not present in /root/reference and
 * cannot be installed here.  The reference's own tests pin no log-likelihood.  This file therefore
 * restates the *published* algorithm at the reference's call sites:
 *
 *   src/IMCoalHMM/hmm.py:16         ziphmm.preprocess_raw_observations(obs, NSYM)
 *                                       -> (new_obs, sym2pair, new_nsyms)
 *   src/IMCoalHMM/hmm.py:20-21      ziphmm.zip_forward(pi, T, E, sym2pair, new_obs, NSYM, new_nsyms) -> logL
 *   src/IMCoalHMM/likelihood.py:33  logL(theta) = sum over forwarders (each file restarts from pi)
 *
 * i.e. the scaled HMM forward recursion
 *     alpha_0 = pi o E[:,o_0];  alpha_t = (T^T alpha_{t-1}) o E[:,o_t];  logL = sum_t log(sum_j alpha_t[j])
 * with T[i][j] = P(next=j | cur=i) (transitions.py:243-246) and E[state][symbol] (emissions.py:95-99),
 * and its zipHMM form (Sand et al. 2013, cited at documentation/development-manual/bibliography.bib:45-47):
 * symbol matrices C_s = diag(E[:,s]) T^T, pair symbols C_(a,b) = C_b C_a, one mat-vec per compressed symbol.
 * It is pinned mathematically (tests/test_oracle_forward.py): brute-force path enumeration, long double
 * vs double, all-missing => 0, single site, chunk additivity, time-reversal invariance, plain == zip.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ---------------------------------------------------------------- plain forward, double ---- */
double imco_forward_plain(const int32_t* obs, int64_t L, int K, int S,
                          const double* pi, const double* T, const double* E)
{
    if (L <= 0) return 0.0;
    double* a = (double*)malloc(sizeof(double) * 2 * (size_t)K);
    double* b = a + K;
    double logl = 0.0, c = 0.0;
    for (int j = 0; j < K; ++j) { a[j] = pi[j] * E[(size_t)j * S + obs[0]]; c += a[j]; }
    for (int j = 0; j < K; ++j) a[j] /= c;
    logl += log(c);
    for (int64_t t = 1; t < L; ++t) {
        const int o = obs[t];
        c = 0.0;
        for (int j = 0; j < K; ++j) {
            double s = 0.0;
            for (int i = 0; i < K; ++i) s += a[i] * T[(size_t)i * K + j];
            b[j] = s * E[(size_t)j * S + o];
            c += b[j];
        }
        for (int j = 0; j < K; ++j) a[j] = b[j] / c;
        logl += log(c);
    }
    free(a);
    return logl;
}

/* ---------------------------------------------------------------- plain forward, long double --
 * Same recursion in x87 extended precision (64-bit mantissa).  Returns the result rounded to
 * double and, through out_hi/out_lo, as an unevaluated double-double sum so that Python (which has
 * no portable long double in ctypes return values) can see the extra bits. */
double imco_forward_plain_ld(const int32_t* obs, int64_t L, int K, int S,
                             const double* pi, const double* T, const double* E,
                             double* out_hi, double* out_lo)
{
    long double total = 0.0L;
    if (L > 0) {
        long double* a = (long double*)malloc(sizeof(long double) * 2 * (size_t)K);
        long double* b = a + K;
        long double c = 0.0L;
        for (int j = 0; j < K; ++j) {
            a[j] = (long double)pi[j] * (long double)E[(size_t)j * S + obs[0]];
            c += a[j];
        }
        for (int j = 0; j < K; ++j) a[j] /= c;
        total += logl(c);
        for (int64_t t = 1; t < L; ++t) {
            const int o = obs[t];
            c = 0.0L;
            for (int j = 0; j < K; ++j) {
                long double s = 0.0L;
                for (int i = 0; i < K; ++i) s += a[i] * (long double)T[(size_t)i * K + j];
                b[j] = s * (long double)E[(size_t)j * S + o];
                c += b[j];
            }
            for (int j = 0; j < K; ++j) a[j] = b[j] / c;
            total += logl(c);
        }
        free(a);
    }
    double hi = (double)total;
    if (out_hi) *out_hi = hi;
    if (out_lo) *out_lo = (double)(total - (long double)hi);
    return hi;
}

/* ---------------------------------------------------------------- zipHMM-style preprocessing --
 * Greedy byte-pair style compression: repeatedly replace the most frequent adjacent pair (a,b) by a
 * fresh symbol id.  Stops when the best pair occurs fewer than min_count times or max_syms is reached.
 * Any such re-encoding is exact (the pair matrix is the product of its parts), so the particular
 * choices need not match mini-ziphmm's.  Outputs are malloc'ed; release with imco_free().
 *   new_obs[newL], sym2pair[2*(new_nsyms - nsym)] with sym2pair[2*(id-nsym)+{0,1}] = (left, right). */
int imco_zip_preprocess(const int32_t* obs, int64_t L, int nsym, int min_count, int max_syms,
                        int32_t** new_obs_out, int64_t* newL_out, int32_t** sym2pair_out, int* new_nsyms_out)
{
    if (min_count < 2) min_count = 2;
    if (max_syms < nsym) max_syms = nsym;
    int32_t* cur = (int32_t*)malloc(sizeof(int32_t) * (size_t)(L > 0 ? L : 1));
    if (L > 0) memcpy(cur, obs, sizeof(int32_t) * (size_t)L);
    int32_t* pairs = (int32_t*)malloc(sizeof(int32_t) * 2 * (size_t)(max_syms - nsym + 1));
    int64_t n = L;
    int ns = nsym;
    int64_t* counts = NULL;
    while (ns < max_syms && n >= 2) {
        counts = (int64_t*)realloc(counts, sizeof(int64_t) * (size_t)ns * (size_t)ns);
        memset(counts, 0, sizeof(int64_t) * (size_t)ns * (size_t)ns);
        for (int64_t t = 0; t + 1 < n; ++t) counts[(size_t)cur[t] * ns + cur[t + 1]]++;
        int64_t best = 0; int ba = 0, bb = 0;
        for (int a = 0; a < ns; ++a)
            for (int b = 0; b < ns; ++b)
                if (counts[(size_t)a * ns + b] > best) { best = counts[(size_t)a * ns + b]; ba = a; bb = b; }
        if (best < min_count) break;
        int64_t w = 0;
        for (int64_t t = 0; t < n; ) {
            if (t + 1 < n && cur[t] == ba && cur[t + 1] == bb) { cur[w++] = ns; t += 2; }
            else { cur[w++] = cur[t]; t += 1; }
        }
        n = w;
        pairs[2 * (ns - nsym)] = ba;
        pairs[2 * (ns - nsym) + 1] = bb;
        ns++;
    }
    free(counts);
    *new_obs_out = cur;
    *newL_out = n;
    *sym2pair_out = pairs;
    *new_nsyms_out = ns;
    return 0;
}

void imco_free(void* p) { free(p); }

/* ---------------------------------------------------------------- zip forward ---------------
 * Symbol matrices are stored scaled to unit total mass with their log-scale kept aside:
 *   M_s = C_s / sum(C_s),  ls_s = log(sum(C_s)).
 * The first (possibly compound) symbol is applied to pi by walking its left spine. */
typedef struct { int K; int nsym; int nsyms; double* M; double* ls; } zipmats;

static void zip_build(zipmats* z, int K, int S, int nsym, int nsyms, const int32_t* sym2pair,
                      const double* T, const double* E)
{
    z->K = K; z->nsym = nsym; z->nsyms = nsyms;
    z->M = (double*)malloc(sizeof(double) * (size_t)nsyms * K * K);
    z->ls = (double*)malloc(sizeof(double) * (size_t)nsyms);
    for (int s = 0; s < nsym; ++s) {
        double* M = z->M + (size_t)s * K * K;
        double tot = 0.0;
        for (int j = 0; j < K; ++j)
            for (int i = 0; i < K; ++i) { M[j * K + i] = E[(size_t)j * S + s] * T[(size_t)i * K + j]; tot += M[j * K + i]; }
        for (int x = 0; x < K * K; ++x) M[x] /= tot;
        z->ls[s] = log(tot);
    }
    for (int s = nsym; s < nsyms; ++s) {
        const int a = sym2pair[2 * (s - nsym)], b = sym2pair[2 * (s - nsym) + 1];
        const double* A = z->M + (size_t)a * K * K;
        const double* B = z->M + (size_t)b * K * K;
        double* M = z->M + (size_t)s * K * K;
        double tot = 0.0;
        for (int r = 0; r < K; ++r)
            for (int c = 0; c < K; ++c) {
                double acc = 0.0;
                for (int k = 0; k < K; ++k) acc += B[r * K + k] * A[k * K + c];   /* C_(a,b) = C_b C_a */
                M[r * K + c] = acc; tot += acc;
            }
        for (int x = 0; x < K * K; ++x) M[x] /= tot;
        z->ls[s] = log(tot) + z->ls[a] + z->ls[b];
    }
}

static void zip_free(zipmats* z) { free(z->M); free(z->ls); }

/* v <- M_s v, renormalise, return log of the scale taken out (including ls_s) */
static double zip_apply(const zipmats* z, int s, double* v, double* tmp)
{
    const int K = z->K;
    const double* M = z->M + (size_t)s * K * K;
    double c = 0.0;
    for (int r = 0; r < K; ++r) {
        double acc = 0.0;
        for (int k = 0; k < K; ++k) acc += M[r * K + k] * v[k];
        tmp[r] = acc; c += acc;
    }
    for (int r = 0; r < K; ++r) v[r] = tmp[r] / c;
    return log(c) + z->ls[s];
}

/* initial vector for a (possibly compound) first symbol */
static double zip_init(const zipmats* z, int s, int S, const int32_t* sym2pair, const double* pi,
                       const double* E, double* v, double* tmp)
{
    const int K = z->K;
    if (s < z->nsym) {
        double c = 0.0;
        for (int j = 0; j < K; ++j) { v[j] = pi[j] * E[(size_t)j * S + s]; c += v[j]; }
        for (int j = 0; j < K; ++j) v[j] /= c;
        return log(c);
    }
    const int a = sym2pair[2 * (s - z->nsym)], b = sym2pair[2 * (s - z->nsym) + 1];
    double l = zip_init(z, a, S, sym2pair, pi, E, v, tmp);
    return l + zip_apply(z, b, v, tmp);
}

double imco_zip_forward(const double* pi, const double* T, const double* E, const int32_t* sym2pair,
                        const int32_t* new_obs, int64_t newL, int nsym, int new_nsyms, int K)
{
    if (newL <= 0) return 0.0;
    zipmats z;
    zip_build(&z, K, nsym, nsym, new_nsyms, sym2pair, T, E);
    double* v = (double*)malloc(sizeof(double) * 2 * (size_t)K);
    double* tmp = v + K;
    double logl = zip_init(&z, new_obs[0], nsym, sym2pair, pi, E, v, tmp);
    for (int64_t t = 1; t < newL; ++t) logl += zip_apply(&z, new_obs[t], v, tmp);
    free(v);
    zip_free(&z);
    return logl;
}

/* ---------------------------------------------------------------- zip forward, tuned for the timed CPU arm ----
 * Same algorithm as imco_zip_forward, written the way a performance-minded C implementation would be, so that the
 * CPU baseline bench.py reports is not an artificially slow one: symbol matrices stored column-major so that the
 * mat-vec and the pair products vectorise without reassociation (independent accumulators per row), rescaling by
 * exact powers of two every 8 symbols with an integer exponent (no division and no log per symbol).
 * tests/test_oracle_forward.py holds it to the plain and the simple zip forward. */
typedef struct { int K; double* M; long long* ex; } zipfast;   /* M[s][k][r] = C_s[r][k] / 2^ex_s */

static long long fast_normalise(double* M, int n)
{
    double mx = 0.0;
    for (int x = 0; x < n; ++x) if (M[x] > mx) mx = M[x];
    if (!(mx > 0.0) || !(mx < 1.7e308)) return 0;
    int e;
    frexp(mx, &e);
    const double f = ldexp(1.0, -e);
    for (int x = 0; x < n; ++x) M[x] *= f;
    return e;
}

static inline __attribute__((always_inline)) double zip_fast_impl(const double* pi, const double* T, const double* E, const int32_t* sym2pair,
                                                              const int32_t* new_obs, int64_t newL, int nsym, int new_nsyms, const int K)
{
    if (newL <= 0) return 0.0;
    const int KK = K * K;
    zipfast z;
    z.K = K;
    z.M = (double*)malloc(sizeof(double) * (size_t)new_nsyms * KK);
    z.ex = (long long*)malloc(sizeof(long long) * (size_t)new_nsyms);
    for (int s = 0; s < nsym; ++s) {          /* C_s[r][k] = E[r][s] T[k][r], stored [k][r] */
        double* M = z.M + (size_t)s * KK;
        for (int k = 0; k < K; ++k)
            for (int r = 0; r < K; ++r) M[k * K + r] = E[(size_t)r * nsym + s] * T[(size_t)k * K + r];
        z.ex[s] = fast_normalise(M, KK);
    }
    for (int s = nsym; s < new_nsyms; ++s) {  /* C_(a,b) = C_b C_a:  out[:,c] = sum_k C_b[:,k] C_a[k,c] */
        const int a = sym2pair[2 * (s - nsym)], b = sym2pair[2 * (s - nsym) + 1];
        const double* A = z.M + (size_t)a * KK;
        const double* B = z.M + (size_t)b * KK;
        double* M = z.M + (size_t)s * KK;
        for (int c = 0; c < K; ++c) {
            double* out = M + c * K;
            for (int r = 0; r < K; ++r) out[r] = 0.0;
            for (int k = 0; k < K; ++k) {
                const double w = A[c * K + k];            /* C_a[k][c] */
                const double* col = B + k * K;            /* C_b[:, k] */
                for (int r = 0; r < K; ++r) out[r] += col[r] * w;
            }
        }
        z.ex[s] = z.ex[a] + z.ex[b] + fast_normalise(M, KK);
    }
    double* v = (double*)malloc(sizeof(double) * 2 * (size_t)K);
    double* tmp = v + K;
    long long scale = 0;
    int dead = 0;
    /* the first (possibly compound) symbol: walk its left spine down to a raw symbol, then apply the right parts */
    int* stack = (int*)malloc(sizeof(int) * (size_t)(new_nsyms + 1));
    int sp = 0, s0 = new_obs[0];
    while (s0 >= nsym) { stack[sp++] = sym2pair[2 * (s0 - nsym) + 1]; s0 = sym2pair[2 * (s0 - nsym)]; }
    for (int j = 0; j < K; ++j) v[j] = pi[j] * E[(size_t)j * nsym + s0];
    int64_t t = 1;
    int pending = sp;
    for (int count = 0;; ++count) {
        int s;
        if (pending > 0) s = stack[--pending];
        else if (t < newL) s = new_obs[t++];
        else break;
        const double* M = z.M + (size_t)s * KK;
        for (int r = 0; r < K; ++r) tmp[r] = 0.0;
        for (int k = 0; k < K; ++k) {
            const double w = v[k];
            const double* col = M + k * K;
            for (int r = 0; r < K; ++r) tmp[r] += col[r] * w;
        }
        scale += z.ex[s];
        double* sw = v; v = tmp; tmp = sw;
        if ((count & 7) == 7) {
            double c = 0.0;
            for (int r = 0; r < K; ++r) c += v[r];
            if (c > 0.0 && c < 1.7e308) {
                int e;
                frexp(c, &e);
                const double f = ldexp(1.0, -e);
                for (int r = 0; r < K; ++r) v[r] *= f;
                scale += e;
            } else if (!(c > 0.0)) { dead = 1; }
        }
    }
    double c = 0.0;
    for (int r = 0; r < K; ++r) c += v[r];
    const double result = (dead || !(c > 0.0)) ? (c != c ? c : -INFINITY) : log(c) + (double)scale * 0.693147180559945309417232121458;
    free(v < tmp ? v : tmp);
    free(stack);
    free(z.M); free(z.ex);
    return result;
}

/* state counts of the benchmark configurations get their own copy with K known at compile time, so that the compiler
 * unrolls and vectorises the length-K loops (what a hand-tuned implementation would do) */
#define ZIP_FAST_K(KC)                                                                                                  \
    static double zip_fast_k##KC(const double* pi, const double* T, const double* E, const int32_t* sym2pair,            \
                                 const int32_t* new_obs, int64_t newL, int nsym, int new_nsyms)                          \
    { return zip_fast_impl(pi, T, E, sym2pair, new_obs, newL, nsym, new_nsyms, KC); }
ZIP_FAST_K(4)
ZIP_FAST_K(8)
ZIP_FAST_K(10)
ZIP_FAST_K(16)
ZIP_FAST_K(20)
ZIP_FAST_K(40)

double imco_zip_forward_fast(const double* pi, const double* T, const double* E, const int32_t* sym2pair,
                             const int32_t* new_obs, int64_t newL, int nsym, int new_nsyms, int K)
{
    switch (K) {
        case 4: return zip_fast_k4(pi, T, E, sym2pair, new_obs, newL, nsym, new_nsyms);
        case 8: return zip_fast_k8(pi, T, E, sym2pair, new_obs, newL, nsym, new_nsyms);
        case 10: return zip_fast_k10(pi, T, E, sym2pair, new_obs, newL, nsym, new_nsyms);
        case 16: return zip_fast_k16(pi, T, E, sym2pair, new_obs, newL, nsym, new_nsyms);
        case 20: return zip_fast_k20(pi, T, E, sym2pair, new_obs, newL, nsym, new_nsyms);
        case 40: return zip_fast_k40(pi, T, E, sym2pair, new_obs, newL, nsym, new_nsyms);
    }
    return zip_fast_impl(pi, T, E, sym2pair, new_obs, newL, nsym, new_nsyms, K);
}

/* ---------------------------------------------------------------- batched CPU baseline ------
 * out[n] = sum_c forward(seq_c; pi_n, T_n, E_n) -- what likelihood.py:33 computes for each of N
 * parameter points, fanned out over the host cores (one (n, c) pair per task).  mode 0 = plain,
 * mode 1 = zip (sequences already preprocessed: obs[c] is new_obs, pairs[c]/nsyms[c] its tables),
 * mode 2 = the same through imco_zip_forward_fast (the timed CPU arm). */
int imco_forward_batch(int mode, int C, const int32_t* const* obs, const int64_t* lens,
                       const int32_t* const* pairs, const int* nsyms, int nsym,
                       int N, int K, const double* pi, const double* T, const double* E,
                       double* out, int nthreads)
{
    double* part = (double*)malloc(sizeof(double) * (size_t)N * C);
    const int64_t tasks = (int64_t)N * C;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#pragma omp parallel for schedule(dynamic, 1)
#endif
    for (int64_t task = 0; task < tasks; ++task) {
        const int n = (int)(task / C), c = (int)(task % C);
        const double* pin = pi + (size_t)n * K;
        const double* Tn = T + (size_t)n * K * K;
        const double* En = E + (size_t)n * K * nsym;
        part[task] = mode == 0 ? imco_forward_plain(obs[c], lens[c], K, nsym, pin, Tn, En)
                   : mode == 1 ? imco_zip_forward(pin, Tn, En, pairs[c], obs[c], lens[c], nsym, nsyms[c], K)
                               : imco_zip_forward_fast(pin, Tn, En, pairs[c], obs[c], lens[c], nsym, nsyms[c], K);
    }
    for (int n = 0; n < N; ++n) {
        double s = 0.0;
        for (int c = 0; c < C; ++c) s += part[(size_t)n * C + c];
        out[n] = s;
    }
    free(part);
    int used = 1;
#ifdef _OPENMP
    used = omp_get_max_threads();
#endif
    return used;
}
