// Host-side zipHMM-style preprocessing for the compressed forward kernel (zip_kernels.cuh).
//
// Reference contract: /root/reference/src/IMCoalHMM/hmm.py:16  ziphmm.preprocess_raw_observations(obs, NSYM)
// -> (new_obs, sym2pair, new_nsyms): the observation sequence re-encoded over an extended alphabet in which
// every id >= NSYM stands for an adjacent pair (left, right) of earlier ids.  Any such re-encoding is exact
// (the pair's matrix is the product of its parts), so the particular merges need not match mini-ziphmm's.
//
// Differences that come from the GPU, not from the algorithm:
//   * ONE dictionary is shared by all chunks of a sequence set (the kernel keeps one copy of every
//     dictionary matrix per parameter point in shared memory), learned on a bounded prefix sample;
//   * at most 256 ids (one byte per token); a forward call with K states uses the first M(K) ids that fit
//     in shared memory, tokens with larger ids being expanded back into their parts (zip_expand);
//   * position 0 of every chunk is kept out of the token stream (alpha_0 = pi o E[:,o_0] has no T factor).
#include <array>
#include <thread>

namespace imc {

struct ZipMerges {
    int nsym = 0;
    std::vector<std::array<uint8_t, 2>> pairs;   // pairs[i] = (left, right) of id nsym + i, creation order
    int size() const { return nsym + (int)pairs.size(); }
};

// one left-to-right pass: replace every (a, b) by id, in place; returns the new length
static inline size_t zip_replace(uint8_t* s, size_t n, uint8_t a, uint8_t b, uint8_t id) {
    size_t w = 0, t = 0;
    // skip the untouched prefix quickly
    while (t + 1 < n && !(s[t] == a && s[t + 1] == b)) ++t;
    w = t;
    while (t < n) {
        if (t + 1 < n && s[t] == a && s[t + 1] == b) { s[w++] = id; t += 2; }
        else { s[w++] = s[t++]; }
    }
    return w;
}

// Greedy most-frequent-adjacent-pair merges learned on `sample` (each entry one chunk without its first symbol;
// the buffers are consumed).  Stops at max_ids or when the best pair occurs fewer than min_count times.
static ZipMerges zip_learn(std::vector<std::vector<uint8_t>>& sample, int nsym, int max_ids, long long min_count) {
    ZipMerges mg;
    mg.nsym = nsym;
    if (max_ids > 256) max_ids = 256;
    std::vector<long long> counts;
    int ns = nsym;
    while (ns < max_ids) {
        counts.assign((size_t)ns * ns, 0);
        for (const auto& s : sample)
            for (size_t t = 0; t + 1 < s.size(); ++t) counts[(size_t)s[t] * ns + s[t + 1]]++;
        long long best = 0;
        int ba = 0, bb = 0;
        for (int a = 0; a < ns; ++a)
            for (int b = 0; b < ns; ++b)
                if (counts[(size_t)a * ns + b] > best) { best = counts[(size_t)a * ns + b]; ba = a; bb = b; }
        if (best < min_count) break;
        for (auto& s : sample) s.resize(zip_replace(s.data(), s.size(), (uint8_t)ba, (uint8_t)bb, (uint8_t)ns));
        mg.pairs.push_back({(uint8_t)ba, (uint8_t)bb});
        ++ns;
    }
    return mg;
}

// apply the merges in creation order to one chunk (symbols 1..L-1)
static void zip_encode(const ZipMerges& mg, const uint8_t* sym, size_t n, std::vector<uint8_t>& out) {
    out.assign(sym, sym + n);
    size_t len = n;
    for (size_t i = 0; i < mg.pairs.size() && len >= 2; ++i)
        len = zip_replace(out.data(), len, mg.pairs[i][0], mg.pairs[i][1], (uint8_t)(mg.nsym + i));
    out.resize(len);
    out.shrink_to_fit();
}

// tokens over the first M ids only: larger ids are expanded into their parts (left first)
static void zip_expand(const ZipMerges& mg, const std::vector<uint8_t>& in, int M, std::vector<uint8_t>& out) {
    out.clear();
    out.reserve(in.size());
    uint8_t stack[512];
    for (uint8_t tok : in) {
        if (tok < M) { out.push_back(tok); continue; }
        int sp = 0;
        stack[sp++] = tok;
        while (sp > 0) {
            const uint8_t x = stack[--sp];
            if (x < M) { out.push_back(x); continue; }
            const auto& p = mg.pairs[x - mg.nsym];
            stack[sp++] = p[1];   // right is applied after left
            stack[sp++] = p[0];
        }
    }
}

// Dictionary of the first M ids renumbered by level (a pair's level = 1 + max level of its parts) so that the
// kernel can build all entries of one level in parallel: perm[old id] = new id, pairs_new[new id] = (left, right)
// in new ids (entries < nsym unused), level_start[l] .. level_start[l+1] = new ids of level l+1 (pairs only).
struct ZipLevels {
    std::vector<uint8_t> perm;
    std::vector<uint8_t> pairs;        // [M][2]
    std::vector<int> level_start;      // size nlevels + 1, level_start[0] = nsym
};

static ZipLevels zip_levels(const ZipMerges& mg, int M) {
    ZipLevels zl;
    std::vector<int> level(M, 0), order;
    for (int id = mg.nsym; id < M; ++id) {
        const auto& p = mg.pairs[id - mg.nsym];
        level[id] = 1 + std::max(level[p[0]], level[p[1]]);
        order.push_back(id);
    }
    std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return level[x] < level[y]; });
    zl.perm.resize(M);
    for (int s = 0; s < mg.nsym && s < M; ++s) zl.perm[s] = (uint8_t)s;
    for (size_t i = 0; i < order.size(); ++i) zl.perm[order[i]] = (uint8_t)(mg.nsym + i);
    zl.pairs.assign((size_t)M * 2, 0);
    zl.level_start.push_back(mg.nsym);
    int cur = 1;
    for (size_t i = 0; i < order.size(); ++i) {
        const int id = order[i];
        while (level[id] > cur) { zl.level_start.push_back(mg.nsym + (int)i); ++cur; }
        const auto& p = mg.pairs[id - mg.nsym];
        zl.pairs[(size_t)(mg.nsym + i) * 2] = zl.perm[p[0]];
        zl.pairs[(size_t)(mg.nsym + i) * 2 + 1] = zl.perm[p[1]];
    }
    zl.level_start.push_back(M);
    return zl;
}

// ------------------------------------------------------------------------------------------------
// Run tokens for the spectral form of the kernel (zip_kernels.cuh, ZipArgs::spec).  The most frequent symbol r of a
// set (the "run symbol": matching sites in a pairwise alignment) is taken out of the dictionary altogether: in the
// eigenbasis of C_r a run of n such sites is the diagonal matrix Lambda^n, whatever n is.  A chunk (symbols 1..L-1)
// becomes   first_run x r,  then tokens (id, n) = "dictionary entry id, followed by n x r",   one 32-bit word each:
//     word = id | n << 8,   id < 256 (ids < nsym: single symbols, larger ids: pairs as above),   n <= RUN_MAX.
// Pairs are learned on the entries only: (x, y) may merge where x carries no run (n == 0); the merged token keeps
// y's run.  Any such encoding is exact, like the pair dictionary itself.  Runs longer than RUN_MAX are continued by
// tokens whose entry is the run symbol itself (one site) -- rare, and it keeps the kernel free of special cases.
// ------------------------------------------------------------------------------------------------
// (RUN_LO_BITS, RUN_MAX, ...: zip_kernels.cuh)
static inline uint32_t run_word(int id, int n) { return (uint32_t)id | ((uint32_t)n << 8); }

// symbols 1..L-1 of a chunk -> leading run + run tokens over the plain symbols (no pairs yet)
static void run_tokenize(const uint8_t* sym, size_t n, int run_sym, int* first_run, std::vector<uint32_t>& out) {
    out.clear();
    size_t t = 0;
    while (t < n && sym[t] == run_sym && (int)t < RUN_MAX) ++t;
    *first_run = (int)t;
    while (t < n) {
        const int id = sym[t++];               // equals run_sym only when it continues an over-long run
        int r = 0;
        while (t < n && sym[t] == run_sym && r < RUN_MAX) { ++t; ++r; }
        out.push_back(run_word(id, r));
    }
}

// one left-to-right pass: (a with no run, b) -> id keeping b's run; in place, returns the new length
static inline size_t run_replace(uint32_t* s, size_t n, uint32_t a, uint32_t b, uint32_t id) {
    size_t w = 0, t = 0;
    while (t < n) {
        if (t + 1 < n && s[t] == a && (s[t + 1] & 0xffu) == b) { s[w++] = (s[t + 1] & ~0xffu) | id; t += 2; }
        else { s[w++] = s[t++]; }
    }
    return w;
}

static ZipMerges run_learn(std::vector<std::vector<uint32_t>>& sample, int nsym, int max_ids, long long min_count) {
    ZipMerges mg;
    mg.nsym = nsym;
    if (max_ids > 256) max_ids = 256;
    std::vector<long long> counts;
    int ns = nsym;
    while (ns < max_ids) {
        counts.assign((size_t)ns * ns, 0);
        for (const auto& s : sample)
            for (size_t t = 0; t + 1 < s.size(); ++t)
                if (s[t] < 256u) counts[(size_t)s[t] * ns + (s[t + 1] & 0xffu)]++;      // s[t] < 256: no run attached
        long long best = 0;
        int ba = 0, bb = 0;
        for (int a = 0; a < ns; ++a)
            for (int b = 0; b < ns; ++b)
                if (counts[(size_t)a * ns + b] > best) { best = counts[(size_t)a * ns + b]; ba = a; bb = b; }
        if (best < min_count) break;
        for (auto& s : sample) s.resize(run_replace(s.data(), s.size(), (uint32_t)ba, (uint32_t)bb, (uint32_t)ns));
        mg.pairs.push_back({(uint8_t)ba, (uint8_t)bb});
        ++ns;
    }
    return mg;
}

static void run_encode(const ZipMerges& mg, const uint8_t* sym, size_t n, int run_sym, int* first_run, std::vector<uint32_t>& out) {
    run_tokenize(sym, n, run_sym, first_run, out);
    size_t len = out.size();
    for (size_t i = 0; i < mg.pairs.size() && len >= 2; ++i)
        len = run_replace(out.data(), len, mg.pairs[i][0], mg.pairs[i][1], (uint32_t)(mg.nsym + i));
    out.resize(len);
    out.shrink_to_fit();
}

// run tokens over the first M ids only (larger ids expanded into their parts, the run stays with the last part)
static void run_expand(const ZipMerges& mg, const std::vector<uint32_t>& in, int M, std::vector<uint32_t>& out) {
    out.clear();
    out.reserve(in.size());
    uint8_t stack[512];
    for (uint32_t w : in) {
        const int id = (int)(w & 0xffu);
        if (id < M) { out.push_back(w); continue; }
        int sp = 0;
        stack[sp++] = (uint8_t)id;
        while (sp > 0) {
            const uint8_t x = stack[--sp];
            if (x < M) { out.push_back(x); continue; }
            const auto& p = mg.pairs[x - mg.nsym];
            stack[sp++] = p[1];
            stack[sp++] = p[0];
        }
        out.back() |= w & ~0xffu;
    }
}

// ------------------------------------------------------------------------------------------------
// Two-run form.  Where shared memory holds only a dozen dictionary entries (K >= 32), a block of the SECOND run symbol r2
// (missing data: runs of mean length 100) costs three or four power-of-two entries.  With C_r2 diagonalised as well
// (zip_spectral_kernel), a run of m >= RUN2_MIN sites of r2 becomes two tokens over two FIXED extra base entries,
//     (nsym     | m << 8 | RUN_TABLE2_BIT)   "B^-1, then Lambda2^m"      and      (nsym + 1 | n << 8)   "B, then n sites of r",
// whatever m is; shorter runs of r2 stay ordinary entries.  The alphabet of this encoding is nsym + 2 base ids.
// ------------------------------------------------------------------------------------------------
constexpr int RUN2_MIN = 3;

static void run2_tokenize(const uint8_t* sym, size_t n, int nsym, int run_sym, int run_sym2, int* first_run, std::vector<uint32_t>& out,
                          long long* run2_sites) {
    out.clear();
    *run2_sites = 0;
    size_t t = 0;
    while (t < n && sym[t] == run_sym && (int)t < RUN_MAX) ++t;
    *first_run = (int)t;
    auto take_run = [&]() { int r = 0; while (t < n && sym[t] == run_sym && r < RUN_MAX) { ++t; ++r; } return r; };
    while (t < n) {
        if (sym[t] == run_sym2) {
            size_t e = t;
            while (e < n && sym[e] == run_sym2) ++e;
            size_t m = e - t;
            if (m >= (size_t)RUN2_MIN) {
                while (m > 0) {                 // blocks longer than RUN_MAX: several (B^-1, B) pairs
                    const int mm = (int)std::min<size_t>(m, RUN_MAX);
                    out.push_back(run_word(nsym, mm) | RUN_TABLE2_BIT);
                    *run2_sites += mm;
                    m -= mm;
                    t += mm;
                    out.push_back(run_word(nsym + 1, m == 0 ? take_run() : 0));
                }
                continue;
            }
        }
        const int id = sym[t++];
        out.push_back(run_word(id, take_run()));
    }
}

static void run2_encode(const ZipMerges& mg, const uint8_t* sym, size_t n, int nsym, int run_sym, int run_sym2, int* first_run,
                        std::vector<uint32_t>& out, long long* run2_sites) {
    run2_tokenize(sym, n, nsym, run_sym, run_sym2, first_run, out, run2_sites);
    size_t len = out.size();
    for (size_t i = 0; i < mg.pairs.size() && len >= 2; ++i)
        len = run_replace(out.data(), len, mg.pairs[i][0], mg.pairs[i][1], (uint32_t)(mg.nsym + i));
    out.resize(len);
    out.shrink_to_fit();
}

// fn(i) for i in [0, n) on the host cores of the affinity mask; returns false if any call threw (out of memory)
template <typename F>
static bool parallel_for(int n, F&& fn) {
    unsigned hw = std::thread::hardware_concurrency();
    int nt = (int)std::min<unsigned>(hw ? hw : 4u, 32u);
    cpu_set_t cs;
    if (sched_getaffinity(0, sizeof cs, &cs) == 0) nt = std::min(nt, std::max(1, CPU_COUNT(&cs)));
    nt = std::min(nt, n);
    std::atomic<bool> ok{true};
    auto guarded = [&](int i) { try { fn(i); } catch (...) { ok = false; } };
    if (nt <= 1) { for (int i = 0; i < n; ++i) guarded(i); return ok; }
    std::atomic<int> next{0};
    std::vector<std::thread> th;
    for (int t = 0; t < nt; ++t)
        th.emplace_back([&]() { for (int i = next++; i < n; i = next++) guarded(i); });
    for (auto& t : th) t.join();
    return ok;
}

}  // namespace imc
