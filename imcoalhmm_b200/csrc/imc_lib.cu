// libimcoalhmm_b200.so -- host side: sequences, packed sequence sets, launchers and the C ABI
// declared in include/imcoalhmm_b200.h.  No CPU fallback: every forward entry point needs the GPU.
#include "../../include/imcoalhmm_b200.h"
#include "forward_kernels.cuh"
#include "zip_kernels.cuh"
#include "model_kernels.cuh"

#include <algorithm>
#include <atomic>
#include <cerrno>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <mutex>
#include <numeric>
#include <dlfcn.h>
#include <nccl.h>      // types and prototypes only: the library is resolved with dlopen when imc_comm_init is called
#include <sched.h>
#include <string>
#include <unistd.h>
#include <vector>

#include "tokenizer.inl"

using namespace imc;

// ------------------------------------------------------------------------------------------ errors
static thread_local std::string g_err;
static thread_local const char* g_last_kernel = "none";
static std::atomic<long long> g_launches{0};

static int fail(int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}
#define CUDA_TRY(x)                                                                              \
    do {                                                                                         \
        cudaError_t e_ = (x);                                                                    \
        if (e_ != cudaSuccess) return fail(IMC_ERR_CUDA, "%s failed: %s", #x, cudaGetErrorString(e_)); \
    } while (0)

// ------------------------------------------------------------------------------------------ context
struct Context {
    std::mutex mu;
    int device = -1;          // requested device (-1 = default 0)
    bool ready = false;
    pid_t pid = 0;
    cudaStream_t stream = nullptr;
    int sm_count = 0;
    long long opt_forward_kernel = 0;
    long long opt_dmma_mtiles = 0;
    long long opt_fold_emission = 0;   // measured slower than the emission-row multiply on B200 (profiles/r01_pair_micro2.txt)
    long long opt_zip_ctas_per_sm = 0; // 1 or 2 resident CTAs per SM for the zip kernel (0 = auto)
    long long opt_zip_lanes = 0;       // lanes per chain in the zip kernel: 8, 4 or 0 = auto
    long long opt_zip_segment_tokens = 0;  // tokens per segment in segmented mode: 0 = auto, -1 = never, > 0 = forced
    long long opt_zip_max_entries = 0; // cap on dictionary entries used (0 = whatever fits in shared memory)
};
static Context g_ctx;

static int ensure_device() {
    std::lock_guard<std::mutex> lk(g_ctx.mu);
    if (g_ctx.ready) {
        if (g_ctx.pid != getpid())
            return fail(IMC_ERR_CUDA, "CUDA was initialised in the parent before fork(); create the context in the "
                                      "child (construct Forwarders before forking, call forward only in children, "
                                      "or use the 'spawn' start method)");
        return IMC_OK;
    }
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(IMC_ERR_CUDA, "no usable CUDA device (%s); this library has no CPU fallback",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    const int dev = g_ctx.device < 0 ? 0 : g_ctx.device;
    if (dev >= count) return fail(IMC_ERR_INVALID, "device %d requested but only %d present", dev, count);
    CUDA_TRY(cudaSetDevice(dev));
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, dev));
    if (prop.major != 10)
        return fail(IMC_ERR_CUDA, "device %d is sm_%d%d; this build contains sm_100a code only", dev, prop.major, prop.minor);
    g_ctx.sm_count = prop.multiProcessorCount;
    CUDA_TRY(cudaStreamCreateWithFlags(&g_ctx.stream, cudaStreamNonBlocking));
    g_ctx.device = dev;
    g_ctx.pid = getpid();
    g_ctx.ready = true;
    return IMC_OK;
}

// ------------------------------------------------------------------------------------------ sequences
struct imc_seq {
    std::vector<uint8_t> sym;
    int nsym = 0;
};

struct DeviceBuf {
    void* p = nullptr;
    size_t cap = 0;
    int reserve(size_t bytes) {
        if (bytes <= cap) return IMC_OK;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        CUDA_TRY(cudaMalloc(&p, bytes));
        cap = bytes;
        return IMC_OK;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

// device copy of the token streams derived for one dictionary size M (see zip_device)
struct ZipSplit {                 // segmented variant of a ZipDevice's chunk list (same token buffer)
    int K = 0, seglen = 0, nchains = 0;
    int n_level1 = 0, n_final = 0, nvec2 = 0;     // fold items of the two levels, vectors written by level 1
    DeviceBuf chunks, items1, items2;
};

struct ZipDevice {
    int M = 0, nlevels = 0;
    long long total_tokens = 0;
    int max_ntok = 0;
    std::vector<ZipChunk> host_chunks;        // sorted by ntok, descending
    DeviceBuf tokens, chunks, pairs, levels;
    std::vector<ZipSplit*> splits;
};

struct imc_seqset {
    int n_chunks = 0;
    int nsym = 0;
    long long total_sites = 0;
    int fold_sym = 0;                     // most frequent symbol over the whole set
    // host-side packed layout (plain kernels)
    bool packable = false;                // nsym <= 3
    std::vector<uint32_t> words;          // bundles of 32 streams, word-interleaved, 16 two-bit symbols per word (lazy)
    long long n_words = 0;
    std::vector<StreamInfo> streams;      // non-empty chunks only, sorted by length (descending)
    // host-side compressed layout (zip kernel): one dictionary for the whole set, tokens per non-empty chunk
    ZipMerges merges;
    std::vector<std::vector<uint8_t>> tok_full;   // per stream, over all merges.size() ids, symbols 1..L-1
    std::vector<uint8_t> first_sym;               // per stream
    std::vector<int> stream_of_chunk;             // chunk index as given to imc_seqset_create -> stream (-1: empty chunk)
    // device side (lazy)
    bool uploaded = false;
    DeviceBuf d_words, d_streams, d_chain, d_pi, d_T, d_E, d_out, d_pnext, d_vec;
    std::vector<ZipDevice*> zip_dev;      // one per dictionary size in use
};

static int seq_finish(imc_seq* s, imc_seq** out) {
    *out = s;
    return IMC_OK;
}

extern "C" const char* imc_last_error(void) { return g_err.c_str(); }
extern "C" int imc_version(void) { return 100; }

extern "C" int imc_init(int device) {
    {
        std::lock_guard<std::mutex> lk(g_ctx.mu);
        if (g_ctx.ready && g_ctx.device != device && g_ctx.pid == getpid())
            return fail(IMC_ERR_INVALID, "context already bound to device %d", g_ctx.device);
        if (!g_ctx.ready) g_ctx.device = device;
    }
    return ensure_device();
}

extern "C" int imc_device_count(int* count_out) {
    if (!count_out) return fail(IMC_ERR_INVALID, "count_out is NULL");
    int c = 0;
    cudaError_t e = cudaGetDeviceCount(&c);
    *count_out = (e == cudaSuccess) ? c : 0;
    return IMC_OK;
}

template <typename Tin>
static int seq_create_impl(const Tin* obs, int64_t L, int nsym, imc_seq** out) {
    if (!out) return fail(IMC_ERR_INVALID, "out is NULL");
    if (L < 0 || (L > 0 && !obs)) return fail(IMC_ERR_INVALID, "bad observation buffer");
    if (nsym < 1 || nsym > 255) return fail(IMC_ERR_INVALID, "nsym must be in [1, 255], got %d", nsym);
    imc_seq* s = new (std::nothrow) imc_seq;
    if (!s) return fail(IMC_ERR_NOMEM, "out of memory");
    s->nsym = nsym;
    try { s->sym.resize((size_t)L); } catch (...) { delete s; return fail(IMC_ERR_NOMEM, "out of memory for %lld symbols", (long long)L); }
    for (int64_t t = 0; t < L; ++t) {
        const long long v = (long long)obs[t];
        if (v < 0 || v >= nsym) {
            delete s;
            return fail(IMC_ERR_INVALID, "symbol %lld at position %lld is outside [0, %d)", v, (long long)t, nsym);
        }
        s->sym[(size_t)t] = (uint8_t)v;
    }
    return seq_finish(s, out);
}

extern "C" int imc_seq_create(const int32_t* obs, int64_t L, int nsym, imc_seq** out) {
    return seq_create_impl(obs, L, nsym, out);
}
extern "C" int imc_seq_create_u8(const uint8_t* obs, int64_t L, int nsym, imc_seq** out) {
    return seq_create_impl(obs, L, nsym, out);
}

extern "C" int imc_seq_from_file(const char* path, int nsym, imc_seq** out) {
    if (!path || !out) return fail(IMC_ERR_INVALID, "NULL argument");
    if (nsym < 1 || nsym > 255) return fail(IMC_ERR_INVALID, "nsym must be in [1, 255], got %d", nsym);
    FILE* f = fopen(path, "rb");
    if (!f) return fail(IMC_ERR_IO, "cannot open '%s': %s", path, strerror(errno));
    imc_seq* s = new (std::nothrow) imc_seq;
    if (!s) { fclose(f); return fail(IMC_ERR_NOMEM, "out of memory"); }
    s->nsym = nsym;
    // whitespace-separated base-10 integers (python: map(int, text.split()), hmm.py:13-14)
    std::vector<char> buf(1 << 20);
    long long cur = 0;
    bool in_num = false, neg = false;
    int rc = IMC_OK;
    size_t got;
    long long pos = 0;
    auto flush = [&]() -> int {
        if (!in_num) return IMC_OK;
        const long long v = neg ? -cur : cur;
        if (v < 0 || v >= nsym) return fail(IMC_ERR_INVALID, "symbol %lld (token %zu) in '%s' is outside [0, %d)", v, s->sym.size(), path, nsym);
        s->sym.push_back((uint8_t)v);
        in_num = false; neg = false; cur = 0;
        return IMC_OK;
    };
    while (rc == IMC_OK && (got = fread(buf.data(), 1, buf.size(), f)) > 0) {
        for (size_t i = 0; i < got && rc == IMC_OK; ++i, ++pos) {
            const char ch = buf[i];
            if (ch >= '0' && ch <= '9') {
                cur = cur * 10 + (ch - '0');
                if (cur > 1000000) cur = 1000000;  // saturate; rejected by the range check
                in_num = true;
            } else if (ch == ' ' || ch == '\n' || ch == '\t' || ch == '\r' || ch == '\f' || ch == '\v') {
                rc = flush();
            } else if ((ch == '-' || ch == '+') && !in_num) {
                neg = (ch == '-'); in_num = true;
            } else {
                rc = fail(IMC_ERR_IO, "unexpected character 0x%02x at byte %lld of '%s'", (unsigned char)ch, pos, path);
            }
        }
    }
    if (rc == IMC_OK) rc = flush();
    fclose(f);
    if (rc != IMC_OK) { delete s; return rc; }
    return seq_finish(s, out);
}

// ------------------------------------------------------------------------------------------ ingest
// Pairwise symbol rule of the reference's preprocessing script (scripts/prepare-alignments.py:99-111):
//   upper-case both bases; 2 if either is not one of A, C, G, T; 0 if equal; 1 otherwise.
static inline uint8_t pair_symbol(unsigned char a, unsigned char b) {
    static const struct Table { uint8_t code[256]; Table() {
        for (int i = 0; i < 256; ++i) code[i] = 4;
        code[(int)'A'] = code[(int)'a'] = 0; code[(int)'C'] = code[(int)'c'] = 1;
        code[(int)'G'] = code[(int)'g'] = 2; code[(int)'T'] = code[(int)'t'] = 3;
    } } tab;
    const uint8_t x = tab.code[a], y = tab.code[b];
    return (x > 3 || y > 3) ? 2 : (x == y ? 0 : 1);
}

extern "C" int imc_seq_from_pair(const char* seq1, const char* seq2, int64_t L, imc_seq** out) {
    if (!out || L < 0 || (L > 0 && (!seq1 || !seq2))) return fail(IMC_ERR_INVALID, "bad arguments");
    imc_seq* s = new (std::nothrow) imc_seq;
    if (!s) return fail(IMC_ERR_NOMEM, "out of memory");
    s->nsym = 3;
    try { s->sym.resize((size_t)L); } catch (...) { delete s; return fail(IMC_ERR_NOMEM, "out of memory for %lld symbols", (long long)L); }
    for (int64_t t = 0; t < L; ++t) s->sym[(size_t)t] = pair_symbol((unsigned char)seq1[t], (unsigned char)seq2[t]);
    return seq_finish(s, out);
}

// FASTA: '>' starts a record, its name is the text up to the first whitespace; sequence lines are concatenated.
static int read_fasta(const char* path, std::vector<std::string>& names, std::vector<std::string>& seqs) {
    FILE* f = fopen(path, "rb");
    if (!f) return fail(IMC_ERR_IO, "cannot open '%s': %s", path, strerror(errno));
    std::vector<char> buf(1 << 20);
    bool in_header = false, at_line_start = true;
    size_t got;
    while ((got = fread(buf.data(), 1, buf.size(), f)) > 0) {
        for (size_t i = 0; i < got; ++i) {
            const char ch = buf[i];
            if (in_header) {
                if (ch == '\n') { in_header = false; at_line_start = true; }
                else names.back().push_back(ch);
                continue;
            }
            if (ch == '\n' || ch == '\r') { at_line_start = (ch == '\n') || at_line_start; continue; }
            if (at_line_start && ch == '>') { names.emplace_back(); seqs.emplace_back(); in_header = true; continue; }
            at_line_start = false;
            if (ch == ' ' || ch == '\t') continue;
            if (seqs.empty()) { fclose(f); return fail(IMC_ERR_IO, "'%s' does not start with a FASTA header", path); }
            seqs.back().push_back(ch);
        }
    }
    fclose(f);
    for (auto& n : names) {
        while (!n.empty() && (n.back() == '\r' || n.back() == ' ' || n.back() == '\t')) n.pop_back();
        const size_t sp = n.find_first_of(" \t");
        if (sp != std::string::npos) n.resize(sp);
    }
    return IMC_OK;
}

extern "C" int imc_seq_from_fasta(const char* path, const char* name1, const char* name2, imc_seq** out) {
    if (!path || !out) return fail(IMC_ERR_INVALID, "NULL argument");
    std::vector<std::string> names, seqs;
    int rc = read_fasta(path, names, seqs);
    if (rc) return rc;
    int i1 = -1, i2 = -1;
    if (!name1 && !name2) {
        if (names.size() != 2) return fail(IMC_ERR_INVALID, "'%s' holds %zu records; name the two to compare", path, names.size());
        i1 = 0; i2 = 1;
    } else {
        if (!name1 || !name2) return fail(IMC_ERR_INVALID, "give both record names or neither");
        for (size_t i = 0; i < names.size(); ++i) { if (names[i] == name1) i1 = (int)i; if (names[i] == name2) i2 = (int)i; }
        if (i1 < 0 || i2 < 0) return fail(IMC_ERR_INVALID, "record '%s' not found in '%s'", i1 < 0 ? name1 : name2, path);
    }
    if (seqs[i1].size() != seqs[i2].size())
        return fail(IMC_ERR_INVALID, "aligned sequences differ in length (%zu vs %zu)", seqs[i1].size(), seqs[i2].size());
    return imc_seq_from_pair(seqs[i1].data(), seqs[i2].data(), (int64_t)seqs[i1].size(), out);
}

// Binary container: "IMCSEQ1\0", int32 nsym, int32 bits per symbol (2 or 8), int64 L, packed symbols (little endian,
// symbol t of a 2-bit file sits in bits 2*(t%4).. of byte t/4).  16x smaller than the text format for NSYM = 3.
extern "C" int imc_seq_save(const imc_seq* seq, const char* path) {
    if (!seq || !path) return fail(IMC_ERR_INVALID, "NULL argument");
    FILE* f = fopen(path, "wb");
    if (!f) return fail(IMC_ERR_IO, "cannot create '%s': %s", path, strerror(errno));
    const int32_t nsym = seq->nsym, bits = seq->nsym <= 4 ? 2 : 8;
    const int64_t L = (int64_t)seq->sym.size();
    bool ok = fwrite("IMCSEQ1", 1, 8, f) == 8 && fwrite(&nsym, 4, 1, f) == 1 && fwrite(&bits, 4, 1, f) == 1 && fwrite(&L, 8, 1, f) == 1;
    if (ok && bits == 8) ok = L == 0 || fwrite(seq->sym.data(), 1, (size_t)L, f) == (size_t)L;
    if (ok && bits == 2) {
        std::vector<uint8_t> packed((size_t)((L + 3) / 4), 0);
        for (int64_t t = 0; t < L; ++t) packed[(size_t)(t >> 2)] |= (uint8_t)(seq->sym[(size_t)t] << (2 * (t & 3)));
        ok = packed.empty() || fwrite(packed.data(), 1, packed.size(), f) == packed.size();
    }
    if (fclose(f) != 0) ok = false;
    return ok ? IMC_OK : fail(IMC_ERR_IO, "short write to '%s'", path);
}

extern "C" int imc_seq_load(const char* path, imc_seq** out) {
    if (!path || !out) return fail(IMC_ERR_INVALID, "NULL argument");
    FILE* f = fopen(path, "rb");
    if (!f) return fail(IMC_ERR_IO, "cannot open '%s': %s", path, strerror(errno));
    char magic[8];
    int32_t nsym = 0, bits = 0;
    int64_t L = -1;
    bool ok = fread(magic, 1, 8, f) == 8 && !memcmp(magic, "IMCSEQ1", 8) && fread(&nsym, 4, 1, f) == 1 && fread(&bits, 4, 1, f) == 1 &&
              fread(&L, 8, 1, f) == 1 && nsym >= 1 && nsym <= 255 && (bits == 2 || bits == 8) && L >= 0 && !(bits == 2 && nsym > 4);
    if (!ok) { fclose(f); return fail(IMC_ERR_IO, "'%s' is not an IMCSEQ1 file", path); }
    imc_seq* s = new (std::nothrow) imc_seq;
    if (!s) { fclose(f); return fail(IMC_ERR_NOMEM, "out of memory"); }
    s->nsym = nsym;
    try {
        s->sym.resize((size_t)L);
        if (bits == 8) ok = L == 0 || fread(s->sym.data(), 1, (size_t)L, f) == (size_t)L;
        else {
            std::vector<uint8_t> packed((size_t)((L + 3) / 4));
            ok = packed.empty() || fread(packed.data(), 1, packed.size(), f) == packed.size();
            for (int64_t t = 0; ok && t < L; ++t) s->sym[(size_t)t] = (packed[(size_t)(t >> 2)] >> (2 * (t & 3))) & 3;
        }
    } catch (...) { fclose(f); delete s; return fail(IMC_ERR_NOMEM, "out of memory for %lld symbols", (long long)L); }
    fclose(f);
    if (ok) for (uint8_t v : s->sym) if (v >= nsym) { ok = false; break; }
    if (!ok) { delete s; return fail(IMC_ERR_IO, "'%s' is truncated or holds symbols outside [0, %d)", path, nsym); }
    return seq_finish(s, out);
}

extern "C" int imc_seq_length(const imc_seq* seq, int64_t* L_out) {
    if (!seq || !L_out) return fail(IMC_ERR_INVALID, "NULL argument");
    *L_out = (int64_t)seq->sym.size();
    return IMC_OK;
}
extern "C" int imc_seq_nsym(const imc_seq* seq, int* nsym_out) {
    if (!seq || !nsym_out) return fail(IMC_ERR_INVALID, "NULL argument");
    *nsym_out = seq->nsym;
    return IMC_OK;
}
extern "C" int imc_seq_symbol_counts(const imc_seq* seq, int64_t* counts) {
    if (!seq || !counts) return fail(IMC_ERR_INVALID, "NULL argument");
    for (int i = 0; i < seq->nsym; ++i) counts[i] = 0;
    for (uint8_t v : seq->sym) counts[v]++;
    return IMC_OK;
}
extern "C" int imc_seq_symbols(const imc_seq* seq, uint8_t* out, int64_t capacity) {
    if (!seq || !out) return fail(IMC_ERR_INVALID, "NULL argument");
    if (capacity < (int64_t)seq->sym.size()) return fail(IMC_ERR_INVALID, "capacity %lld < length %zu", (long long)capacity, seq->sym.size());
    memcpy(out, seq->sym.data(), seq->sym.size());
    return IMC_OK;
}
extern "C" int imc_seq_destroy(imc_seq* seq) {
    delete seq;
    return IMC_OK;
}

// ------------------------------------------------------------------------------------------ sequence sets
extern "C" int imc_seqset_create(const imc_seq* const* seqs, int C, imc_seqset** out) {
    if (!out || C < 0 || (C > 0 && !seqs)) return fail(IMC_ERR_INVALID, "bad arguments");
    imc_seqset* set = new (std::nothrow) imc_seqset;
    if (!set) return fail(IMC_ERR_NOMEM, "out of memory");
    set->n_chunks = C;
    std::vector<int> order;
    for (int c = 0; c < C; ++c) {
        if (!seqs[c]) { delete set; return fail(IMC_ERR_INVALID, "seqs[%d] is NULL", c); }
        if (c == 0) set->nsym = seqs[c]->nsym;
        if (seqs[c]->nsym != set->nsym) { delete set; return fail(IMC_ERR_INVALID, "chunks disagree on nsym (%d vs %d)", seqs[c]->nsym, set->nsym); }
        if (seqs[c]->sym.size() > 0x7fffffffULL) { delete set; return fail(IMC_ERR_UNSUPPORTED, "chunk %d has more than 2^31-1 sites; split it", c); }
        set->total_sites += (long long)seqs[c]->sym.size();
        if (!seqs[c]->sym.empty()) order.push_back(c);   // an empty chunk contributes logL = 0
    }
    set->packable = set->nsym <= 3;
    {
        long long counts[256] = {0};
        for (int c : order) for (uint8_t v : seqs[c]->sym) counts[v]++;
        set->fold_sym = (int)(std::max_element(counts, counts + set->nsym) - counts);
    }
    std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return seqs[x]->sym.size() > seqs[y]->sym.size(); });
    const int ns = (int)order.size();
    set->streams.resize(ns);
    long long word_off = 0;   // in words
    try {
        // ---- zipHMM-style preprocessing (hmm.py:16): learn the merges on a bounded prefix sample, encode every chunk
        {
            // sample: up to 16 M symbols taken from chunks spread evenly over the set (not just the first ones)
            std::vector<std::vector<uint8_t>> sample;
            const long long budget_total = 16LL << 20;
            long long budget = budget_total;
            const int stride = std::max(1, C / 32);
            for (int start = 0; start < stride && budget > 0; ++start)
                for (int c = start; c < C && budget > 0; c += stride) {
                    const auto& sy = seqs[c]->sym;
                    if (sy.size() < 2) continue;
                    const long long share = budget_total / std::min(32, std::max(1, C));
                    const size_t take = (size_t)std::min<long long>({(long long)sy.size() - 1, budget, share});
                    sample.emplace_back(sy.begin() + 1, sy.begin() + 1 + take);
                    budget -= (long long)take;
                }
            set->merges = zip_learn(sample, set->nsym, 256, 16);
        }
        set->tok_full.resize(ns);
        set->first_sym.resize(ns);
        set->stream_of_chunk.assign(C, -1);
        for (int k = 0; k < ns; ++k) set->stream_of_chunk[order[k]] = k;
        if (!parallel_for(ns, [&](int k) {
                const auto& sy = seqs[order[k]]->sym;
                set->first_sym[k] = sy[0];
                zip_encode(set->merges, sy.data() + 1, sy.size() - 1, set->tok_full[k]);
            })) throw std::bad_alloc();
        // stream geometry of the packed 2-bit layout; the words themselves are built on first use (seqset_pack)
        for (int b0 = 0; set->packable && b0 < ns; b0 += 32) {
            const int nb = std::min(32, ns - b0);
            const long long nwords = ((long long)seqs[order[b0]]->sym.size() + 15) / 16;
            for (int k = 0; k < nb; ++k) {
                const long long len = (long long)seqs[order[b0 + k]]->sym.size();
                StreamInfo& si = set->streams[b0 + k];
                si.base = word_off + k;
                si.len = (int)len;
                si.nwords = (int)((len + 15) / 16);
            }
            word_off += nwords * 32;
        }
        set->n_words = word_off;
    } catch (...) { delete set; return fail(IMC_ERR_NOMEM, "out of host memory while preprocessing"); }
    *out = set;
    return IMC_OK;
}

extern "C" int imc_seqset_destroy(imc_seqset* set) {
    if (!set) return IMC_OK;
    const bool mine = g_ctx.pid == getpid();
    if (mine) {
        set->d_words.release(); set->d_streams.release(); set->d_chain.release();
        set->d_pi.release(); set->d_T.release(); set->d_E.release(); set->d_out.release(); set->d_pnext.release();
        set->d_vec.release();
    }
    for (ZipDevice* z : set->zip_dev) {
        if (mine) { z->tokens.release(); z->chunks.release(); z->pairs.release(); z->levels.release(); }
        for (ZipSplit* sp : z->splits) {
            if (mine) { sp->chunks.release(); sp->items1.release(); sp->items2.release(); }
            delete sp;
        }
        delete z;
    }
    delete set;
    return IMC_OK;
}

extern "C" int imc_seqset_info(const imc_seqset* set, int* n_chunks, int64_t* total_sites, int64_t* packed_bytes) {
    if (!set) return fail(IMC_ERR_INVALID, "NULL set");
    if (n_chunks) *n_chunks = set->n_chunks;
    if (total_sites) *total_sites = set->total_sites;
    if (packed_bytes) *packed_bytes = (int64_t)(set->n_words * (long long)sizeof(uint32_t));
    return IMC_OK;
}

// The packed layout of the per-site kernels, rebuilt from the token streams (the set keeps no copy of the raw symbols):
// word w of stream k sits at base + 32 w, symbol t of the word in bits 2 (t % 16)..; code 3 = padding.
static int seqset_pack(imc_seqset* set) {
    if (!set->words.empty() || set->n_words == 0) return IMC_OK;
    try {
        set->words.assign((size_t)set->n_words, 0xffffffffu);
        const int ns = (int)set->streams.size();
        if (!parallel_for(ns, [&](int k) {
                std::vector<uint8_t> sym;
                zip_expand(set->merges, set->tok_full[k], set->nsym, sym);
                const StreamInfo& si = set->streams[k];
                auto put = [&](long long t, uint8_t v) {
                    uint32_t& w = set->words[(size_t)(si.base + (t >> 4) * 32)];
                    const int sh = 2 * (int)(t & 15);
                    w = (w & ~(3u << sh)) | ((uint32_t)v << sh);
                };
                put(0, set->first_sym[k]);
                for (size_t t = 0; t < sym.size(); ++t) put((long long)t + 1, sym[t]);
            })) throw std::bad_alloc();
    } catch (...) { set->words.clear(); return fail(IMC_ERR_NOMEM, "out of host memory while packing"); }
    return IMC_OK;
}

static int seqset_upload(imc_seqset* set) {
    if (set->uploaded) return IMC_OK;
    int rc = ensure_device();
    if (rc) return rc;
    if ((rc = seqset_pack(set))) return rc;
    if (!set->streams.empty()) {
        if ((rc = set->d_words.reserve(set->words.size() * sizeof(uint32_t)))) return rc;
        if ((rc = set->d_streams.reserve(set->streams.size() * sizeof(StreamInfo)))) return rc;
        CUDA_TRY(cudaMemcpy(set->d_words.p, set->words.data(), set->words.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
        CUDA_TRY(cudaMemcpy(set->d_streams.p, set->streams.data(), set->streams.size() * sizeof(StreamInfo), cudaMemcpyHostToDevice));
        std::vector<uint32_t>().swap(set->words);      // the device copy is the only one needed from here on
    }
    set->uploaded = true;
    return IMC_OK;
}


// ------------------------------------------------------------------------------------------ zip (compressed) path
enum { KERNEL_AUTO = 0, KERNEL_GENERIC = 1, KERNEL_PAIR = 2, KERNEL_DMMA = 3, KERNEL_ZIP = 4 };

#define ZIP_K_LIST(X) X(2) X(3) X(4) X(5) X(6) X(8) X(10) X(12) X(16) X(20) X(24) X(32) X(40)
// smallest instantiated tile that holds K states (the kernels take the actual K at run time and leave the padding
// rows / columns of the tile at zero), or 0
static int zip_tile(int K) {
    static const int tiles[] = {
#define X(k) k,
        ZIP_K_LIST(X)
#undef X
    };
    for (int t : tiles) if (t >= K) return t;
    return 0;
}
static bool zip_supported(int K) { return K >= 1 && zip_tile(K) != 0; }

struct ZipPlan { int lanes, threads, ctas_per_sm, M; size_t smem; };
static const size_t ZIP_SMEM_SM = 227 * 1024;    // usable shared memory per SM (1 KB per resident CTA is reserved on top)

// Launch shapes (all persistent, see zip_forward_kernel):
//   lanes per chain 8: two CTAs of 256 threads per SM (K <= 24), or one CTA with all the shared memory for the
//                      dictionary -- 512 threads where the register file allows (K <= 24), else 256;
//   lanes per chain 4: one or (K <= 24) two CTAs of 256 threads (64 chains per CTA), K >= 8 only.
// measured on B200 (gpurun_out/zip_bench_*.log, round 1): K=10 8.3 ms (8 lanes) vs 9.7 ms (4, padded to 12);
// K=20 58.9 vs 43.7 ms; K=40 175 vs 190 ms (4 lanes need twice the exchange buffers, which costs dictionary entries)
// (tools/k_sweep.py: tile 8 1.25 vs 1.34 ms with 4 lanes; tile 10 keeps 8 lanes, where one full load fetches the remainder
// rows of four tokens)
static int zip_default_lanes(int K) { return (K >= 8 && K <= 24 && K != 10) ? 4 : 8; }

template <int K>
static ZipPlan zip_plan_k(int S, int avail_ids, int want_ctas, int want_lanes) {
    ZipPlan p;
    p.lanes = want_lanes ? want_lanes : zip_default_lanes(K);
    if (K < 8) p.lanes = 8;
    if (p.lanes == 32 && K < 10) p.lanes = 8;
    int m1, m2, t1;
    if (p.lanes == 32) {       // one chain per warp, one CTA per SM (latency mode)
        using C = ZipCfg32<K>;
        t1 = K <= 24 ? 512 : 256;
        m2 = 0;
        m1 = ZipSmem<C>::max_entries(ZIP_SMEM_SM, S, t1);
        want_ctas = 1;
    } else if (p.lanes == 8) {
        using C = ZipCfg8<K>;
        t1 = K <= 24 ? 512 : 256;
        m2 = K <= 24 ? ZipSmem<C>::max_entries((ZIP_SMEM_SM - 1024) / 2, S, 256) : 0;
        m1 = ZipSmem<C>::max_entries(ZIP_SMEM_SM, S, t1);
    } else {
        using C = ZipCfg4<K>;
        t1 = 256;
        m2 = K <= 24 ? ZipSmem<C>::max_entries((ZIP_SMEM_SM - 1024) / 2, S, 256) : 0;
        m1 = ZipSmem<C>::max_entries(ZIP_SMEM_SM, S, t1);
    }
    int ctas = want_ctas;
    if (ctas == 2 && K > 24) ctas = 1;
    if (ctas == 0) ctas = (K <= 24 && m2 >= avail_ids) ? 2 : 1;   // the bigger dictionary wins unless everything fits in half
    p.ctas_per_sm = ctas;
    p.threads = ctas == 2 ? 256 : t1;
    p.M = std::min(avail_ids, ctas == 2 ? m2 : m1);
    if (g_ctx.opt_zip_max_entries > 0) p.M = std::min<int>(p.M, (int)g_ctx.opt_zip_max_entries);
    p.M = std::max(p.M, S);
    p.smem = p.lanes == 8 ? ZipSmem<ZipCfg8<K>>::bytes(p.M, S, p.threads)
           : (p.lanes == 4 ? ZipSmem<ZipCfg4<K>>::bytes(p.M, S, p.threads) : ZipSmem<ZipCfg32<K>>::bytes(p.M, S, p.threads));
    return p;
}

static int zip_plan(int K, int S, int avail_ids, ZipPlan* out, int lanes_override = 0) {
    const int want = (int)g_ctx.opt_zip_ctas_per_sm, lanes = lanes_override ? lanes_override : (int)g_ctx.opt_zip_lanes;
    switch (zip_tile(K)) {
#define X(k) case k: *out = zip_plan_k<k>(S, avail_ids, want, lanes); break;
        ZIP_K_LIST(X)
#undef X
        default: return fail(IMC_ERR_UNSUPPORTED, "zip kernel is not instantiated for K = %d", K);
    }
    if (out->smem > ZIP_SMEM_SM) return fail(IMC_ERR_UNSUPPORTED, "zip kernel: %d symbols x K = %d do not fit in shared memory", S, K);
    return IMC_OK;
}

// token streams over the first M dictionary ids, level-ordered, on the device (cached per M)
static int zip_device(imc_seqset* set, int M, ZipDevice** out) {
    for (ZipDevice* z : set->zip_dev) if (z->M == M) { *out = z; return IMC_OK; }
    const int ns = (int)set->streams.size();
    ZipLevels zl = zip_levels(set->merges, M);
    std::vector<std::vector<uint8_t>> tok(ns);
    if (!parallel_for(ns, [&](int k) {
            zip_expand(set->merges, set->tok_full[k], M, tok[k]);
            for (auto& t : tok[k]) t = zl.perm[t];
        })) return fail(IMC_ERR_NOMEM, "out of host memory while deriving the token streams");
    std::vector<int> order(ns);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return tok[x].size() > tok[y].size(); });
    std::vector<ZipChunk> chunks(ns);
    long long off = 0;
    for (int i = 0; i < ns; ++i) {
        const int k = order[i];
        if (tok[k].size() > 0x7fffffffULL) return fail(IMC_ERR_UNSUPPORTED, "a chunk has more than 2^31-1 tokens");
        chunks[i].tok_off = off;
        chunks[i].ntok = (int)tok[k].size();
        chunks[i].first_sym = set->first_sym[k];
        chunks[i].out_index = k;
        chunks[i].pad = 0;
        off += (long long)((tok[k].size() + 15) / 16) * 16 + 16;
    }
    std::vector<uint8_t> flat((size_t)off, 0);
    long long total = 0;
    for (int i = 0; i < ns; ++i) {
        const auto& t = tok[order[i]];
        if (!t.empty()) memcpy(flat.data() + chunks[i].tok_off, t.data(), t.size());
        total += (long long)t.size();
    }
    ZipDevice* z = new (std::nothrow) ZipDevice;
    if (!z) return fail(IMC_ERR_NOMEM, "out of memory");
    z->M = M;
    z->nlevels = (int)zl.level_start.size() - 1;
    z->total_tokens = total;
    z->max_ntok = ns ? chunks[0].ntok : 0;
    z->host_chunks = chunks;
    int rc;
    if ((rc = z->tokens.reserve(std::max<size_t>(flat.size(), 16))) || (rc = z->chunks.reserve(sizeof(ZipChunk) * std::max(ns, 1))) ||
        (rc = z->pairs.reserve(std::max<size_t>(zl.pairs.size(), 16))) || (rc = z->levels.reserve(sizeof(int) * zl.level_start.size()))) {
        z->tokens.release(); z->chunks.release(); z->pairs.release(); z->levels.release();
        delete z;
        return rc;
    }
    cudaError_t e = cudaSuccess;
    if (!flat.empty()) e = cudaMemcpy(z->tokens.p, flat.data(), flat.size(), cudaMemcpyHostToDevice);
    if (e == cudaSuccess && ns) e = cudaMemcpy(z->chunks.p, chunks.data(), sizeof(ZipChunk) * ns, cudaMemcpyHostToDevice);
    if (e == cudaSuccess && !zl.pairs.empty()) e = cudaMemcpy(z->pairs.p, zl.pairs.data(), zl.pairs.size(), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(z->levels.p, zl.level_start.data(), sizeof(int) * zl.level_start.size(), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        z->tokens.release(); z->chunks.release(); z->pairs.release(); z->levels.release();
        delete z;
        return fail(IMC_ERR_CUDA, "uploading token streams failed: %s", cudaGetErrorString(e));
    }
    set->zip_dev.push_back(z);
    *out = z;
    return IMC_OK;
}

// chunk list of z cut into segments of seglen tokens (a multiple of 16) for a K-state model, cached per (K, seglen)
static int zip_split(ZipDevice* z, int K, int seglen, ZipSplit** out) {
    for (ZipSplit* sp : z->splits) if (sp->K == K && sp->seglen == seglen) { *out = sp; return IMC_OK; }
    std::vector<ZipChunk> chains;
    std::vector<ZipFoldItem> items1, items2;
    int nvec2 = 0;
    for (const ZipChunk& ch : z->host_chunks) {
        const int nseg = std::max(1, (ch.ntok + seglen - 1) / seglen);
        const int first_chain = (int)chains.size();
        for (int sg = 0; sg < nseg; ++sg) {
            ZipChunk c = ch;
            c.tok_off = ch.tok_off + (long long)sg * seglen;
            c.ntok = std::max(0, std::min(seglen, ch.ntok - sg * seglen));
            for (int col = 0; col < (sg == 0 ? 1 : K); ++col) {
                c.first_sym = sg == 0 ? ch.first_sym : -1 - col;
                c.out_index = (int)chains.size();
                chains.push_back(c);
            }
        }
        // segment s >= 1, column c sits at first_chain + 1 + (s-1)*K + c
        if (nseg <= 32) {
            items2.push_back({first_chain, first_chain + 1, nseg - 1, ch.out_index, 0, 0});
        } else {        // two levels: groups of gs segments folded in parallel, then the groups
            const int gs = (int)std::ceil(std::sqrt((double)nseg)), ngroups = (nseg + gs - 1) / gs;
            const int base2 = nvec2;
            for (int g = 0; g < ngroups; ++g) {
                const int s0 = g * gs, s1 = std::min(nseg, s0 + gs);     // segments [s0, s1)
                if (g == 0) {
                    items1.push_back({first_chain, first_chain + 1, s1 - 1, nvec2++, 0, 1});
                } else {
                    for (int col = 0; col < K; ++col)
                        items1.push_back({first_chain + 1 + (s0 - 1) * K + col, first_chain + 1 + s0 * K, s1 - s0 - 1, nvec2++, 0, 1});
                }
            }
            items2.push_back({base2, base2 + 1, ngroups - 1, ch.out_index, 1, 0});
        }
    }
    // the kernel takes chunks in list order, longest first: full segments first, tails last (out_index keeps identity)
    std::vector<ZipChunk> sorted = chains;
    std::stable_sort(sorted.begin(), sorted.end(), [](const ZipChunk& x, const ZipChunk& y) { return x.ntok > y.ntok; });
    ZipSplit* sp = new (std::nothrow) ZipSplit;
    if (!sp) return fail(IMC_ERR_NOMEM, "out of memory");
    sp->K = K; sp->seglen = seglen; sp->nchains = (int)sorted.size();
    sp->n_level1 = (int)items1.size(); sp->n_final = (int)items2.size(); sp->nvec2 = nvec2;
    int rc;
    if ((rc = sp->chunks.reserve(sizeof(ZipChunk) * sorted.size())) ||
        (rc = sp->items1.reserve(sizeof(ZipFoldItem) * std::max<size_t>(items1.size(), 1))) ||
        (rc = sp->items2.reserve(sizeof(ZipFoldItem) * std::max<size_t>(items2.size(), 1)))) {
        sp->chunks.release(); sp->items1.release(); sp->items2.release(); delete sp;
        return rc;
    }
    cudaError_t e = cudaMemcpy(sp->chunks.p, sorted.data(), sizeof(ZipChunk) * sorted.size(), cudaMemcpyHostToDevice);
    if (e == cudaSuccess && !items1.empty()) e = cudaMemcpy(sp->items1.p, items1.data(), sizeof(ZipFoldItem) * items1.size(), cudaMemcpyHostToDevice);
    if (e == cudaSuccess && !items2.empty()) e = cudaMemcpy(sp->items2.p, items2.data(), sizeof(ZipFoldItem) * items2.size(), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        sp->chunks.release(); sp->items1.release(); sp->items2.release(); delete sp;
        return fail(IMC_ERR_CUDA, "uploading segment descriptors failed: %s", cudaGetErrorString(e));
    }
    z->splits.push_back(sp);
    *out = sp;
    return IMC_OK;
}

template <class C, int THREADS, int MINB>
static int launch_zip_k(const ZipArgs& a, const ZipPlan& p, int grid, cudaStream_t st) {
    static size_t attr_max = 0;
    if (p.smem > attr_max) {
        CUDA_TRY(cudaFuncSetAttribute(zip_forward_kernel<C, THREADS, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem));
        attr_max = p.smem;
    }
    zip_forward_kernel<C, THREADS, MINB><<<grid, THREADS, p.smem, st>>>(a);
    return IMC_OK;
}

template <int K>
static int launch_zip_shape(const ZipArgs& a, const ZipPlan& p, int grid, cudaStream_t st) {
    if constexpr (K >= 10) {
        if (p.lanes == 32) {
            if constexpr (K <= 24) return launch_zip_k<ZipCfg32<K>, 512, 1>(a, p, grid, st);
            else return launch_zip_k<ZipCfg32<K>, 256, 1>(a, p, grid, st);
        }
    }
    if constexpr (K >= 8) {
        if (p.lanes == 4) {
            if constexpr (K <= 24) { if (p.ctas_per_sm == 2) return launch_zip_k<ZipCfg4<K>, 256, 2>(a, p, grid, st); }
            return launch_zip_k<ZipCfg4<K>, 256, 1>(a, p, grid, st);
        }
    }
    if constexpr (K <= 24) {
        if (p.ctas_per_sm == 2) return launch_zip_k<ZipCfg8<K>, 256, 2>(a, p, grid, st);
        return launch_zip_k<ZipCfg8<K>, 512, 1>(a, p, grid, st);
    } else {
        return launch_zip_k<ZipCfg8<K>, 256, 1>(a, p, grid, st);
    }
}

static int launch_zip(ZipArgs a, const ZipPlan& p, cudaStream_t st) {
    // persistent CTAs: one per resident slot, but never more than there are (point, warp-load of quads) units
    const int cpw = 32 / p.lanes, nunits = (a.nchunks + cpw - 1) / cpw, nw = p.threads / 32;
    (void)nw;
    const long long units = (long long)a.N * nunits;     // scarce work spreads one warp-load per CTA over the SMs
    const int sms = g_ctx.sm_count > 0 ? g_ctx.sm_count : 148;
    const int grid = (int)std::min<long long>(units, (long long)sms * p.ctas_per_sm);
    // with fewer warp-loads than warps on the machine, let only as many warps per CTA claim work as it takes to cover
    // them: the chains then spread over all SMs instead of piling onto the first CTAs that arrive
    a.active_warps = (int)std::min<long long>(p.threads / 32, std::max<long long>(1, (units + grid - 1) / grid));
    switch (zip_tile(a.K)) {
#define X(k) case k: return launch_zip_shape<k>(a, p, grid, st);
        ZIP_K_LIST(X)
#undef X
    }
    return fail(IMC_ERR_UNSUPPORTED, "zip kernel is not instantiated for K = %d", a.K);
}

extern "C" int imc_seqset_zip_info(imc_seqset* set, int K, int* ids_available, int* ids_used, int64_t* tokens, int* levels) {
    if (!set) return fail(IMC_ERR_INVALID, "NULL set");
    if (ids_available) *ids_available = set->merges.size();
    if (!ids_used && !tokens && !levels) return IMC_OK;
    ZipPlan plan;
    int rc = zip_plan(K, set->nsym, set->merges.size(), &plan);
    if (rc) return rc;
    if (ids_used) *ids_used = plan.M;
    if (levels) *levels = (int)zip_levels(set->merges, plan.M).level_start.size() - 1;
    if (tokens) {
        std::vector<long long> len(set->merges.size(), 1);
        for (int id = std::max(plan.M, set->nsym); id < set->merges.size(); ++id) {
            const auto& pr = set->merges.pairs[id - set->nsym];
            len[id] = len[pr[0]] + len[pr[1]];
        }
        long long total = 0;
        for (const auto& t : set->tok_full) for (uint8_t x : t) total += len[x];
        *tokens = total;
    }
    return IMC_OK;
}

extern "C" int imc_seqset_zip_pairs(imc_seqset* set, uint8_t* pairs_out, int capacity_pairs) {
    if (!set || !pairs_out) return fail(IMC_ERR_INVALID, "NULL argument");
    if (capacity_pairs < (int)set->merges.pairs.size()) return fail(IMC_ERR_INVALID, "capacity %d < %zu pairs", capacity_pairs, set->merges.pairs.size());
    for (size_t i = 0; i < set->merges.pairs.size(); ++i) { pairs_out[2 * i] = set->merges.pairs[i][0]; pairs_out[2 * i + 1] = set->merges.pairs[i][1]; }
    return IMC_OK;
}

extern "C" int imc_seqset_zip_tokens(imc_seqset* set, int chunk, int ids, uint8_t* out, int64_t capacity, int64_t* ntokens) {
    if (!set || !ntokens) return fail(IMC_ERR_INVALID, "NULL argument");
    if (chunk < 0 || chunk >= set->n_chunks) return fail(IMC_ERR_INVALID, "chunk %d out of range", chunk);
    if (ids < set->nsym || ids > set->merges.size()) return fail(IMC_ERR_INVALID, "ids must be in [%d, %d]", set->nsym, set->merges.size());
    const int k = set->stream_of_chunk[chunk];
    if (k < 0) { *ntokens = 0; return IMC_OK; }
    std::vector<uint8_t> tok;
    zip_expand(set->merges, set->tok_full[k], ids, tok);
    *ntokens = (int64_t)tok.size();
    if (out) {
        if (capacity < (int64_t)tok.size()) return fail(IMC_ERR_INVALID, "capacity %lld < %zu tokens", (long long)capacity, tok.size());
        if (!tok.empty()) memcpy(out, tok.data(), tok.size());
    }
    return IMC_OK;
}

// ------------------------------------------------------------------------------------------ launchers

static bool pair_supported(int K) { return K == 2 || K == 4 || K == 6 || K == 8 || K == 10 || K == 12; }
static bool dmma_supported(int K) {
    switch (K) { case 10: case 12: case 16: case 20: case 24: case 28: case 32: case 36: case 40: case 48: case 64: return true; }
    return false;
}

template <int K>
static int launch_pair(const FwdArgs& a, cudaStream_t st) {
    using C = PairCfg<K>;
    static bool attr_set = false;
    if (!attr_set) {
        CUDA_TRY(cudaFuncSetAttribute(fwd_pair_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::smem_bytes()));
        attr_set = true;
    }
    const long long warps = (a.nchains + 15) / 16;
    const int wpb = C::THREADS / 32;
    const long long blocks = (warps + wpb - 1) / wpb;
    fwd_pair_kernel<K><<<(unsigned)blocks, C::THREADS, C::smem_bytes(), st>>>(a);
    return IMC_OK;
}

template <int K, int MT>
static int launch_dmma_mt(const FwdArgs& a, cudaStream_t st) {
    const int tiles = (a.nstreams + 7) / 8;
    const int warps = (tiles + MT - 1) / MT;
    const int wpb = 4;
    dim3 grid((warps + wpb - 1) / wpb, a.N);
    fwd_dmma_kernel<K, MT><<<grid, wpb * 32, 0, st>>>(a);
    return IMC_OK;
}

template <int K>
static int launch_dmma(const FwdArgs& a, cudaStream_t st, int mt_opt) {
    constexpr int NT = DmmaCfg<K>::NT;
    constexpr int MAXMT = NT <= 3 ? 4 : (NT <= 5 ? 2 : 1);   // keep 2*MT*NT*2 state doubles + K*K/32 fragment doubles in registers
    int mt = mt_opt;
    if (mt == 0) {
        // enough warps to give every SM sub-partition at least two; otherwise fewer chains per warp
        const long long tiles = (long long)((a.nstreams + 7) / 8) * a.N;
        const long long want = 2LL * 4 * (g_ctx.sm_count > 0 ? g_ctx.sm_count : 148);
        mt = MAXMT;
        while (mt > 1 && tiles / mt < want) mt >>= 1;
    }
    if (mt > MAXMT) mt = MAXMT;
    if constexpr (MAXMT >= 4) { if (mt == 4) return launch_dmma_mt<K, 4>(a, st); }
    if constexpr (MAXMT >= 2) { if (mt >= 2) return launch_dmma_mt<K, 2>(a, st); }
    return launch_dmma_mt<K, 1>(a, st);
}

static int launch_generic(const FwdArgs& a, cudaStream_t st) {
    const int K = a.K;
    int threads = K <= 64 ? 128 : (K <= 100 ? 64 : 32);
    if (a.nstreams <= 32) threads = 32; else if (a.nstreams <= 64 && threads > 64) threads = 64;
    const size_t smem = ((size_t)K * K + 4 * (size_t)K + 2 * (size_t)K * threads) * sizeof(double);
    if (smem > 227 * 1024) return fail(IMC_ERR_UNSUPPORTED, "K = %d needs %zu bytes of shared memory (max 232448)", K, smem);
    static size_t attr_max = 0;
    if (smem > attr_max) {
        CUDA_TRY(cudaFuncSetAttribute(fwd_generic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_max = smem;
    }
    dim3 grid((a.nstreams + threads - 1) / threads, a.N);
    fwd_generic_kernel<<<grid, threads, smem, st>>>(a);
    return IMC_OK;
}

static int forward_local_dev(imc_seqset* set, int N, int K, int S, const double* d_pi, const double* d_T, const double* d_E,
                             double* d_out, cudaStream_t st) {
    if (!set) return fail(IMC_ERR_INVALID, "NULL set");
    if (N < 0 || K < 1 || S < 1) return fail(IMC_ERR_INVALID, "bad sizes N=%d K=%d S=%d", N, K, S);
    if (N == 0) return IMC_OK;
    if (!set->streams.empty() && S != set->nsym)
        return fail(IMC_ERR_INVALID, "emission matrix has %d symbols but the sequences were created with nsym = %d", S, set->nsym);
    if (K > 128) return fail(IMC_ERR_UNSUPPORTED, "K = %d > 128 is not supported", K);
    int rc = ensure_device();
    if (rc) return rc;
    const int ns = (int)set->streams.size();
    if (ns == 0) {   // only empty chunks: logL = 0 for every point
        CUDA_TRY(cudaMemsetAsync(d_out, 0, sizeof(double) * (size_t)N, st));
        return IMC_OK;
    }
    if ((rc = set->d_chain.reserve(sizeof(double) * (size_t)N * ns))) return rc;

    int which = (int)g_ctx.opt_forward_kernel;
    if (which == KERNEL_AUTO) {
        ZipPlan probe;
        // zip wherever its dictionary fits; large alphabets x large K that do not fit fall back to the per-site kernels
        if (zip_supported(K) && (zip_plan(K, S, set->merges.size(), &probe) == IMC_OK || !set->packable)) which = KERNEL_ZIP;
        else which = pair_supported(K) ? KERNEL_PAIR : (dmma_supported(K) ? KERNEL_DMMA : KERNEL_GENERIC);
        // An alignment that hardly compresses (random-looking symbols) is scored faster by the per-site kernels, which
        // keep T in registers and run at the FP64 rate: per chain-step the zip kernel pays ~14 / 35 / 135 SM clocks
        // (K = 10 / 20 / 40) against ~3 / 10 / 31 per site for the pair / DMMA kernels, so below ~5 sites per token
        // -- and with enough chains to fill the machine -- the per-site kernel wins.
        if (which == KERNEL_ZIP && set->packable && (pair_supported(K) || dmma_supported(K)) && (long long)N * ns >= 4096) {
            ZipDevice* zp = nullptr;
            if (zip_device(set, probe.M, &zp) == IMC_OK && zp->total_tokens * 5 > set->total_sites)
                which = pair_supported(K) ? KERNEL_PAIR : KERNEL_DMMA;
        }
    }
    if (which == KERNEL_ZIP) {
        ZipPlan plan;
        if ((rc = zip_plan(K, S, set->merges.size(), &plan))) return rc;
        ZipDevice* z = nullptr;
        if ((rc = zip_device(set, plan.M, &z))) return rc;
        // ---- chain-scarce call (few chunks x few points)?  Three ways to run it, chosen by a cost model in SM clocks whose
        // constants come from the round-1 measurements (profiles/r01_latency_single_point.txt):
        //   (a) as it is: every chain walks its whole chunk; a lone chain advances one step per ~lat clocks
        //       (8 lanes: 36 K, 4 lanes: 49 K: one warp issuing K*K/lanes DFMA and half as many LDS.128 per step);
        //   (b) one warp per chain (32 lanes): the shortest possible step, ~80 + 14.5 K clocks;
        //   (c) segments of s tokens: K times the arithmetic on the segments after the first, but K * #segments times
        //       the chains; throughput cost c clocks of shared-memory pipe per chain-step and SM (K=10: 14, K=20: 35).
        ZipSplit* split = nullptr;
        long long seglen = g_ctx.opt_zip_segment_tokens;
        {
            const int sms = g_ctx.sm_count > 0 ? g_ctx.sm_count : 148;
            const long long slots = (long long)sms * plan.ctas_per_sm * (plan.threads / plan.lanes);
            const bool scarce = (long long)N * ns * 4 <= slots && z->max_ntok >= 256 && K <= 64;
            if (scarce && (seglen == 0 || g_ctx.opt_zip_lanes == 0)) {
                const double c = 1.4 * K * K / 16.0 + 5.0;
                const double lat = (plan.lanes == 4 ? 49.0 : 36.0) * K, lat32 = 80.0 + 14.5 * K;
                auto cost_seg = [&](long long sl) {
                    double steps = 0.0, longest = 0.0, nseg_max = 1.0;
                    for (const ZipChunk& ch : z->host_chunks) {
                        const double first = (double)std::min<long long>(sl, ch.ntok);
                        steps += first + ((double)ch.ntok - first) * K;
                        longest = std::max(longest, first);
                        nseg_max = std::max(nseg_max, std::ceil((double)ch.ntok / (double)sl));
                    }
                    const double fold = nseg_max > 1.0 ? 20000.0 + 1500.0 * 2.0 * std::sqrt(nseg_max) : 0.0;
                    return std::max(longest * lat, 1.3 * steps * N * c / sms) + fold;      // 1.3: tails and imbalance
                };
                double best = cost_seg(z->max_ntok);          // (a)
                long long best_seg = -1;
                int best_lanes = plan.lanes;
                if (seglen == 0) {
                    for (long long sl = 256; sl < z->max_ntok; sl *= 2) {        // (c)
                        const double t = cost_seg(sl);
                        if (t < 0.8 * best) { best = t; best_seg = sl; }
                    }
                }
                if (g_ctx.opt_zip_lanes == 0 && zip_tile(K) >= 10 && seglen <= 0) {   // (b)
                    const long long warps = (long long)sms * (zip_tile(K) <= 24 ? 16 : 8);
                    const double rounds = std::ceil((double)((long long)N * ns) / (double)warps);
                    const double t = rounds * z->max_ntok * lat32;
                    if (t < best) { best = t; best_seg = -1; best_lanes = 32; }
                }
                if (seglen == 0) seglen = best_seg;
                if (best_lanes != plan.lanes) {
                    if ((rc = zip_plan(K, S, set->merges.size(), &plan, best_lanes))) return rc;
                    if ((rc = zip_device(set, plan.M, &z))) return rc;
                }
            }
        }
        ZipArgs za;
        za.tokens = (const uint8_t*)z->tokens.p;
        za.chunks = (const ZipChunk*)z->chunks.p;
        za.nchunks = ns;
        if ((rc = set->d_pnext.reserve(sizeof(int) * (size_t)N))) return rc;
        CUDA_TRY(cudaMemsetAsync(set->d_pnext.p, 0, sizeof(int) * (size_t)N, st));
        za.point_next = (int*)set->d_pnext.p;
        za.pairs = (const uint8_t*)z->pairs.p;
        za.level_start = (const int*)z->levels.p;
        za.nlevels = z->nlevels;
        za.M = plan.M;
        za.N = N; za.K = K; za.S = S;
        za.pi = d_pi; za.T = d_T; za.E = d_E;
        za.chain_out = (double*)set->d_chain.p;
        za.out_stride = ns;
        za.vec_out = nullptr;
        za.vec_stride = 0;
        if (seglen > 0 && K <= 64) {
            seglen = (seglen + 15) / 16 * 16;
            if (seglen < z->max_ntok) {
                if ((rc = zip_split(z, K, (int)seglen, &split))) return rc;
                const size_t vec_bytes = sizeof(double) * (size_t)N * ((size_t)split->nchains + split->nvec2) * (K + 1);
                if (vec_bytes > (size_t)1 << 30) split = nullptr;      // not worth a gigabyte of scratch
                else if ((rc = set->d_vec.reserve(vec_bytes))) return rc;
            }
        }
        if (split) {
            za.chunks = (const ZipChunk*)split->chunks.p;
            za.nchunks = split->nchains;
            za.vec_out = (double*)set->d_vec.p;
            za.vec_stride = K + 1;
        }
        g_last_kernel = split ? "zip-segmented" : (plan.lanes == 32 ? "zip-warp" : "zip");
        if ((rc = launch_zip(za, plan, st))) return rc;
        CUDA_TRY(cudaGetLastError());
        if (split) {
            double* vec2 = za.vec_out + (size_t)N * split->nchains * za.vec_stride;
            if (split->n_level1 > 0) {
                zip_fold_kernel<<<dim3(split->n_level1, N), 64, 0, st>>>(za.vec_out, split->nchains, vec2, split->nvec2, za.vec_stride,
                                                                        (const ZipFoldItem*)split->items1.p, K, za.chain_out, za.out_stride);
                CUDA_TRY(cudaGetLastError());
                g_launches += 1;
            }
            zip_fold_kernel<<<dim3(split->n_final, N), 64, 0, st>>>(za.vec_out, split->nchains, vec2, split->nvec2, za.vec_stride,
                                                                   (const ZipFoldItem*)split->items2.p, K, za.chain_out, za.out_stride);
            CUDA_TRY(cudaGetLastError());
            g_launches += 1;
        }
        reduce_chains_kernel<<<N, 256, 0, st>>>(za.chain_out, ns, d_out);
        CUDA_TRY(cudaGetLastError());
        g_launches += 2;
        return IMC_OK;
    }
    if (!set->packable)
        return fail(IMC_ERR_UNSUPPORTED, "alphabets larger than 3 symbols run on the zip kernel only (nsym = %d, K = %d)", set->nsym, K);
    if ((rc = seqset_upload(set))) return rc;
    FwdArgs a;
    a.words = (const uint32_t*)set->d_words.p;
    a.streams = (const StreamInfo*)set->d_streams.p;
    a.nstreams = ns;
    a.N = N; a.K = K; a.S = S;
    a.pi = d_pi; a.T = d_T; a.E = d_E;
    a.chain_out = (double*)set->d_chain.p;
    a.nchains = (long long)N * ns;
    a.fold_sym = g_ctx.opt_fold_emission ? set->fold_sym : -1;

    if (which == KERNEL_PAIR) {
        if (!pair_supported(K)) return fail(IMC_ERR_UNSUPPORTED, "lane-pair kernel covers even K <= 12, not K = %d", K);
        g_last_kernel = "pair";
        switch (K) {
            case 2: rc = launch_pair<2>(a, st); break;
            case 4: rc = launch_pair<4>(a, st); break;
            case 6: rc = launch_pair<6>(a, st); break;
            case 8: rc = launch_pair<8>(a, st); break;
            case 10: rc = launch_pair<10>(a, st); break;
            case 12: rc = launch_pair<12>(a, st); break;
        }
    } else if (which == KERNEL_DMMA) {
        if (!dmma_supported(K)) return fail(IMC_ERR_UNSUPPORTED, "DMMA kernel is not instantiated for K = %d", K);
        g_last_kernel = "dmma";
        const int mt = (int)g_ctx.opt_dmma_mtiles;
        switch (K) {
            case 10: rc = launch_dmma<10>(a, st, mt); break;
            case 12: rc = launch_dmma<12>(a, st, mt); break;
            case 16: rc = launch_dmma<16>(a, st, mt); break;
            case 20: rc = launch_dmma<20>(a, st, mt); break;
            case 24: rc = launch_dmma<24>(a, st, mt); break;
            case 28: rc = launch_dmma<28>(a, st, mt); break;
            case 32: rc = launch_dmma<32>(a, st, mt); break;
            case 36: rc = launch_dmma<36>(a, st, mt); break;
            case 40: rc = launch_dmma<40>(a, st, mt); break;
            case 48: rc = launch_dmma<48>(a, st, mt); break;
            case 64: rc = launch_dmma<64>(a, st, mt); break;
        }
    } else if (which == KERNEL_GENERIC) {
        g_last_kernel = "generic";
        rc = launch_generic(a, st);
    } else {
        return fail(IMC_ERR_INVALID, "unknown forward_kernel option %d", which);
    }
    if (rc) return rc;
    CUDA_TRY(cudaGetLastError());
    reduce_chains_kernel<<<N, 256, 0, st>>>(a.chain_out, ns, d_out);
    CUDA_TRY(cudaGetLastError());
    g_launches += 2;
    return IMC_OK;
}

// ------------------------------------------------------------------------------------------ multi-GPU: the one collective
// Chunks are sharded over the ranks (one process per GPU); every rank scores the same parameter batch on its shard and
// the partial logL[N] are summed by ONE all-reduce per batch.  NCCL is resolved at run time so that the
// library has no link-time dependency on it: a single-GPU user never needs libnccl, and a process that already loaded
// torch's NCCL reuses that copy.
struct NcclApi {
    void* handle = nullptr;
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclAllReduce) AllReduce = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
};
static NcclApi g_nccl;
static ncclComm_t g_comm = nullptr;

static int nccl_load() {
    if (g_nccl.handle) return IMC_OK;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return fail(IMC_ERR_UNSUPPORTED, "libnccl.so.2 cannot be loaded: %s", dlerror());
    g_nccl.GetUniqueId = (decltype(g_nccl.GetUniqueId))dlsym(h, "ncclGetUniqueId");
    g_nccl.CommInitRank = (decltype(g_nccl.CommInitRank))dlsym(h, "ncclCommInitRank");
    g_nccl.AllReduce = (decltype(g_nccl.AllReduce))dlsym(h, "ncclAllReduce");
    g_nccl.CommDestroy = (decltype(g_nccl.CommDestroy))dlsym(h, "ncclCommDestroy");
    g_nccl.GetErrorString = (decltype(g_nccl.GetErrorString))dlsym(h, "ncclGetErrorString");
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllReduce || !g_nccl.CommDestroy || !g_nccl.GetErrorString)
        return fail(IMC_ERR_UNSUPPORTED, "libnccl.so.2 lacks an expected symbol");
    g_nccl.handle = h;
    return IMC_OK;
}
#define NCCL_TRY(x)                                                                                       \
    do {                                                                                                  \
        ncclResult_t r_ = (x);                                                                            \
        if (r_ != ncclSuccess) return fail(IMC_ERR_CUDA, "%s failed: %s", #x, g_nccl.GetErrorString(r_)); \
    } while (0)

extern "C" int imc_comm_unique_id(void* id_out, int capacity) {
    if (!id_out || capacity < (int)sizeof(ncclUniqueId)) return fail(IMC_ERR_INVALID, "id buffer must hold %zu bytes", sizeof(ncclUniqueId));
    int rc = nccl_load();
    if (rc) return rc;
    ncclUniqueId id;
    NCCL_TRY(g_nccl.GetUniqueId(&id));
    memcpy(id_out, &id, sizeof id);
    return IMC_OK;
}

extern "C" int imc_comm_init(int nranks, int rank, const void* nccl_id) {
    if (nranks < 1 || rank < 0 || rank >= nranks || (nranks > 1 && !nccl_id)) return fail(IMC_ERR_INVALID, "bad communicator arguments");
    if (g_comm) return fail(IMC_ERR_INVALID, "communicator already initialised; call imc_comm_destroy first");
    int rc = ensure_device();
    if (rc) return rc;
    if (nranks == 1) return IMC_OK;       // nothing to sum over
    if ((rc = nccl_load())) return rc;
    ncclUniqueId id;
    memcpy(&id, nccl_id, sizeof id);
    NCCL_TRY(g_nccl.CommInitRank(&g_comm, nranks, id, rank));
    return IMC_OK;
}

extern "C" int imc_comm_destroy(void) {
    if (g_comm && g_ctx.pid == getpid()) g_nccl.CommDestroy(g_comm);
    g_comm = nullptr;
    return IMC_OK;
}

// this rank's partial log-likelihoods, then the sum over ranks when a communicator exists
static int forward_dev(imc_seqset* set, int N, int K, int S, const double* d_pi, const double* d_T, const double* d_E,
                       double* d_out, cudaStream_t st) {
    int rc = forward_local_dev(set, N, K, S, d_pi, d_T, d_E, d_out, st);
    if (rc || !g_comm || N <= 0) return rc;
    NCCL_TRY(g_nccl.AllReduce(d_out, d_out, (size_t)N, ncclDouble, ncclSum, g_comm, st));
    return IMC_OK;
}

extern "C" int imc_forward_batch_dev(imc_seqset* set, int N, int K, int S, const double* d_pi, const double* d_T,
                                     const double* d_E, double* d_out, void* stream) {
    int rc = ensure_device();
    if (rc) return rc;
    if (N > 0 && (!d_pi || !d_T || !d_E || !d_out)) return fail(IMC_ERR_INVALID, "NULL device pointer");
    return forward_dev(set, N, K, S, d_pi, d_T, d_E, d_out, (cudaStream_t)stream);
}

extern "C" int imc_forward_batch(imc_seqset* set, int N, int K, int S, const double* pi, const double* T,
                                 const double* E, double* out) {
    if (!set) return fail(IMC_ERR_INVALID, "NULL set");
    if (N < 0 || K < 1 || S < 1) return fail(IMC_ERR_INVALID, "bad sizes N=%d K=%d S=%d", N, K, S);
    if (N == 0) return IMC_OK;
    if (!pi || !T || !E || !out) return fail(IMC_ERR_INVALID, "NULL host pointer");
    int rc = ensure_device();
    if (rc) return rc;
    const size_t npi = (size_t)N * K, nT = (size_t)N * K * K, nE = (size_t)N * K * S;
    if ((rc = set->d_pi.reserve(npi * sizeof(double)))) return rc;
    if ((rc = set->d_T.reserve(nT * sizeof(double)))) return rc;
    if ((rc = set->d_E.reserve(nE * sizeof(double)))) return rc;
    if ((rc = set->d_out.reserve((size_t)N * sizeof(double)))) return rc;
    cudaStream_t st = g_ctx.stream;
    CUDA_TRY(cudaMemcpyAsync(set->d_pi.p, pi, npi * sizeof(double), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(set->d_T.p, T, nT * sizeof(double), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(set->d_E.p, E, nE * sizeof(double), cudaMemcpyHostToDevice, st));
    rc = forward_dev(set, N, K, S, (const double*)set->d_pi.p, (const double*)set->d_T.p, (const double*)set->d_E.p,
                     (double*)set->d_out.p, st);
    if (rc) return rc;
    CUDA_TRY(cudaMemcpyAsync(out, set->d_out.p, (size_t)N * sizeof(double), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return IMC_OK;
}

extern "C" int imc_forward(imc_seqset* set, int K, int S, const double* pi, const double* T, const double* E,
                           double* logL_out) {
    return imc_forward_batch(set, 1, K, S, pi, T, E, logL_out);
}

// ------------------------------------------------------------------------------------------ FP64 peak probe
extern "C" int imc_measure_fp64_peak(double* dfma_tflops, double* dmma_tflops) {
    int rc = ensure_device();
    if (rc) return rc;
    double* d_out = nullptr;
    CUDA_TRY(cudaMalloc(&d_out, 64));
    cudaEvent_t e0, e1;
    CUDA_TRY(cudaEventCreate(&e0));
    CUDA_TRY(cudaEventCreate(&e1));
    const int sms = g_ctx.sm_count, threads = 512, iters = 20000;
    cudaStream_t st = g_ctx.stream;
    double best_fma = 0.0, best_mma = 0.0;
    for (int rep = 0; rep < 4; ++rep) {
        float ms = 0.f;
        CUDA_TRY(cudaEventRecord(e0, st));
        peak_dfma_kernel<<<sms, threads, 0, st>>>(d_out, iters, 1.0000001, 1e-9);
        CUDA_TRY(cudaEventRecord(e1, st));
        CUDA_TRY(cudaEventSynchronize(e1));
        CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
        if (rep) best_fma = std::max(best_fma, 2.0 * 64.0 * iters * (double)threads * sms / (ms * 1e9));
        CUDA_TRY(cudaEventRecord(e0, st));
        peak_dmma_kernel<<<sms, threads, 0, st>>>(d_out, iters / 4);
        CUDA_TRY(cudaEventRecord(e1, st));
        CUDA_TRY(cudaEventSynchronize(e1));
        CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
        if (rep) best_mma = std::max(best_mma, 512.0 * 32.0 * (iters / 4) * (double)(threads / 32) * sms / (ms * 1e9));
    }
    CUDA_TRY(cudaGetLastError());
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d_out);
    g_launches += 8;
    if (dfma_tflops) *dfma_tflops = best_fma;
    if (dmma_tflops) *dmma_tflops = best_mma;
    return IMC_OK;
}

// ------------------------------------------------------------------------------------------ options
extern "C" int imc_set_option(const char* key, int64_t value) {
    if (!key) return fail(IMC_ERR_INVALID, "NULL key");
    if (!strcmp(key, "forward_kernel")) {
        if (value < 0 || value > 4) return fail(IMC_ERR_INVALID, "forward_kernel must be 0..4");
        g_ctx.opt_forward_kernel = value;
        return IMC_OK;
    }
    if (!strcmp(key, "dmma_mtiles")) {
        if (value != 0 && value != 1 && value != 2 && value != 4) return fail(IMC_ERR_INVALID, "dmma_mtiles must be 0, 1, 2 or 4");
        g_ctx.opt_dmma_mtiles = value;
        return IMC_OK;
    }
    if (!strcmp(key, "fold_emission")) { g_ctx.opt_fold_emission = value ? 1 : 0; return IMC_OK; }
    if (!strcmp(key, "zip_ctas_per_sm")) { if (value < 0 || value > 2) return fail(IMC_ERR_INVALID, "zip_ctas_per_sm must be 0, 1 or 2"); g_ctx.opt_zip_ctas_per_sm = value; return IMC_OK; }
    if (!strcmp(key, "zip_segment_tokens")) { if (value < -1) return fail(IMC_ERR_INVALID, "zip_segment_tokens must be >= -1"); g_ctx.opt_zip_segment_tokens = value; return IMC_OK; }
    if (!strcmp(key, "zip_lanes")) { if (value != 0 && value != 4 && value != 8 && value != 32) return fail(IMC_ERR_INVALID, "zip_lanes must be 0, 4, 8 or 32"); g_ctx.opt_zip_lanes = value; return IMC_OK; }
    if (!strcmp(key, "zip_max_entries")) { if (value < 0 || value > 256) return fail(IMC_ERR_INVALID, "zip_max_entries must be in [0, 256]"); g_ctx.opt_zip_max_entries = value; return IMC_OK; }
    return fail(IMC_ERR_INVALID, "unknown option '%s'", key);
}
extern "C" int imc_get_option(const char* key, int64_t* value_out) {
    if (!key || !value_out) return fail(IMC_ERR_INVALID, "NULL argument");
    if (!strcmp(key, "forward_kernel")) { *value_out = g_ctx.opt_forward_kernel; return IMC_OK; }
    if (!strcmp(key, "dmma_mtiles")) { *value_out = g_ctx.opt_dmma_mtiles; return IMC_OK; }
    if (!strcmp(key, "fold_emission")) { *value_out = g_ctx.opt_fold_emission; return IMC_OK; }
    if (!strcmp(key, "zip_ctas_per_sm")) { *value_out = g_ctx.opt_zip_ctas_per_sm; return IMC_OK; }
    if (!strcmp(key, "zip_segment_tokens")) { *value_out = g_ctx.opt_zip_segment_tokens; return IMC_OK; }
    if (!strcmp(key, "zip_lanes")) { *value_out = g_ctx.opt_zip_lanes; return IMC_OK; }
    if (!strcmp(key, "zip_max_entries")) { *value_out = g_ctx.opt_zip_max_entries; return IMC_OK; }
    return fail(IMC_ERR_INVALID, "unknown option '%s'", key);
}
extern "C" int64_t imc_kernel_launches(void) { return g_launches.load(); }
extern "C" const char* imc_last_forward_kernel(void) { return g_last_kernel; }

// ------------------------------------------------------------------------------------------ model build
#include "model_host.inl"
