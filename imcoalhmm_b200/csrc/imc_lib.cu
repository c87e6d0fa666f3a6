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
// NCCL: types and prototypes only -- the library itself is resolved with dlopen when imc_comm_init is called.  Without the
// development header the six entry points used are declared here (their ABI has been stable since NCCL 2.0), so that the
// single-GPU library builds on machines that have no NCCL at all.
#if defined(__has_include)
#if __has_include(<nccl.h>)
#include <nccl.h>
#define IMC_HAVE_NCCL_H 1
#endif
#endif
#ifndef IMC_HAVE_NCCL_H
extern "C" {
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef enum { ncclSuccess = 0 } ncclResult_t;
typedef enum { ncclSum = 0, ncclProd = 1, ncclMax = 2, ncclMin = 3 } ncclRedOp_t;
typedef enum { ncclInt8 = 0, ncclChar = 0, ncclUint8 = 1, ncclInt32 = 2, ncclInt = 2, ncclUint32 = 3, ncclInt64 = 4, ncclUint64 = 5,
               ncclFloat16 = 6, ncclFloat32 = 7, ncclFloat64 = 8, ncclDouble = 8 } ncclDataType_t;
ncclResult_t ncclGetUniqueId(ncclUniqueId* uniqueId);
ncclResult_t ncclCommInitRank(ncclComm_t* comm, int nranks, ncclUniqueId commId, int rank);
ncclResult_t ncclAllReduce(const void* sendbuff, void* recvbuff, size_t count, ncclDataType_t datatype, ncclRedOp_t op, ncclComm_t comm, cudaStream_t stream);
ncclResult_t ncclAllGather(const void* sendbuff, void* recvbuff, size_t sendcount, ncclDataType_t datatype, ncclComm_t comm, cudaStream_t stream);
ncclResult_t ncclCommDestroy(ncclComm_t comm);
const char* ncclGetErrorString(ncclResult_t result);
}
#endif
#include <sched.h>
#include <string>
#include <unistd.h>
#include <vector>

#include "tokenizer.inl"

// NVTX ranges around the stages of a call (model build / spectral preparation / forward passes / reduction), so that
// Nsight Systems timelines read as the design does.  Header-only (nvtx3); a no-op unless a tool is attached.
#if defined(__has_include)
#if __has_include(<nvtx3/nvToolsExt.h>)
#include <nvtx3/nvToolsExt.h>
#define IMC_HAVE_NVTX 1
#endif
#endif
struct NvtxRange {
#ifdef IMC_HAVE_NVTX
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
#else
    explicit NvtxRange(const char*) {}
#endif
};

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
    long long opt_zip_pipeline = 0;    // pieces per chunk in pipelined mode: 0 = auto, 1 = off, >= 2 forced
    long long opt_zip_spectral = 0;    // spectral form of the zip kernel (run tokens): 0 = auto, 1 = always, 2 = never
    long long opt_zip_mma = 0;         // MMA form of the spectral kernel: 0 = auto, 1 = always (tiles >= 8), 2 = never
    long long opt_zip_align = 0;       // aligned form of the MMA shape (one entry per warp-step, host-built schedules): 0 = auto, 1 = on, 2 = off
    long long opt_zip_mma_shape = 0;   // launch shape of the MMA form: 0 = auto, 1..4 see zip_plan_k
    long long opt_zip_run2 = 0;        // two-run form of the MMA shape: 0 = auto (fewer expected DMMA passes), 1 = always, 2 = never
    long long opt_zip_spectral_force_bad = 0;   // test switch: zip_spectral_kernel declares every point unfit (plain-form pass serves them)
    long long opt_comm_fused = 1;      // map peer mailboxes at imc_comm_init and all-reduce inside the reduction kernel
    long long opt_comm_enabled = 1;    // 0: forward / loglik calls return this rank's partial sums although a communicator exists
    long long opt_comm_timeout_ms = 30000;   // how long the fused all-reduce waits for a peer before it gives up (NaN + error)
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

// One call at a time per handle: the scratch buffers of a sequence set / model (chain sums, work counters, carry vectors,
// pi / T / E staging) are per handle, so a second call -- from another host thread or on another stream -- must not start
// before the first has finished with them.  The guard serialises the host side with a (recursive) mutex and the device
// side with an event: every call waits for the event its predecessor recorded, whatever streams the two were issued on.
struct HandleSerial {
    std::recursive_mutex mu;
    cudaEvent_t done = nullptr;
};
struct CallGuard {
    HandleSerial& h;
    cudaStream_t st;
    CallGuard(HandleSerial& hs, cudaStream_t s) : h(hs), st(s) {
        h.mu.lock();
        if (h.done) cudaStreamWaitEvent(st, h.done, 0);
    }
    ~CallGuard() {
        if (!h.done) cudaEventCreateWithFlags(&h.done, cudaEventDisableTiming);
        if (h.done) cudaEventRecord(h.done, st);
        h.mu.unlock();
    }
};

// device copy of the token streams derived for one dictionary size M (see zip_device)
struct ZipSplit {                 // segmented variant of a ZipDevice's chunk list (same token buffer)
    int K = 0, seglen = 0, nchains = 0, parts_local = 0;
    int n_level1 = 0, n_final = 0, nvec2 = 0;     // fold items of the two levels, vectors written by level 1
    DeviceBuf chunks, items1, items2;
};

struct ZipDevice {
    int M = 0, nlevels = 0;
    bool spec = false;                        // run tokens (32-bit words) for the spectral form of the kernel
    int hot_id = 0;                           // spec: the most frequent entry and its share of the tokens
    double hot_share = 0.0;
    bool run2 = false;                        // two-run form (nsym + 2 base entries, second power table)
    double est_passes = 1.0;                  // expected MMA passes per warp-step: 1 + sum over cold ids of P(some chain of 8 is on it)
    bool sched = false;                       // aligned form: per-quad schedules, streams padded with no-op words (zip_align_quad)
    long long pass_cost = 0;                  // MMA passes of one parameter point over all warp-loads (estimate in lock step, exact when aligned)
    long long total_tokens = 0;
    int max_ntok = 0;
    std::vector<ZipChunk> host_chunks;        // sorted by ntok, descending
    DeviceBuf tokens, chunks, pairs, levels;
    std::vector<ZipSplit*> splits;
};

struct imc_seqset {
    HandleSerial serial;
    // parts mode (imc_seqset_create_parts): the chunks of this set are consecutive PARTS part_first .. of ONE long chunk that
    // is cut into parts_total parts over the ranks of the communicator; 0 = ordinary set of independent chunks
    int parts_total = 0, part_first = 0;
    DeviceBuf d_parts, d_gather;           // this rank's block of part vectors; the all-gathered blocks
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
    // run tokens for the spectral form (tokenizer.inl): the run symbol is fold_sym; own pair dictionary over the entries
    ZipMerges run_merges;
    std::vector<std::vector<uint32_t>> run_tok_full;   // per stream, over all run_merges.size() ids
    std::vector<int> first_run;                        // per stream
    long long zip_tokens_full = 0, run_tokens_full = 0;   // stream lengths over the full dictionaries (which form compresses better)
    // two-run form (tokenizer.inl; built on first use): second run symbol, its own pair dictionary over nsym + 2 base ids
    int run2_state = 0;                                // 0 = not looked at yet, 1 = available, -1 = nothing to gain / not possible
    int run_sym2 = -1;
    ZipMerges run2_merges;
    std::vector<std::vector<uint32_t>> run2_tok_full;
    std::vector<int> run2_first_run;
    std::vector<long long> run2_sites;
    // device side (lazy)
    bool uploaded = false;
    DeviceBuf d_words, d_streams, d_chain, d_pi, d_T, d_E, d_out;
    DeviceBuf d_pnext[2], d_vec[2], d_prog[2];    // scratch of the two passes of a forward call (spectral, plain)
    DeviceBuf d_spec, d_lists;                    // per-point spectral data; ok / bad point lists + their counts
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
    bool in_num = false, neg = false, digits = false;
    int rc = IMC_OK;
    size_t got;
    long long pos = 0;
    auto flush = [&]() -> int {
        if (!in_num) return IMC_OK;
        if (!digits) return fail(IMC_ERR_IO, "token %zu of '%s' is a sign without digits", s->sym.size(), path);
        const long long v = neg ? -cur : cur;
        if (v < 0 || v >= nsym) return fail(IMC_ERR_INVALID, "symbol %lld (token %zu) in '%s' is outside [0, %d)", v, s->sym.size(), path, nsym);
        s->sym.push_back((uint8_t)v);
        in_num = false; neg = false; digits = false; cur = 0;
        return IMC_OK;
    };
    while (rc == IMC_OK && (got = fread(buf.data(), 1, buf.size(), f)) > 0) {
        for (size_t i = 0; i < got && rc == IMC_OK; ++i, ++pos) {
            const char ch = buf[i];
            if (ch >= '0' && ch <= '9') {
                cur = cur * 10 + (ch - '0');
                if (cur > 1000000) cur = 1000000;  // saturate; rejected by the range check
                in_num = true; digits = true;
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

#include "ingest_host.inl"

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
static int seqset_create_impl(const imc_seq* const* seqs, int C, int part_first, int parts_total, imc_seqset** out) {
    if (!out || C < 0 || (C > 0 && !seqs)) return fail(IMC_ERR_INVALID, "bad arguments");
    imc_seqset* set = new (std::nothrow) imc_seqset;
    if (!set) return fail(IMC_ERR_NOMEM, "out of memory");
    set->n_chunks = C;
    set->parts_total = parts_total;
    set->part_first = part_first;
    std::vector<int> order;
    for (int c = 0; c < C; ++c) {
        if (!seqs[c]) { delete set; return fail(IMC_ERR_INVALID, "seqs[%d] is NULL", c); }
        if (c == 0) set->nsym = seqs[c]->nsym;
        if (seqs[c]->nsym != set->nsym) { delete set; return fail(IMC_ERR_INVALID, "chunks disagree on nsym (%d vs %d)", seqs[c]->nsym, set->nsym); }
        if (seqs[c]->sym.size() > 0x7fffffffULL) { delete set; return fail(IMC_ERR_UNSUPPORTED, "chunk %d has more than 2^31-1 sites; split it", c); }
        set->total_sites += (long long)seqs[c]->sym.size();
        if (parts_total > 0 && seqs[c]->sym.empty()) { delete set; return fail(IMC_ERR_INVALID, "part %d of a cut chunk is empty", part_first + c); }
        if (!seqs[c]->sym.empty()) order.push_back(c);   // an empty chunk contributes logL = 0
    }
    set->packable = set->nsym <= 3 && parts_total == 0;      // the per-site kernels know nothing of parts
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
                    if (sy.size() < 2) continue;      // (a later part's position 0 is skipped in the sample like any chunk's: harmless)
                    const long long share = budget_total / std::min(32, std::max(1, C));
                    const size_t take = (size_t)std::min<long long>({(long long)sy.size() - 1, budget, share});
                    sample.emplace_back(sy.begin() + 1, sy.begin() + 1 + take);
                    budget -= (long long)take;
                }
            std::vector<std::vector<uint32_t>> rsample(sample.size());
            for (size_t i = 0; i < sample.size(); ++i) {
                int fr;
                run_tokenize(sample[i].data(), sample[i].size(), set->fold_sym, &fr, rsample[i]);
            }
            set->merges = zip_learn(sample, set->nsym, 256, 16);
            set->run_merges = run_learn(rsample, set->nsym, 256, 16);
        }
        set->tok_full.resize(ns);
        set->run_tok_full.resize(ns);
        set->first_run.assign(ns, 0);
        set->first_sym.resize(ns);
        set->stream_of_chunk.assign(C, -1);
        for (int k = 0; k < ns; ++k) set->stream_of_chunk[order[k]] = k;
        if (!parallel_for(ns, [&](int k) {
                const auto& sy = seqs[order[k]]->sym;
                set->first_sym[k] = sy[0];
                // a later part of a cut chunk has no start of its own: its position 0 is an ordinary site of the stream
                const size_t skip = (parts_total > 0 && part_first + order[k] > 0) ? 0 : 1;
                zip_encode(set->merges, sy.data() + skip, sy.size() - skip, set->tok_full[k]);
                run_encode(set->run_merges, sy.data() + skip, sy.size() - skip, set->fold_sym, &set->first_run[k], set->run_tok_full[k]);
            })) throw std::bad_alloc();
        for (int k = 0; k < ns; ++k) {
            set->zip_tokens_full += (long long)set->tok_full[k].size();
            set->run_tokens_full += (long long)set->run_tok_full[k].size();
        }
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

extern "C" int imc_seqset_create(const imc_seq* const* seqs, int C, imc_seqset** out) {
    return seqset_create_impl(seqs, C, 0, 0, out);
}

extern "C" int imc_seqset_create_parts(const imc_seq* const* parts, int n_local, int part_first, int parts_total, imc_seqset** out) {
    if (n_local < 1 || parts_total < 1 || part_first < 0 || part_first + n_local > parts_total || parts_total % n_local != 0 || part_first % n_local != 0)
        return fail(IMC_ERR_INVALID, "parts: every rank holds the same number n_local of consecutive parts (first = rank * n_local) of parts_total");
    return seqset_create_impl(parts, n_local, part_first, parts_total, out);
}

extern "C" int imc_seqset_destroy(imc_seqset* set) {
    if (!set) return IMC_OK;
    const bool mine = g_ctx.pid == getpid();
    if (mine) {
        set->d_words.release(); set->d_streams.release(); set->d_chain.release();
        set->d_pi.release(); set->d_T.release(); set->d_E.release(); set->d_out.release();
        for (int i = 0; i < 2; ++i) { set->d_pnext[i].release(); set->d_vec[i].release(); set->d_prog[i].release(); }
        set->d_spec.release(); set->d_lists.release(); set->d_parts.release(); set->d_gather.release();
        if (set->serial.done) cudaEventDestroy(set->serial.done);
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


#include "zip_host.inl"

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

static int launch_chain_reduce(const double* chain, int ns, int N, double* d_out, cudaStream_t st);   // comm_host.inl

// One pass of the compressed forward over the points of a list (plist == NULL: all N points): plan, token streams, the
// choice among whole chunks / pipelined pieces / one warp per chain / segments, the launch and (segmented) the folds.
// spec: spectral form over run tokens (points of the ok list), else the plain form (pass 1 scratch).  Results land in
// set->d_chain[n][chunk] for the points served.
static const int MAX_POINTS_PARTS = 32768;
constexpr double ZIP_STEP_LATENCY = 120.0;         // clocks of a lone warp's step beside twice its FP64 pipe time (see the aligned-form decision)
static thread_local bool g_want_sched = false;     // aligned form (implies the two-run form) wanted, if the call turns out not to be chain-scarce
static thread_local bool g_want_run2 = false;      // two-run form wanted for the spectral pass of the current call

static int zip_pass(imc_seqset* set, int N, int K, int S, const double* d_pi, const double* d_T, const double* d_E, bool spec,
                    const int* plist, const int* pcount, const double* d_spec, int spec_stride, cudaStream_t st) {
    NvtxRange nvtx_pass(spec ? "imc: zip pass (spectral form)" : "imc: zip pass (plain form)");
    int rc;
    const int ns = (int)set->streams.size(), pass = spec ? 0 : 1;
    const int avail = spec ? set->run_merges.size() : set->merges.size();
    g_plan_chunks = ns;
    ZipPlan plan;
    if ((rc = zip_plan(K, S, avail, &plan, 0, spec))) return rc;
    ZipDevice* z = nullptr;
    if ((rc = zip_device(set, plan.M, &z, spec))) return rc;
    // MMA form: eight chains per warp as the rows of FP64 tensor-core tiles, the most frequent entry's matrix in registers.
    // Worth it where one entry dominates the streams (the lone mismatch followed by its run: 75-90 % of the tokens of a
    // pairwise alignment); chains on other entries cost a pass of their own.
    bool mma = false;
    if (spec && zip_mma_tile(zip_tile(K)) && g_ctx.opt_zip_mma != 2 && g_ctx.opt_zip_lanes == 0) {
        ZipPlan mp;
        ZipDevice* mz = nullptr;
        const bool r2 = g_want_run2 && set->run2_state == 1;          // two-run form: decided by forward_local_dev before the preparation kernel
        g_mma_shape_hint = g_want_sched && r2 && set->parts_total == 0 && g_ctx.opt_zip_segment_tokens <= 0 ? 2 : 0;
        const int planned = zip_plan(K, S, r2 ? set->run2_merges.size() : avail, &mp, 0, true, true, r2);
        g_mma_shape_hint = 0;
        if (planned == IMC_OK && zip_device(set, mp.M, &mz, true, r2) == IMC_OK && (g_ctx.opt_zip_mma == 1 || mz->hot_share >= 0.5)) {
            plan = mp; z = mz; mma = true;
        }
    }
    const bool try_sched = mma && z->run2 && g_want_sched && set->parts_total == 0 && g_ctx.opt_zip_segment_tokens <= 0;
    ZipDevice* zs = nullptr;           // the aligned streams of the same dictionary
    if (try_sched && (rc = zip_device(set, z->M, &zs, true, true, true))) return rc;
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
        const bool scarce = (long long)N * ns * 4 <= slots && z->max_ntok >= 64 && K <= 64 && set->parts_total == 0;
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
            if (zs) {
                // (a') whole chunks in the aligned form: ~330 clocks per step of a lone warp (8 chunks x 256 points: 1.00 ms where
                // the warp-per-chain shape took 1.63; 16 chunks: 1.27 vs 2.47 ms), one set of DMMAs per step when the SMs fill up
                const int tile = zip_tile(K);
                const double P = 16.0 * ((tile + 3) / 4) * ((tile + 7) / 8) + 8.0 * ((tile + 7) / 8);
                best = std::min(best, std::max((ZIP_STEP_LATENCY + 2.0 * P) * zs->max_ntok, 1.3 * (double)zs->pass_cost * N * P / 0.8 / (4.0 * sms)));
            }
            if (seglen == 0) {
                for (long long sl = 64; sl < z->max_ntok; sl *= 2) {        // (c)
                    const double t = cost_seg(sl);
                    if (t < 0.8 * best) { best = t; best_seg = sl; }
                }
            }
            // (the warp-per-chain shape is for a handful of chains; a batch large enough for the aligned form -- >= 1024 chains -- is
            // not: 8 / 16 chunks x 256 points took 1.63 / 2.47 ms there, 1.00 / 1.27 ms aligned)
            if (g_ctx.opt_zip_lanes == 0 && zip_tile(K) >= 10 && seglen <= 0 && g_ctx.opt_zip_mma != 1 && !zs) {   // (b)
                const long long warps = (long long)sms * (zip_tile(K) <= 24 ? 16 : 8);
                const double rounds = std::ceil((double)((long long)N * ns) / (double)warps);
                const double t = rounds * z->max_ntok * lat32;
                if (t < best) { best = t; best_seg = -1; best_lanes = 32; }
            }
            if (seglen == 0) seglen = best_seg;
            if (best_seg > 0 || best_lanes != plan.lanes) zs = nullptr;          // segments or the warp shape won: lock-step forms
            if (best_lanes != plan.lanes) {
                if ((rc = zip_plan(K, S, avail, &plan, best_lanes, spec))) return rc;
                if ((rc = zip_device(set, plan.M, &z, spec))) return rc;
                mma = false;
            }
        }
    }
    if (zs && mma && seglen <= 0) z = zs;        // not segmented, not the warp shape: the aligned streams
    ZipArgs za;
    za.tokens = (const uint8_t*)z->tokens.p;
    za.chunks = (const ZipChunk*)z->chunks.p;
    za.nchunks = ns;
    if ((rc = set->d_pnext[pass].reserve(sizeof(int) * (size_t)N))) return rc;
    CUDA_TRY(cudaMemsetAsync(set->d_pnext[pass].p, 0, sizeof(int) * (size_t)N, st));
    za.point_next = (int*)set->d_pnext[pass].p;
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
    za.plist = plist; za.pcount = pcount;
    za.spec = d_spec; za.spec_stride = spec_stride;
    za.hot_id = z->hot_id;
    za.nbase = S + (z->run2 ? 2 : 0);
    za.run2 = z->run2 ? 1 : 0;
    za.mma_passes = nullptr;
    za.sched = z->sched ? 1 : 0;
    if (mma) {
        if (!g_mma_passes) {
            CUDA_TRY(cudaMalloc((void**)&g_mma_passes, sizeof(unsigned long long)));
            CUDA_TRY(cudaMemset(g_mma_passes, 0, sizeof(unsigned long long)));
        }
        za.mma_passes = g_mma_passes;
    }
    DeviceBuf& d_vec = set->d_vec[pass];
    DeviceBuf& d_prog = set->d_prog[pass];
    const bool parts = set->parts_total > 0;
    if (parts) {      // parts of one long chunk: always segments (a later part has no start of its own), folded into set->d_parts
        if (K > 64) return fail(IMC_ERR_UNSUPPORTED, "parts mode supports K <= 64");
        if (seglen <= 0) seglen = std::max<long long>(256, std::min<long long>(4096, z->max_ntok / 64));
        seglen = (seglen + 15) / 16 * 16;
        if ((rc = zip_split(z, K, (int)seglen, &split, ns, set->part_first))) return rc;
        if ((rc = d_vec.reserve(sizeof(double) * (size_t)N * (size_t)split->nchains * (K + 1)))) return rc;
    } else if (seglen > 0 && K <= 64) {
        seglen = (seglen + 15) / 16 * 16;
        if (seglen < z->max_ntok) {
            if ((rc = zip_split(z, K, (int)seglen, &split))) return rc;
            const size_t vec_bytes = sizeof(double) * (size_t)N * ((size_t)split->nchains + split->nvec2) * (K + 1);
            if (vec_bytes > (size_t)1 << 30) split = nullptr;      // not worth a gigabyte of scratch
            else if ((rc = d_vec.reserve(vec_bytes))) return rc;
        }
    }
    za.nseg = 1; za.seglen = 0; za.carry = nullptr; za.carry_stride = 0; za.progress = nullptr;
    if (!split) {
        // ---- few work units per warp: walk every chunk in pieces that are separate, ordered work units (pipelined
        // mode of zip_run_unit), so that the end of the launch is not a wait for whole-chunk stragglers.  Config 2 has
        // 2.7 warp-loads per warp: SMs were idle 9 % of the launch.
        const int sms = g_ctx.sm_count > 0 ? g_ctx.sm_count : 148;
        const int cpw = 32 / plan.lanes, nquads = (ns + cpw - 1) / cpw;
        const double waves = (double)N * nquads / ((double)sms * plan.ctas_per_sm * (plan.threads / 32));
        long long nseg = 1;
        if (g_ctx.opt_zip_pipeline >= 2) nseg = g_ctx.opt_zip_pipeline;
        // (also where the warps are not all busy but the POINTS do not divide among the CTAs: 256 points x 8 warp-loads on 148 SMs ran
        // 3.26 ms as whole chunks, 2.62 ms in pieces -- the CTAs without a point of their own in the last round take pieces)
        else if (g_ctx.opt_zip_pipeline == 0 && (waves >= 1.0 || N > sms * plan.ctas_per_sm) && waves < 40.0)
            nseg = (long long)std::ceil(40.0 / std::max(waves, 1.0));
        nseg = std::min<long long>({nseg, 32, z->max_ntok / 256});
        if (nseg >= 2) {
            const int plen = (int)(((z->max_ntok + nseg - 1) / nseg + 15) / 16 * 16);
            nseg = (z->max_ntok + plen - 1) / plen;
            if (nseg >= 2) {
                const int cstride = (zip_tile(K) + 7) / 8 * 8 + 4;       // state registers of a chain's lanes (MMA form: 8 per n-tile) + exponent + flags
                if ((rc = d_vec.reserve(sizeof(double) * (size_t)N * ns * cstride))) return rc;
                if ((rc = d_prog.reserve(sizeof(int) * (size_t)N * ns))) return rc;
                CUDA_TRY(cudaMemsetAsync(d_prog.p, 0, sizeof(int) * (size_t)N * ns, st));
                za.nseg = (int)nseg; za.seglen = plen;
                za.carry = (double*)d_vec.p; za.carry_stride = cstride;
                za.progress = (int*)d_prog.p;
            }
        }
    }
    if (split) {
        za.chunks = (const ZipChunk*)split->chunks.p;
        za.nchunks = split->nchains;
        za.vec_out = (double*)d_vec.p;
        za.vec_stride = K + 1;
    }
    if (spec || !plist) {
        g_last_kernel = spec ? (split ? (mma ? (z->run2 ? "zip-spectral-mma2-segmented" : "zip-spectral-mma-segmented") : "zip-spectral-segmented")
                                      : (plan.lanes == 32 ? "zip-spectral-warp" : (mma ? (z->sched ? "zip-spectral-mma2-aligned" : (z->run2 ? "zip-spectral-mma2" : "zip-spectral-mma")) : "zip-spectral")))
                             : (split ? "zip-segmented" : (plan.lanes == 32 ? "zip-warp" : "zip"));
    }
    if ((rc = launch_zip(za, plan, st))) return rc;
    CUDA_TRY(cudaGetLastError());
    g_launches += 1;
    if (split) {
        double* vec2 = parts ? (double*)set->d_parts.p : za.vec_out + (size_t)N * split->nchains * za.vec_stride;
        for (int n0 = 0; n0 < N; n0 += 65535) {      // gridDim.y <= 65535
            const int nn = std::min(N - n0, 65535);
            // without a list blockIdx.y is the point itself: shift the per-point bases instead
            const size_t voff = plist ? 0 : (size_t)n0;
            if (split->n_level1 > 0) {
                zip_fold_kernel<<<dim3(split->n_level1, nn), 64, 0, st>>>(za.vec_out + voff * split->nchains * za.vec_stride, split->nchains,
                    vec2 + voff * split->nvec2 * za.vec_stride, split->nvec2, za.vec_stride, (const ZipFoldItem*)split->items1.p, K,
                    za.chain_out + voff * za.out_stride, za.out_stride, d_spec ? d_spec + voff * spec_stride : nullptr, spec_stride, plist, pcount, n0);
                CUDA_TRY(cudaGetLastError());
                g_launches += 1;
            }
            if (split->n_final > 0) {
                zip_fold_kernel<<<dim3(split->n_final, nn), 64, 0, st>>>(za.vec_out + voff * split->nchains * za.vec_stride, split->nchains,
                    vec2 + voff * split->nvec2 * za.vec_stride, split->nvec2, za.vec_stride, (const ZipFoldItem*)split->items2.p, K,
                    za.chain_out + voff * za.out_stride, za.out_stride, d_spec ? d_spec + voff * spec_stride : nullptr, spec_stride, plist, pcount, n0);
                CUDA_TRY(cudaGetLastError());
                g_launches += 1;
            }
        }
    }
    return IMC_OK;
}

static int parts_finish(imc_seqset* set, int N, int K, const double* d_spec, int spec_stride, const int* okflag, double* d_out,
                        cudaStream_t st);     // comm_host.inl: all-gather of the part blocks + zip_fold_parts_kernel

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
    if (ns == 0) return launch_chain_reduce(nullptr, 0, N, d_out, st);   // only empty chunks here: this rank adds 0 to every point
    if ((rc = set->d_chain.reserve(sizeof(double) * (size_t)N * ns))) return rc;

    int which = (int)g_ctx.opt_forward_kernel;
    if (set->parts_total > 0) {
        if (!zip_supported(K)) return fail(IMC_ERR_UNSUPPORTED, "parts mode runs on the zip kernel (K <= 40), not K = %d", K);
        which = KERNEL_ZIP;
    }
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
        // every chain result is written by exactly one of the passes below; a slot nobody writes must read as NaN, never as
        // the previous call's value
        CUDA_TRY(cudaMemsetAsync(set->d_chain.p, 0xff, sizeof(double) * (size_t)N * ns, st));
        const bool parts = set->parts_total > 0;
        if (parts) {        // block of this rank: vec[N][ns * K][K + 1], then run_sites[ns]
            if (N > MAX_POINTS_PARTS) return fail(IMC_ERR_UNSUPPORTED, "parts mode: at most %d parameter points per call", MAX_POINTS_PARTS);
            const size_t vecs = (size_t)N * ns * K * (K + 1);
            if ((rc = set->d_parts.reserve(sizeof(double) * (vecs + ns)))) return rc;
            CUDA_TRY(cudaMemsetAsync(set->d_parts.p, 0xff, sizeof(double) * vecs, st));
        }
        // Spectral form where the run symbol's runs carry most of the compression: zip_spectral_kernel diagonalises C_r of
        // every point and sorts the points into those it could serve (ok list) and the others (plain form, bad list).
        bool use_spec = g_ctx.opt_zip_spectral == 1;
        if (g_ctx.opt_zip_spectral == 0) use_spec = set->run_tokens_full * 115 < set->zip_tokens_full * 100;
        if (use_spec && (set->nsym < 2 || K > 64)) use_spec = false;
        ZipPlan probe;
        if (use_spec && zip_plan(K, S, set->run_merges.size(), &probe, 0, true) != IMC_OK) use_spec = false;
        if (!use_spec) {
            g_last_kernel = "zip";
            if ((rc = zip_pass(set, N, K, S, d_pi, d_T, d_E, false, nullptr, nullptr, nullptr, 0, st))) return rc;
            if (parts) return parts_finish(set, N, K, nullptr, 0, nullptr, d_out, st);
            return launch_chain_reduce((const double*)set->d_chain.p, ns, N, d_out, st);
        }
        // Two-run form of the MMA shape (the second run symbol -- missing data -- diagonalised as well): worth it where the
        // dictionary that fits is small (K >= 32), judged by the expected DMMA work = tokens x passes per warp-step.
        bool run2 = false, sched = false;
        if (g_ctx.opt_zip_run2 != 2 && !parts && zip_mma_tile(zip_tile(K)) && g_ctx.opt_zip_mma != 2 && g_ctx.opt_zip_lanes == 0) {
            ZipPlan p1, p2;
            ZipDevice *z1 = nullptr, *z2 = nullptr;
            // only a stream with many cold passes has something to gain: the second encoding is not even prepared otherwise
            if (zip_plan(K, S, set->run_merges.size(), &p1, 0, true, true, false) == IMC_OK && zip_device(set, p1.M, &z1, true, false) == IMC_OK &&
                (g_ctx.opt_zip_run2 == 1 || z1->est_passes >= 2.4)) {
                if ((rc = seqset_run2_prepare(set))) return rc;
                if (set->run2_state == 1 && zip_plan(K, S, set->run2_merges.size(), &p2, 0, true, true, true) == IMC_OK &&
                    zip_device(set, p2.M, &z2, true, true) == IMC_OK)
                    run2 = g_ctx.opt_zip_run2 == 1 || (double)z2->total_tokens * z2->est_passes < 0.93 * (double)z1->total_tokens * z1->est_passes;
            }
            // Aligned form: the two-run streams of every warp-load merged into one schedule with a single entry per warp-step; its
            // pass count is exact, the lock-step ones are estimates (0.9: it pays a little more per pass for the no-op bookkeeping)
            if (z1 && g_ctx.opt_zip_align != 2 && (g_ctx.opt_zip_align == 1 || ((long long)N * ns >= 1024 && z1->est_passes >= 1.5))) {
                ZipPlan p3;
                ZipDevice* z3 = nullptr;
                if ((rc = seqset_run2_prepare(set))) return rc;
                g_mma_shape_hint = 2;
                const int planned3 = set->run2_state == 1 ? zip_plan(K, S, set->run2_merges.size(), &p3, 0, true, true, true) : IMC_ERR_UNSUPPORTED;
                g_mma_shape_hint = 0;
                if (planned3 == IMC_OK && zip_device(set, p3.M, &z3, true, true, true) == IMC_OK) {
                    const ZipDevice* zl = run2 && z2 ? z2 : z1;
                    // Clocks per parameter point and SM.  A warp-step holds the FP64 pipe of its sub-partition for P clocks (its passes
                    // of KT x NT DMMAs of 16 clocks, the DMULs of the table factors) and takes a lone warp ~120 + 2 P clocks (measured
                    // at K=10: 570 in lock step, ~330 aligned); with W warps of the SM busy a round of one step each takes
                    // max(latency, W / 4 x P / 0.8), and a point is steps x warp-loads / W rounds.  Reproduces configs 2-5 and the
                    // few-chunk cases of profiles/r02_aligned_form.txt within 15 %; the aligned form has to win by 5 %.
                    const int tile = zip_tile(K), nq = (ns + 7) / 8;
                    const double pass = 16.0 * ((tile + 3) / 4) * ((tile + 7) / 8), dmul = 8.0 * ((tile + 7) / 8);
                    auto point_cost = [&](const ZipPlan& p, double quad_steps, double passes) {
                        const int W = p.ctas_per_sm * std::min(p.threads / 32, nq);
                        const double P = passes * pass + dmul;
                        return quad_steps * std::max(ZIP_STEP_LATENCY + 2.0 * P, W / 4.0 * P / 0.8) / W;
                    };
                    const double lock = point_cost(run2 && z2 ? p2 : p1, (double)zl->pass_cost / zl->est_passes, zl->est_passes);
                    const double aligned = point_cost(p3, (double)z3->pass_cost, 1.0);
                    sched = g_ctx.opt_zip_align == 1 || aligned < 0.95 * lock;
                    if (getenv("IMC_TRACE_PLAN"))
                        fprintf(stderr, "imc plan: K=%d lock-step steps %.0f x %.2f passes (%s), aligned steps %lld -> cost %.3g vs %.3g: %s\n", K,
                                (double)zl->pass_cost / zl->est_passes, zl->est_passes, zl->run2 ? "two-run" : "one-run", z3->pass_cost, lock, aligned,
                                sched ? "aligned" : "lock step");
                }
            }
            if (sched) run2 = true;
        }
        g_want_run2 = run2;
        g_want_sched = sched;
        const int sstride = zip_spec_stride(K, S, run2);
        if ((rc = set->d_spec.reserve(sizeof(double) * (size_t)N * sstride))) return rc;
        if ((rc = set->d_lists.reserve(sizeof(int) * ((size_t)3 * N + 2)))) return rc;
        int* lists = (int*)set->d_lists.p;
        CUDA_TRY(cudaMemsetAsync(lists, 0, sizeof(int) * 2, st));
        NvtxRange nvtx_spec("imc: spectral forward (prepare + passes)");
        ZipSpecArgs sa;
        sa.N = N; sa.K = K; sa.S = S; sa.run_sym = set->fold_sym;
        sa.run_sym2 = run2 ? set->run_sym2 : -1;
        sa.pi = d_pi; sa.T = d_T; sa.E = d_E;
        sa.spec = (double*)set->d_spec.p; sa.spec_stride = sstride;
        sa.counts = lists; sa.ok_list = lists + 2; sa.bad_list = lists + 2 + N; sa.okflag = lists + 2 + 2 * N;
        sa.force_bad = g_ctx.opt_zip_spectral_force_bad ? 1 : 0;
        sa.point_base = 0;
        {
            static size_t attr_max = 0;
            const size_t sm = zip_spec_smem(K, run2);
            if (sm > 48 * 1024 && sm > attr_max) {
                CUDA_TRY(cudaFuncSetAttribute(zip_spectral_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
                attr_max = sm;
            }
            for (int n0 = 0; n0 < N; n0 += 65535) {       // gridDim.x is not limited, but keep launches bounded
                ZipSpecArgs sb = sa;
                sb.N = std::min(N - n0, 65535);
                sb.pi += (size_t)n0 * K; sb.T += (size_t)n0 * K * K; sb.E += (size_t)n0 * K * S;
                sb.spec += (size_t)n0 * sstride;
                sb.point_base = n0;
                zip_spectral_kernel<<<sb.N, SPEC_THREADS, sm, st>>>(sb);
                CUDA_TRY(cudaGetLastError());
                g_launches += 1;
            }
        }
        g_last_kernel = "zip-spectral";
        if ((rc = zip_pass(set, N, K, S, d_pi, d_T, d_E, true, sa.ok_list, sa.counts, sa.spec, sstride, st))) return rc;
        if ((rc = zip_pass(set, N, K, S, d_pi, d_T, d_E, false, sa.bad_list, sa.counts + 1, nullptr, 0, st))) return rc;
        if (parts) return parts_finish(set, N, K, sa.spec, sstride, sa.okflag, d_out, st);
        return launch_chain_reduce((const double*)set->d_chain.p, ns, N, d_out, st);
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
    g_launches += 1;
    return launch_chain_reduce(a.chain_out, ns, N, d_out, st);
}

#include "comm_host.inl"

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
    if (!strcmp(key, "zip_pipeline")) { if (value < 0 || value > 32) return fail(IMC_ERR_INVALID, "zip_pipeline must be in [0, 32]"); g_ctx.opt_zip_pipeline = value; return IMC_OK; }
    if (!strcmp(key, "zip_spectral")) { if (value < 0 || value > 2) return fail(IMC_ERR_INVALID, "zip_spectral must be 0 (auto), 1 (always) or 2 (never)"); g_ctx.opt_zip_spectral = value; return IMC_OK; }
    if (!strcmp(key, "zip_spectral_force_bad")) { g_ctx.opt_zip_spectral_force_bad = value ? 1 : 0; return IMC_OK; }
    if (!strcmp(key, "zip_mma")) { if (value < 0 || value > 2) return fail(IMC_ERR_INVALID, "zip_mma must be 0 (auto), 1 (always) or 2 (never)"); g_ctx.opt_zip_mma = value; return IMC_OK; }
    if (!strcmp(key, "zip_align")) { if (value < 0 || value > 2) return fail(IMC_ERR_INVALID, "zip_align must be 0, 1 or 2"); g_ctx.opt_zip_align = value; return IMC_OK; }
    if (!strcmp(key, "zip_mma_shape")) { if (value < 0 || value > 4) return fail(IMC_ERR_INVALID, "zip_mma_shape must be in [0, 4]"); g_ctx.opt_zip_mma_shape = value; return IMC_OK; }
    if (!strcmp(key, "zip_run2")) { if (value < 0 || value > 2) return fail(IMC_ERR_INVALID, "zip_run2 must be 0 (auto), 1 (always) or 2 (never)"); g_ctx.opt_zip_run2 = value; return IMC_OK; }
    if (!strcmp(key, "comm_fused")) { g_ctx.opt_comm_fused = value ? 1 : 0; return IMC_OK; }
    if (!strcmp(key, "comm_enabled")) { g_ctx.opt_comm_enabled = value ? 1 : 0; return IMC_OK; }
    if (!strcmp(key, "comm_timeout_ms")) { if (value < 1) return fail(IMC_ERR_INVALID, "comm_timeout_ms must be >= 1"); g_ctx.opt_comm_timeout_ms = value; return IMC_OK; }
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
    if (!strcmp(key, "zip_pipeline")) { *value_out = g_ctx.opt_zip_pipeline; return IMC_OK; }
    if (!strcmp(key, "zip_spectral")) { *value_out = g_ctx.opt_zip_spectral; return IMC_OK; }
    if (!strcmp(key, "zip_spectral_force_bad")) { *value_out = g_ctx.opt_zip_spectral_force_bad; return IMC_OK; }
    if (!strcmp(key, "zip_mma")) { *value_out = g_ctx.opt_zip_mma; return IMC_OK; }
    if (!strcmp(key, "zip_align")) { *value_out = g_ctx.opt_zip_align; return IMC_OK; }
    if (!strcmp(key, "zip_mma_shape")) { *value_out = g_ctx.opt_zip_mma_shape; return IMC_OK; }
    if (!strcmp(key, "zip_run2")) { *value_out = g_ctx.opt_zip_run2; return IMC_OK; }
    if (!strcmp(key, "comm_fused")) { *value_out = g_ctx.opt_comm_fused; return IMC_OK; }
    if (!strcmp(key, "comm_enabled")) { *value_out = g_ctx.opt_comm_enabled; return IMC_OK; }
    if (!strcmp(key, "comm_timeout_ms")) { *value_out = g_ctx.opt_comm_timeout_ms; return IMC_OK; }
    return fail(IMC_ERR_INVALID, "unknown option '%s'", key);
}
extern "C" int64_t imc_kernel_launches(void) { return g_launches.load(); }
extern "C" const char* imc_last_forward_kernel(void) { return g_last_kernel; }

// ------------------------------------------------------------------------------------------ model build
#include "model_host.inl"
