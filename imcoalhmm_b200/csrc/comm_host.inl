// Multi-GPU plumbing (included by imc_lib.cu): NCCL resolved at run time, the one all-reduce, the forward wrapper.

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
    decltype(&ncclAllGather) AllGather = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
};
static NcclApi g_nccl;
static ncclComm_t g_comm = nullptr;
static int g_comm_nranks = 1, g_comm_rank = 0;

// Mailboxes for the fused reduce + all-reduce kernel (reduce_chains_peer_kernel): one cudaMalloc'ed buffer per rank, mapped
// into every other rank's address space through CUDA IPC (same node: NVLink / NVSwitch peer-to-peer stores).
struct PeerBox {
    bool active = false;
    int nmax = 0;
    unsigned epoch = 0;
    void* local = nullptr;
    void* mapped[MAX_PEERS] = {};      // mapped[r] = rank r's mailbox in this process (mapped[rank] == local)
    unsigned* counter = nullptr;
    unsigned* status = nullptr;        // pinned, mapped host word written by reduce_chains_peer_kernel on a timeout
    unsigned* status_dev = nullptr;    // its device address
    bool broken = false;               // a collective timed out: every later collective call fails until imc_comm_destroy
    size_t val_bytes() const { return sizeof(double) * 2 * (size_t)g_comm_nranks * nmax; }
};
static PeerBox g_peer;
static thread_local bool g_reduced_over_ranks = false;   // set by launch_chain_reduce when the fused kernel did the all-reduce

static int nccl_load() {
    if (g_nccl.handle) return IMC_OK;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return fail(IMC_ERR_UNSUPPORTED, "libnccl.so.2 cannot be loaded: %s", dlerror());
    g_nccl.GetUniqueId = (decltype(g_nccl.GetUniqueId))dlsym(h, "ncclGetUniqueId");
    g_nccl.CommInitRank = (decltype(g_nccl.CommInitRank))dlsym(h, "ncclCommInitRank");
    g_nccl.AllReduce = (decltype(g_nccl.AllReduce))dlsym(h, "ncclAllReduce");
    g_nccl.AllGather = (decltype(g_nccl.AllGather))dlsym(h, "ncclAllGather");
    g_nccl.CommDestroy = (decltype(g_nccl.CommDestroy))dlsym(h, "ncclCommDestroy");
    g_nccl.GetErrorString = (decltype(g_nccl.GetErrorString))dlsym(h, "ncclGetErrorString");
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllReduce || !g_nccl.AllGather || !g_nccl.CommDestroy || !g_nccl.GetErrorString)
        return fail(IMC_ERR_UNSUPPORTED, "libnccl.so.2 lacks an expected symbol");
    g_nccl.handle = h;
    return IMC_OK;
}
#define NCCL_TRY(x)                                                                                       \
    do {                                                                                                  \
        ncclResult_t r_ = (x);                                                                            \
        if (r_ != ncclSuccess) return fail(IMC_ERR_CUDA, "%s failed: %s", #x, g_nccl.GetErrorString(r_)); \
    } while (0)

static void peer_teardown() {
    for (int r = 0; r < MAX_PEERS; ++r) {
        if (g_peer.mapped[r] && g_peer.mapped[r] != g_peer.local) cudaIpcCloseMemHandle(g_peer.mapped[r]);
        g_peer.mapped[r] = nullptr;
    }
    if (g_peer.local) cudaFree(g_peer.local);
    if (g_peer.counter) cudaFree(g_peer.counter);
    if (g_peer.status) cudaFreeHost(g_peer.status);
    g_peer = PeerBox();
}

// Called by every rank right after the communicator exists.  Failure to map any peer (ranks on different nodes, IPC not
// permitted, "comm_fused" 0 on some rank, an allocation that failed) is not an error: all ranks then agree to keep the NCCL
// all-reduce.  Every rank takes part in BOTH exchange rounds whatever happened locally -- a rank that left early would
// leave the others blocked inside NCCL -- and carries its local verdict in an "ok" bit instead.
static int peer_setup(int nranks, int rank) {
    cudaStream_t st = g_ctx.stream;
    int ok = (g_ctx.opt_comm_fused && nranks <= MAX_PEERS) ? 1 : 0;
    g_peer.nmax = 1 << 16;
    const size_t bytes = g_peer.val_bytes() + 128 * (size_t)nranks;
    cudaIpcMemHandle_t mine;
    memset(&mine, 0, sizeof mine);
    if (ok && (cudaMalloc(&g_peer.local, bytes) != cudaSuccess || cudaMalloc((void**)&g_peer.counter, sizeof(unsigned)) != cudaSuccess ||
               cudaMemset(g_peer.local, 0, bytes) != cudaSuccess || cudaMemset(g_peer.counter, 0, sizeof(unsigned)) != cudaSuccess ||
               cudaHostAlloc((void**)&g_peer.status, sizeof(unsigned), cudaHostAllocMapped) != cudaSuccess ||
               cudaHostGetDevicePointer((void**)&g_peer.status_dev, g_peer.status, 0) != cudaSuccess ||
               cudaIpcGetMemHandle(&mine, g_peer.local) != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess)) {
        ok = 0;
        cudaGetLastError();
    }
    if (ok) *g_peer.status = 0u;
    // round 1: exchange the handles (and whether everybody has one) with the communicator itself
    const size_t slot = sizeof(cudaIpcMemHandle_t) + 8;
    unsigned char* d_all = nullptr;
    std::vector<unsigned char> h_all(slot * std::max(nranks, 1), 0);
    int local_rc = IMC_OK;          // first local CUDA / NCCL failure; reported after both rounds
    auto note = [&](bool good, const char* what) { if (!good && local_rc == IMC_OK) local_rc = fail(IMC_ERR_CUDA, "peer setup: %s failed", what); if (!good) ok = 0; };
    note(cudaMalloc((void**)&d_all, slot * nranks) == cudaSuccess, "cudaMalloc");
    memcpy(h_all.data() + slot * rank, &mine, sizeof mine);
    h_all[slot * rank + sizeof mine] = (unsigned char)ok;
    if (d_all) {
        note(cudaMemcpy(d_all + slot * rank, h_all.data() + slot * rank, slot, cudaMemcpyHostToDevice) == cudaSuccess, "cudaMemcpy");
        note(g_nccl.AllGather(d_all + slot * rank, d_all, slot, ncclChar, g_comm, st) == ncclSuccess, "ncclAllGather");
        note(cudaStreamSynchronize(st) == cudaSuccess, "cudaStreamSynchronize");
        note(cudaMemcpy(h_all.data(), d_all, slot * nranks, cudaMemcpyDeviceToHost) == cudaSuccess, "cudaMemcpy");
    }
    for (int r = 0; r < nranks; ++r) ok = ok && h_all[slot * r + sizeof mine];
    if (ok) {
        g_peer.mapped[rank] = g_peer.local;
        for (int r = 0; r < nranks && ok; ++r) {
            if (r == rank) continue;
            cudaIpcMemHandle_t h;
            memcpy(&h, h_all.data() + slot * r, sizeof h);
            if (cudaIpcOpenMemHandle(&g_peer.mapped[r], h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
                g_peer.mapped[r] = nullptr;
                ok = 0;
                cudaGetLastError();
            }
        }
    }
    // round 2: did every rank map every mailbox?
    if (d_all) {
        int* d_ok = (int*)d_all;
        note(cudaMemcpy(d_ok, &ok, sizeof(int), cudaMemcpyHostToDevice) == cudaSuccess, "cudaMemcpy");
        note(g_nccl.AllReduce(d_ok, d_ok, 1, ncclInt, ncclMin, g_comm, st) == ncclSuccess, "ncclAllReduce");
        note(cudaStreamSynchronize(st) == cudaSuccess, "cudaStreamSynchronize");
        int all_ok = 0;
        note(cudaMemcpy(&all_ok, d_ok, sizeof(int), cudaMemcpyDeviceToHost) == cudaSuccess, "cudaMemcpy");
        ok = ok && all_ok;
        cudaFree(d_all);
    }
    if (!ok) { peer_teardown(); return local_rc; }
    g_peer.active = true;
    return IMC_OK;
}

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
    g_comm_nranks = nranks;
    g_comm_rank = rank;
    return peer_setup(nranks, rank);
}

extern "C" int imc_comm_destroy(void) {
    if (g_comm && g_ctx.pid == getpid()) {
        if (g_peer.local) {
            // nobody may unmap or free a mailbox while another rank can still write into it: finish all local work, then
            // meet the other ranks once more
            cudaDeviceSynchronize();
            int* d_one = nullptr;
            if (cudaMalloc((void**)&d_one, sizeof(int)) == cudaSuccess) {
                cudaMemset(d_one, 0, sizeof(int));
                g_nccl.AllReduce(d_one, d_one, 1, ncclInt, ncclSum, g_comm, g_ctx.stream);
                cudaStreamSynchronize(g_ctx.stream);
                cudaFree(d_one);
            }
            peer_teardown();
        }
        g_nccl.CommDestroy(g_comm);
    }
    g_comm = nullptr;
    g_comm_nranks = 1;
    g_comm_rank = 0;
    g_peer = PeerBox();
    return IMC_OK;
}

extern "C" int imc_comm_info(int* nranks, int* rank, int* fused) {
    if (nranks) *nranks = g_comm ? g_comm_nranks : 1;
    if (rank) *rank = g_comm ? g_comm_rank : 0;
    if (fused) *fused = (g_comm && g_peer.active) ? 1 : 0;
    return IMC_OK;
}

// out[n] = sum over this rank's chains -- and, when the mailboxes are mapped, over all ranks in the same kernel
static int comm_check() {      // a collective that timed out leaves the ranks out of step: nothing collective may follow
    if (g_peer.active && !g_peer.broken && g_peer.status && (*((volatile unsigned*)g_peer.status) & 0x80000000u)) g_peer.broken = true;
    if (g_peer.broken)
        return fail(IMC_ERR_CUDA, "the fused all-reduce timed out waiting for rank %u (dead peer, or ranks issued different call sequences); "
                                  "results of that call are NaN and the communicator is unusable: call imc_comm_destroy",
                    g_peer.status ? (*g_peer.status & 0x7fffffffu) : 0u);
    return IMC_OK;
}

static int launch_chain_reduce(const double* chain, int ns, int N, double* d_out, cudaStream_t st) {
    NvtxRange nvtx_reduce("imc: chain reduction (+ fused all-reduce)");
    g_reduced_over_ranks = false;
    if (g_comm && g_peer.active && g_ctx.opt_comm_enabled && N <= g_peer.nmax) {
        int rc = comm_check();
        if (rc) return rc;
        // consecutive collectives of this process run in issue order whatever streams they were enqueued on (the mailbox
        // halves and the block counter assume it)
        static HandleSerial comm_serial;
        CallGuard guard(comm_serial, st);
        PeerReduceArgs pa;
        pa.nranks = g_comm_nranks; pa.rank = g_comm_rank; pa.nmax = g_peer.nmax;
        pa.epoch = ++g_peer.epoch;
        for (int r = 0; r < g_comm_nranks; ++r) {
            pa.box[r] = (double*)g_peer.mapped[r];
            pa.flag[r] = (unsigned*)((unsigned char*)g_peer.mapped[r] + g_peer.val_bytes());
        }
        pa.counter = g_peer.counter;
        pa.status = g_peer.status_dev;
        pa.timeout_ns = (unsigned long long)std::max<long long>(1, g_ctx.opt_comm_timeout_ms) * 1000000ull;
        reduce_chains_peer_kernel<<<N, 256, 0, st>>>(chain, ns, d_out, pa);
        g_reduced_over_ranks = true;
    } else {
        reduce_chains_kernel<<<N, 256, 0, st>>>(chain, ns, d_out);
    }
    CUDA_TRY(cudaGetLastError());
    g_launches += 1;
    return IMC_OK;
}

// Parts mode (one long chunk cut over the ranks, SURVEY 8e "fewer chunks than GPUs"): the passes have folded this rank's parts
// into set->d_parts; write the parts' run-site counts behind them, all-gather the blocks of all ranks (one ncclAllGather: the
// only collective of such a call) and fold alpha <- P_part alpha over all parts in order.  Every rank ends with the same
// logL[N] of the whole chunk.  A single process may hold all parts (no communicator needed): the gather is then the block itself.
static int parts_finish(imc_seqset* set, int N, int K, const double* d_spec, int spec_stride, const int* okflag, double* d_out,
                        cudaStream_t st) {
    const int ns = (int)set->streams.size();
    const int nranks = set->parts_total / ns;
    const size_t vecs = (size_t)N * ns * K * (K + 1), block = vecs + ns;
    std::vector<double> rs(ns, 0.0);
    {
        // run sites of every local part (stream k holds chunk set_chunk_of_stream(k)); taken from the spectral streams
        ZipPlan plan;
        ZipDevice* z = nullptr;
        if (d_spec && zip_plan(K, set->nsym, set->run_merges.size(), &plan, 0, true) == IMC_OK && zip_device(set, plan.M, &z, true) == IMC_OK)
            for (const ZipChunk& ch : z->host_chunks) rs[ch.continues - set->part_first] = (double)ch.run_sites;
    }
    CUDA_TRY(cudaMemcpyAsync((double*)set->d_parts.p + vecs, rs.data(), sizeof(double) * ns, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaStreamSynchronize(st));        // rs is a stack buffer; parts-mode calls are rare and long
    const double* gathered = (const double*)set->d_parts.p;
    if (nranks > 1) {
        if (!g_comm || g_comm_nranks != nranks || g_comm_rank * ns != set->part_first)
            return fail(IMC_ERR_INVALID, "parts mode: this set holds parts %d..%d of %d, which needs a communicator of %d ranks with this process as rank %d",
                        set->part_first, set->part_first + ns - 1, set->parts_total, nranks, set->part_first / ns);
        int rc = set->d_gather.reserve(sizeof(double) * block * nranks);
        if (rc) return rc;
        NCCL_TRY(g_nccl.AllGather(set->d_parts.p, set->d_gather.p, block, ncclDouble, g_comm, st));
        gathered = (const double*)set->d_gather.p;
    }
    zip_fold_parts_kernel<<<N, 64, 0, st>>>(gathered, block, N, K, ns, set->parts_total, d_spec, spec_stride, okflag, d_out);
    CUDA_TRY(cudaGetLastError());
    g_launches += 1;
    g_reduced_over_ranks = true;       // the result already is the whole chunk's: no all-reduce on top
    return IMC_OK;
}

// this rank's partial log-likelihoods, then the sum over ranks when a communicator exists
static const int MAX_POINTS_PER_LAUNCH = 32768;     // several kernels put the parameter point on gridDim.y (<= 65535)

static int forward_dev(imc_seqset* set, int N, int K, int S, const double* d_pi, const double* d_T, const double* d_E,
                       double* d_out, cudaStream_t st) {
    if (!set) return fail(IMC_ERR_INVALID, "NULL set");
    NvtxRange nvtx_fwd("imc: forward");
    CallGuard guard(set->serial, st);
    if (N > MAX_POINTS_PER_LAUNCH) {       // large batches (MCMC / swarm populations) run as slices, each with its own all-reduce
        for (int n0 = 0; n0 < N; n0 += MAX_POINTS_PER_LAUNCH) {
            const int nn = std::min(MAX_POINTS_PER_LAUNCH, N - n0);
            int rc = forward_dev(set, nn, K, S, d_pi + (size_t)n0 * K, d_T + (size_t)n0 * K * K, d_E + (size_t)n0 * K * S, d_out + n0, st);
            if (rc) return rc;
        }
        return IMC_OK;
    }
    int rc = forward_local_dev(set, N, K, S, d_pi, d_T, d_E, d_out, st);
    if (rc || !g_comm || N <= 0 || !g_ctx.opt_comm_enabled || g_reduced_over_ranks) return rc;
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
    CallGuard guard(set->serial, st);       // the staging buffers belong to this call until its results are on the host
    CUDA_TRY(cudaMemcpyAsync(set->d_pi.p, pi, npi * sizeof(double), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(set->d_T.p, T, nT * sizeof(double), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(set->d_E.p, E, nE * sizeof(double), cudaMemcpyHostToDevice, st));
    rc = forward_dev(set, N, K, S, (const double*)set->d_pi.p, (const double*)set->d_T.p, (const double*)set->d_E.p,
                     (double*)set->d_out.p, st);
    if (rc) return rc;
    CUDA_TRY(cudaMemcpyAsync(out, set->d_out.p, (size_t)N * sizeof(double), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return g_comm ? comm_check() : IMC_OK;
}

extern "C" int imc_forward(imc_seqset* set, int K, int S, const double* pi, const double* T, const double* E,
                           double* logL_out) {
    return imc_forward_batch(set, 1, K, S, pi, T, E, logL_out);
}

