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
    g_peer = PeerBox();
}

// Called by every rank right after the communicator exists.  Failure to map any peer (ranks on different nodes, IPC not
// permitted) is not an error: all ranks then agree to keep the NCCL all-reduce.
static int peer_setup(int nranks, int rank) {
    if (!g_ctx.opt_comm_fused || nranks > MAX_PEERS) return IMC_OK;
    cudaStream_t st = g_ctx.stream;
    g_peer.nmax = 1 << 16;
    const size_t bytes = g_peer.val_bytes() + 128 * (size_t)nranks;
    int ok = 1;
    cudaIpcMemHandle_t mine;
    memset(&mine, 0, sizeof mine);
    if (cudaMalloc(&g_peer.local, bytes) != cudaSuccess || cudaMalloc((void**)&g_peer.counter, sizeof(unsigned)) != cudaSuccess ||
        cudaMemset(g_peer.local, 0, bytes) != cudaSuccess || cudaMemset(g_peer.counter, 0, sizeof(unsigned)) != cudaSuccess ||
        cudaIpcGetMemHandle(&mine, g_peer.local) != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) {
        ok = 0;
        cudaGetLastError();
    }
    // exchange the handles (and whether everybody has one) with the communicator itself
    const size_t slot = sizeof(cudaIpcMemHandle_t) + 8;
    unsigned char* d_all = nullptr;
    CUDA_TRY(cudaMalloc((void**)&d_all, slot * nranks));
    std::vector<unsigned char> h_all(slot * nranks, 0);
    memcpy(h_all.data() + slot * rank, &mine, sizeof mine);
    h_all[slot * rank + sizeof mine] = (unsigned char)ok;
    CUDA_TRY(cudaMemcpy(d_all + slot * rank, h_all.data() + slot * rank, slot, cudaMemcpyHostToDevice));
    NCCL_TRY(g_nccl.AllGather(d_all + slot * rank, d_all, slot, ncclChar, g_comm, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    CUDA_TRY(cudaMemcpy(h_all.data(), d_all, slot * nranks, cudaMemcpyDeviceToHost));
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
    // second round: did every rank map every mailbox?
    int* d_ok = (int*)d_all;
    CUDA_TRY(cudaMemcpy(d_ok, &ok, sizeof(int), cudaMemcpyHostToDevice));
    NCCL_TRY(g_nccl.AllReduce(d_ok, d_ok, 1, ncclInt, ncclMin, g_comm, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    CUDA_TRY(cudaMemcpy(&ok, d_ok, sizeof(int), cudaMemcpyDeviceToHost));
    cudaFree(d_all);
    if (!ok) { peer_teardown(); return IMC_OK; }
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
static int launch_chain_reduce(const double* chain, int ns, int N, double* d_out, cudaStream_t st) {
    g_reduced_over_ranks = false;
    if (g_comm && g_peer.active && g_ctx.opt_comm_enabled && N <= g_peer.nmax) {
        PeerReduceArgs pa;
        pa.nranks = g_comm_nranks; pa.rank = g_comm_rank; pa.nmax = g_peer.nmax;
        pa.epoch = ++g_peer.epoch;
        for (int r = 0; r < g_comm_nranks; ++r) {
            pa.box[r] = (double*)g_peer.mapped[r];
            pa.flag[r] = (unsigned*)((unsigned char*)g_peer.mapped[r] + g_peer.val_bytes());
        }
        pa.counter = g_peer.counter;
        reduce_chains_peer_kernel<<<N, 256, 0, st>>>(chain, ns, d_out, pa);
        g_reduced_over_ranks = true;
    } else {
        reduce_chains_kernel<<<N, 256, 0, st>>>(chain, ns, d_out);
    }
    CUDA_TRY(cudaGetLastError());
    g_launches += 1;
    return IMC_OK;
}

// this rank's partial log-likelihoods, then the sum over ranks when a communicator exists
static int forward_dev(imc_seqset* set, int N, int K, int S, const double* d_pi, const double* d_T, const double* d_E,
                       double* d_out, cudaStream_t st) {
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

