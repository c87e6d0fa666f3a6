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

