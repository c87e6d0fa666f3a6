// Host side of the batched model build: two-locus ancestry state spaces, model descriptors, launchers.
// Included by imc_lib.cu (after the context / error helpers).
//
// State spaces follow the semantics of /root/reference/src/IMCoalHMM/statespace_generator.py:72-156 and
// state_spaces.py:7-116 (coloured-Petri-net exploration with R / C / M transitions), re-designed as a
// bitmask BFS: a lineage is (population, left-sample mask, right-sample mask), a state is the sorted set of
// its lineages.  States are numbered in the canonical order used by tests/golden/statespaces.json (the
// reference's own numbering is hash-order dependent and never compared).

#include <map>
#include <set>
#include <tuple>

namespace imc {

struct HostSpace {
    int n = 0;
    std::vector<std::vector<uint8_t>> states;     // canonical order; each state = sorted lineage codes
    std::vector<int> edges;                       // (src, dst, label) triples
    std::vector<int> B, L, R, E;
    int i11 = -1, i12 = -1, i22 = -1;
};

static inline uint8_t lineage(int pop, int l, int r) { return (uint8_t)((pop << 4) | (l << 2) | r); }
static inline int lin_pop(uint8_t t) { return t >> 4; }
static inline int lin_l(uint8_t t) { return (t >> 2) & 3; }
static inline int lin_r(uint8_t t) { return t & 3; }
// python orders the sample tuples () < (1,) < (1,2) < (2,)
static inline int mask_rank(int m) { static const int r[4] = {0, 1, 3, 2}; return r[m]; }
static inline int lin_key(uint8_t t) { return (lin_pop(t) << 4) | (mask_rank(lin_l(t)) << 2) | mask_rank(lin_r(t)); }

static bool state_less(const std::vector<uint8_t>& a, const std::vector<uint8_t>& b) {
    const size_t n = std::min(a.size(), b.size());
    for (size_t i = 0; i < n; ++i) {
        if (lin_key(a[i]) != lin_key(b[i])) return lin_key(a[i]) < lin_key(b[i]);
    }
    return a.size() < b.size();
}
static void canon(std::vector<uint8_t>& s) {
    std::sort(s.begin(), s.end(), [](uint8_t x, uint8_t y) { return lin_key(x) < lin_key(y); });
}

// kind: SP_ISO (populations 1,2; R,C), SP_SINGLE (population 0; R,C), SP_MIG (populations 1,2; R,M,C)
static HostSpace build_space(int kind) {
    std::vector<uint8_t> init;
    for (int sample = 1; sample <= 2; ++sample)
        init.push_back(lineage(kind == SP_SINGLE ? 0 : sample, sample, sample));
    canon(init);
    auto cmp = [](const std::vector<uint8_t>& a, const std::vector<uint8_t>& b) { return state_less(a, b); };
    std::set<std::vector<uint8_t>, decltype(cmp)> seen(cmp);
    std::vector<std::vector<uint8_t>> todo{init};
    seen.insert(init);
    struct RawEdge { std::vector<uint8_t> src, dst; int label; };
    std::vector<RawEdge> raw;
    while (!todo.empty()) {
        std::vector<uint8_t> s = todo.back();
        todo.pop_back();
        auto emit = [&](std::vector<uint8_t> d, int label) {
            canon(d);
            raw.push_back({s, d, label});
            if (seen.insert(d).second) todo.push_back(d);
        };
        for (size_t a = 0; a < s.size(); ++a) {
            const uint8_t t = s[a];
            const int pop = lin_pop(t), l = lin_l(t), r = lin_r(t);
            if (l && r) {   // recombination splits a lineage carrying both loci (statespace_generator.py:160-172)
                std::vector<uint8_t> d(s);
                d.erase(d.begin() + a);
                d.push_back(lineage(pop, l, 0));
                d.push_back(lineage(pop, 0, r));
                emit(d, LBL_R);
            }
            if (kind == SP_MIG) {   // migration to the other population (state_spaces.py:75-79)
                std::vector<uint8_t> d(s);
                d[a] = lineage(3 - pop, l, r);
                emit(d, pop == 1 ? LBL_M12 : LBL_M21);
            }
            for (size_t b = 0; b < a; ++b) {   // coalescence of two lineages in the same population (:175-185)
                const uint8_t t2 = s[b];
                if (lin_pop(t2) != pop) continue;
                std::vector<uint8_t> d;
                for (size_t k = 0; k < s.size(); ++k) if (k != a && k != b) d.push_back(s[k]);
                d.push_back(lineage(pop, l | lin_l(t2), r | lin_r(t2)));
                emit(d, pop == 2 ? LBL_C2 : LBL_C1);
            }
        }
    }
    HostSpace sp;
    sp.states.assign(seen.begin(), seen.end());   // canonical order
    sp.n = (int)sp.states.size();
    std::map<std::vector<uint8_t>, int, decltype(cmp)> index(cmp);
    for (int i = 0; i < sp.n; ++i) index[sp.states[i]] = i;
    std::set<std::tuple<int, int, int>> uniq;
    for (auto& e : raw) uniq.insert(std::make_tuple(index[e.src], index[e.dst], e.label));
    for (auto& e : uniq) { sp.edges.push_back(std::get<0>(e)); sp.edges.push_back(std::get<1>(e)); sp.edges.push_back(std::get<2>(e)); }
    for (int i = 0; i < sp.n; ++i) {
        bool left = false, right = false;
        for (uint8_t t : sp.states[i]) { left = left || lin_l(t) == 3; right = right || lin_r(t) == 3; }
        (left ? (right ? sp.E : sp.L) : (right ? sp.R : sp.B)).push_back(i);
    }
    if (kind != SP_SINGLE) {
        auto find = [&](int p1, int p2) {
            std::vector<uint8_t> s{lineage(p1, 1, 1), lineage(p2, 2, 2)};
            canon(s);
            return index.at(s);
        };
        sp.i12 = find(1, 2);
        if (kind == SP_MIG) { sp.i11 = find(1, 1); sp.i22 = find(2, 2); }
    }
    return sp;
}

static const HostSpace& host_space(int kind) {
    static HostSpace spaces[N_SPACES];
    static std::once_flag once[N_SPACES];
    std::call_once(once[kind], [kind] { spaces[kind] = build_space(kind); });
    return spaces[kind];
}

// map every state of `from` to the state of `to` obtained by relabelling (to_single) or keeping populations
static std::vector<int> projection(int from_kind, int to_kind) {
    const HostSpace &a = host_space(from_kind), &b = host_space(to_kind);
    std::vector<int> out(a.n, -1);
    for (int i = 0; i < a.n; ++i) {
        std::vector<uint8_t> s = a.states[i];
        if (to_kind == SP_SINGLE) for (auto& t : s) t = lineage(0, lin_l(t), lin_r(t));
        canon(s);
        for (int j = 0; j < b.n; ++j) if (b.states[j] == s) { out[i] = j; break; }
    }
    return out;
}

}  // namespace imc

// ------------------------------------------------------------------------------------------ model objects
struct imc_model {
    int kind = 0, K = 0, P = 0;
    int n_mig = 0, n_anc = 0, n_epochs = 0, est_split = 0, initial_state = 0, pre_space_n = 0;
    std::vector<int> interval_space, interval_epoch;
    std::vector<double> c_exp_a, c_psmc;
    std::vector<long long> p_off;
    long long p_stride = 0;
    // device side (lazy)
    bool uploaded = false;
    imc::ModelDev dev{};
    DeviceBuf d_static;                           // all static tables in one allocation
    DeviceBuf d_theta, d_params, d_scratch, d_status, d_pbuf, d_prebuf, d_pi, d_T, d_E, d_out;
    HandleSerial serial;
};

extern "C" int imc_statespace_describe(int space, int* n_states, int* n_edges, int* counts /*[4] B,L,R,E*/,
                                       int* special /*[3] i11,i12,i22*/, int32_t* edges /*[n_edges][3]*/,
                                       int32_t* classes /*[n_states] 0=B 1=L 2=R 3=E*/, uint8_t* lineages /*[n_states][4]*/) {
    if (space < 0 || space >= imc::N_SPACES) return fail(IMC_ERR_INVALID, "space must be 0 (isolation), 1 (single) or 2 (migration)");
    const imc::HostSpace& s = imc::host_space(space);
    if (n_states) *n_states = s.n;
    if (n_edges) *n_edges = (int)s.edges.size() / 3;
    if (counts) { counts[0] = (int)s.B.size(); counts[1] = (int)s.L.size(); counts[2] = (int)s.R.size(); counts[3] = (int)s.E.size(); }
    if (special) { special[0] = s.i11; special[1] = s.i12; special[2] = s.i22; }
    if (edges) for (size_t i = 0; i < s.edges.size(); ++i) edges[i] = s.edges[i];
    if (classes) {
        for (int i : s.B) classes[i] = 0;
        for (int i : s.L) classes[i] = 1;
        for (int i : s.R) classes[i] = 2;
        for (int i : s.E) classes[i] = 3;
    }
    if (lineages)
        for (int i = 0; i < s.n; ++i)
            for (int k = 0; k < 4; ++k) lineages[i * 4 + k] = k < (int)s.states[i].size() ? s.states[i][k] : 0xff;
    return IMC_OK;
}

// ---- break points (break_points.py:9-30, 60-78, 81-108).  The same expressions feed the models' constant tables below
// (c_exp_a[i] = bp_exp_unit, c_psmc[i] = bp_psmc_unit at the defaults) and model_params_kernel's per-point arithmetic
// (unit / coal_rate + offset; (i / n) * (end - start) + start).
static inline double bp_exp_unit(int i, int n) { return -std::log1p(-(double)i / n); }                 // scipy expon.ppf(i / n)
static inline double bp_psmc_unit(int i, int n, double t_max, double mu) {
    return i == 0 ? 0.0 : 0.1 * (std::exp((double)i / n * std::log(1 + 10 * t_max * mu)) - 1.0);
}

extern "C" int imc_model_create(int kind, const int32_t* iparams, int n_iparams, imc_model** out) {
    using namespace imc;
    if (!out || (n_iparams > 0 && !iparams)) return fail(IMC_ERR_INVALID, "NULL argument");
    imc_model* m = new (std::nothrow) imc_model;
    if (!m) return fail(IMC_ERR_NOMEM, "out of memory");
    m->kind = kind;
    auto bad = [&](const char* msg) { delete m; return fail(IMC_ERR_INVALID, "%s", msg); };
    std::vector<int> intervals;
    switch (kind) {
        case MODEL_ISOLATION:       // iparams = {no_hmm_states}
            if (n_iparams != 1 || iparams[0] < 2) return bad("isolation model: iparams = {no_hmm_states >= 2}");
            m->K = iparams[0]; m->P = 3; m->pre_space_n = 4;
            m->interval_space.assign(m->K, SP_SINGLE);
            m->interval_epoch.assign(m->K, 0);
            m->initial_state = host_space(SP_ISO).i12;
            for (int i = 0; i < m->K; ++i) m->c_exp_a.push_back(bp_exp_unit(i, m->K));
            break;
        case MODEL_IM:              // iparams = {no_mig_states, no_ancestral_states}
            if (n_iparams != 2 || iparams[0] < 2 || iparams[1] < 1)
                return bad("IM model: iparams = {no_mig_states >= 2, no_ancestral_states >= 1} "
                           "(the reference's joint[0,0] at transitions.py:222 breaks for a single migration state)");
            m->n_mig = iparams[0]; m->n_anc = iparams[1]; m->K = m->n_mig + m->n_anc; m->P = 5; m->pre_space_n = 4;
            for (int i = 0; i < m->K; ++i) m->interval_space.push_back(i < m->n_mig ? SP_MIG : SP_SINGLE);
            m->interval_epoch.assign(m->K, 0);
            m->initial_state = host_space(SP_ISO).i12;
            for (int i = 0; i < m->n_anc; ++i) m->c_exp_a.push_back(bp_exp_unit(i, m->n_anc));
            break;
        case MODEL_PSMC_ISO:        // iparams = {est_split, n_epochs, intervals...}
        case MODEL_VARMIG: {        // iparams = {initial_configuration, n_epochs, intervals...}
            if (n_iparams < 3 || iparams[1] < 1 || n_iparams != 2 + iparams[1]) return bad("variable-rate model: iparams = {flag, n_epochs, intervals[n_epochs]}");
            m->n_epochs = iparams[1];
            for (int e = 0; e < m->n_epochs; ++e) {
                if (iparams[2 + e] < 1) return bad("every epoch needs at least one interval");
                for (int k = 0; k < iparams[2 + e]; ++k) m->interval_epoch.push_back(e);
            }
            m->K = (int)m->interval_epoch.size();
            if (m->K < 2) return bad("at least two intervals are needed");
            if (kind == MODEL_PSMC_ISO) {
                m->est_split = iparams[0] ? 1 : 0;
                m->P = m->n_epochs + 1 + m->est_split; m->pre_space_n = 4;
                m->interval_space.assign(m->K, SP_SINGLE);
                m->initial_state = host_space(SP_ISO).i12;
            } else {
                if (iparams[0] < 0 || iparams[0] > 2) return bad("initial configuration must be 0 (11), 1 (12) or 2 (22)");
                m->P = 4 * m->n_epochs + 1; m->pre_space_n = 0;
                m->interval_space.assign(m->K, SP_MIG);
                const HostSpace& mg = host_space(SP_MIG);
                m->initial_state = iparams[0] == 0 ? mg.i11 : (iparams[0] == 1 ? mg.i12 : mg.i22);
            }
            // psmc_break_points(K, t_max=15, mu=1e-9) with offset 0 (break_points.py:81-108)
            for (int i = 0; i < m->K; ++i)
                m->c_psmc.push_back(bp_psmc_unit(i, m->K, 15, 1e-9));
            break;
        }
        case MODEL_IM_EPOCHS:       // iparams = {no_epochs, no_mig_states, no_ancestral_states}
            if (n_iparams != 3 || iparams[0] < 1 || iparams[1] < 1 || iparams[2] < 1 || iparams[0] * iparams[1] < 2)
                return bad("IM epochs model: iparams = {no_epochs, no_mig_states, no_ancestral_states}, at least two migration intervals");
            m->n_epochs = iparams[0]; m->n_mig = iparams[1]; m->n_anc = iparams[2];
            m->K = m->n_epochs * (m->n_mig + m->n_anc); m->P = 3 + (2 * m->n_epochs + 1) + m->n_epochs; m->pre_space_n = 4;
            for (int i = 0; i < m->K; ++i) m->interval_space.push_back(i < m->n_epochs * m->n_mig ? SP_MIG : SP_SINGLE);
            m->interval_epoch.assign(m->K, 0);
            m->initial_state = host_space(SP_ISO).i12;
            for (int i = 0; i < m->n_epochs * m->n_anc; ++i) m->c_exp_a.push_back(bp_exp_unit(i, m->n_epochs * m->n_anc));
            break;
        default:
            return bad("unknown model kind");
    }
    if (m->K > 128) return bad("more than 128 HMM states are not supported");
    long long off = 0;
    for (int i = 0; i < m->K; ++i) {
        m->p_off.push_back(off);
        if (i + 1 < m->K) off += (long long)host_space(m->interval_space[i]).n * host_space(m->interval_space[i + 1]).n;
    }
    m->p_stride = off;
    *out = m;
    return IMC_OK;
}


extern "C" int imc_break_points(int kind, int no_intervals, double a, double b, double c, double* out) {
    if (!out || no_intervals < 1) return fail(IMC_ERR_INVALID, "bad arguments");
    for (int i = 0; i < no_intervals; ++i) {
        switch (kind) {
            case 0: out[i] = bp_exp_unit(i, no_intervals) / a + b; break;                      // exp_break_points(n, coal_rate = a, offset = b)
            case 1: out[i] = ((double)i / no_intervals) * (b - a) + a; break;                  // uniform_break_points(n, start = a, end = b)
            case 2: out[i] = i == 0 ? c : c + bp_psmc_unit(i, no_intervals, a, b); break;      // psmc_break_points(n, t_max = a, mu = b, offset = c)
            default: return fail(IMC_ERR_INVALID, "break point kind must be 0 (exp), 1 (uniform) or 2 (psmc)");
        }
    }
    return IMC_OK;
}

extern "C" int imc_model_info(const imc_model* m, int* K, int* P) {
    if (!m) return fail(IMC_ERR_INVALID, "NULL model");
    if (K) *K = m->K;
    if (P) *P = m->P;
    return IMC_OK;
}

extern "C" int imc_model_destroy(imc_model* m) {
    if (!m) return IMC_OK;
    if (g_ctx.pid == getpid()) {      // buffers may exist before the static tables were uploaded
        if (m->serial.done) cudaEventDestroy(m->serial.done);
        for (DeviceBuf* b : {&m->d_static, &m->d_theta, &m->d_params, &m->d_scratch, &m->d_status, &m->d_pbuf,
                             &m->d_prebuf, &m->d_pi, &m->d_T, &m->d_E, &m->d_out}) b->release();
    }
    delete m;
    return IMC_OK;
}

static int model_upload(imc_model* m) {
    using namespace imc;
    if (m->uploaded) return IMC_OK;
    int rc = ensure_device();
    if (rc) return rc;
    // pack every static table into one buffer (8-byte aligned sections)
    std::vector<unsigned char> blob;
    auto put = [&](const void* p, size_t bytes) {
        const size_t at = (blob.size() + 7) & ~size_t(7);
        blob.resize(at + bytes);
        if (bytes) memcpy(blob.data() + at, p, bytes);
        return at;
    };
    struct Off { size_t edges, B, L, E; } so[N_SPACES];
    for (int s = 0; s < N_SPACES; ++s) {
        const HostSpace& hs = host_space(s);
        so[s].edges = put(hs.edges.data(), hs.edges.size() * sizeof(int));
        so[s].B = put(hs.B.data(), hs.B.size() * sizeof(int));
        so[s].L = put(hs.L.data(), hs.L.size() * sizeof(int));
        so[s].E = put(hs.E.data(), hs.E.size() * sizeof(int));
    }
    const std::vector<int> pis = projection(SP_ISO, SP_SINGLE), pim = projection(SP_ISO, SP_MIG), pms = projection(SP_MIG, SP_SINGLE);
    const size_t o_isp = put(m->interval_space.data(), m->interval_space.size() * sizeof(int));
    const size_t o_iep = put(m->interval_epoch.data(), m->interval_epoch.size() * sizeof(int));
    const size_t o_pis = put(pis.data(), pis.size() * sizeof(int));
    const size_t o_pim = put(pim.data(), pim.size() * sizeof(int));
    const size_t o_pms = put(pms.data(), pms.size() * sizeof(int));
    const size_t o_cexp = put(m->c_exp_a.data(), m->c_exp_a.size() * sizeof(double));
    const size_t o_psmc = put(m->c_psmc.data(), m->c_psmc.size() * sizeof(double));
    const size_t o_poff = put(m->p_off.data(), m->p_off.size() * sizeof(long long));
    if ((rc = m->d_static.reserve(blob.size()))) return rc;
    CUDA_TRY(cudaMemcpy(m->d_static.p, blob.data(), blob.size(), cudaMemcpyHostToDevice));
    const unsigned char* base = (const unsigned char*)m->d_static.p;
    ModelDev& d = m->dev;
    d.kind = m->kind; d.K = m->K; d.P = m->P; d.n_mig = m->n_mig; d.n_anc = m->n_anc; d.n_epochs = m->n_epochs;
    d.est_split = m->est_split; d.initial_state = m->initial_state; d.pre_space_n = m->pre_space_n;
    for (int s = 0; s < N_SPACES; ++s) {
        const HostSpace& hs = host_space(s);
        d.space[s].n = hs.n; d.space[s].n_edges = (int)hs.edges.size() / 3;
        d.space[s].edges = (const int*)(base + so[s].edges);
        d.space[s].nB = (int)hs.B.size(); d.space[s].nL = (int)hs.L.size(); d.space[s].nE = (int)hs.E.size();
        d.space[s].B = (const int*)(base + so[s].B); d.space[s].L = (const int*)(base + so[s].L); d.space[s].E = (const int*)(base + so[s].E);
    }
    d.interval_space = (const int*)(base + o_isp); d.interval_epoch = (const int*)(base + o_iep);
    d.proj_iso_single = (const int*)(base + o_pis); d.proj_iso_mig = (const int*)(base + o_pim); d.proj_mig_single = (const int*)(base + o_pms);
    d.c_exp_a = (const double*)(base + o_cexp); d.c_psmc = (const double*)(base + o_psmc);
    d.p_off = (const long long*)(base + o_poff); d.p_stride = m->p_stride;
    m->uploaded = true;
    return IMC_OK;
}

// theta (device) -> pi, T, E, status (device).  Enqueued on st; no synchronisation.
static int model_build_dev(imc_model* m, int N, const double* d_theta, double* d_pi, double* d_T, double* d_E,
                           int* d_status, cudaStream_t st) {
    using namespace imc;
    NvtxRange nvtx_build("imc: model build (theta -> pi, T, E)");
    int rc = model_upload(m);
    if (rc) return rc;
    const int K = m->K;
    if (N > 32768) {        // model_expm_kernel puts the parameter point on gridDim.y: large batches run as slices
        for (int n0 = 0; n0 < N; n0 += 32768) {
            const int nn = std::min(32768, N - n0);
            if ((rc = model_build_dev(m, nn, d_theta + (size_t)n0 * m->P, d_pi + (size_t)n0 * K, d_T + (size_t)n0 * K * K,
                                      d_E + (size_t)n0 * K * 3, d_status + n0, st))) return rc;
        }
        return IMC_OK;
    }
    if ((rc = m->d_params.reserve(sizeof(double) * (size_t)N * PointParams::size(K)))) return rc;
    if ((rc = m->d_scratch.reserve(sizeof(double) * (size_t)N * 3 * K))) return rc;
    if ((rc = m->d_pbuf.reserve(sizeof(double) * (size_t)N * m->p_stride))) return rc;
    if ((rc = m->d_prebuf.reserve(sizeof(double) * (size_t)N * 16))) return rc;
    model_params_kernel<<<(N + 127) / 128, 128, 0, st>>>(m->dev, N, d_theta, (double*)m->d_params.p, d_E, d_status,
                                                         (double*)m->d_scratch.p);
    CUDA_TRY(cudaGetLastError());
    int nmax = 4;
    for (int s : m->interval_space) nmax = std::max(nmax, host_space(s).n);
    const size_t smem_expm = sizeof(double) * 3 * (size_t)nmax * nmax;
    if (nmax <= 16) {       // isolation-type models: the lean instantiation (no tensor-path code, 8 CTAs per SM)
        // the CTA is a chain of ~16 tiny products with barriers in between: a batch that fills the machine gains from resident
        // CTAs (64 threads per 15 x 15 exponential: 98 vs 113 us for 256 points), a single point from threads per product (18 vs 42 us)
        model_expm_kernel<true><<<dim3(K, N), (long long)K * N >= 1024 ? 64 : 256, smem_expm, st>>>(m->dev, (const double*)m->d_params.p, d_status,
                                                                    (double*)m->d_pbuf.p, (double*)m->d_prebuf.p);
    } else {
        static size_t expm_attr = 0;
        if (smem_expm > expm_attr) {
            CUDA_TRY(cudaFuncSetAttribute(model_expm_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_expm));
            expm_attr = smem_expm;
        }
        model_expm_kernel<false><<<dim3(K, N), 256, smem_expm, st>>>(m->dev, (const double*)m->d_params.p, d_status,
                                                                     (double*)m->d_pbuf.p, (double*)m->d_prebuf.p);
    }
    CUDA_TRY(cudaGetLastError());
    size_t smem_chain = sizeof(double) * (2 * MAX_STATES + (size_t)K * K + 2 * (size_t)K * MAX_L + 128);
    const int stage = smem_chain + sizeof(double) * (size_t)m->p_stride <= 96 * 1024 ? 1 : 0;     // (two CTAs per SM still fit)
    if (stage) smem_chain += sizeof(double) * (size_t)m->p_stride;
    static size_t chain_attr = 0;
    if (smem_chain > chain_attr && smem_chain > 48 * 1024) {
        CUDA_TRY(cudaFuncSetAttribute(model_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_chain));
        chain_attr = smem_chain;
    }
    model_chain_kernel<<<N, 128, smem_chain, st>>>(m->dev, (const double*)m->d_params.p, (const double*)m->d_pbuf.p,
                                                   (const double*)m->d_prebuf.p, d_status, d_pi, d_T, stage);
    CUDA_TRY(cudaGetLastError());
    g_launches += 3;
    return IMC_OK;
}

extern "C" int imc_model_build_batch_dev(imc_model* m, int N, const double* d_theta, double* d_pi, double* d_T,
                                         double* d_E, int32_t* d_status, void* stream) {
    if (!m) return fail(IMC_ERR_INVALID, "NULL model");
    if (N <= 0) return N == 0 ? IMC_OK : fail(IMC_ERR_INVALID, "N < 0");
    if (!d_theta || !d_pi || !d_T || !d_E || !d_status) return fail(IMC_ERR_INVALID, "NULL device pointer");
    CallGuard guard(m->serial, (cudaStream_t)stream);
    return model_build_dev(m, N, d_theta, d_pi, d_T, d_E, d_status, (cudaStream_t)stream);
}

static int model_stage_theta(imc_model* m, int N, const double* theta, cudaStream_t st) {
    int rc;
    const size_t K = m->K;
    if ((rc = ensure_device())) return rc;
    if ((rc = m->d_theta.reserve(sizeof(double) * (size_t)N * m->P))) return rc;
    if ((rc = m->d_status.reserve(sizeof(int) * (size_t)N))) return rc;
    if ((rc = m->d_pi.reserve(sizeof(double) * N * K))) return rc;
    if ((rc = m->d_T.reserve(sizeof(double) * N * K * K))) return rc;
    if ((rc = m->d_E.reserve(sizeof(double) * N * K * 3))) return rc;
    if ((rc = m->d_out.reserve(sizeof(double) * (size_t)N))) return rc;
    CUDA_TRY(cudaMemcpyAsync(m->d_theta.p, theta, sizeof(double) * (size_t)N * m->P, cudaMemcpyHostToDevice, st));
    return IMC_OK;
}

extern "C" int imc_model_build_batch(imc_model* m, int N, const double* theta, double* pi, double* T, double* E,
                                     int32_t* status) {
    if (!m) return fail(IMC_ERR_INVALID, "NULL model");
    if (N <= 0) return N == 0 ? IMC_OK : fail(IMC_ERR_INVALID, "N < 0");
    if (!theta || !pi || !T || !E) return fail(IMC_ERR_INVALID, "NULL host pointer");
    int rc = ensure_device();
    if (rc) return rc;
    cudaStream_t st = g_ctx.stream;
    CallGuard guard(m->serial, st);
    if ((rc = model_stage_theta(m, N, theta, st))) return rc;
    const size_t K = m->K;
    if ((rc = model_build_dev(m, N, (const double*)m->d_theta.p, (double*)m->d_pi.p, (double*)m->d_T.p, (double*)m->d_E.p,
                              (int*)m->d_status.p, st))) return rc;
    CUDA_TRY(cudaMemcpyAsync(pi, m->d_pi.p, sizeof(double) * N * K, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(T, m->d_T.p, sizeof(double) * N * K * K, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(E, m->d_E.p, sizeof(double) * N * K * 3, cudaMemcpyDeviceToHost, st));
    if (status) CUDA_TRY(cudaMemcpyAsync(status, m->d_status.p, sizeof(int) * (size_t)N, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return IMC_OK;
}

// the break points model_params_kernel computed for each parameter point (what <model>.py build_ctmc_system hands to the
// CTMC system): out[N][K]; rows of invalid points (status 1) are NaN
extern "C" int imc_model_break_points(imc_model* m, int N, const double* theta, double* out) {
    if (!m) return fail(IMC_ERR_INVALID, "NULL model");
    if (N <= 0) return N == 0 ? IMC_OK : fail(IMC_ERR_INVALID, "N < 0");
    if (!theta || !out) return fail(IMC_ERR_INVALID, "NULL host pointer");
    if (N > 32768) return fail(IMC_ERR_UNSUPPORTED, "at most 32768 parameter points per imc_model_break_points call");
    int rc = ensure_device();
    if (rc) return rc;
    cudaStream_t st = g_ctx.stream;
    CallGuard guard(m->serial, st);
    if ((rc = model_stage_theta(m, N, theta, st))) return rc;
    const size_t K = m->K;
    if ((rc = model_build_dev(m, N, (const double*)m->d_theta.p, (double*)m->d_pi.p, (double*)m->d_T.p, (double*)m->d_E.p,
                              (int*)m->d_status.p, st))) return rc;
    std::vector<int> status(N);
    CUDA_TRY(cudaMemcpy2DAsync(out, sizeof(double) * K, m->d_scratch.p, sizeof(double) * 3 * K, sizeof(double) * K, (size_t)N,
                               cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(status.data(), m->d_status.p, sizeof(int) * (size_t)N, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    for (int n = 0; n < N; ++n)
        if (status[n] == 1) for (size_t i = 0; i < K; ++i) out[(size_t)n * K + i] = std::nan("");
    return IMC_OK;
}

// fused: theta -> (pi,T,E) -> forward -> logL, all on the device
static int forward_dev(imc_seqset* set, int N, int K, int S, const double* d_pi, const double* d_T, const double* d_E,
                       double* d_out, cudaStream_t st);

extern "C" int imc_loglik_batch_dev(imc_model* m, imc_seqset* set, int N, const double* d_theta, double* d_out,
                                    int32_t* d_status, void* stream) {
    if (!m || !set) return fail(IMC_ERR_INVALID, "NULL handle");
    if (N <= 0) return N == 0 ? IMC_OK : fail(IMC_ERR_INVALID, "N < 0");
    if (!d_theta || !d_out) return fail(IMC_ERR_INVALID, "NULL device pointer");
    int rc = ensure_device();
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    CallGuard guard(m->serial, st);          // the model's pi / T / E staging is in use until the forward has read it
    const size_t K = m->K;
    if ((rc = m->d_pi.reserve(sizeof(double) * N * K))) return rc;
    if ((rc = m->d_T.reserve(sizeof(double) * N * K * K))) return rc;
    if ((rc = m->d_E.reserve(sizeof(double) * N * K * 3))) return rc;
    int* status = d_status;
    if (!status) {
        if ((rc = m->d_status.reserve(sizeof(int) * (size_t)N))) return rc;
        status = (int*)m->d_status.p;
    }
    if ((rc = model_build_dev(m, N, d_theta, (double*)m->d_pi.p, (double*)m->d_T.p, (double*)m->d_E.p, status, st))) return rc;
    if ((rc = forward_dev(set, N, m->K, 3, (const double*)m->d_pi.p, (const double*)m->d_T.p, (const double*)m->d_E.p, d_out, st))) return rc;
    imc::model_status_fixup_kernel<<<(N + 127) / 128, 128, 0, st>>>(status, N, d_out);
    CUDA_TRY(cudaGetLastError());
    g_launches += 1;
    return IMC_OK;
}

extern "C" int imc_loglik_batch(imc_model* m, imc_seqset* set, int N, const double* theta, double* out, int32_t* status) {
    if (!m || !set) return fail(IMC_ERR_INVALID, "NULL handle");
    if (N <= 0) return N == 0 ? IMC_OK : fail(IMC_ERR_INVALID, "N < 0");
    if (!theta || !out) return fail(IMC_ERR_INVALID, "NULL host pointer");
    int rc = ensure_device();
    if (rc) return rc;
    cudaStream_t st = g_ctx.stream;
    CallGuard guard(m->serial, st);
    if ((rc = model_stage_theta(m, N, theta, st))) return rc;
    if ((rc = imc_loglik_batch_dev(m, set, N, (const double*)m->d_theta.p, (double*)m->d_out.p, (int32_t*)m->d_status.p, st))) return rc;
    CUDA_TRY(cudaMemcpyAsync(out, m->d_out.p, sizeof(double) * (size_t)N, cudaMemcpyDeviceToHost, st));
    if (status) CUDA_TRY(cudaMemcpyAsync(status, m->d_status.p, sizeof(int) * (size_t)N, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return g_comm ? comm_check() : IMC_OK;
}
