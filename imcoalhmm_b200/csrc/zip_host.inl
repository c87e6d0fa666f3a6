// Host side of the compressed (zip) forward path (included by imc_lib.cu): launch plans, device copies of the token
// streams, segment descriptors, introspection entry points and the launcher of zip_forward_kernel.

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

struct ZipPlan { int lanes, threads, ctas_per_sm, M; size_t smem; bool spec, mma, run2; };
#ifndef IMC_ZIP8_THREADS
#define IMC_ZIP8_THREADS 512     // threads of the single-CTA-per-SM shape with 8 lanes per chain, K <= 24 (experiment builds: 640, 768)
#endif
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

// MMA form (ZipCfgM): spectral form, tiles >= 8 states
static bool zip_mma_tile(int K) { return K >= 8; }
static thread_local int g_plan_chunks = 0;     // chunks of the set being planned for (0 = unknown), set by zip_pass
static thread_local int g_mma_shape_hint = 0;           // launch shape zip_plan uses for the MMA form when the option says auto (set while planning the aligned form)
static unsigned long long* g_mma_passes = nullptr;      // device counter (ZipArgs::mma_passes)

template <int K, bool SPEC>
static ZipPlan zip_plan_k(int S, int avail_ids, int want_ctas, int want_lanes, bool mma, bool run2) {
    ZipPlan p;
    p.mma = false;
    p.run2 = false;
    if constexpr (SPEC && K >= 8) {
        if (mma) {
            // shapes: 1 = one CTA of 512 threads per SM (256 for K >= 32), 2 = two CTAs of 256, 3 = four CTAs of 256 threads
            // with 64 registers (K <= 12), 4 = four CTAs of 128 threads (K <= 24).  More CTAs per SM = more parameter
            // points in flight per SM (a point offers chunks / 8 warp-loads at a time) and smaller dictionaries, which cost
            // little in the spectral form (K=10: 62 instead of 162 entries = 5 % more tokens).
            using C = ZipCfgM<K>;
            int shape = (int)g_ctx.opt_zip_mma_shape;
            // a point offers ceil(chunks / 8) independent warp-loads at a time: CTAs with more warps than that would idle
            // lock step: one CTA (c2: 3.52 vs 3.70 ms, K=20: 12.4 vs 13.2 ms with two CTAs of 256).  Aligned form (g_mma_shape_hint): two
            // CTAs -- its dictionary is small, two points in flight fill the warps a single point leaves idle and hide each other's
            // dictionary builds (c2: 3.15 vs 3.48 ms; 128 / 192 chunks: 3.64 / 5.20 vs 3.90 / 5.57; K=20, 100 chunks: 15.8 vs 17.4 ms)
            if (shape == 0) shape = g_mma_shape_hint ? g_mma_shape_hint : 1;
            if (K > 24) shape = 1;
            if (K > 12 && shape == 3) shape = 2;
            p.mma = true;
            p.lanes = 4;                     // four lanes per chain, eight chains per warp
            p.ctas_per_sm = shape == 1 ? 1 : (shape == 2 ? 2 : 4);
            p.threads = shape == 1 ? (K <= 24 ? 512 : 256) : (shape == 4 ? 128 : 256);
            const size_t budget = (ZIP_SMEM_SM - 1024 * (p.ctas_per_sm - 1)) / p.ctas_per_sm;
            p.run2 = run2;
            p.M = std::min(avail_ids, ZipSmem<C, true>::max_entries(budget, S, p.threads, run2));
            if (g_ctx.opt_zip_max_entries > 0) p.M = std::min<int>(p.M, (int)g_ctx.opt_zip_max_entries);
            p.M = std::max(p.M, S + (run2 ? 2 : 0));
            p.spec = true;
            p.smem = ZipSmem<C, true>::bytes(p.M, S, p.threads, run2);
            return p;
        }
    }
    p.lanes = want_lanes ? want_lanes : zip_default_lanes(K);
    if (K < 8) p.lanes = 8;
    if (p.lanes == 32 && K < 10) p.lanes = 8;
    int m1, m2, t1;
    if (p.lanes == 32) {       // one chain per warp, one CTA per SM (latency mode)
        using C = ZipCfg32<K>;
        t1 = K <= 24 ? 512 : 256;
        m2 = 0;
        m1 = ZipSmem<C, SPEC>::max_entries(ZIP_SMEM_SM, S, t1);
        want_ctas = 1;
    } else if (p.lanes == 8) {
        using C = ZipCfg8<K>;
        t1 = K <= 24 ? IMC_ZIP8_THREADS : 256;
        m2 = K <= 24 ? ZipSmem<C, SPEC>::max_entries((ZIP_SMEM_SM - 1024) / 2, S, 256) : 0;
        m1 = ZipSmem<C, SPEC>::max_entries(ZIP_SMEM_SM, S, t1);
    } else {
        using C = ZipCfg4<K>;
        t1 = 256;
        m2 = K <= 24 ? ZipSmem<C, SPEC>::max_entries((ZIP_SMEM_SM - 1024) / 2, S, 256) : 0;
        m1 = ZipSmem<C, SPEC>::max_entries(ZIP_SMEM_SM, S, t1);
    }
    int ctas = want_ctas;
    if (ctas == 2 && K > 24) ctas = 1;
    if (ctas == 0) ctas = (K <= 24 && m2 >= avail_ids) ? 2 : 1;   // the bigger dictionary wins unless everything fits in half
    p.ctas_per_sm = ctas;
    p.threads = ctas == 2 ? 256 : t1;
    p.M = std::min(avail_ids, ctas == 2 ? m2 : m1);
    if (g_ctx.opt_zip_max_entries > 0) p.M = std::min<int>(p.M, (int)g_ctx.opt_zip_max_entries);
    p.M = std::max(p.M, S);
    p.spec = SPEC;
    p.smem = p.lanes == 8 ? ZipSmem<ZipCfg8<K>, SPEC>::bytes(p.M, S, p.threads)
           : (p.lanes == 4 ? ZipSmem<ZipCfg4<K>, SPEC>::bytes(p.M, S, p.threads) : ZipSmem<ZipCfg32<K>, SPEC>::bytes(p.M, S, p.threads));
    return p;
}

// spec: plan for the spectral form (run tokens, power table in shared memory) over avail_ids run-dictionary ids
static int zip_plan(int K, int S, int avail_ids, ZipPlan* out, int lanes_override = 0, bool spec = false, bool mma = false, bool run2 = false) {
    const int want = (int)g_ctx.opt_zip_ctas_per_sm, lanes = lanes_override ? lanes_override : (int)g_ctx.opt_zip_lanes;
    switch (zip_tile(K)) {
#define X(k) case k: *out = spec ? zip_plan_k<k, true>(S, avail_ids, want, lanes, mma, run2 && mma) : zip_plan_k<k, false>(S, avail_ids, want, lanes, false, false); break;
        ZIP_K_LIST(X)
#undef X
        default: return fail(IMC_ERR_UNSUPPORTED, "zip kernel is not instantiated for K = %d", K);
    }
    if (out->smem > ZIP_SMEM_SM) return fail(IMC_ERR_UNSUPPORTED, "zip kernel: %d symbols x K = %d do not fit in shared memory", S, K);
    return IMC_OK;
}

// chunk index (as given at creation) of stream k
static int set_chunk_of_stream(const imc_seqset* set, int k) {
    for (int c = 0; c < set->n_chunks; ++c) if (set->stream_of_chunk[c] == k) return c;
    return 0;
}

// Two-run form, prepared on first use: pick the second run symbol (the one with the most sites in runs of >= RUN2_MIN among
// the symbols other than the run symbol), learn its pair dictionary over nsym + 2 base ids on a sample, encode every chunk.
static int seqset_run2_prepare(imc_seqset* set) {
    if (set->run2_state != 0) return IMC_OK;
    set->run2_state = -1;
    const int ns = (int)set->streams.size(), nsym = set->nsym;
    if (nsym < 2 || nsym + 2 > 200 || ns == 0 || set->parts_total > 0) return IMC_OK;
    try {
        // every step expands a stream back to its symbols and lets go of them again: never more than one chunk per core in memory
        std::vector<std::vector<long long>> inruns_k(ns, std::vector<long long>(nsym, 0));
        if (!parallel_for(ns, [&](int k) {
                std::vector<uint8_t> sy;
                zip_expand(set->merges, set->tok_full[k], nsym, sy);
                for (size_t t = 0; t < sy.size();) {
                    size_t e = t;
                    while (e < sy.size() && sy[e] == sy[t]) ++e;
                    if (sy[t] != set->fold_sym && e - t >= (size_t)RUN2_MIN) inruns_k[k][sy[t]] += (long long)(e - t);
                    t = e;
                }
            })) throw std::bad_alloc();
        std::vector<long long> inruns(nsym, 0);
        for (int k = 0; k < ns; ++k) for (int v = 0; v < nsym; ++v) inruns[v] += inruns_k[k][v];
        const int r2 = (int)(std::max_element(inruns.begin(), inruns.end()) - inruns.begin());
        if (inruns[r2] == 0) return IMC_OK;
        set->run_sym2 = r2;
        set->run2_tok_full.resize(ns);
        set->run2_first_run.assign(ns, 0);
        set->run2_sites.assign(ns, 0);
        {
            std::vector<std::vector<uint32_t>> sample;
            long long budget = 16LL << 20;
            const int stride = std::max(1, ns / 32);
            for (int k = 0; k < ns && budget > 0; k += stride) {
                int fr; long long r2s;
                std::vector<uint8_t> sy;
                zip_expand(set->merges, set->tok_full[k], nsym, sy);
                sample.emplace_back();
                run2_tokenize(sy.data(), sy.size(), nsym, set->fold_sym, r2, &fr, sample.back(), &r2s);
                budget -= (long long)sy.size();
            }
            set->run2_merges = run_learn(sample, nsym + 2, 256, 16);
        }
        if (!parallel_for(ns, [&](int k) {
                std::vector<uint8_t> sy;
                zip_expand(set->merges, set->tok_full[k], nsym, sy);
                run2_encode(set->run2_merges, sy.data(), sy.size(), nsym, set->fold_sym, r2, &set->run2_first_run[k],
                            set->run2_tok_full[k], &set->run2_sites[k]);
            })) throw std::bad_alloc();
    } catch (const std::bad_alloc&) { return fail(IMC_ERR_NOMEM, "out of host memory while preparing the two-run encoding"); }
    set->run2_state = 1;
    return IMC_OK;
}

// token streams over the first M dictionary ids, level-ordered, on the device (cached per M and form)
static int zip_device_build(imc_seqset* set, int M, bool spec, bool run2, bool sched, ZipDevice** out);
static int zip_device(imc_seqset* set, int M, ZipDevice** out, bool spec = false, bool run2 = false, bool sched = false) {
    try { return zip_device_build(set, M, spec, run2, sched, out); }
    catch (const std::bad_alloc&) { return fail(IMC_ERR_NOMEM, "out of host memory while deriving the token streams"); }
}

// Aligned form of the MMA shape (sched): the chains of a warp do not march in lock step through their own token streams but
// follow ONE common schedule per warp-load (quad) -- a supersequence of their entry-id sequences, built here -- in which every
// warp-step applies a single entry: chains whose next token is that entry take it, the others hold a no-op word.  One MMA pass per
// warp-step, where lock step spends 1 + (distinct cold entries among the chains).  Schedule: hot steps until `stall` chains wait at
// a cold entry (or no chain wants the hot one), then a cold slot: one step per waiting cold entry, and once more for the chains that
// moved on to another cold entry (in the two-run form a chain served by "into the second basis" wants "back" next); further rounds
// only while they serve two chains per step.  streams[i] = words of chain i, all of one length (a multiple of 8).
static void zip_align_quad(const std::vector<const std::vector<uint32_t>*>& tok, int hot, int stall,
                           std::vector<std::vector<uint32_t>>* streams, long long* steps_out) {
    const int n = (int)tok.size();
    std::vector<size_t> ptr(n, 0);
    long long steps = 0;
    auto next_id = [&](int i) { return ptr[i] < tok[i]->size() ? (int)((*tok[i])[ptr[i]] & 0xffu) : -1; };
    auto emit = [&](int id) {
        for (int i = 0; i < n; ++i) {
            uint32_t w = RUN_NOP_BIT | (uint32_t)id;          // (sitting this step out, but every lane knows the step's entry)
            if (next_id(i) == id) w = (*tok[i])[ptr[i]++];
            if (streams) (*streams)[i].push_back(w);
        }
        ++steps;
    };
    for (;;) {
        bool left = false;
        for (int i = 0; i < n; ++i) left = left || next_id(i) >= 0;
        if (!left) break;
        for (;;) {
            int nhot = 0, nstalled = 0;
            for (int i = 0; i < n; ++i) { const int id = next_id(i); nhot += id == hot; nstalled += id >= 0 && id != hot; }
            if (nhot == 0 || nstalled >= stall) break;
            emit(hot);
        }
        // cold slot: one step per waiting cold entry; again for the chains it has moved to another cold entry (in the two-run form
        // "into the second basis" is followed by "back"), as long as that serves at least two chains per step
        for (int round = 0; round < 8; ++round) {
            int want[256] = {0}, waiting = 0, kinds = 0;
            for (int i = 0; i < n; ++i) { const int id = next_id(i); if (id >= 0 && id != hot) { kinds += want[id]++ == 0; ++waiting; } }
            if (waiting == 0 || (round >= 2 && waiting < 2 * kinds)) break;
            for (int id = 0; id < 256; ++id) if (want[id]) emit(id);
        }
    }
    if (streams)
        for (int i = 0; i < n; ++i) while ((*streams)[i].size() % 8) (*streams)[i].push_back(RUN_NOP_BIT | (uint32_t)hot);   // (an idle hot step)
    if (steps_out) *steps_out = (steps + 7) / 8 * 8;
}

// The two-run token streams of the set over the first M dictionary ids (ids level-ordered like the device dictionary), the
// order of the streams by length and their histogram figures: what the aligned form is built from (host only, no CUDA).
struct ZipHostStreams {
    std::vector<std::vector<uint32_t>> rtok;
    std::vector<int> order;
    int hot_id = 0;
    double hot_share = 0.0, est_passes = 1.0;
    long long tokens = 0;
};
static void zip_host_streams(const ZipMerges& mg, const std::vector<std::vector<uint32_t>>& rfull, int M, const ZipLevels& zl, ZipHostStreams* h) {
    const int ns = (int)rfull.size();
    h->rtok.assign(ns, {});
    if (!parallel_for(ns, [&](int k) {
            run_expand(mg, rfull[k], M, h->rtok[k]);
            for (auto& t : h->rtok[k]) t = (t & ~0xffu) | zl.perm[t & 0xffu];
        })) throw std::bad_alloc();
    h->order.resize(ns);
    std::iota(h->order.begin(), h->order.end(), 0);
    std::stable_sort(h->order.begin(), h->order.end(), [&](int x, int y) { return h->rtok[x].size() > h->rtok[y].size(); });
    std::vector<long long> hist(256, 0);
    h->tokens = 0;
    for (int k = 0; k < ns; ++k) { for (uint32_t t : h->rtok[k]) hist[t & 0xffu]++; h->tokens += (long long)h->rtok[k].size(); }
    h->hot_id = (int)(std::max_element(hist.begin(), hist.end()) - hist.begin());
    h->hot_share = h->tokens > 0 ? (double)hist[h->hot_id] / (double)h->tokens : 0.0;
    h->est_passes = 1.0;
    for (int id = 0; id < 256; ++id)
        if (id != h->hot_id && hist[id] > 0) h->est_passes += 1.0 - std::pow(1.0 - (double)hist[id] / (double)h->tokens, 8.0);
}
static std::vector<const std::vector<uint32_t>*> zip_quad_tokens(const ZipHostStreams& h, int q) {
    std::vector<const std::vector<uint32_t>*> t;
    const int ns = (int)h.order.size();
    for (int i = q * 8; i < std::min(ns, q * 8 + 8); ++i) t.push_back(&h.rtok[h.order[i]]);
    return t;
}
// the stall threshold that gives the fewest steps on a sample of quads
static int zip_align_pick_stall(const ZipHostStreams& h) {
    const int nq = ((int)h.order.size() + 7) / 8;
    int stall = 3;
    long long best = -1;
    for (int g = 2; g <= 6; ++g) {
        long long tot = 0;
        for (int q = 0; q < nq; q += std::max(1, nq / 4)) { long long st; zip_align_quad(zip_quad_tokens(h, q), h.hot_id, g, nullptr, &st); tot += st; }
        if (best < 0 || tot < best) { best = tot; stall = g; }
    }
    return stall;
}

static int zip_device_build(imc_seqset* set, int M, bool spec, bool run2, bool sched, ZipDevice** out) {
    for (ZipDevice* z : set->zip_dev) if (z->M == M && z->spec == spec && z->run2 == run2 && z->sched == sched) { *out = z; return IMC_OK; }
    if (sched && !spec) return fail(IMC_ERR_INVALID, "the aligned form needs run tokens");
    const int ns = (int)set->streams.size();
    const ZipMerges& mg = run2 ? set->run2_merges : (spec ? set->run_merges : set->merges);
    const std::vector<std::vector<uint32_t>>& rfull = run2 ? set->run2_tok_full : set->run_tok_full;
    ZipLevels zl = zip_levels(mg, M);
    const size_t tsz = spec ? 4 : 1, align = spec ? 32 : 16;      // bytes per token; streams padded to whole blocks + one block of slack
    std::vector<std::vector<uint8_t>> tok(spec ? 0 : ns);
    std::vector<std::vector<uint32_t>> rtok(spec ? ns : 0);
    if (!parallel_for(ns, [&](int k) {
            if (spec) {
                run_expand(mg, rfull[k], M, rtok[k]);
                for (auto& t : rtok[k]) t = (t & ~0xffu) | zl.perm[t & 0xffu];
            } else {
                zip_expand(mg, set->tok_full[k], M, tok[k]);
                for (auto& t : tok[k]) t = zl.perm[t];
            }
        })) return fail(IMC_ERR_NOMEM, "out of host memory while deriving the token streams");
    auto ntok = [&](int k) { return spec ? rtok[k].size() : tok[k].size(); };
    std::vector<int> order(ns);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return ntok(x) > ntok(y); });
    int hot_id = 0;
    double hot_share = 0.0, est_passes = 1.0;
    long long real_tokens = 0;
    if (spec) {        // the most frequent entry of the streams: the MMA form keeps its matrix in registers
        std::vector<long long> hist(256, 0);
        for (int k = 0; k < ns; ++k) { for (uint32_t t : rtok[k]) hist[t & 0xffu]++; real_tokens += (long long)rtok[k].size(); }
        hot_id = (int)(std::max_element(hist.begin(), hist.end()) - hist.begin());
        hot_share = real_tokens > 0 ? (double)hist[hot_id] / (double)real_tokens : 0.0;
        for (int id = 0; id < 256; ++id)
            if (id != hot_id && hist[id] > 0) est_passes += 1.0 - std::pow(1.0 - (double)hist[id] / (double)real_tokens, 8.0);
    }
    // MMA passes the streams cost, summed over the warp-loads of 8 chains (sorted order): lock step = longest stream of the quad x
    // expected passes per warp-step; aligned = the steps of the quad's schedule
    long long pass_cost = 0;
    for (int i = 0; i < ns; i += 8) pass_cost += (long long)std::llround((double)ntok(order[i]) * est_passes);
    if (sched) {
        const int nq = (ns + 7) / 8;
        auto quad_tokens = [&](int q) {
            std::vector<const std::vector<uint32_t>*> t;
            for (int i = q * 8; i < std::min(ns, q * 8 + 8); ++i) t.push_back(&rtok[order[i]]);
            return t;
        };
        int stall = 3;          // the threshold that gives the fewest steps on a sample of quads
        {
            long long best = -1;
            for (int g = 2; g <= 6; ++g) {
                long long tot = 0;
                for (int q = 0; q < nq; q += std::max(1, nq / 4)) { long long st; zip_align_quad(quad_tokens(q), hot_id, g, nullptr, &st); tot += st; }
                if (best < 0 || tot < best) { best = tot; stall = g; }
            }
        }
        std::vector<std::vector<std::vector<uint32_t>>> aligned(nq);
        if (!parallel_for(nq, [&](int q) {
                const auto t = quad_tokens(q);
                aligned[q].resize(t.size());
                zip_align_quad(t, hot_id, stall, &aligned[q], nullptr);
            })) return fail(IMC_ERR_NOMEM, "out of host memory while aligning the token streams");
        pass_cost = 0;
        for (int q = 0; q < nq; ++q) {
            pass_cost += aligned[q].empty() ? 0 : (long long)aligned[q][0].size();
            for (size_t j = 0; j < aligned[q].size(); ++j) rtok[order[q * 8 + j]].swap(aligned[q][j]);      // (run sites below skip the no-op words)
        }
    }
    std::vector<ZipChunk> chunks(ns);
    long long off = 0;
    for (int i = 0; i < ns; ++i) {
        const int k = order[i];
        if (ntok(k) > 0x7fffffffULL) return fail(IMC_ERR_UNSUPPORTED, "a chunk has more than 2^31-1 tokens");
        chunks[i].tok_off = off;
        chunks[i].ntok = (int)ntok(k);
        chunks[i].first_sym = set->first_sym[k];
        chunks[i].out_index = k;
        chunks[i].first_run = spec ? (run2 ? set->run2_first_run[k] : set->first_run[k]) : 0;
        long long rs = chunks[i].first_run;
        if (spec) for (uint32_t t : rtok[k]) if (!(t & (RUN_TABLE2_BIT | RUN_NOP_BIT))) rs += (t >> 8) & RUN_MAX;
        chunks[i].run_sites = (int)rs;
        chunks[i].run2_sites = run2 ? (int)set->run2_sites[k] : 0;
        chunks[i].pad = 0;
        chunks[i].continues = set->parts_total > 0 ? set->part_first + set_chunk_of_stream(set, k) : 0;     // global part index (0: has its own start)
        off += (long long)((ntok(k) * tsz + align - 1) / align * align + align);
    }
    std::vector<uint8_t> flat((size_t)off, 0);
    if (sched) { uint32_t* w = reinterpret_cast<uint32_t*>(flat.data()); for (size_t x = 0; x < flat.size() / 4; ++x) w[x] = RUN_NOP_BIT | (uint32_t)hot_id; }
    long long total = 0;
    for (int i = 0; i < ns; ++i) {
        const int k = order[i];
        if (ntok(k)) memcpy(flat.data() + chunks[i].tok_off, spec ? (const void*)rtok[k].data() : (const void*)tok[k].data(), ntok(k) * tsz);
        total += (long long)ntok(k);
    }
    ZipDevice* z = new (std::nothrow) ZipDevice;
    if (!z) return fail(IMC_ERR_NOMEM, "out of memory");
    z->M = M;
    z->spec = spec;
    z->hot_id = hot_id;
    z->hot_share = hot_share;
    z->est_passes = sched ? 1.0 : est_passes;
    z->pass_cost = pass_cost;
    z->sched = sched;
    if (sched) total = real_tokens;
    z->run2 = run2;
    z->nlevels = (int)zl.level_start.size() - 1;
    z->total_tokens = total;
    z->max_ntok = 0;
    for (const ZipChunk& c : chunks) z->max_ntok = std::max(z->max_ntok, c.ntok);
    z->host_chunks.swap(chunks);      // no copy, cannot throw
    int rc;
    if ((rc = z->tokens.reserve(std::max<size_t>(flat.size(), 64))) || (rc = z->chunks.reserve(sizeof(ZipChunk) * std::max(ns, 1))) ||
        (rc = z->pairs.reserve(std::max<size_t>(zl.pairs.size(), 16))) || (rc = z->levels.reserve(sizeof(int) * zl.level_start.size()))) {
        z->tokens.release(); z->chunks.release(); z->pairs.release(); z->levels.release();
        delete z;
        return rc;
    }
    cudaError_t e = cudaSuccess;
    if (!flat.empty()) e = cudaMemcpy(z->tokens.p, flat.data(), flat.size(), cudaMemcpyHostToDevice);
    if (e == cudaSuccess && ns) e = cudaMemcpy(z->chunks.p, z->host_chunks.data(), sizeof(ZipChunk) * ns, cudaMemcpyHostToDevice);
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

// chunk list of z cut into segments of seglen tokens (a multiple of 16) for a K-state model, cached per (K, seglen).
// parts_local > 0: the chunks are parts part_first .. part_first + parts_local - 1 of ONE long chunk (parts mode): every
// part is folded into vec2 only -- part 0 into one vector, a later part (all of whose segments start from unit vectors)
// into the K columns of its transfer matrix, vector c of local part lp at index lp * K + c -- and zip_fold_parts_kernel
// finishes the job after the all-gather.
static int zip_split_build(ZipDevice* z, int K, int seglen, int parts_local, int part_first, ZipSplit** out);
static int zip_split(ZipDevice* z, int K, int seglen, ZipSplit** out, int parts_local = 0, int part_first = 0) {
    try { return zip_split_build(z, K, seglen, parts_local, part_first, out); }
    catch (const std::bad_alloc&) { return fail(IMC_ERR_NOMEM, "out of host memory while cutting chunks into segments"); }
}
static int zip_split_build(ZipDevice* z, int K, int seglen, int parts_local, int part_first, ZipSplit** out) {
    for (ZipSplit* sp : z->splits) if (sp->K == K && sp->seglen == seglen && sp->parts_local == parts_local) { *out = sp; return IMC_OK; }
    std::vector<ZipChunk> chains;
    std::vector<ZipFoldItem> items1, items2;
    int nvec2 = 0;
    if (parts_local > 0) nvec2 = parts_local * K;
    for (const ZipChunk& ch : z->host_chunks) {
        const int nseg = std::max(1, (ch.ntok + seglen - 1) / seglen);
        const int first_chain = (int)chains.size();
        const bool cont = parts_local > 0 && ch.continues > 0;
        for (int sg = 0; sg < nseg; ++sg) {
            ZipChunk c = ch;
            c.tok_off = ch.tok_off + (long long)sg * seglen * (z->spec ? 4 : 1);
            c.ntok = std::max(0, std::min(seglen, ch.ntok - sg * seglen));
            const bool own_start = sg == 0 && !cont;
            for (int col = 0; col < (own_start ? 1 : K); ++col) {
                c.first_sym = own_start ? ch.first_sym : -1 - col;
                c.first_run = sg == 0 ? ch.first_run : 0;
                c.out_index = (int)chains.size();
                chains.push_back(c);
            }
        }
        if (parts_local > 0) {
            const int lp = ch.continues - part_first;         // local part index
            if (!cont) {      // part 0: segment 0 is one chain, segment s >= 1 column c at first_chain + 1 + (s-1) K + c
                items1.push_back({first_chain, first_chain + 1, nseg - 1, lp * K, 0, 1, 0, 0});
            } else {          // later part: segment s column c at first_chain + s K + c; one fold per start column
                for (int col = 0; col < K; ++col)
                    items1.push_back({first_chain + col, first_chain + K, nseg - 1, lp * K + col, 0, 1, 0, 0});
            }
            continue;
        }
        // segment s >= 1, column c sits at first_chain + 1 + (s-1)*K + c
        if (nseg <= 32) {
            items2.push_back({first_chain, first_chain + 1, nseg - 1, ch.out_index, 0, 0, ch.run_sites, ch.run2_sites});
        } else {        // two levels: groups of gs segments folded in parallel, then the groups
            const int gs = (int)std::ceil(std::sqrt((double)nseg)), ngroups = (nseg + gs - 1) / gs;
            const int base2 = nvec2;
            for (int g = 0; g < ngroups; ++g) {
                const int s0 = g * gs, s1 = std::min(nseg, s0 + gs);     // segments [s0, s1)
                if (g == 0) {
                    items1.push_back({first_chain, first_chain + 1, s1 - 1, nvec2++, 0, 1, 0, 0});
                } else {
                    for (int col = 0; col < K; ++col)
                        items1.push_back({first_chain + 1 + (s0 - 1) * K + col, first_chain + 1 + s0 * K, s1 - s0 - 1, nvec2++, 0, 1, 0, 0});
                }
            }
            items2.push_back({base2, base2 + 1, ngroups - 1, ch.out_index, 1, 0, ch.run_sites, ch.run2_sites});
        }
    }
    // the kernel takes chunks in list order, longest first: full segments first, tails last (out_index keeps identity)
    std::vector<ZipChunk> sorted = chains;
    std::stable_sort(sorted.begin(), sorted.end(), [](const ZipChunk& x, const ZipChunk& y) { return x.ntok > y.ntok; });
    ZipSplit* sp = new (std::nothrow) ZipSplit;
    if (!sp) return fail(IMC_ERR_NOMEM, "out of memory");
    sp->K = K; sp->seglen = seglen; sp->nchains = (int)sorted.size(); sp->parts_local = parts_local;
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

template <class C, int THREADS, int MINB, bool SPEC>
static int launch_zip_k(const ZipArgs& a, const ZipPlan& p, int grid, cudaStream_t st) {
    static size_t attr_max = 0;
    if (p.smem > attr_max) {
        CUDA_TRY(cudaFuncSetAttribute(zip_forward_kernel<C, THREADS, MINB, SPEC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem));
        attr_max = p.smem;
    }
    zip_forward_kernel<C, THREADS, MINB, SPEC><<<grid, THREADS, p.smem, st>>>(a);
    return IMC_OK;
}

template <int K, bool SPEC>
static int launch_zip_shape(const ZipArgs& a, const ZipPlan& p, int grid, cudaStream_t st) {
    if constexpr (SPEC && K >= 8) {
        if (p.mma) {
            if constexpr (K <= 12) { if (p.ctas_per_sm == 4 && p.threads == 256) return launch_zip_k<ZipCfgM<K>, 256, 4, true>(a, p, grid, st); }
            if constexpr (K <= 24) {
                if (p.ctas_per_sm == 4) return launch_zip_k<ZipCfgM<K>, 128, 4, true>(a, p, grid, st);
                if (p.ctas_per_sm == 2) return launch_zip_k<ZipCfgM<K>, 256, 2, true>(a, p, grid, st);
                return launch_zip_k<ZipCfgM<K>, 512, 1, true>(a, p, grid, st);
            } else return launch_zip_k<ZipCfgM<K>, 256, 1, true>(a, p, grid, st);
        }
    }
    if constexpr (K >= 10) {
        if (p.lanes == 32) {
            if constexpr (K <= 24) return launch_zip_k<ZipCfg32<K>, 512, 1, SPEC>(a, p, grid, st);
            else return launch_zip_k<ZipCfg32<K>, 256, 1, SPEC>(a, p, grid, st);
        }
    }
    if constexpr (K >= 8) {
        if (p.lanes == 4) {
            if constexpr (K <= 24) { if (p.ctas_per_sm == 2) return launch_zip_k<ZipCfg4<K>, 256, 2, SPEC>(a, p, grid, st); }
            return launch_zip_k<ZipCfg4<K>, 256, 1, SPEC>(a, p, grid, st);
        }
    }
    if constexpr (K <= 24) {
        if (p.ctas_per_sm == 2) return launch_zip_k<ZipCfg8<K>, 256, 2, SPEC>(a, p, grid, st);
        return launch_zip_k<ZipCfg8<K>, IMC_ZIP8_THREADS, 1, SPEC>(a, p, grid, st);
    } else {
        return launch_zip_k<ZipCfg8<K>, 256, 1, SPEC>(a, p, grid, st);
    }
}

static int launch_zip(ZipArgs a, const ZipPlan& p, cudaStream_t st) {
    // persistent CTAs: one per resident slot, but never more than there are (point, warp-load of quads) units
    const int cpw = 32 / p.lanes, nunits = (a.nchunks + cpw - 1) / cpw * (a.nseg > 1 ? a.nseg : 1), nw = p.threads / 32;
    (void)nw;
    const long long units = (long long)a.N * nunits;     // scarce work spreads one warp-load per CTA over the SMs
    const int sms = g_ctx.sm_count > 0 ? g_ctx.sm_count : 148;
    const int grid = (int)std::min<long long>(units, (long long)sms * p.ctas_per_sm);
    // with fewer warp-loads than warps on the machine, let only as many warps per CTA claim work as it takes to cover
    // them: the chains then spread over all SMs instead of piling onto the first CTAs that arrive
    a.active_warps = (int)std::min<long long>(p.threads / 32, std::max<long long>(1, (units + grid - 1) / grid));
    // pipelined pieces: a point has only (chunks / chains per warp) chains of pieces; more warps than that would just poll
    if (a.nseg > 1) a.active_warps = std::min(a.active_warps, (a.nchunks + cpw - 1) / cpw);
    switch (zip_tile(a.K)) {
#define X(k) case k: return p.spec ? launch_zip_shape<k, true>(a, p, grid, st) : launch_zip_shape<k, false>(a, p, grid, st);
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


// ------------------------------------------------------------------------------------------ run tokens (spectral form)
extern "C" int imc_seqset_run_info(imc_seqset* set, int K, int* run_sym, int* ids_available, int* ids_used, int64_t* tokens, int* levels) {
    if (!set) return fail(IMC_ERR_INVALID, "NULL set");
    if (run_sym) *run_sym = set->fold_sym;
    if (ids_available) *ids_available = set->run_merges.size();
    if (!ids_used && !tokens && !levels) return IMC_OK;
    ZipPlan plan;
    int rc = zip_plan(K, set->nsym, set->run_merges.size(), &plan, 0, true);
    if (rc) return rc;
    if (ids_used) *ids_used = plan.M;
    if (levels) *levels = (int)zip_levels(set->run_merges, plan.M).level_start.size() - 1;
    if (tokens) {
        const ZipMerges& mg = set->run_merges;
        std::vector<long long> len(mg.size(), 1);      // entries a full-dictionary token expands to over the first plan.M ids
        for (int id = std::max(plan.M, set->nsym); id < mg.size(); ++id) {
            const auto& pr = mg.pairs[id - set->nsym];
            len[id] = len[pr[0]] + len[pr[1]];
        }
        long long total = 0;
        for (const auto& t : set->run_tok_full) for (uint32_t x : t) total += len[x & 0xffu];
        *tokens = total;
    }
    return IMC_OK;
}

extern "C" int imc_seqset_run_pairs(imc_seqset* set, uint8_t* pairs_out, int capacity_pairs) {
    if (!set || !pairs_out) return fail(IMC_ERR_INVALID, "NULL argument");
    const auto& pairs = set->run_merges.pairs;
    if (capacity_pairs < (int)pairs.size()) return fail(IMC_ERR_INVALID, "capacity %d < %zu pairs", capacity_pairs, pairs.size());
    for (size_t i = 0; i < pairs.size(); ++i) { pairs_out[2 * i] = pairs[i][0]; pairs_out[2 * i + 1] = pairs[i][1]; }
    return IMC_OK;
}

extern "C" int imc_seqset_run_tokens(imc_seqset* set, int chunk, int ids, uint32_t* out, int64_t capacity, int64_t* ntokens, int* first_run) {
    if (!set || !ntokens) return fail(IMC_ERR_INVALID, "NULL argument");
    if (chunk < 0 || chunk >= set->n_chunks) return fail(IMC_ERR_INVALID, "chunk %d out of range", chunk);
    if (ids < set->nsym || ids > set->run_merges.size()) return fail(IMC_ERR_INVALID, "ids must be in [%d, %d]", set->nsym, set->run_merges.size());
    const int k = set->stream_of_chunk[chunk];
    if (first_run) *first_run = k < 0 ? 0 : set->first_run[k];
    if (k < 0) { *ntokens = 0; return IMC_OK; }
    std::vector<uint32_t> tok;
    try { run_expand(set->run_merges, set->run_tok_full[k], ids, tok); }
    catch (const std::bad_alloc&) { return fail(IMC_ERR_NOMEM, "out of host memory"); }
    *ntokens = (int64_t)tok.size();
    if (out) {
        if (capacity < (int64_t)tok.size()) return fail(IMC_ERR_INVALID, "capacity %lld < %zu tokens", (long long)capacity, tok.size());
        if (!tok.empty()) memcpy(out, tok.data(), tok.size() * sizeof(uint32_t));
    }
    return IMC_OK;
}

// Host-only views of the aligned form for a K-state model (tests, tools/schedule_sim.py): no CUDA call is made.
static int align_host_streams(imc_seqset* set, int K, ZipHostStreams* h, int* M_out) {
    if (!set) return fail(IMC_ERR_INVALID, "NULL set");
    std::lock_guard<std::recursive_mutex> lock(set->serial.mu);       // the two-run encoding is prepared once per set (host side only: no event)
    int rc = seqset_run2_prepare(set);
    if (rc) return rc;
    if (set->run2_state != 1) return fail(IMC_ERR_UNSUPPORTED, "this set has no second run symbol");
    ZipPlan plan;
    if ((rc = zip_plan(K, set->nsym, set->run2_merges.size(), &plan, 0, true, true, true))) return rc;
    try {
        ZipLevels zl = zip_levels(set->run2_merges, plan.M);
        zip_host_streams(set->run2_merges, set->run2_tok_full, plan.M, zl, h);
    } catch (const std::bad_alloc&) { return fail(IMC_ERR_NOMEM, "out of host memory"); }
    if (M_out) *M_out = plan.M;
    return IMC_OK;
}

extern "C" int imc_seqset_align_info(imc_seqset* set, int K, int stall, int64_t* lock_steps, double* est_passes, int64_t* aligned_steps,
                                     int* stall_used, int* hot_id) {
    ZipHostStreams h;
    int rc = align_host_streams(set, K, &h, nullptr);
    if (rc) return rc;
    try {
        const int ns = (int)h.order.size(), nq = (ns + 7) / 8;
        if (stall <= 0) stall = zip_align_pick_stall(h);
        long long lock = 0;
        for (int i = 0; i < ns; i += 8) lock += (long long)h.rtok[h.order[i]].size();
        std::vector<long long> st(nq, 0);
        if (!parallel_for(nq, [&](int q) { zip_align_quad(zip_quad_tokens(h, q), h.hot_id, stall, nullptr, &st[q]); })) throw std::bad_alloc();
        long long al = 0;
        for (long long x : st) al += x;
        if (lock_steps) *lock_steps = lock;
        if (est_passes) *est_passes = h.est_passes;
        if (aligned_steps) *aligned_steps = al;
        if (stall_used) *stall_used = stall;
        if (hot_id) *hot_id = h.hot_id;
    } catch (const std::bad_alloc&) { return fail(IMC_ERR_NOMEM, "out of host memory"); }
    return IMC_OK;
}

// words of warp-load `quad` as out[nchains][steps]: stall > 0 the aligned streams, stall <= 0 the chains' own token streams,
// each padded to the longest with the padding word (RUN_NOP_BIT | 0xff)
extern "C" int imc_seqset_align_quad(imc_seqset* set, int K, int quad, int stall, uint32_t* out, int64_t capacity, int64_t* steps, int* nchains) {
    if (!steps || !nchains) return fail(IMC_ERR_INVALID, "NULL argument");
    ZipHostStreams h;
    int rc = align_host_streams(set, K, &h, nullptr);
    if (rc) return rc;
    const int ns = (int)h.order.size();
    if (quad < 0 || quad * 8 >= ns) return fail(IMC_ERR_INVALID, "quad %d out of range", quad);
    try {
        const auto t = zip_quad_tokens(h, quad);
        std::vector<std::vector<uint32_t>> streams(t.size());
        if (stall > 0) zip_align_quad(t, h.hot_id, stall, &streams, nullptr);
        else {
            size_t mx = 0;
            for (auto* v : t) mx = std::max(mx, v->size());
            for (size_t i = 0; i < t.size(); ++i) { streams[i] = *t[i]; streams[i].resize(mx, RUN_NOP_BIT | 0xffu); }
        }
        *nchains = (int)t.size();
        *steps = streams.empty() ? 0 : (int64_t)streams[0].size();
        if (out) {
            if (capacity < *steps * *nchains) return fail(IMC_ERR_INVALID, "capacity %lld < %lld words", (long long)capacity, (long long)(*steps * *nchains));
            for (size_t i = 0; i < streams.size(); ++i) if (*steps) memcpy(out + i * (size_t)*steps, streams[i].data(), (size_t)*steps * sizeof(uint32_t));
        }
    } catch (const std::bad_alloc&) { return fail(IMC_ERR_NOMEM, "out of host memory"); }
    return IMC_OK;
}

// how the last spectral call on this set split its points (synchronises the device; tests)
extern "C" int imc_seqset_spectral_counts(imc_seqset* set, int* ok_points, int* plain_points) {
    if (!set) return fail(IMC_ERR_INVALID, "NULL set");
    int c[2] = {0, 0};
    if (set->d_lists.p) {
        CUDA_TRY(cudaDeviceSynchronize());
        CUDA_TRY(cudaMemcpy(c, set->d_lists.p, sizeof c, cudaMemcpyDeviceToHost));
    }
    if (ok_points) *ok_points = c[0];
    if (plain_points) *plain_points = c[1];
    return IMC_OK;
}

// passes of the MMA form executed by this process so far (each pass = KT x NT DMMAs of 512 flop per warp); bench.py
extern "C" int imc_mma_passes(int64_t* passes_out) {
    if (!passes_out) return fail(IMC_ERR_INVALID, "NULL argument");
    *passes_out = 0;
    if (!g_mma_passes) return IMC_OK;
    unsigned long long v = 0;
    CUDA_TRY(cudaDeviceSynchronize());
    CUDA_TRY(cudaMemcpy(&v, g_mma_passes, sizeof v, cudaMemcpyDeviceToHost));
    *passes_out = (int64_t)v;
    return IMC_OK;
}
