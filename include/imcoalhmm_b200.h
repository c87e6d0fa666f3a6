/*
 * imcoalhmm_b200.h -- C ABI of the B200-native IMCoalHMM likelihood hot path.
 *
 * The reference (harvardinformatics/IMCoalHMM) has no C ABI of its own: its forward arithmetic lives in
 * the external CPython extension `ziphmm`, reached at exactly two call sites
 *     /root/reference/src/IMCoalHMM/hmm.py:16      ziphmm.preprocess_raw_observations(obs, NSYM)
 *     /root/reference/src/IMCoalHMM/hmm.py:20-21   ziphmm.zip_forward(pi, T, E, sym2pair, new_obs, NSYM, new_nsyms)
 * and its model build is Python (model.py:44-49).  The entry points below are what a ctypes / cffi binding
 * for that path binds instead (see INTEGRATION.md for the stub).  Plain pointers and sizes only.
 *
 * Conventions
 *   - every function returns IMC_OK (0) or a negative IMC_ERR_* code; imc_last_error() returns the
 *     message of the last failure on the calling thread.
 *   - all floating point is IEEE binary64.  pi is [K], T is [K][K] row-major with T[i][j] = P(next=j | cur=i)
 *     (transitions.py:243-246), E is [K][S] with E[state][symbol] (emissions.py:95-99).
 *   - batched arrays are [N][...] contiguous: pi [N][K], T [N][K][K], E [N][K][S], out [N].
 *   - caller owns every host buffer; the library copies in/out and keeps no host pointer past return.
 *   - there is NO CPU fallback: calls that need the GPU fail with IMC_ERR_CUDA when no device is usable.
 *   - CUDA is initialised lazily by the first call that needs the device (imc_init or a forward call),
 *     so a process may create sequences, fork (mcmc.py:113-122), and only then touch the GPU in the child.
 */
#ifndef IMCOALHMM_B200_H
#define IMCOALHMM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IMC_OK 0
#define IMC_ERR_INVALID (-1)     /* bad argument (shape, NULL, symbol out of range, ...) */
#define IMC_ERR_CUDA (-2)        /* CUDA runtime failure or no usable device */
#define IMC_ERR_NOMEM (-3)
#define IMC_ERR_UNSUPPORTED (-4) /* valid request outside what the kernels cover (e.g. K > 128) */
#define IMC_ERR_IO (-5)          /* file could not be read / parsed */

const char* imc_last_error(void);
int imc_version(void);

/* Select the device for the calling process (default 0) and create the context.  Optional: the first
 * forward call does it implicitly on device 0 (or the device given by an earlier imc_init). */
int imc_init(int device);
int imc_device_count(int* count_out);

/* ---- sequences: one alignment chunk == one reference Forwarder (hmm.py:12-16) --------------------- */
typedef struct imc_seq imc_seq;

/* obs[L] symbols in [0, nsym).  Replaces np.array(map(int, ...), dtype=np.int32) + preprocess (hmm.py:14-16). */
int imc_seq_create(const int32_t* obs, int64_t L, int nsym, imc_seq** out);
int imc_seq_create_u8(const uint8_t* obs, int64_t L, int nsym, imc_seq** out);
/* Text file of whitespace-separated integers (hmm.py:13-14; written by scripts/prepare-alignments.py:93-105). */
int imc_seq_from_file(const char* path, int nsym, imc_seq** out);
/* Ingest.  Two aligned sequences -> pairwise symbols by the rule of scripts/prepare-alignments.py:99-111:
 * upper-case; 2 if either base is not one of ACGT, 0 if equal, 1 otherwise (NSYM = 3).  The FASTA reader takes the two
 * named records (record name = header text up to the first blank), or the only two when both names are NULL. */
int imc_seq_from_pair(const char* seq1, const char* seq2, int64_t L, imc_seq** out);
int imc_seq_from_fasta(const char* path, const char* name1, const char* name2, imc_seq** out);
/* 2, 3 or 4 aligned sequences: pairs as above; triplets i1 + 4 i2 + 16 i3 with 64 for any base outside ACGT (NSYM = 65),
 * quartets i1 + 4 i2 + 16 i3 + 32 i4 with 128 -- prepare-alignments.py:113-190, weights as in the script (clean quartet
 * columns therefore reach 159 and 128 is ambiguous; NSYM = 160). */
int imc_seq_from_columns(const char* const* seqs, int n_seqs, int64_t L, imc_seq** out);
int imc_seq_from_fasta_n(const char* path, const char* const* record_names, int n_names, imc_seq** out);
/* The same for any supported alignment format -- prepare-alignments.py:41,66 hands its <input format> argument to BioPython;
 * here: "fasta" (or NULL), "phylip" (interleaved, names in the first 10 columns), "phylip-relaxed" (interleaved, names end at
 * the first blank), "phylip-sequential".  n_names = 0 takes the only two records of the file, else 2, 3 or 4 named ones. */
int imc_seq_from_alignment(const char* path, const char* format, const char* const* record_names, int n_names, imc_seq** out);
/* Binary container for a symbol sequence (2 bits per symbol for NSYM <= 4): 16x smaller than the text format. */
int imc_seq_save(const imc_seq* seq, const char* path);
int imc_seq_load(const char* path, imc_seq** out);
int imc_seq_length(const imc_seq* seq, int64_t* L_out);
int imc_seq_nsym(const imc_seq* seq, int* nsym_out);
/* counts[nsym] <- number of occurrences of each symbol */
int imc_seq_symbol_counts(const imc_seq* seq, int64_t* counts);
/* copy the symbols back (uint8), e.g. to check the encoding byte-for-byte */
int imc_seq_symbols(const imc_seq* seq, uint8_t* out, int64_t capacity);
int imc_seq_destroy(imc_seq* seq);

/* ---- concurrency -------------------------------------------------------------------------------------------------
 * A handle (imc_seqset, imc_model) owns per-handle scratch on the device, so its calls are serialised: every call on a
 * handle first waits -- on the host with a mutex, on the device with an event -- for the previous call on the SAME handle,
 * whatever thread or CUDA stream that call came from.  Calls on different handles run concurrently.  With a communicator
 * the collective calls of a process are likewise executed in issue order, and all ranks must issue the same sequence.
 * Errors are per thread (imc_last_error). */

/* ---- sequence sets: the list of forwarders a Likelihood sums over (likelihood.py:22-25,33), packed into
 * the interleaved 2-bit layout the kernels stream from HBM.  Host-only until the first forward call. */
typedef struct imc_seqset imc_seqset;

int imc_seqset_create(const imc_seq* const* seqs, int C, imc_seqset** out);
/* Fewer chunks than GPUs: ONE long chunk cut into parts_total consecutive parts, every rank holding n_local of
 * them (parts[i] = part part_first + i, part_first = rank * n_local).  A forward / loglik call on such a set is collective
 * over a communicator of parts_total / n_local ranks: every part runs in segmented mode -- part 0 from pi, every later part
 * from the K unit vectors, which yields its K x K transfer matrix --, the per-rank blocks (float64 [N][n_local K][K+1]) are
 * exchanged with ONE ncclAllGather, and every rank folds alpha <- P_part alpha over the parts in order: the log-likelihood of
 * the whole chunk, the same bits on every rank.  With parts_total == n_local a single process holds all parts and no
 * communicator is needed.  K <= 40, at most 32768 parameter points per call. */
int imc_seqset_create_parts(const imc_seq* const* parts, int n_local, int part_first, int parts_total, imc_seqset** out);
int imc_seqset_destroy(imc_seqset* set);
/* any output may be NULL */
int imc_seqset_info(const imc_seqset* set, int* n_chunks, int64_t* total_sites, int64_t* packed_bytes);

/* The zipHMM-style preprocessing of the set (hmm.py:16 preprocess_raw_observations -> new_obs, sym2pair, new_nsyms):
 * ids_available = new_nsyms of the set's shared dictionary (<= 256); for a K-state model the kernel keeps the
 * first ids_used of them in shared memory; tokens = total length of the re-encoded chunks (positions 1..L-1)
 * over those ids; levels = depth of the pair dictionary.  Host only; any output may be NULL. */
int imc_seqset_zip_info(imc_seqset* set, int K, int* ids_available, int* ids_used, int64_t* tokens, int* levels);

/* sym2pair of the set's dictionary in creation order: pairs_out[2*i], pairs_out[2*i+1] = (left, right) of id nsym + i. */
int imc_seqset_zip_pairs(imc_seqset* set, uint8_t* pairs_out, int capacity_pairs);
/* new_obs of one chunk (index as given to imc_seqset_create) over the first `ids` dictionary ids, creation-order
 * numbering.  Position 0 of the chunk is not part of the stream (alpha_0 = pi o E[:,o_0] carries no transition).
 * out may be NULL to query the length. */
int imc_seqset_zip_tokens(imc_seqset* set, int chunk, int ids, uint8_t* out, int64_t capacity, int64_t* ntokens);

/* Run tokens: the second re-encoding of the set, used by the spectral form of the compressed kernel.  The most frequent
 * symbol r of the set ("run symbol") leaves the dictionary: a chunk (positions 1..L-1) is first_run sites of r followed by
 * tokens  word = id | n << 8  = "dictionary entry id, then n sites of r" (n <= 4095; longer runs continue with entries whose
 * id is r itself).  Pairs are learned over the entries (left part without a run).  In the eigenbasis of
 * C_r = diag(E[:,r]) T^T a run of any length is a diagonal matrix, so a token costs one mat-vec however long its run --
 * exact like the pair dictionary (hmm.py:16,20-21 compute the same likelihood).  Same conventions as the zip_* calls above. */
int imc_seqset_run_info(imc_seqset* set, int K, int* run_sym, int* ids_available, int* ids_used, int64_t* tokens, int* levels);
int imc_seqset_run_pairs(imc_seqset* set, uint8_t* pairs_out, int capacity_pairs);
int imc_seqset_run_tokens(imc_seqset* set, int chunk, int ids, uint32_t* out, int64_t capacity, int64_t* ntokens, int* first_run);
/* how the last spectral forward call on this set split its parameter points: served by the spectral form / by the plain
 * form (synchronises the device; for tests).  Both 0 if no spectral call was made yet. */
int imc_seqset_spectral_counts(imc_seqset* set, int* ok_points, int* plain_points);

/* ---- forward log-likelihood ------------------------------------------------------------------------ */
/* logL_out[0] = sum over the set's chunks of log P(chunk | pi, T, E).  Replaces
 * sum(f.forward(pi, T, E) for f in forwarders)  (hmm.py:19-21, likelihood.py:33). */
int imc_forward(imc_seqset* set, int K, int S, const double* pi, const double* T, const double* E,
                double* logL_out);

/* The same for N parameter points in one call (new; the reference evaluates one point per call). */
int imc_forward_batch(imc_seqset* set, int N, int K, int S, const double* pi, const double* T,
                      const double* E, double* out);

/* Device-resident variant: all four pointers are device memory on the set's device; work is enqueued on
 * `stream` (a cudaStream_t, NULL = legacy default stream) and NOT synchronised. */
int imc_forward_batch_dev(imc_seqset* set, int N, int K, int S, const double* d_pi, const double* d_T,
                          const double* d_E, double* d_out, void* stream);

/* ---- multi-GPU (one process per GPU): chunks are sharded over the ranks, every rank scores the same parameter batch on
 * its own sequence set, and once a communicator exists every forward / loglik entry point above and below returns the SUM
 * over ranks -- one all-reduce of float64[N] per batch, the only collective of the path (likelihood.py:33 sums over
 * forwarders; here the forwarders live on different GPUs).  Rank 0 draws the id and hands its bytes to the other ranks by
 * any means (file, pipe, MPI, torch.distributed); all ranks then call imc_comm_init after imc_init(device).  NCCL is
 * loaded with dlopen at that point (no link-time dependency) and used for the bootstrap.
 *
 * On one node the all-reduce is FUSED into the kernel that sums a rank's chains: imc_comm_init maps a small mailbox of
 * every rank into every other rank (CUDA IPC), the reduction kernel stores each point's partial sum into all mailboxes
 * with peer-to-peer writes over NVLink, and its last block adds the rows up in rank order (the same bits on every rank;
 * no extra launch, no NCCL call per batch).  Where the mailboxes cannot be mapped (ranks on different nodes, IPC not
 * permitted) or a batch has more than 65536 points, the sum is one ncclAllReduce instead; imc_comm_info tells which.
 * All ranks must issue the same sequence of forward / loglik calls (as with any collective), set the options
 * "comm_fused" (default 1; 0 = always NCCL; read by imc_comm_init) and "comm_enabled" (default 1; 0 = calls return this
 * rank's partial sums) identically, and call imc_comm_destroy together. */
#define IMC_COMM_ID_BYTES 128
int imc_comm_unique_id(void* id_out, int capacity);
int imc_comm_init(int nranks, int rank, const void* nccl_id);
int imc_comm_destroy(void);
/* any pointer may be NULL; *fused = 1 when the peer-memory kernel performs the all-reduce */
int imc_comm_info(int* nranks, int* rank, int* fused);

/* ---- batched model build: theta -> (pi, T, E) on the GPU ----------------------------------------------
 * Replaces Model.build_hidden_markov_model (model.py:44-49) and everything under it (state_spaces.py, CTMC.py,
 * transitions.py, emissions.py, break_points.py and the model files) for N parameter points per call.
 *   kind 0  IsolationModel(no_hmm_states)                          iparams {K}             theta[3]   isolation_model.py:94-122
 *   kind 1  IsolationMigrationModel(no_mig, no_anc)                iparams {n_mig, n_anc}  theta[5]   isolation_with_migration_model.py:116-164
 *   kind 2  VariableCoalescenceRateIsolationModel(intervals, split) iparams {est_split, n_epochs, intervals...}
 *                                                                   theta[n_epochs+1(+1)]  variable_coalescence_rate_isolation_model.py:90-178
 *   kind 3  VariableCoalAndMigrationRateModel(initial, intervals)  iparams {initial 0|1|2, n_epochs, intervals...}
 *                                                                   theta[4 n_epochs+1]    variable_migration_model.py:86-181
 *   kind 4  IsolationMigrationEpochsModel(epochs, no_mig, no_anc)  iparams {e, n_mig, n_anc} theta[3+(2e+1)+e]
 *                                                                   isolation_with_migration_model_epochs.py:133-211
 * status[n]: 0 ok, 1 invalid parameters (some theta <= 0; model.py:32-42 -> logL = -inf, likelihood.py:29-30),
 *            2 joint matrix does not sum to 1 within 7 decimals (the reference raises AssertionError, transitions.py:239). */
typedef struct imc_model imc_model;
int imc_model_create(int kind, const int32_t* iparams, int n_iparams, imc_model** out);
int imc_model_info(const imc_model* model, int* K_out, int* P_out);
int imc_model_destroy(imc_model* model);
/* host arrays: theta [N][P] in; pi [N][K], T [N][K][K], E [N][K][3], status [N] (may be NULL) out */
int imc_model_build_batch(imc_model* model, int N, const double* theta, double* pi, double* T, double* E,
                          int32_t* status);
int imc_model_build_batch_dev(imc_model* model, int N, const double* d_theta, double* d_pi, double* d_T,
                              double* d_E, int32_t* d_status, void* stream);
/* Break points (break_points.py:9-30 exp, :60-78 uniform, :81-108 psmc).  imc_break_points is the host form of the three
 * functions -- kind 0: exp_break_points(n, coal_rate = a, offset = b); 1: uniform_break_points(n, start = a, end = b);
 * 2: psmc_break_points(n, t_max = a, mu = b, offset = c) -- built from the same expressions as the models' constant
 * tables; imc_model_break_points returns what the model-build kernel computed on the device for each parameter point
 * (out[N][K]; NaN rows where status is 1). */
int imc_break_points(int kind, int no_intervals, double a, double b, double c, double* out);
int imc_model_break_points(imc_model* model, int N, const double* theta, double* out);
/* fused theta -> logL (model build and forward on the device, no host round trip in between; replaces
 * Likelihood.__call__, likelihood.py:27-33, for N points).  out[n] = -inf where status[n] == 1, NaN where 2. */
int imc_loglik_batch(imc_model* model, imc_seqset* set, int N, const double* theta, double* out, int32_t* status);
int imc_loglik_batch_dev(imc_model* model, imc_seqset* set, int N, const double* d_theta, double* d_out,
                         int32_t* d_status /* may be NULL */, void* stream);
/* The two-locus ancestry state spaces (state_spaces.py:7-116): space 0 Isolation, 1 Single, 2 Migration.
 * States are numbered in a canonical order; lineages[i][k] = (population << 4) | (left mask << 2) | right mask,
 * 0xff = unused slot.  Any output may be NULL. */
int imc_statespace_describe(int space, int* n_states, int* n_edges, int* counts, int* special, int32_t* edges,
                            int32_t* classes, uint8_t* lineages);

/* ---- knobs and introspection (tests, bench) ---------------------------------------------------------- */
/* key "forward_kernel": 0 auto (zip where instantiated), 1 generic (shared-memory state, per-step rescale),
 *                       2 lane-pair DFMA, 3 DMMA tiles (1-3 walk every site), 4 zip (compressed token streams,
 *                       dictionary matrices in shared memory).  Forcing a kernel that does not cover (K, S)
 *                       returns IMC_ERR_UNSUPPORTED.
 * key "zip_ctas_per_sm": 1 = one persistent CTA per SM with all of shared memory for the dictionary, 2 = two CTAs of
 *                       256 threads with half each (K <= 24 only); 0 = auto.
 * key "zip_lanes":      lanes that share one chain's mat-vec in the zip kernel: 8, 4 (K >= 8), 32 (K >= 10: one warp
 *                       per chain, the lowest latency per step) or 0 = auto.
 * key "zip_segment_tokens": segmented mode of the zip kernel for chain-scarce calls (few chunks x few points): chunks
 *                       are cut into segments of this many tokens whose K x K transfer matrices are computed
 *                       column by column in parallel and folded afterwards.  0 = auto, -1 = never, > 0 = forced length.
 * key "zip_max_entries": cap on the dictionary ids used (0 = as many as fit).
 * key "zip_pipeline":   pipelined mode of the zip kernel for launches with few work units per warp: every chunk is walked
 *                       in this many pieces that are separate, ordered work units (a piece starts from the state its
 *                       predecessor left in global memory), so that the SMs finish together; bit-identical results.
 *                       0 = auto (about 40 units per warp), 1 = off, 2..32 = forced.
 * key "zip_spectral":   spectral form of the zip kernel (run tokens, see imc_seqset_run_info): per parameter point C_r is
 *                       diagonalised on the device (symmetric Jacobi; needs pi, E[:,r] > 0 and diag(pi) T symmetric, which
 *                       every model of the reference guarantees, transitions.py:231-246); points that do not qualify are
 *                       served by the plain form in the same call.  0 = auto (where run tokens are >= 15 % fewer than
 *                       dictionary tokens), 1 = always, 2 = never.
 * key "zip_mma":        MMA shape of the spectral form (state counts >= 7): the eight chains of a warp are the rows of
 *                       mma.sync.m8n8k4.f64 tiles and the most frequent entry's matrix is held as B fragments in registers,
 *                       so a step moves no matrix through shared memory; chains on other entries take one extra pass per
 *                       distinct entry.  0 = auto (where one entry is >= 50 % of the run tokens), 1 = always, 2 = never.
 * key "zip_run2":       two-run form of the MMA shape: the second run symbol (missing data in a pairwise alignment) is
 *                       diagonalised too, so that a block of m such sites costs two fixed matrices (B^-1, B) and a diagonal
 *                       instead of a product of power-of-two dictionary entries.  0 = auto (where it lowers the expected
 *                       number of DMMA passes: state counts whose dictionary hardly fits, K >= 32), 1 = always, 2 = never.
 * key "zip_align":      aligned form of the MMA shape (the chains of a warp follow one host-built schedule with a single dictionary
 *                       entry per warp-step; two-run tokens): 0 = auto (cost model; IMC_TRACE_PLAN=1 in the environment prints its
 *                       decision to stderr), 1 = wherever it applies, 2 = off.
 * key "zip_mma_shape":  launch shape of the MMA form: 0 = auto, 1 = one CTA of 512 threads per SM (256 for K >= 32),
 *                       2 = two CTAs of 256, 3 = four CTAs of 256 threads with 64 registers (K <= 12), 4 = four CTAs of 128.
 * key "zip_spectral_force_bad": 1 = treat every point as not qualifying (exercises the plain-form pass; tests).
 * key "comm_fused", "comm_enabled": see the multi-GPU section above.
 * key "comm_timeout_ms": how long the fused all-reduce waits for a peer's partial sums (default 30000).  On a timeout the call's
 *                       results are NaN, the synchronous entry points return IMC_ERR_CUDA naming the rank, and every later
 *                       collective call fails until imc_comm_destroy: a dead or out-of-step rank is an error, not a hang.
 * key "dmma_mtiles":    M-tiles (of 8 chains) per warp for the DMMA kernel (1, 2 or 4; 0 = auto).
 * key "fold_emission":  1 fold the most frequent symbol's emission column into the register copy
 *                       of T where E[:,s0] > 0; 0 (default; measured faster on B200) = always multiply by the emission row. */
int imc_set_option(const char* key, int64_t value);
int imc_get_option(const char* key, int64_t* value_out);
/* Measured FP64 peaks of the bound device in TFLOP/s: plain DFMA and DMMA.8x8x4 (mma.sync f64) loops,
 * best of 3 after warm-up (~100 ms).  bench.py uses the larger as the roofline denominator. */
int imc_measure_fp64_peak(double* dfma_tflops, double* dmma_tflops);
/* number of kernel launches issued by this library since load (for bench.py's gpu_launches) */
int64_t imc_kernel_launches(void);
/* MMA passes executed by this process so far (each = KT x NT mma.sync.m8n8k4.f64 of 512 flop per warp, KT = ceil(K/4),
 * NT = ceil(K/8)): the executed FP64 tensor work behind bench.py's roofline.  Synchronises the device. */
/* Aligned form of the MMA kernel (host only, no CUDA call): for a K-state model, the warp-steps of the two-run token streams in
 * lock step (sum over warp-loads of 8 chains of the longest stream) with the expected passes per step, and the warp-steps (= passes)
 * of the host-built schedules with one dictionary entry per step.  stall = 0: pick the threshold as the library does. */
int imc_seqset_align_info(imc_seqset* set, int K, int stall, int64_t* lock_steps, double* est_passes, int64_t* aligned_steps,
                          int* stall_used, int* hot_id);
/* The words of warp-load `quad` as out[nchains][steps] (tests): stall > 0 the aligned streams (bit 21 = no token of this chain in this
 * step, bits 0-7 the step's entry; the padding to a multiple of 8 steps names the hot entry), stall <= 0 the chains' own streams
 * padded to the longest with the word (1 << 21) | 0xff. */
int imc_seqset_align_quad(imc_seqset* set, int K, int quad, int stall, uint32_t* out, int64_t capacity, int64_t* steps, int* nchains);
int imc_mma_passes(int64_t* passes_out);
/* name of the forward kernel chosen by the last forward call on this thread
 * ("generic", "pair", "dmma", "zip", "zip-segmented", "zip-warp": one warp per chain, chosen for chain-scarce calls;
 * "zip-spectral", "zip-spectral-segmented", "zip-spectral-warp": the same three shapes over run tokens;
 * "zip-spectral-mma", "zip-spectral-mma-segmented": the MMA shape of the spectral form; "zip-spectral-mma2...": its
 * two-run form) */
const char* imc_last_forward_kernel(void);

#ifdef __cplusplus
}
#endif
#endif /* IMCOALHMM_B200_H */
