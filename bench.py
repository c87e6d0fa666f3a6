#!/usr/bin/env python
"""bench.py -- forward sites x parameter-points / second on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c3_1gpu|ns_1gpu|c5_1gpu|c1|...]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one batched log-likelihood evaluation: N_theta parameter points scored on every alignment
chunk of the (per-rank) shard, logL[N_theta] out.  Rank 0 prints ONE JSON line.

  value      device-timed (CUDA events on the launching stream, max over ranks) with the token streams and
             the parameter batch resident in HBM; an L2 flush (256 MiB write) separates timed steps.
  e2e        the same metric through the public host API (Model.batched_log_likelihood -> imc_loglik_batch):
             theta[N,P] copied from pinned host memory, logL[N] + status[N] copied back, every step, wall clock
             around the calls.  The sequences stay resident (uploaded once when the Forwarders are built, exactly
             like the reference preprocesses once in Forwarder.__init__, hmm.py:12-16).
  roofline   dominant kernel (zip_forward_kernel) against the pipe that bounds it, the shared-memory pipe: achieved =
             chain-steps * 8K^2 bytes / kernel time (chain-steps = tokens of the form that ran x points), peak =
             128 B/clk/SM.  The SURVEY 8(d) view (plain-forward flops sites*points*(2K^2+3K) / kernel time against
             the FP64 peak MEASURED IN THIS RUN) sits beside it under "plain_forward_equivalent".
  parity     EVERY output of the timed launch (all points, all chunks of the shard) against the CPU oracle's
             zipHMM-style forward (oracle/forward_oracle.c, float64); the first 16 points use the (pi,T,E) the
             REFERENCE's own build_hidden_markov_model produced (tests/golden), the others the GPU model build.
  secondary  (default run on one GPU only) the same measurement + parity for the shapes the north star names:
             c3_1gpu (IM model K=20 shard of configs[2]), ns_1gpu (north-star shard), c4_1gpu (4096 MCMC proposals per
             step on a 375 Mbp shard), c5_1gpu (K=40 shard), and the drop-in latency of ONE Likelihood(theta) call on
             the reference's example alignment (c1).
  cpu_baseline  the CPU oracle's zipHMM-style forward (OpenMP over all host cores) on a bounded sample of the same
             workload -- a reported baseline, not the target.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SEED0 = 20261018

WORKLOADS = {
    # BASELINE.json configs[1]: isolation model, 10 intervals, synthetic 100 Mbp, 256 parameter points, 1 B200
    "c2": dict(model="isolation_k10", ctor=("IsolationModel", (10,)), default=[1e-3, 2000.0, 0.4], K=10,
               chunks=100, chunk_len=1_000_000, points=256,
               desc="configs[1]: isolation model K=10, synthetic 100 Mbp (100 x 1 Mbp chunks), 256 parameter points"),
    "c2_small": dict(model="isolation_k10", ctor=("IsolationModel", (10,)), default=[1e-3, 2000.0, 0.4], K=10,
                     chunks=100, chunk_len=50_000, points=256,
                     desc="reduced configs[1] for smoke runs: 100 x 50 kbp, 256 points"),
    # per-GPU slice of configs[2] (IM model, K=20, 1 Gbp, 1024 points on 8 GPUs = 125 chunks/GPU)
    "c3_1gpu": dict(model="im_k10_10", ctor=("IsolationMigrationModel", (10, 10)),
                    default=[1e-3, 1e-3, 2000.0, 0.4, 200.0], K=20, chunks=125, chunk_len=1_000_000, points=1024,
                    desc="configs[2] per-GPU shard: IM model K=10+10, 125 x 1 Mbp chunks, 1024 parameter points"),
    # per-GPU slice of the north-star target: IM model K=10+10, 3 Gbp on 8 GPUs = 375 chunks/GPU, 1024 parameter points
    "ns_1gpu": dict(model="im_k10_10", ctor=("IsolationMigrationModel", (10, 10)),
                    default=[1e-3, 1e-3, 2000.0, 0.4, 200.0], K=20, chunks=375, chunk_len=1_000_000, points=1024,
                    desc="north-star per-GPU shard: IM model K=10+10, 375 x 1 Mbp chunks (3 Gbp on 8 GPUs), 1024 parameter points"),
    # per-GPU slice of configs[3] (isolation-model-mcmc, 4096 lock-step chains, 3 Gbp on 8 GPUs = 375 chunks/GPU): one step
    # = one proposal of every chain scored in one batched call (imcoalhmm_b200.mcmc.BatchedMCMC.step)
    "c4_1gpu": dict(model="isolation_k10", ctor=("IsolationModel", (10,)), default=[1e-3, 2000.0, 0.4], K=10,
                    chunks=375, chunk_len=1_000_000, points=4096,
                    desc="configs[3] per-GPU shard: isolation model K=10, 375 x 1 Mbp chunks, 4096 MCMC proposals per step"),
    # per-GPU slice of configs[4] (psmc-style isolation model, 40 intervals, 3 Gbp, 1024 points on 8 GPUs = 375 chunks/GPU)
    "c5_1gpu": dict(model="psmc_iso_split_4x10", ctor=("VariableCoalescenceRateIsolationModel", ([4] * 10, True)),
                    default=[1e-3] + [1000.0] * 10 + [0.4], K=40, chunks=375, chunk_len=1_000_000, points=1024,
                    desc="configs[4] per-GPU shard: psmc-style isolation model K=40 (10 epochs x 4), 375 x 1 Mbp chunks, "
                         "1024 parameter points"),
    # BASELINE.json configs[0]: isolation-model.py on examples/example_data.fa (hg18 vs pantro2, 65 255 sites), ONE
    # Likelihood(theta) call at a time -- the latency a drop-in user of the scripts sees (scripts/isolation-model.py:82-100)
    "c1": dict(model="isolation_k10", ctor=("IsolationModel", (10,)), default=[1e-3, 2000.0, 0.4], K=10,
               chunks=1, chunk_len=65255, points=1,
               desc="configs[0]: isolation model K=10 on the reference's example alignment (hg18 vs pantro2, 65 255 sites), "
                    "one Likelihood(theta) call at a time"),
}
SECONDARY = ("c3_1gpu", "ns_1gpu", "c4_1gpu", "c5_1gpu")


def thetas_around(default, n, seed=7, scale=0.1):
    """theta_b = default * exp(0.1 N(0,1)), PCG64(seed=7): the MCMC proposal scale (mcmc.py:26,34-36)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    default = np.asarray(default, dtype=np.float64)
    out = default[None, :] * np.exp(scale * rng.standard_normal((n, default.size)))
    out[0] = default
    return np.ascontiguousarray(out)


def flops_per_site_point(K):
    return 2 * K * K + 3 * K


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


# ---------------------------------------------------------------------------------------- synthetic data
def simulate_chunk(rng, pi, T, E, L, missing=0.04, mean_run=100):
    """Exact sample from the HMM (hidden path from (pi,T), symbols from E[:, :2]) + missing-data runs of
    symbol 2 (geometric length, mean 100, ~4% coverage) -- SURVEY 8(d).  Jump-chain simulation: O(#jumps)."""
    K = pi.size
    stay = np.clip(np.diag(T), 0.0, 1.0 - 1e-15)
    jump = T.copy()
    np.fill_diagonal(jump, 0.0)
    jump /= jump.sum(axis=1, keepdims=True)
    cum = np.cumsum(jump, axis=1)
    states, holds = [], []
    s = int(rng.choice(K, p=pi / pi.sum()))
    total = 0
    while total < L:
        h = int(rng.geometric(1.0 - stay[s]))
        states.append(s)
        holds.append(h)
        total += h
        s = min(int(np.searchsorted(cum[s], rng.random())), K - 1)
    path = np.repeat(np.asarray(states), np.asarray(holds))[:L]
    p1 = (E[:, 1] / (E[:, 0] + E[:, 1]))[path]
    obs = (rng.random(L) < p1).astype(np.uint8)
    n_runs = int(L * missing / mean_run * 1.5) + 8
    gaps = rng.geometric(missing / (mean_run * (1.0 - missing)), size=n_runs)
    runs = rng.geometric(1.0 / mean_run, size=n_runs)
    ends = np.cumsum(gaps + runs)
    starts = ends - runs
    keep = starts < L
    delta = np.zeros(L + 1, dtype=np.int32)
    np.add.at(delta, starts[keep], 1)
    np.add.at(delta, np.minimum(ends[keep], L), -1)
    obs[np.cumsum(delta[:L]) > 0] = 2
    return obs


def load_points(model, n_points):
    """(pi, T, E) built by the REFERENCE's own build_hidden_markov_model: the committed fixture
    tests/golden/model_<name>.npz (16 points theta_b = default * exp(0.1 N(0,1)) = the first 16 points of
    thetas_around) tiled to n_points.  Used to simulate the alignments, by the CPU arm, and by the parity gate."""
    g = np.load(os.path.join(ROOT, "tests", "golden", "model_%s.npz" % model))
    reps = (n_points + g["pi"].shape[0] - 1) // g["pi"].shape[0]
    pis = np.ascontiguousarray(np.tile(g["pi"], (reps, 1))[:n_points])
    Ts = np.ascontiguousarray(np.tile(g["T"], (reps, 1, 1))[:n_points])
    Es = np.ascontiguousarray(np.tile(g["E"], (reps, 1, 1))[:n_points])
    return pis, Ts, Es


def _sim_task(task):
    model, chunk_len, cid = task
    pis, Ts, Es = load_points(model, 1)
    rng = np.random.Generator(np.random.PCG64(SEED0 + int(cid)))
    return simulate_chunk(rng, pis[0], Ts[0], Es[0], chunk_len)


class ChunkFactory(object):
    """Synthetic alignment chunks, simulated from the reference-built model at the scripts' default parameters (point 0 of
    the fixture), chunk c with PCG64(SEED0 + c).  A pool of forked workers (created before CUDA is touched) spreads the
    simulation over the host cores: the K=40 workload alone is a minute of single-core Python otherwise."""

    def __init__(self, nworkers):
        self.pool = None
        if nworkers > 1:
            import multiprocessing as mp
            try:
                self.pool = mp.get_context("fork").Pool(nworkers)
            except Exception:
                self.pool = None

    def make(self, wl, chunk_ids):
        if wl.get("example"):
            return [np.load(os.path.join(ROOT, "tests", "golden", "example_pair.npz"))["symbols"]]
        tasks = [(wl["model"], wl["chunk_len"], int(c)) for c in chunk_ids]
        if self.pool is not None and len(tasks) > 4:
            return self.pool.map(_sim_task, tasks, chunksize=max(1, len(tasks) // 64))
        return [_sim_task(t) for t in tasks]

    def close(self):
        if self.pool is not None:
            self.pool.terminate()
            self.pool = None


def make_chunks(wl, pis, Ts, Es, chunk_ids, **sim_args):
    """Kept for tools/: chunks simulated from the given point-0 model."""
    out = []
    for cid in chunk_ids:
        rng = np.random.Generator(np.random.PCG64(SEED0 + int(cid)))
        out.append(simulate_chunk(rng, pis[0], Ts[0], Es[0], wl["chunk_len"], **sim_args))
    return out


# ---------------------------------------------------------------------------------------- clocks sampler
class ClockSampler(threading.Thread):
    """SM clock and throttle reasons of GPU `index` sampled from the warm-up steps to the end of the timed region: NVML
    (what nvidia-smi reads) every ~5 ms in-process, or an `nvidia-smi --query-gpu` call every ~50 ms where pynvml is
    not usable."""
    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self._stop_evt = index, threading.Event()
        self.sm, self.mx, self.reasons, self.source = [], [], set(), "nvidia-smi"
        self._nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            pynvml.nvmlDeviceGetClockInfo(self._handle, pynvml.NVML_CLOCK_SM)
            self._nvml, self.source = pynvml, "nvml"
        except Exception:
            self._nvml = None

    def _sample_nvml(self):
        nv, h = self._nvml, self._handle
        self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
        self.mx.append(float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)))
        bits = int(nv.nvmlDeviceGetCurrentClocksEventReasons(h))
        for name, bit in (("hw_slowdown", nv.nvmlClocksEventReasonHwSlowdown),
                          ("hw_thermal_slowdown", nv.nvmlClocksEventReasonHwThermalSlowdown),
                          ("sw_thermal_slowdown", nv.nvmlClocksEventReasonSwThermalSlowdown),
                          ("sw_power_cap", nv.nvmlClocksEventReasonSwPowerCap)):
            if bits & int(bit):
                self.reasons.add(name)

    def _sample_smi(self):
        out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.QUERY,
                              "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
        f = [x.strip() for x in out.strip().split(",")]
        if len(f) >= 7:
            if f[0].replace(".", "").isdigit():
                self.sm.append(float(f[0]))
            if f[1].replace(".", "").isdigit():
                self.mx.append(float(f[1]))
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    self.reasons.add(name)

    def run(self):
        while not self._stop_evt.is_set():
            try:
                if self._nvml is not None:
                    self._sample_nvml()
                else:
                    self._sample_smi()
            except Exception:
                if self._nvml is not None:
                    self._nvml, self.source = None, "nvidia-smi"     # fall back for the rest of the run
            self._stop_evt.wait(0.005 if self._nvml is not None else 0.05)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=6)
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": max(self.mx) if self.mx else None,
                "reasons": sorted(self.reasons), "samples": len(self.sm), "source": self.source}


# ---------------------------------------------------------------------------------------- CPU oracle legs
def oracle_zip(chunks, max_syms=256):
    """zipHMM-style preprocessing of every chunk with the oracle (hmm.py:16), spread over the host cores."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import forward as F
    F.build()
    with ThreadPoolExecutor(max(1, min(32, host_cores()))) as ex:
        return list(ex.map(lambda c: F.zip_preprocess(c.astype(np.int32), 3, max_syms=max_syms), chunks))


def parity_check(chunks, logl, pis_ref, Ts_ref, Es_ref, pis_gpu, Ts_gpu, Es_gpu, kernel, budget_s, label_full):
    """Every output of the timed launch against the oracle's zipHMM-style forward on the same chunks: the sum over ALL
    chunks for every point, or -- when the CPU cannot do that inside budget_s -- for an evenly spaced subset of the
    points (stated in `sample`).  Points < 16 use the reference-built (pi,T,E) of the fixture."""
    from oracle import forward as F
    N = logl.size
    t0 = time.perf_counter()
    zipped = oracle_zip(chunks, max_syms=128 if pis_gpu.shape[1] >= 32 else 256)
    t_prep = time.perf_counter() - t0
    nref = min(16, N, pis_ref.shape[0])
    pis, Ts, Es = pis_gpu.copy(), Ts_gpu.copy(), Es_gpu.copy()
    pis[:nref], Ts[:nref], Es[:nref] = pis_ref[:nref], Ts_ref[:nref], Es_ref[:nref]
    # calibrate on 2 points, then as many evenly spaced points as the budget allows (always including the first 16)
    t0 = time.perf_counter()
    F.forward_batch(None, pis[:2], Ts[:2], Es[:2], mode="zip_fast", zipped=zipped, nthreads=host_cores())
    per_point = max((time.perf_counter() - t0) / 2, 1e-6)
    can = int(max(budget_s - t_prep, 1.0) / per_point)
    if can >= N:
        idx = np.arange(N)
    else:
        extra = max(0, can - nref)
        idx = np.unique(np.concatenate([np.arange(nref), np.linspace(nref, N - 1, max(extra, 1)).astype(int)])) if N > nref else np.arange(nref)
    t0 = time.perf_counter()
    want, _ = F.forward_batch(None, pis[idx], Ts[idx], Es[idx], mode="zip_fast", zipped=zipped, nthreads=host_cores())
    t_cpu = time.perf_counter() - t0
    rel = np.abs(logl[idx] - want) / np.abs(want)
    worst = int(idx[int(np.argmax(rel))])
    sites = sum(len(c) for c in chunks)
    return {"max_rel_err": float(rel.max()), "tolerance": 1e-9, "ok": bool(rel.max() <= 1e-9), "kernel": kernel,
            "checked_outputs": int(idx.size), "outputs": int(N), "worst_point": worst,
            "max_rel_err_reference_built_points": float(rel[:nref].max()) if nref else None,
            "sample": ("full: " if idx.size == N else "subset: ") +
                      "%d of the %d outputs of the timed launch (%s), each the sum over all %d chunks (%d sites), vs the oracle's "
                      "zipHMM-style float64 forward; points 0..%d with the reference-built (pi,T,E) of tests/golden, the rest "
                      "with the GPU-built ones" % (idx.size, N, label_full, len(chunks), sites, nref - 1),
            "oracle_seconds": round(t_prep + t_cpu, 2)}


def cpu_reference_run(wl, chunk_factory, steps, warmup, budget_s=12.0):
    """The reference's CPU path restated (oracle, zipHMM-style pair compression + forward, OpenMP over
    (point, chunk) tasks on all host cores).  Each step scores a bounded sample of the workload."""
    from oracle import forward as F
    F.build()
    ncores = host_cores()
    pis, Ts, Es = load_points(wl["model"], wl["points"])
    # bounded sample: n_c chunks x n_p points sized from a calibration run
    n_c = min(wl["chunks"], max(ncores, 8))
    n_p = min(wl["points"], 16)
    chunks = chunk_factory.make(wl, range(n_c))
    # The dictionary size trades K^3 work per new symbol (rebuilt in every zip_forward call, hmm.py:20-21) against
    # K^2 work per remaining symbol; mini-ziphmm tunes it from its own cost estimate.  Give the CPU arm the best of a
    # small sweep so that it is not handicapped by our choice.
    best = None
    for max_syms in (64, 128, 256, 512, 1024):
        t0 = time.perf_counter()
        z = oracle_zip(chunks, max_syms)
        tp = time.perf_counter() - t0
        t0 = time.perf_counter()
        F.forward_batch(None, pis[:n_p], Ts[:n_p], Es[:n_p], mode="zip_fast", zipped=z, nthreads=ncores)
        tc = time.perf_counter() - t0
        if best is None or tc < best[0]:
            best = (tc, tp, z, max_syms)
    t_cal, t_prep, zipped, max_syms = best
    per_step = budget_s / max(1, steps + warmup)
    scale = max(1, min(wl["points"] // n_p, int(per_step / max(t_cal, 1e-4))))
    n_p = min(wl["points"], n_p * scale)
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        _, used = F.forward_batch(None, pis[:n_p], Ts[:n_p], Es[:n_p], mode="zip_fast", zipped=zipped, nthreads=ncores)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    sites = sum(len(c) for c in chunks)
    t = float(np.sum(times))
    value = sites * n_p * len(times) / t
    ratio = float(np.mean([len(c) / max(1, len(z[0])) for c, z in zip(chunks, zipped)]))
    # how the reference actually runs: one process, one thread (mcmc.py forks one such process per chain)
    n1 = max(1, min(n_p, 4))
    t0 = time.perf_counter()
    F.forward_batch(None, pis[:n1], Ts[:n1], Es[:n1], mode="zip_fast", zipped=zipped[:2], nthreads=1)
    single = sum(len(c) for c in chunks[:2]) * n1 / (time.perf_counter() - t0)
    return {"value": value, "unit": "sites*points/s", "cores": int(used), "kind": "port", "single_thread_value": single,
            "sample": "%d chunks x %d bp x %d points per step, zipHMM-style compressed forward (imco_zip_forward_fast: "
                      "column-major matrices, power-of-two rescaling every 8 symbols; per-chunk dictionaries of "
                      "<= %d symbols, the fastest of 64..1024 on this host; %.0fx fewer symbols, preprocess %.1fs excluded "
                      "like hmm.py:16), OpenMP, %d steps; (pi,T,E) from the committed fixture built by the reference's own "
                      "build_hidden_markov_model" % (n_c, wl["chunk_len"], n_p, max_syms, ratio, t_prep, len(times)),
            "ms_per_step": 1e3 * t / len(times)}


def base_config(wl):
    """The same dict for both arms (the driver compares them)."""
    return {"workload": wl["desc"], "chunks_per_gpu": wl["chunks"], "chunk_len": wl["chunk_len"],
            "points": wl["points"], "K": wl["K"], "sharding": "chunks across ranks, no data-path collective except one "
            "all-reduce of float64[points]", "l2": "flushed between timed steps (256 MiB write)"}


# ---------------------------------------------------------------------------------------- one workload on the GPU
class GpuCtx(object):
    def __init__(self, m, torch, dist, dev, local_rank, rank, world, lib_comm, flush, peaks):
        self.m, self.torch, self.dist, self.dev = m, torch, dist, dev
        self.local_rank, self.rank, self.world, self.lib_comm = local_rank, rank, world, lib_comm
        self.flush, self.peaks = flush, peaks


def measure_workload(g, name, wl, chunk_factory, steps, warmup, forward_kernel=0, parity_budget_s=20.0, e2e=True,
                     parity=True, collective="fused"):
    """Times one workload on this rank's GPU; returns (result dict for rank 0, extras)."""
    m, torch, dist, dev = g.m, g.torch, g.dist, g.dev
    K, N, S = wl["K"], wl["points"], 3
    m.set_option("forward_kernel", forward_kernel)
    model = getattr(m, wl["ctor"][0])(*wl["ctor"][1])
    thetas = thetas_around(wl["default"], N)
    # weak scaling: every rank owns wl["chunks"] chunks (distinct seeds), all ranks score the same points
    chunks = chunk_factory.make(wl, range(g.rank * wl["chunks"], (g.rank + 1) * wl["chunks"]))
    t0 = time.perf_counter()
    fset = m.ForwarderSet([m.Forwarder.from_symbols(c, 3) for c in chunks])
    t_preprocess = time.perf_counter() - t0      # one-off, like Forwarder.__init__ (hmm.py:12-16); not in any timed region
    sites_rank = fset.total_sites
    pis, Ts, Es, st = model.build_hidden_markov_models(thetas)       # host copies for the kernel-only leg and the parity gate
    assert (st == 0).all()
    d_theta = torch.tensor(thetas, device=dev)
    d_pi, d_T, d_E = (torch.tensor(x, device=dev) for x in (pis, Ts, Es))
    d_out = torch.empty(N, dtype=torch.float64, device=dev)
    stream = torch.cuda.current_stream()
    reduce_torch = dist is not None and not g.lib_comm

    def step():
        # theta -> (pi,T,E) -> logL on the device: model-build kernels, (spectral preparation,) forward, chunk reduction
        model.batched_log_likelihood_device(d_theta.data_ptr(), fset, d_out.data_ptr(), N, 0, stream.cuda_stream)
        if reduce_torch:
            dist.all_reduce(d_out)        # the only collective: float64[N] partial log-likelihoods (lib_comm: done inside the call)

    sampler = ClockSampler(g.local_rank)
    sampler.start()
    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    launches0 = m.kernel_launches()
    evs = []
    for _ in range(steps):
        g.flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        step()
        b.record(stream)
        evs.append((a, b))
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    clocks = sampler.stop()
    launches = m.kernel_launches() - launches0
    kernel = m.last_forward_kernel()
    t_dev = sum(a.elapsed_time(b) for a, b in evs) * 1e-3
    if dist is not None:
        tt = torch.tensor([t_dev], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_dev = float(tt.item())
    logl_dev = d_out.cpu().numpy().copy()        # the outputs of the LAST timed step: what the parity gate checks

    # ---- N > 1: the fused all-reduce against torch.distributed on this run's own partial sums ----
    collective_check = None
    if dist is not None and g.lib_comm:
        m.set_option("comm_enabled", 0)          # this rank's partial sums
        part = torch.empty(N, dtype=torch.float64, device=dev)
        model.batched_log_likelihood_device(d_theta.data_ptr(), fset, part.data_ptr(), N, 0, stream.cuda_stream)
        torch.cuda.synchronize()
        m.set_option("comm_enabled", 1)
        gathered = [torch.empty_like(part) for _ in range(g.world)]
        dist.all_gather(gathered, part)
        ordered = torch.zeros_like(part)
        for p in gathered:                       # rank order: the order reduce_chains_peer_kernel adds the rows in
            ordered += p
        nccl_sum = part.clone()
        dist.all_reduce(nccl_sum)
        fused = torch.tensor(logl_dev, device=dev)
        collective_check = {"collective": collective, "fused_in_kernel": bool(m._lib.comm_info()["fused"]),
                            "bit_equal_to_rank_ordered_sum": bool(torch.equal(fused, ordered)),
                            "max_rel_diff_vs_torch_all_reduce": float(((fused - nccl_sum).abs() / nccl_sum.abs()).max().item()),
                            "ranks": g.world}
        ok = torch.tensor([1 if collective_check["bit_equal_to_rank_ordered_sum"] else 0], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        collective_check["all_ranks_agree"] = bool(ok.item())

    # ---- kernel-only time of the dominant kernel (forward kernel without model build / all-reduce) ----
    kt = []
    m.set_option("comm_enabled", 0)          # this rank's kernel alone
    passes0 = m.mma_passes()
    for _ in range(min(3, steps)):
        g.flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        fset.forward_batch_device(d_pi.data_ptr(), d_T.data_ptr(), d_E.data_ptr(), d_out.data_ptr(), N, K, S,
                                  stream.cuda_stream)
        b.record(stream)
        torch.cuda.synchronize()
        kt.append(a.elapsed_time(b) * 1e-3)
    passes_per_launch = (m.mma_passes() - passes0) / float(len(kt))
    t_kernel = float(np.mean(kt))
    m.set_option("comm_enabled", 1)

    # ---- end to end through the host API (Model.batched_log_likelihood -> imc_loglik_batch): pinned host theta in,
    # host logL + status out, every step; model build and forward on the device in between ----
    t_e2e = None
    if e2e:
        h_theta = torch.tensor(thetas).pin_memory()
        np_theta = h_theta.numpy()
        for _ in range(2):
            np_out = model.batched_log_likelihood(np_theta, fset)
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            np_out = model.batched_log_likelihood(np_theta, fset)   # synchronous: returns with logL on the host
            if reduce_torch:
                tmp = torch.from_numpy(np_out).to(dev)
                dist.all_reduce(tmp)
                np_out[:] = tmp.cpu().numpy()
        t_e2e = time.perf_counter() - t0
        if dist is not None:
            tt = torch.tensor([t_e2e], dtype=torch.float64, device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            t_e2e = float(tt.item())
        if g.world == 1:
            assert np.allclose(np_out, logl_dev, rtol=1e-12), "host API and device API disagree"
    if g.rank != 0:
        return None

    spectral = kernel.startswith("zip-spectral")
    zinfo = fset.run_info(K) if spectral else fset.zip_info(K)
    total_site_points = float(sites_rank) * N * g.world
    peak_dfma, peak_dmma = g.peaks
    peak = max(peak_dfma, peak_dmma)
    algo_flops = float(sites_rank) * N * flops_per_site_point(K)          # per launch (one rank), SURVEY 8(d)
    achieved = algo_flops / t_kernel / 1e12
    kname = "imc::%s_kernel" % ("zip_forward" if kernel.startswith("zip") else "fwd_" + kernel)
    fp64_view = {"bound": "fp64", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                 "peak_source": "measured in this run (imc_measure_fp64_peak): DFMA %.1f, DMMA %.1f TFLOP/s; "
                 "MEASURED_PEAKS.json has no FP64 entry" % (peak_dfma, peak_dmma),
                 "algorithmic_flop_per_site_point": flops_per_site_point(K)}
    if kernel.startswith("zip"):
        # The kernel runs the reference's own algorithm (zipHMM: one mat-vec per COMPRESSED symbol; in the spectral form one
        # per non-run dictionary entry).  Its necessary work per launch is tokens x points chain-steps = K x K mat-vecs.
        sm_clock = (clocks.get("sm_mhz") or 1965.0) * 1e6
        steps_exec = float(zinfo["tokens"]) * N
        smem_peak = 128.0 * 148 * sm_clock / 1e9
        smem_ach = steps_exec * 8.0 * K * K / t_kernel / 1e9
        exec_tflops = steps_exec * (2 * K * K + K) / t_kernel / 1e12
        compression = {"sites": int(sites_rank), "tokens": int(zinfo["tokens"]), "ratio": sites_rank / max(1, zinfo["tokens"]),
                       "form": "run tokens (spectral)" if spectral else "pair dictionary",
                       "dictionary_ids_used": zinfo["ids_used"], "dictionary_ids_available": zinfo["ids_available"],
                       "dictionary_levels": zinfo["levels"], "preprocess_s_one_off": t_preprocess}
        smem_view = {"bound": "smem", "achieved": smem_ach, "peak": smem_peak, "unit": "GB/s", "frac": smem_ach / smem_peak,
                     "frac_of_measured_pipe": smem_ach / (smem_peak * 122.0 / 128.0),
                     "note": "chain-steps x 8 K^2 bytes / kernel time against 128 B/clk/SM (122 measured with full LDS.128, "
                             "profiles/r01_smem_patterns.txt): the roofline of the FMA shapes, which stream one K x K matrix per "
                             "chain-step out of shared memory"}
        if "mma" in kernel:
            # MMA shape: the hot entry's matrix is in registers (B fragments), chains are rows of mma.sync.m8n8k4.f64 tiles; the pipe
            # that bounds it is the FP64 tensor (DMMA) pipe.  achieved = algorithmic flops (chain-steps x (2 K^2 + K)); the
            # executed DMMA flops (incl. tile padding and the rows of cold passes that serve other chains) are reported beside.
            KT, NT = (K + 3) // 4, (K + 7) // 8
            dmma_tflops = passes_per_launch * KT * NT * 512.0 / t_kernel / 1e12
            roofline = {
                "bound": "fp64-tensor", "kernel": kname + " (spectral form, MMA shape%s)" % (", aligned streams" if "aligned" in kernel else ""),
                "kernel_form": kernel, "achieved": exec_tflops, "peak": peak_dmma,
                "unit": "TFLOP/s", "frac": exec_tflops / peak_dmma, "traffic": None, "kernel_ms": 1e3 * t_kernel,
                "bound_note": "FP64 tensor pipe: achieved = chain-steps x (2 K^2 + K) flop / kernel time; peak = DMMA.8x8x4 rate "
                              "measured in this run (imc_measure_fp64_peak)",
                "algorithmic_flop_per_chain_step": 2 * K * K + K, "chain_steps_per_launch": steps_exec,
                "executed_dmma_tflops": dmma_tflops, "executed_dmma_frac_of_peak": dmma_tflops / peak_dmma,
                "dmma_passes_per_launch": passes_per_launch, "dmma_per_pass": KT * NT,
                "passes_per_warp_step": passes_per_launch / max(1.0, steps_exec / 8.0),
                "frac_note": "frac counts ALGORITHMIC flops only; the tensor pipe itself is busy executed_dmma_frac_of_peak of the time "
                             "(ncu sm__inst_executed_pipe_tensor_subpipe_dmma agrees: profiles/r02_zip*_mma_ncu.txt)",
                "executed_note": "every pass is KT x NT DMMAs of 512 flop for the 8 chains of a warp: tiles are padded to 8 x 4 "
                                 "(K=10: 100 of 192 MACs useful) and a pass serves only the chains whose token is its entry "
                                 "(lock step: hot pass + one per distinct cold entry; aligned: one pass per step of the warp's schedule)",
                "smem_view_of_the_fma_shapes": smem_view, "compression": compression, "plain_forward_equivalent": fp64_view}
        else:
            roofline = dict(smem_view, kernel=kname + (" (spectral form: run tokens)" if spectral else ""), traffic=None,
                            kernel_ms=1e3 * t_kernel, algorithmic_bytes_per_chain_step=8 * K * K, chain_steps_per_launch=steps_exec,
                            executed_tflops_fp64=exec_tflops, executed_frac_of_fp64_peak=exec_tflops / peak,
                            compression=compression, plain_forward_equivalent=fp64_view)
            roofline["bound_note"] = roofline.pop("note")
    else:
        roofline = dict(fp64_view, kernel=kname, traffic=None, kernel_ms=1e3 * t_kernel)
    try:    # DRAM traffic of the dominant kernel per launch, from the committed ncu capture of this workload and kernel form
        tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(name)
        if tr and tr.get("kernel_form") == kernel and forward_kernel == 0:
            roofline["traffic"] = tr["bytes"]
            roofline["traffic_source"] = "%s, ncu --set full: %s" % (tr["kernel"], tr["source"])
    except (OSError, ValueError):
        pass
    try:
        mp = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        hbm_peak, hbm_src = float(mp["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
    except (OSError, ValueError, KeyError):
        hbm_peak, hbm_src = 6650.0, "of fallback (B200_PROFILING.md: 6.65 TB/s)"
    if roofline.get("traffic"):
        ach = roofline["traffic"] / t_kernel / 1e9
        roofline["hbm_view"] = {"bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak,
                                "peak_source": hbm_src, "note": "not the bound: the working set lives in shared memory and registers"}
    res = {"value": total_site_points * steps / t_dev, "unit": "sites*points/s", "ms_per_step": 1e3 * t_dev / steps,
           "steps": steps, "warmup": warmup, "clocks": clocks, "gpu_launches": int(launches), "kernel": kernel,
           "roofline": roofline,
           "logL_check": {"first": float(logl_dev[0]), "finite": bool(np.isfinite(logl_dev).all())}}
    if t_e2e is not None:
        res["e2e"] = {"value": total_site_points * steps / t_e2e, "unit": "sites*points/s",
                      "h2d_bytes_per_step": int(thetas.nbytes), "d2h_bytes_per_step": int(N * 8 + N * 4)}
    if collective_check is not None:
        res["collective_check"] = collective_check
    if parity and g.world == 1:
        pr, Tr, Er = load_points(wl["model"], min(16, N))
        res["parity"] = parity_check(chunks, logl_dev, pr, Tr, Er, pis, Ts, Es, kernel, parity_budget_s,
                                     "%d chunks x %d points" % (len(chunks), N))
    elif parity:
        # N > 1: the timed outputs are sums over all ranks' shards; rank 0 checks them against the collective check above
        # and its own shard's partial sums against the oracle
        m.set_option("comm_enabled", 0)
        part = model.batched_log_likelihood(thetas, fset)
        m.set_option("comm_enabled", 1)
        pr, Tr, Er = load_points(wl["model"], min(16, N))
        res["parity"] = parity_check(chunks, part, pr, Tr, Er, pis, Ts, Es, kernel, parity_budget_s,
                                     "rank 0's partial sums, %d chunks x %d points" % (len(chunks), N))
    return res


def measure_latency(g, chunk_factory, calls=300):
    """configs[0]: one Likelihood(theta) call at a time through the reference-facing Python API
    (scripts/isolation-model.py:82-100), then a 200-evaluation Nelder-Mead run (likelihood.py:36-87)."""
    m = g.m
    wl = dict(WORKLOADS["c1"], example=True)
    obs = chunk_factory.make(wl, [0])[0]
    f = m.Forwarder.from_symbols(obs, 3)
    model = m.IsolationModel(10)
    like = m.Likelihood(model, [f])
    theta = np.array(wl["default"])
    for _ in range(5):
        val = like(theta)
    ts = []
    rng = np.random.default_rng(1)
    for _ in range(calls):
        th = theta * np.exp(0.05 * rng.standard_normal(3))
        t0 = time.perf_counter()
        like(th)
        ts.append(time.perf_counter() - t0)
    kernel = m.last_forward_kernel()
    ts = np.sort(np.asarray(ts)) * 1e3
    evals = [0]

    def counted(th):
        evals[0] += 1
        return like(th)
    import scipy.optimize
    t0 = time.perf_counter()
    res = scipy.optimize.minimize(lambda p: -counted(np.asarray(p)), theta * 1.3, method="Nelder-Mead",
                                  options={"maxfev": 200, "disp": False})
    t_opt = time.perf_counter() - t0
    # CPU: the oracle's zipHMM-style forward, one thread, the way the reference runs (one process per chain)
    from oracle import forward as F
    pis, Ts, Es = load_points(wl["model"], 1)
    z = F.zip_preprocess(obs.astype(np.int32), 3, max_syms=256)
    F.zip_forward_fast(pis[0], Ts[0], Es[0], z[1], z[0], 3, z[2])
    t0 = time.perf_counter()
    for _ in range(20):
        want = F.zip_forward_fast(pis[0], Ts[0], Es[0], z[1], z[0], 3, z[2])
    t_cpu = (time.perf_counter() - t0) / 20
    return {"workload": wl["desc"], "calls": calls, "p50_ms": float(ts[len(ts) // 2]), "p99_ms": float(ts[int(len(ts) * 0.99)]),
            "min_ms": float(ts[0]), "kernel": kernel, "api": "imcoalhmm_b200.Likelihood(IsolationModel(10), [Forwarder])(theta): "
            "model build + forward fused on the device, host theta in, python float out",
            "nelder_mead": {"evaluations": evals[0], "seconds": t_opt, "ms_per_evaluation": 1e3 * t_opt / max(1, evals[0]),
                            "logL_at_optimum": float(-res.fun)},
            "parity": {"logL": float(val), "oracle": float(want), "rel_err": float(abs(val - want) / abs(want)),
                       "note": "theta = script defaults; oracle on the reference-built (pi,T,E) of tests/golden"},
            "cpu_forward_only_ms_single_thread": 1e3 * t_cpu,
            "cpu_note": "oracle zipHMM-style forward alone (no model build; the reference's Python build_hidden_markov_model "
                        "adds 2.4 ms per call for this model, SURVEY 3.1)"}


# ---------------------------------------------------------------------------------------- main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the oracle parity gate (profiling runs)")
    ap.add_argument("--no-secondary", action="store_true", help="skip the secondary workloads of the default run")
    ap.add_argument("--secondary-budget-s", type=float, default=110.0,
                    help="CPU seconds the oracle may spend on the parity gates of the secondary workloads (all together)")
    ap.add_argument("--collective", default="fused", choices=["fused", "nccl", "torch"],
                    help="N > 1: all-reduce inside the library's reduction kernel over peer memory (falls back to its "
                         "ncclAllReduce where the mailboxes cannot be mapped), the library's ncclAllReduce, or "
                         "torch.distributed.all_reduce on the partial sums")
    ap.add_argument("--forward-kernel", type=int, default=0,
                    help="0 auto (zip: compressed token streams), 2 lane-pair DFMA, 3 DMMA (2/3 walk every site)")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    config = base_config(wl)
    # forked before anything touches CUDA; the workers only ever run numpy
    chunk_factory = ChunkFactory(max(1, min(16, host_cores() // max(1, world))))

    if args.impl == "reference":
        if rank != 0:
            chunk_factory.close()
            return
        res = cpu_reference_run(wl, chunk_factory, args.steps, max(args.warmup, 1))
        chunk_factory.close()
        line = {"metric": "forward sites*param-points/sec", "value": res["value"], "unit": "sites*points/s",
                "impl": "reference", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic", "config": config, "gpu_launches": 0,
                "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample", "single_thread_value")},
                "e2e": {"value": res["value"], "unit": "sites*points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "notes": {"model_build": "not part of this arm: (pi,T,E) come from the committed fixture built by the reference's "
                          "own Python build_hidden_markov_model (2.4 ms/theta for the isolation model on one core, SURVEY 3.1)",
                          "kind": "port: ziphmm (birc-aeh/mini-ziphmm) is absent from /root/reference; this is the oracle's "
                                  "restatement of its published algorithm"}}
        print(json.dumps(line))
        return

    import torch
    import imcoalhmm_b200 as m
    assert args.warmup >= 3 or args.steps <= 2, "timing rules: at least 3 warm-up steps"
    dist = None
    torch.cuda.set_device(local_rank)
    if world > 1:
        import torch.distributed as dist
        # NCCL announces its version on stdout when the first communicator is created: keep stdout for the JSON line
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            dist.all_reduce(torch.zeros(1, device=torch.device("cuda", local_rank)))
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    m._lib.check(m._lib.load().imc_init(local_rank))
    dev = torch.device("cuda", local_rank)
    lib_comm = False
    collective_note = "none (one rank)"
    if world > 1 and args.collective != "torch":
        # the library's own communicator: rank 0 draws the id, torch.distributed carries it to the other ranks
        from imcoalhmm_b200.sharding import init_library_comm
        init_library_comm(dev, fused=(args.collective == "fused"))
        lib_comm = True
        collective_note = ("all-reduce fused into the chain-reduction kernel (peer-to-peer stores over NVLink)"
                           if m._lib.comm_info()["fused"] else "ncclAllReduce issued by the library")
    elif world > 1:
        collective_note = "torch.distributed.all_reduce (NCCL)"

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    peaks = m.measure_fp64_peak()
    g = GpuCtx(m, torch, dist, dev, local_rank, rank, world, lib_comm, flush, peaks)

    if args.workload == "c1":
        lat = measure_latency(g, chunk_factory)
        chunk_factory.close()
        sites = WORKLOADS["c1"]["chunk_len"]
        line = {"metric": "forward sites*param-points/sec", "value": sites / (lat["p50_ms"] * 1e-3), "unit": "sites*points/s",
                "n_gpus": 1, "steps": lat["calls"], "warmup": 5, "ms_per_step": lat["p50_ms"], "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "reference example alignment (tests/golden/example_pair.npz)",
                "config": config, "latency": lat, "gpu_launches": int(m.kernel_launches())}
        print(json.dumps(line))
        return

    res = measure_workload(g, args.workload, wl, chunk_factory, args.steps, args.warmup, args.forward_kernel,
                           parity_budget_s=25.0, parity=not args.no_parity, collective=args.collective)
    if rank != 0:
        if lib_comm:
            m._lib.comm_destroy()
        chunk_factory.close()
        if dist is not None:
            dist.destroy_process_group()
        return
    line = {
        "metric": "forward sites*param-points/sec", "value": res["value"], "unit": "sites*points/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": res["ms_per_step"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config, "clocks": res["clocks"], "gpu_launches": res["gpu_launches"], "e2e": res["e2e"],
        "roofline": res["roofline"], "logL_check": res["logL_check"],
        "notes": {"model_build": "inside the timed region, on the GPU (theta -> pi,T,E -> logL fused on the device)",
                  "collective": collective_note,
                  "oracle": "forward parity is against the repository's CPU restatement of zipHMM (oracle/): the reference's "
                            "ziphmm dependency is absent and its tests pin no log-likelihood, so parity at that boundary is "
                            "UNPINNED by the reference; the model-build half is pinned by the reference's own code"},
    }
    if "parity" in res:
        line["parity"] = res["parity"]
    if "collective_check" in res:
        line["collective_check"] = res["collective_check"]
    if lib_comm:
        m._lib.comm_destroy()      # collective; rank 0 goes on alone
    failures = []
    if "parity" in res and not res["parity"]["ok"]:
        failures.append("primary parity %g" % res["parity"]["max_rel_err"])
    if "collective_check" in res and not res["collective_check"]["all_ranks_agree"]:
        failures.append("fused all-reduce differs from the rank-ordered sum")

    # ---- secondary block: the shapes the north star names + the drop-in latency, measured in the same run ----
    if world == 1 and args.workload == "c2" and not args.no_secondary and args.forward_kernel == 0:
        sec = {}
        share = args.secondary_budget_s / len(SECONDARY)
        for name in SECONDARY:
            try:
                r = measure_workload(g, name, WORKLOADS[name], chunk_factory, steps=3 if name == "c4_1gpu" else 5, warmup=3, parity_budget_s=share,
                                     e2e=False, parity=not args.no_parity)
                rf = r["roofline"]
                sec[name] = {"workload": WORKLOADS[name]["desc"], "ms_per_step": r["ms_per_step"], "value": r["value"],
                             "unit": r["unit"], "steps": r["steps"], "warmup": r["warmup"], "kernel": r["kernel"],
                             "kernel_ms": rf["kernel_ms"], "roofline_bound": rf.get("bound"), "roofline_frac": rf.get("frac"),
                             "executed_dmma_frac_of_peak": rf.get("executed_dmma_frac_of_peak"),
                             "passes_per_warp_step": rf.get("passes_per_warp_step"),
                             "smem_view_frac": (rf.get("smem_view_of_the_fma_shapes") or rf).get("frac"),
                             "plain_forward_equivalent_tflops": rf["plain_forward_equivalent"]["achieved"],
                             "compression": rf.get("compression"), "clocks": r["clocks"], "gpu_launches": r["gpu_launches"],
                             "parity": r.get("parity")}
                if r.get("parity") and not r["parity"]["ok"]:
                    failures.append("%s parity %g" % (name, r["parity"]["max_rel_err"]))
            except Exception as e:      # a secondary workload must not cost the run its primary line
                sec[name] = {"error": "%s: %s" % (type(e).__name__, e)}
        try:
            sec["c1"] = measure_latency(g, chunk_factory)
        except Exception as e:
            sec["c1"] = {"error": "%s: %s" % (type(e).__name__, e)}
        line["secondary"] = sec
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = {k: v for k, v in cpu_reference_run(wl, chunk_factory, 3, 1).items() if k != "ms_per_step"}
    chunk_factory.close()
    if failures:
        line["failed"] = failures
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    assert not failures, "; ".join(failures)


if __name__ == "__main__":
    main()
