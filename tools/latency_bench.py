"""Latency of ONE log-likelihood evaluation through the reference-facing API (Likelihood.__call__ pattern:
one theta per call, host arrays in, python float out) -- the way the reference's optimiser and MCMC drive it."""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import imcoalhmm_b200 as m  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--chunks", type=int, default=0)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--chunk-len", type=int, default=0)
    args = ap.parse_args()
    wl = dict(bench.WORKLOADS[args.workload])
    if args.chunks:
        wl["chunks"] = args.chunks
    if args.chunk_len:
        wl["chunk_len"] = args.chunk_len
    model = getattr(m, wl["ctor"][0])(*wl["ctor"][1])
    theta = np.asarray(wl["default"], dtype=np.float64)
    pi, T, E = model.build_hidden_markov_model(theta)
    chunks = bench.make_chunks(wl, pi[None], T[None], E[None], range(wl["chunks"]))
    forwarders = [m.Forwarder.from_symbols(c, 3) for c in chunks]
    lik = m.Likelihood(model, forwarders)
    for seg, label in ((-1, "sequential chains"), (0, "auto (segmented when chain-scarce)")):
        m.set_option("zip_segment_tokens", seg)
        v = lik(theta)
        t0 = time.perf_counter()
        for _ in range(args.reps):
            v = lik(theta)
        dt = (time.perf_counter() - t0) / args.reps
        print("%-40s %d chunks x %d bp, K=%d: %.3f ms per Likelihood(theta) call, logL=%.6f, kernel=%s"
              % (label, wl["chunks"], wl["chunk_len"], wl["K"], dt * 1e3, v, m.last_forward_kernel()), flush=True)


if __name__ == "__main__":
    main()
