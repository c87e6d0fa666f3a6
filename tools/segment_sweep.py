"""Sweep zip_segment_tokens for single-theta calls (development aid for the cost model in forward_dev)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import imcoalhmm_b200 as m  # noqa: E402

for name, nchunks, clen in (("c2", 100, 0), ("c2", 1, 100_000_000), ("c3_1gpu", 125, 0), ("c5_1gpu", 100, 0)):
    wl = dict(bench.WORKLOADS[name])
    wl["chunks"] = nchunks
    if clen:
        wl["chunk_len"] = clen
    model = getattr(m, wl["ctor"][0])(*wl["ctor"][1])
    theta = np.asarray(wl["default"], dtype=np.float64)
    pi, T, E = model.build_hidden_markov_model(theta)
    chunks = bench.make_chunks(wl, pi[None], T[None], E[None], range(wl["chunks"]))
    fset = m.ForwarderSet([m.Forwarder.from_symbols(c, 3) for c in chunks])
    for lanes in (0, 32):
        m.set_option("zip_lanes", lanes)
        out = []
        for seg in ((-1, 0, 128, 256, 512, 768, 1024, 1536, 2048, 3072, 4096) if lanes == 0 else (-1, 1024, 2048, 4096, 8192)):
            m.set_option("zip_segment_tokens", seg)
            v = fset.forward(pi, T, E)
            t0 = time.perf_counter()
            for _ in range(10):
                fset.forward(pi, T, E)
            out.append("%d:%.3f" % (seg, (time.perf_counter() - t0) * 100))
        print(name, nchunks, wl["chunk_len"], "K=%d lanes=%d logL=%.6f" % (wl["K"], lanes, v), " ".join(out), flush=True)
