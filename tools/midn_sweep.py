"""Small and medium batches (N = 1..64 points): automatic choice vs forced shapes (development aid for the cost model)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import imcoalhmm_b200 as m  # noqa: E402

for name in ("c2", "c3_1gpu"):
    wl = dict(bench.WORKLOADS[name])
    model = getattr(m, wl["ctor"][0])(*wl["ctor"][1])
    thetas = bench.thetas_around(wl["default"], 64)
    pis, Ts, Es, st = model.build_hidden_markov_models(thetas)
    chunks = bench.make_chunks(wl, pis, Ts, Es, range(wl["chunks"]))
    fset = m.ForwarderSet([m.Forwarder.from_symbols(c, 3) for c in chunks])
    for N in (1, 2, 4, 8, 16, 32, 64):
        res = []
        for label, lanes, seg in (("auto", 0, 0), ("default-seq", 0, -1), ("warp", 32, -1), ("seg512", 0, 512), ("seg2048", 0, 2048)):
            m.set_option("zip_lanes", lanes)
            m.set_option("zip_segment_tokens", seg)
            fset.forward_batch(pis[:N], Ts[:N], Es[:N])
            t0 = time.perf_counter()
            for _ in range(5):
                fset.forward_batch(pis[:N], Ts[:N], Es[:N])
            res.append("%s %.3f (%s)" % (label, (time.perf_counter() - t0) * 200, m.last_forward_kernel()))
        print(name, "N=%d" % N, " | ".join(res), flush=True)
