"""A few single-theta forward calls with forced kernel options (for ncu): python tools/one_call.py c3_1gpu 32 -1"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import imcoalhmm_b200 as m  # noqa: E402

name, lanes, seg = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
wl = dict(bench.WORKLOADS[name])
model = getattr(m, wl["ctor"][0])(*wl["ctor"][1])
theta = np.asarray(wl["default"], dtype=np.float64)
pi, T, E = model.build_hidden_markov_model(theta)
chunks = bench.make_chunks(wl, pi[None], T[None], E[None], range(wl["chunks"]))
fset = m.ForwarderSet([m.Forwarder.from_symbols(c, 3) for c in chunks])
m.set_option("zip_lanes", lanes)
m.set_option("zip_segment_tokens", seg)
for _ in range(3):
    print(fset.forward(pi, T, E), m.last_forward_kernel())
