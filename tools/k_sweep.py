"""Zip kernel across state counts and lane decompositions on a compressible synthetic alignment (development aid)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import imcoalhmm_b200 as m  # noqa: E402

rng = np.random.default_rng(0)
C, L, N = 96, 300_000, 192
chunks = []
for _ in range(C):
    s = (rng.random(L) < 0.003).astype(np.uint8)
    t = 0
    while t < L:
        t += int(rng.geometric(1.0 / 2500))
        run = int(rng.geometric(1.0 / 100))
        s[t:t + run] = 2
        t += run
    chunks.append(s)
fset = m.ForwarderSet([m.Forwarder.from_symbols(c, 3) for c in chunks])
stream = torch.cuda.current_stream().cuda_stream
for K in (4, 6, 8, 10, 12, 14, 16, 20, 24, 28, 32, 40):
    pis = rng.dirichlet(np.ones(K), size=N)
    Ts = np.stack([0.9995 * np.eye(K) + 0.0005 * rng.dirichlet(np.ones(K), size=K) for _ in range(N)])
    Es = np.empty((N, K, 3))
    Es[:, :, 1] = rng.uniform(0.001, 0.01, size=(N, K))
    Es[:, :, 0] = 1.0 - Es[:, :, 1]
    Es[:, :, 2] = 1.0
    d_pi, d_T, d_E = (torch.tensor(x, device="cuda") for x in (pis, Ts, Es))
    d_out = torch.empty(N, dtype=torch.float64, device="cuda")
    res = []
    ref = None
    for lanes in (8, 4, 0):
        if lanes == 4 and K < 8:
            continue
        m.set_option("forward_kernel", 4)
        m.set_option("zip_lanes", lanes)
        info = fset.zip_info(K)
        best = 1e9
        for _ in range(3):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fset.forward_batch_device(d_pi.data_ptr(), d_T.data_ptr(), d_E.data_ptr(), d_out.data_ptr(), N, K, 3, stream)
            b.record()
            torch.cuda.synchronize()
            best = min(best, a.elapsed_time(b))
        out = d_out.cpu().numpy()
        ref = out if ref is None else ref
        steps = info["tokens"] * N
        res.append("lanes=%d M=%d %.3f ms (%.1f clk/chain-step, %.2f of smem peak) diff %.1e" % (
            lanes, info["ids_used"], best, best * 1e-3 * 148 * 1.965e9 / steps, steps * 8.0 * K * K / (best * 1e-3) / 37.2e12,
            float(np.max(np.abs(out - ref) / np.abs(ref)))))
    print("K=%2d  " % K + " | ".join(res), flush=True)
