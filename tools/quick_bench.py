"""Quick device-side timing of the forward kernels (development aid; bench.py is the contract)."""
import argparse
import ctypes
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import imcoalhmm_b200 as m  # noqa: E402


def flops_per_site(K):
    return 2 * K * K + 3 * K


def run(K, C, L, N, kernel, mt, reps, seed=0):
    rng = np.random.default_rng(seed)
    t0 = time.time()
    chunks = [rng.choice(3, size=L, p=[0.95, 0.01, 0.04]).astype(np.uint8) for _ in range(C)]
    fset = m.ForwarderSet([m.Forwarder.from_symbols(c, 3) for c in chunks])
    t_prep = time.time() - t0
    pis = rng.dirichlet(np.ones(K), size=N)
    Ts = np.stack([0.999 * np.eye(K) + 0.001 * rng.dirichlet(np.ones(K), size=K) for _ in range(N)])
    Es = rng.dirichlet(np.ones(3), size=(N, K))
    Es[:, :, 2] = 1.0
    d_pi, d_T, d_E = (torch.tensor(x, device="cuda") for x in (pis, Ts, Es))
    d_out = torch.empty(N, dtype=torch.float64, device="cuda")
    m.set_option("forward_kernel", kernel)
    m.set_option("dmma_mtiles", mt)
    stream = torch.cuda.current_stream().cuda_stream

    def step():
        fset.forward_batch_device(d_pi.data_ptr(), d_T.data_ptr(), d_E.data_ptr(), d_out.data_ptr(), N, K, 3, stream)

    step()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        step()
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b) * 1e-3)
    sp = C * L * N
    print("K=%d C=%d L=%d N=%d kernel=%s mt=%d: %.4f s  %.3e site-pts/s  %.2f TFLOP/s algorithmic (prep %.1fs) logL[0]=%.6f"
          % (K, C, L, N, m.last_forward_kernel(), mt, best, sp / best, sp * flops_per_site(K) / best / 1e12, t_prep,
             d_out[0].item()), flush=True)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", default="c2")
    ap.add_argument("--custom", action="append", default=[], help="K,C,L,N,kernel,mt,fold")
    args = ap.parse_args()
    for spec in args.custom:
        K, C, L, N, kern, mt, fold = map(int, spec.split(","))
        m.set_option("fold_emission", fold)
        run(K, C, L, N, kern, mt, 2)
    m.set_option("fold_emission", 1)
    if "c2" in args.cases:
        run(10, 100, 1_000_000, 256, 2, 0, 3)
    if "k10small" in args.cases:
        run(10, 100, 100_000, 256, 2, 0, 3)
        run(10, 100, 100_000, 256, 3, 0, 2)
        run(10, 100, 20_000, 256, 1, 0, 1)
    if "ncu_pair" in args.cases:
        run(10, 100, 50_000, 256, 2, 0, 1)
    if "ncu_dmma20" in args.cases:
        run(20, 128, 50_000, 128, 3, 1, 1)
    if "k20" in args.cases:
        for mt in (1, 2, 4):
            run(20, 128, 200_000, 128, 3, mt, 2)
        run(20, 128, 20_000, 128, 1, 0, 1)
    if "k40" in args.cases:
        for mt in (1, 2):
            run(40, 128, 100_000, 64, 3, mt, 2)
