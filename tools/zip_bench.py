"""Device-side timing sweep of the zip (compressed) forward kernel on bench.py's synthetic workloads
(development aid; bench.py is the contract).  python tools/zip_bench.py --workload c2 --chunks 100"""
import argparse
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import imcoalhmm_b200 as m  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--chunks", type=int, default=0)
    ap.add_argument("--points", type=int, default=0)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--sweep", default="0:0:0", help="comma list of ctas_per_sm:max_entries:lanes[:pipeline pieces[:spectral 0 auto|1|2 off[:mma 0 auto|1|2 off[:shape[:run2[:align 0 auto|1|2 off]]]]]]")
    ap.add_argument("--plain", action="store_true", help="also time the uncompressed kernel")
    ap.add_argument("--segment", type=int, default=0, help="zip_segment_tokens: 0 auto, -1 whole chunks, > 0 segment length")
    ap.add_argument("--missing", type=float, default=0.04, help="missing-data coverage of the simulated chunks (bench.py: 0.04)")
    ap.add_argument("--check", type=int, default=2, help="chunks to check against the oracle")
    args = ap.parse_args()
    wl = dict(bench.WORKLOADS[args.workload])
    if args.chunks:
        wl["chunks"] = args.chunks
    if args.points:
        wl["points"] = args.points
    K, N = wl["K"], wl["points"]
    model = getattr(m, wl["ctor"][0])(*wl["ctor"][1])
    thetas = bench.thetas_around(wl["default"], N)
    pis, Ts, Es, st = model.build_hidden_markov_models(thetas)
    assert (st == 0).all()
    t0 = time.time()
    chunks = bench.make_chunks(wl, pis, Ts, Es, range(wl["chunks"]), missing=args.missing)
    t_sim = time.time() - t0
    t0 = time.time()
    fset = m.ForwarderSet([m.Forwarder.from_symbols(c, 3) for c in chunks])
    t_prep = time.time() - t0
    print("simulate %.1fs, preprocess %.1fs, sites %d, zip info %s" % (t_sim, t_prep, fset.total_sites, fset.zip_info(K)), flush=True)
    d_pi, d_T, d_E = (torch.tensor(x, device="cuda") for x in (pis, Ts, Es))
    d_out = torch.empty(N, dtype=torch.float64, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream

    def step():
        fset.forward_batch_device(d_pi.data_ptr(), d_T.data_ptr(), d_E.data_ptr(), d_out.data_ptr(), N, K, 3, stream)

    def timed(label):
        step()
        torch.cuda.synchronize()
        best = 1e30
        for _ in range(args.reps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            step()
            b.record()
            torch.cuda.synchronize()
            best = min(best, a.elapsed_time(b) * 1e-3)
        sp = fset.total_sites * N
        print("%-60s kernel=%-14s %9.3f ms  %.3e site-pts/s  logL[0]=%.6f" %
              (label, m.last_forward_kernel(), best * 1e3, sp / best, d_out[0].item()), flush=True)
        return d_out.cpu().numpy().copy()

    ref = None
    for spec in args.sweep.split(","):
        f = list(map(int, spec.split(":")))
        ctas, cap, lanes = f[:3]
        pipe = f[3] if len(f) > 3 else 0
        spec = f[4] if len(f) > 4 else 0
        m.set_option("zip_pipeline", pipe)
        m.set_option("zip_segment_tokens", args.segment)
        m.set_option("zip_spectral", spec)
        mma = f[5] if len(f) > 5 else 0
        m.set_option("zip_mma", mma)
        m.set_option("zip_mma_shape", f[6] if len(f) > 6 else 0)
        m.set_option("zip_run2", f[7] if len(f) > 7 else 0)
        al = f[8] if len(f) > 8 else 0
        try:
            m.set_option("zip_align", al)
        except m.IMCError:
            pass                      # an older experiment build (IMC_LIB_PATH)
        m.set_option("forward_kernel", 4)
        m.set_option("zip_lanes", lanes)
        m.set_option("zip_ctas_per_sm", ctas)
        m.set_option("zip_max_entries", cap)
        info = fset.zip_info(K)
        m.set_option("zip_ctas_per_sm", ctas)   # zip_info reports the plan in effect
        info = fset.zip_info(K)
        rinfo = fset.run_info(K)
        out = timed("zip lanes=%d ctas=%d cap=%d pipe=%d spec=%d mma=%d align=%d M=%d/%d tok=%d/%d" % (lanes, ctas, cap, pipe, spec, mma, al, info["ids_used"],
                    rinfo["ids_used"], info["tokens"], rinfo["tokens"]))
        if ref is None:
            ref = out
        else:
            print("    max rel diff vs first config: %.2e" % np.max(np.abs(out - ref) / np.abs(ref)))
    if args.plain:
        m.set_option("forward_kernel", 2 if K <= 12 else 3)
        out = timed("plain")
        print("    max rel diff zip vs plain: %.2e" % np.max(np.abs(out - ref) / np.abs(ref)))
    if args.check:
        from oracle import forward as F
        nc, npts = min(args.check, len(chunks)), min(4, N)
        sub = m.ForwarderSet([m.Forwarder.from_symbols(c, 3) for c in chunks[:nc]])
        m.set_option("forward_kernel", 4)
        got = sub.forward_batch(pis[:npts], Ts[:npts], Es[:npts])
        want, _ = F.forward_batch([c.astype(np.int32) for c in chunks[:nc]], pis[:npts], Ts[:npts], Es[:npts])
        print("oracle check (%d chunks x %d points): max rel err %.2e" % (nc, npts, np.max(np.abs(got - want) / np.abs(want))))


if __name__ == "__main__":
    main()
