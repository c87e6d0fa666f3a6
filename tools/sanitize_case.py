"""Small inputs through every launch shape of the forward path, for compute-sanitizer (memcheck / racecheck / synccheck):
    compute-sanitizer --tool racecheck python tools/sanitize_case.py [--ranks 2 --rank R --idfile F]
Shapes: plain form (8 / 4 / 32 lanes, pipelined pieces, segments + folds), spectral form (FMA shapes, MMA shape, pieces,
segments), the plain-form pass for points that do not qualify, the per-site kernels, the model build, and -- with two
processes -- the fused peer-to-peer all-reduce."""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import imcoalhmm_b200 as m  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--ranks", type=int, default=1)
ap.add_argument("--rank", type=int, default=0)
ap.add_argument("--idfile", default="/tmp/imc_sanitize_id")
args = ap.parse_args()
m._lib.check(m._lib.load().imc_init(args.rank))
if args.ranks > 1:
    if args.rank == 0:
        with open(args.idfile + ".tmp", "wb") as f:
            f.write(m._lib.comm_unique_id())
        os.rename(args.idfile + ".tmp", args.idfile)
    else:
        while not os.path.exists(args.idfile):
            time.sleep(0.1)
    m._lib.comm_init(args.ranks, args.rank, open(args.idfile, "rb").read())
    print("communicator:", m._lib.comm_info(), flush=True)

obs = np.load(os.path.join(ROOT, "tests", "golden", "example_pair.npz"))["symbols"]
rng = np.random.default_rng(5 + args.rank)
chunks = [obs[i * 2500:(i + 1) * 2500] for i in range(args.rank, 12, args.ranks)] + [obs[40000:40001], obs[50000:50017]]
fset = m.ForwarderSet([m.Forwarder.from_symbols(c, 3) for c in chunks])
done = []
for name, n in (("isolation_k10", 3), ("im_k10_10", 2), ("psmc_iso_split_4x10", 2)):
    g = np.load(os.path.join(ROOT, "tests", "golden", "model_%s.npz" % name))
    pis, Ts, Es = g["pi"][:n].copy(), g["T"][:n].copy(), g["E"][:n].copy()
    Ts[-1] = 0.9 * np.eye(Ts.shape[1]) + 0.1 * rng.dirichlet(np.ones(Ts.shape[1]), size=Ts.shape[1])   # not reversible: plain-form pass
    ref = None
    for opts in (dict(zip_spectral=2, zip_lanes=8), dict(zip_spectral=2, zip_lanes=4), dict(zip_spectral=2, zip_lanes=32),
                 dict(zip_spectral=2, zip_pipeline=3, zip_segment_tokens=-1), dict(zip_spectral=2, zip_segment_tokens=64),
                 dict(zip_spectral=1, zip_mma=2, zip_segment_tokens=-1), dict(zip_spectral=1, zip_mma=2, zip_pipeline=3, zip_segment_tokens=-1),
                 dict(zip_spectral=1, zip_mma=2, zip_segment_tokens=64), dict(zip_spectral=1, zip_mma=1, zip_segment_tokens=-1),
                 dict(zip_spectral=1, zip_mma=1, zip_pipeline=3, zip_segment_tokens=-1), dict(zip_spectral=1, zip_mma=1, zip_segment_tokens=64),
                 dict(zip_spectral=1, zip_mma=1, zip_run2=1, zip_segment_tokens=-1), dict(zip_spectral=1, zip_mma=1, zip_run2=1, zip_pipeline=3, zip_segment_tokens=-1),
                 dict(zip_spectral=1, zip_mma=1, zip_run2=1, zip_segment_tokens=64),
                 dict(forward_kernel=1), dict(forward_kernel=3)):
        for k in ("forward_kernel", "zip_spectral", "zip_mma", "zip_run2", "zip_lanes", "zip_pipeline", "zip_segment_tokens"):
            m.set_option(k, opts.get(k, 2 if k == "zip_run2" else 0))
        out = fset.forward_batch(pis, Ts, Es)
        ref = out if ref is None else ref
        assert np.allclose(out, ref, rtol=1e-10), (name, opts, out, ref)
        done.append("%s/%s" % (name, m.last_forward_kernel()))
for k in ("forward_kernel", "zip_spectral", "zip_mma", "zip_run2", "zip_lanes", "zip_pipeline", "zip_segment_tokens"):
    m.set_option(k, 0)
g = np.load(os.path.join(ROOT, "tests", "golden", "model_im_k10_10.npz"))
fused = m.IsolationMigrationModel(10, 10).batched_log_likelihood(g["theta"][:3], fset)
assert np.isfinite(fused).all()
if args.ranks > 1:
    m._lib.comm_destroy()
print("sanitize_case ok (rank %d): %d launches over %s" % (args.rank, m.kernel_launches(), sorted(set(done))), flush=True)
