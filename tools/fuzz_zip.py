"""Randomised cross-check of the zip kernel (all launch shapes, segmented / warp-per-chain modes, dictionary caps, the
spectral form over run tokens with reversible, non-reversible and mixed batches) against the CPU oracle.
python tools/fuzz_zip.py [trials] [seed]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import imcoalhmm_b200 as m  # noqa: E402
from oracle import forward as F  # noqa: E402

trials = int(sys.argv[1]) if len(sys.argv) > 1 else 100
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
worst = 0.0
for trial in range(trials):
    K = int(rng.choice([1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 16, 17, 20, 24, 27, 32, 40]))
    N = int(rng.choice([1, 1, 2, 3, 8, 37]))
    C = int(rng.choice([1, 1, 2, 5, 17, 33, 70]))
    nsym = int(rng.choice([3, 3, 3, 2, 5]))
    pis = rng.dirichlet(np.ones(K), size=N)
    stick = rng.choice([0.5, 0.99, 0.9999])
    Ts = np.stack([stick * np.eye(K) + (1 - stick) * rng.dirichlet(np.ones(K), size=K) for _ in range(N)])
    Es = rng.dirichlet(np.ones(nsym), size=(N, K))
    rev = rng.random(N) < rng.choice([0.0, 0.5, 1.0, 1.0])    # reversible points: diag(pi) T symmetric, as the reference's models are
    for n in np.flatnonzero(rev):
        J = rng.random((K, K)) + 0.01
        J = 0.5 * (J + J.T) * (1 - stick) / J.sum()
        J[np.diag_indices(K)] += stick * rng.dirichlet(np.ones(K) * 3.0)
        J /= J.sum()
        pis[n] = J.sum(axis=1)
        Ts[n] = J / pis[n][:, None]
    if rng.random() < 0.2:                                   # an impossible symbol for some states
        Es[:, rng.integers(0, K), rng.integers(0, nsym)] = 0.0
    p = rng.dirichlet(np.ones(nsym) * rng.choice([0.05, 1.0]))
    chunks = []
    for _ in range(C):
        L = int(rng.choice([0, 1, 2, 15, 16, 17, 33, 200, 1500, 6000, 40000]))
        chunks.append(rng.choice(nsym, size=L, p=p).astype(np.int32))
    want, _ = F.forward_batch(chunks, pis, Ts, Es)
    fset = m.ForwarderSet([m.Forwarder.from_symbols(c, nsym) for c in chunks])
    lanes = int(rng.choice([0, 4, 8, 32]))
    ctas = int(rng.choice([0, 1, 2]))
    seg = int(rng.choice([0, 0, -1, 16, 64, 700]))
    cap = int(rng.choice([0, 0, nsym, nsym + 1, 9, 40]))
    cap = cap if cap == 0 or cap >= nsym else nsym
    pipe = int(rng.choice([0, 0, 1, 2, 3, 7, 32]))           # pipelined pieces (needs >= 256 tokens per piece to engage)
    fk = int(rng.choice([4, 4, 4, 0, 1, 2, 3]))              # mostly the zip kernel; also the automatic choice and the per-site kernels
    if fk in (1, 2, 3) and nsym > 3:
        fk = 4                                               # the packed 2-bit layout of the per-site kernels holds 3 symbols
    if fk == 2 and (K % 2 or K > 12):
        fk = 1
    if fk == 3 and K not in (10, 12, 16, 20, 24, 28, 32, 36, 40, 48, 64):
        fk = 1
    spec = int(rng.choice([0, 1, 1, 1, 2]))                  # spectral form: auto, forced (points that do not qualify take the plain form), off
    for k, v in (("forward_kernel", fk), ("zip_lanes", lanes), ("zip_ctas_per_sm", ctas), ("zip_segment_tokens", seg), ("zip_max_entries", cap),
                 ("zip_pipeline", pipe), ("zip_spectral", spec), ("zip_mma", int(rng.choice([0, 1, 2]))), ("zip_run2", int(rng.choice([0, 1, 2]))),
                 ("zip_align", int(rng.choice([0, 1, 1, 2])))):
        m.set_option(k, v)
    got = fset.forward_batch(pis, Ts, Es)
    # the oracle's plain forward divides by the zero scale of an impossible observation and returns NaN where the
    # answer is -inf; accept -inf there
    want = np.where(np.isnan(want) & np.isneginf(got), -np.inf, want)
    with np.errstate(invalid="ignore", divide="ignore"):
        same_inf = np.isneginf(got) == np.isneginf(want)
        fin = np.isfinite(want)
        err = float(np.max(np.abs(got[fin] - want[fin]) / np.maximum(1e-300, np.abs(want[fin])))) if fin.any() else 0.0
    ok = bool(same_inf.all()) and err < 1e-10 and bool((np.isfinite(got) == np.isfinite(want)).all())
    worst = max(worst, err)
    if not ok:
        print("MISMATCH trial %d: K=%d N=%d C=%d nsym=%d lanes=%d ctas=%d seg=%d cap=%d pipe=%d spec=%d rev=%s kernel=%s err=%.3e\n got %s\nwant %s"
              % (trial, K, N, C, nsym, lanes, ctas, seg, cap, pipe, spec, rev.astype(int), m.last_forward_kernel(), err, got[:4], want[:4]))
        sys.exit(1)
print("fuzz ok: %d trials, worst relative error %.2e" % (trials, worst))
