"""configs[3] in miniature: lock-step MCMC chains on a synthetic alignment, every step one batched likelihood call.
    python tools/mcmc_demo.py --chains 4096 --chunks 20 --steps 30"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import imcoalhmm_b200 as m  # noqa: E402
from imcoalhmm_b200.mcmc import BatchedMCMC, ExpLogNormPrior, LogNormPrior  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--chains", type=int, default=4096)
    ap.add_argument("--chunks", type=int, default=20)
    ap.add_argument("--steps", type=int, default=30)
    args = ap.parse_args()
    wl = dict(bench.WORKLOADS["c2"], chunks=args.chunks)
    model = m.IsolationModel(10)
    truth = np.asarray(wl["default"])
    pi, T, E = model.build_hidden_markov_model(truth)
    chunks = bench.make_chunks(wl, pi[None], T[None], E[None], range(wl["chunks"]))
    lik = m.Likelihood(model, [m.Forwarder.from_symbols(c, 3) for c in chunks])
    # priors as in scripts/isolation-model-mcmc.py:120-124 (log-normal split time, exponential rates)
    priors = [LogNormPrior(np.log(1e-3)), ExpLogNormPrior(2000.0), ExpLogNormPrior(0.4)]
    rng = np.random.default_rng(1)
    t0 = time.perf_counter()
    mc = BatchedMCMC(priors, lik, thinning=1, no_chains=args.chains, rng=rng)
    print("init: %d chains scored in %.3f s, mean logL %.1f" % (args.chains, time.perf_counter() - t0, mc.current_likelihood.mean()))
    t0 = time.perf_counter()
    acc = 0.0
    for _ in range(args.steps):
        acc += mc.step().mean()
    dt = time.perf_counter() - t0
    sites = sum(len(c) for c in chunks)
    print("%d steps x %d chains on %d sites: %.3f s/step, %.3e sites*proposals/s, acceptance %.2f"
          % (args.steps, args.chains, sites, dt / args.steps, sites * args.chains * args.steps / dt, acc / args.steps))
    best = mc.current_theta[np.argmax(mc.current_posterior)]
    print("true theta", truth, "best chain", best, "mean logL %.1f" % mc.current_likelihood.mean())


if __name__ == "__main__":
    main()
