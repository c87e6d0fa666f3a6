"""Parameter points spread over +-3 decades around the script defaults: the compressed kernel vs the CPU oracle's plain
forward on the GPU-built (pi, T, E), on the example alignment.  python tools/fuzz_theta.py [points] [seed]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import imcoalhmm_b200 as m  # noqa: E402
from oracle import forward as F  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
obs = np.load(os.path.join(ROOT, "tests", "golden", "example_pair.npz"))["symbols"]
chunks = [obs[:20000], obs[20000:45000], obs[45000:]]
fset = m.ForwarderSet([m.Forwarder.from_symbols(c, 3) for c in chunks])
worst = {}
for name, model, default in (("isolation", m.IsolationModel(10), [1e-3, 2000.0, 0.4]),
                             ("im", m.IsolationMigrationModel(10, 10), [1e-3, 1e-3, 2000.0, 0.4, 200.0]),
                             ("psmc", m.VariableCoalescenceRateIsolationModel([4] * 10, True), [1e-3] + [1000.0] * 10 + [0.4])):
    default = np.asarray(default)
    thetas = default[None, :] * 10.0 ** rng.uniform(-3, 3, size=(n, default.size))
    thetas[0] = default
    pis, Ts, Es, st = model.build_hidden_markov_models(thetas, check_joint=False)
    ok = st == 0
    want, _ = F.forward_batch([c.astype(np.int32) for c in chunks], pis[ok], Ts[ok], Es[ok])
    for seg, lanes in ((0, 0), (-1, 8), (-1, 32), (64, 4)):
        m.set_option("zip_segment_tokens", seg)
        m.set_option("zip_lanes", lanes)
        got = fset.forward_batch(pis[ok], Ts[ok], Es[ok])
        fin = np.isfinite(want)
        bad = (np.isfinite(got) != fin)
        rel = np.abs(got[fin & ~bad] - want[fin & ~bad]) / np.abs(want[fin & ~bad])
        worst[(name, seg, lanes)] = (int(bad.sum()), float(rel.max()) if rel.size else 0.0, int(ok.sum()), int((~fin).sum()))
        if bad.any() or (rel.size and rel.max() > 1e-9):
            i = int(np.argmax(bad)) if bad.any() else int(np.argmax(np.where(fin & ~bad, np.abs(got - want) / np.abs(want), 0)))
            print("PROBLEM", name, seg, lanes, "theta", thetas[ok][i], "got", got[i], "want", want[i])
for k, v in worst.items():
    print(k, "mismatched finiteness %d, worst rel err %.2e, valid points %d, oracle non-finite %d" % v)
