#!/usr/bin/env python
"""Run the reference's OWN Forwarder / Likelihood / model code end to end on the CPU and commit what it returns.

The reference's hmm.py imports `ziphmm` (hmm.py:7), which is absent here; this script puts a stand-in backed by the CPU
oracle (oracle/forward.py: zipHMM's published algorithm restated) into sys.modules, builds the py3 shim of the reference
INCLUDING hmm.py and likelihood.py (tools/make_ref_shim.py, outside the repository), writes the example alignment
(tests/golden/example_pair.npz, the encoding of examples/example_data.fa) in the reference's text format, and evaluates
    IMCoalHMM.likelihood.Likelihood(Model, [IMCoalHMM.hmm.Forwarder(path, 3)])(theta)
exactly as scripts/isolation-model.py:82-100 does.  Output: tests/golden/reference_likelihood.json -- the numbers the GPU
path must reproduce (tests/test_reference_boundary.py).  Only numbers are committed, never reference source."""
import json
import os
import sys
import tempfile
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import make_ref_shim  # noqa: E402
from oracle import forward as F  # noqa: E402

SHIM = "/tmp/imcoalhmm_ref_shim_full"


def oracle_ziphmm():
    mod = types.ModuleType("ziphmm")
    mod.preprocess_raw_observations = lambda obs, nsym: F.zip_preprocess(obs, nsym)
    mod.zip_forward = lambda pi, T, E, sym2pair, new_obs, nsym, new_nsyms: F.zip_forward(
        np.asarray(pi).reshape(-1), np.asarray(T), np.asarray(E), sym2pair, new_obs, nsym, new_nsyms)
    return mod


def cases():
    return [
        ("isolation_10", "IsolationModel", [10], [[1e-3, 2000.0, 0.4], [1.2e-3, 1500.0, 0.5], [0.8e-3, 2600.0, 0.25],
                                                   [1e-3, -1.0, 0.4]]),
        ("im_10_10", "IsolationMigrationModel", [10, 10], [[1e-3, 1e-3, 2000.0, 0.4, 200.0], [0.7e-3, 1.4e-3, 1800.0, 0.5, 350.0],
                                                             [1e-3, 1e-3, 2000.0, 0.4, 0.0]]),
        ("psmc_iso_split_2x3", "VariableCoalescenceRateIsolationModel", [[2, 3], True], [[1e-3, 1000.0, 1400.0, 0.4]]),
    ]


def main():
    make_ref_shim.build(SHIM, with_hmm=True)
    sys.path.insert(0, SHIM)
    sys.modules["ziphmm"] = oracle_ziphmm()
    from IMCoalHMM.hmm import Forwarder
    from IMCoalHMM.likelihood import Likelihood
    from IMCoalHMM.isolation_model import IsolationModel
    from IMCoalHMM.isolation_with_migration_model import IsolationMigrationModel
    from IMCoalHMM.variable_coalescence_rate_isolation_model import VariableCoalescenceRateIsolationModel
    ctors = {"IsolationModel": IsolationModel, "IsolationMigrationModel": IsolationMigrationModel,
             "VariableCoalescenceRateIsolationModel": VariableCoalescenceRateIsolationModel}
    sym = np.load(os.path.join(ROOT, "tests", "golden", "example_pair.npz"))["symbols"]
    out = {"alignment": "tests/golden/example_pair.npz, split at site 30000 into two Forwarders", "cases": []}
    with tempfile.TemporaryDirectory() as tmp:
        paths = []
        for k, part in enumerate((sym[:30000], sym[30000:])):
            p = os.path.join(tmp, "chunk%d.txt" % k)
            with open(p, "w") as f:
                f.write(" ".join(map(str, part.tolist())))       # prepare-alignments.py:93-105: blank-separated integers
            paths.append(p)
        forwarders = [Forwarder(p, 3) for p in paths]
        for name, ctor, args, thetas in cases():
            like = Likelihood(ctors[ctor](*args), forwarders)
            vals = [float(like(np.array(th))) for th in thetas]
            out["cases"].append({"name": name, "model": ctor, "args": args, "thetas": thetas,
                                 "logL": [v if np.isfinite(v) else "-inf" for v in vals]})
            print(name, vals)
    with open(os.path.join(ROOT, "tests", "golden", "reference_likelihood.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
