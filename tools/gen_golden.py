#!/usr/bin/env python
"""Generate the committed golden vectors under tests/golden/ from the REFERENCE.

Runs only in the build container (needs /root/reference); the GPU box and the
test-suite read the committed fixtures, never the reference.

What is pinned:
  * model_*.npz   theta -> (pi, T, E) produced by the reference's own
                  `Model.build_hidden_markov_model` (src/IMCoalHMM/model.py:44-49)
                  executed through the py3 shim of tools/make_ref_shim.py.
  * statespaces.json  the reference's Isolation / Single / Migration state
                  spaces (src/IMCoalHMM/state_spaces.py) in a numbering-free
                  canonical form (state numbering is hash-order dependent).
  * example_pair.npz  hg18 vs pantro2 columns of examples/example_data.fa encoded
                  by the rule of scripts/prepare-alignments.py:99-105
                  (0 equal, 1 differ, 2 either base not in ACGT).

The forward log-likelihoods in forward_kat.json are NOT reference outputs (ziphmm is
absent, see DESIGN.md "parity unpinned"); they come from oracle/ in long double and
are written by tools/gen_forward_kat.py.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_ref_shim  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(HERE), "tests", "golden")


def thetas_around(default, n, seed=7, scale=0.1):
    """theta_b = default * exp(scale * N(0,1)) -- the MCMC proposal scale (mcmc.py:26,34-36)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    default = np.asarray(default, dtype=np.float64)
    out = default[None, :] * np.exp(scale * rng.standard_normal((n, default.size)))
    out[0] = default
    return out


def canon_state(state):
    return tuple(sorted((int(pop), tuple(sorted(int(x) for x in left)), tuple(sorted(int(x) for x in right)))
                        for pop, (left, right) in state))


def dump_statespace(space):
    canon = {canon_state(s): idx for s, idx in space.states.items()}
    ordered = sorted(canon)
    new_index = {canon[c]: k for k, c in enumerate(ordered)}
    edges = sorted([new_index[s], t[0], int(t[1]), int(t[2]), new_index[d]] for s, t, d in space.transitions)
    out = {
        "states": [repr(c) for c in ordered],
        "edges": edges,
        "begin": sorted(new_index[i] for i in space.begin_states),
        "left": sorted(new_index[i] for i in space.left_states),
        "right": sorted(new_index[i] for i in space.right_states),
        "end": sorted(new_index[i] for i in space.end_states),
    }
    for name in ("i11_index", "i12_index", "i22_index"):
        if hasattr(space, name):
            out[name] = new_index[getattr(space, name)]
    return out


def main():
    make_ref_shim.build()
    make_ref_shim.import_shim()
    from IMCoalHMM.isolation_model import IsolationModel
    from IMCoalHMM.isolation_with_migration_model import IsolationMigrationModel
    from IMCoalHMM.variable_coalescence_rate_isolation_model import VariableCoalescenceRateIsolationModel
    from IMCoalHMM.variable_migration_model import VariableCoalAndMigrationRateModel
    from IMCoalHMM.isolation_with_migration_model_epochs import IsolationMigrationEpochsModel
    from IMCoalHMM.state_spaces import Isolation, Single, Migration
    from IMCoalHMM import break_points as bp

    os.makedirs(GOLDEN, exist_ok=True)

    # ---- state spaces -------------------------------------------------------------
    spaces = {"Isolation": dump_statespace(Isolation()),
              "Single": dump_statespace(Single()),
              "Migration": dump_statespace(Migration())}
    with open(os.path.join(GOLDEN, "statespaces.json"), "w") as f:
        json.dump(spaces, f, separators=(",", ":"))

    # ---- break points (extra values beyond the reference's own unit-test goldens) ----
    bps = {
        "exp_10_2000_0.001": list(map(float, bp.exp_break_points(10, 2000.0, 0.001))),
        "uniform_10_0.001_0.002": list(map(float, bp.uniform_break_points(10, 0.001, 0.002))),
        "psmc_40": list(map(float, bp.psmc_break_points(40))),
        "psmc_40_off": list(map(float, bp.psmc_break_points(40, offset=0.001))),
        # trunc_exp_break_points (break_points.py:33-58) raises TypeError upstream
        # ("list + float" at :58) in Python 2 and 3 alike, and no model calls it: not pinned.
    }
    with open(os.path.join(GOLDEN, "break_points.json"), "w") as f:
        json.dump(bps, f)

    # ---- theta -> (pi, T, E) --------------------------------------------------------
    VM = VariableCoalAndMigrationRateModel
    cases = {
        # name: (constructor, ctor kwargs description, default theta, n thetas)
        "isolation_k10": (lambda: IsolationModel(10), {"model": "isolation", "no_hmm_states": 10},
                          [1e-3, 2000.0, 0.4], 16),
        "isolation_k4": (lambda: IsolationModel(4), {"model": "isolation", "no_hmm_states": 4},
                         [1.0, 0.5, 4e-4], 4),
        "im_k10_10": (lambda: IsolationMigrationModel(10, 10),
                      {"model": "im", "no_mig_states": 10, "no_ancestral_states": 10},
                      [1e-3, 1e-3, 2000.0, 0.4, 200.0], 16),
        "im_k3_4": (lambda: IsolationMigrationModel(3, 4),
                    {"model": "im", "no_mig_states": 3, "no_ancestral_states": 4},
                    [0.5, 1.0, 1.0, 0.4, 0.1], 4),
        "psmc_iso_split_4x10": (lambda: VariableCoalescenceRateIsolationModel([4] * 10, est_split=True),
                                {"model": "psmc_iso", "intervals": [4] * 10, "est_split": True},
                                [1e-3] + [1000.0] * 10 + [0.4], 16),
        "psmc_iso_nosplit_2_3": (lambda: VariableCoalescenceRateIsolationModel([2, 3], est_split=False),
                                 {"model": "psmc_iso", "intervals": [2, 3], "est_split": False},
                                 [1000.0, 500.0, 0.4], 4),
        "varmig_i12_4x10": (lambda: VM(VM.INITIAL_12, [4] * 10),
                            {"model": "varmig", "initial": 1, "intervals": [4] * 10},
                            [1000.0] * 20 + [100.0] * 20 + [0.4], 8),
        "varmig_i11_2_2": (lambda: VM(VM.INITIAL_11, [2, 2]),
                           {"model": "varmig", "initial": 0, "intervals": [2, 2]},
                           [1000.0, 800.0, 1200.0, 900.0, 100.0, 50.0, 150.0, 75.0, 0.4], 4),
        "varmig_i22_1_2": (lambda: VM(VM.INITIAL_22, [1, 2]),
                           {"model": "varmig", "initial": 2, "intervals": [1, 2]},
                           [1000.0, 800.0, 1200.0, 900.0, 100.0, 50.0, 150.0, 75.0, 0.4], 2),
        "im_epochs_2_3_3": (lambda: IsolationMigrationEpochsModel(2, 3, 3),
                            {"model": "im_epochs", "no_epochs": 2, "no_mig_states": 3, "no_ancestral_states": 3},
                            [1e-3, 1e-3, 0.4] + [2000.0, 1800.0, 2200.0, 1900.0, 2100.0] + [200.0, 150.0], 4),
    }
    for name, (ctor, desc, default, n) in cases.items():
        model = ctor()
        thetas = thetas_around(default, n)
        pis, Ts, Es = [], [], []
        for th in thetas:
            pi, T, E = model.build_hidden_markov_model(np.array(th))
            pis.append(np.asarray(pi, dtype=np.float64))
            Ts.append(np.asarray(T, dtype=np.float64))
            Es.append(np.asarray(E, dtype=np.float64))
        np.savez_compressed(os.path.join(GOLDEN, "model_%s.npz" % name),
                            desc=json.dumps(desc), theta=thetas,
                            pi=np.stack(pis), T=np.stack(Ts), E=np.stack(Es))
        print(name, "K =", pis[0].size, "P =", thetas.shape[1], "n =", n)

    # ---- example alignment (encoding contract) ------------------------------------
    seqs, cur = {}, None
    with open("/root/reference/examples/example_data.fa") as f:
        for line in f:
            line = line.strip()
            if line.startswith(">"):
                cur = line[1:].split()[0]
                seqs[cur] = []
            elif line:
                seqs[cur].append(line)
    a = "".join(seqs["hg18"]).upper()
    b = "".join(seqs["pantro2"]).upper()
    assert len(a) == len(b)
    clean = set("ACGT")
    sym = np.fromiter((2 if (x not in clean or y not in clean) else (0 if x == y else 1)
                       for x, y in zip(a, b)), dtype=np.uint8, count=len(a))
    np.savez_compressed(os.path.join(GOLDEN, "example_pair.npz"), symbols=sym)
    print("example_pair", len(sym), np.bincount(sym, minlength=3))


if __name__ == "__main__":
    main()
