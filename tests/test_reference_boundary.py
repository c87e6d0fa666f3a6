"""The reference's own code run against this package's boundary.

`/root/reference/src/IMCoalHMM/hmm.py` does `import ziphmm` (hmm.py:7) and calls exactly
`ziphmm.preprocess_raw_observations` (hmm.py:16) and `ziphmm.zip_forward` (hmm.py:20-21).  With
`sys.modules["ziphmm"] = imcoalhmm_b200.ziphmm` the reference's Forwarder, Likelihood and model classes (py3 shim of
the unmodified sources, tools/make_ref_shim.py) run on this library:

  * CPU (here, where /root/reference exists): the reference's Forwarder is constructed through OUR preprocessing and
    its Likelihood is evaluated with our (new_obs, sym2pair, new_nsyms) scored by the CPU oracle -- this pins that what
    we hand back satisfies the ziphmm contract under the reference's own glue, and reproduces the committed values
    tests/golden/reference_likelihood.json (tools/gen_reference_likelihood.py: the reference + the oracle's own
    zipHMM-style preprocessing).
  * GPU (the B200 box, where /root/reference does not exist): imcoalhmm_b200's drop-in classes reproduce the same
    committed values; where both a GPU and the reference are present the reference's code itself drives the kernels.
"""
import json
import os
import sys

import numpy as np
import pytest

from conftest import GOLDEN, ROOT, example_symbols

pytestmark = pytest.mark.filterwarnings("ignore")      # numpy.matrix deprecation noise from the reference's own modules

REFERENCE = "/root/reference/src/IMCoalHMM"
SHIM = "/tmp/imcoalhmm_ref_shim_boundary"
needs_reference = pytest.mark.skipif(not os.path.isdir(REFERENCE), reason="the reference sources are not on this machine")


def golden_cases():
    return json.load(open(os.path.join(GOLDEN, "reference_likelihood.json")))["cases"]


def write_chunks(tmp_path):
    sym = example_symbols()
    paths = []
    for k, part in enumerate((sym[:30000], sym[30000:])):
        p = tmp_path / ("chunk%d.txt" % k)
        p.write_text(" ".join(map(str, part.tolist())))          # prepare-alignments.py:93-105
        paths.append(str(p))
    return paths


@pytest.fixture
def reference_modules():
    """The reference's modules with `ziphmm` resolved to this package; everything is unloaded again afterwards."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import make_ref_shim
    import imcoalhmm_b200.ziphmm as our_ziphmm
    make_ref_shim.build(SHIM, with_hmm=True)
    saved = {k: v for k, v in sys.modules.items() if k == "ziphmm" or k.startswith("IMCoalHMM")}
    for k in saved:
        del sys.modules[k]
    sys.modules["ziphmm"] = our_ziphmm
    sys.path.insert(0, SHIM)
    try:
        import IMCoalHMM.hmm
        import IMCoalHMM.likelihood
        import IMCoalHMM.isolation_model
        import IMCoalHMM.isolation_with_migration_model
        import IMCoalHMM.variable_coalescence_rate_isolation_model
        yield sys.modules["IMCoalHMM"]
    finally:
        sys.path.remove(SHIM)
        for k in [k for k in sys.modules if k == "ziphmm" or k.startswith("IMCoalHMM")]:
            del sys.modules[k]
        sys.modules.update(saved)


def reference_model(ref, case):
    mod = {"IsolationModel": ref.isolation_model, "IsolationMigrationModel": ref.isolation_with_migration_model,
           "VariableCoalescenceRateIsolationModel": ref.variable_coalescence_rate_isolation_model}[case["model"]]
    return getattr(mod, case["model"])(*case["args"])


@needs_reference
def test_reference_forwarder_and_likelihood_over_our_preprocessing(reference_modules, tmp_path, monkeypatch):
    """hmm.py:12-16 through imcoalhmm_b200.ziphmm.preprocess_raw_observations; likelihood.py:27-33 with the forward of
    OUR encoding computed by the CPU oracle (no GPU in this container)."""
    ref = reference_modules
    import imcoalhmm_b200.ziphmm as our_ziphmm
    from oracle import forward as F
    paths = write_chunks(tmp_path)
    forwarders = [ref.hmm.Forwarder(p, 3) for p in paths]
    sym = example_symbols()
    for f, part in zip(forwarders, (sym[:30000], sym[30000:])):
        assert f.NSYM == 3 and f.new_nsyms >= 3 and np.asarray(f.sym2pair).shape == (f.new_nsyms - 3, 2)
        assert len(f.new_obs) < len(part) / 10                                        # really compressed
        back = our_ziphmm._expand(f.sym2pair, np.asarray(f.new_obs), 3, f.new_nsyms)
        assert np.array_equal(back, part)                                             # an exact re-encoding of the file
    monkeypatch.setattr(our_ziphmm, "zip_forward", lambda pi, T, E, s2p, obs, nsym, nn: F.zip_forward(
        np.asarray(pi).reshape(-1), np.asarray(T), np.asarray(E), np.asarray(s2p), np.asarray(obs), nsym, nn))
    for case in golden_cases():
        like = ref.likelihood.Likelihood(reference_model(ref, case), forwarders)
        for th, want in zip(case["thetas"], case["logL"]):
            got = like(np.array(th))
            if want == "-inf":
                assert got == -float("inf")
            else:
                assert got == pytest.approx(want, rel=1e-12)


@pytest.mark.gpu
def test_drop_in_classes_reproduce_the_reference_values(tmp_path):
    """imcoalhmm_b200.Likelihood(Model, [Forwarder(path, 3), ...])(theta) == what the reference's own Likelihood returned."""
    import imcoalhmm_b200 as m
    paths = write_chunks(tmp_path)
    forwarders = [m.Forwarder(p, 3) for p in paths]
    for case in golden_cases():
        model = getattr(m, case["model"])(*case["args"])
        like = m.Likelihood(model, forwarders)
        for th, want in zip(case["thetas"], case["logL"]):
            got = like(np.array(th))
            if want == "-inf":
                assert got == -float("inf")
            else:
                assert got == pytest.approx(want, rel=1e-11), (case["name"], th)
        finite = [(th, w) for th, w in zip(case["thetas"], case["logL"]) if w != "-inf"]
        batched = like.batched(np.array([th for th, _ in finite]))
        np.testing.assert_allclose(batched, [w for _, w in finite], rtol=1e-11)
        # the reference-style two-step route: build_hidden_markov_model on the GPU, then Forwarder.forward per chunk
        pi, T, E = model.build_hidden_markov_model(np.array(finite[0][0]))
        assert sum(f.forward(pi, T, E) for f in forwarders) == pytest.approx(finite[0][1], rel=1e-11)


@pytest.mark.gpu
@needs_reference
def test_reference_code_drives_the_kernels(reference_modules, tmp_path):
    """The reference's unmodified Forwarder + Likelihood + model classes with ziphmm = imcoalhmm_b200.ziphmm: its
    Python builds (pi, T, E) on the CPU, our zip_forward scores them on the GPU (needs a GPU AND the reference)."""
    import imcoalhmm_b200 as m
    ref = reference_modules
    paths = write_chunks(tmp_path)
    forwarders = [ref.hmm.Forwarder(p, 3) for p in paths]
    ours = [m.Forwarder(p, 3) for p in paths]
    for case in golden_cases():
        like = ref.likelihood.Likelihood(reference_model(ref, case), forwarders)
        mine = m.Likelihood(getattr(m, case["model"])(*case["args"]), ours)
        for th, want in zip(case["thetas"], case["logL"]):
            got = like(np.array(th))
            if want == "-inf":
                assert got == -float("inf") and mine(np.array(th)) == -float("inf")
            else:
                assert got == pytest.approx(want, rel=1e-11) and got == pytest.approx(mine(np.array(th)), rel=1e-11)
