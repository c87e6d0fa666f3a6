"""Host-side zipHMM-style preprocessing (csrc/tokenizer.inl; reference contract hmm.py:16) -- no GPU needed.

The re-encoding must be exact: expanding the tokens gives back the symbols bit for bit, and the oracle's
zip_forward over OUR dictionary equals the oracle's plain forward."""
import os

import numpy as np
import pytest

from conftest import ROOT
from oracle import forward as F


def expand(tokens, pairs, nsym):
    table = {}

    def ex(t):
        if t < nsym:
            return [t]
        if t not in table:
            l, r = pairs[t - nsym]
            table[t] = ex(int(l)) + ex(int(r))
        return table[t]
    out = []
    for t in tokens:
        out.extend(ex(int(t)))
    return np.asarray(out, dtype=np.uint8)


def example_chunks():
    obs = np.load(os.path.join(ROOT, "tests", "golden", "example_pair.npz"))["symbols"]
    return [obs[:30000], obs[30000:30001], obs[30001:30003], obs[30003:]]


def test_tokens_expand_back_to_the_symbols_bit_exact():
    import imcoalhmm_b200 as m
    chunks = example_chunks() + [np.zeros(0, dtype=np.uint8)]
    fset = m.ForwarderSet([m.Forwarder.from_symbols(c, 3) for c in chunks])
    info = fset.zip_info()
    pairs = fset.zip_pairs()
    assert 3 < info["ids_available"] <= 256 and pairs.shape == (info["ids_available"] - 3, 2)
    # every pair refers to earlier ids only
    for i, (l, r) in enumerate(pairs):
        assert l < 3 + i and r < 3 + i
    for ids in (3, 4, 17, info["ids_available"]):
        for c, chunk in enumerate(chunks):
            tok = fset.zip_tokens(c, ids)
            assert tok.size == 0 or tok.max() < ids
            assert np.array_equal(expand(tok, pairs, 3), chunk[1:]), (ids, c)
    full = sum(fset.zip_tokens(c).size for c in range(len(chunks)))
    assert full * 8 < sum(len(c) for c in chunks)         # the example alignment compresses well over 8x
    k10 = fset.zip_info(10)
    assert k10["ids_used"] <= k10["ids_available"] and k10["levels"] >= 1
    assert k10["tokens"] == sum(fset.zip_tokens(c, k10["ids_used"]).size for c in range(len(chunks)))
    k40 = fset.zip_info(40)
    assert 3 <= k40["ids_used"] <= k10["ids_used"] and k40["tokens"] >= k10["tokens"]


def test_oracle_zip_forward_over_our_dictionary_equals_plain():
    import imcoalhmm_b200 as m
    g = np.load(os.path.join(ROOT, "tests", "golden", "model_isolation_k10.npz"))
    pi, T, E = g["pi"][0], g["T"][0], g["E"][0]
    chunks = example_chunks()
    fset = m.ForwarderSet([m.Forwarder.from_symbols(c, 3) for c in chunks])
    pairs = fset.zip_pairs().astype(np.int32)
    nsyms = fset.zip_info()["ids_available"]
    for c, chunk in enumerate(chunks):
        new_obs = np.concatenate([chunk[:1], fset.zip_tokens(c)]).astype(np.int32)
        want = F.forward_plain(chunk.astype(np.int32), pi, T, E)
        got = F.zip_forward(pi, T, E, pairs, new_obs, 3, nsyms)
        assert abs(got - want) <= 1e-11 * max(1.0, abs(want)), (c, got, want)


@pytest.mark.parametrize("nsym", [1, 2, 5, 9])
def test_other_alphabets_and_degenerate_inputs(nsym):
    import imcoalhmm_b200 as m
    rng = np.random.default_rng(nsym)
    chunks = [rng.integers(0, nsym, size=n).astype(np.int32) for n in (1, 2, 3, 1000, 20000)]
    chunks.append(np.zeros(5000, dtype=np.int32))                    # one long run
    fset = m.ForwarderSet([m.Forwarder.from_symbols(c, nsym) for c in chunks])
    pairs = fset.zip_pairs()
    for c, chunk in enumerate(chunks):
        assert np.array_equal(expand(fset.zip_tokens(c), pairs, nsym), chunk[1:].astype(np.uint8))
