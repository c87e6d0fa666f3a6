"""Host-side zipHMM-style preprocessing (csrc/tokenizer.inl; reference contract hmm.py:16) -- no GPU needed.

The re-encoding must be exact: expanding the tokens gives back the symbols bit for bit, and the oracle's
zip_forward over OUR dictionary equals the oracle's plain forward."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, ROOT
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


def _expand_run_tokens(first_run, words, pairs, nsym, run_sym):
    table = [[s] for s in range(nsym)]
    for left, right in pairs:
        table.append(table[left] + table[right])
    out = [run_sym] * first_run
    for w in words.tolist():
        out.extend(table[w & 0xff])
        out.extend([run_sym] * (w >> 8))
    return np.array(out, dtype=np.uint8)


def test_run_tokens_round_trip_and_contract():
    """The second re-encoding (spectral form): first_run sites of the run symbol, then words id | n << 8; any dictionary
    cap expands back to the symbols bit for bit; runs longer than 4095 continue through entries of the run symbol itself;
    pairs are learned over the entries only (the left part carries no run)."""
    import imcoalhmm_b200 as m
    rng = np.random.default_rng(42)
    obs = np.load(os.path.join(GOLDEN, "example_pair.npz"))["symbols"]
    long_run = rng.choice(3, size=30000, p=[0.955, 0.005, 0.04]).astype(np.uint8)
    long_run[2000:14000] = 0                                   # 12 000 matching sites in a row
    chunks = [obs[:30000], obs[30000:], long_run, np.zeros(5000, dtype=np.uint8), np.array([1], dtype=np.uint8),
              np.array([2, 0, 0], dtype=np.uint8), np.zeros(0, dtype=np.uint8)]
    s = m.ForwarderSet([m.Forwarder.from_symbols(c, 3) for c in chunks])
    info = s.run_info()
    assert info["run_sym"] == 0 and 3 <= info["ids_available"] <= 256
    pairs = s.run_pairs()
    assert pairs.shape == (info["ids_available"] - 3, 2) and (pairs < np.arange(3, info["ids_available"])[:, None]).all()
    for ids in (None, 3, 4, min(9, info["ids_available"]), info["ids_available"]):
        for c, chunk in enumerate(chunks):
            first_run, words = s.run_tokens(c, ids)
            assert first_run <= 4095 and ((words >> 8) <= 4095).all()
            if ids is not None:
                assert ((words & 0xff) < ids).all()
            if len(chunk) == 0:
                assert first_run == 0 and words.size == 0
                continue
            back = _expand_run_tokens(first_run, words, pairs, 3, 0)
            assert np.array_equal(back, chunk[1:]), (ids, c)
    # the all-matching chunk is one leading run plus one continuation token; the long run costs two continuation entries
    first_run, words = s.run_tokens(3)
    assert first_run == 4095 and words.size == 1 and (words[0] & 0xff) == 0 and (words[0] >> 8) == 5000 - 1 - 4095 - 1
    # far fewer tokens than the pair dictionary on alignment-like data: one per mismatch or missing-data block
    k10 = s.run_info(10)
    assert k10["tokens"] * 2 < s.zip_info(10)["tokens"] * 3 and k10["tokens"] * 15 < s.total_sites
