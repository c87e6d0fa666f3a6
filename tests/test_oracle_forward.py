"""Pins the CPU forward oracle mathematically (parity at the ziphmm boundary is unpinned by
the reference, so the oracle is anchored on the definition and on reference-derived (pi, T, E))."""
import numpy as np
import pytest

from oracle import forward as F
from conftest import golden_model, example_symbols, random_hmm


def test_bruteforce_matches_forward():
    rng = np.random.default_rng(11)
    for K, L in [(2, 8), (3, 7), (4, 6)]:
        pi, T, E = random_hmm(rng, K, missing_col=False)
        obs = rng.integers(0, 3, size=L).astype(np.int32)
        want = F.forward_bruteforce(obs, pi, T, E)
        assert F.forward_plain(obs, pi, T, E) == pytest.approx(want, rel=1e-13)
        assert F.forward_numpy(obs, pi, T, E) == pytest.approx(want, rel=1e-13)


def test_single_site_and_all_missing():
    _, pi, T, E = golden_model("isolation_k10")
    pi, T, E = pi[0], T[0], E[0]
    for o in range(3):
        assert F.forward_plain(np.array([o], dtype=np.int32), pi, T, E) == pytest.approx(np.log(pi @ E[:, o]), rel=1e-14)
    # symbol 2 emits with probability 1 from every state (emissions.py:99) => P(all missing) = sum(pi) = 1
    assert abs(F.forward_plain(np.full(1000, 2, dtype=np.int32), pi, T, E)) < 1e-11
    assert F.forward_plain(np.zeros(0, dtype=np.int32), pi, T, E) == 0.0


def test_survey_probe_values_on_example_alignment():
    """own restatement on hg18 vs pantro2 with reference-built (pi, T, E)."""
    obs = example_symbols().astype(np.int32)
    assert np.bincount(obs, minlength=3).tolist() == [62137, 642, 2476]
    _, pi, T, E = golden_model("isolation_k10")
    assert F.forward_plain(obs, pi[0], T[0], E[0]) == pytest.approx(-3729.5586472699, rel=1e-12)
    _, pi, T, E = golden_model("im_k10_10")
    assert F.forward_plain(obs, pi[0], T[0], E[0]) == pytest.approx(-3650.0493084297, rel=1e-12)


def test_double_vs_long_double():
    obs = example_symbols().astype(np.int32)
    for name in ("isolation_k10", "im_k10_10", "psmc_iso_split_4x10"):
        _, pi, T, E = golden_model(name)
        d = F.forward_plain(obs, pi[1], T[1], E[1])
        hi, lo = F.forward_plain_ld(obs, pi[1], T[1], E[1])
        assert abs(d - (hi + lo)) <= 1e-12 * abs(d)


def test_time_reversal_invariance():
    """J is symmetric by construction (transitions.py:237) => pi_i T_ij = pi_j T_ji => logL(obs) = logL(obs[::-1]).
    Using T^T by mistake breaks this at the 1e-4 level, so the test pins the orientation of T."""
    rng = np.random.default_rng(5)
    _, pi, T, E = golden_model("im_k10_10")
    obs = rng.choice(3, size=5000, p=[0.95, 0.01, 0.04]).astype(np.int32)
    a = F.forward_plain(obs, pi[0], T[0], E[0])
    b = F.forward_plain(obs[::-1].copy(), pi[0], T[0], E[0])
    assert a == pytest.approx(b, rel=1e-12)
    wrong = F.forward_plain(obs, pi[0], T[0].T.copy(), E[0])
    assert abs(wrong - a) > 1e-6 * abs(a)


def test_zip_equals_plain_and_chunk_additivity():
    obs = example_symbols().astype(np.int32)
    _, pi, T, E = golden_model("isolation_k10")
    new_obs, sym2pair, new_nsyms = F.zip_preprocess(obs, 3)
    assert new_nsyms > 3 and new_obs.size < obs.size // 10
    plain = F.forward_plain(obs, pi[2], T[2], E[2])
    assert F.zip_forward(pi[2], T[2], E[2], sym2pair, new_obs, 3, new_nsyms) == pytest.approx(plain, rel=1e-12)
    # batch entry: sum over chunks for several parameter points (likelihood.py:33)
    chunks = [obs[:20000], obs[20000:20001], obs[20001:]]
    out, _ = F.forward_batch(chunks, pi[:4], T[:4], E[:4])
    want = [sum(F.forward_plain(c, pi[n], T[n], E[n]) for c in chunks) for n in range(4)]
    np.testing.assert_allclose(out, want, rtol=1e-13)
    zipped = [F.zip_preprocess(c, 3) for c in chunks]
    outz, _ = F.forward_batch(None, pi[:4], T[:4], E[:4], mode="zip", zipped=zipped)
    np.testing.assert_allclose(outz, want, rtol=1e-12)


def test_tuned_zip_forward_matches_the_simple_one():
    """bench.py times imco_zip_forward_fast; it must be the same function as the simple restatement."""
    import numpy as np
    from conftest import golden_model, example_symbols
    from oracle import forward as F
    obs = example_symbols().astype(np.int32)
    for name in ("isolation_k10", "im_k10_10", "isolation_k4"):
        _, pis, Ts, Es = golden_model(name)
        for max_syms in (3, 4, 64, 1024):
            new_obs, s2p, ns = F.zip_preprocess(obs, 3, max_syms=max_syms)
            want = F.forward_plain(obs, pis[1], Ts[1], Es[1])
            assert abs(F.zip_forward_fast(pis[1], Ts[1], Es[1], s2p, new_obs, 3, ns) - want) <= 1e-11 * abs(want)
            assert abs(F.zip_forward(pis[1], Ts[1], Es[1], s2p, new_obs, 3, ns) - want) <= 1e-11 * abs(want)
        # compound first symbol, tiny inputs, impossible observation
        for n in (1, 2, 3, 17):
            new_obs, s2p, ns = F.zip_preprocess(obs[:n], 3, min_count=2, max_syms=8)
            want = F.forward_plain(obs[:n], pis[0], Ts[0], Es[0])
            assert abs(F.zip_forward_fast(pis[0], Ts[0], Es[0], s2p, new_obs, 3, ns) - want) <= 1e-12 * max(1.0, abs(want))
    rep = np.tile(np.array([0, 0, 0, 0], dtype=np.int32), 64)
    new_obs, s2p, ns = F.zip_preprocess(rep, 3, min_count=2, max_syms=16)
    assert new_obs[0] >= 3                                    # the whole sequence starts with a compound symbol
    _, pis, Ts, Es = golden_model("isolation_k10")
    want = F.forward_plain(rep, pis[0], Ts[0], Es[0])
    assert abs(F.zip_forward_fast(pis[0], Ts[0], Es[0], s2p, new_obs, 3, ns) - want) <= 1e-12 * abs(want)
    E0 = Es[0].copy()
    E0[:, 1] = 0.0
    bad = rep.copy()
    bad[100] = 1
    new_obs, s2p, ns = F.zip_preprocess(bad, 3, min_count=2, max_syms=16)
    assert F.zip_forward_fast(pis[0], Ts[0], E0, s2p, new_obs, 3, ns) == -np.inf
