"""Spectral form of the compressed kernel (run tokens; C_r diagonalised per parameter point on the device) against
the CPU oracle's plain forward, through the C ABI.  Tolerance 1e-11 relative (north_star: 1e-9)."""
import numpy as np
import pytest

from conftest import golden_model, example_symbols, random_hmm

pytestmark = pytest.mark.gpu
RTOL = 1e-11
OPTS = ("forward_kernel", "zip_ctas_per_sm", "zip_max_entries", "zip_lanes", "zip_segment_tokens", "zip_pipeline", "zip_mma", "zip_run2",
        "zip_align", "zip_spectral", "zip_spectral_force_bad")


@pytest.fixture(autouse=True)
def _reset_options():
    import imcoalhmm_b200 as m
    yield
    for k in OPTS:
        m.set_option(k, 0)


def oracle_batch(chunks, pis, Ts, Es):
    from oracle import forward as F
    out, _ = F.forward_batch([np.asarray(c, dtype=np.int32) for c in chunks], pis, Ts, Es)
    return out


def make_set(chunks, nsym=3):
    import imcoalhmm_b200 as m
    return m.ForwarderSet([m.Forwarder.from_symbols(np.asarray(c), nsym) for c in chunks])


def reversible_hmm(rng, K, S=3, stay=0.97):
    """pi, T with diag(pi) T symmetric (what every model of the reference builds, transitions.py:231-246), E > 0."""
    J = rng.random((K, K)) + 0.05
    J = 0.5 * (J + J.T)
    J *= (1.0 - stay) / J.sum()
    J[np.diag_indices(K)] += stay * rng.dirichlet(np.ones(K) * 3.0)
    J /= J.sum()
    pi = J.sum(axis=1)
    T = J / pi[:, None]
    E = rng.dirichlet(np.ones(S) * 2.0, size=K)
    E[:, 0] = 0.6 + 0.39 * rng.random(K)
    if S == 3:
        E[:, 1] = 1.0 - E[:, 0]
        E[:, 2] = 1.0
    return pi, T, E


def ragged_chunks(rng, p=(0.955, 0.005, 0.04)):
    obs = example_symbols()
    cuts = [0, 1, 2, 18, 5000, 5000, 21017, 40000, len(obs)]
    chunks = [obs[a:b] for a, b in zip(cuts[:-1], cuts[1:])]
    chunks += [rng.choice(3, size=n, p=list(p)).astype(np.uint8) for n in (33, 64, 777, 4096, 15, 16, 17, 30000)]
    chunks.append(np.zeros(9000, dtype=np.uint8))                      # one run longer than 4095 and nothing else
    long_run = rng.choice(3, size=20000, p=list(p)).astype(np.uint8)
    long_run[3000:13000] = 0                                           # continuation tokens in mid-stream
    chunks.append(long_run)
    chunks.append(np.array([1], dtype=np.uint8))
    chunks.append(np.array([2, 0], dtype=np.uint8))
    return chunks


@pytest.mark.parametrize("model", ["isolation_k4", "isolation_k10", "im_epochs_2_3_3", "im_k10_10", "psmc_iso_split_4x10",
                                   "varmig_i12_4x10"])
def test_reference_models_all_shapes(model):
    """Reference-built (pi,T,E): every launch shape of the spectral form agrees with the oracle and serves all points."""
    import imcoalhmm_b200 as m
    rng = np.random.default_rng(7)
    _, pis, Ts, Es = golden_model(model)
    K = pis.shape[1]
    chunks = ragged_chunks(rng)
    want = oracle_batch(chunks, pis, Ts, Es)
    s = make_set(chunks)
    m.set_option("zip_spectral", 1)
    shapes = [dict(), dict(zip_segment_tokens=-1, zip_pipeline=1), dict(zip_segment_tokens=-1, zip_pipeline=5),
              dict(zip_segment_tokens=256), dict(zip_segment_tokens=-1, zip_lanes=8), dict(zip_segment_tokens=-1, zip_lanes=4)]
    if K >= 10:
        shapes.append(dict(zip_segment_tokens=-1, zip_lanes=32))
    if K <= 24:
        shapes.append(dict(zip_segment_tokens=-1, zip_ctas_per_sm=2))
    shapes.append(dict(zip_max_entries=5, zip_segment_tokens=-1))
    shapes = [dict(o, zip_mma=2) for o in shapes]                      # the FMA shapes ...
    if K >= 7:                                                        # ... and the MMA form (tiles >= 8): whole chunks, pieces, segments
        shapes += [dict(zip_mma=1, zip_segment_tokens=-1, zip_pipeline=1), dict(zip_mma=1, zip_segment_tokens=-1, zip_pipeline=5),
                   dict(zip_mma=1, zip_segment_tokens=256), dict(zip_mma=1, zip_segment_tokens=-1, zip_max_entries=4), dict(zip_mma=0)]
        shapes = [dict(o, zip_run2=2) for o in shapes]
        # the two-run form (missing-data blocks through B^-1, Lambda2^m, B): whole chunks, pieces, segments, a tiny dictionary
        shapes += [dict(zip_mma=1, zip_run2=1, zip_segment_tokens=-1, zip_pipeline=1), dict(zip_mma=1, zip_run2=1, zip_segment_tokens=-1, zip_pipeline=5),
                   dict(zip_mma=1, zip_run2=1, zip_segment_tokens=256), dict(zip_mma=1, zip_run2=1, zip_segment_tokens=-1, zip_max_entries=5),
                   dict(zip_run2=0)]
        # the aligned form (one dictionary entry per warp-step, host-built schedules): whole chunks, pieces, a tiny dictionary
        shapes += [dict(zip_mma=1, zip_align=1, zip_segment_tokens=-1, zip_pipeline=1), dict(zip_mma=1, zip_align=1, zip_segment_tokens=-1, zip_pipeline=5),
                   dict(zip_mma=1, zip_align=1, zip_segment_tokens=-1, zip_pipeline=32), dict(zip_mma=1, zip_align=1, zip_segment_tokens=-1, zip_max_entries=6)]
    for opts in shapes:
        for k in OPTS[1:9]:
            m.set_option(k, opts.get(k, 0))
        got = s.forward_batch(pis, Ts, Es)
        assert m.last_forward_kernel().startswith("zip-spectral"), m.last_forward_kernel()
        if opts.get("zip_mma") == 1:
            assert "mma" in m.last_forward_kernel(), m.last_forward_kernel()
        if opts.get("zip_run2") == 1:
            assert "mma2" in m.last_forward_kernel(), m.last_forward_kernel()
        if opts.get("zip_align") == 1:
            assert m.last_forward_kernel() == "zip-spectral-mma2-aligned", m.last_forward_kernel()
        np.testing.assert_allclose(got, want, rtol=RTOL, err_msg="%s %s" % (model, opts))
        assert s.spectral_counts() == (len(pis), 0)
        one = s.forward(pis[-1], Ts[-1], Es[-1])
        assert one == pytest.approx(want[-1], rel=RTOL)


@pytest.mark.parametrize("K", [1, 2, 3, 5, 6, 8, 9, 12, 13, 16, 20, 24, 31, 32, 40])
def test_random_reversible_hmms_every_tile(K):
    import imcoalhmm_b200 as m
    rng = np.random.default_rng(500 + K)
    hmms = [reversible_hmm(rng, K) for _ in range(6)]
    pis, Ts, Es = (np.stack([h[i] for h in hmms]) for i in range(3))
    chunks = [rng.choice(3, size=int(n), p=[0.93, 0.03, 0.04]).astype(np.uint8) for n in (3000, 17, 1, 64, 5000, 70000, 2, 129)]
    want = oracle_batch(chunks, pis, Ts, Es)
    s = make_set(chunks)
    m.set_option("zip_spectral", 1)
    for lanes in (4, 8):
        m.set_option("zip_lanes", lanes)
        np.testing.assert_allclose(s.forward_batch(pis, Ts, Es), want, rtol=RTOL, err_msg="K=%d lanes=%d" % (K, lanes))
        assert s.spectral_counts() == (6, 0)
    m.set_option("zip_lanes", 0)
    if K >= 7:
        m.set_option("zip_mma", 1)
        for seg, r2 in ((-1, 2), (128, 2), (-1, 1), (128, 1)):
            m.set_option("zip_segment_tokens", seg)
            m.set_option("zip_run2", r2)
            np.testing.assert_allclose(s.forward_batch(pis, Ts, Es), want, rtol=RTOL, err_msg="K=%d mma seg=%d run2=%d" % (K, seg, r2))
            assert "mma" in m.last_forward_kernel() and s.spectral_counts() == (6, 0)
            assert ("mma2" in m.last_forward_kernel()) == (r2 == 1)
        m.set_option("zip_run2", 0)
        m.set_option("zip_align", 1)
        for pipe in (1, 4):         # aligned form: whole streams, pieces
            m.set_option("zip_segment_tokens", -1)
            m.set_option("zip_pipeline", pipe)
            np.testing.assert_allclose(s.forward_batch(pis, Ts, Es), want, rtol=RTOL, err_msg="K=%d aligned pipe=%d" % (K, pipe))
            assert m.last_forward_kernel() == "zip-spectral-mma2-aligned" and s.spectral_counts() == (6, 0)
        m.set_option("zip_align", 0)
        m.set_option("zip_pipeline", 0)
        m.set_option("zip_mma", 0)
        m.set_option("zip_run2", 0)
        m.set_option("zip_segment_tokens", 0)
    np.testing.assert_allclose(s.forward_batch(pis[:1], Ts[:1], Es[:1]), want[:1], rtol=RTOL)     # chain-scarce shapes


def test_mixed_batch_reversible_and_not():
    """Points whose diag(pi) T is not symmetric (or whose pi has a zero) go to the plain form inside the same call."""
    import imcoalhmm_b200 as m
    rng = np.random.default_rng(11)
    K = 10
    hmms = []
    for i in range(12):
        hmms.append(random_hmm(rng, K) if i % 3 == 1 else reversible_hmm(rng, K))
    pis, Ts, Es = (np.stack([h[i] for h in hmms]) for i in range(3))
    Ts = 0.9 * np.eye(K)[None] + 0.1 * Ts        # still reversible where it was (pi_i delta_ij is symmetric)
    pis[3, 0] = 0.0                              # reversible but with an unreachable start state: not diagonalisable our way
    chunks = [rng.choice(3, size=int(n), p=[0.95, 0.01, 0.04]).astype(np.uint8) for n in (30000, 17, 1, 64, 5000, 70000)]
    want = oracle_batch(chunks, pis, Ts, Es)
    s = make_set(chunks)
    m.set_option("zip_spectral", 1)
    for seg, mma, r2 in ((-1, 2, 0), (0, 2, 0), (512, 2, 0), (-1, 1, 2), (512, 1, 2), (-1, 1, 1), (512, 1, 1), (0, 0, 0)):
        m.set_option("zip_segment_tokens", seg)
        m.set_option("zip_mma", mma)
        m.set_option("zip_run2", r2)
        np.testing.assert_allclose(s.forward_batch(pis, Ts, Es), want, rtol=RTOL, err_msg="seg=%d mma=%d" % (seg, mma))
        ok, plain = s.spectral_counts()
        assert (ok, plain) == (7, 5)
    m.set_option("zip_spectral_force_bad", 1)
    np.testing.assert_allclose(s.forward_batch(pis, Ts, Es), want, rtol=RTOL)
    assert s.spectral_counts() == (0, 12)


def test_run_symbol_is_not_zero_and_larger_alphabets():
    import imcoalhmm_b200 as m
    rng = np.random.default_rng(13)
    m.set_option("zip_spectral", 1)
    for S, K in ((3, 6), (5, 8), (9, 4)):
        hmms = [reversible_hmm(rng, K, S) for _ in range(3)]
        pis, Ts, Es = (np.stack([h[i] for h in hmms]) for i in range(3))
        p = np.full(S, 0.08 / (S - 1))
        p[S - 1] = 0.92                                            # the last symbol is the frequent one
        chunks = [rng.choice(S, size=int(n), p=p).astype(np.uint8) for n in (20000, 300, 5, 1)]
        want = oracle_batch(chunks, pis, Ts, Es)
        s = make_set(chunks, S)
        assert s.run_info()["run_sym"] == S - 1
        np.testing.assert_allclose(s.forward_batch(pis, Ts, Es), want, rtol=RTOL, err_msg="S=%d" % S)
        assert s.spectral_counts() == (3, 0)


def test_auto_choice_and_fused_model_path():
    """Auto: run tokens win on alignment-like data, the plain dictionary on symbols without runs; the fused
    theta -> logL path goes through the spectral form and matches reference-built (pi,T,E) + oracle."""
    import imcoalhmm_b200 as m
    rng = np.random.default_rng(17)
    theta, pis, Ts, Es = golden_model("isolation_k10")
    obs = example_symbols()
    chunks = [obs[i * 4000:(i + 1) * 4000] for i in range(16)] * 8
    s = make_set(chunks)
    want = oracle_batch(chunks, pis, Ts, Es)
    got = s.forward_batch(pis, Ts, Es)
    np.testing.assert_allclose(got, want, rtol=RTOL)
    assert m.last_forward_kernel().startswith("zip-spectral")
    fused = m.IsolationModel(10).batched_log_likelihood(theta, s)
    np.testing.assert_allclose(fused, want, rtol=1e-9)
    assert m.last_forward_kernel().startswith("zip-spectral")
    noise = [rng.integers(0, 3, size=5000).astype(np.uint8) for _ in range(64)]
    s2 = make_set(noise)
    s2.forward_batch(pis, Ts, Es)
    assert not m.last_forward_kernel().startswith("zip-spectral")


def test_impossible_and_degenerate_inputs():
    """An impossible observation gives -inf, NaN parameters give NaN, in the spectral form as in the plain one."""
    import imcoalhmm_b200 as m
    rng = np.random.default_rng(19)
    K = 6
    pi, T, E = reversible_hmm(rng, K)
    E2 = E.copy()
    E2[:, 1] = 0.0                      # symbol 1 can never be emitted
    chunks = [np.array([0] * 50 + [1] + [0] * 50, dtype=np.uint8), np.zeros(40, dtype=np.uint8)]
    s = make_set(chunks)
    m.set_option("zip_spectral", 1)
    m.set_option("zip_segment_tokens", -1)
    assert s.forward(pi, T, E2) == -np.inf
    Tn = T.copy()
    Tn[0, 0] = np.nan
    assert np.isnan(s.forward(pi, Tn, E))
    assert np.isfinite(s.forward(pi, T, E))
