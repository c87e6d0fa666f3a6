"""Parity of the CUDA forward kernels with the CPU oracle, through the C ABI.

Tolerance: north_star asks for 1e-9 relative; the kernels are held to 1e-11 here (expected ~1e-13).
"""
import numpy as np
import pytest

from conftest import golden_model, example_symbols, random_hmm, synthetic_sequence

pytestmark = pytest.mark.gpu
RTOL = 1e-11

KERNELS = {"generic": 1, "pair": 2, "dmma": 3, "zip": 4}
ZIP_MAX_K = 40      # tiles 2..40; any K below runs in the next tile with zero padding


@pytest.fixture(autouse=True)
def _reset_options():
    """These are the tests of the plain (pair-dictionary) form and of the per-site kernels; the spectral form over run
    tokens, which the automatic choice prefers on alignment-like data, has tests/test_spectral_gpu.py."""
    import imcoalhmm_b200 as m
    m.set_option("zip_spectral", 2)
    yield
    m.set_option("zip_spectral", 0)
    m.set_option("forward_kernel", 0)
    m.set_option("dmma_mtiles", 0)
    m.set_option("fold_emission", 0)
    m.set_option("zip_ctas_per_sm", 0)
    m.set_option("zip_max_entries", 0)
    m.set_option("zip_lanes", 0)
    m.set_option("zip_segment_tokens", 0)
    m.set_option("zip_pipeline", 0)


def oracle_batch(chunks, pis, Ts, Es):
    from oracle import forward as F
    out, _ = F.forward_batch([np.asarray(c, dtype=np.int32) for c in chunks], pis, Ts, Es)
    return out


def make_set(chunks, nsym=3):
    import imcoalhmm_b200 as m
    return m.ForwarderSet([m.Forwarder.from_symbols(np.asarray(c), nsym) for c in chunks])


def test_example_alignment_reference_models():
    """config 1: hg18 vs pantro2, reference-built (pi,T,E), single evaluation (SURVEY 7.1c probe values)."""
    import imcoalhmm_b200 as m
    obs = example_symbols()
    f = m.Forwarder.from_symbols(obs, 3)
    _, pi, T, E = golden_model("isolation_k10")
    assert f.forward(pi[0], T[0], E[0]) == pytest.approx(-3729.5586472699, rel=1e-11)
    assert m.last_forward_kernel() in ("zip-segmented", "zip-warp")      # one chunk x one point: chain-scarce
    _, pi, T, E = golden_model("im_k10_10")
    assert f.forward(pi[0], T[0], E[0]) == pytest.approx(-3650.0493084297, rel=1e-11)
    assert m.last_forward_kernel() in ("zip-segmented", "zip-warp")
    m.set_option("zip_segment_tokens", -1)
    m.set_option("zip_lanes", 4)
    assert f.forward(pi[0], T[0], E[0]) == pytest.approx(-3650.0493084297, rel=1e-11)
    assert m.last_forward_kernel() == "zip"
    m.set_option("zip_lanes", 32)
    assert f.forward(pi[0], T[0], E[0]) == pytest.approx(-3650.0493084297, rel=1e-11)
    assert m.last_forward_kernel() == "zip-warp"
    m.set_option("zip_lanes", 0)
    for code, name in ((2, "pair"), (3, "dmma")):
        m.set_option("forward_kernel", code)
        _, pi, T, E = golden_model("isolation_k10")
        assert f.forward(pi[0], T[0], E[0]) == pytest.approx(-3729.5586472699, rel=1e-11)
        assert m.last_forward_kernel() == name


@pytest.mark.parametrize("model,kernels", [
    ("isolation_k4", ["generic", "pair", "zip"]),
    ("isolation_k10", ["generic", "pair", "dmma", "zip"]),
    ("im_k3_4", ["generic"]),
    ("im_epochs_2_3_3", ["generic", "pair", "dmma", "zip"]),
    ("im_k10_10", ["generic", "dmma", "zip"]),
    ("psmc_iso_split_4x10", ["generic", "dmma", "zip"]),
    ("varmig_i12_4x10", ["generic", "dmma", "zip"]),
])
def test_batch_parity_on_reference_models(model, kernels):
    import imcoalhmm_b200 as m
    rng = np.random.default_rng(20261018)
    _, pis, Ts, Es = golden_model(model)
    obs = example_symbols()
    # ragged chunks: includes a length-1 chunk, a chunk that ends mid-word, an empty chunk
    cuts = [0, 1, 18, 5000, 5000, 21017, 40000, len(obs)]
    chunks = [obs[a:b] for a, b in zip(cuts[:-1], cuts[1:])]
    chunks += [rng.choice(3, size=n, p=[0.9, 0.05, 0.05]).astype(np.uint8) for n in (33, 64, 777, 4096, 15, 16, 17)]
    want = oracle_batch(chunks, pis, Ts, Es)
    s = make_set(chunks)
    for k in kernels:
        m.set_option("forward_kernel", KERNELS[k])
        got = s.forward_batch(pis, Ts, Es)
        assert m.last_forward_kernel().split("-")[0] == k     # "zip-segmented" when the call is chain-scarce
        np.testing.assert_allclose(got, want, rtol=RTOL, err_msg="%s / %s" % (model, k))
        one = s.forward(pis[-1], Ts[-1], Es[-1])
        assert one == pytest.approx(want[-1], rel=RTOL)


@pytest.mark.parametrize("K", [2, 6, 8, 12, 16, 24, 28, 32, 36, 48, 64])
def test_random_hmms_all_instantiated_sizes(K):
    import imcoalhmm_b200 as m
    rng = np.random.default_rng(K)
    N = 5
    hmms = [random_hmm(rng, K) for _ in range(N)]
    pis, Ts, Es = (np.stack([h[i] for h in hmms]) for i in range(3))
    chunks = [rng.integers(0, 3, size=n).astype(np.uint8) for n in (1, 2, 31, 32, 33, 500, 1000, 1000, 999, 47, 2048)]
    want = oracle_batch(chunks, pis, Ts, Es)
    s = make_set(chunks)
    for k, code in KERNELS.items():
        if k == "pair" and (K % 2 or K > 12):
            continue
        if k == "dmma" and K not in (10, 12, 16, 20, 24, 28, 32, 36, 40, 48, 64):
            continue
        if k == "zip" and K > ZIP_MAX_K:
            continue
        m.set_option("forward_kernel", code)
        for lanes in ((4, 8) if k == "zip" else (0,)):
            m.set_option("zip_lanes", lanes)
            np.testing.assert_allclose(s.forward_batch(pis, Ts, Es), want, rtol=RTOL, err_msg="K=%d %s %d" % (K, k, lanes))


@pytest.mark.parametrize("K", [3, 5, 7, 9, 11, 13, 20, 40, 65, 100, 128])
def test_generic_kernel_any_K(K):
    import imcoalhmm_b200 as m
    rng = np.random.default_rng(100 + K)
    pi, T, E = random_hmm(rng, K)
    chunks = [rng.integers(0, 3, size=n).astype(np.uint8) for n in (300, 17, 64)]
    want = oracle_batch(chunks, pi[None], T[None], E[None])[0]
    m.set_option("forward_kernel", 1)
    assert make_set(chunks).forward(pi, T, E) == pytest.approx(want, rel=RTOL)


@pytest.mark.parametrize("K", [1, 7, 9, 11, 13, 14, 15, 18, 23, 30, 37])
def test_zip_kernel_any_K_up_to_40(K):
    """State counts between the instantiated tiles run zero-padded in the next tile (both lane decompositions)."""
    import imcoalhmm_b200 as m
    rng = np.random.default_rng(300 + K)
    hmms = [random_hmm(rng, K) for _ in range(3)]
    pis, Ts, Es = (np.stack([h[i] for h in hmms]) for i in range(3))
    Ts = 0.9 * np.eye(K)[None] + 0.1 * Ts
    chunks = [rng.choice(3, size=int(n), p=[0.95, 0.01, 0.04]).astype(np.uint8) for n in (3000, 17, 1, 64, 5000, 70000)]
    want = oracle_batch(chunks, pis, Ts, Es)
    s = make_set(chunks)
    assert s.forward(pis[0], Ts[0], Es[0]) == pytest.approx(want[0], rel=RTOL)
    assert m.last_forward_kernel().startswith("zip")               # the default for every K <= 40
    for lanes in (8, 4, 32):
        for seg in (-1, 64):
            m.set_option("zip_lanes", lanes)
            m.set_option("zip_segment_tokens", seg)
            np.testing.assert_allclose(s.forward_batch(pis, Ts, Es), want, rtol=RTOL, err_msg="K=%d lanes=%d seg=%d" % (K, lanes, seg))


@pytest.mark.parametrize("mt", [1, 2, 4])
def test_dmma_mtile_variants(mt):
    import imcoalhmm_b200 as m
    rng = np.random.default_rng(mt)
    _, pis, Ts, Es = golden_model("im_k10_10")
    chunks = [rng.choice(3, size=int(n), p=[0.95, 0.01, 0.04]).astype(np.uint8) for n in rng.integers(900, 1100, size=70)]
    want = oracle_batch(chunks, pis, Ts, Es)
    m.set_option("forward_kernel", 3)
    m.set_option("dmma_mtiles", mt)
    np.testing.assert_allclose(make_set(chunks).forward_batch(pis, Ts, Es), want, rtol=RTOL)


def test_emission_folding_on_off_and_unfoldable():
    """The pair kernel folds E[:,s0] of the most frequent symbol into T; results must not depend on it,
    and parameter points with a zero in that column (cannot be folded) must take the unfolded loop."""
    import imcoalhmm_b200 as m
    rng = np.random.default_rng(99)
    _, pis, Ts, Es = golden_model("isolation_k10")
    chunks = [rng.choice(3, size=n, p=p).astype(np.uint8)
              for n, p in ((3000, [0.95, 0.01, 0.04]), (2000, [0.1, 0.8, 0.1]), (1000, [0.3, 0.3, 0.4]), (17, [1, 0, 0]))]
    want = oracle_batch(chunks, pis, Ts, Es)
    s = make_set(chunks)
    m.set_option("forward_kernel", 2)
    for fold in (1, 0):
        m.set_option("fold_emission", fold)
        np.testing.assert_allclose(s.forward_batch(pis, Ts, Es), want, rtol=RTOL)
    # most frequent symbol is 1 here -> s0 = 1
    s1 = make_set([chunks[1]])
    m.set_option("fold_emission", 1)
    np.testing.assert_allclose(s1.forward_batch(pis, Ts, Es), oracle_batch([chunks[1]], pis, Ts, Es), rtol=RTOL)
    # unfoldable: some states cannot emit symbol 0 at all
    E2 = Es.copy()
    E2[:, 3, 0] = 0.0
    E2[:, 7, 0] = 0.0
    np.testing.assert_allclose(s.forward_batch(pis, Ts, E2), oracle_batch(chunks, pis, Ts, E2), rtol=RTOL)
    # mixed batch: only some parameter points unfoldable (lanes of one warp disagree)
    E3 = Es.copy()
    E3[::2, 1, 0] = 0.0
    np.testing.assert_allclose(s.forward_batch(pis, Ts, E3), oracle_batch(chunks, pis, Ts, E3), rtol=RTOL)


def test_edge_cases():
    import imcoalhmm_b200 as m
    _, pis, Ts, Es = golden_model("isolation_k10")
    # all-missing: logL = log(sum(pi)) = 0 (emissions.py:99); empty set / empty chunks: 0
    s = make_set([np.full(1000, 2, dtype=np.uint8)])
    assert abs(s.forward(pis[0], Ts[0], Es[0])) < 1e-11
    assert make_set([np.zeros(0, dtype=np.uint8)]).forward(pis[0], Ts[0], Es[0]) == 0.0
    assert m.ForwarderSet([]).forward(pis[0], Ts[0], Es[0]) == 0.0
    # single site
    for o in range(3):
        got = make_set([np.array([o], dtype=np.uint8)]).forward(pis[0], Ts[0], Es[0])
        assert got == pytest.approx(np.log(pis[0] @ Es[0][:, o]), rel=1e-13)
    # an impossible observation gives -inf, not NaN
    E0 = Es[0].copy()
    E0[:, 1] = 0.0
    for code in (1, 2, 4):
        m.set_option("forward_kernel", code)
        assert make_set([np.array([0, 0, 1, 0] * 10, dtype=np.uint8)]).forward(pis[0], Ts[0], E0) == -np.inf
    m.set_option("forward_kernel", 0)
    # shape errors are loud
    with pytest.raises(ValueError):
        s.forward(pis[0][:5], Ts[0], Es[0])
    with pytest.raises(m.IMCError):
        s.forward(pis[0], Ts[0], Es[0][:, :2])       # S != NSYM of the sequences
    # (K,1) column vector for pi as pyZipHMM.Matrix users pass it
    assert s.forward(pis[0].reshape(-1, 1), Ts[0], Es[0]) == pytest.approx(s.forward(pis[0], Ts[0], Es[0]))


def test_properties_at_scale():
    """Size-independent properties on inputs too large for a quick oracle pass: chunk additivity,
    time-reversal invariance (reversible reference models), kernel-vs-kernel agreement."""
    import imcoalhmm_b200 as m
    rng = np.random.default_rng(7)
    _, pis, Ts, Es = golden_model("isolation_k10")
    obs = synthetic_sequence(rng, pis[0], Ts[0], Es[0], 400_000)
    whole = make_set([obs]).forward_batch(pis[:8], Ts[:8], Es[:8])
    rev = make_set([obs[::-1].copy()]).forward_batch(pis[:8], Ts[:8], Es[:8])
    np.testing.assert_allclose(rev, whole, rtol=1e-11)
    m.set_option("forward_kernel", 1)
    np.testing.assert_allclose(make_set([obs]).forward_batch(pis[:8], Ts[:8], Es[:8]), whole, rtol=1e-11)
    m.set_option("forward_kernel", 3)
    np.testing.assert_allclose(make_set([obs]).forward_batch(pis[:8], Ts[:8], Es[:8]), whole, rtol=1e-11)
    m.set_option("forward_kernel", 0)
    # the oracle on the same input (0.4 Mbp x 8 points: ~1 s of CPU)
    np.testing.assert_allclose(whole, oracle_batch([obs], pis[:8], Ts[:8], Es[:8]), rtol=RTOL)


@pytest.mark.parametrize("K", [3, 5, 8, 10, 12, 20, 40])
def test_zip_kernel_configurations(K):
    """The compressed kernel under every launch shape: resident CTAs per SM (256 / 512 threads), dictionary caps
    (3 = base symbols only, i.e. no compression), on compressible and on incompressible data."""
    import imcoalhmm_b200 as m
    rng = np.random.default_rng(500 + K)
    N = 3
    hmms = [random_hmm(rng, K) for _ in range(N)]
    pis, Ts, Es = (np.stack([h[i] for h in hmms]) for i in range(3))
    # keep transitions sticky so that long compressed runs stay well conditioned
    Ts = 0.9 * np.eye(K)[None] + 0.1 * Ts
    chunks = [rng.choice(3, size=int(n), p=[0.95, 0.01, 0.04]).astype(np.uint8) for n in rng.integers(2000, 6000, size=37)]
    chunks += [rng.integers(0, 3, size=n).astype(np.uint8) for n in (1, 2, 3, 16, 17, 18, 300)]
    chunks += [np.zeros(70000, dtype=np.uint8), np.zeros(0, dtype=np.uint8)]
    want = oracle_batch(chunks, pis, Ts, Es)
    s = make_set(chunks)
    m.set_option("forward_kernel", 4)
    for lanes, ctas, cap in ((0, 0, 0), (8, 1, 0), (8, 2, 16), (4, 1, 3), (4, 2, 4), (4, 0, 64), (8, 0, 5), (4, 1, 0), (32, 0, 0), (32, 0, 7)):
        m.set_option("zip_lanes", lanes)
        m.set_option("zip_ctas_per_sm", ctas)
        m.set_option("zip_max_entries", cap)
        got = s.forward_batch(pis, Ts, Es)
        assert m.last_forward_kernel() .startswith("zip")
        np.testing.assert_allclose(got, want, rtol=RTOL, err_msg="K=%d lanes=%d ctas=%d cap=%d" % (K, lanes, ctas, cap))


def test_zip_kernel_work_stealing_many_points():
    """More parameter points than persistent CTAs, and fewer: every (point, chunk) pair is scored exactly once."""
    import imcoalhmm_b200 as m
    rng = np.random.default_rng(77)
    _, pis, Ts, Es = golden_model("isolation_k10")
    chunks = [rng.choice(3, size=int(n), p=[0.95, 0.01, 0.04]).astype(np.uint8) for n in rng.integers(50, 3000, size=45)]
    want16 = oracle_batch(chunks, pis, Ts, Es)
    s = make_set(chunks)
    reps = 44                                             # 704 points > 296 resident CTAs
    big = [np.tile(x, (reps,) + (1,) * (x.ndim - 1)) for x in (pis, Ts, Es)]
    for lanes, ctas in ((8, 1), (8, 2), (4, 1), (4, 2)):
        m.set_option("zip_lanes", lanes)
        m.set_option("zip_ctas_per_sm", ctas)
        got = s.forward_batch(*big)
        np.testing.assert_allclose(got, np.tile(want16, reps), rtol=RTOL)
        np.testing.assert_allclose(s.forward_batch(pis[:3], Ts[:3], Es[:3]), want16[:3], rtol=RTOL)


@pytest.mark.parametrize("model,lanes", [("isolation_k10", 8), ("im_k10_10", 4), ("im_k10_10", 32), ("isolation_k4", 8)])
def test_zip_pipelined_pieces_are_bit_identical(model, lanes):
    """Pipelined mode: a chunk is walked in pieces that are separate, ordered work units handing their state on through
    global memory.  Same operations in the same order, so the result must not change by a single bit -- ragged chunks,
    chunks shorter than a piece, a chain that dies half way, more points than CTAs."""
    import imcoalhmm_b200 as m
    rng = np.random.default_rng(5)
    _, pis, Ts, Es = golden_model(model)
    chunks = [rng.choice(3, size=int(n), p=[0.6, 0.3, 0.1]).astype(np.uint8) for n in rng.integers(1, 20000, size=37)]
    chunks.append(rng.choice(3, size=30000, p=[0.6, 0.3, 0.1]).astype(np.uint8))
    Es = Es.copy()
    P, dead = pis.shape[0], pis.shape[0] - 1
    Es[dead, :, 1] = 0.0                                     # last point: symbol 1 is impossible -> -inf, carried across pieces
    want = oracle_batch(chunks, pis, Ts, Es)
    assert np.isfinite(want[:dead]).all()                    # (the oracle's plain recursion gives 0/0 = nan for the dead point)
    s = make_set(chunks)
    reps = 320 // P                                          # 320 points > 148 or 296 CTAs
    big = [np.tile(x, (reps,) + (1,) * (x.ndim - 1)) for x in (pis, Ts, Es)]
    m.set_option("forward_kernel", 4)                        # (these symbols hardly compress: auto would pick a per-site kernel)
    m.set_option("zip_lanes", lanes)
    m.set_option("zip_segment_tokens", -1)
    m.set_option("zip_pipeline", 1)
    base = s.forward_batch(*big)
    assert m.last_forward_kernel() in ("zip", "zip-warp")
    base2 = base.reshape(reps, P)
    np.testing.assert_allclose(base2[:, :dead], np.tile(want[:dead], (reps, 1)), rtol=RTOL)
    assert np.isneginf(base2[:, dead]).all()
    for pieces in (2, 5, 32, 0):
        m.set_option("zip_pipeline", pieces)
        np.testing.assert_array_equal(s.forward_batch(*big), base)
        np.testing.assert_array_equal(s.forward_batch(pis[:3], Ts[:3], Es[:3]), base[:3])


@pytest.mark.parametrize("model", ["isolation_k10", "im_k10_10", "isolation_k4"])
def test_zip_segmented_mode(model):
    """Chain-scarce calls cut chunks into segments whose transfer matrices are built column by column and folded
    afterwards (the associative form of the forward recursion): same logL as the sequential pass."""
    import imcoalhmm_b200 as m
    rng = np.random.default_rng(5)
    _, pis, Ts, Es = golden_model(model)
    obs = synthetic_sequence(rng, pis[0], Ts[0], Es[0], 300_000)
    chunks = [obs, obs[:70_001], obs[100_000:100_700], obs[:1], np.zeros(0, dtype=np.uint8)]
    want = oracle_batch(chunks, pis[:3], Ts[:3], Es[:3])
    s = make_set(chunks)
    m.set_option("zip_segment_tokens", -1)
    base = s.forward_batch(pis[:3], Ts[:3], Es[:3])
    assert m.last_forward_kernel() in ("zip", "zip-warp")
    np.testing.assert_allclose(base, want, rtol=RTOL)
    for lanes in (8, 4):
        m.set_option("zip_lanes", lanes)
        for seg in (0, 16, 48, 256, 1000):
            m.set_option("zip_segment_tokens", seg)
            got = s.forward_batch(pis[:3], Ts[:3], Es[:3])
            assert m.last_forward_kernel() == "zip-segmented", (lanes, seg)
            np.testing.assert_allclose(got, want, rtol=RTOL, err_msg="lanes=%d seg=%d" % (lanes, seg))
            assert s.forward(pis[1], Ts[1], Es[1]) == pytest.approx(want[1], rel=RTOL)
    # many points: auto mode goes back to the plain zip kernel
    m.set_option("zip_segment_tokens", 0)
    reps = -(-6000 // len(pis))                           # 6000 points x 4 chunks: far more chains than chain slots
    big = [np.tile(x, (reps,) + (1,) * (x.ndim - 1)) for x in (pis, Ts, Es)]
    s.forward_batch(*big)
    assert m.last_forward_kernel() == "zip"
    # an impossible observation inside a later segment still gives -inf
    E0 = Es[0].copy()
    E0[:, 1] = 0.0
    bad = obs.copy()
    bad[250_000] = 1
    m.set_option("zip_segment_tokens", 64)
    assert make_set([bad]).forward(pis[0], Ts[0], E0) == -np.inf
    assert np.isfinite(make_set([np.where(obs == 1, 0, obs)]).forward(pis[0], Ts[0], E0))


@pytest.mark.parametrize("nsym", [1, 2, 4, 9])
def test_zip_kernel_other_alphabets(nsym):
    """NSYM other than 3 (the ILS scripts use alphabetSize=9, scripts/prepare-alignments.py:201) runs on the zip kernel."""
    import imcoalhmm_b200 as m
    rng = np.random.default_rng(nsym)
    K = 6
    pi, T, E = random_hmm(rng, K, S=nsym, missing_col=False)
    chunks = [rng.integers(0, nsym, size=n).astype(np.int32) for n in (1, 50, 1000, 4097)]
    want = oracle_batch(chunks, pi[None], T[None], E[None])[0]
    got = make_set(chunks, nsym).forward(pi, T, E)
    assert m.last_forward_kernel() .startswith("zip")
    assert got == pytest.approx(want, rel=RTOL)


def test_ziphmm_module_contract():
    """The two calls the reference's hmm.py makes (hmm.py:16, :20-21)."""
    from imcoalhmm_b200 import ziphmm
    from oracle import forward as F
    _, pis, Ts, Es = golden_model("isolation_k10")
    obs = example_symbols().astype(np.int32)[:10000]
    new_obs, sym2pair, new_nsyms = ziphmm.preprocess_raw_observations(obs, 3)
    got = ziphmm.zip_forward(pis[0], Ts[0], Es[0], sym2pair, new_obs, 3, new_nsyms)
    want = F.forward_plain(obs, pis[0], Ts[0], Es[0])
    assert got == pytest.approx(want, rel=RTOL)
    # a genuinely pair-compressed encoding produced elsewhere is accepted too
    z_obs, z_pairs, z_n = F.zip_preprocess(obs, 3)
    assert ziphmm.zip_forward(pis[0], Ts[0], Es[0], z_pairs, z_obs, 3, z_n) == pytest.approx(want, rel=RTOL)


def test_auto_selection_prefers_per_site_kernels_on_incompressible_data():
    """Random-looking symbols do not compress; with many chains the per-site kernels are the faster route and the
    automatic choice switches to them (same numbers either way)."""
    import imcoalhmm_b200 as m
    rng = np.random.default_rng(11)
    _, pis, Ts, Es = golden_model("isolation_k10")
    chunks = [rng.integers(0, 3, size=3000).astype(np.uint8) for _ in range(300)]
    s = make_set(chunks)
    assert s.zip_info(10)["tokens"] * 5 > s.total_sites
    got = s.forward_batch(pis, Ts, Es)                      # 16 points x 300 chunks = 4800 chains
    assert m.last_forward_kernel() == "pair"
    m.set_option("forward_kernel", 4)
    np.testing.assert_allclose(s.forward_batch(pis, Ts, Es), got, rtol=1e-11)
    assert m.last_forward_kernel().startswith("zip")
    m.set_option("forward_kernel", 0)
    np.testing.assert_allclose(got, oracle_batch(chunks, pis, Ts, Es), rtol=RTOL)
    comp = make_set([rng.choice(3, size=3000, p=[0.97, 0.01, 0.02]).astype(np.uint8) for _ in range(300)])
    comp.forward_batch(pis, Ts, Es)
    assert m.last_forward_kernel().startswith("zip")       # compressible: stays on the zip kernel
