"""BASELINE.json configs[1] at FULL size (100 x 1 Mbp, 256 parameter points) through size-independent properties,
plus oracle spot checks -- the oracle cannot score 2.6e10 site-points in a test, the GPU does it in milliseconds."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu
sys.path.insert(0, ROOT)


@pytest.fixture(autouse=True)
def _options():
    import imcoalhmm_b200 as m
    yield
    m.set_option("zip_spectral", 0)
    m.set_option("zip_mma", 0)
    m.set_option("forward_kernel", 0)


@pytest.fixture(scope="module")
def full_c2():
    import bench
    import imcoalhmm_b200 as m
    wl = bench.WORKLOADS["c2"]
    model = m.IsolationModel(10)
    thetas = bench.thetas_around(wl["default"], wl["points"])
    pis, Ts, Es, st = model.build_hidden_markov_models(thetas)
    assert (st == 0).all()
    chunks = bench.make_chunks(wl, pis, Ts, Es, range(wl["chunks"]))
    return m, model, thetas, pis, Ts, Es, chunks


def test_full_size_properties(full_c2):
    m, model, thetas, pis, Ts, Es, chunks = full_c2
    mk = lambda cs: m.ForwarderSet([m.Forwarder.from_symbols(c, 3) for c in cs])
    whole_set = mk(chunks)
    assert whole_set.total_sites == 100_000_000
    info = whole_set.zip_info(10)
    assert info["tokens"] * 100 < whole_set.total_sites            # > 100x compression on this alignment
    whole = whole_set.forward_batch(pis, Ts, Es)
    assert m.last_forward_kernel() == "zip-spectral-mma2-aligned" and np.isfinite(whole).all()     # the automatic choice on this alignment
    assert whole_set.spectral_counts() == (256, 0)
    m.set_option("zip_align", 2)                                   # the lock-step one-run form on the same streams
    try:
        lock_step = whole_set.forward_batch(pis, Ts, Es)
        assert m.last_forward_kernel() == "zip-spectral-mma"
    finally:
        m.set_option("zip_align", 0)
    np.testing.assert_allclose(lock_step, whole, rtol=1e-12)
    assert whole_set.run_info(10)["tokens"] * 250 < whole_set.total_sites  # > 250 sites per mat-vec
    # the FMA shape of the spectral form, then the plain form (pair dictionary, no eigenbasis), on the same 1e8 sites
    m.set_option("zip_mma", 2)
    fma_form = whole_set.forward_batch(pis, Ts, Es)
    assert m.last_forward_kernel() == "zip-spectral"
    m.set_option("zip_mma", 0)
    np.testing.assert_allclose(fma_form, whole, rtol=1e-12)
    m.set_option("zip_spectral", 2)
    plain_form = whole_set.forward_batch(pis, Ts, Es)
    assert m.last_forward_kernel() == "zip"
    m.set_option("zip_spectral", 0)
    np.testing.assert_allclose(plain_form, whole, rtol=1e-11)
    # chunk additivity: two halves have their OWN dictionaries and token streams -> an independent evaluation
    halves = mk(chunks[:37]).forward_batch(pis, Ts, Es) + mk(chunks[37:]).forward_batch(pis, Ts, Es)
    np.testing.assert_allclose(halves, whole, rtol=1e-12)
    # the fused theta -> logL entry gives the same numbers as build + forward
    np.testing.assert_allclose(model.batched_log_likelihood(thetas, whole_set), whole, rtol=1e-12)
    # the uncompressed per-site kernel (every one of the 1e8 sites walked, T in registers) agrees on all 256 points
    m.set_option("forward_kernel", 2)
    try:
        plain = whole_set.forward_batch(pis, Ts, Es)
        assert m.last_forward_kernel() == "pair"
    finally:
        m.set_option("forward_kernel", 0)
    np.testing.assert_allclose(plain, whole, rtol=1e-11)
    # time reversal (the reference's chains are reversible, transitions.py:237) on 10 Mbp
    fw = mk(chunks[:10]).forward_batch(pis[:16], Ts[:16], Es[:16])
    bw = mk([c[::-1].copy() for c in chunks[:10]]).forward_batch(pis[:16], Ts[:16], Es[:16])
    np.testing.assert_allclose(bw, fw, rtol=1e-11)
    # single-point calls take the segmented route and must agree with the batch
    one = whole_set.forward(pis[3], Ts[3], Es[3])
    assert m.last_forward_kernel().startswith("zip-spectral") and m.last_forward_kernel().endswith("segmented")
    assert one == pytest.approx(whole[3], rel=1e-12)


def test_full_size_oracle_spot_check(full_c2):
    m, model, thetas, pis, Ts, Es, chunks = full_c2
    from oracle import forward as F
    sub = [chunks[0], chunks[57], chunks[99]]
    got = m.ForwarderSet([m.Forwarder.from_symbols(c, 3) for c in sub]).forward_batch(pis[:4], Ts[:4], Es[:4])
    want, _ = F.forward_batch([c.astype(np.int32) for c in sub], pis[:4], Ts[:4], Es[:4])
    np.testing.assert_allclose(got, want, rtol=1e-11)


def test_full_size_im_model_shard():
    """configs[2] per-GPU shard (IM model K=10+10, 125 x 1 Mbp) on 256 parameter points: the compressed kernel (4 lanes per
    chain), the per-site DMMA kernel that walks all 1.25e8 sites, chunk additivity over independently preprocessed halves,
    and the warp-per-chain route of single-point calls must all agree."""
    import bench
    import imcoalhmm_b200 as m
    wl = bench.WORKLOADS["c3_1gpu"]
    model = m.IsolationMigrationModel(10, 10)
    thetas = bench.thetas_around(wl["default"], 256)
    pis, Ts, Es, st = model.build_hidden_markov_models(thetas)
    assert (st == 0).all()
    chunks = bench.make_chunks(wl, pis, Ts, Es, range(wl["chunks"]))
    mk = lambda cs: m.ForwarderSet([m.Forwarder.from_symbols(c, 3) for c in cs])
    whole_set = mk(chunks)
    whole = whole_set.forward_batch(pis, Ts, Es)
    assert m.last_forward_kernel() == "zip-spectral-mma2-aligned" and np.isfinite(whole).all()      # the automatic choice at K=20
    m.set_option("zip_align", 2)
    try:
        lock_step = whole_set.forward_batch(pis, Ts, Es)
        assert m.last_forward_kernel() == "zip-spectral-mma"
    finally:
        m.set_option("zip_align", 0)
    np.testing.assert_allclose(lock_step, whole, rtol=1e-12)
    m.set_option("zip_spectral", 2)
    plain_form = whole_set.forward_batch(pis[:64], Ts[:64], Es[:64])
    assert m.last_forward_kernel() == "zip"
    m.set_option("zip_spectral", 0)
    np.testing.assert_allclose(plain_form, whole[:64], rtol=1e-11)
    halves = mk(chunks[:60]).forward_batch(pis, Ts, Es) + mk(chunks[60:]).forward_batch(pis, Ts, Es)
    np.testing.assert_allclose(halves, whole, rtol=1e-12)
    m.set_option("forward_kernel", 3)
    try:
        plain = whole_set.forward_batch(pis[:64], Ts[:64], Es[:64])
        assert m.last_forward_kernel() == "dmma"
    finally:
        m.set_option("forward_kernel", 0)
    np.testing.assert_allclose(plain, whole[:64], rtol=1e-11)
    one = whole_set.forward(pis[5], Ts[5], Es[5])
    assert m.last_forward_kernel() in ("zip-spectral-warp", "zip-spectral-segmented", "zip-spectral-mma-segmented")
    assert one == pytest.approx(whole[5], rel=1e-12)
    np.testing.assert_allclose(model.batched_log_likelihood(thetas[:32], whole_set), whole[:32], rtol=1e-12)
