"""Fewer chunks than GPUs: one long chunk cut into consecutive parts -- every part run in segmented mode, a
later part from the unit vectors (its K x K transfer matrix), the parts folded in order.  On one GPU a single process holds
all parts (same kernels, no collective); with two GPUs the part blocks travel through one ncclAllGather."""
import os
import time

import numpy as np
import pytest

from conftest import golden_model, example_symbols, random_hmm

pytestmark = pytest.mark.gpu
OPTS = ("zip_spectral", "zip_mma", "zip_segment_tokens", "zip_lanes")


@pytest.fixture(autouse=True)
def _reset():
    import imcoalhmm_b200 as m
    yield
    for k in OPTS:
        m.set_option(k, 0)


def cut(obs, bounds):
    return [obs[a:b] for a, b in zip(bounds[:-1], bounds[1:])]


@pytest.mark.parametrize("model", ["isolation_k10", "im_k10_10", "psmc_iso_split_4x10", "isolation_k4"])
def test_parts_of_one_chunk_equal_the_whole_chunk(model):
    import imcoalhmm_b200 as m
    from oracle import forward as F
    obs = example_symbols()
    _, pis, Ts, Es = golden_model(model)
    rng = np.random.default_rng(3)
    pis, Ts, Es = pis[:5].copy(), Ts[:5].copy(), Es[:5].copy()
    K = pis.shape[1]
    Ts[3] = 0.9 * np.eye(K) + 0.1 * rng.dirichlet(np.ones(K), size=K)        # one point the spectral form cannot serve
    want, _ = F.forward_batch([obs.astype(np.int32)], pis, Ts, Es)
    whole = m.ForwarderSet([m.Forwarder.from_symbols(obs, 3)]).forward_batch(pis, Ts, Es)
    np.testing.assert_allclose(whole, want, rtol=1e-11)
    for bounds in ([0, 20000, len(obs)], [0, 1, 30000, 30017, len(obs)], [0, 7000, 14000, 40000, 40001, 50000, 60000, 65000, len(obs)]):
        parts = [m.Forwarder.from_symbols(c, 3) for c in cut(obs, bounds)]
        pset = m.ForwarderSet(parts, parts=(0, len(parts)))
        for opts in (dict(), dict(zip_spectral=2), dict(zip_spectral=1, zip_mma=2), dict(zip_spectral=1, zip_mma=1, zip_segment_tokens=64),
                     dict(zip_spectral=2, zip_segment_tokens=48, zip_lanes=8)):
            for k in OPTS:
                m.set_option(k, opts.get(k, 0))
            got = pset.forward_batch(pis, Ts, Es)
            np.testing.assert_allclose(got, want, rtol=1e-11, err_msg="%s %s %s" % (model, bounds, opts))
            assert pset.forward(pis[1], Ts[1], Es[1]) == pytest.approx(want[1], rel=1e-11)
    # the fused theta -> logL entry on parts
    theta = golden_model(model)[0]
    mk = {"isolation_k10": lambda: m.IsolationModel(10), "im_k10_10": lambda: m.IsolationMigrationModel(10, 10),
          "psmc_iso_split_4x10": lambda: m.VariableCoalescenceRateIsolationModel([4] * 10, True), "isolation_k4": lambda: m.IsolationModel(4)}[model]
    parts = [m.Forwarder.from_symbols(c, 3) for c in cut(obs, [0, 33333, len(obs)])]
    pset = m.ForwarderSet(parts, parts=(0, 2))
    fused = mk().batched_log_likelihood(theta[:4], pset)
    ref, _ = F.forward_batch([obs.astype(np.int32)], *golden_model(model)[1:4])
    np.testing.assert_allclose(fused, ref[:4], rtol=1e-9)


def test_parts_argument_checks():
    import imcoalhmm_b200 as m
    obs = example_symbols()
    a, b = m.Forwarder.from_symbols(obs[:100], 3), m.Forwarder.from_symbols(obs[100:300], 3)
    with pytest.raises(m.IMCError):
        m.ForwarderSet([a, b], parts=(1, 3))              # first part of a rank must be rank * n_local
    with pytest.raises(m.IMCError):
        m.ForwarderSet([a, m.Forwarder.from_symbols(obs[:0], 3)], parts=(0, 2))      # empty part
    _, pis, Ts, Es = golden_model("isolation_k10")
    half = m.ForwarderSet([a], parts=(0, 2))              # needs a communicator of two ranks
    with pytest.raises(m.IMCError):
        half.forward(pis[0], Ts[0], Es[0])


def _rank(rank, world, tmp):
    import imcoalhmm_b200 as m
    m._lib.check(m._lib.load().imc_init(rank))
    idfile = os.path.join(tmp, "id")
    if rank == 0:
        with open(idfile + ".tmp", "wb") as f:
            f.write(m._lib.comm_unique_id())
        os.rename(idfile + ".tmp", idfile)
    else:
        while not os.path.exists(idfile):
            time.sleep(0.1)
    m._lib.comm_init(world, rank, open(idfile, "rb").read())
    obs = example_symbols()
    bounds = [0, 31000, len(obs)]
    part = m.Forwarder.from_symbols(obs[bounds[rank]:bounds[rank + 1]], 3)
    pset = m.ForwarderSet([part], parts=(rank, world))
    theta, pis, Ts, Es = golden_model("im_k10_10")
    out = [pset.forward_batch(pis[:6], Ts[:6], Es[:6]), m.IsolationMigrationModel(10, 10).batched_log_likelihood(theta[:6], pset)]
    m._lib.comm_destroy()
    np.save(os.path.join(tmp, "parts%d.npy" % rank), np.stack(out))


def test_parts_over_two_gpus(tmp_path):
    """One chunk, two ranks, one part each: a single ncclAllGather of the part blocks, then the ordered fold on every rank."""
    import ctypes
    import imcoalhmm_b200 as m
    n = ctypes.c_int()
    m._lib.load().imc_device_count(ctypes.byref(n))
    if n.value < 2:
        pytest.skip("needs two GPUs")
    import multiprocessing as mp
    from oracle import forward as F
    ctx = mp.get_context("spawn")
    procs = [ctx.Process(target=_rank, args=(r, 2, str(tmp_path))) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0
    _, pis, Ts, Es = golden_model("im_k10_10")
    want, _ = F.forward_batch([example_symbols().astype(np.int32)], pis[:6], Ts[:6], Es[:6])
    outs = [np.load(tmp_path / ("parts%d.npy" % r)) for r in range(2)]
    np.testing.assert_array_equal(outs[0], outs[1])                  # the same bits on both ranks
    np.testing.assert_allclose(outs[0][0], want, rtol=1e-11)
    np.testing.assert_allclose(outs[0][1], want, rtol=1e-9)
