"""The library's own all-reduce (imc_comm_init): two processes, one GPU each, chunks sharded between them;
every forward / likelihood entry point must return the sum over ranks.  Needs two GPUs (skipped otherwise)."""
import os
import time

import numpy as np
import pytest

from conftest import golden_model, example_symbols

pytestmark = pytest.mark.gpu


def _connect(m, rank, world, idfile):
    if rank == 0:
        uid = m._lib.comm_unique_id()
        with open(idfile + ".tmp", "wb") as f:
            f.write(uid)
        os.rename(idfile + ".tmp", idfile)
    else:
        for _ in range(600):
            if os.path.exists(idfile):
                break
            time.sleep(0.1)
        uid = open(idfile, "rb").read()
    m._lib.comm_init(world, rank, uid)


def _rank(rank, world, tmp, lengths):
    import imcoalhmm_b200 as m
    lib = m._lib.load()
    m._lib.check(lib.imc_init(rank))
    theta, pis, Ts, Es = golden_model("isolation_k10")
    obs = example_symbols()
    cuts = np.concatenate([[0], np.cumsum(lengths)])
    mine = [obs[cuts[c]:cuts[c + 1]] for c in range(len(lengths)) if c % world == rank]
    fset = m.ForwarderSet([m.Forwarder.from_symbols(c, 3) for c in mine])
    # a rank whose shard holds nothing but an empty chunk still takes part in the sum
    lonely = fset if rank == 0 else m.ForwarderSet([m.Forwarder.from_symbols(np.zeros(0, dtype=np.int32), 3)])
    model = m.IsolationModel(10)
    results, fused = [], []
    for phase in range(2):      # 0: all-reduce fused into the reduction kernel over peer memory; 1: ncclAllReduce
        m.set_option("comm_fused", 1 - phase)
        _connect(m, rank, world, os.path.join(tmp, "nccl_id%d" % phase))
        fused.append(m._lib.comm_info()["fused"])
        assert m._lib.comm_info()["nranks"] == world and m._lib.comm_info()["rank"] == rank
        for rep in range(3):    # several epochs through the two mailbox halves
            a = fset.forward_batch(pis[:6], Ts[:6], Es[:6])                   # host arrays in, summed over ranks
        b = model.batched_log_likelihood(theta[:6], fset)                     # fused theta -> logL, summed over ranks
        c = fset.forward(pis[2], Ts[2], Es[2])                                # chain-scarce single point
        d = lonely.forward_batch(pis[:6], Ts[:6], Es[:6])                     # rank 0's shard only
        m.set_option("comm_enabled", 0)
        e = fset.forward_batch(pis[:6], Ts[:6], Es[:6])                       # this rank's partial sums
        m.set_option("comm_enabled", 1)
        results.append(np.concatenate([a, b, [c], d, e]))
        m._lib.comm_destroy()
    assert m._lib.comm_info() == {"nranks": 1, "rank": 0, "fused": False}
    # ---- a rank that does not show up is an error after the timeout, not a hang: rank 1 skips one collective call ----
    m.set_option("comm_fused", 1)
    m.set_option("comm_timeout_ms", 400)
    _connect(m, rank, world, os.path.join(tmp, "nccl_id_timeout"))
    timed_out = -1
    if m._lib.comm_info()["fused"]:
        ok = fset.forward_batch(pis[:6], Ts[:6], Es[:6])                      # both ranks: fine
        assert np.isfinite(ok).all()
        if rank == 0:
            t0 = time.time()
            try:
                fset.forward_batch(pis[:6], Ts[:6], Es[:6])                   # rank 1 never issues this call
                timed_out = 0
            except m.IMCError as e:
                timed_out = 1 if ("timed out" in str(e) and time.time() - t0 < 20) else 0
            try:
                fset.forward_batch(pis[:6], Ts[:6], Es[:6])                   # the communicator is unusable from here on
                timed_out = 0
            except m.IMCError:
                pass
        else:
            time.sleep(3.0)
    m._lib.comm_destroy()
    m.set_option("comm_timeout_ms", 30000)
    np.save(os.path.join(tmp, "timeout%d.npy" % rank), np.array([timed_out]))
    np.save(os.path.join(tmp, "out%d.npy" % rank), np.stack(results))
    np.save(os.path.join(tmp, "fused%d.npy" % rank), np.array(fused))


def test_native_allreduce_sums_over_two_gpus(tmp_path):
    import ctypes
    import imcoalhmm_b200 as m
    n = ctypes.c_int()
    m._lib.load().imc_device_count(ctypes.byref(n))
    if n.value < 2:
        pytest.skip("needs two GPUs")
    import multiprocessing as mp
    from oracle import forward as F
    lengths = [9000, 1, 14000, 8000, 12000, 254, 22000]
    ctx = mp.get_context("spawn")
    procs = [ctx.Process(target=_rank, args=(r, 2, str(tmp_path), lengths)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0
    _, pis, Ts, Es = golden_model("isolation_k10")
    obs = example_symbols().astype(np.int32)
    cuts = np.concatenate([[0], np.cumsum(lengths)])
    want, _ = F.forward_batch([obs[cuts[c]:cuts[c + 1]] for c in range(len(lengths))], pis[:6], Ts[:6], Es[:6])
    chunks = [obs[cuts[c]:cuts[c + 1]] for c in range(len(lengths))]
    want0, _ = F.forward_batch(chunks[0::2], pis[:6], Ts[:6], Es[:6])
    want1, _ = F.forward_batch(chunks[1::2], pis[:6], Ts[:6], Es[:6])
    outs = [np.load(tmp_path / ("out%d.npy" % r)) for r in range(2)]
    fused = [np.load(tmp_path / ("fused%d.npy" % r)) for r in range(2)]
    assert fused[0].tolist() == fused[1].tolist() and not fused[0][1]      # phase 1 is NCCL; phase 0 is fused where IPC works
    for r in range(2):
        for phase in range(2):
            got = outs[r][phase]
            np.testing.assert_allclose(got[:6], want, rtol=1e-11)
            np.testing.assert_allclose(got[6:12], want, rtol=1e-9)      # GPU-built (pi,T,E) vs reference-built fixture
            assert got[12] == pytest.approx(want[2], rel=1e-11)
            np.testing.assert_allclose(got[13:19], want0, rtol=1e-11)
            np.testing.assert_allclose(got[19:25], want1 if r else want0, rtol=1e-11)
        np.testing.assert_array_equal(outs[r][0], outs[r][1])             # both collectives add the same two numbers
    np.testing.assert_array_equal(outs[0][:, :19], outs[1][:, :19])       # and every rank holds the same bits
    print("fused all-reduce used:", bool(fused[0][0]))
    if fused[0][0]:
        assert int(np.load(tmp_path / "timeout0.npy")[0]) == 1             # rank 0 got IMC_ERR_CUDA "timed out", twice
