"""Host-side logic of the multi-GPU path on CPU: chunk partitioning and the single all-reduce, world_size 2, gloo."""
import os
import socket

import numpy as np
import pytest

from conftest import golden_model, example_symbols


def test_partition_properties():
    from imcoalhmm_b200.sharding import partition_chunks
    rng = np.random.default_rng(0)
    for n, world in [(100, 8), (125, 8), (3, 8), (1, 2), (0, 4), (1000, 1), (17, 3)]:
        lengths = rng.integers(1, 2_000_000, size=n)
        blocks = partition_chunks(lengths, world)
        assert len(blocks) == world
        assert blocks[0][0] == 0 and blocks[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(blocks[:-1], blocks[1:]))      # contiguous, ordered, tiling
        assert all(s <= e for s, e in blocks)
        if n >= 4 * world:
            loads = np.array([lengths[s:e].sum() for s, e in blocks], dtype=float)
            assert loads.max() <= loads.mean() + lengths.max()                  # balanced within one chunk
    # equal chunks split evenly (config 3: 1000 chunks on 8 GPUs = 125 each)
    assert partition_chunks([1_000_000] * 1000, 8) == [(125 * r, 125 * (r + 1)) for r in range(8)]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, lengths, ret):
    import torch
    import torch.distributed as dist
    from oracle import forward as F
    from imcoalhmm_b200.sharding import partition_chunks, ShardedLikelihood
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    _, pis, Ts, Es = golden_model("isolation_k10")
    obs = example_symbols().astype(np.int32)
    cuts = np.concatenate([[0], np.cumsum(lengths)])
    start, end = partition_chunks(lengths, world)[rank]
    mine = [obs[cuts[c]:cuts[c + 1]] for c in range(start, end)]

    def scorer(_thetas):   # stands in for the GPU kernels: this rank's partial sums from the CPU oracle
        out, _ = F.forward_batch(mine, pis[:6], Ts[:6], Es[:6]) if mine else (np.zeros(6), 0)
        return torch.from_numpy(np.ascontiguousarray(out))

    got = ShardedLikelihood(scorer).batched(None).numpy()
    if rank == 0:
        ret.put(got)
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_sum_matches_single_process_gloo_world2():
    import torch.multiprocessing as mp
    from oracle import forward as F
    lengths = [9000, 1, 14000, 8000, 12000, 254, 22000]       # ragged chunks of the 65,255-site example
    assert sum(lengths) == 65255
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, lengths, ret)) for r in range(2)]
    for p in procs:
        p.start()
    got = ret.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    _, pis, Ts, Es = golden_model("isolation_k10")
    obs = example_symbols().astype(np.int32)
    cuts = np.concatenate([[0], np.cumsum(lengths)])
    want, _ = F.forward_batch([obs[cuts[c]:cuts[c + 1]] for c in range(len(lengths))], pis[:6], Ts[:6], Es[:6])
    np.testing.assert_allclose(got, want, rtol=1e-13)


def test_library_comm_single_rank_is_a_no_op():
    """init_library_comm with a world of one never touches NCCL or the GPU: comm_info reports one rank, nothing fused."""
    import torch.distributed as dist
    from imcoalhmm_b200 import _lib
    from imcoalhmm_b200.sharding import init_library_comm, ShardedLikelihood
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(_free_port())
    dist.init_process_group("gloo", rank=0, world_size=1)
    try:
        assert init_library_comm("cpu") == {"nranks": 1, "rank": 0, "fused": False}
        import torch
        part = torch.arange(4, dtype=torch.float64)
        assert ShardedLikelihood(lambda th: part.clone()).batched(None).tolist() == part.tolist()
        assert ShardedLikelihood(lambda th: part.clone(), summed_by_library=True).batched(None).tolist() == part.tolist()
    finally:
        dist.destroy_process_group()
    _lib.comm_destroy()                      # harmless without a communicator
    assert _lib.comm_info()["nranks"] == 1


def test_partition_by_cost_balances_mat_vecs_not_sites():
    """GPU time follows the number of run tokens (one mat-vec per non-matching stretch), not the number of sites: equal-length
    chunks of different divergence are balanced by `weights=chunk_cost`."""
    from imcoalhmm_b200.sharding import chunk_cost, partition_chunks
    rng = np.random.default_rng(0)
    chunks = [rng.choice(3, size=20000, p=[1 - d - 0.001, d, 0.001]).astype(np.uint8) for d in (0.001, 0.001, 0.001, 0.001, 0.02, 0.02)]
    cost = [chunk_cost(c) for c in chunks]
    assert cost[4] > 5 * cost[0]
    # cost counts maximal stretches: 0 0 1 1 0 2 2 2 0 1 -> three stretches (+1)
    assert chunk_cost(np.array([0, 0, 1, 1, 0, 2, 2, 2, 0, 1], dtype=np.uint8)) == 4
    assert chunk_cost(np.zeros(0, dtype=np.uint8)) == 0
    by_sites = partition_chunks([len(c) for c in chunks], 2)
    by_cost = partition_chunks([len(c) for c in chunks], 2, weights=cost)
    assert by_sites == [(0, 3), (3, 6)]
    load = lambda blocks: [sum(cost[a:b]) for a, b in blocks]
    assert max(load(by_cost)) < max(load(by_sites))
    assert by_cost[0][1] >= 4 and by_cost[0][0] == 0 and by_cost[-1][1] == 6
