"""Sharding of alignment chunks across the GPUs of one box.

Chunks (= reference Forwarders, one per alignment file) are independent HMM runs whose log-likelihoods are
added (likelihood.py:33), so the path shards by chunk with no data-path exchange: every rank owns a contiguous
block of chunks balanced by total sites, scores the same parameter batch on it, and the partial logL[N] vectors
are summed by ONE all-reduce (NCCL over NVLink on the GPUs; gloo in the CPU tests).
"""
import numpy as np


def chunk_cost(symbols):
    """What a chunk costs the compressed forward kernel, up to a constant: the number of mat-vecs it is walked in.  In the
    run-token form (csrc/tokenizer.inl) that is one per maximal stretch of a non-run symbol -- every lone mismatch, every block
    of missing data -- however many matching sites lie in between; the sites themselves cost nothing.  O(L) in numpy."""
    sym = np.asarray(symbols)
    if sym.size == 0:
        return 0
    run_sym = int(np.argmax(np.bincount(sym.astype(np.int64), minlength=3)))
    other = sym != run_sym
    starts = other & np.concatenate([[True], sym[1:] != sym[:-1]])
    return int(starts.sum()) + 1


def partition_chunks(lengths, world_size, weights=None):
    """Contiguous blocks [(start, end), ...] of chunk indices, one per rank, balanced by total weight.

    weights defaults to the chunk lengths (sites); pass `[chunk_cost(c) for c in chunks]` to balance what the GPU time is
    proportional to -- mat-vecs, not sites: two alignments of equal length differ in cost by their divergence.
    Greedy prefix split at the ideal cumulative boundaries; every rank gets a (possibly empty) block and the
    blocks tile range(len(lengths)) in order, so chunk order -- and therefore the summation order inside a rank --
    is preserved."""
    lengths = np.asarray(lengths if weights is None else weights, dtype=np.int64)
    if weights is not None and len(weights) != len(np.atleast_1d(lengths)):
        raise ValueError("one weight per chunk")
    n = lengths.size
    if world_size < 1:
        raise ValueError("world_size must be >= 1")
    cum = np.concatenate([[0], np.cumsum(lengths)])
    total = cum[-1]
    bounds = [0]
    for r in range(1, world_size):
        target = total * r / world_size
        k = int(np.searchsorted(cum, target, side="left"))
        if k > 0 and abs(cum[k - 1] - target) <= abs(cum[min(k, n)] - target):
            k -= 1
        bounds.append(min(max(k, bounds[-1]), n))
    bounds.append(n)
    return [(bounds[r], bounds[r + 1]) for r in range(world_size)]


def init_library_comm(device, fused=True, process_group=None):
    """Give the library its own communicator from an initialised torch.distributed group: rank 0 draws the id, a
    broadcast carries it to the other ranks, every rank calls imc_comm_init.  From then on every forward / likelihood
    call of this process returns the SUM over ranks -- on one node inside the chain-reduction kernel over peer memory
    (`comm_info()["fused"]`), else by one ncclAllReduce -- and ShardedLikelihood must not all-reduce again
    (`summed_by_library=True`).  Call `_lib.comm_destroy()` on all ranks together when done.  Returns comm_info()."""
    import torch
    import torch.distributed as dist
    from . import _lib, set_option
    world, rank = dist.get_world_size(process_group), dist.get_rank(process_group)
    if world == 1:
        return _lib.comm_info()
    uid = torch.zeros(128, dtype=torch.uint8, device=device)
    if rank == 0:
        uid.copy_(torch.frombuffer(bytearray(_lib.comm_unique_id()), dtype=torch.uint8))
    dist.broadcast(uid, 0, group=process_group)
    set_option("comm_fused", 1 if fused else 0)
    _lib.comm_init(world, rank, bytes(uid.cpu().numpy().tobytes()))
    return _lib.comm_info()


class ShardedLikelihood(object):
    """Likelihood over chunks sharded across ranks.

    local_scorer(thetas) -> tensor float64[N] of this rank's partial log-likelihoods, on the device the process
    group communicates on (cuda for nccl, cpu for gloo).  `batched` all-reduces it (sum) and returns the tensor --
    unless the scorer's result is already the sum over ranks (`summed_by_library`: init_library_comm was called).
    """

    def __init__(self, local_scorer, process_group=None, summed_by_library=False):
        self.local_scorer = local_scorer
        self.process_group = process_group
        self.summed_by_library = summed_by_library

    def batched(self, thetas):
        import torch.distributed as dist
        part = self.local_scorer(thetas)
        if self.summed_by_library:
            return part
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(self.process_group) > 1:
            dist.all_reduce(part, op=dist.ReduceOp.SUM, group=self.process_group)   # the only collective
        return part


def make_gpu_scorer(model, forwarder_set, device):
    """local_scorer for ShardedLikelihood: fused theta -> logL on `device`, result left on the device."""
    import torch

    def scorer(thetas):
        th = torch.as_tensor(np.ascontiguousarray(thetas, dtype=np.float64)).to(device, non_blocking=True)
        out = torch.empty(th.shape[0], dtype=torch.float64, device=device)
        stream = torch.cuda.current_stream(device).cuda_stream
        model.batched_log_likelihood_device(th.data_ptr(), forwarder_set, out.data_ptr(), th.shape[0], 0, stream)
        return out

    return scorer
