"""One long chunk cut over the GPUs (fewer chunks than GPUs): time per batched call with the chunk whole on one
GPU and cut into one part per rank.  torchrun --nproc-per-node G tools/parts_bench.py [--mbp 100] [--points 1,64]"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--mbp", type=int, default=100)
ap.add_argument("--points", default="1,64")
ap.add_argument("--workload", default="c2")
args = ap.parse_args()
rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
wl = bench.WORKLOADS[args.workload]
pis, Ts, Es = bench.load_points(wl["model"], 1)
rng = np.random.Generator(np.random.PCG64(bench.SEED0))
obs = bench.simulate_chunk(rng, pis[0], Ts[0], Es[0], args.mbp * 1_000_000)       # the same chunk on every rank

import torch  # noqa: E402
import imcoalhmm_b200 as m  # noqa: E402
torch.cuda.set_device(rank)
m._lib.check(m._lib.load().imc_init(rank))
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
    from imcoalhmm_b200.sharding import init_library_comm
    init_library_comm(torch.device("cuda", rank))
model = getattr(m, wl["ctor"][0])(*wl["ctor"][1])
bounds = np.linspace(0, len(obs), world + 1).astype(int)
t0 = time.time()
part = m.ForwarderSet([m.Forwarder.from_symbols(obs[bounds[rank]:bounds[rank + 1]], 3)], parts=(rank, world))
whole = m.ForwarderSet([m.Forwarder.from_symbols(obs, 3)]) if rank == 0 else None
t_prep = time.time() - t0
for npts in map(int, args.points.split(",")):
    thetas = bench.thetas_around(wl["default"], npts)
    res = {}
    for name, fset in (("parts", part), ("whole", whole)):
        if fset is None:
            continue
        if name == "whole":
            m.set_option("comm_enabled", 0)
        out = model.batched_log_likelihood(thetas, fset)
        ts = []
        for _ in range(5):
            t0 = time.perf_counter()
            out = model.batched_log_likelihood(thetas, fset)
            ts.append(time.perf_counter() - t0)
        m.set_option("comm_enabled", 1)
        res[name] = (min(ts) * 1e3, out, m.last_forward_kernel())
    if world > 1:
        dist.barrier()
    if rank == 0:
        rel = float(np.max(np.abs(res["parts"][1] - res["whole"][1]) / np.abs(res["whole"][1])))
        print("%d Mbp chunk, %d points: %d parts on %d GPUs %.3f ms (%s) | whole chunk on one GPU %.3f ms (%s) | max rel diff %.1e | preprocess %.1fs"
              % (args.mbp, npts, world, world, res["parts"][0], res["parts"][2], res["whole"][0], res["whole"][2], rel, t_prep), flush=True)
if world > 1:
    m._lib.comm_destroy()
    dist.destroy_process_group()
