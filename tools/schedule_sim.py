"""Offline study: how many chain-steps one MMA pass can serve when the 8 chains of a warp follow a common, host-built schedule
(a supersequence of their token-type sequences) instead of marching in lock step.  Two-run token types from simulated
benchmark chunks: 0 = mismatch (hot), 1 = X (into the missing-data basis), 2 = Y (back), 3 = short missing, 4 = '11'."""
import sys
import numpy as np
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
import bench  # noqa: E402


def events(obs):
    nz = np.flatnonzero(obs); out = []; i = 0; n = len(nz)
    while i < n:
        p = nz[i]
        if obs[p] == 1:
            if i + 1 < n and nz[i + 1] == p + 1 and obs[p + 1] == 1: out.append(4); i += 2
            else: out.append(0); i += 1
        else:
            j = i
            while j + 1 < n and nz[j + 1] == nz[j] + 1 and obs[nz[j + 1]] == 2: j += 1
            L = j - i + 1
            out += [1, 2] if L >= 3 else [3] * L
            i = j + 1
    return out


def lockstep(chains):
    n = max(len(c) for c in chains); passes = 0
    for s in range(n):
        ids = set(c[s] for c in chains if s < len(c))
        passes += 1 + len(ids - {0})
    return sum(len(c) for c in chains) / passes


def periodic(chains, P):
    ptr = [0] * len(chains); passes = 0; total = sum(len(c) for c in chains)
    def nxt(i): return chains[i][ptr[i]] if ptr[i] < len(chains[i]) else -1
    while any(ptr[i] < len(chains[i]) for i in range(len(chains))):
        for _ in range(P):
            hot = [i for i in range(len(chains)) if nxt(i) == 0]
            if not hot: break
            passes += 1
            for i in hot: ptr[i] += 1
        for _ in range(2):                     # cold slot, twice (a chain served by X wants Y right away)
            for t in (1, 2, 3, 4):
                w = [i for i in range(len(chains)) if nxt(i) == t]
                if w:
                    passes += 1
                    for i in w: ptr[i] += 1
    return total / passes


if __name__ == "__main__":
    cf = bench.ChunkFactory(1)
    for wname in ("c2", "c3_1gpu", "c5_1gpu"):
        wl = dict(bench.WORKLOADS[wname]); wl["chunk_len"] = min(wl["chunk_len"], 1000000)
        ev = [events(c) for c in cf.make(wl, range(8))]
        print(wname, "lock step %.2f" % lockstep(ev), " periodic:", " ".join("P=%d %.2f" % (P, periodic(ev, P)) for P in (2, 3, 4, 5, 6, 8, 10, 12)))


def adaptive(chains, g):
    """hot passes until g chains are stalled at a cold entry (or nothing hot is left), then one pass per waiting cold type, twice"""
    n = len(chains); ptr = [0] * n; passes = 0; total = sum(len(c) for c in chains)
    def nxt(i): return chains[i][ptr[i]] if ptr[i] < len(chains[i]) else -1
    while any(ptr[i] < len(chains[i]) for i in range(n)):
        while True:
            hot = [i for i in range(n) if nxt(i) == 0]
            stalled = sum(1 for i in range(n) if nxt(i) > 0)
            if not hot or stalled >= g: break
            passes += 1
            for i in hot: ptr[i] += 1
        for _ in range(2):
            for t in (1, 2, 3, 4):
                w = [i for i in range(n) if nxt(i) == t]
                if w:
                    passes += 1
                    for i in w: ptr[i] += 1
    return total / passes
