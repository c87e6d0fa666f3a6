"""Aligned form of the MMA kernel, host side (no GPU): the per-warp-load schedules built by zip_align_quad (zip_host.inl).  Every
chain keeps its own tokens in their order, every warp-step names exactly one dictionary entry, and the schedule of the library's own threshold is
not longer than lock step with one pass per distinct entry would be."""
import numpy as np
import pytest

NOP = 1 << 21


def _set(rng, n_chunks, length, p=(0.93, 0.03, 0.04), runs=True):
    import imcoalhmm_b200 as m
    chunks = []
    for _ in range(n_chunks):
        c = rng.choice(3, size=int(length * (0.7 + 0.6 * rng.random())), p=list(p)).astype(np.uint8)
        if runs:                                   # blocks of missing data, like the benchmark alignments
            for _ in range(max(1, c.size // 2500)):
                a = int(rng.integers(0, c.size))
                c[a:a + int(rng.geometric(0.01))] = 2
        chunks.append(c)
    return m.ForwarderSet([m.Forwarder.from_symbols(c, 3) for c in chunks])


@pytest.mark.parametrize("K,n_chunks", [(10, 19), (20, 8), (40, 5)])
def test_schedules_keep_every_chain_and_name_one_entry_per_step(K, n_chunks):
    rng = np.random.default_rng(K)
    s = _set(rng, n_chunks, 30000)
    info = s.align_info(K)
    assert 2 <= info["stall"] <= 6 and info["aligned_steps"] > 0 and info["est_passes"] >= 1.0
    total = 0
    for quad in range((n_chunks + 7) // 8):
        raw = s.align_quad(K, quad, -1)
        for stall in (1, 2, info["stall"], 8):
            al = s.align_quad(K, quad, stall)
            assert al.shape[0] == raw.shape[0] == min(8, n_chunks - 8 * quad) and al.shape[1] % 8 == 0
            ids = al & 0xff
            assert (ids == ids[0]).all()                                   # one entry per step, known to every lane
            served = ((al & NOP) == 0).sum(axis=0)
            pad = served == 0                                              # idle steps: only the padding to a multiple of 8, at the very end
            assert pad.sum() < 8 and not pad[:al.shape[1] - int(pad.sum())].any()
            assert (ids[0][pad] == info["hot_id"]).all()                   # ... and they name the hot entry
            for c in range(al.shape[0]):
                mine = al[c][(al[c] & NOP) == 0]
                want = raw[c][(raw[c] & NOP) == 0]
                assert np.array_equal(mine, want)                          # the chain's own tokens, in order, nothing else
                assert (((al[c][(al[c] & NOP) != 0] >> 8) & 0x1fff) == 0).all()     # a no-op word carries no run (rows 0 of the tables)
            # the library's threshold beats lock step with one pass per distinct entry; no threshold is much worse
            lock = 0
            for t in range(raw.shape[1]):
                col = raw[:, t]
                lock += len(set((col[(col & NOP) == 0] & 0xff).tolist()))
            assert al.shape[1] <= 1.05 * lock + 8
            if stall == info["stall"]:
                assert al.shape[1] <= lock + 8
            if stall == info["stall"]:
                total += al.shape[1]
    assert total == info["aligned_steps"]


def test_align_info_of_a_set_without_a_second_run_symbol():
    import imcoalhmm_b200 as m
    obs = np.zeros(5000, dtype=np.uint8)
    obs[::7] = 1                                   # isolated mismatches only: nothing but the run symbol comes in runs
    s = m.ForwarderSet([m.Forwarder.from_symbols(obs, 2)])
    with pytest.raises(m.IMCError):
        s.align_info(10)
