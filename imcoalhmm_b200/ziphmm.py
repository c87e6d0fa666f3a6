"""`ziphmm`-compatible module so that the reference's unmodified hmm.py works against this package.

The reference imports `ziphmm` (hmm.py:7) and uses exactly two functions (hmm.py:16, :20-21):
    preprocess_raw_observations(obs, nsym) -> (new_obs, sym2pair, new_nsyms)
    zip_forward(pi, T, E, sym2pair, new_obs, nsym, new_nsyms) -> float
Putting this file's directory-level alias on sys.path as `ziphmm` routes both to
the B200 kernels.  preprocess_raw_observations returns this library's own pair compression (csrc/tokenizer.inl:
greedy most-frequent-pair merges, at most 256 ids); any exact re-encoding gives the same likelihood, so the merges
need not be mini-ziphmm's.  The Forwarder that owns the device-resident token streams rides along on the returned
array; zip_forward also accepts encodings produced elsewhere (they are expanded and re-encoded).
"""
import numpy as np


class _TaggedObs(np.ndarray):
    """int32 observation array that remembers its device-resident Forwarder."""
    _imc_forwarder = None


def _tag(arr, forwarder):
    out = np.asarray(arr, dtype=np.int32).view(_TaggedObs)
    out._imc_forwarder = forwarder
    return out


def preprocess_raw_observations(obs, nsym):
    from .hmm import Forwarder
    obs = np.ascontiguousarray(obs, dtype=np.int32)
    fwd = Forwarder.from_symbols(obs, nsym)
    return fwd.new_obs, fwd.sym2pair, fwd.new_nsyms


def _expand(sym2pair, new_obs, nsym, new_nsyms):
    """Undo a zipHMM pair encoding: symbol id >= nsym stands for (left, right) = sym2pair[id - nsym]."""
    sym2pair = np.asarray(sym2pair, dtype=np.int64).reshape(-1, 2)
    table = [[s] for s in range(nsym)]
    for left, right in sym2pair[: new_nsyms - nsym]:
        table.append(table[left] + table[right])
    lengths = np.array([len(t) for t in table], dtype=np.int64)
    new_obs = np.asarray(new_obs, dtype=np.int64)
    out = np.empty(int(lengths[new_obs].sum()), dtype=np.int32)
    pos = 0
    for s in new_obs:
        t = table[s]
        out[pos:pos + len(t)] = t
        pos += len(t)
    return out


def zip_forward(pi, T, E, sym2pair, new_obs, nsym, new_nsyms):
    from .hmm import Forwarder
    fwd = getattr(new_obs, "_imc_forwarder", None)
    if fwd is None:   # a foreign (really compressed) encoding: expand it, score it, do not cache
        fwd = Forwarder.from_symbols(_expand(sym2pair, new_obs, int(nsym), int(new_nsyms)), int(nsym))
    return fwd.forward(pi, T, E)
