"""Python face of the CPU forward oracle (oracle/forward_oracle.c) + a brute-force enumerator.

TEST INFRASTRUCTURE ONLY -- see the header of forward_oracle.c.  The product package never
imports this module.  Parity at the ziphmm boundary is UNPINNED by the reference (no ziphmm here,
no log-likelihood in the reference's tests); these functions restate the recursion behind
/root/reference/src/IMCoalHMM/hmm.py:19-21 and likelihood.py:33.
"""
import ctypes
import itertools
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libimco_oracle.so")
_lib = None

_i32p = ctypes.POINTER(ctypes.c_int32)
_f64p = ctypes.POINTER(ctypes.c_double)


def build(force=False):
    """Compile the C oracle (gcc, a few hundred ms)."""
    src = os.path.join(_HERE, "forward_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_SO)
        L.imco_forward_plain.restype = ctypes.c_double
        L.imco_forward_plain.argtypes = [_i32p, ctypes.c_int64, ctypes.c_int, ctypes.c_int, _f64p, _f64p, _f64p]
        L.imco_forward_plain_ld.restype = ctypes.c_double
        L.imco_forward_plain_ld.argtypes = [_i32p, ctypes.c_int64, ctypes.c_int, ctypes.c_int, _f64p, _f64p, _f64p,
                                            _f64p, _f64p]
        L.imco_zip_preprocess.restype = ctypes.c_int
        L.imco_zip_preprocess.argtypes = [_i32p, ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                          ctypes.POINTER(_i32p), ctypes.POINTER(ctypes.c_int64),
                                          ctypes.POINTER(_i32p), ctypes.POINTER(ctypes.c_int)]
        L.imco_free.restype = None
        L.imco_free.argtypes = [ctypes.c_void_p]
        L.imco_zip_forward.restype = ctypes.c_double
        L.imco_zip_forward.argtypes = [_f64p, _f64p, _f64p, _i32p, _i32p, ctypes.c_int64, ctypes.c_int,
                                       ctypes.c_int, ctypes.c_int]
        L.imco_zip_forward_fast.restype = ctypes.c_double
        L.imco_zip_forward_fast.argtypes = L.imco_zip_forward.argtypes
        L.imco_forward_batch.restype = ctypes.c_int
        L.imco_forward_batch.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.POINTER(_i32p),
                                         ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(_i32p),
                                         ctypes.POINTER(ctypes.c_int), ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                         _f64p, _f64p, _f64p, _f64p, ctypes.c_int]
        _lib = L
    return _lib


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _p(a, t):
    return a.ctypes.data_as(t)


def _hmm(pi, T, E):
    pi = _f64(pi).reshape(-1)
    T = _f64(T)
    E = _f64(E)
    K = pi.size
    assert T.shape == (K, K) and E.shape[0] == K
    return pi, T, E, K, E.shape[1]


def forward_plain(obs, pi, T, E):
    """Scaled forward in float64 (hmm.py:19-21 semantics)."""
    obs = np.ascontiguousarray(obs, dtype=np.int32)
    pi, T, E, K, S = _hmm(pi, T, E)
    return lib().imco_forward_plain(_p(obs, _i32p), obs.size, K, S, _p(pi, _f64p), _p(T, _f64p), _p(E, _f64p))


def forward_plain_ld(obs, pi, T, E):
    """Same recursion in x87 long double; returns (hi, lo) with hi+lo the extended result."""
    obs = np.ascontiguousarray(obs, dtype=np.int32)
    pi, T, E, K, S = _hmm(pi, T, E)
    hi, lo = ctypes.c_double(), ctypes.c_double()
    lib().imco_forward_plain_ld(_p(obs, _i32p), obs.size, K, S, _p(pi, _f64p), _p(T, _f64p), _p(E, _f64p),
                                ctypes.byref(hi), ctypes.byref(lo))
    return hi.value, lo.value


def zip_preprocess(obs, nsym, min_count=16, max_syms=1024):
    """zipHMM-style pair compression (hmm.py:16 contract): -> (new_obs, sym2pair[P,2], new_nsyms)."""
    obs = np.ascontiguousarray(obs, dtype=np.int32)
    new_obs, pairs = _i32p(), _i32p()
    newL, nsyms = ctypes.c_int64(), ctypes.c_int()
    lib().imco_zip_preprocess(_p(obs, _i32p), obs.size, nsym, min_count, max_syms, ctypes.byref(new_obs),
                              ctypes.byref(newL), ctypes.byref(pairs), ctypes.byref(nsyms))
    out = np.ctypeslib.as_array(new_obs, shape=(max(newL.value, 1),))[:newL.value].copy()
    npairs = nsyms.value - nsym
    sym2pair = (np.ctypeslib.as_array(pairs, shape=(max(npairs, 1) * 2,))[:2 * npairs].copy().reshape(npairs, 2))
    lib().imco_free(new_obs)
    lib().imco_free(pairs)
    return out, sym2pair, nsyms.value


def zip_forward(pi, T, E, sym2pair, new_obs, nsym, new_nsyms):
    """hmm.py:20-21 contract."""
    pi, T, E, K, S = _hmm(pi, T, E)
    assert S == nsym
    new_obs = np.ascontiguousarray(new_obs, dtype=np.int32)
    sym2pair = np.ascontiguousarray(sym2pair, dtype=np.int32).reshape(-1)
    if sym2pair.size == 0:
        sym2pair = np.zeros(2, dtype=np.int32)
    return lib().imco_zip_forward(_p(pi, _f64p), _p(T, _f64p), _p(E, _f64p), _p(sym2pair, _i32p),
                                  _p(new_obs, _i32p), new_obs.size, nsym, new_nsyms, K)


def zip_forward_fast(pi, T, E, sym2pair, new_obs, nsym, new_nsyms):
    """Same contract as zip_forward, through the tuned implementation the CPU arm of bench.py times."""
    pi, T, E, K, S = _hmm(pi, T, E)
    assert S == nsym
    new_obs = np.ascontiguousarray(new_obs, dtype=np.int32)
    sym2pair = np.ascontiguousarray(sym2pair, dtype=np.int32).reshape(-1)
    if sym2pair.size == 0:
        sym2pair = np.zeros(2, dtype=np.int32)
    return lib().imco_zip_forward_fast(_p(pi, _f64p), _p(T, _f64p), _p(E, _f64p), _p(sym2pair, _i32p),
                                       _p(new_obs, _i32p), new_obs.size, nsym, new_nsyms, K)


def forward_batch(seqs, pis, Ts, Es, mode="plain", zipped=None, nthreads=0):
    """out[n] = sum_c logL(seq_c | pi_n, T_n, E_n) on the host cores (likelihood.py:33, batched).

    mode "plain": seqs = list of int32 arrays.  mode "zip": zipped = list of (new_obs, sym2pair, new_nsyms).
    Returns (out[N], threads_used)."""
    pis = _f64(pis)
    Ts = _f64(Ts)
    Es = _f64(Es)
    N, K = pis.shape
    S = Es.shape[2]
    if mode == "plain":
        obs = [np.ascontiguousarray(s, dtype=np.int32) for s in seqs]
        pairs = [np.zeros(2, dtype=np.int32) for _ in obs]
        nsyms = [S] * len(obs)
    else:
        obs = [np.ascontiguousarray(z[0], dtype=np.int32) for z in zipped]
        pairs = [np.ascontiguousarray(z[1], dtype=np.int32).reshape(-1) if np.size(z[1]) else
                 np.zeros(2, dtype=np.int32) for z in zipped]
        nsyms = [int(z[2]) for z in zipped]
    C = len(obs)
    obs_p = (_i32p * C)(*[_p(o, _i32p) for o in obs])
    pairs_p = (_i32p * C)(*[_p(p, _i32p) for p in pairs])
    lens = (ctypes.c_int64 * C)(*[o.size for o in obs])
    ns = (ctypes.c_int * C)(*nsyms)
    out = np.zeros(N, dtype=np.float64)
    used = lib().imco_forward_batch({"plain": 0, "zip": 1, "zip_fast": 2}[mode], C, obs_p, lens, pairs_p, ns, S, N, K,
                                    _p(pis, _f64p), _p(Ts, _f64p), _p(Es, _f64p), _p(out, _f64p), nthreads)
    return out, used


def forward_numpy(obs, pi, T, E, dtype=np.float64):
    """Pure NumPy restatement (python loop; small inputs only)."""
    pi = np.asarray(pi, dtype=dtype).reshape(-1)
    T = np.asarray(T, dtype=dtype)
    E = np.asarray(E, dtype=dtype)
    a = pi * E[:, obs[0]]
    c = a.sum()
    a = a / c
    logl = np.log(c)
    for o in obs[1:]:
        a = (a @ T) * E[:, o]
        c = a.sum()
        a = a / c
        logl = logl + np.log(c)
    return logl


def forward_bruteforce(obs, pi, T, E):
    """log of the sum over ALL hidden paths -- the definition.  K**L terms: tiny inputs only."""
    pi = np.asarray(pi, dtype=np.float64).reshape(-1)
    T = np.asarray(T, dtype=np.float64)
    E = np.asarray(E, dtype=np.float64)
    K, L = pi.size, len(obs)
    assert K ** L <= 2_000_000
    total = 0.0
    for path in itertools.product(range(K), repeat=L):
        p = pi[path[0]] * E[path[0], obs[0]]
        for t in range(1, L):
            p *= T[path[t - 1], path[t]] * E[path[t], obs[t]]
        total += p
    return float(np.log(total))
