"""Drop-in for the reference's Forwarder boundary (/root/reference/src/IMCoalHMM/hmm.py:10-21).

`Forwarder(input_filename, NSYM)` and `.forward(init_probs, trans_probs, emission_probs) -> float`
keep the reference's names, argument meaning and error behaviour; the arithmetic runs in the
hand-written sm_100a kernels of libimcoalhmm_b200.so through its C ABI.  New on top of the
reference: `forward_batch` (N parameter points per call) and `ForwarderSet` (all the chunks a
Likelihood sums over, likelihood.py:33, scored in one launch).

The legacy pyZipHMM constructors used by the reference's ILS scripts are mirrored as
`Forwarder.fromSequence` (scripts/prepare-alignments.py:201) and `Forwarder.fromDirectory`
(scripts/ils-isolation-model.py:112).
"""
import ctypes
import os

import numpy as np

from . import _lib
from ._lib import IMCError, check  # noqa: F401  (re-exported)


def _as_f64(a, shape_tail, what):
    a = np.ascontiguousarray(np.asarray(a, dtype=np.float64))
    if a.shape[-len(shape_tail):] != tuple(shape_tail):
        raise ValueError("%s has shape %s, expected (..., %s)" % (what, a.shape, ", ".join(map(str, shape_tail))))
    return a


def _from_legacy_matrix(x):
    """pyZipHMM.Matrix objects (getHeight / getWidth / [i, j]) as used by the reference's legacy callers
    (ILS.py:271-276, isolation_model.py:131-146) -> ndarray; anything else is returned unchanged."""
    if hasattr(x, "getHeight") and hasattr(x, "getWidth"):
        h, w = int(x.getHeight()), int(x.getWidth())
        return np.array([[float(x[i, j]) for j in range(w)] for i in range(h)], dtype=np.float64)
    return x


def _hmm_arrays(pis, Ts, Es, batched):
    """Normalise (pi, T, E) to contiguous float64 arrays [N,K], [N,K,K], [N,K,S]."""
    pis, Ts, Es = _from_legacy_matrix(pis), _from_legacy_matrix(Ts), _from_legacy_matrix(Es)
    Ts = np.ascontiguousarray(np.asarray(Ts, dtype=np.float64))
    if not batched:
        Ts = Ts[None]
    if Ts.ndim != 3 or Ts.shape[1] != Ts.shape[2]:
        raise ValueError("transition matrix must be square, got shape %s" % (Ts.shape[1:],))
    N, K = Ts.shape[0], Ts.shape[1]
    pis = np.ascontiguousarray(np.asarray(pis, dtype=np.float64)).reshape(N, -1)   # accepts (K,), (K,1), (1,K)
    if pis.shape[1] != K:
        raise ValueError("initial probabilities have %d entries for %d states" % (pis.shape[1], K))
    Es = np.ascontiguousarray(np.asarray(Es, dtype=np.float64))
    if not batched:
        Es = Es[None]
    if Es.ndim != 3 or Es.shape[0] != N or Es.shape[1] != K:
        raise ValueError("emission matrix has shape %s, expected (%d, S)" % (Es.shape[1:], K))
    return pis, Ts, Es, N, K, Es.shape[2]


class _Seq(object):
    """Owner of one imc_seq handle."""

    def __init__(self, handle):
        self.handle = handle

    def __del__(self):
        h, self.handle = getattr(self, "handle", None), None
        if h:
            try:
                _lib.load().imc_seq_destroy(h)
            except Exception:
                pass

    @property
    def length(self):
        n = ctypes.c_int64()
        check(_lib.load().imc_seq_length(self.handle, ctypes.byref(n)))
        return n.value

    def symbols(self):
        out = np.empty(self.length, dtype=np.uint8)
        check(_lib.load().imc_seq_symbols(self.handle, out.ctypes.data_as(_lib.c_u8p), out.size))
        return out


class ForwarderSet(object):
    """The chunks a Likelihood sums over, packed once and resident on the GPU.

    forward(pi, T, E) == sum(f.forward(pi, T, E) for f in forwarders)   (likelihood.py:33)
    """

    def __init__(self, forwarders, parts=None):
        """parts=(part_first, parts_total): the forwarders are consecutive PARTS of one long chunk that is cut over the
        ranks of the library's communicator (fewer chunks than GPUs; include/imcoalhmm_b200.h: imc_seqset_create_parts)."""
        if isinstance(forwarders, Forwarder):
            forwarders = [forwarders]
        self.forwarders = list(forwarders)
        self._seqs = [f._seq for f in self.forwarders]   # keep the handles alive
        lib = _lib.load()
        arr = (_lib.c_vp * max(1, len(self._seqs)))(*[s.handle for s in self._seqs])
        h = _lib.c_vp()
        if parts is not None:
            check(lib.imc_seqset_create_parts(arr, len(self._seqs), int(parts[0]), int(parts[1]), ctypes.byref(h)))
        else:
            check(lib.imc_seqset_create(arr, len(self._seqs), ctypes.byref(h)))
        self._handle = h
        n, sites, nbytes = ctypes.c_int(), ctypes.c_int64(), ctypes.c_int64()
        check(lib.imc_seqset_info(h, ctypes.byref(n), ctypes.byref(sites), ctypes.byref(nbytes)))
        self.n_chunks, self.total_sites, self.packed_bytes = n.value, sites.value, nbytes.value

    def __del__(self):
        h, self._handle = getattr(self, "_handle", None), None
        if h:
            try:
                _lib.load().imc_seqset_destroy(h)
            except Exception:
                pass

    def __len__(self):
        return self.n_chunks

    def forward(self, init_probs, trans_probs, emission_probs):
        pis, Ts, Es, _, K, S = _hmm_arrays(init_probs, trans_probs, emission_probs, batched=False)
        out = ctypes.c_double()
        check(_lib.load().imc_forward(self._handle, K, S, pis.ctypes.data_as(_lib.c_f64p),
                                      Ts.ctypes.data_as(_lib.c_f64p), Es.ctypes.data_as(_lib.c_f64p),
                                      ctypes.byref(out)))
        return out.value

    def forward_batch(self, pis, Ts, Es, out=None):
        """logL for N parameter points: pis [N,K], Ts [N,K,K], Es [N,K,S] (host arrays) -> float64[N]."""
        pis, Ts, Es, N, K, S = _hmm_arrays(pis, Ts, Es, batched=True)
        if out is None:
            out = np.empty(N, dtype=np.float64)
        assert out.dtype == np.float64 and out.size == N and out.flags.c_contiguous
        check(_lib.load().imc_forward_batch(self._handle, N, K, S, pis.ctypes.data_as(_lib.c_f64p),
                                            Ts.ctypes.data_as(_lib.c_f64p), Es.ctypes.data_as(_lib.c_f64p),
                                            out.ctypes.data_as(_lib.c_f64p)))
        return out

    # -- the zipHMM-style preprocessing of the set (hmm.py:16: new_obs, sym2pair, new_nsyms), host only ----------
    def zip_info(self, K=None):
        """dict(ids_available=new_nsyms of the shared dictionary; with K: ids_used, tokens, levels for a K-state model)."""
        lib = _lib.load()
        avail, used, lev, tok = ctypes.c_int(), ctypes.c_int(), ctypes.c_int(), ctypes.c_int64()
        if K is None:
            check(lib.imc_seqset_zip_info(self._handle, 0, ctypes.byref(avail), None, None, None))
            return {"ids_available": avail.value}
        check(lib.imc_seqset_zip_info(self._handle, int(K), ctypes.byref(avail), ctypes.byref(used), ctypes.byref(tok),
                                      ctypes.byref(lev)))
        return {"ids_available": avail.value, "ids_used": used.value, "tokens": tok.value, "levels": lev.value}

    def zip_pairs(self):
        """sym2pair[P,2] (uint8): row i = (left, right) of id NSYM + i, creation order."""
        n = self.zip_info()["ids_available"] - (self.forwarders[0].NSYM if self.forwarders else 0)
        out = np.zeros((max(n, 0), 2), dtype=np.uint8)
        if n > 0:
            check(_lib.load().imc_seqset_zip_pairs(self._handle, out.ctypes.data_as(_lib.c_u8p), n))
        return out

    def zip_tokens(self, chunk, ids=None):
        """new_obs of chunk `chunk` (positions 1..L-1) over the first `ids` dictionary ids (default: all)."""
        lib = _lib.load()
        if ids is None:
            ids = self.zip_info()["ids_available"]
        n = ctypes.c_int64()
        check(lib.imc_seqset_zip_tokens(self._handle, int(chunk), int(ids), None, 0, ctypes.byref(n)))
        out = np.empty(n.value, dtype=np.uint8)
        if n.value:
            check(lib.imc_seqset_zip_tokens(self._handle, int(chunk), int(ids), out.ctypes.data_as(_lib.c_u8p), out.size,
                                            ctypes.byref(n)))
        return out

    # -- run tokens (spectral form of the compressed kernel; include/imcoalhmm_b200.h) ---------------------------
    def run_info(self, K=None):
        """dict(run_sym, ids_available; with K: ids_used, tokens, levels for a K-state model)."""
        lib = _lib.load()
        sym, avail, used, lev, tok = ctypes.c_int(), ctypes.c_int(), ctypes.c_int(), ctypes.c_int(), ctypes.c_int64()
        if K is None:
            check(lib.imc_seqset_run_info(self._handle, 0, ctypes.byref(sym), ctypes.byref(avail), None, None, None))
            return {"run_sym": sym.value, "ids_available": avail.value}
        check(lib.imc_seqset_run_info(self._handle, int(K), ctypes.byref(sym), ctypes.byref(avail), ctypes.byref(used),
                                      ctypes.byref(tok), ctypes.byref(lev)))
        return {"run_sym": sym.value, "ids_available": avail.value, "ids_used": used.value, "tokens": tok.value,
                "levels": lev.value}

    def align_info(self, K, stall=0):
        """Aligned form for a K-state model (host only): dict(lock_steps, est_passes, aligned_steps, stall, hot_id)."""
        lock, al, est = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_double()
        st, hot = ctypes.c_int32(), ctypes.c_int32()
        check(_lib.load().imc_seqset_align_info(self._handle, int(K), int(stall), ctypes.byref(lock), ctypes.byref(est), ctypes.byref(al),
                                                ctypes.byref(st), ctypes.byref(hot)))
        return {"lock_steps": lock.value, "est_passes": est.value, "aligned_steps": al.value, "stall": st.value, "hot_id": hot.value}

    def align_quad(self, K, quad, stall):
        """uint32[nchains, steps]: the words of warp-load `quad`, aligned (stall > 0) or the chains' own streams (stall <= 0)."""
        lib = _lib.load()
        steps, nch = ctypes.c_int64(), ctypes.c_int32()
        check(lib.imc_seqset_align_quad(self._handle, int(K), int(quad), int(stall), None, 0, ctypes.byref(steps), ctypes.byref(nch)))
        out = np.zeros((nch.value, steps.value), dtype=np.uint32)
        if out.size:
            check(lib.imc_seqset_align_quad(self._handle, int(K), int(quad), int(stall), out.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32)),
                                            out.size, ctypes.byref(steps), ctypes.byref(nch)))
        return out

    def run_pairs(self):
        """[P,2] (uint8): row i = (left, right) of run-dictionary id NSYM + i, creation order."""
        n = self.run_info()["ids_available"] - (self.forwarders[0].NSYM if self.forwarders else 0)
        out = np.zeros((max(n, 0), 2), dtype=np.uint8)
        if n > 0:
            check(_lib.load().imc_seqset_run_pairs(self._handle, out.ctypes.data_as(_lib.c_u8p), n))
        return out

    def run_tokens(self, chunk, ids=None):
        """(first_run, words uint32[ntok]) of chunk `chunk` over the first `ids` run-dictionary ids (default: all);
        word = id | run << 8."""
        lib = _lib.load()
        if ids is None:
            ids = self.run_info()["ids_available"]
        n, fr = ctypes.c_int64(), ctypes.c_int()
        check(lib.imc_seqset_run_tokens(self._handle, int(chunk), int(ids), None, 0, ctypes.byref(n), ctypes.byref(fr)))
        out = np.empty(n.value, dtype=np.uint32)
        if n.value:
            check(lib.imc_seqset_run_tokens(self._handle, int(chunk), int(ids), out.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32)),
                                            out.size, ctypes.byref(n), ctypes.byref(fr)))
        return fr.value, out

    def spectral_counts(self):
        """(points served by the spectral form, points served by the plain form) in the last spectral forward call."""
        a, b = ctypes.c_int(), ctypes.c_int()
        check(_lib.load().imc_seqset_spectral_counts(self._handle, ctypes.byref(a), ctypes.byref(b)))
        return a.value, b.value

    def forward_batch_device(self, d_pi, d_T, d_E, d_out, N, K, S, stream=0):
        """Device-resident variant: arguments are raw device pointers (ints); enqueued on `stream`."""
        check(_lib.load().imc_forward_batch_dev(self._handle, int(N), int(K), int(S), int(d_pi), int(d_T), int(d_E),
                                                int(d_out), int(stream)))


class Forwarder(object):
    """One alignment chunk (reference: hmm.py:10-21)."""

    def __init__(self, input_filename, NSYM):
        lib = _lib.load()
        h = _lib.c_vp()
        rc = lib.imc_seq_from_file(os.fsencode(input_filename), int(NSYM), ctypes.byref(h))
        if rc == -5 and not os.path.exists(input_filename):
            raise IOError(lib.imc_last_error().decode())     # the reference's open() raises IOError
        if rc == -5 or rc == -1:
            raise ValueError(lib.imc_last_error().decode())  # int('x') / out-of-range symbol
        check(rc)
        self._finish(_Seq(h), NSYM)

    def _finish(self, seq, NSYM):
        self._seq = seq
        self.NSYM = int(NSYM)
        self._set = None

    @classmethod
    def from_symbols(cls, obs, NSYM):
        """Build from an in-memory array of symbols in [0, NSYM) (what hmm.py:14 parses from the file)."""
        self = cls.__new__(cls)
        lib = _lib.load()
        h = _lib.c_vp()
        obs = np.asarray(obs)
        if obs.dtype == np.uint8:
            obs = np.ascontiguousarray(obs)
            check(lib.imc_seq_create_u8(obs.ctypes.data_as(_lib.c_u8p), obs.size, int(NSYM), ctypes.byref(h)))
        else:
            obs = np.ascontiguousarray(obs, dtype=np.int32)
            check(lib.imc_seq_create(obs.ctypes.data_as(_lib.c_i32p), obs.size, int(NSYM), ctypes.byref(h)))
        self._finish(_Seq(h), NSYM)
        return self

    # -- ingest (reference: scripts/prepare-alignments.py:77-111, pairwise case) --------------------------------
    @classmethod
    def from_pair(cls, sequence1, sequence2):
        """Two aligned sequences (str / bytes) -> symbols 0 equal, 1 different, 2 either base not in ACGT."""
        s1 = sequence1.encode("ascii", "replace") if isinstance(sequence1, str) else bytes(sequence1)
        s2 = sequence2.encode("ascii", "replace") if isinstance(sequence2, str) else bytes(sequence2)
        if len(s1) != len(s2):
            raise ValueError("aligned sequences differ in length (%d vs %d)" % (len(s1), len(s2)))
        self = cls.__new__(cls)
        h = _lib.c_vp()
        check(_lib.load().imc_seq_from_pair(s1, s2, len(s1), ctypes.byref(h)))
        self._finish(_Seq(h), 3)
        return self

    @classmethod
    def from_sequences(cls, *sequences):
        """2, 3 or 4 aligned sequences -> one symbol per column (prepare-alignments.py:77-190): pairs 0/1/2, triplets
        i1 + 4 i2 + 16 i3 or 64, quartets i1 + 4 i2 + 16 i3 + 32 i4 or 128 (the script's weights: NSYM = 160)."""
        seqs = [s.encode("ascii", "replace") if isinstance(s, str) else bytes(s) for s in sequences]
        if len({len(s) for s in seqs}) > 1:
            raise ValueError("aligned sequences differ in length")
        arr = (ctypes.c_char_p * len(seqs))(*seqs)
        self = cls.__new__(cls)
        h = _lib.c_vp()
        check(_lib.load().imc_seq_from_columns(arr, len(seqs), len(seqs[0]) if seqs else 0, ctypes.byref(h)))
        self._finish(_Seq(h), {2: 3, 3: 65, 4: 160}[len(seqs)])
        return self

    @classmethod
    def from_fasta(cls, path, names=None):
        """The named records of a FASTA alignment (two, three or four; or its only two) as one Forwarder."""
        self = cls.__new__(cls)
        h = _lib.c_vp()
        lib = _lib.load()
        if names is not None and len(names) != 2:
            arr = (ctypes.c_char_p * len(names))(*[str(n).encode() for n in names])
            rc = lib.imc_seq_from_fasta_n(os.fsencode(path), arr, len(names), ctypes.byref(h))
            if rc == -5 and not os.path.exists(path):
                raise IOError(lib.imc_last_error().decode())
            if rc in (-5, -1):
                raise ValueError(lib.imc_last_error().decode())
            check(rc)
            self._finish(_Seq(h), {3: 65, 4: 160}[len(names)])
            return self
        n1, n2 = (None, None) if names is None else (str(names[0]).encode(), str(names[1]).encode())
        rc = lib.imc_seq_from_fasta(os.fsencode(path), n1, n2, ctypes.byref(h))
        if rc == -5 and not os.path.exists(path):
            raise IOError(lib.imc_last_error().decode())
        if rc in (-5, -1):
            raise ValueError(lib.imc_last_error().decode())
        check(rc)
        self._finish(_Seq(h), 3)
        return self

    @classmethod
    def from_alignment(cls, path, fmt="fasta", names=None):
        """The named records (two, three or four; or the only two) of an alignment file in format `fmt`: "fasta",
        "phylip", "phylip-relaxed" or "phylip-sequential" -- the <input format> argument of scripts/prepare-alignments.py."""
        self = cls.__new__(cls)
        h = _lib.c_vp()
        lib = _lib.load()
        n = 0 if names is None else len(names)
        arr = (ctypes.c_char_p * max(n, 1))(*[str(x).encode() for x in (names or [])]) if n else None
        rc = lib.imc_seq_from_alignment(os.fsencode(path), str(fmt).encode(), arr, n, ctypes.byref(h))
        if rc == -5 and not os.path.exists(path):
            raise IOError(lib.imc_last_error().decode())
        if rc in (-5, -1):
            raise ValueError(lib.imc_last_error().decode())
        check(rc)
        self._finish(_Seq(h), {0: 3, 2: 3, 3: 65, 4: 160}[n])
        return self

    def save(self, path):
        """Binary container (2 bits per symbol for NSYM <= 4); read it back with Forwarder.load."""
        check(_lib.load().imc_seq_save(self._seq.handle, os.fsencode(path)))

    @classmethod
    def load(cls, path):
        self = cls.__new__(cls)
        h = _lib.c_vp()
        lib = _lib.load()
        rc = lib.imc_seq_load(os.fsencode(path), ctypes.byref(h))
        if rc == -5:
            raise IOError(lib.imc_last_error().decode())
        check(rc)
        n = ctypes.c_int()
        check(lib.imc_seq_nsym(h, ctypes.byref(n)))
        self._finish(_Seq(h), n.value)
        return self

    def write_text(self, path):
        """The reference's text format (prepare-alignments.py:93-105: integers separated by single blanks)."""
        with open(path, "w", 1 << 16) as f:
            sym = self._seq.symbols()
            for a in range(0, sym.size, 1 << 20):
                f.write(" ".join(map(str, sym[a:a + (1 << 20)].tolist())))
                if a + (1 << 20) < sym.size:
                    f.write(" ")

    # -- legacy pyZipHMM constructors -------------------------------------------------------------
    @classmethod
    def fromSequence(cls, seqFilename, alphabetSize, minNoEvals=500):
        """pyZipHMM.Forwarder.fromSequence (scripts/prepare-alignments.py:201).  minNoEvals tuned zipHMM's
        CPU compression and has no meaning here; it is accepted and ignored."""
        return cls(seqFilename, alphabetSize)

    @classmethod
    def fromDirectory(cls, directory):
        """pyZipHMM.Forwarder.fromDirectory (scripts/ils-isolation-model.py:112).  The legacy directory
        schema is not in the reference; only the file names are evidenced (.gitignore:2-4,38-40):
        read `original_sequence` (same integer text format) and take the alphabet size from the first
        integer of `data_structure` when it parses, else max symbol + 1."""
        seq_file = os.path.join(directory, "original_sequence")
        if not os.path.exists(seq_file):
            raise IOError("no 'original_sequence' in zipHMM directory %r" % (directory,))
        nsym = None
        ds = os.path.join(directory, "data_structure")
        if os.path.exists(ds):
            try:
                with open(ds) as f:
                    for tok in f.read().split():
                        if tok.lstrip("-").isdigit():
                            nsym = int(tok)
                            break
            except (IOError, OSError):
                nsym = None
        with open(seq_file) as f:
            obs = np.array(f.read().split(), dtype=np.int64)
        if nsym is None or nsym <= int(obs.max(initial=0)):
            nsym = int(obs.max(initial=0)) + 1
        return cls.from_symbols(obs.astype(np.int32), nsym)

    # -- reference attributes (hmm.py:15-16): the zipHMM-style re-encoding this library computed on the host
    # (position 0 is kept as a raw symbol, the rest are dictionary ids; sym2pair[i] = parts of id NSYM + i). ----
    @property
    def new_obs(self):
        from .ziphmm import _tag
        n = self._seq.length
        if n == 0:
            return _tag(np.zeros(0, dtype=np.int32), self)
        first = self._seq.symbols()[:1].astype(np.int32)
        return _tag(np.concatenate([first, self._as_set().zip_tokens(0).astype(np.int32)]), self)

    @property
    def sym2pair(self):
        return self._as_set().zip_pairs().astype(np.int32)

    @property
    def new_nsyms(self):
        return self._as_set().zip_info()["ids_available"]

    def __len__(self):
        return self._seq.length

    def _as_set(self):
        if self._set is None:
            self._set = ForwarderSet([self])
        return self._set

    def forward(self, init_probs, trans_probs, emission_probs):
        """log P(sequence | pi, T, E) as a python float (hmm.py:19-21)."""
        return self._as_set().forward(init_probs, trans_probs, emission_probs)

    def forward_batch(self, pis, Ts, Es):
        """The same for N parameter points at once -> float64[N]."""
        return self._as_set().forward_batch(pis, Ts, Es)
