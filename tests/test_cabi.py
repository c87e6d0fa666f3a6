"""The C-ABI library loads and exports every symbol include/imcoalhmm_b200.h declares (no GPU needed)."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "imcoalhmm_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(imc_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound():
    import imcoalhmm_b200 as m
    lib = ctypes.CDLL(m._lib.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 20
    for name in names:
        assert hasattr(lib, name), "header declares %s but the library does not export it" % name
    assert set(names) == set(m._lib.EXPORTS), "ctypes binding and header disagree"
    assert m._lib.load().imc_version() >= 100


def test_host_side_sequence_handling_without_gpu(tmp_path):
    import imcoalhmm_b200 as m
    p = tmp_path / "a.txt"
    p.write_text("0 0 1 2\n0\t1  2 0")          # any whitespace (hmm.py:13-14)
    f = m.Forwarder(str(p), 3)
    assert len(f) == 8 and f.NSYM == 3 and f.new_nsyms == 3      # too short for any pair to pay off
    assert f.new_obs.tolist() == [0, 0, 1, 2, 0, 1, 2, 0] and f.new_obs.dtype == np.int32
    assert f.sym2pair.shape == (0, 2)
    g = m.Forwarder.from_symbols(np.tile(np.array([0, 0, 0, 1], dtype=np.int32), 200), 3)
    assert g.new_nsyms > 3 and g.sym2pair.shape == (g.new_nsyms - 3, 2) and len(g.new_obs) < 200
    from imcoalhmm_b200 import ziphmm
    assert ziphmm._expand(g.sym2pair, g.new_obs, 3, g.new_nsyms).tolist() == [0, 0, 0, 1] * 200
    with pytest.raises(IOError):
        m.Forwarder(str(tmp_path / "missing.txt"), 3)
    bad = tmp_path / "bad.txt"
    bad.write_text("0 1 x 2")
    with pytest.raises(ValueError):
        m.Forwarder(str(bad), 3)
    bad.write_text("0 1 - 2")                       # a sign without digits: int("-") raises in the reference (hmm.py:14)
    with pytest.raises(ValueError):
        m.Forwarder(str(bad), 3)
    bad.write_text("0 +1 2\n")                      # int("+1") == 1
    assert m.Forwarder(str(bad), 3).new_obs.tolist() == [0, 1, 2]
    bad.write_text("0 1 3 2")                       # symbol outside [0, NSYM)
    with pytest.raises(ValueError):
        m.Forwarder(str(bad), 3)
    s = m.ForwarderSet([f, m.Forwarder.from_symbols(np.zeros(0, dtype=np.int32), 3)])
    assert s.n_chunks == 2 and s.total_sites == 8


def test_legacy_constructors(tmp_path):
    import imcoalhmm_b200 as m
    d = tmp_path / "zipdir"
    d.mkdir()
    (d / "original_sequence").write_text("0 1 2 2 0")
    (d / "data_structure").write_text("3\n")
    f = m.Forwarder.fromDirectory(str(d))
    assert len(f) == 5 and f.NSYM == 3
    g = m.Forwarder.fromSequence(seqFilename=str(d / "original_sequence"), alphabetSize=3, minNoEvals=500)
    assert g.new_obs.tolist() == [0, 1, 2, 2, 0]
    with pytest.raises(IOError):
        m.Forwarder.fromDirectory(str(tmp_path))


def test_ziphmm_shim_expand():
    from imcoalhmm_b200 import ziphmm
    from oracle import forward as F
    rng = np.random.default_rng(3)
    obs = rng.choice(3, size=4000, p=[0.9, 0.05, 0.05]).astype(np.int32)
    new_obs, sym2pair, new_nsyms = F.zip_preprocess(obs, 3, min_count=4)
    assert new_nsyms > 3
    np.testing.assert_array_equal(ziphmm._expand(sym2pair, new_obs, 3, new_nsyms), obs)


def test_no_cpu_fallback_without_device(have_gpu):
    if have_gpu:
        pytest.skip("a GPU is present")
    import imcoalhmm_b200 as m
    f = m.Forwarder.from_symbols(np.array([0, 1, 2], dtype=np.int32), 3)
    with pytest.raises(m.IMCError) as e:
        f.forward(np.ones(2) / 2, np.ones((2, 2)) / 2, np.ones((2, 3)) / 3)
    assert "no CPU fallback" in str(e.value)


def test_legacy_matrix_objects_are_accepted():
    """pyZipHMM.Matrix-style arguments (getHeight/getWidth/[i,j]; ILS.py:271-276) are converted like ndarrays."""
    from imcoalhmm_b200.hmm import _hmm_arrays

    class Matrix(object):
        def __init__(self, a):
            self.a = np.asarray(a, dtype=float)

        def getHeight(self):
            return self.a.shape[0]

        def getWidth(self):
            return self.a.shape[1]

        def __getitem__(self, ij):
            return self.a[ij]
    rng = np.random.default_rng(0)
    pi, T, E = rng.random((4, 1)), rng.random((4, 4)), rng.random((4, 3))
    a = _hmm_arrays(Matrix(pi), Matrix(T), Matrix(E), batched=False)
    b = _hmm_arrays(pi, T, E, batched=False)
    assert all(np.array_equal(x, y) for x, y in zip(a[:3], b[:3])) and a[3:] == b[3:] == (1, 4, 3)


def test_options_are_validated_on_the_host():
    import imcoalhmm_b200 as m
    for key, good, bad in (("forward_kernel", 4, 5), ("zip_lanes", 32, 16), ("zip_ctas_per_sm", 2, 3), ("zip_max_entries", 256, 257),
                           ("zip_segment_tokens", -1, -2), ("dmma_mtiles", 4, 3)):
        m.set_option(key, good)
        assert m.get_option(key) == good
        with pytest.raises(m.IMCError):
            m.set_option(key, bad)
        m.set_option(key, 0)
        assert m.get_option(key) == 0
    with pytest.raises(m.IMCError):
        m.set_option("no_such_option", 1)
    with pytest.raises(m.IMCError):
        m.get_option("no_such_option")


def test_forward_without_a_gpu_fails_loudly(have_gpu):
    """No CPU fallback: on a box without a usable device every forward call raises instead of computing on the host."""
    import imcoalhmm_b200 as m
    if have_gpu:
        pytest.skip("a GPU is present")
    f = m.Forwarder.from_symbols(np.array([0, 1, 2, 0], dtype=np.int32), 3)
    K = 4
    with pytest.raises(m.IMCError) as e:
        f.forward(np.full(K, 0.25), np.full((K, K), 0.25), np.full((K, 3), 1.0 / 3))
    assert e.value.code == -2 and "no CPU fallback" in str(e.value)
    with pytest.raises(m.IMCError):
        m.IsolationModel(4).build_hidden_markov_model(np.array([1e-3, 1000.0, 0.4]))
