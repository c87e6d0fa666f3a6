"""The library's C++ state-space enumeration against the reference's own (tests/golden/statespaces.json, generated
from state_spaces.py through the py3 shim).  Host-only: runs without a GPU."""
import json
import os

import numpy as np

from conftest import GOLDEN


def _labels(edges, single):
    out = set()
    for s, d, l in edges:
        if l in (0, 1):
            lab = ("C", 0 if single else (1 if l == 0 else 2))
        elif l == 2:
            lab = ("R",)
        else:
            lab = ("M", 1 if l == 3 else 2)
        out.add((int(s), int(d), lab))
    return out


def test_state_spaces_match_reference():
    from imcoalhmm_b200.models import describe_state_space
    ref_all = json.load(open(os.path.join(GOLDEN, "statespaces.json")))
    expect = {"Isolation": (4, 8, [4, 0, 0, 0]), "Single": (15, 44, [7, 3, 3, 2]), "Migration": (94, 466, [56, 16, 16, 6])}
    for sp, name in enumerate(["Isolation", "Single", "Migration"]):
        d, ref = describe_state_space(sp), ref_all[name]
        assert (len(d["states"]), len(d["edges"]), d["counts"]) == expect[name]
        assert d["states"] == ref["states"]
        theirs = set()
        for s, t, p1, p2, dst in ref["edges"]:
            theirs.add((s, dst, ("C", p1) if t == "C" else (("R",) if t == "R" else ("M", p1))))
        assert _labels(d["edges"], name == "Single") == theirs
        for k, key in enumerate(["begin", "left", "right", "end"]):
            assert sorted(np.nonzero(d["classes"] == k)[0].tolist()) == ref[key]
        for key in ("i11_index", "i12_index", "i22_index"):
            if key in ref:
                assert d[key] == ref[key]


def test_per_label_edge_counts():
    """Migration {R11:33,R22:33,M12:133,M21:133,C11:67,C22:67}; Single {R:13, C:31}."""
    from imcoalhmm_b200.models import describe_state_space
    mig = describe_state_space(2)["edges"][:, 2]
    assert np.bincount(mig, minlength=5).tolist() == [67, 67, 66, 133, 133]
    single = describe_state_space(1)["edges"][:, 2]
    assert np.bincount(single, minlength=5).tolist() == [31, 0, 13, 0, 0]
    iso = describe_state_space(0)["edges"][:, 2]
    assert np.bincount(iso, minlength=5).tolist() == [2, 2, 4, 0, 0]


def test_model_shapes_and_argument_checks():
    import imcoalhmm_b200 as m
    assert (m.IsolationModel(10).no_states_total, m.IsolationModel(10).no_parameters) == (10, 3)
    im = m.IsolationMigrationModel(10, 10)
    assert (im.no_states_total, im.no_parameters) == (20, 5)
    ps = m.VariableCoalescenceRateIsolationModel([4] * 10, est_split=True)
    assert (ps.no_states_total, ps.no_parameters) == (40, 12)
    vm = m.VariableCoalAndMigrationRateModel(m.VariableCoalAndMigrationRateModel.INITIAL_12, [4] * 10)
    assert (vm.no_states_total, vm.no_parameters) == (40, 41)
    ep = m.IsolationMigrationEpochsModel(2, 3, 3)
    assert (ep.no_states_total, ep.no_parameters) == (12, 10)
    import pytest
    with pytest.raises(m.IMCError):
        m.IsolationMigrationModel(1, 5)      # the reference's joint[0,0] breaks for one migration state
    assert im.valid_parameters(np.array([1e-3, 1e-3, 2000.0, 0.4, 200.0]))
    assert not im.valid_parameters(np.array([1e-3, -1e-3, 2000.0, 0.4, 200.0]))
    with pytest.raises(AssertionError):
        im.valid_parameters([1.0, 1.0, 1.0, 1.0, 1.0])     # model.py:40 asserts an ndarray
