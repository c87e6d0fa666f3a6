"""The reference's own known-answer tests for the break points (the only KATs it holds on this path):
/root/reference/tests/IMCoalHMM/break_points_tests.py:39-48, 93-98, 143-162, with the literals copied from there,
against the library's host form (CPU) and against what the model-build kernel computes on the device (GPU)."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN


def test_exp_break_points_reference_literals():
    from imcoalhmm_b200.break_points import exp_break_points
    # break_points_tests.py:22-36: first point is the offset, points scale with 1 / coal_rate
    for n in range(1, 5):
        assert exp_break_points(n, 1.0)[0] == 0.0
        for off in range(10):
            assert exp_break_points(n, 1.0, float(off))[0] == float(off)
        for coal in range(1, 10):
            np.testing.assert_allclose(exp_break_points(n, float(coal)), exp_break_points(n, 1.0) / coal, rtol=0, atol=1e-7)
    # break_points_tests.py:39-48 (assertListEqual: exact)
    assert list(exp_break_points(5, 1.0)) == [0.0, 0.22314355131420976, 0.51082562376599072, 0.916290731874155, 1.6094379124341005]
    assert list(exp_break_points(10, 2.0, -100.0)) == [
        -100.0, -99.947319742171089, -99.888428224342888, -99.821662528030629, -99.744587188117009,
        -99.653426409720026, -99.541854634062929, -99.398013597837036, -99.195281043782956, -98.848707453502982]


def test_uniform_break_points_reference_literals():
    from imcoalhmm_b200.break_points import uniform_break_points
    for n in range(1, 5):
        for start in range(10):
            for end in range(start + 1, 10):
                pts = uniform_break_points(n, float(start), float(end))
                assert pts[0] == float(start) and (pts < float(end)).all()
    # break_points_tests.py:93-98
    assert list(uniform_break_points(7, 1.0, 50.0)) == [1.0, 8.0, 15.0, 22.0, 29.0, 36.0, 43.0]
    assert list(uniform_break_points(10, -20.0, 100.0)) == [-20.0, -8.0, 4.0, 16.0, 28.0, 40.0, 52.0, 64.0, 76.0, 88.0]


def test_psmc_break_points_reference_literals():
    from imcoalhmm_b200.break_points import psmc_break_points
    for n in range(1, 5):
        for t_max in range(0, 50, 5):
            for mu_m in range(10):
                for offset in range(0, 100, 20):
                    mu = mu_m / 100000.0
                    pts = psmc_break_points(n, float(t_max), mu, float(offset))
                    assert pts[0] == float(offset)
                    for i in range(1, n):
                        assert pts[i] >= float(offset)
                        assert pts[i] == pytest.approx(offset + 0.1 * ((1.0 + 10.0 * t_max * mu) ** (float(i) / n) - 1.0), abs=1e-7)
    # break_points_tests.py:143-162
    assert psmc_break_points(4) == [0.0, 3.7499997995738e-09, 7.499999710169902e-09, 1.124999979840169e-08]
    assert psmc_break_points(4, 5, 1) == [0.0, 0.16723451177837886, 0.614142842854285, 1.8084361395018835]
    assert psmc_break_points(4, 50, 0.1) == [0.0, 0.16723451177837886, 0.614142842854285, 1.8084361395018835]
    assert len(psmc_break_points()) == 64


def test_reference_generated_values():
    """tests/golden/break_points.json: values produced by running the reference's own functions (tools/gen_golden.py)."""
    from imcoalhmm_b200.break_points import exp_break_points, psmc_break_points, uniform_break_points
    g = json.load(open(os.path.join(GOLDEN, "break_points.json")))
    np.testing.assert_allclose(exp_break_points(10, 2000.0, 0.001), g["exp_10_2000_0.001"], rtol=1e-15)
    np.testing.assert_allclose(uniform_break_points(10, 0.001, 0.002), g["uniform_10_0.001_0.002"], rtol=1e-15)
    np.testing.assert_allclose(psmc_break_points(40), g["psmc_40"], rtol=1e-15)
    np.testing.assert_allclose(psmc_break_points(40, offset=0.001), g["psmc_40_off"], rtol=1e-15)


@pytest.mark.gpu
def test_device_break_points_match_the_reference():
    """What model_params_kernel computes per parameter point == the reference's functions at the models' arguments
    (isolation_model.py:117, isolation_with_migration_model.py:158-161, variable_coalescence_rate_isolation_model.py:163-167)."""
    import imcoalhmm_b200 as m
    from imcoalhmm_b200.break_points import exp_break_points, psmc_break_points, uniform_break_points
    g = json.load(open(os.path.join(GOLDEN, "break_points.json")))
    bp = m.IsolationModel(10).break_points(np.array([[0.001, 2000.0, 0.4], [0.002, 1500.0, 0.4], [-1.0, 1.0, 1.0]]))
    np.testing.assert_allclose(bp[0], g["exp_10_2000_0.001"], rtol=1e-15)
    np.testing.assert_allclose(bp[1], exp_break_points(10, 1500.0, 0.002), rtol=1e-15)
    assert np.isnan(bp[2]).all()                                   # invalid parameters (model.py:32-42)
    bp = m.IsolationMigrationModel(10, 10).break_points(np.array([[0.001, 0.001, 2000.0, 0.4, 200.0]]))[0]
    np.testing.assert_allclose(bp[:10], g["uniform_10_0.001_0.002"], rtol=1e-15)          # migration period: tau1 .. tau1 + tau2
    np.testing.assert_allclose(bp[10:], exp_break_points(10, 2000.0, 0.002), rtol=1e-15)  # ancestral period
    bp = m.VariableCoalescenceRateIsolationModel([4] * 10, est_split=True).break_points(
        np.array([[0.001] + [1000.0] * 10 + [0.4]]))[0]
    np.testing.assert_allclose(bp, g["psmc_40_off"], rtol=1e-15)
    bp = m.VariableCoalescenceRateIsolationModel([4] * 10, est_split=False).break_points(np.array([[1000.0] * 10 + [0.4]]))[0]
    np.testing.assert_allclose(bp, g["psmc_40"], rtol=1e-15)
    assert psmc_break_points(40) == list(bp)
    assert uniform_break_points(3, 0.0, 3.0).tolist() == [0.0, 1.0, 2.0]
