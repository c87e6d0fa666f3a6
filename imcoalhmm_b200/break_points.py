"""Break points between HMM intervals (reference: /root/reference/src/IMCoalHMM/break_points.py:9-30, 60-78, 81-108).

Same names, arguments and defaults as the reference; the values come from the library's host form of the three functions
(`imc_break_points`), which is built from the very expressions the batched model-build kernel uses on the device
(`csrc/model_host.inl`, `csrc/model_kernels.cuh:model_params_kernel`).  `trunc_exp_break_points` (break_points.py:33-58)
raises TypeError upstream (list + float) and no model calls it: not provided.
"""
import numpy as np

from . import _lib
from ._lib import check


def _call(kind, n, a, b, c):
    n = int(n)
    out = np.empty(max(n, 1), dtype=np.float64)
    check(_lib.load().imc_break_points(kind, n, float(a), float(b), float(c), out.ctypes.data_as(_lib.c_f64p)))
    return out[:n]


def exp_break_points(no_intervals, coal_rate, offset=0.0):
    """Equal-probability intervals of an exponential with rate coal_rate, shifted by offset -> ndarray[no_intervals]."""
    return _call(0, no_intervals, coal_rate, offset, 0.0)


def uniform_break_points(no_intervals, start, end):
    """Equally spaced points from start (included) towards end (excluded) -> ndarray[no_intervals]."""
    return _call(1, no_intervals, start, end, 0.0)


def psmc_break_points(no_intervals=64, t_max=15, mu=1e-9, offset=0.0):
    """Li & Durbin (2011) break points -> list[no_intervals] (the reference returns a list here)."""
    return [float(x) for x in _call(2, no_intervals, t_max, mu, offset)]
