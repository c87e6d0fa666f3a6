"""Demographic models of the reference, built on the GPU in batches.

Class names, constructor arguments, parameter layouts and `valid_parameters` /
`build_hidden_markov_model` follow the reference's model files so that
`Likelihood(model, forwarders)` and the optimiser / MCMC callers work unchanged:

    IsolationModel(no_hmm_states)                               isolation_model.py:94-122
    IsolationMigrationModel(no_mig_states, no_ancestral_states) isolation_with_migration_model.py:116-164
    VariableCoalescenceRateIsolationModel(intervals, est_split) variable_coalescence_rate_isolation_model.py:90-178
    VariableCoalAndMigrationRateModel(initial, intervals)       variable_migration_model.py:86-181
    IsolationMigrationEpochsModel(no_epochs, no_mig, no_anc)    isolation_with_migration_model_epochs.py:133-211

The arithmetic (rate matrices, matrix exponentials, joint matrix, emissions) runs in
libimcoalhmm_b200.so's batched kernels (csrc/model_kernels.cuh); nothing here computes on the CPU.
"""
import ctypes

import numpy as np

from . import _lib
from ._lib import check


class Model(object):
    """Base class (reference: model.py:11-49)."""
    _kind = None

    def __init__(self, iparams):
        lib = _lib.load()
        arr = np.asarray(iparams, dtype=np.int32)
        h = _lib.c_vp()
        check(lib.imc_model_create(self._kind, arr.ctypes.data_as(_lib.c_i32p), arr.size, ctypes.byref(h)))
        self._handle = h
        k, p = ctypes.c_int(), ctypes.c_int()
        check(lib.imc_model_info(h, ctypes.byref(k), ctypes.byref(p)))
        self.no_states_total, self.no_parameters = k.value, p.value

    def __del__(self):
        h, self._handle = getattr(self, "_handle", None), None
        if h:
            try:
                _lib.load().imc_model_destroy(h)
            except Exception:
                pass

    # noinspection PyMethodMayBeStatic
    def valid_parameters(self, parameters):
        """model.py:32-42: all parameters strictly positive (and an ndarray, as the reference asserts)."""
        assert isinstance(parameters, np.ndarray)
        return bool(np.all(parameters > 0))

    def _thetas(self, thetas):
        thetas = np.ascontiguousarray(np.atleast_2d(np.asarray(thetas, dtype=np.float64)))
        if thetas.shape[1] != self.no_parameters:
            raise ValueError("model takes %d parameters, got %d" % (self.no_parameters, thetas.shape[1]))
        return thetas

    def build_hidden_markov_models(self, thetas, check_joint=True):
        """Batched model.py:44-49: thetas[N,P] -> (pi[N,K], T[N,K,K], E[N,K,3], status[N])."""
        thetas = self._thetas(thetas)
        N, K = thetas.shape[0], self.no_states_total
        pi = np.empty((N, K))
        T = np.empty((N, K, K))
        E = np.empty((N, K, 3))
        status = np.empty(N, dtype=np.int32)
        check(_lib.load().imc_model_build_batch(self._handle, N, thetas.ctypes.data_as(_lib.c_f64p),
                                                pi.ctypes.data_as(_lib.c_f64p), T.ctypes.data_as(_lib.c_f64p),
                                                E.ctypes.data_as(_lib.c_f64p), status.ctypes.data_as(_lib.c_i32p)))
        if check_joint and np.any(status == 2):
            # transitions.py:239: assert_almost_equal(joint.sum(), 1.0)
            raise AssertionError("joint genealogy probabilities do not sum to 1 (7 decimals) for parameter point(s) %s"
                                 % np.nonzero(status == 2)[0].tolist())
        return pi, T, E, status

    def build_hidden_markov_model(self, parameters):
        """model.py:44-49 for one parameter point -> (pi[K], T[K,K], E[K,3]) as plain ndarrays."""
        pi, T, E, _ = self.build_hidden_markov_models(np.asarray(parameters, dtype=np.float64)[None])
        return pi[0], T[0], E[0]

    def break_points(self, thetas):
        """The break points the model-build kernel computed on the device for thetas[N,P] -> float64[N,K] (what the
        reference's build_ctmc_system gets from break_points.py); NaN rows for invalid parameter points."""
        thetas = self._thetas(thetas)
        out = np.empty((thetas.shape[0], self.no_states_total))
        check(_lib.load().imc_model_break_points(self._handle, thetas.shape[0], thetas.ctypes.data_as(_lib.c_f64p),
                                                 out.ctypes.data_as(_lib.c_f64p)))
        return out

    def batched_log_likelihood(self, thetas, forwarder_set, return_status=False):
        """Fused theta -> logL on the device for thetas[N,P]; invalid rows give -inf (likelihood.py:29-30)."""
        thetas = self._thetas(thetas)
        N = thetas.shape[0]
        out = np.empty(N)
        status = np.empty(N, dtype=np.int32)
        check(_lib.load().imc_loglik_batch(self._handle, forwarder_set._handle, N, thetas.ctypes.data_as(_lib.c_f64p),
                                           out.ctypes.data_as(_lib.c_f64p), status.ctypes.data_as(_lib.c_i32p)))
        if np.any(status == 2):
            raise AssertionError("joint genealogy probabilities do not sum to 1 (7 decimals) for parameter point(s) %s"
                                 % np.nonzero(status == 2)[0].tolist())
        return (out, status) if return_status else out

    def batched_log_likelihood_device(self, d_theta, forwarder_set, d_out, N, d_status=0, stream=0):
        """Device pointers in/out, enqueued on `stream`, not synchronised."""
        check(_lib.load().imc_loglik_batch_dev(self._handle, forwarder_set._handle, int(N), int(d_theta), int(d_out),
                                               int(d_status), int(stream)))


class IsolationModel(Model):
    """theta = (split_time, coal_rate, recomb_rate)."""
    _kind = 0

    def __init__(self, no_hmm_states):
        self.no_hmm_states = int(no_hmm_states)
        super(IsolationModel, self).__init__([self.no_hmm_states])


class IsolationMigrationModel(Model):
    """theta = (isolation_time, migration_time, coal_rate, recomb_rate, mig_rate)."""
    _kind = 1

    def __init__(self, no_mig_states, no_ancestral_states):
        self.no_mig_states, self.no_ancestral_states = int(no_mig_states), int(no_ancestral_states)
        super(IsolationMigrationModel, self).__init__([self.no_mig_states, self.no_ancestral_states])


class VariableCoalescenceRateIsolationModel(Model):
    """theta = ([split_time,] coal_rate per epoch..., recomb_rate)."""
    _kind = 2

    def __init__(self, intervals, est_split=False):
        self.intervals, self.est_split = [int(x) for x in intervals], bool(est_split)
        super(VariableCoalescenceRateIsolationModel, self).__init__([int(self.est_split), len(self.intervals)] + self.intervals)


class VariableCoalAndMigrationRateModel(Model):
    """theta = (coal_1[e]..., coal_2[e]..., mig_12[e]..., mig_21[e]..., recomb_rate)."""
    _kind = 3
    INITIAL_11 = 0
    INITIAL_12 = 1
    INITIAL_22 = 2

    def __init__(self, initial_configuration, intervals):
        assert initial_configuration in (0, 1, 2), "We should never reach this point!"
        self.initial_configuration, self.intervals = int(initial_configuration), [int(x) for x in intervals]
        self.no_states = sum(self.intervals)
        super(VariableCoalAndMigrationRateModel, self).__init__([self.initial_configuration, len(self.intervals)] + self.intervals)


class IsolationMigrationEpochsModel(Model):
    """theta = (isolation_time, migration_time, recomb_rate, coal_rates[2e+1]..., mig_rates[e]...)."""
    _kind = 4

    def __init__(self, no_epochs, no_mig_states, no_ancestral_states):
        self.no_epochs, self.no_mig_states, self.no_ancestral_states = int(no_epochs), int(no_mig_states), int(no_ancestral_states)
        super(IsolationMigrationEpochsModel, self).__init__([self.no_epochs, self.no_mig_states, self.no_ancestral_states])


def describe_state_space(space):
    """The library's two-locus ancestry state space: 0 Isolation, 1 Single, 2 Migration (state_spaces.py:7-116).
    Returns a dict in the same canonical form as tests/golden/statespaces.json."""
    lib = _lib.load()
    n, ne = ctypes.c_int(), ctypes.c_int()
    counts, special = (ctypes.c_int * 4)(), (ctypes.c_int * 3)()
    check(lib.imc_statespace_describe(space, ctypes.byref(n), ctypes.byref(ne), counts, special, None, None, None))
    edges = np.empty((ne.value, 3), dtype=np.int32)
    classes = np.empty(n.value, dtype=np.int32)
    lin = np.empty((n.value, 4), dtype=np.uint8)
    check(lib.imc_statespace_describe(space, None, None, None, None, edges.ctypes.data_as(_lib.c_i32p),
                                      classes.ctypes.data_as(_lib.c_i32p), lin.ctypes.data_as(_lib.c_u8p)))

    def samples(mask):
        return tuple(s for s in (1, 2) if mask & s)

    states = []
    for row in lin:
        toks = [(int(t) >> 4, samples((int(t) >> 2) & 3), samples(int(t) & 3)) for t in row if t != 0xff]
        states.append(repr(tuple(sorted(toks))))
    label_names = {0: ("C", 1, 1), 1: ("C", 2, 2), 2: ("R", None, None), 3: ("M", 1, 2), 4: ("M", 2, 1)}
    return {"states": states, "edges": edges, "classes": classes, "label_names": label_names,
            "counts": list(counts), "i11_index": special[0], "i12_index": special[1], "i22_index": special[2]}
