"""Likelihood glue (reference: /root/reference/src/IMCoalHMM/likelihood.py:8-33) + the batched entry.

`Likelihood(model, forwarders)(theta)` keeps the reference's contract: invalid parameters give -inf
(likelihood.py:29-30), otherwise the model's (pi, T, E) is scored on every forwarder and the
log-likelihoods are added (likelihood.py:33).  All forwarders are packed into one ForwarderSet so the
sum is one launch.  `batched(thetas)` scores N parameter points per call.
"""
import numpy as np

from .hmm import Forwarder, ForwarderSet


class Likelihood(object):
    def __init__(self, model, forwarders):
        self.model = model
        if isinstance(forwarders, ForwarderSet):
            self.forwarder_set = forwarders
            self.forwarders = forwarders.forwarders
        else:
            if isinstance(forwarders, Forwarder) or not hasattr(forwarders, "__iter__"):
                forwarders = [forwarders]     # likelihood.py:22-25
            self.forwarders = list(forwarders)
            if all(isinstance(f, Forwarder) for f in self.forwarders):
                self.forwarder_set = ForwarderSet(self.forwarders)
            else:
                self.forwarder_set = None     # foreign forwarder objects: fall back to the reference's loop

    def __call__(self, *parameters):
        if not self.model.valid_parameters(*parameters):
            return -float("inf")
        if self.forwarder_set is not None and len(parameters) == 1 and hasattr(self.model, "batched_log_likelihood"):
            # fused on the device: theta -> (pi, T, E) -> logL without the host round trip of the matrices
            return float(self.model.batched_log_likelihood(np.asarray(parameters[0], dtype=np.float64)[None],
                                                           self.forwarder_set)[0])
        pi, T, E = self.model.build_hidden_markov_model(*parameters)
        if self.forwarder_set is not None:
            return self.forwarder_set.forward(pi, T, E)
        return sum(f.forward(pi, T, E) for f in self.forwarders)

    def batched(self, thetas):
        """float64[N] of log-likelihoods for thetas[N, P]; invalid rows give -inf."""
        thetas = np.atleast_2d(np.asarray(thetas, dtype=np.float64))
        out = np.full(thetas.shape[0], -np.inf)
        if self.forwarder_set is None:            # foreign forwarder objects: one reference-style call per point
            return np.array([self(th) for th in thetas], dtype=np.float64)
        if hasattr(self.model, "batched_log_likelihood"):
            return self.model.batched_log_likelihood(thetas, self.forwarder_set)
        ok = np.array([bool(self.model.valid_parameters(th)) for th in thetas])
        if ok.any():
            hmms = [self.model.build_hidden_markov_model(th) for th in thetas[ok]]
            out[ok] = self.forwarder_set.forward_batch(np.stack([h[0] for h in hmms]),
                                                       np.stack([h[1] for h in hmms]),
                                                       np.stack([h[2] for h in hmms]))
        return out


def maximum_likelihood_estimate(log_likelihood, initial_parameters, optimizer_method="Nelder-Mead", log_file=None,
                                log_param_transform=lambda x: x):
    """Maximum likelihood estimation with a scipy optimiser: same signature, options and logging as the reference
    (likelihood.py:36-87; the scripts call it as `maximum_likelihood_estimate(log_likelihood, init, log_file=...)`).
    Every objective evaluation is one fused theta -> logL call on the GPU (chain-scarce form)."""
    import scipy.optimize
    log_callback = None
    if log_file:
        def log_callback(parameters):
            log_file.write("\t".join(str(param) for param in log_param_transform(parameters)) + "\n")

    def minimize_wrapper(parameters):
        return -log_likelihood(np.asarray(parameters, dtype=np.float64))

    options = {"disp": False}
    if optimizer_method in ["Anneal", "L-BFGS-B", "TNC", "SLSQP"]:       # likelihood.py:76-80: positivity bounds
        bounds = [(0, None)] * len(initial_parameters)
        result = scipy.optimize.minimize(fun=minimize_wrapper, x0=initial_parameters, method=optimizer_method, bounds=bounds,
                                         callback=log_callback, options=options)
    else:
        result = scipy.optimize.minimize(fun=minimize_wrapper, x0=initial_parameters, method=optimizer_method,
                                         callback=log_callback, options=options)
    return result.x
