"""imcoalhmm_b200 -- B200-native drop-in for IMCoalHMM's likelihood hot path.

Public surface (mirrors /root/reference/src/IMCoalHMM for this path only):
    Forwarder(input_filename, NSYM).forward(pi, T, E) -> float            hmm.py:10-21
    Forwarder.fromSequence / Forwarder.fromDirectory                      legacy pyZipHMM constructors
    ForwarderSet(forwarders).forward / .forward_batch                     likelihood.py:33, batched
    Likelihood(model, forwarders)(theta) / .batched(thetas)               likelihood.py:8-33
    maximum_likelihood_estimate(log_likelihood, initial_parameters, ...)  likelihood.py:36-87
    ziphmm.preprocess_raw_observations / ziphmm.zip_forward               hmm.py:16,20-21
    mcmc.BatchedMCMC / MC3 / ParticleSwarm / GeneticAlgorithm             mcmc.py, particle_swarm.py, genetic_algorithm.py: one batched call per step

Importing the package needs the in-tree shared library (python imcoalhmm_b200/build.py); there is no
CPU fallback.  CUDA itself is initialised lazily by the first forward call.
"""
from . import _lib

_lib.load()   # fail loudly at import time if the CUDA library is missing

from ._lib import IMCError, set_option, get_option, kernel_launches, last_forward_kernel, measure_fp64_peak, mma_passes  # noqa: E402
from .hmm import Forwarder, ForwarderSet  # noqa: E402
from .likelihood import Likelihood, maximum_likelihood_estimate  # noqa: E402
from .models import (Model, IsolationModel, IsolationMigrationModel, VariableCoalescenceRateIsolationModel,  # noqa: E402
                     VariableCoalAndMigrationRateModel, IsolationMigrationEpochsModel)
from . import ziphmm  # noqa: E402
from . import mcmc  # noqa: E402

__all__ = ["Forwarder", "ForwarderSet", "Likelihood", "maximum_likelihood_estimate", "IMCError", "ziphmm", "mcmc", "Model", "IsolationModel",
           "IsolationMigrationModel", "VariableCoalescenceRateIsolationModel", "VariableCoalAndMigrationRateModel",
           "IsolationMigrationEpochsModel",
           "set_option", "get_option", "kernel_launches", "last_forward_kernel", "measure_fp64_peak", "mma_passes"]
