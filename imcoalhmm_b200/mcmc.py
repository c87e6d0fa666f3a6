"""Batched callers of the likelihood: the reference's samplers and swarm optimiser, restated so that
every step scores ALL chains / particles with one `Likelihood.batched(thetas)` call on the GPU.

Reference: /root/reference/src/IMCoalHMM/mcmc.py (priors :16-56, MCMC :59-97, MC3 :147-196) and
particle_swarm.py:102-214.  The reference runs one process per MC3 chain, each with its own Forwarders
(mcmc.py:113-136), and one likelihood call per proposal; here the chains live in one process and advance in
lock-step.  Proposal, acceptance and swap rules are the reference's:

    proposal   theta' = exp(N(log theta, proposal_sd))                      mcmc.py:34-36
    accept     post' > post  or  u < exp(post'/T - post/T)                  mcmc.py:87-92
    swap i,j   new > cur  or  u < exp(new - cur), posteriors tempered       mcmc.py:178-187
    heat       T_0 = 1, T_k = k * temperature_scale                         mcmc.py:158-162
"""
import numpy as np
from scipy.stats import expon, norm


class LogNormPrior(object):
    """Log-normal prior, random walk in log space (mcmc.py:16-36).  Methods accept scalars or arrays."""

    def __init__(self, log_mean, proposal_sd=None):
        self.log_mean = log_mean
        self.proposal_sd = 0.1 if proposal_sd is None else proposal_sd

    def pdf(self, x):
        return norm.pdf(np.log(x), loc=self.log_mean)

    def sample(self, size=None, rng=None):
        rng = np.random.default_rng() if rng is None else rng
        return np.exp(rng.normal(self.log_mean, 1.0, size=size))

    def proposal(self, x, rng=None):
        rng = np.random.default_rng() if rng is None else rng
        return np.exp(rng.normal(np.log(x), self.proposal_sd))


class ExpLogNormPrior(object):
    """Exponential prior, random walk in log space (mcmc.py:39-56)."""

    def __init__(self, mean, proposal_sd=None):
        self.mean = mean
        self.proposal_sd = 0.1 if proposal_sd is None else proposal_sd

    def pdf(self, x):
        return expon.pdf(x, scale=self.mean)

    def sample(self, size=None, rng=None):
        rng = np.random.default_rng() if rng is None else rng
        return rng.exponential(self.mean, size=size)

    def proposal(self, x, rng=None):
        rng = np.random.default_rng() if rng is None else rng
        return np.exp(rng.normal(np.log(x), self.proposal_sd))


def _batched(log_likelihood):
    """thetas[N,P] -> float64[N]; uses .batched when the callable has it (imcoalhmm_b200.Likelihood)."""
    if hasattr(log_likelihood, "batched"):
        return log_likelihood.batched
    return lambda thetas: np.array([log_likelihood(th) for th in thetas], dtype=np.float64)


class BatchedMCMC(object):
    """`no_chains` independent Metropolis chains (mcmc.py:59-97 each) advancing in lock-step: one batched
    likelihood call per step.  `temperatures` (scalar or [no_chains]) temper the acceptance as in MCMC.step."""

    def __init__(self, priors, log_likelihood, thinning, no_chains, rng=None):
        self.priors, self.thinning, self.no_chains = list(priors), int(thinning), int(no_chains)
        self.rng = np.random.default_rng() if rng is None else rng
        self._loglik = _batched(log_likelihood)
        self.likelihood_calls = 0
        self.current_theta = np.stack([p.sample(size=self.no_chains, rng=self.rng) for p in self.priors], axis=1)
        self.current_prior = self.log_prior(self.current_theta)
        self.current_likelihood = self._score(self.current_theta)
        self.current_posterior = self.current_prior + self.current_likelihood

    def _score(self, thetas):
        self.likelihood_calls += 1
        return np.asarray(self._loglik(np.ascontiguousarray(thetas)), dtype=np.float64)

    def log_prior(self, thetas):
        """mcmc.py:70-78 per row: -inf as soon as one density is non-positive."""
        pdf = np.stack([np.asarray(p.pdf(thetas[:, i]), dtype=np.float64) for i, p in enumerate(self.priors)], axis=1)
        with np.errstate(divide="ignore", invalid="ignore"):
            out = np.log(pdf).sum(axis=1)
        out[~(pdf > 0.0).all(axis=1)] = -np.inf
        return out

    def step(self, temperatures=1.0):
        new_theta = np.stack([p.proposal(self.current_theta[:, i], rng=self.rng) for i, p in enumerate(self.priors)], axis=1)
        new_prior = self.log_prior(new_theta)
        new_likelihood = self._score(new_theta)
        new_posterior = new_prior + new_likelihood
        T = np.broadcast_to(np.asarray(temperatures, dtype=np.float64), (self.no_chains,))
        with np.errstate(over="ignore", invalid="ignore"):
            ratio = np.exp(new_posterior / T - self.current_posterior / T)
        accept = (new_posterior > self.current_posterior) | (self.rng.random(self.no_chains) < ratio)
        accept &= ~np.isnan(new_posterior)
        self.current_theta[accept] = new_theta[accept]
        self.current_prior[accept] = new_prior[accept]
        self.current_likelihood[accept] = new_likelihood[accept]
        self.current_posterior[accept] = new_posterior[accept]
        return accept

    def sample(self, temperatures=1.0):
        for _ in range(self.thinning):
            self.step(temperatures)
        return self.current_theta.copy(), self.current_prior.copy(), self.current_likelihood.copy(), self.current_posterior.copy()


class MC3(object):
    """Metropolis-coupled MCMC (mcmc.py:147-196) with all heated chains in one process.  `log_likelihood` replaces
    the reference's (input_files, model) pair: build it once with Likelihood(model, forwarders)."""

    def __init__(self, priors, log_likelihood, no_chains, thinning, switching, temperature_scale, rng=None):
        self.no_chains, self.thinning, self.switching = int(no_chains), int(thinning), int(switching)
        self.temperature_scale = temperature_scale
        self.rng = np.random.default_rng() if rng is None else rng
        self.chains = BatchedMCMC(priors, log_likelihood, self.switching, self.no_chains, rng=self.rng)
        self.order = np.arange(self.no_chains)      # order[k] = chain currently at temperature slot k

    def chain_temperature(self, chain_no):
        return 1.0 if chain_no == 0 else chain_no * self.temperature_scale

    def sample(self):
        temps_by_slot = np.array([self.chain_temperature(k) for k in range(self.no_chains)])
        for _ in range(int(float(self.thinning) / self.switching)):
            temps = np.empty(self.no_chains)
            temps[self.order] = temps_by_slot
            self.chains.sample(temps)
            i, j = self.rng.integers(0, self.no_chains, size=2)
            if i != j:
                pi_, pj_ = self.chains.current_posterior[self.order[i]], self.chains.current_posterior[self.order[j]]
                ti, tj = temps_by_slot[i], temps_by_slot[j]
                current = pi_ / ti + pj_ / tj
                new = pj_ / ti + pi_ / tj
                with np.errstate(over="ignore", invalid="ignore"):
                    if new > current or self.rng.random() < np.exp(new - current):
                        self.order[i], self.order[j] = self.order[j], self.order[i]
        c = self.order[0]
        return (self.chains.current_theta[c].copy(), self.chains.current_prior[c], self.chains.current_likelihood[c],
                self.chains.current_posterior[c])

    def terminate(self):
        """Kept for the reference's call sequence (mcmc.py:194-196); there are no child processes here."""


class ParticleSwarm(object):
    """Particle swarm (particle_swarm.py:102-214) with the whole swarm scored per iteration in one batched call.
    Positions live in the unit cube like the reference's; `transform` maps a row of positions to model parameters
    (the reference's scripts wrap their likelihood the same way).  The swarm best is updated once per iteration
    (synchronous variant; the reference updates it particle by particle inside the loop)."""

    def __init__(self, particle_count=50, omega=0.9, phi_particle=0.3, phi_swarm=0.1, max_initial_velocity=0.02,
                 max_iterations=None, rng=None):
        self.particle_count, self.omega = int(particle_count), omega
        self.phi_particle, self.phi_swarm = phi_particle, phi_swarm
        self.max_initial_velocity, self.max_iterations = max_initial_velocity, max_iterations
        self.rng = np.random.default_rng() if rng is None else rng

    def maximise(self, fitness_function, parameter_count, transform=None, log_function=None):
        score = _batched(fitness_function)

        def fitness(pos):
            f = np.asarray(score(np.ascontiguousarray(pos if transform is None else transform(pos))), dtype=np.float64)
            f[np.isnan(f)] = -np.inf                                   # particle_swarm.py:112-116
            return f
        n, rng = self.particle_count, self.rng
        pos = rng.uniform(0.0, 1.0, size=(n, parameter_count))
        vel = rng.uniform(-self.max_initial_velocity, self.max_initial_velocity, size=(n, parameter_count))
        cur = fitness(pos)
        best_pos, best_fit = pos.copy(), cur.copy()
        g = int(np.argmax(best_fit))
        swarm_pos, swarm_fit = best_pos[g].copy(), best_fit[g]
        iteration = 0
        while self.max_iterations is None or iteration < self.max_iterations:
            iteration += 1
            if log_function is not None and log_function(iteration, swarm_fit, swarm_pos) is False:
                break
            r_particle, r_swarm = rng.uniform(size=(n, 1)), rng.uniform(size=(n, 1))
            vel = self.omega * vel + self.phi_particle * r_particle * (best_pos - pos) + self.phi_swarm * r_swarm * (swarm_pos - pos)
            pos = pos + vel
            cur = fitness(pos)
            better = cur > best_fit
            best_pos[better], best_fit[better] = pos[better], cur[better]
            g = int(np.argmax(best_fit))
            if best_fit[g] > swarm_fit:
                swarm_pos, swarm_fit = best_pos[g].copy(), best_fit[g]
        return swarm_pos, swarm_fit, iteration


class GeneticAlgorithm(object):
    """Genetic algorithm with the reference's default operators (genetic_algorithm.py:690-840: uniform initialisation in
    the unit cube, tournament selection of 75 % breeders with 10 % tournaments, one-point crossover of breeders i and
    i + n/2 with the fitter parent first, Gaussian point mutation (ratio 0.15, sigma 0.01, clipped to [0, 1]), elitism,
    hall of fame) -- with every generation's offspring scored in ONE batched likelihood call instead of one call each
    (genetic_algorithm.py:813-833).  `transform` maps genomes in the unit cube to model parameters."""

    def __init__(self, population_size=100, elite_count=1, hall_of_fame_size=5, max_generations=500, selection_ratio=0.75,
                 tournament_ratio=0.1, point_mutation_ratio=0.15, mutation_sigma=0.01, rng=None):
        assert elite_count < population_size
        self.population_size, self.elite_count, self.hall_of_fame_size = int(population_size), int(elite_count), int(hall_of_fame_size)
        self.max_generations, self.selection_ratio, self.tournament_ratio = max_generations, selection_ratio, tournament_ratio
        self.point_mutation_ratio, self.mutation_sigma = point_mutation_ratio, mutation_sigma
        self.rng = np.random.default_rng() if rng is None else rng

    def maximise(self, fitness_function, genome_length, transform=None, log_function=None):
        assert genome_length >= 2
        score = _batched(fitness_function)
        rng, n = self.rng, self.population_size

        def fitness(genomes):
            f = np.asarray(score(np.ascontiguousarray(genomes if transform is None else transform(genomes))), dtype=np.float64)
            f[np.isnan(f)] = -np.inf
            return f
        pop = rng.uniform(0.0, 1.0, size=(n, genome_length))
        fit = fitness(pop)
        hall = []                                                  # [(fitness, genome)], best first

        def submit(genomes, fits):
            hall.extend(zip(fits.tolist(), [g.copy() for g in genomes]))
            hall.sort(key=lambda x: -x[0])
            del hall[self.hall_of_fame_size:]
        submit(pop, fit)
        generation = 0
        while self.max_generations is None or generation < self.max_generations:
            generation += 1
            if log_function is not None and log_function(generation, hall[0][0], hall[0][1]) is False:
                break
            # tournament selection over windows of the population (genetic_algorithm.py:342-367)
            size = max(1, int(round(n * self.selection_ratio)))
            tsize = max(1, int(round(n * self.tournament_ratio)))
            starts = rng.integers(0, n - tsize + 1, size=size)
            winners = np.array([a + int(np.argmax(fit[a:a + tsize])) for a in starts])
            winners = winners[np.argsort(-fit[winners], kind="stable")]
            elite = np.argsort(-fit, kind="stable")[:self.elite_count]
            n_off = n - self.elite_count
            i = np.arange(n_off) % size
            j = (np.arange(n_off) + size // 2) % size
            a, b = winners[i], winners[j]
            swap = fit[b] > fit[a]                                 # the fitter parent supplies the left part
            left, right = np.where(swap, b, a), np.where(swap, a, b)
            cut = rng.integers(1, genome_length, size=n_off)       # one-point crossover (:426-445)
            cols = np.arange(genome_length)[None, :]
            off = np.where(cols < cut[:, None], pop[left], pop[right])
            mutate = rng.uniform(size=off.shape) < self.point_mutation_ratio        # Gaussian mutation (:620-640)
            off = np.where(mutate, np.clip(off + rng.normal(0.0, self.mutation_sigma, size=off.shape), 0.0, 1.0), off)
            off_fit = fitness(off)                                 # the whole generation in one batched call
            pop = np.concatenate([pop[elite], off])
            fit = np.concatenate([fit[elite], off_fit])
            submit(off, off_fit)
        return hall[0][1], hall[0][0], generation
