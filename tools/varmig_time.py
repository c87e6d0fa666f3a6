import sys, time, numpy as np
sys.path.insert(0, '' + __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))) + '')
import imcoalhmm_b200 as m
g = np.load('' + __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))) + '/tests/golden/model_varmig_i12_4x10.npz')
model = m.VariableCoalAndMigrationRateModel(1, [4] * 10)
th0 = g['theta'][0]
rng = np.random.default_rng(0)
thetas = th0[None, :] * np.exp(0.1 * rng.standard_normal((1024, th0.size)))
thetas[:8] = g['theta'][:8]
for _ in range(2):
    t0 = time.perf_counter()
    pi, T, E, st = model.build_hidden_markov_models(thetas)
    dt = time.perf_counter() - t0
print('varmig K=40 (25 expm of 94x94 per point) x 1024 points: %.1f ms, status ok %d' % (dt * 1e3, (st == 0).sum()))
print('max rel err vs reference-built fixture (pi, T):', np.max(np.abs(pi[:8] - g['pi'][:8]) / g['pi'][:8]), np.max(np.abs(T[:8] - g['T'][:8]) / np.maximum(g['T'][:8], 1e-300)))
