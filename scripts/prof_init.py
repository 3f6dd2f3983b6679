"""Where does the chain initialisation go?  cProfile of BayesBridge.gibbs(n_iter=1) on a bench workload (GPU box)."""
import cProfile, pstats, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
import bayesbridge_b200 as bb
from bayesbridge_b200 import _lib
from bayesbridge_b200.design_matrix import GpuSparseDesignMatrix
wl = sys.argv[1] if len(sys.argv) > 1 else 'C4'
n, p, dens = bench.WORKLOADS[wl]
X, y = bench.generate_rows(range(bench.N_BLOCKS), n, p, dens)
ctx = _lib.Context.default()
t0 = time.time(); D = GpuSparseDesignMatrix(X, center_predictor=True, add_intercept=True, ctx=ctx); print('design', time.time() - t0, D.build_seconds)
model = bb.RegressionModel(y, D, family='logit')
bridge = bb.BayesBridge(model, bb.RegressionCoefPrior(bridge_exponent=.5))
pr = cProfile.Profile(); pr.enable()
s, info = bridge.gibbs(n_iter=1, n_burnin=0, coef_sampler_type='cg', seed=0)
pr.disable()
print('init_runtime', info['init_runtime'], 'runtime', info['runtime'], info['_init_optim_info'])
pstats.Stats(pr).sort_stats('cumulative').print_stats(25)
