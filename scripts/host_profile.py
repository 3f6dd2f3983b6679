"""cProfile of the Gibbs loop (host side) on a bench workload. Usage: host_profile.py [C3|C4] [steps]"""
import sys, os, time, cProfile, pstats, io
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, warnings
warnings.simplefilter('ignore')
import bench
import bayesbridge_b200 as bb
from bayesbridge_b200 import _lib
from bayesbridge_b200.design_matrix import GpuSparseDesignMatrix
wl = sys.argv[1] if len(sys.argv) > 1 else 'C3'
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
n, p, dens = bench.WORKLOADS[wl]
ctx = _lib.Context.default()
t0 = time.time(); X, y = bench.generate_rows(range(bench.N_BLOCKS), n, p, dens); print('gen', time.time() - t0, X.nnz, flush=True)
t0 = time.time(); D = GpuSparseDesignMatrix(X, center_predictor=True, add_intercept=True, ctx=ctx); print('upload', time.time() - t0, flush=True)
bridge = bb.BayesBridge(bb.RegressionModel(y, D, family='logit'), bb.RegressionCoefPrior(bridge_exponent=.5))
t0 = time.time(); _, info = bridge.gibbs(n_iter=5, coef_sampler_type='cg', seed=0); print('init+5', time.time() - t0, info['_init_optim_info'], flush=True)
ctx.reset_device_ms()
pr = cProfile.Profile(); pr.enable(); t0 = time.time()
s, info2 = bridge.gibbs_resume(info, steps)
wall = time.time() - t0; pr.disable()
print('wall/step ms', 1000 * wall / steps, 'device/step ms', ctx.device_ms() / steps, 'n_cg', info2['_reg_coef_sampling_info']['n_cg_iter'])
st = io.StringIO(); pstats.Stats(pr, stream=st).sort_stats('cumulative').print_stats(25); print(st.getvalue()[:6000])
