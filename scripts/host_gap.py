"""Where the wall-clock time of a Gibbs step goes on the host side: every library call of the timed loop is wrapped with
a wall-clock timer and the library's own device-time accumulator (CUDA events inside the call), so that
wall - device per call (enqueue latency, polls, copies into pageable memory) and the Python time between calls are visible.

    python scripts/host_gap.py [C4shard8|C4|C1|C3] [steps]
"""
import collections
import os
import sys
import time
import warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.simplefilter('ignore')
import bench
import bayesbridge_b200 as bb
from bayesbridge_b200 import _lib
from bayesbridge_b200.design_matrix import GpuSparseDesignMatrix

wl = sys.argv[1] if len(sys.argv) > 1 else 'C4shard8'
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 40
n, p, dens = bench.WORKLOADS[wl]
X, y = bench.generate_rows(range(bench.N_BLOCKS), n, p, dens)
ctx = _lib.Context.default()
D = GpuSparseDesignMatrix(X, center_predictor=True, add_intercept=True, ctx=ctx)
bridge = bb.BayesBridge(bb.RegressionModel(y, D, family='logit'), bb.RegressionCoefPrior(bridge_exponent=.5))
_, info = bridge.gibbs(n_iter=5, coef_sampler_type='cg', seed=0)

lib = _lib.load()
stat = collections.defaultdict(lambda: [0, 0.0, 0.0])
state = {'last_end': None, 'between': 0.0}


def wrap(name):
    f = getattr(lib, name)

    def g(*a):
        t0 = time.perf_counter()
        if state['last_end'] is not None:
            state['between'] += t0 - state['last_end']
        d0 = ctx.device_ms()
        rc = f(*a)
        d1 = ctx.device_ms()
        t1 = time.perf_counter()
        s = stat[name]
        s[0] += 1; s[1] += 1e3 * (t1 - t0); s[2] += d1 - d0
        state['last_end'] = time.perf_counter()
        return rc
    setattr(lib, name, g)


for nm in ('bb_cg_sample_resident', 'bb_pg_from_coef', 'bb_local_scale_resident', 'bb_state_get', 'bb_state_set', 'bb_state_init',
           'bb_get_obs_prec', 'bb_set_outcome'):
    wrap(nm)
ctx.reset_device_ms()
t0 = time.perf_counter()
_, info2 = bridge.gibbs_resume(info, steps)
wall = 1e3 * (time.perf_counter() - t0)
print('%s: %d steps, wall %.3f ms/step, mean n_cg %.1f' % (wl, steps, wall / steps, info2['_reg_coef_sampling_info']['n_cg_iter'].mean()))
tw = td = 0.0
for nm, (c, w, d) in sorted(stat.items(), key=lambda kv: -kv[1][1]):
    print('  %-26s calls %4d  wall %8.3f ms/step  device %8.3f ms/step  wall-device %7.3f ms/step' % (nm, c, w / steps, d / steps, (w - d) / steps))
    tw += w; td += d
print('  in library calls: wall %.3f device %.3f ms/step; Python between calls %.3f ms/step; rest %.3f ms/step'
      % (tw / steps, td / steps, 1e3 * state['between'] / steps, (wall - tw - 1e3 * state['between']) / steps))
