"""Per-CTA time line of one launch of the sliced SpMV kernel (bb_spmv_timeline): where the time of a launch goes --
launch skew, dependency wait, window staging, strip streaming, tail imbalance.

    python scripts/spmv_timeline.py [C4shard8|C4|C3] [cold]
"""
import ctypes
import os
import sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from bayesbridge_b200 import _lib
from bayesbridge_b200.design_matrix import GpuSparseDesignMatrix

wl = sys.argv[1] if len(sys.argv) > 1 else 'C4shard8'
cold = len(sys.argv) > 2 and sys.argv[2] == 'cold'
n, p, dens = bench.WORKLOADS[wl]
X, y = bench.generate_rows(range(bench.N_BLOCKS), n, p, dens)
ctx = _lib.Context.default()
D = GpuSparseDesignMatrix(X, center_predictor=True, add_intercept=True, ctx=ctx)
D.time_kernel('spmv_dot', reps=2, flush_l2=False)        # fills the gather vectors with test data
lib = _lib.load()
W = 40
for which, name in ((0, 'dot'), (1, 'Tdot')):
    out = np.zeros(W * 256, dtype=np.uint64)
    ncta = ctypes.c_int()
    _lib.check(lib.bb_spmv_timeline(D._mat, which, int(cold), out.ctypes.data_as(ctypes.c_void_p), out.size, ctypes.byref(ncta)))
    t = out[:W * ncta.value].reshape(ncta.value, W).astype(np.int64)
    t0 = t[:, 0].min()
    entry, waited, staged, end, nsec = (t[:, 0] - t0) / 1e3, (t[:, 1] - t0) / 1e3, (t[:, 2] - t0) / 1e3, (t[:, 3] - t0) / 1e3, t[:, 4]
    wend = (t[:, 8:] - t0) / 1e3
    q = lambda a: 'min %6.2f  med %6.2f  p90 %6.2f  max %6.2f' % (a.min(), np.median(a), np.percentile(a, 90), a.max())
    print('== %s %s (%s, %d CTAs, %s) -- us since the first CTA entered' % (wl, name, X.shape, ncta.value, 'cold' if cold else 'warm'))
    print('CTA entry              ', q(entry))
    print('first window staged    ', q(staged), '  (staging alone: ' + q(staged - waited) + ')')
    print('CTA end                ', q(end))
    print('CTA duration           ', q(end - entry))
    print('streaming (end-staged) ', q(end - staged))
    print('sections per CTA       ', np.bincount(nsec.astype(int)).tolist())
    wl_ = wend.max(axis=1) - np.median(wend, axis=1)
    print('warp-end spread in CTA (max - median)', q(wl_))
    one = end - entry
    print('CTAs with 1 section: duration', q(one[nsec == 1]) if (nsec == 1).any() else '-')
    print('CTAs with >1 section: duration', q(one[nsec > 1]) if (nsec > 1).any() else '-')
    slow = np.argsort(end)[-5:]
    print('slowest CTAs:', [(int(c), round(float(end[c]), 2), int(nsec[c])) for c in slow])
