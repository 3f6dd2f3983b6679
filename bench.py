#!/usr/bin/env python
"""Benchmark of the hot path: Gibbs iterations/s of the CG-accelerated sampler (logit), BASELINE.json's metric.

    python bench.py --gpus N --steps K --warmup W            # this repo, N GPUs of one node
    python bench.py --impl reference --steps K --warmup W    # the reference's own CPU sampler (oracle/_ref)

Workload (config.workload = "C4"): BASELINE.json configs[3], the one the north-star target is quoted on --
binary sparse X, n = 1,000,000 x p = 100,000, mean density 0.1 % (nnz ~ 1e8), column frequencies
0.5*Beta(0.5, b) as in the reference's simulate_binary_design (simulate_data.py:100-117), logit outcome,
bridge exponent 0.5, coef_sampler_type='cg'.  X is generated in 50 fixed row blocks (seeded per block), so
the matrix is identical for every N; rank r of N owns a contiguous run of blocks (strong scaling: the total
work is fixed).  One "step" = one full Gibbs iteration (beta | omega by CG, omega | beta by Polya-Gamma,
tau, lambda, log-posterior).  X (2.4 GB as CSR+CSC images) is far larger than the 126 MB L2, so no L2
flush is needed between steps.

value   = K / (device time of the K steps: CUDA events on the library stream, summed over the library
          calls of a step, max over ranks)  -- inputs (X, outcome, omega) resident in HBM.
e2e     = K / (wall time of BayesBridge.gibbs_resume(K), the public API, max over ranks) -- includes the
          per-step host<->device copies of the P-length vectors and all host-side Python.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time
import warnings

import numpy as np
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
warnings.simplefilter('ignore')

N_BLOCKS = 50
WORKLOADS = {
    # name: (n, p, mean density)            binary sparse X, logit outcome (BASELINE configs 1, 3, 4)
    'C4': (1_000_000, 100_000, 0.001),
    'C3': (100_000, 20_000, 0.005),
    'C1': (10_000, 1_000, 0.01),
    'C4shard8': (125_000, 100_000, 0.001),     # one rank's share of C4 at N = 8 (profiling aid)
    'C4quarter': (250_000, 100_000, 0.001),    # two such shares (N = 2 reproduces the N = 8 per-rank load)
    # dense fp64 X, linear outcome (BASELINE config 2); density 1.0 marks the dense family
    'C2': (50_000, 5_000, 1.0),
    'C2small': (5_000, 500, 1.0),
    # dense fp64 X, logit outcome, 16 independent chains advanced in lock-step (BASELINE config 5)
    'C5': (200_000, 2_000, 1.0),
    'C5small': (20_000, 400, 1.0),
}
N_CHAINS = {'C5': 16, 'C5small': 16}


def is_dense(workload):
    return WORKLOADS[workload][2] >= 1.0


# ---- synthetic data (shared by both arms) ------------------------------------------------------
def column_frequencies(p, density, seed=0):
    """simulate_data.py:100-111: f_j = 0.5 * Beta(a, b), a = 0.5, b chosen so that E f_j = density."""
    a, max_freq = 0.5, 0.5
    b = a * (max_freq / density - 1)
    return max_freq * np.random.default_rng(seed).beta(a, b, p)


def true_coef(p):
    beta = np.zeros(p)
    beta[:5], beta[5:10], beta[10:15], beta[15:20] = 1.5, -1.0, 1.0, -0.5
    return beta


def generate_block(block, n, p, freq, seed=0):
    """Rows [n*block/N_BLOCKS, n*(block+1)/N_BLOCKS) of X (binary CSR) and of the binary outcome."""
    lo, hi = n * block // N_BLOCKS, n * (block + 1) // N_BLOCKS
    nb = hi - lo
    rng = np.random.default_rng([seed, block])
    counts = rng.binomial(nb, freq)
    cols = np.repeat(np.arange(p, dtype=np.int64), counts)
    rows = rng.integers(0, nb, cols.size)
    key = np.unique(rows * p + cols)                  # row-major order, duplicates dropped
    rows, cols = key // p, (key % p).astype(np.int32)
    indptr = np.zeros(nb + 1, dtype=np.int32)
    np.cumsum(np.bincount(rows, minlength=nb), out=indptr[1:])
    X = sp.csr_matrix((np.ones(cols.size), cols, indptr), shape=(nb, p))
    eta = -2.0 + X @ true_coef(p)
    y = rng.binomial(1, 1 / (1 + np.exp(-eta)))
    return X, y


def generate_dense_rows(blocks, n, p, seed=0, logit=False):
    """BASELINE config 2: X = standard normal n x p fp64, y = X beta + N(0, 1); generated per row
    block so that the matrix is the same for every number of ranks."""
    Xs, ys = [], []
    beta = true_coef(p)
    for b in blocks:
        lo, hi = n * b // N_BLOCKS, n * (b + 1) // N_BLOCKS
        rng = np.random.default_rng([seed, 1000 + b])
        Xb = rng.standard_normal((hi - lo, p))
        Xs.append(Xb)
        if logit:
            ys.append(rng.binomial(1, 1 / (1 + np.exp(-(-1.0 + Xb @ beta)))).astype(np.float64))
        else:
            ys.append(Xb @ beta + rng.standard_normal(hi - lo))
    return np.ascontiguousarray(np.vstack(Xs)), np.concatenate(ys)


def _block_job(job):
    b, n, p, density = job
    return generate_block(b, n, p, column_frequencies(p, density))


def generate_rows(blocks, n, p, density, logit_dense=False):
    if density >= 1.0:
        return generate_dense_rows(blocks, n, p, logit=logit_dense)
    freq = column_frequencies(p, density)
    blocks = list(blocks)
    # the row blocks are independent (seeded per block): generate them on several host cores (forked numpy-only workers;
    # callers generate BEFORE the CUDA context exists).  The result does not depend on the number of workers.
    world = int(os.environ.get('WORLD_SIZE', '1'))
    workers = min(len(blocks), max(1, (os.cpu_count() or 1) // max(world, 1)), 16)
    if workers > 1 and n * p * density > 2e6 and os.environ.get('BENCH_GEN_WORKERS', '') != '1':
        import multiprocessing as mp
        with mp.get_context('fork').Pool(workers) as pool:
            parts = pool.map(_block_job, [(b, n, p, density) for b in blocks], chunksize=1)
    else:
        parts = [generate_block(b, n, p, freq) for b in blocks]
    X = sp.vstack([q[0] for q in parts], format='csr')
    X.indices = X.indices.astype(np.int32)
    X.indptr = X.indptr.astype(np.int32)
    y = np.concatenate([q[1] for q in parts])
    return X, y


# ---- clocks ------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """Samples SM clocks and throttle reasons during the timed region through NVML, in-process.
    On this driver every NVML query costs tens of milliseconds and serialises with kernel launches (spawning
    nvidia-smi every 200 ms slowed the timed loop 3x, NVML every 250 ms 2x), so the sampler backs off to at
    most ~5 % duty: it sleeps max(period, 20 x the measured query time) between samples."""

    def __init__(self, device, period=0.4):
        super().__init__(daemon=True)
        self.device, self.period, self.samples, self.stop_flag = device, period, [], False
        self.query_ms = []
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(device)
            self.max_sm = pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nvml = None

    def run(self):
        if self.nvml is None:
            return
        nv = self.nvml
        while not self.stop_flag:
            t0 = time.perf_counter()
            try:
                sm = nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM)
                try:
                    reasons = nv.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
                except Exception:
                    reasons = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
                self.samples.append((sm, reasons))
            except Exception:
                pass
            dt = time.perf_counter() - t0
            self.query_ms.append(1000 * dt)
            time.sleep(max(self.period, 20 * dt))

    def summary(self):
        if not self.samples:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['unavailable']}
        nv = self.nvml
        sm = sorted(s[0] for s in self.samples)
        bits = 0
        for s_ in self.samples:
            bits |= s_[1]
        table = [('hw_slowdown', 'nvmlClocksThrottleReasonHwSlowdown'),
                 ('hw_thermal_slowdown', 'nvmlClocksThrottleReasonHwThermalSlowdown'),
                 ('sw_thermal_slowdown', 'nvmlClocksThrottleReasonSwThermalSlowdown'),
                 ('sw_power_cap', 'nvmlClocksThrottleReasonSwPowerCap')]
        reasons = [nm for nm, attr in table if bits & getattr(nv, attr, 0)]
        return {'sm_mhz': float(sm[len(sm) // 2]), 'sm_max_mhz': float(self.max_sm), 'reasons': reasons,
                'samples': len(sm), 'source': 'nvml',
                'query_ms_mean': float(np.mean(self.query_ms)) if self.query_ms else None}


# ---- reference arm / cpu baseline ---------------------------------------------------------------
def import_reference():
    ref_dir = os.path.join(ROOT, 'oracle', '_ref')
    if not os.path.isdir(os.path.join(ref_dir, 'bayesbridge')):
        return None
    sys.path.insert(0, ref_dir)
    import bayesbridge
    return bayesbridge


def plan_reference_iterations(fit, warmup, steps, done_w):
    """How many warm-up and timed iterations the reference arm still runs when only `fit` more iterations fit into its
    wall-clock budget (`done_w` warm-up iterations are already done).  Everything fits: as requested.  Otherwise the warm-up
    is kept and timed iterations are cut first; below half the requested steps both are cut (never fewer than 2 timed).
    Returns (warmup_run, steps_run), warmup_run counting the iterations already done."""
    if fit >= (warmup - done_w) + steps:
        return warmup, steps
    steps_run = fit - (warmup - done_w)
    if steps_run < max(2, steps // 2):
        steps_run = max(min(steps, 2), min(steps, int(0.8 * fit)))
    warmup_run = done_w + max(0, min(warmup - done_w, fit - steps_run))
    return warmup_run, steps_run


def reference_run(workload, steps, warmup, sample_blocks=None, init_state=None, data=None, sampler='cg', budget_s=None):
    """The reference's own numpy/scipy/Cython sampler (oracle/_ref, the unmodified package built by
    oracle/build_ref.sh) through ITS public API, on the host cores of this box.

    sample_blocks=None (the default, and what `--impl reference` runs): the FULL workload matrix, `warmup` untimed
    Gibbs iterations then `steps` timed ones -- nothing is extrapolated.  sample_blocks=k times it on the first k of
    the 50 row blocks instead (explicit flag only; reported as such, never scaled).
    init_state: {'coef','obs_prec','local_scale','global_scale'} to start the chain from (bayesbridge.py:279-353
    skips the mode search when all four are given) -- used by the bounded cpu_baseline leg, which starts the
    reference from the state the GPU chain has reached so that its CG iteration counts are comparable.
    Returns (iterations/s, description dict)."""
    n, p, density = WORKLOADS[workload]
    t_start = time.time()
    ref = import_reference()
    blocks = range(N_BLOCKS) if sample_blocks is None else range(sample_blocks)
    t_gen = time.time()
    X, y = data if data is not None else generate_rows(blocks, n, p, density, logit_dense=workload in N_CHAINS)
    t_gen = time.time() - t_gen
    frac = X.shape[0] / n
    budget_note = ''
    steps_run, warmup_run = steps, warmup
    if ref is not None:
        kind = 'reference'
        model = ref.RegressionModel(y, X, family='linear' if (is_dense(workload) and workload not in N_CHAINS) else 'logit')
        bridge = ref.BayesBridge(model, ref.RegressionCoefPrior(bridge_exponent=.5))
        kw = dict(n_burnin=0, coef_sampler_type=sampler, seed=0, params_to_save=('global_scale',))
        linear = is_dense(workload) and workload not in N_CHAINS
        if init_state is not None:
            # the reference's initialize_obs_precision takes len() of a given obs_prec (bayesbridge.py:355-360), which a
            # linear model's scalar precision does not have: let it recompute the precision from the coefficients
            kw['init'] = {k: v for k, v in init_state.items() if not (linear and k == 'obs_prec')}

        def advance(info, k):
            # k more Gibbs iterations of the same chain.  The reference's gibbs_resume fails for the linear model (same
            # len() of a scalar): restart from the reached state instead, which skips the mode search (bayesbridge.py:279-353)
            if linear:
                st = info['_markov_chain_state']
                return bridge.gibbs(n_iter=k, **dict(kw, init={q: st[q] for q in ('coef', 'local_scale', 'global_scale')}))[1]
            return bridge.gibbs_resume(info, k)[1]

        t0 = time.time()
        if warmup > 0:
            # Wall-clock guard (budget_s; the driver's per-run limit is a few hundred seconds and a full-size C4 iteration
            # of the reference takes ~20 s): the first warm-up iteration carries the chain initialisation, the second one
            # is timed on its own; if W + K iterations at that pace would overrun the budget, FEWER full-size iterations
            # are run -- reported as such ('steps' / 'warmup' of the line are the counts actually executed).  Nothing is
            # ever extrapolated or run on a sub-sample.
            _, info = bridge.gibbs(n_iter=1, **kw)
            done_w = 1
            if warmup >= 2:
                t1 = time.time()
                info = advance(info, 1)
                t_iter = time.time() - t1
                done_w = 2
                if budget_s is not None:
                    left = budget_s - (time.time() - t_start) - 5.0
                    fit = int(max(left, 0.0) / max(t_iter, 1e-9))
                    warmup_run, steps_run = plan_reference_iterations(fit, warmup, steps, done_w)
                    if (warmup_run, steps_run) != (warmup, steps):
                        budget_note = ('; wall-clock budget %.0f s: %d + %d of the requested %d + %d iterations run (%.1f s each)'
                                       % (budget_s, warmup_run, steps_run, warmup, steps, t_iter))
            if warmup_run > done_w:
                info = advance(info, warmup_run - done_w)
            t1 = time.time()
            info2 = advance(info, steps_run)
            dt = time.time() - t1
        else:
            _, info2 = bridge.gibbs(n_iter=steps, **kw)
            dt = info2['runtime']               # includes the (skipped or trivial) chain initialisation
        n_cg = (float(np.mean(info2['_reg_coef_sampling_info']['n_cg_iter']))
                if steps_run > 0 and sampler == 'cg' else float('nan'))
    else:
        # the oracle port (numpy restatement) when the reference could not be built
        kind = 'port'
        from oracle import cg_oracle as co
        from oracle.rand_port import PolyaGammaPort, TiltedStablePort
        t0 = time.time()
        if is_dense(workload):
            raise RuntimeError('oracle/_ref is not built and the numpy port only covers the logit family')
        coefs, n_cg_arr = co.gibbs_cg_oracle('logit', (y.astype(float), np.ones(len(y))), X, warmup + steps, 0, 0.5,
                                             float('inf'), float('inf'), 0.1, np.ones(p), PolyaGammaPort,
                                             TiltedStablePort, True)
        dt = (time.time() - t0) * steps / max(warmup + steps, 1)
        n_cg = float(n_cg_arr.mean())
    its = steps_run / dt
    nnz_x = int(X.nnz) if sp.issparse(X) else int(X.size)
    what = ('the full %s matrix (%d x %d, nnz %d)' % (workload, X.shape[0], p, nnz_x) if sample_blocks is None else
            'rows 0..%d of %s (%d of %d row blocks, %.0f%% of the rows; NOT scaled to the full problem)'
            % (X.shape[0] - 1, workload, sample_blocks, N_BLOCKS, 100 * frac))
    desc = {
        'kind': kind, 'cores': (os.cpu_count() if is_dense(workload) else 1),
        'sample': '%s; %d timed Gibbs iterations after %d warm-up%s; %s and the Cython PG / tilted-stable '
                  'samplers are single-threaded (host has %d cores)'
                  % (what, steps_run, warmup_run, ('' if init_state is None else ', chain started from the state the GPU chain reached') + budget_note,
                     'numpy BLAS gemv uses all cores; the rest of the sampler' if is_dense(workload)
                     else 'thread counts left at their defaults (numpy BLAS level-1 may spin up all cores; it does not help: C3 measured at '
                          '2.66 s/iteration with 1 thread, 2.83 s with 8); scipy SpMV',
                     os.cpu_count()),
        'full_size': sample_blocks is None, 'mean_n_cg_iter': (None if n_cg != n_cg else n_cg), 'sample_nnz': nnz_x,
        'seconds_per_iteration': dt / steps_run, 'generate_seconds': t_gen,
        'host_cores': os.cpu_count(), 'steps_run': steps_run, 'warmup_run': warmup_run, 'steps_requested': steps, 'warmup_requested': warmup,
    }
    return its, desc


# ---- main ---------------------------------------------------------------------------------------
KERNEL_SOURCES = {'spmv_dot': ('bb_sell.cu',), 'spmv_tdot': ('bb_sell.cu',), 'fused_op': ('bb_dense.cu',), 'op': ('bb_dense.cu',),
                  'batch_op': ('bb_batch.cu',)}


def kernel_version(files=('bb_sell.cu', 'bb_dense.cu', 'bb_batch.cu')):
    """Identifies the source of a roofline kernel (a DRAM-traffic capture is only quoted for the version it was taken on)."""
    import hashlib
    h = hashlib.sha1()
    for f in files:
        h.update(open(os.path.join(ROOT, 'bayesbridge_b200', 'csrc', f), 'rb').read())
    return h.hexdigest()[:12]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=30)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--workload', default='C4', choices=sorted(WORKLOADS))
    ap.add_argument('--sampler', default='cg', choices=['cg', 'cholesky'],
                    help="coef_sampler_type; 'cholesky' is the comparator of BASELINE config 2 (dense workloads)")
    ap.add_argument('--ref-blocks', type=int, default=0,
                    help='time the CPU reference on the first K of the 50 row blocks only (0 = the full workload, the default)')
    ap.add_argument('--ref-budget-s', type=float, default=float(os.environ.get('BENCH_REF_BUDGET_S', '790')),
                    help='--impl reference: wall-clock budget of the whole run; when W + K full-size iterations would overrun it, '
                         'fewer (still full-size) iterations are run and reported (0 = no limit)')
    ap.add_argument('--cpu-baseline-steps', type=int, default=2, help='full-size reference iterations of the cpu_baseline leg')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--clocks', default=os.environ.get('BENCH_CLOCKS', 'nvml'), choices=['nvml', 'none'])
    ap.add_argument('--profile-host', action='store_true', help='cProfile the timed region (stderr)')
    args = ap.parse_args()
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    n, p, density = WORKLOADS[args.workload]
    dense = is_dense(args.workload)
    n_chains = N_CHAINS.get(args.workload, 1)
    family = 'linear' if (dense and n_chains == 1) else 'logit'
    x_bytes = 8.0 * n * p if dense else 2 * 4.0 * density * n * p
    config = {'workload': args.workload, 'family': family, 'n': n, 'p': p, 'mean_density': density,
              'bridge_exponent': 0.5, 'coef_sampler_type': args.sampler, 'n_chains': n_chains,
              'format': ('dense row-major fp64' if dense else 'binary CSR + CSC, int32 indices, fp64 math'),
              'l2': ('X (%.2f GB) %s the 126 MB L2; the roofline kernel is timed with an L2 flush before every launch'
                     % (x_bytes / 1e9, 'exceeds' if x_bytes > 126e6 else 'FITS in'))}

    if args.impl == 'reference':
        if rank != 0:
            return
        value, desc = reference_run(args.workload, args.steps, max(args.warmup, 1), args.ref_blocks or None, sampler=args.sampler,
                                    budget_s=(args.ref_budget_s if args.ref_budget_s > 0 else None))
        print(json.dumps({
            'impl': 'reference', 'metric': 'gibbs_iters_per_sec', 'value': value, 'unit': 'iter/s',
            'n_gpus': args.gpus, 'steps': desc['steps_run'], 'warmup': desc['warmup_run'] if args.warmup > 0 else 0,
            'ms_per_step': 1000.0 / value,
            'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            # the same config object as the GPU arm prints (nnz of the whole matrix included)
            'config': dict(config, nnz=int(desc['sample_nnz'])) if not args.ref_blocks else dict(config, ref_blocks=args.ref_blocks),
            'cpu_baseline': dict(desc, value=value, unit='iter/s'),
            'e2e': {'value': value, 'unit': 'iter/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        }))
        return

    # this rank's row blocks (generated first: the generator forks numpy-only workers, which must not inherit a CUDA context)
    blocks = range(N_BLOCKS * rank // world, N_BLOCKS * (rank + 1) // world)
    X, y = generate_rows(blocks, n, p, density, logit_dense=n_chains > 1)
    row_offset = n * blocks[0] // N_BLOCKS

    import torch
    import torch.distributed as dist
    import bayesbridge_b200 as bb
    from bayesbridge_b200 import _lib
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    ctx = _lib.Context(local_rank)
    if world > 1:
        ctx.init_comm_from_torch()
    t_build = time.perf_counter()
    nnz_local = 0 if dense else int(X.nnz)
    from bayesbridge_b200.design_matrix import GpuSparseDesignMatrix, GpuDenseDesignMatrix
    nnz_local = int(X.size) if dense else nnz_local
    Design = GpuDenseDesignMatrix if dense else GpuSparseDesignMatrix
    design = Design(X, center_predictor=True, add_intercept=True, ctx=ctx,
                    presharded=(world > 1), n_global=n, row_offset=row_offset)
    t_build = time.perf_counter() - t_build
    model = bb.RegressionModel(y, design, family=family)
    bridge = bb.BayesBridge(model, bb.RegressionCoefPrior(bridge_exponent=.5))
    P = design.shape[1]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ctx.sync()

    batch = None
    if n_chains > 1:
        batch = bb.BatchedBayesBridge(model, bb.RegressionCoefPrior(bridge_exponent=.5), n_chains)
        info = {}
    else:
        # chain initialisation + W warm-up steps (untimed)
        _, info = bridge.gibbs(n_iter=max(args.warmup, 1), n_burnin=0, coef_sampler_type=args.sampler, seed=0,
                               params_to_save=('coef', 'global_scale', 'logp'))
    # roofline of the dominant kernel, measured live with CUDA events on the library stream; every timed launch is
    # preceded by an L2 flush (a 512 MB write), i.e. these are cold-cache times; the warm ones are reported beside them
    roof, roof_warm = {}, {}
    kernels = ('fused_op',) if dense else ('spmv_dot', 'spmv_tdot')
    if batch is not None:
        kernels = ('batch_op',)
    elif dense:
        try:
            design.time_kernel('fused_op', reps=1, flush_l2=False)
        except RuntimeError:          # p too wide for the one-pass streaming kernel: the two-pass products
            kernels = ('op',)
    for what in kernels:
        roof[what] = design.time_kernel(what, reps=10, flush_l2=True)
        roof_warm[what] = design.time_kernel(what, reps=10, flush_l2=False)

    # the same kernel on the 12 B/nnz layout (fp64 values + int32 indices) of the same matrix: the format
    # SURVEY section 8d's byte formulas are written for; measured here so both fractions are on record
    valued = {}
    if world == 1 and not dense and os.environ.get('BENCH_VALUED', '1') == '1':
        Dv = GpuSparseDesignMatrix(X, center_predictor=True, add_intercept=True, ctx=ctx, pattern_only=False)
        for what in ('spmv_dot', 'spmv_tdot'):
            valued[what] = Dv.time_kernel(what, reps=10, flush_l2=True)
        del Dv

    sampler = ClockSampler(local_rank)
    if args.clocks == 'none':
        sampler.nvml = None
    sampler.start()
    prof = None
    if args.profile_host:
        import cProfile
        prof = cProfile.Profile()
    barrier()
    ctx.reset_device_ms()
    ctx.reset_launch_count()
    ncu_range = os.environ.get('BENCH_NCU_RANGE') == '1'      # ncu --profile-from-start off
    if ncu_range:
        torch.cuda.profiler.start()
    t0 = time.perf_counter()
    if prof is not None:
        prof.enable()
    if batch is None:
        samples, info2 = bridge.gibbs_resume(info, args.steps)
    else:
        # the batched sampler has no resume: one call runs W untimed + K timed lock-step iterations; the timed region
        # (device-time accumulator, launch counter, wall clock) is opened by the callback at iteration W + 1
        W = max(args.warmup, 1)
        opened = {}

        def open_timed_region(it):
            if it == W + 1:
                barrier()
                ctx.reset_device_ms()
                ctx.reset_launch_count()
                opened['t0'] = time.perf_counter()
        samples, info2 = batch.gibbs(W + args.steps, 0, seeds=list(range(n_chains)), on_iteration=open_timed_region,
                                     params_to_save=('coef', 'global_scale', 'logp'))
        t0 = opened['t0']
        info = info2
    if prof is not None:
        prof.disable()
        import pstats
        pstats.Stats(prof, stream=sys.stderr).sort_stats('cumulative').print_stats(18)
    ctx.sync()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    if ncu_range:
        torch.cuda.profiler.stop()
    dev_ms = ctx.device_ms()
    launches = ctx.launch_count()
    barrier()
    sampler.stop_flag = True
    sampler.join(timeout=2)

    stats = torch.tensor([wall, dev_ms, float(nnz_local), float(launches)], dtype=torch.float64, device='cuda')
    if world > 1:
        mx = stats.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = stats.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        wall, dev_ms, nnz_total = float(mx[0]), float(mx[1]), float(sm[2])
    else:
        nnz_total = float(nnz_local)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    K = args.steps
    resident_state = os.environ.get('BB_RESIDENT_STATE', '1') != '0' and args.sampler == 'cg'
    if batch is not None:
        n_cg = info2['n_cg_iter'][-K:]
    else:
        n_cg = info2['_reg_coef_sampling_info']['n_cg_iter'] if args.sampler == 'cg' else np.array([float('nan')])
    # batched chains: a step advances every chain by one Gibbs iteration -> chain-iterations per second
    value = n_chains * K / (dev_ms / 1000.0)
    e2e = n_chains * K / wall
    # algorithmic bytes of one launch of the dominant kernel (per rank).  Sparse, pattern-only format:
    # 4 B per nnz index + pointers + gathered and written vectors.  Dense: the one-pass fused operator reads X once.
    n_loc = design.shape[0]
    if batch is not None:
        dom = 'batch_op'
        alg = {dom: 2 * 8 * n_loc * p + 8 * 16 * (3 * n_loc + 2 * P)}       # two passes over X + the [n][16] / [p][16] operands
        kernel_name = 'k_batch_dot<1> + k_batch_tdot (X V and X\'(Omega o U) for 16 chains, mma.sync m8n8k4 f64; X read twice)'
    elif dense:
        dom = kernels[0]
        passes = 1 if dom == 'fused_op' else 2
        alg = {dom: passes * 8 * n_loc * p + 8 * (2 * n_loc + 2 * P)}
        kernel_name = 'k_dense_stream<FUSED> (one pass over X)' if dom == 'fused_op' else 'k_dense_stream<DOT_W> + <TDOT> (two passes)'
    else:
        alg = {'spmv_dot': 4 * nnz_local + 4 * (n_loc + 1) + 8 * P + 8 * n_loc,
               'spmv_tdot': 4 * nnz_local + 4 * (P + 1) + 8 * n_loc + 8 * P}
        dom = 'spmv_tdot' if roof['spmv_tdot'] >= roof['spmv_dot'] else 'spmv_dot'
        variant = ctx.get_option('spmv_variant')
        kernel_name = ('k_sell_spmv<binary> (%s)' if variant == 1 else 'k_seg_spmv (%s)') % dom
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:
        pass
    peak = float(peaks.get('hbm_gbs', 6650.0))
    peak_src = 'measured (MEASURED_PEAKS.json hbm_gbs)' if 'hbm_gbs' in peaks else 'fallback 6650 GB/s'
    ach = alg[dom] / (roof[dom] * 1e-3) / 1e9
    # DRAM traffic of the same kernel: from the `ncu --set full` capture of THIS kernel version committed under
    # profiles/ (ncu cannot run inside the timed bench); matched by kernel, workload and rank count, else null
    traffic, traffic_src = None, None
    try:
        tr = json.load(open(os.path.join(ROOT, 'profiles', 'r02_traffic.json')))
        ent = tr.get('%s/%s/n%d' % (args.workload, dom, world))
        if ent and ent.get('kernel_version') == kernel_version(tuple(ent.get('kernel_sources', ('bb_sell.cu', 'bb_dense.cu', 'bb_batch.cu')))):
            traffic, traffic_src = float(ent['dram_bytes']), ent['source']
    except Exception:
        traffic, traffic_src = None, None
    roofline = None
    if args.sampler == 'cholesky':
        # the dominant kernel of the direct sampler is X'WX on the fp64 tensor cores: flops of the lower-triangle tiles
        from bayesbridge_b200.reg_coef_sampler import generate_gaussian_with_weight
        st_ = info2['_markov_chain_state']
        om = np.full(n_loc, float(st_['obs_prec'])) if np.ndim(st_['obs_prec']) == 0 else np.asarray(st_['obs_prec'], dtype=float)
        _, cst = generate_gaussian_with_weight(design, om, np.ones(P), np.zeros(P), return_stats=True)
        nt = (p + 127) // 128
        flops = 2.0 * n_loc * 128 * 128 * nt * (nt + 1) / 2
        peak_t = ctx.measure_fp64_mma_tflops()
        roofline = {'bound': 'tensor', 'kernel': 'k_fisher_syrk (X\'WX, mma.sync m8n8k4 f64; tcgen05 has no fp64 kind)',
                    'achieved': flops / (cst['fisher_ms'] * 1e-3) / 1e12, 'peak': peak_t, 'unit': 'TFLOP/s',
                    'frac': flops / (cst['fisher_ms'] * 1e-3) / 1e12 / peak_t, 'traffic': None,
                    'peak_source': 'measured here: mma.sync m8n8k4 f64 back to back from registers (bb_measure_fp64_mma)',
                    'algorithmic_flops_per_launch': flops, 'ms_per_launch': cst['fisher_ms'],
                    'other': {'factorisation_ms (cuSOLVER potrf)': cst['factorisation_ms'], 'hbm_kernel': kernel_name,
                              'hbm_frac': ach / peak}}
    line = {
        'metric': 'gibbs_iters_per_sec', 'value': value, 'unit': 'iter/s', 'n_gpus': world, 'steps': K,
        'warmup': args.warmup, 'ms_per_step': dev_ms / K, 'higher_is_better': True, 'scaling': 'strong',
        'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic', 'config': dict(config, nnz=int(nnz_total)),
        'cg_iteration': {'fused_p_side_kernel': bool(ctx.get_option('cg_fused')),
                         'exchange': ('none (one rank)' if world == 1 else
                                      ('two-shot peer-memory all-reduce inside the fused kernel' if ctx.get_option('cg_fused') and getattr(ctx, 'p2p_capacity', 0) > 0
                                       else 'ncclAllReduce'))},
        'e2e': {'value': e2e, 'unit': 'iter/s', 'ms_per_step_wall': 1000 * wall / K,
                # device-resident P-side state: per step only the coefficient draw (P doubles) and a few scalars
                # come back; nothing P-length goes up (BB_RESIDENT_STATE=0: 4P+(P-1) doubles up, 2P-1 down)
                'h2d_bytes_per_step': (int(8 * n_chains * 4 * P + 32 * n_chains) if batch is not None else
                                       (64 if resident_state else int(8 * (4 * P + (P - 1))))),
                'd2h_bytes_per_step': (int(8 * n_chains * P + 16 * n_chains) if batch is not None else
                                       (int(8 * P + 128) if resident_state else int(8 * (P + (P - 1)) + 64))),
                'api': ('BatchedBayesBridge.gibbs (public API; %d chains in lock-step; value counts chain-iterations)' % n_chains
                        if batch is not None else 'BayesBridge.gibbs_resume (public API; host numpy state in, samples out)')},
        'gpu_launches': int(launches),
        # one-off costs outside the timed region: design construction (host CSR preparation, upload, CSC + kernel formats)
        # and the chain initialisation (L-BFGS mode search with the likelihood evaluated on the device)
        'setup': {'design_build_s': t_build, 'design_build_split_s': getattr(design, 'build_seconds', None),
                  'chain_init_s': info.get('init_runtime'),
                  'chain_init_optim': {k: (int(v) if isinstance(v, (int, np.integer)) else v)
                                       for k, v in (info.get('_init_optim_info') or {}).items()
                                       if k in ('n_iter', 'n_logp_eval', 'n_grad_eval', 'n_design_matvec', 'is_success')}},
        'mean_n_cg_iter': (float(np.mean(n_cg)) if args.sampler == 'cg' else None),
        'clocks': sampler.summary(),
        'roofline': roofline if roofline is not None else {
            'bound': 'hbm', 'kernel': kernel_name, 'achieved': ach, 'peak': peak, 'unit': 'GB/s',
            'frac': ach / peak, 'traffic': traffic, 'traffic_source': traffic_src, 'peak_source': peak_src,
            'frac_of_8000GBs_spec': ach / 8000.0,          # the north star's nominal figure (SURVEY section 8d asks for both)
            'algorithmic_bytes_per_launch': int(alg[dom]),
            'ms_per_launch': roof[dom], 'timing': 'CUDA events on the library stream, L2 flushed before every launch (cold)',
            'other': dict([(k + '_ms', v) for k, v in roof.items()] + [(k + '_GBs', alg[k] / (v * 1e-3) / 1e9) for k, v in roof.items()]
                          + [(k + '_ms_warm', v) for k, v in roof_warm.items()]
                          + [(k + '_frac_warm', alg[k] / (v * 1e-3) / 1e9 / peak) for k, v in roof_warm.items()]),
            'valued_format_12B_per_nnz': {
                what: {'ms': ms, 'GBs': (alg[what] + 8 * nnz_local) / (ms * 1e-3) / 1e9,
                       'frac': (alg[what] + 8 * nnz_local) / (ms * 1e-3) / 1e9 / peak}
                for what, ms in valued.items()},
        },
    }
    if world == 1 and not args.no_cpu_baseline:
        try:
            # bounded sample of the same workload: the reference's sampler on the FULL matrix for a few iterations,
            # started from the state this chain has reached (so that it solves equally hard CG problems)
            st = info2['_markov_chain_state']
            if batch is not None:           # the reference runs the chains one after the other: chain 0 stands for all
                st = {k: st[k][0] for k in ('coef', 'obs_prec', 'local_scale', 'global_scale')}
            init_state = {k: np.array(st[k], copy=True) if np.ndim(st[k]) else float(st[k])
                          for k in ('coef', 'obs_prec', 'local_scale', 'global_scale')}
            v, desc = reference_run(args.workload, args.cpu_baseline_steps, 0, args.ref_blocks or None,
                                    init_state=init_state if not args.ref_blocks else None, data=None if args.ref_blocks else (X, y),
                                    sampler=args.sampler)
            line['cpu_baseline'] = dict(desc, value=v, unit='iter/s')
        except Exception as e:      # the baseline is reported, never required
            line['cpu_baseline'] = {'value': None, 'unit': 'iter/s', 'cores': 1, 'kind': 'reference',
                                    'sample': 'failed: %r' % (e,)}
    if batch is not None:
        flops = 2 * (2.0 * n_loc * p * 16)
        peak_t = ctx.measure_fp64_mma_tflops()
        line['roofline']['other'].update({
            'tensor_TFLOPs': flops / (roof[dom] * 1e-3) / 1e12, 'tensor_peak_TFLOPs (measured, mma.sync m8n8k4 f64)': peak_t,
            'tensor_frac': flops / (roof[dom] * 1e-3) / 1e12 / peak_t,
            'one_pass_equivalent_hbm_frac (8 n p bytes)': (8.0 * n_loc * p) / (roof[dom] * 1e-3) / 1e9 / peak})
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
