"""CPU: host-side pieces of the product that run without a GPU -- prior, options, bookkeeping, the
summarizer (vs reference outputs), row sharding, and the sharded-operator algebra over gloo (world size 2)."""
import os
import sys
import numpy as np
import pytest

from conftest import golden, ROOT


def test_summarizer_matches_reference_outputs():
    from bayesbridge_b200.reg_coef_sampler.reg_coef_posterior_summarizer import RegressionCoeffficientPosteriorSummarizer
    g = golden('summarizer_ref.npz')
    S = RegressionCoeffficientPosteriorSummarizer(12, 2, 1.5)
    for it in range(len(g['gscale'])):
        gs, ls = float(g['gscale'][it]), g['lscale'][it]
        assert np.array_equal(S.extrapolate_coef_condmean(gs, ls), g['x0'][it])
        assert np.array_equal(S.estimate_coef_precond_scale_sd(), g['sd'][it])
        S.update(g['coef'][it], gs, ls)


def test_preconditioner_choice():
    from bayesbridge_b200.reg_coef_sampler import ConjugateGradientSampler
    from oracle.cg_oracle import precond_scale_prior
    pps = np.array([0.0, 0.5, 4.0, 10.0])
    sd = np.array([0.3, 0.6, 9., 9.])
    s = ConjugateGradientSampler(2).choose_preconditioner(pps, None, None, 'prior', sd)
    assert np.array_equal(s, precond_scale_prior(pps, 2, sd))
    assert np.array_equal(s, [0.6, 1.2, 0.25, 0.1])


def test_prior_hyperparameters_and_scale_parametrisation():
    from bayesbridge_b200 import RegressionCoefPrior
    prior = RegressionCoefPrior(bridge_exponent=.25, global_scale_prior_hyper_param={'log10_mean': -4., 'log10_sd': 1.})
    hyper = prior.param['gscale_neg_power']
    # moments of log(phi) under Gamma(shape, rate) reproduce the requested log10 mean / sd of the global scale
    from scipy.special import polygamma
    import math
    sd_log_phi = math.sqrt(polygamma(1, hyper['shape']))
    assert sd_log_phi / .25 == pytest.approx(math.log(10.), rel=1e-6)
    g, l = prior.adjust_scale(0.1, np.ones(3), to='raw')
    g2, l2 = prior.adjust_scale(g, l, to='coef_magnitude')
    assert g2 == pytest.approx(0.1) and np.allclose(l2, 1.0)
    clone = prior.clone(bridge_exponent=.5)
    assert clone.bridge_exp == .5 and clone.param['gscale'] == prior.param['gscale']
    with pytest.raises(ValueError):
        RegressionCoefPrior(bridge_exponent=3.)
    assert RegressionCoefPrior.compute_power_exp_ave_magnitude(1.) == pytest.approx(1.0)


def test_sampler_options_gate_device_matrices():
    from bayesbridge_b200.gibbs_util import SamplerOptions

    class FakeDesign:
        use_gpu, use_cupy, is_sparse, shape, n_global, nnz = True, False, True, (10, 3), 10, 5

    opt = SamplerOptions.pick_default_and_create(None, None, 'logit', FakeDesign())
    assert opt.coef_sampler_type == 'cg' and opt.noise == 'device'
    assert SamplerOptions.pick_default_and_create('cholesky', None, 'logit', FakeDesign()).coef_sampler_type == 'cholesky'
    with pytest.raises(ValueError):
        SamplerOptions.pick_default_and_create('hmc', None, 'logit', FakeDesign())
    with pytest.raises(ValueError):
        SamplerOptions.pick_default_and_create('newton', None, 'logit', FakeDesign())
    opt = SamplerOptions.pick_default_and_create(None, {'noise': 'host'}, 'linear', FakeDesign())
    assert opt.get_info()['noise'] == 'host'


def test_chain_manager_storage_and_merge():
    from bayesbridge_b200.gibbs_util import MarkovChainManager
    M = MarkovChainManager(4, 3, 1, 'logit')
    samples, sinfo = {}, {}
    M.pre_allocate(samples, sinfo, 6, 2, ('coef', 'global_scale', 'logp', 'obs_prec', 'local_scale'), 'cg')
    assert samples['coef'].shape == (3, 3) and samples['obs_prec'].shape == (4, 3) and sinfo['n_cg_iter'].shape == (3,)
    for it in range(1, 9):
        M.store_current_state(samples, it, 2, 2, np.full(3, it), np.full(2, it), float(it), lambda: np.full(4, it), -it,
                              ('coef', 'global_scale', 'logp', 'obs_prec', 'local_scale'))
        M.store_sampling_info(sinfo, {'n_cg_iter': it}, it, 2, 2, 'cg')
    assert list(samples['global_scale']) == [4., 6., 8.] and list(sinfo['n_cg_iter']) == [4., 6., 8.]
    assert np.all(samples['obs_prec'][:, 1] == 6)
    merged, info = M.merge_outputs(
        {'coef': np.zeros((3, 2))}, {'_reg_coef_sampling_info': {'n_cg_iter': np.ones(2)}, 'n_iter': 2, 'runtime': 1.,
                                      '_init_optim_info': 'x', 'seed': 5},
        {'coef': np.ones((3, 3))}, {'_reg_coef_sampling_info': {'n_cg_iter': np.zeros(3)}, 'n_iter': 3, 'runtime': 2.})
    assert merged['coef'].shape == (3, 5) and info['n_iter'] == 5 and info['seed'] == 5


def test_row_sharding_covers_all_rows():
    from bayesbridge_b200.design_matrix import AbstractDesignMatrix

    class C:
        pass
    for n in (7, 1000, 1000003):
        for G in (1, 2, 4, 8):
            edges = []
            for r in range(G):
                c = C(); c.nranks, c.rank = G, r
                edges.append(AbstractDesignMatrix.shard_rows(n, c))
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(edges[i][1] == edges[i + 1][0] for i in range(G - 1))
            assert max(hi - lo for lo, hi in edges) - min(hi - lo for lo, hi in edges) <= 1


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    import torch
    import scipy.sparse as sp
    sys.path.insert(0, ROOT)
    from oracle import cg_oracle as co
    from bayesbridge_b200.design_matrix import AbstractDesignMatrix
    os.environ['MASTER_ADDR'], os.environ['MASTER_PORT'] = '127.0.0.1', str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    rs = np.random.RandomState(0)
    n, p = 501, 37
    X = sp.random(n, p, density=0.2, format='csr', random_state=rs, dtype=np.float64)
    omega, v = rs.rand(n) + 0.1, rs.randn(p + 1)
    full = co.DesignOracle(X, True, True)

    class C:
        pass
    c = C(); c.nranks, c.rank = world, rank
    lo, hi = AbstractDesignMatrix.shard_rows(n, c)
    # what one rank of the sharded operator computes: local rows, GLOBAL column means, then
    # traw = [sum w_g ; X_g' w_g] -> allreduce(sum) -> t = [sum w ; X'w - (sum w) c]
    Xg = X[lo:hi]
    cmean = full.c
    u = v[0] + Xg @ v[1:] - cmean @ v[1:]
    w = omega[lo:hi] * u
    traw = torch.from_numpy(np.concatenate(([w.sum()], Xg.T @ w)))
    dist.all_reduce(traw)
    traw = traw.numpy()
    t = np.concatenate(([traw[0]], traw[1:] - traw[0] * cmean))
    ref = full.Tdot(omega * full.dot(v))
    q.put((rank, float(np.linalg.norm(t - ref) / np.linalg.norm(ref))))
    dist.destroy_process_group()


def test_sharded_operator_algebra_gloo_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    res = [q.get(timeout=180) for _ in procs]
    for pr in procs:
        pr.join(timeout=60)
    assert sorted(r for r, _ in res) == [0, 1]
    assert all(err < 1e-13 for _, err in res)


def _twoshot_worker(rank, world, port, q):
    """The protocol of the fused P-side kernel's exchange (csrc/bb_pside.cu), with gloo point-to-point messages in place of
    NVLink stores: phase A pushes chunk c of the local vector into rank c's inbox slot, phase B adds the slots of the own
    chunk in RANK ORDER and pushes the sum to everybody.  Also the sharded local-scale gather (pieces summed against zeros)."""
    import torch
    import torch.distributed as dist
    os.environ['MASTER_ADDR'], os.environ['MASTER_PORT'] = '127.0.0.1', str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    L = 1003                                              # p + 1, not a multiple of the number of ranks
    Cw = (((L + world - 1) // world) + 1) & ~1            # chunk width of bb_pside.cu
    vecs = [np.random.default_rng(100 + r).standard_normal(L) * 10.0 ** np.random.default_rng(r).integers(-8, 8, L) for r in range(world)]
    mine = vecs[rank]
    inbox = np.zeros((world, Cw))
    reqs = []
    for c in range(world):                                # phase A
        lo = c * Cw
        chunk = np.zeros(Cw)
        seg = mine[lo:min(lo + Cw, L)]
        chunk[:len(seg)] = seg
        if c == rank:
            inbox[rank] = chunk
        else:
            reqs.append(dist.isend(torch.from_numpy(chunk.copy()), dst=c, tag=rank))
    for sdr in range(world):
        if sdr != rank:
            buf = torch.zeros(Cw, dtype=torch.float64)
            dist.recv(buf, src=sdr, tag=sdr)
            inbox[sdr] = buf.numpy()
    for r_ in reqs:
        r_.wait()
    acc = np.zeros(Cw)
    for sdr in range(world):                              # phase B: rank order
        acc = acc + inbox[sdr]
    gathered = [torch.zeros(Cw, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(gathered, torch.from_numpy(acc))
    result = np.concatenate([g.numpy() for g in gathered])[:L]
    expect = np.zeros(L)
    for r in range(world):
        expect = expect + vecs[r]                          # the same rank-ordered sum, computed locally
    # sharded local-scale draw: every rank fills its own slice, zeros elsewhere, then a sum (exact: x + 0 == x)
    ns = 777
    lam = np.random.default_rng(5).random(ns)
    lo, hi = ns * rank // world, ns * (rank + 1) // world
    piece = np.zeros(ns)
    piece[lo:hi] = lam[lo:hi]
    t = torch.from_numpy(piece)
    dist.all_reduce(t)
    q.put((rank, bool(np.array_equal(result, expect)), bool(np.array_equal(t.numpy(), lam)), result.tobytes()))
    dist.destroy_process_group()


@pytest.mark.parametrize('world', [2, 3])
def test_two_shot_exchange_protocol_gloo(world):
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 31500 + (os.getpid() + 7 * world) % 2000
    procs = [ctx.Process(target=_twoshot_worker, args=(r, world, port, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    res = [q.get(timeout=180) for _ in procs]
    for pr in procs:
        pr.join(timeout=60)
    assert sorted(r[0] for r in res) == list(range(world))
    assert all(r[1] for r in res), 'the exchanged sum is the rank-ordered sum, bit for bit'
    assert all(r[2] for r in res), 'pieces summed against zeros reproduce the vector exactly'
    assert len({r[3] for r in res}) == 1, 'every rank holds identical bits'


def test_bench_generator_is_block_deterministic():
    """bench.py builds C4 from 50 seeded row blocks so that every N sees the same matrix: a rank's rows must not
    depend on which other blocks the process generated."""
    import bench
    n, p, dens = 5000, 400, 0.01
    full, y_full = bench.generate_rows(range(bench.N_BLOCKS), n, p, dens)
    part, y_part = bench.generate_rows(range(10, 20), n, p, dens)
    lo, hi = n * 10 // bench.N_BLOCKS, n * 20 // bench.N_BLOCKS
    assert (full[lo:hi] != part).nnz == 0 and np.array_equal(y_full[lo:hi], y_part)
    assert full.indices.dtype == np.int32 and np.all(full.data == 1.0)
    assert abs(full.nnz / (n * p) - dens) < 0.5 * dens


def test_reference_arm_line_has_the_contract_keys():
    """`bench.py --impl reference` (the reference's own CPU sampler from oracle/_ref, or the oracle port) prints one JSON
    line with the keys the driver expects; run on the smallest workload so that it takes seconds."""
    import json
    import subprocess
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--workload', 'C1',
                          '--steps', '3', '--warmup', '1', '--ref-blocks', '10'],
                         capture_output=True, text=True, timeout=300, cwd=ROOT)
    line = [l for l in out.stdout.splitlines() if l.startswith('{')][-1]
    d = json.loads(line)
    assert d['impl'] == 'reference' and d['metric'] == 'gibbs_iters_per_sec' and d['value'] > 0
    for key in ('unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling', 'config',
                'cpu_baseline', 'e2e'):
        assert key in d
    assert d['cpu_baseline']['kind'] in ('reference', 'port') and d['e2e']['h2d_bytes_per_step'] == 0


def test_reference_arm_budget_plan():
    """bench.plan_reference_iterations: what the reference arm still runs when its wall-clock budget is short.  It never
    extrapolates and never runs a sub-sample: it runs FEWER full-size iterations and reports the counts it executed."""
    sys.path.insert(0, ROOT)
    import bench
    plan = bench.plan_reference_iterations
    assert plan(23, 5, 20, 2) == (5, 20)            # everything fits (W = 5 of which 2 are done, K = 20)
    assert plan(100, 5, 20, 2) == (5, 20)
    assert plan(22, 5, 20, 2) == (5, 19)            # one short: the warm-up is kept, one timed iteration goes
    assert plan(13, 5, 20, 2) == (5, 10)            # down to half the requested steps the warm-up is still kept
    w, k = plan(10, 5, 20, 2)                       # below that both are cut ...
    assert k == 8 and w == 4 and (w - 2) + k <= 10
    w, k = plan(1, 5, 20, 2)                        # ... but never fewer than two timed iterations
    assert k == 2 and w == 2
    for fit in range(0, 40):
        w, k = plan(fit, 5, 20, 2)
        assert 2 <= k <= 20 and 2 <= w <= 5
        assert (w - 2) + k <= max(fit, 2)           # never plans more than fits (beyond the two-iteration floor)
