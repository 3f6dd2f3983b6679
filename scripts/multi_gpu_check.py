"""Row-sharded path check, run under torchrun (N >= 2 GPUs):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/multi_gpu_check.py
Every rank builds the same small problem; the sharded design (this rank's row block + NCCL allreduce inside
libbbgpu) is compared on every rank against an unsharded design living on the same GPU.
BB_SAME_DEVICE=1 puts every rank on device 0 (single-GPU boxes): NCCL refuses that, so the ranks talk through the
library's own peer-memory exchange (bb_comm_init_local, CUDA IPC between two processes of one device) and torch's
gloo backend only carries the handles."""
import os, sys, warnings
import numpy as np
import scipy.sparse as sp
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.simplefilter('ignore')
import torch, torch.distributed as dist
import bayesbridge_b200 as bb
from bayesbridge_b200 import _lib
from bayesbridge_b200.design_matrix import GpuSparseDesignMatrix, GpuDenseDesignMatrix
from bayesbridge_b200.reg_coef_sampler import ConjugateGradientSampler

rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
same_device = os.environ.get('BB_SAME_DEVICE') == '1'
if same_device:
    local = 0
torch.cuda.set_device(local)
if same_device:
    dist.init_process_group('gloo')
else:
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
tdev = 'cpu' if same_device else 'cuda'
ctx = _lib.Context(local); ctx.init_comm_from_torch()
solo = _lib.Context(local)                      # same GPU, no communicator: the unsharded comparator
ok = True

def rel(a, b): return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))
def check(name, err, tol):
    global ok
    good = err <= tol
    ok &= good
    print(f"[rank {rank}] {name}: {err:.2e} {'ok' if good else 'FAIL'}", flush=True)

rs = np.random.RandomState(0)
n, p = 30001, 1200
X = sp.random(n, p, density=0.01, format='csr', random_state=rs, dtype=np.float64); X.data[:] = 1.0
Xd = rs.randn(2003, 150)
for name, Xm, Cls in (('sparse', X, GpuSparseDesignMatrix), ('dense', Xd, GpuDenseDesignMatrix)):
    nn, pp = Xm.shape
    Ds = Cls(Xm.copy(), center_predictor=True, add_intercept=True, ctx=ctx)       # sharded
    Du = Cls(Xm.copy(), center_predictor=True, add_intercept=True, ctx=solo)      # unsharded
    lo, hi = Ds.row_offset, Ds.row_offset + Ds.shape[0]
    rng = np.random.default_rng(1)
    v, w, wt = rng.standard_normal(pp + 1), rng.standard_normal(nn), rng.random(nn)
    check(name + ' dot', rel(Ds.dot(v), Du.dot(v)[lo:hi]), 1e-13)
    check(name + ' Tdot', rel(Ds.Tdot(w[lo:hi]), Du.Tdot(w)), 1e-12)
    check(name + ' fisher', rel(Ds.compute_fisher_info(wt[lo:hi], True), Du.compute_fisher_info(wt, True)), 1e-12)
    P = pp + 1
    omega = rng.random(nn) * 0.25 + 0.01
    pps = np.concatenate(([0.5], 1 / (0.1 * rng.random(pp) + 1e-3)))
    z, x0, sd = rng.standard_normal(P), 0.01 * rng.standard_normal(P), 0.5 + rng.random(P)
    S = ConjugateGradientSampler(1)
    for atol, tol in ((1e-5 * np.sqrt(P), 1e-5), (1e-12 * np.sqrt(P), 1e-8)):
        np.random.seed(5); e1 = np.random.randn(nn); e2 = np.random.randn(P)
        # inject the same global noise: the sharded call sees only its block of eps1
        import ctypes
        def run(D, om, e1_):
            coef = np.empty(P); ni, info = ctypes.c_int(), ctypes.c_int()
            s = S.choose_preconditioner(pps, None, D, 'prior', sd)
            _lib.check(_lib.load().bb_cg_sample(D._mat, _lib.dptr(om), _lib.dptr(pps), _lib.dptr(z), _lib.dptr(x0), _lib.dptr(s),
                                                float(atol), 500, 0, _lib.dptr(e1_), _lib.dptr(e2), 0, 0, _lib.dptr(coef),
                                                ctypes.byref(ni), ctypes.byref(info), None))
            return coef, ni.value
        cs, ns = run(Ds, np.ascontiguousarray(omega[lo:hi]), np.ascontiguousarray(e1[lo:hi]))
        cu, nu = run(Du, omega, e1)
        check(f'{name} cg atol={atol:.0e} (n_iter {ns}/{nu})', rel(cs, cu), tol)
        # every rank must hold bit-identical coefficients
        t = torch.from_numpy(cs.copy()).to(tdev); t0 = t.clone(); dist.broadcast(t0, 0)
        check(f'{name} cg replicas identical', float((t - t0).abs().max()), 0.0)
    # device noise: sharding-invariant streams -> same draw as unsharded
    a, _ = S.sample(Ds, np.ascontiguousarray(omega[lo:hi]), pps, z, x0, 'prior', sd, maxiter=500, atol=1e-11, noise='device', philox=(9, 3))
    b, _ = S.sample(Du, omega, pps, z, x0, 'prior', sd, maxiter=500, atol=1e-11, noise='device', philox=(9, 3))
    check(name + ' cg device-noise sharded vs unsharded', rel(a, b), 1e-8)

# batched multi-chain CG draw (bb_batch.cu), sharded vs unsharded: C chains in lock-step, exchange of C (p+1) doubles
nn, pp = Xd.shape
Ds = GpuDenseDesignMatrix(Xd.copy(), center_predictor=True, add_intercept=True, ctx=ctx)
Du = GpuDenseDesignMatrix(Xd.copy(), center_predictor=True, add_intercept=True, ctx=solo)
lo, hi = Ds.row_offset, Ds.row_offset + Ds.shape[0]
Cb, P = 3, pp + 1
rng = np.random.default_rng(7)
om_b = rng.random((Cb, nn)) * 0.25 + 0.01
pps_b = np.concatenate((np.full((Cb, 1), 0.5), 1 / (0.1 * rng.random((Cb, pp)) + 1e-3)), axis=1)
z_b, x0_b = rng.standard_normal((Cb, P)), 0.01 * rng.standard_normal((Cb, P))
s_b = np.array([ConjugateGradientSampler(1).choose_preconditioner(pps_b[c], None, Du, 'prior', np.ones(P)) for c in range(Cb)])
e1_b, e2_b = rng.standard_normal((Cb, nn)), rng.standard_normal((Cb, P))
import ctypes
def run_batched(D, om, e1_):
    lib = _lib.load()
    _lib.check(lib.bb_batch_init(D._mat, Cb))
    coef = np.empty((Cb, P)); ni, info = (ctypes.c_int * Cb)(), (ctypes.c_int * Cb)()
    _lib.check(lib.bb_cg_sample_batched(D._mat, _lib.dptr(np.ascontiguousarray(om)), _lib.dptr(pps_b), _lib.dptr(z_b), _lib.dptr(x0_b),
                                        _lib.dptr(s_b), 1e-11 * np.sqrt(P), 500, 0, _lib.dptr(np.ascontiguousarray(e1_)), _lib.dptr(e2_b),
                                        None, None, _lib.dptr(coef), ni, info))
    return coef, list(ni)
cs, ns = run_batched(Ds, om_b[:, lo:hi], e1_b[:, lo:hi])
cu, nu = run_batched(Du, om_b, e1_b)
check(f'batched cg (n_iter {ns}/{nu}) sharded vs unsharded', rel(cs, cu), 1e-8)
t = torch.from_numpy(cs.copy()).to(tdev); t0 = t.clone(); dist.broadcast(t0, 0)
check('batched cg replicas identical', float((t - t0).abs().max()), 0.0)

# log-likelihood and gradient on the device (chain initialisation): sums over the shards
y_ll = rs.binomial(1, 0.3, n).astype(float)
beta_ll = np.random.default_rng(3).standard_normal(p + 1) * 0.1
m_s = bb.RegressionModel(y_ll, X, family='logit', ctx=ctx)
m_u = bb.RegressionModel(y_ll, X, family='logit', ctx=solo)
ll_s, g_s = m_s.compute_loglik_and_gradient(beta_ll)
ll_u, g_u = m_u.compute_loglik_and_gradient(beta_ll)
check('loglik sharded vs unsharded', abs(ll_s - ll_u) / abs(ll_u), 1e-12)
check('gradient sharded vs unsharded', rel(g_s, g_u), 1e-12)

# full chain, sharded vs unsharded, device RNG: identical streams => same chain up to CG tolerance
y = rs.binomial(1, 1 / (1 + np.exp(-(X @ np.concatenate((np.full(5, 1.5), np.zeros(p - 5))) - 1.0))))
chains = []
for c in (ctx, solo):
    model = bb.RegressionModel(y, X, family='logit', ctx=c)
    s, info = bb.BayesBridge(model, bb.RegressionCoefPrior(bridge_exponent=.5)).gibbs(30, 10, coef_sampler_type='cg', seed=2)
    chains.append(s)
check('chain coef mean sharded vs unsharded', float(np.abs(chains[0]['coef'].mean(1) - chains[1]['coef'].mean(1)).max()), 5e-2)
check('chain logp sharded vs unsharded', float(abs(chains[0]['logp'].mean() / chains[1]['logp'].mean() - 1)), 1e-2)
ready, err = ctx.p2p_status()
print(f'[rank {rank}] p2p ready={ready} error={err} (BB_ALLREDUCE={os.environ.get("BB_ALLREDUCE", "nccl")})', flush=True)
ok &= (err == 0) and ready       # the peer-memory buffers are always attached (fused CG iteration)
flag = torch.tensor([1.0 if ok else 0.0]).to(tdev); dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print('MULTI_GPU_CHECK', 'PASS' if flag.item() == 1.0 else 'FAIL', flush=True)
dist.destroy_process_group()
