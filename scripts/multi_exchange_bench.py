"""Latency of the per-CG-iteration exchange (all-reduce of p+1 doubles) under torchrun: NCCL vs the peer-memory kernels."""
import os, sys, warnings
import numpy as np, scipy.sparse as sp
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.simplefilter('ignore')
import torch, torch.distributed as dist
from bayesbridge_b200 import _lib
from bayesbridge_b200.design_matrix import GpuSparseDesignMatrix
rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dist.init_process_group('nccl', device_id=torch.device('cuda', local))
ctx = _lib.Context(local); ctx.init_comm_from_torch()
for p in (100000,):
    rs = np.random.RandomState(rank)
    X = sp.random(2000, p, density=0.002, format='csr', random_state=rs, dtype=np.float64); X.data[:] = 1.0
    X = sp.vstack([X, sp.csr_matrix(np.ones((1, p)))]).tocsr()       # keep every column non-constant globally
    D = GpuSparseDesignMatrix(X, center_predictor=True, add_intercept=True, ctx=ctx, presharded=True,
                              n_global=2001 * world, row_offset=2001 * rank)
    for mode, variant in ((0, 0), (1, 0), (1, 1), (1, 2), (1, 4), (1, 5), (1, 7), (1, 15), (1, 0), (0, 0)):
        ctx.set_option('allreduce_p2p', mode)
        ctx.set_option('p2p_variant', variant)
        dist.barrier(); torch.cuda.synchronize()
        us = D.time_kernel('exchange', reps=300, flush_l2=False) * 1e3
        t = torch.tensor([us], device='cuda'); dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            print(f"N={world} p={D.shape[1]} exchange {'p2p variant %2d' % variant if mode else 'nccl          '}: {t.item():.1f} us", flush=True)
print(f"[rank {rank}] p2p status {ctx.p2p_status()}", flush=True)
dist.destroy_process_group()
