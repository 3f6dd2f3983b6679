"""GPU: the row-sharded path, launched as one process per rank under torchrun.
scripts/multi_gpu_check.py compares, on every rank, the sharded design / CG sampler / chain against an unsharded design
on the same GPU, and checks that all ranks hold bit-identical coefficients.

The set of tests is decided by the box (never skipped):
* always: two ranks on device 0, exchanging through the library's own peer-memory all-reduce (CUDA IPC between the two
  processes; the fused publish / consume kernels of the CG loop) -- NCCL refuses two ranks on one device;
* with >= 2 GPUs visible: additionally two ranks on two devices, once per exchange implementation
  (ncclAllReduce on the library stream, and the peer-memory exchange over NVLink)."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _n_gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _run(port, timeout, **env):
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2',
           '--master-addr', '127.0.0.1', '--master-port', str(port), os.path.join(ROOT, 'scripts', 'multi_gpu_check.py')]
    # the small-problem SpMV variant is switched off so that the production kernels run on the check's small matrices
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, cwd=ROOT,
                         env=dict(os.environ, BB_OPT_ROWWISE_MAX_NNZ='0', **env))
    print(out.stdout[-4000:])
    assert 'MULTI_GPU_CHECK PASS' in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]


def test_sharded_path_matches_unsharded_two_ranks_one_gpu():
    _run(29535, 900, BB_SAME_DEVICE='1', BB_ALLREDUCE='p2p')


if _n_gpus() >= 2:
    @pytest.mark.parametrize('exchange', ['nccl', 'p2p'])
    def test_sharded_path_matches_unsharded_two_gpus(exchange):
        """exchange = 'nccl': ncclAllReduce on the library stream; 'p2p': the fused peer-memory exchange (bb_p2p.cu)."""
        _run(29533 if exchange == 'nccl' else 29534, 600, BB_ALLREDUCE=exchange)
