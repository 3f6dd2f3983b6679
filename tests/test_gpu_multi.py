"""GPU (>= 2 devices): the row-sharded path, launched as one process per GPU under torchrun.
scripts/multi_gpu_check.py compares, on every rank, the sharded design / CG sampler / chain (NCCL allreduce inside
libbbgpu) against an unsharded design on the same GPU, and checks that all ranks hold bit-identical coefficients."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('exchange', ['nccl', 'p2p'])
def test_sharded_path_matches_unsharded_two_gpus(exchange):
    """exchange = 'nccl': ncclAllReduce on the library stream; 'p2p': the fused peer-memory exchange (bb_p2p.cu)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2',
           '--master-addr', '127.0.0.1', '--master-port', '29533' if exchange == 'nccl' else '29534',
           os.path.join(ROOT, 'scripts', 'multi_gpu_check.py')]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT,
                         env=dict(os.environ, BB_ALLREDUCE=exchange))
    assert 'MULTI_GPU_CHECK PASS' in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]


def test_sharded_path_matches_unsharded_two_ranks_one_gpu():
    """The same check with both ranks on device 0, so that it also runs on a single-GPU box: the ranks exchange
    through the library's own peer-memory all-reduce (CUDA IPC; fused publish / consume kernels of the CG loop),
    not NCCL (which refuses two ranks on one device)."""
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2',
           '--master-addr', '127.0.0.1', '--master-port', '29535', os.path.join(ROOT, 'scripts', 'multi_gpu_check.py')]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT,
                         env=dict(os.environ, BB_SAME_DEVICE='1', BB_ALLREDUCE='p2p'))
    print(out.stdout[-4000:])
    assert 'MULTI_GPU_CHECK PASS' in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]
