import os
import sys
import warnings

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, 'tests', 'golden')
REF_DIR = os.path.join(ROOT, 'oracle', '_ref')


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def record_achieved(test, case, achieved, bound, **extra):
    """Append the achieved error of one parity case to gpurun_out/parity_achieved.jsonl (created on demand; merged
    back from the GPU box by gpurun and copied to profiles/ per round) and print it (visible with pytest -s / -rP)."""
    import json
    row = dict(test=test, case=str(case), achieved=float(achieved), bound=float(bound), **extra)
    print('PARITY', json.dumps(row))
    try:
        d = os.path.join(ROOT, 'gpurun_out')
        os.makedirs(d, exist_ok=True)
        with open(os.path.join(d, 'parity_achieved.jsonl'), 'a') as f:
            f.write(json.dumps(row) + '\n')
    except OSError:
        pass


def golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def import_reference():
    """The reference package built into oracle/_ref by oracle/build_ref.sh, or None."""
    if not os.path.isdir(os.path.join(REF_DIR, 'bayesbridge')):
        return None
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    try:
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            import bayesbridge
        return bayesbridge
    except Exception:
        return None


@pytest.fixture(scope='session')
def ctx():
    """The library's default context.  Problems below 2M nnz would by default take the small-problem SpMV variant
    (sub-warp per segment, bb_sparse.cu); the parity tests are about the production kernels, so the switch is turned off
    here and the small-problem variant is tested explicitly (spmv_variant = 2 cases, test_small_problem_variant_*)."""
    from bayesbridge_b200 import _lib
    c = _lib.Context.default()
    c.set_option('rowwise_max_nnz', 0)
    return c


@pytest.fixture(autouse=True)
def _quiet():
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        yield
