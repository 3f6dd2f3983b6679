"""CPU: the executable specification of the next SpMV kernel design (experimental/emulate_spmv_v7.py) against a plain
segmented sum.  The emulator mirrors the planned warp algorithm lane by lane; this keeps its index logic (head flags,
piece ordinals, compact outputs, empty segments, partial tiles, segments straddling tiles and slabs) pinned while the
CUDA translation is developed.  Not a product path."""
import numpy as np
import scipy.sparse as sp
import pytest

from experimental import emulate_spmv_v7 as em


def _check(X, W, seed=0):
    rng = np.random.default_rng(seed)
    X = X.tocsr()
    x, w = rng.standard_normal(X.shape[1]), rng.standard_normal(X.shape[0])
    f = em.build_format(X.indptr, X.indices, X.data, X.shape[1], W)
    assert np.allclose(em.spmv(f, x), X @ x, rtol=1e-12, atol=1e-12)
    C = X.tocsc()
    ft = em.build_format(C.indptr, C.indices, C.data, X.shape[0], W)
    assert np.allclose(em.spmv(ft, w), X.T @ w, rtol=1e-12, atol=1e-12)
    return f


@pytest.mark.parametrize('n,p,density,W', [(300, 40, 0.2, 32), (2000, 700, 0.02, 128), (50, 3000, 0.3, 1024),
                                            (4000, 90, 0.05, 64), (10, 5, 0.9, 32), (3000, 2000, 0.0005, 256)])
def test_random_matrices(n, p, density, W):
    X = sp.random(n, p, density=density, format='csr', random_state=np.random.RandomState(n + p), dtype=np.float64)
    if n > 100:      # a hot column: segments that straddle many tiles in the transposed product
        rows = np.random.RandomState(1).choice(n, n // 2, replace=False)
        X = X + sp.csr_matrix((np.ones(n // 2), (rows, np.full(n // 2, 2))), shape=(n, p))
    _check(X, W)


def test_edge_cases():
    # every segment has one entry: 16 heads in every lane, 512 pieces per tile
    _check(sp.identity(1500, format='csr'), 4096)
    # one row holds everything: a single segment across many tiles and several slabs
    X = sp.csr_matrix((np.arange(1., 3001.), (np.zeros(3000, int), np.arange(3000))), shape=(4, 3000))
    _check(X, 1024)
    # empty matrix, empty rows / columns, a slab without entries
    _check(sp.csr_matrix((7, 9)), 4)
    X = sp.csr_matrix((np.ones(3), ([0, 5, 5], [0, 0, 8])), shape=(6, 9))
    f = _check(X, 2)
    assert f['n_compact'] == 3 and f['V'] == 5 * 6
    # tile boundary exactly at a segment start, partial last tile
    X = sp.csr_matrix(np.ones((4, 256)))
    f = _check(X, 256)
    assert list(f['chead']) == [-1, -1] and f['n_heads'] == [2, 2]


def test_every_compact_entry_is_written_exactly_once():
    X = sp.random(700, 300, density=0.03, format='csr', random_state=np.random.RandomState(3), dtype=np.float64)
    f = em.build_format(X.indptr, X.indices, X.data, 300, 64)
    writes = np.zeros(f['n_compact'], int)

    class Counter(np.ndarray):
        def __setitem__(self, k, v):
            writes[k] += 1
            np.ndarray.__setitem__(self, k, v)
    cpart = np.zeros(f['n_compact']).view(Counter)
    head_part = np.zeros(len(f['tiles']))
    x = np.ones(300)
    for t in range(len(f['tiles'])):
        em.warp_tile(f, t, x, cpart, head_part)
    assert np.all(writes == 1)


def test_cuda_translation_host_mode(tmp_path):
    """experimental/spmv_v7.cu must compile for sm_100a, and its host format builder + the CPU transliteration of its
    kernel (`--host`) must reproduce a plain segmented sum (several slab widths, pattern-only and valued)."""
    import os
    import shutil
    import subprocess
    nvcc = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    if not os.path.exists(nvcc):
        pytest.skip('nvcc not available')
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / 'spmv_v7_test')
    subprocess.run([nvcc, '-gencode', 'arch=compute_100a,code=sm_100a', '-O1', '-o', exe,
                    os.path.join(root, 'experimental', 'spmv_v7.cu')], check=True, timeout=600)
    for args in (['4000', '900', '0.02', '1', '1', '--host'], ['4000', '900', '0.02', '0', '1', '--host', '64'],
                 ['300', '20000', '0.001', '1', '0', '--host', '1024'], ['2000', '40', '0.5', '0', '1', '--host', '32']):
        out = subprocess.run([exe] + args, capture_output=True, text=True, timeout=300)
        assert out.returncode == 0 and 'PASS' in out.stdout, out.stdout + out.stderr
