"""ctypes binding of libbbgpu.so (the C-ABI declared in include/bbgpu.h).

There is no CPU fallback: if the shared library is missing or no B200 is visible, every
compute entry point raises.  Importing this module never touches the GPU.
"""
import ctypes
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libbbgpu.so')

c_int, c_i64, c_u64, c_dbl = ctypes.c_int, ctypes.c_int64, ctypes.c_uint64, ctypes.c_double
c_void_p, c_char_p = ctypes.c_void_p, ctypes.c_char_p
P_dbl = ctypes.POINTER(c_dbl)
P_i32 = ctypes.POINTER(ctypes.c_int32)
P_int = ctypes.POINTER(c_int)
P_i64 = ctypes.POINTER(c_i64)

# name -> (restype, argtypes); mirrors include/bbgpu.h one to one
SIGNATURES = {
    'bb_last_error': (c_char_p, []),
    'bb_version': (c_int, []),
    'bb_device_count': (c_int, [P_int]),
    'bb_init': (c_int, [c_int, ctypes.POINTER(c_void_p)]),
    'bb_destroy': (c_int, [c_void_p]),
    'bb_set_option': (c_int, [c_void_p, c_char_p, c_i64]),
    'bb_get_option': (c_int, [c_void_p, c_char_p, P_i64]),
    'bb_get_launch_count': (c_int, [c_void_p, P_i64]),
    'bb_reset_launch_count': (c_int, [c_void_p]),
    'bb_get_device_ms': (c_int, [c_void_p, P_dbl]),
    'bb_reset_device_ms': (c_int, [c_void_p]),
    'bb_sync': (c_int, [c_void_p]),
    'bb_comm_unique_id': (c_int, [c_char_p, ctypes.c_char_p]),
    'bb_comm_init': (c_int, [c_void_p, c_char_p, c_int, c_int, c_char_p]),
    'bb_comm_init_local': (c_int, [c_void_p, c_int, c_int]),
    'bb_comm_allreduce_host': (c_int, [c_void_p, P_dbl, c_i64]),
    'bb_comm_p2p_export': (c_int, [c_void_p, c_i64, ctypes.c_char_p]),
    'bb_comm_p2p_attach': (c_int, [c_void_p, ctypes.c_char_p]),
    'bb_comm_p2p_status': (c_int, [c_void_p, P_int, P_int]),
    'bb_csr_upload': (c_int, [c_void_p, c_i64, c_i64, c_i64, P_i32, P_i32, P_dbl, P_dbl, c_int, c_i64, c_i64,
                              ctypes.POINTER(c_void_p)]),
    'bb_dense_upload': (c_int, [c_void_p, c_i64, c_i64, P_dbl, P_dbl, c_int, c_i64, c_i64, ctypes.POINTER(c_void_p)]),
    'bb_mat_free': (c_int, [c_void_p]),
    'bb_mat_info': (c_int, [c_void_p, P_i64, P_i64, P_i64, P_int, P_int]),
    'bb_mat_export_csr': (c_int, [c_void_p, P_i32, P_i32, P_dbl]),
    'bb_mat_export_csc': (c_int, [c_void_p, P_i32, P_i32, P_dbl]),
    'bb_dot': (c_int, [c_void_p, P_dbl, P_dbl]),
    'bb_tdot': (c_int, [c_void_p, P_dbl, P_dbl]),
    'bb_fisher_diag': (c_int, [c_void_p, P_dbl, P_dbl]),
    'bb_fisher_full': (c_int, [c_void_p, P_dbl, P_dbl, P_dbl]),
    'bb_measure_fp64_mma': (c_int, [c_void_p, P_dbl]),
    'bb_column_moments': (c_int, [c_void_p, P_dbl, P_dbl]),
    'bb_set_column_offset': (c_int, [c_void_p, P_dbl]),
    'bb_batch_init': (c_int, [c_void_p, c_int]),
    'bb_batch_free': (c_int, [c_void_p]),
    'bb_batch_set_obs_prec': (c_int, [c_void_p, P_dbl]),
    'bb_batch_get_obs_prec': (c_int, [c_void_p, P_dbl]),
    'bb_dot_batched': (c_int, [c_void_p, P_dbl, P_dbl]),
    'bb_tdot_batched': (c_int, [c_void_p, P_dbl, P_dbl]),
    'bb_cg_sample_batched': (c_int, [c_void_p, P_dbl, P_dbl, P_dbl, P_dbl, P_dbl, c_dbl, c_int, c_int, P_dbl, P_dbl,
                                     ctypes.POINTER(c_u64), ctypes.POINTER(c_u64), P_dbl, P_int, P_int]),
    'bb_pg_from_coef_batched': (c_int, [c_void_p, P_dbl, ctypes.POINTER(c_u64), ctypes.POINTER(c_u64), P_dbl]),
    'bb_mode_search': (c_int, [c_void_p, P_dbl, P_dbl, P_dbl, c_dbl, c_int, c_dbl, c_dbl, c_int, P_dbl, P_int, P_int, P_int]),
    'bb_loglik_and_gradient': (c_int, [c_void_p, P_dbl, c_dbl, c_int, P_dbl, P_dbl]),
    'bb_cholesky_sample': (c_int, [c_void_p, P_dbl, P_dbl, P_dbl, P_dbl, P_dbl, P_dbl]),
    'bb_set_outcome': (c_int, [c_void_p, P_dbl, P_dbl]),
    'bb_set_obs_prec': (c_int, [c_void_p, P_dbl]),
    'bb_set_obs_prec_scalar': (c_int, [c_void_p, c_dbl]),
    'bb_get_obs_prec': (c_int, [c_void_p, P_dbl]),
    'bb_get_linear_predictor': (c_int, [c_void_p, P_dbl]),
    'bb_cg_sample': (c_int, [c_void_p, P_dbl, P_dbl, P_dbl, P_dbl, P_dbl, c_dbl, c_int, c_int, P_dbl, P_dbl,
                             c_u64, c_u64, P_dbl, P_int, P_int, P_dbl]),
    'bb_pg_sample': (c_int, [c_void_p, c_i64, P_i32, P_dbl, c_u64, c_u64, c_i64, P_dbl]),
    'bb_pg_from_coef': (c_int, [c_void_p, P_dbl, c_u64, c_u64, P_dbl, P_dbl]),
    'bb_linear_rss': (c_int, [c_void_p, P_dbl, P_dbl]),
    'bb_tilted_stable_sample': (c_int, [c_void_p, c_i64, c_dbl, P_dbl, c_u64, c_u64, c_i64, P_dbl]),
    'bb_philox_normal': (c_int, [c_void_p, c_i64, c_int, c_u64, c_u64, c_i64, P_dbl]),
    'bb_state_init': (c_int, [c_void_p, c_int, P_dbl, c_dbl]),
    'bb_state_set': (c_int, [c_void_p, P_dbl, P_dbl, P_dbl, c_i64]),
    'bb_state_get': (c_int, [c_void_p, P_dbl, P_dbl, P_dbl, P_i64]),
    'bb_cg_sample_resident': (c_int, [c_void_p, P_dbl, c_dbl, c_dbl, c_dbl, c_int, c_u64, c_u64, P_dbl, P_int, P_int, P_dbl]),
    'bb_local_scale_resident': (c_int, [c_void_p, c_dbl, c_dbl, c_u64, c_u64, P_int, P_dbl]),
    'bb_time_kernel': (c_int, [c_void_p, c_char_p, c_int, c_int, P_dbl]),
    'bb_spmv_timeline': (c_int, [c_void_p, c_int, c_int, c_void_p, ctypes.c_int64, ctypes.POINTER(c_int)]),
}

BB_NOISE_INJECT, BB_NOISE_PHILOX = 0, 1

_lib = None


def load():
    """Load libbbgpu.so (once).  Raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "libbbgpu.so not found at {}: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "or `make -C bayesbridge_b200/csrc`. There is no CPU fallback.".format(LIB_PATH))
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc):
    if rc != 0:
        msg = load().bb_last_error()
        raise RuntimeError("libbbgpu: " + (msg.decode() if msg else "error {}".format(rc)))


def dptr(a):
    """Pointer to a C-contiguous float64 array (None -> NULL)."""
    if a is None:
        return None
    assert a.dtype == np.float64 and a.flags['C_CONTIGUOUS']
    return a.ctypes.data_as(P_dbl)


def iptr(a):
    if a is None:
        return None
    assert a.dtype == np.int32 and a.flags['C_CONTIGUOUS']
    return a.ctypes.data_as(P_i32)


def as_f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def nccl_library_path():
    """Path of the NCCL shared object bundled with torch (falls back to the system one)."""
    try:
        import nvidia.nccl
        cand = os.path.join(list(nvidia.nccl.__path__)[0], 'lib', 'libnccl.so.2')
        if os.path.exists(cand):
            return cand
    except Exception:
        pass
    return 'libnccl.so.2'


class Context:
    """One GPU + one stream (+ an optional communicator over the ranks of a torchrun job)."""

    _default = None

    def __init__(self, device=None):
        lib = load()
        if device is None:
            device = int(os.environ.get('LOCAL_RANK', '0'))
        handle = c_void_p()
        check(lib.bb_init(int(device), ctypes.byref(handle)))
        self.handle = handle
        self.device = int(device)
        self.nranks, self.rank = 1, 0
        # tuning knobs of the library can be preset from the environment, e.g. BB_OPT_SPMV_STAGE=2
        for key, val in os.environ.items():
            if key.startswith('BB_OPT_'):
                check(lib.bb_set_option(handle, key[len('BB_OPT_'):].lower().encode(), int(val)))

    @classmethod
    def default(cls):
        if cls._default is None:
            cls._default = cls()
        return cls._default

    def set_option(self, name, value):
        check(load().bb_set_option(self.handle, name.encode(), int(value)))

    def get_option(self, name):
        v = c_i64()
        check(load().bb_get_option(self.handle, name.encode(), ctypes.byref(v)))
        return v.value

    def launch_count(self):
        v = c_i64()
        check(load().bb_get_launch_count(self.handle, ctypes.byref(v)))
        return v.value

    def reset_launch_count(self):
        check(load().bb_reset_launch_count(self.handle))

    def device_ms(self):
        """Device milliseconds spent inside library entry points since the last reset (CUDA events)."""
        v = c_dbl()
        check(load().bb_get_device_ms(self.handle, ctypes.byref(v)))
        return v.value

    def reset_device_ms(self):
        check(load().bb_reset_device_ms(self.handle))

    def sync(self):
        check(load().bb_sync(self.handle))

    def measure_fp64_mma_tflops(self):
        """fp64 tensor-core throughput of this GPU (mma.sync m8n8k4 f64 from registers), TFLOP/s."""
        v = c_dbl()
        check(load().bb_measure_fp64_mma(self.handle, ctypes.byref(v)))
        return v.value

    def init_comm_from_torch(self):
        """Attach an NCCL communicator spanning the ranks of the current torch.distributed job.

        torch.distributed is only the plumbing that carries the 128-byte NCCL unique id from rank 0
        to the other ranks; the allreduces themselves are issued by libbbgpu on its own stream."""
        import torch.distributed as dist
        if not dist.is_initialized() or dist.get_world_size() == 1:
            return
        lib = load()
        rank, world = dist.get_rank(), dist.get_world_size()
        # NCCL refuses two ranks on one device; such a job (the single-GPU test of the sharded path) and
        # BB_COMM=local use the library's own peer-memory exchange for everything
        import socket
        where = [None] * world
        dist.all_gather_object(where, (socket.gethostname(), self.device))
        if os.environ.get('BB_COMM', '') == 'local' or len(set(where)) < world:
            check(lib.bb_comm_init_local(self.handle, world, rank))
            self.nranks, self.rank, self.comm_local = world, rank, True
            self.init_p2p(int(os.environ.get('BB_P2P_CAPACITY', 1 << 18)))
            return
        path = nccl_library_path().encode()
        buf = ctypes.create_string_buffer(128)
        if rank == 0:
            check(lib.bb_comm_unique_id(path, buf))
        box = [bytes(buf.raw)]
        dist.broadcast_object_list(box, src=0)
        check(lib.bb_comm_init(self.handle, path, world, rank, box[0]))
        self.nranks, self.rank = world, rank

    def init_p2p(self, capacity):
        """Map every rank's exchange buffer into every other rank (CUDA IPC over NVLink).  Collective: every rank
        must call it.  torch.distributed only carries the 64-byte IPC handles.

        The buffers serve two protocols of libbbgpu: the two-shot all-reduce fused into the P-side kernel of every CG
        iteration (bb_pside.cu; used whenever the buffers are attached and option cg_fused is on), and a one-shot
        all-reduce for the few other exchanges of a Gibbs step (RHS product, log-likelihood).  The latter go through
        ncclAllReduce unless BB_ALLREDUCE=p2p (or the communicator is local, i.e. has no NCCL behind it).
        BB_P2P=0 skips the attachment altogether (everything NCCL, unfused CG iteration)."""
        import torch.distributed as dist
        if self.nranks == 1 or getattr(self, 'p2p_capacity', 0) >= capacity:
            return
        if getattr(self, 'p2p_capacity', 0) > 0:
            return      # already attached with a smaller capacity: larger vectors fall back to NCCL
        local = getattr(self, 'comm_local', False)
        if os.environ.get('BB_P2P', '1') == '0' and not local:
            return
        lib = load()
        buf = ctypes.create_string_buffer(64)
        check(lib.bb_comm_p2p_export(self.handle, int(capacity), buf))
        handles = [None] * self.nranks
        dist.all_gather_object(handles, bytes(buf.raw))
        check(lib.bb_comm_p2p_attach(self.handle, b''.join(handles)))
        self.p2p_capacity = int(capacity)
        self.set_option('allreduce_p2p', 1 if (local or os.environ.get('BB_ALLREDUCE', 'nccl') == 'p2p') else 0)

    def p2p_status(self):
        ready, err = c_int(), c_int()
        check(load().bb_comm_p2p_status(self.handle, ctypes.byref(ready), ctypes.byref(err)))
        return bool(ready.value), int(err.value)

    def allreduce_host(self, arr):
        arr = as_f64(arr)
        check(load().bb_comm_allreduce_host(self.handle, dptr(arr), arr.size))
        return arr
