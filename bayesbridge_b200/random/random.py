"""Random-variate front end (replaces random/random.py:5-41).

The Polya-Gamma and tilted-stable samplers are device kernels driven by counter-based Philox
streams: a generator's whole state is (seed, offset), where `offset` counts calls; element i of a
call always reads the stream keyed by its GLOBAL index, so a draw does not depend on how the
observations are sharded over GPUs."""
import numpy as np

from .. import _lib


class _PhiloxSampler:

    def __init__(self, ctx=None, seed=None):
        self._ctx = ctx
        self.set_seed(seed)

    @property
    def ctx(self):
        if self._ctx is None:
            self._ctx = _lib.Context.default()
        return self._ctx

    def set_seed(self, seed):
        if seed is None:
            seed = int(np.random.SeedSequence().generate_state(1, dtype=np.uint64)[0] >> 1)
        self.seed = int(seed)
        self.offset = 0

    def get_state(self):
        return {'bit_generator': 'Philox4x32-10', 'seed': self.seed, 'offset': self.offset}

    def set_state(self, state):
        self.seed, self.offset = int(state['seed']), int(state['offset'])

    def _next_offset(self):
        off = self.offset
        self.offset += 1
        return off


class DevicePolyaGamma(_PhiloxSampler):
    """PG(shape, tilt) draws (reference: random/polya_gamma/polya_gamma.pyx:40-74)."""

    def rand_polyagamma(self, shape, tilt, index_offset=0):
        if not (isinstance(shape, np.ndarray) and isinstance(tilt, np.ndarray)):
            raise TypeError('Input must be numpy arrays.')
        if not shape.size == tilt.size:
            raise ValueError('Input arrays must be of the same length.')
        if not np.issubdtype(shape.dtype, np.integer):
            raise ValueError('Shape parameter must be integers.')
        shape = np.ascontiguousarray(shape, dtype=np.int32)
        tilt = _lib.as_f64(tilt)
        out = np.zeros(shape.size)
        _lib.check(_lib.load().bb_pg_sample(
            self.ctx.handle, shape.size, _lib.iptr(shape), _lib.dptr(tilt),
            self.seed, self._next_offset(), int(index_offset), _lib.dptr(out)))
        return out

    def rand_unit_shape_polyagamma(self, tilt):
        if not isinstance(tilt, np.ndarray):
            raise TypeError('Input must be numpy arrays.')
        return self.rand_polyagamma(np.ones(tilt.size, dtype=np.int32), tilt)


class DeviceTiltedStable(_PhiloxSampler):
    """Exponentially tilted positive stable draws (reference: random/tilted_stable/tilted_stable.pyx:65-135)."""

    def sample(self, char_exponent, tilt, method=None, index_offset=0):
        if not isinstance(tilt, np.ndarray):
            raise TypeError('Tilt parameter must be a numpy array.')
        if isinstance(char_exponent, np.ndarray):
            if char_exponent.size != tilt.size:
                raise ValueError('Input arrays must be of the same length.')
            if not np.all(char_exponent == char_exponent.flat[0]):
                raise NotImplementedError('The device sampler takes one characteristic exponent per call.')
            char_exponent = float(char_exponent.flat[0])
        elif not isinstance(char_exponent, (float, np.floating)):
            raise TypeError('Characteristic exponent must be float or numpy array.')
        if not char_exponent < 1:
            raise ValueError('Characteristic exponent must be smaller than 1.')
        if not np.all(tilt > 0):
            raise ValueError('Tilting parameter must be positive.')
        if method is not None:
            raise NotImplementedError('The device sampler chooses the method itself.')
        tilt = _lib.as_f64(tilt)
        out = np.zeros(tilt.size)
        _lib.check(_lib.load().bb_tilted_stable_sample(
            self.ctx.handle, tilt.size, float(char_exponent), _lib.dptr(tilt),
            self.seed, self._next_offset(), int(index_offset), _lib.dptr(out)))
        return out


class BasicRandom():
    """Owns the random streams of one chain: numpy's global generator for the scalar Gamma draws
    (and for the CG noise in `noise='host'` mode), and Philox streams for PG, tilted stable and the
    device-generated CG noise."""

    def __init__(self, seed=None, ctx=None):
        self.np_random = np.random
        self.pg = DevicePolyaGamma(ctx)
        self.ts = DeviceTiltedStable(ctx)
        self.cg = _PhiloxSampler(ctx)     # (seed, offset) of the CG right-hand-side noise
        self.set_seed(seed)

    def set_seed(self, seed):
        # Same draws from numpy's stream as the reference (random.py:17-22), so a chain run with
        # host-generated noise consumes numpy's generator identically.
        self.np_random.seed(seed)
        pg_seed = np.random.randint(1, 1 + np.iinfo(np.int32).max)
        ts_seed = np.random.randint(1, 1 + np.iinfo(np.int32).max)
        self.pg.set_seed(pg_seed)
        self.ts.set_seed(ts_seed)
        self.cg.set_seed((int(pg_seed) << 31) ^ int(ts_seed))

    def get_state(self):
        return {
            'numpy': self.np_random.get_state(),
            'tilted_stable': self.ts.get_state(),
            'polya_gamma': self.pg.get_state(),
            'cg_noise': self.cg.get_state(),
        }

    def set_state(self, state):
        self.np_random.set_state(state['numpy'])
        self.ts.set_state(state['tilted_stable'])
        self.pg.set_state(state['polya_gamma'])
        if 'cg_noise' in state:
            self.cg.set_state(state['cg_noise'])

    def polya_gamma(self, shape, tilt, index_offset=0):
        return self.pg.rand_polyagamma(shape, tilt, index_offset) if isinstance(self.pg, DevicePolyaGamma) \
            else self.pg.rand_polyagamma(shape, tilt)

    def tilted_stable(self, char_exponent, tilt):
        return self.ts.sample(char_exponent, tilt)
