"""TEST INFRASTRUCTURE ONLY -- Philox4x32-10 (Salmon, Moraes, Dror, Shaw 2011; the published algorithm)
and the per-element stream layout of the device kernels, restated in Python so that the oracle's
sampler ports (rand_port.py) can be driven by exactly the uniforms / normals a device thread sees.
Stream key: (seed); counter: (index lo, offset lo, draw counter, stream id | index hi | offset hi)."""
import math

M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85
MASK = 0xFFFFFFFF
STREAM_EPS1, STREAM_EPS2, STREAM_PG, STREAM_TS = 0, 1, 2, 3


def philox4x32_10(ctr, key):
    c = list(ctr)
    k0, k1 = key
    for _ in range(10):
        p0, p1 = M0 * c[0], M1 * c[2]
        hi0, lo0, hi1, lo1 = p0 >> 32, p0 & MASK, p1 >> 32, p1 & MASK
        c = [hi1 ^ c[1] ^ k0, lo1, hi0 ^ c[3] ^ k1, lo0]
        k0, k1 = (k0 + W0) & MASK, (k1 + W1) & MASK
    return c


class PhiloxStream:

    def __init__(self, seed, offset, index, stream_id):
        self.key = (seed & MASK, (seed >> 32) & MASK)
        self.ctr = [index & MASK, offset & MASK, 0,
                    ((stream_id << 24) | (((index >> 32) & 0xFF) << 16) | ((offset >> 32) & 0xFFFF)) & MASK]
        self.spare = None

    def uniform(self):
        if self.spare is not None:
            u, self.spare = self.spare, None
            return u
        o = philox4x32_10(self.ctr, self.key)
        self.ctr[2] = (self.ctr[2] + 1) & MASK
        a = ((o[0] << 32) | o[1]) >> 11
        b = ((o[2] << 32) | o[3]) >> 11
        self.spare = (b + 0.5) / 9007199254740992.0
        return (a + 0.5) / 9007199254740992.0

    def normal(self):
        u1, u2 = self.uniform(), self.uniform()
        return math.sqrt(-2.0 * math.log(u1)) * math.cos(6.283185307179586476925286766559 * u2)


def philox_normals(n, stream_id, seed, offset, index_offset=0):
    import numpy as np
    return np.array([PhiloxStream(seed, offset, index_offset + i, stream_id).normal() for i in range(n)])


def pg_with_device_streams(shape, tilt, seed, offset, index_offset=0):
    """The PG port (rand_port.PolyaGammaPort) fed, element by element, with the device's streams."""
    import numpy as np
    from .rand_port import PolyaGammaPort
    port = PolyaGammaPort(0)
    out = np.zeros(len(shape))
    for i in range(len(shape)):
        rs = PhiloxStream(seed, offset, index_offset + i, STREAM_PG)
        port.uniform, port.normal = rs.uniform, rs.normal
        for _ in range(int(shape[i])):
            out[i] += 0.25 * port.tilted_jacobi(0.5 * abs(float(tilt[i])))
    return out


def ts_with_device_streams(char_exp, tilt, seed, offset, index_offset=0):
    import numpy as np
    from .rand_port import TiltedStablePort
    port = TiltedStablePort(0)
    out = np.zeros(len(tilt))
    for i in range(len(tilt)):
        rs = PhiloxStream(seed, offset, index_offset + i, STREAM_TS)
        port.uniform, port.normal = rs.uniform, rs.normal
        out[i] = port.sample(char_exp, np.array([tilt[i]]))[0]
    return out
