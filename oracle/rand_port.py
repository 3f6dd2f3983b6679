"""TEST INFRASTRUCTURE ONLY -- pure-Python restatement of the reference's random-variate samplers.

* polya_gamma.pyx:40-216 (PolyaGammaDist) and scipy_ndtr.c:367-396 (log_ndtr)
* tilted_stable.pyx:65-332 (ExpTiltedStableDist)

The samplers draw from a numpy Generator over PCG64: Generator.random() and
Generator.standard_normal() call the same C routines (random_standard_uniform / random_standard_normal)
on the same bit generator as the reference's uniform.pyx:26 / normal.pyx:26, so with the same seed these
ports reproduce the compiled reference draw for draw (checked bit-for-bit in tests/test_oracle_rand.py
against fixtures generated from the reference). Loops are scalar Python: small cases only."""
import math
import numpy as np
from numpy.random import Generator, PCG64

PI = math.pi
THRESHOLD = 2.0 / PI
MAX_SERIES_TERMS = 100
DBL_EPSILON = 2.2204460492503131e-16


def ndtr(a):
    # scipy_ndtr.c:210-233 evaluates Phi through Cephes erf/erfc; libm's erfc agrees to a few ulp
    return 0.5 * math.erfc(-a / math.sqrt(2.0))


def log_ndtr(a):
    """scipy_ndtr.c:367-396."""
    if a > 6:
        return -ndtr(-a)
    if a > -20:
        return math.log(ndtr(a))
    log_lhs = -0.5 * a * a - math.log(-a) - 0.5 * math.log(2 * PI)
    last_total, rhs, numerator, denom_factor = 0.0, 1.0, 1.0, 1.0
    denom_cons = 1.0 / (a * a)
    sign, i = 1, 0
    while abs(last_total - rhs) > DBL_EPSILON:
        i += 1
        last_total = rhs
        sign = -sign
        denom_factor *= denom_cons
        numerator *= 2 * i - 1
        rhs += sign * numerator * denom_factor
    return log_lhs + math.log(rhs)


class _PortBase:

    def __init__(self, seed=None):
        self.set_seed(seed)

    def set_seed(self, seed):
        self.gen = Generator(PCG64(seed))

    def get_state(self):
        return self.gen.bit_generator.state

    def set_state(self, state):
        self.gen.bit_generator.state = state


class PolyaGammaPort(_PortBase):

    def uniform(self):
        return self.gen.random()

    def normal(self):
        return self.gen.standard_normal()

    def rand_polyagamma(self, shape, tilt):
        out = np.zeros(len(shape))
        for i in range(len(shape)):
            for _ in range(int(shape[i])):
                out[i] += 0.25 * self.tilted_jacobi(0.5 * abs(float(tilt[i])))   # polya_gamma.pyx:70-73,94-95
        return out

    def tilted_jacobi(self, z):
        """polya_gamma.pyx:97-112."""
        while True:
            X, a0 = self.proposal(z)
            U = self.uniform() * a0
            if self.accept(U, X, a0):
                return X

    def proposal(self, z):
        """polya_gamma.pyx:114-124."""
        K = 0.5 * z ** 2 + 0.125 * PI ** 2
        p_right = self.prob_to_right(z, K)
        if self.uniform() < p_right:
            X = self.left_truncated_exp(1.0 / K, THRESHOLD)
        else:
            X = self.right_truncated_invgauss(z, THRESHOLD)
        return X, self.series_term(0, X)

    @staticmethod
    def prob_to_right(z, K):
        """polya_gamma.pyx:126-139."""
        lm_expo = -math.log(K) - K * THRESHOLD + math.log(0.25 * PI)
        lm_1 = -z + log_ndtr((THRESHOLD * z - 1.0) / math.sqrt(THRESHOLD))
        lm_2 = z + log_ndtr(-(THRESHOLD * z + 1.0) / math.sqrt(THRESHOLD))
        ratio = math.exp(lm_1 - lm_expo) + math.exp(lm_2 - lm_expo)
        return 1.0 / (1.0 + ratio)

    @staticmethod
    def series_term(n, x):
        """polya_gamma.pyx:142-148."""
        lr = math.log(PI * (n + 0.5))
        if x <= THRESHOLD:
            lr += -1.5 * math.log(0.5 * x * PI) - 2 * (n + 0.5) ** 2 / x
        else:
            lr += -0.5 * x * PI ** 2 * (n + 0.5) ** 2
        return math.exp(lr)

    def accept(self, U, X, a0):
        """polya_gamma.pyx:150-174."""
        partial, n, sign = a0, 1, -1
        while True:
            partial += sign * self.series_term(n, X)
            n += 1
            if sign == -1:
                if U <= partial:
                    return True
            else:
                if U > partial:
                    return False
                if n >= MAX_SERIES_TERMS:
                    return True
            sign = -sign

    def left_truncated_exp(self, scale, trunc):
        return trunc - scale * math.log(1.0 - self.uniform())          # :176-177

    def left_truncated_chisq(self, trunc):
        while True:                                                     # :181-188
            X = self.left_truncated_exp(2.0, trunc)
            if self.uniform() <= math.sqrt(0.5 * PI / X):
                return X

    def right_truncated_invgauss(self, rate, trunc):
        """polya_gamma.pyx:191-216."""
        mean = float('inf') if rate == 0 else 1.0 / rate
        if mean > trunc:
            while True:
                X = 1.0 / self.left_truncated_chisq(0.5 * PI)
                if math.log(self.uniform()) < -0.5 * X * rate ** 2:
                    return X
        while True:
            V = self.normal() ** 2
            X = mean + 0.5 * mean * (mean * V - math.sqrt(4.0 * mean * V + mean ** 2 * V ** 2))
            if self.uniform() > mean / (mean + X):
                X = mean ** 2 / X
            if X < trunc:
                return X


def _exp(x):
    if x > 709:
        return float('inf')
    if x < -709:
        return 0.0
    return math.exp(x)


def _sinc(x):
    if abs(x) < 0.01:
        x2 = x * x
        return 1.0 - x2 / 6.0 * (1 - x2 / 20.0)
    return math.sin(x) / x


class TiltedStablePort(_PortBase):
    """tilted_stable.pyx:44-332."""

    def uniform(self):
        return self.gen.random()

    def normal(self):
        return self.gen.standard_normal()

    def sample(self, char_exp, tilt):
        tilt = np.asarray(tilt, dtype=float)
        out = np.zeros(tilt.size)
        for i in range(tilt.size):
            if tilt[i] ** char_exp < 2.0:                               # :103-108
                out[i] = self.divide_and_conquer(char_exp, float(tilt[i]))
            else:
                out[i] = self.double_rejection(char_exp, float(tilt[i]))
        return out

    def zolotarev(self, x, a):
        return math.pow(
            math.pow((1. - a) * _sinc((1. - a) * x), 1. - a) * math.pow(a * _sinc(a * x), a) / _sinc(x),
            1. / (1. - a))

    def zolotarev_pdf_exp(self, x, a):
        return _sinc(x) / (math.pow(_sinc(a * x), a) * math.pow(_sinc((1. - a) * x), 1. - a))

    def divide_and_conquer(self, a, tilt):
        m = max(1, int(math.floor(math.pow(tilt, a))))                  # :137-145
        c = math.pow(1. / m, 1. / a)
        X = 0.
        for _ in range(m):
            while True:                                                 # :147-153
                u1 = self.uniform()
                u2 = self.uniform()
                S = c * math.pow(-self.zolotarev(PI * u1, a) / math.log(u2), (1. - a) / a)
                if self.uniform() < _exp(-tilt * S):
                    break
            X += S
        return X

    def double_rejection(self, a, tilt):
        b = math.pow(tilt, a)                                           # :162-176
        while True:
            U, V, z = self.aux_rv(a, b)
            X, lacc = self.reference_rv(U, a, b, z)
            if lacc > math.log(V):
                return math.pow(X, -(1. - a) / a)

    def aux_rv(self, a, b):
        gam = b * a * (1. - a)                                          # :178-211
        xi = (1. + math.sqrt(2. * gam) * (2. + math.sqrt(.5 * PI))) / PI
        psi = math.sqrt(gam / PI) * (2. + math.sqrt(.5 * PI)) * _exp(-gam * PI * PI / 8.)
        while True:
            U = self.aux2_rv(xi, psi, gam)
            if U > PI:
                continue
            zeta = math.sqrt(self.zolotarev_pdf_exp(U, a))
            z = 1. / (1. - math.pow(1. + a * zeta / math.sqrt(gam), -1. / a))
            p = self.aux2_accept_prob(U, xi, psi, zeta, z, b, gam)
            if p > 0.:
                V = self.uniform() / p
                if U < PI and V <= 1.:
                    return U, V, z

    def aux2_rv(self, xi, psi, gam):
        w1 = math.sqrt(.5 * PI / gam) * xi                              # :213-238
        w2 = 2. * math.sqrt(PI) * psi
        w3 = xi * PI
        V = self.uniform()
        if gam >= 1:
            if V < w1 / (w1 + w2):
                return abs(self.normal()) / math.sqrt(gam)
            W = self.uniform()
            return PI * (1. - W * W)
        W = self.uniform()
        if V < w3 / (w2 + w3):
            return PI * W
        return PI * (1. - W * W)

    def aux2_accept_prob(self, U, xi, psi, zeta, z, b, gam):
        inv = PI * _exp(-b * (1. - 1. / (zeta * zeta))) / ((1. + math.sqrt(.5 * PI)) * math.sqrt(gam) / zeta + z)
        d = 0.                                                          # :240-256
        if U >= 0. and gam >= 1:
            d += xi * _exp(-gam * U * U / 2.)
        if 0. < U < PI:
            d += psi / math.sqrt(PI - U)
        if 0. <= U <= PI and gam < 1.:
            d += xi
        inv *= d
        return 1 / inv if inv != 0 else float('inf')

    def reference_rv(self, U, a, b, z):
        A = self.zolotarev(U, a)                                        # :258-296
        left = math.pow((1. - a) / a / A, a) * b
        right = left + math.sqrt(left * a / A)
        m_left = (right - left) * math.sqrt(.5 * PI)
        m_mid = right - left
        m_right = z / A
        total = m_left + m_mid + m_right
        V = self.uniform()
        N = E = 0.
        if V < m_left / total:
            N = self.normal()
            X = left - (right - left) * abs(N)
        elif V < (m_left + m_mid) / total:
            X = left + (right - left) * self.uniform()
        else:
            E = -math.log(self.uniform())
            X = right + E * m_right
        odds = (1. - a) / a                                             # :298-316
        if X < 0:
            lacc = -float('inf')
        else:
            lacc = -(A * (X - left)
                     + _exp(math.log(b) / a - odds * math.log(left)) * (math.pow(left / X, odds) - 1.))
            if X < left:
                lacc += N * N / 2.
            elif X > right:
                lacc += E
        return X, lacc
