"""bayesbridge_b200: the CG-accelerated Gibbs sampler of OHDSI/bayes-bridge with its coefficient update
and Polya-Gamma draws running as sm_100a CUDA kernels (libbbgpu.so). Same public names as the
reference's `bayesbridge` package."""
from .bayesbridge import BayesBridge
from .batched import BatchedBayesBridge
from .gibbs_util import SamplerOptions
from .prior import RegressionCoefPrior
from .model import RegressionModel
from ._lib import Context

__all__ = ['BayesBridge', 'BatchedBayesBridge', 'SamplerOptions', 'RegressionCoefPrior', 'RegressionModel', 'Context']
