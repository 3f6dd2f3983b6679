from .factory import RegressionModel
from .linear_model import LinearModel
from .logistic_model import LogisticModel
