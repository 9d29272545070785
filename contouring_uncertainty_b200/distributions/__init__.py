from .bivariatenormal import BivariateNormal  # noqa: F401
from .bivariateskewnormal import BivariateSkewNormal  # noqa: F401
