from .sampler import Sampler, NaiveSampler  # noqa: F401
from .posterior_shape_model.psm import PosteriorShapeModelSampler  # noqa: F401
