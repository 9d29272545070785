from .sampler import Sampler, NaiveSampler  # noqa: F401
from .posterior_shape_model.psm import PosteriorShapeModelSampler  # noqa: F401
from .posterior_shape_model.sequence_sampler import SequencePSMSampler  # noqa: F401
from .posterior_shape_model.psm_skew import SkewPosteriorShapeModelSampler  # noqa: F401
from .posterior_shape_model.psm_skew_sequence import SequenceSkewPSMSampler  # noqa: F401
