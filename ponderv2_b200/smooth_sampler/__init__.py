"""Drop-in for `smooth_sampler` (reference: libs/smooth-sampler/smooth_sampler/__init__.py:1)."""
from . import _C  # noqa: F401
from .modules import SmoothSampler, SmoothSamplerBackward, padding_mode_enum  # noqa: F401
