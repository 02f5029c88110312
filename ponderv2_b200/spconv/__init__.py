"""Drop-in for the `spconv` package surface the reference imports (`import spconv.pytorch as spconv`)."""
from . import pytorch  # noqa: F401

__version__ = "2.3.6+pv2b200"
