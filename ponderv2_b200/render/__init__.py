"""NeuS volume renderer for the pretraining hot path (boundary B3): same registry-facing surface as the reference's
ponder/models/ponder/render_utils (`build_renderer(cfg)` -> `NeuSModel`, `RayBundle`)."""
from .neus import (AABBBoxCollider, NeuSModel, NeuSSampler, RayBundle, RGBDecoder, SDFDecoder, SDFField,  # noqa: F401
                   SemanticDecoder, build_renderer)
