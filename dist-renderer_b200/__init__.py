"""B200-native differentiable sphere tracing -- drop-in for DIST-Renderer's SDFRenderer hot path."""
from .decoder import Decoder, load_decoder  # noqa: F401
