"""B200-native differentiable sphere tracing -- drop-in for DIST-Renderer's SDFRenderer hot path.

Public surface (mirrors the reference's names):
  SDFRenderer                      core/sdfrenderer/renderer.py:12
  SDFRenderer_warp                 core/sdfrenderer/renderer_warp.py:13
  SDFRenderer_color                core/sdfrenderer/renderer_rgb.py:12
  SDFRenderer_deepsdf              core/sdfrenderer/renderer_deepsdf.py:10
  decode_sdf, decode_sdf_gradient  core/utils/decoder_utils.py:53,76
  decode_color                     core/utils/decoder_utils.py:94
  Decoder, load_decoder            core/graph/deep_sdf_decoder.py:18, core/utils/decoder_utils.py:7
  evaluation.Evaluator, evaluation.latent_vec_to_points, evaluation.compute_chamfer_distance
                                   core/evaluation/evaluator.py:8, transforms.py:13, eval_func.py:5
"""
from .decoder import Decoder, load_decoder  # noqa: F401
from .functional import decode_sdf, decode_sdf_gradient, decode_color  # noqa: F401
from .renderer import SDFRenderer  # noqa: F401
from .renderer_warp import SDFRenderer_warp  # noqa: F401
from .renderer_rgb import SDFRenderer_color  # noqa: F401
from .renderer_deepsdf import SDFRenderer_deepsdf  # noqa: F401
from . import evaluation  # noqa: F401
