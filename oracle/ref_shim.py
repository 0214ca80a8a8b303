"""TEST INFRASTRUCTURE ONLY -- loader for the *unmodified* reference (B1ueber2y/DIST-Renderer).

Executes the reference's own sources from /root/reference on torch-cpu with the minimal
compatibility remedies of SURVEY.md section 8c / Appendix A (two environment shims, two one-line
load-time text patches, stub packages so core/__init__.py's star-imports of absent third-party
packages are not executed).  No reference source is copied into this repository; the files are
read from where they lie and exec'd.

This module only works where /root/reference exists (the build container).  It is used by
oracle/make_golden.py to produce tests/golden/*.npz and by the CPU tests that pin
oracle/sdf_oracle.py against the real reference.  Nothing on the GPU box imports it.
"""
import os
import sys
import types

import torch

REF = os.environ.get("DIST_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isfile(os.path.join(REF, "core", "sdfrenderer", "renderer.py"))


_PATCH = {
    # decoder_utils.py:84 -- grad_outputs shaped like the (N,3) points for an (N,1) output; torch>=2 rejects it
    "core/utils/decoder_utils.py": [
        ("grad_outputs=torch.ones_like(points_batch)", "grad_outputs=torch.ones_like(sdf)")],
    # renderer.py:873-874 -- index aliases destination
    "core/sdfrenderer/renderer.py": [
        ("valid_mask[valid_mask] = valid_mask_render", "valid_mask[valid_mask.clone()] = valid_mask_render")],
    # loss_utils.py:43 -- uint8 mask only converted for cuda tensors
    "core/utils/loss_utils.py": [
        ("img.type() == 'torch.cuda.ByteTensor'", "img.dtype == torch.uint8"),
        # loss_utils.py:24 -- written for torch 1.1 whose grid_sample convention is align_corners=True
        ("output = F.grid_sample(img, vgrid)", "output = F.grid_sample(img, vgrid, align_corners=True)")],
    # renderer_warp.py:43 -- hard-coded .cuda()
    "core/sdfrenderer/renderer_warp.py": [
        ("torch.ones(xyz_proj.shape[1]).byte().cuda()", "torch.ones(xyz_proj.shape[1]).bool().to(xyz_proj.device)")],
}

_loaded = {}


def _stub(name, rel):
    m = types.ModuleType(name)
    m.__path__ = [REF + rel]
    sys.modules[name] = m


def _scatter_max(src, index, *a, **k):
    # stand-in for torch_scatter.scatter_max (renderer.py:677), bool -> uint8 amax
    s = src.to(torch.uint8) if src.dtype == torch.bool else src
    o = torch.zeros(int(index.max()) + 1, dtype=s.dtype, device=s.device).scatter_reduce(0, index, s, "amax")
    return o.to(src.dtype), None


def _load(mod, rel):
    src = open(f"{REF}/{rel}").read()
    for a, b in _PATCH.get(rel, []):
        assert a in src, (rel, a)
        src = src.replace(a, b)
    m = types.ModuleType(mod)
    m.__file__ = f"{REF}/{rel}"
    sys.modules[mod] = m
    exec(compile(src, m.__file__, "exec"), m.__dict__)
    return m


def load_warp():
    """Returns the reference's SDFRenderer_warp class (core/sdfrenderer/renderer_warp.py), CPU-runnable."""
    if "W" in _loaded:
        return _loaded["W"]
    R, _, _ = load()
    sys.modules["renderer"] = R                      # renderer_warp.py:6 does `from renderer import SDFRenderer`
    up = os.path.join(REF, "core", "utils")
    if up not in sys.path:
        sys.path.append(up)                          # loss_utils.py:7 `from pytorch_ssim import loss_ssim`
    _load("core.utils.loss_utils", "core/utils/loss_utils.py")
    W = _load("core.sdfrenderer.renderer_warp", "core/sdfrenderer/renderer_warp.py")
    _loaded["W"] = W.SDFRenderer_warp
    return _loaded["W"]


def load_color():
    """Returns the reference's SDFRenderer_color class (core/sdfrenderer/renderer_rgb.py), CPU-runnable."""
    if "C" in _loaded:
        return _loaded["C"]
    R, _, _ = load()
    sys.modules["renderer"] = R                      # renderer_rgb.py:5 does `from renderer import SDFRenderer`
    C = _load("core.sdfrenderer.renderer_rgb", "core/sdfrenderer/renderer_rgb.py")
    _loaded["C"] = C.SDFRenderer_color
    return _loaded["C"]


def load_deepsdf():
    """Returns the reference's SDFRenderer_deepsdf class (core/sdfrenderer/renderer_deepsdf.py), CPU-runnable."""
    if "D" in _loaded:
        return _loaded["D"]
    R, _, _ = load()
    sys.modules["renderer"] = R                      # renderer_deepsdf.py:5 does `from renderer import SDFRenderer`
    D = _load("core.sdfrenderer.renderer_deepsdf", "core/sdfrenderer/renderer_deepsdf.py")
    _loaded["D"] = D.SDFRenderer_deepsdf
    return _loaded["D"]


def load_loss_single():
    """The reference's compute_all_loss (core/inv_optimizer/loss_single.py) with its loss_utils patched for CPU uint8 masks."""
    if "LS" in _loaded:
        return _loaded["LS"]
    load()
    up = os.path.join(REF, "core", "utils")
    if up not in sys.path:
        sys.path.append(up)                          # loss_utils.py:7 `from pytorch_ssim import loss_ssim`
    if "core.utils.loss_utils" not in sys.modules or not hasattr(sys.modules["core.utils.loss_utils"], "compute_loss_mask"):
        _load("core.utils.loss_utils", "core/utils/loss_utils.py")
    _stub("core.inv_optimizer", "/core/inv_optimizer")
    LS = _load("core.inv_optimizer.loss_single", "core/inv_optimizer/loss_single.py")
    _loaded["LS"] = LS.compute_all_loss
    return _loaded["LS"]


def load_create_mesh():
    """The reference's core/evaluation/create_mesh.py with skimage / plyfile stubbed (absent here; only the sampling
    half is exercised), `.cuda()` / `.cpu()` round trips neutralised and the torch>=1.6 true-division of
    create_mesh.py:23-24 restored to the integer division upstream DeepSDF intends."""
    if "CM" in _loaded:
        return _loaded["CM"]
    _, DU, _ = load()
    sk = types.ModuleType("skimage")
    sk.measure = types.ModuleType("skimage.measure")
    sys.modules.setdefault("skimage", sk)
    sys.modules.setdefault("skimage.measure", sk.measure)
    sys.modules.setdefault("plyfile", types.ModuleType("plyfile"))
    sys.modules["decoder_utils"] = DU                 # create_mesh.py:8 `from decoder_utils import decode_sdf`
    _PATCH["core/evaluation/create_mesh.py"] = [
        ("samples[:, 1] = (overall_index.long() / N) % N", "samples[:, 1] = (overall_index.long() // N) % N"),
        ("samples[:, 0] = ((overall_index.long() / N) / N) % N", "samples[:, 0] = ((overall_index.long() // N) // N) % N"),
        ("0:3].cuda()", "0:3]"),
    ]
    _loaded["CM"] = _load("core.evaluation.create_mesh", "core/evaluation/create_mesh.py")
    return _loaded["CM"]


def load_eval_func():
    """The reference's core/evaluation/eval_func.py as it is (numpy + scipy cKDTree, both installed here)."""
    if "EF" not in _loaded:
        load()
        _stub("core.evaluation", "/core/evaluation")
        _loaded["EF"] = _load("core.evaluation.eval_func", "core/evaluation/eval_func.py")
    return _loaded["EF"]


def load():
    """Returns (renderer_module, decoder_utils_module, DecoderClass) of the reference."""
    if _loaded:
        return _loaded["R"], _loaded["DU"], _loaded["Decoder"]
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF)
    for n, p in [("core", "/core"), ("core.utils", "/core/utils"), ("core.visualize", "/core/visualize"),
                 ("core.graph", "/core/graph"), ("core.sdfrenderer", "/core/sdfrenderer")]:
        _stub(n, p)
    sys.modules.setdefault("trimesh", types.ModuleType("trimesh"))  # render_utils.py:1
    ts = types.ModuleType("torch_scatter")
    ts.scatter_max = _scatter_max
    sys.modules["torch_scatter"] = ts
    if not torch.cuda.is_available():
        torch.cuda.synchronize = lambda *a, **k: None  # profiler.py:8
        _to = torch.Tensor.to

        def to(self, *a, **k):  # renderer.py:330 etc: .to(get_device()) with -1 on CPU
            if a and type(a[0]) is int and a[0] == -1:
                a = ("cpu",) + a[1:]
            return _to(self, *a, **k)
        torch.Tensor.to = to
    DU = _load("core.utils.decoder_utils", "core/utils/decoder_utils.py")
    R = _load("core.sdfrenderer.renderer", "core/sdfrenderer/renderer.py")
    from core.graph.deep_sdf_decoder import Decoder
    _loaded.update(R=R, DU=DU, Decoder=Decoder)
    return R, DU, Decoder
