"""DeepSDF decoder network -- host-side mirror of the reference's decoder interface.

Drop-in for `core/graph/deep_sdf_decoder.py:18-111` (class ``Decoder``; constructor arguments, submodule and
parameter names -- ``lin{l}.weight_g / weight_v / bias`` for weight-normalised layers, ``lin{l}.weight / bias``
otherwise, ``bn{l}`` LayerNorm -- are kept so reference / upstream-DeepSDF checkpoints load with
``load_state_dict``), and for `core/utils/decoder_utils.py:7-51` (``load_decoder``).

This module only *describes* the network and holds its parameters.  The hot path never runs these layers through
PyTorch: `DecoderPlan` (plan.py) folds weight-norm and the latent code away and hands flat fp32 buffers to the
CUDA kernels.  ``Decoder.inference`` on arbitrary (per-row latent) inputs is the generic module forward that a
training loop would use; it is plain PyTorch by design and is not on the rendering path.
"""
import json
import os

import torch
import torch.nn as nn
import torch.nn.functional as F


class WNLinear(nn.Module):
    """Linear layer with weight normalisation over output rows, parameterised as (weight_g, weight_v).

    Same parameterisation/keys as ``nn.utils.weight_norm(nn.Linear(...))`` used at deep_sdf_decoder.py:59:
    ``weight = weight_g * weight_v / ||weight_v||_row``.
    """

    def __init__(self, in_features, out_features):
        super().__init__()
        lin = nn.Linear(in_features, out_features)  # default init, then split into (g, v)
        self.in_features, self.out_features = in_features, out_features
        self.weight_g = nn.Parameter(lin.weight.detach().norm(2, dim=1, keepdim=True).clone())
        self.weight_v = nn.Parameter(lin.weight.detach().clone())
        self.bias = nn.Parameter(lin.bias.detach().clone())

    @property
    def weight(self):
        return torch._weight_norm(self.weight_v, self.weight_g, 0)

    def forward(self, x):
        return F.linear(x, self.weight, self.bias)


class Decoder(nn.Module):
    """DeepSDF MLP.  Constructor signature follows deep_sdf_decoder.py:19-32."""

    def __init__(self, latent_size, dims, last_dim=1, dropout=None, dropout_prob=0.0, norm_layers=(),
                 latent_in=(), weight_norm=False, xyz_in_all=None, use_tanh=False, latent_dropout=False):
        super().__init__()
        self.latent_size = latent_size
        dims = [latent_size + 3] + list(dims) + [last_dim]
        self.dims = dims
        self.num_layers = len(dims)
        self.norm_layers = norm_layers
        self.latent_in = latent_in
        self.latent_dropout = latent_dropout
        self.xyz_in_all = xyz_in_all
        self.weight_norm = weight_norm
        self.use_tanh = use_tanh
        self.dropout_prob = dropout_prob
        self.dropout = dropout
        for l in range(self.num_layers - 1):
            if l + 1 in latent_in:
                out_dim = dims[l + 1] - dims[0]
            else:
                out_dim = dims[l + 1]
                if xyz_in_all and l != self.num_layers - 2:
                    out_dim -= 3
            if weight_norm and l in norm_layers:
                self.add_module("lin%d" % l, WNLinear(dims[l], out_dim))
            else:
                self.add_module("lin%d" % l, nn.Linear(dims[l], out_dim))
            if (not weight_norm) and norm_layers is not None and l in norm_layers:
                self.add_module("bn%d" % l, nn.LayerNorm(out_dim))

    def layer(self, l):
        return getattr(self, "lin%d" % l)

    def latent_size_regul(self, lat_vecs):  # deep_sdf_decoder.py:75-77
        return lat_vecs.pow(2).mean(1)

    def inference(self, input):
        """Forward, rows = [latent | xyz] -> (K, 1)  (deep_sdf_decoder.py:80-111).

        Under ``torch.no_grad()`` on a CUDA device, rows that all carry the SAME latent code (the layout decode_sdf
        builds, decoder_utils.py:61-62) are evaluated by the fused CUDA engines; anything else (per-row latents,
        gradients w.r.t. the weights, CPU tensors, unsupported specs) runs the generic PyTorch layers below."""
        if (not torch.is_grad_enabled()) and input.is_cuda and input.dim() == 2 and input.shape[0] > 0 \
                and input.shape[1] == self.latent_size + 3 and not self.training:
            out = self._inference_fused(input)
            if out is not None:
                return out
        return self._inference_torch(input)

    def _inference_fused(self, input):
        from .functional import decode_sdf
        if self.dims[-1] != 1:          # colour networks (last_dim = 3) are not covered by the fused engines
            return None
        L = self.latent_size
        lat = input[:1, :L]
        if L > 0 and not bool((input[:, :L] == lat).all()):
            return None
        try:
            return decode_sdf(self, lat if L > 0 else None, input[:, L:], clamp_dist=None, no_grad=True)
        except NotImplementedError:
            return None

    def _inference_torch(self, input):
        xyz = input[:, -3:]
        x = input
        if input.shape[1] > 3 and self.latent_dropout:
            x = torch.cat([F.dropout(input[:, :-3], p=0.2, training=self.training), xyz], 1)
        last = self.num_layers - 2
        for l in range(self.num_layers - 1):
            if l in self.latent_in:
                x = torch.cat([x, input], 1)
            elif l != 0 and self.xyz_in_all:
                x = torch.cat([x, xyz], 1)
            x = self.layer(l)(x)
            if l == last and self.use_tanh:
                x = torch.tanh(x)
            if l < last:
                if self.norm_layers is not None and l in self.norm_layers and not self.weight_norm:
                    x = getattr(self, "bn%d" % l)(x)
                x = F.relu(x)
                if self.dropout is not None and l in self.dropout:
                    x = F.dropout(x, p=self.dropout_prob, training=self.training)
        return torch.tanh(x)

    def forward(self, input):
        return self.inference(input)


def load_decoder(experiment_directory, checkpoint_num=None, color_size=None, experiment_directory_color=None,
                 parallel=True):
    """specs.json + ModelParameters/<ckpt>.pth loader with the reference's signature (decoder_utils.py:7-51).

    * SDF decoder (``color_size=None``): ``Decoder(CodeLength, **NetworkSpecs)``, weights from
      ``<experiment_directory>/ModelParameters/<checkpoint_num>.pth`` (keys carry the ``module.`` prefix of the
      DataParallel wrapper they were saved from).
    * colour decoder (``color_size=k``): latent = shape code + colour code (``CodeLength + k``), ``dims[3] += k`` so the
      ``latent_in`` concatenation still adds up, ``last_dim=3``; weights from ``experiment_directory_color`` (saved
      without the prefix).
    ``parallel=True`` (the reference's default) returns the module wrapped in ``torch.nn.DataParallel`` -- call sites are
    written ``load_decoder(...).module.cuda()``; ``parallel=False`` returns the bare module.
    """
    specs_filename = os.path.join(experiment_directory, "specs.json")
    if not os.path.isfile(specs_filename):
        raise Exception('The experiment directory does not include specifications file "specs.json"')
    with open(specs_filename) as f:
        specs = json.load(f)
    net = dict(specs["NetworkSpecs"])
    latent_size = specs["CodeLength"]
    if color_size is not None:
        net["dims"] = list(net["dims"])
        net["dims"][3] = net["dims"][3] + color_size
        latent_size = latent_size + color_size
        net["last_dim"] = 3
    decoder = Decoder(latent_size, **net)
    if checkpoint_num is not None:
        src = experiment_directory_color if color_size is not None else experiment_directory
        saved = torch.load(os.path.join(src, "ModelParameters", checkpoint_num + ".pth"), map_location="cpu")
        sd = {(k[len("module."):] if k.startswith("module.") else k): v for k, v in saved["model_state_dict"].items()}
        decoder.load_state_dict(sd)
    return torch.nn.DataParallel(decoder) if parallel else decoder
