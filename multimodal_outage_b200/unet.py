"""Feature producers either side of the gwnet call, stock PyTorch.

The reference builds the 320-channel gwnet input per sample and per county with Python loops
(``/root/reference/models/unet.py``): ``Contraction.forward`` (:106-126) runs the five-stage conv stack 67 times,
``Encoder.forward`` (:138-149) runs two Linear layers 67 times, and ``Modified_UNET.forward`` (:219-231) calls the
spatio-temporal network once per batch element with a ``[67, h, 320]`` tensor.  Here the same parameters
(identical ``state_dict`` keys and shapes, so reference checkpoints load with ``strict=True``) are applied with the
counties - and, opt-in, the batch - folded into the batch dimension of ONE conv / matmul / gwnet call:

* ``literal=True`` forwards reproduce the reference loop structure exactly (same BatchNorm batches: one county's ``h``
  frames at a time; one gwnet call per sample with BatchNorm over ``67*h`` positions) - the parity mode;
* the default batched forwards give identical results in eval mode and whenever BatchNorm statistics are frozen; in
  training mode BatchNorm sees all counties (and samples) at once - a deliberate, documented semantic difference,
  which is why batching is a switch on the call and not a silent replacement.

Only the encoder side is here: the decoder / expansion half of the UNet is downstream of the hot path and out of scope.
"""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.nn as nn

from .graph_wavenet import gwnet

# module-level hyper-parameters of the reference (unet.py:33-38)
image_dimension = 128
n_counties = 67
feature_vector_size = 256
time_embed_size = 64
compression_factor = 4


def _conv_pair(c_in: int, c_out: int) -> nn.Sequential:
    layers: List[nn.Module] = []
    for a, b in ((c_in, c_out), (c_out, c_out)):
        layers += [nn.Conv2d(a, b, kernel_size=3, padding=1, bias=False), nn.BatchNorm2d(b), nn.ReLU(inplace=True)]
    return nn.Sequential(*layers)


class DoubleConv(nn.Module):
    """(conv3x3 -> BN -> ReLU) x 2; parameter names ``double_conv.{0,1,3,4}.*`` as in unet.py:40-53."""

    def __init__(self, in_channels: int, out_channels: int):
        super().__init__()
        self.double_conv = _conv_pair(in_channels, out_channels)

    def forward(self, x):
        return self.double_conv(x)


class Down(nn.Module):
    """2x2 max-pool then DoubleConv; parameter names ``maxpool_conv.1.double_conv.*`` as in unet.py:55-66."""

    def __init__(self, in_channels: int, out_channels: int):
        super().__init__()
        self.maxpool_conv = nn.Sequential(nn.MaxPool2d(2), DoubleConv(in_channels, out_channels))

    def forward(self, x):
        return self.maxpool_conv(x)


class Contraction(nn.Module):
    """UNet down path 128^2 -> 8^2, channels C -> 4 -> 8 -> 16 -> 32 -> 64 (unet.py:95-126).

    input ``[..., n_counties, horizon, C, H, W]`` (the reference passes one sample: ``[67, h, C, 128, 128]``) ->
    ``[..., n_counties, horizon, 64 * (H/16) * (W/16)]``; ``feature_maps[i]`` = the four skip tensors stacked over
    counties like the reference's ``self.feature_maps`` (:121-122)."""

    def __init__(self, in_channels: int, horizon: int):
        super().__init__()
        self.horizon = horizon
        self.inc = DoubleConv(in_channels, 4)
        self.down1 = Down(4, 8)
        self.down2 = Down(8, 16)
        self.down3 = Down(16, 32)
        self.down4 = Down(32, 64)
        self.feature_maps: List[torch.Tensor] = []

    def _stack(self, x):
        maps = []
        for stage in (self.inc, self.down1, self.down2, self.down3):
            x = stage(x)
            maps.append(x)
        return self.down4(x), maps

    def forward(self, input: torch.Tensor, literal: bool = False) -> torch.Tensor:
        lead = input.shape[:-3]                       # (..., counties, horizon)
        if literal:
            # the reference's schedule: one county (= a batch of `horizon` frames) per pass through the stack, so a
            # training-mode BatchNorm normalises over that county's frames only and updates its running statistics
            # once per county (unet.py:110-120)
            if input.dim() != 5:
                raise ValueError('literal mode takes one sample [counties, horizon, C, H, W] like the reference')
            outs, maps = [], [[] for _ in range(4)]
            for county in range(input.shape[0]):
                y, m = self._stack(input[county])
                outs.append(y)
                for i in range(4):
                    maps[i].append(m[i])
            self.feature_maps = [torch.stack(m) for m in maps]
            return torch.stack(outs).reshape(*lead, -1)
        y, maps = self._stack(input.reshape(-1, *input.shape[-3:]))          # counties (and batch) as ONE conv batch
        self.feature_maps = [m.reshape(*lead, *m.shape[1:]) for m in maps]
        return y.reshape(*lead, -1)


class Encoder(nn.Module):
    """fc 4096 -> 1024 -> relu -> dropout(0.3) -> fc 256 -> relu per (county, frame) (unet.py:128-149).  The reference
    loops over counties; a Linear layer is row-wise, so one matmul over all rows is the same function (dropout draws
    differ in order only)."""

    def __init__(self):
        super().__init__()
        self.compression_factor = compression_factor
        self.downsized_image_dimension = image_dimension / 16
        self.first_layer_size = int(self.downsized_image_dimension * self.downsized_image_dimension * 64)
        self.fc1 = nn.Linear(self.first_layer_size, int(self.first_layer_size / self.compression_factor))
        self.dropout1 = nn.Dropout(p=0.3)
        self.fc2 = nn.Linear(int(self.first_layer_size / self.compression_factor), feature_vector_size)

    def forward(self, input: torch.Tensor, literal: bool = False) -> torch.Tensor:
        if literal:
            return torch.stack([torch.relu(self.fc2(self.dropout1(torch.relu(self.fc1(input[c])))))
                                for c in range(input.shape[0])])
        return torch.relu(self.fc2(self.dropout1(torch.relu(self.fc1(input)))))


class UNetGWNetEncoder(nn.Module):
    """Encoder half of the reference's ``Modified_UNET`` up to and including the gwnet call (unet.py:201-225), with the
    same sub-module names (``contraction``, ``encoder``, ``st_gnn``), so the matching entries of a reference
    checkpoint load unchanged.

    ``forward(input [B,67,h,C,128,128], time_dim [B,67,h,64]) -> [B,67,h,256]`` (what the reference hands its decoder).

    * ``batched=False`` (default): the literal schedule - one gwnet call per sample on the ``[67,h,320]`` buffer
      (graph_wavenet.py:189 reinterprets it as ``[1,320,67,h]``), BatchNorm over ``67*h`` positions per call.
    * ``batched=True``: the B samples are stacked and sent through ONE gwnet call ``[B,320,67,h]`` - each sample's buffer
      is reinterpreted exactly as in the literal call, so the results are identical whenever BatchNorm does not depend
      on the batch (eval mode); in training mode its statistics span the batch.
    """

    def __init__(self, horizon: int, input_channels: int = 3, device='cuda', **gwnet_kwargs):
        super().__init__()
        self.horizon = horizon
        self.contraction = Contraction(input_channels, horizon)
        self.encoder = Encoder()
        self.st_gnn_in_dim = feature_vector_size + time_embed_size
        self.st_gnn = gwnet(device=device, in_dim=self.st_gnn_in_dim, out_dim=feature_vector_size, horizon=horizon,
                            **gwnet_kwargs)

    def features(self, input: torch.Tensor, time_dim: torch.Tensor, literal: bool = False) -> torch.Tensor:
        """[B,67,h,C,H,W], [B,67,h,64] -> the gwnet input buffers [B,67,h,320]."""
        if literal:
            rows = [self.encoder(self.contraction(input[b], literal=True), literal=True) for b in range(input.shape[0])]
            feat = torch.stack(rows)
        else:
            feat = self.encoder(self.contraction(input))
        return torch.cat((feat, time_dim), dim=-1)

    def forward(self, input: torch.Tensor, time_dim: torch.Tensor, batched: bool = False,
                literal_features: Optional[bool] = None) -> torch.Tensor:
        literal_features = (not batched) if literal_features is None else literal_features
        x = self.features(input, time_dim, literal=literal_features).contiguous()
        B, V, h, F = x.shape
        if not batched:
            return torch.stack([self.st_gnn(x[b]) for b in range(B)])          # unet.py:221-225
        y = self.st_gnn(x.view(B, F, V, h))                                    # the same reinterpretation, B at once
        return y.reshape(B, V, h, -1)


def load_reference_checkpoint(module: nn.Module, path: str, prefix: str = 'model.', strict: bool = True,
                              map_location='cpu', weights_only: bool = True):
    """Loads the reference's on-disk format (lit.py:187-196: Lightning ``ModelCheckpoint`` of ``LitModified_UNET``, whose
    model lives under ``self.model``): a dict with a ``'state_dict'`` entry whose keys are ``model.<module path>``.
    Plain ``state_dict`` files (``torch.save(model.state_dict())``) are accepted too.  Keys outside ``module`` (the
    decoder half when ``module`` is a ``UNetGWNetEncoder``, or everything but ``model.st_gnn.*`` when it is a bare
    ``gwnet`` and ``prefix='model.st_gnn.'``) are dropped; the module's own keys are loaded with ``strict``.
    ``weights_only=False`` is needed for Lightning checkpoints that pickle callback / hyper-parameter objects (only for
    files you trust)."""
    ckpt = torch.load(path, map_location=map_location, weights_only=weights_only)
    sd = ckpt['state_dict'] if isinstance(ckpt, dict) and 'state_dict' in ckpt else ckpt
    own = set(module.state_dict().keys())
    picked = {}
    for k, v in sd.items():
        name = k[len(prefix):] if prefix and k.startswith(prefix) else (k if not prefix else None)
        if name is not None and name in own:
            picked[name] = v
    return module.load_state_dict(picked, strict=strict)
