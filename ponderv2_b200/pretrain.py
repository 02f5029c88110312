"""One pretraining step of the hot path: SpUNet backbone -> densify -> dense projection -> NeuS renderer -> losses.

Host-side mirror of `PonderIndoor.forward` (ponder_indoor_base.py:694-706) with its `extract_feature` (:120-175,
masking off as in the shipped indoor configs), `prepare_volume` (:635-640), `render_func` (:642-674) and `render_loss`
(:676-679).  Ray preparation (`to_unit_cube` / `ray_sample`, :344-620) is host glue that SURVEY.md §8(f) ranks as
"next"; this module consumes already-prepared rays (`ray_o`, `ray_d` [B,R,3]; `rgb` [B*R,3]; `depth` [B*R,1]) exactly
in the form `ray_sample` returns them.

The dense projection between the two hot blocks is plain cuDNN: `SimpleConv3D-v1m1`
(unet3d.py:16-34) restated with torch modules, run in channels_last_3d so that neither side needs a permute copy.
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence

import torch
from torch import nn

from . import densify
from .backbone import SpUNetBase
from .render import RayBundle, build_renderer


class SimpleConv3D(nn.Module):
    """Conv3d(k3,p1) + BatchNorm3d + ReLU — same module tree / state_dict as unet3d.py:16-34."""

    def __init__(self, in_channels, out_channels, kernel_size=3, padding=1, stride=1):
        super().__init__()
        self.conv = nn.Sequential(
            nn.Conv3d(in_channels, out_channels, kernel_size=kernel_size, padding=padding, stride=stride),
            nn.BatchNorm3d(out_channels), nn.ReLU(inplace=True))

    def forward(self, x):
        return self.conv(x)


class _SceneViews(torch.autograd.Function):
    """(B,C,Z,Y,X) volume -> B per-scene (C,Z,Y,X) views (the per-scene render loop of render_func,
    ponder_indoor_base.py:645-669).  `volume[i]`'s own backward zero-fills a *contiguous NCDHW* tensor per scene and
    copies the slice in, after which every elementwise backward of the channels-last projection net runs on mismatched
    layouts (measured: 1.3 ms per step on B200).  This backward assembles the gradient once, in the volume's own
    strides; with one scene it is a pure view."""

    @staticmethod
    def forward(ctx, volume):
        ctx.shape, ctx.strides = tuple(volume.shape), tuple(volume.stride())
        return tuple(volume[i] for i in range(volume.shape[0]))

    @staticmethod
    def backward(ctx, *grads):
        B = ctx.shape[0]
        if B == 1 and grads[0] is not None and tuple(grads[0].stride()) == ctx.strides[1:]:
            return grads[0].unsqueeze(0)
        ref = next(g for g in grads if g is not None)
        out = torch.empty_strided(ctx.shape, ctx.strides, dtype=ref.dtype, device=ref.device)
        for i, g in enumerate(grads):
            if g is None:
                out[i].zero_()
            else:
                out[i].copy_(g)
        return out


class PonderIndoorStep(nn.Module):
    def __init__(self, backbone: dict, renderer: dict, projection: Optional[dict] = None,
                 grid_shape: Sequence[int] = (128, 128, 32), grid_size: float = 0.02, pool_type: str = "mean"):
        super().__init__()
        if pool_type != "mean":
            raise NotImplementedError("pool_type: every shipped config uses 'mean'")
        bb = dict(backbone); bb.pop("type", None)
        self.backbone = SpUNetBase(**bb)
        proj = dict(projection or dict(in_channels=96, out_channels=128)); proj.pop("type", None)
        self.proj_net = SimpleConv3D(**proj).to(memory_format=torch.channels_last_3d)
        self.renderer = build_renderer(renderer)
        self.grid_shape = tuple(int(g) for g in grid_shape)
        self.grid_size = float(grid_size)

    def to_dense(self, data_dict) -> torch.Tensor:
        offset = data_dict["offset"]
        n = data_dict["coord"].shape[0]
        batch = torch.searchsorted(offset, torch.arange(n, device=offset.device), right=True)
        cell = densify.indoor_cells(data_dict["coord"], batch, data_dict["resolution"], self.grid_shape, self.grid_size)
        X, Y, Z = self.grid_shape
        return densify.scatter_mean_volume(data_dict["sparse_backbone_feat"], cell, int(offset.shape[0]), (Z, Y, X))

    def forward(self, data_dict: Dict[str, torch.Tensor], noise: Optional[dict] = None) -> Dict[str, torch.Tensor]:
        data_dict["sparse_backbone_feat"] = self.backbone(data_dict)
        return self.forward_after_backbone(data_dict, noise)

    def forward_after_backbone(self, data_dict: Dict[str, torch.Tensor], noise: Optional[dict] = None) -> Dict[str, torch.Tensor]:
        """densify -> projection -> render -> losses, given `sparse_backbone_feat` (extract_feature's output)."""
        volume = self.proj_net(self.to_dense(data_dict))            # (B,C,Z,Y,X), channels_last_3d
        outs = []
        scene_vols = _SceneViews.apply(volume)
        for i in range(data_dict["ray_o"].shape[0]):                # scenes are independent (render_func :645-669)
            rb = RayBundle(origins=data_dict["ray_o"][i], directions=data_dict["ray_d"][i])
            outs.append(self.renderer(rb, [scene_vols[i]], noise=noise))
        render_out = outs[0] if len(outs) == 1 else {k: torch.cat([o[k] for o in outs], dim=0) for k in outs[0]}
        loss_dict = self.renderer.get_loss(render_out, {"depth": data_dict["depth"], "rgb": data_dict.get("rgb")})
        loss = sum(v for k, v in loss_dict.items() if "loss" in k)
        return dict(loss=loss, **loss_dict)
