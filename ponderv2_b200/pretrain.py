"""One pretraining step of the hot path: SpUNet backbone -> densify -> dense projection -> NeuS renderer -> losses.
`PonderIndoorStep` mirrors PonderIndoor (below), `PonderOutdoorStep` mirrors PonderOutdoor
(ponder_outdoor_base.py:18-265: masking :93-139, prepare_ray :141-176, to_dense :178-210, render_func :218-251).

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


def _completion_order(step: nn.Module) -> list:
    """Parameters in the order backward finishes their gradients (first = earliest): renderer, projection, then the
    backbone from its last executed block to its first (SpUNetBase.forward: conv_input, (down, enc) x 4, (up, dec) x 4
    from the deepest stage up).  FlatParameters lays the gradient buffer out in this order so that contiguous slices
    complete early and can be all-reduced while backward is still running."""
    bb = step.backbone
    mods = [step.renderer, step.proj_net, bb.final]
    for s in range(bb.num_stages):
        mods += [bb.dec[s], bb.up[s]]
    for s in reversed(range(bb.num_stages)):
        mods += [bb.enc[s], bb.down[s]]
    mods.append(bb.conv_input)
    out, seen = [], set()
    for m in mods:
        for p in reversed(list(m.parameters())):
            if id(p) not in seen:
                seen.add(id(p)); out.append(p)
    for p in step.parameters():            # anything not covered above (mtoken, ...)
        if id(p) not in seen:
            seen.add(id(p)); out.append(p)
    return out


class PonderIndoorStep(nn.Module):
    def __init__(self, backbone: dict, renderer: dict, projection: Optional[dict] = None,
                 grid_shape: Sequence[int] = (128, 128, 32), grid_size: float = 0.02, pool_type: str = "mean",
                 val_ray_split: int = 10240):
        super().__init__()
        self.val_ray_split = int(val_ray_split)
        if pool_type != "mean":
            raise NotImplementedError("pool_type: every shipped config uses 'mean'")
        bb = dict(backbone); bb.pop("type", None)
        self.backbone = SpUNetBase(**bb)
        proj = dict(projection or dict(in_channels=96, out_channels=128))
        if proj.pop("type", "SimpleConv3D-v1m1") == "UNet3D-v1m2":   # the ScanNet / S3DIS configs' projection (§8f-1)
            from .models import UNet3Dv1m2
            self.proj_net = UNet3Dv1m2(**proj).to(memory_format=torch.channels_last_3d)
        else:
            self.proj_net = SimpleConv3D(**proj).to(memory_format=torch.channels_last_3d)
        self.renderer = build_renderer(renderer)
        self.grid_shape = tuple(int(g) for g in grid_shape)
        self.grid_size = float(grid_size)

    def grad_completion_order(self) -> list:
        return _completion_order(self)

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
            o, d = data_dict["ray_o"][i], data_dict["ray_d"][i]
            if self.training:
                outs.append(self.renderer(RayBundle(origins=o, directions=d), [scene_vols[i]], noise=noise))
            else:   # eval: rays in chunks of val_ray_split, outputs detached (ponder_indoor_base.py:655-668)
                parts = [self.renderer(RayBundle(origins=oo, directions=dd), [scene_vols[i]], noise=noise)
                         for oo, dd in zip(o.split(self.val_ray_split), d.split(self.val_ray_split))]
                outs.append({k: torch.cat([q[k].detach() for q in parts], 0) for k in parts[0]})
        render_out = outs[0] if len(outs) == 1 else {k: torch.cat([o[k] for o in outs], dim=0) for k in outs[0]}
        targets = {"depth": data_dict["depth"], "rgb": data_dict.get("rgb")}
        if "semantic" in data_dict and data_dict["semantic"].is_floating_point():   # per-ray text embeddings (§8f-4)
            targets["semantic"] = data_dict["semantic"]
        loss_dict = self.renderer.get_loss(render_out, targets)
        loss = sum(v for k, v in loss_dict.items() if "loss" in k)
        return dict(loss=loss, **loss_dict)


class PonderOutdoorStep(nn.Module):
    """`PonderOutdoor.forward` (ponder_outdoor_base.py:258-265) on the B200 kernels, single-dataset form (the shipped
    nuScenes / Waymo / SemanticKITTI base configs list one condition each; `scene_bbox`, `grid_shape`, `grid_size` are
    that condition's entries).  data_dict: grid_coord [N,3] int64, coord [N,3] f32 (metres, sensor frame), feat [N,4],
    offset [B], ray_start / ray_end [sum R,3], ray_offset [B] (cumulative)."""

    def __init__(self, backbone: dict, renderer: dict, projection: Optional[dict] = None, mask: Optional[dict] = None,
                 scene_bbox: Sequence[float] = (-54.0, -54.0, -5.0, 54.0, 54.0, 3.0),
                 grid_shape: Sequence[int] = (180, 180, 5), grid_size: Sequence[float] = (0.6, 0.6, 1.6),
                 val_ray_split: int = 8192, pool_type: str = "mean"):
        super().__init__()
        if pool_type != "mean":
            raise NotImplementedError("pool_type: every shipped config uses 'mean'")
        bb = dict(backbone); bb.pop("type", None)
        self.backbone = SpUNetBase(**bb)
        proj = dict(projection or dict(in_channels=96, out_channels=32)); proj.pop("type", None)
        self.proj_net = SimpleConv3D(**proj).to(memory_format=torch.channels_last_3d)
        self.renderer = build_renderer(renderer)
        self.scene_bbox = tuple(float(v) for v in scene_bbox)
        self.grid_shape = tuple(int(g) for g in grid_shape)
        self.grid_size = tuple(float(g) for g in grid_size)
        self.val_ray_split = int(val_ray_split)
        self.mask = dict(mask) if mask is not None else None
        if self.mask is not None:
            tok = nn.Parameter(torch.zeros(1, int(self.mask["channel"])))
            nn.init.trunc_normal_(tok, mean=0.0, std=0.02, a=-0.02, b=0.02)
            self.register_parameter("mtoken", tok)

    def grad_completion_order(self) -> list:
        return _completion_order(self)

    # -- extract_feature (:93-139): random block masking of the input features, then the backbone
    def mask_features(self, grid_coord, feat, offset, noise: Optional[torch.Tensor] = None):
        """Blocks of `mask.size`^3 voxels; per scene a uniformly random subset of round(n_blocks * (1 - ratio)) blocks is
        kept, every voxel of the other blocks gets the learned `mtoken`.  `noise` [n_blocks] replaces the random keys
        (tests).  One host sync (torch.unique's output size), against one per scene in the reference."""
        n = grid_coord.shape[0]
        batch = torch.searchsorted(offset, torch.arange(n, device=offset.device), right=True)
        blk = torch.cat([batch[:, None], torch.div(grid_coord, int(self.mask["size"]), rounding_mode="floor")], -1)
        ublk, inv = blk.unique(return_inverse=True, dim=0)
        nb = ublk.shape[0]
        scene = ublk[:, 0]
        key = noise if noise is not None else torch.rand(nb, device=feat.device)
        order = torch.argsort(scene.double() * 2.0 + key.double())         # scenes stay together, random inside
        rank = torch.empty_like(order)
        rank[order] = torch.arange(nb, device=order.device)
        per_scene = torch.bincount(scene, minlength=int(offset.shape[0]))
        start = torch.cumsum(per_scene, 0) - per_scene
        keep_n = torch.round(per_scene.double() * (1.0 - float(self.mask["ratio"]))).long()
        keep = (rank - start[scene]) < keep_n[scene]
        voxel_keep = keep[inv]
        return torch.where(voxel_keep[:, None], feat, self.mtoken.to(feat.dtype).expand_as(feat))

    # -- prepare_ray (:141-176)
    @torch.no_grad()
    def prepare_ray(self, data_dict):
        bb = torch.tensor(self.scene_bbox, dtype=data_dict["ray_start"].dtype, device=data_dict["ray_start"].device)
        norm = lambda c: (c - bb[:3]) / (bb[3:] - bb[:3])
        o, e = norm(data_dict["ray_start"]), norm(data_dict["ray_end"])
        rd = dict(ray_offset=data_dict["ray_offset"], ray_o=o, ray_d=torch.nn.functional.normalize(e - o, dim=-1),
                  depth=torch.linalg.norm(e - o, dim=-1, keepdim=True))
        if "ray_color" in data_dict:
            rd["rgb"] = data_dict["ray_color"]
        return rd

    # -- to_dense (:178-210)
    def to_dense(self, data_dict) -> torch.Tensor:
        offset = data_dict["offset"]
        n = data_dict["coord"].shape[0]
        batch = torch.searchsorted(offset, torch.arange(n, device=offset.device), right=True)
        cell = densify.outdoor_cells(data_dict["coord"], batch, self.scene_bbox, self.grid_size, self.grid_shape)
        X, Y, Z = self.grid_shape
        return densify.scatter_mean_volume(data_dict["sparse_backbone_feat"], cell, int(offset.shape[0]), (Z, Y, X))

    def forward(self, data_dict: Dict[str, torch.Tensor], noise: Optional[dict] = None) -> Dict[str, torch.Tensor]:
        noise = noise or {}
        if self.mask is not None:
            data_dict = dict(data_dict)
            data_dict["feat"] = self.mask_features(data_dict["grid_coord"], data_dict["feat"], data_dict["offset"],
                                                   noise.get("mask"))
        data_dict["sparse_backbone_feat"] = self.backbone(data_dict)
        ray = self.prepare_ray(data_dict)
        volume = self.proj_net(self.to_dense(data_dict))
        scene_vols = _SceneViews.apply(volume)
        ro = ray["ray_offset"]
        bounds = [0] + [int(v) for v in (ro.tolist() if torch.is_tensor(ro) else ro)]
        outs = []
        for i in range(len(bounds) - 1):                               # render_func (:218-251)
            o, d = ray["ray_o"][bounds[i]:bounds[i + 1]], ray["ray_d"][bounds[i]:bounds[i + 1]]
            if self.training:
                outs.append(self.renderer(RayBundle(origins=o, directions=d), [scene_vols[i]], noise=noise))
            else:
                parts = [self.renderer(RayBundle(origins=oo, directions=dd), [scene_vols[i]], noise=noise)
                         for oo, dd in zip(o.split(self.val_ray_split), d.split(self.val_ray_split))]
                outs.append({k: torch.cat([q[k].detach() for q in parts], 0) for k in parts[0]})
        render_out = outs[0] if len(outs) == 1 else {k: torch.cat([q[k] for q in outs], dim=0) for k in outs[0]}
        loss_dict = self.renderer.get_loss(render_out, {"depth": ray["depth"], "rgb": ray.get("rgb")})
        loss = sum(v for k, v in loss_dict.items() if "loss" in k)
        return dict(loss=loss, **loss_dict)
