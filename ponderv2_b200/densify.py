"""Densify: sparse backbone features -> dense channels-last volume (scatter-mean), differentiable.

Host-side mirror of `PonderIndoor.to_dense` (ponder_indoor_base.py:177-216,332-342, the pooling branch real scenes
take) and `PonderOutdoor.to_dense` (ponder_outdoor_base.py:178-210).  The cell index arithmetic reproduces the
reference's float floor-divisions exactly; the scatter-mean and its gradient are `pv2_densify_fwd/bwd`.

Layout: the reference returns a contiguous (B,C,Z,Y,X) tensor after a zero-fill + scatter + permute-copy.  Here the
kernel writes the mean straight into memory order [B][Z][Y][X][C] and the function returns the logical
(B,C,Z,Y,X) view of it, i.e. a `torch.channels_last_3d` tensor: equal values, no permute pass, and the layout both
cuDNN's Conv3d and the renderer's vectorised trilinear fetch prefer.
"""
from __future__ import annotations

from typing import Sequence

import torch

from . import _lib


class _DensifyFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feat, cell, cells_total: int):
        lib = _lib.load()
        feat = feat.contiguous().float()
        n, c = feat.shape
        volume = torch.empty((cells_total, c), dtype=torch.float32, device=feat.device)
        count = torch.empty(cells_total, dtype=torch.int32, device=feat.device)
        with _lib.on_device(feat.device):
            _lib.check(lib.pv2_densify_fwd(_lib.ptr(feat), _lib.ptr(cell), n, c, cells_total, _lib.ptr(volume),
                                           _lib.ptr(count), _lib.stream_ptr()), "pv2_densify_fwd")
        ctx.save_for_backward(cell, count)
        ctx.shape = (n, c)
        return volume

    @staticmethod
    def backward(ctx, dvolume):
        cell, count = ctx.saved_tensors
        n, c = ctx.shape
        lib = _lib.load()
        dvolume = dvolume.contiguous()
        dfeat = torch.empty((n, c), dtype=torch.float32, device=dvolume.device)
        with _lib.on_device(dvolume.device):
            _lib.check(lib.pv2_densify_bwd(_lib.ptr(dvolume), _lib.ptr(cell), _lib.ptr(count), n, c, count.shape[0],
                                           _lib.ptr(dfeat), _lib.stream_ptr()), "pv2_densify_bwd")
        return dfeat, None, None


def scatter_mean_volume(feat: torch.Tensor, cell: torch.Tensor, batch_size: int, zyx: Sequence[int]) -> torch.Tensor:
    """feat [N,C], cell [N] int64 index into [B][Z][Y][X] (or -1) -> (B,C,Z,Y,X) channels_last_3d tensor."""
    z, y, x = (int(v) for v in zyx)
    vol = _DensifyFunction.apply(feat, cell.contiguous(), batch_size * z * y * x)
    return vol.view(batch_size, z, y, x, feat.shape[1]).permute(0, 4, 1, 2, 3)


def indoor_cells(coord: torch.Tensor, batch: torch.Tensor, resolution: torch.Tensor, grid_shape: Sequence[int],
                 grid_size: float) -> torch.Tensor:
    """Cell id per voxel, pooling branch of ponder_indoor_base.py:199-213.
    coord [N,3] float (scene frame), batch [N] int64, resolution [B] (voxels along the longest bbox edge)."""
    gs = torch.tensor([float(g) for g in grid_shape], dtype=torch.float32, device=coord.device)
    if coord.is_cuda:
        # only the pooling branch of PonderIndoor.to_dense is implemented (every scene whose longest edge covers the grid,
        # ponder_indoor_base.py:199-216); smaller scenes take the reference's trilinear-upsample branch (:217-330).
        # Checked on the device without a host sync: a violation raises a CUDA device-side assertion.
        torch._assert_async((resolution.to(coord.device) + 1 >= min(int(g) for g in grid_shape)).all(),
                            "densify.indoor_cells: scene resolution below min(grid_shape): the reference would upsample")
    v = torch.floor_divide(coord, grid_size).int()                       # (coord // grid_size).int()
    cur = (resolution.to(coord.device)[batch] + 1).to(torch.int64)       # int(resolution + 1)
    scale = cur.to(torch.float32)[:, None] / gs[None, :]                 # current_resolution / FloatTensor(grid_shape)
    gi = torch.floor_divide(v, scale).long()
    gx, gy, gz = gi[:, 0], gi[:, 1], gi[:, 2]
    X, Y, Z = (int(g) for g in grid_shape)
    ok = (gx >= 0) & (gx < X) & (gy >= 0) & (gy < Y) & (gz >= 0) & (gz < Z)
    cell = ((batch * Z + gz) * Y + gy) * X + gx
    return torch.where(ok, cell, torch.full_like(cell, -1))


def outdoor_cells(coord: torch.Tensor, batch: torch.Tensor, scene_bbox: Sequence[float], grid_size: Sequence[float],
                  grid_shape: Sequence[int]) -> torch.Tensor:
    """ponder_outdoor_base.py:186-203."""
    bb = torch.tensor([float(b) for b in scene_bbox[:3]], dtype=coord.dtype, device=coord.device)
    gsz = torch.tensor([float(g) for g in grid_size], dtype=coord.dtype, device=coord.device)
    c = ((coord - bb) / gsz).long()
    X, Y, Z = (int(g) for g in grid_shape)
    ok = (c[:, 0] >= 0) & (c[:, 0] < X) & (c[:, 1] >= 0) & (c[:, 1] < Y) & (c[:, 2] >= 0) & (c[:, 2] < Z)
    cell = ((batch * Z + c[:, 2]) * Y + c[:, 1]) * X + c[:, 0]
    return torch.where(ok, cell, torch.full_like(cell, -1))
