"""CPU restatement of the reference's densify step (TEST INFRASTRUCTURE, see oracle/__init__.py).

Indoor: PonderIndoor.to_dense pooling branch, /root/reference/ponder/models/ponder/ponder_indoor_base.py:177-216 +
the final view/permute :332-342.  Outdoor: PonderOutdoor.to_dense, ponder_outdoor_base.py:178-210.
torch_scatter.scatter(reduce="mean", out=zeros) is restated as index_add of sums and counts followed by
sum / max(count, 1) (torch_scatter's mean semantics).
"""
from __future__ import annotations

import numpy as np
import torch


def _scatter_mean(feat: torch.Tensor, index: torch.Tensor, size: int) -> torch.Tensor:
    out = torch.zeros(size, feat.shape[1], dtype=feat.dtype).index_add(0, index, feat)
    cnt = torch.zeros(size, dtype=feat.dtype).index_add(0, index, torch.ones(index.shape[0], dtype=feat.dtype))
    return out / cnt.clamp(min=1)[:, None]


def to_dense_indoor(coord, feat, offset, resolution, grid_shape, grid_size):
    """coord [N,3] float, feat [N,C], offset [B] cumulative, resolution [B] -> (B,C,Z,Y,X) contiguous."""
    B = len(offset)
    G = int(np.prod(grid_shape))
    fea = torch.zeros(B, G, feat.shape[1], dtype=feat.dtype)
    for i in range(B):
        lo = int(offset[i - 1]) if i else 0
        c, f = coord[lo:int(offset[i])], feat[lo:int(offset[i])]
        c = (c // grid_size).int()
        cur = int(resolution[i] + 1)
        assert cur >= min(grid_shape), "only the pooling branch is restated"
        gi = (c // (cur / torch.FloatTensor(list(grid_shape)))).long()
        idx = gi[:, 0] * grid_shape[1] * grid_shape[2] + gi[:, 1] * grid_shape[2] + gi[:, 2]
        fea[i] = _scatter_mean(f, idx, G)
    return fea.view(B, grid_shape[0], grid_shape[1], grid_shape[2], -1).permute(0, 4, 3, 2, 1).contiguous()


def to_dense_outdoor(coord, feat, offset, scene_bbox, grid_size, grid_shape):
    B = len(offset)
    counts = np.diff(np.concatenate([[0], np.asarray(offset)]))
    batch = torch.from_numpy(np.repeat(np.arange(B), counts))
    bb = torch.tensor(scene_bbox, dtype=coord.dtype)
    c = ((coord - bb[:3]) / torch.tensor(grid_size, dtype=coord.dtype)).long()
    gs = grid_shape
    idx = batch * gs[0] * gs[1] * gs[2] + c[:, 0] * gs[1] * gs[2] + c[:, 1] * gs[2] + c[:, 2]
    dense = _scatter_mean(feat, idx, B * gs[0] * gs[1] * gs[2])
    return dense.view(B, *gs, -1).permute(0, 4, 3, 2, 1).contiguous()
