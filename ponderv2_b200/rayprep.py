"""Ray preparation of the indoor pretraining step on the device: the collate dict the
dataloader hands over (`coord`, `offset`, `rgb (B,V,H,W,3)`, `depth (B,V,H,W)`, `intrinsic`, `extrinsic (B,V,4,4)`,
`depth_scale`) -> unit-cube voxel coordinates, per-ray origins / directions / colour / point-to-point depth targets.

Host-side mirror of `PonderIndoor.to_unit_cube` / `get_rays` / `get_mask_at_box` / `ray_sample` / `grid_sample`
(ponder/models/ponder/ponder_indoor_base.py:344-633).  The reference loops over scenes and views in Python, builds the
rays of EVERY pixel of every view (full-image meshgrids, 4x4 inverses per view), indexes the sampled ones, and round-trips
through numpy for the box test.  Here everything is batched torch on the device, rays are computed for the sampled pixels
only, and the only data-dependent shape (the valid-pixel list) is replaced by a random-key top-k, so the function never
synchronises with the host.  Arithmetic follows the reference formula by formula (fp32); its 4x4 `torch.linalg.inv`
calls are replaced by the closed forms of the same rigid / similarity transforms (differences ~1e-7).
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F


def _segment_minmax(x: torch.Tensor, batch: torch.Tensor, nb: int) -> Tuple[torch.Tensor, torch.Tensor]:
    idx = batch[:, None].expand_as(x)
    lo = torch.full((nb, x.shape[1]), float("inf"), dtype=x.dtype, device=x.device).scatter_reduce(0, idx, x, "amin")
    hi = torch.full((nb, x.shape[1]), float("-inf"), dtype=x.dtype, device=x.device).scatter_reduce(0, idx, x, "amax")
    return lo, hi


@torch.no_grad()
def to_unit_cube(data_dict: Dict[str, torch.Tensor], z_level: float = -0.5) -> Dict[str, torch.Tensor]:
    """ponder_indoor_base.py:344-444.  Every scene is centred, scaled by 1 / (longest bbox edge) and put on the floor
    z = z_level; the camera poses are composed with the inverse transform; `depth_scale` absorbs the scale; `coord` is
    finally expressed in "grid resolution" units ((c + 0.5) * pc_scale) with its bbox in `bbox`.  Returns a new dict."""
    out = dict(data_dict)
    coord = data_dict["coord"].float()
    offset = data_dict["offset"]
    B = int(offset.shape[0])
    n = coord.shape[0]
    batch = torch.searchsorted(offset, torch.arange(n, device=coord.device), right=True)
    cmin, cmax = _segment_minmax(coord, batch, B)
    b0, b1 = cmin - 1e-5, cmax + 1e-5
    loc = (b0 + b1) / 2                                          # [B,3]
    extent = (b1 - b0).max(dim=1).values                          # [B]
    scale = 1.0 / extent
    tmp = (coord - loc[batch]) * scale[batch, None]
    zmin, _ = _segment_minmax(tmp[:, 2:3], batch, B)
    shift = torch.zeros_like(loc)
    shift[:, 2] = -zmin[:, 0] + z_level
    c = (coord - loc[batch]) * scale[batch, None] + shift[batch]  # S_loc2 @ S_scale @ S_loc applied to the points
    c = c.clamp(min=-0.5 + 1e-5, max=0.5 - 1e-5)
    # cameras: pose' = pose @ S^-1 with S x = scale (x - loc) + shift  ->  S^-1 x' = x' / scale + (loc - shift / scale)
    ext = data_dict["extrinsic"].float().clone()                  # [B,V,4,4]
    ext[:, :, 3, 3] = 1.0
    s_inv = torch.zeros((B, 4, 4), dtype=torch.float32, device=coord.device)
    s_inv[:, 0, 0] = s_inv[:, 1, 1] = s_inv[:, 2, 2] = extent
    s_inv[:, :3, 3] = loc - shift * extent[:, None]
    s_inv[:, 3, 3] = 1.0
    out["extrinsic"] = ext @ s_inv[:, None]
    out["depth_scale"] = scale * data_dict["depth_scale"].float()
    out["pc_scale"] = extent
    nmin, nmax = _segment_minmax(c, batch, B)
    bbox = torch.stack([nmin - 1e-5, nmax + 1e-5], dim=1)         # [B,2,3]
    out["bbox"] = (bbox + 0.5) * extent[:, None, None]
    out["coord"] = (c + 0.5) * extent[batch, None]
    return out


@torch.no_grad()
def grid_sample(data_dict: Dict[str, torch.Tensor], grid_size: float) -> Dict[str, torch.Tensor]:
    """ponder_indoor_base.py:622-627: bbox in voxels and `resolution` = longest edge in voxels + 1."""
    out = dict(data_dict)
    bb = torch.floor_divide(data_dict["bbox"], grid_size).int()
    out["bbox"] = bb
    out["resolution"] = (bb[:, 1] - bb[:, 0]).max(dim=1).values.int() + 1
    return out


@torch.no_grad()
def sample_pixels(depth: torch.Tensor, n: int, keys: Optional[torch.Tensor] = None) -> torch.Tensor:
    """A uniformly random subset of `n` valid (depth > 0) pixels per view, [B,V,n] flat indices y * W + x: the pixels
    with the n smallest random keys among the valid ones (the reference: `torch.where(mask)` + `randperm`, :558-561)."""
    B, V, H, W = depth.shape
    if keys is None:
        keys = torch.rand((B, V, H * W), device=depth.device)
    k = torch.where(depth.reshape(B, V, -1) > 0, keys, torch.full_like(keys, float("inf")))
    return torch.topk(k, n, dim=-1, largest=False).indices


@torch.no_grad()
def ray_sample(data_dict: Dict[str, torch.Tensor], ray_nsample: int, bounds: Sequence[Sequence[float]],
               pixels: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
    """ponder_indoor_base.py:446-620.  `data_dict` after `to_unit_cube`; with `semantic` (B,V,H,W class ids) and
    `index2semantic` ([K, E] text embeddings) in it, the per-ray embedding target of the semantic branch (§8f-4) is added.
    pixels: optional [B,V,n] flat pixel indices (tests inject the reference's choice).
    -> ray_o, ray_d [B, V*n, 3]; rgb [B*V*n, 3]; depth [B*V*n, 1] (point-to-point, -0.001 where the ray misses the box)."""
    rgb = data_dict["rgb"].float()
    dep = data_dict["depth"].float()
    B, V, H, W = dep.shape
    dev = dep.device
    if pixels is None:
        pixels = sample_pixels(dep, ray_nsample)
    n = pixels.shape[-1]
    py, px = torch.div(pixels, W, rounding_mode="floor"), pixels % W
    intr = data_dict["intrinsic"].float()
    K = intr[:, None, :3, :3].expand(B, V, 3, 3) if intr.dim() == 3 else intr[:, :, :3, :3]
    RT = data_dict["extrinsic"].float()                                   # world -> camera, [B,V,4,4]
    R, T = RT[:, :, :3, :3], RT[:, :, :3, 3]
    # inverse pose (get_rays :447-452): camera -> world rotation R^-1 and camera centre -R^-1 T
    Rinv = torch.linalg.inv(R)
    cam_o = -(Rinv @ T[..., None])[..., 0]                                 # [B,V,3]
    # tx = linspace(0, W - 1, W), ty = linspace(0, H - 1, H): pixel centres are the integer coordinates
    p = torch.stack([px.float(), py.float(), torch.ones_like(px, dtype=torch.float32)], dim=-1)     # [B,V,n,3]
    p = (torch.linalg.inv(K)[:, :, None] @ p[..., None])[..., 0]
    v = p / torch.linalg.norm(p, ord=2, dim=-1, keepdim=True)
    v = (Rinv[:, :, None] @ v[..., None])[..., 0]
    ray_d = F.normalize(v, dim=-1)
    ray_o = cam_o[:, :, None, :].expand(B, V, n, 3)
    flat = lambda t: torch.gather(t.reshape(B, V, H * W, -1), 2, pixels[..., None].expand(B, V, n, t.shape[-1]))
    color = flat(rgb)
    d = flat((dep * (dep > 0).float())[..., None])[..., 0] * data_dict["depth_scale"].float()[:, None, None]
    # plane-to-plane -> point-to-point depth (:573-579): divide by the cosine to the optical axis in world space
    plane = (Rinv @ torch.tensor([0.0, 0.0, 1.0], device=dev)[None, None, :, None])[..., 0]      # cam2lidar [0,0,1,1] - origin
    plane = plane / torch.linalg.norm(plane, dim=-1, keepdim=True)
    d = d / (ray_d * plane[:, :, None, :]).sum(-1)
    # get_mask_at_box (:481-498): slab test of the FIRST ray origin of the view against `bounds`, near clamped to 0.1
    vd = ray_d / torch.linalg.norm(ray_d, dim=-1, keepdim=True)
    vd = torch.where((vd < 1e-5) & (vd > -1e-10), torch.full_like(vd, 1e-5), vd)
    vd = torch.where((vd > -1e-5) & (vd < 1e-10), torch.full_like(vd, -1e-5), vd)
    inv = 1.0 / vd
    lo = torch.tensor(bounds[0], dtype=torch.float32, device=dev)
    hi = torch.tensor(bounds[1], dtype=torch.float32, device=dev)
    o1 = ray_o[:, :, :1, :]
    tmin, tmax = (lo - o1) * inv, (hi - o1) * inv
    near = torch.minimum(tmin, tmax).max(dim=-1).values.clamp(min=0.1)
    far = torch.maximum(tmin, tmax).min(dim=-1).values
    hit = near < far
    color = torch.where(hit[..., None], color, torch.zeros_like(color))
    d = torch.where(hit, d, torch.full_like(d, -0.001))
    out = dict(ray_o=ray_o.reshape(B, V * n, 3).contiguous(), ray_d=ray_d.reshape(B, V * n, 3).contiguous(),
               rgb=color.reshape(-1, color.shape[-1]).contiguous(), depth=d.reshape(-1, 1).contiguous())
    if "semantic" in data_dict and "index2semantic" in data_dict:
        # :581-597: class id of the sampled pixel (-1 outside the box), looked up in the text-embedding table; ids <= 0
        # (unlabelled / ignored) keep an all-zero target, which the loss masks out
        table = data_dict["index2semantic"].float()
        ids = torch.gather(data_dict["semantic"].reshape(B, V, H * W).long(), 2, pixels)
        ids = torch.where(hit.expand_as(ids), ids, torch.full_like(ids, -1))
        emb = table[ids.clamp(min=0, max=table.shape[0] - 1)] * (ids > 0)[..., None].float()
        out["semantic"] = emb.reshape(-1, table.shape[-1]).contiguous()
    return out
