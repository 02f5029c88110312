"""Pure-torch, twice-differentiable restatement of the reference sampler (TEST INFRASTRUCTURE, see oracle/__init__.py).

Follows libs/smooth-sampler/smooth_sampler/csrc/smooth_sampler_kernel.cu:73-152 (forward: unnormalise, floor,
8 corner weights, bounds-checked accumulation; optional smoothstep :27-37,93-97) whose coordinate transform is
torch's grid_sample (`grid_sampler_compute_source_index`: align_corners unnormalise `((g+1)/2)*(size-1)`, border
clip, reflection).  Because it is composed of differentiable torch ops on CPU, autograd provides the first and
second derivatives the CUDA kernels (:155-619) compute by hand.
"""
from __future__ import annotations

import torch


def _unnormalize(g, size: int, align: bool):
    if align:
        return ((g + 1.0) / 2.0) * (size - 1)
    return ((g + 1.0) * size - 1.0) / 2.0


def _reflect(x, twice_low: int, twice_high: int):
    if twice_low == twice_high:
        return torch.zeros_like(x)
    mn = twice_low / 2.0
    span = (twice_high - twice_low) / 2.0
    x = (x - mn).abs()
    extra = torch.fmod(x, span)
    flips = torch.floor(x / span)
    even = (torch.remainder(flips, 2.0) == 0)
    return torch.where(even, extra + mn, span - extra + mn)


def _source_index(g, size: int, padding_mode: str, align: bool):
    x = _unnormalize(g, size, align)
    if padding_mode == "border":
        x = x.clamp(0, size - 1)
    elif padding_mode == "reflection":
        x = _reflect(x, 0, 2 * (size - 1)) if align else _reflect(x, -1, 2 * size - 1)
        x = x.clamp(0, size - 1)
    return x


def trilinear_sample(input: torch.Tensor, grid: torch.Tensor, padding_mode: str = "zeros",
                     align_corners: bool = True, apply_smoothstep: bool = False) -> torch.Tensor:
    """input (N,C,D,H,W), grid (N,Do,Ho,Wo,3) -> (N,C,Do,Ho,Wo); differentiable to any order."""
    N, C, D, H, W = input.shape
    out_sp = grid.shape[1:4]
    g = grid.reshape(N, -1, 3)
    P = g.shape[1]
    coords = []
    for comp, size in ((0, W), (1, H), (2, D)):
        x = _source_index(g[..., comp], size, padding_mode, align_corners)
        i0 = torch.floor(x).detach()
        t = x - i0
        if apply_smoothstep:
            t = t * t * (3.0 - 2.0 * t)
        coords.append((i0.long(), t, size))
    (ix0, tx, _), (iy0, ty, _), (iz0, tz, _) = coords
    flat = input.reshape(N, C, D * H * W)
    out = torch.zeros(N, C, P, dtype=input.dtype, device=input.device)
    for s in range(8):
        px, py, pz = s & 1, (s >> 1) & 1, (s >> 2) & 1
        ix, iy, iz = ix0 + px, iy0 + py, iz0 + pz
        wx = tx if px else 1.0 - tx
        wy = ty if py else 1.0 - ty
        wz = tz if pz else 1.0 - tz
        inb = (ix >= 0) & (ix < W) & (iy >= 0) & (iy < H) & (iz >= 0) & (iz < D)
        lin = ((iz.clamp(0, D - 1) * H + iy.clamp(0, H - 1)) * W + ix.clamp(0, W - 1))  # (N,P)
        vals = torch.gather(flat, 2, lin.unsqueeze(1).expand(N, C, P))
        wgt = (wx * wy * wz) * inb.to(input.dtype)
        out = out + vals * wgt.unsqueeze(1)
    return out.reshape(N, C, *out_sp)
