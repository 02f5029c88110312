"""Host wrappers of the per-ray kernels (csrc/render_ray.cu): sampler, compositing and loss of the NeuS renderer.

Each function mirrors one block of the reference's torch code and cites it; the two differentiable ones are
`torch.autograd.Function`s whose backward is the hand-written kernel, not autograd through elementwise ops.
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import torch

from .. import _lib

MAX_S0, MAX_SI, MAX_S = 256, 127, 256   # kMaxS0, kMaxNb - 1, 32 * kMaxRounds in csrc/render_ray.cu


def supported(s0: int, si: int, steps: int) -> bool:
    return steps == 1 and 2 <= s0 <= MAX_S0 and 1 <= si <= MAX_SI and s0 + si <= MAX_S


def _f32c(t: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
    return None if t is None else t.contiguous().float()


def ray_setup(origins: torch.Tensor, directions: torch.Tensor, s0: int, bbox: Sequence[float], near_plane: float,
              noise: Optional[torch.Tensor]):
    """AABBBoxCollider (scene_colliders.py:38-99) + UniformSampler (ray_samplers.py:55-107).
    -> nears [R,1], fars [R,1], spacing bins [R,S0+1], coarse points [R,S0,3] (all detached)."""
    lib = _lib.load()
    o, d = _f32c(origins.detach()), _f32c(directions.detach())
    R, dev = o.shape[0], o.device
    nears = torch.empty((R, 1), dtype=torch.float32, device=dev)
    fars = torch.empty((R, 1), dtype=torch.float32, device=dev)
    bins = torch.empty((R, s0 + 1), dtype=torch.float32, device=dev)
    pts = torch.empty((R, s0, 3), dtype=torch.float32, device=dev)
    nz = _f32c(noise)
    bb = (_lib.C.c_float * 6)(*[float(b) for b in bbox])
    with _lib.on_device(dev):
        _lib.check(lib.pv2_ray_setup(_lib.ptr(o), _lib.ptr(d), _lib.ptr(nz), nz.shape[1] if nz is not None else 0, R, s0,
                                     bb, float(near_plane), _lib.ptr(nears), _lib.ptr(fars), _lib.ptr(bins),
                                     _lib.ptr(pts), _lib.stream_ptr()), "pv2_ray_setup")
    return nears, fars, bins, pts


def ray_resample(origins, directions, nears, fars, bins, sdf_coarse, si: int, inv_s: float, noise, norm_pts: bool,
                 norm_padding: float):
    """NeuSSampler (one upsample step, ray_samplers.py:355-463) + PDFSampler (:227-322) + merge (rays.py:118-153) +
    normalize_3d_coordinate (sdf_field.py:58-74).
    -> starts [R,S], deltas [R,S], normalised points [R,S,3], init_weights [R,S0], new spacing bins [R,Si], minmax [2]."""
    lib = _lib.load()
    o, d = _f32c(origins.detach()), _f32c(directions.detach())
    R, dev = o.shape[0], o.device
    s0 = bins.shape[1] - 1
    S = s0 + si
    sdf = _f32c(sdf_coarse.detach()).view(R, s0)
    starts = torch.empty((R, S), dtype=torch.float32, device=dev)
    deltas = torch.empty((R, S), dtype=torch.float32, device=dev)
    pts = torch.empty((R, S, 3), dtype=torch.float32, device=dev)
    init_w = torch.empty((R, s0), dtype=torch.float32, device=dev)
    new_bins = torch.empty((R, si), dtype=torch.float32, device=dev)
    minmax = torch.empty(2, dtype=torch.int32, device=dev)
    nz = _f32c(noise)
    with _lib.on_device(dev):
        _lib.check(lib.pv2_ray_resample(_lib.ptr(o), _lib.ptr(d), _lib.ptr(nears), _lib.ptr(fars), _lib.ptr(bins),
                                        _lib.ptr(sdf), _lib.ptr(nz), nz.shape[1] if nz is not None else 0, R, s0, si,
                                        float(inv_s), int(bool(norm_pts)), float(norm_padding), _lib.ptr(starts),
                                        _lib.ptr(deltas), _lib.ptr(pts), _lib.ptr(init_w), _lib.ptr(new_bins),
                                        _lib.ptr(minmax), _lib.stream_ptr()), "pv2_ray_resample")
    return starts, deltas, pts, init_w, new_bins, minmax


class RayComposite(torch.autograd.Function):
    """(sdf [R,S], grad [R,S,3], rgbs [R,S,3] | None, variance [1]) -> weights [R,S], rgb [R,3] | None, depth [R],
    normal [R,3].  get_alpha (sdf_field.py:122-146) + weights (rays.py:83-105) + renderers (renderers.py:5-75)."""

    @staticmethod
    def forward(ctx, sdf, grad, rgbs, variance, starts, deltas, dirs, minmax, cos_anneal: float, clamp_rgb: bool):
        lib = _lib.load()
        sdf, grad, rgbs = _f32c(sdf), _f32c(grad), _f32c(rgbs)
        starts, deltas, dirs = _f32c(starts), _f32c(deltas), _f32c(dirs)
        var = _f32c(variance)
        R, S = sdf.shape
        dev = sdf.device
        weights = torch.empty((R, S), dtype=torch.float32, device=dev)
        depth = torch.empty(R, dtype=torch.float32, device=dev)
        normal = torch.empty((R, 3), dtype=torch.float32, device=dev)
        rgb = torch.empty((R, 3), dtype=torch.float32, device=dev) if rgbs is not None else None
        with _lib.on_device(dev):
            _lib.check(lib.pv2_ray_composite_fwd(_lib.ptr(sdf), _lib.ptr(grad), _lib.ptr(rgbs), _lib.ptr(starts),
                                                 _lib.ptr(deltas), _lib.ptr(dirs), _lib.ptr(var), _lib.ptr(minmax),
                                                 float(cos_anneal), R, S, int(bool(clamp_rgb)), _lib.ptr(weights),
                                                 _lib.ptr(rgb), _lib.ptr(depth), _lib.ptr(normal), _lib.stream_ptr()),
                       "pv2_ray_composite_fwd")
        ctx.save_for_backward(sdf, grad, rgbs if rgbs is not None else sdf.new_empty(0), var, starts, deltas, dirs, minmax)
        ctx.has_rgb = rgbs is not None
        ctx.cos_anneal = float(cos_anneal)
        return weights, rgb, depth, normal

    @staticmethod
    def backward(ctx, g_weights, g_rgb, g_depth, g_normal):
        sdf, grad, rgbs, var, starts, deltas, dirs, minmax = ctx.saved_tensors
        lib = _lib.load()
        R, S = sdf.shape
        dev = sdf.device
        rgbs_p = rgbs if ctx.has_rgb else None
        g_sdf = torch.empty((R, S), dtype=torch.float32, device=dev)
        g_grad = torch.empty((R, S, 3), dtype=torch.float32, device=dev)
        g_rgbs = torch.empty((R, S, 3), dtype=torch.float32, device=dev) if ctx.has_rgb else None
        g_var = torch.zeros(1, dtype=torch.float32, device=dev)
        if not ctx.has_rgb:
            g_rgb = None
        with _lib.on_device(dev):
            _lib.check(lib.pv2_ray_composite_bwd(_lib.ptr(sdf), _lib.ptr(grad), _lib.ptr(rgbs_p), _lib.ptr(starts),
                                                 _lib.ptr(deltas), _lib.ptr(dirs), _lib.ptr(var), _lib.ptr(minmax),
                                                 ctx.cos_anneal, R, S, _lib.ptr(_f32c(g_rgb)), _lib.ptr(_f32c(g_depth)),
                                                 _lib.ptr(_f32c(g_normal)), _lib.ptr(_f32c(g_weights)), _lib.ptr(g_sdf),
                                                 _lib.ptr(g_grad), _lib.ptr(g_rgbs), _lib.ptr(g_var), _lib.stream_ptr()),
                       "pv2_ray_composite_bwd")
        return g_sdf, g_grad, g_rgbs, g_var.view_as(var), None, None, None, None, None, None


class RayLoss(torch.autograd.Function):
    """SurfaceModel.get_loss (base_surface_model.py:102-211) -> terms [6] = weighted (depth, rgb, free-space, sdf,
    eikonal) losses and the rgb mean squared error (for psnr).  `wvec` [5] holds the loss weights (0 = term disabled),
    `const` [5] = (1, 3R, 1, 1, R*S): the fixed normalisers; counted normalisers are clamped to >= 1 like the reference."""

    @staticmethod
    def forward(ctx, depth_pred, rgb_pred, sdf, grad, z, depth_gt, rgb_gt, trunc: float, wvec, const):
        lib = _lib.load()
        dp, rp = _f32c(depth_pred).view(-1), _f32c(rgb_pred)
        sdf, grad, z = _f32c(sdf), _f32c(grad), _f32c(z)
        dg, rg = _f32c(depth_gt).view(-1), _f32c(rgb_gt)
        R, S = sdf.shape
        dev = sdf.device
        sums = torch.empty(11, dtype=torch.float32, device=dev)
        with _lib.on_device(dev):
            _lib.check(lib.pv2_ray_loss_fwd(_lib.ptr(dp), _lib.ptr(rp), _lib.ptr(dg), _lib.ptr(rg), _lib.ptr(sdf),
                                            _lib.ptr(z), _lib.ptr(grad), R, S, float(trunc), _lib.ptr(sums),
                                            _lib.stream_ptr()), "pv2_ray_loss_fwd")
        norm = torch.maximum(sums[5:10], const)
        scale = wvec / norm
        terms = torch.cat([sums[0:5] * scale, sums[10:11] / const[1:2]])
        ctx.save_for_backward(dp, rp if rp is not None else dp.new_empty(0), dg, rg if rg is not None else dp.new_empty(0),
                              sdf, z, grad, scale)
        ctx.has_rgb = rp is not None
        ctx.trunc = float(trunc)
        ctx.depth_shape = depth_pred.shape
        return terms

    @staticmethod
    def backward(ctx, g_terms):
        dp, rp, dg, rg, sdf, z, grad, scale = ctx.saved_tensors
        lib = _lib.load()
        R, S = sdf.shape
        dev = sdf.device
        coef = (g_terms[0:5].float() * scale).contiguous()
        g_depth = torch.empty(R, dtype=torch.float32, device=dev)
        g_rgb = torch.empty((R, 3), dtype=torch.float32, device=dev) if ctx.has_rgb else None
        g_sdf = torch.empty((R, S), dtype=torch.float32, device=dev)
        g_grad = torch.empty((R, S, 3), dtype=torch.float32, device=dev)
        with _lib.on_device(dev):
            _lib.check(lib.pv2_ray_loss_bwd(_lib.ptr(dp), _lib.ptr(rp if ctx.has_rgb else None), _lib.ptr(dg),
                                            _lib.ptr(rg if ctx.has_rgb else None), _lib.ptr(sdf), _lib.ptr(z),
                                            _lib.ptr(grad), R, S, ctx.trunc, _lib.ptr(coef), _lib.ptr(g_depth),
                                            _lib.ptr(g_rgb), _lib.ptr(g_sdf), _lib.ptr(g_grad), _lib.stream_ptr()),
                       "pv2_ray_loss_bwd")
        return g_depth.view(ctx.depth_shape), g_rgb, g_sdf, g_grad, None, None, None, None, None, None
