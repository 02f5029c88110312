"""Twice-differentiable `SmoothSampler.apply(input, grid, padding_mode, align_corners, apply_smoothstep)`.

Same public behaviour as libs/smooth-sampler/smooth_sampler/modules.py:14-101 (two nested autograd Functions so
that `autograd.grad(..., create_graph=True)` followed by `backward()` works), minus the two host syncs the
reference performs (`(grad != 0).any().item()` at modules.py:45-47 and `.all().item()` at :90): a zero upstream
gradient simply flows through the kernels.
"""
from __future__ import annotations

import torch

from . import _C

_PADDING = {"zeros": 0, "border": 1, "reflection": 2}


def padding_mode_enum(padding_mode: str) -> int:
    try:
        return _PADDING[padding_mode]
    except KeyError:
        raise ValueError(f"unknown padding_mode {padding_mode!r}") from None


class SmoothSamplerBackward(torch.autograd.Function):
    @staticmethod
    def forward(ctx, input, grid, grad_out, padding_mode="zeros", align_corners=True, apply_smoothstep=False):
        ctx.align_corners = align_corners
        ctx.apply_smoothstep = apply_smoothstep
        ctx.padding_mode = padding_mode
        grad_input, grad_grid = _C.backward(grad_out, input, grid, padding_mode_enum(padding_mode), align_corners,
                                            apply_smoothstep, input.requires_grad)
        ctx.save_for_backward(input, grid, grad_out)
        return grad_input, grad_grid

    @staticmethod
    def backward(ctx, grad_out_input, grad_out_grid):
        input, grid, grad_out = ctx.saved_tensors
        has_goi = grad_out_input is not None
        if grad_out_grid is None:
            grad_out_grid = torch.zeros_like(grid)
        grad_input, grad_grid, grad_grad_out = _C.backward_backward(
            grad_out_input.contiguous() if has_goi else None, grad_out_grid.contiguous(), input, grid, grad_out,
            padding_mode_enum(ctx.padding_mode), ctx.align_corners, ctx.apply_smoothstep, has_goi)
        return grad_input, grad_grid, grad_grad_out, None, None, None


class SmoothSampler(torch.autograd.Function):
    @staticmethod
    def forward(ctx, input, grid, padding_mode="zeros", align_corners=True, apply_smoothstep=False):
        output = _C.forward(input, grid, padding_mode_enum(padding_mode), align_corners, apply_smoothstep)
        ctx.save_for_backward(input, grid)
        ctx.align_corners = align_corners
        ctx.apply_smoothstep = apply_smoothstep
        ctx.padding_mode = padding_mode
        return output

    @staticmethod
    def backward(ctx, grad_out):
        input, grid = ctx.saved_tensors
        d_input, d_grid = SmoothSamplerBackward.apply(input, grid, grad_out.contiguous(), ctx.padding_mode,
                                                      ctx.align_corners, ctx.apply_smoothstep)
        return d_input, d_grid, None, None, None
