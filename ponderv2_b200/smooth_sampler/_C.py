"""`smooth_sampler._C` equivalent: forward / backward / backward_backward with the argument order, checks and
allocation behaviour of libs/smooth-sampler/smooth_sampler/csrc/smooth_sampler.cpp:36-97, on top of the C ABI
(`pv2_trilinear_*`).  Inputs must be CUDA and contiguous (the reference's CHECK_INPUT) or a RuntimeError is raised."""
from __future__ import annotations

import torch

from .. import _lib


def _check_input(t: torch.Tensor, name: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor")
    if not t.is_contiguous():
        raise RuntimeError(f"{name} must be contiguous")


def _dims(input: torch.Tensor, grid: torch.Tensor):
    if input.dim() != 5 or grid.dim() != 5 or grid.shape[-1] != 3 or grid.shape[0] != input.shape[0]:
        raise RuntimeError("expected input (N,C,D,H,W) and grid (N,Do,Ho,Wo,3)")
    if input.dtype != grid.dtype:
        raise RuntimeError("input and grid must have the same dtype")
    N, C, D, H, W = input.shape
    P = grid.shape[1] * grid.shape[2] * grid.shape[3]
    return N, C, D, H, W, P


def forward(input, grid, padding_mode: int, align_corners: bool, apply_smoothstep: bool):
    _check_input(input, "input")
    _check_input(grid, "grid")
    N, C, D, H, W, P = _dims(input, grid)
    out = torch.empty((N, C, grid.shape[1], grid.shape[2], grid.shape[3]), dtype=input.dtype, device=input.device)
    lib = _lib.load()
    with _lib.on_device(input.device):
        _lib.check(lib.pv2_trilinear_fwd(_lib.ptr(input), _lib.ptr(grid), _lib.ptr(out), N, C, D, H, W, P,
                                         int(padding_mode), int(align_corners), int(apply_smoothstep),
                                         _lib.dtype_code(input.dtype), _lib.stream_ptr()), "pv2_trilinear_fwd")
    return out


def backward(grad_output, input, grid, padding_mode: int, align_corners: bool, apply_smoothstep: bool,
             input_requires_grad: bool):
    _check_input(grad_output, "grad_output")
    _check_input(input, "input")
    _check_input(grid, "grid")
    N, C, D, H, W, P = _dims(input, grid)
    grad_input = torch.zeros_like(input) if input_requires_grad else None
    grad_grid = torch.empty_like(grid)
    lib = _lib.load()
    with _lib.on_device(input.device):
        _lib.check(lib.pv2_trilinear_bwd(_lib.ptr(grad_output), _lib.ptr(input), _lib.ptr(grid), _lib.ptr(grad_input),
                                         _lib.ptr(grad_grid), N, C, D, H, W, P, int(padding_mode), int(align_corners),
                                         int(apply_smoothstep), _lib.dtype_code(input.dtype), _lib.stream_ptr()),
                   "pv2_trilinear_bwd")
    return grad_input, grad_grid


def backward_backward(grad_out_input, grad_out_grid, input, grid, grad_output, padding_mode: int,
                      align_corners: bool, apply_smoothstep: bool, input_requires_grad: bool):
    _check_input(grad_out_grid, "grad_out_grid")
    _check_input(input, "input")
    _check_input(grid, "grid")
    _check_input(grad_output, "grad_output")
    if input_requires_grad:
        _check_input(grad_out_input, "grad_out_input")
    N, C, D, H, W, P = _dims(input, grid)
    grad_input = torch.zeros_like(input)
    grad_grid = torch.empty_like(grid)
    grad_grad_out = torch.zeros_like(grad_output)
    lib = _lib.load()
    with _lib.on_device(input.device):
        _lib.check(lib.pv2_trilinear_bwd_bwd(_lib.ptr(grad_out_input) if input_requires_grad else None,
                                             _lib.ptr(grad_out_grid), _lib.ptr(input), _lib.ptr(grid),
                                             _lib.ptr(grad_output), _lib.ptr(grad_input), _lib.ptr(grad_grid),
                                             _lib.ptr(grad_grad_out), N, C, D, H, W, P, int(padding_mode),
                                             int(align_corners), int(apply_smoothstep),
                                             _lib.dtype_code(input.dtype), _lib.stream_ptr()),
                   "pv2_trilinear_bwd_bwd")
    return grad_input, grad_grid, grad_grad_out
