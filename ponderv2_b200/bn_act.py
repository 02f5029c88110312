"""Fused BatchNorm1d (+ residual) (+ ReLU) for voxel features (csrc/bn_act.cu) — row a6 of the hot path.

`bn_act(x, bn, residual=None, relu=True)` applies the `nn.BatchNorm1d` module `bn` exactly as
`relu(bn(x) + residual)` would (reference: BasicBlock.forward, spconv_unet_v1m1_base.py:70-83, and the conv-bn-relu
blocks :111-180): batch statistics in training mode with the running buffers updated in place (momentum, unbiased
variance, num_batches_tracked), so state_dicts stay interchangeable with the reference's.  Two kernels forward, two
backward instead of torch's five and eight passes, for fp32 and bf16 features (statistics and parameters always fp32, as
nn.BatchNorm1d under autocast).  Eval mode (running statistics) is a plain affine map and stays torch, with a one-time note.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn.functional as F
from torch import nn

from . import _lib


class _BNActFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, res, gamma, beta, running_mean, running_var, momentum: float, eps: float, relu: bool):
        lib = _lib.load()
        x = x.contiguous()
        n, c = x.shape
        dev = x.device
        res_c = res.contiguous().to(x.dtype) if res is not None else None
        y = torch.empty_like(x)
        stats = torch.empty((2, c), dtype=torch.float32, device=dev)
        ws_bytes = lib.pv2_bn_workspace_bytes(n, c)
        with _lib.on_device(dev):
            ws = _lib.workspace(max(ws_bytes, 16), dev)
            _lib.check(lib.pv2_bn_act_fwd_t(_lib.ptr(x), _lib.ptr(res_c), _lib.ptr(gamma.contiguous()),
                                            _lib.ptr(beta.contiguous()), _lib.ptr(running_mean), _lib.ptr(running_var),
                                            float(momentum), float(eps), int(relu), n, c, _lib.ptr(y), _lib.ptr(stats[0]),
                                            _lib.ptr(stats[1]), _lib.dtype_code(x.dtype), _lib.ptr(ws), ws.numel(),
                                            _lib.stream_ptr()), "pv2_bn_act_fwd_t")
        ctx.save_for_backward(x, y, gamma, stats)
        ctx.relu, ctx.has_res = bool(relu), res is not None
        # parameters re-homed by dist.FlatParameters: dgamma / dbeta are accumulated by the kernel straight into their
        # slices of the flat gradient buffer (no temporaries, no autograd accumulation kernels)
        sg, sb = getattr(gamma, "_pv2_sink", None), getattr(beta, "_pv2_sink", None)
        ok = (sg is not None and sb is not None and sg[0] is sb[0] and sg[1].dtype == torch.float32
              and sb[1].dtype == torch.float32 and sg[1].is_contiguous() and sb[1].is_contiguous())
        ctx.sinks = (sg, sb, gamma, beta) if ok else None
        return y

    @staticmethod
    def backward(ctx, dy):
        x, y, gamma, stats = ctx.saved_tensors
        lib = _lib.load()
        n, c = x.shape
        dev = x.device
        dy = dy.contiguous().to(x.dtype)
        dx = torch.empty_like(x)
        dres = torch.empty_like(x) if ctx.has_res else None
        if ctx.sinks is not None:
            (flat, g_view), (_, b_view), p_gamma, p_beta = ctx.sinks
            dg_t, db_t, acc = g_view, b_view, 1
        else:
            dgb = torch.empty((2, c), dtype=torch.float32, device=dev)
            dg_t, db_t, acc = dgb[0], dgb[1], 0
        ws_bytes = lib.pv2_bn_workspace_bytes(n, c)
        with _lib.on_device(dev):
            ws = _lib.workspace(max(ws_bytes, 16), dev)
            _lib.check(lib.pv2_bn_act_bwd_t(_lib.ptr(x), _lib.ptr(dy), _lib.ptr(y), _lib.ptr(gamma.contiguous()),
                                            _lib.ptr(stats[0]), _lib.ptr(stats[1]), int(ctx.relu), n, c, _lib.ptr(dx),
                                            _lib.ptr(dres), _lib.ptr(dg_t), _lib.ptr(db_t), acc, _lib.dtype_code(x.dtype),
                                            _lib.ptr(ws), ws.numel(), _lib.stream_ptr()), "pv2_bn_act_bwd_t")
        if acc:
            flat.mark_ready(p_gamma); flat.mark_ready(p_beta)
            return dx, dres, None, None, None, None, None, None, None
        return dx, dres, dg_t, db_t, None, None, None, None, None


_NOTED = set()


def _note_fallback(x: torch.Tensor, bn: nn.BatchNorm1d) -> None:
    """One warning per reason when a call leaves the fused kernels (eval-mode statistics are a plain affine map and stay
    torch by design; anything else is a shape / dtype the kernels do not take)."""
    why = "eval mode (running statistics)" if not bn.training else f"dtype {x.dtype}, shape {tuple(x.shape)}"
    if why not in _NOTED:
        _NOTED.add(why)
        import warnings
        warnings.warn(f"bn_act: torch BatchNorm1d path taken: {why}", RuntimeWarning, stacklevel=3)


def supported(x: torch.Tensor, bn: nn.BatchNorm1d) -> bool:
    return (x.is_cuda and x.dtype in (torch.float32, torch.bfloat16) and x.dim() == 2 and x.shape[0] > 1
            and x.shape[1] % 4 == 0 and x.shape[1] <= 1024 and bn.affine and bn.weight.dtype == torch.float32)


def supported_modulated(x: torch.Tensor) -> bool:
    return (x.is_cuda and x.dtype in (torch.float32, torch.bfloat16) and x.dim() == 2 and x.shape[0] > 1
            and x.shape[1] % 4 == 0 and x.shape[1] <= 1024)


def bn_act_modulated(x: torch.Tensor, bn: nn.BatchNorm1d, scale: Optional[torch.Tensor], shift: Optional[torch.Tensor],
                     residual: Optional[torch.Tensor] = None, relu: bool = True) -> torch.Tensor:
    """relu(bn(x) * (1 + scale) + shift + residual) — prompt-driven normalisation (PDBatchNorm.forward,
    spconv_unet_v1m3_pdnorm.py:60-72) folded into the fused kernels: with per-channel `scale`, `shift` ([C] or [1, C],
    differentiable, produced by the context modulation) the map is a BatchNorm whose affine pair is
    gamma' = gamma (1 + scale), beta' = beta (1 + scale) + shift, so the same two kernels run and autograd carries
    dgamma' / dbeta' back into the modulation Linear (and into gamma / beta when `bn.affine`)."""
    c = x.shape[1]
    one_plus = (1.0 + scale.reshape(c).float()) if scale is not None else None
    if bn.affine:
        gamma = bn.weight * one_plus if one_plus is not None else bn.weight
        beta = bn.bias * one_plus if one_plus is not None else bn.bias
    else:
        gamma = one_plus if one_plus is not None else torch.ones(c, dtype=torch.float32, device=x.device)
        beta = torch.zeros(c, dtype=torch.float32, device=x.device)
    if shift is not None:
        beta = beta + shift.reshape(c).float()
    if bn.training and supported_modulated(x):
        momentum = bn.momentum
        if bn.track_running_stats and bn.num_batches_tracked is not None:
            bn.num_batches_tracked.add_(1)
            if momentum is None:
                momentum = 1.0 / float(bn.num_batches_tracked)
        rm = bn.running_mean if bn.track_running_stats else None
        rv = bn.running_var if bn.track_running_stats else None
        return _BNActFunction.apply(x, residual, gamma, beta, rm, rv, float(momentum or 0.0), float(bn.eps), relu)
    if x.is_cuda:
        _note_fallback(x, bn)
    out = F.batch_norm(x, bn.running_mean, bn.running_var, None, None, bn.training, bn.momentum or 0.0, bn.eps)
    out = out * gamma.to(out.dtype) + beta.to(out.dtype)
    if residual is not None:
        out = out + residual
    return F.relu(out) if relu else out


def bn_act(x: torch.Tensor, bn: nn.BatchNorm1d, residual: Optional[torch.Tensor] = None, relu: bool = True) -> torch.Tensor:
    """relu(bn(x) + residual) with `bn` an nn.BatchNorm1d (its parameters and running buffers are used / updated)."""
    if bn.training and supported(x, bn):
        momentum = bn.momentum
        if bn.track_running_stats and bn.num_batches_tracked is not None:
            bn.num_batches_tracked.add_(1)
            if momentum is None:  # cumulative moving average
                momentum = 1.0 / float(bn.num_batches_tracked)
        rm = bn.running_mean if bn.track_running_stats else None
        rv = bn.running_var if bn.track_running_stats else None
        return _BNActFunction.apply(x, residual, bn.weight, bn.bias, rm, rv, float(momentum or 0.0), float(bn.eps), relu)
    if x.is_cuda:
        _note_fallback(x, bn)
    out = bn(x)
    if residual is not None:
        out = out + residual
    return F.relu(out) if relu else out
