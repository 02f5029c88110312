"""Fused SDF decoder for the outdoor renderer configuration (csrc/render_mlp.cu): `SDFDecoder` with hidden_size 16, in_dim 32
(decoders.py:6-36; configs/nuscenes/pretrain-ponder-spunet-v1m1-0-base.py:36-41) as one forward and one backward kernel,
returning sdf together with u = d sdf / d f and v = d sdf / d p so that the field's `autograd.grad(sdf, points,
create_graph=True)` (fields/sdf_field.py:226-238) becomes  (d f / d points)^T u + v  with the sampler's own backward.

`pack(decoder)` concatenates the decoder's parameters (differentiably: gradients reach the original nn.Linear parameters
through torch.cat's backward) in the layout the kernels read.  The parameter gradients are contractions over the points of
per-layer vectors the backward kernel emits; they run on the tensor-core weight-gradient kernel (`pv2_spconv_wgrad` with the
identity map) — see the derivation in csrc/render_mlp.cu and tests/test_host_cpu.py::test_sdf_mlp_adjoint_math.
"""
from __future__ import annotations

from typing import Tuple

import torch

from .. import _lib

H, F = 16, 32


def eligible(field) -> bool:
    sd = field.sdf_decoder
    return (field.rgb_decoder is None and field.semantic_decoder is None and field.use_gradient
            and sd.fc_p.weight.shape == (H, 3) and sd.fc_c[0].weight.shape == (H, F) and 2 <= sd.num_layers - 1 <= 8
            and all(getattr(sd, f"lin{l}").weight.shape[1] == H for l in range(sd.num_layers - 1)))


def pack(decoder) -> Tuple[torch.Tensor, int, int, float]:
    """-> (packed [n] fp32, L, O, points_factor); layout = pv2_sdf_mlp_param_count's."""
    L = decoder.num_layers - 1
    lins = [getattr(decoder, f"lin{l}") for l in range(L)]
    parts = [decoder.fc_p.weight.reshape(-1), decoder.fc_p.bias]
    parts += [decoder.fc_c[l].weight.reshape(-1) for l in range(L)]
    parts += [decoder.fc_c[l].bias for l in range(L)]
    parts += [m.weight.reshape(-1) for m in lins]
    parts += [m.bias for m in lins]
    return torch.cat([p.float() for p in parts]), L, int(lins[-1].weight.shape[0]), float(decoder.points_factor)


def _contract(x: torch.Tensor, dy: torch.Tensor, out: torch.Tensor) -> None:
    """out[co, ci] += sum_p dy[p, co] x[p, ci] on the tensor-core weight-gradient kernel (dense: identity map, K = 1)."""
    lib = _lib.load()
    n, cin, cout = x.shape[0], x.shape[1], dy.shape[1]
    with _lib.on_device(x.device):
        _lib.check(lib.pv2_spconv_wgrad(_lib.ptr(x), _lib.ptr(dy), None, None, None, _lib.ptr(out), n, n, cin, cout, 1,
                                        _lib.dtype_code(x.dtype), None, 0, _lib.stream_ptr()), "pv2_spconv_wgrad")


def sdf_only(f: torch.Tensor, pts: torch.Tensor, packed: torch.Tensor, L: int, O: int, pf: float) -> torch.Tensor:
    """No-grad SDF (the sampler's coarse pass, ray_samplers.py:366-368)."""
    lib = _lib.load()
    P = f.shape[0]
    sdf = torch.empty(P, dtype=torch.float32, device=f.device)
    with _lib.on_device(f.device):
        _lib.check(lib.pv2_sdf_mlp_fwd(_lib.ptr(f.contiguous()), _lib.ptr(pts.contiguous()), _lib.ptr(packed.contiguous()), L, F,
                                       H, O, pf, P, _lib.ptr(sdf), None, None, _lib.stream_ptr()), "pv2_sdf_mlp_fwd")
    return sdf


class SdfMlpFunction(torch.autograd.Function):
    """(f [P,32], pts [P,3] (no gradient), packed parameters) -> sdf [P], u [P,32] = d sdf/d f, v [P,3] = d sdf/d p."""

    @staticmethod
    def forward(ctx, f, pts, packed, L: int, O: int, pf: float):
        lib = _lib.load()
        f, pts, packed = f.contiguous().float(), pts.contiguous().float(), packed.contiguous().float()
        P, dev = f.shape[0], f.device
        if f.shape[1] != F or packed.numel() != lib.pv2_sdf_mlp_param_count(L, O):
            raise RuntimeError("SdfMlpFunction: feature width / parameter vector do not match the kernel's layout")
        sdf = torch.empty(P, dtype=torch.float32, device=dev)
        u = torch.empty((P, F), dtype=torch.float32, device=dev)
        v = torch.empty((P, 3), dtype=torch.float32, device=dev)
        with _lib.on_device(dev):
            _lib.check(lib.pv2_sdf_mlp_fwd(_lib.ptr(f), _lib.ptr(pts), _lib.ptr(packed), L, F, H, O, pf, P, _lib.ptr(sdf),
                                           _lib.ptr(u), _lib.ptr(v), _lib.stream_ptr()), "pv2_sdf_mlp_fwd")
        ctx.save_for_backward(f, pts, packed)
        ctx.meta = (L, O, pf)
        return sdf, u, v

    @staticmethod
    def backward(ctx, g_sdf, g_u, g_v):
        f, pts, packed = ctx.saved_tensors
        L, O, pf = ctx.meta
        lib = _lib.load()
        P, dev = f.shape[0], f.device
        c = lambda t: None if t is None else t.contiguous().float()
        g_sdf, g_u, g_v = c(g_sdf), c(g_u), c(g_v)
        fbar = torch.empty((P, F), dtype=torch.float32, device=dev)
        A = torch.empty((max(L - 1, 1), P, H), dtype=torch.float32, device=dev)
        C = torch.empty_like(A)
        Z = torch.empty((L, P, H), dtype=torch.float32, device=dev)
        D = torch.empty_like(Z)
        with _lib.on_device(dev):
            _lib.check(lib.pv2_sdf_mlp_bwd(_lib.ptr(f), _lib.ptr(pts), _lib.ptr(packed), L, F, H, O, pf, P, _lib.ptr(g_sdf),
                                           _lib.ptr(g_u), _lib.ptr(g_v), _lib.ptr(fbar), _lib.ptr(A), _lib.ptr(C), _lib.ptr(Z),
                                           _lib.ptr(D), _lib.stream_ptr()), "pv2_sdf_mlp_bwd")
        # ---- parameter gradients: contractions over the points (tensor cores) + a handful of tiny matrix products
        o_wp, o_bp = 0, H * 3
        o_fc = o_bp + H
        o_bc = o_fc + L * H * F
        o_w = o_bc + L * H
        o_b = o_w + (L - 1) * H * H + O * H
        W = [packed[o_w + l * H * H: o_w + (l + 1) * H * H].view(H, H) for l in range(L - 1)]
        w_last0 = packed[o_w + (L - 1) * H * H: o_w + (L - 1) * H * H + H]
        gs = g_sdf if g_sdf is not None else torch.zeros(P, dtype=torch.float32, device=dev)
        gu = g_u if g_u is not None else torch.zeros((P, F), dtype=torch.float32, device=dev)
        gv = g_v if g_v is not None else torch.zeros((P, 3), dtype=torch.float32, device=dev)
        d = torch.zeros_like(packed)
        dwz = torch.zeros((L - 1, H, H), dtype=torch.float32, device=dev)     # A_l^T Z_l + C_l^T D_l
        dff = torch.zeros((L - 1, H, F), dtype=torch.float32, device=dev)     # A_l^T f + C_l^T gu
        for l in range(L - 1):
            _contract(Z[l], A[l], dwz[l]); _contract(D[l], C[l], dwz[l])
            _contract(f, A[l], dff[l]); _contract(gu, C[l], dff[l])
        a_sum = A[:L - 1].sum(1)                                               # [L-1, H]  = sum_p dL/dy_l
        gsum = gs.sum()
        for l in range(L - 1):
            d[o_w + l * H * H: o_w + (l + 1) * H * H] = dwz[l].reshape(-1)
            d[o_b + l * H: o_b + (l + 1) * H] = a_sum[l]
            d[o_fc + l * H * F: o_fc + (l + 1) * H * F] = (W[l].t() @ dff[l]).reshape(-1)
            d[o_bc + l * H: o_bc + (l + 1) * H] = W[l].t() @ a_sum[l]
        # last layer: dL/dy = g_sdf e_0 and ybar = e_0 are rank one
        d[o_w + (L - 1) * H * H: o_w + (L - 1) * H * H + H] = gs @ Z[L - 1] + D[L - 1].sum(0)
        d[o_b + (L - 1) * H] = gsum
        d[o_fc + (L - 1) * H * F: o_fc + L * H * F] = torch.outer(w_last0, gs @ f + gu.sum(0)).reshape(-1)
        d[o_bc + (L - 1) * H: o_bc + L * H] = w_last0 * gsum
        # fc_p: x_0 = pf (Wp p + bp),  v = pf Wp^T zbar_0   (zhat_0 = A_0 W_0, zbar_0 = C_0 W_0; L = 1: rank-one forms)
        if L > 1:
            d[o_wp: o_bp] = (pf * (W[0].t() @ (A[0].t() @ pts + C[0].t() @ gv))).reshape(-1)
            d[o_bp: o_fc] = pf * (W[0].t() @ a_sum[0])
        else:
            d[o_wp: o_bp] = (pf * torch.outer(w_last0, gs @ pts + gv.sum(0))).reshape(-1)
            d[o_bp: o_fc] = pf * w_last0 * gsum
        return fbar, None, d, None, None, None
