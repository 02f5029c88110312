"""Hand-derived SDF field for the indoor renderer configuration: channels-last trilinear fetch, SDF / colour MLPs on
the tensor cores (3xTF32) and the analytic gradient d sdf / d p with its second-order backward.

Replaces, for `sdf_decoder(n_blocks=1)`, `rgb_decoder(n_blocks=0)`, `points_factor=0`, `share_volume=False` (every
indoor pretraining config), the autograd graph the reference builds in fields/sdf_field.py:211-257:
`SmoothSampler.apply` -> SDFDecoder -> `autograd.grad(sdf, points, create_graph=True)` -> RGBDecoder, whose backward
re-enters `smooth_sampler._C.backward_backward`.

Algebra (f = trilinear features, f_s / f_r its two halves, J = d f_s / d p):
  h = M0 f_s + c0,  a = softplus_100(h),  s = sigmoid(100 h)                M0 = lin0.W fc_c[0].W   (decoders.py:28-36:
  out = W1 a + M1 f_s + c1,  sdf = out[0], geo = out[1:]                    M1 = lin1.W fc_c[1].W    x = lin(x + fc_c(f)),
  u = d sdf / d f_s = M0^T (s * W1[0]) + M1[0],   grad = J^T u                                       points_factor = 0)
  rgb = sigmoid(Mr [grad | f_r | geo | dir] + cr)                            Mr = rgb.lin0.W rgb.fc_c[0].W
The folded matrices are formed with torch ops (tiny, differentiable), so gradients reach the original parameters
through ordinary autograd; everything proportional to the number of sample points runs in libpv2_b200 kernels.
"""
from __future__ import annotations

from typing import Tuple

import torch

from .. import _lib


def _linear(x, x_row, x_lo, presplit, w, bias, y, y_row, y_lo, y_split, act, y2, y2_row, y2_lo, rows, cin, cout):
    lib = _lib.load()
    ws_bytes = lib.pv2_linear_workspace_bytes(rows, cin, cout, int(presplit))
    w = w.contiguous()
    with _lib.on_device(w.device):
        ws = _lib.workspace(ws_bytes, w.device)
        _lib.check(lib.pv2_linear(_lib.C.c_void_p(x.data_ptr()), x_row, x_lo, int(presplit), _lib.ptr(w),
                                  _lib.ptr(bias.contiguous()) if bias is not None else None,
                                  _lib.C.c_void_p(y.data_ptr()), y_row, y_lo, int(y_split), act,
                                  _lib.C.c_void_p(y2.data_ptr()) if y2 is not None else None, y2_row, y2_lo, rows, cin,
                                  cout, _lib.ptr(ws), ws_bytes, _lib.stream_ptr()), "pv2_linear")


def _dense_wgrad(x, x_row, x_lo, dy, dy_row, dy_lo, rows, cin, cout):
    lib = _lib.load()
    dw = torch.zeros((cout, cin), dtype=torch.float32, device=dy.device)
    ws_bytes = lib.pv2_wgrad_workspace_bytes(rows, rows, cin, cout)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dy.device)
    with _lib.on_device(dy.device):
        _lib.check(lib.pv2_dense_wgrad(_lib.C.c_void_p(x.data_ptr()), x_row, x_lo, _lib.C.c_void_p(dy.data_ptr()), dy_row,
                                       dy_lo, rows, cin, cout, _lib.ptr(dw), _lib.ptr(ws), ws_bytes, _lib.stream_ptr()),
                   "pv2_dense_wgrad")
    return dw


def _vol_dims(vol_cl: torch.Tensor) -> Tuple[int, int, int, int]:
    if vol_cl.dim() != 4 or not vol_cl.is_contiguous() or vol_cl.dtype != torch.float32:
        raise RuntimeError("fused field: volume must be a contiguous fp32 [Z, Y, X, C] tensor")
    return tuple(int(v) for v in vol_cl.shape)


def coarse_sdf(vol_cl: torch.Tensor, pts: torch.Tensor, M0, c0, wcat4, c14) -> torch.Tensor:
    """No-grad SDF at the sampler's coarse points (NeuSSampler, ray_samplers.py:366-368): fetch the SDF half of the
    channels, two tensor-core layers, return sdf [P]."""
    lib = _lib.load()
    Z, Y, X, C = _vol_dims(vol_cl)
    P = pts.shape[0]
    dev = pts.device
    xa = torch.empty((P, 192), dtype=torch.float32, device=dev)  # [a(128) | f_s(64)], plain fp32 (split on chip)
    pts = pts.contiguous()
    with _lib.on_device(dev):
        _lib.check(lib.pv2_field_sample_fwd(_lib.ptr(vol_cl), _lib.ptr(pts), P, Z, Y, X, C, 64, 64,
                                            _lib.C.c_void_p(xa.data_ptr() + 128 * 4), 192, 0, None, 0,
                                            _lib.stream_ptr()), "pv2_field_sample_fwd")
    _linear(xa[:, 128:], 192, 0, False, M0, c0, xa, 192, 0, False, 1, None, 0, 0, P, 64, 128)
    out = torch.empty((P, 4), dtype=torch.float32, device=dev)
    _linear(xa, 192, 0, False, wcat4, c14, out, 4, 0, False, 0, None, 0, 0, P, 192, 4)
    return out[:, 0]


class FusedFieldFunction(torch.autograd.Function):
    """(volume [Z,Y,X,C], pts [P,3] normalised, dirs [R,3]) -> sdf [P], grad [P,3], rgb [P,3]."""

    @staticmethod
    def forward(ctx, vol_cl, pts, dirs, samples_per_ray: int, M0, c0, wcat, c1, wp, m10, Mr, cr):
        lib = _lib.load()
        Z, Y, X, C = _vol_dims(vol_cl)
        if C != 128:
            raise RuntimeError("fused field: the indoor configuration has C = 128 (64 SDF + 64 colour channels)")
        P = pts.shape[0]
        dev = pts.device
        pts = pts.contiguous()
        dirs = dirs.contiguous()
        xa = torch.empty((P, 192), dtype=torch.float32, device=dev)      # [a | f_s] plain fp32
        s_act = torch.empty((P, 128), dtype=torch.float32, device=dev)   # sigmoid(100 h) = d softplus / d h
        f_r = torch.empty((P, 64), dtype=torch.float32, device=dev)
        out = torch.empty((P, 68), dtype=torch.float32, device=dev)      # [sdf | geo(64) | 0 0 0]
        u = torch.empty((P, 64), dtype=torch.float32, device=dev)
        grad = torch.empty((P, 3), dtype=torch.float32, device=dev)
        rgb = torch.empty((P, 3), dtype=torch.float32, device=dev)
        sp = _lib.stream_ptr
        with _lib.on_device(dev):
            _lib.check(lib.pv2_field_sample_fwd(_lib.ptr(vol_cl), _lib.ptr(pts), P, Z, Y, X, C, 128, 64,
                                                _lib.C.c_void_p(xa.data_ptr() + 128 * 4), 192, 0, _lib.ptr(f_r), 64,
                                                sp()), "pv2_field_sample_fwd")
        # h = M0 f_s + c0 -> a (into xa[:, 0:128]) and s
        _linear(xa[:, 128:], 192, 0, False, M0, c0, xa, 192, 0, False, 1, s_act, 128, 0, P, 64, 128)
        # out = [W1 | M1] [a | f_s] + c1
        _linear(xa, 192, 0, False, wcat, c1, out, 68, 0, False, 0, None, 0, 0, P, 192, 68)
        # u = (M0 * W1[0])^T s + M1[0]
        _linear(s_act, 128, 0, False, wp, m10, u, 64, 0, False, 0, None, 0, 0, P, 128, 64)
        Mr_c, cr_c = Mr.contiguous(), cr.contiguous()
        with _lib.on_device(dev):
            _lib.check(lib.pv2_field_post_fwd(_lib.ptr(vol_cl), _lib.ptr(pts), _lib.ptr(dirs), samples_per_ray,
                                              _lib.ptr(u), _lib.ptr(f_r), _lib.C.c_void_p(out.data_ptr() + 4), 68,
                                              _lib.ptr(Mr_c), _lib.ptr(cr_c), P, Z, Y, X, C, _lib.ptr(grad),
                                              _lib.ptr(rgb), sp()), "pv2_field_post_fwd")
        ctx.save_for_backward(vol_cl, pts, dirs, xa, s_act, f_r, out, u, grad, rgb, M0, wcat, wp, Mr_c)
        ctx.spr = samples_per_ray
        return out[:, 0].contiguous(), grad, rgb

    @staticmethod
    def backward(ctx, g_sdf, g_grad, g_rgb):
        vol_cl, pts, dirs, xa, s, f_r, out, u, grad, rgb, M0, wcat, wp, Mr = ctx.saved_tensors
        lib = _lib.load()
        Z, Y, X, C = _vol_dims(vol_cl)
        P = pts.shape[0]
        dev = pts.device
        z = lambda t, shape: torch.zeros(shape, dtype=torch.float32, device=dev) if t is None else t.contiguous()
        g_sdf, g_grad, g_rgb = z(g_sdf, (P,)), z(g_grad, (P, 3)), z(g_rgb, (P, 3))
        gbar = torch.empty((P, 3), dtype=torch.float32, device=dev)
        dF = torch.empty((P, 128), dtype=torch.float32, device=dev)
        doutbar = torch.empty((P, 72), dtype=torch.float32, device=dev)   # 68 used, rows padded to 8 channels
        ubar = torch.empty((P, 64), dtype=torch.float32, device=dev)
        small = torch.zeros(3 * 134 + 3 + 64 + 68, dtype=torch.float32, device=dev)   # one fill for the four accumulators
        dMr, dcr = small[:402].view(3, 134), small[402:405]
        d_m10, d_c1 = small[405:469], small[469:537]
        sp = _lib.stream_ptr
        with _lib.on_device(dev):
            _lib.check(lib.pv2_field_post_bwd(_lib.ptr(vol_cl), _lib.ptr(pts), _lib.ptr(dirs), ctx.spr, _lib.ptr(f_r),
                                              _lib.C.c_void_p(out.data_ptr() + 4), 68, _lib.ptr(grad), _lib.ptr(rgb),
                                              _lib.ptr(Mr), _lib.ptr(g_rgb), _lib.ptr(g_grad), _lib.ptr(g_sdf), P, Z, Y,
                                              X, C, _lib.ptr(gbar), _lib.ptr(dF), 128, _lib.ptr(doutbar), _lib.ptr(ubar),
                                              _lib.ptr(dMr), _lib.ptr(dcr), _lib.ptr(d_m10), _lib.ptr(d_c1), sp()),
                       "pv2_field_post_bwd")
        # through u = s wp^T + m10, then through s = sigmoid(100 h):  hbar = (ubar wp^T) * 100 s (1 - s)   [epilogue 2]
        hbar = torch.empty((P, 128), dtype=torch.float32, device=dev)
        _linear(ubar, 64, 0, False, wp.t().contiguous(), None, hbar, 128, 0, False, 2, s, 128, 0, P, 64, 128)
        d_wp = _dense_wgrad(s, 128, 0, ubar, 64, 0, P, 128, 64)
        # through out = [a | f_s] wcat^T + c1:  hbar += (doutbar wcat[:, :128]) * s   [epilogue 3];  dF[:, :64] = doutbar wcat[:, 128:]
        wcat_t = torch.nn.functional.pad(wcat.t(), (0, 4))                # [192, 72], zero columns against the row padding
        _linear(doutbar, 72, 0, False, wcat_t[:128], None, hbar, 128, 0, False, 3, s, 128, 0, P, 72, 128)
        _linear(doutbar, 72, 0, False, wcat_t[128:], None, dF, 128, 0, False, 0, None, 0, 0, P, 72, 64)
        d_wcat = _dense_wgrad(xa, 192, 0, doutbar, 72, 0, P, 192, 68)
        # through a = softplus(h), h = f_s M0^T + c0:  dF[:, :64] += hbar M0   [epilogue 4]
        _linear(hbar, 128, 0, False, M0.t().contiguous(), None, dF, 128, 0, False, 4, None, 0, 0, P, 128, 64)
        d_M0 = _dense_wgrad(xa[:, 128:], 192, 0, hbar, 128, 0, P, 64, 128)
        d_c0 = hbar.sum(0)
        dvol = torch.zeros_like(vol_cl)
        with _lib.on_device(dev):
            _lib.check(lib.pv2_field_sample_bwd(_lib.ptr(pts), _lib.ptr(dF), 128, _lib.ptr(u), _lib.ptr(gbar), P, Z, Y, X,
                                                C, 64, _lib.ptr(dvol), sp()), "pv2_field_sample_bwd")
        return dvol, None, None, None, d_M0, d_c0, d_wcat, d_c1.clone(), d_wp, d_m10.clone(), dMr.clone(), dcr.clone()


def fold_parameters(field) -> dict:
    """Differentiable folding of the decoders' bias-connected linear pairs (see module docstring)."""
    sd, rd = field.sdf_decoder, field.rgb_decoder
    W0, b0 = sd.lin0.weight, sd.lin0.bias
    W1, b1 = sd.lin1.weight, sd.lin1.bias
    M0 = W0 @ sd.fc_c[0].weight
    c0 = W0 @ sd.fc_c[0].bias + b0
    M1 = W1 @ sd.fc_c[1].weight
    c1 = W1 @ sd.fc_c[1].bias + b1
    wcat = torch.cat([W1, M1], dim=1)                                   # [65, 192]
    pad = wcat.new_zeros((3, wcat.shape[1]))
    Rw, Rb = rd.lin0.weight, rd.lin0.bias
    return dict(
        M0=M0, c0=c0,
        wcat=torch.cat([wcat, pad], 0), c1=torch.cat([c1, c1.new_zeros(3)]),
        wp=(M0 * W1[0][:, None]).t().contiguous(), m10=M1[0],
        Mr=Rw @ rd.fc_c[0].weight, cr=Rw @ rd.fc_c[0].bias + Rb,
    )


def eligible(field) -> bool:
    sd, rd = field.sdf_decoder, field.rgb_decoder
    return (rd is not None and field.semantic_decoder is None and not field.share_volume and field.use_gradient
            and field.padding_mode == "zeros" and sd.num_layers == 3 and rd.num_layers == 2
            and sd.points_factor == 0.0 and rd.points_factor == 0.0
            and sd.lin0.weight.shape == (128, 128) and sd.fc_c[0].weight.shape == (128, 64)
            and sd.lin1.weight.shape == (65, 128) and rd.fc_c[0].weight.shape[1] == 134
            and rd.lin0.weight.shape[0] == 3)
