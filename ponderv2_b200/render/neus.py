"""Host-side mirror of the reference NeuS renderer, generic ("composed") path.

Module names, constructor arguments, state_dict keys and the forward / get_loss contracts follow
ponder/models/ponder/render_utils/{models/neus.py:7-36, models/base_surface_model.py:13-211, fields/sdf_field.py:77-284,
decoders.py:6-109, ray_samplers.py:55-107,227-463, rays.py:83-153, renderers.py:5-75, scene_colliders.py:26-99}.

This file is the configuration-generic path: torch elementwise ops around the twice-differentiable CUDA sampler
(ponderv2_b200.smooth_sampler, boundary B2).  Differences from the reference, none of which change results:
  * no host synchronisation: `normalize_3d_coordinate`'s data-dependent branches (sdf_field.py:70-73) become
    `torch.where`, DepthRenderer's global min/max (renderers.py:49-50) stay on the device, the sampler Functions do
    not call `.item()`;
  * the jitter noise can be injected (`noise={"uniform","pdf"}`) so parity tests consume the oracle's numbers;
  * the volume may be channels-first or channels_last_3d.
The single-kernel fused path for the indoor configuration is ponderv2_b200.render.fused (when enabled by NeuSModel).
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch
import torch.nn.functional as F
from torch import nn

from ..smooth_sampler import SmoothSampler
from . import fused, mlp, ray


_NOTED = set()


def _note_once(key: str, msg: str) -> None:
    """One warning per process and reason when a configuration leaves the hand-written fast path (never silent)."""
    if key not in _NOTED:
        _NOTED.add(key)
        import warnings
        warnings.warn(msg, RuntimeWarning, stacklevel=3)


class RayBundle:
    """origins [R,3], directions [R,3] (+ nears/fars [R,1] once collided); rays.py:108-116."""

    def __init__(self, origins, directions, nears=None, fars=None, **kwargs):
        self.origins, self.directions, self.nears, self.fars = origins, directions, nears, fars


class _LossCfg(dict):
    def __getattr__(self, k):
        try:
            v = self[k]
        except KeyError:
            raise AttributeError(k) from None
        return _LossCfg(v) if isinstance(v, dict) and not isinstance(v, _LossCfg) else v


# ------------------------------------------------------------------------------------------ decoders
class _Decoder(nn.Module):
    def __init__(self, in_dim, out_dim, hidden_size=256, n_blocks=5, points_factor=1.0, **kwargs):
        super().__init__()
        dims = [hidden_size] * (n_blocks + 1) + [out_dim]
        self.num_layers = len(dims)
        for l in range(self.num_layers - 1):
            setattr(self, f"lin{l}", nn.Linear(dims[l], dims[l + 1]))
        self.fc_c = nn.ModuleList([nn.Linear(in_dim, hidden_size) for _ in range(self.num_layers - 1)])
        self.fc_p = nn.Linear(3, hidden_size)
        self.points_factor = points_factor

    def _act(self, x):
        raise NotImplementedError

    def _final(self, x):
        return x

    def forward(self, points, point_feats):
        x = self.fc_p(points) * self.points_factor
        for l in range(self.num_layers - 1):
            x = getattr(self, f"lin{l}")(x + self.fc_c[l](point_feats))
            if l < self.num_layers - 2:
                x = self._act(x)
        return self._final(x)


class SDFDecoder(_Decoder):
    def __init__(self, *a, **k):
        super().__init__(*a, **k)
        self.activation = nn.Softplus(beta=100)

    def _act(self, x):
        return self.activation(x)


class RGBDecoder(_Decoder):
    def __init__(self, in_dim, out_dim=3, **k):
        super().__init__(in_dim, out_dim, **k)
        self.activation = nn.ReLU()

    def _act(self, x):
        return self.activation(x)

    def _final(self, x):
        return torch.sigmoid(x)


class SemanticDecoder(_Decoder):
    def __init__(self, *a, **k):
        super().__init__(*a, **k)
        self.activation = nn.ReLU()

    def _act(self, x):
        return self.activation(x)


class LaplaceDensity(nn.Module):
    def __init__(self, init_val, beta_min=0.0001):
        super().__init__()
        self.register_parameter("beta_min", nn.Parameter(beta_min * torch.ones(1), requires_grad=False))
        self.register_parameter("beta", nn.Parameter(init_val * torch.ones(1), requires_grad=True))

    def forward(self, sdf, beta=None):
        beta = self.beta.abs() + self.beta_min if beta is None else beta
        return (1.0 / beta) * (0.5 + 0.5 * sdf.sign() * torch.expm1(-sdf.abs() / beta))


class SingleVarianceNetwork(nn.Module):
    def __init__(self, init_val):
        super().__init__()
        self.register_parameter("variance", nn.Parameter(init_val * torch.ones(1), requires_grad=True))

    def get_variance(self):
        return torch.exp(self.variance * 10.0).clip(1e-6, 1e6)


# ------------------------------------------------------------------------------------------ field
class SDFField(nn.Module):
    def __init__(self, sdf_decoder, beta_init, use_gradient=True, volume_type="default", padding_mode="zeros",
                 share_volume=True, rgb_decoder=None, semantic_decoder=None, norm_pts=False, norm_padding=0.1,
                 **kwargs):
        super().__init__()
        if volume_type != "default":
            raise NotImplementedError(volume_type)
        self.beta_init, self.volume_type, self.padding_mode = beta_init, volume_type, padding_mode
        self.share_volume, self.use_gradient = share_volume, use_gradient
        self.sdf_decoder = SDFDecoder(**sdf_decoder)
        self.rgb_decoder = RGBDecoder(**rgb_decoder) if rgb_decoder is not None else None
        self.semantic_decoder = SemanticDecoder(**semantic_decoder) if semantic_decoder is not None else None
        self.laplace_density = LaplaceDensity(init_val=beta_init)
        self.deviation_network = SingleVarianceNetwork(init_val=beta_init)
        self._cos_anneal_ratio = 1.0
        self.norm_pts, self.norm_padding = norm_pts, norm_padding
        # fused SDF-decoder kernels for the outdoor decoder shape (render/mlp.py); NeuSModel.use_fused switches it
        self.use_mlp_kernel = False

    def set_cos_anneal_ratio(self, anneal):
        self._cos_anneal_ratio = anneal

    def feature_sampling(self, pts_norm, volume_feature: List[torch.Tensor]):
        g = (pts_norm * 2 - 1)[None, None].contiguous()
        feats = []
        for vol in volume_feature:
            f = SmoothSampler.apply(vol.unsqueeze(0).to(g.dtype).contiguous(), g, self.padding_mode, True, False)
            feats.append(f.squeeze(0).squeeze(1).permute(1, 2, 0))
        f = torch.stack(feats, dim=-2)
        half = f.shape[-1] // 2
        return torch.cat([f[..., :half].flatten(-2, -1), f[..., half:].flatten(-2, -1)], dim=-1)

    def get_sdf(self, points, volume_feature):
        pf = self.feature_sampling(points, volume_feature)
        h = self.sdf_decoder(points, pf if self.share_volume else torch.chunk(pf, 2, dim=-1)[0])
        return h[..., :1], h[..., 1:], pf

    def mlp_kernel_ok(self, volume_feature) -> bool:
        return (self.use_mlp_kernel and self.share_volume and len(volume_feature) == 1
                and volume_feature[0].shape[0] == mlp.F and mlp.eligible(self))

    def sdf_no_grad(self, points, volume_feature):
        """SDF only, no graph (NeuSSampler's coarse pass): sampler + fused decoder kernel when the shape allows."""
        if not self.mlp_kernel_ok(volume_feature):
            return self.get_sdf(points, volume_feature)[0].squeeze(-1)
        pf = self.feature_sampling(points, volume_feature)
        packed, L, O, pfac = mlp.pack(self.sdf_decoder)
        return mlp.sdf_only(pf.reshape(-1, mlp.F), points.reshape(-1, 3), packed, L, O, pfac).view(points.shape[:-1])

    def _forward_mlp_kernel(self, points, volume_feature):
        """sdf and d sdf / d points with the fused decoder: sdf, u = d sdf/d f, v = d sdf/d p from one kernel, the sampler's
        own (twice differentiable) backward supplies (d f / d points)^T u."""
        pf = self.feature_sampling(points, volume_feature)                    # [R,S,32], differentiable in points / volume
        packed, L, O, pfac = mlp.pack(self.sdf_decoder)
        sdf, u, v = mlp.SdfMlpFunction.apply(pf.reshape(-1, mlp.F), points.reshape(-1, 3), packed, L, O, pfac)
        gp = torch.autograd.grad(pf, points, u.view_as(pf), create_graph=True, retain_graph=True, only_inputs=True)[0]
        return sdf.view(*points.shape[:-1], 1), gp + v.view_as(points)

    def get_alpha(self, directions, deltas, sdf, gradients):
        inv_s = self.deviation_network.get_variance()
        true_cos = (directions * gradients).sum(-1, keepdim=True)
        r = self._cos_anneal_ratio
        iter_cos = -(F.relu(-true_cos * 0.5 + 0.5) * (1.0 - r) + F.relu(-true_cos) * r)
        nxt = sdf + iter_cos * deltas * 0.5
        prv = sdf - iter_cos * deltas * 0.5
        pc, nc = torch.sigmoid(prv * inv_s), torch.sigmoid(nxt * inv_s)
        return ((pc - nc + 1e-5) / (pc + 1e-5)).clip(0.0, 1.0)

    def forward(self, points, directions, deltas, volume_feature, return_alphas=True, normalized=False):
        if self.norm_pts and not normalized:
            points = points / (1 + self.norm_padding + 10e-4) + 0.5
            points = torch.where(points >= 1, torch.full_like(points, 1 - 10e-4), points)
            points = torch.where(points < 0, torch.zeros_like(points), points)
        points = points.detach().requires_grad_(True)
        if self.mlp_kernel_ok(volume_feature):
            with torch.enable_grad():
                sdf, gradients = self._forward_mlp_kernel(points, volume_feature)
            out = dict(sdf=sdf, gradients=gradients, normal=F.normalize(gradients, dim=-1))
            if return_alphas:
                out["alphas"] = self.get_alpha(directions, deltas, sdf, gradients)
            return out
        with torch.enable_grad():
            sdf, geo, pf = self.get_sdf(points, volume_feature)
            gradients = torch.autograd.grad(sdf, points, torch.ones_like(sdf), create_graph=True, retain_graph=True,
                                            only_inputs=True)[0]
        rgb_inputs = ([gradients] if self.use_gradient else []) + [
            pf if self.share_volume else torch.chunk(pf, 2, dim=-1)[1], geo, directions]
        out = {}
        if self.rgb_decoder is not None:
            out["rgb"] = self.rgb_decoder(points, torch.cat(rgb_inputs, dim=-1))
        if self.semantic_decoder is not None:
            out["semantic"] = self.semantic_decoder(points, torch.cat(rgb_inputs[:-1], dim=-1))
        out.update(sdf=sdf, gradients=gradients, normal=F.normalize(gradients, dim=-1))
        if return_alphas:
            out["alphas"] = self.get_alpha(directions, deltas, sdf, gradients)
        return out


# ------------------------------------------------------------------------------------------ collider / sampler
class AABBBoxCollider(nn.Module):
    def __init__(self, bbox, near_plane, **kwargs):
        super().__init__()
        self.bbox, self.near_plane = [float(b) for b in bbox], float(near_plane)

    def forward(self, ray_bundle: RayBundle) -> RayBundle:
        if ray_bundle.nears is not None and ray_bundle.fars is not None:
            return ray_bundle
        o, d = ray_bundle.origins, ray_bundle.directions
        inv = 1.0 / (d + 1e-6)
        lo, hi = o.new_tensor(self.bbox[:3]), o.new_tensor(self.bbox[3:])
        ta, tb = (lo - o) * inv, (hi - o) * inv
        nears = torch.minimum(ta, tb).max(dim=1).values.clamp(min=self.near_plane)
        fars = torch.maximum(ta, tb).min(dim=1).values
        hit = nears < fars
        zero = torch.zeros_like(nears)
        ray_bundle.nears = torch.where(hit, nears, zero)[..., None]
        ray_bundle.fars = torch.where(hit, fars, zero)[..., None]
        return ray_bundle


def weights_from_alphas(alphas):
    trans = torch.cumprod(torch.cat([torch.ones_like(alphas[:, :1]), 1.0 - alphas + 1e-7], 1), 1)
    return alphas * trans[:, :-1]


class NeuSSampler(nn.Module):
    """UniformSampler -> (no-grad SDF, fixed-inv_s alphas, PDF resample, merge) x num_upsample_steps."""

    def __init__(self, initial_sampler, num_samples, num_samples_importance, num_upsample_steps, base_variance=64.0,
                 train_stratified=True, single_jitter=True, **kwargs):
        super().__init__()
        if initial_sampler != "UniformSampler":
            raise NotImplementedError(f"initial_sampler={initial_sampler}: every shipped config uses UniformSampler")
        self.num_samples, self.num_samples_importance = num_samples, num_samples_importance
        self.num_upsample_steps, self.base_variance = num_upsample_steps, base_variance
        self.train_stratified, self.single_jitter = train_stratified, single_jitter

    def uniform_bins(self, R, ref, noise):
        S0 = self.num_samples
        bins = torch.linspace(0.0, 1.0, S0 + 1, device=ref.device, dtype=ref.dtype).expand(R, -1)
        if self.train_stratified and self.training:
            t = noise if noise is not None else torch.rand((R, 1 if self.single_jitter else S0 + 1),
                                                           dtype=ref.dtype, device=ref.device)
            c = (bins[..., 1:] + bins[..., :-1]) / 2.0
            upper, lower = torch.cat([c, bins[..., -1:]], -1), torch.cat([bins[..., :1], c], -1)
            bins = lower + (upper - lower) * t
        return bins

    def pdf_bins(self, weights, existing, num_new, noise, eps=1e-5):
        nb = num_new + 1
        w = weights
        wsum = w.sum(-1, keepdim=True)
        pad = torch.relu(eps - wsum)
        w = w + pad / w.shape[-1]
        pdf = w / (wsum + pad)
        cdf = torch.cat([torch.zeros_like(pdf[..., :1]), torch.cumsum(pdf, -1).clamp(max=1.0)], -1)
        u = torch.linspace(0.0, 1.0 - 1.0 / nb, steps=nb, device=cdf.device, dtype=cdf.dtype)
        if self.train_stratified and self.training:
            r = noise if noise is not None else torch.rand((cdf.shape[0], 1 if self.single_jitter else nb),
                                                           device=cdf.device, dtype=cdf.dtype)
            u = u.expand(cdf.shape[0], nb) + r / nb
        else:
            u = (u + 1.0 / (2 * nb)).expand(cdf.shape[0], nb)
        inds = torch.searchsorted(cdf, u.contiguous(), right=True)
        top = existing.shape[-1] - 1
        below, above = (inds - 1).clamp(0, top), inds.clamp(0, top)
        c0, b0 = torch.gather(cdf, -1, below), torch.gather(existing, -1, below)
        c1, b1 = torch.gather(cdf, -1, above), torch.gather(existing, -1, above)
        den = c1 - c0
        den = torch.where(den < 1e-5, torch.ones_like(den), den)
        return (b0 + ((u - c0) / den).clip(0, 1) * (b1 - b0)).detach()

    @staticmethod
    def fixed_inv_s_alphas(sdf, deltas, inv_s):
        prev, nxt, d = sdf[:, :-1], sdf[:, 1:], deltas[:, :-1]
        mid = (prev + nxt) * 0.5
        cos = (nxt - prev) / (d + 1e-5)
        cos = torch.minimum(torch.cat([torch.zeros_like(cos[:, :1]), cos[:, :-1]], -1), cos).clip(-1e3, 0.0)
        pc = torch.sigmoid((mid - cos * d * 0.5) * inv_s)
        nc = torch.sigmoid((mid + cos * d * 0.5) * inv_s)
        return (pc - nc + 1e-5) / (pc + 1e-5)


# ------------------------------------------------------------------------------------------ model
class NeuSModel(nn.Module):
    def __init__(self, field, collider, sampler, loss, **kwargs):
        super().__init__()
        field = dict(field); field.pop("type", None)
        collider = dict(collider); collider.pop("type", None)
        sampler = dict(sampler); sampler.pop("type", None)
        self.field = SDFField(**field)
        self.collider = AABBBoxCollider(**collider)
        self.sampler = NeuSSampler(**sampler)
        self.loss = _LossCfg(loss)
        self.anneal_end = 50000
        self.background_color = (0.0, 0.0, 0.0)
        # hand-derived tensor-core field (render/fused.py) whenever the configuration is the indoor one; the generic
        # autograd-through-the-sampler path otherwise.  Both are CUDA-only.
        self.use_fused = True
        # per-ray CUDA kernels for collider / sampler / compositing / losses (render/ray.py); torch elementwise otherwise
        self.use_ray_kernels = True
        self._loss_consts = {}

    def forward(self, ray_bundle: RayBundle, volume_feature: List[torch.Tensor], noise: Optional[dict] = None,
                **kwargs) -> Dict[str, torch.Tensor]:
        if not ray_bundle.origins.is_cuda:
            raise RuntimeError("NeuSModel: rays must be CUDA tensors (ponderv2_b200 has no CPU path)")
        # The renderer always computes in fp32 (the kernels take fp32 buffers; the reference runs its sampler in fp32 too,
        # sdf_field.py:162).  Under the trainer's autocast (engines/train.py:183-196) the folded decoder matrices would
        # otherwise come out of `@` as half-precision tensors.
        with torch.autocast(device_type="cuda", enabled=False):
            volume_feature = [v.float() for v in volume_feature]
            return self._forward_fp32(ray_bundle, volume_feature, noise, **kwargs)

    def _forward_fp32(self, ray_bundle: RayBundle, volume_feature: List[torch.Tensor], noise: Optional[dict] = None,
                      **kwargs) -> Dict[str, torch.Tensor]:
        noise = noise or {}
        rb = self.collider(ray_bundle)
        o3, d3 = rb.origins[:, None, :], rb.directions[:, None, :]
        nears, fars = rb.nears, rb.fars
        to_euclid = lambda b: b * fars + (1 - b) * nears
        smp = self.sampler
        R = rb.origins.shape[0]
        use_fused = (self.use_fused and fused.eligible(self.field) and len(volume_feature) == 1
                     and volume_feature[0].shape[0] == 128)
        can_ray = self.use_ray_kernels and ray.supported(smp.num_samples, smp.num_samples_importance,
                                                         smp.num_upsample_steps)
        self.field.use_mlp_kernel = bool(self.use_fused)
        if self.use_fused and not use_fused and not self.field.mlp_kernel_ok(volume_feature):
            _note_once("field", "NeuSModel: this field configuration is neither the indoor one the tensor-core field kernels "
                       "are written for (sdf 64->128->65, rgb 134->128->3, C = 128, share_volume=False) nor the outdoor "
                       "one of the fused decoder kernel (sdf 32->16x<=8, shared volume, no colour head); the SDF/colour "
                       "MLPs run as torch linears around the CUDA trilinear sampler")
        if self.use_ray_kernels and not can_ray:
            _note_once("ray", f"NeuSModel: sampler {smp.num_samples}+{smp.num_samples_importance} x "
                       f"{smp.num_upsample_steps} steps is outside the per-ray kernels' limits ({ray.MAX_S0}+{ray.MAX_SI}, "
                       f"total {ray.MAX_S}, one upsample step); sampling and compositing run as torch ops")
        if use_fused and can_ray:
            return self._forward_ray_kernels(rb, volume_feature[0], noise)
        if can_ray:
            return self._forward_ray_kernels_generic(rb, volume_feature, noise)
        if use_fused:
            vol_cl = volume_feature[0].permute(1, 2, 3, 0).contiguous().float()  # free for channels_last_3d volumes
            fp = fused.fold_parameters(self.field)
            vol_ng = vol_cl.detach()
            M0_ng, c0_ng = fp["M0"].detach(), fp["c0"].detach()
            w4_ng, c4_ng = fp["wcat"][:4].detach().contiguous(), fp["c1"][:4].detach().contiguous()
        bins = smp.uniform_bins(R, rb.origins, noise.get("uniform"))
        starts_sp, end_sp = bins[:, :-1], bins[:, -1:]
        out: Dict[str, torch.Tensor] = {}
        sdf, sorted_index, new_sp = None, None, starts_sp
        n_new = smp.num_samples_importance // max(smp.num_upsample_steps, 1)
        for it in range(smp.num_upsample_steps):
            with torch.no_grad():
                new_pts = o3 + d3 * to_euclid(new_sp)[..., None]
                if use_fused:
                    new_sdf = fused.coarse_sdf(vol_ng, new_pts.reshape(-1, 3), M0_ng, c0_ng, w4_ng, c4_ng).view(
                        new_pts.shape[:-1])
                else:
                    new_sdf = self.field.sdf_no_grad(new_pts, volume_feature)
            sdf = new_sdf if sorted_index is None else torch.gather(torch.cat([sdf, new_sdf], -1), 1, sorted_index)
            eu = to_euclid(torch.cat([starts_sp, end_sp], -1))
            alphas = smp.fixed_inv_s_alphas(sdf, eu[:, 1:] - eu[:, :-1], smp.base_variance * 2 ** it)
            w = weights_from_alphas(alphas)
            w = torch.cat([w, torch.zeros_like(w[:, :1])], 1)
            if it == 0:
                out["init_sampled_points"], out["init_weights"] = new_pts, w[..., None]
            nb = smp.pdf_bins(w, torch.cat([starts_sp, end_sp], -1), n_new, noise.get("pdf"))
            new_sp, new_end = nb[:, :-1], nb[:, -1:]
            npts = o3 + d3 * to_euclid(new_sp)[..., None]
            out["new_sampled_points"] = npts if "new_sampled_points" not in out else torch.cat(
                [out["new_sampled_points"], npts], 1)
            end_sp = torch.maximum(end_sp, new_end)
            starts_sp, sorted_index = torch.sort(torch.cat([starts_sp, new_sp], -1), -1)
        eu = to_euclid(torch.cat([starts_sp, end_sp], -1)).detach()
        starts, ends = eu[:, :-1, None], eu[:, 1:, None]
        deltas = ends - starts
        pts = o3 + d3 * starts
        dirs = d3.expand(-1, starts.shape[1], -1)
        if use_fused:
            fld = self.field
            pn = pts
            if fld.norm_pts:
                pn = pn / (1 + fld.norm_padding + 10e-4) + 0.5
                pn = torch.where(pn >= 1, torch.full_like(pn, 1 - 10e-4), pn)
                pn = torch.where(pn < 0, torch.zeros_like(pn), pn)
            S = starts.shape[1]
            sdf_f, grad_f, rgb_f = fused.FusedFieldFunction.apply(
                vol_cl, pn.reshape(-1, 3), rb.directions, S, fp["M0"], fp["c0"], fp["wcat"], fp["c1"], fp["wp"],
                fp["m10"], fp["Mr"], fp["cr"])
            sdf_f, grad_f = sdf_f.view(R, S, 1), grad_f.view(R, S, 3)
            fo = dict(sdf=sdf_f, gradients=grad_f, rgb=rgb_f.view(R, S, 3), normal=F.normalize(grad_f, dim=-1),
                      alphas=fld.get_alpha(dirs, deltas, sdf_f, grad_f))
        else:
            fo = self.field(pts, dirs, deltas, volume_feature, return_alphas=True)
        weights = weights_from_alphas(fo["alphas"])
        depth = (weights * starts).sum(-2) / (weights.sum(-2) + 1e-10)
        depth = torch.maximum(torch.minimum(depth, starts.amax()), starts.amin())
        out["depth"] = depth
        out["normal"] = (weights * fo["normal"]).sum(-2)
        if "rgb" in fo:
            rgb = (weights * fo["rgb"]).sum(-2)
            rgb = rgb + rgb.new_tensor(self.background_color) * (1.0 - weights.sum(-2))
            out["rgb"] = rgb if self.training else rgb.clamp(0.0, 1.0)
        if "semantic" in fo:
            out["semantic"] = (weights * fo["semantic"]).sum(-2)
        out.update(weights=weights, sdf=fo["sdf"], gradients=fo["gradients"], z_vals=starts, sampled_points=pts)
        return out

    def _forward_ray_kernels(self, rb: RayBundle, volume: torch.Tensor, noise: dict) -> Dict[str, torch.Tensor]:
        """Indoor configuration, everything proportional to rays x samples in libpv2_b200: ray_setup -> coarse SDF
        (tensor cores) -> ray_resample -> fused field (tensor cores, analytic gradient) -> ray_composite."""
        smp, fld = self.sampler, self.field
        o, d = rb.origins, rb.directions
        R, S0, Si = o.shape[0], smp.num_samples, smp.num_samples_importance
        S = S0 + Si
        jitter = smp.train_stratified and self.training
        nz_u = nz_p = None
        if jitter:
            nz_u = noise.get("uniform")
            if nz_u is None:
                nz_u = torch.rand((R, 1 if smp.single_jitter else S0 + 1), dtype=torch.float32, device=o.device)
            nz_p = noise.get("pdf")
            if nz_p is None:
                nz_p = torch.rand((R, 1 if smp.single_jitter else Si + 1), dtype=torch.float32, device=o.device)
        vol_cl = volume.permute(1, 2, 3, 0).contiguous().float()   # free for channels_last_3d volumes
        fp = fused.fold_parameters(fld)
        nears, fars, bins, pts_c = ray.ray_setup(o, d, S0, self.collider.bbox, self.collider.near_plane, nz_u)
        rb.nears, rb.fars = nears, fars
        with torch.no_grad():
            sdf_c = fused.coarse_sdf(vol_cl.detach(), pts_c.view(-1, 3), fp["M0"].detach(), fp["c0"].detach(),
                                     fp["wcat"][:4].detach().contiguous(), fp["c1"][:4].detach().contiguous())
        starts, deltas, pn, init_w, new_bins, minmax = ray.ray_resample(
            o, d, nears, fars, bins, sdf_c, Si, smp.base_variance, nz_p, fld.norm_pts, fld.norm_padding)
        sdf_f, grad_f, rgb_f = fused.FusedFieldFunction.apply(
            vol_cl, pn.view(-1, 3), d, S, fp["M0"], fp["c0"], fp["wcat"], fp["c1"], fp["wp"], fp["m10"], fp["Mr"],
            fp["cr"])
        sdf_f, grad_f, rgb_f = sdf_f.view(R, S), grad_f.view(R, S, 3), rgb_f.view(R, S, 3)
        weights, rgb, depth, normal = ray.RayComposite.apply(
            sdf_f, grad_f, rgb_f, fld.deviation_network.variance, starts, deltas, d, minmax, fld._cos_anneal_ratio,
            not self.training)
        o3, d3 = o[:, None, :], d[:, None, :]
        z = starts[..., None]
        new_e = new_bins * fars + (1 - new_bins) * nears
        return dict(rgb=rgb, depth=depth[:, None], normal=normal, weights=weights[..., None], sdf=sdf_f[..., None],
                    gradients=grad_f, z_vals=z, sampled_points=o3 + d3 * z, init_sampled_points=pts_c,
                    init_weights=init_w[..., None], new_sampled_points=o3 + d3 * new_e[..., None])

    def _forward_ray_kernels_generic(self, rb: RayBundle, volume_feature: List[torch.Tensor], noise: dict):
        """Any field configuration (outdoor: sdf 32->16x5->17, no colour head; ponder_outdoor_base.py:218-251): the
        per-ray kernels do collider / sampler / compositing, the field is SDFField.forward (torch linears around the
        twice-differentiable CUDA sampler, autograd supplies d sdf / d p as in sdf_field.py:226-238)."""
        smp, fld = self.sampler, self.field
        o, d = rb.origins, rb.directions
        R, S0, Si = o.shape[0], smp.num_samples, smp.num_samples_importance
        S = S0 + Si
        jitter = smp.train_stratified and self.training
        nz_u = nz_p = None
        if jitter:
            nz_u = noise.get("uniform")
            if nz_u is None:
                nz_u = torch.rand((R, 1 if smp.single_jitter else S0 + 1), dtype=torch.float32, device=o.device)
            nz_p = noise.get("pdf")
            if nz_p is None:
                nz_p = torch.rand((R, 1 if smp.single_jitter else Si + 1), dtype=torch.float32, device=o.device)
        nears, fars, bins, pts_c = ray.ray_setup(o, d, S0, self.collider.bbox, self.collider.near_plane, nz_u)
        rb.nears, rb.fars = nears, fars
        with torch.no_grad():   # coarse pass: un-normalised points, as the reference's sampler calls get_sdf directly
            sdf_c = fld.sdf_no_grad(pts_c, volume_feature)
        starts, deltas, pn, init_w, new_bins, minmax = ray.ray_resample(
            o, d, nears, fars, bins, sdf_c, Si, smp.base_variance, nz_p, fld.norm_pts, fld.norm_padding)
        dirs = d[:, None, :].expand(-1, S, -1)
        fo = fld(pn, dirs, deltas[..., None], volume_feature, return_alphas=False, normalized=True)
        rgbs = fo["rgb"] if "rgb" in fo else None
        weights, rgb, depth, normal = ray.RayComposite.apply(
            fo["sdf"].squeeze(-1), fo["gradients"], rgbs, fld.deviation_network.variance, starts, deltas, d, minmax,
            fld._cos_anneal_ratio, not self.training)
        o3, d3 = o[:, None, :], d[:, None, :]
        z = starts[..., None]
        new_e = new_bins * fars + (1 - new_bins) * nears
        out = dict(depth=depth[:, None], normal=normal, weights=weights[..., None], sdf=fo["sdf"],
                   gradients=fo["gradients"], z_vals=z, sampled_points=o3 + d3 * z, init_sampled_points=pts_c,
                   init_weights=init_w[..., None], new_sampled_points=o3 + d3 * new_e[..., None])
        if rgb is not None:
            out["rgb"] = rgb
        if "semantic" in fo:
            out["semantic"] = (weights[..., None] * fo["semantic"]).sum(-2)
        return out

    def _fused_loss(self, preds_dict, targets):
        lw = self.loss.weights
        sdf, z, grad = preds_dict["sdf"][..., 0], preds_dict["z_vals"][..., 0], preds_dict["gradients"]
        R, S = sdf.shape
        rgb_gt = targets.get("rgb")
        has_rgb = "rgb" in preds_dict and rgb_gt is not None and lw.get("rgb_loss", 0.0) > 0
        wts = (float(lw.get("depth_loss", 0.0)), float(lw.get("rgb_loss", 0.0)) if has_rgb else 0.0,
               float(lw.get("free_space_loss", 0.0)), float(lw.get("sdf_loss", 0.0)), float(lw.get("eikonal_loss", 0.0)))
        key = (R, S, sdf.device, wts)
        if key not in self._loss_consts:
            self._loss_consts[key] = (torch.tensor(wts, dtype=torch.float32, device=sdf.device),
                                      torch.tensor([1.0, 3.0 * R, 1.0, 1.0, float(R) * S], dtype=torch.float32,
                                                   device=sdf.device))
        wvec, const = self._loss_consts[key]
        terms = ray.RayLoss.apply(preds_dict["depth"], preds_dict["rgb"] if has_rgb else None, sdf, grad, z,
                                  targets["depth"], rgb_gt if has_rgb else None,
                                  float(self.loss.sensor_depth_truncation), wvec, const)
        ld = {}
        for i, name in enumerate(("depth_loss", "rgb_loss", "free_space_loss", "sdf_loss", "eikonal_loss")):
            if wts[i] > 0:
                ld[name] = terms[i]
        if has_rgb:
            ld["psnr"] = 20.0 * torch.log10(1.0 / terms[5].sqrt())
        return ld

    def get_loss(self, preds_dict, targets):
        with torch.autocast(device_type="cuda", enabled=False):
            return self._get_loss_fp32(preds_dict, targets)

    def _semantic_loss(self, preds_dict, targets):
        """base_surface_model.py:123-173 (§8f-4): rendered per-ray feature vs the pixel's text embedding, a contrastive
        cross entropy over the rays of the batch (logits = normalised prediction . every ray's target / temperature, the
        ray's own target is the label; rays without depth or without a class are ignored).  Same value as the reference's
        `F.cross_entropy(..., ignore_index)`, written as a masked mean so that an all-ignored batch gives 0 (the
        reference's explicit branch) without reading the count back to the host.  Eval: mean of the per-chunk losses
        over `loss.val_ray_split`-ray chunks (the reference's chunk loop reads an undefined `chunk_idx`, :161; the
        evident intent - chunk c's masks - is what runs here)."""
        pred = F.normalize(preds_dict["semantic"].float(), dim=-1)
        gt = targets["semantic"].float()
        mask = ((targets["depth"] > 0.0) & gt.any(dim=-1, keepdim=True)).reshape(-1)
        temp = float(self.loss.temperature)

        def ce(p, g, m):
            logits = (p @ g.t()) / temp
            nll = torch.logsumexp(logits, dim=1) - logits.diagonal()
            return (nll * m).sum() / m.sum().clamp(min=1.0)

        if self.training:
            return ce(pred, gt, mask.float())
        chunk = int(self.loss.get("val_ray_split", 128))
        parts = [ce(p, g, m.float()) for p, g, m in zip(pred.split(chunk), gt.split(chunk), mask.split(chunk))]
        return torch.stack(parts).mean()

    def _get_loss_fp32(self, preds_dict, targets):
        lw = self.loss.weights
        if lw.get("semantic_loss", 0.0) > 0:
            if "semantic" not in preds_dict or "semantic" not in targets:
                raise RuntimeError("semantic_loss > 0 needs field.semantic_decoder and a `semantic` ray target")
            sem = {"semantic_loss": self._semantic_loss(preds_dict, targets) * lw["semantic_loss"]}
        else:
            sem = {}
        if self.use_ray_kernels and preds_dict["sdf"].is_cuda and preds_dict["sdf"].dtype == torch.float32:
            return {**self._fused_loss(preds_dict, targets), **sem}
        ld = dict(sem)
        depth_gt = targets["depth"]
        valid = depth_gt > 0.0
        if lw.get("depth_loss", 0.0) > 0:
            ld["depth_loss"] = (valid * (depth_gt - preds_dict["depth"]).abs()).sum() / valid.sum().clamp(min=1.0) \
                * lw["depth_loss"]
        if lw.get("rgb_loss", 0.0) > 0:
            rgb_pred, rgb_gt = preds_dict["rgb"], targets["rgb"]
            ld["rgb_loss"] = F.l1_loss(rgb_pred, rgb_gt) * lw["rgb_loss"]
            ld["psnr"] = 20.0 * torch.log10(1.0 / (rgb_pred - rgb_gt).pow(2).mean().sqrt())
        sdf, z = preds_dict["sdf"][..., 0], preds_dict["z_vals"][..., 0]
        trunc = self.loss.sensor_depth_truncation
        front = valid & (z < (depth_gt - trunc))
        back = valid & (z > (depth_gt + trunc))
        sdf_mask = valid & (~front) & (~back)
        if lw.get("free_space_loss", 0.0) > 0:
            ld["free_space_loss"] = (F.relu(trunc - sdf) * front).sum() / front.sum().clamp(min=1.0) \
                * lw["free_space_loss"]
        if lw.get("sdf_loss", 0.0) > 0:
            ld["sdf_loss"] = ((z + sdf - depth_gt).abs() * sdf_mask).sum() / sdf_mask.sum().clamp(min=1.0) \
                * lw["sdf_loss"]
        if lw.get("eikonal_loss", 0.0) > 0:
            ld["eikonal_loss"] = ((preds_dict["gradients"].norm(2, dim=-1) - 1) ** 2).mean() * lw["eikonal_loss"]
        return ld


def build_renderer(cfg, **kwargs):
    cfg = dict(cfg)
    t = cfg.pop("type", "NeuSModel")
    if t != "NeuSModel":
        raise NotImplementedError(f"renderer {t}: every shipped pretraining config uses NeuSModel")
    return NeuSModel(**cfg, **kwargs)
