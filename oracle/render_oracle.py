"""CPU restatement of the reference's NeuS volume renderer (TEST INFRASTRUCTURE, see oracle/__init__.py).

Pinned against golden vectors generated from the reference's own code (oracle/gen_golden.py ->
tests/golden/render_*.npz).  Each step cites the reference lines it follows; paths are relative to
/root/reference/ponder/models/ponder/render_utils/.

Weights are addressed with the reference's state_dict names (`field.sdf_decoder.lin0.weight`, ...).  Randomness
(stratified jitter) is injected through `noise` so that oracle, reference and CUDA kernels consume identical
numbers: noise["uniform"] is the `torch.rand` of ray_samplers.py:78-84, noise["pdf"] the one of :262-268.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch
import torch.nn.functional as F

from .trilinear_oracle import trilinear_sample


class RenderConfig:
    """The subset of the reference's renderer config dict that changes arithmetic
    (configs/scannet/pretrain-ponder-spunet-v1m1-0-base.py:31-93)."""

    def __init__(self, *, bbox, near_plane, num_samples, num_samples_importance, num_upsample_steps=1,
                 base_variance=64.0, single_jitter=False, share_volume=False, norm_pts=True, norm_padding=0.1,
                 use_gradient=True, padding_mode="zeros", sdf_points_factor=0.0, rgb_points_factor=0.0,
                 has_rgb=True, has_semantic=False, sem_points_factor=0.0, loss_weights=None,
                 sensor_depth_truncation=0.05):
        self.bbox = [float(b) for b in bbox]
        self.near_plane = float(near_plane)
        self.num_samples = int(num_samples)
        self.num_samples_importance = int(num_samples_importance)
        self.num_upsample_steps = int(num_upsample_steps)
        self.base_variance = float(base_variance)
        self.single_jitter = bool(single_jitter)
        self.share_volume = bool(share_volume)
        self.norm_pts = bool(norm_pts)
        self.norm_padding = float(norm_padding)
        self.use_gradient = bool(use_gradient)
        self.padding_mode = padding_mode
        self.sdf_points_factor = float(sdf_points_factor)
        self.rgb_points_factor = float(rgb_points_factor)
        self.sem_points_factor = float(sem_points_factor)
        self.has_rgb = bool(has_rgb)
        self.has_semantic = bool(has_semantic)
        self.loss_weights = dict(loss_weights or {})
        self.sensor_depth_truncation = float(sensor_depth_truncation)


# ---- decoders (decoders.py:6-109) -------------------------------------------------------------
def _decoder(sd, prefix: str, points, feats, points_factor: float, act: str, final_sigmoid: bool):
    n_layers = 0
    while f"{prefix}lin{n_layers}.weight" in sd:
        n_layers += 1
    x = F.linear(points, sd[prefix + "fc_p.weight"], sd[prefix + "fc_p.bias"]) * points_factor  # :29,68,102
    for l in range(n_layers):
        x = x + F.linear(feats, sd[f"{prefix}fc_c.{l}.weight"], sd[f"{prefix}fc_c.{l}.bias"])
        x = F.linear(x, sd[f"{prefix}lin{l}.weight"], sd[f"{prefix}lin{l}.bias"])
        if l < n_layers - 1:
            x = F.softplus(x, beta=100) if act == "softplus" else torch.relu(x)
    return torch.sigmoid(x) if final_sigmoid else x


# ---- collider (scene_colliders.py:38-99) ------------------------------------------------------
def aabb_collide(o, d, bbox, near_plane):
    inv = 1.0 / (d + 1e-6)
    lo = torch.tensor(bbox[:3], dtype=o.dtype)
    hi = torch.tensor(bbox[3:], dtype=o.dtype)
    ta = (lo - o) * inv
    tb = (hi - o) * inv
    nears = torch.minimum(ta, tb).max(dim=1).values
    fars = torch.maximum(ta, tb).min(dim=1).values
    nears = nears.clamp(min=near_plane)
    miss = ~(nears < fars)
    nears = torch.where(miss, torch.zeros_like(nears), nears)
    fars = torch.where(miss, torch.zeros_like(fars), fars)
    return nears[:, None], fars[:, None]


# ---- compositing (rays.py:83-105) -------------------------------------------------------------
def weights_from_alphas(alphas):
    """alphas (R,S,1) -> weights (R,S,1)"""
    ones = torch.ones((alphas.shape[0], 1, 1), dtype=alphas.dtype)
    trans = torch.cumprod(torch.cat([ones, 1.0 - alphas + 1e-7], 1), 1)
    return alphas * trans[:, :-1, :]


class NeusOracle:
    def __init__(self, state_dict: Dict[str, torch.Tensor], cfg: RenderConfig):
        self.sd = state_dict
        self.cfg = cfg

    # fields/sdf_field.py:148-197 ------------------------------------------------------------
    def _features(self, pts, volumes: List[torch.Tensor]):
        g = pts * 2.0 - 1.0  # :156
        feats = []
        for vol in volumes:
            f = trilinear_sample(vol.unsqueeze(0).to(pts.dtype), g[None, None], self.cfg.padding_mode, True, False)
            feats.append(f.squeeze(0).squeeze(1).permute(1, 2, 0))  # (R,S,C)
        f = torch.stack(feats, dim=-2)  # (R,S,L,C)
        half = f.shape[-1] // 2
        return torch.cat([f[..., :half].flatten(-2, -1), f[..., half:].flatten(-2, -1)], dim=-1)

    def get_sdf(self, pts, volumes):
        pf = self._features(pts, volumes)
        fin = pf if self.cfg.share_volume else torch.chunk(pf, 2, dim=-1)[0]
        h = _decoder(self.sd, "field.sdf_decoder.", pts, fin, self.cfg.sdf_points_factor, "softplus", False)
        return h[..., :1], h[..., 1:], pf

    # ray_samplers.py:55-107 ------------------------------------------------------------------
    def _uniform_bins(self, R, dtype, noise, training):
        S0 = self.cfg.num_samples
        bins = torch.linspace(0.0, 1.0, S0 + 1).to(dtype).expand(R, -1)
        if training:
            t_rand = noise["uniform"].to(dtype)
            centers = (bins[..., 1:] + bins[..., :-1]) / 2.0
            upper = torch.cat([centers, bins[..., -1:]], -1)
            lower = torch.cat([bins[..., :1], centers], -1)
            bins = lower + (upper - lower) * t_rand
        return bins

    # ray_samplers.py:241-322 -----------------------------------------------------------------
    def _pdf_bins(self, weights, existing_bins, num_new, noise, training, eps=1e-5):
        num_bins = num_new + 1
        w = weights[..., 0]
        wsum = w.sum(-1, keepdim=True)
        pad = torch.relu(eps - wsum)
        w = w + pad / w.shape[-1]
        wsum = wsum + pad
        pdf = w / wsum
        cdf = torch.min(torch.ones_like(pdf), torch.cumsum(pdf, -1))
        cdf = torch.cat([torch.zeros_like(cdf[..., :1]), cdf], -1)
        u = torch.linspace(0.0, 1.0 - 1.0 / num_bins, steps=num_bins).to(cdf.dtype)
        if training:
            u = u.expand(cdf.shape[0], num_bins) + noise["pdf"].to(cdf.dtype) / num_bins
        else:
            u = (u + 1.0 / (2 * num_bins)).expand(cdf.shape[0], num_bins)
        u = u.contiguous()
        inds = torch.searchsorted(cdf, u, right=True)
        hi_idx = existing_bins.shape[-1] - 1
        below = (inds - 1).clamp(0, hi_idx)
        above = inds.clamp(0, hi_idx)
        c0, b0 = torch.gather(cdf, -1, below), torch.gather(existing_bins, -1, below)
        c1, b1 = torch.gather(cdf, -1, above), torch.gather(existing_bins, -1, above)
        den = c1 - c0
        den = torch.where(den < 1e-5, torch.ones_like(den), den)
        t = ((u - c0) / den).clip(0, 1)
        return (b0 + t * (b1 - b0)).detach()

    # ray_samplers.py:426-463 -----------------------------------------------------------------
    @staticmethod
    def _fixed_inv_s_alphas(sdf, deltas, inv_s):
        prev, nxt = sdf[:, :-1], sdf[:, 1:]
        d = deltas[:, :-1]
        mid = (prev + nxt) * 0.5
        cos = (nxt - prev) / (d + 1e-5)
        prev_cos = torch.cat([torch.zeros((sdf.shape[0], 1), dtype=sdf.dtype), cos[:, :-1]], -1)
        cos = torch.minimum(prev_cos, cos).clip(-1e3, 0.0)
        pe = mid - cos * d * 0.5
        ne = mid + cos * d * 0.5
        pc, nc = torch.sigmoid(pe * inv_s), torch.sigmoid(ne * inv_s)
        return (pc - nc + 1e-5) / (pc + 1e-5)

    def render(self, rays_o, rays_d, volumes: List[torch.Tensor], noise: Optional[dict] = None,
               training: bool = True) -> Dict[str, torch.Tensor]:
        """SurfaceModel.forward -> NeuSModel.sample_and_forward_field -> get_outputs
        (models/base_surface_model.py:34-100, models/neus.py:16-36)."""
        cfg = self.cfg
        R = rays_o.shape[0]
        dt = rays_o.dtype
        nears, fars = aabb_collide(rays_o, rays_d, cfg.bbox, cfg.near_plane)
        to_euclid = lambda b: b * fars + (1 - b) * nears  # UniformSampler: spacing_fn = identity
        o3, d3 = rays_o[:, None, :], rays_d[:, None, :]

        # ---- NeuSSampler.generate_ray_samples (ray_samplers.py:355-424) ----
        bins = self._uniform_bins(R, dt, noise, training)
        sp_starts, sp_end_last = bins[:, :-1], bins[:, -1:]
        out: Dict[str, torch.Tensor] = {}
        sdf = None
        sorted_index = None
        new_sp_starts = sp_starts
        n_new = cfg.num_samples_importance // cfg.num_upsample_steps
        for it in range(cfg.num_upsample_steps):
            with torch.no_grad():
                new_pts = o3 + d3 * to_euclid(new_sp_starts)[..., None]
                new_sdf = self.get_sdf(new_pts, volumes)[0]  # un-normalised points (neus.py:17-21)
            if sorted_index is not None:
                merged = torch.cat([sdf.squeeze(-1), new_sdf.squeeze(-1)], -1)
                sdf = torch.gather(merged, 1, sorted_index).unsqueeze(-1)
            else:
                sdf = new_sdf
            eu = to_euclid(torch.cat([sp_starts, sp_end_last], -1))
            deltas = eu[:, 1:] - eu[:, :-1]
            alphas = self._fixed_inv_s_alphas(sdf.squeeze(-1), deltas, cfg.base_variance * 2 ** it)
            w = weights_from_alphas(alphas.unsqueeze(-1))
            w = torch.cat((w, torch.zeros_like(w[:, :1])), dim=1)
            if it == 0:
                out["init_sampled_points"] = new_pts
                out["init_weights"] = w
            existing = torch.cat([sp_starts, sp_end_last], -1)
            nb = self._pdf_bins(w, existing, n_new, noise, training)
            new_sp_starts, new_end = nb[:, :-1], nb[:, -1:]
            npts = o3 + d3 * to_euclid(new_sp_starts)[..., None]
            out["new_sampled_points"] = npts if "new_sampled_points" not in out else torch.cat(
                [out["new_sampled_points"], npts], 1)
            # RayBundle.merge_ray_samples (rays.py:118-153)
            sp_end_last = torch.maximum(sp_end_last, new_end)
            sp_starts, sorted_index = torch.sort(torch.cat([sp_starts, new_sp_starts], -1), -1)
            sp_starts = sp_starts.detach()

        eu = to_euclid(torch.cat([sp_starts, sp_end_last], -1))
        starts, ends = eu[:, :-1, None], eu[:, 1:, None]
        deltas = ends - starts

        # ---- SDFField.forward (fields/sdf_field.py:211-284) ----
        pts = o3 + d3 * starts
        if cfg.norm_pts:  # normalize_3d_coordinate :58-74
            pts = pts / (1 + cfg.norm_padding + 10e-4) + 0.5
            pts = torch.where(pts >= 1, torch.full_like(pts, 1 - 10e-4), pts)
            pts = torch.where(pts < 0, torch.zeros_like(pts), pts)
        pts = pts.detach().requires_grad_(True)
        with torch.enable_grad():
            sdf, geo, pf = self.get_sdf(pts, volumes)
            grads = torch.autograd.grad(sdf, pts, torch.ones_like(sdf), create_graph=True, retain_graph=True)[0]
        dirs = d3.expand(-1, starts.shape[1], -1)
        rgb_in = ([grads] if cfg.use_gradient else []) + [
            pf if cfg.share_volume else torch.chunk(pf, 2, dim=-1)[1], geo, dirs]
        if cfg.has_rgb:
            rgb = _decoder(self.sd, "field.rgb_decoder.", pts, torch.cat(rgb_in, -1), cfg.rgb_points_factor,
                           "relu", True)
        if cfg.has_semantic:
            sem = _decoder(self.sd, "field.semantic_decoder.", pts, torch.cat(rgb_in[:-1], -1),
                           cfg.sem_points_factor, "relu", False)
        # get_alpha :122-146 (cos_anneal_ratio is never updated -> 1)
        inv_s = torch.exp(self.sd["field.deviation_network.variance"] * 10.0).clip(1e-6, 1e6)
        true_cos = (dirs * grads).sum(-1, keepdim=True)
        iter_cos = -(F.relu(-true_cos * 0.5 + 0.5) * 0.0 + F.relu(-true_cos) * 1.0)
        nxt = sdf + iter_cos * deltas * 0.5
        prv = sdf - iter_cos * deltas * 0.5
        pc, nc = torch.sigmoid(prv * inv_s), torch.sigmoid(nxt * inv_s)
        alpha = ((pc - nc + 1e-5) / (pc + 1e-5)).clip(0.0, 1.0)
        weights = weights_from_alphas(alpha)

        # ---- renderers.py:5-75 ----
        depth = (weights * starts).sum(-2) / (weights.sum(-2) + 1e-10)
        depth = torch.clip(depth, starts.min(), starts.max())
        normal = (weights * F.normalize(grads, dim=-1)).sum(-2)
        if cfg.has_rgb:
            comp = (weights * rgb).sum(-2)
            comp = comp + comp.new_tensor((0.0, 0.0, 0.0)) * (1.0 - weights.sum(-2))
            if not training:
                comp = comp.clamp(0.0, 1.0)
            out["rgb"] = comp
        if cfg.has_semantic:
            out["semantic"] = (weights * sem).sum(-2)
        out.update(depth=depth, normal=normal, weights=weights, sdf=sdf, gradients=grads, z_vals=starts,
                   sampled_points=o3 + d3 * starts)
        return out

    # models/base_surface_model.py:102-211 (without the semantic / sparse-point terms) -----------
    def loss(self, preds, depth_gt, rgb_gt=None) -> Dict[str, torch.Tensor]:
        lw = self.cfg.loss_weights
        trunc = self.cfg.sensor_depth_truncation
        ld = {}
        valid = depth_gt > 0.0
        if lw.get("depth_loss", 0.0) > 0:
            ld["depth_loss"] = (valid * (depth_gt - preds["depth"]).abs()).sum() / valid.sum().clamp(min=1.0) \
                * lw["depth_loss"]
        if lw.get("rgb_loss", 0.0) > 0:
            ld["rgb_loss"] = F.l1_loss(preds["rgb"], rgb_gt) * lw["rgb_loss"]
            ld["psnr"] = 20.0 * torch.log10(1.0 / (preds["rgb"] - rgb_gt).pow(2).mean().sqrt())
        sdf = preds["sdf"][..., 0]
        z = preds["z_vals"][..., 0]
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
            ld["eikonal_loss"] = ((preds["gradients"].norm(2, dim=-1) - 1) ** 2).mean() * lw["eikonal_loss"]
        return ld

    @staticmethod
    def total_loss(loss_dict) -> torch.Tensor:
        """ponder_indoor_base.py:676-679"""
        return sum(v for k, v in loss_dict.items() if "loss" in k)
