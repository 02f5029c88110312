"""Helpers shared by the parity tests: load a tests/golden/render_*.npz fixture into oracle objects."""
import json
from pathlib import Path

import numpy as np
import torch

GOLD = Path(__file__).resolve().parent / "golden"


def load_render_case(name: str, dtype=torch.float32):
    from oracle.render_oracle import RenderConfig

    z = np.load(GOLD / f"render_{name}.npz")
    meta = json.loads(str(z["meta"]))
    arr = {k: torch.from_numpy(z[k]) for k in z.files if k != "meta"}
    sd = {k[len("param."):]: v.to(dtype) for k, v in arr.items() if k.startswith("param.")}
    if meta["kind"] == "indoor":
        cfg = RenderConfig(bbox=[-0.55] * 3 + [0.55] * 3, near_plane=0.01, num_samples=meta["S0"],
                           num_samples_importance=meta["Si"], share_volume=False, norm_pts=True, norm_padding=0.1,
                           sdf_points_factor=0.0, rgb_points_factor=0.0, has_rgb=True,
                           loss_weights=dict(eikonal_loss=0.01, free_space_loss=1.0, sdf_loss=10.0, depth_loss=1.0,
                                             rgb_loss=10.0))
    else:
        cfg = RenderConfig(bbox=[0, 0, 0, 1, 1, 1], near_plane=0.01, num_samples=meta["S0"],
                           num_samples_importance=meta["Si"], share_volume=True, norm_pts=False, norm_padding=0.0,
                           sdf_points_factor=1.0, has_rgb=False,
                           loss_weights=dict(eikonal_loss=0.01, free_space_loss=1.0, sdf_loss=10.0, depth_loss=1.0))
    return meta, arr, sd, cfg


def product_renderer_cfg(meta):
    """The reference-style config dict (configs/scannet|nuscenes pretrain) for a golden case."""
    if meta["kind"] == "indoor":
        field = dict(type="SDFField",
                     sdf_decoder=dict(in_dim=64, out_dim=65, hidden_size=128, n_blocks=1, points_factor=0.0),
                     rgb_decoder=dict(in_dim=134, out_dim=3, hidden_size=128, n_blocks=0, points_factor=0.0),
                     beta_init=0.3, use_gradient=True, volume_type="default", padding_mode="zeros",
                     share_volume=False, norm_pts=True, norm_padding=0.1)
        collider = dict(type="AABBBoxCollider", near_plane=0.01, bbox=[-0.55] * 3 + [0.55] * 3)
        weights = dict(eikonal_loss=0.01, free_space_loss=1.0, sdf_loss=10.0, depth_loss=1.0, rgb_loss=10.0,
                       semantic_loss=0.0)
    else:
        field = dict(type="SDFField",
                     sdf_decoder=dict(in_dim=32, out_dim=17, hidden_size=16, n_blocks=5, points_factor=1.0),
                     rgb_decoder=None, semantic_decoder=None, beta_init=0.3, use_gradient=True,
                     volume_type="default", padding_mode="zeros", share_volume=True, norm_pts=False,
                     norm_padding=0.0)
        collider = dict(type="AABBBoxCollider", near_plane=0.01, bbox=[0, 0, 0, 1, 1, 1])
        weights = dict(eikonal_loss=0.01, free_space_loss=1.0, sdf_loss=10.0, depth_loss=1.0, rgb_loss=0.0,
                       semantic_loss=0.0)
    sampler = dict(type="NeuSSampler", initial_sampler="UniformSampler", num_samples=meta["S0"],
                   num_samples_importance=meta["Si"], num_upsample_steps=1, train_stratified=True,
                   single_jitter=False)
    loss = dict(sensor_depth_truncation=0.05, temperature=0.01, weights=weights)
    if meta.get("semantic"):   # §8f-4 fixtures (oracle/gen_golden.py:renderer_cfg)
        field["semantic_decoder"] = dict(in_dim=131, out_dim=meta["semantic"], hidden_size=128, n_blocks=0,
                                         points_factor=0.0)
        weights["semantic_loss"] = 0.1
        loss["val_ray_split"] = 5
    return dict(type="NeuSModel", field=field, collider=collider, sampler=sampler, loss=loss)
