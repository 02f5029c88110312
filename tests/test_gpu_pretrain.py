"""End-to-end parity of one pretraining step (BASELINE configs[0] shape: 2 k voxels, 128 rays x 32 samples) between the
B200 path (`PonderIndoorStep`: backbone -> densify -> projection -> NeuS renderer -> losses, all through libpv2_b200)
and the CPU oracle chained the same way in fp64 (oracle/spconv_oracle + densify_oracle + render_oracle).  Loss terms
within 1e-3 relative (fp32 storage, 3xTF32 convolutions vs fp64), and the step must be trainable (finite gradients on
every parameter the losses reach)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import densify_oracle as do
from oracle import spconv_oracle as so
from oracle.render_oracle import NeusOracle, RenderConfig
from ponderv2_b200 import synth


def _renderer_cfg(s0, si):
    return dict(
        type="NeuSModel",
        field=dict(type="SDFField",
                   sdf_decoder=dict(in_dim=64, out_dim=65, hidden_size=128, n_blocks=1, points_factor=0.0),
                   rgb_decoder=dict(in_dim=134, out_dim=3, hidden_size=128, n_blocks=0, points_factor=0.0),
                   beta_init=0.3, use_gradient=True, volume_type="default", padding_mode="zeros", share_volume=False,
                   norm_pts=True, norm_padding=0.1),
        collider=dict(type="AABBBoxCollider", near_plane=0.01, bbox=[-0.55] * 3 + [0.55] * 3),
        sampler=dict(type="NeuSSampler", initial_sampler="UniformSampler", num_samples=s0, num_samples_importance=si,
                     num_upsample_steps=1, train_stratified=True, single_jitter=False),
        loss=dict(sensor_depth_truncation=0.05, temperature=0.01,
                  weights=dict(eikonal_loss=0.01, free_space_loss=1.0, sdf_loss=10.0, depth_loss=1.0, rgb_loss=10.0,
                               semantic_loss=0.0)))


def test_pretrain_step_matches_oracle(cuda_lib):
    from ponderv2_b200.pretrain import PonderIndoorStep
    dev = torch.device("cuda:0")
    old_tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False          # the dense projection is plain cuDNN; compare it at fp32
    try:
        torch.manual_seed(11)
        n, R, s0, si = 2000, 128, 24, 8
        grid_shape = (32, 32, 16)
        cloud = synth.indoor_cloud(n, 1001)
        rays = synth.ray_batch(R, 1008)
        gc = cloud["grid_coord"]
        rcfg = _renderer_cfg(s0, si)
        model = PonderIndoorStep(backbone=dict(in_channels=6, num_classes=0), renderer=rcfg,
                                 projection=dict(in_channels=96, out_channels=128), grid_shape=grid_shape,
                                 grid_size=0.02).to(dev).train()
        noise = {"uniform": torch.rand(R, s0 + 1), "pdf": torch.rand(R, si + 1)}
        data = dict(grid_coord=torch.from_numpy(gc).to(dev), coord=torch.from_numpy(cloud["coord"]).to(dev),
                    feat=torch.from_numpy(cloud["feat"]).to(dev), offset=torch.from_numpy(cloud["offset"]).to(dev),
                    resolution=torch.tensor([int(gc.max())], dtype=torch.int64, device=dev),
                    ray_o=torch.from_numpy(rays["rays_o"])[None].to(dev), ray_d=torch.from_numpy(rays["rays_d"])[None].to(dev),
                    rgb=torch.from_numpy(rays["rgb"]).to(dev), depth=torch.from_numpy(rays["depth"]).to(dev))
        sd = {k: v.detach().cpu().double() for k, v in model.state_dict().items()}   # before the step updates BN buffers
        pnames = {k for k, _ in model.named_parameters()}
        for k, v in sd.items():
            if k in pnames:
                v.requires_grad_(True)
        out = model(data, noise={k: v.to(dev) for k, v in noise.items()})
        out["loss"].backward()

        # ---- the same chain on the CPU oracle, fp64 -------------------------------------------------------------
        bsd = {k[len("backbone."):]: v for k, v in sd.items() if k.startswith("backbone.")}
        feats = so.spunet_forward(bsd, gc, torch.from_numpy(cloud["feat"]).double(), cloud["offset"])
        vol = do.to_dense_indoor(torch.from_numpy(cloud["coord"]), feats, cloud["offset"],
                                 np.array([int(gc.max())]), grid_shape, 0.02)            # (1, 96, Z, Y, X)
        w, b = sd["proj_net.conv.0.weight"], sd["proj_net.conv.0.bias"]
        h = torch.nn.functional.conv3d(vol, w, b, padding=1)
        mu = h.mean(dim=(0, 2, 3, 4), keepdim=True)
        var = h.var(dim=(0, 2, 3, 4), unbiased=False, keepdim=True)
        g = sd["proj_net.conv.1.weight"].view(1, -1, 1, 1, 1)
        be = sd["proj_net.conv.1.bias"].view(1, -1, 1, 1, 1)
        vol128 = torch.relu((h - mu) / torch.sqrt(var + 1e-5) * g + be)
        rsd = {k[len("renderer."):]: v for k, v in sd.items() if k.startswith("renderer.")}
        cfg = RenderConfig(bbox=[-0.55] * 3 + [0.55] * 3, near_plane=0.01, num_samples=s0, num_samples_importance=si,
                           share_volume=False, norm_pts=True, norm_padding=0.1, loss_weights=rcfg["loss"]["weights"])
        orc = NeusOracle(rsd, cfg)
        pred = orc.render(torch.from_numpy(rays["rays_o"]).double(), torch.from_numpy(rays["rays_d"]).double(),
                          [vol128[0]], {k: v.double() for k, v in noise.items()}, True)
        ld = orc.loss(pred, torch.from_numpy(rays["depth"]).double(), torch.from_numpy(rays["rgb"]).double())
        total = orc.total_loss(ld)
        total.backward()

        for k, v in ld.items():
            if k in out:
                got, ref = out[k].item(), v.item()
                assert abs(got - ref) < 1e-3 * max(1.0, abs(ref)), (k, got, ref)
        assert abs(out["loss"].item() - total.item()) < 1e-3 * max(1.0, abs(total.item()))
        missing = [k for k, p in model.named_parameters() if p.grad is None or not torch.isfinite(p.grad).all()]
        # laplace_density.beta is created by the reference field but never used by the NeuS losses
        assert all("laplace_density" in k or "fc_p" in k or "semantic" in k for k in missing), missing
        # parameter gradients of the whole step against the fp64 chain: renderer + projection tensors individually,
        # the backbone jointly (its deepest levels normalise over a few dozen rows, see test_spunet_backbone_matches_oracle)
        from tests.conftest import record
        num = den = 0.0
        worst_r, worst_r_name = 0.0, ""
        for k, p in model.named_parameters():
            ref = sd[k].grad
            # the conv bias in front of a train-mode BatchNorm has an exactly-zero gradient (the mean is subtracted)
            if ref is None or p.grad is None or ref.norm().item() < 1e-9:
                continue
            d = p.grad.detach().cpu().double() - ref
            if k.startswith("backbone."):
                num += d.pow(2).sum().item(); den += ref.pow(2).sum().item()
            else:
                e = d.norm().item() / max(ref.norm().item(), 1e-12)
                if e > worst_r:
                    worst_r, worst_r_name = e, k
        joint_b = (num / max(den, 1e-300)) ** 0.5
        record("pretrain_step_matches_oracle", loss=out["loss"].item(), loss_ref=total.item(),
               grad_renderer_projection_worst=worst_r, worst_name=worst_r_name, grad_backbone_joint=joint_b)
        assert worst_r < 5e-3, (worst_r_name, worst_r)
        # 2 k voxels: the deepest levels hold ~10 voxels, train-mode BatchNorm over so few rows amplifies the fp32-grade
        # rounding of the convolutions (measured 3.9e-2; the 8 k-voxel backbone test holds 1.5e-2, 100 k-voxel layers 4e-5)
        assert joint_b < 0.1, joint_b
    finally:
        torch.backends.cudnn.allow_tf32 = old_tf32


def test_outdoor_step_matches_oracle(cuda_lib):
    """PonderOutdoorStep (ponder_outdoor_base.py:141-265; nuScenes-style config shrunk: sdf decoder 32 -> 16 x 5 -> 17,
    shared volume, depth loss only) against the fp64 oracle chain, two scenes in the batch; masking off here (its own
    test below).  Loss within 1e-3 relative, parameter gradients as in the indoor test."""
    from ponderv2_b200.pretrain import PonderOutdoorStep
    from tests.conftest import record
    dev = torch.device("cuda:0")
    old_tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        torch.manual_seed(5)
        s0, si = 18, 6
        bbox = (-54.0, -54.0, -5.0, 54.0, 54.0, 3.0)
        grid_shape, grid_size = (30, 30, 5), (3.6, 3.6, 1.6)
        a, b = synth.outdoor_cloud(2500, 31), synth.outdoor_cloud(1500, 32)
        lo = np.array(bbox[:3], dtype=np.float32)
        coord = np.concatenate([a["coord"], b["coord"]]) + lo
        gc = np.concatenate([a["grid_coord"], b["grid_coord"]])
        feat = np.concatenate([a["feat"], b["feat"]])
        offset = np.array([2500, 4000], dtype=np.int64)
        R = [40, 24]
        rng = np.random.default_rng(9)
        start = np.tile(np.array([0.0, 0.0, -3.2], np.float32), (sum(R), 1)) + rng.normal(0, 0.5, (sum(R), 3)).astype(np.float32)
        th, rr = rng.random(sum(R)) * 2 * np.pi, rng.uniform(5.0, 50.0, sum(R))
        end = start + np.stack([rr * np.cos(th), rr * np.sin(th), rng.uniform(-1.0, 2.0, sum(R))], 1).astype(np.float32)
        ray_offset = np.cumsum(R).astype(np.int64)
        rcfg = dict(
            type="NeuSModel",
            field=dict(type="SDFField", sdf_decoder=dict(in_dim=32, out_dim=17, hidden_size=16, n_blocks=5),
                       beta_init=0.3, use_gradient=True, volume_type="default", padding_mode="zeros", share_volume=True),
            collider=dict(type="AABBBoxCollider", near_plane=0.01, bbox=[0.0, 0.0, 0.0, 1.0, 1.0, 1.0]),
            sampler=dict(type="NeuSSampler", initial_sampler="UniformSampler", num_samples=s0, num_samples_importance=si,
                         num_upsample_steps=1, train_stratified=True, single_jitter=False),
            loss=dict(sensor_depth_truncation=0.01, weights=dict(depth_loss=10.0, eikonal_loss=0.01)))
        model = PonderOutdoorStep(backbone=dict(in_channels=4, num_classes=0), renderer=rcfg,
                                  projection=dict(in_channels=96, out_channels=32), mask=None, scene_bbox=bbox,
                                  grid_shape=grid_shape, grid_size=grid_size).to(dev).train()
        sd = {k: v.detach().cpu().double() for k, v in model.state_dict().items()}
        pnames = {k for k, _ in model.named_parameters()}
        for k, v in sd.items():
            if k in pnames:
                v.requires_grad_(True)
        # one jitter tensor per scene would need per-scene noise; use the same per-ray noise rows split by scene
        nz_u, nz_p = torch.rand(sum(R), s0 + 1), torch.rand(sum(R), si + 1)

        class _Noise(dict):
            """hands each scene its own rows of the injected jitter, in call order"""
            def __init__(self):
                super().__init__(); self.calls = {"uniform": 0, "pdf": 0}
                self["_"] = 1          # non-empty: `noise or {}` keeps this object
            def get(self, key, default=None):
                if key == "mask":
                    return None
                i = self.calls[key]; self.calls[key] += 1
                lo_ = 0 if i == 0 else R[0]
                src = nz_u if key == "uniform" else nz_p
                return src[lo_:lo_ + R[i]].to(dev)
        data = dict(grid_coord=torch.from_numpy(gc).to(dev), coord=torch.from_numpy(coord).to(dev),
                    feat=torch.from_numpy(feat).to(dev), offset=torch.from_numpy(offset).to(dev),
                    ray_start=torch.from_numpy(start).to(dev), ray_end=torch.from_numpy(end).to(dev),
                    ray_offset=torch.from_numpy(ray_offset).to(dev))
        out = model(data, noise=_Noise())
        out["loss"].backward()

        # ---- oracle chain, fp64 ----
        bsd = {k[len("backbone."):]: v for k, v in sd.items() if k.startswith("backbone.")}
        feats = so.spunet_forward(bsd, gc, torch.from_numpy(feat).double(), offset)
        vol = do.to_dense_outdoor(torch.from_numpy(coord).double(), feats, offset, bbox, grid_size, grid_shape)
        w, bb_ = sd["proj_net.conv.0.weight"], sd["proj_net.conv.0.bias"]
        h = torch.nn.functional.conv3d(vol, w, bb_, padding=1)
        mu = h.mean(dim=(0, 2, 3, 4), keepdim=True)
        var = h.var(dim=(0, 2, 3, 4), unbiased=False, keepdim=True)
        g = sd["proj_net.conv.1.weight"].view(1, -1, 1, 1, 1)
        be = sd["proj_net.conv.1.bias"].view(1, -1, 1, 1, 1)
        vol32 = torch.relu((h - mu) / torch.sqrt(var + 1e-5) * g + be)
        rsd = {k[len("renderer."):]: v for k, v in sd.items() if k.startswith("renderer.")}
        cfg = RenderConfig(bbox=[0, 0, 0, 1, 1, 1], near_plane=0.01, num_samples=s0, num_samples_importance=si,
                           share_volume=True, norm_pts=False, norm_padding=0.0, sdf_points_factor=1.0, has_rgb=False,
                           loss_weights=rcfg["loss"]["weights"], sensor_depth_truncation=0.01)
        orc = NeusOracle(rsd, cfg)
        bbt = torch.tensor(bbox, dtype=torch.float64)
        nrm = lambda c: (torch.from_numpy(c).double() - bbt[:3]) / (bbt[3:] - bbt[:3])
        o_n, e_n = nrm(start), nrm(end)
        d_n = torch.nn.functional.normalize(e_n - o_n, dim=-1)
        depth = torch.linalg.norm(e_n - o_n, dim=-1, keepdim=True)
        preds = []
        for i in range(2):
            lo_, hi_ = (0, R[0]) if i == 0 else (R[0], R[0] + R[1])
            preds.append(orc.render(o_n[lo_:hi_], d_n[lo_:hi_], [vol32[i]],
                                    {"uniform": nz_u[lo_:hi_].double(), "pdf": nz_p[lo_:hi_].double()}, True))
        pred = {k: torch.cat([p_[k] for p_ in preds], 0) for k in preds[0]}
        ld = orc.loss(pred, depth, None)
        total = orc.total_loss(ld)
        total.backward()
        for k, v in ld.items():
            if k in out:
                assert abs(out[k].item() - v.item()) < 1e-3 * max(1.0, abs(v.item())), (k, out[k].item(), v.item())
        assert abs(out["loss"].item() - total.item()) < 1e-3 * max(1.0, abs(total.item()))
        num = den = 0.0
        worst_r, worst_r_name = 0.0, ""
        for k, p in model.named_parameters():
            ref = sd[k].grad
            if ref is None or p.grad is None or ref.norm().item() < 1e-9:
                continue
            dd = p.grad.detach().cpu().double() - ref
            if k.startswith("backbone."):
                num += dd.pow(2).sum().item(); den += ref.pow(2).sum().item()
            else:
                e = dd.norm().item() / max(ref.norm().item(), 1e-12)
                if e > worst_r:
                    worst_r, worst_r_name = e, k
        joint_b = (num / max(den, 1e-300)) ** 0.5
        record("outdoor_step_matches_oracle", loss=out["loss"].item(), loss_ref=total.item(),
               grad_renderer_projection_worst=worst_r, worst_name=worst_r_name, grad_backbone_joint=joint_b)
        assert worst_r < 5e-3, (worst_r_name, worst_r)
        assert joint_b < 0.1, joint_b
    finally:
        torch.backends.cudnn.allow_tf32 = old_tf32


def test_outdoor_block_masking(cuda_lib):
    """mask_features (ponder_outdoor_base.py:93-136): per scene, round(n_blocks * (1 - ratio)) blocks of mask.size^3
    voxels survive, chosen by the smallest random keys; every other voxel carries the learned token."""
    from ponderv2_b200.pretrain import PonderOutdoorStep
    dev = torch.device("cuda:0")
    torch.manual_seed(2)
    model = PonderOutdoorStep(backbone=dict(in_channels=4, num_classes=0),
                              renderer=dict(type="NeuSModel",
                                            field=dict(type="SDFField", sdf_decoder=dict(in_dim=32, out_dim=17, hidden_size=16, n_blocks=5),
                                                       beta_init=0.3, share_volume=True),
                                            collider=dict(type="AABBBoxCollider", near_plane=0.01, bbox=[0, 0, 0, 1, 1, 1]),
                                            sampler=dict(type="NeuSSampler", initial_sampler="UniformSampler", num_samples=8,
                                                         num_samples_importance=4, num_upsample_steps=1),
                                            loss=dict(sensor_depth_truncation=0.01, weights=dict(depth_loss=10.0))),
                              mask=dict(ratio=0.8, size=8, channel=4)).to(dev)
    a, b = synth.outdoor_cloud(3000, 41), synth.outdoor_cloud(2000, 42)
    gc = np.concatenate([a["grid_coord"], b["grid_coord"]])
    feat = torch.randn(5000, 4)
    offset = np.array([3000, 5000])
    blk = np.concatenate([np.repeat([0, 1], [3000, 2000])[:, None], gc // 8], 1)
    ublk, inv = np.unique(blk, axis=0, return_inverse=True)
    keys = torch.rand(ublk.shape[0])
    got = model.mask_features(torch.from_numpy(gc).to(dev), feat.to(dev), torch.from_numpy(offset).to(dev), keys.to(dev)).cpu()
    keep_blk = np.zeros(ublk.shape[0], bool)
    for s in range(2):
        ids = np.nonzero(ublk[:, 0] == s)[0]
        n_keep = int(round(len(ids) * (1 - 0.8)))
        keep_blk[ids[np.argsort(keys.numpy()[ids], kind="stable")[:n_keep]]] = True
    keep_vox = torch.from_numpy(keep_blk[inv.reshape(-1)])
    want = torch.where(keep_vox[:, None], feat, model.mtoken.detach().cpu().expand_as(feat))
    assert torch.equal(got, want)


def test_ponder_indoor_v2_from_collate_dict(cuda_lib):
    """`MODELS.build(dict(type="PonderIndoor-v2", ...))` with the ScanNet config's projection (UNet3D-v1m2) on the collate
    dict the reference dataloader produces (coord / grid_coord / feat / offset + rgb, depth, intrinsic, extrinsic,
    depth_scale): one training step runs end to end on the device, and its loss equals the same model fed with rays
    prepared by ponderv2_b200.rayprep on the CPU (the formulation pinned against the reference's ray_sample in
    tests/test_host_cpu.py) -- i.e. device ray preparation and the adapter wiring add nothing of their own."""
    from ponderv2_b200 import rayprep
    from ponderv2_b200.models import MODELS
    dev = torch.device("cuda:0")
    torch.manual_seed(4)
    B, V, H, W, n = 2, 2, 20, 28, 24
    clouds = [synth.indoor_cloud(2500, 51), synth.indoor_cloud(1800, 52)]
    coord = torch.cat([torch.from_numpy(c["coord"]) + torch.tensor([1.0, 2.0, 0.5]) for c in clouds])
    gc = torch.cat([torch.from_numpy(c["grid_coord"]) for c in clouds])
    feat = torch.cat([torch.from_numpy(c["feat"]) for c in clouds])
    offset = torch.tensor([2500, 4300])
    g = torch.Generator().manual_seed(8)
    depth = torch.rand(B, V, H, W, generator=g) * 2.0 + 0.5
    depth[torch.rand(B, V, H, W, generator=g) < 0.15] = 0.0
    rgb = torch.rand(B, V, H, W, 3, generator=g)
    intrinsic = torch.tensor([[30.0, 0, (W - 1) / 2, 0], [0, 30.0, (H - 1) / 2, 0], [0, 0, 1, 0], [0, 0, 0, 1.0]]).repeat(B, 1, 1)
    extrinsic = torch.zeros(B, V, 4, 4)
    for b in range(B):
        ctr = coord[(0 if b == 0 else 2500):(2500 if b == 0 else 4300)].mean(0)
        for v in range(V):
            q, _ = torch.linalg.qr(torch.randn(3, 3, generator=g))
            if torch.det(q) < 0:
                q[:, 0] = -q[:, 0]
            extrinsic[b, v, :3, :3] = q
            extrinsic[b, v, :3, 3] = -(q @ ctr) + torch.tensor([0.0, 0.0, 0.3])     # camera near the scene centre
            extrinsic[b, v, 3, 3] = 1.0
    collate = dict(coord=coord, grid_coord=gc, feat=feat, offset=offset, rgb=rgb, depth=depth, intrinsic=intrinsic,
                   extrinsic=extrinsic, depth_scale=torch.ones(B), condition=["ScanNet"] * B)
    cfg = dict(type="PonderIndoor-v2",
               backbone=dict(type="SpUNet-v1m1", in_channels=6, num_classes=0),
               projection=dict(type="UNet3D-v1m2", in_channels=96, out_channels=128, f_maps=8, num_levels=3),
               renderer=_renderer_cfg(24, 8), mask=None, grid_shape=(32, 32, 16), grid_size=0.02, ray_nsample=n,
               padding=0.1, render_semantic=False, conditions=("ScanNet",))
    model = MODELS.build(cfg).to(dev).train()
    pixels = rayprep.sample_pixels(depth, n, keys=torch.rand(B, V, H * W, generator=g))
    R = V * n
    noise_cpu = {"uniform": torch.rand(R, 25, generator=g), "pdf": torch.rand(R, 9, generator=g)}

    def run(prepared_on_cpu: bool):
        torch.manual_seed(0)
        for m in model.modules():                       # same BatchNorm running state for both runs is irrelevant in train mode
            pass
        noise = {k: v.to(dev) for k, v in noise_cpu.items()}
        dd = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in collate.items()}
        if not prepared_on_cpu:
            noise["pixels"] = pixels.to(dev)
            return model(dd, noise=noise)
        cube = rayprep.to_unit_cube({k: v for k, v in collate.items() if torch.is_tensor(v)})
        ray = rayprep.ray_sample(cube, n, model.bounds, pixels=pixels)
        cube = rayprep.grid_sample(cube, model.grid_size)
        step_in = dict(dd)
        step_in.update({k: cube[k].to(dev) for k in ("coord", "resolution")})
        step_in.update({k: v.to(dev) for k, v in ray.items()})
        step_in["sparse_backbone_feat"] = model.backbone(dd)
        return model.forward_after_backbone(step_in, noise)

    out_dev = run(False)
    out_dev["loss"].backward()
    assert all(torch.isfinite(v).all() for v in out_dev.values())
    n_grad = sum(p.grad is not None and bool(torch.isfinite(p.grad).all()) for p in model.parameters())
    assert n_grad > 150
    out_cpu = run(True)
    for k in out_dev:
        a, b_ = out_dev[k].item(), out_cpu[k].item()
        assert abs(a - b_) < 2e-4 * max(1.0, abs(b_)), (k, a, b_)


def test_flat_buffer_sink_gradients_match_autograd(cuda_lib):
    """With FlatParameters the sparse-conv weight gradients (on a side stream) and the BatchNorm parameter gradients are
    accumulated by the kernels straight into the flat buffer (`_pv2_sink`); they must equal the gradients autograd collects
    without it.  The yardstick is the run-to-run difference of the plain path itself: the split-K convolutions and the
    weight gradients reduce with fp32 atomics in arbitrary order, and train-mode BatchNorm over the few rows of the deep
    levels amplifies that noise."""
    from ponderv2_b200.backbone import SpUNetBase
    from ponderv2_b200.dist import FlatParameters
    from tests.conftest import record
    dev = torch.device("cuda:0")
    cloud = synth.indoor_cloud(20000, 77)
    inp = {k: torch.from_numpy(cloud[k]).to(dev) for k in ("grid_coord", "feat", "offset")}
    torch.manual_seed(1)
    m1 = SpUNetBase(in_channels=6, num_classes=0).to(dev).train()
    torch.manual_seed(1)
    m2 = SpUNetBase(in_channels=6, num_classes=0).to(dev).train()
    flat = FlatParameters(m2)
    g = torch.randn(20000, 96, device=dev)

    def grads_plain():
        m1.zero_grad()
        (m1(dict(inp)) * g).sum().backward()
        torch.cuda.synchronize()
        return [p.grad.clone() for p in m1.parameters()]

    def diff(a, b):
        num = sum((x - y).double().pow(2).sum().item() for x, y in zip(a, b))
        den = sum(x.double().pow(2).sum().item() for x in a)
        worst = max((x - y).norm().item() / max(x.norm().item(), 1e-12) for x, y in zip(a, b))
        return (num / den) ** 0.5, worst

    ga, gb = grads_plain(), grads_plain()
    noise_joint, noise_worst = diff(ga, gb)
    flat.zero_grad()
    (m2(dict(inp)) * g).sum().backward()
    flat.all_reduce_mean()          # joins the side stream
    torch.cuda.synchronize()
    for p2 in m2.parameters():
        assert p2.grad.data_ptr() >= flat.flat_grad.data_ptr()
    gs = [p.grad.clone() for p in m2.parameters()]
    sink_joint, sink_worst = diff(ga, gs)
    record("flat_buffer_sink_gradients", noise_joint=noise_joint, noise_worst=noise_worst, sink_joint=sink_joint,
           sink_worst=sink_worst)
    assert sink_joint < max(5.0 * noise_joint, 1e-5), (sink_joint, noise_joint)
    assert sink_worst < max(5.0 * noise_worst, 1e-4), (sink_worst, noise_worst)
    # two optimizer steps leave the same weights
    o1 = torch.optim.SGD(m1.parameters(), lr=0.01, momentum=0.9)
    o2 = flat.make_optimizer(torch.optim.SGD, lr=0.01, momentum=0.9)
    for _ in range(2):
        o1.zero_grad(); o2.zero_grad()
        (m1(dict(inp)) * g).sum().backward()
        (m2(dict(inp)) * g).sum().backward()
        flat.all_reduce_mean()
        o1.step(); o2.step()
    torch.cuda.synchronize()
    num = sum((p1 - p2).double().pow(2).sum().item() for p1, p2 in zip(m1.parameters(), m2.parameters()))
    den = sum(p1.double().pow(2).sum().item() for p1 in m1.parameters())
    # with |grad| ~ 1e5 here the two steps move the weights by O(1) of their norm; the 0.3 % gradient noise measured
    # above then shows up as a few 1e-3 of the weight norm (measured 3.7e-3) -- a lost or doubled gradient would be O(1)
    assert (num / den) ** 0.5 < 2e-2
