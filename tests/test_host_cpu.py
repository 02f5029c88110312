"""CPU checks of host-side logic that needs no kernel: the differentiable folding of the decoders' bias-connected linear
pairs (render/fused.py) against the reference-shaped decoder modules, and the densify cell arithmetic against the
oracle's restatement of PonderIndoor/PonderOutdoor.to_dense."""
import numpy as np
import torch

from oracle import densify_oracle as do
from ponderv2_b200 import densify, synth
from ponderv2_b200.render import fused
from ponderv2_b200.render.neus import SDFField


def _field():
    torch.manual_seed(3)
    return SDFField(sdf_decoder=dict(in_dim=64, out_dim=65, hidden_size=128, n_blocks=1, points_factor=0.0),
                    rgb_decoder=dict(in_dim=134, out_dim=3, hidden_size=128, n_blocks=0, points_factor=0.0),
                    beta_init=0.3, use_gradient=True, padding_mode="zeros", share_volume=False, norm_pts=True).double()


def test_folded_decoders_equal_module_forward():
    """decoders.py:28-36: x = lin(x + fc_c(f)); with points_factor = 0 the two linears of every layer fold into one
    matrix.  The folded algebra of render/fused.py must reproduce the module's forward exactly (fp64)."""
    fld = _field()
    assert fused.eligible(fld)
    fp = fused.fold_parameters(fld)
    P = 257
    f_s = torch.randn(P, 64, dtype=torch.float64)
    pts = torch.rand(P, 3, dtype=torch.float64)
    ref = fld.sdf_decoder(pts, f_s)                                         # [P, 65]
    h = f_s @ fp["M0"].t() + fp["c0"]
    a = torch.nn.functional.softplus(h, beta=100)
    out = torch.cat([a, f_s], 1) @ fp["wcat"].t() + fp["c1"]                # [P, 68], last three columns are padding
    assert (out[:, :65] - ref).abs().max().item() < 1e-10
    assert out[:, 65:].abs().max().item() == 0.0
    # u = d sdf / d f_s (what the tensor-core layer 3 computes) against autograd through the module
    f2 = f_s.clone().requires_grad_(True)
    sdf = fld.sdf_decoder(pts, f2)[:, 0]
    (u_ref,) = torch.autograd.grad(sdf.sum(), f2)
    s = torch.sigmoid(100.0 * h)
    u = s @ fp["wp"].t() + fp["m10"]
    assert (u - u_ref).abs().max().item() < 1e-9
    # colour head: sigmoid(lin0(fc_c0(x))) with no hidden nonlinearity (n_blocks = 0)
    xin = torch.randn(P, 134, dtype=torch.float64)
    rgb_ref = fld.rgb_decoder(pts, xin)
    rgb = torch.sigmoid(xin @ fp["Mr"].t() + fp["cr"])
    assert (rgb - rgb_ref).abs().max().item() < 1e-12


def test_fold_is_differentiable_to_the_reference_parameters():
    fld = _field()
    fp = fused.fold_parameters(fld)
    (fp["M0"].sum() + fp["wcat"].sum() + fp["wp"].sum() + fp["Mr"].sum() + fp["c0"].sum() + fp["c1"].sum() + fp["cr"].sum()
     + fp["m10"].sum()).backward()
    for name in ["sdf_decoder.lin0.weight", "sdf_decoder.lin1.weight", "sdf_decoder.fc_c.0.weight", "sdf_decoder.fc_c.1.bias",
                 "rgb_decoder.lin0.weight", "rgb_decoder.fc_c.0.weight", "rgb_decoder.fc_c.0.bias"]:
        p = dict(fld.named_parameters())[name]
        assert p.grad is not None and p.grad.abs().sum().item() > 0, name


def test_indoor_cells_match_oracle_indexing():
    """densify.indoor_cells reproduces the reference's float floor-divisions (ponder_indoor_base.py:199-213): scattering
    ones with the product's cell ids gives the oracle's occupancy volume."""
    c = synth.indoor_cloud(3000, 42)
    coord = torch.from_numpy(c["coord"])
    gs = (32, 32, 16)
    res = torch.tensor([int(c["grid_coord"].max())])
    batch = torch.zeros(coord.shape[0], dtype=torch.int64)
    cell = densify.indoor_cells(coord, batch, res, gs, 0.02)
    X, Y, Z = gs
    feat = torch.ones(coord.shape[0], 1, dtype=torch.float64)
    ref = do.to_dense_indoor(coord, feat, c["offset"], res.numpy(), gs, 0.02)[0, 0]       # (Z, Y, X) occupancy (mean of ones)
    got = torch.zeros(Z * Y * X, dtype=torch.float64)
    ok = cell >= 0
    got[cell[ok]] = 1.0
    assert torch.equal(got.view(Z, Y, X), ref)


def test_outdoor_cells_match_oracle_indexing():
    c = synth.outdoor_cloud(4000, 7)
    coord = torch.from_numpy(c["coord"])
    bbox, gsz, gshape = [0, 0, 0, 108, 108, 8], [0.6, 0.6, 1.6], (180, 180, 5)   # configs/nuscenes scene_bbox / grid
    batch = torch.zeros(coord.shape[0], dtype=torch.int64)
    cell = densify.outdoor_cells(coord, batch, bbox, gsz, gshape)
    X, Y, Z = gshape
    feat = torch.ones(coord.shape[0], 1, dtype=torch.float64)
    ref = do.to_dense_outdoor(coord, feat, np.array([coord.shape[0]]), bbox, gsz, list(gshape))[0, 0]
    got = torch.zeros(Z * Y * X, dtype=torch.float64)
    ok = cell >= 0
    got[cell[ok]] = 1.0
    assert torch.equal(got.view(Z, Y, X), ref)


def test_rayprep_matches_reference_golden():
    """ponderv2_b200.rayprep (batched, sync-free) against the reference's own to_unit_cube / ray_sample / grid_sample
    outputs (tests/golden/rayprep_indoor.npz, oracle/gen_golden.py::gen_rayprep_case), given the same pixel choice."""
    import json
    from pathlib import Path
    from ponderv2_b200 import rayprep
    z = np.load(Path(__file__).resolve().parent / "golden" / "rayprep_indoor.npz")
    meta = json.loads(str(z["meta"]))
    t = {k: torch.from_numpy(z[k]) for k in z.files if k != "meta"}
    dd = {k[3:]: v.clone() for k, v in t.items() if k.startswith("in.")}
    cube = rayprep.to_unit_cube(dd)
    for k in ("coord", "extrinsic", "depth_scale", "pc_scale", "bbox"):
        ref = t["cube." + k]
        assert cube[k].shape == ref.shape, k
        assert (cube[k] - ref).abs().max().item() < 2e-5 * max(1.0, ref.abs().max().item()), k
    pad = meta["padding"]
    bounds = [[-0.5 - pad / 2] * 3, [0.5 + pad / 2] * 3]
    cube["index2semantic"] = t["index2semantic"]          # semantic branch (§8f-4): class ids -> text embeddings
    ray = rayprep.ray_sample(cube, meta["n"], bounds, pixels=t["pixels"])
    assert float(t["ray.semantic"].abs().sum()) > 0
    for k in ("ray_o", "ray_d", "rgb", "depth", "semantic"):
        ref = t["ray." + k]
        assert ray[k].shape == ref.shape, (k, ray[k].shape, ref.shape)
        assert (ray[k] - ref).abs().max().item() < 5e-5 * max(1.0, ref.abs().max().item()), k
    grid = rayprep.grid_sample(cube, meta["grid_size"])
    assert torch.equal(grid["resolution"].long(), t["grid.resolution"].long())
    assert torch.equal(grid["bbox"].long(), t["grid.bbox"].long())
    # the sync-free pixel sampler returns n distinct valid pixels per view
    pix = rayprep.sample_pixels(dd["depth"], meta["n"])
    d = dd["depth"].reshape(meta["B"], meta["V"], -1)
    assert (torch.gather(d, 2, pix) > 0).all()
    assert all(len(set(pix[b, v].tolist())) == meta["n"] for b in range(meta["B"]) for v in range(meta["V"]))


def test_unet3d_v1m2_matches_reference_golden():
    """models.UNet3Dv1m2 against the reference's own UNet3D-v1m2 (tests/golden/unet3d_v1m2.npz: parameters, input, output
    and parameter gradients of a tiny instance): same parameter names (strict load), same numbers (pure torch, CPU)."""
    from pathlib import Path
    from ponderv2_b200.models import MODELS
    z = np.load(Path(__file__).resolve().parent / "golden" / "unet3d_v1m2.npz")
    net = MODELS.build(dict(type="UNet3D-v1m2", in_channels=8, out_channels=6, f_maps=4, num_levels=3)).train()
    sd = {k[len("param."):]: torch.from_numpy(z[k]) for k in z.files if k.startswith("param.")}
    net.load_state_dict(sd, strict=True)
    y = net(torch.from_numpy(z["x"]))
    assert (y.detach() - torch.from_numpy(z["y"])).abs().max().item() < 1e-5
    y.square().mean().backward()
    for k, p in net.named_parameters():
        ref = torch.from_numpy(z["grad." + k])
        assert (p.grad - ref).abs().max().item() < 1e-5 * max(1.0, ref.abs().max().item()), k


def test_registry_models_match_reference_state_dicts():
    """`MODELS.build` of the reference's config dicts (PonderIndoor-v2 with the ScanNet config's UNet3D-v1m2 projection,
    PonderOutdoor-v2 with the nuScenes config's) gives modules whose state_dict names and shapes equal the reference
    classes' (tests/golden/ponder_models_state.json, oracle/gen_golden.py::gen_model_contracts): checkpoints interchange."""
    import json
    from pathlib import Path
    from oracle.gen_golden import INDOOR_MODEL_CFG, OUTDOOR_MODEL_CFG
    from ponderv2_b200.models import MODELS
    from tests.golden_util import product_renderer_cfg
    want = json.loads((Path(__file__).resolve().parent / "golden" / "ponder_models_state.json").read_text())
    for name, cfg, meta in (("PonderIndoor-v2", INDOOR_MODEL_CFG, dict(kind="indoor", S0=96, Si=36)),
                            ("PonderOutdoor-v2", OUTDOOR_MODEL_CFG, dict(kind="outdoor", S0=72, Si=24))):
        m = MODELS.build(dict(cfg, renderer=product_renderer_cfg(meta)))
        got = {k: list(v.shape) for k, v in m.state_dict().items()}
        assert set(got) == set(want[name]), (name, sorted(set(got) ^ set(want[name]))[:10])
        assert all(got[k] == want[name][k] for k in got), name
        assert len(m.grad_completion_order()) == len(list(m.parameters()))


def test_sdf_mlp_adjoint_math():
    """The hand-derived double backward of the fused SDF decoder (csrc/render_mlp.cu, render/mlp.py), restated in torch
    and checked against autograd in fp64: (sdf, u = d sdf/d f, v = d sdf/d p) and the gradient of an arbitrary scalar
    of the three with respect to f and every decoder parameter."""
    torch.manual_seed(0)
    dt = torch.float64
    P, F, H, O, L, pf = 50, 32, 16, 17, 6, 1.0
    Wp, bp = torch.randn(H, 3, dtype=dt) * 0.5, torch.randn(H, dtype=dt) * 0.1
    Fc = [torch.randn(H, F, dtype=dt) * 0.2 for _ in range(L)]
    bc = [torch.randn(H, dtype=dt) * 0.1 for _ in range(L)]
    W = [torch.randn(H if l < L - 1 else O, H, dtype=dt) * 0.3 for l in range(L)]
    b = [torch.randn(H if l < L - 1 else O, dtype=dt) * 0.1 for l in range(L)]
    params = [Wp, bp] + Fc + bc + W + b
    for t in params:
        t.requires_grad_(True)
    f = torch.randn(P, F, dtype=dt, requires_grad=True)
    p = torch.rand(P, 3, dtype=dt, requires_grad=True)
    sp = lambda y: torch.nn.functional.softplus(y, beta=100)
    x = pf * (p @ Wp.t() + bp)
    for l in range(L):
        y = (x + f @ Fc[l].t() + bc[l]) @ W[l].t() + b[l]
        if l < L - 1:
            x = sp(y)
    sdf = y[:, 0]
    u, v = torch.autograd.grad(sdf.sum(), [f, p], create_graph=True)
    gs, gu, gv = torch.randn(P, dtype=dt), torch.randn(P, F, dtype=dt), torch.randn(P, 3, dtype=dt)
    ref = torch.autograd.grad((sdf * gs).sum() + (u * gu).sum() + (v * gv).sum(), [f] + params)
    with torch.no_grad():
        x = pf * (p @ Wp.t() + bp)
        Z, S = [], []
        for l in range(L):
            z = x + f @ Fc[l].t() + bc[l]
            Z.append(z)
            if l < L - 1:
                y = z @ W[l].t() + b[l]
                S.append(torch.sigmoid(100 * y))
                x = sp(y)
        zbar, ybar = [None] * L, [None] * (L - 1)
        zbar[L - 1] = W[L - 1][0].expand(P, H).clone()
        um = zbar[L - 1] @ Fc[L - 1]
        for l in range(L - 2, -1, -1):
            ybar[l] = zbar[l + 1] * S[l]
            zbar[l] = ybar[l] @ W[l]
            um = um + zbar[l] @ Fc[l]
        vm = pf * (zbar[0] @ Wp)
        assert torch.allclose(um, u) and torch.allclose(vm, v)
        D, sbar = [None] * L, [None] * (L - 1)
        D[0] = gu @ Fc[0].t() + pf * (gv @ Wp.t())
        for l in range(L - 1):
            a = D[l] @ W[l].t()
            sbar[l] = a * zbar[l + 1]
            D[l + 1] = gu @ Fc[l + 1].t() + a * S[l]
        zhat = gs[:, None] * W[L - 1][0][None, :]
        fbar = zhat @ Fc[L - 1]
        A = [None] * (L - 1)
        for l in range(L - 2, -1, -1):
            A[l] = zhat * S[l] + sbar[l] * 100 * S[l] * (1 - S[l])
            zhat = A[l] @ W[l]
            fbar = fbar + zhat @ Fc[l]
        dW, db, dFc, dbc = [None] * L, [None] * L, [None] * L, [None] * L
        for l in range(L - 1):
            dW[l] = A[l].t() @ Z[l] + ybar[l].t() @ D[l]
            db[l] = A[l].sum(0)
            dFc[l] = W[l].t() @ (A[l].t() @ f + ybar[l].t() @ gu)
            dbc[l] = W[l].t() @ A[l].sum(0)
        dW[L - 1] = torch.zeros_like(W[L - 1]); dW[L - 1][0] = gs @ Z[L - 1] + D[L - 1].sum(0)
        db[L - 1] = torch.zeros_like(b[L - 1]); db[L - 1][0] = gs.sum()
        dFc[L - 1] = torch.outer(W[L - 1][0], gs @ f + gu.sum(0))
        dbc[L - 1] = W[L - 1][0] * gs.sum()
        dWp = pf * (W[0].t() @ (A[0].t() @ p + ybar[0].t() @ gv))
        dbp = pf * (W[0].t() @ A[0].sum(0))
        man = [fbar, dWp, dbp] + dFc + dbc + dW + db
    for i, (m_, r_) in enumerate(zip(man, ref)):
        e = ((m_ - r_).norm() / r_.norm().clamp(min=1e-30)).item()
        assert e < 1e-6, (i, e)


def test_sdf_mlp_pack_layout():
    """render.mlp.pack lays the decoder's parameters out as the kernels (pv2_sdf_mlp_param_count) expect."""
    from ponderv2_b200 import _lib
    from ponderv2_b200.render import mlp
    from ponderv2_b200.render.neus import SDFDecoder
    dec = SDFDecoder(in_dim=32, out_dim=17, hidden_size=16, n_blocks=5)
    packed, L, O, pf = mlp.pack(dec)
    assert (L, O, pf) == (6, 17, 1.0)
    assert packed.numel() == _lib.load().pv2_sdf_mlp_param_count(L, O) == 4881
    assert torch.equal(packed[:48], dec.fc_p.weight.detach().reshape(-1))
    assert torch.equal(packed[64:64 + 512], dec.fc_c[0].weight.detach().reshape(-1))
    assert torch.equal(packed[-17:], dec.lin5.bias.detach())
    packed.sum().backward()
    assert all(q.grad is not None for q in dec.parameters())


# ------------------------------------------------------------------------------------------ SpUNet-v1m3 (PDNorm, §8f-3)
def _pdnorm_case(tag):
    from pathlib import Path
    from ponderv2_b200.backbone_pdnorm import PDBatchNorm
    g = np.load(Path(__file__).parent / "golden" / "pdnorm.npz")
    kw = dict(adaptive=dict(decouple=True, adaptive=True, affine=False), affine=dict(decouple=True, adaptive=False, affine=True),
              both=dict(decouple=False, adaptive=True, affine=True))[tag]
    pd = PDBatchNorm(32, context_channels=16, conditions=("A", "B"), **kw).train()
    sd = {k[len(tag) + 7:]: torch.from_numpy(g[k]) for k in g.files if k.startswith(tag + ".param.")}
    pd.load_state_dict(sd)
    return g, kw, pd


def test_pdnorm_folding_matches_reference_golden():
    """The folding of the context modulation into the BatchNorm affine pair (bn_act.bn_act_modulated; here its torch
    branch, the algebra the fused kernels are handed) against the reference PDBatchNorm's output, running buffers and
    gradients (tests/golden/pdnorm.npz, oracle/gen_golden.py:gen_pdnorm)."""
    for tag in ("adaptive", "affine", "both"):
        g, kw, pd = _pdnorm_case(tag)
        x = torch.from_numpy(g[f"{tag}.x"]).requires_grad_(True)
        ctx = torch.from_numpy(g[f"{tag}.ctx"]).requires_grad_(True)
        y = pd(x, "B", ctx if kw["adaptive"] else None)
        (y * torch.from_numpy(g[f"{tag}.go"])).sum().backward()
        np.testing.assert_allclose(y.detach().numpy(), g[f"{tag}.y"], rtol=1e-5, atol=2e-5)
        np.testing.assert_allclose(x.grad.numpy(), g[f"{tag}.dx"], rtol=1e-4, atol=2e-5)
        if kw["adaptive"]:
            np.testing.assert_allclose(ctx.grad.numpy(), g[f"{tag}.dctx"], rtol=1e-4, atol=1e-3)
        for k, p in pd.named_parameters():
            key = f"{tag}.grad.{k}"
            if key in g.files:
                np.testing.assert_allclose(p.grad.numpy(), g[key], rtol=1e-4, atol=2e-3)
        for k, v in pd.state_dict().items():
            if "running" in k:
                np.testing.assert_allclose(v.numpy(), g[f"{tag}.after.{k}"], rtol=1e-5, atol=1e-6)


def test_spunet_v1m3_state_dict_contract():
    """Parameter / buffer names and shapes equal the reference's SpUNet-v1m3 (PPT checkpoints load unchanged), through
    the registry name."""
    import json
    from pathlib import Path
    from ponderv2_b200.models import MODELS
    want = json.loads((Path(__file__).parent / "golden" / "spunet_v1m3_state.json").read_text())["state"]
    m = MODELS.build(dict(type="SpUNet-v1m3", in_channels=6, num_classes=0,
                          conditions=("ScanNet", "S3DIS", "Structured3D")))
    got = {k: list(v.shape) for k, v in m.state_dict().items()}
    assert got == want
    # zero_init: the modulation starts as the identity
    assert all(float(p.abs().max()) == 0.0 for k, p in m.named_parameters() if "modulation" in k)


def test_semantic_loss_matches_reference_golden():
    """NeuSModel._semantic_loss (§8f-4: sync-free masked contrastive cross entropy; eval = mean over val_ray_split
    chunks) on the reference's own rendered features against the reference's loss value (base_surface_model.py:123-173,
    tests/golden/render_indoor_{train,eval}_semantic.npz)."""
    from ponderv2_b200.render import build_renderer
    from tests.golden_util import load_render_case, product_renderer_cfg
    for case in ("indoor_train_semantic", "indoor_eval_semantic"):
        meta, arr, sd, _ = load_render_case(case)
        m = build_renderer(product_renderer_cfg(meta))
        m.load_state_dict(sd, strict=True)
        m.train(meta["training"])
        got = m._semantic_loss({"semantic": arr["out.semantic"]}, {"semantic": arr["semantic_gt"], "depth": arr["depth_gt"]})
        ref = float(arr["loss.semantic_loss"])
        assert abs(float(got) * 0.1 - ref) < 1e-5 * max(1.0, abs(ref)), (case, float(got) * 0.1, ref)
    # all rays ignored -> exactly 0 (the reference's explicit branch), no NaN
    z = m._semantic_loss({"semantic": arr["out.semantic"]},
                         {"semantic": torch.zeros_like(arr["semantic_gt"]), "depth": arr["depth_gt"]})
    assert float(z) == 0.0


def test_semantic_and_pdnorm_models_construct_through_registry():
    """`PonderIndoor-v2` with the semantic branch (class embeddings handed in, PPT point loss) and with the SpUNet-v1m3
    backbone builds from config dicts through the registry and exposes the reference's extra state (ponder_indoor_base.py:
    66-118: `embedding_table`, `class_embedding`, `logit_scale`, `proj_head`)."""
    from oracle.gen_golden import INDOOR_MODEL_CFG
    from ponderv2_b200.models import MODELS
    from tests.golden_util import product_renderer_cfg
    rcfg = product_renderer_cfg(dict(kind="indoor", S0=96, Si=36, semantic=24))
    cfg = dict(INDOOR_MODEL_CFG, renderer=rcfg, render_semantic=True, class_embedding=torch.randn(13, 24),
               conditions=("ScanNet", "S3DIS"), valid_index=(tuple(range(13)), tuple(range(5))), ppt_loss_weight=1.0,
               ppt_criteria=[dict(type="CrossEntropyLoss", loss_weight=1.0, ignore_index=-1)])
    cfg["backbone"] = dict(type="SpUNet-v1m3", in_channels=6, num_classes=0, conditions=("ScanNet", "S3DIS"))
    m = MODELS.build(cfg)
    sd = m.state_dict()
    assert type(m.backbone).__name__ == "SpUNetPDNorm"
    for key in ("class_embedding", "logit_scale", "proj_head.weight", "embedding_table.weight",
                "renderer.field.semantic_decoder.lin0.weight"):
        assert key in sd, key
    assert sd["class_embedding"].shape == (13, 24) and sd["proj_head.weight"].shape == (24, 96)
    assert abs(float(sd["class_embedding"].norm(dim=-1).mean()) - 1.0) < 1e-5          # unit-norm rows, as load_semantic
    table = m._class_table({"condition": ["S3DIS"]})
    assert table.shape == (5, 24)
    # without embeddings (and without CLIP in this image) construction must fail loudly, not silently skip the branch
    bad = dict(cfg); bad.pop("class_embedding")
    try:
        MODELS.build(bad)
    except RuntimeError as e:
        assert "class_embedding" in str(e)
    else:
        try:
            import clip  # noqa: F401
        except ImportError:
            raise AssertionError("render_semantic=True without embeddings must raise when CLIP is absent")


def test_traffic_tool_books_render_linears_separately(tmp_path):
    """tools/traffic_from_launches.py on the committed ncu launch list: the render-MLP linear launches (same kernel as the
    sparse convolutions) are told apart by their position between the field kernels, so that bench.py's roofline.traffic
    is the sparse convolutions' own DRAM traffic."""
    import json
    import subprocess
    import sys
    from pathlib import Path
    root = Path(__file__).resolve().parent.parent
    src = root / "profiles" / "r2zd_launches.csv"
    out = tmp_path / "t.json"
    subprocess.run([sys.executable, str(root / "tools" / "traffic_from_launches.py"), str(src), str(out)], check=True,
                   capture_output=True)
    k = json.loads(out.read_text())["kernels"]
    lin = [v for n, v in k.items() if "[render linear]" in n]
    assert sum(v["launches"] for v in lin) == 9                      # 2 coarse + 3 fine forward, 4 backward layers per step
    conv = [v for n, v in k.items() if n.startswith("umma_gather_gemm") and "[render linear]" not in n]
    per_launch = sum(v["dram_bytes_total"] for v in conv) / sum(v["launches"] for v in conv)
    assert 5e6 < per_launch < 17.1e6                                 # below the 17.0 MB algorithmic bytes: L2 reuse
