"""GPU parity of the B200 kernels against the CPU oracle, through the C ABI (ctypes) and the B1/B2 shims.

Integer results (rulebooks) are compared bit-exactly in canonical form; floating point within the stated
tolerances: fp32 kernels <= 2e-5 relative to the fp64 oracle's magnitude (fp32 accumulation-order noise), bf16
storage <= 2e-2.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import densify_oracle as do
from oracle import spconv_oracle as so
from oracle.trilinear_oracle import trilinear_sample
from ponderv2_b200 import synth


def _dev():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch.device("cuda:0")


def _random_cloud(n, extent, batch, seed):
    rng = np.random.default_rng(seed)
    pts = set()
    while len(pts) < n:
        pts.add((int(rng.integers(0, batch)), *[int(v) for v in rng.integers(0, extent, size=3)]))
    arr = np.array(sorted(pts), dtype=np.int32).reshape(-1, 4)
    rng.shuffle(arr)
    return arr


def _indoor_indices(n, seed):
    c = synth.indoor_cloud(n, seed)
    ind = np.concatenate([np.zeros((n, 1), np.int64), c["grid_coord"]], 1).astype(np.int32)
    return ind, (c["grid_coord"].max(0) + 96).tolist()


# ------------------------------------------------------------------------------------------ rulebooks
@pytest.mark.parametrize("ksize", [3, 5])
@pytest.mark.parametrize("case", ["random_b2", "indoor_5k", "single", "dense_block"])
def test_subm_rulebook_bit_exact(cuda_lib, ksize, case):
    from ponderv2_b200.spconv.pytorch import build_subm_rulebook
    if case == "random_b2":
        ind, shape = _random_cloud(3000, 24, 2, 5), [30, 30, 30]
    elif case == "indoor_5k":
        ind, shape = _indoor_indices(5000, 7)
    elif case == "single":
        ind, shape = np.array([[0, 3, 4, 5]], dtype=np.int32), [8, 8, 8]
    else:  # every site occupied, touching the boundary of the spatial shape
        g = np.stack(np.meshgrid(np.arange(6), np.arange(5), np.arange(4), indexing="ij"), -1).reshape(-1, 3)
        ind, shape = np.concatenate([np.zeros((g.shape[0], 1)), g], 1).astype(np.int32), [6, 5, 4]
    rb = build_subm_rulebook(torch.from_numpy(ind).to(_dev()), shape, ksize)
    ref = so.subm_rulebook(ind, shape, ksize)
    got = rb.nbr.cpu().numpy()
    assert got.dtype == np.int32 and got.shape == ref.shape
    assert np.array_equal(got, ref)
    assert rb.num_pairs == int((ref >= 0).sum())


def test_subm_rulebook_empty_and_duplicates(cuda_lib):
    from ponderv2_b200.spconv.pytorch import build_subm_rulebook
    empty = torch.zeros((0, 4), dtype=torch.int32, device=_dev())
    rb = build_subm_rulebook(empty, [4, 4, 4], 3)
    assert rb.nbr.shape == (27, 0)
    dup = np.array([[0, 1, 1, 1], [0, 1, 1, 2], [0, 1, 1, 1]], dtype=np.int32)  # rows 0 and 2 collide
    rb = build_subm_rulebook(torch.from_numpy(dup).to(_dev()), [4, 4, 4], 3)
    assert np.array_equal(rb.nbr.cpu().numpy(), so.subm_rulebook(dup, [4, 4, 4], 3))


@pytest.mark.parametrize("case", ["random_b3", "indoor_20k", "odd_shape"])
def test_down_rulebook_canonical_bit_exact(cuda_lib, case):
    from ponderv2_b200.spconv.pytorch import build_down_rulebook
    if case == "random_b3":
        ind, shape = _random_cloud(4000, 21, 3, 9), [24, 24, 24]
    elif case == "indoor_20k":
        ind, shape = _indoor_indices(20000, 3)
    else:
        ind, shape = _random_cloud(500, 9, 1, 4), [9, 9, 9]  # last plane has no complete window -> dropped
    rb = build_down_rulebook(torch.from_numpy(ind).to(_dev()), shape)
    ref_out, ref_i2o, ref_koff, ref_shape = so.down_rulebook(ind, shape)
    assert rb.out_shape == ref_shape
    out = rb.out_indices.cpu().numpy()
    i2o = rb.in2out.cpu().numpy()
    # implementation order: outputs numbered by first contributing input row
    firsts = {}
    for i, o in enumerate(i2o.tolist()):
        if o >= 0:
            firsts.setdefault(o, i)
    order = sorted(firsts, key=lambda o: firsts[o])
    assert order == list(range(len(order)))
    can_out, can_i2o, perm = so.canonical_down(out, i2o, ref_shape)
    assert np.array_equal(can_out, ref_out)
    assert np.array_equal(can_i2o, ref_i2o)
    assert np.array_equal(rb.koff.cpu().numpy(), ref_koff)
    nd, nu = so.down_maps(i2o, ref_koff, out.shape[0])
    assert np.array_equal(rb.nbr_down.cpu().numpy(), nd)
    assert np.array_equal(rb.nbr_up.cpu().numpy(), nu)


@pytest.mark.parametrize("case", ["subm3_indoor", "down_up"])
def test_row_order_is_stable_mask_sort(cuda_lib, case):
    """pv2_rulebook_row_order == numpy stable argsort of the neighbour-presence masks (bit-exact), and is a permutation."""
    import ponderv2_b200.spconv.pytorch as spconv
    ind, shape = _indoor_indices(7000, 77)
    ind_t = torch.from_numpy(ind).to(_dev())
    if case == "subm3_indoor":
        maps = [spconv.build_subm_rulebook(ind_t, shape, 3).nbr]
    else:
        d = spconv.build_down_rulebook(ind_t, shape)
        maps = [d.nbr_down, d.nbr_up]
    for nbr in maps:
        order = spconv.build_row_order(nbr).cpu().numpy()
        m = nbr.cpu().numpy() >= 0
        mask = np.zeros(m.shape[1], np.int64)
        for k in range(m.shape[0]):
            mask |= m[k].astype(np.int64) << k
        want = np.argsort(mask, kind="stable").astype(np.int32)
        assert np.array_equal(order, want)
        tm = spconv.build_tile_map(nbr)
        nbr_np = nbr.cpu().numpy()
        assert np.array_equal(tm.nbr.cpu().numpy(), nbr_np[:, want])            # the map in tile order
        nblk = (nbr_np.shape[1] + 31) // 32
        padded = np.full((nbr_np.shape[0], nblk * 32), -1, np.int32)
        padded[:, :nbr_np.shape[1]] = nbr_np[:, want]
        assert np.array_equal(tm.blk_active.cpu().numpy(), (padded.reshape(nbr_np.shape[0], nblk, 32) >= 0).any(2))
    assert spconv.build_row_order(torch.zeros((200, 10), dtype=torch.int32, device=_dev())) is None  # K > 128: no order
    assert spconv.build_row_order(torch.zeros((125, 10), dtype=torch.int32, device=_dev())) is not None  # 5x5x5: folded key


def test_make_indices(cuda_lib):
    from ponderv2_b200.backbone import make_sparse_indices
    gc = torch.randint(0, 50, (1000, 3), dtype=torch.int64)
    offset = torch.tensor([100, 100, 640, 1000], dtype=torch.int64)  # an empty scene in the middle
    got = make_sparse_indices(gc.to(_dev()), offset.to(_dev())).cpu()
    batch = torch.from_numpy(so.offset2batch(offset.numpy()))
    assert torch.equal(got, torch.cat([batch[:, None].int(), gc.int()], 1))


# ------------------------------------------------------------------------------------------ sparse conv arithmetic
def _conv_case(kind, cin, cout, dtype, seed, n=1500):
    """Runs one conv forward+backward on GPU (product) and CPU (oracle, fp64); returns max relative errors."""
    import ponderv2_b200.spconv.pytorch as spconv
    dev = _dev()
    torch.manual_seed(seed)
    ind = _random_cloud(n, 16, 2, seed)
    shape = [20, 20, 20]
    x64 = torch.randn(n, cin, dtype=torch.float64)
    if kind == "subm3":
        mod = spconv.SubMConv3d(cin, cout, 3, padding=1, bias=True, indice_key="k")
    elif kind == "subm5":
        mod = spconv.SubMConv3d(cin, cout, 5, padding=1, bias=False, indice_key="k")
    elif kind == "subm1":
        mod = spconv.SubMConv3d(cin, cout, 1, bias=False)
    elif kind in ("down", "updown"):
        mod = spconv.SparseConv3d(cin, cout, 2, stride=2, bias=False, indice_key="d")
    mod = mod.to(dev)
    w64 = mod.weight.detach().cpu().double().requires_grad_(True)
    b64 = mod.bias.detach().cpu().double().requires_grad_(True) if mod.bias is not None else None
    xg = x64.to(dev, dtype).requires_grad_(True)
    xo = x64.clone()
    if dtype == torch.bfloat16:  # compare against the oracle fed with the same rounded inputs
        xo = xg.detach().cpu().double()
        w64 = mod.weight.detach().to(torch.bfloat16).cpu().double().requires_grad_(True)
    xo.requires_grad_(True)
    st = spconv.SparseConvTensor(xg, torch.from_numpy(ind).to(dev), shape, 2)
    ctx = torch.autocast("cuda", dtype=torch.bfloat16) if dtype == torch.bfloat16 else torch.autocast("cuda", enabled=False)
    with ctx:
        out = mod(st)
    yo_t = None
    ot = so.OracleSparseTensor(xo, ind, shape, 2)
    if kind.startswith("subm"):
        ks = int(kind[-1])
        yo = so.subm_conv(ot, w64, b64, ks, "k").features
        yg = out.features
    else:
        o_down = so.down_conv(ot, w64, None, "d")
        # bring product rows into canonical order
        _, _, perm = so.canonical_down(out.indices.cpu().numpy(), np.zeros(0, np.int32), o_down.spatial_shape)
        perm_t = torch.from_numpy(perm).to(dev)
        yg = out.features[perm_t]
        yo = o_down.features
        if kind == "updown":
            inv = spconv.SparseInverseConv3d(cout, cin, 2, indice_key="d", bias=False).to(dev)
            wi64 = inv.weight.detach().cpu().double().requires_grad_(True)
            if dtype == torch.bfloat16:
                wi64 = inv.weight.detach().to(torch.bfloat16).cpu().double().requires_grad_(True)
            with ctx:
                out2 = inv(out)
            yg = out2.features
            yo = so.inverse_conv(o_down, wi64, None, "d").features
            assert torch.equal(out2.indices.cpu(), torch.from_numpy(ind))
    g64 = torch.randn(yo.shape, dtype=torch.float64)
    yo.backward(g64)
    yg.backward(g64.to(dev, yg.dtype))
    scale = lambda t: max(t.abs().max().item(), 1e-6)
    errs = {
        "y": (yg.detach().cpu().double() - yo.detach()).abs().max().item() / scale(yo),
        "dx": (xg.grad.cpu().double() - xo.grad).abs().max().item() / scale(xo.grad),
        "dw": (mod.weight.grad.cpu().double() - w64.grad).abs().max().item() / scale(w64.grad),
    }
    if b64 is not None and dtype == torch.float32:
        errs["db"] = (mod.bias.grad.cpu().double() - b64.grad).abs().max().item() / scale(b64.grad)
    return errs


@pytest.mark.parametrize("kind,cin,cout", [
    ("subm3", 32, 32), ("subm3", 64, 96), ("subm3", 192, 128), ("subm5", 6, 32), ("subm1", 128, 96),
    ("subm3", 384, 256),   # dec3: the 384-wide data gradient / weight gradient run as two column slices
    ("subm3", 7, 13), ("down", 32, 64), ("updown", 64, 96), ("subm5", 4, 32),
])
def test_sparse_conv_fp32(cuda_lib, kind, cin, cout):
    errs = _conv_case(kind, cin, cout, torch.float32, seed=cin + cout)
    # fp32 storage runs on the tensor cores as 3xTF32 (hi*hi + lo*hi + hi*lo, TF32 halves rounded to nearest):
    # measured <= 4e-5 of the output magnitude at K*Cin = 6912; the exact-fp32 SIMT kernel gives <= 4e-6
    for k, v in errs.items():
        assert v < 1e-4, (k, v, errs)


@pytest.mark.parametrize("kind,cin,cout", [("subm3", 32, 32), ("subm3", 96, 96), ("down", 32, 64), ("updown", 64, 32)])
def test_sparse_conv_bf16(cuda_lib, kind, cin, cout):
    errs = _conv_case(kind, cin, cout, torch.bfloat16, seed=cin + cout)
    for k, v in errs.items():
        assert v < 2e-2, (k, v, errs)


def test_spunet_backbone_matches_oracle(cuda_lib):
    """Whole SpUNet-v1m1 (59 convs, 10 rulebooks) on an 8 k-voxel two-scene batch: features and parameter grads."""
    from ponderv2_b200.backbone import SpUNetBase
    dev = _dev()
    torch.manual_seed(0)
    a, b = synth.indoor_cloud(4800, 21), synth.indoor_cloud(3200, 22)
    gc = np.concatenate([a["grid_coord"], b["grid_coord"]])
    feat = np.concatenate([a["feat"], b["feat"]])
    offset = np.array([4800, 8000], dtype=np.int64)
    model = SpUNetBase(in_channels=6, num_classes=0).to(dev).train()
    sd = {k: v.detach().cpu().double() for k, v in model.state_dict().items()}
    for k, v in sd.items():
        if v.is_floating_point() and ("weight" in k or "bias" in k):
            v.requires_grad_(True)
    out = model({"grid_coord": torch.from_numpy(gc).to(dev), "feat": torch.from_numpy(feat).to(dev),
                 "offset": torch.from_numpy(offset).to(dev)})
    ref = so.spunet_forward(sd, gc, torch.from_numpy(feat).double(), offset)
    assert out.shape == (8000, 96)
    err = (out.detach().cpu().double() - ref.detach()).abs().max().item() / ref.abs().max().item()
    print('backbone fwd rel err', err)
    assert err < 5e-4, err  # 59 layers of 3xTF32 convs + train-mode BN
    g = torch.randn(ref.shape, dtype=torch.float64, generator=torch.Generator().manual_seed(1))
    ref.backward(g)
    out.backward(g.to(dev, torch.float32))
    # gradients: all parameters jointly (relative L2) and the worst single tensor.  Train-mode BN divides by batch
    # statistics of a few hundred rows at the deepest levels, which amplifies the 3xTF32 rounding of the convolutions
    # (the same test with PV2_FORCE_SIMT=1, i.e. exact fp32 accumulation, gives 2e-5 / 5e-3).
    num = den = 0.0
    worst, worst_name = 0.0, ""
    for name, p in model.named_parameters():
        rg = sd[name].grad
        d = (p.grad.cpu().double() - rg)
        num += d.pow(2).sum().item()
        den += rg.pow(2).sum().item()
        e = d.norm().item() / max(rg.norm().item(), 1e-9)
        if e > worst:
            worst, worst_name = e, name
    joint = (num / den) ** 0.5
    from tests.conftest import record
    record("spunet_backbone_matches_oracle", fwd=err, grad_joint=joint, grad_worst=worst, worst_name=worst_name)
    print("backbone grad err: joint", joint, "worst", worst_name, worst)
    # measured on B200 (profiles/r2d_parity_report.jsonl): joint 6.6e-3, worst tensor 1.2e-2 (enc.0.block1.bn2.weight)
    assert joint < 1.5e-2, joint
    assert worst < 4e-2, (worst_name, worst)


def test_spunet_state_dict_contract(cuda_lib):
    """Parameter/buffer names and shapes equal the reference's SpUNet-v1m1 (checkpoint compatibility)."""
    import json
    from pathlib import Path
    from ponderv2_b200.backbone import SpUNetBase
    want = json.loads((Path(__file__).parent / "golden" / "spunet_v1m1_state.json").read_text())["state"]
    got = {k: list(v.shape) for k, v in SpUNetBase(in_channels=6, num_classes=0).state_dict().items()}
    assert got == want


# ------------------------------------------------------------------------------------------ SpUNet-v1m3 (PDNorm, §8f-3)
@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["adaptive", "affine", "both"])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_pdnorm_fused_matches_reference_golden(cuda_lib, tag, dtype):
    """PDBatchNorm on the fused BN kernels (context modulation folded into the affine pair) against the reference
    module's output, running buffers and gradients (tests/golden/pdnorm.npz, generated by the imported reference)."""
    from pathlib import Path
    from ponderv2_b200.backbone_pdnorm import PDBatchNorm
    dev = _dev()
    g = np.load(Path(__file__).parent / "golden" / "pdnorm.npz")
    kw = dict(adaptive=dict(decouple=True, adaptive=True, affine=False), affine=dict(decouple=True, adaptive=False, affine=True),
              both=dict(decouple=False, adaptive=True, affine=True))[tag]
    pd = PDBatchNorm(32, context_channels=16, conditions=("A", "B"), **kw).train()
    pd.load_state_dict({k[len(tag) + 7:]: torch.from_numpy(g[k]) for k in g.files if k.startswith(tag + ".param.")})
    pd = pd.to(dev)
    x = torch.from_numpy(g[f"{tag}.x"]).to(dev, dtype).requires_grad_(True)
    ctx = torch.from_numpy(g[f"{tag}.ctx"]).to(dev).requires_grad_(True)
    y = pd(x, "B", ctx if kw["adaptive"] else None)
    assert y.dtype == dtype
    (y.float() * torch.from_numpy(g[f"{tag}.go"]).to(dev)).sum().backward()
    lo = dtype == torch.bfloat16
    rel = lambda a, b: float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-12))
    errs = dict(y=rel(y.detach().float().cpu().numpy(), g[f"{tag}.y"]), dx=rel(x.grad.float().cpu().numpy(), g[f"{tag}.dx"]))
    if kw["adaptive"]:
        errs["dctx"] = rel(ctx.grad.cpu().numpy(), g[f"{tag}.dctx"])
    for k, p in pd.named_parameters():
        if f"{tag}.grad.{k}" in g.files:
            errs["grad." + k] = rel(p.grad.cpu().numpy(), g[f"{tag}.grad.{k}"])
    for k, v in pd.state_dict().items():
        if "running" in k:
            errs["after." + k] = rel(v.cpu().numpy(), g[f"{tag}.after.{k}"])
    from tests.conftest import record
    record("pdnorm_fused", tag=tag, dtype=str(dtype), **errs)
    tol = 1.2e-2 if lo else 2e-5     # bf16: input and output rounding (2^-9 each) of unit-scale features
    for k, e in errs.items():
        assert e < tol, (k, e, errs)


@pytest.mark.gpu
def test_spunet_v1m3_matches_oracle(cuda_lib):
    """Whole SpUNet-v1m3 with a NON-trivial context modulation on a 6 k-voxel two-scene batch.  Per norm layer the
    modulation is a per-channel affine map, so the network equals SpUNet-v1m1 with BatchNorm weight = 1 + scale and
    bias = shift: the fp64 oracle (oracle/spconv_oracle.py:spunet_forward) is run on that translated state dict and
    compared in output, conv-weight gradients and modulation-bias gradients (= [d shift, d scale] = the oracle's
    [d bias, d weight])."""
    from ponderv2_b200.backbone_pdnorm import PDBatchNorm, SpUNetPDNorm
    dev = _dev()
    torch.manual_seed(5)
    a, b = synth.indoor_cloud(3600, 31), synth.indoor_cloud(2400, 32)
    gc = np.concatenate([a["grid_coord"], b["grid_coord"]])
    feat = np.concatenate([a["feat"], b["feat"]])
    offset = np.array([3600, 6000], dtype=np.int64)
    model = SpUNetPDNorm(in_channels=6, num_classes=0, context_channels=24, conditions=("ScanNet", "S3DIS"),
                         zero_init=False).train()
    with torch.no_grad():
        for k, p in model.named_parameters():
            if "modulation" in k:
                p.copy_(torch.randn_like(p) * 0.15)
    model = model.to(dev)
    context = torch.randn(1, 24)
    # translate to the v1m1 oracle's state dict
    names = {"conv_input.conv.": "conv_input.0.", "conv_input.bn": "conv_input.1"}
    def v1(k):
        k = k.replace("conv_input.conv.", "conv_input.0.").replace(".proj_conv.", ".proj.0.")
        for pre in ("down", "up"):
            for s in range(4):
                k = k.replace(f"{pre}.{s}.conv.", f"{pre}.{s}.0.")
        return k
    sd, norm_of = {}, {}
    for k, p in model.named_parameters():
        if k.endswith(".weight") and p.dim() == 5:
            sd[v1(k)] = p.detach().cpu().double().requires_grad_(True)
    for name, m in model.named_modules():
        if isinstance(m, PDBatchNorm):
            shift, scale = m.modulation(context.to(dev)).detach().cpu().double().chunk(2, dim=1)
            tgt = name.replace("conv_input.bn", "conv_input.1").replace(".proj_norm", ".proj.1")
            for pre in ("down", "up"):
                for s in range(4):
                    if tgt == f"{pre}.{s}.bn":
                        tgt = f"{pre}.{s}.1"
            sd[tgt + ".weight"] = (1.0 + scale[0]).requires_grad_(True)
            sd[tgt + ".bias"] = shift[0].clone().requires_grad_(True)
            norm_of[name] = tgt
    out = model({"grid_coord": torch.from_numpy(gc).to(dev), "feat": torch.from_numpy(feat).to(dev),
                 "offset": torch.from_numpy(offset).to(dev), "condition": ["S3DIS"], "context": context.to(dev)})
    ref = so.spunet_forward(sd, gc, torch.from_numpy(feat).double(), offset)
    err = (out.detach().cpu().double() - ref.detach()).abs().max().item() / ref.abs().max().item()
    assert err < 5e-4, err
    gr = torch.randn(ref.shape, dtype=torch.float64, generator=torch.Generator().manual_seed(1))
    ref.backward(gr)
    out.backward(gr.to(dev, torch.float32))
    num = den = 0.0
    for k, p in model.named_parameters():
        if k.endswith(".weight") and p.dim() == 5:
            rg = sd[v1(k)].grad
        elif k.endswith("modulation.1.bias"):
            tgt = norm_of[k[:-len(".modulation.1.bias")]]
            rg = torch.cat([sd[tgt + ".bias"].grad, sd[tgt + ".weight"].grad])
        else:
            continue
        d = p.grad.cpu().double() - rg
        num += d.pow(2).sum().item(); den += rg.pow(2).sum().item()
    joint = (num / den) ** 0.5
    # the unused condition's BatchNorms must be untouched (decoupled statistics)
    bn_other = model.conv_input.bn.bns[0]
    assert int(bn_other.num_batches_tracked) == 0 and int(model.conv_input.bn.bns[1].num_batches_tracked) == 1
    from tests.conftest import record
    record("spunet_v1m3_matches_oracle", fwd=err, grad_joint=joint)
    assert joint < 1.5e-2, joint


# ------------------------------------------------------------------------------------------ densify
def test_densify_indoor_and_outdoor(cuda_lib):
    from ponderv2_b200 import densify
    dev = _dev()
    torch.manual_seed(0)
    a, b = synth.indoor_cloud(4000, 31), synth.indoor_cloud(2500, 32)
    coord = torch.from_numpy(np.concatenate([a["coord"], b["coord"]]))
    offset = np.array([4000, 6500])
    feat = torch.randn(6500, 24, dtype=torch.float32)
    grid_shape, grid_size = (16, 16, 8), 0.02
    res = torch.tensor([int(a["grid_coord"].max()), int(b["grid_coord"].max())])
    ref_in = feat.clone().double().requires_grad_(True)
    ref = do.to_dense_indoor(coord, ref_in, offset, res, grid_shape, grid_size)  # coord stays fp32 as in the reference
    batch = torch.from_numpy(so.offset2batch(offset)).to(dev)
    f = feat.to(dev).requires_grad_(True)
    cell = densify.indoor_cells(coord.to(dev), batch, res, grid_shape, grid_size)
    vol = densify.scatter_mean_volume(f, cell, 2, (grid_shape[2], grid_shape[1], grid_shape[0]))
    assert vol.shape == ref.shape and vol.is_contiguous(memory_format=torch.channels_last_3d)
    assert (vol.detach().cpu().double() - ref.detach()).abs().max().item() < 1e-5
    g = torch.randn(ref.shape, dtype=torch.float64)
    ref.backward(g)
    vol.backward(g.to(dev, torch.float32))
    assert (f.grad.cpu().double() - ref_in.grad).abs().max().item() < 1e-5
    # outdoor
    c = synth.outdoor_cloud(5000, 33)
    bbox, gsz, gshape = [0, 0, 0, 108, 108, 8], [0.6, 0.6, 1.6], [180, 180, 5]
    fo = torch.randn(5000, 16)
    ref_o = do.to_dense_outdoor(torch.from_numpy(c["coord"]), fo, np.array([5000]), bbox, gsz, gshape)
    cell = densify.outdoor_cells(torch.from_numpy(c["coord"]).to(dev), torch.zeros(5000, dtype=torch.int64, device=dev),
                                 bbox, gsz, gshape)
    vo = densify.scatter_mean_volume(fo.to(dev), cell, 1, (5, 180, 180))
    assert (vo.cpu() - ref_o).abs().max().item() < 1e-5


# ------------------------------------------------------------------------------------------ trilinear sampler (B2)
@pytest.mark.parametrize("padding_mode", ["zeros", "border", "reflection"])
@pytest.mark.parametrize("align_corners", [True, False])
def test_smooth_sampler_reference_kat(cuda_lib, padding_mode, align_corners):
    """The reference's own self-check, libs/smooth-sampler/smooth_sampler/modules.py:104-156, verbatim protocol."""
    from ponderv2_b200.smooth_sampler import SmoothSampler
    dev = _dev()
    torch.manual_seed(3)
    inp = torch.rand([2, 2, 2, 3, 11], device=dev).requires_grad_(True)
    grid = (torch.rand([2, 2, 1, 5, 3], device=dev) * 2.0 - 1.0).requires_grad_(True)
    out1 = SmoothSampler.apply(inp, grid, padding_mode, align_corners, False)
    out2 = torch.nn.functional.grid_sample(inp, grid, padding_mode=padding_mode, align_corners=align_corners)
    assert torch.allclose(out1, out2)
    g1 = torch.autograd.grad(out1, [inp, grid], torch.ones_like(out1), create_graph=True)
    g2 = torch.autograd.grad(out2, [inp, grid], torch.ones_like(out2), create_graph=True)
    assert torch.allclose(g1[0], g2[0]) and torch.allclose(g1[1], g2[1])
    for smooth in [True, False]:
        i64 = torch.rand([2, 2, 2, 3, 11], device=dev).double().requires_grad_(True)
        g64 = (torch.rand([2, 2, 1, 5, 3], device=dev) * 2.0 - 1.0).double().requires_grad_(True)
        torch.autograd.gradcheck(SmoothSampler.apply, [i64, g64, padding_mode, align_corners, smooth],
                                 eps=1e-4, atol=1e-3, rtol=1e-2)
        torch.autograd.gradgradcheck(SmoothSampler.apply, [i64, g64, padding_mode, align_corners, smooth],
                                     eps=1e-4, atol=1e-3, rtol=1e-2)


@pytest.mark.parametrize("smooth", [False, True])
def test_smooth_sampler_double_backward_vs_oracle(cuda_lib, smooth):
    """Second-order path as the renderer uses it: d/dtheta of a loss on (features, d features/d points)."""
    from ponderv2_b200.smooth_sampler import SmoothSampler
    dev = _dev()
    torch.manual_seed(5)
    vol = torch.randn(1, 8, 5, 6, 7, dtype=torch.float64)
    grid = torch.rand(1, 1, 9, 13, 3, dtype=torch.float64) * 2.6 - 1.3  # some points outside: zeros padding
    wmix = torch.randn(8, dtype=torch.float64)

    def run(sampler, v, g):
        v = v.clone().requires_grad_(True)
        g = g.clone().requires_grad_(True)
        f = sampler(v, g)                                        # (1,C,1,R,S)
        s = (f * wmix.to(f.device).view(1, -1, 1, 1, 1)).sum(1).tanh()
        dg = torch.autograd.grad(s.sum(), g, create_graph=True)[0]
        loss = (dg.norm(dim=-1) - 1).pow(2).mean() + s.pow(2).mean() + (f ** 2).mean()
        loss.backward()
        return loss.detach().cpu(), v.grad.cpu(), g.grad.cpu()

    lo, gvo, ggo = run(lambda v, g: trilinear_sample(v, g, "zeros", True, smooth), vol, grid)
    lg, gvg, ggg = run(lambda v, g: SmoothSampler.apply(v, g, "zeros", True, smooth), vol.to(dev), grid.to(dev))
    assert abs(lo.item() - lg.item()) < 1e-10
    assert (gvo - gvg).abs().max().item() < 1e-10
    assert (ggo - ggg).abs().max().item() < 1e-9


def test_smooth_sampler_rejects_cpu_and_noncontiguous(cuda_lib):
    from ponderv2_b200.smooth_sampler import SmoothSampler
    with pytest.raises(RuntimeError):
        SmoothSampler.apply(torch.rand(1, 2, 3, 3, 3), torch.rand(1, 1, 1, 4, 3), "zeros", True, False)
    dev = _dev()
    with pytest.raises(RuntimeError):
        SmoothSampler.apply(torch.rand(1, 2, 3, 3, 6, device=dev)[..., ::2], torch.rand(1, 1, 1, 4, 3, device=dev),
                            "zeros", True, False)


# ------------------------------------------------------------------------------------------ fused BatchNorm + ReLU
@pytest.mark.parametrize("n", [3001, 381, 1567, 2])     # > 2048 rows: two-level reduction; <= 2048: the one-launch kernels
@pytest.mark.parametrize("c,with_res,relu", [(32, False, True), (96, True, True), (256, False, False), (64, True, True)])
def test_bn_act_matches_torch(cuda_lib, c, with_res, relu, n):
    """bn_act == relu(BatchNorm1d(x) + residual) in training mode: outputs, all gradients and the running buffers
    (fp32 tolerance; statistics are accumulated in a different order than torch's Welford pass)."""
    from torch import nn
    from ponderv2_b200.bn_act import bn_act
    dev = _dev()
    torch.manual_seed(c)
    x0 = (torch.randn(n, c, device=dev) * 2.0 + 0.7)
    r0 = torch.randn(n, c, device=dev) if with_res else None
    g = torch.randn(n, c, device=dev)
    w0 = torch.rand(c, device=dev) + 0.5
    b0 = torch.rand(c, device=dev) - 0.5
    outs = []
    for fused in (False, True):
        bn = nn.BatchNorm1d(c, eps=1e-3, momentum=0.01).to(dev).train()
        with torch.no_grad():
            bn.weight.copy_(w0); bn.bias.copy_(b0)
        x = x0.clone().requires_grad_(True)
        r = r0.clone().requires_grad_(True) if with_res else None
        if fused:
            y = bn_act(x, bn, r, relu)
        else:
            y = bn(x)
            if r is not None:
                y = y + r
            if relu:
                y = torch.relu(y)
        y.backward(g)
        outs.append((y.detach(), x.grad, r.grad if r is not None else None, bn.weight.grad, bn.bias.grad,
                     bn.running_mean.clone(), bn.running_var.clone(), int(bn.num_batches_tracked)))
    ref, got = outs
    for i, name in enumerate(["y", "dx", "dres", "dgamma", "dbeta", "running_mean", "running_var"]):
        if ref[i] is None:
            assert got[i] is None
            continue
        err = (ref[i] - got[i]).abs().max().item()
        assert err < 2e-5 * max(1.0, ref[i].abs().max().item()), (name, err)
    assert ref[7] == got[7] == 1


# ------------------------------------------------------------------------------------------ BASELINE-size properties
def test_full_size_rulebook_and_conv_properties(cuda_lib):
    """BASELINE configs[1] size (100 k voxels): rulebooks bit-exact against the oracle, the submanifold symmetry
    nbr[k][j] = i <=> nbr[K-1-k][i] = j, exactly one pair per fine voxel in the strided rulebook, and size-independent
    properties of the convolution kernels: linearity and independence of the tile order."""
    import ponderv2_b200.spconv.pytorch as spconv
    dev = _dev()
    n = 100_000
    ind, shape = _indoor_indices(n, 2000)
    ind_t = torch.from_numpy(ind).to(dev)
    rb = spconv.build_subm_rulebook(ind_t, shape, 3)
    nbr = rb.nbr.cpu().numpy()
    assert np.array_equal(nbr, so.subm_rulebook(ind, shape, 3))                       # bit-exact at full size
    assert rb.num_pairs == int((nbr >= 0).sum())
    assert np.array_equal(nbr[13], np.arange(n, dtype=np.int32))                      # centre offset = identity
    for k in (0, 5, 12):                                                              # symmetry
        j = np.nonzero(nbr[k] >= 0)[0]
        assert np.array_equal(nbr[26 - k][nbr[k][j]], j.astype(np.int32))
    d = spconv.build_down_rulebook(ind_t, shape)
    oc, in2out, koff, _ = so.down_rulebook(ind, shape)
    assert d.out_indices.shape[0] == oc.shape[0]
    up = d.nbr_up.cpu().numpy()
    assert ((up >= 0).sum(0) == 1).all()                                              # one pair per fine voxel
    got_c, got_map, _ = so.canonical_down(d.out_indices.cpu().numpy(), d.in2out.cpu().numpy(), d.out_shape)
    ref_c, ref_map, _ = so.canonical_down(oc, in2out, d.out_shape)
    assert np.array_equal(got_c, ref_c) and np.array_equal(got_map, ref_map)

    torch.manual_seed(3)
    cin, cout = 32, 64
    x1, x2 = torch.randn(n, cin, device=dev), torch.randn(n, cin, device=dev)
    w3 = torch.randn(cout, 27, cin, device=dev) * 0.05
    f = lambda x, tm: spconv._gather_gemm(x, w3, None, tm, n)
    y1, y2, y12 = f(x1, rb.tmap), f(x2, rb.tmap), f(2.0 * x1 - 3.0 * x2, rb.tmap)
    scale = y12.abs().max().item()
    assert (y12 - (2.0 * y1 - 3.0 * y2)).abs().max().item() < 2e-5 * scale            # linearity
    y_nat = f(x1, spconv.TileMap(rb.nbr))                                             # natural row order, no skipping
    assert (y_nat - y1).abs().max().item() < 2e-5 * y1.abs().max().item()             # tile order changes nothing
    dy = torch.randn(n, cout, device=dev)
    dw_a = spconv._wgrad(x1, dy, rb.tmap, 27)
    dw_b = spconv._wgrad(x1, dy, spconv.TileMap(rb.nbr), 27)
    # fp32 accumulation over 1e5 rows in two different orders (plus atomics across row chunks): 1e-4 of the magnitude
    assert (dw_a - dw_b).abs().max().item() < 1e-4 * dw_b.abs().max().item()
    # adjointness of fwd and dgrad: <conv(x), dy> == <x, dgrad(dy)>
    wt = w3.flip(1).permute(2, 1, 0).contiguous()
    dx = spconv._gather_gemm(dy, wt, None, rb.tmap, n)
    lhs, rhs = (y1.double() * dy.double()).sum().item(), (x1.double() * dx.double()).sum().item()
    assert abs(lhs - rhs) < 1e-5 * (y1.double().abs() * dy.double().abs()).sum().item()


@pytest.mark.parametrize("cin,cout", [(32, 32), (96, 96), (256, 256), (128, 96)])
def test_full_size_conv_matches_oracle(cuda_lib, cin, cout):
    """BASELINE configs[1] size: SubMConv3d on the 100 k-voxel scene against the fp64 oracle — forward, data gradient
    and weight gradient — so that the kernels the benchmark actually launches (782 row tiles: the persistent
    gather-GEMM for the wide layers, two resident one-tile CTAs for the narrow ones, the row-chunked weight gradient)
    meet the oracle on a real rulebook, not only each other.  The forward is also checked against the independent
    plain-C restatement (oracle/spconv_c.c)."""
    import ponderv2_b200.spconv.pytorch as spconv
    from oracle import c_oracle
    from tests.conftest import record
    dev = _dev()
    n = 100_000
    ind, shape = _indoor_indices(n, 2000)
    torch.manual_seed(cin * 7 + cout)
    mod = spconv.SubMConv3d(cin, cout, 3, padding=1, bias=True, indice_key="k").to(dev)
    with torch.no_grad():
        mod.weight.normal_(0.0, 0.05)
        mod.bias.normal_(0.0, 0.1)
    x64 = torch.randn(n, cin, dtype=torch.float64)
    g64 = torch.randn(n, cout, dtype=torch.float64)
    xg = x64.to(dev, torch.float32).requires_grad_(True)
    out = mod(spconv.SparseConvTensor(xg, torch.from_numpy(ind).to(dev), shape, 1)).features
    out.backward(g64.to(dev, torch.float32))
    w64 = mod.weight.detach().cpu().double().requires_grad_(True)
    b64 = mod.bias.detach().cpu().double().requires_grad_(True)
    xo = x64.clone().requires_grad_(True)
    yo = so.subm_conv(so.OracleSparseTensor(xo, ind, shape, 1), w64, b64, 3, "k").features
    yo.backward(g64)
    yc = c_oracle.sparse_conv(x64.numpy(), w64.detach().reshape(cout, 27, cin).numpy(), b64.detach().numpy(),
                              so.subm_rulebook(ind, shape, 3))
    assert np.abs(yc - yo.detach().numpy()).max() < 1e-10 * max(1.0, float(yo.abs().max()))
    scale = lambda t: max(t.abs().max().item(), 1e-6)
    errs = {
        "y": (out.detach().cpu().double() - yo.detach()).abs().max().item() / scale(yo),
        "dx": (xg.grad.cpu().double() - xo.grad).abs().max().item() / scale(xo.grad),
        "dw": (mod.weight.grad.cpu().double() - w64.grad).abs().max().item() / scale(w64.grad),
        "db": (mod.bias.grad.cpu().double() - b64.grad).abs().max().item() / scale(b64.grad),
    }
    record("full_size_conv_matches_oracle", cin=cin, cout=cout, **errs)
    # fp32 storage, fp32-grade tensor-core products, fp32 accumulation over 27 * cin terms (y, dx) / ~1e5 rows (dw, db)
    for k, v in errs.items():
        assert v < 1e-4, (k, v, errs)


def test_full_size_stem_conv_matches_oracle(cuda_lib):
    """The 5^3 stem (spconv_unet_v1m1_base.py:111-119: 6 -> 32 channels, 125 offsets) at the 100 k-voxel size: the
    persistent kernel with two metadata buffers (125 x 128 neighbour indices per tile) and the zero-padded 6 -> 8 input
    channels, forward and weight gradient against the fp64 oracle."""
    import ponderv2_b200.spconv.pytorch as spconv
    from tests.conftest import record
    dev = _dev()
    n = 100_000
    ind, shape = _indoor_indices(n, 2000)
    torch.manual_seed(11)
    mod = spconv.SubMConv3d(6, 32, 5, padding=1, bias=False, indice_key="stem").to(dev)
    with torch.no_grad():
        mod.weight.normal_(0.0, 0.05)
    x64 = torch.randn(n, 6, dtype=torch.float64)
    g64 = torch.randn(n, 32, dtype=torch.float64)
    out = mod(spconv.SparseConvTensor(x64.to(dev, torch.float32), torch.from_numpy(ind).to(dev), shape, 1)).features
    out.backward(g64.to(dev, torch.float32))
    w64 = mod.weight.detach().cpu().double().requires_grad_(True)
    yo = so.subm_conv(so.OracleSparseTensor(x64, ind, shape, 1), w64, None, 5, "stem").features
    yo.backward(g64)
    scale = lambda t: max(t.abs().max().item(), 1e-6)
    errs = {"y": (out.detach().cpu().double() - yo.detach()).abs().max().item() / scale(yo),
            "dw": (mod.weight.grad.cpu().double() - w64.grad).abs().max().item() / scale(w64.grad)}
    record("full_size_stem_conv_matches_oracle", **errs)
    for k, v in errs.items():
        assert v < 1e-4, (k, v, errs)


@pytest.mark.parametrize("cin,cout,tma", [(64, 64, 1), (128, 64, 1), (64, 128, 0), (256, 256, 1), (96, 96, 0)])
def test_full_size_conv_bf16_matches_oracle(cuda_lib, cin, cout, tma):
    """bf16 storage at the BASELINE size (100 k voxels): forward, data gradient and weight gradient against the fp64
    oracle fed with the same bf16-rounded inputs.  tma = 1 forces the TMA gather4 kernel (UTMALDG.2D.GATHER4: missing
    neighbours are out-of-bounds row indices the copy engine zero-fills), tma = 0 the cp.async kernel; the weight gradient
    is the MN-major tcgen05 kernel in both.  bf16 outputs: <= 2^-8 relative per element on top of fp32 accumulation."""
    import ponderv2_b200.spconv.pytorch as spconv
    from tests.conftest import record
    dev = _dev()
    n = 100_000
    ind, shape = _indoor_indices(n, 2000)
    torch.manual_seed(cin + 3 * cout)
    cuda_lib.pv2_set_option(b"gg_tma", tma)
    try:
        mod = spconv.SubMConv3d(cin, cout, 3, padding=1, bias=False, indice_key="k").to(dev)
        with torch.no_grad():
            mod.weight.normal_(0.0, 0.05)
        xb = torch.randn(n, cin, device=dev).to(torch.bfloat16)
        gb = torch.randn(n, cout, device=dev).to(torch.bfloat16)
        xg = xb.clone().requires_grad_(True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            out = mod(spconv.SparseConvTensor(xg, torch.from_numpy(ind).to(dev), shape, 1)).features
        assert out.dtype == torch.bfloat16
        out.backward(gb)
    finally:
        cuda_lib.pv2_set_option(b"gg_tma", -1)
    w64 = mod.weight.detach().to(torch.bfloat16).cpu().double().requires_grad_(True)
    xo = xb.cpu().double().requires_grad_(True)
    yo = so.subm_conv(so.OracleSparseTensor(xo, ind, shape, 1), w64, None, 3, "k").features
    yo.backward(gb.cpu().double())
    scale = lambda t: max(t.abs().max().item(), 1e-6)
    errs = {
        "y": (out.detach().cpu().double() - yo.detach()).abs().max().item() / scale(yo),
        "dx": (xg.grad.cpu().double() - xo.grad).abs().max().item() / scale(xo.grad),
        "dw": (mod.weight.grad.cpu().double() - w64.grad).abs().max().item() / scale(w64.grad),
    }
    record("full_size_conv_bf16_matches_oracle", cin=cin, cout=cout, tma=tma, **errs)
    assert errs["y"] < 1e-2 and errs["dx"] < 1e-2, errs      # bf16 output rounding (2^-8) relative to the max magnitude
    assert errs["dw"] < 1e-3, errs                            # fp32 result of exact bf16 products


@pytest.mark.parametrize("c,with_res,relu", [(32, False, True), (96, True, True), (256, True, False)])
def test_bn_act_bf16_matches_torch(cuda_lib, c, with_res, relu):
    """Fused BatchNorm + residual + ReLU on bf16 features (fp32 statistics / parameters) against fp64 torch on the same
    bf16-rounded inputs: outputs and input gradients to bf16 rounding, parameter gradients and statistics to 1e-3."""
    from ponderv2_b200.bn_act import bn_act
    dev = _dev()
    torch.manual_seed(c)
    n = 5003
    x = (torch.randn(n, c, device=dev) * 2 + 0.5).to(torch.bfloat16)
    res = torch.randn(n, c, device=dev).to(torch.bfloat16) if with_res else None
    g = torch.randn(n, c, device=dev).to(torch.bfloat16)
    bn = torch.nn.BatchNorm1d(c, eps=1e-3, momentum=0.01).to(dev).train()
    with torch.no_grad():
        bn.weight.uniform_(0.5, 1.5); bn.bias.normal_()
    ref = torch.nn.BatchNorm1d(c, eps=1e-3, momentum=0.01).double().train()
    ref.load_state_dict({k: v.detach().cpu().double() if v.is_floating_point() else v.cpu() for k, v in bn.state_dict().items()})
    xg = x.clone().requires_grad_(True)
    rg = res.clone().requires_grad_(True) if with_res else None
    y = bn_act(xg, bn, rg, relu)
    assert y.dtype == torch.bfloat16
    y.backward(g)
    xr = x.cpu().double().requires_grad_(True)
    rr = res.cpu().double().requires_grad_(True) if with_res else None
    yr = ref(xr)
    if with_res:
        yr = yr + rr
    if relu:
        yr = torch.relu(yr)
    # the kernel applies the ReLU mask from its own bf16 output; compare away from the kink
    yr.backward(g.cpu().double())
    sc = lambda t: max(t.abs().max().item(), 1e-6)
    assert (y.detach().cpu().double() - yr.detach()).abs().max().item() < 1e-2 * sc(yr)
    far = (yr.detach().abs() > 0.05) | (not relu)
    assert ((xg.grad.cpu().double() - xr.grad).abs() * far).max().item() < 2e-2 * sc(xr.grad)
    if with_res:
        assert ((rg.grad.cpu().double() - rr.grad).abs() * far).max().item() < 2e-2 * sc(rr.grad)
    assert (bn.weight.grad.cpu().double() - ref.weight.grad).abs().max().item() < 2e-2 * sc(ref.weight.grad)
    assert (bn.bias.grad.cpu().double() - ref.bias.grad).abs().max().item() < 2e-2 * sc(ref.bias.grad)
    assert (bn.running_mean.cpu().double() - ref.running_mean).abs().max().item() < 1e-4
    assert (bn.running_var.cpu().double() - ref.running_var).abs().max().item() < 1e-4 * sc(ref.running_var)


def test_spunet_backbone_bf16_matches_oracle(cuda_lib):
    """Whole SpUNet-v1m1 under bf16 autocast (BASELINE configs[2]'s dtype; the reference runs fp16 autocast) on an
    8 k-voxel batch against the fp64 oracle with the same fp32 master weights: features to bf16-chain accuracy, the
    parameter gradients jointly."""
    from ponderv2_b200.backbone import SpUNetBase
    from tests.conftest import record
    dev = _dev()
    torch.manual_seed(0)
    a, b = synth.indoor_cloud(4800, 21), synth.indoor_cloud(3200, 22)
    gc = np.concatenate([a["grid_coord"], b["grid_coord"]])
    feat = np.concatenate([a["feat"], b["feat"]])
    offset = np.array([4800, 8000], dtype=np.int64)
    model = SpUNetBase(in_channels=6, num_classes=0).to(dev).train()
    sd = {k: v.detach().cpu().double() for k, v in model.state_dict().items()}
    for k, v in sd.items():
        if v.is_floating_point() and ("weight" in k or "bias" in k):
            v.requires_grad_(True)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        out = model({"grid_coord": torch.from_numpy(gc).to(dev), "feat": torch.from_numpy(feat).to(dev),
                     "offset": torch.from_numpy(offset).to(dev)})
    assert out.dtype == torch.bfloat16
    ref = so.spunet_forward(sd, gc, torch.from_numpy(feat).double(), offset)
    err = (out.detach().cpu().double() - ref.detach()).abs().max().item() / ref.abs().max().item()
    rel_l2 = ((out.detach().cpu().double() - ref.detach()).norm() / ref.detach().norm()).item()
    g = torch.randn(ref.shape, dtype=torch.float64, generator=torch.Generator().manual_seed(1))
    ref.backward(g)
    out.backward(g.to(dev, torch.bfloat16))
    num = den = dot = n2 = 0.0
    for name, p in model.named_parameters():
        rg = sd[name].grad
        pg = p.grad.cpu().double()
        d = pg - rg
        num += d.pow(2).sum().item(); den += rg.pow(2).sum().item()
        dot += (pg * rg).sum().item(); n2 += pg.pow(2).sum().item()
    joint = (num / den) ** 0.5
    cos = dot / (den * n2) ** 0.5
    last = "dec.0.block1.conv2.weight"      # the last convolution: its gradient sees one BatchNorm + ReLU of error only
    rg = sd[last].grad
    last_err = ((dict(model.named_parameters())[last].grad.cpu().double() - rg).norm() / rg.norm()).item()
    record("spunet_backbone_bf16_matches_oracle", fwd_max=err, fwd_l2=rel_l2, grad_joint=joint, grad_cosine=cos,
           grad_last_conv=last_err)
    # Forward: 59 layers of bf16 storage, each rounding to 2^-9 relative, accumulate like a random walk (measured 3.3e-2).
    # Backward: a forward deviation eps flips the ReLU mask of the ~0.8 eps of the units that sit that close to zero, and a
    # flipped unit's gradient is 100 % wrong -> relative gradient error ~ sqrt(0.8 eps) PER ReLU layer (measured: 14 % at
    # the last convolution, 63 % jointly over the 59 layers, cosine 0.85).  Inherent to half-precision activations (the
    # reference's fp16 autocast has the same mechanism with a 8x smaller eps); tools/bf16_grad_diag.py prints it per layer.
    assert rel_l2 < 5e-2, rel_l2
    assert err < 0.15, err
    assert last_err < 0.25, last_err
    assert cos > 0.7, cos


# ------------------------------------------------------------------------------------------ dense linear (render MLP)
@pytest.mark.parametrize("rows,cin,cout,act", [(40_000, 64, 128, 1), (40_000, 192, 68, 0), (50_001, 128, 64, 0),
                                                (40_000, 68, 128, 3), (40_000, 64, 128, 2), (40_000, 128, 64, 4),
                                                (700, 192, 68, 0)])
def test_linear_matches_torch(cuda_lib, rows, cin, cout, act):
    """pv2_linear (identity-map gather-GEMM; >= 296 row tiles take the persistent kernel, fewer the one-tile kernel)
    against fp64 torch, including the fused SDF-decoder epilogues."""
    from ponderv2_b200.render.fused import _linear
    dev = _dev()
    torch.manual_seed(rows + cin)
    x = torch.randn(rows, cin, device=dev)
    w = torch.randn(cout, cin, device=dev) * 0.1
    bias = torch.randn(cout, device=dev) if act in (0, 1) else None
    v = x.double() @ w.double().t() + (bias.double() if bias is not None else 0.0)
    aux = torch.rand(rows, cout, device=dev)
    y = torch.randn(rows, cout, device=dev)
    y0 = y.clone()
    y2 = torch.empty(rows, cout, device=dev) if act == 1 else (aux if act in (2, 3) else None)
    _linear(x, cin, 0, False, w, bias, y, cout, 0, False, act, y2, cout, 0, rows, cin, cout)
    if act == 0:
        want = v
    elif act == 1:
        want = torch.nn.functional.softplus(v, beta=100)
        # sigmoid(100 h) has slope 25 at h = 0: the bf16x3 products leave |dh| <= 4e-5 at |h| ~ 3 (3xTF32: 8e-6)
        assert (y2.double() - torch.sigmoid(100 * v)).abs().max().item() < 1e-3
    elif act == 2:
        want = v * 100 * aux.double() * (1 - aux.double())
    elif act == 3:
        want = y0.double() + v * aux.double()
    else:
        want = y0.double() + v
    err = (y.double() - want).abs().max().item()
    assert err < 3e-5 * max(1.0, want.abs().max().item()), err
