"""Pin the CPU oracle (oracle/) to the reference: its own KAT for the sampler and golden vectors generated from the
reference's renderer code (oracle/gen_golden.py).  CPU only."""
import numpy as np
import pytest
import torch

from oracle import spconv_oracle as so
from oracle.render_oracle import NeusOracle
from oracle.trilinear_oracle import trilinear_sample
from tests.golden_util import load_render_case


# --- the reference's only known-answer test, libs/smooth-sampler/smooth_sampler/modules.py:104-156 -------------
@pytest.mark.parametrize("padding_mode", ["zeros", "border", "reflection"])
@pytest.mark.parametrize("align_corners", [True, False])
def test_trilinear_oracle_matches_grid_sample(padding_mode, align_corners):
    torch.manual_seed(3)
    inp = torch.rand(2, 2, 2, 3, 11, dtype=torch.float64, requires_grad=True)
    grid = (torch.rand(2, 2, 1, 5, 3, dtype=torch.float64) * 2.0 - 1.0).requires_grad_(True)
    o1 = trilinear_sample(inp, grid, padding_mode, align_corners, False)
    o2 = torch.nn.functional.grid_sample(inp, grid, padding_mode=padding_mode, align_corners=align_corners)
    assert torch.allclose(o1, o2)
    g1 = torch.autograd.grad(o1, [inp, grid], torch.ones_like(o1), create_graph=True)
    g2 = torch.autograd.grad(o2, [inp, grid], torch.ones_like(o2))
    assert torch.allclose(g1[0], g2[0]) and torch.allclose(g1[1], g2[1])


@pytest.mark.parametrize("padding_mode", ["zeros", "border"])
@pytest.mark.parametrize("smooth", [False, True])
def test_trilinear_oracle_gradgradcheck(padding_mode, smooth):
    torch.manual_seed(3)
    inp = torch.rand(1, 2, 2, 3, 5, dtype=torch.float64, requires_grad=True)
    grid = (torch.rand(1, 1, 1, 4, 3, dtype=torch.float64) * 2.0 - 1.0).requires_grad_(True)
    fn = lambda a, b: trilinear_sample(a, b, padding_mode, True, smooth)
    assert torch.autograd.gradcheck(fn, [inp, grid], eps=1e-4, atol=1e-3, rtol=1e-2)
    assert torch.autograd.gradgradcheck(fn, [inp, grid], eps=1e-4, atol=1e-3, rtol=1e-2)


# --- renderer oracle vs the reference's own NeuSModel ---------------------------------------------------------
@pytest.mark.parametrize("case", ["indoor_train", "indoor_eval", "outdoor_train", "indoor_train_c2", "indoor_eval_c2",
                                  "indoor_train_s128", "outdoor_train_c4"])
def test_render_oracle_matches_reference(case):
    meta, arr, sd, cfg = load_render_case(case)
    for p in sd.values():
        if p.is_floating_point():
            p.requires_grad_(True)
    vol = arr["volume"].clone().requires_grad_(True)
    orc = NeusOracle(sd, cfg)
    noise = {"uniform": arr["noise_uniform"], "pdf": arr["noise_pdf"]}
    out = orc.render(arr["rays_o"], arr["rays_d"], [vol], noise, training=meta["training"])
    for k in ["rgb", "depth", "normal", "weights", "sdf", "gradients", "z_vals", "sampled_points",
              "init_sampled_points", "init_weights", "new_sampled_points"]:
        if "out." + k not in arr:
            continue
        ref = arr["out." + k]
        err = (out[k].detach() - ref).abs().max().item()
        assert err < 2e-6 * max(1.0, ref.abs().max().item()), (k, err)
    ld = orc.loss(out, arr["depth_gt"], arr["rgb_gt"])
    for k, v in ld.items():
        assert abs(v.item() - arr["loss." + k].item()) < 1e-5 * max(1.0, abs(arr["loss." + k].item())), k
    total = orc.total_loss(ld)
    assert abs(total.item() - arr["total_loss"].item()) < 1e-5
    total.backward()
    gv = arr["grad_volume"]
    assert (vol.grad - gv).abs().max().item() < 1e-5 * max(1.0, gv.abs().max().item())
    for k, p in sd.items():
        key = "grad.field." + k[len("field."):] if k.startswith("field.") else "grad." + k
        if key in arr and p.grad is not None:
            ref = arr[key]
            assert (p.grad - ref).abs().max().item() < 2e-5 * max(1.0, ref.abs().max().item()), k


# --- spconv restatement: internal consistency (PARITY UNPINNED, see oracle/__init__.py) ------------------------
def _cloud(n, extent, batch, seed):
    rng = np.random.default_rng(seed)
    pts = set()
    while len(pts) < n:
        b = int(rng.integers(0, batch))
        c = rng.integers(0, extent, size=3)
        pts.add((b, int(c[0]), int(c[1]), int(c[2])))
    arr = np.array(sorted(pts), dtype=np.int32)
    rng.shuffle(arr)
    return arr


def test_subm_rulebook_brute_force():
    coords = _cloud(300, 9, 2, 0)
    shape = [12, 12, 12]
    nbr = so.subm_rulebook(coords, shape, 3)
    lut = {tuple(c): i for i, c in enumerate(coords.tolist())}
    for j, c in enumerate(coords.tolist()):
        for k0 in range(3):
            for k1 in range(3):
                for k2 in range(3):
                    q = (c[0], c[1] + k0 - 1, c[2] + k1 - 1, c[3] + k2 - 1)
                    assert nbr[(k0 * 3 + k1) * 3 + k2, j] == lut.get(q, -1)
    # submanifold symmetry used by the data-gradient kernel
    K = 27
    for k in range(K):
        rows = np.nonzero(nbr[k] >= 0)[0]
        assert np.all(nbr[K - 1 - k][nbr[k][rows]] == rows)


def test_down_rulebook_properties():
    coords = _cloud(500, 14, 2, 1)
    out, in2out, koff, oshape = so.down_rulebook(coords, [16, 16, 16])
    assert oshape == [8, 8, 8]
    assert np.all(out[in2out, 1:] == coords[:, 1:] >> 1) and np.all(out[in2out, 0] == coords[:, 0])
    keys = so.linear_key(out, oshape)
    assert np.all(np.diff(keys) > 0)
    nd, nu = so.down_maps(in2out, koff, out.shape[0])
    assert (nd >= 0).sum() == coords.shape[0]
    assert (nu >= 0).sum() == coords.shape[0]
    # odd extent: the last plane has no complete 2x2x2 window and is dropped
    c2 = np.array([[0, 4, 0, 0], [0, 3, 1, 1]], dtype=np.int32)
    out2, i2o2, _, osh2 = so.down_rulebook(c2, [5, 4, 4])
    assert osh2 == [2, 2, 2] and i2o2[0] == -1 and i2o2[1] == 0 and out2.shape[0] == 1


def test_sparse_conv_matches_dense_conv3d():
    """Submanifold conv == dense conv evaluated at the active sites (zero elsewhere)."""
    torch.manual_seed(0)
    coords = _cloud(200, 7, 1, 2)
    shape = [7, 7, 7]
    cin, cout = 5, 4
    x = torch.randn(200, cin, dtype=torch.float64)
    w = torch.randn(cout, 3, 3, 3, cin, dtype=torch.float64)
    nbr = so.subm_rulebook(coords, shape, 3)
    y = so.sparse_conv(x, w, None, nbr, 200)
    dense = torch.zeros(1, cin, 7, 7, 7, dtype=torch.float64)
    idx = torch.from_numpy(coords.astype(np.int64))
    dense[0, :, idx[:, 1], idx[:, 2], idx[:, 3]] = x.t()
    yd = torch.nn.functional.conv3d(dense, w.permute(0, 4, 1, 2, 3), padding=1)
    assert torch.allclose(y, yd[0, :, idx[:, 1], idx[:, 2], idx[:, 3]].t(), atol=1e-10)
    # strided conv == dense k2 s2 conv at the active outputs
    wd = torch.randn(cout, 2, 2, 2, cin, dtype=torch.float64)
    c8 = _cloud(150, 8, 1, 3)
    x8 = torch.randn(150, cin, dtype=torch.float64)
    out, in2out, koff, oshape = so.down_rulebook(c8, [8, 8, 8])
    nd, nu = so.down_maps(in2out, koff, out.shape[0])
    yd8 = so.sparse_conv(x8, wd, None, nd, out.shape[0])
    dense8 = torch.zeros(1, cin, 8, 8, 8, dtype=torch.float64)
    i8 = torch.from_numpy(c8.astype(np.int64))
    dense8[0, :, i8[:, 1], i8[:, 2], i8[:, 3]] = x8.t()
    ref8 = torch.nn.functional.conv3d(dense8, wd.permute(0, 4, 1, 2, 3), stride=2)
    oi = torch.from_numpy(out.astype(np.int64))
    assert torch.allclose(yd8, ref8[0, :, oi[:, 1], oi[:, 2], oi[:, 3]].t(), atol=1e-10)
    # inverse conv == transposed dense conv restricted to the fine active set
    wi = torch.randn(3, 2, 2, 2, cout, dtype=torch.float64)
    yi = so.sparse_conv(yd8, wi, None, nu, 150)
    dense_c = torch.zeros(1, cout, 4, 4, 4, dtype=torch.float64)
    dense_c[0, :, oi[:, 1], oi[:, 2], oi[:, 3]] = yd8.t()
    refi = torch.nn.functional.conv_transpose3d(dense_c, wi.permute(4, 0, 1, 2, 3), stride=2)
    assert torch.allclose(yi, refi[0, :, i8[:, 1], i8[:, 2], i8[:, 3]].t(), atol=1e-10)


def test_c_restatement_agrees_with_numpy_oracle():
    """oracle/spconv_c.c (plain C) and oracle/spconv_oracle.py (numpy/torch) were written independently from SURVEY
    Appendix B; they must agree bit-exactly on rulebooks and to 1e-12 on the fp64 convolution."""
    from ponderv2_b200 import build, synth
    build.build_oracle_c()
    from oracle import c_oracle
    rng = np.random.default_rng(5)
    cloud = synth.indoor_cloud(4000, 77)
    coords = np.concatenate([np.zeros((4000, 1), np.int64), cloud["grid_coord"]], 1).astype(np.int32)
    coords[2000:, 0] = 1                                   # two scenes in the batch
    shape = (cloud["grid_coord"].max(0) + 96).tolist()
    for ks in (3, 5):
        assert np.array_equal(c_oracle.subm_rulebook(coords, shape, ks), so.subm_rulebook(coords, shape, ks))
    oc, in2out, koff, oshape = c_oracle.down_rulebook(coords, shape)
    rc, rin2out, rkoff, roshape = so.down_rulebook(coords, shape)
    assert oshape == roshape and np.array_equal(oc, rc) and np.array_equal(in2out, rin2out) and np.array_equal(koff, rkoff)
    # a tight shape: windows outside the unpadded input are dropped by both
    tight = (cloud["grid_coord"].max(0) + 1).tolist()
    oc2, in2, _, _ = c_oracle.down_rulebook(coords, tight)
    rc2, rin2, _, _ = so.down_rulebook(coords, tight)
    assert np.array_equal(oc2, rc2) and np.array_equal(in2, rin2)
    nbr = so.subm_rulebook(coords, shape, 3)
    x = rng.standard_normal((4000, 24))
    w = rng.standard_normal((40, 3, 3, 3, 24)) * 0.1
    b = rng.standard_normal(40)
    y_c = c_oracle.sparse_conv(x, w.reshape(40, 27, 24), b, nbr)
    y_np = so.sparse_conv(torch.from_numpy(x), torch.from_numpy(w), torch.from_numpy(b), nbr, 4000).numpy()
    assert np.abs(y_c - y_np).max() < 1e-12 * max(1.0, np.abs(y_np).max())
