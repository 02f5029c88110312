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
