"""Development: per-ray kernels vs the torch formulation on the indoor_train golden case (run on the GPU box)."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from tests.golden_util import load_render_case, product_renderer_cfg
from ponderv2_b200 import _lib
from ponderv2_b200.render import RayBundle, build_renderer
from ponderv2_b200.render import ray

dev = torch.device("cuda:0")
meta, arr, sd, cfg = load_render_case("indoor_train")
res = {}
for mode in ("torch", "kernel"):
    model = build_renderer(product_renderer_cfg(meta)).to(dev)
    model.load_state_dict(sd, strict=True)
    model.train(True)
    model.use_fused = True
    model.use_ray_kernels = mode == "kernel"
    vol = arr["volume"].to(dev).requires_grad_(True)
    noise = {"uniform": arr["noise_uniform"].to(dev), "pdf": arr["noise_pdf"].to(dev)}
    out = model(RayBundle(arr["rays_o"].to(dev), arr["rays_d"].to(dev)), [vol], noise=noise)
    tg = {"depth": arr["depth_gt"].to(dev), "rgb": arr["rgb_gt"].to(dev)}
    ld = model.get_loss(out, tg)
    total = sum(v for k, v in ld.items() if "loss" in k)
    total.backward()
    res[mode] = (out, ld, vol.grad.clone(), {n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None})
    print(mode, {k: round(v.item(), 6) for k, v in ld.items()}, "ref", {k[5:]: round(arr[k].item(), 6) for k in arr if k.startswith("loss.")})
o0, o1 = res["torch"][0], res["kernel"][0]
for k in o0:
    print(f"out.{k}: max|torch-kernel| = {(o0[k] - o1[k]).abs().max().item():.3e}  shape {tuple(o1[k].shape)}")
print("grad_volume rel diff", ((res['torch'][2] - res['kernel'][2]).norm() / res['torch'][2].norm()).item())
for n in res["torch"][3]:
    a, b = res["torch"][3][n], res["kernel"][3].get(n)
    print(f"grad {n}: rel diff {((a - b).norm() / a.norm().clamp(min=1e-12)).item():.3e}" if b is not None else f"grad {n}: missing")
# loss kernel sums on the torch path's predictions
out = o0
sdf, z, grad = out["sdf"][..., 0].contiguous(), out["z_vals"][..., 0].contiguous(), out["gradients"].contiguous()
R, S = sdf.shape
lib = _lib.load()
sums = torch.empty(11, device=dev)
dp, rp = out["depth"].detach().reshape(-1).contiguous(), out["rgb"].detach().contiguous()
dg, rg = arr["depth_gt"].to(dev).reshape(-1).contiguous(), arr["rgb_gt"].to(dev).contiguous()
_lib.check(lib.pv2_ray_loss_fwd(_lib.ptr(dp), _lib.ptr(rp), _lib.ptr(dg), _lib.ptr(rg), _lib.ptr(sdf.detach()), _lib.ptr(z.detach()),
                                _lib.ptr(grad.detach()), R, S, 0.05, _lib.ptr(sums), _lib.stream_ptr()), "loss")
gt = dg[:, None]
valid = gt > 0
front = valid & (z < gt - 0.05); back = valid & (z > gt + 0.05); sm = valid & ~front & ~back
want = [(valid[:, 0] * (dg - dp).abs()).sum(), (rp - rg).abs().sum(), (torch.relu(0.05 - sdf) * front).sum(),
        ((z + sdf - gt).abs() * sm).sum(), ((grad.norm(dim=-1) - 1) ** 2).sum(), valid.sum(), 0, front.sum(), sm.sum(), 0,
        ((rp - rg) ** 2).sum()]
print("kernel sums", [round(v, 5) for v in sums.tolist()])
print("torch  sums", [round(float(v), 5) for v in want])
