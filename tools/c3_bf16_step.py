"""BASELINE configs[2] shape on one GPU (Structured3D-shape: 200 k voxels, 8192 rays x 128 samples) with the backbone under
bf16 autocast (the reference's AMP mode is fp16 autocast, engines/train.py:183-196; B200 runs it in bf16), renderer fp32:
checks the step runs, the loss is finite and close to the fp32 step's, and reports ms/step.  Not a bench line."""
import sys, time
from pathlib import Path
import numpy as np
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench
from ponderv2_b200.dist import FlatParameters
from ponderv2_b200.pretrain import PonderIndoorStep

dev = torch.device("cuda:0")
wl = dict(name="c3", voxels=200_000, rays=8192, s0=96, si=32, grid_shape=(128, 128, 32), cfg_id=3)
scene = bench.make_scene(wl, 3000)
data = {k: torch.from_numpy(np.ascontiguousarray(v)).to(dev) for k, v in scene.items()}
shape = (torch.from_numpy(scene["grid_coord"]).max(0).values + 96).tolist()
res = {}
for mode in ("fp32", "bf16"):
    torch.manual_seed(1234)
    model = PonderIndoorStep(backbone=dict(in_channels=6, num_classes=0), renderer=bench.renderer_cfg(96, 32),
                             projection=dict(in_channels=96, out_channels=128), grid_shape=wl["grid_shape"], grid_size=0.02).to(dev).train()
    flat = FlatParameters(model)
    opt = torch.optim.SGD(flat.optimizer_params(), lr=5e-4, momentum=0.9)
    torch.manual_seed(7)
    noise = {"uniform": torch.rand(8192, 97, device=dev), "pdf": torch.rand(8192, 33, device=dev)}
    def step():
        d = dict(data); d["sparse_shape"] = shape
        flat.zero_grad()
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=(mode == "bf16")):
            d["sparse_backbone_feat"] = None
            out = model(d, noise=noise)
        out["loss"].backward()
        opt.step()
        return out["loss"].detach()
    first = step().item()
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        l = step()
    e1.record(); torch.cuda.synchronize()
    res[mode] = (first, l.item(), e0.elapsed_time(e1) / 5)
    print(f"{mode}: first-step loss {first:.5f}, loss after 9 steps {l.item():.5f}, {res[mode][2]:.2f} ms/step "
          f"({8192 / res[mode][2] * 1e3:.0f} rays/s, 200k voxels)", flush=True)
assert np.isfinite(res["bf16"][0]) and abs(res["bf16"][0] - res["fp32"][0]) < 0.05 * abs(res["fp32"][0]), res
