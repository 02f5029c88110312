"""Per-parameter gradient error of the bf16-autocast SpUNet against the fp64 oracle, in execution order (diagnostic)."""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle import spconv_oracle as so  # noqa: E402
from ponderv2_b200 import synth  # noqa: E402
from ponderv2_b200.backbone import SpUNetBase  # noqa: E402

dev = torch.device("cuda:0")
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
ref = so.spunet_forward(sd, gc, torch.from_numpy(feat).double(), offset)
g = torch.randn(ref.shape, dtype=torch.float64, generator=torch.Generator().manual_seed(1))
ref.backward(g)
inp = {"grid_coord": torch.from_numpy(gc).to(dev), "feat": torch.from_numpy(feat).to(dev), "offset": torch.from_numpy(offset).to(dev)}
for mode in ("fp32", "bf16"):
    model.zero_grad()
    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=mode == "bf16"):
        out = model(inp)
    out.backward(g.to(dev, out.dtype))
    print(f"== {mode}: fwd rel l2 {((out.detach().cpu().double() - ref.detach()).norm() / ref.detach().norm()).item():.3e}")
    num = den = 0.0
    rows = []
    for name, p in model.named_parameters():
        rg = sd[name].grad
        d = p.grad.cpu().double() - rg
        num += d.pow(2).sum().item(); den += rg.pow(2).sum().item()
        rows.append((name, d.norm().item() / max(rg.norm().item(), 1e-30), rg.norm().item()))
    print(f"   joint {np.sqrt(num / den):.3e}")
    if mode == "bf16":
        for name, e, nrm in rows:
            if name.endswith("weight") and ("conv" in name or ".0.weight" in name):
                print(f"   {name:40s} rel {e:9.3e}  |g| {nrm:9.3e}")
