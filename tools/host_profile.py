"""cProfile of the host side of the C2 training step (where do the ~27 ms of enqueue time go?).
    python tools/host_profile.py [--steps 10] [--workload c2]"""
import argparse
import cProfile
import io
import pstats
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--workload", default="c2")
args = ap.parse_args()
wl = bench.WORKLOADS[args.workload]
dev = torch.device("cuda:0")
model, flat, opt = bench.build_model(wl, dev)
scene = bench.make_scene(wl, 1000 * wl["cfg_id"])
data0 = {k: torch.from_numpy(np.ascontiguousarray(v)).to(dev) for k, v in scene.items()}
shape = (torch.from_numpy(scene["grid_coord"]).max(0).values + 96).tolist()


def step():
    data = dict(data0)
    data["sparse_shape"] = shape
    opt.zero_grad()
    out = model(data)
    out["loss"].backward()
    flat.all_reduce_mean()
    opt.step()


for _ in range(3):
    step()
torch.cuda.synchronize()
pr = cProfile.Profile()
with torch.autograd.set_multithreading_enabled(False):   # backward in the calling thread, so the profiler sees it
    pr.enable()
    for _ in range(args.steps):
        step()
    pr.disable()
torch.cuda.synchronize()
for key in ("tottime", "cumulative"):
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats(key).print_stats(35)
    txt = s.getvalue()
    print(f"===== by {key} (totals over {args.steps} steps)")
    print("\n".join(line[:150] for line in txt.splitlines()[4:50]))
