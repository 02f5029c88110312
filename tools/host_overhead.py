"""How much of a pretraining step is the host?  Enqueue time (no synchronisation) against device time, per phase.

    python tools/host_overhead.py [--workload c2] [--steps 10]

For each phase of the step (backbone forward incl. rulebooks, densify + projection + renderer forward + loss, backward,
all-reduce + optimizer) prints the host wall time spent enqueueing it and the CUDA-event time between its first and
last kernel.  If the host column sums to more than the device column the step is launch-bound.
"""
from __future__ import annotations

import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--steps", type=int, default=10)
    args = ap.parse_args()
    from ponderv2_b200.dist import FlatParameters
    wl = bench.WORKLOADS[args.workload]
    dev = torch.device("cuda:0")
    model, flat, opt = bench.build_model(wl, dev)
    scene = bench.make_scene(wl, 1000 * wl["cfg_id"])
    data0 = {k: torch.from_numpy(np.ascontiguousarray(v)).to(dev) for k, v in scene.items()}
    shape = (torch.from_numpy(scene["grid_coord"]).max(0).values + 96).tolist()

    names = ["backbone_fwd", "render_fwd", "backward", "optimizer"]
    host = {n: 0.0 for n in names}
    devt = {n: 0.0 for n in names}
    total_host = total_dev = 0.0
    for it in range(args.steps + 3):
        data = dict(data0)
        data["sparse_shape"] = shape
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
        torch.cuda.synchronize()
        t = [time.perf_counter()]
        ev[0].record()
        flat.zero_grad()
        data["sparse_backbone_feat"] = model.backbone(data)
        t.append(time.perf_counter()); ev[1].record()
        out = model.forward_after_backbone(data)
        t.append(time.perf_counter()); ev[2].record()
        out["loss"].backward()
        t.append(time.perf_counter()); ev[3].record()
        flat.all_reduce_mean()
        opt.step()
        t.append(time.perf_counter()); ev[4].record()
        torch.cuda.synchronize()
        t_end = time.perf_counter()
        if it >= 3:
            for i, n in enumerate(names):
                host[n] += (t[i + 1] - t[i]) * 1e3
                devt[n] += ev[i].elapsed_time(ev[i + 1])
            total_host += (t[4] - t[0]) * 1e3
            total_dev += (t_end - t[0]) * 1e3
    k = args.steps
    res = {"workload": wl["name"], "steps": k,
           "host_enqueue_ms": {n: host[n] / k for n in names}, "device_ms": {n: devt[n] / k for n in names},
           "host_enqueue_total_ms": total_host / k, "wall_ms_incl_drain": total_dev / k}
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
