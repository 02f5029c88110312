"""Benchmark of the PonderV2 pretraining hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c1]

One "step" = one full pretraining iteration on one synthetic scene per GPU: SpUNet backbone forward/backward (rulebooks
rebuilt every step), densify, dense projection, NeuS render of R rays x S samples with its second-order backward,
the losses, the single gradient all-reduce (N > 1) and the optimizer step.  Prints ONE JSON line on rank 0.

  value    rays/s with inputs resident in HBM (device-timed, CUDA events, max over ranks)
  e2e      same metric through the host-facing call: inputs in pinned host memory, H2D inside the timed region,
           loss read back (D2H) every step
  roofline dominant hand-written kernel family (sparse-conv gather-GEMM, forward + data gradient of all 59 layers):
           algorithmic bytes / CUDA-event time, summed over its launches inside the timed region, vs the measured copy
           bandwidth in MEASURED_PEAKS.json; `traffic` = DRAM bytes per launch from the newest committed ncu launch list
           (profiles/*_traffic.json, written by tools/traffic_from_launches.py)
  cpu_baseline  the CPU oracle (port of the reference path) timed on the host cores on a bounded sample

`--impl reference` times the reference's own algorithm on the CPU (oracle port: spconv is not installable offline and
smooth_sampler is CUDA-only) with every host thread, on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

import numpy as np
import torch

WORKLOADS = {
    # BASELINE.json configs[1]: ScanNet-shape, ~100 k active voxels, 4096 rays x 128 samples, fp32, 1 x B200
    "c2": dict(name="ScanNet-shape synthetic: 100k voxels SpUNet-v1m1, 4096 rays x 128 samples (96+32), fp32",
               voxels=100_000, rays=4096, s0=96, si=32, grid_shape=(128, 128, 32), cfg_id=2),
    # BASELINE.json configs[0]: plumbing-size case
    "c1": dict(name="synthetic 1 scene, 2k voxels, 128 rays x 32 samples (24+8), fp32",
               voxels=2_000, rays=128, s0=24, si=8, grid_shape=(32, 32, 16), cfg_id=1),
}


def renderer_cfg(s0: int, si: int) -> dict:
    """configs/scannet/pretrain-ponder-spunet-v1m1-0-base.py:31-93 (semantic branch off: no CLIP offline)."""
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


def make_scene(wl: dict, seed: int) -> dict:
    """Host (numpy) scene: voxel cloud + ray batch, in the layout the dataloader/collate hands to the model."""
    from ponderv2_b200 import synth
    c = synth.indoor_cloud(wl["voxels"], seed)
    r = synth.ray_batch(wl["rays"], seed + 7)
    gc = c["grid_coord"]
    # to_unit_cube (ponder_indoor_base.py:344-444) maps the scene into the renderer's +-0.5 cube; the densify step
    # only needs the voxel-frame coordinates and the longest bbox edge in voxels ("resolution", :622-627)
    return dict(grid_coord=gc, coord=c["coord"], feat=c["feat"], offset=c["offset"],
                resolution=np.array([int(gc.max())], dtype=np.int64),
                ray_o=r["rays_o"][None], ray_d=r["rays_d"][None], rgb=r["rgb"], depth=r["depth"])


# ------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.path = index, None, None

    def start(self):
        try:
            f = tempfile.NamedTemporaryFile("w", suffix=".csv", delete=False)
            self.path = f.name
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=f,
                                         stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in Path(self.path).read_text().splitlines():
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        os.unlink(self.path)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


T_START = time.time()


def log(msg: str) -> None:
    print(f"[bench +{time.time() - T_START:6.1f}s] {msg}", file=sys.stderr, flush=True)


def host_threads() -> int:
    """Threads for the CPU arm: every core up to 32 (torch's intra-op pool stops scaling on these small ops)."""
    return max(1, min(os.cpu_count() or 1, 32))


def profiled_traffic(kernel_prefix: str):
    """DRAM bytes per launch of `kernel_prefix` from the newest committed ncu launch list (profiles/*_traffic.json,
    written by tools/traffic_from_launches.py); None when no profile has been committed."""
    best = None
    for f in sorted((ROOT / "profiles").glob("*_traffic.json")):
        try:
            d = json.loads(f.read_text())
        except ValueError:
            continue
        tot_b = tot_n = 0
        for name, v in d.get("kernels", {}).items():
            if name.startswith(kernel_prefix):
                tot_b += v["dram_bytes_total"]; tot_n += v["launches"]
        if tot_n:
            best = dict(bytes_per_launch=tot_b / tot_n, source=f"profiles/{f.name}")
    return best


def measured_peaks() -> dict:
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return dict(hbm_gbs=float(d["hbm_gbs"]), bf16_tflops=float(d["bf16_tflops"]), source="measured")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, source="fallback")


# ------------------------------------------------------------------------------------------------------------
def cpu_oracle_step(wl: dict, sample_voxels: int, sample_rays: int, threads: int) -> dict:
    """Times the CPU oracle (port of the reference path) forward+backward on a bounded sample; returns seconds and
    the extrapolated whole-step time for the full workload (linear in voxels and in rays)."""
    from oracle import spconv_oracle as so
    from oracle.render_oracle import NeusOracle, RenderConfig
    from ponderv2_b200 import synth
    from ponderv2_b200.backbone import SpUNetBase
    from ponderv2_b200.render import build_renderer

    torch.set_num_threads(threads)
    torch.manual_seed(0)
    cloud = synth.indoor_cloud(sample_voxels, 4242)
    bb = SpUNetBase(in_channels=6, num_classes=0)
    sd = {k: v.detach().clone().requires_grad_(v.is_floating_point() and v.dim() > 0 and "running" not in k)
          for k, v in bb.state_dict().items()}
    t0 = time.perf_counter()
    feats = so.spunet_forward(sd, cloud["grid_coord"], torch.from_numpy(cloud["feat"]), cloud["offset"])
    feats.square().mean().backward()
    t_bb = time.perf_counter() - t0

    rcfg = renderer_cfg(wl["s0"], wl["si"])
    rm = build_renderer(rcfg)
    rsd = {k: v.detach().clone().requires_grad_(v.requires_grad) for k, v in rm.state_dict().items()}
    for k, p in rm.named_parameters():
        rsd[k].requires_grad_(p.requires_grad)
    cfg = RenderConfig(bbox=[-0.55] * 3 + [0.55] * 3, near_plane=0.01, num_samples=wl["s0"],
                       num_samples_importance=wl["si"], share_volume=False, norm_pts=True, norm_padding=0.1,
                       loss_weights=rcfg["loss"]["weights"])
    X, Y, Z = wl["grid_shape"]
    vol = torch.randn(128, Z, Y, X).requires_grad_(True)
    rays = synth.ray_batch(sample_rays, 99)
    noise = {"uniform": torch.rand(sample_rays, wl["s0"] + 1), "pdf": torch.rand(sample_rays, wl["si"] + 1)}
    orc = NeusOracle(rsd, cfg)
    t0 = time.perf_counter()
    out = orc.render(torch.from_numpy(rays["rays_o"]), torch.from_numpy(rays["rays_d"]), [vol], noise, True)
    ld = orc.loss(out, torch.from_numpy(rays["depth"]), torch.from_numpy(rays["rgb"]))
    orc.total_loss(ld).backward()
    t_r = time.perf_counter() - t0
    full = t_bb * wl["voxels"] / sample_voxels + t_r * wl["rays"] / sample_rays
    return dict(backbone_s=t_bb, render_s=t_r, full_step_s=full)


def run_reference(args, wl: dict) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = host_threads()
    sv, sr = min(wl["voxels"], 10_000), min(wl["rays"], 128)
    times = []
    for i in range(args.warmup + args.steps):
        t = cpu_oracle_step(wl, sv, sr, threads)
        if i >= args.warmup:
            times.append(t["full_step_s"])
    per_step = statistics.mean(times)
    value = wl["rays"] / per_step  # one host: the CPU arm does not scale with N
    sample = (f"SpUNet-v1m1 fwd+bwd on {sv} voxels + NeuS render fwd+bwd (2nd order) on {sr} rays x "
              f"{wl['s0'] + wl['si']} samples, extrapolated linearly to {wl['voxels']} voxels / {wl['rays']} rays")
    print(json.dumps({
        "impl": "reference", "metric": "pretrain_rays_per_sec", "value": value, "unit": "rays/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": per_step * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl["name"]},
        "cpu_baseline": {"value": value, "unit": "rays/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "scenes_per_sec": 1.0 / per_step,
    }))


# ------------------------------------------------------------------------------------------------------------
def build_model(wl: dict, dev):
    """Model (identical replicas: fixed seed), flat parameter/gradient buffers, SGD as in the reference configs
    (configs/scannet/pretrain-ponder-spunet-v1m1-0-base.py:96-98)."""
    from ponderv2_b200.dist import FlatParameters, broadcast_parameters
    from ponderv2_b200.pretrain import PonderIndoorStep
    torch.manual_seed(1234)
    model = PonderIndoorStep(backbone=dict(in_channels=6, num_classes=0), renderer=renderer_cfg(wl["s0"], wl["si"]),
                             projection=dict(in_channels=96, out_channels=128), grid_shape=wl["grid_shape"],
                             grid_size=0.02).to(dev).train()
    flat = FlatParameters(model)
    broadcast_parameters(flat)
    opt = torch.optim.SGD(flat.optimizer_params(), lr=5e-4, momentum=0.9, weight_decay=1e-4, nesterov=True)
    return model, flat, opt


def run_ours(args, wl: dict) -> None:
    import torch.distributed as dist
    from ponderv2_b200 import _lib
    from ponderv2_b200.dist import FlatParameters, broadcast_parameters
    from ponderv2_b200.pretrain import PonderIndoorStep

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py --impl ours needs a CUDA device: ponderv2_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()

    model, flat, opt = build_model(wl, dev)

    # per-rank scene (seed = 1000*config + scene index), kept in pinned host memory
    scene = make_scene(wl, 1000 * wl["cfg_id"] + rank)
    host = {k: torch.from_numpy(np.ascontiguousarray(v)).pin_memory() for k, v in scene.items()}
    h2d_bytes = sum(t.numel() * t.element_size() for t in host.values())
    resident = {k: v.to(dev) for k, v in host.items()}
    shape = (torch.from_numpy(scene["grid_coord"]).max(0).values + 96).tolist()
    loss_host = torch.zeros(1).pin_memory()

    def step(inputs: dict) -> torch.Tensor:
        data = dict(inputs)
        data["sparse_shape"] = shape
        flat.zero_grad()
        out = model(data)
        out["loss"].backward()
        flat.all_reduce_mean()
        opt.step()
        return out["loss"].detach()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed_loop(from_host: bool, profile: bool):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        l0 = lib.pv2_launch_count()
        if profile:
            _lib.PROFILE.start(["pv2_spconv_gather_gemm", "pv2_spconv_wgrad"])
        e0.record()
        for _ in range(args.steps):
            if from_host:
                inputs = {k: v.to(dev, non_blocking=True) for k, v in host.items()}
                loss = step(inputs)
                loss_host.copy_(loss.reshape(1), non_blocking=True)
            else:
                loss = step(resident)
        e1.record()
        barrier()
        _lib.PROFILE.stop()
        ms = e0.elapsed_time(e1)
        t = torch.tensor([ms], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item(), lib.pv2_launch_count() - l0, float(loss)

    log(f"model + scene ready on rank {rank}/{world}: {wl['voxels']} voxels, {wl['rays']} rays")
    for _ in range(max(args.warmup, 3)):
        step(resident)
    torch.cuda.synchronize()
    log("warm-up done")
    if args.profile_step:
        # for `ncu --profile-from-start off`: exactly one warmed-up step inside cudaProfilerStart/Stop, no timing
        torch.cuda.profiler.start()
        step(resident)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        return
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    ms_dev, launches, last_loss = timed_loop(from_host=False, profile=True)
    prof = _lib.PROFILE.summary()
    log(f"device-resident loop: {ms_dev / args.steps:.2f} ms/step")
    ms_e2e, _, _ = timed_loop(from_host=True, profile=False)
    log(f"host-fed loop: {ms_e2e / args.steps:.2f} ms/step")
    clk = clocks.stop() if rank == 0 else None

    rays_per_step = wl["rays"] * world
    value = rays_per_step * args.steps / (ms_dev * 1e-3)
    e2e = rays_per_step * args.steps / (ms_e2e * 1e-3)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = host_threads()
        sv, sr = min(wl["voxels"], 10_000), min(wl["rays"], 128)
        log(f"cpu baseline on {threads} threads ({os.cpu_count()} cores) ...")
        t = cpu_oracle_step(wl, sv, sr, threads)
        log(f"cpu baseline done: {t}")
        cpu = {"value": wl["rays"] / t["full_step_s"], "unit": "rays/s", "cores": threads, "kind": "port",
               "sample": f"oracle SpUNet fwd+bwd on {sv} voxels ({t['backbone_s']:.1f} s) + NeuS fwd+bwd on {sr} rays "
                         f"({t['render_s']:.1f} s), extrapolated linearly to the full step"}
    if rank == 0:
        peaks = measured_peaks()
        gg = prof.get("pv2_spconv_gather_gemm", dict(calls=0, ms=0.0, bytes=0))
        achieved = gg["bytes"] / max(gg["ms"], 1e-9) * 1e-6 if gg["calls"] else 0.0  # GB/s
        traffic = profiled_traffic("umma_gather_gemm")
        line = {
            "metric": "pretrain_rays_per_sec", "value": value, "unit": "rays/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_dev / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl["name"], "scenes_per_gpu_per_step": 1, "parallelism": f"dp{world}",
                       "projection": "SimpleConv3D-v1m1 96->128 (cuDNN; UNet3D-v1m2 is SURVEY 8f-1, out of scope)",
                       "l2": "per-step working set (201 MB + 268 MB dense volumes) exceeds the 126 MB L2"},
            "scenes_per_sec": world * args.steps / (ms_dev * 1e-3),
            "e2e": {"value": e2e, "unit": "rays/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 4,
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": int(launches),
            "roofline": {"kernel": "pv2_spconv_gather_gemm (fwd + dgrad, all layers)", "bound": "hbm",
                         "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                         "frac": achieved / peaks["hbm_gbs"],
                         "traffic": traffic["bytes_per_launch"] if traffic else None,
                         "traffic_source": traffic["source"] if traffic else None,
                         "algorithmic_bytes_per_launch": gg["bytes"] / max(gg["calls"], 1),
                         "peak_source": peaks["source"],
                         "launches": gg["calls"], "kernel_ms_per_step": gg["ms"] / args.steps,
                         "share_of_step": gg["ms"] / max(ms_dev, 1e-9)},
            "kernel_ms_per_step": {k: v["ms"] / args.steps for k, v in prof.items()},
            "cpu_baseline": cpu, "clocks": clk, "loss": last_loss,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile-step", action="store_true",
                    help="run one warmed-up step inside cudaProfilerStart/Stop and exit (for ncu --profile-from-start off)")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, wl)
    else:
        run_ours(args, wl)


if __name__ == "__main__":
    main()
