"""Per-kernel DRAM traffic from an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv`
launch list: launches, total time, bytes read + written, per launch.  Writes profiles/<tag>_traffic.json for bench.py."""
import collections, csv, json, sys

path, out = sys.argv[1], sys.argv[2]
lines = [l for l in open(path) if not l.startswith("==")]
agg = collections.defaultdict(lambda: dict(launches=0, us=0.0, dram_bytes=0.0))
seen = set()
# The render MLP's linear layers run on the same tensor-core kernels as the sparse convolutions (identity row map).  They
# are told apart by position in the launch list: forward linears sit between field_sample_fwd and ray_resample /
# field_post_fwd, backward linears between field_post_bwd and field_sample_bwd.  Their launches are booked under
# "<kernel> [render linear]" so that the sparse-conv traffic (bench.py's roofline family) is not mixed with them.
in_render, last_id = False, None
for r in csv.DictReader(lines):
    name = r["Kernel Name"].split("(")[0].replace("void ", "").replace("<unnamed>::", "")
    if r.get("ID") != last_id:
        last_id = r.get("ID")
        if "field_sample_fwd" in name or "field_post_bwd" in name:
            in_render = True
        elif "ray_resample" in name or "field_post_fwd" in name or "field_sample_bwd" in name or "ray_composite" in name:
            in_render = False
    if in_render and "umma_gather_gemm" in name:
        name += " [render linear]"
    try:
        v = float(r["Metric Value"].replace(",", ""))
    except ValueError:
        continue
    u = r["Metric Unit"]
    d = agg[name]
    if r["Metric Name"] == "gpu__time_duration.sum":
        d["us"] += v / 1e3 if u.startswith("n") else v * 1e3 if u.startswith("m") else v
        d["launches"] += 1
    elif r["Metric Name"].startswith("dram__bytes"):
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)
        d["dram_bytes"] += v * scale
res = {k: dict(launches=v["launches"], total_us=round(v["us"], 1), dram_bytes_total=v["dram_bytes"],
               dram_bytes_per_launch=v["dram_bytes"] / max(v["launches"], 1))
       for k, v in sorted(agg.items(), key=lambda kv: -kv[1]["us"])[:25]}
json.dump(dict(source=path, note="cold-cache, serialised per-launch replay under ncu (one warmed-up C2 step)", kernels=res),
          open(out, "w"), indent=1)
for k, v in list(res.items())[:8]:
    print(f"{v['total_us']:9.1f} us {v['launches']:4d}x  {v['dram_bytes_per_launch'] / 1e6:9.2f} MB/launch  {k}")
