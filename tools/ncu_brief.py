"""One-paragraph summary of an `ncu --page details --csv` export: duration, DRAM / L2 / L1 / SM throughput, tensor pipe,
occupancy, registers, DRAM bytes.   python tools/ncu_brief.py <details.csv> [...]"""
import csv
import sys

KEEP = ["Duration", "DRAM Throughput", "L2 Cache Throughput", "L1/TEX Cache Throughput", "Compute (SM) Throughput",
        "Memory Throughput", "Executed Ipc Active", "Issue Slots Busy", "L2 Hit Rate", "L1/TEX Hit Rate",
        "Registers Per Thread", "Achieved Occupancy", "Theoretical Occupancy", "Grid Size", "Block Size",
        "Dynamic Shared Memory Per Block", "Waves Per SM", "No Eligible"]
for path in sys.argv[1:]:
    try:
        rows = list(csv.reader(open(path)))
    except OSError:
        print(path, "missing"); continue
    if len(rows) < 2:
        print(path, "empty"); continue
    h = rows[0]
    if "Metric Name" not in h:
        print(path, "no metric table:", " ".join(rows[0])[:120]); continue
    iname, ival, iunit, ik = h.index("Metric Name"), h.index("Metric Value"), h.index("Metric Unit"), h.index("Kernel Name")
    print(f"== {path}: {rows[1][ik][:90]}")
    seen = set()
    for r in rows[1:]:
        n = r[iname]
        if n in KEEP and n not in seen:
            seen.add(n)
            print(f"   {n:34s} {r[ival]:>12s} {r[iunit]}")
    # tensor pipe utilisation lives in the pipe-utilisation table of the compute section when present
    for r in rows[1:]:
        if "ensor" in r[iname]:
            print(f"   {r[iname]:34s} {r[ival]:>12s} {r[iunit]}")
