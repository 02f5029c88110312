"""Top stall sites of an ncu source-page CSV (--page source --csv --print-source sass): per instruction the sampled
stall reasons, with a few neighbouring instructions for orientation.   python tools/ncu_stalls.py <file.csv> [top]"""
import csv
import sys

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
rows = list(csv.reader(open(path)))
hdr = rows[1]
ia, isrc, ismp = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
ins = [r for r in rows[2:] if len(r) == len(hdr)]
total = sum(int(r[ismp]) for r in ins)
print(f"{len(ins)} instructions, {total} samples")
by_reason = {}
for r in ins:
    for i in stall_cols:
        by_reason[hdr[i]] = by_reason.get(hdr[i], 0) + int(r[i] or 0)
print("by reason:", ", ".join(f"{k[6:]} {v * 100 // max(total, 1)}%" for k, v in sorted(by_reason.items(), key=lambda kv: -kv[1]) if v))
order = sorted(range(len(ins)), key=lambda j: -int(ins[j][ismp]))[:top]
for j in order:
    r = ins[j]
    reasons = sorted(((int(r[i] or 0), hdr[i][6:]) for i in stall_cols), reverse=True)[:3]
    rs = " ".join(f"{n}:{c}" for c, n in reasons if c)
    print(f"--- #{j} {int(r[ismp]) * 100 / max(total, 1):5.1f}%  {r[isrc].strip()[:90]}   [{rs}]")
    for jj in range(max(0, j - 3), j):
        print(f"        {ins[jj][isrc].strip()[:100]}")
