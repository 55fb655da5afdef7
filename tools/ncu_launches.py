#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv --log-file x.csv` launch list per kernel into markdown for profiles/.

    python tools/ncu_launches.py gpurun_out/launches.csv profiles/x.md "title" [skip_first_n_launches]"""
import csv
import sys
from collections import OrderedDict

src, dst, title = sys.argv[1], sys.argv[2], sys.argv[3]
skip = int(sys.argv[4]) if len(sys.argv) > 4 else 0
rows = []
hdr = None
for r in csv.reader(open(src, errors="replace")):
    if hdr is None:
        if "Kernel Name" in r and "Metric Value" in r:
            hdr = r
        continue
    if len(r) == len(hdr):
        rows.append(dict(zip(hdr, r)))
rows = [r for r in rows if r["Metric Name"] == "gpu__time_duration.sum"]
total_launches = len(rows)
rows = rows[skip:]
agg = OrderedDict()
for r in rows:
    unit = r["Metric Unit"]
    v = float(r["Metric Value"].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
    k = r["Kernel Name"]
    d = agg.setdefault(k, dict(n=0, ms=0.0, block=set(), grid=set()))
    d["n"] += 1; d["ms"] += v; d["block"].add(r.get("Block Size", "")); d["grid"].add(r.get("Grid Size", ""))
tot = sum(d["ms"] for d in agg.values())
with open(dst, "w") as f:
    f.write(f"# {title}\n\n{total_launches} launches captured, the first {skip} skipped; per-launch times are cold-cache and serialised under ncu, so only the SHARE of each kernel is meaningful.\n\n")
    f.write("| kernel | launches | block | grids | total ms | share |\n|---|---|---|---|---|---|\n")
    for k, d in sorted(agg.items(), key=lambda kv: -kv[1]["ms"]):
        f.write(f"| `{k[:110]}` | {d['n']} | {','.join(sorted(d['block']))} | {','.join(sorted(d['grid']))[:60]} | {d['ms']:.3f} | {100 * d['ms'] / tot:.1f}% |\n")
print(open(dst).read())
