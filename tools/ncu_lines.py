#!/usr/bin/env python
"""Aggregate an `ncu --page source --print-source cuda,sass --csv` dump per CUDA source line (stall samples by reason)."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = None
agg = {}
fname = ""
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        fname = r[1].split("/")[-1]
        continue
    if len(r) > 10 and r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) < len(hdr) or r[2] != "-":      # source-line rows have '-' as address
        continue
    try:
        n = int(r[hdr.index("# Samples")])
    except ValueError:
        continue
    key = (fname, int(r[0]), r[1].strip())
    d = agg.setdefault(key, dict(n=0, inst=0, bar=0, ssb=0, wait=0, lsb=0, br=0, sel=0, mio=0, noinst=0, conf=0))
    g = lambda name: int(float(r[hdr.index(name)] or 0))
    d["n"] += n; d["inst"] += g("Instructions Executed"); d["bar"] += g("stall_barrier"); d["ssb"] += g("stall_short_sb"); d["wait"] += g("stall_wait")
    d["lsb"] += g("stall_long_sb"); d["br"] += g("stall_branch_resolving"); d["sel"] += g("stall_selected"); d["mio"] += g("stall_mio"); d["noinst"] += g("stall_no_inst")
    d["conf"] += g("L1 Wavefronts Shared Excessive")
tot = sum(d["n"] for d in agg.values())
print("total samples", tot)
for (f, ln, src), d in sorted(agg.items(), key=lambda kv: -kv[1]["n"])[:top]:
    print(f"{d['n']:7d} {100*d['n']/max(tot,1):5.1f}% bar {d['bar']:6d} ssb {d['ssb']:6d} wait {d['wait']:6d} lsb {d['lsb']:6d} br {d['br']:5d} sel {d['sel']:5d} | inst {d['inst']:9d} xs-smem {d['conf']:8d} | {f}:{ln}: {src[:100]}")
