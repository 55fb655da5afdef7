#!/usr/bin/env python
"""Summarise one `ncu --set full --import-source on` capture (.ncu-rep) into a small markdown file for profiles/.

    python tools/ncu_summary.py gpurun_out/x.ncu-rep profiles/x.md "title / command line"
Needs `ncu` (reads the report here, no GPU needed)."""
import csv
import io
import os
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "kernel duration"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "registers / thread"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem / CTA"), ("launch__occupancy_limit_shared_mem", "occupancy limit (smem), CTAs/SM"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy (% of 64 warps)"),
    ("smsp__issue_active.avg.per_cycle_active", "issue slots busy per scheduler (IPC/SMSP)"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "FP64 pipe utilisation (% of peak, active cycles)"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FP32 FMA pipe utilisation"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe utilisation"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe utilisation"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "shared-memory wavefronts (% of peak)"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "shared-memory wavefronts"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "shared-memory bank-conflict wavefronts"),
    ("dram__bytes_read.sum", "DRAM bytes read"), ("dram__bytes_write.sum", "DRAM bytes written"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput (% of peak)"),
    ("lts__t_bytes.sum", "L2 bytes"), ("lts__t_sector_hit_rate.pct", "L2 hit rate"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput (% of peak)"),
    ("smsp__inst_executed.sum", "warp instructions executed"),
]
STALLS = ["barrier", "branch_resolving", "dispatch_stall", "lg_throttle", "long_scoreboard", "math_pipe_throttle", "mio_throttle", "no_instruction",
          "not_selected", "selected", "short_scoreboard", "wait", "membar", "sleeping"]


def ncu(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def main():
    rep, out, title = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else "")
    raw = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "raw", "--csv"]))))
    hdr, units = raw[0], raw[1]
    lines = [f"# {os.path.basename(rep)}", "", title, ""]
    for ki, vals in enumerate(raw[2:]):
        m = dict(zip(hdr, zip(units, vals)))
        lines += [f"## launch {ki}: `{m.get('Kernel Name', ('', '?'))[1]}`", "", "| metric | value | unit |", "|---|---|---|"]
        for k, label in KEYS:
            if k in m:
                lines.append(f"| {label} (`{k}`) | {m[k][1]} | {m[k][0]} |")
        lines += ["", "Warp stall reasons (warps stalled per issue-active cycle, `smsp__average_warps_issue_stalled_*_per_issue_active.ratio`):", "",
                  "| reason | ratio |", "|---|---|"]
        st = []
        for s in STALLS:
            k = f"smsp__average_warps_issue_stalled_{s}_per_issue_active.ratio"
            if k in m:
                st.append((float(m[k][1] or 0), s))
        for v, s in sorted(st, reverse=True):
            lines.append(f"| {s} | {v:.3f} |")
        lines.append("")
    # hottest source lines
    src = ncu(["-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"])
    rows = list(csv.reader(io.StringIO(src)))
    h, fname, agg = None, "", {}
    for r in rows:
        if len(r) >= 2 and r[0] == "File Path":
            fname = r[1].split("/")[-1]
            continue
        if len(r) > 10 and r[0] == "Line No":
            h = r
            continue
        if h is None or len(r) < len(h) or r[2] != "-":
            continue
        try:
            n = int(r[h.index("# Samples")])
        except ValueError:
            continue
        key = (fname, int(r[0]), r[1].strip())
        d = agg.setdefault(key, [0, 0])
        d[0] += n
        d[1] += int(float(r[h.index("Instructions Executed")] or 0))
    tot = sum(d[0] for d in agg.values())
    if tot:
        lines += ["## hottest source lines (PC samples)", "", "| samples | share | warp instructions | source |", "|---|---|---|---|"]
        for (f, ln, s), d in sorted(agg.items(), key=lambda kv: -kv[1][0])[:25]:
            lines.append(f"| {d[0]} | {100 * d[0] / tot:.1f}% | {d[1]} | `{f}:{ln}` `{s[:90].replace('|', '/')}` |")
    open(out, "w").write("\n".join(lines) + "\n")
    print("wrote", out)


if __name__ == "__main__":
    main()
