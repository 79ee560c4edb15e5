#!/usr/bin/env python
"""Turn the raw ncu artefacts that `gpurun` brings back (gpurun_out/*.ncu-rep, gpurun_out/launches*.csv) into the small,
tracked summaries under profiles/.

    python tools/ncu_summary.py rep  gpurun_out/prof.ncu-rep  profiles/r1_xxx.md  ["title"]
    python tools/ncu_summary.py list gpurun_out/launches.csv  profiles/r1_xxx_launches.md ["title"]
"""
import csv
import io
import subprocess
import sys
from collections import OrderedDict, defaultdict

METRICS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs/thread"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem/block"),
    ("launch__occupancy_limit_shared_mem", "occupancy limit (smem), CTAs/SM"),
    ("launch__occupancy_limit_registers", "occupancy limit (regs), CTAs/SM"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram throughput % of peak"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram throughput % of peak (dram__)"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("lts__t_bytes.sum", "L2 bytes"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem wavefronts % of peak"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "active lanes / warp instruction"),
    ("sm__inst_executed_pipe_alu.sum", "ALU pipe warp instructions"),
    ("sm__inst_executed_pipe_fma.sum", "FMA pipe warp instructions"),
    ("sm__inst_executed_pipe_xu.sum", "XU pipe warp instructions (POPC, conversions)"),
    ("sm__inst_executed_pipe_lsu.sum", "LSU pipe warp instructions"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA pipe active %"),
    ("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "ALU pipe active %"),
    ("sm__inst_executed_pipe_tensor.sum", "tensor pipe warp instructions"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall: short scoreboard (smem) / issue"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall: long scoreboard (global) / issue"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall: barrier / issue"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall: math pipe throttle / issue"),
    ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall: MIO throttle / issue"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall: LG throttle / issue"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall: fixed-latency wait / issue"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall: not selected / issue"),
    ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "stall: branch resolving / issue"),
    ("smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio", "stall: dispatch / issue"),
]


def ncu_csv(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    return list(csv.reader(io.StringIO(out)))


def summarize_rep(rep, dst, title):
    rows = ncu_csv(rep)
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    lines = [f"# {title}", "", f"Source: `{rep}` (`ncu --set full --clock-control none --import-source on`, B200, per launch; times under ncu "
             "are serialised and cold-cache — compare shares and ratios, not absolutes).", ""]
    for r in data:
        name = r[col["Kernel Name"]]
        lines.append(f"## launch {r[col['ID']]}: `{name}`")
        lines.append("")
        lines.append("| metric | value | unit |")
        lines.append("|---|---:|---|")
        for key, label in METRICS:
            if key in col and r[col[key]] not in ("", "n/a"):
                lines.append(f"| {label} (`{key}`) | {r[col[key]]} | {units[col[key]]} |")
        lines.append("")
    with open(dst, "w") as f:
        f.write("\n".join(lines) + "\n")
    print(f"wrote {dst} ({len(data)} launches)")


def summarize_list(path, dst, title):
    rows = list(csv.reader(open(path)))
    start = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr = rows[start]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    gi, bi = hdr.index("Grid Size"), hdr.index("Block Size")
    agg = OrderedDict()
    geo = defaultdict(set)
    for r in rows[start + 1:]:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        v_us = v / 1e3 if r[ui] in ("ns", "nsecond") else (v if r[ui] in ("us", "usecond") else v * 1e3)
        name = r[ki]
        name = name[5:] if name.startswith("void ") else name
        cut = name.find("(")
        name = name[:cut] if cut > 0 and not name.startswith("(") else name
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v_us
        geo[name].add(f"{r[gi]}x{r[bi]}")
    total = sum(a[1] for a in agg.values())
    lines = [f"# {title}", "", f"Source: `{path}` (`ncu --metrics gpu__time_duration.sum --clock-control none`, B200; launches are "
             "serialised and cold-cache under ncu — the SHARE per kernel is what to compare with bench.py's CUDA-event stage times).", "",
             "| kernel | launches | total us | avg us | share | grid x block |", "|---|---:|---:|---:|---:|---|"]
    for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        g = sorted(geo[name])
        lines.append(f"| `{name[:110]}` | {n} | {t:.1f} | {t / n:.1f} | {100 * t / total:.1f}% | {g[0] if len(g) == 1 else str(len(g)) + ' shapes'} |")
    lines.append(f"| **total** | {sum(a[0] for a in agg.values())} | {total:.1f} | | 100% | |")
    with open(dst, "w") as f:
        f.write("\n".join(lines) + "\n")
    print(f"wrote {dst}")


if __name__ == "__main__":
    mode, src, dst = sys.argv[1:4]
    title = sys.argv[4] if len(sys.argv) > 4 else src
    (summarize_rep if mode == "rep" else summarize_list)(src, dst, title)
