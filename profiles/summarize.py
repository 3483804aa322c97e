#!/usr/bin/env python
"""Condense `ncu --page raw --csv` exports (gpurun_out/<tag>_<kernel>_raw.csv) into the few counters DESIGN.md argues
from, one row per captured launch:   python profiles/summarize.py gpurun_out/r1d_*_raw.csv > profiles/r1d_ncu_full_summary.csv
and aggregate a launch list:          python profiles/summarize.py --launches gpurun_out/r1d_launches.csv"""
import csv, collections, re, sys

KEYS = [
    ("time_ms", "gpu__time_duration.sum"),
    ("grid", "launch__grid_size"), ("block", "launch__block_size"), ("regs", "launch__registers_per_thread"),
    ("smem_dyn_B", "launch__shared_mem_per_block_dynamic"),
    ("warps_active_per_sm", "sm__warps_active.avg.per_cycle_active"),
    ("fmaheavy_cycles_active_pct", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed"),
    ("fma_inst_pct", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
    ("alu_cycles_active_pct", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed"),
    ("issue_active_pct", "sm__issue_active.avg.pct_of_peak_sustained_elapsed"),
    ("lsu_inst_pct", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
    ("active_threads_per_inst", "smsp__thread_inst_executed_per_inst_executed.ratio"),
    ("dram_read_MB", "dram__bytes_read.sum"), ("dram_write_MB", "dram__bytes_write.sum"),
    ("dram_throughput_pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
    ("l2_hit_pct", "lts__t_sector_hit_rate.pct"),
    ("l1_hit_pct", "l1tex__t_sector_hit_rate.pct"),
    ("local_ld_inst", "smsp__sass_inst_executed_op_local_ld.sum"),
    ("stall_math_pipe", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"),
    ("stall_wait", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"),
    ("stall_long_sb", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"),
    ("stall_dispatch", "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio"),
    ("stall_no_inst", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio"),
    ("stall_barrier", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"),
]


def to_unit(v, unit, want):
    v = float(v.replace(",", ""))
    scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}
    if want == "time_ms":
        return v * scale.get(unit, 1.0)
    if want.endswith("_MB"):
        return v * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(unit, 1.0)
    return v


def full(paths):
    w = csv.writer(sys.stdout)
    w.writerow(["capture", "kernel"] + [k for k, _ in KEYS])
    for p in paths:
        rows = list(csv.reader(open(p)))
        if len(rows) < 3:
            continue
        hdr, units = rows[0], rows[1]
        col = {h: i for i, h in enumerate(hdr)}
        for r in rows[2:]:
            name = re.sub(r"\(.*", "", r[col["Kernel Name"]]).replace("void bz::", "")
            out = [p.split("/")[-1].replace("_raw.csv", ""), name]
            for k, m in KEYS:
                i = col.get(m)
                try:
                    out.append("%.4g" % to_unit(r[i], units[i], k) if i is not None and r[i] not in ("", "n/a", "no data") else "")
                except ValueError:
                    out.append(r[i])
            w.writerow(out)


def launches(path):
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    tot, cnt = collections.defaultdict(float), collections.Counter()
    for r in rows[1:]:
        name = re.sub(r"<.*|\(.*", "", r[ki]).replace("void bz::", "")
        tot[name] += to_unit(r[vi], r[ui], "time_ms")
        cnt[name] += 1
    T = sum(tot.values())
    w = csv.writer(sys.stdout)
    w.writerow(["kernel", "launches", "total_ms", "avg_us", "share_pct"])
    for k, v in sorted(tot.items(), key=lambda x: -x[1]):
        w.writerow([k, cnt[k], "%.3f" % v, "%.1f" % (1e3 * v / cnt[k]), "%.2f" % (100 * v / T)])


if __name__ == "__main__":
    if sys.argv[1] == "--launches":
        launches(sys.argv[2])
    else:
        full(sys.argv[1:])
