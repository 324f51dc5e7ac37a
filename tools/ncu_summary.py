#!/usr/bin/env python
"""Condense ncu output into the small, tracked summaries kept under profiles/.

  python tools/ncu_summary.py rep  gpurun_out/prof_x.ncu-rep  > profiles/rNN_x.md     (one `--set full` capture)
  python tools/ncu_summary.py list gpurun_out/launches_x.csv > profiles/rNN_x_launches.md   (a gpu__time_duration launch list)

Runs on the CPU box (`ncu -i` needs no GPU).  Never run inside a timed region; numbers under ncu are not bench values."""
import csv
import io
import re
import subprocess
import sys
from collections import OrderedDict

KEYS = [
    ("gpu__time_duration.sum", "time"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs/thread"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem/block"),
    ("launch__occupancy_limit_registers", "occ limit regs (CTAs/SM)"),
    ("launch__occupancy_limit_shared_mem", "occ limit smem (CTAs/SM)"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram throughput % of peak"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram throughput % of peak"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("lts__t_bytes.sum", "L2 bytes"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit rate %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput % of peak"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64 pipe inst % of peak"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64 pipe cycles active %"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe cycles active %"),
    ("sm__pipe_tensor_op_dmma_cycles_active.avg.pct_of_peak_sustained_active", "DMMA pipe cycles active %"),
    ("sm__inst_executed_pipe_tensor_op_dmma.avg.pct_of_peak_sustained_active", "DMMA inst % of peak"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU inst % of peak"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
    ("smsp__average_warp_latency_issue_stalled_long_scoreboard_per_warp_active.pct", "stall long scoreboard"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard (warps/issue)"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_scoreboard"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math_pipe_throttle"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait"),
    ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall mio_throttle"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall lg_throttle"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall not_selected"),
]


def raw_rows(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


def short(name):
    return re.sub(r"\(.*", "", name).replace("void ", "").strip()


def cmd_rep(path, extra_regex=None):
    hdr, units, rows = raw_rows(path)
    idx = {h: i for i, h in enumerate(hdr)}
    print(f"source: `{path}` (ncu --set full --clock-control none; replayed, cold cache: use for counters and shares, not for bench values)\n")
    for r in rows:
        print(f"### {short(r[idx['Kernel Name']])}  (launch id {r[idx['ID']]})\n")
        print("| metric | value |")
        print("|---|---|")
        seen = set()
        for key, label in KEYS:
            if key in idx and label not in seen and r[idx[key]] != "":
                seen.add(label)
                print(f"| {label} (`{key}`) | {r[idx[key]]} {units[idx[key]]} |")
        if extra_regex:
            for h in hdr:
                if re.search(extra_regex, h) and h not in dict(KEYS):
                    print(f"| `{h}` | {r[idx[h]]} {units[idx[h]]} |")
        if "dram__bytes_read.sum" in idx:

            def tobytes(v, u):
                m = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
                return float(v.replace(",", "")) * m.get(u, 1)

            tr = tobytes(r[idx["dram__bytes_read.sum"]], units[idx["dram__bytes_read.sum"]]) + tobytes(r[idx["dram__bytes_write.sum"]], units[idx["dram__bytes_write.sum"]])
            print(f"| **traffic** = dram read + write | {tr:.6g} byte |")
        print()


def cmd_list(path):
    rows = list(csv.reader(l for l in open(path) if not l.startswith("==")))
    hdr = rows[0]
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = OrderedDict()
    for r in rows[1:]:
        if len(r) <= iv:
            continue
        v = float(r[iv].replace(",", ""))
        u = r[iu]
        us = v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1e-3)
        a = agg.setdefault(short(r[ik]), [0, 0.0])
        a[0] += 1
        a[1] += us
    tot = sum(a[1] for a in agg.values())
    print(f"source: `{path}` (ncu --metrics gpu__time_duration.sum --clock-control none; serialised cold-cache launches: compare SHARES)\n")
    print("| kernel | launches | total us | share | avg us |")
    print("|---|---|---|---|---|")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| {k} | {a[0]} | {a[1]:.1f} | {a[1] / tot:.3f} | {a[1] / a[0]:.2f} |")
    print(f"| **total** | {sum(a[0] for a in agg.values())} | {tot:.1f} | 1 | |")


if __name__ == "__main__":
    if len(sys.argv) < 3:
        raise SystemExit(__doc__)
    if sys.argv[1] == "rep":
        cmd_rep(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else None)
    else:
        cmd_list(sys.argv[2])
