#!/usr/bin/env python
"""Summarises an `ncu --set full` report (.ncu-rep) into the text table kept under profiles/ and a small JSON
that bench.py reads for `roofline.traffic` and the issue-slot figures.
  python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r1x_ncu_full_wavefront_c4terrain [--note "..."]
writes <out>.txt and <out>.json.  Needs `ncu` on PATH (no GPU)."""
import csv
import io
import json
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from yahr_b200.api import source_fingerprint  # noqa: E402

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__inst_issued.avg.per_cycle_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum",
    "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum", "sass__inst_executed_local_loads",
    "sass__inst_executed_local_stores",
    "l1tex__data_pipe_lsu_wavefronts.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__cycles_active.avg",
    "sm__cycles_elapsed.max", "smsp__cycles_active.avg",
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    note = sys.argv[sys.argv.index("--note") + 1] if "--note" in sys.argv else ""
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    names = [d[col["Kernel Name"]] for d in data]
    lines = ["# " + note] if note else []
    lines.append("%-88s %-16s %s" % ("Kernel Name", "", " | ".join(n[:40] for n in names)))
    for m in METRICS:
        if m in col:
            lines.append("%-88s %-16s %s" % (m, units[col[m]], " | ".join(d[col[m]] for d in data)))
    open(out + ".txt", "w").write("\n".join(lines) + "\n")

    def f(d, m):
        try:
            return float(d[col[m]].replace(",", ""))
        except Exception:
            return None

    def mbytes(d, m):
        v = f(d, m)
        u = units[col[m]].lower()
        scale = {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1.0)
        return None if v is None else v * scale

    # the capture describes the kernels of THIS source tree: bench.py refuses it when the sources have changed since
    js = {"source": out + ".txt", "fingerprint": source_fingerprint(), "kernels": []}
    for d in data:
        js["kernels"].append({
            "name": d[col["Kernel Name"]],
            "ms": f(d, "gpu__time_duration.sum") * {"ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3}.get(
                units[col["gpu__time_duration.sum"]], 1.0),
            "dram_bytes": (mbytes(d, "dram__bytes_read.sum") or 0) + (mbytes(d, "dram__bytes_write.sum") or 0),
            "issue_active": f(d, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
            "threads_per_inst": f(d, "smsp__thread_inst_executed_per_inst_executed.ratio"),
            "l1_hit": f(d, "l1tex__t_sector_hit_rate.pct"), "l2_hit": f(d, "lts__t_sector_hit_rate.pct"),
            "l1_lsu_wavefronts_pct": f(d, "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
            "warps_active_pct": f(d, "sm__warps_active.avg.pct_of_peak_sustained_active"),
            "inst_executed": f(d, "smsp__inst_executed.sum"),
            # `--set full` carries the per-SM average, a `--metrics` pass the sum
            "l1_lsu_wavefronts": (f(d, "l1tex__data_pipe_lsu_wavefronts.sum")
                                  if "l1tex__data_pipe_lsu_wavefronts.sum" in col else
                                  ((f(d, "SM_A.TriageCompute.l1tex__data_pipe_lsu_wavefronts.avg") or 0.0)
                                   * (f(d, "device__attribute_multiprocessor_count") or 148.0) or None)),
            "grid": f(d, "launch__grid_size"), "registers": f(d, "launch__registers_per_thread"),
            "local_sectors": (f(d, "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum") or 0)
            + (f(d, "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum") or 0),
            "global_ld_sectors": f(d, "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum"),
        })
    json.dump(js, open(out + ".json", "w"), indent=1)
    print("\n".join(lines))


if __name__ == "__main__":
    main()
