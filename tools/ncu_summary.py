#!/usr/bin/env python
"""Summarise one kernel launch of an .ncu-rep: duration, issue, pipes, stall reasons, DRAM bytes.

    python tools/ncu_summary.py gpurun_out/x.ncu-rep [> profiles/rNN_x.txt]
"""
import csv
import io
import subprocess
import sys

KEEP_EXACT = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.avg", "smsp__inst_executed.sum", "launch__grid_size",
    "launch__block_size", "launch__registers_per_thread", "launch__occupancy_limit_registers",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.avg.per_cycle_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fmalite.avg.pct_of_peak_sustained_active",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fmalite_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print("kernel:", d.get("Kernel Name"), " grid", d.get("Grid Size"), " block", d.get("Block Size"))
        for k in KEEP_EXACT:
            if k in d:
                print(f"  {k:78s} {d[k]:>18s} {units[hdr.index(k)]}")
        print("  -- warp stall reasons (warps per issue-active cycle) --")
        st = [(float(d[k]), k) for k in hdr if k.startswith("smsp__average_warps_issue_stalled_") and
              k.endswith("_per_issue_active.ratio") and d[k] not in ("", "n/a")]
        for v, k in sorted(st, reverse=True)[:8]:
            print(f"  {k[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]:30s} {v:8.3f}")
        print("  -- instruction mix (sass opcode: thread-level %, if collected) --")
        ops = [(float(d[k].replace(",", "")), k) for k in hdr if k.startswith("sass__inst_executed_per_opcode") and d[k] not in ("", "n/a")]
        tot = sum(v for v, _ in ops) or 1
        for v, k in sorted(ops, reverse=True)[:16]:
            print(f"  {k:60s} {100 * v / tot:6.2f}%")


if __name__ == "__main__":
    main()
