#!/usr/bin/env python
"""ncu_summary.py — condenses an `ncu --set full --import-source on` report into the small CSV kept under profiles/.

usage: python tools/ncu_summary.py REPORT.ncu-rep STEP_ATTEMPTS_PER_LAUNCH "free-text header" > profiles/NAME.csv

Reads the report with `ncu -i … --page raw --csv` (metrics) and `--page source --csv` (per-SASS-line instruction counts).
The per-warp-step instruction mix divides the executed warp instructions by the number of warp-level step attempts,
taken as STEP_ATTEMPTS / (average active threads per executed instruction); an instruction is counted on the FP64 pipe
when its opcode is a double-precision arithmetic / compare one (DADD, DMUL, DFMA, DSETP, DMNMX).
The issue model quoted in DESIGN.md is 2 x FP64 + other issue cycles per warp-step (tools/micro/dp_issue.cu).
"""
import csv
import io
import re
import subprocess
import sys

KEEP = [
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "launch__block_size", "launch__grid_size", "launch__occupancy_limit_registers", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
    "sass__inst_executed_shared_loads", "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__average_t_sectors_per_request_pipe_lsu_mem_global_op_st.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__sass_average_branch_targets_threads_uniform.pct", "smsp__thread_inst_executed_per_inst_executed.ratio",
]


def ncu(rep, page):
    return subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True, check=True).stdout


def is_fp64(op):
    return op.startswith("D") and not op.startswith("DEPBAR")   # DADD, DMUL, DFMA, DSETP, DMNMX (MUFU.RCP64H runs on the XU pipe)


def main():
    rep, attempts, header = sys.argv[1], float(sys.argv[2]), sys.argv[3]
    rows = list(csv.reader(io.StringIO(ncu(rep, "raw"))))
    names, units, vals = rows[0], rows[1], rows[2]
    m = {n: (u, v) for n, u, v in zip(names, units, vals)}
    src = list(csv.reader(io.StringIO(ncu(rep, "source"))))
    hdr = next(i for i, r in enumerate(src) if r and r[0] == "Address")
    col = {n: i for i, n in enumerate(src[hdr])}
    ops = {}
    total = 0.0
    for r in src[hdr + 1:]:
        if len(r) <= col["Instructions Executed"]:
            continue
        sass = r[col["Source"]].strip()
        sass = re.sub(r"^@!?U?P\d+\s+", "", sass)
        op = sass.split()[0] if sass else "?"
        n = float(r[col["Instructions Executed"]] or 0)
        ops[op] = ops.get(op, 0.0) + n
        total += n
    lanes = float(m["smsp__thread_inst_executed_per_inst_executed.ratio"][1])
    warp_steps = attempts / lanes
    fp64 = sum(n for op, n in ops.items() if is_fp64(op))
    by_base = {}
    for op, n in ops.items():
        by_base[op.split(".")[0]] = by_base.get(op.split(".")[0], 0.0) + n
    top = sorted(by_base.items(), key=lambda kv: -kv[1])[:14]
    cycles = float(m["sm__cycles_elapsed.avg"][1])
    grid = float(m["launch__grid_size"][1])
    smsps = 148 * 4
    print(f"# {header}")
    print(f"# instruction mix per warp-step (ncu source page; warp-steps = {attempts:.0f} step attempts / {lanes} active lanes): "
          f"FP64-pipe {fp64 / warp_steps:.1f}, other {(total - fp64) / warp_steps:.1f}; issue model 2*FP64+other = "
          f"{(2 * fp64 + total - fp64) / warp_steps:.0f} cycles; measured {cycles * smsps / warp_steps:.0f} cycles per warp-step per SMSP (grid {grid:.0f} CTAs)")
    print("# top opcodes per warp-step: " + ", ".join(f"{k} {v / warp_steps:.1f}" for k, v in top))
    print("metric,unit,value")
    print(f'"Kernel Name","","{m.get("Kernel Name", ("", "?"))[1]}"')
    for k in KEEP:
        if k in m:
            print(f'"{k}","{m[k][0]}","{m[k][1]}"')


if __name__ == "__main__":
    main()
