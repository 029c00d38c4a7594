#!/usr/bin/env python
"""Summarise an `ncu --set full` report (one kernel) into a small text file for profiles/.

usage: python tools/summarize_ncu.py gpurun_out/prof.ncu-rep profiles/NAME.txt
"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
    "launch__waves_per_multiprocessor", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__warps_eligible.avg.per_cycle_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sectors.sum", "lts__t_sectors.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_hit.sum",
    "smsp__inst_executed_op_shared_atom.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_atom.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
]


def ncu(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def main():
    rep, out = sys.argv[1], sys.argv[2]
    rows = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "raw", "--csv"]))))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {h: (u, v) for h, u, v in zip(hdr, units, vals)}
    lines = [f"source report: {rep} (ncu --set full --clock-control none, 1 launch)", f"kernel: {d.get('Kernel Name', ('', '?'))[1]}", ""]
    for k in KEYS:
        if k in d:
            lines.append(f"{k:90s} {d[k][1]:>18s} {d[k][0]}")
    lines += ["", "warp stall reasons (warps per issue-active cycle):"]
    for h, (u, v) in d.items():
        if "smsp__average_warps_issue_stalled" in h and h.endswith("_per_issue_active.ratio") and float(v) > 0.1:
            lines.append(f"  {h.split('issue_stalled_')[1].split('_per_issue')[0]:30s} {float(v):8.3f}")
    src = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "source", "--csv"]))))
    hdr = src[1]
    ie, so, sm = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("# Samples")
    data = []
    for r in src[2:]:
        try:
            data.append((int(r[ie]), r[so].strip(), int(r[sm])))
        except Exception:
            pass
    tot, ts = sum(x[0] for x in data), max(1, sum(x[2] for x in data))
    thr = sorted(x[0] for x in data)[-min(len(data), 140)]
    lines += ["", f"hottest SASS (instructions executed >= {thr}; total {tot}; stall samples {ts}):"]
    prev = None
    for i, (c, s, n) in enumerate(data):
        if c >= thr:
            if prev is not None and i != prev + 1:
                lines.append("      ...")
            lines.append(f"  {c:12d} {100.0 * n / ts:6.2f}%  {s[:100]}")
            prev = i
    open(out, "w").write("\n".join(lines) + "\n")
    print("wrote", out, len(lines), "lines")


if __name__ == "__main__":
    main()
