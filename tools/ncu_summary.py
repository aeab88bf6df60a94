#!/usr/bin/env python3
"""tools/ncu_summary.py <file.ncu-rep> [out.csv] -- key counters + opcode histogram of one capture"""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[-1]
keep = ("gpu__time_duration", "dram__bytes", "dram__throughput", "sm__throughput", "sm__warps_active", "launch__registers",
        "launch__occupancy", "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst",
        "smsp__issue_active", "sm__pipe_fp64", "sm__inst_executed_pipe_xu", "l1tex__t_sectors_pipe_lsu_mem_global_op_st",
        "l1tex__t_requests_pipe_lsu_mem_global_op_st", "lts__t_sectors_op_write", "lts__t_sectors_op_read", "smsp__warp_issue_stalled",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared", "sm__sass_inst_executed_op_local", "smsp__pcsamp_warps_issue_stalled",
        "sm__inst_executed_pipe", "launch__shared_mem", "sm__pipe_xu", "sm__pipe_alu", "sm__pipe_fma", "smsp__inst_executed_pipe")
lines = []
for h, u, v in zip(hdr, units, vals):
    if h.startswith(keep):
        lines.append("%s,%s,%s" % (h, u, v))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
srows = list(csv.reader(src.splitlines()))
h2 = srows[1]
ie, si, ss = h2.index("Instructions Executed"), h2.index("Source"), h2.index("# Samples")
ops, samp, tot = collections.Counter(), collections.Counter(), 0
for r in srows[2:]:
    if len(r) <= ie:
        continue
    parts = r[si].split()
    if not parts:
        continue
    op = (parts[1] if parts[0].startswith("@") and len(parts) > 1 else parts[0]).split(".")[0]
    n = int(r[ie])
    ops[op] += n
    samp[op] += int(r[ss])
    tot += n
lines.append("sass_instructions_static,,%d" % (len(srows) - 2))
lines.append("warp_instructions_executed,,%d" % tot)
for op, n in ops.most_common(28):
    lines.append("op:%s,warp_inst,%d,samples,%d" % (op, n, samp[op]))
out = "\n".join(lines)
print(out)
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(out + "\n")
