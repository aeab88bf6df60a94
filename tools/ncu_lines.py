#!/usr/bin/env python3
"""tools/ncu_lines.py <file.ncu-rep> [top] -- warp instructions and stall samples per CUDA source line
(needs -lineinfo and `ncu --import-source on`)."""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 60
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
cur_file = "?"
per_line = collections.OrderedDict()
hdr = None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        i_inst = hdr.index("Instructions Executed")
        i_samp = hdr.index("# Samples")
        i_thr = hdr.index("Thread Instructions Executed")
        continue
    if hdr is None or r[0] in ("Function Name", "Kernel Name"):
        continue
    if r[0] == "":      # SASS row under a source line
        continue
    try:
        n, s, t = int(r[i_inst]), int(r[i_samp]), int(r[i_thr])
    except (ValueError, IndexError):
        continue
    per_line[(cur_file, int(r[0]))] = (n, s, t, r[1].strip()[:90])
tot = sum(v[0] for v in per_line.values())
tots = sum(v[1] for v in per_line.values())
print("total warp inst %d, samples %d" % (tot, tots))
files = collections.Counter()
for (f, l), v in per_line.items():
    files[f] += v[0]
for f, n in files.most_common():
    print("  %-24s %6.2f%%" % (f, 100.0 * n / tot))
print("%-22s %7s %7s %5s  %s" % ("file:line", "inst%", "samp%", "thr", "source"))
for (f, l), v in sorted(per_line.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%-22s %6.2f%% %6.2f%% %5.1f  %s" % ("%s:%d" % (f, l), 100.0 * v[0] / tot, 100.0 * v[1] / max(tots, 1),
                                                v[2] / max(v[0], 1), v[3]))
