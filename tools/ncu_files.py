#!/usr/bin/env python3
"""tools/ncu_files.py <rep> <cells> [file] -- warp instructions per 32-cell chunk and stall samples per source file; with a file name: per line (top 40)"""
import subprocess, csv, sys, collections
rep = sys.argv[1]
cells = float(sys.argv[2]) / 32
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
cur = None; hdr = None; per = {}
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1].split('/')[-1]; continue
    if r[0] == "Line No": hdr = r; ii = hdr.index("Instructions Executed"); si = hdr.index("# Samples"); continue
    if hdr is None or r[0] in ("Function Name", "Kernel Name", ""): continue
    try: per[(cur, int(r[0]))] = (int(r[ii]), int(r[si]), r[1].strip())
    except Exception: pass
ts = sum(v[1] for v in per.values())
if len(sys.argv) > 3:
    f = sys.argv[3]
    ks = sorted([k for k in per if k[0] == f], key=lambda k: -per[k][1])[:int(sys.argv[4]) if len(sys.argv) > 4 else 40]
    for k in sorted(ks, key=lambda k: k[1]):
        print("%5d %7.1f inst %5.2f%% | %s" % (k[1], per[k][0] / cells, 100 * per[k][1] / ts, per[k][2][:110]))
else:
    agg = collections.defaultdict(lambda: [0, 0])
    for k, v in per.items():
        agg[k[0]][0] += v[0]; agg[k[0]][1] += v[1]
    for f, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-24s %7.1f inst/chunk %5.1f%% samples" % (f, v[0] / cells, 100 * v[1] / ts))
