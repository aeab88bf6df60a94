#!/usr/bin/env python3
"""tools/ncu_phases.py <rep> <cells> file:lo-hi=name ... -- warp instructions per 32-cell warp and stall samples per source range"""
import subprocess, csv, sys
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
tot = sum(v[0] for v in per.values()); ts = sum(v[1] for v in per.values())
used = set()
for spec in sys.argv[3:]:
    rng, name = spec.split("=")
    f, lh = rng.split(":"); lo, hi = map(int, lh.split("-"))
    ks = [k for k in per if k[0] == f and lo <= k[1] <= hi]
    used.update(ks)
    print("%-22s %7.1f inst/cell %5.1f%% samples" % (name, sum(per[k][0] for k in ks) / cells, 100 * sum(per[k][1] for k in ks) / ts))
rest = [k for k in per if k not in used]
print("%-22s %7.1f inst/cell %5.1f%% samples" % ("(other)", sum(per[k][0] for k in rest) / cells, 100 * sum(per[k][1] for k in rest) / ts))
print("total %.1f" % (tot / cells))
if "-v" in sys.argv:
    pass
