#!/usr/bin/env python3
"""tools/prof_gvcf.py -- the gVCF block merger on one bench step of the cfg4 shape (100 samples, 99 % invariant sites).
usage: python tools/prof_gvcf.py [S] [n_sites] [reps]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vcfgl_b200 import args as vargs  # noqa: E402
from vcfgl_b200 import capi, synth  # noqa: E402

S = int(sys.argv[1]) if len(sys.argv) > 1 else 100
B = int(sys.argv[2]) if len(sys.argv) > 2 else 131072
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
a = vargs.parse_args("--seed 42 -d 10 -e 0.001 -GL 1 -doUnobserved 1 -doGVCF 1 --gvcf-dps 1,5,10 -addGL 1 -addPL 1 -addI16 1 -addQS 1".split())
hap = synth.sfs_genotypes(B, S, 20260004)
hap[np.random.default_rng(7).random(B) < 0.99] = 0
ctx = capi.Context(capi.params_from_args(a, S, max_batch_sites=B, n_slots=1, host_output=False))
ctx.input_buffer(0)[:] = synth.pack_gt(hap)
ctx.submit(0, 0, B)
b = ctx.wait(0)
rid, pos = np.zeros(B, np.int32), np.arange(B, dtype=np.int32)
ms = []
for _ in range(reps):
    r = ctx.gvcf_merge(0, rid, pos, [1, 5, 10])
    ms.append(r["ms_kernels"])
members = int(r["recs"]["n_members"].sum())
alg = 4 * B * S + 16 * members * S + 16 * r["n_blocks"] * S
print("S=%d sites=%d records=%d blocks=%d member sites=%d  kernels %s ms -> %.1f G member cells/s, %.0f GB/s algorithmic"
      % (S, B, len(r["recs"]), r["n_blocks"], members, ["%.3f" % x for x in ms], members * S / (min(ms) * 1e-3) / 1e9, alg / (min(ms) * 1e-3) / 1e9))
