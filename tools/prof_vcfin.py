#!/usr/bin/env python3
"""tools/prof_vcfin.py -- the input path alone on one step of msprime-shaped text (for ncu and quick timing).
usage: python tools/prof_vcfin.py [S] [n_records] [reps]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vcfgl_b200 import args as vargs  # noqa: E402
from vcfgl_b200 import capi, synth  # noqa: E402

S = int(sys.argv[1]) if len(sys.argv) > 1 else 100
B = int(sys.argv[2]) if len(sys.argv) > 2 else 131072
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
hap = synth.sfs_genotypes(B, S, 20260002)
body = synth.vcf_body(hap, np.arange(1, B + 1) * 7)
a = vargs.parse_args("--seed 42 -d 10 -e 0.01 -GL 1 -addGL 1 -addPL 1 -addFormatAD 1".split())
ctx = capi.Context(capi.params_from_args(a, S, max_batch_sites=min(B, 4096), n_slots=1, host_output=False))
ps = ctx.parser(len(body) + 64, B)
t0 = time.perf_counter()
r = ps.parse(body, 0, 0)
t1 = time.perf_counter()
assert r.n_records == B and r.n_errors == 0, (r.n_records, r.n_errors, r.first_error_record)
assert np.array_equal(ps.rows(0, min(B, 4096)), synth.pack_gt(hap)[:min(B, 4096)])
ms = [ps.parse(None, 0, capi.PARSE_TEXT_ON_DEVICE, n_bytes=len(body)).ms_kernels for _ in range(reps)]
alg = len(body) + B * S + 64 * B
print("S=%d records=%d text=%.1f MB  first parse %.1f ms (h2d %.2f ms)  kernels %s ms  -> %.1f G cells/s, %.0f GB/s algorithmic"
      % (S, B, len(body) / 1e6, 1e3 * (t1 - t0), r.ms_h2d, ["%.3f" % x for x in ms], B * S / (min(ms) * 1e-3) / 1e9,
         alg / (min(ms) * 1e-3) / 1e9))
