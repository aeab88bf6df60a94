#!/usr/bin/env python3
"""tools/prof_bcfin.py -- the BCF input path (k_bcf_gt) on msprime-shaped records.  usage: python tools/prof_bcfin.py [S] [n_records] [reps]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bcf_writer as bw  # noqa: E402  (test infrastructure: only used to make the input)
from vcfgl_b200 import args as vargs  # noqa: E402
from vcfgl_b200 import capi, synth  # noqa: E402

S = int(sys.argv[1]) if len(sys.argv) > 1 else 100
B = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
hap = synth.sfs_genotypes(B, S, 20260002)
# records are identical in shape: encode one per distinct genotype row cheaply = encode all (python, slow) for B <= 32768
buf = synth.vcf_header(S, B * 10) + synth.vcf_body(hap, np.arange(1, B + 1) * 7)
bcf, first, offs, ids = bw.vcf_to_bcf(buf)
body, off = bcf[first:], np.array(offs, np.uint32)
a = vargs.parse_args("--seed 42 -d 10 -e 0.01 -GL 1 -addGL 1 -addPL 1 -addFormatAD 1".split())
ctx = capi.Context(capi.params_from_args(a, S, max_batch_sites=min(B, 4096), n_slots=1, host_output=False))
ps = ctx.parser(len(body) + 64, B)
r = ps.parse_bcf(body, off, ids["GT"], 0)
assert r.n_records == B and r.n_errors == 0
assert np.array_equal(ps.rows(0, min(B, 4096)), synth.pack_gt(hap)[:min(B, 4096)])
ms = [ps.parse_bcf(body, off, ids["GT"], 0, capi.PARSE_TEXT_ON_DEVICE).ms_kernels for _ in range(reps)]
alg = len(body) + B * S + 64 * B
print("BCF S=%d records=%d bytes=%.1f MB  kernel %s ms  -> %.1f G cells/s, %.0f GB/s algorithmic"
      % (S, B, len(body) / 1e6, ["%.3f" % x for x in ms], B * S / (min(ms) * 1e-3) / 1e9, alg / (min(ms) * 1e-3) / 1e9))
