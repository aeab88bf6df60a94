"""Helpers of the --depth inf tests -- test infrastructure (imports oracle/)."""
import gzip
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import truth_oracle as to  # noqa: E402

import vcfin_oracle as vo  # noqa: E402
from test_vcfin_oracle import planned_sequence  # noqa: E402
from vcfgl_b200 import args as vargs  # noqa: E402
from vcfgl_b200 import vcfinput  # noqa: E402

TRUTH_DIR = os.path.join(ROOT, "tests", "golden", "truth")
MANIFEST = json.load(open(os.path.join(TRUTH_DIR, "manifest.json")))
CASES = sorted(c for c, m in MANIFEST.items() if not m["ref_failed"])
REFUSED = sorted(c for c, m in MANIFEST.items() if m["ref_failed"])


def case(cid):
    """-> (SimArgs, n_samples, [(pos, gts int8[2S])] in the reference's site order)"""
    m = MANIFEST[cid]
    a = vargs.parse_args(list(m["argv"]))
    buf = vo.load_input(m["input"])
    hdr = vcfinput.read_header(buf)
    S = len(hdr.samples)
    seq = planned_sequence(buf[hdr.body_offset:], S, a.source, a.explode, 0, hdr.contigs, max_run=1000)
    return a, S, seq


def reference_records(cid):
    """the reference's VCF -> [(pos0, [alleles], format keys, {key: values per sample as float arrays [S, n]})]"""
    out = []
    for line in gzip.open(os.path.join(TRUTH_DIR, cid + ".vcf.gz"), "rt"):
        if line.startswith("#"):
            continue
        f = line.rstrip("\n").split("\t")
        alleles = [f[3]] + (f[4].split(",") if f[4] != "." else [])
        keys = f[8].split(":")
        vals = {k: [] for k in keys}
        for col in f[9:]:
            for k, v in zip(keys, col.split(":")):
                vals[k].append([float(x) for x in v.split(",")])
        out.append((int(f[1]) - 1, alleles, keys, {k: np.array(v) for k, v in vals.items()}))
    return out
