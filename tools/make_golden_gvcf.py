#!/usr/bin/env python3
"""tools/make_golden_gvcf.py -- fixtures of the gVCF block merger (container only; needs oracle/_ref).

For each case below the reference is run twice with the same seed: the instrumented binary gives the per-site capture
(tests/golden/gvcf/<id>.vgld.gz: what every site's tags were BEFORE merging), the unmodified binary writes the merged
output with -O u (tests/golden/gvcf/<id>.bcf.gz).  tests/test_gvcf_oracle.py derives the blocks from the capture with
oracle/gvcf_oracle.py and must find exactly the records of the BCF (positions, END, MIN_DP, per-sample DP and PL).
The gVCF cases of the main golden set (test7, test8, test19) are pinned the same way from tests/golden/ + tests/golden/bcf/.
"""
import gzip
import json
import os
import shutil
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import make_golden as mg  # noqa: E402
from make_golden_bcf import with_bcf_output  # noqa: E402
from vcfgl_b200 import synth  # noqa: E402

OUT = os.path.join(ROOT, "tests/golden/gvcf")
TAGS = "-addPL 1 -addQS 1 -addInfoDP 1 -addFormatAD 1"
CASES = [
    # id, (n_sites, S, seed, contig length, missing rate), reference arguments
    ("g_explode_dps135", (25, 6, 301, 400, 0.0), "--seed 11 -O v -explode 1 -d 3 -e 0.01 -GL 1 -doUnobserved 1 -doGVCF 1 --gvcf-dps 1,3,5 " + TAGS),
    ("g_explode_dps2_gl2", (25, 6, 302, 400, 0.0), "--seed 12 -O v -explode 1 -d 4 -e 0.02 -GL 2 -doUnobserved 2 -doGVCF 1 --gvcf-dps 2 " + TAGS + " -addI16 1"),
    ("g_gaps_dps1", (120, 4, 303, 300, 0.0), "--seed 13 -O v -d 2 -e 0.005 -GL 1 -doUnobserved 1 -doGVCF 1 --gvcf-dps 1,2,4,8 " + TAGS),
    ("g_lowdepth_empty", (30, 3, 304, 500, 0.1), "--seed 14 -O v -explode 1 -d 1.6 -e 0.01 -GL 1 -doUnobserved 2 -doGVCF 1 --gvcf-dps 1,2 --rm-empty-sites 1 " + TAGS),
    ("g_s20_dps1510", (40, 20, 305, 600, 0.0), "--seed 15 -O v -explode 1 -d 10 -e 0.001 -GL 1 -doUnobserved 1 -doGVCF 1 --gvcf-dps 1,5,10 -addPL 1 -addI16 1 -addQS 1"),
]


def main():
    shutil.rmtree(OUT, ignore_errors=True)
    os.makedirs(OUT)
    tmp = tempfile.mkdtemp(prefix="vgl_gvcf_")
    manifest = {}
    for cid, (n_sites, S, seed, length, miss), argline in CASES:
        path = os.path.join(tmp, cid + ".in.vcf")
        synth.write_vcf(path, synth.sfs_genotypes(n_sites, S, seed, miss), synth.positions(n_sites, length, seed), length)
        argv = argline.split()
        dump = os.path.join(tmp, cid + ".vgld")
        mg.run(mg.BIN_DUMP, path, argv, os.path.join(tmp, cid + ".a"), dump)
        mg.run(mg.BIN, path, with_bcf_output(argv), os.path.join(tmp, cid + ".b"))
        for src, dst in ((dump, cid + ".vgld.gz"), (os.path.join(tmp, cid + ".b.bcf"), cid + ".bcf.gz")):
            with gzip.GzipFile(os.path.join(OUT, dst), "wb", compresslevel=9, mtime=0) as g:
                g.write(open(src, "rb").read())
        manifest[cid] = dict(argv=argv, n_samples=S, pinned_by="unmodified + instrumented reference binaries, same seed")
        print(cid, os.path.getsize(os.path.join(OUT, cid + ".bcf.gz")), "bytes of BCF")
    json.dump(manifest, open(os.path.join(OUT, "manifest.json"), "w"), indent=1, sort_keys=True)
    shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    main()
