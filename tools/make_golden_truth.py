#!/usr/bin/env python3
"""tools/make_golden_truth.py -- fixtures of --depth inf (container only; needs oracle/_ref).

Runs the UNMODIFIED reference with --depth inf on input files that are already fixtures (tests/golden/inputs/) and stores
the VCF it wrote (tests/golden/truth/<id>.vcf.gz) with the arguments (manifest.json).  `ref_failed` marks runs the reference
refuses (a missing true genotype: ASSERT at vcfgl.cpp:1196)."""
import gzip
import json
import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
BIN = os.path.join(ROOT, "oracle/_ref/vcfgl_ref")
INPUTS = os.path.join(ROOT, "tests/golden/inputs")
OUT = os.path.join(ROOT, "tests/golden/truth")
TAGS = "-addGL 1 -addGP 1 -addPL 1 -addFormatDP 0"
CASES = [("t_test4", "data3.vcf", "--seed 42 -O v --depth inf --error-rate 0 --gl-model 1 --precise-gl 0 -explode 1 --rm-empty-sites 1 --adjust-qs 1 "
                                  "-doUnobserved 1 -addGP 1 -addPL 1 -addI16 0 -addQS 0 -addFormatDP 0")]
for u in range(6):
    CASES.append(("t_acgt_u%d" % u, "data5_acgt_multiallelic.vcf", "--seed 1 -O v --source 1 --depth inf -e 0 -GL 1 -doUnobserved %d %s" % (u, TAGS)))
    CASES.append(("t_s8_u%d_explode" % u, "s8.in.vcf", "--seed 1 -O v -explode 1 --depth inf -e 0 -GL 2 -doUnobserved %d %s" % (u, TAGS)))
CASES.append(("t_s40_pl_only", "s40.in.vcf", "--seed 1 -O v --depth inf -e 0 -GL 1 -doUnobserved 2 -addGL 0 -addPL 1 -addFormatDP 0"))
CASES.append(("t_missing_refused", "s8m.in.vcf", "--seed 1 -O v --depth inf -e 0 -GL 1 -doUnobserved 1 " + TAGS))


def main():
    shutil.rmtree(OUT, ignore_errors=True)
    os.makedirs(OUT)
    tmp = tempfile.mkdtemp(prefix="vgl_truth_")
    manifest = {}
    for cid, name, argline in CASES:
        path = os.path.join(tmp, name)
        open(path, "wb").write(gzip.open(os.path.join(INPUTS, name + ".gz"), "rb").read())
        argv = argline.split()
        r = subprocess.run([BIN, "-i", path, "-o", os.path.join(tmp, cid)] + argv, capture_output=True, text=True)
        ok = r.returncode == 0
        manifest[cid] = dict(input=name, argv=argv, ref_failed=not ok)
        if ok:
            with gzip.GzipFile(os.path.join(OUT, cid + ".vcf.gz"), "wb", compresslevel=9, mtime=0) as g:
                g.write(open(os.path.join(tmp, cid + ".vcf"), "rb").read())
        print(cid, "ok" if ok else "REFUSED: " + r.stderr.strip().splitlines()[-1][:100])
    json.dump(manifest, open(os.path.join(OUT, "manifest.json"), "w"), indent=1, sort_keys=True)
    shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    main()
