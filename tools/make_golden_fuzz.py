#!/usr/bin/env python3
"""tools/make_golden_fuzz.py -- regenerate tests/golden/fuzz/ (container only; needs oracle/_ref).

Draws N seeded random hot-path configurations (tests/fuzz_cases.py), runs each through the unmodified
and the instrumented reference binary (their VCF outputs must agree), and stores the replay captures as
tests/golden/fuzz/fz<NN>.vgld.gz (+ the `-O u` file of the same run, fz<NN>.bcf.gz) with a manifest of the arguments (file arguments inlined), so the GPU
box can check the CUDA path against them without /root/reference."""
import gzip
import json
import os
import random
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import fuzz_cases  # noqa: E402

BIN = os.path.join(ROOT, "oracle/_ref/vcfgl_ref")
BIN_DUMP = os.path.join(ROOT, "oracle/_ref/vcfgl_ref_dump")
OUT = os.path.join(ROOT, "tests/golden/fuzz")
N_CASES = 32
SEED = 9000


def strip_header(path):
    return [l for l in open(path) if not l.startswith("##")]


def main():
    shutil.rmtree(OUT, ignore_errors=True)
    os.makedirs(OUT)
    tmp = tempfile.mkdtemp(prefix="vgl_fuzz_")
    rnd = random.Random(SEED)
    manifest = {}
    k = 0
    while len(manifest) < N_CASES:
        k += 1
        ref_argv, a, vcf, entry = fuzz_cases.draw_case(rnd, tmp, k)
        cid = "fz%02d" % len(manifest)
        dump = os.path.join(tmp, cid + ".vgld")
        ra = subprocess.run([BIN, "-i", vcf, "-o", os.path.join(tmp, cid + ".a")] + ref_argv, capture_output=True, text=True)
        rb = subprocess.run([BIN_DUMP, "-i", vcf, "-o", os.path.join(tmp, cid + ".b")] + ref_argv, capture_output=True, text=True,
                            env=dict(os.environ, VGL_DUMP_PATH=dump))
        if ra.returncode or rb.returncode:
            if fuzz_cases.reference_exited(ra.stderr):
                continue        # the reference's own run-time exits (fuzz_cases.REFERENCE_EXITS)
            raise SystemExit("reference failed on %s: %s" % (ref_argv, ra.stderr[-800:]))
        if strip_header(os.path.join(tmp, cid + ".a.vcf")) != strip_header(os.path.join(tmp, cid + ".b.vcf")):
            raise SystemExit("%s: instrumented and unmodified reference disagree" % cid)
        if not os.path.exists(dump) or os.path.getsize(dump) == 0:
            continue
        with gzip.GzipFile(os.path.join(OUT, cid + ".vgld.gz"), "wb", compresslevel=9, mtime=0) as g:
            g.write(open(dump, "rb").read())
        # the same run with -O u: the records the output-path tests compare with (tests/test_bcf_oracle.py, tests/test_gpu_bcf.py)
        u_argv = [("u" if i and ref_argv[i - 1] == "-O" else x) for i, x in enumerate(ref_argv)]
        ru = subprocess.run([BIN, "-i", vcf, "-o", os.path.join(tmp, cid + ".u")] + u_argv, capture_output=True, text=True)
        if ru.returncode:
            raise SystemExit("reference -O u failed on %s: %s" % (u_argv, ru.stderr[-800:]))
        with gzip.GzipFile(os.path.join(OUT, cid + ".bcf.gz"), "wb", compresslevel=9, mtime=0) as g:
            g.write(open(os.path.join(tmp, cid + ".u.bcf"), "rb").read())
        manifest[cid] = dict(entry, source="tools/make_golden_fuzz.py (seed %d, draw %d)" % (SEED, k),
                             pinned_by="unmodified reference binary run in the build container")
        print(cid, " ".join(entry["argv"]), "| dump", os.path.getsize(dump))
    json.dump(manifest, open(os.path.join(OUT, "manifest.json"), "w"), indent=1, sort_keys=True)
    shutil.rmtree(tmp, ignore_errors=True)
    print("wrote %d cases" % len(manifest))


if __name__ == "__main__":
    main()
