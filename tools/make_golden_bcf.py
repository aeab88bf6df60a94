#!/usr/bin/env python3
"""tools/make_golden_bcf.py -- BCF fixtures for the output-path parity tests (container only; needs
/root/reference and oracle/_ref/vcfgl_ref built by oracle/build_ref.sh).

For every case of tests/golden/manifest.json this re-runs the UNMODIFIED reference binary with the
case's arguments and `-O u` (uncompressed BCF, the format bench.py's reference arm writes) and stores
the file (gzip-wrapped) as tests/golden/bcf/<id>.bcf.gz.  The reference's RNG streams are seeded, so the records are the
BCF encoding of exactly the sites captured in tests/golden/<id>.vgld.gz (tests/test_bcf_oracle.py checks
that: positions, allele strings and every tag value agree).  The committed replay captures are not touched.
"""
import gzip
import os
import shutil
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import make_golden as mg  # noqa: E402


def with_bcf_output(argv):
    out = list(argv)
    for i, x in enumerate(out[:-1]):
        if x in ("-O", "--output-mode"):
            out[i + 1] = "u"
            return out
    return out + ["-O", "u"]


def main():
    dst = os.path.join(mg.GOLD, "bcf")
    os.makedirs(dst, exist_ok=True)
    tmp = tempfile.mkdtemp(prefix="vgl_golden_bcf_")
    import json
    manifest = json.load(open(os.path.join(mg.GOLD, "manifest.json")))
    cases = [(t, f, a) for t, f, a in mg.reference_tests() if t in manifest] + mg.extra_cases(tmp)
    n = 0
    for tid, infile, argv in cases:
        if tid not in manifest:
            raise SystemExit("%s is not in tests/golden/manifest.json" % tid)
        pref = os.path.join(tmp, tid)
        mg.run(mg.BIN, infile, with_bcf_output(argv), pref)
        with gzip.GzipFile(os.path.join(dst, tid + ".bcf.gz"), "wb", compresslevel=9, mtime=0) as g:
            g.write(open(pref + ".bcf", "rb").read())
        n += 1
        print(tid, os.path.getsize(pref + ".bcf"), "bytes")
    shutil.rmtree(tmp, ignore_errors=True)
    print("wrote %d BCF fixtures to %s" % (n, dst))


if __name__ == "__main__":
    main()
