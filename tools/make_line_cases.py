#!/usr/bin/env python3
"""tools/make_line_cases.py -- what the UNMODIFIED reference does with the single-record cases of tests/vcfin_lines.py
(container only).  Each record is wrapped in a header and run through oracle/_ref/vcfgl_ref_dump; the fixture
tests/golden/inputs/line_cases.json records whether the reference exited with an error and, if it ran, the
true_gts_acgt_int it simulated from -- the input-path oracle's status codes and genotypes are pinned on it
(tests/test_vcfin_oracle.py::test_lines_against_reference)."""
import json
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import vcfin_lines  # noqa: E402
import vgl_dump  # noqa: E402

BIN_DUMP = os.path.join(ROOT, "oracle/_ref/vcfgl_ref_dump")
HDR = ("##fileformat=VCFv4.2\n##FILTER=<ID=PASS,Description=\"x\">\n##FILTER=<ID=q,Description=\"x\">\n"
       "##contig=<ID=1,length=100>\n##contig=<ID=c,length=100>\n##INFO=<ID=X,Number=1,Type=Integer,Description=\"x\">\n"
       "##FORMAT=<ID=GT,Number=1,Type=String,Description=\"Genotype\">\n"
       "##FORMAT=<ID=DP,Number=1,Type=Integer,Description=\"d\">\n##FORMAT=<ID=GQ,Number=1,Type=Integer,Description=\"q\">\n"
       "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t%s\n")


def run_line(tmp, k, line, S, source):
    path = os.path.join(tmp, "l%d.vcf" % k)
    with open(path, "wb") as fh:
        fh.write((HDR % "\t".join("s%d" % i for i in range(S))).encode())
        fh.write(line + b"\n")
    dump = os.path.join(tmp, "l%d.vgld" % k)
    env = dict(os.environ, VGL_DUMP_PATH=dump)
    r = subprocess.run([BIN_DUMP, "-i", path, "-o", os.path.join(tmp, "o%d" % k), "--seed", "1", "-O", "v", "--source", str(source),
                        "-d", "1", "-e", "0.01", "-GL", "2"], capture_output=True, env=env, timeout=60)
    ok = r.returncode == 0
    sites = vgl_dump.read_dump(dump) if ok and os.path.exists(dump) and os.path.getsize(dump) else []
    return dict(line=line.decode("latin1"), S=S, source=source, ref_ok=ok,
                sites=[[int(d.pos), [int(x) for x in d.gts]] for d in sites])


def main():
    tmp = tempfile.mkdtemp(prefix="vgl_lines_")
    out = {"bad": [], "good": []}
    for k, (line, S, source, want) in enumerate(vcfin_lines.BAD):
        c = run_line(tmp, k, line, S, source)
        c["oracle_status"] = int(want)
        out["bad"].append(c)
        print("bad  %2d status %2d ref_ok=%s %r" % (k, want, c["ref_ok"], line[:60]))
    for k, (line, S, source, gts, pos, n_allele) in enumerate(vcfin_lines.GOOD):
        c = run_line(tmp, 100 + k, line, S, source)
        out["good"].append(c)
        print("good %2d ref_ok=%s sites=%s" % (k, c["ref_ok"], c["sites"]))
    json.dump(out, open(os.path.join(ROOT, "tests/golden/inputs/line_cases.json"), "w"), indent=0, sort_keys=True)


if __name__ == "__main__":
    main()
