#!/usr/bin/env python3
"""tools/make_golden.py -- regenerate tests/golden/ (container only; needs /root/reference).

For every hot-path golden test of the reference (test/runTests.sh) this runs the
INSTRUMENTED reference binary (oracle/_ref/vcfgl_ref_dump), checks that its VCF
output is byte-identical (ignoring '##' header lines, like the reference's own
harness, runTests.sh:160-198) to test/reference/<id>/<id>.vcf, and stores the
replay capture as tests/golden/<id>.vgld.  It then adds configurations the
reference's tests do not cover (GL model 1 with per-read qs, --error-qs 1,
depth > 255, many samples, missing genotypes ...), for which the unmodified and
the instrumented binary must agree byte-for-byte.

tests/golden/manifest.json records, per case, the hot-path CLI arguments with
file arguments (--qs-bins, --depths-file) inlined, so nothing under
/root/reference is needed when the tests run.
"""
import json
import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vcfgl_b200 import args as vargs  # noqa: E402
from vcfgl_b200 import synth  # noqa: E402

REF = os.environ.get("REF", "/root/reference")
BIN = os.path.join(ROOT, "oracle/_ref/vcfgl_ref")
BIN_DUMP = os.path.join(ROOT, "oracle/_ref/vcfgl_ref_dump")
GOLD = os.path.join(ROOT, "tests/golden")

STUB = r'''
set -uo pipefail
SCRIPTDIR="%(ref)s/test"; DATADIR="$SCRIPTDIR/data"; TESTWD="@TESTWD@"; EXEC=x; TESTTYPE=regular
runTest(){ printf '%%s\x1f%%s\x1f%%s\n' "$1" "$2" "$4" | tr '\n' ' '; printf '\n'; }
runTestDiffVcf(){ :; }; runTestDiffPileup(){ :; }; runTestDiff(){ :; }
source <(awk '/^# TEST1$/{p=1} p' "$SCRIPTDIR/runTests.sh" | grep -v '^exit')
'''


def reference_tests():
    out = subprocess.run(["bash", "-c", STUB % dict(ref=REF)], capture_output=True, text=True).stdout
    tests = []
    for line in out.splitlines():
        if "\x1f" not in line:
            continue
        tid, infile, a = line.split("\x1f")
        argv = a.split()
        for i, x in enumerate(argv):  # file arguments are relative to the reference root
            if i and argv[i - 1] in ("--qs-bins", "--depths-file", "-df") and not os.path.isabs(x):
                argv[i] = os.path.join(REF, x)
        tests.append((tid.strip(), infile.strip(), argv))
    return tests


def strip_header(path):
    return [l for l in open(path) if not l.startswith("##")]


def run(binary, infile, argv, outpref, dump=None):
    env = dict(os.environ)
    if dump:
        env["VGL_DUMP_PATH"] = dump
    cmd = [binary, "-i", infile, "-o", outpref] + argv
    r = subprocess.run(cmd, capture_output=True, text=True, env=env)
    if r.returncode != 0:
        raise RuntimeError("reference failed: %s\n%s" % (" ".join(cmd), r.stderr[-2000:]))


def hot_args(argv):
    """keep the arguments the hot path reads; inline file arguments"""
    a = vargs.parse_args(argv)
    d = {"argv": [x for x in argv], "qs_bins": a.qs_bins, "depths": a.depths}
    # drop file paths from argv; tests re-attach inline values
    clean = []
    i = 0
    while i < len(argv):
        if argv[i] in ("--qs-bins", "--depths-file", "-df"):
            i += 2
            continue
        clean += argv[i:i + 2]
        i += 2
    d["argv"] = clean
    return d


def extra_cases(tmp):
    """configurations beyond the reference's own tests (oracle pinned by running the reference)"""
    bins = os.path.join(REF, "test/data/rta3_qs_bins.csv")
    alltags = "-addGP 1 -addPL 1 -addI16 1 -addQS 1 -addInfoDP 1 -addFormatAD 1 -addInfoAD 1 " \
              "-addFormatADF 1 -addInfoADF 1 -addFormatADR 1 -addInfoADR 1"
    cases = []

    def synth_vcf(name, n_sites, S, seed, missing=0.0, length=None):
        path = os.path.join(tmp, name + ".in.vcf")
        length = length or max(1000, n_sites * 10)
        hap = synth.sfs_genotypes(n_sites, S, seed, missing)
        pos = synth.positions(n_sites, length, seed)
        synth.write_vcf(path, hap, pos, length)
        return path

    v8 = synth_vcf("s8", 60, 8, 101)
    v8m = synth_vcf("s8m", 60, 8, 102, missing=0.15)
    v3 = synth_vcf("s3", 12, 3, 103)
    v40 = synth_vcf("s40", 60, 40, 104)
    cases.append(("x_gl1_eq2", v8, "--seed 42 -O v -d 6 -e 0.01 -eq 2 -bv 1e-5 -GL 1 -doUnobserved 1 " + alltags))
    cases.append(("x_gl1_eq2_bins_adj", v8, "--seed 7 -O v -d 5 -e 0.02 -eq 2 -bv 1e-4 -GL 1 --adjust-qs 3 "
                  "--qs-bins %s -doUnobserved 2 %s" % (bins, alltags)))
    cases.append(("x_gl2_eq2_bins", v8, "--seed 9 -O v -d 5 -e 0.01 -eq 2 -bv 1e-5 -GL 2 --qs-bins %s "
                  "-doUnobserved 1 -addPL 1 -addGP 1 -addQS 1 -addFormatAD 1" % bins))
    cases.append(("x_gl2_eq2_precise1", v8, "--seed 11 -O v -d 5 -e 0.01 -eq 2 -bv 1e-5 -GL 2 --precise-gl 1 "
                  "-doUnobserved 1 -addPL 1 -addGP 1 -addFormatAD 1"))
    cases.append(("x_gl2_eq1", v8, "--seed 13 -O v -d 4 -e 0.05 -eq 1 -bv 1e-4 -GL 2 -doUnobserved 1 " + alltags))
    cases.append(("x_gl1_eq1", v8, "--seed 15 -O v -d 4 -e 0.05 -eq 1 -bv 1e-4 -GL 1 -doUnobserved 5 " + alltags))
    cases.append(("x_gl2_precise1_eq0", v8, "--seed 17 -O v -d 3 -e 0.013 -GL 2 --precise-gl 1 -doUnobserved 0 "
                  "-addPL 1 -addGP 1 -addFormatAD 1 -addInfoAD 1"))
    cases.append(("x_gl1_d300", v3, "--seed 19 -O v -d 300 -e 0.01 -GL 1 -doUnobserved 1 -addPL 1 -addFormatAD 1 -addQS 1 -addI16 1"))
    cases.append(("x_gl1_d300_eq2", v3, "--seed 21 -O v -d 290 -e 0.02 -eq 2 -bv 1e-4 -GL 1 -doUnobserved 1 -addPL 1 -addFormatAD 1"))
    cases.append(("x_gl2_d300", v3, "--seed 23 -O v -d 300 -e 0.01 -GL 2 -doUnobserved 1 -addPL 1 -addFormatAD 1"))
    cases.append(("x_s40_gl1_cfg2", v40, "--seed 42 -O v -d 10 -e 0.01 -GL 1 -addGL 1 -addPL 1 -addFormatAD 1"))
    cases.append(("x_s40_gl1_d30_cfg5", v40, "--seed 42 -O v -d 30 -e 0.01 -GL 1 -addPL 1 -addFormatAD 1"))
    cases.append(("x_s40_gl2_cfg3", v40, "--seed 42 -O v -d 10 -e 0.01 -GL 2 -eq 2 -bv 1e-5 --qs-bins %s -addPL 1" % bins))
    cases.append(("x_s40_gl2_eq1_cfg3", v40, "--seed 42 -O v -d 10 -e 0.01 -GL 2 -eq 1 -bv 1e-5 -addPL 1"))
    cases.append(("x_s40_cfg4tags", v40, "--seed 42 -O v -d 10 -e 0.001 -GL 1 -doUnobserved 1 -addPL 1 -addI16 1 -addQS 1"))
    cases.append(("x_missing_gl1", v8m, "--seed 25 -O v -d 3 -e 0.05 -GL 1 -doUnobserved 1 " + alltags))
    cases.append(("x_missing_gl2", v8m, "--seed 27 -O v -d 0.7 -e 0.05 -GL 2 -doUnobserved 4 --rm-empty-sites 0 " + alltags))
    cases.append(("x_trim_rminvar", v8, "--seed 29 -O v -d 2 -e 0.1 -GL 2 -doUnobserved 0 --rm-invar-sites 4 --rm-empty-sites 1 "
                  "-addPL 1 -addGP 1 -addQS 1 -addInfoDP 1 -addFormatAD 1 -addInfoAD 1 -addFormatADF 1 -addInfoADF 1 -addFormatADR 1 -addInfoADR 1"))
    cases.append(("x_explode3_e05", v8, "--seed 31 -O v -d 8 -e 0.5 -GL 1 -doUnobserved 3 -addPL 1 -addFormatAD 1 -addQS 1 -addI16 1"))
    cases.append(("x_e099", v8, "--seed 33 -O v -d 3 -e 0.99 -GL 2 -doUnobserved 1 -addPL 1 -addGP 1 -addFormatAD 1"))
    cases.append(("x_e0_gl2", v8, "--seed 35 -O v -d 3 -e 0 -GL 2 -doUnobserved 1 -addPL 1 -addGP 1 -addFormatAD 1"))
    cases.append(("x_acgt_multi", os.path.join(REF, "test/data/data5_acgt_multiallelic.vcf"),
                  "--seed 37 -O v --source 1 -d 5 -e 0.02 -GL 1 -doUnobserved 1 " + alltags))
    return [(n, f, a.split()) for n, f, a in cases]


def main():
    shutil.rmtree(GOLD, ignore_errors=True)
    os.makedirs(GOLD)
    manifest = {}
    tmp = tempfile.mkdtemp(prefix="vgl_golden_")
    n_ok = 0
    for tid, infile, argv in reference_tests():
        a = vargs.parse_args([x for x in argv])
        if a.depth == float("inf"):
            print(tid, "skipped (--depth inf: truth mode, not the hot path)")
            continue
        if a.output_mode != "v":
            print(tid, "skipped (run-only BCF/threads test)")
            continue
        pref = os.path.join(tmp, tid)
        dump = os.path.join(GOLD, tid + ".vgld")
        run(BIN_DUMP, infile, argv, pref, dump)
        got = strip_header(pref + ".vcf")
        want = strip_header(os.path.join(REF, "test/reference", tid, tid + ".vcf"))
        if got != want:
            raise SystemExit("%s: instrumented reference output differs from the reference golden VCF" % tid)
        manifest[tid] = dict(hot_args(argv), source="reference test/runTests.sh", input=os.path.basename(infile),
                             pinned_by="test/reference/%s/%s.vcf" % (tid, tid))
        n_ok += 1
        print(tid, "OK (VCF identical to reference golden; dump %d bytes)" % os.path.getsize(dump))
    for tid, infile, argv in extra_cases(tmp):
        pref_a = os.path.join(tmp, tid + ".a")
        pref_b = os.path.join(tmp, tid + ".b")
        dump = os.path.join(GOLD, tid + ".vgld")
        run(BIN, infile, argv, pref_a)
        run(BIN_DUMP, infile, argv, pref_b, dump)
        if strip_header(pref_a + ".vcf") != strip_header(pref_b + ".vcf"):
            raise SystemExit("%s: instrumented and unmodified reference disagree" % tid)
        manifest[tid] = dict(hot_args(argv), source="tools/make_golden.py extra case", input=os.path.basename(infile),
                             pinned_by="unmodified reference binary run in the build container")
        print(tid, "OK (unmodified == instrumented; dump %d bytes)" % os.path.getsize(dump))
    import gzip
    for f in sorted(os.listdir(GOLD)):
        if f.endswith(".vgld"):
            raw = open(os.path.join(GOLD, f), "rb").read()
            with gzip.GzipFile(os.path.join(GOLD, f + ".gz"), "wb", compresslevel=9, mtime=0) as g:
                g.write(raw)
            os.remove(os.path.join(GOLD, f))
    json.dump(manifest, open(os.path.join(GOLD, "manifest.json"), "w"), indent=1, sort_keys=True)
    shutil.rmtree(tmp, ignore_errors=True)
    print("wrote %d cases (%d pinned by reference golden VCFs)" % (len(manifest), n_ok))


if __name__ == "__main__":
    main()
