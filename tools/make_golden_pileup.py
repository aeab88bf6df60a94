#!/usr/bin/env python3
"""tools/make_golden_pileup.py -- pileup fixtures (container only): the UNMODIFIED reference re-run with -printPileup 1 on
golden cases whose captures are already committed (the draws do not depend on the flag), output stored as
tests/golden/pileup/<id>.pileup.gz.  test10.pileup.gz is the reference's own golden file (test/reference/test10/)."""
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

OUT = os.path.join(ROOT, "tests/golden/pileup")
CASES = ["test1", "x_gl1_eq2_bins_adj", "x_gl2_eq1", "x_missing_gl1", "x_trim_rminvar", "x_acgt_multi", "x_gl2_eq2_precise1"]


def main():
    os.makedirs(OUT, exist_ok=True)
    with gzip.GzipFile(os.path.join(OUT, "test10.pileup.gz"), "wb", compresslevel=9, mtime=0) as g:   # the reference's own golden pileup, re-wrapped
        g.write(gzip.open(os.path.join(mg.REF, "test/reference/test10/test10.pileup.gz"), "rb").read())
    tmp = tempfile.mkdtemp(prefix="vgl_pileup_")
    cases = {t: (f, a) for t, f, a in mg.reference_tests()}
    cases.update({t: (f, a) for t, f, a in mg.extra_cases(tmp)})
    for cid in CASES:
        infile, argv = cases[cid]
        argv = [x for x in argv]
        if "-printPileup" in argv:
            argv[argv.index("-printPileup") + 1] = "1"
        else:
            argv += ["-printPileup", "1"]
        pref = os.path.join(tmp, cid)
        mg.run(mg.BIN, infile, argv, pref)
        raw = gzip.open(pref + ".pileup.gz", "rb").read()
        with gzip.GzipFile(os.path.join(OUT, cid + ".pileup.gz"), "wb", compresslevel=9, mtime=0) as g:
            g.write(raw)
        print(cid, len(raw), "bytes of pileup")
    shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    main()
