"""Live check of the CLI mirror (vcfgl_b200/args.py validate(), beta_shape()) against the reference's own argument checks
(io.cpp:860-1000, rng.h:364-388): seeded random option combinations -- most of them invalid in one way or another -- are
handed to the unmodified reference binary on a tiny input; it must exit non-zero exactly when the mirror raises ArgError.

Container only: skipped where oracle/_ref does not exist."""
import os
import random
import subprocess

import pytest

from fuzz_cases import reference_exited
from vcfgl_b200 import args as vargs
from vcfgl_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "oracle", "_ref", "vcfgl_ref")

pytestmark = pytest.mark.skipif(not os.path.exists(BIN), reason="oracle/_ref not built (needs /root/reference)")

# (option, values in range, values out of range, probability the option is given)
OPTIONS = [("-d", ["0", "1", "5", "500", "inf"], ["501", "-2"], 0.97), ("-e", ["0", "0.01", "0.5"], ["1", "1.5", "-0.1"], 0.97),
           ("-eq", ["0", "1", "2"], ["3"], 0.4), ("-bv", ["1e-5", "1e-3"], ["0", "-1"], 0.35), ("-GL", ["1", "2"], ["3", "0"], 0.6),
           ("--gl1-theta", ["0", "0.83", "1"], ["1.5"], 0.15), ("--precise-gl", ["0", "1"], ["2"], 0.2),
           ("--adjust-qs", ["0", "1", "2", "3", "4", "8", "16", "31"], ["32"], 0.3), ("--adjust-by", ["0.499", "1"], ["0"], 0.15),
           ("--i16-mapq", ["0", "20", "60"], ["61"], 0.15), ("-doUnobserved", ["0", "1", "2", "3", "4", "5"], ["6"], 0.5),
           ("--rm-invar-sites", ["0", "1", "3", "7"], ["8"], 0.25), ("--rm-empty-sites", ["0", "1"], [], 0.2), ("-doGVCF", ["0", "1"], [], 0.15),
           ("--gvcf-dps", ["1,5,10", "3"], [], 0.15), ("-printPileup", ["0", "1"], [], 0.15), ("-printQScores", ["0", "1"], [], 0.1),
           ("-printGlError", ["0", "1"], [], 0.1)] + \
          [(t, ["0", "1"], [], 0.4) for t in ("-addGL", "-addGP", "-addPL", "-addI16", "-addQS", "-addFormatDP", "-addInfoDP", "-addFormatAD")]


def test_mirror_accepts_exactly_what_the_reference_accepts(tmp_path):
    vcf = str(tmp_path / "in.vcf")
    synth.write_vcf(vcf, synth.sfs_genotypes(6, 3, 1), synth.positions(6, 100, 1), 100)
    rnd = random.Random(4100)
    n_ok = n_bad = 0
    for k in range(160):
        argv = ["--seed", "3", "-O", "v"]
        for name, vals, bad, p in OPTIONS:
            if rnd.random() < p:
                argv += [name, rnd.choice(bad) if bad and rnd.random() < 0.04 else rnd.choice(vals)]
        try:
            a = vargs.parse_args(argv)
            if a.error_qs:
                vargs.beta_shape(a.error_rate, a.beta_variance)
            ok, why = True, ""
        except vargs.ArgError as e:
            ok, why = False, str(e)
        r = subprocess.run([BIN, "-i", vcf, "-o", str(tmp_path / "o")] + argv, capture_output=True, text=True)
        if r.returncode != 0 and ok and reference_exited(r.stderr):
            continue        # a run-time exit of the simulation itself, not an argument check
        assert (r.returncode == 0) == ok, (argv, why, " ".join(r.stderr.split("*******")[-2].split())[:400] if r.returncode else "")
        n_ok += ok
        n_bad += not ok
    assert n_ok >= 25 and n_bad >= 50, (n_ok, n_bad)
