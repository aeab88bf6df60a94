"""Seeded random hot-path configurations -- test infrastructure shared by tests/test_oracle_live_reference.py
(runs them through the reference in the build container) and tools/make_golden_fuzz.py (commits a fixed
set of their captures under tests/golden/fuzz/ for the GPU parity tests)."""
import os

from vcfgl_b200 import args as vargs
from vcfgl_b200 import synth

RTA3_BINS = [(0, 2, 2), (3, 14, 12), (15, 30, 23), (31, 40, 37)]      # shape of test/data/rta3_qs_bins.csv
WIDE_BINS = [(0, 2, 2), (3, 14, 12), (15, 30, 23), (31, 63, 37)]      # the same with every capped score covered
TAGS = ["-addGL", "-addGP", "-addPL", "-addI16", "-addQS", "-addFormatDP", "-addInfoDP", "-addFormatAD", "-addInfoAD",
        "-addFormatADF", "-addInfoADF", "-addFormatADR", "-addInfoADR"]


def draw_case(rnd, tmp, k):
    """one random, valid hot-path configuration: (argv for the reference, SimArgs, input VCF, manifest entry)"""
    while True:
        S = rnd.choice([1, 2, 5, 9, 33])
        n_sites = rnd.choice([8, 25, 40])
        gl = rnd.choice([1, 2])
        eq = rnd.choice([0, 0, 1, 2])
        argv = ["--seed", str(rnd.randrange(1, 10000)), "-O", "v", "-GL", str(gl), "-eq", str(eq),
                "-e", rnd.choice(["0.001", "0.01", "0.05", "0.2"]) if eq else rnd.choice(["0", "0.002", "0.01", "0.1", "0.6"]),
                "-doUnobserved", str(rnd.randrange(0, 6)),
                "--rm-invar-sites", str(rnd.choice([0, 0, 1, 2, 3, 4, 7])),
                "--rm-empty-sites", str(rnd.choice([0, 1]))]
        if eq:
            argv += ["-bv", rnd.choice(["1e-5", "1e-4", "1e-3"])]
        if gl == 2 and rnd.random() < 0.4:
            argv += ["--precise-gl", "1"]
        adj = rnd.choice([0, 0, 1, 2, 3])
        if adj:
            argv += ["--adjust-qs", str(adj)]
            if rnd.random() < 0.3:
                argv += ["--adjust-by", rnd.choice(["0.2", "0.9"])]
        for t in TAGS:
            if rnd.random() < 0.6:
                argv += [t, "1"]
        if "-addI16" in argv and rnd.random() < 0.3:
            argv += ["--i16-mapq", str(rnd.choice([0, 37, 60]))]
        depths = None
        if rnd.random() < 0.25:
            depths = [rnd.choice([0.3, 1.0, 4.0, 11.0, 14.5, 40.0]) for _ in range(S)]
        else:
            argv += ["-d", rnd.choice(["0.4", "2", "7", "11.9", "12", "25", "270"] if S < 9 else ["0.4", "2", "7", "12", "25"])]
        bins = rnd.choice([RTA3_BINS, WIDE_BINS]) if (eq == 2 and rnd.random() < 0.5) else None
        try:
            a = vargs.parse_args(argv, qs_bins=bins, depths=depths)
            if eq:
                vargs.beta_shape(a.error_rate, a.beta_variance)     # rng.h:373-388: the reference exits on non-positive shapes
        except vargs.ArgError:
            continue
        ref_argv = list(argv)
        if depths:
            path = os.path.join(tmp, "depths%d.txt" % k)
            with open(path, "w") as f:
                f.write("".join("%g\n" % d for d in depths))
            ref_argv += ["--depths-file", path]
        if bins:
            path = os.path.join(tmp, "bins%d.csv" % k)
            with open(path, "w") as f:
                f.write("".join("%d,%d,%d\n" % b for b in bins))
            ref_argv += ["--qs-bins", path]
        vcf = os.path.join(tmp, "in%d.vcf" % k)
        hap = synth.sfs_genotypes(n_sites, S, 500 + k, rnd.choice([0.0, 0.0, 0.2]))
        synth.write_vcf(vcf, hap, synth.positions(n_sites, 1000, 500 + k), 1000)
        return ref_argv, a, vcf, dict(argv=argv, qs_bins=bins, depths=depths)


# Runs the reference itself ends with exit(1); such draws are skipped, not compared:
#   apply_qs_bins(), vcfgl.cpp:63  -- a quality score outside every --qs-bins range (libvgl: batch status VGL_ERANGE)
#   ASSERT(sim->nAlleles > 1), vcfgl.cpp:1038 -- -addI16 1 at a site with one allele (-doUnobserved 0, one base observed)
#   ASSERT(adjqScore_i != -1), vcfgl.cpp:558 -- --error-qs 2 with --adjust-qs: a beta draw of exactly 0.0 (tiny alpha) leaves the
#       adjusted score unset (vcfgl.cpp:499-506)
#   bcf_enc_vfloat: Assertion `n >= 0' (htslib/vcf.c:2339) -- -doGVCF 1 with -addQS 1 on some low-depth runs (a crash of the reference)
#   bcf_update_format: Assertion `nps && nps*line->n_sample==n' (htslib/vcf.c:4452) -- -doGVCF 1 on some multi-allelic / missing-genotype
#       inputs (the reference itself warns that gVCF mode is under development, io.cpp:964)
REFERENCE_EXITS = ("Could not find a range for qs value", "sim->nAlleles > 1", "adjqScore_i != -1", "bcf_enc_vfloat: Assertion",
                   "bcf_update_format: Assertion")


def reference_exited(stderr: str) -> bool:
    return any(m in stderr for m in REFERENCE_EXITS)
