#!/usr/bin/env python3
"""tools/make_stats_golden.py -- distribution fixtures for the native-RNG statistical parity tests
(container only; needs oracle/_ref/vcfgl_ref_dump, the instrumented build of the unmodified reference).

Runs the reference on synthetic msprime-shaped inputs, >= 1e6 cells per case (>= 1e5 sites for the per-site
beta draw of --error-qs 1), and stores SUMMARY COUNTS only (tests/golden/stats.json):
  depth histogram per cell (and per sample with --depths-file); true-base -> read-base matrix; haplotype pick;
  strand split; tail-distance histogram and which side (reference / non-reference allele) holds a site's tail
  mass; per-read qs histogram (--error-qs 2); mis-called reads per site (--error-qs 1: the beta-binomial law);
  genotype-call discordance (argmax-GL genotype vs true genotype, stratified hom/het like
  misc/gtDiscordance.cpp:11-15).
The GPU tests draw the same quantities from the Philox simulator -- once per native kernel set that takes the
case's flags -- and compare with two-sample chi-square, two-sample Kolmogorov-Smirnov (depth, quality score,
tail distance) and two-proportion z tests at alpha = 0.001.

usage: python tools/make_stats_golden.py [case ...]      (no argument: every case)
"""
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from vcfgl_b200 import synth  # noqa: E402
import vgl_dump  # noqa: E402

BIN_DUMP = os.path.join(ROOT, "oracle/_ref/vcfgl_ref_dump")
REF = "/root/reference"
OUT = os.path.join(ROOT, "tests/golden/stats.json")
DF_MEANS = [0.5, 2.0, 5.0, 12.0, 25.0]      # --depths-file: sample s has mean DF_MEANS[s % 5]

CASES = {
    # name: (n_sites, S, argv)
    "gl1_d10": (10000, 100, "--seed 42 -O u -d 10 -e 0.01 -GL 1 -addPL 1 -addFormatAD 1 -addFormatADF 1"),
    "gl1_d30": (10000, 100, "--seed 43 -O u -d 30 -e 0.01 -GL 1 -addPL 1 -addFormatAD 1"),
    "gl1_aux": (10000, 100, "--seed 48 -O u -d 10 -e 0.001 -GL 1 -doUnobserved 1 -addPL 1 -addI16 1 -addQS 1 -addInfoAD 1 -addInfoADF 1 -addInfoADR 1"),
    "gl1_df": (20000, 50, "--seed 49 -O u -df DEPTHSFILE -e 0.01 -GL 1 -addPL 1 -addFormatAD 1"),
    "gl2_d2_e02": (10000, 100, "--seed 44 -O u -d 2 -e 0.2 -GL 2 -addPL 1 -addFormatAD 1 -addFormatADF 1"),
    "gl2_eq2": (10000, 100, "--seed 45 -O u -d 5 -e 0.01 -eq 2 -bv 1e-5 -GL 2 -addPL 1 -addFormatAD 1"),
    "gl2_eq2_bins": (10000, 100, "--seed 46 -O u -d 5 -e 0.02 -eq 2 -bv 1e-4 -GL 2 -addPL 1 --qs-bins %s/test/data/rta3_qs_bins.csv" % REF),
    "gl2_eq1": (100000, 10, "--seed 47 -O u -d 5 -e 0.05 -eq 1 -bv 1e-3 -GL 2 -addPL 1 -addFormatAD 1"),
}
PAIRS = [(a1, a2) for a2 in range(5) for a1 in range(a2 + 1)]


def summarize(sites, S, per_sample_depth):
    depth_hist = np.zeros(200, np.int64)
    depth_by_sample = np.zeros((S, 64), np.int64) if per_sample_depth else None
    conf = np.zeros((4, 4), np.int64)      # [true base][read base], homozygous cells only (unambiguous truth)
    strand = np.zeros(2, np.int64)
    qs_hist = np.zeros(256, np.int64)
    tail_hist = np.zeros(32, np.int64)
    tail_side = np.zeros(2, np.int64)      # kept sites whose tail mass sits on the reference / a non-reference allele
    het_first = np.zeros(2, np.int64)
    site_err_hist = np.zeros(64, np.int64)  # mis-called reads among the site's homozygous cells
    site_reads_hom = 0
    disc = {"hom": [0, 0], "het": [0, 0]}  # [cells called, discordant]
    site_e = []
    n_sites = 0
    p1 = np.array([p[0] for p in PAIRS]), np.array([p[1] for p in PAIRS])
    for d in sites:
        n_sites += 1
        np.add.at(depth_hist, np.minimum(d.depths, 199), 1)
        if depth_by_sample is not None:
            np.add.at(depth_by_sample, (np.arange(S), np.minimum(d.depths, 63)), 1)
        if d.site_eprob is not None:
            site_e.append(d.site_eprob)
        gt = d.gts.reshape(S, 2)
        rs, rb = d.r_sample, d.r_base
        g0, g1 = gt[rs, 0], gt[rs, 1]
        hom = g0 == g1
        np.add.at(conf, (g0[hom], rb[hom]), 1)
        site_err_hist[min(int((rb[hom] != g0[hom]).sum()), 63)] += 1
        site_reads_hom += int(hom.sum())
        het = ~hom
        het_first[0] += int((rb[het] == g0[het]).sum())
        het_first[1] += int((rb[het] == g1[het]).sum())
        np.add.at(strand, d.r_strand, 1)
        if len(d.r_qs) and (d.r_qs >= 0).any():
            np.add.at(qs_hist, np.clip(d.r_qs, 0, 255), 1)
        if len(d.tails):
            np.add.at(tail_hist, np.clip(d.tails, 0, 31), 1)
        if d.ret == 0 and "i16" in d.out and d.info_dp > 0:
            i16 = d.out["i16"]
            tail_side[0] += int(i16[12] > 0)
            tail_side[1] += int(i16[14] > 0)
        if d.ret == 0 and "gl" in d.out and d.info_dp > 0:
            G = d.n_genotypes
            gl = d.out["gl"].reshape(S, G)
            a2b = d.alleles2acgt
            mx = gl.max(axis=1)
            called = d.fmt_dp > 0
            n_best = (gl == mx[:, None]).sum(axis=1)
            best = gl.argmax(axis=1)
            c0, c1 = a2b[p1[0][best]], a2b[p1[1][best]]
            call_lo, call_hi = np.minimum(c0, c1), np.maximum(c0, c1)
            t_lo, t_hi = np.minimum(gt[:, 0], gt[:, 1]), np.maximum(gt[:, 0], gt[:, 1])
            wrong = (n_best != 1) | (call_lo != t_lo) | (call_hi != t_hi)   # a tie is no call: discordant
            is_hom = t_lo == t_hi
            for k, m in (("hom", called & is_hom), ("het", called & ~is_hom)):
                disc[k][0] += int(m.sum())
                disc[k][1] += int((wrong & m).sum())
    out = dict(depth_hist=depth_hist.tolist(), confusion=conf.tolist(), strand=strand.tolist(),
               qs_hist=qs_hist.tolist(), tail_hist=tail_hist.tolist(), tail_side=tail_side.tolist(),
               het_reads=het_first.tolist(), discordance=disc, site_err_hist=site_err_hist.tolist(),
               site_reads_hom=site_reads_hom,
               site_eprob_mean=float(np.mean(site_e)) if site_e else None,
               site_eprob_var=float(np.var(site_e)) if site_e else None, n_site_eprob=len(site_e))
    if depth_by_sample is not None:
        out["depth_by_sample"] = depth_by_sample.tolist()
    assert n_sites > 0
    return out


def main():
    names = sys.argv[1:] or list(CASES)
    out = json.load(open(OUT)) if os.path.exists(OUT) and sys.argv[1:] else {}
    tmp = tempfile.mkdtemp(prefix="vgl_stats_")
    for name in names:
        n_sites, S, argv = CASES[name]
        vcf = os.path.join(tmp, name + ".vcf")
        hap = synth.sfs_genotypes(n_sites, S, 777)
        pos = synth.positions(n_sites, n_sites * 10, 777)
        synth.write_vcf(vcf, hap, pos, n_sites * 10)
        depths = None
        if "DEPTHSFILE" in argv:
            depths = [DF_MEANS[s % len(DF_MEANS)] for s in range(S)]
            dfile = os.path.join(tmp, name + ".depths")
            open(dfile, "w").write("".join("%g\n" % x for x in depths))
            argv = argv.replace("DEPTHSFILE", dfile)
        dump = os.path.join(tmp, name + ".vgld")
        env = dict(os.environ, VGL_DUMP_PATH=dump)
        r = subprocess.run([BIN_DUMP, "-i", vcf, "-o", os.path.join(tmp, name)] + argv.split(), env=env,
                           capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-2000:]
        st = summarize(vgl_dump.iter_dump(dump), S, depths is not None)
        st["argv"] = [x for x in argv.split()]
        st["n_sites"], st["S"], st["gt_seed"] = n_sites, S, 777
        if "--qs-bins" in st["argv"]:
            i = st["argv"].index("--qs-bins")
            st["qs_bins"] = [[0, 2, 2], [3, 14, 12], [15, 30, 23], [31, 40, 37]]
            del st["argv"][i:i + 2]
        if depths is not None:
            i = st["argv"].index("-df")
            del st["argv"][i:i + 2]
            st["depths"] = depths
        out[name] = st
        os.remove(dump)
        os.remove(vcf)
        print(name, "cells", n_sites * S, "reads", sum(st["strand"]), "disc", st["discordance"], flush=True)
    json.dump(out, open(OUT, "w"))
    print("wrote", OUT)


if __name__ == "__main__":
    main()
