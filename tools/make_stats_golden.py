#!/usr/bin/env python3
"""tools/make_stats_golden.py -- distribution fixtures for the native-RNG statistical parity tests
(container only; needs oracle/_ref/vcfgl_ref_dump).

Runs the instrumented reference on synthetic msprime-shaped inputs (>= 1e6 cells over the cases)
and stores SUMMARY COUNTS only (tests/golden/stats.json):
  depth histogram per cell; true-base -> read-base matrix; strand split; per-read qs histogram
  (--error-qs 2); genotype-call discordance (argmax-GL genotype vs true genotype, stratified
  hom/het like misc/gtDiscordance.cpp:11-15).
The GPU tests draw the same quantities from the Philox simulator and compare with two-sample
chi-square / two-proportion z tests at alpha = 0.001.
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

CASES = {
    # name: (n_sites, S, argv)
    "gl1_d10": (4000, 100, "--seed 42 -O u -d 10 -e 0.01 -GL 1 -addPL 1 -addFormatAD 1 -addFormatADF 1"),
    "gl1_d30": (1500, 100, "--seed 43 -O u -d 30 -e 0.01 -GL 1 -addPL 1 -addFormatAD 1"),
    "gl2_d2_e02": (3000, 100, "--seed 44 -O u -d 2 -e 0.2 -GL 2 -addPL 1 -addFormatAD 1 -addFormatADF 1"),
    "gl2_eq2": (1500, 100, "--seed 45 -O u -d 5 -e 0.01 -eq 2 -bv 1e-5 -GL 2 -addPL 1 -addFormatAD 1"),
    "gl2_eq2_bins": (1500, 100, "--seed 46 -O u -d 5 -e 0.02 -eq 2 -bv 1e-4 -GL 2 -addPL 1 --qs-bins %s/test/data/rta3_qs_bins.csv" % REF),
    "gl2_eq1": (1500, 100, "--seed 47 -O u -d 5 -e 0.05 -eq 1 -bv 1e-3 -GL 2 -addPL 1 -addFormatAD 1"),
}


def call_stats(gl, gts, S, G):
    """argmax-GL genotype vs truth; genotypes as unordered allele-index pairs in VCF order"""
    pairs = [(a1, a2) for a2 in range(5) for a1 in range(a2 + 1)]
    return pairs


def summarize(sites, a2b_key="alleles2acgt"):
    depth_hist = np.zeros(200, np.int64)
    conf = np.zeros((4, 4), np.int64)      # [true base as drawn hap][read base] -- only hom cells are unambiguous
    strand = np.zeros(2, np.int64)
    qs_hist = np.zeros(256, np.int64)
    het_first = np.zeros(2, np.int64)      # reads of het cells equal to allele 1 / allele 2 (no-error approx, counts all)
    disc = {"hom": [0, 0], "het": [0, 0]}  # [n_cells_called, n_discordant]
    site_e = []
    pairs = [(a1, a2) for a2 in range(5) for a1 in range(a2 + 1)]
    for d in sites:
        S = d.S
        np.add.at(depth_hist, np.minimum(d.depths, 199), 1)
        if d.site_eprob is not None:
            site_e.append(d.site_eprob)
        gt = d.gts.reshape(S, 2)
        rs, rb = d.r_sample, d.r_base
        g0, g1 = gt[rs, 0], gt[rs, 1]
        hom = g0 == g1
        np.add.at(conf, (g0[hom], rb[hom]), 1)
        het = ~hom
        het_first[0] += int((rb[het] == g0[het]).sum())
        het_first[1] += int((rb[het] == g1[het]).sum())
        np.add.at(strand, d.r_strand, 1)
        if (d.r_qs >= 0).any():
            np.add.at(qs_hist, np.clip(d.r_qs, 0, 255), 1)
        if d.ret == 0 and "gl" in d.out and d.info_dp > 0:
            G = d.n_genotypes
            gl = d.out["gl"].reshape(S, G)
            a2b = d.alleles2acgt
            for s in range(S):
                if d.fmt_dp[s] == 0:
                    continue
                row = gl[s]
                best = np.flatnonzero(row == row.max())
                if len(best) != 1:
                    call = None      # tie: counted as discordant, like an uncalled genotype
                else:
                    a1, a2 = pairs[best[0]]
                    call = tuple(sorted((a2b[a1], a2b[a2])))
                truth = tuple(sorted((gt[s, 0], gt[s, 1])))
                k = "hom" if truth[0] == truth[1] else "het"
                disc[k][0] += 1
                disc[k][1] += int(call != truth)
    return dict(depth_hist=depth_hist.tolist(), confusion=conf.tolist(), strand=strand.tolist(),
                qs_hist=qs_hist.tolist(), het_reads=het_first.tolist(), discordance=disc,
                site_eprob_mean=float(np.mean(site_e)) if site_e else None,
                site_eprob_var=float(np.var(site_e)) if site_e else None, n_site_eprob=len(site_e))


def main():
    out = {}
    tmp = tempfile.mkdtemp(prefix="vgl_stats_")
    for name, (n_sites, S, argv) in CASES.items():
        vcf = os.path.join(tmp, name + ".vcf")
        hap = synth.sfs_genotypes(n_sites, S, 777)
        pos = synth.positions(n_sites, n_sites * 10, 777)
        synth.write_vcf(vcf, hap, pos, n_sites * 10)
        dump = os.path.join(tmp, name + ".vgld")
        env = dict(os.environ, VGL_DUMP_PATH=dump)
        r = subprocess.run([BIN_DUMP, "-i", vcf, "-o", os.path.join(tmp, name)] + argv.split(), env=env,
                           capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-2000:]
        sites = vgl_dump.read_dump(dump)
        st = summarize(sites)
        st["argv"] = [x for x in argv.split()]
        st["n_sites"], st["S"], st["gt_seed"] = n_sites, S, 777
        if "--qs-bins" in st["argv"]:
            i = st["argv"].index("--qs-bins")
            st["qs_bins"] = [[0, 2, 2], [3, 14, 12], [15, 30, 23], [31, 40, 37]]
            del st["argv"][i:i + 2]
        out[name] = st
        os.remove(dump)
        print(name, "cells", n_sites * S, "reads", sum(st["strand"]), "disc", st["discordance"])
    json.dump(out, open(os.path.join(ROOT, "tests/golden/stats.json"), "w"))
    print("wrote tests/golden/stats.json")


if __name__ == "__main__":
    main()
