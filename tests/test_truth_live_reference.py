"""Live pin of oracle/truth_oracle.py (--depth inf, vcfgl.cpp:1089-1262) on the unmodified reference binary: seeded
random VCF texts (tests/test_vcfin_live_reference.random_vcf, without missing genotypes -- the reference asserts on
those, vcfgl.cpp:1196) with random --source / -explode / -doUnobserved 0-5 / GL, GP, PL subsets; every record of the VCF
the reference writes must be the oracle's: alleles and their order, GL / GP / PL of every sample.

Container only: skipped where oracle/_ref does not exist (the GPU box uses tests/golden/truth/ instead)."""
import os
import random
import subprocess

import numpy as np
import pytest

import truth_util as tu
from test_vcfin_live_reference import random_vcf
from test_vcfin_oracle import planned_sequence
from vcfgl_b200 import args as vargs
from vcfgl_b200 import vcfinput

BIN = os.path.join(tu.ROOT, "oracle", "_ref", "vcfgl_ref")

pytestmark = pytest.mark.skipif(not os.path.exists(BIN), reason="oracle/_ref not built (needs /root/reference)")


def records_of(path):
    out = []
    for line in open(path):
        if line.startswith("#"):
            continue
        f = line.rstrip("\n").split("\t")
        keys = f[8].split(":")
        vals = {k: [] for k in keys}
        for col in f[9:]:
            for k, v in zip(keys, col.split(":")):
                vals[k].append([float("nan") if x == "." else float(x) for x in v.split(",")])
        out.append((int(f[1]) - 1, [f[3]] + (f[4].split(",") if f[4] != "." else []), keys, {k: np.array(v) for k, v in vals.items()}))
    return out


@pytest.mark.parametrize("block", range(3))
def test_truth_oracle_equals_live_reference(block, tmp_path):
    rnd = random.Random(9700 + block)
    n_records = 0
    for k in range(12):
        acgt = rnd.random() < 0.6
        while True:
            S, buf = random_vcf(rnd, acgt)
            if b"." not in b"".join(l.split(b"\t", 9)[9] for l in buf.split(b"\n") if l and not l.startswith(b"#")):
                break
        tags = rnd.choice([("GL",), ("PL",), ("GL", "PL"), ("GL", "GP", "PL"), ("GP",)])
        argv = ["--seed", "1", "-O", "v", "--source", str(int(acgt)), "-explode", str(rnd.choice([0, 0, 1])), "--depth", "inf", "-e", "0",
                "-GL", str(rnd.choice([1, 2])), "-doUnobserved", str(rnd.randrange(0, 6)), "-addFormatDP", "0"]
        for t in ("GL", "GP", "PL"):
            argv += ["-add" + t, str(int(t in tags))]
        a = vargs.parse_args(argv)
        vcf = str(tmp_path / ("in%d.vcf" % k))
        open(vcf, "wb").write(buf)
        r = subprocess.run([BIN, "-i", vcf, "-o", str(tmp_path / ("o%d" % k))] + argv, capture_output=True, text=True)
        where = (argv, buf.decode())
        assert r.returncode == 0, (where, r.stderr[-1500:])
        recs = records_of(str(tmp_path / ("o%d.vcf" % k)))
        hdr = vcfinput.read_header(buf)
        seq = planned_sequence(buf[hdr.body_offset:], S, a.source, a.explode, 0, hdr.contigs, max_run=1000)
        assert [p for p, _ in seq] == [r_[0] for r_ in recs], where
        for (pos, gts), (rpos, alleles, keys, vals) in zip(seq, recs):
            o = tu.to.site(gts, a.do_unobserved)
            assert o["alleles"] == alleles, (pos, o["alleles"], alleles, where)
            # FORMAT keys the input record carried besides GT (DP, GQ here) ride along in front of the simulated ones
            assert [t for t in keys if t not in ("DP", "GQ")] == [t for t in ("GL", "GP", "PL") if t in tags]
            G = o["n_genotypes"]
            for t in tags:
                assert np.array_equal(vals[t], o[t.lower()].reshape(S, G).astype(np.float64)), (pos, t, where)
        n_records += len(recs)
    assert n_records > 100
