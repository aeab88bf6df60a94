"""Live pin of oracle/gvcf_oracle.py (block merger, bcf_utils.cpp:662-942) on the reference binaries built by
oracle/build_ref.sh: seeded random -doGVCF 1 runs (depth bins, -explode, -doUnobserved 1/2/4/5, GL model, missing
genotypes, --rm-empty-sites, gaps between sites) go through the instrumented binary (per-site capture before merging)
and the unmodified one (-O u, merged output); the oracle's merge of the capture must give exactly the records written.

Container only: skipped where oracle/_ref does not exist (the GPU box uses tests/golden/gvcf/ instead)."""
import os
import random
import subprocess

import pytest

import gvcf_util as gu
from fuzz_cases import reference_exited
import vgl_dump
from test_gvcf_oracle import check_blocks
from vcfgl_b200 import args as vargs
from vcfgl_b200 import synth

BIN = os.path.join(gu.ROOT, "oracle", "_ref", "vcfgl_ref")
BIN_DUMP = os.path.join(gu.ROOT, "oracle", "_ref", "vcfgl_ref_dump")

pytestmark = pytest.mark.skipif(not os.path.exists(BIN_DUMP), reason="oracle/_ref not built (needs /root/reference)")


def draw(rnd):
    S = rnd.choice([1, 2, 4, 7, 20])
    n_sites = rnd.choice([20, 40, 90])
    length = n_sites * rnd.choice([2, 5, 12])
    dps = sorted(rnd.sample([1, 2, 3, 4, 5, 8, 10, 15, 30], rnd.choice([1, 2, 3, 4])))
    argv = ["--seed", str(rnd.randrange(1, 10000)), "-O", "v", "-explode", str(rnd.choice([0, 1, 1])),
            "-d", rnd.choice(["0.6", "1.5", "3", "6", "12", "20"]), "-e", rnd.choice(["0", "0.001", "0.01", "0.05"]),
            "-GL", str(rnd.choice([1, 2])), "-doUnobserved", str(rnd.choice([1, 2, 4, 5])), "-doGVCF", "1",
            "--gvcf-dps", ",".join(map(str, dps)), "--rm-empty-sites", str(rnd.choice([0, 1])), "-addPL", "1"]
    for t in ("-addQS", "-addInfoDP", "-addFormatAD", "-addI16", "-addGL"):
        if rnd.random() < 0.5:
            argv += [t, "1"]
    return S, n_sites, length, rnd.choice([0.0, 0.0, 0.15]), argv


@pytest.mark.parametrize("block", range(3))
def test_gvcf_oracle_equals_live_reference(block, tmp_path):
    rnd = random.Random(8200 + block)
    n_blocks = n_records = 0
    for k in range(8):
        S, n_sites, length, miss, argv = draw(rnd)
        a = vargs.parse_args(argv)
        vcf = str(tmp_path / ("in%d.vcf" % k))
        synth.write_vcf(vcf, synth.sfs_genotypes(n_sites, S, 900 + k, miss), synth.positions(n_sites, length, 900 + k), length)
        dump = str(tmp_path / ("c%d.vgld" % k))
        u_argv = [("u" if i and argv[i - 1] == "-O" else x) for i, x in enumerate(argv)]
        exited = False
        for binary, av, env in ((BIN_DUMP, argv, dict(os.environ, VGL_DUMP_PATH=dump)), (BIN, u_argv, dict(os.environ))):
            r = subprocess.run([binary, "-i", vcf, "-o", str(tmp_path / ("o%d" % k))] + av, capture_output=True, text=True, env=env)
            exited = exited or (r.returncode != 0 and reference_exited(r.stderr))
            assert r.returncode == 0 or exited, (av, r.stderr[-1500:])
        if exited:
            continue        # the reference's own run-time exits (fuzz_cases.REFERENCE_EXITS)
        kept = [d for d in vgl_dump.read_dump(dump) if d.ret == 0] if os.path.exists(dump) and os.path.getsize(dump) else []
        bcf = gu.bo.read_bcf(str(tmp_path / ("o%d.bcf" % k)))
        nb = check_blocks(a, kept, bcf, where=argv)
        n_blocks += nb
        n_records += len(bcf[2])
    assert n_blocks > 10 and n_records > 50, (n_blocks, n_records)


@pytest.mark.parametrize("block", range(2))
def test_gvcf_oracle_on_multi_contig_inputs(block, tmp_path):
    """the same on the random texts of tests/test_vcfin_live_reference.py: one to three contigs (a block never spans two), ACGT or
    binary alleles, missing genotypes, gaps"""
    from test_vcfin_live_reference import random_vcf
    rnd = random.Random(8400 + 2 * block)
    n_blocks = n_records = 0
    for k in range(10):
        acgt = rnd.random() < 0.5
        S, buf = random_vcf(rnd, acgt)
        _, _, _, _, argv = draw(rnd)
        argv += ["--source", str(int(acgt))]
        a = vargs.parse_args(argv)
        vcf = str(tmp_path / ("in%d.vcf" % k))
        open(vcf, "wb").write(buf)
        dump = str(tmp_path / ("c%d.vgld" % k))
        u_argv = [("u" if i and argv[i - 1] == "-O" else x) for i, x in enumerate(argv)]
        exited = False
        for binary, av, env in ((BIN_DUMP, argv, dict(os.environ, VGL_DUMP_PATH=dump)), (BIN, u_argv, dict(os.environ))):
            r = subprocess.run([binary, "-i", vcf, "-o", str(tmp_path / ("o%d" % k))] + av, capture_output=True, text=True, env=env)
            exited = exited or (r.returncode != 0 and reference_exited(r.stderr))
            assert r.returncode == 0 or exited, (av, r.stderr[-1500:])
        if exited:
            continue
        kept = [d for d in vgl_dump.read_dump(dump) if d.ret == 0] if os.path.exists(dump) and os.path.getsize(dump) else []
        bcf = gu.bo.read_bcf(str(tmp_path / ("o%d.bcf" % k)))
        n_blocks += check_blocks(a, kept, bcf, where=(argv, buf.decode()))
        n_records += len(bcf[2])
    assert n_blocks > 5 and n_records > 50, (n_blocks, n_records)
