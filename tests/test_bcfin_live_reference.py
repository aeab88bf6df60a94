"""Live pin of the BCF-input oracle (oracle/bcf_in_oracle.py): seeded random VCF texts (tests/test_vcfin_live_reference.random_vcf)
are written as uncompressed BCF by the test writer (tests/bcf_writer.py); the unmodified reference must produce the same output
from the .bcf as from the .vcf (which validates the writer against htslib), and the oracle's reading of the BCF records must equal
the text oracle's reading of the VCF lines (pinned on the reference by tests/test_vcfin_live_reference.py).

Container only: skipped where oracle/_ref does not exist (the GPU box uses tests/golden/inputs/*.bcf.gz instead)."""
import os
import random
import subprocess

import numpy as np
import pytest

import bcf_writer as bw
import bcfin_util as bu
import vcfin_oracle as vo
from test_vcfin_live_reference import random_vcf
from vcfgl_b200 import vcfinput

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "oracle", "_ref", "vcfgl_ref")

pytestmark = pytest.mark.skipif(not os.path.exists(BIN), reason="oracle/_ref not built (needs /root/reference)")


def body_lines(path):
    return [l for l in open(path) if not l.startswith("##")]


@pytest.mark.parametrize("block", range(3))
def test_bcf_input_equals_vcf_input(block, tmp_path):
    rnd = random.Random(9900 + block)
    n_records = 0
    for k in range(10):
        acgt = rnd.random() < 0.6
        S, buf = random_vcf(rnd, acgt, extra_values=(".", "3", "17", "99"))      # the test writer stores FORMAT integers as int8
        bcf, first, offs, ids = bw.vcf_to_bcf(buf, gt_width=rnd.choice([1, 1, 2]))
        open(str(tmp_path / ("in%d.vcf" % k)), "wb").write(buf)
        open(str(tmp_path / ("in%d.bcf" % k)), "wb").write(bcf)
        explode, rm = rnd.choice([0, 1]), rnd.choice([0, 1, 2, 3])
        argv = ["--seed", "5", "-O", "v", "--source", str(int(acgt)), "-explode", str(explode), "--rm-invar-sites", str(rm),
                "-d", "2", "-e", "0.01", "-GL", "1", "-addPL", "1"]
        for ext in ("vcf", "bcf"):
            r = subprocess.run([BIN, "-i", str(tmp_path / ("in%d.%s" % (k, ext))), "-o", str(tmp_path / ("o%d_%s" % (k, ext)))] + argv,
                               capture_output=True, text=True)
            assert r.returncode == 0, (ext, argv, buf.decode(), r.stderr[-1500:])
        assert body_lines(str(tmp_path / ("o%d_vcf.vcf" % k))) == body_lines(str(tmp_path / ("o%d_bcf.vcf" % k))), (argv, buf.decode())
        hdr = vcfinput.read_header(buf)
        body = bcf[first:]
        off = bu.record_offsets(body)
        assert off.tolist() == offs
        sites, rows, _ = vo.parse(buf[hdr.body_offset:], S, int(acgt), rm)
        got = bu.oracle(body, off, S, int(acgt), ids["GT"], rm)
        assert len(got) == len(sites)
        for g, s, row in zip(got, sites, rows):
            assert g["status"] == s["status"] == 0
            assert (g["pos"], g["n_allele"], g["skip_code"], g["allele_sum"]) == (s["pos"], s["n_allele"], s["skip_code"], s["allele_sum"])
            assert g["allele_acgt"] == s["allele_acgt"].tolist()
            assert np.array_equal(g["row"], row)
        n_records += len(got)
    assert n_records > 100
