"""-printPileup from draws (vcfgl_b200/pileup.py) against the reference's own pileup output: the golden test10 pileup
(test/reference/test10/test10.pileup.gz, copied to tests/golden/pileup/) is rebuilt from the capture of the same run."""
import gzip
import os

import numpy as np
import pytest

import golden_cases as gc
import replay_util
import vcfin_oracle as vo
from vcfgl_b200 import pileup, vcfinput


CASES = ["test10", "test1", "x_gl1_eq2_bins_adj", "x_gl2_eq1", "x_missing_gl1", "x_trim_rminvar", "x_acgt_multi", "x_gl2_eq2_precise1"]


@pytest.mark.parametrize("cid", CASES)
def test_pileup_equals_reference(cid):
    a = gc.case_args(cid)
    sites = gc.case_sites(cid)
    S = sites[0].S
    gt, rp = replay_util.batch_from_dump(sites, a)
    buf = vo.load_input(gc.MANIFEST[cid]["input"])
    hdr = vcfinput.read_header(buf)
    names = list(hdr.contigs)
    # REF of every site (binary source: A, vcfgl.cpp:103-127; ACGT source: the record's REF; -explode sites: the REF of the
    # record exploding started at) from the input path's oracle + site planner
    from test_vcfin_oracle import planned_sequence
    recs, _, _ = vo.parse(buf[hdr.body_offset:], S, a.source, a.rm_invar_sites & 3)
    ref_of_pos = {int(r["pos"]): int(r["allele_acgt"][0]) for r in recs}
    first_gap = next((int(r["allele_acgt"][0]) for k, r in enumerate(recs) if int(r["pos"]) != k), int(recs[-1]["allele_acgt"][0]))
    ref = [ref_of_pos.get(d.pos, first_gap) for d in sites]
    got = pileup.format_pileup(a, [names[d.rid] for d in sites], [d.pos for d in sites], ref, [d.ret for d in sites], rp, S)
    want = gzip.open(os.path.join(gc.GOLD, "pileup", cid + ".pileup.gz"), "rb").read()
    assert got == want


def test_fixed_qscores():
    from vcfgl_b200 import args as vargs
    a = vargs.parse_args("--seed 1 -d 1 -e 0.2 -GL 1 --adjust-qs 3 -addQS 1".split())
    assert pileup.fixed_qscores(a) == (6, 7)          # e = 0.2 -> qs 6, +0.499 -> 7
    a = vargs.parse_args("--seed 1 -d 1 -e 0 -GL 2".split())
    assert pileup.fixed_qscores(a) == (63, None)
