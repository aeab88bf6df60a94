"""Pins oracle/gvcf_oracle.py (restatement of prepare_gvcf_block, bcf_utils.cpp:662-942) on the reference itself: the
instrumented binary's per-site capture of a -doGVCF 1 run goes through the oracle, and the resulting record list must be
the one the unmodified binary wrote (-O u) for the same run -- block boundaries, END, MIN_DP, per-sample DP and PL."""
import numpy as np
import pytest

import gvcf_util as gu


@pytest.mark.parametrize("cid", gu.CASES)
def test_blocks_equal_reference_output(cid):
    a, kept, bcf = gu.load(cid)
    check_blocks(a, kept, bcf, where=cid)


def check_blocks(a, kept, bcf, where=None):
    """the oracle's merge of the captured sites `kept` == the records of the reference's output `bcf`; -> number of blocks"""
    text, ids, recs = bcf
    out = gu.go.merge(gu.oracle_input(kept), gu.dps_of(a))
    assert len(out) == len(recs), (where, len(out), len(recs))
    n_blocks = 0
    for o, rec in zip(out, recs):
        r = gu.decode(rec, ids)
        if o["kind"] == "site":
            d = kept[o["site"]]
            assert (r["rid"], r["pos"]) == (d.rid, d.pos) and "MIN_DP" not in r["info"]
            assert np.array_equal(r["fmt"]["DP"], d.fmt_dp)
            continue
        n_blocks += 1
        f = kept[o["first"]]
        assert (r["rid"], r["pos"]) == (f.rid, o["start"])
        assert r["rlen"] == o["end"] + 1 - o["start"]
        if o["end"] - o["start"] >= 1:
            assert r["info"]["END"].tolist() == [o["end"] + 1]
        else:
            assert "END" not in r["info"]
        assert r["info"]["MIN_DP"].tolist() == [o["min_dp"]]
        assert np.array_equal(r["fmt"]["DP"], o["dp"])
        assert np.array_equal(r["fmt"]["PL"], o["pl"])
        assert len(r["alleles"]) == 2      # REF, <*> / <NON_REF>
    return n_blocks


def test_dp_range():
    assert [gu.go.dp_range(m, [1, 5, 10]) for m in (0, 1, 4, 5, 9, 10, 99)] == [0, 1, 1, 2, 2, 3, 3]


def test_fixtures_do_contain_blocks():
    n = {c: sum(o["kind"] == "block" for o in gu.go.merge(gu.oracle_input(gu.load(c)[1]), gu.dps_of(gu.load(c)[0]))) for c in gu.CASES}
    assert sum(n.values()) > 60 and sum(v > 0 for v in n.values()) >= 6, n
