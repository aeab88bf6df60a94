"""Host logic of the gVCF path on the CPU: vcfgl_b200/gvcf.py GvcfStitcher joins the per-batch records of vgl_gvcf_merge into
the run's record sequence.  Here each batch's device result is stood in for by the oracle's merge of that batch alone
(oracle/gvcf_oracle.py, pinned on the reference); for every fixture and many random batch splits -- including batches of one
site -- the stitched sequence must equal the oracle's merge of the whole run, i.e. the reference's output."""
import random

import numpy as np
import pytest

import gvcf_util as gu
from vcfgl_b200 import gvcf

REC = np.dtype([("first_site", "<i4"), ("last_site", "<i4"), ("n_members", "<i4"), ("min_dp", "<i4"), ("dp_range", "<i4"), ("plane", "<i4")])


def batch_result(sites, dps):
    """what vgl_gvcf_merge returns for one batch, built from the oracle: records + the planes of the block records"""
    out = gu.go.merge(sites, dps)
    recs = np.zeros(len(out), REC)
    dp, pl = [], []
    for k, o in enumerate(out):
        if o["kind"] == "site":
            recs[k] = (o["site"], o["site"], 0, 0, 0, -1)
            continue
        recs[k] = (o["first"], o["last"], len(o["members"]), o["min_dp"], o["range"], len(dp))
        dp.append(o["dp"])
        pl.append(None if o["pl"] is None else o["pl"].reshape(-1, 3))
    has_pl = any(p is not None for p in pl)
    return dict(recs=recs, dp=dp, pl=pl if has_pl else None)


def same(a, b, kept):
    assert a["kind"] == b["kind"]
    if a["kind"] == "site":
        assert a["site"] == b["site"]
        return
    assert (a["first"], a["rid"], a["start"], a["end"], a["min_dp"], a["range"]) == (b["first"], b["rid"], b["start"], b["end"], b["min_dp"], b["range"])
    assert a["n_members"] == len(b["members"])
    assert np.array_equal(a["dp"], b["dp"])
    if b["pl"] is not None:
        assert np.array_equal(np.asarray(a["pl"]).reshape(-1), b["pl"])


@pytest.mark.parametrize("cid", gu.CASES)
def test_stitched_batches_equal_whole_run(cid):
    a, kept, _ = gu.load(cid)
    dps = gu.dps_of(a)
    sites = gu.oracle_input(kept)
    want = gu.go.merge(sites, dps)
    n_blocks = sum(o["kind"] == "block" for o in want)
    rnd = random.Random(len(sites))
    for trial in range(12):
        size = [1, 2, 3, 5, 7, len(sites)][trial % 6] if trial < 6 else None
        st = gvcf.GvcfStitcher()
        got = []
        i = 0
        while i < len(sites):
            n = size or rnd.randrange(1, 9)
            chunk = sites[i:i + n]
            got += list(st.feed(batch_result(chunk, dps), [s["rid"] for s in chunk], [s["pos"] for s in chunk]))
            i += len(chunk)
        got += list(st.finish())
        assert len(got) == len(want), (cid, trial, len(got), len(want))
        for g, w in zip(got, want):
            same(g, w, kept)
    assert n_blocks > 0 or cid.startswith("test")
