"""Pins oracle/truth_oracle.py (--depth inf, vcfgl.cpp:1089-1262) on the VCFs the unmodified reference wrote
(tests/golden/truth/, tools/make_golden_truth.py; t_test4 = the reference's own golden test4): every record is rebuilt from
the input file's genotypes -- alleles and their order, GL / GP / PL of every sample."""
import numpy as np
import pytest

import truth_util as tu


@pytest.mark.parametrize("cid", tu.CASES)
def test_truth_records_equal_reference_output(cid):
    a, S, seq = tu.case(cid)
    recs = tu.reference_records(cid)
    assert len(recs) == len(seq) > 0
    for (pos, gts), (rpos, alleles, keys, vals) in zip(seq, recs):
        o = tu.to.site(gts, a.do_unobserved)
        assert pos == rpos and o["alleles"] == alleles, (pos, o["alleles"], alleles)
        assert keys == [k for k, on in (("GL", a.add_gl), ("GP", a.add_gp), ("PL", a.add_pl)) if on]
        G = o["n_genotypes"]
        for k in keys:
            assert np.array_equal(vals[k], o[k.lower()].reshape(S, G).astype(np.float64)), (pos, k)


def test_missing_genotype_is_refused_like_the_reference():
    assert tu.REFUSED == ["t_missing_refused"]
    a, S, seq = tu.case("t_missing_refused")
    with pytest.raises(ValueError):
        for pos, gts in seq:
            tu.to.site(gts, a.do_unobserved)
