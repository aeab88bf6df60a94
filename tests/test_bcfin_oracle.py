"""BCF input on the CPU: oracle/bcf_in_oracle.py on the reference-validated BCF fixtures (tools/make_bcf_inputs.py) must equal
oracle/vcf_in_oracle.c on the VCF they were made from (pinned on the reference's captures): positions, allele maps, skip codes,
allele sums, packed genotypes."""
import numpy as np
import pytest

import bcfin_util as bu
import vcfin_oracle as vo
from vcfgl_b200 import vcfinput


@pytest.mark.parametrize("name", sorted(bu.MANIFEST))
def test_bcf_records_equal_vcf_records(name):
    body, off, m = bu.load(name)
    buf = vo.load_input(m["vcf"])
    hdr = vcfinput.read_header(buf)
    S = len(hdr.samples)
    assert len(off) - 1 == m["n_records"]
    for source in (0, 1):
        for rm in (0, 3):
            sites, rows, _ = vo.parse(buf[hdr.body_offset:], S, source, rm)
            got = bu.oracle(body, off, S, source, m["gt_key"], rm)
            assert len(got) == len(sites)
            for g, s, row in zip(got, sites, rows):
                # text-only defects (columns, POS, GT characters) cannot occur in a BCF record; the rest must agree
                assert g["status"] == s["status"], (name, source, g["status"], s["status"])
                assert (g["pos"], g["n_allele"]) == (s["pos"], s["n_allele"])
                assert g["allele_acgt"] == s["allele_acgt"].tolist()
                if g["status"] == 0:
                    assert g["skip_code"] == s["skip_code"] and g["allele_sum"] == s["allele_sum"]
                    assert np.array_equal(g["row"], row)
