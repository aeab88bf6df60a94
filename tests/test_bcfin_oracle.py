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


@pytest.mark.parametrize("cid", ["test1", "test19", "x_acgt_multi", "x_missing_gl1", "x_explode3_e05"])
def test_bcf_site_sequence_equals_reference_capture(cid):
    """BCF records -> oracle -> SitePlanner.feed_rids (contigs from rid, -explode) must give the (pos, true genotypes) sequence the
    instrumented reference dumped for the same input"""
    import golden_cases as gc
    m = gc.MANIFEST[cid]
    a = gc.case_args(cid)
    dumped = gc.case_sites(cid)
    buf = vo.load_input(m["input"])
    hdr = vcfinput.read_header(buf)
    S = len(hdr.samples)
    name = m["input"][:-4] + ".bcf"
    body, off, bm = bu.load(name)
    recs = bu.oracle(body, off, S, a.source, bm["gt_key"], a.rm_invar_sites & 3)
    sites = np.zeros(len(recs), vcfinput.capi.IN_SITE_DTYPE)
    for i, r in enumerate(recs):
        assert r["status"] == 0
        sites[i]["pos"], sites[i]["skip_code"], sites[i]["allele_acgt"] = r["pos"], r["skip_code"], r["allele_acgt"]
    rid = np.array([int.from_bytes(body[int(o) + 8:int(o) + 12], "little", signed=True) for o in off[:-1]], np.int32)
    planner = vcfinput.SitePlanner(a.explode, a.rm_invar_sites & 3, hdr.contigs, 7)
    seq = []
    for run in planner.feed_rids(rid, list(hdr.contigs), sites):
        for p, s in zip(run.pos, run.src):
            seq.append((int(p), vo.unpack_row(recs[s]["row"]) if s >= 0 else np.full(2 * S, planner.fill_acgt, np.int8)))
    for run in planner.finish(int(sites["allele_acgt"][-1][0])):
        for p, s in zip(run.pos, run.src):
            seq.append((int(p), np.full(2 * S, planner.fill_acgt, np.int8)))
    assert len(seq) == len(dumped)
    for (p, g), d in zip(seq, dumped):
        assert p == d.pos and np.array_equal(g, d.gts), (p, g, d.gts)
