"""Input path on the CPU: the oracle (oracle/vcf_in_oracle.c) and the host-side site planner
(vcfgl_b200/vcfinput.py) against what the reference itself did with the same input files.

For every golden case the instrumented reference dumped, per site that reached simulate_record_values, its position
and true_gts_acgt_int (tests/golden/<id>.vgld.gz); tests/golden/inputs/in_cases.json holds the same for the hand-written
inputs (--rm-invar-sites 1/2/3, -explode with --source 1, unphased / missing genotypes, FORMAT with several keys).
oracle parse -> SitePlanner must give exactly that sequence."""
import numpy as np
import pytest

import golden_cases as gc
import vcfin_oracle as vo
from vcfgl_b200 import capi, vcfinput


def planned_sequence(body_text: bytes, S, gt_source, explode, rm_invar, contigs, max_run=7, chunk_records=None):
    """[(pos, gts int8[2S])] via oracle + planner; records are fed in chunks like the device parser would deliver them"""
    text = np.frombuffer(body_text, np.uint8)
    planner = vcfinput.SitePlanner(explode, rm_invar, contigs, max_run)
    out = []
    off = 0
    last_acgt0 = -1
    while off < len(body_text):
        sites, rows, used = vo.parse(body_text[off:], S, gt_source, rm_invar, final=True,
                                     max_records=chunk_records or 10 ** 6)
        assert (sites["status"] == 0).all(), sites["status"]
        chunk = text[off:off + used]
        last_acgt0 = int(sites["allele_acgt"][-1][0])
        for run in planner.feed(chunk, sites):
            for p, s in zip(run.pos, run.src):
                fill = np.full(2 * S, planner.fill_acgt, np.int8)
                out.append((int(p), vo.unpack_row(rows[s]) if s >= 0 else fill))
        off += used
    for run in planner.finish(last_acgt0):
        for p, s in zip(run.pos, run.src):
            assert s < 0
            out.append((int(p), np.full(2 * S, planner.fill_acgt, np.int8)))
    return out


@pytest.mark.parametrize("cid", gc.CASE_IDS)
def test_golden_case_inputs(cid):
    m = gc.MANIFEST[cid]
    a = gc.case_args(cid)
    sites = gc.case_sites(cid)
    buf = vo.load_input(m["input"])
    hdr = vcfinput.read_header(buf)
    S = len(hdr.samples)
    assert S == sites[0].S
    for chunk_records in (None, 3):
        seq = planned_sequence(buf[hdr.body_offset:], S, a.source, a.explode, a.rm_invar_sites & 3, hdr.contigs,
                               chunk_records=chunk_records)
        assert len(seq) == len(sites), (len(seq), len(sites))
        for (p, g), d in zip(seq, sites):
            assert p == d.pos
            assert np.array_equal(g, d.gts), (p, g, d.gts)


@pytest.mark.parametrize("cid", sorted(vo.in_cases()))
def test_hand_written_inputs(cid):
    c = vo.in_cases()[cid]
    buf = vo.load_input(c["input"])
    hdr = vcfinput.read_header(buf)
    S = len(hdr.samples)
    assert S == c["n_samples"]
    for chunk_records, max_run in ((None, 5), (2, 3), (1, 100)):
        seq = planned_sequence(buf[hdr.body_offset:], S, c["source"], c["explode"], c["rm_invar_sites"] & 3, hdr.contigs,
                               max_run=max_run, chunk_records=chunk_records)
        assert [p for p, _ in seq] == [p for p, _ in c["sites"]]
        for (p, g), (_, want) in zip(seq, c["sites"]):
            assert g.tolist() == want, (p, g.tolist(), want)


from vcfin_lines import BAD, GOOD  # noqa: E402


def test_error_codes():
    for line, S, source, want in BAD:
        sites, rows, used = vo.parse(line + b"\n", S, source)
        assert len(sites) == 1 and sites[0]["status"] == want, (line, sites[0]["status"], want)


def test_good_lines():
    for line, S, source, gts, pos, n_allele in GOOD:
        sites, rows, used = vo.parse(line + b"\n", S, source)
        assert sites[0]["status"] == 0, (line, sites[0]["status"])
        assert rows[0].tolist() == gts, (line, rows[0].tolist())
        assert sites[0]["pos"] == pos and sites[0]["n_allele"] == n_allele


def test_chunking_leaves_partial_lines():
    body = b"".join(l + b"\n" for l, *_ in GOOD[:2])
    sites, rows, used = vo.parse(body[:-3], 2, 0, final=False)
    assert len(sites) == 1 and used == len(GOOD[0][0]) + 1
    sites, rows, used = vo.parse(body[:-3], 2, 0, final=True)
    assert len(sites) == 2 and used == len(body) - 3
    sites, rows, used = vo.parse(body, 2, 0, final=False, max_records=1)
    assert len(sites) == 1 and used == len(GOOD[0][0]) + 1


def test_skip_codes():
    body = (b"1\t1\t.\t0\t1\t.\t.\t.\tGT\t0|0\t0|0\n1\t2\t.\t0\t1\t.\t.\t.\tGT\t1|1\t1|1\n"
            b"1\t3\t.\t0\t1\t.\t.\t.\tGT\t1|1\t.|.\n1\t4\t.\t0\t1\t.\t.\t.\tGT\t.|.\t.|.\n")
    for rm, want in ((0, [0, 0, 0, 0]), (1, [-1, 0, 0, -1]), (2, [0, -2, 0, 0]), (3, [-1, -2, 0, -1])):
        sites, _, _ = vo.parse(body, 2, 0, rm_invar=rm)
        assert sites["skip_code"].tolist() == want


def test_lines_against_reference():
    """tests/golden/inputs/line_cases.json = the unmodified reference run on every BAD / GOOD record (tools/make_line_cases.py).
    A record the oracle rejects must be one the reference (a) exits on, or (b) silently stops reading at (a failed
    bcf_read ends its driver loop, vcfgl.cpp:1479, so no site is simulated), or (c) one of the three documented
    deviations where the reference runs on with undefined meaning: 64-bit positions (EPOS), a triploid genotype read
    through a diploid stride (EPLOIDY), a genotype pointing at <*> (ESYMBOLIC: base index 4 out of range)."""
    import json
    import os
    fx = json.load(open(os.path.join(vo.INPUTS, "line_cases.json")))
    assert len(fx["bad"]) == len(BAD) and len(fx["good"]) == len(GOOD)
    for c, (line, S, source, want) in zip(fx["bad"], BAD):
        assert c["line"].encode("latin1") == line and c["oracle_status"] == want
        if c["ref_ok"] and c["sites"]:
            assert want in (capi.IN_EPOS, capi.IN_EPLOIDY, capi.IN_ESYMBOLIC), (line, want)
    for c, (line, S, source, gts, pos, n_allele) in zip(fx["good"], GOOD):
        assert c["ref_ok"] and len(c["sites"]) == 1
        sites, rows, _ = vo.parse(line + b"\n", S, source)
        assert sites[0]["status"] == 0 and sites[0]["pos"] == c["sites"][0][0]
        assert vo.unpack_row(rows[0]).tolist() == c["sites"][0][1]
