"""BCF input on the device (k_bcf_gt behind vgl_parse_bcf) against oracle/bcf_in_oracle.py (pinned through the reference-validated
BCF fixtures and the text oracle, tests/test_bcfin_oracle.py).  Exact.
(1) the fixtures, both --source modes, --rm-invar-sites; (2) msprime-shaped records of 1 .. 1000 samples with int8 / int16 /
int32 genotype vectors; (3) defects: haploid / triploid vectors, vector_end, allele index beyond the alleles, no GT, wrong
sample count; (4) BCF in -> place -> simulate equals the packed-genotype entry."""
import numpy as np
import pytest

import bcf_writer as bw
import bcfin_util as bu
import vcfin_oracle as vo
from vcfgl_b200 import args as vargs
from vcfgl_b200 import capi, synth, vcfinput

pytestmark = pytest.mark.gpu
ARGV = "--seed 42 -d 4 -e 0.01 -GL 1 -addPL 1 -addFormatAD 1"


def check(body, off, S, source, gt_key, rm=0, what=""):
    want = bu.oracle(body, off, S, source, gt_key, rm)
    a = vargs.parse_args(ARGV.split())
    a.rm_invar_sites = rm
    ctx = capi.Context(capi.params_from_args(a, S, max_batch_sites=64, n_slots=1))
    ps = ctx.parser(max(len(body), 64), max(len(off) - 1, 1))
    res = ps.parse_bcf(body, off, gt_key, source)
    assert res.n_records == len(want)
    rows = ps.rows(0, res.n_records)
    for i, w in enumerate(want):
        s = res.sites[i]
        assert s["status"] == w["status"], (what, i, s["status"], w["status"])
        assert (s["pos"], s["n_allele"], s["line_off"], s["line_len"]) == (w["pos"], w["n_allele"], off[i], off[i + 1] - off[i])
        assert s["allele_acgt"].tolist() == w["allele_acgt"]
        if w["status"] == 0:
            assert s["skip_code"] == w["skip_code"] and s["allele_sum"] == (w["allele_sum"] if rm & 3 else 0)
            assert np.array_equal(rows[i], w["row"]), (what, i)
    bad = [i for i, w in enumerate(want) if w["status"] != 0]
    assert res.n_errors == len(bad) and res.first_error_record == (bad[0] if bad else -1)
    assert res.n_kept == sum(w["status"] == 0 and w["skip_code"] == 0 for w in want)
    ps.close()
    ctx.close()
    return want


@pytest.mark.parametrize("name", sorted(bu.MANIFEST))
def test_fixtures(name):
    body, off, m = bu.load(name)
    S = len(vcfinput.read_header(vo.load_input(m["vcf"])).samples)
    for source in (0, 1):
        for rm in (0, 3):
            check(body, off, S, source, m["gt_key"], rm, (name, source, rm))


@pytest.mark.parametrize("S,n_sites,width", [(1, 300, 1), (3, 200, 2), (100, 400, 1), (100, 100, 4), (1000, 60, 1), (129, 80, 2)])
def test_msprime_shaped(S, n_sites, width):
    hap = synth.sfs_genotypes(n_sites, S, 50 + S, missing_rate=0.02)
    pos = synth.positions(n_sites, n_sites * 10, 3)
    buf = synth.vcf_header(S, n_sites * 10) + synth.vcf_body(hap, pos)
    bcf, first, offs, ids = bw.vcf_to_bcf(buf, gt_width=width)
    want = check(bcf[first:], np.array(offs, np.uint32), S, 0, ids["GT"], 0, (S, width))
    assert all(w["status"] == 0 for w in want)
    assert np.array_equal(np.stack([w["row"] for w in want]), synth.pack_gt(hap))
    assert [w["pos"] for w in want] == (pos - 1).tolist()


HDR = ("##fileformat=VCFv4.2\n##FILTER=<ID=PASS,Description=\"x\">\n##contig=<ID=1,length=100>\n"
       "##FORMAT=<ID=DP,Number=1,Type=Integer,Description=\"d\">\n##FORMAT=<ID=GT,Number=1,Type=String,Description=\"g\">\n"
       "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\ts0\ts1\n")
DEFECTS = [
    ("1\t5\t.\t0\t1\t.\tPASS\t.\tGT\t0|1\t1", capi.IN_EPLOIDY),            # second sample haploid: vector_end
    ("1\t5\t.\t0\t1\t.\tPASS\t.\tGT\t0\t1", capi.IN_EPLOIDY),              # ploidy 1
    ("1\t5\t.\t0\t1\t.\tPASS\t.\tGT\t0|1|1\t0|0|0", capi.IN_EPLOIDY),      # ploidy 3
    ("1\t5\t.\t0\t.\t.\tPASS\t.\tGT\t0|1\t0|0", capi.IN_EALLELEIDX),
    ("1\t5\t.\t0\t1\t.\tPASS\t.\tDP\t3\t4", capi.IN_ENOGT),
    ("1\t5\t.\t0\t2\t.\tPASS\t.\tGT\t0|1\t0|0", capi.IN_EALLELE),
    ("1\t5\t.\t0\t1,1\t.\tPASS\t.\tGT\t0|1\t0|0", capi.IN_ENALLELE),
    ("1\t5\tid7\t0\t1\t.\tPASS\t.\tDP:GT\t3:0|1\t.:.|.", capi.IN_OK),
]


def test_defects():
    buf = (HDR + "".join(l + "\n" for l, _ in DEFECTS)).encode()
    bcf, first, offs, ids = bw.vcf_to_bcf(buf)
    want = check(bcf[first:], np.array(offs, np.uint32), 2, 0, ids["GT"], 0, "defects")
    assert [w["status"] for w in want] == [c for _, c in DEFECTS]
    # a header with three samples against records of two: ENSAMPLES
    want = check(bcf[first:], np.array(offs, np.uint32), 3, 0, ids["GT"], 0, "nsamples")
    assert all(w["status"] <= capi.IN_ENSAMPLES and w["status"] != 0 for w in want)


def test_bcf_to_tags_equals_packed_submit():
    S, n = 40, 60
    body, off, m = bu.load("s40.in.bcf")
    a = vargs.parse_args(ARGV.split())
    ctx = capi.Context(capi.params_from_args(a, S, max_batch_sites=n, n_slots=2))
    ps = ctx.parser(len(body), n)
    res = ps.parse_bcf(body, off, m["gt_key"], 0)
    assert res.n_records == n and res.n_errors == 0
    ctx.place_rows(0, ps, n)
    ctx.submit(0, 7, n, flags=capi.SUBMIT_GT_ON_DEVICE)
    b0 = ctx.wait(0)
    got = [{k: (np.array(v, copy=True) if isinstance(v, np.ndarray) else v) for k, v in b0.site(i).items()} for i in range(n)]
    ctx.input_buffer(1)[:n] = ps.rows(0, n)
    ctx.submit(1, 7, n)
    b1 = ctx.wait(1)
    for i in range(n):
        w = b1.site(i)
        for k, v in got[i].items():
            if isinstance(v, np.ndarray):
                assert np.array_equal(v.view(np.uint8), np.ascontiguousarray(w[k]).view(np.uint8)), (i, k)
            else:
                assert v == w[k], (i, k)
    ps.close()
    ctx.close()


@pytest.mark.parametrize("name,source,explode", [("data7.bcf", 1, 1), ("in_acgt.bcf", 1, 1), ("in_binary.bcf", 0, 0), ("s8.in.bcf", 0, 1)])
def test_bcf_driver_equals_text_driver(name, source, explode):
    """the whole driver loop (contigs from rid, -explode, batches of 5 sites) from BCF records gives the sites and tags of the
    same loop from the VCF text"""
    body, off, m = bu.load(name)
    buf = vo.load_input(m["vcf"])
    hdr = vcfinput.read_header(buf)
    S = len(hdr.samples)
    a = vargs.parse_args(ARGV.split())

    def collect(gen):
        out = []
        for run, b in gen:
            for i in range(b.n_sites):
                o = b.site(i)
                out.append((run.contig, int(run.pos[i]), {k: (np.array(v, copy=True) if isinstance(v, np.ndarray) else v) for k, v in o.items()}))
        return out
    ctx = capi.Context(capi.params_from_args(a, S, max_batch_sites=5, n_slots=2))
    ps = ctx.parser(max(len(body), len(buf)) + 64, 16)
    text = collect(vcfinput.simulate_vcf_text(ctx, ps, buf[hdr.body_offset:], gt_source=source, explode=explode, contigs=hdr.contigs))
    bcf = collect(vcfinput.simulate_bcf_records(ctx, ps, body, gt_key=m["gt_key"], gt_source=source, explode=explode,
                                                contig_names=list(hdr.contigs), contig_lengths=hdr.contigs, max_records_per_chunk=3))
    assert len(text) == len(bcf) > 0
    for (c1, p1, o1), (c2, p2, o2) in zip(text, bcf):
        assert (c1, p1) == (c2, p2)
        for k, v in o1.items():
            if isinstance(v, np.ndarray):
                assert np.array_equal(v.view(np.uint8), o2[k].view(np.uint8)), (p1, k)
            else:
                assert v == o2[k], (p1, k)
    ps.close()
    ctx.close()


@pytest.mark.parametrize("name,source,explode,dps", [("data7.bcf", 1, 1, None), ("in_binary.bcf", 0, 0, None), ("s8.in.bcf", 0, 1, "1,2")])
def test_cpp_host_driver_bcf(name, source, explode, dps, tmp_path):
    """vgl::VcfTextSimulator::run_bcf (C++: BCF header, record hops, vgl_parse_bcf, site plan) prints exactly what ::run prints
    for the VCF the BCF was made from"""
    import gzip
    import os
    import subprocess
    exe = os.path.join(bu.ROOT, "vcfgl_b200", "host", "example_driver")
    if not os.path.exists(exe):
        pytest.skip("example_driver not built")
    m = bu.MANIFEST[name]
    pb, pv = tmp_path / "in.bcf", tmp_path / "in.vcf"
    pb.write_bytes(gzip.open(os.path.join(bu.INPUTS, name + ".gz"), "rb").read())
    pv.write_bytes(vo.load_input(m["vcf"]))
    outs = []
    for path, is_bcf in ((pv, False), (pb, True)):
        env = dict(os.environ, VGL_VCF_IN=str(path), VGL_SOURCE=str(source), VGL_EXPLODE=str(explode), VGL_BATCH="5")
        if dps:
            env["VGL_GVCF_DPS"] = dps
        if is_bcf:
            env["VGL_INPUT_IS_BCF"] = "1"
        r = subprocess.run([exe], capture_output=True, text=True, env=env, timeout=120)
        assert r.returncode == 0, r.stderr
        outs.append(r.stdout)
    assert outs[0] == outs[1] and len(outs[0].splitlines()) >= m["n_records"]
