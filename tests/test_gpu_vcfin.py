"""Input path on the device (k_vcf_lines, k_vcf_gt, k_place_rows; include/vgl.h "Input path") against the CPU oracle
(oracle/vcf_in_oracle.c, pinned on the reference's captures by tests/test_vcfin_oracle.py).

Integer / byte work: every comparison is exact.
(1) the reference's own test inputs, the hand-written inputs and the single-record cases: site records and genotype rows;
(2) msprime-shaped text of 1 .. 10000 samples (the BASELINE.json configs' shapes) and a fuzz of ragged records
    (GT anywhere in FORMAT, multi-digit alleles, unphased / missing genotypes, CRLF, every defect the status codes name);
(3) chunking: max_records and partial last lines -> bytes_consumed, VGL_PARSE_FINAL;
(4) text -> parse -> place -> simulate equals submitting the same genotypes as packed bytes (also with -explode 1 and
    --rm-invar-sites), i.e. the input path is a drop-in for the host-side packing.
"""
import numpy as np
import pytest

import golden_cases as gc
import vcfin_oracle as vo
from test_vcfin_oracle import planned_sequence
from vcfin_lines import BAD, GOOD
from vcfgl_b200 import args as vargs
from vcfgl_b200 import capi, synth, vcfinput

pytestmark = pytest.mark.gpu

ARGV = "--seed 42 -d 4 -e 0.01 -GL 1 -addPL 1 -addFormatAD 1"


def make_ctx(S, max_sites=64, rm_invar=0, argv=ARGV, n_slots=2, **kw):
    a = vargs.parse_args(argv.split())
    a.rm_invar_sites = rm_invar
    return capi.Context(capi.params_from_args(a, S, max_batch_sites=max_sites, n_slots=n_slots, **kw))


def compare(res: capi.ParseResult, rows_dev, sites, rows, what="", rm_invar=0):
    assert res.n_records == len(sites), (what, res.n_records, len(sites))
    d = res.sites
    assert np.array_equal(d["status"], sites["status"]), (what, d["status"], sites["status"])
    assert np.array_equal(d["skip_code"], sites["skip_code"]), what
    assert np.array_equal(d["line_off"], sites["line_off"]) and np.array_equal(d["line_len"], sites["line_len"]), what
    cols = sites["status"] != capi.IN_ENCOLS
    for k in ("pos", "n_allele", "allele_acgt", "id_off", "fmt_off", "samples_off"):
        assert np.array_equal(d[k][cols], sites[k][cols]), (what, k, d[k][cols], sites[k][cols])
    ok = sites["status"] == 0
    want_sum = sites["allele_sum"][ok] if rm_invar & 3 else np.zeros(int(ok.sum()), np.int64)   # vgl.h: only with --rm-invar-sites 1 | 2
    assert np.array_equal(d["allele_sum"][ok], want_sum), what
    assert np.array_equal(rows_dev[ok], rows[ok]), what
    assert res.n_errors == int((~ok).sum())
    assert res.first_error_record == (int(np.argmax(~ok)) if (~ok).any() else -1)
    assert res.n_kept == int((ok & (sites["skip_code"] == 0)).sum())


def check_body(body: bytes, S, source, rm_invar=0, what=""):
    sites, rows, used = vo.parse(body, S, source, rm_invar)
    ctx = make_ctx(S, rm_invar=rm_invar)
    ps = ctx.parser(max(len(body), 64), max(len(sites), 1))
    res = ps.parse(body, source, capi.PARSE_FINAL)
    compare(res, ps.rows(0, res.n_records), sites, rows, what, rm_invar)
    assert res.bytes_consumed == len(body)
    ps.close()
    ctx.close()
    return sites


@pytest.mark.parametrize("name", sorted({m["input"] for m in gc.MANIFEST.values()}) + ["in_acgt.vcf", "in_binary.vcf"])
def test_reference_inputs(name):
    buf = vo.load_input(name)
    hdr = vcfinput.read_header(buf)
    for source in (0, 1):
        for rm in (0, 3):
            check_body(buf[hdr.body_offset:], len(hdr.samples), source, rm, (name, source, rm))


def test_single_records():
    for line, S, source, *_ in BAD + GOOD:
        for tail in (b"\n", b""):
            s = check_body(line + tail, S, source, 0, line)
            assert len(s) == (1 if (line or tail) else 0)
    # all of them in one chunk (the statuses must not leak between records)
    for source in (0, 1):
        body = b"".join(l + b"\n" for l, S, *_ in BAD + GOOD if S == 2)
        check_body(body, 2, source, 3, "all")


@pytest.mark.parametrize("S,n_sites", [(1, 5000), (2, 3000), (3, 777), (15, 999), (16, 999), (100, 4000), (127, 500), (1000, 300),
                                       (10000, 40)])
def test_msprime_shaped(S, n_sites):
    hap = synth.sfs_genotypes(n_sites, S, 1000 + S, missing_rate=0.01)
    pos = synth.positions(n_sites, n_sites * 20, 5)
    import io
    import os
    import tempfile
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, "x.vcf")
        synth.write_vcf(path, hap, pos, n_sites * 20)
        buf = open(path, "rb").read()
    hdr = vcfinput.read_header(buf)
    body = buf[hdr.body_offset:]
    sites = check_body(body, S, 0, 0, S)
    assert (sites["status"] == 0).all() and np.array_equal(sites["pos"], pos - 1)
    # and against the generator's own packing (binary source: 0 -> A, 1 -> C)
    ctx = make_ctx(S)
    ps = ctx.parser(len(body), n_sites)
    res = ps.parse(body, 0, 0)
    assert np.array_equal(ps.rows(0, n_sites), synth.pack_gt(hap))
    ps.close()
    ctx.close()


def fuzz_body(rng, n_lines, S, source):
    alleles = [b"A", b"C", b"G", b"T", b"<*>", b"<NON_REF>", b"N", b"AC", b"a", b""] if source else [b"0", b"1", b"2", b"01", b"1x", b""]
    keys = [b"GT", b"DP", b"GQ", b"PL", b"AD"]
    lines = []
    for _ in range(n_lines):
        clean = rng.random() < 0.6          # a record without defects (whatever its shape)
        n_alt = int(rng.choice([0, 1, 1, 1, 2, 3, 4] if clean else [0, 1, 1, 1, 2, 3, 4, 5]))
        good = clean or rng.random() < 0.8
        pool = alleles[:4] if (good and source) else alleles[:2] if good else alleles
        ref = pool[rng.integers(len(pool))]
        alt = b",".join(pool[rng.integers(len(pool))] for _ in range(n_alt)) if n_alt else b"."
        if not source and good and n_alt > 1:
            alt = alt.split(b",")[0]
        if clean and rng.random() < 0.3:      # the tskit shape: FORMAT "GT", every column "a|b"
            line = b"\t".join([b"1", str(int(rng.integers(1, 10 ** 9))).encode(), b".", ref, alt, b".", b"PASS", b".", b"GT"] +
                              [rng.choice([b"0", b"1", b"."], p=[0.6, 0.3, 0.1]) + rng.choice([b"|", b"/"]) +
                               (b"1" if alt != b"." and rng.random() < 0.4 else b"0") for _ in range(S)])
            lines.append(line + (b"\r" if rng.random() < 0.1 else b""))
            continue
        n_al = 1 + (len(alt.split(b",")) if alt != b"." else 0)
        fmt = [keys[i] for i in rng.permutation(len(keys))[:rng.integers(1, 4)]]
        if (clean or rng.random() < 0.9) and b"GT" not in fmt:
            fmt[rng.integers(len(fmt))] = b"GT"
        if rng.random() < 0.7 and b"GT" in fmt:
            fmt.remove(b"GT")
            fmt.insert(0, b"GT")
        cols = []
        n_cols = S if (clean or rng.random() < 0.95) else int(rng.integers(0, S + 3))
        for _s in range(n_cols):
            sub = []
            for k in fmt:
                if k != b"GT":
                    sub.append(rng.choice([b"3", b"17", b".", b"1,2,3", b""]))
                    continue
                r = rng.random() * (0.8 if clean else 1.0)
                if r < 0.8:
                    a = [str(int(rng.integers(0, max(n_al, 1)))).encode() if rng.random() < 0.93 else b"." for _ in range(2)]
                    sub.append(a[0] + rng.choice([b"|", b"/"]) + a[1])
                elif r < 0.85:
                    sub.append(rng.choice([b"0", b".", b"0|0|0", b"1/", b"|1", b"x", b"0|x", b"0|1x", b"", b"+|0", b"+1|+0"]))
                elif r < 0.9:
                    sub.append(str(int(rng.integers(0, 12))).encode() + b"|" + b"0" * int(rng.integers(1, 3)) + str(int(rng.integers(0, 3))).encode())
                else:
                    sub.append(str(int(rng.integers(0, n_al + 1))).encode() + b"/" + str(int(rng.integers(0, n_al + 1))).encode())
            if not clean and rng.random() < 0.03:
                sub = sub[:rng.integers(0, len(sub) + 1)]
            cols.append(b":".join(sub))
        pos = rng.choice([str(int(rng.integers(0, 10 ** 6))).encode(), b"+12", b"0", b"12x", b"99999999999", b"18446744073709551616000"],
                         p=[0.94, 0.02, 0.02, 0.02, 0, 0] if clean else [0.9, 0.02, 0.02, 0.02, 0.02, 0.02])
        fixed = [rng.choice([b"1", b"chr2", b"c"]), pos, rng.choice([b".", b"rs7", b"a;b"]), ref, alt,
                 rng.choice([b".", b"30", b"1e3"]), rng.choice([b".", b"PASS", b"q10;s50"]),
                 rng.choice([b".", b"NS=3;DP=14", b"X" * int(rng.integers(1, 700))]), b":".join(fmt)]
        if not clean and rng.random() < 0.03:
            fixed = fixed[:rng.integers(0, 9)]
            cols = []
        line = b"\t".join(fixed + cols)
        if rng.random() < 0.1:
            line += b"\r"
        lines.append(line)
    return b"\n".join(lines) + b"\n"


@pytest.mark.parametrize("S,source,seed", [(1, 0, 1), (2, 1, 2), (5, 0, 3), (5, 1, 4), (33, 1, 5), (130, 0, 6), (130, 1, 7), (600, 1, 8)])
def test_fuzz(S, source, seed):
    rng = np.random.default_rng(seed)
    body = fuzz_body(rng, 1200 if S < 200 else 200, S, source)
    sites = check_body(body, S, source, int(rng.integers(0, 4)), (S, source, seed))
    assert (sites["status"] == 0).sum() > len(sites) // 10      # the fuzz does reach the genotype rows
    assert len(set(sites["status"].tolist())) >= 6              # ... and most defects


def test_chunking_and_final():
    S = 7
    rng = np.random.default_rng(11)
    lines = [b"1\t%d\t.\t0\t1\t.\tPASS\t.\tGT\t" % (i + 1) + b"\t".join(rng.choice([b"0|0", b"0|1", b"1|1", b".|."], S)) for i in range(50)]
    body = b"\n".join(lines)          # no trailing LF
    sites, rows, _ = vo.parse(body, S, 0)
    ctx = make_ctx(S)
    ps = ctx.parser(len(body) + 8, 16)
    # without FINAL the unterminated last line is left over; max_records = 16 per call
    off, got_rows, n_calls = 0, [], 0
    while off < len(body):
        res = ps.parse(body[off:], 0, 0)
        if res.n_records == 0:
            res = ps.parse(body[off:], 0, capi.PARSE_FINAL)
            assert res.n_records == 1 and res.bytes_consumed == len(body) - off
        assert res.n_records <= 16 and res.n_errors == 0
        got_rows.append(ps.rows(0, res.n_records))
        off += res.bytes_consumed
        n_calls += 1
    assert n_calls == 5 and np.array_equal(np.concatenate(got_rows), rows)
    # text reuse on the device
    res = ps.parse(body, 0, capi.PARSE_FINAL)
    again = ps.parse(None, 0, capi.PARSE_TEXT_ON_DEVICE, n_bytes=len(body))
    assert again.n_records == res.n_records == 16 and np.array_equal(ps.rows(0, 16), rows[:16])
    ps.close()
    ctx.close()


def run_packed(a, S, gts_rows, first_site_id=0):
    """the same sites submitted as packed host genotypes (the established path)"""
    n = len(gts_rows)
    ctx = capi.Context(capi.params_from_args(a, S, max_batch_sites=n, n_slots=1))
    ctx.input_buffer(0)[:n] = gts_rows
    ctx.submit(0, first_site_id, n)
    b = ctx.wait(0)
    out = [b.site(i) for i in range(n)]
    out = [{k: (np.array(v, copy=True) if isinstance(v, np.ndarray) else v) for k, v in o.items()} for o in out]
    ctx.close()
    return out


def same_site(x, y):
    for k in x:
        if isinstance(x[k], np.ndarray):
            if not np.array_equal(np.ascontiguousarray(x[k]).view(np.uint8), np.ascontiguousarray(y[k]).view(np.uint8)):
                return k
        elif x[k] != y[k]:
            return k
    return None


@pytest.mark.parametrize("cid", sorted(vo.in_cases()))
def test_text_to_tags_equals_packed_submit(cid):
    c = vo.in_cases()[cid]
    buf = vo.load_input(c["input"])
    hdr = vcfinput.read_header(buf)
    S = len(hdr.samples)
    body = buf[hdr.body_offset:]
    a = vargs.parse_args(ARGV.split())
    a.rm_invar_sites = c["rm_invar_sites"]
    # expected site sequence from the reference capture -> packed rows
    want = np.array([[(g[2 * s] & 0xF) | ((g[2 * s + 1] & 0xF) << 4) for s in range(S)] for _, g in c["sites"]], np.uint8)
    ref = run_packed(a, S, want, first_site_id=100)
    for max_sites, chunk in ((5, None), (64, 150)):
        ctx = capi.Context(capi.params_from_args(a, S, max_batch_sites=max_sites, n_slots=2))
        ps = ctx.parser(len(body) + 1, max(max_sites, 16))
        got, got_pos = [], []
        for run, b in vcfinput.simulate_vcf_text(ctx, ps, body, gt_source=c["source"], explode=c["explode"], contigs=hdr.contigs,
                                                 chunk_bytes=chunk, first_site_id=100):
            assert b.n_sites == len(run.pos)
            for i in range(b.n_sites):
                o = b.site(i)
                got.append({k: (np.array(v, copy=True) if isinstance(v, np.ndarray) else v) for k, v in o.items()})
            got_pos += run.pos.tolist()
        assert got_pos == [p for p, _ in c["sites"]]
        assert len(got) == len(ref)
        for i, (x, y) in enumerate(zip(got, ref)):
            assert same_site(x, y) is None, (cid, i, same_site(x, y))
        ps.close()
        ctx.close()


def test_text_to_tags_large_explode():
    """cfg4 shape in small: sparse records on a contig, -explode 1 fills the positions between them"""
    S, n_rec, L = 100, 300, 9000
    hap = synth.sfs_genotypes(n_rec, S, 99)
    pos = synth.positions(n_rec, L, 99)
    import os
    import tempfile
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, "x.vcf")
        synth.write_vcf(path, hap, pos, L)
        buf = open(path, "rb").read()
    hdr = vcfinput.read_header(buf)
    body = buf[hdr.body_offset:]
    a = vargs.parse_args("--seed 42 -d 3 -e 0.001 -GL 1 -doUnobserved 1 -addPL 1 -addI16 1 -addQS 1".split())
    want = np.zeros((L, S), np.uint8)
    want[pos - 1] = synth.pack_gt(hap)
    ref = run_packed(a, S, want)
    ctx = capi.Context(capi.params_from_args(a, S, max_batch_sites=2048, n_slots=2))
    ps = ctx.parser(1 << 16, 2048)
    k = 0
    for run, b in vcfinput.simulate_vcf_text(ctx, ps, body, gt_source=0, explode=1, contigs=hdr.contigs):
        assert np.array_equal(run.pos, np.arange(k, k + b.n_sites))
        for i in range(0, b.n_sites, 7):
            assert same_site(b.site(i), ref[k + i]) is None, (k + i, same_site(b.site(i), ref[k + i]))
        k += b.n_sites
    assert k == L
    ps.close()
    ctx.close()


@pytest.mark.parametrize("cid", ["in_acgt_explode", "in_acgt_rm3", "in_binary_plain", "in_binary_rm3_explode"])
def test_cpp_host_driver(cid, tmp_path):
    """vcfgl_b200/host/vgl_host.hpp VcfTextSimulator (the C++ mirror of the reference's driver loop) reads the VCF file itself:
    same site sequence as the reference capture, same DP / AD as the Python host path on the same parameters"""
    import gzip
    import os
    import subprocess
    exe = os.path.join(vo.ROOT, "vcfgl_b200", "host", "example_driver")
    if not os.path.exists(exe):
        pytest.skip("example_driver not built")
    c = vo.in_cases()[cid]
    buf = vo.load_input(c["input"])
    path = tmp_path / "in.vcf"
    path.write_bytes(buf)
    env = dict(os.environ, VGL_VCF_IN=str(path), VGL_SOURCE=str(c["source"]), VGL_EXPLODE=str(c["explode"]),
               VGL_RM_INVAR=str(c["rm_invar_sites"]), VGL_BATCH="4")
    r = subprocess.run([exe], capture_output=True, text=True, env=env, timeout=120)
    assert r.returncode == 0, r.stderr
    lines = [l.split("\t") for l in r.stdout.splitlines()]
    assert [int(l[1]) - 1 for l in lines] == [p for p, _ in c["sites"]]
    # the same sites through the Python host path with the driver's parameters
    hdr = vcfinput.read_header(buf)
    S = len(hdr.samples)
    a = vargs.parse_args("--seed 42 -d 4 -e 0.01 -GL 1 -doUnobserved 1 -addGL 1 -addPL 1 -addFormatAD 1 -addInfoDP 1".split())
    a.rm_invar_sites = c["rm_invar_sites"]
    ctx = capi.Context(capi.params_from_args(a, S, max_batch_sites=4, n_slots=2))
    ps = ctx.parser(1 << 16, 16)
    k = 0
    for run, b in vcfinput.simulate_vcf_text(ctx, ps, buf[hdr.body_offset:], gt_source=c["source"], explode=c["explode"],
                                             contigs=hdr.contigs):
        for i in range(b.n_sites):
            o, l = b.site(i), lines[k]
            k += 1
            if o["skip_code"] != 0:
                assert l[2].startswith("skipped(%d)" % o["skip_code"])
                continue
            assert l[3] == "DP=%d" % o["info_dp"]
            A = o["n_alleles"]
            for s in range(S):
                dp, ad = l[6 + s].split(":")
                assert int(dp) == o["fmt_dp"][s]
                assert [int(x) for x in ad.split(",")] == o["fmt_ad"].reshape(S, A)[s].tolist()
    assert k == len(lines)
    ps.close()
    ctx.close()
