"""gVCF block merger on the device (k_gvcf_key / k_gvcf_plan / k_gvcf_fin / k_gvcf_reduce behind vgl_gvcf_merge) against
oracle/gvcf_oracle.py, which is pinned on the reference's own merged output (tests/test_gvcf_oracle.py).  Integer work: exact.
(1) the reference's -doGVCF runs, replayed on the device from the captures (tags bit-exact), then merged on the device;
(2) native batches of the cfg4 shape (100 samples, 99 % invariant sites, several --gvcf-dps), the oracle fed with the
    batch's own tags; also 1 / 33 / 1000 samples, skipped sites, position gaps, several contigs;
(3) a run cut into batches and stitched on the host equals the run merged in one piece."""
import numpy as np
import pytest

import golden_cases as gc
import gvcf_util as gu
import replay_util
from vcfgl_b200 import args as vargs
from vcfgl_b200 import capi, gvcf, synth

pytestmark = pytest.mark.gpu


def compare(res, want, base=0):
    recs = res["recs"]
    assert len(recs) == len(want), (len(recs), len(want))
    for r, o in zip(recs, want):
        if o["kind"] == "site":
            assert r["n_members"] == 0 and r["first_site"] == r["last_site"] == o["site"] - base and r["plane"] == -1
            continue
        assert (r["first_site"], r["last_site"], r["n_members"]) == (o["first"] - base, o["last"] - base, len(o["members"]))
        assert (r["min_dp"], r["dp_range"]) == (o["min_dp"], o["range"])
        assert np.array_equal(res["dp"][r["plane"]], o["dp"])
        if o["pl"] is not None:
            assert np.array_equal(res["pl"][r["plane"]].reshape(-1), o["pl"])


@pytest.mark.parametrize("cid", gu.CASES)
def test_reference_runs_replayed(cid):
    a, kept, _ = gu.load(cid)
    sites = gc.case_sites(cid) if cid in gu.MAIN_GVCF else gu.vgl_dump.read_dump(gu.os.path.join(gu.GVCF_DIR, cid + ".vgld.gz"))
    S = sites[0].S
    gt, rp = replay_util.batch_from_dump(sites, a)
    ctx = capi.Context(capi.params_from_args(a, S, max_batch_sites=len(sites), n_slots=1))
    ctx.input_buffer(0)[:len(sites)] = gt
    ctx.submit(0, 0, len(sites), replay=rp)
    b = ctx.wait(0)
    assert [b.site(i)["skip_code"] for i in range(len(sites))] == [d.ret for d in sites]
    rid = np.array([d.rid for d in sites], np.int32)
    pos = np.array([d.pos for d in sites], np.int32)
    if a.do_unobserved not in (1, 2):
        with pytest.raises(capi.VglError):
            ctx.gvcf_merge(0, rid, pos, gu.dps_of(a))
        ctx.close()
        return
    res = ctx.gvcf_merge(0, rid, pos, gu.dps_of(a))
    # the oracle works on written sites only; map its indices back to batch indices
    idx = [i for i, d in enumerate(sites) if d.ret == 0]
    want = gu.go.merge(gu.oracle_input(kept), gu.dps_of(a))
    for o in want:
        if o["kind"] == "site":
            o["site"] = idx[o["site"]]
        else:
            o["first"], o["last"] = idx[o["first"]], idx[o["last"]]
    compare(res, want)
    ctx.close()


def native_batch(ctx, a, S, n_sites, seed, invariant=0.9, first_site=0):
    hap = synth.sfs_genotypes(n_sites, S, seed)
    inv = np.random.default_rng(seed + 1).random(n_sites) < invariant
    hap[inv] = 0
    ctx.input_buffer(0)[:n_sites] = synth.pack_gt(hap)
    ctx.submit(0, first_site, n_sites)
    return ctx.wait(0)


def oracle_sites(b, rid, pos):
    out, idx = [], []
    for i in range(b.n_sites):
        o = b.site(i)
        if o["skip_code"] != 0:
            continue
        idx.append(i)
        out.append(dict(rid=int(rid[i]), pos=int(pos[i]), n_alleles_observed=o["n_alleles_observed"], fmt_dp=o["fmt_dp"].copy(),
                        pl=None if o["pl"] is None else np.asarray(o["pl"]).copy()))
    return out, idx


NATIVE = {
    # name: (argv, S, n_sites, dps, invariant share, gap every, contig every)
    "cfg4": ("--seed 42 -d 10 -e 0.001 -GL 1 -doUnobserved 1 -doGVCF 1 --gvcf-dps 1,5,10 -addPL 1 -addI16 1 -addQS 1", 100, 6000, [1, 5, 10], 0.99, 0, 0),
    "lowdepth_gaps": ("--seed 5 -d 2 -e 0.01 -GL 1 -doUnobserved 2 -doGVCF 1 --gvcf-dps 1,2,3 -addPL 1 --rm-empty-sites 1", 3, 5000, [1, 2, 3], 0.9, 37, 1500),
    "s1": ("--seed 6 -d 4 -e 0.01 -GL 1 -doUnobserved 1 -doGVCF 1 --gvcf-dps 2,4 -addPL 1", 1, 4000, [2, 4], 0.8, 0, 0),
    "s33_gl2": ("--seed 7 -d 6 -e 0.01 -GL 2 -doUnobserved 1 -doGVCF 1 --gvcf-dps 1,3 -addPL 1", 33, 3000, [1, 3], 0.95, 211, 0),
    "s300": ("--seed 8 -d 12 -e 0.00001 -GL 1 -doUnobserved 1 -doGVCF 1 --gvcf-dps 1,2,4 -addPL 1", 300, 1200, [1, 2, 4], 0.97, 0, 0),
}


def layout(n_sites, gap_every, contig_every):
    pos = np.arange(n_sites, dtype=np.int32)
    if gap_every:
        pos += (np.arange(n_sites) // gap_every).astype(np.int32) * 3
    rid = np.zeros(n_sites, np.int32) if not contig_every else (np.arange(n_sites) // contig_every).astype(np.int32)
    return rid, pos


@pytest.mark.parametrize("name", sorted(NATIVE))
def test_native_batches(name):
    argv, S, n_sites, dps, inv, gap, ctg = NATIVE[name]
    a = vargs.parse_args(argv.split())
    ctx = capi.Context(capi.params_from_args(a, S, max_batch_sites=n_sites, n_slots=1))
    b = native_batch(ctx, a, S, n_sites, 77, inv)
    rid, pos = layout(n_sites, gap, ctg)
    res = ctx.gvcf_merge(0, rid, pos, dps)
    osites, idx = oracle_sites(b, rid, pos)
    want = gu.go.merge(osites, dps)
    for o in want:
        if o["kind"] == "site":
            o["site"] = idx[o["site"]]
        else:
            o["first"], o["last"] = idx[o["first"]], idx[o["last"]]
    compare(res, want)
    assert res["n_blocks"] > 10 and (res["recs"]["n_members"] > 1).sum() > 5
    ctx.close()


def test_batches_stitched_equal_one_piece():
    argv, S, n_sites, dps, inv, gap, ctg = NATIVE["cfg4"]
    a = vargs.parse_args(argv.split())
    rid, pos = layout(n_sites, 1999, 0)
    ctx = capi.Context(capi.params_from_args(a, S, max_batch_sites=n_sites, n_slots=1))
    b = native_batch(ctx, a, S, n_sites, 78, inv)
    osites, idx = oracle_sites(b, rid, pos)
    want = gu.go.merge(osites, dps)
    hap_gt = ctx.input_buffer(0)[:n_sites].copy()
    st = gvcf.GvcfStitcher()
    got = []
    for lo in range(0, n_sites, 1111):
        n = min(1111, n_sites - lo)
        ctx.input_buffer(0)[:n] = hap_gt[lo:lo + n]
        ctx.submit(0, lo, n)
        ctx.wait(0)
        got += list(st.feed(ctx.gvcf_merge(0, rid[lo:lo + n], pos[lo:lo + n], dps), rid[lo:lo + n], pos[lo:lo + n]))
    got += list(st.finish())
    assert len(got) == len(want)
    for g, o in zip(got, want):
        assert g["kind"] == o["kind"]
        if o["kind"] == "site":
            assert g["site"] == idx[o["site"]]
        else:
            assert (g["first"], g["start"], g["end"], g["min_dp"], g["range"], g["n_members"]) == \
                (idx[o["first"]], o["start"], o["end"], o["min_dp"], o["range"], len(o["members"]))
            assert np.array_equal(g["dp"], o["dp"]) and np.array_equal(g["pl"].reshape(-1), o["pl"])
    ctx.close()


def test_cpp_host_driver_gvcf(tmp_path):
    """the C++ mirror (vgl::VcfTextSimulator + BatchSimulator::enable_gvcf + GvcfStitcher) reads a VCF, explodes it, merges on the
    device and stitches across batches of 4 sites: same record sequence and block values as the Python host path"""
    import os
    import subprocess
    from vcfgl_b200 import vcfinput
    exe = os.path.join(gu.ROOT, "vcfgl_b200", "host", "example_driver")
    if not os.path.exists(exe):
        pytest.skip("example_driver not built")
    S, n_rec, L = 4, 12, 90
    hap = synth.sfs_genotypes(n_rec, S, 31)
    pos = synth.positions(n_rec, L, 31)
    path = tmp_path / "in.vcf"
    synth.write_vcf(str(path), hap, pos, L)
    env = dict(os.environ, VGL_VCF_IN=str(path), VGL_EXPLODE="1", VGL_GVCF_DPS="1,2,4", VGL_BATCH="4")
    r = subprocess.run([exe], capture_output=True, text=True, env=env, timeout=120)
    assert r.returncode == 0, r.stderr
    got = [l.split("\t") for l in r.stdout.splitlines() if "skipped(" not in l]
    # Python host path on the driver's parameters
    a = vargs.parse_args("--seed 42 -d 4 -e 0.01 -GL 1 -doUnobserved 1 -doGVCF 1 --gvcf-dps 1,2,4 -addGL 1 -addPL 1 -addFormatAD 1 -addInfoDP 1".split())
    buf = path.read_bytes()
    hdr = vcfinput.read_header(buf)
    ctx = capi.Context(capi.params_from_args(a, S, max_batch_sites=4, n_slots=2))
    ps = ctx.parser(1 << 16, 16)
    st = gvcf.GvcfStitcher()
    want = []

    def emit(rec, sites_of):
        if rec["kind"] == "site":
            p, o = sites_of[rec["site"]]
            A = o["n_alleles"]
            want.append(["1", str(p + 1), "DP=%d" % o["info_dp"]] +
                        ["%d:%s" % (o["fmt_dp"][s], ",".join(str(x) for x in o["fmt_ad"].reshape(S, A)[s])) for s in range(S)])
        else:
            want.append(["1", str(rec["start"] + 1), "BLOCK", "END=%d" % (rec["end"] + 1), "MIN_DP=%d" % rec["min_dp"], "n=%d" % rec["n_members"]] +
                        ["%d:%d,%d,%d" % (rec["dp"][s], *rec["pl"][s]) for s in range(S)])
    sites_of, base = {}, 0
    for run, b in vcfinput.simulate_vcf_text(ctx, ps, buf[hdr.body_offset:], gt_source=0, explode=1, contigs=hdr.contigs):
        for i in range(b.n_sites):
            o = b.site(i)
            sites_of[base + i] = (int(run.pos[i]), {k: (np.array(v, copy=True) if isinstance(v, np.ndarray) else v) for k, v in o.items()})
        rid = np.zeros(b.n_sites, np.int32)
        for rec in st.feed(ctx.gvcf_merge(run.slot, rid, run.pos.astype(np.int32), [1, 2, 4]), rid, run.pos):
            emit(rec, sites_of)
        base += b.n_sites
    for rec in st.finish():
        emit(rec, sites_of)
    assert base == L and any(w[2] == "BLOCK" for w in want)
    assert len(got) == len(want)
    for g, w in zip(got, want):
        if w[2] == "BLOCK":
            assert [g[0], g[1], g[3], g[4], g[5], g[6]] + g[8:] == w, (g, w)
        else:
            assert [g[0], g[1], g[3]] + g[6:] == w, (g, w)
    ps.close()
    ctx.close()
