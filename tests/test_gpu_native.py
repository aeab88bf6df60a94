"""Native-RNG (Philox) simulator on the GPU.

(1) self-replay: the kernel's own draws (vgl_native_draws), fed to the CPU oracle and back through
    the replay path, must reproduce the native run's tags bit-exactly -- this checks the whole native
    pipeline at sizes where no reference capture exists, and that results do not depend on batch size.
(2) statistical parity with the reference (tests/golden/stats.json, made by
    tools/make_stats_golden.py from the instrumented reference): two-sample chi-square on the
    per-cell depth histogram, the true-base -> read-base matrix, the strand split and the per-read
    qs histogram; two-proportion z-test on genotype-call discordance; alpha = 0.001 each.
"""
import json
import os

import numpy as np
import pytest
from scipy import stats

import golden_cases as gc
import oracle_lib
from vcfgl_b200 import args as vargs
from vcfgl_b200 import capi, synth

pytestmark = pytest.mark.gpu
ALPHA = 1e-3
STATS = json.load(open(os.path.join(gc.GOLD, "stats.json")))
PAIRS = [(a1, a2) for a2 in range(5) for a1 in range(a2 + 1)]


def u32(x):
    return np.ascontiguousarray(x).view(np.uint32)


SELF_CASES = {
    "gl1_fixed_alltags": "--seed 5 -d 6 -e 0.02 -GL 1 -doUnobserved 1 -addGP 1 -addPL 1 -addI16 1 -addQS 1 -addInfoDP 1 "
                         "-addFormatAD 1 -addInfoAD 1 -addFormatADF 1 -addInfoADF 1 -addFormatADR 1 -addInfoADR 1",
    "gl2_fixed": "--seed 6 -d 4 -e 0.05 -GL 2 -doUnobserved 4 -addPL 1 -addFormatAD 1 -addQS 1",
    "gl2_eq2_lut": "--seed 7 -d 4 -e 0.01 -eq 2 -bv 1e-5 -GL 2 --adjust-qs 3 -addPL 1 -addQS 1 -addI16 1 -addFormatAD 1",
    "gl1_eq2": "--seed 8 -d 5 -e 0.02 -eq 2 -bv 1e-4 -GL 1 -addPL 1 -addFormatAD 1",
    "gl2_eq1": "--seed 9 -d 3 -e 0.05 -eq 1 -bv 1e-3 -GL 2 -addPL 1 -addFormatAD 1",
    "gl1_deep": "--seed 10 -d 280 -e 0.01 -GL 1 -addPL 1 -addFormatAD 1",
    "gl2_precise": "--seed 11 -d 4 -e 0.01 -eq 2 -bv 1e-5 -GL 2 --precise-gl 1 -addPL 1",
}


@pytest.mark.parametrize("name", sorted(SELF_CASES))
def test_native_tags_match_oracle_on_own_draws(name):
    S, n_sites = (6, 40) if "deep" in name else (37, 300)
    self_replay(name, SELF_CASES[name], 1, S, n_sites)


def self_replay(name, argv, sampler, S, n_sites, kernels=None, qs_bins=None):
    a = vargs.parse_args(argv.split(), qs_bins=qs_bins)
    hap = synth.sfs_genotypes(n_sites, S, 99, missing_rate=0.05) if S > 1 else \
        np.random.default_rng(1).integers(0, 2, (n_sites, 2)).astype(np.int8)
    gt = synth.pack_gt(hap)
    ctx = capi.Context(capi.params_from_args(a, S, max_batch_sites=n_sites, n_slots=1, sampler=sampler))
    if kernels:
        assert ctx.native_kernels() == kernels, ctx.native_kernels()
    ctx.input_buffer(0)[:n_sites] = gt
    first = 123456789012
    ctx.submit(0, first, n_sites)
    b = ctx.wait(0)
    native = [b.site(i) for i in range(n_sites)]
    native = [{k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in d.items()} for d in native]
    rp = ctx.native_draws(0, first, n_sites)
    assert np.array_equal(rp["depths"], b.dp)
    # (a) oracle on the kernel's draws
    orc = oracle_lib.Oracle(a, S)
    off = rp["read_offsets"]
    n_cmp = 0
    for i in range(n_sites):
        lo, hi = off[i * S], off[(i + 1) * S]
        if "deep" in name:
            break   # the oracle needs the kept-read capture for depth > 255; covered by (b) + golden d300 cases
        sl = slice(lo, hi)
        o = orc.site(hap[i], rp["depths"][i * S:(i + 1) * S], rp["bases"][sl], rp["strands"][sl],
                     None if rp["qs"] is None else rp["qs"][sl].astype(np.int32),
                     None if rp["adj_qs"] is None else rp["adj_qs"][sl].astype(np.int32),
                     None if rp["error_probs"] is None else rp["error_probs"][sl],
                     rp["tail_dists"][sl].astype(np.int32) if a.add_i16 else None)
        d = native[i]
        assert o["ret"] == d["skip_code"], (name, i)
        if o["ret"] != 0:
            continue
        assert o["n_alleles"] == d["n_alleles"] and o["n_genotypes"] == d["n_genotypes"]
        for key in ("pl", "fmt_ad", "fmt_adf", "fmt_adr", "info_ad", "info_adf", "info_adr"):
            if not getattr(a, "add_" + key) or d.get(key) is None:
                continue
            assert np.array_equal(o[key], d[key]), (name, i, key, o[key], d[key])
        if name != "gl2_precise":
            assert np.array_equal(u32(o["gl"]), u32(d["gl"])), (name, i)
        else:
            same = u32(o["gl"]) == u32(d["gl"])   # missing (a NaN payload) compares by bits
            with np.errstate(invalid="ignore"):
                near = np.abs(o["gl"].astype(np.float64) - d["gl"]) <= 1e-6 * np.abs(d["gl"].astype(np.float64))
            assert (same | near).all(), (name, i)
        if a.add_qs:
            assert np.array_equal(u32(o["qs"]), u32(d["qs"])), (name, i)
        if a.add_i16:
            assert np.array_equal(u32(o["i16"]), u32(d["i16"])), (name, i, o["i16"], d["i16"])
        n_cmp += 1
    assert n_cmp > 0 or "deep" in name
    # (b) replay of the kernel's draws through the C ABI, in two batches with different slots sizes
    if "deep" not in name:
        ctx.submit(0, 7, n_sites, replay=rp)
        b2 = ctx.wait(0)
        for i in range(n_sites):
            d, e = native[i], b2.site(i)
            assert d["skip_code"] == e["skip_code"]
            if d["skip_code"] == 0:
                assert np.array_equal(u32(d["gl"]), u32(e["gl"])), (name, i)
    # (c) batch-size independence: the same sites in two half batches give identical tags
    h = n_sites // 2
    ctx.input_buffer(0)[:n_sites - h] = gt[h:]
    ctx.submit(0, first + h, n_sites - h)
    b3 = ctx.wait(0)
    for i in range(h, n_sites):
        d, e = native[i], b3.site(i - h)
        assert d["skip_code"] == e["skip_code"]
        if d["skip_code"] == 0:
            assert np.array_equal(u32(d["gl"]), u32(e["gl"])), (name, i)
            assert np.array_equal(d["fmt_dp"], e["fmt_dp"])
    ctx.close()


def chi2_two_sample(a, b, min_expected=5):
    a, b = np.asarray(a, float).ravel(), np.asarray(b, float).ravel()
    keep = (a + b) > 0
    a, b = a[keep], b[keep]
    # pool sparse categories
    order = np.argsort(a + b)
    a, b = a[order], b[order]
    tot = a.sum() + b.sum()
    while len(a) > 2 and min((a[0] + b[0]) * a.sum() / tot, (a[0] + b[0]) * b.sum() / tot) < min_expected:
        a = np.concatenate([[a[0] + a[1]], a[2:]])
        b = np.concatenate([[b[0] + b[1]], b[2:]])
        order = np.argsort(a + b)
        a, b = a[order], b[order]
    return stats.chi2_contingency(np.stack([a, b]))[1]


@pytest.mark.parametrize("name", sorted(STATS))
def test_native_distributions_match_reference(name):
    distributions(name, 1)


def distributions(name, sampler, kernels=None, strand=True):
    st = STATS[name]
    argv = list(st["argv"])
    if not strand:   # kernels without strand tags: drop the ADF/ADR flags of the fixture's command line
        for flag in ("-addFormatADF", "-addFormatADR", "-addInfoADF", "-addInfoADR"):
            while flag in argv:
                i = argv.index(flag)
                del argv[i:i + 2]
    a = vargs.parse_args(argv, qs_bins=st.get("qs_bins"))
    S, n_sites = st["S"], st["n_sites"]
    hap = synth.sfs_genotypes(n_sites, S, st["gt_seed"])
    gt = synth.pack_gt(hap)
    ctx = capi.Context(capi.params_from_args(a, S, max_batch_sites=n_sites, n_slots=1, sampler=sampler))
    if kernels:
        assert ctx.native_kernels() == kernels, ctx.native_kernels()
    ctx.input_buffer(0)[:n_sites] = gt
    ctx.submit(0, 0, n_sites)
    b = ctx.wait(0)
    rp = ctx.native_draws(0, 0, n_sites)
    pvals = {}
    # depth
    depth_hist = np.bincount(np.minimum(b.dp, 199), minlength=200)
    pvals["depth"] = chi2_two_sample(depth_hist, st["depth_hist"])
    lam = a.depth
    expected = stats.poisson.pmf(np.arange(200), lam) * len(b.dp)
    pvals["depth_vs_poisson"] = chi2_two_sample(depth_hist, expected * 1e3)  # vs (almost) exact expectation
    # read-level draws
    cell_of_read = np.repeat(np.arange(n_sites * S), np.diff(rp["read_offsets"]))
    g0 = hap.reshape(-1, 2)[cell_of_read, 0]
    g1 = hap.reshape(-1, 2)[cell_of_read, 1]
    hom = g0 == g1
    conf = np.zeros((4, 4), np.int64)
    np.add.at(conf, (g0[hom], rp["bases"][hom]), 1)
    ref_conf = np.array(st["confusion"])
    for t in range(4):
        if ref_conf[t].sum() > 0:
            pvals["confusion_true%d" % t] = chi2_two_sample(conf[t], ref_conf[t])
    het = ~hom
    pvals["het_hap_pick"] = chi2_two_sample([(rp["bases"][het] == g0[het]).sum(), (rp["bases"][het] == g1[het]).sum()],
                                            st["het_reads"])
    if strand and sum(st["strand"][1:]) > 0:
        pvals["strand"] = chi2_two_sample(np.bincount(rp["strands"], minlength=2), st["strand"])
    if a.error_qs == 2:
        pvals["qs"] = chi2_two_sample(np.bincount(rp["qs"], minlength=256), st["qs_hist"])
    # genotype-call discordance (argmax GL vs truth), misc/gtDiscordance.cpp semantics
    disc = {"hom": [0, 0], "het": [0, 0]}
    for i in range(n_sites):
        d = b.site(i)
        if d["skip_code"] != 0 or d["info_dp"] == 0:
            continue
        G = d["n_genotypes"]
        gl = d["gl"].reshape(S, G)
        a2b = d["alleles2acgt"]
        mx = gl.max(axis=1)
        for s in np.flatnonzero(d["fmt_dp"] > 0):
            best = np.flatnonzero(gl[s] == mx[s])
            call = None
            if len(best) == 1:
                a1, a2 = PAIRS[best[0]]
                call = tuple(sorted((int(a2b[a1]), int(a2b[a2]))))
            truth = tuple(sorted((int(hap[i, 2 * s]), int(hap[i, 2 * s + 1]))))
            k = "hom" if truth[0] == truth[1] else "het"
            disc[k][0] += 1
            disc[k][1] += int(call != truth)
    for k in ("hom", "het"):
        n1, x1 = disc[k]
        n2, x2 = st["discordance"][k]
        pp = (x1 + x2) / (n1 + n2)
        if 0 < pp < 1:
            z = (x1 / n1 - x2 / n2) / np.sqrt(pp * (1 - pp) * (1 / n1 + 1 / n2))
            pvals["discordance_" + k] = 2 * stats.norm.sf(abs(z))
    bad = {k: v for k, v in pvals.items() if not (v >= ALPHA)}
    assert not bad, (name, bad, pvals)
    ctx.close()
