"""--depth inf on the device (k_truth_site, k_scan, k_truth_emit; include/vgl.h VGL_DEPTH_INF) against oracle/truth_oracle.py,
which is pinned on the reference's own --depth inf output (tests/test_truth_oracle.py).  Exact: the values are 0 / -inf / 1 / 255.
(1) the reference's runs: same genotypes in, same alleles and GL / GP / PL out;
(2) random genotype matrices of 1 .. 1000 samples, all six -doUnobserved modes;
(3) a missing true genotype raises VGL_EMISSING (the reference asserts, vcfgl.cpp:1196); argument rules of io.cpp:783-853."""
import numpy as np
import pytest

import truth_util as tu
from vcfgl_b200 import args as vargs
from vcfgl_b200 import capi


pytestmark = pytest.mark.gpu


def pack(gts_rows):
    g = np.asarray(gts_rows, np.int64)
    g = np.where(g < 0, 0xF, g)
    return (g[:, 0::2] | (g[:, 1::2] << 4)).astype(np.uint8)


def check(a, S, gts_rows):
    n = len(gts_rows)
    ctx = capi.Context(capi.params_from_args(a, S, max_batch_sites=n, n_slots=1))
    assert ctx.native_kernels() == "k_truth_site+k_scan+k_truth_emit"
    ctx.input_buffer(0)[:n] = pack(gts_rows)
    ctx.submit(0, 0, n)
    b = ctx.wait(0)
    assert b.status == 0
    for i in range(n):
        o, w = b.site(i), tu.to.site(gts_rows[i], a.do_unobserved)
        assert (o["skip_code"], o["info_dp"]) == (0, -1)
        for k in ("n_alleles", "n_alleles_observed", "n_genotypes"):
            assert o[k] == w[k], (i, k)
        assert o["alleles2acgt"].tolist() == w["alleles2acgt"] and o["acgt2alleles"].tolist() == w["acgt2alleles"], i
        for k, on in (("gl", a.add_gl), ("gp", a.add_gp), ("pl", a.add_pl)):
            if on:
                assert np.array_equal(np.ascontiguousarray(o[k]).view(np.uint32), w[k].view(np.uint32)), (i, k)
            else:
                assert o[k] is None
    ctx.close()


@pytest.mark.parametrize("cid", tu.CASES)
def test_reference_runs(cid):
    a, S, seq = tu.case(cid)
    check(a, S, np.array([g for _, g in seq]))


@pytest.mark.parametrize("S,n_sites,u", [(1, 500, 0), (2, 500, 1), (7, 300, 2), (100, 300, 3), (129, 200, 4), (1000, 40, 5), (1000, 40, 1)])
def test_random_genotypes(S, n_sites, u):
    rng = np.random.default_rng(S * 10 + u)
    k = rng.integers(1, 5, n_sites)                       # how many different bases a site has
    gts = np.stack([rng.permutation(4)[rng.integers(0, k[i], 2 * S)] for i in range(n_sites)])
    a = vargs.parse_args(("--seed 1 --depth inf -e 0 -GL 1 -doUnobserved %d -addGL 1 -addGP 1 -addPL 1 -addFormatDP 0" % u).split())
    check(a, S, gts)


def test_missing_genotype_and_argument_rules():
    a, S, seq = tu.case("t_missing_refused")
    gts = np.array([g for _, g in seq])
    ctx = capi.Context(capi.params_from_args(a, S, max_batch_sites=len(gts), n_slots=1))
    ctx.input_buffer(0)[:len(gts)] = pack(gts)
    ctx.submit(0, 0, len(gts))
    assert ctx.wait(0).status == capi.VGL_EMISSING
    ctx.close()
    for bad in ("--depth inf -e 0.01 -GL 1 -addFormatDP 0", "--depth inf -e 0 -GL 1", "--depth inf -e 0 -GL 1 -addFormatDP 0 -addFormatAD 1",
                "--depth inf -e 0 -GL 1 -addFormatDP 0 --rm-invar-sites 4"):
        with pytest.raises(vargs.ArgError):
            vargs.parse_args(("--seed 1 " + bad).split())
    p = capi.params_from_args(vargs.parse_args("--seed 1 --depth inf -e 0 -GL 1 -addFormatDP 0".split()), 2, max_batch_sites=4)
    p.tag_mask |= vargs.TAG_FMT_DP
    with pytest.raises(capi.VglError):
        capi.Context(p)
