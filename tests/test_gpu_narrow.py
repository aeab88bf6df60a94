"""VGL_HOST_NARROW (include/vgl.h): the integer planes narrowed on the device (PL u8, AD/ADF/ADR/DP u8 or u16)
must widen back to exactly the int32 planes of VGL_HOST_I32 -- for every kernel set (tile GL model 1, tile GL
model 2, general kernels), in native and replay mode -- and a depth that does not fit raises VGL_EOVERFLOW."""
import copy

import numpy as np
import pytest

import golden_cases as gc
import replay_util
from vcfgl_b200 import args as vargs
from vcfgl_b200 import capi, synth

pytestmark = pytest.mark.gpu


def bits(x):
    return np.ascontiguousarray(x).view(np.uint32)


def run(a, S, gt, n_sites, mode, replay=None, first=777):
    ctx = capi.Context(capi.params_from_args(a, S, max_batch_sites=n_sites, n_slots=1, host_output=mode))
    ctx.input_buffer(0)[:n_sites] = gt
    ctx.submit(0, first, n_sites, replay=replay)
    b = ctx.wait(0)
    sites = [{k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in b.site(i).items()} for i in range(n_sites)]
    out = dict(status=b.status, narrow_bits=b.narrow_bits, sites=sites, kernels=ctx.native_kernels(),
               has_i32=b.raw.pl is not None or b.raw.ad is not None, launches=ctx.launch_count())
    ctx.close()
    return out


def same_sites(x, y):
    n = 0
    for i, (p, q) in enumerate(zip(x, y)):
        assert p.keys() == q.keys()
        for k in p:
            u, v = p[k], q[k]
            if u is None or v is None:
                assert u is None and v is None, (i, k)
            elif isinstance(u, np.ndarray):
                if u.dtype == np.float32:
                    assert np.array_equal(bits(u), bits(v)), (i, k)
                else:
                    assert np.array_equal(u, v), (i, k, u, v)
                n += u.size
            else:
                assert u == v, (i, k, u, v)
    return n


NATIVE = {
    # name: (argv, S, n_sites, kernel set, narrow bits)
    "tile_m1f": ("--seed 3 -d 10 -e 0.01 -GL 1 -addPL 1 -addFormatAD 1", 100, 500, "k_tile_m1f", 8),
    "tile_m1f_unobs": ("--seed 3 -d 3 -e 0.02 -GL 1 -doUnobserved 1 -addPL 1 -addFormatAD 1 -addInfoAD 1", 37, 333, "k_tile_m1f", 8),
    "tile_m2": ("--seed 4 -d 6 -e 0.01 -GL 2 -addPL 1 -addFormatAD 1", 50, 400, "k_tile_m2", 8),
    "tile_m1f_alltags": ("--seed 5 -d 6 -e 0.02 -GL 1 -doUnobserved 1 -addGP 1 -addPL 1 -addI16 1 -addQS 1 -addInfoDP 1 -addFormatAD 1 "
                         "-addInfoAD 1 -addFormatADF 1 -addInfoADF 1 -addFormatADR 1 -addInfoADR 1", 37, 300, "k_tile_m1f", 8),
    "general_alltags": ("--seed 5 -d 6 -e 0.02 -eq 2 -bv 1e-4 -GL 1 -doUnobserved 1 -addGP 1 -addPL 1 -addI16 1 -addQS 1 -addInfoDP 1 -addFormatAD 1 "
                        "-addInfoAD 1 -addFormatADF 1 -addInfoADF 1 -addFormatADR 1 -addInfoADR 1", 37, 300, "k_sim+k_site+k_scan+k_emit", 8),
    "deep_u16": ("--seed 6 -d 300 -e 0.01 -GL 1 -addPL 1 -addFormatAD 1", 6, 40, None, 16),
}


@pytest.mark.parametrize("name", sorted(NATIVE))
def test_narrow_equals_int32_native(name):
    argv, S, n_sites, kernels, nbits = NATIVE[name]
    a = vargs.parse_args(argv.split())
    gt = synth.pack_gt(synth.sfs_genotypes(n_sites, S, 11, missing_rate=0.05))
    wide = run(a, S, gt, n_sites, capi.HOST_I32)
    narrow = run(a, S, gt, n_sites, capi.HOST_NARROW)
    if kernels:
        assert wide["kernels"] == narrow["kernels"] == kernels
    assert wide["status"] == narrow["status"] == 0
    assert wide["narrow_bits"] == 0 and narrow["narrow_bits"] == nbits
    assert not narrow["has_i32"]            # the int32 integer planes do not cross PCIe
    assert narrow["launches"] > wide["launches"]
    assert same_sites(wide["sites"], narrow["sites"]) > 0
    # the sentinel convention: a missing PL appears exactly where FORMAT/DP is 0
    for s in wide["sites"]:
        if s["skip_code"] == 0 and s.get("pl") is not None:
            miss = (s["pl"].reshape(S, -1) == capi.I32_MISSING)
            assert np.array_equal(miss.all(1), s["fmt_dp"] == 0) and np.array_equal(miss.any(1), miss.all(1))


@pytest.mark.parametrize("cid", ["x_s40_gl1_cfg2", "x_s40_cfg4tags", "x_gl2_eq2_bins", "test13"] + gc.FUZZ_IDS)
def test_narrow_equals_int32_replay(cid):
    if cid in gc.FUZZ_MANIFEST:       # the random-configuration captures (tests/golden/fuzz)
        a, sites = gc.fuzz_args(cid), gc.fuzz_sites(cid)
    elif cid in gc.CASE_IDS:
        a, sites = gc.case_args(cid), gc.case_sites(cid)
    else:
        pytest.skip("no such capture")
    S = sites[0].S
    gt, rp = replay_util.batch_from_dump(sites, a)
    wide = run(a, S, gt, len(sites), capi.HOST_I32, replay=rp)
    narrow = run(a, S, gt, len(sites), capi.HOST_NARROW, replay=rp)
    assert wide["status"] == narrow["status"] == 0
    assert same_sites(wide["sites"], narrow["sites"]) > 0
    for k, d in enumerate(sites):     # and the reference's own integers
        if d.ret == 0 and d.out and "pl" in d.out:
            assert np.array_equal(narrow["sites"][k]["pl"], d.out["pl"])


def test_narrow_overflow_is_reported():
    """replayed depths of ~300 reads on a context whose depth law promises <= 255 (8-bit planes)"""
    cid = "x_gl1_d300"
    a = copy.copy(gc.case_args(cid))
    sites = gc.case_sites(cid)
    S = sites[0].S
    gt, rp = replay_util.batch_from_dump(sites, a)
    a.depth = 2.0
    out = run(a, S, gt, len(sites), capi.HOST_NARROW, replay=rp)
    assert out["narrow_bits"] == 8
    assert out["status"] == capi.VGL_EOVERFLOW
    for s in out["sites"]:
        assert s["fmt_dp"].max() <= 255
