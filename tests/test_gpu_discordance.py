"""On-device discordance summary (k_discordance behind vgl_discordance) against oracle/discordance_oracle.py (the definition the
statistical parity tests use).  Counts: exact.
(1) reference captures replayed on the device (GL bit-exact), (2) native batches of the bench shapes incl. skipped sites, cells
without reads and missing genotypes, (3) argument rules."""
import os
import sys

import numpy as np
import pytest

import golden_cases as gc
import replay_util
from vcfgl_b200 import args as vargs
from vcfgl_b200 import capi, synth

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import discordance_oracle as do  # noqa: E402

pytestmark = pytest.mark.gpu


def oracle_counts(b, gts_rows):
    tot = np.zeros(4, np.int64)
    for i in range(b.n_sites):
        o = b.site(i)
        if o["skip_code"] != 0 or o["info_dp"] == 0:
            continue
        tot += do.site_counts(o["gl"], o["fmt_dp"], gts_rows[i], o["alleles2acgt"], o["n_genotypes"])
    return tot.tolist()


CASES = [c for c in gc.CASE_IDS if gc.case_args(c).add_gl and gc.case_args(c).add_fmt_dp and gc.case_args(c).depth != float("inf")]


@pytest.mark.parametrize("cid", CASES)
def test_replayed_reference_runs(cid):
    a = gc.case_args(cid)
    sites = gc.case_sites(cid)
    S = sites[0].S
    gt, rp = replay_util.batch_from_dump(sites, a)
    ctx = capi.Context(capi.params_from_args(a, S, max_batch_sites=len(sites), n_slots=1))
    ctx.input_buffer(0)[:len(sites)] = gt
    ctx.submit(0, 0, len(sites), replay=rp)
    b = ctx.wait(0)
    d = ctx.discordance(0)
    assert d["hom"] + d["het"] == oracle_counts(b, [s.gts for s in sites])
    ctx.close()


@pytest.mark.parametrize("argv,S,n", [
    ("--seed 42 -d 10 -e 0.01 -GL 1 -addGL 1 -addPL 1 -addFormatAD 1", 100, 3000),
    ("--seed 42 -d 1 -e 0.05 -GL 1 -doUnobserved 1 -addGL 1 --rm-empty-sites 1", 3, 4000),
    ("--seed 42 -d 3 -e 0.02 -GL 2 -eq 2 -bv 1e-4 -doUnobserved 4 -addGL 1 --rm-invar-sites 4", 17, 2000),
    ("--seed 42 -d 30 -e 0.01 -GL 1 -addGL 1", 1000, 60),
])
def test_native_batches(argv, S, n):
    a = vargs.parse_args(argv.split())
    hap = synth.sfs_genotypes(n, S, 9, missing_rate=0.03)
    ctx = capi.Context(capi.params_from_args(a, S, max_batch_sites=n, n_slots=1))
    ctx.input_buffer(0)[:n] = synth.pack_gt(hap)
    ctx.submit(0, 50, n)
    b = ctx.wait(0)
    d = ctx.discordance(0)
    want = oracle_counts(b, hap)
    assert d["hom"] + d["het"] == want
    assert want[0] > 0 and want[2] > 0
    ctx.close()


def test_needs_gl_and_a_waited_batch():
    a = vargs.parse_args("--seed 1 -d 2 -e 0.01 -GL 1 -addGL 0 -addPL 1".split())
    ctx = capi.Context(capi.params_from_args(a, 2, max_batch_sites=4, n_slots=1))
    with pytest.raises(capi.VglError):
        ctx.discordance(0)                      # nothing submitted
    ctx.input_buffer(0)[:4] = 0
    ctx.submit(0, 0, 4)
    ctx.wait(0)
    with pytest.raises(capi.VglError):
        ctx.discordance(0)                      # no GL plane
    ctx.close()
