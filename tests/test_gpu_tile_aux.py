"""The AUX variant of the model-1 tile kernel (k_tile_m1f with QS / I16 / INFO ADF, ADR: BASELINE.json configs[3]).

(1) self-replay: the kernel's count-level draws, listed read by read by vgl_native_draws (bases, strands, tail
    distances, the site's last read last), go through the CPU oracle and through the replay kernels (both pinned
    bit-exactly to the reference's captures): every tag incl. the float QS / I16 sums must come back bit-exact --
    also where the sums leave the exactly-representable range and are redone in the reference's order.
(2) the strand and tail-distance draws have the reference's laws (vcfgl.cpp:581-586, 653-656): fair strands,
    tail = min(1 + U{0..49}, 25); all of a site's tail mass lands on ONE base (the stale r_base, vcfgl.cpp:657).
"""
import numpy as np
import pytest

from test_gpu_native import self_replay
from vcfgl_b200 import args as vargs
from vcfgl_b200 import capi, synth

pytestmark = pytest.mark.gpu

CASES = {
    # name: (argv, S, n_sites)
    "cfg4": ("--seed 42 -d 10 -e 0.001 -GL 1 -doUnobserved 1 -doGVCF 1 --gvcf-dps 1,5,10 -addPL 1 -addI16 1 -addQS 1", 100, 300),
    "i16_qs_ad_info": ("--seed 5 -d 6 -e 0.02 -GL 1 -doUnobserved 1 -addPL 1 -addI16 1 -addQS 1 -addInfoDP 1 -addFormatAD 1 -addInfoAD 1 "
                       "-addInfoADF 1 -addInfoADR 1", 37, 300),
    "explode5": ("--seed 7 -d 3 -e 0.05 -GL 1 -doUnobserved 5 -addPL 1 -addI16 1 -addQS 1", 3, 700),
    "explode3_lowdepth": ("--seed 8 -d 0.3 -e 0.05 -GL 1 -doUnobserved 3 -addPL 1 -addI16 1 -addQS 1 -addInfoADF 1", 2, 1500),
    "trim_qs": ("--seed 6 -d 2 -e 0.1 -GL 1 -doUnobserved 0 --rm-invar-sites 4 --rm-empty-sites 1 -addPL 1 -addQS 1 -addInfoADF 1 -addInfoADR 1", 5, 900),
    "adjust_qs2": ("--seed 9 -d 5 -e 0.02 -GL 1 --adjust-qs 2 -doUnobserved 1 -addPL 1 -addQS 1 -addI16 1", 16, 300),
    "s1": ("--seed 10 -d 4 -e 0.05 -GL 1 -doUnobserved 1 -addPL 1 -addI16 1 -addQS 1", 1, 3000),
    "s1300": ("--seed 11 -d 8 -e 0.01 -GL 1 -doUnobserved 1 -addPL 1 -addI16 1 -addQS 1", 1300, 30),
    "s2501_scratch": ("--seed 14 -d 3 -e 0.02 -GL 1 -doUnobserved 1 -addPL 1 -addI16 1 -addQS 1 -addFormatAD 1", 2501, 10),
    "d100": ("--seed 13 -d 100 -e 0.01 -GL 1 -doUnobserved 1 -addPL 1 -addI16 1 -addQS 1", 6, 60),
    # I16 sums beyond 2^24 (mapq 60: 3600 per read): the float sums are order dependent -> sequential path
    "seq_s1300": ("--seed 15 -d 30 -e 0.01 -GL 1 -doUnobserved 1 --i16-mapq 60 -addPL 1 -addI16 1 -addQS 1", 1300, 8),
    "seq_s2501_scratch": ("--seed 16 -d 8 -e 0.01 -GL 1 -doUnobserved 1 --i16-mapq 60 -addPL 1 -addI16 1 -addQS 1", 2501, 6),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_tile_aux_tags_match_oracle_on_own_draws(name):
    argv, S, n_sites = CASES[name]
    self_replay(name, argv, 0, S, n_sites, kernels="k_tile_m1f")


def test_tile_aux_strand_and_tail_laws():
    S, n_sites = 50, 4000
    a = vargs.parse_args("--seed 21 -d 10 -e 0.01 -GL 1 -doUnobserved 1 -addPL 1 -addI16 1 -addInfoAD 1 -addInfoADF 1 -addInfoADR 1".split())
    hap = synth.sfs_genotypes(n_sites, S, 77)
    ctx = capi.Context(capi.params_from_args(a, S, max_batch_sites=n_sites, n_slots=1))
    assert ctx.native_kernels() == "k_tile_m1f"
    ctx.input_buffer(0)[:] = synth.pack_gt(hap)
    ctx.submit(0, 31337, n_sites)
    b = ctx.wait(0)
    s = b.sites
    keep = (s["skip_code"] == 0) & (s["info_dp"] > 0)
    fwd, rev, reads = s["info_adf"][keep].sum(), s["info_adr"][keep].sum(), s["info_dp"][keep].sum()
    assert fwd + rev == reads == s["info_ad"][keep].sum()
    z = (fwd - reads / 2) / np.sqrt(reads / 4)
    assert abs(z) < 4.5, ("strand", fwd, rev, z)
    i16 = s["i16"][keep].astype(np.float64)
    assert np.array_equal(i16[:, 0] + i16[:, 2], s["info_adf"][keep].sum(axis=1))     # forward counts, ref + non-ref
    assert np.array_equal(i16[:, 1] + i16[:, 3], s["info_adr"][keep].sum(axis=1))
    # all tail mass of a site on one side (ref: 12/13, non-ref: 14/15), on exactly one base
    assert ((i16[:, 12] == 0) | (i16[:, 14] == 0)).all()
    t1, t2 = (i16[:, 12] + i16[:, 14]).sum(), (i16[:, 13] + i16[:, 15]).sum()
    # tail = min(1 + U{0..49}, 25): mean 19, mean square 423; variances 62 and 59456.4
    assert abs(t1 / reads - 19.0) < 4.5 * np.sqrt(62.0 / reads), t1 / reads
    assert abs(t2 / reads - 423.0) < 4.5 * np.sqrt(59456.4 / reads), t2 / reads
    # the stale base is a uniformly chosen read of the last cell: at sites whose last cell is heterozygous 0|1 without
    # mis-called reads the tail mass is on the reference allele (allele 0 = the site's most frequent base) about as often
    # as that cell's reads carry it -- checked in aggregate: P(ref side) must be well inside (0, 1)
    frac_ref = float((i16[:, 12] > 0).mean())
    assert 0.5 < frac_ref < 1.0, frac_ref
    ctx.close()
