"""The model-2 tile kernel (k_tile_m2: native RNG, GL model 2 with run-constant constants or the per-read qs LUT).

GL model 2 depends on the ORDER of a cell's reads, so the kernel's own read sequence is exported
(vgl_native_draws) and
(1) the CPU oracle and the replay kernels (pinned bit-exactly to the reference's captures) re-derive every tag from
    that sequence -> bit-exact; results do not depend on batch boundaries;
(2) the sampler has the reference's distributions: tests/test_gpu_native.py runs every gl2 fixture on k_tile_m2 (depth,
    true-base -> read-base matrix, haplotype pick, per-read quality scores, per-site error counts, genotype-call discordance).
"""
import pytest

import numpy as np

from test_gpu_native import self_replay
from vcfgl_b200 import args as vargs
from vcfgl_b200 import capi

pytestmark = pytest.mark.gpu

RTA3 = [(0, 2, 2), (3, 14, 12), (15, 30, 23), (31, 63, 37)]
CASES = {
    # name: (argv, S, n_sites, qs_bins)
    "fixed_cfg3i": ("--seed 3 -d 10 -e 0.01 -GL 2 -eq 1 -bv 1e-5 -addGL 1 -addPL 1", 120, 120, None),
    "fixed_eq0_star_ad": ("--seed 6 -d 4 -e 0.05 -GL 2 -doUnobserved 4 -addPL 1 -addFormatAD 1 -addInfoAD 1 -addInfoDP 1", 37, 300, None),
    "fixed_eq0_trim": ("--seed 16 -d 2 -e 0.1 -GL 2 -doUnobserved 0 --rm-invar-sites 4 --rm-empty-sites 1 -addPL 1 -addFormatAD 1", 5, 900, None),
    "fixed_eq0_explode5": ("--seed 17 -d 3 -e 0.3 -GL 2 -doUnobserved 5 -addPL 1 -addFormatAD 1", 3, 700, None),
    "fixed_eq1": ("--seed 9 -d 3 -e 0.05 -eq 1 -bv 1e-3 -GL 2 -addPL 1 -addFormatAD 1", 37, 300, None),
    "fixed_precise_eq0": ("--seed 18 -d 5 -e 0.02 -GL 2 --precise-gl 1 -addPL 1", 33, 200, None),
    "fixed_e0": ("--seed 19 -d 5 -e 0 -GL 2 -addPL 1 -addFormatAD 1", 33, 200, None),
    "fixed_e_high": ("--seed 12 -d 12 -e 0.9 -GL 2 -addPL 1 -addFormatAD 1", 33, 200, None),
    "fixed_deep": ("--seed 13 -d 70 -e 0.02 -GL 2 -addPL 1 -addFormatAD 1", 6, 60, None),
    "fixed_s1": ("--seed 10 -d 4 -e 0.05 -GL 2 -addPL 1 -addFormatAD 1", 1, 2000, None),
    "fixed_s1300": ("--seed 11 -d 8 -e 0.01 -GL 2 -addPL 1", 1300, 30, None),
    "fixed_s2501_scratch": ("--seed 14 -d 3 -e 0.02 -GL 2 -addGL 1 -addPL 1 -addFormatAD 1", 2501, 10, None),
    "lut_cfg3ii": ("--seed 4 -d 10 -e 0.01 -GL 2 -eq 2 -bv 1e-5 -addGL 1 -addPL 1", 120, 120, RTA3),
    "lut_eq2": ("--seed 7 -d 4 -e 0.01 -eq 2 -bv 1e-5 -GL 2 -addPL 1 -addFormatAD 1", 37, 300, None),
    "lut_eq2_adj": ("--seed 8 -d 4 -e 0.02 -eq 2 -bv 1e-4 -GL 2 --adjust-qs 1 -addPL 1 -addFormatAD 1", 37, 300, None),
    "lut_eq2_wide": ("--seed 20 -d 6 -e 0.2 -eq 2 -bv 0.02 -GL 2 -addPL 1 -addFormatAD 1", 37, 200, None),
    "lut_deep": ("--seed 21 -d 70 -e 0.02 -eq 2 -bv 1e-4 -GL 2 -addPL 1 -addFormatAD 1", 6, 60, None),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_tile_m2_tags_match_oracle_on_own_reads(name):
    argv, S, n_sites, bins = CASES[name]
    self_replay(name.replace("deep", "d70"), argv, 0, S, n_sites, kernels="k_tile_m2", qs_bins=bins)


@pytest.mark.parametrize("argv,kernels", [("--seed 31 -e 0.05 -GL 2 -addPL 1 -addFormatAD 1", "k_tile_m2"),
                                          ("--seed 32 -e 0.01 -eq 2 -bv 1e-4 -GL 2 -addPL 1 -addFormatAD 1", "k_tile_m2"),
                                          ("--seed 33 -e 0.02 -GL 1 -addPL 1 -addFormatAD 1", "k_tile_m1f"),
                                          ("--seed 34 -e 0.02 -GL 1 -addPL 1 -addI16 1 -addQS 1", "k_tile_m1f")])
def test_tile_kernels_with_per_sample_depths(argv, kernels):
    """--depths-file (vcfgl.cpp:1093-1100: each sample its own Poisson mean): one alias table per distinct mean on the device;
    the kernel's planes against the oracle on the kernel's own draws, and the mean depth of each sample against its mean"""
    S, n_sites = 37, 400
    depths = [0.5 + (i % 5) * 2.0 for i in range(S)]
    self_replay("df", argv, 0, S, n_sites, kernels=kernels, depths=depths)
    a = vargs.parse_args(argv.split(), depths=depths)
    ctx = capi.Context(capi.params_from_args(a, S, max_batch_sites=4000, n_slots=1))
    assert ctx.native_kernels() == kernels
    ctx.input_buffer(0)[:4000] = 0
    ctx.submit(0, 0, 4000)
    dp = ctx.wait(0).dp.reshape(4000, S).astype(np.float64)
    ctx.close()
    z = (dp.mean(axis=0) - np.array(depths)) / np.sqrt(np.array(depths) / 4000)
    assert np.abs(z).max() < 4.5, z
