"""Live pin of the CPU oracle (oracle/vgl_oracle.c) on the instrumented reference binary
(oracle/_ref/vcfgl_ref_dump, built by oracle/build_ref.sh from /root/reference): seeded random
configurations beyond tests/golden/ -- every valid combination class of (--gl-model, --error-qs,
--precise-gl, --adjust-qs, --qs-bins, -doUnobserved, --rm-invar-sites, --rm-empty-sites, tags,
missing genotypes, per-sample depths) -- are run through the reference here, and the oracle must
reproduce each captured site bit-for-bit from the captured draws.

Container only: skipped where oracle/_ref does not exist (the GPU box uses the committed captures
in tests/golden/ instead).  The arguments are validated by the CLI mirror first (vcfgl_b200/args.py,
io.cpp:860-1000), so a rejected combination is redrawn rather than handed to the reference."""
import os
import random
import subprocess

import numpy as np
import pytest

from fuzz_cases import draw_case, reference_exited
from vcfgl_b200 import args as vargs
import oracle_lib
import vgl_dump
from test_oracle_golden import OUT_KEYS, bits

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN_DUMP = os.path.join(ROOT, "oracle", "_ref", "vcfgl_ref_dump")

pytestmark = pytest.mark.skipif(not os.path.exists(BIN_DUMP), reason="oracle/_ref not built (needs /root/reference)")


@pytest.mark.parametrize("block", range(4))
def test_oracle_equals_live_reference_on_random_configurations(block, tmp_path):
    rnd = random.Random(7100 + block)
    n_sites_checked = n_values = 0
    for k in range(12):
        ref_argv, a, vcf, entry = draw_case(rnd, str(tmp_path), k)
        if "-addFormatDP" not in ref_argv and k % 2:      # FORMAT/DP is on by default (io.cpp): also run without it
            ref_argv += ["-addFormatDP", "0"]
            a = vargs.parse_args(entry["argv"] + ["-addFormatDP", "0"], qs_bins=entry["qs_bins"], depths=entry["depths"])
        dump = str(tmp_path / ("c%d.vgld" % k))
        r = subprocess.run([BIN_DUMP, "-i", vcf, "-o", str(tmp_path / ("c%d" % k))] + ref_argv,
                           capture_output=True, text=True, env=dict(os.environ, VGL_DUMP_PATH=dump))
        if r.returncode != 0 and reference_exited(r.stderr):
            continue        # the reference's own run-time exits (fuzz_cases.REFERENCE_EXITS)
        assert r.returncode == 0, (ref_argv, r.stderr[-1500:])
        if not os.path.exists(dump) or os.path.getsize(dump) == 0:
            continue
        sites = vgl_dump.read_dump(dump)
        orc = oracle_lib.Oracle(a, sites[0].S)
        for j, d in enumerate(sites):
            o = orc.site_from_dump(d)
            where = (ref_argv, j)
            assert o["ret"] == d.ret, where
            assert np.array_equal(o["fmt_dp"], d.fmt_dp), where
            assert o["info_dp"] == d.info_dp, where
            n_sites_checked += 1
            if d.ret != 0 or not d.out:
                continue
            assert (o["n_alleles"], o["n_alleles_observed"], o["n_genotypes"]) == \
                (d.n_alleles, d.n_alleles_observed, d.n_genotypes), where
            if d.info_dp > 0:
                assert np.array_equal(o["alleles2acgt"], d.alleles2acgt), where
            for key in OUT_KEYS:
                if key in d.out:
                    got, want = bits(o[key]), bits(d.out[key])
                    assert got.shape == want.shape and np.array_equal(got, want), (where, key)
                    n_values += got.size
    assert n_sites_checked > 50 and n_values > 1000
