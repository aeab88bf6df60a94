"""-printPileup from the CUDA simulator's own draws (Context.native_draws -> vcfgl_b200/pileup.py): every sample column must
agree with the tags of the same batch -- depth = FORMAT/DP, base letters count up to FORMAT/AD.  (The text format itself is
pinned on the reference's pileup files by tests/test_pileup.py.)"""
import numpy as np
import pytest

from vcfgl_b200 import args as vargs
from vcfgl_b200 import capi, pileup, synth

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("argv", ["--seed 42 -d 3 -e 0.05 -GL 1 -doUnobserved 1 -addPL 1 -addFormatAD 1",
                                  "--seed 7 -d 2 -e 0.02 -eq 2 -bv 1e-4 -GL 2 -doUnobserved 1 -addPL 1 -addFormatAD 1 --rm-empty-sites 1"])
def test_pileup_of_native_draws_matches_tags(argv):
    S, n = 5, 200
    a = vargs.parse_args(argv.split())
    hap = synth.sfs_genotypes(n, S, 5, missing_rate=0.05)
    ctx = capi.Context(capi.params_from_args(a, S, max_batch_sites=n, n_slots=1))
    ctx.input_buffer(0)[:n] = synth.pack_gt(hap)
    ctx.submit(0, 1000, n)
    b = ctx.wait(0)
    draws = ctx.native_draws(0, 1000, n)
    skip = [b.site(i)["skip_code"] for i in range(n)]
    text = pileup.format_pileup(a, ["1"] * n, list(range(n)), [0] * n, skip, draws, S)
    lines = text.decode().splitlines()
    kept = [i for i in range(n) if skip[i] != -4]
    assert len(lines) == len(kept)
    for line, i in zip(lines, kept):
        f = line.split("\t")
        assert f[:3] == ["1", str(i + 1), "A"] and len(f) == 3 + 3 * S
        o = b.site(i)
        for s in range(S):
            dp, bases, quals = f[3 + 3 * s: 6 + 3 * s]
            assert int(dp) == o["fmt_dp"][s]
            if int(dp) == 0:
                assert (bases, quals) == ("*", "*")
                continue
            assert len(bases) == len(quals) == int(dp)
            if o["skip_code"] == 0:
                ad = o["fmt_ad"].reshape(S, o["n_alleles"])[s]
                for al in range(o["n_alleles"]):
                    base = o["alleles2acgt"][al]
                    if base < 4:
                        assert bases.count("ACGT"[base]) == ad[al]
    ctx.close()
