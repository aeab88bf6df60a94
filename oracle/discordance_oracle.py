"""oracle/discordance_oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

The genotype-call discordance the statistical parity tests use (tools/make_stats_golden.py, tests/test_gpu_native.py): per cell
of a written site with INFO/DP > 0 and FORMAT/DP > 0, call = the genotype with the single largest GL (a tie = no call =
discordant), compared as an unordered base pair with the true genotype; hom / het strata as misc/gtDiscordance.cpp:11-15.

Parity status: definition-level -- the reference has no counterpart in its main binary (misc/gtDiscordance.cpp compares a
bcftools call set with the truth); this file is the same computation tools/make_stats_golden.py applied to the reference's
captures when it produced tests/golden/stats.json."""
import numpy as np

PAIRS = [(a1, a2) for a2 in range(5) for a1 in range(a2 + 1)]


def site_counts(gl, fmt_dp, gts, alleles2acgt, n_genotypes):
    """-> [n_hom, d_hom, n_het, d_het] of one site; gts int[2S] ACGT ints"""
    S = len(fmt_dp)
    gl = np.asarray(gl, np.float32).reshape(S, n_genotypes)
    out = [0, 0, 0, 0]
    for s in np.flatnonzero(np.asarray(fmt_dp) > 0):
        row = gl[s]
        best = np.flatnonzero(row == row.max())
        call = None
        if len(best) == 1:
            a1, a2 = PAIRS[best[0]]
            call = tuple(sorted((int(alleles2acgt[a1]), int(alleles2acgt[a2]))))
        truth = tuple(sorted((int(gts[2 * s]), int(gts[2 * s + 1]))))
        k = 0 if truth[0] == truth[1] else 2
        out[k] += 1
        out[k + 1] += int(call != truth)
    return out
