"""oracle/truth_oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Restatement of the reference's --depth inf ("truth") mode, simulate_record_true_values() (vcfgl.cpp:1089-1262):
alleles = the bases present among the true haplotypes in descending count order (stable insertion sort, :1106-1117),
then the absent bases (-doUnobserved 3/4/5, :1129-1136) and the unobserved allele (-doUnobserved 1/2/4/5, :1138-1144);
per sample GL = 0 / -inf, GP = 1 / 0, PL = 0 / 255 (shared.h:205-212) at bcf_alleles2gt of its true alleles (:1207-1234).

Parity status: PINNED -- tests/test_truth_oracle.py rebuilds every record of the VCFs the unmodified reference wrote with
--depth inf (its own golden test4 and the runs of tools/make_golden_truth.py: all six -doUnobserved modes, --source 0 / 1,
-explode 0 / 1) from the input files' genotypes.
"""
import numpy as np

NEG_INF_BITS = 0xFF800000


def site(gts, do_unobserved, nonref="<*>"):
    """gts: int[2S] ACGT ints -> dict(alleles [str], alleles2acgt, acgt2alleles, n_alleles, n_alleles_observed, n_genotypes,
    gl float32 [S*G], gp float32 [S*G], pl int32 [S*G])"""
    gts = np.asarray(gts, np.int64)
    if (gts < 0).any() or (gts > 3).any():
        raise ValueError("missing / invalid true genotype (the reference asserts, vcfgl.cpp:1196)")
    S = len(gts) // 2
    ac = [int((gts == b).sum()) for b in range(4)]
    order = [0, 1, 2, 3]
    n_obs = 0
    for i in range(4):
        if ac[i] > 0:
            n_obs += 1
        j = i
        while j > 0 and ac[order[j]] > ac[order[j - 1]]:
            order[j], order[j - 1] = order[j - 1], order[j]
            j -= 1
    explode = do_unobserved in (3, 4, 5)
    unobs = do_unobserved in (1, 2, 4, 5)
    n_real = 4 if explode else n_obs
    alleles = ["ACGT"[order[k]] for k in range(n_real)]
    a2b = [-1] * 5
    b2a = [-1] * 5
    for k in range(n_real):
        a2b[k] = order[k]
        b2a[order[k]] = k
    if unobs:
        alleles.append("<NON_REF>" if do_unobserved in (2, 5) else "<*>")
        a2b[n_real] = 4
        b2a[4] = n_real
    A = len(alleles)
    G = A * (A + 1) // 2
    gl = np.full(S * G, -np.inf, np.float32)
    gp = np.zeros(S * G, np.float32)
    pl = np.full(S * G, 255, np.int32)
    for s in range(S):
        a, b = b2a[gts[2 * s]], b2a[gts[2 * s + 1]]
        hi, lo = max(a, b), min(a, b)
        t = hi * (hi + 1) // 2 + lo
        gl[s * G + t] = 0.0
        gp[s * G + t] = 1.0
        pl[s * G + t] = 0
    return dict(alleles=alleles, alleles2acgt=a2b, acgt2alleles=b2a, n_alleles=A, n_alleles_observed=n_obs, n_genotypes=G,
                gl=gl, gp=gp, pl=pl)
