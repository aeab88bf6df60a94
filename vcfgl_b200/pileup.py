"""-printPileup 1 from the simulator's draws: the text the reference writes per site while it
simulates (vcfgl.cpp:414-416, 445-449, 616-643; no-reads sites :230-235), rebuilt from the per-read draws in the replay
layout -- the reference's own capture, or `Context.native_draws()` of the CUDA simulator.

    <contig> <pos> <REF[0]> { <depth> <bases> <quals> | 0 * * } per sample

Qualities are Phred+33 of the per-read quality score (--error-qs 2) or of the run's fixed score (vcfgl.cpp:1661-1703),
adjusted ones with --adjust-qs bit 4 (shared.h:111, 187-188).  Sites dropped by --rm-empty-sites print nothing
(vcfgl.cpp:396-403); sites dropped as simulated-invariant (-3, vcfgl.cpp:677) were already printed."""
from __future__ import annotations

import math
from typing import Optional

import numpy as np

from . import args as vargs

CAP_BASEQ = 63            # shared.h:241
ADJUST_FOR_PILEUP = 1 << 2


def fixed_qscores(a: vargs.SimArgs):
    """(qScore, adj_qScore or None) of a run with --error-qs 0 | 1 (vcfgl.cpp:1661-1703)"""
    e = a.error_rate
    adj = None
    if e == 0.0:
        qs, adj_raw = CAP_BASEQ, CAP_BASEQ
    elif e == 1.0:
        qs, adj_raw = 0, 0
    else:
        t = -10.0 * math.log10(e)
        qs, adj_raw = int(t), int(t + a.adjust_by)

    def bins(q):
        if a.qs_bins:
            for lo, hi, v in a.qs_bins:
                if lo <= q <= hi:
                    return v
            raise ValueError("Could not find a range for qs value %d" % q)
        return min(q, CAP_BASEQ)
    qs = bins(qs)
    if a.adjust_qs:
        adj = bins(adj_raw)
    return qs, adj


def format_pileup(a: vargs.SimArgs, contigs, pos, ref_acgt, skip_codes, draws: dict, n_samples: int) -> bytes:
    """contigs[i] (str), pos[i] (0-based), ref_acgt[i] (ACGT int of REF), skip_codes[i] of the batch's sites;
    draws: replay-layout dict (depths [n*S] as drawn, read_offsets [n*S+1], bases, qs / adj_qs or None)"""
    S = n_samples
    depths = np.asarray(draws["depths"])
    off = np.asarray(draws["read_offsets"])
    bases = np.asarray(draws["bases"]) if draws.get("bases") is not None else np.zeros(0, np.uint8)
    use_adj = bool(a.adjust_qs & ADJUST_FOR_PILEUP)
    per_read = draws.get("adj_qs" if use_adj else "qs") if a.error_qs == 2 else None
    fixed = None
    if per_read is None:
        q, adj = fixed_qscores(a)
        fixed = (adj if use_adj else q) + 33
    acgt = np.frombuffer(b"ACGT", np.uint8)
    out = []
    for i in range(len(pos)):
        if skip_codes[i] == -4:
            continue
        line = [("%s\t%d\t%s" % (contigs[i], pos[i] + 1, "ACGT"[ref_acgt[i]])).encode()]
        for s in range(S):
            c = i * S + s
            n = int(off[c + 1] - off[c])        # reads actually simulated (0 for a missing genotype)
            if n == 0:
                line.append(b"\t0\t*\t*")
                continue
            b = acgt[bases[off[c]:off[c + 1]]].tobytes()
            ql = (np.asarray(per_read[off[c]:off[c + 1]], np.uint8) + 33).tobytes() if per_read is not None else bytes([fixed]) * n
            line.append(b"\t%d\t%s\t%s" % (n, b, ql))
        out.append(b"".join(line) + b"\n")
    return b"".join(out)
