"""oracle/gvcf_oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Sequential restatement of the reference's gVCF block merger, prepare_gvcf_block() (bcf_utils.cpp:662-942, driven by
write_record_values(), vcfgl.cpp:165-207), on the per-site tag arrays:

  a site is a block MEMBER iff  nAllelesObserved == 1 (bcf_utils.cpp:692, 705)  and its dp range is valid: r = number of
  --gvcf-dps thresholds <= min over samples of FORMAT/DP, r >= 1 (bcf_utils.cpp:741-765);
  it JOINS the block in memory iff same contig (:711), pos <= end_pos + 1 (:719) and same dp range (:790); otherwise the
  block is flushed and the site founds a new one (or, if not a member, is written as a regular record);
  block values: start = founder's pos, end = last member's pos, MIN_DP = min of the members' min DP (:838-842),
  DP[s] = min over members (:844-848), PL[s] = (founder's PL[3s], lexicographic min over members of (PL[3s+1], PL[3s+2]))
  (:858-866); alleles / QS come from the founder (:817-832).

Input sites are those the reference WRITES (simulate_record_values returned 0), in order.

Parity status: PINNED -- tests/test_gvcf_oracle.py rebuilds, from the instrumented reference's per-site captures, the record
list of the BCF the unmodified reference wrote for the same run: tests/golden/gvcf/ (5 runs, tools/make_golden_gvcf.py) and the
gVCF cases of the main golden set (test7, test8, test19).
"""
import numpy as np


def dp_range(min_dp, dps):
    r = 0
    for t in dps:
        if min_dp < t:
            break
        r += 1
    return r


def merge(sites, dps):
    """sites: list of dicts(rid, pos, n_alleles_observed, fmt_dp int32[S], pl int32[S*G] or None)
    -> list of records: dict(kind='site', site=i) | dict(kind='block', first=i, last=j, members=[...], rid, start, end,
    min_dp, range, dp int32[S], pl int32[S*3] or None)"""
    out = []
    cur = None
    for i, s in enumerate(sites):
        dp = np.asarray(s["fmt_dp"], np.int32)
        min_dp = int(dp.min())
        r = dp_range(min_dp, dps)
        member = s["n_alleles_observed"] == 1 and r >= 1
        if cur is not None:
            joins = member and s["rid"] == cur["rid"] and s["pos"] <= cur["end"] + 1 and r == cur["range"]
            if not joins:
                out.append(cur)
                cur = None
        if not member:
            out.append(dict(kind="site", site=i))
            continue
        if cur is None:
            pl = None if s.get("pl") is None else np.array(s["pl"], np.int32).copy()
            cur = dict(kind="block", first=i, last=i, members=[i], rid=s["rid"], start=s["pos"], end=s["pos"], min_dp=min_dp,
                       range=r, dp=dp.copy(), pl=pl)
            continue
        cur["members"].append(i)
        cur["last"] = i
        cur["end"] = s["pos"]
        cur["min_dp"] = min(cur["min_dp"], min_dp)
        np.minimum(cur["dp"], dp, out=cur["dp"])
        if cur["pl"] is not None and s.get("pl") is not None:
            g = cur["pl"].reshape(-1, 3)
            m = np.asarray(s["pl"], np.int32).reshape(-1, 3)
            lower = (m[:, 1] < g[:, 1]) | ((m[:, 1] == g[:, 1]) & (m[:, 2] < g[:, 2]))
            g[lower, 1] = m[lower, 1]
            g[lower, 2] = m[lower, 2]
    if cur is not None:
        out.append(cur)
    return out
