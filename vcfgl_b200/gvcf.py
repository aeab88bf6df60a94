"""Host side of the gVCF block merger: stitching the device's per-batch records (vgl_gvcf_merge) into the run's record
sequence.  A batch's first block continues the previous batch's last one under the reference's own three conditions
(same contig, pos <= end + 1, same dp range: bcf_utils.cpp:711, 719, 790), with the same minima (:838-866)."""
from __future__ import annotations

from typing import Iterator, List, Optional

import numpy as np


class GvcfStitcher:
    """feed(batch result, rid[], pos[]) yields finished records; finish() yields the block still open at the end
    (write_record_values(NULL), vcfgl.cpp:169-177).  A record is a dict: kind 'site' (site = global index of the written
    site's batch entry) or 'block' (rid, start, end, min_dp, range, dp[S], pl[S,3] or None, first = founder)."""

    def __init__(self):
        self.open: Optional[dict] = None
        self.base = 0           # global index of the current batch's site 0

    def feed(self, res: dict, rid, pos) -> Iterator[dict]:
        recs = res["recs"]
        for k, r in enumerate(recs):
            f, l = int(r["first_site"]), int(r["last_site"])
            if r["n_members"] == 0:
                if self.open is not None:
                    yield self.open
                    self.open = None
                yield dict(kind="site", site=self.base + f)
                continue
            b = dict(kind="block", first=self.base + f, rid=int(rid[f]), start=int(pos[f]), end=int(pos[l]), min_dp=int(r["min_dp"]),
                     range=int(r["dp_range"]), n_members=int(r["n_members"]), dp=res["dp"][r["plane"]].copy(),
                     pl=None if res["pl"] is None else res["pl"][r["plane"]].copy())
            o = self.open
            if o is not None and k == 0 and o["rid"] == b["rid"] and b["start"] <= o["end"] + 1 and o["range"] == b["range"]:
                o["end"] = b["end"]
                o["min_dp"] = min(o["min_dp"], b["min_dp"])
                o["n_members"] += b["n_members"]
                np.minimum(o["dp"], b["dp"], out=o["dp"])
                if o["pl"] is not None:
                    g, m = o["pl"], b["pl"]
                    lower = (m[:, 1] < g[:, 1]) | ((m[:, 1] == g[:, 1]) & (m[:, 2] < g[:, 2]))
                    g[lower, 1] = m[lower, 1]
                    g[lower, 2] = m[lower, 2]
                continue
            if o is not None:
                yield o
            self.open = b
        self.base += len(rid)

    def finish(self) -> Iterator[dict]:
        if self.open is not None:
            yield self.open
            self.open = None
