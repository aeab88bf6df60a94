"""Turn reference replay captures (tests/vgl_dump.py) into C-ABI replay batches -- test infrastructure."""
import numpy as np

from vcfgl_b200 import synth


def batch_from_dump(sites, args):
    """sites: list[SiteDump] -> (gt uint8 [n_sites, S], replay dict for Context.submit)"""
    S = sites[0].S
    n = len(sites)
    gt = np.stack([synth.pack_gt(d.gts.reshape(1, 2 * S))[0] for d in sites])
    depths = np.concatenate([d.depths for d in sites]).astype(np.int32)
    eff = np.concatenate([d.fmt_dp for d in sites]).astype(np.int64)
    off = np.zeros(n * S + 1, np.int64)
    np.cumsum(eff, out=off[1:])
    cat = lambda f, dt: np.concatenate([getattr(d, f) for d in sites]).astype(dt) if n else np.zeros(0, dt)
    bases = cat("r_base", np.uint8)
    assert len(bases) == off[-1], (len(bases), off[-1])
    rp = dict(depths=depths, read_offsets=off, n_reads=int(off[-1]), bases=bases,
              strands=cat("r_strand", np.uint8))
    if args.error_qs == 2:
        rp["qs"] = np.clip(cat("r_qs", np.int64), 0, 255).astype(np.uint8)
        if args.adjust_qs:
            rp["adj_qs"] = np.clip(cat("r_adjqs", np.int64), 0, 255).astype(np.uint8)
        rp["error_probs"] = cat("r_eprob", np.float64)
    if args.add_i16:
        # sites that return before the tail loop have no tails recorded; pad so indices line up
        tails = []
        for d in sites:
            t = d.tails.astype(np.uint8)
            if len(t) != len(d.r_base):
                t = np.zeros(len(d.r_base), np.uint8)
            tails.append(t)
        rp["tail_dists"] = np.concatenate(tails)
    if args.gl_model == 1:
        codes, n_deep = [], 0
        for d in sites:
            o = 0
            n_site_deep = int((d.fmt_dp > 255).sum())
            if len(d.em_n) == 0 and n_site_deep:
                # the site returned before calculate_gls (--rm-invar-sites, vcfgl.cpp:677): the reference drew no
                # subsample, the ABI still takes one block per deep cell (ignored for a skipped site)
                codes += [np.zeros(255, np.uint16)] * n_site_deep
                n_deep += n_site_deep
                continue
            assert len(d.em_n) == n_site_deep, (len(d.em_n), n_site_deep)
            for k in range(len(d.em_n)):
                m = int(d.em_n[k])
                codes.append(d.em_codes[o:o + 255])
                o += m
                n_deep += 1
        rp["n_deep_cells"] = n_deep
        if n_deep:
            rp["deep_codes"] = np.concatenate(codes).astype(np.uint16)
    return gt, rp
