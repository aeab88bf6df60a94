#!/usr/bin/env python3
"""oracle/patch_ref_for_dump.py -- TEST INFRASTRUCTURE.

Writes instrumented *temporary* copies of the reference's vcfgl.cpp and
gl_methods.cpp (read from REF_DIR, never modified) into OUT_DIR, injecting the
replay-capture hooks of oracle/ref_dump_hooks.h.  The copies are build
intermediates of oracle/build_ref.sh and are deleted after compilation; they
are never committed.

Anchors are matched on exact statement text (not line numbers) and every
anchor must match exactly the expected number of times, otherwise we abort --
so a changed reference cannot be silently mis-instrumented.

usage: patch_ref_for_dump.py REF_DIR OUT_DIR [--vgl]

--vgl: additionally route the hot path through libvgl (oracle/ref_vgl_binding.h): the wrapper of
simulate_record_values() replays the captured draws on the GPU and overwrites the record's arrays.
"""
import sys, os, re

ref, out = sys.argv[1], sys.argv[2]
VGL = "--vgl" in sys.argv[3:]


def patch(text, anchor, repl, count, mode="after"):
    """insert `repl` after/before every line containing `anchor`"""
    lines = text.split("\n")
    n = 0
    res = []
    for ln in lines:
        hit = anchor in ln
        if hit and mode == "before":
            res.append(repl)
        res.append(ln)
        if hit and mode == "after":
            res.append(repl)
        n += hit
    if n != count:
        sys.exit("patch_ref_for_dump: anchor %r matched %d times, expected %d" % (anchor, n, count))
    return "\n".join(res)


# ---- vcfgl.cpp ----------------------------------------------------------
src = open(os.path.join(ref, "vcfgl.cpp")).read()
src = patch(src, '#include "gl_methods.h"',
            '#define VGL_DUMP_NEED_SIMRECORD 1\n#include "ref_dump_hooks.h"' + ('\n#include "ref_vgl_binding.h"' if VGL else ''), 1)
# rename the hot-path entry and wrap it (vcfgl.cpp:327)
a = "static int simulate_record_values(simRecord* sim) {"
if src.count(a) != 1:
    sys.exit("entry anchor mismatch")
src = src.replace(a, "static int simulate_record_values_VGLORIG(simRecord* sim) {")
wrapper = (
    "static int simulate_record_values(simRecord* sim) {\n"
    "    vgl_dump_begin(sim);\n"
    "    int vgl_ret = simulate_record_values_VGLORIG(sim);\n"
    "    vgl_dump_end(sim, vgl_ret);\n"
    + ("    vgl_bind_replace(sim, vgl_ret);\n" if VGL else "") +
    "    return vgl_ret;\n"
    "}\n")
src = patch(src, "static int simulate_record_true_values(simRecord* sim) {", wrapper, 1, mode="before")
# site-level beta draw (vcfgl.cpp:428)
src = patch(src, "base_pick_error_prob = args->betaSampler->sample();",
            "        vgl_dump_site_eprob(base_pick_error_prob);", 1)
# per read (vcfgl.cpp:610)
src = patch(src, "sim->bases[s][read_i] = r_base;",
            "                vgl_dump_read(s, r_base, which_strand, qScore_i, adjqScore_i, error_prob_forQs_i);", 1)
# tail distance (vcfgl.cpp:658)
src = patch(src, "sim->acgt_sum_taildist_sq[r_base] += (tail_dist * tail_dist);",
            "                    vgl_dump_tail(tail_dist, r_base);", 1)
open(os.path.join(out, "vcfgl.cpp"), "w").write(src)

# ---- gl_methods.cpp -------------------------------------------------------
src = open(os.path.join(ref, "gl_methods.cpp")).read()
src = patch(src, '#include "io.h"', '#include "ref_dump_hooks.h"', 1)
# after both errmod_cal call sites (gl_methods.cpp:266,333)
src = patch(src, "errmod_cal(args->gl1errmod, n_sim_reads, 5, ubases, fpls);",
            "            vgl_dump_errmod(s, n_sim_reads, ubases);", 2)
open(os.path.join(out, "gl_methods.cpp"), "w").write(src)
print("patched copies written to", out)
