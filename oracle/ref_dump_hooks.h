/*
 * oracle/ref_dump_hooks.h -- TEST INFRASTRUCTURE (not product code).
 *
 * Replay-capture hooks that oracle/build_ref.sh injects into a *temporary,
 * patched copy* of the read-only reference (vcfgl.cpp / gl_methods.cpp) to
 * build oracle/_ref/vcfgl_ref_dump.  The patched sources never enter this
 * repository; only this hook header (our own code) does.
 *
 * The instrumented binary behaves exactly like the reference (same RNG
 * consumption, same VCF output; tools/make_golden.py re-checks that against
 * the reference's own golden VCFs) and additionally writes, for every call of
 * simulate_record_values(), one binary record holding
 *   - the hot-path INPUT   (true genotypes as ACGT ints, vcfgl.cpp:66,133-146)
 *   - every DRAW           (depths vcfgl.cpp:364-368, site beta error :428,
 *                           per read base/strand/qs/adj-qs/error-prob :473-610,
 *                           tail distances + stale r_base :647-663,
 *                           post-shuffle read codes when depth>255,
 *                           htslib/errmod.c:156-159)
 *   - the bit-exact OUTPUT (return code, allele maps, all tag arrays exactly
 *                           as add_tags() hands them to htslib,
 *                           bcf_utils.cpp:426-507)
 * to the file named by $VGL_DUMP_PATH.  Layout: see tests/vgl_dump.py.
 */
#ifndef VGL_REF_DUMP_HOOKS_H
#define VGL_REF_DUMP_HOOKS_H

#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

struct vgl_dump_state {
    FILE* fp = NULL;
    bool tried = false;
    int has_site_eprob = 0;
    double site_eprob = -1.0;
    int stale_base = -1;
    std::vector<int32_t> r_sample, r_qs, r_adjqs, tails;
    std::vector<uint8_t> r_base, r_strand;
    std::vector<double> r_eprob;
    std::vector<int32_t> em_sample, em_n;
    std::vector<uint16_t> em_codes;
};

inline vgl_dump_state& vgl_ds() {
    static vgl_dump_state st;
    return st;
}

inline FILE* vgl_dump_fp() {
    vgl_dump_state& st = vgl_ds();
    if (!st.tried) {
        st.tried = true;
        const char* p = getenv("VGL_DUMP_PATH");
        if (p && *p) st.fp = fopen(p, "wb");
    }
    return st.fp;
}

inline void vgl_dump_site_eprob(double e) {
    vgl_ds().has_site_eprob = 1;
    vgl_ds().site_eprob = e;
}

inline void vgl_dump_read(int s, int base, int strand, int qs, int adjqs, double eprob) {
    vgl_dump_state& st = vgl_ds();
    st.r_sample.push_back(s);
    st.r_base.push_back((uint8_t)base);
    st.r_strand.push_back((uint8_t)strand);
    st.r_qs.push_back(qs);
    st.r_adjqs.push_back(adjqs);
    st.r_eprob.push_back(eprob);
}

inline void vgl_dump_tail(int tail, int r_base) {
    vgl_ds().tails.push_back(tail);
    vgl_ds().stale_base = r_base;
}

/* called after errmod_cal(): ubases is then shuffled+sorted in place */
inline void vgl_dump_errmod(int s, int n, const uint16_t* codes) {
    if (n <= 255) return;
    vgl_dump_state& st = vgl_ds();
    st.em_sample.push_back(s);
    st.em_n.push_back(n);
    for (int i = 0; i < n; ++i) st.em_codes.push_back(codes[i]);
}

template <typename T>
inline void vgl_w(FILE* fp, const T* p, size_t n) {
    if (n) fwrite(p, sizeof(T), n, fp);
}
template <typename T>
inline void vgl_w1(FILE* fp, T v) {
    fwrite(&v, sizeof(T), 1, fp);
}

#ifdef VGL_DUMP_NEED_SIMRECORD
extern int* true_gts_acgt_int;
extern int* n_sim_reads_arr;

inline void vgl_dump_begin(simRecord* sim) {
    (void)sim;
    vgl_dump_state& st = vgl_ds();
    st.has_site_eprob = 0;
    st.site_eprob = -1.0;
    st.stale_base = -1;
    st.r_sample.clear(); st.r_base.clear(); st.r_strand.clear();
    st.r_qs.clear(); st.r_adjqs.clear(); st.r_eprob.clear(); st.tails.clear();
    st.em_sample.clear(); st.em_n.clear(); st.em_codes.clear();
}

inline void vgl_dump_end(simRecord* sim, int ret) {
    FILE* fp = vgl_dump_fp();
    if (!fp) return;
    vgl_dump_state& st = vgl_ds();
    const int S = sim->nSamples;
    vgl_w1<uint32_t>(fp, 0x444c4756u); /* "VGLD" */
    vgl_w1<int32_t>(fp, ret);
    vgl_w1<int64_t>(fp, (int64_t)sim->rec->pos);
    vgl_w1<int32_t>(fp, (int32_t)sim->rec->rid);
    vgl_w1<int32_t>(fp, S);
    vgl_w1<int32_t>(fp, (int32_t)st.r_base.size());
    vgl_w1<int32_t>(fp, (int32_t)st.tails.size());
    vgl_w1<int32_t>(fp, st.stale_base);
    vgl_w1<int32_t>(fp, st.has_site_eprob);
    vgl_w1<double>(fp, st.site_eprob);
    vgl_w1<int32_t>(fp, sim->nAlleles);
    vgl_w1<int32_t>(fp, sim->nAllelesObserved);
    vgl_w1<int32_t>(fp, sim->nGenotypes);
    vgl_w1<int32_t>(fp, sim->allele_unobserved);
    for (int i = 0; i < 5; ++i) vgl_w1<int32_t>(fp, sim->alleles2acgt[i]);
    for (int i = 0; i < 5; ++i) vgl_w1<int32_t>(fp, sim->acgt2alleles[i]);
    const int full = (ret == 0) && (sim->nAlleles > 0);
    const int sizeG = full ? S * sim->nGenotypes : 0;
    const int sizeR = full ? S * sim->nAlleles : 0;
    const int nA = full ? sim->nAlleles : 0;
    uint32_t flags = 0;
    if (full) {
        flags |= 1u << 0; /* gl always exists */
        if (sim->pl_arr) flags |= 1u << 1;
        if (sim->gp_arr) flags |= 1u << 2;
        if (sim->qs_arr) flags |= 1u << 3;
        if (sim->i16_arr) flags |= 1u << 4;
        if (sim->fmt_ad_arr) flags |= 1u << 5;
        if (sim->fmt_adf_arr) flags |= 1u << 6;
        if (sim->fmt_adr_arr) flags |= 1u << 7;
        if (sim->info_ad_arr) flags |= 1u << 8;
        if (sim->info_adf_arr) flags |= 1u << 9;
        if (sim->info_adr_arr) flags |= 1u << 10;
    }
    vgl_w1<int32_t>(fp, sizeG);
    vgl_w1<int32_t>(fp, sizeR);
    vgl_w1<uint32_t>(fp, flags);
    /* inputs + depth draws */
    for (int i = 0; i < 2 * S; ++i) vgl_w1<int8_t>(fp, (int8_t)true_gts_acgt_int[i]);
    for (int i = 0; i < S; ++i) vgl_w1<int32_t>(fp, (int32_t)n_sim_reads_arr[i]);
    vgl_w(fp, sim->fmt_dp_arr, (size_t)S);
    vgl_w1<int32_t>(fp, sim->info_dp_arr[0]);
    /* reads */
    vgl_w(fp, st.r_sample.data(), st.r_sample.size());
    vgl_w(fp, st.r_base.data(), st.r_base.size());
    vgl_w(fp, st.r_strand.data(), st.r_strand.size());
    vgl_w(fp, st.r_qs.data(), st.r_qs.size());
    vgl_w(fp, st.r_adjqs.data(), st.r_adjqs.size());
    vgl_w(fp, st.r_eprob.data(), st.r_eprob.size());
    vgl_w(fp, st.tails.data(), st.tails.size());
    /* depth>255 errmod cells */
    vgl_w1<int32_t>(fp, (int32_t)st.em_sample.size());
    vgl_w(fp, st.em_sample.data(), st.em_sample.size());
    vgl_w(fp, st.em_n.data(), st.em_n.size());
    vgl_w1<int32_t>(fp, (int32_t)st.em_codes.size());
    vgl_w(fp, st.em_codes.data(), st.em_codes.size());
    /* outputs */
    if (flags & (1u << 0)) vgl_w(fp, sim->gl_arr, (size_t)sizeG);
    if (flags & (1u << 1)) vgl_w(fp, sim->pl_arr, (size_t)sizeG);
    if (flags & (1u << 2)) vgl_w(fp, sim->gp_arr, (size_t)sizeG);
    if (flags & (1u << 3)) vgl_w(fp, sim->qs_arr, (size_t)nA);
    if (flags & (1u << 4)) vgl_w(fp, sim->i16_arr, (size_t)16);
    if (flags & (1u << 5)) vgl_w(fp, sim->fmt_ad_arr, (size_t)sizeR);
    if (flags & (1u << 6)) vgl_w(fp, sim->fmt_adf_arr, (size_t)sizeR);
    if (flags & (1u << 7)) vgl_w(fp, sim->fmt_adr_arr, (size_t)sizeR);
    if (flags & (1u << 8)) vgl_w(fp, sim->info_ad_arr, (size_t)nA);
    if (flags & (1u << 9)) vgl_w(fp, sim->info_adf_arr, (size_t)nA);
    if (flags & (1u << 10)) vgl_w(fp, sim->info_adr_arr, (size_t)nA);
    fflush(fp);
}
#endif /* VGL_DUMP_NEED_SIMRECORD */

#endif
