/*
 * oracle/ref_vgl_binding.h -- TEST INFRASTRUCTURE: the reference-side binding of INTEGRATION.md, compiled for real.
 *
 * oracle/build_ref.sh injects this header into a temporary, patched copy of the reference's vcfgl.cpp (never committed) and
 * links the result against libvgl.so as oracle/_ref/vcfgl_ref_vgl: the reference's own main(), argument parser, htslib
 * record handling, add_tags() and writer, with the hot path -- simulate_record_values(), vcfgl.cpp:327-1087 -- answered by
 * libvgl.  The reference's random streams are sequential libc generators that no parallel code can reproduce, so the binding
 * runs in REPLAY mode: the original function still runs first (it also performs the reference's htslib edits of the
 * record), the capture hooks of ref_dump_hooks.h collect its draws, libvgl recomputes the site from those draws on the GPU,
 * and every array add_tags() will hand to htslib is first poisoned and then filled from libvgl's answer (allele string
 * included).  The files this binary writes therefore carry libvgl's numbers: tests/test_gpu_ref_binding.py requires them to
 * equal the unmodified reference's -O u files byte for byte.
 */
#ifndef VGL_REF_BINDING_H
#define VGL_REF_BINDING_H

#include "vgl.h"

#include <string>
#include <vector>

struct vgl_bind_state {
    vgl_ctx* ctx = NULL;
    long n_sites = 0, n_values = 0;
};
inline vgl_bind_state& vgl_bs()
{
    static vgl_bind_state st;
    return st;
}

static void vgl_bind_die(const char* what, int rc)
{
    fprintf(stderr, "[vcfgl_ref_vgl] %s: %s\n", what, vgl_strerror(rc));
    exit(rc == VGL_ENODEV ? 3 : 1);
}

/* the argStruct fields the hot path reads -> vgl_params (INTEGRATION.md "The reference-side change") */
static vgl_params vgl_bind_params(const argStruct* a, int nSamples)
{
    vgl_params p;
    memset(&p, 0, sizeof p);
    p.abi_version = VGL_ABI_VERSION;
    p.n_samples = nSamples;
    p.seed = a->seed;
    p.depth_mode = a->mps_depths ? VGL_DEPTH_POISSON_PER_SAMPLE : VGL_DEPTH_POISSON;
    p.depth_mean = a->mps_depths ? 0.0 : a->mps_depth;
    p.depth_means = a->mps_depths;
    p.error_rate = a->error_rate;
    p.error_qs = a->error_qs;
    p.beta_variance = a->beta_variance;
    p.gl_model = a->GL;
    p.gl1_theta = a->glModel1_theta;
    p.precise_gl = a->usePreciseGlError;
    p.adjust_qs = a->adjustQs;
    p.adjust_by = a->adjustBy;
    p.n_qs_bins = a->n_qs_bins;
    for (int i = 0; i < a->n_qs_bins; ++i) memcpy(p.qs_bins[i], a->qs_bins[i], 3);
    p.do_unobserved = a->doUnobserved;
    p.rm_invar_sites = a->rmInvarSites;
    p.rm_empty_sites = a->rmEmptySites;
    p.do_gvcf = a->doGVCF;
    p.i16_mapq = a->i16_mapq;
    p.tag_mask = (a->addGL ? VGL_TAG_GL : 0) | (a->addGP ? VGL_TAG_GP : 0) | (a->addPL ? VGL_TAG_PL : 0) | (a->addI16 ? VGL_TAG_I16 : 0) |
                 (a->addQS ? VGL_TAG_QS : 0) | (a->addFormatDP ? VGL_TAG_FMT_DP : 0) | (a->addInfoDP ? VGL_TAG_INFO_DP : 0) |
                 (a->addFormatAD ? VGL_TAG_FMT_AD : 0) | (a->addInfoAD ? VGL_TAG_INFO_AD : 0) | (a->addFormatADF ? VGL_TAG_FMT_ADF : 0) |
                 (a->addInfoADF ? VGL_TAG_INFO_ADF : 0) | (a->addFormatADR ? VGL_TAG_FMT_ADR : 0) | (a->addInfoADR ? VGL_TAG_INFO_ADR : 0);
    p.tag_mask |= VGL_TAG_GL | VGL_TAG_FMT_DP; /* the record object always carries gl_arr and fmt_dp_arr (bcf_utils.h:310) */
    p.device_id = 0;
    p.max_batch_sites = 1;
    p.n_slots = 1;
    p.host_output = VGL_HOST_I32;
    return p;
}

static std::string vgl_bind_alleles(const vgl_site_out& s, int do_unobserved, int do_gvcf)
{
    const char* nonref = (do_unobserved == 2 || do_unobserved == 5) ? "<NON_REF>" : "<*>";
    if (s.info_dp == 0) { /* simulate_site_with_no_reads, vcfgl.cpp:228-315 */
        if (do_gvcf) return "<NON_REF>";
        switch (do_unobserved) {
        case 0: return ".";
        case 1: return "<*>";
        case 2: return "<NON_REF>";
        case 3: return "A,C,G,T";
        case 4: return "A,C,G,T,<*>";
        default: return "A,C,G,T,<NON_REF>";
        }
    }
    std::string out;
    for (int a = 0; a < s.n_alleles; ++a) {
        if (a) out += ',';
        const int b = s.alleles2acgt[a];
        if (b == 4) out += nonref;
        else out += "ACGT"[b];
    }
    return out;
}

/* after the original simulate_record_values(): the same site through libvgl (replay of the captured draws); the record's
 * arrays are overwritten with libvgl's */
static void vgl_bind_replace(simRecord* sim, int ret)
{
    if (ret == -1) return; /* input-invariant record (vcfgl.cpp:337): skipped before anything is drawn; never submitted */
    vgl_bind_state& bs = vgl_bs();
    vgl_dump_state& st = vgl_ds();
    const int S = sim->nSamples;
    if (!bs.ctx) {
        vgl_params p = vgl_bind_params(args, S);
        const int rc = vgl_create(&p, &bs.ctx);
        if (rc != VGL_OK) vgl_bind_die("vgl_create", rc);
    }
    uint8_t* gt = NULL;
    int64_t cap = 0;
    int rc = vgl_input_buffer(bs.ctx, 0, &gt, &cap);
    if (rc != VGL_OK) vgl_bind_die("vgl_input_buffer", rc);
    for (int s = 0; s < S; ++s) {
        const int h0 = true_gts_acgt_int[2 * s], h1 = true_gts_acgt_int[2 * s + 1];
        gt[s] = VGL_GT_PACK(h0 < 0 ? VGL_GT_MISSING : h0, h1 < 0 ? VGL_GT_MISSING : h1);
    }
    /* the captured draws in the replay layout (include/vgl.h vgl_replay) */
    std::vector<int32_t> depths((size_t)S);
    std::vector<int64_t> off((size_t)S + 1, 0);
    std::vector<int> per_sample((size_t)S, 0);
    for (size_t i = 0; i < st.r_sample.size(); ++i) per_sample[(size_t)st.r_sample[i]]++;
    for (int s = 0; s < S; ++s) {
        depths[(size_t)s] = n_sim_reads_arr[s];
        off[(size_t)s + 1] = off[(size_t)s] + per_sample[(size_t)s];
    }
    const size_t nr = st.r_base.size();
    std::vector<uint8_t> qs(nr), adjqs(nr), tails(nr, 0);
    for (size_t i = 0; i < nr; ++i) {
        qs[i] = (uint8_t)(st.r_qs[i] < 0 ? 0 : (st.r_qs[i] > 255 ? 255 : st.r_qs[i]));
        adjqs[i] = (uint8_t)(st.r_adjqs[i] < 0 ? 0 : (st.r_adjqs[i] > 255 ? 255 : st.r_adjqs[i]));
    }
    if (st.tails.size() == nr)
        for (size_t i = 0; i < nr; ++i) tails[i] = (uint8_t)st.tails[i];
    std::vector<uint16_t> deep;
    int64_t n_deep = 0;
    if (args->GL == 1) { /* cells deeper than 255 reads: the 255 codes errmod_cal kept (htslib/errmod.c:156-159) */
        size_t o = 0;
        int n_site_deep = 0;
        for (int s = 0; s < S; ++s) n_site_deep += per_sample[(size_t)s] > 255;
        if (st.em_n.empty() && n_site_deep) { /* the site returned before calculate_gls: one (ignored) block per deep cell */
            deep.assign((size_t)n_site_deep * 255, 0);
            n_deep = n_site_deep;
        } else {
            for (size_t k = 0; k < st.em_n.size(); ++k) {
                deep.insert(deep.end(), st.em_codes.begin() + (long)o, st.em_codes.begin() + (long)o + 255);
                o += (size_t)st.em_n[k];
                ++n_deep;
            }
        }
    }
    vgl_replay rp;
    memset(&rp, 0, sizeof rp);
    rp.depths = depths.data();
    rp.read_offsets = off.data();
    rp.n_reads = (int64_t)nr;
    rp.bases = st.r_base.data();
    rp.strands = st.r_strand.data();
    if (args->error_qs == 2) {
        rp.qs = qs.data();
        if (args->adjustQs) rp.adj_qs = adjqs.data();
        rp.error_probs = st.r_eprob.data();
    }
    if (args->addI16) rp.tail_dists = tails.data();
    rp.n_deep_cells = n_deep;
    rp.deep_codes = deep.empty() ? NULL : deep.data();
    rc = vgl_submit(bs.ctx, 0, (int64_t)bs.n_sites, 1, &rp, 0);
    if (rc != VGL_OK) vgl_bind_die(vgl_last_error(bs.ctx), rc);
    vgl_batch_out out;
    rc = vgl_wait(bs.ctx, 0, &out);
    if (rc != VGL_OK) vgl_bind_die("vgl_wait", rc);
    if (out.status != VGL_OK) vgl_bind_die("batch status", out.status);
    ++bs.n_sites;
    const vgl_site_out& so = out.sites[0];
    if (so.skip_code != ret) {
        fprintf(stderr, "[vcfgl_ref_vgl] site %ld: libvgl returns %d, the reference %d\n", bs.n_sites - 1, so.skip_code, ret);
        exit(1);
    }
    if (ret != 0) return;
    /* poison, then fill from libvgl: what add_tags() (bcf_utils.cpp:426-507) hands to htslib is libvgl's */
    const int G = so.n_genotypes, A = so.n_alleles;
    sim->nAlleles = A;
    sim->nAllelesObserved = so.n_alleles_observed;
    sim->nGenotypes = G;
    for (int i = 0; i < 5; ++i) { sim->alleles2acgt[i] = so.alleles2acgt[i]; sim->acgt2alleles[i] = so.acgt2alleles[i]; }
    const std::string al = vgl_bind_alleles(so, args->doUnobserved, args->doGVCF);
    if (bcf_update_alleles_str(sim->hdr, sim->rec, al.c_str()) != 0) vgl_bind_die("bcf_update_alleles_str", VGL_EINVAL);
    const size_t nG = (size_t)S * G, nR = (size_t)S * A;
#define VGL_FILL(dst, src, n, T) do { if ((dst) && (src)) { memset((dst), 0x5A, (n) * sizeof(T)); memcpy((dst), (src), (n) * sizeof(T)); bs.n_values += (long)(n); } } while (0)
    VGL_FILL(sim->fmt_dp_arr, out.dp, (size_t)S, int32_t);
    sim->info_dp_arr[0] = so.info_dp;
    VGL_FILL(sim->gl_arr, out.gl + so.g_off, nG, float);
    if (out.pl) VGL_FILL(sim->pl_arr, out.pl + so.g_off, nG, int32_t);
    if (out.gp) VGL_FILL(sim->gp_arr, out.gp + so.g_off, nG, float);
    if (out.ad) VGL_FILL(sim->fmt_ad_arr, out.ad + so.r_off, nR, int32_t);
    if (out.adf) VGL_FILL(sim->fmt_adf_arr, out.adf + so.r_off, nR, int32_t);
    if (out.adr) VGL_FILL(sim->fmt_adr_arr, out.adr + so.r_off, nR, int32_t);
    if (args->addQS) VGL_FILL(sim->qs_arr, so.qs, (size_t)A, float);
    if (args->addI16) VGL_FILL(sim->i16_arr, so.i16, (size_t)16, float);
    if (args->addInfoAD) VGL_FILL(sim->info_ad_arr, so.info_ad, (size_t)A, int32_t);
    if (args->addInfoADF) VGL_FILL(sim->info_adf_arr, so.info_adf, (size_t)A, int32_t);
    if (args->addInfoADR) VGL_FILL(sim->info_adr_arr, so.info_adr, (size_t)A, int32_t);
#undef VGL_FILL
}

struct vgl_bind_report { /* one line at exit, so that a test can see the hot path really went through libvgl */
    ~vgl_bind_report()
    {
        vgl_bind_state& bs = vgl_bs();
        if (bs.ctx) {
            fprintf(stderr, "[vcfgl_ref_vgl] %ld sites through libvgl (replay), %ld values written into the records, kernels: %s\n", bs.n_sites,
                    bs.n_values, "k_sim+k_site+k_scan+k_emit");
            vgl_destroy(bs.ctx);
        }
    }
};
static vgl_bind_report vgl_bind_report_instance;

#endif
