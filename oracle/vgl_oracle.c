/*
 * oracle/vgl_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 * See oracle/vgl_oracle.h for scope, citations and parity status (PINNED).
 *
 * Build: gcc -O2 -ffp-contract=off -fPIC -shared (no -march=native, no
 * -ffast-math): the reference is built with plain -O3 on x86-64, i.e. SSE2
 * scalar IEEE arithmetic without FMA contraction, and every float/double
 * mixing below is deliberate.
 */
#include "vgl_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

struct vgo_ctx {
    vgo_params p;
    /* errmod tables, htslib/errmod.c:36-40 */
    double* fk;
    double* beta;
    double* lhet;
    /* preCalc, io.h:22-32 */
    int pre_qs, pre_adj_qs;
    double pre_gl2[3]; /* homT, het, homF */
    /* per-site scratch in ACGT space (simRecord, bcf_utils.h:146-165) */
    int32_t *acgt_ad, *acgt_adf, *acgt_adr, *acgt_qsum, *acgt_qsumsq;
    int error;
};

/* ------------------------------------------------------------------------ */
/* shared.cpp:110-114 -- qScore_to_log10_gl[3][257].  The reference table is  */
/* R output (7 significant digits) of the formulas in shared.h:516-527; we   */
/* regenerate it from those formulas and round through "%.7g".  Equality of   */
/* all 771 doubles with the reference's table is asserted by                 */
/* tests/test_oracle_golden.py::test_lut_matches_reference (container only). */
static double g_lut[3 * 257];
static int g_lut_ready = 0;

static double round7(double v)
{
    char buf[64];
    if (isinf(v)) return v;
    snprintf(buf, sizeof buf, "%.7g", v);
    return strtod(buf, NULL);
}

const double* vgo_lut_log10_gl(void)
{
    if (!g_lut_ready) {
        for (int q = 0; q <= 256; ++q) {
            double p = pow(10.0, -q / 10.0);
            double hom_hit = (p < 1.0) ? log10(1.0 - p) : -INFINITY;
            double het_hit = log10((1.0 - p) / 2.0 + p / 6.0);
            double hom_miss = log10(p) - log10(3.0);
            g_lut[0 * 257 + q] = round7(hom_hit);
            g_lut[1 * 257 + q] = round7(het_hit);
            g_lut[2 * 257 + q] = round7(hom_miss);
        }
        g_lut_ready = 1;
    }
    return g_lut;
}

/* ------------------------------------------------------------------------ */
/* htslib/errmod.c:51-112 (logbinomial_table + cal_coef), depcorr = 1-theta  */
/* (io.cpp:1276), eta = 0.03 (errmod.c:123).                                 */
static int errmod_tables(vgo_ctx* c, double depcorr, double eta)
{
    c->fk = (double*)calloc(256, sizeof(double));
    c->beta = (double*)calloc((size_t)64 * 256 * 256, sizeof(double));
    c->lhet = (double*)calloc(256 * 256, sizeof(double));
    double* lc = (double*)calloc(256 * 256, sizeof(double));
    if (!c->fk || !c->beta || !c->lhet || !lc) return -1;

    /* log C(n,k) for 1<=k<=n<256, zero elsewhere (errmod.c:58-62) */
    for (int n = 1; n < 256; ++n) {
        double lfn = lgamma(n + 1);
        for (int k = 1; k <= n; ++k) lc[n << 8 | k] = lfn - lgamma(k + 1) - lgamma(n - k + 1);
    }
    /* dependency coefficients (errmod.c:75-77) */
    c->fk[0] = 1.0;
    for (int n = 1; n < 256; ++n) c->fk[n] = pow(1. - depcorr, n) * (1.0 - eta) + eta;
    /* beta[q][n][k] = phred-scaled ratio of binomial tail sums (errmod.c:86-99) */
    for (int q = 1; q < 64; ++q) {
        double e = pow(10.0, -q / 10.0);
        double le = log(e);
        double le1 = log(1.0 - e);
        for (int n = 1; n <= 255; ++n) {
            double* b = c->beta + ((size_t)q << 16 | n << 8);
            double tail = lc[n << 8 | n] + n * le; /* log P(K = n) */
            b[n] = HUGE_VAL;
            for (int k = n - 1; k >= 0; --k) {
                double tail_k = tail + log1p(exp(lc[n << 8 | k] + k * le + (n - k) * le1 - tail));
                b[k] = -10. / M_LN10 * (tail - tail_k);
                tail = tail_k;
            }
        }
    }
    /* lhet[n][k] = log C(n,k) - n ln 2 (errmod.c:107-109) */
    for (int n = 0; n < 256; ++n)
        for (int k = 0; k < 256; ++k) c->lhet[n << 8 | k] = lc[n << 8 | k] - M_LN2 * n;
    free(lc);
    return 0;
}

static int cmp_u16(const void* a, const void* b)
{
    uint16_t x = *(const uint16_t*)a, y = *(const uint16_t*)b;
    return (x > y) - (x < y);
}

/* htslib/errmod.c:143-208 with m = 5 (gl_methods.cpp:266,333).  n <= 255:   */
/* the caller passes the post-shuffle, truncated reads when depth > 255.     */
void vgo_errmod_cal(const vgo_ctx* c, int n, const uint16_t* codes_in, float q[25])
{
    enum { M = 5 };
    memset(q, 0, 25 * sizeof(float));
    if (n == 0) return;
    uint16_t codes[255];
    if (n > 255) n = 255;
    memcpy(codes, codes_in, (size_t)n * sizeof(uint16_t));
    qsort(codes, (size_t)n, sizeof(uint16_t), cmp_u16); /* ascending; walked descending */

    double bsum[16];
    uint32_t cnt[16];
    int per_strand[32];
    memset(bsum, 0, sizeof bsum);
    memset(cnt, 0, sizeof cnt);
    memset(per_strand, 0, sizeof per_strand);
    for (int j = n - 1; j >= 0; --j) {
        const uint16_t code = codes[j];
        int qual = code >> 5;
        if (qual < 4) qual = 4;
        if (qual > 63) qual = 63;
        const int bs = code & 0x1f, base = code & 0xf;
        const double f = c->fk[per_strand[bs]];
        bsum[base] += f * c->beta[(size_t)qual << 16 | n << 8 | cnt[base]];
        ++cnt[base];
        ++per_strand[bs];
    }
    for (int j = 0; j < M; ++j) {
        /* homozygous jj: everything that is not j is an error (float accumulator, errmod.c:182-191) */
        float acc = 0.0f;
        int others = 0;
        for (int k = 0; k < M; ++k) {
            if (k == j) continue;
            acc += bsum[k];
            others += cnt[k];
        }
        if (others) q[j * M + j] = acc;
        /* heterozygous jk (errmod.c:193-202) */
        for (int k = j + 1; k < M; ++k) {
            const int cjk = cnt[j] + cnt[k];
            acc = 0.0f;
            others = 0;
            for (int i = 0; i < M; ++i) {
                if (i == j || i == k) continue;
                acc += bsum[i];
                others += cnt[i];
            }
            if (others)
                q[j * M + k] = q[k * M + j] = -4.343 * c->lhet[cjk << 8 | cnt[k]] + acc;
            else
                q[j * M + k] = q[k * M + j] = -4.343 * c->lhet[cjk << 8 | cnt[k]];
        }
        for (int k = 0; k < M; ++k)
            if (q[j * M + k] < 0.0) q[j * M + k] = 0.0;
    }
}

/* ------------------------------------------------------------------------ */
/* vcfgl.cpp:57-64 */
static int apply_bins(vgo_ctx* c, int qs)
{
    for (int i = 0; i < c->p.n_qs_bins; ++i)
        if (qs >= c->p.qs_bins[i][0] && qs <= c->p.qs_bins[i][1]) return c->p.qs_bins[i][2];
    c->error = 1;
    return -1;
}

/* shared.h:459 QS_TO_QSSQ */
static int qs_sq(int q) { return q == 0 ? 0 : (q < 63 ? q * q : 3969); }

/* vcfgl.cpp:1661-1743 */
static void precalc(vgo_ctx* c)
{
    const vgo_params* p = &c->p;
    c->pre_qs = c->pre_adj_qs = -1;
    c->pre_gl2[0] = c->pre_gl2[1] = c->pre_gl2[2] = -1.0;
    if (p->error_qs == 2) return;
    const double e = p->error_rate;
    int qs = -1, adj = -1;
    if (e == 0.0) {
        qs = adj = 63;
    } else if (e == 1.0) {
        qs = adj = 0;
    } else {
        double t = -10.0 * log10(e);
        qs = (int)t;
        if (p->adjust_qs) adj = (int)(t + p->adjust_by);
    }
    if (p->n_qs_bins) {
        qs = apply_bins(c, qs);
        if (p->adjust_qs) adj = apply_bins(c, adj);
    } else {
        if (qs > 63) qs = 63;
        if (p->adjust_qs && adj > 63) adj = 63;
    }
    c->pre_qs = qs;
    if (p->adjust_qs) c->pre_adj_qs = adj;
    if (p->gl_model == 2) {
        if (!p->precise_gl) {
            const int q = (p->adjust_qs & 1) ? c->pre_adj_qs : c->pre_qs;
            const double* lut = vgo_lut_log10_gl();
            c->pre_gl2[0] = lut[0 * 257 + q];
            c->pre_gl2[1] = lut[1 * 257 + q];
            c->pre_gl2[2] = lut[2 * 257 + q];
        } else if (e == 0.0) {
            c->pre_gl2[0] = 0;
            c->pre_gl2[1] = -0.3010299956639812;
            c->pre_gl2[2] = -INFINITY;
        } else {
            c->pre_gl2[0] = log10(1.0 - e);
            c->pre_gl2[1] = log10((1.0 - e) / 2.0 + e / 6.0);
            c->pre_gl2[2] = log10(e) - 0.47712125471966244;
        }
    }
}

vgo_ctx* vgo_create(const vgo_params* p)
{
    vgo_ctx* c = (vgo_ctx*)calloc(1, sizeof(vgo_ctx));
    if (!c) return NULL;
    c->p = *p;
    if (p->gl_model == 1 && errmod_tables(c, 1.0 - p->gl1_theta, 0.03) != 0) {
        vgo_destroy(c);
        return NULL;
    }
    precalc(c);
    const size_t n4 = (size_t)4 * p->n_samples;
    c->acgt_ad = (int32_t*)calloc(n4, sizeof(int32_t));
    c->acgt_adf = (int32_t*)calloc(n4, sizeof(int32_t));
    c->acgt_adr = (int32_t*)calloc(n4, sizeof(int32_t));
    c->acgt_qsum = (int32_t*)calloc(n4, sizeof(int32_t));
    c->acgt_qsumsq = (int32_t*)calloc(n4, sizeof(int32_t));
    return c;
}

void vgo_destroy(vgo_ctx* c)
{
    if (!c) return;
    free(c->fk); free(c->beta); free(c->lhet);
    free(c->acgt_ad); free(c->acgt_adf); free(c->acgt_adr); free(c->acgt_qsum); free(c->acgt_qsumsq);
    free(c);
}

int vgo_precalc_qs(const vgo_ctx* c) { return c->pre_qs; }
int vgo_precalc_adj_qs(const vgo_ctx* c) { return c->pre_adj_qs; }
void vgo_precalc_gl2(const vgo_ctx* c, double out3[3]) { memcpy(out3, c->pre_gl2, sizeof c->pre_gl2); }
const double* vgo_errmod_fk(const vgo_ctx* c) { return c->fk; }
const double* vgo_errmod_beta(const vgo_ctx* c) { return c->beta; }
const double* vgo_errmod_lhet(const vgo_ctx* c) { return c->lhet; }

static float missing_f32(void)
{
    union { uint32_t u; float f; } m;
    m.u = VGO_MISSING_F32_BITS;
    return m.f;
}

static int gt_index(int a, int b) /* htslib/htslib/vcf.h:902 bcf_alleles2gt */
{
    return a > b ? a * (a + 1) / 2 + b : b * (b + 1) / 2 + a;
}

/* ------------------------------------------------------------------------ */
int vgo_site(vgo_ctx* c, const vgo_site_in* in, vgo_site_out* o)
{
    const vgo_params* p = &c->p;
    const int S = p->n_samples;
    const int will_explode = (p->do_unobserved >= 3);
    const int will_add_unobs = (p->do_unobserved == 1 || p->do_unobserved == 2 || p->do_unobserved == 4 || p->do_unobserved == 5);
    const int sample_strand = p->add_i16 || p->add_fmt_adf || p->add_fmt_adr || p->add_info_adf || p->add_info_adr; /* shared.h:160 */
    const int want_adf = p->add_fmt_adf || p->add_info_adf; /* vcfgl.cpp:460 */
    const int want_adr = p->add_fmt_adr || p->add_info_adr; /* vcfgl.cpp:463 */
    const int have_fmt_ad = p->add_fmt_ad || p->add_i16;     /* bcf_utils.cpp:236 */
    const float fmiss = missing_f32();

    /* reset_rec_objects, bcf_utils.h:230-394 */
    memset(c->acgt_ad, 0, sizeof(int32_t) * 4 * S);
    memset(c->acgt_adf, 0, sizeof(int32_t) * 4 * S);
    memset(c->acgt_adr, 0, sizeof(int32_t) * 4 * S);
    memset(c->acgt_qsum, 0, sizeof(int32_t) * 4 * S);
    memset(c->acgt_qsumsq, 0, sizeof(int32_t) * 4 * S);
    int32_t info_acgt_ad[4] = { 0, 0, 0, 0 };
    int n_bases_i16[8] = { 0 };
    float tail_sum[4] = { 0 }, tail_sumsq[4] = { 0 };
    for (int i = 0; i < 5; ++i) {
        o->alleles2acgt[i] = o->acgt2alleles[i] = -1;
        o->info_ad[i] = o->info_adf[i] = o->info_adr[i] = 0;
        o->qs[i] = 0.0f;
    }
    for (int i = 0; i < 16; ++i) o->i16[i] = 0.0f;
    o->n_alleles = o->n_alleles_observed = o->n_genotypes = 0;
    o->allele_unobserved = -1;
    o->info_dp = 0;
    for (int i = 0; i < S * 15; ++i) {
        o->gl[i] = -0.0f;
        o->gp[i] = 0.0f;
        o->pl[i] = 255;
    }
    for (int i = 0; i < S * 5; ++i) o->fmt_ad[i] = o->fmt_adf[i] = o->fmt_adr[i] = 0;

    /* depths, vcfgl.cpp:371-389: a missing GT discards the drawn depth */
    for (int s = 0; s < S; ++s) {
        if (in->gts[2 * s] == -1 || in->gts[2 * s + 1] == -1) {
            o->fmt_dp[s] = 0;
            continue;
        }
        o->fmt_dp[s] = in->depths[s];
        o->info_dp += in->depths[s];
    }

    if (o->info_dp == 0) {
        /* vcfgl.cpp:396-404 and simulate_site_with_no_reads :228-315 */
        if (p->rm_empty_sites) return (o->ret = -4);
        if (p->do_gvcf) return (o->ret = 0); /* tags added as they are; host formats */
        switch (p->do_unobserved) {
        case 0: case 1: case 2:
            o->n_alleles = 1; o->n_genotypes = 1; o->n_alleles_observed = 0; break;
        case 3:
            o->n_alleles = 4; o->n_genotypes = 10; o->n_alleles_observed = 4; break;
        default:
            o->n_alleles = 5; o->n_genotypes = 15; o->n_alleles_observed = 4; break;
        }
        for (int i = 0; i < S * 15; ++i) {
            o->pl[i] = VGO_MISSING_I32;
            o->gp[i] = fmiss;
            o->gl[i] = fmiss;
        }
        return (o->ret = 0);
    }

    /* read loop, vcfgl.cpp:441-640 (draws replayed) */
    int r = 0; /* cursor into the site's reads */
    const int qsum_adj = (p->adjust_qs & 2) != 0; /* shared.h:184 */
    for (int s = 0; s < S; ++s) {
        const int n = o->fmt_dp[s];
        if (n == 0) continue;
        int32_t* ad = c->acgt_ad + 4 * s;
        int32_t* qsum = c->acgt_qsum + 4 * s;
        int32_t* qsumsq = c->acgt_qsumsq + 4 * s;
        for (int i = 0; i < n; ++i, ++r) {
            const int b = in->bases[r];
            int q_for_sum;
            if (p->error_qs == 2)
                q_for_sum = qsum_adj ? in->adj_qs[r] : in->qs[r];
            else
                q_for_sum = qsum_adj ? c->pre_adj_qs : c->pre_qs;
            qsum[b] += q_for_sum;
            qsumsq[b] += qs_sq(q_for_sum);
            ad[b]++;
            int strand = 0;
            if (sample_strand) {
                strand = in->strands[r];
                if (strand == 0) { if (want_adf) c->acgt_adf[4 * s + b]++; }
                else             { if (want_adr) c->acgt_adr[4 * s + b]++; }
            } else if (want_adf) {
                c->acgt_adf[4 * s + b]++;
            }
            n_bases_i16[2 * b + strand]++;
        }
        for (int b = 0; b < 4; ++b) info_acgt_ad[b] += ad[b];
    }
    if (r != in->n_reads) c->error = 2;

    /* tail distances, vcfgl.cpp:647-663: all mass goes to the last simulated base */
    if (p->add_i16) {
        const int stale = in->bases[in->n_reads - 1];
        for (int i = 0; i < in->n_tails; ++i) {
            const int t = in->tails[i];
            tail_sum[stale] += t;
            tail_sumsq[stale] += (t * t);
        }
    }

    int n_obs = 0;
    for (int b = 0; b < 4; ++b) n_obs += info_acgt_ad[b] > 0;
    if ((p->rm_invar_sites & 4) && n_obs == 1) return (o->ret = -3); /* vcfgl.cpp:675-681 */

    /* allele order: stable descending sort of ACGT by INFO/AD, vcfgl.cpp:700-718 */
    for (int b = 0; b < 4; ++b) {
        int rank = 0;
        for (int x = 0; x < 4; ++x)
            if (info_acgt_ad[x] > info_acgt_ad[b] || (info_acgt_ad[x] == info_acgt_ad[b] && x < b)) ++rank;
        o->acgt2alleles[b] = rank;
        o->alleles2acgt[rank] = b;
    }
    /* unobserved bases, vcfgl.cpp:722-762 */
    if (!will_explode)
        for (int b = 0; b < 4; ++b)
            if (info_acgt_ad[b] == 0) {
                o->alleles2acgt[o->acgt2alleles[b]] = -1;
                o->acgt2alleles[b] = -1;
            }
    int n_alleles = 0;
    while (n_alleles < 5 && o->alleles2acgt[n_alleles] != -1) ++n_alleles;
    int n_unobs = 0;
    if (will_add_unobs) {
        o->allele_unobserved = n_alleles;
        o->alleles2acgt[n_alleles] = 4;
        o->acgt2alleles[4] = n_alleles;
        n_unobs = 1;
    }
    o->n_alleles_observed = n_alleles;
    o->n_alleles = n_alleles + n_unobs;
    o->n_genotypes = o->n_alleles * (o->n_alleles + 1) / 2;
    const int A = o->n_alleles, G = o->n_genotypes;

    /* ---- calculate_gls, gl_methods.cpp ---- */
    const int gl_adj = (p->adjust_qs & 1) != 0; /* shared.h:181 */
    const double* lut = vgo_lut_log10_gl();
    r = 0;
    int em_i = 0;
    size_t em_off = 0;
    for (int s = 0; s < S; ++s) {
        const int n = o->fmt_dp[s];
        float* gl = o->gl + (size_t)s * G;
        if (n == 0) {
            for (int g = 0; g < G; ++g) gl[g] = fmiss;
            continue;
        }
        if (p->gl_model == 2) {
            for (int i = 0; i < n; ++i, ++r) {
                double c3[3]; /* [#alleles of the genotype equal to the observed one] */
                if (p->error_qs != 2) { /* gl_methods.cpp:4-69 */
                    c3[2] = c->pre_gl2[0]; c3[1] = c->pre_gl2[1]; c3[0] = c->pre_gl2[2];
                } else if (!p->precise_gl) { /* gl_methods.cpp:71-150 */
                    const int q = gl_adj ? in->adj_qs[r] : in->qs[r];
                    c3[2] = lut[q]; c3[1] = lut[257 + q]; c3[0] = lut[514 + q];
                } else { /* gl_methods.cpp:152-231 */
                    const double e = in->eprob[r];
                    if (e == 0.0) {
                        c3[2] = 0.0; c3[1] = -0.30103; c3[0] = -INFINITY;
                    } else {
                        c3[2] = log10(1.0 - e);
                        c3[1] = log10((1.0 - e) / 2.0 + e / 6.0);
                        c3[0] = log10(e / 3.0);
                    }
                }
                const int ao = o->acgt2alleles[in->bases[r]];
                for (int a2 = 0; a2 < A; ++a2)
                    for (int a1 = 0; a1 <= a2; ++a1)
                        gl[gt_index(a1, a2)] += c3[(a1 == ao) + (a2 == ao)];
                float mx = -INFINITY; /* rescaled after every read, gl_methods.cpp:50-58 */
                for (int g = 0; g < G; ++g)
                    if (gl[g] > mx) mx = gl[g];
                for (int g = 0; g < G; ++g) gl[g] -= mx;
            }
        } else {
            /* gl_methods.cpp:233-369 */
            uint16_t* codes = (uint16_t*)malloc(sizeof(uint16_t) * (size_t)n);
            for (int i = 0; i < n; ++i, ++r) {
                int q;
                if (p->error_qs == 2) q = gl_adj ? in->adj_qs[r] : in->qs[r];
                else q = gl_adj ? c->pre_adj_qs : c->pre_qs;
                codes[i] = (uint16_t)(q << 5 | in->bases[r]);
            }
            float fpl[25];
            if (n > 255) {
                /* errmod.c:156-159: the kept reads come from the capture */
                if (em_i >= in->n_em || in->em_sample[em_i] != s || in->em_n[em_i] != n) {
                    c->error = 3;
                    memset(fpl, 0, sizeof fpl);
                } else {
                    vgo_errmod_cal(c, 255, in->em_codes + em_off, fpl);
                    em_off += (size_t)n;
                    ++em_i;
                }
            } else {
                vgo_errmod_cal(c, n, codes, fpl);
            }
            free(codes);
            float mx = -INFINITY;
            int g = 0;
            for (int a2 = 0; a2 < A; ++a2) {
                const int b2 = o->alleles2acgt[a2];
                for (int a1 = 0; a1 <= a2; ++a1, ++g) {
                    const int b1 = o->alleles2acgt[a1];
                    gl[g] = ((-1.0 * fpl[b1 * 5 + b2]) / 10.0);
                    if (gl[g] > mx) mx = gl[g];
                }
            }
            for (int i = 0; i < g; ++i) gl[i] -= mx;
        }
    }

    /* AD/ADF/ADR into allele order, vcfgl.cpp:806-843 */
    for (int s = 0; s < S; ++s)
        for (int a = 0; a < A; ++a) {
            const int b = o->alleles2acgt[a];
            if (b == 4) continue;
            if (have_fmt_ad) o->fmt_ad[s * A + a] = c->acgt_ad[4 * s + b];
            if (p->add_fmt_adf) o->fmt_adf[s * A + a] = c->acgt_adf[4 * s + b];
            if (p->add_fmt_adr) o->fmt_adr[s * A + a] = c->acgt_adr[4 * s + b];
            if (p->add_info_ad) o->info_ad[a] += c->acgt_ad[4 * s + b];
            if (p->add_info_adf) o->info_adf[a] += c->acgt_adf[4 * s + b];
            if (p->add_info_adr) o->info_adr[a] += c->acgt_adr[4 * s + b];
        }

    /* QS, vcfgl.cpp:845-898 */
    if (p->add_qs)
        for (int s = 0; s < S; ++s) {
            float sum = 0.0;
            const int32_t* qsum = c->acgt_qsum + 4 * s;
            for (int b = 0; b < 4; ++b) sum += qsum[b];
            if (sum != 0.0)
                for (int b = 0; b < 4; ++b) {
                    const int a = o->acgt2alleles[b];
                    if (a == -1) continue;
                    o->qs[a] += (float)((float)(qsum[b]) / sum);
                }
        }

    /* PL, vcfgl.cpp:907-939 */
    if (p->add_pl)
        for (int i = 0; i < S * G; ++i) {
            uint32_t bits;
            memcpy(&bits, &o->gl[i], 4);
            if (bits == VGO_MISSING_F32_BITS) {
                o->pl[i] = VGO_MISSING_I32;
            } else if (o->gl[i] == -INFINITY) {
                o->pl[i] = 255;
            } else {
                int x = (int)lroundf(-10.0 * o->gl[i]);
                o->pl[i] = x > 255 ? 255 : x;
            }
        }

    /* GP, vcfgl.cpp:941-970 */
    if (p->add_gp) {
        for (int i = 0; i < S * G; ++i) {
            uint32_t bits;
            memcpy(&bits, &o->gl[i], 4);
            o->gp[i] = (bits == VGO_MISSING_F32_BITS) ? fmiss : (float)pow(10, o->gl[i]);
        }
        for (int s = 0; s < S; ++s) {
            float* gp = o->gp + (size_t)s * G;
            uint32_t bits;
            memcpy(&bits, &gp[0], 4);
            int miss = 0;
            float sum = 0.0;
            for (int g = 0; g < G; ++g) {
                memcpy(&bits, &gp[g], 4);
                if (bits == VGO_MISSING_F32_BITS) { miss = 1; break; }
                sum += gp[g];
            }
            if (miss) continue;
            for (int g = 0; g < G; ++g) gp[g] /= sum;
        }
    }

    /* I16, vcfgl.cpp:982-1074 */
    if (p->add_i16) {
        float* v = o->i16;
        const int refb = o->alleles2acgt[0];
        v[0] = n_bases_i16[refb * 2 + 0];
        v[1] = n_bases_i16[refb * 2 + 1];
        const int mq = p->i16_mapq;
        for (int s = 0; s < S; ++s) {
            v[4] += c->acgt_qsum[s * 4 + refb];
            v[5] += c->acgt_qsumsq[s * 4 + refb];
            for (int a = 0; a < A; ++a) {
                if (a == o->n_alleles_observed) continue;
                const int k = o->fmt_ad[s * A + a];
                for (int i = 0; i < k; ++i) {
                    if (a == 0) { v[8] += mq; v[9] += mq * mq; }
                    else        { v[10] += mq; v[11] += mq * mq; }
                }
            }
        }
        v[12] = tail_sum[refb];
        v[13] = tail_sumsq[refb];
        for (int a = 1; a < A; ++a) {
            if (a == o->n_alleles_observed) continue;
            const int b = o->alleles2acgt[a];
            v[2] += n_bases_i16[b * 2 + 0];
            v[3] += n_bases_i16[b * 2 + 1];
            for (int s = 0; s < S; ++s) {
                v[6] += c->acgt_qsum[s * 4 + b];
                v[7] += c->acgt_qsumsq[s * 4 + b];
            }
            v[14] += tail_sum[b];
            v[15] += tail_sumsq[b];
        }
    }
    if (c->error) return (o->ret = -100 - c->error);
    return (o->ret = 0);
}
