/*
 * oracle/vgl_oracle.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C, single-threaded CPU restatement of the reference's per-site
 * simulate-and-score arithmetic (isinaltinkaya/vcfgl @ da6a334):
 *   vcfgl.cpp:327-1087 (simulate_record_values, minus the RNG draws),
 *   gl_methods.cpp:4-369 (the five calculate_gls variants),
 *   htslib/errmod.c:51-208 (cal_coef, errmod_cal),
 *   vcfgl.cpp:1661-1743 (preCalc), shared.cpp:110-114 (qs -> log10 GL LUT).
 *
 * It consumes the *draws* (depths, bases, strands, quality scores, error
 * probabilities, tail distances) captured from the instrumented reference
 * (oracle/ref_dump_hooks.h) or produced by the CUDA simulator, and returns
 * the tag arrays exactly as the reference hands them to htslib.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may use
 * anything under oracle/.  The product (vcfgl_b200/, include/) never does.
 *
 * Parity status: PINNED -- tests/test_oracle_golden.py checks this restatement
 * bit-for-bit against tests/golden/<id>.vgld, which are dumps of the reference
 * itself run here on its own 17 hot-path golden tests (whose VCF output is
 * re-verified against test/reference/<id>/<id>.vcf when the dumps are made) plus extra
 * configurations that the reference's tests do not cover.
 */
#ifndef VGL_ORACLE_H
#define VGL_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VGO_MISSING_F32_BITS 0x7F800001u /* htslib/vcf.c:56 bcf_float_missing */
#define VGO_MISSING_I32 INT32_MIN        /* htslib/htslib/vcf.h:1325 */

typedef struct vgo_params {
    int32_t n_samples;
    double error_rate; /* --error-rate, io.h:66 */
    int32_t error_qs;  /* --error-qs 0|1|2, io.h:67 */
    int32_t gl_model;  /* --gl-model 1|2, io.h:69 */
    double gl1_theta;  /* --gl1-theta, io.h:70 */
    int32_t precise_gl; /* --precise-gl, io.h:72 */
    int32_t adjust_qs;  /* --adjust-qs bitmask, shared.h:103-117 */
    double adjust_by;   /* --adjust-by, io.h:76 */
    int32_t n_qs_bins;  /* --qs-bins, io.h:145-146 */
    uint8_t qs_bins[255][3];
    int32_t do_unobserved;  /* -doUnobserved 0..5, shared.h:70-89 */
    int32_t rm_invar_sites; /* --rm-invar-sites bitmask, shared.h:119-125 */
    int32_t rm_empty_sites; /* --rm-empty-sites */
    int32_t do_gvcf;
    int32_t i16_mapq; /* --i16-mapq */
    int32_t add_gl, add_gp, add_pl, add_i16, add_qs;
    int32_t add_fmt_dp, add_info_dp;
    int32_t add_fmt_ad, add_info_ad, add_fmt_adf, add_info_adf, add_fmt_adr, add_info_adr;
} vgo_params;

typedef struct vgo_ctx vgo_ctx;

/* draws of one site (replay input) */
typedef struct vgo_site_in {
    const int8_t* gts;      /* [2*S] true alleles as ACGT ints, -1 missing (vcfgl.cpp:133-146) */
    const int32_t* depths;  /* [S] drawn depths incl. missing-GT samples (vcfgl.cpp:364-368) */
    int32_t n_reads;        /* reads of all non-missing samples, in (sample, read) order */
    const uint8_t* bases;   /* [n_reads] observed base 0..3 */
    const uint8_t* strands; /* [n_reads] 0 fwd / 1 rev */
    const int32_t* qs;      /* [n_reads] raw qs (error_qs 2 only, else ignored) */
    const int32_t* adj_qs;  /* [n_reads] adjusted qs (error_qs 2 and adjust_qs!=0) */
    const double* eprob;    /* [n_reads] beta-drawn error prob (error_qs 2) */
    int32_t n_tails;        /* == n_reads when add_i16, else 0 */
    const int32_t* tails;   /* [n_tails] capped tail distances (vcfgl.cpp:653-656) */
    int32_t n_em;           /* cells with depth>255 under GL model 1 */
    const int32_t* em_sample; /* [n_em] */
    const int32_t* em_n;      /* [n_em] */
    const uint16_t* em_codes; /* concatenated post-shuffle read codes (errmod.c:156-159) */
} vgo_site_in;

/* outputs of one site; arrays are caller-allocated at maximum size */
typedef struct vgo_site_out {
    int32_t ret; /* 0, -3 (simulated invariant), -4 (empty site) */
    int32_t n_alleles, n_alleles_observed, n_genotypes, allele_unobserved;
    int32_t alleles2acgt[5], acgt2alleles[5];
    int32_t info_dp;
    int32_t* fmt_dp;                       /* [S] */
    float* gl;                             /* [S*15] */
    int32_t* pl;                           /* [S*15] */
    float* gp;                             /* [S*15] */
    int32_t *fmt_ad, *fmt_adf, *fmt_adr;   /* [S*5] */
    int32_t info_ad[5], info_adf[5], info_adr[5];
    float qs[5];
    float i16[16];
} vgo_site_out;

vgo_ctx* vgo_create(const vgo_params* p);
void vgo_destroy(vgo_ctx* c);

/* derived constants, vcfgl.cpp:1661-1743 (error_qs 0/1 only; -1 otherwise) */
int vgo_precalc_qs(const vgo_ctx* c);
int vgo_precalc_adj_qs(const vgo_ctx* c);
void vgo_precalc_gl2(const vgo_ctx* c, double out3[3]);

/* one site: draws -> tags.  Returns out->ret. */
int vgo_site(vgo_ctx* c, const vgo_site_in* in, vgo_site_out* out);

/* pieces exposed for table-level tests */
const double* vgo_errmod_fk(const vgo_ctx* c);   /* [256] */
const double* vgo_errmod_beta(const vgo_ctx* c); /* [64*256*256] */
const double* vgo_errmod_lhet(const vgo_ctx* c); /* [256*256] */
void vgo_errmod_cal(const vgo_ctx* c, int n, const uint16_t* codes, float q[25]);
const double* vgo_lut_log10_gl(void); /* [3*257] */

#ifdef __cplusplus
}
#endif
#endif
