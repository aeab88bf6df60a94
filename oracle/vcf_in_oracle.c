/*
 * oracle/vcf_in_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C, byte-at-a-time restatement of what the reference does to turn one VCF text record into
 * the true genotypes its hot path consumes:
 *
 *   htslib/vcf.c:3041-3110   vcf_parse: tab-separated columns, POS (hts_str2uint, 1-based -> 0-based),
 *                            REF, ALT split on ',' (n_allele = 1 + number of ALT entries, ALT "." = none)
 *   htslib/vcf.c:2425-2520   vcf_parse_format: FORMAT keys split on ':', the GT key's index
 *   htslib/vcf.c:2643-2673   the GT sub-field: '.' or hts_str2uint (htslib/textutils_internal.h:273-305,
 *                            optional '+'), separators '/' and '|', anything else ends the vector
 *   htslib/vcf.c:2726-2738   the character after a sub-field must be ':' or the end of the column
 *   htslib/vcf.c:2740-2760, 2668   omitted GT -> missing + vector_end padding
 *   vcfgl.cpp:75-163         check_rec_alleles: allele strings -> ACGT ints (--source 0: first char '0'/'1',
 *                            at most 2 alleles; --source 1: allele_char_to_int vcfgl.cpp:20-52, at most 5),
 *                            gt_arr[2*s + h] read as DIPLOID (SIM_PLOIDY shared.h:235), missing -> -1,
 *                            allele index asserted < n_allele, allelesum, skip codes -1 / -2
 *                            (--rm-invar-sites bits 1 / 2, shared.h:119-125)
 *
 * Where the reference would exit (ERROR / ASSERT) or read out of bounds, this restatement reports a
 * status code instead; when a line has several defects the SMALLEST code wins (so the result does not
 * depend on the order in which columns are looked at).  The codes are those of include/vgl.h
 * (vgl_in_status), restated here so that the oracle does not include product headers' logic.
 *
 * Output genotype byte per sample: low nibble = first haplotype, high nibble = second, value = ACGT
 * int 0..3 (true_gts_acgt_int, vcfgl.cpp:133-146), 0xF = missing (-1 there).
 *
 * Parity status: PINNED -- tests/test_vcfin_oracle.py parses the input VCFs of the golden cases
 * (tests/golden/inputs/, copies of the reference's test/data files and of the synthetic inputs) and
 * must reproduce, record for record, the (pos, true_gts_acgt_int) sequence the instrumented reference
 * dumped for the same input (tests/golden/<id>.vgld.gz, including -explode 1 and --source 1 runs and the
 * --rm-invar-sites 1/2/3 captures in tests/golden/inputs/in_*.json).
 */
#include <stdint.h>
#include <stddef.h>
#include <string.h>

enum {
    VIN_OK = 0,
    VIN_ENCOLS = 1,      /* fewer than 10 columns (no FORMAT / no sample column) */
    VIN_EPOS = 2,        /* POS does not fit (reference: "Position value too large" / 64-bit positions) */
    VIN_ENALLELE = 3,    /* more than 5 alleles (vcfgl.cpp:90-92), or more than 2 with --source 0 (vcfgl.cpp:123-125) */
    VIN_EALLELE = 4,     /* allele is not a valid base (vcfgl.cpp:99-101) / not 0 or 1 (vcfgl.cpp:113-115) */
    VIN_ENOGT = 5,       /* no GT key in FORMAT (vcfgl.cpp:83-85) */
    VIN_ENSAMPLES = 6,   /* fewer sample columns than the header has samples (htslib/vcf.c:2777-2783) */
    VIN_EGTCHAR = 7,     /* GT value not a number or '.', or an invalid character after it (htslib/vcf.c:2666-2669, 2729-2737) */
    VIN_EPLOIDY = 8,     /* a sample is not diploid: the reference indexes gt_arr as [2*S] and asserts on vector_end */
    VIN_EALLELEIDX = 9,  /* GT allele index >= n_allele (vcfgl.cpp:144) */
    VIN_ESYMBOLIC = 10   /* GT points at <*> / <NON_REF> (allele_char_to_int = 4: not a base the simulator can draw) */
};

typedef struct vin_site {
    int32_t status;
    int32_t skip_code;
    int64_t pos;
    int64_t allele_sum;
    uint64_t line_off;
    uint32_t line_len;
    int32_t n_allele;
    int8_t allele_acgt[8];
    uint32_t id_off, fmt_off, samples_off; /* relative to line_off */
    uint32_t _pad;
} vin_site;

static void raise_(int* st, int code)
{
    if (*st == VIN_OK || code < *st) *st = code;
}

/* vcfgl.cpp:20-52 on a length-delimited string */
static int allele_to_acgt(const char* a, size_t n)
{
    if (n == 1) {
        switch (a[0]) {
        case 'A': return 0;
        case 'C': return 1;
        case 'G': return 2;
        case 'T': return 3;
        default: return -1;
        }
    }
    if (n == 3 && !memcmp(a, "<*>", 3)) return 4;
    if (n == 9 && !memcmp(a, "<NON_REF>", 9)) return 4;
    return -1;
}

/* one record line [line, line+len) without its newline; gt_row receives S bytes */
void vin_oracle_line(const char* line, size_t len, int32_t S, int32_t gt_source, int32_t rm_invar, vin_site* out, uint8_t* gt_row)
{
    int st = VIN_OK;
    memset(out, 0, sizeof *out);
    for (int i = 0; i < 8; ++i) out->allele_acgt[i] = -1;
    for (int32_t s = 0; s < S; ++s) gt_row[s] = 0xFF;
    if (len && line[len - 1] == '\r') --len; /* kstring KS_SEP_LINE strips a CR before the LF */
    out->line_len = (uint32_t)len;

    /* column starts */
    size_t col[10];
    int ncol = 0;
    col[ncol++] = 0;
    for (size_t i = 0; i < len && ncol < 10; ++i)
        if (line[i] == '\t') col[ncol++] = i + 1;
    if (ncol < 10) {
        out->status = VIN_ENCOLS;
        return;
    }
    out->id_off = (uint32_t)col[2];
    out->fmt_off = (uint32_t)col[8];
    out->samples_off = (uint32_t)col[9];

    /* POS: hts_str2uint(p, &p, 63, ...) - 1 */
    {
        const char* p = line + col[1];
        const char* e = line + col[2] - 1;
        if (p < e && *p == '+') ++p;
        uint64_t v = 0;
        int big = 0;
        for (; p < e && *p >= '0' && *p <= '9'; ++p) {
            if (v > (UINT64_MAX - 9) / 10) big = 1;
            else v = v * 10 + (uint64_t)(*p - '0');
        }
        if (big || v > (uint64_t)INT32_MAX) raise_(&st, VIN_EPOS);
        out->pos = (int64_t)v - 1;
    }

    /* REF + ALT */
    int n_allele = 1;
    int amap[5] = {-1, -1, -1, -1, -1};
    {
        const char* a[5];
        size_t al[5];
        a[0] = line + col[3];
        al[0] = col[4] - 1 - col[3];
        const char* p = line + col[4];
        const char* e = line + col[5] - 1;
        if (!(e - p == 1 && *p == '.')) {
            const char* t = p;
            for (const char* r = p;; ++r) {
                if (r == e || *r == ',') {
                    if (n_allele < 5) {
                        a[n_allele] = t;
                        al[n_allele] = (size_t)(r - t);
                    }
                    ++n_allele;
                    t = r + 1;
                }
                if (r == e) break;
            }
        }
        if (n_allele > 5 || (gt_source == 0 && n_allele > 2)) raise_(&st, VIN_ENALLELE);
        int lim = n_allele < 5 ? n_allele : 5;
        for (int i = 0; i < lim; ++i) {
            if (gt_source == 0) {
                int x = al[i] ? a[i][0] - '0' : -1;
                if (x != 0 && x != 1) raise_(&st, VIN_EALLELE);
                else amap[i] = x;
            } else {
                amap[i] = allele_to_acgt(a[i], al[i]);
                if (amap[i] < 0) raise_(&st, VIN_EALLELE);
            }
            out->allele_acgt[i] = (int8_t)amap[i];
        }
    }
    out->n_allele = n_allele;

    /* FORMAT: index of the GT key */
    int gt_idx = -1;
    {
        const char* p = line + col[8];
        const char* e = line + col[9] - 1;
        int j = 0;
        const char* t = p;
        for (const char* r = p;; ++r) {
            if (r == e || *r == ':') {
                if (gt_idx < 0 && r - t == 2 && t[0] == 'G' && t[1] == 'T') gt_idx = j;
                ++j;
                t = r + 1;
            }
            if (r == e) break;
        }
        if (gt_idx < 0) raise_(&st, VIN_ENOGT);
    }

    /* sample columns */
    int64_t allele_sum = 0;
    int32_t s = 0;
    size_t p = col[9];
    while (s < S && p <= len) {
        size_t e = p; /* end of this column */
        while (e < len && line[e] != '\t') ++e;
        if (gt_idx >= 0) {
            /* walk to the GT sub-field */
            size_t q = p;
            int j = 0;
            while (j < gt_idx && q < e) {
                if (line[q] == ':') ++j;
                ++q;
            }
            int n = 0, bad = 0;
            int h[2] = {-1, -1};
            if (j == gt_idx) { /* sub-field present (possibly empty) */
                for (;;) {
                    int val = -2; /* -1 missing */
                    if (q < e && line[q] == '.') {
                        val = -1;
                        ++q;
                    } else {
                        size_t q0 = q;
                        if (q < e && line[q] == '+') ++q;
                        int64_t v = 0;
                        while (q < e && line[q] >= '0' && line[q] <= '9') {
                            if (v < (int64_t)1 << 40) v = v * 10 + (line[q] - '0');
                            ++q;
                        }
                        if (q == q0) bad = 1; /* "value not a number or '.'" */
                        val = v > 1000 ? 1000 : (int)v;
                    }
                    if (n < 2) h[n] = val;
                    ++n;
                    if (q < e && (line[q] == '|' || line[q] == '/')) {
                        ++q;
                        continue;
                    }
                    break;
                }
                if (q < e && line[q] != ':') bad = 1; /* invalid character after the GT vector */
            }
            /* (an empty GT sub-field is "not a number": htslib's `if (!l)` branch at vcf.c:2670 is unreachable; a column with
             * fewer sub-fields than the GT index leaves n == 0 -> missing + vector_end, htslib/vcf.c:2740-2746 -> not diploid) */
            if (bad) raise_(&st, VIN_EGTCHAR);
            else if (n != 2) raise_(&st, VIN_EPLOIDY);
            else {
                uint8_t b = 0;
                for (int k = 0; k < 2; ++k) {
                    int nib = 0xF;
                    if (h[k] >= 0) {
                        if (h[k] >= n_allele) raise_(&st, VIN_EALLELEIDX);
                        else {
                            allele_sum += h[k];
                            int m = h[k] < 5 ? amap[h[k]] : -1;
                            if (m == 4) raise_(&st, VIN_ESYMBOLIC);
                            else if (m >= 0) nib = m;
                        }
                    }
                    b |= (uint8_t)(nib << (4 * k));
                }
                gt_row[s] = b;
            }
        }
        ++s;
        p = e + 1;
    }
    if (s < S) raise_(&st, VIN_ENSAMPLES);

    out->allele_sum = allele_sum;
    out->status = st;
    if (st == VIN_OK) {
        if ((rm_invar & 1) && allele_sum == 0) out->skip_code = -1;
        else if (rm_invar & 2)
            for (int a = 1; a < n_allele; ++a)
                if ((int64_t)a * S * 2 == allele_sum) out->skip_code = -2;
    }
}

/* whole chunk: complete lines only (a tail without '\n' is left unconsumed unless `final`);
 * returns the number of records, *consumed = bytes used */
int64_t vin_oracle_parse(const char* text, size_t n, int32_t S, int32_t gt_source, int32_t rm_invar, int32_t final, int64_t max_records,
                         vin_site* sites, uint8_t* rows, size_t* consumed)
{
    int64_t r = 0;
    size_t p = 0;
    while (p < n && r < max_records) {
        const char* nl = memchr(text + p, '\n', n - p);
        size_t e;
        if (nl) e = (size_t)(nl - text);
        else if (final) e = n;
        else break;
        vin_oracle_line(text + p, e - p, S, gt_source, rm_invar, &sites[r], rows + (size_t)r * S);
        sites[r].line_off = p;
        ++r;
        p = e + 1;
    }
    *consumed = p < n ? p : n;
    return r;
}
