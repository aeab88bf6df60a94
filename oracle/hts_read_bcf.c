/* TEST INFRASTRUCTURE (oracle/): reads a BCF file -- uncompressed or BGZF -- with the reference's own htslib (bundled 1.15.1,
 * linked from oracle/_ref/libhts_ref.a) and prints one line per record: rid, pos, rlen, n_allele, n_info, n_fmt, n_sample, the
 * lengths of the shared and individual blocks and an FNV-1a checksum of their bytes.  tests/ use it to show that the library the
 * reference writes and reads its files with (hts_open / bcf_hdr_read / bcf_read: htslib/vcf.c) accepts the BGZF blocks the
 * device compressed (VGL_HOST_BGZF) and sees the same records as in the uncompressed file.  Never part of the product path. */
#include <stdint.h>
#include <stdio.h>

#include "htslib/hts.h"
#include "htslib/vcf.h"

static uint32_t fnv(const char* s, size_t n, uint32_t h)
{
    for (size_t i = 0; i < n; ++i) h = (h ^ (unsigned char)s[i]) * 16777619u;
    return h;
}

int main(int argc, char** argv)
{
    if (argc < 2) { fprintf(stderr, "usage: hts_read_bcf file.bcf\n"); return 2; }
    htsFile* f = hts_open(argv[1], "r");
    if (!f) { fprintf(stderr, "hts_open failed\n"); return 1; }
    const htsFormat* fmt = hts_get_format(f);
    bcf_hdr_t* h = bcf_hdr_read(f);
    if (!h) { fprintf(stderr, "bcf_hdr_read failed\n"); return 1; }
    printf("format %d compression %d samples %d\n", (int)fmt->format, (int)fmt->compression, bcf_hdr_nsamples(h));
    bcf1_t* r = bcf_init();
    int rc, n = 0;
    while ((rc = bcf_read(f, h, r)) == 0) {
        printf("%d %lld %lld %d %d %d %d %zu %zu %08x\n", r->rid, (long long)r->pos, (long long)r->rlen, (int)r->n_allele, (int)r->n_info, (int)r->n_fmt,
               (int)r->n_sample, (size_t)r->shared.l, (size_t)r->indiv.l, fnv(r->indiv.s, r->indiv.l, fnv(r->shared.s, r->shared.l, 2166136261u)));
        ++n;
    }
    if (rc < -1) { fprintf(stderr, "bcf_read failed after %d records\n", n); return 1; }
    printf("records %d\n", n);
    bcf_destroy(r);
    bcf_hdr_destroy(h);
    return hts_close(f) == 0 ? 0 : 1;
}
