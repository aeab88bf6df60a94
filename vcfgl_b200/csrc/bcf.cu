// VGL_HOST_BCF -- the reference's OUTPUT path on the device: every kept site becomes the exact bytes that
// simRecord::add_tags() (bcf_utils.cpp:426-507) + bcf_write() (htslib/vcf.c:1951-2001, bcf1_sync :1773-1917) emit for it.
//
//   k_bcf_plan   block per site: min / max of the integer planes -> the int8 / int16 / int32 choice bcf_enc_vint makes
//                per tag and record (htslib/vcf.c:2249-2294), then the record length
//   k_bcf_scan   one block: exclusive prefix of the record lengths = byte offset of every record
//   k_bcf_emit   block per site: thread 0 lays out the record as <= 24 segments (literal bytes built in shared memory:
//                fixed words, allele strings, INFO values, FORMAT keys and descriptors; pass-through bytes of the input
//                record; the tag planes); every thread then produces aligned 32-bit words of the record from whichever
//                segments they fall in (funnel shift for float / int32 data, narrowing for int8 / int16)
// The planes are read as the simulate kernels left them (int32 / float32, [S][G] and [S][A] blocks at g_off / r_off), so
// the same pass serves every kernel set and both replay and native submits.  Records are packed back to back: the host
// writes the buffer as it is.
#include "vgl_internal.h"

#include <cstdlib>

namespace vgl {

namespace {

enum { BT_NULL = 0, BT_INT8 = 1, BT_INT16 = 2, BT_INT32 = 3, BT_FLOAT = 5, BT_CHAR = 7 };
enum { SEG_LIT = 0, SEG_BLOB = 1, SEG_VERB = 2, SEG_I8 = 3, SEG_I16 = 4 };
constexpr int MAX_SEG = 40, LIT_CAP = 512;

using SiteMinMax = BcfSiteMinMax; // per tag: 0 dp, 1 pl, 2 ad, 3 adf, 4 adr

__device__ __forceinline__ int int_type(int32_t mn, int32_t mx) // htslib/vcf.c:2261-2294; no real value: max = INT32_MIN, min = INT32_MAX
{
    if (mx <= 127 && mn >= -120) return BT_INT8;
    if (mx <= 32767 && mn >= -32760) return BT_INT16;
    return BT_INT32;
}
__device__ __forceinline__ int type_width(int t) { return t == BT_INT8 ? 1 : (t == BT_INT16 ? 2 : 4); }

struct Builder {
    uint8_t* lit;      // shared memory, LIT_CAP bytes (null: count only)
    uint32_t* seg_start;
    uint32_t* seg_kind;
    unsigned long long* seg_src;
    BcfRecPlanes* planes = nullptr; // receives the FORMAT plane layout
    int nseg = 0;
    uint32_t pos = 0;  // bytes of the record so far
    uint32_t nlit = 0;
    bool in_lit = false;

    __device__ void put(uint8_t b)
    {
        if (!in_lit) {
            if (lit) { seg_start[nseg] = pos; seg_kind[nseg] = SEG_LIT; seg_src[nseg] = nlit; }
            ++nseg;
            in_lit = true;
        }
        if (lit) lit[nlit] = b;
        ++nlit;
        ++pos;
    }
    __device__ void put16(uint32_t v) { put((uint8_t)v); put((uint8_t)(v >> 8)); }
    __device__ void put32(uint32_t v) { put16(v); put16(v >> 16); }
    __device__ void ext(int kind, const void* src, uint32_t nbytes)
    {
        if (nbytes == 0) return;
        if (lit) { seg_start[nseg] = pos; seg_kind[nseg] = (uint32_t)kind; seg_src[nseg] = (unsigned long long)src; }
        ++nseg;
        in_lit = false;
        pos += nbytes;
    }
    // bcf_enc_size, htslib/vcf.h:1392-1414
    __device__ void size(int n, int type)
    {
        if (n >= 15) {
            put((uint8_t)(15 << 4 | type));
            if (n >= 128) {
                if (n >= 32768) { put(1 << 4 | BT_INT32); put32((uint32_t)n); }
                else { put(1 << 4 | BT_INT16); put16((uint32_t)n); }
            } else { put(1 << 4 | BT_INT8); put((uint8_t)n); }
        } else put((uint8_t)(n << 4 | type));
    }
    // bcf_enc_int1, htslib/vcf.h:1423-1446
    __device__ void int1(int32_t x)
    {
        if (x == VGL_I32_MISSING) { size(1, BT_INT8); put(0x80); }
        else if (x == VGL_I32_MISSING + 1) { size(1, BT_INT8); put(0x81); }
        else if (x <= 127 && x >= -120) { size(1, BT_INT8); put((uint8_t)x); }
        else if (x <= 32767 && x >= -32760) { size(1, BT_INT16); put16((uint32_t)x); }
        else { size(1, BT_INT32); put32((uint32_t)x); }
    }
    // bcf_enc_vint(s, n, a, -1) of a short INFO vector, htslib/vcf.c:2249-2294
    __device__ void vint_small(const int32_t* a, int n)
    {
        if (n <= 0) { size(0, BT_NULL); return; }
        if (n == 1) { int1(a[0]); return; }
        int32_t mx = INT32_MIN, mn = INT32_MAX;
        for (int i = 0; i < n; ++i) {
            if (a[i] == VGL_I32_MISSING || a[i] == VGL_I32_MISSING + 1) continue;
            mx = max(mx, a[i]);
            mn = min(mn, a[i]);
        }
        const int t = int_type(mn, mx);
        size(n, t);
        for (int i = 0; i < n; ++i) {
            const int32_t v = a[i];
            if (t == BT_INT8) put(v == VGL_I32_MISSING ? 0x80 : (v == VGL_I32_MISSING + 1 ? 0x81 : (uint8_t)v));
            else if (t == BT_INT16) put16(v == VGL_I32_MISSING ? 0x8000u : (v == VGL_I32_MISSING + 1 ? 0x8001u : (uint32_t)v));
            else put32((uint32_t)v);
        }
    }
    __device__ void vfloat_small(const float* a, int n) // bcf_enc_vfloat, htslib/vcf.c:2337-2343
    {
        size(n, BT_FLOAT);
        for (int i = 0; i < n; ++i) put32(__float_as_uint(a[i]));
    }
    __device__ void str(const char* s, int n) // bcf_enc_vchar
    {
        size(n, BT_CHAR);
        for (int i = 0; i < n; ++i) put((uint8_t)s[i]);
    }
};

// typed integer of a BCF stream (bcf_enc_int1's output): advances p
__device__ __forceinline__ int32_t read_typed_int(const uint8_t*& p)
{
    const int t = *p++ & 0xF;
    int32_t v = 0;
    if (t == BT_INT8) { v = (int8_t)p[0]; p += 1; }
    else if (t == BT_INT16) { v = (int16_t)((uint32_t)p[0] | ((uint32_t)p[1] << 8)); p += 2; }
    else { v = (int32_t)((uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24)); p += 4; }
    return v;
}
// one FORMAT block of the input record: key, values per sample x width -> bytes of the whole block
__device__ __forceinline__ uint32_t in_fmt_block(const uint8_t* p0, int S, int32_t& key)
{
    const uint8_t* p = p0;
    key = read_typed_int(p);
    const int d = *p++;
    int n = d >> 4;
    const int ty = d & 0xF;
    if (n == 15) n = read_typed_int(p);
    const int w = ty == BT_INT8 || ty == BT_CHAR ? 1 : (ty == BT_INT16 ? 2 : (ty == BT_NULL ? 0 : 4));
    return (uint32_t)(p - p0) + (uint32_t)S * (uint32_t)n * (uint32_t)w;
}

// Lays out site `i`'s record.  Returns its length; l_shared / l_indiv as bcf_write() counts them.
__device__ uint32_t bcf_layout(const BcfArgs& a, int i, const vgl_site_out& s, const SiteMinMax& mm, Builder& b)
{
    const vgl_bcf_site_in in = a.site_in[i];
    const int S = a.S, A = s.n_alleles, G = s.n_genotypes;
    const uint32_t t = a.tag_mask;
    const int n_info_sim = !!(t & VGL_TAG_INFO_DP) + !!(t & VGL_TAG_QS) + !!(t & VGL_TAG_I16) + !!(t & VGL_TAG_INFO_AD) +
                           !!(t & VGL_TAG_INFO_ADF) + !!(t & VGL_TAG_INFO_ADR);
    // the simulator's FORMAT tags in add_tags() order, and which of them take the slot of an input block with the same key
    const uint32_t fbit[7] = {VGL_TAG_FMT_DP, VGL_TAG_GL, VGL_TAG_PL, VGL_TAG_GP, VGL_TAG_FMT_AD, VGL_TAG_FMT_ADF, VGL_TAG_FMT_ADR};
    const int32_t fkey[7] = {a.dict.dp, a.dict.gl, a.dict.pl, a.dict.gp, a.dict.ad, a.dict.adf, a.dict.adr};
    uint32_t placed = 0u;
    if (in.n_fmt) {
        const uint8_t* q = a.blob + in.fmt_off;
        for (uint32_t k = 0; k < in.n_fmt; ++k) {
            int32_t key;
            const uint32_t len = in_fmt_block(q, S, key);
            for (int f = 0; f < 7; ++f)
                if ((t & fbit[f]) && fkey[f] == key) placed |= 1u << f;
            q += len;
        }
    }
    int n_fmt = (int)in.n_fmt;
    for (int f = 0; f < 7; ++f)
        if ((t & fbit[f]) && !((placed >> f) & 1u)) ++n_fmt;
    // ---- the 32 fixed bytes (htslib/vcf.c:1984-1993); the two lengths are patched at the end
    b.put32(0); b.put32(0);
    b.put32((uint32_t)in.rid);
    b.put32((uint32_t)in.pos);
    const bool no_reads = s.info_dp == 0;
    // allele strings (vcfgl.cpp:739-782; sites without reads :242-277) and rlen = strlen(REF) (htslib/vcf.c:4607-4611)
    const bool nonref_name = a.do_unobserved == 2 || a.do_unobserved == 5 || (no_reads && a.do_gvcf);
    int code[5]; // 0..3 = A,C,G,T; 4 = <*> / <NON_REF>; 5 = "."
    for (int k = 0; k < 5; ++k) code[k] = k < A ? (int)s.alleles2acgt[k] : -1;
    if (no_reads) {
        if (a.do_gvcf || a.do_unobserved == 1 || a.do_unobserved == 2) code[0] = 4;
        else if (a.do_unobserved == 0) code[0] = 5;
        else { code[0] = 0; code[1] = 1; code[2] = 2; code[3] = 3; code[4] = 4; }
    }
    const int len0 = code[0] == 4 ? (nonref_name ? 9 : 3) : 1;
    b.put32((uint32_t)len0);
    b.put32(in.qual_bits);
    b.put16((uint32_t)(in.n_info + n_info_sim));
    b.put16((uint32_t)A);
    b.put32(((uint32_t)n_fmt << 24) | ((uint32_t)S & 0xFFFFFFu));
    // ---- shared block: ID, alleles, FILTER + the input's INFO, the simulator's INFO in add_tags() order
    if (in.id_len) b.ext(SEG_BLOB, a.blob + in.id_off, in.id_len);
    else b.put(0x07);
    for (int k = 0; k < A; ++k) {
        if (code[k] == 4) { if (nonref_name) b.str("<NON_REF>", 9); else b.str("<*>", 3); }
        else if (code[k] == 5) b.str(".", 1);
        else { const char c = "ACGT"[code[k] & 3]; b.str(&c, 1); }
    }
    if (in.flt_info_len) b.ext(SEG_BLOB, a.blob + in.flt_info_off, in.flt_info_len);
    else b.put(0x00);
    if (t & VGL_TAG_INFO_DP) { b.int1(a.dict.dp); b.int1(s.info_dp); }
    if (t & VGL_TAG_QS) { b.int1(a.dict.qs); b.vfloat_small(s.qs, A); }
    if (t & VGL_TAG_I16) { b.int1(a.dict.i16); b.vfloat_small(s.i16, 16); }
    if (t & VGL_TAG_INFO_AD) { b.int1(a.dict.ad); b.vint_small(s.info_ad, A); }
    if (t & VGL_TAG_INFO_ADF) { b.int1(a.dict.adf); b.vint_small(s.info_adf, A); }
    if (t & VGL_TAG_INFO_ADR) { b.int1(a.dict.adr); b.vint_small(s.info_adr, A); }
    const uint32_t l_shared = b.pos - 8;
    // ---- FORMAT block in add_tags() order: typed key, (values per sample, type), S * nps values (htslib/vcf.c:4453-4470)
    auto fmt_int = [&](int32_t key, const int32_t* src, int nps, int which) {
        b.int1(key);
        const int ty = int_type(mm.mn[which], mm.mx[which]);
        b.size(nps, ty);
        if (b.planes && b.planes->n < 7) { b.planes->off[b.planes->n] = b.pos; b.planes->cell[b.planes->n] = (uint16_t)(nps * type_width(ty)); ++b.planes->n; }
        b.ext(ty == BT_INT8 ? SEG_I8 : (ty == BT_INT16 ? SEG_I16 : SEG_VERB), src, (uint32_t)S * nps * type_width(ty));
    };
    auto fmt_float = [&](int32_t key, const float* src, int nps) {
        b.int1(key);
        b.size(nps, BT_FLOAT);
        if (b.planes && b.planes->n < 7) { b.planes->off[b.planes->n] = b.pos; b.planes->cell[b.planes->n] = (uint16_t)(nps * 4); ++b.planes->n; }
        b.ext(SEG_VERB, src, (uint32_t)S * nps * 4u);
    };
    auto put_tag = [&](int f) {
        switch (f) {
        case 0: fmt_int(a.dict.dp, a.dp + (size_t)i * S, 1, 0); break;
        case 1: fmt_float(a.dict.gl, a.gl + s.g_off, G); break;
        case 2: fmt_int(a.dict.pl, a.pl + s.g_off, G, 1); break;
        case 3: fmt_float(a.dict.gp, a.gp + s.g_off, G); break;
        case 4: fmt_int(a.dict.ad, a.ad + s.r_off, A, 2); break;
        case 5: fmt_int(a.dict.adf, a.adf + s.r_off, A, 3); break;
        default: fmt_int(a.dict.adr, a.adr + s.r_off, A, 4); break;
        }
    };
    if (in.n_fmt) { // the input's blocks first, in their order: a simulated tag with the same key takes the block's slot
        const uint8_t* q = a.blob + in.fmt_off;
        for (uint32_t k = 0; k < in.n_fmt; ++k) {
            int32_t key;
            const uint32_t len = in_fmt_block(q, S, key);
            int hit = -1;
            for (int f = 0; f < 7; ++f)
                if ((t & fbit[f]) && fkey[f] == key) hit = f;
            if (hit >= 0) put_tag(hit);
            else b.ext(SEG_BLOB, q, len);
            q += len;
        }
    }
    for (int f = 0; f < 7; ++f)
        if ((t & fbit[f]) && !((placed >> f) & 1u)) put_tag(f);
    const uint32_t l_indiv = b.pos - 8 - l_shared;
    if (b.lit) {
        const uint32_t v[2] = {l_shared, l_indiv};
        for (int k = 0; k < 8; ++k) b.lit[k] = (uint8_t)(v[k >> 2] >> (8 * (k & 3)));
    }
    return b.pos;
}

// A gVCF block record (GVCF_FLUSH_BLOCK, bcf_utils.cpp:896-925): a cleared record with the founder's contig / position /
// alleles, rlen = END - start, INFO END (blocks longer than one position), MIN_DP, QS (the founder's), FORMAT PL then DP (the
// per-sample minima over the members, reduced by k_gvcf_reduce).  `s` is the founder's site record.
__device__ uint32_t bcf_layout_block(const BcfArgs& a, const vgl_gvcf_rec& r, const vgl_site_out& s, const SiteMinMax& mm, Builder& b)
{
    const int S = a.S;
    const vgl_bcf_site_in first = a.site_in[r.first_site], last = a.site_in[r.last_site];
    const int32_t end1 = last.pos + 1; // 1-based END
    const bool has_end = end1 - first.pos >= 2, has_qs = (a.tag_mask & VGL_TAG_QS) != 0, has_pl = a.blk_pl != nullptr;
    b.put32(0); b.put32(0);
    b.put32((uint32_t)first.rid);
    b.put32((uint32_t)first.pos);
    b.put32((uint32_t)(end1 - first.pos)); // rlen
    b.put32(VGL_F32_MISSING_BITS);         // QUAL of a cleared record
    b.put16((uint32_t)((has_end ? 1 : 0) + 1 + (has_qs ? 1 : 0)));
    b.put16((uint32_t)s.n_alleles);
    b.put32(((uint32_t)((has_pl ? 1 : 0) + 1) << 24) | ((uint32_t)S & 0xFFFFFFu));
    b.put(0x07); // ID "."
    const bool nonref_name = a.do_unobserved == 2 || a.do_unobserved == 5;
    for (int k = 0; k < s.n_alleles; ++k) {
        const int code = s.alleles2acgt[k];
        if (code == 4) { if (nonref_name) b.str("<NON_REF>", 9); else b.str("<*>", 3); }
        else { const char c = "ACGT"[code & 3]; b.str(&c, 1); }
    }
    b.put(0x00); // FILTER: none
    if (has_end) { b.int1(a.dict.end); b.int1(end1); }
    b.int1(a.dict.min_dp); b.int1(r.min_dp);
    if (has_qs) { b.int1(a.dict.qs); b.vfloat_small(s.qs, s.n_alleles); }
    const uint32_t l_shared = b.pos - 8;
    auto fmt_int = [&](int32_t key, const int32_t* src, int nps, int which) {
        b.int1(key);
        const int ty = int_type(mm.mn[which], mm.mx[which]);
        b.size(nps, ty);
        if (b.planes && b.planes->n < 7) { b.planes->off[b.planes->n] = b.pos; b.planes->cell[b.planes->n] = (uint16_t)(nps * type_width(ty)); ++b.planes->n; }
        b.ext(ty == BT_INT8 ? SEG_I8 : (ty == BT_INT16 ? SEG_I16 : SEG_VERB), src, (uint32_t)S * nps * type_width(ty));
    };
    if (has_pl) fmt_int(a.dict.pl, a.blk_pl + (size_t)r.plane * S * 3, 3, 1);
    fmt_int(a.dict.dp, a.blk_dp + (size_t)r.plane * S, 1, 0);
    const uint32_t l_indiv = b.pos - 8 - l_shared;
    if (b.lit) {
        const uint32_t v[2] = {l_shared, l_indiv};
        for (int k = 0; k < 8; ++k) b.lit[k] = (uint8_t)(v[k >> 2] >> (8 * (k & 3)));
    }
    return b.pos;
}

__device__ __forceinline__ void mm_update(int32_t v, int32_t& mn, int32_t& mx)
{
    if (v != VGL_I32_MISSING && v != VGL_I32_MISSING + 1) { mn = min(mn, v); mx = max(mx, v); }
}

__global__ void __launch_bounds__(128) k_bcf_plan(const BcfArgs a)
{
    const int rk = blockIdx.x, tid = threadIdx.x; // record index: the site itself, or (-doGVCF) the merger's k-th record
    __shared__ vgl_site_out s;
    __shared__ int32_t red[2][5][4];
    __shared__ vgl_gvcf_rec grec;
    if (a.recs && rk >= a.rec_counts[0]) {
        if (tid == 0) a.rec_len[rk] = 0u;
        return;
    }
    if (tid == 0) {
        if (a.recs) grec = a.recs[rk];
        s = a.sites[a.recs ? a.recs[rk].first_site : rk];
    }
    __syncthreads();
    const int i = a.recs ? grec.first_site : rk;
    const int S = a.S;
    if (a.recs && grec.n_members > 0) { // block record: min / max of its DP and PL planes
        int32_t mn[2] = {INT32_MAX, INT32_MAX}, mx[2] = {INT32_MIN, INT32_MIN};
        const int32_t* pd = a.blk_dp + (size_t)grec.plane * S;
        for (int k = tid; k < S; k += 128) mm_update(pd[k], mn[0], mx[0]);
        if (a.blk_pl) {
            const int32_t* pp = a.blk_pl + (size_t)grec.plane * S * 3;
            for (int k = tid; k < 3 * S; k += 128) mm_update(pp[k], mn[1], mx[1]);
        }
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            mn[k] = __reduce_min_sync(0xffffffffu, mn[k]);
            mx[k] = __reduce_max_sync(0xffffffffu, mx[k]);
            if ((tid & 31) == 0) { red[0][k][tid >> 5] = mn[k]; red[1][k][tid >> 5] = mx[k]; }
        }
        __syncthreads();
        if (tid == 0) {
            SiteMinMax mm;
            for (int k = 0; k < 5; ++k) { mm.mn[k] = INT32_MAX; mm.mx[k] = INT32_MIN; }
            for (int k = 0; k < 2; ++k) {
                mm.mn[k] = min(min(red[0][k][0], red[0][k][1]), min(red[0][k][2], red[0][k][3]));
                mm.mx[k] = max(max(red[1][k][0], red[1][k][1]), max(red[1][k][2], red[1][k][3]));
            }
            a.minmax[rk] = mm;
            Builder b;
            b.lit = nullptr;
            a.rec_len[rk] = bcf_layout_block(a, grec, s, mm, b);
        }
        return;
    }
    if (s.skip_code != 0) {
        if (tid == 0) a.rec_len[rk] = 0u;
        return;
    }
    const int64_t nG = (int64_t)S * s.n_genotypes, nA = (int64_t)S * s.n_alleles;
    int32_t mn[5], mx[5];
#pragma unroll
    for (int k = 0; k < 5; ++k) { mn[k] = INT32_MAX; mx[k] = INT32_MIN; }
    if (a.tag_mask & VGL_TAG_FMT_DP) {
        const int32_t* p = a.dp + (size_t)i * S;
        for (int k = tid; k < S; k += 128) mm_update(p[k], mn[0], mx[0]);
    }
    if (a.pl) {
        const int32_t* p = a.pl + s.g_off;
        for (int64_t k = tid; k < nG; k += 128) mm_update(p[k], mn[1], mx[1]);
    }
    const int32_t* const rp[3] = {a.ad, a.adf, a.adr};
#pragma unroll
    for (int q = 0; q < 3; ++q)
        if (rp[q]) {
            const int32_t* p = rp[q] + s.r_off;
            for (int64_t k = tid; k < nA; k += 128) mm_update(p[k], mn[2 + q], mx[2 + q]);
        }
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        mn[k] = __reduce_min_sync(0xffffffffu, mn[k]);
        mx[k] = __reduce_max_sync(0xffffffffu, mx[k]);
        if ((tid & 31) == 0) { red[0][k][tid >> 5] = mn[k]; red[1][k][tid >> 5] = mx[k]; }
    }
    __syncthreads();
    if (tid == 0) {
        SiteMinMax mm;
        for (int k = 0; k < 5; ++k) {
            mm.mn[k] = min(min(red[0][k][0], red[0][k][1]), min(red[0][k][2], red[0][k][3]));
            mm.mx[k] = max(max(red[1][k][0], red[1][k][1]), max(red[1][k][2], red[1][k][3]));
        }
        a.minmax[rk] = mm;
        Builder b;
        b.lit = nullptr;
        a.rec_len[rk] = bcf_layout(a, i, s, mm, b);
    }
}

// exclusive prefix of rec_len -> rec_off[n + 1]; total also into totals[3] (device) for the host
__global__ void __launch_bounds__(1024) k_bcf_scan(const BcfArgs a)
{
    __shared__ long long warp_sum[32];
    __shared__ long long carry_s;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    if (tid == 0) carry_s = 0;
    __syncthreads();
    for (int base = 0; base < a.n_sites; base += 1024) { // slots beyond the merger's record count hold length 0
        const int i = base + tid;
        const long long v = i < a.n_sites ? (long long)a.rec_len[i] : 0;
        long long x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const long long y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) warp_sum[w] = x;
        __syncthreads();
        if (w == 0) {
            long long t = warp_sum[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const long long y = __shfl_up_sync(0xffffffffu, t, o);
                if (lane >= o) t += y;
            }
            warp_sum[lane] = t;
        }
        __syncthreads();
        const long long carry = carry_s;
        const long long incl = carry + (w ? warp_sum[w - 1] : 0) + x;
        if (i < a.n_sites) a.rec_off[i] = incl - v;
        __syncthreads();
        if (tid == 1023) carry_s = incl;
        __syncthreads();
    }
    if (tid == 0) {
        a.rec_off[a.n_sites] = carry_s;
        a.totals[3] = carry_s;
        if (carry_s > a.out_cap) atomicExch(a.status, (int)VGL_EOVERFLOW);
    }
}

struct SegView {
    const uint32_t* start;
    const uint32_t* kind;
    const unsigned long long* src;
    const uint8_t* lit;
};

__device__ __forceinline__ uint32_t seg_byte(const SegView& v, int sg, uint32_t r) // byte r of segment sg
{
    const unsigned long long src = v.src[sg];
    switch (v.kind[sg]) {
    case SEG_LIT: return v.lit[(uint32_t)src + r];
    case SEG_BLOB: return reinterpret_cast<const uint8_t*>(src)[r];
    case SEG_VERB: return (__ldg(reinterpret_cast<const uint32_t*>(src) + (r >> 2)) >> (8 * (r & 3))) & 0xFFu;
    case SEG_I8: {
        const int32_t x = __ldg(reinterpret_cast<const int32_t*>(src) + r);
        return x == VGL_I32_MISSING ? 0x80u : (x == VGL_I32_MISSING + 1 ? 0x81u : (uint32_t)x & 0xFFu);
    }
    default: {
        const int32_t x = __ldg(reinterpret_cast<const int32_t*>(src) + (r >> 1));
        const uint32_t h = x == VGL_I32_MISSING ? 0x8000u : (x == VGL_I32_MISSING + 1 ? 0x8001u : (uint32_t)x & 0xFFFFu);
        return (h >> (8 * (r & 1))) & 0xFFu;
    }
    }
}

// U = words per thread and round: 4 for long records (the plane loads of a round are all in flight together), 1 for short ones,
// where the kernel is bound by thread 0's layout pass and more resident blocks matter more than longer rounds
template <int U>
__global__ void __launch_bounds__(256, U == 1 ? 8 : (U == 2 ? 6 : 5)) k_bcf_emit(const BcfArgs a)
{
    const int rk = blockIdx.x, tid = threadIdx.x; // record index (see k_bcf_plan)
    __shared__ vgl_site_out s;
    __shared__ uint32_t seg_start[MAX_SEG + 1], seg_kind[MAX_SEG];
    __shared__ unsigned long long seg_src[MAX_SEG];
    __shared__ __align__(4) uint8_t lit[LIT_CAP];
    __shared__ int nseg_s;
    const uint32_t len = a.rec_len[rk];
    if (len == 0) return;
    const long long off = a.rec_off[rk];
    if (off + len > a.out_cap) return; // status already raised by the scan
    if (tid == 0) {
        vgl_gvcf_rec grec;
        grec.n_members = 0;
        grec.first_site = rk;
        if (a.recs) grec = a.recs[rk];
        const int i = grec.first_site;
        s = a.sites[i];
        Builder b;
        b.lit = lit; b.seg_start = seg_start; b.seg_kind = seg_kind; b.seg_src = seg_src;
        BcfRecPlanes pl;
        pl.n = 0;
        for (int k = 0; k < 7; ++k) { pl.off[k] = 0u; pl.cell[k] = 0; }
        if (a.planes) b.planes = &pl;
        const uint32_t got = grec.n_members > 0 ? bcf_layout_block(a, grec, s, a.minmax[rk], b) : bcf_layout(a, i, s, a.minmax[rk], b);
        if (a.planes) a.planes[rk] = pl;
        nseg_s = b.nseg;
        seg_start[b.nseg] = got;
    }
    __syncthreads();
    const int nseg = nseg_s;
    SegView v{seg_start, seg_kind, seg_src, lit};
    uint8_t* const out = a.out;
    // aligned words of the output buffer that overlap [off, off + len)
    const long long w0 = off >> 2, w1 = (off + len + 3) >> 2;
    int sg = 0;
    // U words per thread and round (w, w + 256, ...): the plane loads of all of them are issued before any is used -- the loop is
    // bound by the latency of those loads, not by their number.
    for (long long wb = w0 + tid; wb < w1; wb += U * 256) {
        int32_t x[U][4];
        uint32_t info[U]; // 0: not a plane word (literal bytes, pass-through bytes, or it straddles segments); else kind | r << 8 (r: low bits only)
        int sgu[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long w = wb + 256 * u;
            info[u] = 0u;
            sgu[u] = sg;
            if (w >= w1) continue;
            const long long p0 = w * 4 - off; // record-relative position of the word's first byte (may be < 0)
            const uint32_t rfirst = p0 < 0 ? 0u : (uint32_t)p0;
            while (sg + 1 < nseg && seg_start[sg + 1] <= rfirst) ++sg;
            sgu[u] = sg;
            const uint32_t kind = seg_kind[sg];
            if (p0 >= 0 && (uint32_t)p0 + 4u <= seg_start[sg + 1] && kind >= SEG_VERB) { // whole word inside one plane segment
                const uint32_t r = (uint32_t)p0 - seg_start[sg];
                const int32_t* q = reinterpret_cast<const int32_t*>(seg_src[sg]);
                int n;
                if (kind == SEG_VERB) { q += r >> 2; n = 1 + ((r & 3u) != 0u); }
                else if (kind == SEG_I8) { q += r; n = 4; }
                else { q += r >> 1; n = 2 + (int)(r & 1u); }
                info[u] = kind | ((r & 3u) << 8);
#pragma unroll
                for (int k = 0; k < 4; ++k) x[u][k] = k < n ? __ldg(q + k) : 0;
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long w = wb + 256 * u;
            if (w >= w1) continue;
            if (info[u]) {
                const uint32_t kind = info[u] & 0xFFu, r = info[u] >> 8;
                uint32_t word;
                if (kind == SEG_VERB) {
                    word = r ? __funnelshift_r((uint32_t)x[u][0], (uint32_t)x[u][1], 8 * r) : (uint32_t)x[u][0];
                } else if (kind == SEG_I8) {
                    word = 0;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const int32_t v8 = x[u][k];
                        word |= (v8 == VGL_I32_MISSING ? 0x80u : (v8 == VGL_I32_MISSING + 1 ? 0x81u : (uint32_t)v8 & 0xFFu)) << (8 * k);
                    }
                } else {
                    uint32_t h[3];
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        const int32_t v16 = x[u][k];
                        h[k] = v16 == VGL_I32_MISSING ? 0x8000u : (v16 == VGL_I32_MISSING + 1 ? 0x8001u : (uint32_t)v16 & 0xFFFFu);
                    }
                    word = (r & 1u) ? ((h[0] >> 8) | (h[1] << 8) | (h[2] << 24)) : (h[0] | (h[1] << 16));
                }
                *reinterpret_cast<uint32_t*>(out + w * 4) = word;
            } else { // byte by byte, only this record's bytes
                const long long p0 = w * 4 - off;
                int g = sgu[u];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const long long p = p0 + k;
                    if (p < 0 || p >= (long long)len) continue;
                    while (g + 1 < nseg && seg_start[g + 1] <= (uint32_t)p) ++g;
                    out[w * 4 + k] = (uint8_t)seg_byte(v, g, (uint32_t)p - seg_start[g]);
                }
            }
        }
    }
}

} // namespace

void launch_bcf(const BcfArgs& a, cudaStream_t st)
{
    k_bcf_plan<<<(unsigned)a.n_sites, 128, 0, st>>>(a);
    k_bcf_scan<<<1, 1024, 0, st>>>(a);
    int u = a.S >= 1000 ? 4 : 1;
    if (const char* e = getenv("VGL_EMIT_U")) u = atoi(e); // development: words per thread and round
    if (u >= 4) k_bcf_emit<4><<<(unsigned)a.n_sites, 256, 0, st>>>(a);
    else if (u == 2) k_bcf_emit<2><<<(unsigned)a.n_sites, 256, 0, st>>>(a);
    else k_bcf_emit<1><<<(unsigned)a.n_sites, 256, 0, st>>>(a);
}

} // namespace vgl
