// Input path (include/vgl.h "Input path"): VCF text records -> packed true genotypes.
//
//   k_vcf_count / k_vcf_scan / k_vcf_index   positions of the line feeds = the record index
//   k_vcf_hdr     one thread per record: the nine fixed columns of records whose sample columns can be fixed-width ("a|b" + tab)
//   k_vcf_cells   one thread per eight samples of those records: word compares, a 4-entry byte table, aligned 64-bit stores
//   k_vcf_gt      every other record, one warp per record: the nine fixed columns (POS, REF/ALT -> allele map, FORMAT -> GT index), then every
//                 sample column's GT sub-field -> one packed byte; skip decision of --rm-invar-sites bits 1 / 2
//   k_place_rows  genotype rows -> a slot's genotype matrix (drops skipped records, inserts -explode sites)
//
// What is restated (behaviour, not code): htslib/vcf.c:3041-3110 (vcf_parse columns), 2425-2520 and 2643-2673
// (vcf_parse_format, GT vector), vcfgl.cpp:75-163 (check_rec_alleles).  Bound: HBM -- 4 text bytes in, 1 packed byte out per
// (record, sample) cell for msprime-shaped input ("a|b\t").
#include "vgl_internal.h"

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <new>
#include <string>

namespace vgl {

namespace {

// record index: the text is cut into warp tiles of 4 KiB (8 x 16 bytes per lane, interleaved so that every load is coalesced)
constexpr int WT_CHUNKS = 8;
constexpr int WT_BYTES = WT_CHUNKS * 32 * 16;
constexpr int IDX_WARPS = 8;          // warps per block of k_vcf_count / k_vcf_index

// counters (device words, mirrored to pinned host memory after a parse)
enum { C_NEWLINES = 0, C_NWORK = 1, C_NRECORDS = 2, C_NERRORS = 3, C_FIRSTERR = 4, C_NKEPT = 5, C_CONSUMED = 6, C_DONE = 7, C_COUNT = 8 };

// bit i of the result = byte i of w equals c
__device__ __forceinline__ uint32_t eq_nibble(uint32_t w, uint32_t c4)
{
    const uint32_t x = __vcmpeq4(w, c4) & 0x08040201u;
    return (x * 0x01010101u) >> 24;
}

__device__ __forceinline__ uint32_t eq_mask16(const uint4& v, uint32_t c4)
{
    return eq_nibble(v.x, c4) | (eq_nibble(v.y, c4) << 4) | (eq_nibble(v.z, c4) << 8) | (eq_nibble(v.w, c4) << 12);
}

// ---- record index --------------------------------------------------------------------------------------------------
// line_end[i] = byte offset of the i-th line feed.  Two launches instead of one chained scan (a decoupled
// look-back over ~7000 tiles that each take < 1 us to load spends its time waiting on predecessors: 120 us for 56 MB):
//   k_vcf_count  line feeds per warp tile and per block (streams the text once from HBM); the last block scans the block totals
//                (it also leaves one 16-bit line-feed mask per 16 bytes of text, in text order)
//   k_vcf_index  positions of the line feeds from those masks (1/8 of the text's size; the text is not read again)
__device__ __forceinline__ void load_warp_tile(const uint8_t* __restrict__ text, uint32_t n_bytes, uint32_t tile, int lane, uint32_t (&m)[WT_CHUNKS])
{
    const uint32_t base = tile * (uint32_t)WT_BYTES + (uint32_t)lane * 16u;
#pragma unroll
    for (int j = 0; j < WT_CHUNKS; ++j) {
        const uint32_t off = base + (uint32_t)j * 512u;
        uint32_t mm = 0;
        if (off < n_bytes) { // the buffer is padded beyond n_bytes
            const uint4 v = __ldg(reinterpret_cast<const uint4*>(text + off));
            mm = eq_mask16(v, 0x0A0A0A0Au);
            const uint32_t left = n_bytes - off;
            if (left < 16) mm &= (1u << left) - 1u;
        }
        m[j] = mm;
    }
}

// Block b owns the warp tiles [b * tpb, (b + 1) * tpb).  k_vcf_count leaves the line feeds of every warp tile in tile_count and
// of every block in block_base; the last block to finish turns block_base into an exclusive prefix (at most MAX_IDX_BLOCKS
// values: no separate scan launch) and posts the total.
constexpr int MAX_IDX_BLOCKS = 2048;

__global__ void __launch_bounds__(IDX_WARPS * 32) k_vcf_count(const uint8_t* __restrict__ text, uint32_t n_bytes, uint32_t n_tiles, uint32_t tpb,
                                                              uint32_t* __restrict__ tile_count, uint16_t* __restrict__ tile_masks, uint32_t* block_base, uint32_t* counters)
{
    __shared__ uint32_t s_part[IDX_WARPS];
    __shared__ uint32_t s_last;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const uint32_t t0 = blockIdx.x * tpb, t1 = min(t0 + tpb, n_tiles);
    uint32_t mine = 0;
    for (uint32_t tile = t0 + wid; tile < t1; tile += IDX_WARPS) {
        uint32_t m[WT_CHUNKS];
        load_warp_tile(text, n_bytes, tile, lane, m);
        uint32_t c = 0;
#pragma unroll
        for (int j = 0; j < WT_CHUNKS; ++j) {
            c += __popc(m[j]);
            tile_masks[(size_t)tile * (WT_CHUNKS * 32) + j * 32 + lane] = (uint16_t)m[j]; // in text order: k_vcf_index never reads the text
        }
        c = __reduce_add_sync(0xffffffffu, c);
        if (lane == 0) tile_count[tile] = c;
        mine += c;
    }
    if (lane == 0) s_part[wid] = mine;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t tot = 0;
#pragma unroll
        for (int w = 0; w < IDX_WARPS; ++w) tot += s_part[w];
        block_base[blockIdx.x] = tot;
        __threadfence();
        s_last = atomicAdd(&counters[C_DONE], 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    // exclusive scan of gridDim.x <= MAX_IDX_BLOCKS block totals by this block's 256 threads (8 consecutive values each)
    __shared__ uint32_t s_warp[IDX_WARPS];
    constexpr int PER = MAX_IDX_BLOCKS / (IDX_WARPS * 32);
    const uint32_t i0 = threadIdx.x * PER;
    uint32_t v[PER], sum = 0;
#pragma unroll
    for (int k = 0; k < PER; ++k) {
        v[k] = i0 + k < gridDim.x ? __ldcg(&block_base[i0 + k]) : 0u;
        sum += v[k];
    }
    uint32_t incl = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += t;
    }
    if (lane == 31) s_warp[wid] = incl;
    __syncthreads();
    uint32_t wbase = 0, total = 0;
#pragma unroll
    for (int w = 0; w < IDX_WARPS; ++w) {
        const uint32_t c = s_warp[w];
        if (w < wid) wbase += c;
        total += c;
    }
    uint32_t run = wbase + incl - sum;
#pragma unroll
    for (int k = 0; k < PER; ++k) {
        if (i0 + k < gridDim.x) block_base[i0 + k] = run;
        run += v[k];
    }
    if (threadIdx.x == 0) counters[C_NEWLINES] = total;
}

__global__ void __launch_bounds__(IDX_WARPS * 32) k_vcf_index(const uint16_t* __restrict__ tile_masks, uint32_t n_tiles, uint32_t tpb,
                                                              const uint32_t* __restrict__ tile_count, const uint32_t* __restrict__ block_base,
                                                              uint32_t max_records, uint32_t* __restrict__ line_end)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const uint32_t t0 = blockIdx.x * tpb, t1 = min(t0 + tpb, n_tiles);
    const uint32_t bbase = block_base[blockIdx.x];
    if (bbase >= max_records) return;
    for (uint32_t tile = t0 + wid; tile < t1; tile += IDX_WARPS) {
        // line feeds of this block's tiles before `tile`
        uint32_t before = 0;
        for (uint32_t t = t0 + lane; t < tile; t += 32) before += tile_count[t];
        const uint32_t idx = bbase + __reduce_add_sync(0xffffffffu, before);
        if (idx >= max_records) return;
        // this lane's eight 16-byte chunks are contiguous in the text: masks [8 * lane, 8 * lane + 8) of the tile
        const uint4 mk = __ldg(reinterpret_cast<const uint4*>(tile_masks + (size_t)tile * (WT_CHUNKS * 32)) + lane);
        const uint32_t w[4] = {mk.x, mk.y, mk.z, mk.w};
        const uint32_t cnt = __popc(mk.x) + __popc(mk.y) + __popc(mk.z) + __popc(mk.w);
        if (!__ballot_sync(0xffffffffu, cnt != 0)) continue;
        uint32_t incl = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += t;
        }
        uint32_t k = idx + incl - cnt;
        const uint32_t off = tile * (uint32_t)WT_BYTES + (uint32_t)lane * 128u;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            uint32_t mm = w[q];
            while (mm) {
                const int b = __ffs(mm) - 1;
                mm &= mm - 1;
                if (k < max_records) line_end[k] = off + (uint32_t)q * 32u + (uint32_t)b;
                ++k;
            }
        }
    }
}

// ---- columns and genotypes -----------------------------------------------------------------------------------------
struct Hdr {
    int32_t n_allele;
    uint32_t amap; // nibble i = ACGT code of allele i, 0xE invalid / none
    int32_t gt_idx;
    int32_t st;
    int64_t pos;
};

__device__ __forceinline__ void raise(int& st, int code) { st = min(st, code); } // st starts at 99 = no error

// general GT sub-field parser of one sample column starting at q (htslib/vcf.c:2643-2673, 2726-2738); returns the packed byte
__device__ __noinline__ uint32_t parse_sample_general(const uint8_t* __restrict__ text, uint32_t q, uint32_t le, int gt_idx, int n_allele,
                                                       uint32_t amap, int& st, int& asum)
{
    int j = 0;
    while (j < gt_idx && q < le) {
        const uint32_t c = text[q];
        if (c == '\t') break;
        if (c == ':') ++j;
        ++q;
    }
    if (j < gt_idx) { // the column has no GT sub-field: missing + vector_end in the reference
        raise(st, VGL_IN_EPLOIDY);
        return 0xFF;
    }
    int n = 0, h0 = -1, h1 = -1;
    bool bad = false;
    for (;;) {
        uint32_t c = q < le ? text[q] : '\t';
        int val;
        if (c == '.') {
            val = -1;
            ++q;
        } else {
            const uint32_t q0 = q;
            if (c == '+') ++q;
            long long v = 0;
            while (q < le) {
                c = text[q];
                if (c < '0' || c > '9') break;
                if (v < (1ll << 40)) v = v * 10 + (int)(c - '0');
                ++q;
            }
            if (q == q0) bad = true;
            val = v > 1000 ? 1000 : (int)v;
        }
        if (n == 0) h0 = val;
        else if (n == 1) h1 = val;
        ++n;
        c = q < le ? text[q] : '\t';
        if (c == '|' || c == '/') {
            ++q;
            continue;
        }
        if (c != '\t' && c != ':') bad = true;
        break;
    }
    if (bad) {
        raise(st, VGL_IN_EGTCHAR);
        return 0xFF;
    }
    if (n != 2) {
        raise(st, VGL_IN_EPLOIDY);
        return 0xFF;
    }
    uint32_t b = 0;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const int h = k ? h1 : h0;
        uint32_t nib = 0xF;
        if (h >= 0) {
            if (h >= n_allele) raise(st, VGL_IN_EALLELEIDX);
            else {
                asum += h;
                const uint32_t mcode = h < 5 ? (amap >> (4 * h)) & 0xF : 0xE;
                if (mcode == 4) raise(st, VGL_IN_ESYMBOLIC);
                else if (mcode < 4) nib = mcode;
            }
        }
        b |= nib << (4 * k);
    }
    return b;
}

// the nine fixed columns; tab[] = positions of the first nine tabs of the line.  Executed uniformly by the whole warp.
template <class Text>
__device__ __forceinline__ Hdr parse_header(const Text& text, const uint32_t* tab, int gt_source)
{
    Hdr h;
    h.st = 99;
    // POS (htslib/vcf.c:3073-3083): hts_str2uint - 1
    {
        uint32_t p = tab[0] + 1;
        const uint32_t e = tab[1];
        if (p < e && text[p] == '+') ++p;
        unsigned long long v = 0;
        bool big = false;
        for (; p < e; ++p) {
            const uint32_t c = text[p];
            if (c < '0' || c > '9') break;
            if (v > (0xFFFFFFFFFFFFFFFFull - 9) / 10) big = true;
            else v = v * 10 + (c - '0');
        }
        if (big || v > 0x7FFFFFFFull) raise(h.st, VGL_IN_EPOS);
        h.pos = (long long)v - 1;
    }
    // REF, ALT (htslib/vcf.c:3087-3107) and the allele map (vcfgl.cpp:94-127)
    {
        int n_allele = 1;
        uint32_t amap = 0xEEEEEEEEu;
        auto allele = [&](int i, uint32_t t, uint32_t e) {
            uint32_t code = 0xE;
            const uint32_t len = e - t;
            const uint32_t c0 = len ? text[t] : 0;
            if (gt_source == VGL_SOURCE_BINARY) { // only the first character is looked at (vcfgl.cpp:112)
                if (c0 == '0') code = 0;
                else if (c0 == '1') code = 1;
            } else if (len == 1) {
                code = c0 == 'A' ? 0 : c0 == 'C' ? 1 : c0 == 'G' ? 2 : c0 == 'T' ? 3 : 0xE;
            } else if (len == 3) {
                if (c0 == '<' && text[t + 1] == '*' && text[t + 2] == '>') code = 4;
            } else if (len == 9) {
                const char* nr = "<NON_REF>";
                bool eq = true;
                for (int k = 0; k < 9; ++k) eq = eq && text[t + k] == (uint8_t)nr[k];
                if (eq) code = 4;
            }
            if (code == 0xE) raise(h.st, VGL_IN_EALLELE);
            amap = (amap & ~(0xFu << (4 * i))) | (code << (4 * i));
        };
        allele(0, tab[2] + 1, tab[3]);
        const uint32_t alt0 = tab[3] + 1, alt1 = tab[4];
        if (!((alt1 - alt0 == 1) && text[alt0] == '.')) {
            uint32_t t = alt0;
            for (;;) {
                uint32_t r = t;
                while (r < alt1 && text[r] != ',') ++r;
                if (n_allele < 5) allele(n_allele, t, r);
                ++n_allele;
                if (r >= alt1 || n_allele > 64) break;
                t = r + 1;
            }
        }
        h.n_allele = n_allele;
        h.amap = amap;
        if (n_allele > 5 || (gt_source == VGL_SOURCE_BINARY && n_allele > 2)) raise(h.st, VGL_IN_ENALLELE);
    }
    // FORMAT: index of the GT key (htslib/vcf.c:2455-2493)
    {
        int gt_idx = -1, j = 0;
        uint32_t t = tab[7] + 1;
        const uint32_t e = tab[8];
        for (uint32_t r = t;; ++r) {
            if (r == e || text[r] == ':') {
                if (gt_idx < 0 && r - t == 2 && text[t] == 'G' && text[t + 1] == 'T') gt_idx = j;
                ++j;
                t = r + 1;
            }
            if (r >= e) break;
        }
        h.gt_idx = gt_idx;
        if (gt_idx < 0) raise(h.st, VGL_IN_ENOGT);
    }
    return h;
}

__constant__ uint32_t c_pow10[9] = {1u, 10u, 100u, 1000u, 10000u, 100000u, 1000000u, 10000000u, 100000000u};

// The common shape of a record: REF and ALT one character each, FORMAT exactly "GT", POS of 1..9 digits.  Same results as
// parse_header(); anything else (and every defect) returns false and takes the general parser.  Warp-uniform.
__device__ __forceinline__ bool fast_header(const uint8_t* __restrict__ text, const uint32_t* tab, int gt_source, int lane, Hdr& h)
{
    const uint32_t plen = tab[1] - tab[0] - 1;
    if (tab[3] - tab[2] != 2 || tab[4] - tab[3] != 2 || tab[8] - tab[7] != 3 || plen - 1u > 8u) return false;
    const uint32_t ref = text[tab[2] + 1], alt = text[tab[3] + 1], f0 = text[tab[7] + 1], f1 = text[tab[7] + 2];
    const uint32_t c = (uint32_t)lane < plen ? (uint32_t)text[tab[0] + 1 + lane] - '0' : 0u; // one POS digit per lane
    uint32_t cr, ca;
    if (gt_source == VGL_SOURCE_BINARY) {
        cr = ref - '0';
        ca = alt - '0';
        if (cr > 1u) cr = 0xE;
        if (ca > 1u) ca = 0xE;
    } else {
        cr = ref == 'A' ? 0 : ref == 'C' ? 1 : ref == 'G' ? 2 : ref == 'T' ? 3 : 0xE;
        ca = alt == 'A' ? 0 : alt == 'C' ? 1 : alt == 'G' ? 2 : alt == 'T' ? 3 : 0xE;
    }
    const bool no_alt = alt == '.';
    const bool good = f0 == 'G' && f1 == 'T' && cr != 0xE && (no_alt || ca != 0xE) && c <= 9u;
    if (!__all_sync(0xffffffffu, good)) return false;
    const uint32_t v = __reduce_add_sync(0xffffffffu, (uint32_t)lane < plen ? c * c_pow10[plen - 1 - lane] : 0u);
    h.pos = (long long)v - 1;
    h.n_allele = no_alt ? 1 : 2;
    h.amap = 0xEEEEEE00u | cr | ((no_alt ? 0xEu : ca) << 4);
    h.gt_idx = 0;
    h.st = 99;
    return true;
}

// ---- msprime / tskit shaped records: FORMAT "GT", every sample column "a|b" + tab -----------------------------------------
// Per record a 16-byte descriptor, written by k_vcf_hdr (one THREAD per record: the nine fixed columns are ~40 bytes, a warp
// per record spends most of its issue slots idle on them), consumed by k_vcf_cells (one thread per four samples, records
// back to back: no per-record cost at all) and by k_vcf_gt, which finishes such records and parses every other one.
struct __align__(16) RecMeta {
    uint32_t p0;    // text offset of the first sample column
    uint32_t lut;   // FLAG_BIALLELIC: byte i = packed genotype of haplotypes (i & 1, i >> 1); else the allele nibble map
    uint32_t flags; // FLAG_*, n_allele << 8
    int32_t asum;   // allele-index sum, accumulated by k_vcf_cells
};
enum { FLAG_FIXED = 1,     // candidate: header clean, GT first, sample columns 4 * S - 1 bytes
       FLAG_FAILED = 2,    // some column was not "a|b": the general parser redoes the record
       FLAG_BIALLELIC = 4 };

__device__ __forceinline__ uint32_t slow_cell(uint32_t xv, uint32_t nal, uint32_t amap, bool& ok, int& sum)
{
    const uint32_t sep = xv & 0xFF00FF00u;
    const uint32_t c0 = xv & 0xFFu, c2 = (xv >> 16) & 0xFFu;
    const uint32_t d0 = c0 - '0', d1 = c2 - '0';
    const bool m0 = c0 == '.', m1 = c2 == '.', v0 = d0 < nal, v1 = d1 < nal;
    const uint32_t n0 = m0 ? 0xFu : (amap >> (4u * (d0 & 7u))) & 0xFu;
    const uint32_t n1 = m1 ? 0xFu : (amap >> (4u * (d1 & 7u))) & 0xFu;
    ok = ok && (sep == 0x09007C00u || sep == 0x09002F00u) && (m0 || v0) && (m1 || v1) && n0 < 4u + 12u * m0 && n1 < 4u + 12u * m1;
    sum += (v0 ? (int)d0 : 0) + (v1 ? (int)d1 : 0);
    return n0 | (n1 << 4);
}

// The common shape of a record for ONE thread (k_vcf_hdr): REF and ALT one character each, FORMAT exactly "GT", POS of 1..9
// digits.  Same results as parse_header(); anything else returns false and takes the general parser.
template <class Text>
__device__ __forceinline__ bool thread_fast_header(const Text& text, const uint32_t* tab, int gt_source, Hdr& h)
{
    const uint32_t plen = tab[1] - tab[0] - 1;
    if (tab[3] - tab[2] != 2 || tab[4] - tab[3] != 2 || tab[8] - tab[7] != 3 || plen - 1u > 8u) return false;
    const uint32_t ref = text[tab[2] + 1], alt = text[tab[3] + 1];
    if (text[tab[7] + 1] != 'G' || text[tab[7] + 2] != 'T') return false;
    uint32_t cr, ca;
    if (gt_source == VGL_SOURCE_BINARY) {
        cr = ref - '0', ca = alt - '0';
        if (cr > 1u) cr = 0xE;
        if (ca > 1u) ca = 0xE;
    } else {
        cr = ref == 'A' ? 0 : ref == 'C' ? 1 : ref == 'G' ? 2 : ref == 'T' ? 3 : 0xE;
        ca = alt == 'A' ? 0 : alt == 'C' ? 1 : alt == 'G' ? 2 : alt == 'T' ? 3 : 0xE;
    }
    const bool no_alt = alt == '.';
    if (cr == 0xE || (!no_alt && ca == 0xE)) return false;
    uint32_t v = 0;
    for (uint32_t p = tab[0] + 1; p < tab[1]; ++p) {
        const uint32_t d = text[p] - '0';
        if (d > 9u) return false;
        v = v * 10u + d;
    }
    h.pos = (long long)v - 1;
    h.n_allele = no_alt ? 1 : 2;
    h.amap = 0xEEEEEE00u | cr | ((no_alt ? 0xEu : ca) << 4);
    h.gt_idx = 0;
    h.st = 99;
    return true;
}

// the head of a record line, held in shared memory by the thread that parses the record
constexpr int HDR_THREADS = 128;
struct LineBuf {
    static constexpr int BYTES = 64;
    const uint8_t* g;
    uint32_t base; // text offset of word 0, 16-byte aligned
    uint32_t* w;   // word k at w[k * HDR_THREADS]
    __device__ __forceinline__ uint32_t operator[](uint32_t i) const
    {
        const uint32_t o = i - base;
        return o < (uint32_t)BYTES ? (w[(o >> 2) * HDR_THREADS] >> ((o & 3u) * 8u)) & 0xFFu : (uint32_t)g[i];
    }
};

__global__ void __launch_bounds__(HDR_THREADS) k_vcf_hdr(const uint8_t* __restrict__ text, const uint32_t* __restrict__ line_end, int32_t S, int32_t gt_source,
                                                 uint32_t max_records, uint32_t* counters, RecMeta* __restrict__ meta,
                                                 vgl_in_site* __restrict__ sites, uint32_t* __restrict__ work)
{
    __shared__ uint32_t s_head[LineBuf::BYTES / 4 * HDR_THREADS];
    const uint32_t n_rec = min(counters[C_NEWLINES], max_records);
    uint32_t n_fixed = 0;
    for (uint32_t line = blockIdx.x * blockDim.x + threadIdx.x; line < n_rec; line += gridDim.x * blockDim.x) {
        const uint32_t ls = line ? line_end[line - 1] + 1 : 0;
        uint32_t le = line_end[line];
        // the head of the line: its address depends on ls only, so the loads go out before anything that waits on le
        LineBuf buf;
        buf.g = text;
        buf.base = ls & ~15u;
        buf.w = s_head + threadIdx.x; // word k of this thread at s_head[k * HDR_THREADS + tid]: conflict-free
        uint4 hv[LineBuf::BYTES / 16];
#pragma unroll
        for (int k = 0; k < LineBuf::BYTES / 16; ++k) hv[k] = __ldg(reinterpret_cast<const uint4*>(text + buf.base) + k);
        const uint32_t body = 4u * (uint32_t)S - 1u;
        const uint32_t c_cr = le > ls ? text[le - 1] : 0u, c_tab = le > body + 1u ? text[le - body - 1u] : 0u, c_tab_cr = le > body + 2u ? text[le - body - 2u] : 0u;
        bool tab_ok = c_tab == '\t';
        if (c_cr == '\r') --le, tab_ok = c_tab_cr == '\t'; // KS_SEP_LINE strips the CR of a CRLF
        RecMeta m;
        m.p0 = 0, m.lut = 0, m.flags = 0, m.asum = 0;
        // the sample columns of a candidate are 4 * S - 1 bytes: the ninth tab must sit right before them
        if (le - ls > body + 16u && tab_ok) {
            uint32_t tab[9];
            int nt = 0;
            const uint32_t p0 = le - body;
#pragma unroll
            for (int k = 0; k < LineBuf::BYTES / 16; ++k) {
                const uint4 v = hv[k];
                buf.w[(4 * k) * HDR_THREADS] = v.x, buf.w[(4 * k + 1) * HDR_THREADS] = v.y;
                buf.w[(4 * k + 2) * HDR_THREADS] = v.z, buf.w[(4 * k + 3) * HDR_THREADS] = v.w;
                uint32_t tm = eq_mask16(v, 0x09090909u);
                const uint32_t a = buf.base + 16u * k;
                if (a < ls) tm &= ~((1u << (ls - a)) - 1u);
                while (tm && nt < 9) {
                    const uint32_t p = a + (uint32_t)__ffs(tm) - 1u;
                    tm &= tm - 1;
                    if (p < p0) tab[nt++] = p;
                }
            }
            for (uint32_t p = buf.base + LineBuf::BYTES; p < p0 && nt < 9; ++p)
                if (text[p] == '\t') tab[nt++] = p;
            if (nt == 9 && tab[8] == p0 - 1) {
                Hdr h;
                if (!thread_fast_header(buf, tab, gt_source, h)) h = parse_header(buf, tab, gt_source);
                bool sym = false;
                for (int i = 0; i < 5; ++i) sym = sym || (i < h.n_allele && ((h.amap >> (4 * i)) & 0xFu) == 4u);
                if (h.st == 99 && h.gt_idx == 0 && !sym) {
                    m.p0 = p0;
                    m.flags = FLAG_FIXED | ((uint32_t)h.n_allele << 8);
                    if (h.n_allele <= 2) {
                        const uint32_t r = h.amap & 0xFu, a = (h.amap >> 4) & 0xFu;
                        m.lut = (r | (r << 4)) | ((a | (r << 4)) << 8) | ((r | (a << 4)) << 16) | ((a | (a << 4)) << 24);
                        m.flags |= FLAG_BIALLELIC;
                    } else m.lut = h.amap;
                    vgl_in_site o;
                    o.status = VGL_IN_OK, o.skip_code = 0, o.pos = h.pos, o.allele_sum = 0, o.line_off = ls, o.line_len = le - ls;
                    o.n_allele = h.n_allele;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const uint32_t c = i < 5 && i < h.n_allele ? (h.amap >> (4 * i)) & 0xFu : 0xEu;
                        o.allele_acgt[i] = c == 0xEu ? -1 : (int8_t)c;
                    }
                    o.id_off = tab[1] + 1 - ls, o.fmt_off = tab[7] + 1 - ls, o.samples_off = p0 - ls, o._pad = 0;
                    sites[line] = o;
                }
            }
        }
        meta[line] = m;
        if (m.flags & FLAG_FIXED) ++n_fixed;
        else work[atomicAdd(&counters[C_NWORK], 1u)] = line; // for the general parser (k_vcf_gt)
    }
    // candidates count as kept records up front (k_vcf_cells takes a failed one back, k_vcf_gt applies --rm-invar-sites 1 | 2)
    n_fixed = __reduce_add_sync(0xffffffffu, n_fixed);
    if ((threadIdx.x & 31) == 0 && n_fixed) atomicAdd(&counters[C_NKEPT], n_fixed);
}

// four columns of a candidate record (samples s0 .. s0 + 3) -> four packed bytes
template <bool SUM>
__device__ __forceinline__ uint32_t cells4(uint32_t x0, uint32_t x1, uint32_t x2, uint32_t x3, int s0, int S, const RecMeta& m, bool& ok, int& sum)
{
    const int n = min(4, S - s0);
    if (s0 + 4 >= S) { // the record's last columns: the line end closes the last one, columns beyond it read as "0|0"
        const int last = S - 1 - s0;
        x0 = last == 0 ? (x0 & 0x00FFFFFFu) | 0x09000000u : x0;
        x1 = last == 1 ? (x1 & 0x00FFFFFFu) | 0x09000000u : last < 1 ? 0x09307C30u : x1;
        x2 = last == 2 ? (x2 & 0x00FFFFFFu) | 0x09000000u : last < 2 ? 0x09307C30u : x2;
        x3 = last == 3 ? (x3 & 0x00FFFFFFu) | 0x09000000u : last < 3 ? 0x09307C30u : x3;
    }
    const uint32_t nal = m.flags >> 8;
    const uint32_t dmask = nal == 2 ? 0xFFFEFFFEu : 0xFFFFFFFFu;
    const uint32_t d0 = (x0 & 0x00FF00FFu) - 0x00300030u, d1 = (x1 & 0x00FF00FFu) - 0x00300030u;
    const uint32_t d2 = (x2 & 0x00FF00FFu) - 0x00300030u, d3 = (x3 & 0x00FF00FFu) - 0x00300030u;
    // all four columns "a|b" + tab with a, b in {0, 1} (0 only without an ALT allele)?  one test, one table look-up each
    const uint32_t bad = ((x0 ^ 0x09007C00u) | (x1 ^ 0x09007C00u) | (x2 ^ 0x09007C00u) | (x3 ^ 0x09007C00u)) & 0xFF00FF00u;
    const uint32_t badd = (d0 | d1 | d2 | d3) & dmask;
    if ((m.flags & FLAG_BIALLELIC) && (bad | badd) == 0) {
        const uint32_t b0 = __byte_perm(m.lut, 0, ((d0 | (d0 >> 15)) & 3u) | 0x4440u);
        const uint32_t b1 = __byte_perm(m.lut, 0, ((d1 | (d1 >> 15)) & 3u) | 0x4440u);
        const uint32_t b2 = __byte_perm(m.lut, 0, ((d2 | (d2 >> 15)) & 3u) | 0x4440u);
        const uint32_t b3 = __byte_perm(m.lut, 0, ((d3 | (d3 >> 15)) & 3u) | 0x4440u);
        if (SUM) sum += __popc(d0 | (d1 << 1) | (d2 << 2) | (d3 << 3));
        return b0 | (b1 << 8) | (b2 << 16) | (b3 << 24);
    }
    // '/' separators, missing alleles, more than two alleles, or a defect
    const uint32_t amap = (m.flags & FLAG_BIALLELIC) ? 0xEEEEEE00u | (m.lut & 0xFu) | (((m.lut >> 8) & 0xFu) << 4) : m.lut;
    uint32_t out = slow_cell(x0, nal, amap, ok, sum);
    if (n > 1) out |= slow_cell(x1, nal, amap, ok, sum) << 8;
    if (n > 2) out |= slow_cell(x2, nal, amap, ok, sum) << 16;
    if (n > 3) out |= slow_cell(x3, nal, amap, ok, sum) << 24;
    return out;
}

// One thread per eight samples of a candidate record; the groups of all records are laid end to end (G = ceil(S / 8) per
// record) and every warp takes a contiguous run of them, so its (record, group) position advances without divisions.
// SUM: also accumulate the allele-index sum per record (only --rm-invar-sites 1 / 2 needs it, vcfgl.cpp:150-160).
template <bool SUM>
__global__ void __launch_bounds__(256) k_vcf_cells(const uint8_t* __restrict__ text, int32_t S, uint32_t G, uint32_t magic, uint32_t max_records,
                                                   uint32_t* counters, RecMeta* meta, uint8_t* __restrict__ rows, uint32_t* __restrict__ work)
{
    const uint32_t n_rec = min(counters[C_NEWLINES], max_records);
    const uint32_t total = n_rec * G; // the host guarantees max_records * G < 2^32
    const int lane = threadIdx.x & 31;
    const uint32_t n_warps = gridDim.x * (blockDim.x >> 5), warp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const uint32_t per_warp = ((total + n_warps - 1) / n_warps + 31u) & ~31u; // groups per warp, whole steps of 32
    const unsigned long long first = (unsigned long long)warp * per_warp;
    if (first >= total) return;
    const uint32_t g_end = (uint32_t)min((unsigned long long)total, first + per_warp);
    uint32_t line0 = (uint32_t)first / G, r0 = (uint32_t)first - line0 * G; // record and group of lane 0
    const bool rows_al = (S & 3) == 0, rows_al8 = (S & 7) == 0;
    const int n_seg = G >= 8 ? (int)(31u / G) + 2 : 0; // records a warp step can touch (0: too many, per-lane atomics instead)
    for (uint32_t g0 = (uint32_t)first; g0 < g_end; g0 += 32) {
        const uint32_t r = r0 + (uint32_t)lane;
        const uint32_t dl = G == 1 ? r : __umulhi(r, magic); // r / G, exact for r < G + 32
        const uint32_t line = line0 + dl, g = r - dl * G;
        const bool valid = g0 + (uint32_t)lane < g_end;
        {   // lane 0's position for the next step
            const uint32_t rn = r0 + 32u, q = G == 1 ? rn : __umulhi(rn, magic);
            line0 += q;
            r0 = rn - q * G;
        }
        RecMeta m;
        m.p0 = 0, m.lut = 0, m.flags = 0, m.asum = 0;
        if (valid) m = *reinterpret_cast<const RecMeta*>(__builtin_assume_aligned(&meta[line], 16));
        const bool fixed = valid && (m.flags & FLAG_FIXED);
        bool ok = true;
        int sum = 0;
        if (fixed) {
            const int s0 = 8 * (int)g;
            const uint32_t q = m.p0 + 32u * g, sh = (q & 3u) * 8u;
            const uint32_t* w = reinterpret_cast<const uint32_t*>(text + (q & ~3u));
            // nine independent loads: 32 bytes of columns per lane in flight
            const uint32_t w0 = __ldg(w), w1 = __ldg(w + 1), w2 = __ldg(w + 2), w3 = __ldg(w + 3), w4 = __ldg(w + 4);
            const bool second = s0 + 4 < S;
            const uint32_t w5 = second ? __ldg(w + 5) : 0u, w6 = second ? __ldg(w + 6) : 0u, w7 = second ? __ldg(w + 7) : 0u;
            const uint32_t w8 = second && sh ? __ldg(w + 8) : 0u;
            const uint32_t lo = cells4<SUM>(__funnelshift_r(w0, w1, sh), __funnelshift_r(w1, w2, sh), __funnelshift_r(w2, w3, sh), __funnelshift_r(w3, w4, sh),
                                            s0, S, m, ok, sum);
            uint32_t hi = 0;
            if (second)
                hi = cells4<SUM>(__funnelshift_r(w4, w5, sh), __funnelshift_r(w5, w6, sh), __funnelshift_r(w6, w7, sh), __funnelshift_r(w7, w8, sh),
                                 s0 + 4, S, m, ok, sum);
            uint8_t* dst = rows + (size_t)line * S + s0;
            if (rows_al8) *reinterpret_cast<uint2*>(dst) = make_uint2(lo, hi);
            else if (rows_al) {
                *reinterpret_cast<uint32_t*>(dst) = lo;
                if (second) *reinterpret_cast<uint32_t*>(dst + 4) = hi;
            } else {
                const int n = min(8, S - s0);
                for (int k = 0; k < n; ++k) dst[k] = (uint8_t)((k < 4 ? lo : hi) >> (8 * (k & 3)));
            }
        }
        if (!ok && !(atomicOr(&meta[line].flags, (uint32_t)FLAG_FAILED) & FLAG_FAILED)) { // rare: back to the general parser
            work[atomicAdd(&counters[C_NWORK], 1u)] = line;
            atomicSub(&counters[C_NKEPT], 1u);
        }
        if (SUM) { // per-record allele-index sums: the lanes of one record are contiguous
            const int v = fixed ? sum : 0;
            if (n_seg) {
                for (int j = 0; j < n_seg; ++j) {
                    const int tot = __reduce_add_sync(0xffffffffu, dl == (uint32_t)j ? v : 0);
                    const uint32_t lanes = __ballot_sync(0xffffffffu, fixed && dl == (uint32_t)j);
                    if (tot && lane == __ffs(lanes) - 1) atomicAdd(&meta[line].asum, tot);
                }
            } else if (v) atomicAdd(&meta[line].asum, v);
        }
    }
}

constexpr int GT_WARPS = 8;

__global__ void __launch_bounds__(GT_WARPS * 32) k_vcf_gt(const uint8_t* __restrict__ text, uint32_t n_bytes, const uint32_t* __restrict__ line_end,
                                                         int32_t S, int32_t gt_source, int32_t rm_invar, uint32_t max_records,
                                                         vgl_in_site* __restrict__ sites, uint8_t* __restrict__ rows, uint32_t* counters,
                                                         const RecMeta* __restrict__ meta, const uint32_t* __restrict__ work)
{
    __shared__ uint32_t s_tab[GT_WARPS][12];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const uint32_t n_rec = min(counters[C_NEWLINES], max_records);
    const uint32_t warp0 = blockIdx.x * GT_WARPS + wid, n_warps = gridDim.x * GT_WARPS;
    if (warp0 == 0 && lane == 0) {
        counters[C_NRECORDS] = n_rec;
        counters[C_CONSUMED] = n_rec ? line_end[n_rec - 1] + 1 : 0;
    }
    uint32_t* tab = s_tab[wid];
    uint32_t n_kept = 0, n_err = 0, first_err = 0xFFFFFFFFu; // lane 0's tallies, posted once per warp
    // (a) fixed-width records that k_vcf_cells converted completely are finished (k_vcf_hdr wrote their site records and counted
    //     them as kept) unless --rm-invar-sites 1 | 2 asks for the skip decision: then one thread per record applies it
    uint32_t n_dropped = 0;
    if (rm_invar & 3) {
        for (uint32_t line = blockIdx.x * blockDim.x + threadIdx.x; line < n_rec; line += gridDim.x * blockDim.x) {
            const RecMeta m = meta[line];
            if ((m.flags & (FLAG_FIXED | FLAG_FAILED)) != FLAG_FIXED) continue;
            int skip = 0;
            const int nal = (int)(m.flags >> 8);
            if ((rm_invar & 1) && m.asum == 0) skip = -1;
            else if (rm_invar & 2)
                for (int al = 1; al < nal; ++al)
                    if ((long long)al * S * 2 == (long long)m.asum) skip = -2;
            sites[line].skip_code = skip;
            sites[line].allele_sum = m.asum;
            n_dropped += skip != 0;
        }
        n_dropped = __reduce_add_sync(0xffffffffu, n_dropped);
    }
    // (b) every other record: one warp per record, from the work list k_vcf_hdr and k_vcf_cells left
    const uint32_t n_work = counters[C_NWORK];
    for (uint32_t wi = warp0; wi < n_work; wi += n_warps) {
        const uint32_t line = work[wi];
        const uint32_t ls = line ? line_end[line - 1] + 1 : 0;
        uint32_t le = line_end[line];
        if (le > ls && text[le - 1] == '\r') --le; // KS_SEP_LINE strips the CR of a CRLF
        uint8_t* const row = rows + (size_t)line * S;
        int st = 99, asum = 0;
        int ntab = 0; // tabs before the current window
        bool have_hdr = false;
        Hdr h;
        h.n_allele = 0, h.amap = 0, h.gt_idx = -1, h.st = 99, h.pos = -1;
        for (uint32_t w = ls & ~15u; w < le; w += 512) {
            const uint32_t a = w + lane * 16;
            uint32_t tm = 0;
            if (a < le) {
                const uint4 v = __ldg(reinterpret_cast<const uint4*>(text + a));
                tm = eq_mask16(v, 0x09090909u);
                const uint32_t lo = ls > a ? ls - a : 0, hi = min(16u, le - a);
                tm &= ((1u << hi) - 1u) & ~((1u << lo) - 1u);
            }
            const int cnt = __popc(tm);
            int incl = cnt;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += t;
            }
            const int total = __shfl_sync(0xffffffffu, incl, 31);
            const int first = ntab + incl - cnt; // ordinal (0-based) of this lane's first tab
            if (!have_hdr) {
                if (first < 9) {
                    uint32_t m = tm;
                    for (int k = 0; m && first + k < 9; ++k) {
                        const int b = __ffs(m) - 1;
                        m &= m - 1;
                        tab[first + k] = a + b;
                    }
                }
                __syncwarp();
                if (ntab + total < 9) {
                    ntab += total;
                    continue;
                }
                if (!fast_header(text, tab, gt_source, lane, h)) h = parse_header(text, tab, gt_source);
                have_hdr = true;
                __syncwarp();
            }
            // sample columns: the tab with ordinal o >= 8 starts sample o - 8
            uint32_t m = tm;
            for (int o = first; m; ++o) {
                const int b = __ffs(m) - 1;
                m &= m - 1;
                const int s = o - 8;
                if (s < 0 || s >= S) continue;
                const uint32_t q = a + b + 1;
                uint32_t byte;
                bool fast = false;
                if (h.gt_idx == 0 && q + 3 <= le) {
                    const uint32_t c0 = text[q], c1 = text[q + 1], c2 = text[q + 2];
                    const uint32_t c3 = q + 3 < le ? text[q + 3] : '\t';
                    const uint32_t d0 = c0 - '0', d1 = c2 - '0';
                    const bool ok0 = d0 <= 9u || c0 == '.', ok1 = d1 <= 9u || c2 == '.';
                    if (ok0 && ok1 && (c1 == '|' || c1 == '/') && (c3 == '\t' || c3 == ':')) {
                        fast = true;
                        uint32_t n0 = 0xF, n1 = 0xF;
                        if (d0 <= 9u) {
                            if ((int)d0 >= h.n_allele) raise(st, VGL_IN_EALLELEIDX);
                            else {
                                asum += d0;
                                const uint32_t mc = d0 < 5 ? (h.amap >> (4 * d0)) & 0xF : 0xE;
                                if (mc == 4) raise(st, VGL_IN_ESYMBOLIC);
                                else if (mc < 4) n0 = mc;
                            }
                        }
                        if (d1 <= 9u) {
                            if ((int)d1 >= h.n_allele) raise(st, VGL_IN_EALLELEIDX);
                            else {
                                asum += d1;
                                const uint32_t mc = d1 < 5 ? (h.amap >> (4 * d1)) & 0xF : 0xE;
                                if (mc == 4) raise(st, VGL_IN_ESYMBOLIC);
                                else if (mc < 4) n1 = mc;
                            }
                        }
                        byte = n0 | (n1 << 4);
                    }
                }
                if (!fast) {
                    if (h.gt_idx >= 0) byte = parse_sample_general(text, q, le, h.gt_idx, h.n_allele, h.amap, st, asum);
                    else byte = 0xFF;
                }
                row[s] = (uint8_t)byte;
            }
            ntab += total;
        }
        // line-level results
        st = min(st, h.st);
        if (!have_hdr) st = VGL_IN_ENCOLS;
        else if (ntab - 8 < S) raise(st, VGL_IN_ENSAMPLES);
#pragma unroll
        for (int d = 16; d; d >>= 1) {
            st = min(st, __shfl_xor_sync(0xffffffffu, st, d));
            asum += __shfl_xor_sync(0xffffffffu, asum, d);
        }
        if (lane == 0) {
            vgl_in_site o;
            o.status = st == 99 ? VGL_IN_OK : st;
            o.skip_code = 0;
            if (o.status == VGL_IN_OK) {
                if ((rm_invar & 1) && asum == 0) o.skip_code = -1;
                else if (rm_invar & 2)
                    for (int al = 1; al < h.n_allele; ++al)
                        if ((long long)al * S * 2 == (long long)asum) o.skip_code = -2;
            }
            o.pos = have_hdr ? h.pos : 0;
            o.allele_sum = (rm_invar & 3) ? asum : 0; // only --rm-invar-sites 1 / 2 uses it
            o.line_off = ls;
            o.line_len = le - ls;
            o.n_allele = have_hdr ? h.n_allele : 0;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const uint32_t c = i < 5 && have_hdr && i < h.n_allele ? (h.amap >> (4 * i)) & 0xF : 0xE;
                o.allele_acgt[i] = c == 0xE ? -1 : (int8_t)c;
            }
            o.id_off = have_hdr ? tab[1] + 1 - ls : 0;
            o.fmt_off = have_hdr ? tab[7] + 1 - ls : 0;
            o.samples_off = have_hdr ? tab[8] + 1 - ls : 0;
            o._pad = 0;
            sites[line] = o;
            if (o.status != VGL_IN_OK) {
                ++n_err;
                first_err = min(first_err, line);
            } else if (o.skip_code == 0) ++n_kept;
        }
        __syncwarp();
    }
    if (lane == 0) {
        if (n_kept) atomicAdd(&counters[C_NKEPT], n_kept);
        if (n_dropped) atomicSub(&counters[C_NKEPT], n_dropped);
        if (n_err) {
            atomicAdd(&counters[C_NERRORS], n_err);
            atomicMin(&counters[C_FIRSTERR], first_err);
        }
    }
}

// ---- BCF records (binary input) -------------------------------------------------------------------------------------
// The same per-record work for uncompressed BCF records (what bcf_read + bcf_unpack + bcf_get_genotypes + check_rec_alleles
// do with them, htslib/vcf.c:1535-1600, vcfgl.cpp:75-163): fixed fields, allele strings -> allele map, the FORMAT block
// whose key is GT -> typed integer vector [n_sample][ploidy] -> packed bytes.  Eight lanes per record; the genotype vector is
// contiguous, so the lanes read it coalesced.  Record offsets come from the host (a chain of l_shared + l_indiv hops).
__device__ __forceinline__ uint32_t rd_u32(const uint8_t* __restrict__ t, uint32_t p)
{
    return (uint32_t)t[p] | ((uint32_t)t[p + 1] << 8) | ((uint32_t)t[p + 2] << 16) | ((uint32_t)t[p + 3] << 24);
}

// typed scalar / descriptor at p: returns (count, type) and advances p past the descriptor (BCF2 spec 6.3.3)
__device__ __forceinline__ void rd_desc(const uint8_t* __restrict__ t, uint32_t& p, uint32_t& n, uint32_t& type)
{
    const uint32_t b = t[p++];
    type = b & 0xFu;
    n = b >> 4;
    if (n == 15) { // the count follows as a typed integer
        const uint32_t tb = t[p++] & 0xFu;
        if (tb == 1) n = t[p], p += 1;
        else if (tb == 2) n = (uint32_t)t[p] | ((uint32_t)t[p + 1] << 8), p += 2;
        else n = rd_u32(t, p), p += 4;
    }
}

template <int LANES> // lanes per record: 8 (four records per warp in flight) for narrow records, 32 for wide ones
__global__ void __launch_bounds__(GT_WARPS * 32) k_bcf_gt(const uint8_t* __restrict__ text, const uint32_t* __restrict__ rec_off, uint32_t n_rec, int32_t S,
                                                         int32_t gt_source, int32_t gt_key, int32_t rm_invar, vgl_in_site* __restrict__ sites,
                                                         uint8_t* __restrict__ rows, uint32_t* counters)
{
    // eight lanes per record, four records per warp in flight: a record's fixed fields are a chain of dependent byte loads, and a
    // warp walking its records one after the other would wait on every link (cf. k_gvcf_key)
    const int lane = threadIdx.x & (LANES - 1);
    const uint32_t gmask = LANES == 32 ? 0xFFFFFFFFu : 0xFFu << (threadIdx.x & 24);
    const uint32_t grp0 = (blockIdx.x * blockDim.x + threadIdx.x) / LANES, n_grp = (gridDim.x * blockDim.x) / LANES;
    uint32_t n_kept = 0, n_err = 0, first_err = 0xFFFFFFFFu;
    for (uint32_t r = grp0; r < n_rec; r += n_grp) {
        const uint32_t o = rec_off[r], rec_len = rec_off[r + 1] - o;
        int st = 99, asum = 0;
        const uint32_t l_shared = rd_u32(text, o);
        const int32_t pos = (int32_t)rd_u32(text, o + 12);
        const uint32_t nai = rd_u32(text, o + 24), nfs = rd_u32(text, o + 28);
        const int n_allele = (int)(nai >> 16), n_fmt = (int)(nfs >> 24);
        const uint32_t n_sample = nfs & 0xFFFFFFu;
        uint32_t p = o + 32, n, type;
        rd_desc(text, p, n, type); // ID
        p += n;
        uint32_t amap = 0xEEEEEEEEu;
        for (int a = 0; a < n_allele; ++a) {
            rd_desc(text, p, n, type);
            if (a < 5) {
                uint32_t code = 0xE;
                const uint32_t c0 = n ? text[p] : 0;
                if (gt_source == VGL_SOURCE_BINARY) {
                    if (c0 == '0') code = 0;
                    else if (c0 == '1') code = 1;
                } else if (n == 1) {
                    code = c0 == 'A' ? 0 : c0 == 'C' ? 1 : c0 == 'G' ? 2 : c0 == 'T' ? 3 : 0xE;
                } else if (n == 3) {
                    if (c0 == '<' && text[p + 1] == '*' && text[p + 2] == '>') code = 4;
                } else if (n == 9) {
                    const char* nr = "<NON_REF>";
                    bool eq = true;
                    for (int k = 0; k < 9; ++k) eq = eq && text[p + k] == (uint8_t)nr[k];
                    if (eq) code = 4;
                }
                if (code == 0xE) raise(st, VGL_IN_EALLELE);
                amap = (amap & ~(0xFu << (4 * a))) | (code << (4 * a));
            }
            p += n;
        }
        if (n_allele > 5 || (gt_source == VGL_SOURCE_BINARY && n_allele > 2)) raise(st, VGL_IN_ENALLELE);
        if ((int)n_sample != S) raise(st, VGL_IN_ENSAMPLES);
        // FORMAT blocks: key, (values per sample, type), values
        p = o + 8 + l_shared;
        uint32_t gt_at = 0, gt_n = 0, gt_w = 0;
        bool have_gt = false;
        for (int f = 0; f < n_fmt && !have_gt; ++f) {
            rd_desc(text, p, n, type); // the key: one typed integer
            int key = type == 1 ? (int)(int8_t)text[p] : type == 2 ? (int)(int16_t)((uint32_t)text[p] | ((uint32_t)text[p + 1] << 8)) : (int)rd_u32(text, p);
            p += type == 1 ? 1 : type == 2 ? 2 : 4;
            rd_desc(text, p, n, type);
            const uint32_t w = type == 1 || type == 7 ? 1u : type == 2 ? 2u : 4u;
            if (key == gt_key) have_gt = true, gt_at = p, gt_n = n, gt_w = type == 7 ? 0u : w;
            p += n * w * n_sample;
        }
        uint8_t* const row = rows + (size_t)r * S;
        if (!have_gt || gt_w == 0) raise(st, VGL_IN_ENOGT);
        else if (gt_n != 2) raise(st, VGL_IN_EPLOIDY); // the reference reads gt_arr as [2 * n_samples] (vcfgl.cpp:131-146)
        else if ((int)n_sample == S) {
            for (int s = lane; s < S; s += LANES) {
                uint32_t byte = 0;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const uint32_t q = gt_at + (2u * (uint32_t)s + h) * gt_w;
                    int v;
                    bool end, missing_sentinel;
                    if (gt_w == 1) v = (int8_t)text[q], end = v == -127, missing_sentinel = v == -128;
                    else if (gt_w == 2) v = (int16_t)((uint32_t)text[q] | ((uint32_t)text[q + 1] << 8)), end = v == -32767, missing_sentinel = v == -32768;
                    else v = (int)rd_u32(text, q), end = v == (int)0x80000001, missing_sentinel = v == (int)0x80000000;
                    uint32_t nib = 0xF;
                    if (end) raise(st, VGL_IN_EPLOIDY);          // a haploid sample in a diploid record: vector_end
                    else if (missing_sentinel) raise(st, VGL_IN_EALLELEIDX); // not a GT value; the reference asserts a >= 0
                    else if ((v >> 1) != 0) {                    // bcf_gt_is_missing: (v >> 1) == 0
                        const int a = (v >> 1) - 1;
                        if (a < 0 || a >= n_allele) raise(st, VGL_IN_EALLELEIDX);
                        else {
                            asum += a;
                            const uint32_t mc = a < 5 ? (amap >> (4 * a)) & 0xF : 0xE;
                            if (mc == 4) raise(st, VGL_IN_ESYMBOLIC);
                            else if (mc < 4) nib = mc;
                        }
                    }
                    byte |= nib << (4 * h);
                }
                row[s] = (uint8_t)byte;
            }
        }
#pragma unroll
        for (int d = LANES / 2; d; d >>= 1) {
            st = min(st, __shfl_xor_sync(gmask, st, d));
            asum += __shfl_xor_sync(gmask, asum, d);
        }
        if (lane == 0) {
            vgl_in_site out;
            out.status = st == 99 ? VGL_IN_OK : st;
            out.skip_code = 0;
            if (out.status == VGL_IN_OK) {
                if ((rm_invar & 1) && asum == 0) out.skip_code = -1;
                else if (rm_invar & 2)
                    for (int al = 1; al < n_allele; ++al)
                        if ((long long)al * S * 2 == (long long)asum) out.skip_code = -2;
            }
            out.pos = pos;
            out.allele_sum = (rm_invar & 3) ? asum : 0;
            out.line_off = o;
            out.line_len = rec_len;
            out.n_allele = n_allele;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const uint32_t c = i < 5 && i < n_allele ? (amap >> (4 * i)) & 0xF : 0xE;
                out.allele_acgt[i] = c == 0xE ? -1 : (int8_t)c;
            }
            out.id_off = 32, out.fmt_off = 8 + l_shared, out.samples_off = have_gt ? gt_at - o : 0, out._pad = 0;
            sites[r] = out;
            if (out.status != VGL_IN_OK) ++n_err, first_err = min(first_err, r);
            else if (out.skip_code == 0) ++n_kept;
        }
    }
    if (lane == 0) {
        if (n_kept) atomicAdd(&counters[C_NKEPT], n_kept);
        if (n_err) {
            atomicAdd(&counters[C_NERRORS], n_err);
            atomicMin(&counters[C_FIRSTERR], first_err);
        }
    }
}

// ---- rows -> slot genotype matrix ----------------------------------------------------------------------------------
__global__ void k_place_rows(const uint8_t* __restrict__ rows, const int32_t* __restrict__ row_map, int32_t first_record, int32_t n_sites,
                             int32_t S, uint32_t fill, uint8_t* __restrict__ gt)
{
    const long long n = (long long)n_sites * S;
    const long long stride = (long long)gridDim.x * blockDim.x;
    if ((S & 15) == 0) { // rows are 16-byte aligned: move 16 cells per thread
        const int per = S >> 4;
        const long long nv = (long long)n_sites * per;
        const uint32_t f4 = fill * 0x01010101u;
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += stride) {
            const int r = (int)(i / per), c = (int)(i - (long long)r * per);
            const int src = row_map ? row_map[r] : first_record + r;
            uint4 v = make_uint4(f4, f4, f4, f4);
            if (src >= 0) v = __ldg(reinterpret_cast<const uint4*>(rows + (size_t)src * S) + c);
            reinterpret_cast<uint4*>(gt + (size_t)r * S)[c] = v;
        }
        return;
    }
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const int r = (int)(i / S), c = (int)(i - (long long)r * S);
        const int src = row_map ? row_map[r] : first_record + r;
        gt[i] = src >= 0 ? rows[(size_t)src * S + c] : (uint8_t)fill;
    }
}

} // namespace

void launch_place_rows(const uint8_t* rows, const int32_t* d_row_map, int32_t first_record, int32_t n_sites, int32_t S, uint8_t fill, uint8_t* gt,
                       cudaStream_t st, int n_sms)
{
    k_place_rows<<<n_sms * 8, 256, 0, st>>>(rows, d_row_map, first_record, n_sites, S, fill, gt);
}

int parser_create(int device, int S, int rm_invar, int n_sms, int64_t max_text, int32_t max_records, vgl_parser** out, std::string& err)
{
    *out = nullptr;
    if (max_text < 1 || max_text >= (int64_t)0xFFFF0000ll || max_records < 1 || S < 1) {
        err = "vgl_parser_create: max_text_bytes must be in [1, 4 GiB), max_records >= 1";
        return VGL_EINVAL;
    }
    if ((uint64_t)max_records * (((uint64_t)S + 3) / 4) >= 0xFFFFFFFFull) {
        err = "vgl_parser_create: max_records * n_samples / 4 must stay below 2^32";
        return VGL_EINVAL;
    }
    vgl_parser* ps = new (std::nothrow) vgl_parser();
    if (!ps) return VGL_ENOMEM;
    ps->device = device, ps->S = S, ps->rm_invar = rm_invar & 3, ps->n_sms = n_sms;
    ps->max_records = max_records;
    ps->text_cap = (size_t)max_text;
    const size_t padded = ((ps->text_cap + 1 + WT_BYTES - 1) / WT_BYTES) * WT_BYTES + 1024;
    ps->max_tiles = (uint32_t)(padded / WT_BYTES);
#define PCK(call)                                                             \
    do {                                                                      \
        cudaError_t e_ = (call);                                              \
        if (e_ != cudaSuccess) {                                              \
            err = std::string(#call) + ": " + cudaGetErrorString(e_);         \
            parser_destroy(ps);                                               \
            return e_ == cudaErrorMemoryAllocation ? VGL_ENOMEM : VGL_ECUDA;  \
        }                                                                     \
    } while (0)
    PCK(cudaSetDevice(device));
    PCK(cudaStreamCreateWithFlags(&ps->stream, cudaStreamNonBlocking));
    for (auto& e : ps->ev) PCK(cudaEventCreate(&e));
    PCK(cudaEventCreateWithFlags(&ps->ev_done, cudaEventDisableTiming));
    PCK(cudaEventCreateWithFlags(&ps->ev_placed, cudaEventDisableTiming));
    PCK(cudaHostAlloc((void**)&ps->h_text, ps->text_cap + 1, cudaHostAllocDefault));
    PCK(cudaMalloc((void**)&ps->d_text, padded));
    PCK(cudaMemset(ps->d_text, 0, padded));
    PCK(cudaMalloc((void**)&ps->d_line_end, ((size_t)max_records + 1) * sizeof(uint32_t)));
    PCK(cudaMalloc((void**)&ps->d_tile_count, (size_t)ps->max_tiles * sizeof(uint32_t)));
    PCK(cudaMalloc((void**)&ps->d_tile_masks, (size_t)ps->max_tiles * (WT_CHUNKS * 32) * sizeof(uint16_t)));
    PCK(cudaMalloc((void**)&ps->d_block_base, (size_t)MAX_IDX_BLOCKS * sizeof(uint32_t)));
    PCK(cudaMalloc((void**)&ps->d_work, (size_t)max_records * sizeof(uint32_t)));
    PCK(cudaMalloc((void**)&ps->d_counters, C_COUNT * sizeof(uint32_t)));
    PCK(cudaHostAlloc((void**)&ps->h_counters, C_COUNT * sizeof(uint32_t), cudaHostAllocDefault));
    PCK(cudaMalloc((void**)&ps->d_sites, (size_t)max_records * sizeof(vgl_in_site)));
    PCK(cudaHostAlloc((void**)&ps->h_sites, (size_t)max_records * sizeof(vgl_in_site), cudaHostAllocDefault));
    PCK(cudaMalloc((void**)&ps->d_rows, (size_t)max_records * S + 16));
    PCK(cudaMalloc((void**)&ps->d_row_map, (size_t)max_records * sizeof(int32_t)));
    PCK(cudaMalloc((void**)&ps->d_meta, (size_t)max_records * sizeof(RecMeta)));
#undef PCK
    *out = ps;
    return VGL_OK;
}

void parser_destroy(vgl_parser* ps)
{
    if (!ps) return;
    cudaSetDevice(ps->device);
    if (ps->stream) cudaStreamSynchronize(ps->stream);
    cudaFreeHost(ps->h_text);
    cudaFree(ps->d_text);
    cudaFree(ps->d_line_end);
    cudaFree(ps->d_tile_count);
    cudaFree(ps->d_block_base);
    cudaFree(ps->d_tile_masks);
    cudaFree(ps->d_work);
    cudaFree(ps->d_counters);
    cudaFreeHost(ps->h_counters);
    cudaFree(ps->d_sites);
    cudaFreeHost(ps->h_sites);
    cudaFree(ps->d_rows);
    cudaFree(ps->d_row_map);
    cudaFree(ps->d_meta);
    for (auto& e : ps->ev)
        if (e) cudaEventDestroy(e);
    if (ps->ev_done) cudaEventDestroy(ps->ev_done);
    if (ps->ev_placed) cudaEventDestroy(ps->ev_placed);
    if (ps->stream) cudaStreamDestroy(ps->stream);
    delete ps;
}

} // namespace vgl

using namespace vgl;

#define PCK(call)                                                        \
    do {                                                                 \
        cudaError_t e_ = (call);                                         \
        if (e_ != cudaSuccess) {                                         \
            ps->err = std::string(#call) + ": " + cudaGetErrorString(e_); \
            return VGL_ECUDA;                                            \
        }                                                                \
    } while (0)

extern "C" void vgl_parser_destroy(vgl_parser* ps) { parser_destroy(ps); }

extern "C" int vgl_parser_text_buffer(vgl_parser* ps, uint8_t** text, int64_t* capacity)
{
    if (!ps) return VGL_EINVAL;
    if (text) *text = ps->h_text;
    if (capacity) *capacity = (int64_t)ps->text_cap;
    return VGL_OK;
}

extern "C" int vgl_parse_vcf(vgl_parser* ps, int64_t n_bytes, int32_t gt_source, uint32_t flags, vgl_parse_out* out)
{
    if (!ps || !out || n_bytes < 0 || (size_t)n_bytes > ps->text_cap || gt_source < 0 || gt_source > 1) return VGL_EINVAL;
    memset(out, 0, sizeof *out);
    out->first_error_record = -1;
    out->sites = ps->h_sites;
    ps->n_records = 0;
    if (n_bytes == 0) return VGL_OK;
    PCK(cudaSetDevice(ps->device));
    cudaStream_t st = ps->stream;
    size_t n = (size_t)n_bytes;
    const bool on_device = (flags & VGL_PARSE_TEXT_ON_DEVICE) != 0;
    if (on_device) {
        if (ps->d_text_bytes == 0) {
            ps->err = "VGL_PARSE_TEXT_ON_DEVICE: no text on the device yet";
            return VGL_ESTATE;
        }
        n = ps->d_text_bytes;
    } else if ((flags & VGL_PARSE_FINAL) && ps->h_text[n - 1] != '\n') {
        ps->h_text[n++] = '\n'; // the staging buffer has one spare byte
    }
    if (ps->placed) PCK(cudaStreamWaitEvent(st, ps->ev_placed, 0)); // rows / row map of the previous parse may still be read
    PCK(cudaEventRecord(ps->ev[0], st));
    if (!on_device) {
        PCK(cudaMemcpyAsync(ps->d_text, ps->h_text, n, cudaMemcpyHostToDevice, st));
        ps->d_text_bytes = n;
    }
    PCK(cudaEventRecord(ps->ev[1], st));
    const uint32_t n_tiles = (uint32_t)((n + WT_BYTES - 1) / WT_BYTES);
    static const uint32_t init[C_COUNT] = {0, 0, 0, 0, 0xFFFFFFFFu, 0, 0, 0};
    PCK(cudaMemcpyAsync(ps->d_counters, init, sizeof init, cudaMemcpyHostToDevice, st));
    const uint32_t idx_grid = (uint32_t)std::min<uint64_t>((n_tiles + IDX_WARPS - 1) / IDX_WARPS, (uint64_t)std::min(ps->n_sms * 8, (int)MAX_IDX_BLOCKS));
    const uint32_t tpb = (n_tiles + idx_grid - 1) / idx_grid; // warp tiles per block, contiguous
    k_vcf_count<<<idx_grid, IDX_WARPS * 32, 0, st>>>(ps->d_text, (uint32_t)n, n_tiles, tpb, ps->d_tile_count, ps->d_tile_masks, ps->d_block_base, ps->d_counters);
    k_vcf_index<<<idx_grid, IDX_WARPS * 32, 0, st>>>(ps->d_tile_masks, n_tiles, tpb, ps->d_tile_count, ps->d_block_base, (uint32_t)ps->max_records, ps->d_line_end);
    RecMeta* meta = reinterpret_cast<RecMeta*>(ps->d_meta);
    k_vcf_hdr<<<ps->n_sms * 8, HDR_THREADS, 0, st>>>(ps->d_text, ps->d_line_end, ps->S, gt_source, (uint32_t)ps->max_records, ps->d_counters, meta, ps->d_sites, ps->d_work);
    {   // k_vcf_cells: one thread per eight samples of every record that can be fixed-width (each is at least 4 * S bytes long)
        const uint32_t G = ((uint32_t)ps->S + 7u) / 8u;
        const uint32_t magic = G > 1 ? (uint32_t)((0x100000000ull + G - 1) / G) : 0u;
        const uint64_t max_groups = std::min<uint64_t>((uint64_t)ps->max_records, n / (4ull * (uint64_t)ps->S) + 1) * G;
        const uint32_t cells_grid = std::max(1u, (uint32_t)std::min<uint64_t>((max_groups + 255) / 256, (uint64_t)ps->n_sms * 16));
        if (ps->rm_invar & 3)
            k_vcf_cells<true><<<cells_grid, 256, 0, st>>>(ps->d_text, ps->S, G, magic, (uint32_t)ps->max_records, ps->d_counters, meta, ps->d_rows, ps->d_work);
        else
            k_vcf_cells<false><<<cells_grid, 256, 0, st>>>(ps->d_text, ps->S, G, magic, (uint32_t)ps->max_records, ps->d_counters, meta, ps->d_rows, ps->d_work);
    }
    k_vcf_gt<<<ps->n_sms * 8, GT_WARPS * 32, 0, st>>>(ps->d_text, (uint32_t)n, ps->d_line_end, ps->S, gt_source, ps->rm_invar, (uint32_t)ps->max_records,
                                                     ps->d_sites, ps->d_rows, ps->d_counters, meta, ps->d_work);
    ps->launches += 5;
    PCK(cudaGetLastError());
    PCK(cudaEventRecord(ps->ev[2], st));
    PCK(cudaMemcpyAsync(ps->h_counters, ps->d_counters, C_COUNT * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    PCK(cudaStreamSynchronize(st));
    const uint32_t n_rec = ps->h_counters[C_NRECORDS];
    if (n_rec) {
        PCK(cudaMemcpyAsync(ps->h_sites, ps->d_sites, (size_t)n_rec * sizeof(vgl_in_site), cudaMemcpyDeviceToHost, st));
        PCK(cudaStreamSynchronize(st));
    }
    PCK(cudaEventRecord(ps->ev_done, st));
    ps->n_records = (int32_t)n_rec;
    out->n_records = (int32_t)n_rec;
    out->n_errors = (int32_t)ps->h_counters[C_NERRORS];
    out->first_error_record = out->n_errors ? (int32_t)ps->h_counters[C_FIRSTERR] : -1;
    out->n_kept = (int32_t)ps->h_counters[C_NKEPT];
    out->bytes_consumed = std::min<int64_t>((int64_t)ps->h_counters[C_CONSUMED], n_bytes);
    cudaEventElapsedTime(&out->ms_h2d, ps->ev[0], ps->ev[1]);
    cudaEventElapsedTime(&out->ms_kernels, ps->ev[1], ps->ev[2]);
    return VGL_OK;
}

extern "C" int vgl_parser_rows(vgl_parser* ps, int32_t first_record, int32_t n_records, uint8_t* host_dst)
{
    if (!ps || !host_dst || first_record < 0 || n_records < 0 || first_record + n_records > ps->n_records) return VGL_EINVAL;
    if (n_records == 0) return VGL_OK;
    PCK(cudaSetDevice(ps->device));
    PCK(cudaMemcpyAsync(host_dst, ps->d_rows + (size_t)first_record * ps->S, (size_t)n_records * ps->S, cudaMemcpyDeviceToHost, ps->stream));
    PCK(cudaStreamSynchronize(ps->stream));
    return VGL_OK;
}

extern "C" int vgl_parse_bcf(vgl_parser* ps, int64_t n_bytes, const uint32_t* rec_off, int32_t n_records, int32_t gt_source, int32_t gt_key, uint32_t flags,
                             vgl_parse_out* out)
{
    if (!ps || !out || !rec_off || n_bytes < 0 || (size_t)n_bytes > ps->text_cap || n_records < 0 || n_records > ps->max_records || gt_source < 0 || gt_source > 1)
        return VGL_EINVAL;
    for (int32_t i = 0; i < n_records; ++i) // offsets ascend, every record holds its 32 fixed bytes, the last one ends inside the buffer
        if (rec_off[i + 1] < rec_off[i] + 32u || (int64_t)rec_off[i + 1] > n_bytes) {
            ps->err = "vgl_parse_bcf: record offsets must ascend, leave 32 bytes per record and stay inside n_bytes";
            return VGL_EINVAL;
        }
    memset(out, 0, sizeof *out);
    out->first_error_record = -1;
    out->sites = ps->h_sites;
    ps->n_records = 0;
    if (n_records == 0) return VGL_OK;
    PCK(cudaSetDevice(ps->device));
    cudaStream_t st = ps->stream;
    if (ps->placed) PCK(cudaStreamWaitEvent(st, ps->ev_placed, 0));
    PCK(cudaEventRecord(ps->ev[0], st));
    if (!(flags & VGL_PARSE_TEXT_ON_DEVICE)) {
        PCK(cudaMemcpyAsync(ps->d_text, ps->h_text, (size_t)n_bytes, cudaMemcpyHostToDevice, st));
        ps->d_text_bytes = (size_t)n_bytes;
    }
    PCK(cudaMemcpyAsync(ps->d_line_end, rec_off, ((size_t)n_records + 1) * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
    PCK(cudaEventRecord(ps->ev[1], st));
    static const uint32_t init[C_COUNT] = {0, 0, 0, 0, 0xFFFFFFFFu, 0, 0, 0};
    PCK(cudaMemcpyAsync(ps->d_counters, init, sizeof init, cudaMemcpyHostToDevice, st));
    if (ps->S <= 256)
        k_bcf_gt<8><<<ps->n_sms * 8, GT_WARPS * 32, 0, st>>>(ps->d_text, ps->d_line_end, (uint32_t)n_records, ps->S, gt_source, gt_key, ps->rm_invar, ps->d_sites,
                                                            ps->d_rows, ps->d_counters);
    else
        k_bcf_gt<32><<<ps->n_sms * 8, GT_WARPS * 32, 0, st>>>(ps->d_text, ps->d_line_end, (uint32_t)n_records, ps->S, gt_source, gt_key, ps->rm_invar, ps->d_sites,
                                                             ps->d_rows, ps->d_counters);
    ps->launches += 1;
    PCK(cudaGetLastError());
    PCK(cudaEventRecord(ps->ev[2], st));
    PCK(cudaMemcpyAsync(ps->h_counters, ps->d_counters, C_COUNT * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    PCK(cudaMemcpyAsync(ps->h_sites, ps->d_sites, (size_t)n_records * sizeof(vgl_in_site), cudaMemcpyDeviceToHost, st));
    PCK(cudaStreamSynchronize(st));
    PCK(cudaEventRecord(ps->ev_done, st));
    ps->n_records = n_records;
    out->n_records = n_records;
    out->n_errors = (int32_t)ps->h_counters[C_NERRORS];
    out->first_error_record = out->n_errors ? (int32_t)ps->h_counters[C_FIRSTERR] : -1;
    out->n_kept = (int32_t)ps->h_counters[C_NKEPT];
    out->bytes_consumed = rec_off[n_records];
    cudaEventElapsedTime(&out->ms_h2d, ps->ev[0], ps->ev[1]);
    cudaEventElapsedTime(&out->ms_kernels, ps->ev[1], ps->ev[2]);
    return VGL_OK;
}
