// Input path (include/vgl.h "Input path"): VCF text records -> packed true genotypes.
//
//   k_vcf_lines   one pass over the text: positions of the line feeds (record index), chained scan with decoupled
//                 look-back so the text is read once
//   k_vcf_gt      one warp per record: the nine fixed columns (POS, REF/ALT -> allele map, FORMAT -> GT index), then every
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

constexpr int LINES_THREADS = 256;
constexpr int LINES_BYTES_PER_THREAD = 32;
constexpr int LINES_TILE = LINES_THREADS * LINES_BYTES_PER_THREAD; // 8 KiB of text per tile

// counters (device words, mirrored to pinned host memory after a parse)
enum { C_NEWLINES = 0, C_TICKET = 1, C_NRECORDS = 2, C_NERRORS = 3, C_FIRSTERR = 4, C_NKEPT = 5, C_CONSUMED = 6, C_COUNT = 8 };

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
// tile_state word: bits 63..62 = 0 not ready, 1 = tile aggregate, 2 = inclusive prefix; low 32 bits = count
__global__ void __launch_bounds__(LINES_THREADS) k_vcf_lines(const uint8_t* __restrict__ text, uint32_t n_bytes, uint32_t n_tiles,
                                                               uint32_t max_records, uint32_t* __restrict__ line_end,
                                                               unsigned long long* tile_state, uint32_t* counters)
{
    __shared__ uint32_t s_tile, s_base, s_warp[LINES_THREADS / 32];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    for (;;) {
        __syncthreads();
        if (tid == 0) s_tile = atomicAdd(&counters[C_TICKET], 1u);
        __syncthreads();
        const uint32_t tile = s_tile;
        if (tile >= n_tiles) return;
        const uint32_t off = tile * (uint32_t)LINES_TILE + (uint32_t)tid * LINES_BYTES_PER_THREAD;
        uint32_t m = 0;
        if (off < n_bytes) { // the buffer is padded to a whole tile
            const uint4 a = __ldg(reinterpret_cast<const uint4*>(text + off));
            const uint4 b = __ldg(reinterpret_cast<const uint4*>(text + off + 16));
            m = eq_mask16(a, 0x0A0A0A0Au) | (eq_mask16(b, 0x0A0A0A0Au) << 16);
            const uint32_t left = n_bytes - off;
            if (left < 32) m &= (1u << left) - 1u;
        }
        const uint32_t cnt = __popc(m);
        uint32_t incl = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += t;
        }
        if (lane == 31) s_warp[wid] = incl;
        __syncthreads();
        uint32_t wbase = 0, total = 0;
#pragma unroll
        for (int w = 0; w < LINES_THREADS / 32; ++w) {
            const uint32_t c = s_warp[w];
            if (w < wid) wbase += c;
            total += c;
        }
        if (tid == 0) {
            uint32_t base = 0;
            if (tile > 0) {
                atomicExch(&tile_state[tile], (1ull << 62) | total);
                for (int64_t t = (int64_t)tile - 1; t >= 0; --t) {
                    unsigned long long w;
                    do {
                        w = atomicAdd(&tile_state[t], 0ull);
                    } while ((w >> 62) == 0);
                    base += (uint32_t)w;
                    if ((w >> 62) == 2) break;
                }
            }
            __threadfence();
            atomicExch(&tile_state[tile], (2ull << 62) | (unsigned long long)(base + total));
            s_base = base;
            if (tile == n_tiles - 1) counters[C_NEWLINES] = base + total;
        }
        __syncthreads();
        uint32_t idx = s_base + wbase + incl - cnt;
        while (m) {
            const int b = __ffs(m) - 1;
            m &= m - 1;
            if (idx < max_records) line_end[idx] = off + b;
            ++idx;
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
__device__ __forceinline__ Hdr parse_header(const uint8_t* __restrict__ text, const uint32_t* tab, int gt_source)
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

constexpr int GT_WARPS = 8;

__global__ void __launch_bounds__(GT_WARPS * 32) k_vcf_gt(const uint8_t* __restrict__ text, uint32_t n_bytes, const uint32_t* __restrict__ line_end,
                                                         int32_t S, int32_t gt_source, int32_t rm_invar, uint32_t max_records,
                                                         vgl_in_site* __restrict__ sites, uint8_t* __restrict__ rows, uint32_t* counters)
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
    for (uint32_t line = warp0; line < n_rec; line += n_warps) {
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
                h = parse_header(text, tab, gt_source);
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
            o.allele_sum = asum;
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
                atomicAdd(&counters[C_NERRORS], 1u);
                atomicMin(&counters[C_FIRSTERR], line);
            } else if (o.skip_code == 0) atomicAdd(&counters[C_NKEPT], 1u);
        }
        __syncwarp();
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
    vgl_parser* ps = new (std::nothrow) vgl_parser();
    if (!ps) return VGL_ENOMEM;
    ps->device = device, ps->S = S, ps->rm_invar = rm_invar & 3, ps->n_sms = n_sms;
    ps->max_records = max_records;
    ps->text_cap = (size_t)max_text;
    const size_t padded = ((ps->text_cap + 1 + LINES_TILE - 1) / LINES_TILE) * LINES_TILE + 1024;
    ps->max_tiles = (uint32_t)(padded / LINES_TILE);
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
    PCK(cudaMalloc((void**)&ps->d_tile_state, (size_t)ps->max_tiles * sizeof(unsigned long long)));
    PCK(cudaMalloc((void**)&ps->d_counters, C_COUNT * sizeof(uint32_t)));
    PCK(cudaHostAlloc((void**)&ps->h_counters, C_COUNT * sizeof(uint32_t), cudaHostAllocDefault));
    PCK(cudaMalloc((void**)&ps->d_sites, (size_t)max_records * sizeof(vgl_in_site)));
    PCK(cudaHostAlloc((void**)&ps->h_sites, (size_t)max_records * sizeof(vgl_in_site), cudaHostAllocDefault));
    PCK(cudaMalloc((void**)&ps->d_rows, (size_t)max_records * S + 16));
    PCK(cudaMalloc((void**)&ps->d_row_map, (size_t)max_records * sizeof(int32_t)));
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
    cudaFree(ps->d_tile_state);
    cudaFree(ps->d_counters);
    cudaFreeHost(ps->h_counters);
    cudaFree(ps->d_sites);
    cudaFreeHost(ps->h_sites);
    cudaFree(ps->d_rows);
    cudaFree(ps->d_row_map);
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
    const uint32_t n_tiles = (uint32_t)((n + LINES_TILE - 1) / LINES_TILE);
    PCK(cudaMemsetAsync(ps->d_tile_state, 0, (size_t)n_tiles * sizeof(unsigned long long), st));
    static const uint32_t init[C_COUNT] = {0, 0, 0, 0, 0xFFFFFFFFu, 0, 0, 0};
    PCK(cudaMemcpyAsync(ps->d_counters, init, sizeof init, cudaMemcpyHostToDevice, st));
    const uint32_t lines_grid = (uint32_t)std::min<uint64_t>(n_tiles, (uint64_t)ps->n_sms * 8);
    k_vcf_lines<<<lines_grid, LINES_THREADS, 0, st>>>(ps->d_text, (uint32_t)n, n_tiles, (uint32_t)ps->max_records, ps->d_line_end, ps->d_tile_state,
                                                     ps->d_counters);
    k_vcf_gt<<<ps->n_sms * 8, GT_WARPS * 32, 0, st>>>(ps->d_text, (uint32_t)n, ps->d_line_end, ps->S, gt_source, ps->rm_invar, (uint32_t)ps->max_records,
                                                     ps->d_sites, ps->d_rows, ps->d_counters);
    ps->launches += 2;
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
