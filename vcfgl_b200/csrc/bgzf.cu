// VGL_HOST_BGZF -- the BCF record stream of a batch (bcf.cu) compressed on the device into BGZF blocks, the container the
// reference writes by default (-O b: htslib/bgzf.c, thread pool set up at vcfgl.cpp:1791-1803).  What crosses PCIe is the
// compressed stream; the host appends it to the file after its own header block(s) and ends the file with the BGZF EOF block.
//
// A BGZF block is a gzip member with a "BC" extra field holding the block size (htslib/bgzf.c:  18-byte header, raw deflate
// data, CRC32 and length of the uncompressed bytes).  Every block here holds BGZF_IN = 32768 bytes of the record stream (the
// last one less) and is compressed by one CTA on its own, so every match distance lies inside deflate's 32 KiB window:
//
//   k_bgzf_deflate  one block per CTA.  The record stream is not searched byte by byte: k_bcf_emit leaves a descriptor of every
//                   record's FORMAT planes, so the CTA knows where each sample's vector of a tag ("cell": the 15 GL floats, the
//                   15 PL bytes, the 5 AD counts ...) starts.  Cells are hashed whole; a cell whose bytes occurred earlier in the
//                   block as a cell becomes ONE length/distance pair pointing at the first such occurrence, everything else
//                   goes out as literals.  Cells never overlap, so the parse needs no sequential pass; the simulated tags
//                   repeat heavily (every sample with the same read counts has the same vectors), which is what makes the
//                   record stream compressible at all.  Codes are deflate's fixed Huffman codes (RFC 1951 3.2.6): bit lengths
//                   per segment -> CTA-wide prefix sum -> every segment ORs its bits into the block image in shared memory.
//                   CRC32: 32-byte chunks per thread, combined by multiplication with x^(8 n) mod P (the identity zlib's
//                   crc32_combine uses).
//   k_bgzf_first    thread per block: the first record that reaches into it (binary search over the record offsets)
//   k_bgzf_scan     exclusive prefix of the compressed block sizes
//   k_bgzf_pack     blocks moved back to back into the stream the host receives -- written straight into the slot's pinned host
//                   buffer (mapped memory), so the transfer is part of the stream's work and needs no size known to the host
//
// Parity: inflating the blocks gives back the VGL_HOST_BCF stream byte for byte (tests/test_gpu_bgzf.py: zlib on the host).
#include "vgl_internal.h"

#include <algorithm>
#include <cstdio>
#include <cstdlib>

namespace vgl {

namespace {

constexpr int BGZF_IN = 32768;                  // uncompressed bytes per block
constexpr int BGZF_THREADS = 1024;
constexpr int MAX_SEG = 8192;                   // segments (cells, literal runs) of a block
constexpr int RUN = 64;                         // bytes of a literal run segment
constexpr int OUT_WORDS = (BGZF_IN * 9 / 8 + 64) / 4; // fixed codes: at most 9 bits per input byte, + header bits / end of block / padding
constexpr int HASH_SLOTS = 8192;
constexpr int MAX_RANGES = 32 * 16;

struct Seg { // 8 bytes
    uint16_t pos, len;   // position in the block, bytes
    uint16_t dist;       // 0: literals; else a match of `len` bytes at this distance
    uint16_t cell;       // 1: a cell (match candidate)
};

__device__ __forceinline__ uint32_t rev_bits(uint32_t code, int n) { return __brev(code) >> (32 - n); }

// fixed Huffman code of a literal byte, already bit-reversed (codes go into the stream most significant bit first)
__device__ __forceinline__ void lit_code(uint32_t v, uint32_t& bits, int& n)
{
    if (v < 144u) { bits = rev_bits(0x30u + v, 8); n = 8; }
    else { bits = rev_bits(0x190u + (v - 144u), 9); n = 9; }
}
// length / distance pair (3 <= len <= 258, 1 <= dist <= 32768): up to 31 bits
__device__ __forceinline__ void match_code(int len, int dist, uint32_t& bits, int& n)
{
    int idx, eb;
    uint32_t ev;
    const int t = len - 3;
    if (t < 8) { idx = t; eb = 0; ev = 0u; }
    else if (len == 258) { idx = 28; eb = 0; ev = 0u; }
    else {
        const int hb = 31 - __clz(t);
        eb = hb - 2;
        idx = 4 * (hb - 1) + ((t >> eb) & 3);
        ev = (uint32_t)t & ((1u << eb) - 1u);
    }
    const int sym = 257 + idx;
    uint32_t b;
    int nb;
    if (sym < 280) { b = rev_bits((uint32_t)(sym - 256), 7); nb = 7; }
    else { b = rev_bits(0xC0u + (uint32_t)(sym - 280), 8); nb = 8; }
    b |= ev << nb;
    nb += eb;
    int dc, deb;
    uint32_t dev;
    if (dist <= 4) { dc = dist - 1; deb = 0; dev = 0u; }
    else {
        const int u = dist - 1, hb = 31 - __clz(u);
        deb = hb - 1;
        dc = 2 * hb + ((u >> deb) & 1);
        dev = (uint32_t)u & ((1u << deb) - 1u);
    }
    b |= rev_bits((uint32_t)dc, 5) << nb;
    nb += 5;
    b |= dev << nb;
    nb += deb;
    bits = b;
    n = nb;
}

// ORs `n` (<= 32) bits into the block image at bit offset `at`
__device__ __forceinline__ void put_bits(uint32_t* out, unsigned at, uint32_t bits, int n)
{
    const unsigned w = at >> 5, sh = at & 31u;
    atomicOr(&out[w], bits << sh);
    if (sh + (unsigned)n > 32u) atomicOr(&out[w + 1], bits >> (32u - sh));
}

// a * b mod P over GF(2) in the reflected representation of CRC-32 (P = 0xEDB88320): shifts a CRC across the bytes that follow
__host__ __device__ inline uint32_t crc_mul(uint32_t a, uint32_t b)
{
    uint32_t m = 1u << 31, p = 0u;
    for (;;) {
        if (a & m) {
            p ^= b;
            if ((a & (m - 1u)) == 0u) break;
        }
        m >>= 1;
        b = (b & 1u) ? (b >> 1) ^ 0xEDB88320u : b >> 1;
    }
    return p;
}

// first record whose bytes reach into block b (the last record that starts at or before the block's first byte)
__global__ void __launch_bounds__(256) k_bgzf_first(const BgzfArgs a)
{
    const long long total = a.totals[3];
    const long long nblk = total > a.in_cap ? 0 : (total + BGZF_IN - 1) / BGZF_IN;
    for (long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x; b < nblk; b += (long long)gridDim.x * blockDim.x) {
        const long long b0 = b * BGZF_IN;
        int lo = 0, hi = a.n_sites; // rec_off[lo] <= b0 < rec_off[hi]; rec_off is non-decreasing
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (a.rec_off[mid] <= b0) lo = mid; else hi = mid;
        }
        a.blk_first[b] = lo;
    }
}

// 32-bit word at byte offset `pos` of shared memory (any alignment)
__device__ __forceinline__ uint32_t word_at(const uint8_t* base, uint32_t pos)
{
    const uint32_t* w = reinterpret_cast<const uint32_t*>(base) + (pos >> 2);
    const uint32_t sh = (pos & 3u) * 8u;
    return sh ? __funnelshift_r(w[0], w[1], sh) : w[0];
}

__global__ void __launch_bounds__(BGZF_THREADS, 1) k_bgzf_deflate(const BgzfArgs a)
{
    extern __shared__ __align__(16) unsigned char sm[];
    uint8_t* const in = sm;                                                       // [BGZF_IN]
    uint32_t* const out = reinterpret_cast<uint32_t*>(sm + BGZF_IN);              // [OUT_WORDS]
    Seg* const segs = reinterpret_cast<Seg*>(sm + BGZF_IN + OUT_WORDS * 4);        // [MAX_SEG]
    uint32_t* const seg_bit = reinterpret_cast<uint32_t*>(segs + MAX_SEG);        // [MAX_SEG] bit offset of every segment
    uint32_t* const htab = seg_bit + MAX_SEG;                                     // [HASH_SLOTS] tag << 16 | position
    uint32_t* const rng = htab + HASH_SLOTS;                                      // [MAX_RANGES][4]: pos, len, cell bytes (0: gap), first segment
    __shared__ uint32_t crc_tab[1024]; // slicing-by-four tables of CRC-32
    __shared__ uint16_t lit_tab[256];
    __shared__ uint32_t warp_tot[32];
    __shared__ int n_rng_s, n_seg_s, lit_only_s;
    __shared__ uint32_t crc_s, total_bits_s;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long total = a.totals[3];
    const long long b0 = (long long)blockIdx.x * BGZF_IN;
    if (b0 >= total || total > a.in_cap) return;
    const int L = (int)min((long long)BGZF_IN, total - b0);

    // ---- warps 1..: the block's bytes, the CRC and literal-code tables, cleared tables; meanwhile warp 0: the ranges
    if (warp > 0) {
        const int t = tid - 32, nt = BGZF_THREADS - 32;
        for (int i = t; i < (L + 15) / 16; i += nt) reinterpret_cast<uint4*>(in)[i] = __ldcs(reinterpret_cast<const uint4*>(a.in + b0) + i);
        for (int i = t; i < OUT_WORDS / 4; i += nt) reinterpret_cast<uint4*>(out)[i] = make_uint4(0u, 0u, 0u, 0u);
        for (int i = t; i < HASH_SLOTS / 4; i += nt) reinterpret_cast<uint4*>(htab)[i] = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);
        if (t < 256) {
            uint32_t c = (uint32_t)t;
#pragma unroll
            for (int k = 0; k < 8; ++k) c = (c & 1u) ? (c >> 1) ^ 0xEDB88320u : c >> 1;
            crc_tab[t] = c;
            uint32_t c1 = c;
#pragma unroll
            for (int lvl = 1; lvl < 4; ++lvl) { // T_lvl[i] = the CRC register after byte i and lvl zero bytes
                uint32_t z = c1 & 0xFFu;
#pragma unroll
                for (int k = 0; k < 8; ++k) z = (z & 1u) ? (z >> 1) ^ 0xEDB88320u : z >> 1;
                c1 = z ^ (c1 >> 8);
                crc_tab[256 * lvl + t] = c1;
            }
            uint32_t b;
            int n;
            lit_code((uint32_t)t, b, n);
            lit_tab[t] = (uint16_t)(b | ((uint32_t)(n - 8) << 15)); // 9 code bits, bit 15: one more than 8
        }
        if (t == 0) crc_s = 0u;
        // CRC32 of the block while warp 0 is busy with the ranges: 32-byte chunks from the end (slicing by four), each shifted
        // across the bytes behind it.  Barrier 1 = warps 1 .. 31 only.
        asm volatile("bar.sync 1, %0;" ::"r"(nt) : "memory");
        for (int c = t; c < 1024; c += nt) {
            const int hi = L - c * 32, lo2 = max(hi - 32, 0);
            if (hi <= 0) break;
            uint32_t r = 0xFFFFFFFFu;
            int k = lo2;
            for (; k < hi && ((hi - k) & 3); ++k) r = crc_tab[(r ^ in[k]) & 0xFFu] ^ (r >> 8); // head bytes: the rest is whole words
            for (; k < hi; k += 4) {
                r ^= word_at(in, (uint32_t)k);
                r = crc_tab[768 + (r & 0xFFu)] ^ crc_tab[512 + ((r >> 8) & 0xFFu)] ^ crc_tab[256 + ((r >> 16) & 0xFFu)] ^ crc_tab[r >> 24];
            }
            const uint32_t part = crc_mul(a.crc_pow[c], ~r);
            if (part) atomicXor(&crc_s, part);
        }
    }

    // ---- ranges: the part of every record that lies in the block, split at the FORMAT planes (warp 0, a lane per record)
    if (warp == 0) {
        if (lane == 0) { n_rng_s = 0; n_seg_s = 0; lit_only_s = 0; }
        __syncwarp();
        const int lo = a.blk_first[blockIdx.x];
        const int r = lo + lane;
        int my_n = 0;
        uint32_t my[16][3]; // pos, len, cell
        long long rs = 0, re = 0;
        if (r < a.n_sites) { rs = a.rec_off[r]; re = a.rec_off[r + 1]; }
        const bool has = r < a.n_sites && re > rs && rs < b0 + L && re > b0;
        if (has) {
            const BcfRecPlanes pl = a.planes[r];
            long long cur = max(rs, b0); // next byte of the record not yet put into a range
            const long long end = min(re, b0 + L);
            for (int k = 0; k < (int)pl.n && k < 7; ++k) {
                const long long ps = rs + pl.off[k], pe = ps + (long long)pl.cell[k] * a.S;
                if (pl.cell[k] < 3 || pl.cell[k] > RUN || pe <= cur || ps >= end) continue; // deflate matches are at least 3 bytes long; segments at most RUN
                // full cells of this plane inside [cur, end)
                long long c_lo = ps >= cur ? 0 : (cur - ps + pl.cell[k] - 1) / pl.cell[k];
                long long c_hi = pe <= end ? a.S : (end - ps) / pl.cell[k];
                if (c_hi <= c_lo) continue;
                const long long cs = ps + c_lo * pl.cell[k], ce = ps + c_hi * pl.cell[k];
                if (cs > cur) { my[my_n][0] = (uint32_t)(cur - b0); my[my_n][1] = (uint32_t)(cs - cur); my[my_n][2] = 0u; ++my_n; }
                my[my_n][0] = (uint32_t)(cs - b0); my[my_n][1] = (uint32_t)(ce - cs); my[my_n][2] = pl.cell[k]; ++my_n;
                cur = ce;
            }
            if (cur < end) { my[my_n][0] = (uint32_t)(cur - b0); my[my_n][1] = (uint32_t)(end - cur); my[my_n][2] = 0u; ++my_n; }
        }
        // more than 32 records in one block (a handful of samples): the block goes out as literals, nothing to gain there
        const long long last_end = __shfl_sync(0xffffffffu, has ? re : 0, 31);
        const bool overflow = lo + 32 < a.n_sites && last_end < b0 + L && __shfl_sync(0xffffffffu, (int)has, 31);
        int segs_mine = 0, bytes_mine = 0;
        for (int k = 0; k < my_n; ++k) {
            segs_mine += my[k][2] ? (int)(my[k][1] / my[k][2]) : (int)((my[k][1] + RUN - 1) / RUN);
            bytes_mine += (int)my[k][1];
        }
        const bool covered = __reduce_add_sync(0xffffffffu, bytes_mine) == L; // else: records beyond the 32 lanes (skipped sites in between)
        int inc_r = my_n, inc_s = segs_mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t1 = __shfl_up_sync(0xffffffffu, inc_r, o), t2 = __shfl_up_sync(0xffffffffu, inc_s, o);
            if (lane >= o) { inc_r += t1; inc_s += t2; }
        }
        const int tot_r = __shfl_sync(0xffffffffu, inc_r, 31), tot_s = __shfl_sync(0xffffffffu, inc_s, 31);
        if (overflow || !covered || tot_s > MAX_SEG || tot_r > MAX_RANGES) {
            if (lane == 0) lit_only_s = 1;
        } else {
            int r0 = inc_r - my_n, s0 = inc_s - segs_mine;
            for (int k = 0; k < my_n; ++k) {
                rng[(r0 + k) * 4 + 0] = my[k][0]; rng[(r0 + k) * 4 + 1] = my[k][1]; rng[(r0 + k) * 4 + 2] = my[k][2]; rng[(r0 + k) * 4 + 3] = (uint32_t)s0;
                s0 += my[k][2] ? (int)(my[k][1] / my[k][2]) : (int)((my[k][1] + RUN - 1) / RUN);
            }
            if (lane == 0) { n_rng_s = tot_r; n_seg_s = tot_s; }
        }
    }
    __syncthreads();
    if (lit_only_s) { // one gap over the whole block
        if (tid == 0) {
            rng[0] = 0u; rng[1] = (uint32_t)L; rng[2] = 0u; rng[3] = 0u;
            n_rng_s = 1;
            n_seg_s = (L + RUN - 1) / RUN;
        }
        __syncthreads();
    }
    const int n_rng = n_rng_s, n_seg = n_seg_s;

    // ---- segment table
    for (int j = warp; j < n_rng; j += BGZF_THREADS / 32) {
        const uint32_t pos = rng[j * 4], len = rng[j * 4 + 1], cell = rng[j * 4 + 2], first = rng[j * 4 + 3];
        const int n = cell ? (int)(len / cell) : (int)((len + RUN - 1) / RUN);
        for (int c = lane; c < n; c += 32) {
            Seg s;
            if (cell) { s.pos = (uint16_t)(pos + (uint32_t)c * cell); s.len = (uint16_t)cell; s.cell = 1; }
            else { s.pos = (uint16_t)(pos + (uint32_t)c * RUN); s.len = (uint16_t)min((uint32_t)RUN, len - (uint32_t)c * RUN); s.cell = 0; }
            s.dist = 0;
            segs[first + c] = s;
        }
    }
    __syncthreads();

    // ---- cells: hash of the bytes, first occurrence per hash (atomicMin on tag << 16 | position; open addressing)
    auto cell_hash = [&](const Seg& s) -> uint32_t { // FNV-style over 32-bit words (the tail word masked)
        uint32_t h = 2166136261u ^ s.len;
        const int nw = s.len >> 2;
        for (int k = 0; k < nw; ++k) h = (h ^ word_at(in, s.pos + 4u * k)) * 16777619u;
        if (s.len & 3) h = (h ^ (word_at(in, s.pos + 4u * nw) & ((1u << (8 * (s.len & 3))) - 1u))) * 16777619u;
        h ^= h >> 15;
        return h;
    };
    for (int j = tid; j < n_seg; j += BGZF_THREADS) {
        const Seg s = segs[j];
        if (!s.cell) continue;
        const uint32_t h = cell_hash(s), tag = h >> 16, mine = (tag << 16) | s.pos;
        seg_bit[j] = h; // kept for the look-up pass
        uint32_t slot = h & (HASH_SLOTS - 1);
        for (int probe = 0; probe < 16; ++probe) {
            const uint32_t old = atomicCAS(&htab[slot], 0xFFFFFFFFu, mine);
            if (old == 0xFFFFFFFFu) break;
            if ((old >> 16) == tag) { atomicMin(&htab[slot], mine); break; }
            slot = (slot + 1) & (HASH_SLOTS - 1);
        }
    }
    __syncthreads();
    // ---- match = the first cell of the block with the same bytes; bit length of every segment
    for (int j = tid; j < n_seg; j += BGZF_THREADS) {
        Seg s = segs[j];
        uint32_t nbits = 0;
        if (s.cell) {
            const uint32_t h = seg_bit[j], tag = h >> 16;
            uint32_t slot = h & (HASH_SLOTS - 1);
            for (int probe = 0; probe < 16; ++probe) {
                const uint32_t e = htab[slot];
                if (e == 0xFFFFFFFFu) break;
                if ((e >> 16) == tag) {
                    const uint32_t q = e & 0xFFFFu;
                    if (q < s.pos && q + s.len <= s.pos) { // an earlier cell (cells do not overlap)
                        bool same = true;
                        const int nw = s.len >> 2;
                        for (int k = 0; k < nw && same; ++k) same = word_at(in, q + 4u * k) == word_at(in, s.pos + 4u * k);
                        for (int k = 4 * nw; k < s.len && same; ++k) same = in[q + k] == in[s.pos + k];
                        if (same) s.dist = (uint16_t)(s.pos - q);
                    }
                    break;
                }
                slot = (slot + 1) & (HASH_SLOTS - 1);
            }
        }
        if (s.dist) {
            uint32_t b;
            int n;
            match_code(s.len, s.dist, b, n);
            nbits = (uint32_t)n;
            segs[j].dist = s.dist;
            segs[j].cell = 0;
        } else { // 8 bits per literal, 9 for the values from 144 up; the counts before quarters 1..3 go into the free cell field
            const int ql = (s.len + 3) >> 2; // <= 16: segments are at most 64 bytes long
            uint32_t extra = 0u, cum = 0u;
#pragma unroll
            for (int q = 0; q < 4; ++q) { // 9-bit literals (values from 144 up) of quarter q, four bytes at a time
                const int k0 = q * ql, k1 = min(k0 + ql, (int)s.len);
                uint32_t part = 0u;
                for (int k = k0; k < k1; k += 4) {
                    uint32_t w = __vcmpgeu4(word_at(in, s.pos + (uint32_t)k), 0x90909090u) & 0x01010101u;
                    if (k1 - k < 4) w &= (1u << (8 * (k1 - k))) - 1u;
                    part += __popc(w);
                }
                if (q < 3) cum |= part << (5 * q);
                extra += part;
            }
            nbits = 8u * s.len + extra;
            segs[j].cell = (uint16_t)(0x8000u | cum); // bit 15: literal segment
        }
        seg_bit[j] = nbits;
    }
    __syncthreads();
    // ---- exclusive prefix of the bit lengths (8 segments per thread), 3 header bits in front
    {
        uint32_t v[MAX_SEG / BGZF_THREADS], sum = 0u;
#pragma unroll
        for (int k = 0; k < MAX_SEG / BGZF_THREADS; ++k) {
            const int j = tid * (MAX_SEG / BGZF_THREADS) + k;
            v[k] = j < n_seg ? seg_bit[j] : 0u;
            sum += v[k];
        }
        uint32_t inc = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) warp_tot[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            uint32_t t = warp_tot[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t u = __shfl_up_sync(0xffffffffu, t, o);
                if (lane >= o) t += u;
            }
            warp_tot[lane] = t;
        }
        __syncthreads();
        uint32_t base = 3u + (warp ? warp_tot[warp - 1] : 0u) + inc - sum;
#pragma unroll
        for (int k = 0; k < MAX_SEG / BGZF_THREADS; ++k) {
            const int j = tid * (MAX_SEG / BGZF_THREADS) + k;
            if (j < n_seg) seg_bit[j] = base;
            base += v[k];
        }
        if (tid == BGZF_THREADS - 1) total_bits_s = base;
    }
    __syncthreads();
    // ---- bits: BFINAL = 1, BTYPE = 01 (fixed Huffman), the segments, end of block (7 zero bits: already there).
    // A literal segment is split into quarters (their bit offsets follow from the counts kept with the segment).
    if (tid == 0) atomicOr(&out[0], 3u);
    {
        // a warp takes 32 consecutive segments: the matches go out one per lane, then the literal segments of the group are
        // spread over the lanes a quarter each, so that the byte loops run on full warps
        for (int j0 = warp * 32; j0 < n_seg; j0 += (BGZF_THREADS / 32) * 32) {
            const int j = j0 + lane;
            Seg s;
            s.pos = s.len = s.dist = s.cell = 0;
            if (j < n_seg) s = segs[j];
            if (s.dist) {
                uint32_t bb;
                int n;
                match_code(s.len, s.dist, bb, n);
                put_bits(out, seg_bit[j], bb, n);
            }
            const uint32_t litm = __ballot_sync(0xffffffffu, j < n_seg && (s.cell & 0x8000u));
            const int ntask = 4 * __popc(litm);
            for (int t = lane; t < ntask; t += 32) {
                const int jj = j0 + (int)__fns(litm, 0u, (t >> 2) + 1), q = t & 3;
                const Seg ls = segs[jj];
                const int ql = (ls.len + 3) >> 2, k0 = q * ql, k1 = min(k0 + ql, (int)ls.len);
                if (k0 >= k1) continue;
                const unsigned c = ls.cell; // 9-bit literals in quarters 0, 1, 2: five bits each
                const unsigned cum = (q > 0 ? c & 31u : 0u) + (q > 1 ? (c >> 5) & 31u : 0u) + (q > 2 ? (c >> 10) & 31u : 0u);
                unsigned at = seg_bit[jj] + 8u * (unsigned)k0 + cum;
                unsigned w = at >> 5;
                int fill = (int)(at & 31u);
                unsigned long long acc = 0ull;
                for (int k = k0; k < k1; ++k) {
                    const uint32_t e = lit_tab[in[ls.pos + k]];
                    acc |= (unsigned long long)(e & 0x1FFu) << fill;
                    fill += 8 + (int)(e >> 15);
                    if (fill >= 32) {
                        atomicOr(&out[w], (uint32_t)acc);
                        ++w;
                        acc >>= 32;
                        fill -= 32;
                    }
                }
                if (fill > 0) atomicOr(&out[w], (uint32_t)acc);
            }
        }
    }
    __syncthreads();
    // ---- the BGZF block: header, deflate data, CRC32, ISIZE -> its slot of the staging buffer
    const uint32_t nbytes = (total_bits_s + 7u + 7u) >> 3; // + the 7-bit end-of-block code, rounded up to a byte
    const uint32_t bsize = 18u + nbytes + 8u;
    uint8_t* const dst = a.stage + (size_t)blockIdx.x * BGZF_STRIDE + 2; // + 2: the deflate data (at + 18) starts on a word
    if (tid < 18) {
        const uint8_t hdr[18] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 'B', 'C', 2, 0, (uint8_t)((bsize - 1u) & 0xFFu), (uint8_t)((bsize - 1u) >> 8)};
        dst[tid] = hdr[tid];
    }
    {
        uint32_t* const dw = reinterpret_cast<uint32_t*>(dst + 18);
        const uint8_t* const ob = reinterpret_cast<const uint8_t*>(out);
        for (uint32_t i = tid; i < (nbytes >> 2); i += BGZF_THREADS) __stcs(dw + i, out[i]);
        if (tid < (nbytes & 3u)) dst[18 + (nbytes & ~3u) + tid] = ob[(nbytes & ~3u) + tid];
    }
    if (tid < 8) {
        const uint32_t v = tid < 4 ? crc_s : (uint32_t)L;
        dst[18 + nbytes + tid] = (uint8_t)(v >> (8 * (tid & 3)));
    }
    if (tid == 0) a.blk_size[blockIdx.x] = bsize;
}

// exclusive prefix of the block sizes; the totals for the host: [4] compressed bytes, [5] blocks
__global__ void __launch_bounds__(1024) k_bgzf_scan(const BgzfArgs a)
{
    __shared__ long long warp_sum[32];
    __shared__ long long carry_s;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const long long total = a.totals[3];
    const int nblk = total > a.in_cap ? 0 : (int)((total + BGZF_IN - 1) / BGZF_IN);
    if (tid == 0) carry_s = 0;
    __syncthreads();
    for (int base = 0; base < nblk; base += 1024) {
        const int i = base + tid;
        const long long v = i < nblk ? (long long)a.blk_size[i] : 0;
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
        const long long incl = carry_s + (w ? warp_sum[w - 1] : 0) + x;
        if (i < nblk) a.blk_off[i] = incl - v;
        __syncthreads();
        if (tid == 1023) carry_s = incl;
        __syncthreads();
    }
    if (tid == 0) {
        a.totals[4] = carry_s;
        a.totals[5] = nblk;
    }
}

__global__ void __launch_bounds__(256) k_bgzf_pack(const BgzfArgs a)
{
    const long long total = a.totals[3];
    const int nblk = total > a.in_cap ? 0 : (int)((total + BGZF_IN - 1) / BGZF_IN);
    for (int b = blockIdx.x; b < nblk; b += gridDim.x) {
        const uint8_t* src = a.stage + (size_t)b * BGZF_STRIDE + 2;
        uint8_t* dst = a.out + a.blk_off[b];
        const uint32_t n = a.blk_size[b];
        // aligned words of the destination, edges by bytes
        const uint32_t head = (uint32_t)((4u - ((uintptr_t)dst & 3u)) & 3u);
        for (uint32_t i = threadIdx.x; i < min(head, n); i += 256) dst[i] = src[i];
        if (n > head) {
            const uint32_t words = (n - head) >> 2;
            for (uint32_t w = threadIdx.x; w < words; w += 256) {
                const uint8_t* s = src + head + 4u * w;
                *reinterpret_cast<uint32_t*>(dst + head + 4u * w) = (uint32_t)s[0] | ((uint32_t)s[1] << 8) | ((uint32_t)s[2] << 16) | ((uint32_t)s[3] << 24);
            }
            for (uint32_t i = head + 4u * words + threadIdx.x; i < n; i += 256) dst[i] = src[i];
        }
    }
}

} // namespace

size_t bgzf_dyn_smem() { return (size_t)BGZF_IN + (size_t)OUT_WORDS * 4 + (size_t)MAX_SEG * (sizeof(Seg) + 4) + (size_t)HASH_SLOTS * 4 + (size_t)MAX_RANGES * 16; }

// blocks a record stream of `bytes` bytes makes
int64_t bgzf_blocks_for(int64_t bytes) { return (bytes + BGZF_IN - 1) / BGZF_IN; }

// x^(8 * 32 * k) mod P for k = 0 .. 1023 (reflected CRC-32 representation)
void bgzf_crc_pow_table(uint32_t* t)
{
    uint32_t x8 = 1u << 31; // x^0
    for (int k = 0; k < 8; ++k) x8 = crc_mul(x8, 1u << 30); // times x
    uint32_t x256 = 1u << 31;
    for (int k = 0; k < 32; ++k) x256 = crc_mul(x256, x8);
    t[0] = 1u << 31;
    for (int k = 1; k < 1024; ++k) t[k] = crc_mul(t[k - 1], x256);
}

void launch_bgzf(const BgzfArgs& a, int64_t max_blocks, cudaStream_t st, int n_sms)
{
    cudaFuncSetAttribute(k_bgzf_deflate, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bgzf_dyn_smem()); // per device
    k_bgzf_first<<<(unsigned)std::min<int64_t>((max_blocks + 255) / 256, 4096), 256, 0, st>>>(a);
    k_bgzf_deflate<<<(unsigned)max_blocks, BGZF_THREADS, bgzf_dyn_smem(), st>>>(a);
    k_bgzf_scan<<<1, 1024, 0, st>>>(a);
    k_bgzf_pack<<<(unsigned)(n_sms * 8), 256, 0, st>>>(a);
}

} // namespace vgl
