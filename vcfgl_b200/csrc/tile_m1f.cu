// k_tile_m1f -- the headline path in one kernel, built for a low instruction count per cell:
// native RNG, GL model 1 with one run-constant quality score and error rate (--error-qs 0),
// same-mean Poisson or fixed depth (<= 255 reads per cell), no strand / GP / QS / I16 tags.
// Everything else runs through k_fused_m1f (fused.cu) or the unfused kernels (kernels.cu).
//
// Persistent CTAs pull TILES of whole sites from an atomic ticket.  A tile is a run of "virtual
// cells": every site occupies S4 = ceil4(S) slots, so that any run of 32 slots starts at a sample
// index that is a multiple of 4 -> its span in every [S][G] / [S][A] output block starts and ends
// on a 16-byte boundary (blocks themselves are padded to 16 B).
//   phase A  thread per cell: two Philox blocks -> depth (alias table), number of mis-called reads
//            (threshold table), haplotype split (popcount), error placement -> four 8-bit base
//            counts kept in shared memory; FORMAT/DP written; per-site totals by packed warp REDUX
//   phase B  thread per site: allele order, unobserved allele, skip code, INFO tags
//            (vcfgl.cpp:396-404, 665-782); decoupled look-back gives the tile's base offsets
//   phase C  warp per 32 cells: errmod scores from the counts (m1f.cuh), GL / PL / AD scattered into
//            the warp's own shared-memory slice in allele order, then ONE bulk async copy
//            (cp.async.bulk shared -> global) per plane writes the whole 16-byte aligned span.
// HBM traffic = 1 B/cell in, the tag planes out.
#include "counts_sampler.cuh"
#include "m1f.cuh"

namespace vgl {

#define TILE_BLOCK 256
#define TILE_WARPS (TILE_BLOCK / 32)
#define TILE_MAX_SITES 64
#define TILE_CELLS 4096      // virtual cells of a tile when a site is smaller than this
#define TILE_WST_G 512       // 4-byte elements per warp and G-shaped plane: 32 cells x 15 (+ pads) <= 32 x 16
#define TILE_WST_R 192       // 32 cells x 5 (+ pads) <= 32 x 6

struct __align__(16) TSite {
    uint32_t slot[4];     // byte k = 4 * (allele-space genotype slot of base pair k), 0xFF = pair not at this site
    int32_t g_rel, r_rel; // element offsets of the site's blocks relative to the tile's base
    uint32_t a2b;         // nibble a = base of allele a; 4 = the unobserved allele, 0xF = none
    uint32_t AG;          // A | G << 8 | all15 << 16   (A = G = 0: site skipped)
};

__device__ __forceinline__ unsigned long long ld_state_t(const unsigned long long* p)
{
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_state_t(unsigned long long* p, unsigned long long v)
{
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
#define TS_PACK(flag, g, r) (((unsigned long long)(flag) << 62) | ((unsigned long long)(g) << 31) | (unsigned long long)(r))
#define TS_FLAG(w) ((int)((w) >> 62))
#define TS_G(w) ((long long)(((w) >> 31) & 0x7FFFFFFFull))
#define TS_R(w) ((long long)((w)&0x7FFFFFFFull))

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void bulk_store(void* gdst, uint32_t ssrc, uint32_t bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(ssrc), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// low `k` bits set, 0 <= k <= 32
__device__ __forceinline__ uint32_t low_bits(int k) { return __funnelshift_rc(0xFFFFFFFFu, 0u, 32 - k); }

// One cell of the count-level sampler: returns the four 8-bit base counts (A | C << 8 | G << 16 | T << 24).
// Draws: block 0 = {x,y: depth; z: number of errors; w: haplotype bits 0..31}, block 1 = {x: haplotype
// bits 32..63; y,z,w: placement of errors 1..3}; rarer needs (depth > 64, > 3 errors) continue with
// blocks 2.. of the same (site, sample) counter.
struct TileRng {
    Key key;
    const uint2* alias;  // shared: [256] (t56 << 8 | alias) as {lo, hi}
    const uint4* cdf_e;  // shared: [256] P(E <= j | n) * 2^32, j = 0..3
    double e;
    int fixed_depth;     // >= 0: every cell has this depth; < 0: Poisson via the alias table
    bool has_err;
};

__device__ __forceinline__ uint32_t tile_sample_cell(const TileRng& R, int64_t site, uint32_t sample, uint32_t gt)
{
    const int g0 = gt & 0xF, g1 = gt >> 4;
    const u32x4 b0 = draw(R.key, site, sample, 0, P_COUNTS, 0);
    const u32x4 b1 = draw(R.key, site, sample, 0, P_COUNTS, 1);
    int n;
    if (R.fixed_depth >= 0) {
        n = R.fixed_depth;
    } else { // Walker alias over 256 columns, 64-bit uniform: column = top byte, 56 bits against the column's threshold
        const uint32_t col = b0.x >> 24;
        const uint2 en = R.alias[col];
        const unsigned long long frac = ((unsigned long long)__funnelshift_l(b0.y, b0.x, 8) << 32) | (b0.y << 8);
        const unsigned long long thr = ((unsigned long long)en.y << 32) | (en.x & 0xFFFFFF00u);
        n = frac < thr ? (int)col : (int)(en.x & 0xFFu);
    }
    if (g0 == VGL_GT_MISSING || g1 == VGL_GT_MISSING) n = 0; // depth is drawn but discarded (vcfgl.cpp:371-379)
    int E = 0;
    if (R.has_err) {
        const uint4 c = R.cdf_e[n];
        E = (b0.z >= c.x) + (b0.z >= c.y) + (b0.z >= c.z) + (b0.z >= c.w);
        if (E == 4) E = binom_inversion(n, R.e, u01_32(b0.z)); // beyond the table: exact inversion
    }
    int k0 = n; // reads drawn from haplotype 0
    if (g0 != g1) {
        k0 = __popc(b0.w & low_bits(min(n, 32))) + __popc(b1.x & low_bits(min(max(n - 32, 0), 32)));
        if (n > 64) {
            Stream st;
            st.init(R.key, site, sample, 0, P_COUNTS);
            st.block = 2;
            for (int left = n - 64; left > 0; left -= 32) k0 += __popc(st.next() & low_bits(min(left, 32)));
        }
    }
    uint32_t ad = ((uint32_t)k0 << (8 * g0)) + ((uint32_t)(n - k0) << (8 * g1));
    if (E > 0) { // place the errors: read j of the rem left is hit, it belongs to haplotype 0 w.p. rem0/rem; wrong base uniform
        int rem0 = k0, rem = n;
        Stream st;
        st.init(R.key, site, sample, 0, P_COUNTS);
        st.block = 10;
        for (int i = 0; i < E; ++i) {
            const uint32_t r = i == 0 ? b1.y : i == 1 ? b1.z : i == 2 ? b1.w : st.next();
            const uint32_t j = mulhi32(r, 3u * (uint32_t)rem);
            const uint32_t which = (j * 0xAAABu) >> 17; // j / 3 for j < 2^15 (rem <= 255)
            const uint32_t woff = j - 3u * which;
            const bool from0 = (int)which < rem0;
            const int truth = from0 ? g0 : g1;
            rem0 -= from0;
            --rem;
            const int wrong = (truth + 1 + (int)woff) & 3;
            ad += (1u << (8 * wrong)) - (1u << (8 * truth));
        }
    }
    return ad;
}

// GL / PL of one cell from its scores, scattered into the warp's stage slice in allele order
// (gl_methods.cpp:338-357, vcfgl.cpp:907-939).  ALL15: every base pair is a genotype of the site.
template <bool ALL15>
__device__ __forceinline__ void tile_emit_cell(const float (&q)[15], const TSite& ts, char* cell_g, bool has_gl, bool has_pl)
{
    float v[15];
#pragma unroll
    for (int k = 0; k < 15; ++k) v[k] = neg_div10_fast(q[k]); // gl_methods.cpp:343
    float mx = -CUDART_INF_F;
#pragma unroll
    for (int k = 0; k < 15; ++k) {
        const uint32_t off = (ts.slot[k >> 2] >> (8 * (k & 3))) & 0xFFu;
        mx = fmaxf(mx, (ALL15 || off != 0xFFu) ? v[k] : -CUDART_INF_F);
    }
#pragma unroll
    for (int k = 0; k < 15; ++k) {
        const uint32_t off = (ts.slot[k >> 2] >> (8 * (k & 3))) & 0xFFu;
        if (ALL15 || off != 0xFFu) {
            const float g = __fsub_rn(v[k], mx);
            char* dst = cell_g + off;
            if (has_gl) *reinterpret_cast<float*>(dst) = g;
            if (has_pl) *reinterpret_cast<int*>(dst + TILE_WST_G * 4) = pl_from_gl_magic(g);
        }
    }
}

__global__ void __launch_bounds__(TILE_BLOCK, 3) k_tile_m1f(const __grid_constant__ DevParams p)
{
    extern __shared__ __align__(16) unsigned char tile_smem[];
    // layout: alias [256] u64 | cdf_e [256] uint4 | stage [8 warps][G plane, PL plane, R plane] | st [64] | tot [64][4] | cnt [cap]
    uint2* alias = reinterpret_cast<uint2*>(tile_smem);
    uint4* cdf_e = reinterpret_cast<uint4*>(tile_smem + 2048);
    uint32_t* stage = reinterpret_cast<uint32_t*>(tile_smem + 2048 + 4096);
    constexpr int WST = 2 * TILE_WST_G + TILE_WST_R; // words per warp
    TSite* st = reinterpret_cast<TSite*>(stage + TILE_WARPS * WST);
    int* tot = reinterpret_cast<int*>(st + TILE_MAX_SITES);
    uint32_t* cnt = reinterpret_cast<uint32_t*>(tot + TILE_MAX_SITES * 4);
    __shared__ int64_t s_base[2];
    __shared__ int s_tile, s_tile_g, s_tile_r;
    __shared__ int wsum[4];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int S = p.S, S4 = (S + 3) & ~3, T = p.sites_per_tile;
    for (int i = tid; i < 256; i += TILE_BLOCK) {
        alias[i] = reinterpret_cast<const uint2*>(p.pois_alias)[i];
        cdf_e[i] = reinterpret_cast<const uint4*>(p.err_cdf)[i];
    }
    TileRng R;
    R.key.k0 = p.k0;
    R.key.k1 = p.k1;
    R.alias = alias;
    R.cdf_e = cdf_e;
    R.e = p.error_rate;
    R.fixed_depth = p.depth_mode == VGL_DEPTH_FIXED ? (int)p.depth_mean : -1;
    R.has_err = p.error_rate > 0.0;
    // iv / S4 for iv < 2^16 (exact: S4 >= 32, iv < 65536)
    const uint32_t inv_s4 = (uint32_t)(((1ull << 32) + S4 - 1) / S4);
    const bool explode = p.do_unobserved >= 3;
    const bool add_unobs = p.do_unobserved == 1 || p.do_unobserved == 2 || p.do_unobserved == 4 || p.do_unobserved == 5;
    const bool has_gl = p.gl != nullptr, has_pl = p.pl != nullptr, has_ad = p.ad != nullptr;
    uint32_t* const wg = stage + warp * WST; // this warp's GL slice; PL at +TILE_WST_G, AD at +2*TILE_WST_G
    uint32_t* const wr = wg + 2 * TILE_WST_G;
    bool pending = false; // this warp has bulk copies in flight that read its slice

    for (;;) {
        __syncthreads(); // previous tile fully done with st / tot / cnt
        if (tid == 0) s_tile = (int)atomicAdd(p.ticket, 1u);
        for (int i = tid; i < TILE_MAX_SITES * 4; i += TILE_BLOCK) tot[i] = 0;
        __syncthreads();
        const int tile = s_tile;
        if (tile >= p.n_tiles) break;
        const int site0 = tile * T;
        const int nsl = min(T, p.n_sites - site0);
        const int nv = nsl * S4; // virtual cells
        const int64_t cell0 = (int64_t)site0 * S;
        const uint8_t* __restrict__ gt_t = p.gt + cell0;
        int32_t* __restrict__ dp_t = p.dp + cell0;

        // ---------------- phase A: sample, FORMAT/DP, per-site totals
        for (int iv0 = 0; iv0 < nv; iv0 += TILE_BLOCK) {
            const int iv = iv0 + tid;
            int sl = (int)__umulhi((uint32_t)iv, inv_s4);
            const int v = iv - sl * S4;
            uint32_t ad = 0;
            if (iv < nv && v < S) {
                const int ci = sl * S + v;
                ad = tile_sample_cell(R, p.first_site + site0 + sl, (uint32_t)v, gt_t[ci]);
                dp_t[ci] = (int)__vsadu4(ad, 0u); // sum of the four byte counts
            }
            if (iv < nv) cnt[iv] = ad;
            // a warp of 32 consecutive slots spans at most two sites (S4 >= 32): packed 16-bit fields,
            // two masked REDUX rounds per site, then eight lanes add the eight sums
            const int first = __shfl_sync(0xffffffffu, sl, 0);
            const uint32_t w01 = __byte_perm(ad, 0u, 0x4140), w23 = __byte_perm(ad, 0u, 0x4342);
            const bool in0 = sl == first;
            const uint32_t a01 = __reduce_add_sync(0xffffffffu, in0 ? w01 : 0u), a23 = __reduce_add_sync(0xffffffffu, in0 ? w23 : 0u);
            const uint32_t b01 = __reduce_add_sync(0xffffffffu, in0 ? 0u : w01), b23 = __reduce_add_sync(0xffffffffu, in0 ? 0u : w23);
            if (lane < 8) {
                const uint32_t w = (lane & 4) ? ((lane & 2) ? b23 : b01) : ((lane & 2) ? a23 : a01);
                const uint32_t val = (lane & 1) ? (w >> 16) : (w & 0xFFFFu);
                const int ts = first + (lane >> 2);
                if (val && ts < nsl) atomicAdd(&tot[ts * 4 + (lane & 3)], (int)val);
            }
        }
        __syncthreads();

        // ---------------- phase B: per-site record (vcfgl.cpp:396-404, 665-782, 806-843 INFO part)
        int my_g = 0, my_r = 0; // this site's block sizes in 4-byte elements (padded to 16 B)
        if (tid < nsl) {
            const int* t = tot + tid * 4;
            const int dp = t[0] + t[1] + t[2] + t[3];
            vgl_site_out o;
            o.skip_code = 0;
            o.n_alleles = o.n_alleles_observed = o.n_genotypes = 0;
#pragma unroll
            for (int i = 0; i < 8; ++i) o.alleles2acgt[i] = o.acgt2alleles[i] = -1;
            o.info_dp = dp;
#pragma unroll
            for (int i = 0; i < 5; ++i) { o.info_ad[i] = o.info_adf[i] = o.info_adr[i] = 0; o.qs[i] = 0.0f; }
#pragma unroll
            for (int i = 0; i < 16; ++i) o.i16[i] = 0.0f;
            o._pad = 0;
            o.g_off = o.r_off = 0; // patched after the look-back
            int b2a[5] = {-1, -1, -1, -1, -1};
            uint32_t a2b = 0xFFFFFFFFu;
            if (dp == 0) {
                if (p.rm_empty) o.skip_code = -4;
                else if (!p.do_gvcf) {
                    if (p.do_unobserved <= 2) { o.n_alleles = 1; o.n_genotypes = 1; o.n_alleles_observed = 0; }
                    else if (p.do_unobserved == 3) { o.n_alleles = 4; o.n_genotypes = 10; o.n_alleles_observed = 4; }
                    else { o.n_alleles = 5; o.n_genotypes = 15; o.n_alleles_observed = 4; }
                }
            } else {
                int n_obs = 0;
#pragma unroll
                for (int b = 0; b < 4; ++b) n_obs += t[b] > 0;
                if (p.rm_invar_sim && n_obs == 1) {
                    o.skip_code = -3;
                } else {
                    int n_alleles = 0;
#pragma unroll
                    for (int b = 0; b < 4; ++b) { // stable sort by INFO/AD, descending (vcfgl.cpp:700-718)
                        int rank = 0;
#pragma unroll
                        for (int x = 0; x < 4; ++x) rank += (t[x] > t[b]) || (t[x] == t[b] && x < b);
                        if (t[b] > 0 || explode) {
                            b2a[b] = rank;
                            o.acgt2alleles[b] = (int8_t)rank;
                            a2b = (a2b & ~(0xFu << (4 * rank))) | ((uint32_t)b << (4 * rank));
                            ++n_alleles;
                        }
                    }
                    o.n_alleles_observed = n_alleles;
                    if (add_unobs) {
                        b2a[4] = n_alleles;
                        o.acgt2alleles[4] = (int8_t)n_alleles;
                        a2b = (a2b & ~(0xFu << (4 * n_alleles))) | (4u << (4 * n_alleles));
                        ++n_alleles;
                    }
                    o.n_alleles = n_alleles;
                    o.n_genotypes = n_alleles * (n_alleles + 1) / 2;
#pragma unroll
                    for (int a = 0; a < 5; ++a) {
                        const int b = (int)((a2b >> (4 * a)) & 0xF);
                        o.alleles2acgt[a] = b == 0xF ? (int8_t)-1 : (int8_t)b;
                        if (a < n_alleles && b < 4 && (p.tag_mask & VGL_TAG_INFO_AD)) o.info_ad[a] = t[b];
                    }
                }
            }
            // dp == 0 sites keep all-missing blocks: their "alleles" carry no base (a2b stays 0xF..F -> counts read as 0)
            const bool keep = o.skip_code == 0 && o.n_alleles > 0;
            TSite ts;
            const uint64_t pm = make_pairmap(b2a);
            bool all15 = true;
#pragma unroll
            for (int w = 0; w < 4; ++w) {
                uint32_t x = 0;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int pair = 4 * w + k;
                    uint32_t slot = pair < 15 ? (uint32_t)((pm >> (4 * pair)) & 0xF) : 0xFu;
                    if (pair < 15 && slot == 0xFu) all15 = false;
                    x |= (slot == 0xFu ? 0xFFu : slot * 4u) << (8 * k);
                }
                ts.slot[w] = x;
            }
            // alleles without a base (none / dp == 0 placeholder) read byte 4 = 0 in the AD permute
            uint32_t a2b4 = 0;
#pragma unroll
            for (int a = 0; a < 5; ++a) {
                const uint32_t b = (a2b >> (4 * a)) & 0xF;
                a2b4 |= (b < 4 ? b : 4u) << (4 * a);
            }
            ts.a2b = a2b4;
            ts.AG = keep ? ((uint32_t)o.n_alleles | ((uint32_t)o.n_genotypes << 8) | ((all15 && dp > 0) ? 1u << 16 : 0u)) : 0u;
            ts.g_rel = ts.r_rel = 0;
            if (keep) {
                my_g = (S * o.n_genotypes + 3) & ~3;
                my_r = (S * o.n_alleles + 3) & ~3;
            }
            st[tid] = ts;
            p.sites[site0 + tid] = o;
        }
        // exclusive scan of the tile's block sizes: at most 64 sites -> warps 0..1
        int ig = my_g, ir = my_r;
        if (tid < TILE_MAX_SITES) {
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const int tg = __shfl_up_sync(0xffffffffu, ig, off);
                const int tr = __shfl_up_sync(0xffffffffu, ir, off);
                if (lane >= off) { ig += tg; ir += tr; }
            }
            if (lane == 31) { wsum[2 * warp] = ig; wsum[2 * warp + 1] = ir; }
        }
        __syncthreads();
        if (tid < TILE_MAX_SITES) {
            const int pre_g = warp ? wsum[0] : 0, pre_r = warp ? wsum[1] : 0;
            const int tile_g = wsum[0] + wsum[2], tile_r = wsum[1] + wsum[3];
            if (tid < nsl) {
                st[tid].g_rel = pre_g + ig - my_g;
                st[tid].r_rel = pre_r + ir - my_r;
            }
            // decoupled look-back (warp 0): base offset of this tile = total size of all earlier tiles
            if (warp == 0) {
                if (lane == 0) st_state_t(p.tile_state + tile, TS_PACK(1, tile_g, tile_r));
                int64_t bg = 0, br = 0;
                int look = tile - 1;
                while (look >= 0) {
                    const int idx = look - lane;
                    unsigned long long w = TS_PACK(2, 0, 0); // lanes before tile 0 act as a zero inclusive prefix
                    if (idx >= 0) {
                        do { w = ld_state_t(p.tile_state + idx); } while (TS_FLAG(w) == 0);
                    }
                    const unsigned incl = __ballot_sync(0xffffffffu, TS_FLAG(w) == 2);
                    const int stop = incl ? (__ffs(incl) - 1) : 31; // nearest inclusive prefix
                    int64_t g = lane <= stop ? TS_G(w) : 0, r = lane <= stop ? TS_R(w) : 0;
#pragma unroll
                    for (int off = 16; off > 0; off >>= 1) {
                        g += __shfl_xor_sync(0xffffffffu, g, off);
                        r += __shfl_xor_sync(0xffffffffu, r, off);
                    }
                    bg += g;
                    br += r;
                    if (incl) break;
                    look -= 32;
                }
                if (lane == 0) {
                    st_state_t(p.tile_state + tile, TS_PACK(2, bg + tile_g, br + tile_r));
                    s_base[0] = bg;
                    s_base[1] = br;
                    s_tile_g = tile_g;
                    s_tile_r = tile_r;
                    if (tile == p.n_tiles - 1) { p.totals[0] = bg + tile_g; p.totals[1] = br + tile_r; }
                }
            }
        }
        __syncthreads();
        if (tid < nsl) {
            p.sites[site0 + tid].g_off = s_base[0] + st[tid].g_rel;
            p.sites[site0 + tid].r_off = s_base[1] + st[tid].r_rel;
        }

        // ---------------- phase C: score + emit, one warp per 32 virtual cells
        const int tile_g = s_tile_g, tile_r = s_tile_r;
        float* const gl_t = p.gl ? p.gl + s_base[0] : nullptr;
        int32_t* const pl_t = p.pl ? p.pl + s_base[0] : nullptr;
        int32_t* const ad_t = p.ad ? p.ad + s_base[1] : nullptr;
        for (int iv0 = warp * 32; iv0 < nv; iv0 += TILE_BLOCK) {
            const int iv = iv0 + lane;
            int sl = (int)__umulhi((uint32_t)iv, inv_s4);
            int v = iv - sl * S4;
            if (sl >= nsl) { sl = nsl - 1; v = S4; } // past the tile: sits at the end of the last block
            const TSite ts = st[sl];
            const int A = (int)(ts.AG & 0xFF), G = (int)((ts.AG >> 8) & 0xFF);
            const bool live = v < S && G > 0;
            const int vv = min(v, S);
            const int gpos = ts.g_rel + vv * G, rpos = ts.r_rel + vv * A;
            const int gend = v < S ? gpos + G : ts.g_rel + ((S * G + 3) & ~3);
            const int rend = v < S ? rpos + A : ts.r_rel + ((S * A + 3) & ~3);
            const int g_lo = __shfl_sync(0xffffffffu, gpos, 0), g_hi = __shfl_sync(0xffffffffu, gend, 31);
            const int r_lo = __shfl_sync(0xffffffffu, rpos, 0), r_hi = __shfl_sync(0xffffffffu, rend, 31);
            char* const cell_g = reinterpret_cast<char*>(wg + (gpos - g_lo));
            uint32_t* const cell_r = wr + (rpos - r_lo);
            if (pending) { // the previous copies must have finished reading the slice
                if (lane == 0) bulk_wait_read();
                __syncwarp();
            }
            const uint32_t c4 = live ? cnt[iv] : 0u;
            const int c0 = (int)(c4 & 0xFF), c1 = (int)((c4 >> 8) & 0xFF), c2 = (int)((c4 >> 16) & 0xFF), c3 = (int)(c4 >> 24);
            const int n = c0 + c1 + c2 + c3;
            if (live) {
                float q[15];
                m1f_scores_noclamp(n, c0, c1, c2, c3, p.m1_bsum, p.m1_het, q);
                if (ts.AG >> 16) tile_emit_cell<true>(q, ts, cell_g, has_gl, has_pl);
                else tile_emit_cell<false>(q, ts, cell_g, has_gl, has_pl);
                if (has_ad) { // AD in allele order (vcfgl.cpp:806-831): byte permute, selector 4 reads 0
#pragma unroll
                    for (int a = 0; a < 5; ++a)
                        if (a < A) cell_r[a] = __byte_perm(c4, 0u, ((ts.a2b >> (4 * a)) & 0xFu) | 0x4440u);
                }
                if (n == 0) { // gl_methods.cpp:359-366
#pragma unroll 1
                    for (int g = 0; g < G; ++g) {
                        reinterpret_cast<uint32_t*>(cell_g)[g] = VGL_F32_MISSING_BITS;
                        reinterpret_cast<uint32_t*>(cell_g)[TILE_WST_G + g] = (uint32_t)VGL_I32_MISSING;
                    }
                }
            } else if (v == S) { // first dead slot of a site: zero the block's padding
#pragma unroll 1
                for (int g = gpos; g < gend; ++g) {
                    reinterpret_cast<uint32_t*>(cell_g)[g - gpos] = 0u;
                    reinterpret_cast<uint32_t*>(cell_g)[TILE_WST_G + g - gpos] = 0u;
                }
#pragma unroll 1
                for (int a = rpos; a < rend; ++a) cell_r[a - rpos] = 0u;
            }
            fence_async_smem();
            __syncwarp();
            if (lane == 0) {
                const uint32_t gb = (uint32_t)(g_hi - g_lo) * 4u, rb = (uint32_t)(r_hi - r_lo) * 4u;
                if (gb) {
                    if (has_gl) bulk_store(gl_t + g_lo, smem_u32(wg), gb);
                    if (has_pl) bulk_store(pl_t + g_lo, smem_u32(wg + TILE_WST_G), gb);
                }
                if (rb && has_ad) bulk_store(ad_t + r_lo, smem_u32(wr), rb);
                bulk_commit();
            }
            pending = true;
        }
        (void)tile_g;
        (void)tile_r;
    }
    if (lane == 0) bulk_wait_all(); // global writes of this warp's last copies complete before exit
}

void launch_tile_m1f(const DevParams& p, cudaStream_t st, int n_sms)
{
    const int S4 = (p.S + 3) & ~3;
    const int cap = S4 > TILE_CELLS ? S4 : TILE_CELLS;
    const size_t dyn = 2048 + 4096 + (size_t)TILE_WARPS * (2 * TILE_WST_G + TILE_WST_R) * 4 + TILE_MAX_SITES * sizeof(TSite) +
                       TILE_MAX_SITES * 16 + (size_t)cap * 4;
    cudaFuncSetAttribute(k_tile_m1f, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
    cudaFuncSetAttribute(k_tile_m1f, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    int per_sm = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_tile_m1f, TILE_BLOCK, dyn);
    if (per_sm < 1) per_sm = 1;
    int grid = n_sms * per_sm;
    if (grid > p.n_tiles) grid = p.n_tiles;
    k_tile_m1f<<<grid, TILE_BLOCK, dyn, st>>>(p);
}

// largest S the tile kernel takes: one site's counts must fit the shared-memory cache
int tile_m1f_max_samples() { return 16384; }
int tile_m1f_sites_per_tile(int S)
{
    const int S4 = (S + 3) & ~3;
    int T = TILE_CELLS / S4;
    return T < 1 ? 1 : (T > TILE_MAX_SITES ? TILE_MAX_SITES : T);
}

} // namespace vgl
