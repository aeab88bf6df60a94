// k_fused_m1f -- the whole hot path in ONE kernel for the headline configuration: native RNG,
// GL model 1 with a run-constant quality score (--error-qs 0/1), count-level sampler.
//
// Persistent CTAs (one wave, a multiple of the SM count) pull TILES of whole sites from an atomic
// ticket.  Per tile:
//   phase 1  every cell of the tile is sampled (counts_sampler.cuh); FORMAT/DP is written; per-site
//            totals are reduced in shared memory (warp REDUX + shared atomics)
//   phase 2  one thread per site: allele order, unobserved allele, skip code, INFO tags
//            (vcfgl.cpp:665-782); the tile's output size is published and its base offset obtained by
//            decoupled look-back over earlier tiles -> deterministic, compact, in-order layout
//   phase 3  every cell is sampled AGAIN (counter-based RNG: identical draws, nothing is stored) and
//            scored: errmod GL, PL, GP, AD/ADF/ADR in allele order, staged per 256-cell chunk in
//            shared memory and written with 128-bit stores
// HBM traffic = 1 B/cell in, the tag planes out; nothing intermediate touches DRAM.
#include "counts_sampler.cuh"
#include "m1f.cuh"

#include <cstdio>
#include <cstdlib>

namespace vgl {

#define FUSED_BLOCK 256
#define FUSED_TILE_CELLS 1024
#define FUSED_MAX_SITES 128
#define FUSED_POIS_MAX 1024
#define WARP_STAGE (32 * 16 + 8) // 4-byte elements per warp slice: 32 cells x (15 values padded to 16) + alignment slack
#define FUSED_STAGE (8 * WARP_STAGE)

struct TileSite {
    uint64_t pairmap;
    int64_t g_off, r_off;
    int32_t A, G, skip;
    uint32_t a2b;   // nibble a = base of allele a (0xF none)
    double e;       // base-picking error probability of the site
    float l2, er;   // log2(1-e) (0 = use the slow binomial), e/(1-e)
};

__device__ __forceinline__ unsigned long long ld_state(const unsigned long long* p)
{
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_state(unsigned long long* p, unsigned long long v)
{
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// tile_state word: flag (2 bits: 1 = aggregate, 2 = inclusive prefix) | g elements (31 bits) | r elements (31 bits)
#define TS_PACK(flag, g, r) (((unsigned long long)(flag) << 62) | ((unsigned long long)(g) << 31) | (unsigned long long)(r))
#define TS_FLAG(w) ((int)((w) >> 62))
#define TS_G(w) ((long long)(((w) >> 31) & 0x7FFFFFFFull))
#define TS_R(w) ((long long)((w)&0x7FFFFFFFull))

// Warp-wide copy of a staged span to global memory.  stage[i] holds plane element (base + i), base % 4 == 0;
// only elements in [lo, hi) belong to this warp.  Complete 16-byte chunks go out as 128-bit streaming
// stores without any bounds logic; the (at most 3 + 3) edge elements are written by six lanes.
__device__ __forceinline__ void store_span_w(uint32_t* __restrict__ plane, const uint32_t* stage, int64_t base, int64_t lo,
                                             int64_t hi, int lane)
{
    uint32_t* __restrict__ dst = plane + base;
    const int rlo = (int)(lo - base), rhi = (int)(hi - base);
    const int first_full = (rlo + 3) >> 2, last_full = rhi >> 2;
    const uint4* s4 = reinterpret_cast<const uint4*>(stage);
    uint4* d4 = reinterpret_cast<uint4*>(dst);
    // a warp slice holds at most WARP_STAGE / 4 = 130 chunks -> at most 5 per lane; fixed trip count, predicated
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        const int ch = first_full + lane + 32 * k;
        if (ch < last_full) __stcs(d4 + ch, s4[ch]);
    }
    const int head_end = min(first_full * 4, rhi);       // [rlo, head_end): before the first complete chunk
    const int tail_beg = max(last_full * 4, head_end);   // [tail_beg, rhi): after the last complete chunk
    if (lane < 4) {
        const int e = rlo + lane;
        if (e < head_end) dst[e] = stage[e];
    } else if (lane < 8) {
        const int e = tail_beg + lane - 4;
        if (e < rhi) dst[e] = stage[e];
    }
}

__global__ void __launch_bounds__(FUSED_BLOCK, 4) k_fused_m1f(const __grid_constant__ DevParams p)
{
    __shared__ __align__(16) uint32_t stage_a[FUSED_STAGE];
    __shared__ __align__(16) uint32_t stage_b[FUSED_STAGE];
    __shared__ TileSite st[FUSED_MAX_SITES];
    __shared__ int tot[FUSED_MAX_SITES][9]; // dp, ad[4], fwd[4]
    extern __shared__ unsigned long long dyn_smem[];
    unsigned long long* pois = dyn_smem;                                  // [pois_n] Poisson CDF
    unsigned long long* cnt_sm = dyn_smem + ((p.pois_n + 15) & ~15);     // [1024] per-cell AD counts of the tile
    unsigned long long* fwd_sm = cnt_sm + FUSED_TILE_CELLS;              // [1024] forward counts (strand runs only)
    unsigned short* guide = reinterpret_cast<unsigned short*>(p.sample_strand ? fwd_sm + FUSED_TILE_CELLS : fwd_sm); // [256]
    __shared__ int64_t s_base[2];
    __shared__ int s_tile;

    const int tid = threadIdx.x, lane = tid & 31;
    const int S = p.S, T = p.sites_per_tile;
    for (int i = tid; i < p.pois_n; i += FUSED_BLOCK) pois[i] = p.pois_cdf[i];
    __syncthreads();
    if (p.pois_n > 0) { // guide[j] = smallest k whose CDF exceeds every u with top byte j-1, i.e. a lower bound for top byte j
        int k = 0;
        const unsigned long long floor_u = (unsigned long long)tid << 56;
        while (k < p.pois_n - 1 && pois[k] <= floor_u) ++k;
        guide[tid] = (unsigned short)k;
    }
    CountsParams cp;
    cp.key.k0 = p.k0;
    cp.key.k1 = p.k1;
    cp.depth_mode = p.depth_mode;
    cp.depth_mean = p.depth_mean;
    cp.depth_means = p.depth_means;
    cp.pois_cdf = pois;
    cp.pois_guide = guide;
    cp.pois_n = p.pois_n;
    cp.sample_strand = p.sample_strand;
    // i / S for i < 1024 when a tile holds several sites (exact for S < 1024)
    const uint32_t inv_s = (uint32_t)(((1u << 20) + S - 1) / S);
    const bool explode = p.do_unobserved >= 3;
    const bool add_unobs = p.do_unobserved == 1 || p.do_unobserved == 2 || p.do_unobserved == 4 || p.do_unobserved == 5;

    for (;;) {
        __syncthreads(); // previous tile fully done with shared memory
        if (tid == 0) s_tile = (int)atomicAdd(p.ticket, 1u);
        __syncthreads();
        const int tile = s_tile;
        if (tile >= p.n_tiles) break;
        const int site0 = tile * T;
        const int nsl = min(T, p.n_sites - site0);
        const int ncell = nsl * S;
        const int64_t cell0 = (int64_t)site0 * S;
        const bool keep_counts = ncell <= FUSED_TILE_CELLS; // tile fits the shared-memory count cache

        // ---------------- phase 0: per-site constants
        for (int i = tid; i < nsl * 9; i += FUSED_BLOCK) (&tot[0][0])[i] = 0;
        if (tid < nsl) {
            double e = p.error_rate;
            if (p.error_qs == 1) { // per-site beta-distributed error rate (vcfgl.cpp:425-437)
                Stream bs;
                bs.init(cp.key, p.first_site + site0 + tid, 0xFFFFFFFFu, 0, P_SITE);
                e = beta_draw(bs, p.beta_a, p.beta_b);
            }
            st[tid].e = e;
            st[tid].l2 = (e > 0.0 && e <= 0.5) ? log2f((float)(1.0 - e)) : 0.0f;
            st[tid].er = (float)(e / (1.0 - e));
        }
        __syncthreads();

        // ---------------- phase 1: sample, FORMAT/DP, per-site totals
        for (int i0 = 0; i0 < ncell; i0 += FUSED_BLOCK) {
            const int i = i0 + tid;
            const bool live = i < ncell;
            int sl = 0, sample = i;
            if (T > 1) { sl = (int)(((uint32_t)i * inv_s) >> 20); sample = i - sl * S; }
            CellCounts cc;
            cc.n = 0; cc.ad = cc.fwd = 0;
            if (live) {
                const uint8_t gt = p.gt[cell0 + i];
                cc = sample_counts(cp, p.first_site + site0 + sl, (uint32_t)sample, gt, st[sl].e, st[sl].l2, st[sl].er);
                p.dp[cell0 + i] = cc.n;
                if (keep_counts) {
                    cnt_sm[i] = cc.ad;
                    if (p.sample_strand) fwd_sm[i] = cc.fwd;
                }
            } else {
                sl = -1;
            }
            // warp-aggregated reduction into the site totals: a warp of 32 consecutive cells spans at most
            // two sites when S >= 32 -> two masked REDUX rounds; otherwise per-thread shared atomics
            const int first = __shfl_sync(0xffffffffu, sl, 0);
            const bool two = __all_sync(0xffffffffu, sl < 0 || sl == first || sl == first + 1);
            int v[9];
            v[0] = cc.n;
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                v[1 + b] = (int)((cc.ad >> (16 * b)) & 0xFFFF);
                v[5 + b] = (int)((cc.fwd >> (16 * b)) & 0xFFFF);
            }
            if (two) {
                if (first >= 0) {
                    const bool any_second = __any_sync(0xffffffffu, sl == first + 1);
#pragma unroll
                    for (int k = 0; k < 9; ++k) {
                        if (k >= 5 && !p.sample_strand) break;
                        const int s0 = __reduce_add_sync(0xffffffffu, sl == first ? v[k] : 0);
                        if (lane == 0 && s0) atomicAdd(&tot[first][k], s0);
                        if (any_second) {
                            const int s1 = __reduce_add_sync(0xffffffffu, sl == first + 1 ? v[k] : 0);
                            if (lane == 0 && s1) atomicAdd(&tot[first + 1][k], s1);
                        }
                    }
                }
            } else if (live) {
#pragma unroll
                for (int k = 0; k < 9; ++k)
                    if (v[k]) atomicAdd(&tot[sl][k], v[k]);
            }
        }
        __syncthreads();

        // ---------------- phase 2: per-site record (vcfgl.cpp:396-404, 665-782, 806-843 INFO part)
        int my_g = 0, my_r = 0; // this site's block sizes in 4-byte elements (padded to 16 B)
        if (tid < nsl) {
            const int* t = tot[tid];
            const int dp = t[0];
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
                for (int b = 0; b < 4; ++b) n_obs += t[1 + b] > 0;
                if (p.rm_invar_sim && n_obs == 1) {
                    o.skip_code = -3;
                } else {
                    int n_alleles = 0;
#pragma unroll
                    for (int b = 0; b < 4; ++b) {
                        int rank = 0;
#pragma unroll
                        for (int x = 0; x < 4; ++x) rank += (t[1 + x] > t[1 + b]) || (t[1 + x] == t[1 + b] && x < b);
                        if (t[1 + b] > 0 || explode) {
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
                        if (a < n_alleles && b < 4) {
                            if (p.tag_mask & VGL_TAG_INFO_AD) o.info_ad[a] = t[1 + b];
                            if (p.tag_mask & VGL_TAG_INFO_ADF) o.info_adf[a] = t[5 + b];
                            if (p.tag_mask & VGL_TAG_INFO_ADR) o.info_adr[a] = t[1 + b] - t[5 + b];
                        }
                    }
                }
            }
            const bool keep = o.skip_code == 0 && o.n_alleles > 0;
            st[tid].A = keep ? o.n_alleles : 0;
            st[tid].G = keep ? o.n_genotypes : 0;
            st[tid].skip = !keep;
            st[tid].a2b = a2b;
            st[tid].pairmap = make_pairmap(b2a);
            if (o.skip_code == 0) {
                my_g = (S * o.n_genotypes + 3) & ~3;
                my_r = (S * o.n_alleles + 3) & ~3;
            }
            p.sites[site0 + tid] = o;
        }
        // exclusive scan of the tile's block sizes: at most 128 sites -> warps 0..3
        int* wsum = reinterpret_cast<int*>(stage_a); // [4][2] scratch
        int ig = my_g, ir = my_r;
        if (tid < FUSED_MAX_SITES) {
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const int tg = __shfl_up_sync(0xffffffffu, ig, off);
                const int tr = __shfl_up_sync(0xffffffffu, ir, off);
                if (lane >= off) { ig += tg; ir += tr; }
            }
            if (lane == 31) { wsum[2 * (tid >> 5)] = ig; wsum[2 * (tid >> 5) + 1] = ir; }
        }
        __syncthreads();
        int ex_g = 0, ex_r = 0;
        if (tid < FUSED_MAX_SITES) {
            int pre_g = 0, pre_r = 0, tile_g = 0, tile_r = 0;
#pragma unroll
            for (int w = 0; w < 4; ++w) {
                if (w < (tid >> 5)) { pre_g += wsum[2 * w]; pre_r += wsum[2 * w + 1]; }
                tile_g += wsum[2 * w];
                tile_r += wsum[2 * w + 1];
            }
            ex_g = pre_g + ig - my_g;
            ex_r = pre_r + ir - my_r;
            // decoupled look-back (warp 0): base offset of this tile = total size of all earlier tiles
            if (tid < 32) {
                if (lane == 0) st_state(p.tile_state + tile, TS_PACK(1, tile_g, tile_r));
                int64_t bg = 0, br = 0;
                int look = tile - 1;
                while (look >= 0) {
                    const int idx = look - lane;
                    unsigned long long w = TS_PACK(2, 0, 0); // lanes before tile 0 act as a zero inclusive prefix
                    if (idx >= 0) {
                        do { w = ld_state(p.tile_state + idx); } while (TS_FLAG(w) == 0);
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
                    st_state(p.tile_state + tile, TS_PACK(2, bg + tile_g, br + tile_r));
                    s_base[0] = bg;
                    s_base[1] = br;
                    if (tile == p.n_tiles - 1) { p.totals[0] = bg + tile_g; p.totals[1] = br + tile_r; }
                }
            }
        }
        __syncthreads();
        if (tid < nsl) {
            const int64_t go = s_base[0] + ex_g, ro = s_base[1] + ex_r;
            st[tid].g_off = go;
            st[tid].r_off = ro;
            p.sites[site0 + tid].g_off = go;
            p.sites[site0 + tid].r_off = ro;
        }
        __syncthreads();

        // ---------------- phase 3: score + emit.  Each WARP stages the contiguous output span of its 32
        // cells in its own shared-memory slice and copies it out itself: no CTA barriers in this phase.
        uint32_t* const wa = stage_a + (tid >> 5) * WARP_STAGE; // GL, later AD / ADR
        uint32_t* const wb = stage_b + (tid >> 5) * WARP_STAGE; // PL, later GP, ADF
        for (int i0 = 0; i0 < ncell; i0 += FUSED_BLOCK) {
            const int i = i0 + tid;
            const bool live = i < ncell;
            int sl = nsl - 1, sample = S; // dead lanes sit just past the last cell of the tile
            if (live) {
                sl = 0; sample = i;
                if (T > 1) { sl = (int)(((uint32_t)i * inv_s) >> 20); sample = i - sl * S; }
            }
            const TileSite& ts = st[sl];
            const int A = ts.A, G = ts.G;
            const bool act = live && !ts.skip;
            const bool pad_owner = act && sample == S - 1;
            const int g_pad = pad_owner ? ((S * G + 3) & ~3) - S * G : 0;
            const int r_pad = pad_owner ? ((S * A + 3) & ~3) - S * A : 0;
            int64_t gpos = ts.g_off + (int64_t)sample * G, rpos = ts.r_off + (int64_t)sample * A;
            if (!live && !ts.skip) { // end of the last site's padded block
                gpos = ts.g_off + ((S * G + 3) & ~3);
                rpos = ts.r_off + ((S * A + 3) & ~3);
            }
            // the warp's span: first lane's start .. last lane's end (positions are monotone in the cell index)
            const int64_t g_lo = __shfl_sync(0xffffffffu, gpos, 0), r_lo = __shfl_sync(0xffffffffu, rpos, 0);
            const int64_t g_hi = __shfl_sync(0xffffffffu, gpos + (live ? G + g_pad : 0), 31);
            const int64_t r_hi = __shfl_sync(0xffffffffu, rpos + (live ? A + r_pad : 0), 31);
            const int64_t g_base = g_lo & ~3ll, r_base = r_lo & ~3ll;
            const int go = (int)(gpos - g_base), ro = (int)(rpos - r_base); // this cell's offset in the warp slice
            CellCounts cc;
            cc.n = 0; cc.ad = cc.fwd = 0;
            if (act) {
                const int64_t site = p.first_site + site0 + sl;
                if (keep_counts) {
                    cc.ad = cnt_sm[i];
                    cc.fwd = p.sample_strand ? fwd_sm[i] : 0ull;
                    cc.n = (int)((cc.ad & 0xFFFF) + ((cc.ad >> 16) & 0xFFFF) + ((cc.ad >> 32) & 0xFFFF) + (cc.ad >> 48));
                } else { // big sites: sample again (counter-based RNG: identical draws, nothing stored)
                    cc = sample_counts(cp, site, (uint32_t)sample, p.gt[cell0 + i], ts.e, ts.l2, ts.er);
                }
                if (cc.n == 0) { // gl_methods.cpp:359-366
#pragma unroll 1
                    for (int g = 0; g < G; ++g) { wa[go + g] = VGL_F32_MISSING_BITS; wb[go + g] = (uint32_t)VGL_I32_MISSING; }
                } else {
                    uint64_t ad = cc.ad;
                    int nn = cc.n;
                    if (nn > 255) { ad = subsample_counts_255(cp, site, (uint32_t)sample, ad, nn); nn = 255; } // errmod.c:156-159
                    float q[15];
                    m1f_scores(nn, (int)(ad & 0xFFFF), (int)((ad >> 16) & 0xFFFF), (int)((ad >> 32) & 0xFFFF), (int)(ad >> 48),
                               p.m1_bsum, p.m1_het, q);
                    if (p.fast_div) {
#pragma unroll
                        for (int k = 0; k < 15; ++k) q[k] = neg_div10_fast(q[k]); // gl_methods.cpp:343
                    } else {
#pragma unroll 1
                        for (int k = 0; k < 15; ++k) q[k] = neg_div10_ref(q[k]);
                    }
                    const uint32_t pm_lo = (uint32_t)ts.pairmap, pm_hi = (uint32_t)(ts.pairmap >> 32);
                    float mx = -CUDART_INF_F;
#pragma unroll
                    for (int k = 0; k < 15; ++k) {
                        const uint32_t slot = ((k < 8 ? pm_lo >> (4 * k) : pm_hi >> (4 * (k - 8))) & 0xF);
                        mx = fmaxf(mx, slot != 0xF ? q[k] : -CUDART_INF_F);
                    }
#pragma unroll
                    for (int k = 0; k < 15; ++k) { // gl_methods.cpp:355-357, vcfgl.cpp:907-939
                        const uint32_t slot = ((k < 8 ? pm_lo >> (4 * k) : pm_hi >> (4 * (k - 8))) & 0xF);
                        if (slot != 0xF) {
                            const float v = __fsub_rn(q[k], mx);
                            wa[go + slot] = __float_as_uint(v);
                            wb[go + slot] = (uint32_t)pl_from_gl(v);
                        }
                    }
                }
#pragma unroll 1
                for (int g = 0; g < g_pad; ++g) { wa[go + G + g] = 0u; wb[go + G + g] = 0u; }
            }
            __syncwarp();
            if (p.gl) store_span_w(reinterpret_cast<uint32_t*>(p.gl), wa, g_base, g_lo, g_hi, lane);
            if (p.pl) store_span_w(reinterpret_cast<uint32_t*>(p.pl), wb, g_base, g_lo, g_hi, lane);
            if (p.gp) { // vcfgl.cpp:941-970
                __syncwarp();
                if (act) {
                    if (cc.n == 0) {
                        for (int g = 0; g < G; ++g) wb[go + g] = VGL_F32_MISSING_BITS;
                    } else {
                        float sum = 0.0f;
                        for (int g = 0; g < G; ++g) {
                            const float v = __double2float_rn(exp10((double)__uint_as_float(wa[go + g])));
                            wb[go + g] = __float_as_uint(v);
                            sum = __fadd_rn(sum, v);
                        }
                        for (int g = 0; g < G; ++g) wb[go + g] = __float_as_uint(__fdiv_rn(__uint_as_float(wb[go + g]), sum));
                    }
                }
                __syncwarp();
                store_span_w(reinterpret_cast<uint32_t*>(p.gp), wb, g_base, g_lo, g_hi, lane);
            }
            // AD / ADF / ADR in allele order (vcfgl.cpp:806-831)
            if (p.ad || p.adf || p.adr) {
                __syncwarp();
                if (act) {
                    const uint32_t a2b = ts.a2b;
                    const bool want_f = p.adf || p.adr;
#pragma unroll
                    for (int a = 0; a < 5; ++a) {
                        if (a < A) {
                            const uint32_t b = (a2b >> (4 * a)) & 0xF;
                            const int sh = (int)(b & 3) * 16;
                            const uint32_t c = b < 4 ? (uint32_t)(cc.ad >> sh) & 0xFFFFu : 0u;
                            wa[ro + a] = c;
                            if (want_f) wb[ro + a] = b < 4 ? (uint32_t)(cc.fwd >> sh) & 0xFFFFu : 0u;
                        }
                    }
#pragma unroll 1
                    for (int a = 0; a < r_pad; ++a) { wa[ro + A + a] = 0u; wb[ro + A + a] = 0u; }
                }
                __syncwarp();
                if (p.ad) store_span_w(reinterpret_cast<uint32_t*>(p.ad), wa, r_base, r_lo, r_hi, lane);
                if (p.adf) store_span_w(reinterpret_cast<uint32_t*>(p.adf), wb, r_base, r_lo, r_hi, lane);
                if (p.adr) {
                    __syncwarp();
                    if (act)
                        for (int a = 0; a < A; ++a) wa[ro + a] -= wb[ro + a];
                    __syncwarp();
                    store_span_w(reinterpret_cast<uint32_t*>(p.adr), wa, r_base, r_lo, r_hi, lane);
                }
            }
            __syncwarp();
        }
    }
}

// exhaustive device self-test of the arithmetic shortcuts (all 2^32 float bit patterns)
// diagnostic: range of q >= 0 where the UNGUARDED 3-instruction division differs from __fdiv_rn
__global__ void k_div10_range(unsigned int* lo_hi)
{
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i <= 0x7F800000ull; i += stride) {
        const float x = __uint_as_float((unsigned int)i);
        if (__float_as_uint(neg_div10_fast(x)) != __float_as_uint(neg_div10_ref(x))) {
            atomicMin(lo_hi, (unsigned int)i);
            atomicMax(lo_hi + 1, (unsigned int)i);
            atomicAdd(lo_hi + 2, 1u);
            if (i >= 0x0D800000u && i < 0x7F800000u) atomicAdd(lo_hi + 3, 1u); // mismatches among q >= 2^-100, finite
        }
    }
}

__global__ void k_selftest(unsigned long long* bad, unsigned int* first_bad)
{
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < (1ull << 32); i += stride) {
        const float x = __uint_as_float((unsigned int)i);
        bool ok = true;
        if ((unsigned int)i == 0u || (x >= 7.888609052210118e-31f && x < CUDART_INF_F)) { // +0 or finite >= 2^-100
            ok = __float_as_uint(neg_div10_fast(x)) == __float_as_uint(neg_div10_ref(x));
            float w0, w1;
            unpack2(div10_fast2(pack2(x, x)), w0, w1);
            ok = ok && __float_as_uint(-w0) == __float_as_uint(neg_div10_ref(x)) && __float_as_uint(-w1) == __float_as_uint(neg_div10_ref(x));
        }
        if (x <= 0.0f || __float_as_uint(x) == 0x80000000u) { // a rescaled GL is <= 0 (or -0)
            ok = ok && (pl_from_gl(x) == pl_from_gl_ref(x)) && (pl_from_gl_magic(x) == pl_from_gl_ref(x)) &&
                 (pl_from_gl_magic_uncapped(x) == pl_from_gl_ref(x));
            float u0, u1;
            unpack2(pl_magic2(pack2(x, x)), u0, u1);
            ok = ok && pl_from_magic_bits(u0) == pl_from_gl_ref(x) && pl_from_magic_bits(u1) == pl_from_gl_ref(x);
        }
        if (!ok) {
            atomicAdd(bad, 1ull);
            atomicMin(first_bad, (unsigned int)i);
        }
    }
}

int run_selftest(unsigned long long* n_bad, unsigned int* first_bad)
{
    unsigned long long* d_bad = nullptr;
    unsigned int* d_first = nullptr;
    if (cudaMalloc((void**)&d_bad, 8) != cudaSuccess || cudaMalloc((void**)&d_first, 4) != cudaSuccess) return -1;
    cudaMemset(d_bad, 0, 8);
    cudaMemset(d_first, 0xFF, 4);
    k_selftest<<<148 * 8, 256>>>(d_bad, d_first);
    cudaError_t e = cudaDeviceSynchronize();
    if (getenv("VGL_SELFTEST_VERBOSE")) {
        unsigned int* d_r = nullptr;
        unsigned int h[4] = {0xFFFFFFFFu, 0, 0, 0};
        cudaMalloc((void**)&d_r, 16);
        cudaMemcpy(d_r, h, 16, cudaMemcpyHostToDevice);
        k_div10_range<<<148 * 8, 256>>>(d_r);
        cudaMemcpy(h, d_r, 16, cudaMemcpyDeviceToHost);
        cudaFree(d_r);
        fprintf(stderr, "[vgl selftest] unguarded div10: %u mismatches, bits 0x%08x..0x%08x, %u with q in [2^-100, inf)\n", h[2], h[0], h[1], h[3]);
    }
    cudaMemcpy(n_bad, d_bad, 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(first_bad, d_first, 4, cudaMemcpyDeviceToHost);
    cudaFree(d_bad);
    cudaFree(d_first);
    return e == cudaSuccess ? 0 : -1;
}

void launch_fused_m1f(const DevParams& p, cudaStream_t st, int n_sms)
{
    int per_sm = 1;
    const size_t dyn = (size_t)(((p.pois_n + 15) & ~15) + FUSED_TILE_CELLS * (p.sample_strand ? 2 : 1)) * 8 + 512;
    cudaFuncSetAttribute(k_fused_m1f, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((FUSED_POIS_MAX + 2 * FUSED_TILE_CELLS) * 8 + 512));
    cudaFuncSetAttribute(k_fused_m1f, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_fused_m1f, FUSED_BLOCK, dyn);
    if (per_sm < 1) per_sm = 1;
    int grid = n_sms * per_sm;
    if (grid > p.n_tiles) grid = p.n_tiles;
    k_fused_m1f<<<grid, FUSED_BLOCK, dyn, st>>>(p);
}

} // namespace vgl
