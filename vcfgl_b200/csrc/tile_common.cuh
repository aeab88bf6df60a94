// Shared pieces of the tile kernels (tile_m1f.cu, tile_m2.cu): tile geometry, the per-site record built in
// phase B, shared-memory / bulk-copy helpers, Philox with round keys in the parameter bank, chunk tickets.
#pragma once
#include "counts_sampler.cuh"
#include "m1f.cuh"

namespace vgl {

#ifndef TILE_BLOCK
#define TILE_BLOCK 128
#endif
#define TILE_WARPS (TILE_BLOCK / 32)
#ifndef TILE_MAX_SITES
#define TILE_MAX_SITES 32    // <= 32: phase B is one warp
#endif
#ifndef TILE_CELLS
#define TILE_CELLS 2048      // virtual cells of a tile when a site is smaller than this
#endif
#ifndef TILE_MIN_CTAS
#define TILE_MIN_CTAS 6
#endif
#define TILE_WST_G 528       // 4-byte elements per warp and G-shaped plane: 32 cells x 15 + pads (<= 3 per site end) <= 504, then a scratch cell
#define TILE_WST_R 192       // 32 cells x 5 + pads <= 184, then a scratch cell
#define TILE_SCRATCH_CTAS_PER_SM 8 // BIG variant: resident CTAs per SM the count scratch is sized for
#define TILE_G_TRASH 512
#define TILE_R_TRASH 184

struct __align__(16) TSite {
    uint32_t slot[4];     // byte k = 4 * (allele-space genotype slot of base pair k), 0xFF = pair not at this site
    int32_t g_rel, r_rel; // element offsets of the site's blocks relative to the tile's base
    uint32_t AG;          // A | G << 8 | all15 << 16   (A = G = 0: site skipped)
    uint32_t sel4;        // PRMT selector of allele 4 (see sel01)
    uint32_t sel01, sel23; // 16-bit PRMT selectors of alleles 0..3: byte (base) of the packed counts, 4 = reads 0
    int32_t g_end, r_end; // g_rel / r_rel + the padded block size
    uint32_t cls[4];      // per base x: bits 0..14 = genotype slots that hold x exactly once, bits 16..19 = the slot of xx
};

// per-site scratch of the AUX variant of k_tile_m1f (QS / I16 / INFO ADF, ADR)
struct __align__(16) TAux {
    int fw[4];                 // phase A: forward-strand reads per base
    unsigned long long ts, tq; // phase A: sum / sum of squares of the reads' tail distances
    int last;                  // phase A: last sample that has reads, -1: none
    int tot[4];                // phase B: reads per base (INFO/AD in ACGT order)
    uint32_t b2a;              // phase B: nibble b = allele index of base b (4 = unobserved allele), 0xF = not an allele
    uint32_t info;             // phase B: n_alleles | n_alleles_observed << 8 | (record kept and has reads) << 16
    int _pad;
};

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

// Philox4x32-10 with the round keys read from the kernel parameters (constant bank operands)
__device__ __forceinline__ u32x4 philox_rk(const DevParams& p, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3)
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const unsigned long long p0 = (unsigned long long)0xD2511F53u * c0, p1 = (unsigned long long)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ p.rk[2 * r], n2 = (uint32_t)(p0 >> 32) ^ c3 ^ p.rk[2 * r + 1];
        c1 = (uint32_t)p1;
        c3 = (uint32_t)p0;
        c0 = n0;
        c2 = n2;
    }
    u32x4 o;
    o.x = c0; o.y = c1; o.z = c2; o.w = c3;
    return o;
}

// shared memory by 32-bit address (keeps generic->shared conversions out of the loops)
__device__ __forceinline__ uint32_t lds32(uint32_t a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ uint2 lds64(uint32_t a) { uint2 v; asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a)); return v; }
__device__ __forceinline__ uint4 lds128(uint32_t a)
{
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts32(uint32_t a, uint32_t v) { asm volatile("st.shared.b32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }

// next chunk of 32 virtual cells of the running phase: one shared-memory ticket per warp and chunk, so that
// warps the scheduler favours take more chunks and the phase ends for all warps at about the same time.
// Issue (lane 0's atomic) and use (broadcast) are split so that the atomic's latency overlaps a whole chunk.
__device__ __forceinline__ int tile_ticket_issue(uint32_t s_ctr, int lane)
{
    // one predicated ATOMS by lane 0.  The caller adds a zero it loaded from shared memory to the address: with a
    // provably warp-uniform address ptxas rewrites the atomic into its leader-election / aggregate / broadcast
    // sequence, whose internal shuffle waits for the atomic right away.
    int c = 0;
    asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.u32 p, %2, 0;\n\t@p atom.shared.add.u32 %0, [%1], 1;\n\t}" : "+r"(c) : "r"(s_ctr), "r"(lane) : "memory");
    return c;
}
__device__ __forceinline__ int tile_ticket_get(int raw) { return __shfl_sync(0xffffffffu, raw, 0); }

// Phase B of a tile (one warp, a lane per site): per-site record from the base totals `tot` (vcfgl.cpp:396-404,
// 665-782, 806-843 INFO part), the per-site scatter tables `st`, block offsets within the tile and the tile bases.
__device__ __forceinline__ void tile_phase_b(const DevParams& p, const int lane, const int nsl, const int site0, const int tile, const int S,
                                             const int T, int* tot, TSite* st, const bool explode, const bool add_unobs, int64_t* s_base,
                                             uint32_t* s_ctr, TAux* aux = nullptr)
{
    int my_g = 0, my_r = 0; // this site's block sizes in 4-byte elements (padded to 16 B)
    if (lane < nsl) {
        int t[4];
#pragma unroll
        for (int b = 0; b < 4; ++b) { t[b] = tot[lane * 4 + b]; tot[lane * 4 + b] = 0; }
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
                    if (aux && a < n_alleles && b < 4) { // vcfgl.cpp:833-841
                        const int f = aux[lane].fw[b];
                        if (p.tag_mask & VGL_TAG_INFO_ADF) o.info_adf[a] = f;
                        if (p.tag_mask & VGL_TAG_INFO_ADR) o.info_adr[a] = t[b] - f;
                    }
                }
            }
        }
        // dp == 0 sites keep all-missing blocks: their "alleles" carry no base (counts read as 0)
        const bool keep = o.skip_code == 0 && o.n_alleles > 0;
        if (aux) {
            uint32_t m = 0;
#pragma unroll
            for (int b = 0; b < 5; ++b) m |= (uint32_t)(b2a[b] & 0xF) << (4 * b);
#pragma unroll
            for (int b = 0; b < 4; ++b) aux[lane].tot[b] = t[b];
            aux[lane].b2a = m;
            aux[lane].info = (uint32_t)o.n_alleles | ((uint32_t)o.n_alleles_observed << 8) | ((keep && dp > 0) ? 1u << 16 : 0u);
        }
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
        // AD permute selectors: allele a reads byte (base) of the packed counts; alleles without a base read byte 4 = 0
        uint32_t sel[5];
#pragma unroll
        for (int a = 0; a < 5; ++a) {
            const uint32_t b = (a2b >> (4 * a)) & 0xF;
            sel[a] = (b < 4 ? b : 4u) | 0x4440u;
        }
        ts.sel01 = sel[0] | (sel[1] << 16);
        ts.sel23 = sel[2] | (sel[3] << 16);
        ts.sel4 = sel[4];
        ts.AG = keep ? ((uint32_t)o.n_alleles | ((uint32_t)o.n_genotypes << 8) | ((all15 && dp > 0) ? 1u << 16 : 0u)) : 0u;
#pragma unroll
        for (int x = 0; x < 4; ++x) { // slot classes of a cell whose reads all show base x (pure cells: closed-form GL / PL)
            uint32_t m = 0u;
#pragma unroll
            for (int k = 0; k < 5; ++k)
#pragma unroll
                for (int j = 0; j <= k; ++j) {
                    const uint32_t sl_ = (uint32_t)((pm >> (4 * (k * (k + 1) / 2 + j))) & 0xF);
                    if (sl_ != 0xFu) {
                        if ((j == x) != (k == x)) m |= 1u << sl_;
                        if (j == x && k == x) m |= sl_ << 16;
                    }
                }
            ts.cls[x] = m;
        }
        if (keep) {
            my_g = (S * o.n_genotypes + 3) & ~3;
            my_r = (S * o.n_alleles + 3) & ~3;
        }
        // inclusive scan of the block sizes over the tile's sites (<= 32: this warp)
        ts.g_rel = ts.r_rel = ts.g_end = ts.r_end = 0;
        st[lane] = ts;
        p.sites[site0 + lane] = o;
    }
    int ig = my_g, ir = my_r;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const int tg = __shfl_up_sync(0xffffffffu, ig, off);
        const int tr = __shfl_up_sync(0xffffffffu, ir, off);
        if (lane >= off) { ig += tg; ir += tr; }
    }
    const int tile_g = __shfl_sync(0xffffffffu, ig, 31), tile_r = __shfl_sync(0xffffffffu, ir, 31);
    if (lane < nsl) {
        st[lane].g_rel = ig - my_g;
        st[lane].r_rel = ir - my_r;
        st[lane].g_end = ig;
        st[lane].r_end = ir;
    }
    // Tile bases are fixed: tile t starts at t * T * (padded size of a 15-genotype / 5-allele block).  Blocks are
    // compact within a tile, so every chunk's span is contiguous; the only holes are at tile ends, behind sites
    // with fewer alleles or skipped ones.  No tile waits for another one (a running prefix over all earlier
    // tiles -- decoupled look-back -- left every CTA idle for ~40% of its time on this workload).
    const int64_t bg = (int64_t)tile * T * ((S * 15 + 3) & ~3), br = (int64_t)tile * T * ((S * 5 + 3) & ~3);
    if (lane == 0) {
        s_base[0] = bg;
        s_base[1] = br;
        s_base[2] = tile_g; // used elements of the tile's spans (the rest up to nsl * padded block size is a hole)
        s_base[3] = tile_r;
        s_ctr[1] = 0u;
        if (tile == p.n_tiles - 1) {
            p.totals[0] = p.totals_host[0] = bg + tile_g;
            p.totals[1] = p.totals_host[1] = br + tile_r;
        }
    }
    if (lane < nsl) {
        p.sites[site0 + lane].g_off = bg + (ig - my_g);
        p.sites[site0 + lane].r_off = br + (ir - my_r);
    }
}

// VGL_HOST_I32 / VGL_HOST_NARROW copy the planes to the host as whole spans: the hole at the end of a tile (behind sites with
// fewer than five alleles or skipped ones) is zeroed so that the spans are deterministic and the narrowing pass never sees
// stale words.  Spans and holes are multiples of four elements (blocks are padded to 16 B).
__device__ __forceinline__ void tile_zero_holes(const DevParams& p, const int tid, const int nsl, const int S, const int64_t* s_base)
{
    const int g_full = nsl * ((S * 15 + 3) & ~3), r_full = nsl * ((S * 5 + 3) & ~3);
    const int g_used = (int)s_base[2], r_used = (int)s_base[3];
    const int4 z = make_int4(0, 0, 0, 0);
    for (int i = g_used + tid * 4; i < g_full; i += TILE_BLOCK * 4) {
        if (p.gl) *reinterpret_cast<int4*>(p.gl + s_base[0] + i) = z;
        if (p.pl) *reinterpret_cast<int4*>(p.pl + s_base[0] + i) = z;
        if (p.gp) *reinterpret_cast<int4*>(p.gp + s_base[0] + i) = z;
    }
    for (int i = r_used + tid * 4; i < r_full; i += TILE_BLOCK * 4) {
        if (p.ad) *reinterpret_cast<int4*>(p.ad + s_base[1] + i) = z;
        if (p.adf) *reinterpret_cast<int4*>(p.adf + s_base[1] + i) = z;
        if (p.adr) *reinterpret_cast<int4*>(p.adr + s_base[1] + i) = z;
    }
}

} // namespace vgl
