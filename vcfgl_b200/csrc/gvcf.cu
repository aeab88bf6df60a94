// gVCF block merger on the device (include/vgl.h vgl_gvcf_merge).
//
// The reference merges records one at a time in an order-dependent state machine, prepare_gvcf_block()
// (bcf_utils.cpp:662-942).  Its rules are local: whether a written site is a block MEMBER depends on the site alone
// (one observed allele, dp range >= 1: bcf_utils.cpp:692, 741-765), whether it JOINS the block before it depends on the
// previous written site alone (same contig, contiguous position, same dp range: bcf_utils.cpp:711, 719, 790), and a block's
// values are minima over its members (MIN_DP :838-842, DP[s] :844-848, lexicographic (PL[3s+1], PL[3s+2]) :858-866).
// So the machine is a segmented min-reduction over sites:
//
//   k_gvcf_key     eight lanes per site: min FORMAT/DP over the samples -> dp range, member flag
//   k_gvcf_plan_local / _global   blocks of 1024 sites: head flags from neighbouring written sites, record ids / block
//                  ordinals by a two-level scan
//   k_gvcf_fin     thread per record: last member, number of members
//   k_gvcf_reduce  warp per (block, 128 samples): per-sample minima over the members, MIN_DP
//
// Bound: HBM -- k_gvcf_reduce reads 16 bytes (DP + 3 PL) per member cell, k_gvcf_key 4.
#include "vgl_internal.h"

namespace vgl {

namespace {

enum { K_KEPT = 1, K_MEMBER = 2 };
constexpr int PLAN_THREADS = 1024;

// eight lanes per site, four sites per warp at a time: a warp that walks its sites one after the other waits two dependent
// round trips (site record, then DP row) per site, which is what bounded the first version (34 us for 131072 sites)
__global__ void __launch_bounds__(256) k_gvcf_key(const vgl_site_out* __restrict__ sites, const int32_t* __restrict__ dp, int32_t S, int32_t n_sites,
                                                  GvcfDps dps, int2* __restrict__ key)
{
    const int l8 = threadIdx.x & 7;
    const int grp = (blockIdx.x * blockDim.x + threadIdx.x) >> 3, n_grp = (gridDim.x * blockDim.x) >> 3;
    const uint32_t gmask = 0xFFu << (threadIdx.x & 24); // this group's lanes within the warp
    for (int i = grp; i < n_sites; i += n_grp) {
        const vgl_site_out& so = sites[i];
        if (so.skip_code != 0) { // not written at all (vcfgl.cpp:1553-1558): invisible to the merger
            if (l8 == 0) key[i] = make_int2(0, 0);
            continue;
        }
        int m = 0x7FFFFFFF;
        const int32_t* row = dp + (size_t)i * S;
        if ((S & 3) == 0) { // rows are 16-byte aligned: four samples per load
            const int4* row4 = reinterpret_cast<const int4*>(row);
            for (int s = l8; s < (S >> 2); s += 8) {
                const int4 v = __ldg(row4 + s);
                m = min(m, min(min(v.x, v.y), min(v.z, v.w)));
            }
        } else {
            for (int s = l8; s < S; s += 8) m = min(m, row[s]);
        }
        m = min(m, __shfl_xor_sync(gmask, m, 1));
        m = min(m, __shfl_xor_sync(gmask, m, 2));
        m = min(m, __shfl_xor_sync(gmask, m, 4));
        int r = 0;
#pragma unroll
        for (int k = 0; k < VGL_MAX_GVCF_DPS; ++k) r += (k < dps.n && m >= dps.v[k]) ? 1 : 0; // thresholds ascend (bcf_utils.cpp:752-757)
        const bool member = so.n_alleles_observed == 1 && r >= 1 && so.n_genotypes == 3;
        if (l8 == 0) key[i] = make_int2(m, K_KEPT | (member ? K_MEMBER : 0) | (r << 8));
    }
}

// inclusive block scans over PLAN_THREADS values (sum of a packed 64-bit word, max of an int)
__device__ __forceinline__ void block_scan(unsigned long long& sum, int& mx, unsigned long long* s_sum, int* s_max)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned long long a = __shfl_up_sync(0xffffffffu, sum, d);
        const int b = __shfl_up_sync(0xffffffffu, mx, d);
        if (lane >= d) {
            sum += a;
            mx = max(mx, b);
        }
    }
    if (lane == 31) {
        s_sum[wid] = sum;
        s_max[wid] = mx;
    }
    __syncthreads();
    unsigned long long a = 0;
    int b = -1;
    for (int w = 0; w < wid; ++w) {
        a += s_sum[w];
        b = max(b, s_max[w]);
    }
    sum += a;
    mx = max(mx, b);
    __syncthreads();
}

// last written site before `end` (exclusive), found by the whole block, PLAN_THREADS sites per step (skipped sites are rare:
// one step; bounded for the all-skipped worst case).  Every thread of the block must call it.
__device__ __forceinline__ int block_last_kept_before(const int2* __restrict__ key, int end, int* s_red)
{
    for (int hi = end; hi > 0; hi -= PLAN_THREADS) {
        const int i = hi - 1 - (int)threadIdx.x;
        int v = (i >= 0 && (key[i].y & K_KEPT)) ? i : -1;
        v = __reduce_max_sync(0xffffffffu, v);
        if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
        __syncthreads();
        int m = -1;
        for (int w = 0; w < PLAN_THREADS / 32; ++w) m = max(m, s_red[w]);
        __syncthreads();
        if (m >= 0) return m;
    }
    return -1;
}

// The plan in two parallel steps over blocks of PLAN_THREADS sites.
// packed sums: bits 0..20 records (heads), 21..41 blocks (member heads), 42..62 written sites; bit 63 of local[]: head flag
__global__ void __launch_bounds__(PLAN_THREADS) k_gvcf_plan_local(const int2* __restrict__ key, const vgl_gvcf_site_in* __restrict__ sin, int32_t n_sites,
                                                                  int32_t* __restrict__ prev_kept,
                                                                  unsigned long long* __restrict__ local, unsigned long long* __restrict__ block_sum)
{
    __shared__ unsigned long long s_sum[PLAN_THREADS / 32];
    __shared__ int s_max[PLAN_THREADS / 32];
    __shared__ int s_incl[PLAN_THREADS];
    const int s_before = block_last_kept_before(key, (int)blockIdx.x * PLAN_THREADS, s_max); // last written site before this block
    const int i = blockIdx.x * PLAN_THREADS + threadIdx.x;
    const bool in = i < n_sites;
    const int2 k = in ? key[i] : make_int2(0, 0);
    const bool kept = (k.y & K_KEPT) != 0, member = (k.y & K_MEMBER) != 0;
    unsigned long long dummy = 0;
    int mx = kept ? i : -1;
    block_scan(dummy, mx, s_sum, s_max);
    s_incl[threadIdx.x] = max(mx, s_before);
    __syncthreads();
    const int p = threadIdx.x ? s_incl[threadIdx.x - 1] : s_before; // last written site before i
    bool head = false;
    if (kept) {
        head = true;
        if (member && p >= 0) {
            const int2 kp = key[p];
            const vgl_gvcf_site_in a = sin[p], b = sin[i];
            if ((kp.y & K_MEMBER) && (kp.y >> 8) == (k.y >> 8) && a.rid == b.rid && b.pos <= a.pos + 1) head = false;
        }
    }
    unsigned long long sum = (head ? 1ull : 0ull) | ((head && member) ? 1ull << 21 : 0ull) | (kept ? 1ull << 42 : 0ull);
    int dummy_max = -1;
    block_scan(sum, dummy_max, s_sum, s_max);
    if (in) {
        prev_kept[i] = p;
        local[i] = sum | (head ? 1ull << 63 : 0ull);
    }
    if (threadIdx.x == PLAN_THREADS - 1) block_sum[blockIdx.x] = sum;
}

__global__ void __launch_bounds__(PLAN_THREADS) k_gvcf_plan_global(const int2* __restrict__ key, int32_t n_sites, int32_t* __restrict__ blk_rec,
                                                                   const unsigned long long* __restrict__ local,
                                                                   const unsigned long long* __restrict__ block_sum, vgl_gvcf_rec* __restrict__ recs,
                                                                   int32_t* __restrict__ kept_idx, int32_t* __restrict__ counts)
{
    __shared__ unsigned long long s_part[PLAN_THREADS / 32];
    __shared__ unsigned long long s_off;
    __shared__ int s_red[PLAN_THREADS / 32];
    int last_written = -1;
    if (blockIdx.x == gridDim.x - 1) last_written = block_last_kept_before(key, n_sites, s_red); // the whole block takes part
    unsigned long long part = 0;
    for (int b = threadIdx.x; b < (int)blockIdx.x; b += PLAN_THREADS) part += block_sum[b];
#pragma unroll
    for (int d = 16; d; d >>= 1) part += __shfl_xor_sync(0xffffffffu, part, d);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = part;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long t = 0;
        for (int w = 0; w < PLAN_THREADS / 32; ++w) t += s_part[w];
        s_off = t;
    }
    __syncthreads();
    const int i = blockIdx.x * PLAN_THREADS + threadIdx.x;
    if (i < n_sites) {
        const unsigned long long l = local[i];
        const unsigned long long sum = (l & ~(1ull << 63)) + s_off;
        const int2 k = key[i];
        const bool kept = (k.y & K_KEPT) != 0, member = (k.y & K_MEMBER) != 0, head = (l >> 63) != 0;
        kept_idx[i] = (int)((sum >> 42) & 0x1FFFFF) - (kept ? 1 : 0);
        if (head) {
            vgl_gvcf_rec r;
            r.first_site = i;
            r.last_site = i;
            r.n_members = member ? 1 : 0;
            r.min_dp = k.x;
            r.dp_range = member ? (k.y >> 8) : 0;
            r.plane = member ? (int)((sum >> 21) & 0x1FFFFF) - 1 : -1;
            recs[(int)(sum & 0x1FFFFF) - 1] = r;
            if (member) blk_rec[r.plane] = (int)(sum & 0x1FFFFF) - 1;
        }
        if (i == n_sites - 1) {
            counts[0] = (int)(sum & 0x1FFFFF);         // records
            counts[1] = (int)((sum >> 21) & 0x1FFFFF); // blocks
            counts[2] = last_written;                  // last written site (-1: none)
        }
    }
}

__global__ void k_gvcf_fin(vgl_gvcf_rec* __restrict__ recs, const int32_t* __restrict__ prev_kept, const int32_t* __restrict__ kept_idx,
                           const int32_t* __restrict__ counts)
{
    const int n = counts[0];
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x) {
        if (recs[r].n_members == 0) continue;
        const int first = recs[r].first_site;
        const int last = r + 1 < n ? prev_kept[recs[r + 1].first_site] : counts[2];
        recs[r].last_site = last;
        recs[r].n_members = kept_idx[last] - kept_idx[first] + 1;
    }
}

// warp per (block, 128 samples): four independent 32-sample slices per lane keep four times the loads in flight per dependent
// step (record -> member key -> plane offset -> values), which is what bounds this kernel
constexpr int RED_U = 4;

__global__ void __launch_bounds__(256) k_gvcf_reduce(vgl_gvcf_rec* recs, const int32_t* __restrict__ counts, const int32_t* __restrict__ blk_rec, const int2* __restrict__ key,
                                                     const vgl_site_out* __restrict__ sites, const int32_t* __restrict__ dp,
                                                     const int32_t* __restrict__ pl, int32_t S, int32_t* __restrict__ out_dp,
                                                     int32_t* __restrict__ out_pl)
{
    const int lane = threadIdx.x & 31;
    const int n_blocks = counts[1];
    const int chunks = (S + 32 * RED_U - 1) / (32 * RED_U);
    const unsigned total = (unsigned)n_blocks * (unsigned)chunks;
    const unsigned warp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), n_warps = gridDim.x * (blockDim.x >> 5);
    for (unsigned w = warp; w < total; w += n_warps) {
        const int blk = (int)(w / (unsigned)chunks), c = (int)(w - (unsigned)blk * (unsigned)chunks);
        const int r = blk_rec[blk];
        const vgl_gvcf_rec rec = recs[r];
        int d[RED_U], p0[RED_U], p1[RED_U], p2[RED_U], md = 0x7FFFFFFF;
#pragma unroll
        for (int u = 0; u < RED_U; ++u) d[u] = p1[u] = p2[u] = 0x7FFFFFFF, p0[u] = 0;
        const int sbase = c * 32 * RED_U + lane;
        for (int i = rec.first_site; i <= rec.last_site; ++i) {
            const int2 k = key[i];
            if (!(k.y & K_KEPT)) continue;
            md = min(md, k.x);
            const int32_t* drow = dp + (size_t)i * S;
            const int32_t* prow = pl ? pl + sites[i].g_off : nullptr;
            int dv[RED_U], a[RED_U], b[RED_U], z[RED_U];
#pragma unroll
            for (int u = 0; u < RED_U; ++u) { // all loads first
                const int s = sbase + 32 * u;
                const bool in = s < S;
                dv[u] = in ? drow[s] : 0x7FFFFFFF;
                z[u] = in && prow ? prow[3 * (size_t)s] : 0;
                a[u] = in && prow ? prow[3 * (size_t)s + 1] : 0x7FFFFFFF;
                b[u] = in && prow ? prow[3 * (size_t)s + 2] : 0x7FFFFFFF;
            }
#pragma unroll
            for (int u = 0; u < RED_U; ++u) {
                d[u] = min(d[u], dv[u]);
                if (i == rec.first_site) p0[u] = z[u];
                if (a[u] < p1[u] || (a[u] == p1[u] && b[u] < p2[u])) { // bcf_utils.cpp:858-866 = lexicographic minimum
                    p1[u] = a[u];
                    p2[u] = b[u];
                }
            }
        }
#pragma unroll
        for (int u = 0; u < RED_U; ++u) {
            const int s = sbase + 32 * u;
            if (s < S) {
                out_dp[(size_t)rec.plane * S + s] = d[u];
                if (pl) {
                    int32_t* o = out_pl + ((size_t)rec.plane * S + s) * 3;
                    o[0] = p0[u], o[1] = p1[u], o[2] = p2[u];
                }
            }
        }
        if (c == 0 && lane == 0) recs[r].min_dp = md;
    }
}

__global__ void k_gvcf_sin_from_bcf(const vgl_bcf_site_in* __restrict__ in, vgl_gvcf_site_in* __restrict__ out, int32_t n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        vgl_gvcf_site_in o;
        o.rid = in[i].rid;
        o.pos = in[i].pos;
        out[i] = o;
    }
}

} // namespace

// VGL_HOST_BCF with -doGVCF: the merger's (contig, position) input from the pass-through records already on the device
void launch_gvcf_sin_from_bcf(const vgl_bcf_site_in* in, vgl_gvcf_site_in* out, int32_t n, cudaStream_t st)
{
    k_gvcf_sin_from_bcf<<<(n + 255) / 256, 256, 0, st>>>(in, out, n);
}

void launch_gvcf(const GvcfArgs& a, cudaStream_t st, int n_sms)
{
    const int nb = (a.n_sites + PLAN_THREADS - 1) / PLAN_THREADS;
    k_gvcf_key<<<n_sms * 8, 256, 0, st>>>(a.sites, a.dp, a.S, a.n_sites, a.dps, a.key);
    k_gvcf_plan_local<<<nb, PLAN_THREADS, 0, st>>>(a.key, a.sin, a.n_sites, a.prev_kept, a.local, a.block_sum);
    k_gvcf_plan_global<<<nb, PLAN_THREADS, 0, st>>>(a.key, a.n_sites, a.blk_rec, a.local, a.block_sum, a.recs, a.kept_idx, a.counts);
    k_gvcf_fin<<<n_sms * 2, 256, 0, st>>>(a.recs, a.prev_kept, a.kept_idx, a.counts);
    k_gvcf_reduce<<<n_sms * 8, 256, 0, st>>>(a.recs, a.counts, a.blk_rec, a.key, a.sites, a.dp, a.pl, a.S, a.out_dp, a.out_pl);
}

} // namespace vgl
