// gVCF block merger on the device (include/vgl.h vgl_gvcf_merge).
//
// The reference merges records one at a time in an order-dependent state machine, prepare_gvcf_block()
// (bcf_utils.cpp:662-942).  Its rules are local: whether a written site is a block MEMBER depends on the site alone
// (one observed allele, dp range >= 1: bcf_utils.cpp:692, 741-765), whether it JOINS the block before it depends on the
// previous written site alone (same contig, contiguous position, same dp range: bcf_utils.cpp:711, 719, 790), and a block's
// values are minima over its members (MIN_DP :838-842, DP[s] :844-848, lexicographic (PL[3s+1], PL[3s+2]) :858-866).
// So the machine is a segmented min-reduction over sites:
//
//   k_gvcf_key     warp per site: min FORMAT/DP over the samples -> dp range, member flag
//   k_gvcf_plan    one block: head flags from neighbouring written sites, record ids / block ordinals by scan
//   k_gvcf_fin     thread per record: last member, number of members
//   k_gvcf_reduce  warp per (block, 32 samples): per-sample minima over the members, MIN_DP
//
// Bound: HBM -- k_gvcf_reduce reads 16 bytes (DP + 3 PL) per member cell, k_gvcf_key 4.
#include "vgl_internal.h"

namespace vgl {

namespace {

enum { K_KEPT = 1, K_MEMBER = 2 };

__global__ void __launch_bounds__(256) k_gvcf_key(const vgl_site_out* __restrict__ sites, const int32_t* __restrict__ dp, int32_t S, int32_t n_sites,
                                                  GvcfDps dps, int2* __restrict__ key)
{
    const int lane = threadIdx.x & 31;
    const int warp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), n_warps = gridDim.x * (blockDim.x >> 5);
    for (int i = warp; i < n_sites; i += n_warps) {
        const vgl_site_out& so = sites[i];
        if (so.skip_code != 0) { // not written at all (vcfgl.cpp:1553-1558): invisible to the merger
            if (lane == 0) key[i] = make_int2(0, 0);
            continue;
        }
        int m = 0x7FFFFFFF;
        const int32_t* row = dp + (size_t)i * S;
        for (int s = lane; s < S; s += 32) m = min(m, row[s]);
        m = __reduce_min_sync(0xffffffffu, m);
        int r = 0;
#pragma unroll
        for (int k = 0; k < VGL_MAX_GVCF_DPS; ++k) r += (k < dps.n && m >= dps.v[k]) ? 1 : 0; // thresholds ascend (bcf_utils.cpp:752-757)
        const bool member = so.n_alleles_observed == 1 && r >= 1 && so.n_genotypes == 3;
        if (lane == 0) key[i] = make_int2(m, K_KEPT | (member ? K_MEMBER : 0) | (r << 8));
    }
}

constexpr int PLAN_THREADS = 1024;

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

// packed sums: bits 0..20 records (heads), 21..41 blocks (member heads), 42..62 written sites
__global__ void __launch_bounds__(PLAN_THREADS) k_gvcf_plan(const int2* __restrict__ key, const vgl_gvcf_site_in* __restrict__ sin, int32_t n_sites,
                                                            vgl_gvcf_rec* __restrict__ recs, int32_t* __restrict__ prev_kept,
                                                            int32_t* __restrict__ kept_idx, int32_t* __restrict__ counts)
{
    __shared__ unsigned long long s_sum[PLAN_THREADS / 32];
    __shared__ int s_max[PLAN_THREADS / 32];
    __shared__ int s_incl[PLAN_THREADS];
    __shared__ unsigned long long s_carry_sum;
    __shared__ int s_carry_max;
    if (threadIdx.x == 0) {
        s_carry_sum = 0;
        s_carry_max = -1;
    }
    __syncthreads();
    for (int base = 0; base < n_sites; base += PLAN_THREADS) {
        const int i = base + threadIdx.x;
        const bool in = i < n_sites;
        const int2 k = in ? key[i] : make_int2(0, 0);
        const bool kept = (k.y & K_KEPT) != 0, member = (k.y & K_MEMBER) != 0;
        // index of the last written site before i
        unsigned long long dummy = 0;
        int mx = kept ? i : -1;
        block_scan(dummy, mx, s_sum, s_max);
        const int incl_max = max(mx, s_carry_max);
        s_incl[threadIdx.x] = incl_max;
        __syncthreads();
        const int p = threadIdx.x ? s_incl[threadIdx.x - 1] : s_carry_max; // exclusive maximum
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
        sum += s_carry_sum;
        if (in) {
            prev_kept[i] = p;
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
            }
        }
        __syncthreads();
        if (threadIdx.x == PLAN_THREADS - 1) {
            s_carry_sum = sum;
            s_carry_max = incl_max;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        counts[0] = (int)(s_carry_sum & 0x1FFFFF);         // records
        counts[1] = (int)((s_carry_sum >> 21) & 0x1FFFFF); // blocks
        counts[2] = s_carry_max;                           // last written site (-1: none)
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

__global__ void __launch_bounds__(256) k_gvcf_reduce(vgl_gvcf_rec* recs, const int32_t* __restrict__ counts, const int2* __restrict__ key,
                                                     const vgl_site_out* __restrict__ sites, const int32_t* __restrict__ dp,
                                                     const int32_t* __restrict__ pl, int32_t S, int32_t* __restrict__ out_dp,
                                                     int32_t* __restrict__ out_pl)
{
    const int lane = threadIdx.x & 31;
    const int n = counts[0];
    const int chunks = (S + 31) / 32;
    const long long total = (long long)n * chunks;
    const long long warp = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), n_warps = (long long)gridDim.x * (blockDim.x >> 5);
    for (long long w = warp; w < total; w += n_warps) {
        const int r = (int)(w / chunks), c = (int)(w - (long long)r * chunks);
        const vgl_gvcf_rec rec = recs[r];
        if (rec.n_members == 0) continue;
        const int s = c * 32 + lane;
        int d = 0x7FFFFFFF, p0 = 0, p1 = 0x7FFFFFFF, p2 = 0x7FFFFFFF, md = 0x7FFFFFFF;
        for (int i = rec.first_site; i <= rec.last_site; ++i) {
            const int2 k = key[i];
            if (!(k.y & K_KEPT)) continue;
            md = min(md, k.x);
            if (s < S) {
                d = min(d, dp[(size_t)i * S + s]);
                if (pl) {
                    const int32_t* q = pl + sites[i].g_off + 3 * (size_t)s;
                    const int a = q[1], b = q[2];
                    if (i == rec.first_site) p0 = q[0];
                    if (a < p1 || (a == p1 && b < p2)) { // bcf_utils.cpp:858-866 = lexicographic minimum
                        p1 = a;
                        p2 = b;
                    }
                }
            }
        }
        if (s < S) {
            out_dp[(size_t)rec.plane * S + s] = d;
            if (pl) {
                int32_t* o = out_pl + ((size_t)rec.plane * S + s) * 3;
                o[0] = p0, o[1] = p1, o[2] = p2;
            }
        }
        if (c == 0 && lane == 0) recs[r].min_dp = md;
    }
}

} // namespace

void launch_gvcf(const GvcfArgs& a, cudaStream_t st, int n_sms)
{
    k_gvcf_key<<<n_sms * 8, 256, 0, st>>>(a.sites, a.dp, a.S, a.n_sites, a.dps, a.key);
    k_gvcf_plan<<<1, PLAN_THREADS, 0, st>>>(a.key, a.sin, a.n_sites, a.recs, a.prev_kept, a.kept_idx, a.counts);
    k_gvcf_fin<<<n_sms * 2, 256, 0, st>>>(a.recs, a.prev_kept, a.kept_idx, a.counts);
    k_gvcf_reduce<<<n_sms * 8, 256, 0, st>>>(a.recs, a.counts, a.key, a.sites, a.dp, a.pl, a.S, a.out_dp, a.out_pl);
}

} // namespace vgl
