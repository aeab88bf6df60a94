// On-device genotype-call discordance summary (include/vgl.h vgl_discordance).
//
// The comparison misc/gtDiscordance.cpp makes between a call set and the truth (hom / het strata, misc/gtDiscordance.cpp:11-15),
// applied to the simulator's own likelihoods: a cell's call is the genotype with the single largest GL (a tie is no call and
// counts as discordant); it is compared, as an unordered pair of bases, with the true genotype.  Cells of written sites
// (skip_code 0) with INFO/DP > 0 and FORMAT/DP > 0 take part -- the definition the statistical parity tests use
// (tools/make_stats_golden.py, tests/test_gpu_native.py).
//
//   k_discordance  warp per site, a lane per sample: one pass over the site's GL block (60 B per cell at 15 genotypes);
//                  four counters per warp through REDUX, one atomic per counter and warp.  Bound: HBM (reads GL once).
#include "vgl_internal.h"

namespace vgl {

namespace {

__global__ void __launch_bounds__(256) k_discordance(const vgl_site_out* __restrict__ sites, const uint8_t* __restrict__ gt,
                                                     const int32_t* __restrict__ dp, const float* __restrict__ gl, int32_t S, int32_t n_sites,
                                                     unsigned long long* __restrict__ counts)
{
    const int lane = threadIdx.x & 31;
    const int warp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), n_warps = gridDim.x * (blockDim.x >> 5);
    uint32_t n_hom = 0, d_hom = 0, n_het = 0, d_het = 0;
    for (int i = warp; i < n_sites; i += n_warps) {
        const vgl_site_out& so = sites[i];
        if (so.skip_code != 0 || so.info_dp == 0) continue;
        const int G = so.n_genotypes;
        const float* blk = gl + so.g_off;
        uint32_t a2b = 0; // allele -> base, nibbles
#pragma unroll
        for (int a = 0; a < 5; ++a) a2b |= (uint32_t)(so.alleles2acgt[a] & 0xF) << (4 * a);
        for (int s = lane; s < S; s += 32) {
            if (dp[(size_t)i * S + s] <= 0) continue;
            const float* row = blk + (size_t)s * G;
            float mx = row[0];
            int best = 0, ties = 1;
            for (int g = 1; g < G; ++g) {
                const float v = row[g];
                if (v > mx) mx = v, best = g, ties = 1;
                else if (v == mx) ++ties;
            }
            // genotype index -> allele pair (a1 <= a2): g = a2 (a2 + 1) / 2 + a1 (htslib/vcf.h:902)
            int a2 = 0;
            while ((a2 + 1) * (a2 + 2) / 2 <= best) ++a2;
            const int a1 = best - a2 * (a2 + 1) / 2;
            const uint32_t c1 = (a2b >> (4 * a1)) & 0xF, c2 = (a2b >> (4 * a2)) & 0xF;
            const uint32_t g8 = gt[(size_t)i * S + s], t1 = g8 & 0xF, t2 = g8 >> 4;
            const bool same = ties == 1 && ((c1 == t1 && c2 == t2) || (c1 == t2 && c2 == t1));
            if (t1 == t2) n_hom += 1, d_hom += !same;
            else n_het += 1, d_het += !same;
        }
    }
    n_hom = __reduce_add_sync(0xffffffffu, n_hom);
    d_hom = __reduce_add_sync(0xffffffffu, d_hom);
    n_het = __reduce_add_sync(0xffffffffu, n_het);
    d_het = __reduce_add_sync(0xffffffffu, d_het);
    if (lane == 0) {
        if (n_hom) atomicAdd(&counts[0], (unsigned long long)n_hom);
        if (d_hom) atomicAdd(&counts[1], (unsigned long long)d_hom);
        if (n_het) atomicAdd(&counts[2], (unsigned long long)n_het);
        if (d_het) atomicAdd(&counts[3], (unsigned long long)d_het);
    }
}

} // namespace

void launch_discordance(const vgl_site_out* sites, const uint8_t* gt, const int32_t* dp, const float* gl, int32_t S, int32_t n_sites,
                        unsigned long long* counts, cudaStream_t st, int n_sms)
{
    k_discordance<<<n_sms * 8, 256, 0, st>>>(sites, gt, dp, gl, S, n_sites, counts);
}

} // namespace vgl
