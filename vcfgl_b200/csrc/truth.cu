// --depth inf ("truth" mode): no reads are simulated, every sample's true genotype gets the best score
// and every other genotype the worst -- simulate_record_true_values(), vcfgl.cpp:1089-1262.
//
//   k_truth_site  warp per site: counts of A, C, G, T among the true haplotypes, alleles in descending count order (stable
//                 insertion sort, vcfgl.cpp:1106-1117), unobserved bases / <*> appended as -doUnobserved asks (:1129-1166)
//   k_scan        (kernels.cu) compact plane offsets
//   k_truth_emit  warp per site, a lane per plane element: GL 0 / -inf, PL 0 / 255, GP 1 / 0 (shared.h:205-212) at
//                 bcf_alleles2gt(a, b) of the sample's true alleles (vcfgl.cpp:1207-1234)
//
// Only GL, GP and PL exist in this mode (the reference refuses -addFormatDP 1 with --depth inf, io.cpp:796-800).  A missing
// true genotype is an error, as in the reference (ASSERT at vcfgl.cpp:1196): the batch status becomes VGL_EMISSING.
#include "vgl_internal.h"

namespace vgl {

namespace {

__global__ void __launch_bounds__(256) k_truth_site(const __grid_constant__ DevParams p)
{
    const int lane = threadIdx.x & 31;
    const int warp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), n_warps = gridDim.x * (blockDim.x >> 5);
    const bool explode = p.do_unobserved >= 3, unobs = p.do_unobserved == 1 || p.do_unobserved == 2 || p.do_unobserved >= 4;
    for (int i = warp; i < p.n_sites; i += n_warps) {
        const uint8_t* row = p.gt + (size_t)i * p.S;
        uint32_t packed = 0; // four 8-bit counters, flushed before they can overflow
        int ac[4] = {0, 0, 0, 0};
        bool missing = false;
        int since = 0;
        for (int s = lane; s < p.S; s += 32) {
            const uint32_t g = row[s], h0 = g & 0xF, h1 = g >> 4;
            missing = missing || h0 > 3 || h1 > 3;
            packed += (1u << (8 * (h0 & 3))) + (1u << (8 * (h1 & 3)));
            if (++since == 100) {
#pragma unroll
                for (int b = 0; b < 4; ++b) ac[b] += (packed >> (8 * b)) & 0xFF;
                packed = 0, since = 0;
            }
        }
#pragma unroll
        for (int b = 0; b < 4; ++b) ac[b] = __reduce_add_sync(0xffffffffu, ac[b] + (int)((packed >> (8 * b)) & 0xFF));
        if (__any_sync(0xffffffffu, missing) && lane == 0) atomicCAS(p.status, 0, (int)VGL_EMISSING);
        if (lane != 0) continue;
        int order[4] = {0, 1, 2, 3}, n_obs = 0;
        for (int k = 0; k < 4; ++k) {
            if (ac[k] > 0) ++n_obs;
            for (int j = k; j > 0 && ac[order[j]] > ac[order[j - 1]]; --j) {
                const int t = order[j];
                order[j] = order[j - 1];
                order[j - 1] = t;
            }
        }
        vgl_site_out o;
        memset(&o, 0, sizeof o);
        const int n_real = explode ? 4 : n_obs;
        for (int k = 0; k < 8; ++k) o.alleles2acgt[k] = o.acgt2alleles[k] = -1;
        for (int k = 0; k < n_real; ++k) {
            o.alleles2acgt[k] = (int8_t)order[k];
            o.acgt2alleles[order[k]] = (int8_t)k;
        }
        if (unobs) {
            o.alleles2acgt[n_real] = 4;
            o.acgt2alleles[4] = (int8_t)n_real;
        }
        o.skip_code = 0;
        o.n_alleles = n_real + (unobs ? 1 : 0);
        o.n_alleles_observed = n_obs;
        o.n_genotypes = o.n_alleles * (o.n_alleles + 1) / 2;
        o.info_dp = -1; // not applicable: there are no reads (and this is NOT the no-reads record of vcfgl.cpp:228-315)
        o.g_off = (int64_t)p.S * o.n_genotypes; // sizes; k_scan turns them into offsets
        o.r_off = 0;
        p.sites[i] = o;
    }
}

__global__ void __launch_bounds__(256) k_truth_emit(const __grid_constant__ DevParams p)
{
    const int lane = threadIdx.x & 31;
    const int warp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), n_warps = gridDim.x * (blockDim.x >> 5);
    for (int i = warp; i < p.n_sites; i += n_warps) {
        const vgl_site_out& so = p.sites[i];
        const int G = so.n_genotypes, n = p.S * G;
        const uint8_t* row = p.gt + (size_t)i * p.S;
        const int64_t off = so.g_off;
        // acgt -> allele index, four nibbles
        uint32_t map = 0;
#pragma unroll
        for (int b = 0; b < 4; ++b) map |= (uint32_t)(so.acgt2alleles[b] & 0xF) << (4 * b);
        for (int e = lane; e < n; e += 32) {
            const int s = e / G, g = e - s * G;
            const uint32_t gt = row[s];
            const int a = (map >> (4 * (gt & 3))) & 0xF, b = (map >> (4 * ((gt >> 4) & 3))) & 0xF;
            const int hi = max(a, b), lo = min(a, b);
            const bool best = g == hi * (hi + 1) / 2 + lo; // bcf_alleles2gt, htslib/vcf.h:902
            if (p.gl) p.gl[off + e] = best ? 0.0f : __int_as_float(0xFF800000);
            if (p.pl) p.pl[off + e] = best ? 0 : 255;
            if (p.gp) p.gp[off + e] = best ? 1.0f : 0.0f;
        }
    }
}

} // namespace

void launch_truth_site(const DevParams& p, cudaStream_t st, int n_sms) { k_truth_site<<<n_sms * 8, 256, 0, st>>>(p); }
void launch_truth_emit(const DevParams& p, cudaStream_t st, int n_sms) { k_truth_emit<<<n_sms * 8, 256, 0, st>>>(p); }

} // namespace vgl
