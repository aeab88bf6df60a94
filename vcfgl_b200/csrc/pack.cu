// k_narrow -- VGL_HOST_NARROW: the integer tag planes (PL, AD, ADF, ADR, FORMAT/DP) narrowed on the device to the
// width BCF stores them in (htslib/vcf.c:2249-2294 bcf_enc_vint narrows the int32 arrays add_tags() passes,
// bcf_utils.cpp:426-507, to int8 / int16 per record), so that 81 instead of 145 bytes per cell cross PCIe.
// One pass per plane over the used extent (the element count is read from the device totals the main kernels
// leave), 16 B loads -> 4 / 8 B stores, HBM-bound and ~1 % of the PCIe time it saves.
#include "vgl_internal.h"

namespace vgl {

// PL: 0..255 by construction, bcf_int32_missing -> 0 (the host tests FORMAT/DP == 0, vgl.h).
// COUNT planes: non-negative; a value above the range is saturated and raises VGL_EOVERFLOW.
template <typename OUT, bool IS_PL>
__global__ void __launch_bounds__(256) k_narrow(const int32_t* __restrict__ src, OUT* __restrict__ dst, const int64_t* __restrict__ n_dev,
                                                int64_t n_fixed, int32_t* status)
{
    const int64_t n = n_dev ? *n_dev : n_fixed;
    const int64_t quads = n >> 2;
    constexpr int32_t MAXV = sizeof(OUT) == 1 ? 255 : 65535;
    bool over = false;
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < quads; q += (int64_t)gridDim.x * blockDim.x) {
        const int4 v = __ldcs(reinterpret_cast<const int4*>(src) + q);
        int32_t x[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (IS_PL) x[k] = x[k] == VGL_I32_MISSING ? 0 : x[k];
            over = over || x[k] > MAXV || x[k] < 0;
            x[k] = min(max(x[k], 0), MAXV);
        }
        if (sizeof(OUT) == 1) {
            reinterpret_cast<uint32_t*>(dst)[q] = (uint32_t)x[0] | ((uint32_t)x[1] << 8) | ((uint32_t)x[2] << 16) | ((uint32_t)x[3] << 24);
        } else {
            reinterpret_cast<uint2*>(dst)[q] = make_uint2((uint32_t)x[0] | ((uint32_t)x[1] << 16), (uint32_t)x[2] | ((uint32_t)x[3] << 16));
        }
    }
    if (blockIdx.x == 0 && threadIdx.x < (int)(n & 3)) { // tail of a plane whose extent is not a multiple of four (FORMAT/DP)
        int32_t x = src[(quads << 2) + threadIdx.x];
        if (IS_PL) x = x == VGL_I32_MISSING ? 0 : x;
        over = over || x > MAXV || x < 0;
        dst[(quads << 2) + threadIdx.x] = (OUT)min(max(x, 0), MAXV);
    }
    if (over) atomicExch(status, (int)VGL_EOVERFLOW);
}

// bits: 8 or 16 (PL is always 8).  n_dev: device word holding the element count, or null -> n_fixed.
void launch_narrow(const int32_t* src, void* dst, int bits, bool is_pl, const int64_t* n_dev, int64_t n_fixed, int64_t cap, int32_t* status,
                   cudaStream_t st, int n_sms)
{
    int64_t blocks = (cap / 4 + 255) / 256;
    if (blocks > (int64_t)n_sms * 16) blocks = (int64_t)n_sms * 16;
    if (blocks < 1) blocks = 1;
    if (is_pl) k_narrow<uint8_t, true><<<(unsigned)blocks, 256, 0, st>>>(src, (uint8_t*)dst, n_dev, n_fixed, status);
    else if (bits == 8) k_narrow<uint8_t, false><<<(unsigned)blocks, 256, 0, st>>>(src, (uint8_t*)dst, n_dev, n_fixed, status);
    else k_narrow<uint16_t, false><<<(unsigned)blocks, 256, 0, st>>>(src, (uint16_t*)dst, n_dev, n_fixed, status);
}

} // namespace vgl
