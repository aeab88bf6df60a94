// Device code shared by the unfused kernels (kernels.cu) and the fused tile kernel (fused.cu):
// the per-cell read source (native Philox draws or replayed reference draws) and small helpers.
#pragma once
#include "philox.cuh"
#include "samplers.cuh"
#include "vgl_internal.h"

#include <math_constants.h>

namespace vgl {

#define VGL_BLOCK 256
// staging capacity per CTA in 4-byte elements: 256 cells x 15 values, rounded up
// to 16 per cell for per-site padding (S == 1), plus head alignment slack
#define VGL_STAGE_ELEMS (VGL_BLOCK * 16 + 8)

__device__ __forceinline__ int qs_squared(int q) { return q == 0 ? 0 : (q < 63 ? q * q : 3969); } // shared.h:459

__device__ __forceinline__ float f32_missing() { return __uint_as_float(VGL_F32_MISSING_BITS); }

struct Read {
    int base, strand, qs, adjqs, tail;
    double eprob;
};

// --------------------------------------------------------------------------
// quality score of a read from its (beta-drawn) error probability, vcfgl.cpp:500-523
__device__ __forceinline__ int bin_qs(const DevParams& p, int q)
{
    if (q < 0 || q > p.bin_max) { // apply_qs_bins() -> ERROR, vcfgl.cpp:63
        atomicExch(p.status, (int)VGL_ERANGE);
        return 0;
    }
    return p.bin_lut[q];
}

__device__ __forceinline__ void qs_from_eprob(const DevParams& p, double e, int& qs, int& adj)
{
    qs = -1;
    adj = -1;
    if (e == 0.0) {
        qs = 63;
    } else if (e == 1.0) {
        qs = 0;
    } else {
        const double phred = -10.0 * log10(e);
        qs = (int)phred;
        if (p.adjust_qs) adj = (int)(phred + p.adjust_by);
    }
    if (p.use_bins) {
        qs = bin_qs(p, qs);
        if (p.adjust_qs) adj = bin_qs(p, adj);
    } else {
        qs = qs > 63 ? 63 : qs;
        if (p.adjust_qs) adj = adj > 63 ? 63 : adj;
    }
}

// --------------------------------------------------------------------------
// One cell's read source: native (Philox) or replay (captured reference draws).
struct CellSource {
    Key key;
    int64_t site;     // global site id
    int64_t cell;     // cell index in the batch
    uint32_t sample;
    int g0, g1;       // true alleles as ACGT ints
    double e_pick;    // base-picking error probability of this site (vcfgl.cpp:425-437)
    int64_t rp_off;   // replay: first read of the cell

    __device__ __forceinline__ void init(const DevParams& p, int64_t c, uint8_t gt)
    {
        cell = c;
        const int64_t sl = c / p.S;
        sample = (uint32_t)(c - sl * p.S);
        site = p.first_site + sl;
        key.k0 = p.k0;
        key.k1 = p.k1;
        g0 = gt & 0xF;
        g1 = gt >> 4;
        e_pick = p.error_rate;
        rp_off = 0;
        if (p.replay) {
            rp_off = p.rp_off[c];
        } else if (p.error_qs == 1) {
            Stream st;
            st.init(key, site, 0xFFFFFFFFu, 0, P_SITE);
            e_pick = beta_draw(st, p.beta_a, p.beta_b);
        }
    }

    __device__ __forceinline__ int depth(const DevParams& p) const
    {
        if (p.replay) return p.rp_depths[cell];
        if (p.depth_mode == VGL_DEPTH_FIXED) return (int)p.depth_mean;
        const double lam = p.depth_mode == VGL_DEPTH_POISSON_PER_SAMPLE ? p.depth_means[sample] : p.depth_mean;
        Stream st;
        st.init(key, site, sample, 0, P_DEPTH);
        return poisson(st, lam);
    }

    __device__ __forceinline__ Read read(const DevParams& p, int i) const
    {
        Read r;
        r.qs = r.adjqs = -1;
        r.eprob = -1.0;
        if (p.replay) {
            const int64_t k = rp_off + i;
            r.base = p.rp_bases[k];
            r.strand = p.rp_strands ? p.rp_strands[k] : 0;
            r.tail = p.rp_tails ? p.rp_tails[k] : 0;
            if (p.error_qs == 2) {
                r.qs = p.rp_qs[k];
                r.adjqs = p.rp_adjqs ? (int)p.rp_adjqs[k] : -1;
                if (p.rp_eprob) r.eprob = p.rp_eprob[k];
            }
            return r;
        }
        const u32x4 w = draw(key, site, sample, (uint32_t)i, P_READ, 0);
        const int truth = (w.y >> 31) ? g1 : g0;                       // vcfgl.cpp:473
        r.base = truth;
        if (u01_32(w.x) < e_pick) r.base = (truth + 1 + (int)mulhi32(w.z, 3u)) & 3; // vcfgl.cpp:485-488
        r.strand = p.sample_strand ? (int)((w.y >> 30) & 1u) : 0;       // vcfgl.cpp:581-586
        const int t = 1 + (int)mulhi32(w.w, 50u);                       // vcfgl.cpp:653-656
        r.tail = t > 25 ? 25 : t;
        if (p.error_qs == 2) {                                          // vcfgl.cpp:494-523
            Stream st;
            st.init(key, site, sample, (uint32_t)i, P_QS);
            r.eprob = beta_draw(st, p.beta_a, p.beta_b);
            qs_from_eprob(p, r.eprob, r.qs, r.adjqs);
        }
        return r;
    }
};

// which reads errmod keeps when a cell has more than 255 (htslib/errmod.c:156-159: shuffle,
// keep 255).  Native mode: sequential selection sampling, keyed per cell -> same subset in
// every kernel that asks.
struct Subsampler {
    Stream st;
    int remaining, need;
    __device__ __forceinline__ void init(const CellSource& cs, int n)
    {
        st.init(cs.key, cs.site, cs.sample, 0, P_SUBSAMPLE);
        remaining = n;
        need = 255;
    }
    __device__ __forceinline__ bool keep()
    {
        // P(keep) = need / remaining
        const bool k = (uint64_t)st.next() * (uint64_t)remaining < ((uint64_t)need << 32);
        --remaining;
        if (k) --need;
        return k;
    }
};

} // namespace vgl
