// GL model 1 with one quality score for every read (gl_methods.cpp:304-369 +
// htslib/errmod.c:143-208): the 5x5 errmod matrix depends only on the four base counts.
// Branch-free restatement in base-pair space, bit-exact with the reference's float/double mixing.
#pragma once
#include "vgl_internal.h"

#include <math_constants.h>

namespace vgl {

// base-pair index of the unordered pair (j <= k): k*(k+1)/2 + j, j,k in 0..4 (4 = unobserved allele)
// pairmap: nibble `pair` holds the output (allele-space) genotype slot of that base pair, 0xF = the
// site does not have both alleles.  Built once per site.
__device__ __forceinline__ uint64_t make_pairmap(const int (&b2a)[5])
{
    uint64_t m = 0;
#pragma unroll
    for (int k = 0; k < 5; ++k)
#pragma unroll
        for (int j = 0; j <= k; ++j) {
            const int aj = b2a[j], ak = b2a[k];
            uint64_t slot = 0xF;
            if (aj >= 0 && ak >= 0) {
                const int hi = aj > ak ? aj : ak, lo = aj > ak ? ak : aj;
                slot = (uint64_t)(hi * (hi + 1) / 2 + lo); // bcf_alleles2gt, htslib/vcf.h:902
            }
            m |= slot << (4 * (k * (k + 1) / 2 + j));
        }
    return m;
}

// float accumulator fed a double (errmod.c:182,187,197): tmp1 += bsum  <=>  (float)((double)tmp1 + bsum)
__device__ __forceinline__ float acc_add(float acc, double b) { return __double2float_rn(__dadd_rn((double)acc, b)); }

// (float)((-1.0 * (double)q) / 10.0), gl_methods.cpp:343.  One correctly rounded float division gives
// the same float: q/10 is never closer than 0.1 ulp to a float rounding boundary, so the double
// intermediate cannot change the result.
__device__ __forceinline__ float neg_div10_ref(float q) { return -__fdiv_rn(q, 10.0f); }

// The same value in three instructions: y = q*RN(0.1); r = fma(-10, y, q) (exact residual);
// y' = fma(r, RN(0.1), y).  vgl_selftest() checks it against __fdiv_rn for +0 and EVERY finite float
// >= 2^-100 (tests/test_gpu_selftest.py; it is wrong only for tinier values, -0 and +inf).  The
// host proves at table-build time that no score below 2^-100 other than +0 can occur
// (ErrmodTables::scores_safe_for_fast_div) and passes that as `fast`.
__device__ __forceinline__ float neg_div10_fast(float q)
{
    const float y = __fmul_rn(q, 0.1f);
    const float r = __fmaf_rn(-10.0f, y, q);
    return -__fmaf_rn(r, 0.1f, y);
}
__device__ __forceinline__ float neg_div10(float q, bool fast) { return fast ? neg_div10_fast(q) : neg_div10_ref(q); }

// lroundf((float)(-10.0 * (double)gl)) capped at 255 (vcfgl.cpp:931-934) for gl <= 0: the double
// product is exact so one float multiply rounds identically; round-half-away via trunc + fraction.
__device__ __forceinline__ int pl_from_gl_ref(float gl) // the reference's expression, for the self-test
{
    if (gl == -CUDART_INF_F) return 255;
    const long x = lroundf(__double2float_rn(__dmul_rn(-10.0, (double)gl)));
    return x > 255 ? 255 : (int)x;
}

__device__ __forceinline__ int pl_from_gl(float gl)
{
    const float x = fminf(__fmul_rn(-10.0f, gl), 300.0f);
    int i = __float2int_rz(x);
    i += (x - (float)i) >= 0.5f;
    return i > 255 ? 255 : i;
}

// lroundf((float)(-10.0 * (double)gl)) capped at 255 (vcfgl.cpp:931-934) for gl <= 0 without a
// float->int conversion: the product is exact in double so one float multiply rounds identically;
// floor(x + 0.5) by a round-toward-zero add (x >= 0), and the integer read off the mantissa after
// adding 2^23.  Checked against the reference expression for every float <= 0 (vgl_selftest).
__device__ __forceinline__ int pl_from_gl_magic(float gl)
{
    const float x = fminf(__fmul_rn(-10.0f, gl), 255.0f);
    const float u = __fadd_rz(__fadd_rz(x, 0.5f), 8388608.0f);
    return __float_as_int(u) & 0x1FF;
}

// ---- packed fp32x2 forms (sm_100 FMUL2 / FFMA2 / FADD2: two IEEE operations per instruction, same
// roundings as the scalar forms above)
struct f32x2 {
    unsigned long long v;
};
__device__ __forceinline__ f32x2 pack2(float lo, float hi)
{
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(f32x2 x, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(x.v)); }
// q / 10 for both halves (the value neg_div10_fast negates)
__device__ __forceinline__ f32x2 div10_fast2(f32x2 q)
{
    const f32x2 c01 = pack2(0.1f, 0.1f), cm10 = pack2(-10.0f, -10.0f);
    f32x2 y, r, w;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(y.v) : "l"(q.v), "l"(c01.v));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(cm10.v), "l"(y.v), "l"(q.v));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(w.v) : "l"(r.v), "l"(c01.v), "l"(y.v));
    return w;
}
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b)
{
    f32x2 r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
    return r;
}
// mantissa-encoded PL of both halves: bits = 0x4B000000 + floor(-10 * gl + 0.5) (uncapped), gl <= 0
__device__ __forceinline__ f32x2 pl_magic2(f32x2 gl)
{
    const f32x2 cm10 = pack2(-10.0f, -10.0f), h = pack2(0.5f, 0.5f), m23 = pack2(8388608.0f, 8388608.0f);
    f32x2 x, t, u;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(x.v) : "l"(gl.v), "l"(cm10.v));
    asm("add.rz.f32x2 %0, %1, %2;" : "=l"(t.v) : "l"(x.v), "l"(h.v));
    asm("add.rz.f32x2 %0, %1, %2;" : "=l"(u.v) : "l"(t.v), "l"(m23.v));
    return u;
}
// scalar twin of pl_magic2 + the cap: min(bits - 0x4B000000, 255) in one VIADDMNMX (also maps +inf to 255)
__device__ __forceinline__ int pl_from_magic_bits(float u) { return (int)__viaddmin_u32(__float_as_uint(u), 0xB5000000u, 255u); }
__device__ __forceinline__ int pl_from_gl_magic_uncapped(float gl)
{
    const float u = __fadd_rz(__fadd_rz(__fmul_rn(-10.0f, gl), 0.5f), 8388608.0f);
    return pl_from_magic_bits(u);
}

// phred-scaled errmod likelihoods q[pair] for all 15 base pairs from the counts c0..c3 of a cell with
// n = c0+c1+c2+c3 reads, 1 <= n <= 255.  bsum = fixed-qs running-sum table [n<<8|c], het = -4.343*lhet.
template <bool CLAMP>
__device__ __forceinline__ void m1f_scores_t(int n, int c0, int c1, int c2, int c3, const double* __restrict__ bsum,
                                             const double* __restrict__ het, float (&q)[15])
{
    const double* row = bsum + (n << 8);
    const double b0 = __ldg(row + c0), b1 = __ldg(row + c1), b2 = __ldg(row + c2), b3 = __ldg(row + c3);
    // sequential float sums over base subsets, in base order (adding a 0.0 term is the identity, so
    // "skip if no reads" branches of the reference are not needed)
    const float f0 = __double2float_rn(b0), f1 = __double2float_rn(b1), f2 = __double2float_rn(b2);
    const float p01 = acc_add(f0, b1), p02 = acc_add(f0, b2), p03 = acc_add(f0, b3);
    const float p12 = acc_add(f1, b2), p13 = acc_add(f1, b3), p23 = acc_add(f2, b3);
    const float t012 = acc_add(p01, b2), t013 = acc_add(p01, b3), t023 = acc_add(p02, b3), t123 = acc_add(p12, b3);
    const float q0123 = acc_add(t012, b3);
    // homozygous jj: everything that is not j (errmod.c:185-191)
    q[0] = t123; q[2] = t023; q[5] = t013; q[9] = t012; q[14] = q0123;
    // heterozygous jk, j<k<4 (errmod.c:193-202): -4.343*lhet[cj+ck][ck] + the other two bases
#define VGL_HET(cj, ck, rest) __double2float_rn(__dadd_rn(__ldg(het + (((cj) + (ck)) << 8 | (ck))), (double)(rest)))
    q[1] = VGL_HET(c0, c1, p23);
    q[3] = VGL_HET(c0, c2, p13);
    q[4] = VGL_HET(c1, c2, p03);
    q[6] = VGL_HET(c0, c3, p12);
    q[7] = VGL_HET(c1, c3, p02);
    q[8] = VGL_HET(c2, c3, p01);
    // j with the unobserved allele (index 4, no reads): lhet[cj][0] + everything that is not j
    q[10] = VGL_HET(c0, 0, t123);
    q[11] = VGL_HET(c1, 0, t023);
    q[12] = VGL_HET(c2, 0, t013);
    q[13] = VGL_HET(c3, 0, t012);
#undef VGL_HET
    if (CLAMP) {
#pragma unroll
        for (int i = 0; i < 15; ++i) q[i] = fmaxf(q[i], 0.0f); // errmod.c:204
    }
}

__device__ __forceinline__ void m1f_scores(int n, int c0, int c1, int c2, int c3, const double* __restrict__ bsum,
                                           const double* __restrict__ het, float (&q)[15])
{
    m1f_scores_t<true>(n, c0, c1, c2, c3, bsum, het, q);
}

// errmod.c:204 clamps negative scores to 0.  When the host has proven that every table term is >= +0
// (ErrmodTables::scores_safe_for_fast_div) the scores are sums of non-negative terms, never negative,
// -0 or NaN, and the clamp is the identity.
__device__ __forceinline__ void m1f_scores_noclamp(int n, int c0, int c1, int c2, int c3, const double* __restrict__ bsum,
                                                   const double* __restrict__ het, float (&q)[15])
{
    m1f_scores_t<false>(n, c0, c1, c2, c3, bsum, het, q);
}

// scatter the base-pair scores into the cell's allele-ordered GL slots (shared memory), rescale to
// max 0 (gl_methods.cpp:338-357)
__device__ __forceinline__ void m1f_store_gl(const float (&q)[15], uint64_t pairmap, int G, float* my_gl)
{
    float mx = -CUDART_INF_F;
#pragma unroll
    for (int i = 0; i < 15; ++i) {
        const int slot = (int)((pairmap >> (4 * i)) & 0xF);
        const float v = neg_div10_ref(q[i]);
        if (slot != 0xF) {
            my_gl[slot] = v;
            mx = fmaxf(mx, v);
        }
    }
    for (int g = 0; g < G; ++g) my_gl[g] = __fsub_rn(my_gl[g], mx);
}

} // namespace vgl
