// Distribution samplers of the native simulator (device side).
//
// The reference draws with sequential libc/libstdc++ generators
// (Poisson: rng.h:284-351 Knuth product / NR rejection; Beta: rng.h:354-421
// two std::gamma_distribution on std::mt19937).  Native mode only has to match
// their DISTRIBUTIONS (north_star), so we use counter-based, branch-light
// standard algorithms: CDF inversion + Hoermann's PTRS for Poisson,
// Marsaglia-Tsang for Gamma, Box-Muller for normals.
#pragma once
#include "philox.cuh"
#include <math.h>

namespace vgl {

// Poisson(lam) by sequential CDF search; exact to the 53-bit uniform. Use for lam < 10.
__device__ __forceinline__ int poisson_inversion(Stream& st, double lam)
{
    const double u = st.uniform();
    double p = exp(-lam), cdf = p;
    int k = 0;
    while (u > cdf && k < 1000) {
        ++k;
        p *= lam / (double)k;
        cdf += p;
        if (p < 1e-300) break;
    }
    return k;
}

// W. Hoermann, "The transformed rejection method for generating Poisson random
// variables", Insurance: Mathematics and Economics 12 (1993): algorithm PTRS. lam >= 10.
__device__ __forceinline__ int poisson_ptrs(Stream& st, double lam)
{
    const double slam = sqrt(lam), loglam = log(lam);
    const double b = 0.931 + 2.53 * slam;
    const double a = -0.059 + 0.02483 * b;
    const double invalpha = 1.1239 + 1.1328 / (b - 3.4);
    const double vr = 0.9277 - 3.6224 / (b - 2.0);
    for (int it = 0; it < 1000; ++it) {
        const double U = st.uniform() - 0.5;
        const double V = st.uniform();
        const double us = 0.5 - fabs(U);
        const double kf = floor((2.0 * a / us + b) * U + lam + 0.43);
        if (us >= 0.07 && V <= vr) return (int)kf;
        if (kf < 0.0 || (us < 0.013 && V > us)) continue;
        if (log(V) + log(invalpha) - log(a / (us * us) + b) <= -lam + kf * loglam - lgamma(kf + 1.0)) return (int)kf;
    }
    return (int)lam;
}

__device__ __forceinline__ int poisson(Stream& st, double lam)
{
    if (!(lam > 0.0)) return 0;
    return lam < 10.0 ? poisson_inversion(st, lam) : poisson_ptrs(st, lam);
}

__device__ __forceinline__ double std_normal(Stream& st)
{
    // Box-Muller; one of the pair is discarded to keep the stream stateless
    const double u1 = st.uniform(), u2 = st.uniform();
    return sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
}

// G. Marsaglia, W. W. Tsang, "A simple method for generating gamma variables", ACM TOMS 26 (2000)
__device__ __forceinline__ double gamma_mt(Stream& st, double shape)
{
    double boost = 1.0;
    if (shape < 1.0) {
        boost = pow(st.uniform(), 1.0 / shape);
        shape += 1.0;
    }
    const double d = shape - 1.0 / 3.0, c = 1.0 / sqrt(9.0 * d);
    for (int it = 0; it < 1000; ++it) {
        const double x = std_normal(st);
        double v = 1.0 + c * x;
        if (v <= 0.0) continue;
        v = v * v * v;
        const double u = st.uniform();
        const double x2 = x * x;
        if (u < 1.0 - 0.0331 * x2 * x2) return boost * d * v;
        if (log(u) < 0.5 * x2 + d * (1.0 - v + log(v))) return boost * d * v;
    }
    return boost * d;
}

// Beta(a, b) = X / (X + Y), X ~ Gamma(a), Y ~ Gamma(b)   (rng.h:408-419)
__device__ __forceinline__ double beta_draw(Stream& st, double a, double b)
{
    const double x = gamma_mt(st, a);
    const double y = gamma_mt(st, b);
    return x / (x + y);
}

} // namespace vgl
