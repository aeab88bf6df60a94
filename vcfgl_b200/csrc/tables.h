// Host-side constant tables of the simulation core (built once in vgl_create()).
#pragma once
#include <stdint.h>
#include <vector>

namespace vgl {

// samtools/bcftools "revised MAQ" error-model coefficients
// (reference: htslib/errmod.c:51-125, errmod_init(1 - theta) at io.cpp:1276)
struct ErrmodTables {
    std::vector<double> fk;   // [256]          dependency decay  fk[n] = (1-depcorr)^n (1-eta) + eta
    std::vector<double> beta; // [64][256][256] phred-scaled binomial tail ratios, index q<<16 | n<<8 | k
    std::vector<double> lhet; // [256][256]     log C(n,k) - n ln 2, index n<<8 | k
    void build(double depcorr, double eta = 0.03);
    // for a FIXED quality score q: the running sum errmod_cal() would reach after walking
    // c reads of one base at depth n:  bsum[n<<8 | c] = sum_{i<c} fk[i] * beta[q][n][i]
    // (errmod.c:174-177 with w == c because the strand bit is never set, gl_methods.cpp:329)
    std::vector<double> fixed_q_bsum(int q) const;
    // -4.343 * lhet[n<<8 | k]  (errmod.c:200)
    std::vector<double> het_term() const;
    // true when every errmod score (a float sum of non-negative table terms) is +0 or >= 2^-100 and
    // finite: each positive term of both tables is itself >= 2^-100, and all are finite and >= 0
    static bool scores_safe_for_fast_div(const std::vector<double>& bsum, const std::vector<double>& het);
};

// Poisson(lambda) CDF as 2^64 fixed-point thresholds: n = smallest k with u < cdf[k] for a 64-bit
// uniform u.  At most max_n entries; the last entry is 2^64-1.  (native count-level sampler)
std::vector<unsigned long long> poisson_cdf_u64(double lambda, int max_n);

// Walker alias table over 256 columns for the same distribution (tile kernel): with a 64-bit uniform u,
// column = u >> 56; the cell keeps `column` when (u << 8) < (entry & ~0xFF), else takes entry & 0xFF.
// Built in exact integer arithmetic from the CDF above, so P(n) is reproduced to 2^-64.
// Returns an empty vector when the support does not fit 256 outcomes.
std::vector<unsigned long long> poisson_alias_u64(const std::vector<unsigned long long>& cdf);

// number of mis-called reads E ~ Binomial(n, e): thresholds floor(P(E <= j | n) * 2^32), j = 0..3, n = 0..255
// (saturated at 2^32-1); E = #{j : u >= t[n][j]} for a 32-bit uniform u, 4 = "beyond the table".
std::vector<uint32_t> binomial_cdf4_u32(double e);

// regularized incomplete beta function I_x(a, b) (continued fraction, Lentz), x in [0, 1]
double inc_beta(double a, double b, double x);

// Per-read (quality score used by the GL, mis-called or not) classes of a Beta(a, b) error probability p
// (--error-qs 2; vcfgl.cpp:494-523): phred = -10 log10 p, q = (int)(phred + shift) (shift = --adjust-by when the
// GL uses the adjusted score, else 0), then --qs-bins or the cap at 63.  P(q) = I(p_hi) - I(p_lo) and
// P(q and error) = E[p; p in the class] = a/(a+b) (I_{a+1,b}(p_hi) - I_{a+1,b}(p_lo)) since the read is mis-called
// with probability p (vcfgl.cpp:485).  Returned: 512 words = a Walker alias table over 256 columns
// (threshold24 << 8 | alias; probabilities quantised to 2^-32) followed by the class info words (q | err << 8);
// `prob` receives the quantised class probabilities (for tests).  Empty when the classes do not fit or a class
// with positive probability falls outside the --qs-bins ranges (the reference exits there, vcfgl.cpp:63).
std::vector<uint32_t> qs_class_table(double a, double b, double shift, bool use_bins, const uint8_t* bin_lut, int bin_max,
                                     std::vector<double>* prob = nullptr);

// qScore_to_log10_gl[3][257] (shared.cpp:110-114)
extern const double kLutLog10Gl[3][257];

// derived per-run constants (reference: preCalcStruct io.h:22-32, filled at vcfgl.cpp:1661-1743)
struct PreCalc {
    int qs = -1, adj_qs = -1;
    double homT = -1.0, het = -1.0, homF = -1.0;
};

// returns 0, or -1 when a qs falls outside the --qs-bins ranges (vcfgl.cpp:63)
int precalc(double error_rate, int error_qs, int gl_model, int precise_gl, int adjust_qs, double adjust_by,
            int n_bins, const uint8_t bins[][3], PreCalc* out);

} // namespace vgl
