// pbn_fit.cuh — exact evaluation of one predictor candidate of the Bittner network fitter.
//
// The reference (gym_PBN/envs/bittner/gen/predictor_sets.py:105-124 gen_COD) fits y ~ [x_a, x_b, x_c, 1] . A by
// least squares through pinv(X'X) in float64, ROUNDS the fitted values and scores the predictor by the number of
// misclassified samples.  All of X and y is binary and there are at most 32 samples, so a gene row is one 32-bit
// mask (bit s = sample s) and the whole fit is a function of popcounts.  The normal equations are solved here in
// exact integer arithmetic (fraction-free Gauss-Jordan / Bareiss on the 4x4 Gram matrix; every intermediate is a
// minor of [X'X | X'y] and stays below 2^53), so the rounded fitted value of every occupied input pattern is known
// exactly — except when it is exactly a half-integer, where the reference's answer is decided by float noise of its
// LAPACK build.  Those candidates are reported with both outcomes (k_lo != k_hi) and settled by the caller.
//
// Singular Gram matrices (a constant gene, two identical or complementary rows) need no pivoting: the Schur
// complements of a positive semi-definite matrix stay PSD, so a zero pivot means a zero row and column, i.e. a free
// coefficient, set to 0.  ANY solution of the normal equations gives the same fitted values on the sample rows as the
// reference's minimum-norm pinv solution.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define PBN_HD __host__ __device__ __forceinline__
#else
#define PBN_HD inline
#endif

PBN_HD int fit_popc(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return __popc(x);
#else
    return __builtin_popcount(x);
#endif
}

// Squared-error count sum_s (round(fitted_s) - y_s)^2 over the n_samples samples: k_lo <= k_hi, equal unless a
// half-integer fitted value falls on a pattern whose two roundings cost differently.
PBN_HD void fit_eval(uint32_t ma, uint32_t mb, uint32_t mc, uint32_t my, int n_samples, int *k_lo, int *k_hi) {
    const uint32_t full = n_samples >= 32 ? 0xFFFFFFFFu : ((1u << n_samples) - 1u);
    long long M[4][5];
    // variable order [1, a, b, c]; column 4 = X'y
    M[0][0] = n_samples;
    M[0][1] = M[1][0] = M[1][1] = fit_popc(ma);
    M[0][2] = M[2][0] = M[2][2] = fit_popc(mb);
    M[0][3] = M[3][0] = M[3][3] = fit_popc(mc);
    M[1][2] = M[2][1] = fit_popc(ma & mb);
    M[1][3] = M[3][1] = fit_popc(ma & mc);
    M[2][3] = M[3][2] = fit_popc(mb & mc);
    M[0][4] = fit_popc(my);
    M[1][4] = fit_popc(ma & my);
    M[2][4] = fit_popc(mb & my);
    M[3][4] = fit_popc(mc & my);
    long long prev = 1;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const long long p = M[k][k];
        if (p == 0) continue;  // free coefficient (row and column of the Schur complement are zero)
#if defined(__CUDA_ARCH__)
        // the quotient is an integer below 2^31 and the dividend is below 2^53: dividend * (1/prev) in float64 is
        // within 2^-20 of it, so rounding to nearest recovers it exactly — three instructions instead of a 64-bit divide
        const double inv_prev = 1.0 / (double)prev;
#endif
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            if (i == k) continue;
            const long long f = M[i][k];
#pragma unroll
            for (int j = 0; j < 5; ++j) {
                const long long num = p * M[i][j] - f * M[k][j];
#if defined(__CUDA_ARCH__)
                M[i][j] = k == 0 ? num : __double2ll_rn(__ll2double_rn(num) * inv_prev);
#else
                M[i][j] = num / prev;  // exact division
#endif
            }
        }
        prev = p;
    }
    const long long den = prev;  // > 0; every solved row has M[i][i] == den, free rows are zero
    int lo = 0, hi = 0;
#pragma unroll
    for (int pat = 0; pat < 8; ++pat) {
        const uint32_t sel = ((pat & 4) ? ma : ~ma) & ((pat & 2) ? mb : ~mb) & ((pat & 1) ? mc : ~mc) & full;
        const int n = fit_popc(sel);
        if (n == 0) continue;
        const int s1 = fit_popc(sel & my);
        const long long num = M[0][4] + ((pat & 4) ? M[1][4] : 0) + ((pat & 2) ? M[2][4] : 0) + ((pat & 1) ? M[3][4] : 0);
        // q = floor(fitted + 1/2) with fitted = num/den; r == 0 marks fitted = q - 1/2 exactly.  |fitted| <= sqrt(S), and
        // inside [-1.5, 2.5) three comparisons replace the 64-bit divide.
        const long long t2 = 2 * num;
        long long q, r;
        if (t2 >= -3 * den && t2 < 5 * den) {
            q = -1 + (t2 >= -den) + (t2 >= den) + (t2 >= 3 * den);
            r = (t2 == -3 * den || t2 == -den || t2 == den || t2 == 3 * den) ? 0 : 1;
        } else {
            const long long t = t2 + den, d2 = 2 * den;
            q = t / d2;
            r = t - q * d2;
            if (r < 0) { r += d2; q -= 1; }  // floor division
        }
        const int qi = (int)q;
        const int e_q = s1 * (qi - 1) * (qi - 1) + (n - s1) * qi * qi;
        if (r == 0) {  // fitted value is exactly q - 1/2
            const int e_m = s1 * (qi - 2) * (qi - 2) + (n - s1) * (qi - 1) * (qi - 1);
            lo += e_q < e_m ? e_q : e_m;
            hi += e_q < e_m ? e_m : e_q;
        } else {
            lo += e_q;
            hi += e_q;
        }
    }
    *k_lo = lo;
    *k_hi = hi;
}

// Candidate key: smaller = better.  rank:10 | a:12 | b:12 | c:12 | y:4 | sc:12 — rank = class of the COD (0 = highest),
// the rest is the order in which the reference visits candidates (combination, target row, input-row product:
// predictor_sets.py:60-76), so equal CODs keep the earlier one, as add_to_buff's strict `<` does (:80-102).
#define FIT_ARRIVAL_BITS 52
#define FIT_ARRIVAL_MASK ((1ull << FIT_ARRIVAL_BITS) - 1ull)
#define FIT_KEY_NONE 0xFFFFFFFFFFFFFFFFull
PBN_HD unsigned long long fit_arrival(int a, int b, int c, int y, int sc) {
    return ((unsigned long long)a << 40) | ((unsigned long long)b << 28) | ((unsigned long long)c << 16) |
           ((unsigned long long)y << 12) | (unsigned long long)sc;
}
