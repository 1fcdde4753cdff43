// Shared device/host helpers for the genetest_b200 kernels (sm_100a).
#pragma once

#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/genetest_b200.h"

#define GT_FULL 0xffffffffu
#define GT_HD __host__ __device__ __forceinline__

// ---------------------------------------------------------------------------
// Device-resident description of one design (filled by gt_engine_set_design).
// Base vector of sample i:  v_i = [q_i (k) | y_i | 1 | w_i1 .. w_iI],  L = k+2+I
//   q = row of the orthonormal factor of the covariates (C = Q R)
//   y = outcome (centred when the design has an explicit constant column)
// E row of sample i (file order, NCP doubles):
//   [ m_c * v_i  for c = 0..I ]  (NC1 = (I+1) L, padded to NCA)   weighted by g
//   [ m_a * m_b  for a <= b   ]  (NC2 = (I+1)(I+2)/2, padded NCB) weighted by g^2
//   with multipliers m_0 = 1, m_j = w_ij.
// ---------------------------------------------------------------------------
struct GtDesignDev {
    int n_file;        // samples per packed genotype row
    int n;             // design rows
    int k, I, L, p;    // covariates, interactions, base length, parameters
    int NC1, NC2;      // live contraction columns
    int NCA, NCB;      // padded column blocks (kernel instantiation)
    int NCP;           // E row stride = NCA + NCB
    int NPROD;         // L (L + 1) / 2
    int n_chunks;      // ceil(n_file / 256)
    int raw_mode;
    int has_const;     // design spans a constant (k_constant of statsmodels)
    int icpt_col;      // explicit constant column of C (or -1)
    int use_maf_t;
    int degenerate;    // covariates alone are rank deficient
    int lean;          // no interaction: E rows are [q | y] only (NC1 = k + 1, NC2 = 0)
    double icpt_shift; // ybar / value of the constant column
    double trG0;       // trace of the covariates' full-sample Gram matrix = |R|_F^2
    double cond_t, eig_t, maf_t;
    const double *E;        // [n_chunks*256][NCP]
    const double *zero_one; // {0.0, 1.0}: targets of the gather for padding lanes
    const uint8_t *ormask;  // [pitch] 2-bit pairs: 11 = sample not in design
    const double *FF0;      // [L*L]   sum over design rows of v v'
    const double *R;        // [k*k]   row-major upper triangular
    const double *Rinv;     // [k*k]
    const double *tq;       // [n+1]   t.ppf(0.975, df)
    // sexual_chromosome=True: MAF / flip of every marker of the launch come
    // from sex_maf_*_kernel (NULL = nanmean(g) / 2)
    const double *maf_o;
    const int *flip_o;
};

// ---------------------------------------------------------------------------
// Student t / normal tails
// ---------------------------------------------------------------------------
GT_HD double gt_stirling_tail(double z) {
    double z2 = z * z;
    return (1.0 / 12 - (1.0 / 360 - (1.0 / 1260 - (1.0 / 1680 -
            (1.0 / 1188) / z2) / z2) / z2) / z2) / z;
}

// lgamma(a + 1/2) - lgamma(a), without the cancellation of the direct form
GT_HD double gt_lgamma_half_diff(double a) {
    if (a < 20.0) return lgamma(a + 0.5) - lgamma(a);
    return a * log1p(0.5 / a) - 0.5 + 0.5 * log(a) +
           gt_stirling_tail(a + 0.5) - gt_stirling_tail(a);
}

// Continued fraction of the incomplete beta function, 1/(1+ d1/(1+ d2/(1+ ...))),
// evaluated with Steed's forward algorithm: one reciprocal per partial
// quotient, and the two coefficients of a step share one more.
GT_HD double gt_betacf(double a, double b, double x) {
    const double qab = a + b, qap = a + 1.0, qam = a - 1.0;
    double D = 1.0 / (1.0 - qab * x / qap);
    double delta = D - 1.0;
    double f = 1.0 + delta;
    for (int m = 1; m <= 300; ++m) {
        const double m2 = 2.0 * m;
        const double q1 = qam + m2, q2 = a + m2, q3 = qap + m2;
        const double r = 1.0 / (q1 * q2 * q3);
        const double aa1 = m * (b - m) * x * q3 * r;
        const double aa2 = -(a + m) * (qab + m) * x * q1 * r;
        D = 1.0 / (1.0 + aa1 * D);
        delta *= D - 1.0;
        f += delta;
        D = 1.0 / (1.0 + aa2 * D);
        delta *= D - 1.0;
        f += delta;
        if (fabs(delta) < 1e-16 * fabs(f)) break;
    }
    return f;
}

// log of the t density normaliser: -ln( sqrt(df) B(df/2, 1/2) ) + ...
GT_HD double gt_lbeta_half(double a) {
    return 0.5 * log(3.14159265358979323846) - gt_lgamma_half_diff(a);
}

// I_x(a, 1/2) for large a by the asymptotic expansion of DiDonato & Morris
// (ACM TOMS 18, 1992, "BGRAT": sum of p_n J_n(b, u) with u = -(a + (b-1)/2) ln x),
// specialised to b = 1/2 where Q(b, u) = erfc(sqrt(u)) and the p_n are constants.
// t2 = t^2, a = df / 2 >= 250: at most 12 terms reach 1e-13 relative for every
// |t| (checked against 40-digit references in tests/test_host_api.py); for
// df ~ 1e5 three terms do.  Returns -1 when exp(-u) would underflow.
GT_HD double gt_t_sf2_large_df(double t2, double df) {
    const double pn[12] = {1.0, -0.08333333333333333, 0.00625, -0.0005042989417989418,
                           4.343722442680776e-05, -3.896385732323232e-06, 3.5833354634507906e-07,
                           -3.349747072060578e-08, 3.1675854343161146e-09, -3.0210368517370963e-10,
                           2.900415787669253e-11, -2.799396748375269e-12};
    const double a = 0.5 * df, T = a - 0.25;
    const double lx = -log1p(t2 / df);            // ln x, x = df / (df + t^2)
    const double u = -T * lx;
    if (!(u < 700.0)) return -1.0;
    const double su = sqrt(u);
    const double h = su * exp(-u) * 0.56418958354775628695;          // u^b e^-u / Gamma(b)
    const double prefix = h * exp(gt_lgamma_half_diff(a)) / sqrt(T);
    double j = erfc(su) / h;
    double sum = j;                                                  // p_0 = 1
    const double lx2 = 0.25 * lx * lx, t4inv = 1.0 / (4.0 * T * T);
    double lxp = 1.0, b2n = 0.5;
    for (int n = 1; n < 12; ++n) {
        j = (b2n * (b2n + 1.0) * j + (u + b2n + 1.0) * lxp) * t4inv;
        lxp *= lx2;
        b2n += 2.0;
        const double r = pn[n] * j;
        sum += r;
        if (fabs(r) < 1e-17 * fabs(sum)) break;
    }
    return prefix * sum;
}

// Two-sided tail 2 * P(T_df > |t|) = I_x(df/2, 1/2), x = df / (df + t^2).
GT_HD double gt_t_sf2(double t, double df) {
    if (!(df > 0.0) || t != t) return nan("");
    double t2 = t * t;
    if (t2 == 0.0) return 1.0;
    if (isinf(t2)) return 0.0;
    if (df >= 500.0) {
        // the continued fraction below needs O(sqrt(df)) partial quotients here
        const double v = gt_t_sf2_large_df(t2, df);
        if (v >= 0.0) return v;
    }
    double a = 0.5 * df;
    double lnpre = -gt_lbeta_half(a) - a * log1p(t2 / df) +
                   0.5 * (log(t2) - log(df + t2));
    double x = df / (df + t2);
    // direct fraction in the tail; the complement converges faster (and loses
    // at most ~3 digits) for |t| < 3 once df is not tiny
    bool direct = x < (a + 1.0) / (a + 2.5);
    if (direct && t2 < 9.0 && df > 10.0) direct = false;
    if (direct)
        return exp(lnpre) * gt_betacf(a, 0.5, x) / a;
    return 1.0 - exp(lnpre) * gt_betacf(0.5, a, t2 / (df + t2)) / 0.5;
}

GT_HD double gt_t_pdf(double t, double df) {
    double a = 0.5 * df;
    return exp(-gt_lbeta_half(a) - 0.5 * log(df) - (a + 0.5) * log1p(t * t / df));
}

// t.ppf(0.975, df): Cornish-Fisher start, Newton on the two-sided tail.
GT_HD double gt_t_ppf975(double df) {
    if (!(df > 0.0)) return nan("");
    const double z = 1.959963984540054;
    double t;
    if (df < 1.5) t = 12.706204736174694;
    else if (df < 2.5) t = 4.302652729749464;
    else {
        double z3 = z * z * z, z5 = z3 * z * z, z7 = z5 * z * z;
        t = z + (z3 + z) / (4 * df) + (5 * z5 + 16 * z3 + 3 * z) / (96 * df * df) +
            (3 * z7 + 19 * z5 + 17 * z3 - 15 * z) / (384 * df * df * df);
    }
    for (int it = 0; it < 12; ++it) {
        double f = gt_t_sf2(t, df) - 0.05;
        double step = f / (2.0 * gt_t_pdf(t, df));
        t += step;
        if (fabs(step) <= 1e-15 * t) break;
    }
    return t;
}

// 2 * Phi(-|z|)
GT_HD double gt_norm_sf2(double z) {
    if (z != z) return nan("");
    return erfc(fabs(z) * 0.70710678118654752440);
}

#define GT_Z975 1.959963984540054

// ---------------------------------------------------------------------------
// Warp helpers
// ---------------------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ double gt_warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(GT_FULL, v, o);
    return v;
}
__device__ __forceinline__ int gt_warp_sum_int(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(GT_FULL, v, o);
    return v;
}
__device__ __forceinline__ double gt_warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(GT_FULL, v, o));
    return v;
}

// In-place lower Cholesky of the p x p matrix A (leading dimension ld) held
// in shared memory, executed by one warp.  Returns false when a pivot is not
// positive (relative to the original diagonal).
__device__ __forceinline__ bool gt_warp_cholesky(double *A, int ld, int p, int lane) {
    bool ok = true;
    for (int j = 0; j < p; ++j) {
        double d = A[j * ld + j];
        if (!(d > 0.0)) { ok = false; break; }
        double s = sqrt(d);
        double inv = 1.0 / s;
        __syncwarp();
        for (int i = j + 1 + lane; i < p; i += 32) A[i * ld + j] *= inv;
        if (lane == 0) A[j * ld + j] = s;
        __syncwarp();
        int m = p - j - 1;
        for (int t = lane; t < m * m; t += 32) {
            int i = j + 1 + t / m, c = j + 1 + t % m;
            if (c <= i) A[i * ld + c] -= A[i * ld + j] * A[c * ld + j];
        }
        __syncwarp();
    }
    return ok;
}

// X = (L L')^{-1}: lane c solves column c (forward then backward substitution).
__device__ __forceinline__ void gt_warp_chol_inverse(const double *Lc, int ld, int p,
                                                     double *X, int lane) {
    for (int c = lane; c < p; c += 32) {
        for (int i = 0; i < p; ++i) {
            double s = (i == c) ? 1.0 : 0.0;
            for (int j = 0; j < i; ++j) s -= Lc[i * ld + j] * X[j * ld + c];
            X[i * ld + c] = s / Lc[i * ld + i];
        }
        for (int i = p - 1; i >= 0; --i) {
            double s = X[i * ld + c];
            for (int j = i + 1; j < p; ++j) s -= Lc[j * ld + i] * X[j * ld + c];
            X[i * ld + c] = s / Lc[i * ld + i];
        }
    }
    __syncwarp();
}

#endif  // __CUDACC__
