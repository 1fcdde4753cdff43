// K3b, thread-per-SNP version for designs without interaction and at most 13
// parameters: what linear_solve_kernel (gt_solve.cuh) does with 16 lanes, shared
// memory and a barrier after every elimination step, done by ONE thread with the
// whole p x p system in registers (fully unrolled, no synchronisation at all);
// loads of the corrections and all stores are coalesced over the markers.
// Restates sm.OLS(y, X).fit() + StatsLinear._results_handler
// (genetest/statistics/models/linear.py:46-104) for the complete cases of each
// marker, after _gwas_worker's MAF / flip steps (analysis.py:127-157).
//
// Markers whose condition-number / eigenvalue gate cannot be decided by the
// certified trace bound, or whose Cholesky factorisation breaks down, are left
// to the 16-lane kernel (it runs the Jacobi eigenvalue path): this kernel marks
// them with status GT_SOLVE_DEFER and the other one only picks those up.
#pragma once

#include "gt_common.cuh"

namespace gtk {

constexpr int GT_SOLVE_DEFER = -77;     // internal: finish this marker in linear_solve_kernel

// NI = 1: one SNP x covariate interaction column (design [Q | g | g w], analysis.py:159-166)
template <int P, int NI = 0>
__global__ void __launch_bounds__(128)
linear_solve_thread_kernel(GtDesignDev D, int n_snps, const double *__restrict__ S,
                           const int *__restrict__ counts, const double *__restrict__ gsum,
                           const double *__restrict__ corrT, long long ldc,
                           double *__restrict__ out, int *__restrict__ iout, long long ostride) {
    constexpr int K = P - 1 - NI;        // covariates (orthonormal basis q)
    constexpr int L = K + 2 + NI;        // base vector [q | y | 1 | w]
    constexpr int IY = K, I1 = K + 1, IW = K + 2;
    __shared__ double sFF0[L * L];
    __shared__ double sRinv[K * K + 1];
    for (int t = threadIdx.x; t < L * L; t += blockDim.x) sFF0[t] = D.FF0[t];
    for (int t = threadIdx.x; t < K * K; t += blockDim.x) sRinv[t] = D.Rinv[t];
    __syncthreads();
    const int snp = blockIdx.x * blockDim.x + threadIdx.x;
    if (snp >= n_snps) return;
    constexpr int NF = GT_F_PARAM0 + 6 * P;
    auto fail = [&](int status, int nobs, int flip, double maf, double aux) {
#pragma unroll 1
        for (int f = 0; f < NF; ++f) {
            double v = nan("");
            if (f == GT_F_MAF) v = maf;
            if (f == GT_F_AUX) v = aux;
            out[(long long)f * ostride + snp] = v;
        }
        iout[(long long)GT_I_STATUS * ostride + snp] = status;
        iout[(long long)GT_I_NOBS * ostride + snp] = nobs;
        iout[(long long)GT_I_FLIP * ostride + snp] = flip;
        iout[(long long)GT_I_NITER * ostride + snp] = 0;
    };
    const int n_miss = counts[(long long)snp * 4 + 2];
    const int nj = D.n - n_miss;
    if (nj <= 0) { fail(GT_ALL_MISSING, 0, 0, nan(""), nan("")); return; }
    const double sum_g_all = gsum[(long long)snp * 2], sum_g2_all = gsum[(long long)snp * 2 + 1];
    double maf = nan("");
    int flip = 0;
    if (!D.raw_mode) {
        maf = (sum_g_all / (double)nj) / 2.0;
        if (maf > 0.5) { maf = 1.0 - maf; flip = 1; }
        if (D.maf_o != nullptr) { maf = D.maf_o[snp]; flip = D.flip_o[snp]; }
        if (D.use_maf_t && maf < D.maf_t) { fail(GT_MAF_BELOW, nj, flip, maf, maf); return; }
    }
    // complete-case Gram entry (a <= b) of the base vectors
    const double *cs = corrT + snp;
    auto s0 = [&](int a, int b) -> double {
        double v = sFF0[a * L + b];
        if (n_miss > 0) v -= __ldg(cs + (long long)(a * L - a * (a - 1) / 2 + (b - a)) * ldc);
        return v;
    };
    // A: lower triangle of the normal equations in the [Q | g] basis, row-major packed
    double A[P * (P + 1) / 2];
    double vec[P];                       // rhs -> u -> bq -> beta, in place
#define GT_A(i, j) A[(i) * ((i) + 1) / 2 + (j)]
#pragma unroll
    for (int i = 0; i < K; ++i) {
#pragma unroll
        for (int j = 0; j <= i; ++j) GT_A(i, j) = s0(j, i);
        vec[i] = s0(i, IY);
    }
    const double yy = s0(IY, IY), sy = s0(IY, I1);
    const double *Ssnp = S + (long long)snp * D.NCP;
    if constexpr (NI == 0) {
        // g column: g'q (K), g'y, g'1 = sum g, g'g
        double s1y = Ssnp[K], gg = sum_g2_all;
        if (flip) {      // g -> 2 - g on the complete cases (analysis.py:156-157)
            const double n_s = s0(I1, I1);
            const double ggf = 4.0 * n_s - 4.0 * sum_g_all + sum_g2_all;
            const bool all_zero = ggf == 0.0;
#pragma unroll
            for (int j = 0; j < K; ++j) GT_A(K, j) = all_zero ? 0.0 : 2.0 * s0(j, I1) - Ssnp[j];
            s1y = all_zero ? 0.0 : 2.0 * sy - s1y;
            gg = all_zero ? 0.0 : ggf;
        } else {
#pragma unroll
            for (int j = 0; j < K; ++j) GT_A(K, j) = Ssnp[j];
        }
        GT_A(K, K) = gg;
        vec[K] = s1y;
    } else {
        // rows g and g w: S1[c][b] = sum g m_c v_b (S columns c L + b), S2 = sum g^2 m_a m_b
        // (S columns NCA + {00, 01, 11}); m_0 = 1 (column I1 of v), m_1 = w (column IW)
        double s2_00 = Ssnp[D.NCA], s2_01 = Ssnp[D.NCA + 1], s2_11 = Ssnp[D.NCA + 2];
        bool all_zero = false;
        if (flip) {
            const double f00 = 4.0 * s0(I1, I1) - 4.0 * Ssnp[I1] + s2_00;
            const double f01 = 4.0 * s0(I1, IW) - 4.0 * Ssnp[IW] + s2_01;
            const double f11 = 4.0 * s0(IW, IW) - 4.0 * Ssnp[L + IW] + s2_11;
            all_zero = f00 == 0.0;
            s2_00 = all_zero ? 0.0 : f00; s2_01 = all_zero ? 0.0 : f01; s2_11 = all_zero ? 0.0 : f11;
        }
        auto s1 = [&](int c, int b) -> double {       // after the flip
            const double raw = Ssnp[c * L + b];
            if (!flip) return raw;
            const int cc = c == 0 ? I1 : IW;
            const double full = cc <= b ? s0(cc, b) : s0(b, cc);
            return all_zero ? 0.0 : 2.0 * full - raw;
        };
#pragma unroll
        for (int j = 0; j < K; ++j) { GT_A(K, j) = s1(0, j); GT_A(K + 1, j) = s1(1, j); }
        GT_A(K, K) = s2_00; GT_A(K + 1, K) = s2_01; GT_A(K + 1, K + 1) = s2_11;
        vec[K] = s1(0, IY); vec[K + 1] = s1(1, IY);
    }
    const double gdiag = NI == 0 ? GT_A(K, K) : GT_A(K, K) + GT_A(K + 1, K + 1);
    // ---- Cholesky A = Lc Lc' (in place) ----
    bool ok = !D.degenerate;
#pragma unroll
    for (int j = 0; j < P; ++j) {
        double d = GT_A(j, j);
#pragma unroll
        for (int c = 0; c < j; ++c) d = fma(-GT_A(j, c), GT_A(j, c), d);
        if (!(d > 0.0)) ok = false;
        const double sj = sqrt(d), inv = 1.0 / sj;
        GT_A(j, j) = sj;
#pragma unroll
        for (int i = j + 1; i < P; ++i) {
            double v = GT_A(i, j);
#pragma unroll
            for (int c = 0; c < j; ++c) v = fma(-GT_A(i, c), GT_A(j, c), v);
            GT_A(i, j) = v * inv;
        }
    }
    if (!ok) { iout[(long long)GT_I_STATUS * ostride + snp] = GT_SOLVE_DEFER; return; }
    // ---- Z = Lc^-1 (in place, column by column from the right) ----
#pragma unroll
    for (int c = P - 1; c >= 0; --c) {
        const double dinv = 1.0 / GT_A(c, c);
        // column c of Z below the diagonal needs columns > c of Z (already done) and column c of Lc
        double col[P];
#pragma unroll
        for (int i = c + 1; i < P; ++i) col[i] = GT_A(i, c);
#pragma unroll
        for (int i = P - 1; i > c; --i) {
            // Z[i][c] = - sum_{j = c .. i-1} Z[i][j'] ... computed as -(Z[i][c+1..i] . Lc[c+1..i][c]) / Lc[c][c]
            double s = 0.0;
#pragma unroll
            for (int j = c + 1; j <= i; ++j) s = fma(GT_A(i, j), col[j], s);
            GT_A(i, c) = -s * dinv;
        }
        GT_A(c, c) = dinv;
    }
    // u = Z rhs (descending rows, in place), ssr = yy - u'u
    double uu = 0.0;
#pragma unroll
    for (int r = P - 1; r >= 0; --r) {
        double s = 0.0;
#pragma unroll
        for (int a = 0; a <= r; ++a) s = fma(GT_A(r, a), vec[a], s);
        vec[r] = s;
        uu = fma(s, s, uu);
    }
    const double ssr = yy - uu;
    // bq = Z' u (ascending, in place)
#pragma unroll
    for (int i = 0; i < P; ++i) {
        double s = 0.0;
#pragma unroll
        for (int r = i; r < P; ++r) s = fma(GT_A(r, i), vec[r], s);
        vec[i] = s;
    }
    // back in the reference's parameterisation: beta = T bq, var_i = |column i of Z T'|^2,
    // T = blockdiag(R^-1, 1)
    double var[P];
    double tv = 0.0;
#pragma unroll
    for (int i = 0; i < P; ++i) {
        double v = 0.0;
#pragma unroll
        for (int r = i; r < P; ++r) {
            double w;
            if (i < K) {
                w = 0.0;
#pragma unroll
                for (int a = i; a <= (r < K - 1 ? r : K - 1); ++a) w = fma(GT_A(r, a), sRinv[i * K + a], w);
            } else {
                w = GT_A(r, i);
            }
            v = fma(w, w, v);
        }
        var[i] = v;
        tv += v;
    }
#pragma unroll
    for (int i = 0; i < K; ++i) {
        double b = 0.0;
#pragma unroll
        for (int a = i; a < K; ++a) b = fma(sRinv[i * K + a], vec[a], b);
        vec[i] = b;
    }
#undef GT_A
    // certified gates: cond^2 <= trace(G) trace(G^-1), lambda_min >= 1 / trace(G^-1)
    const double tg = gdiag + D.trG0;
    const bool cond_safe = tg * tv <= D.cond_t * D.cond_t * (1.0 - 1e-9);
    const bool eig_safe = 1.0 / tv >= D.eig_t * (1.0 + 1e-9);
    if (!(cond_safe && eig_safe)) { iout[(long long)GT_I_STATUS * ostride + snp] = GT_SOLVE_DEFER; return; }

    const int df = nj - P;
    const double scale = (df > 0) ? ssr / (double)df : nan("");
    const double q975 = (df > 0) ? D.tq[df] : nan("");
    double pv_snp = 0.0;
#pragma unroll 1
    for (int j = 0; j < P; ++j) {
        // (dynamic index into vec / var: a small local-memory array at this point)
        const double coef = vec[j] + ((j == D.icpt_col) ? D.icpt_shift : 0.0);
        const double se = sqrt(scale * var[j]);
        const double tval = coef / se;
        const double pval = (df > 0) ? gt_t_sf2(tval, (double)df) : nan("");
        double *o = out + (long long)(GT_F_PARAM0 + 6 * j) * ostride + snp;
        o[0] = coef; o[ostride] = se; o[2 * ostride] = coef - q975 * se; o[3 * ostride] = coef + q975 * se;
        o[4 * ostride] = tval; o[5 * ostride] = pval;
        if (j == K) pv_snp = pval;
    }
    int status = GT_OK;
    if (!D.raw_mode && pv_snp != pv_snp) status = GT_NAN_PVALUE;
    const double dn = (double)nj;
    const double tss = D.has_const ? (yy - sy * sy / dn) : yy;
    const double r2 = 1.0 - ssr / tss;
    const double r2a = 1.0 - ((dn - (double)D.has_const) / (double)df) * (1.0 - r2);
    const double llf = -0.5 * dn * log(2.0 * 3.14159265358979323846) - 0.5 * dn * log(ssr / dn) - 0.5 * dn;
    out[(long long)GT_F_MAF * ostride + snp] = maf;
    out[(long long)GT_F_AUX * ostride + snp] = nan("");
    out[(long long)GT_F_STAT * ostride + snp] = r2a;
    out[(long long)GT_F_LLF * ostride + snp] = llf;
    iout[(long long)GT_I_STATUS * ostride + snp] = status;
    iout[(long long)GT_I_NOBS * ostride + snp] = nj;
    iout[(long long)GT_I_FLIP * ostride + snp] = flip;
    iout[(long long)GT_I_NITER * ostride + snp] = 0;
}

}  // namespace gtk
