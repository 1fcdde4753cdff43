// K3b: per-SNP normal-equation solve of the linear model, GW lanes per SNP
// (two SNPs per warp when the design has at most 16 columns).  Restates what
// sm.OLS(y, X).fit() + StatsLinear._results_handler
// (genetest/statistics/models/linear.py:46-104) compute for the complete
// cases of each marker (analysis.py:119,172), after _gwas_worker's MAF /
// flip / interaction steps (analysis.py:127-166):
//   * complete-case Gram matrix = full-sample FF0 - the corrections K3a
//     accumulated over the marker's missing samples,
//   * MAF on the complete cases, flip (g -> 2 - g) applied algebraically,
//   * p x p normal equations in the orthonormal covariate basis (Cholesky),
//     mapped back through R^-1 to the reference's parameterisation,
//   * condition-number / eigenvalue gates (certified trace bound first,
//     Jacobi eigenvalues only when the bound cannot decide),
//   * t statistics, two-sided p-values, 95 % CI, adjusted R^2, log-lik.
// The two groups of a warp run independently: every shuffle / barrier names
// the group's own lanes.
#pragma once

#include "gt_common.cuh"

namespace gtk {

struct SolveLayout {
    int perGroup;   // doubles per SNP
    int oS0, oMo, oMc, oMi, oScratch, oVec;
};

__host__ __device__ inline SolveLayout solve_layout(int L, int p, int I) {
    SolveLayout f;
    int o = 0;
    f.oS0 = o; o += L * L > p * p ? L * L : p * p;     // also the Jacobi work matrix
    f.oMo = o; o += p * p;
    f.oMc = o; o += p * p;
    f.oMi = o; o += p * p;
    f.oScratch = o; o += p * p;
    f.oVec = o; o += 6 * p + (I + 1) * L + (I + 1) * (I + 1) + 8;
    f.perGroup = (o + 1) & ~1;
    return f;
}

template <int GW>
__device__ __forceinline__ double gt_group_sum(unsigned mask, double v) {
#pragma unroll
    for (int o = GW / 2; o > 0; o >>= 1) v += __shfl_xor_sync(mask, v, o);
    return v;
}
template <int GW>
__device__ __forceinline__ double gt_group_max(unsigned mask, double v) {
#pragma unroll
    for (int o = GW / 2; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(mask, v, o));
    return v;
}
template <int GW>
__device__ __forceinline__ double gt_group_min(unsigned mask, double v) {
#pragma unroll
    for (int o = GW / 2; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(mask, v, o));
    return v;
}

// In-place lower Cholesky of the p x p matrix A (shared memory), by the GW
// lanes of a group.  Returns false when a pivot is not positive.
template <int GW>
__device__ __forceinline__ bool gt_group_cholesky(double *A, int ld, int p, int gl, unsigned mask) {
    bool ok = true;
    for (int j = 0; j < p; ++j) {
        double d = A[j * ld + j];
        if (!(d > 0.0)) { ok = false; break; }
        double s = sqrt(d);
        double inv = 1.0 / s;
        __syncwarp(mask);
        for (int i = j + 1 + gl; i < p; i += GW) A[i * ld + j] *= inv;
        if (gl == 0) A[j * ld + j] = s;
        __syncwarp(mask);
        // trailing update, one row per lane (no index division)
        for (int i = j + 1 + gl; i < p; i += GW) {
            const double aij = A[i * ld + j];
            for (int c = j + 1; c <= i; ++c) A[i * ld + c] = fma(-aij, A[c * ld + j], A[i * ld + c]);
        }
        __syncwarp(mask);
    }
    return ok;
}

// Cyclic Jacobi eigenvalues of the symmetric p x p matrix G (shared memory,
// destroyed).  A zero row/column stays exactly zero (monomorphic marker ->
// eigenvalue 0 -> inf).
template <int GW>
__device__ __forceinline__ void gt_group_jacobi_extremes(double *G, int ld, int p, int gl, unsigned mask,
                                                         double *emax, double *emin) {
    for (int sweep = 0; sweep < 40; ++sweep) {
        double off = 0.0, dg = 0.0;
        for (int t = gl; t < p * p; t += GW) {
            int i = t / p, j = t % p;
            double v = G[i * ld + j];
            if (i == j) dg += v * v; else off += v * v;
        }
        off = gt_group_sum<GW>(mask, off);
        dg = gt_group_sum<GW>(mask, dg);
        if (off <= 1e-32 * dg || off == 0.0) break;
        for (int r = 0; r < p - 1; ++r) {
            for (int c = r + 1; c < p; ++c) {
                double apq = G[r * ld + c];
                if (apq == 0.0) continue;
                double app = G[r * ld + r], aqq = G[c * ld + c];
                double tau = (aqq - app) / (2.0 * apq);
                double t = (tau >= 0.0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
                double cs = 1.0 / sqrt(1.0 + t * t), sn = t * cs;
                __syncwarp(mask);
                for (int i = gl; i < p; i += GW) {
                    if (i != r && i != c) {
                        double gir = G[i * ld + r], gic = G[i * ld + c];
                        double nr = cs * gir - sn * gic, nc = sn * gir + cs * gic;
                        G[i * ld + r] = nr; G[r * ld + i] = nr;
                        G[i * ld + c] = nc; G[c * ld + i] = nc;
                    }
                }
                if (gl == 0) {
                    G[r * ld + r] = app - t * apq;
                    G[c * ld + c] = aqq + t * apq;
                    G[r * ld + c] = 0.0;
                    G[c * ld + r] = 0.0;
                }
                __syncwarp(mask);
            }
        }
    }
    double mx = -INFINITY, mn = INFINITY;
    for (int i = gl; i < p; i += GW) {
        double v = G[i * ld + i];
        mx = fmax(mx, v); mn = fmin(mn, v);
    }
    *emax = gt_group_max<GW>(mask, mx);
    *emin = gt_group_min<GW>(mask, mn);
}

template <int GW>
__device__ __forceinline__ void solve_write_fail(double *out, int *iout, long long n_snps, int snp,
                                                 int nf, int status, int nobs, int flip,
                                                 double maf, double aux, int gl) {
    for (int f = gl; f < nf; f += GW) {
        double v = nan("");
        if (f == GT_F_MAF) v = maf;
        if (f == GT_F_AUX) v = aux;
        out[(long long)f * n_snps + snp] = v;
    }
    if (gl == 0) {
        iout[(long long)GT_I_STATUS * n_snps + snp] = status;
        iout[(long long)GT_I_NOBS * n_snps + snp] = nobs;
        iout[(long long)GT_I_FLIP * n_snps + snp] = flip;
        iout[(long long)GT_I_NITER * n_snps + snp] = 0;
    }
}

// corrT[entry][snp] = upper triangle of the sum over the missing samples of v v'
// (K3a), read only for markers with missing genotypes.
template <int GW>
__global__ void __launch_bounds__(256)
linear_solve_kernel(GtDesignDev D, int n_snps, const double *__restrict__ S,
                    const int *__restrict__ counts, const double *__restrict__ gsum,
                    const double *__restrict__ corrT, long long ldc, double *__restrict__ out,
                    int *__restrict__ iout, long long ostride, int only_deferred) {
    extern __shared__ __align__(16) double sm[];
    const int lane = threadIdx.x & 31;
    const int gl = lane % GW;
    const unsigned gmask = GW == 32 ? 0xFFFFFFFFu : (((1u << GW) - 1u) << (lane - gl));
    const int group = threadIdx.x / GW;
    const int snp = blockIdx.x * (blockDim.x / GW) + group;
    if (snp >= n_snps) return;
    // second pass behind linear_solve_thread_kernel: only the markers it left to the Jacobi path
    if (only_deferred && iout[(long long)GT_I_STATUS * ostride + snp] != -77) return;
    const int k = D.k, I = D.I, L = D.L, p = D.p;
    const int iy = k, i1 = k + 1;
    const int nf = GT_F_PARAM0 + 6 * p;
    const SolveLayout lay = solve_layout(L, p, I);

    double *W = sm + (size_t)group * lay.perGroup;
    double *S0 = W + lay.oS0, *Mo = W + lay.oMo, *Mc = W + lay.oMc, *Mi = W + lay.oMi;
    double *scratch = W + lay.oScratch;
    double *rhs = W + lay.oVec, *bq = rhs + p, *beta = bq + p, *var = beta + p,
           *gdiag = var + p, *tmpv = gdiag + p;
    double *S1 = tmpv + p;                  // [(I+1)][L]
    double *S2 = S1 + (I + 1) * L;          // [(I+1)][(I+1)]

    const int n_miss = counts[(long long)snp * 4 + 2];
    if (n_miss > 0) {
        // upper triangle, SoA over the markers (the 16 markers of a CTA share the sectors)
        const double *cs = corrT + snp;
        for (int t = gl; t < L * L; t += GW) {
            const int a = t / L, b = t % L;
            const int lo = a < b ? a : b, hi = a < b ? b : a;
            S0[t] = D.FF0[t] - __ldg(cs + (long long)(lo * L - lo * (lo - 1) / 2 + (hi - lo)) * ldc);
        }
    } else {
        for (int t = gl; t < L * L; t += GW) S0[t] = D.FF0[t];
    }
    __syncwarp(gmask);

    const double *Ssnp = S + (long long)snp * D.NCP;
    const double sum_g_all = gsum[(long long)snp * 2], sum_g2_all = gsum[(long long)snp * 2 + 1];
    if (D.lean) {      // E carries [q | y] only; sum g, sum g^2 come from the counts
        for (int t = gl; t < L; t += GW) S1[t] = (t <= k) ? Ssnp[t] : sum_g_all;
        if (gl == 0) S2[0] = sum_g2_all;
    } else {
        for (int t = gl; t < (I + 1) * L; t += GW) S1[t] = Ssnp[t];
        for (int t = gl; t < (I + 1) * (I + 1); t += GW) {
            int a = t / (I + 1), b = t % (I + 1);
            int lo = a < b ? a : b, hi = a < b ? b : a;
            int idx = lo * (I + 1) - lo * (lo - 1) / 2 + (hi - lo);
            S2[t] = Ssnp[D.NCA + idx];
        }
    }
    __syncwarp(gmask);

    const int nj = D.n - n_miss;
    if (nj <= 0) {
        solve_write_fail<GW>(out, iout, ostride, snp, nf, GT_ALL_MISSING, 0, 0, nan(""), nan(""), gl);
        return;
    }
    // MAF on the complete cases (descriptive.py:43): nanmean(g) / 2
    double maf = nan("");
    int flip = 0;
    if (!D.raw_mode) {
        double sum_g = sum_g_all;
        maf = (sum_g / (double)nj) / 2.0;
        if (maf > 0.5) { maf = 1.0 - maf; flip = 1; }
        if (D.maf_o != nullptr) { maf = D.maf_o[snp]; flip = D.flip_o[snp]; }
        if (D.use_maf_t && maf < D.maf_t) {
            solve_write_fail<GW>(out, iout, ostride, snp, nf, GT_MAF_BELOW, nj, flip, maf, maf, gl);
            return;
        }
    }
    if (flip) {   // g -> 2 - g on the complete cases (analysis.py:156-157)
        for (int t = gl; t < (I + 1) * (I + 1); t += GW) {
            int a = t / (I + 1), b = t % (I + 1);
            int ca = a == 0 ? i1 : k + 1 + a, cb = b == 0 ? i1 : k + 1 + b;
            scratch[t] = 4.0 * S0[ca * L + cb] - 4.0 * S1[a * L + cb] + S2[t];
        }
        __syncwarp(gmask);
        bool all_zero = (scratch[0] == 0.0);
        for (int t = gl; t < (I + 1) * L; t += GW) {
            int c = t / L, b = t % L;
            int cc = c == 0 ? i1 : k + 1 + c;
            S1[t] = all_zero ? 0.0 : 2.0 * S0[cc * L + b] - S1[t];
        }
        __syncwarp(gmask);
        for (int t = gl; t < (I + 1) * (I + 1); t += GW) S2[t] = all_zero ? 0.0 : scratch[t];
        __syncwarp(gmask);
    }

    // normal equations in the [Q | g | g w] basis
    for (int j = gl; j < p; j += GW)                     // one column per lane
        for (int i = 0; i < p; ++i) {
            double v;
            if (i < k && j < k) v = S0[i * L + j];
            else if (i < k) v = S1[(j - k) * L + i];
            else if (j < k) v = S1[(i - k) * L + j];
            else v = S2[(i - k) * (I + 1) + (j - k)];
            Mo[i * p + j] = v; Mc[i * p + j] = v;
        }
    for (int i = gl; i < p; i += GW) rhs[i] = (i < k) ? S0[i * L + iy] : S1[(i - k) * L + iy];
    const double yy = S0[iy * L + iy], sy = S0[iy * L + i1];
    __syncwarp(gmask);

    bool chol_ok = !D.degenerate && gt_group_cholesky<GW>(Mc, p, p, gl, gmask);
    bool need_jacobi = !chol_ok;
    double ssr = 0.0;
    if (chol_ok) {
        // M = Lc Lc'.  Only Z = Lc^-1 is formed (one forward substitution per
        // column): with u = Z rhs, bq = Z' u and ssr = yy - u'u; back in the
        // reference's parameterisation beta = T bq and var_i = |column i of
        // Z T'|^2 with T = blockdiag(R^-1, I).
        double *Z = Mi, *invd = tmpv, *u = gdiag;
        for (int i = gl; i < p; i += GW) invd[i] = 1.0 / Mc[i * p + i];
        for (int t = gl; t < p * p; t += GW) Z[t] = 0.0;
        __syncwarp(gmask);
        for (int c = gl; c < p; c += GW)
            for (int i = c; i < p; ++i) {
                double s = (i == c) ? 1.0 : 0.0;
                for (int j = c; j < i; ++j) s = fma(-Mc[i * p + j], Z[j * p + c], s);
                Z[i * p + c] = s * invd[i];
            }
        __syncwarp(gmask);
        double part = 0.0;
        for (int r = gl; r < p; r += GW) {
            double s = 0.0;
            for (int a = 0; a <= r; ++a) s = fma(Z[r * p + a], rhs[a], s);
            u[r] = s;
            part = fma(s, s, part);
        }
        ssr = yy - gt_group_sum<GW>(gmask, part);
        __syncwarp(gmask);
        for (int i = gl; i < p; i += GW) {
            double s = 0.0;
            for (int r = i; r < p; ++r) s = fma(Z[r * p + i], u[r], s);
            bq[i] = s;
        }
        __syncwarp(gmask);
        // var_i = sum over r of W[r][i]^2, W = Z T': one column per lane
        double tv = 0.0, tg = 0.0;
        for (int i = gl; i < p; i += GW) {
            double v = 0.0, b;
            for (int r = i; r < p; ++r) {
                double w;
                if (i < k) {
                    w = 0.0;
                    const int hi = r < k - 1 ? r : k - 1;
                    for (int a = i; a <= hi; ++a) w = fma(Z[r * p + a], D.Rinv[i * k + a], w);
                } else {
                    w = Z[r * p + i];
                }
                v = fma(w, w, v);
            }
            if (i < k) {
                b = 0.0;
                for (int a = i; a < k; ++a) b = fma(D.Rinv[i * k + a], bq[a], b);
            } else {
                b = bq[i];
                tg += Mo[i * p + i];
            }
            beta[i] = b; var[i] = v;
            tv += v;
        }
        __syncwarp(gmask);
        tv = gt_group_sum<GW>(gmask, tv);
        // trace of the Gram matrix of the original columns: dropping samples
        // only lowers the covariates' diagonal, so |R|_F^2 bounds it
        tg = gt_group_sum<GW>(gmask, tg) + D.trG0;
        // cond^2 <= trace(G) trace(G^-1),  lambda_min >= 1 / trace(G^-1)
        bool cond_safe = tg * tv <= D.cond_t * D.cond_t * (1.0 - 1e-9);
        bool eig_safe = 1.0 / tv >= D.eig_t * (1.0 + 1e-9);
        need_jacobi = !(cond_safe && eig_safe);
    }
    if (need_jacobi) {
        // G = Ti' Mo Ti with Ti = blockdiag(R, I): Gram matrix of the original columns
        double *tmp = scratch, *G = S0;
        for (int t = gl; t < p * p; t += GW) {
            int i = t / p, j = t % p;
            double s;
            if (j < k) { s = 0.0; for (int a = 0; a <= j; ++a) s = fma(Mo[i * p + a], D.R[a * k + j], s); }
            else s = Mo[i * p + j];
            tmp[t] = s;
        }
        __syncwarp(gmask);
        for (int t = gl; t < p * p; t += GW) {
            int i = t / p, j = t % p;
            double s;
            if (i < k) { s = 0.0; for (int a = 0; a <= i; ++a) s = fma(D.R[a * k + i], tmp[a * p + j], s); }
            else s = tmp[i * p + j];
            G[i * p + j] = s;
        }
        __syncwarp(gmask);
        // symmetrise (rounding) so a zero column stays exactly zero on both sides
        for (int t = gl; t < p * p; t += GW) {
            int i = t / p, j = t % p;
            if (i < j) { double v = G[j * p + i]; G[i * p + j] = v; }
        }
        __syncwarp(gmask);
        double emax, emin;
        gt_group_jacobi_extremes<GW>(G, p, p, gl, gmask, &emax, &emin);
        // statsmodels: condition_number = sqrt(max / min) of eigvalsh(X'X): a zero minimum
        // gives inf, a (rounding) negative one NaN -- and NaN passes the `> threshold` test,
        // so that marker fails on the eigenvalue check below instead (linear.py:51-61)
        double cond = (emin > 0.0) ? sqrt(emax / emin) : (emin == 0.0 ? INFINITY : nan(""));
        if (D.degenerate) cond = INFINITY;
        if (cond > D.cond_t) {
            solve_write_fail<GW>(out, iout, ostride, snp, nf, GT_COND_LARGE, nj, flip, maf, cond, gl);
            return;
        }
        if (emin < D.eig_t) {
            solve_write_fail<GW>(out, iout, ostride, snp, nf, GT_EIG_SMALL, nj, flip, maf, emin, gl);
            return;
        }
        if (!chol_ok) {
            solve_write_fail<GW>(out, iout, ostride, snp, nf, GT_COND_LARGE, nj, flip, maf, INFINITY, gl);
            return;
        }
    }

    const int df = nj - p;
    const double scale = (df > 0) ? ssr / (double)df : nan("");
    const double q975 = (df > 0) ? D.tq[df] : nan("");
    double pv_snp = 0.0;
    double my[6];
    for (int j = gl; j < p; j += GW) {
        double coef = beta[j] + ((j == D.icpt_col) ? D.icpt_shift : 0.0);
        double se = sqrt(scale * var[j]);
        double tval = coef / se;
        double pval = (df > 0) ? gt_t_sf2(tval, (double)df) : nan("");
        my[0] = coef; my[1] = se; my[2] = coef - q975 * se; my[3] = coef + q975 * se;
        my[4] = tval; my[5] = pval;
#pragma unroll
        for (int f = 0; f < 6; ++f)
            out[(long long)(GT_F_PARAM0 + 6 * j + f) * ostride + snp] = my[f];
        if (j == k) pv_snp = pval;
    }
    pv_snp = __shfl_sync(gmask, pv_snp, k % GW, GW);
    int status = GT_OK;
    if (!D.raw_mode && pv_snp != pv_snp) status = GT_NAN_PVALUE;
    if (gl == 0) {
        double dn = (double)nj;
        double tss = D.has_const ? (yy - sy * sy / dn) : yy;
        double r2 = 1.0 - ssr / tss;
        double r2a = 1.0 - ((dn - (double)D.has_const) / (double)df) * (1.0 - r2);
        double llf = -0.5 * dn * log(2.0 * 3.14159265358979323846) -
                     0.5 * dn * log(ssr / dn) - 0.5 * dn;
        out[(long long)GT_F_MAF * ostride + snp] = maf;
        out[(long long)GT_F_AUX * ostride + snp] = nan("");
        out[(long long)GT_F_STAT * ostride + snp] = r2a;
        out[(long long)GT_F_LLF * ostride + snp] = llf;
        iout[(long long)GT_I_STATUS * ostride + snp] = status;
        iout[(long long)GT_I_NOBS * ostride + snp] = nj;
        iout[(long long)GT_I_FLIP * ostride + snp] = flip;
        iout[(long long)GT_I_NITER * ostride + snp] = need_jacobi ? 1 : 0;
    }
}

}  // namespace gtk
