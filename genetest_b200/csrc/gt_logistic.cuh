// K4: batched warp-per-SNP IRLS for the logistic model, FP64.
//
// Restates sm.GLM(y, X, family=Binomial()).fit() as called by
// genetest/statistics/models/logistic.py:38-40 -- start mu = (y + .5) / 2,
// one weighted least-squares solve per iteration, perfect-separation check
// allclose(mu - y, 0), stop on abs(delta deviance) <= 1e-8 (maxiter 100),
// params / covariance from the LAST step's working weights -- for every
// marker of a batch, after _gwas_worker's complete-case mask, MAF, flip and
// interaction columns (genetest/analysis.py:119-166).
//
// One warp owns one SNP.  Per 32 samples: (1) each lane builds its sample's
// design row [q | g | g w], eta, mu, working weight and response, (2) the
// warp accumulates X'W[X | z] with FP64 tensor-core MMAs (m8n8k4) out of a
// shared-memory staging tile; the p x p solve is a warp-level Cholesky.
#pragma once

#include <string>
#include <vector>

#include "gt_common.cuh"

namespace gtk {

struct IterDesignDev {
    int n_file, n, k, I, p;
    int XS;                 // row stride of X (doubles): k + I rounded up to even
    int n_pad;
    int raw_mode, use_maf_t;
    int degenerate;         // covariates alone are rank deficient
    int y_binary;           // every outcome of the design is 0 or 1 (logistic: thread-per-SNP kernel)
    double maf_t;
    const double *X;        // [n_pad][XS]  q (k) | w (I), genotype-file order, zeros when excluded
    const double *y;        // [n_pad]      outcome / event, NaN = sample not in the design
    const double *Rinv;     // [k*k]
    // coxph only: schedule in descending time, 32 slots per block, tie groups
    // never straddle a block (host pads with empty slots)
    int n_slots;
    const int *slot_pos;    // [n_slots] genotype position of the slot's sample, -1 = empty
    const int *slot_info;   // [n_slots] bits 0-4 first lane of the tie group, 5-9 last lane,
                            //           bit 10 event, bit 11 block has a tie group
    const double *Xs;       // [n_slots][XS] q | w in slot order
    const int *blk_info;    // [n_slots/32] 1, or the block count of a big tie group on its first block
    const double *colmax;   // [p] coxph: max |x| of every design column (q, g, g w)
    const double *maf_o;    // sex-aware MAF / flip per marker of the launch, or NULL
    const int *flip_o;
    int newton_maxiter;     // coxph: PHReg.fit(maxiter=100); engine option "cox_maxiter" (tests)
    int n_seg;              // coxph: strata (1 without), blocks seg_start[s] .. seg_start[s + 1) each
    const int *seg_start;   // [n_seg + 1]
    // coxph, thread-per-SNP kernel (no interaction): design rows in descending time, stratum by
    // stratum, each padded to whole 256-row blocks (padding rows: position -1, zeros)
    int th_ok, th_n, th_XS, th_tied, th_nseg;
    const double *th_X;     // [th_n][th_XS]  q (k) | flags (1 event, 2 tied, 4 first, 8 last of its group) | group size
    const int *th_pos;      // [th_n] genotype position
    const int *th_order;    // [2 th_n / 256] blocks in the order one round consumes them
    const int *th_seg;      // [th_nseg + 1] first block of every stratum
};

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// genotype of sample i of a row: returns false when missing
template <bool PACKED>
__device__ __forceinline__ bool load_geno(const void *rowp, int i, double &g) {
    if (PACKED) {
        uint32_t byte = ((const uint8_t *)rowp)[i >> 2];
        uint32_t code = (byte >> (2 * (i & 3))) & 3u;
        g = (code == 0) ? 2.0 : ((code == 2) ? 1.0 : 0.0);
        return code != 1;
    } else {
        g = ((const double *)rowp)[i];
        return g == g;
    }
}

template <int NT>
struct LogitCfg {
    static constexpr int PH = 8 * NT;                // padded Hessian dimension
    // staging tile, TRANSPOSED: row a (design column, then z, then the weights)
    // holds the 32 samples of the block; row stride 36 doubles keeps both the
    // per-lane stores and the MMA fragment loads bank-conflict free
    static constexpr int XT = 36;
    // per-warp doubles: xsT (PH + 1 rows), H, Hc, Hinv, beta, rhs, coef, var
    static constexpr int perWarp = (PH + 1) * XT + 3 * PH * PH + 4 * PH;
};

template <bool PACKED, int NT>
__global__ void __launch_bounds__(256)
logistic_irls_kernel(IterDesignDev T, const void *__restrict__ geno, long long pitch, int n_snps,
                     double *__restrict__ out, int *__restrict__ iout, long long ostride, int only_deferred) {
    using Cfg = LogitCfg<NT>;
    constexpr int PH = Cfg::PH, XT = Cfg::XT;
    extern __shared__ __align__(16) double sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int snp = blockIdx.x * (blockDim.x >> 5) + warp;
    if (snp >= n_snps) return;
    // second pass behind logistic_thread_kernel: only the markers it left unfinished
    if (only_deferred && iout[(long long)GT_I_STATUS * ostride + snp] != -77) return;
    const int k = T.k, I = T.I, p = T.p;
    const int nf = GT_F_PARAM0 + 6 * p;
    const double EPS = 2.220446049250313e-16;

    double *W = sm + (size_t)warp * Cfg::perWarp;
    double *xsT = W, *ws = xsT + PH * XT, *H = xsT + (PH + 1) * XT, *Hc = H + PH * PH,
           *Hinv = Hc + PH * PH, *beta = Hinv + PH * PH, *rhs = beta + PH, *coef = rhs + PH,
           *var = coef + PH;

    const void *rowp = PACKED ? (const void *)((const uint8_t *)geno + (long long)snp * pitch)
                              : (const void *)((const double *)geno + (long long)snp * pitch);
    const int nblk = (T.n_file + 31) / 32;

    // ---- complete cases, MAF, flip (analysis.py:119-157) ----
    int nj = 0;
    double sum_g = 0.0, nev = 0.0;
    for (int b = 0; b < nblk; ++b) {
        int i = b * 32 + lane;
        double yv = (i < T.n_file) ? T.y[i] : nan("");
        double g = 0.0;
        bool ok = (yv == yv);
        if (ok) ok = load_geno<PACKED>(rowp, i, g);
        if (ok) { nj += 1; sum_g += g; nev += yv; }
    }
    nj = gt_warp_sum_int(nj);
    sum_g = gt_warp_sum(sum_g);
    nev = gt_warp_sum(nev);
    auto write_fail = [&](int status, int flip, double maf, double aux, int niter) {
        for (int f = lane; f < nf; f += 32) {
            double v = nan("");
            if (f == GT_F_MAF) v = maf;
            if (f == GT_F_AUX) v = aux;
            out[(long long)f * ostride + snp] = v;
        }
        if (lane == 0) {
            iout[(long long)GT_I_STATUS * ostride + snp] = status;
            iout[(long long)GT_I_NOBS * ostride + snp] = nj;
            iout[(long long)GT_I_FLIP * ostride + snp] = flip;
            iout[(long long)GT_I_NITER * ostride + snp] = niter;
        }
    };
    if (nj == 0) { write_fail(GT_ALL_MISSING, 0, nan(""), nan(""), 0); return; }
    double maf = nan("");
    int flip = 0;
    if (!T.raw_mode) {
        maf = (sum_g / (double)nj) / 2.0;
        if (maf > 0.5) { maf = 1.0 - maf; flip = 1; }
        if (T.maf_o != nullptr) { maf = T.maf_o[snp]; flip = T.flip_o[snp]; }
        if (T.use_maf_t && maf < T.maf_t) { write_fail(GT_MAF_BELOW, flip, maf, maf, 0); return; }
    }

    // rank-deficient covariates: statsmodels' pinv gives a zero standard error
    // for the redundant direction -> NaN inference
    if (T.degenerate) { write_fail(GT_NAN_PVALUE, flip, maf, nan(""), 0); return; }

    for (int t = lane; t < (PH + 1) * XT; t += 32) xsT[t] = 0.0;   // padding rows stay 0
    for (int t = lane; t < PH; t += 32) beta[t] = 0.0;
    __syncwarp();

    // pass-0 values of a binary outcome, by the formulas of the general branch
    const double mu1 = (1.0 + 0.5) / 2.0, mu0 = (0.0 + 0.5) / 2.0;
    const double eta1 = log(mu1 / (1.0 - mu1)), eta0 = log(mu0 / (1.0 - mu0));
    const double dev1 = 2.0 * (1.0 * log(fmax(1.0 / (mu1 + 1e-20), EPS)) + 0.0 * log(fmax(0.0 / (1.0 - mu1 + 1e-20), EPS)));
    const double dev0 = 2.0 * (0.0 * log(fmax(0.0 / (mu0 + 1e-20), EPS)) + 1.0 * log(fmax(1.0 / (1.0 - mu0 + 1e-20), EPS)));
    double dev_prev = 0.0, llf = 0.0;
    int niter = 0;
    int status = GT_OK;
    bool done = false;
    // pass 0 uses the statsmodels start (no beta); pass t >= 1 uses beta_t
    for (int pass = 0; pass <= 100 && !done; ++pass) {
        double acc[NT][NT][2];
#pragma unroll
        for (int a = 0; a < NT; ++a)
#pragma unroll
            for (int b = 0; b < NT; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;
        double dev = 0.0, ll = 0.0;
        int sep = 1;
        // The outcome, genotype and design row of block b + 1 are loaded into
        // registers while block b is being worked on (X and y have n_pad >=
        // 32 nblk rows; rows of excluded samples are zero / NaN).
        double ny, nx[PH], ngd = 0.0;
        uint32_t nbyte = 0;
        auto prefetch = [&](int b) {
            const int i = b * 32 + lane;
            ny = T.y[i];
            if (PACKED) nbyte = (i < T.n_file) ? ((const uint8_t *)rowp)[i >> 2] : 0u;   // caller's pitch may be tight
            else ngd = (i < T.n_file) ? ((const double *)rowp)[i] : nan("");
            const double *xi = T.X + (long long)i * T.XS;
#pragma unroll
            for (int a = 0; a < PH; ++a)
                if (a < k + I) nx[a] = xi[a];
        };
        prefetch(0);
        for (int b = 0; b < nblk; ++b) {
            const int i = b * 32 + lane;
            const double yv = (i < T.n_file) ? ny : nan("");
            double g = 0.0;
            bool ok = (yv == yv);
            if (PACKED) {
                const uint32_t code = (nbyte >> (2 * (i & 3))) & 3u;
                g = (code == 0) ? 2.0 : ((code == 2) ? 1.0 : 0.0);
                ok = ok && code != 1;
            } else {
                g = ngd;
                ok = ok && g == g;
            }
            if (!ok) g = 0.0;
            else if (flip) g = 2.0 - g;
            // design row [q | g | g w] into the transposed staging tile, eta = x'beta
            double eta = g * beta[k];
            xsT[k * XT + lane] = g;
#pragma unroll
            for (int a = 0; a < PH; ++a) {
                if (a < k) {
                    const double v = nx[a];
                    xsT[a * XT + lane] = v; eta = fma(v, beta[a], eta);
                } else if (a < k + I) {
                    const double v = g * nx[a];
                    xsT[(a + 1) * XT + lane] = v; eta = fma(v, beta[a + 1], eta);
                }
            }
            if (b + 1 < nblk) prefetch(b + 1);
            double w = 0.0, z = 0.0;
            if (ok) {
                double mu;
                const bool binary = (yv == 0.0) || (yv == 1.0);
                if (pass == 0 && binary) {
                    // statsmodels' start for y in {0, 1}: the same constants for every sample
                    mu = yv != 0.0 ? mu1 : mu0;
                    eta = yv != 0.0 ? eta1 : eta0;
                    dev += yv != 0.0 ? dev1 : dev0;
                } else if (pass == 0) {
                    mu = (yv + 0.5) / 2.0;
                    eta = log(mu / (1.0 - mu));
                    double e1 = fmax(yv / (mu + 1e-20), EPS), e0 = fmax((1.0 - yv) / (1.0 - mu + 1e-20), EPS);
                    dev += 2.0 * (yv * log(e1) + (1.0 - yv) * log(e0));
                } else {
                    // mu = 1 / (1 + exp(-eta)); log(1 - mu) = -(max(eta, 0) + log1p(exp(-|eta|)))
                    double t = exp(-fabs(eta));
                    double inv = 1.0 / (1.0 + t);
                    mu = (eta >= 0.0) ? inv : t * inv;
                    if (binary) {
                        // y in {0, 1}: loglike_i = y eta + log(1 - mu), deviance_i = -2 loglike_i
                        // log(1 + t), t in (0, 1]: absolute error <= 1e-16 per sample
                        double li = yv * eta - (fmax(eta, 0.0) + log(1.0 + t));
                        ll += li;
                        dev -= 2.0 * li;
                    } else {
                        double e1 = fmax(yv / (mu + 1e-20), EPS), e0 = fmax((1.0 - yv) / (1.0 - mu + 1e-20), EPS);
                        dev += 2.0 * (yv * log(e1) + (1.0 - yv) * log(e0));
                        ll += yv * log(mu / (1.0 - mu) + 1e-20) + log(1.0 - mu + 1e-20);
                    }
                    if (!(fabs(mu - yv) <= 1e-8)) sep = 0;
                }
                double pc = fmin(fmax(mu, EPS), 1.0 - EPS);
                w = pc * (1.0 - pc);                 // 1 / (link.deriv^2 * variance)
                z = eta + (yv - mu) / w;             // working response
            }
            xsT[p * XT + lane] = z;
            ws[lane] = w;
            __syncwarp();
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) {
                const int e = ks * 4 + (lane & 3), c = lane >> 2;
                const double wv = ws[e];
                double af[NT], bf[NT];
#pragma unroll
                for (int t = 0; t < NT; ++t) { bf[t] = xsT[(t * 8 + c) * XT + e]; af[t] = wv * bf[t]; }
#pragma unroll
                for (int mt = 0; mt < NT; ++mt)
#pragma unroll
                    for (int nt = mt; nt < NT; ++nt) dmma884(acc[mt][nt][0], acc[mt][nt][1], af[mt], bf[nt]);
            }
            __syncwarp();
        }
        dev = gt_warp_sum(dev);
        ll = gt_warp_sum(ll);
        sep = __all_sync(GT_FULL, sep);

        if (pass >= 1) {
            niter = pass;
            if (sep) { status = GT_PERFECT_SEPARATION; break; }
            if (fabs(dev - dev_prev) <= 1e-8 || pass == 100) { llf = ll; done = true; break; }
        }
        dev_prev = dev;

        // H = X'WX (upper tiles) and rhs = X'Wz (column p)
#pragma unroll
        for (int mt = 0; mt < NT; ++mt)
#pragma unroll
            for (int nt = mt; nt < NT; ++nt) {
                int r = mt * 8 + (lane >> 2), c0 = nt * 8 + 2 * (lane & 3);
                H[r * PH + c0] = acc[mt][nt][0];
                H[r * PH + c0 + 1] = acc[mt][nt][1];
            }
        __syncwarp();
        for (int t = lane; t < p * p; t += 32) {
            int i = t / p, j = t % p;
            Hc[i * p + j] = (i <= j) ? H[i * PH + j] : H[j * PH + i];
        }
        for (int i = lane; i < p; i += 32) rhs[i] = (p < PH) ? H[i * PH + p] : 0.0;
        __syncwarp();
        if (!gt_warp_cholesky(Hc, p, p, lane)) { status = GT_NAN_PVALUE; break; }
        gt_warp_chol_inverse(Hc, p, p, Hinv, lane);
        for (int i = lane; i < p; i += 32) {
            double s = 0.0;
            for (int j = 0; j < p; ++j) s = fma(Hinv[i * p + j], rhs[j], s);
            coef[i] = s;
        }
        __syncwarp();
        for (int i = lane; i < p; i += 32) beta[i] = coef[i];
        __syncwarp();
    }
    if (status != GT_OK) {
        write_fail(status, flip, maf, nan(""), niter);
        return;
    }

    // back to the reference's parameterisation (covariates: R^-1)
    for (int i = lane; i < p; i += 32) {
        double b = 0.0, v = 0.0;
        if (i < k) {
            for (int a = i; a < k; ++a) {
                double ria = T.Rinv[i * k + a];
                b = fma(ria, beta[a], b);
                double s = 0.0;
                for (int c = i; c < k; ++c) s = fma(Hinv[a * p + c], T.Rinv[i * k + c], s);
                v = fma(ria, s, v);
            }
        } else { b = beta[i]; v = Hinv[i * p + i]; }
        coef[i] = b; var[i] = v;
    }
    __syncwarp();
    double pv_snp = 0.0;
    for (int j = lane; j < p; j += 32) {
        double se = sqrt(var[j]);
        double z = coef[j] / se;
        double pval = gt_norm_sf2(z);
        double my[6] = {coef[j], se, coef[j] - GT_Z975 * se, coef[j] + GT_Z975 * se, z, pval};
#pragma unroll
        for (int f = 0; f < 6; ++f) out[(long long)(GT_F_PARAM0 + 6 * j + f) * ostride + snp] = my[f];
        if (j == k) pv_snp = pval;
    }
    pv_snp = __shfl_sync(GT_FULL, pv_snp, k & 31);
    if (!T.raw_mode && pv_snp != pv_snp) status = GT_NAN_PVALUE;
    if (lane == 0) {
        out[(long long)GT_F_MAF * ostride + snp] = maf;
        out[(long long)GT_F_AUX * ostride + snp] = nan("");
        out[(long long)GT_F_STAT * ostride + snp] = nev;
        out[(long long)GT_F_LLF * ostride + snp] = llf;
        iout[(long long)GT_I_STATUS * ostride + snp] = status;
        iout[(long long)GT_I_NOBS * ostride + snp] = nj;
        iout[(long long)GT_I_FLIP * ostride + snp] = flip;
        iout[(long long)GT_I_NITER * ostride + snp] = niter;
    }
}

}  // namespace gtk
