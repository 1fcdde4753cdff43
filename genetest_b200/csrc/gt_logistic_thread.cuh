// K4, thread-per-SNP version: batched IRLS for the logistic model, FP64, for
// .bed rows, a binary outcome, at most one SNP x covariate interaction column
// and at most 13 parameters.
//
// Restates sm.GLM(y, X, family=Binomial()).fit() as called by
// genetest/statistics/models/logistic.py:38-40 (start mu = (y + .5) / 2, one
// weighted least-squares solve per iteration, perfect-separation check, stop on
// abs(delta deviance) <= 1e-8, params / covariance from the LAST step's working
// weights) after _gwas_worker's complete-case mask, MAF and flip
// (genetest/analysis.py:119-157) -- the same arithmetic, sample by sample, as
// the warp-per-SNP kernel (gt_logistic.cuh), laid out the other way round:
//
//   * ONE THREAD OWNS ONE MARKER and keeps X'WX (upper triangle) and X'Wz in
//     registers: p (p + 1) / 2 + p FMAs per sample, nothing wasted (the m8n8k4
//     tensor-core tiles of the warp kernel compute 192 products for the 91 of
//     a 13 x 13 symmetric update, and every warp streams the whole design
//     through L1 for itself);
//   * all markers of a CTA walk the samples together, so the design rows
//     [q | w | y] are read ONCE per CTA: bulk copies (TMA, mbarrier complete_tx,
//     issued by warp 0 between its own blocks, or by a producer warp) fill a
//     shared-memory ring and every lane reads the same address (broadcast, no
//     bank conflicts);
//   * each thread streams its own packed row word by word, one word ahead (the
//     eight words of a 32-byte sector hit L1 after the first);
//   * exp / log / reciprocal are straight-line polynomials for the ranges the
//     pass needs, and the log-likelihood is summed with compensation (the stop
//     rule compares deviances to 1e-8 absolute);
//   * the p x p Cholesky solve, the triangular inverse and the map back to the
//     reference's parameterisation run fully unrolled in registers.
// The markers of a CTA take their IRLS passes in lockstep; one that needs
// more than `maxpass` passes (perfect separation takes dozens) is handed to the
// warp-per-SNP kernel (status GT_ITER_DEFER), which is the better shape for a
// few long-running markers.
#pragma once

#include "gt_common.cuh"
#include "gt_contract_tc.cuh"
#include "gt_contract_tc2.cuh"
#include "gt_logistic.cuh"

namespace gtk {

constexpr int GT_ITER_DEFER = -77;      // internal: finish this marker in the warp-per-SNP kernel
constexpr int kLtBlk = 256;             // samples per shared-memory block
constexpr int kLtBuf = 4;               // blocks in the ring
constexpr int kLtMaxThreads = 224;      // markers per CTA with a producer warp of its own
constexpr int kLtFoldThreads = 256;     // ... with the copies issued by warp 0 (FOLD): 8 x 32 x 255 registers


// ---------------------------------------------------------------------------
// exp(-a), log(1 + t) and 1 / d for the ranges the IRLS pass needs, without the
// special-case handling of the library routines (which cost more instructions
// per sample than the whole X'WX update).  Relative error <= 3e-16 on the
// stated ranges (tests/test_host_api.py checks the host build against numpy).
// The coefficients are plain literals in straight-line code: the compiler feeds
// them to the FMAs from the constant bank.
// ---------------------------------------------------------------------------

// Coefficients of the polynomials below.
#define GT_LT_COEFS { \
    1.4426950408889634, 6755399441055744.0, -6.93147180369123816490e-01, -1.90821492927058770002e-10, \
    1.0 / 6.0, 0.5, 1.0 / 120.0, 1.0 / 24.0, 1.0 / 5040.0, 1.0 / 720.0, 1.0 / 362880.0, 1.0 / 40320.0, \
    1.0 / 39916800.0, 1.0 / 3628800.0, 1.0 / 6227020800.0, 1.0 / 479001600.0, \
    1.4142135623730951, 1.0 / 3.0, 1.0 / 7.0, 1.0 / 5.0, 1.0 / 11.0, 1.0 / 9.0, 1.0 / 15.0, 1.0 / 13.0, \
    1.0 / 19.0, 1.0 / 17.0, 1.0 / 21.0, 0.6931471805599453, 700.0, 2.220446049250313e-16, \
    1.0 - 2.220446049250313e-16, 1e-8 }
// (measured: literals, which the compiler rebuilds with two moves each, beat LDC.64 loads
// from constant memory by 9 % on this kernel -- the loads' latency is not hidden)
__host__ __device__ constexpr double lt_coef(int i) {
    constexpr double t[32] = GT_LT_COEFS;
    return t[i];
}
#define GT_LTC(i) lt_coef(i)
GT_HD double lt_rcp(double d) {            // 1 / d, d in [0.5, 8]
#ifdef __CUDA_ARCH__
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;\n" : "=d"(r) : "d"(d));
    double e = fma(-d, r, 1.0);
    r = fma(r, e, r);
    e = fma(-d, r, 1.0);
    r = fma(r, e, r);
    e = fma(-d, r, 1.0);
    return fma(r, e, r);
#else
    return 1.0 / d;
#endif
}
GT_HD double lt_exp_neg(double a) {        // exp(-a), a >= 0 (a above 700 counts as 700)
    const double x = -fmin(a, GT_LTC(28));
    // n = round(x / ln 2) by the magic-number add; r = x - n ln 2 in two pieces
    const double magic = GT_LTC(1);
    const double tn = fma(x, GT_LTC(0), magic);
    const double nd = tn - magic;
    double r = fma(nd, GT_LTC(2), x);
    r = fma(nd, GT_LTC(3), r);
    // e^r, |r| <= 0.3466: Taylor to r^13 (next term 4e-18), Estrin-style in r^2
    const double r2 = r * r;
    const double q0 = r + 1.0;                                             // 1 + r
    const double q1 = fma(r, GT_LTC(4), GT_LTC(5));                            // 1/2! + r/3!
    const double q2 = fma(r, GT_LTC(6), GT_LTC(7));
    const double q3 = fma(r, GT_LTC(8), GT_LTC(9));
    const double q4 = fma(r, GT_LTC(10), GT_LTC(11));
    const double q5 = fma(r, GT_LTC(12), GT_LTC(13));
    const double q6 = fma(r, GT_LTC(14), GT_LTC(15));
    const double r4 = r2 * r2;
    const double a0 = fma(r2, q1, q0), a1 = fma(r2, q3, q2), a2 = fma(r2, q5, q4);
    const double r8 = r4 * r4;
    const double b0 = fma(r4, a1, a0), b1 = fma(r4, q6, a2);
    const double pr = fma(r8, b1, b0);
    // times 2^n: n >= -1010, the result stays normal
#ifdef __CUDA_ARCH__
    const int ni = __double2loint(tn);
    return __hiloint2double(__double2hiint(pr) + (ni << 20), __double2loint(pr));
#else
    return ldexp(pr, (int)nd);
#endif
}
GT_HD double lt_log1p01(double t, double &inv) {   // log(1 + t) and inv = 1 / (1 + t), t in [0, 1]
    const double d = 1.0 + t;
    inv = lt_rcp(d);
    // log d = e ln 2 + log m, m = d / 2^e in (0.7071, 1.4143]; log m = 2 atanh(f), f = (m - 1) / (m + 1)
    const bool big = d > GT_LTC(16);
    const double m = big ? 0.5 * d : d;
    const double f = (m - 1.0) * lt_rcp(m + 1.0);
    const double z = f * f;                              // <= 0.02944: z^11 / 23 < 1e-18
    const double z2 = z * z, z4 = z2 * z2;
    const double c0 = fma(z, GT_LTC(17), 1.0), c1 = fma(z, GT_LTC(18), GT_LTC(19));
    const double c2 = fma(z, GT_LTC(20), GT_LTC(21)), c3 = fma(z, GT_LTC(22), GT_LTC(23));
    const double c4 = fma(z, GT_LTC(24), GT_LTC(25)), c5 = GT_LTC(26);
    const double d0 = fma(z2, c1, c0), d1 = fma(z2, c3, c2), d2 = fma(z2, c5, c4);
    const double z8 = z4 * z4;
    const double pz = fma(z8, d2, fma(z4, d1, d0));
    const double lm = (f + f) * pz;
    return big ? lm + GT_LTC(27) : lm;
}

// FOLD = 1: no producer warp -- warp 0 issues the bulk copies between its own blocks (one block
// of lookahead less than the ring holds), which makes room for an eighth warp of markers.
// NI = 1: one SNP x covariate interaction column (design row [q | g | g w], analysis.py:159-166).
template <int P, int FOLD = 0, int NI = 0>
__global__ void __launch_bounds__(256, 1)
logistic_thread_kernel(IterDesignDev T, const uint8_t *__restrict__ geno, long long pitch, int n_snps,
                       double *__restrict__ out, int *__restrict__ iout, long long ostride,
                       double *__restrict__ scratch /* [2 P][lds] coef, se of the last solve */, long long lds,
                       int maxpass) {
    constexpr int K = P - 1 - NI;                         // covariates (orthonormal basis q)
    constexpr int NH = P * (P + 1) / 2;
    constexpr int XS = (K + NI + 1) & ~1;                 // row stride of T.X: q (K) | w (NI)
    constexpr int XSs = XS > 0 ? XS : 2;
    constexpr int XBLK = kLtBlk * XSs * 8, YBLK = kLtBlk * 8;
    constexpr int NF = GT_F_PARAM0 + 6 * P;
    const double EPS = 2.220446049250313e-16;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned char *smem = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);
    double *sX = (double *)smem;                                   // kLtBuf x [kLtBlk][XSs]
    double *sY = (double *)(smem + kLtBuf * XBLK);                 // kLtBuf x [kLtBlk]
    uint64_t *full = (uint64_t *)(smem + kLtBuf * (XBLK + YBLK)), *empty = full + kLtBuf;
    int *fail = (int *)(empty + kLtBuf);
    double *sRinv = (double *)(full + 16);                         // [K * K]
    double *sBeta = sRinv + ((K * K + 1) & ~1);                    // [P][nthr]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nthr = blockDim.x - (FOLD ? 0 : 32);                 // compute threads
    const int ncw = nthr >> 5;
    if (tid == 0) {
        for (int s = 0; s < kLtBuf; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], ncw); }
        *fail = 0;
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    for (int t = tid; t < K * K; t += blockDim.x) sRinv[t] = T.Rinv[t];
    __syncthreads();
    const int nblk = (T.n_file + kLtBlk - 1) / kLtBlk;

    if (!FOLD && warp == ncw) {
        // ===== producer: the design rows of one pass after the other, while any marker of the CTA is active =====
        int it = 0;
        for (int pass = 0; pass < maxpass; ++pass) {
            for (int b = 0; b < nblk; ++b, ++it) {
                const int s = it % kLtBuf;
                if (it >= kLtBuf && !mbar_wait2(&empty[s], ((it / kLtBuf) & 1) ^ 1, fail)) *fail = 1;
                if (elect_one()) {
                    mbar_expect_tx(&full[s], XBLK + YBLK);
                    bulk_g2s((unsigned char *)sX + s * XBLK, (const unsigned char *)T.X + (long long)b * XBLK, XBLK, &full[s]);
                    bulk_g2s((unsigned char *)sY + s * YBLK, (const unsigned char *)T.y + (long long)b * YBLK, YBLK, &full[s]);
                }
                __syncwarp();
            }
            if (!__syncthreads_or(0)) break;
        }
        return;
    }

    // ===== one thread per marker =====
    const int snp = blockIdx.x * nthr + tid;
    const bool owner = snp < n_snps;
    const uint4 *rowp = (const uint4 *)(geno + (long long)(owner ? snp : 0) * pitch);
    const int nchunk = (T.n_file + 63) / 64;              // 16-byte chunks of a row that hold samples
    const uint32_t *roww = (const uint32_t *)rowp;
    const int nword = (T.n_file + 15) / 16;               // 4-byte words of a row that hold samples
    // (the eight words of a 32-byte sector are read one after the other: seven L1 hits)
    auto word_at = [&](int w) { return (owner && w < nword) ? __ldg(roww + w) : 0x55555555u; };
    auto chunk_at = [&](int c) { return (owner && c < nchunk) ? __ldg(rowp + c) : make_uint4(0x55555555u, 0x55555555u, 0x55555555u, 0x55555555u); };
    auto write_fail = [&](int status, int nobs, int flip, double maf, double aux, int niter) {
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
        iout[(long long)GT_I_NITER * ostride + snp] = niter;
    };

    // ---- complete cases, MAF, flip (analysis.py:119-157): genotypes and outcome only ----
    int nj = 0, sum_g = 0, nev = 0;
    {
        uint4 cur = chunk_at(0), nxt = chunk_at(1);
        for (int c = 0; c < nchunk; ++c) {
            const uint4 gw = cur;
            cur = nxt; nxt = chunk_at(c + 2);
            const uint32_t words[4] = {gw.x, gw.y, gw.z, gw.w};
#pragma unroll
            for (int w4 = 0; w4 < 4; ++w4) {
                const double *yp = T.y + (long long)c * 64 + w4 * 16;   // n_pad rows: NaN past the design
#pragma unroll
                for (int s = 0; s < 16; ++s) {
                    const double yv = __ldg(yp + s);
                    const uint32_t code = (words[w4] >> (2 * s)) & 3u;
                    const bool ok = (yv == yv) && code != 1u;
                    nj += ok ? 1 : 0;
                    sum_g += ok ? (code == 0u ? 2 : (code == 2u ? 1 : 0)) : 0;
                    nev += (ok && yv != 0.0) ? 1 : 0;
                }
            }
        }
    }
    bool active = owner;
    int status = GT_OK, flip = 0, niter = 0;
    double maf = nan("");
    if (owner) {
        if (nj == 0) { write_fail(GT_ALL_MISSING, 0, 0, nan(""), nan(""), 0); active = false; }
        else if (!T.raw_mode) {
            maf = ((double)sum_g / (double)nj) / 2.0;
            if (maf > 0.5) { maf = 1.0 - maf; flip = 1; }
            if (T.maf_o != nullptr) { maf = T.maf_o[snp]; flip = T.flip_o[snp]; }
            if (T.use_maf_t && maf < T.maf_t) { write_fail(GT_MAF_BELOW, nj, flip, maf, maf, 0); active = false; }
        }
    }
    const double g_hom = flip ? 0.0 : 2.0, g_ref = flip ? 2.0 : 0.0;   // codes 00 and 11

    // pass-0 values of a binary outcome, by the formulas of the general branch (gt_logistic.cuh)
    const double mu1 = (1.0 + 0.5) / 2.0, mu0 = (0.0 + 0.5) / 2.0;
    const double eta1 = log(mu1 / (1.0 - mu1)), eta0 = log(mu0 / (1.0 - mu0));
    const double dev1 = 2.0 * (1.0 * log(fmax(1.0 / (mu1 + 1e-20), EPS)) + 0.0 * log(fmax(0.0 / (1.0 - mu1 + 1e-20), EPS)));
    const double dev0 = 2.0 * (0.0 * log(fmax(0.0 / (mu0 + 1e-20), EPS)) + 1.0 * log(fmax(1.0 / (1.0 - mu0 + 1e-20), EPS)));

    double *myBeta = sBeta + tid;                           // beta[a] at myBeta[a * nthr]
#pragma unroll
    for (int a = 0; a < P; ++a) myBeta[a * nthr] = 0.0;
    double dev_prev = 0.0, llf = 0.0;
    int it = 0;                                             // blocks consumed (ring position)
    int issued = 0;                                         // FOLD, warp 0: blocks whose copies are issued
#define GT_H(a, b) H[(a) * P - (a) * ((a) - 1) / 2 + ((b) - (a))]      /* a <= b */
    for (int pass = 0; pass < maxpass; ++pass) {
        double H[NH], r[P];
#pragma unroll
        for (int i = 0; i < NH; ++i) H[i] = 0.0;
#pragma unroll
        for (int i = 0; i < P; ++i) r[i] = 0.0;
        // The stop rule compares deviances to 1e-8 absolute at a magnitude of 1e4 .. 1e5: the
        // log-likelihood is summed per 16 samples and those sums are added with compensation
        // (a plain running sum of 50,000 terms carries ~5e-10 of rounding noise).
        double dev = 0.0, ll = 0.0, ll_c = 0.0;
        bool sep = true;
        uint32_t wnext = word_at(0);
        for (int b = 0; b < nblk; ++b, ++it) {
            const int st = it % kLtBuf;
            if (FOLD && warp == 0) {
                // copies of this pass up to kLtBuf - 1 blocks ahead (never across the end of the
                // pass: whether another one follows is only known after it)
                const int lim = min(it + kLtBuf - 1, (pass + 1) * nblk - 1);
                for (; issued <= lim; ++issued) {
                    const int s2 = issued % kLtBuf;
                    if (issued >= kLtBuf) {
                        const uint32_t ea = smem_u32(&empty[s2]), ep = ((issued / kLtBuf) & 1) ^ 1;
                        int spin = 0;
                        while (!mbar_try(ea, ep)) if (++spin > (1 << 22)) { *fail = 1; break; }
                    }
                    if (elect_one()) {
                        const int bb = issued - pass * nblk;
                        mbar_expect_tx(&full[s2], XBLK + YBLK);
                        bulk_g2s((unsigned char *)sX + s2 * XBLK, (const unsigned char *)T.X + (long long)bb * XBLK, XBLK, &full[s2]);
                        bulk_g2s((unsigned char *)sY + s2 * YBLK, (const unsigned char *)T.y + (long long)bb * YBLK, YBLK, &full[s2]);
                    }
                    __syncwarp();
                }
            }
            {   // inline wait (a call in this loop would spill the accumulators around it)
                const uint32_t fa = smem_u32(&full[st]), fp = (it / kLtBuf) & 1;
                int spin = 0;
                while (!mbar_try(fa, fp)) if (++spin > (1 << 22)) { *fail = 1; break; }
            }
            const double *bx = sX + st * (kLtBlk * XSs), *by = sY + st * kLtBlk;
#pragma unroll 1
            for (int c = 0; c < kLtBlk / 64; ++c) {
#pragma unroll 1
                for (int w4 = 0; w4 < 4; ++w4) {
                    const uint32_t word = wnext;
                    wnext = word_at((b * (kLtBlk / 64) + c) * 4 + w4 + 1);
                    const double *xw = bx + (c * 64 + w4 * 16) * XSs, *yw = by + c * 64 + w4 * 16;
                    double ll_w = 0.0;
#pragma unroll 2 //SAMPLE
                    for (int s = 0; s < 16; ++s) {
                        const double yv = yw[s];
                        const double *xr = xw + s * XSs;
                        const uint32_t code = (word >> (2 * s)) & 3u;
                        const bool ok = (yv == yv) && code != 1u;
                        const double g = code == 0u ? g_hom : (code == 2u ? 1.0 : g_ref);
                        double x[P];
#pragma unroll
                        for (int a = 0; a < K; ++a) x[a] = xr[a];
                        x[K] = g;
                        if (NI) x[K + 1] = g * xr[K];
                        double eta, mu, li = 0.0;
                        if (pass == 0) {
                            // statsmodels' start for y in {0, 1}: the same constants for every sample
                            mu = yv != 0.0 ? mu1 : mu0;
                            eta = yv != 0.0 ? eta1 : eta0;
                            dev += ok ? (yv != 0.0 ? dev1 : dev0) : 0.0;
                        } else {
                            double e0 = 0.0, e1 = 0.0, e2 = 0.0;     // three chains instead of one
#pragma unroll
                            for (int a = 0; a < P; ++a) {
                                const double bv = myBeta[a * nthr];
                                if (a % 3 == 0) e0 = fma(x[a], bv, e0);
                                else if (a % 3 == 1) e1 = fma(x[a], bv, e1);
                                else e2 = fma(x[a], bv, e2);
                            }
                            eta = (e0 + e1) + e2;
                            // mu = 1 / (1 + exp(-eta)); log(1 - mu) = -(max(eta, 0) + log1p(exp(-|eta|)))
                            const double t = lt_exp_neg(fabs(eta));
                            double inv;
                            const double l1p = lt_log1p01(t, inv);
                            mu = (eta >= 0.0) ? inv : t * inv;
                            // y in {0, 1}: loglike_i = y eta + log(1 - mu), deviance_i = -2 loglike_i
                            li = yv * eta - (fmax(eta, 0.0) + l1p);
                            ll_w += ok ? li : 0.0;
                            if (ok && !(fabs(mu - yv) <= GT_LTC(31))) sep = false;
                        }
                        const double pc = mu < GT_LTC(29) ? GT_LTC(29) : (mu > GT_LTC(30) ? GT_LTC(30) : mu);
                        const double w = ok ? pc * (1.0 - pc) : 0.0;      // 1 / (link.deriv^2 * variance)
                        const double wz = ok ? fma(w, eta, yv - mu) : 0.0; // w * working response
#pragma unroll
                        for (int a = 0; a < P; ++a) {
                            r[a] = fma(x[a], wz, r[a]);
                            const double wx = w * x[a];
#pragma unroll
                            for (int c2 = a; c2 < P; ++c2) GT_H(a, c2) = fma(wx, x[c2], GT_H(a, c2));
                        }
                    }
                    {   // ll += ll_w (Kahan)
                        const double yk = ll_w - ll_c, tk = ll + yk;
                        ll_c = (tk - ll) - yk;
                        ll = tk;
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[st]);
        }

        if (pass >= 1) dev = -2.0 * ll;
        if (active) {
            if (pass >= 1) {
                niter = pass;
                if (sep) { status = GT_PERFECT_SEPARATION; active = false; }
                else if (fabs(dev - dev_prev) <= 1e-8) { llf = ll; active = false; }
            }
            if (!active) {
                if (status != GT_OK) write_fail(status, nj, flip, maf, nan(""), niter);
                else {
                    // converged: parameters and covariance of the last solve
                    double pv_snp = 0.0;
#pragma unroll 1
                    for (int j = 0; j < P; ++j) {
                        const double cf = scratch[(long long)j * lds + snp], se = scratch[(long long)(P + j) * lds + snp];
                        const double z = cf / se;
                        const double pval = gt_norm_sf2(z);
                        double *o = out + (long long)(GT_F_PARAM0 + 6 * j) * ostride + snp;
                        o[0] = cf; o[ostride] = se; o[2 * ostride] = cf - GT_Z975 * se;
                        o[3 * ostride] = cf + GT_Z975 * se; o[4 * ostride] = z; o[5 * ostride] = pval;
                        if (j == K) pv_snp = pval;
                    }
                    if (!T.raw_mode && pv_snp != pv_snp) status = GT_NAN_PVALUE;
                    out[(long long)GT_F_MAF * ostride + snp] = maf;
                    out[(long long)GT_F_AUX * ostride + snp] = nan("");
                    out[(long long)GT_F_STAT * ostride + snp] = (double)nev;
                    out[(long long)GT_F_LLF * ostride + snp] = llf;
                    iout[(long long)GT_I_STATUS * ostride + snp] = status;
                    iout[(long long)GT_I_NOBS * ostride + snp] = nj;
                    iout[(long long)GT_I_FLIP * ostride + snp] = flip;
                    iout[(long long)GT_I_NITER * ostride + snp] = niter;
                }
            }
        }
        if (active) {
            dev_prev = dev;
            // ---- Cholesky H = Lc Lc' in place: Lc(i, j) lives in H(j, i) ----
#define GT_L(i, j) GT_H(j, i)                                           /* i >= j */
            bool okc = true;
#pragma unroll
            for (int j = 0; j < P; ++j) {
                double d = GT_L(j, j);
#pragma unroll
                for (int c = 0; c < j; ++c) d = fma(-GT_L(j, c), GT_L(j, c), d);
                if (!(d > 0.0)) okc = false;
                const double sj = sqrt(d), inv = 1.0 / sj;
                GT_L(j, j) = sj;
#pragma unroll
                for (int i = j + 1; i < P; ++i) {
                    double v = GT_L(i, j);
#pragma unroll
                    for (int c = 0; c < j; ++c) v = fma(-GT_L(i, c), GT_L(j, c), v);
                    GT_L(i, j) = v * inv;
                }
            }
            if (!okc) {
                write_fail(GT_NAN_PVALUE, nj, flip, maf, nan(""), niter);
                active = false;
            } else {
                // beta = H^-1 X'Wz: forward, then back substitution (in r)
#pragma unroll
                for (int i = 0; i < P; ++i) {
                    double v = r[i];
#pragma unroll
                    for (int c = 0; c < i; ++c) v = fma(-GT_L(i, c), r[c], v);
                    r[i] = v / GT_L(i, i);
                }
#pragma unroll
                for (int i = P - 1; i >= 0; --i) {
                    double v = r[i];
#pragma unroll
                    for (int c = i + 1; c < P; ++c) v = fma(-GT_L(c, i), r[c], v);
                    r[i] = v / GT_L(i, i);
                }
#pragma unroll
                for (int a = 0; a < P; ++a) myBeta[a * nthr] = r[a];
                // Z = Lc^-1 in place (lower triangular), column by column from the last
#pragma unroll
                for (int j = P - 1; j >= 0; --j) {
                    const double zjj = 1.0 / GT_L(j, j);
                    GT_L(j, j) = zjj;
                    // x = Z[j+1:, j+1:] * Lc[j+1:, j], bottom row first (in place), then * -zjj
#pragma unroll
                    for (int i = P - 1; i > j; --i) {
                        double v = 0.0;
#pragma unroll
                        for (int c = j + 1; c <= i; ++c) v = fma(GT_L(i, c), GT_L(c, j), v);
                        GT_L(i, j) = -v * zjj;
                    }
                }
                // back to the reference's parameterisation (covariates: R^-1):
                // coef = T beta, var_i = | Z T' e_i |^2 with T = blockdiag(R^-1, 1)
#pragma unroll
                for (int i = 0; i < P; ++i) {
                    double cf = 0.0, var = 0.0;
                    if (i < K) {
#pragma unroll
                        for (int a = i; a < K; ++a) cf = fma(sRinv[i * K + a], r[a], cf);
#pragma unroll
                        for (int rr = i; rr < P; ++rr) {
                            double u = 0.0;
#pragma unroll
                            for (int a = i; a < K; ++a)
                                if (a <= rr) u = fma(GT_L(rr, a), sRinv[i * K + a], u);
                            var = fma(u, u, var);
                        }
                    } else {
                        cf = r[i];
#pragma unroll
                        for (int rr = i; rr < P; ++rr) var = fma(GT_L(rr, i), GT_L(rr, i), var);
                    }
                    scratch[(long long)i * lds + snp] = cf;
                    scratch[(long long)(P + i) * lds + snp] = sqrt(var);
                }
            }
#undef GT_L
        }
        if (!__syncthreads_or(active ? 1 : 0)) break;
    }
#undef GT_H
    // still running after the last pass: the warp-per-SNP kernel takes over
    if (active) iout[(long long)GT_I_STATUS * ostride + snp] = GT_ITER_DEFER;
    if (tid == 0 && *fail) printf("logistic_thread_kernel: pipeline wait timed out (block %d)\n", blockIdx.x);
}

}  // namespace gtk
