// K5, thread-per-SNP version: batched Newton solver for the Cox proportional
// hazards model with Efron's tie correction and strata, FP64, for .bed rows,
// designs without interaction and at most 12 parameters.
//
// Restates sm.PHReg(tte, X, status=event, ties="efron").fit() as called by
// genetest/statistics/models/survival.py:77-82 -- Newton from beta = 0 on
// loglike / nobs with a 1e-10 ridge on the Hessian diagonal, until every
// abs(step) <= 1e-8, covariance inv(-H(beta_hat)) -- after _gwas_worker's
// complete-case mask, MAF and flip (genetest/analysis.py:119-157): the same
// quantities as the warp-per-SNP kernel (gt_coxk.cuh), laid out the other way
// round.  With the samples in descending time, S0 / S1 the running risk-set sums
// and inv_t = 1 / S0 at the event times,
//   -H    = sum_i W_i x_i x_i'  -  sum_t u_t u_t',   u_t = S1 / S0 at event t
//   score = sum_events x_i  -  sum_i W_i x_i
//   W_i   = e_i * (sum of inv_t over the event times <= t_i)
// so ONE THREAD PER MARKER keeps one p x p triangle, S1 and the score in
// registers and walks the samples twice per Newton step: pass 1 totals inv_t,
// pass 2 accumulates.  No warp scans, no shuffles, no tensor-core tiles that
// compute 64 products for the 66 of an 11 x 11 triangle twice over.
// Tied events (Efron): with F0 / F1 the sums over the d tied events of a time
// point, term j of the group has den_j = S0 - (j/d) F0 and u_j = (S1 - (j/d) F1)
// / den_j; the group's totals are gathered by a look-ahead over its rows when
// its first member comes up, its d terms are added there, and its members then
// take W_i = e_i (sum of 1/den at or after the group - sum_j (j/d) / den_j).
// Strata: the sums restart with every stratum (one segment of blocks each).
//   * the design rows [q | flags | group size] (descending time, the same for
//     every marker) come through a TMA-fed shared-memory ring, read by all lanes
//     at the same address (broadcast);
//   * a first sweep gathers the marker's genotype codes into that order once
//     and writes them word-major (word w of all markers side by side): the
//     passes read them coalesced;
//   * the p x p Cholesky solve, the triangular inverse and the map back to the
//     reference's parameterisation run fully unrolled in registers.
// The markers of a CTA take their Newton steps in lockstep; whatever is not
// finished after `maxround` rounds (or needs the exact maximum of the linear
// predictor again) is left to the warp-per-SNP kernel (status GT_ITER_DEFER).
#pragma once

#include "gt_common.cuh"
#include "gt_contract_tc.cuh"
#include "gt_contract_tc2.cuh"
#include "gt_logistic.cuh"
#include "gt_logistic_thread.cuh"

namespace gtk {

// The genotype codes of `mpc` markers at a time in schedule order: their packed rows are
// staged in shared memory (coalesced), every thread then builds 16-code words for one marker
// (lane-fastest over the markers, so that word w of neighbouring markers is written side by
// side) -- a per-thread gather from global memory would fetch a 32-byte sector per code.
__global__ void __launch_bounds__(256)
cox_permute_kernel(const uint8_t *__restrict__ geno, long long pitch, int n_snps, int row_words,
                   const int *__restrict__ th_pos, int th_n, uint32_t *__restrict__ gperm, long long ldw,
                   int mpc, int spitch_words) {
    extern __shared__ __align__(16) uint32_t srow[];        // [mpc][spitch_words]
    const int snp0 = blockIdx.x * mpc;
    for (int m = 0; m < mpc; ++m) {
        const int snp = snp0 + m;
        if (snp >= n_snps) break;
        const uint32_t *src = (const uint32_t *)(geno + (long long)snp * pitch);
        for (int w = threadIdx.x; w < row_words; w += blockDim.x) srow[m * spitch_words + w] = __ldg(src + w);
    }
    __syncthreads();
    const int m = threadIdx.x % mpc, grp = threadIdx.x / mpc, ngrp = blockDim.x / mpc;
    const int snp = snp0 + m;
    if (snp >= n_snps) return;
    const uint8_t *row = (const uint8_t *)(srow + m * spitch_words);
    const int nword = th_n / 16;
    for (int ww = grp; ww < nword; ww += ngrp) {
        uint32_t word = 0;
#pragma unroll
        for (int s = 0; s < 16; ++s) {
            const int ps = __ldg(th_pos + ww * 16 + s);
            const uint32_t code = ps >= 0 ? ((uint32_t)row[ps >> 2] >> (2 * (ps & 3))) & 3u : 1u;
            word |= code << (2 * s);
        }
        gperm[(long long)ww * ldw + snp] = word;
    }
}

// FOLD = 1: warp 0 issues the bulk copies itself (see logistic_thread_kernel), 256 markers per CTA.
// TIES = 1: groups of tied events (Efron): a group's totals are gathered by a look-ahead over its
// rows, its d terms are added at its first member, its members then take the group's weight.
template <int P, int FOLD = 0, int TIES = 0>
__global__ void __launch_bounds__(256, 1)
cox_thread_kernel(IterDesignDev T, const uint8_t *__restrict__ geno, long long pitch, int n_snps,
                  double *__restrict__ out, int *__restrict__ iout, long long ostride,
                  uint32_t *__restrict__ gperm /* [th_n / 16][ldw] codes in schedule order */, long long ldw,
                  int maxround) {
    constexpr int K = P - 1;                              // covariates (Cox has no constant)
    constexpr int NH = P * (P + 1) / 2;
    constexpr int XST = (K + 3) & ~1;                     // q (K) | flags | group size
    constexpr int XBLK = kLtBlk * XST * 8;
    constexpr int NF = GT_F_PARAM0 + 6 * P;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned char *smem = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);
    double *sX = (double *)smem;                                   // kLtBuf x [kLtBlk][XST]
    uint64_t *full = (uint64_t *)(smem + kLtBuf * XBLK), *empty = full + kLtBuf;
    int *fail = (int *)(empty + kLtBuf);
    double *sRinv = (double *)(full + 16);                         // [K * K]
    double *sCmax = sRinv + ((K * K + 1) & ~1);                    // [P] max |x| per column
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nthr = blockDim.x - (FOLD ? 0 : 32);
    const int ncw = nthr >> 5;
    double *sBeta = sCmax + ((P + 1) & ~1);                        // [P][nthr]
    double *sF1 = sBeta + P * nthr;                                // TIES: [P][nthr] sum of e x over a group's events
    double *sS1g = sF1 + (TIES ? P * nthr : 0);                    // TIES: [P][nthr] S1 at the end of the group
    if (tid == 0) {
        for (int s = 0; s < kLtBuf; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], ncw); }
        *fail = 0;
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    for (int t = tid; t < K * K; t += blockDim.x) sRinv[t] = T.Rinv[t];
    for (int t = tid; t < P; t += blockDim.x) sCmax[t] = T.colmax[t];
    __syncthreads();
    const int nblk = T.th_n / kLtBlk;
    const int per_round = 2 * nblk;                        // every block twice (totals, then accumulation)

    if (!FOLD && warp == ncw) {
        // ===== producer: the design rows, stratum by stratum twice per round, while any marker is active =====
        int it = 0;
        for (int round = 0; round < maxround; ++round) {
            for (int b = 0; b < per_round; ++b, ++it) {
                const int s = it % kLtBuf;
                if (it >= kLtBuf && !mbar_wait2(&empty[s], ((it / kLtBuf) & 1) ^ 1, fail)) *fail = 1;
                if (elect_one()) {
                    mbar_expect_tx(&full[s], XBLK);
                    bulk_g2s((unsigned char *)sX + s * XBLK,
                             (const unsigned char *)T.th_X + (long long)__ldg(T.th_order + b) * XBLK, XBLK, &full[s]);
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
    (void)geno; (void)pitch;
    uint32_t *myw = gperm + (owner ? snp : 0);
    const int nword = T.th_n / 16;
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

    // ---- complete cases, MAF, flip from the codes in schedule order (cox_permute_kernel) ----
    int nj = 0, sum_g = 0, nev = 0;
    if (owner) {
#pragma unroll 1
        for (int w = 0; w < nword; ++w) {
            const uint32_t word = myw[(long long)w * ldw];
            const uint32_t s1 = word >> 1;
            const uint32_t hom = ~(word | s1) & 0x55555555u, het = s1 & ~word & 0x55555555u;
            const uint32_t mis = word & ~s1 & 0x55555555u;
            nj += 16 - __popc(mis);
            sum_g += 2 * __popc(hom) + __popc(het);
            // events among the complete cases: the flags of the 16 rows (the same for every lane)
            uint32_t evm = 0;
#pragma unroll
            for (int s = 0; s < 16; ++s)
                evm |= (uint32_t)((int)__ldg(T.th_X + (long long)(w * 16 + s) * XST + K) & 1) << (2 * s);
            nev += __popc(evm & ~mis);
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

    double *myBeta = sBeta + tid;                           // beta[a] at myBeta[a * nthr]
    double *myF1 = sF1 + tid, *myS1g = sS1g + tid;
#pragma unroll
    for (int a = 0; a < P; ++a) myBeta[a * nthr] = 0.0;
    // shift of the linear predictor: an upper bound of its maximum (0 at beta = 0, grows by
    // at most sum_a |step_a| max |x_a| per step); past a drift of 20 the marker is deferred
    double emax = 0.0, slack = 0.0, llf = 0.0;
    bool final_eval = false;
    int it = 0, issued = 0;
    // FOLD: warp 0 keeps the copies of this round up to kLtBuf - 1 blocks ahead (never across its end)
    auto feed = [&](int round) {
        if (!(FOLD && warp == 0)) return;
        const int lim = min(it + kLtBuf - 1, (round + 1) * per_round - 1);
        for (; issued <= lim; ++issued) {
            const int s2 = issued % kLtBuf;
            if (issued >= kLtBuf) {
                const uint32_t ea = smem_u32(&empty[s2]), ep = ((issued / kLtBuf) & 1) ^ 1;
                int spin = 0;
                while (!mbar_try(ea, ep)) if (++spin > (1 << 22)) { *fail = 1; break; }
            }
            if (elect_one()) {
                mbar_expect_tx(&full[s2], XBLK);
                bulk_g2s((unsigned char *)sX + s2 * XBLK,
                         (const unsigned char *)T.th_X + (long long)__ldg(T.th_order + issued % per_round) * XBLK,
                         XBLK, &full[s2]);
            }
            __syncwarp();
        }
    };
    auto wait_block = [&]() {
        const int st = it % kLtBuf;
        const uint32_t fa = smem_u32(&full[st]), fp = (it / kLtBuf) & 1;
        int spin = 0;
        while (!mbar_try(fa, fp)) if (++spin > (1 << 22)) { *fail = 1; break; }
        return st;
    };
    // linear predictor and exp(eta - emax) of schedule row `row`, read from global memory
    // (the look-ahead over a group of tied events); returns false for a missing genotype
    auto row_global = [&](int row, double &e, double (&x)[P]) {
        const uint32_t code = (myw[(long long)(row >> 4) * ldw] >> (2 * (row & 15))) & 3u;
        const double *xr = T.th_X + (long long)row * XST;
        x[K] = code == 0u ? g_hom : (code == 2u ? 1.0 : g_ref);
        double eta = x[K] * myBeta[K * nthr];
#pragma unroll
        for (int a = 0; a < K; ++a) { x[a] = __ldg(xr + a); eta = fma(x[a], myBeta[a * nthr], eta); }
        const bool ok = code != 1u;
        e = ok ? lt_exp_neg(emax - eta) : 0.0;
        return ok;
    };
#define GT_H(a, b) H[(a) * P - (a) * ((a) - 1) / 2 + ((b) - (a))]      /* a <= b */
#define GT_L(i, j) GT_H(j, i)                                           /* i >= j */
    for (int round = 0; round < maxround; ++round) {
        double H[NH], sc[P];
#pragma unroll
        for (int i = 0; i < NH; ++i) H[i] = 0.0;
#pragma unroll
        for (int i = 0; i < P; ++i) sc[i] = 0.0;
        double ll = 0.0;
        for (int seg = 0; seg < T.th_nseg; ++seg) {
        const int b0 = __ldg(T.th_seg + seg), b1 = __ldg(T.th_seg + seg + 1);
        // ---- pass 1: total of 1 / den over the event terms of the stratum ----
        double Ttot = 0.0;
        {
            double S0 = 0.0, F0 = 0.0;
            int dk = 0;
            for (int b = b0; b < b1; ++b, ++it) {
                feed(round);
                const int st = wait_block();
                const double *bx = sX + st * (kLtBlk * XST);
#pragma unroll 1
                for (int w16 = 0; w16 < kLtBlk / 16; ++w16) {
                    const uint32_t word = myw[(long long)(b * (kLtBlk / 16) + w16) * ldw];
                    const double *xw = bx + w16 * 16 * XST;
#pragma unroll 4
                    for (int s = 0; s < 16; ++s) {
                        const double *xr = xw + s * XST;
                        const uint32_t code = (word >> (2 * s)) & 3u;
                        const bool ok = code != 1u;
                        const double g = code == 0u ? g_hom : (code == 2u ? 1.0 : g_ref);
                        double e0 = g * myBeta[K * nthr], e1 = 0.0;
#pragma unroll
                        for (int a = 0; a < K; ++a) {
                            if (a & 1) e1 = fma(xr[a], myBeta[a * nthr], e1);
                            else e0 = fma(xr[a], myBeta[a * nthr], e0);
                        }
                        const double e = ok ? lt_exp_neg(emax - (e0 + e1)) : 0.0;
                        S0 += e;
                        const int fl = (int)xr[K];                          // warp-uniform
                        if (TIES && (fl & 2)) {
                            F0 += e; dk += ok ? 1 : 0;
                            if (fl & 8) {
                                const int dsz = (int)xr[K + 1];
                                const double dinv = dk > 0 ? lt_rcp((double)dk) : 0.0;
#pragma unroll 1
                                for (int j = 0; j < dsz; ++j)
                                    if (j < dk) Ttot += lt_rcp(S0 - ((double)j * dinv) * F0);
                                F0 = 0.0; dk = 0;
                            }
                        } else if (fl & 1) {
                            Ttot += ok ? lt_rcp(S0) : 0.0;
                        }
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty[st]);
            }
        }
        // ---- pass 2: log-likelihood, score, Hessian ----
        double S1[P];
#pragma unroll
        for (int i = 0; i < P; ++i) S1[i] = 0.0;
        double S0 = 0.0, PI = 0.0;
        double Mg = 0.0, Ag = 0.0;                          // TIES: weight of the current group's members, its sum of 1 / den
        for (int b = b0; b < b1; ++b, ++it) {
            feed(round);
            const int st = wait_block();
            const double *bx = sX + st * (kLtBlk * XST);
#pragma unroll 1
            for (int w16 = 0; w16 < kLtBlk / 16; ++w16) {
                const uint32_t word = myw[(long long)(b * (kLtBlk / 16) + w16) * ldw];
                const double *xw = bx + w16 * 16 * XST;
#pragma unroll 2
                for (int s = 0; s < 16; ++s) {
                    const double *xr = xw + s * XST;
                    const int fl = (int)xr[K];                          // warp-uniform
                    if (TIES && (fl & 4)) {
                        // ---- first member of a group of tied events: the group's totals by a
                        //      look-ahead over its rows, then its d Efron terms ----
                        const int dsz = (int)xr[K + 1];
                        const int row0 = (b * (kLtBlk / 16) + w16) * 16 + s;
                        double S0g = S0, F0 = 0.0;
                        int dk = 0;
#pragma unroll
                        for (int a = 0; a < P; ++a) { myS1g[a * nthr] = S1[a]; myF1[a * nthr] = 0.0; }
#pragma unroll 1
                        for (int m = 0; m < dsz; ++m) {
                            double em, xm[P];
                            const bool okm = row_global(row0 + m, em, xm);
                            S0g += em; F0 += em; dk += okm ? 1 : 0;
#pragma unroll
                            for (int a = 0; a < P; ++a) {
                                myS1g[a * nthr] = fma(em, xm[a], myS1g[a * nthr]);
                                myF1[a * nthr] = fma(em, xm[a], myF1[a * nthr]);
                            }
                        }
                        const double dinv = dk > 0 ? lt_rcp((double)dk) : 0.0;
                        double Cg = 0.0;
                        Ag = 0.0;
#pragma unroll 1
                        for (int j = 0; j < dsz; ++j) {
                            if (j < dk) {
                                const double cj = (double)j * dinv;
                                const double den = S0g - cj * F0;
                                const double inv = lt_rcp(den);
                                Ag += inv; Cg = fma(cj, inv, Cg);
                                if (final_eval) ll -= log(den);
                                double u[P];
#pragma unroll
                                for (int a = 0; a < P; ++a) u[a] = fma(-cj, myF1[a * nthr], myS1g[a * nthr]) * inv;
#pragma unroll
                                for (int a = 0; a < P; ++a)
#pragma unroll
                                    for (int c2 = a; c2 < P; ++c2) GT_H(a, c2) = fma(-u[a], u[c2], GT_H(a, c2));
                            }
                        }
                        // members: W_i = e_i (sum of 1 / den over the terms at or after the group - sum_j c_j / den_j)
                        Mg = (Ttot - PI) - Cg;
                    }
                    const uint32_t code = (word >> (2 * s)) & 3u;
                    const bool ok = code != 1u;
                    const double g = code == 0u ? g_hom : (code == 2u ? 1.0 : g_ref);
                    double x[P];
#pragma unroll
                    for (int a = 0; a < K; ++a) x[a] = xr[a];
                    x[K] = g;
                    double e0 = g * myBeta[K * nthr], e1 = 0.0;
#pragma unroll
                    for (int a = 0; a < K; ++a) {
                        if (a & 1) e1 = fma(x[a], myBeta[a * nthr], e1);
                        else e0 = fma(x[a], myBeta[a * nthr], e0);
                    }
                    const double eta = e0 + e1;
                    const double e = ok ? lt_exp_neg(emax - eta) : 0.0;
                    S0 += e;
#pragma unroll
                    for (int a = 0; a < P; ++a) S1[a] = fma(e, x[a], S1[a]);
                    double W = e * (Ttot - PI);
                    if (TIES && (fl & 2)) {
                        // member of a group of tied events: the group's terms are in already
                        W = e * Mg;
                        if (ok) {
                            if (final_eval) ll += eta - emax;
#pragma unroll
                            for (int a = 0; a < P; ++a) sc[a] += x[a];
                        }
                        if (fl & 8) PI += Ag;
                    } else if (fl & 1) {
                        // an event of its own (W counts its term): u = S1 / S0
                        const double inv = ok ? lt_rcp(S0) : 0.0;
                        PI += inv;
                        if (final_eval && ok) ll += (eta - emax) - log(S0);
#pragma unroll
                        for (int a = 0; a < P; ++a) {
                            const double ua = S1[a] * inv;
                            sc[a] += ok ? x[a] : 0.0;
#pragma unroll
                            for (int c2 = a; c2 < P; ++c2) GT_H(a, c2) = fma(-ua, S1[c2] * inv, GT_H(a, c2));
                        }
                    }
#pragma unroll
                    for (int a = 0; a < P; ++a) {
                        const double wx = W * x[a];
                        sc[a] -= wx;
#pragma unroll
                        for (int c2 = a; c2 < P; ++c2) GT_H(a, c2) = fma(wx, x[c2], GT_H(a, c2));
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[st]);
        }
        }   // strata

        if (active) {
            const double dn = (double)nj;
            if (!final_eval) {
                // Newton step on loglike / nobs with the reference's ridge: (-H / n - ridge... ) as in
                // gt_coxk.cuh: Hc = H / n with 1e-10 taken off the diagonal, rhs = score / n
#pragma unroll
                for (int a = 0; a < P; ++a) {
#pragma unroll
                    for (int c2 = a; c2 < P; ++c2) GT_H(a, c2) = GT_H(a, c2) / dn - (a == c2 ? 1e-10 : 0.0);
                    sc[a] /= dn;
                }
            } else llf = ll;
            // ---- Cholesky in place ----
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
                write_fail(GT_SINGULAR, nj, flip, maf, nan(""), niter);
                active = false;
            } else if (!final_eval) {
                // step = Hc^-1 score: forward, then back substitution (in sc)
#pragma unroll
                for (int i = 0; i < P; ++i) {
                    double v = sc[i];
#pragma unroll
                    for (int c = 0; c < i; ++c) v = fma(-GT_L(i, c), sc[c], v);
                    sc[i] = v / GT_L(i, i);
                }
#pragma unroll
                for (int i = P - 1; i >= 0; --i) {
                    double v = sc[i];
#pragma unroll
                    for (int c = i + 1; c < P; ++c) v = fma(-GT_L(c, i), sc[c], v);
                    sc[i] = v / GT_L(i, i);
                }
                // convergence is judged on the reference's parameterisation: T step
                double worst = 0.0, grow = 0.0;
#pragma unroll
                for (int i = 0; i < P; ++i) {
                    double dd = 0.0;
                    if (i < K) {
#pragma unroll
                        for (int a = i; a < K; ++a) dd = fma(sRinv[i * K + a], sc[a], dd);
                    } else dd = sc[i];
                    worst = fmax(worst, fabs(dd));
                    grow = fma(fabs(sc[i]), sCmax[i], grow);
                    myBeta[i * nthr] += sc[i];
                }
                emax += grow; slack += grow;
                niter += 1;
                if (!(worst == worst)) { write_fail(GT_NAN_PVALUE, nj, flip, maf, nan(""), niter); active = false; }
                else if (niter >= T.newton_maxiter) { write_fail(GT_NO_CONVERGENCE, nj, flip, maf, nan(""), niter); active = false; }
                else if (!(worst > 1e-8)) final_eval = true;
                else if (slack > 20.0) {
                    // the bound has drifted: the warp kernel recomputes the exact maximum
                    iout[(long long)GT_I_STATUS * ostride + snp] = GT_ITER_DEFER;
                    active = false;
                }
            } else {
                // ---- converged: covariance inv(-H) = Z'Z with Z = Lc^-1 ----
#pragma unroll
                for (int j = P - 1; j >= 0; --j) {
                    const double zjj = 1.0 / GT_L(j, j);
                    GT_L(j, j) = zjj;
#pragma unroll
                    for (int i = P - 1; i > j; --i) {
                        double v = 0.0;
#pragma unroll
                        for (int c = j + 1; c <= i; ++c) v = fma(GT_L(i, c), GT_L(c, j), v);
                        GT_L(i, j) = -v * zjj;
                    }
                }
                double pv_snp = 0.0;
#pragma unroll
                for (int i = 0; i < P; ++i) {
                    double cf = 0.0, var = 0.0;
                    if (i < K) {
#pragma unroll
                        for (int a = i; a < K; ++a) cf = fma(sRinv[i * K + a], myBeta[a * nthr], cf);
#pragma unroll
                        for (int rr = i; rr < P; ++rr) {
                            double u = 0.0;
#pragma unroll
                            for (int a = i; a < K; ++a)
                                if (a <= rr) u = fma(GT_L(rr, a), sRinv[i * K + a], u);
                            var = fma(u, u, var);
                        }
                    } else {
                        cf = myBeta[K * nthr];
                        var = GT_L(K, K) * GT_L(K, K);
                    }
                    const double se = sqrt(var), z = cf / se, pval = gt_norm_sf2(z);
                    double *o = out + (long long)(GT_F_PARAM0 + 6 * i) * ostride + snp;
                    o[0] = cf; o[ostride] = se; o[2 * ostride] = cf - GT_Z975 * se;
                    o[3 * ostride] = cf + GT_Z975 * se; o[4 * ostride] = z; o[5 * ostride] = pval;
                    if (i == K) pv_snp = pval;
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
                active = false;
            }
        }
        if (!__syncthreads_or(active ? 1 : 0)) break;
    }
#undef GT_L
#undef GT_H
    if (active) iout[(long long)GT_I_STATUS * ostride + snp] = GT_ITER_DEFER;
    if (tid == 0 && *fail) printf("cox_thread_kernel: pipeline wait timed out (block %d)\n", blockIdx.x);
}

}  // namespace gtk
