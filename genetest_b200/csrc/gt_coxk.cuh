// K5: batched warp-per-SNP Newton solver for the Cox proportional hazards
// model with Efron's tie correction, FP64.
//
// Restates sm.PHReg(tte, X, status=event, ties="efron").fit() as called by
// genetest/statistics/models/survival.py:77-82: Newton from beta = 0 on
// loglike / nobs with a 1e-10 ridge on the Hessian diagonal, until every
// abs(step) <= 1e-8 (maxiter 100 -> "failed to converge"), covariance
// inv(-H(beta_hat)) (singular -> "Singular matrix"), for every marker of a
// batch after _gwas_worker's complete-case mask, MAF, flip and interaction
// columns (genetest/analysis.py:119-166).
//
// Samples are scheduled once on the host in descending time, 32 per block,
// tie groups kept inside a block.  With S0, S1 the running risk-set sums and
// F0, F1 the sums over the tied events of a group, every event (group g,
// rank j of d) contributes  eta - log(den),  den = S0 - (j/d) F0.  The
// Hessian is accumulated without prefix sums of p x p matrices:
//   -H = sum_i W_i x_i x_i'  -  sum_(g,j) u u',   u = (S1 - (j/d) F1) / den
//   W_i = e_i * ( sum_{(g,j): t_g <= t_i} 1/den  -  [i is an event] sum_j (j/d)/den_gj )
// so one pass needs scalar / p-vector warp scans and two FP64 tensor-core
// (m8n8k4) accumulations per 32 samples.
#pragma once

#include <type_traits>

#include "gt_common.cuh"
#include "gt_logistic.cuh"

namespace gtk {

__device__ __forceinline__ double warp_scan_incl(double v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        double t = __shfl_up_sync(GT_FULL, v, o);
        if (lane >= o) v += t;
    }
    return v;
}
__device__ __forceinline__ int warp_scan_incl_int(int v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(GT_FULL, v, o);
        if (lane >= o) v += t;
    }
    return v;
}

template <int NT>
struct CoxCfg {
    static constexpr int PH = 8 * NT;
    // staging tiles are TRANSPOSED: row a (design column) holds the 32
    // samples of the block; a row stride of 36 doubles keeps the per-lane
    // stores and the MMA fragment loads bank-conflict free
    static constexpr int XT = 36;
    // xT, uT (p rows each: the MMA fragment loads of rows p .. PH-1 run into
    // the arrays that follow and only feed ignored outputs), per-sample
    // weights (W, G), H, Hc, Hinv, beta, step, carry, coef, var, prev
    static constexpr int fixed = 64 + 3 * PH * PH + 6 * PH;
    __host__ __device__ static constexpr int perWarp(int p) { return 2 * p * XT + fixed; }
};

// one marker, by one warp (sm: the CTA's dynamic shared memory)
template <bool PACKED, int NT>
__device__ __forceinline__ void
cox_newton_one(const IterDesignDev &T, const void *__restrict__ geno, long long pitch, int n_snps,
               double *__restrict__ out, int *__restrict__ iout, long long ostride,
               const int snp, double *sm) {
    using Cfg = CoxCfg<NT>;
    constexpr int XT = Cfg::XT, PH = Cfg::PH;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int k = T.k, I = T.I, p = T.p;
    const int nf = GT_F_PARAM0 + 6 * p;

    double *W = sm + (size_t)warp * Cfg::perWarp(p);
    double *xT = W, *uT = xT + p * XT, *wv = uT + p * XT, *gv = wv + 32, *H = gv + 32,
           *Hc = H + PH * PH, *Hinv = Hc + PH * PH, *beta = Hinv + PH * PH, *step = beta + PH,
           *carry = step + PH, *coef = carry + PH, *var = coef + PH, *prev = var + PH;

    const void *rowp = PACKED ? (const void *)((const uint8_t *)geno + (long long)snp * pitch)
                              : (const void *)((const double *)geno + (long long)snp * pitch);
    const int nblk = T.n_slots / 32;

    // ---- complete cases, MAF, flip ----
    int nj = 0;
    double sum_g = 0.0, nev = 0.0;
    for (int b = 0; b < nblk; ++b) {
        int slot = b * 32 + lane;
        int pos = T.slot_pos[slot];
        double g = 0.0;
        bool ok = pos >= 0;
        if (ok) ok = load_geno<PACKED>(rowp, pos, g);
        if (ok) { nj += 1; sum_g += g; nev += (T.slot_info[slot] >> 10) & 1; }
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

    // rank-deficient covariates: inv(-H) fails in the reference ("Singular matrix")
    if (T.degenerate) { write_fail(GT_SINGULAR, flip, maf, nan(""), 0); return; }

    for (int t = lane; t < Cfg::perWarp(p); t += 32) xT[t] = 0.0;   // nothing non-finite behind the tiles
    __syncwarp();
    for (int t = lane; t < PH; t += 32) { beta[t] = 0.0; prev[t] = INFINITY; }
    __syncwarp();

    // design row of a slot: eta = x'beta and validity; STORE also writes the
    // row into xT[0..p)[lane]
    auto load_row_impl = [&](int slot, double &eta, int &info, auto store_tag) -> bool {
        constexpr bool STORE = decltype(store_tag)::value;
        int pos = T.slot_pos[slot];
        info = T.slot_info[slot];
        double g = 0.0;
        bool ok = pos >= 0;
        if (ok) ok = load_geno<PACKED>(rowp, pos, g);
        double *xr = xT + lane;
        eta = 0.0;
        if (ok) {
            if (flip) g = 2.0 - g;
            const double *xi = T.Xs + (long long)slot * T.XS;
            for (int a = 0; a < k; ++a) {
                double v = xi[a];
                if (STORE) xr[a * XT] = v;
                eta = fma(v, beta[a], eta);
            }
            if (STORE) xr[k * XT] = g;
            eta = fma(g, beta[k], eta);
            for (int j = 0; j < I; ++j) {
                double v = g * xi[k + j];
                if (STORE) xr[(k + 1 + j) * XT] = v;
                eta = fma(v, beta[k + 1 + j], eta);
            }
        } else if (STORE) {
            for (int a = 0; a < p; ++a) xr[a * XT] = 0.0;
        }
        return ok;
    };
    auto load_row = [&](int slot, double &eta, int &info) {
        return load_row_impl(slot, eta, info, std::true_type{});
    };
    auto load_eta = [&](int slot, double &eta, int &info) {
        return load_row_impl(slot, eta, info, std::false_type{});
    };

    int status = GT_OK, niter = 0;
    double llf = 0.0;
    bool final_eval = false;
    // Shift of the linear predictor (the reference subtracts its maximum).
    // At beta = 0 the maximum is 0; after a step it grows by at most
    // sum_a |step_a| max|x_a|, so the pass that finds the exact maximum only
    // runs once that bound has drifted by more than 20.
    double emax = 0.0, slack = 0.0;
    while (true) {
        // ---- pass 0: max of the linear predictor ----
        // (real-valued / raw columns are not bounded by 2: always exact)
        if (!PACKED || T.raw_mode || slack > 20.0 || !(slack == slack)) {
            emax = -INFINITY;
            for (int b = 0; b < nblk; ++b) {
                double eta; int info;
                if (load_eta(b * 32 + lane, eta, info)) emax = fmax(emax, eta);
            }
            emax = gt_warp_max(emax);
            slack = 0.0;
        }
        // ---- passes 1 and 2, stratum by stratum (PHReg(strata=...), survival.py:77-82: the
        // risk sets restart with every stratum; one segment of blocks each, a single one
        // without strata) ----
        double accH[NT][NT][2], accU[NT][NT][2];
#pragma unroll
        for (int a = 0; a < NT; ++a)
#pragma unroll
            for (int b = 0; b < NT; ++b) accH[a][b][0] = accH[a][b][1] = accU[a][b][0] = accU[a][b][1] = 0.0;
        double ll = 0.0;
        // A = x, B = [W x | G]: the weights are applied while the fragments
        // are loaded (column p of B carries the score); U accumulates u u'
        auto mma_block = [&]() {
            __syncwarp();
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) {
                const int e4 = ks * 4 + (lane & 3), cc = lane >> 2;
                const double wsc = wv[e4], gsc = gv[e4];
                double xf[NT], wf[NT], uf[NT];
#pragma unroll
                for (int t = 0; t < NT; ++t) {
                    xf[t] = xT[(t * 8 + cc) * XT + e4];
                    uf[t] = uT[(t * 8 + cc) * XT + e4];
                    wf[t] = (t * 8 + cc == p) ? gsc : wsc * xf[t];
                }
#pragma unroll
                for (int mt = 0; mt < NT; ++mt)
#pragma unroll
                    for (int nt = mt; nt < NT; ++nt) {
                        dmma884(accH[mt][nt][0], accH[mt][nt][1], xf[mt], wf[nt]);
                        dmma884(accU[mt][nt][0], accU[mt][nt][1], uf[mt], uf[nt]);
                    }
            }
            __syncwarp();
        };
        for (int seg = 0; seg < T.n_seg; ++seg) {
        const int seg_b0 = T.seg_start[seg], seg_b1 = T.seg_start[seg + 1];
        // ---- pass 1: total of 1/den over the event terms of the stratum ----
        double Ttot = 0.0;
        {
            double c0 = 0.0;
            for (int b = seg_b0; b < seg_b1; ++b) {
                const int nbig = T.blk_info[b];
                if (nbig > 1) {
                    // tied events over several blocks: den_j = S0g - (j/d) EG
                    double EG = 0.0; int dg = 0;
                    for (int bb = b; bb < b + nbig; ++bb) {
                        double eta; int info;
                        bool ok = load_eta(bb * 32 + lane, eta, info);
                        if (ok) { EG += exp(eta - emax); dg += 1; }
                    }
                    EG = gt_warp_sum(EG); dg = gt_warp_sum_int(dg);
                    const double S0g = c0 + EG;
                    for (int j = lane; j < dg; j += 32)
                        Ttot += 1.0 / (S0g - ((double)j / (double)dg) * EG);
                    c0 = S0g;
                    b += nbig - 1;
                    continue;
                }
                double eta; int info;
                bool ok = load_eta(b * 32 + lane, eta, info);
                const int la = info & 31, lb = (info >> 5) & 31;
                const bool ev = ok && ((info >> 10) & 1);
                double e = ok ? exp(eta - emax) : 0.0;
                double S0 = warp_scan_incl(e, lane) + c0;
                c0 = __shfl_sync(GT_FULL, S0, 31);
                if (!((info >> 11) & 1)) {      // no tie group in this block (warp-uniform)
                    if (ev) Ttot += 1.0 / S0;
                    continue;
                }
                double PF = warp_scan_incl(ev ? e : 0.0, lane);
                int PD = warp_scan_incl_int(ev ? 1 : 0, lane);
                double S0g = __shfl_sync(GT_FULL, S0, lb);
                double PFb = __shfl_sync(GT_FULL, PF, lb), PFa = __shfl_sync(GT_FULL, PF, (la + 31) & 31);
                int PDb = __shfl_sync(GT_FULL, PD, lb), PDa = __shfl_sync(GT_FULL, PD, (la + 31) & 31);
                if (la == 0) { PFa = 0.0; PDa = 0; }
                if (ev) {
                    double c = (double)(PD - 1 - PDa) / (double)(PDb - PDa);
                    Ttot += 1.0 / (S0g - c * (PFb - PFa));
                }
            }
            Ttot = gt_warp_sum(Ttot);
        }
        // ---- pass 2: log-likelihood, score, Hessian ----
        double c0 = 0.0, cI = 0.0;
        for (int t = lane; t < PH; t += 32) carry[t] = 0.0;
        __syncwarp();
        for (int b = seg_b0; b < seg_b1; ++b) {
            const int nbig = T.blk_info[b];
            if (nbig > 1) {
                // ---- tied events over several blocks (all members are events) ----
                // totals of the group first: EG = sum e, d, S1 at its end
                double *s1_before = coef;             // free during the passes
                for (int a = lane; a < p; a += 32) s1_before[a] = carry[a];
                __syncwarp();
                double EG = 0.0; int dg = 0;
                for (int bb = b; bb < b + nbig; ++bb) {
                    double eta; int info;
                    bool ok = load_row(bb * 32 + lane, eta, info);
                    double e = ok ? exp(eta - emax) : 0.0;
                    EG += e; dg += ok ? 1 : 0;
                    for (int a = 0; a < p; ++a) {
                        double z = gt_warp_sum(e * xT[a * XT + lane]);
                        if (lane == 0) carry[a] += z;
                    }
                    __syncwarp();
                }
                EG = gt_warp_sum(EG); dg = gt_warp_sum_int(dg);
                const double S0g = c0 + EG;
                // the d Efron terms, 32 at a time: u_j = (S1g - c_j F1g) / den_j
                double Ag = 0.0, Cg = 0.0;
                for (int j0 = 0; j0 < dg; j0 += 32) {
                    const int j = j0 + lane;
                    const bool valid = j < dg;
                    const double cj = valid ? (double)j / (double)dg : 0.0;
                    const double den = S0g - cj * EG;
                    const double inv = valid ? 1.0 / den : 0.0;
                    Ag += inv; Cg += cj * inv;
                    if (valid && final_eval) ll -= log(den);
                    for (int a = 0; a < p; ++a) {
                        const double s1g = carry[a];
                        uT[a * XT + lane] = (s1g - cj * (s1g - s1_before[a])) * inv;
                    }
                    wv[lane] = 0.0; gv[lane] = 0.0;
                    mma_block();
                }
                Ag = gt_warp_sum(Ag); Cg = gt_warp_sum(Cg);
                // the members: W_i = e_i (sum of 1/den over the terms at or after
                // the group - sum_j c_j / den_j)
                const double Mg = (Ttot - cI) - Cg;
                for (int bb = b; bb < b + nbig; ++bb) {
                    double eta; int info;
                    bool ok = load_row(bb * 32 + lane, eta, info);
                    const double Wi = ok ? exp(eta - emax) * Mg : 0.0;
                    if (ok && final_eval) ll += eta - emax;
                    wv[lane] = Wi; gv[lane] = (ok ? 1.0 : 0.0) - Wi;
                    for (int a = 0; a < p; ++a) uT[a * XT + lane] = 0.0;
                    mma_block();
                }
                c0 = S0g; cI += Ag;
                b += nbig - 1;
                continue;
            }
            double eta; int info;
            bool ok = load_row(b * 32 + lane, eta, info);
            const int la = info & 31, lb = (info >> 5) & 31, lam1 = (la + 31) & 31;
            const bool ev = ok && ((info >> 10) & 1);
            const bool ties = (info >> 11) & 1;
            double e = ok ? exp(eta - emax) : 0.0;
            double S0 = warp_scan_incl(e, lane) + c0;
            c0 = __shfl_sync(GT_FULL, S0, 31);
            // without a tie group in the block (warp-uniform flag) every
            // sample is its own group: c = 0, den = S0 of its own lane
            double c = 0.0, inv = 0.0, den = S0;
            if (ties) {
                double PF = warp_scan_incl(ev ? e : 0.0, lane);
                int PD = warp_scan_incl_int(ev ? 1 : 0, lane);
                double S0g = __shfl_sync(GT_FULL, S0, lb);
                double PFb = __shfl_sync(GT_FULL, PF, lb), PFa = __shfl_sync(GT_FULL, PF, lam1);
                int PDb = __shfl_sync(GT_FULL, PD, lb), PDa = __shfl_sync(GT_FULL, PD, lam1);
                if (la == 0) { PFa = 0.0; PDa = 0; }
                if (ev) {
                    c = (double)(PD - 1 - PDa) / (double)(PDb - PDa);
                    den = S0g - c * (PFb - PFa);
                }
            }
            if (ev) {
                inv = 1.0 / den;
                if (final_eval) ll += (eta - emax) - log(den);    // only the last pass reports it
            }
            double PI = warp_scan_incl(inv, lane) + cI;
            double PIa = __shfl_sync(GT_FULL, PI, lam1);
            if (la == 0) PIa = cI;
            cI = __shfl_sync(GT_FULL, PI, 31);
            double tie_term = 0.0;
            if (ties) {
                double PCI = warp_scan_incl(c * inv, lane);
                double PCIb = __shfl_sync(GT_FULL, PCI, lb), PCIa = __shfl_sync(GT_FULL, PCI, lam1);
                if (la == 0) PCIa = 0.0;
                if (ev) tie_term = PCIb - PCIa;
            }
            double Wi = e * ((Ttot - PIa) - tie_term);
            double Gi = (ev ? 1.0 : 0.0) - Wi;
            double *xr = xT + lane, *ur = uT + lane;
            wv[lane] = Wi; gv[lane] = Gi;
            // running S1 (risk set) and F1 (tied events of the group), CV
            // design columns at a time so that their scans overlap
            constexpr int CV = 4;
            for (int a0 = 0; a0 < p; a0 += CV) {
                double z[CV], cr[CV];
#pragma unroll
                for (int j = 0; j < CV; ++j) {
                    const int a = a0 + j < p ? a0 + j : p - 1;
                    z[j] = e * xr[a * XT];
                    cr[j] = carry[a];
                }
                double s1[CV];
#pragma unroll
                for (int j = 0; j < CV; ++j) s1[j] = z[j];
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    double t[CV];
#pragma unroll
                    for (int j = 0; j < CV; ++j) t[j] = __shfl_up_sync(GT_FULL, s1[j], o);
                    if (lane >= o) {
#pragma unroll
                        for (int j = 0; j < CV; ++j) s1[j] += t[j];
                    }
                }
                double f1g[CV];
#pragma unroll
                for (int j = 0; j < CV; ++j) { s1[j] += cr[j]; f1g[j] = ev ? z[j] : 0.0; }
                if (ties) {
                    double f1[CV];
#pragma unroll
                    for (int j = 0; j < CV; ++j) f1[j] = f1g[j];
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        double t[CV];
#pragma unroll
                        for (int j = 0; j < CV; ++j) t[j] = __shfl_up_sync(GT_FULL, f1[j], o);
                        if (lane >= o) {
#pragma unroll
                            for (int j = 0; j < CV; ++j) f1[j] += t[j];
                        }
                    }
#pragma unroll
                    for (int j = 0; j < CV; ++j) {
                        double fb = __shfl_sync(GT_FULL, f1[j], lb), fa = __shfl_sync(GT_FULL, f1[j], lam1);
                        if (la == 0) fa = 0.0;
                        f1g[j] = fb - fa;
                    }
                }
                __syncwarp();              // every lane has read carry[a0 ..]
#pragma unroll
                for (int j = 0; j < CV; ++j) {
                    const double s1g = ties ? __shfl_sync(GT_FULL, s1[j], lb) : s1[j];
                    if (a0 + j < p) {
                        ur[(a0 + j) * XT] = ev ? (s1g - c * f1g[j]) * inv : 0.0;
                        if (lane == 31) carry[a0 + j] = s1[j];
                    }
                }
            }
            mma_block();
        }
        }   // strata
        ll = gt_warp_sum(ll);
        // -H = accH[:p,:p] - accU ; score = accH[:, p]
#pragma unroll
        for (int mt = 0; mt < NT; ++mt)
#pragma unroll
            for (int nt = mt; nt < NT; ++nt) {
                int r = mt * 8 + (lane >> 2), cc = nt * 8 + 2 * (lane & 3);
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    double v = accH[mt][nt][j];
                    if (cc + j < p) v -= accU[mt][nt][j];
                    H[r * PH + cc + j] = v;
                }
            }
        __syncwarp();
        const double dn = (double)nj;
        if (final_eval) {
            llf = ll;
            for (int t = lane; t < p * p; t += 32) {
                int i = t / p, j = t % p;
                Hc[i * p + j] = (i <= j) ? H[i * PH + j] : H[j * PH + i];
            }
            __syncwarp();
            if (!gt_warp_cholesky(Hc, p, p, lane)) { status = GT_SINGULAR; break; }
            gt_warp_chol_inverse(Hc, p, p, Hinv, lane);
            break;
        }
        // Newton step on loglike / nobs with the reference's ridge
        for (int t = lane; t < p * p; t += 32) {
            int i = t / p, j = t % p;
            double v = ((i <= j) ? H[i * PH + j] : H[j * PH + i]) / dn;
            if (i == j) v -= 1e-10;
            Hc[i * p + j] = v;
        }
        for (int i = lane; i < p; i += 32) coef[i] = H[i * PH + p] / dn;     // score / nobs
        __syncwarp();
        if (!gt_warp_cholesky(Hc, p, p, lane)) { status = GT_SINGULAR; break; }
        gt_warp_chol_inverse(Hc, p, p, Hinv, lane);
        for (int i = lane; i < p; i += 32) {
            double s = 0.0;
            for (int j = 0; j < p; ++j) s = fma(Hinv[i * p + j], coef[j], s);
            step[i] = s;
        }
        __syncwarp();
        // convergence is judged on the reference's parameterisation: T step
        double worst = 0.0;
        for (int i = lane; i < p; i += 32) {
            double d = 0.0;
            if (i < k) { for (int a = i; a < k; ++a) d = fma(T.Rinv[i * k + a], step[a], d); }
            else d = step[i];
            worst = fmax(worst, fabs(d));
        }
        worst = gt_warp_max(worst);
        double grow = 0.0;
        for (int i = lane; i < p; i += 32) grow = fma(fabs(step[i]), T.colmax[i], grow);
        grow = gt_warp_sum(grow);
        emax += grow; slack += grow;
        for (int i = lane; i < p; i += 32) beta[i] += step[i];
        __syncwarp();
        niter += 1;
        if (!(worst == worst)) { status = GT_NAN_PVALUE; break; }
        // statsmodels warns (-> StatsError, survival.py:87-88) whenever the step count
        // reaches maxiter, also when that very step met the tolerance
        if (niter >= T.newton_maxiter) { status = GT_NO_CONVERGENCE; break; }
        if (!(worst > 1e-8)) { final_eval = true; continue; }
    }
    if (status != GT_OK) { write_fail(status, flip, maf, nan(""), niter); return; }

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

// Warps take markers from a shared counter until none is left: the iteration
// count varies from marker to marker, and a static assignment would leave
// the last wave of CTAs partly empty.
template <bool PACKED, int NT>
__global__ void __launch_bounds__(256)
cox_newton_kernel(IterDesignDev T, const void *__restrict__ geno, long long pitch, int n_snps,
                  double *__restrict__ out, int *__restrict__ iout, long long ostride,
                  int *__restrict__ next_marker, int only_deferred) {
    extern __shared__ __align__(16) double sm[];
    const int lane = threadIdx.x & 31;
    for (;;) {
        int snp = 0;
        if (lane == 0) snp = atomicAdd(next_marker, 1);
        snp = __shfl_sync(GT_FULL, snp, 0);
        if (snp >= n_snps) return;
        // second pass behind cox_thread_kernel: only the markers it left unfinished
        if (only_deferred && iout[(long long)GT_I_STATUS * ostride + snp] != -77) continue;
        cox_newton_one<PACKED, NT>(T, geno, pitch, n_snps, out, iout, ostride, snp, sm);
        __syncwarp();
    }
}

}  // namespace gtk
