// Host side of the iterative models (logistic IRLS, Cox Newton): design upload
// and batch dispatch.  The Cox kernel itself (K5) lives below.
#pragma once

#include <functional>
#include <string>
#include <vector>

#include "gt_common.cuh"
#include "gt_host.cuh"
#include "gt_logistic.cuh"

namespace gtk {

// Uploads, into one device allocation:  X [n_pad][XS] | y [n_pad] | Rinv [k*k]
static int iter_set_design(const gt_design_desc *d, const std::vector<int> &pos, int n_file,
                           int n_chunks, int64_t pitch, const uint8_t *dmask, DevBuf &buf,
                           IterDesignDev &T, cudaStream_t st, std::string &err) {
    (void)pitch; (void)dmask; (void)err;
    const int n = d->n, k = d->k, I = d->n_inter;
    const int p = k + 1 + I;
    if (p > 23) return fail(GT_E_UNSUPPORTED, "iterative models support at most 23 design columns (got %d)", p);
    if (d->model == GT_MODEL_COXPH)
        return fail(GT_E_UNSUPPORTED, "coxph kernel not built in this revision");
    HostDesign H;
    H.n = n; H.k = k; H.I = I;
    qr_mgs2(d->C, n, k, H);
    if (H.degenerate) return fail(GT_E_UNSUPPORTED, "covariates are rank deficient (logistic / coxph need a full-rank design)");
    const size_t n_pad = (size_t)n_chunks * 256;
    const int XS = (k + I + 1) & ~1;
    const int XSs = XS > 0 ? XS : 2;
    std::vector<double> host(n_pad * XSs + n_pad + (size_t)k * k + 2, 0.0);
    double *X = host.data(), *y = X + n_pad * XSs, *Rinv = y + n_pad;
    for (size_t i = 0; i < n_pad; ++i) y[i] = nan("");
    for (int i = 0; i < n; ++i) {
        double *xr = X + (size_t)pos[i] * XSs;
        for (int j = 0; j < k; ++j) xr[j] = H.Q[(size_t)j * n + i];
        for (int j = 0; j < I; ++j) xr[k + j] = d->W[(size_t)j * n + i];
        y[pos[i]] = d->y[i];
    }
    for (int t = 0; t < k * k; ++t) Rinv[t] = H.Rinv[t];
    int rc = upload(buf, host.data(), host.size() * 8, st);
    if (rc) return rc;
    T.n_file = n_file; T.n = n; T.k = k; T.I = I; T.p = p; T.XS = XSs; T.n_pad = (int)n_pad;
    T.raw_mode = d->raw_mode; T.use_maf_t = d->use_maf_t; T.maf_t = d->maf_t;
    T.X = (const double *)buf.p; T.y = T.X + n_pad * XSs; T.Rinv = T.y + n_pad;
    T.order = nullptr; T.group_end = nullptr; T.tte = nullptr;
    CK(cudaStreamSynchronize(st));
    return 0;
}

template <bool PACKED, int NT>
static int launch_logistic(const IterDesignDev &T, const void *geno, int64_t pitch, int n_snps,
                           double *out, int32_t *iout, cudaStream_t st) {
    using Cfg = LogitCfg<NT>;
    const int wpc = 8;
    size_t smem = (size_t)wpc * Cfg::perWarp * 8;
    CK(cudaFuncSetAttribute((const void *)logistic_irls_kernel<PACKED, NT>,
                            cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    logistic_irls_kernel<PACKED, NT><<<(n_snps + wpc - 1) / wpc, wpc * 32, smem, st>>>(
        T, geno, pitch, n_snps, out, iout, (long long)n_snps);
    CK(cudaGetLastError());
    return 0;
}

static int iter_run(const IterDesignDev &T, int model, const void *geno, int64_t pitch, bool packed,
                    int32_t n_snps, double *out, int32_t *iout, cudaStream_t st, bool timing,
                    const std::function<int()> &tbegin, const std::function<void(int)> &tend,
                    int64_t &all_launches, std::string &err) {
    (void)timing; (void)err;
    if (model != GT_MODEL_LOGISTIC) return fail(GT_E_UNSUPPORTED, "coxph kernel not built in this revision");
    int NT = (T.p + 1 + 7) / 8;
    int slot = tbegin();
    int rc;
    if (packed) {
        rc = NT == 1 ? launch_logistic<true, 1>(T, geno, pitch, n_snps, out, iout, st)
           : NT == 2 ? launch_logistic<true, 2>(T, geno, pitch, n_snps, out, iout, st)
                     : launch_logistic<true, 3>(T, geno, pitch, n_snps, out, iout, st);
    } else {
        rc = NT == 1 ? launch_logistic<false, 1>(T, geno, pitch, n_snps, out, iout, st)
           : NT == 2 ? launch_logistic<false, 2>(T, geno, pitch, n_snps, out, iout, st)
                     : launch_logistic<false, 3>(T, geno, pitch, n_snps, out, iout, st);
    }
    tend(slot);
    all_launches += 1;
    return rc;
}

}  // namespace gtk
