// Host side of the iterative models (logistic IRLS, Cox Newton): design upload
// and batch dispatch.  The Cox kernel itself (K5) lives below.
#pragma once

#include <functional>
#include <string>
#include <vector>

#include "gt_common.cuh"
#include "gt_host.cuh"
#include "gt_logistic.cuh"
#include "gt_logistic_thread.cuh"
#include "gt_coxk.cuh"
#include "gt_cox_thread.cuh"

namespace gtk {

// Uploads, into one device allocation:  X [n_pad][XS] | y [n_pad] | Rinv [k*k]
static int iter_set_design(const gt_design_desc *d, const std::vector<int> &pos, int n_file,
                           int n_chunks, int64_t pitch, const uint8_t *dmask, DevBuf &buf, DevBuf &bufT,
                           IterDesignDev &T, cudaStream_t st, std::string &err) {
    (void)pitch; (void)dmask; (void)err;
    const int n = d->n, k = d->k, I = d->n_inter;
    const int p = k + 1 + I;
    if (p > 23) return fail(GT_E_UNSUPPORTED, "iterative models support at most 23 design columns (got %d)", p);
    HostDesign H;
    H.n = n; H.k = k; H.I = I;
    qr_mgs2(d->C, n, k, H);
    const size_t n_pad = (size_t)n_chunks * 256;
    const int XS = (k + I + 1) & ~1;
    const int XSs = XS > 0 ? XS : 2;
    // (k * k rounded up to even: what follows stays 16 B aligned)
    const size_t kk2 = ((size_t)k * k + 1) & ~(size_t)1;
    std::vector<double> host(n_pad * XSs + n_pad + kk2 + 2, 0.0);
    double *X = host.data(), *y = X + n_pad * XSs, *Rinv = y + n_pad;
    for (size_t i = 0; i < n_pad; ++i) y[i] = nan("");
    for (int i = 0; i < n; ++i) {
        double *xr = X + (size_t)pos[i] * XSs;
        for (int j = 0; j < k; ++j) xr[j] = H.Q[(size_t)j * n + i];
        for (int j = 0; j < I; ++j) xr[k + j] = d->W[(size_t)j * n + i];
        y[pos[i]] = d->y[i];
    }
    for (int t = 0; t < k * k; ++t) Rinv[t] = H.Rinv[t];
    // coxph: schedule in descending time.  At one time point the censored
    // samples come first, each on its own (they only have to be in the risk
    // set of the events tied with them); the events of that time form one
    // Efron group: inside one 32-slot block when there are at most 32, else
    // over whole blocks of their own ("big group": blk_info = block count on
    // the first block).
    std::vector<int> slot_pos, slot_info, blk_info, seg_start;
    std::vector<double> Xs;
    if (d->model == GT_MODEL_COXPH) {
        std::vector<int> order(n);
        for (int i = 0; i < n; ++i) order[i] = i;
        // strata first (PHReg(strata=...): the risk sets restart with every stratum), then time
        const double *strata = d->strata;
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
            if (strata && strata[a] != strata[b]) return strata[a] < strata[b];
            return d->y[a] > d->y[b];
        });
        auto pad_block = [&]() {
            for (int fill = (int)(slot_pos.size() % 32); fill != 0 && fill < 32; ++fill) {
                slot_pos.push_back(-1);
                slot_info.push_back(fill | (fill << 5));
            }
        };
        auto push = [&](int row, int la, int lb, int ev) {
            slot_pos.push_back(pos[row]);
            slot_info.push_back(la | (lb << 5) | (ev << 10));
        };
        std::vector<int> big_first, big_blocks;
        int i0 = 0;
        seg_start.push_back(0);
        while (i0 < n) {
            if (strata && i0 > 0 && strata[order[i0]] != strata[order[i0 - 1]]) {
                pad_block();                             // a stratum starts on a block of its own
                seg_start.push_back((int)(slot_pos.size() / 32));
            }
            int i1 = i0;
            while (i1 + 1 < n && d->y[order[i1 + 1]] == d->y[order[i0]] &&
                   (!strata || strata[order[i1 + 1]] == strata[order[i0]])) ++i1;
            std::vector<int> events;
            for (int t = i0; t <= i1; ++t) {
                int row = order[t];
                if (d->event[row] != 0.0) { events.push_back(row); continue; }
                int l = (int)(slot_pos.size() % 32);
                push(row, l, l, 0);
            }
            const int m = (int)events.size();
            if (m > 32) {
                pad_block();
                big_first.push_back((int)(slot_pos.size() / 32));
                big_blocks.push_back((m + 31) / 32);
                for (int row : events) { int l = (int)(slot_pos.size() % 32); push(row, l, l, 1); }
                pad_block();
            } else if (m > 0) {
                if ((int)(slot_pos.size() % 32) + m > 32) pad_block();
                int la = (int)(slot_pos.size() % 32), lb = la + m - 1;
                for (int row : events) push(row, la, lb, 1);
            }
            i0 = i1 + 1;
        }
        pad_block();
        const size_t ns = slot_pos.size();
        seg_start.push_back((int)(ns / 32));
        for (size_t b = 0; b < ns; b += 32) {
            bool ties = false;
            for (int t = 0; t < 32; ++t) ties |= ((slot_info[b + t] & 31) != ((slot_info[b + t] >> 5) & 31));
            if (ties) for (int t = 0; t < 32; ++t) slot_info[b + t] |= 1 << 11;
        }
        blk_info.assign(ns / 32, 1);
        for (size_t g = 0; g < big_first.size(); ++g) {
            blk_info[big_first[g]] = big_blocks[g];
            for (int t = 1; t < big_blocks[g]; ++t) blk_info[big_first[g] + t] = -1;
        }
        Xs.assign(ns * XSs, 0.0);
        // slot -> design row: rebuild from the sorted order
        size_t s = 0;
        std::vector<int> row_of_pos(n_file, -1);
        for (int i = 0; i < n; ++i) row_of_pos[pos[i]] = i;
        for (s = 0; s < ns; ++s) {
            if (slot_pos[s] < 0) continue;
            int row = row_of_pos[slot_pos[s]];
            for (int j = 0; j < k; ++j) Xs[s * XSs + j] = H.Q[(size_t)j * n + row];
            for (int j = 0; j < I; ++j) Xs[s * XSs + k + j] = d->W[(size_t)j * n + row];
        }
        // append: Xs | slot_pos | slot_info | blk_info | seg_start  (ints packed two per double)
        size_t base = host.size();
        host.resize(base + Xs.size() + ns + ns / 64 + seg_start.size() / 2 + 3, 0.0);
        std::memcpy(&host[base], Xs.data(), Xs.size() * 8);
        int *ip = (int *)&host[base + Xs.size()];
        std::memcpy(ip, slot_pos.data(), ns * 4);
        std::memcpy(ip + ns, slot_info.data(), ns * 4);
        std::memcpy(ip + 2 * ns, blk_info.data(), (ns / 32) * 4);
        std::memcpy(ip + 2 * ns + ns / 32, seg_start.data(), seg_start.size() * 4);
        // max |x| per design column (q | g | g w): bounds the change of the
        // linear predictor after a Newton step
        size_t cm = host.size();
        host.resize(cm + p, 0.0);
        for (int j = 0; j < k; ++j)
            for (int i = 0; i < n; ++i) host[cm + j] = std::max(host[cm + j], std::fabs(H.Q[(size_t)j * n + i]));
        host[cm + k] = 2.0;
        for (int j = 0; j < I; ++j)
            for (int i = 0; i < n; ++i)
                host[cm + k + 1 + j] = std::max(host[cm + k + 1 + j], 2.0 * std::fabs(d->W[(size_t)j * n + i]));
        // pointers may have moved
        X = host.data(); y = X + n_pad * XSs; Rinv = y + n_pad;
        T.n_slots = (int)ns;
        // ---- schedule of the thread-per-SNP kernel (gt_cox_thread.cuh): the design rows in
        //      descending time, stratum by stratum (each padded to whole 256-row blocks), the
        //      censored samples of a time point before its events; per row the event flag, the
        //      position inside a group of tied events and the size of that group ----
        T.th_ok = 0;
        if (I == 0 && !H.degenerate && p >= 2 && p <= 12) {
            const int XST = (k + 3) & ~1;                 // q (k) | flags | group size
            std::vector<double> rowsT;
            std::vector<int> posT, segblk;
            auto push_row = [&](int row, int flags, int dsz) {
                const size_t base = rowsT.size();
                rowsT.resize(base + XST, 0.0);
                for (int j = 0; j < k; ++j) rowsT[base + j] = H.Q[(size_t)j * n + row];
                rowsT[base + k] = (double)flags;
                rowsT[base + k + 1] = (double)dsz;
                posT.push_back(pos[row]);
            };
            auto pad256 = [&]() {
                while (posT.size() % 256) { posT.push_back(-1); rowsT.resize(rowsT.size() + XST, 0.0); }
            };
            bool any_tied = false;
            int j0 = 0;
            segblk.push_back(0);
            while (j0 < n) {
                if (strata && j0 > 0 && strata[order[j0]] != strata[order[j0 - 1]]) {
                    pad256();
                    segblk.push_back((int)(posT.size() / 256));
                }
                int j1 = j0;
                while (j1 + 1 < n && d->y[order[j1 + 1]] == d->y[order[j0]] &&
                       (!strata || strata[order[j1 + 1]] == strata[order[j0]])) ++j1;
                int m = 0;
                for (int t = j0; t <= j1; ++t) {
                    if (d->event[order[t]] != 0.0) ++m;
                    else push_row(order[t], 0, 0);
                }
                int idx = 0;
                for (int t = j0; t <= j1; ++t) {
                    if (d->event[order[t]] == 0.0) continue;
                    // 1 event, 2 member of a group of tied events, 4 first / 8 last member
                    const int fl = 1 | (m >= 2 ? 2 : 0) | (m >= 2 && idx == 0 ? 4 : 0) | (m >= 2 && idx == m - 1 ? 8 : 0);
                    push_row(order[t], fl, m);
                    ++idx;
                }
                any_tied |= m >= 2;
                j0 = j1 + 1;
            }
            pad256();
            segblk.push_back((int)(posT.size() / 256));
            const size_t nT = posT.size();
            const int nblkT = (int)(nT / 256), nsegT = (int)segblk.size() - 1;
            std::vector<int> ordT;                          // blocks in the order a round consumes them
            for (int sg = 0; sg < nsegT; ++sg)
                for (int rep = 0; rep < 2; ++rep)
                    for (int bb = segblk[sg]; bb < segblk[sg + 1]; ++bb) ordT.push_back(bb);
            const size_t ints = nT + ordT.size() + segblk.size();
            std::vector<double> ht(nT * XST + (ints + 1) / 2 + 2, 0.0);
            std::memcpy(ht.data(), rowsT.data(), nT * XST * 8);
            int *tp = (int *)&ht[nT * XST];
            std::memcpy(tp, posT.data(), nT * 4);
            std::memcpy(tp + nT, ordT.data(), ordT.size() * 4);
            std::memcpy(tp + nT + ordT.size(), segblk.data(), segblk.size() * 4);
            int rct = upload(bufT, ht.data(), ht.size() * 8, st);
            if (rct) return rct;
            CK(cudaStreamSynchronize(st));
            T.th_ok = 1; T.th_n = (int)nT; T.th_XS = XST; T.th_tied = any_tied ? 1 : 0; T.th_nseg = nsegT;
            T.th_X = (const double *)bufT.p;
            T.th_pos = (const int *)((const double *)bufT.p + nT * XST);
            T.th_order = T.th_pos + nT;
            T.th_seg = T.th_order + ordT.size();
            (void)nblkT;
        }
    }
    int rc = upload(buf, host.data(), host.size() * 8, st);
    if (rc) return rc;
    T.n_file = n_file; T.n = n; T.k = k; T.I = I; T.p = p; T.XS = XSs; T.n_pad = (int)n_pad;
    T.raw_mode = d->raw_mode; T.use_maf_t = d->use_maf_t; T.maf_t = d->maf_t;
    T.degenerate = H.degenerate ? 1 : 0;
    T.y_binary = 1;
    for (int i = 0; i < n && T.y_binary; ++i) T.y_binary = d->y[i] == 0.0 || d->y[i] == 1.0;
    T.maf_o = nullptr; T.flip_o = nullptr;
    T.X = (const double *)buf.p; T.y = T.X + n_pad * XSs; T.Rinv = T.y + n_pad;
    if (d->model == GT_MODEL_COXPH) {
        const double *xs = T.Rinv + kk2 + 2;
        T.Xs = xs;
        T.slot_pos = (const int *)(xs + (size_t)T.n_slots * XSs);
        T.slot_info = T.slot_pos + T.n_slots;
        T.blk_info = T.slot_info + T.n_slots;
        T.seg_start = T.blk_info + T.n_slots / 32;
        T.n_seg = (int)seg_start.size() - 1;
        T.colmax = (const double *)buf.p + (host.size() - p);
    } else {
        T.th_ok = 0;
        T.n_slots = 0; T.Xs = nullptr; T.slot_pos = nullptr; T.slot_info = nullptr; T.blk_info = nullptr; T.colmax = nullptr;
        T.n_seg = 0; T.seg_start = nullptr;
    }
    CK(cudaStreamSynchronize(st));
    return 0;
}

template <bool PACKED, int NT>
static int launch_logistic(const IterDesignDev &T, const void *geno, int64_t pitch, int n_snps,
                      double *out, int32_t *iout, long long ldo, cudaStream_t st, int *counter,
                      int only_deferred = 0) {
    using Cfg = LogitCfg<NT>;
    const int wpc = 8;
    size_t smem = (size_t)wpc * Cfg::perWarp * 8;   // NT = 2: 92 KB -> two CTAs per SM
    CK(cudaFuncSetAttribute((const void *)logistic_irls_kernel<PACKED, NT>,
                            cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // (static assignment: the IRLS step count hardly varies between markers;
    // a work counter measured 3 % slower here)
    (void)counter;
    logistic_irls_kernel<PACKED, NT><<<(n_snps + wpc - 1) / wpc, wpc * 32, smem, st>>>(
        T, geno, pitch, n_snps, out, iout, ldo, only_deferred);
    CK(cudaGetLastError());
    return 0;
}

// thread-per-SNP IRLS (gt_logistic_thread.cuh): whole waves of 148 CTAs where the batch allows
template <int P>
static int launch_logistic_thread(const IterDesignDev &T, const void *geno, int64_t pitch, int n_snps,
                                  double *out, int32_t *iout, long long ldo, cudaStream_t st,
                                  double *scratch, long long lds, int maxpass, int fold) {
    // markers per CTA: the batch spread over the 148 SMs (whole warps)
    const int cap = (fold || T.I == 1) ? kLtFoldThreads : kLtMaxThreads;
    const int waves = (n_snps + 148 * cap - 1) / (148 * cap);
    int nthr = ((n_snps + 148 * waves - 1) / (148 * waves) + 31) / 32 * 32;
    nthr = std::max(32, std::min(nthr, cap));
    const int XSs = std::max(2, (P & ~1));                       // (K + 1) & ~1 with K = P - 1
    const size_t smem = 128 + (size_t)kLtBuf * (kLtBlk * XSs * 8 + kLtBlk * 8) + 128 +
                        (size_t)(((P - 1) * (P - 1) + 1) & ~1) * 8 + (size_t)P * nthr * 8;
    if (T.I == 1) {
        // one interaction column: design row [q | g | g w] (always without a producer warp)
        if constexpr (P >= 3) {
            CK(cudaFuncSetAttribute((const void *)logistic_thread_kernel<P, 1, 1>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            logistic_thread_kernel<P, 1, 1><<<(n_snps + nthr - 1) / nthr, nthr, smem, st>>>(
                T, (const uint8_t *)geno, pitch, n_snps, out, iout, ldo, scratch, lds, maxpass);
        }
    } else if (fold) {
        CK(cudaFuncSetAttribute((const void *)logistic_thread_kernel<P, 1>,
                                cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        logistic_thread_kernel<P, 1><<<(n_snps + nthr - 1) / nthr, nthr, smem, st>>>(
            T, (const uint8_t *)geno, pitch, n_snps, out, iout, ldo, scratch, lds, maxpass);
    } else {
        CK(cudaFuncSetAttribute((const void *)logistic_thread_kernel<P, 0>,
                                cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        logistic_thread_kernel<P, 0><<<(n_snps + nthr - 1) / nthr, nthr + 32, smem, st>>>(
            T, (const uint8_t *)geno, pitch, n_snps, out, iout, ldo, scratch, lds, maxpass);
    }
    CK(cudaGetLastError());
    return 0;
}

template <bool PACKED, int NT>
static int launch_cox(const IterDesignDev &T, const void *geno, int64_t pitch, int n_snps,
                      double *out, int32_t *iout, long long ldo, cudaStream_t st, int *counter,
                      int only_deferred = 0) {
    using Cfg = CoxCfg<NT>;
    // two CTAs per SM: as many warps as fit half of the shared memory
    int wpc = 8;
    while (wpc > 1 && (size_t)wpc * Cfg::perWarp(T.p) * 8 > 112 * 1024) --wpc;
    size_t smem = (size_t)wpc * Cfg::perWarp(T.p) * 8;
    CK(cudaFuncSetAttribute((const void *)cox_newton_kernel<PACKED, NT>,
                            cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // persistent warps: two CTAs per SM, markers handed out by a counter
    CK(cudaMemsetAsync(counter, 0, sizeof(int), st));
    const int grid = std::min((n_snps + wpc - 1) / wpc, 148 * 2);
    cox_newton_kernel<PACKED, NT><<<grid, wpc * 32, smem, st>>>(
        T, geno, pitch, n_snps, out, iout, ldo, counter, only_deferred);
    CK(cudaGetLastError());
    return 0;
}

// thread-per-SNP Newton solver (gt_cox_thread.cuh): warp 0 feeds the ring (no producer warp)
template <int P>
static int launch_cox_thread(const IterDesignDev &T, const void *geno, int64_t pitch, int n_snps,
                             double *out, int32_t *iout, long long ldo, cudaStream_t st,
                             uint32_t *gperm, long long ldw, int maxround) {
    const int cap = kLtFoldThreads;
    const int waves = (n_snps + 148 * cap - 1) / (148 * cap);
    int nthr = ((n_snps + 148 * waves - 1) / (148 * waves) + 31) / 32 * 32;
    nthr = std::max(32, std::min(nthr, cap));
    const int XST = (P + 2) & ~1;                               // (K + 3) & ~1 with K = P - 1
    const size_t smem = 128 + (size_t)kLtBuf * (kLtBlk * XST * 8) + 128 +
                        (size_t)((((P - 1) * (P - 1) + 1) & ~1) + ((P + 1) & ~1)) * 8 +
                        (size_t)(T.th_tied ? 3 : 1) * P * nthr * 8;
    {
        // the codes of every marker in schedule order, word-major
        const int row_words = (T.n_file + 15) / 16;
        int spw = row_words | 1;                                  // odd: the markers of a word hit distinct banks
        int mpc = 32;
        while (mpc > 1 && (size_t)mpc * spw * 4 > 200 * 1024) mpc >>= 1;
        const size_t psm = (size_t)mpc * spw * 4;
        if (psm > 220 * 1024) return fail(GT_E_UNSUPPORTED, "thread-per-SNP Cox kernel: %d samples per row", T.n_file);
        CK(cudaFuncSetAttribute((const void *)cox_permute_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)psm));
        cox_permute_kernel<<<(n_snps + mpc - 1) / mpc, 256, psm, st>>>(
            (const uint8_t *)geno, pitch, n_snps, row_words, T.th_pos, T.th_n, gperm, ldw, mpc, spw);
        CK(cudaGetLastError());
    }
    if (T.th_tied) {
        CK(cudaFuncSetAttribute((const void *)cox_thread_kernel<P, 1, 1>,
                                cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        cox_thread_kernel<P, 1, 1><<<(n_snps + nthr - 1) / nthr, nthr, smem, st>>>(
            T, (const uint8_t *)geno, pitch, n_snps, out, iout, ldo, gperm, ldw, maxround);
    } else {
        CK(cudaFuncSetAttribute((const void *)cox_thread_kernel<P, 1, 0>,
                                cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        cox_thread_kernel<P, 1, 0><<<(n_snps + nthr - 1) / nthr, nthr, smem, st>>>(
            T, (const uint8_t *)geno, pitch, n_snps, out, iout, ldo, gperm, ldw, maxround);
    }
    CK(cudaGetLastError());
    return 0;
}

static int iter_run(const IterDesignDev &T, int model, const void *geno, int64_t pitch, bool packed,
                    int32_t n_snps, double *out, int32_t *iout, long long ldo, cudaStream_t st, int *counter, bool timing,
                    double *lt_scratch, bool use_thread, int thread_min, int thread_fold,
                    const std::function<int()> &tbegin, const std::function<void(int)> &tend,
                    int64_t &all_launches, std::string &err) {
    (void)timing; (void)err;
    int NT = (T.p + 1 + 7) / 8;
    int slot = tbegin();
    int rc;
    if (model == GT_MODEL_COXPH) {
        if (packed) {
            // .bed rows, no strata / interaction / tied event times, p <= 12, enough markers:
            // one thread per marker first, the warp kernel finishes what that one deferred
            int deferred = 0;
            if (use_thread && lt_scratch && T.th_ok && T.p >= 2 && T.p <= 12 && n_snps >= thread_min) {
                const long long ldw = (n_snps + 31) / 32 * 32;
                rc = 0;
#define GT_CT(PP) case PP: rc = launch_cox_thread<PP>(T, geno, pitch, n_snps, out, iout, ldo, st, (uint32_t *)lt_scratch, ldw, 8); break;
                switch (T.p) {
                    GT_CT(2) GT_CT(3) GT_CT(4) GT_CT(5) GT_CT(6) GT_CT(7) GT_CT(8) GT_CT(9) GT_CT(10)
                    GT_CT(11) GT_CT(12)
                }
#undef GT_CT
                if (rc) return rc;
                deferred = 1;
                all_launches += 1;
            }
            rc = NT == 1 ? launch_cox<true, 1>(T, geno, pitch, n_snps, out, iout, ldo, st, counter, deferred)
               : NT == 2 ? launch_cox<true, 2>(T, geno, pitch, n_snps, out, iout, ldo, st, counter, deferred)
                         : launch_cox<true, 3>(T, geno, pitch, n_snps, out, iout, ldo, st, counter, deferred);
        } else {
            rc = NT == 1 ? launch_cox<false, 1>(T, geno, pitch, n_snps, out, iout, ldo, st, counter)
               : NT == 2 ? launch_cox<false, 2>(T, geno, pitch, n_snps, out, iout, ldo, st, counter)
                         : launch_cox<false, 3>(T, geno, pitch, n_snps, out, iout, ldo, st, counter);
        }
    } else if (packed) {
        // .bed rows, binary outcome, no interaction, p <= 13, enough markers to fill the SMs:
        // one thread per marker first, the warp kernel then only finishes what that one deferred
        int deferred = 0;
        rc = 0;
        if (use_thread && lt_scratch && T.I <= 1 && T.p >= 2 + T.I && T.p <= 13 && T.y_binary && !T.degenerate &&
            n_snps >= thread_min) {
            const long long lds = (n_snps + 31) / 32 * 32;
            const int maxpass = 9;
#define GT_LT(PP) case PP: rc = launch_logistic_thread<PP>(T, geno, pitch, n_snps, out, iout, ldo, st, lt_scratch, lds, maxpass, thread_fold); break;
            switch (T.p) {
                GT_LT(2) GT_LT(3) GT_LT(4) GT_LT(5) GT_LT(6) GT_LT(7) GT_LT(8) GT_LT(9) GT_LT(10)
                GT_LT(11) GT_LT(12) GT_LT(13)
            }
#undef GT_LT
            if (rc) return rc;
            deferred = 1;
            all_launches += 1;
        }
        rc = NT == 1 ? launch_logistic<true, 1>(T, geno, pitch, n_snps, out, iout, ldo, st, counter, deferred)
           : NT == 2 ? launch_logistic<true, 2>(T, geno, pitch, n_snps, out, iout, ldo, st, counter, deferred)
                     : launch_logistic<true, 3>(T, geno, pitch, n_snps, out, iout, ldo, st, counter, deferred);
    } else {
        rc = NT == 1 ? launch_logistic<false, 1>(T, geno, pitch, n_snps, out, iout, ldo, st, counter)
           : NT == 2 ? launch_logistic<false, 2>(T, geno, pitch, n_snps, out, iout, ldo, st, counter)
                     : launch_logistic<false, 3>(T, geno, pitch, n_snps, out, iout, ldo, st, counter);
    }
    tend(slot);
    all_launches += 1;
    return rc;
}

}  // namespace gtk
