// K3: per-SNP masked normal-equation finish of the linear model, one warp per
// SNP.  Restates what sm.OLS(y, X).fit() + StatsLinear._results_handler
// (genetest/statistics/models/linear.py:46-104) compute for the complete
// cases of each marker (analysis.py:119,172), after _gwas_worker's MAF /
// flip / interaction steps (analysis.py:127-166):
//   * re-scan the marker's row for missing samples and subtract their
//     v v' contributions from the full-sample Gram matrix (exact
//     complete-case sums),
//   * MAF on the complete cases, flip (g -> 2 - g) applied algebraically,
//   * p x p normal equations in the orthonormal covariate basis (Cholesky),
//     mapped back through R^-1 to the reference's parameterisation,
//   * condition-number / eigenvalue gates (certified trace bound first,
//     Jacobi eigenvalues only when the bound cannot decide),
//   * t statistics, two-sided p-values, 95 % CI, adjusted R^2, log-lik.
#pragma once

#include "gt_common.cuh"

namespace gtk {

// stride of the 32-entry staging tile: == 4 (mod 16) doubles -> the MMA
// fragment loads below are bank-conflict free
__host__ __device__ constexpr int finish_stage_stride(int NT2) {
    return NT2 <= 2 ? 20 : (NT2 <= 4 ? 36 : 52);
}

struct FinishLayout {
    int perWarp;   // doubles per warp
    int oS0, oMo, oMc, oMi, oStage, oVec, oQueue;
    int queueInts;
};

__host__ __device__ inline FinishLayout finish_layout(int L, int p, int I) {
    FinishLayout f;
    int o = 0;
    f.oS0 = o; o += L * L;
    f.oMo = o; o += p * p;
    f.oMc = o; o += p * p;
    f.oMi = o; o += p * p;
    o = (o + 1) & ~1;                      // staging rows are copied 16 B at a time
    f.oStage = o; o += 32 * finish_stage_stride((L + 7) / 8);
    f.oVec = o; o += 6 * p + (I + 1) * L + (I + 1) * (I + 1) + 8;
    f.oQueue = o;
    f.queueInts = 64;
    o += (f.queueInts + 1) / 2;
    f.perWarp = (o + 1) & ~1;
    return f;
}

__device__ __forceinline__ void finish_write_fail(double *out, int *iout, long long n_snps, int snp,
                                                  int nf, int status, int nobs, int flip,
                                                  double maf, double aux, int lane) {
    for (int f = lane; f < nf; f += 32) {
        double v = nan("");
        if (f == GT_F_MAF) v = maf;
        if (f == GT_F_AUX) v = aux;
        out[(long long)f * n_snps + snp] = v;
    }
    if (lane == 0) {
        iout[(long long)GT_I_STATUS * n_snps + snp] = status;
        iout[(long long)GT_I_NOBS * n_snps + snp] = nobs;
        iout[(long long)GT_I_FLIP * n_snps + snp] = flip;
        iout[(long long)GT_I_NITER * n_snps + snp] = 0;
    }
}

__device__ __forceinline__ void dmma884_lin(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// Subtract sum over the missing samples of v v' from S0 (one warp).
// PACKED: row = 2-bit codes OR-ed with the design mask; else doubles with NaN.
// Missing samples are compacted 32 at a time (ballot ranking), their base
// vectors gathered into a shared staging tile, and the outer products
// accumulated with FP64 tensor-core MMAs (m8n8k4, upper tiles only).
template <bool PACKED, int NT2>
__device__ __forceinline__ void subtract_missing(const GtDesignDev &D, const void *rowp,
                                                 double *S0, double *stage, int *queue,
                                                 int lane, const int *mlist, int mcountA,
                                                 int mcountB, int hcap) {
    constexpr int LS = finish_stage_stride(NT2);
    const int L = D.L;
    const int nlive = D.lean ? D.k + 1 : L;      // columns present in E
    double acc[NT2][NT2][2];
#pragma unroll
    for (int a = 0; a < NT2; ++a)
#pragma unroll
        for (int b = 0; b < NT2; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;
    int qn = 0;
    const unsigned lt_mask = (1u << lane) - 1u;

    auto process = [&](int count) {     // entries queue[0 .. count), count <= 32
        int idx = (lane < count) ? queue[lane] : -1;
        double *sr = stage + lane * LS;
        if (idx >= 0) {
            // E rows are 16 B aligned (NCP even), staging rows too (LS even)
            const double2 *er = (const double2 *)(D.E + (long long)idx * D.NCP);
            double2 *s2 = (double2 *)sr;
            for (int c = 0; c < (nlive + 1) / 2; ++c) s2[c] = er[c];
            if (!D.lean && (nlive & 1)) sr[nlive] = 0.0;    // keep the padding column zero
            if (D.lean) sr[D.k + 1] = 1.0;
        } else {
            for (int c = 0; c < L; ++c) sr[c] = 0.0;
        }
        __syncwarp();
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
            const int e = ks * 4 + (lane & 3), c = lane >> 2;
            double f[NT2];
#pragma unroll
            for (int t = 0; t < NT2; ++t) f[t] = stage[e * LS + t * 8 + c];
#pragma unroll
            for (int mt = 0; mt < NT2; ++mt)
#pragma unroll
                for (int nt = mt; nt < NT2; ++nt)
                    dmma884_lin(acc[mt][nt][0], acc[mt][nt][1], f[mt], f[nt]);
        }
        __syncwarp();
    };
    auto push_round = [&](bool have, int idx) {
        unsigned has = __ballot_sync(GT_FULL, have);
        if (have) queue[qn + __popc(has & lt_mask)] = idx;
        qn += __popc(has);
        __syncwarp();
        if (qn >= 32) {
            process(32);
            int rem = qn - 32;
            int v = (lane < rem) ? queue[32 + lane] : 0;
            __syncwarp();
            if (lane < rem) queue[lane] = v;
            qn = rem;
            __syncwarp();
        }
        return has;
    };

    if (mlist != nullptr && mcountA >= 0 && mcountB >= 0) {
        // The contraction kernel already listed the missing samples.  Each
        // lane loads its own MMA fragment elements straight from E (no
        // shared staging): lane (c = lane / 4, e = lane % 4) of k-step ks
        // needs elements t * 8 + c of sample 4 ks + e, so the eight lanes of
        // one sample read 64 contiguous bytes -- whole 32 B sectors.  Group
        // g + 1 is in flight while the tensor cores work on group g; the
        // indices run two groups ahead.
        const int mcount = mcountA + mcountB;
        const int ngroups = (mcount + 31) / 32;
        const int c = lane >> 2, e4 = lane & 3;
        const int one_col = D.lean ? D.k + 1 : -1;
        auto index_of = [&](int g) {
            int t = g * 32 + lane;
            return (t < mcount) ? (t < mcountA ? mlist[t] : mlist[hcap + t - mcountA]) : -1;
        };
        // every load is unconditional: lanes without an element (padding
        // columns, samples past the end of the list) read the constant 0.0,
        // the lane of the ones column reads the constant 1.0 -- a predicated
        // load would be followed by a select that waits for it
        long long off[NT2];
#pragma unroll
        for (int t = 0; t < NT2; ++t) {
            const int el = t * 8 + c;
            off[t] = el < nlive ? (long long)el : (el == one_col ? -2 : -1);
        }
        const double *zero_one = D.zero_one;
        auto gather = [&](int idx_lane, double (&f)[8 * NT2]) {
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) {
                const int idx = __shfl_sync(GT_FULL, idx_lane, ks * 4 + e4);
                const double *er = D.E + (long long)idx * D.NCP;
#pragma unroll
                for (int t = 0; t < NT2; ++t) {
                    const double *ptr = er + off[t];
                    if (off[t] < 0) ptr = zero_one + (off[t] == -2 ? 1 : 0);
                    if (idx < 0) ptr = zero_one;
                    f[ks * NT2 + t] = __ldg(ptr);
                }
            }
        };
        auto mma = [&](const double (&f)[8 * NT2]) {
#pragma unroll
            for (int ks = 0; ks < 8; ++ks)
#pragma unroll
                for (int mt = 0; mt < NT2; ++mt)
#pragma unroll
                    for (int nt = mt; nt < NT2; ++nt)
                        dmma884_lin(acc[mt][nt][0], acc[mt][nt][1], f[ks * NT2 + mt], f[ks * NT2 + nt]);
        };
        if (NT2 <= 2) {
            // branch-free software pipeline: groups past the end gather zeros
            // (index -1), so no conditional copy sits between a load and the
            // MMAs two stages later
            double fa[8 * NT2], fb[8 * NT2];
            int i1 = index_of(1), i2 = index_of(2);
            gather(index_of(0), fa);
            for (int g = 0; g < ngroups; g += 2) {
                gather(i1, fb);
                i1 = index_of(g + 3);
                mma(fa);
                gather(i2, fa);
                i2 = index_of(g + 4);
                mma(fb);
            }
        } else {
            double fa[8 * NT2];
            int inext = index_of(0);
            for (int g = 0; g < ngroups; ++g) {
                gather(inext, fa);
                inext = index_of(g + 1);
                mma(fa);
            }
        }
    } else if (PACKED) {
        const int nwords = (D.n_file + 15) / 16;         // 32-bit words of 16 samples
        const int nblocks = (nwords + 127) / 128;        // 4 words (64 samples) per lane
        const uint4 *row4 = (const uint4 *)rowp;
        const uint4 *mask4 = (const uint4 *)D.ormask;
        auto fetch = [&](int blk) {
            uint4 w = make_uint4(~0u, ~0u, ~0u, ~0u);
            if (blk < nblocks && (blk * 32 + lane) * 4 < nwords) {   // rows are 16 B padded
                uint4 r = row4[blk * 32 + lane], m = mask4[blk * 32 + lane];
                w = make_uint4(r.x | m.x, r.y | m.y, r.z | m.z, r.w | m.w);
            }
            return w;
        };
        uint4 nxt = fetch(0);
        for (int blk = 0; blk < nblocks; ++blk) {
            const uint4 w = nxt;
            nxt = fetch(blk + 1);                 // prefetch: hides the DRAM latency of the scan
            unsigned long long lo = (unsigned long long)(w.x & ~(w.x >> 1) & 0x55555555u) |
                                    ((unsigned long long)(w.y & ~(w.y >> 1) & 0x55555555u) << 32);
            unsigned long long hi = (unsigned long long)(w.z & ~(w.z >> 1) & 0x55555555u) |
                                    ((unsigned long long)(w.w & ~(w.w >> 1) & 0x55555555u) << 32);
            const int base = (blk * 32 + lane) * 64;
            if (!__any_sync(GT_FULL, (lo | hi) != 0ull)) continue;
            while (true) {
                bool have = (lo | hi) != 0ull;
                int idx = 0;
                if (lo) { int b = __ffsll((long long)lo) - 1; idx = base + (b >> 1); lo &= lo - 1; }
                else if (hi) { int b = __ffsll((long long)hi) - 1; idx = base + 32 + (b >> 1); hi &= hi - 1; }
                if (!push_round(have, idx)) break;
            }
        }
    } else {
        const int nblocks = (D.n + 31) / 32;
        for (int blk = 0; blk < nblocks; ++blk) {
            int i = blk * 32 + lane;
            double g = (i < D.n) ? ((const double *)rowp)[i] : 0.0;
            push_round(g != g, i);
        }
    }
    if (qn > 0) process(qn);
    __syncwarp();
#pragma unroll
    for (int mt = 0; mt < NT2; ++mt)
#pragma unroll
        for (int nt = mt; nt < NT2; ++nt)
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                int r = mt * 8 + (lane >> 2), c = nt * 8 + 2 * (lane & 3) + j;
                if (r < L && c < L && r <= c) {
                    double v = acc[mt][nt][j];
                    S0[r * L + c] -= v;
                    if (r != c) S0[c * L + r] -= v;
                }
            }
    __syncwarp();
}

// K3a: missing-sample corrections from the lists the tcgen05 contraction
// emitted.  One warp per SNP, register-light (high occupancy hides the two
// dependent loads index -> base vector); indices run two groups ahead, base
// vectors one.  Writes Corr[snp][L*L] = sum over the missing samples of v v'.
template <int NT2>
__global__ void __launch_bounds__(256, 4)
missing_corr_kernel(GtDesignDev D, int n_snps, const int *__restrict__ mlist, int mcap,
                    const int *__restrict__ mcount, double *__restrict__ corr) {
    constexpr int LS = finish_stage_stride(NT2);
    constexpr int NV = 4 * NT2;
    extern __shared__ __align__(16) double sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int snp = blockIdx.x * (blockDim.x >> 5) + warp;
    if (snp >= n_snps) return;
    const int L = D.L;
    const int nlive = D.lean ? D.k + 1 : L;
    const int nv = (nlive + 1) / 2;
    double *stage = sm + (size_t)warp * 32 * LS;
    const int mA = mcount[(long long)snp * 2], mB = mcount[(long long)snp * 2 + 1];
    double *out = corr + (long long)snp * L * L;
    if (mA < 0 || mB < 0) {          // a list overflowed: the finish kernel re-scans the row
        if (lane == 0) out[0] = nan("");
        return;
    }
    const int *ml = mlist + (long long)snp * mcap;
    const int hcap = mcap >> 1, total = mA + mB;
    double acc[NT2][NT2][2];
#pragma unroll
    for (int a = 0; a < NT2; ++a)
#pragma unroll
        for (int b = 0; b < NT2; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;
    for (int t = lane; t < 32 * LS; t += 32) stage[t] = 0.0;
    auto index_of = [&](int g) {
        int t = g * 32 + lane;
        return (t < total) ? (t < mA ? ml[t] : ml[hcap + t - mA]) : -1;
    };
    double2 regs[NV];
    auto fetch = [&](int idx) {
        const double2 *er = (const double2 *)(D.E + (long long)(idx < 0 ? 0 : idx) * D.NCP);
#pragma unroll
        for (int c = 0; c < NV; ++c) regs[c] = (idx >= 0 && c < nv) ? er[c] : make_double2(0.0, 0.0);
    };
    const int ngroups = (total + 31) / 32;
    int idx_cur = ngroups > 0 ? index_of(0) : -1;
    int idx_nxt = ngroups > 1 ? index_of(1) : -1;
    if (ngroups > 0) fetch(idx_cur);
    __syncwarp();
    for (int g = 0; g < ngroups; ++g) {
        double2 *s2 = (double2 *)(stage + lane * LS);
#pragma unroll
        for (int c = 0; c < NV; ++c) if (2 * c < LS) s2[c] = regs[c];
        if (idx_cur >= 0) {
            if (D.lean) stage[lane * LS + D.k + 1] = 1.0;
            else if (nlive & 1) stage[lane * LS + nlive] = 0.0;
        }
        __syncwarp();
        idx_cur = idx_nxt;
        if (g + 1 < ngroups) fetch(idx_cur);
        idx_nxt = (g + 2 < ngroups) ? index_of(g + 2) : -1;
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
            const int e = ks * 4 + (lane & 3), c = lane >> 2;
            double f[NT2];
#pragma unroll
            for (int t = 0; t < NT2; ++t) f[t] = stage[e * LS + t * 8 + c];
#pragma unroll
            for (int mt = 0; mt < NT2; ++mt)
#pragma unroll
                for (int nt = mt; nt < NT2; ++nt)
                    dmma884_lin(acc[mt][nt][0], acc[mt][nt][1], f[mt], f[nt]);
        }
        __syncwarp();
    }
#pragma unroll
    for (int mt = 0; mt < NT2; ++mt)
#pragma unroll
        for (int nt = mt; nt < NT2; ++nt)
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                int r = mt * 8 + (lane >> 2), c = nt * 8 + 2 * (lane & 3) + j;
                if (r < L && c < L && r <= c) {
                    double v = acc[mt][nt][j];
                    out[r * L + c] = v;
                    if (r != c) out[c * L + r] = v;
                }
            }
}

template <bool PACKED, int NT2>
__global__ void __launch_bounds__(256, 2)
linear_finish_kernel(GtDesignDev D, const void *__restrict__ geno, long long pitch, int n_snps,
                     const double *__restrict__ S, const int *__restrict__ counts,
                     const double *__restrict__ gsum, const int *__restrict__ mlist, int mcap,
                     const int *__restrict__ mcount, const double *__restrict__ corr,
                     double *__restrict__ out, int *__restrict__ iout, long long ostride,
                     int wpc) {
    extern __shared__ __align__(16) double sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int k = D.k, I = D.I, L = D.L, p = D.p;
    const int iy = k, i1 = k + 1;
    const int nf = GT_F_PARAM0 + 6 * p;
    const FinishLayout lay = finish_layout(L, p, I);

    const int snp = blockIdx.x * wpc + warp;
    if (warp >= wpc || snp >= n_snps) return;

    double *W = sm + (size_t)warp * lay.perWarp;
    double *S0 = W + lay.oS0, *Mo = W + lay.oMo, *Mc = W + lay.oMc, *Mi = W + lay.oMi;
    double *stage = W + lay.oStage;
    double *rhs = W + lay.oVec, *bq = rhs + p, *beta = bq + p, *var = beta + p,
           *gdiag = var + p, *tmpv = gdiag + p;
    double *S1 = tmpv + p;                  // [(I+1)][L]
    double *S2 = S1 + (I + 1) * L;          // [(I+1)][(I+1)]
    int *queue = (int *)(W + lay.oQueue);

    for (int t = lane; t < L * L; t += 32) S0[t] = D.FF0[t];
    for (int t = lane; t < 32 * finish_stage_stride(NT2); t += 32) stage[t] = 0.0;
    __syncwarp();

    const int n_het = counts[(long long)snp * 4 + 0];
    const int n_hom = counts[(long long)snp * 4 + 1];
    const int n_miss = counts[(long long)snp * 4 + 2];
    // corrections already accumulated by missing_corr_kernel (NaN marks a list overflow)
    bool have_corr = false;
    if (corr != nullptr && n_miss > 0) {
        const double *cs = corr + (long long)snp * L * L;
        have_corr = (cs[0] == cs[0]);
        if (have_corr)
            for (int t = lane; t < L * L; t += 32) S0[t] -= cs[t];
        __syncwarp();
    }
    if (n_miss > 0 && !have_corr) {
        const void *rowp = PACKED ? (const void *)((const uint8_t *)geno + (long long)snp * pitch)
                                  : (const void *)((const double *)geno + (long long)snp * pitch);
        const int mA = mlist ? mcount[(long long)snp * 2] : -1;
        const int mB = mlist ? mcount[(long long)snp * 2 + 1] : -1;
        subtract_missing<PACKED, NT2>(D, rowp, S0, stage, queue, lane,
                                      mlist ? mlist + (long long)snp * mcap : nullptr, mA, mB,
                                      mcap >> 1);
    }
    __syncwarp();

    const double *Ssnp = S + (long long)snp * D.NCP;
    const double sum_g_all = gsum[(long long)snp * 2], sum_g2_all = gsum[(long long)snp * 2 + 1];
    if (D.lean) {      // E carries [q | y] only; sum g, sum g^2 come from the counts
        for (int t = lane; t < L; t += 32) S1[t] = (t <= k) ? Ssnp[t] : sum_g_all;
        if (lane == 0) S2[0] = sum_g2_all;
    } else {
        for (int t = lane; t < (I + 1) * L; t += 32) S1[t] = Ssnp[t];
        for (int t = lane; t < (I + 1) * (I + 1); t += 32) {
            int a = t / (I + 1), b = t % (I + 1);
            int lo = a < b ? a : b, hi = a < b ? b : a;
            int idx = lo * (I + 1) - lo * (lo - 1) / 2 + (hi - lo);
            S2[t] = Ssnp[D.NCA + idx];
        }
    }
    __syncwarp();

    const int nj = D.n - n_miss;
    if (nj <= 0) {
        finish_write_fail(out, iout, ostride, snp, nf, GT_ALL_MISSING, 0, 0, nan(""), nan(""), lane);
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
            finish_write_fail(out, iout, ostride, snp, nf, GT_MAF_BELOW, nj, flip, maf, maf, lane);
            return;
        }
    }
    if (flip) {   // g -> 2 - g on the complete cases (analysis.py:156-157)
        for (int t = lane; t < (I + 1) * (I + 1); t += 32) {
            int a = t / (I + 1), b = t % (I + 1);
            int ca = a == 0 ? i1 : k + 1 + a, cb = b == 0 ? i1 : k + 1 + b;
            stage[t] = 4.0 * S0[ca * L + cb] - 4.0 * S1[a * L + cb] + S2[t];
        }
        __syncwarp();
        bool all_zero = (stage[0] == 0.0);
        for (int t = lane; t < (I + 1) * L; t += 32) {
            int c = t / L, b = t % L;
            int cc = c == 0 ? i1 : k + 1 + c;
            S1[t] = all_zero ? 0.0 : 2.0 * S0[cc * L + b] - S1[t];
        }
        __syncwarp();
        for (int t = lane; t < (I + 1) * (I + 1); t += 32) S2[t] = all_zero ? 0.0 : stage[t];
        __syncwarp();
    }

    // normal equations in the [Q | g | g w] basis
    for (int t = lane; t < p * p; t += 32) {
        int i = t / p, j = t % p;
        double v;
        if (i < k && j < k) v = S0[i * L + j];
        else if (i < k) v = S1[(j - k) * L + i];
        else if (j < k) v = S1[(i - k) * L + j];
        else v = S2[(i - k) * (I + 1) + (j - k)];
        Mo[t] = v; Mc[t] = v;
    }
    for (int i = lane; i < p; i += 32) rhs[i] = (i < k) ? S0[i * L + iy] : S1[(i - k) * L + iy];
    const double yy = S0[iy * L + iy], sy = S0[iy * L + i1];
    __syncwarp();

    bool chol_ok = !D.degenerate && gt_warp_cholesky(Mc, p, p, lane);
    bool need_jacobi = !chol_ok;
    double ssr = 0.0;
    if (chol_ok) {
        // M = Lc Lc'.  Only Z = Lc^-1 is formed (one forward substitution per
        // column): with u = Z rhs, bq = Z' u and ssr = yy - u'u; back in the
        // reference's parameterisation beta = T bq and var_i = |column i of
        // Z T'|^2 with T = blockdiag(R^-1, I).
        double *Z = Mi, *invd = tmpv, *u = gdiag, *W2 = Mc;
        for (int i = lane; i < p; i += 32) invd[i] = 1.0 / Mc[i * p + i];
        for (int t = lane; t < p * p; t += 32) Z[t] = 0.0;
        __syncwarp();
        for (int c = lane; c < p; c += 32)
            for (int i = c; i < p; ++i) {
                double s = (i == c) ? 1.0 : 0.0;
                for (int j = c; j < i; ++j) s = fma(-Mc[i * p + j], Z[j * p + c], s);
                Z[i * p + c] = s * invd[i];
            }
        __syncwarp();
        double part = 0.0;
        for (int r = lane; r < p; r += 32) {
            double s = 0.0;
            for (int a = 0; a <= r; ++a) s = fma(Z[r * p + a], rhs[a], s);
            u[r] = s;
            part = fma(s, s, part);
        }
        ssr = yy - gt_warp_sum(part);
        __syncwarp();
        for (int i = lane; i < p; i += 32) {
            double s = 0.0;
            for (int r = i; r < p; ++r) s = fma(Z[r * p + i], u[r], s);
            bq[i] = s;
        }
        // squares of W = Z T' (the Cholesky factor is not needed any more)
        for (int t = lane; t < p * p; t += 32) {
            const int r = t / p, i = t - r * p;
            double w;
            if (i < k) {
                w = 0.0;
                const int hi = r < k - 1 ? r : k - 1;
                for (int a = i; a <= hi; ++a) w = fma(Z[r * p + a], D.Rinv[i * k + a], w);
            } else {
                w = Z[r * p + i];
            }
            W2[t] = w * w;
        }
        __syncwarp();
        double tv = 0.0, tg = 0.0;
        for (int i = lane; i < p; i += 32) {
            double v = 0.0, b;
            for (int r = 0; r < p; ++r) v += W2[r * p + i];
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
        __syncwarp();
        tv = gt_warp_sum(tv);
        // trace of the Gram matrix of the original columns: dropping samples
        // only lowers the covariates' diagonal, so |R|_F^2 bounds it
        tg = gt_warp_sum(tg) + D.trG0;
        // cond^2 <= trace(G) trace(G^-1),  lambda_min >= 1 / trace(G^-1)
        bool cond_safe = tg * tv <= D.cond_t * D.cond_t * (1.0 - 1e-9);
        bool eig_safe = 1.0 / tv >= D.eig_t * (1.0 + 1e-9);
        need_jacobi = !(cond_safe && eig_safe);
    }
    if (need_jacobi) {
        // G = Ti' Mo Ti with Ti = blockdiag(R, I): Gram matrix of the original columns
        double *tmp = stage, *G = S0;
        for (int t = lane; t < p * p; t += 32) {
            int i = t / p, j = t % p;
            double s;
            if (j < k) { s = 0.0; for (int a = 0; a <= j; ++a) s = fma(Mo[i * p + a], D.R[a * k + j], s); }
            else s = Mo[i * p + j];
            tmp[t] = s;
        }
        __syncwarp();
        for (int t = lane; t < p * p; t += 32) {
            int i = t / p, j = t % p;
            double s;
            if (i < k) { s = 0.0; for (int a = 0; a <= i; ++a) s = fma(D.R[a * k + i], tmp[a * p + j], s); }
            else s = tmp[i * p + j];
            G[i * p + j] = s;
        }
        __syncwarp();
        // symmetrise (rounding) so a zero column stays exactly zero on both sides
        for (int t = lane; t < p * p; t += 32) {
            int i = t / p, j = t % p;
            if (i < j) { double v = G[j * p + i]; G[i * p + j] = v; }
        }
        __syncwarp();
        double emax, emin;
        gt_warp_jacobi_extremes(G, p, p, lane, &emax, &emin);
        double cond = (emin > 0.0) ? sqrt(emax / emin) : INFINITY;
        if (D.degenerate) cond = INFINITY;
        if (cond > D.cond_t) {
            finish_write_fail(out, iout, ostride, snp, nf, GT_COND_LARGE, nj, flip, maf, cond, lane);
            return;
        }
        if (emin < D.eig_t) {
            finish_write_fail(out, iout, ostride, snp, nf, GT_EIG_SMALL, nj, flip, maf, emin, lane);
            return;
        }
        if (!chol_ok) {
            finish_write_fail(out, iout, ostride, snp, nf, GT_COND_LARGE, nj, flip, maf, INFINITY, lane);
            return;
        }
    }

    const int df = nj - p;
    const double scale = (df > 0) ? ssr / (double)df : nan("");
    const double q975 = (df > 0) ? D.tq[df] : nan("");
    double pv_snp = 0.0;
    double my[6];
    for (int j = lane; j < p; j += 32) {
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
    pv_snp = __shfl_sync(GT_FULL, pv_snp, k & 31);
    int status = GT_OK;
    if (!D.raw_mode && pv_snp != pv_snp) status = GT_NAN_PVALUE;
    if (lane == 0) {
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
