// K3a: missing-sample corrections of the linear model.  Every marker has its
// own set of missing samples; the exact complete-case Gram matrix is the
// full-sample one minus the sum of v v' over those samples (analysis.py:119,
// 172: statsmodels fits the complete cases).  One warp per SNP accumulates
// that sum with FP64 tensor-core MMAs; gt_solve.cuh (K3b) does the rest.
#pragma once

#include "gt_common.cuh"

namespace gtk {

// stride of the 32-entry staging tile: == 4 (mod 16) doubles -> the MMA
// fragment loads below are bank-conflict free
__host__ __device__ constexpr int finish_stage_stride(int NT2) {
    return NT2 <= 2 ? 20 : (NT2 <= 4 ? 36 : 52);
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
__device__ __forceinline__ int subtract_missing(const GtDesignDev &D, const void *rowp,
                                                double *S0, double *stage, int *queue,
                                                int lane, const uint2 *mrec, int nrec) {
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

    int n_listed = 0;
    if (mrec != nullptr) {
        // The contraction kernel recorded every 32-sample pair that has missing
        // genotypes as (pair index, flags); flag bit b is sample
        // 32 * pair + 16 * (b & 1) + (b >> 1).  The records are turned into
        // sample indices in the warp's queue (the staging tile is not needed on
        // this path, the queue takes its place), one round of up to kQCap
        // indices at a time, and every round runs the gather / MMA pipeline:
        // each lane loads its own MMA fragment elements straight from E (no
        // shared staging): lane (c = lane / 4, e = lane % 4) of k-step ks
        // needs elements t * 8 + c of sample 4 ks + e, so the eight lanes of
        // one sample read 64 contiguous bytes -- whole 32 B sectors.  Group
        // g + 1 is in flight while the tensor cores work on group g.
        constexpr int kQCap = 32 * finish_stage_stride(NT2) * 2 + 64;   // ints in stage + queue
        int *rq = (int *)stage;                  // the queue follows the staging tile (corr_layout)
        const int c = lane >> 2, e4 = lane & 3;
        const int one_col = D.lean ? D.k + 1 : -1;
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
        int r = 0;
        while (r < nrec) {
            // ---- records -> indices, until the queue cannot take another 32 records ----
            int qn2 = 0;
            while (r < nrec) {
                const int t = r + lane;
                uint2 rec = make_uint2(0u, 0u);
                if (t < nrec) rec = mrec[t];
                uint32_t f = rec.y;
                int incl = __popc(f);
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    int v = __shfl_up_sync(GT_FULL, incl, o);
                    if (lane >= o) incl += v;
                }
                const int total = __shfl_sync(GT_FULL, incl, 31);
                if (qn2 + total > kQCap) break;               // run what is queued first
                int w = qn2 + incl - __popc(f);
                const int base = (int)rec.x * 32;
                while (f) {
                    const int b = __ffs((int)f) - 1;
                    rq[w++] = base + 16 * (b & 1) + (b >> 1);
                    f &= f - 1;
                }
                qn2 += total;
                r += 32;
            }
            __syncwarp();
            n_listed += qn2;
            const int ngroups = (qn2 + 31) / 32;
            auto index_of = [&](int g) {
                int t = g * 32 + lane;
                return (t < qn2) ? rq[t] : -1;
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
            __syncwarp();
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
    return n_listed;
}

// K3a: the missing-sample corrections of a batch.  One warp per SNP; writes
// Corr[snp][L*L] = sum over the missing samples of v v' for every marker that
// has missing genotypes (the solve kernel reads it only for those), as the upper
// triangle in SoA layout corrT[entry][marker].  The
// samples come from the lists the tcgen05 contraction emitted, or from a
// re-scan of the row (FP64 contraction, dosages, an overflowed list).
struct CorrLayout { int perWarp, oS0, oStage, oQueue; };
__host__ __device__ inline CorrLayout corr_layout(int L) {
    CorrLayout f;
    int o = 0;
    f.oS0 = o; o += (L * L + 1) & ~1;
    f.oStage = o; o += 32 * finish_stage_stride((L + 7) / 8);
    f.oQueue = o; o += 32;                       // 64 ints
    f.perWarp = o;
    return f;
}

template <bool PACKED, int NT2>
__global__ void __launch_bounds__(256, 2)
missing_corr_kernel(GtDesignDev D, const void *__restrict__ geno, long long pitch, int n_snps,
                    int *__restrict__ counts, const uint2 *__restrict__ mrec, int rcap,
                    const int *__restrict__ mcount, double *__restrict__ corrT, long long ldc) {
    extern __shared__ __align__(16) double sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int snp = blockIdx.x * (blockDim.x >> 5) + warp;
    if (snp >= n_snps) return;
    const int mA = mrec ? mcount[snp] : 0;
    if (mrec ? (mA == 0) : (counts[(long long)snp * 4 + 2] <= 0)) return;   // nothing missing
    const int L = D.L;
    const CorrLayout lay = corr_layout(L);
    double *W = sm + (size_t)warp * lay.perWarp;
    double *S0 = W + lay.oS0, *stage = W + lay.oStage;
    int *queue = (int *)(W + lay.oQueue);
    for (int t = lane; t < L * L; t += 32) S0[t] = 0.0;
    if (!mrec) for (int t = lane; t < 32 * finish_stage_stride(NT2); t += 32) stage[t] = 0.0;
    __syncwarp();
    const void *rowp = PACKED ? (const void *)((const uint8_t *)geno + (long long)snp * pitch)
                              : (const void *)((const double *)geno + (long long)snp * pitch);
    const int n_listed = subtract_missing<PACKED, NT2>(
        D, rowp, S0, stage, queue, lane, mrec ? mrec + (long long)snp * rcap : nullptr, mA);
    // the tcgen05 contraction leaves the missing count to this kernel
    if (mrec && lane == 0) counts[(long long)snp * 4 + 2] = n_listed;
    // upper triangle, SoA over the markers: entry (a <= b) at [a L - a (a - 1) / 2 + (b - a)][snp]
    for (int t = lane; t < L * L; t += 32) {
        const int a = t / L, b = t % L;
        if (a <= b) corrT[(long long)(a * L - a * (a - 1) / 2 + (b - a)) * ldc + snp] = -S0[t];
    }
}

}  // namespace gtk
