// K1 + K2: stage PLINK 2-bit rows into shared memory with cp.async, decode
// them to dosages in registers and contract them against the design's E
// matrix in FP64.  Output per SNP: S[NCP] = sum_i g_i E1_i | sum_i g_i^2 E2_i,
// the exact integer genotype counts and (sum g, sum g^2).
//
// Replaces (for a tile of 32*ROWS SNPs at once) the per-SNP staging of
// genetest/analysis.py:109-119 (NaN fill, label-aligned scatter, complete-case
// mask) and the X'X / X'y accumulation inside sm.OLS(...).fit()
// (statistics/models/linear.py:102-104).
#pragma once

#include "gt_common.cuh"

namespace gtk {

constexpr int kChunk = 256;            // samples per pipeline stage
constexpr int kRowBytes = kChunk / 4;  // 64 packed bytes per SNP row per stage
constexpr int kRowPitch = kRowBytes + 16;  // shared-memory row pitch (bank spread)

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

template <int NCA, int NCB, int ROWS, int WARPS>
struct ContractCfg {
    static constexpr int NC = NCA + NCB;
    static constexpr int kThreads = 32 * WARPS;
    static constexpr int kSPW = kChunk / WARPS;                      // samples per warp per stage
    static constexpr int kTile = 32 * ROWS;                          // SNPs per CTA
    static constexpr int kTabBytes = 64;                             // decode table (4 x double2)
    static constexpr int kGenoBytes = kTile * kRowPitch;
    static constexpr int kEBytes = kChunk * NC * 8;
    static constexpr int kStageBytes = kGenoBytes + kEBytes + kRowBytes;
    static constexpr int kStages = (2 * kStageBytes + kTabBytes <= 72 * 1024) ? 2
                                   : ((2 * kStageBytes + kTabBytes <= 200 * 1024) ? 2 : 1);
    static constexpr int kRedBytes = WARPS * 32 * NC * 8 + WARPS * 32 * 3 * 4;
    static constexpr int kMain = kStages * kStageBytes;
    static constexpr int kSmem = kTabBytes + (kMain > kRedBytes ? kMain : kRedBytes);
};

// Packed kernel.  `lean` designs (no interaction) carry only [q | y] in E:
// sum g and sum g^2 come from the exact class counts.
template <int NCA, int NCB, int ROWS, int WARPS, int MINB>
__global__ void __launch_bounds__(32 * WARPS, MINB)
contract_packed_kernel(const uint8_t *__restrict__ geno, long long pitch, int n_snps,
                       const uint8_t *__restrict__ ormask, const double *__restrict__ E,
                       int n_chunks, double *__restrict__ S, int *__restrict__ counts,
                       double *__restrict__ gsum) {
    using Cfg = ContractCfg<NCA, NCB, ROWS, WARPS>;
    constexpr int NC = Cfg::NC;
    constexpr int kThreads = Cfg::kThreads;
    constexpr int NW64 = Cfg::kSPW / 32;     // 64-bit words (32 samples) per row per warp
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // code -> (g, g^2): 00 -> 2, 01 (missing) -> 0, 10 -> 1, 11 -> 0
    double2 *tab = (double2 *)smem_raw;
    unsigned char *smem = smem_raw + Cfg::kTabBytes;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tile0 = blockIdx.x * Cfg::kTile;
    if (tid < 4) {
        double g = tid == 0 ? 2.0 : (tid == 2 ? 1.0 : 0.0);
        tab[tid] = make_double2(g, g * g);
    }

    auto stage_geno = [&](int s) { return smem + s * Cfg::kStageBytes; };
    auto stage_E = [&](int s) { return (double *)(smem + s * Cfg::kStageBytes + Cfg::kGenoBytes); };
    auto stage_mask = [&](int s) { return smem + s * Cfg::kStageBytes + Cfg::kGenoBytes + Cfg::kEBytes; };

    auto issue = [&](int chunk, int s) {
        unsigned char *gs = stage_geno(s);
        for (int t = tid; t < Cfg::kTile * 4; t += kThreads) {
            int row = t >> 2, part = t & 3;
            int snp = min(tile0 + row, n_snps - 1);
            const long long off = (long long)chunk * kRowBytes + part * 16;
            if (off + 16 <= pitch)
                cp_async16(gs + row * kRowPitch + part * 16, geno + (long long)snp * pitch + off);
            else        // past the row (a pitch that is not a multiple of 64 B): masked samples anyway
                *(uint4 *)(gs + row * kRowPitch + part * 16) = make_uint4(~0u, ~0u, ~0u, ~0u);
        }
        const double *src = E + (long long)chunk * kChunk * NC;
        double *es = stage_E(s);
        for (int t = tid; t < kChunk * NC / 2; t += kThreads) cp_async16(es + 2 * t, src + 2 * t);
        if (tid < 4) cp_async16(stage_mask(s) + tid * 16, ormask + (long long)chunk * kRowBytes + tid * 16);
        cp_async_commit();
    };

    double acc[ROWS][NC];
    int nhet[ROWS], nhom[ROWS], nmis[ROWS];
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
        nhet[r] = nhom[r] = nmis[r] = 0;
#pragma unroll
        for (int c = 0; c < NC; ++c) acc[r][c] = 0.0;
    }

    if (Cfg::kStages == 2) issue(0, 0);
    for (int chunk = 0; chunk < n_chunks; ++chunk) {
        int s = (Cfg::kStages == 2) ? (chunk & 1) : 0;
        if (Cfg::kStages == 2) {
            if (chunk + 1 < n_chunks) { issue(chunk + 1, s ^ 1); cp_async_wait<1>(); }
            else cp_async_wait<0>();
        } else {
            issue(chunk, 0);
            cp_async_wait<0>();
        }
        __syncthreads();

        const unsigned char *gs = stage_geno(s);
#pragma unroll 1
        for (int h = 0; h < 2 * NW64; ++h) {
            const int sub = warp * NW64 + (h >> 1);           // 32-sample group within the chunk
            const int half = h & 1;                           // 16-sample half of the group
            const double *es = stage_E(s) + (sub * 32 + half * 16) * NC;
            uint32_t w[ROWS];
            {
                const uint32_t mk = *(const uint32_t *)(stage_mask(s) + sub * 8 + half * 4);
#pragma unroll
                for (int r = 0; r < ROWS; ++r) {
                    uint32_t v = *(const uint32_t *)(gs + (lane + 32 * r) * kRowPitch + sub * 8 + half * 4) | mk;
                    w[r] = v;
                    // exact class counts from the bit planes (even bit = low bit of the code)
                    uint32_t lo = v & 0x55555555u, hi = (v >> 1) & 0x55555555u;
                    nmis[r] += __popc(lo & ~hi);
                    nhet[r] += __popc(hi & ~lo);
                    nhom[r] += __popc(~(lo | hi) & 0x55555555u);
                }
            }
#pragma unroll
            for (int sidx = 0; sidx < 16; ++sidx) {
                double e[NC];
                const double2 *ep = (const double2 *)(es + sidx * NC);
#pragma unroll
                for (int c = 0; c < NC / 2; ++c) {
                    double2 v = ep[c];
                    e[2 * c] = v.x; e[2 * c + 1] = v.y;
                }
#pragma unroll
                for (int r = 0; r < ROWS; ++r) {
                    // byte offset of the table entry: code * 16
                    uint32_t off = (sidx >= 2) ? ((w[r] >> (2 * sidx - 4)) & 0x30u)
                                               : ((w[r] << (4 - 2 * sidx)) & 0x30u);
                    if (NCB > 0) {
                        double2 gg = *(const double2 *)((const unsigned char *)tab + off);
#pragma unroll
                        for (int c = 0; c < NCA; ++c) acc[r][c] = fma(gg.x, e[c], acc[r][c]);
#pragma unroll
                        for (int c = 0; c < NCB; ++c)
                            acc[r][NCA + c] = fma(gg.y, e[NCA + c], acc[r][NCA + c]);
                    } else {
                        double g = *(const double *)((const unsigned char *)tab + off);
#pragma unroll
                        for (int c = 0; c < NCA; ++c) acc[r][c] = fma(g, e[c], acc[r][c]);
                    }
                }
            }
        }
        __syncthreads();
    }

    // cross-warp reduction through shared memory, one row set at a time
    double *red = (double *)smem;                        // [WARPS][32][NC]
    int *redc = (int *)(smem + WARPS * 32 * NC * 8);     // [WARPS][32][3]
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
        double *mine = red + (warp * 32 + lane) * NC;
#pragma unroll
        for (int c = 0; c < NC; ++c) mine[c] = acc[r][c];
        int *minec = redc + (warp * 32 + lane) * 3;
        minec[0] = nhet[r]; minec[1] = nhom[r]; minec[2] = nmis[r];
        __syncthreads();
        for (int t = tid; t < 32 * NC; t += kThreads) {
            double sum = 0.0;
#pragma unroll
            for (int wv = 0; wv < WARPS; ++wv) sum += red[wv * 32 * NC + t];
            int snp = tile0 + 32 * r + t / NC;
            if (snp < n_snps) S[(long long)snp * NC + t % NC] = sum;
        }
        for (int t = tid; t < 32; t += kThreads) {
            int c3[3] = {0, 0, 0};
#pragma unroll
            for (int wv = 0; wv < WARPS; ++wv)
#pragma unroll
                for (int q = 0; q < 3; ++q) c3[q] += redc[(wv * 32 + t) * 3 + q];
            int snp = tile0 + 32 * r + t;
            if (snp < n_snps) {
                counts[(long long)snp * 4 + 0] = c3[0];
                counts[(long long)snp * 4 + 1] = c3[1];
                counts[(long long)snp * 4 + 2] = c3[2];
                gsum[(long long)snp * 2 + 0] = (double)(c3[0] + 2 * c3[1]);
                gsum[(long long)snp * 2 + 1] = (double)(c3[0] + 4 * c3[1]);
            }
        }
        __syncthreads();
    }
}

// Real-valued dosage rows (NaN = missing), already aligned to the design
// rows.  Not the bandwidth path: one warp per SNP, lanes stride the samples,
// runtime column count, warp-shuffle reduction.  counts = {-1, -1, n_missing}.
__global__ void __launch_bounds__(256)
contract_dosage_kernel(const double *__restrict__ dosage, long long ld, int n_snps, int n,
                       const double *__restrict__ E, int NCA, int NCB, int NC1, int NC2,
                       double *__restrict__ S, int *__restrict__ counts,
                       double *__restrict__ gsum) {
    const int lane = threadIdx.x & 31;
    const int snp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (snp >= n_snps) return;
    const int NC = NCA + NCB;
    const double *row = dosage + (long long)snp * ld;
    int nm = 0;
    double sg = 0.0, sg2 = 0.0;
    for (int i = lane; i < n; i += 32) {
        double g = row[i];
        if (g != g) { nm += 1; continue; }
        sg += g; sg2 = fma(g, g, sg2);
    }
    nm = gt_warp_sum_int(nm);
    sg = gt_warp_sum(sg);
    sg2 = gt_warp_sum(sg2);
    if (lane == 0) {
        counts[(long long)snp * 4 + 0] = -1;
        counts[(long long)snp * 4 + 1] = -1;
        counts[(long long)snp * 4 + 2] = nm;
        gsum[(long long)snp * 2 + 0] = sg;
        gsum[(long long)snp * 2 + 1] = sg2;
    }
    for (int c = 0; c < NC; ++c) {
        bool live = (c < NCA) ? (c < NC1) : (c - NCA < NC2);
        double sum = 0.0;
        if (live) {
            for (int i = lane; i < n; i += 32) {
                double g = row[i];
                if (g != g) continue;
                double m = (c < NCA) ? g : g * g;
                sum = fma(m, E[(long long)i * NC + c], sum);
            }
            sum = gt_warp_sum(sum);
        }
        if (lane == 0) S[(long long)snp * NC + c] = sum;
    }
}

}  // namespace gtk
