// K1 + K2: stage PLINK 2-bit rows into shared memory with cp.async, decode
// them to dosages in registers and contract them against the design's E
// matrix in FP64.  Output per SNP: S[NCP] = sum_i g_i E1_i | sum_i g_i^2 E2_i
// and the exact integer genotype counts.
//
// Replaces (for a tile of 32*ROWS SNPs at once) the per-SNP staging of
// genetest/analysis.py:109-119 (NaN fill, label-aligned scatter, complete-case
// mask) and the X'X / X'y accumulation inside sm.OLS(...).fit()
// (statistics/models/linear.py:102-104).
#pragma once

#include "gt_common.cuh"

namespace gtk {

constexpr int kThreads = 256;          // 8 warps; warp w owns samples [32w, 32w+32) of a chunk
constexpr int kWarps = kThreads / 32;
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

// dosage (and its square) of the 2-bit code at bit position `pos` of word w,
// built as IEEE doubles with integer ops only: 00 -> 2, 01 (missing) -> 0,
// 10 -> 1, 11 -> 0.
__device__ __forceinline__ void decode_pair(uint32_t w, int pos, double &g, double &g2) {
    uint32_t lo = (w >> pos) & 1u;
    uint32_t hi = (w >> (pos + 1)) & 1u;
    // g  : hi=0 -> 2.0 (0x40000000), hi=1 -> 1.0 (0x3FF00000)
    // g^2: hi=0 -> 4.0 (0x40100000), hi=1 -> 1.0 (0x3FF00000)
    uint32_t h1 = 0x40000000u - (hi << 20);
    uint32_t h2 = 0x40100000u - (hi << 21);
    h1 = lo ? 0u : h1;
    h2 = lo ? 0u : h2;
    g = __hiloint2double((int)h1, 0);
    g2 = __hiloint2double((int)h2, 0);
}

template <int NCA, int NCB, int ROWS>
struct ContractCfg {
    static constexpr int NC = NCA + NCB;
    static constexpr int kTile = 32 * ROWS;                          // SNPs per CTA
    static constexpr int kGenoBytes = kTile * kRowPitch;
    static constexpr int kEBytes = kChunk * NC * 8;
    static constexpr int kStageBytes = kGenoBytes + kEBytes + kRowBytes;
    static constexpr int kStages = (2 * kStageBytes <= 200 * 1024) ? 2 : 1;
    static constexpr int kSmem = kStages * kStageBytes;
};

template <int NCA, int NCB, int ROWS>
__global__ void __launch_bounds__(kThreads, 1)
contract_packed_kernel(const uint8_t *__restrict__ geno, long long pitch, int n_snps,
                       const uint8_t *__restrict__ ormask, const double *__restrict__ E,
                       int n_chunks, double *__restrict__ S, int *__restrict__ counts) {
    using Cfg = ContractCfg<NCA, NCB, ROWS>;
    constexpr int NC = Cfg::NC;
    extern __shared__ __align__(16) unsigned char smem[];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tile0 = blockIdx.x * Cfg::kTile;

    auto stage_geno = [&](int s) { return smem + s * Cfg::kStageBytes; };
    auto stage_E = [&](int s) { return (double *)(smem + s * Cfg::kStageBytes + Cfg::kGenoBytes); };
    auto stage_mask = [&](int s) { return smem + s * Cfg::kStageBytes + Cfg::kGenoBytes + Cfg::kEBytes; };

    auto issue = [&](int chunk, int s) {
        // packed rows: kTile rows x 64 B  (4 x 16 B per row)
        unsigned char *gs = stage_geno(s);
        for (int t = tid; t < Cfg::kTile * 4; t += kThreads) {
            int row = t >> 2, part = t & 3;
            int snp = min(tile0 + row, n_snps - 1);
            cp_async16(gs + row * kRowPitch + part * 16,
                       geno + (long long)snp * pitch + (long long)chunk * kRowBytes + part * 16);
        }
        // E rows of this chunk: contiguous kChunk * NC doubles
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
        const double *es = stage_E(s) + warp * 32 * NC;
        const uint2 mk = *(const uint2 *)(stage_mask(s) + warp * 8);
        uint2 w[ROWS];
#pragma unroll
        for (int r = 0; r < ROWS; ++r) {
            w[r] = *(const uint2 *)(gs + (lane + 32 * r) * kRowPitch + warp * 8);
            w[r].x |= mk.x; w[r].y |= mk.y;
            // exact class counts from the bit planes (even bit = low bit of the code)
            uint32_t lox = w[r].x & 0x55555555u, hix = (w[r].x >> 1) & 0x55555555u;
            uint32_t loy = w[r].y & 0x55555555u, hiy = (w[r].y >> 1) & 0x55555555u;
            nmis[r] += __popc(lox & ~hix) + __popc(loy & ~hiy);
            nhet[r] += __popc(hix & ~lox) + __popc(hiy & ~loy);
            nhom[r] += __popc(~(lox | hix) & 0x55555555u) + __popc(~(loy | hiy) & 0x55555555u);
        }
#pragma unroll 4
        for (int sidx = 0; sidx < 32; ++sidx) {
            double e[NC];
            const double2 *ep = (const double2 *)(es + sidx * NC);
#pragma unroll
            for (int c = 0; c < NC / 2; ++c) {
                double2 v = ep[c];
                e[2 * c] = v.x; e[2 * c + 1] = v.y;
            }
#pragma unroll
            for (int r = 0; r < ROWS; ++r) {
                uint32_t word = (sidx < 16) ? w[r].x : w[r].y;
                double g, g2;
                decode_pair(word, 2 * (sidx & 15), g, g2);
#pragma unroll
                for (int c = 0; c < NCA; ++c) acc[r][c] = fma(g, e[c], acc[r][c]);
#pragma unroll
                for (int c = 0; c < NCB; ++c) acc[r][NCA + c] = fma(g2, e[NCA + c], acc[r][NCA + c]);
            }
        }
        __syncthreads();
    }

    // cross-warp reduction through shared memory, one row set at a time
    double *red = (double *)smem;                 // [kWarps][32][NC]   (<= one E stage)
    int *redc = (int *)(smem + kWarps * 32 * NC * 8);   // [kWarps][32][3]
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
            for (int wv = 0; wv < kWarps; ++wv) sum += red[wv * 32 * NC + t];
            int snp = tile0 + 32 * r + t / NC;
            if (snp < n_snps) S[(long long)snp * NC + t % NC] = sum;
        }
        for (int t = tid; t < 32 * 3; t += kThreads) {
            int sum = 0;
#pragma unroll
            for (int wv = 0; wv < kWarps; ++wv) sum += redc[wv * 32 * 3 + t];
            int snp = tile0 + 32 * r + t / 3;
            if (snp < n_snps) counts[(long long)snp * 4 + t % 3] = sum;
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
                       double *__restrict__ S, int *__restrict__ counts) {
    const int lane = threadIdx.x & 31;
    const int snp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (snp >= n_snps) return;
    const int NC = NCA + NCB;
    const double *row = dosage + (long long)snp * ld;
    int nm = 0;
    for (int i = lane; i < n; i += 32) nm += (row[i] != row[i]);
    nm = gt_warp_sum_int(nm);
    if (lane == 0) {
        counts[(long long)snp * 4 + 0] = -1;
        counts[(long long)snp * 4 + 1] = -1;
        counts[(long long)snp * 4 + 2] = nm;
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
