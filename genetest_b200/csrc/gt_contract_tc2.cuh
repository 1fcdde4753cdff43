// K1 + K2 on the 5th-generation tensor cores (tcgen05, kind::i8): the exact
// integer formulation of the genotype contraction
//     S[snp][c] = sum_i g_i E_ic          (reference: analysis.py:108-172,
//                                          linear.py:93-104 via the normal equations)
// with the packed genotypes staged by TMA and every per-sample side job kept
// branch-free in the expansion threads.
//
// Dosages are the integers {0,1,2}; every E column is split on the host into
// seven balanced base-256 digits (int8, 54-bit fixed point relative to the
// column's maximum), so the sum becomes seven exact int32 dot products that
// the epilogue recombines in FP64.  When the design spans a constant, one
// column needs no digits at all: it follows from sum(g) -- the D column of a
// plane of ones -- and the other columns (see Tc2Params::drv_col).
//
//   packed rows  HBM --TMA 2-D tiles (rows x 128 B = 512 samples, 128B swizzle)--> smem ring
//   A (ROWS x K, uint8 dosages) lives in TENSOR MEMORY: one thread per SNP row
//     reads its 16 B of a 64-sample slot from the swizzled tile (conflict
//     free), expands the 2-bit codes with byte permutes and tcgen05.st's them;
//   B (N x K int8 digit planes, K-major, 128B swizzle) is a pre-swizzled image,
//     one cp.async.bulk per 128 samples, shared by ALL 128-row tiles of the
//     CTA: the kernel is bound by L2 -> SM bandwidth (measured ~5.9 TB/s
//     chip-wide), and with four tiles per CTA the digit planes add 0.625 B per
//     genotype byte instead of 2.5 B with one tile;
//   D (128 x N int32 per tile) accumulates in TMEM; one thread issues
//     tcgen05.mma.cta_group::1.kind::i8 (M=128, N, K=32) for every tile and
//     tcgen05.commit releases the A slot and the B stage.
// Side jobs of the expansion threads, per 32 samples: one popcount (homozygote
// count -> exact MAF / sum g^2) and one predicated 8-byte store of
// (pair index, missing flags) when any genotype of the pair is missing -- the
// correction kernel (gt_linear.cuh) turns those records into sample indices.
// Nothing in the sample loop diverges.
//
// Warps: 0 .. 4*TILES-1 expansion + epilogue, then one lane each of three
// service warps: B producer, genotype (TMA) producer, MMA issuer.
// One CTA per SM (TILES = 4: 512 SNP rows, 608 threads).
#pragma once

#include <cuda.h>
#include <cstdio>

#include "gt_common.cuh"
#include "gt_contract_tc.cuh"   // mbarrier / prmt / descriptor helpers

namespace gtk {

constexpr int kT2Digits = 7;          // base-256 digits per E column
constexpr int kT2MaxCols = 36;        // digit columns a design can have (N <= 256)

struct Tc2Params {
    long long pitch;          // bytes between packed rows
    int n_snps;
    int n_gstages;            // 512-sample genotype stages per row
    int mask_zero_stages;     // leading stages whose ormask bytes are all zero
    const uint8_t *ormask;    // [n_gstages * 128] 11 = sample not in the design
    const uint8_t *Bimg;      // [4 * n_gstages][N * 128] pre-swizzled digit planes
    int ncolL;                // digit columns (7 planes each), then the ones plane
    int out_col[kT2MaxCols];  // S column of digit column cl
    int out_col2[kT2MaxCols]; // second S column holding the same values (interaction designs), or -1
    int ones_out;             // S column that is the plain dosage sum (E column identically one), or -1
    // interaction designs (HOM kernels): the sums of g^2 m over the columns m = 1, w_a, w_a w_b are
    // sum(g m) + 2 sum(h m) with h the homozygote indicator (rows 64..127 of every MMA tile)
    int n_s2;                 // columns of the g^2 block (0: none)
    int s2_cl[12];            // digit column of g^2 column idx (-1: the ones plane)
    int s2_out;               // S column of g^2 column 0 (the others follow)
    int pad0[4], pad1[4];     // S column ranges [pad0[i], pad1[i]) to clear
    double scale[kT2MaxCols]; // column maximum * 2^-54
    // When the constant vector lies in the span of the digit columns plus one more
    // column j (1 = sum_c a_c E_c over the design rows), column j needs no digits:
    // S_j = (sum(g) - sum_{c != j} a_c S_c) / a_j, and sum(g) is the ones plane.
    int drv_col;              // S column j derived that way, or -1
    double drv_a[kT2MaxCols]; // a_c of digit column cl
    double drv_inv;           // 1 / a_j
    int ncol, NCP;            // live S columns, S row stride
    double *S;
    int *counts;              // [n_snps][4] n_het, n_hom, n_mis (-1: K3a fills it in), 0
    double *gsum;             // [n_snps][2] sum g, sum g^2
    uint2 *mrec;              // [n_snps][rcap] (pair index, flags) of the 32-sample pairs with missing genotypes
    int rcap;                 // records per row = 16 * n_gstages (every pair: no overflow)
    int *mcount;              // [n_snps] records of the row
    int *error_flag;
};

__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, int x, int y, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n"
        ::"r"(smem_u32(dst)), "l"(map), "r"(x), "r"(y), "r"(smem_u32(bar)) : "memory");
}
// One try_wait inline (it suspends the thread for a while by itself), the rest
// of a bounded wait out of line: the sample loop stays small.
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    return done != 0;
}
// non-blocking poll
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    return done != 0;
}
__device__ __noinline__ bool mbar_wait_slow(uint32_t bar, uint32_t parity, volatile int *fail) {
#pragma unroll 1
    for (int spin = 0; spin < (1 << 22); ++spin) {
        if (mbar_try(bar, parity)) return true;
        if ((spin & 255) == 255 && *fail) return false;
    }
    return false;
}
__device__ __forceinline__ bool mbar_wait2(uint64_t *bar, uint32_t parity, volatile int *fail) {
    const uint32_t a = smem_u32(bar);
    if (mbar_try(a, parity)) return true;
    return mbar_wait_slow(a, parity, fail);
}
// one lane of a converged warp (the compiler keeps warp-uniform operands in
// uniform registers when the single-thread instructions sit under this predicate
// instead of inside an `if (lane == 0)` region)
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(pred));
    return pred != 0;
}
// x >> 16 as the high half of x * 2^16: an IMAD.HI on the FMA pipe instead of a shift on the
// ALU pipe, which the expansion loop saturates (LOP3 / SHF / PRMT all issue there).  (Doing the
// same to the >> 1 and >> 2 shifts of the flag / selector words measured 3 % slower, and two
// instead of three A slots 18 % slower.)
__device__ __forceinline__ uint32_t hi16(uint32_t x) {
    uint32_t r;
    asm("mul.hi.u32 %0, %1, 65536;\n" : "=r"(r) : "r"(x));
    return r;
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
    uint4 r;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];\n"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(addr));
    return r;
}

// HOM = 1: every 128-row MMA tile holds 64 markers twice -- rows 0..63 their dosages,
// rows 64..127 their homozygote indicators (same packed bytes, another lookup table)
template <int TILES, int HOM = 0> struct Tc2Cfg {
    static constexpr int kExpWarps = 4 * TILES;
    static constexpr int kThreads = (kExpWarps + 3) * 32;
    static constexpr int kRows = (HOM ? 64 : 128) * TILES;  // markers (packed rows) of a CTA
    static constexpr int kGStage = kRows * 128;          // bytes of one genotype stage (512 samples)
    static constexpr int kGStages = kRows >= 512 ? 2 : (kRows == 256 ? 3 : 4);
    static constexpr int kBoxRows = kRows < 256 ? kRows : 256;   // rows of one TMA box
    __host__ __device__ static constexpr int a_slots(int N) {   // 64-sample A slots in TMEM
        return ((512 - TILES * N) / (TILES * 16)) < 8 ? ((512 - TILES * N) / (TILES * 16)) : 8;
    }
    __host__ __device__ static constexpr int b_stages(int N) {
        // what is left of ~210 KB after the genotype ring, at most 6 stages of 128 samples
        return ((210 * 1024 - kGStages * kGStage - kRecRing) / (N * 128)) < 6
                   ? ((210 * 1024 - kGStages * kGStage - kRecRing) / (N * 128)) : 6;
    }
    // staging ring of the missing-pair records: 8 records of 8 B per expansion thread
    static constexpr int kRecRing = kRows * 64;
    __host__ __device__ static constexpr int smem_bytes(int N) {
        return 1024 + kGStages * kGStage + b_stages(N) * N * 128 + 512 + kRecRing + (HOM ? kRows * 13 * 8 : 0);
    }
};

template <int NT /* N / 16 */, int TILES,
          int DBG = 0 /* tuning: 1 no records, 2 no MMA, 4 no TMEM store, 8 genotypes from L2, 16 no counts,
                         32 no B copies, 128 print the roles' wait cycles */,
          int HOM = 0>
__global__ void __launch_bounds__(Tc2Cfg<TILES, HOM>::kThreads, 1)
contract_tc2_kernel(const __grid_constant__ CUtensorMap gmap, const Tc2Params P) {
    using Cfg = Tc2Cfg<TILES, HOM>;
    constexpr int N = 16 * NT;
    constexpr int BSTAGE = N * 128;
    constexpr int BST = Cfg::b_stages(N);
    constexpr int G = Cfg::kGStages;
    constexpr int S = Cfg::a_slots(N);
    constexpr int EW = Cfg::kExpWarps;
    constexpr int TMEM_NEED = TILES * N + S * TILES * 16;
    constexpr int TMEM_COLS = TMEM_NEED <= 128 ? 128 : (TMEM_NEED <= 256 ? 256 : 512);
    static_assert(S >= 2 && TMEM_NEED <= 512, "accumulators + A ring exceed tensor memory");
    static_assert(BST >= 2, "no room for the B ring");

    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned char *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    unsigned char *sG = smem;                                   // G * kGStage, 1024 B aligned stages
    unsigned char *sB = smem + G * Cfg::kGStage;                // BST * BSTAGE
    uint64_t *bars = (uint64_t *)(sB + BST * BSTAGE);
    uint64_t *g_full = bars, *g_empty = bars + G, *b_full = bars + 2 * G, *b_empty = bars + 2 * G + BST,
             *a_full = bars + 2 * G + 2 * BST, *a_empty = a_full + S, *d_full = a_empty + S;
    uint32_t *tmem_slot = (uint32_t *)(d_full + 1);
    int *fail = (int *)(tmem_slot + 1);
    uint32_t *lut_s = tmem_slot + 2;
    unsigned char *sRec = (unsigned char *)(bars + 64);         // [8][kRows] records (uint2)
    double *sHom = (double *)(sRec + Cfg::kRecRing);            // HOM: [13][kRows] sums of h m

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // probe (DBG & 128): cycles the roles of CTA 0 spend in their waits
    long long prof[4] = {0, 0, 0, 0};
    const long long prof_t0 = (DBG & 128) ? clock64() : 0;
#define GT_PROF(i, stmt) do { if (DBG & 128) { long long _t = clock64(); stmt; prof[i] += clock64() - _t; } else { stmt; } } while (0)
    if (tid == 0) {
        for (int s = 0; s < G; ++s) { mbar_init(&g_full[s], 1); mbar_init(&g_empty[s], EW); }
        for (int s = 0; s < BST; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
        for (int s = 0; s < S; ++s) { mbar_init(&a_full[s], EW); mbar_init(&a_empty[s], 1); }
        mbar_init(d_full, 1);
        *fail = 0;
        lut_s[0] = 0x00010002u;
        lut_s[1] = 0x00000001u;
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n"
                     ::"r"(smem_u32(tmem_slot)), "n"(TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n");
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_a = tmem_base + TILES * N;          // [slot][tile] x 16 columns
    const int n_gst = P.n_gstages;
    const int n_slots = 8 * n_gst;                          // 64-sample slots
    const int row0 = blockIdx.x * Cfg::kRows;

    if (warp < EW) {
        // ===== expansion: one thread per SNP row, 64 samples per slot =====
        const int tile = warp >> 2;
        const int lgrp = warp & 3;                       // TMEM lanes 32 * lgrp .. of this warp
        // HOM: warps 0, 1 of a tile expand the dosages of its 64 markers, warps 2, 3 their
        // homozygote indicators (the class counts and the records are the dosage threads' job)
        const bool is_h = HOM && lgrp >= 2;
        const int row = HOM ? tile * 64 + (lgrp & 1) * 32 + lane
                            : tile * 128 + lgrp * 32 + lane;   // marker (packed row) of the CTA
        const int snp = row0 + row;
        const bool owner = snp < P.n_snps;
        const uint32_t lane_base = ((uint32_t)(lgrp * 32)) << 16;
        // 16-byte chunk c of a row sits at chunk c ^ (row & 7) of its 128-byte line
        const uint32_t g_row = smem_u32(sG) + row * 128;
        const uint32_t sw = (uint32_t)(row & 7);
        // rows past the batch read as zeros (TMA fill): they never produce a record,
        // and the record buffer is sized for whole CTAs anyway
        // Records are staged in shared memory (slot k of this thread at [k][row]: every
        // access is conflict free) and leave in whole 32-byte sectors, four at a time:
        // 8-byte stores scattered over 512 rows cost the L2 a read-modify-write each.
        char *const mbase = (char *)(P.mrec + (long long)snp * P.rcap);
        const uint32_t rec_s = smem_u32(sRec) + row * 8;
        uint32_t wcnt = 0, flushed = 0;                  // records staged / written to global
        int nhom = 0;
        // code -> dosage byte: 00 -> 2, 01 -> 0, 10 -> 1, 11 -> 0.  Opaque to the
        // compiler (read back from shared memory) so that it stays in one vector
        // register: a known constant lands in a uniform register and is copied per PRMT.
        const uint32_t lut = ((volatile uint32_t *)lut_s)[is_h ? 1 : 0];

        auto publish = [&](int it) {     // slot `it` is in tensor memory: tell the MMA thread
            GT_PROF(2, asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory"));
            asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
            __syncwarp();
            // lane 0 arrives -- predicated, not branched
            asm volatile(
                "{\n\t.reg .pred p;\n\t"
                "setp.eq.u32 p, %1, 0;\n\t"
                "@p mbarrier.arrive.shared::cta.b64 _, [%0];\n\t}\n"
                ::"r"(smem_u32(&a_full[it % S])), "r"(lane) : "memory");
        };
        // one 64-sample slot: this thread's 16 bytes -> 16 registers of dosages
        auto do_slot = [&](const uint4 pw, const uint4 m4, const int it) {
            const int s = it % S;
            // poll the slot's release early: the answer is needed only before the store
            const uint32_t a_emp = smem_u32(&a_empty[s]);
            const uint32_t a_par = ((it / S) & 1) ^ 1;
            const bool freed = it < S || mbar_try(a_emp, a_par);
            const uint32_t words[4] = {pw.x, pw.y, pw.z, pw.w};
            const uint32_t masks[4] = {m4.x, m4.y, m4.z, m4.w};
            uint32_t a[16];
            uint32_t homf[4], misf[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint32_t w = words[j];
                const uint32_t wm = w | masks[j];
                // flags of an even word sit on the even bits, those of an odd word on the odd
                // bits, so that two words share one flag word without any realignment
                if (!(j & 1)) {
                    const uint32_t s1 = wm >> 1;
                    homf[j] = ~(wm | s1) & 0x55555555u;          // code 00
                    misf[j] = wm & ~s1 & 0x55555555u;            // code 01
                } else {
                    const uint32_t s1 = wm << 1;
                    homf[j] = ~(wm | s1) & 0xaaaaaaaau;
                    misf[j] = s1 & ~wm & 0xaaaaaaaau;
                }
                const uint32_t even = w & 0x33333333u, odd = (w >> 2) & 0x33333333u;
                a[j * 4 + 0] = prmt(lut, 0u, even);
                a[j * 4 + 1] = prmt(lut, 0u, hi16(even));
                a[j * 4 + 2] = prmt(lut, 0u, odd);
                a[j * 4 + 3] = prmt(lut, 0u, hi16(odd));
            }
            if (!is_h) {
#pragma unroll
            for (int pr = 0; pr < 2; ++pr) {
                if (!(DBG & 16) && !HOM) nhom += __popc(homf[2 * pr] | homf[2 * pr + 1]);
                const uint32_t mis2 = misf[2 * pr] | misf[2 * pr + 1];
                if (!(DBG & 1)) {
                    // predicated, not branched: (pair index, flags) when any flag is set
                    asm volatile(
                        "{\n\t.reg .pred p;\n\t"
                        "setp.ne.u32 p, %3, 0;\n\t"
                        "@p st.shared.v2.u32 [%1], {%2, %3};\n\t"
                        "@p add.u32 %0, %0, 1;\n\t}\n"
                        : "+r"(wcnt) : "r"(rec_s + ((wcnt & 7u) * (uint32_t)(Cfg::kRows * 8))),
                          "r"((uint32_t)(it * 2 + pr)), "r"(mis2)
                        : "memory");
                }
            }
            }
            // publish the PREVIOUS slot only now: its TMEM store drained while
            // this slot's bytes were being expanded
            if (it > 0) publish(it - 1);
            GT_PROF(1, if (!freed) { if (!mbar_wait_slow(a_emp, a_par, fail)) *fail = 1; });
            asm volatile("tcgen05.fence::after_thread_sync;\n");
            const uint32_t taddr = tmem_a + lane_base + (s * TILES + tile) * 16;
            if (!(DBG & 4))
            asm volatile(
                "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
                "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};\n"
                ::"r"(taddr),
                  "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]),
                  "r"(a[8]), "r"(a[9]), "r"(a[10]), "r"(a[11]), "r"(a[12]), "r"(a[13]), "r"(a[14]), "r"(a[15])
                : "memory");
        };

        // every two slots (at most four new records): one whole sector leaves when four are staged
        auto flush4 = [&]() {
            if ((DBG & 1) || is_h || wcnt - flushed < 4u) return;
            const uint32_t src = rec_s + ((flushed & 4u) * (uint32_t)(Cfg::kRows * 8));
            uint2 r0, r1, r2, r3;
            asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];\n" : "=r"(r0.x), "=r"(r0.y) : "r"(src));
            asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];\n" : "=r"(r1.x), "=r"(r1.y) : "r"(src + Cfg::kRows * 8));
            asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];\n" : "=r"(r2.x), "=r"(r2.y) : "r"(src + Cfg::kRows * 16));
            asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];\n" : "=r"(r3.x), "=r"(r3.y) : "r"(src + Cfg::kRows * 24));
            uint4 *dst = (uint4 *)(mbase + (size_t)flushed * 8);
            dst[0] = make_uint4(r0.x, r0.y, r1.x, r1.y);
            dst[1] = make_uint4(r2.x, r2.y, r3.x, r3.y);
            flushed += 4;
        };

        for (int gs = 0; gs < n_gst; ++gs) {
            const int st = gs % G;
            GT_PROF(0, if (!mbar_wait2(&g_full[st], (gs / G) & 1, fail)) *fail = 1);
            const uint32_t g_st = g_row + st * Cfg::kGStage;
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {         // four slots' bytes in registers at a time
                uint4 pw[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) pw[q] = lds128(g_st + ((((uint32_t)(4 * hf + q)) ^ sw) << 4));
                if (hf == 1) {
                    // the stage's bytes are in registers: hand it back
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&g_empty[st]);
                }
                if (gs < P.mask_zero_stages) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        do_slot(pw[q], make_uint4(0u, 0u, 0u, 0u), gs * 8 + hf * 4 + q);
                        if (q & 1) flush4();
                    }
                } else {
                    const uint4 *mrow = (const uint4 *)(P.ormask + (long long)gs * 128);
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        do_slot(pw[q], __ldg(mrow + 4 * hf + q), gs * 8 + hf * 4 + q);
                        if (q & 1) flush4();
                    }
                }
            }
        }
        if (n_slots > 0) publish(n_slots - 1);
        if ((DBG & 128) && blockIdx.x == 0 && (tid == 0 || tid == 32 * (EW - 1)))
            printf("expansion warp %d: total %lld  g_full %lld  a_empty %lld  wait_st %lld\n", warp,
                   clock64() - prof_t0, prof[0], prof[1], prof[2]);
        const int nrec = (int)wcnt;
        if (!(DBG & 1) && !is_h) {
            // at most three staged records are left
            for (; flushed < wcnt; ++flushed) {
                uint2 r;
                asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];\n" : "=r"(r.x), "=r"(r.y)
                             : "r"(rec_s + ((flushed & 7u) * (uint32_t)(Cfg::kRows * 8))));
                *(uint2 *)(mbase + (size_t)flushed * 8) = r;
            }
        }
        if (owner && !is_h) {
            P.mcount[snp] = nrec;
            if (!HOM)
                for (int col = P.ncol; col < P.NCP; ++col) P.S[(long long)snp * P.NCP + col] = 0.0;
            else
                for (int r = 0; r < 4; ++r)
                    for (int col = P.pad0[r]; col < P.pad1[r]; ++col) P.S[(long long)snp * P.NCP + col] = 0.0;
        }
        // ===== epilogue: D (int32, TMEM) -> FP64 sums =====
        if (!mbar_wait2(d_full, 0, fail)) *fail = 1;
        asm volatile("tcgen05.fence::after_thread_sync;\n");
        constexpr int NCL = (N - 1) / kT2Digits;             // digit columns that fit N
        double acc[NCL];
#pragma unroll
        for (int i = 0; i < NCL; ++i) acc[i] = 0.0;
        int sum_g = 0;
        const uint32_t tmem_d = tmem_base + tile * N + lane_base;
        const int ones_n = P.ncolL * kT2Digits;
#pragma unroll
        for (int ch = 0; ch < NT; ++ch) {
            uint32_t d[16];
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
                : "=r"(d[0]), "=r"(d[1]), "=r"(d[2]), "=r"(d[3]), "=r"(d[4]), "=r"(d[5]), "=r"(d[6]), "=r"(d[7]),
                  "=r"(d[8]), "=r"(d[9]), "=r"(d[10]), "=r"(d[11]), "=r"(d[12]), "=r"(d[13]), "=r"(d[14]), "=r"(d[15])
                : "r"(tmem_d + ch * 16));
            asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
            for (int u = 0; u < 16; ++u) {
                const int n = ch * 16 + u;                   // plane n = 7 * column + digit
                if (n == ones_n) sum_g = (int)d[u];
                if (n < NCL * kT2Digits) {
                    const int cl = n / kT2Digits, dg = n % kT2Digits;
                    // 256^dg is a power of two: every product is exact, the seven-term sum rounds ~1e-16
                    const double wgt = (double)(1ull << (8 * dg));
                    acc[cl] = fma((double)(int)d[u], wgt, acc[cl]);
                }
            }
        }
        if constexpr (HOM != 0) {
            // value of digit column `c` (or of the ones plane for c < 0) in this thread's D row
            auto pick = [&](int c) {
                double r = (double)sum_g;
#pragma unroll
                for (int cl = 0; cl < NCL; ++cl) r = (c == cl) ? acc[cl] * P.scale[cl] : r;
                return r;
            };
            // the indicator rows hand sum(h), sum(h m) to the dosage thread of the same marker
            if (is_h) {
                sHom[row] = (double)sum_g;
                for (int idx = 0; idx < P.n_s2; ++idx) sHom[(1 + idx) * Cfg::kRows + row] = pick(P.s2_cl[idx]);
            }
            asm volatile("bar.sync 1, %0;\n" ::"n"(EW * 32) : "memory");
            if (!is_h && owner) {
                double *Srow = P.S + (long long)snp * P.NCP;
                double part = 0.0;
#pragma unroll
                for (int cl = 0; cl < NCL; ++cl)
                    if (cl < P.ncolL) {
                        const double sv = acc[cl] * P.scale[cl];
                        Srow[P.out_col[cl]] = sv;
                        if (P.out_col2[cl] >= 0) Srow[P.out_col2[cl]] = sv;
                        part = fma(P.drv_a[cl], sv, part);
                    }
                if (P.drv_col >= 0) Srow[P.drv_col] = ((double)sum_g - part) * P.drv_inv;
                if (P.ones_out >= 0) Srow[P.ones_out] = (double)sum_g;
                // sum g^2 m = sum g m + 2 sum h m   (g in {0, 1, 2}: g^2 = g + 2 [g == 2])
                for (int idx = 0; idx < P.n_s2; ++idx)
                    Srow[P.s2_out + idx] = fma(2.0, sHom[(1 + idx) * Cfg::kRows + row], pick(P.s2_cl[idx]));
                const int nh = (int)sHom[row];
                P.counts[(long long)snp * 4 + 0] = sum_g - 2 * nh;
                P.counts[(long long)snp * 4 + 1] = nh;
                P.counts[(long long)snp * 4 + 2] = nrec > 0 ? -1 : 0;
                P.counts[(long long)snp * 4 + 3] = 0;
                P.gsum[(long long)snp * 2 + 0] = (double)sum_g;
                P.gsum[(long long)snp * 2 + 1] = (double)(sum_g + 2 * nh);
            }
        } else
        if (owner) {
            double part = 0.0;
#pragma unroll
            for (int cl = 0; cl < NCL; ++cl)
                if (cl < P.ncolL) {
                    const double sv = acc[cl] * P.scale[cl];
                    P.S[(long long)snp * P.NCP + P.out_col[cl]] = sv;
                    part = fma(P.drv_a[cl], sv, part);
                }
            if (P.drv_col >= 0)
                P.S[(long long)snp * P.NCP + P.drv_col] = ((double)sum_g - part) * P.drv_inv;
            // exact integers: sum of the dosages over the design rows, class counts
            P.counts[(long long)snp * 4 + 0] = sum_g - 2 * nhom;
            P.counts[(long long)snp * 4 + 1] = nhom;
            P.counts[(long long)snp * 4 + 2] = nrec > 0 ? -1 : 0;    // K3a counts the flags
            P.counts[(long long)snp * 4 + 3] = 0;
            P.gsum[(long long)snp * 2 + 0] = (double)sum_g;
            P.gsum[(long long)snp * 2 + 1] = (double)(sum_g + 2 * nhom);
        }
        asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    } else if (warp == EW) {
        // ===== B producer: one bulk copy per 128 samples (whole warp, one elected lane issues) =====
        const int n_bst = 4 * n_gst;
        for (int ib = 0; ib < n_bst; ++ib) {
            const int s = ib % BST;
            GT_PROF(0, if (ib >= BST && !mbar_wait2(&b_empty[s], ((ib / BST) & 1) ^ 1, fail)) *fail = 1);
            if (elect_one()) {
                if (DBG & 32) {
                    mbar_arrive(&b_full[s]);                              // probe: no B traffic
                } else {
                    mbar_expect_tx(&b_full[s], BSTAGE);
                    bulk_g2s(sB + s * BSTAGE, P.Bimg + (long long)ib * BSTAGE, BSTAGE, &b_full[s]);
                }
            }
            __syncwarp();
        }
        if ((DBG & 128) && blockIdx.x == 0 && lane == 0)
            printf("B producer: total %lld  b_empty %lld\n", clock64() - prof_t0, prof[0]);
    } else if (warp == EW + 1) {
        // ===== genotype producer: TMA tiles (<= 256 rows x 128 B) per 512-sample stage =====
        for (int gs = 0; gs < n_gst; ++gs) {
            const int s = gs % G;
            GT_PROF(0, if (gs >= G && !mbar_wait2(&g_empty[s], ((gs / G) & 1) ^ 1, fail)) *fail = 1);
            if (elect_one()) {
                mbar_expect_tx(&g_full[s], Cfg::kGStage);
#pragma unroll
                for (int bx = 0; bx < Cfg::kRows / Cfg::kBoxRows; ++bx)
                    tma_load_2d(sG + s * Cfg::kGStage + bx * Cfg::kBoxRows * 128, &gmap,
                                (DBG & 8) ? 0 : gs * 128, row0 + bx * Cfg::kBoxRows, &g_full[s]);
            }
            __syncwarp();
        }
        if ((DBG & 128) && blockIdx.x == 0 && lane == 0)
            printf("genotype producer: total %lld  g_empty %lld\n", clock64() - prof_t0, prof[0]);
    } else {
        // ===== MMA issuer: the whole warp runs the loop, one elected lane issues =====
        // instruction descriptor: D=S32, A=U8 (dosages), B=S8 (digits), K-major both
        const uint32_t idesc = (2u << 4) | (0u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) |
                               ((uint32_t)(128 >> 4) << 24);
        // only the 14-bit start address of the B descriptor changes from MMA to MMA
        const uint64_t bdesc0 = tc_smem_desc(smem_u32(sB));
        auto mma = [&](uint32_t dcol, uint32_t acol, uint64_t bdesc, uint32_t accum) {
            asm volatile(
                "{\n\t.reg .pred p;\n\t"
                "setp.ne.b32 p, %4, 0;\n\t"
                "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, {%5, %6, %7, %8}, p;\n\t}\n"
                ::"r"(dcol), "r"(acol), "l"(bdesc), "r"(idesc), "r"(accum),
                  "r"(0u), "r"(0u), "r"(0u), "r"(0u) : "memory");
        };
        for (int it = 0; it < n_slots; ++it) {
            const int s = it % S, ib = it >> 1, sb = ib % BST;
            if (!(it & 1)) GT_PROF(0, if (!mbar_wait2(&b_full[sb], (ib / BST) & 1, fail)) *fail = 1);
            GT_PROF(1, if (!mbar_wait2(&a_full[s], (it / S) & 1, fail)) *fail = 1);
            asm volatile("tcgen05.fence::after_thread_sync;\n");
            const long long tm0 = (DBG & 128) ? clock64() : 0;
            // this slot's 64 samples are bytes [64 (it & 1), +64) of the B stage's 128-byte rows
            const uint64_t bd = bdesc0 + (uint64_t)((sb * BSTAGE + (it & 1) * 64) >> 4);
            const uint32_t a0 = tmem_a + (s * TILES) * 16;
            if (elect_one()) {
                if (!(DBG & 2)) {
#pragma unroll
                    for (int t = 0; t < TILES; ++t) {
                        mma(tmem_base + t * N, a0 + t * 16, bd, it > 0 ? 1u : 0u);
                        mma(tmem_base + t * N, a0 + t * 16 + 8, bd + 2, 1u);
                    }
                }
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n"
                             ::"r"(smem_u32(&a_empty[s])) : "memory");
                if (it & 1)
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n"
                                 ::"r"(smem_u32(&b_empty[sb])) : "memory");
            }
            __syncwarp();
            if (DBG & 128) prof[2] += clock64() - tm0;
        }
        if (elect_one())
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n"
                         ::"r"(smem_u32(d_full)) : "memory");
        __syncwarp();
        if ((DBG & 128) && blockIdx.x == 0 && lane == 0)
            printf("MMA issuer: total %lld  b_full %lld  a_full %lld  mma issue + commits %lld\n",
                   clock64() - prof_t0, prof[0], prof[1], prof[2]);
    }
#undef GT_PROF
    __syncthreads();
    if (tid == 0 && *fail) atomicExch(P.error_flag, 1);
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "n"(TMEM_COLS));
    }
}

}  // namespace gtk
