// K3a, streaming version: missing-sample corrections of the linear model for
// designs of up to 12 contraction columns.  Every marker has its own set of
// missing samples; the exact complete-case Gram matrix is the full-sample one
// minus the sum of v v' over those samples (analysis.py:119,172: statsmodels
// fits the complete cases).
//
// The warp-per-SNP kernel (gt_linear.cuh) gathers 96 B of E per missing sample
// from L2 -- at 1 % missing that is four times the bytes of the packed
// genotypes themselves and it runs at the L2 -> SM bandwidth limit.  Here the
// roles are turned round: ONE THREAD PER SNP keeps the whole upper triangle in
// registers, and the design streams through shared memory once per 224 SNPs:
//   * a producer warp bulk-copies blocks of 512 rows of E (FP64, 96 B per
//     sample, the same rows the old kernel gathered) into a four-buffer ring
//     (TMA, mbarrier complete_tx): three copies are in flight while one block is read;
//   * every thread walks its marker's records (32-sample pair index + missing
//     flags, in sample order, written by the contraction kernel; fetched ahead
//     of time into a private shared-memory ring with cp.async) and, for each
//     missing sample, reads the row from shared memory and accumulates v v' with
//     FMAs -- only the NV (NV + 1) / 2 products that are needed (the m8n8k4
//     tensor-core tiles compute 192 for 91);
//   * a thread may run ahead into the next two blocks (once they have landed)
//     while its warp finishes the current one, so the lanes' different counts
//     even out over the row instead of per block.
// The arithmetic is the old kernel's (FP64 products and sums of the same rows).
// Output: the upper triangle, SoA over the markers (coalesced stores), and the
// missing count of every marker.
#pragma once

#include "gt_common.cuh"
#include "gt_contract_tc.cuh"
#include "gt_contract_tc2.cuh"

#ifndef GT_CS_UNROLL4
#define GT_CS_UNROLL4 0
#endif

namespace gtk {

constexpr int kCsBlk = 512;             // samples per shared-memory block
constexpr int kCsBuf = 4;               // blocks in the ring
constexpr int kCsRow = 12;              // doubles per row (96 bytes)
constexpr int kCsRecRing = 16;          // records of a thread's prefetch ring (8 cp.async of 16 B)
constexpr int kCsSnps = 224;            // markers (threads) per CTA: 7 warps + the producer warp = 256 threads, 255 registers each
constexpr int kCsThreads = kCsSnps + 32;
constexpr int kCsFoldSnps = 256;        // FOLD: warp 0 issues the block copies itself, 8 warps of markers

struct CorrStreamParams {
    const double *Ec;         // [n_blocks * kCsBlk][12] design rows (streamed columns only), file order
    int n_blocks;
    int n_snps;
    const uint2 *mrec;        // [n_snps][rcap]
    int rcap;
    const int *mcount;        // [n_snps]
    int L;                    // base vector length k + 2 + I: [q | y | 1 | w]
    int one;                  // position of the constant in v (k + 1)
    int vcol[12];             // position in v of streamed column c
    int drv_col;              // position in v of the column derived from the constant, or -1
    double drv_a[12];         // 1 = sum_c drv_a[c] v[vcol[c]] + a_j v[drv_col]
    double drv_inv;           // 1 / a_j
    double *corrT;            // [L (L + 1) / 2][ldc] upper triangle (a <= b), entry a * L - a (a - 1) / 2 + (b - a)
    long long ldc;
    int *counts;              // [n_snps][4]: [2] = missing count
    int *error_flag;
};

__device__ __forceinline__ int cs_tri(int a, int b, int L) { return a * L - a * (a - 1) / 2 + (b - a); }

template <int NV, int FOLD = 0>
__global__ void __launch_bounds__(256, 1)
missing_corr_stream_kernel(const CorrStreamParams P) {
    constexpr int kSnps = FOLD ? kCsFoldSnps : kCsSnps;
    constexpr int NP = NV * (NV + 1) / 2;
    constexpr int BLKB = kCsBlk * kCsRow * 8;               // bytes of one block
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned char *smem = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);
    unsigned char *sE = smem;                                // kCsBuf blocks
    unsigned char *sR = smem + kCsBuf * BLKB;                     // [8 pair slots][kSnps] x 16 B: record rings
    uint64_t *full = (uint64_t *)(sR + (kCsRecRing / 2) * kSnps * 16), *empty = full + kCsBuf;
    int *fail = (int *)(empty + kCsBuf);
    double *sZ = (double *)(full + 16);                       // a row of zeros (128-byte aligned: chunk ^ 1 stays inside)
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int s = 0; s < kCsBuf; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], kSnps / 32); }
        *fail = 0;
        for (int c = 0; c < 16; ++c) sZ[c] = 0.0;
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();
    const int nblk = P.n_blocks;

    if (!FOLD && warp == kSnps / 32) {
        // ===== producer: one bulk copy per block =====
        for (int b = 0; b < nblk; ++b) {
            const int s = b % kCsBuf;
            if (b >= kCsBuf && !mbar_wait2(&empty[s], ((b / kCsBuf) & 1) ^ 1, fail)) *fail = 1;
            if (elect_one()) {
                mbar_expect_tx(&full[s], BLKB);
                bulk_g2s(sE + s * BLKB, (const unsigned char *)P.Ec + (long long)b * BLKB, BLKB, &full[s]);
            }
            __syncwarp();
        }
    } else {
        // ===== one thread per marker =====
        const int snp = blockIdx.x * kSnps + tid;
        const bool owner = snp < P.n_snps;
        const int nrec = owner ? P.mcount[snp] : 0;
        const uint2 *recs = P.mrec + (long long)(owner ? snp : 0) * P.rcap;
        double acc[NP], sum[NV];
#pragma unroll
        for (int i = 0; i < NP; ++i) acc[i] = 0.0;
#pragma unroll
        for (int i = 0; i < NV; ++i) sum[i] = 0.0;
        int n_miss = 0;
        // The records come through a private ring of 8 x 16 B (two records per cp.async),
        // six copies ahead of the one being read: slot q & 7 of record pair q.
        const uint32_t ring = smem_u32(sR) + tid * 16;
        auto fetch = [&](int q) {            // records 2 q, 2 q + 1 -> slot q & 7 (reads past the list are unused)
            asm volatile("cp.async.ca.shared.global [%0], [%1], 16;\n"
                         ::"r"(ring + (uint32_t)(q & 7) * (kSnps * 16)), "l"(recs + 2 * q) : "memory");
            asm volatile("cp.async.commit_group;\n" ::: "memory");
        };
        auto rec_at = [&](int r) {           // record r from the ring
            uint2 v;
            asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];\n" : "=r"(v.x), "=r"(v.y)
                         : "r"(ring + (uint32_t)((r >> 1) & 7) * (kSnps * 16) + (uint32_t)(r & 1) * 8));
            return v;
        };
        // invariant: with record pair q current, pairs up to q + 6 are requested, so that
        // "at most five pending" means pair q + 1 has landed as well
#pragma unroll
        for (int q = 0; q < 7; ++q) fetch(q);
        asm volatile("cp.async.wait_group 6;\n" ::: "memory");
        int ri = 0;
        uint32_t f = 0, pair = 0x7FFFFFFFu;
        if (nrec > 0) { const uint2 r0 = rec_at(0); pair = r0.x; f = r0.y; }
        const uint32_t sE_u = smem_u32(sE);
        const uint32_t zrow = smem_u32(sZ);
        const int ppb = kCsBlk / 32;                        // pairs per block
        // Software pipeline over the missing samples of this thread, written WITHOUT branches:
        // a lane that has no sample to take reads a row of zeros and adds zeros, so that the
        // index arithmetic of the next sample, the shared-memory loads and the FMAs of the
        // current one sit in one basic block and overlap (two register buffers, the loop body
        // is unrolled twice).
        double vA[12], vB[12];
#pragma unroll
        for (int c = 0; c < 12; ++c) vA[c] = vB[c] = 0.0;
        uint32_t pairA = 0x7FFFFFFFu, pairB = 0x7FFFFFFFu;   // pair of the row waiting in the buffer
        uint32_t limit = 0;
        // load_next issues the row of the thread's next sample (or of zeros) into a buffer and
        // reads the following record speculatively; advance(), called after the FMAs of the other
        // buffer, consumes that read -- so neither shared-memory latency is waited for.
        uint2 nr = make_uint2(0u, 0u);
        bool adv = false;
        auto load_next = [&](double (&v)[12], uint32_t &vp) {
            // the oldest copy in flight (the next record pair) has landed after this
            asm volatile("cp.async.wait_group 5;\n" ::: "memory");
            nr = rec_at(ri + 1);                             // speculative: used when this record runs out
            const bool have = pair < limit;
            const int bit = __ffs((int)f) - 1;
            const uint32_t fnew = f & (f - 1);
            const uint32_t idx = pair * 32u + 16u * (bit & 1) + (bit >> 1);
            // 16-byte chunk c of row i sits at chunk c ^ ((i >> 2) & 1): rows that are equal
            // mod 4 (same banks, 96-byte pitch) spread over both halves of every 32 bytes
            uint32_t addr0 = (sE_u + (idx & (kCsBuf * kCsBlk - 1)) * (kCsRow * 8)) ^ ((idx & 4u) << 2);
            addr0 = have ? addr0 : zrow;
            const uint32_t addr1 = addr0 ^ 16u;
#pragma unroll
            for (int c = 0; c < (NV + 1) / 2; ++c)
                asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];\n" : "=d"(v[2 * c]), "=d"(v[2 * c + 1])
                             : "r"(((c & 1) ? addr1 : addr0) + 32 * (c >> 1)));
            vp = have ? pair : 0x7FFFFFFFu;
            n_miss += have ? 1 : 0;
            adv = have && fnew == 0;
            f = have ? fnew : f;
        };
        auto advance = [&]() {           // branch-free: everything is selected / predicated on `adv`
            ri += adv ? 1 : 0;
            const bool more = ri < nrec;
            pair = adv ? (more ? nr.x : 0x7FFFFFFFu) : pair;
            f = adv ? (more ? nr.y : 0u) : f;
            const uint32_t q = (uint32_t)(ri >> 1) + 6u;
            asm volatile(
                "{\n\t.reg .pred p;\n\t"
                "setp.ne.u32 p, %0, 0;\n\t"
                "@p cp.async.ca.shared.global [%1], [%2], 16;\n\t"
                "@p cp.async.commit_group;\n\t}\n"
                ::"r"((uint32_t)(adv && !(ri & 1))), "r"(ring + (q & 7u) * (kSnps * 16)), "l"(recs + 2 * q)
                : "memory");
        };
        auto accumulate = [&](const double (&v)[12]) {
            int e = 0;
#pragma unroll
            for (int a = 0; a < NV; ++a) {
                sum[a] += v[a];
#pragma unroll
                for (int c = a; c < NV; ++c, ++e) acc[e] = fma(v[a], v[c], acc[e]);
            }
        };
        int landed = -1;                 // blocks 0 .. landed are known to be in the ring
        int issued = 0;                  // FOLD, warp 0: blocks whose copies are issued
        // FOLD: block `issued` reuses the slot of block issued - kCsBuf, which every warp must have
        // released; only the block this warp needs next is waited for, the lookahead is opportunistic
        auto try_issue = [&](int b) {
            while (issued < nblk && issued < b + kCsBuf) {
                if (issued >= kCsBuf) {
                    const uint32_t ea = smem_u32(&empty[issued % kCsBuf]);
                    const uint32_t ep = ((issued / kCsBuf) & 1) ^ 1;
                    if (!mbar_test(ea, ep)) {
                        if (issued > b) break;
                        int spin = 0;
                        while (!mbar_try(ea, ep)) if (++spin > (1 << 22)) { *fail = 1; break; }
                    }
                }
                if (elect_one()) {
                    mbar_expect_tx(&full[issued % kCsBuf], BLKB);
                    bulk_g2s(sE + (issued % kCsBuf) * BLKB, (const unsigned char *)P.Ec + (long long)issued * BLKB,
                             BLKB, &full[issued % kCsBuf]);
                }
                __syncwarp();
                ++issued;
            }
        };
        for (int b = 0; b < nblk; ++b) {
            if (FOLD && warp == 0) try_issue(b);
            if (landed < b) {
                if (!mbar_wait2(&full[b % kCsBuf], (b / kCsBuf) & 1, fail)) *fail = 1;
                landed = b;
            }
            const uint32_t end_b = (uint32_t)(b + 1) * ppb;
            limit = (uint32_t)(landed + 1) * ppb;
            bool polled = false;
            // until no lane of the warp has a sample of block b left (in a buffer or in its records)
            while (__any_sync(GT_FULL, pair < end_b || pairA < end_b)) {
                // a lane that is done with block b runs ahead into the other blocks of the ring once they have landed
                if (polled) { ++landed; limit += ppb; polled = false; }
                if (FOLD && warp == 0 && issued < nblk && issued < b + kCsBuf) try_issue(b);
                load_next(vB, pairB);
                accumulate(vA);
                advance();
                if (landed < b + kCsBuf - 1 && landed + 1 < nblk)
                    polled = mbar_test(smem_u32(&full[(landed + 1) % kCsBuf]), ((landed + 1) / kCsBuf) & 1);
                load_next(vA, pairA);
                accumulate(vB);
                advance();
#if GT_CS_UNROLL4
                // a second pair of samples before the warp votes again (a lane that is out of
                // samples of the blocks it may touch adds zeros)
                load_next(vB, pairB);
                accumulate(vA);
                advance();
                load_next(vA, pairA);
                accumulate(vB);
                advance();
#endif
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[b % kCsBuf]);
        }
        accumulate(vA);                  // a row of the last block may still wait in the buffer
        asm volatile("cp.async.wait_group 0;\n" ::: "memory");
        // ===== epilogue: derive the constant-direction column, write the upper triangle =====
        if (owner) {
            const int L = P.L;
            double *out = P.corrT + snp;
            const long long ld = P.ldc;
            const double dn = (double)n_miss;
            const int one = P.one;
            const double *sm = sum;
            // streamed x streamed, streamed x ones
            int e = 0;
#pragma unroll
            for (int a = 0; a < NV; ++a) {
#pragma unroll
                for (int c = a; c < NV; ++c, ++e) {
                    const int ia = P.vcol[a], ic = P.vcol[c];
                    out[(long long)cs_tri(min(ia, ic), max(ia, ic), L) * ld] = acc[e];
                }
                out[(long long)cs_tri(min(P.vcol[a], one), max(P.vcol[a], one), L) * ld] = sm[a];
            }
            out[(long long)cs_tri(one, one, L) * ld] = dn;
            if (P.drv_col >= 0) {
                // v_j = (1 - sum_c a_c v_c) / a_j on every design row
                const int j = P.drv_col;
                double s1 = 0.0, s2 = 0.0;       // sum_c a_c sm_c,  sum_cd a_c a_d C_cd
#pragma unroll
                for (int a = 0; a < NV; ++a) {
                    s1 = fma(P.drv_a[a], sm[a], s1);
                    double t = 0.0;              // sum_c a_c C_ca
#pragma unroll
                    for (int c = 0; c < NV; ++c) {
                        const int lo = a < c ? a : c, hi = a < c ? c : a;
                        t = fma(P.drv_a[c], acc[lo * NV - lo * (lo - 1) / 2 + (hi - lo)], t);
                    }
                    s2 = fma(P.drv_a[a], t, s2);
                    const int ia = P.vcol[a];
                    out[(long long)cs_tri(min(ia, j), max(ia, j), L) * ld] = (sm[a] - t) * P.drv_inv;
                }
                out[(long long)cs_tri(min(j, one), max(j, one), L) * ld] = (dn - s1) * P.drv_inv;
                out[(long long)cs_tri(j, j, L) * ld] = (dn - 2.0 * s1 + s2) * P.drv_inv * P.drv_inv;
            }
            P.counts[(long long)snp * 4 + 2] = n_miss;
        }
    }
    __syncthreads();
    if (tid == 0 && *fail) atomicExch(P.error_flag, 1);
}

}  // namespace gtk
