// Sex-aware allele frequency (sexual_chromosome=True): the reference's
// get_sex_maf (genetest/statistics/descriptive.py:51-93) for every marker of
// a batch, ahead of the model kernels, which then take MAF and the flip flag
// from here instead of nanmean(g) / 2.
//
//   maf = (sum g - 0.5 * sum_{males} g) / (n_males + 2 n_females)
//
// over the samples with a genotype, a design row and a known sex; NaN (and no
// flip) when there is none.  One warp per marker.
#pragma once

#include "gt_common.cuh"

namespace gtk {

__device__ __forceinline__ void sex_maf_finish(double dot, double sum, double n_m, double n_f,
                                               double *maf_o, int *flip_o, int snp) {
    double maf = nan("");
    int flip = 0;
    if (n_m + n_f > 0.0) {
        maf = (-0.5 * dot + sum) / (n_m + 2.0 * n_f);
        if (maf > 0.5) { maf = 1.0 - maf; flip = 1; }
    }
    maf_o[snp] = maf;
    flip_o[snp] = flip;
}

// male / female: bit 2s of byte s/4... same 2-bit geometry as the .bed row,
// 01 in the pair of every male (female) design sample.
__global__ void __launch_bounds__(256)
sex_maf_packed_kernel(const uint8_t *__restrict__ geno, long long pitch, int n_snps, int n_file,
                      const uint8_t *__restrict__ ormask, const uint8_t *__restrict__ male,
                      const uint8_t *__restrict__ female, double *__restrict__ maf_o,
                      int *__restrict__ flip_o) {
    const int lane = threadIdx.x & 31;
    const int snp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (snp >= n_snps) return;
    const uint4 *row = (const uint4 *)(geno + (long long)snp * pitch);
    const uint4 *om = (const uint4 *)ormask, *mm = (const uint4 *)male, *fm = (const uint4 *)female;
    // the masks are gt_packed_pitch(n_file) bytes long; a caller's pitch may be larger
    const long long mask_pitch = (((long long)n_file + 3) / 4 + 127) / 128 * 128;
    const int nvec = (int)((pitch < mask_pitch ? pitch : mask_pitch) / 16);
    int het_m = 0, hom_m = 0, n_m = 0, het_f = 0, hom_f = 0, n_f = 0;
    auto word = [&](unsigned w, unsigned o, unsigned m, unsigned f) {
        const unsigned wm = w | o;
        const unsigned lo = wm & 0x55555555u, hi = (wm >> 1) & 0x55555555u;
        const unsigned present = ~(lo & ~hi) & 0x55555555u;   // not code 01
        const unsigned het = hi & ~lo, hom = ~(lo | hi) & 0x55555555u;
        het_m += __popc(het & m); hom_m += __popc(hom & m); n_m += __popc(present & m);
        het_f += __popc(het & f); hom_f += __popc(hom & f); n_f += __popc(present & f);
    };
    for (int v = lane; v < nvec; v += 32) {
        const uint4 w = row[v], o = __ldg(om + v), m = __ldg(mm + v), f = __ldg(fm + v);
        word(w.x, o.x, m.x, f.x); word(w.y, o.y, m.y, f.y);
        word(w.z, o.z, m.z, f.z); word(w.w, o.w, m.w, f.w);
    }
    het_m = gt_warp_sum_int(het_m); hom_m = gt_warp_sum_int(hom_m); n_m = gt_warp_sum_int(n_m);
    het_f = gt_warp_sum_int(het_f); hom_f = gt_warp_sum_int(hom_f); n_f = gt_warp_sum_int(n_f);
    if (lane == 0) {
        const double dot = (double)het_m + 2.0 * (double)hom_m;
        const double sum = dot + (double)het_f + 2.0 * (double)hom_f;
        sex_maf_finish(dot, sum, (double)n_m, (double)n_f, maf_o, flip_o, snp);
    }
}

// dosage rows aligned to the design rows; sex[i] = 0 / 1 / NaN per design row
__global__ void __launch_bounds__(256)
sex_maf_dosage_kernel(const double *__restrict__ geno, long long ld, int n_snps, int n,
                      const double *__restrict__ sex, double *__restrict__ maf_o,
                      int *__restrict__ flip_o) {
    const int lane = threadIdx.x & 31;
    const int snp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (snp >= n_snps) return;
    const double *row = geno + (long long)snp * ld;
    double dot = 0.0, sum = 0.0, n_m = 0.0, n_all = 0.0;
    for (int i = lane; i < n; i += 32) {
        const double g = row[i], s = __ldg(sex + i);
        if (g == g && s == s) { dot = fma(g, s, dot); sum += g; n_m += s; n_all += 1.0; }
    }
    dot = gt_warp_sum(dot); sum = gt_warp_sum(sum); n_m = gt_warp_sum(n_m); n_all = gt_warp_sum(n_all);
    if (lane == 0) sex_maf_finish(dot, sum, n_m, n_all - n_m, maf_o, flip_o, snp);
}

}  // namespace gtk
