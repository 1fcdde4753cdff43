// C ABI of the genetest_b200 engine (see include/genetest_b200.h).
// Host side: design factorisation (QR of the covariates, full-sample Gram
// matrix, E matrix in genotype-file order), workspace management, batching,
// kernel dispatch, timing.  No torch types, no exceptions across the ABI.
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "gt_common.cuh"
#include "gt_host.cuh"
#include "gt_contract.cuh"
#include "gt_contract_tc.cuh"
#include "gt_contract_tc2.cuh"
#include "gt_corr_stream.cuh"
#include "gt_linear.cuh"
#include "gt_solve.cuh"
#include "gt_solve_thread.cuh"
#include "gt_logistic.cuh"
#include "gt_logistic_thread.cuh"
#include "gt_cox.cuh"
#include "gt_format.cuh"
#include "gt_sexmaf.cuh"

struct gt_engine {
    int device = 0;
    cudaStream_t stream = nullptr, own_stream = nullptr;
    bool have_design = false;
    int model = GT_MODEL_LINEAR;
    GtDesignDev D{};
    gtk::IterDesignDev T{};           // logistic / coxph design
    DevBuf dE, dMask, dFF0, dR, dRinv, dTq, dIter;   // design
    DevBuf dBimg, dScale, dErr;                      // tcgen05 contraction (lean designs)
    DevBuf dEi;                                      // compact design rows of the streaming corrections
    gtk::CorrStreamParams cs{};
    int cs_nv = 0;             // columns the streaming correction kernel gathers (0: not available)
    bool use_tc = true;        // lean linear designs run on the tcgen05 int8 contraction
    bool use_stream_corr = true;   // ... and their corrections on the streaming kernel
    int cox_maxiter = 100;     // Newton step limit of the Cox model (statsmodels' default)
    bool use_thread_solve = true;  // thread-per-SNP solve for narrow designs without interaction
    bool use_thread_iter = true;   // thread-per-SNP IRLS for binary outcomes without interaction
    int thread_fold = 1;           // thread kernels without a producer warp: 256 markers per CTA (option "thread_fold")
    int thread_iter_min = 4096;    // ... for batches of at least this many markers (fewer: warp per SNP)
    DevBuf dIterScratch, dIterT;
    int tc_dbg = 0;
    int tc_N = 0;              // 0: the design does not run on the tcgen05 kernel
    int tc_hom = 0;            // interaction design: MMA tiles of 64 markers x (dosage, homozygote) rows
    gtk::Tc2Params tc{};       // constant part of the kernel parameters
    DevBuf dS, dCounts, dGsum, dMlist, dMcount, dCorr; // per-batch workspace
    DevBuf dGeno, dOut, dIout;                 // staging of the *_host entry points
    cudaStream_t s_in = nullptr, s_out = nullptr;     // copy streams of the *_host pipeline
    cudaEvent_t ev_in[2] = {nullptr, nullptr}, ev_done[2] = {nullptr, nullptr}, ev_start = nullptr;
    int host_chunk = 0;        // option "host_chunk": SNPs per pipeline stage (0 = automatic)
    void *pin[2] = {nullptr, nullptr};   // pinned staging of pageable sources (one per pipeline slot)
    size_t pin_bytes = 0;
    cudaEvent_t ev_h2d[2] = {nullptr, nullptr};
    double *file_maf_host = nullptr;     // gt_engine_set_file_maf_output
    DevBuf dFileMaf;
    int contract_idx = -1;
    // sample alignment of the current design; sex-aware MAF (sexual_chromosome=True)
    std::vector<int> pos;
    bool identity_pos = true, sex_maf = false;
    DevBuf dSexM, dSexF, dSexRow, dMafO, dFlipO;
    DevBuf dCounter;           // work counter of the persistent iterative kernels
    // timing
    bool timing = false;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> events;
    std::vector<int> event_class;      // 0 contraction, 1 corrections, 2 finish / iterative
    double ms_class[4] = {0, 0, 0, 0};
    size_t events_used = 0;
    double ms_total = 0.0;
    int64_t launches = 0, all_launches = 0;
    std::string dominant, dominant_tc;
};

// ---------------------------------------------------------------------------
// contraction kernel instantiations
// ---------------------------------------------------------------------------
typedef void (*contract_fn)(const uint8_t *, long long, int, const uint8_t *, const double *, int,
                            double *, int *, double *);
struct ContractEntry { int NCA, NCB, rows, warps, smem; contract_fn fn; };

// (NCA, NCB, ROWS, WARPS, MINB): register tile ROWS x (NCA + NCB) doubles per
// thread; small tiles run 128-thread CTAs, three per SM.
#define GT_ENTRY(A, B, R, W, M)                                              \
    { A, B, R, W, gtk::ContractCfg<A, B, R, W>::kSmem,                        \
      gtk::contract_packed_kernel<A, B, R, W, M> }
static const ContractEntry kContract[] = {
    // lean designs (no interaction): E = [q | y]
    GT_ENTRY(4, 0, 4, 4, 3),  GT_ENTRY(6, 0, 4, 4, 3),  GT_ENTRY(8, 0, 4, 4, 3),
    GT_ENTRY(10, 0, 4, 4, 3), GT_ENTRY(12, 0, 4, 4, 3), GT_ENTRY(14, 0, 4, 4, 2),
    GT_ENTRY(16, 0, 4, 4, 2), GT_ENTRY(20, 0, 3, 4, 2), GT_ENTRY(24, 0, 2, 4, 3),
    GT_ENTRY(28, 0, 2, 4, 2), GT_ENTRY(32, 0, 2, 4, 2),
    // one interaction
    GT_ENTRY(12, 4, 4, 4, 2), GT_ENTRY(20, 4, 3, 4, 2), GT_ENTRY(28, 4, 2, 4, 2),
    GT_ENTRY(36, 4, 1, 4, 3), GT_ENTRY(48, 4, 1, 4, 2),
    // two / three interactions
    GT_ENTRY(18, 6, 3, 4, 2), GT_ENTRY(30, 6, 2, 4, 2), GT_ENTRY(42, 6, 1, 4, 2),
    GT_ENTRY(54, 6, 1, 4, 2),
    GT_ENTRY(32, 10, 1, 4, 2), GT_ENTRY(48, 10, 1, 4, 2), GT_ENTRY(64, 10, 1, 4, 1),
};
static const int kNumContract = sizeof(kContract) / sizeof(kContract[0]);

static int pick_contract(int NC1, int NC2) {
    int best = -1;
    for (int i = 0; i < kNumContract; ++i) {
        if (kContract[i].NCA >= NC1 && kContract[i].NCB >= NC2 && (kContract[i].NCB == 0) == (NC2 == 0)) {
            if (best < 0 || kContract[i].NCA + kContract[i].NCB < kContract[best].NCA + kContract[best].NCB)
                best = i;
        }
    }
    return best;
}

// ---------------------------------------------------------------------------
// small kernels: t quantile table, synthetic .bed rows
// ---------------------------------------------------------------------------
__global__ void tq_table_kernel(double *tq, int n) {
    int df = blockIdx.x * blockDim.x + threadIdx.x;
    if (df <= n) tq[df] = (df > 0) ? gt_t_ppf975((double)df) : nan("");
}

__device__ __forceinline__ uint64_t splitmix(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
__device__ __forceinline__ double u01(uint64_t h) { return (double)(h >> 11) * (1.0 / 9007199254740992.0); }

__global__ void synth_packed_kernel(uint8_t *packed, long long pitch, int n_file, int n_snps,
                                    uint64_t seed, double maf_lo, double maf_hi, double miss) {
    long long bps = (n_file + 3) / 4;
    long long total = (long long)n_snps * bps;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
         t += (long long)gridDim.x * blockDim.x) {
        int snp = (int)(t / bps);
        int byte = (int)(t % bps);
        double maf = maf_lo + (maf_hi - maf_lo) * u01(splitmix(seed ^ (0xABCDull + (uint64_t)snp * 0x100000001B3ull)));
        uint8_t v = 0;
        for (int j = 0; j < 4; ++j) {
            int s = byte * 4 + j;
            if (s >= n_file) break;   // PLINK pads the last byte with zero bits
            uint64_t h = splitmix(seed + ((uint64_t)snp << 32) + (uint64_t)s);
            uint64_t h2 = splitmix(h);
            uint64_t h3 = splitmix(h2);
            int g = (u01(h) < maf) + (u01(h2) < maf);    // copies of the coded allele A1
            uint8_t code = g == 2 ? 0 : (g == 1 ? 2 : 3);
            if (u01(h3) < miss) code = 1;
            v |= code << (2 * j);
        }
        packed[(long long)snp * pitch + byte] = v;
    }
}

// FP64 FMA peak probe (roofline denominator for the FP64-bound kernels;
// MEASURED_PEAKS.json has no FP64 figure).
__global__ void fp64_peak_kernel(double *sink, int iters, double b, double c) {
    double a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5,
           a6 = a0 + 6, a7 = a0 + 7;
    for (int i = 0; i < iters; ++i) {
        a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
        a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
    }
    double s = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    if (s == 123.456) sink[0] = s;
}

__global__ void dmma_peak_kernel(double *sink, int iters, double a, double b) {
    double c[8][2];
#pragma unroll
    for (int i = 0; i < 8; ++i) { c[i][0] = threadIdx.x + i; c[i][1] = i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) gtk::dmma884(c[i][0], c[i][1], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
    if (s == 123.456) sink[0] = s;
}

// Access-pattern probe (tests / tuning only): 128 rows per CTA, `tpr` threads
// per row, each thread streams `vec` x 16 B per step with `ahead` steps in
// flight, XOR-reducing what it reads.
__global__ void __launch_bounds__(512, 1)
read_pattern_kernel(const uint8_t *geno, long long pitch, int n_rows, int tpr, int vec,
                    unsigned *sink) {
    const int row_in = threadIdx.x % 128, part = threadIdx.x / 128;
    if (part >= tpr) return;
    const int row = min(blockIdx.x * 128 + row_in, n_rows - 1);
    const uint4 *p = (const uint4 *)(geno + (long long)row * pitch);
    const long long nvec = pitch / 16;
    unsigned acc = 0;
    // step = tpr * vec uint4 per row; this thread takes [part*vec, part*vec+vec)
    for (long long base = 0; base + (long long)tpr * vec <= nvec; base += (long long)tpr * vec) {
        for (int v = 0; v < vec; ++v) {
            uint4 x = gtk::ldg_stream16(p + base + part * vec + v);
            acc ^= x.x ^ x.y ^ x.z ^ x.w;
        }
    }
    if (acc == 0x12345678u) sink[0] = acc;
}

// MAF of every marker over ALL samples of its file row (what MAFFilter sees:
// get_maf(snp.genotypes), predicates.py:49 / descriptive.py:19-48): exact class
// counts, NaN when every genotype is missing.  One warp per marker.
__global__ void __launch_bounds__(256)
file_maf_packed_kernel(const uint8_t *__restrict__ geno, long long pitch, int n_snps, int n_file,
                       double *__restrict__ maf_o) {
    const int lane = threadIdx.x & 31;
    const int snp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (snp >= n_snps) return;
    const uint32_t *row = (const uint32_t *)(geno + (long long)snp * pitch);
    const int nwords = (n_file + 15) / 16;
    int het = 0, hom = 0, mis = 0;
    for (int w = lane; w < nwords; w += 32) {
        uint32_t x = row[w];
        const int left = n_file - w * 16;              // samples in this word
        if (left < 16) x |= 0xFFFFFFFFu << (2 * left);   // the padding reads as 11 (not counted)
        const uint32_t lo = x & 0x55555555u, hi = (x >> 1) & 0x55555555u;
        het += __popc(hi & ~lo); hom += __popc(~(lo | hi) & 0x55555555u); mis += __popc(lo & ~hi);
    }
    het = gt_warp_sum_int(het); hom = gt_warp_sum_int(hom); mis = gt_warp_sum_int(mis);
    if (lane == 0) {
        const int present = n_file - mis;
        double maf = nan("");
        if (present > 0) {
            maf = ((double)(het + 2 * hom) / (double)present) / 2.0;
            if (maf > 0.5) maf = 1.0 - maf;
        }
        maf_o[snp] = maf;
    }
}

// ---------------------------------------------------------------------------
static int tc_tiles(const gt_engine *e);
extern "C" {

int gt_engine_set_file_maf_output(gt_engine *e, double *maf_host) {
    if (!e) return fail(GT_E_ARG, "engine is NULL");
    e->file_maf_host = maf_host;
    return 0;
}

int gt_debug_read_pattern(gt_engine *e, const uint8_t *packed_dev, int64_t pitch, int32_t n_rows,
                          int tpr, int vec, int ctas_per_sm, double *gbs) {
    if (!e || !gbs) return fail(GT_E_ARG, "gt_debug_read_pattern: NULL argument");
    CK(cudaSetDevice(e->device));
    int rc = e->dS.reserve(64);
    if (rc) return rc;
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    int grid = (n_rows + 127) / 128;
    (void)ctas_per_sm;
    double best = 0;
    for (int rep = 0; rep < 3; ++rep) {
        CK(cudaEventRecord(a, e->stream));
        read_pattern_kernel<<<grid, 128 * tpr, 0, e->stream>>>(packed_dev, pitch, n_rows, tpr, vec,
                                                              (unsigned *)e->dS.p);
        CK(cudaEventRecord(b, e->stream));
        CK(cudaEventSynchronize(b));
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, a, b));
        double g = (double)n_rows * pitch / (ms * 1e-3) / 1e9;
        if (g > best) best = g;
    }
    cudaEventDestroy(a); cudaEventDestroy(b);
    *gbs = best;
    return 0;
}

int gt_abi_version(void) { return GT_ABI_VERSION; }

// rows [r0, r1) into [p, end); cursor[c] = first string of row r0 for the
// string columns.  Returns the end of the text, or NULL (bad column type /
// out of room; *bad = offending type or 0).
static char *format_row_range(int r0, int r1, int32_t n_cols, const int32_t *col_type,
                              const void *const *col_data, const uint8_t *keep, char sep,
                              std::vector<const char *> cursor, char *p, char *end, int *bad) {
    *bad = 0;
    for (int r = r0; r < r1; ++r) {
        const bool emit = !keep || keep[r];
        for (int c = 0; c < n_cols; ++c) {
            if (col_type[c] == 0) {                 // NUL-separated strings, one per row
                const char *s = cursor[c];
                size_t len = std::strlen(s);
                cursor[c] = s + len + 1;
                if (!emit) continue;
                if (p + len + 2 > end) return nullptr;
                std::memcpy(p, s, len); p += len;
            } else {
                if (!emit) continue;
                if (p + 48 > end) return nullptr;
                if (col_type[c] == 1) p = gtk::format_double_repr(((const double *)col_data[c])[r], p);
                else if (col_type[c] == 2) {
                    auto res = std::to_chars(p, p + 24, ((const int64_t *)col_data[c])[r]);
                    p = res.ptr;
                } else if (col_type[c] == 3) {      // one constant string for every row
                    const char *s = (const char *)col_data[c];
                    size_t len = std::strlen(s);
                    if (p + len + 2 > end) return nullptr;
                    std::memcpy(p, s, len); p += len;
                } else { *bad = col_type[c] ? col_type[c] : -1; return nullptr; }
            }
            *p++ = (c + 1 < n_cols) ? sep : '\n';
        }
    }
    return p;
}

int64_t gt_format_rows(int32_t n_rows, int32_t n_cols, const int32_t *col_type,
                       const void *const *col_data, const uint8_t *keep, char sep,
                       char *out, int64_t out_cap) {
    if (!col_type || !col_data || !out) return fail(GT_E_ARG, "gt_format_rows: NULL argument");
    if (n_rows <= 0) return 0;
    // Row ranges are formatted by a few host threads, each into its own
    // worst-case slice of `out`; the slices are then closed up in order.
    int nt = (int)std::min<unsigned>(std::max(1u, std::thread::hardware_concurrency()), 16u);
    nt = std::max(1, std::min(nt, n_rows / 4096));
    int64_t fixed = 2;                              // worst-case bytes of a row without its strings
    for (int c = 0; c < n_cols; ++c) {
        if (col_type[c] == 1 || col_type[c] == 2) fixed += 26;
        else if (col_type[c] == 3) fixed += (int64_t)std::strlen((const char *)col_data[c]) + 1;
        else if (col_type[c] == 0) fixed += 0;          // text + NUL are counted from the cursors
        else return fail(GT_E_ARG, "gt_format_rows: bad column type %d", col_type[c]);
    }
    std::vector<std::vector<const char *>> starts(nt + 1, std::vector<const char *>(n_cols, nullptr));
    {
        std::vector<const char *> cursor(n_cols, nullptr);
        for (int c = 0; c < n_cols; ++c)
            if (col_type[c] == 0) cursor[c] = (const char *)col_data[c];
        int next = 0;
        for (int r = 0;; ++r) {
            while (next <= nt && r == (int)((int64_t)n_rows * next / nt)) starts[next++] = cursor;
            if (next > nt) break;               // r == n_rows: every boundary is set
            for (int c = 0; c < n_cols; ++c)
                if (col_type[c] == 0) cursor[c] += std::strlen(cursor[c]) + 1;
        }
    }
    std::vector<int64_t> off(nt + 1, 0);
    for (int t = 0; t < nt; ++t) {
        const int r0 = (int)((int64_t)n_rows * t / nt), r1 = (int)((int64_t)n_rows * (t + 1) / nt);
        int64_t cap = (int64_t)(r1 - r0) * fixed + 64;
        for (int c = 0; c < n_cols; ++c)
            if (col_type[c] == 0) cap += starts[t + 1][c] - starts[t][c];
        off[t + 1] = off[t] + cap;
    }
    if (off[nt] > out_cap) return fail(GT_E_ARG, "gt_format_rows: output buffer too small (%lld < %lld)", (long long)out_cap, (long long)off[nt]);
    std::vector<char *> ends(nt, nullptr);
    std::vector<int> bad(nt, 0);
    auto work = [&](int t) {
        const int r0 = (int)((int64_t)n_rows * t / nt), r1 = (int)((int64_t)n_rows * (t + 1) / nt);
        ends[t] = format_row_range(r0, r1, n_cols, col_type, col_data, keep, sep, starts[t],
                                   out + off[t], out + off[t + 1], &bad[t]);
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < nt; ++t) pool.emplace_back(work, t);
    work(0);
    for (auto &th : pool) th.join();
    char *p = out;
    for (int t = 0; t < nt; ++t) {
        if (!ends[t]) return fail(GT_E_ARG, "gt_format_rows: formatting failed (column type %d)", bad[t]);
        const size_t len = (size_t)(ends[t] - (out + off[t]));
        if (p != out + off[t]) std::memmove(p, out + off[t], len);
        p += len;
    }
    return (int64_t)(p - out);
}
// host evaluations of the formulas the kernels use for p-values and
// confidence intervals (GT_HD functions of gt_common.cuh): no GPU needed
double gt_host_t_sf2(double t, double df) { return gt_t_sf2(t, df); }
double gt_host_t_ppf975(double df) { return gt_t_ppf975(df); }
double gt_host_norm_sf2(double z) { return gt_norm_sf2(z); }
double gt_host_lt_exp_neg(double a) { return gtk::lt_exp_neg(a); }
double gt_host_lt_log1p01(double t) { double inv; return gtk::lt_log1p01(t, inv); }
const char *gt_last_error(void) { return g_err.c_str(); }

int64_t gt_packed_pitch(int32_t n_file) {
    int64_t bps = ((int64_t)n_file + 3) / 4;
    return (bps + 127) / 128 * 128;
}

int gt_engine_create(int device, gt_engine **out) {
    if (!out) return fail(GT_E_ARG, "gt_engine_create: out is NULL");
    int count = 0;
    CK(cudaGetDeviceCount(&count));
    if (device < 0 || device >= count) return fail(GT_E_ARG, "gt_engine_create: no CUDA device %d (found %d)", device, count);
    CK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fail(GT_E_UNSUPPORTED, "gt_engine_create: device %d is sm_%d%d; this library is built for sm_100a only",
                    device, prop.major, prop.minor);
    gt_engine *e = new gt_engine();
    e->device = device;
    CK(cudaStreamCreateWithFlags(&e->own_stream, cudaStreamNonBlocking));
    e->stream = e->own_stream;
    *out = e;
    return 0;
}

int gt_engine_destroy(gt_engine *e) {
    if (!e) return 0;
    cudaSetDevice(e->device);
    cudaStreamSynchronize(e->stream);
    for (auto &ev : e->events) { cudaEventDestroy(ev.first); cudaEventDestroy(ev.second); }
    DevBuf *bufs[] = {&e->dE, &e->dMask, &e->dFF0, &e->dR, &e->dRinv, &e->dTq, &e->dIter, &e->dBimg, &e->dScale, &e->dErr, &e->dEi,
                      &e->dS, &e->dCounts, &e->dGsum, &e->dMlist, &e->dMcount, &e->dCorr, &e->dGeno, &e->dOut, &e->dIout,
                      &e->dSexM, &e->dSexF, &e->dSexRow, &e->dMafO, &e->dFlipO, &e->dCounter, &e->dFileMaf, &e->dIterScratch, &e->dIterT};
    for (DevBuf *b : bufs) b->release();
    for (int i = 0; i < 2; ++i) {
        if (e->pin[i]) cudaFreeHost(e->pin[i]);
        if (e->ev_h2d[i]) cudaEventDestroy(e->ev_h2d[i]);
    }
    if (e->own_stream) cudaStreamDestroy(e->own_stream);
    if (e->s_in) cudaStreamDestroy(e->s_in);
    if (e->s_out) cudaStreamDestroy(e->s_out);
    for (int i = 0; i < 2; ++i) {
        if (e->ev_in[i]) cudaEventDestroy(e->ev_in[i]);
        if (e->ev_done[i]) cudaEventDestroy(e->ev_done[i]);
    }
    if (e->ev_start) cudaEventDestroy(e->ev_start);
    delete e;
    return 0;
}

int gt_engine_set_stream(gt_engine *e, void *cuda_stream) {
    if (!e) return fail(GT_E_ARG, "engine is NULL");
    // NULL selects the engine's own stream; pass cudaStreamLegacy (0x1) or
    // cudaStreamPerThread (0x2) to run on a default stream.
    e->stream = cuda_stream ? (cudaStream_t)cuda_stream : e->own_stream;
    return 0;
}

int gt_engine_synchronize(gt_engine *e) {
    if (!e) return fail(GT_E_ARG, "engine is NULL");
    CK(cudaSetDevice(e->device));
    CK(cudaStreamSynchronize(e->stream));
    if (e->dErr.p) {
        int flag = 0;
        CK(cudaMemcpy(&flag, e->dErr.p, 4, cudaMemcpyDeviceToHost));
        if (flag) {
            cudaMemset(e->dErr.p, 0, 4);
            return fail(GT_E_CUDA, "tcgen05 contraction: a pipeline barrier timed out");
        }
    }
    for (size_t i = 0; i < e->events_used; ++i) {
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, e->events[i].first, e->events[i].second));
        if (e->event_class[i] == 0) e->ms_total += ms;
        e->ms_class[e->event_class[i] & 3] += ms;
    }
    e->events_used = 0;
    return 0;
}

int gt_engine_set_option(gt_engine *e, const char *name, int value) {
    if (!e || !name) return fail(GT_E_ARG, "gt_engine_set_option: NULL argument");
    if (std::strcmp(name, "contract_tc") == 0) { e->use_tc = value != 0; return 0; }
    if (std::strcmp(name, "tc_debug") == 0) { e->tc_dbg = value; return 0; }
    if (std::strcmp(name, "stream_corrections") == 0) { e->use_stream_corr = value != 0; return 0; }
    if (std::strcmp(name, "thread_solve") == 0) { e->use_thread_solve = value != 0; return 0; }
    if (std::strcmp(name, "thread_iter") == 0) { e->use_thread_iter = value != 0; return 0; }
    if (std::strcmp(name, "thread_fold") == 0) { e->thread_fold = value != 0; return 0; }
    if (std::strcmp(name, "thread_iter_min") == 0) { e->thread_iter_min = value > 0 ? value : 4096; return 0; }
    if (std::strcmp(name, "cox_maxiter") == 0) { e->cox_maxiter = value > 0 ? value : 100; return 0; }
    if (std::strcmp(name, "host_chunk") == 0) { e->host_chunk = (int)std::max<int64_t>(0, value); return 0; }
    return fail(GT_E_ARG, "gt_engine_set_option: unknown option '%s'", name);
}

static int e_n_file(const gt_engine *e) { return e->model == GT_MODEL_LINEAR ? e->D.n_file : e->T.n_file; }

int gt_engine_set_sample_sex(gt_engine *e, const double *sex) {
    if (!e || !e->have_design) return fail(GT_E_STATE, "gt_engine_set_sample_sex: no design set");
    if (!sex) { e->sex_maf = false; return 0; }
    CK(cudaSetDevice(e->device));
    const int n = (int)e->pos.size();
    const int64_t pitch = gt_packed_pitch(e_n_file(e));
    std::vector<uint8_t> male((size_t)pitch, 0), female((size_t)pitch, 0);
    for (int i = 0; i < n; ++i) {
        const int ps = e->pos[i];
        const uint8_t bit = (uint8_t)(1u << (2 * (ps & 3)));
        if (sex[i] == 1.0) male[ps >> 2] |= bit;
        else if (sex[i] == 0.0) female[ps >> 2] |= bit;
        else if (sex[i] == sex[i])
            return fail(GT_E_ARG, "gt_engine_set_sample_sex: sex[%d]=%g (expected 0, 1 or NaN)", i, sex[i]);
    }
    int rc;
    if ((rc = upload(e->dSexM, male.data(), male.size(), e->stream))) return rc;
    if ((rc = upload(e->dSexF, female.data(), female.size(), e->stream))) return rc;
    if ((rc = upload(e->dSexRow, sex, (size_t)n * 8, e->stream))) return rc;
    CK(cudaStreamSynchronize(e->stream));
    e->sex_maf = true;
    return 0;
}

int gt_engine_enable_timing(gt_engine *e, int on) {
    if (!e) return fail(GT_E_ARG, "engine is NULL");
    e->timing = on != 0;
    e->ms_total = 0.0; e->launches = 0; e->all_launches = 0; e->events_used = 0;
    for (double &m : e->ms_class) m = 0.0;
    return 0;
}

int gt_engine_kernel_time(gt_engine *e, double *ms_total, int64_t *launches, int64_t *all_launches) {
    if (!e) return fail(GT_E_ARG, "engine is NULL");
    if (ms_total) *ms_total = e->ms_total;
    if (launches) *launches = e->launches;
    if (all_launches) *all_launches = e->all_launches;
    return 0;
}

const char *gt_engine_dominant_kernel(const gt_engine *e) {
    if (!e) return "";
    if (e->have_design && e->model == GT_MODEL_LINEAR && e->use_tc && e->tc_N > 0) return e->dominant_tc.c_str();
    if (e->have_design && e->model == GT_MODEL_LOGISTIC && e->use_thread_iter && e->T.I <= 1 && e->T.p >= 2 + e->T.I &&
        e->T.p <= 13 && e->T.y_binary && !e->T.degenerate)
        return "logistic_thread_kernel";       // (calls of at least thread_iter_min .bed markers)
    if (e->have_design && e->model == GT_MODEL_COXPH && e->use_thread_iter && e->T.th_ok)
        return "cox_thread_kernel";
    return e->dominant.c_str();
}

int gt_engine_kernel_breakdown(gt_engine *e, double *ms4) {
    if (!e || !ms4) return fail(GT_E_ARG, "gt_engine_kernel_breakdown: NULL argument");
    for (int i = 0; i < 4; ++i) ms4[i] = e->ms_class[i];
    return 0;
}

int gt_engine_num_params(const gt_engine *e) {
    if (!e || !e->have_design) return fail(GT_E_STATE, "no design set");
    return e->model == GT_MODEL_LINEAR ? e->D.p : e->T.p;
}
int gt_engine_num_fields(const gt_engine *e) {
    int p = gt_engine_num_params(e);
    return p < 0 ? p : GT_F_PARAM0 + 6 * p;
}

int gt_engine_set_design(gt_engine *e, const gt_design_desc *d) {
    if (!e || !d) return fail(GT_E_ARG, "gt_engine_set_design: NULL argument");
    if (d->n <= 0 || d->k < 0 || d->n_inter < 0 || !d->y || (d->k > 0 && !d->C))
        return fail(GT_E_ARG, "gt_engine_set_design: bad sizes (n=%d k=%d n_inter=%d)", d->n, d->k, d->n_inter);
    if (d->n_inter > 0 && !d->W) return fail(GT_E_ARG, "gt_engine_set_design: n_inter > 0 but W is NULL");
    if (d->model == GT_MODEL_COXPH && !d->event) return fail(GT_E_ARG, "gt_engine_set_design: coxph needs event");
    const int n = d->n, k = d->k, I = d->n_inter;
    const int n_file = d->sample_pos ? d->n_file : (d->n_file > 0 ? d->n_file : n);
    if (n_file < n) return fail(GT_E_ARG, "gt_engine_set_design: n_file (%d) < n (%d)", n_file, n);
    if (k + 1 + I > 32) return fail(GT_E_UNSUPPORTED, "gt_engine_set_design: at most 32 design columns (got %d)", k + 1 + I);
    CK(cudaSetDevice(e->device));
    std::vector<int> pos(n);
    {
        std::vector<char> seen(n_file, 0);
        for (int i = 0; i < n; ++i) {
            int ps = d->sample_pos ? d->sample_pos[i] : i;
            if (ps < 0 || ps >= n_file || seen[ps])
                return fail(GT_E_ARG, "gt_engine_set_design: sample_pos[%d]=%d invalid or duplicated", i, ps);
            seen[ps] = 1;
            pos[i] = ps;
        }
    }
    e->have_design = false;
    e->model = d->model;
    e->sex_maf = false;
    e->identity_pos = (n_file == n);
    for (int i = 0; i < n && e->identity_pos; ++i) e->identity_pos = pos[i] == i;
    e->pos = pos;
    cudaStream_t st = e->stream;
    const int64_t pitch = gt_packed_pitch(n_file);
    const int n_chunks = (n_file + gtk::kChunk - 1) / gtk::kChunk;
    const size_t n_pad = (size_t)n_chunks * gtk::kChunk;

    // mask of the 2-bit pairs that are not design rows (file padding included)
    std::vector<uint8_t> mask((size_t)pitch, 0xFF);
    for (int i = 0; i < n; ++i) mask[pos[i] >> 2] &= (uint8_t)~(3u << (2 * (pos[i] & 3)));
    int rc = upload(e->dMask, mask.data(), mask.size(), st);
    if (rc) return rc;

    if (d->model == GT_MODEL_LINEAR) {
        HostDesign H;
        H.n = n; H.k = k; H.I = I; H.L = k + 2 + I;
        qr_mgs2(d->C, n, k, H);
        // explicit constant column (statsmodels k_constant) -> centre y on it
        for (int j = 0; j < k && H.icpt_col < 0; ++j) {
            const double *c = d->C + (size_t)j * n;
            double lo = c[0], hi = c[0];
            for (int i = 1; i < n; ++i) { lo = std::min(lo, c[i]); hi = std::max(hi, c[i]); }
            if (lo == hi && lo != 0.0) H.icpt_col = j;
        }
        H.has_const = H.icpt_col >= 0;
        if (!H.has_const && k > 0 && !H.degenerate) {
            // implicit constant: ones in span(Q)
            long double res = n;
            for (int j = 0; j < k; ++j) {
                long double s = 0;
                for (int i = 0; i < n; ++i) s += H.Q[(size_t)j * n + i];
                res -= s * s;
            }
            if (res < 1e-9L * n) H.has_const = 1;
        }
        long double ysum = 0;
        for (int i = 0; i < n; ++i) ysum += d->y[i];
        double ybar = H.icpt_col >= 0 ? (double)(ysum / n) : 0.0;
        H.icpt_shift = H.icpt_col >= 0 ? ybar / d->C[(size_t)H.icpt_col * n] : 0.0;

        const int L = H.L;
        const bool lean = (I == 0);
        const int NC1 = lean ? k + 1 : (I + 1) * L, NC2 = lean ? 0 : (I + 1) * (I + 2) / 2;
        int ci = pick_contract(NC1, NC2);
        if (ci < 0) return fail(GT_E_UNSUPPORTED, "gt_engine_set_design: no contraction kernel for %d+%d columns (k=%d, interactions=%d)", NC1, NC2, k, I);
        const int NCA = kContract[ci].NCA, NCB = kContract[ci].NCB, NCP = NCA + NCB;
        e->contract_idx = ci;

        std::vector<double> V((size_t)n * L);      // row-major base vectors
        for (int i = 0; i < n; ++i) {
            double *v = &V[(size_t)i * L];
            for (int j = 0; j < k; ++j) v[j] = H.Q[(size_t)j * n + i];
            v[k] = d->y[i] - ybar;
            v[k + 1] = 1.0;
            for (int j = 0; j < I; ++j) v[k + 2 + j] = d->W[(size_t)j * n + i];
        }
        std::vector<long double> ff((size_t)L * L, 0.0L);
        for (int i = 0; i < n; ++i) {
            const double *v = &V[(size_t)i * L];
            for (int a = 0; a < L; ++a)
                for (int b = a; b < L; ++b) ff[(size_t)a * L + b] += (long double)v[a] * v[b];
        }
        std::vector<double> FF0((size_t)L * L);
        for (int a = 0; a < L; ++a)
            for (int b = a; b < L; ++b) FF0[(size_t)a * L + b] = FF0[(size_t)b * L + a] = (double)ff[(size_t)a * L + b];

        // two constants follow the rows: [0.0, 1.0] (see subtract_missing)
        std::vector<double> E(n_pad * NCP + 2, 0.0);
        E[n_pad * NCP + 1] = 1.0;
        for (int i = 0; i < n; ++i) {
            const double *v = &V[(size_t)i * L];
            double *er = &E[(size_t)pos[i] * NCP];
            if (lean) {
                for (int b = 0; b <= k; ++b) er[b] = v[b];
                continue;
            }
            for (int c = 0; c <= I; ++c) {
                double m = c == 0 ? 1.0 : v[k + 1 + c];
                for (int b = 0; b < L; ++b) er[c * L + b] = m * v[b];
            }
            int idx = 0;
            for (int a = 0; a <= I; ++a)
                for (int b = a; b <= I; ++b) {
                    double ma = a == 0 ? 1.0 : v[k + 1 + a], mb = b == 0 ? 1.0 : v[k + 1 + b];
                    er[NCA + idx++] = ma * mb;
                }
        }
        if ((rc = upload(e->dE, E.data(), E.size() * 8, st))) return rc;
        // ---- digit planes for the tcgen05 int8 contraction ----
        e->tc_N = 0; e->tc_hom = lean ? 0 : 1; e->cs_nv = 0;
        {
            const int ncol = NC1;
            std::vector<char> is_design(n_pad, 0);
            for (int i = 0; i < n; ++i) is_design[pos[i]] = 1;
            gtk::Tc2Params &T = e->tc;
            T = gtk::Tc2Params{};
            // If the constant vector is a combination of the E columns (a design with
            // an intercept: 1 = Q a), the column with the largest |a_j| needs no digit
            // planes: the kernel derives its sum from sum(g) and the other columns.
            T.ncolL = 0; T.drv_col = -1; T.drv_inv = 0.0; T.ones_out = -1; T.n_s2 = 0;
            std::vector<long double> acoef(ncol, 0.0L);
            if (H.has_const && !H.degenerate && k > 0) {
                for (int c = 0; c < k; ++c) {
                    long double sacc = 0;
                    for (int i = 0; i < n; ++i) sacc += H.Q[(size_t)c * n + i];
                    acoef[c] = sacc;
                }
                long double worst = 0;
                for (int i = 0; i < n; ++i) {
                    long double r = 1.0L;
                    for (int c = 0; c < k; ++c) r -= acoef[c] * H.Q[(size_t)c * n + i];
                    worst = std::max(worst, fabsl(r));
                }
                int j = 0;
                for (int c = 1; c < k; ++c) if (fabsl(acoef[c]) > fabsl(acoef[j])) j = c;
                if (worst <= 1e-12L && fabsl(acoef[j]) > 0.5L) { T.drv_col = j; T.drv_inv = (double)(1.0L / acoef[j]); }
            }
            // Interaction designs: E holds the products m_c v_b (m_0 = 1, m_c = w_c).  The column
            // of ones needs no digits (it is the ones plane), and equal columns (w_a = w_a * 1,
            // w_a w_b twice) share theirs.
            auto same_col = [&](int c1, int c2) {
                for (int i = 0; i < n; ++i)
                    if (E[(size_t)pos[i] * NCP + c1] != E[(size_t)pos[i] * NCP + c2]) return false;
                return true;
            };
            std::vector<int> colmap(ncol, -3);           // digit column of E column c; -1 ones plane, -2 derived
            std::vector<int> dig_src;                    // E column of digit column cl
            bool too_wide = false;
            for (int c = 0; c < ncol && !too_wide; ++c) {
                if (c == T.drv_col) { colmap[c] = -2; continue; }
                if (!lean) {
                    bool ones = true;
                    for (int i = 0; i < n && ones; ++i) ones = E[(size_t)pos[i] * NCP + c] == 1.0;
                    if (ones && T.ones_out < 0) { T.ones_out = c; colmap[c] = -1; continue; }
                    int twin = -1;
                    for (int cl = 0; cl < T.ncolL && twin < 0; ++cl)
                        if (T.out_col2[cl] < 0 && same_col(dig_src[cl], c)) twin = cl;
                    if (twin >= 0) { T.out_col2[twin] = c; colmap[c] = twin; continue; }
                }
                if (T.ncolL >= gtk::kT2MaxCols) { too_wide = true; break; }
                double mx = 0.0;
                for (int i = 0; i < n; ++i) mx = std::max(mx, std::fabs(E[(size_t)pos[i] * NCP + c]));
                T.out_col[T.ncolL] = c;
                T.out_col2[T.ncolL] = -1;
                T.scale[T.ncolL] = mx > 0.0 ? mx : 1.0;
                T.drv_a[T.ncolL] = c < k ? (double)acoef[c] : 0.0;
                colmap[c] = T.ncolL;
                dig_src.push_back(c);
                ++T.ncolL;
            }
            if (!lean && !too_wide) {
                // g^2 block: m_a m_b is the E column of block a, entry w_b (block 0: the base vector)
                T.n_s2 = NC2; T.s2_out = NCA;
                int idx = 0;
                for (int a = 0; a <= I && !too_wide; ++a)
                    for (int b = a; b <= I; ++b, ++idx) {
                        const int c = a * L + (b == 0 ? k + 1 : k + 1 + b);
                        if (colmap[c] < -1 || idx >= 12) { too_wide = true; break; }
                        T.s2_cl[idx] = colmap[c];
                    }
                T.pad0[0] = NC1; T.pad1[0] = NCA; T.pad0[1] = NCA + NC2; T.pad1[1] = NCP;
            }
            // seven digit planes per column + one plane of ones over the design
            // rows (its D column is the exact sum of the dosages)
            const int N = (gtk::kT2Digits * T.ncolL + 1 + 15) / 16 * 16;
            if (!too_wide && N <= 256 && (long long)n_file <= (1ll << 22)) {
                const int n_gst = (int)(pitch / 128);
                const int n_slots = 4 * n_gst;
                static const int perm16[16] = {0, 2, 4, 6, 8, 10, 12, 14, 1, 3, 5, 7, 9, 11, 13, 15};
                std::vector<uint8_t> Bimg((size_t)n_slots * N * 128, 0);
                auto put = [&](uint8_t *img, int nrow, int kk, uint8_t v) {
                    // K-major, 128B swizzle: 8-row groups 1024 B apart, 16-byte chunk ^ row
                    int grp = nrow >> 3, r = nrow & 7, chunk = kk >> 4, b = kk & 15;
                    img[(size_t)grp * 1024 + r * 128 + ((chunk ^ r) << 4) + b] = v;
                };
                for (int sl_i = 0; sl_i < n_slots; ++sl_i) {
                    uint8_t *img = &Bimg[(size_t)sl_i * N * 128];
                    for (int kk = 0; kk < 128; ++kk) {
                        // the expansion threads deliver the 16 samples of a word even ones first
                        long long smp = (long long)sl_i * 128 + (kk / 16) * 16 + perm16[kk % 16];
                        if (smp >= (long long)n_file || !is_design[smp]) continue;
                        const double *er = &E[(size_t)smp * NCP];
                        put(img, gtk::kT2Digits * T.ncolL, kk, 1);
                        for (int cl = 0; cl < T.ncolL; ++cl) {
                            long long Z = llround(std::ldexp(er[dig_src[cl]] / T.scale[cl], 54));
                            for (int dg = 0; dg < gtk::kT2Digits; ++dg) {
                                long long d8 = ((Z + 128) & 255) - 128;    // balanced base-256 digit
                                Z = (Z - d8) / 256;
                                put(img, cl * gtk::kT2Digits + dg, kk, (uint8_t)(int8_t)d8);
                            }
                        }
                    }
                }
                for (int cl = 0; cl < T.ncolL; ++cl) T.scale[cl] = std::ldexp(T.scale[cl], -54);
                if ((rc = upload(e->dBimg, Bimg.data(), Bimg.size(), st))) return rc;
                if ((rc = e->dErr.reserve(16))) return rc;
                CK(cudaMemsetAsync(e->dErr.p, 0, 16, st));
                int zl = 0;
                while ((size_t)(zl + 1) * 128 <= mask.size()) {
                    bool zero = true;
                    for (int b = 0; b < 128 && zero; ++b) zero = mask[(size_t)zl * 128 + b] == 0;
                    if (!zero) break;
                    ++zl;
                }
                T.n_gstages = n_gst; T.mask_zero_stages = zl;
                T.ormask = (const uint8_t *)e->dMask.p; T.Bimg = (const uint8_t *)e->dBimg.p;
                T.ncol = ncol; T.NCP = NCP; T.rcap = 16 * n_gst;
                T.error_flag = (int *)e->dErr.p;
                e->tc_N = N;
                // ---- compact rows for the streaming corrections: the base-vector columns
                //      that carry digits (block 0 of E, without the ones) ----
                std::vector<int> cs_cl;
                for (int cl = 0; cl < T.ncolL; ++cl) if (dig_src[cl] < L) cs_cl.push_back(cl);
                // (every entry of v must be streamed, derived or the constant)
                if ((int)cs_cl.size() <= gtk::kCsRow && (int)cs_cl.size() + (T.drv_col >= 0 ? 1 : 0) + 1 == L) {
                    const int nblk = (n_file + gtk::kCsBlk - 1) / gtk::kCsBlk;
                    std::vector<double> Ec((size_t)nblk * gtk::kCsBlk * gtk::kCsRow, 0.0);
                    gtk::CorrStreamParams &C = e->cs;
                    C = gtk::CorrStreamParams{};
                    for (int cc = 0; cc < (int)cs_cl.size(); ++cc) {
                        const int cl = cs_cl[cc];
                        C.vcol[cc] = dig_src[cl];
                        C.drv_a[cc] = T.drv_a[cl];
                        // 16-byte chunks (column pairs) of row i are stored at chunk ^ ((i >> 2) & 1)
                        for (int i = 0; i < n; ++i) {
                            const int ps = pos[i];
                            const int slot = (((cc >> 1) ^ ((ps >> 2) & 1)) << 1) | (cc & 1);
                            Ec[(size_t)ps * gtk::kCsRow + slot] = E[(size_t)ps * NCP + dig_src[cl]];
                        }
                    }
                    C.drv_col = T.drv_col; C.drv_inv = T.drv_inv;
                    C.n_blocks = nblk; C.L = L; C.one = k + 1; C.rcap = T.rcap;
                    if ((rc = upload(e->dEi, Ec.data(), Ec.size() * 8, st))) return rc;
                    C.Ec = (const double *)e->dEi.p;
                    C.error_flag = (int *)e->dErr.p;
                    e->cs_nv = (int)cs_cl.size();
                }
                CK(cudaStreamSynchronize(st));           // Bimg / Ei go out of scope
            }
        }
        if ((rc = upload(e->dFF0, FF0.data(), FF0.size() * 8, st))) return rc;
        if ((rc = upload(e->dR, H.R.data(), H.R.size() * 8, st))) return rc;
        if ((rc = upload(e->dRinv, H.Rinv.data(), H.Rinv.size() * 8, st))) return rc;
        if ((rc = e->dTq.reserve(((size_t)n + 1) * 8))) return rc;
        tq_table_kernel<<<(n + 1 + 127) / 128, 128, 0, st>>>((double *)e->dTq.p, n);
        CK(cudaGetLastError());

        GtDesignDev &D = e->D;
        D.n_file = n_file; D.n = n; D.k = k; D.I = I; D.L = L; D.p = k + 1 + I;
        D.NC1 = NC1; D.NC2 = NC2; D.NCA = NCA; D.NCB = NCB; D.NCP = NCP;
        D.NPROD = L * (L + 1) / 2;
        D.n_chunks = n_chunks;
        D.raw_mode = d->raw_mode; D.has_const = H.has_const; D.icpt_col = H.icpt_col;
        D.use_maf_t = d->use_maf_t; D.degenerate = H.degenerate ? 1 : 0;
        D.lean = lean ? 1 : 0;
        D.icpt_shift = H.icpt_shift;
        D.trG0 = 0.0;
        for (double r : H.R) D.trG0 += r * r;
        D.cond_t = d->condition_value_t; D.eig_t = d->eigenvals_t; D.maf_t = d->maf_t;
        D.E = (const double *)e->dE.p; D.zero_one = D.E + n_pad * NCP; D.ormask = (const uint8_t *)e->dMask.p;
        D.FF0 = (const double *)e->dFF0.p; D.R = (const double *)e->dR.p;
        D.Rinv = (const double *)e->dRinv.p; D.tq = (const double *)e->dTq.p;
        D.maf_o = nullptr; D.flip_o = nullptr;
        e->dominant = "contract_packed_kernel";
        if (e->tc_N > 0) {
            char nm[96];
            snprintf(nm, sizeof nm, "contract_tc2_kernel<NT=%d,TILES=%d,HOM=%d>", e->tc_N / 16, tc_tiles(e), e->tc_hom);
            e->dominant_tc = nm;
        }
        CK(cudaStreamSynchronize(st));   // host vectors go out of scope
    } else if (d->model == GT_MODEL_LOGISTIC || d->model == GT_MODEL_COXPH) {
        rc = gtk::iter_set_design(d, pos, n_file, n_chunks, pitch, (const uint8_t *)e->dMask.p,
                                  e->dIter, e->dIterT, e->T, st, g_err);
        if (rc) return rc;
        e->dominant = d->model == GT_MODEL_LOGISTIC ? "logistic_irls_kernel" : "cox_newton_kernel";
    } else {
        return fail(GT_E_ARG, "gt_engine_set_design: unknown model %d", d->model);
    }
    e->have_design = true;
    return 0;
}

}  // extern "C"

// ---------------------------------------------------------------------------
static int time_begin(gt_engine *e, int cls = 0) {
    if (!e->timing) return -1;
    if (e->events_used == e->events.size()) {
        cudaEvent_t a, b;
        if (cudaEventCreate(&a) != cudaSuccess || cudaEventCreate(&b) != cudaSuccess) return -1;
        e->events.emplace_back(a, b);
        e->event_class.push_back(0);
    }
    e->event_class[e->events_used] = cls;
    cudaEventRecord(e->events[e->events_used].first, e->stream);
    return (int)e->events_used;
}
static void time_end(gt_engine *e, int slot) {
    if (slot < 0) return;
    cudaEventRecord(e->events[slot].second, e->stream);
    e->events_used++;
    if (e->event_class[slot] == 0) e->launches++;
}


// cuTensorMapEncodeTiled through the runtime (no link-time dependency on libcuda)
typedef CUresult (*tmap_encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                   const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                                   CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                   CUtensorMapFloatOOBfill);
static tmap_encode_fn tmap_encoder() {
    static tmap_encode_fn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) == cudaSuccess &&
            qr == cudaDriverEntryPointSuccess)
            fn = (tmap_encode_fn)p;
    }
    return fn;
}

// SNP rows per CTA of the tcgen05 contraction for this design
static int tc_tiles(const gt_engine *e) {
    if (e->tc_hom) return e->tc_N <= 224 ? 2 : 1;
    return e->tc_N <= 96 ? 4 : (e->tc_N <= 128 ? 2 : 1);
}
static int tc_rows(const gt_engine *e) { return (e->tc_hom ? 64 : 128) * tc_tiles(e); }

static int launch_contract_tc(gt_engine *e, const uint8_t *g, int64_t pitch, int nb) {
    gtk::Tc2Params P = e->tc;
    P.pitch = pitch; P.n_snps = nb;
    P.S = (double *)e->dS.p; P.counts = (int *)e->dCounts.p; P.gsum = (double *)e->dGsum.p;
    P.mrec = (uint2 *)e->dMlist.p; P.mcount = (int *)e->dMcount.p;
    const int rows = tc_rows(e);
    tmap_encode_fn enc = tmap_encoder();
    if (!enc) return fail(GT_E_CUDA, "tcgen05 contraction: cuTensorMapEncodeTiled is not available");
    // the packed rows as a 2-D byte tensor [n_snps][pitch]; tiles of `rows` x 128 B, 128B swizzle;
    // bytes past a row's pitch / rows past the batch read as zero
    CUtensorMap gmap;
    cuuint64_t gdim[2] = {(cuuint64_t)pitch, (cuuint64_t)nb};
    cuuint64_t gstr[1] = {(cuuint64_t)pitch};
    const cuuint32_t box[2] = {128u, (cuuint32_t)std::min(rows, 256)};
    const cuuint32_t estr[2] = {1u, 1u};
    CUresult cr = enc(&gmap, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, (void *)g, gdim, gstr, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                      CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) return fail(GT_E_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)cr);
    const int grid = (nb + rows - 1) / rows;
    const int NT = e->tc_N / 16;
#define GT_TC(T, TL)                                                                              \
    case T: {                                                                                     \
        const int smem = gtk::Tc2Cfg<TL>::smem_bytes(16 * T);                                     \
        CK(cudaFuncSetAttribute((const void *)gtk::contract_tc2_kernel<T, TL>,                    \
                                cudaFuncAttributeMaxDynamicSharedMemorySize, smem));              \
        gtk::contract_tc2_kernel<T, TL><<<grid, gtk::Tc2Cfg<TL>::kThreads, smem, e->stream>>>(gmap, P); \
    } break;
#define GT_TCD(T, TL, DB)                                                                         \
    if (NT == T && e->tc_dbg == DB) {                                                             \
        const int smem = gtk::Tc2Cfg<TL>::smem_bytes(16 * T);                                     \
        CK(cudaFuncSetAttribute((const void *)gtk::contract_tc2_kernel<T, TL, DB>,                \
                                cudaFuncAttributeMaxDynamicSharedMemorySize, smem));              \
        gtk::contract_tc2_kernel<T, TL, DB><<<grid, gtk::Tc2Cfg<TL>::kThreads, smem, e->stream>>>(gmap, P); \
        CK(cudaGetLastError());                                                                   \
        return 0;                                                                                 \
    }
#define GT_TCH(T, TL)                                                                             \
    case T: {                                                                                     \
        const int smem = gtk::Tc2Cfg<TL, 1>::smem_bytes(16 * T);                                  \
        CK(cudaFuncSetAttribute((const void *)gtk::contract_tc2_kernel<T, TL, 0, 1>,              \
                                cudaFuncAttributeMaxDynamicSharedMemorySize, smem));              \
        gtk::contract_tc2_kernel<T, TL, 0, 1><<<grid, gtk::Tc2Cfg<TL, 1>::kThreads, smem, e->stream>>>(gmap, P); \
    } break;
    if (e->tc_hom) {
        // interaction designs: 64 markers x (dosage, homozygote indicator) per MMA tile
        switch (NT) {
            GT_TCH(2, 2) GT_TCH(3, 2) GT_TCH(4, 2) GT_TCH(5, 2) GT_TCH(6, 2) GT_TCH(7, 2) GT_TCH(8, 2)
            GT_TCH(9, 2) GT_TCH(10, 2) GT_TCH(11, 2) GT_TCH(12, 2) GT_TCH(13, 2) GT_TCH(14, 2)
            GT_TCH(15, 1) GT_TCH(16, 1)
            default: return fail(GT_E_UNSUPPORTED, "tcgen05 contraction (interaction design): N=%d not instantiated", e->tc_N);
        }
        CK(cudaGetLastError());
        return 0;
    }
#undef GT_TCH
    // tuning probes (option "tc_debug") of the 5 x 16 column instantiation only
    GT_TCD(5, 4, 1) GT_TCD(5, 4, 2) GT_TCD(5, 4, 6) GT_TCD(5, 4, 14) GT_TCD(5, 4, 61) GT_TCD(5, 4, 63) GT_TCD(5, 4, 128)
#undef GT_TCD
    switch (NT) {
        GT_TC(1, 4) GT_TC(2, 4) GT_TC(3, 4) GT_TC(4, 4) GT_TC(5, 4) GT_TC(6, 4) GT_TC(7, 2) GT_TC(8, 2)
        GT_TC(9, 1) GT_TC(10, 1) GT_TC(11, 1) GT_TC(12, 1) GT_TC(13, 1) GT_TC(14, 1) GT_TC(15, 1) GT_TC(16, 1)
        default: return fail(GT_E_UNSUPPORTED, "tcgen05 contraction: N=%d not instantiated", e->tc_N);
    }
#undef GT_TC
    CK(cudaGetLastError());
    return 0;
}

// K3a (missing-sample corrections, one warp per SNP) then K3b (solve, 16 or
// 32 lanes per SNP) for one batch.
template <bool PACKED>
static int launch_finish(gt_engine *e, const void *g, int64_t pitch, int nb, double *out,
                         int32_t *iout, long long stride, bool use_list, int64_t s0) {
    GtDesignDev D = e->D;
    if (e->sex_maf && !D.raw_mode) {
        D.maf_o = (const double *)e->dMafO.p + s0;
        D.flip_o = (const int *)e->dFlipO.p + s0;
    }
    const double *S = (const double *)e->dS.p;
    int *cnt = (int *)e->dCounts.p;
    const double *gs = (const double *)e->dGsum.p;
    const uint2 *ml = use_list ? (const uint2 *)e->dMlist.p : nullptr;
    const int *mc = (const int *)e->dMcount.p;
    double *corr = (double *)e->dCorr.p;
    const long long ldc = (nb + 31) / 32 * 32;
    const int NT2 = (D.L + 7) / 8;
    const bool stream = use_list && e->cs_nv > 0 && e->use_stream_corr;
    if (stream) {
        // thread-per-SNP streaming corrections (gt_corr_stream.cuh)
        gtk::CorrStreamParams C = e->cs;
        C.n_snps = nb; C.mrec = ml; C.mcount = mc; C.corrT = corr; C.ldc = ldc; C.counts = cnt;
        const int fold = e->thread_fold;
        const int csn = fold ? gtk::kCsFoldSnps : gtk::kCsSnps;
        const int sm = gtk::kCsBuf * gtk::kCsBlk * gtk::kCsRow * 8 + (gtk::kCsRecRing / 2) * csn * 16 + 256;
        const int cgrid = (nb + csn - 1) / csn;
        int cslot = time_begin(e, 1);
#define GT_CS(N)                                                                              \
    case N:                                                                                   \
        if (fold) {                                                                           \
            CK(cudaFuncSetAttribute((const void *)gtk::missing_corr_stream_kernel<N, 1>,      \
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, sm));        \
            gtk::missing_corr_stream_kernel<N, 1><<<cgrid, csn, sm, e->stream>>>(C);          \
        } else {                                                                              \
            CK(cudaFuncSetAttribute((const void *)gtk::missing_corr_stream_kernel<N, 0>,      \
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, sm));        \
            gtk::missing_corr_stream_kernel<N, 0><<<cgrid, csn + 32, sm, e->stream>>>(C);     \
        }                                                                                     \
        break;
        switch (e->cs_nv) {
            GT_CS(1) GT_CS(2) GT_CS(3) GT_CS(4) GT_CS(5) GT_CS(6) GT_CS(7) GT_CS(8) GT_CS(9) GT_CS(10)
            GT_CS(11) GT_CS(12)
            default: return fail(GT_E_UNSUPPORTED, "streaming corrections: %d columns", e->cs_nv);
        }
#undef GT_CS
        time_end(e, cslot);
        CK(cudaGetLastError());
    } else {
        const gtk::CorrLayout cl = gtk::corr_layout(D.L);
        const int cw = 8;
        const size_t csm = (size_t)cw * cl.perWarp * 8;
        const int cgrid = (nb + cw - 1) / cw;
#define GT_COR(N)                                                                          \
    case N:                                                                                \
        CK(cudaFuncSetAttribute((const void *)gtk::missing_corr_kernel<PACKED, N>,         \
                                cudaFuncAttributeMaxDynamicSharedMemorySize, (int)csm));   \
        gtk::missing_corr_kernel<PACKED, N><<<cgrid, cw * 32, csm, e->stream>>>(           \
            D, g, pitch, nb, cnt, ml, e->tc.rcap, mc, corr, ldc);                              \
        break;
        int cslot = time_begin(e, 1);
        switch (NT2) {
            GT_COR(1) GT_COR(2) GT_COR(3) GT_COR(4) GT_COR(5)
            default: return fail(GT_E_UNSUPPORTED, "correction kernel: design too wide");
        }
        time_end(e, cslot);
#undef GT_COR
        CK(cudaGetLastError());
    }
    {
        const gtk::SolveLayout sl = gtk::solve_layout(D.L, D.p, D.I);
        const bool half = D.L <= 16;                     // two SNPs per warp
        int gpc = half ? 16 : 8;                         // SNPs per 256-thread CTA
        while (gpc > 1 && (size_t)gpc * sl.perGroup * 8 > 110 * 1024) gpc >>= 1;
        const int threads = gpc * (half ? 16 : 32);
        const size_t smem = (size_t)gpc * sl.perGroup * 8;
        if (smem > 220 * 1024) return fail(GT_E_UNSUPPORTED, "design too wide for the solve kernel (%zu B shared)", smem);
        const int grid = (nb + gpc - 1) / gpc;
        int fslot = time_begin(e, 2);
        // designs without interaction and up to 13 parameters: one thread per marker first,
        // the 16-lane kernel then only finishes what that one deferred (Jacobi gate)
        int deferred_only = 0;
        if (D.lean && D.I == 0 && D.p >= 2 && D.p <= 13 && e->use_thread_solve) {
            const int tgrid = (nb + 127) / 128;
#define GT_ST(PP)                                                                                   \
    case PP:                                                                                        \
        gtk::linear_solve_thread_kernel<PP><<<tgrid, 128, 0, e->stream>>>(D, nb, S, cnt, gs, corr,  \
                                                                          ldc, out, iout, stride); \
        break;
            switch (D.p) {
                GT_ST(2) GT_ST(3) GT_ST(4) GT_ST(5) GT_ST(6) GT_ST(7) GT_ST(8) GT_ST(9) GT_ST(10)
                GT_ST(11) GT_ST(12) GT_ST(13)
            }
#undef GT_ST
            CK(cudaGetLastError());
            deferred_only = 1;
            e->all_launches += 1;
        } else if (!D.lean && D.I == 1 && D.p >= 3 && D.p <= 13 && e->use_thread_solve) {
            // one interaction column: the same kernel with the rows g and g w
            const int tgrid = (nb + 127) / 128;
#define GT_ST1(PP)                                                                                  \
    case PP:                                                                                        \
        gtk::linear_solve_thread_kernel<PP, 1><<<tgrid, 128, 0, e->stream>>>(D, nb, S, cnt, gs, corr, \
                                                                             ldc, out, iout, stride); \
        break;
            switch (D.p) {
                GT_ST1(3) GT_ST1(4) GT_ST1(5) GT_ST1(6) GT_ST1(7) GT_ST1(8) GT_ST1(9) GT_ST1(10)
                GT_ST1(11) GT_ST1(12) GT_ST1(13)
            }
#undef GT_ST1
            CK(cudaGetLastError());
            deferred_only = 1;
            e->all_launches += 1;
        }
        if (half) {
            CK(cudaFuncSetAttribute((const void *)gtk::linear_solve_kernel<16>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            gtk::linear_solve_kernel<16><<<grid, threads, smem, e->stream>>>(D, nb, S, cnt, gs, corr, ldc, out, iout, stride, deferred_only);
        } else {
            CK(cudaFuncSetAttribute((const void *)gtk::linear_solve_kernel<32>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            gtk::linear_solve_kernel<32><<<grid, threads, smem, e->stream>>>(D, nb, S, cnt, gs, corr, ldc, out, iout, stride, deferred_only);
        }
        time_end(e, fslot);
        CK(cudaGetLastError());
    }
    return 0;
}

static int run_linear(gt_engine *e, const void *geno, int64_t pitch, bool packed, int32_t n_snps,
                      double *out, int32_t *iout, long long ldo) {
    const GtDesignDev &D = e->D;
    const ContractEntry &ce = kContract[e->contract_idx];
    const int tile = 32 * ce.rows;
    const bool tc = packed && e->use_tc && e->tc_N > 0;
    // whole waves of the tcgen05 kernel (148 SMs x one CTA of 128 / 256 SNP rows); the
    // records of the missing genotypes take up to a packed row per SNP
    int kBatch = 148 * 2 * 128 * 2;
    if (tc) {
        const int wave = 148 * tc_rows(e);
        const long long fit = ((long long)4 << 30) / ((long long)e->tc.rcap * 8);
        kBatch = (int)std::max<long long>(wave, std::min<long long>(8, fit / wave) * wave);
    }
    int rc;
    const int bmax = std::min<int>(n_snps, kBatch);
    if ((rc = e->dS.reserve((size_t)bmax * D.NCP * 8))) return rc;
    if ((rc = e->dCounts.reserve((size_t)bmax * 4 * 4))) return rc;
    if ((rc = e->dGsum.reserve((size_t)bmax * 2 * 8))) return rc;
    if (tc) {
        const size_t brows = ((size_t)bmax + tc_rows(e) - 1) / tc_rows(e) * tc_rows(e);   // whole CTAs
        if ((rc = e->dMlist.reserve(brows * e->tc.rcap * 8 + 256))) return rc;     // + read-ahead slack
        if ((rc = e->dMcount.reserve((size_t)bmax * 4))) return rc;
    }
    if ((rc = e->dCorr.reserve(((size_t)bmax + 32) * (D.L * (D.L + 1) / 2) * 8))) return rc;
    CK(cudaFuncSetAttribute((const void *)ce.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, ce.smem));

    for (int s0 = 0; s0 < n_snps; s0 += kBatch) {
        int nb = std::min<int>(kBatch, n_snps - s0);
        double *S = (double *)e->dS.p;
        int *cnt = (int *)e->dCounts.p;
        double *gs = (double *)e->dGsum.p;
        if (packed) {
            const uint8_t *g = (const uint8_t *)geno + (int64_t)s0 * pitch;
            int slot = time_begin(e);
            if (tc) {
                if ((rc = launch_contract_tc(e, g, pitch, nb))) return rc;
            } else {
                ce.fn<<<(nb + tile - 1) / tile, 32 * ce.warps, ce.smem, e->stream>>>(
                    g, pitch, nb, D.ormask, D.E, D.n_chunks, S, cnt, gs);
            }
            time_end(e, slot);
            CK(cudaGetLastError());
            if ((rc = launch_finish<true>(e, g, pitch, nb, out + s0, iout + s0, ldo, tc, s0))) return rc;
        } else {
            const double *g = (const double *)geno + (int64_t)s0 * pitch;
            gtk::contract_dosage_kernel<<<(nb + 7) / 8, 256, 0, e->stream>>>(
                g, pitch, nb, D.n, D.E, D.NCA, D.NCB, D.NC1, D.NC2, S, cnt, gs);
            CK(cudaGetLastError());
            if ((rc = launch_finish<false>(e, g, pitch, nb, out + s0, iout + s0, ldo, false, s0))) return rc;
        }
        e->all_launches += 3;
    }
    return 0;
}

// out[field][ldo], iout[field][ldo]: ldo >= n_snps (0 selects n_snps) lets a
// caller fill a column range of larger result blocks
static int run_any(gt_engine *e, const void *geno, int64_t pitch, bool packed, int32_t n_snps,
                   double *out, int32_t *iout, long long ldo = 0) {
    if (ldo <= 0) ldo = n_snps;
    if (!e || !e->have_design) return fail(GT_E_STATE, "no design set");
    if (n_snps < 0 || !out || !iout || (n_snps > 0 && !geno)) return fail(GT_E_ARG, "run: bad arguments");
    if (n_snps == 0) return 0;
    CK(cudaSetDevice(e->device));
    if (packed) {
        if (pitch % 16 != 0 || pitch < ((int64_t)e_n_file(e) + 3) / 4)
            return fail(GT_E_ARG, "run: pitch %lld must be a multiple of 16 and >= ceil(n_file/4)", (long long)pitch);
    }
    if (!packed && !e->identity_pos)
        return fail(GT_E_STATE, "dosage rows are aligned to the design rows: set the design without sample_pos / n_file");
    gtk::IterDesignDev T = e->T;
    T.newton_maxiter = e->cox_maxiter;
    if (e->sex_maf) {
        int rc;
        if ((rc = e->dMafO.reserve((size_t)n_snps * 8))) return rc;
        if ((rc = e->dFlipO.reserve((size_t)n_snps * 4))) return rc;
        if (packed)
            gtk::sex_maf_packed_kernel<<<(n_snps + 7) / 8, 256, 0, e->stream>>>(
                (const uint8_t *)geno, pitch, n_snps, e_n_file(e), (const uint8_t *)e->dMask.p,
                (const uint8_t *)e->dSexM.p, (const uint8_t *)e->dSexF.p, (double *)e->dMafO.p,
                (int *)e->dFlipO.p);
        else
            gtk::sex_maf_dosage_kernel<<<(n_snps + 7) / 8, 256, 0, e->stream>>>(
                (const double *)geno, pitch, n_snps, (int)e->pos.size(), (const double *)e->dSexRow.p,
                (double *)e->dMafO.p, (int *)e->dFlipO.p);
        CK(cudaGetLastError());
        e->all_launches += 1;
        T.maf_o = (const double *)e->dMafO.p;
        T.flip_o = (const int *)e->dFlipO.p;
    }
    if (e->model == GT_MODEL_LINEAR) return run_linear(e, geno, pitch, packed, n_snps, out, iout, ldo);
    {
        int rc = e->dCounter.reserve(16);
        if (rc) return rc;
        // coef / se of the last solve of every marker (thread-per-SNP logistic kernel)
        if (e->model == GT_MODEL_LOGISTIC && packed && e->use_thread_iter &&
            (rc = e->dIterScratch.reserve((size_t)(n_snps + 32) * 2 * 13 * 8))) return rc;
        // the codes of every marker in schedule order (thread-per-SNP Cox kernel)
        if (e->model == GT_MODEL_COXPH && packed && e->use_thread_iter && e->T.th_ok &&
            n_snps >= e->thread_iter_min &&
            (rc = e->dIterScratch.reserve((size_t)(n_snps + 32) * (size_t)(e->T.th_n / 16) * 4))) return rc;
    }
    return gtk::iter_run(T, e->model, geno, pitch, packed, n_snps, out, iout, ldo, e->stream,
                         (int *)e->dCounter.p, e->timing,
                         (double *)e->dIterScratch.p, e->use_thread_iter, e->thread_iter_min, e->thread_fold,
                         [&]() { return time_begin(e); }, [&](int s) { time_end(e, s); },
                         e->all_launches, g_err);
}


extern "C" {

int gt_run_packed_dev(gt_engine *e, const uint8_t *packed_dev, int64_t pitch, int32_t n_snps,
                      double *out_dev, int32_t *iout_dev) {
    return run_any(e, packed_dev, pitch, true, n_snps, out_dev, iout_dev);
}

int gt_run_dosage_dev(gt_engine *e, const double *dosage_dev, int64_t ld, int32_t n_snps,
                      double *out_dev, int32_t *iout_dev) {
    if (e && e->have_design && ld < (e->model == GT_MODEL_LINEAR ? e->D.n : e->T.n))
        return fail(GT_E_ARG, "gt_run_dosage: ld (%lld) < n", (long long)ld);
    return run_any(e, dosage_dev, ld, false, n_snps, out_dev, iout_dev);
}

// Host-buffer entry points: a two-slot pipeline.  While chunk c computes on
// the engine's stream, chunk c+1 crosses PCIe on s_in and the results of chunk
// c-1 return on s_out; the order of the calls below also overlaps when the
// host buffers are pageable (those copies block the calling thread).
static int run_host(gt_engine *e, const void *src, int64_t src_pitch_bytes, int64_t dev_pitch_bytes,
                    int64_t width_bytes, int64_t dev_pitch_arg, bool packed, int32_t n_snps,
                    double *out_host, int32_t *iout_host) {
    if (!e || !e->have_design) return fail(GT_E_STATE, "no design set");
    if (n_snps <= 0) return n_snps == 0 ? 0 : fail(GT_E_ARG, "n_snps < 0");
    if (!src || !out_host || !iout_host) return fail(GT_E_ARG, "run_host: NULL buffer");
    CK(cudaSetDevice(e->device));
    const int nf = gt_engine_num_fields(e);
    if (!e->s_in) {
        CK(cudaStreamCreateWithFlags(&e->s_in, cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&e->s_out, cudaStreamNonBlocking));
        for (int i = 0; i < 2; ++i) {
            CK(cudaEventCreateWithFlags(&e->ev_in[i], cudaEventDisableTiming));
            CK(cudaEventCreateWithFlags(&e->ev_done[i], cudaEventDisableTiming));
        }
        CK(cudaEventCreateWithFlags(&e->ev_start, cudaEventDisableTiming));
    }
    // stage size: about an eighth of the call, at least 8192 markers (every
    // SM busy) and at most 512 MiB of genotypes per slot
    int64_t chunk = e->host_chunk > 0 ? e->host_chunk : (((int64_t)n_snps + 7) / 8 + 1023) / 1024 * 1024;
    if (e->host_chunk <= 0) {
        // the iterative models are compute-bound by orders of magnitude and
        // pay a partial last wave per chunk: keep their chunks large
        chunk = std::max<int64_t>(chunk, e->model == GT_MODEL_LINEAR ? 8192 : 65536);
        // thread-per-SNP IRLS: whole waves of 148 CTAs x 224 markers
        const int64_t tw = 148 * (int64_t)(e->thread_fold ? gtk::kLtFoldThreads : gtk::kLtMaxThreads);
        if (packed && e->use_thread_iter && chunk >= tw &&
            ((e->model == GT_MODEL_LOGISTIC && e->T.I <= 1 && e->T.p <= 13 && e->T.y_binary) ||
             (e->model == GT_MODEL_COXPH && e->T.th_ok)))
            chunk = chunk / tw * tw;
        chunk = std::min<int64_t>(chunk, std::max<int64_t>(1024, ((int64_t)512 << 20) / dev_pitch_bytes / 1024 * 1024));
        // whole half-waves of the tcgen05 contraction (148 CTAs x 128 .. 512 SNP rows)
        if (packed && e->model == GT_MODEL_LINEAR && e->use_tc && e->tc_N > 0) {
            const int64_t hw = 148 * (int64_t)tc_rows(e) / 2;
            if (chunk > hw) chunk = chunk / hw * hw;
        }
    }
    chunk = std::min<int64_t>(chunk, n_snps);
    int rc;
    if ((rc = e->dGeno.reserve((size_t)dev_pitch_bytes * chunk * 2))) return rc;
    if ((rc = e->dOut.reserve((size_t)nf * n_snps * 8))) return rc;
    if ((rc = e->dIout.reserve((size_t)GT_NUM_IFIELDS * n_snps * 4))) return rc;
    const bool want_maf = packed && e->file_maf_host != nullptr;
    if (want_maf && (rc = e->dFileMaf.reserve((size_t)n_snps * 8))) return rc;
    // A pageable source (an mmap-ed .bed, a numpy array) is staged through two pinned
    // buffers by a few host threads: cudaMemcpyAsync from pageable memory is a
    // synchronous, single-threaded bounce copy at a fraction of the link rate.
    bool stage = false;
    {
        cudaPointerAttributes attr;
        cudaError_t pe = cudaPointerGetAttributes(&attr, src);
        if (pe != cudaSuccess) { cudaGetLastError(); stage = true; }
        else stage = attr.type == cudaMemoryTypeUnregistered;
    }
    const size_t slot_bytes = (size_t)width_bytes * chunk;
    if (stage) {
        if (e->pin_bytes < slot_bytes) {
            for (int i = 0; i < 2; ++i) { if (e->pin[i]) cudaFreeHost(e->pin[i]); e->pin[i] = nullptr; }
            e->pin_bytes = 0;
            for (int i = 0; i < 2; ++i)
                if (cudaHostAlloc(&e->pin[i], slot_bytes, cudaHostAllocDefault) != cudaSuccess) {
                    cudaGetLastError();
                    for (int j = 0; j < 2; ++j) { if (e->pin[j]) cudaFreeHost(e->pin[j]); e->pin[j] = nullptr; }
                    stage = false;                       // no pinned memory: fall back to the direct copy
                    break;
                }
            if (stage) e->pin_bytes = slot_bytes;
        }
        for (int i = 0; i < 2 && stage; ++i)
            if (!e->ev_h2d[i]) CK(cudaEventCreateWithFlags(&e->ev_h2d[i], cudaEventDisableTiming));
    }
    const int n_threads = (int)std::min<unsigned>(std::max(1u, std::thread::hardware_concurrency()), 16u);
    auto stage_rows = [&](int64_t c0, int nb, char *dst) {
        // rows [c0, c0 + nb) -> dst, `width_bytes` each, tightly packed
        auto work = [&](int t) {
            const int64_t r0 = (int64_t)nb * t / n_threads, r1 = (int64_t)nb * (t + 1) / n_threads;
            if (src_pitch_bytes == width_bytes) {
                std::memcpy(dst + r0 * width_bytes, (const char *)src + (c0 + r0) * src_pitch_bytes,
                            (size_t)(r1 - r0) * width_bytes);
            } else {
                for (int64_t r = r0; r < r1; ++r)
                    std::memcpy(dst + r * width_bytes, (const char *)src + (c0 + r) * src_pitch_bytes,
                                (size_t)width_bytes);
            }
        };
        std::vector<std::thread> pool;
        for (int t = 1; t < n_threads; ++t) pool.emplace_back(work, t);
        work(0);
        for (auto &th : pool) th.join();
    };
    double *dout = (double *)e->dOut.p;
    int32_t *diout = (int32_t *)e->dIout.p;
    CK(cudaEventRecord(e->ev_start, e->stream));       // earlier work on the engine's stream
    CK(cudaStreamWaitEvent(e->s_in, e->ev_start, 0));
    auto copy_back = [&](int64_t c0, int nb, int slot) -> int {
        CK(cudaStreamWaitEvent(e->s_out, e->ev_done[slot], 0));
        CK(cudaMemcpy2DAsync(out_host + c0, (size_t)n_snps * 8, dout + c0, (size_t)n_snps * 8,
                             (size_t)nb * 8, nf, cudaMemcpyDeviceToHost, e->s_out));
        CK(cudaMemcpy2DAsync(iout_host + c0, (size_t)n_snps * 4, diout + c0, (size_t)n_snps * 4,
                             (size_t)nb * 4, GT_NUM_IFIELDS, cudaMemcpyDeviceToHost, e->s_out));
        if (want_maf)
            CK(cudaMemcpyAsync(e->file_maf_host + c0, (double *)e->dFileMaf.p + c0, (size_t)nb * 8,
                               cudaMemcpyDeviceToHost, e->s_out));
        return 0;
    };
    int64_t prev0 = -1;
    int prev_nb = 0, prev_slot = 0, c = 0;
    for (int64_t c0 = 0; c0 < n_snps; c0 += chunk, ++c) {
        const int nb = (int)std::min<int64_t>(chunk, n_snps - c0), slot = c & 1;
        char *dg = (char *)e->dGeno.p + (size_t)slot * dev_pitch_bytes * chunk;
        const char *hsrc = (const char *)src + c0 * src_pitch_bytes;
        size_t hpitch = (size_t)src_pitch_bytes;
        if (stage) {
            if (c >= 2) CK(cudaEventSynchronize(e->ev_h2d[slot]));    // the copy out of this buffer is done
            stage_rows(c0, nb, (char *)e->pin[slot]);
            hsrc = (const char *)e->pin[slot];
            hpitch = (size_t)width_bytes;
        }
        if (c >= 2) CK(cudaStreamWaitEvent(e->s_in, e->ev_done[slot], 0));   // slot read by chunk c-2
        CK(cudaMemcpy2DAsync(dg, dev_pitch_bytes, hsrc, hpitch,
                             width_bytes, nb, cudaMemcpyHostToDevice, e->s_in));
        if (stage) CK(cudaEventRecord(e->ev_h2d[slot], e->s_in));
        CK(cudaEventRecord(e->ev_in[slot], e->s_in));
        CK(cudaStreamWaitEvent(e->stream, e->ev_in[slot], 0));
        if (want_maf) {
            file_maf_packed_kernel<<<(nb + 7) / 8, 256, 0, e->stream>>>(
                (const uint8_t *)dg, dev_pitch_bytes, nb, e_n_file(e), (double *)e->dFileMaf.p + c0);
            CK(cudaGetLastError());
            e->all_launches += 1;
        }
        if ((rc = run_any(e, dg, dev_pitch_arg, packed, nb, dout + c0, diout + c0, n_snps))) return rc;
        CK(cudaEventRecord(e->ev_done[slot], e->stream));
        if (prev0 >= 0 && (rc = copy_back(prev0, prev_nb, prev_slot))) return rc;
        prev0 = c0; prev_nb = nb; prev_slot = slot;
    }
    if ((rc = copy_back(prev0, prev_nb, prev_slot))) return rc;
    CK(cudaStreamSynchronize(e->s_out));
    return gt_engine_synchronize(e);
}

int gt_run_packed_host(gt_engine *e, const uint8_t *packed_host, int64_t row_bytes, int32_t n_snps,
                       double *out_host, int32_t *iout_host) {
    if (!e || !e->have_design) return fail(GT_E_STATE, "no design set");
    int64_t bps = ((int64_t)e_n_file(e) + 3) / 4;
    if (row_bytes < bps) return fail(GT_E_ARG, "gt_run_packed_host: row_bytes %lld < ceil(n_file/4)=%lld", (long long)row_bytes, (long long)bps);
    int64_t pitch = gt_packed_pitch(e_n_file(e));
    return run_host(e, packed_host, row_bytes, pitch, bps, pitch, true, n_snps, out_host, iout_host);
}

int gt_run_dosage_host(gt_engine *e, const double *dosage_host, int64_t ld, int32_t n_snps,
                       double *out_host, int32_t *iout_host) {
    if (!e || !e->have_design) return fail(GT_E_STATE, "no design set");
    int n = e->model == GT_MODEL_LINEAR ? e->D.n : e->T.n;
    if (ld < n) return fail(GT_E_ARG, "gt_run_dosage_host: ld (%lld) < n (%d)", (long long)ld, n);
    return run_host(e, dosage_host, ld * 8, (int64_t)n * 8, (int64_t)n * 8, n, false, n_snps, out_host, iout_host);
}

int gt_synth_packed_dev(gt_engine *e, uint8_t *packed_dev, int64_t pitch, int32_t n_file,
                        int32_t n_snps, uint64_t seed, double maf_lo, double maf_hi,
                        double missing_rate) {
    if (!e || !packed_dev) return fail(GT_E_ARG, "gt_synth_packed_dev: NULL argument");
    if (pitch < ((int64_t)n_file + 3) / 4) return fail(GT_E_ARG, "gt_synth_packed_dev: pitch too small");
    CK(cudaSetDevice(e->device));
    CK(cudaMemsetAsync(packed_dev, 0xFF, (size_t)pitch * n_snps, e->stream));
    synth_packed_kernel<<<148 * 16, 256, 0, e->stream>>>(packed_dev, pitch, n_file, n_snps, seed,
                                                         maf_lo, maf_hi, missing_rate);
    CK(cudaGetLastError());
    return 0;
}

int gt_measure_fp64_peak(gt_engine *e, double *tflops) {
    if (!e || !tflops) return fail(GT_E_ARG, "gt_measure_fp64_peak: NULL argument");
    CK(cudaSetDevice(e->device));
    int rc = e->dS.reserve(64);
    if (rc) return rc;
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    const int blocks = 148 * 8, threads = 256, iters = 1 << 15;
    double best = 0.0;
    for (int rep = 0; rep < 6; ++rep) {
        CK(cudaEventRecord(a, e->stream));
        fp64_peak_kernel<<<blocks, threads, 0, e->stream>>>((double *)e->dS.p, iters, 0.999999, 1e-9);
        CK(cudaEventRecord(b, e->stream));
        CK(cudaEventSynchronize(b));
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, a, b));
        double tf = 2.0 * blocks * threads * 8.0 * iters / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
    }
    cudaEventDestroy(a); cudaEventDestroy(b);
    *tflops = best;
    return 0;
}

int gt_measure_dmma_peak(gt_engine *e, double *tflops) {
    if (!e || !tflops) return fail(GT_E_ARG, "gt_measure_dmma_peak: NULL argument");
    CK(cudaSetDevice(e->device));
    int rc = e->dS.reserve(64);
    if (rc) return rc;
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    const int blocks = 148 * 8, threads = 256, iters = 1 << 12;
    double best = 0.0;
    for (int rep = 0; rep < 6; ++rep) {
        CK(cudaEventRecord(a, e->stream));
        dmma_peak_kernel<<<blocks, threads, 0, e->stream>>>((double *)e->dS.p, iters, 0.5, 0.25);
        CK(cudaEventRecord(b, e->stream));
        CK(cudaEventSynchronize(b));
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, a, b));
        // one m8n8k4 MMA = 256 FMA = 512 FLOP per warp
        double tf = 512.0 * blocks * (threads / 32) * 8.0 * iters / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
    }
    cudaEventDestroy(a); cudaEventDestroy(b);
    *tflops = best;
    return 0;
}

}  // extern "C"
