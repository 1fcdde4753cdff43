/*
 * genetest_b200 -- C ABI of the B200-native per-SNP association engine.
 *
 * This is the drop-in boundary for ONE path of pgxcentre/genetest: the
 * per-SNP loop  analysis.execute -> _execute_gwas -> _gwas_worker ->
 * Stats{Linear,Logistic,CoxPH}.fit  (reference: genetest/analysis.py:593-728,
 * :68-205; statistics/models/linear.py:93-104, logistic.py:29-74,
 * survival.py:60-122).  The reference has no FFI of its own (pure Python over
 * statsmodels); the entry points below are what a binding for that path
 * needs: "set the design once" (what execute() prepares before the loop,
 * analysis.py:265-286), "run a batch of SNPs" (the body of _gwas_worker for
 * many markers at once) and fixed-size result blocks (the dict that
 * Subscriber.handle() receives, one row per SNP).
 *
 * Conventions
 *   - every function returns 0 on success, a negative GT_E_* code on error;
 *     gt_last_error() gives the message (thread-local).  Nothing throws.
 *   - plain pointers and sizes only.  "dev" pointers are CUDA device pointers
 *     on the engine's device; "host" pointers are ordinary memory (pinned
 *     memory makes the copies asynchronous).
 *   - the caller owns every buffer it passes in; the engine owns its
 *     workspace.  One engine per GPU; an engine is not thread-safe.
 */
#ifndef GENETEST_B200_H
#define GENETEST_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GT_ABI_VERSION 2

/* models (genetest/statistics/__init__.py:35-40 model_map keys) */
enum { GT_MODEL_LINEAR = 0, GT_MODEL_LOGISTIC = 1, GT_MODEL_COXPH = 2 };

/* per-SNP status; the host formats the reference's failure strings
 * (analysis.py:123,152; linear.py:51,59,80; logistic.py:43; survival.py:84-88) */
enum {
    GT_OK = 0,
    GT_ALL_MISSING = 1,         /* "All genotypes are missing"               */
    GT_MAF_BELOW = 2,           /* "MAF: {maf} < {maf_t}"                    */
    GT_COND_LARGE = 3,          /* "condition number is large, {aux}"        */
    GT_EIG_SMALL = 4,           /* "smallest eigenvalue is small, {aux}"     */
    GT_NAN_PVALUE = 5,          /* "Inference did not converge."             */
    GT_PERFECT_SEPARATION = 6,  /* "Perfect separation detected, ..."        */
    GT_SINGULAR = 7,            /* "Singular matrix"                         */
    GT_NO_CONVERGENCE = 8       /* "Maximum Likelihood optimization failed to converge." */
};

/* error codes */
enum {
    GT_E_ARG = -1, GT_E_CUDA = -2, GT_E_STATE = -3, GT_E_UNSUPPORTED = -4,
    GT_E_NOMEM = -5
};

/* result block layout: doubles out[gt_engine_num_fields()][n_snps]
 * (field-major, one contiguous row of n_snps doubles per field)          */
enum {
    GT_F_MAF = 0,      /* minor allele frequency on the complete cases      */
    GT_F_AUX = 1,      /* offending scalar of a failure (cond, eigenvalue)  */
    GT_F_STAT = 2,     /* linear: r_squared_adj; logistic/coxph: nevents    */
    GT_F_LLF = 3,      /* log-likelihood                                    */
    GT_F_PARAM0 = 4    /* then 6 doubles per design column, in fit() order:
                          coef, std_err, lower_ci, upper_ci, t(z)_value,
                          p_value                                            */
};
/* integer block: int32 iout[GT_NUM_IFIELDS][n_snps] = status, nobs, flip,
 * iterations (IRLS / Newton steps; linear: 1 when the Jacobi gate ran)   */
enum { GT_I_STATUS = 0, GT_I_NOBS = 1, GT_I_FLIP = 2, GT_I_NITER = 3,
       GT_NUM_IFIELDS = 4 };

typedef struct gt_engine gt_engine;

/* What execute() hands to every worker (analysis.py:619-630): y, X without
 * the SNP column, the gwas_interaction columns, thresholds. */
typedef struct gt_design_desc {
    int32_t model;          /* GT_MODEL_*                                    */
    int32_t n;              /* design rows (samples after dropna)            */
    int32_t k;              /* covariate columns in C (intercept included)   */
    int32_t n_inter;        /* SNP x covariate interaction columns           */
    int32_t n_file;         /* samples per genotype row of the packed source */
    int32_t raw_mode;       /* 1: genotype column used as is (no MAF, flip)  */
    const double *C;        /* host, n*k column-major                        */
    const double *y;        /* host, n: outcome (linear, logistic) or tte    */
    const double *event;    /* host, n: coxph event indicator, else NULL     */
    const double *W;        /* host, n*n_inter column-major multipliers: the
                               product of the covariate columns each
                               interaction multiplies SNPs with
                               (analysis.py:162-166), or NULL                */
    const int32_t *sample_pos; /* host, n: position of each design row in
                               the genotype source's sample order
                               (analysis.py:43-65), NULL = identity          */
    double condition_value_t;  /* StatsLinear(condition_value_t=1000)        */
    double eigenvals_t;        /* StatsLinear(eigenvals_t=1e-10)             */
    double maf_t;              /* execute(maf_t=...)                         */
    int32_t use_maf_t;
    int32_t reserved;
    const double *strata;   /* host, n: coxph stratum label of every design row
                               (y.strata -> PHReg(strata=...), survival.py:54-56,
                               77-82), or NULL                                  */
} gt_design_desc;

int gt_abi_version(void);
const char *gt_last_error(void);

int gt_engine_create(int device, gt_engine **out);
int gt_engine_destroy(gt_engine *e);
/* cudaStream_t to launch on.  NULL = the engine's own non-blocking stream;
 * use cudaStreamLegacy / cudaStreamPerThread for a default stream. */
int gt_engine_set_stream(gt_engine *e, void *cuda_stream);
int gt_engine_synchronize(gt_engine *e);
/* Engine options.  "contract_tc" = 1 (default): run the linear contraction
 * (designs with and without SNP x covariate interactions) on the tcgen05 int8
 * tensor-core kernel (exact fixed point) instead of the FP64 kernel.  "host_chunk" = n: markers per
 * stage of the copy-in / compute / copy-back pipeline of the *_host entry
 * points (0 = automatic).  "stream_corrections" = 1 (default): missing-sample
 * corrections of those designs on the streaming thread-per-SNP kernel instead
 * of the gather kernel.  "cox_maxiter" = n: Newton step limit of the Cox model
 * (PHReg.fit's maxiter, default 100; reaching it is "Maximum Likelihood
 * optimization failed to converge", survival.py:87-88).  "thread_solve" = 1
 * (default): thread-per-SNP solve of narrow linear designs.  "thread_iter" = 1
 * (default): logistic fits of .bed rows with a binary outcome, at most one
 * interaction column and at most 13 parameters, and Cox fits without
 * interaction and at most 12 parameters, run on the thread-per-SNP kernels
 * when the call has at least "thread_iter_min" markers (default 4096), the
 * warp-per-SNP kernels finishing what they defer.  "thread_fold" = 1
 * (default): those kernels and the streaming corrections run without a
 * producer warp (warp 0 issues the bulk copies), 256 markers per CTA.  "tc_debug": tuning probes of the tcgen05
 * kernel. */
int gt_engine_set_option(gt_engine *e, const char *name, int value);

/* Replaces the per-worker setup of _gwas_worker (analysis.py:68-106) and the
 * once-per-analysis part of Stats*.fit. */
int gt_engine_set_design(gt_engine *e, const gt_design_desc *d);
/* sexual_chromosome=True (analysis.py:294-300, 141-148): sex of every design
 * row (host, n doubles: 0 = female, 1 = male, NaN = unknown), after
 * gt_engine_set_design.  MAF and the flip decision of every marker then follow
 * get_sex_maf (statistics/descriptive.py:51-93).  NULL switches it off. */
int gt_engine_set_sample_sex(gt_engine *e, const double *sex);
int gt_engine_num_params(const gt_engine *e);   /* p = k + 1 + n_inter (coxph: without the constant column) */
int gt_engine_num_fields(const gt_engine *e);   /* GT_F_PARAM0 + 6 p        */
/* row pitch (bytes) the engine wants for device-resident packed genotypes */
int64_t gt_packed_pitch(int32_t n_file);

/* The hot path: one batch of SNPs, PLINK .bed SNP-major 2-bit rows
 * (00 hom A1 = 2 copies of the coded allele, 01 missing, 10 het, 11 hom A2).
 * Replaces the body of _gwas_worker for n_snps markers (analysis.py:108-197).
 * Device-resident variant: `packed_dev` rows are `pitch` bytes apart
 * (pitch % 16 == 0, >= ceil(n_file/4)); results stay on the device. Asynchronous. */
int gt_run_packed_dev(gt_engine *e, const uint8_t *packed_dev, int64_t pitch,
                      int32_t n_snps, double *out_dev, int32_t *iout_dev);
/* Host-buffer variant: rows `row_bytes` apart in host memory; copies in,
 * runs, copies the result blocks back and synchronises. */
int gt_run_packed_host(gt_engine *e, const uint8_t *packed_host,
                       int64_t row_bytes, int32_t n_snps, double *out_host,
                       int32_t *iout_host);

/* MAFFilter support (genetest/modelspec/predicates.py:33-57): when `maf_host` is
 * not NULL, every following gt_run_packed_host call also writes the MAF of each
 * marker over ALL samples of its file row -- get_maf(snp.genotypes), what the
 * predicate evaluates before a marker enters the analysis -- to
 * maf_host[0 .. n_snps).  NULL switches it off. */
int gt_engine_set_file_maf_output(gt_engine *e, double *maf_host);

/* Same for real-valued dosages (impute2 / bgen / dataframe readers):
 * row-major [n_snps][ld] doubles already aligned to the design rows
 * (ld >= n), NaN = missing. */
int gt_run_dosage_dev(gt_engine *e, const double *dosage_dev, int64_t ld,
                      int32_t n_snps, double *out_dev, int32_t *iout_dev);
int gt_run_dosage_host(gt_engine *e, const double *dosage_host, int64_t ld,
                       int32_t n_snps, double *out_host, int32_t *iout_host);

/* Device-side generator of synthetic .bed rows (bench / tests only):
 * per-SNP maf ~ U(maf_lo, maf_hi), g ~ Binomial(2, maf), iid missing. */
int gt_synth_packed_dev(gt_engine *e, uint8_t *packed_dev, int64_t pitch,
                        int32_t n_file, int32_t n_snps, uint64_t seed,
                        double maf_lo, double maf_hi, double missing_rate);

/* Kernel timing (CUDA events on the launch stream around the dominant
 * kernel of the current model); read after gt_engine_synchronize(). */
int gt_engine_enable_timing(gt_engine *e, int on);
int gt_engine_kernel_time(gt_engine *e, double *ms_total, int64_t *launches,
                          int64_t *all_launches);
const char *gt_engine_dominant_kernel(const gt_engine *e);
/* ms[0] contraction, ms[1] missing-sample corrections, ms[2] finish / iterative
 * kernels, ms[3] unused; accumulated while timing is enabled. */
int gt_engine_kernel_breakdown(gt_engine *e, double *ms4);
/* FP64 FMA throughput of the device (TFLOP/s), measured with a register-only
 * FMA kernel: the roofline denominator of the FP64-bound kernels. */
int gt_measure_fp64_peak(gt_engine *e, double *tflops);
/* same for the FP64 tensor-core path (mma.sync m8n8k4 f64). */
int gt_measure_dmma_peak(gt_engine *e, double *tflops);
/* tuning probe: bandwidth of the row-streaming access pattern (128 rows per
 * CTA, `tpr` threads per row, `vec` x 16 B per thread and step). */
int gt_debug_read_pattern(gt_engine *e, const uint8_t *packed_dev, int64_t pitch,
                          int32_t n_rows, int tpr, int vec, int ctas_per_sm,
                          double *gbs);

/* The p-value / quantile formulas of the kernels, evaluated on the host (the
 * same source compiled for the CPU): 2 P(T_df > |t|), t.ppf(0.975, df),
 * 2 Phi(-|z|) -- scipy.stats.t.sf / t.ppf / norm.sf of linear.py:64-89,
 * logistic.py:49-72, survival.py:100-120.  For tests; no device needed. */
double gt_host_t_sf2(double t, double df);
double gt_host_t_ppf975(double df);
double gt_host_norm_sf2(double z);
/* exp(-a), a >= 0, and log(1 + t), t in [0, 1], as the thread-per-SNP logistic kernel
 * evaluates them (GLM Binomial mean and deviance, statsmodels genmod/families). */
double gt_host_lt_exp_neg(double a);
double gt_host_lt_log1p01(double t);

/* Streaming result sink (host only): format `n_rows` result rows as delimited
 * text, the way GWASWriter / RowWriter.handle() does one dict at a time
 * (genetest/subscribers.py:192-205).  Column types: 0 = NUL-separated strings
 * (one per row), 1 = double[n_rows] printed like Python's repr(float) (what
 * the reference's str(float) writes), 2 = int64[n_rows], 3 = one constant
 * string.  `keep` (or NULL) selects the rows to emit.  Returns the number of
 * bytes written, or a negative error. */
int64_t gt_format_rows(int32_t n_rows, int32_t n_cols, const int32_t *col_type,
                       const void *const *col_data, const uint8_t *keep, char sep,
                       char *out, int64_t out_cap);

#ifdef __cplusplus
}
#endif
#endif /* GENETEST_B200_H */
