// Host-side helpers shared by the API translation unit: error reporting,
// device buffers, covariate factorisation.
#pragma once

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "gt_common.cuh"

static thread_local std::string g_err;

static int fail(int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CK(call)                                                                        \
    do {                                                                                \
        cudaError_t _e = (call);                                                        \
        if (_e != cudaSuccess)                                                          \
            return fail(GT_E_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), \
                        __FILE__, __LINE__);                                            \
    } while (0)

// ---------------------------------------------------------------------------
struct DevBuf {
    void *p = nullptr;
    size_t bytes = 0;
    int reserve(size_t need) {
        if (need <= bytes) return 0;
        if (p) cudaFree(p);
        p = nullptr; bytes = 0;
        cudaError_t e = cudaMalloc(&p, need);
        if (e != cudaSuccess) return fail(GT_E_NOMEM, "cudaMalloc(%zu) failed: %s", need, cudaGetErrorString(e));
        bytes = need;
        return 0;
    }
    void release() { if (p) cudaFree(p); p = nullptr; bytes = 0; }
};

// ---------------------------------------------------------------------------
// host-side design factorisation
// ---------------------------------------------------------------------------
struct HostDesign {
    int n, k, I, L;
    std::vector<double> Q;      // n*k column-major, orthonormal
    std::vector<double> R, Rinv;  // k*k row-major upper
    std::vector<double> yc;     // centred outcome
    bool degenerate = false;
    int has_const = 0, icpt_col = -1;
    double icpt_shift = 0.0;
};

// Modified Gram-Schmidt with re-orthogonalisation, long double accumulation.
static void qr_mgs2(const double *C, int n, int k, HostDesign &H) {
    H.Q.assign(C, C + (size_t)n * k);
    H.R.assign((size_t)k * k, 0.0);
    H.Rinv.assign((size_t)k * k, 0.0);
    for (int j = 0; j < k; ++j) {
        double *qj = &H.Q[(size_t)j * n];
        long double norm0 = 0;
        for (int i = 0; i < n; ++i) norm0 += (long double)qj[i] * qj[i];
        for (int pass = 0; pass < 2; ++pass) {
            for (int a = 0; a < j; ++a) {
                const double *qa = &H.Q[(size_t)a * n];
                long double dot = 0;
                for (int i = 0; i < n; ++i) dot += (long double)qa[i] * qj[i];
                double d = (double)dot;
                for (int i = 0; i < n; ++i) qj[i] -= d * qa[i];
                H.R[(size_t)a * k + j] += d;
            }
        }
        long double nrm = 0;
        for (int i = 0; i < n; ++i) nrm += (long double)qj[i] * qj[i];
        double r = (double)sqrtl(nrm);
        if (!(nrm > 1e-24L * norm0) || !(r > 0.0)) {
            H.degenerate = true;
            r = 1.0;
            for (int i = 0; i < n; ++i) qj[i] = 0.0;
        } else {
            for (int i = 0; i < n; ++i) qj[i] /= r;
        }
        H.R[(size_t)j * k + j] = r;
    }
    // Rinv by back substitution (upper triangular)
    for (int c = 0; c < k; ++c) {
        for (int i = c; i >= 0; --i) {
            long double s = (i == c) ? 1.0L : 0.0L;
            for (int a = i + 1; a <= c; ++a) s -= (long double)H.R[(size_t)i * k + a] * H.Rinv[(size_t)a * k + c];
            H.Rinv[(size_t)i * k + c] = (double)(s / H.R[(size_t)i * k + i]);
        }
    }
}

static int upload(DevBuf &b, const void *src, size_t bytes, cudaStream_t st) {
    int rc = b.reserve(bytes ? bytes : 16);
    if (rc) return rc;
    if (bytes) CK(cudaMemcpyAsync(b.p, src, bytes, cudaMemcpyHostToDevice, st));
    return 0;
}

