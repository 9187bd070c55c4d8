/*
 * TEST INFRASTRUCTURE — NOT PRODUCT CODE.
 *
 * CPU restatement (plain C, sequential, fp64 + fp32) of the reference's
 * `grouped_cumprod` extension ops.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may call into this file; the
 * shipped path (simplegaussiansplat_tk71_b200/csrc) never links or loads it.
 *
 * Parity pin: checked in tests/test_oracle.py against the only two known-answer
 * vectors the reference holds for this path —
 *   KAT1  /root/reference/cuda_test.py:19-34   (fwd [.4,.08,.1,.08,.2], bwd [.44,.08,.74,.08,.2])
 *   KAT2  /root/reference/uitility.py:383-393  ([1..7] grouped -> [1,2,3,8,5,18,35])
 * and, on the GPU box, against the reference's own CUDA ops built unchanged by
 * oracle/build_ref.sh (oracle/_ref/grouped_cumprod_ref.so).
 *
 * The forward arithmetic of the reference lives in NVIDIA CCCL (Thrust/CUB 2.8.2,
 * CUDA 12.9 toolkit headers; not vendored under /root/reference):
 * thrust::inclusive_scan_by_key(keys, vals, out, equal_to<int>, BinaryOp) is
 * specified as: out[i] = vals[i] when i == 0 or !pred(keys[i-1], keys[i]),
 * otherwise out[i] = op(out[i-1], vals[i])  — i.e. an inclusive scan restarted at
 * every *adjacent-run* boundary (keys need not be globally sorted or unique).
 * CUB's combination order is policy dependent, so fp32 parity is toleranced
 * (|a-b| <= 1e-6 + 1e-5|b|) against the fp64 functions below; integers are exact.
 *
 * Build: make -C oracle   (gcc -O2 -fopenmp -shared -fPIC)
 */
#include <stdint.h>
#include <stddef.h>
#include <stdlib.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define HEAD(key, i) ((i) == 0 || (key)[(i)] != (key)[(i) - 1])

/* grouped_cumprod_forward — /root/reference/cuda_kernel/grouped_cumprod_forward.cu:17-23 */
void gcp_oracle_cumprod_fwd_f64(const float *x, const int32_t *key, double *y, int64_t n) {
    double acc = 1.0;
    for (int64_t i = 0; i < n; ++i) {
        acc = HEAD(key, i) ? (double)x[i] : acc * (double)x[i];
        y[i] = acc;
    }
}

/* same, fp32 sequential (the order a single CPU thread / one CUDA thread would use) */
void gcp_oracle_cumprod_fwd_f32(const float *x, const int32_t *key, float *y, int64_t n) {
    float acc = 1.0f;
    for (int64_t i = 0; i < n; ++i) {
        acc = HEAD(key, i) ? x[i] : acc * x[i];
        y[i] = acc;
    }
}

/* grouped_cumsum_forward — /root/reference/cuda_kernel/grouped_cumsum_forward.cu:17-23 */
void gcp_oracle_cumsum_fwd_f64(const float *x, const int32_t *key, double *y, int64_t n) {
    double acc = 0.0;
    for (int64_t i = 0; i < n; ++i) {
        acc = HEAD(key, i) ? (double)x[i] : acc + (double)x[i];
        y[i] = acc;
    }
}

void gcp_oracle_cumsum_fwd_f32(const float *x, const int32_t *key, float *y, int64_t n) {
    float acc = 0.0f;
    for (int64_t i = 0; i < n; ++i) {
        acc = HEAD(key, i) ? x[i] : acc + x[i];
        y[i] = acc;
    }
}

/*
 * grouped_cumprod_backward, the reference's formula, literally —
 * /root/reference/cuda_kernel/grouped_cumprod_backward.cu:18-29:
 *   gid = inv[idx]; i_max = inv_len[gid];
 *   p = (param[idx] != 0) ? param[idx] : 1e-8f;
 *   val = sum_{i=idx}^{i_max-1} grad_out[i] * (param_cumprod[i] / p)
 * fp32, same loop order, O(sum L^2): small inputs only.
 * (nvcc may contract the mul+add into an FMA; gcc -O2 on x86-64 without -mfma
 *  does not, so this is "same formula", not guaranteed bit-identical.)
 */
void gcp_oracle_cumprod_bwd_refloop_f32(const float *x, const float *y, const float *g,
                                         const int32_t *inv, const int32_t *seg_end,
                                         float *gin, int64_t n) {
    for (int64_t idx = 0; idx < n; ++idx) {
        int64_t end = seg_end[inv[idx]];
        float p = (x[idx] != 0.0f) ? x[idx] : 1e-8f;
        float val = 0.0f;
        for (int64_t i = idx; i < end; ++i) val += g[i] * (y[i] / p);
        gin[idx] = val;
    }
}

/*
 * The same reference formula evaluated in fp64 in O(n): a reverse running sum
 * of g[k]*y[k] inside each segment, divided by p.  Mathematically identical to
 * the loop above (sum_k g_k*y_k/p == (sum_k g_k*y_k)/p); used to check the GPU
 * path against the *reference semantics* (including its x==0 -> 0 behaviour,
 * SURVEY.md §3.6-4) at full size.
 * Segment end = seg_end[inv[idx]], exactly as the reference indexes it.
 */
void gcp_oracle_cumprod_bwd_ref_f64(const float *x, const float *y, const float *g,
                                     const int32_t *inv, const int32_t *seg_end,
                                     double *gin, int64_t n) {
    double suffix = 0.0;
    for (int64_t idx = n - 1; idx >= 0; --idx) {
        int64_t end = seg_end[inv[idx]];
        if (idx == end - 1) suffix = 0.0; /* tail of its segment */
        suffix += (double)g[idx] * (double)y[idx];
        double p = (x[idx] != 0.0f) ? (double)x[idx] : (double)1e-8f;
        gin[idx] = suffix / p;
    }
}

/*
 * Exact gradient of L = sum_k g_k * y_k,  y = segmented inclusive cumprod(x):
 *   dL/dx_i = E_i * S_i,
 *   E_i = prod_{j<i, same seg} x_j   (1 at the segment head)
 *   S_i = g_i + x_{i+1} * S_{i+1}    (S = g at the segment tail)
 * Division-free (自動微分の成分表示.md eq.(14) re-derived without the 1/x_i), so it is
 * exact where x_i == 0.  Equals the reference formula wherever no x in the
 * segment is 0.  Segments: adjacent runs of equal inv.
 */
void gcp_oracle_cumprod_bwd_exact_f64(const float *x, const float *g, const int32_t *inv,
                                       double *gin, int64_t n) {
    /* pass 1 (reverse): S_i into gin */
    double s = 0.0;
    for (int64_t i = n - 1; i >= 0; --i) {
        int tail = (i == n - 1) || (inv[i + 1] != inv[i]);
        s = tail ? (double)g[i] : (double)g[i] + (double)x[i + 1] * s;
        gin[i] = s;
    }
    /* pass 2 (forward): multiply by E_i */
    double e = 1.0;
    for (int64_t i = 0; i < n; ++i) {
        if (HEAD(inv, i)) e = 1.0;
        gin[i] *= e;
        e *= (double)x[i];
    }
}

/* ------------------------------------------------------------------------- */
/* CPU baseline ("port"): fp32, all host threads, parallel over segments.     */
/* Two phases: (1) find segment starts (parallel), (2) per segment fwd+bwd.   */
/* Timed by bench.py's cpu_baseline leg and by `bench.py --impl reference`.   */
/* ------------------------------------------------------------------------- */

/* returns number of segments; starts[] must hold n+1 entries (worst case) */
int64_t gcp_oracle_segment_starts(const int32_t *key, int64_t n, int64_t *starts) {
    int64_t k = 0;
    for (int64_t i = 0; i < n; ++i)
        if (HEAD(key, i)) starts[k++] = i;
    starts[k] = n;
    return k;
}

int gcp_oracle_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* fwd (y) + exact division-free bwd (gin) in fp32 for segments [starts[s], starts[s+1]) */
void gcp_oracle_fwd_bwd_f32_omp(const float *x, const float *g, const int64_t *starts,
                                int64_t nseg, float *y, float *gin, int nthreads) {
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
    (void)nthreads;
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t s = 0; s < nseg; ++s) {
        int64_t b = starts[s], e = starts[s + 1];
        float acc = 1.0f;
        for (int64_t i = b; i < e; ++i) {
            acc = (i == b) ? x[i] : acc * x[i];
            y[i] = acc;
        }
        float sfx = 0.0f;
        for (int64_t i = e - 1; i >= b; --i) {
            sfx = (i == e - 1) ? g[i] : g[i] + x[i + 1] * sfx;
            float ex = (i == b) ? 1.0f : y[i - 1];
            gin[i] = ex * sfx;
        }
    }
}
