/*
 * gcp_abi.h — C ABI of libgcp_b200.so: B200 (sm_100a) segmented cumulative
 * product / sum over depth-sorted per-pixel Gaussian lists.
 *
 * This is the drop-in boundary for the reference's torch extension module
 * `grouped_cumprod` (/root/reference/cuda_kernel/cuda_kernel.cpp:17-22).  Each
 * entry point below names the reference function it replaces.  Plain pointers
 * and sizes only: no torch types, nothing allocated or freed inside, nothing
 * retained after return.  All pointers are DEVICE pointers unless stated.  All
 * work is enqueued on `stream` and returns immediately (fully asynchronous).
 *
 * Return value: 0 (GCP_OK) or a negative GCP_ERR_* / a positive cudaError_t.
 * Nothing throws across this boundary.
 *
 * Segment layout (identical to the reference):
 *   - forward ops: `key` i32[n]; a new segment starts at i == 0 and wherever
 *     key[i] != key[i-1] (adjacent-run semantics of thrust::inclusive_scan_by_key,
 *     grouped_cumprod_forward.cu:17-23).  Keys need not be globally sorted.
 *   - backward op: `inv` i32[n] dense segment ids 0..k-1 in non-decreasing
 *     order (cuda_test.py:21) and `seg_end` i32[k], the EXCLUSIVE end offset of
 *     each segment (cuda_test.py:27, "inv_len").
 *
 * Workspace: caller-owned device buffer of gcp_workspace_bytes(n_max) bytes,
 * zero-initialised ONCE with gcp_workspace_init and then reusable for any call
 * with n <= n_max on the same stream (the kernels reset it themselves; no
 * per-call memset).  Use one workspace per stream.
 */
#ifndef GCP_ABI_H
#define GCP_ABI_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GCP_ABI_VERSION 1

#define GCP_OK 0
#define GCP_ERR_INVALID_ARG (-1)
#define GCP_ERR_WORKSPACE (-2)   /* workspace null / too small / misaligned */
#define GCP_ERR_WATCHDOG (-3)    /* a bounded spin expired inside a kernel (reported by gcp_workspace_status) */
#define GCP_ERR_SEGMENTS (-4)    /* inv / seg_end inconsistent (gcp_validate_segments) */

/* opaque: a cudaStream_t */
typedef void *gcp_stream_t;

int gcp_abi_version(void);

/* Bytes of workspace needed for calls with up to n elements (n >= 0). */
size_t gcp_workspace_bytes(int64_t n);

/* Zero the workspace (once after allocation).  Async on `stream`. */
int gcp_workspace_init(void *ws, size_t ws_bytes, gcp_stream_t stream);

/* Synchronises `stream`, then reports the sticky watchdog flag of the workspace
 * in *status (GCP_OK or GCP_ERR_WATCHDOG).  Test / debug aid. */
int gcp_workspace_status(const void *ws, gcp_stream_t stream, int *status);

/*
 * Replaces grouped_cumprod_forward(x, key, y)
 *   (/root/reference/cuda_kernel/grouped_cumprod_forward.cu:6-24):
 *   y[i] = x[i]              if i == 0 or key[i] != key[i-1]
 *        = y[i-1] * x[i]     otherwise                       (inclusive)
 * x f32[n], key i32[n], y f32[n] (written in place).  Any alignment (a 16-byte
 * aligned fast path is chosen automatically).  n may be 0.
 */
int gcp_cumprod_fwd_f32(const float *x, const int32_t *key, float *y, int64_t n,
                        void *ws, size_t ws_bytes, gcp_stream_t stream);

/*
 * Replaces grouped_cumsum_forward(x, key, y)
 *   (/root/reference/cuda_kernel/grouped_cumsum_forward.cu:6-24): same with +.
 */
int gcp_cumsum_fwd_f32(const float *x, const int32_t *key, float *y, int64_t n,
                       void *ws, size_t ws_bytes, gcp_stream_t stream);

/*
 * Replaces grouped_cumprod_backward(param, param_cumprod, grad_out, inv, grad_in, inv_len)
 *   (/root/reference/cuda_kernel/grouped_cumprod_backward.cu:43-65, kernel :9-41).
 * Computes dL/dx for L = sum_k grad_out[k] * y[k], y = segmented inclusive cumprod(x):
 *   grad_in[i] = E_i * S_i,  E_i = prod_{j<i in seg} x_j,  S_i = g_i + x_{i+1} S_{i+1}
 * which equals the reference's  sum_{k>=i} g_k * y_k / x_i  wherever no x in the
 * segment is 0, and is exact (non-zero) where the reference returns 0 because of
 * its division.  One pass, O(n), division-free.
 *   x (param) f32[n], y (param_cumprod = forward output) f32[n], gout f32[n],
 *   inv i32[n], seg_end (inv_len) i32[k], gin f32[n] out.
 * `y` is read only at tile boundaries (one value per 2-4 Ki elements); `seg_end`
 * is accepted for signature parity and is implied by `inv` (tail <=> inv[i+1] !=
 * inv[i]); gcp_validate_segments checks the two agree.
 */
int gcp_cumprod_bwd_f32(const float *x, const float *y, const float *gout,
                        const int32_t *inv, const int32_t *seg_end, float *gin,
                        int64_t n, int64_t k, void *ws, size_t ws_bytes,
                        gcp_stream_t stream);

/*
 * Bit-exact check of the integer side of the contract: inv non-decreasing dense
 * ids 0..k-1, seg_end[s] == 1 + last index with inv == s.  Synchronises; writes
 * the number of violations to *violations (host pointer).
 */
int gcp_validate_segments(const int32_t *inv, const int32_t *seg_end, int64_t n, int64_t k,
                          void *ws, size_t ws_bytes, gcp_stream_t stream, int64_t *violations);

/*
 * Tuning hook (bench / tests): choose the kernel variant used by subsequent calls
 * of this process.  op: 0 = forward scans, 1 = backward.  variant: -1 = default
 * heuristic, otherwise an index into the table printed by gcp_variant_name.
 * Returns GCP_ERR_INVALID_ARG for unknown values.
 */
int gcp_set_variant(int op, int variant);
/* Tuning options.  option 0 (GCP_OPT_HALO): 1 (default) = the persistent kernels resolve each tile's
 * cross-tile carry from the 128 elements beside the tile and touch the look-back descriptors only
 * for segments longer than that; 0 = always use the decoupled look-back. */
#define GCP_OPT_HALO 0
int gcp_set_option(int option, int value);
int gcp_num_variants(int op);
const char *gcp_variant_name(int op, int variant);

/* Number of kernels launched by the last call on this thread (for bench.py's gpu_launches). */
int gcp_last_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* GCP_ABI_H */
