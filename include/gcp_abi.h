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

#define GCP_ABI_VERSION 2

#define GCP_OK 0
#define GCP_ERR_INVALID_ARG (-1)
#define GCP_ERR_WORKSPACE (-2)   /* workspace null / too small / misaligned */
#define GCP_ERR_WATCHDOG (-3)    /* a bounded spin expired inside a kernel: the results of that op are invalid and the
                                  * workspace is poisoned until re-initialised (gcp_workspace_status / the attached flag) */
#define GCP_ERR_SEGMENTS (-4)    /* inv / seg_end inconsistent (gcp_validate_segments) */

/* elements per call of the three scan ops (int32 indexing like the reference, minus the alignment peel) */
#define GCP_MAX_ELEMENTS 2147483644LL

/* opaque: a cudaStream_t */
typedef void *gcp_stream_t;

int gcp_abi_version(void);

/* Bytes of workspace needed for calls with up to n elements (n >= 0). */
size_t gcp_workspace_bytes(int64_t n);

/* Zero the workspace (once after allocation).  Async on `stream`. */
int gcp_workspace_init(void *ws, size_t ws_bytes, gcp_stream_t stream);

/* Synchronises `stream`, then reports the sticky watchdog flag of the workspace
 * in *status (GCP_OK or GCP_ERR_WATCHDOG). */
int gcp_workspace_status(const void *ws, gcp_stream_t stream, int *status);

/* Failing loudly without synchronising.  The kernels' waits (mbarriers fed by the CTA's own producer warp, the
 * grid barrier of the cooperative launch) are bounded; should one ever expire, the kernel finishes with invalid
 * results and raises a sticky flag in the workspace.  Attach a 4-byte word of PINNED, device-mapped host memory
 * (cudaHostAlloc / cudaHostRegister; pass its device-visible address — the same pointer under unified
 * addressing) and the kernel also stores 1 there: the owner checks the word with a plain host read before every
 * call on that workspace and treats non-zero as GCP_ERR_WATCHDOG (ops.py does; it then re-initialises the
 * workspace).  Call after gcp_workspace_init (which clears the attachment); the word must outlive the workspace's
 * use.  Pass NULL to detach.  Async on `stream`. */
int gcp_workspace_attach_flag(void *ws, size_t ws_bytes, void *host_flag_device_address, gcp_stream_t stream);
/* Test hook: raises the watchdog flag of `ws` from a kernel, exactly as an expired wait would (tests of the
 * fail-loudly path; the workspace must be re-initialised afterwards). */
int gcp_workspace_selftest_abort(void *ws, size_t ws_bytes, gcp_stream_t stream);

/*
 * Replaces grouped_cumprod_forward(x, key, y)
 *   (/root/reference/cuda_kernel/grouped_cumprod_forward.cu:6-24):
 *   y[i] = x[i]              if i == 0 or key[i] != key[i-1]
 *        = y[i-1] * x[i]     otherwise                       (inclusive)
 * x f32[n], key i32[n], y f32[n] (written in place).  n may be 0; at most GCP_MAX_ELEMENTS = 2^31 - 4 per call
 * (the reference ops index with int, grouped_cumprod_backward.cu:52; more is GCP_ERR_INVALID_ARG).  Longer arrays
 * are cut at segment boundaries into independent calls — ops.py does that (the counterpart of the reference's
 * chunk loop, gs_model.py:428, :675; no per-pixel carry is needed because no segment is ever split).
 * Any 4-byte alignment.  The fast path (one persistent kernel, cooperative launch, TMA) needs x and key in the
 * SAME 16-byte phase — true for equal slices of separately allocated tensors, the reference's [cutting_number:]
 * case — and peels the 1-3 leading elements; other layouts, n below one tile, or a device that refuses the
 * cooperative launch are served by the plain-load kernel pair (K1 + fix-up), about 0.7 / 0.4 of the fast
 * path's bandwidth.
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
 * Test hook: choose the path used by subsequent calls of this process.  op: 0 = forward scans, 1 = backward.
 * variant: -1 = default (1, falling back to 0 where 1 cannot serve the call), 0 = the plain-load kernel pair,
 * 1 = the persistent blocked kernel (gcp_variant_name describes them).  GCP_ERR_INVALID_ARG for unknown values.
 */
int gcp_set_variant(int op, int variant);
/* Tuning options.  option 0 (GCP_OPT_HALO): 1 (default) = the persistent kernels resolve each tile's
 * cross-tile carry from the 128 elements beside the tile and touch the look-back descriptors only
 * for segments longer than that; 0 = always use the decoupled look-back. */
#define GCP_OPT_HALO 0
/* option 1 (GCP_OPT_CHAIN): the blocked backward kernel gives every CTA one contiguous range of tiles instead of
 * tickets and hands each tile's outgoing carry to the next in registers: inside a segment spanning many tiles only
 * the first tile of a CTA's range needs the fix-up phase, and only that tile reads the halo window.  1 (default) =
 * always, 0 = tickets, 2 = chained when more than 1/32 of the tiles of the op that ran last on the workspace lay
 * strictly inside a segment. */
#define GCP_OPT_CHAIN 1
/* option 2 (GCP_OPT_CHAIN_FWD): the same schedule for the blocked forward kernel; default 0 (tickets), which is
 * 5-7 % faster there (C3 0.93 vs 0.86, C4 0.94 vs 0.90 of the copy peak). */
#define GCP_OPT_CHAIN_FWD 2
int gcp_set_option(int option, int value);
int gcp_num_variants(int op);
const char *gcp_variant_name(int op, int variant);

/* Number of kernels launched by the last call on this thread (for bench.py's gpu_launches). */
int gcp_last_launch_count(void);

/* ------------------------------------------------------------------------------------------------
 * Compositor rows: streaming kernels around the two scans.  They replace the
 * torch op chains of /root/reference/gs_model.py:480-514 (box expansion, Gaussian kernel, pixel
 * accumulation), :538-548 (pixel key, sort, gather) and :627-663,:733-783 (per-element gradients,
 * scatter to Gaussians).  Per-Gaussian tables: mean f32[n,2], lam f32[n,2,2], opac f32[n], l_d f32[n,3];
 * sp/ep i32[n,2] inclusive box corners; goff i64[n+1] exclusive element offsets (cumsum of boxsize).
 * Sorted element list: key_s i32[N] (y*10000+x, non-decreasing), gid_s i32[N] (Gaussian of the element;
 * inside a pixel in depth = index order).  image f32[(H+1)*(W+1)*3], as gs_model.py:505.
 * ------------------------------------------------------------------------------------------------ */

/* key[e], gid[e] of every element in Gaussian-major order (make_rect_points_parallel, uitility.py:336-366,
 * + unique(), gs_model.py:538-541). */
int gcp_splat_expand(const int32_t *sp, const int32_t *ep, const int64_t *goff, int64_t n, int64_t N,
                     int32_t *key, int32_t *gid, gcp_stream_t stream);

/* Stable sort of the (key, gid) pairs by key (replaces torch.sort + gathers, gs_model.py:547-548). */
size_t gcp_splat_sort_bytes(int64_t N);
int gcp_splat_sort(const int32_t *key_in, const int32_t *gid_in, int32_t *key_out, int32_t *gid_out, int64_t N,
                   int max_key, void *temp, size_t temp_bytes, gcp_stream_t stream);

/* Sort-free construction of the SAME sorted (key_s, gid_s) list (preferred over expand + sort): a counting
 * placement that exploits that the Gaussians arrive in depth order — per image row, the intervals [sx,ex] of
 * the boxes crossing the row are walked in Gaussian order and every element lands at
 * offset[pixel] + (number of earlier Gaussians on that pixel), i.e. exactly its stable-sort position.
 * poff i64[n+1] = exclusive offsets of (box height) x (number of strips the box touches) =
 * (ey-sy+1) * ((ex>>S) - (sx>>S) + 1) with S = gcp_splat_seg_shift(), P = poff[n].  seg_off i32[(H+1)*(W+1)+1]
 * receives the offset of every pixel list (pixels in key order; last entry = N).  Bit-identical output.
 * key_s must be 16-byte aligned (vector stores). */
size_t gcp_splat_place_bytes(int64_t P, int W, int H);
int gcp_splat_set_fill_blocks(int blocks); /* tuning hook: persistent grid of the fill kernel, 0 = default */
/* tuning hook: from this many (cell, Gaussian) pairs per pixel (P / pixels) on, lists count as long and the
 * warp-per-list key kernel and the transposed, ballot-compacting fill are used (default 8) */
int gcp_splat_set_long_list_threshold(int pairs_per_pixel);
int gcp_splat_seg_shift(void);             /* log2 of the strip width the library was built with */
int gcp_splat_place(const int32_t *sp, const int32_t *ep, const int64_t *poff, int64_t n, int64_t P, int W, int H,
                    int32_t *key_s, int32_t *gid_s, int32_t *seg_off, int32_t *cell_start, int32_t *pair_gid,
                    int32_t *batch_table, void *temp, size_t temp_bytes, gcp_stream_t stream);
/* Optional outputs (NULL: not kept) that gcp_splat_bwd_elem_cells needs in the backward of a long-list view:
 * cell_start i32[gcp_splat_num_cells(W,H)+1] and pair_gid i32[P] = the (cell, Gaussian) pair list sorted by cell;
 * batch_table i32[gcp_splat_batch_table_ints(P,W,H)] = per batch of 32 pairs of a cell: cell, first pair, and the
 * list position of each of the strip's pixels at the start of the batch (written only when
 * gcp_splat_long_lists(P,W,H)). */
int64_t gcp_splat_batch_table_ints(int64_t P, int W, int H);
int gcp_splat_num_cells(int W, int H);
int gcp_splat_long_lists(int64_t P, int W, int H); /* 1 if the long-list kernels are selected for this view */

/* View prologue in one call: goff i64[n+1] = exclusive offsets of boxsize i64[n] (element offsets per Gaussian,
 * the reference's cumsum, gs_model.py:427), poff i64[n+1] = exclusive offsets of the (cell, Gaussian) pair
 * counts gcp_splat_place expects, totals i64[2] = {N, P} (device memory; one D2H copy gives the host both). */
size_t gcp_splat_prepare_bytes(int64_t n);
int gcp_splat_prepare(const int64_t *boxsize, const int32_t *sp, const int32_t *ep, int64_t n, int64_t *goff,
                      int64_t *poff, int64_t *totals, void *temp, size_t temp_bytes, gcp_stream_t stream);

/* The per-Gaussian tables packed into two 32-byte records per Gaussian (both arrays 32-byte aligned), so that
 * every per-element gather of the kernels below is a single L2 sector:
 *   rec_a f32[n,8] = {mx, my, l00, l01, l10, l11, opacity, 0}
 *   rec_b i32[n,8] = {bits(l0), bits(l1), bits(l2), sx, sy, box width, goff low, goff high}
 * (mean f32[n,2], lam f32[n,4] row-major Lambda, opac f32[n], l_d f32[n,3], sp/ep i32[n,2], goff i64[n+1]). */
int gcp_splat_pack(const float *mean, const float *lam, const float *opac, const float *l_d, const int32_t *sp,
                   const int32_t *ep, const int64_t *goff, int64_t n, float *rec_a, int32_t *rec_b,
                   gcp_stream_t stream);

/* x_s[e] = 1 - opacity * exp(-1/2 (r-m) Lambda (r-m)^T)  (gs_model.py:493-495, :533-535), sorted order. */
int gcp_splat_alpha(const int32_t *key_s, const int32_t *gid_s, const float *rec_a, int64_t N, float *x_s,
                    gcp_stream_t stream);

/* image[pixel] += sum_i T_i alpha_i l_i with T_i the EXCLUSIVE product taken from the inclusive scan
 * `incl` (no division; replaces gs_model.py:562, :498-514).  image must be zeroed by the caller. */
int gcp_splat_color(const float *incl, const float *x_s, const int32_t *key_s, const int32_t *gid_s,
                    const int32_t *rec_b, int64_t N, int W, float *image, gcp_stream_t stream);

/* gshift[k] = w_{k+1} inside a pixel list (0 at its tail), w_k = <dL/dI(pixel), alpha_k l_k>: the grad_out
 * for which gcp_cumprod_bwd_f32 returns T_k*U_k (division-free replacement of gs_model.py:716-722). */
int gcp_splat_bwd_w(const float *incl, const float *x_s, const int32_t *key_s, const int32_t *gid_s,
                    const int32_t *rec_b, const float *grad_image, int64_t N, int W, float *gshift,
                    gcp_stream_t stream);

/* Backward in two atomic-free steps (preferred over gcp_splat_bwd_grads):
 *  gcp_splat_bwd_elem   (sorted order) writes (dalpha, d) of every element at its Gaussian-major position
 *                       goff[g] + (y-sy)*w + (x-sx) of elem f32[N,2] — the un-sort, without a permutation;
 *  gcp_splat_bwd_reduce (Gaussian-major) sums the per-element gradients of gs_model.py:733-766 over each box
 *                       (replaces scatter_reduce, :776-783); deterministic, balanced over box sizes (small
 *                       boxes 8 lanes each, large boxes cut into 1024-element pieces).  temp: at least
 *                       gcp_splat_bwd_reduce_bytes(N, n) bytes of scratch, 16-byte aligned. */
int gcp_splat_bwd_elem(const float *incl, const float *x_s, const float *tu, const int32_t *key_s,
                       const int32_t *gid_s, const int32_t *rec_b, const float *grad_image, int64_t N, int W,
                       float *elem, gcp_stream_t stream);
/* Same result as gcp_splat_bwd_elem for views with long pixel lists (gcp_splat_long_lists): the un-sort as a
 * shared-memory transposition over the placement's batches, every global access coalesced.  seg_off,
 * cell_start, pair_gid, batch_table: outputs of gcp_splat_place for the same view. */
int gcp_splat_bwd_elem_cells(const float *incl, const float *x_s, const float *tu, const int32_t *rec_b,
                             const float *grad_image, const int32_t *seg_off, const int32_t *cell_start,
                             const int32_t *pair_gid, const int32_t *batch_table, int64_t P, int W, int H,
                             float *elem, gcp_stream_t stream);
size_t gcp_splat_bwd_reduce_bytes(int64_t N, int64_t n);
int gcp_splat_bwd_reduce(const float *elem, const int32_t *sp, const int32_t *ep, const int64_t *goff,
                         const float *mean, const float *lam, const float *opac, const float *l_d, int64_t N,
                         int64_t n, float *g_mean, float *g_lam, float *g_opac, float *g_l, void *temp,
                         size_t temp_bytes, gcp_stream_t stream);

/* Per-element gradients (gs_model.py:733-766) accumulated per Gaussian (:776-783); tu = T*U from
 * gcp_cumprod_bwd_f32.  Outputs must be zeroed by the caller: g_mean[n,2], g_lam[n,4], g_opac[n], g_l[n,3]. */
int gcp_splat_bwd_grads(const float *incl, const float *x_s, const float *tu, const int32_t *key_s,
                        const int32_t *gid_s, const float *mean, const float *lam, const float *opac,
                        const float *l_d, const float *grad_image, int64_t N, int W, float *g_mean, float *g_lam,
                        float *g_opac, float *g_l, gcp_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Fused compositor route (csrc/gcp_tile.cu).  The same per-pixel segmented scan —
 * T_i = prod_{j<i}(1-alpha_j), C = sum T_i alpha_i l_i, gs_model.py:544-566 + :498-514 forward, :627-663 +
 * :733-783 backward — evaluated one pixel per lane without materialising the element lists: the image is cut
 * into tiles of gcp_tile_width() x gcp_tile_height() = 32 pixels, a box contributes one (tile, Gaussian) pair
 * per tile it touches, every tile's pairs are put in Gaussian (= depth) order — the order of the reference's
 * torch.sort at gs_model.py:547, hence the order inside every pixel list — and one warp walks each tile's list.
 * No float atomics: results are bitwise reproducible.  No library (CUB / thrust) kernels.  Boxes are clipped
 * to [0,W] x [0,H].
 *
 * A view is THREE calls — gcp_view_plan, gcp_view_render (or both at once: gcp_view_forward), gcp_view_backward —
 * replacing custom_autograd_grouped_cumprod.forward / .backward (gs_model.py:666-692, :786-820), and lives in two
 * caller-owned device arenas (256-byte aligned; nothing is allocated, freed or retained by the library):
 *   plan arena  gcp_view_plan_bytes(n, W, H): pair counts / offsets per Gaussian and per tile, packed records,
 *               work lists.  Written by gcp_view_plan + gcp_view_render, read by gcp_view_backward.
 *   pair arena  gcp_view_pair_bytes(pair_cap, W, H): the tile-ordered pair list, one checkpoint of T per 8 pairs
 *               (all the backward keeps of the forward walk: 16 B per pair), the state of the pieces of long
 *               lists, the gradient partials (32 B per pair).  pair_cap >= the pair count of the view.
 * gcp_view_plan also stores the pair count into totals_host[0] — a pointer the DEVICE can write: pinned, mapped
 * host memory (cudaHostAlloc; the same address under unified addressing), int64[2], or NULL.  The host waits
 * for the plan (an event behind it) and sizes the pair arena from it; this is the one host sync of a view, the
 * counterpart of the reference's .item() at uitility.py:348.  A caller that already owns a pair arena it
 * believes large enough may skip the wait and call gcp_view_forward: when the view turns out to have more
 * pairs than pair_cap the render kernels do nothing (the image is NOT written) and the caller must compare
 * totals_host[0] with pair_cap before using any result, and redo the view with a larger arena.
 * A PIECE is at most gcp_tile_piece_pairs() (default 128) consecutive pairs of one tile: longer lists are cut
 * into pieces walked by different warps, the per-pixel carries between them (T forward, U backward — the
 * segmented scan's cross-block carries) resolved by two small combine kernels; the setting must not change
 * between the render and the backward of a view.
 * Inputs: sp/ep i32[n,2] (8-byte aligned), mean f32[n,2], lam f32[n,4], opac f32[n], l_d f32[n,3] (any 4-byte
 * alignment), grad_image / image f32[(H+1)*(W+1)*3].  image is written completely (no need to zero it); so are
 * g_mean[n,2], g_lam[n,4], g_opac[n], g_l[n,3] (d_l = (sum d)/l, gs_model.py:763-766).  keep = 0: a render no
 * backward will follow (no checkpoints are written).  n < 2^31 - 64, pairs < 2^31 - 64 per view, any number of
 * ELEMENTS (the reference's 2^29-element chunks and their per-pixel carry, gs_model.py:428, :582-594, have no
 * counterpart: a view is one pass); the arenas of a view at that limit would exceed 100 GB, compositor.py raises.
 * Binning: the Gaussian-major pair list is put in tile order by a stable radix sort on the tile id (no atomics,
 * nothing to sort afterwards) — inside a tile the Gaussians keep their depth order, as after the reference's stable
 * torch.sort (gs_model.py:547).
 * ------------------------------------------------------------------------------------------------ */
int gcp_tile_width(void);
int gcp_tile_height(void);
int gcp_tile_num_tiles(int W, int H);
int gcp_tile_set_piece_pairs(int pairs);   /* tuning / tests: multiple of 32 */
int gcp_tile_piece_pairs(void);
size_t gcp_view_plan_bytes(int64_t n, int W, int H);
size_t gcp_view_pair_bytes(int64_t pair_cap, int W, int H);
int gcp_view_plan(const int32_t *sp, const int32_t *ep, int64_t n, int W, int H, void *plan, size_t plan_bytes,
                  int64_t *totals_host, gcp_stream_t stream);
int gcp_view_render(const int32_t *sp, const int32_t *ep, const float *mean, const float *lam, const float *opac,
                    const float *l_d, int64_t n, int W, int H, void *plan, size_t plan_bytes, void *pairs,
                    size_t pair_bytes, int64_t pair_cap, int keep, float *image, gcp_stream_t stream);
int gcp_view_forward(const int32_t *sp, const int32_t *ep, const float *mean, const float *lam, const float *opac,
                     const float *l_d, int64_t n, int W, int H, void *plan, size_t plan_bytes, void *pairs,
                     size_t pair_bytes, int64_t pair_cap, int keep, float *image, int64_t *totals_host,
                     gcp_stream_t stream);
int gcp_view_backward(void *plan, size_t plan_bytes, void *pairs, size_t pair_bytes, int64_t pair_cap,
                      const float *grad_image, int64_t n, int W, int H, float *g_mean, float *g_lam, float *g_opac,
                      float *g_l, gcp_stream_t stream);
/* The same backward with the view's gradients ADDED into parameter-sized arrays: row index[g] of g_mean f32[*,2]
 * (8-byte aligned), g_lam f32[*,4] (16-byte aligned), g_opac f32[*], g_l f32[*,3] receives Gaussian g's gradient.
 * This is what autograd makes of the reference's per-view selection `param[mask]` (gs_model.py:405-413) when the
 * views of a batch are summed (gs_control.py:180-185): a scatter-add per view.  index holds distinct rows;
 * index == NULL is gcp_view_backward (row g of view-sized arrays, overwritten). */
int gcp_view_backward_scatter(void *plan, size_t plan_bytes, void *pairs, size_t pair_bytes, int64_t pair_cap,
                              const float *grad_image, int64_t n, int W, int H, const int32_t *index, float *g_mean,
                              float *g_lam, float *g_opac, float *g_l, gcp_stream_t stream);
/* A batch of views in ONE call: the per-view loop of gs_model.py:402-449 around the compositor — render, loss
 * gradient, backward — with every view's gradients added into the parameters' gradient arrays (gcp_view_backward_scatter).
 * Nothing waits for the device: every view is rendered on the caller's pair capacity, its pair count is stored in
 * totals_host[v] (pinned, device-writable), and a view with more pairs than pair_cap is skipped as a whole — the
 * caller reads the counts after its next synchronisation and repeats the step on larger arenas if any exceeds
 * pair_cap.  Views alternate between `lanes` streams (lane 0 = the caller's stream), each lane with its own plan
 * and pair arena (plan[lane], pairs[lane], sized for the largest view: gcp_view_plan_bytes(max n),
 * gcp_view_pair_bytes(pair_cap)); the scatter-adds are chained in view order, so every sum has a fixed order.
 * On return the caller's stream waits for all lanes. */
#define GCP_VIEWS_MAX_LANES 4
typedef struct gcp_view_desc {
    const int32_t *sp, *ep;      /* i32[n,2] inclusive box corners, clamped to the image (gs_model.py:419-425) */
    const float *mean, *lam, *opac, *l_d;   /* f32[n,2], f32[n,4], f32[n], f32[n,3], depth order = index order */
    const int32_t *index;        /* i32[n] parameter row of every Gaussian (the view's selection), or NULL: row g */
    const float *target;         /* f32[(H+1)(W+1)3] or NULL.  Not NULL: grad_image is WRITTEN as the gradient of
                                    mean((image - target)^2) and *loss += that mean; NULL: grad_image is an input */
    float *grad_image;           /* f32[(H+1)(W+1)3] dL/d image */
    float *image;                /* f32[(H+1)(W+1)3] out (views of the same lane may share one buffer) */
    int64_t n;
} gcp_view_desc;
typedef struct gcp_views_ctx gcp_views_ctx;   /* the side streams and ordering events of a batch */
int gcp_views_ctx_create(int lanes, gcp_views_ctx **out);
void gcp_views_ctx_destroy(gcp_views_ctx *ctx);
int gcp_views_step(gcp_views_ctx *ctx, const gcp_view_desc *views, int n_views, int W, int H, void *const *plan,
                   size_t plan_bytes, void *const *pairs, size_t pair_bytes, int64_t pair_cap, float *g_mean,
                   float *g_lam, float *g_opac, float *g_l, float *loss, int64_t *totals_host, gcp_stream_t stream);
/* The same batch with a TAIL: views first_tail_view.. add their gradients into a second set of arrays, and `event`
 * (a cudaEvent_t, or NULL) is recorded as soon as the last view in front of the tail has been added to the main
 * arrays — while the tail views are still running.  A data-parallel caller starts the all-reduce of the main
 * bucket on that event and only has the tail's small arrays left to reduce when the batch ends (bench.py,
 * multi_view_leg); the lanes stay full across the split, which two separate batches cannot do. */
typedef struct gcp_views_split {
    int first_tail_view;
    float *g_mean, *g_lam, *g_opac, *g_l;   /* the tail's arrays, same shapes and alignment rules as the main ones */
    void *event;                            /* cudaEvent_t or NULL */
} gcp_views_split;
int gcp_views_step_split(gcp_views_ctx *ctx, const gcp_view_desc *views, int n_views, int W, int H, void *const *plan,
                         size_t plan_bytes, void *const *pairs, size_t pair_bytes, int64_t pair_cap, float *g_mean,
                         float *g_lam, float *g_opac, float *g_l, float *loss, int64_t *totals_host,
                         const gcp_views_split *split, gcp_stream_t stream);
/* Test aid: byte offsets of the integer arrays the tests compare bit for bit with oracle/tile_oracle.py.
 * out[0..5] (plan arena): toff i32[n+1] (Gaussian-major pair offsets), tile_count i32[tiles], tile_start
 * i32[tiles+1], piece_extra i32[tiles], header, records; out[6..7] (pair arena): pair_gid i32[cap],
 * extra-piece table i32[out[8]]. */
int gcp_view_layout(int64_t n, int W, int H, int64_t pair_cap, int64_t *out);
/* kernels launched by the last gcp_view_* call of this thread (bench.py's gpu_launches) */
int gcp_view_last_launch_count(void);

/* ------------------------------------------------------------------------------------------------
 * Host-buffer entry point helpers (csrc/gcp_host.cu, used by simplegaussiansplat_tk71_b200/host.py).  With the
 * element arrays in HOST memory the PCIe link bounds the step; the ops only need to know where the runs of equal
 * adjacent keys start (grouped_cumprod_forward.cu:17-23), so `key` crosses the link as one bit per element.
 * ------------------------------------------------------------------------------------------------ */
/* HOST function (OpenMP, `threads` threads): bits[i>>5] bit (i&31) = 1 iff i == 0 or key[i] != key[i-1];
 * key i32[n] and bits u32[(n+31)/32] are HOST pointers. */
int gcp_host_boundary_bits(const int32_t *key, int64_t n, uint32_t *bits, int threads);
/* Device: ids[i] = number of run starts in (0, i] — dense segment ids 0..k-1, usable as `key` of the forward ops
 * and as `inv` of the backward op.  temp >= gcp_ids_from_bits_bytes(n). */
size_t gcp_ids_from_bits_bytes(int64_t n);
int gcp_ids_from_bits(const uint32_t *bits, int64_t n, int32_t *ids, void *temp, size_t temp_bytes,
                      gcp_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* GCP_ABI_H */
