// gcp_tile.cu — the fused compositor route: exclusive transmittance, colour sum and the
// division-free backward in ONE walk per direction, without materialising the per-pixel element lists, behind two
// or three C-ABI calls per view (gcp_view_plan + gcp_view_render, or gcp_view_forward; gcp_view_backward) that work
// entirely inside caller-owned arenas, and one call per batch of views (gcp_views_step).
//
// The per-pixel segmented scan  T_i = prod_{j<i} (1 - alpha_j),  C = sum_i T_i alpha_i l_i  (gs_model.py:544-566,
// :498-514) is evaluated with one pixel per lane.  The image is cut into tiles of 8 x 4 pixels = one warp; a box
// contributes one (tile, Gaussian) pair per tile it touches; every tile gets the list of its Gaussians in index
// (= depth) order — the order torch.sort gives the reference at gs_model.py:547, hence the order inside every
// pixel list — and a warp walks its tile's list with the running T of its pixels in registers.
//
//   plan      k_view_cnt      : pairs per Gaussian                                              (uitility.py:336-366)
//             k_view_scan     : their exclusive offsets (chained scan); the pair total goes to the host (.item() :348)
//   render    k_view_pack     : packed per-Gaussian records
//             k_view_pairs    : pair q of the Gaussian-major numbering -> {tile, Gaussian}
//             k_view_bin_rowscan + k_view_bin_scatter (+ k_view_bin_hist) per digit: a STABLE radix sort of the pairs by
//                               tile id — inside a tile the pairs keep the Gaussian (= depth) order, which is the
//                               stable torch.sort of the reference, bit for bit                (gs_model.py:538-548)
//             k_view_tiles    : every tile's range in the sorted list; long lists are cut into pieces (the walk
//                               kernels' work units)
//             k_view_render   : alpha = o * exp(-1/2 d Lambda d^T) (:493-495,:533-535), T, colour; only a
//                               CHECKPOINT of T every 8 pairs is kept for the backward (16 B per pair)
//             k_view_combine_fwd / _bwd : carries between the pieces of a long list          (:582-594)
//   backward  k_view_backward : per 8 pairs, (1) T and the Gaussian kernel value are recomputed from the checkpoint
//                               into registers, (2) the pairs are walked in reverse with
//                               U_i = w_{i+1} + (1-alpha_{i+1}) U_{i+1}  (w = <dL/dI, alpha l>),
//                               dL/dalpha_i = T_i (<dL/dI, l_i> - U_i) — no division by 1-alpha (:736,:747,:757
//                               divide), (3) the moments of g*dalpha over the pair's pixels, from which the
//                               reference's per-element gradients (:733-766) follow per Gaussian, are summed
//                               through shared memory (lane = pair x tile row) in a fixed order;
//             k_view_reduce (+ _big) : the partial sums of a Gaussian's pairs added in pair order (:776-783),
//                               written to the view's gradients or scatter-added into the parameters' (gcp_views_step)
//   batch     gcp_views_step  : all views of a training step in one call over alternating stream lanes
// No float atomics anywhere: a pixel belongs to one lane, a partial to one pair — bitwise reproducible.
// Elements whose inclusive product is 0 contribute nothing and get no gradient (gs_model.py:575-578).
// No library kernels (CUB / thrust) on this route.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <algorithm>
#include <new>
#include <utility>

#include "gcp_abi.h"

namespace {

constexpr int TSX = 3, TSY = 2;                 // tile = 8 x 4 pixels, lane = (y & 3) * 8 + (x & 7)
constexpr int TW = 1 << TSX, TH = 1 << TSY;
static_assert(TW * TH == 32, "one tile is one warp");
constexpr int SUB = 8;                          // pairs per T checkpoint
constexpr int SUB_SHIFT = 3;

inline unsigned blocks_for(int64_t work, int per_block, unsigned cap = 0x7fffffffu) {
    int64_t b = (work + per_block - 1) / per_block;
    if (b < 1) b = 1;
    if (b > cap) b = cap;
    return static_cast<unsigned>(b);
}
inline size_t align256(size_t v) { return (v + 255) & ~static_cast<size_t>(255); }

// Programmatic dependent launch: a view is a chain of ~16 kernels on one stream, half of them a few microseconds long.
// Every kernel is launched with the programmatic-serialization attribute and starts with grid_dep_sync(): its blocks
// may be scheduled while the previous kernel drains, wait there until that kernel has completed and its writes are
// visible (griddepcontrol.wait), and then let the next kernel's blocks be scheduled in turn.  Nothing is read or
// written before the wait, so the semantics are those of plain stream order; behind a memset or an event wait the
// attribute changes nothing.
__device__ __forceinline__ void grid_dep_sync() {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
thread_local bool t_dep_launch = true;   // false inside a batch over several lanes (gcp_views_step): there the gaps of
                                         // one lane are filled by the other lanes' kernels, and blocks parked on an SM
                                         // ahead of their turn only take room from them (64 views: 34.6 -> 35.9 ms)
struct DepLaunch {
    dim3 grid, block;
    size_t smem;
    cudaStream_t st;
    template <typename... KArgs, typename... Args>
    void operator()(void (*kern)(KArgs...), Args &&...args) const {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = grid;
        cfg.blockDim = block;
        cfg.dynamicSmemBytes = smem;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        static const bool off = getenv("GCP_NO_DEP_LAUNCH") != nullptr;   // per-kernel timings: a kernel's time then
        cfg.numAttrs = (t_dep_launch && !off) ? 1 : 0;                    // excludes its wait for the one in front
        cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(std::forward<Args>(args))...);   // errors: cudaGetLastError
    }
};

// ---- arena headers (u32 words; zeroed by gcp_view_plan) ----
constexpr int H_TICKET_S1 = 0, H_TICKET_FWD = 2, H_XPIECES = 3, H_NMULTI = 4;
constexpr int H_TICKET_BWD = 6, H_NBIG = 7, H_TICKET_RED = 8;   // adjacent u32 words: reset together by every backward
constexpr int H_P64 = 8;                        // u64 index (byte 64): the pair count of the view
constexpr int HDR_WORDS = 64;

// A view rendered on a caller-provided pair capacity (gcp_view_forward) may turn out to have more pairs than that:
// every kernel behind the plan then does nothing, the host sees the count and redoes the view on a larger arena.
__device__ __forceinline__ bool overflowed(const unsigned int *hdr, int64_t cap) {
    return static_cast<int64_t>(reinterpret_cast<const unsigned long long *>(hdr)[H_P64]) > cap;
}

// box of Gaussian g clipped to the image [0,W] x [0,H] (the caller clamps already, gs_model.py:419-425)
struct Box {
    int sx, sy, ex, ey;
};
__device__ __forceinline__ Box clip_box(const int32_t *__restrict__ sp, const int32_t *__restrict__ ep, int64_t g, int W,
                                        int H) {
    const int2 s = __ldg(reinterpret_cast<const int2 *>(sp) + g), e = __ldg(reinterpret_cast<const int2 *>(ep) + g);
    Box b;
    b.sx = s.x > 0 ? s.x : 0;
    b.sy = s.y > 0 ? s.y : 0;
    b.ex = e.x < W ? e.x : W;
    b.ey = e.y < H ? e.y : H;
    return b;
}

// ---------------------------------------------------------------------------------------------------------------
// plan: pairs per Gaussian + packed records
// Lambda is stored pre-multiplied by -log2(e)/2, so that the walk kernels get g = exp(-1/2 d Lambda d^T) as one
// ex2.approx of d Lambda' d^T (relative error 2^-22; expf costs eight instructions, this one two); k_view_reduce
// multiplies the d_mean sums by EXP2_UNSCALE = -2 ln 2 to undo the factor.
// rec[g] = {mx, my, l00', l01' | l10', l11', o, l0 || l1, l2, sx, sy | ex, ey, pair offset, -}
// ---------------------------------------------------------------------------------------------------------------
constexpr float EXP2_SCALE = -0.72134752044448170368f;    // -log2(e) / 2
constexpr float EXP2_UNSCALE = -1.38629436111989061883f;  // 1 / EXP2_SCALE

// pairs per Gaussian = tiles its (clipped) box touches (uitility.py:336-366), and {first tile, tiles per row} of the
// box: what the binning reads instead of the 64-byte records
__global__ void __launch_bounds__(256)
k_view_cnt(const int32_t *__restrict__ sp, const int32_t *__restrict__ ep, int64_t n, int W, int H,
           int32_t *__restrict__ cnt, int2 *__restrict__ tbox) {
    grid_dep_sync();
    const int64_t g = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (g >= n) return;
    const Box b = clip_box(sp, ep, g, W, H);
    const int tx0 = b.sx >> TSX, ty0 = b.sy >> TSY, nx = (b.ex >> TSX) - tx0 + 1;
    int c = 0;
    if (b.ex >= b.sx && b.ey >= b.sy) c = nx * ((b.ey >> TSY) - ty0 + 1);
    cnt[g] = c;
    tbox[g] = make_int2(tx0 | (ty0 << 16), nx);
}

constexpr int BIN_CHUNK_SHIFT = 12;   // the binning's blocks of 4096 pairs (BIN_CHUNK below)

template <bool VEC>
__global__ void __launch_bounds__(256)
k_view_pack(const int32_t *__restrict__ sp, const int32_t *__restrict__ ep, const float *__restrict__ mean,
            const float *__restrict__ lam, const float *__restrict__ opac, const float *__restrict__ l_d, int64_t n,
            int W, int H, const int32_t *__restrict__ toff, int4 *__restrict__ rec, int32_t *__restrict__ bstart,
            int bcount, const unsigned int *__restrict__ hdr, int64_t cap) {
    grid_dep_sync();
    __shared__ int4 stage[4][257];
    const int64_t g0 = static_cast<int64_t>(blockIdx.x) * 256;
    const int64_t g = g0 + threadIdx.x;
    if (g < n) {
        const Box b = clip_box(sp, ep, g, W, H);
        if (bstart != nullptr && !overflowed(hdr, cap)) {   // (beyond 2^31 pairs the 32-bit offsets are meaningless)
            // bstart[m] = the Gaussian that owns pair m * 4096: k_view_pairs' block m starts there (saves it a search)
            const int q0 = __ldg(toff + g), q1 = __ldg(toff + g + 1);
            for (int m = (q0 + (1 << BIN_CHUNK_SHIFT) - 1) >> BIN_CHUNK_SHIFT; m < bcount && (static_cast<int64_t>(m) << BIN_CHUNK_SHIFT) < q1; ++m)
                bstart[m] = static_cast<int32_t>(g);
        }
        auto f = [](float v) { return __float_as_int(v); };
        float mx, my, l00, l01, l10, l11;
        if (VEC) {
            const float2 m = __ldg(reinterpret_cast<const float2 *>(mean) + g);
            const float4 L = __ldg(reinterpret_cast<const float4 *>(lam) + g);
            mx = m.x; my = m.y; l00 = L.x; l01 = L.y; l10 = L.z; l11 = L.w;
        } else {
            mx = __ldg(mean + 2 * g); my = __ldg(mean + 2 * g + 1);
            l00 = __ldg(lam + 4 * g); l01 = __ldg(lam + 4 * g + 1); l10 = __ldg(lam + 4 * g + 2); l11 = __ldg(lam + 4 * g + 3);
        }
        stage[0][threadIdx.x] = make_int4(f(mx), f(my), f(l00 * EXP2_SCALE), f(l01 * EXP2_SCALE));
        stage[1][threadIdx.x] = make_int4(f(l10 * EXP2_SCALE), f(l11 * EXP2_SCALE), f(__ldg(opac + g)), f(__ldg(l_d + 3 * g)));
        stage[2][threadIdx.x] = make_int4(f(__ldg(l_d + 3 * g + 1)), f(__ldg(l_d + 3 * g + 2)), b.sx, b.sy);
        stage[3][threadIdx.x] = make_int4(b.ex, b.ey, __ldg(toff + g), 0);
    }
    __syncthreads();
    const int64_t words = 4 * min(static_cast<int64_t>(256), n - g0);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int j = threadIdx.x + 256 * k;
        if (j < words) rec[4 * g0 + j] = stage[j & 3][j >> 2];
    }
}

// ---------------------------------------------------------------------------------------------------------------
// exclusive pair offsets of the Gaussians (toff): the classic single-pass chained scan — a block publishes its
// aggregate, then walks back over its predecessors' descriptors (32 per step) to the nearest inclusive prefix.  Tile
// order comes from an atomic ticket, so a block only ever waits for blocks that are already running.
// descriptor = status (2 bits: 1 aggregate, 2 inclusive) << 62 | value.
// 8192 Gaussians per block: 1 M Gaussians = 111 blocks, four look-back steps at most (2048 per block: 443 blocks,
// 14 steps: 16.7 us against 13.6; counting the pairs inside this kernel instead of a kernel of its own, 5.6 us,
// was slower: 27.7 us — sixteen consecutive boxes per thread are sixteen uncoalesced loads).
// ---------------------------------------------------------------------------------------------------------------
constexpr int SCAN_ITEMS = 16, SCAN_THREADS = 512, SCAN_TILE = SCAN_ITEMS * SCAN_THREADS;

__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_u64(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__global__ void __launch_bounds__(SCAN_THREADS)
k_view_scan(const int32_t *__restrict__ in, int64_t n, int32_t *__restrict__ out, unsigned int *ticket,
            unsigned long long *desc, unsigned long long *total_dev, int64_t *total_host) {
    grid_dep_sync();
    __shared__ unsigned long long s_warp[SCAN_THREADS / 32];
    __shared__ unsigned long long s_prefix;
    __shared__ unsigned int s_tile;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
    __syncthreads();
    const unsigned int tile = s_tile;
    const int64_t base = static_cast<int64_t>(tile) * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    int v[SCAN_ITEMS];
    unsigned long long sum = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        const int x = (base + i < n) ? in[base + i] : 0;
        v[i] = x;
        sum += static_cast<unsigned long long>(x);
    }
    unsigned long long inc = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned long long t = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += t;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    unsigned long long wpre = 0, agg = 0;
#pragma unroll
    for (int w = 0; w < SCAN_THREADS / 32; ++w) {
        if (w < warp) wpre += s_warp[w];
        agg += s_warp[w];
    }
    if (warp == 0) {
        unsigned long long prefix = 0;
        if (tile == 0) {
            if (lane == 0) st_release_u64(desc, (2ull << 62) | agg);
        } else {
            if (lane == 0) st_release_u64(desc + tile, (1ull << 62) | agg);
            int64_t pb = static_cast<int64_t>(tile) - 1;
            while (true) {
                const int64_t idx = pb - lane;
                unsigned long long d = 2ull << 62;   // before tile 0: an inclusive prefix of 0
                if (idx >= 0) {
                    do { d = ld_acquire_u64(desc + idx); } while ((d >> 62) == 0ull);
                }
                const unsigned inc_mask = __ballot_sync(0xffffffffu, (d >> 62) == 2ull);
                const int stop = inc_mask ? (__ffs(inc_mask) - 1) : 31;
                unsigned long long c = (lane <= stop) ? (d & ((1ull << 62) - 1)) : 0ull;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
                prefix += c;
                if (inc_mask) break;
                pb -= 32;
            }
            if (lane == 0) st_release_u64(desc + tile, (2ull << 62) | (prefix + agg));
        }
        if (lane == 0) s_prefix = prefix;
    }
    __syncthreads();
    unsigned long long run = s_prefix + wpre + (inc - sum);
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        if (base + i < n) out[base + i] = static_cast<int32_t>(run);
        run += static_cast<unsigned long long>(v[i]);
    }
    if (static_cast<int64_t>(tile + 1) * SCAN_TILE >= n && threadIdx.x == SCAN_THREADS - 1) {   // the block holding item n-1
        const unsigned long long tot = s_prefix + agg;
        out[n] = static_cast<int32_t>(tot < 0x7fffffffull ? tot : 0x7fffffffull);
        *total_dev = tot;
        if (total_host != nullptr) {
            *reinterpret_cast<volatile int64_t *>(total_host) = static_cast<int64_t>(tot);
            __threadfence_system();
        }
    }
}

// empty views (n == 0): the offsets of nothing
__global__ void k_view_scan_empty(int32_t *toff, unsigned int *hdr, int64_t *totals_host) {
    grid_dep_sync();
    if (threadIdx.x == 0) {
        toff[0] = 0;
        if (totals_host) { totals_host[0] = 0; __threadfence_system(); }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// render: binning
// ---------------------------------------------------------------------------------------------------------------
// ---------------------------------------------------------------------------------------------------------------
// The Gaussian-major pair list (pair q of a Gaussian = tile q of its box, row-major) is put in tile order by a STABLE
// least-significant-digit radix sort on the tile id — stable, so inside a tile the pairs keep the Gaussian (= depth)
// order they were emitted in: exactly the reference's stable torch.sort of the expanded list (gs_model.py:547),
// bit for bit, with no atomic on a tile counter and no sort afterwards.  (Round 2's first binning took a slot per
// pair from an atomic counter per tile and then sorted every tile's list by Gaussian id: 3.6 M returning atomics cost
// ~45 us per 1080p view, twice that on the bundled scene whose 8 600 tiles are hit ~500 times each, and the lists of
// thousands of ids needed block-wide sorts — 126 / 228 us against 90 / 90 us now.)
//   k_view_pairs        : pair q -> {tile, Gaussian}: a block stages the pair offsets and tile boxes of its Gaussians
//                         in shared memory, every thread emits 8 consecutive pairs (one binary search for the owner of
//                         the first, then a walk); per-block counts of the first digit
//   per digit           : k_view_bin_rowscan scans the [digit value][block] counts along the blocks; k_view_bin_scatter
//                         sorts the block's 4096 pairs by the digit in shared memory — ranked in list order: a warp takes
//                         512 consecutive pairs, 32 per round; one ballot per digit bit gives the lanes holding the same
//                         digit value, the lowest of them bumps the warp's running counter of that value (no atomics);
//                         the warps' counters are then offset in warp order — and copies them out, every digit value's
//                         run to its place (block's row-scan value + start of the value's run = scan of the totals);
//                         k_view_bin_hist counts the next digit
//   k_view_tiles        : every tile's range in the sorted list (lower bound of its id) and its pieces
// Digits: ceil(log2(tiles) / passes) bits each, at most 9 (1080p: 2 x 8, 4K: 2 x 9, the largest image: 3 x 9).
// ---------------------------------------------------------------------------------------------------------------
constexpr int BIN_THREADS = 256, BIN_WARPS = BIN_THREADS / 32, BIN_ROUNDS = 16;
constexpr int BIN_PER_WARP = 32 * BIN_ROUNDS;             // 512 consecutive pairs per warp
constexpr int BIN_CHUNK = BIN_THREADS * BIN_ROUNDS;       // 4096 pairs per block
static_assert(BIN_CHUNK == 1 << BIN_CHUNK_SHIFT, "k_view_pack marks the first Gaussian of every block of pairs");
constexpr int BIN_MAX_BITS = 9;

constexpr int PW_GAUSS = BIN_CHUNK + 64;   // Gaussians a block of k_view_pairs can stage
constexpr int PW_SMEM = (PW_GAUSS + 2) * 4 + PW_GAUSS * 8 + (4 << BIN_MAX_BITS);

// Where k_view_pairs reads the pair offsets and tile boxes of Gaussian g_lo + j: shared memory (staged) or global.
struct PairSourceShared {
    const int *toff;      // [ng + 1]
    const int2 *box;      // [ng] {tx0 | ty0 << 16, nx}
    __device__ __forceinline__ int off(int j) const { return toff[j]; }
    __device__ __forceinline__ int2 tiles(int j) const { return box[j]; }
};
struct PairSourceGlobal {
    const int32_t *toff;  // + g_lo
    const int2 *tbox;     // + g_lo
    __device__ __forceinline__ int off(int j) const { return __ldg(toff + j); }
    __device__ __forceinline__ int2 tiles(int j) const { return __ldg(tbox + j); }
};
__device__ __forceinline__ PairSourceGlobal gsrc_of(const int32_t *toff, const int2 *tbox, int g_lo) {
    return PairSourceGlobal{toff + g_lo, tbox + g_lo};
}
constexpr int PPT = 8;    // consecutive pairs per thread of k_view_pairs
constexpr int PAIR_THREADS = BIN_CHUNK / PPT;   // 512: twice the warps per block for the same shared memory — the
                                                // kernel is a chain of memory latencies, the warps are what hides them
constexpr int PAIR_PAD = 16;                    // padding of the pair buffers (pairs)

// PPT consecutive pairs from q0 on: one binary search for the owner of the first, then a walk — the tile id of the
// next pair of a box is the previous one + 1, or the start of the next tile row (no division per pair)
template <typename Source>
__device__ __forceinline__ void emit_pairs(const Source &src, int ng, int g_lo, int q0, int P, int ntx, int mask,
                                           int *__restrict__ s_cnt, int2 *__restrict__ out) {
    int lo = 0, hi = ng - 1;   // first j with off(j + 1) > q0
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (src.off(mid + 1) > q0) hi = mid;
        else lo = mid + 1;
    }
    int j = lo, end = src.off(j + 1);
    int2 box = src.tiles(j);
    int nx = box.y;
    const int local = q0 - src.off(j), row = local / nx;
    int col = local - row * nx;
    int t = ((box.x >> 16) + row) * ntx + (box.x & 0xffff) + col;
    int2 v[PPT];
#pragma unroll
    for (int i = 0; i < PPT; ++i) {
        const int q = q0 + i;
        v[i] = make_int2(-1, 0);
        if (q < P) {
            while (q >= end) {   // skips Gaussians without pairs too
                ++j;
                end = src.off(j + 1);
                if (q < end) {
                    box = src.tiles(j);
                    nx = box.y;
                    col = 0;
                    t = (box.x >> 16) * ntx + (box.x & 0xffff);
                }
            }
            v[i] = make_int2(t, g_lo + j);
            atomicAdd(&s_cnt[t & mask], 1);
            if (++col == nx) { col = 0; t += ntx - nx + 1; }
            else ++t;
        }
    }
#pragma unroll
    for (int i = 0; i < PPT; i += 2)   // 16-byte stores; the buffers are padded
        *reinterpret_cast<int4 *>(out + q0 + i) = make_int4(v[i].x, v[i].y, v[i + 1].x, v[i + 1].y);
}

// pair {tile, Gaussian}: one 8-byte word, so that a pair moves with ONE store (scattered stores cost per request,
// not per byte).  A block emits 4096 consecutive pairs of the Gaussian-major numbering:
// the Gaussian of its first pair was marked by k_view_pack (bstart), the block stages the pair offsets and tile boxes
// of its Gaussians in shared memory (coalesced loads, four in flight per thread), and every thread emits 16
// consecutive pairs from there — no dependent global loads per pair.  (A range holding more than PW_GAUSS Gaussians — long runs of Gaussians without any pair — is
// walked in global memory instead.)  Also counts the first digit of the block's pairs.
__global__ void __launch_bounds__(PAIR_THREADS, 3)
k_view_pairs(const int2 *__restrict__ tbox, const int32_t *__restrict__ toff, const int32_t *__restrict__ bstart,
             int64_t n, int ntx, int64_t cap, const unsigned int *__restrict__ hdr, int bits, int nb,
             int2 *__restrict__ out, int32_t *__restrict__ hist) {
    grid_dep_sync();
    extern __shared__ __align__(16) int s_dyn[];
    int *s_toff = s_dyn;                                               // [PW_GAUSS + 1]
    int2 *s_box = reinterpret_cast<int2 *>(s_dyn + PW_GAUSS + 2);      // [PW_GAUSS] {tx0 | ty0 << 16, nx}
    int *s_cnt = s_dyn + PW_GAUSS + 2 + 2 * PW_GAUSS;                  // [nbins]
    if (overflowed(hdr, cap)) return;
    const int nbins = 1 << bits;
    for (int d = threadIdx.x; d < nbins; d += PAIR_THREADS) s_cnt[d] = 0;
    const int P = __ldg(toff + n);
    const int64_t base = static_cast<int64_t>(blockIdx.x) * BIN_CHUNK;
    if (base < P) {   // block-uniform
        // Gaussians of the block's pairs: from the owner of its first pair to (at most) the owner of the next block's
        const int g_lo = __ldg(bstart + blockIdx.x);
        const int g_hi = base + BIN_CHUNK < P ? __ldg(bstart + blockIdx.x + 1) : static_cast<int>(n - 1);
        const int ng = g_hi - g_lo + 1;
        const int q0 = static_cast<int>(base) + threadIdx.x * PPT;
        if (ng <= PW_GAUSS) {
            const PairSourceGlobal gsrc = gsrc_of(toff, tbox, g_lo);
            for (int j0 = 0; j0 <= ng; j0 += 4 * PAIR_THREADS) {   // up to four loads of each kind in flight per thread
                int o[4];
                int2 bx[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int j = j0 + k * PAIR_THREADS + threadIdx.x;
                    if (j <= ng) o[k] = gsrc.off(j);
                    if (j < ng) bx[k] = gsrc.tiles(j);
                }
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int j = j0 + k * PAIR_THREADS + threadIdx.x;
                    if (j <= ng) s_toff[j] = o[k];
                    if (j < ng) s_box[j] = bx[k];
                }
            }
            __syncthreads();
            if (q0 < P) emit_pairs(PairSourceShared{s_toff, s_box}, ng, g_lo, q0, P, ntx, nbins - 1, s_cnt, out);
        } else {
            __syncthreads();   // s_cnt is zero
            if (q0 < P) emit_pairs(gsrc_of(toff, tbox, g_lo), ng, g_lo, q0, P, ntx, nbins - 1, s_cnt, out);
        }
    }
    __syncthreads();
    for (int d = threadIdx.x; d < nbins; d += PAIR_THREADS) hist[static_cast<int64_t>(d) * nb + blockIdx.x] = s_cnt[d];
}

// Exclusive scan of nbins (<= 512) ints by a block of 256 threads: s_out[d] = src[0] + .. + src[d-1], s_out[nbins] =
// the total.  src may be global or shared, and may be s_out itself.  Thread t owns the values [t K, t K + K).
__device__ __forceinline__ void scan_bins_256(const int *src, int nbins, int *s_out, int *s_wsum) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int K = (nbins + BIN_THREADS - 1) / BIN_THREADS;
    int v[(1 << BIN_MAX_BITS) / BIN_THREADS];
    int mysum = 0;
#pragma unroll
    for (int k = 0; k < (1 << BIN_MAX_BITS) / BIN_THREADS; ++k) {
        const int d = threadIdx.x * K + k;
        v[k] = (k < K && d < nbins) ? src[d] : 0;
        mysum += v[k];
    }
    int inc = mysum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int u = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += u;
    }
    __syncthreads();   // s_wsum free again, every src value read
    if (lane == 31) s_wsum[warp] = inc;
    __syncthreads();
    int pre = inc - mysum;
#pragma unroll
    for (int w = 0; w < BIN_WARPS; ++w)
        if (w < warp) pre += s_wsum[w];
#pragma unroll
    for (int k = 0; k < (1 << BIN_MAX_BITS) / BIN_THREADS; ++k) {
        const int d = threadIdx.x * K + k;
        if (k < K && d < nbins) s_out[d] = pre;
        pre += v[k];
    }
    if (threadIdx.x == BIN_THREADS - 1) s_out[nbins] = pre;
    __syncthreads();
}

// The [digit value][block] counts of a pass, scanned along the blocks: rowpre[d][b] = pairs with digit value d in the
// blocks before b, tot[d] = all of them.  One block per digit value, no dependency between blocks (the chained scan
// over the whole matrix spent 10 us per pass, most of it in its look-back; where a digit value's run starts in the
// output — the exclusive scan of tot — is 256..512 numbers that every consumer block scans for itself).
__global__ void __launch_bounds__(BIN_THREADS)
k_view_bin_rowscan(const int32_t *__restrict__ hist, int nb, int64_t cap, const unsigned int *__restrict__ hdr,
                   int32_t *__restrict__ rowpre, int32_t *__restrict__ tot) {
    grid_dep_sync();
    __shared__ int s_w[BIN_WARPS];
    if (overflowed(hdr, cap)) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int32_t *row = hist + static_cast<int64_t>(blockIdx.x) * nb;
    int32_t *out = rowpre + static_cast<int64_t>(blockIdx.x) * nb;
    int carry = 0;
    for (int c0 = 0; c0 < nb; c0 += BIN_THREADS * 8) {
        const int b0 = c0 + threadIdx.x * 8;
        int v[8], sum = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            v[i] = b0 + i < nb ? __ldg(row + b0 + i) : 0;
            sum += v[i];
        }
        int inc = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int u = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += u;
        }
        if (lane == 31) s_w[warp] = inc;
        __syncthreads();
        int pre = carry + inc - sum, all = 0;
#pragma unroll
        for (int w = 0; w < BIN_WARPS; ++w) {
            if (w < warp) pre += s_w[w];
            all += s_w[w];
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (b0 + i < nb) out[b0 + i] = pre;
            pre += v[i];
        }
        carry += all;
        __syncthreads();
    }
    if (threadIdx.x == 0) tot[blockIdx.x] = carry;
}

__global__ void __launch_bounds__(BIN_THREADS)
k_view_bin_hist(const int2 *__restrict__ in, int shift, int bits, int nb, int64_t cap,
                const unsigned int *__restrict__ hdr, int32_t *__restrict__ hist) {
    grid_dep_sync();
    __shared__ int s_cnt[1 << BIN_MAX_BITS];
    if (overflowed(hdr, cap)) return;
    const int nbins = 1 << bits;
    for (int d = threadIdx.x; d < nbins; d += BIN_THREADS) s_cnt[d] = 0;
    __syncthreads();
    const int P = static_cast<int>(reinterpret_cast<const unsigned long long *>(hdr)[H_P64]);
    const int64_t base = static_cast<int64_t>(blockIdx.x) * BIN_CHUNK;
#pragma unroll 4
    for (int r = 0; r < BIN_ROUNDS; ++r) {
        const int64_t i = base + r * BIN_THREADS + threadIdx.x;
        if (i < P) atomicAdd(&s_cnt[(__ldg(&in[i].x) >> shift) & (nbins - 1)], 1);
    }
    __syncthreads();
    for (int d = threadIdx.x; d < nbins; d += BIN_THREADS) hist[static_cast<int64_t>(d) * nb + blockIdx.x] = s_cnt[d];
}

// One digit pass over the block's 4096 pairs.  The block first sorts them by the digit in shared memory (stable:
// rank inside the warp's 512 pairs from MATCH.ANY + a running counter per warp and digit value, warps in order,
// digit values in order), then copies them out: the pairs of one digit value are one contiguous run of the output,
// so consecutive threads write consecutive addresses — a pair thrown straight at its final address costs one
// memory request per pair (measured: 75 us per pass at 1080p against ~12 this way).
// LAST: the final pass writes the Gaussian ids (the tile-ordered pair list) and the tile ids as two int arrays.
template <bool LAST, int BITS>   // BITS: ballots per rank (>= the digit's bits; 7, 8 or 9)
__global__ void __launch_bounds__(BIN_THREADS, 3)
k_view_bin_scatter(const int2 *__restrict__ in, const int32_t *__restrict__ rowpre, const int32_t *__restrict__ tot,
                   int shift, int bits, int nb, int64_t cap, const unsigned int *__restrict__ hdr,
                   int2 *__restrict__ out, int32_t *__restrict__ out_keys, int32_t *__restrict__ out_gid) {
    grid_dep_sync();
    // [BIN_WARPS][nbins] running count per warp and digit value, then its offset | [nbins + 1] start of the digit's run
    // inside the block | [nbins + 1] start of the digit's run in the output | [nbins] start of the block's run of the
    // digit in the output, relative to its place inside the block | the sorted pairs
    extern __shared__ __align__(16) int s_dyn[];
    if (overflowed(hdr, cap)) return;
    const int nbins = 1 << bits;
    int *s_wcnt = s_dyn, *s_lstart = s_dyn + BIN_WARPS * nbins, *s_dstart = s_lstart + nbins + 1;
    int *s_gbase = s_dstart + nbins + 1;
    int2 *s_stage = reinterpret_cast<int2 *>(s_gbase + nbins);
    __shared__ int s_wsum[BIN_WARPS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int P = static_cast<int>(reinterpret_cast<const unsigned long long *>(hdr)[H_P64]);
    const int64_t base = static_cast<int64_t>(blockIdx.x) * BIN_CHUNK;
    if (base >= P) return;
    const int count = static_cast<int>(min(static_cast<int64_t>(BIN_CHUNK), P - base));
    for (int d = threadIdx.x; d < BIN_WARPS * nbins; d += BIN_THREADS) s_wcnt[d] = 0;
    for (int d = threadIdx.x; d < nbins; d += BIN_THREADS) s_gbase[d] = __ldg(rowpre + static_cast<int64_t>(d) * nb + blockIdx.x);
    __syncthreads();
    int *mine = s_wcnt + warp * nbins;
    const unsigned lt = (1u << lane) - 1u;
    int2 it[BIN_ROUNDS];
    int rank[BIN_ROUNDS];
#pragma unroll
    for (int r = 0; r < BIN_ROUNDS; ++r) {
        const int i = warp * BIN_PER_WARP + r * 32 + lane;
        it[r] = make_int2(-1, 0);
        if (i < count) it[r] = __ldcs(in + base + i);
    }
#pragma unroll
    for (int r = 0; r < BIN_ROUNDS; ++r) {
        const bool valid = it[r].x >= 0;
        const int d = valid ? ((it[r].x >> shift) & (nbins - 1)) : 0;
        // the lanes holding the same digit value: one ballot per digit bit (MATCH.ANY gives the same mask but
        // iterates over the distinct values of the warp — measured 40 us per pass against the ballots' fixed cost)
        unsigned peers = __ballot_sync(0xffffffffu, valid);
#pragma unroll
        for (int b = 0; b < BITS; ++b) {   // bits above the digit are 0 in every lane: no-ops
            const bool bit = (d & (1 << b)) != 0;
            const unsigned has = __ballot_sync(0xffffffffu, bit);
            if (bit) peers &= has;
            else peers &= ~has;
        }
        if (!valid) peers = 1u << lane;
        const int before = __popc(peers & lt);
        int old = 0;
        if (valid && before == 0) {   // the lowest lane of the group
            old = mine[d];
            mine[d] = old + __popc(peers);
        }
        old = __shfl_sync(0xffffffffu, old, __ffs(peers) - 1);
        rank[r] = old + before;
        __syncwarp();
    }
    __syncthreads();
    // per digit value: the warps' offsets inside its run and the block's total
    for (int d = threadIdx.x; d < nbins; d += BIN_THREADS) {
        int run = 0;
#pragma unroll
        for (int w = 0; w < BIN_WARPS; ++w) {
            const int c = s_wcnt[w * nbins + d];
            s_wcnt[w * nbins + d] = run;
            run += c;
        }
        s_lstart[d] = run;
    }
    __syncthreads();
    scan_bins_256(s_lstart, nbins, s_lstart, s_wsum);   // -> where the digit value's run starts inside the block
    scan_bins_256(tot, nbins, s_dstart, s_wsum);        // -> where it starts in the output
    // s_gbase[d] becomes (start of the block's run of d in the output) - (start of d's run inside the block)
    for (int d = threadIdx.x; d < nbins; d += BIN_THREADS) s_gbase[d] += s_dstart[d] - s_lstart[d];
#pragma unroll
    for (int r = 0; r < BIN_ROUNDS; ++r) {
        if (it[r].x >= 0) {
            const int d = (it[r].x >> shift) & (nbins - 1);
            s_stage[s_lstart[d] + mine[d] + rank[r]] = it[r];
        }
    }
    __syncthreads();
#pragma unroll 4
    for (int r = 0; r < BIN_ROUNDS; ++r) {
        const int i = r * BIN_THREADS + threadIdx.x;
        if (i < count) {
            const int2 v = s_stage[i];
            const int d = (v.x >> shift) & (nbins - 1);
            const int pos = s_gbase[d] + i;
            if (LAST) {
                out_keys[pos] = v.x;
                out_gid[pos] = v.y;
            } else {
                out[pos] = v;
            }
        }
    }
}

// first index i in [lo, hi) with keys[i] >= t (hi if none)
__device__ __forceinline__ int lower_bound_in(const int32_t *__restrict__ keys, int lo, int hi, int t) {
    while (lo < hi) {
        const int mid = lo + ((hi - lo) >> 1);
        if (__ldg(keys + mid) < t) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}
// one thread per tile: its range in the tile-ordered pair list (a lower bound of its id, searched only inside the run
// of its last digit, whose bounds are the exclusive scan of the last pass's totals) and its pieces ("Work units")
__global__ void __launch_bounds__(BIN_THREADS)
k_view_tiles(const int32_t *__restrict__ keys, const int32_t *__restrict__ tot, int shift, int nbins, int ntiles,
             int piece, int64_t cap, unsigned int *__restrict__ hdr, int32_t *__restrict__ tcount,
             int32_t *__restrict__ tstart, int32_t *__restrict__ pextra, int32_t *__restrict__ ptile_x,
             int32_t *__restrict__ mlist) {
    grid_dep_sync();
    __shared__ int s_dstart[(1 << BIN_MAX_BITS) + 1];
    __shared__ int s_wsum[BIN_WARPS];
    if (overflowed(hdr, cap)) return;
    const int P = static_cast<int>(reinterpret_cast<const unsigned long long *>(hdr)[H_P64]);
    if (tot != nullptr) scan_bins_256(tot, nbins, s_dstart, s_wsum);
    auto lower = [&](int t) {
        if (tot == nullptr) return 0;   // no pairs at all
        const int d = t >> shift;
        return d >= nbins ? P : lower_bound_in(keys, s_dstart[d], s_dstart[d + 1], t);
    };
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const int tt = t < ntiles ? t : ntiles;
    const int lo = lower(tt);
    int hi = __shfl_down_sync(0xffffffffu, lo, 1);   // the next tile's lower bound ...
    if (lane == 31) hi = lower(tt + 1);              // ... which the last lane has to find itself
    if (t > ntiles) return;
    tstart[t] = lo;
    if (t == ntiles) return;
    const int len = hi - lo;
    tcount[t] = len;
    int x = -1;
    if (len > piece) {
        const int np = (len + piece - 1) / piece;
        x = static_cast<int>(atomicAdd(hdr + H_XPIECES, static_cast<unsigned>(np)));
        ptile_x[x] = -1;
        for (int k = 1; k < np; ++k) ptile_x[x + k] = t;
        mlist[atomicAdd(hdr + H_NMULTI, 1u)] = t;
    }
    pextra[t] = x;
}

// Work units of the walk kernels are PIECES: at most `piece` consecutive pairs of one tile's list (a multiple of
// 32).  Most tiles are one piece: work item t.  A tile with a longer list (bundled scene: thousands of Gaussians
// over one tile, ten times the mean) is cut into np pieces walked by different warps at the same time, each
// starting from the identity; it reserves np consecutive slots [x, x + np) of the extra-piece table (slot x stands
// for its first piece, which is work item t; slots x+1.. are work items of their own) and of the piece state:
//   forward : T_i = carry_p * Tlocal_i,  carry_p = prod of the aggregates of the earlier pieces; the colour is
//             linear in the carry, C = sum_p carry_p * C_p                                  (k_view_combine_fwd)
//   backward: U after the last element of piece p:  U_in(p) = Tagg_{p+1} U_in(p+1) + <dL/dI, C_{p+1}>  — the
//             affine map of a piece is made of the SAME two per-pixel quantities the forward already produced,
//             its aggregate and its colour, so the backward needs no pass of its own for them  (k_view_combine_bwd)
// piece state per slot (f32[192]): {Tagg[32], C0[32], C1[32], C2[32], carry[32], U_in[32]}.
constexpr int PIECE_STATE = 192;

// ---------------------------------------------------------------------------------------------------------------
// the walks
// ---------------------------------------------------------------------------------------------------------------
struct Piece {
    int t, np, slot;     // tile (-1: a hole in the table, nothing to do), pieces of that tile, piece-state slot
    int lo, hi;          // pair range of the piece inside the tile-ordered pair list
};
// next work item of this warp (dynamic ticket): the extra pieces first (they are full-length), then one per tile
__device__ __forceinline__ bool next_piece(unsigned int *ticket, const int32_t *__restrict__ tcount,
                                           const int32_t *__restrict__ tstart, const int32_t *__restrict__ pextra,
                                           const int32_t *__restrict__ ptile_x, int nx, int ntiles, int piece, int lane,
                                           Piece &w) {
    unsigned i = 0;
    if (lane == 0) i = atomicAdd(ticket, 1u);
    i = __shfl_sync(0xffffffffu, i, 0);
    if (i >= static_cast<unsigned>(nx + ntiles)) return false;
    int k = 0;
    if (i < static_cast<unsigned>(nx)) {
        w.t = __ldg(ptile_x + i);
        if (w.t < 0) return true;
        k = static_cast<int>(i) - __ldg(pextra + w.t);
    } else {
        w.t = static_cast<int>(i) - nx;
    }
    const int len = __ldg(tcount + w.t), base = __ldg(tstart + w.t);
    w.np = len <= piece ? 1 : (len + piece - 1) / piece;
    w.slot = __ldg(pextra + w.t) + k;
    w.lo = base + k * piece;
    w.hi = min(w.lo + piece, base + len);
    return true;
}

// one 32-byte sector in one instruction (LDG.256, sm_100)
__device__ __forceinline__ void ldg256(const void *p, int4 &u, int4 &v) {
    asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w), "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "l"(p)
                 : "memory");
}
struct RecRegs {
    int4 a, b, c, d;
};
__device__ __forceinline__ RecRegs load_rec(const int4 *__restrict__ rec, int g) {
    RecRegs r;
    ldg256(rec + 4 * static_cast<int64_t>(g), r.a, r.b);
    ldg256(rec + 4 * static_cast<int64_t>(g) + 2, r.c, r.d);
    return r;
}
// lanes of the tile at (x0, y0) inside the record's box: columns [xa, xb] of rows [ya, yb]
__device__ __forceinline__ uint32_t coverage_mask(const RecRegs &r, bool live, int x0, int y0) {
    const int sx = r.c.z, sy = r.c.w, ex = r.d.x, ey = r.d.y;
    const int xa = max(sx, x0) - x0, xb = min(ex, x0 + TW - 1) - x0;
    const int ya = max(sy, y0) - y0, yb = min(ey, y0 + TH - 1) - y0;
    uint32_t mask = 0;
    if (live && xa <= xb && ya <= yb) {
        const uint32_t xm = ((2u << xb) - 1u) & ~((1u << xa) - 1u);
        const uint32_t rows = (0x01010101u >> (8 * (TH - 1 - (yb - ya)))) << (8 * ya);
        mask = xm * rows;
    }
    return mask;
}
// Gaussian-major id of the pair (record's Gaussian with pair offset goff, tile (tx, ty)): where its partial goes
__device__ __forceinline__ int pair_id(const RecRegs &r, int tx, int ty) {
    const int tx0 = r.c.z >> TSX, ty0 = r.c.w >> TSY, nx = (r.d.x >> TSX) - tx0 + 1;
    return r.d.z + (ty - ty0) * nx + (tx - tx0);
}
// the record of Gaussian g copied into shared memory without passing through registers (LDGSTS, 4 x 16 bytes)
// quarter i of the record lands in dst[i * 32]: with dst = plane base + lane, the 32 lanes of a warp write
// consecutive 16-byte words per quarter (no bank conflicts)
__device__ __forceinline__ void copy_rec_async(int4 *dst, const int4 *__restrict__ rec, int g) {
    const unsigned d = static_cast<unsigned>(__cvta_generic_to_shared(dst));
    const int4 *src = rec + 4 * static_cast<int64_t>(g);
#pragma unroll
    for (int i = 0; i < 4; ++i)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(d + 512 * i), "l"(src + i) : "memory");
}
__device__ __forceinline__ void copy_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void copy_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ float i2f(int v) { return __int_as_float(v); }

struct PairEval {
    float d0, d1, gk, x;
};
// g = exp(-1/2 (r-m) Lambda (r-m)^T) (gs_model.py:495); x = 1 - o g (:533-535).  a = {mx, my, l00', l01'},
// (l10', l11', o): Lambda' = EXP2_SCALE Lambda, so g = 2^(d Lambda' d^T).
__device__ __forceinline__ PairEval eval_pair(const float4 &a, float l10, float l11, float o, float px, float py) {
    PairEval e;
    e.d0 = px - a.x;
    e.d1 = py - a.y;
    const float X0 = e.d0 * a.z + e.d1 * l10;
    const float X1 = e.d0 * a.w + e.d1 * l11;
    const float p2 = X0 * e.d0 + X1 * e.d1;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.gk) : "f"(p2));
    e.x = 1.0f - o * e.gk;
    return e;
}

// ---- forward ----
// What the forward walk needs of a pair, staged in shared memory by the lane that loaded it (three broadcast
// LDS.128 per pair): {mx, my, l00, l01} {l10, l11, o, coverage mask} {l0, l1, l2, -}
// (component planes: lane l writes a[l], b[l], c[l] — consecutive 16-byte words, no bank conflicts; an
// array of 48-byte structs costs a 12-way conflict per staging store)
struct FSlot {
    float4 a[32], b[32], c[32];
};
constexpr int TILE_WARPS = 8;

template <bool KEEP>
__global__ void __launch_bounds__(TILE_WARPS * 32, 4)
k_view_render(const int32_t *__restrict__ tcount, const int32_t *__restrict__ tstart,
              const int32_t *__restrict__ pextra, const int32_t *__restrict__ ptile_x,
              const int32_t *__restrict__ pgid, const int4 *__restrict__ rec, unsigned int *__restrict__ hdr,
              int64_t cap, int piece, int ntx, int ntiles, int W, int H, float *__restrict__ image,
              float *__restrict__ tck, float *__restrict__ pstate) {
    grid_dep_sync();
    __shared__ FSlot slots[TILE_WARPS];
    if (overflowed(hdr, cap)) return;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const unsigned lanebit = 1u << lane;
    FSlot &sl = slots[wib];
    const int nx = static_cast<int>(hdr[H_XPIECES]);
    Piece w;
    while (next_piece(hdr + H_TICKET_FWD, tcount, tstart, pextra, ptile_x, nx, ntiles, piece, lane, w)) {
        if (w.t < 0) continue;   // warp-uniform; only __syncwarp inside
        const int t = w.t;
        const int ty = t / ntx, tx = t - ty * ntx;
        const int x0 = tx << TSX, y0 = ty << TSY;
        const int ix = x0 + (lane & (TW - 1)), iy = y0 + (lane >> TSX);
        const float px = static_cast<float>(ix), py = static_cast<float>(iy);
        const int lo = w.lo, hi = w.hi;
        float T = 1.0f, c0 = 0.f, c1 = 0.f, c2 = 0.f;   // T: local to the piece (its carry is applied afterwards)
        // software pipeline over batches of 32 pairs: ids two batches ahead, records one batch ahead
        int g1 = 0;
        RecRegs r = {};
        if (lo + lane < hi) r = load_rec(rec, __ldg(pgid + lo + lane));
        if (lo + 32 + lane < hi) g1 = __ldg(pgid + lo + 32 + lane);
        for (int b = lo; b < hi; b += 32) {
            {
                const uint32_t mask = coverage_mask(r, b + lane < hi, x0, y0);
                sl.a[lane] = make_float4(i2f(r.a.x), i2f(r.a.y), i2f(r.a.z), i2f(r.a.w));
                sl.b[lane] = make_float4(i2f(r.b.x), i2f(r.b.y), i2f(r.b.z), __uint_as_float(mask));
                sl.c[lane] = make_float4(i2f(r.b.w), i2f(r.c.x), i2f(r.c.y), 0.0f);
            }
            int g2 = 0;
            if (b + 64 + lane < hi) g2 = __ldg(pgid + b + 64 + lane);
            if (b + 32 + lane < hi) r = load_rec(rec, g1);
            g1 = g2;
            __syncwarp();
            const int m = min(hi - b, 32);
            // checkpoint rows: row (b >> 3) + t for the 8 pairs from b on (b - tile start is a multiple of 8): the
            // rows of a tile are distinct and below the first row of the next tile
            float *ck = tck + ((static_cast<int64_t>(b) >> SUB_SHIFT) + t) * 32 + lane;
#pragma unroll 4
            for (int k = 0; k < m; ++k) {
                const float4 A = sl.a[k], B = sl.b[k], C = sl.c[k];
                const bool cov = (__float_as_uint(B.w) & lanebit) != 0u;   // one LOP3 into a predicate
                const PairEval e = eval_pair(A, B.x, B.y, B.z, px, py);
                if (KEEP && (k & (SUB - 1)) == 0) __stcs(ck + (k >> SUB_SHIFT) * 32, T);   // T checkpoint for the backward
                const float tin = T * e.x;
                // branch-free: an element outside the box, or dead (inclusive product 0, gs_model.py:575-578), adds 0
                const bool alive = cov & (tin != 0.0f);
                const float ta = alive ? T * (1.0f - e.x) : 0.0f;
                c0 = fmaf(ta, C.x, c0);
                c1 = fmaf(ta, C.y, c1);
                c2 = fmaf(ta, C.z, c2);
                T = cov ? tin : T;
            }
            __syncwarp();
        }
        if (w.np > 1) {
            float *ps = pstate + static_cast<int64_t>(w.slot) * PIECE_STATE + lane;
            ps[0] = T; ps[32] = c0; ps[64] = c1; ps[96] = c2;
        } else if (ix <= W && iy <= H) {
            float *p = image + 3 * (static_cast<int64_t>(iy) * (W + 1) + ix);
            p[0] = c0; p[1] = c1; p[2] = c2;
        }
    }
}

// tiles of several pieces: the carry of every piece (kept for the backward), and the pixel's colour
__global__ void __launch_bounds__(256)
k_view_combine_fwd(const int32_t *__restrict__ tcount, const int32_t *__restrict__ pextra,
                   const int32_t *__restrict__ mlist, const unsigned int *__restrict__ hdr, int64_t cap, int piece,
                   int ntx, int W, int H, float *__restrict__ pstate, float *__restrict__ image) {
    grid_dep_sync();
    if (overflowed(hdr, cap)) return;
    const int lane = threadIdx.x & 31;
    const unsigned int nm = hdr[H_NMULTI];
    for (unsigned int i = blockIdx.x * 8 + (threadIdx.x >> 5); i < nm; i += gridDim.x * 8) {
        const int t = mlist[i];
        const int np = (__ldg(tcount + t) + piece - 1) / piece, s0 = __ldg(pextra + t);
        float carry = 1.0f, c0 = 0.f, c1 = 0.f, c2 = 0.f;
        float *ps = pstate + static_cast<int64_t>(s0) * PIECE_STATE + lane;
        float a0 = ps[0], a1 = ps[32], a2 = ps[64], a3 = ps[96];   // the next piece's state is loaded a piece ahead
        for (int k = 0; k < np; ++k) {
            const float t0 = a0, t1 = a1, t2 = a2, t3 = a3;
            if (k + 1 < np) {
                const float *pn = ps + PIECE_STATE;
                a0 = pn[0]; a1 = pn[32]; a2 = pn[64]; a3 = pn[96];
            }
            ps[128] = carry;
            c0 = fmaf(carry, t1, c0);
            c1 = fmaf(carry, t2, c1);
            c2 = fmaf(carry, t3, c2);
            carry *= t0;
            ps += PIECE_STATE;
        }
        const int ty = t / ntx, tx = t - ty * ntx;
        const int ix = (tx << TSX) + (lane & (TW - 1)), iy = (ty << TSY) + (lane >> TSX);
        if (ix <= W && iy <= H) {
            float *p = image + 3 * (static_cast<int64_t>(iy) * (W + 1) + ix);
            p[0] = c0; p[1] = c1; p[2] = c2;
        }
    }
}

// tiles of several pieces: U after the last element of every piece, from the pieces behind it
__global__ void __launch_bounds__(256)
k_view_combine_bwd(const int32_t *__restrict__ tcount, const int32_t *__restrict__ pextra,
                   const int32_t *__restrict__ mlist, const unsigned int *__restrict__ hdr, int64_t cap, int piece,
                   const float *__restrict__ gimg, int ntx, int W, int H, float *__restrict__ pstate) {
    grid_dep_sync();
    if (overflowed(hdr, cap)) return;   // a view that did not fit its arena was not rendered: nothing to walk back
    const int lane = threadIdx.x & 31;
    const unsigned int nm = hdr[H_NMULTI];
    for (unsigned int i = blockIdx.x * 8 + (threadIdx.x >> 5); i < nm; i += gridDim.x * 8) {
        const int t = mlist[i];
        const int np = (__ldg(tcount + t) + piece - 1) / piece, s0 = __ldg(pextra + t);
        const int ty = t / ntx, tx = t - ty * ntx;
        const int ix = (tx << TSX) + (lane & (TW - 1)), iy = (ty << TSY) + (lane >> TSX);
        float pg0 = 0.f, pg1 = 0.f, pg2 = 0.f;
        if (ix <= W && iy <= H) {
            const float *p = gimg + 3 * (static_cast<int64_t>(iy) * (W + 1) + ix);
            pg0 = __ldg(p); pg1 = __ldg(p + 1); pg2 = __ldg(p + 2);
        }
        float U = 0.0f;
        float *ps = pstate + static_cast<int64_t>(s0 + np - 1) * PIECE_STATE + lane;
        float a0 = ps[0], a1 = ps[32], a2 = ps[64], a3 = ps[96];   // the next piece's state is loaded a piece ahead
        for (int k = np - 1; k >= 0; --k) {
            const float t0 = a0, t1 = a1, t2 = a2, t3 = a3;
            if (k > 0) {
                const float *pn = ps - PIECE_STATE;
                a0 = pn[0]; a1 = pn[32]; a2 = pn[64]; a3 = pn[96];
            }
            ps[160] = U;
            U = fmaf(t0, U, pg0 * t1 + pg1 * t2 + pg2 * t3);
            ps -= PIECE_STATE;
        }
    }
}

// ---- backward ----
// Staged per pair (64 bytes, broadcast LDS.128: two in the recompute sweep, one in the reverse walk, one per
// pair in the moment phase):
//   f0 {mx, my, l00, l01}  f1 {l10, l11, o, coverage mask}  |  r0 {o, l0, l1, l2}  r1 {mx, my, pair id, -}
// Component planes (lane l stages f0[l], f1[l], r0[l], r1[l]: consecutive 16-byte words, conflict-free; as an
// array of 64-byte structs every staging store was a 16-way bank conflict).
struct BSlot {
    float4 f0[32], f1[32], r0[32], r1[32];
};
constexpr int BWD_WARPS = 8;
// exchange buffer of a warp: for each of the SUB pairs of a sub-batch, c[32] and dv[32] of its pixels.  The pair
// stride of 68 words keeps the moment phase's LDS.128 (lanes = 8 pairs x 4 tile rows) free of bank conflicts.
constexpr int XCH_STRIDE = 68;
constexpr int BWD_SMEM_PER_WARP = static_cast<int>(sizeof(BSlot)) + 32 * 64 + SUB * XCH_STRIDE * 4;
constexpr int BWD_SMEM = BWD_WARPS * BWD_SMEM_PER_WARP;

// partial[q] = {sum c, sum d, sum c d0, sum c d1, sum c d0 d0, sum c d0 d1, sum c d1 d1, -} over the pixels of
// pair q, with c = g * dL/dalpha and d = T alpha <dL/dI, l>: the moments from which k_view_reduce forms the
// reference's per-element gradients (gs_model.py:733-766) once per GAUSSIAN instead of once per pixel.
// Nothing of the forward walk is read back but one T per lane and 8 pairs.  A sub-batch of 8 pairs goes through
// three phases:
//   (1) sweep, lane = pixel   : the 8 pairs re-evaluated forward from the checkpoint; T and g stay in registers
//   (2) reverse, lane = pixel : U, dL/dalpha -> c and d of every (pair, pixel) dropped into shared memory
//   (3) moments, lane = (pair, tile row): a lane reads the 8 pixels of its row and pair, forms the row's sums of
//       c {1, d0, d0 d0} and of d in registers (d1 is constant along a row), and the 4 rows of a pair are added
//       by a two-level halving exchange that leaves lane r with terms {2r, 2r+1}: one 8-byte store per lane.
// Against summing 8 values over the 32 lanes of every pair with shuffles (the previous version: 32 of its 96
// instructions per pair), the cross-lane traffic per pair is 2 STS + 4/8 LDS.128 + 6/8 SHFL.  The order of every
// sum is fixed: bitwise reproducible.
template <bool FULL>
__device__ __forceinline__ void backward_sub_batch(const BSlot &sl, float *__restrict__ xch, int s0, int ms,
                                                   int lane, float px, float py, float pg0, float pg1, float pg2,
                                                   float T, float &U, int x0, int y0, float *__restrict__ partial) {
    // (1) recompute sweep, forward: T and g of every (pair, lane) of the sub-batch; coverage bits kept per lane
    float Tk[SUB], Gk[SUB];
    unsigned covbits = 0;
    const unsigned lanebit = 1u << lane;
#pragma unroll
    for (int k = 0; k < SUB; ++k) {
        Tk[k] = 0.0f; Gk[k] = 0.0f;
        if (FULL || k < ms) {   // warp-uniform
            const float4 A = sl.f0[s0 + k], B = sl.f1[s0 + k];
            const bool cov = (__float_as_uint(B.w) & lanebit) != 0u;
            const PairEval e = eval_pair(A, B.x, B.y, B.z, px, py);
            Tk[k] = T;
            Gk[k] = e.gk;
            T = cov ? T * e.x : T;
            covbits |= cov ? (1u << k) : 0u;
        }
    }
    // (2) reverse walk
#pragma unroll
    for (int k = SUB - 1; k >= 0; --k) {
        if (FULL || k < ms) {   // warp-uniform
            const float4 R0 = sl.r0[s0 + k];
            const bool cov = (covbits >> k) & 1u;
            const float Tt = Tk[k], gk = Gk[k];
            const float alpha = R0.x * gk, x = 1.0f - alpha;
            const bool alive = cov & (Tt * x != 0.0f);
            const float pgl = pg0 * R0.y + pg1 * R0.z + pg2 * R0.w;
            const float wv = alive ? alpha * pgl : 0.0f;
            const float dalpha = alive ? Tt * (pgl - U) : 0.0f;
            U = cov ? fmaf(x, U, wv) : U;   // U_{i-1} = w_i + x_i U_i
            xch[k * XCH_STRIDE + lane] = gk * dalpha;
            xch[k * XCH_STRIDE + 32 + lane] = Tt * wv;
        }
    }
    __syncwarp();
    // (3) moments: lane = (pair p of the sub-batch, row r of the tile)
    {
        const int p = lane >> 2, r = lane & 3;
        const float4 R1 = sl.r1[s0 + p];
        const float *xc = xch + p * XCH_STRIDE + r * TW;
        const float4 ca = *reinterpret_cast<const float4 *>(xc), cb = *reinterpret_cast<const float4 *>(xc + 4);
        const float4 da = *reinterpret_cast<const float4 *>(xc + 32), db = *reinterpret_cast<const float4 *>(xc + 36);
        const float c[8] = {ca.x, ca.y, ca.z, ca.w, cb.x, cb.y, cb.z, cb.w};
        const float d1 = static_cast<float>(y0 + r) - R1.y;
        float R_0 = 0.0f, R_1 = 0.0f, R_2 = 0.0f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float d0 = static_cast<float>(x0 + i) - R1.x;
            const float cd0 = c[i] * d0;
            R_0 += c[i];
            R_1 += cd0;
            R_2 = fmaf(cd0, d0, R_2);
        }
        const float Rd = ((da.x + da.y) + (da.z + da.w)) + ((db.x + db.y) + (db.z + db.w));
        const float d1R0 = d1 * R_0;
        // terms 0..7 of this row; rows added by halving: after xor 2 a lane keeps terms 4*bit1..4*bit1+3, after
        // xor 1 terms 2r, 2r+1
        const float v[8] = {R_0, Rd, R_1, d1R0, R_2, d1 * R_1, d1 * d1R0, 0.0f};
        const unsigned F = 0xffffffffu;
        float k2[4];
        {
            const bool h = r & 2;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float keep = h ? v[4 + i] : v[i], send = h ? v[i] : v[4 + i];
                k2[i] = keep + __shfl_xor_sync(F, send, 2);
            }
        }
        float2 out;
        {
            const bool h = r & 1;
            out.x = (h ? k2[2] : k2[0]) + __shfl_xor_sync(F, h ? k2[0] : k2[2], 1);
            out.y = (h ? k2[3] : k2[1]) + __shfl_xor_sync(F, h ? k2[1] : k2[3], 1);
        }
        if (FULL || p < ms)
            *reinterpret_cast<float2 *>(partial + static_cast<int64_t>(__float_as_int(R1.z)) * 8 + 2 * r) = out;
    }
    __syncwarp();
}

__global__ void __launch_bounds__(BWD_WARPS * 32, 3)
k_view_backward(const int32_t *__restrict__ tcount, const int32_t *__restrict__ tstart,
                const int32_t *__restrict__ pextra, const int32_t *__restrict__ ptile_x,
                const int32_t *__restrict__ pgid, const int4 *__restrict__ rec, unsigned int *__restrict__ hdr,
                int64_t cap, int piece, const float *__restrict__ tck, const float *__restrict__ pstate,
                const float *__restrict__ gimg, int ntx, int ntiles, int W, int H, float *__restrict__ partial) {
    grid_dep_sync();
    if (overflowed(hdr, cap)) return;
    // dynamic shared memory, per warp: 32 staged slots | 32 raw records (the next batch's arrive here by cp.async;
    // a lane reads back and overwrites only its own record, so one buffer is enough) | the exchange buffer
    extern __shared__ __align__(16) unsigned char bwd_smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    unsigned char *wbase = bwd_smem + static_cast<size_t>(wib) * BWD_SMEM_PER_WARP;
    BSlot &sl = *reinterpret_cast<BSlot *>(wbase);
    int4 *raw = reinterpret_cast<int4 *>(wbase + sizeof(BSlot));          // [4 quarters][32 lanes]
    float *xch = reinterpret_cast<float *>(wbase + sizeof(BSlot) + 32 * 64);
    const int nx = static_cast<int>(hdr[H_XPIECES]);
    Piece w;
    while (next_piece(hdr + H_TICKET_BWD, tcount, tstart, pextra, ptile_x, nx, ntiles, piece, lane, w)) {
        if (w.t < 0) continue;
        const int t = w.t;
        const int lo = w.lo, hi = w.hi;
        if (lo >= hi) continue;
        const int ty = t / ntx, tx = t - ty * ntx;
        const int x0 = tx << TSX, y0 = ty << TSY;
        const int ix = x0 + (lane & (TW - 1)), iy = y0 + (lane >> TSX);
        const float px = static_cast<float>(ix), py = static_cast<float>(iy);
        float pg0 = 0.f, pg1 = 0.f, pg2 = 0.f;
        if (ix <= W && iy <= H) {
            const float *p = gimg + 3 * (static_cast<int64_t>(iy) * (W + 1) + ix);
            pg0 = __ldg(p); pg1 = __ldg(p + 1); pg2 = __ldg(p + 2);
        }
        // a piece of a longer list starts from the carries the combine kernels left: T = carry * Tlocal, U = U_in
        float carry = 1.0f, U = 0.0f;
        if (w.np > 1) {
            const float *ps = pstate + static_cast<int64_t>(w.slot) * PIECE_STATE + lane;
            carry = __ldg(ps + 128);
            U = __ldg(ps + 160);
        }
        // batches of 32 pairs, aligned to the piece start, from the last one down; ids two batches ahead, records
        // one batch ahead — straight into shared memory, so that no register holds them during the walk
        const int last = lo + ((hi - lo - 1) & ~31);
        int g1 = 0;
        if (last + lane < hi) copy_rec_async(raw + lane, rec, __ldg(pgid + last + lane));
        copy_commit();
        if (last > lo) g1 = __ldg(pgid + last - 32 + lane);
        const float *ckl = tck + static_cast<int64_t>(t) * 32 + lane;
        // checkpoint of the sub-batch walked next, loaded one sub-batch ahead
        float tnext = __ldcs(ckl + (static_cast<int64_t>(last + ((hi - last - 1) & ~(SUB - 1))) >> SUB_SHIFT) * 32);
        for (int bb = last; bb >= lo; bb -= 32) {
            const int m = min(hi - bb, 32);
            copy_wait_all();   // a lane reads back only what it copied itself
            {
                RecRegs r = {};
                if (lane < m) {
                    const int4 *rr = raw + lane;
                    r.a = rr[0]; r.b = rr[32]; r.c = rr[64]; r.d = rr[96];
                }
                const uint32_t mask = coverage_mask(r, lane < m, x0, y0);
                const int q = pair_id(r, tx, ty);
                sl.f0[lane] = make_float4(i2f(r.a.x), i2f(r.a.y), i2f(r.a.z), i2f(r.a.w));
                sl.f1[lane] = make_float4(i2f(r.b.x), i2f(r.b.y), i2f(r.b.z), __uint_as_float(mask));
                sl.r0[lane] = make_float4(i2f(r.b.z), i2f(r.b.w), i2f(r.c.x), i2f(r.c.y));
                sl.r1[lane] = make_float4(i2f(r.a.x), i2f(r.a.y), i2f(q), 0.0f);
            }
            if (bb - 32 >= lo) copy_rec_async(raw + lane, rec, g1);
            copy_commit();
            g1 = (bb - 64 >= lo) ? __ldg(pgid + bb - 64 + lane) : 0;
            __syncwarp();
            for (int sb = (m - 1) >> SUB_SHIFT; sb >= 0; --sb) {   // sub-batches of SUB = 8 pairs, the last one first
                const int s0 = sb * SUB, ms = min(m - s0, SUB);
                const float T = carry * tnext;
                {   // the checkpoint after this one in walking order: the previous sub-batch of the piece
                    const int nb = bb + s0 - SUB;
                    if (nb >= lo) tnext = __ldcs(ckl + (static_cast<int64_t>(nb) >> SUB_SHIFT) * 32);
                }
                if (ms == SUB) backward_sub_batch<true>(sl, xch, s0, ms, lane, px, py, pg0, pg1, pg2, T, U, x0, y0, partial);
                else backward_sub_batch<false>(sl, xch, s0, ms, lane, px, py, pg0, pg1, pg2, T, U, x0, y0, partial);
            }
            __syncwarp();
        }
    }
}

// ---- per-Gaussian sums ----
constexpr int RED_BIG = 64;  // Gaussians with more pairs than this go to k_view_reduce_big (one block each)

// Where the four gradients of a view go.  index == nullptr: row g of the view's own arrays (the autograd contract,
// gs_model.py:820).  index != nullptr: row index[g] of parameter-sized arrays, ADDED to what is there — the
// multi-view step: the reference's boolean-mask selection of a view's Gaussians (gs_model.py:405-413) turns, in
// autograd's backward, into exactly this scatter-add of the view's gradients into the parameters' .grad.  A view's
// index holds distinct rows and views follow each other on the stream, so the sums have a fixed order.
struct GradOut {
    float *g_mean, *g_lam, *g_opac, *g_l;
    const int32_t *index;
};

// S[0..6] = the seven sums of a Gaussian's partials; what the reference's per-element formulas
// (gs_model.py:733-766) give when summed over the Gaussian's pixels, with X = d Lambda, coef = o g dalpha:
//   d_opacity = sum g dalpha                      = S0                    (the d/o term of :739 cancels in dalpha)
//   d_l[k]    = (sum d) / l[k]                    = S1 / l[k]             (:763-766)
//   d_mean    = sum coef X                        = o Lambda^T (S2, S3)   (:743-750)
//   d_Lambda  = -1/2 sum coef d^T d               = -o/2 [[S4, S5], [S5, S6]]   (:753-760)
__device__ __forceinline__ void store_sums(const float (&S)[7], int64_t g, const int4 *__restrict__ rec,
                                           const GradOut &out) {
    int4 ra, rb;
    ldg256(rec + 4 * g, ra, rb);
    const int4 rc = __ldg(rec + 4 * g + 2);
    const float o = i2f(rb.z);
    // the records hold Lambda' = EXP2_SCALE Lambda
    const float2 gm = make_float2(EXP2_UNSCALE * o * (i2f(ra.z) * S[2] + i2f(rb.x) * S[3]),
                                  EXP2_UNSCALE * o * (i2f(ra.w) * S[2] + i2f(rb.y) * S[3]));
    const float h = -0.5f * o;
    const float4 gL = make_float4(h * S[4], h * S[5], h * S[5], h * S[6]);
    const float gl0 = S[1] / i2f(rb.w), gl1 = S[1] / i2f(rc.x), gl2 = S[1] / i2f(rc.y);
    if (out.index == nullptr) {
        reinterpret_cast<float2 *>(out.g_mean)[g] = gm;
        reinterpret_cast<float4 *>(out.g_lam)[g] = gL;
        out.g_opac[g] = S[0];
        out.g_l[3 * g] = gl0; out.g_l[3 * g + 1] = gl1; out.g_l[3 * g + 2] = gl2;
    } else {
        const int64_t j = __ldg(out.index + g);
        float2 *pm = reinterpret_cast<float2 *>(out.g_mean) + j;
        float4 *pL = reinterpret_cast<float4 *>(out.g_lam) + j;
        const float2 m0 = *pm;
        const float4 L0 = *pL;
        *pm = make_float2(m0.x + gm.x, m0.y + gm.y);
        *pL = make_float4(L0.x + gL.x, L0.y + gL.y, L0.z + gL.z, L0.w + gL.w);
        out.g_opac[j] += S[0];
        out.g_l[3 * j] += gl0; out.g_l[3 * j + 1] += gl1; out.g_l[3 * j + 2] += gl2;
    }
}

__device__ __forceinline__ void add_partial(float (&S)[7], const int4 &u, const int4 &v) {
    S[0] += i2f(u.x); S[1] += i2f(u.y); S[2] += i2f(u.z); S[3] += i2f(u.w);
    S[4] += i2f(v.x); S[5] += i2f(v.y); S[6] += i2f(v.z);
}

// one thread per Gaussian: its pairs' partials (32 bytes each, contiguous in pair order — consecutive lanes read
// consecutive ranges) are added in pair order, four pairs = four LDG.256 in flight at a time.  Gaussians with more
// than RED_BIG pairs (boxes of thousands of pixels, bundled scene) are appended to `big` instead; the order of
// that list does not enter any float sum.
__global__ void __launch_bounds__(256)
k_view_reduce(const float *__restrict__ partial, const int32_t *__restrict__ toff, const int4 *__restrict__ rec,
              int64_t n, GradOut out, unsigned int *__restrict__ hdr, int64_t cap, int32_t *__restrict__ big) {
    grid_dep_sync();
    if (overflowed(hdr, cap)) return;
    const int4 zero = make_int4(0, 0, 0, 0);
    // grid-stride: resident blocks that live for the whole kernel keep the memory system fuller than 3 500 blocks
    // of a few microseconds each (ncu: 50 % of the warp slots active, 2.8 TB/s)
    for (int64_t g = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; g < n;
         g += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int b = __ldg(toff + g), e = __ldg(toff + g + 1);
        if (e - b > RED_BIG) {
            big[atomicAdd(hdr + H_NBIG, 1u)] = static_cast<int32_t>(g);
            continue;
        }
        float S[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        for (int q = b; q < e; q += 4) {
            int4 u[4], v[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                u[i] = zero; v[i] = zero;
                if (q + i < e) ldg256(partial + static_cast<int64_t>(q + i) * 8, u[i], v[i]);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if (q + i < e) add_partial(S, u[i], v[i]);
        }
        store_sums(S, g, rec, out);
    }
}

// Big Gaussians.  Up to RED_HUGE pairs: one warp each (ticket) — lane l adds the partials of pairs l, l+32, ... (one
// LDG.256 per pair, four in flight) and the 32 lane sums are added by a butterfly.  More pairs than that (the
// bundled scene has boxes of 5 000 tiles): a whole block each, thread t takes pairs t, t+256, ..., the warps'
// butterfly sums are added in warp order.  Fixed orders: bitwise reproducible.
constexpr int RED_HUGE = 1024;

template <int STRIDE>
__device__ __forceinline__ void sum_partials_strided(const float *__restrict__ partial, int first, int e, float (&S)[7]) {
    const int4 zero = make_int4(0, 0, 0, 0);
    for (int q = first; q < e; q += 4 * STRIDE) {
        int4 u[4], v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            u[k] = zero; v[k] = zero;
            if (q + STRIDE * k < e) ldg256(partial + static_cast<int64_t>(q + STRIDE * k) * 8, u[k], v[k]);
        }
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (q + STRIDE * k < e) add_partial(S, u[k], v[k]);
    }
#pragma unroll
    for (int c = 0; c < 7; ++c) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) S[c] += __shfl_xor_sync(0xffffffffu, S[c], o);
    }
}

__global__ void __launch_bounds__(256)
k_view_reduce_big(const float *__restrict__ partial, const int32_t *__restrict__ toff, const int4 *__restrict__ rec,
                  GradOut out, unsigned int *__restrict__ hdr, int64_t cap, const int32_t *__restrict__ big) {
    grid_dep_sync();
    __shared__ float s_w[8][8];
    if (overflowed(hdr, cap)) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned int nb = hdr[H_NBIG];
    // (1) the huge ones, a block each
    for (unsigned int i = blockIdx.x; i < nb; i += gridDim.x) {
        const int64_t g = __ldg(big + i);
        const int b = __ldg(toff + g), e = __ldg(toff + g + 1);
        if (e - b <= RED_HUGE) continue;   // block-uniform
        float S[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        sum_partials_strided<256>(partial, b + threadIdx.x, e, S);
        if (lane == 0) {
#pragma unroll
            for (int c = 0; c < 7; ++c) s_w[warp][c] = S[c];
        }
        __syncthreads();
        if (threadIdx.x == 0) {
#pragma unroll
            for (int c = 0; c < 7; ++c) {
                float t = 0.0f;
#pragma unroll
                for (int w = 0; w < 8; ++w) t += s_w[w][c];
                S[c] = t;
            }
            store_sums(S, g, rec, out);
        }
        __syncthreads();
    }
    // (2) the others, a warp each
    while (true) {
        unsigned i = 0;
        if (lane == 0) i = atomicAdd(hdr + H_TICKET_RED, 1u);
        i = __shfl_sync(0xffffffffu, i, 0);
        if (i >= nb) return;
        const int64_t g = __ldg(big + i);
        const int b = __ldg(toff + g), e = __ldg(toff + g + 1);
        if (e - b > RED_HUGE) continue;
        float S[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        sum_partials_strided<32>(partial, b + lane, e, S);
        if (lane == 0) store_sums(S, g, rec, out);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------
int g_piece = 128;  // pairs per piece (gcp_tile_set_piece_pairs)

inline bool bad_image(int W, int H) { return W < 0 || H < 0 || W >= 32768 || H >= 32768; }
inline int tiles_x(int W) { return (W + TW) >> TSX; }  // ceil((W+1)/TW): pixels 0..W inclusive (gs_model.py:505)
inline int tiles_y(int H) { return (H + TH) >> TSY; }

// plan arena: everything whose size is known from (n, W, H) alone.  [0, zero_bytes) is cleared by gcp_view_plan.
struct PlanLayout {
    size_t hdr, desc1, tcount, zero_bytes, cnt, tbox, toff, tstart, pextra, mlist, big, rec, total;
    unsigned nb1;
};
PlanLayout plan_layout(int64_t n, int ntiles) {
    PlanLayout L;
    L.nb1 = blocks_for(n, SCAN_TILE);
    size_t o = 0;
    auto take = [&](size_t bytes) { const size_t at = o; o += align256(bytes); return at; };
    L.hdr = take(HDR_WORDS * 4);
    L.desc1 = take(static_cast<size_t>(L.nb1) * 8);
    L.tcount = take(static_cast<size_t>(ntiles) * 4);
    L.zero_bytes = o;
    L.cnt = take(static_cast<size_t>(n > 0 ? n : 1) * 4);
    L.tbox = take(static_cast<size_t>(n > 0 ? n : 1) * 8);
    L.toff = take(static_cast<size_t>(n + 1) * 4);
    L.tstart = take(static_cast<size_t>(ntiles + 1) * 4);
    L.pextra = take(static_cast<size_t>(ntiles) * 4);
    L.mlist = take(static_cast<size_t>(ntiles) * 4);
    L.big = take(static_cast<size_t>(n > 0 ? n : 1) * 4);
    L.rec = take(static_cast<size_t>(n > 0 ? n : 1) * 64);
    L.total = o;
    return L;
}
// pair arena: everything sized by the pair capacity.  What the binning needs (dead once the pair list is built) shares
// the space of the backward's partials (32 B per pair): the two {tile, Gaussian} buffers of the radix passes (8 B per
// pair each), the [digit value][block] counts and their row scans, the digit totals and the block starts.
struct BinPlan {
    int passes, bits, nb;        // digit passes, bits per digit, blocks of BIN_CHUNK pairs
};
BinPlan bin_plan(int64_t cap, int ntiles) {
    int total = 1;
    while ((int64_t(1) << total) < ntiles) ++total;
    BinPlan b;
    b.passes = (total + BIN_MAX_BITS - 1) / BIN_MAX_BITS;   // <= BIN_MAX_PASSES: tiles < 2^30 (bad_image)
    b.bits = (total + b.passes - 1) / b.passes;
    b.nb = static_cast<int>(blocks_for(cap, BIN_CHUNK));
    return b;
}
struct PairLayout {
    size_t pgid, tck, ptile_x, pstate, partial, total;
    size_t rb[2], rb_hist, rb_base, rb_tot, rb_bstart;   // the binning's buffers
    int64_t xcap;
};
PairLayout pair_layout(int64_t cap, int ntiles) {
    PairLayout L;
    if (cap < 16) cap = 16;
    L.xcap = 2 * (cap / g_piece) + 2;   // slots of the extra-piece table: sum over multi-piece tiles of their pieces
    size_t o = 0;
    auto take = [&](size_t bytes) { const size_t at = o; o += align256(bytes); return at; };
    L.pgid = take(static_cast<size_t>(cap) * 4);
    L.tck = take((static_cast<size_t>(cap >> SUB_SHIFT) + ntiles + 2) * 32 * 4);
    L.ptile_x = take(static_cast<size_t>(L.xcap) * 4);
    L.pstate = take(static_cast<size_t>(L.xcap) * PIECE_STATE * 4);
    L.partial = take(static_cast<size_t>(cap) * 32);
    {
        const BinPlan b = bin_plan(cap, ntiles);
        const size_t cells = (static_cast<size_t>(1) << b.bits) * b.nb;
        size_t r = L.partial;
        auto rtake = [&](size_t bytes) { const size_t at = r; r += align256(bytes); return at; };
        for (int k = 0; k < 2; ++k) L.rb[k] = rtake(static_cast<size_t>(cap + PAIR_PAD) * 8);
        L.rb_hist = rtake(cells * 4);
        L.rb_base = rtake(cells * 4);
        L.rb_tot = rtake(static_cast<size_t>(4) << BIN_MAX_BITS);
        L.rb_bstart = rtake(static_cast<size_t>(b.nb + 1) * 4);
        if (r > o) o = r;
    }
    L.total = o;
    return L;
}

// persistent grid of the walk kernels: every resident warp slot of the device
thread_local int t_walk_per_sm = 0;   // > 0: upper bound of resident walk CTAs per SM (set by gcp_views_step)

unsigned walk_grid(const void *kernel, int threads, int dyn_smem = 0) {
    static const void *known[4] = {nullptr, nullptr, nullptr, nullptr};
    static unsigned slots[4] = {0, 0, 0, 0};
    static unsigned n_sms = 148;
    int i = 0;
    while (i < 3 && known[i] != nullptr && known[i] != kernel) ++i;
    if (known[i] != kernel) {
        int dev = 0, sms = 148, per_sm = 1;
        if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (dyn_smem > 48 * 1024) cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn_smem);
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, dyn_smem) != cudaSuccess || per_sm < 1)
            per_sm = 1;
        slots[i] = static_cast<unsigned>(sms * per_sm);
        n_sms = static_cast<unsigned>(sms);
        known[i] = kernel;
    }
    // a batch over three or more lanes (gcp_views_step) caps the walks at two CTAs per SM: the third of the
    // register file that stays free lets the other lanes' binning and reduce kernels (bound by memory latency, few
    // issue slots) run beside a walk instead of queueing behind its persistent CTAs: 64 views 39.5 -> 38.2 ms
    if (t_walk_per_sm > 0) return std::max(1u, std::min(slots[i], n_sms * static_cast<unsigned>(t_walk_per_sm)));
    return std::max(1u, slots[i]);
}

template <typename T>
inline T *at(void *base, size_t off) { return reinterpret_cast<T *>(static_cast<unsigned char *>(base) + off); }
template <typename T>
inline const T *at(const void *base, size_t off) {
    return reinterpret_cast<const T *>(static_cast<const unsigned char *>(base) + off);
}

// one digit pass: the instantiation with just enough ballots for the digit
template <bool LAST, int BITS>
void launch_bin_scatter_as(const BinPlan &bp, size_t smem, cudaStream_t st, const int2 *in, const int32_t *rowpre,
                           const int32_t *tot, int shift, int64_t cap, const unsigned int *hdr, int2 *out, int32_t *pgid) {
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(k_view_bin_scatter<LAST, BITS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             ((BIN_WARPS + 3) << BIN_MAX_BITS) * 4 + 8 + BIN_CHUNK * 8);
        attr_set = true;
    }
    DepLaunch{bp.nb, BIN_THREADS, smem, st}(k_view_bin_scatter<LAST, BITS>, in, rowpre, tot, shift, bp.bits, bp.nb, cap, hdr,
                                                                     LAST ? nullptr : out,
                                                                     LAST ? reinterpret_cast<int32_t *>(out) : nullptr,
                                                                     LAST ? pgid : nullptr);
}
template <bool LAST>
void launch_bin_scatter_last(const BinPlan &bp, size_t smem, cudaStream_t st, const int2 *in, const int32_t *rowpre,
                             const int32_t *tot, int shift, int64_t cap, const unsigned int *hdr, int2 *out,
                             int32_t *pgid) {
    if (bp.bits <= 7) launch_bin_scatter_as<LAST, 7>(bp, smem, st, in, rowpre, tot, shift, cap, hdr, out, pgid);
    else if (bp.bits == 8) launch_bin_scatter_as<LAST, 8>(bp, smem, st, in, rowpre, tot, shift, cap, hdr, out, pgid);
    else launch_bin_scatter_as<LAST, 9>(bp, smem, st, in, rowpre, tot, shift, cap, hdr, out, pgid);
}
void launch_bin_scatter(bool last, const BinPlan &bp, size_t smem, cudaStream_t st, const int2 *in, const int32_t *rowpre,
                        const int32_t *tot, int shift, int64_t cap, const unsigned int *hdr, int2 *out, int32_t *pgid) {
    if (last) launch_bin_scatter_last<true>(bp, smem, st, in, rowpre, tot, shift, cap, hdr, out, pgid);
    else launch_bin_scatter_last<false>(bp, smem, st, in, rowpre, tot, shift, cap, hdr, out, pgid);
}

thread_local int t_view_launches = 0;

// experiments only: an integer from the environment
int env_int(const char *name, int fallback) {
    const char *v = getenv(name);
    return (v && *v) ? atoi(v) : fallback;
}

}  // namespace

extern "C" {

int gcp_tile_width(void) { return TW; }
int gcp_tile_height(void) { return TH; }
int gcp_tile_num_tiles(int W, int H) { return bad_image(W, H) ? 0 : tiles_x(W) * tiles_y(H); }
int gcp_tile_set_piece_pairs(int pairs) {
    if (pairs < 32 || pairs > (1 << 20) || (pairs & 31)) return GCP_ERR_INVALID_ARG;
    g_piece = pairs;
    return GCP_OK;
}
int gcp_tile_piece_pairs(void) { return g_piece; }
int gcp_view_last_launch_count(void) { return t_view_launches; }

size_t gcp_view_plan_bytes(int64_t n, int W, int H) {
    return (n < 0 || bad_image(W, H)) ? 0 : plan_layout(n, tiles_x(W) * tiles_y(H)).total;
}
size_t gcp_view_pair_bytes(int64_t pair_cap, int W, int H) {
    return (pair_cap < 0 || bad_image(W, H)) ? 0 : pair_layout(pair_cap, tiles_x(W) * tiles_y(H)).total;
}
int gcp_view_layout(int64_t n, int W, int H, int64_t pair_cap, int64_t *out) {
    if (n < 0 || bad_image(W, H) || pair_cap < 0 || !out) return GCP_ERR_INVALID_ARG;
    const int ntiles = tiles_x(W) * tiles_y(H);
    const PlanLayout A = plan_layout(n, ntiles);
    const PairLayout B = pair_layout(pair_cap, ntiles);
    out[0] = static_cast<int64_t>(A.toff); out[1] = static_cast<int64_t>(A.tcount); out[2] = static_cast<int64_t>(A.tstart);
    out[3] = static_cast<int64_t>(A.pextra); out[4] = static_cast<int64_t>(A.hdr); out[5] = static_cast<int64_t>(A.rec);
    out[6] = static_cast<int64_t>(B.pgid); out[7] = static_cast<int64_t>(B.ptile_x); out[8] = B.xcap;
    return GCP_OK;
}

int gcp_view_plan(const int32_t *sp, const int32_t *ep, int64_t n, int W, int H, void *plan, size_t plan_bytes,
                  int64_t *totals_host, gcp_stream_t stream) {
    t_view_launches = 0;
    if (n < 0 || n >= (int64_t(1) << 31) - 64 || bad_image(W, H) || !plan) return GCP_ERR_INVALID_ARG;
    if (reinterpret_cast<uintptr_t>(plan) & 255) return GCP_ERR_WORKSPACE;
    const int ntiles = tiles_x(W) * tiles_y(H);
    const PlanLayout L = plan_layout(n, ntiles);
    if (plan_bytes < L.total) return GCP_ERR_WORKSPACE;
    if (n > 0 && (!sp || !ep || (reinterpret_cast<uintptr_t>(sp) & 7) || (reinterpret_cast<uintptr_t>(ep) & 7)))
        return GCP_ERR_INVALID_ARG;
    auto st = reinterpret_cast<cudaStream_t>(stream);
    cudaError_t e = cudaMemsetAsync(plan, 0, L.zero_bytes, st);
    if (e != cudaSuccess) return static_cast<int>(e);
    unsigned int *hdr = at<unsigned int>(plan, L.hdr);
    if (n == 0) {
        DepLaunch{1, 32, 0, st}(k_view_scan_empty, at<int32_t>(plan, L.toff), hdr, totals_host);
        ++t_view_launches;
        return static_cast<int>(cudaGetLastError());
    }
    DepLaunch{blocks_for(n, 256), 256, 0, st}(k_view_cnt, sp, ep, n, W, H, at<int32_t>(plan, L.cnt), at<int2>(plan, L.tbox));
    DepLaunch{L.nb1, SCAN_THREADS, 0, st}(k_view_scan, at<int32_t>(plan, L.cnt), n, at<int32_t>(plan, L.toff),
                                                hdr + H_TICKET_S1, at<unsigned long long>(plan, L.desc1),
                                                reinterpret_cast<unsigned long long *>(hdr) + H_P64, totals_host);
    t_view_launches += 2;
    return static_cast<int>(cudaGetLastError());
}

int gcp_view_render(const int32_t *sp, const int32_t *ep, const float *mean, const float *lam, const float *opac,
                    const float *l_d, int64_t n, int W, int H, void *plan, size_t plan_bytes, void *pairs,
                    size_t pair_bytes, int64_t pair_cap, int keep, float *image, gcp_stream_t stream) {
    t_view_launches = 0;
    if (n < 0 || bad_image(W, H) || !plan || !pairs || !image || pair_cap < 0 || pair_cap >= (int64_t(1) << 31) - 64)
        return GCP_ERR_INVALID_ARG;
    if ((reinterpret_cast<uintptr_t>(plan) & 255) || (reinterpret_cast<uintptr_t>(pairs) & 255)) return GCP_ERR_WORKSPACE;
    if (n > 0 && (!sp || !ep || !mean || !lam || !opac || !l_d)) return GCP_ERR_INVALID_ARG;
    const int ntx = tiles_x(W), ntiles = ntx * tiles_y(H);
    const PlanLayout A = plan_layout(n, ntiles);
    const PairLayout B = pair_layout(pair_cap, ntiles);
    if (plan_bytes < A.total || pair_bytes < B.total) return GCP_ERR_WORKSPACE;
    auto st = reinterpret_cast<cudaStream_t>(stream);
    unsigned int *hdr = at<unsigned int>(plan, A.hdr);
    int32_t *tcount = at<int32_t>(plan, A.tcount), *tstart = at<int32_t>(plan, A.tstart);
    int32_t *pextra = at<int32_t>(plan, A.pextra), *ptile_x = at<int32_t>(pairs, B.ptile_x);
    int32_t *pgid = at<int32_t>(pairs, B.pgid);
    int4 *rec = at<int4>(plan, A.rec);
    int32_t *bstart = at<int32_t>(pairs, B.rb_bstart);
    const int bcount = bin_plan(pair_cap < 16 ? 16 : pair_cap, ntiles).nb + 1;
    if (n > 0) {
        if (((reinterpret_cast<uintptr_t>(mean) & 7) | (reinterpret_cast<uintptr_t>(lam) & 15)) == 0)
            DepLaunch{blocks_for(n, 256), 256, 0, st}(k_view_pack<true>, sp, ep, mean, lam, opac, l_d, n, W, H,
                                                                  at<int32_t>(plan, A.toff), rec, bstart, bcount, hdr, pair_cap);
        else
            DepLaunch{blocks_for(n, 256), 256, 0, st}(k_view_pack<false>, sp, ep, mean, lam, opac, l_d, n, W, H,
                                                                   at<int32_t>(plan, A.toff), rec, bstart, bcount, hdr, pair_cap);
        ++t_view_launches;
    }
    if (n > 0) {
        // stable radix sort of the pair list by tile: no atomics on tile counters, no sort afterwards
        const BinPlan bp = bin_plan(pair_cap < 16 ? 16 : pair_cap, ntiles);
        const int nbins = 1 << bp.bits;
        int32_t *hist = at<int32_t>(pairs, B.rb_hist), *rowpre = at<int32_t>(pairs, B.rb_base);
        int32_t *tot = at<int32_t>(pairs, B.rb_tot);
        const size_t smem = (static_cast<size_t>(BIN_WARPS + 3) * nbins + 2) * 4 + static_cast<size_t>(BIN_CHUNK) * 8;
        {
            static bool attr_set = false;   // (per process; setting the attribute again costs nothing)
            if (!attr_set) {
                cudaFuncSetAttribute(k_view_pairs, cudaFuncAttributeMaxDynamicSharedMemorySize, PW_SMEM);
                attr_set = true;
            }
        }
        DepLaunch{bp.nb, PAIR_THREADS, PW_SMEM, st}(k_view_pairs, at<int2>(plan, A.tbox), at<int32_t>(plan, A.toff), bstart, n, ntx, pair_cap, hdr, bp.bits, bp.nb,
                                                    at<int2>(pairs, B.rb[0]), hist);
        ++t_view_launches;
        for (int k = 0; k < bp.passes; ++k) {
            const int2 *in = at<int2>(pairs, B.rb[k & 1]);
            int2 *out = at<int2>(pairs, B.rb[(k + 1) & 1]);
            if (k > 0) {
                DepLaunch{bp.nb, BIN_THREADS, 0, st}(k_view_bin_hist, in, k * bp.bits, bp.bits, bp.nb, pair_cap, hdr, hist);
                ++t_view_launches;
            }
            DepLaunch{nbins, BIN_THREADS, 0, st}(k_view_bin_rowscan, hist, bp.nb, pair_cap, hdr, rowpre, tot);
            launch_bin_scatter(k + 1 == bp.passes, bp, smem, st, in, rowpre, tot, k * bp.bits, pair_cap, hdr, out, pgid);
            t_view_launches += 2;
        }
        DepLaunch{blocks_for(ntiles + 1, BIN_THREADS), BIN_THREADS, 0, st}(k_view_tiles, 
            at<int32_t>(pairs, B.rb[bp.passes & 1]), tot, (bp.passes - 1) * bp.bits, nbins, ntiles, g_piece, pair_cap, hdr,
            tcount, tstart, pextra, ptile_x, at<int32_t>(plan, A.mlist));
        ++t_view_launches;
    } else {
        // no Gaussians: every tile's list is empty (P = 0: nothing is read)
        DepLaunch{blocks_for(ntiles + 1, BIN_THREADS), BIN_THREADS, 0, st}(k_view_tiles, nullptr, nullptr, 0, 1, ntiles, g_piece,
                                                                                  pair_cap, hdr, tcount, tstart, pextra,
                                                                                  ptile_x, at<int32_t>(plan, A.mlist));
        ++t_view_launches;
    }
    float *tck = at<float>(pairs, B.tck), *pstate = at<float>(pairs, B.pstate);
    if (keep) {
        const unsigned grid = walk_grid(reinterpret_cast<const void *>(k_view_render<true>), TILE_WARPS * 32);
        DepLaunch{grid, TILE_WARPS * 32, 0, st}(k_view_render<true>, tcount, tstart, pextra, ptile_x, pgid, rec, hdr, pair_cap,
                                                              g_piece, ntx, ntiles, W, H, image, tck, pstate);
    } else {
        const unsigned grid = walk_grid(reinterpret_cast<const void *>(k_view_render<false>), TILE_WARPS * 32);
        DepLaunch{grid, TILE_WARPS * 32, 0, st}(k_view_render<false>, tcount, tstart, pextra, ptile_x, pgid, rec, hdr,
                                                               pair_cap, g_piece, ntx, ntiles, W, H, image, tck, pstate);
    }
    DepLaunch{296, 256, 0, st}(k_view_combine_fwd, tcount, pextra, at<int32_t>(plan, A.mlist), hdr, pair_cap, g_piece, ntx, W,
                                           H, pstate, image);
    t_view_launches += 2;
    return static_cast<int>(cudaGetLastError());
}

int gcp_view_forward(const int32_t *sp, const int32_t *ep, const float *mean, const float *lam, const float *opac,
                     const float *l_d, int64_t n, int W, int H, void *plan, size_t plan_bytes, void *pairs,
                     size_t pair_bytes, int64_t pair_cap, int keep, float *image, int64_t *totals_host,
                     gcp_stream_t stream) {
    int rc = gcp_view_plan(sp, ep, n, W, H, plan, plan_bytes, totals_host, stream);
    if (rc != GCP_OK) return rc;
    const int l1 = t_view_launches;
    rc = gcp_view_render(sp, ep, mean, lam, opac, l_d, n, W, H, plan, plan_bytes, pairs, pair_bytes, pair_cap, keep,
                         image, stream);
    t_view_launches += l1;
    return rc;
}

}  // extern "C"

namespace {

struct BackwardArgs {
    void *plan, *pairs;
    int64_t pair_cap, n;
    const float *grad_image;
    int W, H;
    GradOut out;
};
int check_backward(const BackwardArgs &a, size_t plan_bytes, size_t pair_bytes) {
    if (a.n < 0 || bad_image(a.W, a.H) || !a.plan || !a.pairs || !a.grad_image || a.pair_cap < 0) return GCP_ERR_INVALID_ARG;
    if (a.n == 0) return GCP_OK;
    if (!a.out.g_mean || !a.out.g_lam || !a.out.g_opac || !a.out.g_l) return GCP_ERR_INVALID_ARG;
    if ((reinterpret_cast<uintptr_t>(a.out.g_mean) & 7) || (reinterpret_cast<uintptr_t>(a.out.g_lam) & 15)) return GCP_ERR_INVALID_ARG;
    const int ntiles = tiles_x(a.W) * tiles_y(a.H);
    if (plan_bytes < plan_layout(a.n, ntiles).total || pair_bytes < pair_layout(a.pair_cap, ntiles).total) return GCP_ERR_WORKSPACE;
    return GCP_OK;
}
// the reverse walk: carries of multi-piece tiles, then the pieces -> one partial per pair
int launch_backward_walk(const BackwardArgs &a, cudaStream_t st) {
    const int ntx = tiles_x(a.W), ntiles = ntx * tiles_y(a.H);
    const PlanLayout A = plan_layout(a.n, ntiles);
    const PairLayout B = pair_layout(a.pair_cap, ntiles);
    unsigned int *hdr = at<unsigned int>(a.plan, A.hdr);
    cudaError_t e = cudaMemsetAsync(hdr + H_TICKET_BWD, 0, 3 * sizeof(unsigned int), st);
    if (e != cudaSuccess) return static_cast<int>(e);
    const int32_t *tcount = at<int32_t>(a.plan, A.tcount), *tstart = at<int32_t>(a.plan, A.tstart);
    const int32_t *pextra = at<int32_t>(a.plan, A.pextra), *ptile_x = at<int32_t>(a.pairs, B.ptile_x);
    float *pstate = at<float>(a.pairs, B.pstate);
    DepLaunch{296, 256, 0, st}(k_view_combine_bwd, tcount, pextra, at<int32_t>(a.plan, A.mlist), hdr, a.pair_cap, g_piece,
                                           a.grad_image, ntx, a.W, a.H, pstate);
    // (3 resident CTAs per SM; 2 and 4 — 114 / 64 registers — were measured within 5 %: 284 / 280 vs 270 us)
    const unsigned grid = walk_grid(reinterpret_cast<const void *>(k_view_backward), BWD_WARPS * 32, BWD_SMEM);
    DepLaunch{grid, BWD_WARPS * 32, BWD_SMEM, st}(k_view_backward, tcount, tstart, pextra, ptile_x, at<int32_t>(a.pairs, B.pgid),
                                                            at<int4>(a.plan, A.rec), hdr, a.pair_cap, g_piece,
                                                            at<float>(a.pairs, B.tck), pstate, a.grad_image, ntx,
                                                            ntiles, a.W, a.H, at<float>(a.pairs, B.partial));
    t_view_launches += 2;
    return static_cast<int>(cudaGetLastError());
}
// the partials of every Gaussian's pairs summed and turned into the four gradients
int launch_backward_reduce(const BackwardArgs &a, cudaStream_t st) {
    const int ntiles = tiles_x(a.W) * tiles_y(a.H);
    const PlanLayout A = plan_layout(a.n, ntiles);
    const PairLayout B = pair_layout(a.pair_cap, ntiles);
    unsigned int *hdr = at<unsigned int>(a.plan, A.hdr);
    const int4 *rec = at<int4>(a.plan, A.rec);
    const int32_t *toff = at<int32_t>(a.plan, A.toff);
    const float *partial = at<float>(a.pairs, B.partial);
    int32_t *big = at<int32_t>(a.plan, A.big);
    DepLaunch{blocks_for(a.n, 256, 148 * 8), 256, 0, st}(k_view_reduce, partial, toff, rec, a.n, a.out, hdr, a.pair_cap, big);
    DepLaunch{148 * 4, 256, 0, st}(k_view_reduce_big, partial, toff, rec, a.out, hdr, a.pair_cap, big);
    t_view_launches += 2;
    return static_cast<int>(cudaGetLastError());
}

}  // namespace

extern "C" {

int gcp_view_backward_scatter(void *plan, size_t plan_bytes, void *pairs, size_t pair_bytes, int64_t pair_cap,
                              const float *grad_image, int64_t n, int W, int H, const int32_t *index, float *g_mean,
                              float *g_lam, float *g_opac, float *g_l, gcp_stream_t stream) {
    t_view_launches = 0;
    const BackwardArgs a = {plan, pairs, pair_cap, n, grad_image, W, H, {g_mean, g_lam, g_opac, g_l, index}};
    int rc = check_backward(a, plan_bytes, pair_bytes);
    if (rc != GCP_OK || n == 0) return rc;
    auto st = reinterpret_cast<cudaStream_t>(stream);
    rc = launch_backward_walk(a, st);
    if (rc != GCP_OK) return rc;
    return launch_backward_reduce(a, st);
}

int gcp_view_backward(void *plan, size_t plan_bytes, void *pairs, size_t pair_bytes, int64_t pair_cap,
                      const float *grad_image, int64_t n, int W, int H, float *g_mean, float *g_lam, float *g_opac,
                      float *g_l, gcp_stream_t stream) {
    return gcp_view_backward_scatter(plan, plan_bytes, pairs, pair_bytes, pair_cap, grad_image, n, W, H, nullptr, g_mean,
                                     g_lam, g_opac, g_l, stream);
}

// ---------------------------------------------------------------------------------------------------------------
// A batch of views in one call (the per-view loop of gs_model.py:402-449 around the compositor, forward AND
// backward, with the views' gradients summed into the parameters' gradient arrays as autograd does for the
// reference, gs_control.py:180-185).  The host enqueues everything without waiting for the device: every view is
// rendered on the pair capacity the caller provides (gcp_view_forward), its pair count lands in totals_host[v], a
// view that does not fit is skipped entirely (forward and backward) and the caller, who sees the count, repeats
// the step on larger arenas.  Views alternate between `lanes` streams with their own arenas, so that the binning
// kernels of one view (bound by memory latency) share the device with the walk kernels of another (bound by
// instruction issue); the scatter-adds into the shared gradient arrays are chained by events in view order, which
// keeps every float sum in a fixed order.
// ---------------------------------------------------------------------------------------------------------------
struct gcp_views_ctx {
    int lanes;
    cudaStream_t side[GCP_VIEWS_MAX_LANES];      // lane 0 is the caller's stream
    cudaEvent_t ev_in, ev_reduced[GCP_VIEWS_MAX_LANES], ev_done[GCP_VIEWS_MAX_LANES];
};

}  // extern "C"

namespace {
// grad_image = d/d image of mean((image - target)^2); the loss itself is added to *loss (float atomics: the
// reported number only, no gradient depends on it)
__global__ void __launch_bounds__(256)
k_view_mse_grad(const float *__restrict__ image, const float *__restrict__ target, int64_t count, float scale,
                const unsigned int *__restrict__ hdr, int64_t cap, float *__restrict__ gimg, float *loss) {
    grid_dep_sync();
    if (overflowed(hdr, cap)) return;
    float acc = 0.0f;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < count;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const float d = image[i] - __ldg(target + i);
        gimg[i] = scale * d;
        acc = fmaf(d, d, acc);
    }
    if (loss != nullptr) {
        __shared__ float s_acc[8];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if ((threadIdx.x & 31) == 0) s_acc[threadIdx.x >> 5] = acc;
        __syncthreads();
        if (threadIdx.x == 0) {
            float t = 0.0f;
            for (int w = 0; w < 8; ++w) t += s_acc[w];
            atomicAdd(loss, t * (0.5f * scale));
        }
    }
}
}  // namespace

extern "C" {

int gcp_views_ctx_create(int lanes, gcp_views_ctx **out) {
    if (!out || lanes < 1 || lanes > GCP_VIEWS_MAX_LANES) return GCP_ERR_INVALID_ARG;
    gcp_views_ctx *c = new (std::nothrow) gcp_views_ctx();
    if (!c) return GCP_ERR_INVALID_ARG;
    c->lanes = lanes;
    cudaError_t e = cudaEventCreateWithFlags(&c->ev_in, cudaEventDisableTiming);
    for (int l = 0; l < lanes && e == cudaSuccess; ++l) {
        c->side[l] = nullptr;
        if (l > 0) e = cudaStreamCreateWithFlags(&c->side[l], cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_reduced[l], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_done[l], cudaEventDisableTiming);
    }
    if (e != cudaSuccess) { delete c; return static_cast<int>(e); }
    *out = c;
    return GCP_OK;
}

void gcp_views_ctx_destroy(gcp_views_ctx *c) {
    if (!c) return;
    cudaEventDestroy(c->ev_in);
    for (int l = 0; l < c->lanes; ++l) {
        if (l > 0) cudaStreamDestroy(c->side[l]);
        cudaEventDestroy(c->ev_reduced[l]);
        cudaEventDestroy(c->ev_done[l]);
    }
    delete c;
}

int gcp_views_step(gcp_views_ctx *ctx, const gcp_view_desc *views, int n_views, int W, int H, void *const *plan,
                   size_t plan_bytes, void *const *pairs, size_t pair_bytes, int64_t pair_cap, float *g_mean,
                   float *g_lam, float *g_opac, float *g_l, float *loss, int64_t *totals_host, gcp_stream_t stream) {
    return gcp_views_step_split(ctx, views, n_views, W, H, plan, plan_bytes, pairs, pair_bytes, pair_cap, g_mean, g_lam,
                                g_opac, g_l, loss, totals_host, nullptr, stream);
}

int gcp_views_step_split(gcp_views_ctx *ctx, const gcp_view_desc *views, int n_views, int W, int H, void *const *plan,
                         size_t plan_bytes, void *const *pairs, size_t pair_bytes, int64_t pair_cap, float *g_mean,
                         float *g_lam, float *g_opac, float *g_l, float *loss, int64_t *totals_host,
                         const gcp_views_split *split, gcp_stream_t stream) {
    if (!ctx || n_views < 0 || (n_views > 0 && !views) || bad_image(W, H) || !plan || !pairs || !totals_host)
        return GCP_ERR_INVALID_ARG;
    if (split && (split->first_tail_view < 0 || !split->g_mean || !split->g_lam || !split->g_opac || !split->g_l))
        return GCP_ERR_INVALID_ARG;
    const int first_tail = split ? split->first_tail_view : n_views;
    auto main_st = reinterpret_cast<cudaStream_t>(stream);
    const int lanes = ctx->lanes;
    int launches = 0;
    cudaError_t e = cudaSuccess;
    if (lanes > 1) {
        e = cudaEventRecord(ctx->ev_in, main_st);
        for (int l = 1; l < lanes && e == cudaSuccess; ++l) e = cudaStreamWaitEvent(ctx->side[l], ctx->ev_in, 0);
        if (e != cudaSuccess) return static_cast<int>(e);
    }
    const int64_t count = static_cast<int64_t>(W + 1) * (H + 1) * 3;
    const int ntiles = tiles_x(W) * tiles_y(H);
    struct WalkCap {   // restored on every return path
        WalkCap(int v, bool dep) { t_walk_per_sm = v; t_dep_launch = dep; }
        ~WalkCap() { t_walk_per_sm = 0; t_dep_launch = true; }
    } walk_cap(env_int("GCP_WALK_PER_SM", lanes >= 3 ? 2 : 0), env_int("GCP_BATCH_PDL", lanes <= 2) != 0);
    for (int v = 0; v < n_views; ++v) {
        const gcp_view_desc &d = views[v];
        const int lane = v % lanes;
        cudaStream_t st = lane == 0 ? main_st : ctx->side[lane];
        if (!d.image || !d.grad_image) return GCP_ERR_INVALID_ARG;
        int rc = gcp_view_forward(d.sp, d.ep, d.mean, d.lam, d.opac, d.l_d, d.n, W, H, plan[lane], plan_bytes, pairs[lane],
                                  pair_bytes, pair_cap, 1, d.image, totals_host + v, st);
        if (rc != GCP_OK) return rc;
        launches += t_view_launches;
        if (d.n > 0) {
            const unsigned int *hdr = at<unsigned int>(plan[lane], plan_layout(d.n, ntiles).hdr);
            if (d.target != nullptr) {
                DepLaunch{148 * 8, 256, 0, st}(k_view_mse_grad, d.image, d.target, count, 2.0f / static_cast<float>(count), hdr,
                                                         pair_cap, d.grad_image, loss);
                ++launches;
            }
            const bool tail = v >= first_tail;   // the last views add into arrays of their own (see gcp_views_split)
            const BackwardArgs a = {plan[lane], pairs[lane], pair_cap, d.n, d.grad_image, W, H,
                                    {tail ? split->g_mean : g_mean, tail ? split->g_lam : g_lam,
                                     tail ? split->g_opac : g_opac, tail ? split->g_l : g_l, d.index}};
            rc = check_backward(a, plan_bytes, pair_bytes);
            if (rc != GCP_OK) return rc;
            t_view_launches = 0;
            rc = launch_backward_walk(a, st);
            if (rc != GCP_OK) return rc;
            // the sums into the shared gradient arrays follow each other in view order
            if (lanes > 1 && v > 0) {
                e = cudaStreamWaitEvent(st, ctx->ev_reduced[(v - 1) % lanes], 0);
                if (e != cudaSuccess) return static_cast<int>(e);
            }
            rc = launch_backward_reduce(a, st);
            if (rc != GCP_OK) return rc;
            launches += t_view_launches;
        } else if (lanes > 1 && v > 0) {
            e = cudaStreamWaitEvent(st, ctx->ev_reduced[(v - 1) % lanes], 0);   // keeps the chain of events unbroken
            if (e != cudaSuccess) return static_cast<int>(e);
        }
        if (lanes > 1) {
            e = cudaEventRecord(ctx->ev_reduced[lane], st);
            if (e != cudaSuccess) return static_cast<int>(e);
        }
        // the main arrays are complete once the last view in front of the tail has been added: the caller's event
        // fires here, in the middle of the batch (the reduces are chained in view order)
        if (split && split->event && v == first_tail - 1) {
            e = cudaEventRecord(reinterpret_cast<cudaEvent_t>(split->event), st);
            if (e != cudaSuccess) return static_cast<int>(e);
        }
    }
    if (split && split->event && (first_tail <= 0 || first_tail > n_views)) {   // no view in front of the tail
        e = cudaEventRecord(reinterpret_cast<cudaEvent_t>(split->event), main_st);
        if (e != cudaSuccess) return static_cast<int>(e);
    }
    for (int l = 1; l < lanes; ++l) {
        e = cudaEventRecord(ctx->ev_done[l], ctx->side[l]);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(main_st, ctx->ev_done[l], 0);
        if (e != cudaSuccess) return static_cast<int>(e);
    }
    t_view_launches = launches;
    return GCP_OK;
}

}  // extern "C"
