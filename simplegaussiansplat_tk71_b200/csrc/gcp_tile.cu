// gcp_tile.cu — the fused compositor route: exclusive transmittance, colour sum and the
// division-free backward in ONE pass per direction, without materialising the per-pixel element lists.
//
// The per-pixel segmented scan  T_i = prod_{j<i} (1 - alpha_j),  C = sum_i T_i alpha_i l_i  (gs_model.py:544-566,
// :498-514) is evaluated with one pixel per lane: the image is cut into tiles of 8 x 4 pixels = one warp, every
// box contributes one (tile, Gaussian) pair per tile it touches, the pairs are sorted by tile with a stable radix
// sort — the Gaussians arrive in depth order, so every tile list (and with it every pixel list) stays in depth
// order, the same order torch.sort gives the reference at gs_model.py:547 — and a warp walks its tile's list
// with the running T of its pixels in registers.
//
//   forward   k_tile_render   : alpha = o * exp(-1/2 d Lambda d^T) (:493-495,:533-535), T, colour; the exclusive
//                               T of every (pair, lane) is kept for the backward (128 contiguous bytes per pair)
//   backward  k_tile_backward : the list walked in reverse with U_i = w_{i+1} + (1-alpha_{i+1}) U_{i+1}
//                               (w = <dL/dI, alpha l>), dL/dalpha_i = T_i <dL/dI, l_i> - T_i U_i — no division by
//                               1-alpha (:736,:747,:757 divide) — and the reference's per-element gradients
//                               (:733-766) summed over the lanes of the pair in a fixed order;
//             k_tile_reduce   : the partial sums of a Gaussian's pairs added in pair order (:776-783).
// No float atomics anywhere: a pixel belongs to one lane, a partial to one pair — bitwise reproducible.
// Elements whose inclusive product is 0 contribute nothing and get no gradient (gs_model.py:575-578).
#include <cuda_runtime.h>
#include <stdint.h>
#include <algorithm>

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <thrust/iterator/counting_iterator.h>
#include <thrust/iterator/transform_iterator.h>

#include "gcp_abi.h"

namespace {

constexpr int TSX = 3, TSY = 2;                 // tile = 8 x 4 pixels, lane = (y & 3) * 8 + (x & 7)
constexpr int TW = 1 << TSX, TH = 1 << TSY;
static_assert(TW * TH == 32, "one tile is one warp");

inline unsigned blocks_for(int64_t work, int per_block, unsigned cap = 0x7fffffffu) {
    int64_t b = (work + per_block - 1) / per_block;
    if (b < 1) b = 1;
    if (b > cap) b = cap;
    return static_cast<unsigned>(b);
}
inline int key_bits(int max_key) {
    int b = 1;
    while (b < 31 && (1 << b) <= max_key) ++b;
    return b;
}
inline size_t align256(size_t v) { return (v + 255) & ~static_cast<size_t>(255); }

// box of Gaussian g clipped to the image [0,W] x [0,H] (the caller clamps already, gs_model.py:419-425)
struct Box {
    int sx, sy, ex, ey;
};
__host__ __device__ __forceinline__ Box clip_box(const int32_t *__restrict__ sp, const int32_t *__restrict__ ep,
                                                 int64_t g, int W, int H) {
    Box b;
    b.sx = sp[2 * g] > 0 ? sp[2 * g] : 0;
    b.sy = sp[2 * g + 1] > 0 ? sp[2 * g + 1] : 0;
    b.ex = ep[2 * g] < W ? ep[2 * g] : W;
    b.ey = ep[2 * g + 1] < H ? ep[2 * g + 1] : H;
    return b;
}
// (tile, Gaussian) pairs of one box (0 for an empty / inverted box)
struct TilePairCount {
    const int32_t *sp, *ep;
    int W, H;
    __host__ __device__ __forceinline__ int64_t operator()(int64_t g) const {
        const Box b = clip_box(sp, ep, g, W, H);
        if (b.ex < b.sx || b.ey < b.sy) return 0;
        return static_cast<int64_t>((b.ex >> TSX) - (b.sx >> TSX) + 1) * ((b.ey >> TSY) - (b.sy >> TSY) + 1);
    }
};

__global__ void k_tile_totals(const int64_t *__restrict__ toff, int64_t n, int64_t *__restrict__ totals) {
    totals[0] = toff[n];
}

// Lambda is stored pre-multiplied by -log2(e)/2, so that the walk kernels get g = exp(-1/2 d Lambda d^T) as one
// ex2.approx of d Lambda' d^T (relative error 2^-22; expf costs eight instructions, this one two); the backward
// multiplies its d_mean terms by EXP2_UNSCALE = -2 ln 2 to undo the factor in X = d Lambda.
constexpr float EXP2_SCALE = -0.72134752044448170368f;    // -log2(e) / 2
constexpr float EXP2_UNSCALE = -1.38629436111989061883f;  // 1 / EXP2_SCALE

// rec[g] = {mx, my, l00', l01' | l10', l11', o, l0 || l1, l2, sx, sy | ex, ey, toff, 0}
__global__ void __launch_bounds__(256)
k_tile_pack(const float *__restrict__ mean, const float *__restrict__ lam, const float *__restrict__ opac,
            const float *__restrict__ l_d, const int32_t *__restrict__ sp, const int32_t *__restrict__ ep,
            const int64_t *__restrict__ toff, int64_t n, int W, int H, int4 *__restrict__ rec) {
    const int64_t g = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (g >= n) return;
    const float2 m = __ldg(reinterpret_cast<const float2 *>(mean) + g);
    const float4 L = __ldg(reinterpret_cast<const float4 *>(lam) + g);
    const Box b = clip_box(sp, ep, g, W, H);
    auto f = [](float v) { return __float_as_int(v); };
    rec[4 * g] = make_int4(f(m.x), f(m.y), f(L.x * EXP2_SCALE), f(L.y * EXP2_SCALE));
    rec[4 * g + 1] = make_int4(f(L.z * EXP2_SCALE), f(L.w * EXP2_SCALE), f(__ldg(opac + g)), f(__ldg(l_d + 3 * g)));
    rec[4 * g + 2] = make_int4(f(__ldg(l_d + 3 * g + 1)), f(__ldg(l_d + 3 * g + 2)), b.sx, b.sy);
    rec[4 * g + 3] = make_int4(b.ex, b.ey, static_cast<int>(__ldg(toff + g)), 0);
}

// pair p of Gaussian g (Gaussian-major, row-major over the tiles of its box): ptile[p] = tile index, pgid[p] = g.
// One thread per Gaussian writes the pairs of a small box (adjacent threads write adjacent runs); a box of more
// than 32 tiles is written by the whole warp afterwards.
__global__ void __launch_bounds__(256)
k_tile_pairs(const int32_t *__restrict__ sp, const int32_t *__restrict__ ep, const int64_t *__restrict__ toff,
             int64_t n, int64_t cap, int W, int H, int ntx, int32_t *__restrict__ ptile, int32_t *__restrict__ pgid) {
    const int64_t g = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    int64_t beg = 0;
    int cnt = 0, tx0 = 0, ty0 = 0, nx = 1;
    if (g < n) {
        beg = __ldg(toff + g);
        cnt = static_cast<int>(__ldg(toff + g + 1) - beg);
        const Box b = clip_box(sp, ep, g, W, H);
        tx0 = b.sx >> TSX; ty0 = b.sy >> TSY; nx = (b.ex >> TSX) - tx0 + 1;
    }
    if (cnt <= 32) {
        int r = 0, c = 0;
        for (int i = 0; i < cnt && beg + i < cap; ++i) {   // cap: the buffers' capacity (>= the pair count,
            ptile[beg + i] = (ty0 + r) * ntx + tx0 + c;      // except in a speculative call that guessed too low)
            pgid[beg + i] = static_cast<int32_t>(g);
            if (++c == nx) { c = 0; ++r; }
        }
    }
    unsigned big = __ballot_sync(0xffffffffu, cnt > 32);
    while (big) {
        const int src = __ffs(big) - 1;
        big &= big - 1;
        const int64_t bbeg = __shfl_sync(0xffffffffu, beg, src);
        const int bcnt = __shfl_sync(0xffffffffu, cnt, src);
        const int btx0 = __shfl_sync(0xffffffffu, tx0, src), bty0 = __shfl_sync(0xffffffffu, ty0, src);
        const int bnx = __shfl_sync(0xffffffffu, nx, src);
        const int32_t bg = static_cast<int32_t>(g - lane + src);
        for (int i = lane; i < bcnt && bbeg + i < cap; i += 32) {
            const int r = i / bnx;
            ptile[bbeg + i] = (bty0 + r) * ntx + btx0 + (i - r * bnx);
            pgid[bbeg + i] = bg;
        }
    }
}

// first index g with off[g+1] > e
__device__ __forceinline__ int64_t find_owner(const int64_t *__restrict__ off, int64_t n, int64_t e) {
    int64_t lo = 0, hi = n - 1;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (__ldg(off + mid + 1) > e) hi = mid;
        else lo = mid + 1;
    }
    return lo;
}

// the same list, parallel over PAIRS (8 consecutive ones per thread, one search for the first): for views whose
// boxes span many tiles each (bundled scene: ten pairs per Gaussian on average, thousands for some)
constexpr int CHP = 8;
__global__ void __launch_bounds__(256)
k_tile_pairs_flat(const int32_t *__restrict__ sp, const int32_t *__restrict__ ep, const int64_t *__restrict__ toff,
                  int64_t n, int64_t cap, int W, int H, int ntx, int32_t *__restrict__ ptile,
                  int32_t *__restrict__ pgid) {
    const int64_t p0 = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) * CHP;
    const int64_t total = __ldg(toff + n);          // the pair count lives on the device; cap = buffer capacity
    const int64_t P = total < cap ? total : cap;
    if (p0 >= P) return;
    int64_t g = find_owner(toff, n, p0);
    int64_t gbeg = __ldg(toff + g), gend = __ldg(toff + g + 1);
    Box b = clip_box(sp, ep, g, W, H);
    int tx0 = b.sx >> TSX, ty0 = b.sy >> TSY, nx = (b.ex >> TSX) - tx0 + 1;
    for (int i = 0; i < CHP && p0 + i < P; ++i) {
        const int64_t p = p0 + i;
        while (p >= gend) {  // skips Gaussians without pairs too
            ++g;
            gbeg = gend;
            gend = __ldg(toff + g + 1);
            b = clip_box(sp, ep, g, W, H);
            tx0 = b.sx >> TSX; ty0 = b.sy >> TSY; nx = (b.ex >> TSX) - tx0 + 1;
        }
        const int local = static_cast<int>(p - gbeg);
        const int r = local / nx;
        ptile[p] = (ty0 + r) * ntx + tx0 + (local - r * nx);
        pgid[p] = static_cast<int32_t>(g);
    }
}

// speculative binning (the host guessed a capacity instead of waiting for the pair count): the slots behind the
// real pairs get the key `ntiles`, which sorts behind every tile and is what k_tile_start expects after the last pair
__global__ void __launch_bounds__(256)
k_tile_pad(const int64_t *__restrict__ toff, int64_t n, int64_t cap, int ntiles, int32_t *__restrict__ ptile,
           int32_t *__restrict__ pgid) {
    const int64_t total = __ldg(toff + n);
    for (int64_t p = total + static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; p < cap;
         p += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        ptile[p] = ntiles;
        pgid[p] = 0;
    }
}

// tstart[t] = first pair of tile t in the tile-sorted pair list (tstart[ntiles] = P), empty tiles included
__global__ void __launch_bounds__(256)
k_tile_start(const int32_t *__restrict__ ptile_s, int64_t P, int ntiles, int32_t *__restrict__ tstart) {
    const int64_t p = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (p > P) return;
    const int prev = (p == 0) ? -1 : __ldg(ptile_s + p - 1);
    const int cur = (p == P) ? ntiles : __ldg(ptile_s + p);
    for (int c = prev + 1; c <= cur; ++c) tstart[c] = static_cast<int32_t>(p);
}

// Work units of the walk kernels are PIECES: at most `piece` consecutive pairs of one tile's list (a multiple of
// 32).  Most tiles are one piece.  A tile with a long list (bundled scene: thousands of Gaussians over one tile,
// ten times the mean) is cut into several, walked by different warps at the same time, each starting from the
// identity; the per-pixel carries between the pieces of a tile are the segmented scan's cross-block carries:
//   forward : T_i = carry_p * Tlocal_i,  carry_p = prod of the aggregates of the earlier pieces; the colour is
//             linear in the carry, C = sum_p carry_p * C_p                                  (k_tile_combine_fwd)
//   backward: U after the last element of piece p:  U_in(p) = Tagg_{p+1} U_in(p+1) + <dL/dI, C_{p+1}>  — the
//             affine map of a piece is made of the SAME two per-pixel quantities the forward already produced,
//             its aggregate and its colour, so the backward needs no pass of its own for them  (k_tile_combine_bwd)
// piece state per piece (f32[192]): {Tagg[32], C0[32], C1[32], C2[32], carry[32], U_in[32]}, touched only for the
// pieces of multi-piece tiles.
constexpr int PIECE_STATE = 192;

struct PieceCount {   // pieces of tile t: ceil(len / piece), 1 for an empty tile (its pixels still get written)
    const int32_t *tstart;
    int piece;
    __host__ __device__ __forceinline__ int32_t operator()(int32_t t) const {
        const int len = tstart[t + 1] - tstart[t];
        return len <= piece ? 1 : (len + piece - 1) / piece;
    }
};

// ptile[p] = tile of piece p  (pstart: exclusive offsets of the tiles' pieces, pstart[ntiles] = number of pieces)
__global__ void __launch_bounds__(256)
k_tile_pieces(const int32_t *__restrict__ pstart, int ntiles, int32_t *__restrict__ ptile) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ntiles) return;
    const int b = __ldg(pstart + t), e = __ldg(pstart + t + 1);
    for (int p = b; p < e; ++p) ptile[p] = t;
}

struct Piece {
    int p, t, np;        // piece id, tile, pieces of that tile
    int64_t lo, hi;      // pair range of the piece inside the tile-sorted pair list
};
// next piece of this warp (dynamic ticket); false when none is left
__device__ __forceinline__ bool next_piece(unsigned int *ticket, const int32_t *__restrict__ tstart,
                                           const int32_t *__restrict__ pstart, const int32_t *__restrict__ ptile,
                                           int npieces, int piece, int lane, Piece &w) {
    unsigned i = 0;
    if (lane == 0) i = atomicAdd(ticket, 1u);
    i = __shfl_sync(0xffffffffu, i, 0);
    if (i >= static_cast<unsigned>(npieces)) return false;
    w.p = static_cast<int>(i);
    w.t = __ldg(ptile + i);
    const int p0 = __ldg(pstart + w.t);
    w.np = __ldg(pstart + w.t + 1) - p0;
    const int64_t tlo = __ldg(tstart + w.t), thi = __ldg(tstart + w.t + 1);
    w.lo = tlo + static_cast<int64_t>(w.p - p0) * piece;
    w.hi = (w.lo + piece < thi) ? w.lo + piece : thi;
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

// What the walk needs of a pair, staged in shared memory by the lane that loaded it (three broadcast LDS.128 per
// pair in the walk): {mx, my, l00, l01} {l10, l11, o, l0} {l1, l2, coverage mask of the tile, Gaussian-major pair id}
struct PairSlot {
    float4 a, b;
    float2 c;
    uint32_t mask;
    int32_t q;
};
static_assert(sizeof(PairSlot) == 48, "three 16-byte words");

__device__ __forceinline__ void stage_pair(PairSlot *slot, const RecRegs &r, bool live, int tx, int ty, int x0, int y0) {
    const int sx = r.c.z, sy = r.c.w, ex = r.d.x, ey = r.d.y;
    // lanes of the tile inside the box: columns [xa, xb] of rows [ya, yb]
    const int xa = max(sx, x0) - x0, xb = min(ex, x0 + TW - 1) - x0;
    const int ya = max(sy, y0) - y0, yb = min(ey, y0 + TH - 1) - y0;
    uint32_t mask = 0;
    if (live && xa <= xb && ya <= yb) {
        const uint32_t xm = ((2u << xb) - 1u) & ~((1u << xa) - 1u);
        const uint32_t rows = (0x01010101u >> (8 * (TH - 1 - (yb - ya)))) << (8 * ya);
        mask = xm * rows;
    }
    const int tx0 = sx >> TSX, ty0 = sy >> TSY, nx = (ex >> TSX) - tx0 + 1;
    slot->a = make_float4(__int_as_float(r.a.x), __int_as_float(r.a.y), __int_as_float(r.a.z), __int_as_float(r.a.w));
    slot->b = make_float4(__int_as_float(r.b.x), __int_as_float(r.b.y), __int_as_float(r.b.z), __int_as_float(r.b.w));
    slot->c = make_float2(__int_as_float(r.c.x), __int_as_float(r.c.y));
    slot->mask = mask;
    slot->q = r.d.z + (ty - ty0) * nx + (tx - tx0);
}

struct PairEval {
    float d0, d1, X0, X1, gk, x;
};
// g = exp(-1/2 (r-m) Lambda (r-m)^T) with X = (r-m) Lambda (gs_model.py:495, :745); x = 1 - o g (:533-535).
// The record holds Lambda' = EXP2_SCALE Lambda: X0, X1 come out scaled by EXP2_SCALE and g = 2^(X' . d).
__device__ __forceinline__ PairEval eval_pair(const float4 &a, const float4 &b, float px, float py) {
    PairEval e;
    e.d0 = px - a.x;
    e.d1 = py - a.y;
    e.X0 = e.d0 * a.z + e.d1 * b.x;
    e.X1 = e.d0 * a.w + e.d1 * b.y;
    const float p2 = e.X0 * e.d0 + e.X1 * e.d1;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.gk) : "f"(p2));
    e.x = 1.0f - b.z * e.gk;
    return e;
}

constexpr int TILE_WARPS = 8;

template <bool KEEP>
__global__ void __launch_bounds__(TILE_WARPS * 32, 4)
k_tile_render(const int32_t *__restrict__ tstart, const int32_t *__restrict__ pgid_s, const int4 *__restrict__ rec,
              const int32_t *__restrict__ pstart, const int32_t *__restrict__ ptile, unsigned int *__restrict__ ticket,
              int piece, int ntx, int ntiles, int W, int H, float *__restrict__ image, float *__restrict__ tkeep,
              float *__restrict__ pstate) {
    __shared__ PairSlot slots[TILE_WARPS][32];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    PairSlot *sl = slots[wib];
    const int npieces = __ldg(pstart + ntiles);
    Piece w;
    while (next_piece(ticket, tstart, pstart, ptile, npieces, piece, lane, w)) {  // warp-uniform; only __syncwarp inside
        const int t = w.t;
        const int ty = t / ntx, tx = t - ty * ntx;
        const int x0 = tx << TSX, y0 = ty << TSY;
        const int ix = x0 + (lane & (TW - 1)), iy = y0 + (lane >> TSX);
        const float px = static_cast<float>(ix), py = static_cast<float>(iy);
        const int64_t lo = w.lo, hi = w.hi;
        float T = 1.0f, c0 = 0.f, c1 = 0.f, c2 = 0.f;   // T: local to the piece (its carry is applied afterwards)
        // software pipeline over batches of 32 pairs: ids two batches ahead, records one batch ahead
        int g1 = 0;
        RecRegs r = {};
        if (lo + lane < hi) r = load_rec(rec, __ldg(pgid_s + lo + lane));
        if (lo + 32 + lane < hi) g1 = __ldg(pgid_s + lo + 32 + lane);
        for (int64_t b = lo; b < hi; b += 32) {
            stage_pair(sl + lane, r, b + lane < hi, tx, ty, x0, y0);
            int g2 = 0;
            if (b + 64 + lane < hi) g2 = __ldg(pgid_s + b + 64 + lane);
            if (b + 32 + lane < hi) r = load_rec(rec, g1);
            g1 = g2;
            __syncwarp();
            const int m = static_cast<int>(hi - b < 32 ? hi - b : 32);
            float *tk = tkeep + b * 32 + lane;
#pragma unroll 4
            for (int k = 0; k < m; ++k) {
                const float4 A = sl[k].a, B = sl[k].b;
                const float2 C = sl[k].c;
                const bool cov = (sl[k].mask >> lane) & 1u;
                const PairEval e = eval_pair(A, B, px, py);
                const float tin = T * e.x;
                if (KEEP) __stcs(tk + k * 32, T);   // not kept for a render without backward
                // branch-free: an element outside the box, or dead (inclusive product 0, gs_model.py:575-578), adds 0
                const float ta = (cov && tin != 0.0f) ? T * (1.0f - e.x) : 0.0f;
                c0 = fmaf(ta, B.w, c0);
                c1 = fmaf(ta, C.x, c1);
                c2 = fmaf(ta, C.y, c2);
                T = cov ? tin : T;
            }
            __syncwarp();
        }
        if (w.np > 1) {
            float *ps = pstate + static_cast<int64_t>(w.p) * PIECE_STATE + lane;
            ps[0] = T; ps[32] = c0; ps[64] = c1; ps[96] = c2;
        } else if (ix <= W && iy <= H) {
            float *p = image + 3 * (static_cast<int64_t>(iy) * (W + 1) + ix);
            p[0] = c0; p[1] = c1; p[2] = c2;
        }
    }
}

// tiles of several pieces: the carry of every piece (kept for the backward), and the pixel's colour
__global__ void __launch_bounds__(256)
k_tile_combine_fwd(const int32_t *__restrict__ pstart, int ntx, int ntiles, int W, int H, float *__restrict__ pstate,
                   float *__restrict__ image) {
    const int lane = threadIdx.x & 31;
    const int t = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (t >= ntiles) return;
    const int p0 = __ldg(pstart + t), np = __ldg(pstart + t + 1) - p0;
    if (np <= 1) return;
    float carry = 1.0f, c0 = 0.f, c1 = 0.f, c2 = 0.f;
    for (int k = 0; k < np; ++k) {
        float *ps = pstate + static_cast<int64_t>(p0 + k) * PIECE_STATE + lane;
        ps[128] = carry;
        c0 = fmaf(carry, ps[32], c0);
        c1 = fmaf(carry, ps[64], c1);
        c2 = fmaf(carry, ps[96], c2);
        carry *= ps[0];
    }
    const int ty = t / ntx, tx = t - ty * ntx;
    const int ix = (tx << TSX) + (lane & (TW - 1)), iy = (ty << TSY) + (lane >> TSX);
    if (ix <= W && iy <= H) {
        float *p = image + 3 * (static_cast<int64_t>(iy) * (W + 1) + ix);
        p[0] = c0; p[1] = c1; p[2] = c2;
    }
}

// tiles of several pieces: U after the last element of every piece, from the pieces behind it
__global__ void __launch_bounds__(256)
k_tile_combine_bwd(const int32_t *__restrict__ pstart, const float *__restrict__ gimg, int ntx, int ntiles, int W,
                   int H, float *__restrict__ pstate) {
    const int lane = threadIdx.x & 31;
    const int t = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (t >= ntiles) return;
    const int p0 = __ldg(pstart + t), np = __ldg(pstart + t + 1) - p0;
    if (np <= 1) return;
    const int ty = t / ntx, tx = t - ty * ntx;
    const int ix = (tx << TSX) + (lane & (TW - 1)), iy = (ty << TSY) + (lane >> TSX);
    float pg0 = 0.f, pg1 = 0.f, pg2 = 0.f;
    if (ix <= W && iy <= H) {
        const float *p = gimg + 3 * (static_cast<int64_t>(iy) * (W + 1) + ix);
        pg0 = __ldg(p); pg1 = __ldg(p + 1); pg2 = __ldg(p + 2);
    }
    float U = 0.0f;
    for (int k = np - 1; k >= 0; --k) {
        float *ps = pstate + static_cast<int64_t>(p0 + k) * PIECE_STATE + lane;
        ps[160] = U;
        U = fmaf(ps[0], U, pg0 * ps[32] + pg1 * ps[64] + pg2 * ps[96]);
    }
}

// sum of 8 values per lane over the 32 lanes in 9 shuffles (halving butterfly): afterwards the four lanes
// 4c' .. 4c'+3 hold component comp(c') = 4*bit2(c') + 2*bit1(c') + bit0(c'), c' = lane >> 2 with its bits read as
// (lane bit 4, lane bit 3, lane bit 2).  The order of the additions is fixed: bitwise reproducible.
__device__ __forceinline__ float reduce8(float (&v)[8], int lane) {
    const unsigned F = 0xffffffffu;
    float k4[4], k2[2];
    {
        const bool h = lane & 16;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float keep = h ? v[4 + i] : v[i], send = h ? v[i] : v[4 + i];
            k4[i] = keep + __shfl_xor_sync(F, send, 16);
        }
    }
    {
        const bool h = lane & 8;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const float keep = h ? k4[2 + i] : k4[i], send = h ? k4[i] : k4[2 + i];
            k2[i] = keep + __shfl_xor_sync(F, send, 8);
        }
    }
    const bool h = lane & 4;
    float s = (h ? k2[1] : k2[0]) + __shfl_xor_sync(F, h ? k2[0] : k2[1], 4);
    s += __shfl_xor_sync(F, s, 2);
    s += __shfl_xor_sync(F, s, 1);
    return s;
}

// partial[q] = {sum g dalpha, sum d, sum coef X0', sum coef X1', sum coef d0 d0, sum coef d0 d1, sum coef d1 d1, 0}
// over the pixels of pair q (coef = alpha dalpha; X' = EXP2_SCALE X), see gs_model.py:733-766; the constant factors
// -1/2 (d_Lambda) and EXP2_UNSCALE (d_mean) are applied to the per-Gaussian sums by k_tile_reduce
__global__ void __launch_bounds__(TILE_WARPS * 32, 4)
k_tile_backward(const int32_t *__restrict__ tstart, const int32_t *__restrict__ pgid_s, const int4 *__restrict__ rec,
                const int32_t *__restrict__ pstart, const int32_t *__restrict__ ptile, unsigned int *__restrict__ ticket,
                int piece, const float *__restrict__ tkeep, const float *__restrict__ pstate,
                const float *__restrict__ gimg, int ntx, int ntiles, int W, int H, float *__restrict__ partial) {
    __shared__ PairSlot slots[TILE_WARPS][32];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    PairSlot *sl = slots[wib];
    const int comp = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
    float *const pcomp = partial + comp;   // the component this lane's group ends up holding after reduce8
    const int npieces = __ldg(pstart + ntiles);
    Piece w;
    while (next_piece(ticket, tstart, pstart, ptile, npieces, piece, lane, w)) {
        const int t = w.t;
        const int64_t lo = w.lo, hi = w.hi;
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
            const float *ps = pstate + static_cast<int64_t>(w.p) * PIECE_STATE + lane;
            carry = __ldg(ps + 128);
            U = __ldg(ps + 160);
        }
        // batches of 32 pairs from the END of the piece: [bb, be), be = hi, hi - 32, ...
        int g1 = 0;
        RecRegs r = {};
        {
            const int64_t bb = (hi - 32 > lo) ? hi - 32 : lo;
            if (bb + lane < hi) r = load_rec(rec, __ldg(pgid_s + bb + lane));
            const int64_t b1 = (bb - 32 > lo) ? bb - 32 : lo;
            if (b1 + lane < bb) g1 = __ldg(pgid_s + b1 + lane);
        }
        // the kept T of the four pairs walked next, loaded one group (four pairs) ahead of their use, across batch
        // boundaries: every batch but the last one walked holds 32 pairs, so the groups never straddle a batch
        const float *tl = tkeep + lane;
        float tc[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) tc[j] = (hi - 1 - j >= lo) ? __ldcs(tl + (hi - 1 - j) * 32) : 0.0f;
        for (int64_t be = hi; be > lo;) {
            const int64_t bb = (be - 32 > lo) ? be - 32 : lo;
            const int m = static_cast<int>(be - bb);
            stage_pair(sl + lane, r, lane < m, tx, ty, x0, y0);
            // next batch [b1, bb), the one after [b2, b1)
            const int64_t b1 = (bb - 32 > lo) ? bb - 32 : lo;
            const int64_t b2 = (b1 - 32 > lo) ? b1 - 32 : lo;
            int g2 = 0;
            if (b2 + lane < b1) g2 = __ldg(pgid_s + b2 + lane);
            if (b1 + lane < bb) r = load_rec(rec, g1);
            g1 = g2;
            __syncwarp();
            for (int k0 = m - 1; k0 >= 0; k0 -= 4) {
                float tn[4];
                const int64_t pn = bb + k0 - 4;   // first pair of the next group
#pragma unroll
                for (int j = 0; j < 4; ++j) tn[j] = (pn - j >= lo) ? __ldcs(tl + (pn - j) * 32) : 0.0f;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int k = k0 - j;
                    if (k < 0) break;   // warp-uniform (last batch walked only)
                    const float4 A = sl[k].a, B = sl[k].b;
                    const float2 C = sl[k].c;
                    const bool cov = (sl[k].mask >> lane) & 1u;
                    const int q = sl[k].q;
                    const float T = carry * tc[j];
                    const PairEval e = eval_pair(A, B, px, py);
                    const bool alive = cov && (T * e.x != 0.0f);
                    const float alpha = 1.0f - e.x;
                    const float pgl = pg0 * B.w + pg1 * C.x + pg2 * C.y;
                    const float dalpha = alive ? T * pgl - T * U : 0.0f;
                    const float d = alive ? T * alpha * pgl : 0.0f;
                    U = cov ? fmaf(e.x, U, alive ? alpha * pgl : 0.0f) : U;   // U_{i-1} = w_i + x_i U_i
                    // per-pixel terms of gs_model.py:733-766 WITHOUT their constant factors (-1/2 for d_Lambda,
                    // EXP2_UNSCALE for d_mean because X0, X1 carry EXP2_SCALE): k_tile_reduce applies them to the sums
                    const float coef = B.z * e.gk * dalpha;
                    const float c0 = coef * e.d0, c1 = coef * e.d1;
                    float v[8] = {e.gk * dalpha, d, coef * e.X0, coef * e.X1, c0 * e.d0, c0 * e.d1, c1 * e.d1, 0.0f};
                    const float s = reduce8(v, lane);
                    if ((lane & 3) == 0) pcomp[static_cast<int64_t>(q) * 8] = s;
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) tc[j] = tn[j];
            }
            __syncwarp();
            be = bb;
        }
    }
}

// lane c of a group of 8 owns component c of a Gaussian's gradient sums and writes what derives from it:
// d_l[k] = (sum d) / l[k] is the reference's d / l (gs_model.py:763-766)
__device__ __forceinline__ void store_component(int c, float s, int64_t g, const float *__restrict__ l_d,
                                                float *__restrict__ g_mean, float *__restrict__ g_lam,
                                                float *__restrict__ g_opac, float *__restrict__ g_l) {
    // constant factors the backward walk left out of its per-pixel terms
    if (c == 2 || c == 3) s *= EXP2_UNSCALE;
    if (c >= 4 && c <= 6) s *= -0.5f;
    switch (c) {
        case 0: g_opac[g] = s; break;
        case 1:
            g_l[3 * g] = s / __ldg(l_d + 3 * g);
            g_l[3 * g + 1] = s / __ldg(l_d + 3 * g + 1);
            g_l[3 * g + 2] = s / __ldg(l_d + 3 * g + 2);
            break;
        case 2: g_mean[2 * g] = s; break;
        case 3: g_mean[2 * g + 1] = s; break;
        case 4: g_lam[4 * g] = s; break;
        case 5: g_lam[4 * g + 1] = s; g_lam[4 * g + 2] = s; break;
        case 6: g_lam[4 * g + 3] = s; break;
        default: break;
    }
}

constexpr int RED_BIG = 64;  // Gaussians with more pairs than this go to k_tile_reduce_big (one block each)

// 8 lanes per Gaussian: lane c adds component c of the Gaussian's partials in pair order.  Gaussians with more than
// RED_BIG pairs (boxes of thousands of pixels, bundled scene) are appended to `big` instead; the order of that
// list does not enter any float sum.
__global__ void __launch_bounds__(256)
k_tile_reduce(const float *__restrict__ partial, const int64_t *__restrict__ toff, const float *__restrict__ l_d,
              int64_t n, float *__restrict__ g_mean, float *__restrict__ g_lam, float *__restrict__ g_opac,
              float *__restrict__ g_l, unsigned int *__restrict__ nbig, int32_t *__restrict__ big) {
    const int c = threadIdx.x & 7;
    const int64_t g = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 3;
    if (g >= n) return;
    const int64_t b = __ldg(toff + g), e = __ldg(toff + g + 1);
    if (e - b > RED_BIG) {
        if (c == 0) big[atomicAdd(nbig, 1u)] = static_cast<int32_t>(g);
        return;
    }
    float s = 0.0f;
    int64_t q = b;
    for (; q + 4 <= e; q += 4) {
        const float v0 = __ldcs(partial + q * 8 + c), v1 = __ldcs(partial + (q + 1) * 8 + c);
        const float v2 = __ldcs(partial + (q + 2) * 8 + c), v3 = __ldcs(partial + (q + 3) * 8 + c);
        s += v0; s += v1; s += v2; s += v3;
    }
    for (; q < e; ++q) s += __ldcs(partial + q * 8 + c);
    store_component(c, s, g, l_d, g_mean, g_lam, g_opac, g_l);
}

// one block per big Gaussian: 32 groups of 8 lanes stride over its pairs (group j takes pairs j, j+32, ...), the 32
// group sums are added in group order — a fixed order, bitwise reproducible
__global__ void __launch_bounds__(256)
k_tile_reduce_big(const float *__restrict__ partial, const int64_t *__restrict__ toff, const float *__restrict__ l_d,
                  float *__restrict__ g_mean, float *__restrict__ g_lam, float *__restrict__ g_opac,
                  float *__restrict__ g_l, const unsigned int *__restrict__ nbig, const int32_t *__restrict__ big) {
    __shared__ float sums[32][8];
    const int c = threadIdx.x & 7, j = threadIdx.x >> 3;
    const unsigned int nb = *nbig;
    for (unsigned int i = blockIdx.x; i < nb; i += gridDim.x) {
        const int64_t g = __ldg(big + i);
        const int64_t b = __ldg(toff + g), e = __ldg(toff + g + 1);
        float s = 0.0f;
        int64_t q = b + j;
        for (; q + 96 < e; q += 128) {
            const float v0 = __ldcs(partial + q * 8 + c), v1 = __ldcs(partial + (q + 32) * 8 + c);
            const float v2 = __ldcs(partial + (q + 64) * 8 + c), v3 = __ldcs(partial + (q + 96) * 8 + c);
            s += v0; s += v1; s += v2; s += v3;
        }
        for (; q < e; q += 32) s += __ldcs(partial + q * 8 + c);
        sums[j][c] = s;
        __syncthreads();
        if (j == 0) {
            float t = 0.0f;
#pragma unroll
            for (int k = 0; k < 32; ++k) t += sums[k][c];
            store_component(c, t, g, l_d, g_mean, g_lam, g_opac, g_l);
        }
        __syncthreads();
    }
}

struct BinLayout {
    size_t ptile, pgid, ptile_s, cub, total, cub_bytes;
};
BinLayout bin_layout(int64_t P, int ntiles) {
    BinLayout L;
    const size_t pb = align256(static_cast<size_t>(P > 0 ? P : 1) * 4);
    size_t a = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, a, static_cast<const int32_t *>(nullptr), static_cast<int32_t *>(nullptr),
                                    static_cast<const int32_t *>(nullptr), static_cast<int32_t *>(nullptr),
                                    P > 0 ? P : 1, 0, 24);
    size_t b = 0;   // the scan of the tiles' piece counts shares the sort's scratch
    cub::DeviceScan::InclusiveSum(nullptr, b, static_cast<const int32_t *>(nullptr), static_cast<int32_t *>(nullptr),
                                  ntiles > 0 ? ntiles : 1);
    L.cub_bytes = align256(std::max(a, b) + 256);
    L.ptile = 0;
    L.pgid = L.ptile + pb;
    L.ptile_s = L.pgid + pb;
    L.cub = L.ptile_s + pb;
    L.total = L.cub + L.cub_bytes;
    return L;
}

// persistent grid of the walk kernels: every resident warp slot of the device, never more warps than tiles
unsigned walk_grid(const void *kernel, int64_t nwork) {
    static const void *known[2] = {nullptr, nullptr};
    static unsigned slots[2] = {0, 0};  // resident blocks of the (two) walk kernels, queried once
    int i = (known[0] == kernel || known[0] == nullptr) ? 0 : 1;
    if (known[i] != kernel) {
        int dev = 0, sms = 148, per_sm = 4;
        if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, TILE_WARPS * 32, 0) != cudaSuccess ||
            per_sm < 1)
            per_sm = 1;
        slots[i] = static_cast<unsigned>(sms * per_sm);
        known[i] = kernel;
    }
    return std::max(1u, std::min(blocks_for(nwork, TILE_WARPS), slots[i]));
}

int g_piece = 128;  // pairs per piece (gcp_tile_set_piece_pairs)

inline bool bad_image(int W, int H) { return W < 0 || H < 0 || W >= 32768 || H >= 32768; }
inline int tiles_x(int W) { return (W + TW) >> TSX; }  // ceil((W+1)/TW): pixels 0..W inclusive (gs_model.py:505)
inline int tiles_y(int H) { return (H + TH) >> TSY; }

// views into the piece_plan buffer: piece_start[ntiles+1] | piece_tile[cap] | 4 counter words
struct PlanView {
    const int32_t *pstart, *ptile;
    unsigned int *tickets;
    int64_t cap;
};
PlanView plan_view(int32_t *piece_plan, int64_t P, int W, int H) {
    const int ntiles = tiles_x(W) * tiles_y(H);
    const int64_t cap = gcp_tile_piece_cap(P, W, H);
    return PlanView{piece_plan, piece_plan + ntiles + 1,
                    reinterpret_cast<unsigned int *>(piece_plan + ntiles + 1 + cap), cap};
}
}  // namespace

extern "C" {

int gcp_tile_width(void) { return TW; }
int gcp_tile_height(void) { return TH; }
int gcp_tile_num_tiles(int W, int H) { return bad_image(W, H) ? 0 : tiles_x(W) * tiles_y(H); }
/* longest run of one tile's pairs a single warp walks (multiple of 32); longer lists are cut into pieces */
int gcp_tile_set_piece_pairs(int pairs) {
    if (pairs < 32 || pairs > (1 << 20) || (pairs & 31)) return GCP_ERR_INVALID_ARG;
    g_piece = pairs;
    return GCP_OK;
}
int gcp_tile_piece_pairs(void) { return g_piece; }
int64_t gcp_tile_piece_cap(int64_t P, int W, int H) {
    return (P < 0 || bad_image(W, H)) ? 0 : tiles_x(W) * tiles_y(H) + P / g_piece + 1;
}
/* piece_plan i32[]: piece_start[ntiles+1] | piece_tile[cap] | 4 counter words */
int64_t gcp_tile_plan_ints(int64_t P, int W, int H) {
    return (P < 0 || bad_image(W, H)) ? 0 : tiles_x(W) * tiles_y(H) + 1 + gcp_tile_piece_cap(P, W, H) + 4;
}
int64_t gcp_tile_state_floats(int64_t P, int W, int H) { return gcp_tile_piece_cap(P, W, H) * PIECE_STATE; }

size_t gcp_tile_prepare_bytes(int64_t n) {
    size_t a = 0;
    cub::DeviceScan::InclusiveSum(nullptr, a, static_cast<const int64_t *>(nullptr), static_cast<int64_t *>(nullptr),
                                  n > 0 ? n : 1);
    return a + 256;
}

int gcp_tile_prepare(const int32_t *sp, const int32_t *ep, int64_t n, int W, int H, int64_t *toff, int64_t *totals,
                     void *temp, size_t temp_bytes, gcp_stream_t stream) {
    if (n < 0 || bad_image(W, H) || !toff || !totals) return GCP_ERR_INVALID_ARG;
    auto st = reinterpret_cast<cudaStream_t>(stream);
    cudaError_t e = cudaMemsetAsync(toff, 0, 8, st);
    if (e != cudaSuccess) return static_cast<int>(e);
    if (n > 0) {
        if (!sp || !ep || !temp) return GCP_ERR_INVALID_ARG;
        auto pairs = thrust::make_transform_iterator(thrust::counting_iterator<int64_t>(0),
                                                     TilePairCount{sp, ep, W, H});
        size_t need = 0;
        cub::DeviceScan::InclusiveSum(nullptr, need, pairs, toff + 1, n);
        if (temp_bytes < need) return GCP_ERR_WORKSPACE;
        size_t tb = temp_bytes;
        e = cub::DeviceScan::InclusiveSum(temp, tb, pairs, toff + 1, n, st);
        if (e != cudaSuccess) return static_cast<int>(e);
    }
    k_tile_totals<<<1, 1, 0, st>>>(toff, n, totals);
    return static_cast<int>(cudaGetLastError());
}

int gcp_tile_pack(const float *mean, const float *lam, const float *opac, const float *l_d, const int32_t *sp,
                  const int32_t *ep, const int64_t *toff, int64_t n, int W, int H, int32_t *rec,
                  gcp_stream_t stream) {
    if (n < 0 || bad_image(W, H)) return GCP_ERR_INVALID_ARG;
    if (n == 0) return GCP_OK;
    if (!mean || !lam || !opac || !l_d || !sp || !ep || !toff || !rec) return GCP_ERR_INVALID_ARG;
    if (reinterpret_cast<uintptr_t>(rec) & 31) return GCP_ERR_INVALID_ARG;
    if ((reinterpret_cast<uintptr_t>(mean) & 7) || (reinterpret_cast<uintptr_t>(lam) & 15)) return GCP_ERR_INVALID_ARG;
    k_tile_pack<<<blocks_for(n, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        mean, lam, opac, l_d, sp, ep, toff, n, W, H, reinterpret_cast<int4 *>(rec));
    return static_cast<int>(cudaGetLastError());
}

size_t gcp_tile_bin_bytes(int64_t P, int W, int H) {
    return (P < 0 || bad_image(W, H)) ? 0 : bin_layout(P, tiles_x(W) * tiles_y(H)).total;
}

static int tile_bin_impl(const int32_t *sp, const int32_t *ep, const int64_t *toff, int64_t n, int64_t P, int W,
                         int H, int32_t *tile_start, int32_t *piece_plan, int32_t *pair_gid, void *temp,
                         size_t temp_bytes, gcp_stream_t stream, bool speculative) {
    if (n < 0 || P < 0 || P >= (int64_t(1) << 31) - 64 || bad_image(W, H)) return GCP_ERR_INVALID_ARG;
    if (!tile_start || !piece_plan || !temp || (P > 0 && (!sp || !ep || !toff || !pair_gid)))
        return GCP_ERR_INVALID_ARG;
    const int ntx = tiles_x(W), ntiles = ntx * tiles_y(H);
    const BinLayout L = bin_layout(P, ntiles);
    if (temp_bytes < L.total) return GCP_ERR_WORKSPACE;
    auto st = reinterpret_cast<cudaStream_t>(stream);
    unsigned char *t = static_cast<unsigned char *>(temp);
    int32_t *ptile = reinterpret_cast<int32_t *>(t + L.ptile), *pgid = reinterpret_cast<int32_t *>(t + L.pgid);
    int32_t *ptile_s = reinterpret_cast<int32_t *>(t + L.ptile_s);
    cudaError_t e;
    if (P > 0) {
        if (P <= 6 * n)   // small boxes: a thread per Gaussian; boxes of many tiles: parallel over the pairs
            k_tile_pairs<<<blocks_for(n, 256), 256, 0, st>>>(sp, ep, toff, n, P, W, H, ntx, ptile, pgid);
        else
            k_tile_pairs_flat<<<blocks_for(P, 256 * CHP), 256, 0, st>>>(sp, ep, toff, n, P, W, H, ntx, ptile, pgid);
        if (speculative)   // P is a capacity here: mark the slots behind the real pairs
            k_tile_pad<<<blocks_for(P / 4 + 1, 256, 1024), 256, 0, st>>>(toff, n, P, ntiles, ptile, pgid);
        size_t cb = L.cub_bytes;
        // stable LSD radix sort on the tile bits only: inside a tile the Gaussians keep their (depth) order
        e = cub::DeviceRadixSort::SortPairs(t + L.cub, cb, ptile, ptile_s, pgid, pair_gid, P, 0, key_bits(ntiles), st);
        if (e != cudaSuccess) return static_cast<int>(e);
    }
    k_tile_start<<<blocks_for(P + 1, 256), 256, 0, st>>>(ptile_s, P, ntiles, tile_start);
    // the pieces: exclusive offsets per tile, then the tile of every piece
    int32_t *pstart = piece_plan, *piece_tile = piece_plan + ntiles + 1;
    e = cudaMemsetAsync(pstart, 0, sizeof(int32_t), st);
    if (e != cudaSuccess) return static_cast<int>(e);
    auto counts = thrust::make_transform_iterator(thrust::counting_iterator<int32_t>(0), PieceCount{tile_start, g_piece});
    size_t need = 0;
    cub::DeviceScan::InclusiveSum(nullptr, need, counts, pstart + 1, ntiles);
    if (need > L.cub_bytes) return GCP_ERR_WORKSPACE;
    size_t cb = L.cub_bytes;
    e = cub::DeviceScan::InclusiveSum(t + L.cub, cb, counts, pstart + 1, ntiles, st);
    if (e != cudaSuccess) return static_cast<int>(e);
    k_tile_pieces<<<blocks_for(ntiles, 256), 256, 0, st>>>(pstart, ntiles, piece_tile);
    return static_cast<int>(cudaGetLastError());
}

int gcp_tile_bin(const int32_t *sp, const int32_t *ep, const int64_t *toff, int64_t n, int64_t P, int W, int H,
                 int32_t *tile_start, int32_t *piece_plan, int32_t *pair_gid, void *temp, size_t temp_bytes,
                 gcp_stream_t stream) {
    return tile_bin_impl(sp, ep, toff, n, P, W, H, tile_start, piece_plan, pair_gid, temp, temp_bytes, stream, false);
}

int gcp_tile_bin_speculative(const int32_t *sp, const int32_t *ep, const int64_t *toff, int64_t n, int64_t cap, int W,
                             int H, int32_t *tile_start, int32_t *piece_plan, int32_t *pair_gid, void *temp,
                             size_t temp_bytes, gcp_stream_t stream) {
    return tile_bin_impl(sp, ep, toff, n, cap, W, H, tile_start, piece_plan, pair_gid, temp, temp_bytes, stream, true);
}

int gcp_tile_render(const int32_t *tile_start, int32_t *piece_plan, const int32_t *pair_gid, const int32_t *rec,
                    int64_t P, int W, int H, float *image, float *t_keep, float *piece_state, gcp_stream_t stream) {
    if (P < 0 || bad_image(W, H) || !tile_start || !piece_plan || !image || !piece_state) return GCP_ERR_INVALID_ARG;
    if (P > 0 && (!pair_gid || !rec)) return GCP_ERR_INVALID_ARG;   // t_keep may be NULL: forward only
    auto st = reinterpret_cast<cudaStream_t>(stream);
    const int ntx = tiles_x(W), ntiles = ntx * tiles_y(H);
    const PlanView pv = plan_view(piece_plan, P, W, H);
    cudaError_t e = cudaMemsetAsync(pv.tickets, 0, sizeof(unsigned int), st);
    if (e != cudaSuccess) return static_cast<int>(e);
    const unsigned grid = walk_grid(reinterpret_cast<const void *>(k_tile_render<true>), pv.cap);
    if (t_keep != nullptr)
        k_tile_render<true><<<grid, TILE_WARPS * 32, 0, st>>>(tile_start, pair_gid, reinterpret_cast<const int4 *>(rec),
                                                              pv.pstart, pv.ptile, pv.tickets, g_piece, ntx, ntiles, W,
                                                              H, image, t_keep, piece_state);
    else
        k_tile_render<false><<<grid, TILE_WARPS * 32, 0, st>>>(tile_start, pair_gid,
                                                               reinterpret_cast<const int4 *>(rec), pv.pstart, pv.ptile,
                                                               pv.tickets, g_piece, ntx, ntiles, W, H, image, t_keep,
                                                               piece_state);
    k_tile_combine_fwd<<<blocks_for(ntiles, 8), 256, 0, st>>>(pv.pstart, ntx, ntiles, W, H, piece_state, image);
    return static_cast<int>(cudaGetLastError());
}

int gcp_tile_backward(const int32_t *tile_start, int32_t *piece_plan, const int32_t *pair_gid, const int32_t *rec,
                      const float *t_keep, float *piece_state, const float *grad_image, int64_t P, int W, int H,
                      float *partial, gcp_stream_t stream) {
    if (P < 0 || bad_image(W, H) || !tile_start || !piece_plan || !grad_image) return GCP_ERR_INVALID_ARG;
    if (P == 0) return GCP_OK;
    if (!pair_gid || !rec || !t_keep || !partial || !piece_state) return GCP_ERR_INVALID_ARG;
    auto st = reinterpret_cast<cudaStream_t>(stream);
    const int ntx = tiles_x(W), ntiles = ntx * tiles_y(H);
    const PlanView pv = plan_view(piece_plan, P, W, H);
    cudaError_t e = cudaMemsetAsync(pv.tickets + 1, 0, sizeof(unsigned int), st);
    if (e != cudaSuccess) return static_cast<int>(e);
    k_tile_combine_bwd<<<blocks_for(ntiles, 8), 256, 0, st>>>(pv.pstart, grad_image, ntx, ntiles, W, H, piece_state);
    k_tile_backward<<<walk_grid(reinterpret_cast<const void *>(k_tile_backward), pv.cap), TILE_WARPS * 32, 0, st>>>(
        tile_start, pair_gid, reinterpret_cast<const int4 *>(rec), pv.pstart, pv.ptile, pv.tickets + 1, g_piece,
        t_keep, piece_state, grad_image, ntx, ntiles, W, H, partial);
    return static_cast<int>(cudaGetLastError());
}

size_t gcp_tile_reduce_bytes(int64_t n) { return n < 0 ? 0 : 256 + align256(static_cast<size_t>(n) * 4); }

int gcp_tile_reduce(const float *partial, const int64_t *toff, const float *l_d, int64_t n, float *g_mean,
                    float *g_lam, float *g_opac, float *g_l, void *temp, size_t temp_bytes, gcp_stream_t stream) {
    if (n < 0) return GCP_ERR_INVALID_ARG;
    if (n == 0) return GCP_OK;
    if (!toff || !l_d || !g_mean || !g_lam || !g_opac || !g_l) return GCP_ERR_INVALID_ARG;
    if (!temp || temp_bytes < gcp_tile_reduce_bytes(n) || (reinterpret_cast<uintptr_t>(temp) & 3))
        return GCP_ERR_WORKSPACE;
    auto st = reinterpret_cast<cudaStream_t>(stream);
    unsigned int *nbig = static_cast<unsigned int *>(temp);
    int32_t *big = reinterpret_cast<int32_t *>(static_cast<unsigned char *>(temp) + 256);
    cudaError_t e = cudaMemsetAsync(nbig, 0, sizeof(unsigned int), st);
    if (e != cudaSuccess) return static_cast<int>(e);
    k_tile_reduce<<<blocks_for(n, 32), 256, 0, st>>>(partial, toff, l_d, n, g_mean, g_lam, g_opac, g_l, nbig, big);
    k_tile_reduce_big<<<blocks_for(n, 1, 148 * 8), 256, 0, st>>>(partial, toff, l_d, g_mean, g_lam, g_opac, g_l, nbig,
                                                                 big);
    return static_cast<int>(cudaGetLastError());
}

}  // extern "C"
