// gcp_bwd.cuh — backward of the segmented cumprod: one streaming pass, division-free, wait-free.
//
// Replaces grouped_cumprod_backward_kernel
//   (/root/reference/cuda_kernel/grouped_cumprod_backward.cu:9-41, launcher :43-65),
// whose per-thread serial loop costs O(sum L^2) and divides by x.
//
//   grad_in[i] = E_i * S_i
//   E_i = prod_{j<i, same segment} x_j          forward exclusive product (1 at a head)
//   S_i = g_i + x_{i+1} * S_{i+1}               reverse recurrence (S = g at a tail)
//
// S is a reverse segmented scan of affine maps (a_i, b_i) = (tail ? 0 : x_{i+1}, g_i);
// a tail's a = 0 is the segment reset.  E needs the other direction: inside a tile it is
// recomputed from x by the forward's segmented warp scan (no extra traffic), and across the
// tile's left edge it is the forward output itself, E(first element) = y[base-1] — ONE value
// of `param_cumprod` per tile, an argument the reference op already receives.  So the pass
// reads x, grad_out, inv (12 B/elem) and writes grad_in (4 B/elem): 16 B/elem.
//
// Cross-tile carry of S, mirrored from gcp_fwd.cuh and equally wait-free:
//   K1 (k_bwd_blk in gcp_blk.cuh / k_bwd_ldg) resolves R = S(first element after the tile) from the HALO,
//      the 128 elements after the tile.  Unresolved tiles store provisional values for
//      their trailing run (the elements after the tile's last tail), publish a carry
//      descriptor (TERM R | AGG (a,b)) and a fix-up request {needs, trail start}.
//   Fix-up (bwd_fix_tile) one warp per unresolved tile, once all descriptors are complete:
//      walks forward over the descriptors to the nearest TERM / already-fixed tile and
//      recomputes the trailing run (reads x, g and y there: E_i = y[i-1]).  Second phase of the
//      same launch in the persistent kernel (grid barrier), separate kernel K2 on the LDG path.
#pragma once
#include "gcp_device.cuh"
#include "gcp_fwd.cuh"

namespace gcp {

template <int WARPS>
struct BwdShared {
    float wv[WARPS];     // forward product aggregates
    uint32_t wf[WARPS];
    float wa[WARPS];     // reverse affine aggregates
    float wb[WARPS];
    int32_t lt[WARPS];   // offset of the last tail inside the warp span (-1 if none)
    uint32_t res;        // LDG kernel: halo result
    float rn;
    uint32_t tile;
};

// All 32 lanes of one warp: resolve S at the first element after the tile (position `end`)
// from the 128 elements that follow.  True when the tile's last element is a tail (R
// irrelevant) or the composite of the window annihilates (a tail, or an exact zero, inside
// it); then R = S(end).  Also returns the two halo values the consumers need:
// inext = inv[end], xnext = x[end] (-1 / 0 beyond the array).
__device__ __forceinline__ bool halo_suffix(const float *__restrict__ x, const float *__restrict__ g,
                                            const int32_t *__restrict__ inv, int64_t end, int64_t n, int lane,
                                            bool vec, float &R, int32_t &inext, float &xnext) {
    if (end >= n) {
        R = 0.0f;
        inext = -1;
        xnext = 0.0f;
        return true;
    }
    const int64_t h0 = end + lane * 4;
    float xv[4], gv[4];
    int32_t iv[4];
    if (vec && h0 + 3 < n) {
        const float4 a = __ldg(reinterpret_cast<const float4 *>(x + h0));
        const float4 b = __ldg(reinterpret_cast<const float4 *>(g + h0));
        const int4 c = __ldg(reinterpret_cast<const int4 *>(inv + h0));
        xv[0] = a.x; xv[1] = a.y; xv[2] = a.z; xv[3] = a.w;
        gv[0] = b.x; gv[1] = b.y; gv[2] = b.z; gv[3] = b.w;
        iv[0] = c.x; iv[1] = c.y; iv[2] = c.z; iv[3] = c.w;
    } else {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const bool in = h0 + e < n;
            xv[e] = in ? __ldg(x + h0 + e) : 1.0f;
            gv[e] = in ? __ldg(g + h0 + e) : 0.0f;
            iv[e] = in ? __ldg(inv + h0 + e) : -1;  // padding differs from every id: the last element is a tail
        }
    }
    const int32_t ilast = __ldg(inv + end - 1);
    int32_t qn = __shfl_down_sync(0xffffffffu, iv[0], 1);
    float xq = __shfl_down_sync(0xffffffffu, xv[0], 1);
    const bool known = lane < 31;  // what follows the window is unknown
    if (!known) xq = 1.0f;
    inext = __shfl_sync(0xffffffffu, iv[0], 0);
    xnext = __shfl_sync(0xffffffffu, xv[0], 0);
    Affine m = Affine{(known && qn != iv[3]) ? 0.0f : xq, gv[3]};
    m = compose(Affine{(iv[3] != iv[2]) ? 0.0f : xv[3], gv[2]}, m);
    m = compose(Affine{(iv[2] != iv[1]) ? 0.0f : xv[2], gv[1]}, m);
    m = compose(Affine{(iv[1] != iv[0]) ? 0.0f : xv[1], gv[0]}, m);
    m = warp_compose_all(m, lane);  // everything after the first tail is annihilated
    if (ilast != inext) {
        R = 0.0f;
        return true;
    }
    R = m.b;
    return m.a == 0.0f;
}

// halo_suffix split in two for a software-pipelined producer (issue in iteration i, finish in i+1).
// HQ = float4 quads per lane: the window is the 128*HQ elements after the tile.
template <int HQ>
struct HaloSuffixRegs {
    float xv[4 * HQ], gv[4 * HQ];
    int32_t iv[4 * HQ];
    int32_t ilast;
    bool beyond;  // end >= n: nothing follows the tile
};
template <int HQ>
__device__ __forceinline__ void halo_suffix_issue(const float *__restrict__ x, const float *__restrict__ g,
                                                  const int32_t *__restrict__ inv, int64_t end, int64_t n, int lane,
                                                  bool full_window, HaloSuffixRegs<HQ> &r) {
    r.beyond = end >= n;
    if (r.beyond) return;
    r.ilast = __ldg(inv + end - 1);
    if (!full_window) {  // halo resolution off: only the two values the consumers need
        r.xv[0] = __ldg(x + end);
        r.iv[0] = __ldg(inv + end);
        return;
    }
    const int64_t h0 = end + lane * (4 * HQ);
#pragma unroll
    for (int c = 0; c < HQ; ++c) {
        const int64_t hc = h0 + 4 * c;
        if (hc + 3 < n) {
            const float4 a = __ldg(reinterpret_cast<const float4 *>(x + hc));
            const float4 b = __ldg(reinterpret_cast<const float4 *>(g + hc));
            const int4 k = __ldg(reinterpret_cast<const int4 *>(inv + hc));
            r.xv[4 * c] = a.x; r.xv[4 * c + 1] = a.y; r.xv[4 * c + 2] = a.z; r.xv[4 * c + 3] = a.w;
            r.gv[4 * c] = b.x; r.gv[4 * c + 1] = b.y; r.gv[4 * c + 2] = b.z; r.gv[4 * c + 3] = b.w;
            r.iv[4 * c] = k.x; r.iv[4 * c + 1] = k.y; r.iv[4 * c + 2] = k.z; r.iv[4 * c + 3] = k.w;
        } else {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const bool in = hc + e < n;
                r.xv[4 * c + e] = in ? __ldg(x + hc + e) : 1.0f;
                r.gv[4 * c + e] = in ? __ldg(g + hc + e) : 0.0f;
                r.iv[4 * c + e] = in ? __ldg(inv + hc + e) : -1;  // padding differs from every id: a tail before it
            }
        }
    }
}
template <int HQ>
__device__ __forceinline__ bool halo_suffix_finish(const HaloSuffixRegs<HQ> &r, int lane, bool full_window, float &R,
                                                   int32_t &inext, float &xnext) {
    if (r.beyond) {
        R = 0.0f;
        inext = -1;
        xnext = 0.0f;
        return true;
    }
    if (!full_window) {
        R = 0.0f;
        inext = r.iv[0];
        xnext = r.xv[0];
        return false;
    }
    constexpr int E = 4 * HQ;
    int32_t qn = __shfl_down_sync(0xffffffffu, r.iv[0], 1);
    float xq = __shfl_down_sync(0xffffffffu, r.xv[0], 1);
    const bool known = lane < 31;  // what follows the window is unknown
    if (!known) xq = 1.0f;
    inext = __shfl_sync(0xffffffffu, r.iv[0], 0);
    xnext = __shfl_sync(0xffffffffu, r.xv[0], 0);
    Affine m = Affine{(known && qn != r.iv[E - 1]) ? 0.0f : xq, r.gv[E - 1]};
#pragma unroll
    for (int e = E - 1; e >= 1; --e) m = compose(Affine{(r.iv[e] != r.iv[e - 1]) ? 0.0f : r.xv[e], r.gv[e - 1]}, m);
    m = warp_compose_all(m, lane);  // everything after the first tail is annihilated
    if (r.ilast != inext) {
        R = 0.0f;
        return true;
    }
    R = m.b;
    return m.a == 0.0f;
}

// Everything after x / g / inv of the tile are in registers.
//   iprev : inv of the element before this warp's span (used by lane 0), -1 if none
//   inext : inv of the element after this warp's span (used by lane 0!), -1 if none
//   xnext : x of the element after this warp's span (used by lane 0)
//   y_prev: y[base-1] (forward inclusive product just before the tile), any value if base == 0
//   resolved/rn : CTA-uniform halo result (TMA: from the producer; LDG: read from sh)
template <int WARPS, int ROWS, bool HALO_IN_SH>
__device__ __forceinline__ void bwd_tile_body(const float (&x)[ROWS][4], const float (&g)[ROWS][4],
                                              const int32_t (&iv)[ROWS][4], int32_t iprev, int32_t inext,
                                              float xnext, float y_prev, bool resolved, float rn, uint32_t tile,
                                              int64_t base, int64_t n, float *__restrict__ gin, bool out_vec,
                                              uint32_t epoch, uint32_t *__restrict__ hdr,
                                              uint64_t *__restrict__ desc, uint32_t *__restrict__ ulist,
                                              uint32_t *__restrict__ ulist2, BwdShared<WARPS> *sh, int warp,
                                              int lane) {
    static_assert(WARPS < 32, "cross-warp step uses one lane per warp (and lane WARPS as the identity)");
    constexpr int TILE = WARPS * ROWS * 128;
    const uint32_t lanes_lt = (1u << lane) - 1u;
    const uint32_t lanes_le = lanes_lt | (1u << lane);
    // ---- tail bits (one rotate-shuffle of inv per row), head bits derived from them,
    //      x of the next element, last tail of the warp span ----
    uint32_t hm = 0u, tm = 0u;
    float xn[ROWS];
    uint32_t stop_room = 0u;  // 5 bits per row: (first lane >= me holding a tail) - lane
    uint32_t carry_h = (iv[0][0] != iprev) ? 1u : 0u;  // only lane 0's value is used
    int32_t lt = -1;
    const int src_lane = (lane + 1) & 31;
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
        // lane 0 lends the first element of the NEXT row (or the warp halo) to lane 31
        constexpr int RL = ROWS - 1;
        const int rnx = (r + 1 < ROWS) ? r + 1 : RL;
        const int32_t lend_i = (r + 1 < ROWS) ? iv[rnx][0] : inext;
        const float lend_x = (r + 1 < ROWS) ? x[rnx][0] : xnext;
        const int32_t q = __shfl_sync(0xffffffffu, lane == 0 ? lend_i : iv[r][0], src_lane);
        xn[r] = __shfl_sync(0xffffffffu, lane == 0 ? lend_x : x[r][0], src_lane);
        const uint32_t t = (iv[r][1] != iv[r][0] ? 1u : 0u) | (iv[r][2] != iv[r][1] ? 2u : 0u) |
                           (iv[r][3] != iv[r][2] ? 4u : 0u) | (q != iv[r][3] ? 8u : 0u);
        const uint32_t m3 = __ballot_sync(0xffffffffu, (t & 8u) != 0u);  // lanes whose LAST element is a tail
        const uint32_t mt = __ballot_sync(0xffffffffu, t != 0u);         // lanes holding any tail
        const uint32_t h0 = lane ? ((m3 >> (lane - 1)) & 1u) : carry_h;
        carry_h = m3 >> 31;
        hm |= (h0 | ((t & 7u) << 1)) << (4 * r);
        tm |= t << (4 * r);
        const uint32_t ahead = mt & ~lanes_lt;  // tails at or after my lane
        const uint32_t stop = ahead ? static_cast<uint32_t>(__ffs(ahead) - 1) : 31u;
        stop_room |= (stop - lane) << (5 * r);
        if (mt) {  // warp-uniform
            const int l1 = 31 - __clz(mt);
            const uint32_t t1 = __shfl_sync(0xffffffffu, t, l1);
            lt = r * 128 + l1 * 4 + (31 - __clz(t1));
        }
    }
    // ---- pass 1: per-lane aggregates (forward product after the last head; reverse affine
    //      composite up to the first tail) ----
    float fagg[ROWS];
    Affine ragg[ROWS];
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
        const uint32_t h = (hm >> (4 * r)) & 15u;
        const uint32_t t = (tm >> (4 * r)) & 15u;
        float p = x[r][0];
        p = (h & 2u) ? x[r][1] : p * x[r][1];
        p = (h & 4u) ? x[r][2] : p * x[r][2];
        p = (h & 8u) ? x[r][3] : p * x[r][3];
        fagg[r] = p;
        float A = (t & 8u) ? 0.0f : xn[r];
        float B = g[r][3];
        B = (t & 4u) ? g[r][2] : fmaf(x[r][3], B, g[r][2]);
        A = (t & 4u) ? 0.0f : x[r][3] * A;
        B = (t & 2u) ? g[r][1] : fmaf(x[r][2], B, g[r][1]);
        A = (t & 2u) ? 0.0f : x[r][2] * A;
        B = (t & 1u) ? g[r][0] : fmaf(x[r][1], B, g[r][0]);
        A = (t & 1u) ? 0.0f : x[r][1] * A;
        ragg[r] = Affine{A, B};
    }
    // ---- forward segmented warp scan of the products ----
    float cv[ROWS];
    uint32_t cf, wf, fh;
    float wv;
    warp_seg_scan_rows<OP_MUL, ROWS>(fagg, hm, lane, cv, cf, wv, wf, fh);
    // ---- reverse warp scan of the affine maps, masked by the tail ballot: a lane only ever
    //      combines with lanes of its own segment (plus the lane that holds the tail) ----
    Affine sx[ROWS];
    Affine rs = affine_id();
#pragma unroll
    for (int r = ROWS - 1; r >= 0; --r) {
        const int room = static_cast<int>((stop_room >> (5 * r)) & 31u);
        float A = ragg[r].a, B = ragg[r].b;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const float ta = __shfl_down_sync(0xffffffffu, A, d);
            const float tb = __shfl_down_sync(0xffffffffu, B, d);
            if (d <= room) {
                B = fmaf(A, tb, B);
                A = A * ta;
            }
        }
        Affine exc;
        exc.a = __shfl_down_sync(0xffffffffu, A, 1);
        exc.b = __shfl_down_sync(0xffffffffu, B, 1);
        if (lane == 31) exc = affine_id();
        Affine row;
        row.a = __shfl_sync(0xffffffffu, A, 0);
        row.b = __shfl_sync(0xffffffffu, B, 0);
        sx[r] = compose(exc, rs);
        rs = compose(row, rs);
    }
    if (lane == 0) {
        sh->wv[warp] = wv;
        sh->wf[warp] = wf;
        sh->wa[warp] = rs.a;
        sh->wb[warp] = rs.b;
        sh->lt[warp] = lt;
    }
    named_bar_sync<WARPS * 32>(1);
    if (HALO_IN_SH) {
        resolved = sh->res != 0u;
        rn = sh->rn;
    }
    // ---- across warps, one lane per warp: forward prefix of the earlier warps, reverse suffix
    //      of the later warps, tile aggregate, position after the tile's last tail ----
    const bool wl = lane < WARPS;
    float jv = wl ? sh->wv[lane] : 1.0f;
    const uint32_t jf = wl ? sh->wf[lane] : 0u;
    Affine jm = wl ? Affine{sh->wa[lane], sh->wb[lane]} : affine_id();
    const int32_t jl = wl ? sh->lt[lane] : -1;
    const uint32_t fm = __ballot_sync(0xffffffffu, jf != 0u);
    {
        const int start = max(31 - __clz(fm & lanes_le), 0);
#pragma unroll
        for (int d = 1; d < WARPS; d <<= 1) {
            const float tv = __shfl_up_sync(0xffffffffu, jv, d);
            if (lane - d >= start) jv = tv * jv;
        }
#pragma unroll
        for (int d = 1; d < WARPS; d <<= 1) {
            Affine q;
            q.a = __shfl_down_sync(0xffffffffu, jm.a, d);
            q.b = __shfl_down_sync(0xffffffffu, jm.b, d);
            if (lane + d < 32) jm = compose(jm, q);
        }
    }
    const float wp_v = __shfl_sync(0xffffffffu, jv, warp > 0 ? warp - 1 : 0);  // used only if warp > 0
    const bool wp_f = (fm & ((1u << warp) - 1u)) != 0u;
    Affine ws, ta;
    ws.a = __shfl_sync(0xffffffffu, jm.a, warp + 1 < 32 ? warp + 1 : 31);  // lanes >= WARPS hold the identity
    ws.b = __shfl_sync(0xffffffffu, jm.b, warp + 1 < 32 ? warp + 1 : 31);
    ta.a = __shfl_sync(0xffffffffu, jm.a, 0);
    ta.b = __shfl_sync(0xffffffffu, jm.b, 0);
    const uint32_t lm = __ballot_sync(0xffffffffu, jl >= 0);
    uint32_t trail = 0u;
    if (lm) {
        const int jw = 31 - __clz(lm);
        trail = static_cast<uint32_t>(jw * ROWS * 128 + __shfl_sync(0xffffffffu, jl, jw) + 1);
    }
    // ---- publish the carry descriptor (+ fix-up request when the halo did not resolve R) ----
    if (warp == 0 && lane == 0) {
        uint64_t *slot = desc + static_cast<int64_t>(tile) * 4;
        const bool term = resolved || (ta.a == 0.0f);
        slot[0] = term ? pack_desc(epoch, ST_TERM, 0u, resolved ? apply(ta, rn) : ta.b)
                       : pack_desc(epoch, ST_AGG, 0u, ta.a);
        slot[1] = static_cast<uint64_t>(trail);
        slot[3] = static_cast<uint64_t>(__float_as_uint(ta.b));
        if (!resolved) {
            if (TILE - trail > LONG_RUN) ulist2[atomicAdd(hdr + HDR_UCOUNT2, 1u)] = tile;
            else ulist[atomicAdd(hdr + HDR_UCOUNT, 1u)] = tile;
        }
    }
    // S at the first element after my warp's span
    const float s_after_warp = apply(ws, resolved ? rn : 0.0f);
    // ---- pass 2: per-element S and E, store E*S ----
    const int64_t wbase = base + static_cast<int64_t>(warp) * (ROWS * 128);
    const bool full = (base + TILE <= n);
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
        const uint32_t h = (hm >> (4 * r)) & 15u;
        const uint32_t t = (tm >> (4 * r)) & 15u;
        // S of the element right after this lane's e=3 (only used when e=3 is not a tail)
        const float sn = apply(sx[r], s_after_warp);
        const float s3 = (t & 8u) ? g[r][3] : fmaf(xn[r], sn, g[r][3]);
        const float s2 = (t & 4u) ? g[r][2] : fmaf(x[r][3], s3, g[r][2]);
        const float s1 = (t & 2u) ? g[r][1] : fmaf(x[r][2], s2, g[r][1]);
        const float s0 = (t & 1u) ? g[r][0] : fmaf(x[r][1], s1, g[r][0]);
        // forward carry into e=0: product of x from the segment head up to the previous element
        float c = cv[r];
        bool f = (cf >> r) & 1u;
        if (!f) {
            c = (warp > 0) ? wp_v * c : c;
            f = wp_f;
        }
        if (!f) c = y_prev * c;
        const float e0 = (h & 1u) ? 1.0f : c;
        const float e1 = (h & 2u) ? 1.0f : e0 * x[r][0];
        const float e2 = (h & 4u) ? 1.0f : e1 * x[r][1];
        const float e3 = (h & 8u) ? 1.0f : e2 * x[r][2];
        const float o0 = e0 * s0, o1 = e1 * s1, o2 = e2 * s2, o3 = e3 * s3;
        const int64_t gi = wbase + r * 128 + lane * 4;
        if (full && out_vec) {
            stcs_f4(gin + gi, o0, o1, o2, o3);
        } else {
            if (gi + 0 < n) __stcs(gin + gi + 0, o0);
            if (gi + 1 < n) __stcs(gin + gi + 1, o1);
            if (gi + 2 < n) __stcs(gin + gi + 2, o2);
            if (gi + 3 < n) __stcs(gin + gi + 3, o3);
        }
    }
}

// ---------------------------------------------------------------------------
// Fix-up of ONE tile whose R the halo could not resolve (one warp); runs when every descriptor
// of the launch is complete, so it never waits.  The trailing run [tile start + trail, tile end)
// contains no tail, so S there is a plain reverse affine scan seeded with R;
// E_i = y[i-1] (1 at the run start when it is a segment head).
// ---------------------------------------------------------------------------
// Walk forward over the (complete) descriptors to the nearest tile whose S(first element) is known;
// returns S at the first element of tile t+1 and publishes tile t's own inclusive carry (word2).
__device__ __forceinline__ float bwd_fix_walk(int64_t t, uint32_t num_tiles, uint32_t epoch, uint64_t *desc, int lane) {
    // ---- walk forward to the nearest tile whose S(first element) is known ----
    Affine carry = affine_id();
    int64_t nb = t + 1;
    while (true) {
        const int64_t idx = nb + lane;
        Affine m = Affine{0.0f, 0.0f};  // beyond the last tile: nothing follows
        bool term = true;
        if (idx < static_cast<int64_t>(num_tiles)) {
            const uint64_t d0 = ld_relaxed_u64(desc + idx * 4);
            if (desc_status(d0) == ST_TERM) {
                m = Affine{0.0f, desc_value(d0)};
            } else {
                const uint64_t d2 = ld_relaxed_u64(desc + idx * 4 + 2);
                if (desc_valid(d2, epoch) && desc_status(d2) == ST_INCL) {
                    m = Affine{0.0f, desc_value(d2)};
                } else {
                    m = Affine{desc_value(d0),
                               __uint_as_float(static_cast<uint32_t>(ld_relaxed_u64(desc + idx * 4 + 3)))};
                    term = (m.a == 0.0f);
                }
            }
        }
        const uint32_t tmk = __ballot_sync(0xffffffffu, term);
        const int last = tmk ? (__ffs(tmk) - 1) : 31;
        Affine w = (lane <= last) ? m : affine_id();
        w = warp_compose_all(w, lane);
        carry = compose(carry, w);
        if (tmk) break;
        nb += 32;
    }
    const float S = carry.b;  // S at the first element of tile t+1
    const uint64_t d0 = ld_relaxed_u64(desc + t * 4);
    if (desc_status(d0) == ST_AGG && lane == 0) {
        const Affine ta = Affine{desc_value(d0), __uint_as_float(static_cast<uint32_t>(ld_relaxed_u64(desc + t * 4 + 3)))};
        st_relaxed_u64(desc + t * 4 + 2, pack_desc(epoch, ST_INCL, 0u, apply(ta, S)));
    }
    return S;
}

__device__ __forceinline__ void bwd_fix_tile(int64_t t, const float *__restrict__ x, const float *__restrict__ y,
                                             const float *__restrict__ g, const int32_t *__restrict__ inv,
                                             float *gin, int64_t n, uint32_t num_tiles, int tile_elems,
                                             uint32_t epoch, uint64_t *desc, int lane, int64_t lead = 0) {
    // lead: the first `lead` (< 4) positions of the arrays are phantom elements in front of the caller's element 0
    // (alignment peel of the blocked kernels, see gcp_blk.cuh): never read as data, never written.
    const uint32_t trail = static_cast<uint32_t>(ld_relaxed_u64(desc + t * 4 + 1));
    float S = bwd_fix_walk(t, num_tiles, epoch, desc, lane);
    // ---- recompute the trailing run, 128 elements per step, from the tile end ----
    int64_t rs = t * tile_elems + trail;
    const int64_t re = (t + 1) * tile_elems;  // < n: the last tile always resolves
    if (rs >= re) return;
    bool rs_head = trail > 0u;
    if (rs < lead) rs = lead;                 // (trail == 0 in tile 0: the run starts at the caller's element 0)
    if (!rs_head) rs_head = (rs == lead) || (__ldg(inv + rs) != __ldg(inv + rs - 1));
    for (int64_t ce = re; ce > rs; ce -= 128) {
        const int64_t i0 = ce - 128 + lane * 4;
        float xn[4], gv[4];
        bool ok[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int64_t i = i0 + e;
            ok[e] = (i >= rs);
            xn[e] = ok[e] ? __ldg(x + i + 1) : 1.0f;  // i+1 <= re < n
            gv[e] = ok[e] ? __ldg(g + i) : 0.0f;
        }
        Affine m = Affine{xn[3], gv[3]};
        m = compose(Affine{xn[2], gv[2]}, m);
        m = compose(Affine{xn[1], gv[1]}, m);
        m = compose(Affine{xn[0], gv[0]}, m);
        Affine inc = m;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            Affine q;
            q.a = __shfl_down_sync(0xffffffffu, inc.a, d);
            q.b = __shfl_down_sync(0xffffffffu, inc.b, d);
            if (lane + d < 32) inc = compose(inc, q);
        }
        Affine exc;
        exc.a = __shfl_down_sync(0xffffffffu, inc.a, 1);
        exc.b = __shfl_down_sync(0xffffffffu, inc.b, 1);
        if (lane == 31) exc = affine_id();
        Affine tot;
        tot.a = __shfl_sync(0xffffffffu, inc.a, 0);
        tot.b = __shfl_sync(0xffffffffu, inc.b, 0);
        const float sn = apply(exc, S);
        const float s3 = fmaf(xn[3], sn, gv[3]);
        const float s2 = fmaf(xn[2], s3, gv[2]);
        const float s1 = fmaf(xn[1], s2, gv[1]);
        const float s0 = fmaf(xn[0], s1, gv[0]);
        const float sv[4] = {s0, s1, s2, s3};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int64_t i = i0 + e;
            if (ok[e]) {
                const float ev = (i == rs && rs_head) ? 1.0f : __ldg(y + i - 1);
                gin[i] = ev * sv[e];
            }
        }
        S = apply(tot, S);
    }
}

__device__ __forceinline__ void load_row_global_bwd(const float *__restrict__ x, const float *__restrict__ g,
                                                    const int32_t *__restrict__ inv, int64_t gi, int64_t n,
                                                    bool vec, float (&xv)[4], float (&gv)[4], int32_t (&iv)[4]) {
    if (vec && gi + 3 < n) {
        const float4 a = ldcs_f4(x + gi);
        const float4 b = ldcs_f4(g + gi);
        const int4 c = ldcs_i4(inv + gi);
        xv[0] = a.x; xv[1] = a.y; xv[2] = a.z; xv[3] = a.w;
        gv[0] = b.x; gv[1] = b.y; gv[2] = b.z; gv[3] = b.w;
        iv[0] = c.x; iv[1] = c.y; iv[2] = c.z; iv[3] = c.w;
    } else {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const bool in = gi + e < n;
            xv[e] = in ? __ldcs(x + gi + e) : 1.0f;
            gv[e] = in ? __ldcs(g + gi + e) : 0.0f;
            iv[e] = in ? __ldcs(inv + gi + e) : -1;
        }
    }
}

// ---------------------------------------------------------------------------
// K1, LDG variant: one tile per CTA, any alignment.
// ---------------------------------------------------------------------------
template <int WARPS, int ROWS>
__global__ void __launch_bounds__(WARPS * 32)
k_bwd_ldg(const float *__restrict__ x, const float *__restrict__ y, const float *__restrict__ g,
          const int32_t *__restrict__ inv, float *__restrict__ gin, int64_t n, uint32_t num_tiles,
          uint32_t *__restrict__ hdr, uint64_t *__restrict__ desc, uint32_t *__restrict__ ulist, int in_vec,
          int out_vec, int use_halo) {
    constexpr int TILE = WARPS * ROWS * 128;
    __shared__ BwdShared<WARPS> sh;
    __shared__ uint32_t s_epoch;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        sh.tile = atomicAdd(hdr + HDR_TICKET, 1u);
        s_epoch = ld_relaxed_u32(hdr + HDR_EPOCH);
    }
    __syncthreads();
    const uint32_t ticket = sh.tile;
    const uint32_t epoch = s_epoch;
    if (ticket < num_tiles) {
        const uint32_t tile = num_tiles - 1u - ticket;
        const int64_t base = static_cast<int64_t>(tile) * TILE;
        const int64_t wbase = base + static_cast<int64_t>(warp) * (ROWS * 128);
        const int64_t wend = wbase + ROWS * 128;
        float xv[ROWS][4], gv[ROWS][4];
        int32_t iv[ROWS][4];
        int32_t iprev = -1, inext = -1;
        float xnext = 0.0f;
        if (lane == 0 && wbase > 0 && wbase - 1 < n) iprev = __ldg(inv + wbase - 1);
        const float y_prev = (base > 0) ? __ldg(y + base - 1) : 1.0f;
#pragma unroll
        for (int r = 0; r < ROWS; ++r)
            load_row_global_bwd(x, g, inv, wbase + r * 128 + lane * 4, n, in_vec != 0, xv[r], gv[r], iv[r]);
        if (warp == WARPS - 1) {
            // the last warp's right neighbour is the next tile: resolve R from the halo there
            float R = 0.0f;
            bool res;
            const int64_t end = base + TILE;
            if (use_halo) {
                res = halo_suffix(x, g, inv, end, n, lane, in_vec != 0, R, inext, xnext);
            } else {
                res = (end >= n);
                if (end < n) {
                    inext = __ldg(inv + end);
                    xnext = __ldg(x + end);
                }
            }
            if (lane == 0) {
                sh.res = res ? 1u : 0u;
                sh.rn = R;
            }
        } else if (lane == 0 && wend < n) {
            inext = __ldg(inv + wend);
            xnext = __ldg(x + wend);
        }
        bwd_tile_body<WARPS, ROWS, true>(xv, gv, iv, iprev, inext, xnext, y_prev, false, 0.0f, tile, base, n, gin,
                                         out_vec != 0, epoch, hdr, desc, ulist, ulist + num_tiles, &sh, warp, lane);
    }
    if (threadIdx.x == 0) finish_stream_kernel(hdr);
}

// K2 of the LDG path: the same fix-up as a separate launch (one warp per list entry).
__global__ void __launch_bounds__(256)
k_bwd_fix(const float *__restrict__ x, const float *__restrict__ y, const float *__restrict__ g,
          const int32_t *__restrict__ inv, float *gin, int64_t n, uint32_t num_tiles, int tile_elems, uint32_t *hdr,
          uint64_t *desc, const uint32_t *ulist) {
    const int lane = threadIdx.x & 31;
    const uint32_t gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t nw = (gridDim.x * blockDim.x) >> 5;
    const uint32_t epoch = ld_relaxed_u32(hdr + HDR_EPOCH);
    const uint32_t ucount = ld_relaxed_u32(hdr + HDR_UCOUNT);
    for (uint32_t u = gw; u < ucount; u += nw)
        bwd_fix_tile(static_cast<int64_t>(__ldcg(ulist + u)), x, y, g, inv, gin, n, num_tiles, tile_elems, epoch, desc,
                     lane);
    const uint32_t ucount2 = ld_relaxed_u32(hdr + HDR_UCOUNT2);
    for (uint32_t u = gw; u < ucount2; u += nw)
        bwd_fix_tile(static_cast<int64_t>(__ldcg(ulist + num_tiles + u)), x, y, g, inv, gin, n, num_tiles, tile_elems,
                     epoch, desc, lane);
    __syncthreads();
    if (threadIdx.x == 0) finish_op(hdr, epoch);
}

// ---------------------------------------------------------------------------
// Integer-side validation of (inv, seg_end): see gcp_validate_segments.
// ---------------------------------------------------------------------------
__global__ void k_validate_segments(const int32_t *__restrict__ inv, const int32_t *__restrict__ seg_end, int64_t n,
                                    int64_t k, unsigned long long *__restrict__ violations) {
    unsigned long long bad = 0;
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
        const int32_t s = inv[i];
        if (s < 0 || s >= k) { ++bad; continue; }
        if (i == 0 && s != 0) ++bad;
        if (i > 0) {
            const int32_t p = inv[i - 1];
            if (s != p && s != p + 1) ++bad;
        }
        const bool tail = (i == n - 1) || (inv[i + 1] != s);
        if (tail != (seg_end[s] == i + 1)) ++bad;
        if (i == n - 1 && s != k - 1) ++bad;
    }
    if (n == 0 && k != 0 && blockIdx.x == 0 && threadIdx.x == 0) ++bad;
    if (bad) atomicAdd(violations, bad);
}

}  // namespace gcp
