// gcp_fwd.cuh — forward segmented inclusive scan (cumprod / cumsum), wait-free.
//
// Replaces thrust::inclusive_scan_by_key at
//   /root/reference/cuda_kernel/grouped_cumprod_forward.cu:17-23  (OP_MUL)
//   /root/reference/cuda_kernel/grouped_cumsum_forward.cu:17-23   (OP_ADD)
//
// Layout inside a tile of TILE = WARPS*ROWS*128 elements: warp w owns the contiguous
// span [w*ROWS*128, (w+1)*ROWS*128); in row r lane l holds the four consecutive
// elements at r*128 + l*4 ("striped float4"): every global / shared access is a
// coalesced, bank-conflict-free 128-bit access and no shared-memory transpose is
// needed.  Scan = thread-serial over the float4 -> ballot-masked segmented warp scan
// per row (5 SHFL) -> serial chain over the ROWS rows -> WARPS warp aggregates
// through shared memory (ONE named barrier per tile).
//
// Cross-tile carry, without any inter-CTA waiting:
//   K1 (k_fwd_blk in gcp_blk.cuh / k_fwd_ldg)  resolves the tile's exclusive prefix from the HALO, the
//      128 elements before the tile (a per-pixel list is ~20-50 elements, so a segment
//      head is almost always inside it).  If no head is found the tile is "unresolved":
//      it stores its results as if the prefix were the identity and publishes a carry
//      descriptor {TERM|AGG, value} plus a fix-up request {needs, lead = #elements before
//      the tile's first head}.  Nothing ever polls: K1 never blocks on another CTA.
//   Fix-up (fwd_fix_tile), one warp per unresolved tile, once all descriptors of the launch
//      are complete: walk back over the descriptors 32 tiles per round to the nearest TERM /
//      already-fixed tile and apply prefix (x) y over the `lead` elements.  In the persistent
//      kernel it is a second phase of the SAME launch behind a grid barrier (all CTAs are
//      resident); the LDG path launches it as a separate kernel K2 (k_fwd_fix).
//
//   k_fwd_blk : (gcp_blk.cuh, the default) persistent CTAs, tensor-map TMA ring, blocked layout.
//   k_fwd_ldg : one tile per CTA, direct streaming loads; any alignment, any n (the fallback: arrays
//               whose 16-byte phases differ, n smaller than a tile, cooperative launch unavailable).
#pragma once
#include "gcp_device.cuh"

namespace gcp {

constexpr uint32_t NO_POS = 0xFFFFFFFFu;

template <int WARPS>
struct FwdShared {
    float wv[WARPS];     // warp aggregates
    uint32_t wf[WARPS];  // "warp span contains a head"
    uint32_t fh[WARPS];  // offset of the first head inside the warp span (NO_POS if none)
    uint32_t res;        // LDG kernel: halo result of warp 0
    float tp;
    uint32_t tile;
};

// Segmented inclusive scan of one aggregate per (row, lane) across the warp and down its rows.
//   agg[r] : op over this lane's 4 elements of row r after its last head
//   hm     : head bits of the lane's elements, bit r*4+e
//   cv/cf  : out, per-row carry INTO this lane from earlier lanes/rows of the warp
//            (value; bit r of cf = "a head lies between the warp start and this lane")
//   wv/wf  : out, warp aggregate;  fh : out, offset of the first head in the warp span
template <int OP, int ROWS>
__device__ __forceinline__ void warp_seg_scan_rows(const float (&agg)[ROWS], uint32_t hm, int lane,
                                                   float (&cv)[ROWS], uint32_t &cf, float &wv, uint32_t &wf,
                                                   uint32_t &fh) {
    using O = ScanOp<OP>;
    const uint32_t lt = (1u << lane) - 1u;
    const uint32_t le = lt | (1u << lane);
    float rp_v = O::id();
    uint32_t rp_f = 0u;
    cf = 0u;
    fh = NO_POS;
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
        const uint32_t h = (hm >> (4 * r)) & 15u;
        const uint32_t m = __ballot_sync(0xffffffffu, h != 0u);
        const int start = max(31 - __clz(m & le), 0);  // nearest lane <= me holding a head (0 if none)
        float inc = agg[r];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            float t = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane - d >= start) inc = O::f(t, inc);
        }
        float exc = __shfl_up_sync(0xffffffffu, inc, 1);
        if (lane == 0) exc = O::id();
        const bool ef = (m & lt) != 0u;
        const float row_v = __shfl_sync(0xffffffffu, inc, 31);
        const bool row_f = m != 0u;
        if (row_f && rp_f == 0u) {  // first row of the warp that holds a head (warp-uniform branch)
            const int l0 = __ffs(m) - 1;
            const uint32_t h0 = __shfl_sync(0xffffffffu, h, l0);
            fh = static_cast<uint32_t>(r * 128 + l0 * 4 + (__ffs(h0) - 1));
        }
        cv[r] = ef ? exc : O::f(rp_v, exc);
        cf |= ((ef || rp_f) ? 1u : 0u) << r;
        rp_v = row_f ? row_v : O::f(rp_v, row_v);
        rp_f |= row_f ? 1u : 0u;
    }
    wv = rp_v;
    wf = rp_f;
}

// op over the 32 lanes' values, every lane gets the result.  Only ADJACENT lane ranges are ever combined (a
// shfl_down tree, then a broadcast), like the reference's thrust scan: a butterfly would also multiply
// non-adjacent ranges, whose product can overflow although every contiguous sub-range product is finite.
template <int OP>
__device__ __forceinline__ float warp_reduce_ordered(float w) {
    using O = ScanOp<OP>;
#pragma unroll
    for (int d = 1; d <= 16; d <<= 1) w = O::f(w, __shfl_down_sync(0xffffffffu, w, d));
    return __shfl_sync(0xffffffffu, w, 0);
}

// All 32 lanes of one warp: resolve the exclusive prefix of the tile that starts at `base`
// from the 128 elements before it.  True when the tile's first element is a head (prefix
// irrelevant) or a segment head lies inside the window; then P = op over [last head, base).
// kprev = key[base-1].  Needs 128 <= base < n.
template <int OP>
__device__ __forceinline__ bool halo_prefix(const float *__restrict__ x, const int32_t *__restrict__ key,
                                            int64_t base, int lane, bool vec, float &P, int32_t &kprev) {
    using O = ScanOp<OP>;
    const int64_t h0 = base - 128 + lane * 4;
    float4 a;
    int4 b;
    if (vec) {
        a = __ldg(reinterpret_cast<const float4 *>(x + h0));
        b = __ldg(reinterpret_cast<const int4 *>(key + h0));
    } else {
        a = make_float4(__ldg(x + h0), __ldg(x + h0 + 1), __ldg(x + h0 + 2), __ldg(x + h0 + 3));
        b = make_int4(__ldg(key + h0), __ldg(key + h0 + 1), __ldg(key + h0 + 2), __ldg(key + h0 + 3));
    }
    const int32_t kfirst = __ldg(key + base);
    const int32_t pk = __shfl_up_sync(0xffffffffu, b.w, 1);
    const uint32_t h = ((lane > 0 && b.x != pk) ? 1u : 0u) | (b.y != b.x ? 2u : 0u) | (b.z != b.y ? 4u : 0u) |
                       (b.w != b.z ? 8u : 0u);
    kprev = __shfl_sync(0xffffffffu, b.w, 31);
    const uint32_t m = __ballot_sync(0xffffffffu, h != 0u);
    float v = a.x;
    v = (h & 2u) ? a.y : O::f(v, a.y);
    v = (h & 4u) ? a.z : O::f(v, a.z);
    v = (h & 8u) ? a.w : O::f(v, a.w);
    const int last = 31 - __clz(m);  // highest lane holding a head, -1 if none
    float w = (lane >= last) ? v : O::id();
    w = warp_reduce_ordered<OP>(w);
    if (kfirst != kprev) {
        P = O::id();
        return true;
    }
    P = w;
    return m != 0u;
}

// The same halo resolution split in two so that a producer can issue the loads of tile i and
// consume them one iteration later (software pipelining).  16-byte aligned inputs only.
// HQ = float4 quads per lane: the window is the 128*HQ elements before the tile (base >= 128*HQ).
template <int HQ>
struct HaloPrefixRegs {
    float4 a[HQ];
    int4 b[HQ];
    int32_t kfirst;  // key[base]; with halo resolution off: key[base-1]
};
template <int HQ>
__device__ __forceinline__ void halo_prefix_issue(const float *__restrict__ x, const int32_t *__restrict__ key,
                                                  int64_t base, int lane, HaloPrefixRegs<HQ> &r) {
    const int64_t h0 = base - 128 * HQ + lane * (4 * HQ);
#pragma unroll
    for (int c = 0; c < HQ; ++c) {
        r.a[c] = __ldg(reinterpret_cast<const float4 *>(x + h0) + c);
        r.b[c] = __ldg(reinterpret_cast<const int4 *>(key + h0) + c);
    }
    r.kfirst = __ldg(key + base);
}
template <int OP, int HQ>
__device__ __forceinline__ bool halo_prefix_finish(const HaloPrefixRegs<HQ> &r, int lane, float &P, int32_t &kprev) {
    using O = ScanOp<OP>;
    float xv[4 * HQ];
    int32_t kv[4 * HQ];
#pragma unroll
    for (int c = 0; c < HQ; ++c) {
        xv[4 * c] = r.a[c].x; xv[4 * c + 1] = r.a[c].y; xv[4 * c + 2] = r.a[c].z; xv[4 * c + 3] = r.a[c].w;
        kv[4 * c] = r.b[c].x; kv[4 * c + 1] = r.b[c].y; kv[4 * c + 2] = r.b[c].z; kv[4 * c + 3] = r.b[c].w;
    }
    const int32_t pk = __shfl_up_sync(0xffffffffu, kv[4 * HQ - 1], 1);
    kprev = __shfl_sync(0xffffffffu, kv[4 * HQ - 1], 31);
    bool any = lane > 0 && kv[0] != pk;
    float v = xv[0];  // op over the lane's elements since its last head
#pragma unroll
    for (int e = 1; e < 4 * HQ; ++e) {
        const bool h = kv[e] != kv[e - 1];
        v = h ? xv[e] : O::f(v, xv[e]);
        any |= h;
    }
    const uint32_t m = __ballot_sync(0xffffffffu, any);
    const int last = 31 - __clz(m);  // highest lane holding a head, -1 if none
    float w = (lane >= last) ? v : O::id();
    w = warp_reduce_ordered<OP>(w);
    if (r.kfirst != kprev) {
        P = O::id();
        return true;
    }
    P = w;
    return m != 0u;
}

// Everything after the tile's x/key values are in registers.
//   kprev      : key of the element just before this warp's span (only lane 0 needs it)
//   first_head : this warp's first element is global element 0
//   resolved/tp: CTA-uniform halo result (TMA kernel: from the producer; LDG kernel: read from sh)
template <int OP, int WARPS, int ROWS, bool HALO_IN_SH>
__device__ __forceinline__ void fwd_tile_body(float (&v)[ROWS][4], const int32_t (&k)[ROWS][4], int32_t kprev,
                                              bool first_head, bool resolved, float tp, uint32_t tile, int64_t base,
                                              int64_t n, float *__restrict__ y, bool y_vec, uint32_t epoch,
                                              uint32_t *__restrict__ hdr, uint64_t *__restrict__ desc,
                                              uint32_t *__restrict__ ulist, FwdShared<WARPS> *sh, int warp,
                                              int lane) {
    using O = ScanOp<OP>;
    constexpr int TILE = WARPS * ROWS * 128;
    // ---- head flags ----
    uint32_t hm = 0u;
    int32_t carry_key = kprev;  // key just before row r's lane 0
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
        int32_t p = __shfl_up_sync(0xffffffffu, k[r][3], 1);
        if (lane == 0) p = carry_key;
        carry_key = __shfl_sync(0xffffffffu, k[r][3], 31);
        uint32_t h = (k[r][0] != p ? 1u : 0u) | (k[r][1] != k[r][0] ? 2u : 0u) | (k[r][2] != k[r][1] ? 4u : 0u) |
                     (k[r][3] != k[r][2] ? 8u : 0u);
        if (r == 0 && lane == 0 && first_head) h |= 1u;
        hm |= h << (4 * r);
    }
    // ---- thread-serial + warp-level scan ----
    float agg[ROWS];
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
        const uint32_t h = (hm >> (4 * r)) & 15u;
        v[r][1] = (h & 2u) ? v[r][1] : O::f(v[r][0], v[r][1]);
        v[r][2] = (h & 4u) ? v[r][2] : O::f(v[r][1], v[r][2]);
        v[r][3] = (h & 8u) ? v[r][3] : O::f(v[r][2], v[r][3]);
        agg[r] = v[r][3];
    }
    float cv[ROWS];
    uint32_t cf, wf, fh;
    float wv;
    warp_seg_scan_rows<OP, ROWS>(agg, hm, lane, cv, cf, wv, wf, fh);
    if (lane == 0) {
        sh->wv[warp] = wv;
        sh->wf[warp] = wf;
        sh->fh[warp] = fh;
    }
    named_bar_sync<WARPS * 32>(1);
    if (HALO_IN_SH) {
        resolved = sh->res != 0u;
        tp = sh->tp;
    }
    // ---- exclusive prefix over warps, tile aggregate, first head of the tile ----
    float wp_v = O::id(), ta_v = O::id();
    uint32_t wp_f = 0u, ta_f = 0u, lead = TILE;
#pragma unroll
    for (int j = 0; j < WARPS; ++j) {
        const float jv = sh->wv[j];
        const uint32_t jf = sh->wf[j];
        if (j < warp) {
            wp_v = jf ? jv : O::f(wp_v, jv);
            wp_f |= jf;
        }
        if (jf && ta_f == 0u) lead = static_cast<uint32_t>(j * ROWS * 128) + sh->fh[j];
        ta_v = jf ? jv : O::f(ta_v, jv);
        ta_f |= jf;
    }
    // ---- publish the carry descriptor (+ fix-up request when the halo held no head) ----
    if (warp == 0 && lane == 0) {
        uint64_t *slot = desc + static_cast<int64_t>(tile) * 4;
        const bool term = (ta_f != 0u) || resolved;
        const float val = ta_f ? ta_v : (resolved ? O::f(tp, ta_v) : ta_v);
        slot[0] = pack_desc(epoch, term ? ST_TERM : ST_AGG, ta_f, val);
        slot[1] = static_cast<uint64_t>(lead);
        if (!resolved) ulist[atomicAdd(hdr + HDR_UCOUNT, 1u)] = tile;
    }
    const float tp_v = resolved ? tp : O::id();
    // ---- apply carries, store ----
    const int64_t wbase = base + static_cast<int64_t>(warp) * (ROWS * 128);
    const bool full = (base + TILE <= n);
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
        float c = cv[r];
        bool f = (cf >> r) & 1u;
        if (!f) {
            c = O::f(wp_v, c);
            f = wp_f != 0u;
        }
        if (!f) c = O::f(tp_v, c);
        const uint32_t h = (hm >> (4 * r)) & 15u;
        const bool n0 = !(h & 1u), n1 = n0 && !(h & 2u), n2 = n1 && !(h & 4u), n3 = n2 && !(h & 8u);
        const float o0 = n0 ? O::f(c, v[r][0]) : v[r][0];
        const float o1 = n1 ? O::f(c, v[r][1]) : v[r][1];
        const float o2 = n2 ? O::f(c, v[r][2]) : v[r][2];
        const float o3 = n3 ? O::f(c, v[r][3]) : v[r][3];
        const int64_t gi = wbase + r * 128 + lane * 4;
        if (full && y_vec) {
            stcs_f4(y + gi, o0, o1, o2, o3);
        } else {
            if (gi + 0 < n) __stcs(y + gi + 0, o0);
            if (gi + 1 < n) __stcs(y + gi + 1, o1);
            if (gi + 2 < n) __stcs(y + gi + 2, o2);
            if (gi + 3 < n) __stcs(y + gi + 3, o3);
        }
    }
}

// ---------------------------------------------------------------------------
// Fix-up of ONE tile whose prefix the halo could not resolve (one warp).  Runs when every
// descriptor of the launch is complete (after the grid barrier of the persistent kernel, or in
// K2 after the LDG K1), so the walk never waits.  An AGG tile that has been fixed publishes its
// inclusive carry in word2 (epoch-tagged) so that later walkers stop there; reading a word2 that
// is not there yet only makes a walk longer, never wrong.
// ---------------------------------------------------------------------------
template <int OP>
__device__ __forceinline__ void fwd_fix_tile(uint32_t t, float *y, int64_t n, int tile_elems, uint32_t epoch,
                                             uint64_t *desc, int lane) {
    using O = ScanOp<OP>;
    const uint32_t lead = static_cast<uint32_t>(ld_relaxed_u64(desc + static_cast<int64_t>(t) * 4 + 1));
    float carry = O::id();
    int64_t pb = static_cast<int64_t>(t) - 1;
    while (true) {
        const int64_t idx = pb - lane;
        bool term = true;
        float v = O::id();
        if (idx >= 0) {
            const uint64_t d0 = ld_relaxed_u64(desc + idx * 4);
            term = desc_status(d0) == ST_TERM;
            v = desc_value(d0);
            if (!term) {
                const uint64_t d2 = ld_relaxed_u64(desc + idx * 4 + 2);
                if (desc_valid(d2, epoch) && desc_status(d2) == ST_INCL) {
                    term = true;
                    v = desc_value(d2);
                }
            }
        }
        const uint32_t tm = __ballot_sync(0xffffffffu, term);
        const int last = tm ? (__ffs(tm) - 1) : 31;
        float w = (lane <= last) ? v : O::id();
        w = warp_reduce_ordered<OP>(w);
        carry = O::f(w, carry);
        if (tm) break;
        pb -= 32;
    }
    const uint64_t d0 = ld_relaxed_u64(desc + static_cast<int64_t>(t) * 4);
    if (desc_status(d0) == ST_AGG && lane == 0)
        st_relaxed_u64(desc + static_cast<int64_t>(t) * 4 + 2,
                       pack_desc(epoch, ST_INCL, 0u, O::f(carry, desc_value(d0))));
    const int64_t base = static_cast<int64_t>(t) * tile_elems;
    int64_t end = base + lead;
    if (end > n) end = n;
    for (int64_t i = base + lane; i < end; i += 32) y[i] = O::f(carry, __ldcg(y + i));
}

// Guarded / unaligned-capable global load of one thread-row (4 elements).
template <int OP>
__device__ __forceinline__ void load_row_global(const float *__restrict__ x, const int32_t *__restrict__ key,
                                                int64_t gi, int64_t n, bool vec, float (&v)[4], int32_t (&k)[4]) {
    using O = ScanOp<OP>;
    if (vec && gi + 3 < n) {
        float4 a = ldcs_f4(x + gi);
        int4 b = ldcs_i4(key + gi);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
        k[0] = b.x; k[1] = b.y; k[2] = b.z; k[3] = b.w;
    } else {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const bool in = gi + e < n;
            v[e] = in ? __ldcs(x + gi + e) : O::id();
            k[e] = in ? __ldcs(key + gi + e) : 0;
        }
    }
}

// ---------------------------------------------------------------------------
// K1, LDG variant: one tile per CTA (ticket taken at entry), any alignment.
// ---------------------------------------------------------------------------
template <int OP, int WARPS, int ROWS>
__global__ void __launch_bounds__(WARPS * 32)
k_fwd_ldg(const float *__restrict__ x, const int32_t *__restrict__ key, float *__restrict__ y, int64_t n,
          uint32_t num_tiles, uint32_t *__restrict__ hdr, uint64_t *__restrict__ desc,
          uint32_t *__restrict__ ulist, int in_vec, int y_vec, int use_halo) {
    using O = ScanOp<OP>;
    constexpr int TILE = WARPS * ROWS * 128;
    __shared__ FwdShared<WARPS> sh;
    __shared__ uint32_t s_epoch;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        sh.tile = atomicAdd(hdr + HDR_TICKET, 1u);
        s_epoch = ld_relaxed_u32(hdr + HDR_EPOCH);
    }
    __syncthreads();
    const uint32_t tile = sh.tile;
    const uint32_t epoch = s_epoch;
    if (tile < num_tiles) {
        const int64_t base = static_cast<int64_t>(tile) * TILE;
        const int64_t wbase = base + static_cast<int64_t>(warp) * (ROWS * 128);
        float v[ROWS][4];
        int32_t k[ROWS][4];
        int32_t kprev = 0;
#pragma unroll
        for (int r = 0; r < ROWS; ++r)
            load_row_global<OP>(x, key, wbase + r * 128 + lane * 4, n, in_vec != 0, v[r], k[r]);
        if (warp == 0) {
            float P = O::id();
            bool res = (tile == 0u);
            if (tile > 0u) {
                if (use_halo) res = halo_prefix<OP>(x, key, base, lane, in_vec != 0, P, kprev);
                else kprev = __ldg(key + base - 1);
            }
            if (lane == 0) {
                sh.res = res ? 1u : 0u;
                sh.tp = P;
            }
        } else if (lane == 0 && wbase - 1 < n) {
            kprev = __ldg(key + wbase - 1);
        }
        fwd_tile_body<OP, WARPS, ROWS, true>(v, k, kprev, wbase == 0, false, 0.0f, tile, base, n, y, y_vec != 0,
                                             epoch, hdr, desc, ulist, &sh, warp, lane);
    }
    if (threadIdx.x == 0) finish_stream_kernel(hdr);
}

// K2 of the LDG path: the same fix-up as a separate launch (one warp per list entry).
template <int OP>
__global__ void __launch_bounds__(256)
k_fwd_fix(float *y, int64_t n, int tile_elems, uint32_t *hdr, uint64_t *desc, const uint32_t *ulist) {
    const int lane = threadIdx.x & 31;
    const uint32_t gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t nw = (gridDim.x * blockDim.x) >> 5;
    const uint32_t epoch = ld_relaxed_u32(hdr + HDR_EPOCH);
    const uint32_t ucount = ld_relaxed_u32(hdr + HDR_UCOUNT);
    for (uint32_t u = gw; u < ucount; u += nw) fwd_fix_tile<OP>(__ldcg(ulist + u), y, n, tile_elems, epoch, desc, lane);
    __syncthreads();
    if (threadIdx.x == 0) finish_op(hdr, epoch);
}

}  // namespace gcp
