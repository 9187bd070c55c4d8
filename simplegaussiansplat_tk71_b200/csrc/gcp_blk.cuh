// gcp_blk.cuh — "blocked" persistent kernels: every lane owns 16 CONSECUTIVE elements.
//
// The striped-float4 kernels of gcp_fwd.cuh / gcp_bwd.cuh pay one warp scan (ballots + 5-10
// shuffles + bookkeeping) per 4 elements of a lane, which makes the backward issue-bound.
// Here a tile is loaded with 2-D tensor-map TMA copies (cp.async.bulk.tensor.2d -> UTMALDG) using
// the 128-byte swizzle: the array is viewed as [n/32][32] f32 (128-byte rows), and the hardware
// XORs the 16-byte chunk index with (row & 7) while writing shared memory.  With that swizzle a
// lane can read its 16 consecutive elements as four LDS.128 without bank conflicts
// (lanes 0..7 of a quarter warp hit chunks 0,4,1,5,2,6,3,7), so the layout change costs nothing
// and each lane now runs ONE thread-serial scan over 16 elements and ONE warp scan per direction
// per tile.  Outputs go back through the (already consumed) shared-memory stage to be written
// with fully coalesced 128-bit streaming stores.
//
// Carry logic, descriptors, halo resolution and the in-kernel fix-up phase are exactly those of
// gcp_fwd.cuh / gcp_bwd.cuh (the fix-up functions are shared).
//
// Alignment peel (`lead`).  TMA needs 16-byte aligned global addresses; a sliced tensor (the reference
// slices with [cutting_number:], gs_model.py:557) starts 4, 8 or 12 bytes past one.  When all streamed
// arrays of an op share that phase, the host passes every pointer moved DOWN to the 16-byte boundary,
// n grown by the same `lead` = 1..3 elements, and `lead` itself: the kernels then see `lead` phantom
// elements in front of the caller's element 0.  They sit in tile 0 / warp 0 / lane 0 only, where they
// are replaced in registers by the scan identity with the key of the first real element (so they neither
// start nor end a segment, and contribute nothing), and that warp stores its outputs with scalar stores
// that skip them.  Everything else — the other 99.99 % of the tiles — runs the aligned fast path
// unchanged.  (The bytes in front of a 4-byte aligned pointer up to the 16-byte boundary belong to the same
// allocation: cudaMalloc and every pooling allocator hand out blocks aligned to at least 256 bytes.)
#pragma once
#include <type_traits>
#include <cuda.h>

#include "gcp_bwd.cuh"
#include "gcp_device.cuh"
#include "gcp_fwd.cuh"

namespace gcp {

constexpr int BLK_EPL = 16;              // elements per lane
constexpr int BLK_WSPAN = 32 * BLK_EPL;  // elements per warp = 512
#ifndef GCP_BLK_HQ
#define GCP_BLK_HQ 4
#endif
constexpr int BLK_HQ = GCP_BLK_HQ;       // halo window = 128 * BLK_HQ elements (resolves a tile's carry in place when a
                                         // segment boundary lies that close; the data are the neighbouring tile's, in L2)

// byte offset inside a swizzled stage array for logical byte offset `b` (128B swizzle)
__device__ __forceinline__ uint32_t swz(uint32_t b) { return b ^ (((b >> 7) & 7u) << 4); }

__device__ __forceinline__ void tma_load_2d(void *dst_smem, const CUtensorMap *tm, int c0, int c1, uint64_t *bar,
                                            uint64_t policy) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint"
        " [%0], [%1, {%2, %3}], [%4], %5;"
        ::"r"(smem_u32(dst_smem)), "l"(tm), "r"(c0), "r"(c1), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap *tm) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(tm) : "memory");
}

// lane's 16 consecutive elements of one stage array -> registers (4 conflict-free LDS.128)
template <typename T>
__device__ __forceinline__ void lds_blocked(const unsigned char *arr, int warp, int lane, T (&out)[16]) {
    const uint32_t b0 = static_cast<uint32_t>(warp * BLK_WSPAN + lane * BLK_EPL) * 4u;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const uint4 q = *reinterpret_cast<const uint4 *>(arr + swz(b0 + 16u * j));
        out[4 * j + 0] = *reinterpret_cast<const T *>(&q.x);
        out[4 * j + 1] = *reinterpret_cast<const T *>(&q.y);
        out[4 * j + 2] = *reinterpret_cast<const T *>(&q.z);
        out[4 * j + 3] = *reinterpret_cast<const T *>(&q.w);
    }
}
template <typename T>
__device__ __forceinline__ T lds_one(const unsigned char *arr, int elem) {
    return *reinterpret_cast<const T *>(arr + swz(static_cast<uint32_t>(elem) * 4u));
}

// registers (blocked) -> global, coalesced, through the warp's own span of a consumed stage array
__device__ __forceinline__ void store_blocked_via_smem(unsigned char *arr, int warp, int lane, const float (&o)[16],
                                                       float *__restrict__ dst_warp) {
    const uint32_t b0 = static_cast<uint32_t>(warp * BLK_WSPAN + lane * BLK_EPL) * 4u;
#pragma unroll
    for (int j = 0; j < 4; ++j)
        *reinterpret_cast<float4 *>(arr + swz(b0 + 16u * j)) = make_float4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
    __syncwarp();
    const uint32_t w0 = static_cast<uint32_t>(warp * BLK_WSPAN) * 4u;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float4 q = *reinterpret_cast<const float4 *>(arr + swz(w0 + (j * 32u + lane) * 16u));
        __stcs(reinterpret_cast<float4 *>(dst_warp + (j * 32 + lane) * 4), q);
    }
    // the next user of this shared memory is the TMA engine (async proxy)
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// guarded global load of a lane's 16 consecutive elements (partial last tile)
template <typename T>
__device__ __forceinline__ void ldg_blocked(const T *__restrict__ p, int64_t gi, int64_t n, T pad, T (&out)[16]) {
#pragma unroll
    for (int e = 0; e < 16; ++e) out[e] = (gi + e < n) ? __ldg(p + gi + e) : pad;
}
// full-tile variant: four 128-bit loads per lane (each lane reads 64 contiguous bytes)
template <typename T>
__device__ __forceinline__ void ldg_blocked_vec(const T *__restrict__ p, int64_t gi, T (&out)[16]) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const uint4 q = __ldg(reinterpret_cast<const uint4 *>(p + gi) + j);
        out[4 * j + 0] = *reinterpret_cast<const T *>(&q.x);
        out[4 * j + 1] = *reinterpret_cast<const T *>(&q.y);
        out[4 * j + 2] = *reinterpret_cast<const T *>(&q.z);
        out[4 * j + 3] = *reinterpret_cast<const T *>(&q.w);
    }
}
__device__ __forceinline__ void stg_blocked_guarded(float *__restrict__ p, int64_t gi, int64_t n, const float (&o)[16],
                                                    int64_t lo = 0) {
#pragma unroll
    for (int e = 0; e < 16; ++e)
        if (gi + e < n && gi + e >= lo) p[gi + e] = o[e];
}
// phantom elements in front of the caller's element 0 (alignment peel): identity value, key of the first real element
template <typename T>
__device__ __forceinline__ void mask_lead(T (&a)[16], int lead, T fill) {
#pragma unroll
    for (int e = 0; e < 3; ++e)
        if (e < lead) a[e] = fill;
}
template <typename T>
__device__ __forceinline__ void mask_lead_key(T (&a)[16], int lead) {
    const T k = lead == 1 ? a[1] : (lead == 2 ? a[2] : a[3]);
#pragma unroll
    for (int e = 0; e < 3; ++e)
        if (e < lead) a[e] = k;
}

// ============================================================================================
// forward
// ============================================================================================
template <int WARPS>
struct FwdBlkShared {
    float wv[WARPS];
    uint32_t wf[WARPS];
    uint32_t fh[WARPS];
};

// v: in x, out final y of the lane's 16 elements
template <int OP, int WARPS>
__device__ __forceinline__ void fwd_blk_compute(float (&v)[16], const int32_t (&k)[16], int32_t kprev, bool first_head,
                                                bool resolved, float tp, uint32_t tile, uint32_t epoch,
                                                uint32_t *__restrict__ hdr, uint64_t *__restrict__ desc,
                                                uint32_t *__restrict__ ulist, FwdBlkShared<WARPS> *sh, int warp,
                                                int lane, bool &term_out, float &carry_out) {
    using O = ScanOp<OP>;
    static_assert(WARPS < 32, "one lane per warp in the cross-warp step");
    const uint32_t lanes_lt = (1u << lane) - 1u;
    const uint32_t lanes_le = lanes_lt | (1u << lane);
    // ---- head bits ----
    int32_t p = __shfl_up_sync(0xffffffffu, k[15], 1);
    if (lane == 0) p = kprev;
    uint32_t hm = (k[0] != p) ? 1u : 0u;
    if (lane == 0 && first_head) hm = 1u;
#pragma unroll
    for (int e = 1; e < 16; ++e) hm |= (k[e] != k[e - 1] ? 1u : 0u) << e;
    // ---- thread-serial inclusive scan ----
#pragma unroll
    for (int e = 1; e < 16; ++e) v[e] = ((hm >> e) & 1u) ? v[e] : O::f(v[e - 1], v[e]);
    // ---- segmented warp scan of the lane aggregates (one per tile) ----
    const uint32_t m = __ballot_sync(0xffffffffu, hm != 0u);
    const int start = max(31 - __clz(m & lanes_le), 0);
    float inc = v[15];
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const float t = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane - d >= start) inc = O::f(t, inc);
    }
    float c = __shfl_up_sync(0xffffffffu, inc, 1);
    if (lane == 0) c = O::id();
    bool cf = (m & lanes_lt) != 0u;
    if (lane == 31) {
        sh->wv[warp] = inc;
        sh->wf[warp] = m != 0u ? 1u : 0u;
    }
    if (m != 0u) {  // warp-uniform
        const int l0 = __ffs(m) - 1;
        const uint32_t h0 = __shfl_sync(0xffffffffu, hm, l0);
        if (lane == 0) sh->fh[warp] = static_cast<uint32_t>(l0 * BLK_EPL + (__ffs(h0) - 1));
    }
    named_bar_sync<WARPS * 32>(1);
    // ---- across warps, one lane per warp ----
    const bool wl = lane < WARPS;
    float jv = wl ? sh->wv[lane] : O::id();
    const uint32_t jf = wl ? sh->wf[lane] : 0u;
    const uint32_t fm = __ballot_sync(0xffffffffu, jf != 0u);
    {
        const int st = max(31 - __clz(fm & lanes_le), 0);
#pragma unroll
        for (int d = 1; d < WARPS; d <<= 1) {
            const float tv = __shfl_up_sync(0xffffffffu, jv, d);
            if (lane - d >= st) jv = O::f(tv, jv);
        }
    }
    const float wp_v = __shfl_sync(0xffffffffu, jv, warp > 0 ? warp - 1 : 0);
    const bool wp_f = (fm & ((1u << warp) - 1u)) != 0u;
    const float ta_v = __shfl_sync(0xffffffffu, jv, WARPS - 1);
    const bool ta_f = fm != 0u;
    // the tile's outgoing carry (inclusive value of its last element) is final when the tile holds a head or its own
    // incoming carry was final; every thread knows it, so a CTA walking contiguous tiles hands it to the next tile
    const bool term = ta_f || resolved;
    const float val = ta_f ? ta_v : (resolved ? O::f(tp, ta_v) : ta_v);
    term_out = term;
    carry_out = val;
    if (warp == 0 && lane == 0) {
        uint32_t lead = WARPS * BLK_WSPAN;
        if (ta_f) {
            const int jw = __ffs(fm) - 1;
            lead = static_cast<uint32_t>(jw * BLK_WSPAN) + sh->fh[jw];
        }
        uint64_t *slot = desc + static_cast<int64_t>(tile) * 4;
        slot[0] = pack_desc(epoch, term ? ST_TERM : ST_AGG, ta_f ? 1u : 0u, val);
        slot[1] = static_cast<uint64_t>(lead);
        if (!resolved) ulist[atomicAdd(hdr + HDR_UCOUNT, 1u)] = tile;
        if (!ta_f) atomicAdd(hdr + HDR_INTERIOR, 1u);   // no head inside: the tile is the interior of a long segment
    }
    // ---- carry into the lane, applied up to its first head ----
    if (!cf) {
        if (warp > 0) c = O::f(wp_v, c);
        cf = wp_f;
    }
    if (!cf) c = O::f(resolved ? tp : O::id(), c);
    const int nfirst = hm ? (__ffs(hm) - 1) : 16;  // elements before the lane's first head take the carry
#pragma unroll
    for (int e = 0; e < 16; ++e) v[e] = (e < nfirst) ? O::f(c, v[e]) : v[e];
}

template <int WARPS, int STAGES>
struct FwdBlkSmem {
    static constexpr int TILE = WARPS * BLK_WSPAN;
    static constexpr int ARR_BYTES = TILE * 4;  // multiple of 1024 (swizzle atom = 8 rows x 128 B)
    static constexpr int STAGE_BYTES = 2 * ARR_BYTES;
    struct Ctl {
        uint64_t full[STAGES];
        uint64_t empty[STAGES];
        uint32_t tile[STAGES];
        int32_t halo[STAGES];
        uint32_t mode[STAGES];
        uint32_t resolved[STAGES];
        float tp[STAGES];
        uint32_t epoch;
        FwdBlkShared<WARPS> sh[2];
    };
    static constexpr int BYTES = STAGES * STAGE_BYTES + static_cast<int>(sizeof(Ctl)) + 1024;  // + alignment slack
};

template <int OP, int WARPS, int STAGES>
__global__ void __launch_bounds__((WARPS + 1) * 32)
k_fwd_blk(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_k,
          const float *__restrict__ x, const int32_t *__restrict__ key, float *__restrict__ y, int64_t n,
          uint32_t num_tiles, uint32_t *__restrict__ hdr, uint64_t *desc, uint32_t *ulist, int y_vec, int use_halo,
          int lead) {
    using L = FwdBlkSmem<WARPS, STAGES>;
    using O = ScanOp<OP>;
    constexpr int TILE = L::TILE;
    extern __shared__ unsigned char smem_raw[];
    // 1024-byte alignment for the 128B swizzle atom; plain pointer arithmetic keeps the shared address space
    unsigned char *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    typename L::Ctl *ctl = reinterpret_cast<typename L::Ctl *>(smem + STAGES * L::STAGE_BYTES);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // Chained mode (GCP_OPT_CHAIN_FWD, see k_bwd_blk): one contiguous ASCENDING range of tiles per CTA, each tile's
    // outgoing carry handed to the next in registers, the halo window read for the first tile of the range only.
    // Off by default: unlike the backward, the forward is 5-7 % slower with ranges than with tickets (C3 0.86 vs
    // 0.93, C4 0.90 vs 0.94 of the copy peak) — its fix-up is a cheap multiply of a leading run, so there is
    // little to win, and the compact window of the ticketed order is lost.
    const uint32_t hint = ld_relaxed_u32(hdr + HDR_HINT);
    const bool chain = (use_halo & 2) != 0 || ((use_halo & 4) != 0 && hint > num_tiles / 32u);
    use_halo &= 1;
    const uint32_t per_cta = num_tiles / gridDim.x, rem_cta = num_tiles % gridDim.x;
    const uint32_t first_ticket = blockIdx.x * per_cta + min(blockIdx.x, rem_cta);
    const uint32_t end_ticket = first_ticket + per_cta + (blockIdx.x < rem_cta ? 1u : 0u);

    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&ctl->full[s], 2);
            mbar_init(&ctl->empty[s], WARPS);
        }
        mbar_fence_init();
        ctl->epoch = ld_relaxed_u32(hdr + HDR_EPOCH);
    }
    __syncthreads();
    const uint32_t epoch = ctl->epoch;

    if (warp == WARPS) {
        // ===================== producer warp =====================
        // Software pipelined: the halo loads of tile i are issued right after its TMA copies and
        // consumed at the top of iteration i+1, so their latency overlaps the next tile's issue;
        // tickets are fetched two iterations ahead (lane 0 keeps q0 = ready, q1 = in flight).
        // One instantiation per mode (a run-time branch around the ticket fetch would expose the atomic's latency).
        auto produce = [&](auto chain_tag) {
        constexpr bool CHAIN = decltype(chain_tag)::value;
        const uint64_t pol = policy_evict_first();
        uint32_t q0 = 0, q1 = 0;
        if (lane == 0) {
            tma_prefetch_desc(&tm_x);
            tma_prefetch_desc(&tm_k);
            if (CHAIN) {   // tickets past the CTA's range read as "no tile left"
                q0 = first_ticket < end_ticket ? first_ticket : num_tiles;
                q1 = first_ticket + 1u < end_ticket ? first_ticket + 1u : num_tiles;
            } else {
                q0 = atomicAdd(hdr + HDR_TICKET, 1u);
                q1 = atomicAdd(hdr + HDR_TICKET, 1u);
            }
        }
        bool pend_window = false;
        bool pending = false;
        int ps = 0;
        int64_t pbase = 0;
        HaloPrefixRegs<BLK_HQ> hp;
        hp.kfirst = 0;
        for (uint32_t it = 0;; ++it) {
            if (pending) {
                float P = O::id();
                int32_t kprev = 0;
                bool res;
                if (pend_window) {
                    res = halo_prefix_finish<OP, BLK_HQ>(hp, lane, P, kprev);
                } else {
                    res = false;
                    kprev = hp.kfirst;
                }
                if (lane == 0) {
                    ctl->halo[ps] = kprev;
                    ctl->resolved[ps] = res ? 1u : 0u;
                    ctl->tp[ps] = P;
                    mbar_arrive(&ctl->full[ps]);
                }
                pending = false;
            }
            const int s = it % STAGES;
            const uint32_t ph = (it / STAGES) & 1u;
            if (lane == 0) mbar_wait(&ctl->empty[s], ph ^ 1u, hdr);
            const uint32_t t = __shfl_sync(0xffffffffu, q0, 0);
            if (t >= num_tiles) {
                if (lane == 0) {
                    ctl->tile[s] = t;
                    mbar_arrive(&ctl->full[s]);
                    mbar_arrive(&ctl->full[s]);
                }
                break;
            }
            const int64_t base = static_cast<int64_t>(t) * TILE;
            if (lane == 0) {
                ctl->tile[s] = t;
                if (base + TILE <= n) {
                    unsigned char *st = smem + s * L::STAGE_BYTES;
                    mbar_arrive_expect_tx(&ctl->full[s], L::STAGE_BYTES);
                    tma_load_2d(st, &tm_x, 0, static_cast<int>(base >> 5), &ctl->full[s], pol);
                    tma_load_2d(st + L::ARR_BYTES, &tm_k, 0, static_cast<int>(base >> 5), &ctl->full[s], pol);
                    ctl->mode[s] = 1u;
                } else {
                    ctl->mode[s] = 0u;
                    mbar_arrive(&ctl->full[s]);
                }
                q0 = q1;
                if (CHAIN) q1 = (q1 + 1u < end_ticket) ? q1 + 1u : num_tiles;
                else q1 = atomicAdd(hdr + HDR_TICKET, 1u);
            }
            if (t == 0u) {
                if (lane == 0) {
                    ctl->halo[s] = 0;
                    ctl->resolved[s] = 1u;
                    ctl->tp[s] = O::id();
                    mbar_arrive(&ctl->full[s]);
                }
            } else {
                // issue the halo loads now, use them next iteration (chained mode: only the first tile of the range
                // can use the window, the others take their carry from the tile before them)
                const bool window = use_halo != 0 && (!CHAIN || t == first_ticket);
                if (window) {
                    halo_prefix_issue<BLK_HQ>(x, key, base, lane, hp);
                } else {
                    hp.kfirst = __ldg(key + base - 1);
                }
                pend_window = window;
                pending = true;
                ps = s;
                pbase = base;
            }
        }
        (void)pbase;
        };
        if (chain) produce(std::true_type{});
        else produce(std::false_type{});
        return;
    }

    // ===================== consumers =====================
    auto consume = [&](auto chain_tag) {
    constexpr bool CHAIN = decltype(chain_tag)::value;
    bool chain_term = false;        // the tile walked last published a final outgoing carry ...
    float chain_carry = O::id();    // ... this one: the inclusive value of its last element
    uint32_t chain_tile = 0xffffffffu;
    for (uint32_t it = 0;; ++it) {
        const int s = it % STAGES;
        const uint32_t ph = (it / STAGES) & 1u;
        mbar_wait(&ctl->full[s], ph, hdr);
        const uint32_t tile = ctl->tile[s];
        if (tile >= num_tiles) break;
        const int64_t base = static_cast<int64_t>(tile) * TILE;
        const int64_t wbase = base + warp * BLK_WSPAN;
        bool resolved = ctl->resolved[s] != 0u;
        float tp_res = ctl->tp[s];
        if (CHAIN && !resolved && chain_term && chain_tile + 1u == tile) {
            resolved = true;
            tp_res = chain_carry;
        }
        const bool staged = ctl->mode[s] != 0u;
        unsigned char *xs = smem + s * L::STAGE_BYTES;
        const unsigned char *ks = xs + L::ARR_BYTES;
        float v[16];
        int32_t k[16];
        int32_t kprev = 0;
        if (staged) {
            lds_blocked<float>(xs, warp, lane, v);
            lds_blocked<int32_t>(ks, warp, lane, k);
            if (lane == 0) kprev = (warp == 0) ? ctl->halo[s] : lds_one<int32_t>(ks, warp * BLK_WSPAN - 1);
        } else {
            ldg_blocked<float>(x, wbase + lane * BLK_EPL, n, O::id(), v);
            ldg_blocked<int32_t>(key, wbase + lane * BLK_EPL, n, 0, k);
            if (lane == 0 && wbase > 0 && wbase - 1 < n) kprev = __ldg(key + wbase - 1);
        }
        const bool peel = lead != 0 && wbase == 0;   // warp-uniform: the warp that holds the phantom elements
        if (peel && lane == 0) {
            mask_lead<float>(v, lead, O::id());
            mask_lead_key<int32_t>(k, lead);
        }
        fwd_blk_compute<OP, WARPS>(v, k, kprev, wbase == 0, resolved, tp_res, tile, epoch, hdr, desc, ulist,
                                   &ctl->sh[it & 1u], warp, lane, chain_term, chain_carry);
        chain_tile = tile;
        if (staged && y_vec && !peel) {
            // the x array of the stage is dead (every warp has read its own span only): reuse it
            store_blocked_via_smem(xs, warp, lane, v, y + wbase);
        } else {
            stg_blocked_guarded(y, wbase + lane * BLK_EPL, n, v, lead);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&ctl->empty[s]);
    }
    };
    if (chain) consume(std::true_type{});
    else consume(std::false_type{});

    // ===================== fix-up phase (same launch) =====================
    grid_phase_barrier<WARPS * 32>(hdr, threadIdx.x);
    const uint32_t ucount = ld_relaxed_u32(hdr + HDR_UCOUNT);
    for (uint32_t u = blockIdx.x * WARPS + warp; u < ucount; u += gridDim.x * WARPS)
        fwd_fix_tile<OP>(__ldcg(ulist + u), y, n, TILE, epoch, desc, lane);
    named_bar_sync<WARPS * 32>(1);
    if (threadIdx.x == 0) finish_op(hdr, epoch);
}

// ============================================================================================
// backward
// ============================================================================================
template <int WARPS>
struct BwdBlkShared {
    float wv[WARPS];
    uint32_t wf[WARPS];
    float wa[WARPS];
    float wb[WARPS];
    int32_t lt[WARPS];
};

// out: grad_in of the lane's 16 elements.  PUBLISH = false: fix-up re-run of a tile (descriptors untouched).
template <int WARPS, bool PUBLISH>
__device__ __forceinline__ void bwd_blk_compute(const float (&x)[16], const float (&g)[16], const int32_t (&iv)[16],
                                                int32_t iprev, int32_t inext, float xnext, float y_prev,
                                                bool resolved, float rn, uint32_t tile, uint32_t epoch,
                                                uint32_t *__restrict__ hdr, uint64_t *__restrict__ desc,
                                                uint32_t *__restrict__ ulist, uint32_t *__restrict__ ulist2,
                                                BwdBlkShared<WARPS> *sh, int warp, int lane, float (&out)[16],
                                                bool &term_out, float &carry_out) {
    static_assert(WARPS < 32, "one lane per warp in the cross-warp step");
    const uint32_t lanes_lt = (1u << lane) - 1u;
    const uint32_t lanes_le = lanes_lt | (1u << lane);
    // ---- tail bits, head bits derived from them, x of the element after the lane ----
    const int src_lane = (lane + 1) & 31;
    const int32_t q = __shfl_sync(0xffffffffu, lane == 0 ? inext : iv[0], src_lane);  // lane 0 lends the warp halo
    const float xq = __shfl_sync(0xffffffffu, lane == 0 ? xnext : x[0], src_lane);
    uint32_t tm = (q != iv[15]) ? (1u << 15) : 0u;
#pragma unroll
    for (int e = 0; e < 15; ++e) tm |= (iv[e + 1] != iv[e] ? 1u : 0u) << e;
    const uint32_t m15 = __ballot_sync(0xffffffffu, (tm >> 15) != 0u);  // lanes whose LAST element is a tail
    const uint32_t mt = __ballot_sync(0xffffffffu, tm != 0u);           // lanes holding any tail
    const uint32_t h0 = lane ? ((m15 >> (lane - 1)) & 1u) : ((iv[0] != iprev) ? 1u : 0u);
    const uint32_t hm = (h0 | (tm << 1)) & 0xFFFFu;
    const uint32_t ahead = mt & ~lanes_lt;
    const int room = (ahead ? (__ffs(ahead) - 1) : 31) - lane;
    // ---- pass 1: lane aggregates ----
    float p = x[0];
#pragma unroll
    for (int e = 1; e < 16; ++e) p = ((hm >> e) & 1u) ? x[e] : p * x[e];
    float A = (tm >> 15) ? 0.0f : xq;
    float B = g[15];
#pragma unroll
    for (int e = 14; e >= 0; --e) {
        const bool t = (tm >> e) & 1u;
        B = t ? g[e] : fmaf(x[e + 1], B, g[e]);
        A = t ? 0.0f : x[e + 1] * A;
    }
    // ---- forward segmented warp scan of p ----
    const uint32_t mh = __ballot_sync(0xffffffffu, hm != 0u);
    const int start = max(31 - __clz(mh & lanes_le), 0);
    float finc = p;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const float t = __shfl_up_sync(0xffffffffu, finc, d);
        if (lane - d >= start) finc = t * finc;
    }
    float c = __shfl_up_sync(0xffffffffu, finc, 1);
    if (lane == 0) c = 1.0f;
    bool cf = (mh & lanes_lt) != 0u;
    // ---- reverse warp scan of (A,B), masked by the tail ballot ----
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const float ta_ = __shfl_down_sync(0xffffffffu, A, d);
        const float tb_ = __shfl_down_sync(0xffffffffu, B, d);
        if (d <= room) {
            B = fmaf(A, tb_, B);
            A = A * ta_;
        }
    }
    Affine sx;
    sx.a = __shfl_down_sync(0xffffffffu, A, 1);
    sx.b = __shfl_down_sync(0xffffffffu, B, 1);
    if (lane == 31) sx = affine_id();
    if (lane == 31) {
        sh->wv[warp] = finc;
        sh->wf[warp] = mh != 0u ? 1u : 0u;
    }
    if (lane == 0) {
        sh->wa[warp] = A;
        sh->wb[warp] = B;
    }
    {
        int32_t lt = -1;
        if (mt) {  // warp-uniform
            const int l1 = 31 - __clz(mt);
            const uint32_t t1 = __shfl_sync(0xffffffffu, tm, l1);
            lt = l1 * BLK_EPL + (31 - __clz(t1));
        }
        if (lane == 0) sh->lt[warp] = lt;
    }
    named_bar_sync<WARPS * 32>(1);
    // ---- across warps, one lane per warp ----
    const bool wl = lane < WARPS;
    float jv = wl ? sh->wv[lane] : 1.0f;
    const uint32_t jf = wl ? sh->wf[lane] : 0u;
    Affine jm = wl ? Affine{sh->wa[lane], sh->wb[lane]} : affine_id();
    const int32_t jl = wl ? sh->lt[lane] : -1;
    const uint32_t fm = __ballot_sync(0xffffffffu, jf != 0u);
    {
        const int st = max(31 - __clz(fm & lanes_le), 0);
#pragma unroll
        for (int d = 1; d < WARPS; d <<= 1) {
            const float tv = __shfl_up_sync(0xffffffffu, jv, d);
            if (lane - d >= st) jv = tv * jv;
        }
#pragma unroll
        for (int d = 1; d < WARPS; d <<= 1) {
            Affine r;
            r.a = __shfl_down_sync(0xffffffffu, jm.a, d);
            r.b = __shfl_down_sync(0xffffffffu, jm.b, d);
            if (lane + d < 32) jm = compose(jm, r);
        }
    }
    const float wp_v = __shfl_sync(0xffffffffu, jv, warp > 0 ? warp - 1 : 0);
    const bool wp_f = (fm & ((1u << warp) - 1u)) != 0u;
    Affine ws, ta;
    ws.a = __shfl_sync(0xffffffffu, jm.a, warp + 1);  // lane WARPS holds the identity
    ws.b = __shfl_sync(0xffffffffu, jm.b, warp + 1);
    ta.a = __shfl_sync(0xffffffffu, jm.a, 0);
    ta.b = __shfl_sync(0xffffffffu, jm.b, 0);
    const uint32_t lm = __ballot_sync(0xffffffffu, jl >= 0);
    uint32_t trail = 0u;
    if (lm) {
        const int jw = 31 - __clz(lm);
        trail = static_cast<uint32_t>(jw * BLK_WSPAN + __shfl_sync(0xffffffffu, jl, jw) + 1);
    }
    // the tile's outgoing carry (S of its first element) is final when its own incoming carry was, or when the tile
    // holds a tail; every thread knows it (ta is broadcast), so a CTA walking contiguous tiles can hand it to the
    // next (lower) tile without going through the descriptors
    const bool term = resolved || (ta.a == 0.0f);
    term_out = term;
    carry_out = resolved ? apply(ta, rn) : ta.b;
    if (PUBLISH && warp == 0 && lane == 0) {
        uint64_t *slot = desc + static_cast<int64_t>(tile) * 4;
        if (ta.a != 0.0f) atomicAdd(hdr + HDR_INTERIOR, 1u);   // no tail inside: interior of a long segment
        slot[0] = term ? pack_desc(epoch, ST_TERM, 0u, carry_out)
                       : pack_desc(epoch, ST_AGG, 0u, ta.a);
        slot[1] = static_cast<uint64_t>(trail);
        slot[3] = static_cast<uint64_t>(__float_as_uint(ta.b));
        if (!resolved) {
            if (WARPS * BLK_WSPAN - trail > LONG_RUN) ulist2[atomicAdd(hdr + HDR_UCOUNT2, 1u)] = tile;
            else ulist[atomicAdd(hdr + HDR_UCOUNT, 1u)] = tile;
        }
    }
    // ---- pass 2 ----
    float s = apply(sx, apply(ws, resolved ? rn : 0.0f));  // S of the element right after the lane
    if (!cf) {
        if (warp > 0) c = wp_v * c;
        cf = wp_f;
    }
    if (!cf) c = y_prev * c;
    // reverse: S_e ; forward: E_e.  Two independent chains.
    float sv[16];
    sv[15] = (tm >> 15) ? g[15] : fmaf(xq, s, g[15]);
#pragma unroll
    for (int e = 14; e >= 0; --e) sv[e] = ((tm >> e) & 1u) ? g[e] : fmaf(x[e + 1], sv[e + 1], g[e]);
    float ev = (hm & 1u) ? 1.0f : c;
    out[0] = ev * sv[0];
#pragma unroll
    for (int e = 1; e < 16; ++e) {
        ev = ((hm >> e) & 1u) ? 1.0f : ev * x[e - 1];
        out[e] = ev * sv[e];
    }
}

template <int WARPS, int STAGES>
struct BwdBlkSmem {
    static constexpr int TILE = WARPS * BLK_WSPAN;
    static constexpr int ARR_BYTES = TILE * 4;
    static constexpr int STAGE_BYTES = 3 * ARR_BYTES;
    struct Ctl {
        uint64_t full[STAGES];
        uint64_t empty[STAGES];
        uint32_t tile[STAGES];
        uint32_t mode[STAGES];
        int32_t iprev[STAGES];
        int32_t inext[STAGES];
        float xnext[STAGES];
        float yprev[STAGES];
        uint32_t resolved[STAGES];
        float rn[STAGES];
        uint32_t epoch;
        BwdBlkShared<WARPS> sh[2];
    };
    static constexpr int BYTES = STAGES * STAGE_BYTES + static_cast<int>(sizeof(Ctl)) + 1024;
};

template <int WARPS, int STAGES, int MINB>
__global__ void __launch_bounds__((WARPS + 1) * 32, MINB)
k_bwd_blk(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_g,
          const __grid_constant__ CUtensorMap tm_i, const float *__restrict__ x, const float *__restrict__ y,
          const float *__restrict__ g, const int32_t *__restrict__ inv, float *__restrict__ gin, int64_t n,
          uint32_t num_tiles, uint32_t *__restrict__ hdr, uint64_t *desc, uint32_t *ulist, int out_vec,
          int use_halo, int lead) {
    using L = BwdBlkSmem<WARPS, STAGES>;
    constexpr int TILE = L::TILE;
    extern __shared__ unsigned char smem_raw[];
    // 1024-byte alignment for the 128B swizzle atom; plain pointer arithmetic keeps the shared address space
    unsigned char *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    typename L::Ctl *ctl = reinterpret_cast<typename L::Ctl *>(smem + STAGES * L::STAGE_BYTES);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // Chained mode (GCP_OPT_CHAIN, the default): every CTA walks ONE CONTIGUOUS range of tiles (descending) instead
    // of taking tickets, and hands each tile's outgoing carry to the next one in registers.
    //  * Inside a segment of many tiles only the first tile of a CTA's range is left for the fix-up phase, not
    //    every tile of the segment (C4: backward 0.73 -> 0.95 of the copy peak).
    //  * Only the first tile of a range needs the 512-element halo window; the others load two boundary values.
    //    The window is 12 % of a tile's reads — L2 hits with tickets (the neighbour tile is in flight in another
    //    CTA), DRAM reads with ranges — and dropping it is what makes ranges faster than tickets on short lists
    //    too (C3: 0.932 vs 0.925); with the windows kept, ranges were 5 % slower there.
    // use_halo bit 1 = chained, bit 2 = chained only when the op that ran last on this workspace (normally the
    // forward over the same list) had more than 1/32 of its tiles strictly inside a segment, 0 = tickets.  The mode
    // only picks a schedule; results do not depend on it beyond fp32 rounding.
    // (Tickets for RUNS of 8 contiguous tiles do not work: a run that starts inside a long segment has no resolved
    // carry to hand on, so every tile of the segment stays unresolved as with single-tile tickets — measured 0.73.
    // The range of a CTA has to be long against the segments, which is what one range per CTA gives.)
    const uint32_t hint = ld_relaxed_u32(hdr + HDR_HINT);   // rewritten only by the last CTA out
    const bool chain = (use_halo & 2) != 0 || ((use_halo & 4) != 0 && hint > num_tiles / 32u);
    use_halo &= 1;
    // balanced contiguous ranges: the first (num_tiles % grid) CTAs take one tile more
    const uint32_t per_cta = num_tiles / gridDim.x, rem_cta = num_tiles % gridDim.x;
    const uint32_t first_ticket = blockIdx.x * per_cta + min(blockIdx.x, rem_cta);
    const uint32_t end_ticket = first_ticket + per_cta + (blockIdx.x < rem_cta ? 1u : 0u);

    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&ctl->full[s], 2);
            mbar_init(&ctl->empty[s], WARPS);
        }
        mbar_fence_init();
        ctl->epoch = ld_relaxed_u32(hdr + HDR_EPOCH);
    }
    __syncthreads();
    const uint32_t epoch = ctl->epoch;

    if (warp == WARPS) {
        // ===================== producer warp =====================
        // Instantiated per mode like the consumer loop below: with a run-time branch around the ticket fetch the
        // atomic's result is needed where the two paths merge, i.e. at once, instead of two iterations later —
        // 1 us of exposed latency per tile in the producer (C3 backward 0.1855 -> 0.1926 ms).
        auto produce = [&](auto chain_tag) {
            constexpr bool CHAIN = decltype(chain_tag)::value;
            // Software pipelined like the forward's: halo loads issued in iteration i are consumed at
            // the top of iteration i+1; tickets are fetched two iterations ahead.
            const uint64_t pol = policy_evict_first();
            uint32_t q0 = 0, q1 = 0;
            if (lane == 0) {
                tma_prefetch_desc(&tm_x);
                tma_prefetch_desc(&tm_g);
                tma_prefetch_desc(&tm_i);
                if (CHAIN) {   // tickets past the CTA's range read as "no tile left"
                    q0 = first_ticket < end_ticket ? first_ticket : num_tiles;
                    q1 = first_ticket + 1u < end_ticket ? first_ticket + 1u : num_tiles;
                } else {
                    q0 = atomicAdd(hdr + HDR_TICKET, 1u);
                    q1 = atomicAdd(hdr + HDR_TICKET, 1u);
                }
            }
            bool pending = false, pend_window = false;
            int ps = 0;
            HaloSuffixRegs<BLK_HQ> hr;
            int32_t ip = -1;
            float yp = 1.0f;
            for (uint32_t it = 0;; ++it) {
                if (pending) {
                    float R = 0.0f, xq = 0.0f;
                    int32_t in = -1;
                    const bool res = halo_suffix_finish(hr, lane, pend_window, R, in, xq);
                    if (lane == 0) {
                        ctl->iprev[ps] = ip;
                        ctl->yprev[ps] = yp;
                        ctl->inext[ps] = in;
                        ctl->xnext[ps] = xq;
                        ctl->resolved[ps] = res ? 1u : 0u;
                        ctl->rn[ps] = R;
                        mbar_arrive(&ctl->full[ps]);
                    }
                    pending = false;
                }
                const int s = it % STAGES;
                const uint32_t ph = (it / STAGES) & 1u;
                if (lane == 0) mbar_wait(&ctl->empty[s], ph ^ 1u, hdr);
                const uint32_t t = __shfl_sync(0xffffffffu, q0, 0);
                if (t >= num_tiles) {
                    if (lane == 0) {
                        ctl->tile[s] = t;
                        mbar_arrive(&ctl->full[s]);
                        mbar_arrive(&ctl->full[s]);
                    }
                    break;
                }
                const uint32_t tile = num_tiles - 1u - t;
                const int64_t base = static_cast<int64_t>(tile) * TILE;
                const int64_t end = base + TILE;
                if (lane == 0) {
                    ctl->tile[s] = t;
                    if (end <= n) {
                        unsigned char *st = smem + s * L::STAGE_BYTES;
                        const int row0 = static_cast<int>(base >> 5);
                        mbar_arrive_expect_tx(&ctl->full[s], L::STAGE_BYTES);
                        tma_load_2d(st, &tm_x, 0, row0, &ctl->full[s], pol);
                        tma_load_2d(st + L::ARR_BYTES, &tm_g, 0, row0, &ctl->full[s], pol);
                        tma_load_2d(st + 2 * L::ARR_BYTES, &tm_i, 0, row0, &ctl->full[s], pol);
                        ctl->mode[s] = 1u;
                    } else {
                        ctl->mode[s] = 0u;
                        mbar_arrive(&ctl->full[s]);
                    }
                    q0 = q1;
                    if (CHAIN) q1 = (q1 + 1u < end_ticket) ? q1 + 1u : num_tiles;
                    else q1 = atomicAdd(hdr + HDR_TICKET, 1u);
                    ip = -1;
                    yp = 1.0f;
                    if (base > 0) {
                        ip = __ldg(inv + base - 1);
                        yp = __ldg(y + base - 1);
                    }
                }
                // chained mode: only the first tile of the CTA's range can make use of the halo window — every other
                // tile takes its carry from the tile above it (or, inside a segment that began above the range, has
                // no tail within the window anyway); the others load just the two boundary values
                const bool window = use_halo != 0 && (!CHAIN || t == first_ticket);
                halo_suffix_issue(x, g, inv, end, n, lane, window, hr);
                pend_window = window;
                pending = true;
                ps = s;
            }
        };
        if (chain) produce(std::true_type{});
        else produce(std::false_type{});
        return;
    }

    // ===================== consumers =====================
    // Two instantiations of the same loop: the ticketed one carries no chain state at all (it is the kernel of the
    // short-list workloads, where a few instructions per tile are measurable), the chained one keeps the previous
    // tile's outgoing carry in registers.
    auto consume = [&](auto chain_tag) {
        constexpr bool CHAIN = decltype(chain_tag)::value;
        bool chain_term = false;        // the tile walked last published a final outgoing carry ...
        float chain_carry = 0.0f;       // ... this one: S of its first element = the incoming carry of the tile below it
        uint32_t chain_tile = 0xffffffffu;
        for (uint32_t it = 0;; ++it) {
            const int s = it % STAGES;
            const uint32_t ph = (it / STAGES) & 1u;
            mbar_wait(&ctl->full[s], ph, hdr);
            const uint32_t ticket = ctl->tile[s];
            if (ticket >= num_tiles) break;
            const uint32_t tile = num_tiles - 1u - ticket;
            const int64_t base = static_cast<int64_t>(tile) * TILE;
            const int64_t wbase = base + warp * BLK_WSPAN;
            const int64_t wend = wbase + BLK_WSPAN;
            const float y_prev = ctl->yprev[s];
            bool resolved = ctl->resolved[s] != 0u;
            float rn_res = ctl->rn[s];
            if (CHAIN && !resolved && chain_term && chain_tile == tile + 1u) {
                resolved = true;        // the halo window did not reach a tail, but the tile above was walked by this CTA
                rn_res = chain_carry;
            }
            const bool staged = ctl->mode[s] != 0u;
            unsigned char *xs = smem + s * L::STAGE_BYTES;
            unsigned char *gs = xs + L::ARR_BYTES;
            const unsigned char *is = xs + 2 * L::ARR_BYTES;
            float xv[16], gv[16];
            int32_t iv[16];
            int32_t iprev = -1, inext = -1;
            float xnext = 0.0f;
            if (staged) {
                lds_blocked<float>(xs, warp, lane, xv);
                lds_blocked<float>(gs, warp, lane, gv);
                lds_blocked<int32_t>(is, warp, lane, iv);
                if (lane == 0) {
                    iprev = (warp == 0) ? ctl->iprev[s] : lds_one<int32_t>(is, warp * BLK_WSPAN - 1);
                    if (warp == WARPS - 1) {
                        inext = ctl->inext[s];
                        xnext = ctl->xnext[s];
                    } else {
                        inext = lds_one<int32_t>(is, (warp + 1) * BLK_WSPAN);
                        xnext = lds_one<float>(xs, (warp + 1) * BLK_WSPAN);
                    }
                }
            } else {
                ldg_blocked<float>(x, wbase + lane * BLK_EPL, n, 1.0f, xv);
                ldg_blocked<float>(g, wbase + lane * BLK_EPL, n, 0.0f, gv);
                ldg_blocked<int32_t>(inv, wbase + lane * BLK_EPL, n, -1, iv);
                if (lane == 0 && wbase > 0 && wbase - 1 < n) iprev = __ldg(inv + wbase - 1);
                if (lane == 0 && wend < n) {
                    inext = __ldg(inv + wend);
                    xnext = __ldg(x + wend);
                }
            }
            const bool peel = lead != 0 && wbase == 0;   // warp-uniform: the warp that holds the phantom elements
            if (peel && lane == 0) {
                mask_lead<float>(xv, lead, 1.0f);
                mask_lead<float>(gv, lead, 0.0f);
                mask_lead_key<int32_t>(iv, lead);
            }
            float out[16];
            bwd_blk_compute<WARPS, true>(xv, gv, iv, iprev, inext, xnext, y_prev, resolved, rn_res, tile, epoch, hdr, desc,
                                         ulist, ulist + num_tiles, &ctl->sh[it & 1u], warp, lane, out, chain_term,
                                         chain_carry);
            chain_tile = tile;
            if (staged && out_vec && !peel) {
                // the g array of the stage is only ever read by the warp that owns the span: reuse it
                // (x is read across warp boundaries for x_next, inv for the head/tail halos)
                store_blocked_via_smem(gs, warp, lane, out, gin + wbase);
            } else {
                stg_blocked_guarded(gin, wbase + lane * BLK_EPL, n, out, lead);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&ctl->empty[s]);
        }
    };
    if (chain) consume(std::true_type{});
    else consume(std::false_type{});

    // ===================== fix-up phase (same launch) =====================
    grid_phase_barrier<WARPS * 32>(hdr, threadIdx.x);
    // (a) short trailing runs: one warp per tile recomputes just the run
    const uint32_t ucount = ld_relaxed_u32(hdr + HDR_UCOUNT);
    for (uint32_t u = blockIdx.x * WARPS + warp; u < ucount; u += gridDim.x * WARPS)
        bwd_fix_tile(static_cast<int64_t>(__ldcg(ulist + u)), x, y, g, inv, gin, n, num_tiles, TILE, epoch, desc, lane,
                     lead);
    // (b) long trailing runs (tiles inside segments of thousands of elements): warp 0 finds R by
    //     walking the descriptors, then the whole CTA re-runs the tile with R known.
    const uint32_t ucount2 = ld_relaxed_u32(hdr + HDR_UCOUNT2);
    unsigned char *gs0 = smem + L::ARR_BYTES;  // every stage is idle now: staging area for the stores
    for (uint32_t u = blockIdx.x, k2 = 0; u < ucount2; u += gridDim.x, ++k2) {
        const int64_t t = static_cast<int64_t>(__ldcg(ulist + num_tiles + u));
        float *rbox = reinterpret_cast<float *>(&ctl->rn[0]);
        if (warp == 0) {
            const float R = bwd_fix_walk(t, num_tiles, epoch, desc, lane);
            if (lane == 0) *rbox = R;
        }
        named_bar_sync<WARPS * 32>(1);
        const float R = *rbox;
        const int64_t base = t * TILE;  // t <= num_tiles-2: a full tile with a successor
        const int64_t wbase = base + warp * BLK_WSPAN;
        float xv[16], gv[16], out[16];
        int32_t iv[16];
        ldg_blocked_vec<float>(x, wbase + lane * BLK_EPL, xv);
        ldg_blocked_vec<float>(g, wbase + lane * BLK_EPL, gv);
        ldg_blocked_vec<int32_t>(inv, wbase + lane * BLK_EPL, iv);
        int32_t iprev = -1, inext = -1;
        float xnext = 0.0f;
        if (lane == 0) {
            if (wbase > 0) iprev = __ldg(inv + wbase - 1);
            inext = __ldg(inv + wbase + BLK_WSPAN);
            xnext = __ldg(x + wbase + BLK_WSPAN);
        }
        const float y_prev = (base > 0) ? __ldg(y + base - 1) : 1.0f;
        const bool peel = lead != 0 && wbase == 0;
        if (peel && lane == 0) {
            mask_lead<float>(xv, lead, 1.0f);
            mask_lead<float>(gv, lead, 0.0f);
            mask_lead_key<int32_t>(iv, lead);
        }
        bool term_unused;
        float carry_unused;
        bwd_blk_compute<WARPS, false>(xv, gv, iv, iprev, inext, xnext, y_prev, true, R, static_cast<uint32_t>(t), epoch,
                                      hdr, desc, ulist, ulist, &ctl->sh[k2 & 1u], warp, lane, out, term_unused,
                                      carry_unused);
        if (out_vec && !peel) {
            store_blocked_via_smem(gs0, warp, lane, out, gin + wbase);
        } else {
            stg_blocked_guarded(gin, wbase + lane * BLK_EPL, n, out, lead);
        }
    }
    named_bar_sync<WARPS * 32>(1);
    if (threadIdx.x == 0) finish_op(hdr, epoch);
}

}  // namespace gcp
