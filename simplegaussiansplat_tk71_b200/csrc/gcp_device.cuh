// gcp_device.cuh — device-side building blocks shared by the scan kernels (sm_100a).
//
//  * PTX wrappers: mbarrier, relaxed gpu-scope descriptor loads/stores, streaming vector
//    loads/stores.
//  * workspace layout + per-tile descriptors.
//  * scan operators (segmented product/sum, affine maps of the backward).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace gcp {

// ----------------------------------------------------------------------------
// Workspace layout (device memory, caller owned; see include/gcp_abi.h)
//   [0,256)    header: u32 words  ticket@0, done@16, epoch@32, abort@33, violations(u64)@byte 160,
//              ucount@48 (number of unresolved tiles), exit@56, host flag address (u64)@byte 240
//   [256, ...) one 32-byte slot (4 x u64) per tile of >= MIN_TILE elements, then the list of
//              unresolved tile indices (u32 per tile):
//     word0  K1: carry descriptor of the tile   {epoch, status TERM|AGG, flag, f32 value}
//     word1  K1: fix-up request                 {bit31 = needs fix-up, low 31 bits = run boundary}
//     word2  K2: inclusive carry of an AGG tile {epoch, ST_INCL, f32 value}  (epoch-tagged: never cleared)
//     word3  K1 (backward only): the `b` of an AGG tile's affine aggregate
// ----------------------------------------------------------------------------
constexpr int WS_HEADER_BYTES = 256;
constexpr int WS_SLOT_BYTES = 32;
constexpr int MIN_TILE = 1024;
constexpr int HDR_TICKET = 0;
constexpr int HDR_DONE = 16;
constexpr int HDR_EPOCH = 32;
constexpr int HDR_ABORT = 33;
constexpr int HDR_INTERIOR = 40;  // blocked kernels: tiles of this op lying strictly inside one segment (no boundary)
constexpr int HDR_HINT = 41;      // that count for the op that finished last on this workspace (survives finish_op)
constexpr int HDR_VIOL64 = 20;  // index in u64 units (byte 160)
constexpr int HDR_UCOUNT = 48;
constexpr int HDR_UCOUNT2 = 52;  // backward: unresolved tiles with a long trailing run (fixed by a whole CTA)
constexpr int HDR_EXIT = 56;
constexpr int HDR_HOSTFLAG64 = 30;  // index in u64 units (byte 240): device-visible address of a HOST word (pinned,
                                    // mapped) that a tripped watchdog sets to 1, or 0 (gcp_workspace_attach_flag)
constexpr int WS_LIST_BYTES = 8;  // per tile: two u32 lists
constexpr uint32_t LONG_RUN = 1024;  // trailing runs longer than this are recomputed by a whole CTA

constexpr uint32_t ST_INVALID = 0, ST_AGG = 1, ST_TERM = 2, ST_INCL = 3;
constexpr uint32_t EPOCH_MASK = 0x1FFFFFFFu;
constexpr uint32_t WAIT_LIMIT = 1u << 20;  // bounded spins: never hang; expiry raises the sticky abort flag
constexpr uint32_t FIX_FLAG = 0x80000000u;

__device__ __forceinline__ uint64_t pack_desc(uint32_t epoch, uint32_t status, uint32_t flag, float v) {
    uint32_t hi = ((epoch & EPOCH_MASK) << 3) | (status << 1) | (flag & 1u);
    return (static_cast<uint64_t>(hi) << 32) | static_cast<uint64_t>(__float_as_uint(v));
}
__device__ __forceinline__ bool desc_valid(uint64_t d, uint32_t epoch) {
    uint32_t hi = static_cast<uint32_t>(d >> 32);
    return ((hi >> 3) == (epoch & EPOCH_MASK)) && (((hi >> 1) & 3u) != ST_INVALID);
}
__device__ __forceinline__ uint32_t desc_status(uint64_t d) { return (static_cast<uint32_t>(d >> 32) >> 1) & 3u; }
__device__ __forceinline__ uint32_t desc_flag(uint64_t d) { return static_cast<uint32_t>(d >> 32) & 1u; }
__device__ __forceinline__ float desc_value(uint64_t d) { return __uint_as_float(static_cast<uint32_t>(d)); }

__device__ __forceinline__ uint64_t ld_relaxed_u64(const uint64_t *p) {
    uint64_t v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_u64(uint64_t *p, uint64_t v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_relaxed_u32(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// ----------------------------------------------------------------------------
// mbarrier + bulk async copy
// ----------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// A bounded wait expired.  Sets the sticky flag in the workspace header AND, when the owner of the workspace
// attached one (gcp_workspace_attach_flag), a word in pinned host memory, so that the host learns about it
// without a synchronising read: the next op on that workspace refuses to launch (GCP_ERR_WATCHDOG) instead of
// producing garbage silently.  Only a bug can trip it: the kernels wait on mbarriers fed by the CTA's own
// producer warp, and on the grid barrier of a COOPERATIVE launch (all CTAs resident by contract).
__device__ __noinline__ void signal_abort(uint32_t *hdr) {
    if (atomicExch(hdr + HDR_ABORT, 1u) == 0u) {
        const uint64_t host = reinterpret_cast<const uint64_t *>(hdr)[HDR_HOSTFLAG64];
        if (host != 0ull) {
            *reinterpret_cast<volatile uint32_t *>(host) = 1u;
            __threadfence_system();
        }
    }
}
// Bounded wait: on expiry signal_abort and carry on (the results of THIS op are then wrong, the kernel
// terminates, and the flag makes every later use of the workspace fail loudly).
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity, uint32_t *hdr) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        ++spins;
        if ((spins & 63u) == 0u) {
            if (spins >= WAIT_LIMIT) { signal_abort(hdr); break; }
            if (ld_relaxed_u32(hdr + HDR_ABORT) != 0u) break;
        }
    }
}
__device__ __forceinline__ uint64_t policy_evict_first() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
template <int COUNT>
__device__ __forceinline__ void named_bar_sync(int id) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(COUNT) : "memory");
}

// ----------------------------------------------------------------------------
// streaming global access
// ----------------------------------------------------------------------------
__device__ __forceinline__ float4 ldcs_f4(const float *p) { return __ldcs(reinterpret_cast<const float4 *>(p)); }
__device__ __forceinline__ int4 ldcs_i4(const int32_t *p) { return __ldcs(reinterpret_cast<const int4 *>(p)); }
__device__ __forceinline__ void stcs_f4(float *p, float a, float b, float c, float d) {
    __stcs(reinterpret_cast<float4 *>(p), make_float4(a, b, c, d));
}

// ----------------------------------------------------------------------------
// scan operators
// ----------------------------------------------------------------------------
constexpr int OP_MUL = 0;
constexpr int OP_ADD = 1;
template <int OP>
struct ScanOp;
template <>
struct ScanOp<OP_MUL> {
    static __device__ __forceinline__ float id() { return 1.0f; }
    static __device__ __forceinline__ float f(float a, float b) { return a * b; }
};
template <>
struct ScanOp<OP_ADD> {
    static __device__ __forceinline__ float id() { return 0.0f; }
    static __device__ __forceinline__ float f(float a, float b) { return a + b; }
};

// Affine map S -> b + a*S (reverse scan of the backward).  a == 0 is a hard reset:
// the right operand is then ignored even if it is Inf/NaN (tail of a segment).
struct Affine {
    float a, b;
};
__device__ __forceinline__ Affine affine_id() { return Affine{1.0f, 0.0f}; }
// (l ∘ r)(S) = l(r(S)) : l is nearer (lower index), r is further (higher index)
__device__ __forceinline__ Affine compose(Affine l, Affine r) {
    Affine o;
    o.b = (l.a == 0.0f) ? l.b : fmaf(l.a, r.b, l.b);
    o.a = (l.a == 0.0f) ? 0.0f : l.a * r.a;
    return o;
}
__device__ __forceinline__ float apply(Affine m, float s) { return (m.a == 0.0f) ? m.b : fmaf(m.a, s, m.b); }

// Ordered warp reduction of affine maps: lane 0 ends with m_0 ∘ m_1 ∘ ... ∘ m_31, broadcast to all.
__device__ __forceinline__ Affine warp_compose_all(Affine m, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        Affine t;
        t.a = __shfl_down_sync(0xffffffffu, m.a, d);
        t.b = __shfl_down_sync(0xffffffffu, m.b, d);
        if (lane + d < 32) m = compose(m, t);
    }
    m.a = __shfl_sync(0xffffffffu, m.a, 0);
    m.b = __shfl_sync(0xffffffffu, m.b, 0);
    return m;
}

// Last-CTA-out protocols.  The workspace resets itself: no memset between launches.
//  * finish_stream_kernel: end of a K1 that is followed by a separate K2 (LDG path): the last CTA
//    clears the ticket/done counters; the unresolved list and the epoch are left for K2.
//  * finish_op: end of an op (K2, or the persistent kernel that contains its own fix-up phase): the
//    last CTA clears every counter and advances the epoch (which invalidates all word2 entries).
__device__ __forceinline__ void finish_stream_kernel(uint32_t *hdr) {
    __threadfence();
    const uint32_t prev = atomicAdd(hdr + HDR_DONE, 1u);
    if (prev == gridDim.x - 1u) {
        hdr[HDR_TICKET] = 0u;
        hdr[HDR_DONE] = 0u;
        __threadfence();
    }
}
__device__ __forceinline__ void finish_op(uint32_t *hdr, uint32_t epoch) {
    __threadfence();
    const uint32_t prev = atomicAdd(hdr + HDR_EXIT, 1u);
    if (prev == gridDim.x - 1u) {
        hdr[HDR_TICKET] = 0u;
        hdr[HDR_DONE] = 0u;
        hdr[HDR_UCOUNT] = 0u;
        hdr[HDR_UCOUNT2] = 0u;
        hdr[HDR_EXIT] = 0u;
        hdr[HDR_HINT] = hdr[HDR_INTERIOR];   // a property of the segment layout, not of how the op resolved carries
        hdr[HDR_INTERIOR] = 0u;
        hdr[HDR_EPOCH] = epoch + 1u;
        __threadfence();
    }
}

// Grid-wide barrier between the streaming phase and the fix-up phase of a persistent kernel.
// Called by all consumer threads of the CTA (COUNT of them, named barrier 1).  The persistent kernels
// are launched COOPERATIVELY (cudaLaunchAttributeCooperative, gcp_abi.cu): the runtime starts the grid
// only when every CTA can be resident at once — also when another stream, an MPS limit or a long
// kernel holds SMs — so the barrier cannot starve; where a cooperative launch is refused the host
// falls back to the two-kernel LDG path, which has no cross-CTA wait at all.  Bounded like every spin.
template <int COUNT>
__device__ __forceinline__ void grid_phase_barrier(uint32_t *hdr, int tid) {
    __threadfence();
    named_bar_sync<COUNT>(1);
    if (tid == 0) {
        atomicAdd(hdr + HDR_DONE, 1u);
        uint32_t spins = 0;
        while (ld_relaxed_u32(hdr + HDR_DONE) < gridDim.x) {
            __nanosleep(64);
            if (((++spins) & 63u) == 0u) {
                if (spins >= WAIT_LIMIT) { signal_abort(hdr); break; }
                if (ld_relaxed_u32(hdr + HDR_ABORT) != 0u) break;
            }
        }
        __threadfence();
    }
    named_bar_sync<COUNT>(1);
}

}  // namespace gcp
