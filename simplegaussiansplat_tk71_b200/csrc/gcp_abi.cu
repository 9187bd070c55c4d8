// gcp_abi.cu — C-ABI entry points of libgcp_b200.so (see include/gcp_abi.h) and the
// host-side launch logic: the two paths (persistent blocked kernel, cooperative launch | LDG K1 + K2),
// grid sizing (persistent = resident CTAs x 148 SMs), the alignment peel, workspace checks.
// No allocation, no host sync on the compute entry points.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <utility>

#include "gcp_abi.h"
#include "gcp_blk.cuh"
#include "gcp_bwd.cuh"
#include "gcp_fwd.cuh"

namespace {

using namespace gcp;

thread_local int t_launches = 0;
int g_variant[2] = {-1, -1};
int g_option[4] = {1, 1, 0, 0};  // [0] = resolve tile carries from the halo window (1) or always look back (0)
                                 // [1] = blocked backward, contiguous tile range per CTA with carries chained in registers:
                                 //       0 never, 1 always (default), 2 when the last op on the workspace saw long segments
                                 // [2] = the same for the blocked forward; default 0: there tickets are 5-7 % faster

struct DeviceInfo {
    bool init = false;
    int sms = 0;
};
DeviceInfo g_dev[64];

int current_device() {
    int d = 0;
    cudaGetDevice(&d);
    return d;
}
int sm_count() {
    const int d = current_device();
    DeviceInfo &di = g_dev[d & 63];
    if (!di.init) {
        cudaDeviceGetAttribute(&di.sms, cudaDevAttrMultiProcessorCount, d);
        di.init = true;
    }
    return di.sms;
}

inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

struct Ws {
    uint32_t *hdr;
    uint64_t *desc;
    // lists of unresolved tiles (two, num_tiles entries each): right behind the descriptors of THIS launch's tile count
    uint32_t *ulist(uint32_t num_tiles) const { return reinterpret_cast<uint32_t *>(desc + 4ull * num_tiles); }
};
int check_ws(void *ws, size_t ws_bytes, int64_t n, Ws *out) {
    if (ws == nullptr || (reinterpret_cast<uintptr_t>(ws) & 15u) != 0) return GCP_ERR_WORKSPACE;
    if (ws_bytes < gcp_workspace_bytes(n)) return GCP_ERR_WORKSPACE;
    out->hdr = reinterpret_cast<uint32_t *>(ws);
    out->desc = reinterpret_cast<uint64_t *>(reinterpret_cast<unsigned char *>(ws) + WS_HEADER_BYTES);
    return GCP_OK;
}

// resident CTAs per SM for a persistent kernel, cached per (kernel, device).  The kernel is a
// non-type template parameter: distinct instantiations share one function-pointer TYPE, so a cache
// keyed on the type alone would skip the shared-memory opt-in of all but the first of them.
template <auto kernel>
int persistent_ctas_per_sm(int threads, size_t smem) {
    static int cache[64];
    static bool have[64];
    const int d = current_device() & 63;
    if (!have[d]) {
        if (smem > 48 * 1024) cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        int nb = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kernel, threads, smem);
        cache[d] = nb > 0 ? nb : 1;
        have[d] = true;
    }
    return cache[d];
}

inline uint32_t tiles_for(int64_t n, int tile) { return static_cast<uint32_t>((n + tile - 1) / tile); }

// ---- 2-D tensor maps ([n/32][32] elements, 128-byte rows, 128B swizzle) for the blocked kernels ----
using EncodeTiledFn = CUresult (*)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                   const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}
bool make_tile_map(CUtensorMap *tm, const void *ptr, int64_t n, int tile, bool is_int) {
    EncodeTiledFn fn = encode_tiled_fn();
    if (fn == nullptr || n < tile) return false;
    const cuuint64_t dims[2] = {32, static_cast<cuuint64_t>(n / 32)};
    const cuuint64_t strides[1] = {128};
    const cuuint32_t box[2] = {32, static_cast<cuuint32_t>(tile / 32)};
    const cuuint32_t estr[2] = {1, 1};
    return fn(tm, is_int ? CU_TENSOR_MAP_DATA_TYPE_INT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
              const_cast<void *>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// K2 grid: one warp per tile, at most 4 CTAs of 8 warps per SM (grid-stride beyond that)
inline unsigned fix_grid(uint32_t nt) {
    unsigned blocks = (nt + 7u) / 8u;
    const unsigned cap = static_cast<unsigned>(sm_count()) * 4u;
    if (blocks > cap) blocks = cap;
    return blocks < 1u ? 1u : blocks;
}

// 16-byte phase of a 4-byte aligned pointer, in elements (0..3); -1 when the pointer is not even 4-byte aligned
inline int phase_of(const void *p) {
    const uintptr_t a = reinterpret_cast<uintptr_t>(p);
    return (a & 3u) ? -1 : static_cast<int>((a & 15u) >> 2);
}
template <typename T>
inline const T *back(const T *p, int lead) { return p - lead; }
template <typename T>
inline T *back(T *p, int lead) { return p - lead; }

// Persistent kernels contain a grid barrier: they are launched COOPERATIVELY, so the runtime guarantees that
// all CTAs are resident at once (or refuses the launch, in which case the caller falls back to the LDG path).
template <typename... KArgs, typename... Args>
cudaError_t launch_cooperative(void (*kern)(KArgs...), unsigned grid, unsigned threads, size_t smem, cudaStream_t s,
                               Args &&...args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(threads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;
    attr[0].val.cooperative = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
}

// ------------------------------ forward ------------------------------------
template <int OP>
int launch_fwd_fix(float *y, int64_t n, uint32_t nt, int tile, Ws ws, cudaStream_t s) {
    // always launched behind an LDG K1 (even when nothing is unresolved): it also closes the op
    // (resets the workspace counters, advances the epoch)
    k_fwd_fix<OP><<<fix_grid(nt), 256, 0, s>>>(y, n, tile, ws.hdr, ws.desc, ws.ulist(nt));
    ++t_launches;
    return static_cast<int>(cudaGetLastError());
}

// The fallback: one tile per CTA, plain (vector where aligned) loads, fix-up as a second launch.  Any
// alignment, any n, no cross-CTA wait.
template <int OP, int WARPS, int ROWS>
int launch_fwd_ldg(const float *x, const int32_t *key, float *y, int64_t n, Ws ws, cudaStream_t s) {
    constexpr int TILE = WARPS * ROWS * 128;
    const uint32_t nt = tiles_for(n, TILE);
    const int in_vec = aligned16(x) && aligned16(key);
    const int y_vec = aligned16(y);
    k_fwd_ldg<OP, WARPS, ROWS><<<nt, WARPS * 32, 0, s>>>(x, key, y, n, nt, ws.hdr, ws.desc, ws.ulist(nt), in_vec,
                                                          y_vec, g_option[0]);
    ++t_launches;
    const int rc = static_cast<int>(cudaGetLastError());
    return rc != 0 ? rc : launch_fwd_fix<OP>(y, n, nt, TILE, ws, s);
}

// The default: persistent blocked kernel, ONE cooperative launch (streaming phase, grid barrier, sparse fix-up).
// Returns -100 when this path cannot serve the call (the caller then uses the LDG path).
constexpr int NOT_SERVED = -100;
template <int OP, int WARPS, int STAGES>
int launch_fwd_blk(const float *x, const int32_t *key, float *y, int64_t n, Ws ws, cudaStream_t s) {
    using L = FwdBlkSmem<WARPS, STAGES>;
    const int lead = phase_of(x);
    if (lead < 0 || phase_of(key) != lead || (reinterpret_cast<uintptr_t>(y) & 3u)) return NOT_SERVED;
    // alignment peel: every pointer moved down to its 16-byte boundary, `lead` phantom elements in front
    x = back(x, lead); key = back(key, lead); y = back(y, lead);
    n += lead;
    CUtensorMap tmx, tmk;
    if (!make_tile_map(&tmx, x, n, L::TILE, false) || !make_tile_map(&tmk, key, n, L::TILE, true))
        return NOT_SERVED;  // n smaller than a tile, or no driver entry point
    const uint32_t nt = tiles_for(n, L::TILE);
    constexpr auto kern = k_fwd_blk<OP, WARPS, STAGES>;
    const int threads = (WARPS + 1) * 32;
    const int per_sm = persistent_ctas_per_sm<kern>(threads, L::BYTES);
    uint32_t grid = static_cast<uint32_t>(per_sm) * static_cast<uint32_t>(sm_count());
    if (grid > nt) grid = nt;
    const cudaError_t e = launch_cooperative(
        kern, grid, threads, L::BYTES, s, tmx, tmk, x, key, y, n, nt, ws.hdr, ws.desc, ws.ulist(nt),
        aligned16(y) ? 1 : 0, (g_option[0] & 1) | (g_option[2] == 1 ? 2 : 0) | (g_option[2] == 2 ? 4 : 0), lead);
    if (e == cudaErrorCooperativeLaunchTooLarge || e == cudaErrorNotSupported || e == cudaErrorLaunchOutOfResources) {
        (void)cudaGetLastError();   // co-residency cannot be guaranteed here (MPS limit, ...): not an error, use K1 + K2
        return NOT_SERVED;
    }
    ++t_launches;
    return static_cast<int>(e);
}

constexpr int FWD_NUM_VARIANTS = 2;
const char *const kFwdNames[FWD_NUM_VARIANTS] = {
    "ldg_w4_r4 (tile 2048, 1 tile/CTA, 128 threads, separate fix-up launch; any alignment)",
    "blk_w8_s2 (persistent, cooperative; 16 contiguous elements/lane from a swizzled tensor-map TMA ring)"};
constexpr int FWD_DEFAULT = 1;

template <int OP>
int dispatch_fwd(const float *x, const int32_t *key, float *y, int64_t n, void *wsp, size_t ws_bytes,
                 gcp_stream_t stream) {
    t_launches = 0;
    if (n < 0 || n > GCP_MAX_ELEMENTS || (n > 0 && (x == nullptr || key == nullptr || y == nullptr)))
        return GCP_ERR_INVALID_ARG;
    if (n == 0) return GCP_OK;
    Ws ws;
    int rc = check_ws(wsp, ws_bytes, n, &ws);
    if (rc != GCP_OK) return rc;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    const int v = g_variant[0] < 0 ? FWD_DEFAULT : g_variant[0];
    if (v == 1) {
        rc = launch_fwd_blk<OP, 8, 2>(x, key, y, n, ws, s);
        if (rc != NOT_SERVED) return rc;
    } else if (v != 0) {
        return GCP_ERR_INVALID_ARG;
    }
    return launch_fwd_ldg<OP, 4, 4>(x, key, y, n, ws, s);
}

// ------------------------------ backward -----------------------------------
inline int launch_bwd_fix(const float *x, const float *y, const float *g, const int32_t *inv, float *gin, int64_t n,
                          uint32_t nt, int tile, Ws ws, cudaStream_t s) {
    k_bwd_fix<<<fix_grid(nt), 256, 0, s>>>(x, y, g, inv, gin, n, nt, tile, ws.hdr, ws.desc, ws.ulist(nt));
    ++t_launches;
    return static_cast<int>(cudaGetLastError());
}

template <int WARPS, int ROWS>
int launch_bwd_ldg(const float *x, const float *y, const float *g, const int32_t *inv, float *gin, int64_t n,
                   Ws ws, cudaStream_t s) {
    constexpr int TILE = WARPS * ROWS * 128;
    const uint32_t nt = tiles_for(n, TILE);
    const int in_vec = aligned16(x) && aligned16(g) && aligned16(inv);
    k_bwd_ldg<WARPS, ROWS><<<nt, WARPS * 32, 0, s>>>(x, y, g, inv, gin, n, nt, ws.hdr, ws.desc, ws.ulist(nt), in_vec,
                                                      aligned16(gin) ? 1 : 0, g_option[0]);
    ++t_launches;
    const int rc = static_cast<int>(cudaGetLastError());
    return rc != 0 ? rc : launch_bwd_fix(x, y, g, inv, gin, n, nt, TILE, ws, s);
}

template <int WARPS, int STAGES, int MINB>
int launch_bwd_blk(const float *x, const float *y, const float *g, const int32_t *inv, float *gin, int64_t n,
                   Ws ws, cudaStream_t s) {
    using L = BwdBlkSmem<WARPS, STAGES>;
    const int lead = phase_of(x);
    if (lead < 0 || phase_of(g) != lead || phase_of(inv) != lead) return NOT_SERVED;
    if ((reinterpret_cast<uintptr_t>(gin) & 3u) || (reinterpret_cast<uintptr_t>(y) & 3u)) return NOT_SERVED;
    // alignment peel (y is only read with scalar loads: it is shifted for the indexing alone)
    x = back(x, lead); g = back(g, lead); inv = back(inv, lead); y = back(y, lead); gin = back(gin, lead);
    n += lead;
    CUtensorMap tmx, tmg, tmi;
    if (!make_tile_map(&tmx, x, n, L::TILE, false) || !make_tile_map(&tmg, g, n, L::TILE, false) ||
        !make_tile_map(&tmi, inv, n, L::TILE, true))
        return NOT_SERVED;
    const uint32_t nt = tiles_for(n, L::TILE);
    constexpr auto kern = k_bwd_blk<WARPS, STAGES, MINB>;
    const int threads = (WARPS + 1) * 32;
    const int per_sm = persistent_ctas_per_sm<kern>(threads, L::BYTES);
    uint32_t grid = static_cast<uint32_t>(per_sm) * static_cast<uint32_t>(sm_count());
    if (grid > nt) grid = nt;
    const cudaError_t e = launch_cooperative(
        kern, grid, threads, L::BYTES, s, tmx, tmg, tmi, x, y, g, inv, gin, n, nt, ws.hdr, ws.desc, ws.ulist(nt),
        aligned16(gin) ? 1 : 0, (g_option[0] & 1) | (g_option[1] == 1 ? 2 : 0) | (g_option[1] == 2 ? 4 : 0), lead);
    if (e == cudaErrorCooperativeLaunchTooLarge || e == cudaErrorNotSupported || e == cudaErrorLaunchOutOfResources) {
        (void)cudaGetLastError();
        return NOT_SERVED;
    }
    ++t_launches;
    return static_cast<int>(e);
}

constexpr int BWD_NUM_VARIANTS = 2;
const char *const kBwdNames[BWD_NUM_VARIANTS] = {
    "ldg_w4_r4 (tile 2048, 1 tile/CTA, 128 threads, separate fix-up launch; any alignment)",
    "blk_w8_s2 (persistent, cooperative, 2 CTA/SM; 16 contiguous elements/lane from a swizzled tensor-map TMA ring)"};
constexpr int BWD_DEFAULT = 1;

__global__ void k_selftest_abort(uint32_t *hdr) { gcp::signal_abort(hdr); }

}  // namespace

extern "C" {

int gcp_abi_version(void) { return GCP_ABI_VERSION; }

size_t gcp_workspace_bytes(int64_t n) {
    if (n < 0) n = 0;
    const int64_t slots = (n + gcp::MIN_TILE - 1) / gcp::MIN_TILE + 1;
    return static_cast<size_t>(gcp::WS_HEADER_BYTES) +
           static_cast<size_t>(slots) * (gcp::WS_SLOT_BYTES + gcp::WS_LIST_BYTES);
}

int gcp_workspace_init(void *ws, size_t ws_bytes, gcp_stream_t stream) {
    if (ws == nullptr || ws_bytes < static_cast<size_t>(gcp::WS_HEADER_BYTES)) return GCP_ERR_WORKSPACE;
    return static_cast<int>(cudaMemsetAsync(ws, 0, ws_bytes, reinterpret_cast<cudaStream_t>(stream)));
}

int gcp_workspace_attach_flag(void *ws, size_t ws_bytes, void *host_flag_device_address, gcp_stream_t stream) {
    if (ws == nullptr || ws_bytes < static_cast<size_t>(gcp::WS_HEADER_BYTES) ||
        (reinterpret_cast<uintptr_t>(ws) & 15u) != 0 || (reinterpret_cast<uintptr_t>(host_flag_device_address) & 3u))
        return GCP_ERR_WORKSPACE;
    const uint64_t v = reinterpret_cast<uint64_t>(host_flag_device_address);
    return static_cast<int>(cudaMemcpyAsync(reinterpret_cast<uint64_t *>(ws) + gcp::HDR_HOSTFLAG64, &v, sizeof(v),
                                            cudaMemcpyHostToDevice, reinterpret_cast<cudaStream_t>(stream)));
}

int gcp_workspace_selftest_abort(void *ws, size_t ws_bytes, gcp_stream_t stream) {
    if (ws == nullptr || ws_bytes < static_cast<size_t>(gcp::WS_HEADER_BYTES)) return GCP_ERR_WORKSPACE;
    k_selftest_abort<<<1, 1, 0, reinterpret_cast<cudaStream_t>(stream)>>>(reinterpret_cast<uint32_t *>(ws));
    return static_cast<int>(cudaGetLastError());
}

int gcp_workspace_status(const void *ws, gcp_stream_t stream, int *status) {
    if (ws == nullptr || status == nullptr) return GCP_ERR_INVALID_ARG;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    cudaError_t e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) return static_cast<int>(e);
    uint32_t flag = 0;
    e = cudaMemcpy(&flag, reinterpret_cast<const uint32_t *>(ws) + gcp::HDR_ABORT, sizeof(flag),
                   cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) return static_cast<int>(e);
    *status = flag ? GCP_ERR_WATCHDOG : GCP_OK;
    return GCP_OK;
}

int gcp_cumprod_fwd_f32(const float *x, const int32_t *key, float *y, int64_t n, void *ws, size_t ws_bytes,
                        gcp_stream_t stream) {
    return dispatch_fwd<gcp::OP_MUL>(x, key, y, n, ws, ws_bytes, stream);
}

int gcp_cumsum_fwd_f32(const float *x, const int32_t *key, float *y, int64_t n, void *ws, size_t ws_bytes,
                       gcp_stream_t stream) {
    return dispatch_fwd<gcp::OP_ADD>(x, key, y, n, ws, ws_bytes, stream);
}

int gcp_cumprod_bwd_f32(const float *x, const float *y, const float *gout, const int32_t *inv,
                        const int32_t *seg_end, float *gin, int64_t n, int64_t k, void *wsp, size_t ws_bytes,
                        gcp_stream_t stream) {
    t_launches = 0;
    (void)seg_end;  // implied by inv (tail <=> inv[i+1] != inv[i]); see gcp_validate_segments
    if (n < 0 || n > GCP_MAX_ELEMENTS || k < 0) return GCP_ERR_INVALID_ARG;
    if (n > 0 && (x == nullptr || y == nullptr || gout == nullptr || inv == nullptr || gin == nullptr))
        return GCP_ERR_INVALID_ARG;
    if (n == 0) return GCP_OK;
    Ws ws;
    int rc = check_ws(wsp, ws_bytes, n, &ws);
    if (rc != GCP_OK) return rc;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    const int v = g_variant[1] < 0 ? BWD_DEFAULT : g_variant[1];
    if (v == 1) {
        rc = launch_bwd_blk<8, 2, 2>(x, y, gout, inv, gin, n, ws, s);
        if (rc != NOT_SERVED) return rc;
    } else if (v != 0) {
        return GCP_ERR_INVALID_ARG;
    }
    return launch_bwd_ldg<4, 4>(x, y, gout, inv, gin, n, ws, s);
}

int gcp_validate_segments(const int32_t *inv, const int32_t *seg_end, int64_t n, int64_t k, void *wsp,
                          size_t ws_bytes, gcp_stream_t stream, int64_t *violations) {
    t_launches = 0;
    if (violations == nullptr || n < 0 || k < 0) return GCP_ERR_INVALID_ARG;
    if (n > 0 && (inv == nullptr || (k > 0 && seg_end == nullptr))) return GCP_ERR_INVALID_ARG;
    Ws ws;
    int rc = check_ws(wsp, ws_bytes, 0, &ws);
    if (rc != GCP_OK) return rc;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    unsigned long long *ctr = reinterpret_cast<unsigned long long *>(ws.hdr) + gcp::HDR_VIOL64;
    cudaError_t e = cudaMemsetAsync(ctr, 0, sizeof(unsigned long long), s);
    if (e != cudaSuccess) return static_cast<int>(e);
    if (n == 0 && k == 0) {
        *violations = 0;
        return GCP_OK;
    }
    const int threads = 256;
    int64_t blocks = (n + threads - 1) / threads;
    const int64_t cap = static_cast<int64_t>(sm_count()) * 16;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    gcp::k_validate_segments<<<static_cast<unsigned>(blocks), threads, 0, s>>>(inv, seg_end, n, k, ctr);
    ++t_launches;
    e = cudaGetLastError();
    if (e != cudaSuccess) return static_cast<int>(e);
    unsigned long long host = 0;
    e = cudaMemcpyAsync(&host, ctr, sizeof(host), cudaMemcpyDeviceToHost, s);
    if (e != cudaSuccess) return static_cast<int>(e);
    e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) return static_cast<int>(e);
    *violations = static_cast<int64_t>(host);
    return GCP_OK;
}

int gcp_set_variant(int op, int variant) {
    if (op < 0 || op > 1) return GCP_ERR_INVALID_ARG;
    const int nv = op == 0 ? FWD_NUM_VARIANTS : BWD_NUM_VARIANTS;
    if (variant < -1 || variant >= nv) return GCP_ERR_INVALID_ARG;
    g_variant[op] = variant;
    return GCP_OK;
}

int gcp_set_option(int option, int value) {
    if (option < 0 || option >= 4) return GCP_ERR_INVALID_ARG;
    g_option[option] = value;
    return GCP_OK;
}

int gcp_num_variants(int op) { return op == 0 ? FWD_NUM_VARIANTS : (op == 1 ? BWD_NUM_VARIANTS : 0); }

const char *gcp_variant_name(int op, int variant) {
    if (op == 0 && variant >= 0 && variant < FWD_NUM_VARIANTS) return kFwdNames[variant];
    if (op == 1 && variant >= 0 && variant < BWD_NUM_VARIANTS) return kBwdNames[variant];
    return "";
}

int gcp_last_launch_count(void) { return t_launches; }

}  // extern "C"
