// gcp_host.cu — helpers of the host-buffer entry point (simplegaussiansplat_tk71_b200/host.py): when the element
// arrays live in HOST memory the PCIe link bounds the step, so the segment keys do not cross it as 4 bytes per
// element.  A segment is a run of equal adjacent keys (grouped_cumprod_forward.cu:17-23), so all the scan ops need
// of `key` is where the runs start: the host packs that into one BIT per element (OpenMP, memory-bound), 1/32 of
// the bytes go up, and the device rebuilds dense segment ids with one scan (k_ids_from_bits: a hand-written
// single-pass chained scan over the bit words, no library kernel).
#include <cuda_runtime.h>
#include <stdint.h>
#if defined(__x86_64__)
#include <immintrin.h>
#endif

#include "gcp_abi.h"

namespace {
// ids[i] = number of run starts in elements 1..i (element 0 starts the first run and counts as id 0).  One thread
// per 32-bit word: its popcount is the word's aggregate; the block scans its 256 aggregates, publishes the block
// aggregate and walks back over the predecessors' descriptors to the nearest inclusive prefix (blocks take their
// tile from an atomic ticket, so a block only waits for blocks that are already running); then every warp expands
// its 32 words one after the other, lane = element, 128 contiguous bytes per store.
// descriptor = status (1 aggregate, 2 inclusive) << 62 | value;  temp = {ticket u32, pad} + one descriptor per block.
constexpr int IDS_THREADS = 256;

__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_u64(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__global__ void __launch_bounds__(IDS_THREADS)
k_ids_from_bits(const uint32_t *__restrict__ bits, int64_t n, int32_t *__restrict__ ids, unsigned int *ticket,
                unsigned long long *desc) {
    __shared__ unsigned int s_warp[IDS_THREADS / 32];
    __shared__ unsigned long long s_prefix;
    __shared__ unsigned int s_tile;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
    __syncthreads();
    const unsigned int tile = s_tile;
    const int64_t words = (n + 31) >> 5;
    const int64_t w = static_cast<int64_t>(tile) * IDS_THREADS + threadIdx.x;
    uint32_t word = 0;
    if (w < words) {
        word = bits[w];
        const int64_t left = n - (w << 5);
        if (left < 32) word &= (1u << left) - 1u;     // the ragged tail
        if (w == 0) word &= ~1u;                      // element 0 is id 0
    }
    const unsigned int cnt = __popc(word);
    unsigned int inc = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned int t = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += t;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    unsigned int wpre = 0, agg = 0;
#pragma unroll
    for (int k = 0; k < IDS_THREADS / 32; ++k) {
        if (k < warp) wpre += s_warp[k];
        agg += s_warp[k];
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
    const unsigned int before = static_cast<unsigned int>(s_prefix) + wpre + (inc - cnt);   // run starts before my word
    const int64_t w0 = w - lane;                                                             // the warp's first word
#pragma unroll 4
    for (int k = 0; k < 32; ++k) {
        const uint32_t wk = __shfl_sync(0xffffffffu, word, k);
        const unsigned int bk = __shfl_sync(0xffffffffu, before, k);
        const int64_t i = ((w0 + k) << 5) + lane;
        if (i < n) ids[i] = static_cast<int32_t>(bk + __popc(wk & ((2u << lane) - 1u)));
    }
}

// run-start bits of the (up to 32) elements of word w
inline uint32_t word_scalar(const int32_t *key, int64_t n, int64_t w) {
    const int64_t i0 = w << 5;
    const int m = static_cast<int>(n - i0 < 32 ? n - i0 : 32);
    uint32_t v = 0;
    int32_t prev = i0 > 0 ? key[i0 - 1] : ~key[0];
    for (int b = 0; b < m; ++b) {
        const int32_t k = key[i0 + b];
        v |= static_cast<uint32_t>(k != prev) << b;
        prev = k;
    }
    return v;
}

#if defined(__x86_64__)
// a full interior word: four 8-wide compares of key[i..] against key[i-1..]
__attribute__((target("avx2"))) inline uint32_t word_avx2(const int32_t *key, int64_t w) {
    const int32_t *p = key + (w << 5);
    uint32_t v = 0;
    for (int g = 0; g < 4; ++g) {
        const __m256i cur = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(p + 8 * g));
        const __m256i prv = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(p + 8 * g - 1));
        const unsigned eq = static_cast<unsigned>(_mm256_movemask_ps(_mm256_castsi256_ps(_mm256_cmpeq_epi32(cur, prv))));
        v |= ((~eq) & 0xffu) << (8 * g);
    }
    return v;
}
__attribute__((target("avx2"))) void words_avx2(const int32_t *key, int64_t w0, int64_t w1, uint32_t *bits) {
    for (int64_t w = w0; w < w1; ++w) bits[w] = word_avx2(key, w);
}
#endif
}  // namespace

extern "C" {

int gcp_host_boundary_bits(const int32_t *key, int64_t n, uint32_t *bits, int threads) {
    if (n < 0 || (n > 0 && (!key || !bits))) return GCP_ERR_INVALID_ARG;
    const int64_t words = (n + 31) >> 5;
    if (words == 0) return GCP_OK;
    bits[0] = word_scalar(key, n, 0);                       // element 0 always starts a run
    if (words > 1) bits[words - 1] = word_scalar(key, n, words - 1);   // the ragged tail
    const int64_t w0 = 1, w1 = words - 1;                   // full interior words
#if defined(__x86_64__)
    const bool avx2 = __builtin_cpu_supports("avx2");
#else
    const bool avx2 = false;
#endif
    constexpr int64_t BLK = 4096;                           // words per task (512 KB of keys)
    const int64_t nblk = (w1 - w0 + BLK - 1) / BLK;
#pragma omp parallel for schedule(static) num_threads(threads > 0 ? threads : 1)
    for (int64_t b = 0; b < nblk; ++b) {
        const int64_t a = w0 + b * BLK, e = (a + BLK < w1) ? a + BLK : w1;
#if defined(__x86_64__)
        if (avx2) {
            words_avx2(key, a, e, bits);
            continue;
        }
#endif
        for (int64_t w = a; w < e; ++w) bits[w] = word_scalar(key, n, w);
    }
    return GCP_OK;
}

size_t gcp_ids_from_bits_bytes(int64_t n) {
    const int64_t words = n > 0 ? (n + 31) >> 5 : 1;
    const int64_t blocks = (words + IDS_THREADS - 1) / IDS_THREADS;
    return 256 + static_cast<size_t>(blocks) * 8;
}

int gcp_ids_from_bits(const uint32_t *bits, int64_t n, int32_t *ids, void *temp, size_t temp_bytes,
                      gcp_stream_t stream) {
    if (n < 0) return GCP_ERR_INVALID_ARG;
    if (n == 0) return GCP_OK;
    if (!bits || !ids || !temp || (reinterpret_cast<uintptr_t>(temp) & 7)) return GCP_ERR_INVALID_ARG;
    const size_t need = gcp_ids_from_bits_bytes(n);
    if (temp_bytes < need) return GCP_ERR_WORKSPACE;
    auto st = reinterpret_cast<cudaStream_t>(stream);
    cudaError_t e = cudaMemsetAsync(temp, 0, need, st);     // ticket + descriptors
    if (e != cudaSuccess) return static_cast<int>(e);
    const int64_t words = (n + 31) >> 5;
    const unsigned blocks = static_cast<unsigned>((words + IDS_THREADS - 1) / IDS_THREADS);
    k_ids_from_bits<<<blocks, IDS_THREADS, 0, st>>>(bits, n, ids, static_cast<unsigned int *>(temp),
                                                    reinterpret_cast<unsigned long long *>(static_cast<unsigned char *>(temp) + 256));
    return static_cast<int>(cudaGetLastError());
}

}  // extern "C"
