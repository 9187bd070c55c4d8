// gcp_host.cu — helpers of the host-buffer entry point (simplegaussiansplat_tk71_b200/host.py): when the element
// arrays live in HOST memory the PCIe link bounds the step, so the segment keys do not cross it as 4 bytes per
// element.  A segment is a run of equal adjacent keys (grouped_cumprod_forward.cu:17-23), so all the scan ops need
// of `key` is where the runs start: the host packs that into one BIT per element (OpenMP, memory-bound), 1/32 of
// the bytes go up, and the device rebuilds dense segment ids with one scan.
#include <cuda_runtime.h>
#include <stdint.h>
#if defined(__x86_64__)
#include <immintrin.h>
#endif

#include <cub/device/device_scan.cuh>
#include <thrust/iterator/counting_iterator.h>
#include <thrust/iterator/transform_iterator.h>

#include "gcp_abi.h"

namespace {
struct BitAt {   // 1 where element i starts a run; element 0 contributes 0 so that the ids start at 0
    const uint32_t *bits;
    __host__ __device__ __forceinline__ int32_t operator()(int64_t i) const {
        return i == 0 ? 0 : static_cast<int32_t>((bits[i >> 5] >> (i & 31)) & 1u);
    }
};
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
    size_t a = 0;
    auto it = thrust::make_transform_iterator(thrust::counting_iterator<int64_t>(0), BitAt{nullptr});
    cub::DeviceScan::InclusiveSum(nullptr, a, it, static_cast<int32_t *>(nullptr), n > 0 ? n : 1);
    return a + 256;
}

int gcp_ids_from_bits(const uint32_t *bits, int64_t n, int32_t *ids, void *temp, size_t temp_bytes,
                      gcp_stream_t stream) {
    if (n < 0) return GCP_ERR_INVALID_ARG;
    if (n == 0) return GCP_OK;
    if (!bits || !ids || !temp) return GCP_ERR_INVALID_ARG;
    auto it = thrust::make_transform_iterator(thrust::counting_iterator<int64_t>(0), BitAt{bits});
    size_t need = 0;
    cub::DeviceScan::InclusiveSum(nullptr, need, it, ids, n);
    if (temp_bytes < need) return GCP_ERR_WORKSPACE;
    size_t tb = temp_bytes;
    return static_cast<int>(cub::DeviceScan::InclusiveSum(temp, tb, it, ids, n, reinterpret_cast<cudaStream_t>(stream)));
}

}  // extern "C"
