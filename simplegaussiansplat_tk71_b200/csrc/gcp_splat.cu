// gcp_splat.cu — kernels around the scan ops for the compositor rows:
// box -> element expansion straight into (key, gaussian-id) pairs, the stable pixel-key sort,
// per-element alpha in sorted order, per-pixel colour reduction, and the per-element /
// per-Gaussian backward.  Everything here is a plain streaming kernel (no cross-CTA carries): the
// two scans of a step are the ops of gcp_abi.cu (gcp_cumprod_fwd_f32 / gcp_cumprod_bwd_f32).
//
// Reference being replaced: the torch op chains of gs_model.py:480-514 (expansion, Gaussian kernel,
// pixel accumulation with index_put_ atomics), :538-548 (key build + sort + gather), :627-663 and
// :733-783 (per-element gradients and scatter_reduce to Gaussians).
//
// Element = (Gaussian j, pixel (x,y)) with (x,y) inside the inclusive box [sp_j, ep_j]; element order
// before sorting = Gaussian-major, row-major inside the box (uitility.py:336-366); pixel key =
// y*10000 + x (gs_model.py:541).  The sort is a stable LSD radix sort on the key bits only, so inside a
// pixel the depth order (= Gaussian index order) is preserved bit-exactly.
#include <cuda_runtime.h>
#include <stdint.h>
#include <algorithm>

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <thrust/iterator/counting_iterator.h>
#include <thrust/iterator/transform_iterator.h>

#include "gcp_abi.h"

namespace {

constexpr int KEY_STRIDE = 10000;
constexpr int CH = 8;  // consecutive elements per thread

// first index g with goff[g+1] > e  (goff: exclusive offsets, n+1 entries)
__device__ __forceinline__ int64_t find_gaussian(const int64_t *__restrict__ goff, int64_t n, int64_t e) {
    int64_t lo = 0, hi = n - 1;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (__ldg(goff + mid + 1) > e) hi = mid;
        else lo = mid + 1;
    }
    return lo;
}

__global__ void __launch_bounds__(256)
k_splat_expand(const int32_t *__restrict__ sp, const int32_t *__restrict__ ep, const int64_t *__restrict__ goff,
               int64_t n, int64_t N, int32_t *__restrict__ key, int32_t *__restrict__ gid) {
    const int64_t e0 = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) * CH;
    if (e0 >= N) return;
    int64_t g = find_gaussian(goff, n, e0);
    int64_t gbeg = __ldg(goff + g), gend = __ldg(goff + g + 1);
    int sx = __ldg(sp + 2 * g), sy = __ldg(sp + 2 * g + 1);
    int w = __ldg(ep + 2 * g) - sx + 1;
    int32_t ko[CH], go[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) {
        const int64_t e = e0 + i;
        if (e < N) {
            while (e >= gend) {  // skips empty boxes too
                ++g;
                gbeg = gend;
                gend = __ldg(goff + g + 1);
                sx = __ldg(sp + 2 * g);
                sy = __ldg(sp + 2 * g + 1);
                w = __ldg(ep + 2 * g) - sx + 1;
            }
            const int local = static_cast<int>(e - gbeg);
            const int iy = local / w;
            const int ix = local - iy * w;
            ko[i] = (sy + iy) * KEY_STRIDE + sx + ix;
            go[i] = static_cast<int32_t>(g);
        } else {
            ko[i] = 0;
            go[i] = 0;
        }
    }
    if (e0 + CH <= N) {
        reinterpret_cast<int4 *>(key + e0)[0] = make_int4(ko[0], ko[1], ko[2], ko[3]);
        reinterpret_cast<int4 *>(key + e0)[1] = make_int4(ko[4], ko[5], ko[6], ko[7]);
        reinterpret_cast<int4 *>(gid + e0)[0] = make_int4(go[0], go[1], go[2], go[3]);
        reinterpret_cast<int4 *>(gid + e0)[1] = make_int4(go[4], go[5], go[6], go[7]);
    } else {
        for (int i = 0; i < CH && e0 + i < N; ++i) {
            key[e0 + i] = ko[i];
            gid[e0 + i] = go[i];
        }
    }
}

struct Gauss {
    float mx, my, l00, l01, l10, l11, o;
};
__device__ __forceinline__ Gauss load_gauss(const float *__restrict__ mean, const float *__restrict__ lam,
                                            const float *__restrict__ opac, int g) {
    const float2 m = __ldg(reinterpret_cast<const float2 *>(mean) + g);
    const float4 L = __ldg(reinterpret_cast<const float4 *>(lam) + g);
    return Gauss{m.x, m.y, L.x, L.y, L.z, L.w, __ldg(opac + g)};
}
// g = exp(-1/2 (r-m) Lambda (r-m)^T)  with X = (r-m) Lambda  (gs_model.py:495, :745)
__device__ __forceinline__ float gauss_kernel(const Gauss &G, int key, float &d0, float &d1, float &X0, float &X1) {
    const int py = key / KEY_STRIDE;
    const int px = key - py * KEY_STRIDE;
    d0 = static_cast<float>(px) - G.mx;
    d1 = static_cast<float>(py) - G.my;
    X0 = d0 * G.l00 + d1 * G.l10;
    X1 = d0 * G.l01 + d1 * G.l11;
    return expf(-0.5f * (X0 * d0 + X1 * d1));
}

// Per-Gaussian tables packed into 32-byte records, one L2 sector per gather in the per-element kernels
// (consecutive elements of a pixel list belong to unrelated Gaussians: every gather is its own sector):
//   rec_a[g] = {mx, my, l00, l01 | l10, l11, o, 0}                 -> k_splat_alpha
//   rec_b[g] = {l0, l1, l2, bits(sx) | sy, w, goff_lo, goff_hi}    -> colour / backward kernels (first half only
//                                                                     where the box is not needed)
__global__ void __launch_bounds__(256)
k_splat_pack(const float *__restrict__ mean, const float *__restrict__ lam, const float *__restrict__ opac,
             const float *__restrict__ l_d, const int32_t *__restrict__ sp, const int32_t *__restrict__ ep,
             const int64_t *__restrict__ goff, int64_t n, float4 *__restrict__ rec_a, int4 *__restrict__ rec_b) {
    const int64_t g = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (g >= n) return;
    const Gauss G = load_gauss(mean, lam, opac, static_cast<int>(g));
    rec_a[2 * g] = make_float4(G.mx, G.my, G.l00, G.l01);
    rec_a[2 * g + 1] = make_float4(G.l10, G.l11, G.o, 0.f);
    const int sx = __ldg(sp + 2 * g), sy = __ldg(sp + 2 * g + 1);
    const int64_t off = __ldg(goff + g);
    rec_b[2 * g] = make_int4(__float_as_int(__ldg(l_d + 3 * g)), __float_as_int(__ldg(l_d + 3 * g + 1)),
                             __float_as_int(__ldg(l_d + 3 * g + 2)), sx);
    rec_b[2 * g + 1] = make_int4(sy, __ldg(ep + 2 * g) - sx + 1, static_cast<int>(off & 0xffffffffll),
                                 static_cast<int>(off >> 32));
}

// one 32-byte record in one instruction (LDG.256, sm_100)
__device__ __forceinline__ void ldg256(const void *p, int4 &u, int4 &v) {
    // volatile + memory clobber: keeps the compiler from sinking a gather below later stores / into branches, so
    // that several gathers issued back to back stay in flight together
    asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w), "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "l"(p)
                 : "memory");
}
__device__ __forceinline__ Gauss load_rec_a(const float4 *__restrict__ rec_a, int g) {
    int4 u, v;
    ldg256(rec_a + 2 * static_cast<int64_t>(g), u, v);
    return Gauss{__int_as_float(u.x), __int_as_float(u.y), __int_as_float(u.z), __int_as_float(u.w),
                 __int_as_float(v.x), __int_as_float(v.y), __int_as_float(v.z)};
}

// x_s[e] = 1 - o*g in sorted order (the scan's input); 4 consecutive elements per thread
__global__ void __launch_bounds__(256, 3)
k_splat_alpha(const int32_t *__restrict__ key_s, const int32_t *__restrict__ gid_s,
              const float4 *__restrict__ rec_a, int64_t N, float *__restrict__ x_s) {
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x * 4;
    for (int64_t e = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) * 4; e < N; e += stride) {
        int k[4], g[4];
        if (e + 4 <= N) {
            const int4 kk = __ldg(reinterpret_cast<const int4 *>(key_s + e));
            const int4 gg = __ldg(reinterpret_cast<const int4 *>(gid_s + e));
            k[0] = kk.x; k[1] = kk.y; k[2] = kk.z; k[3] = kk.w;
            g[0] = gg.x; g[1] = gg.y; g[2] = gg.z; g[3] = gg.w;
        } else {
            for (int i = 0; i < 4; ++i) {
                k[i] = (e + i < N) ? __ldg(key_s + e + i) : 0;
                g[i] = (e + i < N) ? __ldg(gid_s + e + i) : 0;
            }
        }
        Gauss G[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) G[i] = load_rec_a(rec_a, g[i]);
        float out[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float d0, d1, X0, X1;
            out[i] = 1.0f - G[i].o * gauss_kernel(G[i], k[i], d0, d1, X0, X1);
        }
        if (e + 4 <= N) {
            *reinterpret_cast<float4 *>(x_s + e) = make_float4(out[0], out[1], out[2], out[3]);
        } else {
            for (int i = 0; i < 4 && e + i < N; ++i) x_s[e + i] = out[i];
        }
    }
}

__device__ __forceinline__ int pixel_index(int key, int W) {
    const int py = key / KEY_STRIDE;
    return py * (W + 1) + (key - py * KEY_STRIDE);
}

// 8 consecutive elements of arrays a (int) / f (float) starting at e (e % 8 == 0 for every thread: vector loads)
__device__ __forceinline__ void load8(const int32_t *__restrict__ a, int64_t e, int64_t N, int (&v)[8], int fill) {
    if (e + 8 <= N) {
        const int4 p = __ldg(reinterpret_cast<const int4 *>(a + e)), q = __ldg(reinterpret_cast<const int4 *>(a + e) + 1);
        v[0] = p.x; v[1] = p.y; v[2] = p.z; v[3] = p.w; v[4] = q.x; v[5] = q.y; v[6] = q.z; v[7] = q.w;
    } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = (e + i < N) ? __ldg(a + e + i) : fill;
    }
}
__device__ __forceinline__ void load8(const float *__restrict__ a, int64_t e, int64_t N, float (&v)[8], float fill) {
    if (e + 8 <= N) {
        const float4 p = __ldg(reinterpret_cast<const float4 *>(a + e)),
                     q = __ldg(reinterpret_cast<const float4 *>(a + e) + 1);
        v[0] = p.x; v[1] = p.y; v[2] = p.z; v[3] = p.w; v[4] = q.x; v[5] = q.y; v[6] = q.z; v[7] = q.w;
    } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = (e + i < N) ? __ldg(a + e + i) : fill;
    }
}

__device__ __forceinline__ void color_flush(float *__restrict__ image, int key, int W, float a0, float a1, float a2) {
    if (key >= 0 && (a0 != 0.f || a1 != 0.f || a2 != 0.f)) {
        float *p = image + 3 * static_cast<int64_t>(pixel_index(key, W));
        atomicAdd(p, a0); atomicAdd(p + 1, a1); atomicAdd(p + 2, a2);
    }
}

// image[pixel] += sum_i T_i alpha_i l_i  (T exclusive = previous inclusive product, 1 at a head;
// elements whose inclusive product is 0 contribute nothing, gs_model.py:575-578).
// 8 consecutive elements per thread; a thread adds the partial sum of every pixel run it sees (one atomic
// triple per run and thread — a run of a few hundred elements is a few dozen adds).
__global__ void __launch_bounds__(256)
k_splat_color(const float *__restrict__ incl, const float *__restrict__ x_s, const int32_t *__restrict__ key_s,
              const int32_t *__restrict__ gid_s, const int4 *__restrict__ rec_b, int64_t N, int W,
              float *__restrict__ image) {
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x * 8;
    for (int64_t e0 = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) * 8; e0 < N; e0 += stride) {
        int k[8], g[8];
        float y[8], x[8];
        load8(key_s, e0, N, k, -1);
        load8(gid_s, e0, N, g, 0);
        load8(incl, e0, N, y, 0.0f);
        load8(x_s, e0, N, x, 1.0f);
        int kprev = (e0 > 0) ? __ldg(key_s + e0 - 1) : -1;
        float yprev = (e0 > 0) ? __ldg(incl + e0 - 1) : 1.0f;
        int4 l[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) l[i] = (y[i] != 0.0f) ? __ldg(rec_b + 2 * static_cast<int64_t>(g[i])) : make_int4(0, 0, 0, 0);
        float a0 = 0.f, a1 = 0.f, a2 = 0.f;
        int kcur = -1;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float T = (k[i] != kprev) ? 1.0f : yprev;
            if (k[i] != kcur) {
                color_flush(image, kcur, W, a0, a1, a2);
                kcur = k[i];
                a0 = a1 = a2 = 0.f;
            }
            if (y[i] != 0.0f) {
                const float ta = T * (1.0f - x[i]);
                a0 = fmaf(ta, __int_as_float(l[i].x), a0);
                a1 = fmaf(ta, __int_as_float(l[i].y), a1);
                a2 = fmaf(ta, __int_as_float(l[i].z), a2);
            }
            kprev = k[i];
            yprev = y[i];
        }
        color_flush(image, kcur, W, a0, a1, a2);
    }
}

// gshift[k] = w_{k+1} inside a pixel list, 0 at its tail, with w_k = <dL/dI(pixel), alpha_k l_k> (0 for dead
// elements): the grad_out that makes grouped_cumprod_backward return T_k * U_k.  4 elements per thread.
__global__ void __launch_bounds__(256)
k_splat_bwd_w(const float *__restrict__ incl, const float *__restrict__ x_s, const int32_t *__restrict__ key_s,
              const int32_t *__restrict__ gid_s, const int4 *__restrict__ rec_b, const float *__restrict__ gimg,
              int64_t N, int W, float *__restrict__ gshift) {
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x * 4;
    for (int64_t e = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) * 4; e < N; e += stride) {
        // elements e+1 .. e+4 are needed (w of the successor): slots 0..4 hold elements e .. e+4
        int k[5], g[5];
        float y[5], x[5];
        if (e + 4 <= N) {
            const int4 kk = __ldg(reinterpret_cast<const int4 *>(key_s + e));
            const int4 gg = __ldg(reinterpret_cast<const int4 *>(gid_s + e));
            const float4 yy = __ldg(reinterpret_cast<const float4 *>(incl + e));
            const float4 xv = __ldg(reinterpret_cast<const float4 *>(x_s + e));
            k[0] = kk.x; k[1] = kk.y; k[2] = kk.z; k[3] = kk.w;
            g[0] = gg.x; g[1] = gg.y; g[2] = gg.z; g[3] = gg.w;
            y[0] = yy.x; y[1] = yy.y; y[2] = yy.z; y[3] = yy.w;
            x[0] = xv.x; x[1] = xv.y; x[2] = xv.z; x[3] = xv.w;
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const bool in = e + i < N;
                k[i] = in ? __ldg(key_s + e + i) : -1;
                g[i] = in ? __ldg(gid_s + e + i) : 0;
                y[i] = in ? __ldg(incl + e + i) : 0.0f;
                x[i] = in ? __ldg(x_s + e + i) : 1.0f;
            }
        }
        {
            const bool in = e + 4 < N;
            k[4] = in ? __ldg(key_s + e + 4) : -1;
            g[4] = in ? __ldg(gid_s + e + 4) : 0;
            y[4] = in ? __ldg(incl + e + 4) : 0.0f;
            x[4] = in ? __ldg(x_s + e + 4) : 1.0f;
        }
        float out[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            out[i] = 0.0f;
            if (k[i + 1] == k[i] && k[i] >= 0 && y[i + 1] != 0.0f) {
                const int4 l = __ldg(rec_b + 2 * static_cast<int64_t>(g[i + 1]));
                const float *pg = gimg + 3 * static_cast<int64_t>(pixel_index(k[i + 1], W));
                const float pgl = __ldg(pg) * __int_as_float(l.x) + __ldg(pg + 1) * __int_as_float(l.y) +
                                  __ldg(pg + 2) * __int_as_float(l.z);
                out[i] = (1.0f - x[i + 1]) * pgl;
            }
        }
        if (e + 4 <= N) {
            *reinterpret_cast<float4 *>(gshift + e) = make_float4(out[0], out[1], out[2], out[3]);
        } else {
            for (int i = 0; i < 4 && e + i < N; ++i) gshift[e + i] = out[i];
        }
    }
}

// per element: dalpha = T <dL/dI, l> - T U  (tu = T*U from grouped_cumprod_backward), then the reference's
// per-element gradients (gs_model.py:733-766) accumulated per Gaussian (:776-783):
//   d_opacity += g * dalpha          d_l[c] += d / l[c]   (the reference's d/l, d = T w)
//   d_mean    += alpha dalpha (r-m)Lambda        d_Lambda += -1/2 alpha dalpha (r-m)^T (r-m)
__global__ void __launch_bounds__(256)
k_splat_bwd_grads(const float *__restrict__ incl, const float *__restrict__ x_s, const float *__restrict__ tu,
                  const int32_t *__restrict__ key_s, const int32_t *__restrict__ gid_s,
                  const float *__restrict__ mean, const float *__restrict__ lam, const float *__restrict__ opac,
                  const float *__restrict__ l_d, const float *__restrict__ gimg, int64_t N, int W,
                  float *__restrict__ g_mean, float *__restrict__ g_lam, float *__restrict__ g_opac,
                  float *__restrict__ g_l) {
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t e = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; e < N; e += stride) {
        const float y = __ldg(incl + e);
        if (y == 0.0f) continue;  // dead element: no contribution, no gradient
        const int k = __ldg(key_s + e);
        const bool head = (e == 0) || (__ldg(key_s + e - 1) != k);
        const float T = head ? 1.0f : __ldg(incl + e - 1);
        const int g = __ldg(gid_s + e);
        const Gauss G = load_gauss(mean, lam, opac, g);
        float d0, d1, X0, X1;
        const float gk = gauss_kernel(G, k, d0, d1, X0, X1);
        const float alpha = G.o * gk;
        const float *pg = gimg + 3 * static_cast<int64_t>(pixel_index(k, W));
        const float l0 = __ldg(l_d + 3 * g), l1 = __ldg(l_d + 3 * g + 1), l2 = __ldg(l_d + 3 * g + 2);
        const float pgl = __ldg(pg) * l0 + __ldg(pg + 1) * l1 + __ldg(pg + 2) * l2;
        const float dalpha = T * pgl - __ldg(tu + e);
        const float d = T * alpha * pgl;
        const float coef = alpha * dalpha;
        atomicAdd(g_opac + g, gk * dalpha);
        atomicAdd(g_l + 3 * g, d / l0);
        atomicAdd(g_l + 3 * g + 1, d / l1);
        atomicAdd(g_l + 3 * g + 2, d / l2);
        atomicAdd(g_mean + 2 * g, coef * X0);
        atomicAdd(g_mean + 2 * g + 1, coef * X1);
        const float hc = -0.5f * coef;
        atomicAdd(g_lam + 4 * g, hc * d0 * d0);
        atomicAdd(g_lam + 4 * g + 1, hc * d0 * d1);
        atomicAdd(g_lam + 4 * g + 2, hc * d1 * d0);
        atomicAdd(g_lam + 4 * g + 3, hc * d1 * d1);
    }
}

// Backward, step 1 of 2 (sorted order): per element (dalpha, d) written at the element's GAUSSIAN-MAJOR
// position e = goff[g] + (y - sy)*w + (x - sx), i.e. un-sorted without a permutation array.  Neighbouring
// pixels of a box row are neighbouring pixel lists, so the scattered 8-byte stores of one tile land in
// the same few sectors.
//   dalpha = T <dL/dI, l> - T U   (tu = T*U from grouped_cumprod_backward);   d = T alpha <dL/dI, l>
__global__ void __launch_bounds__(256, 3)
k_splat_bwd_elem(const float *__restrict__ incl, const float *__restrict__ x_s, const float *__restrict__ tu,
                 const int32_t *__restrict__ key_s, const int32_t *__restrict__ gid_s,
                 const int4 *__restrict__ rec_b, const float *__restrict__ gimg, int64_t N, int W,
                 float2 *__restrict__ elem) {
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x * 4;
    for (int64_t e0 = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) * 4; e0 < N; e0 += stride) {
        int k[4], g[4];
        float y[4], xx[4], t[4];
        if (e0 + 4 <= N) {
            // read-once streams: evict-first, so that the dirty lines of `elem` stay in L2 until their box is complete
            const int4 kk = __ldcs(reinterpret_cast<const int4 *>(key_s + e0));
            const int4 gg = __ldcs(reinterpret_cast<const int4 *>(gid_s + e0));
            const float4 yy = __ldcs(reinterpret_cast<const float4 *>(incl + e0));
            const float4 xv = __ldcs(reinterpret_cast<const float4 *>(x_s + e0));
            const float4 tv = __ldcs(reinterpret_cast<const float4 *>(tu + e0));
            k[0] = kk.x; k[1] = kk.y; k[2] = kk.z; k[3] = kk.w;
            g[0] = gg.x; g[1] = gg.y; g[2] = gg.z; g[3] = gg.w;
            y[0] = yy.x; y[1] = yy.y; y[2] = yy.z; y[3] = yy.w;
            xx[0] = xv.x; xx[1] = xv.y; xx[2] = xv.z; xx[3] = xv.w;
            t[0] = tv.x; t[1] = tv.y; t[2] = tv.z; t[3] = tv.w;
        } else {
            for (int i = 0; i < 4; ++i) {
                const bool in = e0 + i < N;
                k[i] = in ? __ldg(key_s + e0 + i) : -1;
                g[i] = in ? __ldg(gid_s + e0 + i) : 0;
                y[i] = in ? __ldg(incl + e0 + i) : 0.0f;
                xx[i] = in ? __ldg(x_s + e0 + i) : 1.0f;
                t[i] = in ? __ldg(tu + e0 + i) : 0.0f;
            }
        }
        int kprev = (e0 > 0) ? __ldg(key_s + e0 - 1) : -1;
        float yprev = (e0 > 0) ? __ldg(incl + e0 - 1) : 1.0f;
        int4 ra[4], rb[4];
        float pg[4][3];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            ldg256(rec_b + 2 * static_cast<int64_t>(g[i]), ra[i], rb[i]);
            const float *q = gimg + 3 * static_cast<int64_t>(pixel_index(max(k[i], 0), W));
            pg[i][0] = __ldg(q); pg[i][1] = __ldg(q + 1); pg[i][2] = __ldg(q + 2);
        }
        // branch-free per element (only the store is predicated): the four gathers above stay in flight together
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int kk = max(k[i], 0);
            const int py = kk / KEY_STRIDE, px = kk - py * KEY_STRIDE;
            const int sx = ra[i].w, sy = rb[i].x, w = rb[i].y;
            const int64_t off = (static_cast<int64_t>(rb[i].w) << 32) | static_cast<uint32_t>(rb[i].z);
            const int64_t dst = off + static_cast<int64_t>(py - sy) * w + (px - sx);
            const float T = (k[i] != kprev) ? 1.0f : yprev;
            const float pgl = pg[i][0] * __int_as_float(ra[i].x) + pg[i][1] * __int_as_float(ra[i].y) +
                              pg[i][2] * __int_as_float(ra[i].z);
            // dead elements (inclusive product 0) carry no gradient, gs_model.py:575-578
            const bool alive = y[i] != 0.0f;
            const float2 out = make_float2(alive ? T * pgl - t[i] : 0.0f, alive ? T * (1.0f - xx[i]) * pgl : 0.0f);
            if (e0 + i < N) elem[dst] = out;
            kprev = k[i];
            yprev = y[i];
        }
    }
}

// Backward, step 2 of 2 (Gaussian-major order): the reference's per-element gradients (gs_model.py:733-766)
// summed over each Gaussian's box — a segmented reduction without float atomics, deterministic:
//   d_opacity = sum g dalpha            d_l[c] = (sum d) / l[c]      (the reference's d/l, :763-766)
//   d_mean    = sum alpha dalpha (r-m)Lambda          d_Lambda = sum -1/2 alpha dalpha (r-m)^T (r-m)
// Work is balanced over box sizes from a few pixels to >100 K (bundled scene, C2):
//   k_splat_bwd_reduce        boxes of <= RED_SMALL elements: 8 lanes per Gaussian.  Larger boxes are cut into
//                             pieces of RED_PIECE elements and appended to a piece list (an integer atomic
//                             reserves the contiguous slots; slot order does not enter any float sum).
//   k_splat_bwd_reduce_pieces one warp per piece -> one partial 7-vector per slot.
//   k_splat_bwd_reduce_final  the partials of one Gaussian are added in piece order.
constexpr int RED_SMALL = 256;
constexpr int RED_PIECE = 1024;

struct RedAcc {
    float o = 0.f, d = 0.f, m0 = 0.f, m1 = 0.f, a00 = 0.f, a01 = 0.f, a11 = 0.f;
};

// inv_w = 1/w in float: (local + 0.5) * inv_w truncates to local / w for local < 2^21 (the +-1 fix-up below covers
// the rounding of inv_w); larger boxes take the integer division
__device__ __forceinline__ void red_add(RedAcc &A, const float2 v, const int local, const int w, const float inv_w,
                                        const int sx, const int sy, const Gauss &G) {
    int iy, ix;
    if (local < (1 << 21)) {
        iy = __float2int_rz((static_cast<float>(local) + 0.5f) * inv_w);
        ix = local - iy * w;
        if (ix < 0) { ix += w; --iy; }
        if (ix >= w) { ix -= w; ++iy; }
    } else {
        iy = local / w;
        ix = local - iy * w;
    }
    const float d0 = static_cast<float>(sx + ix) - G.mx, d1 = static_cast<float>(sy + iy) - G.my;
    const float X0 = d0 * G.l00 + d1 * G.l10, X1 = d0 * G.l01 + d1 * G.l11;
    const float gk = expf(-0.5f * (X0 * d0 + X1 * d1));
    const float coef = G.o * gk * v.x;
    A.o = fmaf(gk, v.x, A.o);
    A.d += v.y;
    A.m0 = fmaf(coef, X0, A.m0);
    A.m1 = fmaf(coef, X1, A.m1);
    const float hc = -0.5f * coef;
    A.a00 = fmaf(hc * d0, d0, A.a00);
    A.a01 = fmaf(hc * d0, d1, A.a01);
    A.a11 = fmaf(hc * d1, d1, A.a11);
}

template <int LANES>
__device__ __forceinline__ void red_lanes(RedAcc &A) {
#pragma unroll
    for (int d = LANES / 2; d >= 1; d >>= 1) {
        A.o += __shfl_xor_sync(0xffffffffu, A.o, d);
        A.d += __shfl_xor_sync(0xffffffffu, A.d, d);
        A.m0 += __shfl_xor_sync(0xffffffffu, A.m0, d);
        A.m1 += __shfl_xor_sync(0xffffffffu, A.m1, d);
        A.a00 += __shfl_xor_sync(0xffffffffu, A.a00, d);
        A.a01 += __shfl_xor_sync(0xffffffffu, A.a01, d);
        A.a11 += __shfl_xor_sync(0xffffffffu, A.a11, d);
    }
}

__device__ __forceinline__ void red_store(const RedAcc &A, const int64_t g, const float *__restrict__ l_d,
                                          float *__restrict__ g_mean, float *__restrict__ g_lam,
                                          float *__restrict__ g_opac, float *__restrict__ g_l) {
    g_opac[g] = A.o;
    g_l[3 * g] = A.d / __ldg(l_d + 3 * g);
    g_l[3 * g + 1] = A.d / __ldg(l_d + 3 * g + 1);
    g_l[3 * g + 2] = A.d / __ldg(l_d + 3 * g + 2);
    g_mean[2 * g] = A.m0;
    g_mean[2 * g + 1] = A.m1;
    g_lam[4 * g] = A.a00;
    g_lam[4 * g + 1] = A.a01;
    g_lam[4 * g + 2] = A.a01;
    g_lam[4 * g + 3] = A.a11;
}

__global__ void __launch_bounds__(256)
k_splat_bwd_reduce(const float2 *__restrict__ elem, const int32_t *__restrict__ sp, const int32_t *__restrict__ ep,
                   const int64_t *__restrict__ goff, const float *__restrict__ mean, const float *__restrict__ lam,
                   const float *__restrict__ opac, const float *__restrict__ l_d, int64_t n,
                   float *__restrict__ g_mean, float *__restrict__ g_lam, float *__restrict__ g_opac,
                   float *__restrict__ g_l, unsigned int *__restrict__ pcount, int32_t *__restrict__ piece_g,
                   int32_t *__restrict__ piece_i) {
    constexpr int GL = 8;
    const int sub = threadIdx.x & (GL - 1);
    const int64_t grp0 = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) / GL;
    const int64_t ngrp = (static_cast<int64_t>(gridDim.x) * blockDim.x) / GL;
    const int64_t iters = (n + ngrp - 1) / ngrp;  // same trip count for every lane: shuffles stay converged
    for (int64_t it = 0; it < iters; ++it) {
        const int64_t g = grp0 + it * ngrp;
        const bool live = g < n;
        const int64_t gg = live ? g : 0;
        const int64_t b = __ldg(goff + gg);
        int64_t eend = live ? __ldg(goff + gg + 1) : b;
        const bool big = eend - b > RED_SMALL;
        if (big) {
            if (sub == 0) {
                const int np = static_cast<int>((eend - b + RED_PIECE - 1) / RED_PIECE);
                const unsigned int base = atomicAdd(pcount, static_cast<unsigned int>(np));
                for (int i = 0; i < np; ++i) {
                    piece_g[base + i] = static_cast<int32_t>(g);
                    piece_i[base + i] = i;
                }
            }
            eend = b;
        }
        const Gauss G = load_gauss(mean, lam, opac, static_cast<int>(gg));
        const int sx = __ldg(sp + 2 * gg), sy = __ldg(sp + 2 * gg + 1);
        const int w = __ldg(ep + 2 * gg) - sx + 1;
        const float inv_w = 1.0f / static_cast<float>(w);
        RedAcc A;
        for (int64_t e = b + sub; e < eend; e += GL)
            red_add(A, __ldg(elem + e), static_cast<int>(e - b), w, inv_w, sx, sy, G);
        red_lanes<GL>(A);
        if (live && !big && sub == 0) red_store(A, g, l_d, g_mean, g_lam, g_opac, g_l);
    }
}

__global__ void __launch_bounds__(256)
k_splat_bwd_reduce_pieces(const float2 *__restrict__ elem, const int32_t *__restrict__ sp,
                          const int32_t *__restrict__ ep, const int64_t *__restrict__ goff,
                          const float *__restrict__ mean, const float *__restrict__ lam,
                          const float *__restrict__ opac, const unsigned int *__restrict__ pcount,
                          const int32_t *__restrict__ piece_g, const int32_t *__restrict__ piece_i,
                          float4 *__restrict__ partial) {
    const int lane = threadIdx.x & 31;
    const unsigned int np = *pcount;
    const unsigned int nwarps = gridDim.x * (blockDim.x >> 5);
    for (unsigned int p = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); p < np; p += nwarps) {
        const int g = __ldg(piece_g + p);
        const int64_t g0 = __ldg(goff + g);
        const int64_t b = g0 + static_cast<int64_t>(__ldg(piece_i + p)) * RED_PIECE;
        const int64_t eend = min(__ldg(goff + g + 1), b + RED_PIECE);
        const Gauss G = load_gauss(mean, lam, opac, g);
        const int sx = __ldg(sp + 2 * g), sy = __ldg(sp + 2 * g + 1);
        const int w = __ldg(ep + 2 * g) - sx + 1;
        const float inv_w = 1.0f / static_cast<float>(w);
        RedAcc A;
#pragma unroll 4
        for (int64_t e = b + lane; e < eend; e += 32)
            red_add(A, __ldg(elem + e), static_cast<int>(e - g0), w, inv_w, sx, sy, G);
        red_lanes<32>(A);
        if (lane == 0) {
            partial[2 * static_cast<size_t>(p)] = make_float4(A.o, A.d, A.m0, A.m1);
            partial[2 * static_cast<size_t>(p) + 1] = make_float4(A.a00, A.a01, A.a11, 0.f);
        }
    }
}

__global__ void __launch_bounds__(256)
k_splat_bwd_reduce_final(const int64_t *__restrict__ goff, const float *__restrict__ l_d,
                         const unsigned int *__restrict__ pcount, const int32_t *__restrict__ piece_g,
                         const int32_t *__restrict__ piece_i, const float4 *__restrict__ partial,
                         float *__restrict__ g_mean, float *__restrict__ g_lam, float *__restrict__ g_opac,
                         float *__restrict__ g_l) {
    const unsigned int np = *pcount;
    for (unsigned int p = blockIdx.x * blockDim.x + threadIdx.x; p < np; p += gridDim.x * blockDim.x) {
        if (__ldg(piece_i + p) != 0) continue;
        const int g = __ldg(piece_g + p);
        const int cnt = static_cast<int>((__ldg(goff + g + 1) - __ldg(goff + g) + RED_PIECE - 1) / RED_PIECE);
        RedAcc A;
        for (int i = 0; i < cnt; ++i) {
            const float4 u = partial[2 * static_cast<size_t>(p + i)], v = partial[2 * static_cast<size_t>(p + i) + 1];
            A.o += u.x; A.d += u.y; A.m0 += u.z; A.m1 += u.w;
            A.a00 += v.x; A.a01 += v.y; A.a11 += v.z;
        }
        red_store(A, g, l_d, g_mean, g_lam, g_opac, g_l);
    }
}

// =====================================================================================================
// Sort-free placement: the sorted element list is built directly, without sorting
// the N elements.  A stable sort by pixel key is a counting sort whose rank of element (Gaussian j, pixel p)
// is the number of Gaussians i < j that cover p — and the input already arrives in depth (= index) order.
// The image is cut into cells of one row x 64 pixels (GCP_SEG_SHIFT).
//   1. k_place_pairs : one (cell, Gaussian) pair per box row and strip the box touches, P ~ N/6 pairs; the
//                      pairs are sorted by cell (a small stable radix sort on <= 16 bits, Gaussian order kept).
//   2. k_place_count : one warp per cell adds +1/-1 at the ends of every (clipped) interval [sx,ex] of the
//                      cell into a shared-memory difference array and prefix-sums it: per-pixel list lengths.
//   3. exclusive scan of the lengths over the pixels in key order: segment offsets.
//   4. k_place_fill  : one warp per cell walks the cell's intervals in Gaussian order; lanes cover the pixels
//                      of the interval, each pixel keeps a running counter in shared memory, and the element
//                      is written at offset[pixel] + counter: exactly the stable-sort position.
// Output is bit-identical to expand + stable sort (checked in tests/test_compositor.py).
// =====================================================================================================
#ifndef GCP_SEG_SHIFT
#define GCP_SEG_SHIFT 6
#endif
constexpr int SEG_SHIFT = GCP_SEG_SHIFT;  // image rows are cut into strips of 2^SEG_SHIFT pixels:
constexpr int SEGW = 1 << SEG_SHIFT;      // one warp owns one (row, strip) cell at a time

// pair key of cell (row y, strip s) = y * nseg + s
__global__ void __launch_bounds__(256)
k_place_pairs(const int32_t *__restrict__ sp, const int32_t *__restrict__ ep, const int64_t *__restrict__ poff,
              int64_t n, int64_t P, int nseg, int32_t *__restrict__ pcell, int32_t *__restrict__ pgid) {
    const int64_t p0 = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) * CH;
    if (p0 >= P) return;
    int64_t g = find_gaussian(poff, n, p0);
    int64_t gbeg = __ldg(poff + g), gend = __ldg(poff + g + 1);
    int sy = __ldg(sp + 2 * g + 1);
    int s0 = __ldg(sp + 2 * g) >> SEG_SHIFT;
    int ns = (__ldg(ep + 2 * g) >> SEG_SHIFT) - s0 + 1;
    for (int i = 0; i < CH && p0 + i < P; ++i) {
        const int64_t p = p0 + i;
        while (p >= gend) {
            ++g;
            gbeg = gend;
            gend = __ldg(poff + g + 1);
            sy = __ldg(sp + 2 * g + 1);
            s0 = __ldg(sp + 2 * g) >> SEG_SHIFT;
            ns = (__ldg(ep + 2 * g) >> SEG_SHIFT) - s0 + 1;
        }
        const int local = static_cast<int>(p - gbeg);   // row-major over (box row, strip)
        const int r = local / ns;
        pcell[p] = (sy + r) * nseg + s0 + (local - r * ns);
        pgid[p] = static_cast<int32_t>(g);
    }
}

// cstart[c] = first pair of cell c inside the cell-sorted pair list (cstart[ncell] = P): one thread per pair
// writes the entries of every cell that begins at its position (empty cells included)
__global__ void __launch_bounds__(256)
k_place_cellstart(const int32_t *__restrict__ pcell_s, int64_t P, int ncell, int32_t *__restrict__ cstart) {
    const int64_t p = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (p > P) return;
    const int prev = (p == 0) ? -1 : __ldg(pcell_s + p - 1);
    const int cur = (p == P) ? ncell : __ldg(pcell_s + p);
    for (int c = prev + 1; c <= cur; ++c) cstart[c] = static_cast<int32_t>(p);
}

// one warp per (row, strip) cell: per-pixel list lengths from a shared-memory difference array
__global__ void __launch_bounds__(256)
k_place_count(const int32_t *__restrict__ cstart, const int32_t *__restrict__ pgid_s,
              const int32_t *__restrict__ sp, const int32_t *__restrict__ ep, int64_t P, int W, int H, int nseg,
              int32_t *__restrict__ cnt) {
    __shared__ int32_t sm[8][SEGW + 1];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int cell = blockIdx.x * 8 + wib;
    if (cell >= (H + 1) * nseg) return;
    const int y = cell / nseg, x0 = (cell - y * nseg) << SEG_SHIFT;
    const int x1 = min(W, x0 + SEGW - 1);
    int32_t *c = sm[wib];
    for (int i = lane; i <= SEGW; i += 32) c[i] = 0;
    __syncwarp();
    const int64_t lo = __ldg(cstart + cell), hi = __ldg(cstart + cell + 1);
    for (int64_t p = lo + lane; p < hi; p += 32) {
        const int g = __ldg(pgid_s + p);
        atomicAdd(c + (max(__ldg(sp + 2 * g), x0) - x0), 1);
        atomicAdd(c + (min(__ldg(ep + 2 * g), x1) - x0 + 1), -1);
    }
    __syncwarp();
    int carry = 0;
    for (int i0 = 0; i0 < SEGW; i0 += 32) {
        const int i = i0 + lane;
        int v = c[i];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, v, d);
            if (lane >= d) v += t;
        }
        v += carry;
        if (x0 + i <= x1) cnt[static_cast<int64_t>(y) * (W + 1) + x0 + i] = v;
        carry = __shfl_sync(0xffffffffu, v, 31);
    }
}

// key_s from the pixel-list offsets: streaming, coalesced (the fill below then scatters only the Gaussian ids).
// One block per 256 consecutive pixels (image order = key order): its elements form one contiguous range.
__global__ void __launch_bounds__(256)
k_place_keys_short(const int32_t *__restrict__ off, int npix, int W, int32_t *__restrict__ key_s) {
    __shared__ int32_t so[257];
    __shared__ int32_t sk[256];
    const int p0 = blockIdx.x * 256;
    const int np = min(256, npix - p0);
    for (int i = threadIdx.x; i <= np; i += 256) so[i] = __ldg(off + p0 + i);
    if (threadIdx.x < np) {
        const int p = p0 + threadIdx.x, y = p / (W + 1);
        sk[threadIdx.x] = y * KEY_STRIDE + (p - y * (W + 1));
    }
    __syncthreads();
    const int s0 = so[0], s1 = so[np];
    // aligned groups of 4 elements per thread: one search, then a walk; interior groups are one 16-byte store
    for (int c = (s0 >> 2) + threadIdx.x; c < ((s1 + 3) >> 2); c += 256) {
        const int e0 = c << 2;
        const int lo = max(e0, s0), hi = min(e0 + 4, s1);
        int a = 0, b = np;  // last pixel j with so[j] <= lo (empty lists share their successor's offset)
        while (b - a > 1) {
            const int m = (a + b) >> 1;
            if (so[m] <= lo) a = m;
            else b = m;
        }
        int v[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int e = e0 + i;
            if (e >= lo && e < hi) {
                while (so[a + 1] <= e) ++a;
                v[i] = sk[a];
            } else {
                v[i] = 0;
            }
        }
        if (lo == e0 && hi == e0 + 4) {
            *reinterpret_cast<int4 *>(key_s + e0) = make_int4(v[0], v[1], v[2], v[3]);
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if (e0 + i >= lo && e0 + i < hi) key_s[e0 + i] = v[i];
        }
    }
}

// long lists (hundreds of elements per pixel, C2): one warp per 32 consecutive pixels, each list in turn is
// written by the whole warp
__global__ void __launch_bounds__(256)
k_place_keys_long(const int32_t *__restrict__ off, int npix, int W, int32_t *__restrict__ key_s) {
    const int lane = threadIdx.x & 31;
    const int p0 = (blockIdx.x * 8 + (threadIdx.x >> 5)) * 32;
    if (p0 >= npix) return;
    const int p = min(p0 + lane, npix - 1);
    const int st = __ldg(off + p), en = (p0 + lane < npix) ? __ldg(off + p + 1) : st;
    const int y = p / (W + 1);
    const int key = y * KEY_STRIDE + (p - y * (W + 1));
    const int cnt = min(32, npix - p0);
    for (int j = 0; j < cnt; ++j) {
        const int s = __shfl_sync(0xffffffffu, st, j), e = __shfl_sync(0xffffffffu, en, j);
        const int k = __shfl_sync(0xffffffffu, key, j);
        for (int i = s + lane; i < e; i += 32) key_s[i] = k;
    }
}

constexpr int FILL_STAGE = 2048;  // staged elements per warp (8 KB)

// one warp per (row, strip) cell at a time: walk the cell's intervals in Gaussian (depth) order.
// Lane l owns pixels l, l+32, ... of the strip: list offset and running count live in REGISTERS, so an interval
// costs three shuffles and a predicated store per owned pixel — no shared-memory read-modify-write chain between
// consecutive intervals (that chain, not the store traffic, bounded the shared-memory-counter version; a FIFO
// that merged eight appends into one 32-byte store was slower still).
// The grid is persistent and small on purpose (see gcp_splat_place): cells are taken in index order, so the
// lists being filled at any moment form one compact address window that stays in L2 until every 32-byte
// sector is complete — the 4-byte scattered stores then cost no DRAM read-modify-write.
__global__ void __launch_bounds__(256)
k_place_fill(const int32_t *__restrict__ cstart, const int32_t *__restrict__ pgid_s,
             const int32_t *__restrict__ sp, const int32_t *__restrict__ ep, const int32_t *__restrict__ off,
             int64_t P, int W, int H, int nseg, int32_t *__restrict__ gid_s) {
    constexpr int PPL = SEGW / 32;  // pixels per lane
    static_assert(SEGW % 32 == 0, "strip width must be a multiple of the warp size");
    extern __shared__ int32_t fill_stage[];  // FILL_STAGE entries per warp
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    int32_t *stage = fill_stage + wib * FILL_STAGE;
    const int ncell = (H + 1) * nseg;
    for (int cell = blockIdx.x * 8 + wib; cell < ncell; cell += gridDim.x * 8) {
        const int y = cell / nseg, x0 = (cell - y * nseg) << SEG_SHIFT;
        const int x1 = min(W, x0 + SEGW - 1);
        const int64_t lo = __ldg(cstart + cell), hi = __ldg(cstart + cell + 1);
        // the lists of the cell's pixels are one contiguous range [c0, c1) of gid_s: if it fits, the scattered
        // appends go to shared memory and the range leaves as coalesced stores
        const int64_t prow = static_cast<int64_t>(y) * (W + 1);
        const int c0 = __ldg(off + prow + x0), c1 = __ldg(off + prow + x1 + 1);
        const bool staged = c1 - c0 <= FILL_STAGE;
        int32_t *const base = staged ? stage - c0 : gid_s;
        int32_t *dst[PPL];  // next free slot of the owned pixels' lists
#pragma unroll
        for (int q = 0; q < PPL; ++q) {
            const int x = x0 + lane + 32 * q;
            dst[q] = base + ((x <= x1) ? __ldg(off + prow + x) : c0);
        }
        // intervals in batches of 32 (one per lane), software-pipelined two deep: the ids of batch b+2 and the box
        // columns of batch b+1 are in flight while batch b is walked (the columns depend on the ids)
        int g = 0, g1 = 0, rs = KEY_STRIDE, re = -1;
        if (lo + lane < hi) g = __ldg(pgid_s + lo + lane);
        if (lo + 32 + lane < hi) g1 = __ldg(pgid_s + lo + 32 + lane);
        if (lo + lane < hi) { rs = __ldg(sp + 2 * g); re = __ldg(ep + 2 * g); }
        for (int64_t b = lo; b < hi; b += 32) {
            const int a = max(rs, x0) - x0, z = min(re, x1) - x0;
            int g2 = 0, rs1 = KEY_STRIDE, re1 = -1;
            if (b + 64 + lane < hi) g2 = __ldg(pgid_s + b + 64 + lane);
            if (b + 32 + lane < hi) { rs1 = __ldg(sp + 2 * g1); re1 = __ldg(ep + 2 * g1); }
            const int m = static_cast<int>(hi - b < 32 ? hi - b : 32);
#pragma unroll 4
            for (int k = 0; k < m; ++k) {
                const int gg = __shfl_sync(0xffffffffu, g, k);
                const int ia = __shfl_sync(0xffffffffu, a, k);
                const int iz = __shfl_sync(0xffffffffu, z, k);
#pragma unroll
                for (int q = 0; q < PPL; ++q) {
                    const int i = lane + 32 * q;
                    if (i >= ia && i <= iz) *dst[q]++ = gg;
                }
            }
            g = g1; g1 = g2; rs = rs1; re = re1;
        }
        if (staged) {
            __syncwarp();
            for (int i = lane; i < c1 - c0; i += 32) gid_s[c0 + i] = stage[i];
            __syncwarp();
        }
    }
}

// Long lists (hundreds of elements per pixel, wide boxes — the bundled scene, C2): the scattered stores of the
// kernel above, one 4-byte request per element, saturate the load/store path.  Transposed walk instead: the
// lanes hold 32 intervals of the cell; for every pixel of the strip in turn, the intervals covering it are
// compacted with a ballot and written as one contiguous run.  Same output, coalesced stores.
__global__ void __launch_bounds__(256)
k_place_fill_long(const int32_t *__restrict__ cstart, const int32_t *__restrict__ pgid_s,
                  const int32_t *__restrict__ sp, const int32_t *__restrict__ ep, const int32_t *__restrict__ off,
                  int64_t P, int W, int H, int nseg, int32_t *__restrict__ gid_s, unsigned int *__restrict__ ticket,
                  int32_t *__restrict__ btab, int64_t nslot) {
    constexpr int PPL = SEGW / 32;
    const int lane = threadIdx.x & 31;
    const int ncell = (H + 1) * nseg;
    const unsigned lt = (1u << lane) - 1u;
    for (;;) {
        // cells are handed out dynamically (their sizes differ by an order of magnitude)
        int cell = 0;
        if (lane == 0) cell = static_cast<int>(atomicAdd(ticket, 1u));
        cell = __shfl_sync(0xffffffffu, cell, 0);
        if (cell >= ncell) break;
        const int y = cell / nseg, x0 = (cell - y * nseg) << SEG_SHIFT;
        const int x1 = min(W, x0 + SEGW - 1);
        const int64_t lo = __ldg(cstart + cell), hi = __ldg(cstart + cell + 1);
        int pos[PPL];  // next free slot of the owned pixels' lists (lane l owns pixels l, l+32, ...)
#pragma unroll
        for (int q = 0; q < PPL; ++q) {
            const int x = x0 + lane + 32 * q;
            pos[q] = (x <= x1) ? __ldg(off + static_cast<int64_t>(y) * (W + 1) + x) : 0;
        }
        int g = 0, g1 = 0, rs = KEY_STRIDE, re = -1;
        if (lo + lane < hi) g = __ldg(pgid_s + lo + lane);
        if (lo + 32 + lane < hi) g1 = __ldg(pgid_s + lo + 32 + lane);
        if (lo + lane < hi) { rs = __ldg(sp + 2 * g); re = __ldg(ep + 2 * g); }
        for (int64_t b = lo; b < hi; b += 32) {
            const int a = max(rs, x0) - x0, z = min(re, x1) - x0;  // empty (a > z) for lanes past the end
            int g2 = 0, rs1 = KEY_STRIDE, re1 = -1;
            if (b + 64 + lane < hi) g2 = __ldg(pgid_s + b + 64 + lane);
            if (b + 32 + lane < hi) { rs1 = __ldg(sp + 2 * g1); re1 = __ldg(ep + 2 * g1); }
            if (btab != nullptr) {
                // batch table for the backward (gcp_splat_bwd_elem_cells): slot -> cell, first pair, list positions
                // at the start of the batch.  slot = b/32 + cell is unique and increasing along the pair list.
                const int64_t slot = (b >> 5) + cell;
                if (lane == 0) {
                    btab[slot] = cell;
                    btab[nslot + slot] = static_cast<int32_t>(b);
                }
#pragma unroll
                for (int q = 0; q < PPL; ++q) btab[2 * nslot + slot * SEGW + lane + 32 * q] = pos[q];
            }
#pragma unroll
            for (int q = 0; q < PPL; ++q) {
                int add = 0;  // elements appended to the pixel this lane owns
#pragma unroll 8
                for (int j = 0; j < 32; ++j) {
                    const int i = 32 * q + j;
                    const bool cov = (a <= i) && (i <= z);
                    const unsigned m = __ballot_sync(0xffffffffu, cov);
                    const int base = __shfl_sync(0xffffffffu, pos[q], j);
                    if (cov) gid_s[base + __popc(m & lt)] = g;
                    if (lane == j) add = __popc(m);
                }
                pos[q] += add;
            }
            g = g1; g1 = g2; rs = rs1; re = re1;
        }
    }
}

// Backward step 1 for LONG pixel lists (C2): the un-sort as a transposition through shared memory, every global
// access coalesced.  The per-element version above issues one scattered 8-byte store (and one 32-byte gather)
// per element, and the request rate of scattered accesses bounds it.  Here a warp walks a (row, strip) cell of
// the forward placement again: its lanes hold 32 of the cell's (Gaussian, interval) pairs — box, colour and
// Gaussian-major row address in registers, loaded once per pair instead of once per element.
//   read phase : for every pixel of a 32-pixel half strip, the pairs covering it are compacted with a ballot —
//                they are exactly the next elements of that pixel's list, in order — so incl / x / T*U are read
//                as contiguous runs; (dalpha, d) goes to a padded shared-memory tile [pair][pixel];
//   write phase: for every pair, the lanes run along its pixels and store the row segment of the Gaussian-major
//                array elem as one contiguous run.
// Same values as k_splat_bwd_elem, bit for bit (same expressions).
constexpr int BEL_WARPS = 8;
constexpr int BEL_STRIDE = 33;  // float2 entries per tile row: odd -> conflict-free column writes and row reads

__global__ void __launch_bounds__(BEL_WARPS * 32)
k_splat_bwd_elem_cells(const float *__restrict__ incl, const float *__restrict__ x_s, const float *__restrict__ tu,
                       const int4 *__restrict__ rec_b, const float *__restrict__ gimg,
                       const int32_t *__restrict__ off, const int32_t *__restrict__ cstart,
                       const int32_t *__restrict__ pgid_s, const int32_t *__restrict__ btab, int64_t nslot, int W,
                       int nseg, float2 *__restrict__ elem) {
    extern __shared__ float2 bel_tiles[];
    constexpr int PPL = SEGW / 32;
    const int lane = threadIdx.x & 31;
    float2 *tile = bel_tiles + (threadIdx.x >> 5) * (32 * BEL_STRIDE);
    const unsigned lt = (1u << lane) - 1u;
    const int64_t nwarp = static_cast<int64_t>(gridDim.x) * BEL_WARPS;
    // one warp per batch (32 pairs of one cell); the batches are independent: the forward recorded the list
    // positions at the start of every batch
    for (int64_t slot = static_cast<int64_t>(blockIdx.x) * BEL_WARPS + (threadIdx.x >> 5); slot < nslot; slot += nwarp) {
        const int cell = __ldg(btab + slot);
        if (cell < 0) continue;  // unused slot
        const int64_t b = __ldg(btab + nslot + slot);
        const int y = cell / nseg, x0 = (cell - y * nseg) << SEG_SHIFT;
        const int x1 = min(W, x0 + SEGW - 1);
        const int64_t hi = __ldg(cstart + cell + 1);
        const int nv = static_cast<int>(hi - b < 32 ? hi - b : 32);
        int4 ra = make_int4(0, 0, 0, KEY_STRIDE), rb = make_int4(0, 0, 0, 0);  // empty interval: sx > any pixel, w = 0
        if (lane < nv) ldg256(rec_b + 2 * static_cast<int64_t>(__ldg(pgid_s + b + lane)), ra, rb);
        int pos[PPL], st[PPL];  // next element at the start of the batch / first element of the owned pixels' lists
#pragma unroll
        for (int q = 0; q < PPL; ++q) {
            const int x = x0 + lane + 32 * q;
            pos[q] = __ldg(btab + 2 * nslot + slot * SEGW + lane + 32 * q);
            st[q] = (x <= x1) ? __ldg(off + static_cast<int64_t>(y) * (W + 1) + x) : 0;
        }
        const float *grow = gimg + 3 * (static_cast<int64_t>(y) * (W + 1) + x0);
        const int sx = ra.w, sy = rb.x, w = rb.y;
        const int a = max(sx, x0) - x0, z = min(sx + w - 1, x1) - x0;  // a > z for lanes past the end
        const float l0 = __int_as_float(ra.x), l1 = __int_as_float(ra.y), l2 = __int_as_float(ra.z);
        // Gaussian-major address of strip pixel 0 on this row (may lie before the box: only [a, z] is used)
        const int64_t rowbase = ((static_cast<int64_t>(rb.w) << 32) | static_cast<uint32_t>(rb.z)) +
                                static_cast<int64_t>(y - sy) * w + (x0 - sx);
#pragma unroll
        for (int q = 0; q < PPL; ++q) {
            if (x0 + 32 * q > x1) break;
            // read phase in groups of 8 pixels: positions first (ballots only), then all loads, then the math
#pragma unroll 1
            for (int j0 = 0; j0 < 32; j0 += 8) {
                int e[8], first[8];
                bool cov[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int i = 32 * q + j0 + u;
                    cov[u] = (a <= i) && (i <= z);
                    const unsigned m = __ballot_sync(0xffffffffu, cov[u]);
                    e[u] = __shfl_sync(0xffffffffu, pos[q], j0 + u) + __popc(m & lt);
                    first[u] = __shfl_sync(0xffffffffu, st[q], j0 + u);
                }
                float yv[8], xv[8], tv[8], yp[8], p0[8], p1[8], p2[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int i = 32 * q + j0 + u;
                    yv[u] = cov[u] ? __ldg(incl + e[u]) : 0.0f;
                    xv[u] = cov[u] ? __ldg(x_s + e[u]) : 1.0f;
                    tv[u] = cov[u] ? __ldg(tu + e[u]) : 0.0f;
                    yp[u] = (cov[u] && e[u] != first[u]) ? __ldg(incl + e[u] - 1) : 1.0f;
                    p0[u] = cov[u] ? __ldg(grow + 3 * i) : 0.0f;
                    p1[u] = cov[u] ? __ldg(grow + 3 * i + 1) : 0.0f;
                    p2[u] = cov[u] ? __ldg(grow + 3 * i + 2) : 0.0f;
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    if (cov[u]) {
                        const float T = yp[u];
                        const float pgl = p0[u] * l0 + p1[u] * l1 + p2[u] * l2;
                        // dead elements (inclusive product 0) carry no gradient, gs_model.py:575-578
                        const bool alive = yv[u] != 0.0f;
                        tile[lane * BEL_STRIDE + j0 + u] =
                            make_float2(alive ? T * pgl - tv[u] : 0.0f, alive ? T * (1.0f - xv[u]) * pgl : 0.0f);
                    }
                }
            }
            __syncwarp();
            for (int k = 0; k < nv; ++k) {
                const int ka = __shfl_sync(0xffffffffu, a, k), kz = __shfl_sync(0xffffffffu, z, k);
                if (kz < 32 * q || ka > 32 * q + 31) continue;  // the pair does not touch this half strip
                const int64_t kb = __shfl_sync(0xffffffffu, rowbase, k);
                const int i = 32 * q + lane;
                if (ka <= i && i <= kz) elem[kb + i] = tile[k * BEL_STRIDE + lane];
            }
            __syncwarp();
        }
    }
}

// (cell, Gaussian) pairs of one box: rows x strips it touches (0 for an empty / inverted box)
struct PairCount {
    const int32_t *sp, *ep;
    __host__ __device__ __forceinline__ int64_t operator()(int64_t g) const {
        const int sx = sp[2 * g], sy = sp[2 * g + 1], ex = ep[2 * g], ey = ep[2 * g + 1];
        if (ex < sx || ey < sy) return 0;
        return static_cast<int64_t>(ey - sy + 1) * ((ex >> SEG_SHIFT) - (sx >> SEG_SHIFT) + 1);
    }
};

__global__ void k_prepare_totals(const int64_t *__restrict__ goff, const int64_t *__restrict__ poff, int64_t n,
                                 int64_t *__restrict__ totals) {
    totals[0] = goff[n];
    totals[1] = poff[n];
}

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

}  // namespace

extern "C" {

int gcp_splat_expand(const int32_t *sp, const int32_t *ep, const int64_t *goff, int64_t n, int64_t N,
                     int32_t *key, int32_t *gid, gcp_stream_t stream) {
    if (n < 0 || N < 0) return GCP_ERR_INVALID_ARG;
    if (N == 0) return GCP_OK;
    if (!sp || !ep || !goff || !key || !gid) return GCP_ERR_INVALID_ARG;
    k_splat_expand<<<blocks_for(N, 256 * CH), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(sp, ep, goff, n, N,
                                                                                                  key, gid);
    return static_cast<int>(cudaGetLastError());
}

size_t gcp_splat_sort_bytes(int64_t N) {
    size_t bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, bytes, static_cast<const int32_t *>(nullptr),
                                    static_cast<int32_t *>(nullptr), static_cast<const int32_t *>(nullptr),
                                    static_cast<int32_t *>(nullptr), N > 0 ? N : 1, 0, 32);
    return bytes;
}

int gcp_splat_sort(const int32_t *key_in, const int32_t *gid_in, int32_t *key_out, int32_t *gid_out, int64_t N,
                   int max_key, void *temp, size_t temp_bytes, gcp_stream_t stream) {
    if (N < 0 || max_key < 0) return GCP_ERR_INVALID_ARG;
    if (N == 0) return GCP_OK;
    if (!key_in || !gid_in || !key_out || !gid_out || !temp) return GCP_ERR_INVALID_ARG;
    // LSD radix sort over the significant key bits only: stable, so depth order inside a pixel survives
    return static_cast<int>(cub::DeviceRadixSort::SortPairs(temp, temp_bytes, key_in, key_out, gid_in, gid_out, N, 0,
                                                            key_bits(max_key),
                                                            reinterpret_cast<cudaStream_t>(stream)));
}

size_t gcp_splat_prepare_bytes(int64_t n) {
    size_t a = 0;
    cub::DeviceScan::InclusiveSum(nullptr, a, static_cast<const int64_t *>(nullptr), static_cast<int64_t *>(nullptr),
                                  n > 0 ? n : 1);
    return a + 256;  // the transform-iterator scan needs no more than the plain one of the same value type
}

int gcp_splat_prepare(const int64_t *boxsize, const int32_t *sp, const int32_t *ep, int64_t n, int64_t *goff,
                      int64_t *poff, int64_t *totals, void *temp, size_t temp_bytes, gcp_stream_t stream) {
    if (n < 0 || !goff || !poff || !totals) return GCP_ERR_INVALID_ARG;
    auto st = reinterpret_cast<cudaStream_t>(stream);
    cudaError_t e = cudaMemsetAsync(goff, 0, 8, st);
    if (e != cudaSuccess) return static_cast<int>(e);
    e = cudaMemsetAsync(poff, 0, 8, st);
    if (e != cudaSuccess) return static_cast<int>(e);
    if (n > 0) {
        if (!boxsize || !sp || !ep || !temp) return GCP_ERR_INVALID_ARG;
        size_t need = 0;
        auto pairs = thrust::make_transform_iterator(thrust::counting_iterator<int64_t>(0), PairCount{sp, ep});
        cub::DeviceScan::InclusiveSum(nullptr, need, pairs, poff + 1, n);
        size_t need2 = 0;
        cub::DeviceScan::InclusiveSum(nullptr, need2, boxsize, goff + 1, n);
        if (need2 > need) need = need2;
        if (temp_bytes < need) return GCP_ERR_WORKSPACE;
        size_t tb = temp_bytes;
        e = cub::DeviceScan::InclusiveSum(temp, tb, boxsize, goff + 1, n, st);
        if (e != cudaSuccess) return static_cast<int>(e);
        tb = temp_bytes;
        e = cub::DeviceScan::InclusiveSum(temp, tb, pairs, poff + 1, n, st);
        if (e != cudaSuccess) return static_cast<int>(e);
    }
    k_prepare_totals<<<1, 1, 0, st>>>(goff, poff, n, totals);
    return static_cast<int>(cudaGetLastError());
}

int gcp_splat_pack(const float *mean, const float *lam, const float *opac, const float *l_d, const int32_t *sp,
                   const int32_t *ep, const int64_t *goff, int64_t n, float *rec_a, int32_t *rec_b,
                   gcp_stream_t stream) {
    if (n < 0) return GCP_ERR_INVALID_ARG;
    if (n == 0) return GCP_OK;
    if ((reinterpret_cast<uintptr_t>(rec_a) | reinterpret_cast<uintptr_t>(rec_b)) & 31) return GCP_ERR_INVALID_ARG;
    k_splat_pack<<<blocks_for(n, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        mean, lam, opac, l_d, sp, ep, goff, n, reinterpret_cast<float4 *>(rec_a), reinterpret_cast<int4 *>(rec_b));
    return static_cast<int>(cudaGetLastError());
}

int gcp_splat_alpha(const int32_t *key_s, const int32_t *gid_s, const float *rec_a, int64_t N, float *x_s,
                    gcp_stream_t stream) {
    if (N < 0) return GCP_ERR_INVALID_ARG;
    if (N == 0) return GCP_OK;
    k_splat_alpha<<<blocks_for(N, 256 * 4, 148 * 16), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        key_s, gid_s, reinterpret_cast<const float4 *>(rec_a), N, x_s);
    return static_cast<int>(cudaGetLastError());
}

int gcp_splat_color(const float *incl, const float *x_s, const int32_t *key_s, const int32_t *gid_s,
                    const int32_t *rec_b, int64_t N, int W, float *image, gcp_stream_t stream) {
    if (N < 0) return GCP_ERR_INVALID_ARG;
    if (N == 0) return GCP_OK;
    k_splat_color<<<blocks_for(N, 256 * 8, 148 * 16), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        incl, x_s, key_s, gid_s, reinterpret_cast<const int4 *>(rec_b), N, W, image);
    return static_cast<int>(cudaGetLastError());
}

int gcp_splat_bwd_w(const float *incl, const float *x_s, const int32_t *key_s, const int32_t *gid_s,
                    const int32_t *rec_b, const float *grad_image, int64_t N, int W, float *gshift,
                    gcp_stream_t stream) {
    if (N < 0) return GCP_ERR_INVALID_ARG;
    if (N == 0) return GCP_OK;
    k_splat_bwd_w<<<blocks_for(N, 256 * 4, 148 * 16), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        incl, x_s, key_s, gid_s, reinterpret_cast<const int4 *>(rec_b), grad_image, N, W, gshift);
    return static_cast<int>(cudaGetLastError());
}

int gcp_splat_bwd_grads(const float *incl, const float *x_s, const float *tu, const int32_t *key_s,
                        const int32_t *gid_s, const float *mean, const float *lam, const float *opac,
                        const float *l_d, const float *grad_image, int64_t N, int W, float *g_mean, float *g_lam,
                        float *g_opac, float *g_l, gcp_stream_t stream) {
    if (N < 0) return GCP_ERR_INVALID_ARG;
    if (N == 0) return GCP_OK;
    k_splat_bwd_grads<<<blocks_for(N, 256, 148 * 32), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        incl, x_s, tu, key_s, gid_s, mean, lam, opac, l_d, grad_image, N, W, g_mean, g_lam, g_opac, g_l);
    return static_cast<int>(cudaGetLastError());
}

int gcp_splat_bwd_elem(const float *incl, const float *x_s, const float *tu, const int32_t *key_s,
                       const int32_t *gid_s, const int32_t *rec_b, const float *grad_image, int64_t N, int W,
                       float *elem, gcp_stream_t stream) {
    if (N < 0) return GCP_ERR_INVALID_ARG;
    if (N == 0) return GCP_OK;
    k_splat_bwd_elem<<<blocks_for(N, 256 * 4, 148 * 16), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        incl, x_s, tu, key_s, gid_s, reinterpret_cast<const int4 *>(rec_b), grad_image, N, W,
        reinterpret_cast<float2 *>(elem));
    return static_cast<int>(cudaGetLastError());
}

namespace {
struct ReduceLayout {
    size_t piece_g, piece_i, partial, total, cap;
};
ReduceLayout reduce_layout(int64_t N, int64_t n) {
    ReduceLayout l;
    const int64_t nbig = std::min<int64_t>(n, N / RED_SMALL);
    l.cap = static_cast<size_t>(N / RED_PIECE + nbig + 1);
    size_t off = 256;  // header: piece counter
    l.piece_g = off; off += (l.cap * 4 + 255) & ~size_t(255);
    l.piece_i = off; off += (l.cap * 4 + 255) & ~size_t(255);
    l.partial = off; off += l.cap * 32;
    l.total = off;
    return l;
}
}  // namespace

size_t gcp_splat_bwd_reduce_bytes(int64_t N, int64_t n) { return (N < 0 || n < 0) ? 0 : reduce_layout(N, n).total; }

int gcp_splat_bwd_reduce(const float *elem, const int32_t *sp, const int32_t *ep, const int64_t *goff,
                         const float *mean, const float *lam, const float *opac, const float *l_d, int64_t N,
                         int64_t n, float *g_mean, float *g_lam, float *g_opac, float *g_l, void *temp,
                         size_t temp_bytes, gcp_stream_t stream) {
    if (n < 0 || N < 0) return GCP_ERR_INVALID_ARG;
    if (n == 0) return GCP_OK;
    const ReduceLayout l = reduce_layout(N, n);
    if (temp == nullptr || temp_bytes < l.total) return GCP_ERR_WORKSPACE;
    if (reinterpret_cast<uintptr_t>(temp) & 15) return GCP_ERR_INVALID_ARG;
    auto st = reinterpret_cast<cudaStream_t>(stream);
    char *t = static_cast<char *>(temp);
    auto *pcount = reinterpret_cast<unsigned int *>(t);
    auto *piece_g = reinterpret_cast<int32_t *>(t + l.piece_g);
    auto *piece_i = reinterpret_cast<int32_t *>(t + l.piece_i);
    auto *partial = reinterpret_cast<float4 *>(t + l.partial);
    cudaError_t e = cudaMemsetAsync(pcount, 0, 256, st);
    if (e != cudaSuccess) return static_cast<int>(e);
    k_splat_bwd_reduce<<<blocks_for(n, 32, 148 * 16), 256, 0, st>>>(
        reinterpret_cast<const float2 *>(elem), sp, ep, goff, mean, lam, opac, l_d, n, g_mean, g_lam, g_opac, g_l,
        pcount, piece_g, piece_i);
    const int piece_blocks = static_cast<int>(std::min<size_t>(148 * 8, (l.cap + 7) / 8));
    k_splat_bwd_reduce_pieces<<<piece_blocks, 256, 0, st>>>(reinterpret_cast<const float2 *>(elem), sp, ep, goff,
                                                            mean, lam, opac, pcount, piece_g, piece_i, partial);
    k_splat_bwd_reduce_final<<<std::min<int>(148 * 4, static_cast<int>((l.cap + 255) / 256)), 256, 0, st>>>(
        goff, l_d, pcount, piece_g, piece_i, partial, g_mean, g_lam, g_opac, g_l);
    return static_cast<int>(cudaGetLastError());
}

// ---- sort-free placement -------------------------------------------------------------------------
namespace {
int g_fill_blocks = 0;  // tuning: size of the persistent k_place_fill grid (0 = default)
int g_long_min = 8;     // tuning: pairs per pixel from which the long-list kernels are used
struct PlaceLayout {
    size_t prow, pgid, prow_s, pgid_s, cnt, cstart, cub, total;
    size_t cub_bytes;
};
inline size_t align256(size_t v) { return (v + 255) & ~static_cast<size_t>(255); }
PlaceLayout place_layout(int64_t P, int W, int H) {
    PlaceLayout L;
    const size_t pb = align256(static_cast<size_t>(P > 0 ? P : 1) * 4);
    const size_t npix = static_cast<size_t>(H + 1) * (W + 1) + 1;
    size_t a = 0, b = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, a, static_cast<const int32_t *>(nullptr), static_cast<int32_t *>(nullptr),
                                    static_cast<const int32_t *>(nullptr), static_cast<int32_t *>(nullptr),
                                    P > 0 ? P : 1, 0, 24);
    cub::DeviceScan::ExclusiveSum(nullptr, b, static_cast<const int32_t *>(nullptr), static_cast<int32_t *>(nullptr),
                                  static_cast<int64_t>(npix));
    L.cub_bytes = align256(a > b ? a : b);
    L.prow = 0;
    L.pgid = L.prow + pb;
    L.prow_s = L.pgid + pb;
    L.pgid_s = L.prow_s + pb;
    L.cnt = L.pgid_s + pb;
    L.cstart = L.cnt + align256(npix * 4);  // cells + 1 offsets, then one ticket word
    const size_t ncell = static_cast<size_t>(H + 1) * ((W + SEGW) >> SEG_SHIFT);
    L.cub = L.cstart + align256((ncell + 2) * 4);
    L.total = L.cub + L.cub_bytes;
    return L;
}
}  // namespace

size_t gcp_splat_place_bytes(int64_t P, int W, int H) { return place_layout(P, W, H).total; }

int gcp_splat_seg_shift(void) { return SEG_SHIFT; }

int gcp_splat_set_fill_blocks(int blocks) {
    g_fill_blocks = blocks;
    return GCP_OK;
}

int gcp_splat_set_long_list_threshold(int pairs_per_pixel) {
    if (pairs_per_pixel < 0) return GCP_ERR_INVALID_ARG;
    g_long_min = pairs_per_pixel;
    return GCP_OK;
}

int gcp_splat_long_lists(int64_t P, int W, int H) {
    if (P < 0 || W < 0 || H < 0) return 0;
    return P / (static_cast<int64_t>(H + 1) * (W + 1)) >= g_long_min ? 1 : 0;
}

int gcp_splat_num_cells(int W, int H) { return (W < 0 || H < 0) ? 0 : (H + 1) * ((W + SEGW) >> SEG_SHIFT); }

int64_t gcp_splat_batch_table_ints(int64_t P, int W, int H) {
    if (P < 0 || W < 0 || H < 0) return 0;
    const int64_t nslot = (P >> 5) + gcp_splat_num_cells(W, H) + 1;
    return nslot * (2 + SEGW);
}

int gcp_splat_bwd_elem_cells(const float *incl, const float *x_s, const float *tu, const int32_t *rec_b,
                             const float *grad_image, const int32_t *seg_off, const int32_t *cell_start,
                             const int32_t *pair_gid, const int32_t *batch_table, int64_t P, int W, int H,
                             float *elem, gcp_stream_t stream) {
    if (P < 0 || W < 0 || H < 0 || W >= KEY_STRIDE) return GCP_ERR_INVALID_ARG;
    if (!incl || !x_s || !tu || !rec_b || !grad_image || !seg_off || !cell_start || !pair_gid || !batch_table || !elem)
        return GCP_ERR_INVALID_ARG;
    auto st = reinterpret_cast<cudaStream_t>(stream);
    constexpr int smem = BEL_WARPS * 32 * BEL_STRIDE * static_cast<int>(sizeof(float2));
    static int blocks_per_sm = 0;
    if (blocks_per_sm == 0) {
        cudaError_t e = cudaFuncSetAttribute(k_splat_bwd_elem_cells, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return static_cast<int>(e);
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, k_splat_bwd_elem_cells, BEL_WARPS * 32, smem);
        if (e != cudaSuccess) return static_cast<int>(e);
        if (blocks_per_sm < 1) blocks_per_sm = 1;
    }
    const int nseg = (W + SEGW) >> SEG_SHIFT;
    const int64_t nslot = (P >> 5) + static_cast<int64_t>(H + 1) * nseg + 1;
    const int grid = static_cast<int>(std::min<int64_t>((nslot + BEL_WARPS - 1) / BEL_WARPS, 148 * blocks_per_sm));
    k_splat_bwd_elem_cells<<<grid, BEL_WARPS * 32, smem, st>>>(
        incl, x_s, tu, reinterpret_cast<const int4 *>(rec_b), grad_image, seg_off, cell_start, pair_gid, batch_table,
        nslot, W, nseg, reinterpret_cast<float2 *>(elem));
    return static_cast<int>(cudaGetLastError());
}

int gcp_splat_place(const int32_t *sp, const int32_t *ep, const int64_t *poff, int64_t n, int64_t P, int W, int H,
                    int32_t *key_s, int32_t *gid_s, int32_t *seg_off, int32_t *cell_start, int32_t *pair_gid,
                    int32_t *batch_table, void *temp, size_t temp_bytes, gcp_stream_t stream) {
    if (n < 0 || P < 0 || W < 0 || H < 0 || W >= KEY_STRIDE) return GCP_ERR_INVALID_ARG;
    if (!seg_off || !temp) return GCP_ERR_INVALID_ARG;
    if (reinterpret_cast<uintptr_t>(key_s) & 15) return GCP_ERR_INVALID_ARG;
    const PlaceLayout L = place_layout(P, W, H);
    if (temp_bytes < L.total) return GCP_ERR_WORKSPACE;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    unsigned char *t = static_cast<unsigned char *>(temp);
    int32_t *prow = reinterpret_cast<int32_t *>(t + L.prow), *pgid = reinterpret_cast<int32_t *>(t + L.pgid);
    int32_t *prow_s = reinterpret_cast<int32_t *>(t + L.prow_s);
    int32_t *pgid_s = pair_gid ? pair_gid : reinterpret_cast<int32_t *>(t + L.pgid_s);
    int32_t *cnt = reinterpret_cast<int32_t *>(t + L.cnt);
    const int64_t npix = static_cast<int64_t>(H + 1) * (W + 1) + 1;
    cudaError_t e;
    const int nseg = (W + SEGW) >> SEG_SHIFT;  // ceil((W+1)/SEGW) strips per row
    const int cells = (H + 1) * nseg;
    if (P > 0) {
        k_place_pairs<<<blocks_for(P, 256 * CH), 256, 0, s>>>(sp, ep, poff, n, P, nseg, prow, pgid);
        size_t cb = L.cub_bytes;
        e = cub::DeviceRadixSort::SortPairs(t + L.cub, cb, prow, prow_s, pgid, pgid_s, P, 0, key_bits(cells), s);
        if (e != cudaSuccess) return static_cast<int>(e);
    }
    int32_t *cstart = cell_start ? cell_start : reinterpret_cast<int32_t *>(t + L.cstart);
    unsigned int *ticket = reinterpret_cast<unsigned int *>(reinterpret_cast<int32_t *>(t + L.cstart) + cells + 1);
    k_place_cellstart<<<blocks_for(P + 1, 256), 256, 0, s>>>(prow_s, P, cells, cstart);
    const unsigned blocks = static_cast<unsigned>((cells + 7) / 8);
    k_place_count<<<blocks, 256, 0, s>>>(cstart, pgid_s, sp, ep, P, W, H, nseg, cnt);
    e = cudaMemsetAsync(cnt + (npix - 1), 0, 4, s);  // sentinel: the scan's last output is the element count
    if (e != cudaSuccess) return static_cast<int>(e);
    size_t cb = L.cub_bytes;
    e = cub::DeviceScan::ExclusiveSum(t + L.cub, cb, cnt, seg_off, npix, s);
    if (e != cudaSuccess) return static_cast<int>(e);
    // persistent: three 64 KB-staging blocks per SM; cells taken in index order keep the output window compact
    const unsigned cap = g_fill_blocks > 0 ? static_cast<unsigned>(g_fill_blocks) : 444u;
    unsigned fill_blocks = blocks < cap ? blocks : cap;
    // short lists: per-element search inside 256-pixel blocks; long lists (>= 8 intervals per pixel): warp per list
    const unsigned key_blocks = static_cast<unsigned>((npix - 1 + 255) / 256);
    const bool long_lists = P / (npix - 1) >= g_long_min;
    if (long_lists)
        k_place_keys_long<<<key_blocks, 256, 0, s>>>(seg_off, static_cast<int>(npix - 1), W, key_s);
    else
        k_place_keys_short<<<key_blocks, 256, 0, s>>>(seg_off, static_cast<int>(npix - 1), W, key_s);
    if (long_lists) {
        e = cudaMemsetAsync(ticket, 0, 4, s);
        if (e != cudaSuccess) return static_cast<int>(e);
        const int64_t nslot = (P >> 5) + cells + 1;
        if (batch_table) {
            e = cudaMemsetAsync(batch_table, 0xff, static_cast<size_t>(nslot) * 4, s);  // slot -> cell = -1: unused
            if (e != cudaSuccess) return static_cast<int>(e);
        }
        k_place_fill_long<<<std::min(blocks, 148u * 8u), 256, 0, s>>>(cstart, pgid_s, sp, ep, seg_off, P, W, H, nseg,
                                                                      gid_s, ticket, batch_table, nslot);
    } else {
        constexpr int fill_smem = 8 * FILL_STAGE * 4;
        static bool fill_attr = false;
        if (!fill_attr) {
            e = cudaFuncSetAttribute(k_place_fill, cudaFuncAttributeMaxDynamicSharedMemorySize, fill_smem);
            if (e != cudaSuccess) return static_cast<int>(e);
            fill_attr = true;
        }
        k_place_fill<<<fill_blocks, 256, fill_smem, s>>>(cstart, pgid_s, sp, ep, seg_off, P, W, H, nseg, gid_s);
    }
    return static_cast<int>(cudaGetLastError());
}

}  // extern "C"
