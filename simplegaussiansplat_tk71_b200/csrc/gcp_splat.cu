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

#include <cub/device/device_radix_sort.cuh>

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

// x_s[e] = 1 - o*g in sorted order (the scan's input); 4 consecutive elements per thread
__global__ void __launch_bounds__(256)
k_splat_alpha(const int32_t *__restrict__ key_s, const int32_t *__restrict__ gid_s, const float *__restrict__ mean,
              const float *__restrict__ lam, const float *__restrict__ opac, int64_t N, float *__restrict__ x_s) {
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
        float out[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const Gauss G = load_gauss(mean, lam, opac, g[i]);
            float d0, d1, X0, X1;
            out[i] = 1.0f - G.o * gauss_kernel(G, k[i], d0, d1, X0, X1);
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

// image[pixel] += sum_i T_i alpha_i l_i  (T exclusive = previous inclusive product, 1 at a head;
// elements whose inclusive product is 0 contribute nothing, gs_model.py:575-578)
__global__ void __launch_bounds__(256)
k_splat_color(const float *__restrict__ incl, const float *__restrict__ x_s, const int32_t *__restrict__ key_s,
              const int32_t *__restrict__ gid_s, const float *__restrict__ l_d, int64_t N, int W,
              float *__restrict__ image) {
    const int64_t e0 = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) * CH;
    if (e0 >= N) return;
    int kprev = (e0 > 0) ? __ldg(key_s + e0 - 1) : -1;
    float yprev = (e0 > 0) ? __ldg(incl + e0 - 1) : 1.0f;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f;
    int kcur = -1;
    for (int i = 0; i < CH && e0 + i < N; ++i) {
        const int64_t e = e0 + i;
        const int k = __ldg(key_s + e);
        const float y = __ldg(incl + e);
        const float T = (k != kprev) ? 1.0f : yprev;
        if (k != kcur) {
            if (kcur >= 0 && (a0 != 0.f || a1 != 0.f || a2 != 0.f)) {
                float *p = image + 3 * static_cast<int64_t>(pixel_index(kcur, W));
                atomicAdd(p, a0); atomicAdd(p + 1, a1); atomicAdd(p + 2, a2);
            }
            kcur = k;
            a0 = a1 = a2 = 0.f;
        }
        if (y != 0.0f) {
            const float ta = T * (1.0f - __ldg(x_s + e));
            const int g = __ldg(gid_s + e);
            a0 = fmaf(ta, __ldg(l_d + 3 * g), a0);
            a1 = fmaf(ta, __ldg(l_d + 3 * g + 1), a1);
            a2 = fmaf(ta, __ldg(l_d + 3 * g + 2), a2);
        }
        kprev = k;
        yprev = y;
    }
    if (kcur >= 0 && (a0 != 0.f || a1 != 0.f || a2 != 0.f)) {
        float *p = image + 3 * static_cast<int64_t>(pixel_index(kcur, W));
        atomicAdd(p, a0); atomicAdd(p + 1, a1); atomicAdd(p + 2, a2);
    }
}

// w_k = <dL/dI(pixel), alpha_k l_k> (0 for dead elements); gshift[k] = w_{k+1} inside a pixel list, 0 at its tail:
// the grad_out that makes grouped_cumprod_backward return T_k * U_k.
__device__ __forceinline__ float elem_w(const float *__restrict__ incl, const float *__restrict__ x_s,
                                        const int32_t *__restrict__ gid_s, const float *__restrict__ l_d,
                                        const float *__restrict__ gimg, int64_t e, int key, int W, float &pgl) {
    const int g = __ldg(gid_s + e);
    const float *pg = gimg + 3 * static_cast<int64_t>(pixel_index(key, W));
    pgl = __ldg(pg) * __ldg(l_d + 3 * g) + __ldg(pg + 1) * __ldg(l_d + 3 * g + 1) +
          __ldg(pg + 2) * __ldg(l_d + 3 * g + 2);
    return (__ldg(incl + e) != 0.0f) ? (1.0f - __ldg(x_s + e)) * pgl : 0.0f;
}

__global__ void __launch_bounds__(256)
k_splat_bwd_w(const float *__restrict__ incl, const float *__restrict__ x_s, const int32_t *__restrict__ key_s,
              const int32_t *__restrict__ gid_s, const float *__restrict__ l_d, const float *__restrict__ gimg,
              int64_t N, int W, float *__restrict__ gshift) {
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x * 4;
    for (int64_t e = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) * 4; e < N; e += stride) {
        // elements e+1 .. e+4 are needed (w of the successor); load keys e .. e+4
        int k[5];
#pragma unroll
        for (int i = 0; i < 5; ++i) k[i] = (e + i < N) ? __ldg(key_s + e + i) : -1;
        float out[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            out[i] = 0.0f;
            if (e + i + 1 < N && k[i + 1] == k[i]) {
                float pgl;
                out[i] = elem_w(incl, x_s, gid_s, l_d, gimg, e + i + 1, k[i + 1], W, pgl);
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
__global__ void __launch_bounds__(256)
k_splat_bwd_elem(const float *__restrict__ incl, const float *__restrict__ x_s, const float *__restrict__ tu,
                 const int32_t *__restrict__ key_s, const int32_t *__restrict__ gid_s,
                 const int32_t *__restrict__ sp, const int32_t *__restrict__ ep, const int64_t *__restrict__ goff,
                 const float *__restrict__ l_d, const float *__restrict__ gimg, int64_t N, int W,
                 float2 *__restrict__ elem) {
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x * 4;
    for (int64_t e0 = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) * 4; e0 < N; e0 += stride) {
        int k[4], g[4];
        float y[4], xx[4], t[4];
        if (e0 + 4 <= N) {
            const int4 kk = __ldg(reinterpret_cast<const int4 *>(key_s + e0));
            const int4 gg = __ldg(reinterpret_cast<const int4 *>(gid_s + e0));
            const float4 yy = __ldg(reinterpret_cast<const float4 *>(incl + e0));
            const float4 xv = __ldg(reinterpret_cast<const float4 *>(x_s + e0));
            const float4 tv = __ldg(reinterpret_cast<const float4 *>(tu + e0));
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
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            if (e0 + i < N) {
                const int py = k[i] / KEY_STRIDE, px = k[i] - py * KEY_STRIDE;
                const int2 s2 = __ldg(reinterpret_cast<const int2 *>(sp) + g[i]);
                const int w = __ldg(ep + 2 * g[i]) - s2.x + 1;
                const int64_t dst = __ldg(goff + g[i]) + static_cast<int64_t>(py - s2.y) * w + (px - s2.x);
                float2 out = make_float2(0.0f, 0.0f);
                if (y[i] != 0.0f) {  // dead elements (inclusive product 0) carry no gradient, gs_model.py:575-578
                    const float T = (k[i] != kprev) ? 1.0f : yprev;
                    const float *pg = gimg + 3 * static_cast<int64_t>(py * (W + 1) + px);
                    const float pgl = __ldg(pg) * __ldg(l_d + 3 * g[i]) + __ldg(pg + 1) * __ldg(l_d + 3 * g[i] + 1) +
                                      __ldg(pg + 2) * __ldg(l_d + 3 * g[i] + 2);
                    out.x = T * pgl - t[i];
                    out.y = T * (1.0f - xx[i]) * pgl;
                }
                elem[dst] = out;
            }
            kprev = k[i];
            yprev = y[i];
        }
    }
}

// Backward, step 2 of 2 (Gaussian-major order): one warp per Gaussian sums the reference's per-element
// gradients (gs_model.py:733-766) over its box — a segmented reduction without atomics, deterministic:
//   d_opacity = sum g dalpha            d_l[c] = (sum d) / l[c]      (the reference's d/l, :763-766)
//   d_mean    = sum alpha dalpha (r-m)Lambda          d_Lambda = sum -1/2 alpha dalpha (r-m)^T (r-m)
__global__ void __launch_bounds__(256)
k_splat_bwd_reduce(const float2 *__restrict__ elem, const int32_t *__restrict__ sp, const int32_t *__restrict__ ep,
                   const int64_t *__restrict__ goff, const float *__restrict__ mean, const float *__restrict__ lam,
                   const float *__restrict__ opac, const float *__restrict__ l_d, int64_t n,
                   float *__restrict__ g_mean, float *__restrict__ g_lam, float *__restrict__ g_opac,
                   float *__restrict__ g_l) {
    // a box holds ~40 elements on average: 8 lanes per Gaussian, four Gaussians per warp at a time
    constexpr int GL = 8;
    const int sub = threadIdx.x & (GL - 1);
    const int64_t grp0 = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) / GL;
    const int64_t ngrp = (static_cast<int64_t>(gridDim.x) * blockDim.x) / GL;
    const int64_t iters = (n + ngrp - 1) / ngrp;  // same trip count for every lane: shuffles stay converged
    for (int64_t it = 0; it < iters; ++it) {
        const int64_t g = grp0 + it * ngrp;
        const bool live = g < n;
        const int64_t gg = live ? g : 0;
        const int64_t b = __ldg(goff + gg), eend = live ? __ldg(goff + gg + 1) : b;
        const Gauss G = load_gauss(mean, lam, opac, static_cast<int>(gg));
        const int sx = __ldg(sp + 2 * gg), sy = __ldg(sp + 2 * gg + 1);
        const int w = __ldg(ep + 2 * gg) - sx + 1;
        float a_o = 0.f, a_d = 0.f, a_m0 = 0.f, a_m1 = 0.f, a_00 = 0.f, a_01 = 0.f, a_11 = 0.f;
        for (int64_t e = b + sub; e < eend; e += GL) {
            const float2 v = __ldg(elem + e);
            const int local = static_cast<int>(e - b);
            const int iy = local / w, ix = local - iy * w;
            const float d0 = static_cast<float>(sx + ix) - G.mx, d1 = static_cast<float>(sy + iy) - G.my;
            const float X0 = d0 * G.l00 + d1 * G.l10, X1 = d0 * G.l01 + d1 * G.l11;
            const float gk = expf(-0.5f * (X0 * d0 + X1 * d1));
            const float coef = G.o * gk * v.x;
            a_o = fmaf(gk, v.x, a_o);
            a_d += v.y;
            a_m0 = fmaf(coef, X0, a_m0);
            a_m1 = fmaf(coef, X1, a_m1);
            const float hc = -0.5f * coef;
            a_00 = fmaf(hc * d0, d0, a_00);
            a_01 = fmaf(hc * d0, d1, a_01);
            a_11 = fmaf(hc * d1, d1, a_11);
        }
#pragma unroll
        for (int d = GL / 2; d >= 1; d >>= 1) {
            a_o += __shfl_xor_sync(0xffffffffu, a_o, d);
            a_d += __shfl_xor_sync(0xffffffffu, a_d, d);
            a_m0 += __shfl_xor_sync(0xffffffffu, a_m0, d);
            a_m1 += __shfl_xor_sync(0xffffffffu, a_m1, d);
            a_00 += __shfl_xor_sync(0xffffffffu, a_00, d);
            a_01 += __shfl_xor_sync(0xffffffffu, a_01, d);
            a_11 += __shfl_xor_sync(0xffffffffu, a_11, d);
        }
        if (live && sub == 0) {
            g_opac[g] = a_o;
            g_l[3 * g] = a_d / __ldg(l_d + 3 * g);
            g_l[3 * g + 1] = a_d / __ldg(l_d + 3 * g + 1);
            g_l[3 * g + 2] = a_d / __ldg(l_d + 3 * g + 2);
            g_mean[2 * g] = a_m0;
            g_mean[2 * g + 1] = a_m1;
            g_lam[4 * g] = a_00;
            g_lam[4 * g + 1] = a_01;
            g_lam[4 * g + 2] = a_01;
            g_lam[4 * g + 3] = a_11;
        }
    }
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

int gcp_splat_alpha(const int32_t *key_s, const int32_t *gid_s, const float *mean, const float *lam,
                    const float *opac, int64_t N, float *x_s, gcp_stream_t stream) {
    if (N < 0) return GCP_ERR_INVALID_ARG;
    if (N == 0) return GCP_OK;
    k_splat_alpha<<<blocks_for(N, 256 * 4, 148 * 16), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        key_s, gid_s, mean, lam, opac, N, x_s);
    return static_cast<int>(cudaGetLastError());
}

int gcp_splat_color(const float *incl, const float *x_s, const int32_t *key_s, const int32_t *gid_s,
                    const float *l_d, int64_t N, int W, float *image, gcp_stream_t stream) {
    if (N < 0) return GCP_ERR_INVALID_ARG;
    if (N == 0) return GCP_OK;
    k_splat_color<<<blocks_for(N, 256 * CH), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(incl, x_s, key_s, gid_s,
                                                                                                 l_d, N, W, image);
    return static_cast<int>(cudaGetLastError());
}

int gcp_splat_bwd_w(const float *incl, const float *x_s, const int32_t *key_s, const int32_t *gid_s,
                    const float *l_d, const float *grad_image, int64_t N, int W, float *gshift,
                    gcp_stream_t stream) {
    if (N < 0) return GCP_ERR_INVALID_ARG;
    if (N == 0) return GCP_OK;
    k_splat_bwd_w<<<blocks_for(N, 256 * 4, 148 * 16), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        incl, x_s, key_s, gid_s, l_d, grad_image, N, W, gshift);
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
                       const int32_t *gid_s, const int32_t *sp, const int32_t *ep, const int64_t *goff,
                       const float *l_d, const float *grad_image, int64_t N, int W, float *elem,
                       gcp_stream_t stream) {
    if (N < 0) return GCP_ERR_INVALID_ARG;
    if (N == 0) return GCP_OK;
    k_splat_bwd_elem<<<blocks_for(N, 256 * 4, 148 * 16), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        incl, x_s, tu, key_s, gid_s, sp, ep, goff, l_d, grad_image, N, W, reinterpret_cast<float2 *>(elem));
    return static_cast<int>(cudaGetLastError());
}

int gcp_splat_bwd_reduce(const float *elem, const int32_t *sp, const int32_t *ep, const int64_t *goff,
                         const float *mean, const float *lam, const float *opac, const float *l_d, int64_t n,
                         float *g_mean, float *g_lam, float *g_opac, float *g_l, gcp_stream_t stream) {
    if (n < 0) return GCP_ERR_INVALID_ARG;
    if (n == 0) return GCP_OK;
    k_splat_bwd_reduce<<<blocks_for(n, 32, 148 * 16), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const float2 *>(elem), sp, ep, goff, mean, lam, opac, l_d, n, g_mean, g_lam, g_opac, g_l);
    return static_cast<int>(cudaGetLastError());
}

}  // extern "C"
