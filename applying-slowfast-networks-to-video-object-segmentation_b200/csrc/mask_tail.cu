// Tail of the mask branch: mask_fcn_logits (1x1 conv C->n_cls, TV/models/detection/mask_rcnn.py:344), the
// class-channel BCE-with-logits loss of maskrcnn_loss (TV/models/detection/roi_heads.py:100-129), their fused
// backward, and the sigmoid/select of maskrcnn_inference (roi_heads.py:56-82).
// N = n_cls (2) is far too narrow for a tensor-core tile, so these are warp-per-pixel dot products over the
// contiguous channel axis (16-byte vector loads, shuffle reductions): HBM-bound on reading x once.
#include <stdlib.h>

#include "common.cuh"

namespace {

__device__ __forceinline__ void ld8(const float* p, float (&v)[8]) {
    const float4 a = *reinterpret_cast<const float4*>(p);
    const float4 b = *reinterpret_cast<const float4*>(p + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void ld8(const __nv_bfloat16* p, float (&v)[8]) {
    const uint4 u = *reinterpret_cast<const uint4*>(p);
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) { v[2 * i] = __uint_as_float(w[i] << 16); v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u); }
}
__device__ __forceinline__ void st8(float* p, const float (&v)[8]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ void st8(__nv_bfloat16* p, const float (&v)[8]) {
    uint4 u;
    u.x = pack_bf16x2(v[0], v[1]); u.y = pack_bf16x2(v[2], v[3]); u.z = pack_bf16x2(v[4], v[5]); u.w = pack_bf16x2(v[6], v[7]);
    *reinterpret_cast<uint4*>(p) = u;
}

constexpr int MAX_CLS = 8;

// pixel_order 1: the rows of x are the 2x2 space-to-depth order of the S x S map -- row (h*(S/2) + w)*4 + 2i + j holds spatial
// pixel (2h+i, 2w+j) -- which is how the ConvTranspose2d(k2,s2) of the mask predictor leaves its output when its four taps are
// computed as ONE 1x1 convolution with 4*C output channels.  logits / glogits stay in the reference's spatial order.
__device__ __forceinline__ long long row_to_spatial(long long r, int S, int pixel_order) {
    if (!pixel_order) return r;
    const int tap = (int)(r & 3);
    const int lw = (int)(r >> 2), hs = S >> 1;
    const int h = lw / hs, w = lw - h * hs;
    return (long long)(2 * h + (tap >> 1)) * S + 2 * w + (tap & 1);
}

template <typename XT>
__global__ void __launch_bounds__(256)
mask_logits_fwd_kernel(const XT* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b, float* logits,
                       long long K, int S, int C, int n_cls, int pixel_order) {
    const int lane = threadIdx.x & 31;
    const long long pix = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long long ss = (long long)S * S;
    if (pix >= K * ss) return;
    float acc[MAX_CLS] = {};
    for (int c8 = lane; c8 < C / 8; c8 += 32) {
        float v[8];
        ld8(x + pix * C + c8 * 8, v);
        for (int cls = 0; cls < n_cls; ++cls) {
            float wv[8];
            ld8(w + (long long)cls * C + c8 * 8, wv);
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[cls] = fmaf(v[j], wv[j], acc[cls]);
        }
    }
    const long long k = pix / ss, p = pix - k * ss;
    for (int cls = 0; cls < n_cls; ++cls) {
        const float s = warp_sum(acc[cls]);
        if (lane == 0) logits[(k * n_cls + cls) * ss + row_to_spatial(p, S, pixel_order)] = s + b[cls];
    }
}

// n_cls == 2, C == 256 (the reference's mask predictor): a warp takes 8 consecutive pixels -- 8 independent 512-byte row
// loads in flight per lane, weights held in registers -- and reduces the 16 partial dot products with one 15-shuffle
// transpose-reduce instead of 10 shuffles per pixel.
template <typename XT>
__global__ void __launch_bounds__(256)
mask_logits_fwd2_kernel(const XT* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b, float* logits,
                        long long npix, long long ss, int S, int pixel_order) {
    constexpr int C = 256;
    const int lane = threadIdx.x & 31;
    const long long pix0 = ((long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 8;
    if (pix0 >= npix) return;
    float w0[8], w1[8];
    ld8(w + lane * 8, w0);
    ld8(w + C + lane * 8, w1);
    float v[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        if (pix0 + i < npix) ld8(x + (pix0 + i) * C + lane * 8, v[i]);
        else {
#pragma unroll
            for (int j = 0; j < 8; ++j) v[i][j] = 0.f;
        }
    }
    float s[16];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        float a0 = 0.f, a1 = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) { a0 = fmaf(v[i][j], w0[j], a0); a1 = fmaf(v[i][j], w1[j], a1); }
        s[2 * i] = a0; s[2 * i + 1] = a1;
    }
#pragma unroll
    for (int st = 0; st < 4; ++st) {
        const int m = 16 >> st, n = 8 >> st;
        const bool upper = (lane & m) != 0;
#pragma unroll
        for (int t = 0; t < n; ++t) {
            const float send = upper ? s[t] : s[t + n];
            const float keep = upper ? s[t + n] : s[t];
            s[t] = keep + __shfl_xor_sync(0xffffffffu, send, m);
        }
    }
    s[0] += __shfl_xor_sync(0xffffffffu, s[0], 1);
    if ((lane & 1) == 0) {                       // lane pair l holds value index l >> 1 = 2 * pixel + class
        const int idx = lane >> 1, i = idx >> 1, cls = idx & 1;
        const long long pix = pix0 + i;
        if (pix < npix) {
            const long long k = pix / ss, q = pix - k * ss;
            logits[(k * 2 + cls) * ss + row_to_spatial(q, S, pixel_order)] = s[0] + b[cls];
        }
    }
}

__global__ void __launch_bounds__(256)
mask_bce_fwd_kernel(const float* __restrict__ logits, const long long* __restrict__ labels, const float* __restrict__ targets,
                    float* loss, long long K, int S, int n_cls) {
    __shared__ float part[8];
    const long long ss = (long long)S * S, total = K * ss;
    float acc = 0.f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long k = i / ss, p = i - k * ss;
        const float z = logits[(k * n_cls + labels[k]) * ss + p];
        const float t = targets[i];
        acc += fmaxf(z, 0.f) - z * t + log1pf(expf(-fabsf(z)));
    }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f;
        for (int i = 0; i < (blockDim.x >> 5); ++i) s += part[i];
        atomicAdd(loss, s / (float)total);
    }
}

// one CTA per ROI (8 warps stride over its S*S pixels): dx = sum_cls g_cls * w[cls], dw[cls] += g_cls * x, db[cls] += g_cls
// RELU: x is the output of a ReLU (the ConvTranspose2d + ReLU of the mask predictor) and the kernel also applies that ReLU's
// backward: dx = (x > 0) ? sum_cls g_cls * w[cls] : 0 and dbias_x[c] += sum_pixels dx -- the separate relu_bwd pass (read dx,
// read x, write dx again: 1.2 GB at the bench size) disappears.
template <typename XT, typename DxT, bool RELU>
__global__ void __launch_bounds__(256)
mask_logits_bwd_kernel(const XT* __restrict__ x, const float* __restrict__ w, const float* __restrict__ glogits, DxT* dx,
                       float* dw, float* db, float* dbias_x, long long K, int S, int C, int n_cls, int pixel_order) {
    extern __shared__ float s_dw[];          // [8 warps][n_cls][C] (+ [8 warps][C] for RELU)
    __shared__ float s_db[8][MAX_CLS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long k = blockIdx.x;
    const int ss = S * S;
    float dbacc[MAX_CLS] = {};
    float* s_dbx = s_dw + 8 * n_cls * C;
    for (int c8 = lane; c8 < C / 8; c8 += 32) {
        float dbx[8] = {};
        for (int cls0 = 0; cls0 < n_cls; cls0 += 2) {            // two classes per pass keeps registers bounded
            const int ncl = min(2, n_cls - cls0);
            const bool last_pass = cls0 + 2 >= n_cls;
            float wv[2][8], dwacc[2][8] = {};
            for (int q = 0; q < ncl; ++q) ld8(w + (long long)(cls0 + q) * C + c8 * 8, wv[q]);
            float vn[8], gn[2] = {0.f, 0.f};                      // next pixel's activations / logit gradients, loaded one trip ahead
            if (warp < ss) {
                ld8(x + (k * ss + warp) * C + c8 * 8, vn);
                for (int q = 0; q < ncl; ++q) gn[q] = glogits[(k * n_cls + cls0 + q) * ss + row_to_spatial(warp, S, pixel_order)];
            }
            for (int p = warp; p < ss; p += 8) {
                float v[8], o[8];
                const float gc[2] = {gn[0], gn[1]};
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] = vn[j];
                if (p + 8 < ss) {
                    ld8(x + (k * ss + p + 8) * C + c8 * 8, vn);
                    for (int q = 0; q < ncl; ++q) gn[q] = glogits[(k * n_cls + cls0 + q) * ss + row_to_spatial(p + 8, S, pixel_order)];
                }
                if (cls0 == 0) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) o[j] = 0.f;
                } else {
                    ld8(dx + (k * ss + p) * C + c8 * 8, o);
                }
                for (int q = 0; q < ncl; ++q) {
                    const float g = gc[q];
#pragma unroll
                    for (int j = 0; j < 8; ++j) { o[j] = fmaf(g, wv[q][j], o[j]); dwacc[q][j] = fmaf(g, v[j], dwacc[q][j]); }
                    if (c8 == 0) dbacc[cls0 + q] += g;
                }
                if (RELU && last_pass) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        o[j] = v[j] > 0.f ? o[j] : 0.f;
                        dbx[j] += o[j];
                    }
                }
                st8(dx + (k * ss + p) * C + c8 * 8, o);
            }
            for (int q = 0; q < ncl; ++q)
#pragma unroll
                for (int j = 0; j < 8; ++j) s_dw[(warp * n_cls + cls0 + q) * C + c8 * 8 + j] = dwacc[q][j];
        }
        if (RELU) {
#pragma unroll
            for (int j = 0; j < 8; ++j) s_dbx[warp * C + c8 * 8 + j] = dbx[j];
        }
    }
    if (lane == 0)
        for (int q = 0; q < n_cls; ++q) s_db[warp][q] = dbacc[q];
    __syncthreads();
    for (int i = threadIdx.x; i < n_cls * C; i += blockDim.x) {
        float s = 0.f;
        for (int wi = 0; wi < 8; ++wi) s += s_dw[wi * n_cls * C + i];
        atomicAdd(dw + i, s);
    }
    if (threadIdx.x < n_cls) {
        float s = 0.f;
        for (int wi = 0; wi < 8; ++wi) s += s_db[wi][threadIdx.x];
        atomicAdd(db + threadIdx.x, s);
    }
    if (RELU) {
        for (int i = threadIdx.x; i < C; i += blockDim.x) {
            float s = 0.f;
            for (int wi = 0; wi < 8; ++wi) s += s_dbx[wi * C + i];
            atomicAdd(dbias_x + i, s);
        }
    }
}

// The product path's instance of the kernel above (bf16, C = 256: one 8-channel group per lane, RELU, <= 2 classes), with the
// memory-level parallelism the general kernel lacks.  ncu / timeline, round 2: the general kernel ran at 2.1 TB/s (33 % of the
// HBM roofline): 119 registers -> 2 CTAs = 16 warps per SM, each with ONE 512-byte row in flight (+ one prefetched), i.e. a
// latency chain of 98 trips per warp.  Here a warp loads FOUR pixels' rows and logit gradients per trip before it touches
// any of them, and 3 CTAs fit an SM.  Same pixel -> warp assignment and accumulation order as the general kernel, so the
// results are bit-identical to it.
template <int NCLS>
__global__ void __launch_bounds__(256, 3)
mask_logits_relu_bwd_c256_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ w, const float* __restrict__ glogits,
                                 __nv_bfloat16* dx, float* dw, float* db, float* dbias_x, int S, int pixel_order) {
    constexpr int C = 256, UNR = 4;
    extern __shared__ float s_dw[];          // [8 warps][NCLS][C] + [8 warps][C]
    __shared__ float s_db[8][NCLS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long k = blockIdx.x;
    const int ss = S * S;
    float* s_dbx = s_dw + 8 * NCLS * C;
    const __nv_bfloat16* xk = x + k * ss * C + lane * 8;
    __nv_bfloat16* dxk = dx + k * ss * C + lane * 8;
    const float* gk = glogits + k * NCLS * ss;
    float wv[NCLS][8], dwacc[NCLS][8], dbx[8], dbacc[NCLS];
#pragma unroll
    for (int q = 0; q < NCLS; ++q) {
        ld8(w + q * C + lane * 8, wv[q]);
        dbacc[q] = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) dwacc[q][j] = 0.f;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) dbx[j] = 0.f;
    const int n_it = warp < ss ? (ss - warp + 7) / 8 : 0;             // this warp's pixels: warp, warp + 8, ...
    for (int i0 = 0; i0 < n_it; i0 += UNR) {
        uint4 xv[UNR];
        float g[UNR][NCLS];
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
            const int p = warp + 8 * (i0 + u);
            xv[u] = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
            for (int q = 0; q < NCLS; ++q) g[u][q] = 0.f;
            if (i0 + u < n_it) {
                xv[u] = __ldg(reinterpret_cast<const uint4*>(xk + (long long)p * C));
                const long long sp = row_to_spatial(p, S, pixel_order);
#pragma unroll
                for (int q = 0; q < NCLS; ++q) g[u][q] = __ldg(gk + (long long)q * ss + sp);
            }
        }
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
            if (i0 + u < n_it) {
                const int p = warp + 8 * (i0 + u);
                const uint32_t xw[4] = {xv[u].x, xv[u].y, xv[u].z, xv[u].w};
                float v[8], o[8];
#pragma unroll
                for (int i = 0; i < 4; ++i) { v[2 * i] = __uint_as_float(xw[i] << 16); v[2 * i + 1] = __uint_as_float(xw[i] & 0xffff0000u); }
#pragma unroll
                for (int j = 0; j < 8; ++j) o[j] = 0.f;
#pragma unroll
                for (int q = 0; q < NCLS; ++q) {
                    const float gq = g[u][q];
#pragma unroll
                    for (int j = 0; j < 8; ++j) { o[j] = fmaf(gq, wv[q][j], o[j]); dwacc[q][j] = fmaf(gq, v[j], dwacc[q][j]); }
                    if (lane == 0) dbacc[q] += gq;
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    o[j] = v[j] > 0.f ? o[j] : 0.f;
                    dbx[j] += o[j];
                }
                st8(dxk + (long long)p * C, o);
            }
        }
    }
#pragma unroll
    for (int q = 0; q < NCLS; ++q)
#pragma unroll
        for (int j = 0; j < 8; ++j) s_dw[(warp * NCLS + q) * C + lane * 8 + j] = dwacc[q][j];
#pragma unroll
    for (int j = 0; j < 8; ++j) s_dbx[warp * C + lane * 8 + j] = dbx[j];
    if (lane == 0)
        for (int q = 0; q < NCLS; ++q) s_db[warp][q] = dbacc[q];
    __syncthreads();
    for (int i = threadIdx.x; i < NCLS * C; i += blockDim.x) {
        float sacc = 0.f;
        for (int wi = 0; wi < 8; ++wi) sacc += s_dw[wi * NCLS * C + i];
        atomicAdd(dw + i, sacc);
    }
    if (threadIdx.x < NCLS) {
        float sacc = 0.f;
        for (int wi = 0; wi < 8; ++wi) sacc += s_db[wi][threadIdx.x];
        atomicAdd(db + threadIdx.x, sacc);
    }
    for (int i = threadIdx.x; i < C; i += blockDim.x) {
        float sacc = 0.f;
        for (int wi = 0; wi < 8; ++wi) sacc += s_dbx[wi * C + i];
        atomicAdd(dbias_x + i, sacc);
    }
}

// glogits[k, cls, p] = (cls == labels[k]) ? gloss * (sigmoid(z) - t) / (K*S*S) : 0
__global__ void mask_bce_bwd_kernel(const float* __restrict__ logits, const long long* __restrict__ labels,
                                    const float* __restrict__ targets, const float* __restrict__ gloss, float* glogits,
                                    long long K, int S, int n_cls) {
    const long long ss = (long long)S * S, total = K * n_cls * ss;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const long long p = i % ss, cls = (i / ss) % n_cls, k = i / (ss * n_cls);
    float g = 0.f;
    if (cls == labels[k]) g = gloss[0] / (float)(K * ss) * (1.f / (1.f + expf(-logits[i])) - targets[k * ss + p]);
    glogits[i] = g;
}

__global__ void mask_probs_kernel(const float* __restrict__ logits, const long long* __restrict__ labels, float* prob,
                                  long long K, int S, int n_cls) {
    const long long ss = (long long)S * S, total = K * ss;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const long long k = i / ss, p = i - k * ss;
    prob[i] = 1.f / (1.f + expf(-logits[(k * n_cls + labels[k]) * ss + p]));
}

}  // namespace

#define CS(s) reinterpret_cast<cudaStream_t>(s)

extern "C" int sfvos_mask_logits_fwd(const void* x, int32_t x_dtype, const float* w, const float* b, float* logits,
                                     int64_t K, int64_t S, int64_t C, int32_t n_cls, int32_t pixel_order, sfvos_stream stream) {
    SF_CHECK(pixel_order == 0 || (pixel_order == 1 && S % 2 == 0), "mask_logits: pixel_order 1 needs an even S");
    SF_CHECK(C % 8 == 0 && n_cls >= 1 && n_cls <= MAX_CLS, "mask_logits: C %% 8 == 0 and n_cls <= 8 required");
    if (K == 0) return SFVOS_OK;
    const long long npix = K * S * S;
    const int grid = (int)((npix + 7) / 8);
    if (n_cls == 2 && C == 256) {
        const int grid2 = (int)((npix + 63) / 64);
        if (x_dtype == SFVOS_BF16) mask_logits_fwd2_kernel<__nv_bfloat16><<<grid2, 256, 0, CS(stream)>>>(reinterpret_cast<const __nv_bfloat16*>(x), w, b, logits, npix, S * S, (int)S, pixel_order);
        else mask_logits_fwd2_kernel<float><<<grid2, 256, 0, CS(stream)>>>(reinterpret_cast<const float*>(x), w, b, logits, npix, S * S, (int)S, pixel_order);
        SF_LAUNCH_CHECK();
        return SFVOS_OK;
    }
    if (x_dtype == SFVOS_BF16) mask_logits_fwd_kernel<__nv_bfloat16><<<grid, 256, 0, CS(stream)>>>(reinterpret_cast<const __nv_bfloat16*>(x), w, b, logits, K, (int)S, (int)C, n_cls, pixel_order);
    else mask_logits_fwd_kernel<float><<<grid, 256, 0, CS(stream)>>>(reinterpret_cast<const float*>(x), w, b, logits, K, (int)S, (int)C, n_cls, pixel_order);
    SF_LAUNCH_CHECK();
    return SFVOS_OK;
}

extern "C" int sfvos_mask_bce_fwd(const float* logits, const int64_t* labels, const float* targets, float* loss, int64_t K,
                                  int64_t S, int32_t n_cls, sfvos_stream stream) {
    SF_CUDA(cudaMemsetAsync(loss, 0, sizeof(float), CS(stream)));
    if (K == 0) return SFVOS_OK;
    const long long total = K * S * S;
    long long grid = (total + 255) / 256;
    if (grid > 2 * sfvos_num_sms()) grid = 2 * sfvos_num_sms();
    mask_bce_fwd_kernel<<<(int)grid, 256, 0, CS(stream)>>>(logits, reinterpret_cast<const long long*>(labels), targets, loss, K, (int)S, n_cls);
    SF_LAUNCH_CHECK();
    return SFVOS_OK;
}

extern "C" int sfvos_mask_logits_bwd(const void* x, int32_t x_dtype, const float* w, const float* glogits, void* dx,
                                     int32_t dx_dtype, float* dw, float* db, int64_t K, int64_t S, int64_t C,
                                     int32_t n_cls, int32_t pixel_order, sfvos_stream stream) {
    SF_CHECK(pixel_order == 0 || (pixel_order == 1 && S % 2 == 0), "mask_logits_bwd: pixel_order 1 needs an even S");
    SF_CHECK(C % 8 == 0 && n_cls >= 1 && n_cls <= MAX_CLS, "mask_logits_bwd: C %% 8 == 0 and n_cls <= 8 required");
    SF_CHECK(x_dtype == dx_dtype, "mask_logits_bwd: x and dx must share a dtype");
    if (K == 0) return SFVOS_OK;
    const size_t sm = (size_t)8 * n_cls * C * sizeof(float);
    SF_CHECK(sm <= 48 * 1024, "mask_logits_bwd: n_cls*C too large");
    if (x_dtype == SFVOS_BF16)
        mask_logits_bwd_kernel<__nv_bfloat16, __nv_bfloat16, false><<<(int)K, 256, sm, CS(stream)>>>(reinterpret_cast<const __nv_bfloat16*>(x), w, glogits, reinterpret_cast<__nv_bfloat16*>(dx), dw, db, nullptr, K, (int)S, (int)C, n_cls, pixel_order);
    else
        mask_logits_bwd_kernel<float, float, false><<<(int)K, 256, sm, CS(stream)>>>(reinterpret_cast<const float*>(x), w, glogits, reinterpret_cast<float*>(dx), dw, db, nullptr, K, (int)S, (int)C, n_cls, pixel_order);
    SF_LAUNCH_CHECK();
    return SFVOS_OK;
}

extern "C" int sfvos_mask_logits_relu_bwd(const void* x, int32_t x_dtype, const float* w, const float* glogits, void* dx,
                                          int32_t dx_dtype, float* dw, float* db, float* dbias_x, int64_t K, int64_t S,
                                          int64_t C, int32_t n_cls, int32_t pixel_order, sfvos_stream stream) {
    SF_CHECK(pixel_order == 0 || (pixel_order == 1 && S % 2 == 0), "mask_logits_relu_bwd: pixel_order 1 needs an even S");
    SF_CHECK(C % 8 == 0 && n_cls >= 1 && n_cls <= MAX_CLS, "mask_logits_relu_bwd: C %% 8 == 0 and n_cls <= 8 required");
    SF_CHECK(x_dtype == dx_dtype, "mask_logits_relu_bwd: x and dx must share a dtype");
    SF_CHECK(dbias_x != nullptr, "mask_logits_relu_bwd: dbias_x required");
    if (K == 0) return SFVOS_OK;
    const size_t sm = (size_t)8 * (n_cls + 1) * C * sizeof(float);
    SF_CHECK(sm <= 48 * 1024, "mask_logits_relu_bwd: n_cls*C too large");
    if (x_dtype == SFVOS_BF16 && C == 256 && n_cls <= 2 && getenv("SFVOS_MASK_LOGITS_GENERIC") == nullptr) {
        using bf = __nv_bfloat16;
        if (n_cls == 2)
            mask_logits_relu_bwd_c256_kernel<2><<<(int)K, 256, sm, CS(stream)>>>(reinterpret_cast<const bf*>(x), w, glogits, reinterpret_cast<bf*>(dx), dw, db, dbias_x, (int)S, pixel_order);
        else
            mask_logits_relu_bwd_c256_kernel<1><<<(int)K, 256, sm, CS(stream)>>>(reinterpret_cast<const bf*>(x), w, glogits, reinterpret_cast<bf*>(dx), dw, db, dbias_x, (int)S, pixel_order);
        SF_LAUNCH_CHECK();
        return SFVOS_OK;
    }
    if (x_dtype == SFVOS_BF16)
        mask_logits_bwd_kernel<__nv_bfloat16, __nv_bfloat16, true><<<(int)K, 256, sm, CS(stream)>>>(reinterpret_cast<const __nv_bfloat16*>(x), w, glogits, reinterpret_cast<__nv_bfloat16*>(dx), dw, db, dbias_x, K, (int)S, (int)C, n_cls, pixel_order);
    else
        mask_logits_bwd_kernel<float, float, true><<<(int)K, 256, sm, CS(stream)>>>(reinterpret_cast<const float*>(x), w, glogits, reinterpret_cast<float*>(dx), dw, db, dbias_x, K, (int)S, (int)C, n_cls, pixel_order);
    SF_LAUNCH_CHECK();
    return SFVOS_OK;
}

extern "C" int sfvos_mask_bce_bwd(const float* logits, const int64_t* labels, const float* targets, const float* gloss,
                                  float* glogits, int64_t K, int64_t S, int32_t n_cls, sfvos_stream stream) {
    if (K == 0) return SFVOS_OK;
    const long long total = K * n_cls * S * S;
    mask_bce_bwd_kernel<<<(int)((total + 255) / 256), 256, 0, CS(stream)>>>(logits, reinterpret_cast<const long long*>(labels), targets, gloss, glogits, K, (int)S, n_cls);
    SF_LAUNCH_CHECK();
    return SFVOS_OK;
}

extern "C" int sfvos_mask_probs(const float* logits, const int64_t* labels, float* prob, int64_t K, int64_t S, int32_t n_cls,
                                sfvos_stream stream) {
    if (K == 0) return SFVOS_OK;
    const long long total = K * S * S;
    mask_probs_kernel<<<(int)((total + 255) / 256), 256, 0, CS(stream)>>>(logits, reinterpret_cast<const long long*>(labels), prob, K, (int)S, n_cls);
    SF_LAUNCH_CHECK();
    return SFVOS_OK;
}
