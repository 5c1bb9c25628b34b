// Multi-level ROIAlign (legacy aligned=False) on channels-last features, forward and backward, one launch for all
// FPN levels and all clips.  Semantics follow SURVEY.md 8(a) R2 (torch.ops.torchvision.roi_align as called from
// TV/ops/poolers.py:204-210 with sampling_ratio=2): start = x1*s, size = max(x2*s - x1*s, 1), samples at
// start + p*bin + (i+0.5)*bin/grid, a sample with y<-1 || y>H || x<-1 || x>W contributes 0, clamp to >=0,
// lo=(int)y, lo>=H-1 -> lo=hi=H-1 and y=lo, bilinear, mean over the grid.
//
// Mapping: one warp per output bin (roi, ph, pw); the bin geometry is computed once per warp (lane-uniform) and
// each lane owns 8 consecutive channels, so every bilinear tap is one contiguous row read of C*elem bytes
// (512 B for bf16, 1 KB for f32 at C=256) with 16-byte vector loads.  Backward scatters with 16-byte vector
// atomics (red.global.add.v4.f32) into f32 gradient maps.
#include "common.cuh"

namespace {

struct RoiArgs {
    const void* feat[4];
    float* dfeat[4];
    int H[4], W[4];
    float scale[4];
    int n_levels, feat_bf16;
    int N, C;
    long long cstride;
    const float* rois;
    const int* levels;
    long long K;
    int P, sr;
    void* out;
    int out_bf16, out_nchw;
};

struct Tap {
    int off[4];      // pixel offsets (y*W+x) of the 4 neighbours
    float w[4];
    bool valid;
};

__device__ __forceinline__ Tap make_tap(float y, float x, int H, int W) {
    Tap t;
    t.valid = !(y < -1.0f || y > (float)H || x < -1.0f || x > (float)W);
    if (!t.valid) return t;
    if (y <= 0.f) y = 0.f;
    if (x <= 0.f) x = 0.f;
    int yl = (int)y, xl = (int)x, yh, xh;
    if (yl >= H - 1) { yh = yl = H - 1; y = (float)yl; } else { yh = yl + 1; }
    if (xl >= W - 1) { xh = xl = W - 1; x = (float)xl; } else { xh = xl + 1; }
    const float ly = y - yl, lx = x - xl, hy = 1.f - ly, hx = 1.f - lx;
    t.off[0] = yl * W + xl; t.off[1] = yl * W + xh; t.off[2] = yh * W + xl; t.off[3] = yh * W + xh;
    t.w[0] = hy * hx; t.w[1] = hy * lx; t.w[2] = ly * hx; t.w[3] = ly * lx;
    return t;
}

__device__ __forceinline__ void ld8(const float* p, float (&v)[8]) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(p));
    const float4 b = __ldg(reinterpret_cast<const float4*>(p + 4));
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void ld8(const __nv_bfloat16* p, float (&v)[8]) {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) { v[2 * i] = __uint_as_float(w[i] << 16); v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u); }
}

struct BinGeom { int lvl, b, H, W; float y0, x0, bh, bw; };

__device__ __forceinline__ BinGeom bin_geom(const RoiArgs& a, long long k, int ph, int pw) {
    BinGeom g;
    const float* r = a.rois + k * 5;
    g.lvl = a.n_levels > 1 ? a.levels[k] : 0;
    g.b = (int)r[0];
    const float s = a.scale[g.lvl];
    g.H = a.H[g.lvl]; g.W = a.W[g.lvl];
    const float sw = r[1] * s, sh = r[2] * s, ew = r[3] * s, eh = r[4] * s;
    const float rw = fmaxf(ew - sw, 1.0f), rh = fmaxf(eh - sh, 1.0f);
    g.bh = rh / (float)a.P; g.bw = rw / (float)a.P;
    g.y0 = sh + ph * g.bh; g.x0 = sw + pw * g.bw;
    return g;
}

template <typename FT>
__global__ void __launch_bounds__(256) roi_align_fwd_kernel(const RoiArgs a) {
    const int lane = threadIdx.x & 31;
    const long long bin = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long long nbins = a.K * a.P * a.P;
    if (bin >= nbins) return;
    const int pw = (int)(bin % a.P), ph = (int)((bin / a.P) % a.P);
    const long long k = bin / ((long long)a.P * a.P);
    const BinGeom g = bin_geom(a, k, ph, pw);
    const FT* base = reinterpret_cast<const FT*>(a.feat[g.lvl]) + (long long)g.b * g.H * g.W * a.cstride;
    const float inv_count = 1.0f / (float)(a.sr * a.sr);
    for (int c8 = lane; c8 < a.C / 8; c8 += 32) {
        float acc[8] = {};
        for (int iy = 0; iy < a.sr; ++iy) {
            const float y = g.y0 + ((float)iy + 0.5f) * g.bh / (float)a.sr;
            for (int ix = 0; ix < a.sr; ++ix) {
                const float x = g.x0 + ((float)ix + 0.5f) * g.bw / (float)a.sr;
                const Tap t = make_tap(y, x, g.H, g.W);
                if (!t.valid) continue;
#pragma unroll
                for (int n = 0; n < 4; ++n) {
                    float v[8];
                    ld8(base + (long long)t.off[n] * a.cstride + c8 * 8, v);
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc[j] = fmaf(t.w[n], v[j], acc[j]);
                }
            }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] *= inv_count;
        if (!a.out_nchw) {
            const long long o = bin * a.C + c8 * 8;
            if (a.out_bf16) {
                uint4 u;
                u.x = pack_bf16x2(acc[0], acc[1]); u.y = pack_bf16x2(acc[2], acc[3]);
                u.z = pack_bf16x2(acc[4], acc[5]); u.w = pack_bf16x2(acc[6], acc[7]);
                *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(a.out) + o) = u;
            } else {
                float* d = reinterpret_cast<float*>(a.out) + o;
                *reinterpret_cast<float4*>(d) = make_float4(acc[0], acc[1], acc[2], acc[3]);
                *reinterpret_cast<float4*>(d + 4) = make_float4(acc[4], acc[5], acc[6], acc[7]);
            }
        } else {
            const long long pp = (long long)a.P * a.P;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const long long o = (k * a.C + c8 * 8 + j) * pp + ph * a.P + pw;
                if (a.out_bf16) reinterpret_cast<__nv_bfloat16*>(a.out)[o] = __float2bfloat16(acc[j]);
                else reinterpret_cast<float*>(a.out)[o] = acc[j];
            }
        }
    }
}

__global__ void __launch_bounds__(256) roi_align_bwd_kernel(const RoiArgs a) {
    const int lane = threadIdx.x & 31;
    const long long bin = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long long nbins = a.K * a.P * a.P;
    if (bin >= nbins) return;
    const int pw = (int)(bin % a.P), ph = (int)((bin / a.P) % a.P);
    const long long k = bin / ((long long)a.P * a.P);
    const BinGeom g = bin_geom(a, k, ph, pw);
    float* base = a.dfeat[g.lvl] + (long long)g.b * g.H * g.W * a.cstride;
    const float inv_count = 1.0f / (float)(a.sr * a.sr);
    const long long pp = (long long)a.P * a.P;
    for (int c8 = lane; c8 < a.C / 8; c8 += 32) {
        float go[8];
        if (!a.out_nchw) {
            if (a.out_bf16) ld8(reinterpret_cast<const __nv_bfloat16*>(a.out) + bin * a.C + c8 * 8, go);
            else ld8(reinterpret_cast<const float*>(a.out) + bin * a.C + c8 * 8, go);
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const long long o = (k * a.C + c8 * 8 + j) * pp + ph * a.P + pw;
                go[j] = a.out_bf16 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(a.out)[o])
                                   : reinterpret_cast<const float*>(a.out)[o];
            }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) go[j] *= inv_count;
        for (int iy = 0; iy < a.sr; ++iy) {
            const float y = g.y0 + ((float)iy + 0.5f) * g.bh / (float)a.sr;
            for (int ix = 0; ix < a.sr; ++ix) {
                const float x = g.x0 + ((float)ix + 0.5f) * g.bw / (float)a.sr;
                const Tap t = make_tap(y, x, g.H, g.W);
                if (!t.valid) continue;
#pragma unroll
                for (int n = 0; n < 4; ++n) {
                    float* d = base + (long long)t.off[n] * a.cstride + c8 * 8;
                    const float w = t.w[n];
                    atomicAdd(reinterpret_cast<float4*>(d), make_float4(w * go[0], w * go[1], w * go[2], w * go[3]));
                    atomicAdd(reinterpret_cast<float4*>(d + 4), make_float4(w * go[4], w * go[5], w * go[6], w * go[7]));
                }
            }
        }
    }
}

__global__ void roi_levels_kernel(const float* rois, long long K, int k_min, int k_max, int* levels) {
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= K) return;
    const float* r = rois + k * 5;
    const float area = (r[3] - r[1]) * (r[4] - r[2]);
    const float s = sqrtf(area);
    float lv = floorf(4.0f + log2f(s / 224.0f) + 1e-6f);
    lv = fminf(fmaxf(lv, (float)k_min), (float)k_max);
    levels[k] = (int)lv - k_min;
}

// project_masks_on_boxes: single-channel u8 image, spatial_scale 1, adaptive grid = ceil(roi_size / M)
__global__ void mask_targets_kernel(const uint8_t* __restrict__ masks, int n_obj, int H, int W, const float* __restrict__ rois,
                                    long long K, int M, float* out) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= K * M * M) return;
    const int pw = (int)(idx % M), ph = (int)((idx / M) % M);
    const long long k = idx / ((long long)M * M);
    const float* r = rois + k * 5;
    const int obj = (int)r[0];
    const float sw = r[1], sh = r[2];
    const float rw = fmaxf(r[3] - r[1], 1.0f), rh = fmaxf(r[4] - r[2], 1.0f);
    const float bh = rh / (float)M, bw = rw / (float)M;
    const int gh = (int)ceilf(rh / (float)M), gw = (int)ceilf(rw / (float)M);
    const float count = fmaxf((float)(gh * gw), 1.0f);
    const uint8_t* img = masks + (long long)obj * H * W;
    float acc = 0.f;
    if (obj >= 0 && obj < n_obj)
        for (int iy = 0; iy < gh; ++iy) {
            const float y = sh + ph * bh + ((float)iy + 0.5f) * bh / (float)gh;
            for (int ix = 0; ix < gw; ++ix) {
                const float x = sw + pw * bw + ((float)ix + 0.5f) * bw / (float)gw;
                const Tap t = make_tap(y, x, H, W);
                if (!t.valid) continue;
                acc += t.w[0] * (float)img[t.off[0]] + t.w[1] * (float)img[t.off[1]] +
                       t.w[2] * (float)img[t.off[2]] + t.w[3] * (float)img[t.off[3]];
            }
        }
    out[idx] = acc / count;
}

int fill_args(const sfvos_roi_params* p, RoiArgs* a, bool bwd) {
    SF_CHECK(p != nullptr, "roi_align: null params");
    SF_CHECK(p->n_levels >= 1 && p->n_levels <= 4, "roi_align: n_levels must be 1..4");
    SF_CHECK(p->C % 8 == 0 && p->cstride % 8 == 0, "roi_align: C and cstride must be multiples of 8");
    SF_CHECK(p->sampling_ratio > 0, "roi_align: sampling_ratio must be positive");
    SF_CHECK(p->P > 0, "roi_align: bad output size");
    SF_CHECK(p->n_levels == 1 || p->levels != nullptr, "roi_align: levels required for multi-level pooling");
    for (int i = 0; i < 4; ++i) {
        a->feat[i] = i < p->n_levels ? p->feat[i] : nullptr;
        a->dfeat[i] = i < p->n_levels ? reinterpret_cast<float*>(p->dfeat[i]) : nullptr;
        a->H[i] = (int)p->H[i]; a->W[i] = (int)p->W[i]; a->scale[i] = p->scale[i];
        if (i < p->n_levels) SF_CHECK(bwd ? a->dfeat[i] != nullptr : a->feat[i] != nullptr, "roi_align: level %d buffer is NULL", i);
    }
    a->n_levels = p->n_levels; a->feat_bf16 = (p->feat_dtype == SFVOS_BF16);
    a->N = (int)p->N; a->C = (int)p->C; a->cstride = p->cstride;
    a->rois = p->rois; a->levels = p->levels; a->K = p->K; a->P = p->P; a->sr = p->sampling_ratio;
    a->out = p->out; a->out_bf16 = (p->out_dtype == SFVOS_BF16); a->out_nchw = p->out_nchw;
    return SFVOS_OK;
}

}  // namespace

extern "C" int sfvos_roi_levels(const float* rois, int64_t K, int32_t k_min, int32_t k_max, int32_t* levels,
                                sfvos_stream stream) {
    if (K == 0) return SFVOS_OK;
    roi_levels_kernel<<<(int)((K + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(rois, K, k_min, k_max, levels);
    SF_LAUNCH_CHECK();
    return SFVOS_OK;
}

extern "C" int sfvos_roi_align_fwd(const sfvos_roi_params* p, sfvos_stream stream) {
    RoiArgs a;
    int rc = fill_args(p, &a, false);
    if (rc) return rc;
    if (a.K == 0) return SFVOS_OK;
    const long long nbins = a.K * a.P * a.P;
    const int grid = (int)((nbins + 7) / 8);
    if (a.feat_bf16) roi_align_fwd_kernel<__nv_bfloat16><<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(a);
    else roi_align_fwd_kernel<float><<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(a);
    SF_LAUNCH_CHECK();
    return SFVOS_OK;
}

extern "C" int sfvos_roi_align_bwd(const sfvos_roi_params* p, sfvos_stream stream) {
    RoiArgs a;
    int rc = fill_args(p, &a, true);
    if (rc) return rc;
    if (a.K == 0) return SFVOS_OK;
    const long long nbins = a.K * a.P * a.P;
    const int grid = (int)((nbins + 7) / 8);
    roi_align_bwd_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(a);
    SF_LAUNCH_CHECK();
    return SFVOS_OK;
}

extern "C" int sfvos_mask_targets(const uint8_t* masks, int64_t n_obj, int64_t H, int64_t W, const float* rois, int64_t K,
                                  int32_t M, float* out, sfvos_stream stream) {
    if (K == 0) return SFVOS_OK;
    const long long total = K * M * M;
    mask_targets_kernel<<<(int)((total + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(masks, (int)n_obj, (int)H, (int)W, rois, K, M, out);
    SF_LAUNCH_CHECK();
    return SFVOS_OK;
}
