// Mask paste-back: torchvision's paste_masks_in_image (TV/models/detection/roi_heads.py:415-501), reached from
// code/helpers/model.py:347 via GeneralizedRCNNTransform.postprocess (TV/models/detection/transform.py:257-279).
// Per detection: the M x M mask probabilities are zero-padded by `padding`, the box is grown by (M + 2 padding) / M
// about its centre and truncated to integers, the padded mask is resized to the box's (h, w) with bilinear
// interpolation (align_corners = False) and written into an im_h x im_w canvas of zeros, clipped to the image.
// The reference path runs ~10 tiny kernels per mask in a Python loop; here every output pixel of every mask of a whole
// chunk of frames is one thread of ONE launch: a pure streaming write of K * im_h * im_w floats (HBM-bound), the
// (M + 2)^2 source mask stays in L1/L2.  Arithmetic follows the fp32 op sequence of expand_boxes and of ATen's
// upsample_bilinear2d (area_pixel_compute_source_index); the box arithmetic uses explicit round-to-nearest intrinsics so
// no FMA contraction changes an integer truncation.
#include "common.cuh"

namespace {

struct PasteBox { int x0, y0, w, h; };

__device__ __forceinline__ PasteBox expand_box(const float* b, float scale) {
    // expand_boxes: half extents and centre in fp32, half extents scaled, then .to(int64) = truncation toward zero
    const float w_half = __fmul_rn(__fmul_rn(__fsub_rn(b[2], b[0]), 0.5f), scale);
    const float h_half = __fmul_rn(__fmul_rn(__fsub_rn(b[3], b[1]), 0.5f), scale);
    const float x_c = __fmul_rn(__fadd_rn(b[2], b[0]), 0.5f);
    const float y_c = __fmul_rn(__fadd_rn(b[3], b[1]), 0.5f);
    const long long x1 = (long long)__fsub_rn(x_c, w_half), x2 = (long long)__fadd_rn(x_c, w_half);
    const long long y1 = (long long)__fsub_rn(y_c, h_half), y2 = (long long)__fadd_rn(y_c, h_half);
    PasteBox p;
    p.x0 = (int)x1; p.y0 = (int)y1;
    long long w = x2 - x1 + 1, h = y2 - y1 + 1;
    p.w = (int)(w < 1 ? 1 : w); p.h = (int)(h < 1 ? 1 : h);
    return p;
}

// source coordinate of output index i (ATen area_pixel_compute_source_index, align_corners = False)
__device__ __forceinline__ void src_coord(int i, int in_size, int out_size, int* i0, int* step, float* l1) {
    const float scale = (float)in_size / (float)out_size;
    float s = fmaf(scale, (float)i + 0.5f, -0.5f);       // ATen's CUDA kernel: scale * (i + 0.5) - 0.5, contracted by nvcc
    if (s < 0.f) s = 0.f;
    const int lo = (int)s;
    *i0 = lo;
    *step = (lo < in_size - 1) ? 1 : 0;
    *l1 = s - (float)lo;
}

__global__ void __launch_bounds__(256)
paste_masks_kernel(const float* __restrict__ masks, const float* __restrict__ boxes, int M, int padding, float scale,
                   int im_h, int im_w, float* __restrict__ out) {
    const int k = blockIdx.z, y = blockIdx.y;
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= im_w) return;
    const PasteBox pb = expand_box(boxes + 4 * (long long)k, scale);
    const int iy = y - pb.y0, ix = x - pb.x0;
    float v = 0.f;
    if (iy >= 0 && iy < pb.h && ix >= 0 && ix < pb.w) {
        const int S = M + 2 * padding;
        int h1, hp, w1, wp;
        float hl1, wl1;
        src_coord(iy, S, pb.h, &h1, &hp, &hl1);
        src_coord(ix, S, pb.w, &w1, &wp, &wl1);
        const float hl0 = 1.f - hl1, wl0 = 1.f - wl1;
        const float* m = masks + (long long)k * M * M;
        auto at = [&](int r, int c) -> float {          // padded mask: zeros in the `padding` border
            r -= padding; c -= padding;
            return (r >= 0 && r < M && c >= 0 && c < M) ? __ldg(m + r * M + c) : 0.f;
        };
        v = hl0 * (wl0 * at(h1, w1) + wl1 * at(h1, w1 + wp)) + hl1 * (wl0 * at(h1 + hp, w1) + wl1 * at(h1 + hp, w1 + wp));
    }
    out[((long long)k * im_h + y) * im_w + x] = v;
}

}  // namespace

extern "C" int sfvos_paste_masks(const float* masks, const float* boxes, int64_t K, int32_t M, int32_t padding, int64_t im_h,
                                 int64_t im_w, float* out, sfvos_stream stream) {
    SF_CHECK(M > 0 && padding >= 0 && im_h > 0 && im_w > 0, "paste_masks: bad sizes (M=%d, padding=%d, image %lld x %lld)", M, padding,
             (long long)im_h, (long long)im_w);
    SF_CHECK(K <= 65535 && im_h <= 65535, "paste_masks: at most 65535 masks / rows per call");
    int rc = sfvos_device_check();
    if (rc) return rc;
    if (K == 0) return SFVOS_OK;
    const float scale = (float)((double)(M + 2 * padding) / (double)M);
    dim3 grid((unsigned)((im_w + 255) / 256), (unsigned)im_h, (unsigned)K);
    paste_masks_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(masks, boxes, M, padding, scale, (int)im_h, (int)im_w, out);
    SF_LAUNCH_CHECK();
    return SFVOS_OK;
}
