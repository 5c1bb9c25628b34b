/* ORACLE (test infrastructure only) -- plain-C restatement of torchvision's legacy ROIAlign.
 *
 * The algorithm lives in the un-vendored third-party dependency torchvision (reference pin
 * `torchvision>=0.6.0`, code/requirements.txt:2; installed 0.26.0), whose compiled op
 * torch.ops.torchvision.roi_align is reached from the reference at code/helpers/model.py:346 via
 * torchvision/models/detection/roi_heads.py:772,815 -> torchvision/ops/poolers.py:204-210
 * (spatial_scale per level, sampling_ratio=2, aligned=False) and from maskrcnn_loss at
 * torchvision/models/detection/roi_heads.py:85-97 (output 28x28, spatial_scale 1, sampling_ratio=-1).
 * The C++ source is not on disk; this file restates the published algorithm:
 *   start = x1*s (no -0.5 offset: aligned=False), w = max(x2*s - x1*s, 1), bin = w/P,
 *   grid = sampling_ratio > 0 ? sampling_ratio : ceil(roi_size / P), samples at
 *   start + p*bin + (i+0.5)*bin/grid; sample outside [-1, size] contributes 0; clamp to >=0;
 *   lo = (int)y; if lo >= size-1 -> lo=hi=size-1, y=lo; bilinear; output = mean over grid_h*grid_w.
 * tests/test_oracle.py pins it against torch.ops.torchvision.roi_align (forward and backward).
 *
 * Layout: input NCHW, rois [K,5] = (batch_idx, x1, y1, x2, y2), output [K,C,PH,PW]. T = float or double.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#define DEFINE_ROI_ALIGN(T, SUFFIX)                                                                   \
static void bilinear_##SUFFIX(int H, int W, T y, T x, int *ylo, int *xlo, int *yhi, int *xhi,        \
                              T *w1, T *w2, T *w3, T *w4, int *valid) {                               \
    if (y < (T)-1.0 || y > (T)H || x < (T)-1.0 || x > (T)W) { *valid = 0; return; }                   \
    *valid = 1;                                                                                       \
    if (y <= 0) y = 0;                                                                                \
    if (x <= 0) x = 0;                                                                                \
    int yl = (int)y, xl = (int)x, yh, xh;                                                             \
    if (yl >= H - 1) { yh = yl = H - 1; y = (T)yl; } else { yh = yl + 1; }                            \
    if (xl >= W - 1) { xh = xl = W - 1; x = (T)xl; } else { xh = xl + 1; }                            \
    T ly = y - yl, lx = x - xl, hy = (T)1.0 - ly, hx = (T)1.0 - lx;                                   \
    *ylo = yl; *xlo = xl; *yhi = yh; *xhi = xh;                                                       \
    *w1 = hy * hx; *w2 = hy * lx; *w3 = ly * hx; *w4 = ly * lx;                                       \
}                                                                                                     \
void roi_align_forward_##SUFFIX(const T *input, const T *rois, T *output, int64_t N, int64_t C,       \
                                int64_t H, int64_t W, int64_t K, int PH, int PW, T spatial_scale,     \
                                int sampling_ratio) {                                                 \
    (void)N;                                                                                          \
    for (int64_t k = 0; k < K; ++k) {                                                                 \
        const T *r = rois + k * 5;                                                                    \
        int64_t b = (int64_t)r[0];                                                                    \
        T sw = r[1] * spatial_scale, sh = r[2] * spatial_scale;                                       \
        T ew = r[3] * spatial_scale, eh = r[4] * spatial_scale;                                       \
        T rw = ew - sw, rh = eh - sh;                                                                 \
        if (rw < (T)1.0) rw = (T)1.0;                                                                 \
        if (rh < (T)1.0) rh = (T)1.0;                                                                 \
        T bh = rh / (T)PH, bw = rw / (T)PW;                                                           \
        int gh = sampling_ratio > 0 ? sampling_ratio : (int)ceil((double)(rh / (T)PH));              \
        int gw = sampling_ratio > 0 ? sampling_ratio : (int)ceil((double)(rw / (T)PW));              \
        T count = (T)(gh * gw > 1 ? gh * gw : 1);                                                     \
        for (int64_t c = 0; c < C; ++c) {                                                             \
            const T *plane = input + (b * C + c) * H * W;                                             \
            for (int ph = 0; ph < PH; ++ph) for (int pw = 0; pw < PW; ++pw) {                         \
                T acc = 0;                                                                            \
                for (int iy = 0; iy < gh; ++iy) {                                                     \
                    T y = sh + ph * bh + ((T)iy + (T)0.5) * bh / (T)gh;                               \
                    for (int ix = 0; ix < gw; ++ix) {                                                 \
                        T x = sw + pw * bw + ((T)ix + (T)0.5) * bw / (T)gw;                           \
                        int yl, xl, yh, xh, valid; T w1, w2, w3, w4;                                  \
                        bilinear_##SUFFIX((int)H, (int)W, y, x, &yl, &xl, &yh, &xh,                   \
                                          &w1, &w2, &w3, &w4, &valid);                                \
                        if (!valid) continue;                                                         \
                        acc += w1 * plane[yl * W + xl] + w2 * plane[yl * W + xh] +                    \
                               w3 * plane[yh * W + xl] + w4 * plane[yh * W + xh];                     \
                    }                                                                                 \
                }                                                                                     \
                output[((k * C + c) * PH + ph) * PW + pw] = acc / count;                              \
            }                                                                                         \
        }                                                                                             \
    }                                                                                                 \
}                                                                                                     \
void roi_align_backward_##SUFFIX(const T *grad_out, const T *rois, T *grad_in, int64_t N, int64_t C,  \
                                 int64_t H, int64_t W, int64_t K, int PH, int PW, T spatial_scale,    \
                                 int sampling_ratio) {                                                \
    memset(grad_in, 0, sizeof(T) * (size_t)(N * C * H * W));                                          \
    for (int64_t k = 0; k < K; ++k) {                                                                 \
        const T *r = rois + k * 5;                                                                    \
        int64_t b = (int64_t)r[0];                                                                    \
        T sw = r[1] * spatial_scale, sh = r[2] * spatial_scale;                                       \
        T ew = r[3] * spatial_scale, eh = r[4] * spatial_scale;                                       \
        T rw = ew - sw, rh = eh - sh;                                                                 \
        if (rw < (T)1.0) rw = (T)1.0;                                                                 \
        if (rh < (T)1.0) rh = (T)1.0;                                                                 \
        T bh = rh / (T)PH, bw = rw / (T)PW;                                                           \
        int gh = sampling_ratio > 0 ? sampling_ratio : (int)ceil((double)(rh / (T)PH));              \
        int gw = sampling_ratio > 0 ? sampling_ratio : (int)ceil((double)(rw / (T)PW));              \
        T count = (T)(gh * gw > 1 ? gh * gw : 1);                                                     \
        for (int64_t c = 0; c < C; ++c) {                                                             \
            T *plane = grad_in + (b * C + c) * H * W;                                                 \
            for (int ph = 0; ph < PH; ++ph) for (int pw = 0; pw < PW; ++pw) {                         \
                T g = grad_out[((k * C + c) * PH + ph) * PW + pw];                                    \
                for (int iy = 0; iy < gh; ++iy) {                                                     \
                    T y = sh + ph * bh + ((T)iy + (T)0.5) * bh / (T)gh;                               \
                    for (int ix = 0; ix < gw; ++ix) {                                                 \
                        T x = sw + pw * bw + ((T)ix + (T)0.5) * bw / (T)gw;                           \
                        int yl, xl, yh, xh, valid; T w1, w2, w3, w4;                                  \
                        bilinear_##SUFFIX((int)H, (int)W, y, x, &yl, &xl, &yh, &xh,                   \
                                          &w1, &w2, &w3, &w4, &valid);                                \
                        if (!valid) continue;                                                         \
                        plane[yl * W + xl] += g * w1 / count;                                         \
                        plane[yl * W + xh] += g * w2 / count;                                         \
                        plane[yh * W + xl] += g * w3 / count;                                         \
                        plane[yh * W + xh] += g * w4 / count;                                         \
                    }                                                                                 \
                }                                                                                     \
            }                                                                                         \
        }                                                                                             \
    }                                                                                                 \
}

DEFINE_ROI_ALIGN(float, f32)
DEFINE_ROI_ALIGN(double, f64)

/* FPN level of a box, torchvision/ops/poolers.py:73-84: floor(4 + log2(sqrt(area)/224) + 1e-6) clamped to
 * [k_min,k_max], minus k_min; evaluated in fp32 like the reference's tensor ops. */
void level_mapper_f32(const float *boxes, int64_t K, int k_min, int k_max, int64_t *levels) {
    for (int64_t k = 0; k < K; ++k) {
        const float *b = boxes + 4 * k;
        float area = (b[2] - b[0]) * (b[3] - b[1]);
        float s = sqrtf(area);
        float lv = floorf(4.0f + log2f(s / 224.0f) + 1e-6f);
        if (lv < (float)k_min) lv = (float)k_min;
        if (lv > (float)k_max) lv = (float)k_max;
        levels[k] = (int64_t)lv - k_min;
    }
}
