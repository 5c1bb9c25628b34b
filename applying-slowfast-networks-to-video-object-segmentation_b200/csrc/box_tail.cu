// Detection losses of the box branch: torchvision's fastrcnn_loss (TV/models/detection/roi_heads.py:12-53), reached
// from code/helpers/model.py:346 via RoIHeads.forward (TV roi_heads.py:783).
//   loss_classifier = mean_m( logsumexp(z_m) - z_m[label_m] )                              (F.cross_entropy)
//   loss_box_reg    = sum_{m: label_m > 0} sum_j smooth_l1(r_m[label_m, j] - t_m[j]; beta) / M   (beta = 1/9, "sum" / numel)
// One ROI per thread.  The forward pass is a single CTA (M is a few thousand ROIs) accumulating in double, so the two
// scalars are deterministic and need no pre-zeroed output; the backward pass is element-wise and writes the gradient
// of every class logit / box delta (zeros where the loss does not look), so the caller needs no fill either.
#include "common.cuh"

namespace {

constexpr int LOSS_THREADS = 1024;

__device__ __forceinline__ float row_lse(const float* z, int n_cls, float* zmax_out) {
    float zmax = z[0];
    for (int c = 1; c < n_cls; ++c) zmax = fmaxf(zmax, z[c]);
    float s = 0.f;
    for (int c = 0; c < n_cls; ++c) s += expf(z[c] - zmax);
    *zmax_out = zmax;
    return s;
}

__global__ void __launch_bounds__(LOSS_THREADS)
fastrcnn_loss_fwd_kernel(const float* __restrict__ cls, long long cls_stride, const float* __restrict__ box, long long box_stride,
                         const long long* __restrict__ labels, const float* __restrict__ tgt, long long M, int n_cls,
                         float beta, float* losses) {
    __shared__ double red[2][LOSS_THREADS / 32];
    double ce = 0.0, sl = 0.0;
    for (long long m = threadIdx.x; m < M; m += LOSS_THREADS) {
        const float* z = cls + m * cls_stride;
        const int lab = (int)labels[m];
        float zmax;
        const float s = row_lse(z, n_cls, &zmax);
        ce += (double)(logf(s) + zmax - z[lab]);
        if (lab > 0) {
            const float* r = box + m * box_stride + 4 * lab;
            const float* t = tgt + m * 4;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float d = r[j] - t[j];
                const float ad = fabsf(d);
                sl += (double)(ad < beta ? 0.5f * d * d / beta : ad - 0.5f * beta);
            }
        }
    }
    for (int o = 16; o > 0; o >>= 1) {
        ce += __shfl_down_sync(0xffffffffu, ce, o);
        sl += __shfl_down_sync(0xffffffffu, sl, o);
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { red[0][warp] = ce; red[1][warp] = sl; }
    __syncthreads();
    if (warp == 0) {
        ce = red[0][lane]; sl = red[1][lane];
        for (int o = 16; o > 0; o >>= 1) {
            ce += __shfl_down_sync(0xffffffffu, ce, o);
            sl += __shfl_down_sync(0xffffffffu, sl, o);
        }
        if (lane == 0) {
            losses[0] = (float)(ce / (double)M);
            losses[1] = (float)(sl / (double)M);
        }
    }
}

__global__ void __launch_bounds__(256)
fastrcnn_loss_bwd_kernel(const float* __restrict__ cls, long long cls_stride, const float* __restrict__ box, long long box_stride,
                         const long long* __restrict__ labels, const float* __restrict__ tgt, const float* __restrict__ gloss,
                         long long M, int n_cls, float beta, float* dcls, long long dcls_stride, float* dbox,
                         long long dbox_stride) {
    const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M) return;
    const float g_cls = gloss[0] / (float)M, g_box = gloss[1] / (float)M;
    const float* z = cls + m * cls_stride;
    const int lab = (int)labels[m];
    float zmax;
    const float s = row_lse(z, n_cls, &zmax);
    const float inv = 1.0f / s;
    float* dz = dcls + m * dcls_stride;
    for (int c = 0; c < n_cls; ++c) dz[c] = g_cls * (expf(z[c] - zmax) * inv - (c == lab ? 1.0f : 0.0f));
    float* dr = dbox + m * dbox_stride;
    const float* r = box + m * box_stride;
    const float* t = tgt + m * 4;
    for (int c = 0; c < n_cls; ++c)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float g = 0.f;
            if (c == lab && lab > 0) {
                const float d = r[4 * c + j] - t[j];
                g = g_box * (fabsf(d) < beta ? d / beta : (d > 0.f ? 1.0f : -1.0f));
            }
            dr[4 * c + j] = g;
        }
}

}  // namespace

extern "C" int sfvos_fastrcnn_loss_fwd(const float* cls_logits, int64_t cls_stride, const float* box_reg, int64_t box_stride,
                                       const int64_t* labels, const float* reg_targets, int64_t M, int32_t n_cls, float beta,
                                       float* losses, sfvos_stream stream) {
    SF_CHECK(M > 0 && n_cls >= 2, "fastrcnn_loss_fwd: needs M > 0 ROIs and >= 2 classes (M=%lld, n_cls=%d)", (long long)M, n_cls);
    SF_CHECK(beta > 0.f, "fastrcnn_loss_fwd: beta must be positive");
    int rc = sfvos_device_check();
    if (rc) return rc;
    fastrcnn_loss_fwd_kernel<<<1, LOSS_THREADS, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        cls_logits, cls_stride, box_reg, box_stride, reinterpret_cast<const long long*>(labels), reg_targets, M, n_cls, beta, losses);
    SF_LAUNCH_CHECK();
    return SFVOS_OK;
}

extern "C" int sfvos_fastrcnn_loss_bwd(const float* cls_logits, int64_t cls_stride, const float* box_reg, int64_t box_stride,
                                       const int64_t* labels, const float* reg_targets, const float* gloss, int64_t M,
                                       int32_t n_cls, float beta, float* dcls, int64_t dcls_stride, float* dbox,
                                       int64_t dbox_stride, sfvos_stream stream) {
    SF_CHECK(M > 0 && n_cls >= 2, "fastrcnn_loss_bwd: needs M > 0 ROIs and >= 2 classes (M=%lld, n_cls=%d)", (long long)M, n_cls);
    SF_CHECK(beta > 0.f, "fastrcnn_loss_bwd: beta must be positive");
    int rc = sfvos_device_check();
    if (rc) return rc;
    fastrcnn_loss_bwd_kernel<<<(unsigned)((M + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        cls_logits, cls_stride, box_reg, box_stride, reinterpret_cast<const long long*>(labels), reg_targets, gloss, M, n_cls, beta,
        dcls, dcls_stride, dbox, dbox_stride);
    SF_LAUNCH_CHECK();
    return SFVOS_OK;
}
