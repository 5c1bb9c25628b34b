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
#include <stdlib.h>

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


// ---- sampling_ratio == 2 fast path -------------------------------------------------------------------------------
// The bilinear weights of a sample factor into a row part and a column part, and so does the sum over the 2 x 2
// sampling grid of a bin:  out = 1/4 * sum_r sum_c Wy[r] * Wx[c] * F[r][c]  with Wy / Wx the per-axis weights of the
// two samples merged by feature row / column.  Neighbouring samples are usually < 2 cells apart, so the 4 x 4 = 16
// taps of the reference loop collapse to ~3 x 3 distinct cells: fewer L2 reads forward, fewer atomics backward.
// (Same rule set per sample: -1 / L range test, clamp at 0, lo >= L-1 -> lo = hi = L-1.)
struct Axis {
    int idx[4];
    float w[4];
    int n;
};

__device__ __forceinline__ void axis_add(Axis& a, int i, float w) {
    bool found = false;
#pragma unroll
    for (int e = 0; e < 4; ++e)
        if (e < a.n && a.idx[e] == i) { a.w[e] += w; found = true; }
    if (!found) {
#pragma unroll
        for (int e = 0; e < 4; ++e)
            if (e == a.n) { a.idx[e] = i; a.w[e] = w; }
        a.n++;
    }
}

__device__ __forceinline__ Axis make_axis2(float start, float bin, int L) {
    Axis a;
    a.n = 0;
#pragma unroll
    for (int e = 0; e < 4; ++e) { a.idx[e] = 0; a.w[e] = 0.f; }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        float v = start + ((float)i + 0.5f) * bin / 2.0f;
        if (v < -1.0f || v > (float)L) continue;
        if (v <= 0.f) v = 0.f;
        int lo = (int)v, hi;
        if (lo >= L - 1) { hi = lo = L - 1; v = (float)lo; } else { hi = lo + 1; }
        const float l = v - lo, h = 1.f - l;
        axis_add(a, lo, h);
        axis_add(a, hi, l);
    }
    return a;
}

__device__ __forceinline__ float4 ld4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float4 ld4(const __nv_bfloat16* p) {
    const uint2 u = __ldg(reinterpret_cast<const uint2*>(p));
    return make_float4(__uint_as_float(u.x << 16), __uint_as_float(u.x & 0xffff0000u), __uint_as_float(u.y << 16),
                       __uint_as_float(u.y & 0xffff0000u));
}

// One CTA per ROI.  The merged taps depend only on (roi, ph) for rows and (roi, pw) for columns, so the first 2P
// threads build the 2P axis records once into shared memory; the 8 warps then walk the P*P bins (one warp per bin, lane l
// owns channels [c0 + 4l, c0 + 4l + 4) of every 128-channel chunk c0: each tap is one fully coalesced 512 B (f32) /
// 256 B (bf16) warp access) with no geometry arithmetic on their critical path.
constexpr int MAX_P = 16;

struct AxisRec {
    int off[4];         // rows: feature row * W; columns: feature column
    float w[4];         // 0 for unused entries (never loaded / never added)
};

struct RoiPlan {
    AxisRec y[MAX_P], x[MAX_P];
    int lvl, b, H, W;
};

__device__ __forceinline__ void build_plan(const RoiArgs& a, long long k, RoiPlan* plan) {
    const int t = threadIdx.x;
    if (t < 2 * a.P) {
        const bool is_x = t >= a.P;
        const int i = is_x ? t - a.P : t;
        const BinGeom g = bin_geom(a, k, is_x ? 0 : i, is_x ? i : 0);
        const Axis ax = is_x ? make_axis2(g.x0, g.bw, g.W) : make_axis2(g.y0, g.bh, g.H);
        AxisRec r;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            r.off[e] = e < ax.n ? (is_x ? ax.idx[e] : ax.idx[e] * g.W) : 0;
            r.w[e] = e < ax.n ? ax.w[e] : 0.f;
        }
        if (is_x) plan->x[i] = r; else plan->y[i] = r;
        if (t == 0) { plan->lvl = g.lvl; plan->b = g.b; plan->H = g.H; plan->W = g.W; }
    }
}

// Pooled value of one bin for channels [c, c+4): the distinct cells (<= 16, typically 9) are fetched with independent
// predicated loads, all in flight together, ahead of the FMAs.
template <typename FT>
__device__ __forceinline__ float4 pool_bin4(const FT* base, const AxisRec& ry, const AxisRec& rx, long long cstride, int c) {
    float4 v[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int q = 0; q < 4; ++q)
            v[r][q] = (ry.w[r] != 0.f && rx.w[q] != 0.f) ? ld4(base + (long long)(ry.off[r] + rx.off[q]) * cstride + c)
                                                         : make_float4(0.f, 0.f, 0.f, 0.f);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float w = ry.w[r] * rx.w[q];
            acc.x = fmaf(w, v[r][q].x, acc.x); acc.y = fmaf(w, v[r][q].y, acc.y);
            acc.z = fmaf(w, v[r][q].z, acc.z); acc.w = fmaf(w, v[r][q].w, acc.w);
        }
    acc.x *= 0.25f; acc.y *= 0.25f; acc.z *= 0.25f; acc.w *= 0.25f;
    return acc;
}

// NCHW = false: output [K,P,P,C] written straight from registers (coalesced rows).
// NCHW = true : output [K,C,P,P] (what the torch box head flattens) staged as a [C][P*P] tile in shared memory and
//               written as one contiguous, fully coalesced C*P*P block.
template <typename FT, bool NCHW>
__global__ void __launch_bounds__(256, 2) roi_align_fwd2_kernel(const RoiArgs a) {
    extern __shared__ float s_tile[];                     // NCHW only: [C][P*P]
    __shared__ RoiPlan plan;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    const long long k = blockIdx.x;
    const int pp = a.P * a.P;
    build_plan(a, k, &plan);
    __syncthreads();
    const FT* base = reinterpret_cast<const FT*>(a.feat[plan.lvl]) + (long long)plan.b * plan.H * plan.W * a.cstride;
    for (int b = warp; b < pp; b += nwarp) {
        const int ph = b / a.P, pw = b - ph * a.P;
        const AxisRec ry = plan.y[ph], rx = plan.x[pw];
        for (int c = lane * 4; c < a.C; c += 128) {
            const float4 acc = pool_bin4(base, ry, rx, a.cstride, c);
            if (NCHW) {
                s_tile[(c + 0) * pp + b] = acc.x; s_tile[(c + 1) * pp + b] = acc.y;
                s_tile[(c + 2) * pp + b] = acc.z; s_tile[(c + 3) * pp + b] = acc.w;
            } else {
                const long long o = (k * pp + b) * a.C + c;
                if (a.out_bf16) {
                    uint2 u;
                    u.x = pack_bf16x2(acc.x, acc.y); u.y = pack_bf16x2(acc.z, acc.w);
                    *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(a.out) + o) = u;
                } else {
                    *reinterpret_cast<float4*>(reinterpret_cast<float*>(a.out) + o) = acc;
                }
            }
        }
    }
    if (NCHW) {
        __syncthreads();
        const long long o = k * a.C * pp;
        const int n = a.C * pp;
        if (a.out_bf16 && (n & 7) == 0) {
            // bf16 rows for the native box head: 8 elements (16 bytes) per store
            uint4* d = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(a.out) + o);
            for (int i = threadIdx.x; i < n / 8; i += blockDim.x) {
                const float* s = s_tile + 8 * i;
                uint4 u;
                u.x = pack_bf16x2(s[0], s[1]); u.y = pack_bf16x2(s[2], s[3]);
                u.z = pack_bf16x2(s[4], s[5]); u.w = pack_bf16x2(s[6], s[7]);
                d[i] = u;
            }
        } else {
            for (int i = threadIdx.x; i < n; i += blockDim.x) {
                if (a.out_bf16) reinterpret_cast<__nv_bfloat16*>(a.out)[o + i] = __float2bfloat16(s_tile[i]);
                else reinterpret_cast<float*>(a.out)[o + i] = s_tile[i];
            }
        }
    }
}

// Tap-list variant of the forward kernel.  ncu showed roi_align_fwd2 ISSUE-bound on the 14 x 14 mask pooling (61 % issue-slot
// utilisation at 3.7 warps per scheduler, ~330 instructions per bin and 128-channel chunk): every warp re-derived the 16
// (row, column) products, predicates and 64-bit addresses of a bin and ran all 64 FMAs including the zero-weight ones.  Here
// one thread per bin first compacts the bin's distinct cells into a list of (cell offset, weight) pairs in shared memory
// (typically 9 of 16), and the pooling loop is: broadcast-read a pair, one address multiply-add, one 16-byte load, 4 FMAs.
// Same cells, same weights, same accumulation order as roi_align_fwd2 (bit-identical results).
constexpr int MAX_TAPS = 16;

template <typename FT, bool NCHW>
__global__ void __launch_bounds__(256, NCHW ? 3 : 4) roi_align_fwd3_kernel(const RoiArgs a) {
    extern __shared__ float s_dyn[];                      // [NCHW: C * P*P floats] [P*P * 16 (offset, weight) pairs] [P*P counts]
    __shared__ RoiPlan plan;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    const long long k = blockIdx.x;
    const int pp = a.P * a.P;
    float* s_tile = s_dyn;
    uint2* s_taps = reinterpret_cast<uint2*>(s_dyn + (NCHW ? a.C * pp : 0));
    int* s_cnt = reinterpret_cast<int*>(s_taps + pp * MAX_TAPS);
    build_plan(a, k, &plan);
    __syncthreads();
    for (int b = threadIdx.x; b < pp; b += blockDim.x) {
        const int ph = b / a.P, pw = b - ph * a.P;
        const AxisRec ry = plan.y[ph], rx = plan.x[pw];
        int n = 0;
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int q = 0; q < 4; ++q)
                if (ry.w[r] != 0.f && rx.w[q] != 0.f)
                    s_taps[b * MAX_TAPS + n++] = make_uint2((unsigned)(ry.off[r] + rx.off[q]), __float_as_uint(ry.w[r] * rx.w[q]));
        s_cnt[b] = n;
        for (; n < MAX_TAPS; ++n) s_taps[b * MAX_TAPS + n] = make_uint2(0u, 0u);     // padding: cell 0, weight 0 (value discarded)
    }
    __syncthreads();
    const FT* base = reinterpret_cast<const FT*>(a.feat[plan.lvl]) + (long long)plan.b * plan.H * plan.W * a.cstride;
    for (int b = warp; b < pp; b += nwarp) {
        const int n = s_cnt[b];
        const uint2* tp = s_taps + b * MAX_TAPS;
        for (int c = lane * 4; c < a.C; c += 128) {
            // all of the bin's loads in flight before the first FMA.  The first 9 slots (the typical 3 x 3 neighbourhood) are
            // loaded unconditionally -- a padding slot reads cell 0 and its value is discarded below, so NaNs cannot leak in --
            // which lets the compiler issue them back to back; slots 9..15 (bins wider than 2 cells) are predicated.
            float4 v[9];
#pragma unroll
            for (int t = 0; t < 9; ++t) v[t] = ld4(base + (long long)tp[t].x * a.cstride + c);
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int t = 0; t < 9; ++t) {
                const float w = __uint_as_float(tp[t].y);
                const float4 x = (t < n) ? v[t] : make_float4(0.f, 0.f, 0.f, 0.f);
                acc.x = fmaf(w, x.x, acc.x); acc.y = fmaf(w, x.y, acc.y);
                acc.z = fmaf(w, x.z, acc.z); acc.w = fmaf(w, x.w, acc.w);
            }
            if (n > 9) {
                // bins wider than 2 cells per axis (rare): the remaining slots in a second round of loads, same accumulation
                // order.  Keeping only 9 vectors live (instead of all 16) is what lets three CTAs share an SM.
                float4 u[MAX_TAPS - 9];
#pragma unroll
                for (int t = 9; t < MAX_TAPS; ++t) {
                    u[t - 9] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (t < n) u[t - 9] = ld4(base + (long long)tp[t].x * a.cstride + c);
                }
#pragma unroll
                for (int t = 9; t < MAX_TAPS; ++t)
                    if (t < n) {
                        const float w = __uint_as_float(tp[t].y);
                        acc.x = fmaf(w, u[t - 9].x, acc.x); acc.y = fmaf(w, u[t - 9].y, acc.y);
                        acc.z = fmaf(w, u[t - 9].z, acc.z); acc.w = fmaf(w, u[t - 9].w, acc.w);
                    }
            }
            acc.x *= 0.25f; acc.y *= 0.25f; acc.z *= 0.25f; acc.w *= 0.25f;
            if (NCHW) {
                s_tile[(c + 0) * pp + b] = acc.x; s_tile[(c + 1) * pp + b] = acc.y;
                s_tile[(c + 2) * pp + b] = acc.z; s_tile[(c + 3) * pp + b] = acc.w;
            } else {
                const long long o = (k * pp + b) * a.C + c;
                if (a.out_bf16) {
                    uint2 u;
                    u.x = pack_bf16x2(acc.x, acc.y); u.y = pack_bf16x2(acc.z, acc.w);
                    *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(a.out) + o) = u;
                } else {
                    *reinterpret_cast<float4*>(reinterpret_cast<float*>(a.out) + o) = acc;
                }
            }
        }
    }
    if (NCHW) {
        __syncthreads();
        const long long o = k * a.C * pp;
        const int n = a.C * pp;
        if (a.out_bf16 && (n & 7) == 0) {
            uint4* d = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(a.out) + o);
            for (int i = threadIdx.x; i < n / 8; i += blockDim.x) {
                const float* sv = s_tile + 8 * i;
                uint4 u;
                u.x = pack_bf16x2(sv[0], sv[1]); u.y = pack_bf16x2(sv[2], sv[3]);
                u.z = pack_bf16x2(sv[4], sv[5]); u.w = pack_bf16x2(sv[6], sv[7]);
                d[i] = u;
            }
        } else {
            for (int i = threadIdx.x; i < n; i += blockDim.x) {
                if (a.out_bf16) reinterpret_cast<__nv_bfloat16*>(a.out)[o + i] = __float2bfloat16(s_tile[i]);
                else reinterpret_cast<float*>(a.out)[o + i] = s_tile[i];
            }
        }
    }
}

// Backward: one warp per bin (the atomics are fire-and-forget, so what counts is how many warps issue them); the
// merged taps cut the 16 vector atomics per bin and channel chunk of the reference loop to ~9.
__global__ void __launch_bounds__(256) roi_align_bwd2_kernel(const RoiArgs a) {
    const int lane = threadIdx.x & 31;
    const long long bin = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long long nbins = a.K * a.P * a.P;
    if (bin >= nbins) return;
    const int pw = (int)(bin % a.P), ph = (int)((bin / a.P) % a.P);
    const long long k = bin / ((long long)a.P * a.P);
    const BinGeom g = bin_geom(a, k, ph, pw);
    const Axis ay = make_axis2(g.y0, g.bh, g.H), ax = make_axis2(g.x0, g.bw, g.W);
    float* base = a.dfeat[g.lvl] + (long long)g.b * g.H * g.W * a.cstride;
    const long long pp = (long long)a.P * a.P;
    for (int c = lane * 4; c < a.C; c += 128) {
        float4 go;
        if (!a.out_nchw) {
            go = a.out_bf16 ? ld4(reinterpret_cast<const __nv_bfloat16*>(a.out) + bin * a.C + c)
                            : ld4(reinterpret_cast<const float*>(a.out) + bin * a.C + c);
        } else {
            float v4[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const long long o = (k * a.C + c + j) * pp + ph * a.P + pw;
                v4[j] = a.out_bf16 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(a.out)[o])
                                   : reinterpret_cast<const float*>(a.out)[o];
            }
            go = make_float4(v4[0], v4[1], v4[2], v4[3]);
        }
        go.x *= 0.25f; go.y *= 0.25f; go.z *= 0.25f; go.w *= 0.25f;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            if (r >= ay.n) break;
            float* row = base + (long long)ay.idx[r] * g.W * a.cstride + c;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                if (q >= ax.n) break;
                const float w = ay.w[r] * ax.w[q];
                atomicAdd(reinterpret_cast<float4*>(row + (long long)ax.idx[q] * a.cstride),
                          make_float4(w * go.x, w * go.y, w * go.z, w * go.w));
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

// project_masks_on_boxes: single-channel u8 image, spatial_scale 1, adaptive grid = ceil(roi_size / M).
// One WARP per output bin: a 700-pixel box has 25 x 25 samples per bin, so the lanes stride over the sampling grid
// (a thread per bin leaves the whole launch waiting on the few largest boxes).
__global__ void __launch_bounds__(256)
mask_targets_kernel(const uint8_t* __restrict__ masks, int n_obj, int H, int W, const float* __restrict__ rois,
                    long long K, int M, float* out) {
    const int lane = threadIdx.x & 31;
    const long long idx = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
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
        for (int s = lane; s < gh * gw; s += 32) {
            const int iy = s / gw, ix = s - iy * gw;
            const float y = sh + ph * bh + ((float)iy + 0.5f) * bh / (float)gh;
            const float x = sw + pw * bw + ((float)ix + 0.5f) * bw / (float)gw;
            const Tap t = make_tap(y, x, H, W);
            if (!t.valid) continue;
            acc += t.w[0] * (float)img[t.off[0]] + t.w[1] * (float)img[t.off[1]] +
                   t.w[2] * (float)img[t.off[2]] + t.w[3] * (float)img[t.off[3]];
        }
    acc = warp_sum(acc);
    if (lane == 0) out[idx] = acc / count;
}


// Separable formulation of project_masks_on_boxes: the sampling grid of a bin is dense (spacing <= 1 pixel), so the sum
// over its gh x gw bilinear samples is  sum_r sum_c Wy[ph][r] * Wx[pw][c] * img[r][c]  with per-axis weights accumulated
// once per ROI.  One CTA per ROI: (1) 2M threads build the row / column weight tables, (2) T[r][pw] = sum_c Wx[pw][c] *
// img[r][c] for every image row the ROI touches, (3) out[ph][pw] = sum_r Wy[ph][r] * T[r][pw] / count.  Every mask pixel is
// read ~once instead of once per sample and tap (4 gh gw byte loads per bin in the reference loop).
constexpr int MT_SPAN = 64;      // max image rows / columns under one bin (+2): boxes up to M * 62 pixels

struct MtAxis {
    float w[MT_SPAN];
    int base, len;
};

__global__ void __launch_bounds__(256)
mask_targets_sep_kernel(const uint8_t* __restrict__ masks, int n_obj, int H, int W, const float* __restrict__ rois, int M,
                        int max_rows, float* out) {
    extern __shared__ float mt_smem[];
    MtAxis* ax = reinterpret_cast<MtAxis*>(mt_smem);            // [2][M]: y tables then x tables
    float* T = reinterpret_cast<float*>(ax + 2 * M);              // [max_rows][M]
    const long long k = blockIdx.x;
    const float* r = rois + k * 5;
    const int obj = (int)r[0];
    const float sw = r[1], sh = r[2];
    const float rw = fmaxf(r[3] - r[1], 1.0f), rh = fmaxf(r[4] - r[2], 1.0f);
    const float bh = rh / (float)M, bw = rw / (float)M;
    const int gh = (int)ceilf(rh / (float)M), gw = (int)ceilf(rw / (float)M);
    const float inv_count = 1.0f / fmaxf((float)(gh * gw), 1.0f);
    float* o = out + k * M * M;
    if (obj < 0 || obj >= n_obj) {
        for (int i = threadIdx.x; i < M * M; i += blockDim.x) o[i] = 0.f;
        return;
    }
    if (threadIdx.x < 2 * M) {
        const bool is_x = threadIdx.x >= M;
        const int p = is_x ? threadIdx.x - M : threadIdx.x;
        MtAxis& a = ax[threadIdx.x];
        const float start = is_x ? sw + p * bw : sh + p * bh, bin = is_x ? bw : bh;
        const int g = is_x ? gw : gh, L = is_x ? W : H;
        for (int i = 0; i < MT_SPAN; ++i) a.w[i] = 0.f;
        int base = -1, last = -1;
        for (int i = 0; i < g; ++i) {
            float v = start + ((float)i + 0.5f) * bin / (float)g;
            if (v < -1.0f || v > (float)L) continue;
            if (v <= 0.f) v = 0.f;
            int lo = (int)v, hi;
            if (lo >= L - 1) { hi = lo = L - 1; v = (float)lo; } else { hi = lo + 1; }
            const float l = v - lo, h = 1.f - l;
            if (base < 0) base = lo;
            if (hi - base < MT_SPAN) { a.w[lo - base] += h; a.w[hi - base] += l; last = hi; }
        }
        a.base = base < 0 ? 0 : base;
        a.len = base < 0 ? 0 : last - base + 1;
    }
    __syncthreads();
    const MtAxis* ay = ax;
    const MtAxis* axx = ax + M;
    // rows touched by the ROI
    int rbase = 1 << 30, rend = 0;
    for (int p = 0; p < M; ++p)
        if (ay[p].len > 0) { rbase = min(rbase, ay[p].base); rend = max(rend, ay[p].base + ay[p].len); }
    int nrows = rend > rbase ? rend - rbase : 0;
    if (nrows > max_rows) nrows = max_rows;
    const uint8_t* img = masks + (long long)obj * H * W;
    for (int i = threadIdx.x; i < nrows * M; i += blockDim.x) {
        const int row = i / M, pw = i - row * M;
        const MtAxis& a = axx[pw];
        const uint8_t* src = img + (long long)(rbase + row) * W + a.base;
        float acc = 0.f;
        for (int c = 0; c < a.len; ++c) acc = fmaf(a.w[c], (float)src[c], acc);
        T[i] = acc;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < M * M; i += blockDim.x) {
        const int ph = i / M, pw = i - ph * M;
        const MtAxis& a = ay[ph];
        float acc = 0.f;
        for (int j = 0; j < a.len; ++j) {
            const int row = a.base - rbase + j;
            if (row < nrows) acc = fmaf(a.w[j], T[row * M + pw], acc);
        }
        o[i] = acc * inv_count;
    }
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
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const long long nbins = a.K * a.P * a.P;
    const size_t tile_bytes = a.out_nchw ? (size_t)a.C * a.P * a.P * sizeof(float) : 0;
    const bool fast = a.sr == 2 && a.C % 4 == 0 && a.P <= MAX_P && tile_bytes <= 160 * 1024 && a.K < (1LL << 31) &&
                      getenv("SFVOS_ROI_GENERIC") == nullptr;
    const char* tl = getenv("SFVOS_ROI_TAPLIST");
    const size_t list_bytes = (size_t)a.P * a.P * (MAX_TAPS * sizeof(uint2) + sizeof(int));
    if (fast && (tl == nullptr || atoi(tl) != 0) && tile_bytes + list_bytes <= 200 * 1024) {
        const size_t smem = tile_bytes + list_bytes;
#define SF_ROI_FWD3(FT, NCHW)                                                                                          \
    do {                                                                                                               \
        SF_CUDA(cudaFuncSetAttribute(roi_align_fwd3_kernel<FT, NCHW>, cudaFuncAttributeMaxDynamicSharedMemorySize,     \
                                     (int)smem));                                                                      \
        roi_align_fwd3_kernel<FT, NCHW><<<(int)a.K, 256, smem, st>>>(a);                                               \
    } while (0)
        if (a.feat_bf16 && a.out_nchw) SF_ROI_FWD3(__nv_bfloat16, true);
        else if (a.feat_bf16) SF_ROI_FWD3(__nv_bfloat16, false);
        else if (a.out_nchw) SF_ROI_FWD3(float, true);
        else SF_ROI_FWD3(float, false);
#undef SF_ROI_FWD3
    } else if (fast) {
#define SF_ROI_FWD(FT, NCHW)                                                                                           \
    do {                                                                                                               \
        SF_CUDA(cudaFuncSetAttribute(roi_align_fwd2_kernel<FT, NCHW>, cudaFuncAttributeMaxDynamicSharedMemorySize,     \
                                     (int)tile_bytes));                                                                \
        roi_align_fwd2_kernel<FT, NCHW><<<(int)a.K, 256, tile_bytes, st>>>(a);                                         \
    } while (0)
        if (a.feat_bf16 && a.out_nchw) SF_ROI_FWD(__nv_bfloat16, true);
        else if (a.feat_bf16) SF_ROI_FWD(__nv_bfloat16, false);
        else if (a.out_nchw) SF_ROI_FWD(float, true);
        else SF_ROI_FWD(float, false);
#undef SF_ROI_FWD
    } else {
        const int grid = (int)((nbins + 7) / 8);
        if (a.feat_bf16) roi_align_fwd_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(a);
        else roi_align_fwd_kernel<float><<<grid, 256, 0, st>>>(a);
    }
    SF_LAUNCH_CHECK();
    return SFVOS_OK;
}

extern "C" int sfvos_roi_align_bwd(const sfvos_roi_params* p, sfvos_stream stream) {
    RoiArgs a;
    int rc = fill_args(p, &a, true);
    if (rc) return rc;
    if (a.K == 0) return SFVOS_OK;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const long long nbins = a.K * a.P * a.P;
    const size_t tile_bytes = a.out_nchw ? (size_t)a.C * a.P * a.P * sizeof(float) : 0;
    const bool fast = a.sr == 2 && a.C % 4 == 0 && a.P <= MAX_P && tile_bytes <= 160 * 1024 && a.K < (1LL << 31) &&
                      getenv("SFVOS_ROI_GENERIC") == nullptr;
    if (fast) {
        roi_align_bwd2_kernel<<<(int)((nbins + 7) / 8), 256, 0, st>>>(a);
    } else {
        roi_align_bwd_kernel<<<(int)((nbins + 7) / 8), 256, 0, st>>>(a);
    }
    SF_LAUNCH_CHECK();
    return SFVOS_OK;
}

extern "C" int sfvos_mask_targets(const uint8_t* masks, int64_t n_obj, int64_t H, int64_t W, const float* rois, int64_t K,
                                  int32_t M, float* out, sfvos_stream stream) {
    if (K == 0) return SFVOS_OK;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    // separable kernel: a bin may cover at most MT_SPAN - 2 pixels per axis, and the ROI's row band must fit shared memory
    const int max_dim = (int)(H > W ? H : W);
    const int max_rows = (int)H + 2;
    const size_t smem = (size_t)2 * M * sizeof(MtAxis) + (size_t)max_rows * M * sizeof(float);
    if (getenv("SFVOS_MASK_TARGETS_GENERIC") == nullptr && M <= 64 && max_dim / M + 3 <= MT_SPAN && smem <= 200 * 1024 && K < (1LL << 31)) {
        SF_CUDA(cudaFuncSetAttribute(mask_targets_sep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        mask_targets_sep_kernel<<<(int)K, 256, smem, st>>>(masks, (int)n_obj, (int)H, (int)W, rois, M, max_rows, out);
    } else {
        const long long total = K * M * M;
        mask_targets_kernel<<<(int)((total + 7) / 8), 256, 0, st>>>(masks, (int)n_obj, (int)H, (int)W, rois, K, M, out);
    }
    SF_LAUNCH_CHECK();
    return SFVOS_OK;
}
