// HBM-bound kernels around the GEMMs: weight packing, BatchNorm statistics / finalize / apply / backward,
// ReLU backward, NCHW <-> channels-last layout conversion.  All are coalesced along the contiguous channel axis
// with 16/32-byte vector accesses; per-channel reductions use registers -> shared memory -> one atomic per
// channel per CTA.
#include "common.cuh"

namespace {

// ---- 8-channel vector access helpers ----------------------------------------------------------------------------
__device__ __forceinline__ void load8(const float* p, float (&v)[8]) {
    const float4 a = *reinterpret_cast<const float4*>(p);
    const float4 b = *reinterpret_cast<const float4*>(p + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void load8(const __nv_bfloat16* p, float (&v)[8]) {
    const uint4 u = *reinterpret_cast<const uint4*>(p);
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        v[2 * i] = __uint_as_float(w[i] << 16);
        v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
}
__device__ __forceinline__ void store8(float* p, const float (&v)[8]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ void store8(__nv_bfloat16* p, const float (&v)[8]) {
    uint4 u;
    u.x = pack_bf16x2(v[0], v[1]); u.y = pack_bf16x2(v[2], v[3]);
    u.z = pack_bf16x2(v[4], v[5]); u.w = pack_bf16x2(v[6], v[7]);
    *reinterpret_cast<uint4*>(p) = u;
}

// ---- weight packing ---------------------------------------------------------------------------------------------
struct PackArgs {
    const float* w; void* out;
    int out_bf16, mode, Cout, Cin, kt, kh, kw, Cp, tap_i, tap_j, N, Kc, taps;
};
__global__ void pack_weights_kernel(const PackArgs a) {
    const long long total = (long long)a.N * a.taps * a.Cp;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(idx % a.Cp);
        const int tap = (int)((idx / a.Cp) % a.taps);
        const int n = (int)(idx / ((long long)a.Cp * a.taps));
        float v = 0.f;
        if (c < a.Kc) {
            if (a.mode == 0 || a.mode == 1) {
                int tj = tap % a.kw, ti = (tap / a.kw) % a.kh, ta = tap / (a.kw * a.kh);
                int co, ci;
                if (a.mode == 0) { co = n; ci = c; }
                else { co = c; ci = n; ta = a.kt - 1 - ta; ti = a.kh - 1 - ti; tj = a.kw - 1 - tj; }
                v = a.w[((((long long)co * a.Cin + ci) * a.kt + ta) * a.kh + ti) * a.kw + tj];
            } else {
                const int ci = a.mode == 2 ? c : n;
                const int co = a.mode == 2 ? n : c;
                v = a.w[(((long long)ci * a.Cout + co) * a.kh + a.tap_i) * a.kw + a.tap_j];
            }
        }
        const long long k = (long long)tap * a.Cp + c;
        if (a.out_bf16) reinterpret_cast<__nv_bfloat16*>(a.out)[(long long)n * a.taps * a.Cp + k] = __float2bfloat16(v);
        else reinterpret_cast<float*>(a.out)[k * a.N + n] = v;
    }
}

struct UnpackArgs { const float* dw; float* grad; int mode, Cout, Cin, kt, kh, kw, tap_i, tap_j; };
__global__ void unpack_wgrad_kernel(const UnpackArgs a) {
    if (a.mode == 0) {
        const long long total = (long long)a.Cout * a.Cin * a.kt * a.kh * a.kw;
        for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
             idx += (long long)gridDim.x * blockDim.x) {
            long long r = idx;
            const int tj = (int)(r % a.kw); r /= a.kw;
            const int ti = (int)(r % a.kh); r /= a.kh;
            const int ta = (int)(r % a.kt); r /= a.kt;
            const int ci = (int)(r % a.Cin); const int co = (int)(r / a.Cin);
            const int tap = (ta * a.kh + ti) * a.kw + tj;
            a.grad[idx] += a.dw[((long long)tap * a.Cin + ci) * a.Cout + co];
        }
    } else {
        const long long total = (long long)a.Cin * a.Cout;
        for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
             idx += (long long)gridDim.x * blockDim.x) {
            const int co = (int)(idx % a.Cout), ci = (int)(idx / a.Cout);
            a.grad[(((long long)ci * a.Cout + co) * a.kh + a.tap_i) * a.kw + a.tap_j] += a.dw[idx];
        }
    }
}

// 2-D transposes for the single-tap fully-connected layers of the box head (fc6: 1024 x 12544), where the generic
// gather kernels above would touch one 32-byte sector per element: dst[r][c] (+)= src[c][r], 32 x 32 tiles through
// shared memory, both sides coalesced.
template <typename OutT, bool ACC>
__global__ void __launch_bounds__(256) transpose2d_kernel(const float* __restrict__ src, OutT* dst, int R, int Ccols) {
    __shared__ float tile[32][33];
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int i = ty; i < 32; i += 8) {
        const int c = c0 + i, r = r0 + tx;
        tile[i][tx] = (c < Ccols && r < R) ? src[(long long)c * R + r] : 0.f;
    }
    __syncthreads();
    for (int i = ty; i < 32; i += 8) {
        const int r = r0 + i, c = c0 + tx;
        if (r < R && c < Ccols) {
            OutT* d = dst + (long long)r * Ccols + c;
            float v = tile[tx][i];
            if (ACC) v += static_cast<float>(*d);
            *d = static_cast<OutT>(v);
        }
    }
}

// fc fprop operand: [N][K] f32 -> [N][K] bf16, 8 elements per thread (the generic pack kernel spends its time on index arithmetic)
__global__ void __launch_bounds__(256) convert_bf16x8_kernel(const float* __restrict__ src, __nv_bfloat16* dst, long long n8) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(src) + 2 * i);
        const float4 b = __ldg(reinterpret_cast<const float4*>(src) + 2 * i + 1);
        uint4 u;
        u.x = pack_bf16x2(a.x, a.y); u.y = pack_bf16x2(a.z, a.w);
        u.z = pack_bf16x2(b.x, b.y); u.w = pack_bf16x2(b.z, b.w);
        reinterpret_cast<uint4*>(dst)[i] = u;
    }
}

// ---- per-channel reductions -------------------------------------------------------------------------------------
// Block of 256 threads: thread -> (pixel lane, 8-channel group).  G = C/8 groups, L = 256/G pixel lanes.
constexpr int RED_THREADS = 256;
// Pixels per CTA: >= 32 pixel iterations per thread, so the per-channel constants a thread loads into registers are
// amortised also for the 32-channel fast-pathway tensors (C = 32 -> 2048 px, C >= 128 -> 512 px); small chunks keep
// >= 1000 CTAs even for one frame of 192x336.
__host__ __device__ inline int red_ppc(int C) {
    const int p = 65536 / C;
    return p < 512 ? 512 : p;
}

template <int NQ>
__device__ __forceinline__ void block_reduce_to_global(float (&acc)[NQ][8], int g, int lane_pix, int G, int L,
                                                       float* const (&dst)[NQ], float* smem) {
    // smem: [NQ][L][G*8]
#pragma unroll
    for (int q = 0; q < NQ; ++q)
#pragma unroll
        for (int j = 0; j < 8; ++j)
            if (g < G && lane_pix < L) smem[(q * L + lane_pix) * G * 8 + g * 8 + j] = acc[q][j];
    __syncthreads();
    for (int i = threadIdx.x; i < NQ * G * 8; i += blockDim.x) {
        const int q = i / (G * 8), c = i - q * G * 8;
        float s = 0.f;
        for (int l = 0; l < L; ++l) s += smem[(q * L + l) * G * 8 + c];
        atomicAdd(dst[q] + c, s);
    }
}

// Deterministic variant of the reduction tail (fp32 validation mode): the CTA's per-channel sums are formed in a fixed
// order in fp64 and written to partial[blockIdx.x][q][c]; reduce_partials_kernel then adds the CTAs' rows in index order.
// No atomics anywhere, so the result does not depend on scheduling.
template <int NQ>
__device__ __forceinline__ void block_reduce_to_partial(double (&acc)[NQ][8], int g, int lane_pix, int G, int L,
                                                        double* partial, double* smem) {
#pragma unroll
    for (int q = 0; q < NQ; ++q)
#pragma unroll
        for (int j = 0; j < 8; ++j)
            if (g < G && lane_pix < L) smem[(q * L + lane_pix) * G * 8 + g * 8 + j] = acc[q][j];
    __syncthreads();
    for (int i = threadIdx.x; i < NQ * G * 8; i += blockDim.x) {
        const int q = i / (G * 8), c = i - q * G * 8;
        double s = 0.0;
        for (int l = 0; l < L; ++l) s += smem[(q * L + l) * G * 8 + c];
        partial[(long long)blockIdx.x * NQ * G * 8 + i] = s;
    }
}

// out[i] = sum over CTAs (in index order) of partial[cta][i], i < n; OutT = double (BN statistics) or float (backward sums)
template <typename OutT>
__global__ void reduce_partials_kernel(const double* __restrict__ partial, int n_ctas, int n, OutT* out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double s = 0.0;
    for (int b = 0; b < n_ctas; ++b) s += partial[(long long)b * n + i];
    out[i] = static_cast<OutT>(s);
}

// fp32 validation mode: fp64 sums (|mean| >> std must not cancel), fixed reduction order
__global__ void __launch_bounds__(RED_THREADS)
channel_stats_kernel(const float* __restrict__ x, long long npix, int C, long long cstride, double* partial, int ppc) {
    extern __shared__ double red_smem_d[];
    const int G = C / 8, L = RED_THREADS / G;
    const int g = threadIdx.x % G, lp = threadIdx.x / G;
    double acc[2][8] = {};
    const long long p0 = (long long)blockIdx.x * ppc;
    long long p1 = p0 + ppc; if (p1 > npix) p1 = npix;
    if (lp < L)
        for (long long p = p0 + lp; p < p1; p += L) {
            float v[8];
            load8(x + p * cstride + g * 8, v);
#pragma unroll
            for (int j = 0; j < 8; ++j) { const double d = v[j]; acc[0][j] += d; acc[1][j] = fma(d, d, acc[1][j]); }
        }
    block_reduce_to_partial<2>(acc, g, lp, G, L, partial, red_smem_d);
}

// AccT = float: per-CTA sums merged into ``sums`` with atomics (product path).  AccT = double: fp64 sums written to
// ``partial`` for the fixed-order merge (validation mode).
template <typename DyT, typename AccT>
__global__ void __launch_bounds__(RED_THREADS)
bn_bwd_reduce_kernel(const DyT* __restrict__ dy, long long dy_cstride, const float* __restrict__ x, long long x_cstride,
                     const float* __restrict__ scale, const float* __restrict__ shift, const float* __restrict__ mean,
                     const float* __restrict__ rstd, int relu, long long npix, int C, float* sums, double* partial, int ppc) {
    extern __shared__ double red_smem_d[];
    const int G = C / 8, L = RED_THREADS / G;
    const int g = threadIdx.x % G, lp = threadIdx.x / G;
    AccT acc[2][8] = {};
    float sc[8], sh[8], mu[8], rs[8];
    if (lp < L) {
#pragma unroll
        for (int j = 0; j < 8; ++j) { sc[j] = scale[g * 8 + j]; sh[j] = shift[g * 8 + j]; mu[j] = mean[g * 8 + j]; rs[j] = rstd[g * 8 + j]; }
    }
    const long long p0 = (long long)blockIdx.x * ppc;
    long long p1 = p0 + ppc; if (p1 > npix) p1 = npix;
    if (lp < L)
        for (long long p = p0 + lp; p < p1; p += 2 * L) {
            const long long q = p + L;
            const bool two = q < p1;
            float d0[8], v0[8], d1[8], v1[8];
            load8(dy + p * dy_cstride + g * 8, d0);
            load8(x + p * x_cstride + g * 8, v0);
            if (two) { load8(dy + q * dy_cstride + g * 8, d1); load8(x + q * x_cstride + g * 8, v1); }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float dm = (relu && fmaf(v0[j], sc[j], sh[j]) <= 0.f) ? 0.f : d0[j];
                acc[0][j] += dm;
                acc[1][j] += (AccT)dm * ((AccT)(v0[j] - mu[j]) * (AccT)rs[j]);
            }
            if (two) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float dm = (relu && fmaf(v1[j], sc[j], sh[j]) <= 0.f) ? 0.f : d1[j];
                    acc[0][j] += dm;
                    acc[1][j] += (AccT)dm * ((AccT)(v1[j] - mu[j]) * (AccT)rs[j]);
                }
            }
        }
    if constexpr (sizeof(AccT) == 8) {
        block_reduce_to_partial<2>(acc, g, lp, G, L, partial, red_smem_d);
    } else {
        float* const dst[2] = {sums, sums + C};
        block_reduce_to_global<2>(acc, g, lp, G, L, dst, reinterpret_cast<float*>(red_smem_d));
    }
}

// dx = k1*dm - k2 - k3*(x - mean) with k1 = gamma*rstd, k2 = k1*sum(dm)/n, k3 = k1*rstd*sum(dm*xhat)/n.
// Thread -> (pixel lane, fixed 8-channel group): the per-channel constants live in registers for the whole pixel loop.
template <typename DyT, typename DxT>
__global__ void __launch_bounds__(RED_THREADS)
bn_bwd_apply_kernel(const DyT* __restrict__ dy, long long dy_cstride, const float* __restrict__ x, long long x_cstride,
                    const float* __restrict__ scale, const float* __restrict__ shift, const float* __restrict__ mean,
                    const float* __restrict__ rstd, const float* __restrict__ gamma, int relu, long long npix, int C,
                    const float* __restrict__ sums, DxT* dx, long long dx_cstride, float* dgamma, float* dbeta,
                    int fixed_stats, float* dbias, int ppc) {
    const int G = C / 8, L = RED_THREADS / G;
    const int g = threadIdx.x % G, lp = threadIdx.x / G;
    // fixed_stats (eval-mode BN: mean / rstd are constants, not functions of x): dx = gamma*rstd*dm, nothing else
    const float inv_n = fixed_stats ? 0.f : 1.0f / (float)npix;
    if (blockIdx.x == 0 && dgamma != nullptr)
        for (int c = threadIdx.x; c < C; c += blockDim.x) {     // atomics: pyramid levels on different streams share these accumulators
            atomicAdd(dbeta + c, sums[c]);
            atomicAdd(dgamma + c, sums[C + c]);
            if (fixed_stats && dbias != nullptr) atomicAdd(dbias + c, gamma[c] * rstd[c] * sums[c]);   // = sum over pixels of dx
        }
    if (lp >= L) return;
    float sc[8], sh[8], mu[8], k1[8], k2[8], k3[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int c = g * 8 + j;
        sc[j] = scale[c]; sh[j] = shift[c]; mu[j] = mean[c];
        k1[j] = gamma[c] * rstd[c];
        k2[j] = k1[j] * sums[c] * inv_n;
        k3[j] = k1[j] * rstd[c] * sums[C + c] * inv_n;
    }
    const long long p0 = (long long)blockIdx.x * ppc;
    long long p1 = p0 + ppc; if (p1 > npix) p1 = npix;
    for (long long p = p0 + lp; p < p1; p += 2 * L) {
        const long long q = p + L;
        const bool two = q < p1;
        float d0[8], v0[8], d1[8], v1[8];
        load8(dy + p * dy_cstride + g * 8, d0);
        load8(x + p * x_cstride + g * 8, v0);
        if (two) { load8(dy + q * dy_cstride + g * 8, d1); load8(x + q * x_cstride + g * 8, v1); }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float dm = (relu && fmaf(v0[j], sc[j], sh[j]) <= 0.f) ? 0.f : d0[j];
            d0[j] = fmaf(k1[j], dm, -k2[j]) - k3[j] * (v0[j] - mu[j]);
        }
        store8(dx + p * dx_cstride + g * 8, d0);
        if (two) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float dm = (relu && fmaf(v1[j], sc[j], sh[j]) <= 0.f) ? 0.f : d1[j];
                d1[j] = fmaf(k1[j], dm, -k2[j]) - k3[j] * (v1[j] - mu[j]);
            }
            store8(dx + q * dx_cstride + g * 8, d1);
        }
    }
}

template <typename DyT, typename YT, typename DxT>
__global__ void __launch_bounds__(RED_THREADS)
relu_bwd_kernel(const DyT* __restrict__ dy, long long dy_cstride, const YT* __restrict__ y, long long y_cstride, DxT* dx,
                long long dx_cstride, float* dbias, long long npix, int C, int ppc) {
    extern __shared__ float red_smem[];
    const int G = C / 8, L = RED_THREADS / G;
    const int g = threadIdx.x % G, lp = threadIdx.x / G;
    float acc[1][8] = {};
    const long long p0 = (long long)blockIdx.x * ppc;
    long long p1 = p0 + ppc; if (p1 > npix) p1 = npix;
    if (lp < L)
        for (long long p = p0 + lp; p < p1; p += L) {
            float d[8], v[8];
            load8(dy + p * dy_cstride + g * 8, d);
            load8(y + p * y_cstride + g * 8, v);
#pragma unroll
            for (int j = 0; j < 8; ++j) { d[j] = v[j] > 0.f ? d[j] : 0.f; acc[0][j] += d[j]; }
            store8(dx + p * dx_cstride + g * 8, d);
        }
    if (dbias != nullptr) {
        float* const dst[1] = {dbias};
        block_reduce_to_global<1>(acc, g, lp, G, L, dst, red_smem);
    }
}

template <typename XT, typename YT>
__global__ void __launch_bounds__(RED_THREADS)
affine_act_kernel(const XT* __restrict__ x, long long x_cstride, YT* y, long long y_cstride, const float* __restrict__ scale,
                  const float* __restrict__ shift, int relu, long long npix, int C, int ppc) {
    const int G = C / 8, L = RED_THREADS / G;
    const int g = threadIdx.x % G, lp = threadIdx.x / G;
    if (lp >= L) return;
    float sc[8], sh[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { sc[j] = scale[g * 8 + j]; sh[j] = shift[g * 8 + j]; }
    const long long p0 = (long long)blockIdx.x * ppc;
    long long p1 = p0 + ppc; if (p1 > npix) p1 = npix;
    for (long long p = p0 + lp; p < p1; p += 2 * L) {
        const long long q = p + L;
        const bool two = q < p1;
        float v0[8], v1[8];
        load8(x + p * x_cstride + g * 8, v0);
        if (two) load8(x + q * x_cstride + g * 8, v1);
#pragma unroll
        for (int j = 0; j < 8; ++j) { const float r = fmaf(v0[j], sc[j], sh[j]); v0[j] = relu ? fmaxf(r, 0.f) : r; }
        store8(y + p * y_cstride + g * 8, v0);
        if (two) {
#pragma unroll
            for (int j = 0; j < 8; ++j) { const float r = fmaf(v1[j], sc[j], sh[j]); v1[j] = relu ? fmaxf(r, 0.f) : r; }
            store8(y + q * y_cstride + g * 8, v1);
        }
    }
}

__device__ __forceinline__ double stat_at(const void* p, int f64, int c) {
    return f64 ? reinterpret_cast<const double*>(p)[c] : (double)reinterpret_cast<const float*>(p)[c];
}

__global__ void bn_finalize_kernel(const void* sum, const void* sumsq, int stats_f64, double count, const float* conv_bias,
                                   const float* gamma, const float* beta, float* running_mean, float* running_var,
                                   long long* nbt, double momentum, double eps, float* scale, float* shift, float* mean,
                                   float* rstd, int C) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c == 0 && nbt != nullptr) *nbt += 1;
    if (c >= C) return;
    const double m = stat_at(sum, stats_f64, c) / count;
    double var = stat_at(sumsq, stats_f64, c) / count - m * m;
    if (var < 0.0) var = 0.0;
    const double rs = 1.0 / sqrt(var + eps);
    const double g = gamma[c];
    scale[c] = (float)(g * rs);
    shift[c] = (float)((double)beta[c] - m * g * rs);
    mean[c] = (float)m;
    rstd[c] = (float)rs;
    if (running_mean != nullptr) {
        const double mb = m + (conv_bias ? (double)conv_bias[c] : 0.0);
        const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
        running_mean[c] = (float)((1.0 - momentum) * (double)running_mean[c] + momentum * mb);
        running_var[c] = (float)((1.0 - momentum) * (double)running_var[c] + momentum * unbiased);
    }
}

// Running-statistics update of ONE BatchNorm for several consecutive forward calls (the pyramid levels of one
// temporally_enhance_features call), applied in call order: exactly the sequence of exponential-moving-average steps that
// bn_finalize performs when the calls run one after another, but issued once after the levels - which run concurrently on
// their own streams - have joined.
__global__ void bn_running_update_kernel(const sfvos_bn_running_params p) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c == 0 && p.num_batches_tracked != nullptr) *reinterpret_cast<long long*>(p.num_batches_tracked) += p.n_calls;
    if (c >= p.C) return;
    float rm = p.running_mean[c], rv = p.running_var[c];
    for (int i = 0; i < p.n_calls; ++i) {
        const double count = p.count[i];
        const int f64 = p.stats_dtype == SFVOS_F64;
        const double m = stat_at(p.sum[i], f64, c) / count;
        double var = stat_at(p.sumsq[i], f64, c) / count - m * m;
        if (var < 0.0) var = 0.0;
        const double mb = m + (p.conv_bias ? (double)p.conv_bias[c] : 0.0);
        const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
        rm = (float)((1.0 - p.momentum) * (double)rm + p.momentum * mb);
        rv = (float)((1.0 - p.momentum) * (double)rv + p.momentum * unbiased);
    }
    p.running_mean[c] = rm;
    p.running_var[c] = rv;
}

__global__ void bn_fold_eval_kernel(const float* conv_bias, const float* gamma, const float* beta, const float* rm,
                                    const float* rv, double eps, float* scale, float* shift, float* mean, float* rstd, int C) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const double rs = 1.0 / sqrt((double)rv[c] + eps);
    const double s = (double)gamma[c] * rs;
    const double b = conv_bias ? (double)conv_bias[c] : 0.0;
    scale[c] = (float)s;
    shift[c] = (float)((double)beta[c] + (b - (double)rm[c]) * s);
    if (mean != nullptr) mean[c] = (float)((double)rm[c] - b);      // in terms of the bias-free conv output
    if (rstd != nullptr) rstd[c] = (float)rs;
}

// ---- layout -----------------------------------------------------------------------------------------------------
// [F][C][HW] (f32 | bf16) -> [F][HW][cstride]; tile = 64 channels x 128 pixels, 256 threads.
// Load: 8 independent 4-pixel loads per thread along the pixel axis (16 B f32 / 8 B bf16; 512 / 256 B per warp row).  Store:
// each thread writes 8 consecutive channels of one pixel (16 B bf16 / 32 B f32), 8 threads cover the tile's 64 channels.
__device__ __forceinline__ float4 load4_px(const float* r, long long p, long long HW, bool vec_ok) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (vec_ok && p + 3 < HW) return __ldg(reinterpret_cast<const float4*>(r));
    if (p < HW) v.x = r[0];
    if (p + 1 < HW) v.y = r[1];
    if (p + 2 < HW) v.z = r[2];
    if (p + 3 < HW) v.w = r[3];
    return v;
}
__device__ __forceinline__ float4 load4_px(const __nv_bfloat16* r, long long p, long long HW, bool vec_ok) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (vec_ok && p + 3 < HW) {
        const uint2 u = __ldg(reinterpret_cast<const uint2*>(r));
        v.x = __uint_as_float(u.x << 16); v.y = __uint_as_float(u.x & 0xffff0000u);
        v.z = __uint_as_float(u.y << 16); v.w = __uint_as_float(u.y & 0xffff0000u);
        return v;
    }
    if (p < HW) v.x = __bfloat162float(r[0]);
    if (p + 1 < HW) v.y = __bfloat162float(r[1]);
    if (p + 2 < HW) v.z = __bfloat162float(r[2]);
    if (p + 3 < HW) v.w = __bfloat162float(r[3]);
    return v;
}

template <typename InT, typename OutT>
__global__ void __launch_bounds__(256)
nchw_to_nhwc_kernel(const InT* __restrict__ src, long long src_fstride, OutT* dst, long long dst_cstride, int C, long long HW) {
    __shared__ float tile[64][129];
    const int f = blockIdx.z;
    const long long p0 = (long long)blockIdx.x * 128;
    const int c0 = blockIdx.y * 64;
    const InT* s = src + (long long)f * src_fstride;
    const int t = threadIdx.x;
    const int lp4 = (t & 31) * 4, lc = t >> 5;
    // 4-pixel vector loads need every channel row to start on a vector boundary
    const bool vec_ok = (HW % 4 == 0) && (src_fstride % 4 == 0) && ((reinterpret_cast<uintptr_t>(src) & (4 * sizeof(InT) - 1)) == 0);
    float4 buf[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int c = c0 + lc + 8 * i;
        const long long p = p0 + lp4;
        buf[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c < C && p < HW) buf[i] = load4_px(s + (long long)c * HW + p, p, HW, vec_ok);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        float* row = &tile[lc + 8 * i][lp4];
        row[0] = buf[i].x; row[1] = buf[i].y; row[2] = buf[i].z; row[3] = buf[i].w;
    }
    __syncthreads();
    const int cg = t & 7, pl = t >> 3;                 // 8 channel groups x 32 pixels per pass
#pragma unroll
    for (int pass = 0; pass < 4; ++pass) {
        const int pp = pass * 32 + pl;
        const long long p = p0 + pp;
        const int c = c0 + cg * 8;
        if (p < HW && c < C) {
            float v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = tile[cg * 8 + j][pp];
            OutT* d = dst + ((long long)f * HW + p) * dst_cstride + c;
            if (c + 8 <= C) store8(d, v);
            else for (int j = 0; j < 8 && c + j < C; ++j) d[j] = static_cast<OutT>(v[j]);
        }
    }
}

// bf16 -> bf16 (the product path's layout pass: FPN features arrive as bf16 [frames,C,H,W]).  Pure 16-bit transposition, no
// conversion: a thread loads 8 pixels of TWO adjacent channels (2 x 16 B), byte-permutes them into 8 words (pixel i: channel
// pair), stores the words into a [128 px][64 ch] tile whose 16-byte chunks are XOR-swizzled by pixel (conflict-free for the
// word stores - a warp holds 4 pixel groups x 8 channel pairs - and for the 16-byte row reads), and the tile leaves as
// 16-byte vectors: 8 lanes cover one pixel's 128 bytes.  16 smem words written + 4 vectors read per thread instead of
// 32 + 32 conflicted scalar accesses (ncu, round 2: the scalar version ran at 42 % of the HBM roofline with the L1 / shared
// pipe as its busiest unit).
__global__ void __launch_bounds__(256, 8)
nchw_to_nhwc_bf16_kernel(const __nv_bfloat16* __restrict__ src, long long src_fstride, __nv_bfloat16* dst, long long dst_cstride,
                         long long HW) {
    __shared__ __align__(16) uint32_t tile[128 * 32];
    const int f = blockIdx.z;
    const long long p0 = (long long)blockIdx.x * 128;
    const int c0 = blockIdx.y * 64;
    const __nv_bfloat16* s = src + (long long)f * src_fstride;
    const int t = threadIdx.x;
    uint4 va[2], vb[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const int u = t + 256 * k;
        const int pg = ((u >> 5) & 3) * 4 + (u & 3), cp = ((u >> 7) & 3) * 8 + ((u >> 2) & 7);
        const long long p = p0 + pg * 8;
        va[k] = make_uint4(0u, 0u, 0u, 0u); vb[k] = va[k];
        if (p < HW) {
            const __nv_bfloat16* q = s + (long long)(c0 + 2 * cp) * HW + p;
            va[k] = __ldg(reinterpret_cast<const uint4*>(q));
            vb[k] = __ldg(reinterpret_cast<const uint4*>(q + HW));
        }
    }
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const int u = t + 256 * k;
        const int pg = ((u >> 5) & 3) * 4 + (u & 3), cp = ((u >> 7) & 3) * 8 + ((u >> 2) & 7);
        const uint32_t a[4] = {va[k].x, va[k].y, va[k].z, va[k].w}, b[4] = {vb[k].x, vb[k].y, vb[k].z, vb[k].w};
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const uint32_t w = __byte_perm(a[i >> 1], b[i >> 1], (i & 1) ? 0x7632 : 0x5410);
            const int px = pg * 8 + i;
            const int chunk = (cp >> 2) ^ (px & 7) ^ (((px >> 3) & 3) << 1);
            tile[px * 32 + chunk * 4 + (cp & 3)] = w;
        }
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int v = t + 256 * j;
        const int k8 = v & 7, px = v >> 3;
        const long long p = p0 + px;
        if (p < HW) {
            const int chunk = k8 ^ (px & 7) ^ (((px >> 3) & 3) << 1);
            const uint4 o = *reinterpret_cast<const uint4*>(&tile[px * 32 + chunk * 4]);
            *reinterpret_cast<uint4*>(dst + ((long long)f * HW + p) * dst_cstride + c0 + k8 * 8) = o;
        }
    }
}

template <typename InT>
__global__ void __launch_bounds__(256)
nhwc_to_nchw_kernel(const InT* __restrict__ src, long long src_cstride, float* dst, int C, long long HW) {
    __shared__ float tile[64][33];
    const int f = blockIdx.z;
    const long long p0 = (long long)blockIdx.x * 32;
    const int c0 = blockIdx.y * 64;
    for (int py = threadIdx.y; py < 32; py += 8) {
        const long long p = p0 + py;
        const int c = c0 + 2 * threadIdx.x;
        float v0 = 0.f, v1 = 0.f;
        if (p < HW && c < C) {
            const InT* s = src + ((long long)f * HW + p) * src_cstride + c;
            if constexpr (sizeof(InT) == 2) {
                const uint32_t u = *reinterpret_cast<const uint32_t*>(s);
                v0 = __uint_as_float(u << 16); v1 = __uint_as_float(u & 0xffff0000u);
            } else {
                const float2 t = *reinterpret_cast<const float2*>(s);
                v0 = t.x; v1 = t.y;
            }
        }
        tile[2 * threadIdx.x][py] = v0;
        tile[2 * threadIdx.x + 1][py] = v1;
    }
    __syncthreads();
    for (int cy = threadIdx.y; cy < 64; cy += 8) {
        const long long p = p0 + threadIdx.x;
        const int c = c0 + cy;
        if (c < C && p < HW) dst[((long long)f * C + c) * HW + p] = tile[cy][threadIdx.x];
    }
}

// ---- ResNet body helpers ---------------------------------------------------------------------
// Stem patches: f32 NCHW image [N,Cin,H,W] -> rows [N*Ho*Wo, Kp] (bf16 | f32) with row[(i*kw + j)*Cin + c] =
// x[n, c, oy*stride + i - pad, ox*stride + j - pad] (0 outside the image, 0 in the Kp - kh*kw*Cin padding columns): the 7x7
// stride-2 stem convolution of 3 input channels becomes a [pixels, 192] x [192, 64] GEMM on the tensor cores.
// One thread per (output pixel, 8 consecutive K columns).
template <typename OutT>
__global__ void __launch_bounds__(256)
im2col_kernel(const float* __restrict__ x, OutT* rows, int Cin, int H, int W, int Ho, int Wo, int kh, int kw, int stride, int pad,
              int Kp, long long total) {
    const int K8 = Kp / 8;
    const int K = kh * kw * Cin;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int kg = (int)(i % K8);
        long long pix = i / K8;
        const int ox = (int)(pix % Wo); pix /= Wo;
        const int oy = (int)(pix % Ho);
        const long long n = pix / Ho;
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int k = kg * 8 + e;
            float val = 0.f;
            if (k < K) {
                const int c = k % Cin, t = k / Cin;
                const int tj = t % kw, ti = t / kw;
                const int iy = oy * stride + ti - pad, ix = ox * stride + tj - pad;
                if (iy >= 0 && iy < H && ix >= 0 && ix < W) val = __ldg(x + ((n * Cin + c) * H + iy) * (long long)W + ix);
            }
            v[e] = val;
        }
        store8(rows + (((n * Ho + oy) * Wo + ox) * (long long)Kp) + kg * 8, v);
    }
}

// max_pool2d(kernel 3, stride 2, padding 1) on a channels-last tensor; 8 channels per thread
template <typename T>
__global__ void __launch_bounds__(256)
maxpool3x3s2_kernel(const T* __restrict__ x, T* y, int H, int W, int Ho, int Wo, int C, long long total8) {
    const int C8 = C / 8;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total8; i += (long long)gridDim.x * blockDim.x) {
        const int cg = (int)(i % C8);
        long long pix = i / C8;
        const int ox = (int)(pix % Wo); pix /= Wo;
        const int oy = (int)(pix % Ho);
        const long long n = pix / Ho;
        float m[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) m[j] = -INFINITY;
        for (int di = 0; di < 3; ++di) {
            const int iy = 2 * oy + di - 1;
            if (iy < 0 || iy >= H) continue;
            for (int dj = 0; dj < 3; ++dj) {
                const int ix = 2 * ox + dj - 1;
                if (ix < 0 || ix >= W) continue;
                float v[8];
                load8(x + (((n * H + iy) * W + ix) * (long long)C) + cg * 8, v);
#pragma unroll
                for (int j = 0; j < 8; ++j) m[j] = fmaxf(m[j], v[j]);
            }
        }
        store8(y + (((n * Ho + oy) * Wo + ox) * (long long)C) + cg * 8, m);
    }
}

// residual join of a bottleneck block: out = relu(a + b) on f32 (the residual stream stays f32), plus the bf16 copy the next
// block's convolutions read
__global__ void __launch_bounds__(256)
add_relu_kernel(const float* __restrict__ a, const float* __restrict__ b, float* out, __nv_bfloat16* out_bf16, long long n8) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
        float u[8], v[8];
        load8(a + i * 8, u);
        load8(b + i * 8, v);
#pragma unroll
        for (int j = 0; j < 8; ++j) u[j] = fmaxf(u[j] + v[j], 0.f);
        if (out != nullptr) store8(out + i * 8, u);
        if (out_bf16 != nullptr) store8(out_bf16 + i * 8, u);
    }
}

// FPN top-down merge (TV/ops/feature_pyramid_network.py: F.interpolate(last_inner, size, mode="nearest") + inner_lateral):
// inner[n,h,w,:] += top[n, floor(h*Ht/H), floor(w*Wt/W), :] in place on the f32 lateral output, plus the bf16 copy the 3x3
// output convolution reads.  top == nullptr (coarsest level): only the bf16 copy.  8 channels per thread.
__global__ void __launch_bounds__(256)
upsample_add_kernel(const float* __restrict__ top, int Ht, int Wt, float* inner, __nv_bfloat16* out, int H, int W, int C, long long total8) {
    const int C8 = C / 8;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total8; i += (long long)gridDim.x * blockDim.x) {
        const int cg = (int)(i % C8);
        long long pix = i / C8;
        const int w = (int)(pix % W); pix /= W;
        const int h = (int)(pix % H);
        const long long n = pix / H;
        float v[8];
        float* ip = inner + (((n * H + h) * W + w) * (long long)C) + cg * 8;
        load8(ip, v);
        if (top != nullptr) {
            // torch 'nearest': src = min(floor(dst * in / out), in - 1), computed in float like ATen (scale = in / out)
            const int hs = min((int)floorf(h * ((float)Ht / (float)H)), Ht - 1);
            const int ws = min((int)floorf(w * ((float)Wt / (float)W)), Wt - 1);
            float t[8];
            load8(top + (((n * Ht + hs) * Wt + ws) * (long long)C) + cg * 8, t);
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] += t[j];
            store8(ip, v);
        }
        if (out != nullptr) store8(out + (((n * H + h) * W + w) * (long long)C) + cg * 8, v);
    }
}

__global__ void axpby_kernel(const float* __restrict__ x, float* y, float a, float b, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        y[i] = a * x[i] + (b == 0.f ? 0.f : b * y[i]);
}

inline int grid_for(long long work_items, int threads, int max_waves = 8) {
    long long g = (work_items + threads - 1) / threads;
    long long cap = (long long)sfvos_num_sms() * max_waves;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}

}  // namespace

#define CS(s) reinterpret_cast<cudaStream_t>(s)
#define CHECK_C8(C) SF_CHECK((C) % 8 == 0 && (C) >= 8 && (C) <= 2048, "channel count %lld must be a multiple of 8 in [8,2048]", (long long)(C))

extern "C" int sfvos_pack_weights(const float* w, void* out, int32_t out_dtype, int32_t mode, int64_t Cout, int64_t Cin,
                                  int64_t kt, int64_t kh, int64_t kw, int64_t Cp, int64_t tap_i, int64_t tap_j,
                                  sfvos_stream stream) {
    SF_CHECK(mode >= 0 && mode <= 3, "pack_weights: bad mode %d", mode);
    PackArgs a;
    a.w = w; a.out = out; a.out_bf16 = (out_dtype == SFVOS_BF16); a.mode = mode;
    a.Cout = (int)Cout; a.Cin = (int)Cin; a.kt = (int)kt; a.kh = (int)kh; a.kw = (int)kw; a.Cp = (int)Cp;
    a.tap_i = (int)tap_i; a.tap_j = (int)tap_j;
    a.N = (mode == 0 || mode == 2) ? (int)Cout : (int)Cin;
    a.Kc = (mode == 0 || mode == 2) ? (int)Cin : (int)Cout;
    a.taps = (mode <= 1) ? (int)(kt * kh * kw) : 1;
    SF_CHECK(Cp >= a.Kc, "pack_weights: Cp=%lld smaller than the channel count %d", (long long)Cp, a.Kc);
    const long long total = (long long)a.N * a.taps * a.Cp;
    if (mode == 0 && a.taps == 1 && a.out_bf16 && Cp == Cin && total % 8 == 0 && total >= (1 << 20) &&
        (reinterpret_cast<uintptr_t>(w) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0) {
        // fc fprop operand: the state_dict layout [Cout][Cin] already is the K-major operand -> a pure conversion
        convert_bf16x8_kernel<<<grid_for(total / 8, 256), 256, 0, CS(stream)>>>(w, reinterpret_cast<__nv_bfloat16*>(out), total / 8);
        SF_LAUNCH_CHECK();
        return SFVOS_OK;
    }
    if (mode == 1 && a.taps == 1 && a.out_bf16 && Cp == Cout && Cout >= 256 && Cin >= 256) {
        // fc dgrad operand: out[ci][co] = w[co][ci]
        dim3 grid((unsigned)((Cout + 31) / 32), (unsigned)((Cin + 31) / 32));
        transpose2d_kernel<__nv_bfloat16, false><<<grid, 256, 0, CS(stream)>>>(w, reinterpret_cast<__nv_bfloat16*>(out), (int)Cin, (int)Cout);
        SF_LAUNCH_CHECK();
        return SFVOS_OK;
    }
    pack_weights_kernel<<<grid_for(total, 256), 256, 0, CS(stream)>>>(a);
    SF_LAUNCH_CHECK();
    return SFVOS_OK;
}

extern "C" int sfvos_unpack_wgrad(const float* dw, float* grad, int32_t mode, int64_t Cout, int64_t Cin, int64_t kt,
                                  int64_t kh, int64_t kw, int64_t tap_i, int64_t tap_j, sfvos_stream stream) {
    SF_CHECK(mode == 0 || mode == 2, "unpack_wgrad: bad mode %d", mode);
    UnpackArgs a{dw, grad, mode, (int)Cout, (int)Cin, (int)kt, (int)kh, (int)kw, (int)tap_i, (int)tap_j};
    const long long total = mode == 0 ? Cout * Cin * kt * kh * kw : Cout * Cin;
    if (mode == 0 && kt * kh * kw == 1 && Cout >= 256 && Cin >= 256) {
        // fc weight gradient: grad[co][ci] += dw[ci][co]
        dim3 grid((unsigned)((Cin + 31) / 32), (unsigned)((Cout + 31) / 32));
        transpose2d_kernel<float, true><<<grid, 256, 0, CS(stream)>>>(dw, grad, (int)Cout, (int)Cin);
        SF_LAUNCH_CHECK();
        return SFVOS_OK;
    }
    unpack_wgrad_kernel<<<grid_for(total, 256), 256, 0, CS(stream)>>>(a);
    SF_LAUNCH_CHECK();
    return SFVOS_OK;
}

// Pixels per CTA.  A thread walks its CTA's pixels L = 256 / (C/8) at a time, i.e. ppc / L dependent load round trips: fine
// when thousands of CTAs hide each other's latency, but a small pyramid level (grid < #SMs) exposed the whole chain --
// ncu: ~32 us for EVERY 192-channel launch of levels 2..4 whatever its size.  Shrink the chunk until the grid covers the
// machine a few times over (floor: 32 pixels).
static inline int pick_ppc(long long npix, int base) {
    int ppc = base;
    const long long want = 6LL * sfvos_num_sms();
    while (ppc > 32 && (npix + ppc - 1) / ppc < want) ppc >>= 1;
    return ppc;
}
static inline size_t red_smem_bytes(int nq, int C, size_t esz = sizeof(float)) { return (size_t)nq * (RED_THREADS / (C / 8)) * C * esz; }

// grid of the per-channel reductions (channel_stats, bn_bwd_reduce) = rows of fp64 partials the deterministic mode needs
static inline int red_grid(long long npix, int C, int* ppc_out) {
    const int ppc = pick_ppc(npix, red_ppc(C));
    if (ppc_out) *ppc_out = ppc;
    return (int)((npix + ppc - 1) / ppc);
}

extern "C" int64_t sfvos_reduce_workspace_bytes(int64_t npix, int64_t C) {
    if (npix <= 0 || C <= 0 || C % 8) return 0;
    return (int64_t)red_grid(npix, (int)C, nullptr) * 2 * C * (int64_t)sizeof(double);
}

extern "C" int sfvos_channel_stats(const float* x, int64_t npix, int64_t C, int64_t cstride, double* stats,
                                   void* workspace, int64_t workspace_bytes, sfvos_stream stream) {
    CHECK_C8(C);
    SF_CHECK(cstride % 4 == 0, "channel_stats: cstride must be a multiple of 4");
    SF_CHECK(C <= 512, "channel_stats: at most 512 channels (fp64 staging in shared memory)");
    if (npix == 0) return SFVOS_OK;
    int ppc;
    const int grid = red_grid(npix, (int)C, &ppc);
    SF_CHECK(workspace != nullptr && workspace_bytes >= sfvos_reduce_workspace_bytes(npix, C) &&
             (reinterpret_cast<uintptr_t>(workspace) & 7) == 0 && (reinterpret_cast<uintptr_t>(stats) & 7) == 0,
             "channel_stats: needs an 8-byte aligned workspace of sfvos_reduce_workspace_bytes(npix, C) = %lld bytes",
             (long long)sfvos_reduce_workspace_bytes(npix, C));
    double* partial = reinterpret_cast<double*>(workspace);
    channel_stats_kernel<<<grid, RED_THREADS, red_smem_bytes(2, (int)C, sizeof(double)), CS(stream)>>>(x, npix, (int)C, cstride, partial, ppc);
    SF_LAUNCH_CHECK();
    reduce_partials_kernel<double><<<(int)((2 * C + 127) / 128), 128, 0, CS(stream)>>>(partial, grid, (int)(2 * C), stats);
    SF_LAUNCH_CHECK();
    return SFVOS_OK;
}

extern "C" int sfvos_bn_finalize(const void* sum, const void* sumsq, int32_t stats_dtype, double count, const float* conv_bias,
                                 const float* gamma, const float* beta, float* running_mean, float* running_var,
                                 int64_t* num_batches_tracked, double momentum, double eps, float* scale, float* shift,
                                 float* mean, float* rstd, int64_t C, sfvos_stream stream) {
    SF_CHECK(count > 0, "bn_finalize: empty batch");
    SF_CHECK(stats_dtype == SFVOS_F32 || stats_dtype == SFVOS_F64, "bn_finalize: statistics are f32 or f64");
    bn_finalize_kernel<<<(int)((C + 127) / 128), 128, 0, CS(stream)>>>(sum, sumsq, stats_dtype == SFVOS_F64, count, conv_bias, gamma, beta,
        running_mean, running_var, reinterpret_cast<long long*>(num_batches_tracked), momentum, eps, scale, shift, mean, rstd, (int)C);
    SF_LAUNCH_CHECK();
    return SFVOS_OK;
}

extern "C" int sfvos_bn_running_update(const sfvos_bn_running_params* p, sfvos_stream stream) {
    SF_CHECK(p != nullptr && p->n_calls >= 1 && p->n_calls <= SFVOS_BN_MAX_CALLS, "bn_running_update: 1..%d calls", SFVOS_BN_MAX_CALLS);
    SF_CHECK(p->running_mean != nullptr && p->running_var != nullptr, "bn_running_update: running buffers required");
    SF_CHECK(p->stats_dtype == SFVOS_F32 || p->stats_dtype == SFVOS_F64, "bn_running_update: statistics are f32 or f64");
    for (int i = 0; i < p->n_calls; ++i) SF_CHECK(p->count[i] > 0 && p->sum[i] && p->sumsq[i], "bn_running_update: empty call %d", i);
    bn_running_update_kernel<<<(int)((p->C + 127) / 128), 128, 0, CS(stream)>>>(*p);
    SF_LAUNCH_CHECK();
    return SFVOS_OK;
}

extern "C" int sfvos_bn_fold_eval(const float* conv_bias, const float* gamma, const float* beta, const float* running_mean,
                                  const float* running_var, double eps, float* scale, float* shift, float* mean,
                                  float* rstd, int64_t C, sfvos_stream stream) {
    bn_fold_eval_kernel<<<(int)((C + 127) / 128), 128, 0, CS(stream)>>>(conv_bias, gamma, beta, running_mean, running_var, eps, scale, shift, mean, rstd, (int)C);
    SF_LAUNCH_CHECK();
    return SFVOS_OK;
}

extern "C" int sfvos_affine_act(const void* x, int32_t x_dtype, int64_t x_cstride, void* y, int32_t y_dtype,
                                int64_t y_cstride, const float* scale, const float* shift, int32_t relu, int64_t npix,
                                int64_t C, sfvos_stream stream) {
    CHECK_C8(C);
    SF_CHECK(x_cstride % 8 == 0 && y_cstride % 8 == 0, "affine_act: strides must be multiples of 8");
    if (npix == 0) return SFVOS_OK;
    const int ppc = pick_ppc(npix, 512);
    const int grid = (int)((npix + ppc - 1) / ppc);
    using bf = __nv_bfloat16;
#define LAUNCH(XT, YT) affine_act_kernel<XT, YT><<<grid, RED_THREADS, 0, CS(stream)>>>(reinterpret_cast<const XT*>(x), x_cstride, reinterpret_cast<YT*>(y), y_cstride, scale, shift, relu, npix, (int)C, ppc)
    if (x_dtype == SFVOS_F32 && y_dtype == SFVOS_F32) LAUNCH(float, float);
    else if (x_dtype == SFVOS_F32) LAUNCH(float, bf);
    else if (y_dtype == SFVOS_F32) LAUNCH(bf, float);
    else LAUNCH(bf, bf);
#undef LAUNCH
    SF_LAUNCH_CHECK();
    return SFVOS_OK;
}

extern "C" int sfvos_bn_bwd_reduce(const void* dy, int32_t dy_dtype, int64_t dy_cstride, const float* x, int64_t x_cstride,
                                   const float* scale, const float* shift, const float* mean, const float* rstd,
                                   int32_t relu, int64_t npix, int64_t C, float* sums, void* workspace,
                                   int64_t workspace_bytes, sfvos_stream stream) {
    CHECK_C8(C);
    SF_CHECK(dy_cstride % 8 == 0 && x_cstride % 8 == 0, "bn_bwd_reduce: strides must be multiples of 8");
    if (npix == 0) return SFVOS_OK;
    int ppc;
    const int grid = red_grid(npix, (int)C, &ppc);
    if (workspace != nullptr) {
        // validation mode: fp64 partial rows + fixed-order merge; ``sums`` is overwritten (not accumulated into)
        SF_CHECK(dy_dtype == SFVOS_F32 && C <= 512, "bn_bwd_reduce: the deterministic mode takes f32 gradients, <= 512 channels");
        SF_CHECK(workspace_bytes >= sfvos_reduce_workspace_bytes(npix, C) && (reinterpret_cast<uintptr_t>(workspace) & 7) == 0,
                 "bn_bwd_reduce: workspace needs sfvos_reduce_workspace_bytes(npix, C) = %lld bytes, 8-byte aligned",
                 (long long)sfvos_reduce_workspace_bytes(npix, C));
        double* partial = reinterpret_cast<double*>(workspace);
        bn_bwd_reduce_kernel<float, double><<<grid, RED_THREADS, red_smem_bytes(2, (int)C, sizeof(double)), CS(stream)>>>(
            reinterpret_cast<const float*>(dy), dy_cstride, x, x_cstride, scale, shift, mean, rstd, relu, npix, (int)C, sums, partial, ppc);
        SF_LAUNCH_CHECK();
        reduce_partials_kernel<float><<<(int)((2 * C + 127) / 128), 128, 0, CS(stream)>>>(partial, grid, (int)(2 * C), sums);
        SF_LAUNCH_CHECK();
        return SFVOS_OK;
    }
    const size_t sm = red_smem_bytes(2, (int)C);
    if (dy_dtype == SFVOS_F32)
        bn_bwd_reduce_kernel<float, float><<<grid, RED_THREADS, sm, CS(stream)>>>(reinterpret_cast<const float*>(dy), dy_cstride, x, x_cstride, scale, shift, mean, rstd, relu, npix, (int)C, sums, nullptr, ppc);
    else
        bn_bwd_reduce_kernel<__nv_bfloat16, float><<<grid, RED_THREADS, sm, CS(stream)>>>(reinterpret_cast<const __nv_bfloat16*>(dy), dy_cstride, x, x_cstride, scale, shift, mean, rstd, relu, npix, (int)C, sums, nullptr, ppc);
    SF_LAUNCH_CHECK();
    return SFVOS_OK;
}

extern "C" int sfvos_bn_bwd_apply(const void* dy, int32_t dy_dtype, int64_t dy_cstride, const float* x, int64_t x_cstride,
                                  const float* scale, const float* shift, const float* mean, const float* rstd,
                                  const float* gamma, int32_t relu, int64_t npix, int64_t C, const float* sums, void* dx,
                                  int32_t dx_dtype, int64_t dx_cstride, float* dgamma, float* dbeta, int32_t fixed_stats,
                                  float* dbias, sfvos_stream stream) {
    CHECK_C8(C);
    SF_CHECK(dy_cstride % 8 == 0 && x_cstride % 8 == 0 && dx_cstride % 8 == 0, "bn_bwd_apply: strides must be multiples of 8");
    SF_CHECK((dgamma == nullptr) == (dbeta == nullptr), "bn_bwd_apply: dgamma and dbeta go together");
    if (npix == 0) return SFVOS_OK;
    const int ppc = pick_ppc(npix, 512);
    const int grid = (int)((npix + ppc - 1) / ppc);
    using bf = __nv_bfloat16;
#define LAUNCH(DT, XT) bn_bwd_apply_kernel<DT, XT><<<grid, RED_THREADS, 0, CS(stream)>>>(reinterpret_cast<const DT*>(dy), dy_cstride, x, x_cstride, scale, shift, mean, rstd, gamma, relu, npix, (int)C, sums, reinterpret_cast<XT*>(dx), dx_cstride, dgamma, dbeta, fixed_stats, dbias, ppc)
    if (dy_dtype == SFVOS_F32 && dx_dtype == SFVOS_F32) LAUNCH(float, float);
    else if (dy_dtype == SFVOS_F32) LAUNCH(float, bf);
    else if (dx_dtype == SFVOS_F32) LAUNCH(bf, float);
    else LAUNCH(bf, bf);
#undef LAUNCH
    SF_LAUNCH_CHECK();
    return SFVOS_OK;
}

extern "C" int sfvos_relu_bwd(const void* dy, int32_t dy_dtype, int64_t dy_cstride, const void* y, int32_t y_dtype,
                              int64_t y_cstride, void* dx, int32_t dx_dtype, int64_t dx_cstride, float* dbias, int64_t npix,
                              int64_t C, sfvos_stream stream) {
    CHECK_C8(C);
    SF_CHECK(dy_cstride % 8 == 0 && y_cstride % 8 == 0 && dx_cstride % 8 == 0, "relu_bwd: strides must be multiples of 8");
    if (npix == 0) return SFVOS_OK;
    const int ppc = pick_ppc(npix, red_ppc((int)C));
    const int grid = (int)((npix + ppc - 1) / ppc);
    const size_t sm = red_smem_bytes(1, (int)C);
    using bf = __nv_bfloat16;
#define LAUNCH(DT, YT, XT) relu_bwd_kernel<DT, YT, XT><<<grid, RED_THREADS, sm, CS(stream)>>>(reinterpret_cast<const DT*>(dy), dy_cstride, reinterpret_cast<const YT*>(y), y_cstride, reinterpret_cast<XT*>(dx), dx_cstride, dbias, npix, (int)C, ppc)
    if (dy_dtype == SFVOS_F32 && y_dtype == SFVOS_F32 && dx_dtype == SFVOS_F32) LAUNCH(float, float, float);
    else if (dy_dtype == SFVOS_F32 && y_dtype == SFVOS_BF16 && dx_dtype == SFVOS_BF16) LAUNCH(float, bf, bf);
    else if (dy_dtype == SFVOS_BF16 && y_dtype == SFVOS_BF16 && dx_dtype == SFVOS_BF16) LAUNCH(bf, bf, bf);
    else { sfvos_set_error("relu_bwd: unsupported dtype combination (%d,%d,%d)", dy_dtype, y_dtype, dx_dtype); return SFVOS_ERR_INVALID; }
#undef LAUNCH
    SF_LAUNCH_CHECK();
    return SFVOS_OK;
}

extern "C" int sfvos_nchw_to_nhwc(const void* src, int32_t src_dtype, int64_t src_fstride, void* dst, int32_t dst_dtype,
                                  int64_t dst_cstride, int64_t F, int64_t C, int64_t HW, sfvos_stream stream) {
    SF_CHECK(C % 8 == 0 && dst_cstride % 8 == 0, "nchw_to_nhwc: C and cstride must be multiples of 8");
    SF_CHECK(F <= 65535, "nchw_to_nhwc: too many frames in one call");
    SF_CHECK(src_dtype == SFVOS_F32 || src_dtype == SFVOS_BF16, "nchw_to_nhwc: source must be f32 or bf16");
    if (F == 0 || HW == 0) return SFVOS_OK;
    dim3 grid((unsigned)((HW + 127) / 128), (unsigned)((C + 63) / 64), (unsigned)F), block(256);
    using bf = __nv_bfloat16;
#define LAUNCH(IT, OT) nchw_to_nhwc_kernel<IT, OT><<<grid, block, 0, CS(stream)>>>(reinterpret_cast<const IT*>(src), src_fstride, reinterpret_cast<OT*>(dst), dst_cstride, (int)C, HW)
    if (src_dtype == SFVOS_BF16 && dst_dtype == SFVOS_BF16 && C % 64 == 0 && HW % 8 == 0 && src_fstride % 8 == 0 &&
        (reinterpret_cast<uintptr_t>(src) & 15) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
        nchw_to_nhwc_bf16_kernel<<<grid, block, 0, CS(stream)>>>(reinterpret_cast<const bf*>(src), src_fstride, reinterpret_cast<bf*>(dst),
                                                                 dst_cstride, HW);
        SF_LAUNCH_CHECK();
        return SFVOS_OK;
    }
    if (src_dtype == SFVOS_F32 && dst_dtype == SFVOS_BF16) LAUNCH(float, bf);
    else if (src_dtype == SFVOS_F32) LAUNCH(float, float);
    else if (dst_dtype == SFVOS_BF16) LAUNCH(bf, bf);
    else LAUNCH(bf, float);
#undef LAUNCH
    SF_LAUNCH_CHECK();
    return SFVOS_OK;
}

extern "C" int sfvos_nhwc_to_nchw(const void* src, int32_t src_dtype, int64_t src_cstride, float* dst, int64_t F, int64_t C,
                                  int64_t HW, sfvos_stream stream) {
    SF_CHECK(C % 2 == 0 && src_cstride % 2 == 0, "nhwc_to_nchw: C and cstride must be even");
    SF_CHECK(F <= 65535, "nhwc_to_nchw: too many frames in one call");
    if (F == 0 || HW == 0) return SFVOS_OK;
    dim3 grid((unsigned)((HW + 31) / 32), (unsigned)((C + 63) / 64), (unsigned)F), block(32, 8);
    if (src_dtype == SFVOS_BF16) nhwc_to_nchw_kernel<__nv_bfloat16><<<grid, block, 0, CS(stream)>>>(reinterpret_cast<const __nv_bfloat16*>(src), src_cstride, dst, (int)C, HW);
    else nhwc_to_nchw_kernel<float><<<grid, block, 0, CS(stream)>>>(reinterpret_cast<const float*>(src), src_cstride, dst, (int)C, HW);
    SF_LAUNCH_CHECK();
    return SFVOS_OK;
}

extern "C" int sfvos_im2col(const float* x, void* rows, int32_t rows_dtype, int64_t N, int64_t Cin, int64_t H, int64_t W, int64_t kh,
                            int64_t kw, int64_t stride, int64_t pad, int64_t Kp, sfvos_stream stream) {
    SF_CHECK(Kp % 8 == 0 && Kp >= kh * kw * Cin, "im2col: Kp=%lld must be a multiple of 8 and >= kh*kw*Cin", (long long)Kp);
    SF_CHECK(stride >= 1 && kh >= 1 && kw >= 1, "im2col: bad geometry");
    const long long Ho = (H + 2 * pad - kh) / stride + 1, Wo = (W + 2 * pad - kw) / stride + 1;
    const long long total = N * Ho * Wo * (Kp / 8);
    if (total <= 0) return SFVOS_OK;
#define LAUNCH(T) im2col_kernel<T><<<grid_for(total, 256), 256, 0, CS(stream)>>>(x, reinterpret_cast<T*>(rows), (int)Cin, (int)H, (int)W, (int)Ho, (int)Wo, (int)kh, (int)kw, (int)stride, (int)pad, (int)Kp, total)
    if (rows_dtype == SFVOS_BF16) LAUNCH(__nv_bfloat16); else LAUNCH(float);
#undef LAUNCH
    SF_LAUNCH_CHECK();
    return SFVOS_OK;
}

extern "C" int sfvos_maxpool3x3s2(const void* x, void* y, int32_t dtype, int64_t N, int64_t H, int64_t W, int64_t C, sfvos_stream stream) {
    SF_CHECK(C % 8 == 0, "maxpool3x3s2: C must be a multiple of 8");
    const long long Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
    const long long total8 = N * Ho * Wo * (C / 8);
    if (total8 <= 0) return SFVOS_OK;
#define LAUNCH(T) maxpool3x3s2_kernel<T><<<grid_for(total8, 256), 256, 0, CS(stream)>>>(reinterpret_cast<const T*>(x), reinterpret_cast<T*>(y), (int)H, (int)W, (int)Ho, (int)Wo, (int)C, total8)
    if (dtype == SFVOS_BF16) LAUNCH(__nv_bfloat16); else LAUNCH(float);
#undef LAUNCH
    SF_LAUNCH_CHECK();
    return SFVOS_OK;
}

extern "C" int sfvos_add_relu(const float* a, const float* b, float* out, void* out_bf16, int64_t n, sfvos_stream stream) {
    SF_CHECK(n % 8 == 0, "add_relu: n must be a multiple of 8");
    if (n == 0) return SFVOS_OK;
    add_relu_kernel<<<grid_for(n / 8, 256), 256, 0, CS(stream)>>>(a, b, out, reinterpret_cast<__nv_bfloat16*>(out_bf16), n / 8);
    SF_LAUNCH_CHECK();
    return SFVOS_OK;
}

extern "C" int sfvos_upsample_add(const float* top, int64_t Ht, int64_t Wt, float* inner, void* out_bf16, int64_t N, int64_t H,
                                  int64_t W, int64_t C, sfvos_stream stream) {
    SF_CHECK(C % 8 == 0 && inner != nullptr, "upsample_add: C must be a multiple of 8 and inner non-null");
    SF_CHECK(top == nullptr || (Ht >= 1 && Wt >= 1 && Ht <= H && Wt <= W), "upsample_add: the coarser map must not be larger");
    const long long total8 = N * H * W * (C / 8);
    if (total8 == 0) return SFVOS_OK;
    upsample_add_kernel<<<grid_for(total8, 256), 256, 0, CS(stream)>>>(top, (int)Ht, (int)Wt, inner, reinterpret_cast<__nv_bfloat16*>(out_bf16),
                                                                     (int)H, (int)W, (int)C, total8);
    SF_LAUNCH_CHECK();
    return SFVOS_OK;
}

extern "C" int sfvos_axpby(const float* x, float* y, float a, float b, int64_t n, sfvos_stream stream) {
    if (n == 0) return SFVOS_OK;
    axpby_kernel<<<grid_for(n, 256), 256, 0, CS(stream)>>>(x, y, a, b, n);
    SF_LAUNCH_CHECK();
    return SFVOS_OK;
}
