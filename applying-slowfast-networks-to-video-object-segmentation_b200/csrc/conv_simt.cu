// CUDA-core implicit-GEMM convolution (fprop/dgrad) and weight gradient: the "fp32 validation mode" of the
// north star (max-normalised error <= 1e-4 against the reference).  Same semantics as the tcgen05 kernels in
// conv_umma.cu / wgrad_umma.cu, same packed-weight K ordering, but f32 activations and f32 [K][N] weights.
//
// Validation mode means two things here, both needed for a gradient check through ReLU layers to be meaningful:
//   * products are accumulated in fp64 and rounded ONCE to the f32 result, so a pre-activation is the correctly
//     rounded value of the exact sum over the stored f32 operands -- its sign cannot be an artefact of the order
//     of an fp32 summation (a ReLU whose pre-activation sits at 1e-7 flips under fp32 accumulation noise, and one
//     flipped mask moves single weight-gradient entries by ~1e-3 of the maximum);
//   * no float atomics: split-K partial sums go to a caller-provided workspace and are reduced in a fixed order,
//     so two runs are bit-identical.
#include "common.cuh"

namespace {

constexpr int TM = 64, TN = 64, TK = 16;

struct SimtArgs {
    const float* x; const float* w; float* yf; __nv_bfloat16* yb;
    long long npix;
    int B, T, H, W, C, To, N, Cp;
    long long x_cstride, x_hstride, x_tstride, x_bstride, y_cstride;
    int kt, kh, kw, pad_t, pad_h, pad_w;
    const float* scale; const float* shift;
    int relu, accumulate;
    int OH, OW, oy_mul, oy_off, ox_mul, ox_off;
};

__global__ void __launch_bounds__(256) conv_simt_kernel(const SimtArgs a) {
    __shared__ float As[TK][TM + 4];
    __shared__ float Bs[TK][TN + 4];
    const int tid = threadIdx.x;
    const long long pix0 = (long long)blockIdx.x * TM;
    const int n0 = blockIdx.y * TN;
    const int tx = tid & 15, ty = tid >> 4;
    // loader roles
    const int lp = tid >> 2, lc = (tid & 3) * 4;          // A: pixel lp, channels lc..lc+3
    const int lk = tid >> 4, ln = (tid & 15) * 4;         // B: k row lk, columns ln..ln+3
    const long long mypix = pix0 + lp;
    const bool pix_ok = mypix < a.npix;
    int pb = 0, pt = 0, ph = 0, pw = 0;
    if (pix_ok) {
        long long r = mypix;
        pw = (int)(r % a.W); r /= a.W;
        ph = (int)(r % a.H); r /= a.H;
        pt = (int)(r % a.To); pb = (int)(r / a.To);
    }
    double acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;

    int tap = 0;
    for (int ta = 0; ta < a.kt; ++ta)
        for (int ti = 0; ti < a.kh; ++ti)
            for (int tj = 0; tj < a.kw; ++tj, ++tap) {
                const int it = pt + ta - a.pad_t, ih = ph + ti - a.pad_h, iw = pw + tj - a.pad_w;
                const bool inb = pix_ok && it >= 0 && it < a.T && ih >= 0 && ih < a.H && iw >= 0 && iw < a.W;
                const float* xp = a.x + (long long)pb * a.x_bstride + (long long)it * a.x_tstride + (long long)ih * a.x_hstride + (long long)iw * a.x_cstride;
                for (int c0 = 0; c0 < a.C; c0 += TK) {
                    float4 av = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (inb && c0 + lc < a.C) av = *reinterpret_cast<const float4*>(xp + c0 + lc);
                    As[lc + 0][lp] = av.x; As[lc + 1][lp] = av.y; As[lc + 2][lp] = av.z; As[lc + 3][lp] = av.w;
                    float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (c0 + lk < a.C && n0 + ln < a.N)
                        bv = *reinterpret_cast<const float4*>(a.w + ((long long)tap * a.Cp + c0 + lk) * a.N + n0 + ln);
                    *reinterpret_cast<float4*>(&Bs[lk][ln]) = bv;
                    __syncthreads();
#pragma unroll
                    for (int k = 0; k < TK; ++k) {
                        const float4 a4 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
                        const float4 b4 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
                        const double ar[4] = {a4.x, a4.y, a4.z, a4.w};
                        const double br[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
                        for (int i = 0; i < 4; ++i)
#pragma unroll
                            for (int j = 0; j < 4; ++j) acc[i][j] = fma(ar[i], br[j], acc[i][j]);
                    }
                    __syncthreads();
                }
            }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const long long pix = pix0 + ty * 4 + i;
        if (pix >= a.npix) continue;
        long long r = pix;
        const int w = (int)(r % a.W); r /= a.W;
        const int h = (int)(r % a.H); r /= a.H;   // r = frame index b*To + t
        const long long opix = (r * a.OH + (h * a.oy_mul + a.oy_off)) * a.OW + (w * a.ox_mul + a.ox_off);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n >= a.N) continue;
            double v = acc[i][j];
            if (a.scale) v *= (double)a.scale[n];
            if (a.shift) v += (double)a.shift[n];
            if (a.relu) v = fmax(v, 0.0);
            if (a.yf) {
                float* d = a.yf + opix * a.y_cstride + n;
                *d = (float)(a.accumulate ? ((double)*d + v) : v);
            } else {
                a.yb[opix * a.y_cstride + n] = __float2bfloat16((float)v);
            }
        }
    }
}

struct WgArgs {
    const float* x; const float* dy; float* dw; double* partial;   // partial: [splits][taps*C*N] when splits > 1
    long long npix, pix_per_split;
    int B, T, H, W, C, To, N;
    long long x_cstride, x_hstride, x_tstride, x_bstride, dy_cstride, dy_hstride, dy_tstride, dy_bstride;
    int kt, kh, kw, pad_t, pad_h, pad_w, ctiles;
    long long total;      // taps*C*N
};

// dw[tap][c][n] += sum_pix x[pix + tap offset][c] * dy[pix][n];  grid = (ctiles*ntiles, taps, splits).
// One split: the CTA owns its dw tile and adds to it directly.  Several splits: each writes its fp64 partial tile to the
// workspace and wgrad_reduce_kernel sums them in split order (no atomics -> bit-reproducible).
__global__ void __launch_bounds__(256) wgrad_simt_kernel(const WgArgs a) {
    __shared__ float As[TK][TM + 4];   // [pixel][c]
    __shared__ float Bs[TK][TN + 4];   // [pixel][n]
    const int tid = threadIdx.x;
    const int c0 = (blockIdx.x % a.ctiles) * TM;
    const int n0 = (blockIdx.x / a.ctiles) * TN;
    const int tap = blockIdx.y;
    const int tj = tap % a.kw, ti = (tap / a.kw) % a.kh, ta = tap / (a.kw * a.kh);
    const long long p_begin = (long long)blockIdx.z * a.pix_per_split;
    long long p_end = p_begin + a.pix_per_split;
    if (p_end > a.npix) p_end = a.npix;
    const int tx = tid & 15, ty = tid >> 4;
    const int lk = tid >> 4, l4 = (tid & 15) * 4;   // loader: pixel row lk, 4 consecutive channels l4..
    double acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
    for (long long p0 = p_begin; p0 < p_end; p0 += TK) {
        const long long pix = p0 + lk;
        float4 av = make_float4(0.f, 0.f, 0.f, 0.f), bv = av;
        if (pix < p_end) {
            long long r = pix;
            const int pw = (int)(r % a.W); r /= a.W;
            const int ph = (int)(r % a.H); r /= a.H;
            const int pt = (int)(r % a.To); const int pb = (int)(r / a.To);
            const int it = pt + ta - a.pad_t, ih = ph + ti - a.pad_h, iw = pw + tj - a.pad_w;
            if (it >= 0 && it < a.T && ih >= 0 && ih < a.H && iw >= 0 && iw < a.W && c0 + l4 < a.C)
                av = *reinterpret_cast<const float4*>(a.x + (long long)pb * a.x_bstride + (long long)it * a.x_tstride + (long long)ih * a.x_hstride + (long long)iw * a.x_cstride + c0 + l4);
            if (n0 + l4 < a.N) bv = *reinterpret_cast<const float4*>(a.dy + (long long)pb * a.dy_bstride + (long long)pt * a.dy_tstride + (long long)ph * a.dy_hstride + (long long)pw * a.dy_cstride + n0 + l4);
        }
        *reinterpret_cast<float4*>(&As[lk][l4]) = av;
        *reinterpret_cast<float4*>(&Bs[lk][l4]) = bv;
        __syncthreads();
#pragma unroll
        for (int k = 0; k < TK; ++k) {
            const float4 a4 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
            const float4 b4 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
            const double ar[4] = {a4.x, a4.y, a4.z, a4.w};
            const double br[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fma(ar[i], br[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int c = c0 + ty * 4 + i;
        if (c >= a.C) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n >= a.N) continue;
            const long long e = ((long long)tap * a.C + c) * a.N + n;
            if (a.partial) a.partial[(long long)blockIdx.z * a.total + e] = acc[i][j];
            else a.dw[e] = (float)((double)a.dw[e] + acc[i][j]);
        }
    }
}

__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const double* __restrict__ partial, float* dw, long long total, int splits) {
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        double s = 0.0;
        for (int z = 0; z < splits; ++z) s += partial[(long long)z * total + e];
        dw[e] = (float)((double)dw[e] + s);
    }
}

}  // namespace

extern "C" int sfvos_conv_simt(const sfvos_conv_params* p, sfvos_stream stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    SF_CHECK(p != nullptr, "conv_simt: null params");
    SF_CHECK(p->C % 4 == 0 && p->x_cstride % 4 == 0 && p->N % 4 == 0, "conv_simt: C, x_cstride, N must be multiples of 4");
    SF_CHECK(p->sum == nullptr && p->sumsq == nullptr, "conv_simt: fused statistics are umma-only; call sfvos_channel_stats");
    SF_CHECK(!(p->accumulate && p->y_dtype != SFVOS_F32), "conv_simt: accumulate needs an f32 output");
    SF_CHECK(p->addend == nullptr, "conv_simt: addend is umma-only (use accumulate)");
    int rc = sfvos_device_check();
    if (rc) return rc;
    SimtArgs a;
    a.x = reinterpret_cast<const float*>(p->x); a.w = reinterpret_cast<const float*>(p->w);
    a.yf = p->y_dtype == SFVOS_F32 ? reinterpret_cast<float*>(p->y) : nullptr;
    a.yb = p->y_dtype == SFVOS_BF16 ? reinterpret_cast<__nv_bfloat16*>(p->y) : nullptr;
    a.B = (int)p->B; a.T = (int)p->T; a.H = (int)p->H; a.W = (int)p->W; a.C = (int)p->C; a.To = (int)p->To;
    a.N = (int)p->N; a.Cp = (int)p->Cp;
    a.npix = p->B * p->To * p->H * p->W;
    a.x_cstride = p->x_cstride; a.y_cstride = p->y_cstride;
    a.x_hstride = p->x_hstride ? p->x_hstride : p->x_cstride * p->W;
    a.x_tstride = p->x_tstride ? p->x_tstride : a.x_hstride * p->H;
    a.x_bstride = p->x_bstride ? p->x_bstride : a.x_tstride * p->T;
    a.kt = (int)p->kt; a.kh = (int)p->kh; a.kw = (int)p->kw;
    a.pad_t = (int)p->pad_t; a.pad_h = (int)p->pad_h; a.pad_w = (int)p->pad_w;
    a.scale = p->scale; a.shift = p->shift; a.relu = p->relu; a.accumulate = p->accumulate;
    a.OH = (int)(p->OH ? p->OH : p->H); a.OW = (int)(p->OW ? p->OW : p->W);
    a.oy_mul = (int)(p->oy_mul ? p->oy_mul : 1); a.ox_mul = (int)(p->ox_mul ? p->ox_mul : 1);
    a.oy_off = (int)p->oy_off; a.ox_off = (int)p->ox_off;
    if (a.npix == 0) return SFVOS_OK;
    dim3 grid((unsigned)((a.npix + TM - 1) / TM), (unsigned)((a.N + TN - 1) / TN));
    conv_simt_kernel<<<grid, 256, 0, stream>>>(a);
    sfvos_set_kernel("conv_simt");
    SF_LAUNCH_CHECK();
    return SFVOS_OK;
}

// split-K factor that fills the machine a few times over (the same on every call with the same shape)
static long long wgrad_simt_splits(const sfvos_wgrad_params* p) {
    const long long npix = p->B * p->To * p->H * p->W;
    const long long base = ((p->C + TM - 1) / TM) * ((p->N + TN - 1) / TN) * (p->kt * p->kh * p->kw);
    if (npix <= 0 || base <= 0) return 1;
    long long splits = (8LL * sfvos_num_sms() + base - 1) / base;
    const long long max_splits = (npix + 255) / 256;
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    if (splits > 65535) splits = 65535;
    return splits;
}

extern "C" int64_t sfvos_wgrad_simt_workspace_bytes(const sfvos_wgrad_params* p) {
    if (p == nullptr) return 0;
    const long long splits = wgrad_simt_splits(p);
    return splits > 1 ? splits * (p->kt * p->kh * p->kw) * p->C * p->N * (long long)sizeof(double) : 0;
}

extern "C" int sfvos_wgrad_simt(const sfvos_wgrad_params* p, sfvos_stream stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    SF_CHECK(p != nullptr, "wgrad_simt: null params");
    SF_CHECK(p->C % 4 == 0 && p->x_cstride % 4 == 0 && p->N % 4 == 0 && p->dy_cstride % 4 == 0,
             "wgrad_simt: C, N and strides must be multiples of 4");
    int rc = sfvos_device_check();
    if (rc) return rc;
    WgArgs a;
    a.x = reinterpret_cast<const float*>(p->x); a.dy = reinterpret_cast<const float*>(p->dy); a.dw = p->dw;
    a.B = (int)p->B; a.T = (int)p->T; a.H = (int)p->H; a.W = (int)p->W; a.C = (int)p->C; a.To = (int)p->To; a.N = (int)p->N;
    a.npix = p->B * p->To * p->H * p->W;
    a.x_cstride = p->x_cstride; a.dy_cstride = p->dy_cstride;
    a.x_hstride = p->x_hstride ? p->x_hstride : p->x_cstride * p->W;
    a.x_tstride = p->x_tstride ? p->x_tstride : a.x_hstride * p->H;
    a.x_bstride = p->x_bstride ? p->x_bstride : a.x_tstride * p->T;
    a.dy_hstride = p->dy_hstride ? p->dy_hstride : p->dy_cstride * p->W;
    a.dy_tstride = p->dy_tstride ? p->dy_tstride : a.dy_hstride * p->H;
    a.dy_bstride = p->dy_bstride ? p->dy_bstride : a.dy_tstride * p->To;
    a.kt = (int)p->kt; a.kh = (int)p->kh; a.kw = (int)p->kw;
    a.pad_t = (int)p->pad_t; a.pad_h = (int)p->pad_h; a.pad_w = (int)p->pad_w;
    if (a.npix == 0) return SFVOS_OK;
    a.ctiles = (a.C + TM - 1) / TM;
    const int ntn = (a.N + TN - 1) / TN;
    const int taps = a.kt * a.kh * a.kw;
    long long splits = wgrad_simt_splits(p);
    a.total = (long long)taps * a.C * a.N;
    // split-K needs room for the fp64 partial tiles; without a (large enough) workspace every CTA walks all pixels
    const long long ws_splits = p->workspace ? (long long)(p->workspace_bytes / (a.total * (long long)sizeof(double))) : 0;
    if (splits > ws_splits) splits = ws_splits;
    if (splits < 2) splits = 1;
    SF_CHECK(p->workspace == nullptr || (reinterpret_cast<uintptr_t>(p->workspace) & 7) == 0, "wgrad_simt: workspace must be 8-byte aligned");
    a.pix_per_split = ((a.npix + splits - 1) / splits + TK - 1) / TK * TK;
    splits = (a.npix + a.pix_per_split - 1) / a.pix_per_split;
    a.partial = splits > 1 ? reinterpret_cast<double*>(p->workspace) : nullptr;
    dim3 grid((unsigned)(a.ctiles * ntn), (unsigned)taps, (unsigned)splits);
    wgrad_simt_kernel<<<grid, 256, 0, stream>>>(a);
    SF_LAUNCH_CHECK();
    if (splits > 1) {
        long long g = (a.total + 255) / 256;
        if (g > 8LL * sfvos_num_sms()) g = 8LL * sfvos_num_sms();
        wgrad_reduce_kernel<<<(unsigned)g, 256, 0, stream>>>(a.partial, a.dw, a.total, (int)splits);
        SF_LAUNCH_CHECK();
    }
    sfvos_set_kernel("wgrad_simt");
    return SFVOS_OK;
}
