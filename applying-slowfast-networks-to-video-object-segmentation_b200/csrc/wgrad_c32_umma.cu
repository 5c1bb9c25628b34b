// Weight gradient of the narrow fast-pathway convolutions (fast_conv2 / fast_conv3: Cin = Cout = 32, k_t x 3 x 3;
// code/helpers/model.py:53-54,59-60), tcgen05 / sm_100a.
//
//     dw[ta][di][dj][c][n] = sum_{b,t,h,w} x[b, t+ta-pad_t, h+di-1, w+dj-1, c] * dy[b,t,h,w,n]
//
// A 32 x 32 GEMM per tap cannot fill a 128-row tensor-core tile, so taps are stacked along BOTH output dims and every
// operand byte is fetched once:
//   * K tile = 16 x 4 dy pixels; one MMA K step = one 16-pixel line.
//   * A = x box of frame tau: the 18 x 6 neighbourhood (tile + halo) x 32 ch = 108 rows of 64 B (64B swizzle).  Seen as
//     an MN-major operand with leading byte offset 64 B (= ONE ROW), its four 32-channel M-atom columns are the same
//     16 rows shifted by 0,1,2,3 pixels: M rows [32 dj, 32 dj + 32) = tap dj (rows 96..127 are junk, never stored);
//     the start row (hl + di) * 18 selects the tap line di.
//   * B = the dy tiles of ALL output frames of this pixel tile, resident side by side (4 KB apart = the leading byte
//     offset of B): x frame tau meets dy frames t = tau+pad_t-ta, so N-atom column i = frame t_lo + i = temporal tap
//     ta_hi - i, and D column block (di*G + G-1-ta_rel) collects tap (ta, di, .) over the whole run.
//   => one 128 x (32 nt) x 16 MMA per (x frame, di, line) replaces 3 nt separate 32 x 32 GEMM steps, each x frame is
//      read once (not 9 k_t times) and each dy tile once per pixel tile.
//   * work item = (temporal tap group of <= 5 taps, range of pixel tiles); partial sums merged with vector atomics.
#include <stdlib.h>

#include "common.cuh"

namespace {

constexpr int PW = 16, PH = 4;
constexpr int HW_ = PW + 2, HH_ = PH + 2;
constexpr int X_TX = HW_ * HH_ * 64;                        // 108 rows x 64 B
constexpr int X_BYTES = (X_TX + 4 * 64 + 1023) & ~1023;     // + slack: the junk 4th M-atom column reads 1 row past the box
constexpr int DY_TILE = PW * PH * 64;                       // 4 KB per output frame
constexpr int NUM_THREADS = 192;                            // warp0 TMA, warp1 MMA + TMEM alloc, warps 2..5 epilogue
constexpr int MAX_G = 5;                                    // 3 * 5 * 32 = 480 TMEM columns

struct WcArgs {
    int B, T, To, H, W;
    int tiles_w, tiles_per_frame, ntiles, splits;
    int kt, pad_t, G, ngroups, x_stages, dy_bytes;
    float* dw;
};

__device__ __forceinline__ void tmem_st_zero_32x32(uint32_t taddr) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, "
        "%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};"
        ::"r"(taddr), "r"(0u)
        : "memory");
}

__global__ void __launch_bounds__(NUM_THREADS, 1)
wgrad_c32_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_dy, const WcArgs a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smem_x = smem;
    uint8_t* smem_dy = smem + a.x_stages * X_BYTES;
    uint64_t* x_full = reinterpret_cast<uint64_t*>(smem_dy + 2 * a.dy_bytes);
    uint64_t* x_empty = x_full + a.x_stages;
    uint64_t* dy_full = x_empty + a.x_stages;
    uint64_t* dy_empty = dy_full + 2;
    uint64_t* done_bar = dy_empty + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done_bar + 1);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    const int tg = blockIdx.x % a.ngroups;
    const int split = blockIdx.x / a.ngroups;
    const int ta0 = tg * a.G, gn = min(a.G, a.kt - ta0);
    // x frames meeting a dy frame through a tap of this group: t = tau + pad_t - ta in [0, To)
    const int tau_lo = max(0, ta0 - a.pad_t), tau_hi = min(a.T - 1, a.To - 1 + ta0 + gn - 1 - a.pad_t);
    const int per = (a.ntiles + a.splits - 1) / a.splits;
    const int tile_begin = split * per;
    const int tile_end = min(a.ntiles, tile_begin + per);
    const bool active = tile_end > tile_begin && tau_hi >= tau_lo;

    if (warp == 0 && elect_one()) {
        tma_prefetch_desc(&tmap_x);
        tma_prefetch_desc(&tmap_dy);
    }
    if (warp == 1) {
        if (elect_one()) {
            for (int i = 0; i < a.x_stages; ++i) { mbar_init(&x_full[i], 1); mbar_init(&x_empty[i], 1); }
            for (int i = 0; i < 2; ++i) { mbar_init(&dy_full[i], 1); mbar_init(&dy_empty[i], 1); }
            mbar_init(done_bar, 1);
            fence_mbar_init();
        }
        __syncwarp();
        tmem_alloc(tmem_slot, 512);
    }
    for (int s = 0; s < a.x_stages; ++s) {                 // slack rows behind every x box: finite junk
        uint4* z = reinterpret_cast<uint4*>(smem_x + s * X_BYTES + X_TX);
        for (int i = threadIdx.x; i < (X_BYTES - X_TX) / 16; i += NUM_THREADS) z[i] = make_uint4(0, 0, 0, 0);
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (warp >= 2) {
        // the accumulators live for the whole kernel: clear them once so that every MMA accumulates
        const uint32_t t_addr = tmem_base + (uint32_t((warp & 3) * 32) << 16);
        for (int c0 = 0; c0 < 3 * a.G * 32; c0 += 32) tmem_st_zero_32x32(t_addr + c0);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    if (warp == 0) {
        if (elect_one() && active) {
            int xs = 0, ds = 0;
            uint32_t xphase = 0, dphase = 0;
            for (int tile = tile_begin; tile < tile_end; ++tile) {
                const int b = tile / a.tiles_per_frame;
                const int rem = tile - b * a.tiles_per_frame;
                const int th_i = rem / a.tiles_w;
                const int h0 = th_i * PH, w0 = (rem - th_i * a.tiles_w) * PW;
                mbar_wait(&dy_empty[ds], dphase ^ 1);
                mbar_arrive_expect_tx(&dy_full[ds], a.To * DY_TILE);
                tma_load_5d(smem_dy + ds * a.dy_bytes, &tmap_dy, &dy_full[ds], 0, w0, h0, 0, b);
                if (++ds == 2) { ds = 0; dphase ^= 1; }
                for (int tau = tau_lo; tau <= tau_hi; ++tau) {
                    mbar_wait(&x_empty[xs], xphase ^ 1);
                    mbar_arrive_expect_tx(&x_full[xs], X_TX);
                    tma_load_5d(smem_x + xs * X_BYTES, &tmap_x, &x_full[xs], 0, w0 - 1, h0 - 1, tau, b);
                    if (++xs == a.x_stages) { xs = 0; xphase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (elect_one() && active) {
            int xs = 0, ds = 0;
            uint32_t xphase = 0, dphase = 0;
            const uint32_t idesc0 = umma_idesc_bf16(128, 0, 1, 1);
            const uint64_t adesc0 = umma_smem_desc(0, 64, 512, 4);          // MN-major SW64, M-atom columns ONE ROW apart
            const uint64_t bdesc0 = umma_smem_desc(0, DY_TILE, 512, 4);     // MN-major SW64, N-atom columns one frame apart
            for (int tile = tile_begin; tile < tile_end; ++tile) {
                mbar_wait(&dy_full[ds], dphase);
                const uint32_t dy_addr = smem_u32(smem_dy + ds * a.dy_bytes);
                for (int tau = tau_lo; tau <= tau_hi; ++tau) {
                    // dy frames met through this group's taps, ascending t = descending ta
                    const int t_lo = max(0, tau + a.pad_t - (ta0 + gn - 1));
                    const int t_hi = min(a.To - 1, tau + a.pad_t - ta0);
                    mbar_wait(&x_full[xs], xphase);
                    tc_fence_after();
                    const uint32_t x_addr = smem_u32(smem_x + xs * X_BYTES);
                    for (int t = t_lo; t <= t_hi; t += 8) {                 // <= 8 frames (N <= 256) per MMA
                        const int nt = min(8, t_hi - t + 1);
                        const int ta_rel = tau + a.pad_t - t - ta0;         // tap of N-atom column 0, relative to the group
                        const uint32_t col0 = uint32_t(a.G - 1 - ta_rel) * 32;
                        const uint32_t idesc = idesc0 | (uint32_t(nt * 4) << 17);
#pragma unroll
                        for (int hl = 0; hl < PH; ++hl) {
                            const uint64_t bdesc = bdesc0 + ((dy_addr + t * DY_TILE + hl * PW * 64) >> 4);
#pragma unroll
                            for (int di = 0; di < 3; ++di) {
                                const uint64_t adesc = adesc0 + ((x_addr + (hl + di) * HW_ * 64) >> 4);
                                umma_bf16(tmem_base + di * a.G * 32 + col0, adesc, bdesc, idesc, 1u);
                            }
                        }
                    }
                    umma_commit(&x_empty[xs]);
                    if (++xs == a.x_stages) { xs = 0; xphase ^= 1; }
                }
                umma_commit(&dy_empty[ds]);
                if (++ds == 2) { ds = 0; dphase ^= 1; }
            }
            umma_commit(done_bar);
        }
    } else {
        const int q = warp & 3;                 // TMEM lane quarter = tap dj (quarter 3 = junk rows)
        if (active && q < 3) {
            mbar_wait(done_bar, 0);
            tc_fence_after();
            for (int di = 0; di < 3; ++di)
                for (int r = 0; r < gn; ++r) {
                    const int ta = ta0 + r;
                    uint32_t v[32];
                    tmem_ld_32x32(tmem_base + (uint32_t(q * 32) << 16) + (di * a.G + (a.G - 1 - r)) * 32, v);
                    tmem_ld_wait();
                    float* dst = a.dw + ((long long)((ta * 3 + di) * 3 + q) * 32 + lane) * 32;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        float4 u = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                                               __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
                        atomicAdd(reinterpret_cast<float4*>(dst + 4 * j), u);
                    }
                }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

}  // namespace

int sfvos_wgrad_c32_applicable(const sfvos_wgrad_params* p) {
    const char* e = getenv("SFVOS_WGRAD_C32");
    if (e && atoi(e) == 0) return 0;
    return p->N == 32 && p->C == 32 && p->kh == 3 && p->kw == 3 && p->pad_h == 1 && p->pad_w == 1 && p->To <= 16 &&
           p->kt <= 64;
}

int sfvos_wgrad_c32_launch(const sfvos_wgrad_params* p, cudaStream_t stream) {
    WcArgs a;
    a.B = (int)p->B; a.T = (int)p->T; a.To = (int)p->To; a.H = (int)p->H; a.W = (int)p->W;
    a.tiles_w = (a.W + PW - 1) / PW;
    a.tiles_per_frame = a.tiles_w * ((a.H + PH - 1) / PH);
    a.ntiles = a.B * a.tiles_per_frame;
    a.kt = (int)p->kt; a.pad_t = (int)p->pad_t;
    a.ngroups = (a.kt + MAX_G - 1) / MAX_G;
    a.G = (a.kt + a.ngroups - 1) / a.ngroups;
    a.ngroups = (a.kt + a.G - 1) / a.G;
    int splits = sfvos_num_sms() / a.ngroups;
    if (splits > a.ntiles) splits = a.ntiles;
    if (splits < 1) splits = 1;
    a.splits = splits;
    a.dy_bytes = a.To * DY_TILE;
    a.x_stages = (227 * 1024 - 2048 - 2 * a.dy_bytes) / X_BYTES;
    if (a.x_stages > 12) a.x_stages = 12;
    SF_CHECK(a.x_stages >= 2, "wgrad_c32: not enough shared memory");
    a.dw = p->dw;

    CUtensorMap tx, tdy;
    int rc;
    {
        uint64_t dims[5] = {32, (uint64_t)p->W, (uint64_t)p->H, (uint64_t)p->T, (uint64_t)p->B};
        const uint64_t cs = (uint64_t)p->x_cstride;
        const uint64_t hs = p->x_hstride ? (uint64_t)p->x_hstride : cs * p->W;
        const uint64_t ts = p->x_tstride ? (uint64_t)p->x_tstride : hs * p->H;
        const uint64_t bs = p->x_bstride ? (uint64_t)p->x_bstride : ts * p->T;
        uint64_t str[4] = {cs * 2, hs * 2, ts * 2, bs * 2};
        uint32_t box[5] = {32, HW_, HH_, 1, 1};
        rc = sfvos_make_tmap(&tx, p->x, 5, dims, str, box, 64);
        if (rc) return rc;
    }
    {
        uint64_t dims[5] = {32, (uint64_t)p->W, (uint64_t)p->H, (uint64_t)p->To, (uint64_t)p->B};
        const uint64_t cs = (uint64_t)p->dy_cstride;
        const uint64_t hs = p->dy_hstride ? (uint64_t)p->dy_hstride : cs * p->W;
        const uint64_t ts = p->dy_tstride ? (uint64_t)p->dy_tstride : hs * p->H;
        const uint64_t bs = p->dy_bstride ? (uint64_t)p->dy_bstride : ts * p->To;
        uint64_t str[4] = {cs * 2, hs * 2, ts * 2, bs * 2};
        uint32_t box[5] = {32, PW, PH, (uint32_t)a.To, 1};
        rc = sfvos_make_tmap(&tdy, p->dy, 5, dims, str, box, 64);
        if (rc) return rc;
    }
    const int smem_bytes = a.x_stages * X_BYTES + 2 * a.dy_bytes + 1024 + 1024;
    SF_CUDA(cudaFuncSetAttribute(wgrad_c32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    wgrad_c32_kernel<<<a.ngroups * splits, NUM_THREADS, smem_bytes, stream>>>(tx, tdy, a);
    sfvos_set_kernel("wgrad_c32");
    SF_LAUNCH_CHECK();
    return SFVOS_OK;
}
