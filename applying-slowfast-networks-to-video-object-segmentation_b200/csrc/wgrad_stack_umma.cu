// Weight-gradient GEMM for the NARROW (Cout = 32) fast-pathway convolutions, tcgen05 / sm_100a.
//
// wgrad_umma.cu computes one tap per CTA: D[128 cin, 32] per 64-pixel K step ingests 24 KB for 4 tiny MMAs, which is
// bound by the ~64 B/clk/SM L2->SMEM fill rate at ~15 % tensor utilisation (measured: 195 TFLOP/s on fast_conv1).
// Here the activation tile is fetched ONCE per pixel block and reused for a whole GROUP of taps:
//     dw[tap][c][n] = sum_q x[q][c] * dy[q - delta_tap][n]          (q runs over INPUT pixels)
// A = x tile (fixed), B = the dy tiles of up to 14 taps, shifted by -delta_tap and stacked along N, so one
// 128 x (32*taps) x 16 MMA pair replaces up to 14 separate 128x32x16 MMAs and the bytes ingested per FLOP drop ~5x.
// dy tiles are 32-channel MN-major atoms (64-byte rows, 64B swizzle, 4 KB per tap per K step); out-of-range
// shifted coordinates are zero-filled by TMA, which implements the valid-range limits of the sum.
#include "common.cuh"

namespace {

constexpr int KPIX = 64;
constexpr int A_ATOM = KPIX * 128;          // 64 channels x 64 pixels, SW128
constexpr int B_ATOM = KPIX * 64;           // 32 channels x 64 pixels, SW64
constexpr int MAX_TG = 14;                  // taps per group: 14 * 32 = 448 TMEM columns
constexpr int NUM_THREADS = 192;

struct WsArgs {
    int B, T, To, H, W, C;
    int PW, PH, tiles_w, tiles_h, ntiles, tiles_per_split;
    int kt, kh, kw, pad_t, pad_h, pad_w, taps, tg, groups, mblks, stages;
    uint32_t tmem_cols;
    float* dw;
};

__global__ void __launch_bounds__(NUM_THREADS, 1)
wgrad_stack_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_dy, const WsArgs a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int a_bytes = 2 * A_ATOM;
    const int stage_bytes = a_bytes + a.tg * B_ATOM;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + a.stages * stage_bytes);
    uint64_t* empty_bar = full_bar + a.stages;
    uint64_t* done_bar = empty_bar + a.stages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done_bar + 1);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    const int group = blockIdx.x % a.groups;
    const int mblk = (blockIdx.x / a.groups) % a.mblks;
    const int split = blockIdx.x / (a.groups * a.mblks);
    const int tap0 = group * a.tg;
    const int ntap = min(a.tg, a.taps - tap0);
    const int tile_begin = split * a.tiles_per_split;
    int tile_end = tile_begin + a.tiles_per_split;
    if (tile_end > a.ntiles) tile_end = a.ntiles;
    const int c_base = mblk * 128;
    const bool second_atom = (c_base + 64 < a.C);

    if (warp == 0 && elect_one()) {
        tma_prefetch_desc(&tmap_x);
        tma_prefetch_desc(&tmap_dy);
    }
    if (warp == 1) {
        if (elect_one()) {
            for (int i = 0; i < a.stages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
            mbar_init(done_bar, 1);
            fence_mbar_init();
        }
        __syncwarp();
        tmem_alloc(tmem_slot, a.tmem_cols);
    }
    if (!second_atom) {
        for (int s = 0; s < a.stages; ++s) {
            uint4* z = reinterpret_cast<uint4*>(smem + s * stage_bytes + A_ATOM);
            for (int i = threadIdx.x; i < A_ATOM / 16; i += NUM_THREADS) z[i] = make_uint4(0, 0, 0, 0);
        }
        fence_proxy_async();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int tiles_per_frame = a.tiles_w * a.tiles_h;

    if (warp == 0) {
        if (elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            const uint32_t tx_bytes = (second_atom ? 2 : 1) * A_ATOM + ntap * B_ATOM;
            for (int tile = tile_begin; tile < tile_end; ++tile) {
                const int frame = tile / tiles_per_frame;          // frame of the INPUT tensor: b * T + t
                const int rem = tile - frame * tiles_per_frame;
                const int th_i = rem / a.tiles_w;
                const int tw_i = rem - th_i * a.tiles_w;
                const int b = frame / a.T, t = frame - b * a.T;
                const int h0 = th_i * a.PH, w0 = tw_i * a.PW;
                mbar_wait(&empty_bar[stage], phase ^ 1);
                uint8_t* sa = smem + stage * stage_bytes;
                mbar_arrive_expect_tx(&full_bar[stage], tx_bytes);
                tma_load_5d(sa, &tmap_x, &full_bar[stage], c_base, w0, h0, t, b);
                if (second_atom) tma_load_5d(sa + A_ATOM, &tmap_x, &full_bar[stage], c_base + 64, w0, h0, t, b);
                for (int n = 0; n < ntap; ++n) {
                    const int tap = tap0 + n;
                    const int tj = tap % a.kw, ti = (tap / a.kw) % a.kh, ta = tap / (a.kw * a.kh);
                    tma_load_5d(sa + a_bytes + n * B_ATOM, &tmap_dy, &full_bar[stage], 0, w0 - (tj - a.pad_w),
                                h0 - (ti - a.pad_h), t - (ta - a.pad_t), b);
                }
                if (++stage == a.stages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            uint32_t accum = 0;
            const int n_total = ntap * 32;
            const int n0 = n_total > 256 ? 256 : n_total;          // first MMA covers taps [0,8)
            const int n1 = n_total - n0;                           // second MMA the rest (multiple of 32)
            const uint32_t idesc0 = umma_idesc_bf16(128, n0, 1, 1);
            const uint32_t idesc1 = umma_idesc_bf16(128, n1 > 0 ? n1 : 32, 1, 1);
            for (int tile = tile_begin; tile < tile_end; ++tile) {
                mbar_wait(&full_bar[stage], phase);
                tc_fence_after();
                const uint32_t a_addr = smem_u32(smem + stage * stage_bytes);
                const uint32_t b_addr = a_addr + a_bytes;
#pragma unroll
                for (int k = 0; k < KPIX / 16; ++k) {
                    const uint64_t adesc = umma_smem_desc(a_addr + k * 2048, A_ATOM, 1024, 2);
                    const uint64_t bdesc0 = umma_smem_desc(b_addr + k * 1024, B_ATOM, 512, 4);
                    umma_bf16(tmem_base, adesc, bdesc0, idesc0, accum);
                    if (n1 > 0) {
                        const uint64_t bdesc1 = umma_smem_desc(b_addr + 8 * B_ATOM + k * 1024, B_ATOM, 512, 4);
                        umma_bf16(tmem_base + 256, adesc, bdesc1, idesc1, accum);
                    }
                    accum = 1;
                }
                umma_commit(&empty_bar[stage]);
                if (++stage == a.stages) { stage = 0; phase ^= 1; }
            }
            umma_commit(done_bar);
        }
    } else {
        const int q = warp & 3;
        const int c = c_base + q * 32 + lane;
        if (tile_end > tile_begin) {
            mbar_wait(done_bar, 0);
            tc_fence_after();
            for (int n = 0; n < ntap; ++n) {
                uint32_t v[32];
                tmem_ld_32x32(tmem_base + (uint32_t(q * 32) << 16) + n * 32, v);
                tmem_ld_wait();
                if (c < a.C) {
                    float* dst = a.dw + ((long long)(tap0 + n) * a.C + c) * 32;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        float4 u = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                                               __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
                        atomicAdd(reinterpret_cast<float4*>(dst + 4 * j), u);
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, a.tmem_cols);
    }
}

void choose_ktile(int H, int W, int* PW, int* PH) {
    double best = -1.0;
    int bw = 64, bh = 1;
    for (int pw = 1; pw <= KPIX; pw <<= 1) {
        const int ph = KPIX / pw;
        long long tiles = (long long)((W + pw - 1) / pw) * ((H + ph - 1) / ph);
        double eff = (double)H * W / (double)(tiles * KPIX);
        if (eff > best + 1e-9) { best = eff; bw = pw; bh = ph; }
    }
    *PW = bw; *PH = bh;
}

}  // namespace

// Called by sfvos_wgrad_umma (wgrad_umma.cu) when N == 32 and dy is dense in its channel axis.
int sfvos_wgrad_stack_launch(const sfvos_wgrad_params* p, cudaStream_t stream) {
    WsArgs a;
    a.B = (int)p->B; a.T = (int)p->T; a.To = (int)p->To; a.H = (int)p->H; a.W = (int)p->W; a.C = (int)p->C;
    choose_ktile(a.H, a.W, &a.PW, &a.PH);
    a.tiles_w = (a.W + a.PW - 1) / a.PW;
    a.tiles_h = (a.H + a.PH - 1) / a.PH;
    a.ntiles = a.B * a.T * a.tiles_w * a.tiles_h;             // blocks over the INPUT domain
    a.kt = (int)p->kt; a.kh = (int)p->kh; a.kw = (int)p->kw;
    a.pad_t = (int)p->pad_t; a.pad_h = (int)p->pad_h; a.pad_w = (int)p->pad_w;
    a.taps = a.kt * a.kh * a.kw;
    a.groups = (a.taps + MAX_TG - 1) / MAX_TG;
    a.tg = (a.taps + a.groups - 1) / a.groups;                // balanced groups, <= 14 taps each
    a.mblks = (a.C + 127) / 128;
    const int base_items = a.groups * a.mblks;
    int splits = (2 * sfvos_num_sms()) / base_items;          // 1 CTA per SM: keep the grid within whole waves
    int max_splits = (a.ntiles + 7) / 8;
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    a.tiles_per_split = (a.ntiles + splits - 1) / splits;
    splits = (a.ntiles + a.tiles_per_split - 1) / a.tiles_per_split;
    const int stage_bytes = 2 * A_ATOM + a.tg * B_ATOM;
    a.stages = (227 * 1024 - 2048) / stage_bytes;
    if (a.stages > 6) a.stages = 6;
    SF_CHECK(a.stages >= 2, "wgrad_stack: not enough shared memory");
    uint32_t cols = 32;
    while (cols < (uint32_t)(a.tg * 32)) cols <<= 1;
    a.tmem_cols = cols;
    a.dw = p->dw;

    CUtensorMap tx, tdy;
    int rc;
    {
        uint64_t dims[5] = {(uint64_t)p->C, (uint64_t)p->W, (uint64_t)p->H, (uint64_t)p->T, (uint64_t)p->B};
        const uint64_t cs = (uint64_t)p->x_cstride;
        const uint64_t hs = p->x_hstride ? (uint64_t)p->x_hstride : cs * p->W;
        const uint64_t ts = p->x_tstride ? (uint64_t)p->x_tstride : hs * p->H;
        const uint64_t bs = p->x_bstride ? (uint64_t)p->x_bstride : ts * p->T;
        uint64_t str[4] = {cs * 2, hs * 2, ts * 2, bs * 2};
        uint32_t box[5] = {64, (uint32_t)a.PW, (uint32_t)a.PH, 1, 1};
        rc = sfvos_make_tmap(&tx, p->x, 5, dims, str, box, 128);
        if (rc) return rc;
    }
    {
        uint64_t dims[5] = {32, (uint64_t)p->W, (uint64_t)p->H, (uint64_t)p->To, (uint64_t)p->B};
        const uint64_t cs = (uint64_t)p->dy_cstride;
        const uint64_t hs = p->dy_hstride ? (uint64_t)p->dy_hstride : cs * p->W;
        const uint64_t ts = p->dy_tstride ? (uint64_t)p->dy_tstride : hs * p->H;
        const uint64_t bs = p->dy_bstride ? (uint64_t)p->dy_bstride : ts * p->To;
        uint64_t str[4] = {cs * 2, hs * 2, ts * 2, bs * 2};
        uint32_t box[5] = {32, (uint32_t)a.PW, (uint32_t)a.PH, 1, 1};
        rc = sfvos_make_tmap(&tdy, p->dy, 5, dims, str, box, 64);
        if (rc) return rc;
    }
    const int smem_bytes = a.stages * stage_bytes + 1024 + 1024;
    SF_CUDA(cudaFuncSetAttribute(wgrad_stack_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    wgrad_stack_kernel<<<base_items * splits, NUM_THREADS, smem_bytes, stream>>>(tx, tdy, a);
    sfvos_set_kernel("wgrad_stack");
    SF_LAUNCH_CHECK();
    return SFVOS_OK;
}
