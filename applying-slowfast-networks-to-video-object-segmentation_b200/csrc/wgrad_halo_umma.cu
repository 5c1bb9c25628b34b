// Weight gradient of the wide-input / narrow-output fast-pathway convolution (fast_conv1: Cin = 256 -> Cout = 32,
// k_t x 3 x 3; code/helpers/model.py:47-48), tcgen05 / sm_100a.
//
//     dw[ta][di][dj][c][n] = sum_{b,tau,h,w} x[b,tau,h,w,c] * dy[b, tau-ta+pad_t, h-di+1, w-dj+1, n]
//
// GEMM per (ta, 128-channel block): D[128 c, 9 taps x 32] with the x PIXELS as the reduction dim.
//   * A = x tile, 16 x 4 pixels x 128 ch, MN-major SW128 atoms (2 TMA boxes), unshifted
//   * B = ONE dy box per K tile: the 18 x 6 pixel neighbourhood (tile + halo) x 32 ch, 64-byte rows, 64B swizzle.
//     All 9 spatial taps read that buffer through shifted descriptors: a K step is one 16-pixel tile line, so the rows
//     of tap (di,dj) are 16 CONSECUTIVE halo rows starting at line (hl-di+2), pixel (2-dj); and because consecutive
//     dj differ by exactly one row (64 B), the three dj taps are the three N-atom columns of one MN-major operand with
//     leading byte offset 64 B -> one 128 x 96 x 16 MMA per (di, line) instead of nine 128 x 32 x 16 ones.
//     (wgrad_stack_umma.cu fetched the shifted dy tile once per tap: 14 x 64 rows per K tile against 108 here; that
//     kernel was bound by the TMA row rate.)
//   * work item = (ta, channel block, pixel split); split-K partial sums merged with vector atomics.
#include <stdlib.h>

#include "common.cuh"

namespace {

constexpr int PW = 16, PH = 4;                  // K tile = 64 pixels; one MMA K step = one 16-pixel line
constexpr int HW_ = PW + 2, HH_ = PH + 2;       // dy halo box
constexpr int A_ATOM = PW * PH * 128;           // 64 px x 64 ch, SW128
constexpr int A_BYTES = 2 * A_ATOM;
constexpr int B_TX = HW_ * HH_ * 64;            // 108 rows x 64 B
constexpr int B_BYTES = (B_TX + 4 * 64 + 1023) & ~1023;     // + slack rows so no descriptor ever points past the stage
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int NUM_THREADS = 192;                // warp0 TMA, warp1 MMA + TMEM alloc, warps 2..5 epilogue
constexpr int NCOLS = 9 * 32;

struct WhArgs {
    int B, T, To, H, W, C;
    int tiles_w, tiles_h, kt, pad_t, mblks, stages, splits;
    float* dw;
};

__global__ void __launch_bounds__(NUM_THREADS, 1)
wgrad_halo_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_dy, const WhArgs a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + a.stages * STAGE_BYTES);
    uint64_t* empty_bar = full_bar + a.stages;
    uint64_t* done_bar = empty_bar + a.stages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done_bar + 1);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    const int ta = blockIdx.x % a.kt;
    const int mblk = (blockIdx.x / a.kt) % a.mblks;
    const int split = blockIdx.x / (a.kt * a.mblks);
    const int c_base = mblk * 128;
    const bool second_atom = (c_base + 64 < a.C);
    // x frames that pair with a dy frame for this temporal tap: t = tau - ta + pad_t in [0, To)
    const int tau_lo = max(0, ta - a.pad_t), tau_hi = min(a.T - 1, a.To - 1 + ta - a.pad_t);
    const int nfr = max(0, tau_hi - tau_lo + 1);
    const int tiles_per_frame = a.tiles_w * a.tiles_h;
    const int ntiles = a.B * nfr * tiles_per_frame;
    const int per = (ntiles + a.splits - 1) / a.splits;
    const int tile_begin = split * per;
    const int tile_end = min(ntiles, tile_begin + per);

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
        tmem_alloc(tmem_slot, 512);
    }
    // zero what TMA never writes: the slack rows behind every dy halo and (C <= 64) the second A atom
    for (int s = 0; s < a.stages; ++s) {
        uint4* z = reinterpret_cast<uint4*>(smem + s * STAGE_BYTES + A_BYTES + B_TX);
        for (int i = threadIdx.x; i < (B_BYTES - B_TX) / 16; i += NUM_THREADS) z[i] = make_uint4(0, 0, 0, 0);
        if (!second_atom) {
            uint4* z2 = reinterpret_cast<uint4*>(smem + s * STAGE_BYTES + A_ATOM);
            for (int i = threadIdx.x; i < A_ATOM / 16; i += NUM_THREADS) z2[i] = make_uint4(0, 0, 0, 0);
        }
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            const uint32_t tx_bytes = (second_atom ? 2 : 1) * A_ATOM + B_TX;
            for (int tile = tile_begin; tile < tile_end; ++tile) {
                const int fr = tile / tiles_per_frame;
                const int rem = tile - fr * tiles_per_frame;
                const int th_i = rem / a.tiles_w;
                const int tw_i = rem - th_i * a.tiles_w;
                const int b = fr / nfr, tau = tau_lo + (fr - b * nfr);
                const int h0 = th_i * PH, w0 = tw_i * PW;
                mbar_wait(&empty_bar[stage], phase ^ 1);
                uint8_t* sa = smem + stage * STAGE_BYTES;
                mbar_arrive_expect_tx(&full_bar[stage], tx_bytes);
                tma_load_5d(sa, &tmap_x, &full_bar[stage], c_base, w0, h0, tau, b);
                if (second_atom) tma_load_5d(sa + A_ATOM, &tmap_x, &full_bar[stage], c_base + 64, w0, h0, tau, b);
                tma_load_5d(sa + A_BYTES, &tmap_dy, &full_bar[stage], 0, w0 - 1, h0 - 1, tau - ta + a.pad_t, b);
                if (++stage == a.stages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            uint32_t accum = 0;
            const uint32_t idesc = umma_idesc_bf16(128, 96, 1, 1);
            const uint64_t adesc0 = umma_smem_desc(0, A_ATOM, 1024, 2);     // MN-major SW128: LBO = atom column, SBO = 8 rows
            const uint64_t bdesc0 = umma_smem_desc(0, 64, 512, 4);          // MN-major SW64: LBO = ONE ROW (next dj tap)
            for (int tile = tile_begin; tile < tile_end; ++tile) {
                mbar_wait(&full_bar[stage], phase);
                tc_fence_after();
                const uint32_t a_addr = smem_u32(smem + stage * STAGE_BYTES);
                const uint32_t b_addr = a_addr + A_BYTES;
#pragma unroll
                for (int hl = 0; hl < PH; ++hl) {
                    const uint64_t adesc = adesc0 + ((a_addr + hl * 2048) >> 4);
#pragma unroll
                    for (int di = 0; di < 3; ++di) {
                        const uint64_t bdesc = bdesc0 + ((b_addr + (hl - di + 2) * HW_ * 64) >> 4);
                        umma_bf16(tmem_base + di * 96, adesc, bdesc, idesc, accum);
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
            for (int blk = 0; blk < 9; ++blk) {
                const int di = blk / 3, dj = 2 - (blk - di * 3);        // N-atom column j of group di holds tap dj = 2 - j
                uint32_t v[32];
                tmem_ld_32x32(tmem_base + (uint32_t(q * 32) << 16) + blk * 32, v);
                tmem_ld_wait();
                if (c < a.C) {
                    float* dst = a.dw + ((long long)((ta * 3 + di) * 3 + dj) * a.C + c) * 32;
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
        tmem_dealloc(tmem_base, 512);
    }
}

}  // namespace

int sfvos_wgrad_halo_applicable(const sfvos_wgrad_params* p) {
    const char* e = getenv("SFVOS_WGRAD_HALO");
    if (e && atoi(e) == 0) return 0;
    return p->N == 32 && p->C >= 64 && p->kh == 3 && p->kw == 3 && p->pad_h == 1 && p->pad_w == 1;
}

int sfvos_wgrad_halo_launch(const sfvos_wgrad_params* p, cudaStream_t stream) {
    WhArgs a;
    a.B = (int)p->B; a.T = (int)p->T; a.To = (int)p->To; a.H = (int)p->H; a.W = (int)p->W; a.C = (int)p->C;
    a.tiles_w = (a.W + PW - 1) / PW;
    a.tiles_h = (a.H + PH - 1) / PH;
    a.kt = (int)p->kt; a.pad_t = (int)p->pad_t;
    a.mblks = (a.C + 127) / 128;
    const int base_items = a.kt * a.mblks;
    const long long ntiles = (long long)a.B * a.To * a.tiles_w * a.tiles_h;      // upper bound per temporal tap
    int splits = sfvos_num_sms() / base_items;                                  // one wave: grid <= #SMs (1 CTA per SM)
    const long long max_splits = (ntiles + 7) / 8;
    if (splits > max_splits) splits = (int)max_splits;
    if (splits < 1) splits = 1;
    a.splits = splits;
    a.stages = (227 * 1024 - 2048) / STAGE_BYTES;
    if (a.stages > 8) a.stages = 8;
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
        uint32_t box[5] = {64, PW, PH, 1, 1};
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
        uint32_t box[5] = {32, HW_, HH_, 1, 1};
        rc = sfvos_make_tmap(&tdy, p->dy, 5, dims, str, box, 64);
        if (rc) return rc;
    }
    const int smem_bytes = a.stages * STAGE_BYTES + 1024 + 1024;
    SF_CUDA(cudaFuncSetAttribute(wgrad_halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    wgrad_halo_kernel<<<base_items * splits, NUM_THREADS, smem_bytes, stream>>>(tx, tdy, a);
    sfvos_set_kernel("wgrad_halo");
    SF_LAUNCH_CHECK();
    return SFVOS_OK;
}
