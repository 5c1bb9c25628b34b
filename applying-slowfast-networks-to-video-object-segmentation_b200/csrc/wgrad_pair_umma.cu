// CTA-pair (tcgen05 cta_group::2) weight-gradient GEMM for the wide layers (C = 256 input channels, N >= 128):
// slow_conv1/2/3 and the mask-head convolutions.
//     dw[tap][c][n] += sum_{pixels} x[pixel + tap offset][c] * dy[pixel][n]
// wgrad_umma.cu gives each CTA one (tap, 128-channel block): per 64-pixel K step it ingests 16 KB of x and 24-32 KB of
// dy for four 128 x N x 16 MMAs, and both the TMA fill and the MMA operand reads (10 KB per MMA) run into the 128 B/clk
// shared-memory port.  Here the two channel blocks of a tap are the two CTAs of a cluster and one 256 x N x 16 MMA:
// each CTA loads its own x block and only HALF of the dy tile (N/2 channels); tcgen05.mma.cta_group::2 reads dy from both.
// Barrier protocol as in conv_pair_umma.cu (leader issues the MMAs, both CTAs' TMA loads complete on the leader's
// "full" barriers, commits are multicast to both CTAs).
#include <stdlib.h>

#include "common.cuh"

namespace {

constexpr int KPIX = 64;
constexpr int ATOM_BYTES = KPIX * 128;      // 64 channels x 64 pixels, SW128
constexpr int A_BYTES = 2 * ATOM_BYTES;     // this CTA's 128 channels of the x tile
constexpr int B_BYTES = 2 * ATOM_BYTES;     // this CTA's N/2 (<= 128) channels of the dy tile
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int NUM_THREADS = 192;            // warp0 TMA, warp1 MMA + TMEM alloc, warps 2..5 epilogue

struct WpArgs {
    int B, To, H, W, C, N;
    int PW, PH, tiles_w, tiles_h, ntiles, tiles_per_split;
    int kt, kh, kw, pad_t, pad_h, pad_w, stages;
    uint32_t idesc, tmem_cols;
    float* dw;
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t map_to_cta(const void* p, uint32_t rank) {
    uint32_t out;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(out) : "r"(smem_u32(p)), "r"(rank));
    return out;
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tma_load_5d_pair(void* dst, const CUtensorMap* m, uint32_t bar_cluster, int c0, int c1, int c2,
                                                 int c3, int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
        : "memory");
}
__device__ __forceinline__ void umma_bf16_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
    asm volatile(
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
        ::"r"(smem_u32(bar)), "h"((uint16_t)3)
        : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
wgrad_pair_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_dy, const WpArgs a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + a.stages * STAGE_BYTES);
    uint64_t* empty_bar = full_bar + a.stages;
    uint64_t* done_bar = empty_bar + a.stages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done_bar + 1);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();          // = channel block of this CTA
    const int taps = a.kt * a.kh * a.kw;
    const int item = blockIdx.x >> 1;
    const int tap = item % taps;
    const int split = item / taps;
    const int tj = tap % a.kw, ti = (tap / a.kw) % a.kh, ta = tap / (a.kw * a.kh);
    const int tile_begin = split * a.tiles_per_split;
    int tile_end = tile_begin + a.tiles_per_split;
    if (tile_end > a.ntiles) tile_end = a.ntiles;
    const int c_base = (int)rank * 128;
    const int n_base = (int)rank * (a.N / 2);

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
        tmem_alloc_pair(tmem_slot, a.tmem_cols);
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int tiles_per_frame = a.tiles_w * a.tiles_h;

    if (warp == 0) {
        if (elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = tile_begin; tile < tile_end; ++tile) {
                const int frame = tile / tiles_per_frame;
                const int rem = tile - frame * tiles_per_frame;
                const int th_i = rem / a.tiles_w;
                const int tw_i = rem - th_i * a.tiles_w;
                const int b = frame / a.To, t = frame - b * a.To;
                const int h0 = th_i * a.PH, w0 = tw_i * a.PW;
                mbar_wait(&empty_bar[stage], phase ^ 1);
                uint8_t* sa = smem + stage * STAGE_BYTES;
                if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * STAGE_BYTES);
                const uint32_t bar = map_to_cta(&full_bar[stage], 0);
                const int xw = w0 + tj - a.pad_w, xh = h0 + ti - a.pad_h, xt = t + ta - a.pad_t;
                tma_load_5d_pair(sa, &tmap_x, bar, c_base, xw, xh, xt, b);
                tma_load_5d_pair(sa + ATOM_BYTES, &tmap_x, bar, c_base + 64, xw, xh, xt, b);
                tma_load_5d_pair(sa + A_BYTES, &tmap_dy, bar, n_base, w0, h0, t, b);
                tma_load_5d_pair(sa + A_BYTES + ATOM_BYTES, &tmap_dy, bar, n_base + 64, w0, h0, t, b);
                if (++stage == a.stages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (rank == 0 && elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            uint32_t accum = 0;
            const uint64_t desc0 = umma_smem_desc(0, ATOM_BYTES, 1024, 2);      // MN-major SW128
            for (int tile = tile_begin; tile < tile_end; ++tile) {
                mbar_wait(&full_bar[stage], phase);
                tc_fence_after();
                const uint32_t a_addr = smem_u32(smem + stage * STAGE_BYTES);
#pragma unroll
                for (int k = 0; k < KPIX / 16; ++k) {
                    const uint64_t adesc = desc0 + ((a_addr + k * 2048) >> 4);
                    const uint64_t bdesc = desc0 + ((a_addr + A_BYTES + k * 2048) >> 4);
                    umma_bf16_pair(tmem_base, adesc, bdesc, a.idesc, accum);
                    accum = 1;
                }
                umma_commit_pair(&empty_bar[stage]);
                if (++stage == a.stages) { stage = 0; phase ^= 1; }
            }
            umma_commit_pair(done_bar);
        }
    } else {
        const int q = warp & 3;
        const int c = c_base + q * 32 + lane;
        if (tile_end > tile_begin) {
            mbar_wait(done_bar, 0);
            tc_fence_after();
            float* dst_row = a.dw + ((long long)tap * a.C + c) * a.N;
            for (int n0 = 0; n0 < a.N; n0 += 32) {
                uint32_t v[32];
                tmem_ld_32x32(tmem_base + (uint32_t(q * 32) << 16) + n0, v);
                tmem_ld_wait();
                if (c < a.C) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        float4 u = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                                               __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
                        atomicAdd(reinterpret_cast<float4*>(dst_row + n0 + 4 * j), u);
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc_pair(tmem_base, a.tmem_cols);
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

int sfvos_wgrad_pair_applicable(const sfvos_wgrad_params* p) {
    const char* e = getenv("SFVOS_WGRAD_PAIR");
    if (e && atoi(e) == 0) return 0;
    return p->C > 128 && p->C <= 256 && p->C % 64 == 0 && p->N >= 128 && p->N <= 256 && p->N % 32 == 0;
}

int sfvos_wgrad_pair_launch(const sfvos_wgrad_params* p, cudaStream_t stream) {
    WpArgs a;
    a.B = (int)p->B; a.To = (int)p->To; a.H = (int)p->H; a.W = (int)p->W; a.C = (int)p->C; a.N = (int)p->N;
    choose_ktile(a.H, a.W, &a.PW, &a.PH);
    a.tiles_w = (a.W + a.PW - 1) / a.PW;
    a.tiles_h = (a.H + a.PH - 1) / a.PH;
    a.ntiles = a.B * a.To * a.tiles_w * a.tiles_h;
    a.kt = (int)p->kt; a.kh = (int)p->kh; a.kw = (int)p->kw;
    a.pad_t = (int)p->pad_t; a.pad_h = (int)p->pad_h; a.pad_w = (int)p->pad_w;
    const int taps = a.kt * a.kh * a.kw;
    const int clusters_per_wave = sfvos_num_sms() / 2;
    int splits = (4 * clusters_per_wave) / taps;              // 1 cluster per SM pair: keep the grid within whole waves
    const int max_splits = (a.ntiles + 7) / 8;
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    a.tiles_per_split = (a.ntiles + splits - 1) / splits;
    splits = (a.ntiles + a.tiles_per_split - 1) / a.tiles_per_split;
    a.stages = (227 * 1024 - 2048) / STAGE_BYTES;
    if (a.stages > 6) a.stages = 6;
    a.idesc = umma_idesc_bf16(256, a.N, 1, 1);
    uint32_t cols = 32;
    while (cols < (uint32_t)a.N) cols <<= 1;
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
        // channels past N read as zero (TMA bounds): the half tile of CTA 1 may reach beyond N when N/2 is not 64 or 128
        uint64_t dims[5] = {(uint64_t)p->N, (uint64_t)p->W, (uint64_t)p->H, (uint64_t)p->To, (uint64_t)p->B};
        const uint64_t cs = (uint64_t)p->dy_cstride;
        const uint64_t hs = p->dy_hstride ? (uint64_t)p->dy_hstride : cs * p->W;
        const uint64_t ts = p->dy_tstride ? (uint64_t)p->dy_tstride : hs * p->H;
        const uint64_t bs = p->dy_bstride ? (uint64_t)p->dy_bstride : ts * p->To;
        uint64_t str[4] = {cs * 2, hs * 2, ts * 2, bs * 2};
        uint32_t box[5] = {64, (uint32_t)a.PW, (uint32_t)a.PH, 1, 1};
        rc = sfvos_make_tmap(&tdy, p->dy, 5, dims, str, box, 128);
        if (rc) return rc;
    }
    const int smem_bytes = a.stages * STAGE_BYTES + 1024 + 1024;
    SF_CUDA(cudaFuncSetAttribute(wgrad_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    wgrad_pair_kernel<<<2 * taps * splits, NUM_THREADS, smem_bytes, stream>>>(tx, tdy, a);
    sfvos_set_kernel("wgrad_pair");
    SF_LAUNCH_CHECK();
    return SFVOS_OK;
}
