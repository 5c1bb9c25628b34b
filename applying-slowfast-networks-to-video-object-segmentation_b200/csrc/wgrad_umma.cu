// Weight-gradient GEMM on tcgen05 tensor cores (sm_100a).
//   dw[tap][c][n] += sum_{pixels} x[pixel + tap offset][c] * dy[pixel][n]
// GEMM view per (tap, 128-channel block of c): D[M=128 c, N] = A^T * B with the PIXELS as the reduction dim.
// Both operands are channels-last activations, i.e. "MN-major" for the tensor core (the channel index is the
// contiguous one), so the kernel uses the UMMA major-ness bits instead of transposing anything:
//   * A = x tile  : 2 TMA boxes {64ch, PW, PH, 1, 1} (PW*PH = 64 pixels) at the tap-shifted coordinate
//   * B = dy tile : ceil(N/64) TMA boxes {64ch, PW, PH, 1, 1}
//   each box lands as 64 rows (pixels) x 128 B with the 128B swizzle = canonical MN-major SW128 atom
//   (8 K-rows x 64 MN elements), SBO = 1024 B between 8-pixel groups, LBO = 8192 B between 64-channel atoms.
// Out-of-image pixels are zero-filled by TMA in both operands, so halo / overhang terms vanish.
// Work item = (tap, channel block, pixel split); split-K partial sums are merged with vector atomics
// (red.global.add.v4.f32) into the f32 dw buffer.
#include <stdlib.h>

#include "common.cuh"

namespace {

constexpr int KPIX = 64;                    // pixels per K step
constexpr int ATOM_BYTES = KPIX * 128;      // one 64-channel x 64-pixel atom column = 8 KB
constexpr int NUM_THREADS = 192;            // warp0 TMA, warp1 MMA + TMEM alloc, warps 2..5 epilogue

struct WgArgs {
    int B, To, H, W, C, N;
    int PW, PH, tiles_w, tiles_h, ntiles, tiles_per_split;
    int kt, kh, kw, pad_t, pad_h, pad_w, mblks, nb_atoms, stages;
    int nblks, Ntot;                            // wide fc layers: N = 256-column blocks of Ntot output channels
    uint32_t idesc, tmem_cols;
    float* dw;
};

__global__ void __launch_bounds__(NUM_THREADS, 1)
wgrad_umma_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_dy,
                  const WgArgs a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int a_bytes = 2 * ATOM_BYTES;
    const int b_bytes = a.nb_atoms * ATOM_BYTES;
    const int stage_bytes = a_bytes + b_bytes;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + a.stages * stage_bytes);
    uint64_t* empty_bar = full_bar + a.stages;
    uint64_t* done_bar = empty_bar + a.stages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done_bar + 1);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    // work item
    // N block fastest: the nblks CTAs that read the same x block run side by side and share it in L2
    const int nblk = blockIdx.x % a.nblks;
    const int tap = (blockIdx.x / a.nblks) % (a.kt * a.kh * a.kw);
    const int mblk = (blockIdx.x / (a.nblks * a.kt * a.kh * a.kw)) % a.mblks;
    const int split = blockIdx.x / (a.kt * a.kh * a.kw * a.mblks * a.nblks);
    const int n_base = nblk * a.N;
    const int tj = tap % a.kw, ti = (tap / a.kw) % a.kh, ta = tap / (a.kw * a.kh);
    const int tile_begin = split * a.tiles_per_split;
    int tile_end = tile_begin + a.tiles_per_split;
    if (tile_end > a.ntiles) tile_end = a.ntiles;
    const int c_base = mblk * 128;
    const bool second_atom = (c_base + 64 < a.C);       // C=32/64: only one 64-channel atom carries data

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
        // rows 64..127 of A are never loaded: keep them zero so D rows 64..127 stay finite (they are not stored)
        for (int s = 0; s < a.stages; ++s) {
            uint4* z = reinterpret_cast<uint4*>(smem + s * stage_bytes + ATOM_BYTES);
            for (int i = threadIdx.x; i < ATOM_BYTES / 16; i += NUM_THREADS) z[i] = make_uint4(0, 0, 0, 0);
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
            const uint32_t tx_bytes = (second_atom ? 2 : 1) * ATOM_BYTES + b_bytes;
            for (int tile = tile_begin; tile < tile_end; ++tile) {
                const int frame = tile / tiles_per_frame;
                const int rem = tile - frame * tiles_per_frame;
                const int th_i = rem / a.tiles_w;
                const int tw_i = rem - th_i * a.tiles_w;
                const int b = frame / a.To, t = frame - b * a.To;
                const int h0 = th_i * a.PH, w0 = tw_i * a.PW;
                mbar_wait(&empty_bar[stage], phase ^ 1);
                uint8_t* sa = smem + stage * stage_bytes;
                mbar_arrive_expect_tx(&full_bar[stage], tx_bytes);
                const int xw = w0 + tj - a.pad_w, xh = h0 + ti - a.pad_h, xt = t + ta - a.pad_t;
                tma_load_5d(sa, &tmap_x, &full_bar[stage], c_base, xw, xh, xt, b);
                if (second_atom) tma_load_5d(sa + ATOM_BYTES, &tmap_x, &full_bar[stage], c_base + 64, xw, xh, xt, b);
                for (int nb = 0; nb < a.nb_atoms; ++nb)
                    tma_load_5d(sa + a_bytes + nb * ATOM_BYTES, &tmap_dy, &full_bar[stage], n_base + nb * 64, w0, h0, t, b);
                if (++stage == a.stages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            uint32_t accum = 0;
            for (int tile = tile_begin; tile < tile_end; ++tile) {
                mbar_wait(&full_bar[stage], phase);
                tc_fence_after();
                const uint32_t a_addr = smem_u32(smem + stage * stage_bytes);
                const uint32_t b_addr = a_addr + a_bytes;
#pragma unroll
                for (int k = 0; k < KPIX / 16; ++k) {
                    // 16 pixels (K) per MMA = two 8-row groups = 2048 B further into every atom
                    const uint64_t adesc = umma_smem_desc(a_addr + k * 2048, ATOM_BYTES, 1024, 2);
                    const uint64_t bdesc = umma_smem_desc(b_addr + k * 2048, ATOM_BYTES, 1024, 2);
                    umma_bf16(tmem_base, adesc, bdesc, a.idesc, accum);
                    accum = 1;
                }
                umma_commit(&empty_bar[stage]);
                if (++stage == a.stages) { stage = 0; phase ^= 1; }
            }
            umma_commit(done_bar);
        }
    } else {
        // epilogue: warps 2..5 -> TMEM lane quarter = warp id % 4
        const int q = warp & 3;
        const int r = q * 32 + lane;                 // channel within the block
        const int c = c_base + r;
        if (tile_end > tile_begin) {
            mbar_wait(done_bar, 0);
            tc_fence_after();
            float* dst_row = a.dw + ((long long)tap * a.C + c) * a.Ntot + n_base;
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

int sfvos_wgrad_stack_launch(const sfvos_wgrad_params* p, cudaStream_t stream);   // wgrad_stack_umma.cu
int sfvos_wgrad_halo_applicable(const sfvos_wgrad_params* p);                     // wgrad_halo_umma.cu
int sfvos_wgrad_halo_launch(const sfvos_wgrad_params* p, cudaStream_t stream);
int sfvos_wgrad_pair_applicable(const sfvos_wgrad_params* p);                     // wgrad_pair_umma.cu
int sfvos_wgrad_pair_launch(const sfvos_wgrad_params* p, cudaStream_t stream);
int sfvos_wgrad_c32_applicable(const sfvos_wgrad_params* p);                      // wgrad_c32_umma.cu
int sfvos_wgrad_c32_launch(const sfvos_wgrad_params* p, cudaStream_t stream);

extern "C" int sfvos_wgrad_umma(const sfvos_wgrad_params* p, sfvos_stream stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    SF_CHECK(p != nullptr, "wgrad_umma: null params");
    SF_CHECK(p->N >= 32 && p->N % 32 == 0 && (p->N <= 256 || p->N % 256 == 0),
             "wgrad_umma: N=%lld must be a multiple of 32 in [32,256] or a multiple of 256", (long long)p->N);
    SF_CHECK(p->C % 8 == 0 && p->x_cstride % 8 == 0 && p->dy_cstride % 8 == 0, "wgrad_umma: C and strides must be multiples of 8");
    SF_CHECK((p->N % 4) == 0 && (reinterpret_cast<uintptr_t>(p->dw) & 15) == 0, "wgrad_umma: dw must be 16-byte aligned");
    SF_CHECK(p->B > 0 && p->To > 0 && p->H > 0 && p->W > 0, "wgrad_umma: empty tensor");
    int rc = sfvos_device_check();
    if (rc) return rc;
    if (sfvos_wgrad_pair_applicable(p)) return sfvos_wgrad_pair_launch(p, stream);
    if (sfvos_wgrad_c32_applicable(p)) return sfvos_wgrad_c32_launch(p, stream);
    if (sfvos_wgrad_halo_applicable(p)) return sfvos_wgrad_halo_launch(p, stream);
    if (p->N == 32 && getenv("SFVOS_NO_WGRAD_STACK") == nullptr)
        return sfvos_wgrad_stack_launch(p, stream);     // narrow fast-pathway GEMMs: taps stacked along N

    WgArgs a;
    a.B = (int)p->B; a.To = (int)p->To; a.H = (int)p->H; a.W = (int)p->W; a.C = (int)p->C;
    a.Ntot = (int)p->N;
    a.nblks = p->N > 256 ? (int)(p->N / 256) : 1;
    a.N = p->N > 256 ? 256 : (int)p->N;
    choose_ktile(a.H, a.W, &a.PW, &a.PH);
    a.tiles_w = (a.W + a.PW - 1) / a.PW;
    a.tiles_h = (a.H + a.PH - 1) / a.PH;
    a.ntiles = a.B * a.To * a.tiles_w * a.tiles_h;
    a.kt = (int)p->kt; a.kh = (int)p->kh; a.kw = (int)p->kw;
    a.pad_t = (int)p->pad_t; a.pad_h = (int)p->pad_h; a.pad_w = (int)p->pad_w;
    a.mblks = (a.C + 127) / 128;
    a.nb_atoms = (a.N + 63) / 64;
    const int taps = a.kt * a.kh * a.kw;
    const int base_items = taps * a.mblks * a.nblks;
    int splits = (4 * sfvos_num_sms()) / base_items;          // 1 CTA per SM: keep the grid within whole waves
    int max_splits = (a.ntiles + 7) / 8;
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    a.tiles_per_split = (a.ntiles + splits - 1) / splits;
    splits = (a.ntiles + a.tiles_per_split - 1) / a.tiles_per_split;
    const int stage_bytes = (2 + a.nb_atoms) * ATOM_BYTES;
    a.stages = (227 * 1024 - 2048) / stage_bytes;
    if (a.stages > 6) a.stages = 6;
    a.idesc = umma_idesc_bf16(128, a.N, 1, 1);
    uint32_t cols = 32;
    while (cols < (uint32_t)a.N) cols <<= 1;
    a.tmem_cols = cols;
    a.dw = p->dw;

    CUtensorMap tx, tdy;
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
    const int smem_bytes = a.stages * stage_bytes + 1024 + 1024;
    SF_CUDA(cudaFuncSetAttribute(wgrad_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    const int grid = base_items * splits;
    wgrad_umma_kernel<<<grid, NUM_THREADS, smem_bytes, stream>>>(tx, tdy, a);
    sfvos_set_kernel("wgrad_umma");
    SF_LAUNCH_CHECK();
    return SFVOS_OK;
}
