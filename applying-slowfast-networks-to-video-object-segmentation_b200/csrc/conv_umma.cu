// Implicit-GEMM convolution on tcgen05 tensor cores (sm_100a): fprop and dgrad of every Conv3d of SlowFastLayers
// (code/helpers/model.py:72-76,83-90) and of the mask-head Conv2d / ConvTranspose2d
// (TV/models/detection/mask_rcnn.py:284-296,342-344).
//
// GEMM view: D[M,N] = A[M,K] * B[K,N] with M = output pixels, N = output channels, K = taps * channels.
//   * one M tile = TW x TH output pixels of one (clip, frame) (<=128 rows; the rest of the 128 UMMA rows are
//     never stored); N is a single tile (<=256) so an activation tile is fetched once for all output channels
//   * A is never materialised.  Two ways to feed it, both pure TMA (out-of-range rows / columns / frames are
//     zero-filled by TMA = the convolution padding):
//       per-tap mode : one 5-D box {BK ch, TW, TH, 1, 1} per (tap, channel chunk) at the shifted coordinate
//       halo mode    : (3x3 spatial taps) ONE box {64 ch, 16, TH+2, 1, 1} per (temporal tap, channel chunk) holding
//                      the tile plus its halo; the 9 spatial taps are 9 UMMA descriptors into that same buffer
//                      (start += (i*16 + j) rows, stride between 8-row groups = one 16-pixel line = 2048 B, swizzle
//                      phase carried in the descriptor's base-offset field).  TMA row requests per tap drop from
//                      128 to 32 -- the L2->SMEM fill rate (~1 128-byte row / 2 clk / SM) is what bounds this kernel.
//     Rows land as 128-byte (BK=64) or 64-byte (BK=32) swizzled rows = the canonical K-major UMMA operand layout.
//   * B = pre-packed bf16 weights [N][taps*Cp], one 2-D TMA box {BK, N} per (tap, chunk), in its own smem ring
//   * fp32 accumulators live in TMEM, double buffered (2*N columns) so the epilogue of tile i overlaps the
//     MMAs of tile i+1; persistent CTAs, one per SM
//   * warp 0 = TMA producer, warp 1 = MMA issuer (one elected thread), warp 2 = TMEM allocator,
//     warps 4..11 = epilogue, two per TMEM lane quarter splitting the accumulator's column blocks (TMEM -> registers ->
//     shared-memory transpose (common.cuh: epi_block) -> per-channel affine/ReLU -> whole-sector stores), which also reduces
//     the per-channel sum / sum-of-squares of the raw fp32 accumulators for train-mode BatchNorm.
#include <math.h>
#include <stdlib.h>

#include "common.cuh"

namespace {

constexpr int BM = 128;
constexpr int NUM_THREADS = 384;             // warps 0-3: TMA / MMA / TMEM alloc / idle; warps 4-11: epilogue
constexpr int EPI_WARP0 = 4;
constexpr int EPI_THREADS = 256;             // TWO warps per TMEM lane quarter, each draining half of the accumulator's columns
constexpr int HALO_LP = 16;                 // pixels per halo line in smem (tile width 8 + 2 halo, padded to 16)

struct ConvArgs {
    int B, To, H, W;
    int TW, TH, tiles_w, tiles_h, ntiles;
    int N, kt, kh, kw, pad_t, pad_h, pad_w, cchunks;
    int nchunks;                                // N tiles of one pixel tile (wide fc layers: Cout = nchunks * 256)
    int b_resident;                             // per-tap mode: ALL weight tiles stay in shared memory for the CTA's lifetime
    int halo, use_bo, a_stages, b_stages, a_stage_bytes, b_group;     // b_group = taps per B stage
    int epi_stage;                              // epilogue through the shared-memory transpose (common.cuh: epi_block)
    uint32_t idesc, tmem_cols, a_tx_bytes;
    void* y;
    int y_bf16, relu, accumulate;
    long long y_cstride;
    const float* scale;
    const float* shift;
    float* sum;
    float* sumsq;
    int OH, OW, oy_mul, oy_off, ox_mul, ox_off;
    const __nv_bfloat16* relu_mask;          // fused ReLU backward (see sfvos_conv_params)
    long long mask_cstride;
    const void* addend;                      // y = act(...) + addend (see sfvos_conv_params)
    long long addend_cstride;
    int addend_bf16;
};

// BK = channels per K step: 64 (128-byte rows, 128B swizzle) or 32 (64-byte rows, 64B swizzle; Cin = 32 layers).
template <int BK>
__global__ void __launch_bounds__(NUM_THREADS, 1)
conv_umma_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
                 const ConvArgs a) {
    constexpr uint32_t LAYOUT = BK == 64 ? 2u : 4u;          // SWIZZLE_128B : SWIZZLE_64B
    constexpr uint32_t ROW = BK * 2;                          // bytes per smem row
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int b_bytes = a.N * BK * 2;                         // one tap's weight tile
    const int b_stage_bytes = a.b_group * b_bytes;
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + a.a_stages * a.a_stage_bytes;
    uint64_t* a_full = reinterpret_cast<uint64_t*>(smem_b + a.b_stages * b_stage_bytes);
    uint64_t* a_empty = a_full + a.a_stages;
    uint64_t* b_full = a_empty + a.a_stages;
    uint64_t* b_empty = b_full + a.b_stages;
    uint64_t* tmem_full = b_empty + a.b_stages;
    uint64_t* tmem_empty = tmem_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
    float* s_scale = reinterpret_cast<float*>(tmem_slot + 4);
    float* s_shift = s_scale + 256;
    float* s_sum = s_shift + 256;
    float* s_sq = s_sum + 256;
    float* s_stage = s_sq + 256;                              // EPI_STAGE_BYTES per epilogue warp

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && elect_one()) {
        tma_prefetch_desc(&tmap_x);
        tma_prefetch_desc(&tmap_w);
    }
    if (warp == 1 && elect_one()) {
        for (int i = 0; i < a.a_stages; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
        for (int i = 0; i < a.b_stages; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], EPI_THREADS); }
        fence_mbar_init();
    }
    if (warp == 2) tmem_alloc(tmem_slot, a.tmem_cols);
    for (int i = threadIdx.x; i < 256; i += NUM_THREADS) {
        s_scale[i] = (a.scale && i < a.N) ? a.scale[i] : 1.0f;
        s_shift[i] = (a.shift && i < a.N) ? a.shift[i] : 0.0f;
        s_sum[i] = 0.0f;
        s_sq[i] = 0.0f;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int tiles_per_frame = a.tiles_w * a.tiles_h;

    if (warp == 0) {
        if (elect_one()) {
            // ------------------------------ TMA producer ------------------------------
            int as = 0, bs = 0;
            uint32_t aphase = 0, bphase = 0;
            const int taps_hw = a.kh * a.kw;
            if (a.b_resident) {
                // small weight sets (1x1 lateral convs, the 2x2 ConvTranspose taps, the box predictor): every (tap, chunk)
                // tile is loaded ONCE per CTA and reused by all its pixel tiles -- the per-tile L2 -> SM traffic drops to A
                // (with several N chunks the grid is a multiple of nchunks, so this CTA only ever sees chunk blockIdx.x % nchunks)
                const int nb = a.kt * taps_hw * a.cchunks;
                const int n_res = (int)(blockIdx.x % a.nchunks) * a.N;
                mbar_arrive_expect_tx(&b_full[0], (uint32_t)(nb * b_bytes));
                for (int tp = 0; tp < a.kt * taps_hw; ++tp)
                    for (int cc = 0; cc < a.cchunks; ++cc)
                        tma_load_3d(smem_b + (tp * a.cchunks + cc) * b_bytes, &tmap_w, &b_full[0], cc * BK, n_res, tp);
            }
            for (int item = blockIdx.x; item < a.ntiles * a.nchunks; item += gridDim.x) {
                // consecutive items = the N chunks of one pixel tile: concurrent CTAs share the activation tile in L2
                const int tile = item / a.nchunks;
                const int n0 = (item - tile * a.nchunks) * a.N;
                const int frame = tile / tiles_per_frame;
                const int rem = tile - frame * tiles_per_frame;
                const int th_i = rem / a.tiles_w;
                const int tw_i = rem - th_i * a.tiles_w;
                const int b = frame / a.To;
                const int t = frame - b * a.To;
                const int h0 = th_i * a.TH, w0 = tw_i * a.TW;
                if (!a.halo) {
                    // per-tap mode: A tile and B tile of one K step share a stage and a barrier
                    for (int ta = 0; ta < a.kt; ++ta)
                        for (int ti = 0; ti < a.kh; ++ti)
                            for (int tj = 0; tj < a.kw; ++tj)
                                for (int cc = 0; cc < a.cchunks; ++cc) {
                                    mbar_wait(&a_empty[as], aphase ^ 1);
                                    mbar_arrive_expect_tx(&a_full[as], a.a_tx_bytes + (a.b_resident ? 0 : b_bytes));
                                    tma_load_5d(smem_a + as * a.a_stage_bytes, &tmap_x, &a_full[as], cc * BK,
                                                w0 + tj - a.pad_w, h0 + ti - a.pad_h, t + ta - a.pad_t, b);
                                    if (!a.b_resident)
                                        tma_load_3d(smem_b + as * b_bytes, &tmap_w, &a_full[as], cc * BK, n0,
                                                    (ta * a.kh + ti) * a.kw + tj);
                                    if (++as == a.a_stages) { as = 0; aphase ^= 1; }
                                }
                } else {
                    // halo mode: one A box (tile + halo) per (temporal tap, chunk); weights in groups of b_group taps
                    for (int ta = 0; ta < a.kt; ++ta)
                        for (int cc = 0; cc < a.cchunks; ++cc) {
                            mbar_wait(&a_empty[as], aphase ^ 1);
                            mbar_arrive_expect_tx(&a_full[as], a.a_tx_bytes);
                            tma_load_5d(smem_a + as * a.a_stage_bytes, &tmap_x, &a_full[as], cc * BK, w0 - a.pad_w,
                                        h0 - a.pad_h, t + ta - a.pad_t, b);
                            if (++as == a.a_stages) { as = 0; aphase ^= 1; }
                            for (int g = 0; g < taps_hw; g += a.b_group) {
                                mbar_wait(&b_empty[bs], bphase ^ 1);
                                mbar_arrive_expect_tx(&b_full[bs], b_stage_bytes);
                                tma_load_3d(smem_b + bs * b_stage_bytes, &tmap_w, &b_full[bs], cc * BK, n0, ta * taps_hw + g);
                                if (++bs == a.b_stages) { bs = 0; bphase ^= 1; }
                            }
                        }
                }
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            // ------------------------------ MMA issuer ------------------------------
            int as = 0, bs = 0;
            uint32_t aphase = 0, bphase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            const int taps_hw = a.kh * a.kw;
            const int outer = a.halo ? a.kt * a.cchunks : a.kt * taps_hw * a.cchunks;
            if (a.b_resident) mbar_wait(&b_full[0], 0);
            for (int item = blockIdx.x; item < a.ntiles * a.nchunks; item += gridDim.x) {
                mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * a.N;
                uint32_t accum = 0;
                for (int o = 0; o < outer; ++o) {
                    mbar_wait(&a_full[as], aphase);
                    const uint32_t a_addr = smem_u32(smem_a + as * a.a_stage_bytes);
                    if (!a.halo) {
                        tc_fence_after();
                        const uint32_t b_addr = smem_u32(smem_b + (a.b_resident ? o : as) * b_bytes);
#pragma unroll
                        for (int k = 0; k < BK / 16; ++k) {
                            const uint64_t adesc = umma_smem_desc(a_addr + k * 32, 16, 8 * ROW, LAYOUT);
                            const uint64_t bdesc = umma_smem_desc(b_addr + k * 32, 16, 8 * ROW, LAYOUT);
                            umma_bf16(d_tmem, adesc, bdesc, a.idesc, accum);
                            accum = 1;
                        }
                    } else {
                        for (int g = 0; g < taps_hw; g += a.b_group) {
                            mbar_wait(&b_full[bs], bphase);
                            tc_fence_after();
                            const uint32_t b_base = smem_u32(smem_b + bs * b_stage_bytes);
                            for (int tl = 0; tl < a.b_group; ++tl) {
                                const int sp = g + tl;                      // spatial tap index i*kw + j
                                const int ti = sp / a.kw, tj = sp - ti * a.kw;
                                // the tap's A operand = the same halo buffer, start shifted by (i lines + j pixels);
                                // consecutive 8-row groups are one 16-pixel line (HALO_LP rows) apart
                                const uint32_t a_tap = a_addr + (ti * HALO_LP + tj) * ROW;
                                const uint32_t b_addr = b_base + tl * b_bytes;
#pragma unroll
                                for (int k = 0; k < BK / 16; ++k) {
                                    uint64_t adesc = umma_smem_desc(a_tap + k * 32, 16, HALO_LP * ROW, LAYOUT);
                                    if (a.use_bo) adesc |= static_cast<uint64_t>((a_tap >> 7) & 7u) << 49;
                                    const uint64_t bdesc = umma_smem_desc(b_addr + k * 32, 16, 8 * ROW, LAYOUT);
                                    umma_bf16(d_tmem, adesc, bdesc, a.idesc, accum);
                                    accum = 1;
                                }
                            }
                            umma_commit(&b_empty[bs]);
                            if (++bs == a.b_stages) { bs = 0; bphase ^= 1; }
                        }
                    }
                    umma_commit(&a_empty[as]);
                    if (++as == a.a_stages) { as = 0; aphase ^= 1; }
                }
                umma_commit(&tmem_full[acc]);            // accumulator complete -> epilogue
                acc ^= 1;
                if (acc == 0) acc_phase ^= 1;
            }
        }
    } else if (warp >= EPI_WARP0) {
        // ------------------------------ epilogue ------------------------------
        // A single epilogue warp per scheduler runs its dependent chains (TMEM load -> transpose -> convert -> store) at
        // ~0.15 IPC; measured on the mask predictor's ConvTranspose GEMM (K = 256, N = 1024, bf16 output): main loop alone
        // 105 us, with the 4-warp epilogue 383 us.  Two warps per lane quarter split the accumulator's column blocks.
        const int ew = warp - EPI_WARP0;
        const int q = ew & 3;                            // TMEM lane quarter == warp id % 4
        const int c_split = ((a.N / 32 + 1) / 2) * 32;
        const int c_begin = (ew >> 2) ? c_split : 0, c_end = (ew >> 2) ? a.N : c_split;
        const int r = q * 32 + lane;                     // accumulator row = pixel within the tile
        const int hl = r / a.TW;
        const int wl = r - hl * a.TW;
        const bool do_stats = (a.sum != nullptr) && (a.relu_mask == nullptr);
        const bool affine = (a.scale != nullptr) || (a.shift != nullptr);
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int item = blockIdx.x; item < a.ntiles * a.nchunks; item += gridDim.x) {
            const int tile = item / a.nchunks;
            const int nbase = (item - tile * a.nchunks) * a.N;     // first output channel of this N chunk
            const int frame = tile / tiles_per_frame;
            const int rem = tile - frame * tiles_per_frame;
            const int th_i = rem / a.tiles_w;
            const int tw_i = rem - th_i * a.tiles_w;
            const int h = th_i * a.TH + hl, w = tw_i * a.TW + wl;
            const bool valid = (hl < a.TH) && (h < a.H) && (w < a.W);
            const long long pix = ((long long)frame * a.OH + (h * a.oy_mul + a.oy_off)) * a.OW + (w * a.ox_mul + a.ox_off);
            mbar_wait(&tmem_full[acc], acc_phase);
            tc_fence_after();
            const uint32_t t_addr = tmem_base + (uint32_t(q * 32) << 16) + acc * a.N;
            if (a.epi_stage) {
                const EpiRows rows = epi_rows(valid ? (int)pix : -1, lane);
                EpiOut eo;
                eo.y = a.y; eo.y_cstride = a.y_cstride; eo.y_bf16 = a.y_bf16; eo.relu = a.relu; eo.accumulate = a.accumulate;
                eo.relu_mask = a.relu_mask; eo.mask_cstride = a.mask_cstride;
                eo.addend = a.addend; eo.addend_cstride = a.addend_cstride; eo.addend_bf16 = a.addend_bf16;
                // wide fc layers (several N chunks): bias / scale straight from global memory, indexed by absolute channel
                const float* scp = !affine ? nullptr : (a.nchunks > 1 ? a.scale : s_scale);
                const float* shp = !affine ? nullptr : (a.nchunks > 1 ? a.shift : s_shift);
                float* sum_dst = do_stats ? s_sum : (a.relu_mask != nullptr ? a.sum : nullptr);
                float* sq_dst = do_stats ? s_sq : nullptr;
                EpiRows rows_d = rows;
                if (a.epi_stage & 2) {                   // DEBUG: nothing is stored
#pragma unroll
                    for (int i = 0; i < 8; ++i) rows_d.pix[i] = -1;
                }
                for (int c0 = c_begin; c0 < ((a.epi_stage & 4) ? 0 : c_end); c0 += 32) {      // DEBUG bit 2: accumulators released unread
                    uint32_t v[32];
                    tmem_ld_32x32(t_addr + c0, v);
                    tmem_ld_wait();
                    epi_block(s_stage + ew * (EPI_STAGE_BYTES / 4), v, rows_d, lane, nbase + c0, scp, shp, eo, sum_dst, sq_dst);
                }
            } else
            for (int c0 = c_begin; c0 < c_end; c0 += 32) {
                uint32_t v[32];
                tmem_ld_32x32(t_addr + c0, v);
                tmem_ld_wait();
                if (do_stats) {
                    float f[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) f[j] = valid ? __uint_as_float(v[j]) : 0.0f;
                    float s = warp_transpose_reduce32(f, lane);
                    atomicAdd(&s_sum[c0 + lane], s);
#pragma unroll
                    for (int j = 0; j < 32; ++j) { float x = valid ? __uint_as_float(v[j]) : 0.0f; f[j] = x * x; }
                    s = warp_transpose_reduce32(f, lane);
                    atomicAdd(&s_sq[c0 + lane], s);
                }
                if (a.relu_mask != nullptr) {
                    // fused ReLU backward of the layer below (see conv_pair_umma.cu): mask, then column sums of the masked gradient
                    if (valid) {
                        const uint4* mp = reinterpret_cast<const uint4*>(a.relu_mask + pix * a.mask_cstride + nbase + c0);
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const uint4 m = __ldg(mp + j);
                            const uint32_t mw[4] = {m.x, m.y, m.z, m.w};
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const uint32_t lo = mw[e] & 0xffffu, hi = mw[e] >> 16;
                                if (!(lo != 0 && lo < 0x8000u)) v[8 * j + 2 * e] = 0u;
                                if (!(hi != 0 && hi < 0x8000u)) v[8 * j + 2 * e + 1] = 0u;
                            }
                        }
                    }
                    if (a.sum != nullptr) {
                        float f[32];
#pragma unroll
                        for (int j = 0; j < 32; ++j) f[j] = valid ? __uint_as_float(v[j]) : 0.0f;
                        const float s = warp_transpose_reduce32(f, lane);
                        atomicAdd(&a.sum[nbase + c0 + lane], s);      // straight to global: N chunks do not share s_sum
                    }
                }
                if (valid) {
                    float o[32];
                    if (affine && a.nchunks > 1) {      // wide fc layers: bias (and scale) straight from global memory
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const float sc = a.scale ? __ldg(a.scale + nbase + c0 + j) : 1.0f;
                            const float sh = a.shift ? __ldg(a.shift + nbase + c0 + j) : 0.0f;
                            o[j] = fmaf(__uint_as_float(v[j]), sc, sh);
                        }
                    } else if (affine) {    // per-channel scale / shift from shared memory, 16 bytes per load
                        const float4* sc4 = reinterpret_cast<const float4*>(s_scale + c0);
                        const float4* sh4 = reinterpret_cast<const float4*>(s_shift + c0);
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const float4 sc = sc4[j], sh = sh4[j];
                            o[4 * j + 0] = fmaf(__uint_as_float(v[4 * j + 0]), sc.x, sh.x);
                            o[4 * j + 1] = fmaf(__uint_as_float(v[4 * j + 1]), sc.y, sh.y);
                            o[4 * j + 2] = fmaf(__uint_as_float(v[4 * j + 2]), sc.z, sh.z);
                            o[4 * j + 3] = fmaf(__uint_as_float(v[4 * j + 3]), sc.w, sh.w);
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j) o[j] = __uint_as_float(v[j]);
                    }
                    if (a.relu) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) o[j] = fmaxf(o[j], 0.0f);
                    }
                    if (a.y_bf16) {
                        uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(a.y) + pix * a.y_cstride + nbase + c0);
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            uint4 u;
                            u.x = pack_bf16x2(o[8 * j + 0], o[8 * j + 1]);
                            u.y = pack_bf16x2(o[8 * j + 2], o[8 * j + 3]);
                            u.z = pack_bf16x2(o[8 * j + 4], o[8 * j + 5]);
                            u.w = pack_bf16x2(o[8 * j + 6], o[8 * j + 7]);
                            dst[j] = u;
                        }
                    } else {
                        float4* dst = reinterpret_cast<float4*>(reinterpret_cast<float*>(a.y) + pix * a.y_cstride + nbase + c0);
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            float4 u = make_float4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
                            if (a.accumulate) {
                                float4 old = dst[j];
                                u.x += old.x; u.y += old.y; u.z += old.z; u.w += old.w;
                            }
                            dst[j] = u;
                        }
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(&tmem_empty[acc]);
            acc ^= 1;
            if (acc == 0) acc_phase ^= 1;
        }
        if (do_stats) {
            asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory");
            for (int i = threadIdx.x - EPI_WARP0 * 32; i < a.N; i += EPI_THREADS) {
                atomicAdd(&a.sum[i], s_sum[i]);
                atomicAdd(&a.sumsq[i], s_sq[i]);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, a.tmem_cols);
    }
}

double choose_tile(int H, int W, int* TW, int* TH) {
    double best = -1.0;
    int bw = 1, bh = 1;
    for (int tw = 1; tw <= W && tw <= BM; ++tw) {
        int th = BM / tw;
        if (th > H) th = H;
        if (th < 1) continue;
        long long tiles = (long long)((W + tw - 1) / tw) * ((H + th - 1) / th);
        double eff = (double)H * W / (double)(tiles * BM);
        if (eff > best + 1e-9 || (fabs(eff - best) <= 1e-9 && tw > bw)) {
            best = eff; bw = tw; bh = th;
        }
    }
    *TW = bw; *TH = bh;
    return best;
}

int env_int(const char* name, int dflt) {
    const char* v = getenv(name);
    return v ? atoi(v) : dflt;
}

}  // namespace

// conv_tstack_umma.cu: temporally-stacked kernel for the Cout = 32 layers
int sfvos_conv_tstack_applicable(const sfvos_conv_params* p);
int sfvos_conv_tstack_launch(const sfvos_conv_params* p, cudaStream_t stream);
// conv_pair_umma.cu: CTA-pair (cta_group::2) kernel for the wide 3x3 layers
int sfvos_conv_pair_applicable(const sfvos_conv_params* p);
int sfvos_conv_pair_launch(const sfvos_conv_params* p, cudaStream_t stream);

extern "C" int sfvos_conv_umma(const sfvos_conv_params* p, sfvos_stream stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    SF_CHECK(p != nullptr, "conv_umma: null params");
    SF_CHECK(p->N >= 32 && p->N % 32 == 0 && (p->N <= 256 || p->N % 256 == 0),
             "conv_umma: N=%lld must be a multiple of 32 in [32,256] or a multiple of 256", (long long)p->N);
    SF_CHECK(p->N <= 256 || p->sum == nullptr || p->relu_mask != nullptr, "conv_umma: fused statistics need N <= 256");
    SF_CHECK(p->Cp % 32 == 0 && p->Cp >= p->C, "conv_umma: Cp=%lld must be a multiple of 32 and >= C=%lld", (long long)p->Cp, (long long)p->C);
    const int BK = (p->Cp % 64 == 0) ? 64 : 32;
    SF_CHECK(p->C % 8 == 0 && p->x_cstride % 8 == 0, "conv_umma: C and x_cstride must be multiples of 8");
    SF_CHECK(p->y_cstride % 8 == 0 && (reinterpret_cast<uintptr_t>(p->y) & 15) == 0, "conv_umma: y must be 16-byte aligned with cstride %% 8 == 0");
    SF_CHECK(!(p->accumulate && p->y_dtype != SFVOS_F32), "conv_umma: accumulate needs an f32 output");
    SF_CHECK(p->addend == nullptr || (!p->accumulate && p->addend_cstride % 8 == 0 && (reinterpret_cast<uintptr_t>(p->addend) & 15) == 0 &&
                                      (p->addend_dtype == SFVOS_F32 || p->addend_dtype == SFVOS_BF16)),
             "conv_umma: addend must be a 16-byte aligned f32 / bf16 tensor with cstride %% 8 == 0, and excludes accumulate");
    SF_CHECK((p->sum == nullptr) == (p->sumsq == nullptr) || p->relu_mask != nullptr, "conv_umma: sum and sumsq must be given together");
    SF_CHECK(p->B > 0 && p->To > 0 && p->H > 0 && p->W > 0 && p->T > 0, "conv_umma: empty tensor");
    if (p->relu_mask != nullptr) {
        SF_CHECK(p->scale == nullptr && p->shift == nullptr && !p->relu && !p->accumulate && p->sumsq == nullptr,
                 "conv_umma: relu_mask (fused ReLU backward) excludes scale/shift/relu/accumulate/sumsq");
        SF_CHECK(p->relu_mask_cstride % 8 == 0 && (reinterpret_cast<uintptr_t>(p->relu_mask) & 15) == 0,
                 "conv_umma: relu_mask must be 16-byte aligned with cstride %% 8 == 0");
        SF_CHECK((!p->OH || p->OH == p->H) && (!p->OW || p->OW == p->W) && (!p->oy_mul || p->oy_mul == 1) && (!p->ox_mul || p->ox_mul == 1),
                 "conv_umma: relu_mask needs the identity output mapping");
    }
    SF_CHECK(p->B * p->To * p->H * p->W < (1LL << 31), "conv_umma: too many output pixels");
    SF_CHECK(p->B * p->To * (p->OH ? p->OH : p->H) * (p->OW ? p->OW : p->W) < (1LL << 31), "conv_umma: too many output pixels");
    int rc = sfvos_device_check();
    if (rc) return rc;
    if (sfvos_conv_tstack_applicable(p)) return sfvos_conv_tstack_launch(p, stream);
    if (sfvos_conv_pair_applicable(p)) return sfvos_conv_pair_launch(p, stream);

    ConvArgs a;
    a.B = (int)p->B; a.To = (int)p->To; a.H = (int)p->H; a.W = (int)p->W;
    const double eff_tap = choose_tile(a.H, a.W, &a.TW, &a.TH);
    // halo mode: 8 x 16 tiles; worth it when its tiling wastes little more than the per-tap tiling
    a.halo = 0;
    if (p->kh == 3 && p->kw == 3 && env_int("SFVOS_HALO", 1)) {
        const long long tiles = (long long)((a.W + 7) / 8) * ((a.H + 15) / 16);
        const double eff_halo = (double)a.H * a.W / (double)(tiles * BM);
        if (eff_halo >= eff_tap - 0.15) { a.halo = 1; a.TW = 8; a.TH = 16; }
    }
    a.b_resident = 0;
    a.use_bo = env_int("SFVOS_HALO_BO", 0);   // measured on B200: the swizzle XOR uses absolute smem address bits, so a shifted start needs NO base offset
    a.tiles_w = (a.W + a.TW - 1) / a.TW;
    a.tiles_h = (a.H + a.TH - 1) / a.TH;
    a.ntiles = a.B * a.To * a.tiles_w * a.tiles_h;
    a.nchunks = p->N > 256 ? (int)(p->N / 256) : 1;
    a.N = p->N > 256 ? 256 : (int)p->N;
    a.kt = (int)p->kt; a.kh = (int)p->kh; a.kw = (int)p->kw;
    a.pad_t = (int)p->pad_t; a.pad_h = (int)p->pad_h; a.pad_w = (int)p->pad_w;
    a.cchunks = (int)(p->Cp / BK);
    const int b_bytes = a.N * BK * 2;
    const int small_bytes = 8192 /*barriers, scale/shift, stats*/ + (EPI_THREADS / 32) * EPI_STAGE_BYTES /*epilogue transpose tiles*/;
    const int smem_budget = 227 * 1024 - 1024 /*align*/ - small_bytes;
    a.epi_stage = env_int("SFVOS_EPI_STAGE", 1);
    uint32_t abox_w, abox_h;
    if (a.halo) {
        abox_w = HALO_LP; abox_h = (uint32_t)(a.TH + 2);
        a.a_stage_bytes = HALO_LP * (a.TH + 2) * BK * 2;                    // 36 KB (BK=64) / 18 KB (BK=32)
        // narrow N: the MMA-issuing thread, not the tensor pipe, is the limiter -> all 9 taps' weights per barrier
        a.b_group = (9 * b_bytes <= 72 * 1024 && a.N <= 64) ? 9 : 1;
        if (a.b_group == 9 && smem_budget - 2 * 9 * b_bytes < 2 * a.a_stage_bytes) a.b_group = 3;   // N = 64 at BK = 64: 3 taps per barrier
        const int b_stage = a.b_group * b_bytes;
        a.b_stages = a.b_group == 9 ? 2 : 4;
        while (a.b_stages > 2 && smem_budget - a.b_stages * b_stage < 2 * a.a_stage_bytes) --a.b_stages;     // wide N: fewer weight stages
        a.a_stages = (smem_budget - a.b_stages * b_stage) / a.a_stage_bytes;
        if (a.a_stages > 4) a.a_stages = 4;
        if (a.b_group == 1 && a.a_stages > 2 && b_bytes <= 8192) a.b_stages = 8;
        SF_CHECK(a.a_stages >= 2, "conv_umma: not enough shared memory for the halo pipeline");
    } else {
        abox_w = (uint32_t)a.TW; abox_h = (uint32_t)a.TH;
        a.a_stage_bytes = BM * BK * 2;
        a.b_group = 1;
        const long long all_b = (long long)p->kt * p->kh * p->kw * a.cchunks * b_bytes;
        if (all_b <= 144 * 1024 && all_b + 3 * a.a_stage_bytes <= smem_budget && env_int("SFVOS_B_RESIDENT", 1) &&
            (a.nchunks == 1 || sfvos_num_sms() >= a.nchunks)) {
            a.b_resident = 1;
            a.b_stages = (int)(all_b / b_bytes);
            int st = (int)((smem_budget - all_b) / a.a_stage_bytes);
            a.a_stages = st > 8 ? 8 : st;
            SF_CHECK(a.a_stages >= 3, "conv_umma: not enough shared memory next to the resident weights");
        } else {
            int st = smem_budget / (a.a_stage_bytes + b_bytes);
            if (st > 8) st = 8;
            SF_CHECK(st >= 2, "conv_umma: not enough shared memory for 2 stages");
            a.a_stages = a.b_stages = st;
        }
    }
    a.a_tx_bytes = abox_w * abox_h * BK * 2;
    a.idesc = umma_idesc_bf16(BM, a.N, 0, 0);
    uint32_t cols = 32;
    while (cols < (uint32_t)(2 * a.N)) cols <<= 1;
    a.tmem_cols = cols;
    a.y = p->y; a.y_bf16 = (p->y_dtype == SFVOS_BF16); a.relu = p->relu; a.accumulate = p->accumulate;
    a.y_cstride = p->y_cstride;
    a.scale = p->scale; a.shift = p->shift; a.sum = p->sum; a.sumsq = p->sumsq;
    a.relu_mask = reinterpret_cast<const __nv_bfloat16*>(p->relu_mask); a.mask_cstride = p->relu_mask_cstride;
    a.addend = p->addend; a.addend_cstride = p->addend_cstride; a.addend_bf16 = (p->addend_dtype == SFVOS_BF16);
    SF_CHECK(p->addend == nullptr || (a.epi_stage & 1), "conv_umma: addend needs the transposing epilogue (SFVOS_EPI_STAGE)");
    a.OH = (int)(p->OH ? p->OH : p->H); a.OW = (int)(p->OW ? p->OW : p->W);
    a.oy_mul = (int)(p->oy_mul ? p->oy_mul : 1); a.ox_mul = (int)(p->ox_mul ? p->ox_mul : 1);
    a.oy_off = (int)p->oy_off; a.ox_off = (int)p->ox_off;

    CUtensorMap tx, tw;
    {
        uint64_t dims[5] = {(uint64_t)p->C, (uint64_t)p->W, (uint64_t)p->H, (uint64_t)p->T, (uint64_t)p->B};
        const uint64_t cs = (uint64_t)p->x_cstride;
        const uint64_t hs = p->x_hstride ? (uint64_t)p->x_hstride : cs * p->W;
        const uint64_t ts = p->x_tstride ? (uint64_t)p->x_tstride : hs * p->H;
        const uint64_t bs = p->x_bstride ? (uint64_t)p->x_bstride : ts * p->T;
        uint64_t str[4] = {cs * 2, hs * 2, ts * 2, bs * 2};
        uint32_t box[5] = {(uint32_t)BK, abox_w, abox_h, 1, 1};
        rc = sfvos_make_tmap(&tx, p->x, 5, dims, str, box, BK * 2);
        if (rc) return rc;
    }
    {
        // packed weights [N][taps][Cp] viewed as {Cp, N, taps}: a box {BK, N, G} lands as G consecutive [N x BK] tiles
        const uint64_t taps = (uint64_t)(p->kt * p->kh * p->kw);
        uint64_t dims[3] = {(uint64_t)p->Cp, (uint64_t)p->N, taps};
        uint64_t str[2] = {taps * p->Cp * 2, (uint64_t)p->Cp * 2};
        uint32_t box[3] = {(uint32_t)BK, (uint32_t)a.N, (uint32_t)a.b_group};
        rc = sfvos_make_tmap(&tw, p->w, 3, dims, str, box, BK * 2);
        if (rc) return rc;
    }
    const int smem_bytes = a.a_stages * a.a_stage_bytes + a.b_stages * a.b_group * b_bytes + 1024 + small_bytes;
    int grid = sfvos_num_sms();
    if (grid > a.ntiles * a.nchunks) grid = a.ntiles * a.nchunks;
    if (a.b_resident && a.nchunks > 1) grid -= grid % a.nchunks;     // every CTA keeps ONE N chunk (its resident weights)
    if (BK == 64) {
        SF_CUDA(cudaFuncSetAttribute(conv_umma_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        conv_umma_kernel<64><<<grid, NUM_THREADS, smem_bytes, stream>>>(tx, tw, a);
    } else {
        SF_CUDA(cudaFuncSetAttribute(conv_umma_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        conv_umma_kernel<32><<<grid, NUM_THREADS, smem_bytes, stream>>>(tx, tw, a);
    }
    sfvos_set_kernel("conv_umma");
    SF_LAUNCH_CHECK();
    return SFVOS_OK;
}
