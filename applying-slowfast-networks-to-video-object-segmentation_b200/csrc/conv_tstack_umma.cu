// Temporally-stacked implicit-GEMM convolution for the NARROW (Cout = 32) fast-pathway layers, tcgen05 / sm_100a:
// fprop of fast_conv1/2/3 (code/helpers/model.py:47-48,53-54,59-60, invoked :124,136,147) and dgrad of fast_conv2/3.
//
// conv_umma.cu computes one output frame per tile: D[128 px, 32] per MMA.  A 128x32x16 MMA reads 4 KB of A for 16
// cycles of math, so the tensor pipe idles on operand reads, and every input frame is fetched k_t times.  Here one
// work item owns a spatial tile for a whole GROUP of output frames and walks the INPUT frames once:
//     y[t][px][n] = sum_{ta} sum_{s,c} x[t + ta - pad_t][px + s][c] * W[ta][s][c][n]
// For input frame tau the products with ALL temporal taps ta are one MMA: the accumulators of the output frames
// t = tau + pad_t - ta sit side by side in TMEM (frame t at column block (t1-1-t)*32, i.e. descending t), and the
// weights of the taps are stacked along N in ascending ta, so
//     D[128 px, 32*G cols starting at block(t_hi)] += A[tau tile, tap s, 16 ch] * [W[ta_lo] | W[ta_lo+1] | ...]
// One activation box per (input frame, 64-ch chunk) -- tile + halo, as in conv_umma's halo mode -- feeds 9 spatial
// taps x G temporal taps: bytes ingested and A-operand reads per FLOP drop by G (3 for fast_conv1 at fp = 8).
//   warp 0 = activation TMA producer, warp 3 = weight TMA producer (9 per-spatial-tap pieces, reloaded per
//   (tap group, chunk) as soon as the MMAs of the previous chunk's last frame release them), warp 1 = MMA issuer,
//   warp 2 = TMEM allocator, warps 4..11 = epilogue (per output frame: BN statistics, affine/ReLU, store; two warps per
//   TMEM lane quarter take alternate frames).
#include <stdlib.h>

#include "common.cuh"

namespace {

constexpr int BM = 128;
constexpr int NC = 32;                      // output channels per frame
constexpr int NUM_THREADS = 384;             // warps 0-3: TMA (activations) / MMA / TMEM alloc / TMA (weights); warps 4-11: epilogue
constexpr int EPI_WARP0 = 4;
constexpr int EPI_THREADS = 256;             // TWO warps per TMEM lane quarter, taking alternate output frames
constexpr int TW = 8, TH = 16;              // spatial tile (8 wide x 16 tall = 128 accumulator rows)
constexpr int SMALL_BYTES = 2048;           // barriers, scale / shift, statistics
constexpr int STAGE_RESERVE = (EPI_THREADS / 32) * EPI_STAGE_BYTES;   // one 32 x 32 f32 transpose tile per epilogue warp (common.cuh)

struct TsArgs {
    int B, To, Ti, H, W;
    int tiles_w, tiles_per_frame, nitems;
    int kt, pad_t, G, ngroups, Fg, nfg, cchunks;
    int LP, a_stages, a_stage_bytes, piece_bytes;
    uint32_t a_tx_bytes, tmem_cols;
    int nbuf, b_resident, stage_mode;            // stage_mode: f32 epilogue through the shared-memory transpose
    int dbg;                                     // SFVOS_TSTACK_DBG (measurement only): 1 = the epilogue releases accumulators unread
    void* y;
    int y_bf16, relu, accumulate;
    long long y_cstride;
    const float* scale;
    const float* shift;
    float* sum;
    float* sumsq;
    const void* addend;                          // y = act(...) + addend (see sfvos_conv_params)
    long long addend_cstride;
    int addend_bf16;
};

struct Item {
    int b, t0, t1, h0, w0;
};

__device__ __forceinline__ Item decode_item(const TsArgs& a, int item) {
    Item it;
    const int tile = item % a.tiles_per_frame;
    const int r = item / a.tiles_per_frame;
    const int fg = r % a.nfg;
    it.b = r / a.nfg;
    it.t0 = fg * a.Fg;
    it.t1 = min(it.t0 + a.Fg, a.To);
    const int th_i = tile / a.tiles_w;
    it.h0 = th_i * TH;
    it.w0 = (tile - th_i * a.tiles_w) * TW;
    return it;
}

// input frames that meet at least one (output frame in [t0,t1), tap in [ta0,ta0+gn)) pair
__device__ __forceinline__ void tau_range(const TsArgs& a, const Item& it, int ta0, int gn, int* lo, int* hi) {
    *lo = max(0, it.t0 + ta0 - a.pad_t);
    *hi = min(a.Ti - 1, it.t1 - 1 + ta0 + gn - 1 - a.pad_t);
}

template <int BK, int NSP>
__global__ void __launch_bounds__(NUM_THREADS, 1)
conv_tstack_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
                   const TsArgs a) {
    constexpr uint32_t LAYOUT = BK == 64 ? 2u : 4u;          // SWIZZLE_128B : SWIZZLE_64B
    constexpr uint32_t ROW = BK * 2;                          // bytes per smem row
    constexpr int HALO = NSP == 9 ? 1 : 0;
    constexpr int KW = NSP == 9 ? 3 : 1;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + a.a_stages * a.a_stage_bytes;
    uint64_t* a_full = reinterpret_cast<uint64_t*>(smem_b + NSP * a.piece_bytes);
    uint64_t* a_empty = a_full + a.a_stages;
    uint64_t* b_full = a_empty + a.a_stages;
    uint64_t* b_empty = b_full + NSP;
    uint64_t* tmem_full = b_empty + NSP;
    uint64_t* tmem_empty = tmem_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
    float* s_scale = reinterpret_cast<float*>(tmem_slot + 4);
    float* s_shift = s_scale + NC;
    float* s_sum = s_shift + NC;
    float* s_sq = s_sum + NC;
    float* s_stage = reinterpret_cast<float*>(smem_b + NSP * a.piece_bytes + SMALL_BYTES);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && elect_one()) {
        tma_prefetch_desc(&tmap_x);
        tma_prefetch_desc(&tmap_w);
    }
    if (warp == 1 && elect_one()) {
        for (int i = 0; i < a.a_stages; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
        for (int i = 0; i < NSP; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], EPI_THREADS); }
        fence_mbar_init();
    }
    if (warp == 2) tmem_alloc(tmem_slot, a.tmem_cols);
    if (threadIdx.x < NC) {
        const int i = threadIdx.x;
        s_scale[i] = a.scale ? a.scale[i] : 1.0f;
        s_shift[i] = a.shift ? a.shift[i] : 0.0f;
        s_sum[i] = 0.0f;
        s_sq[i] = 0.0f;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (elect_one()) {
            // ------------------------------ activation producer ------------------------------
            int as = 0;
            uint32_t aphase = 0;
            for (int item = blockIdx.x; item < a.nitems; item += gridDim.x) {
                const Item it = decode_item(a, item);
                for (int tg = 0; tg < a.ngroups; ++tg) {
                    const int ta0 = tg * a.G, gn = min(a.G, a.kt - ta0);
                    int lo, hi;
                    tau_range(a, it, ta0, gn, &lo, &hi);
                    if (lo > hi) continue;
                    for (int cc = 0; cc < a.cchunks; ++cc)
                        for (int tau = lo; tau <= hi; ++tau) {
                            mbar_wait(&a_empty[as], aphase ^ 1);
                            mbar_arrive_expect_tx(&a_full[as], a.a_tx_bytes);
                            tma_load_5d(smem_a + as * a.a_stage_bytes, &tmap_x, &a_full[as], cc * BK, it.w0 - HALO,
                                        it.h0 - HALO, tau, it.b);
                            if (++as == a.a_stages) { as = 0; aphase ^= 1; }
                        }
                }
            }
        }
    } else if (warp == 3) {
        if (elect_one()) {
            // ------------------------------ weight producer ------------------------------
            // resident mode (one tap group, one channel chunk, e.g. the Cin = 32 layers): the 9 pieces ARE the whole
            // weight set, loaded once per CTA; otherwise they are re-streamed per (item, tap group, chunk)
            uint32_t bphase = 0;
            for (int item = blockIdx.x; item < a.nitems; item += gridDim.x) {
                const Item it = decode_item(a, item);
                for (int tg = 0; tg < a.ngroups; ++tg) {
                    const int ta0 = tg * a.G, gn = min(a.G, a.kt - ta0);
                    int lo, hi;
                    tau_range(a, it, ta0, gn, &lo, &hi);
                    if (lo > hi) continue;
                    for (int cc = 0; cc < a.cchunks; ++cc) {
                        for (int s = 0; s < NSP; ++s) {
                            mbar_wait(&b_empty[s], bphase ^ 1);
                            mbar_arrive_expect_tx(&b_full[s], a.piece_bytes);
                            tma_load_4d(smem_b + s * a.piece_bytes, &tmap_w, &b_full[s], cc * BK, 0, ta0, s);
                        }
                        bphase ^= 1;
                    }
                }
                if (a.b_resident) break;
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            // ------------------------------ MMA issuer ------------------------------
            int as = 0;
            uint32_t aphase = 0, bphase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            const uint32_t idesc0 = umma_idesc_bf16(BM, 0, 0, 0);
            const uint64_t adesc0 = umma_smem_desc(0, 16, a.LP * ROW, LAYOUT);
            const uint64_t bdesc0 = umma_smem_desc(0, 16, 8 * ROW, LAYOUT);
            const uint32_t b_base = smem_u32(smem_b);
            bool b_loaded = false;                      // resident mode: the weights have been waited for once
            for (int item = blockIdx.x; item < a.nitems; item += gridDim.x) {
                const Item it = decode_item(a, item);
                mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_base = tmem_base + acc * a.Fg * NC;
                uint32_t touched = 0;                       // bit f: accumulator of frame slot f holds a partial sum
                for (int tg = 0; tg < a.ngroups; ++tg) {
                    const int ta0 = tg * a.G, gn = min(a.G, a.kt - ta0);
                    int lo, hi;
                    tau_range(a, it, ta0, gn, &lo, &hi);
                    if (lo > hi) continue;
                    for (int cc = 0; cc < a.cchunks; ++cc) {
                        for (int tau = lo; tau <= hi; ++tau) {
                            const int ta_lo = max(ta0, tau + a.pad_t - (it.t1 - 1));
                            const int ta_hi = min(ta0 + gn - 1, tau + a.pad_t - it.t0);
                            const int nfr = ta_hi - ta_lo + 1;                      // output frames hit by this input frame
                            const int slot = it.t1 - 1 - (tau + a.pad_t - ta_lo);   // column block of the latest of them
                            const uint32_t range = ((1u << nfr) - 1u) << slot;
                            const bool split = (range & ~touched) != 0;
                            const uint32_t d_tmem = d_base + slot * NC;
                            const uint32_t idesc = idesc0 | (uint32_t(nfr * (NC >> 3)) << 17);
                            const uint32_t b_row0 = uint32_t(ta_lo - ta0) * NC * ROW;
                            mbar_wait(&a_full[as], aphase);
                            const uint32_t a_addr = smem_u32(smem_a + as * a.a_stage_bytes);
#pragma unroll
                            for (int s = 0; s < NSP; ++s) {
                                const int ti = s / KW, tj = s - ti * KW;
                                if (tau == lo && !(a.b_resident && b_loaded)) mbar_wait(&b_full[s], bphase);
                                tc_fence_after();
                                const uint64_t adesc = adesc0 + ((a_addr + (ti * a.LP + tj) * ROW) >> 4);
                                const uint64_t bdesc = bdesc0 + ((b_base + s * a.piece_bytes + b_row0) >> 4);
#pragma unroll
                                for (int k = 0; k < BK / 16; ++k) {
                                    if (s == 0 && k == 0 && split) {
                                        // first products of some of these accumulators: one MMA per run of equal state
                                        int i = 0;
                                        while (i < nfr) {
                                            const uint32_t bit = (touched >> (slot + i)) & 1u;
                                            int j = i + 1;
                                            while (j < nfr && ((touched >> (slot + j)) & 1u) == bit) ++j;
                                            umma_bf16(d_tmem + i * NC, adesc, bdesc + ((uint32_t(i) * NC * ROW) >> 4),
                                                      idesc0 | (uint32_t((j - i) * (NC >> 3)) << 17), bit);
                                            i = j;
                                        }
                                    } else {
                                        umma_bf16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, 1u);
                                    }
                                }
                                if (tau == hi && !a.b_resident) umma_commit(&b_empty[s]);
                            }
                            touched |= range;
                            b_loaded = true;
                            umma_commit(&a_empty[as]);
                            if (++as == a.a_stages) { as = 0; aphase ^= 1; }
                        }
                        bphase ^= 1;
                    }
                }
                umma_commit(&tmem_full[acc]);
                if (++acc == a.nbuf) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else if (warp >= EPI_WARP0) {
        // ------------------------------ epilogue ------------------------------
        const int ew = warp - EPI_WARP0;
        const int q = ew & 3;                            // TMEM lane quarter == warp id % 4
        const int par = ew >> 2;                         // this warp's frames: t0 + par, t0 + par + 2, ...
        const int r = q * 32 + lane;                     // accumulator row = pixel within the tile
        const int hl = r / TW, wl = r - hl * TW;
        const bool do_stats = (a.sum != nullptr);
        const bool affine = (a.scale != nullptr) || (a.shift != nullptr);
        const bool staged = a.stage_mode != 0;
        const long long frame_pix = (long long)a.H * a.W;
        float* stage = s_stage + ew * (EPI_STAGE_BYTES / 4);
        EpiOut eo;
        eo.y = a.y; eo.y_cstride = a.y_cstride; eo.y_bf16 = a.y_bf16; eo.relu = a.relu; eo.accumulate = a.accumulate;
        eo.relu_mask = nullptr; eo.mask_cstride = 0;
        eo.addend = a.addend; eo.addend_cstride = a.addend_cstride; eo.addend_bf16 = a.addend_bf16;
        int acc = 0;
        uint32_t acc_phase = 0;
        // BatchNorm statistics: in the transposed ownership of the epilogue (common.cuh) a lane sees 4 channels of 8 pixels per
        // frame, so its column sums of x and x^2 are 8 REGISTERS that live across every frame of every work item of the CTA; the
        // cross-lane reduction (a 6-shuffle butterfly) runs ONCE at the end.  (Round 2, first version: per-(item, frame)
        // transpose-reduces were 314 of the epilogue's 407 instructions per frame at 0.15 IPC.)
        float cs[4] = {0.f, 0.f, 0.f, 0.f}, cq[4] = {0.f, 0.f, 0.f, 0.f};
        for (int item = blockIdx.x; item < a.nitems; item += gridDim.x) {
            const Item it = decode_item(a, item);
            const int h = it.h0 + hl, w = it.w0 + wl;
            const bool valid = (h < a.H) && (w < a.W);
            const long long pix0 = (((long long)it.b * a.To + it.t0) * a.H + h) * a.W + w;      // this lane's pixel in frame t0
            const EpiRows rows0 = epi_rows(valid ? (int)pix0 : -1, lane);
            mbar_wait(&tmem_full[acc], acc_phase);
            tc_fence_after();
            const uint32_t t_addr = tmem_base + (uint32_t(q * 32) << 16) + acc * a.Fg * NC;
            for (int t = it.t0 + par; t < (a.dbg & 1 ? it.t0 : it.t1); t += 2) {
                EpiRows rt;
#pragma unroll
                for (int i = 0; i < 8; ++i) rt.pix[i] = rows0.pix[i] >= 0 ? rows0.pix[i] + (t - it.t0) * (int)frame_pix : -1;
                uint32_t v[32];
                tmem_ld_32x32(t_addr + (it.t1 - 1 - t) * NC, v);
                tmem_ld_wait();
                if (staged) {
                    float4 x[8];
                    epi_transpose(stage, v, lane, x);
                    if (do_stats) epi_colsum(x, rt, cs, cq);
                    // (issuing the addend loads before the TMEM load costs 32 more live registers: spills, 94 -> 112 us)
                    uint4 ad[8];
                    epi_addend_load(ad, rt, 4 * (lane & 7), eo);
                    epi_store(x, rt, 4 * (lane & 7), affine ? s_scale : nullptr, affine ? s_shift : nullptr, eo, ad);
                } else if (valid) {
                    // bf16 outputs without statistics (eval-mode folded layers): a thread owns one pixel's 64 bytes
                    const long long pix = pix0 + (t - it.t0) * frame_pix;
                    float o[32];
                    if (affine) {           // per-channel scale / shift from shared memory, 16 bytes per load
                        const float4* sc4 = reinterpret_cast<const float4*>(s_scale);
                        const float4* sh4 = reinterpret_cast<const float4*>(s_shift);
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
                    uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(a.y) + pix * a.y_cstride);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        uint4 u;
                        u.x = pack_bf16x2(o[8 * j + 0], o[8 * j + 1]);
                        u.y = pack_bf16x2(o[8 * j + 2], o[8 * j + 3]);
                        u.z = pack_bf16x2(o[8 * j + 4], o[8 * j + 5]);
                        u.w = pack_bf16x2(o[8 * j + 6], o[8 * j + 7]);
                        dst[j] = u;
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(&tmem_empty[acc]);
            if (++acc == a.nbuf) { acc = 0; acc_phase ^= 1; }
        }
        if (do_stats) {
            epi_colsum_flush(cs, cq, lane, s_sum, s_sq, 4 * (lane & 7));
            asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory");
            const int i = threadIdx.x - EPI_WARP0 * 32;
            if (i < NC) {
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

int env_int(const char* name, int dflt) {
    const char* v = getenv(name);
    return v ? atoi(v) : dflt;
}

}  // namespace

// 1 if sfvos_conv_umma should hand this problem to the temporally-stacked kernel.
int sfvos_conv_tstack_applicable(const sfvos_conv_params* p) {
    if (!env_int("SFVOS_TSTACK", 1)) return 0;
    if (p->N != NC || p->relu_mask != nullptr) return 0;
    const bool k3 = p->kh == 3 && p->kw == 3 && p->pad_h == 1 && p->pad_w == 1;
    const bool k1 = p->kh == 1 && p->kw == 1 && p->pad_h == 0 && p->pad_w == 0 && p->kt > 1;    // lateral dgrad
    if (!k3 && !k1) return 0;
    if ((p->OH && p->OH != p->H) || (p->OW && p->OW != p->W)) return 0;
    if ((p->oy_mul && p->oy_mul != 1) || (p->ox_mul && p->ox_mul != 1) || p->oy_off || p->ox_off) return 0;
    if (p->kt > 64 || p->To > 4096) return 0;
    return 1;
}

int sfvos_conv_tstack_launch(const sfvos_conv_params* p, cudaStream_t stream) {
    const int BK = (p->Cp % 64 == 0) ? 64 : 32;
    const int ROW = BK * 2;
    const int NSP = (int)(p->kh * p->kw);                   // 9 or 1
    const int HALO = NSP == 9 ? 1 : 0;
    TsArgs a;
    a.B = (int)p->B; a.To = (int)p->To; a.Ti = (int)p->T; a.H = (int)p->H; a.W = (int)p->W;
    a.tiles_w = (a.W + TW - 1) / TW;
    const int tiles_h = (a.H + TH - 1) / TH;
    a.tiles_per_frame = a.tiles_w * tiles_h;
    a.kt = (int)p->kt; a.pad_t = (int)p->pad_t;
    a.cchunks = (int)(p->Cp / BK);
    // output frames per work item: 2 accumulator sets of Fg*32 columns must fit the 512 TMEM columns
    a.nfg = (a.To + 7) / 8;
    a.Fg = (a.To + a.nfg - 1) / a.nfg;
    a.nfg = (a.To + a.Fg - 1) / a.Fg;
    a.nbuf = 2;
    uint32_t cols = 32;
    while (cols < (uint32_t)(a.nbuf * a.Fg * NC)) cols <<= 1;
    a.tmem_cols = cols;
    a.nitems = a.B * a.nfg * a.tiles_per_frame;
    // halo line pitch in smem rows: 10 = exact (tile + 2), 16 = padded
    a.LP = HALO ? env_int("SFVOS_TSTACK_LP", 10) : TW;
    SF_CHECK(a.LP >= TW + 2 * HALO && a.LP <= 16, "conv_tstack: SFVOS_TSTACK_LP=%d out of range", a.LP);
    a.a_tx_bytes = (uint32_t)(a.LP * (TH + 2 * HALO) * ROW);
    a.a_stage_bytes = (int)((a.a_tx_bytes + 1023u) & ~1023u);
    // temporal taps stacked per MMA: all 9 spatial pieces of a (tap group, chunk) stay resident next to >= 2 (3) A stages
    const int smem_budget = 227 * 1024 - 1024 /*align*/ - SMALL_BYTES - STAGE_RESERVE /*epilogue transpose buffers*/;
    int gmax = env_int("SFVOS_TSTACK_G", BK == 64 ? 3 : 8);
    if (gmax < 1) gmax = 1;
    if (gmax > 8) gmax = 8;
    while (gmax > 1 && smem_budget - NSP * gmax * NC * ROW < 2 * a.a_stage_bytes) --gmax;
    a.ngroups = (a.kt + gmax - 1) / gmax;
    a.G = (a.kt + a.ngroups - 1) / a.ngroups;
    a.ngroups = (a.kt + a.G - 1) / a.G;
    a.piece_bytes = a.G * NC * ROW;
    a.b_resident = (a.ngroups == 1 && a.cchunks == 1) ? 1 : 0;
    a.a_stages = (smem_budget - NSP * a.piece_bytes) / a.a_stage_bytes;
    if (a.a_stages > 8) a.a_stages = 8;
    SF_CHECK(a.a_stages >= 2, "conv_tstack: not enough shared memory");
    a.y = p->y; a.y_bf16 = (p->y_dtype == SFVOS_BF16); a.relu = p->relu; a.accumulate = p->accumulate;
    a.y_cstride = p->y_cstride;
    a.scale = p->scale; a.shift = p->shift; a.sum = p->sum; a.sumsq = p->sumsq;
    // f32 outputs, statistics and addends leave through the transposing epilogue; bf16 outputs without them (eval-mode folded
    // layers, bf16 partial gradients) store row-per-lane - measured faster there (fast_conv2 dgrad 138 vs 147 us, fast_conv3
    // 66 vs 74); SFVOS_TSTACK_STAGE=1 forces the transpose for A/B measurements
    a.addend = p->addend; a.addend_cstride = p->addend_cstride; a.addend_bf16 = (p->addend_dtype == SFVOS_BF16);
    a.stage_mode = (!a.y_bf16 || a.sum != nullptr || a.addend != nullptr || env_int("SFVOS_TSTACK_STAGE", 0) >= 1) ? 1 : 0;
    SF_CHECK(p->B * p->To * p->H * p->W < (1LL << 31), "conv_tstack: too many output pixels");

    a.dbg = env_int("SFVOS_TSTACK_DBG", 0);

    CUtensorMap tx, tw;
    int rc;
    {
        uint64_t dims[5] = {(uint64_t)p->C, (uint64_t)p->W, (uint64_t)p->H, (uint64_t)p->T, (uint64_t)p->B};
        const uint64_t cs = (uint64_t)p->x_cstride;
        const uint64_t hs = p->x_hstride ? (uint64_t)p->x_hstride : cs * p->W;
        const uint64_t ts = p->x_tstride ? (uint64_t)p->x_tstride : hs * p->H;
        const uint64_t bs = p->x_bstride ? (uint64_t)p->x_bstride : ts * p->T;
        uint64_t str[4] = {cs * 2, hs * 2, ts * 2, bs * 2};
        uint32_t box[5] = {(uint32_t)BK, (uint32_t)a.LP, (uint32_t)(TH + 2 * HALO), 1, 1};
        rc = sfvos_make_tmap(&tx, p->x, 5, dims, str, box, ROW);
        if (rc) return rc;
    }
    {
        // packed weights [N][kt][9][Cp] viewed as {Cp, N, kt, 9}: a box {BK, 32, G, 1} lands as G stacked [32 x BK] tiles
        const uint64_t taps = (uint64_t)(p->kt * NSP);
        uint64_t dims[4] = {(uint64_t)p->Cp, (uint64_t)NC, (uint64_t)p->kt, (uint64_t)NSP};
        uint64_t str[3] = {taps * p->Cp * 2, (uint64_t)NSP * p->Cp * 2, (uint64_t)p->Cp * 2};
        uint32_t box[4] = {(uint32_t)BK, (uint32_t)NC, (uint32_t)a.G, 1};
        rc = sfvos_make_tmap(&tw, p->w, 4, dims, str, box, ROW);
        if (rc) return rc;
    }
    const int smem_bytes = a.a_stages * a.a_stage_bytes + NSP * a.piece_bytes + 1024 + SMALL_BYTES + STAGE_RESERVE;
    int grid = sfvos_num_sms();
    if (grid > a.nitems) grid = a.nitems;
#define SF_TSTACK_LAUNCH(BK_, NSP_)                                                                                  \
    do {                                                                                                             \
        SF_CUDA(cudaFuncSetAttribute(conv_tstack_kernel<BK_, NSP_>, cudaFuncAttributeMaxDynamicSharedMemorySize,     \
                                     227 * 1024));                                                                   \
        conv_tstack_kernel<BK_, NSP_><<<grid, NUM_THREADS, smem_bytes, stream>>>(tx, tw, a);                         \
    } while (0)
    if (BK == 64 && NSP == 9) SF_TSTACK_LAUNCH(64, 9);
    else if (BK == 32 && NSP == 9) SF_TSTACK_LAUNCH(32, 9);
    else if (BK == 64) SF_TSTACK_LAUNCH(64, 1);
    else SF_TSTACK_LAUNCH(32, 1);
#undef SF_TSTACK_LAUNCH
    sfvos_set_kernel("conv_tstack");
    SF_LAUNCH_CHECK();
    return SFVOS_OK;
}
