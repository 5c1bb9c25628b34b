// CTA-pair (tcgen05 cta_group::2) implicit-GEMM 3x3 convolution for the WIDE layers (N >= 128 output channels):
// slow_conv1/2/3 fprop and dgrad (code/helpers/model.py:50-51,56-57,62-63) and the mask-head convolutions
// (TV/models/detection/mask_rcnn.py:284-296).
//
// conv_umma.cu is bound by the L2 -> SMEM fill path on these layers: every CTA streams ALL the weights (9 taps x
// N x 64 ch = 216-288 KB per 64-channel chunk) past a 23-36 KB activation box, ~10 TB/s chip-wide.  Here two CTAs of a
// cluster (the two SMs of a TPC) compute two adjacent 128-pixel tiles as ONE 256 x N x 16 MMA:
//   * each CTA loads its own activation box (tile + halo) and only HALF of every weight tile (N/2 rows);
//     tcgen05.mma.cta_group::2 reads A from both CTAs (128 rows each) and B from both (N/2 columns each) and writes
//     each CTA's 128 x N accumulator into its own TMEM  -> weight bytes ingested per SM halve;
//   * CTA 0 issues the MMAs.  Both CTAs' TMA loads complete on CTA 0's "full" barriers (cp.async.bulk.tensor
//     .cta_group::2), MMA completion is multicast to both CTAs' "empty" / "accumulator full" barriers
//     (tcgen05.commit ... multicast::cluster), and CTA 1's epilogue warps release the accumulator on CTA 0's barrier
//     through its cluster address (mapa);
//   * everything else as conv_umma's halo mode: 9 spatial taps = 9 shifted UMMA descriptors into one halo box,
//     double-buffered accumulators, epilogue warps doing BN statistics / affine / ReLU / stores.
#include <stdlib.h>

#include "common.cuh"

namespace {

constexpr int BM = 128;                     // rows per CTA (the MMA is 256 x N)
constexpr int BK = 64;
constexpr uint32_t ROW = BK * 2;
constexpr int NUM_THREADS = 384;             // warps 0-3: TMA / MMA / TMEM alloc / idle; warps 4-11: epilogue
constexpr int EPI_WARP0 = 4;
constexpr int EPI_THREADS = 256;             // TWO warps per TMEM lane quarter, each draining half of the accumulator's columns
constexpr int TW = 8, TH = 16;

struct PairArgs {
    int B, To, H, W;
    int tiles_w, tiles_h, ntiles, npairs;
    int N, kt, pad_t, cchunks;
    int LP, a_stages, b_stages, a_stage_bytes, b_half_bytes;
    int epi_stage;                               // epilogue through the shared-memory transpose (common.cuh: epi_block)
    uint32_t idesc, tmem_cols, a_tx_bytes;
    void* y;
    int y_bf16, relu, accumulate;
    long long y_cstride;
    const float* scale;
    const float* shift;
    float* sum;
    float* sumsq;
    const __nv_bfloat16* relu_mask;          // fused ReLU backward (see sfvos_conv_params)
    long long mask_cstride;
    const void* addend;                      // y = act(...) + addend (see sfvos_conv_params)
    long long addend_cstride;
    int addend_bf16;
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `p` (an address in this CTA's shared memory) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t map_to_cta(const void* p, uint32_t rank) {
    uint32_t out;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(out) : "r"(smem_u32(p)), "r"(rank));
    return out;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* smem_dst, uint32_t ncols) {     // one full warp, in BOTH CTAs
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// TMA loads whose completion is signalled on a barrier of the pair's leader CTA (cluster address)
__device__ __forceinline__ void tma_load_5d_pair(void* dst, const CUtensorMap* m, uint32_t bar_cluster, int c0, int c1, int c2,
                                                 int c3, int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d_pair(void* dst, const CUtensorMap* m, uint32_t bar_cluster, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2)
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
// arrive (once all MMAs issued so far are complete) on the barrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
    asm volatile(
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
        ::"r"(smem_u32(bar)), "h"((uint16_t)3)
        : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
conv_pair_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w, const PairArgs a) {
    constexpr uint32_t LAYOUT = 2u;                           // SWIZZLE_128B
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + a.a_stages * a.a_stage_bytes;
    uint64_t* a_full = reinterpret_cast<uint64_t*>(smem_b + a.b_stages * a.b_half_bytes);
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
    const uint32_t rank = cluster_ctarank();                  // 0 = leader (issues the MMAs), 1 = peer
    const int pair0 = blockIdx.x >> 1, pair_stride = gridDim.x >> 1;

    if (warp == 0 && elect_one()) {
        tma_prefetch_desc(&tmap_x);
        tma_prefetch_desc(&tmap_w);
    }
    if (warp == 1 && elect_one()) {
        for (int i = 0; i < a.a_stages; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
        for (int i = 0; i < a.b_stages; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], 16); }     // 8 epilogue warps x 2 CTAs
        fence_mbar_init();
    }
    if (warp == 2) tmem_alloc_pair(tmem_slot, a.tmem_cols);
    for (int i = threadIdx.x; i < 256; i += NUM_THREADS) {
        s_scale[i] = (a.scale && i < a.N) ? a.scale[i] : 1.0f;
        s_shift[i] = (a.shift && i < a.N) ? a.shift[i] : 0.0f;
        s_sum[i] = 0.0f;
        s_sq[i] = 0.0f;
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                                       // the peer's barriers are initialised before anyone signals them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int tiles_per_frame = a.tiles_w * a.tiles_h;

    if (warp == 0) {
        if (elect_one()) {
            // ------------------------------ TMA producer (both CTAs) ------------------------------
            int as = 0, bs = 0;
            uint32_t aphase = 0, bphase = 0;
            for (int pair = pair0; pair < a.npairs; pair += pair_stride) {
                const int tile = 2 * pair + (int)rank;
                int b = a.B, t = 0, h0 = 0, w0 = 0;           // past-the-end tile (odd tile count): fully out of range -> zeros
                if (tile < a.ntiles) {
                    const int frame = tile / tiles_per_frame;
                    const int rem = tile - frame * tiles_per_frame;
                    const int th_i = rem / a.tiles_w;
                    b = frame / a.To; t = frame - b * a.To;
                    h0 = th_i * TH; w0 = (rem - th_i * a.tiles_w) * TW;
                }
                for (int ta = 0; ta < a.kt; ++ta)
                    for (int cc = 0; cc < a.cchunks; ++cc) {
                        mbar_wait(&a_empty[as], aphase ^ 1);
                        if (rank == 0) mbar_arrive_expect_tx(&a_full[as], 2 * a.a_tx_bytes);
                        tma_load_5d_pair(smem_a + as * a.a_stage_bytes, &tmap_x, map_to_cta(&a_full[as], 0), cc * BK, w0 - 1, h0 - 1,
                                         t + ta - a.pad_t, b);
                        if (++as == a.a_stages) { as = 0; aphase ^= 1; }
                        for (int sp = 0; sp < 9; ++sp) {
                            mbar_wait(&b_empty[bs], bphase ^ 1);
                            if (rank == 0) mbar_arrive_expect_tx(&b_full[bs], 2 * a.b_half_bytes);
                            tma_load_3d_pair(smem_b + bs * a.b_half_bytes, &tmap_w, map_to_cta(&b_full[bs], 0), cc * BK,
                                             (int)rank * (a.N / 2), ta * 9 + sp);
                            if (++bs == a.b_stages) { bs = 0; bphase ^= 1; }
                        }
                    }
            }
        }
    } else if (warp == 1) {
        if (rank == 0 && elect_one()) {
            // ------------------------------ MMA issuer (leader CTA only) ------------------------------
            int as = 0, bs = 0;
            uint32_t aphase = 0, bphase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            const uint64_t adesc0 = umma_smem_desc(0, 16, a.LP * ROW, LAYOUT);
            const uint64_t bdesc0 = umma_smem_desc(0, 16, 8 * ROW, LAYOUT);
            const int outer = a.kt * a.cchunks;
            for (int pair = pair0; pair < a.npairs; pair += pair_stride) {
                mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * a.N;
                uint32_t accum = 0;
                for (int o = 0; o < outer; ++o) {
                    mbar_wait(&a_full[as], aphase);
                    const uint32_t a_addr = smem_u32(smem_a + as * a.a_stage_bytes);
#pragma unroll
                    for (int sp = 0; sp < 9; ++sp) {
                        const int ti = sp / 3, tj = sp - ti * 3;
                        mbar_wait(&b_full[bs], bphase);
                        tc_fence_after();
                        const uint64_t adesc = adesc0 + ((a_addr + (ti * a.LP + tj) * ROW) >> 4);
                        const uint64_t bdesc = bdesc0 + (smem_u32(smem_b + bs * a.b_half_bytes) >> 4);
#pragma unroll
                        for (int k = 0; k < BK / 16; ++k) {
                            umma_bf16_pair(d_tmem, adesc + 2 * k, bdesc + 2 * k, a.idesc, accum);
                            accum = 1;
                        }
                        umma_commit_pair(&b_empty[bs]);
                        if (++bs == a.b_stages) { bs = 0; bphase ^= 1; }
                    }
                    umma_commit_pair(&a_empty[as]);
                    if (++as == a.a_stages) { as = 0; aphase ^= 1; }
                }
                umma_commit_pair(&tmem_full[acc]);
                acc ^= 1;
                if (acc == 0) acc_phase ^= 1;
            }
        }
    } else if (warp >= EPI_WARP0) {
        // ------------------------------ epilogue (both CTAs, own tile) ------------------------------
        // Two warps share every TMEM lane quarter (a warp reads lanes 32 * (warp % 4) ..): the first drains the lower half of
        // the accumulator's column blocks, the second the upper half.  With one warp per scheduler the epilogue ran its
        // dependent shuffle / select chains (fused BN statistics: ~310 instructions per 32-column block) at ~0.15 IPC and a
        // statistics-carrying fprop tile took about as long to drain as to compute (ncu, round 2: 67 % tensor-pipe-active with
        // statistics against 79 % without); two warps per scheduler halve the drain and hide each other's latencies.
        const int ew = warp - EPI_WARP0;
        const int q = ew & 3;
        const int c_split = ((a.N / 32 + 1) / 2) * 32;
        const int c_begin = (ew >> 2) ? c_split : 0, c_end = (ew >> 2) ? a.N : c_split;
        const int r = q * 32 + lane;
        const int hl = r / TW, wl = r - hl * TW;
        const bool do_stats = (a.sum != nullptr) && (a.relu_mask == nullptr);
        const bool affine = (a.scale != nullptr) || (a.shift != nullptr);
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int pair = pair0; pair < a.npairs; pair += pair_stride) {
            const int tile = 2 * pair + (int)rank;
            const int frame = tile / tiles_per_frame;
            const int rem = tile - frame * tiles_per_frame;
            const int th_i = rem / a.tiles_w;
            const int h = th_i * TH + hl, w = (rem - th_i * a.tiles_w) * TW + wl;
            const bool valid = (tile < a.ntiles) && (h < a.H) && (w < a.W);
            const long long pix = ((long long)frame * a.H + h) * a.W + w;
            mbar_wait(&tmem_full[acc], acc_phase);
            tc_fence_after();
            const uint32_t t_addr = tmem_base + (uint32_t(q * 32) << 16) + acc * a.N;
            if (a.epi_stage) {
                const EpiRows rows = epi_rows(valid ? (int)pix : -1, lane);
                EpiOut eo;
                eo.y = a.y; eo.y_cstride = a.y_cstride; eo.y_bf16 = a.y_bf16; eo.relu = a.relu; eo.accumulate = a.accumulate;
                eo.relu_mask = a.relu_mask; eo.mask_cstride = a.mask_cstride;
                eo.addend = a.addend; eo.addend_cstride = a.addend_cstride; eo.addend_bf16 = a.addend_bf16;
                float* sum_dst = (do_stats || (a.relu_mask != nullptr && a.sum != nullptr)) ? s_sum : nullptr;
                for (int c0 = c_begin; c0 < c_end; c0 += 32) {
                    uint32_t v[32];
                    tmem_ld_32x32(t_addr + c0, v);
                    tmem_ld_wait();
                    epi_block(s_stage + ew * (EPI_STAGE_BYTES / 4), v, rows, lane, c0, affine ? s_scale : nullptr,
                              affine ? s_shift : nullptr, eo, sum_dst, do_stats ? s_sq : nullptr);
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
                    // fused ReLU backward of the layer below: zero the gradient where its activation is not positive, and
                    // reduce the column sums of the masked gradient (= that layer's bias gradient)
                    float f[32];
                    if (valid) {
                        const uint4* mp = reinterpret_cast<const uint4*>(a.relu_mask + pix * a.mask_cstride + c0);
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const uint4 m = __ldg(mp + j);
                            const uint32_t mw[4] = {m.x, m.y, m.z, m.w};
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                // bf16 > 0  <=>  sign bit clear and not zero
                                const uint32_t lo = mw[e] & 0xffffu, hi = mw[e] >> 16;
                                const bool p0 = lo != 0 && lo < 0x8000u, p1 = hi != 0 && hi < 0x8000u;
                                if (!p0) v[8 * j + 2 * e] = 0u;
                                if (!p1) v[8 * j + 2 * e + 1] = 0u;
                            }
                        }
                    }
                    if (a.sum != nullptr) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) f[j] = valid ? __uint_as_float(v[j]) : 0.0f;     // fp32, like relu_bwd
                        const float s = warp_transpose_reduce32(f, lane);
                        atomicAdd(&s_sum[c0 + lane], s);
                    }
                }
                if (valid) {
                    float o[32];
                    if (affine) {           // per-channel scale / shift from shared memory, 16 bytes per load
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
                        uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(a.y) + pix * a.y_cstride + c0);
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
                        float4* dst = reinterpret_cast<float4*>(reinterpret_cast<float*>(a.y) + pix * a.y_cstride + c0);
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
            // this warp's TMEM reads are done: release the accumulator to the leader's MMA thread (one arrive per warp)
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (rank == 0) mbar_arrive(&tmem_empty[acc]);
                else mbar_arrive_cluster(map_to_cta(&tmem_empty[acc], 0));
            }
            acc ^= 1;
            if (acc == 0) acc_phase ^= 1;
        }
        if (do_stats || (a.relu_mask != nullptr && a.sum != nullptr)) {
            asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory");
            for (int i = threadIdx.x - EPI_WARP0 * 32; i < a.N; i += EPI_THREADS) {
                atomicAdd(&a.sum[i], s_sum[i]);
                if (do_stats) atomicAdd(&a.sumsq[i], s_sq[i]);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                                       // nobody frees TMEM / exits while the pair still uses it
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc_pair(tmem_base, a.tmem_cols);
    }
}

int env_int(const char* name, int dflt) {
    const char* v = getenv(name);
    return v ? atoi(v) : dflt;
}

}  // namespace

int sfvos_conv_pair_applicable(const sfvos_conv_params* p) {
    if (!env_int("SFVOS_PAIR", 1)) return 0;
    if (p->kh != 3 || p->kw != 3 || p->pad_h != 1 || p->pad_w != 1) return 0;
    if (p->N < 128 || p->N > 256 || p->N % 32 != 0 || p->Cp % 64 != 0) return 0;
    if ((p->OH && p->OH != p->H) || (p->OW && p->OW != p->W)) return 0;
    if ((p->oy_mul && p->oy_mul != 1) || (p->ox_mul && p->ox_mul != 1) || p->oy_off || p->ox_off) return 0;
    // 8 x 16 tiles must not waste much more than the best free-form tiling (tiny feature maps stay on conv_umma)
    const long long tiles = (long long)((p->W + TW - 1) / TW) * ((p->H + TH - 1) / TH);
    const double eff = (double)p->H * p->W / (double)(tiles * BM);
    return eff >= 0.6;
}

int sfvos_conv_pair_launch(const sfvos_conv_params* p, cudaStream_t stream) {
    PairArgs a;
    a.B = (int)p->B; a.To = (int)p->To; a.H = (int)p->H; a.W = (int)p->W;
    a.tiles_w = (a.W + TW - 1) / TW;
    a.tiles_h = (a.H + TH - 1) / TH;
    a.ntiles = a.B * a.To * a.tiles_w * a.tiles_h;
    a.npairs = (a.ntiles + 1) / 2;
    a.N = (int)p->N; a.kt = (int)p->kt; a.pad_t = (int)p->pad_t;
    a.cchunks = (int)(p->Cp / BK);
    a.LP = env_int("SFVOS_TSTACK_LP", 10);
    SF_CHECK(a.LP >= TW + 2 && a.LP <= 16, "conv_pair: SFVOS_TSTACK_LP=%d out of range", a.LP);
    a.a_tx_bytes = (uint32_t)(a.LP * (TH + 2) * ROW);
    a.a_stage_bytes = (int)((a.a_tx_bytes + 1023u) & ~1023u);
    a.b_half_bytes = (a.N / 2) * (int)ROW;
    const int small_bytes = 8192 /*barriers, scale/shift, stats*/ + (EPI_THREADS / 32) * EPI_STAGE_BYTES /*epilogue transpose tiles*/;
    const int smem_budget = 227 * 1024 - 1024 /*align*/ - small_bytes;
    // Measured (B200, level 0, B = 8): f32 outputs gain from the transposed epilogue (slow_conv1 fprop with statistics 350 -> 310 us,
    // its f32 dgrad 340 -> 299 us); bf16 outputs without statistics do not (mask-head conv 191 -> 203 us) - this kernel runs at the
    // shared-memory port's limit and the transpose tile adds 8 KB of traffic per 32-column block.
    const int es = env_int("SFVOS_EPI_STAGE", -1);
    a.epi_stage = es < 0 ? (p->y_dtype != SFVOS_BF16 || p->addend != nullptr) : (es != 0);
    a.addend = p->addend; a.addend_cstride = p->addend_cstride; a.addend_bf16 = (p->addend_dtype == SFVOS_BF16);
    SF_CHECK(p->addend == nullptr || a.epi_stage, "conv_pair: addend needs the transposing epilogue (SFVOS_EPI_STAGE)");
    // Pipeline depth (measured, round 2, slow_conv1 fprop / slow_conv2 dgrad / slow_conv3 fprop at level 0, same box):
    // 3 A x 10 B stages 360 / 347 / 406 us, 2 x 6 353 / 335 / 395, 2 x 8 346 / 331 / 394, 2 x 16 361 / 344 / 406, 2 x 3 380 / 365 / 422.
    // Deeper prefetch does not help - the MMA thread's waits for weight stages are L2 -> SM BANDWIDTH (every pixel tile
    // re-streams the layer's weights: 2.16 GB per launch against ~6300 B/clk chip-wide), not latency - and more bytes in flight
    // make it slightly worse, so: 2 activation stages, at most 8 weight stages.
    a.a_stages = env_int("SFVOS_PAIR_ASTAGES", 2);
    if (a.a_stages < 2) a.a_stages = 2;
    if (a.a_stages > 4) a.a_stages = 4;
    a.b_stages = (smem_budget - a.a_stages * a.a_stage_bytes) / a.b_half_bytes;
    const int b_cap = env_int("SFVOS_PAIR_BSTAGES", 8);
    if (a.b_stages > b_cap) a.b_stages = b_cap;
    SF_CHECK(a.b_stages >= 3, "conv_pair: not enough shared memory");
    a.idesc = umma_idesc_bf16(2 * BM, a.N, 0, 0);
    uint32_t cols = 32;
    while (cols < (uint32_t)(2 * a.N)) cols <<= 1;
    a.tmem_cols = cols;
    a.y = p->y; a.y_bf16 = (p->y_dtype == SFVOS_BF16); a.relu = p->relu; a.accumulate = p->accumulate;
    a.y_cstride = p->y_cstride;
    a.scale = p->scale; a.shift = p->shift; a.sum = p->sum; a.sumsq = p->sumsq;
    a.relu_mask = reinterpret_cast<const __nv_bfloat16*>(p->relu_mask); a.mask_cstride = p->relu_mask_cstride;

    CUtensorMap tx, tw;
    int rc;
    {
        uint64_t dims[5] = {(uint64_t)p->C, (uint64_t)p->W, (uint64_t)p->H, (uint64_t)p->T, (uint64_t)p->B};
        const uint64_t cs = (uint64_t)p->x_cstride;
        const uint64_t hs = p->x_hstride ? (uint64_t)p->x_hstride : cs * p->W;
        const uint64_t ts = p->x_tstride ? (uint64_t)p->x_tstride : hs * p->H;
        const uint64_t bs = p->x_bstride ? (uint64_t)p->x_bstride : ts * p->T;
        uint64_t str[4] = {cs * 2, hs * 2, ts * 2, bs * 2};
        uint32_t box[5] = {(uint32_t)BK, (uint32_t)a.LP, (uint32_t)(TH + 2), 1, 1};
        rc = sfvos_make_tmap(&tx, p->x, 5, dims, str, box, ROW);
        if (rc) return rc;
    }
    {
        // packed weights [N][taps][Cp] viewed as {Cp, N, taps}: a box {BK, N/2, 1} is one CTA's half of a tap's tile
        const uint64_t taps = (uint64_t)(p->kt * 9);
        uint64_t dims[3] = {(uint64_t)p->Cp, (uint64_t)p->N, taps};
        uint64_t str[2] = {taps * p->Cp * 2, (uint64_t)p->Cp * 2};
        uint32_t box[3] = {(uint32_t)BK, (uint32_t)(p->N / 2), 1};
        rc = sfvos_make_tmap(&tw, p->w, 3, dims, str, box, ROW);
        if (rc) return rc;
    }
    const int smem_bytes = a.a_stages * a.a_stage_bytes + a.b_stages * a.b_half_bytes + 1024 + small_bytes;
    int clusters = sfvos_num_sms() / 2;
    if (clusters > a.npairs) clusters = a.npairs;
    SF_CUDA(cudaFuncSetAttribute(conv_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    conv_pair_kernel<<<2 * clusters, NUM_THREADS, smem_bytes, stream>>>(tx, tw, a);
    sfvos_set_kernel("conv_pair");
    SF_LAUNCH_CHECK();
    return SFVOS_OK;
}
