// Shared device/host helpers for the sm_100a kernels: error plumbing, mbarrier / TMA / tcgen05 PTX wrappers.
// Everything here is hand-written inline PTX for sm_100a (no CUTLASS, no Triton).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/sfvos.h"

// ----------------------------------------------------------------------------------------------------------
// host-side error plumbing: every extern "C" entry returns 0 on success, !=0 on failure with a message
// retrievable through sfvos_last_error().
// ----------------------------------------------------------------------------------------------------------
void sfvos_set_error(const char* fmt, ...);

#define SF_CHECK(cond, ...)                                                                       \
    do { if (!(cond)) { sfvos_set_error(__VA_ARGS__); return SFVOS_ERR_INVALID; } } while (0)

#define SF_CUDA(expr)                                                                             \
    do { cudaError_t _e = (expr); if (_e != cudaSuccess) {                                        \
        sfvos_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
        return SFVOS_ERR_CUDA; } } while (0)

#define SF_LAUNCH_CHECK()                                                                         \
    do { cudaError_t _e = cudaGetLastError(); if (_e != cudaSuccess) {                            \
        sfvos_set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
        return SFVOS_ERR_CUDA; } } while (0)

int sfvos_num_sms();
void sfvos_set_kernel(const char* name);      // records which kernel an entry point dispatched to (sfvos_last_kernel)

// Tensor-map (TMA descriptor) construction, tmap.cu.  dims/strides innermost first; strides in BYTES for
// dims 1..rank-1.  swizzle: 0 none, 64, 128.
int sfvos_make_tmap(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                    const uint64_t* strides_bytes, const uint32_t* box, int swizzle_bytes);

// ----------------------------------------------------------------------------------------------------------
// device helpers
// ----------------------------------------------------------------------------------------------------------
#ifdef __CUDACC__

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0;
}

// ---- mbarrier ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a pipeline bug must fail loudly (trap -> CUDA error on the host) instead of hanging the GPU box.
#ifndef SFVOS_WATCHDOG_CYCLES
#define SFVOS_WATCHDOG_CYCLES (4000000000LL)   // ~2 s at 1.9 GHz
#endif
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > SFVOS_WATCHDOG_CYCLES) {
            printf("sfvos: mbarrier watchdog expired (block %d thread %d bar %u parity %u)\n", blockIdx.x,
                   threadIdx.x, smem_u32(bar), parity);
            __trap();
        }
    }
}

// ---- TMA ----------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2,
                                            int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2),
        "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tma_load_5d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1,
                                            int c2, int c3, int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2),
        "r"(c3), "r"(c4)
        : "memory");
}

// ---- tcgen05 / TMEM -------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {   // one full warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
                 "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {     // same warp that allocated
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc]; bf16 inputs, fp32 accumulate.  One thread issues.
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// Arrive on an mbarrier once all previously issued tcgen05.mma of this thread have completed.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
// 32 lanes x 32 consecutive fp32 columns: thread i of the warp gets row (lane_base+i), columns [col, col+32).
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Shared-memory matrix descriptor (tcgen05 "SmemDescriptor"): start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46),
// version=1 [46,48), layout type [61,64) (2 = SWIZZLE_128B, 4 = SWIZZLE_64B, 0 = none).
__device__ __forceinline__ uint64_t umma_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                                   uint32_t layout_type) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
    d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(layout_type & 7) << 61;
    return d;
}
// Instruction descriptor for kind::f16 with bf16 A/B and fp32 D.  a_mn / b_mn: 1 = MN-major operand.
__host__ __device__ inline uint32_t umma_idesc_bf16(uint32_t m, uint32_t n, uint32_t a_mn, uint32_t b_mn) {
    uint32_t d = 0;
    d |= 1u << 4;            // D format fp32
    d |= 1u << 7;            // A format bf16
    d |= 1u << 10;           // B format bf16
    d |= (a_mn & 1u) << 15;
    d |= (b_mn & 1u) << 16;
    d |= ((n >> 3) & 0x3Fu) << 17;
    d |= ((m >> 4) & 0x1Fu) << 24;
    return d;
}

// ---- misc ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// Column sums across the 32 lanes of a warp: on return lane l holds sum_over_lanes(v[l]) in v[0] (31 shuffles).
__device__ __forceinline__ float warp_transpose_reduce32(float (&v)[32], int lane) {
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) {
        const bool upper = (lane & s) != 0;
#pragma unroll
        for (int i = 0; i < s; ++i) {
            float send = upper ? v[i] : v[i + s];
            float keep = upper ? v[i + s] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
        }
    }
    return v[0];
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&h);
}

// ---- coalescing epilogue ---------------------------------------------------------------------------------------
// tcgen05.ld 32x32b hands every lane ONE accumulator row (= one pixel): a direct store is 32 lanes x 16 B on 32 different
// 128-byte lines per instruction, per-channel scale / shift cost 16 shared-memory loads per 32 columns, and the fused
// BatchNorm statistics need a 31-shuffle transpose-reduce per quantity.  The epilogue warps therefore pass every
// 32 row x 32 column f32 block through a warp-private 4 KB shared-memory tile (16-byte chunks XOR-swizzled by row:
// conflict-free both ways) and continue in the TRANSPOSED ownership: lane = (sub = lane / 8, part = lane % 8) holds
// columns 4 part .. 4 part + 3 of rows 4 i + sub, i = 0..7.  A store instruction then covers 4 pixels x 128 B (f32) or
// 4 x 64 B (bf16) of whole sectors, a lane needs ONE float4 of scale / shift per block, and the column sums are 8 adds
// in registers plus a 6-shuffle butterfly over the 4 lanes that share a column set.
constexpr int EPI_STAGE_BYTES = 32 * 32 * 4;          // per epilogue warp

struct EpiRows {
    int pix[8];          // pixel index (NOT yet multiplied by the channel stride) of tile row 4 i + sub, -1 = row not stored
};
// pix_own: this lane's own pixel index (row = lane of the warp's 32 rows) or -1
__device__ __forceinline__ EpiRows epi_rows(int pix_own, int lane) {
    EpiRows r;
#pragma unroll
    for (int i = 0; i < 8; ++i) r.pix[i] = __shfl_sync(0xffffffffu, pix_own, 4 * i + (lane >> 3));
    return r;
}

struct EpiOut {
    void* y;                         // output tensor, channels-last
    long long y_cstride;             // elements per pixel
    int y_bf16, relu, accumulate;
    const __nv_bfloat16* relu_mask;  // fused ReLU backward: zero where the activation is not positive (nullable)
    long long mask_cstride;
    const void* addend;              // optional: y = act(...) + addend[pixel, channel] (f32 or bf16 tensor indexed like y)
    long long addend_cstride;
    int addend_bf16;
};

// Step 1: the lane's raw accumulator row v of a 32-column block -> x[i] = columns 4 part .. 4 part + 3 of row 4 i + sub.
// All 32 lanes must call (the tile is warp-collective).
__device__ __forceinline__ void epi_transpose(float* stage, const uint32_t (&v)[32], int lane, float4 (&x)[8]) {
    const int sub = lane >> 3, part = lane & 7;
    {
        uint4* s = reinterpret_cast<uint4*>(stage) + lane * 8;
#pragma unroll
        for (int c = 0; c < 8; ++c) s[c ^ (lane & 7)] = make_uint4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
    }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int row = 4 * i + sub;
        x[i] = reinterpret_cast<const float4*>(stage)[row * 8 + (part ^ (row & 7))];
    }
    __syncwarp();                                        // the tile may be overwritten by the next block from here on
}
// Fused ReLU backward: zero the gradient where the bf16 activation of the layer below is not positive.
__device__ __forceinline__ void epi_relu_mask(float4 (&x)[8], const EpiRows& rows, const EpiOut& o, int ch) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
        if (rows.pix[i] >= 0) {
            const uint2 m = __ldg(reinterpret_cast<const uint2*>(o.relu_mask + (long long)rows.pix[i] * o.mask_cstride + ch));
            // bf16 > 0  <=>  sign bit clear and not zero
            const uint32_t m0 = m.x & 0xffffu, m1 = m.x >> 16, m2 = m.y & 0xffffu, m3 = m.y >> 16;
            if (!(m0 != 0 && m0 < 0x8000u)) x[i].x = 0.0f;
            if (!(m1 != 0 && m1 < 0x8000u)) x[i].y = 0.0f;
            if (!(m2 != 0 && m2 < 0x8000u)) x[i].z = 0.0f;
            if (!(m3 != 0 && m3 < 0x8000u)) x[i].w = 0.0f;
        }
}
// Column sums / sums of squares of the stored rows, accumulated into the caller's registers (columns 4 part ..).
__device__ __forceinline__ void epi_colsum(const float4 (&x)[8], const EpiRows& rows, float (&s)[4], float (&q)[4]) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
        if (rows.pix[i] >= 0) {
            s[0] += x[i].x; s[1] += x[i].y; s[2] += x[i].z; s[3] += x[i].w;
            q[0] = fmaf(x[i].x, x[i].x, q[0]); q[1] = fmaf(x[i].y, x[i].y, q[1]);
            q[2] = fmaf(x[i].z, x[i].z, q[2]); q[3] = fmaf(x[i].w, x[i].w, q[3]);
        }
}
// Butterfly over the 4 lanes of a column set (lane bits 3, 4): afterwards lane bit 4 selects sums / squares and lane bit 3
// the column pair, every lane holding 2 finished values, which it adds to sum_dst / sq_dst (nullable) at channel ch + ...
__device__ __forceinline__ void epi_colsum_flush(const float (&s)[4], const float (&q)[4], int lane, float* sum_dst, float* sq_dst,
                                                 int ch) {
    const bool up16 = (lane & 16) != 0, up8 = (lane & 8) != 0;
    float k[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float send = up16 ? s[j] : q[j];
        k[j] = (up16 ? q[j] : s[j]) + __shfl_xor_sync(0xffffffffu, send, 16);
    }
    float f[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const float send = up8 ? k[j] : k[j + 2];
        f[j] = (up8 ? k[j + 2] : k[j]) + __shfl_xor_sync(0xffffffffu, send, 8);
    }
    float* dst = up16 ? sq_dst : sum_dst;
    if (dst != nullptr) {
        atomicAdd(dst + ch + (up8 ? 2 : 0), f[0]);
        atomicAdd(dst + ch + (up8 ? 2 : 0) + 1, f[1]);
    }
}
// Raw loads of the addend rows (EpiOut.addend), kept apart from their use so that all 8 are in flight at once - and so that a
// kernel can issue them BEFORE it waits for the accumulators (with the bf16 unpacking next to its load the compiler
// serialised the 8 loads: lateral dgrad 177 us with an f32 addend, 418 us with a bf16 one).
__device__ __forceinline__ void epi_addend_load(uint4 (&ad)[8], const EpiRows& rows, int ch, const EpiOut& o) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        ad[i] = make_uint4(0u, 0u, 0u, 0u);
        if (o.addend != nullptr && rows.pix[i] >= 0) {
            if (o.addend_bf16) {
                const uint2 u = *reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(o.addend) +
                                                                (long long)rows.pix[i] * o.addend_cstride + ch);
                ad[i].x = u.x; ad[i].y = u.y;
            } else {
                ad[i] = *reinterpret_cast<const uint4*>(reinterpret_cast<const float*>(o.addend) +
                                                        (long long)rows.pix[i] * o.addend_cstride + ch);
            }
        }
    }
}
// Per-channel affine / ReLU, the addend (ad: epi_addend_load) and the store (bf16, f32 or f32 read-modify-write).  scale / shift:
// nullable, indexed by absolute channel (shared or global memory, 16-byte aligned at ch).
__device__ __forceinline__ void epi_store(float4 (&x)[8], const EpiRows& rows, int ch, const float* scale, const float* shift,
                                          const EpiOut& o, const uint4 (&ad)[8]) {
    if (scale != nullptr || shift != nullptr) {
        const float4 sc = scale ? *reinterpret_cast<const float4*>(scale + ch) : make_float4(1.f, 1.f, 1.f, 1.f);
        const float4 sh = shift ? *reinterpret_cast<const float4*>(shift + ch) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            x[i].x = fmaf(x[i].x, sc.x, sh.x); x[i].y = fmaf(x[i].y, sc.y, sh.y);
            x[i].z = fmaf(x[i].z, sc.z, sh.z); x[i].w = fmaf(x[i].w, sc.w, sh.w);
        }
    }
    if (o.relu) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            x[i].x = fmaxf(x[i].x, 0.f); x[i].y = fmaxf(x[i].y, 0.f); x[i].z = fmaxf(x[i].z, 0.f); x[i].w = fmaxf(x[i].w, 0.f);
        }
    }
    if (o.addend != nullptr) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (o.addend_bf16) {
                x[i].x += __uint_as_float(ad[i].x << 16); x[i].y += __uint_as_float(ad[i].x & 0xffff0000u);
                x[i].z += __uint_as_float(ad[i].y << 16); x[i].w += __uint_as_float(ad[i].y & 0xffff0000u);
            } else {
                x[i].x += __uint_as_float(ad[i].x); x[i].y += __uint_as_float(ad[i].y);
                x[i].z += __uint_as_float(ad[i].z); x[i].w += __uint_as_float(ad[i].w);
            }
        }
    }
    if (o.y_bf16) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            if (rows.pix[i] >= 0) {
                uint2 u;
                u.x = pack_bf16x2(x[i].x, x[i].y);
                u.y = pack_bf16x2(x[i].z, x[i].w);
                *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(o.y) + (long long)rows.pix[i] * o.y_cstride + ch) = u;
            }
    } else if (o.accumulate) {
        float4 old[8];
#pragma unroll
        for (int i = 0; i < 8; ++i)                      // all reads of the read-modify-write in flight at once
            old[i] = rows.pix[i] >= 0
                         ? *reinterpret_cast<const float4*>(reinterpret_cast<float*>(o.y) + (long long)rows.pix[i] * o.y_cstride + ch)
                         : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int i = 0; i < 8; ++i)
            if (rows.pix[i] >= 0)
                *reinterpret_cast<float4*>(reinterpret_cast<float*>(o.y) + (long long)rows.pix[i] * o.y_cstride + ch) =
                    make_float4(x[i].x + old[i].x, x[i].y + old[i].y, x[i].z + old[i].z, x[i].w + old[i].w);
    } else {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            if (rows.pix[i] >= 0)
                *reinterpret_cast<float4*>(reinterpret_cast<float*>(o.y) + (long long)rows.pix[i] * o.y_cstride + ch) = x[i];
    }
}
// The whole block: v = the lane's raw accumulator row, col = first output channel of the block.  sum_dst / sq_dst: nullable
// atomicAdd targets indexed by absolute channel: column sums (and sums of squares) of the RAW (masked) accumulators over
// the stored rows.
__device__ __forceinline__ void epi_block(float* stage, const uint32_t (&v)[32], const EpiRows& rows, int lane, int col,
                                          const float* scale, const float* shift, const EpiOut& o, float* sum_dst,
                                          float* sq_dst) {
    float4 x[8];
    uint4 ad[8];
    const int ch = col + 4 * (lane & 7);
    epi_transpose(stage, v, lane, x);
    epi_addend_load(ad, rows, ch, o);
    if (o.relu_mask != nullptr) epi_relu_mask(x, rows, o, ch);
    if (sum_dst != nullptr) {
        float s[4] = {0.f, 0.f, 0.f, 0.f}, q[4] = {0.f, 0.f, 0.f, 0.f};
        epi_colsum(x, rows, s, q);
        epi_colsum_flush(s, q, lane, sum_dst, sq_dst, ch);
    }
    epi_store(x, rows, ch, scale, shift, o, ad);
}
#endif  // __CUDACC__
