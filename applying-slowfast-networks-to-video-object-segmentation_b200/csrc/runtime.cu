// Host-side runtime shared by all entry points: error string, device check, SM count, TMA descriptor encode.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

static thread_local char g_err[1024] = "";

void sfvos_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

static thread_local const char* g_kernel = "";
void sfvos_set_kernel(const char* name) { g_kernel = name; }

extern "C" const char* sfvos_last_error(void) { return g_err; }
extern "C" const char* sfvos_last_kernel(void) { return g_kernel; }
extern "C" int sfvos_version(void) { return 100; }

extern "C" int sfvos_device_check(void) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) {
        sfvos_set_error("no CUDA device: %s (libsfvos has no CPU fallback)", cudaGetErrorString(e));
        return SFVOS_ERR_UNSUPPORTED;
    }
    int major = 0, minor = 0;
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
    if (major != 10) {
        sfvos_set_error("device %d is sm_%d%d; libsfvos is built for sm_100a only", dev, major, minor);
        return SFVOS_ERR_UNSUPPORTED;
    }
    return SFVOS_OK;
}

int sfvos_num_sms() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int n = 0;
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        cached[dev] = n > 0 ? n : 148;
    }
    return cached[dev];
}

// cuTensorMapEncodeTiled is a driver-API symbol; fetch it through the runtime so the library links without -lcuda
// (the build container has no libcuda.so.1).
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (fn) return fn;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
        return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(p);
    return fn;
}

int sfvos_make_tmap(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                    const uint64_t* strides_bytes, const uint32_t* box, int swizzle_bytes) {
    EncodeTiledFn fn = get_encode_fn();
    SF_CHECK(fn != nullptr, "cuTensorMapEncodeTiled not available from the driver");
    SF_CHECK((reinterpret_cast<uintptr_t>(base) & 15) == 0, "TMA base address must be 16-byte aligned");
    cuuint64_t gdim[5];
    cuuint64_t gstr[4];
    cuuint32_t bdim[5];
    cuuint32_t estr[5];
    for (int i = 0; i < rank; ++i) {
        gdim[i] = dims[i];
        bdim[i] = box[i];
        estr[i] = 1;
        SF_CHECK(box[i] >= 1 && box[i] <= 256, "TMA box dim %d = %u out of range", i, box[i]);
    }
    for (int i = 0; i + 1 < rank; ++i) {
        gstr[i] = strides_bytes[i];
        SF_CHECK((strides_bytes[i] & 15) == 0, "TMA stride %d = %llu not a multiple of 16 bytes", i,
                 (unsigned long long)strides_bytes[i]);
    }
    CUtensorMapSwizzle sw = swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                            : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                            : swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B
                                                  : CU_TENSOR_MAP_SWIZZLE_NONE;
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr,
                    bdim, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        sfvos_set_error("cuTensorMapEncodeTiled failed with CUresult %d (rank %d dims %llu,%llu,.. box %u,%u,..)",
                        (int)r, rank, (unsigned long long)dims[0], (unsigned long long)dims[1], box[0], box[1]);
        return SFVOS_ERR_CUDA;
    }
    return SFVOS_OK;
}

// Can a tensor map describe OVERLAPPING windows, i.e. a batch stride smaller than the extent of the dimensions below it
// (window b = frames [b, b+T) of one sequence buffer: stride_B = one frame)?  The TMA unit only forms base + sum(coord * stride),
// but the driver validates the descriptor, so ask it once: encode a 5-D map {C, W, H, T, B} with stride_B = stride_T on the
// caller's (16-byte aligned, never dereferenced) device pointer.
extern "C" int sfvos_tma_overlap_supported(const void* dev_ptr) {
    static int cached = -1;
    if (cached >= 0) return cached;
    if (sfvos_device_check() != SFVOS_OK || dev_ptr == nullptr) return 0;
    CUtensorMap m;
    const uint64_t C = 64, W = 16, H = 8, T = 4, B = 3;
    const uint64_t dims[5] = {C, W, H, T, B};
    const uint64_t strides[4] = {C * 2, W * C * 2, H * W * C * 2, H * W * C * 2};      // stride_B == stride_T: windows overlap
    const uint32_t box[5] = {64, 16, 8, 1, 1};
    cached = sfvos_make_tmap(&m, dev_ptr, 5, dims, strides, box, 128) == SFVOS_OK ? 1 : 0;
    return cached;
}

// ----------------------------------------------------------------------------------------------------------------------
// Measurement probes (tools/bench_l2.py): the denominators of the kernels that are bound by the L2 rather than by HBM.
//   kind 0: streaming 16-byte loads (grid-stride, summed)  -> L2 -> SM read bandwidth when the buffer fits the L2
//   kind 1: streaming red.global.add.v4.f32                 -> vector-reduction throughput of the L2 (ROIAlign backward)
// Every pass touches each 16-byte vector of the buffer once; `iters` passes per launch.
// ----------------------------------------------------------------------------------------------------------------------
namespace {
__global__ void __launch_bounds__(256) probe_l2_kernel(int kind, float4* buf, long long nvec, int iters, float* sink) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    float acc = 0.f;
    for (int it = 0; it < iters; ++it)
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride) {
            if (kind == 0) {
                const float4 v = __ldcg(buf + i);
                acc += v.x + v.y + v.z + v.w;
            } else {
                atomicAdd(buf + i, make_float4(1.f, 1.f, 1.f, 1.f));
            }
        }
    if (kind == 0 && acc == 123.456f) *sink = acc;          // keeps the loads alive
}
}  // namespace

extern "C" int sfvos_probe_l2(int32_t kind, void* buf, int64_t nbytes, int32_t iters, float* sink, sfvos_stream stream) {
    SF_CHECK(kind == 0 || kind == 1, "probe_l2: kind must be 0 (loads) or 1 (vector reductions)");
    SF_CHECK(buf != nullptr && sink != nullptr && nbytes >= 16 && (reinterpret_cast<uintptr_t>(buf) & 15) == 0, "probe_l2: bad buffer");
    int rc = sfvos_device_check();
    if (rc) return rc;
    const long long nvec = nbytes / 16;
    long long blocks = (nvec + 255) / 256;
    const long long cap = (long long)sfvos_num_sms() * 16;
    if (blocks > cap) blocks = cap;
    probe_l2_kernel<<<(int)blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(kind, reinterpret_cast<float4*>(buf), nvec, iters, sink);
    SF_LAUNCH_CHECK();
    return SFVOS_OK;
}
