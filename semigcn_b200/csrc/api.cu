// Library-wide plumbing: version, thread-local error string, cached device attributes.
#include "common.cuh"
#include "umma.cuh"
#include <string.h>
#include <atomic>

namespace sgb {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int num_sms() {
    static thread_local int cached_dev = -1;
    static thread_local int cached_sms = 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (dev != cached_dev) {
        int v = 0;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
        cached_dev = dev;
        cached_sms = v;
    }
    return cached_sms;
}

// Opt a kernel into > 48 KB of dynamic shared memory.  The attribute is per (function, device): `done_mask` (one static
// per kernel at the call site) remembers the devices it has been set on, so a process that moves from cuda:0 to cuda:1
// sets it again there, and the steady state costs one cudaGetDevice.
int smem_optin(const void* func, int bytes, std::atomic<uint64_t>* done_mask) {
    int dev = 0;
    SGB_CUDA(cudaGetDevice(&dev));
    const uint64_t bit = dev < 64 ? (uint64_t)1 << dev : 0;
    if (bit && (done_mask->load(std::memory_order_relaxed) & bit)) return SGB_OK;
    SGB_CUDA(cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    if (bit) done_mask->fetch_or(bit, std::memory_order_relaxed);
    return SGB_OK;
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time dependency on libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int make_tmap_2d(CUtensorMap* out, const float* base, int64_t rows, int64_t cols, int64_t ld, int box_cols, int box_rows, bool swizzle128) {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) != cudaSuccess || qr != cudaDriverEntryPointSuccess || !p) {
            set_error("cuTensorMapEncodeTiled is not available from this driver");
            return SGB_ECUDA;
        }
        fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(float)};
    const cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (%d): rows=%lld cols=%lld ld=%lld box=%dx%d", (int)r, (long long)rows, (long long)cols, (long long)ld,
                  box_cols, box_rows);
        return SGB_ECUDA;
    }
    return SGB_OK;
}

}  // namespace sgb

extern "C" int sgb_version(void) { return SGB_VERSION; }
extern "C" const char* sgb_last_error(void) { return sgb::g_err; }
extern "C" int sgb_num_sms(void) { return sgb::num_sms(); }
