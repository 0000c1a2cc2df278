// Library-wide state: error string, launch counters, SM count, tensor-map encoding.
#include "common.cuh"

#include <atomic>
#include <mutex>
#include <stdarg.h>
#include <string.h>

namespace pg {

static thread_local char g_err[512] = "";
std::atomic<long long> g_tc_launches{0}, g_simt_launches{0}, g_other_launches{0};
int g_conv_backend = PG_BACKEND_AUTO;

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int num_sms() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, []() {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    });
    return fn;
}

int encode_tmap(CUtensorMap* map, int dtype, int rank, const void* base, const uint64_t* dims,
                const uint64_t* strides_bytes, const uint32_t* box, int swizzle_bytes) {
    EncodeTiledFn fn = get_encode_fn();
    PG_CHECK(fn != nullptr, PG_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUtensorMapSwizzle sw = swizzle_bytes == 12832 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B
                            : swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                            : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                            : swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B
                                                  : CU_TENSOR_MAP_SWIZZLE_NONE;
    CUtensorMapDataType dt = dtype == PG_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
                             : dtype == PG_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16
                                               : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
    CUresult r = fn(map, dt, (cuuint32_t)rank, const_cast<void*>(base), (const cuuint64_t*)dims,
                    (const cuuint64_t*)strides_bytes, (const cuuint32_t*)box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (CUresult %d): rank %d dims [%llu,%llu,%llu,%llu,..] box [%u,%u,%u,%u,..]",
                  (int)r, rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
                  (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)(rank > 3 ? dims[3] : 0), box[0],
                  rank > 1 ? box[1] : 0, rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0);
        return PG_ERR_CUDA;
    }
    return PG_OK;
}

}  // namespace pg

extern "C" int pggan_version(void) { return 100; }
extern "C" const char* pggan_last_error(void) { return pg::g_err; }
extern "C" int pggan_set_conv_backend(int backend) {
    if (backend != PG_BACKEND_AUTO && backend != PG_BACKEND_SIMT) {
        pg::set_error("pggan_set_conv_backend: unknown backend %d", backend);
        return PG_ERR_ARG;
    }
    pg::g_conv_backend = backend;
    return PG_OK;
}
extern "C" int pggan_launch_counts(int64_t* tc, int64_t* simt, int64_t* other) {
    if (tc) *tc = pg::g_tc_launches.load();
    if (simt) *simt = pg::g_simt_launches.load();
    if (other) *other = pg::g_other_launches.load();
    return PG_OK;
}

namespace pg {
int pdl_mode() {
    // programmatic dependent launch of this library's kernels: PGGAN_PDL = 0 (default) off, 1 every kernel, 2 the
    // tensor-core kernels only.  Measured on whole steps (tools/gpu_pdl_modes.sh, one box, C5 at full clocks): 20.67 ms off,
    // 20.90 ms mode 2, 21.37 ms mode 1; C4 27.1-27.6 ms in every mode; C1 3.20 / 3.18 / 3.13 ms — the 0.5-0.7 us per
    // launch it saves in back-to-back micro-benchmarks does not survive in the step (pre-launched CTAs take their SMs in
    // the order SMs happen to free up, which unbalances the bandwidth kernels), so it stays off.
    static const int mode = [] {
        const char* e = getenv("PGGAN_PDL");
        if (e == nullptr || e[0] == '\0') return 0;
        return (e[0] >= '0' && e[0] <= '2') ? e[0] - '0' : 0;
    }();
    return mode;
}
}  // namespace pg
