// Shared device/host helpers for the pggan_b200 sm_100a kernels.
// PTX wrappers for mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (UMMA) and TMEM.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/pggan_b200.h"

namespace pg {

// ----------------------------------------------------------------------------- errors
void set_error(const char* fmt, ...);
#define PG_CHECK(cond, code, ...)                                   \
    do {                                                            \
        if (!(cond)) {                                              \
            pg::set_error(__VA_ARGS__);                             \
            return (code);                                          \
        }                                                           \
    } while (0)
#define PG_CUDA(expr)                                                              \
    do {                                                                           \
        cudaError_t _e = (expr);                                                   \
        if (_e != cudaSuccess) {                                                   \
            pg::set_error("%s:%d CUDA error %s: %s", __FILE__, __LINE__,           \
                          cudaGetErrorName(_e), cudaGetErrorString(_e));           \
            return PG_ERR_CUDA;                                                    \
        }                                                                          \
    } while (0)
#define PG_LAUNCH_CHECK() PG_CUDA(cudaPeekAtLastError())
// Programmatic dependent launch (griddepcontrol), OPT-IN (PGGAN_PDL=1|2, see pdl_mode() in common.cu for the measurements
// that keep it off by default): a kernel launched with the attribute may be SCHEDULED while its predecessor in the stream
// is still running — its CTAs take an SM as soon as one is free, run their prologue (barrier init, TMEM allocation,
// descriptor prefetch) and block in pdl_wait() until the predecessor has completed and its writes are visible.  Every
// kernel calls PG_PDL_ENTER() before its first global-memory access: wait for the predecessor, then allow the successor
// to be scheduled.  Without the launch attribute both instructions are no-ops.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#define PG_PDL_ENTER()     \
    do {                   \
        pg::pdl_wait();    \
        pg::pdl_trigger(); \
    } while (0)

int pdl_mode();            // common.cu: env PGGAN_PDL — 0 off, 1 every kernel, 2 the tensor-core kernels only

template <typename... KArgs, typename... Args>
inline cudaError_t launch_attr(bool pdl, void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                               Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}
// bandwidth / glue kernels
template <typename... KArgs, typename... Args>
inline cudaError_t launch(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
    return launch_attr(pdl_mode() == 1, kern, grid, block, smem, stream, static_cast<Args&&>(args)...);
}
// tcgen05 kernels (one persistent CTA per SM with a prologue worth overlapping: barrier init, TMEM allocation)
template <typename... KArgs, typename... Args>
inline cudaError_t launch_tc(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
    return launch_attr(pdl_mode() >= 1, kern, grid, block, smem, stream, static_cast<Args&&>(args)...);
}

// opt a kernel into > 48 KB of dynamic shared memory: the attribute is per (function, DEVICE), so remember it per device
#define PG_SET_SMEM_ONCE(kern, bytes)                                                                         \
    do {                                                                                                      \
        static bool done_[64] = {};                                                                           \
        int dev_ = 0;                                                                                         \
        if (cudaGetDevice(&dev_) != cudaSuccess) dev_ = -1;                                                   \
        if (dev_ < 0 || dev_ >= 64 || !done_[dev_]) {                                                         \
            PG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes)));   \
            if (dev_ >= 0 && dev_ < 64) done_[dev_] = true;                                                   \
        }                                                                                                     \
    } while (0)

int num_sms();

// tensor-map encode through the driver entry point (no link-time libcuda dependency)
int encode_tmap(CUtensorMap* map, int dtype, int rank, const void* base, const uint64_t* dims,
                const uint64_t* strides_bytes, const uint32_t* box, int swizzle_bytes);

// 3xTF32 operand split (elementwise.cu); see pggan_split3 in the header
int split3_launch(const float* in, float* out, long long m, int c, int along_channels, int lo_slot, cudaStream_t s);

inline int dtype_size(int dt) { return (dt == PG_BF16 || dt == PG_F16) ? 2 : 4; }
__host__ __device__ inline bool is16(int dt) { return dt == PG_BF16 || dt == PG_F16; }

// ------------------------------------------------------------------------------ dtypes
template <typename T> struct DT;
template <> struct DT<__nv_bfloat16> {
    static constexpr int kEnum = PG_BF16;
    static __device__ __forceinline__ float to_f(__nv_bfloat16 v) { return __bfloat162float(v); }
    static __device__ __forceinline__ __nv_bfloat16 from_f(float v) { return __float2bfloat16_rn(v); }
};
template <> struct DT<__half> {
    static constexpr int kEnum = PG_F16;
    static __device__ __forceinline__ float to_f(__half v) { return __half2float(v); }
    static __device__ __forceinline__ __half from_f(float v) {      // saturating (see pack16)
        uint16_t h;
        asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(h) : "f"(v));
        return __ushort_as_half(h);
    }
};
template <> struct DT<float> {
    static constexpr int kEnum = PG_F32;
    static __device__ __forceinline__ float to_f(float v) { return v; }
    // fp32 feature maps are stored UNROUNDED: the fp32 mode runs every GEMM as 3xTF32
    // (hi*hi + hi*lo + lo*hi, see pggan_split3) so the tensor core's truncation to tf32 is compensated.
    static __device__ __forceinline__ float from_f(float v) { return v; }
};

// 16-byte vector of T
template <typename T> struct Vec16 { static constexpr int kN = 16 / sizeof(T); };

// value of a pg_coef: c0 + c1 * (*alpha) — alpha lives in device memory so a captured graph follows the fade-in
__device__ __forceinline__ float coef_val(const pg_coef& c) {
    return c.alpha != nullptr ? fmaf(c.c1, __ldg(c.alpha), c.c0) : c.c0;
}

__device__ __forceinline__ float lrelu_f(float v, float slope) { return v > 0.f ? v : v * slope; }

// The tensor-core kernels are instantiated once for 16-bit operands (T = __nv_bfloat16) and told at run time whether
// the bits are bf16 or IEEE half (both are kind::f16 MMAs at the same rate): only three things depend on the format.
//   1. the instruction-descriptor operand format (0 = f16, 1 = bf16)          -> fmt16()
//   2. packing two fp32 results into one 32-bit word                          -> pack16()
//   3. the LeakyReLU-mask test "x > 0" on a stored activation: both formats are sign-magnitude, so it is the
//      signed-integer test on the raw 16 bits                                 -> pos16()
__host__ __device__ inline uint32_t fmt16(int dt) { return dt == PG_F16 ? 0u : 1u; }
// fp32 -> IEEE half conversions SATURATE to +-65504 instead of producing inf (one F2FP.SATFINITE instruction): fp16
// activations of the mixed-precision mode cannot overflow into inf/NaN; bf16 has fp32's exponent range and needs none.
__device__ __forceinline__ uint32_t pack16(float a, float b, bool f16) {
    if (f16) {
        uint32_t r;
        asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
        return r;
    }
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ bool pos16(uint16_t bits) { return static_cast<int16_t>(bits) > 0; }
__device__ __forceinline__ uint16_t cvt16(float v, bool f16) {
    if (f16) {
        uint16_t h;
        asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(h) : "f"(v));
        return h;
    }
    return __bfloat16_as_ushort(__float2bfloat16_rn(v));
}
__device__ __forceinline__ float f16bits_to_f(uint16_t bits, bool f16) {
    if (f16) return __half2float(__ushort_as_half(bits));
    return __uint_as_float(static_cast<uint32_t>(bits) << 16);
}
// "stored activation > 0" for a vector element of type T (16-bit: raw sign-magnitude test; fp32: the comparison)
template <typename T> __device__ __forceinline__ bool act_pos(const T& v) {
    if constexpr (sizeof(T) == 2) return pos16(*reinterpret_cast<const uint16_t*>(&v));
    else return v > 0.f;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// block-wide sum; every thread gets the result. `red` is >= 32 floats of shared memory.
__device__ __forceinline__ float block_sum(float v, float* red) {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) red[wid] = v;
    __syncthreads();
    const int nw = (blockDim.x + 31) >> 5;
    float t = (lane < nw) ? red[lane] : 0.f;
    t = warp_sum(t);
    return t;
}

// --------------------------------------------------------------------------- PTX: misc
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n\t.reg .pred P;\n\t"
        "elect.sync _|P, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t}"
        : "=r"(pred));
    return pred != 0;
}

// ----------------------------------------------------------------------- PTX: mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
// Wait on an mbarrier phase (try_wait suspends in hardware between polls).  A protocol bug would
// otherwise hang the GPU, so after ~4 s of wall time the kernel traps instead (PG_NO_WATCHDOG removes it).
__device__ __forceinline__ uint64_t global_ns() {
    uint64_t t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar_addr, uint32_t parity) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred P;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t}"
        : "=r"(done)
        : "r"(bar_addr), "r"(parity)
        : "memory");
    return done != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t a = smem_u32(bar);
#pragma unroll 1
    for (int i = 0; i < 1024; ++i)
        if (mbar_try_wait(a, parity)) return;
#ifndef PG_NO_WATCHDOG
    const uint64_t t0 = global_ns();
#endif
#pragma unroll 1
    while (!mbar_try_wait(a, parity)) {
#ifndef PG_NO_WATCHDOG
        if (global_ns() - t0 > 4000000000ull) __trap();
#endif
    }
}

// ---------------------------------------------------------------------------- PTX: TMA
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
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1,
                                            int c2, int c3) {
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

// ------------------------------------------------------------------------ PTX: tcgen05
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

template <int kCols>
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst) {   // whole warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
                 "n"(kCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int kCols>
__device__ __forceinline__ void tmem_dealloc(uint32_t addr) {      // whole warp (the allocating one)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "n"(kCols) : "memory");
}
// tcgen05.commit: arrive on `bar` once all previously issued MMAs of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]; kind::f16 covers bf16/fp16 operands, kind::tf32 fp32 operands.
template <bool kTF32>
__device__ __forceinline__ void umma_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                        uint32_t accumulate) {
    if constexpr (kTF32) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
            ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
            : "memory");
    } else {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
            ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
            : "memory");
    }
}

// Shared-memory matrix descriptor (sm_100 "version 1"), SWIZZLE_128B layouts only.
//   K-major  operand: rows of 128 B (one swizzle span of K), 8-row groups SBO bytes apart.
//   MN-major operand: rows of 128 B (64 bf16 / 32 tf32 along MN), 8 K-rows per group SBO bytes apart,
//                     successive 128 B MN spans LBO bytes apart.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                                   uint32_t layout_type = 2 /* SWIZZLE_128B */) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
    d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= 1ull << 46;   // descriptor version (Blackwell)
    d |= static_cast<uint64_t>(layout_type) << 61;   // 2 = SWIZZLE_128B, 1 = SWIZZLE_128B_BASE32B
    return d;
}
// Instruction descriptor for kind::f16 / kind::tf32, fp32 accumulate, dense.
//   a_fmt/b_fmt: 1 = bf16, 2 = tf32; major: 0 = K-major, 1 = MN-major.
//   The A and B formats are separate fields (bits 7-9 / 10-12), but kind::f16 wants them EQUAL: a bf16 x f16 descriptor
//   raises an illegal-instruction fault on B200 (measured with tools/probe_mixed.py) — callers pass one format.
__host__ __device__ constexpr uint32_t make_idesc2(int a_fmt, int b_fmt, int a_major, int b_major, int m, int n) {
    return (1u << 4) | (uint32_t(a_fmt) << 7) | (uint32_t(b_fmt) << 10) | (uint32_t(a_major) << 15) |
           (uint32_t(b_major) << 16) | (uint32_t(n >> 3) << 17) | (uint32_t(m >> 4) << 24);
}
__host__ __device__ constexpr uint32_t make_idesc(int fmt, int a_major, int b_major, int m, int n) {
    return make_idesc2(fmt, fmt, a_major, b_major, m, n);
}


// TMEM -> registers: 32 lanes x 32 consecutive fp32 columns (this warp's lane quarter)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
}
// 32 lanes x 16 consecutive fp32 columns
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

}  // namespace pg
