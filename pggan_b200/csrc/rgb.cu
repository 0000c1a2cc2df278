// 1x1 RGB convolutions and RGB-space fades: pure bandwidth kernels (N=3 / K=3 is not tensor-core
// shaped).  Feature maps NHWC (bf16 / fp32), images fp32 NCHW as the reference passes them.
// Reference: toRGBBlock / fromRGBBlock (lib/blocks.py:153-170, 271-292), G's RGB fade
// (pggan/nets.py:155-156), the real-image fade (pggan/model.py:217-221).
#include "common.cuh"

#include <atomic>

namespace pg {
extern std::atomic<long long> g_other_launches;

constexpr int kMaxImgC = 4;

template <typename T> struct RV {
    static constexpr int N = 16 / sizeof(T);
    uint4 raw;
    __device__ __forceinline__ float get(int i) const { return DT<T>::to_f(reinterpret_cast<const T*>(&raw)[i]); }
    __device__ __forceinline__ void set(int i, float v) { reinterpret_cast<T*>(&raw)[i] = DT<T>::from_f(v); }
};

__device__ __forceinline__ void load_pixel(const float* __restrict__ img, int cimg, int pool, long long b, int h, int w,
                                           int H, int W, float (&p)[kMaxImgC]) {
    // (H, W) are the FEATURE-map dims; the image is [B,cimg,H,W] or, if pool, [B,cimg,2H,2W]
    if (!pool) {
        const long long plane = static_cast<long long>(H) * W;
#pragma unroll
        for (int j = 0; j < kMaxImgC; ++j)
            p[j] = j < cimg ? __ldg(img + (b * cimg + j) * plane + static_cast<long long>(h) * W + w) : 0.f;
    } else {
        const int Wi = 2 * W;
        const long long plane = 4LL * H * W;
#pragma unroll
        for (int j = 0; j < kMaxImgC; ++j) {
            if (j < cimg) {
                const float* q = img + (b * cimg + j) * plane + (2LL * h) * Wi + 2 * w;
                const float2 r0 = __ldg(reinterpret_cast<const float2*>(q));
                const float2 r1 = __ldg(reinterpret_cast<const float2*>(q + Wi));
                p[j] = (((r0.x + r0.y) + r1.x) + r1.y) * 0.25f;      // same order as avg_pool2d's window sum
            } else {
                p[j] = 0.f;
            }
        }
    }
}

// ------------------------------------------------------------------------------ expand (cimg -> C)
template <typename T>
__global__ void rgb_expand_kernel(const float* __restrict__ img, const float* __restrict__ w,
                                  const float* __restrict__ bias, T* __restrict__ out, int B, int H, int W, int C,
                                  int cimg, int pool, float scale, float slope) {
    constexpr int VN = RV<T>::N;
    extern __shared__ float sw[];          // [C][cimg] weights then [C] bias
    float* sb = sw + C * cimg;
    for (int i = threadIdx.x; i < C * cimg; i += blockDim.x) sw[i] = w[i];
    for (int i = threadIdx.x; i < C; i += blockDim.x) sb[i] = bias ? bias[i] : 0.f;
    __syncthreads();
    const int CV = C / VN;
    const long long total = static_cast<long long>(B) * H * W * CV;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int cv = static_cast<int>(i % CV);
        const long long pix = i / CV;
        const int ww = static_cast<int>(pix % W);
        const int hh = static_cast<int>((pix / W) % H);
        const long long b = pix / (static_cast<long long>(W) * H);
        float p[kMaxImgC];
        load_pixel(img, cimg, pool, b, hh, ww, H, W, p);
        RV<T> o;
#pragma unroll
        for (int e = 0; e < VN; ++e) {
            const int c = cv * VN + e;
            float acc = 0.f;
            for (int j = 0; j < cimg; ++j) acc += sw[c * cimg + j] * p[j];
            o.set(e, lrelu_f((acc + sb[c]) * scale, slope));
        }
        *reinterpret_cast<uint4*>(out + pix * C + cv * VN) = o.raw;
    }
}

// ------------------------------------------------------------------------------ reduce (C -> cimg)
// LP lanes cooperate on one pixel (LP = power of two <= 32); a warp covers 32/LP pixels per step.
template <typename T>
__global__ void rgb_reduce_kernel(const T* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                                  float* __restrict__ out, int B, int H, int W, int C, int cimg, int up, float scale,
                                  float up_scale, int LP) {
    constexpr int VN = RV<T>::N;
    extern __shared__ float sw[];          // [cimg][C]
    for (int i = threadIdx.x; i < C * cimg; i += blockDim.x) sw[i] = w[i];
    __syncthreads();
    const int CV = C / VN;
    const int lane = threadIdx.x & 31;
    const int sub = lane % LP;             // lane within the pixel group
    const int grp = lane / LP;             // which pixel of this warp's batch
    const int ppw = 32 / LP;               // pixels per warp step
    const long long P = static_cast<long long>(B) * H * W;
    const long long warp0 = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5;
    const long long nwarps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
    for (long long base = warp0 * ppw; base < P; base += nwarps * ppw) {
        const long long pix = base + grp;
        float acc[kMaxImgC] = {0.f, 0.f, 0.f, 0.f};
        if (pix < P) {
            for (int cv = sub; cv < CV; cv += LP) {
                RV<T> v;
                v.raw = __ldg(reinterpret_cast<const uint4*>(x + pix * C + cv * VN));
#pragma unroll
                for (int e = 0; e < VN; ++e) {
                    const float xv = v.get(e);
#pragma unroll
                    for (int j = 0; j < kMaxImgC; ++j)
                        if (j < cimg) acc[j] += xv * sw[j * C + cv * VN + e];
                }
            }
        }
#pragma unroll
        for (int j = 0; j < kMaxImgC; ++j)
            for (int o = LP >> 1; o > 0; o >>= 1) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], o);
        if (pix < P && sub < cimg) {
            float v = 0.f;
#pragma unroll
            for (int j = 0; j < kMaxImgC; ++j)
                if (j == sub) v = acc[j];
            v = (v + (bias ? __ldg(bias + sub) : 0.f)) * scale;
            const int ww = static_cast<int>(pix % W);
            const int hh = static_cast<int>((pix / W) % H);
            const long long b = pix / (static_cast<long long>(W) * H);
            if (!up) {
                out[((b * cimg + sub) * H + hh) * W + ww] = v;
            } else {
                v *= up_scale;
                float* o = out + ((b * cimg + sub) * 2 * H + 2 * hh) * 2 * W + 2 * ww;
                *reinterpret_cast<float2*>(o) = make_float2(v, v);
                *reinterpret_cast<float2*>(o + 2 * W) = make_float2(v, v);
            }
        }
    }
}

// ------------------------------------------------------------------------------ 1x1 weight gradient
// acc[c][j] += scale * sum_p f[p,c] * img_j(p).  Threads (ty pixel-lane, tx channel-vector lane).
template <typename T>
__global__ void rgb_wgrad_kernel(const T* __restrict__ f, const float* __restrict__ img, float* __restrict__ acc, int B,
                                 int H, int W, int C, int cimg, int pool, float scale, int TX,
                                 long long pix_per_block) {
    constexpr int VN = RV<T>::N;
    extern __shared__ float red[];         // [blockDim][VN*kMaxImgC]
    const int CV = C / VN;
    const int TY = blockDim.x / TX;
    const int tx = threadIdx.x % TX, ty = threadIdx.x / TX;
    const long long P = static_cast<long long>(B) * H * W;
    const long long p0 = blockIdx.x * pix_per_block;
    long long p1 = p0 + pix_per_block;
    if (p1 > P) p1 = P;
    for (int cv0 = 0; cv0 < CV; cv0 += TX) {
        const int cv = cv0 + tx;
        float a[VN][kMaxImgC];
#pragma unroll
        for (int e = 0; e < VN; ++e)
#pragma unroll
            for (int j = 0; j < kMaxImgC; ++j) a[e][j] = 0.f;
        if (cv < CV) {
            for (long long pix = p0 + ty; pix < p1; pix += TY) {
                const int ww = static_cast<int>(pix % W);
                const int hh = static_cast<int>((pix / W) % H);
                const long long b = pix / (static_cast<long long>(W) * H);
                float p[kMaxImgC];
                load_pixel(img, cimg, pool, b, hh, ww, H, W, p);
                RV<T> v;
                v.raw = __ldg(reinterpret_cast<const uint4*>(f + pix * C + cv * VN));
#pragma unroll
                for (int e = 0; e < VN; ++e) {
                    const float fv = v.get(e);
#pragma unroll
                    for (int j = 0; j < kMaxImgC; ++j) a[e][j] += fv * p[j];
                }
            }
        }
        __syncthreads();
#pragma unroll
        for (int e = 0; e < VN; ++e)
#pragma unroll
            for (int j = 0; j < kMaxImgC; ++j) red[(threadIdx.x * VN + e) * kMaxImgC + j] = a[e][j];
        __syncthreads();
        if (ty == 0 && cv < CV) {
#pragma unroll
            for (int e = 0; e < VN; ++e)
                for (int j = 0; j < cimg; ++j) {
                    float t = 0.f;
                    for (int yy = 0; yy < TY; ++yy) t += red[((yy * TX + tx) * VN + e) * kMaxImgC + j];
                    atomicAdd(acc + (cv * VN + e) * cimg + j, t * scale);
                }
        }
    }
}

// ------------------------------------------------------------------------------ fp32 image kernels
__global__ void image_fade_kernel(const float* __restrict__ x, const float* __restrict__ low, float* __restrict__ out,
                                  long long planes, int H, int W, float alpha) {
    // one thread per 2x2 output block
    const int Hh = H / 2, Wh = W / 2;
    const long long total = planes * Hh * Wh;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int w = static_cast<int>(i % Wh);
        const int h = static_cast<int>((i / Wh) % Hh);
        const long long pl = i / (static_cast<long long>(Wh) * Hh);
        const float* q = x + (pl * H + 2 * h) * W + 2 * w;
        const float2 r0 = *reinterpret_cast<const float2*>(q);
        const float2 r1 = *reinterpret_cast<const float2*>(q + W);
        float lo;
        if (low != nullptr)
            lo = low[(pl * Hh + h) * Wh + w];
        else
            lo = (((r0.x + r0.y) + r1.x) + r1.y) * 0.25f;
        const float s = (1.f - alpha) * lo;
        float* o = out + (pl * H + 2 * h) * W + 2 * w;
        *reinterpret_cast<float2*>(o) = make_float2(s + alpha * r0.x, s + alpha * r0.y);
        *reinterpret_cast<float2*>(o + W) = make_float2(s + alpha * r1.x, s + alpha * r1.y);
    }
}

__global__ void image_pool_sum_kernel(const float* __restrict__ x, float* __restrict__ out, long long planes, int H,
                                      int W, float scale) {
    const int Hh = H / 2, Wh = W / 2;
    const long long total = planes * Hh * Wh;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int w = static_cast<int>(i % Wh);
        const int h = static_cast<int>((i / Wh) % Hh);
        const long long pl = i / (static_cast<long long>(Wh) * Hh);
        const float* q = x + (pl * H + 2 * h) * W + 2 * w;
        const float2 r0 = *reinterpret_cast<const float2*>(q);
        const float2 r1 = *reinterpret_cast<const float2*>(q + W);
        out[i] = (((r0.x + r0.y) + r1.x) + r1.y) * scale;
    }
}

static inline int grid_for(long long items, int threads, int per_sm) {
    long long blocks = (items + threads - 1) / threads;
    const long long cap = static_cast<long long>(num_sms()) * per_sm;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return static_cast<int>(blocks);
}
static inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace pg

using namespace pg;

extern "C" int pggan_rgb_expand(const float* img, const float* w, const float* bias, void* out, int B, int H, int W,
                                int C, int cimg, int pool, float scale, float slope, int dtype, void* stream) {
    PG_CHECK(img && w && out && B > 0 && H > 0 && W > 0 && C > 0, PG_ERR_ARG, "pggan_rgb_expand: bad args");
    PG_CHECK(cimg >= 1 && cimg <= kMaxImgC, PG_ERR_UNSUPPORTED, "pggan_rgb_expand: cimg=%d must be 1..4", cimg);
    PG_CHECK(al16(out) && (!pool || (reinterpret_cast<uintptr_t>(img) & 7) == 0), PG_ERR_ARG,
             "pggan_rgb_expand: misaligned pointer");
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    const size_t smem = (size_t)C * (cimg + 1) * sizeof(float);
    PG_CHECK(smem <= 48 * 1024, PG_ERR_UNSUPPORTED, "pggan_rgb_expand: C=%d too large", C);
    if (dtype == PG_BF16) {
        PG_CHECK(C % 8 == 0, PG_ERR_UNSUPPORTED, "pggan_rgb_expand: C=%d must be a multiple of 8", C);
        rgb_expand_kernel<__nv_bfloat16><<<grid_for((long long)B * H * W * (C / 8), 256, 8), 256, smem, s>>>(
            img, w, bias, (__nv_bfloat16*)out, B, H, W, C, cimg, pool, scale, slope);
    } else if (dtype == PG_F32) {
        PG_CHECK(C % 4 == 0, PG_ERR_UNSUPPORTED, "pggan_rgb_expand: C=%d must be a multiple of 4", C);
        rgb_expand_kernel<float><<<grid_for((long long)B * H * W * (C / 4), 256, 8), 256, smem, s>>>(
            img, w, bias, (float*)out, B, H, W, C, cimg, pool, scale, slope);
    } else {
        PG_CHECK(false, PG_ERR_ARG, "pggan_rgb_expand: bad dtype %d", dtype);
    }
    PG_LAUNCH_CHECK();
    g_other_launches++;
    return PG_OK;
}

extern "C" int pggan_rgb_reduce(const void* x, const float* w, const float* bias, float* out, int B, int H, int W, int C,
                                int cimg, int up, float scale, float up_scale, int dtype, void* stream) {
    PG_CHECK(x && w && out && B > 0 && H > 0 && W > 0 && C > 0, PG_ERR_ARG, "pggan_rgb_reduce: bad args");
    PG_CHECK(cimg >= 1 && cimg <= kMaxImgC, PG_ERR_UNSUPPORTED, "pggan_rgb_reduce: cimg=%d must be 1..4", cimg);
    PG_CHECK(al16(x) && (!up || (reinterpret_cast<uintptr_t>(out) & 7) == 0), PG_ERR_ARG,
             "pggan_rgb_reduce: misaligned pointer");
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    const size_t smem = (size_t)C * cimg * sizeof(float);
    PG_CHECK(smem <= 48 * 1024, PG_ERR_UNSUPPORTED, "pggan_rgb_reduce: C=%d too large", C);
    const int vn = dtype == PG_BF16 ? 8 : 4;
    PG_CHECK(dtype == PG_BF16 || dtype == PG_F32, PG_ERR_ARG, "pggan_rgb_reduce: bad dtype %d", dtype);
    PG_CHECK(C % vn == 0, PG_ERR_UNSUPPORTED, "pggan_rgb_reduce: C=%d must be a multiple of %d", C, vn);
    const int CV = C / vn;
    int LP = 4;                                   // >= cimg lanes are needed to write the outputs
    while (LP * 2 <= CV && LP * 2 <= 32) LP *= 2;
    const long long P = (long long)B * H * W;
    const long long warps = (P + (32 / LP) - 1) / (32 / LP);
    const int grid = grid_for(warps * 32, 256, 8);
    if (dtype == PG_BF16)
        rgb_reduce_kernel<__nv_bfloat16><<<grid, 256, smem, s>>>((const __nv_bfloat16*)x, w, bias, out, B, H, W, C, cimg, up,
                                                               scale, up_scale, LP);
    else
        rgb_reduce_kernel<float><<<grid, 256, smem, s>>>((const float*)x, w, bias, out, B, H, W, C, cimg, up, scale,
                                                       up_scale, LP);
    PG_LAUNCH_CHECK();
    g_other_launches++;
    return PG_OK;
}

extern "C" int pggan_rgb_wgrad(const void* f, const float* img, float* acc, int B, int H, int W, int C, int cimg,
                               int pool, float scale, int dtype, void* stream) {
    PG_CHECK(f && img && acc && B > 0 && H > 0 && W > 0 && C > 0, PG_ERR_ARG, "pggan_rgb_wgrad: bad args");
    PG_CHECK(cimg >= 1 && cimg <= kMaxImgC, PG_ERR_UNSUPPORTED, "pggan_rgb_wgrad: cimg=%d must be 1..4", cimg);
    PG_CHECK(al16(f) && (!pool || (reinterpret_cast<uintptr_t>(img) & 7) == 0), PG_ERR_ARG,
             "pggan_rgb_wgrad: misaligned pointer");
    PG_CHECK(dtype == PG_BF16 || dtype == PG_F32, PG_ERR_ARG, "pggan_rgb_wgrad: bad dtype %d", dtype);
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    const int vn = dtype == PG_BF16 ? 8 : 4;
    PG_CHECK(C % vn == 0, PG_ERR_UNSUPPORTED, "pggan_rgb_wgrad: C=%d must be a multiple of %d", C, vn);
    const int CV = C / vn;
    const int threads = 128;
    int TX = 1;
    while (TX * 2 <= CV && TX * 2 <= threads) TX *= 2;
    const long long P = (long long)B * H * W;
    long long blocks = (P + 255) / 256;
    const long long cap = (long long)num_sms() * 4;
    if (blocks > cap) blocks = cap;
    const long long ppb = (P + blocks - 1) / blocks;
    const int grid = (int)((P + ppb - 1) / ppb);
    const size_t smem = (size_t)threads * vn * kMaxImgC * sizeof(float);
    if (dtype == PG_BF16)
        rgb_wgrad_kernel<__nv_bfloat16><<<grid, threads, smem, s>>>((const __nv_bfloat16*)f, img, acc, B, H, W, C, cimg, pool,
                                                                 scale, TX, ppb);
    else
        rgb_wgrad_kernel<float><<<grid, threads, smem, s>>>((const float*)f, img, acc, B, H, W, C, cimg, pool, scale, TX, ppb);
    PG_LAUNCH_CHECK();
    g_other_launches++;
    return PG_OK;
}

extern "C" int pggan_image_fade(const float* x, const float* low, float* out, int B, int Cimg, int H, int W, float alpha,
                                void* stream) {
    PG_CHECK(x && out && B > 0 && Cimg > 0 && H > 1 && W > 1 && H % 2 == 0 && W % 2 == 0, PG_ERR_ARG,
             "pggan_image_fade: bad args (H=%d W=%d must be even)", H, W);
    PG_CHECK((reinterpret_cast<uintptr_t>(x) & 7) == 0 && (reinterpret_cast<uintptr_t>(out) & 7) == 0, PG_ERR_ARG,
             "pggan_image_fade: misaligned pointer");
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    const long long planes = (long long)B * Cimg;
    image_fade_kernel<<<grid_for(planes * (H / 2) * (W / 2), 256, 8), 256, 0, s>>>(x, low, out, planes, H, W, alpha);
    PG_LAUNCH_CHECK();
    g_other_launches++;
    return PG_OK;
}

extern "C" int pggan_image_pool_sum(const float* x, float* out, int B, int Cimg, int H, int W, float scale, void* stream) {
    PG_CHECK(x && out && B > 0 && Cimg > 0 && H > 1 && W > 1 && H % 2 == 0 && W % 2 == 0, PG_ERR_ARG,
             "pggan_image_pool_sum: bad args");
    PG_CHECK((reinterpret_cast<uintptr_t>(x) & 7) == 0, PG_ERR_ARG, "pggan_image_pool_sum: misaligned pointer");
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    const long long planes = (long long)B * Cimg;
    image_pool_sum_kernel<<<grid_for(planes * (H / 2) * (W / 2), 256, 8), 256, 0, s>>>(x, out, planes, H, W, scale);
    PG_LAUNCH_CHECK();
    g_other_launches++;
    return PG_OK;
}
