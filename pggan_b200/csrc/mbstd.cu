// Minibatch standard deviation (reference lib/blocks.py:204-233): forward statistic, its gradient and
// the gradient of that gradient (needed by the R1 / WGAN-GP double backward).
// x is [B][K]; a group is `n` consecutive samples; per feature k:
//   mu_k = mean_i x_ik, var_k = sum_i (x_ik-mu_k)^2/(n-1), sigma_k = sqrt(var_k + 1e-8), stat = mean_k sigma_k
// One thread owns a 16-byte vector of features for one group and walks the n samples (n is 4 in
// practice); per-group sums go through a block reduction and one atomicAdd.
#include "common.cuh"

#include <atomic>

namespace pg {
extern std::atomic<long long> g_other_launches;

constexpr float kMbEps = 1e-8f;

template <typename T> struct MV {
    static constexpr int N = 16 / sizeof(T);
    uint4 raw;
    __device__ __forceinline__ float get(int i) const { return DT<T>::to_f(reinterpret_cast<const T*>(&raw)[i]); }
    __device__ __forceinline__ void set(int i, float v) { reinterpret_cast<T*>(&raw)[i] = DT<T>::from_f(v); }
};
template <typename T> __device__ __forceinline__ MV<T> mld(const T* p) {
    MV<T> v;
    v.raw = __ldg(reinterpret_cast<const uint4*>(p));
    return v;
}

// mean and sigma of one feature vector over the group's n samples
template <typename T>
__device__ __forceinline__ void group_moments(const T* __restrict__ xg, long long K, int n, long long k0,
                                              float (&mu)[MV<T>::N], float (&sigma)[MV<T>::N]) {
    constexpr int VN = MV<T>::N;
#pragma unroll
    for (int e = 0; e < VN; ++e) mu[e] = 0.f;
    for (int i = 0; i < n; ++i) {
        MV<T> v = mld(xg + i * K + k0);
#pragma unroll
        for (int e = 0; e < VN; ++e) mu[e] += v.get(e);
    }
#pragma unroll
    for (int e = 0; e < VN; ++e) {
        mu[e] /= n;
        sigma[e] = 0.f;
    }
    for (int i = 0; i < n; ++i) {
        MV<T> v = mld(xg + i * K + k0);
#pragma unroll
        for (int e = 0; e < VN; ++e) {
            const float d = v.get(e) - mu[e];
            sigma[e] += d * d;
        }
    }
#pragma unroll
    for (int e = 0; e < VN; ++e) sigma[e] = sqrtf(sigma[e] / (n - 1) + kMbEps);
}

template <typename T>
__global__ void mbstd_fwd_kernel(const T* __restrict__ x, float* __restrict__ stat, long long K, int n) {
    PG_PDL_ENTER();
    constexpr int VN = MV<T>::N;
    __shared__ float red[32];
    const int grp = blockIdx.y;
    const T* xg = x + static_cast<long long>(grp) * n * K;
    const long long KV = K / VN;
    float local = 0.f;
    for (long long kv = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; kv < KV;
         kv += static_cast<long long>(gridDim.x) * blockDim.x) {
        float mu[VN], sg[VN];
        group_moments<T>(xg, K, n, kv * VN, mu, sg);
#pragma unroll
        for (int e = 0; e < VN; ++e) local += sg[e];
    }
    const float tot = block_sum(local, red);
    if (threadIdx.x == 0) atomicAdd(stat + grp, tot / static_cast<float>(K));
}

// TG: element type of the gradient tensors when it differs from the activation type T (mixed precision)
template <typename T, typename TG = T>
__global__ void mbstd_bwd_kernel(const T* __restrict__ x, const float* __restrict__ gs, TG* __restrict__ gx, long long K,
                                 int n) {
    PG_PDL_ENTER();
    constexpr int VN = MV<T>::N;
    const int grp = blockIdx.y;
    const T* xg = x + static_cast<long long>(grp) * n * K;
    TG* gg = gx + static_cast<long long>(grp) * n * K;
    const long long KV = K / VN;
    const float coef = gs[grp] / (static_cast<float>(K) * (n - 1));
    for (long long kv = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; kv < KV;
         kv += static_cast<long long>(gridDim.x) * blockDim.x) {
        float mu[VN], sg[VN];
        group_moments<T>(xg, K, n, kv * VN, mu, sg);
        for (int i = 0; i < n; ++i) {
            MV<T> v = mld(xg + i * K + kv * VN);
            MV<TG> o;
#pragma unroll
            for (int e = 0; e < VN; ++e) o.set(e, coef * (v.get(e) - mu[e]) / sg[e]);
            *reinterpret_cast<uint4*>(gg + i * K + kv * VN) = o.raw;
        }
    }
}

// u is the cotangent of gx.  ggs[grp] += sum_k S_k/(K(n-1)sigma_k),  S_k = sum_i u_ik d_ik
// gx2_jk = gs/(K(n-1)) * [ (u_jk - ubar_k)/sigma_k - d_jk S_k / ((n-1) sigma_k^3) ]
template <typename T, typename TG = T>
__global__ void mbstd_bwd2_kernel(const T* __restrict__ x, const TG* __restrict__ u, const float* __restrict__ gs,
                                  float* __restrict__ ggs, TG* __restrict__ gx2, long long K, int n) {
    PG_PDL_ENTER();
    constexpr int VN = MV<T>::N;
    __shared__ float red[32];
    const int grp = blockIdx.y;
    const long long goff = static_cast<long long>(grp) * n * K;
    const T* xg = x + goff;
    const TG* ug = u + goff;
    TG* og = gx2 + goff;
    const long long KV = K / VN;
    const float inv = 1.f / (static_cast<float>(K) * (n - 1));
    const float coef = gs[grp] * inv;
    float local = 0.f;
    for (long long kv = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; kv < KV;
         kv += static_cast<long long>(gridDim.x) * blockDim.x) {
        float mu[VN], sg[VN], ub[VN], S[VN];
        group_moments<T>(xg, K, n, kv * VN, mu, sg);
#pragma unroll
        for (int e = 0; e < VN; ++e) ub[e] = S[e] = 0.f;
        for (int i = 0; i < n; ++i) {
            MV<T> xv = mld(xg + i * K + kv * VN);
            MV<TG> uv = mld(ug + i * K + kv * VN);
#pragma unroll
            for (int e = 0; e < VN; ++e) {
                ub[e] += uv.get(e);
                S[e] += uv.get(e) * (xv.get(e) - mu[e]);
            }
        }
#pragma unroll
        for (int e = 0; e < VN; ++e) {
            ub[e] /= n;
            local += S[e] / sg[e];
        }
        for (int i = 0; i < n; ++i) {
            MV<T> xv = mld(xg + i * K + kv * VN);
            MV<TG> uv = mld(ug + i * K + kv * VN), o;
#pragma unroll
            for (int e = 0; e < VN; ++e) {
                const float d = xv.get(e) - mu[e];
                const float s3 = sg[e] * sg[e] * sg[e];
                o.set(e, coef * ((uv.get(e) - ub[e]) / sg[e] - d * S[e] / ((n - 1) * s3)));
            }
            *reinterpret_cast<uint4*>(og + i * K + kv * VN) = o.raw;
        }
    }
    const float tot = block_sum(local, red);
    if (threadIdx.x == 0) atomicAdd(ggs + grp, tot * inv);
}

static int mb_check(const char* name, int B, int64_t K, int group, int dtype) {
    PG_CHECK(B > 0 && K > 0 && group > 1 && B % group == 0, PG_ERR_ARG, "%s: bad B=%d group=%d", name, B, group);
    PG_CHECK(dtype == PG_BF16 || dtype == PG_F16 || dtype == PG_F32, PG_ERR_ARG, "%s: bad dtype %d", name, dtype);
    const int vn = is16(dtype) ? 8 : 4;
    PG_CHECK(K % vn == 0, PG_ERR_UNSUPPORTED, "%s: K=%lld must be a multiple of %d", name, (long long)K, vn);
    return PG_OK;
}
static dim3 mb_grid(int64_t K, int vn, int groups) {
    long long bx = (K / vn + 255) / 256;
    if (bx > 64) bx = 64;
    if (bx < 1) bx = 1;
    return dim3((unsigned)bx, (unsigned)groups);
}

}  // namespace pg

using namespace pg;

extern "C" int pggan_mbstd_fwd(const void* x, float* stat, int B, int64_t K, int group, int dtype, void* stream) {
    PG_CHECK(x && stat, PG_ERR_ARG, "pggan_mbstd_fwd: null pointer");
    int rc = mb_check("pggan_mbstd_fwd", B, K, group, dtype);
    if (rc != PG_OK) return rc;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    PG_CUDA(cudaMemsetAsync(stat, 0, sizeof(float) * (B / group), s));
    if (dtype == PG_BF16)
        PG_CUDA(pg::launch(mbstd_fwd_kernel<__nv_bfloat16>, mb_grid(K, 8, B / group), 256, 0, s, (const __nv_bfloat16*)x, stat, K, group));
    else if (dtype == PG_F16)
        PG_CUDA(pg::launch(mbstd_fwd_kernel<__half>, mb_grid(K, 8, B / group), 256, 0, s, (const __half*)x, stat, K, group));
    else
        PG_CUDA(pg::launch(mbstd_fwd_kernel<float>, mb_grid(K, 4, B / group), 256, 0, s, (const float*)x, stat, K, group));
    PG_LAUNCH_CHECK();
    g_other_launches++;
    return PG_OK;
}

extern "C" int pggan_mbstd_bwd(const void* x, const float* gs, void* gx, int B, int64_t K, int group, int dtype,
                               int g_dtype, void* stream) {
    PG_CHECK(x && gs && gx, PG_ERR_ARG, "pggan_mbstd_bwd: null pointer");
    int rc = mb_check("pggan_mbstd_bwd", B, K, group, dtype);
    if (rc != PG_OK) return rc;
    PG_CHECK(g_dtype == dtype || (is16(dtype) && is16(g_dtype)), PG_ERR_ARG, "pggan_mbstd_bwd: dtype %d / g_dtype %d cannot be mixed", dtype, g_dtype);
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    const dim3 grid = mb_grid(K, is16(dtype) ? 8 : 4, B / group);
    if (dtype == PG_BF16 && g_dtype == PG_BF16)
        PG_CUDA(pg::launch(mbstd_bwd_kernel<__nv_bfloat16>, grid, 256, 0, s, (const __nv_bfloat16*)x, gs, (__nv_bfloat16*)gx, K, group));
    else if (dtype == PG_F16 && g_dtype == PG_F16)
        PG_CUDA(pg::launch(mbstd_bwd_kernel<__half>, grid, 256, 0, s, (const __half*)x, gs, (__half*)gx, K, group));
    else if (dtype == PG_F16)
        PG_CUDA(pg::launch(mbstd_bwd_kernel<__half, __nv_bfloat16>, grid, 256, 0, s, (const __half*)x, gs, (__nv_bfloat16*)gx, K, group));
    else if (dtype == PG_BF16)
        PG_CUDA(pg::launch(mbstd_bwd_kernel<__nv_bfloat16, __half>, grid, 256, 0, s, (const __nv_bfloat16*)x, gs, (__half*)gx, K, group));
    else
        PG_CUDA(pg::launch(mbstd_bwd_kernel<float>, grid, 256, 0, s, (const float*)x, gs, (float*)gx, K, group));
    PG_LAUNCH_CHECK();
    g_other_launches++;
    return PG_OK;
}

extern "C" int pggan_mbstd_bwd2(const void* x, const void* u, const float* gs, float* ggs, void* gx2, int B, int64_t K,
                                int group, int dtype, int g_dtype, void* stream) {
    PG_CHECK(x && u && gs && ggs && gx2, PG_ERR_ARG, "pggan_mbstd_bwd2: null pointer");
    int rc = mb_check("pggan_mbstd_bwd2", B, K, group, dtype);
    if (rc != PG_OK) return rc;
    PG_CHECK(g_dtype == dtype || (is16(dtype) && is16(g_dtype)), PG_ERR_ARG, "pggan_mbstd_bwd2: dtype %d / g_dtype %d cannot be mixed", dtype, g_dtype);
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    PG_CUDA(cudaMemsetAsync(ggs, 0, sizeof(float) * (B / group), s));
    const dim3 grid = mb_grid(K, is16(dtype) ? 8 : 4, B / group);
    if (dtype == PG_BF16 && g_dtype == PG_BF16)
        PG_CUDA(pg::launch(mbstd_bwd2_kernel<__nv_bfloat16>, grid, 256, 0, s, (const __nv_bfloat16*)x, (const __nv_bfloat16*)u, gs, ggs,
                                                             (__nv_bfloat16*)gx2, K, group));
    else if (dtype == PG_F16 && g_dtype == PG_F16)
        PG_CUDA(pg::launch(mbstd_bwd2_kernel<__half>, grid, 256, 0, s, (const __half*)x, (const __half*)u, gs, ggs, (__half*)gx2, K, group));
    else if (dtype == PG_F16)
        PG_CUDA(pg::launch(mbstd_bwd2_kernel<__half, __nv_bfloat16>, grid, 256, 0, s, (const __half*)x, (const __nv_bfloat16*)u, gs, ggs,
                                                                     (__nv_bfloat16*)gx2, K, group));
    else if (dtype == PG_BF16)
        PG_CUDA(pg::launch(mbstd_bwd2_kernel<__nv_bfloat16, __half>, grid, 256, 0, s, (const __nv_bfloat16*)x, (const __half*)u, gs, ggs,
                                                                     (__half*)gx2, K, group));
    else
        PG_CUDA(pg::launch(mbstd_bwd2_kernel<float>, grid, 256, 0, s, (const float*)x, (const float*)u, gs, ggs, (float*)gx2, K, group));
    PG_LAUNCH_CHECK();
    g_other_launches++;
    return PG_OK;
}
