// WGAN-GP / R1 reductions over fp32 image gradients and the fused multi-tensor Adam step.
// Reference: pggan/loss.py:85-87 (per-sample ||grad||_2, (n-1)^2 sum), pggan/loss.py:98-100 (drift),
// lib/loss.py:132-134 (R1), torch.optim.Adam as configured in lib/model.py:95-97.
#include "common.cuh"

#include <atomic>

namespace pg {
extern std::atomic<long long> g_other_launches;

__global__ void rows_sumsq_kernel(const float* __restrict__ g, float* __restrict__ sumsq, long long n,
                                  long long per_block) {
    PG_PDL_ENTER();
    __shared__ float red[32];
    const int b = blockIdx.y;
    const float* row = g + static_cast<long long>(b) * n;
    const long long k0 = blockIdx.x * per_block;
    long long k1 = k0 + per_block;
    if (k1 > n) k1 = n;
    float acc = 0.f;
    // rows start 16-byte aligned when n % 4 == 0; otherwise fall back to scalar loads
    if ((n & 3) == 0 && (k0 & 3) == 0) {
        const long long v1 = k0 + ((k1 - k0) & ~3LL);
        for (long long k = k0 + 4LL * threadIdx.x; k < v1; k += 4LL * blockDim.x) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(row + k));
            acc += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
        }
        for (long long k = v1 + threadIdx.x; k < k1; k += blockDim.x) acc += row[k] * row[k];
    } else {
        for (long long k = k0 + threadIdx.x; k < k1; k += blockDim.x) acc += row[k] * row[k];
    }
    const float tot = block_sum(acc, red);
    if (threadIdx.x == 0) atomicAdd(sumsq + b, tot);
}

__global__ void rows_scale_kernel(const float* __restrict__ g, const float* __restrict__ coef, float* __restrict__ out,
                                  long long n, long long total) {
    PG_PDL_ENTER();
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x)
        out[i] = g[i] * __ldg(coef + i / n);
}

__global__ void gp_finalize_kernel(const float* __restrict__ sumsq, const float* __restrict__ pred,
                                   float* __restrict__ res, float* __restrict__ coef, int B, float w_gp, float w_drift) {
    PG_PDL_ENTER();
    __shared__ float red[32];
    float gp = 0.f, dr = 0.f;
    for (int b = threadIdx.x; b < B; b += blockDim.x) {
        const float nrm = sqrtf(sumsq[b]);
        gp += (nrm - 1.f) * (nrm - 1.f);
        if (coef != nullptr) coef[b] = w_gp * 2.f * (nrm - 1.f) / nrm;
        if (pred != nullptr) dr += pred[b] * pred[b];
    }
    gp = block_sum(gp, red);
    dr = block_sum(dr, red);
    if (threadIdx.x == 0) {
        res[0] = gp * w_gp;
        res[1] = dr * w_drift;
    }
}

// ----------------------------------------------------------------------------------------- Adam
constexpr int kAdamChunk = 16384;    // elements per block

__global__ void adam_kernel(const pg_adam_entry* __restrict__ table, const int32_t* __restrict__ chunk_tensor,
                            const int64_t* __restrict__ chunk_offset, const int64_t* __restrict__ step, float lr,
                            float beta1, float beta2, float eps, float grad_scale) {
    PG_PDL_ENTER();
    __shared__ float s_c[2];
    if (threadIdx.x == 0) {
        const double t = static_cast<double>(*step + 1);
        s_c[0] = static_cast<float>(1.0 - pow(static_cast<double>(beta1), t));
        s_c[1] = static_cast<float>(sqrt(1.0 - pow(static_cast<double>(beta2), t)));
    }
    __syncthreads();
    const float bc1 = s_c[0], bc2_sqrt = s_c[1];
    const float step_size = lr / bc1;
    const pg_adam_entry e = table[chunk_tensor[blockIdx.x]];
    const int64_t off = chunk_offset[blockIdx.x];
    int64_t end = off + kAdamChunk;
    if (end > e.numel) end = e.numel;
    auto upd = [&](float& p, float g, float& m, float& v) {
        g *= grad_scale;
        m = beta1 * m + (1.f - beta1) * g;
        v = beta2 * v + (1.f - beta2) * g * g;
        const float denom = sqrtf(v) / bc2_sqrt + eps;
        p = p - step_size * (m / denom);
    };
    const bool vec = ((reinterpret_cast<uintptr_t>(e.param) | reinterpret_cast<uintptr_t>(e.grad) |
                       reinterpret_cast<uintptr_t>(e.exp_avg) | reinterpret_cast<uintptr_t>(e.exp_avg_sq)) & 15) == 0 &&
                     (reinterpret_cast<uintptr_t>(e.shadow) & 7) == 0;
    uint16_t* sh = reinterpret_cast<uint16_t*>(e.shadow);        // packed 16-bit copy of the new value (may be NULL)
    const bool sh_f16 = e.shadow_dtype == PG_F16;
    int64_t i0 = off;
    if (vec) {                                        // chunk offsets are multiples of 4 elements
        const int64_t vend = off + ((end - off) & ~int64_t(3));
        for (int64_t i = off + 4 * threadIdx.x; i < vend; i += 4 * blockDim.x) {
            float4 p = *reinterpret_cast<float4*>(e.param + i);
            const float4 g = __ldg(reinterpret_cast<const float4*>(e.grad + i));
            float4 m = *reinterpret_cast<float4*>(e.exp_avg + i);
            float4 v = *reinterpret_cast<float4*>(e.exp_avg_sq + i);
            upd(p.x, g.x, m.x, v.x);
            upd(p.y, g.y, m.y, v.y);
            upd(p.z, g.z, m.z, v.z);
            upd(p.w, g.w, m.w, v.w);
            *reinterpret_cast<float4*>(e.param + i) = p;
            *reinterpret_cast<float4*>(e.exp_avg + i) = m;
            *reinterpret_cast<float4*>(e.exp_avg_sq + i) = v;
            if (sh != nullptr)
                *reinterpret_cast<uint2*>(sh + i) = make_uint2(pack16(p.x, p.y, sh_f16), pack16(p.z, p.w, sh_f16));
        }
        i0 = vend;
    }
    for (int64_t i = i0 + threadIdx.x; i < end; i += blockDim.x) {
        float p = e.param[i], m = e.exp_avg[i], v = e.exp_avg_sq[i];
        upd(p, e.grad[i], m, v);
        e.param[i] = p;
        e.exp_avg[i] = m;
        e.exp_avg_sq[i] = v;
        if (sh != nullptr) sh[i] = cvt16(p, sh_f16);
    }
}
__global__ void adam_incr_step_kernel(int64_t* step) {
    PG_PDL_ENTER(); *step += 1; }

}  // namespace pg

using namespace pg;

extern "C" int pggan_rows_sumsq(const float* g, float* sumsq, int B, int64_t n, void* stream) {
    PG_CHECK(g && sumsq && B > 0 && n > 0, PG_ERR_ARG, "pggan_rows_sumsq: bad args");
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    long long bx = (n + 8191) / 8192;
    const long long cap = ((long long)num_sms() * 8 + B - 1) / B;
    if (bx > cap) bx = cap;
    if (bx < 1) bx = 1;
    long long per_block = (n + bx - 1) / bx;
    per_block = (per_block + 3) & ~3LL;
    bx = (n + per_block - 1) / per_block;
    PG_CUDA(pg::launch(rows_sumsq_kernel, dim3((unsigned)bx, (unsigned)B), 256, 0, s, g, sumsq, n, per_block));
    PG_LAUNCH_CHECK();
    g_other_launches++;
    return PG_OK;
}

extern "C" int pggan_rows_scale(const float* g, const float* coef, float* out, int B, int64_t n, void* stream) {
    PG_CHECK(g && coef && out && B > 0 && n > 0, PG_ERR_ARG, "pggan_rows_scale: bad args");
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    const long long total = (long long)B * n;
    long long blocks = (total + 255) / 256;
    const long long cap = (long long)num_sms() * 16;
    if (blocks > cap) blocks = cap;
    PG_CUDA(pg::launch(rows_scale_kernel, (int)blocks, 256, 0, s, g, coef, out, n, total));
    PG_LAUNCH_CHECK();
    g_other_launches++;
    return PG_OK;
}

extern "C" int pggan_gp_finalize(const float* sumsq, const float* pred, float* res, float* coef, int B, float w_gp,
                                 float w_drift, void* stream) {
    PG_CHECK(sumsq && res && B > 0, PG_ERR_ARG, "pggan_gp_finalize: bad args");
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    PG_CUDA(pg::launch(gp_finalize_kernel, 1, 256, 0, s, sumsq, pred, res, coef, B, w_gp, w_drift));
    PG_LAUNCH_CHECK();
    g_other_launches++;
    return PG_OK;
}

extern "C" int pggan_adam_chunk_size(void) { return pg::kAdamChunk; }

extern "C" int pggan_adam_step(const pg_adam_entry* table, int n_tensors, int64_t total_chunks,
                               const int32_t* chunk_tensor, const int64_t* chunk_offset, int64_t* step, float lr,
                               float beta1, float beta2, float eps, float grad_scale, void* stream) {
    PG_CHECK(table && chunk_tensor && chunk_offset && step && n_tensors > 0 && total_chunks > 0, PG_ERR_ARG,
             "pggan_adam_step: bad args");
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    PG_CUDA(pg::launch(adam_kernel, (unsigned)total_chunks, 256, 0, s, table, chunk_tensor, chunk_offset, step, lr, beta1, beta2, eps,
                                                       grad_scale));
    PG_LAUNCH_CHECK();
    PG_CUDA(pg::launch(adam_incr_step_kernel, 1, 1, 0, s, step));
    PG_LAUNCH_CHECK();
    g_other_launches += 2;
    return PG_OK;
}
