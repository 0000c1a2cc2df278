// GPU input pipeline of the training step: the transform chain of the reference's UnsupervisedDataset
// (lib/dataset.py:100-111 — Resize((s,s)) -> RandomHorizontalFlip -> ColorJitter(0.2,0.2,0.2,0.01) -> ToTensor ->
// Normalize(0.5,0.5)) and the real-image fade that follows it in train_step (pggan/model.py:217-221), as three
// bandwidth kernels on decoded uint8 / fp32 batches: an anti-aliased separable bilinear resize (the filter PIL and
// torchvision apply when down-scaling: a triangle of half-width max(scale,1)), a per-image statistics pass (ColorJitter's
// contrast blends with the mean grey level of the image as it is WHEN contrast is applied), and one pass that applies the
// sample's jitter operations in its random order, the flip, the normalisation to [-1,1] and the fade.
// The arithmetic is torchvision's float-tensor functional backend (transforms/_functional_tensor.py): same formulas,
// same operation order, fp32 — the PIL backend the reference happens to use quantises to uint8 after every operation.
#include "common.cuh"

#include <atomic>

namespace pg {
extern std::atomic<long long> g_other_launches;

// ------------------------------------------------------------------------------------------ resize
// ATen's _upsample_bilinear2d_aa weights (aten/src/ATen/native/cpu/UpSampleKernel.cpp, _compute_indices_min_size_weights_aa):
//   scale = in/out, support = max(scale,1), center = scale*(i+0.5), xmin = max(int(center-support+0.5),0),
//   xsize = min(int(center+support+0.5), in) - xmin, w_j = tri((j+xmin-center+0.5)/max(scale,1)), normalised.
__device__ __forceinline__ float tri(float x) {
    x = fabsf(x);
    return x < 1.f ? 1.f - x : 0.f;
}

struct AaSpan {
    int xmin, xsize;
    float center, invscale, inv_total;
};
__device__ __forceinline__ AaSpan aa_span(int i, int in_size, float scale) {
    AaSpan s;
    const float support = scale >= 1.f ? scale : 1.f;
    s.invscale = scale >= 1.f ? 1.f / scale : 1.f;
    s.center = scale * (i + 0.5f);
    s.xmin = max(static_cast<int>(s.center - support + 0.5f), 0);
    s.xsize = min(static_cast<int>(s.center + support + 0.5f), in_size) - s.xmin;
    float total = 0.f;
    for (int j = 0; j < s.xsize; ++j) total += tri((j + s.xmin - s.center + 0.5f) * s.invscale);
    s.inv_total = total != 0.f ? 1.f / total : 0.f;
    return s;
}

// one pass along one axis: out[p][o] = sum_j w_j * in[p][xmin + j]   (elements along the axis are `stride` apart,
// `inner` contiguous elements ride along: horizontal pass inner = 1, vertical pass inner = row length)
template <typename TI>
__global__ void resize_axis_kernel(const TI* __restrict__ in, float* __restrict__ out, long long planes, int in_size,
                                   int out_size, long long inner, float scale, float in_mul) {
    PG_PDL_ENTER();
    const long long total = planes * out_size * inner;
    for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long k = idx % inner;
        const int o = static_cast<int>((idx / inner) % out_size);
        const long long pl = idx / (inner * out_size);
        const AaSpan s = aa_span(o, in_size, scale);
        const TI* src = in + (pl * in_size + s.xmin) * inner + k;
        float acc = 0.f;
        for (int j = 0; j < s.xsize; ++j) {
            const float w = tri((j + s.xmin - s.center + 0.5f) * s.invscale) * s.inv_total;
            acc += w * (static_cast<float>(src[static_cast<long long>(j) * inner]) * in_mul);
        }
        out[idx] = acc;
    }
}

// ------------------------------------------------------------------------------------------ colour jitter
// per-sample parameters: [flip, op0, op1, op2, op3] (ints; op: 0 brightness, 1 contrast, 2 saturation, 3 hue, -1 none)
// and [brightness, contrast, saturation, hue] factors
struct Jit {
    int flip, op[4];
    float f[4];
};
__device__ __forceinline__ Jit load_jit(const int* __restrict__ ip, const float* __restrict__ fp, int b) {
    Jit j;
    j.flip = ip[b * 5];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        j.op[k] = ip[b * 5 + 1 + k];
        j.f[k] = fp[b * 4 + k];
    }
    return j;
}
__device__ __forceinline__ float clamp01(float v) { return fminf(fmaxf(v, 0.f), 1.f); }
__device__ __forceinline__ float grey(float r, float g, float b) { return 0.2989f * r + 0.587f * g + 0.114f * b; }

// torchvision _rgb2hsv / adjust_hue / _hsv2rgb on one pixel
__device__ __forceinline__ void hue_shift(float& r, float& g, float& b, float hue) {
    const float maxc = fmaxf(r, fmaxf(g, b)), minc = fminf(r, fminf(g, b));
    const bool eqc = maxc == minc;
    const float cr = maxc - minc;
    const float s = cr / (eqc ? 1.f : maxc);
    const float div = eqc ? 1.f : cr;
    const float rc = (maxc - r) / div, gc = (maxc - g) / div, bc = (maxc - b) / div;
    const float hr = (maxc == r) ? (bc - gc) : 0.f;
    const float hg = ((maxc == g) && (maxc != r)) ? (2.f + rc - bc) : 0.f;
    const float hb = ((maxc != g) && (maxc != r)) ? (4.f + gc - rc) : 0.f;
    float h = fmodf((hr + hg + hb) / 6.f + 1.f, 1.f);
    h = h + hue;
    h = h - floorf(h);                                  // python's % 1.0 (result in [0,1))
    const float v = maxc;
    const float i6 = floorf(h * 6.f);
    const float f = h * 6.f - i6;
    int i = static_cast<int>(i6) % 6;
    const float p = clamp01(v * (1.f - s)), q = clamp01(v * (1.f - s * f)), t = clamp01(v * (1.f - s * (1.f - f)));
    switch (i) {
        case 0: r = v; g = t; b = p; break;
        case 1: r = q; g = v; b = p; break;
        case 2: r = p; g = v; b = t; break;
        case 3: r = p; g = q; b = v; break;
        case 4: r = t; g = p; b = v; break;
        default: r = v; g = p; b = q; break;
    }
}

// apply the sample's operations in order; `upto_contrast`: stop right before the contrast op (statistics pass)
__device__ __forceinline__ void apply_jitter(float& r, float& g, float& b, const Jit& j, float mean, bool upto_contrast) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int op = j.op[k];
        const float f = j.f[op < 0 ? 0 : op];
        if (op == 0) {
            r = clamp01(r * f); g = clamp01(g * f); b = clamp01(b * f);
        } else if (op == 1) {
            if (upto_contrast) return;
            r = clamp01(f * r + (1.f - f) * mean); g = clamp01(f * g + (1.f - f) * mean); b = clamp01(f * b + (1.f - f) * mean);
        } else if (op == 2) {
            const float gr = grey(r, g, b);
            r = clamp01(f * r + (1.f - f) * gr); g = clamp01(f * g + (1.f - f) * gr); b = clamp01(f * b + (1.f - f) * gr);
        } else if (op == 3) {
            hue_shift(r, g, b, f);
        }
    }
}

// mean grey level of every image at the moment its contrast op runs
__global__ void jitter_stats_kernel(const float* __restrict__ img, const int* __restrict__ ip, const float* __restrict__ fp,
                                    float* __restrict__ mean, int HW) {
    PG_PDL_ENTER();
    __shared__ float red[32];
    const int b = blockIdx.y;
    const Jit j = load_jit(ip, fp, b);
    const float* base = img + static_cast<size_t>(b) * 3 * HW;
    float acc = 0.f;
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < HW; p += gridDim.x * blockDim.x) {
        float r = base[p], g = base[HW + p], bl = base[2 * HW + p];
        apply_jitter(r, g, bl, j, 0.f, true);
        acc += grey(r, g, bl);
    }
    const float tot = block_sum(acc, red);
    if (threadIdx.x == 0) atomicAdd(mean + b, tot / static_cast<float>(HW));
}

// out = fade( normalise( flip( jitter(img) ) ) ); one thread per 2x2 output block (the fade needs the block's mean)
__global__ void jitter_apply_kernel(const float* __restrict__ img, const int* __restrict__ ip, const float* __restrict__ fp,
                                    const float* __restrict__ mean, float* __restrict__ out, int B, int S, int fade,
                                    pg_coef calpha) {
    PG_PDL_ENTER();
    const float alpha = coef_val(calpha);
    const int Sh = S / 2, HW = S * S;
    const long long total = static_cast<long long>(B) * Sh * Sh;
    for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int xh = static_cast<int>(idx % Sh), yh = static_cast<int>((idx / Sh) % Sh);
        const int b = static_cast<int>(idx / (static_cast<long long>(Sh) * Sh));
        const Jit j = load_jit(ip, fp, b);
        const float m = mean[b];
        const float* base = img + static_cast<size_t>(b) * 3 * HW;
        float* obase = out + static_cast<size_t>(b) * 3 * HW;
        float v[4][3];
        float avg[3] = {0.f, 0.f, 0.f};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int y = 2 * yh + (q >> 1), x = 2 * xh + (q & 1);
            const int xs = j.flip ? (S - 1 - x) : x;                    // hflip: output x reads source S-1-x
            float r = base[y * S + xs], g = base[HW + y * S + xs], bl = base[2 * HW + y * S + xs];
            apply_jitter(r, g, bl, j, m, false);
            v[q][0] = (r - 0.5f) / 0.5f;                                // Normalize((0.5,)*3, (0.5,)*3)
            v[q][1] = (g - 0.5f) / 0.5f;
            v[q][2] = (bl - 0.5f) / 0.5f;
        }
        if (fade) {
#pragma unroll
            for (int c = 0; c < 3; ++c) avg[c] = (((v[0][c] + v[1][c]) + v[2][c]) + v[3][c]) * 0.25f;
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int y = 2 * yh + (q >> 1), x = 2 * xh + (q & 1);
#pragma unroll
            for (int c = 0; c < 3; ++c)
                obase[c * HW + y * S + x] = fade ? ((1.f - alpha) * avg[c] + alpha * v[q][c]) : v[q][c];
        }
    }
}

static inline int grid_of(long long items, int threads, int per_sm) {
    long long blocks = (items + threads - 1) / threads;
    const long long cap = static_cast<long long>(num_sms()) * per_sm;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return static_cast<int>(blocks);
}

}  // namespace pg

using namespace pg;

extern "C" int pggan_resize_aa(const void* src, int src_is_u8, float* tmp, float* dst, int planes, int Hs, int Ws, int Ho,
                               int Wo, float in_mul, void* stream) {
    PG_CHECK(src && tmp && dst && planes > 0 && Hs > 0 && Ws > 0 && Ho > 0 && Wo > 0, PG_ERR_ARG, "pggan_resize_aa: bad args");
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    // horizontal pass: [planes*Hs rows][Ws] -> tmp [planes*Hs][Wo]; vertical pass: [planes][Hs][Wo] -> [planes][Ho][Wo]
    const float sw = static_cast<float>(Ws) / Wo, sh = static_cast<float>(Hs) / Ho;
    const long long rows = static_cast<long long>(planes) * Hs;
    if (src_is_u8)
        PG_CUDA(pg::launch(resize_axis_kernel<uint8_t>, grid_of(rows * Wo, 256, 8), 256, 0, s, reinterpret_cast<const uint8_t*>(src), tmp, rows,
                                                                               Ws, Wo, 1, sw, in_mul));
    else
        PG_CUDA(pg::launch(resize_axis_kernel<float>, grid_of(rows * Wo, 256, 8), 256, 0, s, reinterpret_cast<const float*>(src), tmp, rows, Ws,
                                                                             Wo, 1, sw, in_mul));
    PG_LAUNCH_CHECK();
    PG_CUDA(pg::launch(resize_axis_kernel<float>, grid_of(static_cast<long long>(planes) * Ho * Wo, 256, 8), 256, 0, s, tmp, dst, planes, Hs, Ho,
                                                                                                       Wo, sh, 1.f));
    PG_LAUNCH_CHECK();
    g_other_launches += 2;
    return PG_OK;
}

extern "C" int pggan_jitter_normalize_fade(const float* img, const int32_t* iparams, const float* fparams, float* mean_ws,
                                           float* out, int B, int S, int fade, pg_coef alpha, void* stream) {
    PG_CHECK(img && iparams && fparams && mean_ws && out && B > 0 && S >= 2 && S % 2 == 0, PG_ERR_ARG,
             "pggan_jitter_normalize_fade: bad args (S=%d must be even)", S);
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    PG_CUDA(cudaMemsetAsync(mean_ws, 0, sizeof(float) * B, s));
    const int HW = S * S;
    int bx = (HW + 255) / 256;
    if (bx > 64) bx = 64;
    PG_CUDA(pg::launch(jitter_stats_kernel, dim3(bx, B), 256, 0, s, img, iparams, fparams, mean_ws, HW));
    PG_LAUNCH_CHECK();
    PG_CUDA(pg::launch(jitter_apply_kernel, grid_of(static_cast<long long>(B) * (S / 2) * (S / 2), 256, 8), 256, 0, s, img, iparams, fparams,
                                                                                                       mean_ws, out, B, S, fade,
                                                                                                       alpha));
    PG_LAUNCH_CHECK();
    g_other_launches += 2;
    return PG_OK;
}
