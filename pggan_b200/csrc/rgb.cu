// 1x1 RGB convolutions and RGB-space fades: pure bandwidth kernels (N=3 / K=3 is not tensor-core
// shaped).  Feature maps NHWC (bf16 / fp32), images fp32 NCHW as the reference passes them.
// Reference: toRGBBlock / fromRGBBlock (lib/blocks.py:153-170, 271-292), G's RGB fade
// (pggan/nets.py:155-156), the real-image fade (pggan/model.py:217-221).
#include "common.cuh"

#include <mutex>
#include <unordered_map>

#include <atomic>
#include <stdlib.h>

namespace pg {
extern std::atomic<long long> g_other_launches;

constexpr int kMaxImgC = 4;

template <typename T> struct RV {
    static constexpr int N = 16 / sizeof(T);
    uint4 raw;
    __device__ __forceinline__ float get(int i) const { return DT<T>::to_f(reinterpret_cast<const T*>(&raw)[i]); }
    __device__ __forceinline__ void set(int i, float v) { reinterpret_cast<T*>(&raw)[i] = DT<T>::from_f(v); }
};

// pixel index -> (b, h, w) with 32-bit arithmetic (host code guarantees B*H*W < 2^31)
__device__ __forceinline__ void pix_decomp(int pix, int H, int W, int& b, int& h, int& w) {
    // every PGGAN stage is a power of two: shifts instead of two runtime integer divisions (~40 instructions per pixel,
    // which these issue-bound kernels feel: ncu showed them at 50-60 % issue-slot utilisation and 11-27 % of DRAM peak)
    if (((W & (W - 1)) | (H & (H - 1))) == 0) {
        const int lw = __ffs(W) - 1, lh = __ffs(H) - 1;
        w = pix & (W - 1);
        h = (pix >> lw) & (H - 1);
        b = pix >> (lw + lh);
        return;
    }
    w = pix % W;
    const int t = pix / W;
    h = t % H;
    b = t / H;
}

__device__ __forceinline__ void load_pixel(const float* __restrict__ img, int cimg, int pool, int b, int h, int w,
                                           int H, int W, float (&p)[kMaxImgC]) {
    // (H, W) are the FEATURE-map dims; the image is [B,cimg,H,W] or, if pool, [B,cimg,2H,2W]
    if (!pool) {
        const size_t plane = static_cast<size_t>(H) * W;
#pragma unroll
        for (int j = 0; j < kMaxImgC; ++j)
            p[j] = j < cimg ? __ldg(img + (static_cast<size_t>(b) * cimg + j) * plane + h * W + w) : 0.f;
    } else {
        const int Wi = 2 * W;
        const size_t plane = static_cast<size_t>(4) * H * W;
#pragma unroll
        for (int j = 0; j < kMaxImgC; ++j) {
            if (j < cimg) {
                const float* q = img + (static_cast<size_t>(b) * cimg + j) * plane + static_cast<size_t>(2 * h) * Wi + 2 * w;
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
// Thread = one 16-byte channel vector (fixed for the whole kernel: its weights/bias live in registers)
// x a strided set of pixels; a block covers 256/TX consecutive pixels per iteration, so a warp writes
// 512 contiguous bytes.  TX = C/VN must divide 256.
template <typename T>
__global__ void __launch_bounds__(256)
rgb_expand_kernel(const float* __restrict__ img, const float* __restrict__ w, const float* __restrict__ bias,
                  T* __restrict__ out, int B, int H, int W, int C, int cimg, int pool, float scale, float slope,
                  int TX) {
    PG_PDL_ENTER();
    constexpr int VN = RV<T>::N;
    const int tx = threadIdx.x % TX, ty = threadIdx.x / TX, TY = blockDim.x / TX;
    float wr[VN][kMaxImgC], br[VN];
#pragma unroll
    for (int e = 0; e < VN; ++e) {
        const int c = tx * VN + e;
#pragma unroll
        for (int j = 0; j < kMaxImgC; ++j) wr[e][j] = j < cimg ? __ldg(w + c * cimg + j) * scale : 0.f;
        br[e] = bias ? __ldg(bias + c) * scale : 0.f;
    }
    const int P = B * H * W;
    for (int pix = blockIdx.x * TY + ty; pix < P; pix += gridDim.x * TY) {
        int b, hh, ww;
        pix_decomp(pix, H, W, b, hh, ww);
        float p[kMaxImgC];
        load_pixel(img, cimg, pool, b, hh, ww, H, W, p);
        RV<T> o;
#pragma unroll
        for (int e = 0; e < VN; ++e) {
            float acc = br[e];
#pragma unroll
            for (int j = 0; j < kMaxImgC; ++j) acc += wr[e][j] * p[j];
            o.set(e, lrelu_f(acc, slope));
        }
        *reinterpret_cast<uint4*>(out + static_cast<size_t>(pix) * C + tx * VN) = o.raw;
    }
}

// Narrow feature maps (C <= 64): one thread per pixel.  The vector-lane kernel above spends most of its issue slots on
// per-pixel index arithmetic that is amortised over only C/8 threads (ncu at C = 16, 1024^2: 69 us, 29 % of DRAM
// peak, issue slots 59 % busy); here a thread loads its pixel once, keeps the per-channel (w0,w1,w2,bias) in shared
// memory (one broadcast LDS.128 per output channel) and writes its pixel's C channels as CV 16-byte vectors.
template <typename T, int CV>
__global__ void __launch_bounds__(256)
rgb_expand_pixel_kernel(const float* __restrict__ img, const float* __restrict__ w, const float* __restrict__ bias,
                        T* __restrict__ out, int B, int H, int W, int cimg, int pool, float scale, float slope) {
    PG_PDL_ENTER();
    constexpr int VN = RV<T>::N;
    constexpr int C = CV * VN;
    __shared__ float4 ws[C];
    for (int c = threadIdx.x; c < C; c += blockDim.x)
        ws[c] = make_float4(__ldg(w + c * cimg) * scale, cimg > 1 ? __ldg(w + c * cimg + 1) * scale : 0.f,
                            cimg > 2 ? __ldg(w + c * cimg + 2) * scale : 0.f, bias ? __ldg(bias + c) * scale : 0.f);
    __syncthreads();
    const int P = B * H * W;
    for (int pix = blockIdx.x * blockDim.x + threadIdx.x; pix < P; pix += gridDim.x * blockDim.x) {
        int b, hh, ww;
        pix_decomp(pix, H, W, b, hh, ww);
        float p[kMaxImgC];
        load_pixel(img, cimg, pool, b, hh, ww, H, W, p);
        T* o = out + static_cast<size_t>(pix) * C;
#pragma unroll
        for (int k = 0; k < CV; ++k) {
            RV<T> v;
#pragma unroll
            for (int e = 0; e < VN; ++e) {
                const float4 wv = ws[k * VN + e];
                const float t = wv.w + wv.x * p[0] + wv.y * p[1] + wv.z * p[2];
                v.set(e, lrelu_f(t, slope));
            }
            reinterpret_cast<uint4*>(o)[k] = v.raw;
        }
    }
}

// ------------------------------------------------------------------------------ reduce (C -> cimg)
// LP lanes (power of two <= 32) cooperate on one pixel; each lane keeps the weights of its channel
// vectors in registers (NV = C/(VN*LP) vectors, <= 2 for C <= 512) and a warp handles 32/LP pixels x
// UNR pixels per iteration so several 16-byte loads are in flight per lane.
template <typename T, int NV>
__global__ void __launch_bounds__(256)
rgb_reduce_kernel(const T* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                  float* __restrict__ out, int B, int H, int W, int C, int cimg, int up, float scale, float up_scale,
                  int LP) {
    PG_PDL_ENTER();
    constexpr int VN = RV<T>::N;
    constexpr int UNR = 4;
    const int lane = threadIdx.x & 31;
    const int sub = lane % LP, grp = lane / LP, ppw = 32 / LP;
    float wr[NV][VN][kMaxImgC];
#pragma unroll
    for (int v = 0; v < NV; ++v)
#pragma unroll
        for (int j = 0; j < kMaxImgC; ++j) {
            const int c0 = (sub + v * LP) * VN;
#pragma unroll
            for (int e4 = 0; e4 < VN; e4 += 4) {
                float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
                if (j < cimg && c0 < C) t = __ldg(reinterpret_cast<const float4*>(w + j * C + c0 + e4));
                wr[v][e4 + 0][j] = t.x;
                wr[v][e4 + 1][j] = t.y;
                wr[v][e4 + 2][j] = t.z;
                wr[v][e4 + 3][j] = t.w;
            }
        }
    const int P = B * H * W;
    const int warp0 = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int base = warp0 * ppw * UNR; base < P; base += nwarps * ppw * UNR) {
        RV<T> v[UNR][NV];
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
            const int pix = base + u * ppw + grp;
#pragma unroll
            for (int k = 0; k < NV; ++k) {
                const int cv = sub + k * LP;
                if (pix < P && cv * VN < C)
                    v[u][k].raw = __ldg(reinterpret_cast<const uint4*>(x + static_cast<size_t>(pix) * C + cv * VN));
                else
                    v[u][k].raw = make_uint4(0, 0, 0, 0);
            }
        }
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
            const int pix = base + u * ppw + grp;
            float acc[kMaxImgC] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int k = 0; k < NV; ++k)
#pragma unroll
                for (int e = 0; e < VN; ++e) {
                    const float xv = v[u][k].get(e);
#pragma unroll
                    for (int j = 0; j < kMaxImgC; ++j) acc[j] += xv * wr[k][e][j];
                }
#pragma unroll
            for (int j = 0; j < kMaxImgC; ++j)
                for (int o = LP >> 1; o > 0; o >>= 1) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], o);
            if (pix < P && sub < cimg) {
                float r = 0.f;
#pragma unroll
                for (int j = 0; j < kMaxImgC; ++j)
                    if (j == sub) r = acc[j];
                r = (r + (bias ? __ldg(bias + sub) : 0.f)) * scale;
                int b, hh, ww;
                pix_decomp(pix, H, W, b, hh, ww);
                if (!up) {
                    out[(static_cast<size_t>(b * cimg + sub) * H + hh) * W + ww] = r;
                } else {
                    r *= up_scale;
                    float* o = out + (static_cast<size_t>(b * cimg + sub) * 2 * H + 2 * hh) * 2 * W + 2 * ww;
                    *reinterpret_cast<float2*>(o) = make_float2(r, r);
                    *reinterpret_cast<float2*>(o + 2 * W) = make_float2(r, r);
                }
            }
        }
    }
}

// Narrow feature maps (C <= 64): ONE THREAD PER PIXEL.  ncu on the lane-cooperative kernel above at C = 16, 1024^2:
// 178 us, 11 % of DRAM peak, issue-bound (half the lanes idle, shuffle reductions, 4-byte scattered plane stores).
// Here a thread reads its pixel's CV 16-byte vectors (consecutive threads -> consecutive pixels), keeps three fp32
// dot products, and writes one float per image plane: warp-coalesced 128-byte plane stores, no shuffles.
template <typename T, int CV>
__global__ void __launch_bounds__(256)
rgb_reduce_pixel_kernel(const T* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                        float* __restrict__ out, int B, int H, int W, int cimg, int up, float scale, float up_scale) {
    PG_PDL_ENTER();
    constexpr int VN = RV<T>::N;
    constexpr int C = CV * VN;
    __shared__ float4 ws[C];                // per channel: the (up to) three image-channel weights, one LDS.128
    for (int c = threadIdx.x; c < C; c += blockDim.x)
        ws[c] = make_float4(__ldg(w + c), cimg > 1 ? __ldg(w + C + c) : 0.f, cimg > 2 ? __ldg(w + 2 * C + c) : 0.f, 0.f);
    __syncthreads();
    const int P = B * H * W;
    for (int pix = blockIdx.x * blockDim.x + threadIdx.x; pix < P; pix += gridDim.x * blockDim.x) {
        RV<T> v[CV];
#pragma unroll
        for (int k = 0; k < CV; ++k) v[k].raw = __ldg(reinterpret_cast<const uint4*>(x + static_cast<size_t>(pix) * C) + k);
        float acc[3] = {0.f, 0.f, 0.f};
#pragma unroll
        for (int k = 0; k < CV; ++k)
#pragma unroll
            for (int e = 0; e < VN; ++e) {
                const float xv = v[k].get(e);
                const float4 wv = ws[k * VN + e];                                      // smem broadcast read
                acc[0] += xv * wv.x;
                acc[1] += xv * wv.y;
                acc[2] += xv * wv.z;
            }
        int b, hh, ww;
        pix_decomp(pix, H, W, b, hh, ww);
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            if (j < cimg) {
                float r = (acc[j] + (bias ? __ldg(bias + j) : 0.f)) * scale;
                if (!up) {
                    out[(static_cast<size_t>(b * cimg + j) * H + hh) * W + ww] = r;
                } else {
                    r *= up_scale;
                    float* o = out + (static_cast<size_t>(b * cimg + j) * 2 * H + 2 * hh) * 2 * W + 2 * ww;
                    *reinterpret_cast<float2*>(o) = make_float2(r, r);
                    *reinterpret_cast<float2*>(o + 2 * W) = make_float2(r, r);
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------ 1x1 weight gradient
// acc[c][j] += scale * sum_p f[p,c] * img_j(p), and (optionally) bsum[c] += scale * sum_p f[p,c].
// Threads (ty pixel lane, tx channel-vector lane, fixed per thread); each block owns a slab of pixels,
// 4 pixels per thread in flight; partials are combined in shared memory, one atomicAdd per (block, c, j).
template <typename T>
__global__ void __launch_bounds__(256)
rgb_wgrad_kernel(const T* __restrict__ f, const float* __restrict__ img, float* __restrict__ acc,
                 float* __restrict__ bsum, int B, int H, int W, int C, int cimg, int pool, float scale, int TX,
                 long long pix_per_block) {
    PG_PDL_ENTER();
    constexpr int VN = RV<T>::N;
    constexpr int NA = kMaxImgC + 1;
    extern __shared__ float red[];         // [blockDim][VN*NA]
    const int TY = blockDim.x / TX;
    const int tx = threadIdx.x % TX, ty = threadIdx.x / TX;
    const int P = B * H * W;
    const int p0 = static_cast<int>(blockIdx.x * pix_per_block);
    int p1 = p0 + static_cast<int>(pix_per_block);
    if (p1 > P) p1 = P;
    float a[VN][NA];
#pragma unroll
    for (int e = 0; e < VN; ++e)
#pragma unroll
        for (int j = 0; j < NA; ++j) a[e][j] = 0.f;
    for (int pix = p0 + ty; pix < p1; pix += TY) {
        int b, hh, ww;
        pix_decomp(pix, H, W, b, hh, ww);
        float p[kMaxImgC];
        load_pixel(img, cimg, pool, b, hh, ww, H, W, p);
        RV<T> v;
        v.raw = __ldg(reinterpret_cast<const uint4*>(f + static_cast<size_t>(pix) * C + tx * VN));
#pragma unroll
        for (int e = 0; e < VN; ++e) {
            const float fv = v.get(e);
#pragma unroll
            for (int j = 0; j < kMaxImgC; ++j) a[e][j] += fv * p[j];
            a[e][kMaxImgC] += fv;
        }
    }
#pragma unroll
    for (int e = 0; e < VN; ++e)
#pragma unroll
        for (int j = 0; j < NA; ++j) red[(threadIdx.x * VN + e) * NA + j] = a[e][j];
    __syncthreads();
    // TX*VN*NA outputs per block, reduced over TY by the first threads
    for (int o = threadIdx.x; o < TX * VN * NA; o += blockDim.x) {
        const int j = o % NA;
        const int ce = o / NA;            // tx*VN + e
        float t = 0.f;
        for (int yy = 0; yy < TY; ++yy) t += red[((yy * TX) * VN + ce) * NA + j];
        if (j < cimg)
            atomicAdd(acc + ce * cimg + j, t * scale);
        else if (j == kMaxImgC && bsum != nullptr)
            atomicAdd(bsum + ce, t * scale);
    }
}

// Narrow feature maps (C <= 32): one thread per pixel STREAM.  Each thread walks a strided set of pixels with all
// C x (cimg + 1) partial sums in registers (no per-pixel work is shared between threads, so the image values are
// loaded once per pixel instead of once per channel vector), then the block reduces with shuffles + one atomicAdd
// per output per block.
template <typename T, int CV>
__global__ void __launch_bounds__(256)
rgb_wgrad_pixel_kernel(const T* __restrict__ f, const float* __restrict__ img, float* __restrict__ acc,
                       float* __restrict__ bsum, int B, int H, int W, int cimg, int pool, float scale) {
    PG_PDL_ENTER();
    constexpr int VN = RV<T>::N;
    constexpr int C = CV * VN;
    __shared__ float red[8][C * 4];
    float a[C][4];
#pragma unroll
    for (int c = 0; c < C; ++c)
#pragma unroll
        for (int j = 0; j < 4; ++j) a[c][j] = 0.f;
    const int P = B * H * W;
    for (int pix = blockIdx.x * blockDim.x + threadIdx.x; pix < P; pix += gridDim.x * blockDim.x) {
        int b, hh, ww;
        pix_decomp(pix, H, W, b, hh, ww);
        float p[kMaxImgC];
        load_pixel(img, cimg, pool, b, hh, ww, H, W, p);
        RV<T> v[CV];
#pragma unroll
        for (int k = 0; k < CV; ++k) v[k].raw = __ldg(reinterpret_cast<const uint4*>(f + static_cast<size_t>(pix) * C) + k);
#pragma unroll
        for (int k = 0; k < CV; ++k)
#pragma unroll
            for (int e = 0; e < VN; ++e) {
                const float fv = v[k].get(e);
                a[k * VN + e][0] += fv * p[0];
                a[k * VN + e][1] += fv * p[1];
                a[k * VN + e][2] += fv * p[2];
                a[k * VN + e][3] += fv;
            }
    }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int c = 0; c < C; ++c)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float t = warp_sum(a[c][j]);
            if (lane == 0) red[wid][c * 4 + j] = t;
        }
    __syncthreads();
    for (int o = threadIdx.x; o < C * 4; o += blockDim.x) {
        float t = 0.f;
#pragma unroll
        for (int w2 = 0; w2 < 8; ++w2) t += red[w2][o];
        const int c = o >> 2, j = o & 3;
        if (j < cimg && j < 3)
            atomicAdd(acc + c * cimg + j, t * scale);
        else if (j == 3 && bsum != nullptr)
            atomicAdd(bsum + c, t * scale);
    }
}

// ------------------------------------------------------------------------------ tile kernels (round 2)
// ncu on the kernels above (profiles/r1_ncu_elementwise.md): 19-34 % of the HBM peak with the issue slots 50-60 % busy —
// every (pixel, channel-vector) visit paid the pixel's index arithmetic and its three image loads again.  The tile
// kernels below stage the image values of a 256-pixel tile ONCE in shared memory (coalesced plane loads, one float4
// per pixel), keep each thread on a FIXED channel vector (its weights live in registers) and walk the tile's pixels
// with purely linear addressing: per 16-byte feature vector one LDS.128, one LDG/STG.128 and the 24 FMAs the math
// needs.  A warp touches 512 contiguous bytes of the feature map per instruction.
constexpr int kTilePix = 256;

template <bool POOL>
__device__ __forceinline__ void stage_tile(const float* __restrict__ img, int cimg, int B, int H, int W, long long tile,
                                           int P, float4* spix) {
    const int pix = static_cast<int>(tile * kTilePix) + threadIdx.x;
    float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
    if (pix < P) {
        int b, hh, ww;
        pix_decomp(pix, H, W, b, hh, ww);
        float p[kMaxImgC];
        load_pixel(img, cimg, POOL ? 1 : 0, b, hh, ww, H, W, p);
        q = make_float4(p[0], p[1], p[2], p[3]);
    }
    spix[threadIdx.x] = q;
}

// out[pix][c] = lrelu(scale*(sum_j w[c][j]*p_j + bias[c]))        (fromRGB; also the dgrad of toRGB)
template <typename T, bool POOL>
__global__ void __launch_bounds__(256)
rgb_expand_tile_kernel(const float* __restrict__ img, const float* __restrict__ w, const float* __restrict__ bias,
                       T* __restrict__ out, int B, int H, int W, int C, int cimg, float scale, float slope, int TX) {
    PG_PDL_ENTER();
    constexpr int VN = RV<T>::N;
    __shared__ float4 spix[kTilePix];
    const int tx = threadIdx.x % TX, ty = threadIdx.x / TX, TY = kTilePix / TX;
    float w0[VN], w1[VN], w2[VN], br[VN];               // cimg <= 3 (RGB)
#pragma unroll
    for (int e = 0; e < VN; ++e) {
        const int c = tx * VN + e;
        w0[e] = __ldg(w + c * cimg) * scale;
        w1[e] = cimg > 1 ? __ldg(w + c * cimg + 1) * scale : 0.f;
        w2[e] = cimg > 2 ? __ldg(w + c * cimg + 2) * scale : 0.f;
        br[e] = bias ? __ldg(bias + c) * scale : 0.f;
    }
    const bool act = slope != 1.f;
    const int P = B * H * W;
    const long long tiles = (P + kTilePix - 1) / kTilePix;
    for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        __syncthreads();
        stage_tile<POOL>(img, cimg, B, H, W, tile, P, spix);
        __syncthreads();
        const int npx = (P - tile * kTilePix) < kTilePix ? static_cast<int>(P - tile * kTilePix) : kTilePix;
        T* obase = out + static_cast<size_t>(tile) * kTilePix * C + tx * VN;
#pragma unroll 4
        for (int p = ty; p < npx; p += TY) {
            const float4 q = spix[p];
            RV<T> o;
#pragma unroll
            for (int e = 0; e < VN; ++e) {
                float t = fmaf(w0[e], q.x, br[e]);
                t = fmaf(w1[e], q.y, t);
                t = fmaf(w2[e], q.z, t);
                if (act) t = t > 0.f ? t : t * slope;
                o.set(e, t);
            }
            *reinterpret_cast<uint4*>(obase + static_cast<size_t>(p) * C) = o.raw;
        }
    }
}

// out[b][j][h][w] = scale*(sum_c w[j][c]*x[pix][c] + bias[j])  (toRGB; also the dgrad of fromRGB), optionally written
// to the 2x2 block of a twice larger image.  A pixel's C/VN vectors are spread over LX = min(TX, 32) consecutive lanes
// (NV = TX/LX vectors per lane), partial dot products meet through shuffles, results are staged in shared memory and
// written as coalesced plane rows.
template <typename T, int NV>
__global__ void __launch_bounds__(256)
rgb_reduce_tile_kernel(const T* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                       float* __restrict__ out, int B, int H, int W, int C, int cimg, int up, float scale,
                       float up_scale, int LX) {
    PG_PDL_ENTER();
    constexpr int VN = RV<T>::N;
    __shared__ float sout[3][kTilePix];
    const int tx = threadIdx.x % LX, ty = threadIdx.x / LX, TY = kTilePix / LX;
    float w0[NV][VN], w1[NV][VN], w2[NV][VN];
#pragma unroll
    for (int k = 0; k < NV; ++k)
#pragma unroll
        for (int e = 0; e < VN; ++e) {
            const int c = (tx + k * LX) * VN + e;
            w0[k][e] = __ldg(w + c);
            w1[k][e] = cimg > 1 ? __ldg(w + C + c) : 0.f;
            w2[k][e] = cimg > 2 ? __ldg(w + 2 * C + c) : 0.f;
        }
    const int P = B * H * W;
    const long long tiles = (P + kTilePix - 1) / kTilePix;
    for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const int npx = (P - tile * kTilePix) < kTilePix ? static_cast<int>(P - tile * kTilePix) : kTilePix;
        const T* xbase = x + static_cast<size_t>(tile) * kTilePix * C + tx * VN;
        __syncthreads();                              // sout of the previous tile has been written out
#pragma unroll 2
        for (int p0 = ty; p0 < kTilePix; p0 += TY) {  // whole warps stay converged for the shuffles
            const bool live = p0 < npx;
            float a0 = 0.f, a1 = 0.f, a2 = 0.f;
            RV<T> v[NV];
#pragma unroll
            for (int k = 0; k < NV; ++k)
                v[k].raw = live ? __ldg(reinterpret_cast<const uint4*>(xbase + static_cast<size_t>(p0) * C + k * LX * VN))
                                : make_uint4(0, 0, 0, 0);
#pragma unroll
            for (int k = 0; k < NV; ++k)
#pragma unroll
                for (int e = 0; e < VN; ++e) {
                    const float xv = v[k].get(e);
                    a0 = fmaf(xv, w0[k][e], a0);
                    a1 = fmaf(xv, w1[k][e], a1);
                    a2 = fmaf(xv, w2[k][e], a2);
                }
            for (int o = LX >> 1; o > 0; o >>= 1) {
                a0 += __shfl_xor_sync(0xffffffffu, a0, o);
                a1 += __shfl_xor_sync(0xffffffffu, a1, o);
                a2 += __shfl_xor_sync(0xffffffffu, a2, o);
            }
            if (tx == 0 && live) {
                sout[0][p0] = a0;
                sout[1][p0] = a1;
                sout[2][p0] = a2;
            }
        }
        __syncthreads();
        if (static_cast<int>(threadIdx.x) < npx) {
            const int pix = static_cast<int>(tile * kTilePix) + threadIdx.x;
            int b, hh, ww;
            pix_decomp(pix, H, W, b, hh, ww);
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                if (j < cimg) {
                    float r = (sout[j][threadIdx.x] + (bias ? __ldg(bias + j) : 0.f)) * scale;
                    if (!up) {
                        out[(static_cast<size_t>(b * cimg + j) * H + hh) * W + ww] = r;
                    } else {
                        r *= up_scale;
                        float* o = out + (static_cast<size_t>(b * cimg + j) * 2 * H + 2 * hh) * 2 * W + 2 * ww;
                        *reinterpret_cast<float2*>(o) = make_float2(r, r);
                        *reinterpret_cast<float2*>(o + 2 * W) = make_float2(r, r);
                    }
                }
            }
        }
    }
}

// acc[c][j] += scale * sum_pix f[pix][c]*p_j,  bsum[c] += scale * sum_pix f[pix][c]
template <typename T, bool POOL>
__global__ void __launch_bounds__(256, 4)
rgb_wgrad_tile_kernel(const T* __restrict__ f, const float* __restrict__ img, float* __restrict__ acc,
                      float* __restrict__ bsum, int B, int H, int W, int C, int cimg, float scale, int TX) {
    PG_PDL_ENTER();
    constexpr int VN = RV<T>::N;
    __shared__ float4 spix[kTilePix];
    __shared__ float red[kTilePix][4 * VN + 1];
    const int tx = threadIdx.x % TX, ty = threadIdx.x / TX, TY = kTilePix / TX;
    float a0[VN], a1[VN], a2[VN], as[VN];
#pragma unroll
    for (int e = 0; e < VN; ++e) a0[e] = a1[e] = a2[e] = as[e] = 0.f;
    const int P = B * H * W;
    const long long tiles = (P + kTilePix - 1) / kTilePix;
    for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        __syncthreads();
        stage_tile<POOL>(img, cimg, B, H, W, tile, P, spix);
        __syncthreads();
        const int npx = (P - tile * kTilePix) < kTilePix ? static_cast<int>(P - tile * kTilePix) : kTilePix;
        const T* fbase = f + static_cast<size_t>(tile) * kTilePix * C + tx * VN;
#pragma unroll 4
        for (int p = ty; p < npx; p += TY) {
            RV<T> v;
            v.raw = __ldg(reinterpret_cast<const uint4*>(fbase + static_cast<size_t>(p) * C));
            const float4 q = spix[p];
#pragma unroll
            for (int e = 0; e < VN; ++e) {
                const float fv = v.get(e);
                a0[e] = fmaf(fv, q.x, a0[e]);
                a1[e] = fmaf(fv, q.y, a1[e]);
                a2[e] = fmaf(fv, q.z, a2[e]);
                as[e] += fv;
            }
        }
    }
    __syncthreads();
#pragma unroll
    for (int e = 0; e < VN; ++e) {
        red[threadIdx.x][4 * e + 0] = a0[e];
        red[threadIdx.x][4 * e + 1] = a1[e];
        red[threadIdx.x][4 * e + 2] = a2[e];
        red[threadIdx.x][4 * e + 3] = as[e];
    }
    __syncthreads();
    // TX*VN*4 outputs per block, each the sum over the TY threads that share a channel vector
    for (int o = threadIdx.x; o < TX * VN * 4; o += blockDim.x) {
        const int j = o & 3, ce = o >> 2;               // ce = tx*VN + e
        const int txx = ce / VN, e = ce - txx * VN;
        float t = 0.f;
        for (int yy = 0; yy < TY; ++yy) t += red[yy * TX + txx][4 * e + j];
        if (j < 3) {
            if (j < cimg) atomicAdd(acc + ce * cimg + j, t * scale);
        } else if (bsum != nullptr) {
            atomicAdd(bsum + ce, t * scale);
        }
    }
}

// ------------------------------------------------------------------------------ fp32 image kernels

__global__ void image_fade_kernel(const float* __restrict__ x, const float* __restrict__ low, float* __restrict__ out,
                                  long long planes, int H, int W, pg_coef calpha) {
    PG_PDL_ENTER();
    // one thread per 2x2 output block
    const float alpha = coef_val(calpha);
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
                                      int W, pg_coef cscale) {
    PG_PDL_ENTER();
    const float scale = coef_val(cscale);
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

// blocks of a persistent tile kernel: what the machine holds at once (a grid of, say, 8 blocks per SM with 5 resident
// runs 1.6 waves — the second one on 60 % of the slots; ncu on the tile kernels: 30-50 % achieved occupancy)
template <typename K>
static int resident_grid(K kern, long long tiles, int threads = 256) {
    // (the occupancy of a kernel does not change: asked once per kernel, the boxes' GPUs being identical)
    static std::mutex mu;
    static std::unordered_map<const void*, int> cache;
    int per_sm = 0;
    {
        std::lock_guard<std::mutex> lock(mu);
        auto it = cache.find(reinterpret_cast<const void*>(kern));
        if (it != cache.end()) {
            per_sm = it->second;
        } else {
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, 0) != cudaSuccess || per_sm < 1) per_sm = 1;
            cache.emplace(reinterpret_cast<const void*>(kern), per_sm);
        }
    }
    const long long cap = static_cast<long long>(num_sms()) * per_sm;
    return static_cast<int>(tiles < cap ? (tiles < 1 ? 1 : tiles) : cap);
}

}  // namespace pg

using namespace pg;

static inline int pow2_le(int v, int cap) {
    int r = 1;
    while (r * 2 <= v && r * 2 <= cap) r *= 2;
    return r;
}

extern "C" int pggan_rgb_expand(const float* img, const float* w, const float* bias, void* out, int B, int H, int W,
                                int C, int cimg, int pool, float scale, float slope, int dtype, void* stream) {
    PG_CHECK(img && w && out && B > 0 && H > 0 && W > 0 && C > 0, PG_ERR_ARG, "pggan_rgb_expand: bad args");
    PG_CHECK(cimg >= 1 && cimg <= kMaxImgC, PG_ERR_UNSUPPORTED, "pggan_rgb_expand: cimg=%d must be 1..4", cimg);
    PG_CHECK(al16(out) && (!pool || (reinterpret_cast<uintptr_t>(img) & 7) == 0), PG_ERR_ARG,
             "pggan_rgb_expand: misaligned pointer");
    PG_CHECK(dtype == PG_BF16 || dtype == PG_F16 || dtype == PG_F32, PG_ERR_ARG, "pggan_rgb_expand: bad dtype %d", dtype);
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    const int vn = is16(dtype) ? 8 : 4;
    PG_CHECK(C % vn == 0, PG_ERR_UNSUPPORTED, "pggan_rgb_expand: C=%d must be a multiple of %d", C, vn);
    const int TX = C / vn;
    PG_CHECK(TX <= 256 && 256 % TX == 0, PG_ERR_UNSUPPORTED,
             "pggan_rgb_expand: C/%d = %d must divide 256 (power-of-two channel counts up to %d)", vn, TX, 256 * vn);
    const long long P = (long long)B * H * W;
    const int TY = 256 / TX;
    PG_CHECK(P * (long long)C < (1LL << 40) && P < (1LL << 31), PG_ERR_UNSUPPORTED, "pggan_rgb_expand: tensor too large");
    static const bool rgb_old = getenv("PGGAN_RGB_OLD") != nullptr;
    if (!rgb_old && cimg <= 3) {
        // tile kernel: image values of 256 pixels staged once in smem, fixed channel vector per thread
        const long long tiles = (P + kTilePix - 1) / kTilePix;
        auto launch_t = [&](auto tag) -> int {
            using TT = decltype(tag);
            if (pool)
                PG_CUDA(pg::launch(rgb_expand_tile_kernel<TT, true>, resident_grid(rgb_expand_tile_kernel<TT, true>, tiles), 256, 0, s, 
                    img, w, bias, (TT*)out, B, H, W, C, cimg, scale, slope, TX));
            else
                PG_CUDA(pg::launch(rgb_expand_tile_kernel<TT, false>, resident_grid(rgb_expand_tile_kernel<TT, false>, tiles), 256, 0, s, 
                    img, w, bias, (TT*)out, B, H, W, C, cimg, scale, slope, TX));
            return PG_OK;
        };
        const int lrc = dtype == PG_BF16 ? launch_t(__nv_bfloat16()) : (dtype == PG_F16 ? launch_t(__half()) : launch_t(float()));
        if (lrc != PG_OK) return lrc;
        PG_LAUNCH_CHECK();
        g_other_launches++;
        return PG_OK;
    }
    const int CVn = C / vn;
    const int g2 = grid_for(P, 256, 8);
    const int grid = grid_for((P + TY - 1) / TY * 256 / 8 + 1, 256, 4);     // >= 8 pixels per thread
    auto launch = [&](auto tag) -> int {
        using TT = decltype(tag);
        // thread-per-pixel for C <= 32 (measured: at C = 64 the vector-lane kernel is faster, 67 vs 98 us)
        if (cimg <= 3 && CVn == 2)
            PG_CUDA(pg::launch(rgb_expand_pixel_kernel<TT, 2>, g2, 256, 0, s, img, w, bias, (TT*)out, B, H, W, cimg, pool, scale, slope));
        else if (cimg <= 3 && CVn == 4)
            PG_CUDA(pg::launch(rgb_expand_pixel_kernel<TT, 4>, g2, 256, 0, s, img, w, bias, (TT*)out, B, H, W, cimg, pool, scale, slope));
        else
            PG_CUDA(pg::launch(rgb_expand_kernel<TT>, grid, 256, 0, s, img, w, bias, (TT*)out, B, H, W, C, cimg, pool, scale, slope, TX));
        return PG_OK;
    };
    const int lrc2 = dtype == PG_BF16 ? launch(__nv_bfloat16()) : (dtype == PG_F16 ? launch(__half()) : launch(float()));
    if (lrc2 != PG_OK) return lrc2;
    PG_LAUNCH_CHECK();
    g_other_launches++;
    return PG_OK;
}

template <typename T>
static int rgb_reduce_launch(const void* x, const float* w, const float* bias, float* out, int B, int H, int W, int C,
                             int cimg, int up, float scale, float up_scale, cudaStream_t s) {
    constexpr int VN = 16 / sizeof(T);
    const int CV = C / VN;
    static const bool rgb_old = getenv("PGGAN_RGB_OLD") != nullptr;
    if (!rgb_old && cimg <= 3 && CV >= 1 && CV <= 128 && (CV & (CV - 1)) == 0 && (long long)B * H * W < (1LL << 31)) {
        // tile kernel: LX = min(CV, 32) lanes per pixel, NV = CV / LX vectors per lane
        const int LX = CV < 32 ? CV : 32, nv = CV / LX;
        const long long Pn = (long long)B * H * W;
        const long long tiles = (Pn + kTilePix - 1) / kTilePix;
        const T* xq = reinterpret_cast<const T*>(x);
        if (nv == 1)
            PG_CUDA(pg::launch(rgb_reduce_tile_kernel<T, 1>, resident_grid(rgb_reduce_tile_kernel<T, 1>, tiles), 256, 0, s, 
                xq, w, bias, out, B, H, W, C, cimg, up, scale, up_scale, LX));
        else if (nv == 2)
            PG_CUDA(pg::launch(rgb_reduce_tile_kernel<T, 2>, resident_grid(rgb_reduce_tile_kernel<T, 2>, tiles), 256, 0, s, 
                xq, w, bias, out, B, H, W, C, cimg, up, scale, up_scale, LX));
        else
            PG_CUDA(pg::launch(rgb_reduce_tile_kernel<T, 4>, resident_grid(rgb_reduce_tile_kernel<T, 4>, tiles), 256, 0, s, 
                xq, w, bias, out, B, H, W, C, cimg, up, scale, up_scale, LX));
        PG_LAUNCH_CHECK();
        g_other_launches++;
        return PG_OK;
    }
    if (cimg <= 3 && (CV == 2 || CV == 4 || CV == 8) && (long long)B * H * W < (1LL << 31)) {
        const T* xq = reinterpret_cast<const T*>(x);
        const int P = B * H * W;
        const int grid = grid_for(P, 256, 8);
        if (CV == 2) PG_CUDA(pg::launch(rgb_reduce_pixel_kernel<T, 2>, grid, 256, 0, s, xq, w, bias, out, B, H, W, cimg, up, scale, up_scale));
        else if (CV == 4) PG_CUDA(pg::launch(rgb_reduce_pixel_kernel<T, 4>, grid, 256, 0, s, xq, w, bias, out, B, H, W, cimg, up, scale, up_scale));
        else PG_CUDA(pg::launch(rgb_reduce_pixel_kernel<T, 8>, grid, 256, 0, s, xq, w, bias, out, B, H, W, cimg, up, scale, up_scale));
        PG_LAUNCH_CHECK();
        g_other_launches++;
        return PG_OK;
    }
    int LP = pow2_le(CV < 4 ? 4 : CV, 32);      // >= cimg lanes are needed to write the outputs
    const int nv = (CV + LP - 1) / LP;
    PG_CHECK(nv <= 4, PG_ERR_UNSUPPORTED, "pggan_rgb_reduce: C=%d too large (max %d)", C, 4 * 32 * VN);
    const long long P = (long long)B * H * W;
    PG_CHECK(P < (1LL << 31), PG_ERR_UNSUPPORTED, "pggan_rgb_reduce: tensor too large");
    const long long warps = (P + (32 / LP) * 4 - 1) / ((32 / LP) * 4);
    const int grid = grid_for(warps * 32 / 8 + 1, 256, nv >= 2 ? 1 : 2);   // >= 8 iterations per warp; register-heavy
    const T* xp = reinterpret_cast<const T*>(x);
    switch (nv) {
        case 1: PG_CUDA(pg::launch(rgb_reduce_kernel<T, 1>, grid, 256, 0, s, xp, w, bias, out, B, H, W, C, cimg, up, scale, up_scale, LP)); break;
        case 2: PG_CUDA(pg::launch(rgb_reduce_kernel<T, 2>, grid, 256, 0, s, xp, w, bias, out, B, H, W, C, cimg, up, scale, up_scale, LP)); break;
        default: PG_CUDA(pg::launch(rgb_reduce_kernel<T, 4>, grid, 256, 0, s, xp, w, bias, out, B, H, W, C, cimg, up, scale, up_scale, LP)); break;
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
    PG_CHECK(dtype == PG_BF16 || dtype == PG_F16 || dtype == PG_F32, PG_ERR_ARG, "pggan_rgb_reduce: bad dtype %d", dtype);
    const int vn = is16(dtype) ? 8 : 4;
    PG_CHECK(C % vn == 0, PG_ERR_UNSUPPORTED, "pggan_rgb_reduce: C=%d must be a multiple of %d", C, vn);
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    if (dtype == PG_BF16)
        return rgb_reduce_launch<__nv_bfloat16>(x, w, bias, out, B, H, W, C, cimg, up, scale, up_scale, s);
    if (dtype == PG_F16) return rgb_reduce_launch<__half>(x, w, bias, out, B, H, W, C, cimg, up, scale, up_scale, s);
    return rgb_reduce_launch<float>(x, w, bias, out, B, H, W, C, cimg, up, scale, up_scale, s);
}

extern "C" int pggan_rgb_wgrad(const void* f, const float* img, float* acc, float* bsum, int B, int H, int W, int C,
                               int cimg, int pool, float scale, int dtype, void* stream) {
    PG_CHECK(f && img && acc && B > 0 && H > 0 && W > 0 && C > 0, PG_ERR_ARG, "pggan_rgb_wgrad: bad args");
    PG_CHECK(cimg >= 1 && cimg <= kMaxImgC, PG_ERR_UNSUPPORTED, "pggan_rgb_wgrad: cimg=%d must be 1..4", cimg);
    PG_CHECK(al16(f) && (!pool || (reinterpret_cast<uintptr_t>(img) & 7) == 0), PG_ERR_ARG,
             "pggan_rgb_wgrad: misaligned pointer");
    PG_CHECK(dtype == PG_BF16 || dtype == PG_F16 || dtype == PG_F32, PG_ERR_ARG, "pggan_rgb_wgrad: bad dtype %d", dtype);
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    const int vn = is16(dtype) ? 8 : 4;
    PG_CHECK(C % vn == 0, PG_ERR_UNSUPPORTED, "pggan_rgb_wgrad: C=%d must be a multiple of %d", C, vn);
    const int TX = C / vn;
    PG_CHECK(TX <= 256 && 256 % TX == 0, PG_ERR_UNSUPPORTED, "pggan_rgb_wgrad: C/%d = %d must divide 256", vn, TX);
    const long long P = (long long)B * H * W;
    PG_CHECK(P < (1LL << 31), PG_ERR_UNSUPPORTED, "pggan_rgb_wgrad: tensor too large");
    static const bool rgb_old = getenv("PGGAN_RGB_OLD") != nullptr;
    if (!rgb_old && cimg <= 3) {
        const long long tiles = (P + kTilePix - 1) / kTilePix;
        // (each block ends with TX*VN*4 atomics; one resident wave keeps their count at a few per SM)
        auto launch_t = [&](auto tag) -> int {
            using TT = decltype(tag);
            if (pool)
                PG_CUDA(pg::launch(rgb_wgrad_tile_kernel<TT, true>, resident_grid(rgb_wgrad_tile_kernel<TT, true>, tiles), 256, 0, s, 
                    (const TT*)f, img, acc, bsum, B, H, W, C, cimg, scale, TX));
            else
                PG_CUDA(pg::launch(rgb_wgrad_tile_kernel<TT, false>, resident_grid(rgb_wgrad_tile_kernel<TT, false>, tiles), 256, 0, s, 
                    (const TT*)f, img, acc, bsum, B, H, W, C, cimg, scale, TX));
            return PG_OK;
        };
        const int lrc = dtype == PG_BF16 ? launch_t(__nv_bfloat16()) : (dtype == PG_F16 ? launch_t(__half()) : launch_t(float()));
        if (lrc != PG_OK) return lrc;
        PG_LAUNCH_CHECK();
        g_other_launches++;
        return PG_OK;
    }
    const int TY = 256 / TX;
    long long ppb = (long long)TY * 16;                   // 16 pixels per thread
    long long blocks = (P + ppb - 1) / ppb;
    const long long cap = (long long)num_sms() * 16;     // (measured: fewer, longer blocks are slower here)
    if (blocks > cap) {
        blocks = cap;
        ppb = (P + blocks - 1) / blocks;
    }
    const int grid = (int)((P + ppb - 1) / ppb);
    const size_t smem = (size_t)256 * vn * (kMaxImgC + 1) * sizeof(float);
    const int g2 = grid_for(P / 8 + 1, 256, 4);          // pixel kernel: >= 8 pixels per thread, <= 4 blocks per SM
    cudaError_t attr_err = cudaSuccess;
    auto launch = [&](auto tag) -> int {
        using TT = decltype(tag);
        // thread-per-pixel-stream for C = 16 (C = 32 would need 128 accumulators per thread: 203 registers, slower)
        if (cimg <= 3 && TX == 2) {
            PG_CUDA(pg::launch(rgb_wgrad_pixel_kernel<TT, 2>, g2, 256, 0, s, (const TT*)f, img, acc, bsum, B, H, W, cimg, pool, scale));
            return PG_OK;
        }
        if (smem > 48 * 1024) {
            attr_err = cudaFuncSetAttribute(rgb_wgrad_kernel<TT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
        }
        PG_CUDA(pg::launch(rgb_wgrad_kernel<TT>, grid, 256, smem, s, (const TT*)f, img, acc, bsum, B, H, W, C, cimg, pool, scale, TX, ppb));
        return PG_OK;
    };
    const int lrc2 = dtype == PG_BF16 ? launch(__nv_bfloat16()) : (dtype == PG_F16 ? launch(__half()) : launch(float()));
    if (lrc2 != PG_OK) return lrc2;
    PG_CUDA(attr_err);
    PG_LAUNCH_CHECK();
    g_other_launches++;
    return PG_OK;
}

extern "C" int pggan_image_fade(const float* x, const float* low, float* out, int B, int Cimg, int H, int W, pg_coef alpha,
                                void* stream) {
    PG_CHECK(x && out && B > 0 && Cimg > 0 && H > 1 && W > 1 && H % 2 == 0 && W % 2 == 0, PG_ERR_ARG,
             "pggan_image_fade: bad args (H=%d W=%d must be even)", H, W);
    PG_CHECK((reinterpret_cast<uintptr_t>(x) & 7) == 0 && (reinterpret_cast<uintptr_t>(out) & 7) == 0, PG_ERR_ARG,
             "pggan_image_fade: misaligned pointer");
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    const long long planes = (long long)B * Cimg;
    PG_CUDA(pg::launch(image_fade_kernel, grid_for(planes * (H / 2) * (W / 2), 256, 8), 256, 0, s, x, low, out, planes, H, W, alpha));
    PG_LAUNCH_CHECK();
    g_other_launches++;
    return PG_OK;
}

extern "C" int pggan_image_pool_sum(const float* x, float* out, int B, int Cimg, int H, int W, pg_coef scale, void* stream) {
    PG_CHECK(x && out && B > 0 && Cimg > 0 && H > 1 && W > 1 && H % 2 == 0 && W % 2 == 0, PG_ERR_ARG,
             "pggan_image_pool_sum: bad args");
    PG_CHECK((reinterpret_cast<uintptr_t>(x) & 7) == 0, PG_ERR_ARG, "pggan_image_pool_sum: misaligned pointer");
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    const long long planes = (long long)B * Cimg;
    PG_CUDA(pg::launch(image_pool_sum_kernel, grid_for(planes * (H / 2) * (W / 2), 256, 8), 256, 0, s, x, out, planes, H, W, scale));
    PG_LAUNCH_CHECK();
    g_other_launches++;
    return PG_OK;
}
