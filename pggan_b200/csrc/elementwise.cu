// Vectorised elementwise / reduction kernels on NHWC feature maps: LeakyReLU mask, nearest-2x
// up-sampling, 2x2 pooling (+ fade blend), axpby, pixel norm fwd/bwd, column sums, weight (un)packing.
// All are HBM-bound: 16-byte accesses, grid sized in multiples of the SM count, fp32 math.
#include "common.cuh"

#include <atomic>

namespace pg {

extern std::atomic<long long> g_other_launches;

static inline int ew_grid(long long work_items, int threads) {
    long long blocks = (work_items + threads - 1) / threads;
    const long long cap = static_cast<long long>(num_sms()) * 16;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return static_cast<int>(blocks);
}

template <typename T> struct V16 {
    static constexpr int N = 16 / sizeof(T);
    uint4 raw;
    __device__ __forceinline__ float get(int i) const { return DT<T>::to_f(reinterpret_cast<const T*>(&raw)[i]); }
    __device__ __forceinline__ void set(int i, float v) { reinterpret_cast<T*>(&raw)[i] = DT<T>::from_f(v); }
    // "stored activation > 0" (LeakyReLU mask): the sign-magnitude test on the raw bits for the 16-bit types, so a mask
    // source may be bf16 or IEEE half whatever T is (mixed precision)
    __device__ __forceinline__ bool pos(int i) const { return act_pos(reinterpret_cast<const T*>(&raw)[i]); }
};
template <typename T> __device__ __forceinline__ V16<T> ld16(const T* p) {
    V16<T> v;
    v.raw = __ldg(reinterpret_cast<const uint4*>(p));
    return v;
}
template <typename T> __device__ __forceinline__ void st16(T* p, const V16<T>& v) {
    *reinterpret_cast<uint4*>(p) = v.raw;
}


#define GRID_STRIDE(i, n)                                                                      \
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < (n); \
         i += static_cast<long long>(gridDim.x) * blockDim.x)

// ------------------------------------------------------------------------- lrelu mask multiply
template <typename T>
__global__ void lrelu_mask_mul_kernel(const T* __restrict__ g, const T* __restrict__ y, T* __restrict__ out,
                                      long long nvec, long long n, float slope) {
    PG_PDL_ENTER();
    constexpr int VN = V16<T>::N;
    GRID_STRIDE(i, nvec) {
        V16<T> gv = ld16(g + i * VN), yv = ld16(y + i * VN), o;
#pragma unroll
        for (int e = 0; e < VN; ++e) o.set(e, yv.pos(e) ? gv.get(e) : gv.get(e) * slope);
        st16(out + i * VN, o);
    }
    // scalar tail
    const long long tail0 = nvec * VN;
    GRID_STRIDE(i, n - tail0) {
        const long long k = tail0 + i;
        const float gv = DT<T>::to_f(g[k]);
        out[k] = DT<T>::from_f(act_pos(y[k]) ? gv : gv * slope);
    }
}

// --------------------------------------------------------------------------------- axpby
template <typename T>
__global__ void axpby_kernel(const T* __restrict__ a, const T* __restrict__ b, T* __restrict__ out, long long nvec,
                             long long n, pg_coef ca, pg_coef cb) {
    PG_PDL_ENTER();
    constexpr int VN = V16<T>::N;
    const float sa = coef_val(ca), sb = coef_val(cb);
    GRID_STRIDE(i, nvec) {
        V16<T> av = ld16(a + i * VN), o;
        if (b != nullptr) {
            V16<T> bv = ld16(b + i * VN);
#pragma unroll
            for (int e = 0; e < VN; ++e) o.set(e, sa * av.get(e) + sb * bv.get(e));
        } else {
#pragma unroll
            for (int e = 0; e < VN; ++e) o.set(e, sa * av.get(e));
        }
        st16(out + i * VN, o);
    }
    const long long tail0 = nvec * VN;
    GRID_STRIDE(i, n - tail0) {
        const long long k = tail0 + i;
        float v = sa * DT<T>::to_f(a[k]);
        if (b != nullptr) v += sb * DT<T>::to_f(b[k]);
        out[k] = DT<T>::from_f(v);
    }
}


// (pixel, channel-vector) index -> (cv, w, h, b).  All PGGAN shapes are powers of two: shifts instead of three runtime
// integer divisions per element (these kernels run at 45-70 % issue-slot utilisation)
__device__ __forceinline__ void idx_decomp(long long i, int CV, int W, int H, int& cv, int& w, int& h, long long& b,
                                           long long& pix) {
    if (((CV & (CV - 1)) | (W & (W - 1)) | (H & (H - 1))) == 0) {
        const int lc = __ffs(CV) - 1, lw = __ffs(W) - 1, lh = __ffs(H) - 1;
        cv = static_cast<int>(i) & (CV - 1);
        pix = i >> lc;
        w = static_cast<int>(pix) & (W - 1);
        h = static_cast<int>(pix >> lw) & (H - 1);
        b = pix >> (lw + lh);
        return;
    }
    cv = static_cast<int>(i % CV);
    pix = i / CV;
    w = static_cast<int>(pix % W);
    h = static_cast<int>((pix / W) % H);
    b = pix / (static_cast<long long>(W) * H);
}

// ------------------------------------------------------------------------ upsample / pool
// one thread per (input pixel, 16-byte channel vector): read once, write the 2x2 block
template <typename T>
__global__ void upsample2x_kernel(const T* __restrict__ x, const T* __restrict__ mask_src, T* __restrict__ out, int B,
                                  int H, int W, int CV, pg_coef cscale, float slope) {
    PG_PDL_ENTER();
    constexpr int VN = V16<T>::N;
    const float scale = coef_val(cscale);
    const long long total = static_cast<long long>(B) * H * W * CV;
    const long long C = static_cast<long long>(CV) * VN;
    GRID_STRIDE(i, total) {
        int cv, w, h;
        long long b, pix;
        idx_decomp(i, CV, W, H, cv, w, h, b, pix);
        V16<T> v = ld16(x + pix * C + cv * VN);
        if (scale != 1.f) {
#pragma unroll
            for (int e = 0; e < VN; ++e) v.set(e, v.get(e) * scale);
        }
        const long long ooff = ((b * 2 * H + 2 * h) * 2 * W + 2 * w) * C + cv * VN;
        T* o = out + ooff;
        if (mask_src == nullptr) {
            st16(o, v);
            st16(o + C, v);
            st16(o + 2 * W * C, v);
            st16(o + 2 * W * C + C, v);
        } else {
            // avg-pool backward fused with the LeakyReLU backward of the pooled activation (mask_src = that
            // activation, full resolution): out = up(x) * (mask_src > 0 ? 1 : slope)
            const long long offs[4] = {0, C, 2 * W * C, 2 * W * C + C};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const V16<T> mv = ld16(mask_src + ooff + offs[k]);
                V16<T> r;
#pragma unroll
                for (int e = 0; e < VN; ++e) r.set(e, mv.pos(e) ? v.get(e) : v.get(e) * slope);
                st16(o + offs[k], r);
            }
        }
    }
}

// avg-pool backward + LeakyReLU mask as above, plus the bias gradient of the conv that produced the pooled activation:
// colsum[c] += cs_scale * sum over OUTPUT pixels of out[., c] (the values as rounded to T, so the result equals a
// separate column-sum pass over `out` — which this kernel saves).  Threads: tx = fixed channel vector, ty = input-pixel
// lane; per-thread partial sums meet in shared memory, one atomicAdd per (block, channel).
template <typename T>
__global__ void __launch_bounds__(256)
upsample2x_mask_colsum_kernel(const T* __restrict__ x, const T* __restrict__ mask_src, T* __restrict__ out,
                              float* __restrict__ colsum, int B, int H, int W, int C, pg_coef cscale, float slope,
                              float cs_scale, int TX, long long rows_per_block) {
    PG_PDL_ENTER();
    constexpr int VN = V16<T>::N;
    __shared__ float red[256 * VN];
    const float scale = coef_val(cscale);
    const int TY = blockDim.x / TX;
    const int tx = threadIdx.x % TX, ty = threadIdx.x / TX;
    const long long M = static_cast<long long>(B) * H * W;
    const long long r0 = blockIdx.x * rows_per_block;
    long long r1 = r0 + rows_per_block;
    if (r1 > M) r1 = M;
    const long long Cl = C;
    const long long offs[4] = {0, Cl, 2LL * W * Cl, 2LL * W * Cl + Cl};
    float acc[VN];
#pragma unroll
    for (int e = 0; e < VN; ++e) acc[e] = 0.f;
    for (long long m = r0 + ty; m < r1; m += TY) {
        const int w = static_cast<int>(m % W);
        const long long t = m / W;
        const int h = static_cast<int>(t % H);
        const long long b = t / H;
        V16<T> v = ld16(x + m * Cl + tx * VN);
        const long long ooff = ((b * 2 * H + 2 * h) * 2 * W + 2 * w) * Cl + tx * VN;
        V16<T> mv[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) mv[k] = ld16(mask_src + ooff + offs[k]);
        float vs[VN];
#pragma unroll
        for (int e = 0; e < VN; ++e) vs[e] = v.get(e) * scale;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            V16<T> r;
#pragma unroll
            for (int e = 0; e < VN; ++e) {
                r.set(e, mv[k].pos(e) ? vs[e] : vs[e] * slope);
                acc[e] += r.get(e);
            }
            st16(out + ooff + offs[k], r);
        }
    }
#pragma unroll
    for (int e = 0; e < VN; ++e) red[threadIdx.x * VN + e] = acc[e];
    __syncthreads();
    for (int o = threadIdx.x; o < TX * VN; o += blockDim.x) {
        float t = 0.f;
        for (int yy = 0; yy < TY; ++yy) t += red[yy * TX * VN + o];
        atomicAdd(colsum + o, t * cs_scale);
    }
}

// one thread per (output pixel, channel vector); x is [B,H,W,C], out [B,H/2,W/2,C]
template <typename T>
__global__ void pool2x2_kernel(const T* __restrict__ x, const T* __restrict__ blend, T* __restrict__ out, int B, int H,
                               int W, int CV, pg_coef cscale, pg_coef calpha) {
    PG_PDL_ENTER();
    constexpr int VN = V16<T>::N;
    const float alpha = coef_val(calpha), scale = coef_val(cscale);
    const int Ho = H / 2, Wo = W / 2;
    const long long total = static_cast<long long>(B) * Ho * Wo * CV;
    const long long C = static_cast<long long>(CV) * VN;
    GRID_STRIDE(i, total) {
        int cv, w, h;
        long long b, pix;
        idx_decomp(i, CV, Wo, Ho, cv, w, h, b, pix);
        const T* p = x + ((b * H + 2 * h) * W + 2 * w) * C + cv * VN;
        V16<T> v00 = ld16(p), v01 = ld16(p + C), v10 = ld16(p + W * C), v11 = ld16(p + W * C + C), o;
        if (blend != nullptr) {
            V16<T> bl = ld16(blend + pix * C + cv * VN);
#pragma unroll
            for (int e = 0; e < VN; ++e) {
                const float s = (((v00.get(e) + v01.get(e)) + v10.get(e)) + v11.get(e)) * scale;
                o.set(e, (1.f - alpha) * bl.get(e) + alpha * s);
            }
        } else {
#pragma unroll
            for (int e = 0; e < VN; ++e) o.set(e, (((v00.get(e) + v01.get(e)) + v10.get(e)) + v11.get(e)) * scale);
        }
        st16(out + pix * C + cv * VN, o);
    }
}

// --------------------------------------------------------------------------- pixel norm
// one warp per pixel row of C channels (C % VN == 0); row cached in registers when C <= 32*VN*MAXV
template <typename T, int MAXV>
__global__ void pixelnorm_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, float* __restrict__ rnorm,
                                     long long M, int C, float eps) {
    PG_PDL_ENTER();
    constexpr int VN = V16<T>::N;
    const int lane = threadIdx.x & 31;
    const long long warp0 = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5;
    const long long nwarps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
    const int nvec = C / VN;
    for (long long m = warp0; m < M; m += nwarps) {
        const T* xr = x + m * C;
        V16<T> v[MAXV];
        float ss = 0.f;
#pragma unroll
        for (int k = 0; k < MAXV; ++k) {
            const int vi = lane + 32 * k;
            if (vi < nvec) {
                v[k] = ld16(xr + vi * VN);
#pragma unroll
                for (int e = 0; e < VN; ++e) ss += v[k].get(e) * v[k].get(e);
            }
        }
        ss = warp_sum(ss);
        const float r = rsqrtf(ss / C + eps);
        if (lane == 0 && rnorm != nullptr) rnorm[m] = r;
        T* yr = y + m * C;
#pragma unroll
        for (int k = 0; k < MAXV; ++k) {
            const int vi = lane + 32 * k;
            if (vi < nvec) {
                V16<T> o;
#pragma unroll
                for (int e = 0; e < VN; ++e) o.set(e, v[k].get(e) * r);
                st16(yr + vi * VN, o);
            }
        }
    }
}

// gx = r * (gy - y * mean_c(gy*y)), then optional LeakyReLU mask by the sign of y; optionally the
// bias gradient colsum[c] += cs_scale * sum_m gx[m,c] in the same pass (lanes own fixed channel vectors)
// TY: element type of the saved activation y when it differs from the gradient type T (mixed precision: fp16 / bf16)
template <typename T, int MAXV, typename TY = T>
__global__ void __launch_bounds__(256)
pixelnorm_bwd_kernel(const T* __restrict__ gy, const TY* __restrict__ y, const float* __restrict__ rnorm,
                     T* __restrict__ gx, long long M, int C, float act_slope, float* __restrict__ colsum,
                     float cs_scale) {
    PG_PDL_ENTER();
    constexpr int VN = V16<T>::N;
    __shared__ float red[8][32 * VN + 1];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const long long warp0 = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5;
    const long long nwarps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
    const int nvec = C / VN;
    float cs[MAXV][VN];
#pragma unroll
    for (int k = 0; k < MAXV; ++k)
#pragma unroll
        for (int e = 0; e < VN; ++e) cs[k][e] = 0.f;
    for (long long m = warp0; m < M; m += nwarps) {
        V16<T> g[MAXV];
        V16<TY> yv[MAXV];
        float dot = 0.f;
#pragma unroll
        for (int k = 0; k < MAXV; ++k) {
            const int vi = lane + 32 * k;
            if (vi < nvec) {
                g[k] = ld16(gy + m * C + vi * VN);
                yv[k] = ld16(y + m * C + vi * VN);
#pragma unroll
                for (int e = 0; e < VN; ++e) dot += g[k].get(e) * yv[k].get(e);
            }
        }
        dot = warp_sum(dot) / C;
        const float r = rnorm[m];
#pragma unroll
        for (int k = 0; k < MAXV; ++k) {
            const int vi = lane + 32 * k;
            if (vi < nvec) {
                V16<T> o;
#pragma unroll
                for (int e = 0; e < VN; ++e) {
                    const float yy = yv[k].get(e);
                    float t = r * (g[k].get(e) - yy * dot);
                    if (!(yy > 0.f)) t *= act_slope;
                    o.set(e, t);
                    cs[k][e] += o.get(e);        // sum what is stored (rounded), as a separate pass would
                }
                st16(gx + m * C + vi * VN, o);
            }
        }
    }
    if (colsum != nullptr) {
#pragma unroll
        for (int k = 0; k < MAXV; ++k) {
            __syncthreads();
#pragma unroll
            for (int e = 0; e < VN; ++e) red[wid][lane * VN + e] = cs[k][e];
            __syncthreads();
            const int nw = blockDim.x >> 5;
            for (int o = threadIdx.x; o < 32 * VN; o += blockDim.x) {
                const int vi = o / VN + 32 * k;
                if (vi < nvec) {
                    float t = 0.f;
                    for (int w2 = 0; w2 < nw; ++w2) t += red[w2][o];
                    atomicAdd(colsum + vi * VN + (o % VN), t * cs_scale);
                }
            }
        }
    }
}

// out = g * mask(y) (y may be NULL: plain copy-free column sum) and colsum[c] += scale * sum_m out[m,c].
// Thread (ty row lane, tx channel-vector lane, fixed per thread); 4 rows in flight per thread.
template <typename T>
__global__ void __launch_bounds__(256)
mask_colsum_kernel(const T* __restrict__ g, const T* __restrict__ y, T* __restrict__ out, float* __restrict__ colsum,
                   long long M, int C, float slope, float scale, int TX, long long rows_per_block) {
    PG_PDL_ENTER();
    constexpr int VN = V16<T>::N;
    __shared__ float red[256 * VN];
    const int TY = blockDim.x / TX;
    const int tx = threadIdx.x % TX, ty = threadIdx.x / TX;
    const long long r0 = blockIdx.x * rows_per_block;
    long long r1 = r0 + rows_per_block;
    if (r1 > M) r1 = M;
    float acc[VN];
#pragma unroll
    for (int e = 0; e < VN; ++e) acc[e] = 0.f;
    for (long long m = r0 + ty; m < r1; m += 4LL * TY) {
        V16<T> gv[4], yv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const long long mm = m + static_cast<long long>(u) * TY;
            if (mm < r1) {
                gv[u] = ld16(g + mm * C + tx * VN);
                if (y != nullptr) yv[u] = ld16(y + mm * C + tx * VN);
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const long long mm = m + static_cast<long long>(u) * TY;
            if (mm < r1) {
                V16<T> o;
#pragma unroll
                for (int e = 0; e < VN; ++e) {
                    float t = gv[u].get(e);
                    if (y != nullptr && !yv[u].pos(e)) t *= slope;
                    o.set(e, t);
                    acc[e] += o.get(e);
                }
                if (out != nullptr) st16(out + mm * C + tx * VN, o);
            }
        }
    }
    if (colsum != nullptr) {
#pragma unroll
        for (int e = 0; e < VN; ++e) red[threadIdx.x * VN + e] = acc[e];
        __syncthreads();
        for (int o = threadIdx.x; o < TX * VN; o += blockDim.x) {
            float t = 0.f;
            for (int yy = 0; yy < TY; ++yy) t += red[yy * TX * VN + o];
            atomicAdd(colsum + o, t * scale);
        }
    }
}

// column sums only (what is left of the LeakyReLU-backward passes once the mask is applied by the producing dgrad / pool
// backward: the bias gradient): no mask, no output tensor -> half the registers of the general kernel, eight 16-byte
// loads in flight per thread (ncu on the general kernel in this mode: 57 % of the measured HBM peak, occupancy 33 %)
template <typename T>
__global__ void __launch_bounds__(256)
colsum_only_kernel(const T* __restrict__ g, float* __restrict__ colsum, long long M, int C, float scale, int TX,
                   long long rows_per_block) {
    PG_PDL_ENTER();
    constexpr int VN = V16<T>::N;
    __shared__ float red[256 * VN];
    const int TY = blockDim.x / TX;
    const int tx = threadIdx.x % TX, ty = threadIdx.x / TX;
    const long long r0 = blockIdx.x * rows_per_block;
    long long r1 = r0 + rows_per_block;
    if (r1 > M) r1 = M;
    float acc[VN];
#pragma unroll
    for (int e = 0; e < VN; ++e) acc[e] = 0.f;
    constexpr int U = 8;
    for (long long m = r0 + ty; m < r1; m += static_cast<long long>(U) * TY) {
        V16<T> gv[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long mm = m + static_cast<long long>(u) * TY;
            if (mm < r1) gv[u] = ld16(g + mm * C + tx * VN);
            else gv[u].raw = make_uint4(0, 0, 0, 0);
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
            for (int e = 0; e < VN; ++e) acc[e] += gv[u].get(e);
    }
#pragma unroll
    for (int e = 0; e < VN; ++e) red[threadIdx.x * VN + e] = acc[e];
    __syncthreads();
    for (int o = threadIdx.x; o < TX * VN; o += blockDim.x) {
        float t = 0.f;
        for (int yy = 0; yy < TY; ++yy) t += red[yy * TX * VN + o];
        atomicAdd(colsum + o, t * scale);
    }
}

// -------------------------------------------------------------------------------- column sum
// out[c] += scale * sum_m g[m,c].  Each block owns a slab of rows; threads are (ty row-lane, tx
// channel-vector lane); per-thread partial sums are combined through shared memory and one
// atomicAdd per (block, channel) goes to global memory.
template <typename T>
__global__ void colsum_kernel(const T* __restrict__ g, float* __restrict__ out, long long M, int C, float scale,
                              int TX, long long rows_per_block) {
    PG_PDL_ENTER();
    constexpr int VN = V16<T>::N;
    __shared__ float red[256 * VN];
    const int CV = C / VN;
    const int TY = blockDim.x / TX;
    const int tx = threadIdx.x % TX, ty = threadIdx.x / TX;
    const long long r0 = blockIdx.x * rows_per_block;
    long long r1 = r0 + rows_per_block;
    if (r1 > M) r1 = M;
    for (int cv0 = 0; cv0 < CV; cv0 += TX) {
        const int cv = cv0 + tx;
        float acc[VN];
#pragma unroll
        for (int e = 0; e < VN; ++e) acc[e] = 0.f;
        if (cv < CV) {
            for (long long m = r0 + ty; m < r1; m += TY) {
                V16<T> v = ld16(g + m * C + cv * VN);
#pragma unroll
                for (int e = 0; e < VN; ++e) acc[e] += v.get(e);
            }
        }
        __syncthreads();
#pragma unroll
        for (int e = 0; e < VN; ++e) red[(ty * TX + tx) * VN + e] = acc[e];
        __syncthreads();
        if (ty == 0 && cv < CV) {
#pragma unroll
            for (int e = 0; e < VN; ++e) {
                float t = 0.f;
                for (int yy = 0; yy < TY; ++yy) t += red[(yy * TX + tx) * VN + e];
                atomicAdd(out + cv * VN + e, t * scale);
            }
        }
    }
}

// ---------------------------------------------------------------------------- weight packing
// logical operand [n][t][c]  <->  parameter memory: off = base + n1*s_n1 + n2*s_n2 + t*s_t + c1*s_c1 + c2*s_c2
// with n = n1*N2 + n2, c = c1*C2 + c2; rows n >= n_valid and columns c >= c_valid are zero padding.
struct PackMap {
    int n_total, n_valid, n2, taps, c_total, c_valid, c2;
    long long base, s_n1, s_n2, s_t, s_c1, s_c2;
};
__device__ __forceinline__ long long pack_off(const PackMap& m, int n, int t, int c) {
    return m.base + (n / m.n2) * m.s_n1 + (n % m.n2) * m.s_n2 + t * m.s_t + (c / m.c2) * m.s_c1 + (c % m.c2) * m.s_c2;
}

template <typename T>
__global__ void pack_weight_kernel(const float* __restrict__ in, T* __restrict__ out, PackMap m, long long total) {
    PG_PDL_ENTER();
    GRID_STRIDE(i, total) {
        const int c = static_cast<int>(i % m.c_total);
        const long long r = i / m.c_total;
        const int t = static_cast<int>(r % m.taps);
        const int n = static_cast<int>(r / m.taps);
        float v = 0.f;
        if (n < m.n_valid && c < m.c_valid) v = in[pack_off(m, n, t, c)];
        out[i] = DT<T>::from_f(v);
    }
}

__global__ void unpack_wgrad_kernel(const float* __restrict__ in, float* __restrict__ out, PackMap m, long long total,
                                    float scale, int accumulate, int slices, long long slice_elems) {
    PG_PDL_ENTER();
    GRID_STRIDE(i, total) {
        const int c = static_cast<int>(i % m.c_valid);
        const long long r = i / m.c_valid;
        const int t = static_cast<int>(r % m.taps);
        const int n = static_cast<int>(r / m.taps);
        float v = 0.f;
        for (int sl = 0; sl < slices; ++sl) v += in[sl * slice_elems + (static_cast<long long>(n) * m.taps + t) * m.c_total + c];
        v *= scale;
        float* o = out + pack_off(m, n, t, c);
        *o = accumulate ? (*o + v) : v;
    }
}

// Row-span path for the permuting maps of the big linear layers (and any other map whose logical row n lives in ONE
// short contiguous span of the parameter: D's 8192 -> 512 layer read in (h,w,c) order, G's 512 -> 8192 format layer
// with permuted rows, a [out,in] linear as it is).  One block per logical row: the span is moved between global and
// shared memory with coalesced 16-byte accesses, the permutation happens inside shared memory (index + index/32
// padding: the stride-16 walks of the (c,hw)<->(hw,c) flatten orders are bank-conflict free).  ncu on the generic
// kernels these replace: 30-78 us per 4.2 M-element weight at 72 % issue-slot utilisation and 6 % of the DRAM peak
// (64-bit div/mod per element).
constexpr int kRowSpanMax = 10240;      // floats of parameter span per logical row (padded: +1/32) -> < 48 KB static-free smem

__device__ __forceinline__ int pad32(int i) { return i + (i >> 5); }

struct RowSpan {
    long long span0;       // lowest relative offset of a row's elements (taps flipped: negative s_t)
    int span;              // elements in the span
};

__device__ __forceinline__ long long row_base(const PackMap& m, int n) {
    return m.base + static_cast<long long>(n / m.n2) * m.s_n1 + static_cast<long long>(n % m.n2) * m.s_n2;
}
__device__ __forceinline__ int rel_off(const PackMap& m, int t, int c, long long span0) {
    return static_cast<int>(t * m.s_t + (c / m.c2) * m.s_c1 + (c % m.c2) * m.s_c2 - span0);
}

template <typename T>
__global__ void __launch_bounds__(256) pack_rowspan_kernel(const float* __restrict__ in, T* __restrict__ out, PackMap m,
                                                           RowSpan rs) {
    PG_PDL_ENTER();
    extern __shared__ float srow[];
    const int n = blockIdx.x;
    T* orow = out + static_cast<size_t>(n) * m.taps * m.c_total;
    const int rowlen = m.taps * m.c_total;
    if (n >= m.n_valid) {
        for (int i = threadIdx.x; i < rowlen; i += blockDim.x) orow[i] = DT<T>::from_f(0.f);
        return;
    }
    const float* src = in + row_base(m, n) + rs.span0;
    if ((reinterpret_cast<uintptr_t>(src) & 15) == 0 && (rs.span & 3) == 0) {
        for (int i = threadIdx.x; i < rs.span / 4; i += blockDim.x) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(src) + i);
            srow[pad32(4 * i)] = v.x; srow[pad32(4 * i + 1)] = v.y; srow[pad32(4 * i + 2)] = v.z; srow[pad32(4 * i + 3)] = v.w;
        }
    } else {
        for (int i = threadIdx.x; i < rs.span; i += blockDim.x) srow[pad32(i)] = __ldg(src + i);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < rowlen; i += blockDim.x) {
        const int t = i / m.c_total, c = i - t * m.c_total;
        orow[i] = DT<T>::from_f(c < m.c_valid ? srow[pad32(rel_off(m, t, c, rs.span0))] : 0.f);
    }
}

__global__ void __launch_bounds__(256) unpack_rowspan_kernel(const float* __restrict__ in, float* __restrict__ out, PackMap m,
                                                             RowSpan rs, float scale, int accumulate, int slices,
                                                             long long slice_elems) {
    PG_PDL_ENTER();
    extern __shared__ float srow[];
    const int n = blockIdx.x;                    // < n_valid
    const int rowlen = m.taps * m.c_total;
    for (int i = threadIdx.x; i < rs.span + (rs.span >> 5) + 1; i += blockDim.x) srow[i] = 0.f;
    __syncthreads();
    const float* irow = in + static_cast<size_t>(n) * rowlen;
    for (int i = threadIdx.x; i < rowlen; i += blockDim.x) {
        const int t = i / m.c_total, c = i - t * m.c_total;
        if (c >= m.c_valid) continue;
        float v = 0.f;
        for (int sl = 0; sl < slices; ++sl) v += __ldg(irow + sl * slice_elems + i);
        srow[pad32(rel_off(m, t, c, rs.span0))] = v * scale;
    }
    __syncthreads();
    float* dst = out + row_base(m, n) + rs.span0;
    for (int i = threadIdx.x; i < rs.span; i += blockDim.x) {
        const float v = srow[pad32(i)];
        dst[i] = accumulate ? dst[i] + v : v;
    }
}

// Fast paths for the common OIHW 3x3 maps: one output channel (or a tile of them) per block, staged
// through shared memory so both the fp32 parameter side and the packed side are accessed contiguously
// with 16-byte transactions and enough bytes in flight per block.
constexpr int kPackMaxI = 1024;         // fast paths handle up to 1024 (padded) input channels

// fprop: out[o][t][i] = W[o][i][t]   (i padded to Ip with zeros); grid (O), dynamic smem 9*I floats.
// The parameter row is copied linearly (stride-9 reads from the tile are bank-conflict free: 9 is odd).
template <typename T>
__global__ void __launch_bounds__(256) pack_oihw_f_kernel(const float* __restrict__ W, T* __restrict__ out, int I, int Ip) {
    PG_PDL_ENTER();
    extern __shared__ float tile[];                        // [I][9] exactly as in the parameter
    const int o = blockIdx.x;
    const float* src = W + static_cast<size_t>(o) * I * 9;
    const int n = I * 9;
    if ((n & 3) == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
        for (int k = threadIdx.x; k < n / 4; k += blockDim.x)
            reinterpret_cast<float4*>(tile)[k] = __ldg(reinterpret_cast<const float4*>(src) + k);
    } else {
        for (int k = threadIdx.x; k < n; k += blockDim.x) tile[k] = src[k];
    }
    __syncthreads();
    T* dst = out + static_cast<size_t>(o) * 9 * Ip;
    for (int k = threadIdx.x; k < 9 * Ip; k += blockDim.x) {
        const int t = k / Ip, i = k - t * Ip;
        dst[k] = DT<T>::from_f(i < I ? tile[i * 9 + t] : 0.f);
    }
}
// gradients back: gW[o][i][t] (+)= scale * sum_s acc[s][o][t][i]; grid (O), dynamic smem 9*I floats
__global__ void __launch_bounds__(256) unpack_oihw_kernel(const float* __restrict__ acc, float* __restrict__ gW, int I,
                                                          int Ip, float scale, int accumulate, int slices,
                                                          long long slice_elems, int slices_per_block) {
    PG_PDL_ENTER();
    extern __shared__ float tile[];                        // [I][9]
    const int o = blockIdx.x;
    const float* src = acc + static_cast<size_t>(o) * 9 * Ip;
    // many slices (wgrad_rows writes one per CTA): blockIdx.y owns a group of slices and the groups meet through
    // atomics on gW (the caller has zeroed it when accumulate == 0)
    const int sl0 = blockIdx.y * slices_per_block;
    const int sl1 = (sl0 + slices_per_block < slices) ? (sl0 + slices_per_block) : slices;
    // blockIdx.z owns a chunk of input channels (more blocks in flight on the wide layers: 512 rows alone leave
    // most of the machine idle)
    const int cb = (I + gridDim.z - 1) / gridDim.z;
    const int i0 = blockIdx.z * cb;
    const int ni = (I - i0) < cb ? (I - i0) : cb;
    if (ni <= 0) return;
    for (int k = threadIdx.x; k < 9 * ni; k += blockDim.x) {
        const int t = k / ni, ii = k - t * ni;
        const int ks = t * Ip + i0 + ii;
        float v0 = 0.f, v1 = 0.f, v2 = 0.f, v3 = 0.f;
        int sl = sl0;
        for (; sl + 4 <= sl1; sl += 4) {
            v0 += __ldg(src + (sl + 0) * slice_elems + ks);
            v1 += __ldg(src + (sl + 1) * slice_elems + ks);
            v2 += __ldg(src + (sl + 2) * slice_elems + ks);
            v3 += __ldg(src + (sl + 3) * slice_elems + ks);
        }
        for (; sl < sl1; ++sl) v0 += __ldg(src + sl * slice_elems + ks);
        tile[ii * 9 + t] = ((v0 + v1) + (v2 + v3)) * scale;
    }
    __syncthreads();
    float* dst = gW + (static_cast<size_t>(o) * I + i0) * 9;
    const int n = ni * 9;
    if (gridDim.y > 1) {
        for (int k = threadIdx.x; k < n; k += blockDim.x) atomicAdd(dst + k, tile[k]);
        return;
    }
    if ((n & 3) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
        for (int k = threadIdx.x; k < n / 4; k += blockDim.x) {
            float4 v = reinterpret_cast<const float4*>(tile)[k];
            if (accumulate) {
                const float4 u = reinterpret_cast<const float4*>(dst)[k];
                v.x += u.x; v.y += u.y; v.z += u.z; v.w += u.w;
            }
            reinterpret_cast<float4*>(dst)[k] = v;
        }
    } else {
        for (int k = threadIdx.x; k < n; k += blockDim.x) dst[k] = accumulate ? (dst[k] + tile[k]) : tile[k];
    }
}
// dgrad: out[i][8-t][o] = W[o][i][t]   (rows i >= I zero); grid (ceil(Np/8), O/64), O % 64 == 0
template <typename T>
__global__ void __launch_bounds__(256) pack_oihw_t_kernel(const float* __restrict__ W, T* __restrict__ out, int O, int I,
                                                          int Np) {
    PG_PDL_ENTER();
    __shared__ float tile[64][73];
    const int i0 = blockIdx.x * 8, o0 = blockIdx.y * 64;
    const int ni = min(8, I - i0);
    for (int k = threadIdx.x; k < 64 * 72; k += blockDim.x) {
        const int oo = k / 72, r = k % 72;
        tile[oo][r] = (r < ni * 9) ? W[(static_cast<size_t>(o0 + oo) * I + i0) * 9 + r] : 0.f;
    }
    __syncthreads();
    for (int k = threadIdx.x; k < 72 * 64; k += blockDim.x) {
        const int oo = k % 64, r = k / 64;        // r = ii*9 + t'
        const int ii = r / 9, tp = r % 9;
        if (i0 + ii < Np)
            out[(static_cast<size_t>(i0 + ii) * 9 + tp) * O + o0 + oo] = DT<T>::from_f(tile[oo][ii * 9 + (8 - tp)]);
    }
}

// ------------------------------------------------------------------------------ channels_last (OHWI) weights
// A 3x3 conv weight kept in torch.channels_last memory format is physically [O][kh][kw][I] = the packed fprop operand
// [n][t][c] itself: the fprop pack is a cast (fused into the Adam kernel), the weight gradient un-pack a plain sum of
// the split-K slices.  Only the dgrad operand [i][8-t][o] needs a transpose: one 32x32 tile per block through smem.
template <typename T>
__global__ void __launch_bounds__(256) pack_ohwi_t_kernel(const float* __restrict__ W, T* __restrict__ out, int O, int I,
                                                          int Ip, int taps) {
    PG_PDL_ENTER();
    __shared__ float tile[32][33];
    const int t = blockIdx.z, i0 = blockIdx.x * 32, o0 = blockIdx.y * 32;
    const int lx = threadIdx.x & 31, ly = threadIdx.x >> 5;           // 32 x 8
#pragma unroll
    for (int r = ly; r < 32; r += 8) {                                 // read W[o0+r][t][i0+lx] (i contiguous)
        const int o = o0 + r, i = i0 + lx;
        tile[r][lx] = (o < O && i < I) ? __ldg(W + (static_cast<size_t>(o) * taps + t) * I + i) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int r = ly; r < 32; r += 8) {                                 // write out[i0+r][taps-1-t][o0+lx] (o contiguous)
        const int i = i0 + r, o = o0 + lx;
        if (i < Ip && o < O) out[(static_cast<size_t>(i) * taps + (taps - 1 - t)) * O + o] = DT<T>::from_f(tile[lx][r]);
    }
}

// gW[k] (+)= scale * sum_s acc[s][k]  for a dense [N][taps][C] slice layout that equals the gradient's own layout
__global__ void __launch_bounds__(256) unpack_dense_kernel(const float* __restrict__ acc, float* __restrict__ gW,
                                                           long long n4, float scale, int accumulate, int slices,
                                                           long long slice_elems) {
    PG_PDL_ENTER();
    GRID_STRIDE(i, n4) {
        float4 v = __ldg(reinterpret_cast<const float4*>(acc) + i);
        for (int sl = 1; sl < slices; ++sl) {
            const float4 u = __ldg(reinterpret_cast<const float4*>(acc + sl * slice_elems) + i);
            v.x += u.x; v.y += u.y; v.z += u.z; v.w += u.w;
        }
        v.x *= scale; v.y *= scale; v.z *= scale; v.w *= scale;
        if (accumulate) {
            const float4 u = reinterpret_cast<const float4*>(gW)[i];
            v.x += u.x; v.y += u.y; v.z += u.z; v.w += u.w;
        }
        reinterpret_cast<float4*>(gW)[i] = v;
    }
}

// ------------------------------------------------------------------------------ format conversion
template <typename TI, typename TO>
__global__ void convert_kernel(const TI* __restrict__ in, TO* __restrict__ out, long long n) {
    PG_PDL_ENTER();
    // 8 elements per thread: one or two 16-byte loads, one or two 16-byte stores
    const long long nvec = n / 8;
    GRID_STRIDE(i, nvec) {
        float v[8];
        if constexpr (sizeof(TI) == 2) {
            const uint4 raw = __ldg(reinterpret_cast<const uint4*>(in) + i);
            const TI* e = reinterpret_cast<const TI*>(&raw);
#pragma unroll
            for (int k = 0; k < 8; ++k) v[k] = DT<TI>::to_f(e[k]);
        } else {
            const float4 a = __ldg(reinterpret_cast<const float4*>(in) + 2 * i), b = __ldg(reinterpret_cast<const float4*>(in) + 2 * i + 1);
            v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
        }
        if constexpr (sizeof(TO) == 2) {
            uint4 o;
            TO* e = reinterpret_cast<TO*>(&o);
#pragma unroll
            for (int k = 0; k < 8; ++k) e[k] = DT<TO>::from_f(v[k]);
            reinterpret_cast<uint4*>(out)[i] = o;
        } else {
            reinterpret_cast<float4*>(out)[2 * i] = make_float4(v[0], v[1], v[2], v[3]);
            reinterpret_cast<float4*>(out)[2 * i + 1] = make_float4(v[4], v[5], v[6], v[7]);
        }
    }
    GRID_STRIDE(i, n - nvec * 8) {
        const long long k = nvec * 8 + i;
        out[k] = DT<TO>::from_f(DT<TI>::to_f(in[k]));
    }
}

// ------------------------------------------------------------------------------ 3xTF32 operand split
// in [M][C] fp32 -> three copies, one of which holds lo = x - trunc_tf32(x) (the tensor core truncates
// fp32 operands to tf32, so the plain copies act as hi).  Concatenated along channels ([M][3C]) for the
// K-major operands of the fprop-style GEMM, or along rows ([3M][C]) for the pixel-reduction of wgrad.
__global__ void split3_kernel(const float* __restrict__ in, float* __restrict__ out, long long M, int CV,
                              int along_channels, int lo_slot) {
    PG_PDL_ENTER();
    const long long total = M * CV;
    const long long C = 4LL * CV;
    GRID_STRIDE(i, total) {
        const long long m = i / CV;
        const int cv = static_cast<int>(i % CV);
        const float4 v = __ldg(reinterpret_cast<const float4*>(in + m * C) + cv);
        float4 lo;
        lo.x = v.x - __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u);
        lo.y = v.y - __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u);
        lo.z = v.z - __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u);
        lo.w = v.w - __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u);
#pragma unroll
        for (int slot = 0; slot < 3; ++slot) {
            float* dst = along_channels ? out + m * 3 * C + slot * C + cv * 4 : out + (slot * M + m) * C + cv * 4;
            *reinterpret_cast<float4*>(dst) = (slot == lo_slot) ? lo : v;
        }
    }
}

// narrow feature maps (C/VN = LP <= 16 vectors per pixel, LP a power of two): LP lanes per pixel, 32/LP pixels
// per warp, so every lane of a load instruction is live (the warp-per-pixel kernels above leave 30 of 32 lanes
// idle at C = 16: 131 / 357 us per 1024^2 map where the HBM floor is 41 / 61 us)
template <typename T>
__global__ void __launch_bounds__(256)
pixelnorm_fwd_small_kernel(const T* __restrict__ x, T* __restrict__ y, float* __restrict__ rnorm, long long M, int C,
                           float eps, int LP) {
    PG_PDL_ENTER();
    constexpr int VN = V16<T>::N;
    const long long nvec_total = M * LP;
    const float inv_c = 1.f / C;
    GRID_STRIDE(i, (nvec_total + 31) / 32 * 32) {           // whole warps stay converged for the shuffles
        const bool live = i < nvec_total;
        V16<T> v;
        float ss = 0.f;
        if (live) {
            v = ld16(x + i * VN);
#pragma unroll
            for (int e = 0; e < VN; ++e) ss += v.get(e) * v.get(e);
        }
        for (int o = LP >> 1; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
        if (live) {
            const float r = rsqrtf(ss * inv_c + eps);
            if (rnorm != nullptr && (i % LP) == 0) rnorm[i / LP] = r;
            V16<T> o;
#pragma unroll
            for (int e = 0; e < VN; ++e) o.set(e, v.get(e) * r);
            st16(y + i * VN, o);
        }
    }
}

template <typename T, typename TY = T>
__global__ void __launch_bounds__(256)
pixelnorm_bwd_small_kernel(const T* __restrict__ gy, const TY* __restrict__ y, const float* __restrict__ rnorm,
                           T* __restrict__ gx, long long M, int C, float act_slope, float* __restrict__ colsum,
                           float cs_scale, int LP) {
    PG_PDL_ENTER();
    constexpr int VN = V16<T>::N;
    __shared__ float red[8][16 * VN];
    const long long nvec_total = M * LP;
    const float inv_c = 1.f / C;
    float cs[VN];
#pragma unroll
    for (int e = 0; e < VN; ++e) cs[e] = 0.f;
    // the grid stride is a multiple of 32 and of LP, so a lane always owns the same channel vector (lane % LP)
    GRID_STRIDE(i, (nvec_total + 31) / 32 * 32) {
        const bool live = i < nvec_total;
        V16<T> g;
        V16<TY> yv;
        float dot = 0.f;
        if (live) {
            g = ld16(gy + i * VN);
            yv = ld16(y + i * VN);
#pragma unroll
            for (int e = 0; e < VN; ++e) dot += g.get(e) * yv.get(e);
        }
        for (int o = LP >> 1; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
        if (live) {
            dot *= inv_c;
            const float r = rnorm[i / LP];
            V16<T> o;
#pragma unroll
            for (int e = 0; e < VN; ++e) {
                const float yy = yv.get(e);
                float t = r * (g.get(e) - yy * dot);
                if (!(yy > 0.f)) t *= act_slope;
                o.set(e, t);
                cs[e] += o.get(e);
            }
            st16(gx + i * VN, o);
        }
    }
    if (colsum != nullptr) {
        const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
        for (int e = 0; e < VN; ++e)
            for (int o = 16; o >= LP; o >>= 1) cs[e] += __shfl_xor_sync(0xffffffffu, cs[e], o);
        if (lane < LP)
#pragma unroll
            for (int e = 0; e < VN; ++e) red[wid][lane * VN + e] = cs[e];
        __syncthreads();
        const int nw = blockDim.x >> 5;
        for (int o = threadIdx.x; o < LP * VN; o += blockDim.x) {
            float t = 0.f;
            for (int w2 = 0; w2 < nw; ++w2) t += red[w2][o];
            atomicAdd(colsum + o, t * cs_scale);
        }
    }
}

}  // namespace pg

using namespace pg;

#define DISPATCH_T(dtype, ...)                                       \
    do {                                                             \
        if ((dtype) == PG_BF16) {                                    \
            using T = __nv_bfloat16;                                 \
            __VA_ARGS__;                                             \
        } else if ((dtype) == PG_F16) {                              \
            using T = __half;                                        \
            __VA_ARGS__;                                             \
        } else if ((dtype) == PG_F32) {                              \
            using T = float;                                         \
            __VA_ARGS__;                                             \
        } else {                                                     \
            pg::set_error("bad dtype %d", (int)(dtype));             \
            return PG_ERR_ARG;                                       \
        }                                                            \
    } while (0)

static inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

extern "C" int pggan_lrelu_mask_mul(const void* g, const void* y, void* out, int64_t n, float slope, int dtype,
                                    void* stream) {
    PG_CHECK(g && y && out && n >= 0, PG_ERR_ARG, "pggan_lrelu_mask_mul: bad args");
    PG_CHECK(al16(g) && al16(y) && al16(out), PG_ERR_ARG, "pggan_lrelu_mask_mul: pointers must be 16-byte aligned");
    if (n == 0) return PG_OK;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    DISPATCH_T(dtype, {
        const long long nvec = n / V16<T>::N;
        PG_CUDA(pg::launch(lrelu_mask_mul_kernel<T>, ew_grid(nvec + 1, 256), 256, 0, s, (const T*)g, (const T*)y, (T*)out, nvec, n, slope));
    });
    PG_LAUNCH_CHECK();
    g_other_launches++;
    return PG_OK;
}

extern "C" int pggan_axpby(const void* a, const void* b, void* out, int64_t n, pg_coef sa, pg_coef sb, int dtype,
                           void* stream) {
    PG_CHECK(a && out && n >= 0, PG_ERR_ARG, "pggan_axpby: bad args");
    PG_CHECK(al16(a) && al16(out) && (b == nullptr || al16(b)), PG_ERR_ARG, "pggan_axpby: pointers must be 16-byte aligned");
    if (n == 0) return PG_OK;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    DISPATCH_T(dtype, {
        const long long nvec = n / V16<T>::N;
        PG_CUDA(pg::launch(axpby_kernel<T>, ew_grid(nvec + 1, 256), 256, 0, s, (const T*)a, (const T*)b, (T*)out, nvec, n, sa, sb));
    });
    PG_LAUNCH_CHECK();
    g_other_launches++;
    return PG_OK;
}

static int upsample2x_impl(const void* x, const void* mask_src, void* out, int B, int H, int W, int C, pg_coef scale,
                           float slope, int dtype, void* stream, const char* who) {
    PG_CHECK(x && out && B > 0 && H > 0 && W > 0 && C > 0, PG_ERR_ARG, "%s: bad args", who);
    PG_CHECK(al16(x) && al16(out) && (!mask_src || al16(mask_src)), PG_ERR_ARG, "%s: pointers must be 16-byte aligned", who);
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    DISPATCH_T(dtype, {
        PG_CHECK(C % V16<T>::N == 0, PG_ERR_UNSUPPORTED, "%s: C=%d must be a multiple of %d", who, C, V16<T>::N);
        const int CV = C / V16<T>::N;
        PG_CUDA(pg::launch(upsample2x_kernel<T>, ew_grid((long long)B * H * W * CV, 256), 256, 0, s, (const T*)x, (const T*)mask_src, (T*)out,
                                                                                     B, H, W, CV, scale, slope));
    });
    PG_LAUNCH_CHECK();
    g_other_launches++;
    return PG_OK;
}

extern "C" int pggan_upsample2x(const void* x, void* out, int B, int H, int W, int C, pg_coef scale, int dtype,
                                void* stream) {
    return upsample2x_impl(x, nullptr, out, B, H, W, C, scale, 1.f, dtype, stream, "pggan_upsample2x");
}

extern "C" int pggan_upsample2x_mask(const void* x, const void* mask_src, void* out, int B, int H, int W, int C,
                                     pg_coef scale, float slope, int dtype, void* stream) {
    PG_CHECK(mask_src != nullptr, PG_ERR_ARG, "pggan_upsample2x_mask: null mask_src");
    return upsample2x_impl(x, mask_src, out, B, H, W, C, scale, slope, dtype, stream, "pggan_upsample2x_mask");
}

extern "C" int pggan_upsample2x_mask_colsum(const void* x, const void* mask_src, void* out, float* colsum, int B, int H,
                                            int W, int C, pg_coef scale, float slope, float cs_scale, int dtype,
                                            void* stream) {
    PG_CHECK(x && mask_src && out && colsum && B > 0 && H > 0 && W > 0 && C > 0, PG_ERR_ARG,
             "pggan_upsample2x_mask_colsum: bad args");
    PG_CHECK(al16(x) && al16(mask_src) && al16(out), PG_ERR_ARG, "pggan_upsample2x_mask_colsum: pointers must be 16-byte aligned");
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    DISPATCH_T(dtype, {
        constexpr int VN = V16<T>::N;
        PG_CHECK(C % VN == 0 && (C / VN) <= 256 && 256 % (C / VN) == 0, PG_ERR_UNSUPPORTED,
                 "pggan_upsample2x_mask_colsum: C/%d = %d must divide 256", VN, C / VN);
        const int TX = C / VN, TY = 256 / TX;
        const long long m = (long long)B * H * W;
        long long rpb = (long long)TY * 8;
        long long blocks = (m + rpb - 1) / rpb;
        const long long cap = (long long)num_sms() * 16;
        if (blocks > cap) {
            blocks = cap;
            rpb = (m + blocks - 1) / blocks;
        }
        const int grid = (int)((m + rpb - 1) / rpb);
        PG_CUDA(pg::launch(upsample2x_mask_colsum_kernel<T>, grid, 256, 0, s, (const T*)x, (const T*)mask_src, (T*)out, colsum, B, H, W, C,
                                                              scale, slope, cs_scale, TX, rpb));
    });
    PG_LAUNCH_CHECK();
    g_other_launches++;
    return PG_OK;
}

extern "C" int pggan_pool2x2(const void* x, const void* blend_with, void* out, int B, int H, int W, int C, pg_coef scale,
                             pg_coef alpha, int dtype, void* stream) {
    PG_CHECK(x && out && B > 0 && H > 1 && W > 1 && C > 0, PG_ERR_ARG, "pggan_pool2x2: bad args");
    PG_CHECK(H % 2 == 0 && W % 2 == 0, PG_ERR_ARG, "pggan_pool2x2: H=%d W=%d must be even", H, W);
    PG_CHECK(al16(x) && al16(out) && (blend_with == nullptr || al16(blend_with)), PG_ERR_ARG,
             "pggan_pool2x2: pointers must be 16-byte aligned");
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    DISPATCH_T(dtype, {
        PG_CHECK(C % V16<T>::N == 0, PG_ERR_UNSUPPORTED, "pggan_pool2x2: C=%d must be a multiple of %d", C, V16<T>::N);
        const int CV = C / V16<T>::N;
        PG_CUDA(pg::launch(pool2x2_kernel<T>, ew_grid((long long)B * (H / 2) * (W / 2) * CV, 256), 256, 0, s, 
            (const T*)x, (const T*)blend_with, (T*)out, B, H, W, CV, scale, alpha));
    });
    PG_LAUNCH_CHECK();
    g_other_launches++;
    return PG_OK;
}

extern "C" int pggan_pixelnorm_fwd(const void* x, void* y, float* rnorm, int64_t m, int C, float eps, int dtype,
                                   void* stream) {
    PG_CHECK(x && y && m >= 0 && C > 0, PG_ERR_ARG, "pggan_pixelnorm_fwd: bad args");
    PG_CHECK(al16(x) && al16(y), PG_ERR_ARG, "pggan_pixelnorm_fwd: pointers must be 16-byte aligned");
    if (m == 0) return PG_OK;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    DISPATCH_T(dtype, {
        constexpr int VN = V16<T>::N;
        PG_CHECK(C % VN == 0 && C <= 32 * VN * 4, PG_ERR_UNSUPPORTED,
                 "pggan_pixelnorm_fwd: C=%d must be a multiple of %d and <= %d", C, VN, 32 * VN * 4);
        const int lp = C / VN;
        if (lp <= 16 && (lp & (lp - 1)) == 0) {
            PG_CUDA(pg::launch(pixelnorm_fwd_small_kernel<T>, ew_grid(m * lp, 256), 256, 0, s, (const T*)x, (T*)y, rnorm, m, C, eps, lp));
            PG_LAUNCH_CHECK();
            g_other_launches++;
            return PG_OK;
        }
        const int grid = ew_grid(m * 32, 256);
        if (C <= 32 * VN)
            PG_CUDA(pg::launch(pixelnorm_fwd_kernel<T, 1>, grid, 256, 0, s, (const T*)x, (T*)y, rnorm, m, C, eps));
        else if (C <= 64 * VN)
            PG_CUDA(pg::launch(pixelnorm_fwd_kernel<T, 2>, grid, 256, 0, s, (const T*)x, (T*)y, rnorm, m, C, eps));
        else
            PG_CUDA(pg::launch(pixelnorm_fwd_kernel<T, 4>, grid, 256, 0, s, (const T*)x, (T*)y, rnorm, m, C, eps));
    });
    PG_LAUNCH_CHECK();
    g_other_launches++;
    return PG_OK;
}

template <typename T, typename TY>
static int pixelnorm_bwd_launch(const void* gy, const void* y, const float* rnorm, void* gx, int64_t m, int C,
                                float act_slope, float* colsum, float cs_scale, cudaStream_t s) {
    constexpr int VN = V16<T>::N;
    PG_CHECK(C % VN == 0 && C <= 32 * VN * 4, PG_ERR_UNSUPPORTED,
             "pggan_pixelnorm_bwd: C=%d must be a multiple of %d and <= %d", C, VN, 32 * VN * 4);
    const int lp = C / VN;
    if (lp <= 16 && (lp & (lp - 1)) == 0) {
        PG_CUDA(pg::launch(pixelnorm_bwd_small_kernel<T, TY>, ew_grid(m * lp / 4 + 1, 256), 256, 0, s, 
            (const T*)gy, (const TY*)y, rnorm, (T*)gx, m, C, act_slope, colsum, cs_scale, lp));
    } else {
        const int grid = ew_grid(m * 32 / 4 + 1, 256);      // ~4 rows per warp
        if (C <= 32 * VN)
            PG_CUDA(pg::launch(pixelnorm_bwd_kernel<T, 1, TY>, grid, 256, 0, s, (const T*)gy, (const TY*)y, rnorm, (T*)gx, m, C, act_slope, colsum, cs_scale));
        else if (C <= 64 * VN)
            PG_CUDA(pg::launch(pixelnorm_bwd_kernel<T, 2, TY>, grid, 256, 0, s, (const T*)gy, (const TY*)y, rnorm, (T*)gx, m, C, act_slope, colsum, cs_scale));
        else
            PG_CUDA(pg::launch(pixelnorm_bwd_kernel<T, 4, TY>, grid, 256, 0, s, (const T*)gy, (const TY*)y, rnorm, (T*)gx, m, C, act_slope, colsum, cs_scale));
    }
    PG_LAUNCH_CHECK();
    g_other_launches++;
    return PG_OK;
}

extern "C" int pggan_pixelnorm_bwd(const void* gy, const void* y, const float* rnorm, void* gx, int64_t m, int C,
                                   float act_slope, float* colsum, float cs_scale, int dtype, int y_dtype, void* stream) {
    PG_CHECK(gy && y && rnorm && gx && m >= 0 && C > 0, PG_ERR_ARG, "pggan_pixelnorm_bwd: bad args");
    PG_CHECK(al16(gy) && al16(y) && al16(gx), PG_ERR_ARG, "pggan_pixelnorm_bwd: pointers must be 16-byte aligned");
    if (m == 0) return PG_OK;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    if (y_dtype != dtype) {
        // mixed precision: bf16 gradients, fp16 saved activation (or the reverse)
        PG_CHECK(is16(dtype) && is16(y_dtype), PG_ERR_ARG, "pggan_pixelnorm_bwd: dtype %d / y_dtype %d cannot be mixed", dtype, y_dtype);
        if (dtype == PG_BF16)
            return pixelnorm_bwd_launch<__nv_bfloat16, __half>(gy, y, rnorm, gx, m, C, act_slope, colsum, cs_scale, s);
        return pixelnorm_bwd_launch<__half, __nv_bfloat16>(gy, y, rnorm, gx, m, C, act_slope, colsum, cs_scale, s);
    }
    DISPATCH_T(dtype, { return pixelnorm_bwd_launch<T, T>(gy, y, rnorm, gx, m, C, act_slope, colsum, cs_scale, s); });
    return PG_OK;
}

extern "C" int pggan_mask_colsum(const void* g, const void* y, void* out, float* colsum, int64_t m, int C, float slope,
                                 float scale, int dtype, void* stream) {
    PG_CHECK(g && m >= 0 && C > 0 && (out || colsum), PG_ERR_ARG, "pggan_mask_colsum: bad args");
    PG_CHECK(al16(g) && (!y || al16(y)) && (!out || al16(out)), PG_ERR_ARG, "pggan_mask_colsum: pointers must be 16-byte aligned");
    if (m == 0) return PG_OK;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    DISPATCH_T(dtype, {
        constexpr int VN = V16<T>::N;
        PG_CHECK(C % VN == 0 && (C / VN) <= 256 && 256 % (C / VN) == 0, PG_ERR_UNSUPPORTED,
                 "pggan_mask_colsum: C/%d = %d must divide 256", VN, C / VN);
        const int TX = C / VN, TY = 256 / TX;
        long long rpb = (long long)TY * 16;
        long long blocks = (m + rpb - 1) / rpb;
        const long long cap = (long long)num_sms() * 16;
        if (blocks > cap) {
            blocks = cap;
            rpb = (m + blocks - 1) / blocks;
        }
        const int grid = (int)((m + rpb - 1) / rpb);
        if (y == nullptr && out == nullptr)
            PG_CUDA(pg::launch(colsum_only_kernel<T>, grid, 256, 0, s, (const T*)g, colsum, m, C, scale, TX, rpb));
        else
            PG_CUDA(pg::launch(mask_colsum_kernel<T>, grid, 256, 0, s, (const T*)g, (const T*)y, (T*)out, colsum, m, C, slope, scale, TX, rpb));
    });
    PG_LAUNCH_CHECK();
    g_other_launches++;
    return PG_OK;
}

extern "C" int pggan_colsum(const void* g, float* out, int64_t m, int C, float scale, int dtype, void* stream) {
    PG_CHECK(g && out && m >= 0 && C > 0, PG_ERR_ARG, "pggan_colsum: bad args");
    PG_CHECK(al16(g), PG_ERR_ARG, "pggan_colsum: pointer must be 16-byte aligned");
    if (m == 0) return PG_OK;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    DISPATCH_T(dtype, {
        constexpr int VN = V16<T>::N;
        PG_CHECK(C % VN == 0, PG_ERR_UNSUPPORTED, "pggan_colsum: C=%d must be a multiple of %d", C, VN);
        const int CV = C / VN;
        int TX = 1;
        while (TX * 2 <= CV && TX * 2 <= 256) TX *= 2;
        long long blocks = (m + 63) / 64;
        const long long cap = (long long)num_sms() * 4;
        if (blocks > cap) blocks = cap;
        const long long rows_per_block = (m + blocks - 1) / blocks;
        const int grid = (int)((m + rows_per_block - 1) / rows_per_block);
        PG_CUDA(pg::launch(colsum_kernel<T>, grid, 256, 0, s, (const T*)g, out, m, C, scale, TX, rows_per_block));
    });
    PG_LAUNCH_CHECK();
    g_other_launches++;
    return PG_OK;
}

static int make_map(PackMap& m, const pg_pack_map* p, const char* who) {
    PG_CHECK(p != nullptr, PG_ERR_ARG, "%s: null map", who);
    PG_CHECK(p->n_total > 0 && p->n_valid > 0 && p->n_valid <= p->n_total && p->n2 > 0 && p->taps > 0 && p->c_total > 0 &&
                 p->c_valid > 0 && p->c_valid <= p->c_total && p->c2 > 0,
             PG_ERR_ARG, "%s: bad map", who);
    m.n_total = p->n_total; m.n_valid = p->n_valid; m.n2 = p->n2; m.taps = p->taps;
    m.c_total = p->c_total; m.c_valid = p->c_valid; m.c2 = p->c2;
    m.base = p->base; m.s_n1 = p->s_n1; m.s_n2 = p->s_n2; m.s_t = p->s_t; m.s_c1 = p->s_c1; m.s_c2 = p->s_c2;
    return PG_OK;
}

// does every logical row of the map live in one short span of the parameter, disjoint from the other rows' spans?
// (then the row-span kernels apply).  The valid (t, c) of a row must cover the span exactly once (a permutation): true
// for the layouts ops.py builds; checked here through the element count.
static bool rowspan_ok(const PackMap& m, RowSpan& rs) {
    const long long c1_len = (m.c_valid + m.c2 - 1) / m.c2;
    const long long c2_len = m.c_valid < m.c2 ? m.c_valid : m.c2;
    if (m.c_valid > m.c2 && m.c_valid % m.c2 != 0) return false;
    long long lo = 0, hi = 0;
    const long long ext[3][2] = {{m.taps - 1, m.s_t}, {c1_len - 1, m.s_c1}, {c2_len - 1, m.s_c2}};
    for (int k = 0; k < 3; ++k) {
        const long long e = ext[k][0] * ext[k][1];
        if (e < 0) lo += e; else hi += e;
    }
    const long long span = hi - lo + 1;
    if (span > kRowSpanMax || span != (long long)m.taps * m.c_valid) return false;
    // rows must not overlap: distinct row bases at least a span apart (two-level n: one index is the major one)
    const long long a1 = m.s_n1 < 0 ? -m.s_n1 : m.s_n1, a2 = m.s_n2 < 0 ? -m.s_n2 : m.s_n2;
    const long long n1_len = (m.n_valid + m.n2 - 1) / m.n2, n2_len = m.n_valid < m.n2 ? m.n_valid : m.n2;
    if (n2_len > 1 && a2 < span) return false;
    if (n1_len > 1 && a1 < span) return false;
    if (n1_len > 1 && n2_len > 1 && !(a1 >= n2_len * a2 || a2 >= n1_len * a1)) return false;
    rs.span0 = lo;
    rs.span = (int)span;
    return true;
}

extern "C" int pggan_pack_weight(const float* in, void* out, int dtype, const pg_pack_map* map, void* stream) {
    PG_CHECK(in && out, PG_ERR_ARG, "pggan_pack_weight: null pointer");
    PackMap m;
    int rc = make_map(m, map, "pggan_pack_weight");
    if (rc != PG_OK) return rc;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    const long long total = (long long)m.n_total * m.taps * m.c_total;
    const bool oihw_f = m.taps == 9 && m.n2 == 1 && m.c2 == 1 && m.base == 0 && m.s_t == 1 && m.s_c1 == 9 &&
                        m.s_n1 == 9LL * m.c_valid && m.n_valid == m.n_total && m.c_total <= kPackMaxI;
    const bool oihw_t = m.taps == 9 && m.n2 == 1 && m.c2 == 1 && m.base == 8 && m.s_t == -1 && m.s_n1 == 9 &&
                        m.s_c1 == 9LL * m.n_valid && m.c_valid == m.c_total && m.c_total % 64 == 0;
    // channels_last parameter, dgrad operand: out[i][T-1-t][o] = W[o][t][i]
    const bool ohwi_t = m.taps > 1 && m.n2 == 1 && m.c2 == 1 && m.s_n1 == 1 && m.s_t == -(long long)m.n_valid && m.c_valid == m.c_total &&
                        m.s_c1 == (long long)m.taps * m.n_valid && m.base == (long long)(m.taps - 1) * m.n_valid;
    if (ohwi_t) {
        DISPATCH_T(dtype, {
            PG_CUDA(pg::launch(pack_ohwi_t_kernel<T>, dim3((m.n_total + 31) / 32, (m.c_total + 31) / 32, m.taps), 256, 0, s, 
                in, (T*)out, m.c_total, m.n_valid, m.n_total, m.taps));
        });
    } else if (oihw_f) {
        DISPATCH_T(dtype, {
            PG_CUDA(pg::launch(pack_oihw_f_kernel<T>, m.n_total, 256, 9 * m.c_valid * sizeof(float), s, in, (T*)out, m.c_valid, m.c_total));
        });
    } else if (oihw_t) {
        DISPATCH_T(dtype, {
            PG_CUDA(pg::launch(pack_oihw_t_kernel<T>, dim3((m.n_total + 7) / 8, m.c_total / 64), 256, 0, s, in, (T*)out, m.c_total, m.n_valid,
                                                                                          m.n_total));
        });
    } else if (RowSpan rs; rowspan_ok(m, rs) && m.n_total >= 64) {
        const size_t smem = sizeof(float) * (size_t)(rs.span + (rs.span >> 5) + 1);
        DISPATCH_T(dtype, { PG_CUDA(pg::launch(pack_rowspan_kernel<T>, m.n_total, 256, smem, s, in, (T*)out, m, rs)); });
    } else {
        DISPATCH_T(dtype, { PG_CUDA(pg::launch(pack_weight_kernel<T>, ew_grid(total, 256), 256, 0, s, in, (T*)out, m, total)); });
    }
    PG_LAUNCH_CHECK();
    g_other_launches++;
    return PG_OK;
}

extern "C" int pggan_unpack_wgrad(const float* in, float* out, const pg_pack_map* map, int slices, float scale,
                                  int accumulate, void* stream) {
    PG_CHECK(in && out && slices >= 1, PG_ERR_ARG, "pggan_unpack_wgrad: bad args");
    PackMap m;
    int rc = make_map(m, map, "pggan_unpack_wgrad");
    if (rc != PG_OK) return rc;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    const long long total = (long long)m.n_valid * m.taps * m.c_valid;
    const bool oihw_f = m.taps == 9 && m.n2 == 1 && m.c2 == 1 && m.base == 0 && m.s_t == 1 && m.s_c1 == 9 &&
                        m.s_n1 == 9LL * m.c_valid && m.n_valid == m.n_total && m.c_total <= kPackMaxI;
    const long long slice_elems = (long long)m.n_total * m.taps * m.c_total;
    // the gradient's own layout is the slice layout (channels_last conv weight, plain [out,in] linear weight)
    const bool dense = m.n2 == 1 && m.c2 == 1 && m.base == 0 && m.s_c1 == 1 && (m.taps == 1 || m.s_t == (long long)m.c_total) &&
                       m.s_n1 == (long long)m.taps * m.c_total && m.n_valid == m.n_total && m.c_valid == m.c_total &&
                       slice_elems % 4 == 0 && al16(in) && al16(out);
    if (dense) {
        PG_CUDA(pg::launch(unpack_dense_kernel, ew_grid(slice_elems / 4, 256), 256, 0, s, in, out, slice_elems / 4, scale, accumulate, slices,
                                                                          slice_elems));
    } else if (oihw_f) {
        int zchunks = 1;                 // channel chunks of >= 128 (a multiple of 4 keeps the float4 path)
        while (zchunks < 4 && m.c_valid / (zchunks * 2) >= 128 && m.c_valid % (zchunks * 2 * 4) == 0) zchunks *= 2;
        // few output rows but several slices (narrow layers: ncu showed 64 blocks summing 29 slices in 38 us): split
        // the slices over enough blocks to cover the machine twice; the slice groups meet through atomics
        int spb = slices, groups = 1;
        const int want_blocks = 2 * num_sms();
        if (slices >= 4 && m.n_total * zchunks < want_blocks) {
            groups = (want_blocks + m.n_total * zchunks - 1) / (m.n_total * zchunks);
            if (groups > slices) groups = slices;
            spb = (slices + groups - 1) / groups;
            groups = (slices + spb - 1) / spb;
            if (groups > 1 && !accumulate)
                PG_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * (size_t)m.n_total * 9 * m.c_valid, s));
        }
        PG_CUDA(pg::launch(unpack_oihw_kernel, dim3(m.n_total, groups, zchunks), 256, 9 * m.c_valid * sizeof(float), s, 
            in, out, m.c_valid, m.c_total, scale, accumulate, slices, slice_elems, spb));
    } else if (RowSpan rs; rowspan_ok(m, rs) && m.n_valid >= 64) {
        const size_t smem = sizeof(float) * (size_t)(rs.span + (rs.span >> 5) + 1);
        PG_CUDA(pg::launch(unpack_rowspan_kernel, m.n_valid, 256, smem, s, in, out, m, rs, scale, accumulate, slices, slice_elems));
    } else
        PG_CUDA(pg::launch(unpack_wgrad_kernel, ew_grid(total, 256), 256, 0, s, in, out, m, total, scale, accumulate, slices, slice_elems));
    PG_LAUNCH_CHECK();
    g_other_launches++;
    return PG_OK;
}

extern "C" int pggan_convert(const void* in, void* out, int64_t n, int from_dtype, int to_dtype, void* stream) {
    PG_CHECK(in && out && n >= 0, PG_ERR_ARG, "pggan_convert: bad args");
    PG_CHECK(al16(in) && al16(out), PG_ERR_ARG, "pggan_convert: pointers must be 16-byte aligned");
    if (n == 0) return PG_OK;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    const int grid = ew_grid(n / 8 + 1, 256);
#define PG_CVT(FI, TI, FO, TO)                                                                             \
    if (from_dtype == FI && to_dtype == FO) {                                                              \
        PG_CUDA(pg::launch(convert_kernel<TI, TO>, grid, 256, 0, s, (const TI*)in, (TO*)out, n));                           \
        PG_LAUNCH_CHECK();                                                                                 \
        g_other_launches++;                                                                                \
        return PG_OK;                                                                                      \
    }
    PG_CVT(PG_F16, __half, PG_BF16, __nv_bfloat16)
    PG_CVT(PG_BF16, __nv_bfloat16, PG_F16, __half)
    PG_CVT(PG_F32, float, PG_BF16, __nv_bfloat16)
    PG_CVT(PG_F32, float, PG_F16, __half)
    PG_CVT(PG_BF16, __nv_bfloat16, PG_F32, float)
    PG_CVT(PG_F16, __half, PG_F32, float)
#undef PG_CVT
    pg::set_error("pggan_convert: unsupported conversion %d -> %d", from_dtype, to_dtype);
    return PG_ERR_ARG;
}

namespace pg {
int split3_launch(const float* in, float* out, long long m, int c, int along_channels, int lo_slot, cudaStream_t s) {
    PG_CHECK(in && out && m > 0 && c > 0 && c % 4 == 0 && lo_slot >= 0 && lo_slot < 3, PG_ERR_ARG,
             "pggan_split3: bad args (c=%d must be a multiple of 4)", c);
    PG_CHECK(al16(in) && al16(out), PG_ERR_ARG, "pggan_split3: pointers must be 16-byte aligned");
    PG_CUDA(pg::launch(split3_kernel, ew_grid(m * (c / 4), 256), 256, 0, s, in, out, m, c / 4, along_channels, lo_slot));
    PG_LAUNCH_CHECK();
    g_other_launches++;
    return PG_OK;
}
}  // namespace pg

extern "C" int pggan_split3(const float* in, float* out, int64_t m, int c, int along_channels, int lo_slot, void* stream) {
    return pg::split3_launch(in, out, m, c, along_channels, lo_slot, reinterpret_cast<cudaStream_t>(stream));
}
