// Weight-gradient GEMM on tcgen05 tensor cores, fed by TMA (split-K over pixels).
//
//   acc[n][tap][c] += sum_p g[p][n] * x[p + shift(tap)][c]
//
// The reduction runs over pixels, which are the OUTER dimension of both NHWC operands, so both
// operands are "MN-major" for the tensor core: a stage holds BKP pixels x 128 B of channels per TMA
// box (SWIZZLE_128B), and the UMMA descriptors walk K (pixels) in 8-row groups (SBO = 1024 B) and
// MN (channels) in 128-byte spans (LBO = BKP*128 B).  The x box is shifted by the tap offset and
// TMA's out-of-bounds zero fill provides the conv padding, exactly as in conv_gemm.cu.
//
// D tile: 128 output channels (g) x BLOCK_C input channels (x), fp32 in TMEM.  Split-K partial sums go
// to per-split slices of the accumulator with plain vector stores (no atomics, no pre-zeroing);
// pggan_unpack_wgrad sums the slices while it scales/transposes into the parameter layout.
//
// Replaces cuDNN wgrad / cuBLAS behind autograd's convolution_backward for
// EqualizedConv2d/EqualizedLinear (reference lib/layers.py:60) and provides the
// d/dW term of the double backward.
#include "common.cuh"

#include <atomic>
#include <stdlib.h>

namespace pg {

extern std::atomic<long long> g_tc_launches, g_simt_launches;
extern int g_conv_backend;

// wgrad_rows.cu: narrow layers on wide images (all nine taps in one MMA over smem-resident rows)
bool wgrad_rows_supported(const pg_wgrad_params* p);
int wgrad_rows_slices(const pg_wgrad_params* p);
int wgrad_rows_launch(const pg_wgrad_params* p, cudaStream_t stream);

constexpr int kWgThreads = 192;

template <typename T, int BLOCK_C> struct WgCfg {
    static constexpr int ES = sizeof(T);
    static constexpr int UMMA_K = 32 / ES;             // pixels per MMA
    static constexpr int BKP = 4 * UMMA_K;             // pixels per stage (64 bf16 / 32 fp32)
    static constexpr int kGroupBytes = BKP * 128;      // one 128-byte channel span x BKP pixels
    static constexpr int GA = 128 * ES / 128;          // channel spans of the g tile (128 rows of D)
    static constexpr int GB = BLOCK_C * ES / 128;      // channel spans of the x tile
    static constexpr int kStageBytes = (GA + GB) * kGroupBytes;
    static constexpr int kStages = (196608 / kStageBytes) > 8 ? 8 : (196608 / kStageBytes);
    static constexpr int kTmemCols = (2 * BLOCK_C < 32) ? 32 : 2 * BLOCK_C;
    static constexpr int kSmemBytes = kStages * kStageBytes + 1024 + 256;
    static constexpr int CPS = 128 / ES;               // channels per 128-byte span
};

struct WgKArgs {
    float* acc;
    long long slice_elems;          // N*taps*C: split s writes its partial sums to acc + s*slice_elems
    float* direct;                  // != NULL: add dscale * partial sums straight into this [N][taps][C] tensor instead
    float dscale;
    int B, H, W, C, N, taps;
    int PW, PH, PB;                 // pixel box of one stage
    int chunks_w, chunks_h, chunks_b;
    int tiles_n, tiles_c, splits, chunks_per_split;
    int f16;                        // 16-bit g is IEEE half (else bf16)
    int xf16;                       // ... and x (mixed precision: g bf16, x fp16)
    int span5d;                     // 1: tensor maps are 5-D (channel spans are the 5th dim): one TMA box per operand
    int tpu;                        // taps per unit: C < BLOCK_C layers stack tpu = BLOCK_C/C shifted x tiles side by
                                    // side along the MMA N dimension (one g tile feeds several taps at full N = 256)
    int tap_groups;                 // ceil(taps / tpu)
};

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

template <typename T, int BLOCK_C>
__global__ void __launch_bounds__(kWgThreads, 1)
wgrad_gemm_kernel(const __grid_constant__ CUtensorMap tmap_g, const __grid_constant__ CUtensorMap tmap_x,
                  const WgKArgs a) {
    using Cfg = WgCfg<T, BLOCK_C>;
    constexpr bool kTF32 = (sizeof(T) == 4);
    constexpr int kStages = Cfg::kStages;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kStages * Cfg::kStageBytes);
    uint64_t* empty_bar = full_bar + kStages;
    uint64_t* tfull_bar = empty_bar + kStages;
    uint64_t* tempty_bar = tfull_bar + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_g);
        tma_prefetch_desc(&tmap_x);
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&tfull_bar[s], 1);
            mbar_init(&tempty_bar[s], 128);
        }
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc<Cfg::kTmemCols>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    PG_PDL_ENTER();        // prologue done (barriers, TMEM, descriptor prefetch): now wait for the predecessor's data

    const int total_chunks = a.chunks_w * a.chunks_h * a.chunks_b;
    const int num_units = a.tap_groups * a.tiles_n * a.tiles_c * a.splits;

    // unit -> (split, tap, n tile, c tile); c fastest so neighbouring CTAs share the g tile in L2
    auto decode = [&](int unit, int& ct, int& nt, int& tap, int& k0, int& k1, int& sp) {
        ct = unit % a.tiles_c;
        unit /= a.tiles_c;
        nt = unit % a.tiles_n;
        unit /= a.tiles_n;
        tap = unit % a.tap_groups;          // tap group (== tap when tpu == 1)
        sp = unit / a.tap_groups;
        k0 = sp * a.chunks_per_split;
        k1 = k0 + a.chunks_per_split;
        if (k1 > total_chunks) k1 = total_chunks;
    };

    if (warp == 0) {
        if (elect_one()) {    // (not lane == 0: ptxas keeps descriptors in uniform registers only under elect.sync)
            int stage = 0;
            uint32_t phase = 0;
            for (int unit = blockIdx.x; unit < num_units; unit += gridDim.x) {
                int ct, nt, tap, k0, k1, sp;
                decode(unit, ct, nt, tap, k0, k1, sp);
                int dh = 0, dw = 0;
                if (a.taps == 9 && a.tpu == 1) {
                    dh = tap / 3 - 1;
                    dw = tap % 3 - 1;
                }
                for (int kc = k0; kc < k1; ++kc) {
                    const int w0 = (kc % a.chunks_w) * a.PW;
                    const int h0 = ((kc / a.chunks_w) % a.chunks_h) * a.PH;
                    const int b0 = (kc / (a.chunks_w * a.chunks_h)) * a.PB;
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t* sa = smem + stage * Cfg::kStageBytes;
                    mbar_arrive_expect_tx(&full_bar[stage], Cfg::kStageBytes);
                    uint8_t* sb = sa + Cfg::GA * Cfg::kGroupBytes;
                    if (a.span5d && a.tpu > 1) {
                        // tap stacking: tpu shifted x boxes (each all C channels = GB/tpu spans) fill the N = BLOCK_C tile
                        tma_load_5d(sa, &tmap_g, &full_bar[stage], 0, w0, h0, b0, nt * Cfg::GA);
                        const int spans = Cfg::GB / a.tpu;
                        for (int j = 0; j < a.tpu; ++j) {
                            const int t = tap * a.tpu + j;
                            // taps past the filter (last group) are loaded fully out of bounds: zero fill
                            const int bb = (t < a.taps) ? b0 : (a.B + a.PB);
                            const int tdh = (a.taps == 9 && t < a.taps) ? t / 3 - 1 : 0;
                            const int tdw = (a.taps == 9 && t < a.taps) ? t % 3 - 1 : 0;
                            tma_load_5d(sb + j * spans * Cfg::kGroupBytes, &tmap_x, &full_bar[stage], 0, w0 + tdw, h0 + tdh, bb, 0);
                        }
                    } else if (a.span5d) {
                        // one box per operand: the channel spans are the outermost box dimension, so the box lands
                        // span-major ([span][pixel][128 B]) — the same smem image as the per-span loads below, with
                        // 2 TMA instructions per stage instead of GA+GB (the single producer thread was issue-bound)
                        tma_load_5d(sa, &tmap_g, &full_bar[stage], 0, w0, h0, b0, nt * Cfg::GA);
                        tma_load_5d(sb, &tmap_x, &full_bar[stage], 0, w0 + dw, h0 + dh, b0, ct * Cfg::GB);
                    } else {
#pragma unroll
                        for (int gidx = 0; gidx < Cfg::GA; ++gidx)
                            tma_load_4d(sa + gidx * Cfg::kGroupBytes, &tmap_g, &full_bar[stage],
                                        nt * 128 + gidx * Cfg::CPS, w0, h0, b0);
#pragma unroll
                        for (int gidx = 0; gidx < Cfg::GB; ++gidx)
                            tma_load_4d(sb + gidx * Cfg::kGroupBytes, &tmap_x, &full_bar[stage],
                                        ct * BLOCK_C + gidx * Cfg::CPS, w0 + dw, h0 + dh, b0);
                    }
                    if (++stage == kStages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {    // (not lane == 0: ptxas keeps descriptors in uniform registers only under elect.sync)
            const uint32_t idesc = make_idesc2(kTF32 ? 2 : (a.f16 ? 0 : 1), kTF32 ? 2 : (a.xf16 ? 0 : 1), 1, 1, 128, BLOCK_C);
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int unit = blockIdx.x; unit < num_units; unit += gridDim.x, ++it) {
                int ct, nt, tap, k0, k1, sp;
                decode(unit, ct, nt, tap, k0, k1, sp);
                const int acc = it & 1;
                const uint32_t acc_phase = (it >> 1) & 1;
                mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BLOCK_C;
                for (int kc = k0; kc < k1; ++kc) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + stage * Cfg::kStageBytes);
                    const uint32_t sb = sa + Cfg::GA * Cfg::kGroupBytes;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint32_t koff = k * Cfg::UMMA_K * 128;   // UMMA_K pixel rows of 128 B
                        // bf16: SWIZZLE_128B, 8-row K groups 1024 B apart; tf32 MN-major operands only exist in
                        // the 32-byte-atom flavour: SWIZZLE_128B_BASE32B, 4-row K groups 512 B apart
                        const uint64_t adesc = make_smem_desc(sa + koff, Cfg::kGroupBytes, kTF32 ? 512 : 1024, kTF32 ? 1 : 2);
                        const uint64_t bdesc = make_smem_desc(sb + koff, Cfg::kGroupBytes, kTF32 ? 512 : 1024, kTF32 ? 1 : 2);
                        umma_ss<kTF32>(d_tmem, adesc, bdesc, idesc, (kc > k0 || k > 0) ? 1u : 0u);
                    }
                    umma_commit(&empty_bar[stage]);
                    if (++stage == kStages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                umma_commit(&tfull_bar[acc]);
            }
        }
    } else {
        const int q = warp & 3;
        const int row = q * 32 + lane;
        int it = 0;
        for (int unit = blockIdx.x; unit < num_units; unit += gridDim.x, ++it) {
            int ct, nt, tap, k0, k1, sp;
            decode(unit, ct, nt, tap, k0, k1, sp);
            const int acc = it & 1;
            const uint32_t acc_phase = (it >> 1) & 1;
            const int n = nt * 128 + row;
            const bool valid = (n < a.N) && (k1 > k0);
            float* slice = a.acc + sp * a.slice_elems;
            mbar_wait(&tfull_bar[acc], acc_phase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BLOCK_C;
#pragma unroll 1
            for (int j = 0; j < BLOCK_C / 32; ++j) {
                __syncwarp();
                uint32_t r[32];
                tmem_ld32(taddr + j * 32, r);
                tmem_ld_wait();
                // accumulator column -> (tap, input channel): stacked taps own C consecutive columns each
                int t = tap, c0 = ct * BLOCK_C + j * 32;
                if (a.tpu > 1) {
                    t = tap * a.tpu + (j * 32) / a.C;
                    c0 = (j * 32) % a.C;
                }
                if (valid && t < a.taps && a.direct != nullptr) {
                    // gradient accumulation in place: the splits (and earlier backward passes) meet through fp32 reductions
                    float* dst = a.direct + (static_cast<size_t>(n) * a.taps + t) * a.C + c0;
#pragma unroll
                    for (int c = 0; c < 8; ++c)
                        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + c * 4),
                                     "f"(__uint_as_float(r[c * 4 + 0]) * a.dscale), "f"(__uint_as_float(r[c * 4 + 1]) * a.dscale),
                                     "f"(__uint_as_float(r[c * 4 + 2]) * a.dscale), "f"(__uint_as_float(r[c * 4 + 3]) * a.dscale)
                                     : "memory");
                } else if (valid && t < a.taps) {     // every (split, tap, n, c) element is owned by exactly one unit: plain stores
                    float* dst = slice + (static_cast<size_t>(n) * a.taps + t) * a.C + c0;
#pragma unroll
                    for (int c = 0; c < 8; ++c)
                        *reinterpret_cast<float4*>(dst + c * 4) =
                            make_float4(__uint_as_float(r[c * 4 + 0]), __uint_as_float(r[c * 4 + 1]),
                                        __uint_as_float(r[c * 4 + 2]), __uint_as_float(r[c * 4 + 3]));
                }
            }
            tc_fence_before();
            mbar_arrive(&tempty_bar[acc]);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        tc_fence_after();
        tmem_dealloc<Cfg::kTmemCols>(tmem_base);
    }
}

// ----------------------------------------------------------------------------- SIMT direct kernel
// 16x16 output tile per block, pixels split across blockIdx.x, partial sums via atomics.
template <typename T>
__global__ void wgrad_direct_kernel(const T* __restrict__ g, const T* __restrict__ x, float* __restrict__ acc, int B,
                                    int H, int W, int C, int N, int taps, long long pix_per_block, int gf16, int xf16) {
    PG_PDL_ENTER();
    __shared__ float gs[16][17], xs[16][17];
    // 16-bit instantiation: bf16 / IEEE half per operand are run-time flags
    auto ldv = [](const T* p, int f16) -> float {
        if constexpr (sizeof(T) == 2) return f16bits_to_f(*reinterpret_cast<const uint16_t*>(p), f16 != 0);
        else return *p;
    };
    const long long P = static_cast<long long>(B) * H * W;
    const long long p0 = blockIdx.x * pix_per_block;
    long long p1 = p0 + pix_per_block;
    if (p1 > P) p1 = P;
    const int tap = blockIdx.y;
    const int tiles_c = (C + 15) / 16;
    const int n0 = (blockIdx.z / tiles_c) * 16, c0 = (blockIdx.z % tiles_c) * 16;
    const int ty = threadIdx.x / 16, tx = threadIdx.x % 16;
    const int dh = taps == 9 ? tap / 3 - 1 : 0, dw = taps == 9 ? tap % 3 - 1 : 0;
    float sum = 0.f;
    for (long long p = p0; p < p1; p += 16) {
        const long long pp = p + ty;
        float gv = 0.f, xv = 0.f;
        if (pp < p1) {
            if (n0 + tx < N) gv = ldv(g + pp * N + n0 + tx, gf16);
            const int w = static_cast<int>(pp % W), h = static_cast<int>((pp / W) % H);
            const long long b = pp / (static_cast<long long>(W) * H);
            const int hh = h + dh, ww = w + dw;
            if (hh >= 0 && hh < H && ww >= 0 && ww < W && c0 + tx < C)
                xv = ldv(x + ((b * H + hh) * W + ww) * C + c0 + tx, xf16);
        }
        __syncthreads();
        gs[ty][tx] = gv;
        xs[ty][tx] = xv;
        __syncthreads();
#pragma unroll
        for (int i = 0; i < 16; ++i) sum += gs[i][ty] * xs[i][tx];
    }
    if (n0 + ty < N && c0 + tx < C) atomicAdd(acc + (static_cast<size_t>(n0 + ty) * taps + tap) * C + c0 + tx, sum);
}

// ------------------------------------------------------------------ SIMT kernel for tiny channel counts
// 3x3 weight gradient when both channel counts are <= 32 (the 512^2 / 1024^2 stages).  These layers are
// bandwidth-shaped: a block stages one image-row segment of g and the 3 neighbouring rows of x in
// shared memory; a thread owns one input channel c and 4 output channels n for all 9 taps (36 fp32
// accumulators, 13 shared loads per 36 FMAs) over a strided subset of the segment's pixels.
template <typename T, int TW, int NT>
__global__ void __launch_bounds__(256)
wgrad_small_kernel(const T* __restrict__ g, const T* __restrict__ x, float* __restrict__ acc, int B, int H, int W, int C,
                   int N, int gf16, int xf16) {
    PG_PDL_ENTER();
    extern __shared__ float sm[];
    auto cvt = [](const T& v, int f16) -> float {
        if constexpr (sizeof(T) == 2) return f16bits_to_f(*reinterpret_cast<const uint16_t*>(&v), f16 != 0);
        else return v;
    };
    float* sg = sm;                                   // [TW][N]
    float* sx = sm + TW * N;                          // [3][TW + 2][C]
    constexpr int VN = 16 / sizeof(T);
    const int tpl = C * (N / NT);                     // threads per pixel lane
    const int PL = blockDim.x / tpl;                  // pixel lanes
    const int lane_id = threadIdx.x / tpl;
    const int tt = threadIdx.x % tpl;
    const int c = tt % C, n0 = (tt / C) * NT;
    float a[NT][9];
#pragma unroll
    for (int i = 0; i < NT; ++i)
#pragma unroll
        for (int t = 0; t < 9; ++t) a[i][t] = 0.f;
    const int segs_w = (W + TW - 1) / TW;
    const long long nseg = static_cast<long long>(B) * H * segs_w;
    const int xrow = (TW + 2) * C;                    // floats per staged x row
    for (long long sidx = blockIdx.x; sidx < nseg; sidx += gridDim.x) {
        const int sw = static_cast<int>(sidx % segs_w);
        const int h = static_cast<int>((sidx / segs_w) % H);
        const long long b = sidx / (static_cast<long long>(segs_w) * H);
        const int w0 = sw * TW;
        const int tw = (W - w0) < TW ? (W - w0) : TW;
        __syncthreads();
        // g: tw*N contiguous elements; x: three rows of (tw+2)*C contiguous elements (NHWC), 16-byte vectors
        const T* gsrc = g + ((b * H + h) * W + w0) * N;
        for (int k = threadIdx.x; k < tw * N / VN; k += blockDim.x) {
            const uint4 raw = __ldg(reinterpret_cast<const uint4*>(gsrc) + k);
            const T* v = reinterpret_cast<const T*>(&raw);
#pragma unroll
            for (int e = 0; e < VN; ++e) sg[k * VN + e] = cvt(v[e], gf16);
        }
        const int nvx = xrow / VN;
        for (int k = threadIdx.x; k < 3 * nvx; k += blockDim.x) {
            const int r = k / nvx, kv = k - r * nvx;
            const int hh = h + r - 1;
            const int px = (kv * VN) / C;             // staged pixel index (0 = w0-1); C % VN == 0 so a vector stays in one pixel
            const int ww = w0 + px - 1;
            uint4 raw = make_uint4(0, 0, 0, 0);
            if (hh >= 0 && hh < H && ww >= 0 && ww < W && px < tw + 2)
                raw = __ldg(reinterpret_cast<const uint4*>(x + ((b * H + hh) * W + ww) * C + (kv * VN) % C));
            const T* v = reinterpret_cast<const T*>(&raw);
#pragma unroll
            for (int e = 0; e < VN; ++e) sx[r * xrow + kv * VN + e] = cvt(v[e], xf16);
        }
        __syncthreads();
        if (lane_id < PL) {
            for (int p = lane_id; p < tw; p += PL) {
                float gv[NT];
#pragma unroll
                for (int i = 0; i < NT; i += 4) {
                    const float4 t4 = *reinterpret_cast<const float4*>(sg + p * N + n0 + i);
                    gv[i] = t4.x; gv[i + 1] = t4.y; gv[i + 2] = t4.z; gv[i + 3] = t4.w;
                }
#pragma unroll
                for (int t = 0; t < 9; ++t) {
                    const float xv = sx[(t / 3) * xrow + (p + (t % 3)) * C + c];
#pragma unroll
                    for (int i = 0; i < NT; ++i) a[i][t] += gv[i] * xv;
                }
            }
        }
    }
    if (lane_id < PL) {
#pragma unroll
        for (int i = 0; i < NT; ++i)
#pragma unroll
            for (int t = 0; t < 9; ++t) atomicAdd(acc + (static_cast<size_t>(n0 + i) * 9 + t) * C + c, a[i][t]);
    }
}

static bool is_pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }

template <typename T, int BLOCK_C>
static void plan_wg(const pg_wgrad_params* p, WgKArgs& a) {
    using Cfg = WgCfg<T, BLOCK_C>;
    a.acc = p->acc;
    a.direct = p->direct;
    a.dscale = p->direct_scale;
    a.B = p->B; a.H = p->H; a.W = p->W; a.C = p->C; a.N = p->N; a.taps = p->taps;
    a.slice_elems = (long long)p->N * p->taps * p->C;
    a.PW = p->W < Cfg::BKP ? p->W : Cfg::BKP;
    a.PH = p->H < (Cfg::BKP / a.PW) ? p->H : (Cfg::BKP / a.PW);
    a.PB = Cfg::BKP / (a.PW * a.PH);
    a.chunks_w = p->W / a.PW;
    a.chunks_h = p->H / a.PH;
    a.chunks_b = (p->B + a.PB - 1) / a.PB;
    a.tiles_n = (p->N + 127) / 128;
    a.tpu = 1;
    if (p->C < BLOCK_C && BLOCK_C % p->C == 0 && p->C % Cfg::CPS == 0 && p->N % Cfg::CPS == 0) a.tpu = BLOCK_C / p->C;
    if (a.tpu > p->taps) a.tpu = p->taps;
    a.tap_groups = (p->taps + a.tpu - 1) / a.tpu;
    a.tiles_c = a.tpu > 1 ? 1 : p->C / BLOCK_C;
    const int total_chunks = a.chunks_w * a.chunks_h * a.chunks_b;
    const int base_units = a.tap_groups * a.tiles_n * a.tiles_c;
    // split-K over pixels: pick the split count whose unit count fills whole waves of SMs best
    // (fewest slices among equally good fills), keeping >= 8 chunks per unit
    int max_splits = total_chunks / 8;
    if (max_splits < 1) max_splits = 1;
    if (max_splits > 128) max_splits = 128;
    // every split writes its own fp32 slice: keep the slices within 64 MB
    const long long slice_bytes = a.slice_elems * 4;
    while (max_splits > 1 && (long long)max_splits * slice_bytes > (64LL << 20)) --max_splits;
    int splits = 1;
    double best = -1.0;
    const int sms = num_sms();
    for (int sct = 1; sct <= max_splits; ++sct) {
        const long long units = (long long)base_units * sct;
        const long long waves = (units + sms - 1) / sms;
        double eff = (double)units / (double)(waves * sms);
        if (eff > best + 0.03) {
            best = eff;
            splits = sct;
        }
    }
    a.chunks_per_split = (total_chunks + splits - 1) / splits;
    a.splits = (total_chunks + a.chunks_per_split - 1) / a.chunks_per_split;
}

template <typename T, int BLOCK_C>
static int launch_wg(const pg_wgrad_params* p, cudaStream_t stream) {
    using Cfg = WgCfg<T, BLOCK_C>;
    WgKArgs a;
    plan_wg<T, BLOCK_C>(p, a);
    const int base_units = a.tap_groups * a.tiles_n * a.tiles_c;
    const uint64_t es = sizeof(T);
    CUtensorMap tg, tx;
    const int sw = sizeof(T) == 4 ? 12832 : 128;
    a.f16 = p->dtype == PG_F16;
    a.xf16 = p->dtype == PG_F16;
    const int xdt = sizeof(T) == 4 ? PG_F32 : p->dtype;
    a.span5d = (p->N % Cfg::CPS == 0 && p->C % Cfg::CPS == 0) ? 1 : 0;
    if (a.span5d) {
        // dims: (channel within span, w, h, b, span); the span stride is one 128-byte span
        const uint64_t span_bytes = 128;
        {
            uint64_t dims[5] = {(uint64_t)Cfg::CPS, (uint64_t)p->W, (uint64_t)p->H, (uint64_t)p->B, (uint64_t)(p->N / Cfg::CPS)};
            uint64_t strides[4] = {(uint64_t)p->N * es, (uint64_t)p->W * p->N * es, (uint64_t)p->H * p->W * p->N * es, span_bytes};
            uint32_t box[5] = {(uint32_t)Cfg::CPS, (uint32_t)a.PW, (uint32_t)a.PH, (uint32_t)a.PB, (uint32_t)Cfg::GA};
            int rc = encode_tmap(&tg, p->dtype == PG_F16 ? PG_F16 : DT<T>::kEnum, 5, p->g, dims, strides, box, sw);
            if (rc != PG_OK) return rc;
        }
        {
            uint64_t dims[5] = {(uint64_t)Cfg::CPS, (uint64_t)p->W, (uint64_t)p->H, (uint64_t)p->B, (uint64_t)(p->C / Cfg::CPS)};
            uint64_t strides[4] = {(uint64_t)p->C * es, (uint64_t)p->W * p->C * es, (uint64_t)p->H * p->W * p->C * es, span_bytes};
            uint32_t box[5] = {(uint32_t)Cfg::CPS, (uint32_t)a.PW, (uint32_t)a.PH, (uint32_t)a.PB,
                               (uint32_t)(a.tpu > 1 ? Cfg::GB / a.tpu : Cfg::GB)};
            int rc = encode_tmap(&tx, xdt, 5, p->x, dims, strides, box, sw);
            if (rc != PG_OK) return rc;
        }
    } else {
        uint32_t box[4] = {(uint32_t)Cfg::CPS, (uint32_t)a.PW, (uint32_t)a.PH, (uint32_t)a.PB};
        {
            uint64_t dims[4] = {(uint64_t)p->N, (uint64_t)p->W, (uint64_t)p->H, (uint64_t)p->B};
            uint64_t strides[3] = {(uint64_t)p->N * es, (uint64_t)p->W * p->N * es, (uint64_t)p->H * p->W * p->N * es};
            int rc = encode_tmap(&tg, p->dtype == PG_F16 ? PG_F16 : DT<T>::kEnum, 4, p->g, dims, strides, box, sw);
            if (rc != PG_OK) return rc;
        }
        {
            uint64_t dims[4] = {(uint64_t)p->C, (uint64_t)p->W, (uint64_t)p->H, (uint64_t)p->B};
            uint64_t strides[3] = {(uint64_t)p->C * es, (uint64_t)p->W * p->C * es, (uint64_t)p->H * p->W * p->C * es};
            int rc = encode_tmap(&tx, xdt, 4, p->x, dims, strides, box, sw);
            if (rc != PG_OK) return rc;
        }
    }
    auto kern = wgrad_gemm_kernel<T, BLOCK_C>;
    PG_SET_SMEM_ONCE(kern, Cfg::kSmemBytes);
    const int num_units = base_units * a.splits;
    const int grid = num_units < num_sms() ? num_units : num_sms();
    PG_CUDA(pg::launch_tc(kern, grid, kWgThreads, Cfg::kSmemBytes, stream, tg, tx, a));
    PG_LAUNCH_CHECK();
    g_tc_launches++;
    return PG_OK;
}

template <typename T, int BLOCK_C>
static int slices_wg(const pg_wgrad_params* p) {
    WgKArgs a;
    plan_wg<T, BLOCK_C>(p, a);
    return a.splits;
}

// SIMT kernels do arithmetic on the element values: the 16-bit instantiation takes the element formats of g and x
// (bf16 / IEEE half, possibly different: mixed precision) as run-time flags
template <typename T>
static int wgrad_simt(const pg_wgrad_params* p, cudaStream_t stream) {
    const int gf16 = p->dtype == PG_F16, xf16 = p->dtype == PG_F16;
    // SIMT path: one slice, partial sums meet through atomics on a zeroed accumulator
    PG_CUDA(cudaMemsetAsync(p->acc, 0, sizeof(float) * (size_t)p->N * p->taps * p->C, stream));
    constexpr int VNT = 16 / sizeof(T);
    if (p->taps == 9 && p->N <= 64 && p->C <= 32 && p->N % 8 == 0 && p->C % VNT == 0 && p->N % VNT == 0 &&
        (reinterpret_cast<uintptr_t>(p->g) % 16 == 0) && (reinterpret_cast<uintptr_t>(p->x) % 16 == 0)) {
        constexpr int TW = 64;
        const size_t smem = (size_t)(TW * p->N + 3 * (TW + 2) * p->C) * sizeof(float);
        const long long nseg = (long long)p->B * p->H * ((p->W + TW - 1) / TW);
        long long grid = (long long)num_sms() * 4;
        if (grid > nseg) grid = nseg;
        const T* gp = reinterpret_cast<const T*>(p->g);
        const T* xp = reinterpret_cast<const T*>(p->x);
        if (p->C * (p->N / 4) <= 256)
            PG_CUDA(pg::launch(wgrad_small_kernel<T, TW, 4>, (int)grid, 256, smem, stream, gp, xp, p->acc, p->B, p->H, p->W, p->C, p->N, gf16, xf16));
        else
            PG_CUDA(pg::launch(wgrad_small_kernel<T, TW, 8>, (int)grid, 256, smem, stream, gp, xp, p->acc, p->B, p->H, p->W, p->C, p->N, gf16, xf16));
        PG_LAUNCH_CHECK();
        g_simt_launches++;
        return PG_OK;
    }
    const long long P = (long long)p->B * p->H * p->W;
    long long pix_per_block = 256;
    while ((P + pix_per_block - 1) / pix_per_block > 2048) pix_per_block *= 2;
    dim3 grid((unsigned)((P + pix_per_block - 1) / pix_per_block), (unsigned)p->taps,
              (unsigned)(((p->N + 15) / 16) * ((p->C + 15) / 16)));
    PG_CUDA(pg::launch(wgrad_direct_kernel<T>, grid, 256, 0, stream, reinterpret_cast<const T*>(p->g), reinterpret_cast<const T*>(p->x),
                                                      p->acc, p->B, p->H, p->W, p->C, p->N, p->taps, pix_per_block, gf16, xf16));
    PG_LAUNCH_CHECK();
    g_simt_launches++;
    return PG_OK;
}

// slices_out != nullptr: only report how many partial-sum slices the launch would write (no launch)
template <typename T>
static int wgrad_dispatch(const pg_wgrad_params* p_in, cudaStream_t stream, int* slices_out) {
    constexpr int CPS = 128 / sizeof(T);
    constexpr bool kF32 = sizeof(T) == 4;
    pg_wgrad_params local = *p_in;
    const pg_wgrad_params* p = &local;
    const bool tc_ok = g_conv_backend == PG_BACKEND_AUTO && is_pow2(p->H) && is_pow2(p->W) && (p->C % CPS == 0) &&
                       (p->N % 8 == 0) && (!kF32 || slices_out || (p->ws_g && p->ws_x)) &&
                       (reinterpret_cast<uintptr_t>(p->g) % 16 == 0) &&
                       (reinterpret_cast<uintptr_t>(p->x) % 16 == 0) && (reinterpret_cast<uintptr_t>(p->acc) % 16 == 0) &&
                       (reinterpret_cast<uintptr_t>(p->direct) % 16 == 0);
    if constexpr (!kF32) {
        static const bool no_rows = getenv("PGGAN_NO_WROWS") != nullptr;
        const bool rows_ok = g_conv_backend == PG_BACKEND_AUTO && (reinterpret_cast<uintptr_t>(p->g) % 16 == 0) &&
                             (reinterpret_cast<uintptr_t>(p->x) % 16 == 0) && (reinterpret_cast<uintptr_t>(p->acc) % 16 == 0) &&
                             (reinterpret_cast<uintptr_t>(p->direct) % 16 == 0);
        if (rows_ok && !no_rows && wgrad_rows_supported(p)) {
            if (slices_out) {
                *slices_out = wgrad_rows_slices(p);
                return PG_OK;
            }
            return wgrad_rows_launch(p, stream);
        }
    }
    if (tc_ok && slices_out) {
        if (kF32) local.B = 3 * p->B;
        const bool stack = p->taps > 1 && p->C < 256 && 256 % p->C == 0 && p->C % CPS == 0 && p->N % CPS == 0;
        if (p->C % 256 == 0 || stack) *slices_out = slices_wg<T, 256>(p);
        else if (p->C % 128 == 0) *slices_out = slices_wg<T, 128>(p);
        else if (p->C % 64 == 0) *slices_out = slices_wg<T, 64>(p);
        else *slices_out = slices_wg<T, 32>(p);
        return PG_OK;
    }
    if (slices_out) {
        *slices_out = 1;
        return PG_OK;
    }
    if (tc_ok) {
        if (kF32) {
            // 3xTF32 along the pixel reduction: (g, g_lo, g) and (x, x, x_lo) stacked along the batch
            const long long M = (long long)p->B * p->H * p->W;
            int rc = split3_launch(reinterpret_cast<const float*>(p->g), reinterpret_cast<float*>(p->ws_g), M, p->N, 0, 1, stream);
            if (rc != PG_OK) return rc;
            rc = split3_launch(reinterpret_cast<const float*>(p->x), reinterpret_cast<float*>(p->ws_x), M, p->C, 0, 2, stream);
            if (rc != PG_OK) return rc;
            local.g = p->ws_g;
            local.x = p->ws_x;
            local.B = 3 * p->B;
        }
        const bool stack = p->taps > 1 && p->C < 256 && 256 % p->C == 0 && p->C % CPS == 0 && p->N % CPS == 0;
        if (p->C % 256 == 0 || stack) return launch_wg<T, 256>(p, stream);
        if (p->C % 128 == 0) return launch_wg<T, 128>(p, stream);
        if (p->C % 64 == 0) return launch_wg<T, 64>(p, stream);
        if constexpr (sizeof(T) == 4) {
            if (p->C % 32 == 0) return launch_wg<T, 32>(p, stream);
        }
    }
    PG_CHECK(p->direct == nullptr, PG_ERR_UNSUPPORTED,
             "pggan_conv_wgrad: direct accumulation is not available for this shape (ask pggan_conv_wgrad_direct_ok)");
    return wgrad_simt<T>(p, stream);
}

}  // namespace pg

extern "C" int pggan_conv_wgrad(const pg_wgrad_params* p, void* stream) {
    using namespace pg;
    PG_CHECK(p != nullptr && p->g && p->x && (p->acc || p->direct), PG_ERR_ARG, "pggan_conv_wgrad: null pointer");
    PG_CHECK(p->B > 0 && p->H > 0 && p->W > 0 && p->C > 0 && p->N > 0, PG_ERR_ARG,
             "pggan_conv_wgrad: bad dims B=%d H=%d W=%d C=%d N=%d", p->B, p->H, p->W, p->C, p->N);
    PG_CHECK(p->taps == 1 || p->taps == 9, PG_ERR_ARG, "pggan_conv_wgrad: taps must be 1 or 9 (got %d)", p->taps);
    PG_CHECK(p->dtype == PG_BF16 || p->dtype == PG_F16 || p->dtype == PG_F32, PG_ERR_ARG, "pggan_conv_wgrad: bad dtype %d", p->dtype);
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    if (is16(p->dtype)) return wgrad_dispatch<__nv_bfloat16>(p, s, nullptr);
    return wgrad_dispatch<float>(p, s, nullptr);
}

extern "C" int pggan_conv_wgrad_slices(const pg_wgrad_params* p) {
    using namespace pg;
    if (p == nullptr || p->B <= 0 || p->H <= 0 || p->W <= 0 || p->C <= 0 || p->N <= 0) return 1;
    int slices = 1;
    pg_wgrad_params q = *p;
    if (!q.g) q.g = reinterpret_cast<const void*>(uintptr_t(256));      // alignment checks only
    if (!q.x) q.x = reinterpret_cast<const void*>(uintptr_t(256));
    if (!q.acc) q.acc = reinterpret_cast<float*>(uintptr_t(256));
    if (is16(q.dtype))
        wgrad_dispatch<__nv_bfloat16>(&q, nullptr, &slices);
    else if (q.dtype == PG_F32)
        wgrad_dispatch<float>(&q, nullptr, &slices);
    return slices;
}

extern "C" int pggan_conv_wgrad_direct_ok(const pg_wgrad_params* p) {
    using namespace pg;
    if (p == nullptr || p->B <= 0 || p->H <= 0 || p->W <= 0 || p->C <= 0 || p->N <= 0 || g_conv_backend != PG_BACKEND_AUTO) return 0;
    if (reinterpret_cast<uintptr_t>(p->direct) % 16 != 0 || reinterpret_cast<uintptr_t>(p->g) % 16 != 0 ||
        reinterpret_cast<uintptr_t>(p->x) % 16 != 0)
        return 0;
    if (p->taps != 1 && p->taps != 9) return 0;
    if (p->dtype == PG_F32) {
        return is_pow2(p->H) && is_pow2(p->W) && p->C % 32 == 0 && p->N % 8 == 0;
    }
    if (!is16(p->dtype)) return 0;
    if (wgrad_rows_supported(p)) return 1;
    return is_pow2(p->H) && is_pow2(p->W) && p->C % 64 == 0 && p->N % 8 == 0;
}
