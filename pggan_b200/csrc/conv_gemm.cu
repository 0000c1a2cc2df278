// Implicit-GEMM 3x3 / 1x1 convolution and linear layers on tcgen05 tensor cores, fed by TMA.
//
//   D[128 pixels x BLOCK_N] (fp32, TMEM) += A[128 pixels x 128 B of K] * B[BLOCK_N x 128 B of K]^T
//
// A tile  = a (BW x BH x BB) box of pixels of the NHWC input, one filter tap, one 128-byte channel
//           chunk: a single 4-D TMA box load whose (w,h) origin is shifted by the tap offset; TMA
//           zero-fills out-of-bounds coordinates, which IS the conv zero padding (no im2col buffer).
// B tile  = BLOCK_N rows of the K-major packed weights [N][taps*C].
// Both land in the canonical SWIZZLE_128B K-major layout that tcgen05.mma consumes directly.
//
// Warp roles of the plain kernel (352 threads, 1 CTA/SM, persistent over tiles):
//   warp 0     : TMA producer        (one elected lane, elect.sync)
//   warp 1     : TMEM allocator + MMA issuer of M tile 0 (one elected lane)
//   warps 2-9  : epilogue, two groups of four warps taking alternate tiles — tcgen05.ld the accumulator
//                (double-buffered in TMEM so the next tile's mainloop overlaps), apply
//                scale*(acc+bias) -> LeakyReLU -> optional pixel norm / dgrad mask, store NHWC.
//   warp 10    : MMA issuer of M tile 1 when the CTA tile is 256 x N (MT = 2)
// The swapped-orientation kernel further down runs 192 threads (TMA warp, MMA warp, four epilogue warps).
//
// Replaces cuDNN fprop/dgrad behind EqualizedConv2d/EqualizedLinear (reference lib/layers.py:58-63)
// with the He scale, bias and LeakyReLU (lib/blocks.py:127,137,185,192) folded into the epilogue.
#include "common.cuh"

#include <atomic>
#include <stdlib.h>

namespace pg {

extern std::atomic<long long> g_tc_launches, g_simt_launches;
extern int g_conv_backend;

// conv_rows.cu: smem-resident input rows, nine shifted views (narrow layers on wide images)
bool conv_rows_supported(const pg_conv_params* p);
int conv_rows_launch(const pg_conv_params* p, cudaStream_t stream);

constexpr int kBlockM = 128;
constexpr int kConvThreads = 192;          // swapped kernel: TMA warp, MMA warp, 4 epilogue warps
constexpr int kConvThreadsPlain = 352;     // plain kernel: 2 x 4 epilogue warps taking alternate tiles + a second MMA warp

// SPAN = bytes of K per chunk = the TMA/UMMA swizzle span: 128 normally; 64 / 32 for layers whose channel
// count times element size is only 64 / 32 bytes (the 32- and 16-channel stages of the 1024^2 network)
// MT = 128-pixel M tiles per CTA tile.  MT = 2 with BLOCK_N = 128 gives a 256 x 128 tile: two MMAs per K-step share the
// weight tile, so the L2->smem bytes per FLOP equal those of the 128 x 256 tile (a 128 x 128 tile moves 1.33x more and
// measured 1 094 vs 1 525 TFLOP/s on the 32^2 layer); used for the big N = 128 layers of the 128^2 / 256^2 stages.
template <int BLOCK_N, int SPAN = 128, int MT = 1> struct ConvCfg {
    static constexpr int kABytesS = MT * kBlockM * SPAN;
    static constexpr int kBBytes = BLOCK_N * SPAN;
    static constexpr int kStageBytes = kABytesS + kBBytes;
    static constexpr int kStagesRaw = 196608 / kStageBytes;
    static constexpr int kStagesWant = SPAN == 128 ? ((MT * BLOCK_N >= 256) ? 4 : (BLOCK_N >= 128 ? 6 : 8)) : 8;
    static constexpr int kStages = kStagesRaw < kStagesWant ? kStagesRaw : kStagesWant;
    static constexpr int kTmemCols = (2 * MT * BLOCK_N < 32) ? 32 : 2 * MT * BLOCK_N;
    static_assert(2 * MT * BLOCK_N <= 512, "TMEM");
    static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/;
};

struct ConvKArgs {
    const float* bias;
    const void* mask_src;
    void* y;
    int B, H, W, N;
    int BW, BH, BB;
    int tiles_w, tiles_h, tiles_b, tiles_n;
    int kc_per_tap, taps;
    int out_dtype;
    int f16;                 // 16-bit x / y are IEEE half (else bf16); see common.cuh fmt16()
    int wf16;                // ... and the packed weights (may differ from f16: mixed precision, pg_conv_params.w_fmt)
    float scale, slope, mask_slope;
    float* pn_rnorm;         // fused pixel norm (plain kernel with tiles_n == 1 only): per-pixel rsqrt factor out
    float pn_eps;
    int up2;                 // fused nearest-2x up-sampling: the tile geometry (BW/BH/BB, tiles_*) is in LOW-RES pixels,
                             // every box yields four tiles (one per output-pixel parity); H, W are the output dims
    // split-K (plain kernel, MT = 1): a layer with few output tiles and a long K (the 4^2 / 8^2 trunk layers: each CTA
    // would pull its whole K extent through ONE SM's L2 port — ~0.2-0.35 us per k-step whatever the tile,
    // tools/splitk_micro.py) is cut into ksplit K ranges per output tile.  CTA ks adds its fp32 partial accumulator into
    // pair buffer ks/2 of the tile in `ws_part` ([tile][pair][column][row], zero between launches) with red.global: a
    // buffer receives at most two addends, so its content does not depend on their order.  The CTA that arrives last at
    // the tile's counter adds the pair buffers in pair order, clears them and runs the epilogue: results are
    // reproducible bit for bit.  Counters return to zero.
    int ksplit;
    float* ws_part;
    unsigned int* ws_count;
};

template <typename T, int BLOCK_N, int SPAN = 128, int MT = 1>
__global__ void __launch_bounds__(kConvThreadsPlain, 1)
conv_gemm_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
                 const ConvKArgs a) {
    using Cfg = ConvCfg<BLOCK_N, SPAN, MT>;
    constexpr bool kTF32 = (sizeof(T) == 4);
    constexpr int BK = SPAN / sizeof(T);          // elements per K chunk (one swizzle span)
    constexpr int kMmaPerChunk = SPAN / 32;       // 32 B per UMMA K-step
    constexpr int kStages = Cfg::kStages;
    constexpr int kABytes = Cfg::kABytesS;        // shadows the 128-byte-span constant
    constexpr uint32_t kLayout = SPAN == 128 ? 2u : (SPAN == 64 ? 4u : 6u);     // SWIZZLE_128B / 64B / 32B
    constexpr uint32_t kSBO = 8 * SPAN;           // 8-row core-matrix groups

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
        tma_prefetch_desc(&tmap_x);
        tma_prefetch_desc(&tmap_w);
        // MT = 2: each 128-pixel M tile has its own MMA warp (one issuing thread sustains an MMA per ~62 cycles, two
        // reach ~48 at N = 64: tools/probe_issue.cu), so a smem stage / an accumulator is released by MT commits
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], MT);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&tfull_bar[s], MT);
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

    const int num_m_tiles = a.tiles_w * a.tiles_h * a.tiles_b * (a.up2 ? 4 : 1);
    const int num_tiles = num_m_tiles * a.tiles_n * a.ksplit;      // a tile index = (output tile, K split), split fastest
    const int num_kc = a.taps * a.kc_per_tap;
    const int kc_per_split = num_kc / a.ksplit;

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer
        if (elect_one()) {    // (not lane == 0: ptxas keeps descriptors in uniform registers only under elect.sync)
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int ks = tile % a.ksplit, mn = tile / a.ksplit;
                const int nt = mn % a.tiles_n;
                int mt = mn / a.tiles_n;
                int pi = 0, pj = 0;                    // output-pixel parity of this tile (fused up-sampling)
                if (a.up2) {
                    pj = mt & 1;
                    pi = (mt >> 1) & 1;
                    mt >>= 2;
                }
                const int w0 = (mt % a.tiles_w) * a.BW;
                const int h0 = ((mt / a.tiles_w) % a.tiles_h) * a.BH;
                const int b0 = (mt / (a.tiles_w * a.tiles_h)) * a.BB;
                const int n0 = nt * BLOCK_N;
                for (int kc = ks * kc_per_split; kc < (ks + 1) * kc_per_split; ++kc) {
                    const int tap = kc / a.kc_per_tap;
                    const int cc = kc - tap * a.kc_per_tap;
                    int dh = 0, dw = 0;
                    if (a.taps == 9) {
                        dh = tap / 3 - 1;
                        dw = tap % 3 - 1;
                    }
                    if (a.up2) {                       // low-res source offset: floor((parity + d) / 2) in {-1, 0, 1}
                        dh = (pi + dh + 2) / 2 - 1;
                        dw = (pj + dw + 2) / 2 - 1;
                    }
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t* sa = smem + stage * Cfg::kStageBytes;
                    mbar_arrive_expect_tx(&full_bar[stage], Cfg::kStageBytes);
                    tma_load_4d(sa, &tmap_x, &full_bar[stage], cc * BK, w0 + dw, h0 + dh, b0);
                    tma_load_2d(sa + kABytes, &tmap_w, &full_bar[stage], kc * BK, n0);
                    if (++stage == kStages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp == 1 || (warp == 10 && MT == 2)) {
        // -------------------------------------------------------------------- MMA issuer(s): M tile `mi`
        const int mi = warp == 1 ? 0 : 1;
        if (elect_one()) {    // (not lane == 0: ptxas keeps descriptors in uniform registers only under elect.sync)
            const uint32_t idesc = make_idesc2(kTF32 ? 2 : (a.f16 ? 0 : 1), kTF32 ? 2 : (a.wf16 ? 0 : 1), 0, 0, kBlockM, BLOCK_N);
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
                const int acc = it & 1;
                const uint32_t acc_phase = (it >> 1) & 1;
                mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * (MT * BLOCK_N);
                for (int kc = 0; kc < kc_per_split; ++kc) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + stage * Cfg::kStageBytes);
                    const uint64_t adesc = make_smem_desc(sa, 16, kSBO, kLayout);
                    const uint64_t bdesc = make_smem_desc(sa + kABytes, 16, kSBO, kLayout);
#pragma unroll
                    for (int k = 0; k < kMmaPerChunk; ++k) {
                        // advance 32 bytes along K inside the swizzle span: +2 in 16-byte units; M tile m starts
                        // 128 rows further down the A box and owns its own BLOCK_N accumulator columns
                        umma_ss<kTF32>(d_tmem + mi * BLOCK_N, adesc + 2 * k + mi * ((kBlockM * SPAN) >> 4), bdesc + 2 * k,
                                       idesc, (kc | k) != 0);
                    }
                    umma_commit(&empty_bar[stage]);     // frees the smem slot when these MMAs retire
                    if (++stage == kStages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                umma_commit(&tfull_bar[acc]);           // accumulator complete -> epilogue
            }
        }
    } else if (warp >= 2 && warp < 10) {
        // ---------------------------------------------------------------------- epilogue
        const int q = warp & 3;                        // TMEM lane quarter this warp may read
        // one epilogue warp per SM sub-partition is latency-bound on short-K layers (the epilogue of a 128 x N tile
        // is ~N/32 x 450 dependent instructions): two groups of four warps drain the two TMEM accumulators concurrently
        const int egroup = (warp - 2) >> 2;
        int it = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
            if ((it & 1) != egroup) continue;
            const int acc = it & 1;
            const uint32_t acc_phase = (it >> 1) & 1;
            const int ks = tile % a.ksplit, mn = tile / a.ksplit;
            const int nt = mn % a.tiles_n;
            int mt = mn / a.tiles_n;
            int pi = 0, pj = 0;
            if (a.up2) {
                pj = mt & 1;
                pi = (mt >> 1) & 1;
                mt >>= 2;
            }
            const int w0 = (mt % a.tiles_w) * a.BW;
            const int h0 = ((mt / a.tiles_w) % a.tiles_h) * a.BH;
            const int b0 = (mt / (a.tiles_w * a.tiles_h)) * a.BB;
            const int n0 = nt * BLOCK_N;
#pragma unroll 1
            for (int m = 0; m < MT; ++m) {                   // this thread's row in each 128-pixel M tile of the box
            const int row = m * kBlockM + q * 32 + lane;
            const int wo = row % a.BW;
            const int ho = (row / a.BW) % a.BH;
            const int bo = row / (a.BW * a.BH);
            const int b = b0 + bo;
            const bool valid = b < a.B;
            // fused up-sampling: the box is in low-res pixels, this tile owns the output pixels of parity (pi, pj)
            const int hout = a.up2 ? 2 * (h0 + ho) + pi : (h0 + ho);
            const int wout = a.up2 ? 2 * (w0 + wo) + pj : (w0 + wo);
            const size_t pix = (static_cast<size_t>(b) * a.H + hout) * a.W + wout;
            const size_t off = pix * a.N + n0;

            constexpr int CH = BLOCK_N >= 32 ? 32 : 16;      // accumulator columns handled per step
            constexpr int VN = 16 / sizeof(T);
            constexpr int MV = CH / VN;                      // 16-byte mask vectors per step
            // the dgrad mask (a saved activation, same layout as y) is fetched one step AHEAD of its use — the first
            // step's vectors before the wait on the accumulator — so its global-load latency hides behind the MMAs
            const bool masked = valid && a.mask_src != nullptr;
            uint4 mcur[MV];
            if (masked) {
                const uint4* mp = reinterpret_cast<const uint4*>(reinterpret_cast<const T*>(a.mask_src) + off);
#pragma unroll
                for (int c = 0; c < MV; ++c) mcur[c] = __ldg(mp + c);
            }
            mbar_wait(&tfull_bar[acc], acc_phase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * (MT * BLOCK_N) + m * BLOCK_N;
            // split-K: add this CTA's partial accumulator to the tile's fp32 sum in the workspace (red.global: fire and
            // forget, the L2 does the adds), count the arrival; only the last CTA of the tile goes on, reading the sum
            // where the code below reads the accumulator and leaving zeros behind for the next launch
            float* part0 = nullptr;
            if (a.ksplit > 1) {
                const int npairs = (a.ksplit + 1) >> 1;
                float* part = a.ws_part + (static_cast<size_t>(mn) * npairs + (ks >> 1)) * (kBlockM * BLOCK_N) + q * 32 + lane;
#pragma unroll 1
                for (int j = 0; j < BLOCK_N / CH; ++j) {
                    __syncwarp();
                    uint32_t r[32];
                    if constexpr (CH == 32) {
                        tmem_ld32(taddr + j * 32, r);
                    } else {
                        tmem_ld16(taddr + j * 16, r);
                    }
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < CH; ++i)
                        asm volatile("red.global.add.f32 [%0], %1;" ::"l"(part + static_cast<size_t>(j * CH + i) * kBlockM),
                                     "f"(__uint_as_float(r[i]))
                                     : "memory");
                }
                __threadfence();
                uint32_t* flag = tmem_slot + 1 + egroup;
                asm volatile("bar.sync %0, 128;" ::"r"(1 + egroup) : "memory");
                if (q == 0 && lane == 0) {
                    const unsigned int prev = atomicAdd(a.ws_count + mn, 1u);
                    if (prev == static_cast<unsigned int>(a.ksplit - 1)) a.ws_count[mn] = 0u;     // ready for the next launch
                    *flag = prev;
                }
                asm volatile("bar.sync %0, 128;" ::"r"(1 + egroup) : "memory");
                const bool last = *flag == static_cast<unsigned int>(a.ksplit - 1);
                asm volatile("bar.sync %0, 128;" ::"r"(1 + egroup) : "memory");          // flag may be rewritten for the next tile
                if (!last) continue;              // (MT == 1: leaves the m loop; the accumulator is released below)
                __threadfence();
                part0 = a.ws_part + static_cast<size_t>(mn) * npairs * (kBlockM * BLOCK_N) + q * 32 + lane;
            }
            // accumulator chunk j -> r: from TMEM, or (split-K, last CTA) the pair sums of the workspace added in pair order
            auto load_acc = [&](int j, uint32_t (&r)[32], bool clear) {
                if (part0 == nullptr) {
                    tmem_ld_wait();
                    return;
                }
                const int npairs = (a.ksplit + 1) >> 1;
                constexpr size_t kTile = static_cast<size_t>(kBlockM) * BLOCK_N;
                float t[CH];
#pragma unroll
                for (int i = 0; i < CH; ++i) t[i] = 0.f;
#pragma unroll 1
                for (int pr = 0; pr < npairs; pr += 2) {          // two pair buffers in flight per round
                    float* pa = part0 + pr * kTile + static_cast<size_t>(j * CH) * kBlockM;
                    const bool two = pr + 1 < npairs;
                    float ua[CH], ub[CH];
#pragma unroll
                    for (int i = 0; i < CH; ++i) ua[i] = __ldcg(pa + static_cast<size_t>(i) * kBlockM);
#pragma unroll
                    for (int i = 0; i < CH; ++i) ub[i] = two ? __ldcg(pa + kTile + static_cast<size_t>(i) * kBlockM) : 0.f;
#pragma unroll
                    for (int i = 0; i < CH; ++i) t[i] = (t[i] + ua[i]) + ub[i];
                    if (clear) {
#pragma unroll
                        for (int i = 0; i < CH; ++i) __stcg(pa + static_cast<size_t>(i) * kBlockM, 0.f);
                        if (two) {
#pragma unroll
                            for (int i = 0; i < CH; ++i) __stcg(pa + kTile + static_cast<size_t>(i) * kBlockM, 0.f);
                        }
                    }
                }
#pragma unroll
                for (int i = 0; i < CH; ++i) r[i] = __float_as_uint(t[i]);
            };
            // fused pixel norm (one N tile covers the layer: this thread owns all channels of its pixel): a first pass
            // over the accumulator for mean(lrelu(..)^2), then the normal pass multiplies by the factor
            float rn = 1.f;
            if (a.pn_rnorm != nullptr) {
                float ss = 0.f;
#pragma unroll 1
                for (int j = 0; j < BLOCK_N / CH; ++j) {
                    __syncwarp();
                    uint32_t r[32];
                    if (part0 == nullptr) {
                        if constexpr (CH == 32) {
                            tmem_ld32(taddr + j * 32, r);
                        } else {
                            tmem_ld16(taddr + j * 16, r);
                        }
                    }
                    load_acc(j, r, false);
#pragma unroll
                    for (int i = 0; i < CH; ++i) {
                        float t = __uint_as_float(r[i]);
                        if (a.bias != nullptr) t += __ldg(a.bias + n0 + j * CH + i);
                        t = lrelu_f(t * a.scale, a.slope);
                        ss += t * t;
                    }
                }
                rn = rsqrtf(ss * (1.f / BLOCK_N) + a.pn_eps);
                if (valid) a.pn_rnorm[pix] = rn;
            }
#pragma unroll 1
            for (int j = 0; j < BLOCK_N / CH; ++j) {
                __syncwarp();
                uint32_t r[32];
                if (part0 == nullptr) {
                    if constexpr (CH == 32) {
                        tmem_ld32(taddr + j * 32, r);
                    } else {
                        tmem_ld16(taddr + j * 16, r);
                    }
                }
                uint4 mnxt[MV];
                if (masked && j + 1 < BLOCK_N / CH) {
                    const uint4* mp = reinterpret_cast<const uint4*>(reinterpret_cast<const T*>(a.mask_src) + off + (j + 1) * CH);
#pragma unroll
                    for (int c = 0; c < MV; ++c) mnxt[c] = __ldg(mp + c);
                }
                load_acc(j, r, true);
                float v[32];
#pragma unroll
                for (int i = 0; i < CH; ++i) {
                    float t = __uint_as_float(r[i]);
                    if (a.bias != nullptr) t += __ldg(a.bias + n0 + j * CH + i);
                    t *= a.scale;
                    v[i] = lrelu_f(t, a.slope) * rn;
                }
                if (valid) {
                    if (masked) {
#pragma unroll
                        for (int c = 0; c < MV; ++c) {
                            const T* mv = reinterpret_cast<const T*>(&mcur[c]);
#pragma unroll
                            for (int e = 0; e < VN; ++e)
                                if (!act_pos(mv[e])) v[c * VN + e] *= a.mask_slope;
                        }
#pragma unroll
                        for (int c = 0; c < MV; ++c) mcur[c] = mnxt[c];
                    }
                    if (is16(a.out_dtype)) {
                        uint16_t* op = reinterpret_cast<uint16_t*>(a.y) + off + j * CH;
                        const bool f16 = a.f16 != 0;
#pragma unroll
                        for (int c = 0; c < CH / 8; ++c) {
                            uint4 pk;
                            uint32_t* h2 = reinterpret_cast<uint32_t*>(&pk);
#pragma unroll
                            for (int e = 0; e < 4; ++e) h2[e] = pack16(v[c * 8 + 2 * e], v[c * 8 + 2 * e + 1], f16);
                            reinterpret_cast<uint4*>(op)[c] = pk;
                        }
                    } else {
                        float* op = reinterpret_cast<float*>(a.y) + off + j * CH;
#pragma unroll
                        for (int c = 0; c < CH / 4; ++c) {
                            float4 pk;
                            pk.x = v[c * 4 + 0];
                            pk.y = v[c * 4 + 1];
                            pk.z = v[c * 4 + 2];
                            pk.w = v[c * 4 + 3];
                            reinterpret_cast<float4*>(op)[c] = pk;
                        }
                    }
                }
            }
            }   // m
            tc_fence_before();
            mbar_arrive(&tempty_bar[acc]);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<Cfg::kTmemCols>(tmem_base);
    }
}

// ------------------------------------------------------------------- swapped orientation (N <= 128)
// tcgen05.mma reads its A operand from shared memory at a fixed ~128 cycles per instruction (M = 128
// rows x 32 B), which only the N = 256 shape hides: with pixels on M, a layer with 64 / 128 output
// channels runs the tensor pipe at 25 % / 50 %.  For those layers the roles are swapped:
//   D[128 output channels x 256 pixels] += A[128 weight rows x 128 B of K] * B[256 pixels x 128 B of K]^T
// (A = packed weights, rows beyond N zero-filled by TMA; B = one 256-pixel TMA box, shifted per tap).
// The accumulator is channel-major (TMEM lane = channel), so the epilogue transposes 32-pixel slabs
// through shared memory and writes NHWC rows with coalesced 16-byte stores.
constexpr int kSwPix = 256;
constexpr int kSwABytes = 128 * 128;
constexpr int kSwBBytes = kSwPix * 128;
constexpr int kSwStageBytes = kSwABytes + kSwBBytes;
constexpr int kSwStages = 4;
constexpr int kSwStagingBytes = 32 * 128 * 4;     // 32 pixels x 128 channels x fp32
constexpr int kSwSmemBytes = kSwStages * kSwStageBytes + kSwStagingBytes + 1024 + 256;

template <typename T>
__global__ void __launch_bounds__(kConvThreads, 1)
conv_gemm_swapped_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
                         const ConvKArgs a) {
    constexpr bool kTF32 = (sizeof(T) == 4);
    constexpr int BK = 128 / sizeof(T);
    constexpr int kStages = kSwStages;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* staging = smem + kStages * kSwStageBytes;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(staging + kSwStagingBytes);
    uint64_t* empty_bar = full_bar + kStages;
    uint64_t* tfull_bar = empty_bar + kStages;
    uint64_t* tempty_bar = tfull_bar + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_x);
        tma_prefetch_desc(&tmap_w);
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
    if (warp == 1) tmem_alloc<512>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    PG_PDL_ENTER();        // prologue done (barriers, TMEM, descriptor prefetch): now wait for the predecessor's data

    const int num_m_tiles = a.tiles_w * a.tiles_h * a.tiles_b;      // pixel tiles (256 pixels each)
    const int num_tiles = num_m_tiles * a.tiles_n;                  // x channel tiles (128 channels each)
    const int num_kc = a.taps * a.kc_per_tap;

    if (warp == 0) {
        if (elect_one()) {    // (not lane == 0: ptxas keeps descriptors in uniform registers only under elect.sync)
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int nt = tile % a.tiles_n;
                const int mt = tile / a.tiles_n;
                const int w0 = (mt % a.tiles_w) * a.BW;
                const int h0 = ((mt / a.tiles_w) % a.tiles_h) * a.BH;
                const int b0 = (mt / (a.tiles_w * a.tiles_h)) * a.BB;
                for (int kc = 0; kc < num_kc; ++kc) {
                    const int tap = kc / a.kc_per_tap;
                    const int cc = kc - tap * a.kc_per_tap;
                    int dh = 0, dw = 0;
                    if (a.taps == 9) {
                        dh = tap / 3 - 1;
                        dw = tap % 3 - 1;
                    }
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t* sa = smem + stage * kSwStageBytes;
                    mbar_arrive_expect_tx(&full_bar[stage], kSwStageBytes);
                    tma_load_2d(sa, &tmap_w, &full_bar[stage], kc * BK, nt * 128);
                    tma_load_4d(sa + kSwABytes, &tmap_x, &full_bar[stage], cc * BK, w0 + dw, h0 + dh, b0);
                    if (++stage == kStages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {    // (not lane == 0: ptxas keeps descriptors in uniform registers only under elect.sync)
            const uint32_t idesc = make_idesc2(kTF32 ? 2 : (a.wf16 ? 0 : 1), kTF32 ? 2 : (a.f16 ? 0 : 1), 0, 0, 128, kSwPix);
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
                const int acc = it & 1;
                const uint32_t acc_phase = (it >> 1) & 1;
                mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * kSwPix;
                for (int kc = 0; kc < num_kc; ++kc) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + stage * kSwStageBytes);
                    const uint64_t adesc = make_smem_desc(sa, 16, 1024);
                    const uint64_t bdesc = make_smem_desc(sa + kSwABytes, 16, 1024);
#pragma unroll
                    for (int k = 0; k < 4; ++k) umma_ss<kTF32>(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kc | k) != 0);
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
        // epilogue: thread = one output channel (TMEM lane); 8 slabs of 32 pixels per tile
        const int q = warp & 3;
        const int cl = q * 32 + lane;                       // channel within the 128-channel tile
        const int et = threadIdx.x - 64;                    // 0..127 among the epilogue threads
        const int osz = is16(a.out_dtype) ? 2 : 4;
        const bool f16 = a.f16 != 0;
        int it = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
            const int acc = it & 1;
            const uint32_t acc_phase = (it >> 1) & 1;
            const int nt = tile % a.tiles_n;
            const int mt = tile / a.tiles_n;
            const int w0 = (mt % a.tiles_w) * a.BW;
            const int h0 = ((mt / a.tiles_w) % a.tiles_h) * a.BH;
            const int b0 = (mt / (a.tiles_w * a.tiles_h)) * a.BB;
            const int n0 = nt * 128;
            const int cvalid = (a.N - n0) < 128 ? (a.N - n0) : 128;
            const float bias = (a.bias != nullptr && cl < cvalid) ? __ldg(a.bias + n0 + cl) : 0.f;
            const int row_bytes = cvalid * osz;             // one pixel's channel slice in the output
            const int vec_per_row = row_bytes / 16;

            // dgrad mask vectors of a slab are fetched one slab ahead (slab 0 before the accumulator wait), so the
            // global-load latency is off the store loop's critical path
            constexpr int MAXI = sizeof(T) == 2 ? 4 : 8;    // 16-byte vectors per thread per slab (<= 32*32/128)
            const bool masked = a.mask_src != nullptr && osz == static_cast<int>(sizeof(T));
            auto load_masks = [&](int j, uint4 (&m)[MAXI]) {
#pragma unroll
                for (int i = 0; i < MAXI; ++i) {
                    const int v = et + i * 128;
                    if (v < 32 * vec_per_row) {
                        const int p = v / vec_per_row, vv = v - p * vec_per_row;
                        const int row = j * 32 + p;
                        const int wo = row % a.BW;
                        const int ho = (row / a.BW) % a.BH;
                        const int b = b0 + row / (a.BW * a.BH);
                        if (b < a.B) {
                            const size_t pix = (static_cast<size_t>(b) * a.H + (h0 + ho)) * a.W + (w0 + wo);
                            m[i] = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const T*>(a.mask_src) + pix * a.N + n0) + vv);
                        }
                    }
                }
            };
            uint4 mcur[MAXI], mnxt[MAXI];
            if (masked) load_masks(0, mcur);
            mbar_wait(&tfull_bar[acc], acc_phase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * kSwPix;
#pragma unroll 1
            for (int j = 0; j < kSwPix / 32; ++j) {
                __syncwarp();
                uint32_t r[32];
                tmem_ld32(taddr + j * 32, r);
                if (masked && j + 1 < kSwPix / 32) load_masks(j + 1, mnxt);
                tmem_ld_wait();
                // transpose through smem: staging[pixel p][channel cl]
                if (cl < cvalid) {
                    if (osz == 2) {
                        uint16_t* st = reinterpret_cast<uint16_t*>(staging);
#pragma unroll
                        for (int p = 0; p < 32; ++p)
                            st[p * cvalid + cl] = cvt16(lrelu_f((__uint_as_float(r[p]) + bias) * a.scale, a.slope), f16);
                    } else {
                        float* st = reinterpret_cast<float*>(staging);
#pragma unroll
                        for (int p = 0; p < 32; ++p)
                            st[p * cvalid + cl] = lrelu_f((__uint_as_float(r[p]) + bias) * a.scale, a.slope);
                    }
                }
                asm volatile("bar.sync 1, 128;" ::: "memory");
                // coalesced store of the 32 x cvalid slab (+ optional LeakyReLU mask of the saved activation;
                // mask_src has the operand dtype T, so only the out dtype == T case is masked)
#pragma unroll
                for (int i = 0; i < 2 * MAXI; ++i) {        // fp32 output of bf16 operands has twice the vectors
                    const int v = et + i * 128;
                    if (v >= 32 * vec_per_row) break;
                    const int p = v / vec_per_row, vv = v - p * vec_per_row;
                    const int row = j * 32 + p;
                    const int wo = row % a.BW;
                    const int ho = (row / a.BW) % a.BH;
                    const int b = b0 + row / (a.BW * a.BH);
                    if (b >= a.B) continue;
                    const size_t pix = (static_cast<size_t>(b) * a.H + (h0 + ho)) * a.W + (w0 + wo);
                    uint4 val = *reinterpret_cast<const uint4*>(staging + static_cast<size_t>(p) * row_bytes + vv * 16);
                    if (masked && i < MAXI) {
                        constexpr int VN = 16 / sizeof(T);
                        T* ov = reinterpret_cast<T*>(&val);
                        const T* mv = reinterpret_cast<const T*>(&mcur[i < MAXI ? i : 0]);
#pragma unroll
                        for (int e = 0; e < VN; ++e)
                            if (!act_pos(mv[e])) {
                                if constexpr (sizeof(T) == 2) {
                                    uint16_t* ob = reinterpret_cast<uint16_t*>(&ov[e]);
                                    *ob = cvt16(f16bits_to_f(*ob, f16) * a.mask_slope, f16);
                                } else {
                                    ov[e] = ov[e] * a.mask_slope;
                                }
                            }
                    }
                    *reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(a.y) + (pix * a.N + n0) * osz + vv * 16) = val;
                }
                if (masked) {
#pragma unroll
                    for (int i = 0; i < MAXI; ++i) mcur[i] = mnxt[i];
                }
                asm volatile("bar.sync 1, 128;" ::: "memory");
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
        tmem_dealloc<512>(tmem_base);
    }
}

// ------------------------------------------------------------------------- SIMT direct kernel
// One thread per (pixel, 4 consecutive n); any shape.  Bring-up cross-check and the path for
// shapes the tensor-core kernel does not take (C*elsize % 128 != 0, non power-of-two H/W).
template <typename T>
__global__ void conv_direct_kernel(const T* __restrict__ x, const T* __restrict__ wp, const float* __restrict__ bias,
                                   const T* __restrict__ mask_src, void* __restrict__ y, int B, int H, int W, int C,
                                   int N, int taps, int out_dtype, float scale, float slope, float mask_slope, int xf16,
                                   int wf16) {
    PG_PDL_ENTER();
    // 16-bit instantiation: the element formats (bf16 / IEEE half) of x (= y, mask) and wp are run-time flags
    auto ldx = [&](const T* p) -> float {
        if constexpr (sizeof(T) == 2) return f16bits_to_f(*reinterpret_cast<const uint16_t*>(p), xf16 != 0);
        else return *p;
    };
    auto ldw = [&](const T* p) -> float {
        if constexpr (sizeof(T) == 2) return f16bits_to_f(*reinterpret_cast<const uint16_t*>(p), wf16 != 0);
        else return *p;
    };
    const long long total = static_cast<long long>(B) * H * W * N;
    for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int n = static_cast<int>(idx % N);
        const long long pix = idx / N;
        const int w = static_cast<int>(pix % W);
        const int h = static_cast<int>((pix / W) % H);
        const int b = static_cast<int>(pix / (static_cast<long long>(W) * H));
        float acc = 0.f;
        for (int t = 0; t < taps; ++t) {
            const int hh = h + (taps == 9 ? t / 3 - 1 : 0);
            const int ww = w + (taps == 9 ? t % 3 - 1 : 0);
            if (hh < 0 || hh >= H || ww < 0 || ww >= W) continue;
            const T* xp = x + ((static_cast<long long>(b) * H + hh) * W + ww) * C;
            const T* wr = wp + (static_cast<long long>(n) * taps + t) * C;
            for (int c = 0; c < C; ++c) acc += ldx(xp + c) * ldw(wr + c);
        }
        if (bias != nullptr) acc += bias[n];
        float v = lrelu_f(acc * scale, slope);
        if (mask_src != nullptr && !act_pos(mask_src[idx])) v *= mask_slope;
        if constexpr (sizeof(T) == 2) {
            if (is16(out_dtype)) {
                reinterpret_cast<uint16_t*>(y)[idx] = cvt16(v, xf16 != 0);
                continue;
            }
        }
        reinterpret_cast<float*>(y)[idx] = v;
    }
}

static bool is_pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }

// swapped orientation (256 pixels on the MMA N dimension) vs the plain kernel with a small N tile, measured on the
// 128^2 / 256^2 layer shapes after the issue fix: 128->64 266 vs 285 us, 256->128 132 vs 156 us (swapped wins),
// 128->128 116 vs 89 us (plain wins): the transposing epilogue needs K >= 1152 at N <= 64, K >= 2304 at N = 128
// (with the 256 x 128 tile of the plain kernel the N = 128 case goes the other way too — 256->128: 126 vs 132 us — so a
// big N = 128 layer never swaps; `big` = that tile shape applies)
static bool swap_pays(int n, long long k, bool big) {
    if (big && (n == 128 || n == 64)) return false;      // 256 x N tiles: 128->64 @256^2 189 us vs 266 us swapped
    return (n <= 64 && k >= 1152) || (n <= 128 && k >= 2304);
}
// the 256 x 128 tile (MT = 2) applies to N % 128 == 0 layers with at least 4 tiles per SM
static bool m2_applies(const pg_conv_params* p, int sms, int bn = 128) {
    // (fused up-sampling: the pixel boxes are cut from the low-resolution tensor, four parity tiles per box)
    const int XW = p->up2 ? p->W / 2 : p->W, XH = p->up2 ? p->H / 2 : p->H, mult = p->up2 ? 4 : 1;
    const int bw = XW < 128 ? XW : 128;
    const int bh = XH < (128 / bw) ? XH : (128 / bw);
    const int bb = 128 / (bw * bh);
    const long long m_tiles = (long long)(XW / bw) * (XH / bh) * ((p->B + bb - 1) / bb) * mult;
    return p->N % bn == 0 && m_tiles * (p->N / bn) >= 4LL * sms && (long long)XW * XH * p->B >= 512 &&
           (XW >= 256 || (256 % XW == 0));
}

// split-K workspace layout (pg_conv_params.ws_k): kSplitCounters zero-initialised arrival counters, then the partials
constexpr int kSplitCounters = 1024;
constexpr size_t kSplitHeader = kSplitCounters * sizeof(unsigned int);

template <typename T, int BLOCK_N, int SPAN = 128, int MT = 1>
static int launch_tc(const pg_conv_params* p, cudaStream_t stream, int ksplit = 1) {
    using Cfg = ConvCfg<BLOCK_N, SPAN, MT>;
    constexpr int kTileM = MT * kBlockM;
    constexpr int BK = SPAN / sizeof(T);
    ConvKArgs a;
    a.bias = p->bias;
    a.mask_src = p->mask_src;
    a.y = p->y;
    a.B = p->B; a.H = p->H; a.W = p->W; a.N = p->N;
    a.up2 = p->up2 ? 1 : 0;
    const int XW = a.up2 ? p->W / 2 : p->W, XH = a.up2 ? p->H / 2 : p->H;      // dims of the tensor x is read from
    a.BW = XW < kTileM ? XW : kTileM;
    a.BH = XH < (kTileM / a.BW) ? XH : (kTileM / a.BW);
    a.BB = kTileM / (a.BW * a.BH);
    a.tiles_w = XW / a.BW;
    a.tiles_h = XH / a.BH;
    a.tiles_b = (p->B + a.BB - 1) / a.BB;
    a.tiles_n = p->N / BLOCK_N;
    a.kc_per_tap = p->C / BK;
    a.taps = p->taps;
    a.out_dtype = p->out_dtype;
    a.f16 = p->dtype == PG_F16;
    a.wf16 = p->dtype == PG_F16;
    a.scale = p->scale; a.slope = p->slope; a.mask_slope = p->mask_slope;
    a.pn_rnorm = p->pn_rnorm; a.pn_eps = p->pn_eps;
    PG_CHECK(p->pn_rnorm == nullptr || a.tiles_n == 1, PG_ERR_UNSUPPORTED, "fused pixel norm needs one N tile (N=%d)", p->N);
    a.ksplit = 1;
    a.ws_part = nullptr;
    a.ws_count = nullptr;
    if (ksplit > 1) {
        const long long mn_tiles = (long long)a.tiles_w * a.tiles_h * a.tiles_b * a.tiles_n * (a.up2 ? 4 : 1);
        const size_t need = kSplitHeader + (size_t)mn_tiles * ((ksplit + 1) / 2) * kBlockM * BLOCK_N * sizeof(float);
        PG_CHECK(MT == 1 && p->ws_k != nullptr && (size_t)p->ws_k_bytes >= need && mn_tiles <= kSplitCounters &&
                     (p->taps * a.kc_per_tap) % ksplit == 0,
                 PG_ERR_ARG, "split-K: internal dispatch error (ksplit=%d)", ksplit);
        a.ksplit = ksplit;
        a.ws_count = reinterpret_cast<unsigned int*>(p->ws_k);
        a.ws_part = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(p->ws_k) + kSplitHeader);
    }

    const uint64_t es = sizeof(T);
    CUtensorMap tx, tw;
    {
        uint64_t dims[4] = {(uint64_t)p->C, (uint64_t)XW, (uint64_t)XH, (uint64_t)p->B};
        uint64_t strides[3] = {(uint64_t)p->C * es, (uint64_t)XW * p->C * es, (uint64_t)XH * XW * p->C * es};
        uint32_t box[4] = {(uint32_t)BK, (uint32_t)a.BW, (uint32_t)a.BH, (uint32_t)a.BB};
        int rc = encode_tmap(&tx, p->dtype, 4, p->x, dims, strides, box, SPAN);
        if (rc != PG_OK) return rc;
    }
    {
        const uint64_t K = (uint64_t)p->taps * p->C;
        uint64_t dims[2] = {K, (uint64_t)p->N};
        uint64_t strides[1] = {K * es};
        uint32_t box[2] = {(uint32_t)BK, (uint32_t)BLOCK_N};
        int rc = encode_tmap(&tw, p->dtype, 2, p->wp, dims, strides, box, SPAN);
        if (rc != PG_OK) return rc;
    }
    auto kern = conv_gemm_kernel<T, BLOCK_N, SPAN, MT>;
    PG_SET_SMEM_ONCE(kern, Cfg::kSmemBytes);
    const int num_tiles = a.tiles_w * a.tiles_h * a.tiles_b * a.tiles_n * (a.up2 ? 4 : 1) * a.ksplit;
    int grid = num_tiles < num_sms() ? num_tiles : num_sms();
    PG_CUDA(pg::launch_tc(kern, grid, kConvThreadsPlain, Cfg::kSmemBytes, stream, tx, tw, a));
    PG_LAUNCH_CHECK();
    g_tc_launches++;
    return PG_OK;
}

template <typename T>
static int launch_swapped(const pg_conv_params* p, cudaStream_t stream) {
    constexpr int BK = 128 / sizeof(T);
    ConvKArgs a;
    a.bias = p->bias;
    a.mask_src = p->mask_src;
    a.y = p->y;
    a.B = p->B; a.H = p->H; a.W = p->W; a.N = p->N;
    a.BW = p->W < kSwPix ? p->W : kSwPix;
    a.BH = p->H < (kSwPix / a.BW) ? p->H : (kSwPix / a.BW);
    a.BB = kSwPix / (a.BW * a.BH);
    a.tiles_w = p->W / a.BW;
    a.tiles_h = p->H / a.BH;
    a.tiles_b = (p->B + a.BB - 1) / a.BB;
    a.tiles_n = (p->N + 127) / 128;
    a.kc_per_tap = p->C / BK;
    a.taps = p->taps;
    a.out_dtype = p->out_dtype;
    a.f16 = p->dtype == PG_F16;
    a.wf16 = p->dtype == PG_F16;
    a.scale = p->scale; a.slope = p->slope; a.mask_slope = p->mask_slope;
    const uint64_t es = sizeof(T);
    CUtensorMap tx, tw;
    {
        uint64_t dims[4] = {(uint64_t)p->C, (uint64_t)p->W, (uint64_t)p->H, (uint64_t)p->B};
        uint64_t strides[3] = {(uint64_t)p->C * es, (uint64_t)p->W * p->C * es, (uint64_t)p->H * p->W * p->C * es};
        uint32_t box[4] = {(uint32_t)BK, (uint32_t)a.BW, (uint32_t)a.BH, (uint32_t)a.BB};
        int rc = encode_tmap(&tx, p->dtype, 4, p->x, dims, strides, box, 128);
        if (rc != PG_OK) return rc;
    }
    {
        const uint64_t K = (uint64_t)p->taps * p->C;
        uint64_t dims[2] = {K, (uint64_t)p->N};
        uint64_t strides[1] = {K * es};
        uint32_t box[2] = {(uint32_t)BK, 128u};
        int rc = encode_tmap(&tw, p->dtype, 2, p->wp, dims, strides, box, 128);
        if (rc != PG_OK) return rc;
    }
    auto kern = conv_gemm_swapped_kernel<T>;
    PG_SET_SMEM_ONCE(kern, kSwSmemBytes);
    const int num_tiles = a.tiles_w * a.tiles_h * a.tiles_b * a.tiles_n;
    const int grid = num_tiles < num_sms() ? num_tiles : num_sms();
    PG_CUDA(pg::launch_tc(kern, grid, kConvThreads, kSwSmemBytes, stream, tx, tw, a));
    PG_LAUNCH_CHECK();
    g_tc_launches++;
    return PG_OK;
}

template <typename T>
static int conv_dispatch(const pg_conv_params* p_in, cudaStream_t stream, int* kind_out = nullptr) {
    // kind_out != nullptr: only report which kernel the shape would run on (PG_CONV_KIND_*), no launch
    constexpr bool kF32 = sizeof(T) == 4;
    pg_conv_params local = *p_in;
    const pg_conv_params* p = &local;
    // fp32 runs as 3xTF32: K is tripled ((x,x,x_lo) . (w,w_lo,w)) through caller-provided workspaces
    const int Ceff = kF32 ? 3 * p->C : p->C;
    const int cbytes = Ceff * (int)sizeof(T);
    const int span = cbytes % 128 == 0 ? 128 : (cbytes % 64 == 0 ? 64 : (cbytes % 32 == 0 ? 32 : 0));
    const bool tc_ok = g_conv_backend == PG_BACKEND_AUTO && is_pow2(p->H) && is_pow2(p->W) && span != 0 &&
                       (p->N % 16 == 0) && (p->taps == 1 || p->taps == 9) && (!kF32 || (p->C % 4 == 0 && p->ws_x && p->ws_w)) &&
                       (reinterpret_cast<uintptr_t>(p->x) % 16 == 0) && (reinterpret_cast<uintptr_t>(p->wp) % 16 == 0) &&
                       (reinterpret_cast<uintptr_t>(p->y) % 16 == 0) &&
                       (p->mask_src == nullptr || reinterpret_cast<uintptr_t>(p->mask_src) % 16 == 0);
    if (p->up2 && kind_out == nullptr)
        PG_CHECK(tc_ok && p->taps == 9 && p->H % 2 == 0 && p->W % 2 == 0 && p->mask_src == nullptr, PG_ERR_UNSUPPORTED,
                 "fused up-sampling needs the tensor-core path (3x3, even H/W, no mask): B=%d H=%d W=%d C=%d N=%d", p->B,
                 p->H, p->W, p->C, p->N);
    if (kind_out != nullptr) {
        *kind_out = PG_CONV_KIND_SIMT;
        if (tc_ok) {
            const int osz_k = is16(p->out_dtype) ? 2 : 4;
            static const bool no_rows_k = getenv("PGGAN_NO_ROWS") != nullptr;
            static const bool no_swap_k = getenv("PGGAN_NO_SWAP") != nullptr;
            if (p->up2)
                *kind_out = PG_CONV_KIND_PLAIN;          // the fused up-sampling prologue exists in the plain kernel only
            else if (!kF32 && !no_rows_k && conv_rows_supported(p))
                *kind_out = PG_CONV_KIND_ROWS;
            else if (!no_swap_k && span == 128 && swap_pays(p->N, (long long)Ceff * p->taps, m2_applies(p, num_sms(), p->N == 64 ? 64 : 128)) && (p->N * osz_k) % 16 == 0 &&
                     (long long)p->B * p->H * p->W >= 2 * kSwPix)
                *kind_out = PG_CONV_KIND_SWAPPED;
            else
                *kind_out = PG_CONV_KIND_PLAIN;
        }
        return PG_OK;
    }
    if (tc_ok) {
        if (kF32) {
            const long long xpix = (long long)p->B * p->H * p->W / (p->up2 ? 4 : 1);
            int rc = split3_launch(reinterpret_cast<const float*>(p->x), reinterpret_cast<float*>(p->ws_x),
                                   xpix, p->C, 1, 2, stream);
            if (rc != PG_OK) return rc;
            rc = split3_launch(reinterpret_cast<const float*>(p->wp), reinterpret_cast<float*>(p->ws_w),
                               (long long)p->N * p->taps, p->C, 1, 1, stream);
            if (rc != PG_OK) return rc;
            local.x = p->ws_x;
            local.wp = p->ws_w;
            local.C = Ceff;
        }
        // narrow layers on wide images: input rows resident in smem, taps as shifted views (conv_rows.cu)
        static const bool no_rows = getenv("PGGAN_NO_ROWS") != nullptr;
        if (!kF32 && !no_rows && !p->up2 && conv_rows_supported(p)) return conv_rows_launch(p, stream);
        // few output channels: swapped orientation (channels on M, 256 pixels on N) keeps the tensor pipe fed
        const int osz = is16(p->out_dtype) ? 2 : 4;
        const bool mask_ok = p->mask_src == nullptr || osz == (int)sizeof(T);
        static const bool no_swap = getenv("PGGAN_NO_SWAP") != nullptr;
        // (its transposing epilogue costs a fixed ~5k cycles per 128x256 tile: only worth it when K is long enough)
        if (!no_swap && !p->up2 && span == 128 && swap_pays(p->N, (long long)p->C * p->taps, m2_applies(p, num_sms(), p->N == 64 ? 64 : 128)) && (p->N * osz) % 16 == 0 && mask_ok &&
            (long long)p->B * p->H * p->W >= 2 * kSwPix && p->pn_rnorm == nullptr)
            return launch_swapped<T>(p, stream);
        // largest N tile that still yields enough CTAs to occupy the machine
        const int XW = p->up2 ? p->W / 2 : p->W, XH = p->up2 ? p->H / 2 : p->H;
        const int bw = XW < 128 ? XW : 128;
        const int bh = XH < (128 / bw) ? XH : (128 / bw);
        const int bb = 128 / (bw * bh);
        const long long m_tiles = (long long)(XW / bw) * (XH / bh) * ((p->B + bb - 1) / bb) * (p->up2 ? 4 : 1);
        const int sms = num_sms();
        // N tile: minimise (waves over the SMs) x (per-tile cost).  Per-tile cost is the MMA time (~N) until the
        // fixed 16 KB A-tile shared-memory read per k-step dominates (N <= 64).
        if (span != 128) {
            // narrow layers (C*elsize = 64 or 32 bytes): N is small too, one N tile covers it
            if (span == 64) {
                if (p->N % 64 == 0) return launch_tc<T, 64, 64>(p, stream);
                if (p->N % 32 == 0) return launch_tc<T, 32, 64>(p, stream);
                return launch_tc<T, 16, 64>(p, stream);
            }
            if (p->N % 64 == 0) return launch_tc<T, 64, 32>(p, stream);
            if (p->N % 32 == 0) return launch_tc<T, 32, 32>(p, stream);
            return launch_tc<T, 16, 32>(p, stream);
        }
        // Few output tiles, long K: every CTA pulls its K extent through one SM's L2 port (measured per k-step, GPU time
        // inside a CUDA graph, tools/splitk_micro.py: 0.21 / 0.33 / 0.62 us at N tile 64 / 128 / 256 — 21-24 us for the
        // 72 k-steps of a 512-channel 3x3 layer however few pixels it has).  With a workspace the K extent is split over
        // the idle SMs instead: choose (N tile, splits) minimising   k-steps/S * time per k-step  +  the latency rounds
        // of the CTA that adds the pair buffers,  all CTAs in one wave; only for layers with at most ~40 tiles at N = 128
        // (beyond that the unsplit kernel already covers the machine: 16^2 x 16 measured 26.8 us unsplit, 30 us split).
        static const bool no_splitk = getenv("PGGAN_NO_SPLITK") != nullptr;
        if (!no_splitk && p->ws_k != nullptr && reinterpret_cast<uintptr_t>(p->ws_k) % 16 == 0 && !p->up2 && span == 128) {
            const int num_kc = p->taps * (p->C / (128 / (int)sizeof(T)));
            double t_best = 1e30;
            int bn_best = 0, s_best = 1;
            const int bns[3] = {64, 128, 256};
            const double tks[3] = {0.21, 0.33, 0.62};
            const bool few = m_tiles * ((p->N + 127) / 128) <= 40 && num_kc >= 16;
            for (int i = 0; i < 3 && few; ++i) {
                const int b = bns[i];
                if (p->N % b != 0) continue;
                if (p->pn_rnorm != nullptr && b != p->N) continue;
                const long long mn = m_tiles * (p->N / b);
                if (mn > sms || mn > kSplitCounters) continue;
                for (int S = 2; S <= 8; ++S) {
                    if (num_kc % S != 0 || mn * S > sms) continue;
                    const int npairs = (S + 1) / 2;
                    const size_t need = kSplitHeader + (size_t)mn * npairs * kBlockM * b * sizeof(float);
                    if (need > (size_t)p->ws_k_bytes) continue;
                    const double t = (num_kc / S) * tks[i] + 0.7 * (b / 32) * ((npairs + 1) / 2) * 0.5 + 0.1 * S * (b / 64);
                    if (t < t_best - 1e-9) {
                        t_best = t;
                        bn_best = b;
                        s_best = S;
                    }
                }
            }
            if (const char* f = getenv("PGGAN_FORCE_SPLITK")) {          // tuning experiments only: "bn,S"; "x,1": never split
                int fb = 0, fs = 0;
                if (sscanf(f, "%d,%d", &fb, &fs) == 2 && fs == 1) s_best = 1;
                if (sscanf(f, "%d,%d", &fb, &fs) == 2 && (fb == 64 || fb == 128 || fb == 256) && p->N % fb == 0 && fs >= 2 &&
                    (p->pn_rnorm == nullptr || fb == p->N) &&
                    num_kc % fs == 0 && m_tiles * (p->N / fb) * fs <= 2LL * sms && m_tiles * (p->N / fb) <= kSplitCounters &&
                    kSplitHeader + (size_t)m_tiles * (p->N / fb) * ((fs + 1) / 2) * kBlockM * fb * sizeof(float) <= (size_t)p->ws_k_bytes) {
                    bn_best = fb;
                    s_best = fs;
                }
            }
            if (s_best > 1) {
                switch (bn_best) {
                    case 256: return launch_tc<T, 256>(p, stream, s_best);
                    case 128: return launch_tc<T, 128>(p, stream, s_best);
                    default: return launch_tc<T, 64>(p, stream, s_best);
                }
            }
        }
        int bn = 16;
        double best = 1e30;
        const int cands[5] = {256, 128, 64, 32, 16};
        // measured per-tile cost (tools/sweep_bn.sh, 32^2 512->512): 1525 / 1094 / 599 TFLOP/s for N = 256 / 128 / 64
        // (smaller tiles move more L2->smem bytes per FLOP)
        const double cost[5] = {1.0, 0.70, 0.64, 0.62, 0.60};
        for (int i = 0; i < 5; ++i) {
            if (p->N % cands[i] != 0) continue;
            const long long tiles = m_tiles * (p->N / cands[i]);
            const double t = (double)((tiles + sms - 1) / sms) * cost[i];
            if (t < best - 1e-9) {
                best = t;
                bn = cands[i];
            }
        }
        if (p->pn_rnorm != nullptr) {                        // fused pixel norm: one N tile must cover the layer
            PG_CHECK(p->N == 16 || p->N == 32 || p->N == 64 || p->N == 128 || p->N == 256, PG_ERR_UNSUPPORTED,
                     "fused pixel norm: N=%d must be 16..256 (power of two)", p->N);
            bn = p->N;
        }
        if (const char* f = getenv("PGGAN_FORCE_BN")) {      // tuning experiments only
            const int v = atoi(f);
            if ((v == 16 || v == 32 || v == 64 || v == 128 || v == 256) && p->N % v == 0) bn = v;
        }
        // N tile 128 on a big layer: 256 x 128 tiles (two MMAs per K-step share the weight tile)
        static const bool no_m2 = getenv("PGGAN_NO_M2") != nullptr;
        if (bn == 128 && !no_m2 && m2_applies(p, sms)) return launch_tc<T, 128, 128, 2>(p, stream);
        if (bn == 64 && !no_m2 && m2_applies(p, sms, 64)) return launch_tc<T, 64, 128, 2>(p, stream);
        switch (bn) {
            case 256: return launch_tc<T, 256>(p, stream);
            case 128: return launch_tc<T, 128>(p, stream);
            case 64: return launch_tc<T, 64>(p, stream);
            case 32: return launch_tc<T, 32>(p, stream);
            default: return launch_tc<T, 16>(p, stream);
        }
    }
    PG_CHECK(p->pn_rnorm == nullptr, PG_ERR_UNSUPPORTED, "fused pixel norm is not available on the SIMT conv path");
    const long long total = (long long)p->B * p->H * p->W * p->N;
    if (total == 0) return PG_OK;
    long long blocks = (total + 255) / 256;
    if (blocks > 148LL * 32) blocks = 148LL * 32;
    PG_CUDA(pg::launch(conv_direct_kernel<T>, (int)blocks, 256, 0, stream, 
        reinterpret_cast<const T*>(p->x), reinterpret_cast<const T*>(p->wp), p->bias,
        reinterpret_cast<const T*>(p->mask_src), p->y, p->B, p->H, p->W, p->C, p->N, p->taps, p->out_dtype, p->scale,
        p->slope, p->mask_slope, p->dtype == PG_F16, p->dtype == PG_F16));
    PG_LAUNCH_CHECK();
    g_simt_launches++;
    return PG_OK;
}

}  // namespace pg

extern "C" int pggan_conv_kernel_kind(const pg_conv_params* p) {
    using namespace pg;
    if (p == nullptr || p->B <= 0 || p->H <= 0 || p->W <= 0 || p->C <= 0 || p->N <= 0) return PG_CONV_KIND_SIMT;
    pg_conv_params q = *p;
    // alignment checks only: absent pointers count as aligned
    if (!q.x) q.x = reinterpret_cast<const void*>(uintptr_t(256));
    if (!q.wp) q.wp = reinterpret_cast<const void*>(uintptr_t(256));
    if (!q.y) q.y = reinterpret_cast<void*>(uintptr_t(256));
    if (q.dtype == PG_F32) {
        if (!q.ws_x) q.ws_x = reinterpret_cast<void*>(uintptr_t(256));
        if (!q.ws_w) q.ws_w = reinterpret_cast<void*>(uintptr_t(256));
    }
    int kind = PG_CONV_KIND_SIMT;
    if (is16(q.dtype))
        conv_dispatch<__nv_bfloat16>(&q, nullptr, &kind);
    else if (q.dtype == PG_F32)
        conv_dispatch<float>(&q, nullptr, &kind);
    return kind;
}

extern "C" int pggan_conv_up2_fusable(const pg_conv_params* p) {
    // the fused call is offered where the UNFUSED convolution of the up-sampled tensor would run on the plain
    // tensor-core kernel anyway (the rows / swapped kernels of the narrow layers have no up-sampling prologue)
    if (p == nullptr || p->taps != 9 || p->H % 2 != 0 || p->W % 2 != 0 || p->H < 2 || p->W < 2) return 0;
    pg_conv_params q = *p;
    q.up2 = 0;
    q.mask_src = nullptr;
    return pggan_conv_kernel_kind(&q) == PG_CONV_KIND_PLAIN ? 1 : 0;
}

extern "C" int pggan_conv_pn_fusable(const pg_conv_params* p) {
    if (p == nullptr || p->taps != 9) return 0;
    const int kind = pggan_conv_kernel_kind(p);
    const bool pow2n = p->N == 16 || p->N == 32 || p->N == 64 || p->N == 128 || p->N == 256;
    if (kind == PG_CONV_KIND_ROWS) return 1;
    // the swapped-orientation kernel (channel-major accumulator) cannot; a layer it would take is left unfused
    const int cbytes = p->C * (p->dtype == PG_F32 ? 12 : 2);           // fp32 runs as 3xTF32: K tripled
    if (kind == PG_CONV_KIND_PLAIN && pow2n && cbytes % 128 == 0) return 1;   // (128-byte K span: N tile == N is chosen)
    return 0;
}

extern "C" int pggan_sizeof_conv_params(void) { return (int)sizeof(pg_conv_params); }
extern "C" int pggan_sizeof_wgrad_params(void) { return (int)sizeof(pg_wgrad_params); }

extern "C" int pggan_conv_fprop(const pg_conv_params* p, void* stream) {
    using namespace pg;
    PG_CHECK(p != nullptr, PG_ERR_ARG, "pggan_conv_fprop: null params");
    PG_CHECK(p->x && p->wp && p->y, PG_ERR_ARG, "pggan_conv_fprop: null tensor pointer");
    PG_CHECK(p->B > 0 && p->H > 0 && p->W > 0 && p->C > 0 && p->N > 0, PG_ERR_ARG,
             "pggan_conv_fprop: bad dims B=%d H=%d W=%d C=%d N=%d", p->B, p->H, p->W, p->C, p->N);
    PG_CHECK(p->taps == 1 || p->taps == 9, PG_ERR_ARG, "pggan_conv_fprop: taps must be 1 or 9 (got %d)", p->taps);
    PG_CHECK(p->dtype == PG_BF16 || p->dtype == PG_F16 || p->dtype == PG_F32, PG_ERR_ARG, "pggan_conv_fprop: bad dtype %d", p->dtype);
    PG_CHECK(p->out_dtype == p->dtype || p->out_dtype == PG_F32 || p->out_dtype == PG_F32_RAW, PG_ERR_ARG,
             "pggan_conv_fprop: bad out_dtype %d", p->out_dtype);
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    if (is16(p->dtype)) return conv_dispatch<__nv_bfloat16>(p, s);      // bf16 and f16 share the 16-bit instantiation
    return conv_dispatch<float>(p, s);
}
