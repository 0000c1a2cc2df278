// 3x3 weight gradient for the NARROW layers on wide images (16 / 32 / 64 channels on both sides, W % 128 == 0:
// the 256^2 / 512^2 / 1024^2 stages), on tcgen05 with BOTH operands taken as overlapping views of NHWC rows that
// are resident in shared memory.
//
//   dW[n][dy][dx][c] = sum_{b,y,xp} g[b,y,xp][n] * x[b, y+dy-1, xp+dx-1][c]
//
// One tcgen05.mma consumes 16 pixels of the reduction and costs ~128 cycles whatever its M x N (the A operand is
// read from smem at 32 B/clk), so the only way to make these tiny-channel layers cheap is to put ALL NINE taps of
// one pixel run into a single instruction:
//   A (MN-major, M = 128): the x row as (128/C) spans of C channels, span j = the row shifted by j pixels
//                          (descriptor LBO = one pixel = C*2 bytes: the spans OVERLAP)      -> M index = (dx, c)
//   B (MN-major, N = 3n) : three consecutive g rows, span i = row (r-1+i) (LBO = the row pitch) -> N index = (dy, n)
//   D[(dx,c)][(dy,n)] += sum over the 16 pixels.                 (layout facts measured in tools/probe_mn.cu)
// A CTA walks down a 128-pixel column strip: x rows and g rows stream through two smem rings (one TMA box per row,
// zero-filled outside the image = the conv padding; g rows outside the CTA's own range are loaded as zeros so each
// g row is counted by exactly one CTA), 8 (C = 64: 16) MMAs per row, accumulators live in TMEM for the whole kernel
// and are written once, as this CTA's fp32 partial-sum slice [n][tap][c]; pggan_unpack_wgrad sums the slices.
// The SIMT kernel this replaces ran at ~20 TFLOP/s (1.0-1.8 ms per 1024^2 layer, 36 % of the 1024^2 step).
//
// Replaces autograd's convolution_backward (weight) for EqualizedConv2d (reference lib/layers.py:60,86-89).
#include "common.cuh"

#include <atomic>

namespace pg {

extern std::atomic<long long> g_tc_launches;

constexpr int kWrThreads = 224;      // TMA warp, MMA warp, 4 epilogue warps, second MMA warp

struct WrKArgs {
    float* acc;
    long long slice_elems;
    float* direct;           // != NULL: add dscale * partial sums straight into this [N][9][C] tensor (no slices)
    float dscale;
    int B, H, W, C, N;
    int nseg;
    int f16;                 // g is IEEE half (else bf16)
    int xf16;                // ... and x (mixed precision: g bf16, x fp16)
    long long total, per_cta;
};

template <int XSPAN, int GSPAN> struct WrCfg {
    static constexpr int C = XSPAN / 2, NCH = GSPAN / 2;
    static constexpr int kSpansPerA = 128 / C;                       // dx values covered by one M = 128 operand
    static constexpr int kNumA = (3 + kSpansPerA - 1) / kSpansPerA;  // MMAs per 16-pixel step (2 for C = 64)
    static constexpr int NCOLS = 3 * NCH;
    // MMAs into ONE accumulator are a dependent chain; small-N MMAs are then paced by the accumulate latency, not by
    // tensor work.  kSets independent accumulator sets are used round-robin over the 16-pixel steps and summed in the
    // epilogue (as many as fit in TMEM, up to 4)
    static constexpr int kSets = (4 * kNumA * NCOLS <= 512) ? 4 : ((2 * kNumA * NCOLS <= 512) ? 2 : 1);
    // one issuing thread sustains an MMA every ~62 cycles (tools/probe_issue.cu), two issuing warps ~40-48: with
    // >= 2 accumulator sets the 16-pixel steps are split even / odd between two MMA warps
    static constexpr int kIssuers = kSets >= 2 ? 2 : 1;
    static constexpr int kTmemRaw = kSets * kNumA * NCOLS;
    static constexpr int kTmemCols = kTmemRaw <= 32 ? 32 : (kTmemRaw <= 64 ? 64 : (kTmemRaw <= 128 ? 128 : (kTmemRaw <= 256 ? 256 : 512)));
    // x row pitch: 128 + 2 halo + the extra pixels the unused high spans reach into, rounded to 128 bytes
    static constexpr int kXPixRaw = 128 + kNumA * kSpansPerA;
    static constexpr int kXPW = (kXPixRaw * XSPAN + 127) / 128 * 128 / XSPAN;
    static constexpr int kXSlot = kXPW * XSPAN;
    static constexpr int kGSlot = 128 * GSPAN;
    static constexpr int kBudget = 200 * 1024;
    static constexpr int kRingRaw = (kBudget - 2 * kGSlot) / (kXSlot + kGSlot);
    static constexpr int kRing = kRingRaw > 12 ? 12 : kRingRaw;
    static constexpr int kSmemBytes = kRing * kXSlot + (kRing + 2) * kGSlot + 1024 + 512;
    static_assert(kRing >= 4, "ring too small");
    static_assert(kXPW <= 256, "TMA box");
};

template <int XSPAN, int GSPAN>
__global__ void __launch_bounds__(kWrThreads, 1)
wgrad_rows_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_g,
                  const WrKArgs a) {
    using Cfg = WrCfg<XSPAN, GSPAN>;
    constexpr int R = Cfg::kRing;
    constexpr int C = Cfg::C, NCH = Cfg::NCH, NCOLS = Cfg::NCOLS;
    constexpr int kXSlot = Cfg::kXSlot, kGSlot = Cfg::kGSlot;
    constexpr uint32_t kLayoutX = XSPAN == 128 ? 2u : (XSPAN == 64 ? 4u : 6u);
    constexpr uint32_t kLayoutG = GSPAN == 128 ? 2u : (GSPAN == 64 ? 4u : 6u);

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* xring = smem;
    uint8_t* gring = smem + R * kXSlot;                  // R + 2 slots: slots R, R+1 mirror slots 0, 1
    uint64_t* xfull = reinterpret_cast<uint64_t*>(gring + (R + 2) * kGSlot);
    uint64_t* xempty = xfull + R;
    uint64_t* gfull = xempty + R;
    uint64_t* gempty = gfull + R;
    uint64_t* tfull = gempty + R;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull + 1);

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_x);
        tma_prefetch_desc(&tmap_g);
        for (int s = 0; s < R; ++s) {
            mbar_init(&xfull[s], 1);
            mbar_init(&xempty[s], Cfg::kIssuers);
            mbar_init(&gfull[s], 1);
            mbar_init(&gempty[s], Cfg::kIssuers);
        }
        mbar_init(tfull, Cfg::kIssuers);
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc<Cfg::kTmemCols>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    PG_PDL_ENTER();        // prologue done (barriers, TMEM, descriptor prefetch): now wait for the predecessor's data

    const long long u_begin = static_cast<long long>(blockIdx.x) * a.per_cta;
    const long long u_end = (u_begin + a.per_cta < a.total) ? (u_begin + a.per_cta) : a.total;

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer
        if (elect_one()) {    // (not lane == 0: ptxas keeps descriptors in uniform registers only under elect.sync)
            uint32_t LX = 0, LG = 0;
            auto load_g = [&](int yrow, int b, int seg) {         // yrow < 0: a zero row (fully out of bounds)
                const uint32_t slot = LG % R;
                mbar_wait(&gempty[slot], ((LG / R) & 1) ^ 1);
                mbar_arrive_expect_tx(&gfull[slot], slot < 2 ? 2 * kGSlot : kGSlot);
                tma_load_4d(gring + slot * kGSlot, &tmap_g, &gfull[slot], 0, seg * 128, yrow, b);
                if (slot < 2) tma_load_4d(gring + (R + slot) * kGSlot, &tmap_g, &gfull[slot], 0, seg * 128, yrow, b);
                ++LG;
            };
            long long u = u_begin;
            while (u < u_end) {
                const long long col = u / a.H;
                const int y0 = static_cast<int>(u - col * a.H);
                const long long strip_end = (col + 1) * a.H < u_end ? (col + 1) * a.H : u_end;
                const int th = static_cast<int>(strip_end - u);
                const int b = static_cast<int>(col / a.nseg);
                const int seg = static_cast<int>(col - static_cast<long long>(b) * a.nseg);
                // g index mg <-> image row y0-2+mg, real only inside [y0, y0+th); x index mx <-> image row y0-1+mx
                load_g(-8, b, seg);
                load_g(-8, b, seg);
                for (int m = 0; m < th + 2; ++m) {
                    const int gy = y0 + m;                        // g index m+2
                    load_g(m < th ? gy : -8, b, seg);
                    const uint32_t slot = LX % R;
                    mbar_wait(&xempty[slot], ((LX / R) & 1) ^ 1);
                    mbar_arrive_expect_tx(&xfull[slot], kXSlot);
                    tma_load_4d(xring + slot * kXSlot, &tmap_x, &xfull[slot], 0, seg * 128 - 1, y0 - 1 + m, b);
                    ++LX;
                }
                u = strip_end;
            }
        }
    } else if (warp == 1 || (warp == 6 && Cfg::kIssuers == 2)) {
        // -------------------------------------------------------------------- MMA issuer(s)
        const int iss = warp == 1 ? 0 : 1;
        if (elect_one()) {    // (not lane == 0: ptxas keeps descriptors in uniform registers only under elect.sync)
            const uint32_t idesc = make_idesc2(a.xf16 ? 0 : 1, a.f16 ? 0 : 1, 1, 1, 128, NCOLS);
            const uint32_t xring_addr = smem_u32(xring), gring_addr = smem_u32(gring);
            uint32_t LX = 0, LG = 0;
            bool first = true;
            long long u = u_begin;
            while (u < u_end) {
                const long long col = u / a.H;
                const long long strip_end = (col + 1) * a.H < u_end ? (col + 1) * a.H : u_end;
                const int th = static_cast<int>(strip_end - u);
                mbar_wait(&gfull[LG % R], (LG / R) & 1);
                mbar_wait(&gfull[(LG + 1) % R], ((LG + 1) / R) & 1);
                for (int m = 0; m < th + 2; ++m) {
                    const uint32_t g2 = LG + m + 2;
                    mbar_wait(&gfull[g2 % R], (g2 / R) & 1);
                    const uint32_t lx = LX + m;
                    mbar_wait(&xfull[lx % R], (lx / R) & 1);
                    tc_fence_after();
                    const uint32_t sx = xring_addr + (lx % R) * kXSlot;
                    const uint32_t sg = gring_addr + ((LG + m) % R) * kGSlot;     // rows m, m+1, m+2 follow physically
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        if (Cfg::kIssuers == 2 && (k & 1) != iss) continue;
                        const uint64_t bdesc = make_smem_desc(sg + 16 * k * GSPAN, kGSlot, 8 * GSPAN, kLayoutG);
#pragma unroll
                        for (int aa = 0; aa < Cfg::kNumA; ++aa) {
                            const uint64_t adesc = make_smem_desc(sx + (16 * k + aa * Cfg::kSpansPerA) * XSPAN, XSPAN,
                                                                  8 * XSPAN, kLayoutX);
                            umma_ss<false>(tmem_base + ((k % Cfg::kSets) * Cfg::kNumA + aa) * NCOLS, adesc, bdesc, idesc,
                                           (first && k < Cfg::kSets) ? 0u : 1u);
                        }
                    }
                    first = false;
                    umma_commit(&xempty[lx % R]);
                    umma_commit(&gempty[(LG + m) % R]);           // g index m is dead after x row m
                }
                umma_commit(&gempty[(LG + th + 2) % R]);
                umma_commit(&gempty[(LG + th + 3) % R]);
                LX += th + 2;
                LG += th + 4;
                u = strip_end;
            }
            umma_commit(tfull);
        }
    } else if (warp >= 2 && warp <= 5) {
        // ---------------------------------------------------- epilogue: once, this CTA's partial-sum slice
        const int q = warp & 3;
        const int L = q * 32 + lane;                   // TMEM lane = (dx within the operand, c)
        float* slice = a.acc + static_cast<size_t>(blockIdx.x) * a.slice_elems;
        mbar_wait(tfull, 0);
        tc_fence_after();
#pragma unroll 1
        for (int aa = 0; aa < Cfg::kNumA; ++aa) {
            const int dx = aa * Cfg::kSpansPerA + L / C;
            const int c = L % C;
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + aa * NCOLS;
#pragma unroll 1
            for (int j = 0; j < NCOLS / 16; ++j) {
                __syncwarp();
                uint32_t r[32];
                tmem_ld16(taddr + j * 16, r);
                tmem_ld_wait();
#pragma unroll
                for (int st = 1; st < Cfg::kSets; ++st) {          // add the other accumulator sets
                    uint32_t r2[32];
                    tmem_ld16(taddr + st * Cfg::kNumA * NCOLS + j * 16, r2);
                    tmem_ld_wait();
#pragma unroll
                    for (int e = 0; e < 16; ++e) r[e] = __float_as_uint(__uint_as_float(r[e]) + __uint_as_float(r2[e]));
                }
                const int i = (j * 16) / NCH;          // g row of the stack: dy = 2 - i
                const int n0 = (j * 16) % NCH;
                if (dx < 3 && u_begin < u_end) {
                    const int t = (2 - i) * 3 + dx;
                    if (a.direct != nullptr) {
#pragma unroll
                        for (int e = 0; e < 16; ++e)
                            asm volatile("red.global.add.f32 [%0], %1;" ::"l"(a.direct + (static_cast<size_t>(n0 + e) * 9 + t) * C + c),
                                         "f"(__uint_as_float(r[e]) * a.dscale)
                                         : "memory");
                    } else {
#pragma unroll
                        for (int e = 0; e < 16; ++e)
                            slice[(static_cast<size_t>(n0 + e) * 9 + t) * C + c] = __uint_as_float(r[e]);
                    }
                }
            }
        }
        tc_fence_before();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        tc_fence_after();
        tmem_dealloc<Cfg::kTmemCols>(tmem_base);
    }
}

static void wr_plan(const pg_wgrad_params* p, long long& total, long long& per_cta, int& grid) {
    const int nseg = p->W / 128;
    total = static_cast<long long>(p->B) * nseg * p->H;
    const int sms = num_sms();
    long long g = total < sms ? total : sms;
    per_cta = (total + g - 1) / g;
    grid = static_cast<int>((total + per_cta - 1) / per_cta);
}

bool wgrad_rows_supported(const pg_wgrad_params* p) {
    auto ok = [](int v) { return v == 16 || v == 32 || v == 64; };
    return is16(p->dtype) && p->taps == 9 && p->W % 128 == 0 && p->H >= 2 && ok(p->C) && ok(p->N);
}

int wgrad_rows_slices(const pg_wgrad_params* p) {
    long long total, per;
    int grid;
    wr_plan(p, total, per, grid);
    return grid;
}

template <int XSPAN, int GSPAN>
static int launch_wr(const pg_wgrad_params* p, cudaStream_t stream) {
    using Cfg = WrCfg<XSPAN, GSPAN>;
    WrKArgs a;
    a.acc = p->acc;
    a.direct = p->direct;
    a.dscale = p->direct_scale;
    a.slice_elems = static_cast<long long>(p->N) * 9 * p->C;
    a.B = p->B; a.H = p->H; a.W = p->W; a.C = p->C; a.N = p->N;
    a.nseg = p->W / 128;
    a.f16 = p->dtype == PG_F16;
    a.xf16 = p->dtype == PG_F16;
    int grid;
    wr_plan(p, a.total, a.per_cta, grid);
    CUtensorMap tx, tg;
    {
        uint64_t dims[4] = {(uint64_t)p->C, (uint64_t)p->W, (uint64_t)p->H, (uint64_t)p->B};
        uint64_t strides[3] = {(uint64_t)p->C * 2, (uint64_t)p->W * p->C * 2, (uint64_t)p->H * p->W * p->C * 2};
        uint32_t box[4] = {(uint32_t)p->C, (uint32_t)Cfg::kXPW, 1u, 1u};
        int rc = encode_tmap(&tx, p->dtype, 4, p->x, dims, strides, box, XSPAN);
        if (rc != PG_OK) return rc;
    }
    {
        uint64_t dims[4] = {(uint64_t)p->N, (uint64_t)p->W, (uint64_t)p->H, (uint64_t)p->B};
        uint64_t strides[3] = {(uint64_t)p->N * 2, (uint64_t)p->W * p->N * 2, (uint64_t)p->H * p->W * p->N * 2};
        uint32_t box[4] = {(uint32_t)p->N, 128u, 1u, 1u};
        int rc = encode_tmap(&tg, p->dtype, 4, p->g, dims, strides, box, GSPAN);
        if (rc != PG_OK) return rc;
    }
    auto kern = wgrad_rows_kernel<XSPAN, GSPAN>;
    PG_SET_SMEM_ONCE(kern, Cfg::kSmemBytes);
    PG_CUDA(pg::launch_tc(kern, grid, kWrThreads, Cfg::kSmemBytes, stream, tx, tg, a));
    PG_LAUNCH_CHECK();
    g_tc_launches++;
    return PG_OK;
}

int wgrad_rows_launch(const pg_wgrad_params* p, cudaStream_t stream) {
    const int xs = p->C * 2, gs = p->N * 2;
#define PG_WR_CASE(X, G) \
    if (xs == X && gs == G) return launch_wr<X, G>(p, stream);
    PG_WR_CASE(32, 32) PG_WR_CASE(32, 64) PG_WR_CASE(32, 128)
    PG_WR_CASE(64, 32) PG_WR_CASE(64, 64) PG_WR_CASE(64, 128)
    PG_WR_CASE(128, 32) PG_WR_CASE(128, 64) PG_WR_CASE(128, 128)
#undef PG_WR_CASE
    set_error("wgrad_rows_launch: unsupported shape C=%d N=%d", p->C, p->N);
    return PG_ERR_UNSUPPORTED;
}

}  // namespace pg
