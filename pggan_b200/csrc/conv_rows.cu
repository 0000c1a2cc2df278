// 3x3 convolution for the NARROW, WIDE-IMAGE layers (C*2 bytes = one swizzle span: 64 / 32 / 16 bf16 channels at
// 256^2 / 512^2 / 1024^2): input rows stay resident in shared memory and the nine filter taps are nine
// SHIFTED VIEWS of them.
//
// Why: conv_gemm_kernel re-fetches every input pixel through TMA once per tap (9x L2->SM traffic).  With N <= 128
// output channels there is too little tensor work per fetched byte to hide that: the 64->64 layers at 256^2 ran at
// ~30 % of the tensor peak and the 16-/32-channel layers at 5 % of the HBM roofline (profiles/r1_*).
//
// Here a CTA walks down a 128-pixel-wide column strip of one image:
//   * ring of R input rows in smem, each row = (128 + 2 halo) pixels x SPAN bytes, one TMA box per row (zero-filled
//     outside the image = the conv padding), 128B/64B/32B-swizzled exactly as the K-major UMMA layout wants;
//   * all 9 x [N x SPAN] weight tiles resident in smem for the whole kernel (loaded once per CTA);
//   * output row y = 9 x (SPAN/32) tcgen05.mma of shape 128 pixels x N x 16: the A descriptor of tap (dy,dx) simply
//     STARTS at  ring[(y+dy) % R] + dx*SPAN  — a row shift inside the swizzled tile.  The swizzle XOR is a function
//     of the absolute smem address (measured: tools/probe_shift.cu), so a descriptor may begin at any row;
//   * accumulators are 4-deep in TMEM; 4 epilogue warps apply bias/He scale/LeakyReLU/(dgrad mask) and store NHWC.
// Each input byte is read from L2/HBM ~once (+ 2 halo rows per strip, + 2 halo pixels per row).
//
// Pixel grouping (J = 4 for 16 channels, J = 2 for 32): one tcgen05.mma costs ~128 cycles whatever its N (the 128-row
// A operand is read from smem at 32 B/clk), so a 16-channel layer issuing 9 MMAs of K = 16 per 128 pixels is bound by
// the MMA count.  With J > 1 the NHWC tensor is VIEWED as [B][H][W/J][J*C = 64]: an MMA row is a group of J pixels,
// its N dimension the J*N outputs of the group, and the filter becomes a 3 x 3 filter over groups whose [J*N x 64]
// tiles are Toeplitz expansions of the 3x3 weights (built in smem at kernel start).  The left / right group taps
// only touch one pixel of the neighbour group, so only those K-steps are issued: 3*(J+2)*C/16 MMAs per 128*J pixels
// instead of 9*J*C/16 — 2x fewer at C = 16, 1.5x at C = 32 — and the same epilogue (a thread stores J*N contiguous
// channels-of-the-group = the NHWC bytes of J pixels).
//
// Replaces the same reference code as conv_gemm.cu (EqualizedConv2d 3x3 fprop / dgrad, lib/layers.py:58-63,86-89).
#include "common.cuh"

#include <atomic>
#include <stdlib.h>

namespace pg {

extern std::atomic<long long> g_tc_launches;

constexpr int kRowsThreads = 320;     // TMA warp, MMA warp, 2 x 4 epilogue warps (the groups take alternate output rows)
constexpr int kRowPix = 128;          // output pixels per row segment = MMA M

struct RowsKArgs {
    const float* bias;
    const void* mask_src;
    void* y;
    const __nv_bfloat16* wp;   // packed [Nr][9][Cr] weights (only read by the grouped variant)
    int Nr;                    // real output channels (bias index = column % Nr)
    int B, H, W, N, C;         // in GROUP units when J > 1: W = pixels/J per row, N = J*Nr, C = 64
    int nseg;                // W / 128
    long long total;         // B * nseg * H row segments
    long long per_cta;
    int out_dtype;
    int f16;                 // 16-bit x / y are IEEE half (else bf16)
    int wf16;                // ... and the packed weights (mixed precision: may differ)
    float scale, slope, mask_slope;
    float* pn_rnorm;         // fused pixel norm: per-pixel rsqrt factor out (NULL: none)
    float pn_eps;
};

template <int SPAN> struct RowsGeom {
    // halo row pitch in pixels: 128 + 2, rounded so that a row is a multiple of 128 bytes (TMA smem alignment)
    static constexpr int kPW = SPAN == 32 ? 132 : 130;
    static constexpr int kSlotBytes = kPW * SPAN;
};

template <int SPAN, int BLOCK_N> struct RowsCfg {
    static constexpr int kWBytes = 9 * BLOCK_N * SPAN;
    static constexpr int kSlot = RowsGeom<SPAN>::kSlotBytes;
    static constexpr int kAvail = 227 * 1024 - 1024 - 512 - kWBytes;
    static constexpr int kRingRaw = kAvail / kSlot;
    static constexpr int kRing = kRingRaw > 10 ? 10 : kRingRaw;
    static constexpr int kAcc = 4;
    static constexpr int kTmemCols = (kAcc * BLOCK_N < 32) ? 32 : kAcc * BLOCK_N;
    static constexpr int kSmemBytes = kWBytes + kRing * kSlot + 1024 + 512;
    static_assert(kRing >= 4, "ring too small");
    static_assert(kTmemCols <= 512, "TMEM");
};

template <int SPAN, int BLOCK_N, int J = 1>
__global__ void __launch_bounds__(kRowsThreads, 1)
conv_rows_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
                 const RowsKArgs a) {
    static_assert(J == 1 || SPAN == 128, "pixel groups fill one 128-byte span");
    using Cfg = RowsCfg<SPAN, BLOCK_N>;
    using T = __nv_bfloat16;
    constexpr int R = Cfg::kRing;
    constexpr int ACC = Cfg::kAcc;
    constexpr int kSlot = Cfg::kSlot;
    constexpr int kMmaPerTap = SPAN / 32;
    constexpr int kWTile = BLOCK_N * SPAN;
    constexpr uint32_t kLayout = SPAN == 128 ? 2u : (SPAN == 64 ? 4u : 6u);
    constexpr uint32_t kSBO = 8 * SPAN;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sw = smem;                                  // 9 weight tiles
    uint8_t* ring = smem + Cfg::kWBytes;                 // kWBytes is a multiple of 128
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(ring + R * kSlot);
    uint64_t* empty_bar = full_bar + R;
    uint64_t* tfull_bar = empty_bar + R;
    uint64_t* tempty_bar = tfull_bar + ACC;
    uint64_t* w_bar = tempty_bar + ACC;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_bar + 1);

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_x);
        tma_prefetch_desc(&tmap_w);
        for (int s = 0; s < R; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int s = 0; s < ACC; ++s) {
            mbar_init(&tfull_bar[s], 1);
            mbar_init(&tempty_bar[s], 128);
        }
        mbar_init(w_bar, 1);
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc<Cfg::kTmemCols>(tmem_slot);
    PG_PDL_ENTER();        // barriers and TMEM are set up: wait for the predecessor's data before the first global read
    if constexpr (J > 1) {
        // Toeplitz weight tiles, written straight into the SWIZZLE_128B K-major layout (16-byte chunk q of row r
        // lives at chunk q ^ (r & 7)); tile t = (dy, group tap gx), row = (pixel j of the group, n),
        // K = (pixel j' of the input group, c):  w[n][dy][dx][c] with dx = gx*J + j' - j + 1 if that is a tap
        constexpr int CS = 4 / J;                        // 32-byte K-steps per real pixel (C/16)
        constexpr int Cr = 64 / J;
        const int Nr = BLOCK_N / J;
        for (int idx = threadIdx.x; idx < 9 * BLOCK_N * 8; idx += kRowsThreads) {
            const int qq = idx & 7;
            const int row = (idx >> 3) % BLOCK_N;
            const int t = idx / (8 * BLOCK_N);
            const int dy = t / 3, gx = t % 3 - 1;
            const int j = row / Nr, n = row - j * Nr;
            const int k = qq >> 1, half = qq & 1;
            const int jp = k / CS, cs = k - jp * CS;
            const int dx = gx * J + jp - j + 1;
            uint4 val = make_uint4(0, 0, 0, 0);
            if (dx >= 0 && dx <= 2)
                val = __ldg(reinterpret_cast<const uint4*>(a.wp + (static_cast<size_t>(n) * 9 + dy * 3 + dx) * Cr + cs * 16 + half * 8));
            *reinterpret_cast<uint4*>(sw + t * kWTile + row * 128 + ((qq ^ (row & 7)) << 4)) = val;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");    // generic-proxy stores -> UMMA (async proxy) reads
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // this CTA's contiguous range of row segments u = (b*nseg + seg)*H + y, cut into strips at image-column ends
    const long long u_begin = static_cast<long long>(blockIdx.x) * a.per_cta;
    const long long u_end = (u_begin + a.per_cta < a.total) ? (u_begin + a.per_cta) : a.total;

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer
        if (u_begin < u_end && elect_one()) {
            if constexpr (J == 1) {
                mbar_arrive_expect_tx(w_bar, Cfg::kWBytes);
                for (int t = 0; t < 9; ++t) tma_load_2d(sw + t * kWTile, &tmap_w, w_bar, t * a.C, 0);
            }
            uint32_t L = 0;                 // running count of input rows loaded
            long long u = u_begin;
            while (u < u_end) {
                const long long col = u / a.H;                 // b*nseg + seg
                const int y0 = static_cast<int>(u - col * a.H);
                const long long strip_end = (col + 1) * a.H < u_end ? (col + 1) * a.H : u_end;
                const int th = static_cast<int>(strip_end - u);
                const int b = static_cast<int>(col / a.nseg);
                const int seg = static_cast<int>(col - static_cast<long long>(b) * a.nseg);
                for (int yy = y0 - 1; yy <= y0 + th; ++yy, ++L) {
                    const uint32_t slot = L % R;
                    mbar_wait(&empty_bar[slot], ((L / R) & 1) ^ 1);
                    mbar_arrive_expect_tx(&full_bar[slot], kSlot);
                    tma_load_4d(ring + slot * kSlot, &tmap_x, &full_bar[slot], 0, seg * kRowPix - 1, yy, b);
                }
                u = strip_end;
            }
        }
    } else if (warp == 1) {
        // -------------------------------------------------------------------- MMA issuer
        if (u_begin < u_end && elect_one()) {
            const uint32_t idesc = make_idesc2(a.f16 ? 0 : 1, a.wf16 ? 0 : 1, 0, 0, kRowPix, BLOCK_N);
            if constexpr (J == 1) {
                mbar_wait(w_bar, 0);
                tc_fence_after();
            }
            const uint32_t sw_addr = smem_u32(sw);
            const uint32_t ring_addr = smem_u32(ring);
            uint32_t L = 0, it = 0;
            long long u = u_begin;
            while (u < u_end) {
                const long long col = u / a.H;
                const long long strip_end = (col + 1) * a.H < u_end ? (col + 1) * a.H : u_end;
                const int th = static_cast<int>(strip_end - u);
                const uint32_t base = L;
                mbar_wait(&full_bar[base % R], (base / R) & 1);
                mbar_wait(&full_bar[(base + 1) % R], ((base + 1) / R) & 1);
                for (int i = 0; i < th; ++i, ++it) {
                    const uint32_t l2 = base + i + 2;
                    mbar_wait(&full_bar[l2 % R], (l2 / R) & 1);
                    const uint32_t acc = it % ACC;
                    mbar_wait(&tempty_bar[acc], ((it / ACC) & 1) ^ 1);
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
#pragma unroll
                    for (int dy = 0; dy < 3; ++dy) {
                        const uint32_t row_addr = ring_addr + ((base + i + dy) % R) * kSlot;
#pragma unroll
                        for (int dx = 0; dx < 3; ++dx) {
                            const uint64_t adesc = make_smem_desc(row_addr + dx * SPAN, 16, kSBO, kLayout);
                            const uint64_t bdesc = make_smem_desc(sw_addr + (dy * 3 + dx) * kWTile, 16, kSBO, kLayout);
                            // grouped: the left / right group taps only reach the neighbour group's last / first pixel
                            constexpr int CSK = J > 1 ? 4 / J : kMmaPerTap;
                            const int k0 = (J > 1 && dx == 0) ? kMmaPerTap - CSK : 0;
                            const int k1 = (J > 1 && dx == 2) ? CSK : kMmaPerTap;
#pragma unroll
                            for (int k = 0; k < kMmaPerTap; ++k)
                                if (k >= k0 && k < k1)
                                    umma_ss<false>(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (dy | dx | (k - k0)) != 0);
                        }
                    }
                    umma_commit(&tfull_bar[acc]);
                    umma_commit(&empty_bar[(base + i) % R]);      // input row i is dead once these MMAs retire
                }
                umma_commit(&empty_bar[(base + th) % R]);          // the two trailing halo rows
                umma_commit(&empty_bar[(base + th + 1) % R]);
                L = base + th + 2;
                u = strip_end;
            }
        }
    } else {
        // ---------------------------------------------------------------------- epilogue
        const int q = warp & 3;
        const int r = q * 32 + lane;                   // pixel within the 128-pixel row segment = TMEM lane
        // a lone warp per SM sub-partition leaves the epilogue latency-bound (ncu: ~900 dependent instructions per
        // row per warp); two groups of four warps take alternate rows, each accumulator still sees 128 arrivals
        const uint32_t egroup = (warp - 2) >> 2;
        uint32_t it = 0;
        long long u = u_begin;
        while (u < u_end) {
            const long long col = u / a.H;
            const int y0 = static_cast<int>(u - col * a.H);
            const long long strip_end = (col + 1) * a.H < u_end ? (col + 1) * a.H : u_end;
            const int th = static_cast<int>(strip_end - u);
            const int b = static_cast<int>(col / a.nseg);
            const int seg = static_cast<int>(col - static_cast<long long>(b) * a.nseg);
            for (int i = 0; i < th; ++i, ++it) {
                if ((it & 1u) != egroup) continue;
                const uint32_t acc = it % ACC;
                const size_t pix = (static_cast<size_t>(b) * a.H + (y0 + i)) * a.W + seg * kRowPix + r;
                const size_t off = pix * a.N;
                constexpr int CH = BLOCK_N >= 32 ? 32 : 16;
                constexpr int MV = CH / 8;                   // 16-byte mask vectors per step
                // dgrad mask: fetched one step ahead of its use (the first step's before the accumulator wait)
                const bool masked = a.mask_src != nullptr;
                uint4 mcur[MV];
                if (masked) {
                    const uint4* mp = reinterpret_cast<const uint4*>(reinterpret_cast<const T*>(a.mask_src) + off);
#pragma unroll
                    for (int c = 0; c < MV; ++c) mcur[c] = __ldg(mp + c);
                }
                mbar_wait(&tfull_bar[acc], (it / ACC) & 1);
                tc_fence_after();
                const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BLOCK_N;
                // fused pixel norm: this thread owns all channels of its pixel (J = 1) or of its J pixels (grouped:
                // columns [jj*Nr, (jj+1)*Nr) belong to pixel jj).  First pass: per-pixel mean of lrelu(..)^2.
                constexpr int NR = BLOCK_N / J;              // real channels per pixel
                float rnp[J];
#pragma unroll
                for (int jj = 0; jj < J; ++jj) rnp[jj] = 1.f;
                if (a.pn_rnorm != nullptr) {
                    float ssp[J];
#pragma unroll
                    for (int jj = 0; jj < J; ++jj) ssp[jj] = 0.f;
#pragma unroll 1
                    for (int j = 0; j < BLOCK_N / CH; ++j) {
                        __syncwarp();
                        uint32_t rr[32];
                        if constexpr (CH == 32) {
                            tmem_ld32(taddr + j * 32, rr);
                        } else {
                            tmem_ld16(taddr + j * 16, rr);
                        }
                        tmem_ld_wait();
                        constexpr int SEG = NR >= CH ? CH : NR;      // columns of one pixel inside this step
#pragma unroll
                        for (int s2 = 0; s2 < CH / SEG; ++s2) {
                            float sc = 0.f;
#pragma unroll
                            for (int e = 0; e < SEG; ++e) {
                                float t = __uint_as_float(rr[s2 * SEG + e]);
                                if (a.bias != nullptr) t += __ldg(a.bias + ((j * CH + s2 * SEG + e) & (a.Nr - 1)));
                                t = lrelu_f(t * a.scale, a.slope);
                                sc += t * t;
                            }
                            const int pj = (j * CH + s2 * SEG) / NR;
#pragma unroll
                            for (int jj = 0; jj < J; ++jj) ssp[jj] += (pj == jj) ? sc : 0.f;
                        }
                    }
#pragma unroll
                    for (int jj = 0; jj < J; ++jj) {
                        rnp[jj] = rsqrtf(ssp[jj] * (1.f / NR) + a.pn_eps);
                        a.pn_rnorm[pix * J + jj] = rnp[jj];
                    }
                }
#pragma unroll 1
                for (int j = 0; j < BLOCK_N / CH; ++j) {
                    __syncwarp();
                    uint32_t rr[32];
                    if constexpr (CH == 32) {
                        tmem_ld32(taddr + j * 32, rr);
                    } else {
                        tmem_ld16(taddr + j * 16, rr);
                    }
                    uint4 mnxt[MV];
                    if (masked && j + 1 < BLOCK_N / CH) {
                        const uint4* mp = reinterpret_cast<const uint4*>(reinterpret_cast<const T*>(a.mask_src) + off + (j + 1) * CH);
#pragma unroll
                        for (int c = 0; c < MV; ++c) mnxt[c] = __ldg(mp + c);
                    }
                    tmem_ld_wait();
                    float v[32];
#pragma unroll
                    for (int e = 0; e < CH; ++e) {
                        float t = __uint_as_float(rr[e]);
                        if (a.bias != nullptr) t += __ldg(a.bias + ((j * CH + e) & (a.Nr - 1)));
                        t *= a.scale;
                        v[e] = lrelu_f(t, a.slope);
                    }
                    if (a.pn_rnorm != nullptr) {
                        // factor of the pixel(s) this 32-/16-column step belongs to (selects, no dynamic register index)
                        if constexpr (NR >= CH) {
                            const int pj = (j * CH) / NR;
                            float rn = rnp[0];
#pragma unroll
                            for (int jj = 1; jj < J; ++jj) rn = (pj == jj) ? rnp[jj] : rn;
#pragma unroll
                            for (int e = 0; e < CH; ++e) v[e] *= rn;
                        } else {
                            constexpr int PPC = CH / NR;         // pixels per step
#pragma unroll
                            for (int s2 = 0; s2 < PPC; ++s2) {
                                const int pj = j * PPC + s2;
                                float rn = rnp[0];
#pragma unroll
                                for (int jj = 1; jj < J; ++jj) rn = (pj == jj) ? rnp[jj] : rn;
#pragma unroll
                                for (int e = 0; e < NR; ++e) v[s2 * NR + e] *= rn;
                            }
                        }
                    }
                    if (masked) {
#pragma unroll
                        for (int c = 0; c < MV; ++c) {
                            const uint16_t* mv = reinterpret_cast<const uint16_t*>(&mcur[c]);
#pragma unroll
                            for (int e = 0; e < 8; ++e)
                                if (!pos16(mv[e])) v[c * 8 + e] *= a.mask_slope;
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
                tc_fence_before();
                mbar_arrive(&tempty_bar[acc]);
            }
            u = strip_end;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<Cfg::kTmemCols>(tmem_base);
    }
}

template <int SPAN, int BLOCK_N, int J = 1>
static int launch_rows(const pg_conv_params* p_in, cudaStream_t stream) {
    using Cfg = RowsCfg<SPAN, BLOCK_N>;
    // J > 1: the same memory viewed as [B][H][W/J][J*C] -> [B][H][W/J][J*N]
    pg_conv_params view = *p_in;
    view.W = p_in->W / J;
    view.C = p_in->C * J;
    view.N = p_in->N * J;
    const pg_conv_params* p = &view;
    RowsKArgs a;
    a.bias = p->bias;
    a.mask_src = p->mask_src;
    a.y = p->y;
    a.wp = reinterpret_cast<const __nv_bfloat16*>(p->wp);
    a.Nr = p_in->N;
    a.B = p->B; a.H = p->H; a.W = p->W; a.N = p->N; a.C = p->C;
    a.nseg = p->W / kRowPix;
    a.total = static_cast<long long>(p->B) * a.nseg * p->H;
    a.out_dtype = p->out_dtype;
    a.f16 = p->dtype == PG_F16;
    a.wf16 = p->dtype == PG_F16;
    a.scale = p->scale; a.slope = p->slope; a.mask_slope = p->mask_slope;
    a.pn_rnorm = p->pn_rnorm; a.pn_eps = p->pn_eps;
    const int sms = num_sms();
    long long grid = a.total < sms ? a.total : sms;
    a.per_cta = (a.total + grid - 1) / grid;
    grid = (a.total + a.per_cta - 1) / a.per_cta;

    CUtensorMap tx, tw;
    {
        uint64_t dims[4] = {(uint64_t)p->C, (uint64_t)p->W, (uint64_t)p->H, (uint64_t)p->B};
        uint64_t strides[3] = {(uint64_t)p->C * 2, (uint64_t)p->W * p->C * 2, (uint64_t)p->H * p->W * p->C * 2};
        uint32_t box[4] = {(uint32_t)p->C, (uint32_t)RowsGeom<SPAN>::kPW, 1u, 1u};
        int rc = encode_tmap(&tx, p->dtype, 4, p->x, dims, strides, box, SPAN);
        if (rc != PG_OK) return rc;
    }
    if (J == 1) {
        const uint64_t K = 9ull * p->C;
        uint64_t dims[2] = {K, (uint64_t)p->N};
        uint64_t strides[1] = {K * 2};
        uint32_t box[2] = {(uint32_t)p->C, (uint32_t)BLOCK_N};
        int rc = encode_tmap(&tw, p->dtype, 2, p->wp, dims, strides, box, SPAN);
        if (rc != PG_OK) return rc;
    } else {
        tw = tx;                      // unused by the grouped variant (it builds its weight tiles itself)
    }
    auto kern = conv_rows_kernel<SPAN, BLOCK_N, J>;
    PG_SET_SMEM_ONCE(kern, Cfg::kSmemBytes);
    PG_CUDA(pg::launch_tc(kern, (int)grid, kRowsThreads, Cfg::kSmemBytes, stream, tx, tw, a));
    PG_LAUNCH_CHECK();
    g_tc_launches++;
    return PG_OK;
}

// true if the layer is in this kernel's envelope (the caller has checked pointer alignment and dtype bf16)
bool conv_rows_supported(const pg_conv_params* p) {
    const int cbytes = p->C * 2;
    return is16(p->dtype) && p->taps == 9 && (cbytes == 128 || cbytes == 64 || cbytes == 32) &&
           p->W % kRowPix == 0 && p->H >= 4 && (p->N == 16 || p->N == 32 || p->N == 64 || p->N == 128);
}

int conv_rows_launch(const pg_conv_params* p, cudaStream_t stream) {
    const int span = p->C * 2;
    // pixel grouping for 16 / 32 channels when the grouped row still splits into 128-group segments
    static const bool no_group = getenv("PGGAN_NO_GROUP") != nullptr;
    if (!no_group && span < 128) {
        const int J = 128 / span;
        // (J*N = 128 measured slower than ungrouped: 147 KB of weight tiles leave a 4-row ring and a long epilogue)
        if (p->W % (128 * J) == 0 && p->N * J <= 64) {
            const int np = p->N * J;
            if (J == 4 && np == 64) return launch_rows<128, 64, 4>(p, stream);
            if (J == 2 && np == 32) return launch_rows<128, 32, 2>(p, stream);
            if (J == 2 && np == 64) return launch_rows<128, 64, 2>(p, stream);
        }
    }
#define PG_ROWS_CASE(S, BN) \
    if (span == S && p->N == BN) return launch_rows<S, BN>(p, stream);
    PG_ROWS_CASE(128, 16) PG_ROWS_CASE(128, 32) PG_ROWS_CASE(128, 64) PG_ROWS_CASE(128, 128)
    PG_ROWS_CASE(64, 16) PG_ROWS_CASE(64, 32) PG_ROWS_CASE(64, 64) PG_ROWS_CASE(64, 128)
    PG_ROWS_CASE(32, 16) PG_ROWS_CASE(32, 32) PG_ROWS_CASE(32, 64) PG_ROWS_CASE(32, 128)
#undef PG_ROWS_CASE
    set_error("conv_rows_launch: unsupported shape C=%d N=%d", p->C, p->N);
    return PG_ERR_UNSUPPORTED;
}

}  // namespace pg
