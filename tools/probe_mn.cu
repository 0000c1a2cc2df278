// Probe for the narrow-layer weight-gradient design (wgrad_rows.cu): MN-major tcgen05 operands taken as OVERLAPPING,
// row-shifted views of NHWC rows resident in shared memory.
//
//   D[(j, c)][(i, n)] = sum_{p < 128}  x[p + j][c] * g[i][p][n]
//
//   A (MN-major, M = 128 = (128/C) spans of C channels): span j = the x row shifted by j pixels -> LBO = XSPAN bytes
//   B (MN-major, N = 3*n = 3 spans of n channels): span i = g row i                             -> LBO = g row pitch
//   K = pixels, 16 per MMA, 8 MMAs per 128-pixel row.
//
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o tools/probe_mn tools/probe_mn.cu pggan_b200/csrc/common.cu
#include "../pggan_b200/csrc/common.cuh"

#include <math.h>
#include <stdlib.h>
#include <vector>

using namespace pg;

constexpr int kXPix = 160;    // x pixels held in smem (128 + shifts)
constexpr int kGPix = 128;

template <int XSPAN, int GSPAN>
__global__ void __launch_bounds__(128, 1)
probe_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_g, float* out,
             int dst_shift, int sbo_mode) {
    constexpr int C = XSPAN / 2, NCH = GSPAN / 2;
    constexpr int NCOLS = 3 * NCH;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sx = smem + dst_shift * 128;
    uint8_t* sg = smem + 32768;
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 32768 + 3 * kGPix * GSPAN + 1024);
    uint64_t* done = bar + 1;
    uint32_t* slot = reinterpret_cast<uint32_t*>(done + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        mbar_init(done, 1);
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc<256>(slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *slot;
    if (threadIdx.x == 0) {
        mbar_arrive_expect_tx(bar, kXPix * XSPAN + 3 * kGPix * GSPAN);
        tma_load_2d(sx, &tmap_x, bar, 0, 0);
        tma_load_3d(sg, &tmap_g, bar, 0, 0, 0);
        mbar_wait(bar, 0);
        tc_fence_after();
        constexpr uint32_t kLayoutX = XSPAN == 128 ? 2u : (XSPAN == 64 ? 4u : 6u);
        constexpr uint32_t kLayoutG = GSPAN == 128 ? 2u : (GSPAN == 64 ? 4u : 6u);
        constexpr uint32_t idesc = make_idesc(1, 1, 1, 128, NCOLS);
        for (int k = 0; k < 8; ++k) {
            const uint64_t adesc = make_smem_desc(smem_u32(sx) + 16 * k * XSPAN, XSPAN, sbo_mode ? 1024 : 8 * XSPAN, kLayoutX);
            const uint64_t bdesc = make_smem_desc(smem_u32(sg) + 16 * k * GSPAN, kGPix * GSPAN, sbo_mode ? 1024 : 8 * GSPAN, kLayoutG);
            umma_ss<false>(tmem, adesc, bdesc, idesc, k != 0);
        }
        umma_commit(done);
    }
    mbar_wait(done, 0);
    tc_fence_after();
    const uint32_t taddr = tmem + (static_cast<uint32_t>(warp * 32) << 16);
    for (int j = 0; j < NCOLS / 16; ++j) {
        uint32_t r[32];
        tmem_ld16(taddr + j * 16, r);
        tmem_ld_wait();
        for (int i = 0; i < 16; ++i) out[(warp * 32 + lane) * NCOLS + j * 16 + i] = __uint_as_float(r[i]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<256>(tmem);
    }
}

static float bf(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }

template <int XSPAN, int GSPAN>
static int run() {
    constexpr int C = XSPAN / 2, NCH = GSPAN / 2, NCOLS = 3 * NCH;
    std::vector<__nv_bfloat16> hx(kXPix * C), hg(3 * kGPix * NCH);
    std::vector<float> fx(kXPix * C), fg(3 * kGPix * NCH);
    srand(77 + XSPAN + GSPAN);
    for (size_t i = 0; i < fx.size(); ++i) { fx[i] = bf((rand() % 2001 - 1000) / 500.f); hx[i] = __float2bfloat16_rn(fx[i]); }
    for (size_t i = 0; i < fg.size(); ++i) { fg[i] = bf((rand() % 2001 - 1000) / 500.f); hg[i] = __float2bfloat16_rn(fg[i]); }
    __nv_bfloat16 *dx, *dg;
    float* dout;
    cudaMalloc(&dx, hx.size() * 2);
    cudaMalloc(&dg, hg.size() * 2);
    cudaMalloc(&dout, 128 * NCOLS * 4);
    cudaMemcpy(dx, hx.data(), hx.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dg, hg.data(), hg.size() * 2, cudaMemcpyHostToDevice);
    CUtensorMap tx, tg;
    {
        uint64_t dims[2] = {(uint64_t)C, (uint64_t)kXPix};
        uint64_t strides[1] = {(uint64_t)XSPAN};
        uint32_t box[2] = {(uint32_t)C, (uint32_t)kXPix};
        if (encode_tmap(&tx, PG_BF16, 2, dx, dims, strides, box, XSPAN) != PG_OK) { printf("tmap x failed\n"); return 1; }
    }
    {
        uint64_t dims[3] = {(uint64_t)NCH, (uint64_t)kGPix, 3};
        uint64_t strides[2] = {(uint64_t)GSPAN, (uint64_t)GSPAN * kGPix};
        uint32_t box[3] = {(uint32_t)NCH, (uint32_t)kGPix, 3};
        if (encode_tmap(&tg, PG_BF16, 3, dg, dims, strides, box, GSPAN) != PG_OK) { printf("tmap g failed\n"); return 1; }
    }
    const int smem_bytes = 32768 + 3 * kGPix * GSPAN + 1024 + 1024 + 256;
    cudaFuncSetAttribute(probe_kernel<XSPAN, GSPAN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    std::vector<float> hout(128 * NCOLS);
    for (int sbo_mode = 0; sbo_mode < 2; ++sbo_mode)
        for (int dst_shift : {0, 1, 3}) {
            cudaMemset(dout, 0, hout.size() * 4);
            probe_kernel<XSPAN, GSPAN><<<1, 128, smem_bytes>>>(tx, tg, dout, dst_shift, sbo_mode);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("XSPAN %d GSPAN %d: cuda error %s\n", XSPAN, GSPAN, cudaGetErrorString(e)); return 2; }
            cudaMemcpy(hout.data(), dout, hout.size() * 4, cudaMemcpyDeviceToHost);
            int bad = 0, bad_span[8] = {0};
            for (int m = 0; m < 128; ++m)
                for (int col = 0; col < NCOLS; ++col) {
                    const int j = m / C, c = m % C, i = col / NCH, n = col % NCH;
                    double ref = 0;
                    for (int p = 0; p < 128; ++p) ref += (double)fx[(p + j) * C + c] * fg[(i * kGPix + p) * NCH + n];
                    if (fabs(ref - hout[m * NCOLS + col]) > 0.05) { ++bad; ++bad_span[j < 8 ? j : 7]; }
                }
            printf("XSPAN %3d GSPAN %3d sbo_mode %d dst_shift %d : %s (%d bad; per A-span:", XSPAN, GSPAN, sbo_mode, dst_shift,
                   bad ? "BAD" : "ok", bad);
            for (int j = 0; j < 128 / C && j < 8; ++j) printf(" %d", bad_span[j]);
            printf(")\n");
        }
    cudaFree(dx); cudaFree(dg); cudaFree(dout);
    return 0;
}

int main() {
    int rc = 0;
    rc |= run<128, 128>();
    rc |= run<128, 64>();
    rc |= run<64, 64>();
    rc |= run<64, 32>();
    rc |= run<32, 32>();
    rc |= run<32, 64>();
    rc |= run<32, 128>();
    return rc;
}
