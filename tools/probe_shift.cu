// Probe: can a tcgen05 shared-memory descriptor start at an arbitrary ROW of a swizzled K-major tile?
// (needed by the halo-resident 3x3 conv kernel: one input tile in smem, 9 tap-shifted A views of it)
//
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o tools/probe_shift tools/probe_shift.cu pggan_b200/csrc/common.cu
//   run  : tools/probe_shift
//
// For SPAN in {128,64,32}: TMA-load R rows x SPAN bytes (bf16) at smem offset dst_shift*SPAN, run one
// 128 x 64 x (SPAN/2) MMA chain with the A descriptor starting `shift` rows further, compare with the host product.
#include "../pggan_b200/csrc/common.cuh"

#include <math.h>
#include <stdlib.h>
#include <vector>

using namespace pg;

constexpr int kRows = 320;     // rows held in smem
constexpr int kN = 64;

template <int SPAN>
__global__ void __launch_bounds__(128, 1)
probe_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, float* out,
             int shift, int dst_shift, int base_off_mode) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sa = smem;                                  // (kRows + 16) * SPAN
    uint8_t* sb = smem + ((kRows + 16) * SPAN + 1023) / 1024 * 1024;
    uint64_t* bar = reinterpret_cast<uint64_t*>(sb + kN * SPAN);
    uint64_t* done = bar + 1;
    uint32_t* slot = reinterpret_cast<uint32_t*>(done + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        mbar_init(done, 1);
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc<64>(slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *slot;
    if (threadIdx.x == 0) {
        mbar_arrive_expect_tx(bar, kRows * SPAN + kN * SPAN);
        tma_load_3d(sa + dst_shift * SPAN, &tmap_a, bar, 0, 0, 0);
        tma_load_2d(sb, &tmap_b, bar, 0, 0);
        mbar_wait(bar, 0);
        tc_fence_after();
        constexpr uint32_t kLayout = SPAN == 128 ? 2u : (SPAN == 64 ? 4u : 6u);
        const uint32_t a_addr = smem_u32(sa) + (dst_shift + shift) * SPAN;
        uint64_t adesc = make_smem_desc(a_addr, 16, 8 * SPAN, kLayout);
        if (base_off_mode == 1) adesc |= static_cast<uint64_t>((a_addr >> 7) & 7) << 49;
        const uint64_t bdesc = make_smem_desc(smem_u32(sb), 16, 8 * SPAN, kLayout);
        constexpr uint32_t idesc = make_idesc(1, 0, 0, 128, kN);
        for (int k = 0; k < SPAN / 32; ++k) umma_ss<false>(tmem, adesc + 2 * k, bdesc + 2 * k, idesc, k != 0);
        umma_commit(done);
    }
    mbar_wait(done, 0);
    tc_fence_after();
    const uint32_t taddr = tmem + (static_cast<uint32_t>(warp * 32) << 16);
    for (int j = 0; j < kN / 32; ++j) {
        uint32_t r[32];
        tmem_ld32(taddr + j * 32, r);
        tmem_ld_wait();
        for (int i = 0; i < 32; ++i) out[(warp * 32 + lane) * kN + j * 32 + i] = __uint_as_float(r[i]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<64>(tmem);
    }
}

static float bf(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }

template <int SPAN>
static int run_span() {
    constexpr int K = SPAN / 2;
    std::vector<__nv_bfloat16> ha(kRows * K), hb(kN * K);
    std::vector<float> fa(kRows * K), fb(kN * K);
    srand(1234 + SPAN);
    for (int i = 0; i < kRows * K; ++i) { fa[i] = bf((rand() % 2001 - 1000) / 500.f); ha[i] = __float2bfloat16_rn(fa[i]); }
    for (int i = 0; i < kN * K; ++i) { fb[i] = bf((rand() % 2001 - 1000) / 500.f); hb[i] = __float2bfloat16_rn(fb[i]); }
    __nv_bfloat16 *da, *db;
    float* dout;
    cudaMalloc(&da, ha.size() * 2);
    cudaMalloc(&db, hb.size() * 2);
    cudaMalloc(&dout, 128 * kN * 4);
    cudaMemcpy(da, ha.data(), ha.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(db, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice);
    CUtensorMap ta, tb;
    {
        // rows split as a 3-D box so that more than 256 rows fit one TMA instruction
        uint64_t dims[3] = {(uint64_t)K, 64, (uint64_t)kRows / 64};
        uint64_t strides[2] = {(uint64_t)SPAN, (uint64_t)SPAN * 64};
        uint32_t box[3] = {(uint32_t)K, 64, (uint32_t)kRows / 64};
        if (encode_tmap(&ta, PG_BF16, 3, da, dims, strides, box, SPAN) != PG_OK) { printf("tmap a failed: %s\n", pggan_last_error()); return 1; }
    }
    {
        uint64_t dims[2] = {(uint64_t)K, (uint64_t)kN};
        uint64_t strides[1] = {(uint64_t)SPAN};
        uint32_t box[2] = {(uint32_t)K, (uint32_t)kN};
        if (encode_tmap(&tb, PG_BF16, 2, db, dims, strides, box, SPAN) != PG_OK) { printf("tmap b failed\n"); return 1; }
    }
    const int smem_bytes = (kRows + 16) * SPAN + 1024 + kN * SPAN + 1024 + 256;
    cudaFuncSetAttribute(probe_kernel<SPAN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    std::vector<float> hout(128 * kN);
    const int shifts[] = {0, 1, 2, 3, 7, 8, 9, 66, 67, 130, 131, 132};
    for (int dst_shift : {0, 1, 5}) {
        for (int mode = 0; mode < 2; ++mode) {
            printf("SPAN %3d dst_shift %d base_off_mode %d :", SPAN, dst_shift, mode);
            for (int shift : shifts) {
                cudaMemset(dout, 0, 128 * kN * 4);
                probe_kernel<SPAN><<<1, 128, smem_bytes>>>(ta, tb, dout, shift, dst_shift, mode);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf(" [cuda error %s]\n", cudaGetErrorString(e)); return 2; }
                cudaMemcpy(hout.data(), dout, hout.size() * 4, cudaMemcpyDeviceToHost);
                double maxerr = 0;
                for (int m = 0; m < 128; ++m)
                    for (int n = 0; n < kN; ++n) {
                        double ref = 0;
                        for (int k = 0; k < K; ++k) ref += (double)fa[(m + shift) * K + k] * fb[n * K + k];
                        maxerr = fmax(maxerr, fabs(ref - hout[m * kN + n]));
                    }
                printf(" s%d:%s", shift, maxerr < 1e-2 ? "ok" : "BAD");
            }
            printf("\n");
        }
    }
    cudaFree(da); cudaFree(db); cudaFree(dout);
    return 0;
}

int main() {
    int rc = run_span<128>();
    if (rc == 0) rc = run_span<64>();
    if (rc == 0) rc = run_span<32>();
    return rc;
}
