// Probe: what paces back-to-back tcgen05.mma (SS, M = 128, K = 16, bf16)?  One thread issues `iters` MMAs on fixed
// smem operands and accumulators; variants: N, number of independent accumulators, one or two issuing warps.
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o tools/probe_issue tools/probe_issue.cu pggan_b200/csrc/common.cu
#include "../pggan_b200/csrc/common.cuh"
#include <vector>
using namespace pg;

__global__ void __launch_bounds__(128, 1)
issue_kernel(long long* out, int N, int nacc, int issuers, int iters, int a_rows_shift, int use_elect) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* done = reinterpret_cast<uint64_t*>(smem + 65536);
    uint32_t* slot = reinterpret_cast<uint32_t*>(done + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 65536 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (threadIdx.x == 0) {
        mbar_init(&done[0], 1);
        mbar_init(&done[1], 1);
        fence_mbar_init();
    }
    if (warp == 2) tmem_alloc<512>(slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *slot;
    if (warp < issuers && (use_elect ? elect_one() : (lane == 0))) {
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (uint32_t(N >> 3) << 17) | (uint32_t(128 >> 4) << 24);
        const uint64_t adesc = make_smem_desc(smem_u32(smem), 16, 1024, 2);
        const uint64_t bdesc = make_smem_desc(smem_u32(smem) + 32768, 16, 1024, 2);
        const long long t0 = clock64();
        const uint32_t d0 = tmem + ((warp * nacc) * N) % 512, d1 = tmem + ((warp * nacc + (nacc - 1)) * N) % 512;
        // a_rows_shift != 0: every MMA reads a DIFFERENT 4 KB A tile (walks 8 tiles) — does the A fetch set the pace?
        for (int i = 0; i < iters; i += 2) {
            const uint64_t ao = a_rows_shift ? static_cast<uint64_t>(((i >> 1) & 7) * 256) : 0ull;
            umma_ss<false>(d0, adesc + ao, bdesc, idesc, 1);
            umma_ss<false>(d1, adesc + ao + 2, bdesc + 2, idesc, 1);
        }
        const long long t1 = clock64();
        umma_commit(&done[warp]);
        mbar_wait(&done[warp], 0);
        const long long t2 = clock64();
        out[blockIdx.x * 4 + warp * 2 + 0] = t1 - t0;
        out[blockIdx.x * 4 + warp * 2 + 1] = t2 - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc<512>(tmem);
    }
}

int main() {
    long long* dout;
    cudaMalloc(&dout, 148 * 4 * 8);
    cudaFuncSetAttribute(issue_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 70000);
    const int iters = 2000;
    for (int use_elect : {1})
    for (int walk : {0, 1})
    for (int grid : {148})
        for (int N : {16, 64, 128, 256})
            for (int issuers : {1, 2})
                for (int nacc : {1, 2}) {
                    if (issuers * nacc * N > 512) continue;
                    cudaMemset(dout, 0, 148 * 4 * 8);
                    issue_kernel<<<grid, 128, 70000>>>(dout, N, nacc, issuers, iters, walk, use_elect);
                    cudaError_t e = cudaDeviceSynchronize();
                    if (e != cudaSuccess) { printf("cuda error %s\n", cudaGetErrorString(e)); return 1; }
                    long long h[4];
                    cudaMemcpy(h, dout, sizeof(h), cudaMemcpyDeviceToHost);
                    printf("walkA %d elect %d grid %3d N %3d issuers %d nacc %d : issue %.1f cyc/MMA, complete %.1f cyc/MMA (per issuer; x%d MMAs in flight total)\n",
                           walk, use_elect, grid, N, issuers, nacc, (double)h[0] / iters, (double)h[1] / iters, issuers);
                }
    return 0;
}
