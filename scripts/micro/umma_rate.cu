// Issue rate of tcgen05.mma (cta_group::1, kind::f16, M=128, K=16) from shared memory:
// clocks per MMA versus N, swizzle mode (64 B / 128 B rows) and row shift of the A start address.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../iterseg_b200/csrc/sm100.cuh"
using namespace isg::sm100;

__global__ void __launch_bounds__(128, 1) k(long long *out, int N, int rb, int shift_rows, int nmma, int nacc, int mode) {
    extern __shared__ uint8_t smem_dyn[];
    const uint32_t raw = smem_u32(smem_dyn);
    uint8_t *base = smem_dyn + (((raw + 1023u) & ~1023u) - raw);
    __shared__ uint64_t bar;
    __shared__ uint64_t bar2[4];
    __shared__ uint32_t tslot;
    for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(base)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); for (int i = 0; i < 4; ++i) mbar_init(&bar2[i], 1); fence_barrier_init(); }
    if (threadIdx.x < 32) { tmem_alloc(&tslot, 512); tmem_relinquish(); }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tslot;
    if (threadIdx.x == 0) {
        const uint32_t idesc = make_idesc_f16(128, (uint32_t)N, 0);
        const uint64_t proto = make_kmajor_desc(0, (uint32_t)rb, 0);
        const uint32_t hi = (uint32_t)(proto >> 32), lo = (uint32_t)proto;
        const uint32_t a0 = lo | (smem_u32(base) >> 4);
        const uint32_t b0 = lo | (smem_u32(base + 96 * 1024) >> 4);
        const uint32_t U = (uint32_t)rb >> 4;
        const uint32_t shU = (uint32_t)shift_rows * U;
        const uint32_t acc1 = nacc > 1 ? (uint32_t)N : 0u;
        long long t0 = clock64();
        for (int i = 0; i < nmma; i += 18) {
#pragma unroll
            for (int j = 0; j < 18; ++j) {
                const uint32_t sh = (uint32_t)(j % 9) * shU;                     // tap-like row shifts
                const uint32_t ksel = (uint32_t)(j & 1) * 2u;
                const uint64_t ad = ((uint64_t)hi << 32) | (a0 + sh + ksel);
                const uint64_t bd = ((uint64_t)hi << 32) | (b0 + ksel);
                umma_f16(tmem + (j & 1) * acc1, ad, bd, idesc, 1u);
            }
            if (mode & 1) { umma_commit(&bar2[0]); umma_commit(&bar2[1]); }       // two commits per 18 MMAs
            if (mode & 2) { mbar_wait(&bar2[2], 1u); tc_fence_after(); }          // a wait that passes at once
            if (mode & 4) { mbar_wait(&bar2[0], (uint32_t)(i / 18) & 1u); tc_fence_after(); }   // wait for this tile's commit
        }
        umma_commit(&bar);
        long long t1 = clock64();
        mbar_wait(&bar, 0);
        long long t2 = clock64();
        out[0] = t1 - t0;
        out[1] = t2 - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(tmem, 512);
}

int main() {
    long long *d, h[2];
    cudaMalloc(&d, sizeof(h));
    const int smem = 200 * 1024;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const int nmma = 4608;
    printf("%5s %4s %6s %5s | %10s %10s\n", "N", "rb", "shift", "nacc", "issue clk", "done clk/MMA");
    for (int rb : {64})
        for (int N : {32, 48, 96, 192})
            for (int shift : {1})
                for (int nacc : {1, 2, 10, 11, 12, 13}) {
                    const int mode = nacc >= 10 ? nacc - 9 : 0;     // 10: commits, 11: free wait, 12: both, 13: dependent wait
                    const int na = nacc >= 10 ? 2 : nacc;
                    if (na * N > 512) continue;
                    k<<<1, 128, smem>>>(d, N, rb, shift, nmma, na, mode == 4 ? 5 : mode);
                    cudaError_t e = cudaDeviceSynchronize();
                    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
                    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
                    printf("%5d %4d %6d %5d | %10.1f %10.1f\n", N, rb, shift, nacc, (double)h[0] / nmma, (double)h[1] / nmma);
                }
    return 0;
}
