// Latency of the warp primitives the flood heap is built from (one warp, dependent chains).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define N 256
__global__ void k(long long *out, uint32_t seed) {
    __shared__ uint64_t sm[1024];
    const unsigned lane = threadIdx.x;
    for (int i = lane; i < 1024; i += 32) sm[i] = (uint64_t)(i * 2654435761u + seed) % 1024;
    __syncwarp();
    uint32_t x = lane * 7919u + seed;
    long long t0, t1;
    // 1. REDUX min chain
    t0 = clock64();
    for (int i = 0; i < N; ++i) x = __reduce_min_sync(0xFFFFFFFFu, x + lane) + i;
    t1 = clock64();
    if (lane == 0) out[0] = (t1 - t0) / N;
    // 2. ballot + ffs chain
    t0 = clock64();
    for (int i = 0; i < N; ++i) x = __ffs(__ballot_sync(0xFFFFFFFFu, ((x + lane) & 3) == 0)) + x;
    t1 = clock64();
    if (lane == 0) out[1] = (t1 - t0) / N;
    // 3. shfl chain
    t0 = clock64();
    for (int i = 0; i < N; ++i) x = __shfl_sync(0xFFFFFFFFu, x, (x + i) & 31) + 1;
    t1 = clock64();
    if (lane == 0) out[2] = (t1 - t0) / N;
    // 4. LDS.64 dependent chain
    uint64_t a = x & 1023;
    t0 = clock64();
    for (int i = 0; i < N; ++i) a = sm[(a + lane) & 1023];
    t1 = clock64();
    if (lane == 0) out[3] = (t1 - t0) / N;
    // 5. STS + syncwarp + LDS (lane 0 writes, all read)
    t0 = clock64();
    for (int i = 0; i < N; ++i) {
        if (lane == 0) sm[(a + i) & 1023] = a + i;
        __syncwarp();
        a = sm[(a + i) & 1023] + lane;
        __syncwarp();
    }
    t1 = clock64();
    if (lane == 0) out[4] = (t1 - t0) / N;
    // 6. 64-bit min via 2 REDUX + ballot (one heap level without memory)
    uint64_t kk = ((uint64_t)x << 32) | lane;
    t0 = clock64();
    for (int i = 0; i < N; ++i) {
        uint32_t hi = (uint32_t)(kk >> 32), mhi = __reduce_min_sync(0xFFFFFFFFu, hi);
        uint32_t lo = hi == mhi ? (uint32_t)kk : 0xFFFFFFFFu, mlo = __reduce_min_sync(0xFFFFFFFFu, lo);
        uint32_t win = __ffs(__ballot_sync(0xFFFFFFFFu, hi == mhi && lo == mlo)) - 1;
        kk = (((uint64_t)mhi << 32) | mlo) + win + lane * 977u + ((uint64_t)(lane ^ i) << 33);
    }
    t1 = clock64();
    if (lane == 0) out[5] = (t1 - t0) / N;
    // 7. 64-bit min via 5-step shfl butterfly
    t0 = clock64();
    for (int i = 0; i < N; ++i) {
        uint64_t m = kk;
        for (int o = 16; o > 0; o >>= 1) { uint64_t y = __shfl_xor_sync(0xFFFFFFFFu, m, o); m = y < m ? y : m; }
        kk = m + lane * 977u + ((uint64_t)(lane ^ i) << 33);
    }
    t1 = clock64();
    if (lane == 0) out[6] = (t1 - t0) / N;
    // 8. match_any
    t0 = clock64();
    for (int i = 0; i < N; ++i) x = __match_any_sync(0xFFFFFFFFu, (x + lane) & 7) + x;
    t1 = clock64();
    if (lane == 0) out[7] = (t1 - t0) / N;
    if (lane == 0) out[15] = x + a + kk;
}
int main() {
    long long *d, h[16];
    cudaMalloc(&d, sizeof(h));
    k<<<1, 32>>>(d, 12345u);
    k<<<1, 32>>>(d, 777u);
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    const char *nm[] = {"redux.min", "ballot+ffs", "shfl", "lds64 chain", "sts+syncwarp+lds+syncwarp", "min64 = 2 redux + ballot",
                        "min64 shfl butterfly", "match_any"};
    for (int i = 0; i < 8; ++i) printf("%-28s %lld clk\n", nm[i], h[i]);
    return 0;
}
