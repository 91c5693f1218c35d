// Single-thread latencies of the operations the bucket-queue flood is built from.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define N 512
__global__ void k(long long *out, const uint4 *g, uint32_t gn, uint32_t seed) {
    extern __shared__ __align__(16) unsigned char raw[];
    uint32_t *sm = (uint32_t *)raw;
    uint16_t *sh = (uint16_t *)(raw + 16384);
    for (int i = threadIdx.x; i < 4096; i += 32) { sm[i] = (i * 2654435761u + seed) & 4095; }
    for (int i = threadIdx.x; i < 8192; i += 32) { sh[i] = (uint16_t)((i * 40503u + seed) & 8191); }
    __syncwarp();
    if (threadIdx.x != 0) return;
    long long t0, t1;
    uint32_t a = seed & 4095;
    t0 = clock64(); for (int i = 0; i < N; ++i) a = sm[a]; t1 = clock64(); out[0] = (t1 - t0) / N;          // LDS chain
    uint32_t b = seed & 8191;
    t0 = clock64(); for (int i = 0; i < N; ++i) b = sh[b]; t1 = clock64(); out[1] = (t1 - t0) / N;          // LDS.U16 chain
    t0 = clock64(); for (int i = 0; i < N; ++i) { atomicOr(sm + a, 0u); a = sm[a]; } t1 = clock64(); out[2] = (t1 - t0) / N;  // ATOMS + LDS same word
    t0 = clock64(); for (int i = 0; i < N; ++i) { atomicOr(sm + ((a + i * 7) & 4095), 0u); } t1 = clock64(); out[3] = (t1 - t0) / N;  // ATOMS issue
    uint32_t c = a | 1;
    t0 = clock64(); for (int i = 0; i < N; ++i) c = (__clz(c) + 0x10001u) * (c | 3); t1 = clock64(); out[4] = (t1 - t0) / N;  // clz + imad chain
    t0 = clock64(); for (int i = 0; i < N; ++i) { sm[a] = a; a = sm[a] ^ 0; a = (a + 1) & 4095; } t1 = clock64(); out[5] = (t1 - t0) / N;  // STS -> LDS same addr
    // global: dependent chain over records (L2 resident)
    uint32_t gi = seed % gn;
    t0 = clock64(); for (int i = 0; i < N; ++i) { uint4 r = __ldg(g + 2 * (size_t)gi); gi = r.x % gn; } t1 = clock64(); out[6] = (t1 - t0) / N;
    // prefetch issue cost
    t0 = clock64(); for (int i = 0; i < N; ++i) { asm volatile("prefetch.global.L1 [%0];" ::"l"(g + 2 * (size_t)((gi + i * 977u) % gn))); } t1 = clock64(); out[7] = (t1 - t0) / N;
    // prefetch then (much later) load: L1 hit?
    uint32_t idx[64];
    for (int i = 0; i < 64; ++i) idx[i] = (gi * 31u + i * 7919u + 5) % gn;
    for (int i = 0; i < 64; ++i) asm volatile("prefetch.global.L1 [%0];" ::"l"(g + 2 * (size_t)idx[i]));
    for (int i = 0; i < 200; ++i) a = sm[a];
    uint32_t acc = 0;
    t0 = clock64(); for (int i = 0; i < 64; ++i) { uint4 r = __ldg(g + 2 * (size_t)idx[i]); acc += r.x; if (acc == 0x7fffffff) idx[(i + 1) & 63] = 0; } t1 = clock64(); out[8] = (t1 - t0) / 64;
    // same without prefetch
    for (int i = 0; i < 64; ++i) idx[i] = (gi * 17u + i * 104729u + 11) % gn;
    t0 = clock64(); for (int i = 0; i < 64; ++i) { uint4 r = __ldg(g + 2 * (size_t)idx[i]); acc += r.x; if (acc == 0x7fffffff) idx[(i + 1) & 63] = 0; } t1 = clock64(); out[9] = (t1 - t0) / 64;
    // ld.global.ca (plain) after a plain load of the same line (L1 hit latency)
    t0 = clock64(); for (int i = 0; i < 64; ++i) { uint4 r = __ldg(g + 2 * (size_t)idx[i]); acc += r.x; if (acc == 0x7fffffff) idx[(i + 1) & 63] = 0; } t1 = clock64(); out[10] = (t1 - t0) / 64;
    out[15] = a + b + c + gi + acc;
}
int main() {
    long long *d, h[16];
    cudaMalloc(&d, sizeof(h));
    const uint32_t gn = 563194;
    uint4 *g; cudaMalloc(&g, (size_t)gn * 32);
    uint32_t *hg = (uint32_t *)malloc((size_t)gn * 32);
    for (size_t i = 0; i < (size_t)gn * 8; ++i) hg[i] = (uint32_t)(i * 2654435761u >> 3);
    cudaMemcpy(g, hg, (size_t)gn * 32, cudaMemcpyHostToDevice);
    for (int smem : {32768, 232448}) {
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        k<<<1, 32, smem>>>(d, g, gn, 12345u);
        k<<<1, 32, smem>>>(d, g, gn, 777u);
        cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
        const char *nm[] = {"LDS chain", "LDS.U16 chain", "ATOMS.OR + LDS same word", "ATOMS.OR issue", "clz+imad chain", "STS->LDS same addr (+2 alu)",
                            "LDG.128 chain (L2)", "prefetch.L1 issue", "LDG after prefetch", "LDG cold (L2)", "LDG again (L1)"};
        printf("dynamic smem %d: %s\n", smem, cudaGetErrorString(cudaGetLastError()));
        for (int i = 0; i < 11; ++i) printf("  %-30s %lld clk\n", nm[i], h[i]);
    }
    return 0;
}
