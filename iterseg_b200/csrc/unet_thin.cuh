// c0.conv0 (1 -> 32 channels, src/iterseg/unet.py:93 of the first ConvModule) on tcgen05 tensor
// cores, straight from the frame.
//
// One input channel is too thin for the TMA-fed implicit GEMM of unet_conv.cuh (a K step of 16
// would be 1/16 full), and on CUDA cores the layer costs 27*32 FMAs per voxel (1.7 ms per
// frame, profiles/r01_notes.md).  Here the threads build an explicit im2col tile in shared
// memory and one thread issues two MMAs per 128 voxels:
//   unit : 128 output voxels = 4 rows x 32 columns of one z-plane of one chunk (all valid
//          outputs; the 3 x 6 x 34 input halo is staged once in shared memory as fp16),
//   A    : [128 voxels][K = 27 taps + 5 zero] fp16, K-major, 64-byte rows, built by the 128
//          threads (thread r writes row r) in the 64-byte-swizzled layout the UMMA descriptor
//          expects (16-byte chunk index XOR address bits 7-8 -- the operand fetch swizzles on
//          absolute shared-memory address bits, profiles/r01_notes.md),
//   B    : the weights [32 out][K] fp16, converted into the same layout once per CTA,
//   D    : 32 TMEM columns per CTA; thread r reads row r back (tcgen05.ld 32x32b), writes the
//          raw convolution output and adds its tile's BatchNorm sums (per-tile fp32 sums in a
//          fixed order -> 2^-24 fixed point -> integer atomics: reproducible).
// Seven CTAs share an SM (29 KB of shared memory each), which is what overlaps the
// stage / build / MMA / epilogue phases of different units.
// (The same scheme was tried for c8_0.conv1, 5 -> 5 with K = 27 x 8: its 56 KB im2col tile
// leaves 3 CTAs per SM and it ran at 1.8 ms against 1.3 ms on CUDA cores -- not kept.)
#pragma once
#include <cuda_fp16.h>

#include "common.cuh"
#include "sm100.cuh"
#include "unet_elem.cuh"

namespace isg {

static constexpr int THIN_THREADS = 128;
static constexpr int THIN_HALO = 3 * 6 * 34;          // staged input voxels per unit

__host__ __device__ constexpr size_t thin_smem_bytes() {
    return 1024 /* alignment */ + 8192 /* A */ + 32 * 64 /* B */ + ((THIN_HALO * 2 + 15) & ~15) /* halo */ +
           4 * 32 * 33 * sizeof(float) /* statistics transposes */ + 16 /* barrier, TMEM slot */;
}

struct ThinArgs {
    const float *src;                     // frame (Z,Y,X) fp32 (device memory, or pinned host memory mapped
                                          // into the device address space: "zero-copy" staging)
    const float *norm_max;                // nullable: every input voxel is divided by norm_max[0] on its way in
                                          // (vol /= max, segmentation.py:889; IEEE division, bit-identical to
                                          // isg_frame_divide_by_max followed by a plain load)
    const int *starts;                    // [N][3] chunk origins
    int Y, X;
    const float *wgt;                     // [27][32] fp32
    __half *out;                          // [N][vox][32] raw convolution output
    unsigned long long *stats;            // [N][32][2]
    int N, D, H, W;                       // chunks and chunk extents
};

__global__ void __launch_bounds__(THIN_THREADS)
conv_in_tc_kernel(const ThinArgs a) {
    using namespace sm100;
    extern __shared__ uint8_t smem_dyn[];
    const uint32_t raw_base = smem_u32(smem_dyn);
    uint8_t *a_smem = smem_dyn + (((raw_base + 1023u) & ~1023u) - raw_base);      // [128][64 B]
    uint8_t *b_smem = a_smem + 8192;                                              // [32][64 B]
    unsigned short *hin = reinterpret_cast<unsigned short *>(b_smem + 2048);      // [3][6][34] fp16 bits
    float *stat_t = reinterpret_cast<float *>(reinterpret_cast<uint8_t *>(hin) + ((THIN_HALO * 2 + 15) & ~15));
    uint64_t *bar = reinterpret_cast<uint64_t *>(stat_t + 4 * 32 * 33);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bar + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int D = a.D, H = a.H, W = a.W;
    const size_t vox = (size_t)D * H * W;
    const int tiles_w = (W + 31) / 32, tiles_h = (H + 3) / 4;
    const uint32_t units = (uint32_t)a.N * D * tiles_h * tiles_w;       // < 2^31, checked by the launcher

    if (tid == 32) {
        mbar_init(bar, 1);
        fence_barrier_init();
    }
    if (warp == 0) {
        tmem_alloc(tmem_slot, 32);
        tmem_relinquish();
    }
    {   // weights -> B[co][k = tap] (fp16, swizzled): thread t fills one 16-byte chunk
        const int co = tid >> 2, c = tid & 3;
        uint4 pk;
        __half *h = reinterpret_cast<__half *>(&pk);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int tap = c * 8 + e;
            h[e] = __float2half_rn(tap < 27 ? a.wgt[tap * 32 + co] : 0.0f);
        }
        *reinterpret_cast<uint4 *>(b_smem + co * 64 + ((c ^ ((co >> 1) & 3)) << 4)) = pk;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
    const uint32_t idesc = make_idesc_f16(128, 32, 0 /* fp16 */);
    const uint64_t adesc = make_kmajor_desc(smem_u32(a_smem), 64, 0);
    const uint64_t bdesc = make_kmajor_desc(smem_u32(b_smem), 64, 0);

    // halo slots of this thread: i = tid + 128 * j  ->  (iz, iy, ix), the same for every unit
    constexpr int HSLOTS = (THIN_HALO + THIN_THREADS - 1) / THIN_THREADS;
    int h_off[HSLOTS];                                  // iz | iy << 8 | ix << 16, or -1
#pragma unroll
    for (int j = 0; j < HSLOTS; ++j) {
        const int i = tid + THIN_THREADS * j;
        h_off[j] = i < THIN_HALO ? ((i / (34 * 6)) | (((i / 34) % 6) << 8) | ((i % 34) << 16)) : -1;
    }
    const int hy = tid >> 5, wx = tid & 31;             // this thread's output voxel inside the unit
    const uint32_t sw = (uint32_t)((tid >> 1) & 3);     // swizzle term of A row `tid`
    float *st = stat_t + warp * (32 * 33);
    long long csum = 0, csq = 0;                        // channel `lane` of the current chunk
    int cur_n = -1;
    uint32_t phase = 0;
    auto flush = [&](int n) {
        if (n < 0) return;
        atomicAdd(a.stats + ((size_t)n * 32 + lane) * 2 + 0, (unsigned long long)csum);
        atomicAdd(a.stats + ((size_t)n * 32 + lane) * 2 + 1, (unsigned long long)csq);
        csum = csq = 0;
    };

    // unit index -> (chunk, plane, tile origin)
    auto decode = [&](uint32_t u, int &n, int &d, int &h0, int &w0) {
        uint32_t t = u;
        w0 = (int)(t % (uint32_t)tiles_w) * 32; t /= (uint32_t)tiles_w;
        h0 = (int)(t % (uint32_t)tiles_h) * 4; t /= (uint32_t)tiles_h;
        d = (int)(t % (uint32_t)D);
        n = (int)(t / (uint32_t)D);
    };
    // stage the input halo of unit u (zero outside the CHUNK: the reference pads the sliced chunk)
    auto stage = [&](uint32_t u) {
        int n, d, h0, w0;
        decode(u, n, d, h0, w0);
        const int z0 = __ldg(a.starts + n * 3 + 0), y0 = __ldg(a.starts + n * 3 + 1), x0 = __ldg(a.starts + n * 3 + 2);
        const float dmax = a.norm_max ? __ldg(a.norm_max) : 1.0f;
        float v[HSLOTS];
#pragma unroll
        for (int j = 0; j < HSLOTS; ++j) {
            const int dd = d + (h_off[j] & 255) - 1, hh = h0 + ((h_off[j] >> 8) & 255) - 1, ww = w0 + (h_off[j] >> 16) - 1;
            v[j] = 0.0f;
            if (h_off[j] >= 0 && dd >= 0 && dd < D && hh >= 0 && hh < H && ww >= 0 && ww < W)
                v[j] = __ldg(a.src + ((size_t)(z0 + dd) * a.Y + (y0 + hh)) * a.X + (x0 + ww));
        }
        if (a.norm_max) {
#pragma unroll
            for (int j = 0; j < HSLOTS; ++j) v[j] = __fdiv_rn(v[j], dmax);
        }
#pragma unroll
        for (int j = 0; j < HSLOTS; ++j)
            if (h_off[j] >= 0) hin[tid + THIN_THREADS * j] = __half_as_ushort(__float2half_rn(v[j]));
    };

    if (blockIdx.x < units) stage(blockIdx.x);
    for (uint32_t u = blockIdx.x; u < units; u += gridDim.x) {
        int n, d, h0, w0;
        decode(u, n, d, h0, w0);
        if (n != cur_n) {
            flush(cur_n);
            cur_n = n;
        }
        __syncthreads();                                 // the staged halo of this unit is complete
        // ---- im2col: row `tid` of A = the 27 taps around this thread's voxel ----
        uint32_t w32[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const int t0 = 2 * k, t1 = 2 * k + 1;
            uint32_t lo = 0, hi = 0;
            if (t0 < 27) lo = hin[((t0 / 9) * 6 + hy + (t0 / 3) % 3) * 34 + wx + t0 % 3];
            if (t1 < 27) hi = hin[((t1 / 9) * 6 + hy + (t1 / 3) % 3) * 34 + wx + t1 % 3];
            w32[k] = lo | (hi << 16);
        }
#pragma unroll
        for (int c = 0; c < 4; ++c)
            *reinterpret_cast<uint4 *>(a_smem + tid * 64 + ((c ^ sw) << 4)) =
                make_uint4(w32[4 * c], w32[4 * c + 1], w32[4 * c + 2], w32[4 * c + 3]);
        fence_proxy_async();                             // generic-proxy writes -> visible to the MMA's reads
        tc_fence_before();                               // orders the previous unit's TMEM reads
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
            umma_f16(tmem_base, adesc, bdesc, idesc, 0u);
            umma_f16(tmem_base, adesc + 2, bdesc + 2, idesc, 1u);
            umma_commit(bar);
        }
        // the halo tile is free again: stage the next unit's while the MMAs run
        if (u + gridDim.x < units) stage(u + gridDim.x);
        mbar_wait(bar, phase);
        phase ^= 1u;
        tc_fence_after();
        // ---- epilogue: row `tid` of D ----
        const int h = h0 + hy, w = w0 + wx;
        const bool valid = h < H && w < W;
        uint32_t v[32];
        tmem_ld_32x32(taddr, v);
        tmem_ld_wait();
        if (valid) {
            __half *o = a.out + ((size_t)n * vox + ((size_t)d * H + h) * W + w) * 32;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                uint4 pk;
                uint32_t *pw = reinterpret_cast<uint32_t *>(&pk);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    __half2 h2 = __floats2half2_rn(__uint_as_float(v[q * 8 + e * 2]),
                                                   __uint_as_float(v[q * 8 + e * 2 + 1]));
                    pw[e] = *reinterpret_cast<uint32_t *>(&h2);
                }
                reinterpret_cast<uint4 *>(o)[q] = pk;
            }
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) st[lane * 33 + j] = valid ? __uint_as_float(v[j]) : 0.0f;
        __syncwarp();
        float s = 0.0f, q2 = 0.0f;
#pragma unroll
        for (int r = 0; r < 32; ++r) {
            const float x = st[r * 33 + lane];
            s += x;
            q2 = fmaf(x, x, q2);
        }
        __syncwarp();
        stat_guard(q2);
        csum += __float2ll_rn(s * 16777216.0f);
        csq += __float2ll_rn(q2 * 16777216.0f);
    }
    flush(cur_n);
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, 32);
}

}  // namespace isg
