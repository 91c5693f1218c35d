// conv3d (k=3, stride 1, zero pad 1) as an implicit GEMM on tcgen05 tensor cores.
//
// Replaces the 16 tensor-core-sized nn.Conv3d calls of ConvModule.forward
// (src/iterseg/unet.py:63-76 used at :93,:96); train-mode BatchNorm statistics
// (unet.py:80-81) are reduced in the epilogue.
//
// Mapping (one CTA per SM, persistent over output tiles):
//   M = 128 output voxels of one z-plane: a patch of Ht rows x P columns, linearised
//       row-major with pitch P (the last 2 columns of every patch row are halo, so
//       Wt = P-2 outputs per row are valid);
//   N = Cout (16..256);  K = 27 taps x Cin, walked as (channel block of CBLK) x (tap).
//   A: ONE TMA box per (tile, channel block): the halo patch {CBLK ch, P, Ht+2, 3 planes}
//      of the channels-last activation tensor, out-of-bounds -> 0 (the conv padding).
//      It lands in smem as rows of CBLK*2 bytes (hardware 128B/64B swizzle); every tap
//      (dz,dy,dx) is the SAME smem tile read through a K-major UMMA descriptor whose
//      start address is advanced by (dz*(Ht+2)*P + dy*P + dx) rows.  So each input
//      voxel is fetched from L2 ~ (3*(Ht+2)*P)/(Ht*Wt) times instead of 27.
//   B: weights packed [tap][Cout][Cin] fp16, one TMA box {CBLK, Cout} per (tap, block),
//      streamed through its own ring.
//   D: fp32 accumulators in TMEM, two stages of Cout columns (epilogue of tile i
//      overlaps the MMAs of tile i+1).
// Warp roles: 0 = A producer, 3 = B producer, 1 = MMA issuer (one thread),
//             2 = TMEM allocator, 4..7 = epilogue (TMEM -> regs -> global + statistics).
#pragma once
#include <cuda.h>

#include "common.cuh"
#include "sm100.cuh"

namespace isg {

struct ConvGeom {
    int N, D, H, W;              // batch (chunks) and spatial extents (in == out)
    int P, Ht, Wt;               // patch pitch, patch rows, valid outputs per row
    int tiles_w, tiles_h;
    int n_tiles;
    int cout;                    // UMMA N (multiple of 16, <= 256)
    int nkb0, nkb1;              // channel blocks taken from source 0 / source 1 (concat)
    int a_rows;                  // 3 * (Ht + 2) * P
    int a_stage_bytes;           // a_rows * CBLK * 2 rounded up to 1024
    int b_stage_bytes;           // taps_per_b * cout * CBLK * 2
    int n_b_stages;
    int n_a_stages;              // 2..4
    int taps_per_b;              // taps per B stage (template parameter G): 1, 3, 9 or 27
    int out_mode;                // 0: fp16 [vox][cout]   1: fp32 [vox][8] (first 8 columns)
    int base_off_mode;           // 0 (correct on B200): descriptor base_offset field = 0;  1: (addr >> 7) & 7
    void *out;
    unsigned long long *stats;   // [N][cout][2] (sum, sum of squares) as 2^-24 fixed point:
                                 // integer atomics are order-independent -> reproducible
};

static constexpr int CONV_THREADS = 256;
static constexpr float STAT_SCALE = 16777216.0f;      // 2^24
static constexpr int CONV_SLACK = 4096;      // garbage rows the last taps of invalid rows touch

__host__ __device__ inline size_t conv_smem_bytes(const ConvGeom &g) {
    return 1024 /* alignment */ + (size_t)g.n_a_stages * g.a_stage_bytes +
           (size_t)g.n_b_stages * g.b_stage_bytes + CONV_SLACK + 512 /* barriers */ +
           4 * 32 * 33 * sizeof(float);
}

template <int CBLK, int G>
__global__ void __launch_bounds__(CONV_THREADS, 1)
conv3d_tc_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                 const __grid_constant__ CUtensorMap tmB, const ConvGeom g) {
    using namespace sm100;
    constexpr uint32_t RB = CBLK * 2;            // smem row bytes (128 -> SW128, 64 -> SW64)
    constexpr int KSTEPS = CBLK / 16;
    extern __shared__ uint8_t smem_dyn[];
    const uint32_t raw_base = smem_u32(smem_dyn);
    const uint32_t pad = ((raw_base + 1023u) & ~1023u) - raw_base;
    uint8_t *base = smem_dyn + pad;
    uint8_t *a_smem = base;
    uint8_t *b_smem = a_smem + (size_t)g.n_a_stages * g.a_stage_bytes;
    uint8_t *tail = b_smem + (size_t)g.n_b_stages * g.b_stage_bytes + CONV_SLACK;
    uint64_t *bars = reinterpret_cast<uint64_t *>(tail);
    uint64_t *a_full = bars, *a_empty = bars + 4, *acc_full = bars + 8, *acc_empty = bars + 10;
    uint64_t *b_full = bars + 12, *b_empty = bars + 12 + 16;        // up to 16 B stages
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 12 + 32);
    float *stat_t = reinterpret_cast<float *>(tail + 512);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int nkb = g.nkb0 + g.nkb1;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmA0);
        prefetch_tmap(&tmA1);
        prefetch_tmap(&tmB);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < g.n_a_stages; ++i) {
            mbar_init(&a_full[i], 1);
            mbar_init(&a_empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&acc_full[i], 1);
            mbar_init(&acc_empty[i], 4);
        }
        for (int i = 0; i < g.n_b_stages; ++i) {
            mbar_init(&b_full[i], 1);
            mbar_init(&b_empty[i], 1);
        }
        fence_barrier_init();
    }
    if (warp == 2) {
        tmem_alloc(tmem_slot, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== A producer =====================
        if (lane == 0) {
            uint32_t it = 0;
            const uint32_t na = (uint32_t)g.n_a_stages;
            for (int tile = blockIdx.x; tile < g.n_tiles; tile += gridDim.x) {
                int t = tile;
                const int wb = t % g.tiles_w; t /= g.tiles_w;
                const int hb = t % g.tiles_h; t /= g.tiles_h;
                const int d = t % g.D;
                const int n = t / g.D;
                for (int kb = 0; kb < nkb; ++kb, ++it) {
                    const uint32_t s = it % na, ph = (it / na) & 1u;
                    mbar_wait(&a_empty[s], ph ^ 1u);
                    mbar_expect_tx(&a_full[s], (uint32_t)g.a_rows * RB);
                    const bool first = kb < g.nkb0;
                    tma_load_5d(a_smem + (size_t)s * g.a_stage_bytes, first ? &tmA0 : &tmA1,
                                &a_full[s], (first ? kb : kb - g.nkb0) * CBLK, wb * g.Wt - 1,
                                hb * g.Ht - 1, d - 1, n);
                }
            }
        }
    } else if (warp == 3) {
        // ===================== B producer =====================
        if (lane == 0) {
            uint32_t it = 0;
            const uint32_t nb = (uint32_t)g.n_b_stages;
            for (int tile = blockIdx.x; tile < g.n_tiles; tile += gridDim.x) {
                for (int kb = 0; kb < nkb; ++kb) {
                    for (int tap = 0; tap < 27; tap += G, ++it) {
                        const uint32_t s = it % nb, ph = (it / nb) & 1u;
                        mbar_wait(&b_empty[s], ph ^ 1u);
                        mbar_expect_tx(&b_full[s], (uint32_t)g.b_stage_bytes);
                        tma_load_3d(b_smem + (size_t)s * g.b_stage_bytes, &tmB, &b_full[s],
                                    kb * CBLK, 0, tap);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        // One thread issues everything, so the instruction count per MMA is what bounds
        // small-N layers: descriptors are (constant high word) | (start address >> 4) and
        // the tap loop is fully unrolled, leaving ~2 integer adds per tcgen05.mma.
        if (lane == 0) {
            const uint32_t idesc = make_idesc_f16(128, (uint32_t)g.cout, 0 /* fp16 */);
            const uint32_t nb = (uint32_t)g.n_b_stages, na = (uint32_t)g.n_a_stages;
            uint32_t ita = 0, itb = 0, tcount = 0;
            const uint64_t dproto = make_kmajor_desc(0, RB, 0);
            const uint32_t d_hi = (uint32_t)(dproto >> 32), d_lo = (uint32_t)dproto;
            constexpr uint32_t U = RB >> 4;                       // one row in 16-byte units
            const uint32_t cz = (uint32_t)((g.Ht + 2) * g.P) * U, cy = (uint32_t)g.P * U;
            const uint32_t b_tap_units = (uint32_t)(g.cout * CBLK * 2) >> 4;
            for (int tile = blockIdx.x; tile < g.n_tiles; tile += gridDim.x, ++tcount) {
                const uint32_t as = tcount & 1u, aph = (tcount >> 1) & 1u;
                mbar_wait(&acc_empty[as], aph ^ 1u);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + as * (uint32_t)g.cout;
                for (int kb = 0; kb < nkb; ++kb, ++ita) {
                    const uint32_t s = ita % na, ph = (ita / na) & 1u;
                    mbar_wait(&a_full[s], ph);
                    const uint32_t a_lo = d_lo | (smem_u32(a_smem + (size_t)s * g.a_stage_bytes) >> 4);
                    uint32_t b_lo = 0, bs = 0;
#pragma unroll
                    for (int tap = 0; tap < 27; ++tap) {
                        if (tap % G == 0) {
                            bs = itb % nb;
                            mbar_wait(&b_full[bs], (itb / nb) & 1u);
                            tc_fence_after();
                            b_lo = d_lo | (smem_u32(b_smem + (size_t)bs * g.b_stage_bytes) >> 4);
                        }
                        const int dz = tap / 9, dy = (tap / 3) % 3, dx = tap % 3;
                        const uint32_t a_tap = a_lo + dz * cz + dy * cy + dx * U;
                        const uint32_t b_tap = b_lo + (tap % G) * b_tap_units;
#pragma unroll
                        for (int k = 0; k < KSTEPS; ++k) {
                            const uint64_t adesc = ((uint64_t)d_hi << 32) | (a_tap + 2 * k);
                            const uint64_t bdesc = ((uint64_t)d_hi << 32) | (b_tap + 2 * k);
                            umma_f16(tmem_d, adesc, bdesc, idesc, (tap | k) != 0 ? 1u : (kb != 0 ? 1u : 0u));
                        }
                        if (tap % G == G - 1) {
                            umma_commit(&b_empty[bs]);
                            ++itb;
                        }
                    }
                    umma_commit(&a_empty[s]);
                }
                umma_commit(&acc_full[as]);
            }
        }
    } else if (warp >= 4) {
        // ===================== epilogue =====================
        const int ew = warp - 4;                     // TMEM lanes 32*ew .. 32*ew+31
        float *st = stat_t + ew * (32 * 33);
        long long csum[8], csq[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) csum[i] = csq[i] = 0;
        int cur_n = -1;
        uint32_t tcount = 0;
        const int row = ew * 32 + lane;
        const int hy = row / g.P, wx = row - hy * g.P;
        auto flush = [&](int n) {
            if (n < 0) return;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int c = i * 32 + lane;
                if (i * 32 < g.cout && c < g.cout) {
                    atomicAdd(g.stats + ((size_t)n * g.cout + c) * 2 + 0, (unsigned long long)csum[i]);
                    atomicAdd(g.stats + ((size_t)n * g.cout + c) * 2 + 1, (unsigned long long)csq[i]);
                }
                csum[i] = csq[i] = 0;
            }
        };
        for (int tile = blockIdx.x; tile < g.n_tiles; tile += gridDim.x, ++tcount) {
            int t = tile;
            const int wb = t % g.tiles_w; t /= g.tiles_w;
            const int hb = t % g.tiles_h; t /= g.tiles_h;
            const int d = t % g.D;
            const int n = t / g.D;
            if (n != cur_n) {
                flush(cur_n);
                cur_n = n;
            }
            const int h = hb * g.Ht + hy, w = wb * g.Wt + wx;
            const bool valid = hy < g.Ht && wx < g.Wt && h < g.H && w < g.W;
            const size_t vox = (((size_t)n * g.D + d) * g.H + h) * g.W + w;
            const uint32_t as = tcount & 1u, aph = (tcount >> 1) & 1u;
            mbar_wait(&acc_full[as], aph);
            tc_fence_after();
            const uint32_t taddr = tmem_base + as * (uint32_t)g.cout + ((uint32_t)(ew * 32) << 16);
            if (g.cout >= 32) {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    if (i * 32 < g.cout) {
                        uint32_t v[32];
                        tmem_ld_32x32(taddr + i * 32, v);
                        tmem_ld_wait();
                        if (valid) {
                            __half *o = reinterpret_cast<__half *>(g.out) + vox * g.cout + i * 32;
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
                        csum[i] += __float2ll_rn(s * STAT_SCALE);
                        csq[i] += __float2ll_rn(q2 * STAT_SCALE);
                    }
                }
            } else {
                uint32_t v[16];
                tmem_ld_32x16(taddr, v);
                tmem_ld_wait();
                if (valid) {
                    float4 *o = reinterpret_cast<float4 *>(reinterpret_cast<float *>(g.out) + vox * 8);
                    o[0] = make_float4(__uint_as_float(v[0]), __uint_as_float(v[1]),
                                       __uint_as_float(v[2]), __uint_as_float(v[3]));
                    o[1] = make_float4(__uint_as_float(v[4]), __uint_as_float(v[5]),
                                       __uint_as_float(v[6]), __uint_as_float(v[7]));
                }
#pragma unroll
                for (int j = 0; j < 16; ++j) st[lane * 33 + j] = valid ? __uint_as_float(v[j]) : 0.0f;
                __syncwarp();
                float s = 0.0f, q2 = 0.0f;
                if (lane < 16) {
#pragma unroll
                    for (int r = 0; r < 32; ++r) {
                        const float x = st[r * 33 + lane];
                        s += x;
                        q2 = fmaf(x, x, q2);
                    }
                }
                __syncwarp();
                csum[0] += __float2ll_rn(s * STAT_SCALE);
                csq[0] += __float2ll_rn(q2 * STAT_SCALE);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[as]);
        }
        flush(cur_n);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, 512);
}

}  // namespace isg
