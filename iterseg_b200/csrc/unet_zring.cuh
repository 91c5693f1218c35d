// c8_0.conv0: conv3d 64 -> 5 (k=3, zero pad 1) on the concatenation [up3 (32 ch), c0 skip (32 ch)]
// (src/iterseg/unet.py:93 of the last ConvModule; the cat is unet.py:345).
//
// With 5 output channels the generic kernel (unet_conv.cuh, dx-fold, N = 48) is bound by the
// NUMBER of tcgen05.mma it issues -- 36 per 128 voxels at ~70 clocks each whatever their N --
// and ran at 238 TFLOP/s, 1.75 ms per frame.  This kernel folds the NINE (dz,dx) taps into N:
//   D_p[r, (dz,dx,co)] = sum_{dy,ci} A_p[r + dy*P, ci] * W[dz,dy,dx][ci,co]      (input plane p)
//   out_z[r, co]       = sum_{dz,dx} D_{z+dz-1}[r + dx, (dz,dx,co)]
// i.e. one accumulator per INPUT plane (N = 9 taps x 8 channels + 8 zero rows = 80 columns,
// 12 MMAs per plane: 2 channel blocks x 3 dy x 2 K steps), kept in a RING of 3 accumulators (256 TMEM
// columns, so TWO CTAs share an SM and fill each other's MMA -> epilogue bubbles) while
// the CTA walks a z-column of the chunk: output plane z is emitted once the accumulators of
// planes z-1, z, z+1 are complete, reading the dz = 0 / 1 / 2 column blocks of the three, and an
// accumulator is released after the third output plane that reads it.  The dx shift is the same
// warp-shuffle row shift as in the dx-fold epilogue (a patch row is one warp, P = 32).
// 36 -> 12 MMAs per 128 voxels: 1.75 -> 1.07 ms (one CTA per SM, ring of 6) -> 0.95 ms (two per SM).
//
// A: TMA halo planes {32 ch, P = 32, Ht + 2 = 6} (64-byte rows, hardware swizzle) through a ring
//    of 4 slots, one per (plane, channel block).  B: the whole weight tensor resident in shared
//    memory, packed [dy][(dz,dx) x 8 + co][cin] fp16 (pack_conv_w_zring_kernel).
// Warp roles: 0 = A producer + column scheduler, 1 = MMA issuer, 2 = TMEM allocator,
//             3 = B loader, 4..7 = epilogue.  Columns are handed out dynamically (see unet_conv.cuh).
#pragma once
#include <cuda.h>

#include "common.cuh"
#include "sm100.cuh"
#include "unet_elem.cuh"

namespace isg {

struct ZringGeom {
    int N, D, H, W;              // chunks, chunk extents
    int tiles_w, tiles_h;
    int n_cols;                  // N * tiles_h * tiles_w z-columns
    __half *out;                 // fp16 [N][vox][8] raw (pre-BatchNorm) output: one 16-byte store per voxel
    unsigned long long *stats;   // [N][16][2] fixed point (see unet_conv.cuh)
    unsigned int *sched;         // column counter, zeroed by the caller
};

static constexpr int ZR_P = 32, ZR_HT = 4, ZR_WT = 30;
static constexpr int ZR_PLANE_ROWS = (ZR_HT + 2) * ZR_P;              // 192
static constexpr int ZR_PLANE_BYTES = ZR_PLANE_ROWS * 64;             // 12288
#ifndef ISG_ZR_SLOTS
#define ISG_ZR_SLOTS 4
#endif
static constexpr int ZR_SLOTS = ISG_ZR_SLOTS;                         // plane slots per CTA (TMA prefetch depth)
static constexpr int ZR_NA = 3;                                       // accumulators in the ring: 3 x 80 columns ->
                                                                      // 256 TMEM columns, TWO CTAs per SM (they fill
                                                                      // each other's MMA -> epilogue bubbles)
static constexpr int ZR_TMEM_COLS = 256;
static constexpr int ZR_N = 80;                                       // UMMA N
static constexpr int ZR_B_STAGE = ZR_N * 64;                          // one (channel block, dy): 5120 B
static constexpr int ZR_SCHED = 4;
static constexpr int ZR_THREADS = 256;

__host__ __device__ constexpr size_t zring_smem_bytes() {
    return 1024 + (size_t)ZR_SLOTS * ZR_PLANE_BYTES + 6 * ZR_B_STAGE + 512 /* barriers */ +
           4 * 32 * 9 * sizeof(float);
}

// Epilogue of the z-ring kernels (4 warps, warp ew = TMEM lanes 32*ew.. = patch row ew): walks the
// columns published by the scheduler; output plane z = sum over the input planes z-1, z, z+1 of
// their (dz, dx) column blocks, dx by warp shuffle; fp16 [vox][8] store + BatchNorm sums (of the fp32 values).
// NA accumulators of ZR_N columns in the ring; statistics laid out [N][SSTRIDE][2].
template <int NA, int SSTRIDE, bool BATCH_LD>
__device__ __forceinline__ void zring_epilogue(const ZringGeom &g, uint32_t tmem_base, uint64_t *acc_full,
                                               uint64_t *acc_empty, uint64_t *sched_full, uint64_t *sched_empty,
                                               volatile int *sched_col, float *st, int ew, int lane) {
    using namespace sm100;
    const int D = g.D;
    long long csum = 0, csq = 0;                     // lane < 8: channel `lane` of the current chunk
    int cur_n = -1;
    auto flush = [&](int n) {
        if (n >= 0 && lane < 5) {
            atomicAdd(g.stats + ((size_t)n * SSTRIDE + lane) * 2 + 0, (unsigned long long)csum);
            atomicAdd(g.stats + ((size_t)n * SSTRIDE + lane) * 2 + 1, (unsigned long long)csq);
        }
        csum = csq = 0;
    };
    const size_t vox_chunk = (size_t)D * g.H * g.W;
    uint32_t pc0 = 0;
    for (uint32_t sidx = 0;; ++sidx) {
        const uint32_t slot = sidx % ZR_SCHED;
        mbar_wait(&sched_full[slot], (sidx / ZR_SCHED) & 1u);
        const int col = sched_col[slot];
        __syncwarp();
        if (lane == 0) mbar_arrive(&sched_empty[slot]);
        if (col >= g.n_cols) break;
        const int wb = col % g.tiles_w;
        const int hb = (col / g.tiles_w) % g.tiles_h;
        const int n = col / (g.tiles_w * g.tiles_h);
        if (n != cur_n) {
            flush(cur_n);
            cur_n = n;
        }
        const int h = hb * ZR_HT + ew, w = wb * ZR_WT + lane;
        const bool valid = lane < ZR_WT && h < g.H && w < g.W;
        int waited = 0;
        for (int z = 0; z < D; ++z) {
            const int need = z + 1 < D ? z + 1 : D - 1;
            while (waited <= need) {
                const uint32_t q = pc0 + (uint32_t)waited;
                mbar_wait(&acc_full[q % NA], (q / NA) & 1u);
                ++waited;
            }
            tc_fence_after();
            float S[3][8];
            if (BATCH_LD) {
                // all loads are issued before the one wait (72 registers)
                uint32_t v[3][3][8];
#pragma unroll
                for (int dz = 0; dz < 3; ++dz) {
                    const int p = z + dz - 1;
                    if (p < 0 || p >= D) continue;                      // zero padding in z
                    const uint32_t a = (pc0 + (uint32_t)p) % NA;
                    const uint32_t taddr = tmem_base + a * ZR_N + dz * 24 + ((uint32_t)(ew * 32) << 16);
                    tmem_ld_32x8(taddr, v[dz][0]);
                    tmem_ld_32x8(taddr + 8, v[dz][1]);
                    tmem_ld_32x8(taddr + 16, v[dz][2]);
                }
                tmem_ld_wait();
#pragma unroll
                for (int dx = 0; dx < 3; ++dx)
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        float acc = __uint_as_float(v[1][dx][j]);       // plane z always exists
                        if (z >= 1) acc += __uint_as_float(v[0][dx][j]);
                        if (z + 1 < D) acc += __uint_as_float(v[2][dx][j]);
                        S[dx][j] = acc;
                    }
            } else {
                // one plane at a time (the variant that shares an SM with a second CTA: <= 112 registers)
#pragma unroll
                for (int dx = 0; dx < 3; ++dx)
#pragma unroll
                    for (int j = 0; j < 8; ++j) S[dx][j] = 0.0f;
#pragma unroll
                for (int dz = 0; dz < 3; ++dz) {
                    const int p = z + dz - 1;
                    if (p < 0 || p >= D) continue;
                    const uint32_t a = (pc0 + (uint32_t)p) % NA;
                    const uint32_t taddr = tmem_base + a * ZR_N + dz * 24 + ((uint32_t)(ew * 32) << 16);
                    uint32_t v0[8], v1[8], v2[8];
                    tmem_ld_32x8(taddr, v0);
                    tmem_ld_32x8(taddr + 8, v1);
                    tmem_ld_32x8(taddr + 16, v2);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        S[0][j] += __uint_as_float(v0[j]);
                        S[1][j] += __uint_as_float(v1[j]);
                        S[2][j] += __uint_as_float(v2[j]);
                    }
                }
            }
            float o[8];
#pragma unroll
            for (int j = 0; j < 8; ++j)
                o[j] = S[0][j] + __shfl_down_sync(0xFFFFFFFFu, S[1][j], 1) + __shfl_down_sync(0xFFFFFFFFu, S[2][j], 2);
            if (valid) {
                uint4 pk;
                __half2 *h2 = reinterpret_cast<__half2 *>(&pk);
#pragma unroll
                for (int j = 0; j < 4; ++j) h2[j] = __floats2half2_rn(o[2 * j], o[2 * j + 1]);
                *reinterpret_cast<uint4 *>(g.out + ((size_t)n * vox_chunk + ((size_t)z * g.H + h) * g.W + w) * 8) = pk;
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) st[lane * 9 + j] = valid ? o[j] : 0.0f;
            __syncwarp();
            {   // lane = (quarter of the rows, channel): 8 rows each, then two shuffle steps
                const int c = lane & 7, r0 = (lane >> 3) * 8;
                float s = 0.0f, q2 = 0.0f;
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    const float x = st[(r0 + r) * 9 + c];
                    s += x;
                    q2 = fmaf(x, x, q2);
                }
                s += __shfl_xor_sync(0xFFFFFFFFu, s, 8);
                q2 += __shfl_xor_sync(0xFFFFFFFFu, q2, 8);
                s += __shfl_xor_sync(0xFFFFFFFFu, s, 16);
                q2 += __shfl_xor_sync(0xFFFFFFFFu, q2, 16);
                stat_guard(q2);
                csum += __float2ll_rn(s * 16777216.0f);            // lanes >= 8 hold copies; only lanes < 5 flush
                csq += __float2ll_rn(q2 * 16777216.0f);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (z >= 1) mbar_arrive(&acc_empty[(pc0 + (uint32_t)z - 1u) % NA]);
                if (z == D - 1) mbar_arrive(&acc_empty[(pc0 + (uint32_t)z) % NA]);
            }
        }
        pc0 += (uint32_t)D;
    }
    flush(cur_n);
}

__global__ void __launch_bounds__(ZR_THREADS, 2)
conv3d_zring_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                    const __grid_constant__ CUtensorMap tmB, const ZringGeom g) {
    using namespace sm100;
    extern __shared__ uint8_t smem_dyn[];
    const uint32_t raw_base = smem_u32(smem_dyn);
    uint8_t *a_smem = smem_dyn + (((raw_base + 1023u) & ~1023u) - raw_base);
    uint8_t *b_smem = a_smem + ZR_SLOTS * ZR_PLANE_BYTES;
    uint64_t *bars = reinterpret_cast<uint64_t *>(b_smem + 6 * ZR_B_STAGE);
    uint64_t *plane_full = bars, *plane_empty = bars + ZR_SLOTS;
    uint64_t *acc_full = bars + 2 * ZR_SLOTS, *acc_empty = acc_full + ZR_NA;
    uint64_t *b_full = acc_empty + ZR_NA;
    uint64_t *sched_full = b_full + 1, *sched_empty = sched_full + ZR_SCHED;
    volatile int *sched_col = reinterpret_cast<volatile int *>(sched_empty + ZR_SCHED);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(const_cast<int *>(sched_col) + ZR_SCHED);
    float *stat_t = reinterpret_cast<float *>(reinterpret_cast<uint8_t *>(bars) + 512);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int D = g.D;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmA0);
        prefetch_tmap(&tmA1);
        prefetch_tmap(&tmB);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < ZR_SLOTS; ++i) {
            mbar_init(&plane_full[i], 1);
            mbar_init(&plane_empty[i], 1);
        }
        for (int i = 0; i < ZR_NA; ++i) {
            mbar_init(&acc_full[i], 1);
            mbar_init(&acc_empty[i], 4);
        }
        mbar_init(b_full, 1);
        for (int i = 0; i < ZR_SCHED; ++i) {
            mbar_init(&sched_full[i], 1);
            mbar_init(&sched_empty[i], 5);                 // MMA warp + 4 epilogue warps
        }
        fence_barrier_init();
    }
    if (warp == 2) {
        tmem_alloc(tmem_slot, ZR_TMEM_COLS);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    auto decode = [&](int col, int &wb, int &hb, int &n) {
        wb = col % g.tiles_w;
        const int t = col / g.tiles_w;
        hb = t % g.tiles_h;
        n = t / g.tiles_h;
    };

    if (warp == 0) {
        // ===================== A producer + scheduler =====================
        if (lane == 0) {
            uint32_t lc = 0, sidx = 0;
            int col = blockIdx.x;
            for (;;) {
                const uint32_t slot = sidx % ZR_SCHED;
                mbar_wait(&sched_empty[slot], ((sidx / ZR_SCHED) & 1u) ^ 1u);
                sched_col[slot] = col;
                mbar_arrive(&sched_full[slot]);
                ++sidx;
                if (col >= g.n_cols) break;
                const int next = (int)gridDim.x + (int)atomicAdd(g.sched, 1u);
                int wb, hb, n;
                decode(col, wb, hb, n);
                for (int p = 0; p < D; ++p)
                    for (int kb = 0; kb < 2; ++kb, ++lc) {
                        const uint32_t s = lc % ZR_SLOTS;
                        mbar_wait(&plane_empty[s], ((lc / ZR_SLOTS) & 1u) ^ 1u);
                        mbar_expect_tx(&plane_full[s], ZR_PLANE_BYTES);
                        tma_load_5d(a_smem + (size_t)s * ZR_PLANE_BYTES, kb == 0 ? &tmA0 : &tmA1, &plane_full[s],
                                    0, wb * ZR_WT - 1, hb * ZR_HT - 1, p, n);
                    }
                col = next;
            }
        }
    } else if (warp == 3) {
        // ===================== B loader (once) =====================
        if (lane == 0) {
            mbar_expect_tx(b_full, 6 * ZR_B_STAGE);
            for (int kb = 0; kb < 2; ++kb)
                for (int dy = 0; dy < 3; ++dy)
                    tma_load_3d(b_smem + (size_t)(kb * 3 + dy) * ZR_B_STAGE, &tmB, b_full, kb * 32, 0, dy);
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        const bool leader = elect_one();
        const uint32_t idesc = make_idesc_f16(128, ZR_N, 0 /* fp16 */);
        const uint64_t dproto = make_kmajor_desc(0, 64, 0);
        const uint32_t d_hi = (uint32_t)(dproto >> 32), d_lo = (uint32_t)dproto;
        const uint32_t a_lo = d_lo | (smem_u32(a_smem) >> 4);
        const uint32_t b_lo = d_lo | (smem_u32(b_smem) >> 4);
        constexpr uint32_t U = 64 >> 4;                        // one row in 16-byte units
        uint32_t lc = 0, pc = 0;
        mbar_wait(b_full, 0u);
        for (uint32_t sidx = 0;; ++sidx) {
            const uint32_t slot = sidx % ZR_SCHED;
            mbar_wait(&sched_full[slot], (sidx / ZR_SCHED) & 1u);
            const int col = sched_col[slot];
            __syncwarp();
            if (leader) mbar_arrive(&sched_empty[slot]);
            if (col >= g.n_cols) break;
            for (int p = 0; p < D; ++p, ++pc) {
                const uint32_t a = pc % ZR_NA;
                mbar_wait(&acc_empty[a], ((pc / ZR_NA) & 1u) ^ 1u);
                const uint32_t tmem_d = tmem_base + a * ZR_N;
#pragma unroll
                for (int kb = 0; kb < 2; ++kb, ++lc) {
                    const uint32_t s = lc % ZR_SLOTS;
                    mbar_wait(&plane_full[s], (lc / ZR_SLOTS) & 1u);
                    tc_fence_after();
                    if (leader) {
                        const uint32_t a_pl = a_lo + s * (ZR_PLANE_BYTES >> 4);
#pragma unroll
                        for (int dy = 0; dy < 3; ++dy) {
                            const uint32_t a_tap = a_pl + dy * ZR_P * U;
                            const uint32_t b_tap = b_lo + (kb * 3 + dy) * (ZR_B_STAGE >> 4);
#pragma unroll
                            for (int k = 0; k < 2; ++k)
                                umma_f16(tmem_d, ((uint64_t)d_hi << 32) | (a_tap + 2 * k),
                                         ((uint64_t)d_hi << 32) | (b_tap + 2 * k), idesc, (kb | dy | k) != 0 ? 1u : 0u);
                        }
                        umma_commit(&plane_empty[s]);
                        if (kb == 1) umma_commit(&acc_full[a]);
                    }
                    __syncwarp();
                }
            }
        }
    } else if (warp >= 4) {
        // ===================== epilogue =====================
        zring_epilogue<ZR_NA, 16, false>(g, tmem_base, acc_full, acc_empty, sched_full, sched_empty, sched_col,
                                  stat_t + (warp - 4) * (32 * 9), warp - 4, lane);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, ZR_TMEM_COLS);
}

// ---------------------------------------------------------------------------------------------
// c8_0.conv1: conv3d 5 -> 5 on relu(bn(raw8)) (unet.py:96 of the last ConvModule) with the same z-ring.
//
// On CUDA cores the layer costs 675 FMAs per voxel and ran at a third of the FFMA rate (1.39 ms
// per frame).  Here the input is K = 16 (5 channels + zeros), so an input plane needs only THREE
// MMAs (one per dy; the nine (dz,dx) taps are the N = 80 columns as above).  The input cannot come
// through TMA because BatchNorm + ReLU of c8_0.conv0's raw output sit in front of it: four producer
// warps load the fp16 [vox][8] halo plane, normalise, convert and write the rows themselves --
// one 16-byte chunk per voxel at its 64-byte-swizzled position (rows are 64 bytes so that the
// layout is the one the other kernels use; only K step 0 is ever read, its second chunk is zeroed
// once) -- then fence.proxy.async and arrive on the plane's mbarrier.  The next plane's loads are
// issued before the current plane is stored.  3 accumulators of 80 columns (256 TMEM columns), so
// two CTAs share an SM and overlap each other's MMA / epilogue phases.
// Warp roles: 0..3 = producers (warp 0 lane 0 also schedules columns), 4..7 = epilogue, 8 = MMA
// issuer + TMEM allocator.
static constexpr int ZO_NA = 3;
static constexpr int ZO_SLOTS = 3;
static constexpr int ZO_THREADS = 288;

struct ZoutArgs {
    ZringGeom g;                          // out = raw9 fp16 [N][vox][8], stats = stats9 [N][5][2]
    const __half *src;                    // raw8 fp16 [N][vox][8]
    const unsigned long long *stats_in;   // [N][16][2]
    const float *gamma_in, *beta_in, *eps_in;
    const float *wgt;                     // [27][5 in][5 out] fp32
};

__host__ __device__ constexpr size_t zout_smem_bytes() {
    return 1024 + (size_t)ZO_SLOTS * ZR_PLANE_BYTES + 3 * ZR_B_STAGE + 512 /* barriers */ + 4 * 32 * 9 * sizeof(float);
}

__global__ void __launch_bounds__(ZO_THREADS, 2)
conv_out_zring_kernel(const ZoutArgs a) {
    using namespace sm100;
    const ZringGeom &g = a.g;
    extern __shared__ uint8_t smem_dyn[];
    const uint32_t raw_base = smem_u32(smem_dyn);
    uint8_t *a_smem = smem_dyn + (((raw_base + 1023u) & ~1023u) - raw_base);
    uint8_t *b_smem = a_smem + ZO_SLOTS * ZR_PLANE_BYTES;
    uint64_t *bars = reinterpret_cast<uint64_t *>(b_smem + 3 * ZR_B_STAGE);
    uint64_t *plane_full = bars, *plane_empty = bars + ZO_SLOTS;
    uint64_t *acc_full = bars + 2 * ZO_SLOTS, *acc_empty = acc_full + ZO_NA;
    uint64_t *sched_full = acc_empty + ZO_NA, *sched_empty = sched_full + ZR_SCHED;
    volatile int *sched_col = reinterpret_cast<volatile int *>(sched_empty + ZR_SCHED);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(const_cast<int *>(sched_col) + ZR_SCHED);
    float *stat_t = reinterpret_cast<float *>(reinterpret_cast<uint8_t *>(bars) + 512);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int D = g.D, H = g.H, W = g.W;

    if (tid == 32) {
        for (int i = 0; i < ZO_SLOTS; ++i) {
            mbar_init(&plane_full[i], 128);                // every producer thread arrives
            mbar_init(&plane_empty[i], 1);
        }
        for (int i = 0; i < ZO_NA; ++i) {
            mbar_init(&acc_full[i], 1);
            mbar_init(&acc_empty[i], 4);
        }
        for (int i = 0; i < ZR_SCHED; ++i) {
            mbar_init(&sched_full[i], 1);
            mbar_init(&sched_empty[i], 9);                 // 4 producer warps + MMA warp + 4 epilogue warps
        }
        fence_barrier_init();
    }
    if (warp == 8) {
        tmem_alloc(tmem_slot, 256);
        tmem_relinquish();
    }
    // weights -> B[dy][(dz*3 + dx) * 8 + co][k = ci] fp16, 64-byte swizzled rows (chunks 0 and 1 = K step 0;
    // chunk 1 and the rows >= 72 are zero); zero chunk 1 of every A row of every slot once
    for (int i = tid; i < 3 * ZR_N * 2; i += ZO_THREADS) {
        const int c = i & 1, row = (i >> 1) % ZR_N, dy = i / (2 * ZR_N);
        const int co = row & 7, t9 = row >> 3;
        uint4 pk = make_uint4(0u, 0u, 0u, 0u);
        if (c == 0 && t9 < 9 && co < 5) {
            const int tap = (t9 / 3) * 9 + dy * 3 + t9 % 3;
            __half *h = reinterpret_cast<__half *>(&pk);
#pragma unroll
            for (int ci = 0; ci < 5; ++ci) h[ci] = __float2half_rn(a.wgt[(tap * 5 + ci) * 5 + co]);
        }
        *reinterpret_cast<uint4 *>(b_smem + dy * ZR_B_STAGE + row * 64 + ((c ^ ((row >> 1) & 3)) << 4)) = pk;
    }
    for (int i = tid; i < ZO_SLOTS * ZR_PLANE_ROWS; i += ZO_THREADS) {
        const int r = i % ZR_PLANE_ROWS, s = i / ZR_PLANE_ROWS;
        *reinterpret_cast<uint4 *>(a_smem + s * ZR_PLANE_BYTES + r * 64 + ((1 ^ ((r >> 1) & 3)) << 4)) =
            make_uint4(0u, 0u, 0u, 0u);
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < 4) {
        // ===================== producers (+ scheduler) =====================
        const size_t vox_chunk = (size_t)D * H * W;
        // this thread's rows of a plane: r0 = tid, r1 = tid + 128 (< 192 for tid < 64)
        const int r0 = tid, r1 = tid + 128;
        const bool has1 = r1 < ZR_PLANE_ROWS;
        const uint32_t off0 = (uint32_t)r0 * 64 + ((uint32_t)((r0 >> 1) & 3) << 4);       // chunk 0 ^ swizzle
        const uint32_t off1 = (uint32_t)r1 * 64 + ((uint32_t)((r1 >> 1) & 3) << 4);
        float sc[5], sh[5];
        int cur_n = -1;
        uint32_t lc = 0, sidx = 0;
        int next_col = blockIdx.x;
        for (;; ++sidx) {
            const uint32_t slot = sidx % ZR_SCHED;
            if (tid == 0) {
                mbar_wait(&sched_empty[slot], ((sidx / ZR_SCHED) & 1u) ^ 1u);
                sched_col[slot] = next_col;
                mbar_arrive(&sched_full[slot]);
                if (next_col < g.n_cols) next_col = (int)gridDim.x + (int)atomicAdd(g.sched, 1u);
            }
            mbar_wait(&sched_full[slot], (sidx / ZR_SCHED) & 1u);
            const int col = sched_col[slot];
            __syncwarp();
            if (lane == 0) mbar_arrive(&sched_empty[slot]);
            if (col >= g.n_cols) break;
            const int wb = col % g.tiles_w;
            const int hb = (col / g.tiles_w) % g.tiles_h;
            const int n = col / (g.tiles_w * g.tiles_h);
            if (n != cur_n) {
                cur_n = n;
#pragma unroll
                for (int c = 0; c < 5; ++c)
                    bn_coeffs(a.stats_in + ((size_t)n * 16 + c) * 2, a.gamma_in[c], a.beta_in[c], a.eps_in[c],
                              1.0f / (float)vox_chunk, sc[c], sh[c]);
            }
            // in-plane positions of the two rows (the same for every plane of the column)
            const int h0 = hb * ZR_HT - 1 + (r0 >> 5), w0 = wb * ZR_WT - 1 + (r0 & 31);
            const int h1 = hb * ZR_HT - 1 + (r1 >> 5), w1 = wb * ZR_WT - 1 + (r1 & 31);
            const bool in0 = h0 >= 0 && h0 < H && w0 >= 0 && w0 < W;
            const bool in1 = has1 && h1 >= 0 && h1 < H && w1 >= 0 && w1 < W;
            const __half *p0 = a.src + ((size_t)n * vox_chunk + (size_t)(in0 ? h0 : 0) * W + (in0 ? w0 : 0)) * 8;
            const __half *p1 = a.src + ((size_t)n * vox_chunk + (size_t)(in1 ? h1 : 0) * W + (in1 ? w1 : 0)) * 8;
            const size_t zstride = (size_t)H * W * 8;
            float4 c0lo, c1lo;
            float c0hi, c1hi;
            auto load = [&](int p, float4 &lo0, float &hi0, float4 &lo1, float &hi1) {
                lo0 = lo1 = make_float4(0.f, 0.f, 0.f, 0.f);
                hi0 = hi1 = 0.f;
                if (in0) {                                        // one 16-byte load per voxel: 8 halves, 5 used
                    const uint4 u = __ldg(reinterpret_cast<const uint4 *>(p0 + p * zstride));
                    const __half2 *h = reinterpret_cast<const __half2 *>(&u);
                    const float2 a01 = __half22float2(h[0]), a23 = __half22float2(h[1]), a45 = __half22float2(h[2]);
                    lo0 = make_float4(a01.x, a01.y, a23.x, a23.y);
                    hi0 = a45.x;
                }
                if (in1) {
                    const uint4 u = __ldg(reinterpret_cast<const uint4 *>(p1 + p * zstride));
                    const __half2 *h = reinterpret_cast<const __half2 *>(&u);
                    const float2 a01 = __half22float2(h[0]), a23 = __half22float2(h[1]), a45 = __half22float2(h[2]);
                    lo1 = make_float4(a01.x, a01.y, a23.x, a23.y);
                    hi1 = a45.x;
                }
            };
            auto pack = [&](const float4 &lo, float hi, bool in) {
                uint4 pk = make_uint4(0u, 0u, 0u, 0u);
                if (in) {
                    __half2 *h2 = reinterpret_cast<__half2 *>(&pk);
                    h2[0] = __floats2half2_rn(fmaxf(fmaf(lo.x, sc[0], sh[0]), 0.f), fmaxf(fmaf(lo.y, sc[1], sh[1]), 0.f));
                    h2[1] = __floats2half2_rn(fmaxf(fmaf(lo.z, sc[2], sh[2]), 0.f), fmaxf(fmaf(lo.w, sc[3], sh[3]), 0.f));
                    h2[2] = __floats2half2_rn(fmaxf(fmaf(hi, sc[4], sh[4]), 0.f), 0.f);
                }
                return pk;
            };
            load(0, c0lo, c0hi, c1lo, c1hi);
            for (int p = 0; p < D; ++p, ++lc) {
                float4 n0lo, n1lo;
                float n0hi, n1hi;
                if (p + 1 < D) load(p + 1, n0lo, n0hi, n1lo, n1hi);      // in flight while this plane is stored
                const uint32_t s = lc % ZO_SLOTS;
                mbar_wait(&plane_empty[s], ((lc / ZO_SLOTS) & 1u) ^ 1u);
                uint8_t *pl = a_smem + (size_t)s * ZR_PLANE_BYTES;
                *reinterpret_cast<uint4 *>(pl + off0) = pack(c0lo, c0hi, in0);
                if (has1) *reinterpret_cast<uint4 *>(pl + off1) = pack(c1lo, c1hi, in1);
                fence_proxy_async();
                mbar_arrive(&plane_full[s]);
                c0lo = n0lo; c0hi = n0hi; c1lo = n1lo; c1hi = n1hi;
            }
        }
    } else if (warp == 8) {
        // ===================== MMA issuer =====================
        const bool leader = elect_one();
        const uint32_t idesc = make_idesc_f16(128, ZR_N, 0 /* fp16 */);
        const uint64_t dproto = make_kmajor_desc(0, 64, 0);
        const uint32_t d_hi = (uint32_t)(dproto >> 32), d_lo = (uint32_t)dproto;
        const uint32_t a_lo = d_lo | (smem_u32(a_smem) >> 4);
        const uint32_t b_lo = d_lo | (smem_u32(b_smem) >> 4);
        constexpr uint32_t U = 64 >> 4;
        uint32_t lc = 0, pc = 0;
        for (uint32_t sidx = 0;; ++sidx) {
            const uint32_t slot = sidx % ZR_SCHED;
            mbar_wait(&sched_full[slot], (sidx / ZR_SCHED) & 1u);
            const int col = sched_col[slot];
            __syncwarp();
            if (leader) mbar_arrive(&sched_empty[slot]);
            if (col >= g.n_cols) break;
            for (int p = 0; p < D; ++p, ++pc, ++lc) {
                const uint32_t acc = pc % ZO_NA, s = lc % ZO_SLOTS;
                mbar_wait(&acc_empty[acc], ((pc / ZO_NA) & 1u) ^ 1u);
                mbar_wait(&plane_full[s], (lc / ZO_SLOTS) & 1u);
                tc_fence_after();
                if (leader) {
                    const uint32_t a_pl = a_lo + s * (ZR_PLANE_BYTES >> 4);
#pragma unroll
                    for (int dy = 0; dy < 3; ++dy)
                        umma_f16(tmem_base + acc * ZR_N, ((uint64_t)d_hi << 32) | (a_pl + dy * ZR_P * U),
                                 ((uint64_t)d_hi << 32) | (b_lo + dy * (ZR_B_STAGE >> 4)), idesc, dy != 0 ? 1u : 0u);
                    umma_commit(&plane_empty[s]);
                    umma_commit(&acc_full[acc]);
                }
                __syncwarp();
            }
        }
    } else {
        // ===================== epilogue (warps 4..7) =====================
        zring_epilogue<ZO_NA, 5, false>(g, tmem_base, acc_full, acc_empty, sched_full, sched_empty, sched_col,
                                 stat_t + (warp - 4) * (32 * 9), warp - 4, lane);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 8) tmem_dealloc(tmem_base, 256);
}

// nn.Conv3d weight (5, 64, 3,3,3) fp32 -> [dy][(dz*3 + dx) * 8 + co][cin] fp16, 80 rows per dy
// (co >= 5 and rows >= 72 zero)
__global__ void pack_conv_w_zring_kernel(const float *__restrict__ src, const float *__restrict__ inv_s,
                                         __half *__restrict__ dst, int cout, int cin) {
    const int total = 3 * ZR_N * cin;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int ci = i % cin;
        const int row = (i / cin) % ZR_N;
        const int dy = i / (cin * ZR_N);
        const int co = row & 7, t9 = row >> 3;              // t9 = dz * 3 + dx
        float v = 0.0f;
        if (t9 < 9 && co < cout) {
            const int dz = t9 / 3, dx = t9 % 3;
            v = src[((size_t)co * cin + ci) * 27 + dz * 9 + dy * 3 + dx] * inv_s[co];
        }
        dst[i] = __float2half_rn(v);
    }
}

}  // namespace isg
