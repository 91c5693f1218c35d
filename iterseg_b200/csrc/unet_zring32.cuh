// 32 -> 32 convolutions at the two finest levels (c0.conv1 and c7_0.conv1; unet.py:96) on a TMEM
// ring along z with the three dz taps folded into N.
//
// In the generic kernel (unet_conv.cuh) these layers use the dx-fold, whose epilogue pays three TMEM
// reads + 64 warp shuffles per tile (25 % of the kernel, profiles/r01_notes.md) and whose z-groups
// re-load (T+2)/T halo planes.  Here one accumulator belongs to an INPUT plane p:
//     D_p[r, (dz,co)] = sum_{dy,dx,ci} A_p[r + dy*P + dx, ci] * W[dz,dy,dx][ci,co]      (N = 96, 18 MMAs)
// kept in a ring of 5 accumulators while the CTA walks a z-column of the chunk; output plane
// z = D_{z-1}[., dz=0] + D_z[., dz=1] + D_{z+1}[., dz=2]: three TMEM reads and plain adds, no row
// shifts.  Every input plane is loaded once per column (TMA halo box {32 ch, P = 32, 6 rows}, 64-byte
// swizzled rows, out-of-bounds = 0 = the conv padding).  Weights resident in shared memory, packed
// [dy*3+dx][dz*32 + co][cin] (pack_conv_w_zring32_kernel).
// Warp roles: 0 = A producer + column scheduler, 1 = MMA issuer, 2 = TMEM allocator, 3 = B loader,
//             4.. = epilogue (EPI = 4 warps: one per TMEM lane quarter = patch row, 32 channels each; EPI = 8:
//             two per quarter, 16 channels each -- the epilogue of a plane is a latency chain of ~300
//             instructions in ONE warp per scheduler, and a plane's 18 MMAs take only ~1 000 clocks).
//             Columns are handed out dynamically (see unet_conv.cuh).
#pragma once
#include <cuda.h>

#include "common.cuh"
#include "sm100.cuh"
#include "unet_zring.cuh"

namespace isg {

static constexpr int Z32_NA = 5;
static constexpr int Z32_SLOTS = 6;
static constexpr int Z32_N = 96;
static constexpr int Z32_B_STAGE = Z32_N * 64;                       // one (dy,dx) tap: 6144 B
static constexpr int Z32_SLACK = 1024;                               // rows the last taps of halo rows touch

struct Z32Args {
    int N, D, H, W;
    int tiles_w, tiles_h, n_cols;
    __half *out;                          // raw fp16 [N][vox][32]
    unsigned long long *stats;            // [N][32][2]
    unsigned int *sched;
};

__host__ __device__ constexpr size_t z32_smem_bytes() {
    return 1024 + (size_t)Z32_SLOTS * ZR_PLANE_BYTES + Z32_SLACK + 9 * Z32_B_STAGE + 512 /* barriers */ +
           8 * 32 * 17 * sizeof(float);                 // >= 4 * 32 * 33: the statistics transposes of either variant
}

template <int CH>
__device__ __forceinline__ void z32_tmem_ld(uint32_t taddr, uint32_t (&v)[CH]) {
    if constexpr (CH == 32) sm100::tmem_ld_32x32(taddr, v);
    else sm100::tmem_ld_32x16(taddr, v);
}

template <int EPI>
__global__ void __launch_bounds__(128 + 32 * EPI, 1)
conv3d_zring32_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                      const Z32Args g) {
    using namespace sm100;
    extern __shared__ uint8_t smem_dyn[];
    const uint32_t raw_base = smem_u32(smem_dyn);
    uint8_t *a_smem = smem_dyn + (((raw_base + 1023u) & ~1023u) - raw_base);
    uint8_t *b_smem = a_smem + Z32_SLOTS * ZR_PLANE_BYTES + Z32_SLACK;
    uint64_t *bars = reinterpret_cast<uint64_t *>(b_smem + 9 * Z32_B_STAGE);
    uint64_t *plane_full = bars, *plane_empty = bars + Z32_SLOTS;
    uint64_t *acc_full = bars + 2 * Z32_SLOTS, *acc_empty = acc_full + Z32_NA;
    uint64_t *b_full = acc_empty + Z32_NA;
    uint64_t *sched_full = b_full + 1, *sched_empty = sched_full + ZR_SCHED;
    volatile int *sched_col = reinterpret_cast<volatile int *>(sched_empty + ZR_SCHED);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(const_cast<int *>(sched_col) + ZR_SCHED);
    float *stat_t = reinterpret_cast<float *>(reinterpret_cast<uint8_t *>(bars) + 512);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int D = g.D, H = g.H, W = g.W;
    const size_t vox_chunk = (size_t)D * H * W;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmA);
        prefetch_tmap(&tmB);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < Z32_SLOTS; ++i) {
            mbar_init(&plane_full[i], 1);
            mbar_init(&plane_empty[i], 1);
        }
        for (int i = 0; i < Z32_NA; ++i) {
            mbar_init(&acc_full[i], 1);
            mbar_init(&acc_empty[i], EPI);
        }
        mbar_init(b_full, 1);
        for (int i = 0; i < ZR_SCHED; ++i) {
            mbar_init(&sched_full[i], 1);
            mbar_init(&sched_empty[i], 1 + EPI);           // MMA warp + the epilogue warps
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

    auto decode = [&](int col, int &wb, int &hb, int &n) {
        wb = col % g.tiles_w;
        hb = (col / g.tiles_w) % g.tiles_h;
        n = col / (g.tiles_w * g.tiles_h);
    };

    if (warp == 0) {
        // ===================== A producer + scheduler =====================
        if (lane == 0) {
            uint32_t lc = 0, sidx = 0;
            int col = blockIdx.x;
#ifdef ISG_Z32_PROF
            long long prof_w = 0;
            const long long prof_t0 = clock64();
#endif
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
                for (int p = 0; p < D; ++p, ++lc) {
                    const uint32_t s = lc % Z32_SLOTS;
#ifdef ISG_Z32_PROF
                    const long long pt0 = clock64();
#endif
                    mbar_wait(&plane_empty[s], ((lc / Z32_SLOTS) & 1u) ^ 1u);
#ifdef ISG_Z32_PROF
                    prof_w += clock64() - pt0;
#endif
                    mbar_expect_tx(&plane_full[s], ZR_PLANE_BYTES);
                    tma_load_5d(a_smem + (size_t)s * ZR_PLANE_BYTES, &tmA, &plane_full[s], 0, wb * ZR_WT - 1,
                                hb * ZR_HT - 1, p, n);
                }
                col = next;
            }
#ifdef ISG_Z32_PROF
            if (blockIdx.x == 0 || blockIdx.x == 77)
                printf("z32 blk %d producer: total %lld, wait plane_empty %lld, planes %u\n", blockIdx.x,
                       clock64() - prof_t0, prof_w, lc);
#endif
        }
    } else if (warp == 3) {
        // ===================== B loader (once) =====================
        if (lane == 0) {
            mbar_expect_tx(b_full, 9 * Z32_B_STAGE);
            for (int t = 0; t < 9; ++t) tma_load_3d(b_smem + (size_t)t * Z32_B_STAGE, &tmB, b_full, 0, 0, t);
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        const bool leader = elect_one();
        const uint32_t idesc = make_idesc_f16(128, Z32_N, 0 /* fp16 */);
        const uint64_t dproto = make_kmajor_desc(0, 64, 0);
        const uint32_t d_hi = (uint32_t)(dproto >> 32), d_lo = (uint32_t)dproto;
        const uint32_t a_lo = d_lo | (smem_u32(a_smem) >> 4);
        const uint32_t b_lo = d_lo | (smem_u32(b_smem) >> 4);
        constexpr uint32_t U = 64 >> 4;
        uint32_t lc = 0, pc = 0;
        mbar_wait(b_full, 0u);
#ifdef ISG_Z32_PROF
        long long prof_acc = 0, prof_pl = 0, prof_sched = 0;
        const long long prof_t0 = clock64();
#endif
        for (uint32_t sidx = 0;; ++sidx) {
            const uint32_t slot = sidx % ZR_SCHED;
#ifdef ISG_Z32_PROF
            const long long st0 = clock64();
#endif
            mbar_wait(&sched_full[slot], (sidx / ZR_SCHED) & 1u);
#ifdef ISG_Z32_PROF
            prof_sched += clock64() - st0;
#endif
            const int col = sched_col[slot];
            __syncwarp();
            if (leader) mbar_arrive(&sched_empty[slot]);
            if (col >= g.n_cols) break;
            for (int p = 0; p < D; ++p, ++pc, ++lc) {
                const uint32_t acc = pc % Z32_NA, s = lc % Z32_SLOTS;
#ifdef ISG_Z32_PROF
                const long long at0 = clock64();
#endif
                mbar_wait(&acc_empty[acc], ((pc / Z32_NA) & 1u) ^ 1u);
#ifdef ISG_Z32_PROF
                const long long at1 = clock64();
                prof_acc += at1 - at0;
#endif
                mbar_wait(&plane_full[s], (lc / Z32_SLOTS) & 1u);
#ifdef ISG_Z32_PROF
                prof_pl += clock64() - at1;
#endif
                tc_fence_after();
                if (leader) {
                    const uint32_t a_pl = a_lo + s * (ZR_PLANE_BYTES >> 4);
                    const uint32_t tmem_d = tmem_base + acc * Z32_N;
#pragma unroll
                    for (int t = 0; t < 9; ++t) {
                        const uint32_t a_tap = a_pl + ((t / 3) * ZR_P + t % 3) * U;
                        const uint32_t b_tap = b_lo + t * (Z32_B_STAGE >> 4);
#pragma unroll
                        for (int k = 0; k < 2; ++k)
                            umma_f16(tmem_d, ((uint64_t)d_hi << 32) | (a_tap + 2 * k),
                                     ((uint64_t)d_hi << 32) | (b_tap + 2 * k), idesc, (t | k) != 0 ? 1u : 0u);
                    }
                    umma_commit(&plane_empty[s]);
                    umma_commit(&acc_full[acc]);
                }
                __syncwarp();
            }
        }
#ifdef ISG_Z32_PROF
        if (leader && (blockIdx.x == 0 || blockIdx.x == 77))
            printf("z32 blk %d mma: total %lld, wait acc_empty %lld, wait plane_full %lld, wait sched %lld, planes %u\n",
                   blockIdx.x, clock64() - prof_t0, prof_acc, prof_pl, prof_sched, pc);
#endif
    } else if (warp >= 4) {
        // ===================== epilogue =====================
        constexpr int CH = 128 / EPI;                    // channels per warp: 32 or 16
        const int ew = (warp - 4) & 3;                   // TMEM lanes 32*ew.. = patch row ew (= warp % 4)
        const int c_base = ((warp - 4) >> 2) * CH;
        float *st = stat_t + (warp - 4) * (32 * (CH + 1));
        long long csum = 0, csq = 0;                     // channel c_base + lane % CH of the current chunk
        int cur_n = -1;
        auto flush = [&](int n) {
            if (n >= 0 && lane < CH) {
                atomicAdd(g.stats + ((size_t)n * 32 + c_base + lane) * 2 + 0, (unsigned long long)csum);
                atomicAdd(g.stats + ((size_t)n * 32 + c_base + lane) * 2 + 1, (unsigned long long)csq);
            }
            csum = csq = 0;
        };
        const uint32_t lane_base = ((uint32_t)(ew * 32) << 16) + (uint32_t)c_base;
        uint32_t pc0 = 0;
#ifdef ISG_Z32_PROF
        long long prof_full = 0;
        const long long prof_t0 = clock64();
#endif
        for (uint32_t sidx = 0;; ++sidx) {
            const uint32_t slot = sidx % ZR_SCHED;
            mbar_wait(&sched_full[slot], (sidx / ZR_SCHED) & 1u);
            const int col = sched_col[slot];
            __syncwarp();
            if (lane == 0) mbar_arrive(&sched_empty[slot]);
            if (col >= g.n_cols) break;
            int wb, hb, n;
            decode(col, wb, hb, n);
            if (n != cur_n) {
                flush(cur_n);
                cur_n = n;
            }
            const int h = hb * ZR_HT + ew, w = wb * ZR_WT + lane;
            const bool valid = lane < ZR_WT && h < H && w < W;
            int waited = 0;
            for (int z = 0; z < D; ++z) {
                const int need = z + 1 < D ? z + 1 : D - 1;
#ifdef ISG_Z32_PROF
                const long long ft0 = clock64();
#endif
                while (waited <= need) {
                    const uint32_t qq = pc0 + (uint32_t)waited;
                    mbar_wait(&acc_full[qq % Z32_NA], (qq / Z32_NA) & 1u);
                    ++waited;
                }
#ifdef ISG_Z32_PROF
                prof_full += clock64() - ft0;
#endif
                tc_fence_after();
                uint32_t v[CH], v0[CH], v2[CH];
                z32_tmem_ld<CH>(tmem_base + ((pc0 + (uint32_t)z) % Z32_NA) * Z32_N + 32 + lane_base, v);
                if (z >= 1) z32_tmem_ld<CH>(tmem_base + ((pc0 + (uint32_t)z - 1u) % Z32_NA) * Z32_N + lane_base, v0);
                if (z + 1 < D) z32_tmem_ld<CH>(tmem_base + ((pc0 + (uint32_t)z + 1u) % Z32_NA) * Z32_N + 64 + lane_base, v2);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < CH; ++j) {
                    float x = __uint_as_float(v[j]);
                    if (z >= 1) x += __uint_as_float(v0[j]);
                    if (z + 1 < D) x += __uint_as_float(v2[j]);
                    v[j] = __float_as_uint(x);
                }
                if (valid) {
                    __half *o = g.out + ((size_t)n * vox_chunk + ((size_t)z * H + h) * W + w) * 32 + c_base;
#pragma unroll
                    for (int qd = 0; qd < CH / 8; ++qd) {
                        uint4 pk;
                        uint32_t *pw = reinterpret_cast<uint32_t *>(&pk);
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            __half2 h2 = __floats2half2_rn(__uint_as_float(v[qd * 8 + e * 2]),
                                                           __uint_as_float(v[qd * 8 + e * 2 + 1]));
                            pw[e] = *reinterpret_cast<uint32_t *>(&h2);
                        }
                        reinterpret_cast<uint4 *>(o)[qd] = pk;
                    }
                }
#pragma unroll
                for (int j = 0; j < CH; ++j) st[lane * (CH + 1) + j] = valid ? __uint_as_float(v[j]) : 0.0f;
                __syncwarp();
                float s = 0.0f, q2 = 0.0f;
                if constexpr (CH == 32) {
#pragma unroll
                    for (int r = 0; r < 32; ++r) {
                        const float x = st[r * 33 + lane];
                        s += x;
                        q2 = fmaf(x, x, q2);
                    }
                } else {
                    // lane = (half of the rows, channel): 16 rows each in a fixed order, then one exchange
                    const int c = lane & 15, r0 = (lane >> 4) * 16;
#pragma unroll
                    for (int r = 0; r < 16; ++r) {
                        const float x = st[(r0 + r) * 17 + c];
                        s += x;
                        q2 = fmaf(x, x, q2);
                    }
                    s += __shfl_xor_sync(0xFFFFFFFFu, s, 16);
                    q2 += __shfl_xor_sync(0xFFFFFFFFu, q2, 16);
                }
                __syncwarp();
                stat_guard(q2);
                csum += __float2ll_rn(s * 16777216.0f);  // CH = 16: lanes >= 16 hold copies, lanes < 16 flush
                csq += __float2ll_rn(q2 * 16777216.0f);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    if (z >= 1) mbar_arrive(&acc_empty[(pc0 + (uint32_t)z - 1u) % Z32_NA]);
                    if (z == D - 1) mbar_arrive(&acc_empty[(pc0 + (uint32_t)z) % Z32_NA]);
                }
            }
            pc0 += (uint32_t)D;
        }
        flush(cur_n);
#ifdef ISG_Z32_PROF
        if (lane == 0 && warp == 4 && (blockIdx.x == 0 || blockIdx.x == 77))
            printf("z32 blk %d epilogue: total %lld, wait acc_full %lld, planes %u\n", blockIdx.x, clock64() - prof_t0,
                   prof_full, pc0);
#endif
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, 512);
}

// ---------------------------------------------------------------------------------------------
// The same layers with the sum over dz taken INSIDE TMEM.  The kernel above reads three accumulators
// per output plane, and tcgen05.ld moves only 64 B per clock per SM: 48 KB = 768 clocks per plane beside
// ~1 150 clocks of MMAs (measured with clock64 counters per role: the issuing thread spends 1 220 clocks
// per plane inside its 18 MMAs and waits 250 for an accumulator; the epilogue warps are busy 1 300, with 4
// warps or with 8).  Here an accumulator belongs to an OUTPUT plane: output plane z of the current column
// sits in TMEM columns 32 * (15 - z) (D <= 16), so that the planes z+1, z, z-1 an input plane z
// contributes to (dz = 0, 1, 2) are 96 CONSECUTIVE columns in the order of the packed weights -- one
// N = 96 MMA per (tap, K step) adds all three (N = 64 at the two ends of the column), the epilogue reads 32
// columns per plane (256 clocks) and adds nothing.  Every MMA accumulates: a block is zeroed by the
// epilogue warp that has just read it (tcgen05.st, 256 B per clock), and once at kernel start.  [A first
// version that cleared fresh blocks with a non-accumulating first MMA (split by freshness) and walked a
// ring of 16 blocks (MMAs split where the ring wrapped) was correct but 45 % SLOWER than the kernel above:
// changing the D address / shape between consecutive MMAs is expensive, so a plane's MMAs must all be alike.]
static constexpr int ZS_NB = 16;

template <int EPI>
__global__ void __launch_bounds__(128 + 32 * EPI, 1)
conv3d_zslide32_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                       const Z32Args g) {
    using namespace sm100;
    extern __shared__ uint8_t smem_dyn[];
    const uint32_t raw_base = smem_u32(smem_dyn);
    uint8_t *a_smem = smem_dyn + (((raw_base + 1023u) & ~1023u) - raw_base);
    uint8_t *b_smem = a_smem + Z32_SLOTS * ZR_PLANE_BYTES + Z32_SLACK;
    uint64_t *bars = reinterpret_cast<uint64_t *>(b_smem + 9 * Z32_B_STAGE);
    uint64_t *plane_full = bars, *plane_empty = bars + Z32_SLOTS;
    uint64_t *acc_full = bars + 2 * Z32_SLOTS, *acc_empty = acc_full + ZS_NB;
    uint64_t *b_full = acc_empty + ZS_NB;
    uint64_t *sched_full = b_full + 1, *sched_empty = sched_full + ZR_SCHED;
    volatile int *sched_col = reinterpret_cast<volatile int *>(sched_empty + ZR_SCHED);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(const_cast<int *>(sched_col) + ZR_SCHED);
    static_assert((2 * Z32_SLOTS + 2 * ZS_NB + 1 + 2 * ZR_SCHED) * 8 + ZR_SCHED * 4 + 4 <= 512, "barrier block");
    float *stat_t = reinterpret_cast<float *>(reinterpret_cast<uint8_t *>(bars) + 512);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int D = g.D, H = g.H, W = g.W;
    const size_t vox_chunk = (size_t)D * H * W;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmA);
        prefetch_tmap(&tmB);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < Z32_SLOTS; ++i) {
            mbar_init(&plane_full[i], 1);
            mbar_init(&plane_empty[i], 1);
        }
        for (int i = 0; i < ZS_NB; ++i) {
            mbar_init(&acc_full[i], 1);
            mbar_init(&acc_empty[i], EPI);
        }
        mbar_init(b_full, 1);
        for (int i = 0; i < ZR_SCHED; ++i) {
            mbar_init(&sched_full[i], 1);
            mbar_init(&sched_empty[i], 1 + EPI);           // MMA warp + the epilogue warps
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

    auto decode = [&](int col, int &wb, int &hb, int &n) {
        wb = col % g.tiles_w;
        hb = (col / g.tiles_w) % g.tiles_h;
        n = col / (g.tiles_w * g.tiles_h);
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
                for (int p = 0; p < D; ++p, ++lc) {
                    const uint32_t s = lc % Z32_SLOTS;
                    mbar_wait(&plane_empty[s], ((lc / Z32_SLOTS) & 1u) ^ 1u);
                    mbar_expect_tx(&plane_full[s], ZR_PLANE_BYTES);
                    tma_load_5d(a_smem + (size_t)s * ZR_PLANE_BYTES, &tmA, &plane_full[s], 0, wb * ZR_WT - 1,
                                hb * ZR_HT - 1, p, n);
                }
                col = next;
            }
        }
    } else if (warp == 3) {
        // ===================== B loader (once) =====================
        if (lane == 0) {
            mbar_expect_tx(b_full, 9 * Z32_B_STAGE);
            for (int t = 0; t < 9; ++t) tma_load_3d(b_smem + (size_t)t * Z32_B_STAGE, &tmB, b_full, 0, 0, t);
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        const bool leader = elect_one();
        const uint32_t idesc0 = make_idesc_f16(128, 0, 0 /* fp16 */);      // + (N >> 3) << 17 per MMA: N = 32, 64 or 96
        const uint64_t dproto = make_kmajor_desc(0, 64, 0);
        const uint32_t d_hi = (uint32_t)(dproto >> 32), d_lo = (uint32_t)dproto;
        const uint32_t a_lo = d_lo | (smem_u32(a_smem) >> 4);
        const uint32_t b_lo = d_lo | (smem_u32(b_smem) >> 4);
        constexpr uint32_t U = 64 >> 4;
        constexpr uint32_t BROW32 = (32 * 64) >> 4;          // 32 weight rows (one dz block) in 16-byte units
        uint32_t lc = 0;                                     // planes loaded
        mbar_wait(b_full, 0u);
        for (uint32_t sidx = 0;; ++sidx) {                   // sidx = columns done: the phase of the block barriers
            const uint32_t slot = sidx % ZR_SCHED;
            mbar_wait(&sched_full[slot], (sidx / ZR_SCHED) & 1u);
            const int col = sched_col[slot];
            __syncwarp();
            if (leader) mbar_arrive(&sched_empty[slot]);
            if (col >= g.n_cols) break;
            for (int p = 0; p < D; ++p, ++lc) {
                const uint32_t s = lc % Z32_SLOTS;
                // blocks this plane is the first to touch: output p + 1 (dz = 0) and, for p = 0, output 0 (dz = 1);
                // phase 0 of acc_empty = the zeroing at kernel start, phase c = read + zeroed in column c - 1
                if (p + 1 < D) mbar_wait(&acc_empty[p + 1], sidx & 1u);
                if (p == 0) mbar_wait(&acc_empty[0], sidx & 1u);
                mbar_wait(&plane_full[s], (lc / Z32_SLOTS) & 1u);
                tc_fence_after();
                if (leader) {
                    // dz blocks e_lo..e_hi -> outputs p + 1 - e: columns 32 * (15 - (p + 1 - e_lo)) onwards
                    const int e_lo = p + 1 < D ? 0 : 1, e_hi = p >= 1 ? 2 : 1;
                    const uint32_t tmem_d = tmem_base + 32u * (uint32_t)(15 - (p + 1 - e_lo));
                    const uint32_t idesc = idesc0 + ((4u * (uint32_t)(e_hi - e_lo + 1)) << 17);
                    const uint32_t b_off = (uint32_t)e_lo * BROW32;
                    const uint32_t a_pl = a_lo + s * (ZR_PLANE_BYTES >> 4);
#pragma unroll
                    for (int t = 0; t < 9; ++t) {
                        const uint32_t a_tap = a_pl + ((t / 3) * ZR_P + t % 3) * U;
                        const uint32_t b_tap = b_lo + t * (Z32_B_STAGE >> 4) + b_off;
#pragma unroll
                        for (int k = 0; k < 2; ++k)
                            umma_f16(tmem_d, ((uint64_t)d_hi << 32) | (a_tap + 2 * k),
                                     ((uint64_t)d_hi << 32) | (b_tap + 2 * k), idesc, 1u);
                    }
                    umma_commit(&plane_empty[s]);
                    if (p >= 1) umma_commit(&acc_full[p - 1]);          // output p - 1 is complete
                    if (p == D - 1) umma_commit(&acc_full[p]);          // and so is the last one
                }
                __syncwarp();
            }
        }
    } else if (warp >= 4) {
        // ===================== epilogue =====================
        constexpr int CH = 128 / EPI;                    // channels per warp: 32 or 16
        const int ew = (warp - 4) & 3;                   // TMEM lanes 32*ew.. = patch row ew (= warp % 4)
        const int c_base = ((warp - 4) >> 2) * CH;
        float *st = stat_t + (warp - 4) * (32 * (CH + 1));
        long long csum = 0, csq = 0;                     // channel c_base + lane % CH of the current chunk
        int cur_n = -1;
        auto flush = [&](int n) {
            if (n >= 0 && lane < CH) {
                atomicAdd(g.stats + ((size_t)n * 32 + c_base + lane) * 2 + 0, (unsigned long long)csum);
                atomicAdd(g.stats + ((size_t)n * 32 + c_base + lane) * 2 + 1, (unsigned long long)csq);
            }
            csum = csq = 0;
        };
        const uint32_t lane_base = ((uint32_t)(ew * 32) << 16) + (uint32_t)c_base;
        auto zero_block = [&](uint32_t taddr) {
            if constexpr (CH == 32) tmem_st_zero_32x32(taddr);
            else tmem_st_zero_32x16(taddr);
        };
        for (uint32_t b = 0; b < (uint32_t)ZS_NB; ++b) zero_block(tmem_base + 32u * b + lane_base);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0)
            for (int z = 0; z < D; ++z) mbar_arrive(&acc_empty[z]);
        for (uint32_t sidx = 0;; ++sidx) {
            const uint32_t slot = sidx % ZR_SCHED;
            mbar_wait(&sched_full[slot], (sidx / ZR_SCHED) & 1u);
            const int col = sched_col[slot];
            __syncwarp();
            if (lane == 0) mbar_arrive(&sched_empty[slot]);
            if (col >= g.n_cols) break;
            int wb, hb, n;
            decode(col, wb, hb, n);
            if (n != cur_n) {
                flush(cur_n);
                cur_n = n;
            }
            const int h = hb * ZR_HT + ew, w = wb * ZR_WT + lane;
            const bool valid = lane < ZR_WT && h < H && w < W;
            for (int z = 0; z < D; ++z) {
                mbar_wait(&acc_full[z], sidx & 1u);
                tc_fence_after();
                uint32_t v[CH];
                const uint32_t taddr = tmem_base + 32u * (uint32_t)(15 - z) + lane_base;
                z32_tmem_ld<CH>(taddr, v);
                tmem_ld_wait();
                zero_block(taddr);                                 // the next column accumulates from zero
                tmem_st_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&acc_empty[z]);
                if (valid) {
                    __half *o = g.out + ((size_t)n * vox_chunk + ((size_t)z * H + h) * W + w) * 32 + c_base;
#pragma unroll
                    for (int qd = 0; qd < CH / 8; ++qd) {
                        uint4 pk;
                        uint32_t *pw = reinterpret_cast<uint32_t *>(&pk);
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            __half2 h2 = __floats2half2_rn(__uint_as_float(v[qd * 8 + e * 2]),
                                                           __uint_as_float(v[qd * 8 + e * 2 + 1]));
                            pw[e] = *reinterpret_cast<uint32_t *>(&h2);
                        }
                        reinterpret_cast<uint4 *>(o)[qd] = pk;
                    }
                }
#pragma unroll
                for (int j = 0; j < CH; ++j) st[lane * (CH + 1) + j] = valid ? __uint_as_float(v[j]) : 0.0f;
                __syncwarp();
                float s = 0.0f, q2 = 0.0f;
                if constexpr (CH == 32) {
#pragma unroll
                    for (int r = 0; r < 32; ++r) {
                        const float x = st[r * 33 + lane];
                        s += x;
                        q2 = fmaf(x, x, q2);
                    }
                } else {
                    const int c = lane & 15, r0 = (lane >> 4) * 16;
#pragma unroll
                    for (int r = 0; r < 16; ++r) {
                        const float x = st[(r0 + r) * 17 + c];
                        s += x;
                        q2 = fmaf(x, x, q2);
                    }
                    s += __shfl_xor_sync(0xFFFFFFFFu, s, 16);
                    q2 += __shfl_xor_sync(0xFFFFFFFFu, q2, 16);
                }
                __syncwarp();
                stat_guard(q2);
                csum += __float2ll_rn(s * 16777216.0f);
                csq += __float2ll_rn(q2 * 16777216.0f);
            }
        }
        flush(cur_n);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, 512);
}

// nn.Conv3d weight (32, 32, 3,3,3) fp32 -> [dy*3 + dx][dz*32 + co][cin] fp16
__global__ void pack_conv_w_zring32_kernel(const float *__restrict__ src, const float *__restrict__ inv_s,
                                           __half *__restrict__ dst) {
    const int total = 9 * Z32_N * 32;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int ci = i & 31;
        const int row = (i >> 5) % Z32_N;
        const int t = i / (32 * Z32_N);                  // dy*3 + dx
        const int dz = row >> 5, co = row & 31;
        dst[i] = __float2half_rn(src[((size_t)co * 32 + ci) * 27 + dz * 9 + t] * inv_s[co]);
    }
}

}  // namespace isg
