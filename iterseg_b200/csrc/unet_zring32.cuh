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
//             4..7 = epilogue.  Columns are handed out dynamically (see unet_conv.cuh).
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
static constexpr int Z32_THREADS = 256;
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
           4 * 32 * 33 * sizeof(float);
}

__global__ void __launch_bounds__(Z32_THREADS, 1)
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
        const uint32_t idesc = make_idesc_f16(128, Z32_N, 0 /* fp16 */);
        const uint64_t dproto = make_kmajor_desc(0, 64, 0);
        const uint32_t d_hi = (uint32_t)(dproto >> 32), d_lo = (uint32_t)dproto;
        const uint32_t a_lo = d_lo | (smem_u32(a_smem) >> 4);
        const uint32_t b_lo = d_lo | (smem_u32(b_smem) >> 4);
        constexpr uint32_t U = 64 >> 4;
        uint32_t lc = 0, pc = 0;
        mbar_wait(b_full, 0u);
        for (uint32_t sidx = 0;; ++sidx) {
            const uint32_t slot = sidx % ZR_SCHED;
            mbar_wait(&sched_full[slot], (sidx / ZR_SCHED) & 1u);
            const int col = sched_col[slot];
            __syncwarp();
            if (leader) mbar_arrive(&sched_empty[slot]);
            if (col >= g.n_cols) break;
            for (int p = 0; p < D; ++p, ++pc, ++lc) {
                const uint32_t acc = pc % Z32_NA, s = lc % Z32_SLOTS;
                mbar_wait(&acc_empty[acc], ((pc / Z32_NA) & 1u) ^ 1u);
                mbar_wait(&plane_full[s], (lc / Z32_SLOTS) & 1u);
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
    } else if (warp >= 4) {
        // ===================== epilogue =====================
        const int ew = warp - 4;                         // TMEM lanes 32*ew.. = patch row ew
        float *st = stat_t + ew * (32 * 33);
        long long csum = 0, csq = 0;                     // channel `lane` of the current chunk
        int cur_n = -1;
        auto flush = [&](int n) {
            if (n >= 0) {
                atomicAdd(g.stats + ((size_t)n * 32 + lane) * 2 + 0, (unsigned long long)csum);
                atomicAdd(g.stats + ((size_t)n * 32 + lane) * 2 + 1, (unsigned long long)csq);
            }
            csum = csq = 0;
        };
        const uint32_t lane_base = (uint32_t)(ew * 32) << 16;
        uint32_t pc0 = 0;
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
                while (waited <= need) {
                    const uint32_t qq = pc0 + (uint32_t)waited;
                    mbar_wait(&acc_full[qq % Z32_NA], (qq / Z32_NA) & 1u);
                    ++waited;
                }
                tc_fence_after();
                uint32_t v[32], v0[32], v2[32];
                tmem_ld_32x32(tmem_base + ((pc0 + (uint32_t)z) % Z32_NA) * Z32_N + 32 + lane_base, v);
                if (z >= 1) tmem_ld_32x32(tmem_base + ((pc0 + (uint32_t)z - 1u) % Z32_NA) * Z32_N + lane_base, v0);
                if (z + 1 < D) tmem_ld_32x32(tmem_base + ((pc0 + (uint32_t)z + 1u) % Z32_NA) * Z32_N + 64 + lane_base, v2);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    float x = __uint_as_float(v[j]);
                    if (z >= 1) x += __uint_as_float(v0[j]);
                    if (z + 1 < D) x += __uint_as_float(v2[j]);
                    v[j] = __float_as_uint(x);
                }
                if (valid) {
                    __half *o = g.out + ((size_t)n * vox_chunk + ((size_t)z * H + h) * W + w) * 32;
#pragma unroll
                    for (int qd = 0; qd < 4; ++qd) {
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
