// conv3d (k=3, stride 1, zero pad 1) as an implicit GEMM on tcgen05 tensor cores.
//
// Replaces the 16 tensor-core-sized nn.Conv3d calls of ConvModule.forward
// (src/iterseg/unet.py:63-76 used at :93,:96); train-mode BatchNorm statistics
// (unet.py:80-81) are reduced in the epilogue.
//
// Mapping (one CTA per SM, persistent over GROUPS of output tiles):
//   tile : M = 128 output voxels of one z-plane: a patch of Ht rows x P columns, linearised
//          row-major with pitch P (the last 2 columns of every patch row are halo, so
//          Wt = P-2 outputs per row are valid);  N = Cout (16..256);
//          K = 27 taps x Cin, walked as (channel block of CBLK) x (tap).
//   group: T tiles stacked along z (same patch, output planes d0..d0+T-1), one fp32
//          accumulator each in TMEM (T x Cout <= 512 columns).
//   A: one TMA box per (input plane, channel block): the halo patch {CBLK ch, P, Ht+2} of the
//      channels-last activation tensor, out-of-bounds -> 0 (the conv padding), landing in
//      one of T+2 plane slots as rows of CBLK*2 bytes (hardware 128B/64B swizzle).  A group
//      needs T+2 planes, so an input voxel is fetched from L2 ~ (T+2)/T * (Ht+2)*P/(Ht*Wt)
//      times instead of 27; every tap (dz,dy,dx) of tile t is plane slot t+dz+1 read
//      through a K-major UMMA descriptor whose start address is advanced by dy*P + dx rows.
//   B: weights packed [tap][Cout][Cin] fp16, one TMA box {CBLK, Cout, G taps} per stage.
//      The loop is WEIGHT-STATIONARY: a weight stage is used for all T tiles of the group
//      before it is released, so the weight stream from L2 -- what bounded the previous,
//      tile-at-a-time kernel (27*Cout*Cin*2 bytes per 128 voxels) -- shrinks T-fold; thin
//      layers keep all their weights resident in shared memory.
//   FOLD (thin layers, Cout <= 64 with P = 32): a tcgen05.mma costs >= ~45 clocks however small
//      N is (measured, scripts/micro/umma_rate.cu), so with N = Cout = 32 the tensor pipe idles.
//      The three dx taps of a (dz,dy) pair are therefore folded into N: one MMA with
//      N = 3*Cout multiplies the un-shifted A view with the weights of dx = 0,1,2 (they are
//      adjacent in the [tap][Cout][Cin] packing), and the epilogue adds the three column
//      blocks with row shifts 0,1,2 -- out[r] = Y0[r] + Y1[r+1] + Y2[r+2] -- which are warp
//      shuffles because a patch row is exactly one warp (P = 32; lanes 30,31 are halo).
//      27 -> 9 MMAs per channel step.
//   Plane slots and accumulators are released one by one at their last use, so the loads of
//   the next channel block / group and the epilogue overlap the MMAs.
//   SCHEDULER: a CTA starts with group blockIdx.x and then takes the next free group from a global
//      counter (atomicAdd by the A producer, one group ahead of its loads), published to the
//      other roles through a 4-deep shared-memory ring guarded by mbarriers.  The post stage of the
//      previous frame runs beside the U-Net on a second stream and its long flood blocks
//      own whole SMs (227 KB of shared memory) for milliseconds: with a static grid-stride
//      partition the conv CTA that cannot be placed starts late and the whole launch waits for
//      its share; with the counter the placed CTAs simply take the work.
// Warp roles: 0 = A producer + scheduler, 3 = B producer, 1,2 = MMA issuers (one elected lane each;
//             2 also allocates TMEM), 4..7 = epilogue (TMEM -> regs -> global + statistics).
#pragma once
#include <cuda.h>

#include "common.cuh"
#include "sm100.cuh"

namespace isg {

struct ConvGeom {
    int N, D, H, W;              // batch (chunks) and spatial extents (in == out)
    int P, Ht, Wt;               // patch pitch, patch rows, valid outputs per row
    int flat;                    // 1: band-flat tiling -- a tile is 128 CONSECUTIVE positions f = y*P + x' of a
                                 // column band (P = Wt + 2 wide, all H rows), tile fb covers f in [128 fb, 128 fb + 128);
                                 // the plane slot holds the Ht + 2 padded rows from y0 = 128 fb / P on, and every
                                 // tap is still one constant row shift (dy*P + dx) of the A descriptor.  No rows are
                                 // lost to patch-height quantisation: only the 2 halo columns per row and the last
                                 // tile of a band are idle.  tiles_h = tiles per band = ceil(H*P / 128).
    int tiles_w, tiles_h;
    int T;                       // output planes (accumulators) per group
    int nsets;                   // 1: one accumulator set; 2: groups alternate between two sets, so the
                                 // epilogue of a group overlaps the MMAs of the next one
    int dgroups;                 // ceil(D / T)
    int n_groups;                // N * dgroups * tiles_h * tiles_w
    int cout;                    // output channels (multiple of 16, <= 256)
    int acc_cols;                // TMEM columns per accumulator = UMMA N: cout, or 3 * cout when
                                 // the three dx taps are folded into N (template parameter FOLD)
    int nkb0, nkb1;              // channel blocks taken from source 0 / source 1 (concat)
    int plane_rows;              // (Ht + 2) * P
    int plane_bytes;             // plane_rows * CBLK * 2 rounded up to 1024
    int b_stage_bytes;           // taps_per_b * cout * CBLK * 2
    int n_b_stages;
    int taps_per_b;              // taps per B stage (template parameter G): 1, 3 or 9
    int b_resident;              // 1: all (nkb x 27/G) weight stages are loaded once and kept
    int out_mode;                // 0: fp16 [vox][cout]   1: fp16 [vox][8] (first 8 columns)
    int debug;                   // ISG_CONV_DEBUG (diagnosis only): 1 skip epilogue body, 2 skip A loads, 4 skip B loads
    void *out;
    unsigned long long *stats;   // [N][cout][2] (sum, sum of squares) as 2^-24 fixed point:
                                 // integer atomics are order-independent -> reproducible
    unsigned int *sched;         // group counter of this launch (zeroed by the caller): groups beyond the
                                 // first one of a CTA are handed out dynamically, see "scheduler" below
};

// fp16 range guard.  The pre-BatchNorm activations are stored as fp16 (ceiling 65504); the filters
// are prescaled at pack time so that they are O(1), but nothing can rule out pathological
// weights.  Every kernel that reduces BatchNorm sums checks the per-warp sum of 32 squares it has
// anyway: >= 1e9 (some |x| >= 5590), inf or NaN raises this sticky device flag, which the host
// reads back after every forward pass and turns into an error (isg_unet_plan_overflowed) --
// never a silent inf / NaN in the feature volume.
__device__ unsigned int g_unet_overflow;
__device__ __forceinline__ void stat_guard(float q2) {
    if (!(q2 < 1.0e9f)) g_unet_overflow = 1u;
}

static constexpr int CONV_THREADS = 256;
static constexpr float STAT_SCALE = 16777216.0f;      // 2^24
static constexpr int CONV_SLACK = 4096;      // garbage rows the last taps of invalid rows touch
static constexpr int CONV_MAX_T = 10;
static constexpr int CONV_MAX_B_STAGES = 24;
static constexpr int CONV_BAR_BYTES = 1024;
static constexpr int CONV_SCHED_SLOTS = 4;   // ring of published group indices
static_assert((66 + 2 * CONV_MAX_B_STAGES + 2 * CONV_SCHED_SLOTS) * 8 + CONV_SCHED_SLOTS * 4 <= CONV_BAR_BYTES,
              "barrier block too small");

__host__ __device__ inline size_t conv_smem_bytes(const ConvGeom &g) {
    return 1024 /* alignment */ + (size_t)(g.T + 2) * g.plane_bytes +
           (size_t)g.n_b_stages * g.b_stage_bytes + CONV_SLACK + CONV_BAR_BYTES +
           4 * 32 * 33 * sizeof(float);
}

template <int CBLK, int G, bool FOLD>
__global__ void __launch_bounds__(CONV_THREADS, 1)
conv3d_tc_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                 const __grid_constant__ CUtensorMap tmB, const ConvGeom g) {
    using namespace sm100;
    constexpr uint32_t RB = CBLK * 2;            // smem row bytes (128 -> SW128, 64 -> SW64)
    constexpr int KSTEPS = CBLK / 16;
    constexpr int NBG = 27 / G;                  // weight stages per channel block
    extern __shared__ uint8_t smem_dyn[];
    const uint32_t raw_base = smem_u32(smem_dyn);
    const uint32_t pad = ((raw_base + 1023u) & ~1023u) - raw_base;
    uint8_t *base = smem_dyn + pad;
    uint8_t *a_smem = base;
    uint8_t *b_smem = a_smem + (size_t)(g.T + 2) * g.plane_bytes;
    uint8_t *tail = b_smem + (size_t)g.n_b_stages * g.b_stage_bytes + CONV_SLACK;
    uint64_t *bars = reinterpret_cast<uint64_t *>(tail);
    uint64_t *plane_full = bars, *plane_empty = bars + 12;
    uint64_t *acc_full = bars + 24, *acc_empty = bars + 44;            // [nsets * T] <= 20 each
    uint64_t *b_full = bars + 64, *b_empty = bars + 64 + CONV_MAX_B_STAGES;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 64 + 2 * CONV_MAX_B_STAGES);
    uint64_t *sched_full = bars + 66 + 2 * CONV_MAX_B_STAGES, *sched_empty = sched_full + CONV_SCHED_SLOTS;
    volatile int *sched_grp = reinterpret_cast<volatile int *>(sched_empty + CONV_SCHED_SLOTS);
    float *stat_t = reinterpret_cast<float *>(tail + CONV_BAR_BYTES);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int nkb = g.nkb0 + g.nkb1;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmA0);
        prefetch_tmap(&tmA1);
        prefetch_tmap(&tmB);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < g.T + 2; ++i) {
            mbar_init(&plane_full[i], 1);
            mbar_init(&plane_empty[i], 2);        // both MMA-issuing warps commit
        }
        for (int i = 0; i < g.T * g.nsets; ++i) {
            mbar_init(&acc_full[i], 2);
            mbar_init(&acc_empty[i], 4);
        }
        for (int i = 0; i < g.n_b_stages; ++i) {
            mbar_init(&b_full[i], 1);
            mbar_init(&b_empty[i], 2);
        }
        for (int i = 0; i < CONV_SCHED_SLOTS; ++i) {
            mbar_init(&sched_full[i], 1);
            // readers: 2 MMA warps + 4 epilogue warps (+ the B producer when it streams weights)
            mbar_init(&sched_empty[i], g.b_resident ? 6 : 7);
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

    // group index -> (patch column, patch row, z group, chunk)
    auto decode = [&](int grp, int &wb, int &hb, int &d0, int &n, int &tg) {
        int t = grp;
        wb = t % g.tiles_w; t /= g.tiles_w;
        hb = t % g.tiles_h; t /= g.tiles_h;
        const int dg = t % g.dgroups;
        n = t / g.dgroups;
        d0 = dg * g.T;
        tg = g.D - d0 < g.T ? g.D - d0 : g.T;
    };

    if (warp == 0) {
        // ===================== A producer + scheduler =====================
        if (lane == 0) {
            uint32_t ph = 0;                                  // per plane slot: uses so far (parity)
            uint32_t sidx = 0;
            int grp = blockIdx.x;
            for (;;) {
                const uint32_t slot = sidx % CONV_SCHED_SLOTS, sph = (sidx / CONV_SCHED_SLOTS) & 1u;
                mbar_wait(&sched_empty[slot], sph ^ 1u);
                sched_grp[slot] = grp;                        // >= n_groups: the stop mark
                mbar_arrive(&sched_full[slot]);               // release: readers see the store
                ++sidx;
                if (grp >= g.n_groups) break;
                // the following group, requested now and needed after this group's loads
                const int next = (int)gridDim.x + (int)atomicAdd(g.sched, 1u);
                int wb, hb, d0, n, tg;
                decode(grp, wb, hb, d0, n, tg);
                for (int kb = 0; kb < nkb; ++kb) {
                    const bool first = kb < g.nkb0;
                    const CUtensorMap *tm = first ? &tmA0 : &tmA1;
                    const int c = (first ? kb : kb - g.nkb0) * CBLK;
                    for (int pi = 0; pi < tg + 2; ++pi) {
                        mbar_wait(&plane_empty[pi], ((ph >> pi) & 1u) ^ 1u);
                        if ((g.debug & 2) && grp != (int)blockIdx.x) { mbar_arrive(&plane_full[pi]); continue; }
                        mbar_expect_tx(&plane_full[pi], (uint32_t)g.plane_rows * RB);
                        tma_load_5d(a_smem + (size_t)pi * g.plane_bytes, tm, &plane_full[pi], c,
                                    wb * g.Wt - 1, (g.flat ? (hb * 128) / g.P : hb * g.Ht) - 1, d0 - 1 + pi, n);
                    }
                    ph ^= (1u << (tg + 2)) - 1u;
                }
                grp = next;
            }
        }
    } else if (warp == 3) {
        // ===================== B producer =====================
        if (lane == 0) {
            if (g.b_resident) {
                for (int kb = 0; kb < nkb; ++kb)
                    for (int bg = 0; bg < NBG; ++bg) {
                        const int s = kb * NBG + bg;
                        mbar_expect_tx(&b_full[s], (uint32_t)g.b_stage_bytes);
                        tma_load_3d(b_smem + (size_t)s * g.b_stage_bytes, &tmB, &b_full[s], kb * CBLK, 0,
                                    bg * G);
                    }
            } else {
                uint32_t it = 0, sidx = 0;
                const uint32_t nb = (uint32_t)g.n_b_stages;
                for (;; ++sidx) {
                    const uint32_t slot = sidx % CONV_SCHED_SLOTS;
                    mbar_wait(&sched_full[slot], (sidx / CONV_SCHED_SLOTS) & 1u);
                    const int grp = sched_grp[slot];
                    mbar_arrive(&sched_empty[slot]);
                    if (grp >= g.n_groups) break;
                    for (int kb = 0; kb < nkb; ++kb) {
                        for (int bg = 0; bg < NBG; ++bg, ++it) {
                            const uint32_t s = it % nb, ph = (it / nb) & 1u;
                            mbar_wait(&b_empty[s], ph ^ 1u);
                            if ((g.debug & 4) && it >= nb) { mbar_arrive(&b_full[s]); continue; }
                            mbar_expect_tx(&b_full[s], (uint32_t)g.b_stage_bytes);
                            tma_load_3d(b_smem + (size_t)s * g.b_stage_bytes, &tmB, &b_full[s], kb * CBLK,
                                        0, bg * G);
                        }
                    }
                }
            }
        }
    } else if (warp == 1 || warp == 2) {
        // ===================== MMA issuers (two warps) =====================
        // Tiles are split by parity between warps 1 and 2 (different SM sub-partitions): one
        // thread issuing every tcgen05.mma of the CTA is what bounded the previous version.  A
        // tile's accumulator is only ever touched by one issuer, so no ordering between the two
        // is needed; BOTH warps walk the identical loops and commit to every barrier (count 2) --
        // a commit covers the MMAs its own thread issued, and each warp reaches a release point
        // only after its own last use of that plane / weight stage.
        // The whole warp walks the loops (so that all the address arithmetic is warp-uniform)
        // and one elected lane issues; what bounds thin layers is the instruction count per
        // tcgen05.mma: descriptors are (constant high word) | (start address >> 4) and the tap
        // loops are fully unrolled, leaving ~2 integer adds per MMA.
        {
            const bool leader = elect_one();
            const int par = warp - 1;                          // this issuer's tiles: t % 2 == par
            const uint32_t idesc = make_idesc_f16(128, (uint32_t)g.acc_cols, 0 /* fp16 */);
            const uint32_t nb = (uint32_t)g.n_b_stages;
            uint32_t plane_ph = 0, acc_ph = 0;
            bool b_waited = false;
            const uint64_t dproto = make_kmajor_desc(0, RB, 0);
            const uint32_t d_hi = (uint32_t)(dproto >> 32), d_lo = (uint32_t)dproto;
            constexpr uint32_t U = RB >> 4;                       // one row in 16-byte units
            const uint32_t cy = (uint32_t)g.P * U;
            const uint32_t plane_units = (uint32_t)g.plane_bytes >> 4;
            const uint32_t b_tap_units = (uint32_t)(g.cout * CBLK * 2) >> 4;
            const uint32_t a_lo0 = d_lo | (smem_u32(a_smem) >> 4);
            const int flat = g.flat, tiles_w = g.tiles_w, tiles_h = g.tiles_h, Pp = g.P;
            const uint32_t b_lo0 = d_lo | (smem_u32(b_smem) >> 4);
            const uint32_t b_stage_units = (uint32_t)g.b_stage_bytes >> 4;
            const uint32_t acc_cols = (uint32_t)g.acc_cols;
            const int T = g.T, D = g.D, nsets = g.nsets, n_groups = g.n_groups, b_resident = g.b_resident;
            const int dgroups = g.dgroups, tiles_hw = g.tiles_h * g.tiles_w;
            uint32_t gcount = 0, bs = 0, bph = 0;                  // weight ring position / parity
            for (;; ++gcount) {
                const uint32_t slot = gcount % CONV_SCHED_SLOTS;
                mbar_wait(&sched_full[slot], (gcount / CONV_SCHED_SLOTS) & 1u);
                const int grp = sched_grp[slot];
                __syncwarp();                                  // every lane has read the slot
                if (leader) mbar_arrive(&sched_empty[slot]);
                if (grp >= n_groups) break;
                const int dg = (grp / tiles_hw) % dgroups;
                const int d0 = dg * T;
                const int tg = D - d0 < T ? D - d0 : T;
                const int a0 = nsets == 2 ? (int)(gcount & 1u) * T : 0;   // accumulator set of this group
                const uint32_t tmem_g = tmem_base + (uint32_t)a0 * acc_cols;
                // band-flat tiling: the tile starts (128 fb) % P rows into the slot's first padded row
                const uint32_t a_lo = a_lo0 + (flat ? (uint32_t)((((grp / tiles_w) % tiles_h) * 128) % Pp) * U : 0u);
                if (b_resident) {
                    // ---- tile-major: all taps of a tile back to back (weights never move) ----
                    for (int kb = 0; kb < nkb; ++kb) {
                        int ready = 0;                             // plane slots known to have landed
                        uint32_t tmem_d = tmem_g, a_pl = a_lo;
                        for (int t = 0; t < tg; ++t, tmem_d += acc_cols, a_pl += plane_units) {
                            if (kb == 0) mbar_wait(&acc_empty[a0 + t], ((acc_ph >> (a0 + t)) & 1u) ^ 1u);
                            while (ready <= t + 2) {
                                mbar_wait(&plane_full[ready], (plane_ph >> ready) & 1u);
                                ++ready;
                            }
                            if (!b_waited)
                                for (int bg = 0; bg < NBG; ++bg) mbar_wait(&b_full[kb * NBG + bg], 0u);
                            tc_fence_after();
                            if (leader && (t & 1) == par) {
                                const uint32_t b_kb = b_lo0 + (uint32_t)(kb * NBG) * b_stage_units;
#pragma unroll
                                for (int tap = 0; tap < 27; tap += (FOLD ? 3 : 1)) {
                                    const int dz = tap / 9, dy = (tap / 3) % 3, dx = tap % 3;   // FOLD: dx = 0
                                    const uint32_t a_tap = a_pl + dz * plane_units + dy * cy + dx * U;
                                    const uint32_t b_tap = b_kb + (tap / G) * b_stage_units + (tap % G) * b_tap_units;
#pragma unroll
                                    for (int k = 0; k < KSTEPS; ++k) {
                                        const uint64_t adesc = ((uint64_t)d_hi << 32) | (a_tap + 2 * k);
                                        const uint64_t bdesc = ((uint64_t)d_hi << 32) | (b_tap + 2 * k);
                                        umma_f16(tmem_d, adesc, bdesc, idesc, (tap | k) != 0 ? 1u : (kb != 0 ? 1u : 0u));
                                    }
                                }
                            }
                            if (leader) {
                                umma_commit(&plane_empty[t]);                 // last use of plane t
                                if (t == tg - 1) {
                                    umma_commit(&plane_empty[tg]);
                                    umma_commit(&plane_empty[tg + 1]);
                                }
                                if (kb == nkb - 1) umma_commit(&acc_full[a0 + t]);
                            }
                            __syncwarp();
                        }
                        b_waited = b_waited || kb == nkb - 1;
                        plane_ph ^= (1u << (tg + 2)) - 1u;
                    }
                    acc_ph ^= ((1u << tg) - 1u) << a0;
                    continue;
                }
                // ---- weight-stationary: a weight stage serves all tiles of the group ----
                for (int kb = 0; kb < nkb; ++kb) {
                    int ready = 0;
#pragma unroll
                    for (int bg = 0; bg < NBG; ++bg) {
                        mbar_wait(&b_full[bs], bph);
                        const uint32_t b_lo = b_lo0 + bs * b_stage_units;
                        const int dz = (bg * G) / 9;               // G <= 9: one dz per stage
                        uint32_t tmem_d = tmem_g, a_pl = a_lo + (uint32_t)dz * plane_units;
                        for (int t = 0; t < tg; ++t, tmem_d += acc_cols, a_pl += plane_units) {
                            if (kb == 0 && bg == 0) mbar_wait(&acc_empty[a0 + t], ((acc_ph >> (a0 + t)) & 1u) ^ 1u);
                            while (ready <= t + dz) {
                                mbar_wait(&plane_full[ready], (plane_ph >> ready) & 1u);
                                ++ready;
                            }
                            tc_fence_after();
                            if (leader && (t & 1) == par) {
#pragma unroll
                                for (int j = 0; j < G; j += (FOLD ? 3 : 1)) {
                                    const int tap0 = bg * G;
                                    const int dy = ((tap0 + j) / 3) % 3, dx = (tap0 + j) % 3;   // FOLD: dx = 0
                                    const uint32_t a_tap = a_pl + dy * cy + dx * U;
                                    const uint32_t b_tap = b_lo + j * b_tap_units;
#pragma unroll
                                    for (int k = 0; k < KSTEPS; ++k) {
                                        const uint64_t adesc = ((uint64_t)d_hi << 32) | (a_tap + 2 * k);
                                        const uint64_t bdesc = ((uint64_t)d_hi << 32) | (b_tap + 2 * k);
                                        umma_f16(tmem_d, adesc, bdesc, idesc,
                                                 (bg | j | k) != 0 ? 1u : (kb != 0 ? 1u : 0u));
                                    }
                                }
                            }
                            if (leader) {
                                if (bg == NBG - 1) {
                                    umma_commit(&plane_empty[t + 2]);          // last use of plane t+2
                                    if (kb == nkb - 1) umma_commit(&acc_full[a0 + t]);
                                }
                            }
                        }
                        if (leader) {
                            if (bg == NBG / 3 - 1) umma_commit(&plane_empty[0]);       // end of the dz=-1 taps
                            if (bg == 2 * NBG / 3 - 1) umma_commit(&plane_empty[1]);   // end of the dz=0 taps
                            umma_commit(&b_empty[bs]);
                        }
                        if (++bs == nb) { bs = 0; bph ^= 1u; }
                        __syncwarp();
                    }
                    plane_ph ^= (1u << (tg + 2)) - 1u;
                }
                acc_ph ^= ((1u << tg) - 1u) << a0;
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
        const int row = ew * 32 + lane;
        int hy = row / g.P, wx = row - hy * g.P;
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
        uint32_t acc_ph = 0, gcount = 0;
        for (;; ++gcount) {
          const uint32_t slot = gcount % CONV_SCHED_SLOTS;
          mbar_wait(&sched_full[slot], (gcount / CONV_SCHED_SLOTS) & 1u);
          const int grp = sched_grp[slot];
          __syncwarp();
          if (lane == 0) mbar_arrive(&sched_empty[slot]);
          if (grp >= g.n_groups) break;
          int wb, hb, d0, n, tg;
          decode(grp, wb, hb, d0, n, tg);
          const int a0 = g.nsets == 2 ? (int)(gcount & 1u) * g.T : 0;
          if (n != cur_n) {
              flush(cur_n);
              cur_n = n;
          }
          int h, w;
          bool valid;
          if (g.flat) {
              const int f = hb * 128 + row;
              h = f / g.P;
              wx = f - h * g.P;
              w = wb * g.Wt + wx;
              valid = wx < g.Wt && h < g.H && w < g.W;
          } else {
              h = hb * g.Ht + hy;
              w = wb * g.Wt + wx;
              valid = hy < g.Ht && wx < g.Wt && h < g.H && w < g.W;
          }
          for (int ti = 0; ti < tg; ++ti) {
            const int d = d0 + ti;
            const size_t vox = (((size_t)n * g.D + d) * g.H + h) * g.W + w;
            mbar_wait(&acc_full[a0 + ti], (acc_ph >> (a0 + ti)) & 1u);
            tc_fence_after();
            if (g.debug & 1) {
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&acc_empty[a0 + ti]);
                continue;
            }
            const uint32_t taddr = tmem_base + (uint32_t)((a0 + ti) * g.acc_cols) + ((uint32_t)(ew * 32) << 16);
            if (g.cout >= 32) {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    if (i * 32 < g.cout) {
                        uint32_t v[32];
                        tmem_ld_32x32(taddr + i * 32, v);
                        if (FOLD) {
                            uint32_t v1[32], v2[32];
                            tmem_ld_32x32(taddr + g.cout + i * 32, v1);
                            tmem_ld_32x32(taddr + 2 * g.cout + i * 32, v2);
                            tmem_ld_wait();
#pragma unroll
                            for (int j = 0; j < 32; ++j)
                                v[j] = __float_as_uint(__uint_as_float(v[j]) +
                                                       __shfl_down_sync(0xFFFFFFFFu, __uint_as_float(v1[j]), 1) +
                                                       __shfl_down_sync(0xFFFFFFFFu, __uint_as_float(v2[j]), 2));
                        } else {
                            tmem_ld_wait();
                        }
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
                        stat_guard(q2);
                        csum[i] += __float2ll_rn(s * STAT_SCALE);
                        csq[i] += __float2ll_rn(q2 * STAT_SCALE);
                    }
                }
            } else {
                uint32_t v[16];
                tmem_ld_32x16(taddr, v);
                if (FOLD) {
                    uint32_t v1[16], v2[16];
                    tmem_ld_32x16(taddr + g.cout, v1);
                    tmem_ld_32x16(taddr + 2 * g.cout, v2);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        v[j] = __float_as_uint(__uint_as_float(v[j]) +
                                               __shfl_down_sync(0xFFFFFFFFu, __uint_as_float(v1[j]), 1) +
                                               __shfl_down_sync(0xFFFFFFFFu, __uint_as_float(v2[j]), 2));
                } else {
                    tmem_ld_wait();
                }
                if (valid) {
                    uint4 pk;
                    __half2 *h2 = reinterpret_cast<__half2 *>(&pk);
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                        h2[e] = __floats2half2_rn(__uint_as_float(v[2 * e]), __uint_as_float(v[2 * e + 1]));
                    *reinterpret_cast<uint4 *>(reinterpret_cast<__half *>(g.out) + vox * 8) = pk;
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
                stat_guard(q2);
                csum[0] += __float2ll_rn(s * STAT_SCALE);
                csq[0] += __float2ll_rn(q2 * STAT_SCALE);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[a0 + ti]);
          }
          acc_ph ^= ((1u << tg) - 1u) << a0;
        }
        flush(cur_n);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, 512);
}

}  // namespace isg
