// Affinity-keyed priority flood, exact w.r.t. the reference heap order.
//
// Replaces raveled_affinity_watershed (src/iterseg/watershed.py:95-159).
//
// Exact decomposition (verified against the reference):
//   * the flood never crosses between 6-connected components of the flood
//     domain (mask, plus seed voxels), and the (value, age) order inside a
//     component does not depend on what happens in other components, so every
//     component is flooded on its own with a local age counter;
//   * a component holding exactly one seed is filled with that seed's label
//     (a plain parallel pass); a component without seeds stays 0;
//   * only components with >= 2 seeds run the ordered flood: one warp per
//     component, a 32-ary min-heap of 64-bit keys (order-preserving float bits
//     << 32 | age) in shared memory (global arena for big components), the six
//     neighbour tests of a popped voxel done by six lanes at once.
// Seeds carry value 0.0 / age 0 in the reference and are ordered by flat index
// (third tuple field, watershed.py:162); here they get ages 0..k-1 in index
// order and pushes continue from k, which preserves every comparison.
#pragma once
#include "ccl.cuh"
#include "flood_stage.h"

namespace isg {

static constexpr int FLOOD_SMEM_ENTRIES = 2048;      // 24 KB per warp-CTA


struct FloodWork {
    const uint64_t *seed_keys;     // sorted (root << 32 | padded flat index)
    const uint32_t *seed_labels;   // label of the seed at the same sorted position
    const uint32_t *comp_start;    // n_comp + 1 offsets into the sorted arrays
    const uint64_t *arena_off;     // per component offset into the heap arena
    const uint32_t *n_comp;        // device scalar
    uint64_t *arena_keys;
    uint32_t *arena_idx;
    uint32_t *counter;             // work-stealing cursor (zeroed by the caller)
};

__device__ __forceinline__ float flood_key_value(const FloodGeom &g, int axis, float div,
                                                 float scale, uint32_t z, uint32_t y,
                                                 uint32_t x) {
    // (z,y,x) are padded coordinates; out-of-plane reads are the implicit zero pad
    uint32_t az = z - g.origin, ay = y - g.origin, ax = x - g.origin;
    float a = 0.0f;
    if (az < g.za && ay < g.ya && ax < g.xa)
        a = __ldg(g.aff + (int64_t)axis * g.plane_stride + ((uint64_t)az * g.ya + ay) * g.xa + ax);
    // IEEE division, then the optional |scale| multiply (watershed.py:195, :23-24);
    // "+ 0.0f" folds -0.0 into +0.0, which compare equal in the reference
    float v = __fmul_rn(__fdiv_rn(a, div), scale);
    return v + 0.0f;
}

__global__ void __launch_bounds__(32)
flood_components_kernel(FloodGeom g, FloodWork w, const uint8_t *__restrict__ mask,
                        uint32_t *labels) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const unsigned FULL = 0xFFFFFFFFu;
    const unsigned lane = threadIdx.x;
    const uint32_t n_comp = *w.n_comp;
    const uint64_t plane = (uint64_t)g.yp * g.xp;
    const uint64_t npix = plane * g.zp;

    // neighbour order and key axis exactly as watershed.py:84-92:
    // offsets [-YX, -X, -1, +1, +X, +YX], axes [0, 1, 2, 2, 1, 0]
    int dz = 0, dy = 0, dx = 0, axis = 0;
    switch (lane) {
        case 0: dz = -1; axis = 0; break;
        case 1: dy = -1; axis = 1; break;
        case 2: dx = -1; axis = 2; break;
        case 3: dx = +1; axis = 2; break;
        case 4: dy = +1; axis = 1; break;
        case 5: dz = +1; axis = 0; break;
        default: break;
    }
    const int64_t my_off = (int64_t)dz * (int64_t)plane + (int64_t)dy * g.xp + dx;
    const float my_div = __ldg(g.div + axis);
    const float my_scale = axis == 0 ? g.scale[0] : (axis == 1 ? g.scale[1] : g.scale[2]);

    for (;;) {
        uint32_t c = 0;
        if (lane == 0) c = atomicAdd(w.counter, 1u);
        c = __shfl_sync(FULL, c, 0);
        if (c >= n_comp) break;
        const uint32_t s0 = w.comp_start[c], s1 = w.comp_start[c + 1];
        const uint32_t cnt = s1 - s0;
        if (cnt < 2) continue;
        const uint64_t cap = w.arena_off[c + 1] - w.arena_off[c];
        uint64_t *keys;
        uint32_t *idx;
        if (cap <= (uint64_t)FLOOD_SMEM_ENTRIES) {
            keys = reinterpret_cast<uint64_t *>(smem_raw);
            idx = reinterpret_cast<uint32_t *>(smem_raw + sizeof(uint64_t) * FLOOD_SMEM_ENTRIES);
        } else {
            keys = w.arena_keys + w.arena_off[c];
            idx = w.arena_idx + w.arena_off[c];
        }
        const uint64_t zero_hi = (uint64_t)f32_ord(0.0f) << 32;
        for (uint32_t i = lane; i < cnt; i += 32) {
            keys[i] = zero_hi | i;                       // ascending array == valid heap
            idx[i] = (uint32_t)(w.seed_keys[s0 + i] & 0xFFFFFFFFu);
        }
        uint32_t n = cnt;
        uint32_t age = cnt;
        __syncwarp();

        while (n > 0) {
            // ---- pop the minimum -------------------------------------------------
            const uint32_t p = idx[0];
            --n;
            const uint64_t lastk = keys[n];
            const uint32_t lasti = idx[n];
            __syncwarp();                                // all lanes hold p / last before slot 0 changes
            if (n > 0) {
                uint32_t i = 0;
                for (;;) {
                    const uint32_t c0 = i * 32u + 1u;
                    if (c0 >= n) break;
                    const uint32_t ch = c0 + lane;
                    const uint64_t k = ch < n ? keys[ch] : ~0ull;
                    const uint32_t hi = (uint32_t)(k >> 32);
                    const uint32_t mhi = __reduce_min_sync(FULL, hi);
                    const uint32_t lo = hi == mhi ? (uint32_t)k : 0xFFFFFFFFu;
                    const uint32_t mlo = __reduce_min_sync(FULL, lo);
                    const uint64_t mk = ((uint64_t)mhi << 32) | mlo;
                    if (mk >= lastk) break;
                    const uint32_t win = __ffs(__ballot_sync(FULL, hi == mhi && lo == mlo)) - 1;
                    const uint32_t wc = c0 + win;
                    if (lane == 0) {
                        keys[i] = mk;
                        idx[i] = idx[wc];
                    }
                    __syncwarp();
                    i = wc;
                }
                if (lane == 0) {
                    keys[i] = lastk;
                    idx[i] = lasti;
                }
                __syncwarp();
            }
            // ---- expand the popped voxel (watershed.py:135-154) -------------------
            const uint32_t lab = labels[p];
            const uint32_t pz = (uint32_t)(p / plane);
            const uint32_t prem = (uint32_t)(p - (uint64_t)pz * plane);
            const uint32_t py = prem / g.xp;
            const uint32_t px = prem - py * g.xp;
            const int64_t nb = (int64_t)p + my_off;
            const bool valid = lane < 6 && nb >= 0 && (uint64_t)nb < npix;
            bool claim = false;
            float val = 0.0f;
            if (valid) {
                const uint8_t m = mask[nb];
                const uint32_t l = labels[nb];
                // key = affinity of the edge (popped, neighbour): stored at the popped voxel
                // for the three negative directions, at the neighbour for the positive ones
                const bool neg = lane < 3;
                val = flood_key_value(g, axis, my_div, my_scale, neg ? pz : pz + dz,
                                      neg ? py : py + dy, neg ? px : px + dx);
                claim = m != 0 && l == 0;
            }
            unsigned bits = __ballot_sync(FULL, claim);
            if (claim) labels[nb] = lab;                 // labelled at push time (:149)
            const uint64_t mykey = ((uint64_t)f32_ord(val) << 32) |
                                   (uint64_t)(age + __popc(bits & ((1u << lane) - 1u)));
            age += __popc(bits);
            while (bits) {
                const int src = __ffs(bits) - 1;
                bits &= bits - 1;
                const uint64_t k = __shfl_sync(FULL, mykey, src);
                const uint32_t ix = (uint32_t)__shfl_sync(FULL, (uint32_t)nb, src);
                uint32_t i = n++;
                while (i > 0) {
                    const uint32_t par = (i - 1u) >> 5;
                    const uint64_t pk = keys[par];
                    if (pk <= k) break;
                    if (lane == 0) {
                        keys[i] = pk;
                        idx[i] = idx[par];
                    }
                    i = par;
                }
                if (lane == 0) {
                    keys[i] = k;
                    idx[i] = ix;
                }
                __syncwarp();
            }
            __syncwarp();
        }
    }
}

}  // namespace isg
