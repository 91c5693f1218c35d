// 6-connected component labelling of a padded (Zp,Yp,Xp) byte volume by
// concurrent union-find (atomicMin hooking), warp-level run merging along x.
//
// Replaces scipy.ndimage.label as used by _remove_unwanted_objects
// (src/iterseg/watershed.py:239-240).  Component numbering is free (the
// reference only uses sizes and membership), so a component is identified by
// the smallest flat index it contains (its root).
//
// HBM-bound integer work: one byte read + one u32 write per voxel for init, then
// neighbour unions only where a run starts (the redundant ones are skipped).
#pragma once
#include "flood_stage.h"

namespace isg {

__device__ __forceinline__ uint32_t ccl_find(const uint32_t *parent, uint32_t x) {
    // parent pointers only ever decrease; a plain (L2) read is enough
    uint32_t p = __ldcg(parent + x);
    while (p != x) {
        x = p;
        p = __ldcg(parent + x);
    }
    return x;
}

__device__ __forceinline__ void ccl_union(uint32_t *parent, uint32_t a, uint32_t b) {
    for (;;) {
        a = ccl_find(parent, a);
        b = ccl_find(parent, b);
        if (a == b) return;
        if (a < b) { uint32_t t = a; a = b; b = t; }        // a > b: hook a under b
        uint32_t old = atomicMin(parent + a, b);
        if (old == a) return;
        a = old;
    }
}

// One warp handles 32 consecutive voxels of a row: every voxel of an x-run is
// pointed at the first voxel of the run that lies in the same 32-segment.
__global__ void __launch_bounds__(256)
ccl_init_kernel(const uint8_t *__restrict__ dom, uint32_t *__restrict__ parent, uint64_t n) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const unsigned lane = threadIdx.x & 31;
    for (uint64_t base = i - lane; base < n; base += stride) {
        uint64_t v = base + lane;
        bool in = v < n && dom[v] != 0;
        unsigned bits = __ballot_sync(0xFFFFFFFFu, in);
        if (in) {
            unsigned below = ~bits & ((1u << lane) - 1u);   // zeros below me
            unsigned start = below ? (32u - __clz(below)) : 0u;
            parent[v] = (uint32_t)(base + start);
        } else if (v < n) {
            parent[v] = CCL_NONE;
        }
    }
}

// Unions across 32-segment boundaries in x, and along y and z.  A (v, v-stride)
// union is skipped when the pair one step to the left is also inside the domain
// and inside the same row: that pair already makes the same connection.
__global__ void __launch_bounds__(256)
ccl_union_kernel(const uint8_t *__restrict__ dom, uint32_t *parent,
                 uint32_t zp, uint32_t yp, uint32_t xp) {
    const uint64_t n = (uint64_t)zp * yp * xp;
    const uint64_t plane = (uint64_t)yp * xp;
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; v < n; v += stride) {
        if (!dom[v]) continue;
        uint32_t x = (uint32_t)(v % xp);
        bool left = x > 0 && dom[v - 1];
        if (left && (v & 31u) == 0) ccl_union(parent, (uint32_t)v, (uint32_t)(v - 1));
        if (v >= xp) {
            uint64_t u = v - xp;
            if (dom[u] && !(left && dom[u - 1])) ccl_union(parent, (uint32_t)v, (uint32_t)u);
        }
        if (v >= plane) {
            uint64_t u = v - plane;
            if (dom[u] && !(left && dom[u - 1])) ccl_union(parent, (uint32_t)v, (uint32_t)u);
        }
    }
}

// Path compression to the root + component sizes (warp-aggregated atomics).
__global__ void __launch_bounds__(256)
ccl_flatten_count_kernel(uint32_t *parent, uint32_t *__restrict__ comp_size, uint64_t n) {
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const unsigned lane = threadIdx.x & 31;
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (uint64_t base = i - lane; base < n; base += stride) {
        uint64_t v = base + lane;
        uint32_t r = CCL_NONE;
        if (v < n) {
            uint32_t p = parent[v];
            if (p != CCL_NONE) {
                r = ccl_find(parent, p);
                parent[v] = r;
            }
        }
        // NB: writing parent[v] = root while other threads still chase pointers is
        // safe: roots never change in this kernel and every stored value stays an
        // ancestor of v.
        unsigned active = __ballot_sync(0xFFFFFFFFu, r != CCL_NONE);
        if (r != CCL_NONE) {
            unsigned peers = __match_any_sync(active, r);
            if ((unsigned)(__ffs(peers) - 1) == lane) atomicAdd(comp_size + r, (uint32_t)__popc(peers));
        }
    }
}

}  // namespace isg
