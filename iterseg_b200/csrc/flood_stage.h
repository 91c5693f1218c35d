// Shared flood stage (CCL + component grouping + single-seed fill + ordered flood):
// declarations only, the kernels live in flood.cu.
#pragma once
#include "common.cuh"

namespace isg {

static constexpr uint32_t CCL_NONE = 0xFFFFFFFFu;

struct FloodGeom {
    const float *aff;          // 3 planes
    int64_t plane_stride;
    int origin;                // 0: planes padded like labels, 1: unpadded
    uint32_t za, ya, xa;       // plane extents
    const float *div;          // device float[3]
    float scale[3];
    uint32_t zp, yp, xp;
    // node-keyed mode (classic marker watershed of the DoG blob path): when non-null, a voxel is
    // queued with node_key[voxel] (order-preserving uint32, smaller pops first) instead of the
    // affinity of the edge it was claimed through, and seeds enter with their own key (not 0.0)
    const uint32_t *node_key;
};

struct FloodStageBuffers {
    uint64_t *keys_a, *keys_b;
    uint32_t *vals_a, *vals_b;
    uint32_t *comp_start;
    uint64_t *arena_off;
    uint32_t *order_keys_a, *order_keys_b, *order_a, *order_b;   // components sorted by size
    uint32_t *cbase, *ccursor;                                   // compact arena slices per component
    int64_t max_seeds;
    uint32_t *scalars;          // see comp_group_kernel
    unsigned char *cub_tmp;
    size_t cub_bytes;
    uint64_t *arena_keys;
    uint32_t *arena_idx;
    uint64_t arena_cap;
    uint64_t node_cap;                                           // compact nodes the arenas hold (<= npix)
    uint32_t *lidmap, *vox, *key, *nbr, *nlab;                   // compact component graphs
    uint4 *rec;                                                  // bucket-queue node records
    uint32_t *ebase;                                             // edge-arena segment per component
    uint64_t *ekeys_a, *ekeys_b;                                 // edge arena: (component << 32 | key)
    uint32_t *evals_a, *evals_b;                                 // (sort double buffers)
    uint64_t edge_cap;
    uint32_t *seedpos, *seedgs, *complab;
    unsigned char *seg_tmp;
    size_t seg_bytes;
};

// 6-connected components of the non-zero voxels of `dom` (padded volume):
// parent[v] = smallest flat index of v's component (CCL_NONE outside),
// comp_size[root] = voxel count (comp_size must be zero on entry).
int ccl_run(const uint8_t *dom, uint32_t *parent, uint32_t *comp_size, uint32_t zp, uint32_t yp,
            uint32_t xp, cudaStream_t st);

// node_cap: how many voxels of multi-seed components the compact arenas can hold (clamped to
// [1, npix]; npix = the worst case, every voxel).  A run that needs more fails with
// ISG_ERR_WORKSPACE and leaves the multi-seed components unflooded.
size_t flood_stage_workspace(FloodStageBuffers *b, Carver &cv, uint64_t npix, int64_t max_seeds,
                             uint64_t node_cap);

// parent: flattened CCL roots of the flood domain (CCL_NONE outside);
// comp_size: voxels per root; comp_label: zeroed scratch indexed by root;
// seeds: padded flat indices in label order, labels[seed] already set; n_seeds is the
// host-side (upper bound on the) count, n_seeds_dev the exact device-side count or NULL.
// seed_labels: NULL (seed i carries label i + 1) or the label of every seed (several seeds may
// share one: multi-voxel markers).
int flood_stage_run(const FloodStageBuffers &b, const FloodGeom &geom, const uint8_t *mask,
                    const uint32_t *parent, const uint32_t *comp_size, uint32_t *comp_label,
                    const int64_t *seeds, int64_t n_seeds, const uint32_t *n_seeds_dev,
                    uint32_t *labels, cudaStream_t st, const uint32_t *seed_labels = nullptr);

}  // namespace isg
