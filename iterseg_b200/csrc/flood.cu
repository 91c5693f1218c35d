// Host orchestration of the component-wise exact flood + the C-ABI entry
// isg_affinity_flood (replaces affinity_watershed / _prep_data /
// raveled_affinity_watershed, src/iterseg/watershed.py:17-159).
#include <cub/cub.cuh>

#include "flood.cuh"

namespace isg {

// ---------------------------------------------------------------------------
// small kernels
// ---------------------------------------------------------------------------

// output[raveled_markers] = 1..N (watershed.py:61-62); duplicates: the last
// (= largest) label wins, as with numpy fancy assignment.
__global__ void seed_label_kernel(const int64_t *__restrict__ seeds, int64_t n, uint32_t *labels,
                                  uint8_t *dom, uint64_t npix) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int64_t s = seeds[i];
    if (s < 0 || (uint64_t)s >= npix) return;
    atomicMax(labels + s, (uint32_t)(i + 1));
    if (dom) dom[s] = 1;
}

// flood domain of the generic entry point: claimable voxels (in mask and not
// pre-labelled); seed voxels are added by seed_label_kernel afterwards.
__global__ void domain_kernel(const uint8_t *__restrict__ mask, const uint32_t *__restrict__ labels,
                              uint8_t *__restrict__ dom, uint64_t n) {
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; v < n; v += stride)
        dom[v] = (mask[v] != 0 && labels[v] == 0) ? 1 : 0;
}

// key = root << 32 | padded flat index, value = label; seeds outside the
// domain (cannot happen through the public entry points) sort to the end.
__global__ void seed_key_kernel(const int64_t *__restrict__ seeds, int64_t n,
                                const uint32_t *__restrict__ n_dev,
                                const uint32_t *__restrict__ parent, uint64_t npix,
                                uint64_t *__restrict__ keys, uint32_t *__restrict__ vals,
                                const uint32_t *__restrict__ seed_labels) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const bool live = n_dev == nullptr || i < (int64_t)*n_dev;
    int64_t s = live ? seeds[i] : -1;
    uint64_t k = ~0ull;
    if (s >= 0 && (uint64_t)s < npix) {
        uint32_t r = parent[s];
        if (r != CCL_NONE) k = ((uint64_t)r << 32) | (uint64_t)s;
    }
    keys[i] = k;
    vals[i] = seed_labels ? seed_labels[i] : (uint32_t)(i + 1);
}

// Single CTA: split the sorted seed list into components, give single-seed
// components their fill label, multi-seed components their index (MULTI_FLAG | c), a
// size class (XL = heap kernel; L / M / S = bucket-queue kernel with 227 / 54 / 13.5 KB of
// shared memory), slices of the compact node arena and of the edge arena (bucket-queue
// classes) or of the global heap arena (class XL), and a sort key that orders the work
// list class by class, largest first.
// scalars: [0] n_comp [1] n_multi [2,3] XL cursor/end [4,5] L [6,7] M [8,9] S
//          [12] compact nodes [13] edge-arena entries
__global__ void __launch_bounds__(1024)
comp_group_kernel(const uint64_t *__restrict__ keys, const uint32_t *__restrict__ vals, uint32_t n,
                  const uint32_t *__restrict__ comp_size, uint32_t *__restrict__ comp_label,
                  uint32_t *__restrict__ comp_start, uint64_t *__restrict__ arena_off,
                  uint32_t *__restrict__ cbase, uint32_t *__restrict__ ebase,
                  uint32_t *__restrict__ scalars, uint32_t *__restrict__ order_keys,
                  uint32_t *__restrict__ order_vals, uint64_t node_cap, uint64_t edge_cap,
                  uint64_t arena_cap) {
    typedef cub::BlockScan<uint32_t, 1024> Scan32;
    typedef cub::BlockScan<uint64_t, 1024> Scan64;
    __shared__ union {
        typename Scan32::TempStorage s32;
        typename Scan64::TempStorage s64;
    } tmp;
    __shared__ uint32_t carry32, carry_e;
    __shared__ uint64_t carry64;
    __shared__ uint32_t cls[4];            // XL, L, M, S counts
    const uint32_t t = threadIdx.x;
    if (t == 0) { carry32 = 0; carry64 = 0; carry_e = 0; }
    if (t < 4) cls[t] = 0;
    __syncthreads();
    // pass 1: component heads
    for (uint32_t base = 0; base < n; base += 1024) {
        uint32_t i = base + t;
        uint32_t head = 0;
        if (i < n) {
            uint64_t k = keys[i];
            if (k != ~0ull) {
                uint32_t r = (uint32_t)(k >> 32);
                head = (i == 0 || (uint32_t)(keys[i - 1] >> 32) != r) ? 1u : 0u;
            }
        }
        uint32_t pos, total;
        Scan32(tmp.s32).ExclusiveSum(head, pos, total);
        if (head) comp_start[carry32 + pos] = i;
        __syncthreads();
        if (t == 0) carry32 += total;
        __syncthreads();
    }
    const uint32_t n_comp = carry32;
    // number of valid (in-domain) seeds = first index holding the sentinel
    if (t == 0) {
        uint32_t lo = 0, hi = n;
        while (lo < hi) {
            uint32_t mid = (lo + hi) >> 1;
            if (keys[mid] == ~0ull) hi = mid; else lo = mid + 1;
        }
        comp_start[n_comp] = lo;
        scalars[0] = n_comp;
        carry32 = 0;                       // reused: running compact-arena offset
    }
    __syncthreads();
    // pass 2: fill labels, size classes, arena offsets
    for (uint32_t base = 0; base < n_comp; base += 1024) {
        uint32_t c = base + t;
        uint64_t arena = 0;                // global heap-arena entries (class XL only)
        uint32_t nodes = 0;                // compact-arena nodes
        uint32_t edges = 0;                // edge-arena entries (bucket-queue classes)
        uint32_t okey = 0xFFFFFFFFu;
        if (c < n_comp) {
            uint32_t s0 = comp_start[c], s1 = comp_start[c + 1];
            uint32_t root = (uint32_t)(keys[s0] >> 32);
            uint32_t cnt = s1 - s0;
            if (cnt == 1) {
                comp_label[root] = vals[s0];
            } else {
                comp_label[root] = MULTI_FLAG | c;
                nodes = comp_size[root];
                const uint64_t need = (uint64_t)nodes + cnt;       // heap entries it can hold
                const uint64_t P = 3ull * nodes + cnt;             // bucket-queue positions
                const uint32_t bytes = P <= BQ_MAX_POS ? bq_smem_bytes(nodes, cnt) : 0xFFFFFFFFu;
                int k;
                if (bytes > BQ_SMEM_L) {
                    k = 0;
                    arena = need;
                    okey = 0x7FFFFFFFu - (uint32_t)(need > 0x7FFFFFF0ull ? 0x7FFFFFF0ull : need);
                } else {
                    k = bytes > BQ_SMEM_M ? 1 : (bytes > BQ_SMEM_S ? 2 : 3);
                    edges = (uint32_t)P;
                    okey = 0x80000000u + (BQ_SMEM_L - bytes);
                }
                atomicAdd(&cls[k], 1u);
            }
        }
        uint64_t pos64, total64;
        Scan64(tmp.s64).ExclusiveSum(arena, pos64, total64);
        __syncthreads();
        uint32_t pos32, total32;
        Scan32(tmp.s32).ExclusiveSum(nodes, pos32, total32);
        __syncthreads();
        uint32_t pos_e, total_e;
        Scan32(tmp.s32).ExclusiveSum(edges, pos_e, total_e);
        if (c < n_comp) {
            arena_off[c] = carry64 + pos64;
            cbase[c] = carry32 + pos32;
            ebase[c] = carry_e + pos_e;
            order_keys[c] = okey;
            order_vals[c] = c;
        }
        __syncthreads();
        if (t == 0) { carry64 += total64; carry32 += total32; carry_e += total_e; }
        __syncthreads();
    }
    if (t == 0) {
        arena_off[n_comp] = carry64;
        cbase[n_comp] = carry32;
        uint32_t acc = 0;
        for (int k = 0; k < 4; ++k) {          // work-list segments in sort order: XL, L, M, S
            scalars[2 + 2 * k] = acc;
            acc += cls[k];
            scalars[3 + 2 * k] = acc;
        }
        scalars[1] = acc;
        scalars[12] = carry32;
        scalars[13] = carry_e;
        // The compact node / edge / heap arenas may be smaller than the worst case (a caller that
        // knows its mask is sparse passes a smaller workspace): when the multi-seed components do
        // not fit, nothing is flooded -- the work lists are emptied, the compaction kernels see
        // scalars[14] and return, and the host reports ISG_ERR_WORKSPACE after its read-back.
        if ((uint64_t)carry32 > node_cap || (uint64_t)carry_e > edge_cap || carry64 > arena_cap) {
            scalars[14] = 1;
            for (int k = 1; k < 10; ++k) scalars[k] = 0;
            scalars[12] = 0;
        }
    }
    for (uint32_t c = n_comp + t; c <= n; c += 1024) {     // padding up to the host-side count
        if (c < n) {
            order_keys[c] = 0xFFFFFFFFu;
            order_vals[c] = c;
        }
        ebase[c] = carry_e;                                 // empty segments (carry_e is final)
    }
}

// ---------------------------------------------------------------------------
// stage drivers shared by isg_affinity_flood and isg_segment_features
// ---------------------------------------------------------------------------
int ccl_run(const uint8_t *dom, uint32_t *parent, uint32_t *comp_size, uint32_t zp, uint32_t yp,
            uint32_t xp, cudaStream_t st) {
    const uint64_t npix = (uint64_t)zp * yp * xp;
    const int grid = num_sms() * 8;
    ccl_init_kernel<<<grid, 256, 0, st>>>(dom, parent, npix);
    ISG_LAUNCHED();
    ccl_union_kernel<<<grid, 256, 0, st>>>(dom, parent, zp, yp, xp);
    ISG_LAUNCHED();
    ccl_flatten_count_kernel<<<grid, 256, 0, st>>>(parent, comp_size, npix);
    ISG_LAUNCHED();
    return ISG_OK;
}

size_t flood_stage_workspace(FloodStageBuffers *b, Carver &cv, uint64_t npix, int64_t max_seeds,
                             uint64_t node_cap) {
    if (max_seeds < 1) max_seeds = 1;
    if (node_cap > npix) node_cap = npix;
    if (node_cap < 1) node_cap = 1;
    b->node_cap = node_cap;
    b->keys_a = cv.take<uint64_t>(max_seeds);
    b->keys_b = cv.take<uint64_t>(max_seeds);
    b->vals_a = cv.take<uint32_t>(max_seeds);
    b->vals_b = cv.take<uint32_t>(max_seeds);
    b->comp_start = cv.take<uint32_t>(max_seeds + 1);
    b->arena_off = cv.take<uint64_t>(max_seeds + 1);
    b->order_keys_a = cv.take<uint32_t>(max_seeds);
    b->order_keys_b = cv.take<uint32_t>(max_seeds);
    b->order_a = cv.take<uint32_t>(max_seeds);
    b->order_b = cv.take<uint32_t>(max_seeds);
    b->cbase = cv.take<uint32_t>(max_seeds + 1);
    b->ccursor = cv.take<uint32_t>(max_seeds);
    b->max_seeds = max_seeds;
    b->scalars = cv.take<uint32_t>(64);
    size_t cub_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, (uint64_t *)nullptr, (uint64_t *)nullptr,
                                    (uint32_t *)nullptr, (uint32_t *)nullptr, (int)max_seeds);
    b->cub_bytes = cub_bytes + 256;
    b->cub_tmp = cv.take<unsigned char>(b->cub_bytes);
    // heap arena (class XL): every domain voxel enters a heap at most once, plus the seeds
    b->arena_cap = node_cap + (uint64_t)max_seeds;
    b->arena_keys = cv.take<uint64_t>(b->arena_cap);
    b->arena_idx = cv.take<uint32_t>(b->arena_cap);
    // compact component graphs: at most one node per voxel
    b->lidmap = cv.take<uint32_t>(npix);
    b->vox = cv.take<uint32_t>(node_cap);
    b->nbr = cv.take<uint32_t>(node_cap * 6);
    b->key = cv.take<uint32_t>(node_cap * 3);
    b->nlab = cv.take<uint32_t>(node_cap);
    b->rec = cv.take<uint4>(node_cap * 2);
    // edge arena of the bucket-queue classes: 3 stored edges per node + one entry per seed
    b->ebase = cv.take<uint32_t>(max_seeds + 1);
    b->edge_cap = node_cap * 3 + (uint64_t)max_seeds;
    b->ekeys_a = cv.take<uint64_t>(b->edge_cap);
    b->ekeys_b = cv.take<uint64_t>(b->edge_cap);
    b->evals_a = cv.take<uint32_t>(b->edge_cap);
    b->evals_b = cv.take<uint32_t>(b->edge_cap);
    b->seedpos = cv.take<uint32_t>(max_seeds);
    b->seedgs = cv.take<uint32_t>(max_seeds);
    b->complab = cv.take<uint32_t>(max_seeds);
    {
        cub::DoubleBuffer<uint64_t> dk(nullptr, nullptr);
        cub::DoubleBuffer<uint32_t> dv(nullptr, nullptr);
        size_t seg_bytes = 0;
        const int items = (int)(b->edge_cap > 0x7FFFFFF0ull ? 0x7FFFFFF0ull : b->edge_cap);
        cub::DeviceRadixSort::SortPairs(nullptr, seg_bytes, dk, dv, items);
        b->seg_bytes = seg_bytes + 256;
        b->seg_tmp = cv.take<unsigned char>(b->seg_bytes);
    }
    return cv.off;
}

// side streams so that the size classes of the ordered flood run concurrently
static constexpr int FLOOD_SIDE_STREAMS = 3;
struct FloodStreams {
    cudaStream_t s[FLOOD_SIDE_STREAMS];
    cudaEvent_t fork, join[FLOOD_SIDE_STREAMS];
    bool ok;
};
static FloodStreams *flood_streams() {
    static thread_local FloodStreams fs[16];
    static thread_local bool init[16] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 16) dev = 0;
    if (!init[dev]) {
        fs[dev].ok = true;
        for (int i = 0; i < FLOOD_SIDE_STREAMS; ++i) {
            fs[dev].ok &= cudaStreamCreateWithFlags(&fs[dev].s[i], cudaStreamNonBlocking) == cudaSuccess;
            fs[dev].ok &= cudaEventCreateWithFlags(&fs[dev].join[i], cudaEventDisableTiming) == cudaSuccess;
        }
        fs[dev].ok &= cudaEventCreateWithFlags(&fs[dev].fork, cudaEventDisableTiming) == cudaSuccess;
        init[dev] = true;
    }
    return &fs[dev];
}

int flood_stage_run(const FloodStageBuffers &b, const FloodGeom &geom, const uint8_t *mask,
                    const uint32_t *parent, const uint32_t *comp_size, uint32_t *comp_label,
                    const int64_t *seeds, int64_t n_seeds, const uint32_t *n_seeds_dev,
                    uint32_t *labels, cudaStream_t st, const uint32_t *seed_labels) {
    const uint64_t npix = (uint64_t)geom.zp * geom.yp * geom.xp;
    ISG_CUDA(cudaMemsetAsync(b.scalars, 0, 64 * sizeof(uint32_t), st));
    if (n_seeds <= 0) return ISG_OK;
    ISG_REQUIRE(n_seeds <= b.max_seeds, ISG_ERR_WORKSPACE, "flood stage: %lld seeds exceed the workspace (%lld)",
                (long long)n_seeds, (long long)b.max_seeds);
    ISG_CUDA(cudaMemsetAsync(b.ccursor, 0, (size_t)n_seeds * sizeof(uint32_t), st));
    int blocks = (int)((n_seeds + 255) / 256);
    seed_key_kernel<<<blocks, 256, 0, st>>>(seeds, n_seeds, n_seeds_dev, parent, npix, b.keys_a,
                                            b.vals_a, seed_labels);
    ISG_LAUNCHED();
    size_t cub_bytes = b.cub_bytes;
    ISG_CUDA(cub::DeviceRadixSort::SortPairs(b.cub_tmp, cub_bytes, b.keys_a, b.keys_b, b.vals_a,
                                             b.vals_b, (int)n_seeds, 0, 64, st));
    count_launch(4);
    comp_group_kernel<<<1, 1024, 0, st>>>(b.keys_b, b.vals_b, (uint32_t)n_seeds, comp_size,
                                          comp_label, b.comp_start, b.arena_off, b.cbase, b.ebase,
                                          b.scalars, b.order_keys_a, b.order_a, b.node_cap, b.edge_cap,
                                          b.arena_cap);
    ISG_LAUNCHED();
    {
        size_t cb = b.cub_bytes;      // sized for 64-bit keys + 32-bit values: enough for 32/32
        ISG_CUDA(cub::DeviceRadixSort::SortPairs(b.cub_tmp, cb, b.order_keys_a, b.order_keys_b,
                                                 b.order_a, b.order_b, (int)n_seeds, 0, 32, st));
        count_launch(3);
    }
    const int sms = num_sms();
    fill_assign_kernel<<<sms * 8, 256, 0, st>>>(parent, comp_label, mask, labels, b.cbase, b.ccursor,
                                                b.lidmap, b.vox, npix, b.scalars + 14);
    ISG_LAUNCHED();
    seed_edge_kernel<<<blocks, 256, 0, st>>>(b.keys_b, (uint32_t)n_seeds, comp_label, b.comp_start,
                                             b.ebase, b.ekeys_a, b.evals_a, geom.node_key, b.scalars + 14);
    ISG_LAUNCHED();
    compact_graph_kernel<<<sms * 8, 256, 0, st>>>(geom, parent, comp_label, b.comp_start, b.cbase, b.ebase,
                                                  b.lidmap, b.vox, b.scalars + 12, b.nbr, b.key, b.nlab,
                                                  b.rec, b.ekeys_a, b.evals_a);
    ISG_LAUNCHED();
    // rank the edge values of every bucket-queue component: ONE device-wide stable radix sort of
    // (component << 32 | value) -- component c's entries land exactly in its arena slice
    // [ebase[c], ebase[c+1]).  The entry count lives on the device: one small read-back.
    cub::DoubleBuffer<uint64_t> dk(b.ekeys_a, b.ekeys_b);
    cub::DoubleBuffer<uint32_t> dv(b.evals_a, b.evals_b);
    {
        uint32_t tail[2] = {0, 0};                        // scalars[13] = edge entries, [14] = arena overflow
        ISG_CUDA(cudaMemcpyAsync(tail, b.scalars + 13, sizeof(tail), cudaMemcpyDeviceToHost, st));
        ISG_CUDA(cudaStreamSynchronize(st));
        ISG_REQUIRE(tail[1] == 0, ISG_ERR_WORKSPACE,
                    "flood stage: the multi-seed components need more than the %llu compact nodes this "
                    "workspace holds (pass the full isg_*_workspace_bytes size)", (unsigned long long)b.node_cap);
        const uint32_t total_edges = tail[0];
        ISG_REQUIRE((uint64_t)total_edges <= b.edge_cap, ISG_ERR_WORKSPACE, "flood stage: edge arena overflow");
        if (total_edges > 1) {
            int comp_bits = 1;
            while (comp_bits < 32 && (1ll << comp_bits) <= n_seeds) ++comp_bits;
            size_t sb = b.seg_bytes;
            ISG_CUDA(cub::DeviceRadixSort::SortPairs(b.seg_tmp, sb, dk, dv, (int)total_edges, 0, 32 + comp_bits, st));
            count_launch(4);
        }
    }
    {
        const int64_t g = n_seeds < (int64_t)sms * 8 ? n_seeds : (int64_t)sms * 8;
        edge_group_kernel<<<(int)g, 256, 0, st>>>(b.scalars, b.ebase, b.cbase, b.comp_start, dk.Current(),
                                                  dv.Current(), reinterpret_cast<uint16_t *>(b.rec),
                                                  b.seedpos, b.seedgs, geom.node_key != nullptr);
        ISG_LAUNCHED();
    }

    FloodWork w;
    w.seed_keys = b.keys_b;
    w.comp_start = b.comp_start;
    w.arena_off = b.arena_off;
    w.order = b.order_b;
    w.arena_keys = b.arena_keys;
    w.arena_idx = b.arena_idx;
    CompactGraph cg;
    cg.cbase = b.cbase;
    cg.vox = b.vox;
    cg.nbr = b.nbr;
    cg.key = b.key;
    cg.lab = b.nlab;
    cg.lidmap = b.lidmap;
    BqGraph bq;
    bq.cbase = b.cbase;
    bq.ebase = b.ebase;
    bq.vox = b.vox;
    bq.lidmap = b.lidmap;
    bq.rec = b.rec;
    bq.seedpos = b.seedpos;
    bq.seedgs = b.seedgs;
    bq.complab = b.complab;

    const size_t smem_xl = (size_t)FLOOD_SMEM_ENTRIES * (sizeof(uint64_t) + sizeof(uint32_t));
    static bool attr_set_dev[16] = {false};
    int cur_dev = 0;
    cudaGetDevice(&cur_dev);
    bool &attr_set = attr_set_dev[cur_dev >= 0 && cur_dev < 16 ? cur_dev : 0];
    if (!attr_set) {
        ISG_CUDA(cudaFuncSetAttribute(flood_heap_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)smem_xl));
        ISG_CUDA(cudaFuncSetAttribute(flood_bq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)BQ_SMEM_L));
        attr_set = true;
    }
    FloodStreams *fs = flood_streams();
    ISG_REQUIRE(fs->ok, ISG_ERR_CUDA, "flood stage: could not create side streams");
    // upper bounds on the per-class work (the exact counts live on the device)
    const int64_t max_multi = n_seeds / 2 + 1;
    const int flood_sms = post_sms() > 0 ? post_sms() : sms;
    auto grid_for = [&](int64_t per_sm) {
        int64_t gsz = (int64_t)flood_sms * per_sm;
        return (int)(gsz < max_multi ? gsz : max_multi);
    };
    ISG_CUDA(cudaEventRecord(fs->fork, st));
    for (int i = 0; i < FLOOD_SIDE_STREAMS; ++i) ISG_CUDA(cudaStreamWaitEvent(fs->s[i], fs->fork, 0));
    // the biggest classes first: L on the caller's stream, XL / M / S beside it
    w.work_end = b.scalars + 5;
    flood_bq_kernel<<<grid_for(1), 32, BQ_SMEM_L, st>>>(w, bq, b.scalars + 4, labels);
    ISG_LAUNCHED();
    w.work_end = b.scalars + 3;
    flood_heap_kernel<<<grid_for(1), 32, smem_xl, fs->s[0]>>>(w, cg, (uint32_t)FLOOD_SMEM_ENTRIES,
                                                               b.scalars + 2, labels, geom.node_key);
    ISG_LAUNCHED();
    w.work_end = b.scalars + 7;
    flood_bq_kernel<<<grid_for(4), 32, BQ_SMEM_M, fs->s[1]>>>(w, bq, b.scalars + 6, labels);
    ISG_LAUNCHED();
    w.work_end = b.scalars + 9;
    flood_bq_kernel<<<grid_for(16), 32, BQ_SMEM_S, fs->s[2]>>>(w, bq, b.scalars + 8, labels);
    ISG_LAUNCHED();
    for (int i = 0; i < FLOOD_SIDE_STREAMS; ++i) {
        ISG_CUDA(cudaEventRecord(fs->join[i], fs->s[i]));
        ISG_CUDA(cudaStreamWaitEvent(st, fs->join[i], 0));
    }
    return ISG_OK;
}

}  // namespace isg

using namespace isg;

extern "C" size_t isg_flood_workspace_bytes(int64_t zp, int64_t yp, int64_t xp, int64_t max_seeds) {
    const uint64_t npix = (uint64_t)zp * yp * xp;
    Carver cv(nullptr, 0);
    cv.take<uint8_t>(npix);        // domain
    cv.take<uint32_t>(npix);       // parent
    cv.take<uint32_t>(npix);       // comp_size
    cv.take<uint32_t>(npix);       // comp_label
    FloodStageBuffers b;
    flood_stage_workspace(&b, cv, npix, max_seeds, npix);
    return cv.off + 512;
}

extern "C" int isg_affinity_flood(const float *aff, int64_t aff_plane_stride, int aff_origin,
                                  const float *aff_div, const uint8_t *mask, const int64_t *seeds,
                                  int64_t n_seeds, uint32_t *labels, int64_t zp, int64_t yp,
                                  int64_t xp, const float *aff_scale_host, void *workspace,
                                  size_t workspace_bytes, void *stream) {
    ISG_REQUIRE(aff && aff_div && mask && labels, ISG_ERR_ARG, "isg_affinity_flood: null pointer");
    ISG_REQUIRE(zp >= 3 && yp >= 3 && xp >= 3, ISG_ERR_ARG,
                "isg_affinity_flood: padded extents must be >= 3 (got %lld,%lld,%lld)",
                (long long)zp, (long long)yp, (long long)xp);
    ISG_REQUIRE(aff_origin == 0 || aff_origin == 1, ISG_ERR_ARG, "aff_origin must be 0 or 1");
    ISG_REQUIRE(n_seeds >= 0 && (n_seeds == 0 || seeds), ISG_ERR_ARG, "bad seeds");
    const uint64_t npix = (uint64_t)zp * yp * xp;
    ISG_REQUIRE(npix < 0xFFFFFFF0ull, ISG_ERR_OVERFLOW, "volume too large for 32-bit voxel ids");
    cudaStream_t st = (cudaStream_t)stream;
    Carver cv(workspace, workspace_bytes);
    uint8_t *dom = cv.take<uint8_t>(npix);
    uint32_t *parent = cv.take<uint32_t>(npix);
    uint32_t *comp_size = cv.take<uint32_t>(npix);
    uint32_t *comp_label = cv.take<uint32_t>(npix);
    FloodStageBuffers b;
    flood_stage_workspace(&b, cv, npix, n_seeds, npix);
    ISG_REQUIRE(workspace && cv.ok, ISG_ERR_WORKSPACE,
                "isg_affinity_flood: workspace too small (%zu < %zu)", workspace_bytes, cv.off);
    const int sms = num_sms();
    const int grid = sms * 8;
    domain_kernel<<<grid, 256, 0, st>>>(mask, labels, dom, npix);
    ISG_LAUNCHED();
    if (n_seeds > 0) {
        seed_label_kernel<<<(int)((n_seeds + 255) / 256), 256, 0, st>>>(seeds, n_seeds, labels, dom,
                                                                        npix);
        ISG_LAUNCHED();
    }
    ISG_CUDA(cudaMemsetAsync(comp_size, 0, npix * sizeof(uint32_t), st));
    ISG_CUDA(cudaMemsetAsync(comp_label, 0, npix * sizeof(uint32_t), st));
    {
        int rc = ccl_run(dom, parent, comp_size, (uint32_t)zp, (uint32_t)yp, (uint32_t)xp, st);
        if (rc != ISG_OK) return rc;
    }
    FloodGeom g;
    g.aff = aff;
    g.plane_stride = aff_plane_stride;
    g.origin = aff_origin;
    g.za = (uint32_t)(zp - 2 * aff_origin);
    g.ya = (uint32_t)(yp - 2 * aff_origin);
    g.xa = (uint32_t)(xp - 2 * aff_origin);
    g.div = aff_div;
    for (int a = 0; a < 3; ++a) g.scale[a] = aff_scale_host ? fabsf(aff_scale_host[a]) : 1.0f;
    g.zp = (uint32_t)zp;
    g.yp = (uint32_t)yp;
    g.xp = (uint32_t)xp;
    g.node_key = nullptr;
    return flood_stage_run(b, g, mask, parent, comp_size, comp_label, seeds, n_seeds, nullptr, labels,
                           st);
}

#ifdef FLOOD_PROF
extern "C" int isg_debug_flood_prof(unsigned long long *out16, int reset) {
    ISG_CUDA(cudaDeviceSynchronize());
    ISG_CUDA(cudaMemcpyFromSymbol(out16, isg::g_flood_prof, sizeof(unsigned long long) * 16));
    if (reset) {
        unsigned long long z[16] = {0};
        ISG_CUDA(cudaMemcpyToSymbol(isg::g_flood_prof, z, sizeof(z)));
    }
    return ISG_OK;
}
#endif
