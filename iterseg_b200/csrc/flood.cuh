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
//   * only components with >= 2 seeds run the ordered flood, each on its own compacted
//     graph: flood_bq_kernel (an exact O(1) bucket queue over the component's ranked edge
//     values, one thread per component, queue in shared memory) and, for components too
//     big for that, flood_heap_kernel (one warp per component, 32-ary heap).
// Seeds carry value 0.0 / age 0 in the reference and are ordered by flat index
// (third tuple field, watershed.py:162); both kernels queue them first, in index order,
// among the entries of value 0.0, which preserves every comparison.
#pragma once
#include "ccl.cuh"
#include "flood_stage.h"

namespace isg {

static constexpr int FLOOD_SMEM_ENTRIES = 18816;     // class XL: heap slots [0, 18816) (12 B each, 220.5 KB)
                                                     // live in shared memory, the rest in the global arena
static constexpr uint32_t MULTI_FLAG = 0x80000000u;  // comp_label[root] = MULTI_FLAG | component index
static constexpr uint32_t NO_NODE = 0xFFFFu;
static constexpr uint32_t NO_NODE32 = 0xFFFFFFFFu;

struct FloodWork {
    const uint64_t *seed_keys;     // sorted (root << 32 | padded flat index)
    const uint32_t *comp_start;    // n_comp + 1 offsets into the sorted arrays
    const uint64_t *arena_off;     // per component offset into the heap arena (class XL)
    const uint32_t *work_end;      // device scalar: number of work-list entries for this kernel
    const uint32_t *order;         // component indices, largest first
    uint64_t *arena_keys;
    uint32_t *arena_idx;
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

// Every multi-seed component is first compacted (fill_assign_kernel, compact_graph_kernel):
// its voxels get dense local ids (their order is irrelevant: no comparison ever reaches the
// id bits), and per node the six neighbour ids (NO_NODE = not claimable) and the edge keys
// (order-preserving bits of aff / channel_max).

// ---------------------------------------------------------------------------
// flood_heap_kernel (class XL: components too big for the bucket queue below)
// ---------------------------------------------------------------------------
// One warp per component, a 32-ary min-heap of (key << 32 | age, node) pairs: the first
// `cap` slots in shared memory, the rest in the global arena; pop = 2x redux.sync.min per
// level, the six neighbour tests done by six lanes, the graph in global memory (L2) with
// the loads of a popped node issued before / during the sift-down.
struct CompactGraph {
    const uint32_t *cbase;         // [n_comp + 1] first node of component c (compact arena)
    const uint32_t *vox;           // [total] padded flat voxel index of every node
    const uint32_t *nbr;           // [total][6]  (NO_NODE32 = not claimable)
    const uint32_t *key;           // [total][3]
    uint32_t *lab;                 // [total] node labels (zeroed by compact_graph_kernel)
    const uint32_t *lidmap;        // [npix] voxel -> local id
};

__global__ void __launch_bounds__(32)
flood_heap_kernel(FloodWork w, CompactGraph cg, uint32_t cap, uint32_t *cursor, uint32_t *labels,
                  const uint32_t *node_key /* nullable: node-keyed mode */) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const unsigned FULL = 0xFFFFFFFFu;
    const unsigned lane = threadIdx.x;
    const uint32_t work_end = *w.work_end;
    uint64_t *const heap = reinterpret_cast<uint64_t *>(smem_raw);
    uint32_t *const hnode = reinterpret_cast<uint32_t *>(smem_raw + (size_t)cap * 8);
    const int axis = lane < 3 ? (int)lane : (lane < 6 ? 5 - (int)lane : 0);   // [0,1,2,2,1,0]

    for (;;) {
        uint32_t wi = 0;
        if (lane == 0) wi = atomicAdd(cursor, 1u);
        wi = __shfl_sync(FULL, wi, 0);
        if (wi >= work_end) break;
        const uint32_t c = w.order[wi];
        const uint32_t s0 = w.comp_start[c], cnt = w.comp_start[c + 1] - s0;
        const uint32_t b0 = cg.cbase[c], nv = cg.cbase[c + 1] - b0;
        const uint32_t *const gkey = cg.key + (size_t)b0 * 3;
        const uint32_t *const gnbr = cg.nbr + (size_t)b0 * 6;
        uint32_t *const glab = cg.lab + b0;
        uint64_t *const akeys = w.arena_keys + w.arena_off[c];       // heap overflow
        uint32_t *const anode = w.arena_idx + w.arena_off[c];
        __syncwarp();                                   // previous component's smem reads are done
        auto hk_ld = [&](uint32_t i) -> uint64_t { return i < cap ? heap[i] : akeys[i]; };
        auto hk_st = [&](uint32_t i, uint64_t k) { if (i < cap) heap[i] = k; else akeys[i] = k; };
        auto hn_ld = [&](uint32_t i) -> uint32_t { return i < cap ? hnode[i] : anode[i]; };
        auto hn_st = [&](uint32_t i, uint32_t x) { if (i < cap) hnode[i] = x; else anode[i] = x; };

        const uint64_t zero_hi = (uint64_t)f32_ord(0.0f) << 32;
        for (uint32_t i = lane; i < cnt; i += 32) {
            const uint32_t v = (uint32_t)(w.seed_keys[s0 + i] & 0xFFFFFFFFu);
            const uint32_t lid = cg.lidmap[v];
            // ascending array == valid heap (node-keyed mode: fixed up below)
            hk_st(i, (node_key ? (uint64_t)__ldg(node_key + v) << 32 : zero_hi) | i);
            hn_st(i, lid);
            __stcg(glab + lid, labels[v]);   // the final seed label (duplicates: largest, set upstream)
        }
        uint32_t n = cnt;
        uint32_t age = cnt;
        uint32_t pref_node = NO_NODE32, pref_nb = NO_NODE32;   // speculative adjacency prefetch
        __syncwarp();
        if (node_key) {
            // the seeds carry their own keys: establish the heap order (32-ary Floyd build)
            for (int64_t root = ((int64_t)n - 2) / 32; root >= 0; --root) {
                uint32_t i = (uint32_t)root;
                const uint64_t lk = hk_ld(i);
                const uint32_t ln = hn_ld(i);
                for (;;) {
                    const uint32_t c0 = i * 32u + 1u;
                    if (c0 >= n) break;
                    const uint32_t ch = c0 + lane;
                    const uint64_t k = ch < n ? hk_ld(ch) : ~0ull;
                    const uint32_t hi = (uint32_t)(k >> 32);
                    const uint32_t mhi = __reduce_min_sync(FULL, hi);
                    const uint32_t lo = hi == mhi ? (uint32_t)k : 0xFFFFFFFFu;
                    const uint32_t mlo = __reduce_min_sync(FULL, lo);
                    const uint64_t mk = ((uint64_t)mhi << 32) | mlo;
                    if (mk >= lk) break;
                    const uint32_t wc = c0 + __ffs(__ballot_sync(FULL, hi == mhi && lo == mlo)) - 1;
                    if (lane == 0) {
                        hk_st(i, mk);
                        hn_st(i, hn_ld(wc));
                    }
                    __syncwarp();
                    i = wc;
                }
                if (lane == 0) {
                    hk_st(i, lk);
                    hn_st(i, ln);
                }
                __syncwarp();
            }
        }

        while (n > 0) {
            // ---- pop the minimum -------------------------------------------------
            const uint32_t p = hnode[0];
            --n;
            const uint64_t lastk = hk_ld(n);
            const uint32_t lastn = hn_ld(n);
            // expansion loads, part 1 (in flight during the sift-down): the adjacency of the
            // node that was on top after the previous sift-down was requested back then
            uint32_t nb = NO_NODE32;
            const bool pref_hit = pref_node == p;
            if (pref_hit) nb = pref_nb;
            else if (lane < 6) nb = __ldg(gnbr + (size_t)p * 6u + lane);
            const uint32_t labp = __ldcg(glab + p);
            uint32_t labn = 1, kord = 0;
            bool fetched = false;
            auto fetch2 = [&]() {        // expansion loads, part 2: need nb
                if (nb != NO_NODE32) {
                    // edge key: stored at the popped voxel for the three negative directions,
                    // at the neighbour for the positive ones
                    const size_t kn = node_key ? (size_t)nb * 3u : (size_t)(lane < 3 ? p : nb) * 3u + axis;
                    labn = __ldcg(glab + nb);
                    kord = __ldg(gkey + kn);
                }
                fetched = true;
            };
            if (pref_hit) fetch2();
            __syncwarp();                                // all lanes hold p / last before slot 0 changes
            if (n > 0) {
                uint32_t i = 0;
                for (;;) {
                    const uint32_t c0 = i * 32u + 1u;
                    if (c0 >= n) break;
                    const uint32_t ch = c0 + lane;
                    const uint64_t k = ch < n ? hk_ld(ch) : ~0ull;
                    const uint32_t hi = (uint32_t)(k >> 32);
                    const uint32_t mhi = __reduce_min_sync(FULL, hi);
                    const uint32_t lo = hi == mhi ? (uint32_t)k : 0xFFFFFFFFu;
                    const uint32_t mlo = __reduce_min_sync(FULL, lo);
                    const uint64_t mk = ((uint64_t)mhi << 32) | mlo;
                    if (mk >= lastk) break;
                    const uint32_t win = __ffs(__ballot_sync(FULL, hi == mhi && lo == mlo)) - 1;
                    const uint32_t wc = c0 + win;
                    if (lane == 0) {
                        hk_st(i, mk);
                        hn_st(i, hn_ld(wc));
                    }
                    __syncwarp();
                    i = wc;
                    if (!fetched) fetch2();
                }
                if (lane == 0) {
                    hk_st(i, lastk);
                    hn_st(i, lastn);
                }
                __syncwarp();
                pref_node = hnode[0];                    // request the adjacency of the likely next pop
                pref_nb = NO_NODE32;
                if (lane < 6) pref_nb = __ldg(gnbr + (size_t)pref_node * 6u + lane);
            }
            if (!fetched) fetch2();
            // ---- expand the popped node (watershed.py:135-154) ---------------------
            const bool claim = nb != NO_NODE32 && labn == 0;
            unsigned bits = __ballot_sync(FULL, claim);
            if (claim) __stcg(glab + nb, labp);           // labelled at push time (:149)
            const uint32_t myage = age + __popc(bits & ((1u << lane) - 1u));
            const uint64_t mykey = ((uint64_t)kord << 32) | myage;
            age += __popc(bits);
            __syncwarp();
            while (bits) {
                const int src = __ffs(bits) - 1;
                bits &= bits - 1;
                const uint64_t k = __shfl_sync(FULL, mykey, src);
                const uint32_t kn = __shfl_sync(FULL, nb, src);
                uint32_t i = n++;
                while (i > 0) {
                    const uint32_t par = (i - 1u) >> 5;
                    const uint64_t pk = hk_ld(par);
                    if (pk <= k) break;
                    if (lane == 0) {
                        hk_st(i, pk);
                        hn_st(i, hn_ld(par));
                    }
                    i = par;
                }
                if (lane == 0) {
                    hk_st(i, k);
                    hn_st(i, kn);
                }
                __syncwarp();
            }
        }
        // ---- write the component's labels back ---------------------------------------
        __syncwarp();
        for (uint32_t i = lane; i < nv; i += 32) labels[cg.vox[b0 + i]] = __ldcg(glab + i);
    }
}

// ---------------------------------------------------------------------------
// flood_bq_kernel: the fast ordered flood -- an exact bucket queue
// ---------------------------------------------------------------------------
// The reference heap orders entries by (value, age) (watershed.py:162).  Inside one
// component the values that can ever be queued are known before the flood starts: a voxel
// is queued with the affinity of the edge it was claimed through, so the universe is the
// component's edge list (+ one value-0.0 pseudo edge per seed).  The edges are ranked once,
// in parallel (segmented radix sort of the order-preserving float bits, one segment per
// component; edge_group_kernel turns ranks into "group start" positions, a group being a
// run of equal values).  [The ranking is one device-wide radix sort of (component << 32 | value).]  The queue is then a set of POSITIONS:
//   push(edge) : position = group start + tail[group]++ ;  slots[position] = node
//   pop        : the lowest queued position
// Equal values pop in insertion order because a group's positions are handed out in
// ascending order: exactly the reference's age tie-break, with no age counter at all.  An
// edge is used for at most one push (after it both end points are labelled), so a group
// never overflows.  Seeds are the first entries of the 0.0 group in flat-index order (the
// reference's (0.0, 0, index) tuples).
//
// One warp per component.  The set of queued positions is split in two:
//   * the FRONT: at most 32 entries (position << 16 | node), one per lane, in registers;
//     every front entry is below `bmin`;
//   * the BACK: a two-level bitmap over positions in shared memory; every back entry is at
//     or above `bmin`.
// pop = redux.min over the front (one instruction); a push below bmin goes to a free front
// lane, the others set a bit in the back.  When the front runs empty it is refilled from
// the first 32 non-empty-or-not bitmap words at once (one word per lane) and bmin moves up.
// The six neighbour tests of a popped node are done by six lanes.  The read-only node
// record {6 neighbours, 6 group starts} (32 B) stays in global memory (L2); the lane that
// receives a front entry requests it right away with cp.async into its own staging slot, so
// that the fetch overlaps the pops in front of it (a cp.async holds no register scoreboard,
// a plain load would serialise the whole warp behind it).
// Queue + labels take ~14.1 B of shared memory per node -> components of up to ~16 000
// voxels; bigger ones take the heap kernel above (class XL).
static constexpr uint32_t BQ_SMEM_S = 13824;         // 16 components per SM
static constexpr uint32_t BQ_SMEM_M = 55296;         // 4 per SM
static constexpr uint32_t BQ_SMEM_L = 232448;        // 1 per SM (227 KB)
static constexpr uint32_t BQ_SEED_FLAG = 0x80000000u;
static constexpr uint32_t BQ_MAX_POS = 65535u;       // positions and node ids are 16-bit
static constexpr uint32_t BQ_HEADER = 1024u + 128u;  // 32 record staging slots + refill scratch
#ifndef BQ_SWEEP
#define BQ_SWEEP true                                // the dead-entry sweep of flood_bq_kernel (-DBQ_SWEEP=false: off)
#endif
#ifndef BQ_SWEEP_EVERY
#define BQ_SWEEP_EVERY 4u
#endif

__host__ __device__ __forceinline__ uint32_t bq_smem_bytes(uint32_t nodes, uint32_t seeds) {
    const uint32_t P = 3u * nodes + seeds;
    const uint32_t n0 = (P + 31u) / 32u, n1 = (n0 + 31u) / 32u;
    return BQ_HEADER + ((4u * P + 2u * nodes + 3u) & ~3u) + 4u * n0 + 4u * n1;
}

#ifdef FLOOD_PROF
// debug: [0] pops [1] refills [2] front pushes [3] back pushes [4] evictions [6] total clocks
__device__ unsigned long long g_flood_prof[16];
#define FP_ADD(i, v) prof[i] += (unsigned long long)(v)
#else
#define FP_ADD(i, v)
#endif

struct BqGraph {
    const uint32_t *cbase;         // [n_comp + 1] first node of component c (compact arena)
    const uint32_t *ebase;         // [n_comp + 1] first queue position (edge arena); empty = class XL
    const uint32_t *vox;           // [total] padded flat voxel index of every node
    const uint32_t *lidmap;        // [npix] voxel -> local id
    const uint4 *rec;              // [total][2] u16 x {6 neighbours, 6 group starts, 4 pad}
    const uint32_t *seedpos;       // [n_seeds] queue position of the i-th sorted seed
    const uint32_t *seedgs;        // [n_seeds] start of the value group that position belongs to
    uint32_t *complab;             // [n_seeds] final label of the i-th sorted seed
};

__global__ void __launch_bounds__(32)
flood_bq_kernel(FloodWork w, BqGraph g, uint32_t *cursor, uint32_t *labels) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const unsigned FULL = 0xFFFFFFFFu;
    const unsigned lane = threadIdx.x;
    const uint32_t work_end = *w.work_end;
    const uint32_t EMPTY = 0xFFFFFFFFu;
    const uint16_t *const stage16 = reinterpret_cast<const uint16_t *>(smem_raw);
    uint32_t *const scratch = reinterpret_cast<uint32_t *>(smem_raw + 1024);
    const uint32_t my_stage = (uint32_t)__cvta_generic_to_shared(smem_raw) + lane * 32u;
    for (;;) {
        uint32_t wi = 0;
        if (lane == 0) wi = atomicAdd(cursor, 1u);
        wi = __shfl_sync(FULL, wi, 0);
        if (wi >= work_end) break;
        const uint32_t c = w.order[wi];
        const uint32_t s0 = w.comp_start[c], cnt = w.comp_start[c + 1] - s0;
        const uint32_t b0 = g.cbase[c], nv = g.cbase[c + 1] - b0;
        const uint32_t P = 3u * nv + cnt;
        const uint32_t n0 = (P + 31u) / 32u, n1 = (n0 + 31u) / 32u;
        uint16_t *const slots = reinterpret_cast<uint16_t *>(smem_raw + BQ_HEADER);
        uint16_t *const tail = slots + P;
        uint16_t *const slab = tail + P;
        uint32_t *const L0 = reinterpret_cast<uint32_t *>(smem_raw + BQ_HEADER + ((4u * P + 2u * nv + 3u) & ~3u));
        uint32_t *const L1 = L0 + n0;
        const uint4 *const rec = g.rec + (size_t)b0 * 2;
        __syncwarp();                                   // previous component's smem reads are done
        for (uint32_t i = lane; i < P; i += 32) tail[i] = 0;
        for (uint32_t i = lane; i < nv; i += 32) slab[i] = 0;
        for (uint32_t i = lane; i < n0 + n1; i += 32) L0[i] = 0;
        __syncwarp();
        for (uint32_t i = lane; i < cnt; i += 32) {
            const uint32_t v = (uint32_t)(w.seed_keys[s0 + i] & 0xFFFFFFFFu);
            const uint32_t lid = g.lidmap[v];
            const uint32_t pos = g.seedpos[s0 + i];
            g.complab[s0 + i] = labels[v];   // the final seed label (duplicates: largest, set upstream)
            slots[pos] = (uint16_t)lid;
            slab[lid] = (uint16_t)(i + 1);   // duplicated seeds carry the same final label
            atomicOr(L0 + (pos >> 5), 1u << (pos & 31u));
            atomicOr(L1 + (pos >> 10), 1u << ((pos >> 5) & 31u));
        }
        if (lane == 0)                                   // the seeds open their value groups (affinity
            for (uint32_t i = 0; i < cnt; ++i)             // mode: all of them the 0.0 group)
                tail[g.seedgs[s0 + i]] += 1;
        __syncwarp();

        // asynchronous copy of a node record into this lane's staging slot
        auto request = [&](uint32_t node) {
            const uint4 *src = rec + 2u * node;
            asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(my_stage), "l"(src) : "memory");
            asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(my_stage + 16u), "l"(src + 1) : "memory");
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        auto back_insert = [&](uint32_t qq) {
            atomicOr(L0 + (qq >> 5), 1u << (qq & 31u));
            atomicOr(L1 + (qq >> 10), 1u << ((qq >> 5) & 31u));
        };
        uint32_t F = EMPTY;          // this lane's front entry
        uint32_t bmin = 0;           // every back entry is >= bmin, every front entry < bmin
        uint32_t age = 0;            // pops since this lane requested its record (the sweep waits for age >= 3)
        uint32_t sweep_it = 0;
        const bool sweep_on = BQ_SWEEP;
#ifdef FLOOD_PROF
        unsigned long long prof[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        const long long tc0 = clock64();
#endif
        for (;;) {
            // ---- dead-entry sweep -------------------------------------------------------------
            // About half of all pops claim nothing: every neighbour of the popped node has been
            // labelled in the meantime.  Labels are only ever added, so an entry whose node has no
            // unlabelled neighbour left stays a no-op for ever -- and a no-op pop touches neither
            // labels nor tail counters nor positions.  Such entries are dropped HERE, by all 32 lanes at
            // once (each lane tests the node of its own front entry against the record it has
            // staged), instead of one by one on the ordered critical path.  The result is identical.
            if (sweep_on) {
                ++age;
                // every BQ_SWEEP_EVERY-th iteration: the sweep costs ~130 clk, a dead pop ~300
                if ((++sweep_it % BQ_SWEEP_EVERY) == 0u && F != EMPTY && age >= 3u) {
                    asm volatile("cp.async.wait_group 0;" ::: "memory");
                    const uint4 r = *reinterpret_cast<const uint4 *>(smem_raw + lane * 32u);
                    // branch-free: a missing neighbour reads the entry's own node, which is labelled
                    const uint32_t self = F & 0xFFFFu;
                    uint32_t n0 = r.x & 0xFFFFu, n1 = r.x >> 16, n2 = r.y & 0xFFFFu, n3 = r.y >> 16,
                             n4 = r.z & 0xFFFFu, n5 = r.z >> 16;
                    n0 = n0 == NO_NODE ? self : n0;
                    n1 = n1 == NO_NODE ? self : n1;
                    n2 = n2 == NO_NODE ? self : n2;
                    n3 = n3 == NO_NODE ? self : n3;
                    n4 = n4 == NO_NODE ? self : n4;
                    n5 = n5 == NO_NODE ? self : n5;
                    const uint32_t l0 = slab[n0], l1 = slab[n1], l2 = slab[n2], l3 = slab[n3], l4 = slab[n4],
                                   l5 = slab[n5];
                    const bool dead = (l0 != 0) & (l1 != 0) & (l2 != 0) & (l3 != 0) & (l4 != 0) & (l5 != 0);
                    F = dead ? EMPTY : F;
                }
            }
            uint32_t mn = __reduce_min_sync(FULL, F);
            if (mn == EMPTY) {
                // ---- refill the front from the back: the first non-empty word and the 31 after it ----
                FP_ADD(1, 1);
                const uint32_t a = lane < n1 ? L1[lane] : 0u;
                const uint32_t b = lane + 32u < n1 ? L1[lane + 32u] : 0u;
                const unsigned ba = __ballot_sync(FULL, a != 0), bb = __ballot_sync(FULL, b != 0);
                if ((ba | bb) == 0) break;                              // queue empty: component done
                const uint32_t i2 = ba ? (uint32_t)__ffs((int)ba) - 1u : 32u + (uint32_t)__ffs((int)bb) - 1u;
                const uint32_t w1 = __shfl_sync(FULL, ba ? a : b, i2 & 31u);
                const uint32_t i1 = i2 * 32u + (uint32_t)__ffs((int)w1) - 1u;
                const uint32_t wj = i1 + lane < n0 ? L0[i1 + lane] : 0u;
                const uint32_t cj = (uint32_t)__popc(wj);
                uint32_t sj = cj;                                        // inclusive prefix of the bit counts
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t t = __shfl_up_sync(FULL, sj, o);
                    if (lane >= (unsigned)o) sj += t;
                }
                const bool take = sj <= 32u;                             // whole words only; lane 0 always fits
                const unsigned tb = __ballot_sync(FULL, take);
                const uint32_t k = (uint32_t)__ffs((int)~tb) - 1u;       // taken words = lanes [0, k)  (k = 32: ffs(0) - 1)
                const uint32_t kk = tb == FULL ? 32u : k;
                const uint32_t total = __shfl_sync(FULL, sj, kk - 1u);
                if (lane < kk && wj) {
                    uint32_t e = sj - cj, x = wj;
                    while (x) {
                        scratch[e++] = (i1 + lane) * 32u + (uint32_t)__ffs((int)x) - 1u;
                        x &= x - 1u;
                    }
                    L0[i1 + lane] = 0u;
                    atomicAnd(L1 + ((i1 + lane) >> 5), ~(1u << ((i1 + lane) & 31u)));
                }
                __syncwarp();
                if (lane < total) {
                    const uint32_t pos = scratch[lane];
                    const uint32_t node = slots[pos];
                    F = (pos << 16) | node;
                    request(node);
                    age = 0;
                }
                bmin = (i1 + kk) * 32u;
                __syncwarp();
                continue;
            }
            // ---- pop the lowest front entry -------------------------------------------
            FP_ADD(0, 1);
            const uint32_t o = (uint32_t)__ffs((int)__ballot_sync(FULL, F == mn)) - 1u;
            if (lane == o) {
                F = EMPTY;
                asm volatile("cp.async.wait_group 0;" ::: "memory");
            }
            __syncwarp();
            const uint32_t p = mn & 0xFFFFu;
            // ---- expand (watershed.py:135-154): lanes 0..5 = neighbours [-z,-y,-x,+x,+y,+z] ----
            const uint32_t labp = slab[p];
            uint32_t nb = NO_NODE, gp = 0;
            if (lane < 6) {
                nb = stage16[o * 16u + lane];
                gp = stage16[o * 16u + 6u + lane];
            }
            const bool claim = nb != NO_NODE && slab[nb] == 0;
            const unsigned cb = __ballot_sync(FULL, claim);
            if (cb == 0) continue;
            uint32_t qq = 0;
            if (claim) slab[nb] = (uint16_t)labp;                         // labelled at push time (:149)
            if ((cb & (cb - 1u)) == 0) {                                   // one claim (the common case)
                if (claim) {
                    const uint32_t t = tail[gp];
                    tail[gp] = (uint16_t)(t + 1u);
                    qq = gp + t;
                }
            } else {                                                       // in direction order: equal values
                for (unsigned rest = cb; rest; rest &= rest - 1u) {        // share a group
                    if (lane == (unsigned)__ffs((int)rest) - 1u) {
                        const uint32_t t = tail[gp];
                        tail[gp] = (uint16_t)(t + 1u);
                        qq = gp + t;
                    }
                    __syncwarp();
                }
            }
            if (claim) slots[qq] = (uint16_t)nb;
            const bool to_back = claim && qq >= bmin;
            if (to_back) { back_insert(qq); FP_ADD(3, 1); }
            unsigned fb = __ballot_sync(FULL, claim && qq < bmin);
            const uint32_t x = (qq << 16) | nb;
            while (fb) {
                const int src = __ffs((int)fb) - 1;
                fb &= fb - 1u;
                const uint32_t xs = __shfl_sync(FULL, x, src);
                const unsigned free_lanes = __ballot_sync(FULL, F == EMPTY);
                if (free_lanes) {
                    FP_ADD(2, 1);
                    if (lane == (unsigned)__ffs((int)free_lanes) - 1u) {
                        F = xs;
                        request(xs & 0xFFFFu);
                        age = 0;
                    }
                } else {
                    // front full: the larger of (new entry, front maximum) moves to the back
                    FP_ADD(4, 1);
                    const uint32_t mx = __reduce_max_sync(FULL, F);
                    const uint32_t y = xs > mx ? xs : mx;
                    if (mx > xs && F == mx) {
                        F = xs;
                        request(xs & 0xFFFFu);
                        age = 0;
                    }
                    if (lane == 0) back_insert(y >> 16);
                    bmin = y >> 16;
                }
            }
            __syncwarp();
        }
#ifdef FLOOD_PROF
        if (lane == 0 && nv > 5000) {
            prof[6] = (unsigned long long)(clock64() - tc0);
            for (int qi = 0; qi < 8; ++qi) atomicAdd(&g_flood_prof[qi], prof[qi]);
        }
#endif
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncwarp();
        // ---- write the component's labels back ---------------------------------------
        for (uint32_t i = lane; i < nv; i += 32) {
            const uint32_t l = slab[i];
            if (l) labels[g.vox[b0 + i]] = g.complab[s0 + l - 1u];
        }
    }
}

static constexpr int EG_ITEMS = 8;
struct MaxU32 {
    __device__ __forceinline__ uint32_t operator()(uint32_t a, uint32_t b) const { return a > b ? a : b; }
};

// One CTA per component: group starts of the sorted edge list, scattered into the node
// records of both end points; queue positions of the seeds.
__global__ void __launch_bounds__(256)
edge_group_kernel(const uint32_t *__restrict__ n_comp_dev, const uint32_t *__restrict__ ebase,
                  const uint32_t *__restrict__ cbase, const uint32_t *__restrict__ comp_start,
                  const uint64_t *__restrict__ skeys, const uint32_t *__restrict__ svals,
                  uint16_t *__restrict__ rec16, uint32_t *__restrict__ seedpos,
                  uint32_t *__restrict__ seedgs, int node_mode) {
    typedef cub::BlockScan<uint32_t, 256> Scan;
    __shared__ typename Scan::TempStorage tmp;
    __shared__ uint32_t carry;
    const uint32_t t = threadIdx.x;
    for (uint32_t c = blockIdx.x; c < *n_comp_dev; c += gridDim.x) {
        const uint32_t e0 = ebase[c], P = ebase[c + 1] - e0;
        if (P == 0) continue;
        const uint32_t b0 = cbase[c], s0 = comp_start[c];
        __syncthreads();
        if (t == 0) carry = 0;
        __syncthreads();
        for (uint32_t base = 0; base < P; base += 256 * EG_ITEMS) {
            // blocked arrangement: thread t owns EG_ITEMS consecutive entries
            const uint32_t i0 = base + t * EG_ITEMS;
            uint32_t gs[EG_ITEMS];
            uint64_t prev = (i0 > 0 && i0 - 1 < P) ? skeys[e0 + i0 - 1] : 0ull;
#pragma unroll
            for (int u = 0; u < EG_ITEMS; ++u) {
                const uint32_t i = i0 + u;
                const uint64_t k = i < P ? skeys[e0 + i] : 0ull;
                gs[u] = (i > 0 && i < P && k != prev) ? i : 0u;
                prev = k;
            }
            Scan(tmp).InclusiveScan(gs, gs, MaxU32());
            const uint32_t cr = carry;
            __syncthreads();
            if (t == 255) carry = gs[EG_ITEMS - 1] > cr ? gs[EG_ITEMS - 1] : cr;
            __syncthreads();
#pragma unroll
            for (int u = 0; u < EG_ITEMS; ++u) {
                const uint32_t i = i0 + u;
                if (i >= P) break;
                const uint32_t start = gs[u] > cr ? gs[u] : cr;
                const uint32_t id = svals[e0 + i];
                if (id & BQ_SEED_FLAG) {
                    seedpos[s0 + (id & ~BQ_SEED_FLAG)] = i;
                    seedgs[s0 + (id & ~BQ_SEED_FLAG)] = start;
                } else if (node_mode) {
                    // entry of node j (axis 0 only): every neighbour that can claim j queues it here
                    const uint32_t j = id / 3u, axis = id - 3u * j;
                    if (axis == 0) {
#pragma unroll
                        for (uint32_t d = 0; d < 6; ++d) {
                            const uint32_t nb = rec16[(size_t)(b0 + j) * 16u + d];
                            if (nb != NO_NODE) rec16[(size_t)(b0 + nb) * 16u + 6u + (5u - d)] = (uint16_t)start;
                        }
                    }
                } else {
                    const uint32_t j = id / 3u, axis = id - 3u * j;
                    const uint32_t nb = rec16[(size_t)(b0 + j) * 16u + axis];
                    if (nb != NO_NODE) {      // edge (j, j - e_axis): same value from either side
                        rec16[(size_t)(b0 + j) * 16u + 6u + axis] = (uint16_t)start;
                        rec16[(size_t)(b0 + nb) * 16u + 6u + (5u - axis)] = (uint16_t)start;
                    }
                }
            }
        }
    }
}

// value-0.0 pseudo edges of the seeds, placed in front of the component's edge segment (the
// radix sort is stable, so they open the 0.0 group in flat-index order)
__global__ void seed_edge_kernel(const uint64_t *__restrict__ seed_keys, uint32_t n,
                                 const uint32_t *__restrict__ comp_label,
                                 const uint32_t *__restrict__ comp_start,
                                 const uint32_t *__restrict__ ebase, uint64_t *__restrict__ ekeys,
                                 uint32_t *__restrict__ evals, const uint32_t *__restrict__ node_key,
                                 const uint32_t *__restrict__ overflow) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || *overflow) return;
    const uint64_t k = seed_keys[i];
    if (k == ~0ull) return;
    const uint32_t cl = comp_label[(uint32_t)(k >> 32)];
    if (!(cl & MULTI_FLAG)) return;
    const uint32_t c = cl & ~MULTI_FLAG;
    const uint32_t e0 = ebase[c];
    if (ebase[c + 1] == e0) return;
    const uint32_t idx = i - comp_start[c];
    ekeys[e0 + idx] = ((uint64_t)c << 32) |
                      (node_key ? __ldg(node_key + (uint32_t)(k & 0xFFFFFFFFu)) : f32_ord(0.0f));
    evals[e0 + idx] = BQ_SEED_FLAG | idx;
}

// Pass 1 of the compaction, fused with the single-seed fill: every voxel of a
// single-seed component takes the seed's label; every voxel of a compacted multi-seed
// component gets a local id (warp-aggregated atomics per component).
__global__ void __launch_bounds__(256)
fill_assign_kernel(const uint32_t *__restrict__ parent, const uint32_t *__restrict__ comp_label,
                   const uint8_t *__restrict__ mask, uint32_t *__restrict__ labels,
                   const uint32_t *__restrict__ cbase, uint32_t *__restrict__ ccursor,
                   uint32_t *__restrict__ lidmap, uint32_t *__restrict__ vox, uint64_t n,
                   const uint32_t *__restrict__ overflow) {
    const bool no_compaction = *overflow != 0;          // arenas too small: single-seed fill only
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const unsigned lane = threadIdx.x & 31;
    const uint64_t i0 = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (uint64_t base = i0 - lane; base < n; base += stride) {
        const uint64_t v = base + lane;
        uint32_t c = 0xFFFFFFFFu;
        if (v < n) {
            const uint32_t r = parent[v];
            if (r != CCL_NONE) {
                const uint32_t cl = comp_label[r];
                if (cl & MULTI_FLAG) {
                    if (!no_compaction) c = cl & ~MULTI_FLAG;
                } else if (cl != 0 && mask[v] && labels[v] == 0) {
                    labels[v] = cl;
                }
            }
        }
        const unsigned active = __ballot_sync(0xFFFFFFFFu, c != 0xFFFFFFFFu);
        if (c != 0xFFFFFFFFu) {
            const unsigned peers = __match_any_sync(active, c);
            const unsigned leader = __ffs(peers) - 1;
            uint32_t first = 0;
            if (lane == leader) first = atomicAdd(ccursor + c, (uint32_t)__popc(peers));
            first = __shfl_sync(peers, first, leader);
            const uint32_t lid = first + __popc(peers & ((1u << lane) - 1u));
            lidmap[v] = lid;
            vox[cbase[c] + lid] = (uint32_t)v;
        }
    }
}

// Pass 2: adjacency and edge keys of every compacted node.  Components of the bucket-queue
// classes get a 32-byte record per node and one (key, id) entry per stored edge in their
// segment of the edge arena; class XL keeps the wide arrays of the heap kernel.
__global__ void __launch_bounds__(256)
compact_graph_kernel(FloodGeom g, const uint32_t *__restrict__ parent,
                     const uint32_t *__restrict__ comp_label, const uint32_t *__restrict__ comp_start,
                     const uint32_t *__restrict__ cbase, const uint32_t *__restrict__ ebase,
                     const uint32_t *__restrict__ lidmap, const uint32_t *__restrict__ vox,
                     const uint32_t *__restrict__ total_dev, uint32_t *__restrict__ nbr,
                     uint32_t *__restrict__ key, uint32_t *__restrict__ lab, uint4 *__restrict__ rec,
                     uint64_t *__restrict__ ekeys, uint32_t *__restrict__ evals) {
    const uint32_t total = *total_dev;
    const uint32_t plane = g.yp * g.xp;
    const uint64_t npix = (uint64_t)plane * g.zp;
    const float d0 = g.node_key ? 1.0f : __ldg(g.div + 0), d1 = g.node_key ? 1.0f : __ldg(g.div + 1),
                d2 = g.node_key ? 1.0f : __ldg(g.div + 2);
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < total; j += gridDim.x * blockDim.x) {
        const uint32_t v = vox[j];
        const uint32_t r = parent[v];
        const uint32_t c = comp_label[r] & ~MULTI_FLAG;
        const uint32_t z = v / plane, rem = v - z * plane, y = rem / g.xp, x = rem - y * g.xp;
        const int64_t offs[6] = {-(int64_t)plane, -(int64_t)g.xp, -1, 1, (int64_t)g.xp, (int64_t)plane};
        uint32_t id[6];
#pragma unroll
        for (int d = 0; d < 6; ++d) {
            const int64_t nb = (int64_t)v + offs[d];
            id[d] = NO_NODE32;
            if (nb >= 0 && (uint64_t)nb < npix && parent[nb] == r) id[d] = lidmap[nb];
        }
        uint32_t k[3];
        if (g.node_key) {
            k[0] = __ldg(g.node_key + v);
            k[1] = k[2] = 0xFFFFFFFFu;
        } else {
            k[0] = f32_ord(flood_key_value(g, 0, d0, g.scale[0], z, y, x));
            k[1] = f32_ord(flood_key_value(g, 1, d1, g.scale[1], z, y, x));
            k[2] = f32_ord(flood_key_value(g, 2, d2, g.scale[2], z, y, x));
        }
        const uint32_t e0 = ebase[c];
        if (ebase[c + 1] != e0) {
            uint32_t h[6];
#pragma unroll
            for (int d = 0; d < 6; ++d) h[d] = id[d] == NO_NODE32 ? NO_NODE : id[d];
            rec[(size_t)j * 2] = make_uint4(h[0] | (h[1] << 16), h[2] | (h[3] << 16), h[4] | (h[5] << 16), 0u);
            rec[(size_t)j * 2 + 1] = make_uint4(0u, 0u, 0u, 0u);
            const uint32_t lid = j - cbase[c];
            const uint32_t eb = e0 + (comp_start[c + 1] - comp_start[c]) + 3u * lid;
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                ekeys[eb + a] = ((uint64_t)c << 32) | (g.node_key ? k[a] : (id[a] != NO_NODE32 ? k[a] : 0xFFFFFFFFu));
                evals[eb + a] = 3u * lid + a;
            }
        } else {
#pragma unroll
            for (int d = 0; d < 6; ++d) nbr[(size_t)j * 6 + d] = id[d];
            lab[j] = 0;
#pragma unroll
            for (int a = 0; a < 3; ++a) key[(size_t)j * 3 + a] = k[a];
        }
    }
}

}  // namespace isg
