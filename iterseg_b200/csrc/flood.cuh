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

static constexpr int FLOOD_SMEM_ENTRIES = 18816;     // class XL: heap slots [0, 18816) (12 B each, 220.5 KB)
                                                     // live in shared memory, the rest in the global arena

// Size classes of the shared-memory flood (flood_compact_kernel): a component whose
// node count + seed count fits CAP runs entirely out of shared memory
// (36 B per node: heap u64, 3 edge keys u32, label u32, 6 neighbour ids u16).
static constexpr uint32_t FLOOD_CAP_S = 384;         // 13.5 KB -> 16 warps-CTAs per SM
static constexpr uint32_t FLOOD_CAP_M = 1536;        // 54 KB   -> 4 per SM
static constexpr uint32_t FLOOD_CAP_L = 6272;        // 220.5 KB -> 1 per SM
static constexpr uint32_t FLOOD_NODE_BYTES = 36;
static constexpr uint32_t MULTI_FLAG = 0x80000000u;  // comp_label[root] = MULTI_FLAG | component index
static constexpr uint32_t NO_NODE = 0xFFFFu;

struct FloodWork {
    const uint64_t *seed_keys;     // sorted (root << 32 | padded flat index)
    const uint32_t *seed_labels;   // label of the seed at the same sorted position
    const uint32_t *comp_start;    // n_comp + 1 offsets into the sorted arrays
    const uint64_t *arena_off;     // per component offset into the heap arena
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

// ---------------------------------------------------------------------------
// Ordered flood of one component per warp over its compacted graph
// ---------------------------------------------------------------------------
// Every multi-seed component is first compacted (fill_assign_kernel, compact_graph_kernel):
// its voxels get dense local ids (their order is irrelevant: no comparison ever reaches the id
// bits), and per node the six neighbour ids (NO_NODE = not claimable) and the three edge keys
// stored at the voxel (order-preserving bits of aff / channel_max).
//   SMEM = true  (classes S/M/L): graph, labels and heap live in shared memory; heap entries
//                are single 64-bit words (key << 32 | age << 16 | node).  No global access on
//                the pop -> expand -> push critical path.
//   SMEM = false (class XL): the graph stays in global memory (L2-resident), the loads of a
//                popped node are issued before / during the sift-down so their latency hides
//                behind it; heap = (key << 32 | age, node) pairs, the first `cap` slots in
//                shared memory and the rest in the global arena.
struct CompactGraph {
    const uint32_t *cbase;         // [n_comp + 1] first node of component c (compact arena)
    const uint32_t *vox;           // [total] padded flat voxel index of every node
    const uint32_t *nbr;           // [total][6]  (NO_NODE32 = not claimable)
    const uint32_t *key;           // [total][3]
    const uint32_t *rec;           // [total][12] {6 neighbours, 6 edge keys} (global-graph mode)
    uint32_t *lab;                 // [total] node labels (class XL; zeroed by compact_graph_kernel)
    const uint32_t *lidmap;        // [npix] voxel -> local id
};
static constexpr uint32_t NO_NODE32 = 0xFFFFFFFFu;

#ifdef FLOOD_PROF
// debug: [SMEM][0..7] = pops, sift clocks, expand clocks, push clocks, pushes, prefetch hits, levels, max heap
__device__ unsigned long long g_flood_prof[2][8];
#define FP_T(x) const long long x = clock64()
#define FP_ADD(i, v) prof[i] += (unsigned long long)(v)
#else
#define FP_T(x)
#define FP_ADD(i, v)
#endif

template <bool SMEM>
__global__ void __launch_bounds__(32)
flood_graph_kernel(FloodWork w, CompactGraph cg, uint32_t cap, uint32_t *cursor, uint32_t *labels) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const unsigned FULL = 0xFFFFFFFFu;
    const unsigned lane = threadIdx.x;
    const uint32_t work_end = *w.work_end;
    // SMEM layout: heap u64[cap] | key u32[3 cap] | lab u32[cap] | nbr u16[6 cap]
    // XL layout:   heap keys u64[cap] | heap nodes u32[cap]
    uint64_t *const heap = reinterpret_cast<uint64_t *>(smem_raw);
    uint32_t *const skey = reinterpret_cast<uint32_t *>(smem_raw + (size_t)cap * 8);
    uint32_t *const slab = skey + (size_t)cap * 3;
    uint16_t *const snbr = reinterpret_cast<uint16_t *>(slab + cap);
    uint32_t *const hnode = skey;                                    // XL only
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
        uint64_t *const akeys = w.arena_keys + w.arena_off[c];       // XL heap overflow
        uint32_t *const anode = w.arena_idx + w.arena_off[c];
        __syncwarp();                                   // previous component's smem reads are done
        if (SMEM) {   // stage the component graph (coalesced copies)
            for (uint32_t i = lane; i < nv * 3; i += 32) skey[i] = gkey[i];
            for (uint32_t i = lane; i < nv * 6; i += 32) snbr[i] = (uint16_t)gnbr[i];   // NO_NODE32 -> NO_NODE
            for (uint32_t i = lane; i < nv; i += 32) slab[i] = 0;
            __syncwarp();
        }
        // heap accessors (slot -> storage)
        auto hk_ld = [&](uint32_t i) -> uint64_t {
            if (SMEM) return heap[i];
            return i < cap ? heap[i] : akeys[i];
        };
        auto hk_st = [&](uint32_t i, uint64_t k) {
            if (SMEM || i < cap) heap[i] = k; else akeys[i] = k;
        };
        auto hn_ld = [&](uint32_t i) -> uint32_t { return i < cap ? hnode[i] : anode[i]; };
        auto hn_st = [&](uint32_t i, uint32_t x) { if (i < cap) hnode[i] = x; else anode[i] = x; };

        const uint64_t zero_hi = (uint64_t)f32_ord(0.0f) << 32;
        for (uint32_t i = lane; i < cnt; i += 32) {
            const uint32_t v = (uint32_t)(w.seed_keys[s0 + i] & 0xFFFFFFFFu);
            const uint32_t lid = cg.lidmap[v];
            const uint32_t l = labels[v];    // the final seed label (duplicates: largest, set upstream)
            if (SMEM) {
                heap[i] = zero_hi | ((uint64_t)i << 16) | lid;     // ascending array == valid heap
                slab[lid] = l;
            } else {
                hk_st(i, zero_hi | i);
                hn_st(i, lid);
                __stcg(glab + lid, l);
            }
        }
        uint32_t n = cnt;
        uint32_t age = cnt;
        uint32_t pref_node = NO_NODE32, pref_nb = NO_NODE32;   // XL: speculative adjacency prefetch
#ifdef FLOOD_PROF
        unsigned long long prof[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#endif
        __syncwarp();

        while (n > 0) {
            // ---- pop the minimum -------------------------------------------------
            FP_T(t0);
            FP_ADD(0, 1);
#ifdef FLOOD_PROF
            if (n > prof[7]) prof[7] = n;
#endif
            uint32_t p;
            if (SMEM) p = (uint32_t)heap[0] & 0xFFFFu; else p = hnode[0];
            --n;
            const uint64_t lastk = hk_ld(n);
            uint32_t lastn = 0;
            if (!SMEM) lastn = hn_ld(n);
            // expansion loads, part 1 (in flight during the sift-down)
            uint32_t nb = NO_NODE32, labp;
            if (SMEM) {
                if (lane < 6) { const uint32_t t = snbr[p * 6u + lane]; nb = t == NO_NODE ? NO_NODE32 : t; }
                labp = slab[p];
            } else {
                // the adjacency of the node that was on top after the previous sift-down was
                // requested back then; unless a push displaced it, it has long arrived
                if (pref_node == p) nb = pref_nb;
                else if (lane < 6) nb = __ldg(gnbr + (size_t)p * 6u + lane);
                labp = __ldcg(glab + p);
            }
            const bool pref_hit = !SMEM && pref_node == p;
            uint32_t labn = 1, kord = 0;
            bool fetched = false;
            auto fetch2 = [&]() {        // expansion loads, part 2: need nb
                if (nb != NO_NODE32) {
                    // edge key: stored at the popped voxel for the three negative directions,
                    // at the neighbour for the positive ones
                    const size_t kn = (size_t)(lane < 3 ? p : nb) * 3u + axis;
                    if (SMEM) { labn = slab[nb]; kord = skey[kn]; }
                    else { labn = __ldcg(glab + nb); kord = __ldg(gkey + kn); }
                }
                fetched = true;
            };
            if (pref_hit) { fetch2(); FP_ADD(5, 1); }    // adjacency already here: start the dependent loads
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
                        if (!SMEM) hn_st(i, hn_ld(wc));
                    }
                    __syncwarp();
                    i = wc;
                    FP_ADD(6, 1);
                    if (!fetched) fetch2();
                }
                if (lane == 0) {
                    hk_st(i, lastk);
                    if (!SMEM) hn_st(i, lastn);
                }
                __syncwarp();
                if (!SMEM) {                             // request the adjacency of the likely next pop
                    pref_node = hnode[0];
                    pref_nb = NO_NODE32;
                    if (lane < 6) pref_nb = __ldg(gnbr + (size_t)pref_node * 6u + lane);
                }
            }
            if (!fetched) fetch2();
            FP_T(t1);
            // ---- expand the popped node (watershed.py:135-154) ---------------------
            const bool claim = nb != NO_NODE32 && labn == 0;
            unsigned bits = __ballot_sync(FULL, claim);
            FP_T(t2);
            FP_ADD(4, __popc(bits));
            if (claim) {                                  // labelled at push time (:149)
                if (SMEM) slab[nb] = labp; else __stcg(glab + nb, labp);
            }
            const uint32_t myage = age + __popc(bits & ((1u << lane) - 1u));
            const uint64_t mykey = SMEM ? (((uint64_t)kord << 32) | ((uint64_t)myage << 16) | nb)
                                        : (((uint64_t)kord << 32) | myage);
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
                        if (!SMEM) hn_st(i, hn_ld(par));
                    }
                    i = par;
                }
                if (lane == 0) {
                    hk_st(i, k);
                    if (!SMEM) hn_st(i, kn);
                }
                __syncwarp();
            }
            FP_T(t3);
            FP_ADD(1, t1 - t0);
            FP_ADD(2, t2 - t1);
            FP_ADD(3, t3 - t2);
        }
#ifdef FLOOD_PROF
        if (lane == 0 && nv > 5000) {
            for (int q = 0; q < 7; ++q) atomicAdd(&g_flood_prof[SMEM ? 1 : 0][q], prof[q]);
            atomicMax(&g_flood_prof[SMEM ? 1 : 0][7], prof[7]);
        }
#endif
        // ---- write the component's labels back ---------------------------------------
        __syncwarp();
        for (uint32_t i = lane; i < nv; i += 32)
            labels[cg.vox[b0 + i]] = SMEM ? slab[i] : __ldcg(glab + i);
    }
}

// ---------------------------------------------------------------------------
// flood_pq_kernel: the fast ordered flood (classes S / M / L with the graph in shared
// memory, class G with the graph in global memory)
// ---------------------------------------------------------------------------
// Priority queue = a 32-ary min-heap in shared memory PLUS a small pending buffer of
// PS entries held in (warp-uniform) registers.  New keys go to the pending buffer for
// free; a pop takes min(pending minimum, heap top).  The very common cascade "a freshly
// pushed voxel is the next one popped" therefore never touches the heap, and a heap pop
// re-uses the vacated root for the largest pending entry (one sift-down instead of a
// sift-down plus a sift-up).  Keys are unique 64-bit words (value bits | age | node), so
// any implementation of "pop the minimum" yields the reference order.
//   GSMEM = true : per node {6 x u16 neighbour, 3 x u32 key, u32 label} staged in smem.
//   GSMEM = false: per node one 48-byte record {6 neighbours, 6 edge keys} in global
//                  memory, requested when the node enters the queue (pending slot
//                  registers) or becomes the heap top, so that its latency hides behind
//                  the queue work; claimed-bits in smem, labels in global memory.
static constexpr int FLOOD_PS = 8;                   // pending-buffer slots
static constexpr uint32_t FLOOD_CAP_G = 28160;       // class G: heap entries (8 B) + claimed bits: 223.4 KB

template <bool GSMEM>
__global__ void __launch_bounds__(32)
flood_pq_kernel(FloodWork w, CompactGraph cg, uint32_t cap, uint32_t *cursor, uint32_t *labels) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const unsigned FULL = 0xFFFFFFFFu;
    const unsigned lane = threadIdx.x;
    const uint32_t work_end = *w.work_end;
    // GSMEM layout: heap u64[cap] | key u32[3 cap] | lab u32[cap] | nbr u16[6 cap]
    // global-graph layout: heap u64[cap] | claimed bits u32[cap / 32]
    uint64_t *const heap = reinterpret_cast<uint64_t *>(smem_raw);
    uint32_t *const skey = reinterpret_cast<uint32_t *>(smem_raw + (size_t)cap * 8);
    uint32_t *const slab = skey + (size_t)cap * 3;
    uint16_t *const snbr = reinterpret_cast<uint16_t *>(slab + cap);
    uint32_t *const sbits = skey;                                    // global-graph mode
    const int axis = lane < 3 ? (int)lane : (lane < 6 ? 5 - (int)lane : 0);   // [0,1,2,2,1,0]
    const uint64_t EMPTY = ~0ull;

    for (;;) {
        uint32_t wi = 0;
        if (lane == 0) wi = atomicAdd(cursor, 1u);
        wi = __shfl_sync(FULL, wi, 0);
        if (wi >= work_end) break;
        const uint32_t c = w.order[wi];
        const uint32_t s0 = w.comp_start[c], cnt = w.comp_start[c + 1] - s0;
        const uint32_t b0 = cg.cbase[c], nv = cg.cbase[c + 1] - b0;
        const uint32_t *const grec = cg.rec + (size_t)b0 * 12;
        uint32_t *const glab = cg.lab + b0;
        __syncwarp();                                   // previous component's smem reads are done
        if (GSMEM) {   // stage the component graph (coalesced copies)
            const uint32_t *const gkey = cg.key + (size_t)b0 * 3;
            const uint32_t *const gnbr = cg.nbr + (size_t)b0 * 6;
            for (uint32_t i = lane; i < nv * 3; i += 32) skey[i] = gkey[i];
            for (uint32_t i = lane; i < nv * 6; i += 32) snbr[i] = (uint16_t)gnbr[i];   // NO_NODE32 -> NO_NODE
            for (uint32_t i = lane; i < nv; i += 32) slab[i] = 0;
        } else {
            for (uint32_t i = lane; i < (nv + 31) / 32; i += 32) sbits[i] = 0;
        }
        __syncwarp();
        const uint64_t zero_hi = (uint64_t)f32_ord(0.0f) << 32;
        for (uint32_t i = lane; i < cnt; i += 32) {
            const uint32_t v = (uint32_t)(w.seed_keys[s0 + i] & 0xFFFFFFFFu);
            const uint32_t lid = cg.lidmap[v];
            const uint32_t l = labels[v];    // the final seed label (duplicates: largest, set upstream)
            heap[i] = zero_hi | ((uint64_t)i << 16) | lid;         // ascending array == valid heap
            if (GSMEM) slab[lid] = l;
            else { atomicOr(sbits + (lid >> 5), 1u << (lid & 31)); __stcg(glab + lid, l); }
        }
        uint32_t n = cnt;
        uint32_t age = cnt;
        __syncwarp();

        // pending buffer (warp-uniform) + per-slot node data (lanes 0..5, global-graph mode)
        uint64_t pk[FLOOD_PS];
        uint32_t plab[FLOOD_PS], pnb[FLOOD_PS], pkey[FLOOD_PS];
#pragma unroll
        for (int s = 0; s < FLOOD_PS; ++s) { pk[s] = EMPTY; plab[s] = 0; pnb[s] = NO_NODE32; pkey[s] = 0; }
        uint64_t topk = heap[0];                                     // cached heap top (EMPTY = heap empty)
        // data of the heap top (global-graph mode): requested as soon as the top is known
        uint64_t hfor = EMPTY;
        uint32_t hnb = NO_NODE32, hkey = 0, hlab = 0;
        auto request_top = [&]() {
            if (!GSMEM && topk != EMPTY && hfor != topk) {
                const uint32_t node = (uint32_t)topk & 0xFFFFu;
                hnb = NO_NODE32;
                if (lane < 6) {
                    hnb = __ldg(grec + (size_t)node * 12 + lane);
                    hkey = __ldg(grec + (size_t)node * 12 + 6 + lane);
                }
                hlab = __ldcg(glab + node);
                hfor = topk;
            }
        };
        request_top();
#ifdef FLOOD_PROF
        unsigned long long prof[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        const long long tc0 = clock64();
#endif

        for (;;) {
            // ---- minimum and maximum of the pending buffer (uniform, no communication) ----
            uint64_t mP = EMPTY, xP = 0;
            int sP = -1, tP = -1;
#pragma unroll
            for (int s = 0; s < FLOOD_PS; ++s) {
                if (pk[s] < mP) { mP = pk[s]; sP = s; }
                if (pk[s] != EMPTY && pk[s] >= xP) { xP = pk[s]; tP = s; }
            }
            if (mP == EMPTY && topk == EMPTY) break;
            FP_ADD(0, 1);
            uint32_t p, nb = NO_NODE32, kord = 0, labp = 0;
            if (mP < topk) {
                // ---- pop from the pending buffer ----
                FP_ADD(1, 1);
                p = (uint32_t)mP & 0xFFFFu;
#pragma unroll
                for (int s = 0; s < FLOOD_PS; ++s)
                    if (s == sP) { nb = pnb[s]; kord = pkey[s]; labp = plab[s]; pk[s] = EMPTY; }
            } else {
                // ---- pop the heap top; the root is refilled with the largest pending entry
                //      (or the last heap entry) and sifted down ----
                p = (uint32_t)topk & 0xFFFFu;
                request_top();
                nb = hnb; kord = hkey; labp = hlab;
                uint64_t ins;
                if (tP >= 0) {
                    ins = xP;
#pragma unroll
                    for (int s = 0; s < FLOOD_PS; ++s)
                        if (s == tP) pk[s] = EMPTY;
                } else {
                    --n;
                    ins = n > 0 ? heap[n] : EMPTY;
                }
                __syncwarp();
                if (n == 0) {
                    topk = EMPTY;
                } else {
                    uint32_t i = 0;
                    bool top_known = false;
                    for (;;) {
                        const uint32_t c0 = i * 32u + 1u;
                        if (c0 >= n) break;
                        const uint32_t ch = c0 + lane;
                        const uint64_t k = ch < n ? heap[ch] : EMPTY;
                        const uint32_t hi = (uint32_t)(k >> 32);
                        const uint32_t mhi = __reduce_min_sync(FULL, hi);
                        const uint32_t lo = hi == mhi ? (uint32_t)k : 0xFFFFFFFFu;
                        const uint32_t mlo = __reduce_min_sync(FULL, lo);
                        const uint64_t mk = ((uint64_t)mhi << 32) | mlo;
                        if (mk >= ins) break;
                        const uint32_t win = __ffs(__ballot_sync(FULL, hi == mhi && lo == mlo)) - 1;
                        if (lane == 0) heap[i] = mk;
                        __syncwarp();
                        if (!top_known) { topk = mk; top_known = true; request_top(); }
                        i = c0 + win;
                        FP_ADD(6, 1);
                    }
                    if (lane == 0) heap[i] = ins;
                    __syncwarp();
                    if (!top_known) { topk = ins; request_top(); }
                }
            }
            // ---- expand the popped node (watershed.py:135-154) ---------------------
            bool claim = false;
            if (GSMEM) {
                labp = slab[p];
                if (lane < 6) { const uint32_t t = snbr[p * 6u + lane]; nb = t == NO_NODE ? NO_NODE32 : t; }
                if (nb != NO_NODE32) {
                    claim = slab[nb] == 0;
                    // edge key: stored at the popped voxel for the three negative directions,
                    // at the neighbour for the positive ones
                    kord = skey[(lane < 3 ? p : nb) * 3u + axis];
                }
            } else {
                if (nb != NO_NODE32) claim = ((sbits[nb >> 5] >> (nb & 31)) & 1u) == 0;
            }
            unsigned bits = __ballot_sync(FULL, claim);
            if (claim) {                                  // labelled at push time (:149)
                if (GSMEM) slab[nb] = labp;
                else { atomicOr(sbits + (nb >> 5), 1u << (nb & 31)); __stcg(glab + nb, labp); }
            }
            const uint32_t myage = age + __popc(bits & ((1u << lane) - 1u));
            const uint64_t mykey = ((uint64_t)kord << 32) | ((uint64_t)myage << 16) | (nb & 0xFFFFu);
            age += __popc(bits);
            FP_ADD(4, __popc(bits));
            __syncwarp();
            // ---- queue the claimed neighbours ------------------------------------------
            while (bits) {
                const int src = __ffs(bits) - 1;
                bits &= bits - 1;
                const uint64_t k = __shfl_sync(FULL, mykey, src);
                int fs = -1;
#pragma unroll
                for (int s = FLOOD_PS - 1; s >= 0; --s)
                    if (pk[s] == EMPTY) fs = s;
                if (fs >= 0) {
                    const uint32_t node = (uint32_t)k & 0xFFFFu;
#pragma unroll
                    for (int s = 0; s < FLOOD_PS; ++s)
                        if (s == fs) {
                            pk[s] = k;
                            if (!GSMEM) {
                                plab[s] = labp;
                                pnb[s] = NO_NODE32;
                                if (lane < 6) {
                                    pnb[s] = __ldg(grec + (size_t)node * 12 + lane);
                                    pkey[s] = __ldg(grec + (size_t)node * 12 + 6 + lane);
                                }
                            }
                        }
                } else {
                    // pending buffer full: regular heap push
                    FP_ADD(5, 1);
                    uint32_t i = n++;
                    while (i > 0) {
                        const uint32_t par = (i - 1u) >> 5;
                        const uint64_t pkv = heap[par];
                        if (pkv <= k) break;
                        if (lane == 0) heap[i] = pkv;
                        i = par;
                    }
                    if (lane == 0) heap[i] = k;
                    __syncwarp();
                    if (i == 0) { topk = k; request_top(); }
                }
            }
        }
#ifdef FLOOD_PROF
        if (lane == 0 && nv > 5000) {
            prof[2] = (unsigned long long)(clock64() - tc0);
            for (int q = 0; q < 7; ++q) atomicAdd(&g_flood_prof[GSMEM ? 1 : 0][q], prof[q]);
        }
#endif
        // ---- write the component's labels back ---------------------------------------
        __syncwarp();
        for (uint32_t i = lane; i < nv; i += 32)
            labels[cg.vox[b0 + i]] = GSMEM ? slab[i] : __ldcg(glab + i);
    }
}

// Pass 1 of the compaction, fused with the single-seed fill: every voxel of a
// single-seed component takes the seed's label; every voxel of a compacted multi-seed
// component gets a local id (warp-aggregated atomics per component).
__global__ void __launch_bounds__(256)
fill_assign_kernel(const uint32_t *__restrict__ parent, const uint32_t *__restrict__ comp_label,
                   const uint8_t *__restrict__ mask, uint32_t *__restrict__ labels,
                   const uint32_t *__restrict__ cbase, uint32_t *__restrict__ ccursor,
                   uint32_t *__restrict__ lidmap, uint32_t *__restrict__ vox, uint64_t n) {
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
                    c = cl & ~MULTI_FLAG;
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

// Pass 2: adjacency and edge keys of every compacted node.
__global__ void __launch_bounds__(256)
compact_graph_kernel(FloodGeom g, const uint32_t *__restrict__ parent,
                     const uint32_t *__restrict__ lidmap, const uint32_t *__restrict__ vox,
                     const uint32_t *__restrict__ total_dev, uint32_t *__restrict__ nbr,
                     uint32_t *__restrict__ key, uint32_t *__restrict__ lab, uint32_t *__restrict__ rec) {
    const uint32_t total = *total_dev;
    const uint32_t plane = g.yp * g.xp;
    const uint64_t npix = (uint64_t)plane * g.zp;
    const float d0 = __ldg(g.div + 0), d1 = __ldg(g.div + 1), d2 = __ldg(g.div + 2);
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < total; j += gridDim.x * blockDim.x) {
        const uint32_t v = vox[j];
        const uint32_t r = parent[v];
        const uint32_t z = v / plane, rem = v - z * plane, y = rem / g.xp, x = rem - y * g.xp;
        const int64_t offs[6] = {-(int64_t)plane, -(int64_t)g.xp, -1, 1, (int64_t)g.xp, (int64_t)plane};
#pragma unroll
        for (int d = 0; d < 6; ++d) {
            const int64_t nb = (int64_t)v + offs[d];
            uint32_t id = NO_NODE32;
            if (nb >= 0 && (uint64_t)nb < npix && parent[nb] == r) id = lidmap[nb];
            nbr[(size_t)j * 6 + d] = id;
            rec[(size_t)j * 12 + d] = id;
        }
        lab[j] = 0;
        const uint32_t k0 = f32_ord(flood_key_value(g, 0, d0, g.scale[0], z, y, x));
        const uint32_t k1 = f32_ord(flood_key_value(g, 1, d1, g.scale[1], z, y, x));
        const uint32_t k2 = f32_ord(flood_key_value(g, 2, d2, g.scale[2], z, y, x));
        key[(size_t)j * 3 + 0] = k0;
        key[(size_t)j * 3 + 1] = k1;
        key[(size_t)j * 3 + 2] = k2;
        // all six edge keys of the node, in neighbour order [-z,-y,-x,+x,+y,+z]: the positive
        // directions read the value stored at the neighbour
        rec[(size_t)j * 12 + 6] = k0;
        rec[(size_t)j * 12 + 7] = k1;
        rec[(size_t)j * 12 + 8] = k2;
        rec[(size_t)j * 12 + 9] = f32_ord(flood_key_value(g, 2, d2, g.scale[2], z, y, x + 1));
        rec[(size_t)j * 12 + 10] = f32_ord(flood_key_value(g, 1, d1, g.scale[1], z, y + 1, x));
        rec[(size_t)j * 12 + 11] = f32_ord(flood_key_value(g, 0, d0, g.scale[0], z + 1, y, x));
    }
}

}  // namespace isg
