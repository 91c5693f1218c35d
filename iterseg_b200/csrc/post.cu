// Feature map -> labels on the device: isg_segment_features.
//
// Replaces segment_output_image (src/iterseg/watershed.py:165-223):
//   affinity normalisation   :194-201  (per-channel max; the divide is done inside the flood)
//   _get_centroids           :232-236  (gaussian (0,1,1) -> 3x3x3 peaks > 0.04, sorted)
//   _get_mask                :226-229  (otsu of the sigma=2 smoothed channel; raw > thr)
//   _remove_unwanted_objects :239-251  (6-connected components, size window, seed filter)
//   affinity_watershed       :17-35    (flood_stage_run)
//
// Arithmetic contracts (so that results are identical to scipy / numpy, see
// SURVEY.md Appendix B and oracle/skimage_shim.py):
//   * each 1-D Gaussian pass accumulates in float64 in scipy's order
//     x[c]*w[0] + sum_{j=r..1} (x[c-j] + x[c+j])*w[j], no FMA contraction, boundary
//     'nearest', and stores float32 between passes;
//   * the histogram bin index, the float32 bin edges / centres and the float32
//     sequential cumulative sums follow numpy.histogram / numpy.cumsum.
#include <cub/cub.cuh>

#include "flood_stage.h"
#include "gauss.cuh"

namespace isg {

// per-channel maximum of up to 3 planes (np.max(affinities, axis=(1,2,3)), watershed.py:195)
__global__ void __launch_bounds__(256)
chan_max_kernel(const float *__restrict__ feats, uint64_t chan_stride, uint64_t n, int c0, int c1, int c2,
                uint32_t *__restrict__ out_ord) {
    const int ch = blockIdx.y == 0 ? c0 : (blockIdx.y == 1 ? c1 : c2);
    const float *base = feats + (uint64_t)ch * chan_stride;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    float m = -INFINITY;
    if ((reinterpret_cast<uintptr_t>(base) & 15u) == 0) {
        const float4 *p = reinterpret_cast<const float4 *>(base);
        const uint64_t n4 = n / 4;
        for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
            float4 v = __ldg(p + i);
            m = fmaxf(fmaxf(m, fmaxf(v.x, v.y)), fmaxf(v.z, v.w));
        }
        if (blockIdx.x == 0 && threadIdx.x < (n & 3)) m = fmaxf(m, base[n4 * 4 + threadIdx.x]);
    } else {
        for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
            m = fmaxf(m, __ldg(base + i));
    }
    uint32_t k = __reduce_max_sync(0xFFFFFFFFu, f32_ord(m));
    if ((threadIdx.x & 31) == 0) atomicMax(out_ord + blockIdx.y, k);
}

__global__ void ord_to_float_kernel(const uint32_t *__restrict__ in, float *__restrict__ out, int n) {
    int i = threadIdx.x;
    if (i < n) out[i] = ord_f32(in[i]);
}

// 3x3x3 peaks of the smoothed centre map (peak_local_max):
// candidates = interior voxels > thr that no voxel of their 3x3x3 neighbourhood exceeds.
// `nontrivial` reproduces skimage's "every voxel equals its local maximum -> no peaks" rule: on a
// connected grid that happens exactly when the image is constant, i.e. no voxel differs from
// voxel 0.  Only voxels above the threshold (few) pay for the 26 neighbour loads.
// One warp per (z, y) row, lanes along x.
__global__ void __launch_bounds__(256)
local_max_kernel(const float *__restrict__ cs, uint32_t Z, uint32_t Y, uint32_t X, float thr,
                 uint64_t *__restrict__ cand, uint32_t cap, uint32_t *__restrict__ n_cand,
                 uint32_t *__restrict__ nontrivial) {
    const uint32_t rows = Z * Y;
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    const float c0 = __ldg(cs);
    bool differs = false;
    for (uint32_t row = wid; row < rows; row += nw) {
        const uint32_t z = row / Y, y = row - z * Y;
        const float *r = cs + (uint64_t)row * X;
        const bool row_interior = z > 0 && z + 1 < Z && y > 0 && y + 1 < Y;
        for (uint32_t x = lane; x < X; x += 32) {
            const float c = r[x];
            differs |= c != c0;
            if (!(row_interior && x > 0 && x + 1 < X && c > thr)) continue;
            bool is_max = true;
#pragma unroll
            for (int dz = -1; dz <= 1; ++dz)
#pragma unroll
                for (int dy = -1; dy <= 1; ++dy) {
                    const float *q = r + ((int64_t)dz * Y + dy) * (int64_t)X + x;
                    is_max &= !(__ldg(q - 1) > c) & !(__ldg(q) > c) & !(__ldg(q + 1) > c);
                }
            if (is_max) {
                uint32_t slot = atomicAdd(n_cand, 1u);
                if (slot < cap) cand[slot] = ((uint64_t)(~f32_ord(c)) << 32) | ((uint64_t)row * X + x);
            }
        }
    }
    if (__any_sync(0xFFFFFFFFu, differs) && lane == 0) atomicOr(nontrivial, 1u);
}

// numpy.histogram(smoothed, 256) bin index (uniform-bin fast path, float32)
__device__ __forceinline__ int np_hist_bin(float a, float first, float denom, const float *edges) {
    float f = __fmul_rn(__fdiv_rn(__fsub_rn(a, first), denom), 256.0f);
    int idx = (int)f;
    if (idx == 256) idx = 255;
    if (a < edges[idx]) idx -= 1;
    else if (a >= edges[idx + 1] && idx != 255) idx += 1;
    return idx;
}

__global__ void hist_edges_kernel(const uint32_t *__restrict__ minmax, float *__restrict__ edges) {
    // np.linspace(first, last, 257, dtype=float32): arange * step + first, last forced
    const float first = ord_f32(minmax[0]), last = ord_f32(minmax[1]);
    const float step = __fdiv_rn(__fsub_rn(last, first), 256.0f);
    int i = threadIdx.x;
    if (i <= 256) {
        float e;
        if (step == 0.0f) e = __fadd_rn(__fmul_rn(__fdiv_rn((float)i, 256.0f), __fsub_rn(last, first)), first);
        else e = __fadd_rn(__fmul_rn((float)i, step), first);
        if (i == 256) e = last;
        edges[i] = e;
    }
}

__global__ void __launch_bounds__(256)
hist_kernel(const float *__restrict__ s, uint64_t n, const uint32_t *__restrict__ minmax,
            const float *__restrict__ edges_g, unsigned long long *__restrict__ hist) {
    __shared__ float edges[257];
    __shared__ uint32_t h[256];
    for (int i = threadIdx.x; i < 257; i += blockDim.x) edges[i] = edges_g[i];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) h[i] = 0;
    __syncthreads();
    const float first = ord_f32(minmax[0]), last = ord_f32(minmax[1]);
    const float denom = __fsub_rn(last, first);
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    if (denom > 0.0f) {
        for (uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; v < n; v += stride)
            atomicAdd(&h[np_hist_bin(__ldg(s + v), first, denom, edges)], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 256; i += blockDim.x)
        if (h[i]) atomicAdd(hist + i, (unsigned long long)h[i]);
}

// skimage.filters.threshold_otsu on the 256-bin histogram, float32 sequential
// arithmetic exactly as numpy evaluates it (oracle/skimage_shim.py).
__global__ void otsu_kernel(const unsigned long long *__restrict__ hist,
                            const float *__restrict__ edges, const uint32_t *__restrict__ minmax,
                            float *__restrict__ thr_out) {
    __shared__ float cnt[256], ctr[256], w1[256], w2[256], m1[256], m2[256];
    if (threadIdx.x != 0) return;
    const float first = ord_f32(minmax[0]), last = ord_f32(minmax[1]);
    if (!(last > first)) {          // constant image: threshold_otsu returns that value
        *thr_out = first;
        return;
    }
    for (int i = 0; i < 256; ++i) {
        cnt[i] = (float)hist[i];
        ctr[i] = __fdiv_rn(__fadd_rn(edges[i], edges[i + 1]), 2.0f);
    }
    float acc = 0.0f, accm = 0.0f;
    for (int i = 0; i < 256; ++i) {
        acc = i == 0 ? cnt[0] : __fadd_rn(acc, cnt[i]);
        float pm = __fmul_rn(cnt[i], ctr[i]);
        accm = i == 0 ? pm : __fadd_rn(accm, pm);
        w1[i] = acc;
        m1[i] = __fdiv_rn(accm, acc);
    }
    for (int i = 255; i >= 0; --i) {
        acc = i == 255 ? cnt[255] : __fadd_rn(acc, cnt[i]);
        float pm = __fmul_rn(cnt[i], ctr[i]);
        accm = i == 255 ? pm : __fadd_rn(accm, pm);
        w2[i] = acc;
        m2[i] = __fdiv_rn(accm, acc);
    }
    int best = 0;
    float bestv = 0.0f;
    for (int i = 0; i < 255; ++i) {
        float d = __fsub_rn(m1[i], m2[i + 1]);
        float var = __fmul_rn(__fmul_rn(w1[i], w2[i + 1]), __fmul_rn(d, d));
        if (i == 0 || var > bestv) {          // np.argmax: first maximum; NaN never arises here
            best = i;
            bestv = var;
        }
    }
    *thr_out = ctr[best];
}

// mask = raw > thr, zero-padded by one voxel (watershed.py:207-213)
__global__ void __launch_bounds__(256)
mask_kernel(const float *__restrict__ raw, uint32_t Z, uint32_t Y, uint32_t X,
            const float *__restrict__ thr_dev, float thr_abs, int use_abs,
            uint8_t *__restrict__ mask) {
    const uint32_t Yp = Y + 2, Xp = X + 2;
    const uint64_t np = (uint64_t)(Z + 2) * Yp * Xp;
    const float thr = use_abs ? thr_abs : *thr_dev;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; v < np; v += stride) {
        uint32_t x = (uint32_t)(v % Xp);
        uint64_t t = v / Xp;
        uint32_t y = (uint32_t)(t % Yp);
        uint32_t z = (uint32_t)(t / Yp);
        uint8_t m = 0;
        if (x >= 1 && x <= X && y >= 1 && y <= Y && z >= 1 && z <= Z)
            m = __ldg(raw + ((uint64_t)(z - 1) * Y + (y - 1)) * X + (x - 1)) > thr ? 1 : 0;
        mask[v] = m;
    }
}

// Single CTA: walk the sorted candidates, keep those whose component survives the
// size window (watershed.py:241-249), compact in order, label them 1..N.
__global__ void __launch_bounds__(1024)
seed_filter_kernel(const uint64_t *__restrict__ cand_sorted, uint32_t n_cand,
                   const uint32_t *__restrict__ nontrivial, const uint32_t *__restrict__ parent,
                   const uint32_t *__restrict__ comp_size, uint64_t min_area, uint64_t max_area,
                   uint32_t Z, uint32_t Y, uint32_t X, int64_t *__restrict__ seeds_out,
                   uint32_t cap, uint32_t *__restrict__ labels, uint32_t *__restrict__ n_kept_out,
                   unsigned long long *__restrict__ keys_out /* nullable */) {
    typedef cub::BlockScan<uint32_t, 1024> Scan;
    __shared__ typename Scan::TempStorage tmp;
    __shared__ uint32_t carry;
    const uint32_t t = threadIdx.x;
    if (t == 0) carry = 0;
    __syncthreads();
    if (*nontrivial == 0) n_cand = 0;          // constant image: peak_local_max returns nothing
    const uint32_t Yp = Y + 2, Xp = X + 2;
    for (uint32_t base = 0; base < n_cand; base += 1024) {
        uint32_t i = base + t;
        uint32_t keep = 0;
        uint64_t p = 0;
        if (i < n_cand) {
            uint32_t u = (uint32_t)(cand_sorted[i] & 0xFFFFFFFFu);
            uint32_t x = u % X;
            uint32_t r = u / X;
            uint32_t y = r % Y;
            uint32_t z = r / Y;
            p = ((uint64_t)(z + 1) * Yp + (y + 1)) * Xp + (x + 1);
            uint32_t root = parent[p];
            if (root != CCL_NONE) {
                uint64_t sz = comp_size[root];
                keep = (sz >= min_area && sz < max_area) ? 1u : 0u;
            }
        }
        uint32_t pos, total;
        Scan(tmp).ExclusiveSum(keep, pos, total);
        if (keep) {
            uint32_t k = carry + pos;
            if (k < cap) {
                seeds_out[k] = (int64_t)p;
                labels[p] = k + 1;
                if (keys_out) keys_out[k] = cand_sorted[i];
            }
        }
        __syncthreads();
        if (t == 0) carry += total;
        __syncthreads();
    }
    if (t == 0) *n_kept_out = carry < cap ? carry : cap;
}

__global__ void __launch_bounds__(256)
mask_keep_kernel(const uint8_t *__restrict__ mask0, const uint32_t *__restrict__ parent,
                 const uint32_t *__restrict__ comp_size, uint64_t min_area, uint64_t max_area,
                 uint8_t *__restrict__ mask_out, uint64_t np, uint32_t *__restrict__ n_components) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    uint32_t roots = 0;
    for (uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; v < np; v += stride) {
        uint8_t m = 0;
        uint32_t r = parent[v];
        if (r != CCL_NONE) {
            uint64_t sz = comp_size[r];
            m = (sz >= min_area && sz < max_area) ? 1 : 0;
            if (r == (uint32_t)v && m) roots++;
        }
        mask_out[v] = m;
    }
    roots = __reduce_add_sync(0xFFFFFFFFu, roots);
    if ((threadIdx.x & 31) == 0 && roots) atomicAdd(n_components, roots);
}

// Slab mode: components that touch an open z face of the slab are only partially known.
__global__ void __launch_bounds__(256)
slab_face_kernel(const uint32_t *__restrict__ parent, uint32_t *__restrict__ flag, uint32_t Zp,
                 uint32_t Yp, uint32_t Xp, int open_faces) {
    const uint64_t plane = (uint64_t)Yp * Xp;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < 2 * plane; i += stride) {
        const int face = i >= plane;
        if (!((open_faces >> face) & 1)) continue;
        const uint64_t v = (face ? (uint64_t)(Zp - 2) : 1ull) * plane + (i - face * plane);
        const uint32_t r = parent[v];
        if (r != CCL_NONE) flag[r] = 1u;
    }
}
__global__ void __launch_bounds__(256)
slab_guard_kernel(const uint32_t *__restrict__ parent, const uint32_t *__restrict__ flag, uint32_t Yp,
                  uint32_t Xp, uint32_t own_z0, uint32_t own_z1, uint32_t *__restrict__ violation) {
    const uint64_t plane = (uint64_t)Yp * Xp;
    const uint64_t v0 = (uint64_t)(own_z0 + 1) * plane, v1 = (uint64_t)(own_z1 + 1) * plane;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    bool bad = false;
    for (uint64_t v = v0 + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; v < v1; v += stride) {
        const uint32_t r = parent[v];
        if (r != CCL_NONE && flag[r]) bad = true;
    }
    if (__any_sync(0xFFFFFFFFu, bad) && (threadIdx.x & 31) == 0) atomicOr(violation, 1u);
}

__global__ void counts_kernel(const uint32_t *n_kept, const uint32_t *n_cand, const uint32_t *n_components,
                              const uint32_t *n_multi, const uint32_t *violation, int64_t *out) {
    out[0] = *n_kept;
    out[1] = *n_cand;
    out[2] = *n_components;
    out[3] = *n_multi;
    out[4] = *violation;
    out[5] = out[6] = out[7] = 0;
}

struct PostBuffers {
    float *tmp_a, *tmp_b;
    uint64_t *cand_a, *cand_b;
    unsigned char *cub_tmp;
    size_t cub_bytes;
    uint32_t *scal;        // 0,1: minmax ord; 2: n_cand; 3: nontrivial; 4: n_kept; 5: n_components; 6: halo violation; 8..10: aff max ord
    float *fscal;          // 0..2: aff max; 3: otsu thr
    float *edges;
    unsigned long long *hist;
    uint8_t *mask0;
    uint32_t *parent, *comp_size, *comp_label;
    FloodStageBuffers flood;
};

static void post_carve(PostBuffers *b, Carver &cv, uint64_t n, uint64_t np, int64_t max_seeds, uint64_t node_cap) {
    b->tmp_a = cv.take<float>(n);
    b->tmp_b = cv.take<float>(n);
    b->cand_a = cv.take<uint64_t>(max_seeds);
    b->cand_b = cv.take<uint64_t>(max_seeds);
    size_t cub_bytes = 0;
    cub::DeviceRadixSort::SortKeys(nullptr, cub_bytes, (uint64_t *)nullptr, (uint64_t *)nullptr,
                                   (int)max_seeds);
    b->cub_bytes = cub_bytes + 256;
    b->cub_tmp = cv.take<unsigned char>(b->cub_bytes);
    b->scal = cv.take<uint32_t>(64);
    b->fscal = cv.take<float>(64);
    b->edges = cv.take<float>(320);
    b->hist = cv.take<unsigned long long>(256);
    b->mask0 = cv.take<uint8_t>(np);
    b->parent = cv.take<uint32_t>(np);
    b->comp_size = cv.take<uint32_t>(np);
    b->comp_label = cv.take<uint32_t>(np);
    flood_stage_workspace(&b->flood, cv, np, max_seeds, node_cap);
}

}  // namespace isg

using namespace isg;

static size_t post_bytes(uint64_t n, uint64_t np, int64_t max_seeds, uint64_t node_cap) {
    Carver cv(nullptr, 0);
    PostBuffers b;
    post_carve(&b, cv, n, np, max_seeds, node_cap);
    return cv.off + 512;
}

extern "C" size_t isg_post_workspace_bytes(int64_t z, int64_t y, int64_t x, int64_t max_seeds) {
    if (z <= 0 || y <= 0 || x <= 0 || max_seeds <= 0) return 0;
    const uint64_t np = (uint64_t)(z + 2) * (y + 2) * (x + 2);
    return post_bytes((uint64_t)z * y * x, np, max_seeds, np);
}

extern "C" size_t isg_post_workspace_bytes_capped(int64_t z, int64_t y, int64_t x, int64_t max_seeds,
                                                  int64_t max_flood_nodes) {
    if (z <= 0 || y <= 0 || x <= 0 || max_seeds <= 0 || max_flood_nodes <= 0) return 0;
    const uint64_t np = (uint64_t)(z + 2) * (y + 2) * (x + 2);
    return post_bytes((uint64_t)z * y * x, np, max_seeds, (uint64_t)max_flood_nodes);
}

// the largest compact-arena capacity (voxels of multi-seed components) that fits `bytes`
static uint64_t post_node_cap_for(uint64_t n, uint64_t np, int64_t max_seeds, size_t bytes) {
    if (post_bytes(n, np, max_seeds, np) <= bytes) return np;
    const size_t lo_b = post_bytes(n, np, max_seeds, 1);
    if (lo_b > bytes) return 0;
    uint64_t lo = 1, hi = np;                  // post_bytes is monotone in the cap
    while (hi - lo > 1) {
        const uint64_t mid = lo + (hi - lo) / 2;
        if (post_bytes(n, np, max_seeds, mid) <= bytes) lo = mid; else hi = mid;
    }
    return lo;
}

extern "C" int isg_segment_features(const float *feats, int n_chan, int64_t z, int64_t y, int64_t x,
                                    const isg_post_params *prm, const double *gauss1_host,
                                    const double *gauss2_host, uint32_t *labels, uint8_t *mask_out,
                                    int64_t *seeds_out, int64_t max_seeds, int64_t *counts_out,
                                    float *otsu_out, void *workspace, size_t workspace_bytes,
                                    void *stream) {
    ISG_REQUIRE(feats && prm && labels && mask_out && seeds_out && counts_out && otsu_out,
                ISG_ERR_ARG, "isg_segment_features: null pointer");
    ISG_REQUIRE(z >= 1 && y >= 1 && x >= 1 && max_seeds >= 1, ISG_ERR_ARG, "bad extents");
    ISG_REQUIRE(prm->r1 >= 0 && prm->r1 <= 11 && prm->r2 >= 0 && prm->r2 <= 11, ISG_ERR_ARG,
                "gaussian radius must be <= 11");
    for (int i = 0; i < 3; ++i)
        ISG_REQUIRE(prm->aff_ch[i] >= 0 && prm->aff_ch[i] < n_chan, ISG_ERR_ARG, "bad affinity channel");
    ISG_REQUIRE(prm->mask_ch >= 0 && prm->mask_ch < n_chan && prm->cent_ch >= 0 && prm->cent_ch < n_chan,
                ISG_ERR_ARG, "bad channel index");
    const uint64_t n = (uint64_t)z * y * x;
    const uint64_t np = (uint64_t)(z + 2) * (y + 2) * (x + 2);
    ISG_REQUIRE(np < 0xFFFFFFF0ull, ISG_ERR_OVERFLOW, "volume too large for 32-bit voxel ids");
    // the flood reads the affinity planes through one base pointer + stride
    const int c0 = prm->aff_ch[0];
    ISG_REQUIRE(prm->aff_ch[1] - c0 == prm->aff_ch[2] - prm->aff_ch[1], ISG_ERR_ARG,
                "affinity channels must be equally spaced");
    cudaStream_t st = (cudaStream_t)stream;
    // the compact arenas of the ordered flood take what the workspace offers beyond the fixed part:
    // the full isg_post_workspace_bytes covers every voxel, a smaller block (sparse masks, big slabs)
    // covers fewer and fails loudly when the multi-seed components do not fit
    const uint64_t node_cap = post_node_cap_for(n, np, max_seeds, workspace ? workspace_bytes : 0);
    ISG_REQUIRE(workspace && node_cap > 0, ISG_ERR_WORKSPACE,
                "isg_segment_features: workspace too small (%zu < %zu)", workspace_bytes,
                post_bytes(n, np, max_seeds, 1));
    Carver cv(workspace, workspace_bytes);
    PostBuffers b;
    post_carve(&b, cv, n, np, max_seeds, node_cap);
    ISG_REQUIRE(cv.ok, ISG_ERR_WORKSPACE,
                "isg_segment_features: workspace too small (%zu < %zu)", workspace_bytes, cv.off);
    const uint32_t Z = (uint32_t)z, Y = (uint32_t)y, X = (uint32_t)x;
    const int sms = num_sms();
    const int grid = sms * 8;
    GaussW g1, g2;
    g1.r = prm->r1;
    g2.r = prm->r2;
    for (int i = 0; i <= prm->r1; ++i) g1.w[i] = gauss1_host[i];
    for (int i = 0; i <= prm->r2; ++i) g2.w[i] = gauss2_host[i];

    ISG_CUDA(cudaMemsetAsync(b.scal, 0, 64 * sizeof(uint32_t), st));
    ISG_CUDA(cudaMemsetAsync(b.hist, 0, 256 * sizeof(unsigned long long), st));
    {   // scal[0] = min accumulates downwards
        const uint32_t init = 0xFFFFFFFFu;
        ISG_CUDA(cudaMemcpyAsync(b.scal, &init, sizeof(uint32_t), cudaMemcpyHostToDevice, st));
    }
    // ---- affinity channel maxima -------------------------------------------------
    if (prm->use_aff_div) {
        ISG_CUDA(cudaMemcpyAsync(b.fscal, prm->aff_div, 3 * sizeof(float), cudaMemcpyHostToDevice, st));
    } else {
        chan_max_kernel<<<dim3(sms * 2, 3), 256, 0, st>>>(feats, n, n, prm->aff_ch[0], prm->aff_ch[1],
                                                         prm->aff_ch[2], b.scal + 8);
        ISG_LAUNCHED();
        ord_to_float_kernel<<<1, 32, 0, st>>>(b.scal + 8, b.fscal, 3);
        ISG_LAUNCHED();
    }
    // ---- seeds --------------------------------------------------------------------
    const float *cent = feats + (uint64_t)prm->cent_ch * n;
    const float *smoothed_c = cent;
    if (prm->r1 > 0) {
        int grc = gauss_axis(cent, b.tmp_a, Z, Y, X, 1, g1, 0, nullptr, 0, 0, st);
        if (grc) return grc;
        grc = gauss_axis(b.tmp_a, b.tmp_b, Z, Y, X, 2, g1, 0, nullptr, 0, 0, st);
        if (grc) return grc;
        smoothed_c = b.tmp_b;
    }
    local_max_kernel<<<grid, 256, 0, st>>>(smoothed_c, Z, Y, X, prm->peak_thresh, b.cand_a,
                                           (uint32_t)max_seeds, b.scal + 2, b.scal + 3);
    ISG_LAUNCHED();
    uint32_t n_cand = 0;
    ISG_CUDA(cudaMemcpyAsync(&n_cand, b.scal + 2, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    ISG_CUDA(cudaStreamSynchronize(st));
    ISG_REQUIRE(n_cand <= (uint64_t)max_seeds, ISG_ERR_OVERFLOW,
                "isg_segment_features: %u seed candidates exceed max_seeds=%lld", n_cand,
                (long long)max_seeds);
    const uint64_t *cand_sorted = b.cand_a;
    if (n_cand > 1) {
        size_t cb = b.cub_bytes;
        ISG_CUDA(cub::DeviceRadixSort::SortKeys(b.cub_tmp, cb, b.cand_a, b.cand_b, (int)n_cand, 0, 64, st));
        count_launch(4);
        cand_sorted = b.cand_b;
    }
    // ---- mask ---------------------------------------------------------------------
    const float *mraw = feats + (uint64_t)prm->mask_ch * n;
    if (!prm->use_absolute_thresh) {
        const float *s = mraw;
        if (prm->r2 > 0) {
            int grc = gauss_axis(mraw, b.tmp_a, Z, Y, X, 0, g2, 0, nullptr, 0, 0, st);
            if (grc) return grc;
            grc = gauss_axis(b.tmp_a, b.tmp_b, Z, Y, X, 1, g2, 0, nullptr, 0, 0, st);
            if (grc) return grc;
            grc = gauss_axis(b.tmp_b, b.tmp_a, Z, Y, X, 2, g2, 0, b.scal, 0, 0xFFFFFFFFu, st);
            if (grc) return grc;
            s = b.tmp_a;
        } else {
            // sigma = 0: min/max of the raw channel via an identity pass
            GaussW id;
            id.r = 0;
            id.w[0] = 1.0;
            int grc = gauss_axis(mraw, b.tmp_a, Z, Y, X, 2, id, 0, b.scal, 0, 0xFFFFFFFFu, st);
            if (grc) return grc;
            s = b.tmp_a;
        }
        hist_edges_kernel<<<1, 288, 0, st>>>(b.scal, b.edges);
        ISG_LAUNCHED();
        hist_kernel<<<sms * 4, 256, 0, st>>>(s, n, b.scal, b.edges, b.hist);
        ISG_LAUNCHED();
        otsu_kernel<<<1, 32, 0, st>>>(b.hist, b.edges, b.scal, b.fscal + 3);
        ISG_LAUNCHED();
    }
    if (prm->use_absolute_thresh)
        ISG_CUDA(cudaMemcpyAsync(b.fscal + 3, &prm->absolute_thresh, sizeof(float),
                                 cudaMemcpyHostToDevice, st));
    mask_kernel<<<grid, 256, 0, st>>>(mraw, Z, Y, X, b.fscal + 3, prm->absolute_thresh,
                                      prm->use_absolute_thresh, b.mask0);
    ISG_LAUNCHED();
    // ---- components + size window ---------------------------------------------------
    ISG_CUDA(cudaMemsetAsync(b.comp_size, 0, np * sizeof(uint32_t), st));
    ISG_CUDA(cudaMemsetAsync(b.comp_label, 0, np * sizeof(uint32_t), st));
    {
        int rc = ccl_run(b.mask0, b.parent, b.comp_size, Z + 2, Y + 2, X + 2, st);
        if (rc != ISG_OK) return rc;
    }
    mask_keep_kernel<<<grid, 256, 0, st>>>(b.mask0, b.parent, b.comp_size, (uint64_t)prm->min_area,
                                           (uint64_t)prm->max_area, mask_out, np, b.scal + 5);
    ISG_LAUNCHED();
    seed_filter_kernel<<<1, 1024, 0, st>>>(cand_sorted, n_cand, b.scal + 3, b.parent, b.comp_size,
                                           (uint64_t)prm->min_area, (uint64_t)prm->max_area, Z, Y, X,
                                           seeds_out, (uint32_t)max_seeds, labels, b.scal + 4,
                                           prm->seed_keys_out);
    ISG_LAUNCHED();
    if (prm->open_faces) {
        // slab mode: is every component that reaches the own planes completely inside the slab?
        const uint32_t oz0 = (uint32_t)prm->own_z0, oz1 = prm->own_z1 > 0 ? (uint32_t)prm->own_z1 : Z;
        ISG_REQUIRE(oz0 < oz1 && oz1 <= Z, ISG_ERR_ARG, "bad own plane range");
        uint32_t *flag = b.flood.lidmap;          // scratch until the flood stage writes it
        ISG_CUDA(cudaMemsetAsync(flag, 0, np * sizeof(uint32_t), st));
        slab_face_kernel<<<grid, 256, 0, st>>>(b.parent, flag, Z + 2, Y + 2, X + 2, prm->open_faces);
        ISG_LAUNCHED();
        slab_guard_kernel<<<grid, 256, 0, st>>>(b.parent, flag, Y + 2, X + 2, oz0, oz1, b.scal + 6);
        ISG_LAUNCHED();
    }
    // ---- flood ----------------------------------------------------------------------
    FloodGeom g;
    g.aff = feats + (uint64_t)c0 * n;
    g.plane_stride = (int64_t)(prm->aff_ch[1] - c0) * (int64_t)n;
    g.origin = 1;
    g.za = Z;
    g.ya = Y;
    g.xa = X;
    g.div = b.fscal;
    for (int a = 0; a < 3; ++a) g.scale[a] = fabsf(prm->scale[a]);
    g.zp = Z + 2;
    g.yp = Y + 2;
    g.xp = X + 2;
    g.node_key = nullptr;
    int rc = flood_stage_run(b.flood, g, mask_out, b.parent, b.comp_size, b.comp_label, seeds_out,
                             (int64_t)n_cand, b.scal + 4, labels, st);
    if (rc != ISG_OK) return rc;
    counts_kernel<<<1, 1, 0, st>>>(b.scal + 4, b.scal + 2, b.scal + 5, b.flood.scalars + 1, b.scal + 6,
                                   counts_out);
    ISG_LAUNCHED();
    ISG_CUDA(cudaMemcpyAsync(otsu_out, b.fscal + 3, sizeof(float), cudaMemcpyDeviceToDevice, st));
    return ISG_OK;
}

// ---------------------------------------------------------------------------
// slab mode: volume-wide statistics, slab by slab
// ---------------------------------------------------------------------------
namespace isg {
__global__ void float_to_ord_kernel(const float *__restrict__ in, uint32_t *__restrict__ out, int n) {
    int i = threadIdx.x;
    if (i < n) out[i] = f32_ord(in[i]);
}
__global__ void __launch_bounds__(256)
lut_kernel(const unsigned long long *__restrict__ local_keys, int64_t n_local,
           const unsigned long long *__restrict__ sorted, int64_t n_global, uint32_t *__restrict__ lut,
           int *__restrict__ missing) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_local) return;
    const unsigned long long k = local_keys[i];
    int64_t lo = 0, hi = n_global;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (sorted[mid] < k) lo = mid + 1; else hi = mid;
    }
    if (lo < n_global && sorted[lo] == k) lut[i] = (uint32_t)(lo + 1);
    else { lut[i] = 0u; *missing = 1; }
}
__global__ void __launch_bounds__(256)
relabel_kernel(uint32_t *__restrict__ labels, int64_t n, const uint32_t *__restrict__ lut, int64_t n_local,
               int *__restrict__ missing) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint32_t l = labels[i];
        if (l == 0) continue;
        if ((int64_t)l > n_local) { labels[i] = 0; *missing = 1; continue; }
        labels[i] = lut[l - 1];
    }
}
}  // namespace isg

extern "C" int isg_slab_stats(const float *feats, int n_chan, int64_t z, int64_t y, int64_t x,
                              const isg_post_params *prm, const double *gauss2_host, int stage,
                              float *minmax_io, float *chan_max_out, unsigned long long *hist_out,
                              void *workspace, size_t workspace_bytes, void *stream) {
    ISG_REQUIRE(feats && prm && minmax_io, ISG_ERR_ARG, "isg_slab_stats: null pointer");
    ISG_REQUIRE(z >= 1 && y >= 1 && x >= 1 && prm->r2 >= 0 && prm->r2 <= 11, ISG_ERR_ARG, "bad extents");
    ISG_REQUIRE(stage == 0 || stage == 1, ISG_ERR_ARG, "stage must be 0 or 1");
    ISG_REQUIRE(prm->mask_ch >= 0 && prm->mask_ch < n_chan, ISG_ERR_ARG, "bad channel index");
    const uint64_t n = (uint64_t)z * y * x;
    const uint32_t Z = (uint32_t)z, Y = (uint32_t)y, X = (uint32_t)x;
    const uint32_t oz0 = (uint32_t)prm->own_z0, oz1 = prm->own_z1 > 0 ? (uint32_t)prm->own_z1 : Z;
    ISG_REQUIRE(oz0 < oz1 && oz1 <= Z, ISG_ERR_ARG, "bad own plane range");
    cudaStream_t st = (cudaStream_t)stream;
    Carver cv(workspace, workspace_bytes);
    float *tmp_a = cv.take<float>(n), *tmp_b = cv.take<float>(n);
    uint32_t *scal = cv.take<uint32_t>(64);
    float *edges = cv.take<float>(320);
    ISG_REQUIRE(workspace && cv.ok, ISG_ERR_WORKSPACE, "isg_slab_stats: workspace too small (%zu < %zu)",
                workspace_bytes, cv.off);
    const int sms = num_sms();
    GaussW g2;
    g2.r = prm->r2;
    for (int i = 0; i <= prm->r2; ++i) g2.w[i] = gauss2_host[i];
    const uint64_t plane = (uint64_t)Y * X;
    const float *mraw = feats + (uint64_t)prm->mask_ch * n;
    ISG_CUDA(cudaMemsetAsync(scal, 0, 64 * sizeof(uint32_t), st));
    {
        const uint32_t init = 0xFFFFFFFFu;
        ISG_CUDA(cudaMemcpyAsync(scal, &init, sizeof(uint32_t), cudaMemcpyHostToDevice, st));
    }
    uint32_t *mm = stage == 0 ? scal : nullptr;
    const float *s = tmp_a;
    if (prm->r2 > 0) {
        int grc = gauss_axis(mraw, tmp_a, Z, Y, X, 0, g2, 0, nullptr, 0, 0, st);
        if (grc) return grc;
        grc = gauss_axis(tmp_a, tmp_b, Z, Y, X, 1, g2, 0, nullptr, 0, 0, st);
        if (grc) return grc;
        grc = gauss_axis(tmp_b, tmp_a, Z, Y, X, 2, g2, 0, mm, oz0, oz1, st);
        if (grc) return grc;
    } else {
        GaussW id;
        id.r = 0;
        id.w[0] = 1.0;
        int grc = gauss_axis(mraw, tmp_a, Z, Y, X, 2, id, 0, mm, oz0, oz1, st);
        if (grc) return grc;
    }
    if (stage == 0) {
        ISG_REQUIRE(chan_max_out, ISG_ERR_ARG, "isg_slab_stats: chan_max_out is NULL");
        for (int i = 0; i < 3; ++i)
            ISG_REQUIRE(prm->aff_ch[i] >= 0 && prm->aff_ch[i] < n_chan, ISG_ERR_ARG, "bad affinity channel");
        ord_to_float_kernel<<<1, 32, 0, st>>>(scal, minmax_io, 2);
        ISG_LAUNCHED();
        chan_max_kernel<<<dim3(sms * 2, 3), 256, 0, st>>>(feats + oz0 * plane, n, (uint64_t)(oz1 - oz0) * plane,
                                                         prm->aff_ch[0], prm->aff_ch[1], prm->aff_ch[2],
                                                         scal + 8);
        ISG_LAUNCHED();
        ord_to_float_kernel<<<1, 32, 0, st>>>(scal + 8, chan_max_out, 3);
        ISG_LAUNCHED();
        return ISG_OK;
    }
    ISG_REQUIRE(hist_out, ISG_ERR_ARG, "isg_slab_stats: hist_out is NULL");
    float_to_ord_kernel<<<1, 32, 0, st>>>(minmax_io, scal, 2);
    ISG_LAUNCHED();
    ISG_CUDA(cudaMemsetAsync(hist_out, 0, 256 * sizeof(unsigned long long), st));
    hist_edges_kernel<<<1, 288, 0, st>>>(scal, edges);
    ISG_LAUNCHED();
    hist_kernel<<<sms * 4, 256, 0, st>>>(s + oz0 * plane, (uint64_t)(oz1 - oz0) * plane, scal, edges, hist_out);
    ISG_LAUNCHED();
    return ISG_OK;
}

extern "C" int isg_otsu_from_hist(const unsigned long long *hist, const float *minmax, float *thr_out,
                                  void *scratch, size_t scratch_bytes, void *stream) {
    ISG_REQUIRE(hist && minmax && thr_out && scratch, ISG_ERR_ARG, "isg_otsu_from_hist: null pointer");
    ISG_REQUIRE(scratch_bytes >= 2048, ISG_ERR_WORKSPACE, "isg_otsu_from_hist: scratch must hold 2048 bytes");
    cudaStream_t st = (cudaStream_t)stream;
    uint32_t *scal = reinterpret_cast<uint32_t *>(scratch);             // ordered min / max
    float *edges = reinterpret_cast<float *>(scratch) + 64;             // 257 float32 bin edges
    float_to_ord_kernel<<<1, 32, 0, st>>>(minmax, scal, 2);
    ISG_LAUNCHED();
    hist_edges_kernel<<<1, 288, 0, st>>>(scal, edges);
    ISG_LAUNCHED();
    otsu_kernel<<<1, 32, 0, st>>>(hist, edges, scal, thr_out);
    ISG_LAUNCHED();
    return ISG_OK;
}

extern "C" size_t isg_sort_tmp_bytes(int64_t n) {
    if (n < 1) n = 1;
    size_t b = 0;
    cub::DeviceRadixSort::SortKeys(nullptr, b, (unsigned long long *)nullptr, (unsigned long long *)nullptr, (int)n);
    return b + 256 + (size_t)n * sizeof(unsigned long long);
}

extern "C" int isg_sort_keys_u64(unsigned long long *keys, int64_t n, void *tmp, size_t tmp_bytes, void *stream) {
    ISG_REQUIRE(n >= 0 && (n == 0 || (keys && tmp)), ISG_ERR_ARG, "isg_sort_keys_u64: null pointer");
    if (n <= 1) return ISG_OK;
    ISG_REQUIRE(tmp_bytes >= isg_sort_tmp_bytes(n), ISG_ERR_WORKSPACE, "isg_sort_keys_u64: tmp too small");
    cudaStream_t st = (cudaStream_t)stream;
    unsigned long long *alt = (unsigned long long *)tmp;
    size_t cb = tmp_bytes - (size_t)n * sizeof(unsigned long long);
    char *cub_tmp = (char *)tmp + (((size_t)n * sizeof(unsigned long long) + 255) & ~(size_t)255);
    cb -= (size_t)(cub_tmp - ((char *)tmp + (size_t)n * sizeof(unsigned long long)));
    ISG_CUDA(cub::DeviceRadixSort::SortKeys(cub_tmp, cb, keys, alt, (int)n, 0, 64, st));
    count_launch(4);
    ISG_CUDA(cudaMemcpyAsync(keys, alt, (size_t)n * sizeof(unsigned long long), cudaMemcpyDeviceToDevice, st));
    return ISG_OK;
}

extern "C" int isg_relabel_by_keys(uint32_t *labels, int64_t n, const unsigned long long *local_keys,
                                   int64_t n_local, const unsigned long long *global_sorted_keys,
                                   int64_t n_global, uint32_t *lut_scratch, int *missing_out, void *stream) {
    ISG_REQUIRE(labels && missing_out && n >= 0 && n_local >= 0 && n_global >= 0, ISG_ERR_ARG,
                "isg_relabel_by_keys: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    ISG_CUDA(cudaMemsetAsync(missing_out, 0, sizeof(int), st));
    if (n_local > 0) {
        ISG_REQUIRE(local_keys && global_sorted_keys && lut_scratch, ISG_ERR_ARG, "isg_relabel_by_keys: null pointer");
        lut_kernel<<<(int)((n_local + 255) / 256), 256, 0, st>>>(local_keys, n_local, global_sorted_keys, n_global,
                                                                  lut_scratch, missing_out);
        ISG_LAUNCHED();
    }
    if (n > 0) {
        relabel_kernel<<<num_sms() * 8, 256, 0, st>>>(labels, n, lut_scratch, n_local, missing_out);
        ISG_LAUNCHED();
    }
    return ISG_OK;
}
