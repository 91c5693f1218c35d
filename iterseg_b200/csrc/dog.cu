// DoG blob segmenter on the device: isg_dog_blob_segment.
//
// Replaces dog_blob_watershed_for_chunks + dog_image (src/iterseg/segmentation.py:592-650,
// :678-680):
//   input_volume = np.pad(input_volume, 1)                                       :634
//   mask    = gaussian(v, min_sigma) - gaussian(v, max_sigma) > threshold        :635-636, :678-680
//   markers = blob_dog(v, min_sigma, max_sigma, threshold)                        :637-638
//   distance = ndi.distance_transform_edt(v)                                      :639
//   markers, n = ndi.label(centroids at the blob coordinates)                     :641-644
//   labels  = watershed(-distance, markers, mask=mask)                            :645
// The three scikit-image functions are restated from their published algorithms (scikit-image
// is not available in the build environment; parity unpinned, see oracle/dog.py for the
// exact statements and the two documented tie rules).  Arithmetic contracts as in post.cu:
// every 1-D Gaussian pass accumulates in float64 in scipy's order and stores float32.
//   * Gaussians: separable, 'nearest' for the mask, 'reflect' for blob_dog;
//   * blob_dog with one DoG layer (k = int(log(max/min)/log(1.6) + 1) == 1, i.e.
//     max_sigma / min_sigma < 1.6): 3x3x3 maxima of (G(s0) - G(1.6 s0)) / 0.6 above the
//     threshold, no border exclusion, a constant cube has none; _prune_blobs with one common
//     sigma reduces to "a blob dies iff a LATER blob (peak order) lies within the overlap
//     distance" (pairs walked in lexicographic order), which is data parallel;
//   * exact Euclidean distance transform: separable squared-distance passes (x scan, then z
//     and y lower envelopes by bounded search); the flood is keyed by ~d2 (uint32), which orders
//     voxels exactly like float64 -sqrt(d2);
//   * ndi.label numbering: 6-connected components of the marker voxels, ids in raster order of
//     each component's first voxel;
//   * the marker watershed: flood_stage_run in node-keyed mode (flood.cuh).
#include <cub/cub.cuh>

#include "flood_stage.h"
#include "gauss.cuh"

namespace isg {

__global__ void __launch_bounds__(256)
pad_kernel(const float *__restrict__ vol, float *__restrict__ vp, uint32_t Z, uint32_t Y, uint32_t X) {
    const uint32_t Yp = Y + 2, Xp = X + 2;
    const uint64_t np = (uint64_t)(Z + 2) * Yp * Xp;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; v < np; v += stride) {
        const uint32_t x = (uint32_t)(v % Xp);
        const uint64_t t = v / Xp;
        const uint32_t y = (uint32_t)(t % Yp), z = (uint32_t)(t / Yp);
        float o = 0.0f;
        if (x >= 1 && x <= X && y >= 1 && y <= Y && z >= 1 && z <= Z)
            o = __ldg(vol + ((uint64_t)(z - 1) * Y + (y - 1)) * X + (x - 1));
        vp[v] = o;
    }
}

// mask = a - b > thr;  cube = (c - d) * sf (float32 arithmetic, as numpy evaluates it)
__global__ void __launch_bounds__(256)
dog_combine_kernel(const float *__restrict__ a, const float *__restrict__ b, float thr,
                   uint8_t *__restrict__ mask, uint64_t n) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; v < n; v += stride)
        mask[v] = __fsub_rn(a[v], b[v]) > thr ? 1 : 0;
}
__global__ void __launch_bounds__(256)
dog_cube_kernel(const float *__restrict__ c, const float *__restrict__ d, float sf, float *__restrict__ cube,
                uint64_t n) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; v < n; v += stride)
        cube[v] = __fmul_rn(__fsub_rn(c[v], d[v]), sf);
}

// 3x3x3 maxima ('nearest' = clamped window), value > thr, no border exclusion
__global__ void __launch_bounds__(256)
dog_peak_kernel(const float *__restrict__ cs, uint32_t Z, uint32_t Y, uint32_t X, float thr,
                uint64_t *__restrict__ cand, uint32_t cap, uint32_t *__restrict__ n_cand,
                uint32_t *__restrict__ nontrivial) {
    const uint64_t n = (uint64_t)Z * Y * X;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    bool saw_nonmax = false;
    for (uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; v < n; v += stride) {
        const uint32_t x = (uint32_t)(v % X);
        const uint64_t t = v / X;
        const uint32_t y = (uint32_t)(t % Y), z = (uint32_t)(t / Y);
        const float c = cs[v];
        bool is_max = true;
        const uint32_t z0 = z ? z - 1 : 0, z1 = z + 1 < Z ? z + 1 : Z - 1;
        const uint32_t y0 = y ? y - 1 : 0, y1 = y + 1 < Y ? y + 1 : Y - 1;
        const uint32_t x0 = x ? x - 1 : 0, x1 = x + 1 < X ? x + 1 : X - 1;
        for (uint32_t zz = z0; zz <= z1; ++zz)
            for (uint32_t yy = y0; yy <= y1; ++yy) {
                const float *row = cs + ((uint64_t)zz * Y + yy) * X;
                for (uint32_t xx = x0; xx <= x1; ++xx) is_max &= !(__ldg(row + xx) > c);
            }
        if (!is_max) saw_nonmax = true;
        if (is_max && c > thr) {
            const uint32_t slot = atomicAdd(n_cand, 1u);
            if (slot < cap) cand[slot] = ((uint64_t)(~f32_ord(c)) << 32) | (uint64_t)v;
        }
    }
    if (__any_sync(0xFFFFFFFFu, saw_nonmax) && (threadIdx.x & 31) == 0) atomicOr(nontrivial, 1u);
}

// multi-layer: 3^4 maxima over (z, y, x, layer) of the cube [layer][z][y][x] ('nearest' = clamped window
// on all four axes), value > thr; candidate key = (~ord(value) << 32) | (voxel * K + layer), i.e. ties in
// argwhere order of the (z, y, x, layer) array; max_layer = the largest layer index that holds a peak
__global__ void __launch_bounds__(256)
dog_peak4_kernel(const float *__restrict__ cube, int K, uint32_t Z, uint32_t Y, uint32_t X, float thr,
                 uint64_t *__restrict__ cand, uint32_t cap, uint32_t *__restrict__ n_cand,
                 uint32_t *__restrict__ nontrivial, uint32_t *__restrict__ max_layer) {
    const uint64_t n = (uint64_t)Z * Y * X;
    const uint64_t total = n * (uint64_t)K;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    bool saw_nonmax = false;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int k = (int)(i / n);
        const uint64_t v = i - (uint64_t)k * n;
        const uint32_t x = (uint32_t)(v % X);
        const uint64_t t = v / X;
        const uint32_t y = (uint32_t)(t % Y), z = (uint32_t)(t / Y);
        const float c = cube[i];
        bool is_max = true;
        const int k0 = k ? k - 1 : 0, k1 = k + 1 < K ? k + 1 : K - 1;
        const uint32_t z0 = z ? z - 1 : 0, z1 = z + 1 < Z ? z + 1 : Z - 1;
        const uint32_t y0 = y ? y - 1 : 0, y1 = y + 1 < Y ? y + 1 : Y - 1;
        const uint32_t x0 = x ? x - 1 : 0, x1 = x + 1 < X ? x + 1 : X - 1;
        if (c > thr || !saw_nonmax) {                       // below the threshold only the "constant cube" test needs it
            for (int kk = k0; kk <= k1; ++kk)
                for (uint32_t zz = z0; zz <= z1; ++zz)
                    for (uint32_t yy = y0; yy <= y1; ++yy) {
                        const float *row = cube + (uint64_t)kk * n + ((uint64_t)zz * Y + yy) * X;
                        for (uint32_t xx = x0; xx <= x1; ++xx) is_max &= !(__ldg(row + xx) > c);
                    }
            if (!is_max) saw_nonmax = true;
            if (is_max && c > thr) {
                const uint32_t slot = atomicAdd(n_cand, 1u);
                if (slot < cap) cand[slot] = ((uint64_t)(~f32_ord(c)) << 32) | (v * (uint64_t)K + (uint64_t)k);
                atomicMax(max_layer, (uint32_t)k);
            }
        }
    }
    if (__any_sync(0xFFFFFFFFu, saw_nonmax) && (threadIdx.x & 31) == 0) atomicOr(nontrivial, 1u);
}

// _prune_blobs with per-blob sigma, one CTA.  Blobs in peak order; a pair (i, j), i < j, is "close" when
// sqrt(d2) <= dist; pairs are walked in lexicographic order and use the CURRENT sigmas (a dead blob has
// sigma 0, overlaps nothing it could kill and is skipped).  For a live blob i all its pairs are
// evaluated in parallel: j* = the first close live j that kills i (overlap > thr and sigma_i <= sigma_j);
// every close live j < j* (all of them when there is no j*) with overlap > thr and sigma_i > sigma_j dies.
// O(N^2 / 1024) distance tests: milliseconds for the thousands of blobs of a frame.
__device__ __forceinline__ double blob_overlap_frac(double s1, double s2, double d2) {
    const double root3 = sqrt(3.0);
    double r1 = __dmul_rn(s1, root3), r2 = __dmul_rn(s2, root3);
    if (r2 > r1) { const double t = r1; r1 = r2; r2 = t; }
    const double d = sqrt(d2);
    if (d > __dadd_rn(r1, r2)) return 0.0;
    if (d <= fabs(__dsub_rn(r1, r2))) return 1.0;
    const double pi = 3.141592653589793;
    const double a = __dsub_rn(__dadd_rn(r1, r2), d);
    const double rm = __dsub_rn(r1, r2);
    double poly = __dadd_rn(__dmul_rn(d, d), __dmul_rn(__dmul_rn(2.0, d), __dadd_rn(r1, r2)));
    poly = __dsub_rn(poly, __dmul_rn(3.0, __dmul_rn(rm, rm)));
    const double vol = __dmul_rn(__dmul_rn(__ddiv_rn(pi, __dmul_rn(12.0, d)), __dmul_rn(a, a)), poly);
    const double mn = r1 < r2 ? r1 : r2;
    const double full = __dmul_rn(__dmul_rn(__ddiv_rn(4.0, 3.0), pi), pow(mn, 3.0));
    return __ddiv_rn(vol, full);
}
__global__ void __launch_bounds__(1024)
blob_prune_multi_kernel(const uint64_t *__restrict__ cand_sorted, uint32_t n, int K,
                        const uint32_t *__restrict__ nontrivial, uint32_t Y, uint32_t X,
                        const double *__restrict__ sigmas, double dist, double thr,
                        float *__restrict__ sig /* [n] scratch */, uint8_t *__restrict__ centroids,
                        uint32_t *__restrict__ n_blobs) {
    if (*nontrivial == 0) return;
    __shared__ uint32_t jstar;
    const uint64_t plane = (uint64_t)Y * X;
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
        const uint64_t id = cand_sorted[i] & 0xFFFFFFFFull;
        sig[i] = (float)(int)(id % (uint64_t)K + 1);          // layer + 1 while alive, 0 when dead
    }
    __syncthreads();
    for (uint32_t i = 0; i + 1 < n; ++i) {
        if (sig[i] == 0.0f) continue;                          // uniform: written before the last barrier
        if (threadIdx.x == 0) jstar = 0xFFFFFFFFu;
        __syncthreads();
        const uint64_t idi = cand_sorted[i] & 0xFFFFFFFFull;
        const uint64_t vi = idi / (uint64_t)K;
        const int zi = (int)(vi / plane), yi = (int)((vi % plane) / X), xi = (int)(vi % X);
        const double si = sigmas[(int)sig[i] - 1];
        uint32_t mine[4];                                      // this thread's close live js that i may kill
        int n_mine = 0;
        bool overflow = false;
        for (uint32_t j = i + 1 + threadIdx.x; j < n; j += blockDim.x) {
            const float sj_f = sig[j];
            if (sj_f == 0.0f) continue;
            const uint64_t vj = (cand_sorted[j] & 0xFFFFFFFFull) / (uint64_t)K;
            const int dz = (int)(vj / plane) - zi, dy = (int)((vj % plane) / X) - yi, dx = (int)(vj % X) - xi;
            const double d2 = (double)(dz * dz + dy * dy + dx * dx);
            if (sqrt(d2) > dist) continue;
            const double sj = sigmas[(int)sj_f - 1];
            if (!(blob_overlap_frac(si, sj, d2) > thr)) continue;
            if (si > sj) {
                if (n_mine < 4) mine[n_mine++] = j; else overflow = true;
            } else {
                atomicMin(&jstar, j);
            }
        }
        __syncthreads();
        const uint32_t js = jstar;
        for (int q = 0; q < n_mine; ++q)
            if (mine[q] < js) sig[mine[q]] = 0.0f;
        if (overflow) {                                        // more than 4 victims in one thread's stride: redo them
            for (uint32_t j = i + 1 + threadIdx.x; j < n && j < js; j += blockDim.x) {
                const float sj_f = sig[j];
                if (sj_f == 0.0f) continue;
                const uint64_t vj = (cand_sorted[j] & 0xFFFFFFFFull) / (uint64_t)K;
                const int dz = (int)(vj / plane) - zi, dy = (int)((vj % plane) / X) - yi, dx = (int)(vj % X) - xi;
                const double d2 = (double)(dz * dz + dy * dy + dx * dx);
                if (sqrt(d2) > dist) continue;
                const double sj = sigmas[(int)sj_f - 1];
                if (blob_overlap_frac(si, sj, d2) > thr && si > sj) sig[j] = 0.0f;
            }
        }
        if (threadIdx.x == 0 && js != 0xFFFFFFFFu) sig[i] = 0.0f;
        __syncthreads();
    }
    uint32_t alive = 0;
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x)
        if (sig[i] != 0.0f) {
            centroids[(cand_sorted[i] & 0xFFFFFFFFull) / (uint64_t)K] = 1;
            ++alive;
        }
    if (alive) atomicAdd(n_blobs, alive);
}

// blob index grid (rank + 1 in peak order) and the prune rule
__global__ void blob_grid_kernel(const uint64_t *__restrict__ cand_sorted, uint32_t n,
                                 const uint32_t *__restrict__ nontrivial, uint32_t *__restrict__ grid) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || *nontrivial == 0) return;
    grid[(uint32_t)(cand_sorted[i] & 0xFFFFFFFFu)] = i + 1;
}
__global__ void blob_prune_kernel(const uint64_t *__restrict__ cand_sorted, uint32_t n,
                                  const uint32_t *__restrict__ nontrivial, const uint32_t *__restrict__ grid,
                                  uint32_t Z, uint32_t Y, uint32_t X, int rad, int d2max,
                                  uint8_t *__restrict__ centroids, uint32_t *__restrict__ n_blobs) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || *nontrivial == 0) return;
    const uint32_t v = (uint32_t)(cand_sorted[i] & 0xFFFFFFFFu);
    const int x = (int)(v % X), y = (int)((v / X) % Y), z = (int)(v / ((uint64_t)X * Y));
    bool dies = false;
    for (int dz = -rad; dz <= rad && !dies; ++dz)
        for (int dy = -rad; dy <= rad && !dies; ++dy)
            for (int dx = -rad; dx <= rad; ++dx) {
                const int q = dz * dz + dy * dy + dx * dx;
                if (q == 0 || q > d2max) continue;
                const int zz = z + dz, yy = y + dy, xx = x + dx;
                if (zz < 0 || yy < 0 || xx < 0 || zz >= (int)Z || yy >= (int)Y || xx >= (int)X) continue;
                const uint32_t j = grid[((uint64_t)zz * Y + yy) * X + xx];
                if (j > i + 1) { dies = true; break; }       // a later blob overlaps: this one is pruned
            }
    if (!dies) {
        centroids[v] = 1;
        atomicAdd(n_blobs, 1u);
    }
}

// ---- exact squared Euclidean distance to the nearest zero voxel ------------------------------
static constexpr uint32_t EDT_INF = 1u << 28;
// x pass: one thread per row, two sweeps
__global__ void __launch_bounds__(128)
edt_x_kernel(const float *__restrict__ vp, uint32_t *__restrict__ d2, uint64_t rows, uint32_t X) {
    const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    const float *in = vp + r * X;
    uint32_t *out = d2 + r * X;
    uint32_t d = EDT_INF;
    for (uint32_t x = 0; x < X; ++x) {
        d = in[x] == 0.0f ? 0u : (d >= EDT_INF ? EDT_INF : d + 1u);
        out[x] = d;
    }
    d = EDT_INF;
    for (uint32_t x = X; x-- > 0;) {
        d = in[x] == 0.0f ? 0u : (d >= EDT_INF ? EDT_INF : d + 1u);
        const uint32_t f = out[x] < d ? out[x] : d;
        out[x] = f >= EDT_INF ? EDT_INF : f * f;
    }
}
// lower envelope along one axis: out[c] = min_j in[j] + (c - j)^2, searched outwards until (c-j)^2 >= best
template <int AXIS>
__global__ void __launch_bounds__(256)
edt_axis_kernel(const uint32_t *__restrict__ in, uint32_t *__restrict__ out, uint32_t Z, uint32_t Y, uint32_t X) {
    const uint64_t n = (uint64_t)Z * Y * X;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const int len = (int)(AXIS == 0 ? Z : Y);
    const uint64_t step = AXIS == 0 ? (uint64_t)Y * X : (uint64_t)X;
    for (uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; v < n; v += stride) {
        const uint64_t t = v / X;
        const int c = (int)(AXIS == 0 ? t / Y : t % Y);
        const uint64_t line0 = v - (uint64_t)c * step;
        uint32_t best = in[v];
        for (int o = 1; o < len; ++o) {
            const uint32_t oo = (uint32_t)o * (uint32_t)o;
            if (oo >= best) break;
            if (c - o >= 0) {
                const uint32_t a = __ldg(in + line0 + (uint64_t)(c - o) * step);
                if (a < EDT_INF && a + oo < best) best = a + oo;
            }
            if (c + o < len) {
                const uint32_t a = __ldg(in + line0 + (uint64_t)(c + o) * step);
                if (a < EDT_INF && a + oo < best) best = a + oo;
            }
        }
        out[v] = best;
    }
}
__global__ void __launch_bounds__(256)
edt_finish_kernel(const uint32_t *__restrict__ d2, uint32_t *__restrict__ key, double *__restrict__ dist,
                  uint64_t n) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; v < n; v += stride) {
        const uint32_t q = d2[v];
        key[v] = ~q;                                   // -distance ascending == d2 descending
        if (dist) dist[v] = sqrt((double)q);
    }
}

// ---- ndi.label numbering of the marker voxels and the seed list -------------------------------
__global__ void __launch_bounds__(256)
root_flag_kernel(const uint32_t *__restrict__ parent, uint32_t *__restrict__ flag, uint64_t n) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; v < n; v += stride)
        flag[v] = parent[v] == (uint32_t)v ? 1u : 0u;
}
__global__ void __launch_bounds__(256)
marker_kernel(const uint32_t *__restrict__ parent, const uint32_t *__restrict__ rank_excl,
              uint32_t *__restrict__ labels, int64_t *__restrict__ seeds, uint32_t *__restrict__ seed_labels,
              uint32_t cap, uint32_t *__restrict__ n_seeds, uint64_t n) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; v < n; v += stride) {
        const uint32_t r = parent[v];
        if (r == CCL_NONE) continue;
        const uint32_t l = rank_excl[r] + 1u;
        labels[v] = l;
        const uint32_t slot = atomicAdd(n_seeds, 1u);
        if (slot < cap) {
            seeds[slot] = (int64_t)v;
            seed_labels[slot] = l;
        }
    }
}
__global__ void __launch_bounds__(256)
dog_domain_kernel(const uint8_t *__restrict__ mask, const uint8_t *__restrict__ centroids,
                  uint8_t *__restrict__ dom, uint64_t n) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; v < n; v += stride)
        dom[v] = (mask[v] | centroids[v]) ? 1 : 0;
}
__global__ void dog_counts_kernel(const uint32_t *n_cand, const uint32_t *n_blobs, const uint32_t *n_seeds,
                                  const uint32_t *n_markers_src, uint64_t last, int64_t *out) {
    out[0] = *n_cand;
    out[1] = *n_blobs;
    out[2] = *n_seeds;
    out[3] = n_markers_src[last];
}

struct DogBuffers {
    float *cube;                   // multi-layer: [K][np] DoG layers
    float *sig;                    // multi-layer: per-candidate state of the prune sweep
    double *sigmas;                // multi-layer: sigma_list on the device
    float *vp, *ga, *gb, *gc;
    uint32_t *d2a, *d2b, *key, *grid, *flag, *rank;
    uint64_t *cand_a, *cand_b;
    uint8_t *centroids, *dom;
    uint32_t *parent, *comp_size, *comp_label, *seed_labels, *scal;
    int64_t *seeds;
    unsigned char *cub_tmp;
    size_t cub_bytes;
    FloodStageBuffers flood;
};

static void dog_carve(DogBuffers *b, Carver &cv, uint64_t np, int64_t max_seeds, int n_layers) {
    b->cube = n_layers > 1 ? cv.take<float>(np * (uint64_t)n_layers) : nullptr;
    b->sig = n_layers > 1 ? cv.take<float>(max_seeds) : nullptr;
    b->sigmas = n_layers > 1 ? cv.take<double>(ISG_DOG_MAX_SIGMAS) : nullptr;
    b->vp = cv.take<float>(np);
    b->ga = cv.take<float>(np);
    b->gb = cv.take<float>(np);
    b->gc = cv.take<float>(np);
    b->d2a = cv.take<uint32_t>(np);
    b->d2b = cv.take<uint32_t>(np);
    b->key = cv.take<uint32_t>(np);
    b->grid = cv.take<uint32_t>(np);
    b->flag = cv.take<uint32_t>(np);
    b->rank = cv.take<uint32_t>(np + 1);
    b->cand_a = cv.take<uint64_t>(max_seeds);
    b->cand_b = cv.take<uint64_t>(max_seeds);
    b->centroids = cv.take<uint8_t>(np);
    b->dom = cv.take<uint8_t>(np);
    b->parent = cv.take<uint32_t>(np);
    b->comp_size = cv.take<uint32_t>(np);
    b->comp_label = cv.take<uint32_t>(np);
    b->seed_labels = cv.take<uint32_t>(max_seeds);
    b->seeds = cv.take<int64_t>(max_seeds);
    b->scal = cv.take<uint32_t>(64);
    size_t s1 = 0, s2 = 0;
    cub::DeviceRadixSort::SortKeys(nullptr, s1, (uint64_t *)nullptr, (uint64_t *)nullptr, (int)max_seeds);
    cub::DeviceScan::ExclusiveSum(nullptr, s2, (uint32_t *)nullptr, (uint32_t *)nullptr, (int)(np + 1));
    b->cub_bytes = (s1 > s2 ? s1 : s2) + 256;
    b->cub_tmp = cv.take<unsigned char>(b->cub_bytes);
    flood_stage_workspace(&b->flood, cv, np, max_seeds, np);
}

// 3-axis Gaussian in -> out using tmp (in is preserved)
static int gauss3(const float *in, float *tmp, float *out, uint32_t Z, uint32_t Y, uint32_t X, const GaussW &g,
                  int reflect, cudaStream_t st) {
    int rc = gauss_axis(in, out, Z, Y, X, 0, g, reflect, nullptr, 0, 0, st);
    if (rc) return rc;
    rc = gauss_axis(out, tmp, Z, Y, X, 1, g, reflect, nullptr, 0, 0, st);
    if (rc) return rc;
    return gauss_axis(tmp, out, Z, Y, X, 2, g, reflect, nullptr, 0, 0, st);
}

}  // namespace isg

using namespace isg;

extern "C" size_t isg_dog_workspace_bytes(int64_t z, int64_t y, int64_t x, int64_t max_seeds) {
    if (z <= 0 || y <= 0 || x <= 0 || max_seeds <= 0) return 0;
    return isg_dog_workspace_bytes_layers(z, y, x, max_seeds, 1);
}

extern "C" size_t isg_dog_workspace_bytes_layers(int64_t z, int64_t y, int64_t x, int64_t max_seeds, int n_layers) {
    if (z <= 0 || y <= 0 || x <= 0 || max_seeds <= 0 || n_layers >= ISG_DOG_MAX_SIGMAS) return 0;
    Carver cv(nullptr, 0);
    DogBuffers b;
    dog_carve(&b, cv, (uint64_t)(z + 2) * (y + 2) * (x + 2), max_seeds, n_layers);
    return cv.off + 512;
}

extern "C" int isg_dog_blob_segment(const float *vol, int64_t z, int64_t y, int64_t x,
                                    const isg_dog_params *prm, uint32_t *labels, uint8_t *mask_out,
                                    double *distance_out, int64_t max_seeds, int64_t *counts_out,
                                    void *workspace, size_t workspace_bytes, void *stream) {
    ISG_REQUIRE(vol && prm && labels && mask_out && counts_out, ISG_ERR_ARG, "isg_dog_blob_segment: null pointer");
    ISG_REQUIRE(z >= 1 && y >= 1 && x >= 1 && max_seeds >= 1, ISG_ERR_ARG, "bad extents");
    const int K = prm->n_layers > 1 ? prm->n_layers : 1;
    for (int i = 0; i < 4 && K == 1; ++i)
        ISG_REQUIRE(prm->radius[i] >= 0 && prm->radius[i] <= 11, ISG_ERR_ARG, "gaussian radius must be <= 11");
    for (int i = 0; i < 2 && K > 1; ++i)
        ISG_REQUIRE(prm->mask_radius[i] >= 0 && prm->mask_radius[i] <= ISG_GAUSS_MAX_RADIUS, ISG_ERR_ARG,
                    "gaussian radius must be <= %d", ISG_GAUSS_MAX_RADIUS);
    ISG_REQUIRE(K < ISG_DOG_MAX_SIGMAS, ISG_ERR_ARG, "isg_dog_blob_segment: at most %d DoG layers", ISG_DOG_MAX_SIGMAS - 1);
    if (K > 1)
        for (int i = 0; i <= K; ++i)
            ISG_REQUIRE(prm->layer_radius[i] >= 0 && prm->layer_radius[i] <= ISG_GAUSS_MAX_RADIUS && prm->layer_sigma[i] > 0,
                        ISG_ERR_ARG, "isg_dog_blob_segment: layer %d: gaussian radius must be <= %d", i, ISG_GAUSS_MAX_RADIUS);
    const uint32_t Z = (uint32_t)z + 2, Y = (uint32_t)y + 2, X = (uint32_t)x + 2;      // padded extents
    const uint64_t np = (uint64_t)Z * Y * X;
    ISG_REQUIRE(np * (uint64_t)K < 0xFFFFFFF0ull, ISG_ERR_OVERFLOW, "volume too large for 32-bit (voxel, layer) ids");
    cudaStream_t st = (cudaStream_t)stream;
    Carver cv(workspace, workspace_bytes);
    DogBuffers b;
    dog_carve(&b, cv, np, max_seeds, K);
    ISG_REQUIRE(workspace && cv.ok, ISG_ERR_WORKSPACE, "isg_dog_blob_segment: workspace too small (%zu < %zu)",
                workspace_bytes, cv.off);
    const int grid = num_sms() * 8;
    GaussW g[4];
    for (int k = 0; k < 4; ++k) {
        if (K > 1 && k < 2) {
            g[k].r = prm->mask_radius[k];
            for (int i = 0; i <= g[k].r; ++i) g[k].w[i] = prm->mask_weights[k][i];
        } else {
            g[k].r = K > 1 ? 0 : prm->radius[k];
            for (int i = 0; i <= g[k].r; ++i) g[k].w[i] = prm->weights[k][i];
        }
    }
    ISG_CUDA(cudaMemsetAsync(b.scal, 0, 64 * sizeof(uint32_t), st));
    pad_kernel<<<grid, 256, 0, st>>>(vol, b.vp, (uint32_t)z, (uint32_t)y, (uint32_t)x);
    ISG_LAUNCHED();
    int rc;
    // ---- mask = gaussian(v, min) - gaussian(v, max) > threshold ('nearest') ------------------
    if ((rc = gauss3(b.vp, b.gc, b.ga, Z, Y, X, g[0], 0, st))) return rc;
    if ((rc = gauss3(b.vp, b.gc, b.gb, Z, Y, X, g[1], 0, st))) return rc;
    dog_combine_kernel<<<grid, 256, 0, st>>>(b.ga, b.gb, prm->threshold, mask_out, np);
    ISG_LAUNCHED();
    // ---- blob_dog: DoG layers of 'reflect' Gaussians, scaled, local maxima over space (and scale) ------
    if (K == 1) {
        if ((rc = gauss3(b.vp, b.gc, b.ga, Z, Y, X, g[2], 1, st))) return rc;
        if ((rc = gauss3(b.vp, b.gc, b.gb, Z, Y, X, g[3], 1, st))) return rc;
        dog_cube_kernel<<<grid, 256, 0, st>>>(b.ga, b.gb, prm->scale_factor, b.gc, np);
        ISG_LAUNCHED();
        dog_peak_kernel<<<grid, 256, 0, st>>>(b.gc, Z, Y, X, prm->threshold, b.cand_a, (uint32_t)max_seeds,
                                              b.scal + 0, b.scal + 1);
        ISG_LAUNCHED();
    } else {
        float *prev = b.ga, *cur = b.gb;
        for (int i = 0; i <= K; ++i) {
            GaussW gl;
            gl.r = prm->layer_radius[i];
            for (int j = 0; j <= gl.r; ++j) gl.w[j] = prm->layer_weights[i][j];
            if ((rc = gauss3(b.vp, b.gc, cur, Z, Y, X, gl, 1, st))) return rc;
            if (i > 0) {
                dog_cube_kernel<<<grid, 256, 0, st>>>(prev, cur, prm->scale_factor, b.cube + (uint64_t)(i - 1) * np, np);
                ISG_LAUNCHED();
            }
            float *t = prev; prev = cur; cur = t;
        }
        dog_peak4_kernel<<<grid, 256, 0, st>>>(b.cube, K, Z, Y, X, prm->threshold, b.cand_a, (uint32_t)max_seeds,
                                               b.scal + 0, b.scal + 1, b.scal + 4);
        ISG_LAUNCHED();
    }
    uint32_t n_cand = 0, max_layer = 0;
    {
        uint32_t head[5] = {0, 0, 0, 0, 0};
        ISG_CUDA(cudaMemcpyAsync(head, b.scal, sizeof(head), cudaMemcpyDeviceToHost, st));
        ISG_CUDA(cudaStreamSynchronize(st));
        n_cand = head[0];
        max_layer = head[4];
    }
    ISG_REQUIRE(n_cand <= (uint64_t)max_seeds, ISG_ERR_OVERFLOW,
                "isg_dog_blob_segment: %u blob candidates exceed max_seeds=%lld", n_cand, (long long)max_seeds);
    const uint64_t *cand_sorted = b.cand_a;
    if (n_cand > 1) {
        size_t cb = b.cub_bytes;
        ISG_CUDA(cub::DeviceRadixSort::SortKeys(b.cub_tmp, cb, b.cand_a, b.cand_b, (int)n_cand, 0, 64, st));
        count_launch(4);
        cand_sorted = b.cand_b;
    }
    ISG_CUDA(cudaMemsetAsync(b.grid, 0, np * sizeof(uint32_t), st));
    ISG_CUDA(cudaMemsetAsync(b.centroids, 0, np, st));
    if (n_cand > 0 && K == 1) {
        const int blocks = (int)((n_cand + 255) / 256);
        blob_grid_kernel<<<blocks, 256, 0, st>>>(cand_sorted, n_cand, b.scal + 1, b.grid);
        ISG_LAUNCHED();
        blob_prune_kernel<<<blocks, 256, 0, st>>>(cand_sorted, n_cand, b.scal + 1, b.grid, Z, Y, X,
                                                  prm->prune_radius, prm->prune_d2, b.centroids, b.scal + 2);
        ISG_LAUNCHED();
    } else if (n_cand > 0) {
        ISG_CUDA(cudaMemcpyAsync(b.sigmas, prm->layer_sigma, sizeof(double) * ISG_DOG_MAX_SIGMAS,
                                 cudaMemcpyHostToDevice, st));
        // _prune_blobs: distance = 2 * sigma_max * sqrt(ndim), sigma_max over the DETECTED blobs
        const double dist = 2.0 * prm->layer_sigma[max_layer] * sqrt(3.0);
        blob_prune_multi_kernel<<<1, 1024, 0, st>>>(cand_sorted, n_cand, K, b.scal + 1, Y, X, b.sigmas, dist,
                                                    prm->overlap, b.sig, b.centroids, b.scal + 2);
        ISG_LAUNCHED();
    }
    // ---- exact EDT of (v != 0) ----------------------------------------------------------------------
    edt_x_kernel<<<(int)(((uint64_t)Z * Y + 127) / 128), 128, 0, st>>>(b.vp, b.d2a, (uint64_t)Z * Y, X);
    ISG_LAUNCHED();
    edt_axis_kernel<0><<<grid, 256, 0, st>>>(b.d2a, b.d2b, Z, Y, X);
    ISG_LAUNCHED();
    edt_axis_kernel<1><<<grid, 256, 0, st>>>(b.d2b, b.d2a, Z, Y, X);
    ISG_LAUNCHED();
    edt_finish_kernel<<<grid, 256, 0, st>>>(b.d2a, b.key, distance_out, np);
    ISG_LAUNCHED();
    // ---- markers = ndi.label(centroids): raster-order ids -----------------------------------------
    ISG_CUDA(cudaMemsetAsync(b.comp_size, 0, np * sizeof(uint32_t), st));
    if ((rc = ccl_run(b.centroids, b.parent, b.comp_size, Z, Y, X, st))) return rc;
    root_flag_kernel<<<grid, 256, 0, st>>>(b.parent, b.flag, np);
    ISG_LAUNCHED();
    {
        size_t cb = b.cub_bytes;
        ISG_CUDA(cub::DeviceScan::ExclusiveSum(b.cub_tmp, cb, b.flag, b.rank, (int)np, st));
        count_launch(2);
    }
    marker_kernel<<<grid, 256, 0, st>>>(b.parent, b.rank, labels, b.seeds, b.seed_labels, (uint32_t)max_seeds,
                                        b.scal + 3, np);
    ISG_LAUNCHED();
    // ---- labels = watershed(-distance, markers, mask) ----------------------------------------------
    dog_domain_kernel<<<grid, 256, 0, st>>>(mask_out, b.centroids, b.dom, np);
    ISG_LAUNCHED();
    ISG_CUDA(cudaMemsetAsync(b.comp_size, 0, np * sizeof(uint32_t), st));
    ISG_CUDA(cudaMemsetAsync(b.comp_label, 0, np * sizeof(uint32_t), st));
    if ((rc = ccl_run(b.dom, b.parent, b.comp_size, Z, Y, X, st))) return rc;
    FloodGeom fg;
    fg.aff = nullptr;
    fg.plane_stride = 0;
    fg.origin = 0;
    fg.za = Z; fg.ya = Y; fg.xa = X;
    fg.div = nullptr;
    fg.scale[0] = fg.scale[1] = fg.scale[2] = 1.0f;
    fg.zp = Z; fg.yp = Y; fg.xp = X;
    fg.node_key = b.key;
    uint32_t n_seeds = 0;
    ISG_CUDA(cudaMemcpyAsync(&n_seeds, b.scal + 3, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    ISG_CUDA(cudaStreamSynchronize(st));
    ISG_REQUIRE(n_seeds <= (uint64_t)max_seeds, ISG_ERR_OVERFLOW, "isg_dog_blob_segment: %u marker voxels exceed max_seeds",
                n_seeds);
    // marker voxels were collected with atomics: the flood stage sorts them by (component, index)
    rc = flood_stage_run(b.flood, fg, mask_out, b.parent, b.comp_size, b.comp_label, b.seeds, (int64_t)n_seeds,
                         nullptr, labels, st, b.seed_labels);
    if (rc != ISG_OK) return rc;
    dog_counts_kernel<<<1, 1, 0, st>>>(b.scal + 0, b.scal + 2, b.scal + 3, b.rank, 0, counts_out);
    ISG_LAUNCHED();
    return ISG_OK;
}
