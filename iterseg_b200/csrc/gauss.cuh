// One axis of scipy.ndimage.gaussian_filter (correlate1d with a symmetric kernel), tiled through
// shared memory.  Arithmetic contract: double accumulation in scipy's
// pairing order  x[c]*w[0] + sum_{j=r..1} (x[c-j] + x[c+j])*w[j],  no FMA contraction, float32
// stored between passes; boundary 'nearest' (clamp) or 'reflect' (d c b a | a b c d | d c b a).
// The volume is seen as (outer, L, inner): L = the filtered axis, inner = its element stride.
#pragma once
#include "common.cuh"

namespace isg {

struct GaussW {
    double w[28];
    int r;
};

static constexpr int GAUSS_TL = 32;      // outputs along the filtered axis per tile
static constexpr int GAUSS_TI = 32;      // contiguous elements per tile row
static constexpr int GAUSS_RMAX = 27;     // sigma <= 6.8 at truncate = 4 (multi-layer blob_dog: sigma_list grows by 1.6x per layer)

__device__ __forceinline__ int gauss_src_index(int i, int len, int reflect) {
    if (reflect) {
        while (i < 0 || i >= len) i = i < 0 ? -i - 1 : 2 * len - i - 1;
        return i;
    }
    return i < 0 ? 0 : (i >= len ? len - 1 : i);
}

// Strided axes (inner >= 1 elements between neighbours along L, tiles of TI contiguous elements):
// grid = (ceil(inner / TI), ceil(L / TL), outer), block = (TI, 8).
// minmax (nullable): ordered-uint min / max of the outputs whose coordinate along the OUTERMOST
// volume axis z lies in [mm_z0, mm_z1); z_of_outer / z_of_l say where z lives for this view.
static __global__ void __launch_bounds__(GAUSS_TI * 8)
gauss_tiled_kernel(const float *__restrict__ in, float *__restrict__ out, int L, int64_t inner, GaussW gw,
                   int reflect, uint32_t *minmax, int z_is_l, int64_t outer_per_z, uint32_t mm_z0,
                   uint32_t mm_z1) {
    // converted to double once on the way in: every input is used 2r+1 times, and on this part the
    // float -> double conversions run on the (slow) FP64 pipe like the adds and multiplies
    __shared__ double tile[GAUSS_TL + 2 * GAUSS_RMAX][GAUSS_TI + 1];
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int64_t i0 = (int64_t)blockIdx.x * GAUSS_TI + tx;
    const int l0 = blockIdx.y * GAUSS_TL;
    const int64_t o = blockIdx.z;
    const float *src = in + o * (int64_t)L * inner;
    float *dst = out + o * (int64_t)L * inner;
    const int r = gw.r;
    __shared__ double w[GAUSS_RMAX + 1];                 // dynamically indexed: not from the parameter block
    if (ty == 0 && tx <= GAUSS_RMAX) w[tx] = gw.w[tx <= r ? tx : 0];
    for (int k = ty; k < GAUSS_TL + 2 * r; k += 8) {
        const int l = gauss_src_index(l0 + k - r, L, reflect);
        tile[k][tx] = i0 < inner ? (double)__ldg(src + (int64_t)l * inner + i0) : 0.0;
    }
    __syncthreads();
    uint32_t lo = 0xFFFFFFFFu, hi = 0u;
    for (int k = ty; k < GAUSS_TL; k += 8) {
        const int l = l0 + k;
        if (l >= L || i0 >= inner) continue;
        double acc = __dmul_rn(tile[k + r][tx], w[0]);
        for (int j = r; j >= 1; --j)
            acc = __dadd_rn(acc, __dmul_rn(__dadd_rn(tile[k + r - j][tx], tile[k + r + j][tx]), w[j]));
        const float v = (float)acc;
        dst[(int64_t)l * inner + i0] = v;
        if (minmax) {
            const uint32_t z = z_is_l ? (uint32_t)l : (uint32_t)(o / outer_per_z);
            if (z >= mm_z0 && z < mm_z1) {
                const uint32_t kk = f32_ord(v);
                lo = min(lo, kk);
                hi = max(hi, kk);
            }
        }
    }
    if (minmax) {
        lo = __reduce_min_sync(0xFFFFFFFFu, lo);
        hi = __reduce_max_sync(0xFFFFFFFFu, hi);
        if (tx == 0) {
            atomicMin(minmax + 0, lo);
            atomicMax(minmax + 1, hi);
        }
    }
}

// The contiguous axis (x): one row segment of 256 outputs per WARP (its own slice of shared
// memory, converted to double on the way in, no block-wide barrier); 8 outputs per lane.
// Work item = (row, segment), walked with a grid stride over warps.
static __global__ void __launch_bounds__(256)
gauss_row_kernel(const float *__restrict__ in, float *__restrict__ out, int X, int64_t rows, int64_t rows_per_z,
                 GaussW gw, int reflect, uint32_t *minmax, uint32_t mm_z0, uint32_t mm_z1) {
    __shared__ double seg_all[8][256 + 2 * GAUSS_RMAX];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double *seg = seg_all[warp];
    const int r = gw.r;
    __shared__ double w[GAUSS_RMAX + 1];
    if (threadIdx.x <= GAUSS_RMAX) w[threadIdx.x] = gw.w[(int)threadIdx.x <= r ? threadIdx.x : 0];
    __syncthreads();
    const int nseg = (X + 255) / 256;
    const int64_t items = rows * nseg;
    uint32_t lo = 0xFFFFFFFFu, hi = 0u;
    for (int64_t it = (int64_t)blockIdx.x * 8 + warp; it < items; it += (int64_t)gridDim.x * 8) {
        const int64_t row = it / nseg;
        const int x0 = (int)(it - row * nseg) * 256;
        const float *src = in + row * (int64_t)X;
        __syncwarp();
        for (int k = lane; k < 256 + 2 * r; k += 32)
            seg[k] = (double)__ldg(src + gauss_src_index(x0 + k - r, X, reflect));
        __syncwarp();
        const uint32_t z = (uint32_t)(row / rows_per_z);
        const bool mm = minmax && z >= mm_z0 && z < mm_z1;
#pragma unroll 2
        for (int q = 0; q < 8; ++q) {
            const int x = x0 + q * 32 + lane;
            if (x >= X) break;
            const int c = q * 32 + lane + r;
            double acc = __dmul_rn(seg[c], w[0]);
            for (int j = r; j >= 1; --j)
                acc = __dadd_rn(acc, __dmul_rn(__dadd_rn(seg[c - j], seg[c + j]), w[j]));
            const float v = (float)acc;
            out[row * (int64_t)X + x] = v;
            if (mm) {
                const uint32_t kk = f32_ord(v);
                lo = min(lo, kk);
                hi = max(hi, kk);
            }
        }
    }
    if (minmax) {
        lo = __reduce_min_sync(0xFFFFFFFFu, lo);
        hi = __reduce_max_sync(0xFFFFFFFFu, hi);
        if (lane == 0) {
            atomicMin(minmax + 0, lo);
            atomicMax(minmax + 1, hi);
        }
    }
}

// axis: 0 = z, 1 = y, 2 = x of a (Z, Y, X) volume
static inline int gauss_axis(const float *in, float *out, uint32_t Z, uint32_t Y, uint32_t X, int axis,
                             const GaussW &gw, int reflect, uint32_t *minmax, uint32_t mm_z0, uint32_t mm_z1,
                             cudaStream_t st) {
    if (axis == 2) {
        const int64_t rows = (int64_t)Z * Y;
        const int64_t items = rows * ((X + 255) / 256);
        int64_t blocks = (items + 7) / 8;
        const int64_t cap = (int64_t)num_sms() * 16;
        if (blocks > cap) blocks = cap;
        const unsigned grid = (unsigned)(blocks < 1 ? 1 : blocks);
        gauss_row_kernel<<<grid, 256, 0, st>>>(in, out, (int)X, rows, (int64_t)Y, gw, reflect, minmax, mm_z0, mm_z1);
        ISG_LAUNCHED();
        return ISG_OK;
    }
    const int L = axis == 0 ? (int)Z : (int)Y;
    const int64_t inner = axis == 0 ? (int64_t)Y * X : (int64_t)X;
    const int64_t outer = axis == 0 ? 1 : (int64_t)Z;
    ISG_REQUIRE(outer <= 65535, ISG_ERR_OVERFLOW, "gaussian: more than 65535 planes");
    dim3 grid((unsigned)((inner + GAUSS_TI - 1) / GAUSS_TI), (unsigned)((L + GAUSS_TL - 1) / GAUSS_TL), (unsigned)outer);
    gauss_tiled_kernel<<<grid, dim3(GAUSS_TI, 8), 0, st>>>(in, out, L, inner, gw, reflect, minmax, axis == 0 ? 1 : 0,
                                                          1, mm_z0, mm_z1);
    ISG_LAUNCHED();
    return ISG_OK;
}

}  // namespace isg
