// Memory-bound companions of the tensor-core convolutions (all channels-last fp16,
// fp32 math): train-mode BatchNorm apply + ReLU, fused with max-pool (encoder) or
// with the depthwise transposed convolution + crop (decoder); the final BatchNorm + sigmoid +
// crop-and-place.  (The two thin convolutions are tensor-core kernels of their own: c0.conv0 in
// unet_thin.cuh, c8_0.conv1 in unet_zring.cuh.)
//
// Reference: ConvModule.forward (src/iterseg/unet.py:91-106), MaxPool3d layers
// (unet.py:166-187), ConvTranspose3d layers (unet.py:216-242), crops + concat
// (unet.py:329-345), process_chunks placement (predict.py:89-95).
#pragma once
#include <cuda_fp16.h>

#include "common.cuh"

namespace isg {

static constexpr float BN_EPS = 1e-5f;

// BatchNorm3d in training mode: batch statistics (biased variance) of ONE chunk.
// stats_c = (sum, sum of squares) in 2^-24 fixed point (see STAT_SCALE in unet_conv.cuh).
// eps: BN_EPS / s^2 where s is the power of two the layer's weights of this output channel were
// divided by when they were packed (isg_unet_weights_pack): the stored convolution output is
// raw / s, and gamma * (raw - mean) / sqrt(var + eps) == gamma * (raw/s - mean/s) / sqrt(var/s^2 + eps/s^2).
__device__ __forceinline__ void bn_coeffs(const unsigned long long *stats_c, float gamma, float beta, float eps,
                                          float inv_count, float &scale, float &shift) {
    const double k = (double)inv_count / 16777216.0;
    const double mean = (double)(long long)stats_c[0] * k;
    const double var = fmax((double)(long long)stats_c[1] * k - mean * mean, 0.0);
    const float inv = 1.0f / sqrtf((float)var + eps);
    scale = gamma * inv;
    shift = beta - (float)mean * scale;
}

__device__ __forceinline__ void load8h(const __half *p, float (&v)[8]) {
    const uint4 u = *reinterpret_cast<const uint4 *>(p);
    const __half2 *h = reinterpret_cast<const __half2 *>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 f = __half22float2(h[i]);
        v[2 * i] = f.x;
        v[2 * i + 1] = f.y;
    }
}
__device__ __forceinline__ void store8h(__half *p, const float (&v)[8]) {
    uint4 u;
    __half2 *h = reinterpret_cast<__half2 *>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
    *reinterpret_cast<uint4 *>(p) = u;
}

// Shared scale/shift table for the chunk a block works on (C <= 256), as two float4 arrays per
// quantity -- [scale | shift][half][C/8] -- so that the 8
// channels of a thread are two float4 reads, consecutive lanes reading consecutive 16 bytes
// (a scalar s_scale[c0 + j] read has lanes 32 bytes apart: an 8-way bank conflict).
// Float index of channel c inside one quantity:
__device__ __forceinline__ int bn_slot(int c, int groups) {
    return ((c >> 2) & 1) * (groups * 4) + (c >> 3) * 4 + (c & 3);
}
__device__ __forceinline__ void bn_table4(float4 *tab, const unsigned long long *stats, const float *gamma,
                                          const float *beta, const float *eps, int C, int n, float inv_count) {
    float *t = reinterpret_cast<float *>(tab);
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float sc, sh;
        bn_coeffs(stats + ((size_t)n * C + c) * 2, gamma[c], beta[c], eps[c], inv_count, sc, sh);
        t[bn_slot(c, C / 8)] = sc;
        t[C + bn_slot(c, C / 8)] = sh;
    }
    __syncthreads();
}
// v = relu(v * scale + shift) for the 8 channels of group g
__device__ __forceinline__ void bn_relu8(const float4 *tab, int groups, int g, float (&v)[8]) {
    const float4 sc0 = tab[g], sc1 = tab[groups + g], sh0 = tab[2 * groups + g], sh1 = tab[3 * groups + g];
    v[0] = fmaxf(fmaf(v[0], sc0.x, sh0.x), 0.0f);
    v[1] = fmaxf(fmaf(v[1], sc0.y, sh0.y), 0.0f);
    v[2] = fmaxf(fmaf(v[2], sc0.z, sh0.z), 0.0f);
    v[3] = fmaxf(fmaf(v[3], sc0.w, sh0.w), 0.0f);
    v[4] = fmaxf(fmaf(v[4], sc1.x, sh1.x), 0.0f);
    v[5] = fmaxf(fmaf(v[5], sc1.y, sh1.y), 0.0f);
    v[6] = fmaxf(fmaf(v[6], sc1.z, sh1.z), 0.0f);
    v[7] = fmaxf(fmaf(v[7], sc1.w, sh1.w), 0.0f);
}

// raw -> relu(bn(raw)), same shape.  grid = (blocks, N)
__global__ void __launch_bounds__(256)
bn_relu_kernel(const __half *__restrict__ raw, __half *__restrict__ act, const unsigned long long *__restrict__ stats,
               const float *__restrict__ gamma, const float *__restrict__ beta, const float *__restrict__ eps,
               int C, size_t vox) {
    __shared__ float4 tab[128];                          // C <= 256
    const int n = blockIdx.y;
    bn_table4(tab, stats, gamma, beta, eps, C, n, 1.0f / (float)vox);
    const int groups = C / 8;
    const size_t total = vox * groups;
    const __half *src = raw + (size_t)n * vox * C;
    __half *dst = act + (size_t)n * vox * C;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (size_t)gridDim.x * blockDim.x) {
        float v[8];
        load8h(src + i * 8, v);
        bn_relu8(tab, groups, (int)(i % groups), v);
        store8h(dst + i * 8, v);
    }
}

// Encoder: skip = relu(bn(raw)) at the fine level and pooled = maxpool(skip) with
// kernel = stride = (PZ,2,2), padding (0,1,1) (-inf padding: border windows just see
// fewer voxels).  One thread per (coarse voxel, 8 channels).  grid = (blocks, N)
template <int PZ>
__global__ void __launch_bounds__(256)
bn_relu_pool_kernel(const __half *__restrict__ raw, __half *__restrict__ skip,
                    __half *__restrict__ pooled, const unsigned long long *__restrict__ stats,
                    const float *__restrict__ gamma, const float *__restrict__ beta,
                    const float *__restrict__ eps, int C, int D,
                    int H, int W, int Dc, int Hc, int Wc) {
    __shared__ float4 tab[128];                          // C <= 256
    const int n = blockIdx.y;
    const size_t vox = (size_t)D * H * W;
    bn_table4(tab, stats, gamma, beta, eps, C, n, 1.0f / (float)vox);
    const int groups = C / 8;
    const size_t total = (size_t)Dc * Hc * Wc * groups;
    const __half *src = raw + (size_t)n * vox * C;
    __half *dsk = skip + (size_t)n * vox * C;
    __half *dpo = pooled + (size_t)n * Dc * Hc * Wc * C;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (size_t)gridDim.x * blockDim.x) {
        const int c0 = (int)(i % groups) * 8;
        size_t t = i / groups;
        const int wc = (int)(t % Wc); t /= Wc;
        const int hc = (int)(t % Hc);
        const int dc = (int)(t / Hc);
        float m[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) m[j] = 0.0f;           // post-ReLU values are >= 0
        const int g = c0 >> 3;
        const float4 sc0 = tab[g], sc1 = tab[groups + g], sh0 = tab[2 * groups + g], sh1 = tab[3 * groups + g];
        const float sc[8] = {sc0.x, sc0.y, sc0.z, sc0.w, sc1.x, sc1.y, sc1.z, sc1.w};
        const float sh[8] = {sh0.x, sh0.y, sh0.z, sh0.w, sh1.x, sh1.y, sh1.z, sh1.w};
#pragma unroll
        for (int kz = 0; kz < PZ; ++kz) {
            const int d = dc * PZ + kz;
            if (d >= D) continue;
#pragma unroll
            for (int a = 0; a < 2; ++a) {
                const int h = 2 * hc - 1 + a;
                if (h < 0 || h >= H) continue;
#pragma unroll
                for (int b = 0; b < 2; ++b) {
                    const int w = 2 * wc - 1 + b;
                    if (w < 0 || w >= W) continue;
                    const size_t off = ((((size_t)d * H + h) * W + w) * C) + c0;
                    float v[8];
                    load8h(src + off, v);
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        v[j] = fmaxf(fmaf(v[j], sc[j], sh[j]), 0.0f);
                        // the pooled tensor holds the max of the fp16-ROUNDED skip values so that
                        // it equals maxpool(skip) exactly
                        v[j] = __half2float(__float2half_rn(v[j]));
                        m[j] = fmaxf(m[j], v[j]);
                    }
                    store8h(dsk + off, v);
                }
            }
        }
        store8h(dpo + ((((size_t)dc * Hc + hc) * Wc + wc) * C) + c0, m);
    }
}

// Decoder: up = crop(conv_transpose3d(relu(bn(raw)), w, b, kernel = stride = (KZ,2,2),
// groups = C)).  out[c, KZ*d+kz, 2h+a-off, 2w+b-off] = wgt[c,kz,a,b] * x[c,d,h,w] + bias[c];
// off = 0 with crop [:-1,:-1] (up0..up2), off = 1 with crop [1:-1,1:-1] (up3).
// One thread per (coarse voxel, 8 channels).  grid = (blocks, N)
// All per-channel constants sit in shared memory as [table][half][C/8] float4 -- the 8 channels of
// a thread are two float4 reads per table, consecutive lanes read consecutive 16 bytes (no bank
// conflicts).  Reading the (C, KZ*4) weights straight from global memory costs 16 strided
// 4-byte loads per tap, 32 sectors each: that, not HBM, bounded the first version (12x off the
// roofline at C = 256).
template <int KZ>
__global__ void __launch_bounds__(256)
bn_relu_up_kernel(const __half *__restrict__ raw, __half *__restrict__ up,
                  const unsigned long long *__restrict__ stats, const float *__restrict__ gamma,
                  const float *__restrict__ beta, const float *__restrict__ eps,
                  const float *__restrict__ wgt,
                  const float *__restrict__ bias, int C, int Dc, int Hc, int Wc, int Df, int Hf,
                  int Wf, int off) {
    constexpr int TAPS = KZ * 4;
    extern __shared__ float4 up_tab[];                  // [3 + TAPS][2][C/8]: scale, shift, bias, weights per tap
    const int n = blockIdx.y;
    const size_t vox = (size_t)Dc * Hc * Wc;
    const int groups = C / 8;
    {
        float *tab = reinterpret_cast<float *>(up_tab);
        // channel c -> float index inside one table: [half = (c & 4) >> 2][group = c >> 3][c & 3]
        auto slot = [&](int c) { return ((c >> 2) & 1) * (groups * 4) + (c >> 3) * 4 + (c & 3); };
        for (int c = threadIdx.x; c < C; c += blockDim.x) {
            float sc, sh;
            bn_coeffs(stats + ((size_t)n * C + c) * 2, gamma[c], beta[c], eps[c], 1.0f / (float)vox, sc, sh);
            tab[0 * C + slot(c)] = sc;
            tab[1 * C + slot(c)] = sh;
            tab[2 * C + slot(c)] = bias[c];
        }
        for (int i = threadIdx.x; i < C * TAPS; i += blockDim.x) {
            const int c = i / TAPS, tap = i - c * TAPS;
            tab[(3 + tap) * C + slot(c)] = wgt[i];
        }
    }
    __syncthreads();
    const size_t total = vox * groups;
    const __half *src = raw + (size_t)n * vox * C;
    __half *dst = up + (size_t)n * Df * Hf * Wf * C;
    const int tstride = 2 * groups;                     // float4 per table
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (size_t)gridDim.x * blockDim.x) {
        const int g = (int)(i % groups);
        size_t t = i / groups;
        const int wc = (int)(t % Wc); t /= Wc;
        const int hc = (int)(t % Hc);
        const int dc = (int)(t / Hc);
        float x[8], bs[8];
        load8h(src + i * 8, x);
        {
            const float4 sc0 = up_tab[g], sc1 = up_tab[groups + g];
            const float4 sh0 = up_tab[tstride + g], sh1 = up_tab[tstride + groups + g];
            const float4 b0 = up_tab[2 * tstride + g], b1 = up_tab[2 * tstride + groups + g];
            const float sc[8] = {sc0.x, sc0.y, sc0.z, sc0.w, sc1.x, sc1.y, sc1.z, sc1.w};
            const float sh[8] = {sh0.x, sh0.y, sh0.z, sh0.w, sh1.x, sh1.y, sh1.z, sh1.w};
            const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                x[j] = fmaxf(fmaf(x[j], sc[j], sh[j]), 0.0f);
                bs[j] = bb[j];
            }
        }
#pragma unroll
        for (int tap = 0; tap < TAPS; ++tap) {
            const int kz = tap >> 2, a = (tap >> 1) & 1, b = tap & 1;
            const int d = dc * KZ + kz, h = 2 * hc + a - off, w = 2 * wc + b - off;
            if (d >= Df || h < 0 || h >= Hf || w < 0 || w >= Wf) continue;
            const float4 w0 = up_tab[(3 + tap) * tstride + g], w1 = up_tab[(3 + tap) * tstride + groups + g];
            const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
            float o[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = fmaf(wv[j], x[j], bs[j]);
            store8h(dst + ((((size_t)d * Hf + h) * Wf + w) * C) + g * 8, o);
        }
    }
}

// sigmoid(bn(raw9)) -> the cropped interior of every chunk is placed into the
// (5, Z, Y, X) feature volume (predict.py:89-95): each voxel is written by exactly one chunk.
__global__ void __launch_bounds__(256)
place_kernel(const __half *__restrict__ raw9, const unsigned long long *__restrict__ stats9,
             const float *__restrict__ gamma, const float *__restrict__ beta, const float *__restrict__ eps,
             const int *__restrict__ starts, const int *__restrict__ crop_lo,
             const int *__restrict__ crop_hi, float *__restrict__ feats, int Z, int Y, int X, int D,
             int H, int W) {
    __shared__ float sc[5], sh[5];
    const int n = blockIdx.y;
    const size_t vox = (size_t)D * H * W;
    if (threadIdx.x < 5)
        bn_coeffs(stats9 + ((size_t)n * 5 + threadIdx.x) * 2, gamma[threadIdx.x], beta[threadIdx.x],
                  eps[threadIdx.x], 1.0f / (float)vox, sc[threadIdx.x], sh[threadIdx.x]);
    __syncthreads();
    const int lz = crop_lo[n * 3], ly = crop_lo[n * 3 + 1], lx = crop_lo[n * 3 + 2];
    const int cd = crop_hi[n * 3] - lz, ch = crop_hi[n * 3 + 1] - ly, cw = crop_hi[n * 3 + 2] - lx;
    const int z0 = starts[n * 3], y0 = starts[n * 3 + 1], x0 = starts[n * 3 + 2];
    const size_t total = (size_t)cd * ch * cw;
    const size_t plane = (size_t)Z * Y * X;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (size_t)gridDim.x * blockDim.x) {
        const int w = lx + (int)(i % cw);
        const size_t t = i / cw;
        const int h = ly + (int)(t % ch);
        const int d = lz + (int)(t / ch);
        const uint4 u = __ldg(reinterpret_cast<const uint4 *>(
            raw9 + ((size_t)n * vox + ((size_t)d * H + h) * W + w) * 8));
        const __half2 *hh = reinterpret_cast<const __half2 *>(&u);
        const float2 a01 = __half22float2(hh[0]), a23 = __half22float2(hh[1]), a45 = __half22float2(hh[2]);
        const float x[5] = {a01.x, a01.y, a23.x, a23.y, a45.x};
        const size_t o = ((size_t)(z0 + d) * Y + (y0 + h)) * X + (x0 + w);
#pragma unroll
        for (int c = 0; c < 5; ++c) {
            const float y = fmaf(x[c], sc[c], sh[c]);
            feats[c * plane + o] = 1.0f / (1.0f + expf(-y));
        }
    }
}

// ---- weight packing -------------------------------------------------------------
// nn.Conv3d weight (Cout, Cin, 3,3,3) fp32 -> [tap][cout_pad][cin] fp16 (rows >= Cout zero)
// Per output channel: s = 2^ceil(log2(||w_co||_2)) (1 for an all-zero filter), inv_s = 1 / s,
// eps = BN_EPS / s^2.  Every convolution is followed by a train-mode BatchNorm, which makes the
// network's output invariant to the scale of a filter (up to eps, handled in bn_coeffs); packing
// w / s keeps the fp16 weights and the fp16 pre-BatchNorm activations O(1) whatever scale a
// training run left the filters at -- fp16 has a 65504 ceiling that the reference's fp32 (and
// bf16) do not have.  Powers of two: the division is exact.
__global__ void __launch_bounds__(256)
conv_prescale_kernel(const float *__restrict__ src, int cin, float *__restrict__ inv_s, float *__restrict__ eps,
                     int enable) {
    __shared__ double part[256];
    const int co = blockIdx.x;
    double acc = 0.0;
    for (int i = threadIdx.x; i < cin * 27; i += blockDim.x) {
        const double v = (double)src[(size_t)co * cin * 27 + i];
        acc += v * v;
    }
    part[threadIdx.x] = acc;
    __syncthreads();
    for (int st = 128; st > 0; st >>= 1) {
        if ((int)threadIdx.x < st) part[threadIdx.x] += part[threadIdx.x + st];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        int e = 0;
        const double nrm = sqrt(part[0]);
        if (enable && nrm > 0.0 && isfinite(nrm)) {
            e = (int)ceil(log2(nrm));
            e = e < -60 ? -60 : (e > 60 ? 60 : e);
        }
        inv_s[co] = (float)ldexp(1.0, -e);
        eps[co] = (float)ldexp((double)BN_EPS, -2 * e);
    }
}

__global__ void pack_conv_w_kernel(const float *__restrict__ src, const float *__restrict__ inv_s,
                                   __half *__restrict__ dst, int cout,
                                   int cout_pad, int cin) {
    const size_t total = (size_t)27 * cout_pad * cin;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (size_t)gridDim.x * blockDim.x) {
        const int ci = (int)(i % cin);
        const size_t t = i / cin;
        const int co = (int)(t % cout_pad);
        const int tap = (int)(t / cout_pad);
        float v = 0.0f;
        if (co < cout) v = src[((size_t)co * cin + ci) * 27 + tap] * inv_s[co];
        dst[i] = __float2half_rn(v);
    }
}
// (Cout, Cin, 27) fp32 -> [tap][cin][cout] fp32 (CUDA-core layers)
__global__ void pack_conv_w_f32_kernel(const float *__restrict__ src, const float *__restrict__ inv_s,
                                       float *__restrict__ dst, int cout, int cin) {
    const int total = 27 * cin * cout;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int co = i % cout;
        const int t = i / cout;
        const int ci = t % cin;
        const int tap = t / cin;
        dst[i] = src[((size_t)co * cin + ci) * 27 + tap] * inv_s[co];
    }
}
__global__ void copy_f32_kernel(const float *__restrict__ src, float *__restrict__ dst, int n) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) dst[i] = src[i];
}

// fp16 channels-last [vox][C] of one chunk -> fp32 [C][vox] (debug / parity only)
__global__ void debug_to_ncdhw_kernel(const void *__restrict__ src, int is_f32, int cstride, int C,
                                      size_t vox, const float *__restrict__ inv_s, float *__restrict__ dst) {
    const size_t total = vox * C;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (size_t)gridDim.x * blockDim.x) {
        const size_t v = i % vox;
        const int c = (int)(i / vox);
        const float x = is_f32 ? reinterpret_cast<const float *>(src)[v * cstride + c]
                               : __half2float(reinterpret_cast<const __half *>(src)[v * cstride + c]);
        dst[i] = x / inv_s[c];                       // undo the filter's power-of-two prescale (exact)
    }
}

}  // namespace isg
