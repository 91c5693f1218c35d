// U-Net over chunks: plan (buffers, TMA tensor maps, tile geometry), weight packing
// and the forward pass.  C-ABI: isg_unet_* (include/iterseg_b200.h).
//
// Replaces process_chunks + predict_chunk_feature_map (src/iterseg/predict.py:64-126)
// and UNet.forward (src/iterseg/unet.py:284-364) for UNet(in_channels=1, out_channels=5).
// BatchNorm runs with per-chunk batch statistics exactly like the reference, which
// never calls .eval() (predict.py:25-35,118-123); conv biases are dropped because the
// mean subtraction cancels them.
#include <cuda.h>
#include <cuda_fp16.h>
#include <string.h>

#include <string>
#include <vector>

#include "unet_conv.cuh"
#include "unet_elem.cuh"
#include "unet_thin.cuh"
#include "unet_zring.cuh"
#include "unet_zring32.cuh"

namespace isg {

// ---- static architecture tables (unet.py:192-210, decoder_instructions :8-21) ----------
struct ConvDef {
    const char *name;
    int cin, cout, level;
};
static const ConvDef CONVS[18] = {
    {"c0.conv0", 1, 32, 0},     {"c0.conv1", 32, 32, 0},    {"c1.conv0", 32, 64, 1},
    {"c1.conv1", 64, 64, 1},    {"c2.conv0", 64, 128, 2},   {"c2.conv1", 128, 128, 2},
    {"c3.conv0", 128, 256, 3},  {"c3.conv1", 256, 256, 3},  {"c4.conv0", 256, 256, 4},
    {"c4.conv1", 256, 256, 4},  {"c5_0.conv0", 512, 128, 3}, {"c5_0.conv1", 128, 128, 3},
    {"c6_0.conv0", 256, 64, 2}, {"c6_0.conv1", 64, 64, 2},  {"c7_0.conv0", 128, 32, 1},
    {"c7_0.conv1", 32, 32, 1},  {"c8_0.conv0", 64, 5, 0},   {"c8_0.conv1", 5, 5, 0}};
static const int UP_C[4] = {256, 128, 64, 32};
static const int UP_KZ[4] = {2, 1, 1, 1};

static inline int cout_pad(int i) { return CONVS[i].cout == 5 ? 16 : CONVS[i].cout; }
static inline bool is_tc(int i) { return i != 0 && i != 17; }

struct PackLayout {
    size_t w[18];          // conv weights (fp16 tap-major for tensor-core layers, fp32 otherwise)
    size_t gamma[18], beta[18];
    size_t inv_s[18], eps[18];   // per output channel: 1 / (power-of-two prescale of the filter), BN_EPS / s^2
    size_t up_w[4], up_b[4];
    size_t total;
};
static PackLayout pack_layout() {
    PackLayout L;
    size_t off = 0;
    auto take = [&](size_t bytes) {
        size_t at = off;
        off += (bytes + 255) & ~(size_t)255;
        return at;
    };
    for (int i = 0; i < 18; ++i) {
        if (is_tc(i)) L.w[i] = take((size_t)27 * cout_pad(i) * CONVS[i].cin * sizeof(__half));
        else L.w[i] = take((size_t)27 * CONVS[i].cin * CONVS[i].cout * sizeof(float));
        L.gamma[i] = take(CONVS[i].cout * sizeof(float));
        L.beta[i] = take(CONVS[i].cout * sizeof(float));
        L.inv_s[i] = take(CONVS[i].cout * sizeof(float));
        L.eps[i] = take(CONVS[i].cout * sizeof(float));
    }
    for (int u = 0; u < 4; ++u) {
        L.up_w[u] = take((size_t)UP_C[u] * UP_KZ[u] * 4 * sizeof(float));
        L.up_b[u] = take(UP_C[u] * sizeof(float));
    }
    L.total = off;
    return L;
}

// ---- driver entry point for tensor maps (no link-time dependency on libcuda) ----------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *,
                                  const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                  const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

static bool make_act_map(CUtensorMap *m, const void *base, int C, int W, int H, int D, int N,
                         int cblk, int P, int Ht) {
    cuuint64_t dims[5] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)N};
    cuuint64_t strides[4] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2,
                             (cuuint64_t)D * H * W * C * 2};
    cuuint32_t box[5] = {(cuuint32_t)cblk, (cuuint32_t)P, (cuuint32_t)(Ht + 2), 1u, 1u};
    cuuint32_t es[5] = {1, 1, 1, 1, 1};
    CUresult r = encode_fn()(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, const_cast<void *>(base), dims,
                             strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             cblk == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                             CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) set_error("cuTensorMapEncodeTiled(activation C=%d W=%d H=%d D=%d N=%d box %d,%d,%d) -> %d",
                                     C, W, H, D, N, cblk, P, Ht + 2, (int)r);
    return r == CUDA_SUCCESS;
}
static bool make_w_map(CUtensorMap *m, const void *base, int cin, int coutp, int cblk, int G) {
    cuuint64_t dims[3] = {(cuuint64_t)cin, (cuuint64_t)coutp, 27};
    cuuint64_t strides[2] = {(cuuint64_t)cin * 2, (cuuint64_t)coutp * cin * 2};
    cuuint32_t box[3] = {(cuuint32_t)cblk, (cuuint32_t)coutp, (cuuint32_t)G};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = encode_fn()(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void *>(base), dims,
                             strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             cblk == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                             CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) set_error("cuTensorMapEncodeTiled(weights cin=%d cout=%d) -> %d", cin, coutp, (int)r);
    return r == CUDA_SUCCESS;
}

// c8_0.conv0 weights packed [dy][80 rows][cin] (unet_zring.cuh): one box = (32 channels, 80 rows, 1 dy)
static bool make_w_map_zring(CUtensorMap *m, const void *base, int cin) {
    cuuint64_t dims[3] = {(cuuint64_t)cin, (cuuint64_t)ZR_N, 3};
    cuuint64_t strides[2] = {(cuuint64_t)cin * 2, (cuuint64_t)ZR_N * cin * 2};
    cuuint32_t box[3] = {32u, (cuuint32_t)ZR_N, 1u};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = encode_fn()(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void *>(base), dims,
                             strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
                             CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) set_error("cuTensorMapEncodeTiled(z-ring weights) -> %d", (int)r);
    return r == CUDA_SUCCESS;
}

// 32 -> 32 z-ring weights packed [dy*3+dx][96 rows][32] (unet_zring32.cuh): one box = one tap
static bool make_w_map_z32(CUtensorMap *m, const void *base) {
    cuuint64_t dims[3] = {32, (cuuint64_t)Z32_N, 9};
    cuuint64_t strides[2] = {32 * 2, (cuuint64_t)Z32_N * 32 * 2};
    cuuint32_t box[3] = {32u, (cuuint32_t)Z32_N, 1u};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = encode_fn()(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void *>(base), dims,
                             strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
                             CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) set_error("cuTensorMapEncodeTiled(z-ring 32->32 weights) -> %d", (int)r);
    return r == CUDA_SUCCESS;
}

// ---- plan -------------------------------------------------------------------------------
struct TcLayer {
    ConvGeom g;
    CUtensorMap tmA0, tmA1, tmB;
    int cblk;
    int fold;
    size_t smem;
    int grid;
    int zring;                   // 1: c8_0.conv0 runs conv3d_zring_kernel (unet_zring.cuh) with `zg`
    ZringGeom zg;
    Z32Args z32;                 // zring == 2: c0.conv1 / c7_0.conv1 run conv3d_zring32_kernel (unet_zring32.cuh)
};

}  // namespace isg

struct isg_unet_plan {
    int N, Z, Y, X;
    int cz, cy, cx;
    int D[5], H[5], W[5];
    const unsigned char *packed;
    isg::PackLayout L;
    // device buffers
    int *starts, *crop_lo, *crop_hi;
    // chunk tables: host copy in pinned memory, uploaded by the next forward pass with
    // cudaMemcpyAsync on ITS stream (ordered against that stream's earlier kernels)
    unsigned int *overflow_host;  // pinned: g_unet_overflow as of the last completed forward pass
    int *tabs_host;               // pinned, 9 * N ints: starts | crop_lo | crop_hi
    int tabs_dirty;
    cudaEvent_t tabs_ev;          // recorded after an upload: the pinned copy may be rewritten once it is done
    int tabs_ev_pending;
    __half *raw[5], *act[5], *skip[4], *pooled[5], *up[4];
    __half *raw8, *raw9;          // the two 5-channel layers: fp16 [vox][8] (16 bytes per voxel)
    unsigned long long *stats[18];
    unsigned long long *stats_all;
    unsigned int *sched[18];
    size_t stats_bytes;
    isg::TcLayer tc[18];
    int base_off_mode;
    double flops;
    double tc_flops;              // algorithmic FLOPs of the 16 tensor-core convolutions
    // optional per-launch CUDA-event timing (bench roofline): kind 0 = tcgen05 conv, 1 = other
    int profiling;
    std::vector<cudaEvent_t> ev;  // pairs (start, stop)
    std::vector<int> ev_kind;
    size_t ev_used;
};

namespace isg {

static void choose_tile(int H, int W, int rb, int &P, int &Ht) {
    long best = -1;
    P = 32; Ht = 4;
    for (int p = 10; p <= 130; ++p) {
        const int ht = 130 / p;
        if (ht < 1) continue;
        const int wt = p - 2;
        if ((long)(ht + 2) * p * rb > 27 * 1024) continue;       // one plane slot
        const long tiles = (long)((W + wt - 1) / wt) * ((H + ht - 1) / ht);
        const long cost = tiles * 4096 + (long)(ht + 2) * p;      // fewest tiles, then smallest halo
        if (best < 0 || cost < best) {
            best = cost;
            P = p;
            Ht = ht;
        }
    }
}

// Band-flat tiling (ConvGeom::flat): pitch P = band width + 2, R = the padded rows a 128-position
// tile plus its 2P + 2 tap reach can touch, fewest tiles first, then the smallest slot.
static bool choose_flat(int H, int W, int rb, long max_slot, int &P, int &R, long &tiles) {
    long best_t = -1, best_s = 0;
    for (int p = 10; p <= 256; ++p) {
        const int wb = p - 2;
        const int r = (p - 1 + 128 + 2 * p + 2 + p - 1) / p;
        const long slot = (((long)r * p * rb) + 1023) & ~1023L;
        if (slot > max_slot || r > 256) continue;
        const long t = (long)((W + wb - 1) / wb) * (((long)H * p + 127) / 128);
        if (best_t < 0 || t < best_t || (t == best_t && slot < best_s)) {
            best_t = t; best_s = slot; P = p; R = r;
        }
    }
    tiles = best_t;
    return best_t > 0;
}

static size_t plan_carve(isg_unet_plan *p, Carver &cv) {
    const int N = p->N;
    auto vox = [&](int l) { return (size_t)N * p->D[l] * p->H[l] * p->W[l]; };
    static const int CH[5] = {32, 64, 128, 256, 256};
    p->starts = cv.take<int>(N * 9);                  // starts | crop_lo | crop_hi, one upload per forward
    p->crop_lo = p->starts ? p->starts + 3 * N : nullptr;
    p->crop_hi = p->starts ? p->starts + 6 * N : nullptr;
    for (int l = 0; l < 5; ++l) {
        p->raw[l] = cv.take<__half>(vox(l) * CH[l]);
        p->act[l] = cv.take<__half>(vox(l) * CH[l]);
        if (l < 4) {
            p->skip[l] = cv.take<__half>(vox(l) * CH[l]);
            p->up[l] = cv.take<__half>(vox(l) * CH[l]);
        }
        if (l > 0) p->pooled[l] = cv.take<__half>(vox(l) * CH[l - 1]);
        else p->pooled[l] = nullptr;
    }
    p->raw8 = cv.take<__half>(vox(0) * 8);
    p->raw9 = cv.take<__half>(vox(0) * 8);
    size_t floats = 0;
    for (int i = 0; i < 18; ++i) floats += (size_t)N * cout_pad(i) * 2;
    // + one 8-byte slot per conv for the group counter of its dynamic scheduler (unet_conv.cuh);
    // the memset that clears the statistics at the start of a forward pass clears these too
    p->stats_all = cv.take<unsigned long long>(floats + 18);
    p->stats_bytes = (floats + 18) * sizeof(unsigned long long);
    if (p->stats_all) {
        unsigned long long *s = p->stats_all;
        for (int i = 0; i < 18; ++i) {
            p->stats[i] = s;
            s += (size_t)N * cout_pad(i) * 2;
        }
        for (int i = 0; i < 18; ++i) p->sched[i] = reinterpret_cast<unsigned int *>(s + i);
    }
    return cv.off;
}

static bool level_dims(isg_unet_plan *p, int cz, int cy, int cx) {
    p->D[0] = cz; p->H[0] = cy; p->W[0] = cx;
    for (int l = 1; l < 5; ++l) {
        p->D[l] = l == 4 ? p->D[l - 1] / 2 : p->D[l - 1];
        p->H[l] = p->H[l - 1] / 2 + 1;
        p->W[l] = p->W[l - 1] / 2 + 1;
    }
    // the decoder's crops must reproduce the skip shapes (unet.py:329-345), else torch.cat fails
    bool ok = cz >= 2 && cz % 2 == 0;
    for (int l = 3; l >= 1; --l) ok = ok && 2 * p->H[l + 1] - 1 == p->H[l] && 2 * p->W[l + 1] - 1 == p->W[l];
    ok = ok && 2 * p->H[1] - 2 == p->H[0] && 2 * p->W[1] - 2 == p->W[0];
    return ok;
}

static bool setup_tc_layer(isg_unet_plan *p, int i, const __half *src0, int c0, const __half *src1,
                           int c1, void *out, int out_mode) {
    TcLayer &t = p->tc[i];
    const int l = CONVS[i].level;
    const int cin = CONVS[i].cin;
    t.zring = 0;
    if ((i == 1 || i == 15) && getenv("ISG_NO_ZRING32") == nullptr) {
        // 32 -> 32 at the two finest levels: dz folded into N on a TMEM ring (unet_zring32.cuh)
        Z32Args &z = t.z32;
        z.N = p->N; z.D = p->D[l]; z.H = p->H[l]; z.W = p->W[l];
        z.tiles_w = (z.W + ZR_WT - 1) / ZR_WT;
        z.tiles_h = (z.H + ZR_HT - 1) / ZR_HT;
        z.n_cols = z.N * z.tiles_h * z.tiles_w;
        z.out = reinterpret_cast<__half *>(out);
        z.stats = p->stats[i];
        z.sched = p->sched[i];
        t.zring = 2;
        t.cblk = 32;
        t.fold = 0;
        t.g = ConvGeom{};
        t.smem = z32_smem_bytes();
        t.grid = z.n_cols < num_sms() ? z.n_cols : num_sms();
        if (c0 != 32 || c1 != 0 || out_mode != 0 || CONVS[i].cout != 32) {
            set_error("conv %s: the 32 -> 32 z-ring kernel does not fit", CONVS[i].name);
            return false;
        }
        return make_act_map(&t.tmA0, src0, 32, z.W, z.H, z.D, z.N, 32, ZR_P, ZR_HT) &&
               make_w_map_z32(&t.tmB, p->packed + p->L.w[i]);
    }
    if (i == 16) {
        // 64 -> 5 on [up3, skip0]: the z-ring kernel (nine (dz,dx) taps folded into N)
        ZringGeom &z = t.zg;
        z.N = p->N; z.D = p->D[l]; z.H = p->H[l]; z.W = p->W[l];
        z.tiles_w = (z.W + ZR_WT - 1) / ZR_WT;
        z.tiles_h = (z.H + ZR_HT - 1) / ZR_HT;
        z.n_cols = z.N * z.tiles_h * z.tiles_w;
        z.out = reinterpret_cast<__half *>(out);
        z.stats = p->stats[i];
        z.sched = p->sched[i];
        t.zring = 1;
        t.cblk = 32;
        t.fold = 0;
        t.g = ConvGeom{};
        t.smem = zring_smem_bytes();
        t.grid = z.n_cols < 2 * num_sms() ? z.n_cols : 2 * num_sms();      // two CTAs per SM
        if (c0 != 32 || c1 != 32 || out_mode != 1) {
            set_error("conv %s: the z-ring kernel expects two 32-channel sources", CONVS[i].name);
            return false;
        }
        return make_act_map(&t.tmA0, src0, 32, z.W, z.H, z.D, z.N, 32, ZR_P, ZR_HT) &&
               make_act_map(&t.tmA1, src1, 32, z.W, z.H, z.D, z.N, 32, ZR_P, ZR_HT) &&
               make_w_map_zring(&t.tmB, p->packed + p->L.w[i], cin);
    }
    // Channel block of the K loop (= shared-memory row: 64 channels -> 128-byte swizzle, 32 -> 64-byte).
    // What decides (measured per layer on B200, profiles/r02_notes.md): the MMAs issued back to back into ONE
    // accumulator -- taps per weight stage x K steps per block -- because every change of the D tile costs the
    // tensor pipe ~150-300 clocks (the accumulator is written back and the next one fetched), and the room the
    // plane slots leave for weight stages:
    //   Cout <= 64 : 32-channel blocks, 9 taps per stage (18 MMAs per visit; 64-channel blocks lose the flat tiles
    //                and planes per group: c1.conv1 1.43 -> 1.31 ms);
    //   Cout = 128 : 64-channel blocks, 3 taps per 48 KB stage (12 MMAs per visit; c2.conv1 1.00 -> 0.98 ms,
    //                c5_0.conv0 1.02 -> 1.00 against 32-channel blocks with 9 taps per stage) -- with at least two
    //                blocks: c2.conv0 (Cin = 64) keeps 32-channel blocks and its flat tiles (0.60 vs 0.615 ms);
    //   Cout = 256 : 32-channel blocks, 3 taps per 48 KB stage (6 MMAs per visit instead of 4 with one 64-channel
    //                tap: c3.conv1 1.09 -> 1.06 ms, c4.* 0.45 -> 0.43);
    //   dx-fold    : 64-channel blocks (c7_0.conv0 1.31 vs 1.75 ms).
    // ISG_CONV_CBLK = "<layer index>:<32|64>,..." overrides (diagnosis).
    int cblk = (c0 % 64 == 0 && (c1 % 64 == 0)) ? 64 : 32;
    const bool fold_layer = cout_pad(i) <= 32 && p->W[l] >= 100;
    if (!fold_layer && (cout_pad(i) <= 64 || cout_pad(i) > 128 || cin < 128)) cblk = 32;
    if (const char *ov = getenv("ISG_CONV_CBLK")) {
        char key[16];
        snprintf(key, sizeof key, ",%d:", i);
        std::string lst = std::string(",") + ov + ",";
        const size_t at = lst.find(key);
        if (at != std::string::npos) {
            const int v = atoi(lst.c_str() + at + strlen(key));
            if (v == 32 || (v == 64 && c0 % 64 == 0 && c1 % 64 == 0)) cblk = v;
        }
    }
    t.cblk = cblk;
    ConvGeom &g = t.g;
    g.N = p->N; g.D = p->D[l]; g.H = p->H[l]; g.W = p->W[l];
    g.cout = cout_pad(i);
    g.nkb0 = c0 / cblk;
    g.nkb1 = c1 / cblk;
    const int nkb = g.nkb0 + g.nkb1;
    const long budget = 232448 - (1024 + CONV_SLACK + CONV_BAR_BYTES + 4 * 32 * 33 * 4);
    const int tap_bytes = g.cout * cblk * 2;
    // Candidate configurations, best first:
    //   fold    : the three dx taps folded into the MMA's N (thin, wide layers; needs P = 32, a
    //             patch row per warp), two accumulator sets;
    //   plain   : N = Cout; two accumulator sets (the epilogue of a group overlaps the next
    //             group's MMAs) except for Cout = 256, where one set of T = 2 keeps the weight
    //             stream from L2 halved.
    // Per candidate: T output planes per group = as many accumulators as the set holds, balanced
    // over the z extent; then the weight ring: everything resident if it fits, else G taps per
    // stage (a stage <= 48 KB) and as many stages as fit (>= 2).
    // fold pays for Cout = 32 only: with Cout = 64 the plain kernel (N = 64 at ~48 clk per MMA, a plain
    // epilogue) beats the fold's N = 192 with its three TMEM reads and 64 shuffles per 32 columns
    // (c1.conv0: 1.15 -> 0.74 ms, profiles/r02_notes.md)
    const bool want_fold = g.cout <= 32 && g.W >= 100 && getenv("ISG_CONV_NOFOLD") == nullptr;
    // band-flat tiling for the plain layers (ISG_CONV_NOFLAT=1: the rectangular patches of round 1);
    // ISG_CONV_FLAT_KB bounds the plane slot (default 32 KB)
    const bool want_flat = getenv("ISG_CONV_NOFLAT") == nullptr;
    const long flat_kb = getenv("ISG_CONV_FLAT_KB") ? atol(getenv("ISG_CONV_FLAT_KB")) : 32;
    // pass 0: weights resident -> tile-major issue order, where ONE set of T accumulators already
    //         works as a ring (tile t's epilogue overlaps tiles t+1..), so T can use all of TMEM;
    // pass 1: streamed weights -> weight-stationary order, two accumulator sets.
    long stage_cap = (getenv("ISG_CONV_STAGE_KB") ? atol(getenv("ISG_CONV_STAGE_KB")) : 48) * 1024;
    int force_g = 0;                                    // ISG_CONV_G = "<layer>:<taps per weight stage>,..." (experiments)
    if (const char *ov = getenv("ISG_CONV_G")) {
        char key[16];
        snprintf(key, sizeof key, ",%d:", i);
        std::string lst = std::string(",") + ov + ",";
        const size_t at = lst.find(key);
        if (at != std::string::npos) force_g = atoi(lst.c_str() + at + strlen(key));
    }
    auto place = [&](bool allow_flat, long slot_cap) -> bool {
        bool placed = false;
        for (int pass = 0; pass < 2 && !placed; ++pass) {
            for (int cand = want_fold ? 0 : 1; cand < 2 && !placed; ++cand) {
                const bool fold = cand == 0;
                const int acc = fold ? 3 * g.cout : g.cout;
                const int nsets = pass == 0 ? 1 : ((!fold && g.cout >= 256) ? 1 : 2);
                int tmax = (512 / nsets) / acc;
                if (tmax < 1) continue;
                g.flat = 0;
                if (fold) { g.P = 32; g.Ht = 4; }
                else {
                    choose_tile(g.H, g.W, cblk * 2, g.P, g.Ht);
                    int fp = 0, fr = 0;
                    long ft = 0;
                    const long rect_tiles = (long)((g.W + g.P - 3) / (g.P - 2)) * ((g.H + g.Ht - 1) / g.Ht);
                    if (allow_flat && choose_flat(g.H, g.W, cblk * 2, slot_cap, fp, fr, ft) && ft < rect_tiles) {
                        g.flat = 1;
                        g.P = fp;
                        g.Ht = fr - 2;                       // the TMA box holds Ht + 2 = R padded rows
                    }
                }
                g.plane_rows = (g.Ht + 2) * g.P;
                g.plane_bytes = (g.plane_rows * cblk * 2 + 1023) & ~1023;
                if (tmax > CONV_MAX_T) tmax = CONV_MAX_T;
                if (tmax > g.D) tmax = g.D;
                for (; tmax >= 1 && !placed; --tmax) {
                    const int ngr = (g.D + tmax - 1) / tmax;
                    const int T = (g.D + ngr - 1) / ngr;
                    const long avail = budget - (long)(T + 2) * g.plane_bytes;
                    if (avail <= 0) continue;
                    if ((long)nkb * 27 * tap_bytes <= avail && nkb * 3 <= CONV_MAX_B_STAGES) {
                        g.T = T; g.taps_per_b = 9; g.b_stage_bytes = 9 * tap_bytes; g.n_b_stages = nkb * 3;
                        g.b_resident = 1;
                        placed = true;
                    } else if (pass == 1 && (!fold || T >= 2)) {
                        for (int G : {9, 3, 1}) {
                            if (fold && G == 1) continue;
                            if (force_g && G != force_g) continue;
                            const long sb = (long)G * tap_bytes;
                            if (sb > stage_cap && G > 1) continue;
                            long nb = avail / sb;
                            if (nb < 2) continue;
                            if (nb > 6) nb = 6;
                            g.T = T; g.taps_per_b = G; g.b_stage_bytes = (int)sb; g.n_b_stages = (int)nb;
                            g.b_resident = 0;
                            placed = true;
                            break;
                        }
                    }
                    if (placed) {
                        g.nsets = nsets;
                        g.acc_cols = acc;
                        t.fold = fold ? 1 : 0;
                    }
                    if (pass == 0 && !placed && T <= 2) break;      // resident only pays with a few tiles per group
                }
            }
        }
        return placed;
    };
    // The rectangular placement is the reference point; a band-flat tiling is taken when it needs
    // fewer tiles AND leaves the pipeline as it was (same planes per group, accumulator sets, weight
    // ring granularity, at most one weight stage fewer): measured on B200, flat tiles are worth
    // 7-11 % on the T = 2 layers, while a layer whose bigger plane slots cost it planes per group or
    // weight stages loses more than the tiles gain (c1.conv1 1.46 -> 1.98 ms, c6_0.conv0 1.38 -> 1.83 ms).
    bool placed = place(false, 0);
    if (placed && want_flat && !t.fold) {
        const ConvGeom rect = g;
        const int rect_fold = t.fold;
        bool took = false;
        for (long kb : {flat_kb, 28L, 24L, 20L}) {
            if (kb > flat_kb) continue;
            const bool relax = getenv("ISG_CONV_FLAT_RELAX") != nullptr;      // experiment: accept fewer planes per group
            if (place(true, kb * 1024) && g.flat && !t.fold && (g.T == rect.T || (relax && g.T >= 3)) && g.nsets == rect.nsets &&
                g.b_resident == rect.b_resident && g.taps_per_b == rect.taps_per_b &&
                g.n_b_stages >= rect.n_b_stages - 1) {
                took = true;
                break;
            }
        }
        if (!took) {
            g = rect;
            t.fold = rect_fold;
        }
    }
    // More taps per weight stage are worth more than more stages (every stage boundary is a barrier
    // wait + commit per tile of the group: c1.conv1 G = 9 / 3 / 1 -> 1.30 / 1.61 / 2.92 ms): when the tiling just
    // chosen leaves room for two 9-tap stages of up to 80 KB, take them -- but never at the price of the
    // tiling, the planes per group or the accumulator sets (c5_0.conv0 1.11 -> 1.04 ms; for c2.* the bigger
    // stages would push out the flat tiles, which are worth more).
    if (placed && !g.b_resident && g.taps_per_b < 9 && getenv("ISG_CONV_STAGE_KB") == nullptr) {
        const ConvGeom prev = g;
        const int prev_fold = t.fold;
        stage_cap = 80 * 1024;
        const bool ok2 = place(prev.flat != 0, (long)prev.plane_bytes) && g.flat == prev.flat && g.P == prev.P &&
                         g.Ht == prev.Ht && g.T == prev.T && g.nsets == prev.nsets && t.fold == prev_fold &&
                         !g.b_resident && g.taps_per_b > prev.taps_per_b && g.n_b_stages >= 2;
        stage_cap = 48 * 1024;
        if (!ok2) {
            g = prev;
            t.fold = prev_fold;
        }
    }
    if (!placed) {
        set_error("conv %s: shared memory budget exceeded", CONVS[i].name);
        return false;
    }
    g.Wt = g.P - 2;
    g.tiles_w = (g.W + g.Wt - 1) / g.Wt;
    g.tiles_h = g.flat ? (g.H * g.P + 127) / 128 : (g.H + g.Ht - 1) / g.Ht;
    if (getenv("ISG_CONV_VERBOSE"))
        fprintf(stderr, "conv %-11s %s P=%d Ht=%d tiles/plane=%d T=%d sets=%d resident=%d G=%d stages=%d slot=%d B\n",
                CONVS[i].name, t.fold ? "fold" : (g.flat ? "flat" : "rect"), g.P, g.Ht, g.tiles_w * g.tiles_h, g.T,
                g.nsets, g.b_resident, g.taps_per_b, g.n_b_stages, g.plane_bytes);
    g.dgroups = (g.D + g.T - 1) / g.T;
    g.n_groups = g.N * g.dgroups * g.tiles_h * g.tiles_w;
    g.out_mode = out_mode;
    { const char *dbg = getenv("ISG_CONV_DEBUG"); g.debug = dbg ? atoi(dbg) : 0; }
    g.out = out;
    g.stats = p->stats[i];
    g.sched = p->sched[i];
    t.smem = conv_smem_bytes(g);
    t.grid = g.n_groups < num_sms() ? g.n_groups : num_sms();
    if (!make_act_map(&t.tmA0, src0, c0, g.W, g.H, g.D, g.N, cblk, g.P, g.Ht)) return false;
    if (c1 > 0) {
        if (!make_act_map(&t.tmA1, src1, c1, g.W, g.H, g.D, g.N, cblk, g.P, g.Ht)) return false;
    } else {
        t.tmA1 = t.tmA0;
    }
    if (!make_w_map(&t.tmB, p->packed + p->L.w[i], cin, g.cout, cblk, g.taps_per_b)) return false;
    return true;
}

struct ProfScope {
    isg_unet_plan *p;
    cudaStream_t st;
    size_t slot;
    ProfScope(isg_unet_plan *plan, int kind, cudaStream_t s) : p(plan), st(s), slot((size_t)-1) {
        // level 1: the TMA-fed tcgen05 convolutions and the whole forward (bench.py's live roofline);
        // level 2: every launch (scripts/time_unet.py)
        if (!p->profiling || (kind == 2 && p->profiling < 2)) return;
        if (p->ev_used + 2 > p->ev.size()) {
            cudaEvent_t a, b;
            cudaEventCreate(&a);
            cudaEventCreate(&b);
            p->ev.push_back(a);
            p->ev.push_back(b);
            p->ev_kind.push_back(kind);
        }
        slot = p->ev_used;
        p->ev_kind[slot / 2] = kind;
        p->ev_used += 2;
        cudaEventRecord(p->ev[slot], st);
    }
    ~ProfScope() {
        if (slot != (size_t)-1) cudaEventRecord(p->ev[slot + 1], st);
    }
};

template <int CBLK, int G, bool FOLD>
static int launch_tc_inst(const TcLayer &t, cudaStream_t st) {
    ISG_CUDA(cudaFuncSetAttribute(conv3d_tc_kernel<CBLK, G, FOLD>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)t.smem));
    int grid = t.grid;
    const int avail = num_sms() - post_sms();
    if (grid > avail) grid = avail;
    conv3d_tc_kernel<CBLK, G, FOLD><<<grid, CONV_THREADS, t.smem, st>>>(t.tmA0, t.tmA1, t.tmB, t.g);
    ISG_LAUNCHED();
    return ISG_OK;
}

static int launch_tc(const TcLayer &t, cudaStream_t st) {
    if (t.zring == 2) {
        int grid = t.grid;
        const int avail = num_sms() - post_sms();
        if (grid > avail) grid = avail;
        // ISG_Z32_MODE=ring: the round-1 kernel (three accumulators read per output plane); ISG_Z32_EPI=4|8 epilogue warps
        static const bool ring = getenv("ISG_Z32_MODE") != nullptr && strcmp(getenv("ISG_Z32_MODE"), "ring") == 0;
        static const int epi = getenv("ISG_Z32_EPI") ? atoi(getenv("ISG_Z32_EPI")) : 8;
#define ISG_Z32_LAUNCH(KERNEL, EPI)                                                                              \
    do {                                                                                                         \
        ISG_CUDA(cudaFuncSetAttribute(KERNEL<EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)t.smem));   \
        KERNEL<EPI><<<grid, 128 + 32 * EPI, t.smem, st>>>(t.tmA0, t.tmB, t.z32);                                 \
    } while (0)
        if (ring || t.z32.D > ZS_NB) {                      // one 32-column TMEM block per output plane of a column
            if (epi == 8) ISG_Z32_LAUNCH(conv3d_zring32_kernel, 8);
            else ISG_Z32_LAUNCH(conv3d_zring32_kernel, 4);
        } else {
            if (epi == 8) ISG_Z32_LAUNCH(conv3d_zslide32_kernel, 8);
            else ISG_Z32_LAUNCH(conv3d_zslide32_kernel, 4);
        }
#undef ISG_Z32_LAUNCH
        ISG_LAUNCHED();
        return ISG_OK;
    }
    if (t.zring) {
        ISG_CUDA(cudaFuncSetAttribute(conv3d_zring_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)t.smem));
        int grid = t.grid;
        const int avail = 2 * (num_sms() - post_sms());
        if (grid > avail) grid = avail;
        conv3d_zring_kernel<<<grid, ZR_THREADS, t.smem, st>>>(t.tmA0, t.tmA1, t.tmB, t.zg);
        ISG_LAUNCHED();
        return ISG_OK;
    }
    const int G = t.g.taps_per_b;
    if (t.fold) {
        if (t.cblk == 64) return G == 9 ? launch_tc_inst<64, 9, true>(t, st) : launch_tc_inst<64, 3, true>(t, st);
        return G == 9 ? launch_tc_inst<32, 9, true>(t, st) : launch_tc_inst<32, 3, true>(t, st);
    }
    if (t.cblk == 64) {
        switch (G) {
            case 9: return launch_tc_inst<64, 9, false>(t, st);
            case 3: return launch_tc_inst<64, 3, false>(t, st);
            default: return launch_tc_inst<64, 1, false>(t, st);
        }
    }
    switch (G) {
        case 9: return launch_tc_inst<32, 9, false>(t, st);
        case 3: return launch_tc_inst<32, 3, false>(t, st);
        default: return launch_tc_inst<32, 1, false>(t, st);
    }
}

static int launch_thin(const ThinArgs &a, cudaStream_t st) {
    constexpr size_t smem = thin_smem_bytes();
    static int per_sm[16] = {0};
    int dev = 0;
    ISG_CUDA(cudaGetDevice(&dev));
    if (per_sm[dev & 15] == 0) {
        ISG_CUDA(cudaFuncSetAttribute(conv_in_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)smem));
        ISG_CUDA(cudaFuncSetAttribute(conv_in_tc_kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                                      cudaSharedmemCarveoutMaxShared));
        // resident CTAs per SM from the kernel's own footprint: shared memory (+1 KB the hardware
        // reserves per CTA), registers, and 32 of the SM's 512 TMEM columns each
        cudaFuncAttributes fa;
        ISG_CUDA(cudaFuncGetAttributes(&fa, conv_in_tc_kernel));
        int smem_sm = 0, regs_sm = 0;
        ISG_CUDA(cudaDeviceGetAttribute(&smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev));
        ISG_CUDA(cudaDeviceGetAttribute(&regs_sm, cudaDevAttrMaxRegistersPerMultiprocessor, dev));
        int occ = (int)((size_t)smem_sm / (smem + 1024 + fa.sharedSizeBytes));
        const int by_regs = regs_sm / (((fa.numRegs + 7) & ~7) * THIN_THREADS);
        if (by_regs < occ) occ = by_regs;
        per_sm[dev & 15] = occ < 1 ? 1 : (occ > 16 ? 16 : occ);
    }
    const long long units = (long long)a.N * a.D * ((a.H + 3) / 4) * ((a.W + 31) / 32);
    if (units >= (1ll << 31)) {
        set_error("thin conv: %lld units exceed the 32-bit unit index", units);
        return ISG_ERR_ARG;
    }
    long long grid = (long long)num_sms() * per_sm[dev & 15];
    if (grid > units) grid = units;
    conv_in_tc_kernel<<<(unsigned)grid, THIN_THREADS, smem, st>>>(a);
    ISG_LAUNCHED();
    return ISG_OK;
}

// grid = (blocks per chunk, N): ~48 blocks per SM over ALL chunks -- enough waves that kernels with
// 2..8 resident blocks per SM balance, few enough that the per-block prologue (the chunk's
// BatchNorm coefficient table) is not paid once per 256-thread sliver of every chunk
static inline dim3 egrid(size_t work, int N) {
    size_t b = (work + 255) / 256;
    const size_t cap = ((size_t)num_sms() * 48 + (size_t)N - 1) / (size_t)N;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return dim3((unsigned)b, (unsigned)N);
}

// run the network up to and including conv `stop` (17 = everything incl. placement)
static int forward(isg_unet_plan *p, const float *frame, float *feats, int stop, cudaStream_t st,
                   const float *norm_max = nullptr) {
    const int N = p->N;
    auto run_tc = [&](int i) {
        ProfScope ps(p, 0, st);
        return launch_tc(p->tc[i], st);
    };
    ProfScope whole(p, 1, st);
    {
        // every pass uploads its tables: plans may share the workspace, so whatever another plan's
        // pass left at these addresses is not ours (three small asynchronous copies from pinned memory)
        if (p->tabs_dirty < 0) {
            set_error("isg_unet_forward_chunks: no chunk tables set (isg_unet_plan_set_chunks)");
            return ISG_ERR_ARG;
        }
        const size_t nb = sizeof(int) * 3 * (size_t)N;
        static const bool once = getenv("ISG_TABLES_ONCE") != nullptr;      // diagnosis only (unsafe with shared workspaces)
        if (!(once && p->tabs_dirty == 0)) {
        ISG_CUDA(cudaMemcpyAsync(p->starts, p->tabs_host, 3 * nb, cudaMemcpyHostToDevice, st));
        ISG_CUDA(cudaEventRecord(p->tabs_ev, st));
        p->tabs_ev_pending = 1;
        p->tabs_dirty = 0;
        }
    }
    const unsigned char *pk = p->packed;
    const PackLayout &L = p->L;
    auto G = [&](int i) { return reinterpret_cast<const float *>(pk + L.gamma[i]); };
    auto B = [&](int i) { return reinterpret_cast<const float *>(pk + L.beta[i]); };
    auto E = [&](int i) { return reinterpret_cast<const float *>(pk + L.eps[i]); };
    auto vox = [&](int l) { return (size_t)p->D[l] * p->H[l] * p->W[l]; };
    ISG_CUDA(cudaMemsetAsync(p->stats_all, 0, p->stats_bytes, st));
    // ---- c0.conv0 (1 -> 32, im2col on tensor cores, straight from the frame) ----
    {
        ProfScope ps(p, 2, st);
        ThinArgs a{};
        a.src = frame; a.norm_max = norm_max; a.starts = p->starts; a.Y = p->Y; a.X = p->X;
        a.wgt = reinterpret_cast<const float *>(pk + L.w[0]);
        a.out = p->raw[0]; a.stats = p->stats[0];
        a.N = N; a.D = p->D[0]; a.H = p->H[0]; a.W = p->W[0];
        int rc = launch_thin(a, st);
        if (rc) return rc;
    }
    if (stop == 0) return ISG_OK;
    static const int CH[5] = {32, 64, 128, 256, 256};
    // ---- encoder ----
    for (int l = 0; l < 5; ++l) {
        const int i0 = 2 * l, i1 = 2 * l + 1;
        if (l > 0) {
            int rc = run_tc(i0);
            if (rc) return rc;
            if (stop == i0) return ISG_OK;
        }
        {
            ProfScope ps(p, 2, st);
            bn_relu_kernel<<<egrid(vox(l) * CH[l] / 8, N), 256, 0, st>>>(p->raw[l], p->act[l], p->stats[i0],
                                                                         G(i0), B(i0), E(i0), CH[l], vox(l));
            ISG_LAUNCHED();
        }
        int rc = run_tc(i1);
        if (rc) return rc;
        if (stop == i1) return ISG_OK;
        if (l < 4) {
            ProfScope ps(p, 2, st);
            const size_t work = vox(l + 1) * CH[l] / 8;
            if (l == 3)
                bn_relu_pool_kernel<2><<<egrid(work, N), 256, 0, st>>>(
                    p->raw[l], p->skip[l], p->pooled[l + 1], p->stats[i1], G(i1), B(i1), E(i1), CH[l], p->D[l],
                    p->H[l], p->W[l], p->D[l + 1], p->H[l + 1], p->W[l + 1]);
            else
                bn_relu_pool_kernel<1><<<egrid(work, N), 256, 0, st>>>(
                    p->raw[l], p->skip[l], p->pooled[l + 1], p->stats[i1], G(i1), B(i1), E(i1), CH[l], p->D[l],
                    p->H[l], p->W[l], p->D[l + 1], p->H[l + 1], p->W[l + 1]);
            ISG_LAUNCHED();
        }
    }
    // ---- decoder: level 4 -> 3 -> 2 -> 1 -> 0 ----
    for (int u = 0; u < 4; ++u) {
        const int lc = 4 - u, lf = 3 - u;           // coarse (source) and fine (target) level
        const int src_conv = u == 0 ? 9 : 9 + 2 * u; // conv whose raw output feeds this upsample
        const int C = UP_C[u];
        const float *uw = reinterpret_cast<const float *>(pk + L.up_w[u]);
        const float *ub = reinterpret_cast<const float *>(pk + L.up_b[u]);
        const size_t work = vox(lc) * C / 8;
        const int off = u == 3 ? 1 : 0;
        ProfScope *ups = new ProfScope(p, 2, st);
        if (u == 0)
            bn_relu_up_kernel<2><<<egrid(work, N), 256, (size_t)(3 + 8) * C * sizeof(float), st>>>(
                p->raw[lc], p->up[lf], p->stats[src_conv], G(src_conv), B(src_conv), E(src_conv), uw, ub, C, p->D[lc],
                p->H[lc], p->W[lc], p->D[lf], p->H[lf], p->W[lf], off);
        else
            bn_relu_up_kernel<1><<<egrid(work, N), 256, (size_t)(3 + 4) * C * sizeof(float), st>>>(
                p->raw[lc], p->up[lf], p->stats[src_conv], G(src_conv), B(src_conv), E(src_conv), uw, ub, C, p->D[lc],
                p->H[lc], p->W[lc], p->D[lf], p->H[lf], p->W[lf], off);
        delete ups;
        ISG_LAUNCHED();
        const int i0 = 10 + 2 * u, i1 = i0 + 1;
        int rc = run_tc(i0);                        // reads [up, skip] of level lf
        if (rc) return rc;
        if (stop == i0) return ISG_OK;
        if (u < 3) {
            const int Cm = CONVS[i0].cout;
            {
                ProfScope ps(p, 2, st);
                bn_relu_kernel<<<egrid(vox(lf) * Cm / 8, N), 256, 0, st>>>(p->raw[lf], p->act[lf], p->stats[i0],
                                                                           G(i0), B(i0), E(i0), Cm, vox(lf));
                ISG_LAUNCHED();
            }
            rc = run_tc(i1);
            if (rc) return rc;
            if (stop == i1) return ISG_OK;
        }
    }
    // ---- c8_0.conv1 (5 -> 5) + BN + sigmoid + placement ----
    {
        ProfScope ps(p, 2, st);
        ZoutArgs a{};
        ZringGeom &z = a.g;
        z.N = N; z.D = p->D[0]; z.H = p->H[0]; z.W = p->W[0];
        z.tiles_w = (z.W + ZR_WT - 1) / ZR_WT;
        z.tiles_h = (z.H + ZR_HT - 1) / ZR_HT;
        z.n_cols = z.N * z.tiles_h * z.tiles_w;
        z.out = p->raw9; z.stats = p->stats[17]; z.sched = p->sched[17];
        a.src = p->raw8; a.stats_in = p->stats[16]; a.gamma_in = G(16); a.beta_in = B(16); a.eps_in = E(16);
        a.wgt = reinterpret_cast<const float *>(pk + L.w[17]);
        ISG_CUDA(cudaFuncSetAttribute(conv_out_zring_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)zout_smem_bytes()));
        int grid = 2 * (num_sms() - post_sms());
        if (grid > z.n_cols) grid = z.n_cols;
        conv_out_zring_kernel<<<grid, ZO_THREADS, zout_smem_bytes(), st>>>(a);
        ISG_LAUNCHED();
    }
    if (feats == nullptr) return ISG_OK;
    ISG_CUDA(cudaMemcpyFromSymbolAsync(p->overflow_host, g_unet_overflow, sizeof(unsigned int), 0,
                                       cudaMemcpyDeviceToHost, st));
    {
        ProfScope ps(p, 2, st);
        place_kernel<<<egrid(vox(0), N), 256, 0, st>>>(p->raw9, p->stats[17], G(17), B(17), E(17), p->starts,
                                                       p->crop_lo, p->crop_hi, feats, p->Z, p->Y, p->X,
                                                       p->D[0], p->H[0], p->W[0]);
        ISG_LAUNCHED();
    }
    return ISG_OK;
}

}  // namespace isg

using namespace isg;

// Geometry of the band-flat tiling for an (H, W) plane with rows of `row_bytes` and plane slots of at most
// `max_slot_bytes` (host only, no device needed: the CPU tests check that every output voxel lies in
// exactly one tile and that no tap leaves the plane slot).  out4 = {P, R, tiles per plane, slot bytes}.
extern "C" int isg_debug_flat_tiling(int H, int W, int row_bytes, int64_t max_slot_bytes, int64_t *out4) {
    ISG_REQUIRE(out4 && H > 0 && W > 0 && row_bytes > 0, ISG_ERR_ARG, "isg_debug_flat_tiling: bad argument");
    int P = 0, R = 0;
    long tiles = 0;
    if (!isg::choose_flat(H, W, row_bytes, (long)max_slot_bytes, P, R, tiles)) {
        isg::set_error("isg_debug_flat_tiling: no band fits %lld bytes", (long long)max_slot_bytes);
        return ISG_ERR_ARG;
    }
    out4[0] = P; out4[1] = R; out4[2] = tiles;
    out4[3] = (((int64_t)R * P * row_bytes) + 1023) & ~(int64_t)1023;
    return ISG_OK;
}

extern "C" size_t isg_unet_packed_weight_bytes(void) { return pack_layout().total; }

extern "C" int isg_unet_weights_pack(const void *const *tensors, int n_tensors, void *packed, void *stream) {
    ISG_REQUIRE(tensors && packed, ISG_ERR_ARG, "isg_unet_weights_pack: null pointer");
    ISG_REQUIRE(n_tensors == 134, ISG_ERR_ARG,
                "isg_unet_weights_pack: expected the 134 tensors of UNet(1,5).state_dict(), got %d", n_tensors);
    cudaStream_t st = (cudaStream_t)stream;
    const PackLayout L = pack_layout();
    unsigned char *pk = (unsigned char *)packed;
    ISG_CUDA(cudaMemsetAsync(packed, 0, L.total, st));
    for (int i = 0; i < 18; ++i) {
        const int mi = i / 2, ci = i % 2;
        const float *w = (const float *)tensors[14 * mi + 2 * ci];
        const float *gam = (const float *)tensors[14 * mi + 4 + 5 * ci];
        const float *bet = (const float *)tensors[14 * mi + 5 + 5 * ci];
        ISG_REQUIRE(w && gam && bet, ISG_ERR_ARG, "isg_unet_weights_pack: missing tensor for %s", CONVS[i].name);
        float *inv_s = (float *)(pk + L.inv_s[i]);
        // ISG_NO_WEIGHT_PRESCALE=1 (diagnosis / the overflow test only) packs the filters as they are
        conv_prescale_kernel<<<CONVS[i].cout, 256, 0, st>>>(w, CONVS[i].cin, inv_s, (float *)(pk + L.eps[i]),
                                                            getenv("ISG_NO_WEIGHT_PRESCALE") == nullptr ? 1 : 0);
        ISG_LAUNCHED();
        if ((i == 1 || i == 15) && getenv("ISG_NO_ZRING32") == nullptr)
            pack_conv_w_zring32_kernel<<<64, 256, 0, st>>>(w, inv_s, (__half *)(pk + L.w[i]));
        else if (i == 16)
            pack_conv_w_zring_kernel<<<64, 256, 0, st>>>(w, inv_s, (__half *)(pk + L.w[i]), CONVS[i].cout, CONVS[i].cin);
        else if (is_tc(i))
            pack_conv_w_kernel<<<256, 256, 0, st>>>(w, inv_s, (__half *)(pk + L.w[i]), CONVS[i].cout, cout_pad(i), CONVS[i].cin);
        else
            pack_conv_w_f32_kernel<<<16, 256, 0, st>>>(w, inv_s, (float *)(pk + L.w[i]), CONVS[i].cout, CONVS[i].cin);
        ISG_LAUNCHED();
        copy_f32_kernel<<<1, 256, 0, st>>>(gam, (float *)(pk + L.gamma[i]), CONVS[i].cout);
        ISG_LAUNCHED();
        copy_f32_kernel<<<1, 256, 0, st>>>(bet, (float *)(pk + L.beta[i]), CONVS[i].cout);
        ISG_LAUNCHED();
    }
    for (int u = 0; u < 4; ++u) {
        const float *w = (const float *)tensors[126 + 2 * u];
        const float *b = (const float *)tensors[126 + 2 * u + 1];
        ISG_REQUIRE(w && b, ISG_ERR_ARG, "isg_unet_weights_pack: missing up%d tensors", u);
        copy_f32_kernel<<<4, 256, 0, st>>>(w, (float *)(pk + L.up_w[u]), UP_C[u] * UP_KZ[u] * 4);
        ISG_LAUNCHED();
        copy_f32_kernel<<<1, 256, 0, st>>>(b, (float *)(pk + L.up_b[u]), UP_C[u]);
        ISG_LAUNCHED();
    }
    return ISG_OK;
}

extern "C" size_t isg_unet_workspace_bytes(int n_chunks, int cz, int cy, int cx) {
    if (n_chunks <= 0) return 0;
    isg_unet_plan p;
    p.N = n_chunks;
    if (!level_dims(&p, cz, cy, cx)) return 0;
    Carver cv(nullptr, 0);
    plan_carve(&p, cv);
    return cv.off + 1024;
}

extern "C" isg_unet_plan *isg_unet_plan_create(const void *packed_weights, int n_chunks, int cz, int cy,
                                               int cx, int64_t z, int64_t y, int64_t x,
                                               const int32_t *starts_host, const int32_t *crop_lo_host,
                                               const int32_t *crop_hi_host, void *workspace,
                                               size_t workspace_bytes) {
    if (!packed_weights || !workspace || n_chunks <= 0 ||
        ((starts_host || crop_lo_host || crop_hi_host) && !(starts_host && crop_lo_host && crop_hi_host))) {
        set_error("isg_unet_plan_create: bad argument");
        return nullptr;
    }
    if (!encode_fn()) {
        set_error("isg_unet_plan_create: cuTensorMapEncodeTiled is not available from the driver");
        return nullptr;
    }
    isg_unet_plan *p = new isg_unet_plan();
    p->profiling = 0;
    p->ev_used = 0;
    p->N = n_chunks;
    p->Z = (int)z; p->Y = (int)y; p->X = (int)x;
    p->packed = (const unsigned char *)packed_weights;
    p->L = pack_layout();
    // Measured on B200: the UMMA operand fetch swizzles on absolute shared-memory address
    // bits, so a descriptor whose start address is advanced by whole rows needs base_offset 0
    // (profiles/r01_notes.md).  ISG_CONV_BASE_OFFSET=1 re-enables the (wrong) alternative for
    // diagnosis only.
    const char *bm = getenv("ISG_CONV_BASE_OFFSET");
    p->base_off_mode = bm ? atoi(bm) : 0;
    if (!level_dims(p, cz, cy, cx)) {
        set_error("chunk shape (%d,%d,%d) is not valid for this U-Net: z must be even and y/x must survive "
                  "four poolings and the decoder crops (e.g. 10,256,256)", cz, cy, cx);
        delete p;
        return nullptr;
    }
    Carver cv(workspace, workspace_bytes);
    plan_carve(p, cv);
    if (!cv.ok) {
        set_error("isg_unet_plan_create: workspace too small (%zu < %zu)", workspace_bytes, cv.off);
        delete p;
        return nullptr;
    }
    p->cz = cz; p->cy = cy; p->cx = cx;
    p->tabs_host = nullptr;
    p->overflow_host = nullptr;
    p->tabs_ev = nullptr;
    p->tabs_ev_pending = 0;
    p->tabs_dirty = -1;                                   // nothing set yet
    if (cudaHostAlloc((void **)&p->overflow_host, sizeof(unsigned int), cudaHostAllocDefault) == cudaSuccess)
        *p->overflow_host = 0;
    if (!p->overflow_host ||
        cudaHostAlloc((void **)&p->tabs_host, sizeof(int) * 9 * (size_t)n_chunks, cudaHostAllocDefault) != cudaSuccess ||
        cudaEventCreateWithFlags(&p->tabs_ev, cudaEventDisableTiming) != cudaSuccess) {
        set_error("isg_unet_plan_create: pinned staging for the chunk tables: %s",
                  cudaGetErrorString(cudaGetLastError()));
        isg_unet_plan_destroy(p);
        return nullptr;
    }
    if (starts_host && isg_unet_plan_set_chunks(p, starts_host, crop_lo_host, crop_hi_host) != ISG_OK) {
        isg_unet_plan_destroy(p);
        return nullptr;
    }
    bool ok = true;
    // encoder
    ok = ok && setup_tc_layer(p, 1, p->act[0], 32, nullptr, 0, p->raw[0], 0);
    ok = ok && setup_tc_layer(p, 2, p->pooled[1], 32, nullptr, 0, p->raw[1], 0);
    ok = ok && setup_tc_layer(p, 3, p->act[1], 64, nullptr, 0, p->raw[1], 0);
    ok = ok && setup_tc_layer(p, 4, p->pooled[2], 64, nullptr, 0, p->raw[2], 0);
    ok = ok && setup_tc_layer(p, 5, p->act[2], 128, nullptr, 0, p->raw[2], 0);
    ok = ok && setup_tc_layer(p, 6, p->pooled[3], 128, nullptr, 0, p->raw[3], 0);
    ok = ok && setup_tc_layer(p, 7, p->act[3], 256, nullptr, 0, p->raw[3], 0);
    ok = ok && setup_tc_layer(p, 8, p->pooled[4], 256, nullptr, 0, p->raw[4], 0);
    ok = ok && setup_tc_layer(p, 9, p->act[4], 256, nullptr, 0, p->raw[4], 0);
    // decoder: concat order is [upsampled, skip] (unet.py:332,337,341,345)
    ok = ok && setup_tc_layer(p, 10, p->up[3], 256, p->skip[3], 256, p->raw[3], 0);
    ok = ok && setup_tc_layer(p, 11, p->act[3], 128, nullptr, 0, p->raw[3], 0);
    ok = ok && setup_tc_layer(p, 12, p->up[2], 128, p->skip[2], 128, p->raw[2], 0);
    ok = ok && setup_tc_layer(p, 13, p->act[2], 64, nullptr, 0, p->raw[2], 0);
    ok = ok && setup_tc_layer(p, 14, p->up[1], 64, p->skip[1], 64, p->raw[1], 0);
    ok = ok && setup_tc_layer(p, 15, p->act[1], 32, nullptr, 0, p->raw[1], 0);
    ok = ok && setup_tc_layer(p, 16, p->up[0], 32, p->skip[0], 32, p->raw8, 1);
    if (!ok) {
        isg_unet_plan_destroy(p);
        return nullptr;
    }
    double macs = 0, tc_macs = 0;
    for (int i = 0; i < 18; ++i) {
        const int l = CONVS[i].level;
        const double m = 27.0 * CONVS[i].cin * CONVS[i].cout * p->D[l] * p->H[l] * p->W[l];
        macs += m;
        if (is_tc(i)) tc_macs += m;
    }
    p->tc_flops = 2.0 * tc_macs * n_chunks;
    for (int u = 0; u < 4; ++u) {
        const int lc = 4 - u;
        macs += (double)UP_C[u] * UP_KZ[u] * 4 * p->D[lc] * p->H[lc] * p->W[lc];
    }
    p->flops = 2.0 * macs * n_chunks;
    return p;
}

extern "C" void isg_unet_plan_destroy(isg_unet_plan *plan) {
    if (!plan) return;
    for (cudaEvent_t e : plan->ev) cudaEventDestroy(e);
    if (plan->tabs_ev) {
        if (plan->tabs_ev_pending) cudaEventSynchronize(plan->tabs_ev);   // an upload may still read tabs_host
        cudaEventDestroy(plan->tabs_ev);
    }
    if (plan->tabs_host) cudaFreeHost(plan->tabs_host);
    if (plan->overflow_host) cudaFreeHost(plan->overflow_host);
    delete plan;
}

extern "C" int isg_unet_plan_overflowed(const isg_unet_plan *plan) {
    return plan && plan->overflow_host ? (int)*reinterpret_cast<volatile unsigned int *>(plan->overflow_host) : 0;
}

extern "C" int isg_unet_plan_clear_overflow(isg_unet_plan *plan, void *stream) {
    ISG_REQUIRE(plan, ISG_ERR_ARG, "isg_unet_plan_clear_overflow: null plan");
    const unsigned int zero = 0;
    *plan->overflow_host = 0;
    ISG_CUDA(cudaMemcpyToSymbolAsync(g_unet_overflow, &zero, sizeof(zero), 0, cudaMemcpyHostToDevice,
                                     (cudaStream_t)stream));
    return ISG_OK;
}

extern "C" int isg_unet_plan_set_chunks(isg_unet_plan *plan, const int32_t *starts_host,
                                        const int32_t *crop_lo_host, const int32_t *crop_hi_host) {
    ISG_REQUIRE(plan && starts_host && crop_lo_host && crop_hi_host, ISG_ERR_ARG,
                "isg_unet_plan_set_chunks: null pointer");
    const int N = plan->N;
    for (int n = 0; n < N; ++n) {
        const int32_t *s = starts_host + 3 * n, *lo = crop_lo_host + 3 * n, *hi = crop_hi_host + 3 * n;
        const int64_t ext[3] = {plan->Z, plan->Y, plan->X};
        const int chk[3] = {plan->cz, plan->cy, plan->cx};
        for (int a = 0; a < 3; ++a) {
            ISG_REQUIRE(s[a] >= 0 && s[a] + chk[a] <= ext[a], ISG_ERR_ARG, "chunk %d starts outside the frame", n);
            ISG_REQUIRE(lo[a] >= 0 && lo[a] <= hi[a] && hi[a] <= chk[a], ISG_ERR_ARG,
                        "chunk %d: crop [%d,%d) outside the chunk extent %d", n, lo[a], hi[a], chk[a]);
        }
    }
    const size_t nb = sizeof(int) * 3 * (size_t)N;
    if (plan->tabs_dirty >= 0 && memcmp(plan->tabs_host, starts_host, nb) == 0 &&
        memcmp(plan->tabs_host + 3 * N, crop_lo_host, nb) == 0 &&
        memcmp(plan->tabs_host + 6 * N, crop_hi_host, nb) == 0)
        return ISG_OK;                                     // the pinned copy already holds these tables
    if (plan->tabs_ev_pending) {                           // an enqueued upload still reads the pinned copy
        ISG_CUDA(cudaEventSynchronize(plan->tabs_ev));
        plan->tabs_ev_pending = 0;
    }
    memcpy(plan->tabs_host, starts_host, nb);
    memcpy(plan->tabs_host + 3 * N, crop_lo_host, nb);
    memcpy(plan->tabs_host + 6 * N, crop_hi_host, nb);
    plan->tabs_dirty = 1;
    return ISG_OK;
}

extern "C" int isg_unet_plan_profile(isg_unet_plan *plan, int enable) {
    ISG_REQUIRE(plan, ISG_ERR_ARG, "isg_unet_plan_profile: null plan");
    plan->profiling = enable < 0 ? 0 : (enable > 2 ? 2 : enable);
    plan->ev_used = 0;
    return ISG_OK;
}

// Sum the event-timed launches recorded since profiling was (re-)enabled.  The caller
// must have synchronised the stream.  out[0] = ms in tcgen05 convolutions, out[1] = number
// of tcgen05 conv launches, out[2] = ms of whole forward passes, out[3] = number of
// forward passes, out[4] = algorithmic FLOPs of the tcgen05 convolutions of one forward.
extern "C" int isg_unet_plan_profile_read(isg_unet_plan *plan, double *out) {
    ISG_REQUIRE(plan && out, ISG_ERR_ARG, "isg_unet_plan_profile_read: null pointer");
    double tc_ms = 0, all_ms = 0;
    int n_tc = 0, n_fw = 0;
    for (size_t s = 0; s + 1 < plan->ev_used; s += 2) {
        float ms = 0;
        ISG_CUDA(cudaEventElapsedTime(&ms, plan->ev[s], plan->ev[s + 1]));
        if (plan->ev_kind[s / 2] == 0) { tc_ms += ms; ++n_tc; }
        else if (plan->ev_kind[s / 2] == 1) { all_ms += ms; ++n_fw; }
    }
    out[0] = tc_ms; out[1] = n_tc; out[2] = all_ms; out[3] = n_fw; out[4] = plan->tc_flops;
    return ISG_OK;
}

// Per-launch times of the forward passes recorded since profiling was enabled, in launch order
// (35 per forward: conv_in, then alternating companions and convolutions, conv_out, place).
// Returns the number of launches written (<= cap); kind_out[i]: 0 tcgen05 conv (TMA-fed), 2 other.
extern "C" int isg_unet_plan_profile_launches(isg_unet_plan *plan, double *ms_out, int *kind_out, int cap) {
    if (!plan || !ms_out) return 0;
    int n = 0;
    for (size_t s = 0; s + 1 < plan->ev_used && n < cap; s += 2) {
        if (plan->ev_kind[s / 2] == 1) continue;                  // the whole-forward scope
        float ms = 0;
        if (cudaEventElapsedTime(&ms, plan->ev[s], plan->ev[s + 1]) != cudaSuccess) break;
        ms_out[n] = ms;
        if (kind_out) kind_out[n] = plan->ev_kind[s / 2];
        ++n;
    }
    return n;
}

// Start / end of every recorded WHOLE forward pass in ms since the first one started (diagnosis of the
// gaps between consecutive forward passes on a stream).  out[2i], out[2i+1]; returns the number of passes.
extern "C" int isg_unet_plan_profile_timeline(isg_unet_plan *plan, double *out, int cap) {
    if (!plan || !out) return 0;
    int n = 0;
    cudaEvent_t first = nullptr;
    for (size_t s = 0; s + 1 < plan->ev_used && n < cap; s += 2) {
        if (plan->ev_kind[s / 2] != 1) continue;
        if (!first) first = plan->ev[s];
        float a = 0, b = 0;
        if (cudaEventElapsedTime(&a, first, plan->ev[s]) != cudaSuccess) break;
        if (cudaEventElapsedTime(&b, first, plan->ev[s + 1]) != cudaSuccess) break;
        out[2 * n] = a;
        out[2 * n + 1] = b;
        ++n;
    }
    return n;
}

extern "C" double isg_unet_plan_flops(const isg_unet_plan *plan) { return plan ? plan->flops : 0.0; }

extern "C" int isg_unet_forward_chunks(isg_unet_plan *plan, const float *frame, float *feats, void *stream) {
    ISG_REQUIRE(plan && frame && feats, ISG_ERR_ARG, "isg_unet_forward_chunks: null pointer");
    ISG_REQUIRE(!isg_unet_plan_overflowed(plan), ISG_ERR_OVERFLOW,
                "isg_unet_forward_chunks: an earlier forward pass of this plan overflowed the fp16 range of the "
                "pre-BatchNorm activations (its feature volume is invalid); isg_unet_plan_clear_overflow re-arms");
    return forward(plan, frame, feats, 17, (cudaStream_t)stream);
}

extern "C" int isg_unet_forward_chunks_norm(isg_unet_plan *plan, const float *frame, const float *norm_max,
                                            float *feats, void *stream) {
    ISG_REQUIRE(plan && frame && feats && norm_max, ISG_ERR_ARG, "isg_unet_forward_chunks_norm: null pointer");
    ISG_REQUIRE(!isg_unet_plan_overflowed(plan), ISG_ERR_OVERFLOW,
                "isg_unet_forward_chunks_norm: an earlier forward pass of this plan overflowed the fp16 range");
    return forward(plan, frame, feats, 17, (cudaStream_t)stream, norm_max);
}

extern "C" int isg_unet_debug_activation(isg_unet_plan *plan, const float *frame, const char *name,
                                         int chunk, float *out, int64_t out_elems, void *stream) {
    ISG_REQUIRE(plan && frame && name && out, ISG_ERR_ARG, "isg_unet_debug_activation: null pointer");
    ISG_REQUIRE(chunk >= 0 && chunk < plan->N, ISG_ERR_ARG, "bad chunk index");
    int idx = -1;
    for (int i = 0; i < 18; ++i)
        if (strcmp(name, CONVS[i].name) == 0) idx = i;
    ISG_REQUIRE(idx >= 0, ISG_ERR_ARG, "unknown layer '%s' (use e.g. c0.conv0 ... c8_0.conv1)", name);
    cudaStream_t st = (cudaStream_t)stream;
    int rc = forward(plan, frame, nullptr, idx, st);
    if (rc) return rc;
    const int l = CONVS[idx].level;
    const size_t vox = (size_t)plan->D[l] * plan->H[l] * plan->W[l];
    const int C = CONVS[idx].cout;
    ISG_REQUIRE((int64_t)(vox * C) == out_elems, ISG_ERR_ARG, "out must hold %zu floats", vox * C);
    const void *src;
    int is_f32 = 0, cstride = cout_pad(idx);
    if (idx == 16) { src = plan->raw8 + (size_t)chunk * vox * 8; cstride = 8; }
    else if (idx == 17) { src = plan->raw9 + (size_t)chunk * vox * 8; cstride = 8; }
    else src = plan->raw[l] + (size_t)chunk * vox * cstride;
    debug_to_ncdhw_kernel<<<num_sms() * 4, 256, 0, st>>>(
        src, is_f32, cstride, C, vox, reinterpret_cast<const float *>(plan->packed + plan->L.inv_s[idx]), out);
    ISG_LAUNCHED();
    return ISG_OK;
}
