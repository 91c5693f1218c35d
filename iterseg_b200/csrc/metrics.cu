// Label-permutation-invariant assessment of a segmentation against a ground truth on the device:
// variation of information (skimage.metrics.variation_of_information, log base 2) and the
// IoU-matched object counts TP / FP / FN (src/iterseg/metrics.py:107, :205-227).
//
// contingency table = sort of the (gt << 32 | seg) voxel pairs + run-length encoding; everything
// else is one pass over the K distinct pairs.  HBM bound: 8 B per voxel read, then the 64-bit
// radix sort of the pairs.
#include <cub/cub.cuh>

#include "common.cuh"

namespace isg {

__global__ void __launch_bounds__(256)
pair_kernel(const uint32_t *__restrict__ gt, const uint32_t *__restrict__ seg, uint64_t n,
            unsigned long long *__restrict__ keys, uint32_t *__restrict__ maxes) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    uint32_t mg = 0, ms = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint32_t a = gt[i], b = seg[i];
        keys[i] = ((unsigned long long)a << 32) | b;
        mg = max(mg, a);
        ms = max(ms, b);
    }
    mg = __reduce_max_sync(0xFFFFFFFFu, mg);
    ms = __reduce_max_sync(0xFFFFFFFFu, ms);
    if ((threadIdx.x & 31) == 0) {
        atomicMax(maxes + 0, mg);
        atomicMax(maxes + 1, ms);
    }
}

__global__ void __launch_bounds__(256)
area_kernel(const unsigned long long *__restrict__ pairs, const uint32_t *__restrict__ counts,
            const int *__restrict__ n_pairs, unsigned long long *__restrict__ area_gt,
            unsigned long long *__restrict__ area_sg) {
    const int k = *n_pairs;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < k; i += gridDim.x * blockDim.x) {
        const unsigned long long p = pairs[i];
        atomicAdd(area_gt + (uint32_t)(p >> 32), (unsigned long long)counts[i]);
        atomicAdd(area_sg + (uint32_t)p, (unsigned long long)counts[i]);
    }
}

// out: [0] H(seg|gt) [1] H(gt|seg) [2] tp [3] n_seg_objects [4] n_gt_objects (doubles)
__global__ void __launch_bounds__(256)
score_kernel(const unsigned long long *__restrict__ pairs, const uint32_t *__restrict__ counts,
             const int *__restrict__ n_pairs, const unsigned long long *__restrict__ area_gt,
             const unsigned long long *__restrict__ area_sg, double n, double iou_thr,
             double *__restrict__ out) {
    typedef cub::BlockReduce<double, 256> Red;
    __shared__ typename Red::TempStorage tmp;
    const int k = *n_pairs;
    double h_sg = 0.0, h_gs = 0.0, tp = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < k; i += gridDim.x * blockDim.x) {
        const unsigned long long p = pairs[i];
        const uint32_t a = (uint32_t)(p >> 32), b = (uint32_t)p;
        const double c = (double)counts[i], ag = (double)area_gt[a], as = (double)area_sg[b];
        const double pxy = c / n;
        h_sg -= pxy * log2(c / ag);          // H(seg | gt): pxy * log2(pxy / px), px = ag / n
        h_gs -= pxy * log2(c / as);
        if (a && b && c / (ag + as - c) > iou_thr) tp += 1.0;
    }
    h_sg = Red(tmp).Sum(h_sg);
    __syncthreads();
    h_gs = Red(tmp).Sum(h_gs);
    __syncthreads();
    tp = Red(tmp).Sum(tp);
    if (threadIdx.x == 0) {
        atomicAdd(out + 0, h_sg);
        atomicAdd(out + 1, h_gs);
        atomicAdd(out + 2, tp);
    }
}

__global__ void __launch_bounds__(256)
object_count_kernel(const unsigned long long *__restrict__ area, uint32_t n_labels, double *__restrict__ out) {
    uint32_t c = 0;
    for (uint32_t i = 1 + blockIdx.x * blockDim.x + threadIdx.x; i <= n_labels; i += gridDim.x * blockDim.x)
        c += area[i] != 0;
    c = __reduce_add_sync(0xFFFFFFFFu, c);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(out, (double)c);
}

struct MetricBuffers {
    unsigned long long *keys_a, *keys_b, *uniq, *area_gt, *area_sg;
    uint32_t *counts, *maxes;
    int *n_pairs;
    unsigned char *cub_tmp;
    size_t cub_bytes;
};

static void metric_carve(MetricBuffers *b, Carver &cv, uint64_t n, uint64_t max_labels) {
    b->keys_a = cv.take<unsigned long long>(n);
    b->keys_b = cv.take<unsigned long long>(n);
    b->uniq = cv.take<unsigned long long>(n);
    b->counts = cv.take<uint32_t>(n);
    b->area_gt = cv.take<unsigned long long>(max_labels + 1);
    b->area_sg = cv.take<unsigned long long>(max_labels + 1);
    b->maxes = cv.take<uint32_t>(8);
    b->n_pairs = cv.take<int>(8);
    size_t s1 = 0, s2 = 0;
    cub::DeviceRadixSort::SortKeys(nullptr, s1, (unsigned long long *)nullptr, (unsigned long long *)nullptr, (int)n);
    cub::DeviceRunLengthEncode::Encode(nullptr, s2, (unsigned long long *)nullptr, (unsigned long long *)nullptr,
                                       (uint32_t *)nullptr, (int *)nullptr, (int)n);
    b->cub_bytes = (s1 > s2 ? s1 : s2) + 256;
    b->cub_tmp = cv.take<unsigned char>(b->cub_bytes);
}

}  // namespace isg

using namespace isg;

extern "C" size_t isg_metrics_workspace_bytes(int64_t n, int64_t max_label) {
    if (n <= 0 || max_label < 0) return 0;
    Carver cv(nullptr, 0);
    MetricBuffers b;
    metric_carve(&b, cv, (uint64_t)n, (uint64_t)max_label);
    return cv.off + 512;
}

extern "C" int isg_label_metrics(const uint32_t *gt, const uint32_t *seg, int64_t n, int64_t max_label,
                                 double iou_threshold, double *out8, void *workspace,
                                 size_t workspace_bytes, void *stream) {
    ISG_REQUIRE(gt && seg && out8 && n > 0, ISG_ERR_ARG, "isg_label_metrics: bad argument");
    ISG_REQUIRE(n < 0x7FFFFFF0ll, ISG_ERR_OVERFLOW, "isg_label_metrics: more than 2^31 voxels");
    ISG_REQUIRE(iou_threshold >= 0.5, ISG_ERR_ARG,
                "isg_label_metrics: IoU thresholds below 0.5 need a one-to-one assignment (not implemented)");
    cudaStream_t st = (cudaStream_t)stream;
    Carver cv(workspace, workspace_bytes);
    MetricBuffers b;
    metric_carve(&b, cv, (uint64_t)n, (uint64_t)max_label);
    ISG_REQUIRE(workspace && cv.ok, ISG_ERR_WORKSPACE, "isg_label_metrics: workspace too small (%zu < %zu)",
                workspace_bytes, cv.off);
    const int grid = num_sms() * 8;
    ISG_CUDA(cudaMemsetAsync(b.maxes, 0, 8 * sizeof(uint32_t), st));
    ISG_CUDA(cudaMemsetAsync(out8, 0, 8 * sizeof(double), st));
    ISG_CUDA(cudaMemsetAsync(b.area_gt, 0, ((size_t)max_label + 1) * sizeof(unsigned long long), st));
    ISG_CUDA(cudaMemsetAsync(b.area_sg, 0, ((size_t)max_label + 1) * sizeof(unsigned long long), st));
    pair_kernel<<<grid, 256, 0, st>>>(gt, seg, (uint64_t)n, b.keys_a, b.maxes);
    ISG_LAUNCHED();
    uint32_t mx[2];
    ISG_CUDA(cudaMemcpyAsync(mx, b.maxes, sizeof(mx), cudaMemcpyDeviceToHost, st));
    ISG_CUDA(cudaStreamSynchronize(st));
    ISG_REQUIRE((int64_t)mx[0] <= max_label && (int64_t)mx[1] <= max_label, ISG_ERR_ARG,
                "isg_label_metrics: label %u exceeds max_label=%lld", mx[0] > mx[1] ? mx[0] : mx[1],
                (long long)max_label);
    size_t cb = b.cub_bytes;
    ISG_CUDA(cub::DeviceRadixSort::SortKeys(b.cub_tmp, cb, b.keys_a, b.keys_b, (int)n, 0, 64, st));
    count_launch(4);
    cb = b.cub_bytes;
    ISG_CUDA(cub::DeviceRunLengthEncode::Encode(b.cub_tmp, cb, b.keys_b, b.uniq, b.counts, b.n_pairs, (int)n, st));
    count_launch(2);
    area_kernel<<<grid, 256, 0, st>>>(b.uniq, b.counts, b.n_pairs, b.area_gt, b.area_sg);
    ISG_LAUNCHED();
    score_kernel<<<grid, 256, 0, st>>>(b.uniq, b.counts, b.n_pairs, b.area_gt, b.area_sg, (double)n,
                                       iou_threshold, out8);
    ISG_LAUNCHED();
    object_count_kernel<<<64, 256, 0, st>>>(b.area_sg, mx[1], out8 + 3);
    ISG_LAUNCHED();
    object_count_kernel<<<64, 256, 0, st>>>(b.area_gt, mx[0], out8 + 4);
    ISG_LAUNCHED();
    return ISG_OK;
}
