// Library-level entry points: version, error text, launch counter.
#include <stdarg.h>

#include <atomic>
#include <mutex>

#include "common.cuh"

namespace isg {
static thread_local char g_err[1024] = "";
static std::atomic<uint64_t> g_launches{0};

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }
static std::atomic<int> g_post_sms{0};
int post_sms() { return g_post_sms.load(std::memory_order_relaxed); }
}  // namespace isg

extern "C" int isg_version(void) { return 100; }
extern "C" const char *isg_last_error(void) { return isg::g_err; }
extern "C" uint64_t isg_launch_count(void) { return isg::g_launches.load(); }
extern "C" int isg_set_post_sm_reservation(int n_sms) {
    ISG_REQUIRE(n_sms >= 0 && n_sms <= isg::num_sms() / 2, ISG_ERR_ARG, "isg_set_post_sm_reservation: %d out of range", n_sms);
    isg::g_post_sms.store(n_sms);
    return ISG_OK;
}
extern "C" int isg_device_check(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
        isg::set_error("no CUDA device visible");
        return ISG_ERR_DEVICE;
    }
    int dev = 0, major = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    if (major != 10) {
        isg::set_error("device %d has compute capability %d.x; this library is sm_100a only", dev, major);
        return ISG_ERR_DEVICE;
    }
    return ISG_OK;
}

namespace isg {
__global__ void __launch_bounds__(256)
add_label_offset_kernel(uint32_t *labels, uint64_t n, uint32_t offset) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint32_t l = labels[i];
        if (l) labels[i] = l + offset;
    }
}
}  // namespace isg

namespace isg {
// input staging (segmentation.py:887-889): min / max of the frame, then frame /= max
__global__ void __launch_bounds__(256)
frame_minmax_kernel(const float *__restrict__ v, uint64_t n, uint32_t *__restrict__ mm) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    uint32_t lo = 0xFFFFFFFFu, hi = 0u;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint32_t k = f32_ord(v[i]);
        lo = min(lo, k);
        hi = max(hi, k);
    }
    lo = __reduce_min_sync(0xFFFFFFFFu, lo);
    hi = __reduce_max_sync(0xFFFFFFFFu, hi);
    if ((threadIdx.x & 31) == 0) {
        atomicMin(mm + 0, lo);
        atomicMax(mm + 1, hi);
    }
}
__global__ void minmax_init_kernel(uint32_t *mm) {
    mm[0] = 0xFFFFFFFFu;
    mm[1] = 0u;
}
// padded (Z+2,Y+2,X+2) labels -> the (Z,Y,X) interior, contiguous (segmentation.py:896,900), with an
// optional device-resident offset added to every non-zero label (frame-sharded series)
__global__ void __launch_bounds__(256)
crop_labels_kernel(const uint32_t *__restrict__ lab, int Z, int Y, int X, uint32_t *__restrict__ out,
                   const long long *__restrict__ offset) {
    const uint32_t off = offset ? (uint32_t)*offset : 0u;
    const uint64_t n = (uint64_t)Z * Y * X, stride = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t yp = (uint64_t)Y + 2, xp = (uint64_t)X + 2;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint64_t x = i % X, t = i / X, y = t % Y, z = t / Y;
        const uint32_t l = lab[((z + 1) * yp + (y + 1)) * xp + (x + 1)];
        out[i] = l ? l + off : 0u;
    }
}
__global__ void minmax_out_kernel(const uint32_t *__restrict__ mm, float *__restrict__ out) {
    out[0] = ord_f32(mm[0]);
    out[1] = ord_f32(mm[1]);
}
__global__ void __launch_bounds__(256)
frame_divide_kernel(float *__restrict__ v, uint64_t n, const float *__restrict__ minmax) {
    const float d = minmax[1];
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        v[i] = __fdiv_rn(v[i], d);               // IEEE float32 division, as numpy's in-place /=
}
}  // namespace isg

extern "C" int isg_frame_minmax(const float *frame, int64_t n, float *minmax_out, void *scratch,
                                size_t scratch_bytes, void *stream) {
    ISG_REQUIRE(frame && minmax_out && scratch && n > 0, ISG_ERR_ARG, "isg_frame_minmax: bad argument");
    ISG_REQUIRE(scratch_bytes >= 8, ISG_ERR_WORKSPACE, "isg_frame_minmax: scratch must hold 8 bytes");
    cudaStream_t st = (cudaStream_t)stream;
    uint32_t *mm = reinterpret_cast<uint32_t *>(scratch);
    isg::minmax_init_kernel<<<1, 1, 0, st>>>(mm);       // (a copy from pageable host memory would stage synchronously)
    ISG_LAUNCHED();
    isg::frame_minmax_kernel<<<isg::num_sms() * 8, 256, 0, st>>>(frame, (uint64_t)n, mm);
    ISG_LAUNCHED();
    isg::minmax_out_kernel<<<1, 1, 0, st>>>(mm, minmax_out);
    ISG_LAUNCHED();
    return ISG_OK;
}

extern "C" int isg_frame_divide_by_max(float *frame, int64_t n, const float *minmax, void *stream) {
    ISG_REQUIRE(frame && minmax && n > 0, ISG_ERR_ARG, "isg_frame_divide_by_max: bad argument");
    isg::frame_divide_kernel<<<isg::num_sms() * 8, 256, 0, (cudaStream_t)stream>>>(frame, (uint64_t)n, minmax);
    ISG_LAUNCHED();
    return ISG_OK;
}

extern "C" int isg_add_label_offset(uint32_t *labels, int64_t n, uint32_t offset, void *stream) {
    ISG_REQUIRE(labels && n >= 0, ISG_ERR_ARG, "isg_add_label_offset: bad argument");
    if (n == 0 || offset == 0) return ISG_OK;
    isg::add_label_offset_kernel<<<isg::num_sms() * 8, 256, 0, (cudaStream_t)stream>>>(labels, (uint64_t)n, offset);
    ISG_LAUNCHED();
    return ISG_OK;
}

extern "C" int isg_crop_labels(const uint32_t *labels_padded, int64_t z, int64_t y, int64_t x, uint32_t *out,
                               const int64_t *offset_dev, void *stream) {
    ISG_REQUIRE(labels_padded && out && z > 0 && y > 0 && x > 0, ISG_ERR_ARG, "isg_crop_labels: bad argument");
    ISG_REQUIRE(z < (1ll << 30) && y < (1ll << 30) && x < (1ll << 30), ISG_ERR_ARG, "isg_crop_labels: extent too large");
    isg::crop_labels_kernel<<<isg::num_sms() * 8, 256, 0, (cudaStream_t)stream>>>(
        labels_padded, (int)z, (int)y, (int)x, out, reinterpret_cast<const long long *>(offset_dev));
    ISG_LAUNCHED();
    return ISG_OK;
}
