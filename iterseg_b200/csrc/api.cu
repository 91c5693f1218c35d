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

extern "C" int isg_add_label_offset(uint32_t *labels, int64_t n, uint32_t offset, void *stream) {
    ISG_REQUIRE(labels && n >= 0, ISG_ERR_ARG, "isg_add_label_offset: bad argument");
    if (n == 0 || offset == 0) return ISG_OK;
    isg::add_label_offset_kernel<<<isg::num_sms() * 8, 256, 0, (cudaStream_t)stream>>>(labels, (uint64_t)n, offset);
    ISG_LAUNCHED();
    return ISG_OK;
}
