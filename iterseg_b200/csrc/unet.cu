// placeholder (replaced by the tcgen05 U-Net)
#include "common.cuh"
extern "C" size_t isg_unet_packed_weight_bytes(void) { return 0; }
extern "C" int isg_unet_weights_pack(const void *const *, int, void *, void *) { isg::set_error("unet: not built yet"); return ISG_ERR_ARG; }
extern "C" size_t isg_unet_workspace_bytes(int, int, int, int) { return 0; }
extern "C" isg_unet_plan *isg_unet_plan_create(const void *, int, int, int, int, int64_t, int64_t, int64_t, const int32_t *, const int32_t *, const int32_t *, void *, size_t) { isg::set_error("unet: not built yet"); return nullptr; }
extern "C" void isg_unet_plan_destroy(isg_unet_plan *) {}
extern "C" int isg_unet_forward_chunks(isg_unet_plan *, const float *, float *, void *) { return ISG_ERR_ARG; }
extern "C" int isg_unet_debug_activation(isg_unet_plan *, const char *, int, float *, int64_t, void *) { return ISG_ERR_ARG; }
extern "C" double isg_unet_plan_flops(const isg_unet_plan *) { return 0.0; }
