/* iterseg_b200 -- C-ABI of the B200-native affinity U-Net watershed path.
 *
 * The reference (AbigailMcGovern/iterseg) is pure Python; its "FFI" for this
 * path is the Python plug-in protocol of src/iterseg/segmentation.py,
 * predict.py and watershed.py.  Every entry point below replaces the body of
 * one of those functions; the Python host (iterseg_b200/*.py) keeps the
 * reference names/signatures and calls these through ctypes.
 *
 * Conventions
 *   - plain pointers and sizes; no torch / C++ types.
 *   - all data pointers are DEVICE pointers unless the name ends in `_host`.
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream).
 *   - no allocation inside compute calls: the caller provides a workspace whose
 *     size comes from the matching *_workspace_bytes() query.
 *   - return 0 on success, non-zero on failure; isg_last_error() gives the text.
 *   - volumes are C-contiguous zyx; "padded" means one zero voxel on every face,
 *     i.e. shape (Z+2, Y+2, X+2), exactly as segmentation.py:890-895 allocates.
 */
#ifndef ITERSEG_B200_H
#define ITERSEG_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ISG_OK 0
#define ISG_ERR_CUDA 1
#define ISG_ERR_ARG 2
#define ISG_ERR_WORKSPACE 3
#define ISG_ERR_OVERFLOW 4
#define ISG_ERR_DEVICE 5

/* ---- library ------------------------------------------------------------ */
int isg_version(void);
const char *isg_last_error(void);
/* 0 iff a CUDA device of compute capability 10.x is current. */
int isg_device_check(void);
/* how many kernels this library has launched since load (bench `gpu_launches`) */
uint64_t isg_launch_count(void);
/* Input staging of segment_single_volume (segmentation.py:887-889) on the device:
 * isg_frame_minmax: minmax_out[2] = {min, max} of the n floats (scratch: >= 8 device bytes);
 * isg_frame_divide_by_max: frame[i] /= minmax[1] in place (IEEE float32 division). */
int isg_frame_minmax(const float *frame, int64_t n, float *minmax_out, void *scratch,
                     size_t scratch_bytes, void *stream);
int isg_frame_divide_by_max(float *frame, int64_t n, const float *minmax, void *stream);
/* Frame pipelining (iterseg_b200/pipeline.py): reserve n_sms SMs for the post stage of one frame
 * while the U-Net of the next frame runs on another stream (0 = off, the default). */
int isg_set_post_sm_reservation(int n_sms);

/* ---- affinity flood ------------------------------------------------------
 * Replaces raveled_affinity_watershed (watershed.py:95-159) together with the
 * data preparation of _prep_data (watershed.py:38-63).
 *
 *   aff        3 affinity planes (z,y,x), float32, plane p at aff + p*aff_plane_stride,
 *              each of shape (Za,Ya,Xa) = padded shape minus 2*aff_origin per axis:
 *              aff_origin = 0 -> planes are padded like `labels` (the reference layout,
 *              watershed.py:196-201); aff_origin = 1 -> planes are the unpadded U-Net
 *              channels and the zero border is implicit.
 *   aff_div    3 floats on the device: every key is IEEE fp32 aff / aff_div[axis]
 *              (the per-channel max normalisation of watershed.py:195); pass ones
 *              for already normalised input.
 *   mask       (Zp,Yp,Xp) uint8, non-zero = flood may spread here; faces must be 0.
 *   seeds      n_seeds flat indices into the padded volume, in label order
 *              (label i+1 for seeds[i], watershed.py:61-62).
 *   aff_scale_host  3 host floats or NULL: keys are additionally multiplied by |scale|
 *              (the `scale` argument of affinity_watershed, watershed.py:23-24).
 *   labels     (Zp,Yp,Xp) uint32, in/out.  Written: seeds[i] -> i+1, then the flood.
 *              Voxels that are already non-zero on entry are never claimed.
 * Result: bit-identical to the reference kernel for finite keys.
 */
size_t isg_flood_workspace_bytes(int64_t zp, int64_t yp, int64_t xp, int64_t max_seeds);
int isg_affinity_flood(const float *aff, int64_t aff_plane_stride, int aff_origin,
                       const float *aff_div, const uint8_t *mask,
                       const int64_t *seeds, int64_t n_seeds, uint32_t *labels,
                       int64_t zp, int64_t yp, int64_t xp,
                       const float *aff_scale_host,
                       void *workspace, size_t workspace_bytes, void *stream);

/* ---- feature map -> labels ----------------------------------------------
 * Replaces segment_output_image (watershed.py:165-223) and its helpers
 * _get_centroids (:232-236), _get_mask (:226-229), _remove_unwanted_objects
 * (:239-251), affinity_watershed (:17-35).
 *
 *   feats            (n_chan, Z, Y, X) float32, channel c at feats + c*Z*Y*X
 *   aff_ch[3], mask_ch, cent_ch   channel indices (segmentation.py:192-193)
 *   gauss1_host / gauss2_host     float64 half-kernels w[0..r] of the sigma=1 and
 *                                 sigma=2 Gaussians (host memory; numpy computes them
 *                                 exactly as scipy does), radii r1, r2
 *   peak_thresh      threshold_abs of peak_local_max (0.04, watershed.py:235)
 *   absolute_thresh  if use_absolute_thresh != 0 the mask is feats[mask_ch] > absolute_thresh
 *                    (watershed.py:209-212), otherwise Otsu of the sigma=2 smoothed channel
 *   min_area/max_area  keep components with min_area <= size < max_area (watershed.py:215)
 *   labels           (Z+2, Y+2, X+2) uint32, must be zero on entry; written in place
 *   mask_out         (Z+2, Y+2, X+2) uint8: the kept mask (third return value)
 *   seeds_out        capacity max_seeds flat PADDED indices of the kept seeds, label order
 *   counts_out       device int64[8]: {n_seeds_kept, n_candidates, n_components,
 *                    n_multi_seed_components, halo_violation, 0, 0, 0}
 *   otsu_out         device float[1]: the threshold used
 *
 * Slab mode (one z-slab of a larger volume, extended by halo planes):
 *   use_aff_div / aff_div   per-channel affinity maxima of the WHOLE volume (an all-reduce of
 *                           isg_slab_stats stage 0) instead of the local maxima (watershed.py:195)
 *   own_z0, own_z1          the planes [own_z0, own_z1) of this volume that the caller will keep
 *                           (own_z1 == 0: everything)
 *   open_faces              bit 0: the volume continues below plane 0, bit 1: above plane z-1.
 *                           A mask component that touches an open face AND the own planes cannot
 *                           be segmented exactly from this slab: counts_out[4] is set to 1
 *   seed_keys_out           optional device uint64[max_seeds]: for every kept seed (label order)
 *                           (~order_preserving_bits(smoothed centre value) << 32) | flat UNPADDED
 *                           voxel index -- the key the seeds are sorted by, for a global relabel
 */
typedef struct {
    int aff_ch[3];
    int mask_ch;
    int cent_ch;
    int r1;
    int r2;
    float peak_thresh;
    int use_absolute_thresh;
    float absolute_thresh;
    int64_t min_area;
    int64_t max_area;
    float scale[3]; /* |scale| multiplies the affinities (watershed.py:23-24); 1,1,1 = None */
    int use_aff_div;
    float aff_div[3];
    int own_z0;
    int own_z1;
    int open_faces;
    unsigned long long *seed_keys_out;
} isg_post_params;

size_t isg_post_workspace_bytes(int64_t z, int64_t y, int64_t x, int64_t max_seeds);
/* isg_post_workspace_bytes covers the worst case -- every voxel inside a multi-seed mask
 * component (about 200 bytes per voxel).  isg_segment_features also accepts a SMALLER workspace:
 * whatever lies beyond the fixed part becomes the compact node / edge arenas of the ordered flood,
 * and a frame whose multi-seed components do not fit fails with ISG_ERR_WORKSPACE (nothing is
 * silently dropped).  This query sizes a workspace for at most `max_flood_nodes` such voxels --
 * what the z-slab path uses, where a slab of 2048 x 2048 planes would otherwise need > 60 GB. */
size_t isg_post_workspace_bytes_capped(int64_t z, int64_t y, int64_t x, int64_t max_seeds,
                                       int64_t max_flood_nodes);
int isg_segment_features(const float *feats, int n_chan, int64_t z, int64_t y, int64_t x,
                         const isg_post_params *params,
                         const double *gauss1_host, const double *gauss2_host,
                         uint32_t *labels, uint8_t *mask_out, int64_t *seeds_out,
                         int64_t max_seeds, int64_t *counts_out, float *otsu_out,
                         void *workspace, size_t workspace_bytes, void *stream);

/* ---- spatial (z-slab) sharding of one large volume: global statistics -----
 * The scalars of segment_output_image that couple the whole volume, computed slab by slab so
 * that they can be combined with an all-reduce (MAX / MIN / SUM) and fed back through
 * isg_post_params (aff_div, absolute_thresh):
 *   stage 0: chan_max_out[3] = maxima of the affinity channels over the own planes
 *            (watershed.py:195); minmax_io[2] = min / max of the sigma=2 smoothed mask channel
 *            over the own planes (the smoothing sees the halo planes of the slab)
 *   stage 1: hist_out[256] = numpy.histogram(smoothed own planes, 256, range = minmax_io)
 *            with minmax_io holding the GLOBAL min / max
 * isg_otsu_from_hist: threshold_otsu on the (summed) histogram, numpy float32 arithmetic.
 * All outputs are device pointers; feats is the (C, z, y, x) slab INCLUDING its halo planes. */
int isg_slab_stats(const float *feats, int n_chan, int64_t z, int64_t y, int64_t x,
                   const isg_post_params *params, const double *gauss2_host, int stage,
                   float *minmax_io, float *chan_max_out, unsigned long long *hist_out,
                   void *workspace, size_t workspace_bytes, void *stream);
int isg_otsu_from_hist(const unsigned long long *hist, const float *minmax, float *thr_out,
                       void *scratch /* >= 2048 bytes, device */, size_t scratch_bytes, void *stream);
/* keys[0..n) ascending, in place (tmp: n uint64 of scratch + isg_sort_tmp_bytes(n) bytes) */
size_t isg_sort_tmp_bytes(int64_t n);
int isg_sort_keys_u64(unsigned long long *keys, int64_t n, void *tmp, size_t tmp_bytes, void *stream);
/* labels[i] (non-zero, <= n_local) -> 1 + position of local_keys[labels[i]-1] in the sorted
 * global key list (a label whose key is missing becomes 0 and *missing_out is set to 1) */
int isg_relabel_by_keys(uint32_t *labels, int64_t n, const unsigned long long *local_keys,
                        int64_t n_local, const unsigned long long *global_sorted_keys,
                        int64_t n_global, uint32_t *lut_scratch, int *missing_out, void *stream);

/* ---- DoG blob segmenter ------------------------------------------------------
 * Replaces dog_blob_watershed_for_chunks + dog_image (segmentation.py:592-650, :678-680).
 *   vol          (Z,Y,X) float32 frame (the function pads it by one voxel itself, :634)
 *   params       Gaussian half kernels (float64, scipy's) and radii of, in order:
 *                  [0] min_sigma 'nearest'  [1] max_sigma 'nearest'   (the mask, dog_image)
 *                  [2] sigma_list[0] 'reflect'  [3] sigma_list[1] 'reflect'   (blob_dog, one DoG layer)
 *                threshold (mask and blob threshold), scale_factor = 1/(sigma_ratio-1),
 *                prune_d2 / prune_radius: largest squared lattice distance at which two blobs of
 *                sigma_list[0] overlap by more than 0.5 (_prune_blobs) and its integer radius
 *   labels       (Z+2,Y+2,X+2) uint32, zero on entry: markers + watershed result, in place
 *   mask_out     (Z+2,Y+2,X+2) uint8: dog > threshold
 *   distance_out optional (Z+2,Y+2,X+2) float64: ndi.distance_transform_edt(padded volume)
 *   counts_out   device int64[4]: {peak candidates, blobs after pruning, marker voxels, 0}
 */
#define ISG_DOG_MAX_SIGMAS 9     /* sigma_list of blob_dog: up to 8 DoG layers */
#define ISG_GAUSS_MAX_RADIUS 27  /* sigma <= 6.8 at truncate = 4 */
typedef struct {
    double weights[4][12];
    int radius[4];
    float threshold;
    float scale_factor;
    int prune_d2;
    int prune_radius;
    /* multi-layer blob_dog (max_sigma / min_sigma >= sigma_ratio): n_layers = k >= 2 DoG layers from the
     * k + 1 'reflect' Gaussians of sigma_list[i] = min_sigma * sigma_ratio^i; 3^4 maxima over (z,y,x,layer);
     * _prune_blobs with per-blob sigma: pairs (i < j, peak order) closer than 2 * max(sigma) * sqrt(3), walked
     * in lexicographic order, the blob that is not larger dies when the spheres overlap by more than
     * `overlap`.  n_layers <= 1 selects the single-layer fields above (the default configuration). */
    int n_layers;
    double overlap;
    double layer_sigma[ISG_DOG_MAX_SIGMAS];
    int layer_radius[ISG_DOG_MAX_SIGMAS];
    double layer_weights[ISG_DOG_MAX_SIGMAS][ISG_GAUSS_MAX_RADIUS + 1];
    /* the two 'nearest' Gaussians of the mask (min_sigma, max_sigma) when n_layers > 1: max_sigma may
     * need a radius beyond the 11 of weights[0..1] */
    int mask_radius[2];
    double mask_weights[2][ISG_GAUSS_MAX_RADIUS + 1];
} isg_dog_params;
size_t isg_dog_workspace_bytes(int64_t z, int64_t y, int64_t x, int64_t max_seeds);
/* workspace for n_layers DoG layers (n_layers <= 1: the same as isg_dog_workspace_bytes) */
size_t isg_dog_workspace_bytes_layers(int64_t z, int64_t y, int64_t x, int64_t max_seeds, int n_layers);
int isg_dog_blob_segment(const float *vol, int64_t z, int64_t y, int64_t x, const isg_dog_params *params,
                         uint32_t *labels, uint8_t *mask_out, double *distance_out, int64_t max_seeds,
                         int64_t *counts_out, void *workspace, size_t workspace_bytes, void *stream);

/* ---- assessment metrics (src/iterseg/metrics.py:107, :205-227) -------------
 * gt / seg: n uint32 labels each (device).  out8 (device doubles):
 *   [0] H(seg|gt)  [1] H(gt|seg)   (variation of information, log base 2, background included)
 *   [2] TP = label pairs (both non-zero) with IoU > iou_threshold (>= 0.5: one-to-one by itself)
 *   [3] number of seg objects   [4] number of gt objects   (FP = [3]-[2], FN = [4]-[2])
 * max_label: upper bound of both label sets (sizes the area tables). */
size_t isg_metrics_workspace_bytes(int64_t n, int64_t max_label);
int isg_label_metrics(const uint32_t *gt, const uint32_t *seg, int64_t n, int64_t max_label,
                      double iou_threshold, double *out8, void *workspace, size_t workspace_bytes,
                      void *stream);

/* ---- 3-D U-Net over chunks ----------------------------------------------
 * Replaces process_chunks + predict_chunk_feature_map + UNet.forward
 * (predict.py:64-126, unet.py:284-364) for the bundled architecture
 * UNet(in_channels=1, out_channels=5) with train-mode BatchNorm (per-chunk
 * statistics), see include comments in iterseg_b200/csrc/unet.cuh.
 *
 * Weights: the caller passes the state_dict tensors (device, float32, in the
 * key order of SURVEY.md Appendix A, running stats and biases of convs may be
 * NULL: they do not influence train-mode output) through isg_unet_weights_pack,
 * which writes the packed fp16 tap-major layout into `packed`.
 */
typedef struct isg_unet_plan isg_unet_plan;

size_t isg_unet_packed_weight_bytes(void);
/* tensors: array of 134 device pointers in state_dict order (NULL allowed for
 * entries that are unused: conv biases, running_mean/var, num_batches_tracked) */
int isg_unet_weights_pack(const void *const *tensors, int n_tensors, void *packed, void *stream);

size_t isg_unet_workspace_bytes(int n_chunks, int cz, int cy, int cx);
/* chunk tables are HOST arrays of n_chunks*3 int32: start (z,y,x), crop_lo, crop_hi
 * exactly as make_chunks returns them (predict.py:38-61).  They may be NULL at creation and
 * (re)set at any time with isg_unet_plan_set_chunks: a plan is a (frame extent, chunk extent,
 * chunk count) geometry over a workspace, the tables are per-call data.  The tables are copied
 * into pinned memory owned by the plan and uploaded by EVERY isg_unet_forward_chunks with
 * cudaMemcpyAsync on that call's stream -- no blocking copy, nothing on the legacy stream.
 * Several plans may share one workspace as long as their forward passes are enqueued on the
 * same stream (or otherwise ordered): a forward pass leaves nothing in the workspace that a
 * later one needs. */
int isg_unet_plan_set_chunks(isg_unet_plan *plan, const int32_t *starts_host,
                             const int32_t *crop_lo_host, const int32_t *crop_hi_host);
isg_unet_plan *isg_unet_plan_create(const void *packed_weights, int n_chunks,
                                    int cz, int cy, int cx,
                                    int64_t z, int64_t y, int64_t x,
                                    const int32_t *starts_host, const int32_t *crop_lo_host,
                                    const int32_t *crop_hi_host,
                                    void *workspace, size_t workspace_bytes);
void isg_unet_plan_destroy(isg_unet_plan *plan);
/* frame (Z,Y,X) float32 -> feats (5,Z,Y,X) float32: every voxel written by exactly
 * one chunk's cropped interior (predict.py:89-95). */
int isg_unet_forward_chunks(isg_unet_plan *plan, const float *frame, float *feats, void *stream);
/* The same with the input normalisation of segment_single_volume (vol /= max, segmentation.py:889)
 * fused into the first kernel's loads: every voxel is divided by norm_max[0] (device float, e.g.
 * minmax_out + 1 of isg_frame_minmax) -- bit-identical to isg_frame_divide_by_max + forward.
 * `frame` may then be PINNED HOST memory mapped into the device address space (cudaHostAlloc /
 * cudaHostRegister with unified addressing): the chunks are staged straight from the pinned
 * volume, no H2D copy of the frame ("zero-copy").  Measured on B200 against the copy-engine path
 * in profiles/r02_notes.md. */
int isg_unet_forward_chunks_norm(isg_unet_plan *plan, const float *frame, const float *norm_max,
                                 float *feats, void *stream);
/* host-only query of the band-flat conv tiling for tests: out4 = {P, R, tiles per
 * plane, plane-slot bytes} for an (H, W) plane, rows of row_bytes, slots of at most max_slot_bytes */
int isg_debug_flat_tiling(int H, int W, int row_bytes, int64_t max_slot_bytes, int64_t *out4);

/* fp16 range guard.  Filters are divided by a per-output-channel power of two at pack time (the
 * train-mode BatchNorm that follows every convolution cancels it; BN_EPS is rescaled to match), so
 * the fp16 weights and pre-BatchNorm activations stay O(1) for any filter scale.  Should an
 * activation still leave the fp16 range (|x| >= ~5.6e3 counts), the kernels raise a sticky device
 * flag that every forward pass copies to pinned host memory:
 *   isg_unet_plan_overflowed   != 0 iff a forward pass of this plan that has COMPLETED overflowed
 *                              (no synchronisation: call it after waiting for the features);
 *                              isg_unet_forward_chunks then refuses with ISG_ERR_OVERFLOW
 *   isg_unet_plan_clear_overflow   re-arms the flag (stream-ordered). */
int isg_unet_plan_overflowed(const isg_unet_plan *plan);
int isg_unet_plan_clear_overflow(isg_unet_plan *plan, void *stream);
/* debugging / parity: run the network on `frame` up to and including the convolution
 * `name` ("c0.conv0" ... "c8_0.conv1") and return its raw (pre-BatchNorm, bias-free)
 * output for chunk `chunk` as fp32 NCDHW (out_elems = Cout*D*H*W of that level). */
int isg_unet_debug_activation(isg_unet_plan *plan, const float *frame, const char *name, int chunk,
                              float *out, int64_t out_elems, void *stream);
/* algorithmic FLOPs of one forward over all chunks of the plan (2*MAC of 18 conv + 4 tconv) */
double isg_unet_plan_flops(const isg_unet_plan *plan);

/* per-launch CUDA-event timing of the forward pass (used by bench.py for the roofline
 * object): enable, run forwards, synchronise the stream, read.
 * out[5] = {ms in tcgen05 convolutions, #tcgen05 launches, ms of whole forwards, #forwards,
 *           algorithmic FLOPs of the tcgen05 convolutions of ONE forward}. */
int isg_unet_plan_profile(isg_unet_plan *plan, int enable);   /* 0 off, 1 tcgen05 convs + whole forward, 2 every launch */
int isg_unet_plan_profile_read(isg_unet_plan *plan, double *out);
/* per-launch times (ms) of the recorded forward passes in launch order, 35 launches per forward;
 * kind_out (nullable): 0 = TMA-fed tcgen05 convolution, 2 = any other kernel.  Returns the count. */
int isg_unet_plan_profile_launches(isg_unet_plan *plan, double *ms_out, int *kind_out, int cap);
/* (start, end) of every recorded whole forward pass in ms since the first start; returns the count */
int isg_unet_plan_profile_timeline(isg_unet_plan *plan, double *out, int cap);

/* ---- label bookkeeping for frame-sharded time series ------------------------
 * labels[i] += offset for every non-zero label (global label ids across frames:
 * an addition of this implementation, the reference restarts at 1 in every frame,
 * watershed.py:61-62). */
int isg_add_label_offset(uint32_t *labels, int64_t n, uint32_t offset, void *stream);
/* The crop the driver applies to the padded output (segmentation.py:896,900) on the device:
 * labels_padded (z+2,y+2,x+2) -> out (z,y,x) contiguous, ready for ONE contiguous device->host copy.
 * offset_dev: NULL, or a device int64 added to every non-zero label (the running global label
 * offset of a frame-sharded series stays on the device: no host round trip per frame). */
int isg_crop_labels(const uint32_t *labels_padded, int64_t z, int64_t y, int64_t x, uint32_t *out,
                    const int64_t *offset_dev, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* ITERSEG_B200_H */
