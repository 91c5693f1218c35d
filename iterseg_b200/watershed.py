"""Feature map -> labels: the reference's watershed module on the B200.

Mirrors the public functions of src/iterseg/watershed.py (same names, argument
meaning and return values) on top of the C-ABI (include/iterseg_b200.h):

  segment_output_image        watershed.py:165-223  -> isg_segment_features
  affinity_watershed          watershed.py:17-35    -> isg_affinity_flood
  raveled_affinity_watershed  watershed.py:95-159   -> isg_affinity_flood

Arrays may be numpy (host; copied to the device and back, results written in
place through `out` exactly like the reference) or torch CUDA tensors (no
copies).  There is no CPU implementation here.
"""
import ctypes

import numpy as np
import torch

from . import _lib

__all__ = ['segment_output_image', 'affinity_watershed', 'raveled_affinity_watershed',
           'segment_features_device', 'gaussian_half_kernel']

_ws_cache = {}


def _workspace(kind, key, nbytes, device):
    k = (kind, key, str(device))
    t = _ws_cache.get(k)
    if t is None or t.numel() < nbytes:
        _ws_cache.clear() if len(_ws_cache) > 8 else None
        t = torch.empty(int(nbytes), dtype=torch.uint8, device=device)
        _ws_cache[k] = t
    return t


def gaussian_half_kernel(sigma, truncate=4.0):
    """w[0..r] of scipy.ndimage's 1-D Gaussian (float64), r = int(truncate*sigma + 0.5);
    sigma == 0 means "skip this filter" (r = 0)."""
    if sigma <= 0:
        return np.ones(1, dtype=np.float64), 0
    r = int(truncate * float(sigma) + 0.5)
    x = np.arange(-r, r + 1)
    phi = np.exp(-0.5 / (float(sigma) * float(sigma)) * x ** 2)
    phi = phi / phi.sum()
    return np.ascontiguousarray(phi[r:], dtype=np.float64), r


def _as_device(a, dtype, device):
    if isinstance(a, torch.Tensor):
        return a.to(device=device, dtype=dtype).contiguous()
    return torch.from_numpy(np.ascontiguousarray(a)).to(device=device, dtype=dtype)


def post_params(affinities_channels=(0, 1, 2), centroids_channel=4, thresholding_channel=3,
                scale=None, absolute_thresh=None, min_area=10, max_area=10000000):
    """(isg_post_params, sigma=1 half kernel, sigma=2 half kernel) with the reference's constants
    (watershed.py:227,234-235,241-246)."""
    p = _lib.PostParams()
    for i, c in enumerate(affinities_channels):
        p.aff_ch[i] = int(c)
    p.mask_ch = int(thresholding_channel)
    p.cent_ch = int(centroids_channel)
    w1, r1 = gaussian_half_kernel(1.0)
    w2, r2 = gaussian_half_kernel(2.0)
    p.r1, p.r2 = r1, r2
    p.peak_thresh = 0.04
    p.use_absolute_thresh = 0 if absolute_thresh is None else 1
    p.absolute_thresh = 0.0 if absolute_thresh is None else float(absolute_thresh)
    p.min_area, p.max_area = int(min_area), int(max_area)
    sc = np.ones(3, np.float32) if scale is None else np.abs(
        np.broadcast_to(np.asarray(scale, np.float32).reshape(-1), (3,)))
    for i in range(3):
        p.scale[i] = float(sc[i])
    return p, w1, w2


def segment_features_device(feats, labels, affinities_channels=(0, 1, 2), centroids_channel=4,
                            thresholding_channel=3, scale=None, absolute_thresh=None,
                            max_seeds=None, min_area=10, max_area=10000000, slab=None,
                            max_flood_nodes=None):
    """Device-resident core of segment_output_image.

    feats  : (C,Z,Y,X) float32 CUDA tensor (not modified)
    labels : (Z+2,Y+2,X+2) uint32-as-int32 CUDA tensor, all zero; written in place
    Returns (seeds_padded_flat int64[max], counts int64[8], mask uint8 padded, otsu float[1]),
    all CUDA tensors; counts = (n_seeds, n_candidates, n_components, n_multi_seed_components,
    halo_violation, 0, 0, 0).
    slab: None, or a dict for one z-slab of a larger volume (iterseg_b200/slab.py):
    aff_div (3 floats), own_z0, own_z1, open_faces, seed_keys (int64 CUDA tensor [max_seeds]).
    max_flood_nodes: None = a workspace for the worst case (every voxel inside a multi-seed mask
    component, ~200 B per voxel); a number sizes the ordered flood's compact arenas for at most
    that many such voxels (isg_post_workspace_bytes_capped) -- a frame that needs more raises
    IsgError with status ISG_ERR_WORKSPACE and must be re-run on zeroed labels.
    """
    lib = _lib.load()
    assert feats.is_cuda and feats.dtype == torch.float32 and feats.is_contiguous()
    assert labels.is_cuda and labels.dtype == torch.int32 and labels.is_contiguous()
    C, Z, Y, X = feats.shape
    assert tuple(labels.shape) == (Z + 2, Y + 2, X + 2)
    dev = feats.device
    if max_seeds is None:
        max_seeds = max(1 << 16, (Z * Y * X) // 8)
    p, w1, w2 = post_params(affinities_channels, centroids_channel, thresholding_channel, scale,
                            absolute_thresh, min_area, max_area)
    if slab is not None:
        p.use_aff_div = 1
        for i in range(3):
            p.aff_div[i] = float(slab['aff_div'][i])
        p.own_z0, p.own_z1, p.open_faces = int(slab['own_z0']), int(slab['own_z1']), int(slab['open_faces'])
        keys = slab['seed_keys']
        assert keys.is_cuda and keys.dtype == torch.int64 and keys.numel() >= max_seeds
        p.seed_keys_out = keys.data_ptr()
    if max_flood_nodes is None:
        nbytes = lib.isg_post_workspace_bytes(Z, Y, X, max_seeds)
    else:
        nbytes = lib.isg_post_workspace_bytes_capped(Z, Y, X, max_seeds, max(int(max_flood_nodes), 1))
    ws = _workspace('post', (Z, Y, X, max_seeds, max_flood_nodes), nbytes, dev)
    mask = torch.empty((Z + 2, Y + 2, X + 2), dtype=torch.uint8, device=dev)
    seeds = torch.empty(max_seeds, dtype=torch.int64, device=dev)
    counts = torch.zeros(8, dtype=torch.int64, device=dev)
    otsu = torch.zeros(1, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        rc = lib.isg_segment_features(
            feats.data_ptr(), C, Z, Y, X, ctypes.byref(p), w1.ctypes.data, w2.ctypes.data,
            labels.data_ptr(), mask.data_ptr(), seeds.data_ptr(), max_seeds, counts.data_ptr(),
            otsu.data_ptr(), ws.data_ptr(), ws.numel(), _lib.stream_ptr())
    _lib.check(rc, 'isg_segment_features')
    return seeds, counts, mask, otsu


def _unravel_padded(flat, shape_p):
    zp, yp, xp = shape_p
    z = flat // (yp * xp)
    r = flat - z * (yp * xp)
    y = r // xp
    x = r - y * xp
    return np.stack([z, y, x], axis=1)


def segment_output_image(unet_output, affinities_channels, centroids_channel,
                         thresholding_channel, scale=None, absolute_thresh=None, out=None,
                         py_func=False):
    """Same contract as the reference (watershed.py:165-223): returns
    (segmentation (Z,Y,X), seeds (N,3), mask (Z+2,Y+2,X+2) bool); when `out` (a
    flat or padded-shaped uint32/int32 array of (Z+2)(Y+2)(X+2) zeros) is given
    the labels are written into it in place.  `py_func` is accepted and ignored."""
    _lib.require_device()
    dev = torch.device('cuda', torch.cuda.current_device())
    on_device = isinstance(unet_output, torch.Tensor) and unet_output.is_cuda
    if isinstance(unet_output, torch.Tensor):
        feats = unet_output.squeeze()
    else:
        feats = torch.from_numpy(np.ascontiguousarray(np.squeeze(np.asarray(unet_output)),
                                                      dtype=np.float32))
    feats = feats.to(device=dev, dtype=torch.float32).contiguous()
    if feats.dim() != 4:
        raise ValueError(f'unet_output must squeeze to (c, z, y, x); got {tuple(feats.shape)}')
    C, Z, Y, X = feats.shape
    shape_p = (Z + 2, Y + 2, X + 2)
    if out is not None and isinstance(out, torch.Tensor) and out.is_cuda:
        labels = out.view(shape_p)
        assert labels.dtype == torch.int32
    else:
        labels = torch.zeros(shape_p, dtype=torch.int32, device=dev)
    seeds_d, counts_d, mask_d, _ = segment_features_device(
        feats, labels, affinities_channels, centroids_channel, thresholding_channel,
        scale=scale, absolute_thresh=absolute_thresh)
    counts = counts_d.cpu().numpy()
    n_seeds = int(counts[0])
    if on_device and (out is None or isinstance(out, torch.Tensor)):
        seeds = seeds_d[:n_seeds]
        zp, yp, xp = shape_p
        z = torch.div(seeds, yp * xp, rounding_mode='floor')
        r = seeds - z * (yp * xp)
        y = torch.div(r, xp, rounding_mode='floor')
        coords = torch.stack([z, y, r - y * xp], dim=1) - 1
        return labels[1:-1, 1:-1, 1:-1], coords, mask_d.bool()
    seeds = _unravel_padded(seeds_d[:n_seeds].cpu().numpy(), shape_p) - 1
    mask = mask_d.cpu().numpy().astype(bool)
    lab_host = labels.cpu().numpy()
    if out is not None:
        out_arr = out.reshape(shape_p)          # a view for the flat array the reference passes
        out_arr[...] = lab_host.view(np.uint32).astype(out_arr.dtype, copy=False)
        seg = out_arr
    else:
        seg = lab_host.view(np.uint32)
    return seg[1:-1, 1:-1, 1:-1], seeds, mask


def _run_flood(aff_d, origin, div_d, mask_d, seeds_d, labels_d, shape_p, scale):
    lib = _lib.load()
    zp, yp, xp = shape_p
    n_seeds = int(seeds_d.numel())
    nbytes = lib.isg_flood_workspace_bytes(zp, yp, xp, max(n_seeds, 1))
    ws = _workspace('flood', (zp, yp, xp, n_seeds), nbytes, aff_d.device)
    sc = None
    if scale is not None:
        sc = np.ascontiguousarray(np.abs(np.broadcast_to(
            np.asarray(scale, np.float32).reshape(-1), (3,))), dtype=np.float32)
    with torch.cuda.device(aff_d.device):
        rc = lib.isg_affinity_flood(
            aff_d.data_ptr(), aff_d.stride(0), origin, div_d.data_ptr(), mask_d.data_ptr(),
            seeds_d.data_ptr() if n_seeds else None, n_seeds, labels_d.data_ptr(), zp, yp, xp,
            sc.ctypes.data if sc is not None else None, ws.data_ptr(), ws.numel(),
            _lib.stream_ptr())
    _lib.check(rc, 'isg_affinity_flood')


def affinity_watershed(image, marker_coords, mask, scale=None, out=None, py_func=False):
    """watershed.py:17-35.  image (3,Z,Y,X) float32 affinities (already padded),
    marker_coords (N,3) voxel coordinates, mask (Z,Y,X) bool with False faces (None: all
    interior voxels), out: optional flat/shaped integer array written in place
    (must be zero).  Returns the label volume of shape image.shape[1:].
    An empty marker list gives an all-zero result (the reference raises inside
    numpy's apply_along_axis, watershed.py:50)."""
    _lib.require_device()
    dev = torch.device('cuda', torch.cuda.current_device())
    shape_p = tuple(int(s) for s in image.shape[1:])
    if len(shape_p) != 3 or image.shape[0] != 3:
        raise ValueError('image must have shape (3, z, y, x)')
    aff_d = _as_device(image, torch.float32, dev)
    if mask is None:
        m = torch.zeros(shape_p, dtype=torch.uint8, device=dev)
        m[1:-1, 1:-1, 1:-1] = 1
        mask_d = m
    else:
        mask_d = _as_device(mask, torch.uint8, dev)
        if tuple(mask_d.shape) != shape_p:
            mask_d = mask_d.reshape(shape_p)
    coords = np.asarray(marker_coords.cpu() if isinstance(marker_coords, torch.Tensor)
                        else marker_coords, dtype=np.int64).reshape(-1, 3)
    if len(coords) and ((coords < 0).any() or (coords >= np.array(shape_p)).any()):
        raise IndexError('marker coordinate outside the volume')
    strides = np.array([shape_p[1] * shape_p[2], shape_p[2], 1], dtype=np.int64)
    seeds_d = torch.from_numpy(coords @ strides).to(dev)
    out_is_dev = isinstance(out, torch.Tensor) and out.is_cuda
    if out_is_dev:
        labels_d = out.view(shape_p)
    elif out is not None:
        labels_d = torch.from_numpy(np.ascontiguousarray(out).astype(np.uint32).view(np.int32)
                                    .reshape(shape_p)).to(dev)
    else:
        labels_d = torch.zeros(shape_p, dtype=torch.int32, device=dev)
    ones = torch.ones(3, dtype=torch.float32, device=dev)
    _run_flood(aff_d, 0, ones, mask_d, seeds_d, labels_d, shape_p, scale)
    if out_is_dev:
        return labels_d
    res = labels_d.cpu().numpy().view(np.uint32)
    if out is not None:
        o = out.reshape(shape_p)
        o[...] = res.astype(o.dtype, copy=False)
        return o
    return res.astype(np.int32)          # _prep_data allocates int32 when out is None (:58-59)


def raveled_affinity_watershed(image_raveled, marker_coords, offsets, mask, output):
    """watershed.py:95-159 on raveled arrays.  `offsets` must be the 6-connected table
    [[0,-YX],[1,-X],[2,-1],[2,1],[1,X],[0,YX]] that _indices_to_raveled_affinities builds
    (watershed.py:84-92); the padded shape is recovered from it.  `output` must already
    hold the seed labels (watershed.py:61-62) and is modified in place and returned."""
    _lib.require_device()
    dev = torch.device('cuda', torch.cuda.current_device())
    offsets = np.asarray(offsets, dtype=np.int64)
    if offsets.shape != (6, 2) or list(offsets[:, 0]) != [0, 1, 2, 2, 1, 0] or \
            list(offsets[:3, 1]) != list(-offsets[::-1][:3, 1]) or offsets[3, 1] != 1:
        raise NotImplementedError('only the 3-D 6-connected offsets table of the reference is supported')
    xp = int(offsets[4, 1])
    yx = int(offsets[5, 1])
    npix = int(np.asarray(image_raveled.shape)[1])
    if xp <= 0 or yx % xp or npix % yx:
        raise ValueError('offsets table inconsistent with the raveled image length')
    shape_p = (npix // yx, yx // xp, xp)
    aff_d = _as_device(image_raveled, torch.float32, dev).view((3,) + shape_p)
    mask_d = _as_device(mask, torch.uint8, dev).view(shape_p)
    seeds = np.asarray(marker_coords, dtype=np.int64).reshape(-1)
    seeds_d = torch.from_numpy(seeds).to(dev)
    out_np = np.asarray(output)
    # Seed voxels carry the caller's labels (watershed.py:61-62); any OTHER voxel that is non-zero
    # on entry is never claimed (the reference tests `output[neighbor_index]`, :150) and keeps its
    # value.  The kernel labels seed i with i+1, so the flood runs on a copy whose seed voxels
    # are cleared, and the caller's seed labels are restored through a LUT afterwards -- only on
    # voxels the flood itself wrote (`claimed`, taken BEFORE the flood modifies `work`).
    labels_d = torch.from_numpy(out_np.astype(np.uint32).view(np.int32).reshape(shape_p)).to(dev)
    work = labels_d.clone()
    if len(seeds):
        work.view(-1)[seeds_d] = 0
    claimable = (work == 0).cpu().numpy().reshape(-1)
    ones = torch.ones(3, dtype=torch.float32, device=dev)
    _run_flood(aff_d, 0, ones, mask_d, seeds_d, work, shape_p, None)
    res = work.cpu().numpy().view(np.uint32).reshape(-1)
    if len(seeds):
        lab_seed = out_np.reshape(-1)[seeds].astype(np.uint32)
        canon = np.arange(1, len(seeds) + 1, dtype=np.uint32)
        if not np.array_equal(lab_seed, canon):
            lut = np.zeros(len(seeds) + 1, dtype=np.uint32)
            lut[1:] = lab_seed
            res = np.where(claimable, lut[np.minimum(res, len(seeds))], res)
    output[...] = res.astype(out_np.dtype, copy=False).reshape(out_np.shape)
    return output
