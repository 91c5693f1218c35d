"""CPU restatement of the DoG blob segmenter.   TEST INFRASTRUCTURE (see oracle/__init__.py).

Follows dog_blob_watershed_for_chunks / dog_image (src/iterseg/segmentation.py:592-650,
:678-680) over scipy.ndimage plus restatements of the three scikit-image functions it calls.
scikit-image is not installed here and there is no network, so these are written from the
published algorithms (scikit-image 0.19-0.22) -- **PARITY UNPINNED**:

* `gaussian(img, sigma)` (mode='nearest', truncate=4): scipy.ndimage.gaussian_filter, float32
  in -> float32 out (oracle/skimage_shim.py).
* `blob_dog(img, min_sigma, max_sigma, threshold)` with the defaults sigma_ratio=1.6,
  overlap=0.5, exclude_border=False: k = int(log(max/min)/log(1.6) + 1) DoG layers of
  mode='reflect' Gaussians scaled by 1/(sigma_ratio-1); peak_local_max over the (z,y,x,scale)
  cube with a 3^4 footprint ('nearest'), value > threshold, a constant cube has no peak, peaks
  ordered by descending value (stable); `_prune_blobs`: pairs closer than 2*sigma*sqrt(3) whose
  sphere overlap exceeds 0.5 lose the blob that is not larger (the first of the pair for equal
  sigmas).  scikit-image walks the pairs in the iteration order of a Python set; here pairs
  are walked in lexicographic (i, j) order.
* `watershed(-distance, markers, mask=mask)`: oracle/flood.c::isg_oracle_node_flood (equal-
  valued markers pop in raveled-index order, see there).  Keys: the squared distance is an
  exact integer and -sqrt is strictly decreasing in it, so the int64 key -d2 orders voxels
  exactly like the float64 -distance.
"""
import ctypes
import math

import numpy as np
from scipy import ndimage as ndi

from . import flood as oflood


def gaussian(img, sigma, mode='nearest'):
    return ndi.gaussian_filter(np.asarray(img, np.float32), sigma, mode=mode, truncate=4.0)


def dog_image(vol, sigma_min, sigma_max):
    """segmentation.py:678-680."""
    return gaussian(vol, sigma_min) - gaussian(vol, sigma_max)


def sigma_list(min_sigma, max_sigma, sigma_ratio=1.6):
    k = int(math.log(float(max_sigma) / float(min_sigma)) / math.log(sigma_ratio) + 1)
    return [float(min_sigma) * sigma_ratio ** i for i in range(k + 1)]


def overlap_distance2(sigma, overlap=0.5, ndim=3):
    """Largest squared integer-lattice distance at which two blobs of this sigma overlap by more
    than `overlap` (spheres of radius sigma*sqrt(ndim), scikit-image's _blob_overlap)."""
    r = float(sigma) * math.sqrt(ndim)
    best = 0
    for d2 in range(1, int((2 * r) ** 2) + 2):
        d = math.sqrt(d2)
        if d > 2 * r:
            break
        vol = math.pi / (12 * d) * (2 * r - d) ** 2 * (d * d + 4 * d * r)
        if vol / (4.0 / 3 * math.pi * r ** 3) > overlap:
            best = d2
    return best


def blob_dog(vol, min_sigma, max_sigma, threshold, sigma_ratio=1.6, overlap=0.5):
    """-> (N, 4) float array (z, y, x, sigma) of the surviving blobs, peak order."""
    vol = np.asarray(vol, np.float32)
    sl = sigma_list(min_sigma, max_sigma, sigma_ratio)
    gs = [gaussian(vol, s, mode='reflect') for s in sl]
    sf = np.float32(1.0 / (sigma_ratio - 1))
    cube = np.stack([(gs[i] - gs[i + 1]) * sf for i in range(len(sl) - 1)], axis=-1)
    mx = ndi.maximum_filter(cube, footprint=np.ones((3,) * cube.ndim, bool), mode='nearest')
    peak = cube == mx
    if peak.all():
        peak[:] = False
    peak &= cube > np.float32(threshold)
    coords = np.argwhere(peak)
    vals = cube[peak]
    order = np.argsort(-vals, kind='stable')
    coords = coords[order]
    blobs = np.concatenate([coords[:, :3].astype(np.float64),
                            np.asarray(sl)[coords[:, 3]][:, None]], axis=1)
    # _prune_blobs
    if len(blobs) > 1:
        smax = blobs[:, 3].max()
        dist = 2 * smax * math.sqrt(3)
        from scipy.spatial import cKDTree
        pairs = sorted(cKDTree(blobs[:, :3]).query_pairs(dist))
        for i, j in pairs:
            b1, b2 = blobs[i], blobs[j]
            if _blob_overlap(b1, b2) > overlap:
                if b1[3] > b2[3]:
                    b2[3] = 0
                else:
                    b1[3] = 0
        blobs = blobs[blobs[:, 3] > 0]
    return blobs


def _blob_overlap(b1, b2, ndim=3):
    if b1[3] == b2[3] == 0:
        return 0.0
    r1, r2 = b1[3] * math.sqrt(ndim), b2[3] * math.sqrt(ndim)
    if r2 > r1:
        r1, r2 = r2, r1
    d = math.sqrt(float(np.sum((b1[:3] - b2[:3]) ** 2)))
    if d > r1 + r2:
        return 0.0
    if d <= abs(r1 - r2):
        return 1.0
    vol = math.pi / (12 * d) * (r1 + r2 - d) ** 2 * (d * d + 2 * d * (r1 + r2) - 3 * (r1 - r2) ** 2)
    return vol / (4.0 / 3 * math.pi * min(r1, r2) ** 3)


def node_flood(keys, markers, mask):
    """keys int64 (smaller pops first), markers int32 (modified copy returned), mask bool."""
    lib = oflood._load()
    lib.isg_oracle_node_flood.restype = ctypes.c_int64
    lib.isg_oracle_node_flood.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64,
                                          ctypes.c_void_p, ctypes.c_void_p]
    shape = keys.shape
    k = np.ascontiguousarray(keys, np.int64).ravel()
    out = np.ascontiguousarray(markers, np.int32).ravel().copy()
    m = np.ascontiguousarray(mask).astype(np.uint8).ravel()
    offs = np.ascontiguousarray(oflood.neighbor_table(shape)[:, 1], np.int64)
    age = lib.isg_oracle_node_flood(k.ctypes.data, k.size, offs.ctypes.data, len(offs), m.ctypes.data,
                                    out.ctypes.data)
    if age < 0:
        raise MemoryError('oracle node flood: heap allocation failed')
    return out.reshape(shape)


def dog_blob_watershed_for_chunks(input_volume, current_output, chunk_size=None, margin=None,
                                  min_sigma=1, max_sigma=1.5, threshold=0.02, **kwargs):
    """segmentation.py:592-650: labels written IN PLACE into the padded `current_output`."""
    vol = np.pad(np.asarray(input_volume, np.float32), 1)
    mask = dog_image(vol, min_sigma, max_sigma) > threshold
    blobs = blob_dog(vol, min_sigma, max_sigma, threshold)
    d2 = np.rint(ndi.distance_transform_edt(vol) ** 2).astype(np.int64)
    centroids = np.zeros(vol.shape, bool)
    if len(blobs):
        idx = tuple(blobs[:, :3].T.astype(int))
        centroids[idx] = True
    markers, n = ndi.label(centroids)
    labels = node_flood(-d2, markers.astype(np.int32), mask)
    current_output[...] = labels.astype(current_output.dtype)
    return {'mask': mask, 'blobs': blobs, 'd2': d2, 'markers': markers}
