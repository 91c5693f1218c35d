"""Feature map -> labels, restated on numpy/scipy.  TEST INFRASTRUCTURE.

Restates segment_output_image and helpers (src/iterseg/watershed.py:165-251):
  :192-201  squeeze, gather affinity channels, per-channel max normalise, pad 1
  :203-205  seeds = peak_local_max(gaussian(centre, (0,1,1)), 0.04) + 1
  :207-213  mask = raw mask channel > otsu(gaussian(mask channel, 2)), pad 0
  :214-216  drop 6-connected components with < 10 or >= 10 000 000 voxels and
            the seeds outside kept components (:239-251)
  :218-223  affinity watershed into `out`; crop
The scikit-image calls go through oracle/skimage_shim.py: PARITY UNPINNED for
those (see oracle/__init__.py); the control flow is pinned by golden vectors made
with the verbatim reference module on top of the same shim.
"""
import numpy as np
from scipy import ndimage as ndi

from . import skimage_shim as sk
from .flood import affinity_watershed


def normalise_pad_affinities(unet_output, channels=(0, 1, 2)):
    aff = np.array(unet_output[list(channels)], dtype=np.float32, copy=True)
    aff /= aff.max(axis=(1, 2, 3)).reshape(-1, 1, 1, 1)
    return np.pad(aff, ((0, 0), (1, 1), (1, 1), (1, 1)))


def get_centroids(cent):
    return sk.peak_local_max(sk.gaussian(cent, sigma=(0, 1, 1)), threshold_abs=.04)


def otsu_threshold_of_smoothed(img, sigma=2):
    return sk.threshold_otsu(sk.gaussian(img, sigma=sigma))


def get_mask(img, sigma=2):
    return img > otsu_threshold_of_smoothed(img, sigma)


def remove_unwanted_objects(mask, centroids, min_area=10, max_area=10000000):
    labels, _ = ndi.label(mask)
    sizes = np.bincount(labels.ravel())
    keep = (sizes >= min_area) & (sizes < max_area)
    keep[0] = False
    new_mask = keep[labels]
    sel = new_mask[tuple(np.asarray(centroids).T)] if len(centroids) else np.zeros(0, bool)
    return new_mask, np.asarray(centroids)[sel]


def segment_output_image(unet_output, affinities_channels=(0, 1, 2), centroids_channel=4,
                         thresholding_channel=3, scale=None, absolute_thresh=None,
                         out=None, impl='c'):
    unet_output = np.asarray(np.squeeze(unet_output))
    aff = normalise_pad_affinities(unet_output, affinities_channels)
    cents = get_centroids(unet_output[centroids_channel]) + 1
    mimg = unet_output[thresholding_channel]
    mask = get_mask(mimg) if absolute_thresh is None else mimg > absolute_thresh
    mask = np.pad(mask, 1, constant_values=False)
    mask, cents = remove_unwanted_objects(mask, cents, 10, 10000000)
    seg = affinity_watershed(aff, cents, mask, scale=scale, out=out, impl=impl)
    return seg[1:-1, 1:-1, 1:-1], cents - 1, mask


def segment_single_volume(vol, unet_forward_frame):
    """segment_single_volume + affinity_watershed_for_chunks
    (segmentation.py:885-900,147-195) for strictly positive input:
    vol /= max; features = chunked U-Net; labels = segment_output_image."""
    vol = np.asarray(vol, dtype=np.float32)
    vol = vol / vol.max()
    feats = unet_forward_frame(vol)
    out = np.zeros(tuple(s + 2 for s in vol.shape), dtype=np.uint32)
    segment_output_image(feats, out=out.ravel())
    return out[1:-1, 1:-1, 1:-1]
