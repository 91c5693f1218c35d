"""scipy/numpy restatement of the scikit-image calls on the hot path.

TEST INFRASTRUCTURE (see oracle/__init__.py).  scikit-image is an un-vendored,
unpinned dependency of the reference (setup.cfg:40) and is not installed in this
image, so the six functions the path calls are restated here from their
published algorithms (scikit-image 0.19 - 0.25 behaviour).
PARITY UNPINNED: these cannot be checked against scikit-image itself here.

Call sites in the reference:
  filters.gaussian          watershed.py:227,234   (mode='nearest', truncate=4)
  filters.threshold_otsu    watershed.py:227
  feature.peak_local_max    watershed.py:235
  morphology.remove_small_objects            watershed.py:241-246
  morphology._util._validate_connectivity    watershed.py:47
  morphology._util._offsets_to_raveled_neighbors   watershed.py:85
"""
import numpy as np
from scipy import ndimage as ndi


# ---------------------------------------------------------------- filters
def gaussian(image, sigma=1, *, mode='nearest', cval=0, preserve_range=False,
             truncate=4.0, channel_axis=None, output=None):
    """skimage.filters.gaussian: float32 stays float32, scipy separable filter,
    one 1-D pass per axis with sigma > 0, float32 stored between passes."""
    img = np.asarray(image)
    if img.dtype not in (np.float32, np.float64):
        img = img.astype(np.float64)
    if np.isscalar(sigma):
        sigma = (float(sigma),) * img.ndim
    return ndi.gaussian_filter(img, sigma, mode=mode, cval=cval,
                               truncate=truncate)


def threshold_otsu(image=None, nbins=256):
    """skimage.filters.threshold_otsu (nbins=256) on a float image."""
    image = np.asarray(image)
    first = image.reshape(-1)[0]
    if np.all(image == first):
        return first
    counts, edges = np.histogram(image.reshape(-1), nbins)
    centers = (edges[:-1] + edges[1:]) / 2
    counts = counts.astype('float32', copy=False)
    w1 = np.cumsum(counts)
    w2 = np.cumsum(counts[::-1])[::-1]
    m1 = np.cumsum(counts * centers) / w1
    m2 = (np.cumsum((counts * centers)[::-1]) / w2[::-1])[::-1]
    var12 = w1[:-1] * w2[1:] * (m1[:-1] - m2[1:]) ** 2
    return centers[np.argmax(var12)]


# ---------------------------------------------------------------- feature
def peak_local_max(image, min_distance=1, threshold_abs=None,
                   threshold_rel=None, exclude_border=True,
                   num_peaks=np.inf, footprint=None, labels=None,
                   num_peaks_per_label=np.inf, p_norm=np.inf):
    """skimage.feature.peak_local_max for the defaults the path uses
    (min_distance=1, exclude_border=True, no labels, unlimited peaks).

    mask = (image == maximum_filter(image, 3^ndim, 'nearest')) & (image > thr),
    all-True plateau -> no peaks, borders of width min_distance cleared,
    coordinates in C order then stably sorted by descending intensity.
    ensure_spacing(spacing=1) only rejects points closer than 1 -> no-op on a grid.
    """
    assert labels is None and footprint is None and np.isinf(num_peaks)
    image = np.asarray(image)
    thr = threshold_abs if threshold_abs is not None else image.min()
    if threshold_rel is not None:
        thr = max(thr, threshold_rel * image.max())
    size = 2 * min_distance + 1
    image_max = ndi.maximum_filter(image, size=size, mode='nearest')
    out = image == image_max
    if np.all(out):
        out[:] = False
    out &= image > thr
    if exclude_border:
        b = min_distance if exclude_border is True else int(exclude_border)
        if b > 0:
            for ax in range(out.ndim):
                sl = [slice(None)] * out.ndim
                sl[ax] = slice(None, b)
                out[tuple(sl)] = False
                sl[ax] = slice(-b, None)
                out[tuple(sl)] = False
    coord = np.nonzero(out)
    inten = image[coord]
    order = np.argsort(-inten, kind='stable')
    return np.transpose(coord)[order]


# ------------------------------------------------------------- morphology
def remove_small_objects(ar, min_size=64, connectivity=1, *, out=None):
    """skimage.morphology.remove_small_objects on an integer label image:
    labels whose voxel count is < min_size are zeroed."""
    ar = np.asarray(ar)
    res = ar.copy()
    if min_size == 0:
        return res
    assert res.dtype != bool, 'shim: label-image form only'
    sizes = np.bincount(res.ravel())
    res[(sizes < min_size)[res]] = 0
    return res


def _validate_connectivity(image_dim, connectivity, offset):
    if connectivity is None:
        connectivity = 1
    if np.isscalar(connectivity):
        c_connectivity = ndi.generate_binary_structure(image_dim, connectivity)
    else:
        c_connectivity = np.array(connectivity, bool)
    if offset is None:
        offset = np.array(c_connectivity.shape) // 2
    return c_connectivity, np.asarray(offset)


def _offsets_to_raveled_neighbors(image_shape, footprint, center, order='C'):
    """Raveled offsets of the footprint's non-centre voxels, sorted (stably) by
    Euclidean distance from the centre, C-order strides of image_shape."""
    footprint = np.asarray(footprint, bool)
    idx = np.stack(np.nonzero(footprint), axis=-1)
    offs = idx - np.asarray(center)
    strides = np.cumprod((1,) + tuple(image_shape[::-1][:-1]))[::-1]
    rav = (offs * strides).sum(axis=1)
    dist = (offs ** 2).sum(axis=1)
    keep = dist > 0
    rav, dist = rav[keep], dist[keep]
    return rav[np.argsort(dist, kind='stable')]
