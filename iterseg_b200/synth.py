"""Seeded synthetic platelet volumes and analytic feature maps.

There is no network for datasets and the reference's bundled network file is
missing, so benchmarks and parity tests run on synthetic data:

* `platelet_labels`  -- oblate ellipsoids ("platelets"), rejection-placed so
  that each keeps >= 80 % of its volume (objects may touch and form small
  aggregates).
* `platelet_frame`   -- a fluorescence-like zyx float32 frame rendered from the
  labels (PSF blur, noise, strictly positive so that the reference's
  remove_sum_zero_slices branch, segmentation.py:887-888, is never taken).
* `analytic_features` -- the 5 channels the U-Net would predict, derived from
  the ground truth with the reference's target conventions: affinities
  aff[a][v] = 1 where labels differ between v-1 and v along axis a
  (labels.py:87-109), mask = labels > 0, centreness peaking at each centroid
  (labels.py:143-205).

Pure numpy/scipy host code: data generation is not on the measured path.
"""
import numpy as np
from scipy import ndimage as ndi


def platelet_labels(shape=(33, 512, 512), n_objects=None, seed=0, max_tries=None):
    rng = np.random.default_rng(seed)
    Z, Y, X = shape
    if n_objects is None:
        n_objects = int(round(1500 * (Z * Y * X) / (33 * 512 * 512)))
    labels = np.zeros(shape, dtype=np.int32)
    placed, tries = 0, 0
    max_tries = max_tries or 20 * n_objects + 100
    while placed < n_objects and tries < max_tries:
        tries += 1
        rz = rng.uniform(1.0, 2.0)
        ra, rb = rng.uniform(4.0, 8.0, 2)
        th = rng.uniform(0, np.pi)
        cz, cy, cx = rng.uniform(0, Z), rng.uniform(0, Y), rng.uniform(0, X)
        R = int(np.ceil(max(ra, rb))) + 1
        Rz = int(np.ceil(rz)) + 1
        z0, z1 = max(0, int(cz) - Rz), min(Z, int(cz) + Rz + 1)
        y0, y1 = max(0, int(cy) - R), min(Y, int(cy) + R + 1)
        x0, x1 = max(0, int(cx) - R), min(X, int(cx) + R + 1)
        if z0 >= z1 or y0 >= y1 or x0 >= x1:
            continue
        zz, yy, xx = np.meshgrid(np.arange(z0, z1) - cz, np.arange(y0, y1) - cy,
                                 np.arange(x0, x1) - cx, indexing='ij')
        u = np.cos(th) * yy + np.sin(th) * xx
        v = -np.sin(th) * yy + np.cos(th) * xx
        inside = (zz / rz) ** 2 + (u / ra) ** 2 + (v / rb) ** 2 <= 1.0
        vol = int(inside.sum())
        if vol < 20:
            continue
        sub = labels[z0:z1, y0:y1, x0:x1]
        free = inside & (sub == 0)
        if free.sum() < 0.8 * vol:
            continue
        placed += 1
        sub[free] = placed
    return labels


def platelet_frame(shape=(33, 512, 512), seed=0, n_objects=None, return_labels=False):
    labels = platelet_labels(shape, n_objects, seed)
    rng = np.random.default_rng(seed + 7919)
    n = int(labels.max())
    inten = np.concatenate([[0.0], rng.uniform(0.4, 1.0, n)]).astype(np.float32)
    vol = inten[labels]
    vol = ndi.gaussian_filter(vol, (0.7, 1.5, 1.5), mode='nearest')
    vol = vol + rng.normal(0.0, 0.03, shape).astype(np.float32) + np.float32(0.02)
    vol = np.clip(vol, 1e-3, None).astype(np.float32)
    vol /= vol.max()
    return (vol, labels) if return_labels else vol


def timeseries(n_frames, shape=(33, 512, 512), seed0=0):
    for t in range(n_frames):
        yield platelet_frame(shape, seed0 + t)


def analytic_features(labels, seed=0, noise=0.01):
    """(5,Z,Y,X) float32: z/y/x affinities, mask, centreness."""
    rng = np.random.default_rng(seed + 104729)
    shape = labels.shape
    feats = np.zeros((5,) + shape, dtype=np.float32)
    for a in range(3):
        pad = [(0, 0)] * 3
        pad[a] = (1, 0)
        lp = np.pad(labels, pad, mode='reflect')
        sl0 = [slice(None)] * 3
        sl1 = [slice(None)] * 3
        sl0[a] = slice(0, shape[a])
        sl1[a] = slice(1, shape[a] + 1)
        aff = (lp[tuple(sl0)] != lp[tuple(sl1)]).astype(np.float32)
        feats[a] = ndi.gaussian_filter(aff, 0.5, mode='nearest')
    feats[3] = ndi.gaussian_filter((labels > 0).astype(np.float32), 0.7, mode='nearest')
    n = int(labels.max())
    if n > 0:
        idx = np.arange(1, n + 1)
        com = np.array(ndi.center_of_mass(np.ones(shape, np.float32), labels, idx))
        zz, yy, xx = np.nonzero(labels)
        lab = labels[zz, yy, xx]
        c = com[lab - 1]
        d = np.sqrt((4.0 * (zz - c[:, 0])) ** 2 + (yy - c[:, 1]) ** 2 + (xx - c[:, 2]) ** 2)
        feats[4][zz, yy, xx] = (1.0 / (1.0 + 0.35 * d)).astype(np.float32)
    if noise:
        feats += rng.normal(0.0, noise, feats.shape).astype(np.float32)
    np.clip(feats, 1e-4, 1.0, out=feats)
    return feats
