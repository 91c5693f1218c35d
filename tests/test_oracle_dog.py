"""CPU checks of the DoG restatement (oracle/dog.py): internal consistency only -- scikit-image is
not available, so nothing here pins it against the real library (parity unpinned)."""
import math

import numpy as np
from scipy import ndimage as ndi


def test_prune_rule_equals_sequential_pair_walk():
    """With one common sigma, walking the overlapping pairs in lexicographic order and zeroing the
    first blob of each overlapping live pair is the same as: blob i dies iff a LATER blob lies
    within the overlap distance (the data-parallel rule the CUDA kernel uses)."""
    from oracle import dog
    rng = np.random.default_rng(0)
    for sigma in (1.0, 1.6, 2.0):
        d2max = dog.overlap_distance2(sigma)
        pts = np.unique(rng.integers(0, 14, size=(300, 3)), axis=0)
        rng.shuffle(pts)
        blobs = np.concatenate([pts.astype(float), np.full((len(pts), 1), sigma)], axis=1)
        # sequential restatement (as in oracle.dog.blob_dog)
        from scipy.spatial import cKDTree
        seq = blobs.copy()
        for i, j in sorted(cKDTree(seq[:, :3]).query_pairs(2 * sigma * math.sqrt(3))):
            if dog._blob_overlap(seq[i], seq[j]) > 0.5:
                if seq[i][3] > seq[j][3]:
                    seq[j][3] = 0
                else:
                    seq[i][3] = 0
        alive_seq = seq[:, 3] > 0
        # parallel rule
        d2 = ((pts[:, None, :] - pts[None, :, :]) ** 2).sum(-1)
        later = np.triu(np.ones((len(pts), len(pts)), bool), 1)
        alive_par = ~((d2 <= d2max) & (d2 > 0) & later).any(1)
        assert np.array_equal(alive_seq, alive_par)


def test_integer_key_orders_like_float_distance():
    """-sqrt(d2) in float64 and the integer key -d2 order voxels identically, ties included."""
    d2 = np.arange(0, 3 * 520 ** 2, 7, dtype=np.int64)
    dist = -np.sqrt(d2.astype(np.float64))
    assert np.all(np.diff(dist) < 0)            # strictly decreasing: no two integers share a sqrt


def test_node_flood_small_case():
    from oracle import dog
    keys = np.array([[[5, 4, 3, 4, 5]]], np.int64)          # a valley at x = 2
    markers = np.zeros((1, 1, 5), np.int32)
    markers[0, 0, 0], markers[0, 0, 4] = 1, 2
    mask = np.ones((1, 1, 5), bool)
    out = dog.node_flood(keys, markers, mask)
    # equal-valued markers pop in index order: marker 1 pops first and claims x=1 (key 4), which
    # pops before marker 2 (key 5) and runs down the valley and up to x=3 (key 4 < 5)
    assert out.tolist() == [[[1, 1, 1, 1, 2]]]


def test_dog_oracle_runs_and_labels_every_marker():
    from iterseg_b200 import synth
    from oracle import dog
    vol = synth.platelet_frame((10, 96, 96), seed=4)
    out = np.zeros((12, 98, 98), np.int32)
    info = dog.dog_blob_watershed_for_chunks(vol, out)
    n = int(info['markers'].max())
    assert n == out.max() and n > 10
    assert set(np.unique(out[info['markers'] > 0])) == set(range(1, n + 1))
    assert not out[~(info['mask'] | (info['markers'] > 0))].any()
    assert np.array_equal(info['d2'], np.rint(ndi.distance_transform_edt(np.pad(vol, 1)) ** 2).astype(np.int64))


def test_dog_params_multi_layer_match_the_restatement():
    """The host side of multi-layer blob_dog (no device): sigma list, Gaussian radii and weights handed to the
    C-ABI are those of the scipy restatement (scikit-image: sigma_list = min_sigma * 1.6**i, truncate 4)."""
    from iterseg_b200 import segmentation, watershed as ws
    from oracle import dog
    for mn, mx in ((1.0, 1.5), (1.0, 2.0), (1.0, 3.0), (0.8, 4.0)):
        p = segmentation._dog_params(mn, mx, 0.02)
        sl = dog.sigma_list(mn, mx)
        assert p.n_layers == len(sl) - 1
        if p.n_layers > 1:
            assert [p.layer_sigma[i] for i in range(len(sl))] == sl
            for i, s in enumerate(sl):
                w, r = ws.gaussian_half_kernel(s)
                assert p.layer_radius[i] == r == int(4.0 * s + 0.5)
                assert [p.layer_weights[i][j] for j in range(r + 1)] == [float(x) for x in w]
            assert p.mask_radius[1] == int(4.0 * mx + 0.5)
        assert abs(p.scale_factor - 1.0 / 0.6) < 1e-6 and p.overlap == 0.5
