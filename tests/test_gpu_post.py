"""GPU parity: flood / CCL / seeds / mask through the C-ABI vs the oracle and the
golden vectors.  Bit-exact (integer / index work)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SCENES = ['random', 'ties4', 'ties1', 'signed', 'sparse', 'oneseed_full',
          'adjacent_seeds', 'platelets']


@pytest.fixture(scope='module')
def ws():
    from iterseg_b200 import _lib, watershed
    _lib.require_device()
    return watershed


@pytest.fixture(scope='module')
def scenes(golden_dir):
    return np.load(os.path.join(golden_dir, 'flood_scenes.npz'))


@pytest.mark.parametrize('name', SCENES)
def test_flood_golden(ws, scenes, name):
    aff, seeds, mask, want = (scenes[f'{name}_{k}'] for k in ('aff', 'seeds', 'mask', 'labels'))
    out = np.zeros(mask.size, np.uint32)
    got = ws.affinity_watershed(aff, seeds, mask, out=out)
    assert np.array_equal(got, want)
    assert np.array_equal(out.reshape(mask.shape), want)      # in place through `out`


def _random_scene(rng, shape, p_mask, n_seeds, quant=None):
    aff = rng.random((3,) + shape, dtype=np.float32)
    if quant:
        aff = (np.round(aff * quant) / quant).astype(np.float32)
    mask = rng.random(shape) < p_mask
    mask = np.pad(mask[1:-1, 1:-1, 1:-1], 1, constant_values=False)
    cand = np.argwhere(mask)
    sel = rng.choice(len(cand), size=min(n_seeds, len(cand)), replace=False)
    return aff, cand[sel].astype(np.int64), mask


@pytest.mark.parametrize('shape,p,n,quant', [
    ((12, 64, 64), 0.55, 200, None),      # many components
    ((12, 64, 64), 0.9, 50, 8),           # one big component, heavy ties -> global-arena heap
    ((20, 130, 130), 0.45, 3000, 4),      # near percolation, many multi-seed components
    ((6, 300, 300), 0.8, 20, 2),          # big component, few seeds
    ((35, 130, 140), 0.35, 5000, None),
])
def test_flood_random_vs_oracle(ws, shape, p, n, quant):
    from oracle import flood as oflood
    rng = np.random.default_rng(hash((shape, n)) & 0xFFFF)
    aff, seeds, mask = _random_scene(rng, shape, p, n, quant)
    want = oflood.affinity_watershed(aff, seeds, mask)
    got = ws.affinity_watershed(aff, seeds, mask, out=np.zeros(mask.size, np.uint32))
    assert np.array_equal(got, want)


@pytest.mark.parametrize('quant', [None, 4])
def test_flood_size_classes(ws, quant):
    """Disjoint boxes sized to hit every flood size class (bucket-queue S / M / L and the
    heap-kernel XL path), several seeds each, a duplicated seed and a seed outside the mask."""
    from oracle import flood as oflood
    rng = np.random.default_rng(17 if quant else 18)
    shape = (30, 120, 120)
    aff = rng.random((3,) + shape, dtype=np.float32)
    if quant:
        aff = (np.round(aff * quant) / quant).astype(np.float32)
    mask = np.zeros(shape, bool)
    seeds = []
    boxes = [(1, 1, 1, 5), (1, 10, 1, 5), (1, 20, 1, 7), (1, 30, 1, 10), (1, 45, 1, 11),
             (1, 60, 1, 16), (1, 80, 1, 18), (1, 1, 40, 20), (1, 30, 40, 24), (1, 60, 40, 3),
             (1, 64, 70, 26)]
    for z0, y0, x0, e in boxes:
        mask[z0:z0 + e, y0:y0 + e, x0:x0 + e] = True
        k = 1 if e == 3 else int(rng.integers(2, 9))
        pts = np.stack([rng.integers(z0, z0 + e, k), rng.integers(y0, y0 + e, k),
                        rng.integers(x0, x0 + e, k)], 1)
        seeds.append(pts)
    seeds = np.concatenate(seeds).astype(np.int64)
    seeds = np.concatenate([seeds, seeds[3:4], np.array([[1, 6, 3]])])   # duplicate + outside-mask seed
    assert not mask[1, 6, 3]
    want = oflood.affinity_watershed(aff, seeds, mask)
    got = ws.affinity_watershed(aff, seeds, mask, out=np.zeros(mask.size, np.uint32))
    assert np.array_equal(got, want)


def test_flood_scale_and_no_mask(ws):
    from oracle import flood as oflood
    rng = np.random.default_rng(3)
    aff, seeds, mask = _random_scene(rng, (8, 40, 40), 1.0, 30, 4)
    sc = np.array([4.0, -1.0, 0.5], np.float32)
    want = oflood.affinity_watershed(aff, seeds, None, scale=sc)
    got = ws.affinity_watershed(aff, seeds, None, scale=sc)
    assert np.array_equal(got.astype(np.uint32), want)


def test_flood_zero_seeds(ws):
    rng = np.random.default_rng(4)
    aff, _, mask = _random_scene(rng, (6, 20, 20), 0.7, 0)
    got = ws.affinity_watershed(aff, np.zeros((0, 3), np.int64), mask)
    assert not got.any()


def test_raveled_entry_point(ws, scenes):
    from oracle import flood as oflood
    aff, seeds, mask, want = (scenes[f'ties4_{k}'] for k in ('aff', 'seeds', 'mask', 'labels'))
    shape = mask.shape
    raveled = np.stack([a.ravel() for a in aff])
    flat = oflood.ravel_seeds(seeds, shape)
    output = np.zeros(mask.size, np.uint32)
    output[flat] = np.arange(1, len(flat) + 1)
    ws.raveled_affinity_watershed(raveled, flat, oflood.neighbor_table(shape), mask.ravel(), output)
    assert np.array_equal(output.reshape(shape), want)


def test_raveled_entry_point_keeps_caller_labels(ws, scenes):
    """the seeds carry whatever labels the caller put into `output` (watershed.py:61-62
    is the caller's business) and pre-labelled non-seed voxels are barriers that keep their value
    -- also when that value collides with 1..N."""
    from oracle import flood as oflood
    aff, seeds, mask = (scenes[f'ties4_{k}'] for k in ('aff', 'seeds', 'mask'))
    shape = mask.shape
    raveled = np.stack([a.ravel() for a in aff])
    flat = oflood.ravel_seeds(seeds, shape)
    table = oflood.neighbor_table(shape)
    rng = np.random.default_rng(5)
    output = np.zeros(mask.size, np.uint32)
    output[flat] = rng.permutation(len(flat)).astype(np.uint32) * 3 + 7        # not 1..N, not sorted
    free = np.flatnonzero(mask.ravel() & (output == 0))
    barrier = rng.choice(free, size=max(4, len(free) // 50), replace=False)
    output[barrier] = rng.integers(1, len(flat) + 1, len(barrier)).astype(np.uint32)   # collide with 1..N
    want = output.copy()
    oflood.raveled_flood_c(raveled, flat, table, mask.ravel(), want)
    got = output.copy()
    ws.raveled_affinity_watershed(raveled, flat, table, mask.ravel(), got)
    assert np.array_equal(got, want)
    assert np.array_equal(got[barrier], output[barrier])


def test_segment_output_image_golden(ws, golden_dir):
    g = np.load(os.path.join(golden_dir, 'post_small.npz'))
    feats = g['feats'].astype(np.float32)
    out = np.zeros(tuple(s + 2 for s in feats.shape[1:]), np.uint32)
    seg, seeds, mask = ws.segment_output_image(feats, (0, 1, 2), 4, 3, out=out.ravel())
    assert np.array_equal(mask, g['mask'])
    assert np.array_equal(seeds, g['seeds'])
    assert np.array_equal(seg, g['seg'])
    assert np.array_equal(out[1:-1, 1:-1, 1:-1], g['seg'])


def test_segment_output_image_absolute_thresh(ws, golden_dir):
    from oracle import post as opost
    g = np.load(os.path.join(golden_dir, 'post_small.npz'))
    feats = g['feats'].astype(np.float32)
    seg, seeds, mask = ws.segment_output_image(feats, (0, 1, 2), 4, 3, absolute_thresh=0.4)
    seg_o, seeds_o, mask_o = opost.segment_output_image(feats, absolute_thresh=0.4)
    assert np.array_equal(mask, mask_o) and np.array_equal(seeds, seeds_o)
    assert np.array_equal(seg, seg_o)


def test_segment_output_image_full_frame(ws):
    """BASELINE configs[0..1] size: one 33x512x512 frame of analytic features."""
    from iterseg_b200 import synth
    from oracle import post as opost
    lab = synth.platelet_labels((33, 512, 512), seed=0)
    feats = synth.analytic_features(lab, 0)
    out = np.zeros((35, 514, 514), np.uint32)
    seg, seeds, mask = ws.segment_output_image(feats, (0, 1, 2), 4, 3, out=out.ravel())
    want = np.zeros((35, 514, 514), np.uint32)
    seg_o, seeds_o, mask_o = opost.segment_output_image(feats, out=want.ravel())
    assert np.array_equal(mask, mask_o)
    assert np.array_equal(seeds, seeds_o)
    assert np.array_equal(out, want)
    # size-independent properties: labels only inside the mask, every seed keeps its label,
    # running it again is idempotent
    assert not out[~mask].any()
    assert np.array_equal(out[tuple((seeds + 1).T)], np.arange(1, len(seeds) + 1))
    out2 = np.zeros_like(out)
    ws.segment_output_image(feats, (0, 1, 2), 4, 3, out=out2.ravel())
    assert np.array_equal(out, out2)


def test_noise_features_percolating(ws):
    """Unstructured features: speckle mask, one percolating component."""
    from oracle import post as opost
    rng = np.random.default_rng(9)
    feats = rng.random((5, 12, 96, 96), dtype=np.float32)
    seg, seeds, mask = ws.segment_output_image(feats, (0, 1, 2), 4, 3)
    seg_o, seeds_o, mask_o = opost.segment_output_image(feats)
    assert np.array_equal(mask, mask_o) and np.array_equal(seeds, seeds_o)
    assert np.array_equal(seg, seg_o)


def test_constant_centre_map_has_no_seeds(ws):
    feats = np.random.default_rng(1).random((5, 6, 40, 40), dtype=np.float32)
    feats[4] = 0.5
    seg, seeds, mask = ws.segment_output_image(feats, (0, 1, 2), 4, 3)
    assert len(seeds) == 0 and not np.asarray(seg).any()


def test_gpu_metrics_equal_oracle():
    """VI and IoU-matched TP/FP/FN on the device == the oracle's scipy restatement (metrics.py)."""
    from iterseg_b200 import metrics, synth
    from oracle import metrics as om
    gt = synth.platelet_labels((16, 128, 128), seed=2).astype(np.uint32)
    rng = np.random.default_rng(0)
    seg = gt.copy()
    ids = np.unique(gt)[1:]
    seg[np.isin(gt, ids[:5])] = 0                               # false negatives
    seg[np.isin(gt, ids[5:9])] = ids[5]                         # a merge
    seg[2:4, 5:9, 5:9] = gt.max() + 7                           # a false positive
    perm = rng.permutation(int(seg.max()) + 1).astype(np.uint32); perm[0] = 0
    seg = perm[seg]                                             # label permutation must not matter
    m = metrics.label_metrics(gt, seg)
    want_vi = om.variation_of_information(gt, seg)
    assert np.allclose([m['vi_seg_given_gt'], m['vi_gt_given_seg']], want_vi, rtol=1e-9, atol=1e-12)
    assert (m['tp'], m['fp'], m['fn']) == om.matched_counts(gt, seg)
    assert m['f1'] == pytest.approx(om.matched_f1(gt, seg))
    same = metrics.label_metrics(gt, gt)
    assert same['f1'] == 1.0 and same['vi_seg_given_gt'] == 0.0 and same['vi_gt_given_seg'] == 0.0
    with pytest.raises(Exception):
        metrics.matched_counts(gt, seg, 0.3)
