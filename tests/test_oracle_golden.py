"""The CPU oracle (oracle/) against the golden vectors minted from the verbatim
reference (scripts/make_golden.py).  CPU only."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import chunks, flood, post, unet_ref

SCENES = ['random', 'ties4', 'ties1', 'signed', 'sparse', 'oneseed_full',
          'adjacent_seeds', 'platelets']


@pytest.fixture(scope='module')
def scenes(golden_dir):
    return np.load(os.path.join(golden_dir, 'flood_scenes.npz'))


def test_chunk_grids(golden_dir):
    grids = json.load(open(os.path.join(golden_dir, 'chunk_grids.json')))
    assert len(grids) >= 7
    for key, g in grids.items():
        shape, chunk, margin = json.loads(key)
        st, cr = chunks.make_chunks(tuple(shape), tuple(chunk), margin if isinstance(margin, int) else tuple(margin))
        assert [list(map(int, s)) for s in st] == g['starts'], key
        assert [[list(map(int, ab)) for ab in c] for c in cr] == g['crops'], key


def test_chunk_coverage_exactly_once():
    for shape in [(33, 512, 512), (12, 300, 300), (18, 384, 260)]:
        cover = np.zeros((1,) + shape, np.float32)
        chunks.process_chunks(np.zeros(shape, np.float32), (10, 256, 256), cover, (1, 64, 64),
                              lambda c: np.ones((1,) + c.shape, np.float32) + 0)
        cnt = np.zeros(shape, np.int32)
        st, cr = chunks.make_chunks(shape, (10, 256, 256), (1, 64, 64))
        for s, c in zip(st, cr):
            sl = tuple(slice(s0 + a, s0 + b) for s0, (a, b) in zip(s, c))
            cnt[sl] += 1
        assert cnt.min() == 1 and cnt.max() == 1


def test_chunk_too_small_raises():
    with pytest.raises(ValueError):
        chunks.make_chunks((8, 256, 256), (10, 256, 256), (1, 64, 64))


def test_provenance(golden_dir):
    prov = np.load(os.path.join(golden_dir, 'provenance_12x300x300.npz'))['provenance']
    out = np.zeros((1, 12, 300, 300), np.float32)
    k = {'i': 0}

    def fn(c):
        k['i'] += 1
        return np.full((1,) + c.shape, k['i'], np.float32)

    chunks.process_chunks(np.zeros((12, 300, 300), np.float32), (10, 256, 256), out, (1, 64, 64), fn)
    assert np.array_equal(out[0].astype(np.uint8), prov)


@pytest.mark.parametrize('name', SCENES)
def test_flood_c_oracle(scenes, name):
    aff, seeds, mask, want = (scenes[f'{name}_{k}'] for k in ('aff', 'seeds', 'mask', 'labels'))
    got = flood.affinity_watershed(aff, seeds, mask, impl='c')
    assert got.dtype == np.uint32
    assert np.array_equal(got, want)


@pytest.mark.parametrize('name', ['ties1', 'signed', 'adjacent_seeds'])
def test_flood_py_oracle(scenes, name):
    aff, seeds, mask, want = (scenes[f'{name}_{k}'] for k in ('aff', 'seeds', 'mask', 'labels'))
    got = flood.affinity_watershed(aff, seeds, mask, impl='py')
    assert np.array_equal(got, want)


def test_flood_zero_seeds_is_all_zero():
    aff = np.random.default_rng(0).random((3, 5, 6, 7), dtype=np.float32)
    mask = np.pad(np.ones((3, 4, 5), bool), 1)
    got = flood.affinity_watershed(aff, np.zeros((0, 3), np.int64), mask)
    assert not got.any()


def test_post_stage(golden_dir):
    g = np.load(os.path.join(golden_dir, 'post_small.npz'))
    feats = g['feats'].astype(np.float32)
    out = np.zeros(tuple(s + 2 for s in feats.shape[1:]), np.uint32)
    seg, seeds, mask = post.segment_output_image(feats, out=out.ravel())
    assert np.array_equal(seeds, g['seeds'])
    assert np.array_equal(mask, g['mask'])
    assert np.array_equal(seg, g['seg'])
    assert np.array_equal(out[1:-1, 1:-1, 1:-1], g['seg'])     # written in place through `out`


def test_unet_restatement(golden_dir):
    g = np.load(os.path.join(golden_dir, 'unet_small.npz'))
    sd = unet_ref.synth_state_dict(0)
    y = unet_ref.unet_forward(torch.from_numpy(g['x']), sd).numpy()
    # same torch build -> bit-identical; allow 1e-6 for a different CPU's oneDNN kernels
    assert np.abs(y - g['y']).max() <= 1e-6


def test_state_dict_format():
    sd = unet_ref.synth_state_dict(0)
    assert len(sd) == 134
    n = sum(v.numel() for k, v in sd.items() if v.dtype.is_floating_point and 'running' not in k)
    assert n == 9972673
    assert list(sd)[:4] == ['c0.conv0.weight', 'c0.conv0.bias', 'c0.conv1.weight', 'c0.conv1.bias']
