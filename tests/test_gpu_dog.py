"""DoG blob segmenter (BASELINE configs[4]): the CUDA path against the scipy
restatement in oracle/dog.py (parity unpinned w.r.t. scikit-image itself, see there)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _run_gpu(vol, **kw):
    from iterseg_b200 import segmentation
    dev = torch.device('cuda', 0)
    frame = torch.from_numpy(np.ascontiguousarray(vol, np.float32)).to(dev)
    shape_p = tuple(s + 2 for s in vol.shape)
    labels = torch.zeros(shape_p, dtype=torch.int32, device=dev)
    dist = torch.zeros(shape_p, dtype=torch.float64, device=dev)
    mask, counts = segmentation.dog_blob_segment_device(frame, labels, distance=dist, **kw)
    return labels.cpu().numpy(), mask.cpu().numpy().astype(bool), dist.cpu().numpy(), counts.cpu().numpy()


@pytest.mark.parametrize('shape,seed,kw', [
    ((12, 96, 96), 1, {}),
    ((10, 128, 128), 2, {'min_sigma': 1, 'max_sigma': 1.5, 'threshold': 0.01}),
    ((33, 160, 160), 3, {'min_sigma': 1.2, 'max_sigma': 1.8, 'threshold': 0.02}),
])
def test_dog_segment_equals_oracle(shape, seed, kw):
    from iterseg_b200 import synth
    from oracle import dog
    from scipy import ndimage as ndi
    vol = synth.platelet_frame(shape, seed=seed)
    want = np.zeros(tuple(s + 2 for s in shape), np.int32)
    info = dog.dog_blob_watershed_for_chunks(vol, want, **kw)
    lab, mask, dist, counts = _run_gpu(vol, **kw)
    assert np.array_equal(mask, info['mask'])
    assert np.array_equal(dist, ndi.distance_transform_edt(np.pad(vol, 1)))       # bit-exact float64
    assert int(counts[1]) == len(info['blobs'])
    assert np.array_equal(lab, want)
    assert want.max() > 10


def test_dog_zero_voxels_inside_and_plateaus():
    """Zero voxels inside the volume (EDT sources, holes) and flat maxima (the prune rule)."""
    from oracle import dog
    rng = np.random.default_rng(5)
    vol = np.zeros((14, 64, 64), np.float32)
    for _ in range(30):
        z, y, x = rng.integers(2, 12), rng.integers(6, 58), rng.integers(6, 58)
        vol[z - 1:z + 1, y - 3:y + 3, x - 3:x + 3] = np.float32(rng.choice([0.5, 0.75, 1.0]))     # flat tops: ties
    want = np.zeros(tuple(s + 2 for s in vol.shape), np.int32)
    info = dog.dog_blob_watershed_for_chunks(vol, want)
    lab, mask, dist, counts = _run_gpu(vol)
    assert np.array_equal(mask, info['mask'])
    assert np.array_equal(dist ** 2, info['d2'].astype(np.float64)) or np.allclose(dist ** 2, info['d2'])
    assert int(counts[1]) == len(info['blobs'])
    assert np.array_equal(lab, want)


def test_dog_through_the_plugin(tmp_path):
    """segment_data(..., segmenter='DoG-blob-watershed') on a short series == the oracle per frame."""
    from iterseg_b200 import _dock_widgets, synth, viewer
    from oracle import dog
    shape = (10, 96, 96)
    data = np.stack([synth.platelet_frame(shape, seed=s) for s in (7, 8)])
    v = viewer.HeadlessViewer()
    layer = viewer.Image(data, name='img', scale=(1, 4, 1, 1), translate=(0, 0, 0, 0))
    out_layer = _dock_widgets.segment_data(v, layer, save_dir=None, name='dog', segmenter='DoG-blob-watershed',
                                           network_or_config_file=None, debug=True)
    labels = np.asarray(out_layer.data)
    for t in range(2):
        vol = data[t].astype(np.float32)
        vol = vol / np.max(vol)
        want = np.zeros(tuple(s + 2 for s in shape), np.int32)
        dog.dog_blob_watershed_for_chunks(vol, want)
        assert np.array_equal(labels[t], want[1:-1, 1:-1, 1:-1])
        assert labels[t].max() > 5


@pytest.mark.parametrize('shape,seed,kw', [
    ((12, 96, 96), 11, {'min_sigma': 1, 'max_sigma': 2, 'threshold': 0.02}),          # 2 DoG layers
    ((16, 128, 128), 12, {'min_sigma': 1, 'max_sigma': 3, 'threshold': 0.01}),        # 3 layers, radius up to 16
    ((33, 160, 160), 13, {'min_sigma': 0.8, 'max_sigma': 4, 'threshold': 0.02}),      # 4 layers
])
def test_dog_multi_layer_equals_oracle(shape, seed, kw):
    """blob_dog with several DoG layers (max_sigma / min_sigma >= 1.6; segmentation.py:637-638 leaves the
    sigmas to the config file): 3^4 maxima over (z, y, x, layer) and _prune_blobs with per-blob sigma,
    against the scipy restatement (parity unpinned w.r.t. scikit-image: the pair order of the prune
    and the heap tie rule are ours, oracle/dog.py)."""
    from iterseg_b200 import synth
    from oracle import dog
    vol = synth.platelet_frame(shape, seed=seed)
    want = np.zeros(tuple(s + 2 for s in shape), np.int32)
    info = dog.dog_blob_watershed_for_chunks(vol, want, **kw)
    lab, mask, dist, counts = _run_gpu(vol, **kw)
    assert len(dog.sigma_list(kw['min_sigma'], kw['max_sigma'])) >= 3
    assert np.array_equal(mask, info['mask'])
    assert int(counts[1]) == len(info['blobs'])
    assert len(np.unique(info['blobs'][:, 3])) >= 2                # blobs of different scales survive
    assert np.array_equal(lab, want)
    assert want.max() > 10


def test_dog_too_many_layers_are_refused():
    from iterseg_b200 import segmentation
    with pytest.raises(NotImplementedError):
        segmentation._dog_params(0.5, 60.0, 0.02)
    with pytest.raises(ValueError):
        segmentation._dog_params(1.0, 20.0, 0.02)       # a layer's Gaussian radius exceeds the kernel's limit
