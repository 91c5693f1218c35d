"""Spatial slab sharding (BASELINE configs[3]): R virtual ranks on one GPU must reproduce the
single-device result bit for bit (iterseg_b200/slab.py)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def mods():
    from iterseg_b200 import slab, synth, watershed, unet as unet_mod, predict, segmentation
    return dict(slab=slab, synth=synth, ws=watershed, unet=unet_mod, predict=predict, seg=segmentation)


def _features(synth, shape, seed):
    lab = synth.platelet_labels(shape, seed=seed)
    return synth.analytic_features(lab, seed=seed).astype(np.float32), lab


@pytest.mark.parametrize('world,halo', [(2, 24), (3, 20), (4, 16)])
def test_slab_post_stage_equals_single_pass(mods, world, halo):
    """Analytic feature maps (objects straddle every seam): emulated ranks == one pass."""
    slab, synth, ws = mods['slab'], mods['synth'], mods['ws']
    shape = (96, 160, 160)
    feats, lab = _features(synth, shape, seed=5)
    seg, seeds, mask = ws.segment_output_image(feats, (0, 1, 2), 4, 3)
    want = np.asarray(seg).astype(np.uint32)
    assert want.max() > 50
    got, n, info = slab.segment_volume_emulated(None, None, (10, 64, 64), (1, 16, 16), world, halo=halo,
                                                features=feats)
    assert n == len(seeds) == want.max()
    assert np.array_equal(got, want)
    # the seams really cut objects
    sl = slab.plan_slabs(shape, (10, 64, 64), (1, 16, 16), world)[0]
    for s in sl[1:]:
        assert (want[s.z0 - 1] != 0).any() and ((want[s.z0 - 1] == want[s.z0]) & (want[s.z0] != 0)).any()


def test_halo_guard_raises(mods):
    """An object longer than the halo that reaches the own planes must be refused, not approximated."""
    slab, synth = mods['slab'], mods['synth']
    shape = (96, 96, 96)
    feats, lab = _features(synth, shape, seed=7)
    # a column of foreground through all planes: mask channel high, one centre
    feats[3, :, 40:46, 40:46] = 1.0
    feats[0:3, :, 40:46, 40:46] = 0.9
    with pytest.raises(slab.HaloTooSmall):
        slab.segment_volume_emulated(None, None, (10, 64, 64), (1, 16, 16), 3, halo=10, features=feats)


def test_slab_unet_and_post_equals_single_pass(mods):
    """Full path with the network: chunk ownership by slab, global input maximum, halo exchange."""
    slab, synth, unet_mod, predict, ws = mods['slab'], mods['synth'], mods['unet'], mods['predict'], mods['ws']
    shape, chunk, margin = (42, 100, 100), (10, 64, 64), (1, 16, 16)
    vol = synth.platelet_frame(shape, seed=3).astype(np.float32)
    vol = vol * np.float32(0.5) + np.float32(0.01)            # maximum != 1: the normalisation matters
    dev = torch.device('cuda', 0)
    net = unet_mod.UNet()
    net.load_state_dict(synth.structured_state_dict(0))
    net.to(dev)
    frame = torch.from_numpy(vol / np.max(vol)).to(dev)
    feats = predict.predict_frame_device(net, frame, chunk, margin)
    seg, seeds, mask = ws.segment_output_image(feats, (0, 1, 2), 4, 3)
    want = seg.cpu().numpy().view(np.uint32)
    got, n, info = slab.segment_volume_emulated(vol.copy(), net, chunk, margin, 2, halo=16)
    assert n == int(want.max())
    assert np.array_equal(got, want)
