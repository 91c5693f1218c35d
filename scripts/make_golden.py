"""Mint tests/golden/ from the VERBATIM reference modules.

Run in the build container only (needs /root/reference):
    python scripts/make_golden.py
The fixtures are small, committed, and are what the oracle restatement
(oracle/*.py, oracle/flood.c) and the CUDA path are pinned against on machines
where the reference checkout does not exist (the GPU box).
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_harness, unet_ref            # noqa: E402
from iterseg_b200 import synth                      # noqa: E402

GOLD = os.path.join(ROOT, 'tests', 'golden')


def flood_scene(ref, aff, seeds_zyx, mask):
    """Run the reference numba kernel AND its py_func through the reference's
    own affinity_watershed wrapper; they must agree (watershed.py:294)."""
    outs = []
    for py in (False, True):
        out = np.zeros(mask.size, dtype=np.uint32)
        ref.watershed.affinity_watershed(aff.copy(), seeds_zyx.copy(), mask.copy(),
                                         out=out, py_func=py)
        outs.append(out.reshape(mask.shape))
    assert np.array_equal(outs[0], outs[1])
    return outs[0]


def random_scene(rng, shape, p_mask, n_seeds, quant=None, signed=False):
    Z, Y, X = shape
    aff = rng.random((3, Z, Y, X), dtype=np.float32)
    if signed:
        aff = aff - np.float32(0.5)
    if quant:
        aff = (np.round(aff * quant) / quant).astype(np.float32)
    mask = rng.random(shape) < p_mask
    mask = np.pad(mask[1:-1, 1:-1, 1:-1], 1, constant_values=False)
    cand = np.argwhere(mask)
    sel = rng.choice(len(cand), size=min(n_seeds, len(cand)), replace=False)
    return aff, cand[sel].astype(np.int64), mask


def main():
    ref = ref_harness.load()
    os.makedirs(GOLD, exist_ok=True)

    # 1. chunk grids ---------------------------------------------------------
    grids = {}
    for shape, chunk, margin in [((33, 512, 512), (10, 256, 256), (1, 64, 64)),
                                 ((256, 2048, 2048), (10, 256, 256), (1, 64, 64)),
                                 ((12, 300, 300), (10, 256, 256), (1, 64, 64)),
                                 ((10, 256, 256), (10, 256, 256), (1, 64, 64)),
                                 ((18, 384, 260), (10, 256, 256), (1, 64, 64)),
                                 ((40, 100, 90), (8, 32, 32), (2, 4, 6)),
                                 ((16, 64, 64), (8, 32, 32), 0)]:
        st, cr = ref.predict.make_chunks(shape, chunk, margin)
        key = json.dumps([shape, chunk, margin])
        grids[key] = {'starts': [[int(v) for v in s] for s in st],
                      'crops': [[[int(a), int(b)] for a, b in c] for c in cr]}
    with open(os.path.join(GOLD, 'chunk_grids.json'), 'w') as f:
        json.dump(grids, f)

    # 2. process_chunks provenance -------------------------------------------
    vol = np.zeros((12, 300, 300), np.float32)
    outv = np.zeros((1, 12, 300, 300), np.float32)
    counter = {'i': 0}

    def fake(input_volume, sl, **kw):
        counter['i'] += 1
        return np.full((1, 1) + input_volume[sl[1:]].shape, counter['i'], np.float32)

    ref.predict.process_chunks(vol, (10, 256, 256), outv, (1, 64, 64), fake)
    np.savez_compressed(os.path.join(GOLD, 'provenance_12x300x300.npz'),
                        provenance=outv[0].astype(np.uint8))

    # 3. flood scenes ----------------------------------------------------------
    rng = np.random.default_rng(20240)
    scenes = {}
    scenes['random'] = random_scene(rng, (10, 34, 38), 0.7, 40)
    scenes['ties4'] = random_scene(rng, (12, 40, 40), 0.6, 60, quant=4)
    scenes['ties1'] = random_scene(rng, (8, 30, 30), 0.8, 25, quant=1)
    scenes['signed'] = random_scene(rng, (8, 26, 26), 0.7, 20, signed=True)
    scenes['sparse'] = random_scene(rng, (10, 40, 40), 0.3, 80)
    # (zero seeds is not a golden case: the reference itself raises in
    #  np.apply_along_axis, watershed.py:50-52)
    a, s, m = random_scene(rng, (6, 16, 16), 1.0, 1)
    scenes['oneseed_full'] = (a, s, m)
    a, s, m = random_scene(rng, (6, 16, 16), 1.0, 0)
    scenes['adjacent_seeds'] = (a, np.array([[2, 5, 5], [2, 5, 6], [3, 5, 5], [2, 6, 5]], np.int64), m)
    lab = synth.platelet_labels((8, 96, 96), n_objects=40, seed=3)
    feats = synth.analytic_features(lab, 3)
    out = np.zeros((10, 98, 98), np.uint32)
    seg, seeds, mask = ref.watershed.segment_output_image(
        feats.copy(), (0, 1, 2), 4, 3, out=out.ravel())
    affp = feats[[0, 1, 2]].copy()
    affp /= affp.max(axis=(1, 2, 3)).reshape(-1, 1, 1, 1)
    affp = np.pad(affp, ((0, 0), (1, 1), (1, 1), (1, 1)))
    scenes['platelets'] = (affp, (seeds + 1).astype(np.int64), mask)
    pack = {}
    for name, (aff, seeds_, mask_) in scenes.items():
        lab_out = flood_scene(ref, aff, seeds_.reshape(-1, 3), mask_)
        pack[name + '_aff'] = aff
        pack[name + '_seeds'] = seeds_.reshape(-1, 3)
        pack[name + '_mask'] = mask_
        pack[name + '_labels'] = lab_out
    assert np.array_equal(pack['platelets_labels'][1:-1, 1:-1, 1:-1], seg)
    np.savez_compressed(os.path.join(GOLD, 'flood_scenes.npz'), **pack)

    # 4. post-U-Net stage on a small analytic feature map ----------------------
    np.savez_compressed(os.path.join(GOLD, 'post_small.npz'),
                        feats=feats.astype(np.float16), seg=seg, seeds=seeds, mask=mask)
    # NB feats are stored as float16 to keep the fixture small; the golden
    # outputs are recomputed from the float16-rounded features below.
    feats16 = feats.astype(np.float16).astype(np.float32)
    out[:] = 0
    seg, seeds, mask = ref.watershed.segment_output_image(
        feats16.copy(), (0, 1, 2), 4, 3, out=out.ravel())
    np.savez_compressed(os.path.join(GOLD, 'post_small.npz'),
                        feats=feats16.astype(np.float16), seg=seg.copy(), seeds=seeds, mask=mask)

    # 5. U-Net forward on a tiny chunk -----------------------------------------
    sd = unet_ref.synth_state_dict(0)
    net = ref.unet.UNet(in_channels=1, out_channels=5)
    net.load_state_dict(sd)
    x = np.random.default_rng(5).random((1, 1, 4, 32, 32), dtype=np.float32)
    y = net(torch.from_numpy(x)).detach().numpy()      # train-mode BN, as predict.py:118-123
    np.savez_compressed(os.path.join(GOLD, 'unet_small.npz'), x=x, y=y)
    print('golden fixtures written to', GOLD)
    for fn in sorted(os.listdir(GOLD)):
        print(' ', fn, os.path.getsize(os.path.join(GOLD, fn)))


if __name__ == '__main__':
    main()
