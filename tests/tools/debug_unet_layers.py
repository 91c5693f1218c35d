"""Per-layer parity report of the CUDA U-Net against the fp32 oracle (GPU box)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from iterseg_b200 import unet as U            # noqa: E402
from oracle import unet_ref                   # noqa: E402


def report(chunk_shape, seed=0, mode=None):
    if mode is not None:
        os.environ['ISG_CONV_BASE_OFFSET'] = str(mode)
    if '--structured' in sys.argv:
        from iterseg_b200 import synth
        sd = synth.structured_state_dict(0)
    else:
        sd = unet_ref.synth_state_dict(0)
    net = U.UNet()
    net.load_state_dict(sd)
    net.cuda()
    rng = np.random.default_rng(seed)
    x = rng.random((1, 1) + chunk_shape, dtype=np.float32)
    if '--structured' in sys.argv:
        from iterseg_b200 import synth
        x = synth.platelet_frame(chunk_shape, seed=1)[None, None]
    ref = {}
    y_ref = unet_ref.unet_forward(torch.from_numpy(x), sd, hook=lambda k, v: ref.__setitem__(k, v.clone()))
    frame = torch.from_numpy(x[0, 0]).cuda()
    zeros = np.zeros((1, 3), np.int32)
    hi = np.asarray([chunk_shape], np.int32)
    names = [f'{m}.conv{i}' for m in ('c0', 'c1', 'c2', 'c3', 'c4', 'c5_0', 'c6_0', 'c7_0', 'c8_0') for i in (0, 1)]
    print(f'--- chunk {chunk_shape} base_offset_mode={os.environ.get("ISG_CONV_BASE_OFFSET", "1")}')
    for name in names:
        got = net.debug_conv_output(frame, chunk_shape, zeros, zeros, hi, name).cpu()
        bias = sd[name + '.bias'].view(-1, 1, 1, 1)
        want = ref[name][0] - bias
        err = (got - want).abs()
        print(f'{name:12s} shape {tuple(got.shape)} max|ref| {float(want.abs().max()):8.4f} '
              f'max err {float(err.max()):9.5f} mean err {float(err.mean()):9.6f} '
              f'nan {int(torch.isnan(got).sum())}')
    y = net(torch.from_numpy(x)).cpu()
    d = (y - y_ref).abs()
    print(f'final: max abs {float(d.max()):.5f} mean {float(d.mean()):.6f}')
    return float(d.max())


if __name__ == '__main__':
    shapes = [(4, 32, 32)]
    if '--full' in sys.argv:
        shapes.append((10, 256, 256))
    for m in ((0, 1) if '--modes' in sys.argv else (0,)):
        for s in shapes:
            try:
                report(s, mode=m)
            except Exception as e:       # keep going: this is a diagnosis tool
                print('FAILED', s, m, repr(e))
