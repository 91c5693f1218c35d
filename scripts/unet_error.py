"""Max-abs error of the device U-Net against a float64-free fp32 torch CPU forward on one full chunk
(GPU box; the fp32 forward here is the Python mirror's own state_dict run through torch.nn.functional,
NOT the oracle -- scripts may not import oracle/).  Prints the margin to the 1e-2 gate."""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from iterseg_b200 import synth, unet as U      # noqa: E402


def conv_module(x, sd, name, last=False):
    for i in (0, 1):
        x = F.conv3d(x, sd[f'{name}.conv{i}.weight'], sd[f'{name}.conv{i}.bias'], padding=1)
        x = F.batch_norm(x, None, None, sd[f'{name}.batch{i}.weight'], sd[f'{name}.batch{i}.bias'], True, 0.1, 1e-5)
        x = torch.sigmoid(x) if (last and i == 1) else F.relu(x)
    return x


def forward(x, sd):
    skips = []
    pools = [((1, 2, 2), (0, 1, 1))] * 3 + [((2, 2, 2), (0, 1, 1))]
    for l in range(4):
        x = conv_module(x, sd, f'c{l}')
        skips.append(x)
        x = F.max_pool3d(x, pools[l][0], pools[l][0], pools[l][1])
    x = conv_module(x, sd, 'c4')
    for u in range(4):
        w = sd[f'up{u}.weight']
        st = tuple(w.shape[2:])
        x = F.conv_transpose3d(x, w, sd[f'up{u}.bias'], stride=st, groups=w.shape[0])
        s = skips[3 - u]
        if u == 3:
            x = x[:, :, :, 1:-1, 1:-1]
        x = x[:, :, :s.shape[2], :s.shape[3], :s.shape[4]]
        x = torch.cat([x, s], 1)
        x = conv_module(x, sd, f'c{5 + u}_0', last=(u == 3))
    return x


if __name__ == '__main__':
    torch.set_num_threads(os.cpu_count())
    for sdname, sd in (('structured', synth.structured_state_dict(0)),):
        net = U.UNet()
        net.load_state_dict(sd)
        net.cuda()
        vol = synth.platelet_frame((10, 256, 256), seed=2)
        x = torch.from_numpy(vol[None, None])
        with torch.no_grad():
            want = forward(x, {k: v.float() for k, v in sd.items()}).numpy()
        got = net(x).cpu().numpy()
        err = np.abs(got - want)
        print(f'{sdname}: max-abs {err.max():.3e}  mean {err.mean():.3e}  (gate 1e-2)')
