"""Time every kernel of one U-Net forward over the 36 chunks of a 33x512x512 frame (GPU box)."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from iterseg_b200 import predict, synth, unet as U      # noqa: E402

net = U.UNet()
net.load_state_dict(synth.structured_state_dict(0))
net.cuda()
shape = (33, 512, 512)
vol = torch.from_numpy(synth.platelet_frame(shape, seed=0)).cuda()
out = torch.zeros((5,) + shape, device='cuda')
for _ in range(2):
    predict.predict_frame_device(net, vol, (10, 256, 256), (1, 64, 64), out=out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 5
e0.record()
for _ in range(reps):
    predict.predict_frame_device(net, vol, (10, 256, 256), (1, 64, 64), out=out)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
plan = list(net._plans.values())[0]
# per-launch CUDA-event times (each launch bracketed by its own pair of events)
import ctypes
from iterseg_b200 import _lib
lib = _lib.load()
lib.isg_unet_plan_profile(plan.ptr, 2)
for _ in range(reps):
    predict.predict_frame_device(net, vol, (10, 256, 256), (1, 64, 64), out=out)
torch.cuda.synchronize()
buf = (ctypes.c_double * 1024)()
kinds = (ctypes.c_int * 1024)()
n = lib.isg_unet_plan_profile_launches(plan.ptr, buf, kinds, 1024)
lib.isg_unet_plan_profile(plan.ptr, 0)
NAMES = ['conv_in', 'bn0', 'c0.conv1', 'pool0', 'c1.conv0', 'bn1', 'c1.conv1', 'pool1', 'c2.conv0', 'bn2', 'c2.conv1',
         'pool2', 'c3.conv0', 'bn3', 'c3.conv1', 'pool3', 'c4.conv0', 'bn4', 'c4.conv1', 'up0', 'c5.conv0', 'bn5',
         'c5.conv1', 'up1', 'c6.conv0', 'bn6', 'c6.conv1', 'up2', 'c7.conv0', 'bn7', 'c7.conv1', 'up3', 'c8.conv0',
         'conv_out', 'place']
per = n // reps
if per == len(NAMES):
    t = np.array(buf[:n]).reshape(reps, per).mean(0) * 1e3
    print(' | '.join(f'{nm} {v:.0f}' for nm, v in zip(NAMES, t)))
    k = np.array(kinds[:per])
    print(f'sum {t.sum():.0f} us: tcgen05 TMA convs {t[k == 0].sum():.0f}, other {t[k != 0].sum():.0f}')
else:
    print('unexpected launch count', n, per)
print(f'frame U-Net: {ms:.2f} ms  -> {plan.flops / ms / 1e9:.1f} TFLOP/s  ({np.prod(shape) / ms / 1e3:.1f} Mvox/s)')
