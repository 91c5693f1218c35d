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
print(f'frame U-Net: {ms:.2f} ms  -> {plan.flops / ms / 1e9:.1f} TFLOP/s  ({np.prod(shape) / ms / 1e3:.1f} Mvox/s)')
