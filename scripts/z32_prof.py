"""One U-Net forward with the instrumented round-1 z-ring kernel: wait-time counters of the producer / MMA / epilogue
roles of two CTAs, printed by the kernel (profiles/r02_notes.md, "what a change of the D tile costs").

    ISG_NVCC_EXTRA=-DISG_Z32_PROF python -m iterseg_b200._build --force
    ISG_Z32_MODE=ring [ISG_Z32_EPI=4|8] python scripts/z32_prof.py        # on the GPU box
    python -m iterseg_b200._build --force                                  # back to the shipped library
"""
import os
import sys

os.environ.setdefault('ISG_Z32_MODE', 'ring')     # the counters live in conv3d_zring32_kernel

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
for _ in range(3):
    predict.predict_frame_device(net, vol, (10, 256, 256), (1, 64, 64), out=out)
    torch.cuda.synchronize()
