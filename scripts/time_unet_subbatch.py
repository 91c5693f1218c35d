"""Experiment: the U-Net of a frame in sub-batches of B chunks (B = 36: one batch, the default) --
does keeping a sub-batch's activations L2-resident (a (10,256,256) chunk's level-0 tensor is 42 MB fp16,
L2 is ~126 MB) beat the launch overhead of many small launches?  (GPU box)"""
import os
import sys

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
st, lo, hi = predict._chunk_tables(shape, (10, 256, 256), (1, 64, 64))
ref = None
for B in (36, 18, 12, 6, 4, 3, 2, 1):
    def run():
        for b in range(0, len(st), B):
            sl = slice(b, b + B)
            net.forward_chunks(vol, (10, 256, 256), st[sl], lo[sl], hi[sl], out=out)
    for _ in range(2):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(4):
        run()
    e1.record()
    torch.cuda.synchronize()
    if ref is None:
        ref = out.clone()
    print(f'sub-batch {B:2d} chunks: {e0.elapsed_time(e1) / 4:.2f} ms per frame, identical {bool(torch.equal(out, ref))}', flush=True)
