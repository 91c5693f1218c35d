"""The post-U-Net stage alone on one 33x512x512 frame of network features (GPU box): three timed
repetitions of isg_segment_features, for `ncu` captures of the seeds / mask / components / flood kernels."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from iterseg_b200 import predict, synth, unet as U, watershed as ws      # noqa: E402

net = U.UNet()
net.load_state_dict(synth.structured_state_dict(0))
net.cuda()
shape = (33, 512, 512)
vol = torch.from_numpy(synth.platelet_frame(shape, seed=0)).cuda()
feats = predict.predict_frame_device(net, vol, (10, 256, 256), (1, 64, 64))
labels = torch.zeros(tuple(s + 2 for s in shape), dtype=torch.int32, device='cuda')
ws.segment_features_device(feats, labels)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    labels.zero_()
    seeds, counts, mask, otsu = ws.segment_features_device(feats, labels)
e1.record()
torch.cuda.synchronize()
c = counts.cpu().numpy()
print(f'post stage: {e0.elapsed_time(e1) / 3:.2f} ms; seeds {c[0]}, components {c[2]}, multi-seed {c[3]}')
