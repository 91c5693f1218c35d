"""Run-to-run and batched-vs-single differences of the U-Net (GPU box)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from iterseg_b200 import predict, synth, unet as U
net = U.UNet(); net.load_state_dict(synth.structured_state_dict(0)); net.cuda()
vol = torch.from_numpy(synth.platelet_frame((12, 300, 300), seed=5)).cuda()
a = predict.predict_frame_device(net, vol, (10, 256, 256), (1, 64, 64)).clone()
b = predict.predict_frame_device(net, vol, (10, 256, 256), (1, 64, 64)).clone()
print('batched run-to-run max diff', float((a - b).abs().max()))
st, lo, hi = predict._chunk_tables(vol.shape, (10, 256, 256), (1, 64, 64))
c = torch.zeros_like(a)
for i in range(len(st)):
    predict.predict_frame_device(net, vol, (10, 256, 256), (1, 64, 64), out=c, tables=(st[i:i+1], lo[i:i+1], hi[i:i+1]))
print('batched vs one-chunk-at-a-time max diff', float((a - c).abs().max()))
# per-layer comparison for chunk 3 between the N=8 plan and the N=1 plan
names = [f'{m}.conv{i}' for m in ('c0','c1','c2','c3','c4','c5_0','c6_0','c7_0','c8_0') for i in (0,1)]
for name in names:
    x8 = net.debug_conv_output(vol, (10,256,256), st, lo, hi, name, chunk=3)
    x1 = net.debug_conv_output(vol, (10,256,256), st[3:4], lo[3:4], hi[3:4], name, chunk=0)
    d = (x8 - x1).abs()
    print(f'{name:12s} max diff {float(d.max()):.3e}  max|x| {float(x1.abs().max()):.3f}  n_diff {int((d>0).sum())}/{d.numel()}')
