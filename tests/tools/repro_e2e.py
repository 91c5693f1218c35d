"""Diagnosis: GPU features vs the fp32 oracle on the 12x300x300 structured-network frame."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from iterseg_b200 import predict, synth, unet, segmentation
from oracle import unet_ref, post, metrics
CHUNK, MARGIN = (10, 256, 256), (1, 64, 64)
frame = synth.platelet_frame((12, 300, 300), seed=1)
sd = synth.structured_state_dict(0)
net = unet.UNet(); net.load_state_dict(sd); net.cuda()
fo = unet_ref.predict_frame(frame, sd)
fg = predict.predict_frame_device(net, torch.from_numpy(frame).cuda(), CHUNK, MARGIN).cpu().numpy()
d = np.abs(fo - fg)
print('feature max abs diff per channel', d.reshape(5, -1).max(1), 'mean', d.mean())
so, seeds_o, mask_o = post.segment_output_image(fo)
sg, seeds_g, mask_g = post.segment_output_image(fg)
print('oracle post: seeds', len(seeds_o), len(seeds_g), 'labels', so.max(), sg.max(), 'mask', mask_o.sum(), mask_g.sum())
print('VI', sum(metrics.variation_of_information(so, sg)), 'F1', metrics.matched_f1(so, sg, 0.5))
cur = np.zeros(tuple(s + 2 for s in frame.shape), np.uint32)
segmentation.affinity_watershed_for_chunks(frame.copy(), cur, CHUNK, MARGIN, unet=net, output_volume=np.zeros((5,) + frame.shape, np.float32))
print('gpu path labels', cur.max(), 'equal to oracle post on gpu feats', np.array_equal(cur[1:-1, 1:-1, 1:-1], sg))
