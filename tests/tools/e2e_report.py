"""End-to-end report on the GPU box: GPU pipeline vs the full CPU oracle pipeline."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from iterseg_b200 import segmentation, synth, unet as U, predict, watershed as ws
from oracle import unet_ref, post, metrics
shape = tuple(int(x) for x in sys.argv[1].split(',')) if len(sys.argv) > 1 else (12, 300, 300)
sd = synth.structured_state_dict(0)
net = U.UNet(); net.load_state_dict(sd); net.cuda()
vol, gt = synth.platelet_frame(shape, seed=1, return_labels=True)
cur = np.zeros(tuple(s + 2 for s in shape), np.uint32)
t = time.time()
segmentation.affinity_watershed_for_chunks(vol.copy(), cur, (10, 256, 256), (1, 64, 64), unet=net, output_volume=np.zeros(1))
torch.cuda.synchronize(); print('gpu pipeline (first call)', time.time() - t)
seg_gpu = cur[1:-1, 1:-1, 1:-1]
feats_gpu = predict.predict_frame_device(net, torch.from_numpy(vol).cuda(), (10, 256, 256), (1, 64, 64)).cpu().numpy()
t = time.time(); feats_cpu = unet_ref.predict_frame(vol, sd); print('oracle unet', time.time() - t)
print('feature max abs diff', np.abs(feats_gpu - feats_cpu).max(), 'mean', np.abs(feats_gpu - feats_cpu).mean())
seg_a, seeds_a, mask_a = post.segment_output_image(feats_gpu)
print('plumbing: gpu pipeline == oracle post on gpu feats:', np.array_equal(seg_a, seg_gpu))
seg_o, seeds_o, mask_o = post.segment_output_image(feats_cpu)
print('seeds gpu/oracle', len(seeds_a), len(seeds_o), 'mask diff voxels', int((mask_a != mask_o).sum()), 'of', int(mask_o.sum()))
vi = metrics.variation_of_information(seg_o, seg_gpu)
print('VI(oracle, gpu) =', vi, 'sum', sum(vi), ' F1@0.5 =', metrics.matched_f1(seg_o, seg_gpu), ' label equal frac', float((seg_o == seg_gpu).mean()))
