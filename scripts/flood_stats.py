"""Component statistics of the flood stage on the bench frame (GPU): sizes and seed counts of
the mask components, to size the shared-memory classes of the ordered flood."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from scipy import ndimage as ndi
from iterseg_b200 import predict, synth, unet as unet_mod, watershed as ws

FRAME, CHUNK, MARGIN = (33, 512, 512), (10, 256, 256), (1, 64, 64)
dev = torch.device('cuda', 0)
vol, lab = synth.platelet_frame(FRAME, seed=0, return_labels=True)
net = unet_mod.UNet(); net.load_state_dict(synth.structured_state_dict(0)); net.to(dev)
frame = torch.from_numpy(vol).to(dev)
feats = torch.zeros((5,) + FRAME, dtype=torch.float32, device=dev)
labels = torch.zeros(tuple(s + 2 for s in FRAME), dtype=torch.int32, device=dev)
predict.predict_frame_device(net, frame, CHUNK, MARGIN, out=feats)
seeds, counts, mask, otsu = ws.segment_features_device(feats, labels)
torch.cuda.synchronize()
n_seeds = int(counts[0].item())
m = mask.cpu().numpy().astype(bool)
s = seeds.cpu().numpy()[:n_seeds]
cc, n = ndi.label(m)
sizes = np.bincount(cc.ravel())[1:]
seed_comp = cc.ravel()[s]
nseed = np.bincount(seed_comp, minlength=n + 1)[1:]
multi = nseed >= 2
print('components', n, 'multi', multi.sum(), 'seeds', n_seeds, 'mask voxels', m.sum())
order = np.argsort(-sizes)
print('largest components (size, seeds):', [(int(sizes[i]), int(nseed[i])) for i in order[:12]])
ms = np.sort(sizes[multi])[::-1]
print('multi sizes quantiles', np.percentile(ms, [50, 90, 99]).tolist(), 'sum', int(ms.sum()))

# optional phase profile (library built with -DFLOOD_PROF)
import ctypes
from iterseg_b200 import _lib
lib = _lib.load()
if hasattr(lib, 'isg_debug_flood_prof'):
    buf = (ctypes.c_ulonglong * 16)()
    lib.isg_debug_flood_prof(buf, 1)
    labels.zero_()
    ws.segment_features_device(feats, labels)
    torch.cuda.synchronize()
    lib.isg_debug_flood_prof(buf, 0)
    v = list(buf)
    pops = max(v[0], 1)
    print(f'bucket-queue flood, comps > 5000 nodes: pops {v[0]}  clk/pop {v[6]/pops:.0f}; per pop: refills {v[1]/pops:.3f} '
          f'front pushes {v[2]/pops:.2f} back pushes {v[3]/pops:.2f} evictions {v[4]/pops:.4f} swept {v[5]/pops:.3f}')
