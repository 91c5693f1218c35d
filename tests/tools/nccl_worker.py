"""Worker of tests/test_gpu_multi.py: torchrun starts one copy per GPU (NCCL).

Checks, over real NCCL ranks, that
 * a 3-D volume through the public `segmentation_loop` (z-slabs, halo exchange, all-reduced
   statistics, seam label merge) equals rank 0's single-device run bit for bit, and
 * a tzyx series through `segmentation_loop` (frames t = rank mod world, global label offsets,
   all ranks writing ONE zarr store) equals the single-process result.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from iterseg_b200 import _io, segmentation, synth, unet as unet_mod      # noqa: E402

rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
torch.cuda.set_device(int(os.environ['LOCAL_RANK']))
dev = torch.device('cuda', int(os.environ['LOCAL_RANK']))
dist.init_process_group('nccl', device_id=dev)
tmp = sys.argv[1]
net = unet_mod.UNet()
net.load_state_dict(synth.structured_state_dict(0))
net.to(dev)
chunk, margin = (10, 64, 64), (1, 16, 16)
fn = segmentation.affinity_watershed_for_chunks

# ---- one volume, z-slabs ---------------------------------------------------------------------
shape = (64, 128, 128)
vol = synth.platelet_frame(shape, seed=3) * np.float32(0.5) + np.float32(0.01)
store = _io.open_zarr(os.path.join(tmp, 'vol'), shape=shape, chunks=(16, 64, 64), dtype=np.int32)
cfg = {'unet': net, 'output_volume': np.zeros(1), 'slab_halo': 16}
assert list(segmentation.segmentation_loop(None, vol.copy(), chunk, margin, store, fn, cfg)) == [0]
dist.barrier()
if rank == 0:
    want = np.zeros(shape, np.int32)
    single = dict(cfg, shard=False)
    assert list(segmentation.segmentation_loop(None, vol.copy(), chunk, margin, want, fn, single)) == [0]
    got = np.asarray(_io.open_zarr(os.path.join(tmp, 'vol')))
    assert want.max() > 20 and np.array_equal(got, want), (want.max(), got.max())

# ---- one series, frames sharded, one store, global label ids ----------------------------------
T, fshape = 5, (10, 128, 128)
data = np.stack([synth.platelet_frame(fshape, seed=50 + t) for t in range(T)])
store = _io.open_zarr(os.path.join(tmp, 'series'), shape=(T,) + fshape, chunks=(10, 64, 64), dtype=np.int32)
cfg = {'unet': net, 'output_volume': np.zeros(1), 'global_label_offsets': True}
done = list(segmentation.segmentation_loop(None, data, chunk, margin, store, fn, cfg))
assert done == list(range(rank, T, world)), done
dist.barrier()
if rank == 0:
    want = np.zeros((T,) + fshape, np.int32)
    single = dict(cfg, shard=False)
    assert list(segmentation.segmentation_loop(None, data, chunk, margin, want, fn, single)) == list(range(T))
    got = np.asarray(_io.open_zarr(os.path.join(tmp, 'series')))
    assert np.array_equal(got, want), [(int(got[t].max()), int(want[t].max())) for t in range(T)]
    assert want[-1].max() > want[0].max() > 0
dist.barrier()
dist.destroy_process_group()
print(f'rank {rank} ok', flush=True)
