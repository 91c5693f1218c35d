"""Throughput of the DoG blob segmenter (BASELINE.json configs[4]) on synthetic platelet frames.

    python scripts/dog_bench.py [--steps K]          (one GPU)
    torchrun --nproc-per-node N scripts/dog_bench.py   (frame-wise sharding, one rank per GPU)

Rank 0 prints one JSON line: voxels/s device-resident (CUDA events, max over ranks) and end to end
through `segmentation.dog_blob_watershed_for_chunks` with pinned host buffers, plus the
single-thread CPU restatement (oracle/dog.py) on the same frame.
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from iterseg_b200 import segmentation, synth          # noqa: E402

FRAME = (33, 512, 512)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--no-cpu', action='store_true')
    a = ap.parse_args()
    rank, world = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
    local = int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=dev)
    vol = synth.platelet_frame(FRAME, seed=rank)
    frame = torch.from_numpy(vol).to(dev)
    shape_p = tuple(s + 2 for s in FRAME)
    labels = torch.zeros(shape_p, dtype=torch.int32, device=dev)
    cfg = dict(min_sigma=1, max_sigma=1.5, threshold=0.02)

    def step():
        labels.zero_()
        return segmentation.dog_blob_segment_device(frame, labels, **cfg)

    for _ in range(a.warmup):
        mask, counts = step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        mask, counts = step()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / a.steps], dtype=torch.float64, device=dev)
    vol_pinned = torch.from_numpy(vol.copy()).pin_memory()
    out_pinned = torch.zeros(shape_p, dtype=torch.int32).pin_memory()
    for _ in range(2):
        segmentation.dog_blob_watershed_for_chunks(vol_pinned.numpy(), out_pinned.numpy().view(np.uint32), None, None, **cfg)
    t0 = time.perf_counter()
    for _ in range(a.steps):
        segmentation.dog_blob_watershed_for_chunks(vol_pinned.numpy(), out_pinned.numpy().view(np.uint32), None, None, **cfg)
    torch.cuda.synchronize()
    te = torch.tensor([(time.perf_counter() - t0) / a.steps * 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    if rank == 0:
        nvox = float(np.prod(FRAME))
        c = counts.cpu().numpy()
        line = {'workload': 'configs[4]: DoG blob watershed (min_sigma 1, max_sigma 1.5, threshold 0.02), one synthetic '
                            '33x512x512 frame per rank and step',
                'n_gpus': world, 'ms_per_step': float(t.item()), 'voxels_per_s': nvox * world / (float(t.item()) * 1e-3),
                'e2e_voxels_per_s': nvox * world / (float(te.item()) * 1e-3),
                'blobs': int(c[1]), 'labels': int(labels.max().item()), 'mask_fraction': float(mask.float().mean().item())}
        if not a.no_cpu and world == 1:
            from oracle import dog
            out = np.zeros(shape_p, np.int32)
            t0 = time.perf_counter()
            dog.dog_blob_watershed_for_chunks(vol, out, **cfg)
            dt = time.perf_counter() - t0
            line['cpu_baseline'] = {'value': nvox / dt, 'unit': 'voxels/s', 'cores': 1, 'kind': 'port',
                                    'sample': 'one whole frame, scipy.ndimage + C heap flood'}
            line['identical_to_cpu_restatement'] = bool(np.array_equal(out, labels.cpu().numpy()))
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
