"""Slab-sharded segmentation of one large synthetic volume (BASELINE.json configs[3]).

    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/slab_volume.py \
        [--shape 128 1024 1024] [--halo 24] [--check]

Every rank builds the same synthetic volume (a 32x512x512 platelet block tiled with per-tile
gain), takes its z-slab, and the ranks produce globally numbered labels
(iterseg_b200/slab.py).  Rank 0 prints one JSON line: voxels/s (max over ranks, CUDA events
around U-Net + halo exchange + statistics all-reduces + post stage + label merge; the host
volume is resident, the H2D copy of the own planes is inside).  --check: rank 0 also runs the
single-device pass and asserts bit-identical labels.
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from iterseg_b200 import predict, slab, synth, unet as unet_mod, watershed     # noqa: E402


def make_volume(shape, seed=1000):
    base = synth.platelet_frame((32, 512, 512), seed=seed).astype(np.float32)
    reps = [(s + b - 1) // b for s, b in zip(shape, base.shape)]
    rng = np.random.default_rng(seed)
    vol = np.empty(tuple(r * b for r, b in zip(reps, base.shape)), np.float32)
    for k in range(reps[0]):
        for j in range(reps[1]):
            for i in range(reps[2]):
                g = np.float32(rng.uniform(0.8, 1.0))
                vol[k * 32:(k + 1) * 32, j * 512:(j + 1) * 512, i * 512:(i + 1) * 512] = \
                    np.roll(base, (int(rng.integers(32)), int(rng.integers(512)), int(rng.integers(512))),
                            (0, 1, 2)) * g
    return np.ascontiguousarray(vol[:shape[0], :shape[1], :shape[2]])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--shape', type=int, nargs=3, default=[128, 1024, 1024])
    ap.add_argument('--halo', type=int, default=24)
    ap.add_argument('--check', action='store_true')
    ap.add_argument('--reps', type=int, default=2)
    a = ap.parse_args()
    rank, world = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
    local = int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist.init_process_group('nccl', device_id=dev)
    shape, chunk, margin = tuple(a.shape), (10, 256, 256), (1, 64, 64)
    vol = make_volume(shape)
    net = unet_mod.UNet()
    net.load_state_dict(synth.structured_state_dict(0))
    net.to(dev)
    times = []
    for rep in range(a.reps + 1):                       # first repetition = warm-up
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        own, (z0, z1), n = slab.segment_volume_slabs(vol, net, chunk, margin, halo=a.halo)
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rep:
            times.append(float(t.item()))
    ok = None
    if a.check:
        parts = [None] * world if rank == 0 else None
        dist.gather_object((z0, z1, own.cpu().numpy().view(np.uint32)), parts, dst=0)
        if rank == 0:
            got = np.zeros(shape, np.uint32)
            for p0, p1, lab in parts:
                got[p0:p1] = lab
            frame = torch.from_numpy(vol / np.max(vol)).to(dev)
            feats = predict.predict_frame_device(net, frame, chunk, margin)
            seg, seeds, mask = watershed.segment_output_image(feats, (0, 1, 2), 4, 3)
            want = seg.cpu().numpy().view(np.uint32)
            ok = bool(np.array_equal(got, want)) and int(want.max()) == n
    if rank == 0:
        ms = float(np.median(times))
        print(json.dumps({'workload': f'configs[3]-style: one {shape[0]}x{shape[1]}x{shape[2]} volume, z-slabs with halo '
                                      f'{a.halo}, global statistics all-reduced, seam label merge',
                          'n_gpus': world, 'ms': ms, 'voxels_per_s': float(np.prod(shape)) / ms * 1e3,
                          'labels': n, 'identical_to_single_device': ok}))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
