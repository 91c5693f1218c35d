"""Frame-wise sharding of tzyx series over the GPUs of one box.

The reference is single-process / single-device; frames
are independent units there (segmentation.py:873-882: per-frame normalisation
:889, per-frame labels watershed.py:61), so the natural partition is one
process per GPU, each taking whole frames -- no data-path collective.  The one
exchange step is bookkeeping: making label ids unique across the series needs
the per-frame label counts of all ranks (an all-gather of T int64 values over
NCCL), an exclusive prefix sum, and an offset added to every non-zero label.
That is an ADDITION (the reference restarts labels at 1 in every frame), so it
is opt-in.
"""
import numpy as np
import torch

from . import _lib


def shard_frames(n_frames, rank, world_size):
    """Frames t = rank (mod world_size): consecutive frames land on different GPUs, every rank
    walks the series in time order, and step s of all ranks covers the frames s*R .. s*R+R-1
    (what makes a running global label offset one all-gather per step)."""
    return list(range(int(rank), int(n_frames), int(world_size)))


def world(group=None):
    """(rank, world_size) of the initialised torch.distributed job, else (0, 1)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


class LabelOffsets:
    """Running global label offset of a frame-sharded series, kept on the device.

    Step s: every rank contributes the label count of its frame s*R + rank (0 when it has
    none); one all-gather of R int64 (NCCL over NVLink; gloo on CPU tensors) gives the exclusive
    prefix inside the step, a device-resident total carries the earlier steps.  Nothing is
    read back to the host: `step` returns a 1-element int64 tensor that the crop kernel
    (isg_crop_labels) adds to the non-zero labels of this rank's frame.  All ranks must call
    `step` the same number of times (ceil(T / R))."""

    def __init__(self, rank, world_size, device, group=None):
        self.rank, self.world, self.group = int(rank), int(world_size), group
        self.total = torch.zeros(1, dtype=torch.int64, device=device)
        self.gathered = torch.zeros(self.world, dtype=torch.int64, device=device)

    def step(self, count):
        """count: 1-element int64 tensor on the device (or None: no frame in this step)."""
        c = count.reshape(1).to(torch.int64) if count is not None else torch.zeros_like(self.total)
        if self.world > 1:
            import torch.distributed as dist
            dist.all_gather_into_tensor(self.gathered, c.contiguous(), group=self.group)
            off = self.total + self.gathered[:self.rank].sum()
            self.total = self.total + self.gathered.sum()
        else:
            off = self.total.clone()
            self.total = self.total + c
        return off


def exclusive_offsets(counts):
    counts = np.asarray(counts, dtype=np.int64)
    off = np.zeros_like(counts)
    if len(counts) > 1:
        off[1:] = np.cumsum(counts[:-1])
    return off


def gather_label_counts(local_counts, n_frames, rank, world_size, device=None, group=None):
    """local_counts: {frame index: number of labels in that frame} for this rank's frames.
    Returns the dense (n_frames,) int64 count vector on every rank (torch.distributed
    all_gather; NCCL when `device` is a CUDA device, gloo on the CPU)."""
    import torch.distributed as dist
    mine = shard_frames(n_frames, rank, world_size)
    per_rank = (n_frames + world_size - 1) // world_size
    buf = torch.zeros(per_rank, dtype=torch.int64, device=device)
    for i, t in enumerate(mine):
        buf[i] = int(local_counts.get(t, 0))
    if world_size == 1 or not (dist.is_available() and dist.is_initialized()):
        gathered = [buf]
    else:
        gathered = [torch.empty_like(buf) for _ in range(world_size)]
        dist.all_gather(gathered, buf, group=group)
    counts = np.zeros(n_frames, dtype=np.int64)
    for r, g in enumerate(gathered):
        g = g.cpu().numpy()
        for i, t in enumerate(shard_frames(n_frames, r, world_size)):
            counts[t] = g[i]
    return counts


def add_label_offset_(labels, offset):
    """labels: int32 CUDA tensor (uint32 bit pattern), in place: non-zero labels += offset."""
    if offset == 0:
        return labels
    lib = _lib.load()
    assert labels.is_cuda and labels.dtype == torch.int32 and labels.is_contiguous()
    with torch.cuda.device(labels.device):
        rc = lib.isg_add_label_offset(labels.data_ptr(), labels.numel(), int(offset), _lib.stream_ptr())
    _lib.check(rc, 'isg_add_label_offset')
    return labels


def add_label_offset_host(labels, offset):
    """numpy variant used for stores that are already on the host."""
    if offset:
        np.add(labels, np.asarray(offset, dtype=labels.dtype), out=labels, where=labels != 0)
    return labels
